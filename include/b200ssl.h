/*
 * b200ssl — C ABI of the B200-native (sm_100a) hot path of the GipMed self-supervised ViT:
 * the DINO-style training step (ViT encoder, projection head, centred cross-entropy, teacher EMA,
 * fused optimiser). Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * Conventions
 *  - every function returns 0 on success, a negative code otherwise; b200ssl_last_error() then holds a
 *    human-readable message (thread-local). Nothing throws across the boundary.
 *  - all pointers are DEVICE pointers owned by the caller (16-byte aligned unless stated); the library
 *    allocates nothing persistent and never synchronises: kernels are enqueued on `stream`
 *    (a cudaStream_t passed as void*), so every call is CUDA-graph capturable.
 *  - activations are bf16 row-major [rows, features]; parameters / statistics / gradients of
 *    parameters are fp32. "ld*" are leading dimensions in ELEMENTS.
 *  - VT.pyc@Lnn = source line nn of the reference's nn_encoder_arch/__pycache__/
 *    vision_transformer.cpython-37.pyc (the reference ships this path as bytecode only);
 *    train.py:nn = /root/reference/train.py line nn.
 */
#ifndef B200SSL_H_
#define B200SSL_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------- */
int b200ssl_version(void);
const char* b200ssl_last_error(void);
/* 0 iff the current CUDA device is compute capability 10.x (there is no fallback path). */
int b200ssl_device_check(void);

/* ---- K3: tcgen05 GEMM  D[M,N] = epilogue(A . B^T), bf16 x bf16 -> fp32 accumulate in TMEM -------
 * Replaces every nn.Linear fprop / dgrad / wgrad cuBLASLt call of the encoder and head
 * (qkv VT.pyc@L114,121; proj @L116,129; fc1/fc2 @L93-95,99-102; DINOHead @L303-318,327-329) and the
 * patch-embed Conv2d-as-GEMM (@L165,169).
 *   a_mn_major / b_mn_major: 0 = operand stored [M|N, K] row-major (K-major), 1 = stored [K, M|N].
 *   epilogue: 0 D=acc+bias | 1 D=gelu'(acc+bias), D2=gelu(acc+bias) (exact erf) | 2 D=acc+bias+aux(bf16)
 *             3 D=acc*aux(bf16) | 4 D(fp32)+=acc split-K atomics, and bias (if non-null) is the OUTPUT
 *             db[M] += column sums of A (wgrad bias gradient) | 5 D(fp32)=acc+bias+aux(fp32) (residual stream; with
 *             D2 != NULL, D2 is a fp32 per-row scale s[M] and D = s[row]*(acc+bias)+aux: stochastic depth)
 *             6 like 4 but stored TRANSPOSED: D(fp32)[n, m] += acc[m, n] (ldd >= M) and bias is the OUTPUT
 *             db[N] += column sums of B (wgrad of a layer with more inputs than outputs; needs block_n 384)
 *             7 D=gelu(acc+bias) only (no-grad forward, e.g. the teacher: nothing saved for backward)
 *   split_k: K splits for epilogues 4/6 (0 = auto); block_n: 0 = auto, else 64/128/192/256/384 dividing N
 *   (384 = 256 x 384 CTA-pair tiles: M > 128, epilogues 0, 4, 5, 6).
 * Constraints: N % 64 == 0, lda/ldb % 8 == 0. */
int b200ssl_gemm(const void* A, long long lda, int a_mn_major, const void* B, long long ldb, int b_mn_major,
                 void* D, long long ldd, void* D2, const float* bias, const void* aux, long long ldaux,
                 int M, int N, int K, int epilogue, int split_k, int block_n, void* stream);

/* LayerNorm fused into the GEMM that consumes it (Block.norm1 -> attn.qkv VT.pyc@L147,121; Block.norm2 -> mlp.fc1
 * @L151,99): D = epilogue(LN(x) W^T + bias) with x the fp32 residual stream [M, 384], W bf16 [N, 384] K-major.
 * ln_out (bf16 [M,384] contiguous), mean, rstd (fp32 [M]) are optional side outputs for backward (NULL: skipped,
 * e.g. the no-grad teacher). epilogue 0 (bias), 1 (GELU' -> D, GELU -> D2) or 7 (GELU only). M > 128, K == 384,
 * N % 192 == 0 or N % 256 == 0. */
int b200ssl_ln_gemm(const float* x, long long ldx, const float* gamma, const float* beta, float eps, void* ln_out,
                    float* mean, float* rstd, const void* W, long long ldw, void* D, long long ldd, void* D2,
                    const float* bias, int M, int N, int K, int epilogue, void* stream);
/* Residual GEMM with a LayerNorm TAIL: D(fp32) = rowscale * (A W^T + bias) + aux, and in the same kernel
 * ln_out(bf16) = LayerNorm(D; gamma, beta, eps) plus the row statistics (mean / rstd nullable, both or none). A cluster
 * computes both 192-column halves of its 256 rows back to back and its epilogue warps normalise the finished rows out of
 * L2: attn.proj + Block.norm2 and mlp.fc2 + the following Block.norm1 (VT.pyc@L147-151) without a second pass over the
 * fp32 stream. N must be 384 (ViT-S), M > 128. rowscale / bias nullable. */
int b200ssl_gemm_res_ln(const void* A, long long lda, const void* W, long long ldw, float* D, long long ldd,
                        const float* rowscale, const float* bias, const float* aux, long long ldaux, int M, int N, int K,
                        const float* gamma, const float* beta, float eps, void* ln_out, float* mean, float* rstd,
                        void* stream);


/* 1 = independent CTAs; 2 (default) = CTA pairs (tcgen05 cta_group::2: 256-row tiles, each CTA loads half of B). */
int b200ssl_set_gemm_cluster(int ctas);
/* 1 (default) = for K <= 384 keep the B tile stationary in shared memory and stream only A; 0 = always stream both. */
int b200ssl_set_gemm_stationary(int on);
/* 1 (default) = choose 256 x 384 CTA-pair tiles automatically (wgrad; long-K plain / fp32-residual epilogues). */
int b200ssl_set_gemm_wide(int on);
/* 1 = GEMM / LayerNorm / attention kernels use programmatic dependent launch (default 0: measured slower in-step): each kernel's set-up (block
 * scheduling, mbarrier + TMEM allocation, descriptor prefetch) overlaps the tail of its predecessor in the stream and
 * it waits (griddepcontrol.wait) before its first global-memory access. 0 = plain stream order. */
int b200ssl_set_pdl(int on);
/* Developer A/B switch: LayerNorm backward with rows staged through shared memory by bulk copies (1, default) or the
 * register-resident version (0). */
int b200ssl_set_ln_bwd_staged(int on);
/* Developer instrumentation: device buffer of 16 uint64 cycle counters the GEMM kernels accumulate into (NULL = off). */
int b200ssl_set_gemm_prof(void* counters);
/* Same for attention forward: 16 uint64 (2 slots x {wait S, max pass, barrier, exp pass, barrier, wait O, epilogue, tiles}). */
int b200ssl_set_attn_prof(void* counters);

/* ---- K2: LayerNorm (Block.norm1/norm2 VT.pyc@L138,142,147,151; VisionTransformer.norm @L195,252) -----
 * x is bf16 (x_f32 = 0) or the fp32 residual stream (x_f32 = 1); y bf16; mean/rstd fp32 [rows].
 * bwd: dx(bf16) = LN'(dy) + dres (dres nullable: the residual-branch gradient is added in-kernel);
 * dw/db fp32 [D] are ACCUMULATED into. D in {192,256,384,512,768,1024}. */
int b200ssl_layernorm_fwd(const void* x, int x_f32, const float* w, const float* b, void* y, float* mean,
                          float* rstd, long long rows, int D, float eps, void* stream);
int b200ssl_layernorm_bwd(const void* x, int x_f32, const void* dy, const float* w, const float* mean,
                          const float* rstd, const void* dres, void* dx, float* dw, float* db, long long rows,
                          int D, void* stream);

/* ---- K4: fused attention (Attention.forward VT.pyc@L119-131) --------------------------------------
 * qkv [B,N,3,H,64] bf16 (the QKV GEMM output as is), out/dout [B,N,H,64] bf16, lse2 [B,H,N] fp32
 * (log2-sum-exp of the scaled scores), dqkv like qkv (32-byte aligned: the backward writes whole 128-byte rows with
 * 256-bit stores). head_dim must be 64. */
int b200ssl_attention_fwd(const void* qkv, void* out, float* lse2, int B, int N, int H, int head_dim,
                          float scale, void* stream);
int b200ssl_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse2, void* dqkv,
                          int B, int N, int H, int head_dim, float scale, void* stream);
/* Forward for sequences of more than 128 tokens has a STREAMING kernel (one 128-row query tile per CTA, key / value
 * blocks of 128 tokens streamed through shared memory, online softmax, two softmax teams on alternating key blocks,
 * no workspace): the default for N > 256 (the reference's native 256^2 tiles give 257 tokens; ViT-S/8 at 224^2 gives
 * 785; up to 65536). b200ssl_set_attn_stream: 0 (default) = streaming kernel for N > 256, 1 = for every N > 128,
 * -1 = never (developer A/B: N > 256 then runs as blocks of <= 256 queries x <= 256 keys through the two-tile
 * kernel, partial results merged by their log-sum-exp in a caller-provided scratch buffer, N <= 4096).
 * Backward for N > 256 runs (query block, key block) pairs of <= 256 x 256 tokens through the two-tile backward kernel
 * (dQ accumulates over key blocks, dK/dV over query blocks through TMA reduce-add; N <= 4096): four launches for
 * 257..512 tokens, ONE launch over all pairs (into a zeroed dqkv) above. b200ssl_attention_fwd == b200ssl_attention_fwd_ws with no
 * workspace; b200ssl_attention_fwd_workspace_bytes returns 0 whenever none is needed. */
int b200ssl_set_attn_stream(int mode);
/* Two-tile forward (N <= 256): 1 (default) = the MMA issuer serves whichever of the two query-tile slots is ready and
 * starts them half a period apart, so one slot's exp2 pass runs under the other's epilogue / row-max pass; 0 = both
 * slots issued round by round in lock-step (developer A/B). Results are bit-identical. */
int b200ssl_set_attn_dephase(int on);
long long b200ssl_attention_fwd_workspace_bytes(int B, int N, int H);
int b200ssl_attention_fwd_ws(const void* qkv, void* out, float* lse2, int B, int N, int H, int head_dim, float scale,
                             void* workspace, long long workspace_bytes, void* stream);

/* ---- §8f-3: multi-crop tile augmentation (replaces the per-tile host pipeline of datasets.py:498-502 /
 *      transformations.py:103-208 and adds the DINO multi-crop geometry). One CTA per source tile: the uint8
 *      256 x 256 x 3 tile (HWC, RGB) is staged in shared memory once and every crop is produced from it.
 *      tiles uint8 [B,256,256,3]; params [B, n_global + n_local, 16] 32-bit words per (tile, crop):
 *        {top, left, h, w (int32 crop box), flags (bit 0 hflip, bit 1 vflip, bits 2-3 rot90 k, bits 4-11 the order of
 *         brightness(0) / contrast(1) / saturation(2) / hue(3) as 4 x 2 bits, bit 12 colour jitter on),
 *         brightness, contrast, saturation, hue (float factors, torchvision semantics), noise sigma (float),
 *         noise seed (uint32), Cutout hole in output-frame pixels as y1 | y2 << 16 and x1 | x2 << 16 (two uint32; the
 *         hole [y1,y2) x [x1,x2) is written as 0 AFTER Normalize like Cutout at transformations.py:10-45 / :206-207;
 *         0 = no hole), 3 reserved};
 *      out_global bf16 [n_global,B,3,Sg,Sg], out_local bf16 [n_local,B,3,Sl,Sl] (crop-major); mean / std: HOST
 *      pointers to the 3 channel statistics of Normalize (transformations.py:104-116). Bilinear resize with
 *      align_corners = false and no antialiasing; GaussianBlur(3, sigma <= 0.1) is the identity to 2e-22 and skipped. */
int b200ssl_multicrop_augment(const void* tiles, const void* params, void* out_global, void* out_local, int B,
                              int n_global, int n_local, int size_global, int size_local, const float* mean,
                              const float* std, void* stream);

/* ---- Element dropout on the encoder path: nn.Dropout(drop_rate) of the reference's VisionTransformer -- pos_drop
 *      (VT.pyc@L196,245), Attention.proj_drop (@L117,130), Mlp.drop behind the activation and behind fc2 (@L96,101-104);
 *      train.py --drop (train.py:283,487). The keep mask is a pure function of (seed, site, element index) -- one
 *      splitmix64 word per four consecutive elements, 16 bits each, dropped when below round(p * 65536), kept elements
 *      scaled by 1 / (1 - p) -- so backward regenerates it (nothing is stored). seed: DEVICE pointer to one 64-bit word.
 *      b200ssl_dropout: dst = src * mask (bf16, or fp32 with is_f32; src == dst allowed; n % 8 == 0); `second` (bf16,
 *      optional) is masked in place with the same mask (the saved gelu' next to gelu).
 *      b200ssl_dropout_residual: y = residual + rowscale[row] * dropout(branch) (branch bf16 [rows, D]; residual, y fp32;
 *      rowscale optional, the stochastic-depth row scale): x + drop_path(drop(branch)) of Block.forward @L150-151. */
int b200ssl_dropout(const void* src, void* dst, void* second, long long n, int is_f32, float p, const void* seed,
                    unsigned site, void* stream);
int b200ssl_dropout_residual(const void* branch, const float* residual, const float* rowscale, float* y, long long rows,
                             int D, float p, const void* seed, unsigned site, void* stream);

/* ---- K1 helpers: patch gathering and token assembly (PatchEmbed.forward VT.pyc@L167-170,
 *      prepare_tokens @L235-246). img [B,C,H,W] bf16 -> cols [B*Np, C*P*P] bf16 (then b200ssl_gemm);
 *      x(fp32)[b,0]=cls+pos[0], x[b,1+p]=y[b*Np+p]+pos[1+p]; bwd: dy=dx[:,1:], dpos=sum_b dx, dcls=dpos[0]. */
int b200ssl_patchify(const void* img, void* cols, int B, int C, int H, int W, int P, void* stream);
int b200ssl_assemble_tokens(const void* y, const float* cls, const float* pos, void* x, int B, int Np, int D,
                            void* stream);
int b200ssl_assemble_tokens_bwd(const void* dx, void* dy, float* dpos, float* dcls, int B, int Np, int D,
                                void* stream);

/* ---- small HBM-bound helpers ------------------------------------------------------------------------ */
/* out[c] (+)= sum_r x[r,c]; bf16 in, fp32 out (accumulate != 0 keeps the previous content). */
int b200ssl_colsum(const void* x, long long ldx, float* out, long long rows, int ncols, int accumulate,
                   void* stream);
/* y[r,:] = x[r,:] * scale[r] (bf16 rows, fp32 scale[rows]): the branch gradient under stochastic depth. */
int b200ssl_scale_rows(const void* x, const float* scale, void* y, long long rows, int D, void* stream);
int b200ssl_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream);
/* the reverse (bf16 gradient buckets back into the fp32 buckets after a compressed all-reduce) */
int b200ssl_cast_bf16_to_f32(const void* src, float* dst, long long n, void* stream);
/* F.normalize(dim=-1, p=2, eps) (DINOHead.forward VT.pyc@L328): y = x / max(||x||, eps). */
int b200ssl_l2norm_fwd(const void* x, void* y, float* norm, long long rows, int D, float eps, void* stream);
int b200ssl_l2norm_bwd(const void* y, const void* dy, const float* norm, void* dx, long long rows, int D,
                       float eps, void* stream);
/* weight_norm(Linear) rows (VT.pyc@L315-318): w(bf16) = g * v / ||v||; g nullable (= 1). */
int b200ssl_weightnorm_fwd(const float* v, const float* g, void* w, float* norm, long long rows, int D,
                           void* stream);
/* accumulate != 0: dv / dg are ADDED to (they are the parameters' persistent .grad buffers), else overwritten. */
int b200ssl_weightnorm_bwd(const float* v, const float* g, const float* norm, const float* dw, float* dv,
                           float* dg, long long rows, int D, int accumulate, void* stream);

/* ---- BatchNorm1d (+ fused exact-erf GELU) for DINOHead(use_bn=True) (VT.pyc@L304-305,309-310) ----------------
 * x, h, dh, dx bf16 [rows, C]; w, b (nullable), running_mean / running_var, save_mean / save_invstd fp32 [C];
 * num_batches: int64 device scalar (nullable). training != 0: batch statistics (biased variance for the output,
 * unbiased for running_var, torch.nn.BatchNorm1d); training == 0: running statistics. dw / db are ACCUMULATED. */
int b200ssl_bn_gelu_fwd(const void* x, const float* w, const float* b, float* running_mean, float* running_var,
                        long long* num_batches, void* h, float* save_mean, float* save_invstd, long long rows, int C,
                        float momentum, float eps, int training, int apply_gelu, void* stream);
int b200ssl_bn_gelu_bwd(const void* x, const void* dh, const float* w, const float* b, const float* save_mean,
                        const float* save_invstd, void* dx, float* dw, float* db, long long rows, int C, int training,
                        int apply_gelu, void* stream);

/* ---- step plumbing (memset / copy nodes and three tiny kernels that replace PyTorch glue inside the captured
 *      step: torch.zeros, clone, x[:, 0] slicing + cat (VT.pyc@L253), its indexed-assignment backward, the
 *      accumulation of cls_token / pos_embed gradients and the position-table resize (VT.pyc@L213-233)). */
int b200ssl_zero_bytes(void* p, long long nbytes, void* stream);
int b200ssl_copy_bytes(void* dst, const void* src, long long nbytes, void* stream);
/* dst[r*dst_stride .. +row_bytes) = src[r*src_stride .. +row_bytes), r < rows; everything in multiples of 16 bytes. */
int b200ssl_copy_rows(const void* src, long long src_stride_bytes, void* dst, long long dst_stride_bytes,
                      long long rows, int row_bytes, void* stream);
/* dst[i] += src[i], fp32. */
int b200ssl_add_f32(float* dst, const float* src, long long n, void* stream);
/* interpolate_pos_encoding as the fixed linear map mat [Mo, Ki] (fp32; the bicubic weights incl. the +0.1 fudge):
 * transpose = 0: out[0]=in[0], out[1+i] = sum_j mat[i,j] in[1+j]   (in = pos_embed [1+Ki, D], out [1+Mo, D])
 * transpose = 1: out[0]+=in[0], out[1+j] += sum_i mat[i,j] in[1+i] (in = d(table) [1+Mo, D], out = d(pos_embed)). */
int b200ssl_pos_interp(const float* mat, const float* in, float* out, int Mo, int Ki, int D, int transpose,
                       void* stream);

/* ---- K7: DINO loss (not in the reference; call slot train.py:1053 loss_fn(output, target)) -----------
 * student [ncrops*B, K], teacher [2*B, K] bf16, crop-major rows; center fp32 [K]; loss fp32 [1];
 * s_lse [ncrops*B], t_lse [2*B] fp32 saved for backward; gout: device scalar (nullable = 1). */
int b200ssl_dino_loss_fwd(const void* student, const void* teacher, const float* center, float* loss,
                          float* s_lse, float* t_lse, int B, int ncrops, int K, float student_temp,
                          float teacher_temp, void* stream);
int b200ssl_dino_loss_bwd(const void* student, const void* teacher, const float* center, const float* s_lse,
                          const float* t_lse, const float* gout, void* dstudent, int B, int ncrops, int K,
                          float student_temp, float teacher_temp, void* stream);
/* center = m*center + (1-m)*batch_sum/total_rows (batch_sum = all-reduced column sums of teacher logits). */
int b200ssl_center_update(float* center, const float* batch_sum, int K, long long total_rows, float momentum,
                          void* stream);

/* ---- K8: multi-tensor EMA / AdamW / grad-norm (train.py:1081 model_ema.update; :1063-1078 clip+step) --
 * `table` is a device array of int64 rows, one per <= 65536-element chunk:
 *   ema   : {dst*, src*, n, bf16_shadow*}      dst = m*dst + (1-m)*src; shadow (nullable) = bf16(dst), same pass
 *   sumsq : {g*, n}                            out = float[1+n_rows] workspace, out[0] = sum g^2 (fixed order)
 *   adamw : {p*, g*, m*, v*, n, decay, ema*, bf16_shadow*}   torch.optim.AdamW rule, grads pre-scaled by
 *           min(1, max_norm/(sqrt(*gnorm_sq)+1e-6)) when gnorm_sq != NULL and max_norm > 0. */
/* momentum_dev / dev_hyper are DEVICE pointers: per-step scalars are read on the GPU so a captured CUDA
 * graph can be replayed with new values. dev_hyper = float[5] {lr, weight_decay, 1-beta1^t, 1-beta2^t, ema_m}. */
int b200ssl_ema_multi_tensor(const void* table, int n_rows, const float* momentum_dev, void* stream);
int b200ssl_sumsq_multi_tensor(const void* table, int n_rows, float* out, void* stream);
int b200ssl_adamw_multi_tensor(const void* table, int n_rows, const float* gnorm_sq, const float* dev_hyper,
                               float beta1, float beta2, float eps, float max_norm, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SSL_H_ */
