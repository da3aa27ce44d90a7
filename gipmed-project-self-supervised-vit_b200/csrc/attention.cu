// K4: fused multi-head attention, forward and backward, on tcgen05/TMEM fed by TMA.
//
// Replaces the materialised-score attention of the reference (VT.pyc@L119-131:
// qkv.reshape.permute -> (q @ k^T) * scale -> softmax -> attn @ v -> transpose.reshape) for
// head_dim 64 and sequences of up to 256 tokens (197 / 37 / 257->unsupported).
//
// Layout: qkv is the QKV-GEMM output as it lies in HBM, [B, N, 3, h, 64] bf16 (no permute copy);
// out / d_out are [B, N, h, 64]; lse2 is [B, h, N] fp32 holding log2-sum-exp of the scaled scores.
// A 3-D tensor map (cols, N, B) lets TMA clip rows >= N and batches >= B, so ragged tails need no
// special code. Sequences with N <= 64 are packed G = 128 / N per 128-row tile with a
// block-diagonal mask (local crops: N = 37 -> 3 sequences per tile).
//
// The whole key range of a sequence fits one MMA N extent (<= 256), so the softmax is single-pass.
// CTA = 1 control warp (TMA + tcgen05.mma issue) + 8 softmax/epilogue warps (2 threads per row).
#include "common.cuh"

namespace b200ssl {

constexpr int ATT_THREADS = 128 + 256;
constexpr int TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16

struct AttnArgs {
  int B, N, H;       // batch, tokens per sequence, heads
  int G;             // sequences packed per 128-row tile (NT == 1), else 1
  int rows;          // valid rows per tile group: G*N (NT == 1) or N (NT == 2)
  int keys_n;        // round_up(rows, 16): MMA N extent over keys
  float scale_log2;  // softmax scale * log2(e)
  float scale;
  float* lse2;
  const __nv_bfloat16* out;   // bwd only
  const __nv_bfloat16* dout;  // bwd only
};

// valid key range [lo, hi) for tile-row r
__device__ __forceinline__ void key_range(const AttnArgs& a, int nt, int r_in_group, int& lo, int& hi,
                                          bool& row_valid) {
  if (nt == 1) {
    row_valid = r_in_group < a.rows;
    const int g = row_valid ? r_in_group / a.N : 0;
    lo = g * a.N;
    hi = lo + a.N;
  } else {
    row_valid = r_in_group < a.N;
    lo = 0;
    hi = a.N;
  }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// smem: Q tiles (NT x 16K) | K (NT x 16K) | V (NT x 16K) | P (2NT x 16K) | barriers
template <int NT>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                     const AttnArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + NT * TILE_BYTES;
  uint8_t* sV = sK + NT * TILE_BYTES;
  uint8_t* sP = sV + NT * TILE_BYTES;  // 2*NT chunks of [128 rows][64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * NT * TILE_BYTES);
  uint64_t* bar_load = bars;        // 1
  uint64_t* bar_s = bars + 1;       // [NT]
  uint64_t* bar_p = bars + 3;       // [NT]
  uint64_t* bar_o = bars + 5;       // [NT]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* xchg = reinterpret_cast<float*>(bars + 10);  // [2][128]
  constexpr int TMEM_COLS = NT == 1 ? 128 : 512;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x % args.H;
  const int grp = blockIdx.x / args.H;  // sequence group (NT==1: G sequences; NT==2: one sequence)
  const int b0 = grp * args.G;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    mbar_init(bar_load, 1);
    for (int t = 0; t < NT; ++t) {
      mbar_init(&bar_s[t], 1);
      mbar_init(&bar_p[t], 256);
      mbar_init(&bar_o[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  if (NT == 1) {
    // rows the TMA box never writes must not feed NaN bit patterns into P(=0) x V
    const int first = args.rows * 128, last = args.keys_n * 128;
    for (int i = first + threadIdx.x * 16; i < last; i += ATT_THREADS * 16)
      *reinterpret_cast<uint4*>(sV + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int D3 = 3 * args.H * 64;
  (void)D3;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- control thread
      const int cq = head * 64, ck = (args.H + head) * 64, cv = (2 * args.H + head) * 64;
      if (NT == 1) {
        mbar_expect_tx(bar_load, 3 * args.rows * 128);
        tma_load_3d(sQ, &tmQKV, bar_load, cq, 0, b0);
        tma_load_3d(sK, &tmQKV, bar_load, ck, 0, b0);
        tma_load_3d(sV, &tmQKV, bar_load, cv, 0, b0);
      } else {
        mbar_expect_tx(bar_load, 3 * NT * TILE_BYTES);
        for (int t = 0; t < NT; ++t) {
          tma_load_3d(sQ + t * TILE_BYTES, &tmQKV, bar_load, cq, t * 128, b0);
          tma_load_3d(sK + t * TILE_BYTES, &tmQKV, bar_load, ck, t * 128, b0);
          tma_load_3d(sV + t * TILE_BYTES, &tmQKV, bar_load, cv, t * 128, b0);
        }
      }
      mbar_wait(bar_load, 0);
      tcgen05_fence_after();
      const uint32_t idesc_s = make_idesc_bf16(128, args.keys_n, false, false);
      const uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);
      for (int t = 0; t < NT; ++t) {
        const uint32_t a0 = smem_u32(sQ + t * TILE_BYTES), bk = smem_u32(sK);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base + t * 256, make_smem_desc_sw128(a0 + k * 32, 16, 1024),
                       make_smem_desc_sw128(bk + k * 32, 16, 1024), idesc_s, k > 0);
        umma_commit(&bar_s[t]);
      }
      const int ksteps = args.keys_n / 16;
      for (int t = 0; t < NT; ++t) {
        mbar_wait(&bar_p[t], 0);
        tcgen05_fence_after();
        const uint32_t p0 = smem_u32(sP), v0 = smem_u32(sV);
        for (int j = 0; j < ksteps; ++j)
          umma_bf16_ss(tmem_base + t * 256,
                       make_smem_desc_sw128(p0 + (j >> 2) * TILE_BYTES + (j & 3) * 32, 16, 1024),
                       make_smem_desc_sw128(v0 + j * 2048, 8192, 1024), idesc_o, j > 0);
        umma_commit(&bar_o[t]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax + epilogue warps
    const int q = warp & 3;
    const int hf = (warp - 4) >> 2;  // which half of the key columns this thread owns
    const int r = q * 32 + lane;     // row within the tile
    const int nchunks = args.keys_n / 16;
    const int c_begin = hf == 0 ? 0 : (nchunks + 1) / 2;
    const int c_end = hf == 0 ? (nchunks + 1) / 2 : nchunks;
    float inv_sum[NT];

    for (int t = 0; t < NT; ++t) {
      int lo, hi;
      bool row_valid;
      key_range(args, NT, NT == 1 ? r : t * 128 + r, lo, hi, row_valid);
      mbar_wait(&bar_s[t], 0);
      tcgen05_fence_after();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * 256;
      float s[8 * 16];  // up to 8 chunks of 16 columns (keys_n <= 256 -> <= 8 chunks per half)
      float mx = -INFINITY;
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        const int c = c_begin + ci;
        if (c < c_end) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(trow + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = c * 16 + j;
            const float x = (col >= lo && col < hi) ? __uint_as_float(v[j]) * args.scale_log2 : -INFINITY;
            s[ci * 16 + j] = x;
            mx = fmaxf(mx, x);
          }
        }
      }
      // all TMEM reads of S_t by this thread are done (O_t will overwrite its first 64 columns)
      xchg[hf * 128 + r] = mx;
      named_bar_sync(1, 256);
      mx = fmaxf(mx, xchg[(hf ^ 1) * 128 + r]);
      if (mx == -INFINITY) mx = 0.f;  // rows with no valid key in this tile (padding rows)
      named_bar_sync(1, 256);
      if (t > 0) {  // the previous P.V MMA must have finished reading sP
        mbar_wait(&bar_o[t - 1], 0);
      }
      float sum = 0.f;
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        const int c = c_begin + ci;
        if (c < c_end) {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float p0 = exp2f(s[ci * 16 + 2 * j] - mx);
            const float p1 = exp2f(s[ci * 16 + 2 * j + 1] - mx);
            const __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
            // sum what the MMA will actually see (bf16-rounded), keeps rows normalised
            sum += __low2float(pb) + __high2float(pb);
            pk[j] = *reinterpret_cast<const uint32_t*>(&pb);
          }
          uint8_t* chunk = sP + (c >> 2) * TILE_BYTES;
          const int c16 = (c & 3) * 2;
          *reinterpret_cast<uint4*>(chunk + sw128_offset(r, c16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(chunk + sw128_offset(r, c16 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      mbar_arrive(&bar_p[t]);
      xchg[hf * 128 + r] = sum;
      named_bar_sync(1, 256);
      sum += xchg[(hf ^ 1) * 128 + r];
      named_bar_sync(1, 256);
      inv_sum[t] = sum > 0.f ? 1.f / sum : 0.f;
      if (hf == 0 && row_valid) {
        const int rr = NT == 1 ? r : t * 128 + r;
        const int b = b0 + (NT == 1 ? rr / args.N : 0);
        const int n = NT == 1 ? rr % args.N : rr;
        if (b < args.B)
          args.lse2[(static_cast<long long>(b) * args.H + head) * args.N + n] = mx + log2f(sum);
      }
    }
    // epilogue: O_t / rowsum -> bf16 -> swizzled staging (re-uses the dead Q tile) -> TMA store
    for (int t = 0; t < NT; ++t) {
      mbar_wait(&bar_o[t], 0);
      tcgen05_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * 256 + hf * 32, v);
      tmem_ld_wait();
      uint8_t* stg = sQ + t * TILE_BYTES;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 pk;
        pk.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]) * inv_sum[t], __uint_as_float(v[8 * j + 1]) * inv_sum[t]);
        pk.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]) * inv_sum[t], __uint_as_float(v[8 * j + 3]) * inv_sum[t]);
        pk.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]) * inv_sum[t], __uint_as_float(v[8 * j + 5]) * inv_sum[t]);
        pk.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]) * inv_sum[t], __uint_as_float(v[8 * j + 7]) * inv_sum[t]);
        *reinterpret_cast<uint4*>(stg + sw128_offset(r, hf * 4 + j)) = pk;
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 256);
      if (threadIdx.x == 128) {
        tma_store_3d(&tmO, stg, head * 64, NT == 1 ? 0 : t * 128, b0);
        tma_store_commit();
      }
    }
    if (threadIdx.x == 128) tma_store_wait_all<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// smem: Q (NT) | dO (NT) | K (NT) | V (NT) | P (2 chunks) | dS (2 chunks) | staging | barriers
// TMEM: S [0,128) | dP [128,256) | dQ_t [256+64t) | dK [384,448) | dV [448,512)
template <int NT>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                     const __grid_constant__ CUtensorMap tmDQKV, const AttnArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + NT * TILE_BYTES;
  uint8_t* sK = sdO + NT * TILE_BYTES;
  uint8_t* sV = sK + NT * TILE_BYTES;
  uint8_t* sP = sV + NT * TILE_BYTES;
  uint8_t* sdS = sP + 2 * TILE_BYTES;
  uint8_t* stg = sdS + 2 * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + TILE_BYTES);
  uint64_t* bar_load = bars;
  uint64_t* bar_sdp = bars + 1;       // S and dP ready in TMEM
  uint64_t* bar_sdp_free = bars + 2;  // softmax threads done reading S/dP
  uint64_t* bar_pds = bars + 3;       // P and dS written to smem
  uint64_t* bar_mma = bars + 4;       // dQ/dK/dV MMAs of the pair finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x % args.H;
  const int grp = blockIdx.x / args.H;
  const int b0 = grp * args.G;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQKV);
    mbar_init(bar_load, 1);
    mbar_init(bar_sdp, 1);
    mbar_init(bar_sdp_free, 256);
    mbar_init(bar_pds, 256);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  if (NT == 1) {
    // rows never written by the TMA boxes: zero them so 0 x garbage cannot become NaN
    const int first = args.rows * 128, last = 128 * 128;
    for (int i = first + threadIdx.x * 16; i < last; i += ATT_THREADS * 16) {
      *reinterpret_cast<uint4*>(sQ + i) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sdO + i) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sK + i) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sV + i) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t T_S = tmem_base, T_DP = tmem_base + 128, T_DQ = tmem_base + 256, T_DK = tmem_base + 384,
                 T_DV = tmem_base + 448;

  if (warp == 0) {
    if (lane == 0) {
      const int cq = head * 64, ck = (args.H + head) * 64, cv = (2 * args.H + head) * 64;
      if (NT == 1) {
        mbar_expect_tx(bar_load, 4 * args.rows * 128);
        tma_load_3d(sQ, &tmQKV, bar_load, cq, 0, b0);
        tma_load_3d(sK, &tmQKV, bar_load, ck, 0, b0);
        tma_load_3d(sV, &tmQKV, bar_load, cv, 0, b0);
        tma_load_3d(sdO, &tmDO, bar_load, cq, 0, b0);
      } else {
        mbar_expect_tx(bar_load, 4 * NT * TILE_BYTES);
        for (int t = 0; t < NT; ++t) {
          tma_load_3d(sQ + t * TILE_BYTES, &tmQKV, bar_load, cq, t * 128, b0);
          tma_load_3d(sK + t * TILE_BYTES, &tmQKV, bar_load, ck, t * 128, b0);
          tma_load_3d(sV + t * TILE_BYTES, &tmQKV, bar_load, cv, t * 128, b0);
          tma_load_3d(sdO + t * TILE_BYTES, &tmDO, bar_load, cq, t * 128, b0);
        }
      }
      mbar_wait(bar_load, 0);
      tcgen05_fence_after();
      const uint32_t idesc_q = make_idesc_bf16(128, 64, false, true);  // dQ: A K-major, B MN-major
      const uint32_t idesc_kv = make_idesc_bf16(128, 64, true, true);  // dK/dV: both MN-major
      int pair = 0;
      for (int u = 0; u < NT; ++u) {
        const int ku = min(128, args.keys_n - u * 128);  // keys in this key tile (multiple of 16)
        const uint32_t idesc_s = make_idesc_bf16(128, ku, false, false);
        const uint32_t k_u = smem_u32(sK + u * TILE_BYTES), v_u = smem_u32(sV + u * TILE_BYTES);
        for (int t = 0; t < NT; ++t, ++pair) {
          const uint32_t q_t = smem_u32(sQ + t * TILE_BYTES), do_t = smem_u32(sdO + t * TILE_BYTES);
          if (pair > 0) {
            mbar_wait(bar_sdp_free, (pair - 1) & 1);
            tcgen05_fence_after();
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(T_S, make_smem_desc_sw128(q_t + k * 32, 16, 1024),
                         make_smem_desc_sw128(k_u + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(T_DP, make_smem_desc_sw128(do_t + k * 32, 16, 1024),
                         make_smem_desc_sw128(v_u + k * 32, 16, 1024), idesc_s, k > 0);
          umma_commit(bar_sdp);
          mbar_wait(bar_pds, pair & 1);
          tcgen05_fence_after();
          const uint32_t p0 = smem_u32(sP), ds0 = smem_u32(sdS);
          // dQ_t (+)= dS[128 q, ku keys] . K_u[ku keys, 64]
          for (int j = 0; j < ku / 16; ++j)
            umma_bf16_ss(T_DQ + t * 64,
                         make_smem_desc_sw128(ds0 + (j >> 2) * TILE_BYTES + (j & 3) * 32, 16, 1024),
                         make_smem_desc_sw128(k_u + j * 2048, 8192, 1024), idesc_q, (u > 0 || j > 0));
          // dV_u (+)= P^T[128 keys, 128 q] . dO_t[128 q, 64] ; dK_u (+)= dS^T . Q_t
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_bf16_ss(T_DV, make_smem_desc_sw128(p0 + j * 2048, TILE_BYTES, 1024),
                         make_smem_desc_sw128(do_t + j * 2048, 8192, 1024), idesc_kv, (t > 0 || j > 0));
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_bf16_ss(T_DK, make_smem_desc_sw128(ds0 + j * 2048, TILE_BYTES, 1024),
                         make_smem_desc_sw128(q_t + j * 2048, 8192, 1024), idesc_kv, (t > 0 || j > 0));
          umma_commit(bar_mma);
        }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int hf = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;

    // per-row constants: delta = sum_d dO*O, lse2; invalid rows get lse2 = +inf -> P = 0
    float delta[NT], lse2[NT];
    int lo[NT], hi[NT];
    for (int t = 0; t < NT; ++t) {
      bool row_valid;
      const int rr = NT == 1 ? r : t * 128 + r;
      key_range(args, NT, rr, lo[t], hi[t], row_valid);
      const int b = b0 + (NT == 1 ? rr / args.N : 0);
      const int n = NT == 1 ? rr % args.N : rr;
      delta[t] = 0.f;
      lse2[t] = INFINITY;
      if (row_valid && b < args.B) {
        const long long off = ((static_cast<long long>(b) * args.N + n) * args.H + head) * 64;
        const uint4* po = reinterpret_cast<const uint4*>(args.out + off);
        const uint4* pd = reinterpret_cast<const uint4*>(args.dout + off);
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 a = __ldg(po + j), d = __ldg(pd + j);
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 x = unpack_bf16x2(aw[e]), y = unpack_bf16x2(dw[e]);
            acc += x.x * y.x + x.y * y.y;
          }
        }
        delta[t] = acc;
        lse2[t] = args.lse2[(static_cast<long long>(b) * args.H + head) * args.N + n];
      }
    }

    int pair = 0;
    for (int u = 0; u < NT; ++u) {
      const int ku = min(128, args.keys_n - u * 128);
      for (int t = 0; t < NT; ++t, ++pair) {
        mbar_wait(bar_sdp, pair & 1);
        tcgen05_fence_after();
        // this thread owns key columns [hf*64, hf*64+64) of the tile (skipped when beyond ku)
        const int col0 = hf * 64;
        const bool active = col0 < ku;
        uint32_t sv[64], dv[64];
        if (active) {
          uint32_t a[32], b[32];
          tmem_ld_32x32b_x32(T_S + lane_off + col0, a);
          tmem_ld_32x32b_x32(T_S + lane_off + col0 + 32, b);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { sv[j] = a[j]; sv[32 + j] = b[j]; }
          tmem_ld_32x32b_x32(T_DP + lane_off + col0, a);
          tmem_ld_32x32b_x32(T_DP + lane_off + col0 + 32, b);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { dv[j] = a[j]; dv[32 + j] = b[j]; }
        }
        tcgen05_fence_before();
        mbar_arrive(bar_sdp_free);
        if (pair > 0) mbar_wait(bar_mma, (pair - 1) & 1);  // previous MMAs done with sP / sdS
        if (active) {
          uint8_t* pc = sP + hf * TILE_BYTES;
          uint8_t* dc = sdS + hf * TILE_BYTES;
#pragma unroll
          for (int j8 = 0; j8 < 8; ++j8) {
            uint32_t pp[4], dd[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float pv[2], dsv[2];
#pragma unroll
              for (int w = 0; w < 2; ++w) {
                const int j = j8 * 8 + e * 2 + w;
                const int col = u * 128 + col0 + j;
                const bool ok = col >= lo[t] && col < hi[t];
                const float p = ok ? exp2f(__uint_as_float(sv[j]) * args.scale_log2 - lse2[t]) : 0.f;
                pv[w] = p;
                dsv[w] = p * (__uint_as_float(dv[j]) - delta[t]) * args.scale;
              }
              pp[e] = pack_bf16x2(pv[0], pv[1]);
              dd[e] = pack_bf16x2(dsv[0], dsv[1]);
            }
            *reinterpret_cast<uint4*>(pc + sw128_offset(r, j8)) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
            *reinterpret_cast<uint4*>(dc + sw128_offset(r, j8)) = make_uint4(dd[0], dd[1], dd[2], dd[3]);
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(bar_pds);

        if (t == NT - 1) {
          // dK_u and dV_u are complete once this pair's MMAs retire
          mbar_wait(bar_mma, pair & 1);
          tcgen05_fence_after();
          for (int which = 0; which < 2; ++which) {
            uint32_t v[32];
            tmem_ld_32x32b_x32((which == 0 ? T_DK : T_DV) + lane_off + hf * 32, v);
            tmem_ld_wait();
            if (threadIdx.x == 128) tma_store_wait_read<0>();
            named_bar_sync(1, 256);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 pk;
              pk.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
              pk.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
              pk.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
              pk.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
              *reinterpret_cast<uint4*>(stg + sw128_offset(r, hf * 4 + j)) = pk;
            }
            fence_proxy_async_smem();
            named_bar_sync(1, 256);
            if (threadIdx.x == 128) {
              tma_store_3d(&tmDQKV, stg, ((which + 1) * args.H + head) * 64, NT == 1 ? 0 : u * 128, b0);
              tma_store_commit();
            }
          }
          // the dK/dV accumulators are re-used by the next key tile: order these reads before its MMAs
          tcgen05_fence_before();
        }
      }
    }
    // dQ tiles (complete after the last pair; bar_mma already waited on above)
    for (int t = 0; t < NT; ++t) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(T_DQ + t * 64 + lane_off + hf * 32, v);
      tmem_ld_wait();
      if (threadIdx.x == 128) tma_store_wait_read<0>();
      named_bar_sync(1, 256);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 pk;
        pk.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
        pk.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
        pk.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
        pk.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
        *reinterpret_cast<uint4*>(stg + sw128_offset(r, hf * 4 + j)) = pk;
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 256);
      if (threadIdx.x == 128) {
        tma_store_3d(&tmDQKV, stg, head * 64, NT == 1 ? 0 : t * 128, b0);
        tma_store_commit();
      }
    }
    if (threadIdx.x == 128) tma_store_wait_all<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

static int setup_args(AttnArgs& a, int B, int N, int H, float scale, int& nt, int& groups) {
  B200SSL_CHECK(N >= 1 && N <= 256, -2, "attention: sequence length %d unsupported (1..256)", N);
  a.B = B; a.N = N; a.H = H;
  nt = N <= 128 ? 1 : 2;
  a.G = nt == 1 ? 128 / N : 1;
  if (a.G > B) a.G = B;
  a.rows = nt == 1 ? a.G * N : N;
  a.keys_n = (a.rows + 15) / 16 * 16;
  a.scale = scale;
  a.scale_log2 = scale * 1.4426950408889634f;
  groups = (B + a.G - 1) / a.G;
  return 0;
}

static int make_bnd_map(CUtensorMap* tm, const void* base, int cols, int N, int B, int nt, int G) {
  uint64_t dims[3] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(N), static_cast<uint64_t>(B)};
  uint64_t strides[3] = {2, static_cast<uint64_t>(cols) * 2, static_cast<uint64_t>(cols) * 2 * N};
  uint32_t box[3] = {64, static_cast<uint32_t>(nt == 1 ? N : 128), static_cast<uint32_t>(nt == 1 ? G : 1)};
  return make_tensor_map(tm, base, 2, 3, dims, strides, box, true);
}

}  // namespace b200ssl

using namespace b200ssl;

extern "C" int b200ssl_attention_fwd(const void* qkv, void* out, float* lse2, int B, int N, int H, int head_dim,
                                     float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200SSL_CHECK(head_dim == 64, -2, "attention: head_dim %d unsupported (64 only)", head_dim);
  B200SSL_CHECK(B > 0 && H > 0, -2, "attention: empty problem");
  AttnArgs a{};
  int nt, groups;
  if (int rc = setup_args(a, B, N, H, scale, nt, groups)) return rc;
  a.lse2 = lse2;
  CUtensorMap tq, to;
  if (int rc = make_bnd_map(&tq, qkv, 3 * H * 64, N, B, nt, a.G)) return rc;
  if (int rc = make_bnd_map(&to, out, H * 64, N, B, nt, a.G)) return rc;
  const int grid = groups * H;
  if (nt == 1) {
    const int smem = 5 * TILE_BYTES + 2048 + 1024;
    static bool cfg = false;
    if (!cfg) {
      B200SSL_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cfg = true;
    }
    attention_fwd_kernel<1><<<grid, ATT_THREADS, smem, stream>>>(tq, to, a);
  } else {
    const int smem = 10 * TILE_BYTES + 2048 + 1024;
    static bool cfg = false;
    if (!cfg) {
      B200SSL_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cfg = true;
    }
    attention_fwd_kernel<2><<<grid, ATT_THREADS, smem, stream>>>(tq, to, a);
  }
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200ssl_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse2,
                                     void* dqkv, int B, int N, int H, int head_dim, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200SSL_CHECK(head_dim == 64, -2, "attention: head_dim %d unsupported (64 only)", head_dim);
  B200SSL_CHECK(B > 0 && H > 0, -2, "attention: empty problem");
  AttnArgs a{};
  int nt, groups;
  if (int rc = setup_args(a, B, N, H, scale, nt, groups)) return rc;
  a.lse2 = const_cast<float*>(lse2);
  a.out = static_cast<const __nv_bfloat16*>(out);
  a.dout = static_cast<const __nv_bfloat16*>(dout);
  CUtensorMap tq, tdo, tdq;
  if (int rc = make_bnd_map(&tq, qkv, 3 * H * 64, N, B, nt, a.G)) return rc;
  if (int rc = make_bnd_map(&tdo, dout, H * 64, N, B, nt, a.G)) return rc;
  if (int rc = make_bnd_map(&tdq, dqkv, 3 * H * 64, N, B, nt, a.G)) return rc;
  const int grid = groups * H;
  if (nt == 1) {
    const int smem = 9 * TILE_BYTES + 1024 + 1024;
    static bool cfg = false;
    if (!cfg) {
      B200SSL_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cfg = true;
    }
    attention_bwd_kernel<1><<<grid, ATT_THREADS, smem, stream>>>(tq, tdo, tdq, a);
  } else {
    const int smem = 13 * TILE_BYTES + 1024 + 1024;
    static bool cfg = false;
    if (!cfg) {
      B200SSL_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cfg = true;
    }
    attention_bwd_kernel<2><<<grid, ATT_THREADS, smem, stream>>>(tq, tdo, tdq, a);
  }
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}
