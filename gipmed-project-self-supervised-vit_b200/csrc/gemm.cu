// K3: persistent, warp-specialised tcgen05 GEMM for every Linear on the hot path.
//
//   D[M,N] = epilogue( A[M,K] . B[N,K]^T )            bf16 operands, fp32 accumulation in TMEM
//
// Replaces the cuBLASLt calls PyTorch makes for nn.Linear in the reference encoder/head
// (qkv VT.pyc@L114,121; proj @L116,129; fc1/fc2 @L93-95; DINOHead mlp/last_layer @L303-318,327-329)
// and their autograd dgrad / wgrad GEMMs.
//
// Operand layouts (runtime flags, so one kernel serves fprop, dgrad and wgrad):
//   A K-major  : row-major [M,K]          A MN-major : row-major [K,M]  (A^T stored)
//   B K-major  : row-major [N,K]          B MN-major : row-major [K,N]  (B^T stored)
//   fprop  Y = X W^T        : A = X (K-major),  B = W  (K-major)
//   dgrad  dX = dY W        : A = dY (K-major), B = W  (MN-major, reduction over W's rows)
//   wgrad  dW = dY^T X      : A = dY (MN-major), B = X (MN-major), reduction over tokens, split-K
//
// Structure per CTA (one CTA per SM, persistent over output tiles of 128 x BN):
//   warp 0      : TMA producer   (global -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      : MMA issuer     (one elected lane issues tcgen05.mma; accumulators in TMEM,
//                                 2 accumulator stages so tile i's epilogue overlaps tile i+1's MMAs)
//   warps 4..19 : epilogue       (4 groups x 4 warps, 32 columns at a time; tcgen05.ld -> bias / GELU+GELU' /
//                                 residual / x aux -> bf16|fp32 -> swizzled smem -> TMA store; or fp32 red.add
//                                 for split-K wgrad, whose bias gradient they fold from the smem stages)
#include "common.cuh"

namespace b200ssl {

enum GemmEpilogue : int {
  EPI_BIAS = 0,       // D = acc (+ bias)
  EPI_BIAS_GELU = 1,  // x = acc + bias ; D = gelu'(x) (saved for backward) ; D2 = gelu(x)
  EPI_BIAS_RES = 2,   // D = acc (+ bias) + aux
  EPI_MUL_AUX = 3,    // D = acc * aux   (dgrad through GELU: aux = saved gelu'(x))
  EPI_ATOMIC_F32 = 4, // D(fp32) += acc   (split-K)
  EPI_BIAS_RES_F32 = 5, // D(fp32) = acc (+ bias) + aux(fp32): the fp32 residual stream
  EPI_ATOMIC_F32_T = 6, // D(fp32)[n, m] += acc[m, n] (split-K, TRANSPOSED store); bias output = column sums of B
  EPI_BIAS_GELU_FWD = 7, // D = gelu(acc + bias): no-grad forward (teacher), nothing saved for backward
};
__host__ __device__ constexpr bool is_wgrad_epi(int e) { return e == EPI_ATOMIC_F32 || e == EPI_ATOMIC_F32_T; }

struct GemmArgs {
  int M, N, K;
  int a_mn, b_mn;
  int num_m_blocks, num_n_blocks, k_splits, k_blocks_per_split, k_blocks_total;
  const float* bias;  // fprop bias [N]; for EPI_ATOMIC_F32 (wgrad) the OUTPUT db[M] += column sums of A, or null
  const void* aux;  // bf16 (EPI_BIAS_RES, EPI_MUL_AUX) or fp32 (EPI_BIAS_RES_F32)
  long long ldaux;
  float* out_f32;
  long long ldd;
  const float* rowscale;     // EPI_BIAS_RES_F32 only (nullable): D = rowscale[row] * (acc + bias) + aux (stochastic depth)
  unsigned long long* prof;  // developer instrumentation (null = off): per-role wait / busy cycle counters
  // LayerNorm-fused mode (MODE 2): A = LN(x) is produced inside the kernel from the fp32 residual stream
  const float* ln_x;
  long long ln_ldx;
  const float* ln_gamma;
  const float* ln_beta;
  float ln_eps;
  __nv_bfloat16* ln_out;  // optional: the normalised rows (bf16, [M, K]) saved for the backward wgrad
  float* ln_mean;         // optional: per-row statistics saved for the LayerNorm backward
  float* ln_rstd;
};

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int SLAB_BYTES = 32 * 64;                    // one epilogue slab: 32 rows x 64 B (32 bf16 or 16 fp32 columns)
constexpr int STAGING_BYTES = 4 * 2 * SLAB_BYTES;      // per epilogue group: 4 warps x ring of two slabs
constexpr int NUM_EPI_GROUPS = 4;                      // 4 groups x 4 warps: 4 warps per TMEM lane quarter
constexpr int GEMM_THREADS = 128 + NUM_EPI_GROUPS * 128;

constexpr int kSmemBudget = 232448 - 2048;  // 227 KB opt-in limit minus barrier block and alignment slack

constexpr int kMaxPanelKBlocks = 6;  // B-stationary mode: the whole K extent (<= 384) of the B tile stays in smem

// MODE 0: A and B stream through the ring. MODE 1 (B-stationary): the B tile stays in a smem panel, the ring carries
// A only. MODE 2 (LayerNorm-fused, A-stationary): the epilogue warps normalise 128 rows of the fp32 residual stream
// straight into a bf16 A panel in the UMMA layout, every n block of those rows is computed from it, the ring
// carries B only. MODE 3 (LayerNorm TAIL, fp32-residual epilogue, N = 384 as two 192-column n blocks): a cluster computes
// BOTH n blocks of its m unit back to back, and one unit later two otherwise idle warps read the finished fp32 rows
// back -- from L2, they were written microseconds ago -- and emit LayerNorm(rows) in bf16 plus the row statistics:
// the LayerNorm that follows every residual add (Block.norm2 after attn.proj, the next Block.norm1 after mlp.fc2,
// VT.pyc@L147-151) then costs its bf16 write only, not a second pass over the fp32 stream from HBM.
template <int BN, int CL, int MODE = 0, bool WG = false>
struct GemmCfg {
  static constexpr bool BS = MODE == 1;
  static constexpr bool LN = MODE == 2;
  // per-CTA bytes of one k-block of B: (in CTA-pair mode half of) the B tile
  static constexpr int kBStageBytes = BN * BLOCK_K * 2 / CL;
  // B-stationary: B lives in a panel of kMaxPanelKBlocks k-blocks loaded once; the ring carries A only
  static constexpr int kPanelBytes = BS ? kMaxPanelKBlocks * kBStageBytes : LN ? kMaxPanelKBlocks * A_STAGE_BYTES : 0;
  static constexpr int kStageBytes = BS ? A_STAGE_BYTES : LN ? kBStageBytes : A_STAGE_BYTES + kBStageBytes;
  // split-K wgrad (WG) stores straight from registers (warp-shuffle transpose + coalesced reds): no epilogue
  // staging, so its long K loop gets one or two more pipeline stages
  static constexpr int kStagingBytes = WG ? 0 : NUM_EPI_GROUPS * STAGING_BYTES;
  // as many stages as fit: the operand feed is latency bound (bytes in flight per SM / ~1 us L2 latency),
  // so depth matters more than anything else; pair mode gets 5-6 stages where single-CTA mode gets 3-4
  static constexpr int kStagesFit = (kSmemBudget - kStagingBytes - kPanelBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  // BN = 384 (CTA pair only): one 128 x 384 accumulator (MMAs of N = 256 + 128 per k step); else two stages
  static constexpr int kAccStages = BN > 256 ? 1 : 2;
  static constexpr int kTmemCols = BN * kAccStages <= 128 ? 128 : BN * kAccStages <= 256 ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kPanelBytes + kStagingBytes + 1024 /*barriers*/ + 1024 /*align*/;
};

// Work distribution. Default: output tiles round-robin over the clusters, n fastest (all n blocks of an m unit
// run at about the same time on different clusters, so A is read from HBM once and hits L2 afterwards).
// B-stationary (BS): cluster c keeps ONE n block (c % num_n_blocks) for the whole kernel and sweeps the m
// units rank, rank + group_size, ...; the groups of different n blocks sweep m in the same order at the same
// pace, which keeps the L2 reuse of A.
struct TileIter {
  int t, step, total;        // default mode: linear tile index; other modes: m unit index
  int n_blk, m_unit, ks;
  // default mode: t = (ks * num_m_units + m_unit) * num_n_blocks + n_blk advances by `step` tiles; the three coordinates are
  // carried along incrementally (step_n = step % num_n_blocks, step_m = step / num_n_blocks) -- a division by a run-time
  // value is ~40 dependent instructions, and the producer / MMA-issuer threads run alone: dividing afresh for every tile
  // (and, in the epilogue's aux prefetch, for every chunk) put hundreds of cycles between tiles
  int step_n, step_m, nmu;
  template <int MODE>
  __device__ __forceinline__ void init(const GemmArgs& a, int cluster_id, int num_clusters, int num_m_units) {
    if (MODE == 1) {
      n_blk = cluster_id % a.num_n_blocks;
      const int rank = cluster_id / a.num_n_blocks;
      step = num_clusters / a.num_n_blocks + (n_blk < num_clusters % a.num_n_blocks ? 1 : 0);
      t = rank;
      total = num_m_units;
    } else if (MODE >= 2) {
      // LayerNorm-fused (2: A panel, 3: LayerNorm tail): a cluster owns whole m units and walks all their n blocks
      n_blk = 0;
      t = cluster_id;
      step = num_clusters;
      total = num_m_units;
    } else {
      t = cluster_id;
      step = num_clusters;
      total = num_m_units * a.num_n_blocks * a.k_splits;
      step_n = step % a.num_n_blocks;
      step_m = step / a.num_n_blocks;
      nmu = num_m_units;
      n_blk = t % a.num_n_blocks;
      const int m_lin = t / a.num_n_blocks;
      ks = m_lin / num_m_units;
      m_unit = m_lin - ks * num_m_units;
    }
  }
  template <int MODE>
  __device__ __forceinline__ bool valid(const GemmArgs& a, int num_m_units) {
    if (t >= total) return false;
    if (MODE >= 1) {
      m_unit = t;
      ks = 0;
    }
    return true;
  }
  template <int MODE>
  __device__ __forceinline__ void next(const GemmArgs& a) {
    if (MODE >= 2) {
      if (++n_blk == a.num_n_blocks) { n_blk = 0; t += step; }
    } else if (MODE == 1) {
      t += step;
    } else {
      t += step;
      n_blk += step_n;
      m_unit += step_m;
      if (n_blk >= a.num_n_blocks) { n_blk -= a.num_n_blocks; ++m_unit; }
      while (m_unit >= nmu) { m_unit -= nmu; ++ks; }
    }
  }
};

// CL = CTAs per cluster (1 or 2). CL == 2 is the CTA-pair mode (tcgen05 cta_group::2): the two CTAs of a
// cluster own vertically adjacent 128-row tiles of one 256 x BN output tile. Each CTA loads its own 128 rows
// of A and only HALF of the B tile; the leader CTA (rank 0) issues 256 x BN x 16 MMAs that read both halves
// of B from the two SMs' shared memory and write each CTA's 128 rows into its own TMEM. L2 -> smem bytes per
// MMA drop by a third (the 128 x BN tiling is L2-bandwidth bound, ~11 TB/s measured).
//   * both CTAs' TMA loads complete on the LEADER's full barrier (peer-bit mask), count 2 = leader's
//     arrive.expect_tx (both CTAs' bytes) + the peer's remote arrive;
//   * the leader's tcgen05.commit is multicast to both CTAs (stage-empty / accumulator-full barriers);
//   * both CTAs' epilogue warps arrive on the leader's accumulator-empty barrier.
template <int BN, int EPI, int CL, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmD2,
            const GemmArgs args) {
  using Cfg = GemmCfg<BN, CL, MODE, is_wgrad_epi(EPI)>;
  constexpr bool BS = MODE == 1;
  constexpr bool LN = MODE == 2;
  constexpr bool LNT = MODE == 3;
  static_assert(!LNT || (EPI == EPI_BIAS_RES_F32 && CL == 2 && BN == 192), "the LayerNorm tail is built for the fp32-residual epilogue, 256 x 192 pair tiles");
  constexpr int kStages = Cfg::kStages;
  static_assert(kStages >= 2, "smem ring too shallow");
  static_assert(!LN || (CL == 2 && !is_wgrad_epi(EPI)), "the LayerNorm-fused mode is built for CTA pairs, fprop only");
  constexpr int kAccStages = Cfg::kAccStages;
  constexpr bool kWgrad = is_wgrad_epi(EPI);
  static_assert(!BS || !kWgrad, "B-stationary mode is for the non-split-K epilogues");
  static_assert(BN <= 256 || CL == 2, "BN = 384 needs the CTA-pair mode");
  static_assert(EPI != EPI_ATOMIC_F32_T || BN == 384, "the transposed wgrad store is built for BN = 384");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* panel = smem + kStages * Cfg::kStageBytes;  // BS: the stationary B tile, kMaxPanelKBlocks k-blocks
  uint8_t* staging = panel + Cfg::kPanelBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;                 // [kStages]
  uint64_t* empty_bar = bars + kStages;      // [kStages]
  uint64_t* tmem_full = bars + 2 * kStages;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint64_t* consumed_bar = tmem_empty + 2;   // [kStages] CTA-pair wgrad: "MMA is done reading this stage"
  uint64_t* panel_full = consumed_bar + kStages;  // [1] BS: the B panel has landed; LN: the A panel is written
  uint64_t* panel_empty = panel_full + 1;         // [1] LN: every MMA that reads the A panel has retired
  uint64_t* ln_go = panel_empty + 1;              // [2] LayerNorm tail: this CTA's rows of an m unit have landed in memory
  uint64_t* ln_done = ln_go + 2;                  // [2] ... and the helper warps have normalised them
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ln_done + 2);

  const long long t_entry = args.prof ? clock64() : 0;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = CL > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int cluster_id = CL > 1 ? blockIdx.x / CL : blockIdx.x;
  const int num_clusters = CL > 1 ? gridDim.x / CL : gridDim.x;
  const int num_m_units = (args.num_m_blocks + CL - 1) / CL;  // CL vertically adjacent m blocks per unit

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (!kWgrad) tma_prefetch_desc(&tmD);
    if (EPI == EPI_BIAS_GELU) tma_prefetch_desc(&tmD2);
  }
  if (warp == 1 && lane == 0) {
    constexpr int kEpiWarps = NUM_EPI_GROUPS * 4;
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], CL);
      // stage release: the MMA commit, plus (wgrad) the epilogue warps that fold the bias gradient from the
      // stage; in CTA-pair wgrad the commit goes to consumed_bar instead and only the warps release
      mbar_init(&empty_bar[s], kWgrad ? (CL == 1 ? 1 + kEpiWarps : kEpiWarps) : 1);
      mbar_init(&consumed_bar[s], 1);
    }
    mbar_init(panel_full, LN ? CL * NUM_EPI_GROUPS * 4 : CL);  // LN: one arrive per transform warp of the pair
    mbar_init(panel_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ln_go[s], NUM_EPI_GROUPS * 4);  // one arrive per epilogue warp of THIS CTA
      mbar_init(&ln_done[s], 2);                 // the two helper warps
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], CL * kEpiWarps);  // one arrive per epilogue warp of every CTA of the pair
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CL == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    else tmem_alloc_2sm<Cfg::kTmemCols>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform for the compiler (MMA issuer)
  // everything above touched only this CTA's shared / tensor memory: let the next kernel in the stream start its
  // own set-up, then wait for the previous one's results before the first global access
  pdl_launch_dependents();
  pdl_wait();

  // Registers: 20 warps cap the kernel at 96 per thread, and at 96 the fused epilogues (GELU + GELU', x aux) spill a few
  // values inside their chunk loops (here the re-loads still hit what is left of L1 -- ncu: 100 % -- in the attention
  // backward, with less L1 to spare, they were L2 round trips). Warpgroup 0 (producer,
  // MMA issuer, TMEM owner, spare) needs far fewer: it shrinks to GEMM_WG0_REGS and the sixteen epilogue warps grow to
  // GEMM_EPI_REGS (128 x 40 given back >= 512 x 8 taken). Not in the LayerNorm-fused modes, whose helper warps 2 / 3 do
  // arithmetic of their own.
  if constexpr (!LN && !LNT) {
    if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;\n");   // the epilogue warps' inc sits at the top of their branch
  }

  TileIter it;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // (elect.sync rather than lane == 0, here and in the MMA issuer: the compiler then knows ONE thread runs the block
    //  and emits UTMALDG / UTCHMMA straight instead of wrapping each in a loop over the active lanes)
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      it.init<MODE>(args, cluster_id, num_clusters, num_m_units);
      if (BS && it.t < it.total) {
        // the stationary B tile: all k-blocks of this cluster's n block, (pair mode) own half of the rows
        constexpr int kHalfN = BN / CL;
        const int n0 = it.n_blk * BN + cta_rank * kHalfN;
        for (int kb = 0; kb < args.k_blocks_total; ++kb) {
          uint8_t* sB = panel + kb * Cfg::kBStageBytes;
          const int k0 = kb * BLOCK_K;
          if (!args.b_mn) {
            if (CL == 1) tma_load_2d(sB, &tmB, panel_full, k0, n0);
            else tma_load_2d_2sm(sB, &tmB, panel_full, k0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < kHalfN / 64; ++c) {
              if (CL == 1) tma_load_2d(sB + c * 8192, &tmB, panel_full, n0 + c * 64, k0);
              else tma_load_2d_2sm(sB + c * 8192, &tmB, panel_full, n0 + c * 64, k0);
            }
          }
        }
        if (cta_rank == 0) mbar_expect_tx(panel_full, CL * args.k_blocks_total * Cfg::kBStageBytes);
        else mbar_arrive_leader(panel_full);
      }
      for (; it.valid<MODE>(args, num_m_units); it.next<MODE>(args)) {
        const int n_blk = it.n_blk;
        const int m_blk = it.m_unit * CL + cta_rank;
        const int kb0 = it.ks * args.k_blocks_per_split;
        const int kb1 = min(kb0 + args.k_blocks_per_split, args.k_blocks_total);
        const int m0 = m_blk * BLOCK_M, n0 = n_blk * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::kStageBytes;
          uint8_t* sB = LN ? sA : sA + A_STAGE_BYTES;  // LN: the ring carries B only (A is produced in-kernel)
          const int k0 = kb * BLOCK_K;
          if (CL == 1) {
            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            if (!args.a_mn) {
              tma_load_2d(sA, &tmA, &full_bar[stage], k0, m0);
            } else {
#pragma unroll
              for (int c = 0; c < BLOCK_M / 64; ++c)
                tma_load_2d(sA + c * 8192, &tmA, &full_bar[stage], m0 + c * 64, k0);
            }
            if (BS) {
              // B is stationary
            } else if (!args.b_mn) {
              tma_load_2d(sB, &tmB, &full_bar[stage], k0, n0);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c)
                tma_load_2d(sB + c * 8192, &tmB, &full_bar[stage], n0 + c * 64, k0);
            }
          } else {
            // CTA pair: own 128 rows of A, own half of the B tile; completion lands on the leader's barrier
            constexpr int kHalfN = BN / 2;
            if (LN) {
              // A comes from the in-kernel LayerNorm panel
            } else if (!args.a_mn) {
              tma_load_2d_2sm(sA, &tmA, &full_bar[stage], k0, m0);
            } else {
#pragma unroll
              for (int c = 0; c < BLOCK_M / 64; ++c)
                tma_load_2d_2sm(sA + c * 8192, &tmA, &full_bar[stage], m0 + c * 64, k0);
            }
            if (BS) {
              // B is stationary
            } else if (BN == 384) {
              // per CTA 192 of the 384 columns as three 64-wide pieces: two for the N = 256 MMA (columns
              // [128 r, 128 r + 128)) and one for the N = 128 MMA (columns 256 + [64 r, 64 r + 64))
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const int col = c < 2 ? n0 + cta_rank * 128 + c * 64 : n0 + 256 + cta_rank * 64;
                if (!args.b_mn) tma_load_2d_2sm(sB + c * 8192, &tmB, &full_bar[stage], k0, col);
                else tma_load_2d_2sm(sB + c * 8192, &tmB, &full_bar[stage], col, k0);
              }
            } else if (!args.b_mn) {
              tma_load_2d_2sm(sB, &tmB, &full_bar[stage], k0, n0 + cta_rank * kHalfN);
            } else {
#pragma unroll
              for (int c = 0; c < kHalfN / 64; ++c)
                tma_load_2d_2sm(sB + c * 8192, &tmB, &full_bar[stage], n0 + cta_rank * kHalfN + c * 64, k0);
            }
            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            else mbar_arrive_leader(&full_bar[stage]);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (cta_rank == 0 && elect_one_sync()) {  // CTA pair: only the leader issues MMAs
      const uint32_t idesc = make_idesc_bf16(BLOCK_M * CL, BN == 384 ? 256 : BN, args.a_mn != 0, args.b_mn != 0);
      const uint32_t idesc2 = make_idesc_bf16(BLOCK_M * CL, 128, args.a_mn != 0, args.b_mn != 0);  // BN = 384 only
      // Descriptors (SWIZZLE_128B, SBO 1024): high word constant; low word = address field | LBO field. Stepping an
      // operand by `bytes` is ONE add of bytes >> 4 to the low word -- the issuing thread runs alone on the uniform
      // datapath and the mask / shift / or chain of make_smem_desc_sw128 per operand per MMA took about as long as a
      // 128 x 192 x 16 MMA itself.
      const uint32_t desc_hi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      const uint32_t a_lbo_f = (((args.a_mn ? 8192u : 16u) >> 4) & 0x3FFFu) << 16;
      const uint32_t b_lbo_f = (((args.b_mn ? 8192u : 16u) >> 4) & 0x3FFFu) << 16;
      const uint32_t a_kstep = (args.a_mn ? UMMA_K * 128u : UMMA_K * 2u) >> 4;
      const uint32_t b_kstep = (args.b_mn ? UMMA_K * 128u : UMMA_K * 2u) >> 4;
      const uint32_t lo_smem = (smem_u32(smem) & 0x3FFFFu) >> 4, lo_panel = (smem_u32(panel) & 0x3FFFFu) >> 4;
      auto mk_desc = [&](uint32_t lo) -> uint64_t {
        uint64_t d;
        asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(desc_hi));
        return d;
      };
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      it.init<MODE>(args, cluster_id, num_clusters, num_m_units);
      if (BS && it.t < it.total) {
        mbar_wait(panel_full, 0);
        tcgen05_fence_after();
      }
      long long w_full = 0, w_acc = 0, w_panel = 0;
      uint32_t ln_phase = 0;
      const long long t_begin = clock64();
      for (; it.valid<MODE>(args, num_m_units); it.next<MODE>(args)) {
        const int kb0 = it.ks * args.k_blocks_per_split;
        const int kb1 = min(kb0 + args.k_blocks_per_split, args.k_blocks_total);
        if (LN && it.n_blk == 0) {
          // a new m unit: both CTAs' transform warps have written (and proxy-fenced) their A panels
          const long long tp = clock64();
          mbar_wait_cluster(panel_full, ln_phase);
          w_panel += clock64() - tp;
          ln_phase ^= 1;
          tcgen05_fence_after();
        }
        long long tw = clock64();
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        w_acc += clock64() - tw;
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          tw = clock64();
          mbar_wait(&full_bar[stage], phase);
          w_full += clock64() - tw;
          tcgen05_fence_after();
          const uint32_t loStage = lo_smem + static_cast<uint32_t>(stage) * (Cfg::kStageBytes >> 4);
          const uint32_t loA = (LN ? lo_panel + static_cast<uint32_t>(kb) * (A_STAGE_BYTES >> 4) : loStage) + a_lbo_f;
          const uint32_t loB = (BS ? lo_panel + static_cast<uint32_t>(kb) * (Cfg::kBStageBytes >> 4)
                                   : LN ? loStage : loStage + (A_STAGE_BYTES >> 4)) + b_lbo_f;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adesc = mk_desc(loA + k * a_kstep);
            const uint64_t bdesc = mk_desc(loB + k * b_kstep);
            if (CL == 1) umma_bf16_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16_ss_2sm(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            if (BN == 384) {
              // second MMA of the k step: output columns [256, 384) from the third 64-wide B piece of each CTA
              const uint64_t bdesc2 = mk_desc(loB + (16384u >> 4) + k * b_kstep);
              umma_bf16_ss_2sm(d_tmem + 256, adesc, bdesc2, idesc2, (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
          if (CL == 1) umma_commit(&empty_bar[stage]);
          else umma_commit_2sm_mc(kWgrad ? &consumed_bar[stage] : &empty_bar[stage], 3);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (CL == 1) umma_commit(&tmem_full[acc]);
        else umma_commit_2sm_mc(&tmem_full[acc], 3);
        if (LN && it.n_blk == args.num_n_blocks - 1) umma_commit_2sm_mc(panel_empty, 3);  // the A panels may be rewritten
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
      if (args.prof) {
        atomicAdd(args.prof + 0, static_cast<unsigned long long>(w_full));
        atomicAdd(args.prof + 1, static_cast<unsigned long long>(w_acc));
        atomicAdd(args.prof + 2, static_cast<unsigned long long>(clock64() - t_begin));
        atomicAdd(args.prof + 3, 1ull);
        if (LN) atomicAdd(args.prof + 6, static_cast<unsigned long long>(w_panel));
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    if constexpr (!LN && !LNT) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;\n");
    // 16 fully independent warps. Warp (q, group) owns TMEM lane quarter q (rows 32q .. 32q+31 of the tile) and
    // the column chunks c with (c + tile_iter) % 4 == group (rotating, so uneven chunk counts even out). Its
    // private staging is a ring of two 2 KB slabs (32 rows x 64 B, 64B swizzle) from which lane 0 issues 32-row
    // TMA stores -- no block-level barrier anywhere in the epilogue. Everything with global-memory latency
    // (bias, the aux slab of the first chunk) is fetched BEFORE the wait for the accumulator; the aux slab of
    // chunk i+1 is fetched while chunk i is processed.
    const int q = warp & 3;
    const int group = (warp - 4) >> 2;
    const uint32_t slab_base = smem_u32(staging) + static_cast<uint32_t>(warp - 4) * (2 * SLAB_BYTES);
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    constexpr int W = EPI == EPI_BIAS_RES_F32 ? 16 : 32;  // columns per chunk (one slab row = 64 B)
    constexpr int kChunks = BN / W;
    constexpr int kMaxOwn = (kChunks + NUM_EPI_GROUPS - 1) / NUM_EPI_GROUPS;
    constexpr bool kAuxBf16 = EPI == EPI_BIAS_RES || EPI == EPI_MUL_AUX;
    constexpr bool kAuxF32 = EPI == EPI_BIAS_RES_F32;
    constexpr bool kAux = kAuxBf16 || kAuxF32;
    constexpr int kAuxEsize = kAuxF32 ? 4 : 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t ring = 0;  // slab ring position
    int cs_stage = 0;   // wgrad only: position in the smem ring (bias-gradient pass)
    uint32_t cs_phase = 0;
    int tile_iter = 0;
    int ln_unit = 0;    // LN mode: m units transformed so far

    // aux operand (fp32 residual / bf16 residual / saved gelu'): fetched TWO owned chunks ahead -- across tile
    // boundaries -- into two alternating register sets, in a coalesced mapping (lane -> row k*8 + lane/4,
    // 16-byte piece lane%4 of the 32-row x 64-byte slab); the set is parked in the slab and refilled right away.
    TileIter pit;
    int p_tile_iter = 0, p_i = 0;
    uint4 aux_a[4], aux_b[4];
    uint32_t aux_par = 0;
    auto prefetch_aux = [&](uint4 (&dst)[4]) {
      while (pit.valid<MODE>(args, num_m_units)) {
        const int c = ((group - p_tile_iter) & 3) + NUM_EPI_GROUPS * p_i;
        if (c < kChunks) {
          const int prow0 = (pit.m_unit * CL + cta_rank) * BLOCK_M + q * 32;
          const int pcol0 = pit.n_blk * BN + c * W;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int arow = prow0 + k * 8 + (lane >> 2);
            dst[k] = make_uint4(0, 0, 0, 0);
            if (arow < args.M) {
              const uint8_t* ap = static_cast<const uint8_t*>(args.aux) +
                                  (static_cast<long long>(arow) * args.ldaux + pcol0) * kAuxEsize;
              dst[k] = __ldg(reinterpret_cast<const uint4*>(ap) + (lane & 3));
            }
          }
          ++p_i;
          return;
        }
        pit.next<MODE>(args);
        ++p_tile_iter;
        p_i = 0;
      }
    };
    if (kAux) {
      pit.init<MODE>(args, cluster_id, num_clusters, num_m_units);
      prefetch_aux(aux_a);
      prefetch_aux(aux_b);
    }

    it.init<MODE>(args, cluster_id, num_clusters, num_m_units);
    for (; it.valid<MODE>(args, num_m_units); it.next<MODE>(args), ++tile_iter) {
      const int n_blk = it.n_blk;
      const int m_blk = it.m_unit * CL + cta_rank;
      const int m0 = m_blk * BLOCK_M, n0 = n_blk * BN;
      const int row0 = m0 + q * 32;  // first row of this warp's 32-row slab

      if (kWgrad) {
        // wgrad: the epilogue warps are idle during the (long, split-K) mainloop, and the operand that is dY
        // (A = dY^T as MN-major [64 k][64 m] pieces; for the transposed variant B) is already in smem for the
        // MMA, so they fold the bias gradient db[col] += sum_k dY[k, col] from there -- on the tiles that see
        // each (column block, k split) exactly once -- and release every stage like a second consumer.
        const int kb0 = it.ks * args.k_blocks_per_split;
        const int kb1 = min(kb0 + args.k_blocks_per_split, args.k_blocks_total);
        float* db = const_cast<float*>(args.bias);
        constexpr bool kT = EPI == EPI_ATOMIC_F32_T;
        constexpr int kGroups = kT ? 24 : 16;       // 16-byte column groups (8 columns) in the folded operand
        constexpr int kSubs = 512 / kGroups;        // threads per column group, striding over the 64 k rows
        const bool mine = db != nullptr && args.a_mn && (kT ? it.m_unit == 0 : n_blk == 0);
        const int etid = threadIdx.x - 128;  // 0..511
        const int jc = etid % kGroups;
        const int sub = etid / kGroups;
        const int piece = jc >> 3;           // 64-column piece of the operand tile
        float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int kb = kb0; kb < kb1; ++kb) {
          // single CTA: the stage is readable as soon as TMA filled it; CTA pair: only the leader's full
          // barrier is signalled, so wait for the (multicast) "MMA consumed this stage" commit instead
          mbar_wait(CL == 1 ? &full_bar[cs_stage] : &consumed_bar[cs_stage], cs_phase);
          if (mine && sub < kSubs) {
            const uint32_t src = smem_u32(smem) + cs_stage * Cfg::kStageBytes + (kT ? A_STAGE_BYTES : 0) + piece * 8192;
#pragma unroll
            for (int rr = 0; rr < (64 + kSubs - 1) / kSubs; ++rr) {
              const int r = sub + kSubs * rr;
              if (r < 64) {
                const uint4 u = lds128(src + r * 128 + (((jc & 7) ^ (r & 7)) << 4));
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f2 = unpack_bf16x2(w[e]);
                  cs[2 * e] += f2.x;
                  cs[2 * e + 1] += f2.y;
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[cs_stage]);
          if (++cs_stage == kStages) { cs_stage = 0; cs_phase ^= 1; }
        }
        if (mine && sub < kSubs) {
          if (!kT) {
#pragma unroll
            for (int e = 0; e < 8; ++e) cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 16);
            if (lane < 16) {
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (m0 + jc * 8 + e < args.M) atomicAdd(db + m0 + jc * 8 + e, cs[e]);
            }
          } else {
            const int colbase = (piece < 2 ? n0 + cta_rank * 128 + piece * 64 : n0 + 256 + cta_rank * 64) + (jc & 7) * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (colbase + e < args.N) atomicAdd(db + colbase + e, cs[e]);
          }
        }
        mbar_wait(&tmem_full[acc], acc_phase);
        tcgen05_fence_after();
        const long long te0 = args.prof ? clock64() : 0;
        const int row = row0 + lane;
        for (int c = group; c < BN / 32; c += NUM_EPI_GROUPS) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(lane_taddr + static_cast<uint32_t>(acc * BN + c * 32), v);
          tmem_ld_wait();
          if (row < args.M) {
            if (!kT) {
              // handled below (needs the whole warp)
            } else {
              // transposed store: lanes hold consecutive rows m, i.e. consecutive addresses of D[n, :]
              float* dst = args.out_f32 + static_cast<long long>(n0 + c * 32) * args.ldd + row;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + static_cast<long long>(j) * args.ldd),
                             "f"(__uint_as_float(v[j]))
                             : "memory");
            }
          }
          if (!kT) {
            // row-major store: a lane owns a ROW, so reds straight from its registers touch 32 cache lines per
            // instruction (measured 3x slower than the mainloop could hide). Transpose the 32 x 32 fp32 chunk inside
            // the warp with butterfly shuffles (no shared memory: the smem all goes to pipeline stages) so that lane
            // c holds COLUMN c of every row, and issue one coalesced 128-byte red per row.
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) {
              const bool upper = (lane & sft) != 0;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if ((i & sft) == 0) {
                  // lanes with bit `sft` clear keep v[i] and receive the partner's v[i] into v[i + sft];
                  // lanes with the bit set keep v[i + sft] and receive the partner's v[i + sft] into v[i]
                  const uint32_t send = upper ? v[i] : v[i + sft];
                  const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, sft);
                  if (upper) v[i] = recv;
                  else v[i + sft] = recv;
                }
              }
            }
            // now v[r] = element (row r, column lane) of the chunk
            float* dst = args.out_f32 + static_cast<long long>(row0) * args.ldd + n0 + c * 32 + lane;
            const int rows_here = min(32, args.M - row0);
#pragma unroll
            for (int r = 0; r < 32; ++r)
              if (r < rows_here)
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + static_cast<long long>(r) * args.ldd),
                             "f"(__uint_as_float(v[r]))
                             : "memory");
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CL == 1) mbar_arrive(&tmem_empty[acc]);
          else mbar_arrive_leader(&tmem_empty[acc]);
        }
        if (args.prof && threadIdx.x == 128) atomicAdd(args.prof + 6, static_cast<unsigned long long>(clock64() - te0));
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        continue;
      }

      if (LN && n_blk == 0) {
        // ---- LayerNorm transform: this warp normalises 8 of the CTA's 128 rows of the fp32 residual stream into
        // the bf16 A panel (UMMA K-major SW128 layout, k block = 64 columns), optionally saving the normalised
        // rows and the statistics for backward. Lane l owns columns 4l..4l+3 of each 128-column third of the row.
        constexpr int kK = 384;
        static_assert(kMaxPanelKBlocks * BLOCK_K == kK, "the LayerNorm-fused mode is built for D = 384");
        const long long tt0 = args.prof ? clock64() : 0;
        if (ln_unit > 0) mbar_wait(panel_empty, (ln_unit - 1) & 1);  // MMAs of the previous unit are done with the panel
        ++ln_unit;
        const long long tt1 = args.prof ? clock64() : 0;
        float g[12], bt[12];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float4 gv = __ldg(reinterpret_cast<const float4*>(args.ln_gamma + 128 * j) + lane);
          const float4 bv = __ldg(reinterpret_cast<const float4*>(args.ln_beta + 128 * j) + lane);
          g[4 * j] = gv.x; g[4 * j + 1] = gv.y; g[4 * j + 2] = gv.z; g[4 * j + 3] = gv.w;
          bt[4 * j] = bv.x; bt[4 * j + 1] = bv.y; bt[4 * j + 2] = bv.z; bt[4 * j + 3] = bv.w;
        }
        const uint32_t panel_s = smem_u32(panel);
        const int wrow0 = (warp - 4) * 8;  // first tile row of this warp
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float4 xv[4][3];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int grow = m0 + wrow0 + half * 4 + i;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              xv[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (grow < args.M)
                xv[i][j] = __ldg(reinterpret_cast<const float4*>(args.ln_x + static_cast<long long>(grow) * args.ln_ldx + 128 * j) + lane);
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int trow = wrow0 + half * 4 + i;
            const int grow = m0 + trow;
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) sum += (xv[i][j].x + xv[i][j].y) + (xv[i][j].z + xv[i][j].w);
            const float mean = warp_sum(sum) * (1.f / kK);
            float var = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const float a = xv[i][j].x - mean, b = xv[i][j].y - mean, c = xv[i][j].z - mean, d = xv[i][j].w - mean;
              var += (a * a + b * b) + (c * c + d * d);
            }
            const float rstd = rsqrtf(warp_sum(var) * (1.f / kK) + args.ln_eps);
            const bool live = grow < args.M;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const float y0 = (xv[i][j].x - mean) * rstd * g[4 * j] + bt[4 * j];
              const float y1 = (xv[i][j].y - mean) * rstd * g[4 * j + 1] + bt[4 * j + 1];
              const float y2 = (xv[i][j].z - mean) * rstd * g[4 * j + 2] + bt[4 * j + 2];
              const float y3 = (xv[i][j].w - mean) * rstd * g[4 * j + 3] + bt[4 * j + 3];
              const uint32_t p0 = live ? pack_bf16x2(y0, y1) : 0u, p1 = live ? pack_bf16x2(y2, y3) : 0u;
              // column 4 lane + 128 j -> k block 2j + (lane >= 16), 16-byte chunk (lane % 16) / 2, half (lane & 1)
              const int kb = 2 * j + (lane >> 4);
              const uint32_t dst = panel_s + kb * A_STAGE_BYTES + trow * 128 +
                                   (((((lane & 15) >> 1)) ^ (trow & 7)) << 4) + (lane & 1) * 8;
              asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(p0), "r"(p1) : "memory");
              if (live && args.ln_out != nullptr)
                *reinterpret_cast<uint2*>(args.ln_out + static_cast<long long>(grow) * kK + 128 * j + 4 * lane) = make_uint2(p0, p1);
            }
            if (live && lane == 0 && args.ln_mean != nullptr) {
              args.ln_mean[grow] = mean;
              args.ln_rstd[grow] = rstd;
            }
          }
        }
        fence_proxy_async_smem();  // generic-proxy panel writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive_leader_release(panel_full);
        if (args.prof && threadIdx.x == 128) {
          atomicAdd(args.prof + 5, static_cast<unsigned long long>(clock64() - tt1));  // transform proper
          atomicAdd(args.prof + 7, static_cast<unsigned long long>(tt1 - tt0));        // waiting for the panel
        }
      }

      const int c_first = (group - tile_iter) & 3;
      const bool has_bias = EPI != EPI_MUL_AUX && args.bias != nullptr;
      // bias: the fp32-residual epilogue adds it to the aux registers (per-lane float4 of its column piece);
      // the others broadcast it through the slab from one register per lane, loaded before the accumulator wait
      const bool slab_bias = has_bias && EPI != EPI_BIAS_RES_F32;
      float bias_r[kMaxOwn];  // lane l: bias[n0 + c*W + l] of the i-th owned chunk
#pragma unroll
      for (int i = 0; i < kMaxOwn; ++i) {
        const int c = c_first + NUM_EPI_GROUPS * i;
        bias_r[i] = 0.f;
        if (slab_bias && c < kChunks && lane < W) bias_r[i] = __ldg(args.bias + n0 + c * W + lane);
      }

      // stochastic depth: per-row scale of the branch output (this lane's own row, and below the rows of its aux pieces)
      const bool has_rs = EPI == EPI_BIAS_RES_F32 && args.rowscale != nullptr;
      const float rs_own = (has_rs && row0 + lane < args.M) ? __ldg(args.rowscale + row0 + lane) : 1.f;

      const long long te0 = args.prof ? clock64() : 0;
      mbar_wait(&tmem_full[acc], acc_phase);
      if (args.prof && threadIdx.x == 128) atomicAdd(args.prof + 4, static_cast<unsigned long long>(clock64() - te0));
      tcgen05_fence_after();

      bool released = false;
#pragma unroll
      for (int i = 0; i < kMaxOwn; ++i) {
        const int c = c_first + NUM_EPI_GROUPS * i;
        if (c >= kChunks) break;
        const int col0 = n0 + c * W;
        const bool lap_on = args.prof != nullptr && threadIdx.x == 128;
        long long lap_t = lap_on ? clock64() : 0;
        auto lap = [&](int idx) {  // developer instrumentation: cycles of the first epilogue thread per chunk phase
          if (lap_on) {
            const long long now = clock64();
            atomicAdd(args.prof + 8 + idx, static_cast<unsigned long long>(now - lap_t));
            lap_t = now;
          }
        };
        uint32_t v[W];
        if constexpr (W == 32) tmem_ld_32x32b_x32(lane_taddr + static_cast<uint32_t>(acc * BN + c * W), v);
        else tmem_ld_32x32b_x16(lane_taddr + static_cast<uint32_t>(acc * BN + c * W), v);
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (EPI == EPI_BIAS_RES_F32 && has_bias) b4 = __ldg(reinterpret_cast<const float4*>(args.bias + col0) + (lane & 3));
        uint32_t slab = slab_base + (ring & 1) * SLAB_BYTES;
        if (!LNT || lane == 0) tma_store_wait_read<1>();  // the store that last read this slab has drained (every lane asks:
                                                          // a no-op for lanes without bulk groups, and no assumption about WHICH lane
                                                          // elect.sync picks for the stores below)
        __syncwarp();
        lap(0);
        // park this chunk's aux set in the slab (coalesced mapping) and refill the registers two chunks ahead
        auto stage_and_refill = [&](uint4 (&cur)[4]) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint4 a = cur[k];
            if (EPI == EPI_BIAS_RES_F32) {
              float rk = 1.f;  // the bias belongs to the (scaled) branch, not to the residual
              if (has_rs && row0 + k * 8 + (lane >> 2) < args.M) rk = __ldg(args.rowscale + row0 + k * 8 + (lane >> 2));
              a.x = __float_as_uint(fmaf(rk, b4.x, __uint_as_float(a.x)));
              a.y = __float_as_uint(fmaf(rk, b4.y, __uint_as_float(a.y)));
              a.z = __float_as_uint(fmaf(rk, b4.z, __uint_as_float(a.z)));
              a.w = __float_as_uint(fmaf(rk, b4.w, __uint_as_float(a.w)));
            }
            sts128(slab + sw64_offset(k * 8 + (lane >> 2), lane & 3), a);
          }
          prefetch_aux(cur);
        };
        auto stage_aux = [&]() {
          if (aux_par) stage_and_refill(aux_b);
          else stage_and_refill(aux_a);
          aux_par ^= 1;
          __syncwarp();
        };
        if (slab_bias) {
          if (lane < W) sts32(slab + lane * 4, __float_as_uint(bias_r[i]));
          __syncwarp();
        } else if (kAux) {
          stage_aux();  // while the TMEM load is in flight
        }
        tmem_ld_wait();
        lap(1);
        if (c + NUM_EPI_GROUPS >= kChunks) {
          // last TMEM read of this tile by this warp: hand the accumulator stage back (to the leader's MMA)
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CL == 1) mbar_arrive(&tmem_empty[acc]);
            else mbar_arrive_leader(&tmem_empty[acc]);
          }
          released = true;
        }
        float f[W];
#pragma unroll
        for (int j = 0; j < W; ++j) f[j] = __uint_as_float(v[j]);
        if (slab_bias) {
#pragma unroll
          for (int j = 0; j < W / 4; ++j) {
            const uint4 b = lds128(slab + j * 16);  // broadcast read
            f[4 * j] += __uint_as_float(b.x);
            f[4 * j + 1] += __uint_as_float(b.y);
            f[4 * j + 2] += __uint_as_float(b.z);
            f[4 * j + 3] += __uint_as_float(b.w);
          }
          __syncwarp();  // every lane has its bias before the slab is overwritten
          if (kAux) stage_aux();
        }
        if (kAux) {
          // own row back from the slab (conflict-free), combined with the accumulator
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 a = lds128(slab + sw64_offset(lane, j));
            const uint32_t w[4] = {a.x, a.y, a.z, a.w};
            if (kAuxF32) {
#pragma unroll
              for (int e = 0; e < 4; ++e) f[(4 * j + e) % W] = fmaf(f[(4 * j + e) % W], rs_own, __uint_as_float(w[e]));
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 x = unpack_bf16x2(w[e]);
                if (EPI == EPI_BIAS_RES) {
                  f[(8 * j + 2 * e) % W] += x.x;
                  f[(8 * j + 2 * e + 1) % W] += x.y;
                } else {
                  f[(8 * j + 2 * e) % W] *= x.x;
                  f[(8 * j + 2 * e + 1) % W] *= x.y;
                }
              }
            }
          }
          // each lane re-uses only its own row below, so no further warp sync is needed before the writes
        }
        lap(2);
        const bool rows_ok = row0 < args.M;
        if (EPI == EPI_BIAS_RES_F32) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(slab + sw64_offset(lane, j),
                   make_uint4(__float_as_uint(f[(4 * j) % W]), __float_as_uint(f[(4 * j + 1) % W]),
                              __float_as_uint(f[(4 * j + 2) % W]), __float_as_uint(f[(4 * j + 3) % W])));
          lap(3);
          fence_proxy_async_smem();
          __syncwarp();
          lap(4);
          if (LNT) {
            if (lane == 0) {   // LayerNorm tail: one group per chunk ALWAYS, on lane 0 (its bookkeeping counts groups)
              if (rows_ok) tma_store_2d_s(&tmD, slab, col0, row0);
              tma_store_commit();
            }
          } else if (rows_ok && elect_one_sync()) {   // elect.sync: the compiler emits UTMASTG straight (see the producer)
            tma_store_2d_s(&tmD, slab, col0, row0);
            tma_store_commit();
          }
          ++ring;
          lap(5);
          if (lap_on) atomicAdd(args.prof + 15, 1ull);
          continue;
        }
        if (EPI == EPI_BIAS_GELU_FWD) {
#pragma unroll
          for (int j = 0; j < W / 2; ++j) unpack_f32x2(gelu2(f[2 * j], f[2 * j + 1]), f[2 * j], f[2 * j + 1]);
        }
        uint32_t hpk[EPI == EPI_BIAS_GELU ? W / 2 : 1];  // gelu(x) packed: the second output
        if (EPI == EPI_BIAS_GELU) {
#pragma unroll
          for (int j = 0; j < W / 2; ++j) {
            f32x2 h2, g2;
            gelu_and_grad2(f[2 * j], f[2 * j + 1], h2, g2);
            float h0, h1;
            unpack_f32x2(h2, h0, h1);
            unpack_f32x2(g2, f[2 * j], f[2 * j + 1]);
            hpk[j % (EPI == EPI_BIAS_GELU ? W / 2 : 1)] = pack_bf16x2(h0, h1);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts128(slab + sw64_offset(lane, j),
                 make_uint4(pack_bf16x2(f[(8 * j) % W], f[(8 * j + 1) % W]), pack_bf16x2(f[(8 * j + 2) % W], f[(8 * j + 3) % W]),
                            pack_bf16x2(f[(8 * j + 4) % W], f[(8 * j + 5) % W]), pack_bf16x2(f[(8 * j + 6) % W], f[(8 * j + 7) % W])));
        lap(3);
        fence_proxy_async_smem();
        __syncwarp();
        lap(4);
        if (rows_ok && elect_one_sync()) {
          tma_store_2d_s(&tmD, slab, col0, row0);
          tma_store_commit();
        }
        ++ring;
        lap(5);
        if (lap_on) atomicAdd(args.prof + 15, 1ull);
        if (EPI == EPI_BIAS_GELU) {
          slab = slab_base + (ring & 1) * SLAB_BYTES;
          if (!LNT || lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          constexpr int HM = EPI == EPI_BIAS_GELU ? W / 2 : 1;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(slab + sw64_offset(lane, j),
                   make_uint4(hpk[(4 * j) % HM], hpk[(4 * j + 1) % HM], hpk[(4 * j + 2) % HM], hpk[(4 * j + 3) % HM]));
          fence_proxy_async_smem();
          __syncwarp();
          if (rows_ok && elect_one_sync()) {
            tma_store_2d_s(&tmD2, slab, col0, row0);
            tma_store_commit();
          }
          ++ring;
        }
      }
      if (!released) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CL == 1) mbar_arrive(&tmem_empty[acc]);
          else mbar_arrive_leader(&tmem_empty[acc]);
        }
      }
      if (LNT && n_blk == args.num_n_blocks - 1) {
        // ---- LayerNorm tail, epilogue side: this warp has now ISSUED its stores of m unit `ln_unit` (six groups: empty
        // groups are committed for row groups past M). Everything older has landed once at most those six are pending:
        // tell the helper warps that the PREVIOUS unit's rows are in memory (a unit of slack: no stall here).
        if (ln_unit > 0 && lane == 0) {
          const int n = ln_unit - 1;
          tma_store_wait_all<6>();
          if (n >= 2) mbar_wait(&ln_done[n & 1], ((n - 2) >> 1) & 1);   // the helpers are done with this barrier's last use
          mbar_arrive(&ln_go[n & 1]);
        }
        ++ln_unit;
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
    if (!LNT || lane == 0) tma_store_wait_all<0>();
    if (LNT && ln_unit > 0 && lane == 0) {
      const int n = ln_unit - 1;   // the last unit
      if (n >= 2) mbar_wait(&ln_done[n & 1], ((n - 2) >> 1) & 1);
      mbar_arrive(&ln_go[n & 1]);
    }
  } else if (LNT && warp >= 2) {
    // ------------------------------------------------------------------ LayerNorm tail, helper warps 2 and 3
    // One m unit behind the epilogue: normalise this CTA's 128 finished rows (64 per warp) of the fp32 stream -- read back
    // from L2, where they were written a unit ago -- into bf16 LayerNorm rows + statistics. The epilogue warps, the
    // bottleneck of these HBM-bound GEMMs, spend nothing on it.
    constexpr int kK = 384;
    int n = 0;
    for (int t = cluster_id; t < num_m_units; t += num_clusters, ++n) {
      const int r_base = (t * CL + cta_rank) * BLOCK_M + (warp - 2) * 64;
      mbar_wait(&ln_go[n & 1], (n >> 1) & 1);
#pragma unroll 1
      for (int rr = 0; rr < 64; rr += 4) {
        float4 xv[4][3];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int grow = r_base + rr + i;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            xv[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (grow < args.M)
              xv[i][j] = ld_global_cg_f4(args.out_f32 + static_cast<long long>(grow) * args.ldd + 128 * j + 4 * lane);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int grow = r_base + rr + i;
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < 3; ++j) sum += (xv[i][j].x + xv[i][j].y) + (xv[i][j].z + xv[i][j].w);
          const float mean = warp_sum(sum) * (1.f / kK);
          float var = 0.f;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float a = xv[i][j].x - mean, b = xv[i][j].y - mean, c = xv[i][j].z - mean, d = xv[i][j].w - mean;
            var += (a * a + b * b) + (c * c + d * d);
          }
          const float rstd = rsqrtf(warp_sum(var) * (1.f / kK) + args.ln_eps);
          if (grow < args.M) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const float4 gv = __ldg(reinterpret_cast<const float4*>(args.ln_gamma + 128 * j) + lane);
              const float4 bv = __ldg(reinterpret_cast<const float4*>(args.ln_beta + 128 * j) + lane);
              const float y0 = (xv[i][j].x - mean) * rstd * gv.x + bv.x;
              const float y1 = (xv[i][j].y - mean) * rstd * gv.y + bv.y;
              const float y2 = (xv[i][j].z - mean) * rstd * gv.z + bv.z;
              const float y3 = (xv[i][j].w - mean) * rstd * gv.w + bv.w;
              *reinterpret_cast<uint2*>(args.ln_out + static_cast<long long>(grow) * kK + 128 * j + 4 * lane) =
                  make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
            }
            if (lane == 0 && args.ln_mean != nullptr) {
              args.ln_mean[grow] = mean;
              args.ln_rstd[grow] = rstd;
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ln_done[n & 1]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // the peer may still arrive on / read from this CTA until it is done
  if (args.prof && threadIdx.x == 32 && !LN) {
    atomicAdd(args.prof + 5, static_cast<unsigned long long>(clock64() - t_entry));
    atomicAdd(args.prof + 7, 1ull);
  }
  if (warp == 2) {
    if (CL == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
static int g_gemm_cluster = 2;     // CTAs per cluster (1 disables the CTA-pair mode)
static int g_gemm_stationary = 1;  // 0 disables the B-stationary mode (developer A/B switch)
static int g_gemm_wide = 1;        // 0 disables the automatic choice of 256 x 384 CTA-pair tiles
static unsigned long long* g_gemm_prof = nullptr;

template <int BN, int EPI, int CL, int MODE>
static int launch_gemm_cl(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                          const CUtensorMap& tmD2, const GemmArgs& args, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CL, MODE, is_wgrad_epi(EPI)>;
  static bool configured = false;
  if (!configured) {
    B200SSL_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, EPI, CL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::kSmemBytes));
    configured = true;
  }
  const int units = ((args.num_m_blocks + CL - 1) / CL) * args.num_n_blocks * args.k_splits;
  const int max_clusters = sm_count() / CL;
  const int clusters = units < max_clusters ? units : max_clusters;
  B200SSL_CUDA(launch_pdl(gemm_kernel<BN, EPI, CL, MODE>, dim3(clusters * CL), dim3(GEMM_THREADS), Cfg::kSmemBytes, stream,
                          CL, tmA, tmB, tmD, tmD2, args));
  return 0;
}

// mode: 1 = independent CTAs, 2 = CTA pairs, 3 = CTA pairs with a stationary B tile (K <= 384)
template <int BN, int EPI>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                       const CUtensorMap& tmD2, const GemmArgs& args, int mode, cudaStream_t stream) {
  if constexpr (BN == 384) {
    // 256 x 384 CTA-pair tiles: wgrad (both store orders), plain dgrad and the fp32-residual fprop
    if constexpr (EPI == EPI_BIAS || EPI == EPI_ATOMIC_F32 || EPI == EPI_ATOMIC_F32_T || EPI == EPI_BIAS_RES_F32) {
      if (mode >= 2) return launch_gemm_cl<BN, EPI, 2, 0>(tmA, tmB, tmD, tmD2, args, stream);
    }
    set_last_error("gemm: block_n 384 needs the CTA-pair mode and epilogue 0, 4, 5 or 6 (got %d)", EPI);
    return -2;
  } else if constexpr (EPI == EPI_ATOMIC_F32_T) {
    set_last_error("gemm: epilogue 6 (transposed wgrad store) needs block_n 384");
    return -2;
  } else {
    if constexpr (!is_wgrad_epi(EPI) && (BN == 192 || BN == 256)) {
      if (mode == 3) return launch_gemm_cl<BN, EPI, 2, 1>(tmA, tmB, tmD, tmD2, args, stream);
    }
    if (mode >= 2) return launch_gemm_cl<BN, EPI, 2, 0>(tmA, tmB, tmD, tmD2, args, stream);
    return launch_gemm_cl<BN, EPI, 1, 0>(tmA, tmB, tmD, tmD2, args, stream);
  }
}

template <int BN>
static int dispatch_epi(int epi, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& d,
                        const CUtensorMap& d2, const GemmArgs& args, int cl, cudaStream_t s) {
  switch (epi) {
    case EPI_BIAS: return launch_gemm<BN, EPI_BIAS>(a, b, d, d2, args, cl, s);
    case EPI_BIAS_GELU: return launch_gemm<BN, EPI_BIAS_GELU>(a, b, d, d2, args, cl, s);
    case EPI_BIAS_RES: return launch_gemm<BN, EPI_BIAS_RES>(a, b, d, d2, args, cl, s);
    case EPI_MUL_AUX: return launch_gemm<BN, EPI_MUL_AUX>(a, b, d, d2, args, cl, s);
    case EPI_ATOMIC_F32: return launch_gemm<BN, EPI_ATOMIC_F32>(a, b, d, d2, args, cl, s);
    case EPI_BIAS_RES_F32: return launch_gemm<BN, EPI_BIAS_RES_F32>(a, b, d, d2, args, cl, s);
    case EPI_ATOMIC_F32_T: return launch_gemm<BN, EPI_ATOMIC_F32_T>(a, b, d, d2, args, cl, s);
    case EPI_BIAS_GELU_FWD: return launch_gemm<BN, EPI_BIAS_GELU_FWD>(a, b, d, d2, args, cl, s);
  }
  set_last_error("gemm: unknown epilogue %d", epi);
  return -2;
}

static int pick_block_n(int N, int M, int k_splits) {
  // largest tile that divides N; fall back to smaller tiles when the grid would not fill the SMs
  const int cands[4] = {256, 192, 128, 64};
  int best = 0;
  for (int i = 0; i < 4; ++i) {
    if (N % cands[i]) continue;
    if (!best) best = cands[i];
    const long long tiles = static_cast<long long>((M + BLOCK_M - 1) / BLOCK_M) * (N / cands[i]) * k_splits;
    if (tiles >= sm_count()) return cands[i];
  }
  // nothing fills the machine: use the smallest dividing tile for the most parallelism
  for (int i = 3; i >= 0; --i)
    if (N % cands[i] == 0) return cands[i];
  return best;
}

}  // namespace b200ssl

using namespace b200ssl;

extern "C" int b200ssl_gemm(const void* A, long long lda, int a_mn_major, const void* B,
                            long long ldb, int b_mn_major, void* D, long long ldd, void* D2,
                            const float* bias, const void* aux, long long ldaux, int M, int N, int K,
                            int epilogue, int split_k, int block_n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200SSL_CHECK(M > 0 && N > 0 && K > 0, -2, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  B200SSL_CHECK(N % 64 == 0, -2, "gemm: N=%d must be a multiple of 64", N);
  B200SSL_CHECK(lda % 8 == 0 && ldb % 8 == 0, -2, "gemm: lda/ldb must be multiples of 8 elements");
  B200SSL_CHECK((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(D) & 15) == 0,
                -2, "gemm: operands must be 16-byte aligned");
  B200SSL_CHECK(epilogue >= 0 && epilogue <= 7, -2, "gemm: unknown epilogue %d", epilogue);
  const bool wgrad = is_wgrad_epi(epilogue);
  if (epilogue == EPI_BIAS_RES || epilogue == EPI_MUL_AUX)
    B200SSL_CHECK(aux != nullptr && ldaux % 8 == 0 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0, -2,
                  "gemm: epilogue %d needs a 16B-aligned aux operand", epilogue);
  if (epilogue == EPI_BIAS_RES_F32)
    B200SSL_CHECK(aux != nullptr && ldaux % 4 == 0 && ldd % 4 == 0 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0, -2,
                  "gemm: fp32 residual epilogue needs a 16B-aligned fp32 aux operand and ldd %% 4 == 0");
  if (epilogue == EPI_BIAS_GELU) B200SSL_CHECK(D2 != nullptr, -2, "gemm: GELU epilogue needs D2");
  if (wgrad) {
    B200SSL_CHECK(ldd % 4 == 0, -2, "gemm: fp32 ldd must be a multiple of 4");
  } else {
    if (epilogue != EPI_BIAS_RES_F32) B200SSL_CHECK(ldd % 8 == 0, -2, "gemm: bf16 ldd must be a multiple of 8");
    split_k = 1;
  }
  if (bias && !wgrad)
    B200SSL_CHECK((reinterpret_cast<uintptr_t>(bias) & 15) == 0, -2, "gemm: bias must be 16B aligned");

  const int k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  int bn = block_n;
  const bool pair_ok = g_gemm_cluster == 2 && M > BLOCK_M;
  // 256 x 384 CTA-pair tiles (one accumulator, no mainloop/epilogue overlap) halve the shared-memory traffic per
  // MMA: used for wgrad (one tile per CTA anyway), for long-K problems with a plain epilogue and for every plain dgrad
  // with 384 outputs (proj dgrad, K = 384: 75.9 -> 67.6 us at 195,584 rows, tests/gpu_checks/proj_gemm_ab.py)
  if (bn == 0 && N % 384 == 0 && pair_ok &&
      (epilogue == EPI_ATOMIC_F32_T ||
       (g_gemm_wide && ((epilogue == EPI_ATOMIC_F32 && M >= 512) ||
                        (epilogue == EPI_BIAS && (K >= 768 || (b_mn_major && K >= 384)))))))
    bn = 384;
  if (epilogue == EPI_ATOMIC_F32_T) B200SSL_CHECK(bn == 384, -2, "gemm: epilogue 6 needs N %% 384 == 0 and M > 128");
  if (wgrad && split_k <= 0) {
    // auto split-K: the largest tile that divides N, then as many K splits as fill one wave of SMs
    if (bn == 0) bn = N % 256 == 0 ? 256 : N % 192 == 0 ? 192 : N % 128 == 0 ? 128 : 64;
    const int cl_ = (g_gemm_cluster == 2 && M > BLOCK_M && (!b_mn_major || bn % 128 == 0 || bn == 384)) ? 2 : 1;
    const int tiles_mn = (((M + BLOCK_M - 1) / BLOCK_M + cl_ - 1) / cl_) * (N / bn);  // work units per k split
    split_k = (sm_count() / cl_) / (tiles_mn > 0 ? tiles_mn : 1);
    if (split_k > k_blocks / 2) split_k = k_blocks / 2;
  }
  if (split_k < 1) split_k = 1;
  if (split_k > k_blocks) split_k = k_blocks;
  int kbps = (k_blocks + split_k - 1) / split_k;
  split_k = (k_blocks + kbps - 1) / kbps;  // no empty splits

  if (bn == 0) bn = pick_block_n(N, M, split_k);
  B200SSL_CHECK((bn == 64 || bn == 128 || bn == 192 || bn == 256 || bn == 384) && N % bn == 0, -2,
                "gemm: block_n=%d does not divide N=%d", bn, N);
  B200SSL_CHECK(bn != 384 || pair_ok, -2, "gemm: block_n 384 needs the CTA-pair mode (M > 128)");

  GemmArgs args;
  args.M = M; args.N = N; args.K = K;
  args.a_mn = a_mn_major; args.b_mn = b_mn_major;
  args.num_m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
  args.num_n_blocks = N / bn;
  args.k_splits = split_k;
  args.k_blocks_per_split = kbps;
  args.k_blocks_total = k_blocks;
  args.bias = bias;
  args.aux = aux;
  args.ldaux = ldaux;
  args.out_f32 = static_cast<float*>(D);
  args.ldd = ldd;
  args.prof = g_gemm_prof;
  args.rowscale = epilogue == EPI_BIAS_RES_F32 ? static_cast<const float*>(D2) : nullptr;
  args.ln_x = nullptr; args.ln_ldx = 0; args.ln_gamma = nullptr; args.ln_beta = nullptr; args.ln_eps = 0.f;
  args.ln_out = nullptr; args.ln_mean = nullptr; args.ln_rstd = nullptr;

  // CTA-pair mode needs two m blocks and, for an MN-major B, a half tile made of whole 64-column chunks
  int cluster = (g_gemm_cluster == 2 && args.num_m_blocks > 1 && (!b_mn_major || bn % 128 == 0 || bn == 384)) ? 2 : 1;
  // B-stationary: short K (the whole B tile fits beside the A ring), enough clusters to give every n block one
  if (cluster == 2 && g_gemm_stationary && !wgrad && k_blocks <= kMaxPanelKBlocks &&
      (bn == 192 || bn == 256) && args.num_n_blocks <= sm_count() / 2 && args.num_m_blocks >= 8)
    cluster = 3;
  CUtensorMap tmA, tmB, tmD, tmD2;
  {
    // A: K-major -> dims (K, M), box (64, 128); MN-major -> dims (M, K), box (64, 64)
    uint64_t dims[2], strides[2];
    uint32_t box[2];
    if (!a_mn_major) { dims[0] = K; dims[1] = M; box[0] = 64; box[1] = BLOCK_M; }
    else             { dims[0] = M; dims[1] = K; box[0] = 64; box[1] = 64; }
    strides[0] = 2; strides[1] = static_cast<uint64_t>(lda) * 2;
    if (int rc = make_tensor_map(&tmA, A, 2, 2, dims, strides, box, 128)) return rc;
    if (!b_mn_major) { dims[0] = K; dims[1] = N; box[0] = 64; box[1] = bn == 384 ? 64 : bn / (cluster >= 2 ? 2 : 1); }  // pair: half the rows
    else             { dims[0] = N; dims[1] = K; box[0] = 64; box[1] = 64; }
    strides[1] = static_cast<uint64_t>(ldb) * 2;
    if (int rc = make_tensor_map(&tmB, B, 2, 2, dims, strides, box, 128)) return rc;
    if (epilogue == EPI_BIAS_RES_F32) {
      dims[0] = N; dims[1] = M; box[0] = 16; box[1] = 32;               // per-warp slab: 32 rows x 64 B
      strides[0] = 4; strides[1] = static_cast<uint64_t>(ldd) * 4;
      if (int rc = make_tensor_map(&tmD, D, 4, 2, dims, strides, box, 64)) return rc;
      tmD2 = tmD;
    } else if (!wgrad) {
      dims[0] = N; dims[1] = M; box[0] = 32; box[1] = 32;               // per-warp slab: 32 rows x 64 B, 64B swizzle
      strides[1] = static_cast<uint64_t>(ldd) * 2;
      if (int rc = make_tensor_map(&tmD, D, 2, 2, dims, strides, box, 64)) return rc;
      if (int rc = make_tensor_map(&tmD2, D2 ? D2 : D, 2, 2, dims, strides, box, 64)) return rc;
    } else {
      tmD = tmA; tmD2 = tmA;  // unused
    }
  }

  switch (bn) {
    case 64: return dispatch_epi<64>(epilogue, tmA, tmB, tmD, tmD2, args, cluster, stream);
    case 128: return dispatch_epi<128>(epilogue, tmA, tmB, tmD, tmD2, args, cluster, stream);
    case 192: return dispatch_epi<192>(epilogue, tmA, tmB, tmD, tmD2, args, cluster, stream);
    case 384: return dispatch_epi<384>(epilogue, tmA, tmB, tmD, tmD2, args, cluster, stream);
    default: return dispatch_epi<256>(epilogue, tmA, tmB, tmD, tmD2, args, cluster, stream);
  }
}

// LayerNorm fused into the GEMM that consumes it (Block.norm1 -> attn.qkv, Block.norm2 -> mlp.fc1; VT.pyc@L147,151):
//   D = epilogue( LN(x) W^T + bias ),  x fp32 [M, 384] (the residual stream), W bf16 [N, 384].
// The epilogue warps of each CTA pair normalise 2 x 128 rows straight into bf16 A panels in the UMMA layout; every n
// block of those rows is computed from the panel while only W streams through the smem ring. ln_out (bf16 [M, 384],
// contiguous) / mean / rstd (fp32 [M]) are optional side outputs for backward. Epilogues 0, 1, 7.
extern "C" int b200ssl_ln_gemm(const float* x, long long ldx, const float* gamma, const float* beta, float eps,
                               void* ln_out, float* mean, float* rstd, const void* W, long long ldw, void* D,
                               long long ldd, void* D2, const float* bias, int M, int N, int K, int epilogue,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200SSL_CHECK(K == kMaxPanelKBlocks * BLOCK_K, -2, "ln_gemm: K=%d unsupported (the fused mode is built for D = 384)", K);
  B200SSL_CHECK(M > BLOCK_M, -2, "ln_gemm: needs more than 128 rows (CTA-pair tiles), got %d", M);
  B200SSL_CHECK(epilogue == EPI_BIAS || epilogue == EPI_BIAS_GELU || epilogue == EPI_BIAS_GELU_FWD, -2,
                "ln_gemm: epilogue %d unsupported (0, 1 or 7)", epilogue);
  B200SSL_CHECK(N % 192 == 0 || N % 256 == 0, -2, "ln_gemm: N=%d must be a multiple of 192 or 256", N);
  B200SSL_CHECK(ldx % 4 == 0 && ldw % 8 == 0 && ldd % 8 == 0, -2, "ln_gemm: ldx %% 4, ldw %% 8, ldd %% 8 must be 0");
  B200SSL_CHECK(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(D) |
                  reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
                  reinterpret_cast<uintptr_t>(ln_out) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0,
                -2, "ln_gemm: operands must be 16-byte aligned");
  B200SSL_CHECK((mean == nullptr) == (rstd == nullptr), -2, "ln_gemm: mean and rstd go together");
  if (epilogue == EPI_BIAS_GELU) B200SSL_CHECK(D2 != nullptr, -2, "ln_gemm: GELU epilogue needs D2");
  const int bn = N % 256 == 0 ? 256 : 192;

  GemmArgs args;
  args.M = M; args.N = N; args.K = K;
  args.a_mn = 0; args.b_mn = 0;
  args.num_m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
  args.num_n_blocks = N / bn;
  args.k_splits = 1;
  args.k_blocks_per_split = kMaxPanelKBlocks;
  args.k_blocks_total = kMaxPanelKBlocks;
  args.bias = bias;
  args.aux = nullptr; args.ldaux = 0;
  args.out_f32 = nullptr;
  args.ldd = ldd;
  args.prof = g_gemm_prof;
  args.rowscale = nullptr;
  args.ln_x = x; args.ln_ldx = ldx; args.ln_gamma = gamma; args.ln_beta = beta; args.ln_eps = eps;
  args.ln_out = static_cast<__nv_bfloat16*>(ln_out); args.ln_mean = mean; args.ln_rstd = rstd;

  CUtensorMap tmB, tmD, tmD2;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t strides[2] = {2, static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2] = {64, static_cast<uint32_t>(bn / 2)};
    if (int rc = make_tensor_map(&tmB, W, 2, 2, dims, strides, box, 128)) return rc;
    dims[0] = N; dims[1] = M; box[0] = 32; box[1] = 32;
    strides[1] = static_cast<uint64_t>(ldd) * 2;
    if (int rc = make_tensor_map(&tmD, D, 2, 2, dims, strides, box, 64)) return rc;
    if (int rc = make_tensor_map(&tmD2, D2 ? D2 : D, 2, 2, dims, strides, box, 64)) return rc;
  }
  if (bn == 256) {
    switch (epilogue) {
      case EPI_BIAS: return launch_gemm_cl<256, EPI_BIAS, 2, 2>(tmB, tmB, tmD, tmD2, args, stream);
      case EPI_BIAS_GELU: return launch_gemm_cl<256, EPI_BIAS_GELU, 2, 2>(tmB, tmB, tmD, tmD2, args, stream);
      default: return launch_gemm_cl<256, EPI_BIAS_GELU_FWD, 2, 2>(tmB, tmB, tmD, tmD2, args, stream);
    }
  }
  switch (epilogue) {
    case EPI_BIAS: return launch_gemm_cl<192, EPI_BIAS, 2, 2>(tmB, tmB, tmD, tmD2, args, stream);
    case EPI_BIAS_GELU: return launch_gemm_cl<192, EPI_BIAS_GELU, 2, 2>(tmB, tmB, tmD, tmD2, args, stream);
    default: return launch_gemm_cl<192, EPI_BIAS_GELU_FWD, 2, 2>(tmB, tmB, tmD, tmD2, args, stream);
  }
}

// Residual GEMM with a LayerNorm tail (MODE 3):  D = rowscale * (A W^T + bias) + aux  on the fp32 residual stream, and
// ln_out = LayerNorm(D; gamma, beta, eps) in bf16 (+ per-row mean / rstd, nullable) from the same kernel. For N = 384
// (ViT-S): attn.proj + Block.norm2 and mlp.fc2 + the next Block.norm1 (VT.pyc@L147-151). A [M,K] bf16, W [384,K] bf16,
// aux / D fp32 [M,384], ln_out bf16 [M,384] contiguous. rowscale (stochastic depth) and bias are nullable.
extern "C" int b200ssl_gemm_res_ln(const void* A, long long lda, const void* W, long long ldw, float* D, long long ldd,
                                   const float* rowscale, const float* bias, const float* aux, long long ldaux, int M,
                                   int N, int K, const float* gamma, const float* beta, float eps, void* ln_out,
                                   float* mean, float* rstd, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200SSL_CHECK(N == 384, -2, "gemm_res_ln: N=%d unsupported (the LayerNorm tail is built for D = 384)", N);
  B200SSL_CHECK(M > BLOCK_M && K > 0 && g_gemm_cluster == 2, -2, "gemm_res_ln: needs more than 128 rows and the CTA-pair mode");
  B200SSL_CHECK(lda % 8 == 0 && ldw % 8 == 0 && ldaux % 4 == 0 && ldd % 4 == 0, -2, "gemm_res_ln: lda/ldw %% 8, ldaux/ldd %% 4 must be 0");
  B200SSL_CHECK(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(D) |
                  reinterpret_cast<uintptr_t>(aux) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
                  reinterpret_cast<uintptr_t>(ln_out) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0,
                -2, "gemm_res_ln: operands must be 16-byte aligned");
  B200SSL_CHECK(aux != nullptr && gamma != nullptr && beta != nullptr && ln_out != nullptr, -2, "gemm_res_ln: aux, gamma, beta, ln_out are required");
  B200SSL_CHECK((mean == nullptr) == (rstd == nullptr), -2, "gemm_res_ln: mean and rstd go together");
  constexpr int bn = 192;
  const int k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  GemmArgs args;
  args.M = M; args.N = N; args.K = K;
  args.a_mn = 0; args.b_mn = 0;
  args.num_m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
  args.num_n_blocks = N / bn;
  args.k_splits = 1;
  args.k_blocks_per_split = k_blocks;
  args.k_blocks_total = k_blocks;
  args.bias = bias;
  args.aux = aux;
  args.ldaux = ldaux;
  args.out_f32 = D;
  args.ldd = ldd;
  args.prof = g_gemm_prof;
  args.rowscale = rowscale;
  args.ln_x = nullptr; args.ln_ldx = 0;
  args.ln_gamma = gamma; args.ln_beta = beta; args.ln_eps = eps;
  args.ln_out = static_cast<__nv_bfloat16*>(ln_out); args.ln_mean = mean; args.ln_rstd = rstd;
  CUtensorMap tmA, tmB, tmD;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t strides[2] = {2, static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {64, BLOCK_M};
    if (int rc = make_tensor_map(&tmA, A, 2, 2, dims, strides, box, 128)) return rc;
    dims[1] = N; strides[1] = static_cast<uint64_t>(ldw) * 2; box[1] = bn / 2;   // pair: each CTA loads half of the B tile
    if (int rc = make_tensor_map(&tmB, W, 2, 2, dims, strides, box, 128)) return rc;
    dims[0] = N; dims[1] = M; box[0] = 16; box[1] = 32;                          // per-warp slab: 32 rows x 64 B
    strides[0] = 4; strides[1] = static_cast<uint64_t>(ldd) * 4;
    if (int rc = make_tensor_map(&tmD, D, 4, 2, dims, strides, box, 64)) return rc;
  }
  return launch_gemm_cl<bn, EPI_BIAS_RES_F32, 2, 3>(tmA, tmB, tmD, tmD, args, stream);
}

// 1 = independent CTAs, 2 = clusters of two CTAs sharing the B tile through TMA multicast (default).
extern "C" int b200ssl_set_gemm_cluster(int ctas) {
  B200SSL_CHECK(ctas == 1 || ctas == 2, -2, "gemm cluster size must be 1 or 2");
  b200ssl::g_gemm_cluster = ctas;
  return 0;
}

// Developer instrumentation: device buffer of 16 uint64 counters the GEMM kernels add to (null = off); [8..13] are
// per-chunk phases of the first epilogue thread (wait for a free slab, TMEM load, bias/aux combine, pack + st.shared,
// proxy fence, TMA store issue) and [15] the number of chunks it processed:
// [0] MMA-issuer cycles waiting for operand stages, [1] waiting for a free accumulator, [2] MMA-issuer loop cycles,
// [3] number of issuing CTAs, [4] cycles the first epilogue warp waited for accumulators.
extern "C" int b200ssl_set_gemm_prof(void* counters) {
  b200ssl::g_gemm_prof = static_cast<unsigned long long*>(counters);
  return 0;
}

// 1 (default) = pick 256 x 384 CTA-pair tiles automatically where they apply; 0 = only when block_n = 384 is asked for.
extern "C" int b200ssl_set_gemm_wide(int on) {
  b200ssl::g_gemm_wide = on ? 1 : 0;
  return 0;
}

// 1 = keep the B tile stationary in shared memory for K <= 384 problems (default), 0 = always stream it.
extern "C" int b200ssl_set_gemm_stationary(int on) {
  b200ssl::g_gemm_stationary = on ? 1 : 0;
  return 0;
}
