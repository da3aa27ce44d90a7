// K4: fused multi-head attention, forward and backward, on tcgen05/TMEM fed by TMA.
//
// Replaces the materialised-score attention of the reference (VT.pyc@L119-131:
// qkv.reshape.permute -> (q @ k^T) * scale -> softmax -> attn @ v -> transpose.reshape) for
// head_dim 64: sequences of up to 256 tokens in one fused launch (two-tile kernels below; 197 / 37 tokens for the
// 224^2 / 96^2 crops), longer ones through the streaming forward kernel and a block-pair backward (257 tokens for the
// reference's native 256^2 tiles, 785 / 1025 for ViT-S/8).
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

constexpr int TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16

struct AttnArgs {
  int B, N, H;       // batch, query tokens per sequence handled by this launch, heads
  int Nk;            // key tokens handled by this launch (== N except for the block-decomposed long-sequence path)
  int Ns;            // tokens per sequence in memory (row stride of lse2 / out / dout); == N on the main path
  int acc_dq, acc_dkv;  // bwd, long-sequence path: add this launch's dQ / dK,dV to what is already there
  int nblk;             // bwd, NT == 2: > 1 = ONE launch over all (query block, key block) pairs of 256 x 256 tokens of
                        // every (sequence, head); N / Nk / keys_n are then per item and Ns is the sequence length
  int G;             // sequences packed per 128-row tile (NT == 1), else 1
  int dephase;       // fwd two-tile kernel: event-driven MMA issue with the two slots half a period apart
  int rows;          // valid rows per tile group: G*N (NT == 1) or N (NT == 2)
  int keys_n;        // round_up(rows, 16): MMA N extent over keys
  float scale_log2;  // softmax scale * log2(e)
  float scale;
  float* lse2;
  const __nv_bfloat16* out;   // bwd only
  const __nv_bfloat16* dout;  // bwd only
  const __nv_bfloat16* qkv_q; // bwd only: qkv at the first query / key row of this launch's block (L2 prefetch of the next
  const __nv_bfloat16* qkv_k; //           items' operand rows by the otherwise idle warps)
  __nv_bfloat16* dq;          // bwd only: dqkv at the first query row of this launch's block ([B, Ns, 3, H, 64] layout)
  __nv_bfloat16* dkv;         // bwd only: dqkv at the first key row of this launch's block
  int dq_rows, dkv_rows;      // bwd only: query / key rows that exist behind dq / dkv (rows past them are not stored)
  unsigned long long* prof;   // developer instrumentation (null = off): per-phase cycle counters of one softmax thread
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
    hi = a.Nk;
  }
}

// ------------------------------------------------------------------------------------------------
// forward: persistent CTAs (one per SM), two query tiles in flight
// ------------------------------------------------------------------------------------------------
// A "round" is two 128-row query tiles ("slots"): for 128 < N <= 256 the two tiles of one (sequence, head)
// sharing K/V; for N <= 128 two consecutive (packed group, head) items with their own K/V. Slot s is
// owned by softmax group s (8 warps, two threads per query row) and TMEM columns [256 s, 256 s + 256).
//   warp 16 : TMA producer — prefetches round r+1 into the other 96 KB smem buffer while round r computes
//   warp 17 : tcgen05.mma issuer — S = Q K^T for both slots, then P V per slot as its softmax finishes
//   warps 0-15: softmax + epilogue. Two passes over S straight from TMEM (max, then exp2 / sum / bf16
//             pack); P overwrites the first columns of S and is the A operand of the second MMA (read from
//             TMEM); O accumulates in the last 64 columns of the slot, the dead tail of S.
// smem per buffer: NT==2: Q0 | Q1 | K (2 tiles) | V (2 tiles); NT==1: per slot Q | K | V. The Q tile of a
// slot is re-used as the staging tile of its TMA output store.
constexpr int FWD_THREADS = 18 * 32;
constexpr int FWD_BUF_BYTES = 6 * TILE_BYTES;

template <int NT>
__global__ void __launch_bounds__(FWD_THREADS, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmKV,
                     const __grid_constant__ CUtensorMap tmO, const AttnArgs args, const int num_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * FWD_BUF_BYTES);
  uint64_t* load_full = bars;       // [2] per smem buffer
  uint64_t* buf_free = bars + 2;    // [2] per smem buffer (both groups' stores drained)
  uint64_t* bar_s = bars + 4;       // [2] per slot: S ready
  uint64_t* bar_p = bars + 6;       // [2] per slot: P written to TMEM
  uint64_t* bar_o = bars + 8;       // [2] per slot: O ready
  uint64_t* tmem_free = bars + 10;  // [2] per slot: O read back, slot's TMEM reusable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* xchg_all = reinterpret_cast<float*>(bars + 14);  // [2 groups][2 (max,sum)][2 halves][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rounds: NT==2 -> one item (sequence, head) per round; NT==1 -> two items per round
  const int num_rounds = NT == 2 ? num_items : (num_items + 1) / 2;

  auto q_tile = [&](int buf, int slot) -> uint8_t* {
    return smem + buf * FWD_BUF_BYTES + (NT == 2 ? slot * TILE_BYTES : slot * 3 * TILE_BYTES);
  };
  auto k_tile = [&](int buf, int slot) -> uint8_t* {
    return smem + buf * FWD_BUF_BYTES + (NT == 2 ? 2 * TILE_BYTES : slot * 3 * TILE_BYTES + TILE_BYTES);
  };
  auto v_tile = [&](int buf, int slot) -> uint8_t* {
    return smem + buf * FWD_BUF_BYTES + (NT == 2 ? 4 * TILE_BYTES : slot * 3 * TILE_BYTES + 2 * TILE_BYTES);
  };
  // item handled by (round, slot); -1 when the slot is empty (odd tail of NT == 1)
  auto slot_item = [&](int round, int slot) -> int {
    if (NT == 2) return round;
    const int it = 2 * round + slot;
    return it < num_items ? it : -1;
  };

  if (warp == 17) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV);
      tma_prefetch_desc(&tmO);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&load_full[i], 1);
        mbar_init(&buf_free[i], 2);
        mbar_init(&bar_s[i], 1);
        mbar_init(&bar_p[i], 8);       // one arrival per softmax warp of the slot (lane 0, after __syncwarp)
        mbar_init(&bar_o[i], 1);
        mbar_init(&tmem_free[i], 8);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  if (NT == 1) {
    // V rows the TMA boxes never write must not feed NaN bit patterns into P(=0) x V
    const int first = args.rows * 128, last = args.keys_n * 128;
    for (int buf = 0; buf < 2; ++buf)
      for (int slot = 0; slot < 2; ++slot)
        for (int i = first + threadIdx.x * 16; i < last; i += FWD_THREADS * 16)
          *reinterpret_cast<uint4*>(v_tile(buf, slot) + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform for the compiler (MMA issuer)
  pdl_launch_dependents();  // set-up done: the next kernel may begin its own; then wait for the QKV producer
  pdl_wait();

  if (warp == 16) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one_sync()) {
      int k = 0;
      for (int round = blockIdx.x; round < num_rounds; round += gridDim.x, ++k) {
        const int buf = k & 1;
        mbar_wait(&buf_free[buf], ((k >> 1) & 1) ^ 1);
        if (NT == 2) {
          const int head = round % args.H, b0 = round / args.H;
          const int cq = head * 64, ck = (args.H + head) * 64, cv = (2 * args.H + head) * 64;
          mbar_expect_tx(&load_full[buf], 6 * TILE_BYTES);
          for (int t = 0; t < 2; ++t) {
            tma_load_3d(q_tile(buf, t), &tmQKV, &load_full[buf], cq, t * 128, b0);
            tma_load_3d(k_tile(buf, 0) + t * TILE_BYTES, &tmKV, &load_full[buf], ck, t * 128, b0);
            tma_load_3d(v_tile(buf, 0) + t * TILE_BYTES, &tmKV, &load_full[buf], cv, t * 128, b0);
          }
        } else {
          const int n_valid = slot_item(round, 1) >= 0 ? 2 : 1;
          mbar_expect_tx(&load_full[buf], n_valid * 3 * args.rows * 128);
          for (int slot = 0; slot < n_valid; ++slot) {
            const int item = slot_item(round, slot);
            const int head = item % args.H, b0 = (item / args.H) * args.G;
            tma_load_3d(q_tile(buf, slot), &tmQKV, &load_full[buf], head * 64, 0, b0);
            tma_load_3d(k_tile(buf, slot), &tmQKV, &load_full[buf], (args.H + head) * 64, 0, b0);
            tma_load_3d(v_tile(buf, slot), &tmQKV, &load_full[buf], (2 * args.H + head) * 64, 0, b0);
          }
        }
      }
    }
  } else if (warp == 17) {
    // ------------------------------------------------------------------ MMA issuer
    // One thread, chosen by elect.sync (the compiler then emits UTCHMMA straight, without a per-active-lane loop), and
    // descriptors that are one add behind a base computed once (sw128_desc_at): the P.V MMAs are 32 cycles of tensor
    // work each, the issue code around them used to take ~100.
    if (elect_one_sync()) {
      const uint32_t idesc_s = make_idesc_bf16(128, args.keys_n, false, false);
      const uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);
      const int ksteps = args.keys_n / 16;
      const uint32_t lo0 = (smem_u32(smem) & 0x3FFFFu) >> 4;
      auto tile_lo = [&](uint8_t* tile) -> uint32_t { return lo0 + (static_cast<uint32_t>(tile - smem) >> 4); };
      auto issue_s = [&](int buf, int slot) {
        const uint32_t a0 = tile_lo(q_tile(buf, slot)), bk = tile_lo(k_tile(buf, slot));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16_ss(tmem_base + slot * 256, sw128_desc_at(a0, kk * 32, 16, 1024), sw128_desc_at(bk, kk * 32, 16, 1024),
                       idesc_s, kk > 0);
        umma_commit(&bar_s[slot]);
      };
      auto issue_pv = [&](int buf, int slot) {
        const uint32_t v0 = tile_lo(v_tile(buf, slot));
        const int ch0 = (ksteps + 1) / 2;  // P column map of the softmax halves (see below)
        const uint32_t t_o = tmem_base + slot * 256 + 192, t_p0 = tmem_base + slot * 256, t_p1 = t_p0 + 8 * ch0;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < ksteps)
            umma_bf16_ts(t_o, (j < ch0 ? t_p0 : t_p1) + 8 * j, sw128_desc_at(v0, j * 2048, 8192, 1024), idesc_o, j > 0);
        umma_commit(&bar_o[slot]);
      };
      if (args.dephase) {
        // Event-driven issue: each slot walks S(k) -> P V(k) -> S(k+1) ... on its OWN barriers, whichever slot is ready
        // goes next, and slot 1's first S is held back until slot 0 has finished its first exp2 pass. The two softmax
        // groups then run half a period apart: one is in its (MUFU-bound) exp2 pass while the other waits for O, stores
        // its output and takes the row maxima of its next tile -- instead of both queueing for the MUFU pipe together
        // and both leaving it idle together (round-locked issue: exp2 pass 3,358 of 7,955 cycles per tile).
        const int n_my = static_cast<int>(blockIdx.x) < num_rounds
                             ? (num_rounds - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                             : 0;
        int ks[2] = {0, 0}, kp[2] = {0, 0};   // next round (CTA-local index) whose S / whose P V is to be issued, per slot
        long long spins = 0;
        while (kp[0] < n_my || kp[1] < n_my) {
          bool progress = false;
#pragma unroll
          for (int slot = 0; slot < 2; ++slot) {
            if (kp[slot] >= n_my) continue;
            if (ks[slot] == kp[slot]) {
              const int k = ks[slot];
              if (slot_item(static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x), slot) < 0) {  // empty tail slot
                ++ks[slot]; ++kp[slot]; progress = true;
                continue;
              }
              if (slot == 1 && k == 0 && kp[0] == 0) continue;   // the half-period offset
              if (!mbar_test_wait(&load_full[k & 1], (k >> 1) & 1)) continue;
              if (!mbar_test_wait(&tmem_free[slot], (k & 1) ^ 1)) continue;
              tcgen05_fence_after();
              issue_s(k & 1, slot);
              ++ks[slot];
              progress = true;
            } else {
              const int k = kp[slot];
              if (!mbar_test_wait(&bar_p[slot], k & 1)) continue;
              tcgen05_fence_after();
              issue_pv(k & 1, slot);
              ++kp[slot];
              progress = true;
            }
          }
          if (progress) {
            spins = 0;
          } else {
            __nanosleep(args.dephase > 1 ? args.dephase : 0);   // keep the probe loop off the issue port of this warp's scheduler
            if (++spins > (1ll << 26)) __trap();   // a dead pipeline must fail the launch, not hang the GPU
          }
        }
      } else {
        int k = 0;
        for (int round = blockIdx.x; round < num_rounds; round += gridDim.x, ++k) {
          const int buf = k & 1;
          const uint32_t par = k & 1;
          mbar_wait(&load_full[buf], (k >> 1) & 1);
          tcgen05_fence_after();
          for (int slot = 0; slot < 2; ++slot) {
            if (slot_item(round, slot) < 0) continue;
            mbar_wait(&tmem_free[slot], par ^ 1);
            tcgen05_fence_after();
            issue_s(buf, slot);
          }
          for (int slot = 0; slot < 2; ++slot) {
            if (slot_item(round, slot) < 0) continue;
            mbar_wait(&bar_p[slot], par);
            tcgen05_fence_after();
            issue_pv(buf, slot);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue groups
    const int slot = warp >> 3;          // group 0: warps 0-7, group 1: warps 8-15
    const int q = warp & 3;              // TMEM lane quarter
    const int hf = (warp >> 2) & 1;      // which half of the key columns this thread owns
    const int r = q * 32 + lane;         // row within the tile
    const int gtid = threadIdx.x - slot * 256;
    float* xchg = xchg_all + slot * 512;
    const int nchunks = args.keys_n / 16;
    const int ch0 = (nchunks + 1) / 2;  // half 0 owns key chunks [0, ch0), half 1 owns [ch0, nchunks)
    const int c_begin = hf == 0 ? 0 : ch0;
    const int c_end = hf == 0 ? ch0 : nchunks;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + slot * 256;
    // P (bf16 pairs, 8 columns per 16-key chunk) is written into the thread's OWN S columns, behind its
    // read pointer, so no cross-thread hazard exists and nothing has to be buffered in registers:
    // half 0 -> columns [0, 8 ch0), half 1 -> columns [16 ch0, ...). The MMA issuer uses the same map.
    const uint32_t p_base = trow + (hf == 0 ? 0 : 16 * ch0) - 8 * c_begin;
    constexpr int O_COL = 192;

    int k = 0;
    int store_buf = -1;   // gtid 0: the smem buffer whose output store is committed but not yet known to have been read
    // The group's store thread hands a buffer back one step late: it commits the TMA store of round k, goes on, and only
    // after the NEXT round's max pass does it wait for the store to have read the staging tile (long done by then) and
    // release the buffer -- the read-completion latency (~500 cycles) is off the group's per-tile chain, and the producer
    // still has most of a round to re-fill the buffer for round k + 2.
    auto release_store_buf = [&]() {
      if (gtid == 0 && store_buf >= 0) {
        tma_store_wait_read<0>();
        mbar_arrive(&buf_free[store_buf]);
        store_buf = -1;
      }
    };
    for (int round = blockIdx.x; round < num_rounds; round += gridDim.x, ++k) {
      const int item = slot_item(round, slot);
      if (item < 0) {  // empty tail slot: still hand the smem buffer back
        release_store_buf();
        if (gtid == 0) mbar_arrive(&buf_free[k & 1]);
        continue;
      }
      const int buf = k & 1;
      const uint32_t par = k & 1;
      const int head = item % args.H;
      const int b0 = (item / args.H) * args.G;
      const int t = NT == 2 ? slot : 0;  // query tile within the sequence
      int lo, hi;
      bool row_valid;
      key_range(args, NT, NT == 1 ? r : t * 128 + r, lo, hi, row_valid);
      if (!row_valid) hi = lo;  // no valid keys: the row is all padding
      // A warp whose 32 rows ALL lie past the last query row skips both passes (N = 197: rows 224..255 of the second
      // tile, two of the sixteen softmax warps -- an eighth of the exp2 work of a MUFU-bound pass). Its P columns keep
      // whatever bits S left there: the P.V MMA is row-wise, and output rows past the sequence are clipped by the store.
      const bool warp_live = (NT == 1 ? 0 : t * 128) + q * 32 < (NT == 1 ? args.rows : args.N);
      const int c_end_w = warp_live ? c_end : c_begin;
      // NT == 2: no masking. Key rows >= N are zero-filled by TMA, so their scores are exactly 0 (harmless
      // in the max), their V rows are 0 (no contribution to O) and their exp2(-m) terms are subtracted
      // from the row sum below. NT == 1 (packed sequences): whole 16-key chunks are classified as
      // outside / inside / straddling the row's own sequence [lo, hi).
      auto chunk_class = [&](int c) -> int {
        if (NT == 2) return 1;
        const int c0 = c * 16;
        if (c0 + 16 <= lo || c0 >= hi) return 0;
        return (c0 >= lo && c0 + 16 <= hi) ? 1 : 2;
      };

      const bool prof_on = args.prof != nullptr && gtid == 0;
      long long tp0 = prof_on ? clock64() : 0;
      auto lap = [&](int idx) {
        if (prof_on) {
          const long long now = clock64();
          atomicAdd(args.prof + slot * 8 + idx, static_cast<unsigned long long>(now - tp0));
          tp0 = now;
        }
      };
      mbar_wait(&bar_s[slot], par);
      tcgen05_fence_after();
      lap(0);
      // ---- pass 1: row max of the raw scores over this thread's column range
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = c_begin; c < c_end_w; ++c) {
        // tcgen05.ld/st are warp-collective (.sync.aligned): every lane issues them for every chunk; only the
        // arithmetic in between may diverge on the per-row chunk class
        const int cls = chunk_class(c);
        uint32_t v[16];
        tmem_ld_32x32b_x16(trow + c * 16, v);
        tmem_ld_wait();
        if (cls == 1) {
          float m0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
          float m1 = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
#pragma unroll
          for (int j = 4; j < 16; j += 4) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
          }
          mx = fmaxf(mx, fmaxf(m0, m1));
        } else if (cls == 2) {
          for (int j = 0; j < 16; ++j)
            if (c * 16 + j >= lo && c * 16 + j < hi) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
      }
      xchg[hf * 128 + r] = mx;
      release_store_buf();   // the previous round's output store has had the whole max pass to read its staging tile
      lap(1);
      named_bar_sync(1 + slot, 256);  // also: every thread of the group is done with pass 1
      lap(2);
      mx = fmaxf(mx, xchg[(hf ^ 1) * 128 + r]);
      const float m2 = mx == -INFINITY ? 0.f : mx * args.scale_log2;
      // ---- pass 2: p = 2^(s*c - m) -> bf16 pairs -> written over the already-consumed S columns
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll 1
      for (int c = c_begin; c < c_end_w; ++c) {
        const int cls = chunk_class(c);
        uint32_t pk[8];
        uint32_t v[16];
        tmem_ld_32x32b_x16(trow + c * 16, v);
        tmem_ld_wait();
        if (cls == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = 0u;
        } else {
          if (cls == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), args.scale_log2, -m2));
              const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), args.scale_log2, -m2));
              sum0 += p0;
              sum1 += p1;
              pk[j] = pack_bf16x2(p0, p1);
            }
          } else {
            for (int j = 0; j < 8; ++j) {
              float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), args.scale_log2, -m2));
              float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), args.scale_log2, -m2));
              if (!(c * 16 + 2 * j >= lo && c * 16 + 2 * j < hi)) p0 = 0.f;
              if (!(c * 16 + 2 * j + 1 >= lo && c * 16 + 2 * j + 1 < hi)) p1 = 0.f;
              sum0 += p0;
              sum1 += p1;
              pk[j] = pack_bf16x2(p0, p1);
            }
          }
        }
        tmem_st_32x32b_x8(p_base + c * 8, pk);
      }
      float sum = sum0 + sum1;
      xchg[256 + hf * 128 + r] = sum;
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_p[slot]);   // one arrival per warp (256 per-thread arrivals measured the same)
      lap(3);
      named_bar_sync(1 + slot, 256);
      sum += xchg[256 + (hf ^ 1) * 128 + r];
      if (NT == 2) sum -= static_cast<float>(args.keys_n - args.Nk) * ex2_approx(-m2);  // zero-padded keys
      const float inv = sum > 0.f ? 1.f / sum : 0.f;
      if (hf == 0 && row_valid) {
        const int rr = NT == 1 ? r : t * 128 + r;
        const int b = b0 + (NT == 1 ? rr / args.N : 0);
        const int n = NT == 1 ? rr % args.N : rr;
        if (b < args.B) args.lse2[(static_cast<long long>(b) * args.H + head) * args.Ns + n] = m2 + log2f(sum);
      }
      // ---- epilogue: O / rowsum -> bf16 -> swizzled staging (the dead Q tile) -> TMA store
      lap(4);
      mbar_wait(&bar_o[slot], par);
      tcgen05_fence_after();
      lap(5);
      uint32_t v[32];
      tmem_ld_32x32b_x32(trow + O_COL + hf * 32, v);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_free[slot]);  // S of the next round may overwrite this slot's TMEM
      uint8_t* stg = q_tile(buf, slot);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]) * inv, __uint_as_float(v[8 * j + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv);
        *reinterpret_cast<uint4*>(stg + sw128_offset(r, hf * 4 + j)) = o;
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + slot, 256);
      if (gtid == 0) {
        tma_store_3d(&tmO, stg, head * 64, NT == 1 ? 0 : t * 128, b0);
        tma_store_commit();
        store_buf = buf;               // released in release_store_buf(): staging (the dead Q tile) is part of the buffer
      }
      lap(6);
      if (prof_on) atomicAdd(args.prof + slot * 8 + 7, 1ull);
    }
    release_store_buf();
    if (gtid == 0) tma_store_wait_all<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// backward: persistent CTAs (one per SM) looping over (sequence | packed group, head) items
// ------------------------------------------------------------------------------------------------
// smem: operand sets [Q (NT) | dO (NT) | K (NT) | V (NT)] x SETS | P (2 chunks) | dS (2 chunks) | barriers |
//       row constants (2 buffers). NT == 1 (N <= 128, packed local crops) has two operand sets, so the loads of
//       item i+1 fly during the whole of item i; NT == 2 re-loads its single set as soon as the last MMA of the
//       item has retired.
// TMEM: S [0,128) | dP [128,256) | dQ_t [256+64t) | dK [384,448) | dV [448,512), allocated once per CTA.
// warps: 0      = control: TMA loads + every tcgen05.mma. S/dP of the NEXT pair is issued as soon as the math
//                 warps hold the current S/dP in registers, i.e. before the dQ/dK/dV MMAs of the current pair.
//        1..3   = L2 prefetch (`prefetch.global.L2`, one 128-byte line per row and tensor) of the Q / K / V / dO / O rows of
//                 the item two ahead of the one in work: the TMA loads can only be issued when the previous item's tiles
//                 are dead, and the row-constant loads are latency-bound -- both should find their data in L2
//        20..23 = auxiliary group (one warp per TMEM lane quarter): the finished dK / dV / dQ accumulators go TMEM -> bf16
//                 -> global memory straight from registers (each thread owns one 128-byte output row; 256-bit stores,
//                 vector red.add in the accumulating long-sequence mode), then the row constants delta = sum_d dO*O and
//                 lse2 of item i+2. The math warps never wait for the dK/dV MMAs and never store: they go from one pair's
//                 P/dS straight to the next pair's scores. (Measured alternative: row constants on warps 1..3, two items
//                 ahead and off this group's chain -- slower, 134 -> 257 us on the packed crops: inside warpgroup 0's
//                 64-register budget the producer spills, and every spill re-load is an L2 round trip.)
//        4..19  = math, four threads per query row (32 key columns each).
// Registers: 24 warps cap the kernel at 80 registers per thread, and at 80 the math warps spill a handful of loop
// variables -- with 227 KB of shared memory carved out of L1 every spill re-load is an L2 round trip on the critical path
// (ncu: a quarter of all stall samples). So warpgroup 0 (control + idle warps) and the auxiliary warpgroup shrink to 64
// registers with setmaxnreg and the four math warpgroups grow to 88: 2 x 128 x 16 given back = 4 x 128 x 8 taken (only what
// the CTA was launched with can be re-distributed; asking for more blocks forever). No spills anywhere then.
constexpr int BWD_MATH_THREADS = 512;
constexpr int BWD_AUX_THREADS = 128;
constexpr int BWD_PF_THREADS = 96;
constexpr int BWD_THREADS = 128 + BWD_MATH_THREADS + BWD_AUX_THREADS;

template <int NT>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmKV,
                     const __grid_constant__ CUtensorMap tmDO, const AttnArgs args, const int num_items) {
  constexpr int SETS = NT == 1 ? 2 : 1;
  constexpr int PPI = NT * NT;                    // (key tile, query tile) pairs per item
  constexpr int SET_BYTES = 4 * NT * TILE_BYTES;  // Q | dO | K | V
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sP = smem + SETS * SET_BYTES;
  uint8_t* sdS = sP + 2 * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 2 * TILE_BYTES);
  uint64_t* bar_load = bars;          // [2] per operand set
  uint64_t* bar_sdp = bars + 2;       // S and dP ready in TMEM
  uint64_t* bar_sdp_free = bars + 3;  // math threads hold S/dP in registers
  uint64_t* bar_pds = bars + 4;       // P and dS written to smem
  uint64_t* bar_mma = bars + 5;       // dQ/dK/dV MMAs of the pair finished
  uint64_t* bar_rowc_full = bars + 6; // [2] row constants of an item written
  uint64_t* bar_rowc_free = bars + 8; // [2] ... and copied to registers by every math thread
  uint64_t* bar_dkv_ready = bars + 10; // every MMA up to the last pair of a key tile has retired: dK / dV complete
  uint64_t* bar_dkv_free = bars + 11;  // the auxiliary group has read dK / dV out of TMEM
  uint64_t* bar_dq_free = bars + 12;   // ... and the dQ tiles of the item
  uint64_t* bar_grp = bars + 13;       // [4] NT == 2: operand groups {K0,V0} {Q0,dO0} {Q1,dO1} {K1,V1} of the single set
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  volatile int* item_flag = reinterpret_cast<volatile int*>(bars + 18);  // the item (CTA-local index) the math warps work on
  float* rowc = reinterpret_cast<float*>(bars + 24);  // [2][NT*128][2]: delta, lse2
  // NT == 1: tile row -> (sequence within the packed group) << 16 | token, -1 for rows past the group; filled once, so the
  // per-item loops of the helper warps never divide (a division by a run-time N is ~40 dependent instructions, and a lone
  // warp runs them at ~4 cycles each: index arithmetic, not memory, was what bounded the row constants)
  int* rowtab = reinterpret_cast<int*>(rowc + 2 * NT * 128 * 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = (num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  auto item_of = [&](int k) { return static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x); };
  auto set_base = [&](int k) -> uint8_t* { return smem + (SETS == 2 ? (k & 1) : 0) * SET_BYTES; };
  // what an item is: (packed group | sequence, head) -- or, in the paired long-sequence mode, additionally a (query block,
  // key block) pair of up to 256 x 256 tokens of that sequence, addressed by row offsets into whole-sequence tensor maps
  struct Item { int head, b0, q0, k0, nq, keys_n; };
  auto decode = [&](int item) {
    Item d;
    if (NT == 2 && args.nblk > 1) {
      const int npairs = args.nblk * args.nblk;
      const int pair = item % npairs, bh = item / npairs;
      d.head = bh % args.H;
      d.b0 = bh / args.H;
      d.q0 = (pair / args.nblk) * 256;
      d.k0 = (pair % args.nblk) * 256;
      d.nq = min(256, args.Ns - d.q0);
      d.keys_n = (min(256, args.Ns - d.k0) + 15) & ~15;
    } else {
      d.head = item % args.H;
      d.b0 = (item / args.H) * args.G;
      d.q0 = d.k0 = 0;
      d.nq = NT == 1 ? args.rows : args.N;
      d.keys_n = args.keys_n;
    }
    return d;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(&bar_load[0], 1);
    mbar_init(&bar_load[1], 1);
    mbar_init(bar_sdp, 1);
    // arrivals are per WARP (lane 0 after __syncwarp); 512 per-thread arrivals measured the same (23,958 vs 24,089 clk / item)
    mbar_init(bar_sdp_free, BWD_MATH_THREADS / 32);
    mbar_init(bar_pds, BWD_MATH_THREADS / 32);
    mbar_init(bar_mma, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_rowc_full[i], BWD_AUX_THREADS / 32);
      mbar_init(&bar_rowc_free[i], BWD_MATH_THREADS / 32);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&bar_grp[i], 1);
    *item_flag = -1;
    mbar_init(bar_dkv_ready, 1);
    mbar_init(bar_dkv_free, BWD_AUX_THREADS / 32);
    mbar_init(bar_dq_free, BWD_AUX_THREADS / 32);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  if (NT == 1 && threadIdx.x < 128) {
    const int idx = threadIdx.x;
    rowtab[idx] = idx < args.rows ? ((idx / args.N) << 16) | (idx % args.N) : -1;
  }
  if (NT == 1) {
    // rows never written by the TMA boxes: zero them once so 0 x garbage cannot become NaN
    const int first = args.rows * 128, last = 128 * 128;
    for (int t = 0; t < SETS * 4; ++t)
      for (int i = first + threadIdx.x * 16; i < last; i += BWD_THREADS * 16)
        *reinterpret_cast<uint4*>(smem + t * TILE_BYTES + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform for the compiler: the MMA issuer
                                                                       // keeps it in a uniform register
  const uint32_t T_S = tmem_base, T_DP = tmem_base + 128, T_DQ = tmem_base + 256, T_DK = tmem_base + 384,
                 T_DV = tmem_base + 448;
  pdl_launch_dependents();
  pdl_wait();

  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;\n");
  else if (warp < 20) asm volatile("setmaxnreg.inc.sync.aligned.u32 88;\n");
  if (warp == 0) {
    // ------------------------------------------------------------------ control: TMA + MMA issue
    // (elect.sync, not lane == 0: the compiler then knows a single thread runs this and issues the uniform-datapath
    //  instructions -- UTCHMMA, UTMALDG -- straight, without a per-active-lane loop around each of them)
    if (n_my > 0 && elect_one_sync()) {
      // NT == 1: a whole operand set (Q, K, V, dO of one packed group) per barrier, two sets in flight
      auto issue_load = [&](int k) {
        const Item it = decode(item_of(k));
        const int head = it.head, b0 = it.b0;
        uint8_t* sQ = set_base(k);
        uint8_t* sdO = sQ + NT * TILE_BYTES;
        uint8_t* sK = sdO + NT * TILE_BYTES;
        uint8_t* sV = sK + NT * TILE_BYTES;
        uint64_t* bar = &bar_load[k & 1];
        const int cq = head * 64, ck = (args.H + head) * 64, cv = (2 * args.H + head) * 64;
        mbar_expect_tx(bar, 4 * args.rows * 128);
        tma_load_3d(sQ, &tmQKV, bar, cq, 0, b0);
        tma_load_3d(sK, &tmQKV, bar, ck, 0, b0);
        tma_load_3d(sV, &tmQKV, bar, cv, 0, b0);
        tma_load_3d(sdO, &tmDO, bar, cq, 0, b0);
      };
      auto wait_load = [&](int k) {
        mbar_wait(&bar_load[k & 1], (k >> 1) & 1);  // k >> 1 completed phases of this set's barrier before item k
        tcgen05_fence_after();
      };
      const uint32_t idesc_q = make_idesc_bf16(128, 64, false, true);  // dQ: A K-major, B MN-major
      const uint32_t idesc_kv = make_idesc_bf16(128, 64, true, true);  // dK/dV: both MN-major
      // descriptors: every operand lies a compile-time number of bytes behind the 1024-aligned smem base (plus the set
      // offset when there are two sets), so each one is a single add to `lo0` / `los` (common.cuh: sw128_desc_at)
      const uint32_t lo0 = (smem_u32(smem) & 0x3FFFFu) >> 4;
      constexpr uint32_t P_OFF = SETS * SET_BYTES, DS_OFF = P_OFF + 2 * TILE_BYTES;
      auto set_lo = [&](int k) -> uint32_t { return lo0 + (SETS == 2 ? (k & 1) : 0) * (SET_BYTES >> 4); };
      auto issue_sdp = [&](uint32_t los, int p, int keys_n) {
        const int u = p / NT, t = p % NT;
        const int ku = max(16, min(128, keys_n - u * 128));  // keys in this key tile (multiple of 16; an empty tile of a
                                                             // short key block still runs on zeros)
        const uint32_t idesc_s = make_idesc_bf16(128, ku, false, false);
        const uint32_t q_t = t * TILE_BYTES, do_t = (NT + t) * TILE_BYTES;
        const uint32_t k_u = (2 * NT + u) * TILE_BYTES, v_u = (3 * NT + u) * TILE_BYTES;
        // S and dP are independent accumulation chains: issued alternately, so an MMA never queues right behind the one
        // it accumulates onto (back-to-back MMAs into ONE accumulator ran at ~100 cycles each whatever their N)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          umma_bf16_ss(T_S, sw128_desc_at(los, q_t + kk * 32, 16, 1024), sw128_desc_at(los, k_u + kk * 32, 16, 1024), idesc_s,
                       kk > 0);
          umma_bf16_ss(T_DP, sw128_desc_at(los, do_t + kk * 32, 16, 1024), sw128_desc_at(los, v_u + kk * 32, 16, 1024), idesc_s,
                       kk > 0);
        }
        umma_commit(bar_sdp);
      };
      auto issue_dqkv = [&](uint32_t los, int p, int keys_n, int nq) {
        const int u = p / NT, t = p % NT;
        const int nk = max(16, min(128, keys_n - u * 128)) >> 4;
        const uint32_t q_t = t * TILE_BYTES, do_t = (NT + t) * TILE_BYTES, k_u = (2 * NT + u) * TILE_BYTES;
        // dQ_t (+)= dS[128 q, ku keys] . K_u[ku keys, 64] ; dV_u (+)= P^T[128 keys, 128 q] . dO_t[128 q, 64] ;
        // dK_u (+)= dS^T . Q_t. The dK / dV reductions run over the query rows of tile t: rows past the sequence end hold
        // P = dS = 0 (lse2 = +inf), so only the live 16-row steps are issued. Three independent accumulation chains,
        // issued round-robin.
        const int qsteps = (max(0, min(128, nq - t * 128)) + 15) >> 4;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < nk)
            umma_bf16_ss(T_DQ + t * 64, sw128_desc_at(lo0, DS_OFF + (j >> 2) * TILE_BYTES + (j & 3) * 32, 16, 1024),
                         sw128_desc_at(los, k_u + j * 2048, 8192, 1024), idesc_q, (u > 0 || j > 0));
          if (j < qsteps) {
            umma_bf16_ss(T_DV, sw128_desc_at(lo0, P_OFF + j * 2048, TILE_BYTES, 1024),
                         sw128_desc_at(los, do_t + j * 2048, 8192, 1024), idesc_kv, (t > 0 || j > 0));
            umma_bf16_ss(T_DK, sw128_desc_at(lo0, DS_OFF + j * 2048, TILE_BYTES, 1024),
                         sw128_desc_at(los, q_t + j * 2048, 8192, 1024), idesc_kv, (t > 0 || j > 0));
          }
        }
        umma_commit(bar_mma);
      };

      // NT == 2 has ONE operand set (128 KB) and re-fills it group by group as the pairs of the item retire: pair order is
      // (u,t) = (0,0) (0,1) (1,0) (1,1), so {K0,V0} is dead after pair 1 and {Q0,dO0} after pair 2 -- the next item's first
      // scores are issued right behind the last dQ/dK/dV MMAs of this one, with no load latency in between.
      auto issue_group = [&](int k, int g) {
        const Item it = decode(item_of(k));
        const int head = it.head, b0 = it.b0;
        uint8_t* sQ = set_base(k);
        uint8_t* sdO = sQ + NT * TILE_BYTES;
        uint8_t* sK = sdO + NT * TILE_BYTES;
        uint8_t* sV = sK + NT * TILE_BYTES;
        const int cq = head * 64, ck = (args.H + head) * 64, cv = (2 * args.H + head) * 64;
        uint64_t* bar = &bar_grp[g];
        mbar_expect_tx(bar, 2 * TILE_BYTES);
        const int t = (g == 0 || g == 1) ? 0 : 1;
        if (g == 0 || g == 3) {
          tma_load_3d(sK + t * TILE_BYTES, &tmKV, bar, ck, it.k0 + t * 128, b0);
          tma_load_3d(sV + t * TILE_BYTES, &tmKV, bar, cv, it.k0 + t * 128, b0);
        } else {
          tma_load_3d(sQ + t * TILE_BYTES, &tmQKV, bar, cq, it.q0 + t * 128, b0);
          tma_load_3d(sdO + t * TILE_BYTES, &tmDO, bar, cq, it.q0 + t * 128, b0);
        }
      };
      auto wait_group = [&](int k, int g) { mbar_wait(&bar_grp[g], k & 1); };

      if (NT == 2) {
        for (int g = 0; g < 4; ++g) issue_group(0, g);
        wait_group(0, 0);
        wait_group(0, 1);
      } else {
        issue_load(0);
        if (n_my > 1) issue_load(1);
        wait_load(0);
      }
      Item cur = decode(item_of(0));
      issue_sdp(set_lo(0), 0, cur.keys_n);
      // developer instrumentation of the issuing thread (counters 9..15): cycles waiting for [9] the math warps to take
      // S/dP, [10] operand loads, [11] P/dS in smem, [12] the accumulator drain, [14] the item's last MMAs; [13] / [15] =
      // cycles inside the S/dP / dQ,dK,dV issue code (back-pressure of the MMA queue shows up here)
      const bool cprof = args.prof != nullptr;
      long long ct0 = cprof ? clock64() : 0;
      auto clap = [&](int idx) {
        if (cprof) {
          const long long now = clock64();
          atomicAdd(args.prof + idx, static_cast<unsigned long long>(now - ct0));
          ct0 = now;
        }
      };
      int gp = 0;
      for (int k = 0; k < n_my; ++k) {
        const Item nxt = k + 1 < n_my ? decode(item_of(k + 1)) : cur;
        const uint32_t los = set_lo(k), los_n = set_lo(k + 1);
#pragma unroll 1   // (unrolled, the compiler hoists every descriptor of every pair into registers and the issuer spills)
        for (int p = 0; p < PPI; ++p, ++gp) {
          mbar_wait(bar_sdp_free, gp & 1);  // the math warps hold S/dP(gp) in registers
          tcgen05_fence_after();
          clap(9);
          const bool last = p + 1 == PPI;
          if (!last) {
            if (NT == 2 && p == 0) wait_group(k, 2);  // pair 1 = (K0,V0,Q1,dO1)
            if (NT == 2 && p == 1) wait_group(k, 3);  // pair 2 = (K1,V1,Q0,dO0)
            clap(10);
            issue_sdp(los, p + 1, cur.keys_n);
            clap(13);
          } else if (NT == 1 && k + 1 < n_my) {
            wait_load(k + 1);
            clap(10);
            issue_sdp(los_n, 0, nxt.keys_n);
            clap(13);
          }
          mbar_wait(bar_pds, gp & 1);
          clap(11);
          if (p % NT == 0) {
            // first pair of a key tile: its MMAs overwrite dK / dV (at p == 0 dQ as well) -- the auxiliary group must
            // have read the previous contents out of TMEM
            const int gk = k * NT + p / NT;
            if (gk > 0) mbar_wait(bar_dkv_free, (gk - 1) & 1);
            if (p == 0 && k > 0) mbar_wait(bar_dq_free, (k - 1) & 1);
          }
          tcgen05_fence_after();
          clap(12);
          issue_dqkv(los, p, cur.keys_n, cur.nq);
          if (p % NT == NT - 1) umma_commit(bar_dkv_ready);
          clap(15);
          if (NT == 2 && k + 1 < n_my) {
            // bar_pds(gp) implies bar_mma(gp - 1): every math warp waited for it before writing P / dS of this pair
            if (p == 2) issue_group(k + 1, 0);
            if (p == 3) {
              issue_group(k + 1, 1);
              wait_group(k + 1, 0);
              wait_group(k + 1, 1);
              clap(10);
              issue_sdp(los_n, 0, nxt.keys_n);  // the math warps hold S/dP of pair 3 (bar_sdp_free above): T_S / T_DP are free
              clap(13);
            }
          }
          if (last) {
            mbar_wait(bar_mma, gp & 1);  // every MMA of the item has retired: its operand set is free again
            clap(14);
            if (NT == 1) {
              if (k + 2 < n_my) issue_load(k + 2);
            } else if (k + 1 < n_my) {
              issue_group(k + 1, 2);
              issue_group(k + 1, 3);
            }
          }
        }
        cur = nxt;
#ifndef B200SSL_ATTN_BWD_PROF
        if (cprof) atomicAdd(args.prof + 8, 1ull);   // items (the math leader counts them when its own counters are compiled in)
#endif
      }
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------------ L2 prefetch, PF_AHEAD items ahead of the math warps
    const int ptid = threadIdx.x - 32;  // 0..95
    // how far ahead: far enough for an HBM round trip before the TMA loads are issued, near enough that the lines are still
    // in L2 when they are -- the whole GPU streams ~3 MB per microsecond through a 126 MB L2. Two-tile items last ~10 us
    // and load their operands during the previous item: one item ahead; packed items last ~3.5 us and load two items
    // ahead: two.
    constexpr int PF_AHEAD = NT == 2 ? 1 : 2;
    const long long row_qkv = 3LL * args.H * 64, row_o = static_cast<long long>(args.H) * 64;
    for (int k = PF_AHEAD; k < n_my; ++k) {
      while (*item_flag < k - PF_AHEAD) __nanosleep(256);
      const Item it = decode(item_of(k));
      for (int rr = ptid; rr < NT * 128; rr += BWD_PF_THREADS) {   // one row per thread and turn: five lines
        int b = it.b0, nq_ = it.q0 + rr, nk_ = it.k0 + rr;
        bool okq = nq_ < args.dq_rows, okk = nk_ < args.dkv_rows;
        if (NT == 1) {
          const int e = rowtab[rr];
          b += e >> 16;
          nq_ = nk_ = e & 0xffff;
          okq = okk = e >= 0 && b < args.B;
        }
        if (okq) {
          const long long bn = static_cast<long long>(b) * args.Ns + nq_;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(args.qkv_q + bn * row_qkv + it.head * 64));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(args.dout + bn * row_o + it.head * 64));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(args.out + bn * row_o + it.head * 64));
        }
        if (okk) {
          const long long bn = static_cast<long long>(b) * args.Ns + nk_;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(args.qkv_k + bn * row_qkv + (args.H + it.head) * 64));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(args.qkv_k + bn * row_qkv + (2 * args.H + it.head) * 64));
        }
      }
    }
  } else if (warp >= 20) {
    // ------------------------------------------------------------------ auxiliary group: row constants + accumulator drain
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;\n");
    const int aq = warp & 3;          // TMEM lane quarter
    const int ar = aq * 32 + lane;    // the tile row this thread drains; also its index in the group
    const uint32_t lane_off = static_cast<uint32_t>(aq * 32) << 16;
    // 128x64 fp32 accumulator at TMEM column `tcol` -> bf16 -> row `row0 + ar` of `base` (a [B, Ns, 3, H, 64] tensor seen
    // from its first row of this launch), 64 columns from `gcol`. One thread = one 128-byte output row: four 256-bit
    // stores, or eight 128-bit bf16x2 vector reductions when the launch accumulates (long-sequence block pairs).
    auto drain = [&](uint32_t tcol, __nv_bfloat16* base, int rows_behind, bool accumulate, int gcol, int row0, int b0) {
      bool valid;
      int b, n;
      if (NT == 1) {
        valid = ar < args.rows;
        b = b0 + (valid ? ar / args.N : 0);
        n = valid ? ar % args.N : 0;
        valid = valid && b < args.B;
      } else {
        b = b0;
        n = row0 + ar;
        valid = n < rows_behind;
      }
      __nv_bfloat16* dst = base + (static_cast<long long>(b) * args.Ns + n) * (3LL * args.H * 64) + gcol;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tcol + lane_off + hf * 32, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              pk[e] = pack_bf16x2(__uint_as_float(v[16 * j + 2 * e]), __uint_as_float(v[16 * j + 2 * e + 1]));
            __nv_bfloat16* d16 = dst + hf * 32 + j * 16;
            if (accumulate) {
              asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(d16), "r"(pk[0]), "r"(pk[1]),
                           "r"(pk[2]), "r"(pk[3])
                           : "memory");
              asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(d16 + 8), "r"(pk[4]), "r"(pk[5]),
                           "r"(pk[6]), "r"(pk[7])
                           : "memory");
            } else {
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(d16), "r"(pk[0]), "r"(pk[1]),
                           "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                           : "memory");
            }
          }
        }
      }
    };

    // row constants of item k: one thread per row (16 + 16 sixteen-byte loads of O and dO, one of lse2)
    auto produce_rowc = [&](int k) {
      const Item it = decode(item_of(k));
      const int head = it.head, b0 = it.b0;
      float* rc = rowc + (k & 1) * (NT * 128 * 2);
      mbar_wait(&bar_rowc_free[k & 1], ((k >> 1) & 1) ^ 1);
      for (int idx = ar; idx < NT * 128; idx += BWD_AUX_THREADS) {
        bool row_valid;
        int b = b0, n;
        if (NT == 1) {
          const int e = rowtab[idx];
          row_valid = e >= 0;
          b += e >> 16;
          n = e & 0xffff;
        } else {
          row_valid = idx < it.nq;
          n = it.q0 + idx;
        }
        float delta = 0.f, l2 = INFINITY;  // invalid rows: lse2 = +inf so that P = 0
        if (row_valid && b < args.B) {
          const long long off = ((static_cast<long long>(b) * args.Ns + n) * args.H + head) * 64;
          const uint4* po = reinterpret_cast<const uint4*>(args.out + off);
          const uint4* pd = reinterpret_cast<const uint4*>(args.dout + off);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 a = __ldg(po + j), d = __ldg(pd + j);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 x = unpack_bf16x2(aw[e]), y = unpack_bf16x2(dw[e]);
              delta += x.x * y.x + x.y * y.y;
            }
          }
          l2 = args.lse2[(static_cast<long long>(b) * args.H + head) * args.Ns + n];
        }
        rc[idx * 2] = delta;
        rc[idx * 2 + 1] = l2;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_rowc_full[k & 1]);
    };
    // the drain is what the MMA issuer waits for (TMEM re-use), the row constants have a whole item of slack: item k + 2's
    // are produced after item k's accumulators are out (their buffer was released when the math warps started item k)
    if (n_my > 0) produce_rowc(0);
    if (n_my > 1) produce_rowc(1);
    int gk = 0;
    for (int k = 0; k < n_my; ++k) {
      const Item it = decode(item_of(k));
      for (int u = 0; u < NT; ++u, ++gk) {
        mbar_wait(bar_dkv_ready, gk & 1);
        tcgen05_fence_after();
        const int row0 = NT == 1 ? 0 : it.k0 + u * 128;
        drain(T_DK, args.dkv, args.dkv_rows, args.acc_dkv != 0, (args.H + it.head) * 64, row0, it.b0);
        drain(T_DV, args.dkv, args.dkv_rows, args.acc_dkv != 0, (2 * args.H + it.head) * 64, row0, it.b0);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dkv_free);
      }
      // the last key tile's barrier covers every MMA of the item: dQ is complete as well
#pragma unroll
      for (int t = 0; t < NT; ++t)
        drain(T_DQ + t * 64, args.dq, args.dq_rows, args.acc_dq != 0, it.head * 64, NT == 1 ? 0 : it.q0 + t * 128, it.b0);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_dq_free);
      if (k + 2 < n_my) produce_rowc(k + 2);
    }
  } else {
    // ------------------------------------------------------------------ math: P and dS of every pair
    const int q = warp & 3;           // TMEM lane quarter
    const int qc = (warp - 4) >> 2;   // which 32 key columns of the 128-wide key tile
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const bool leader = threadIdx.x == 128;

    int lo[NT], hi[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      bool row_valid;
      key_range(args, NT, NT == 1 ? r : t * 128 + r, lo[t], hi[t], row_valid);
    }
    int gp = 0;
    // developer instrumentation: cycles of the leader math thread per phase. Compiled in only with -DB200SSL_ATTN_BWD_PROF:
    // the clock and pointer registers it keeps alive are what tips the math warps (88 registers) into spilling their
    // loop counters, and a spill re-load is an L2 round trip here. The issuer's counters (9..15) are always available.
#ifdef B200SSL_ATTN_BWD_PROF
    const bool prof_on = args.prof != nullptr && leader;
    long long tp0 = prof_on ? clock64() : 0;
    auto lap = [&](int idx) {
      if (prof_on) {
        const long long now = clock64();
        atomicAdd(args.prof + idx, static_cast<unsigned long long>(now - tp0));
        tp0 = now;
      }
    };
#else
    constexpr bool prof_on = false;
    auto lap = [](int) {};
    (void)leader;
#endif

    for (int k = 0; k < n_my; ++k) {
      const Item it = decode(item_of(k));
      float delta[NT], lse2[NT];
      if (threadIdx.x == 128) *item_flag = k;   // paces the L2 prefetch warps
      lap(7);
      {
        const float* rc = rowc + (k & 1) * (NT * 128 * 2);
        mbar_wait(&bar_rowc_full[k & 1], (k >> 1) & 1);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          delta[t] = rc[(t * 128 + r) * 2];
          lse2[t] = rc[(t * 128 + r) * 2 + 1];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_rowc_free[k & 1]);
      }
      lap(0);

#pragma unroll
      for (int u = 0; u < NT; ++u) {
        const int ku = max(16, min(128, it.keys_n - u * 128));
#pragma unroll
        for (int t = 0; t < NT; ++t, ++gp) {
          mbar_wait(bar_sdp, gp & 1);
          tcgen05_fence_after();
          lap(1);
          const int col0 = qc * 32;             // first key column (within the tile) of this thread
          // warp-uniform: this warp's 32 key columns lie inside the key tile, and at least one of its 32 query rows exists.
          // A warp of padding rows only (N = 197: rows 224..255, the last quarter of query tile 1) skips the pair: the
          // dK / dV MMAs read the live 16-row steps only (qsteps), and the dQ rows it would feed are never stored.
          const bool active = col0 < ku && (NT == 1 ? 0 : t * 128) + q * 32 < (NT == 1 ? args.rows : it.nq);
          uint32_t pp[16], dd[16];              // 32 columns of P and dS, packed bf16 pairs
          if (active) {                         // warp-uniform: col0 depends on the warp only
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t sv[16], dv[16];
              tmem_ld_32x32b_x16(T_S + lane_off + col0 + h * 16, sv);
              tmem_ld_32x32b_x16(T_DP + lane_off + col0 + h * 16, dv);
              tmem_ld_wait();
              const int gcol = u * 128 + col0 + h * 16;
              // NT == 2 needs no key masking: K / V rows >= N are zero-filled by TMA, so whatever P and dS hold
              // in those columns only reaches dQ through zero K rows and dK / dV rows the TMA store clips;
              // padded query rows have lse2 = +inf, hence P = dS = 0. NT == 1 (packed sequences): classify the
              // 16-key chunk against the row's own sequence [lo, hi) -- 0 outside, 1 inside, 2 straddling.
              int cls = 1;
              if (NT == 1) cls = (gcol + 16 <= lo[t] || gcol >= hi[t]) ? 0 : ((gcol >= lo[t] && gcol + 16 <= hi[t]) ? 1 : 2);
              if (cls == 0) {
#pragma unroll
                for (int e = 0; e < 8; ++e) { pp[h * 8 + e] = 0u; dd[h * 8 + e] = 0u; }
              } else if (cls == 1) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * e]), args.scale_log2, -lse2[t]));
                  const float p1 = ex2_approx(fmaf(__uint_as_float(sv[2 * e + 1]), args.scale_log2, -lse2[t]));
                  const float d0 = p0 * ((__uint_as_float(dv[2 * e]) - delta[t]) * args.scale);
                  const float d1 = p1 * ((__uint_as_float(dv[2 * e + 1]) - delta[t]) * args.scale);
                  pp[h * 8 + e] = pack_bf16x2(p0, p1);
                  dd[h * 8 + e] = pack_bf16x2(d0, d1);
                }
              } else {
                for (int e = 0; e < 8; ++e) {
                  float p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * e]), args.scale_log2, -lse2[t]));
                  float p1 = ex2_approx(fmaf(__uint_as_float(sv[2 * e + 1]), args.scale_log2, -lse2[t]));
                  if (!(gcol + 2 * e >= lo[t] && gcol + 2 * e < hi[t])) p0 = 0.f;
                  if (!(gcol + 2 * e + 1 >= lo[t] && gcol + 2 * e + 1 < hi[t])) p1 = 0.f;
                  const float d0 = p0 * ((__uint_as_float(dv[2 * e]) - delta[t]) * args.scale);
                  const float d1 = p1 * ((__uint_as_float(dv[2 * e + 1]) - delta[t]) * args.scale);
                  pp[h * 8 + e] = pack_bf16x2(p0, p1);
                  dd[h * 8 + e] = pack_bf16x2(d0, d1);
                }
              }
            }
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_sdp_free);
          lap(2);
          if (gp > 0) mbar_wait(bar_mma, (gp - 1) & 1);  // previous MMAs done with sP / sdS
          if (active) {
            const uint32_t pc = smem_u32(sP) + (qc >> 1) * TILE_BYTES;
            const uint32_t dc = smem_u32(sdS) + (qc >> 1) * TILE_BYTES;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t off = sw128_offset(r, (qc & 1) * 4 + j);
              sts128(pc + off, make_uint4(pp[4 * j], pp[4 * j + 1], pp[4 * j + 2], pp[4 * j + 3]));
              sts128(dc + off, make_uint4(dd[4 * j], dd[4 * j + 1], dd[4 * j + 2], dd[4 * j + 3]));
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_pds);
          lap(3);

        }
      }
      if (prof_on) atomicAdd(args.prof + 8, 1ull);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// forward, streaming variant: any sequence length > 128; key / value blocks of 128 tokens streamed through smem
// ------------------------------------------------------------------------------------------------
// One 128-row query tile per CTA at a time (persistent over (sequence, head, query tile) items, query tile fastest so
// that the CTAs running at the same time share K / V in L2). The keys are walked in blocks of 128 with an online
// softmax. Two softmax TEAMS of 128 threads (one thread per query row, no cross-thread reductions) take the even and
// the odd key blocks; each keeps its own running (max, sum, O) and the two are merged once per item through shared
// memory (split-softmax merge). TMEM: three S buffers of 128 columns (P, bf16, overwrites the consumed columns of its
// own S buffer and is the A operand of P.V) + one 64-column O buffer per team. The MMA issuer keeps S three blocks
// ahead of P.V -- across item boundaries -- so a team never waits for scores, and P.V of block j writes a FRESH
// accumulator that the team folds into its registers (O_acc = O_acc * alpha + O_j) while it is already working on
// block j + 2: no TMEM read-modify-write, no correction warps, nothing on the softmax's critical path but exp2.
//   warps 0-3 / 4-7 : teams A / B      warp 8 : TMA producer (Q double-buffered per item; K ring, V ring)
//   warp 9 : tcgen05.mma issuer (+ TMEM allocation)
// Measured alternatives (profiles/README.md, round 2): two threads per row per team (16 softmax warps, one named
// barrier per block for the half-row maxima): same speed -- the teams lock in phase (both in their exp2 pass, then
// both out of it), the MUFU pipe idles ~60 % either way; ONE team of 16 warps with the next block's max pass riding
// inside the exp pass: 234 us (every step of the per-block chain is exposed); exp2 passes taking turns: 236 us.
constexpr int FS_THREADS = 10 * 32;
constexpr bool FS_EXP_TURNS = false;  // measured: exp2 passes taking turns in block order 201 -> 236 us at N = 785 (one warp per scheduler cannot fill the MUFU pipe alone)
constexpr int FS_KSTAGES = 4;
constexpr int FS_VSTAGES = 4;
constexpr int FS_SCRATCH_BYTES = 66 * 128 * 4;  // team B -> team A: O[64], max, sum per row, column-major
constexpr int FS_SMEM_BYTES = (2 + FS_KSTAGES + FS_VSTAGES) * TILE_BYTES + FS_SCRATCH_BYTES + 512 /*barriers*/ + 1024 /*align*/;

__global__ void __launch_bounds__(FS_THREADS, 1)
attention_fwd_stream_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                            const AttnArgs args, const int num_items, const int nqt, const int nkb) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                                   // [2] tiles
  uint8_t* sK = sQ + 2 * TILE_BYTES;                    // [FS_KSTAGES]
  uint8_t* sV = sK + FS_KSTAGES * TILE_BYTES;           // [FS_VSTAGES]
  float* scratch = reinterpret_cast<float*>(sV + FS_VSTAGES * TILE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(scratch) + FS_SCRATCH_BYTES);
  uint64_t* q_full = bars;                  // [2]
  uint64_t* q_free = bars + 2;              // [2] the item's output store has been read out of the (dead) Q tile
  uint64_t* k_full = bars + 4;              // [FS_KSTAGES]
  uint64_t* k_free = k_full + FS_KSTAGES;   // [FS_KSTAGES] S = Q K^T of the block retired
  uint64_t* v_full = k_free + FS_KSTAGES;   // [FS_VSTAGES]
  uint64_t* v_free = v_full + FS_VSTAGES;   // [FS_VSTAGES] P V of the block retired
  uint64_t* s_ready = v_free + FS_VSTAGES;  // [3] per S buffer
  uint64_t* p_ready = s_ready + 3;          // [3] per S buffer: the team wrote P (and holds the previous O in registers)
  uint64_t* o_ready = p_ready + 3;          // [2] per team
  uint64_t* b_done = o_ready + 2;           // team B's partial result of an item is in the scratch
  uint64_t* scratch_free = b_done + 1;      // team A has read it
  uint64_t* exp_turn = scratch_free + 1;    // the exp2 pass of block g is over (one phase per block, in block order)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(exp_turn + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = (num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  auto item_of = [&](int k) { return static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x); };
  auto keys_of = [&](int j) { return min(128, args.N - 128 * j); };           // real keys in block j
  auto keys_n_of = [&](int j) { return (keys_of(j) + 15) & ~15; };            // MMA extent (TMA zero-fills the rest)

  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_free[i], 1);
      mbar_init(&o_ready[i], 1);
    }
    for (int i = 0; i < FS_KSTAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_free[i], 1); }
    for (int i = 0; i < FS_VSTAGES; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_free[i], 1); }
    for (int i = 0; i < 3; ++i) { mbar_init(&s_ready[i], 1); mbar_init(&p_ready[i], 4); }   // per-warp arrivals (lane 0 after __syncwarp)
    mbar_init(b_done, 4);
    mbar_init(scratch_free, 4);
    mbar_init(exp_turn, 4);
    fence_barrier_init();
  }
  if (warp == 9) {
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform for the compiler (MMA issuer)
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one_sync()) {
      int g = 0;
      for (int k = 0; k < n_my; ++k) {
        const int item = item_of(k);
        const int qt = item % nqt, bh = item / nqt;
        const int head = bh % args.H, b = bh / args.H;
        mbar_wait(&q_free[k & 1], ((k >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[k & 1], TILE_BYTES);
        tma_load_3d(sQ + (k & 1) * TILE_BYTES, &tmQKV, &q_full[k & 1], head * 64, qt * 128, b);
        for (int j = 0; j < nkb; ++j, ++g) {
          const int ks = g % FS_KSTAGES, vs = g % FS_VSTAGES;
          mbar_wait(&k_free[ks], ((g / FS_KSTAGES) & 1) ^ 1);
          mbar_expect_tx(&k_full[ks], TILE_BYTES);
          tma_load_3d(sK + ks * TILE_BYTES, &tmQKV, &k_full[ks], (args.H + head) * 64, j * 128, b);
          mbar_wait(&v_free[vs], ((g / FS_VSTAGES) & 1) ^ 1);
          mbar_expect_tx(&v_full[vs], TILE_BYTES);
          tma_load_3d(sV + vs * TILE_BYTES, &tmQKV, &v_full[vs], (2 * args.H + head) * 64, j * 128, b);
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer (elect.sync + one-add descriptors,
    // see the two-tile kernel: the issue code, not the tensor pipe, used to bound the N = 64 P.V MMAs)
    if (elect_one_sync()) {
      const uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);
      const uint32_t lo0 = (smem_u32(smem) & 0x3FFFFu) >> 4;
      auto tile_lo = [&](uint8_t* tile) -> uint32_t { return lo0 + (static_cast<uint32_t>(tile - smem) >> 4); };
      const int total = n_my * nkb;
      int gs = 0, gp = 0;        // blocks whose S / P.V have been issued
      int ks_item = 0, ks_j = 0; // (item, block) of gs
      int kp_j = 0;              // block of gp within its item
      while (gp < total) {
        while (gs < total && gs < gp + 3) {
          // S(gs) overwrites S buffer gs % 3 = the P of block gs - 3, whose P.V was issued before (in-order pipe)
          if (ks_j == 0) mbar_wait(&q_full[ks_item & 1], (ks_item >> 1) & 1);
          const int ks = gs % FS_KSTAGES;
          mbar_wait(&k_full[ks], (gs / FS_KSTAGES) & 1);
          tcgen05_fence_after();
          const uint32_t idesc_s = make_idesc_bf16(128, keys_n_of(ks_j), false, false);
          const uint32_t a0 = tile_lo(sQ + (ks_item & 1) * TILE_BYTES), b0 = tile_lo(sK + ks * TILE_BYTES);
          const uint32_t t_s = tmem_base + (gs % 3) * 128;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ss(t_s, sw128_desc_at(a0, kk * 32, 16, 1024), sw128_desc_at(b0, kk * 32, 16, 1024), idesc_s, kk > 0);
          umma_commit(&k_free[ks]);
          umma_commit(&s_ready[gs % 3]);
          ++gs;
          if (++ks_j == nkb) { ks_j = 0; ++ks_item; }
        }
        const int team = kp_j & 1, vs = gp % FS_VSTAGES;
        mbar_wait(&p_ready[gp % 3], (gp / 3) & 1);
        mbar_wait(&v_full[vs], (gp / FS_VSTAGES) & 1);
        tcgen05_fence_after();
        const uint32_t v0 = tile_lo(sV + vs * TILE_BYTES);
        const int ksteps = keys_n_of(kp_j) / 16;
        const uint32_t t_o = tmem_base + 384 + team * 64, t_p = tmem_base + (gp % 3) * 128;
#pragma unroll
        for (int s = 0; s < 8; ++s)
          if (s < ksteps) umma_bf16_ts(t_o, t_p + 8 * s, sw128_desc_at(v0, s * 2048, 8192, 1024), idesc_o, s > 0);
        umma_commit(&v_free[vs]);
        umma_commit(&o_ready[team]);
        ++gp;
        if (++kp_j == nkb) kp_j = 0;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax teams
    const int team = warp >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t o_addr = tmem_base + lane_off + 384 + team * 64;
    float o_acc[64];
    float m_run = -INFINITY, l_run = 0.f, alpha_pend = 0.f;
    bool pend = false;
    int pend_k = 0;          // item the pending O (and the accumulators) belong to
    bool pend_valid = false; // ... and whether this warp had live rows in it
    uint32_t npv = 0, nfin = 0;
#pragma unroll
    for (int i = 0; i < 64; ++i) o_acc[i] = 0.f;
    // developer instrumentation (args.prof, 8 counters per team): cycles of the team's first thread in
    // [0] wait for S, [1] max pass, [2] wait for O, [3] fold O, [4] finalize, [5] exp pass, [6] blocks, [7] whole loop
    const bool prof_on = args.prof != nullptr && (threadIdx.x & 127) == 0;
    long long tp0 = prof_on ? clock64() : 0;
    const long long tp_begin = tp0;
    auto lap = [&](int idx) {
      if (prof_on) {
        const long long now = clock64();
        atomicAdd(args.prof + team * 8 + idx, static_cast<unsigned long long>(now - tp0));
        tp0 = now;
      }
    };

    // fold the P.V result of this team's previous block into the register accumulator
    auto resolve = [&]() {
      lap(1);
      mbar_wait(&o_ready[team], npv & 1);
      ++npv;
      tcgen05_fence_after();
      lap(2);
      if (pend_valid) {
        uint32_t va[32], vb[32];
        tmem_ld_32x32b_x32(o_addr, va);
        tmem_ld_32x32b_x32(o_addr + 32, vb);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[i] = fmaf(o_acc[i], alpha_pend, __uint_as_float(va[i]));
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[32 + i] = fmaf(o_acc[32 + i], alpha_pend, __uint_as_float(vb[i]));
      }
      pend = false;
      lap(3);
    };
    // an item is complete for this team: B parks its partial result, A merges, normalises and stores
    auto finalize = [&]() {
      const int item = item_of(pend_k);
      const int qt = item % nqt, bh = item / nqt;
      const int head = bh % args.H, b = bh / args.H;
      if (team == 1) {
        mbar_wait(scratch_free, (nfin & 1) ^ 1);
        if (pend_valid) {
#pragma unroll
          for (int i = 0; i < 64; ++i) scratch[i * 128 + r] = o_acc[i];
          scratch[64 * 128 + r] = m_run;
          scratch[65 * 128 + r] = l_run;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(b_done);
      } else {
        mbar_wait(b_done, nfin & 1);
        uint8_t* stg = sQ + (pend_k & 1) * TILE_BYTES;  // every S = Q K^T of the item has retired: Q is dead
        if (pend_valid) {
          const float mb = scratch[64 * 128 + r], lb = scratch[65 * 128 + r];
          const float m = fmaxf(m_run, mb);
          const float wa = ex2_approx(m_run - m), wb = ex2_approx(mb - m);
          const float l = l_run * wa + lb * wb;
          const float inv = l > 0.f ? 1.f / l : 0.f;
          const float sa = wa * inv, sb = wb * inv;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = o_acc[8 * c + e] * sa + scratch[(8 * c + e) * 128 + r] * sb;
            uint4 pk;
            pk.x = pack_bf16x2(o[0], o[1]);
            pk.y = pack_bf16x2(o[2], o[3]);
            pk.z = pack_bf16x2(o[4], o[5]);
            pk.w = pack_bf16x2(o[6], o[7]);
            *reinterpret_cast<uint4*>(stg + sw128_offset(r, c)) = pk;
          }
          const int n = qt * 128 + r;
          if (n < args.N) args.lse2[(static_cast<long long>(b) * args.H + head) * args.Ns + n] = m + log2f(l);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(scratch_free);
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (warp == 0 && lane == 0) {
          tma_store_3d(&tmO, stg, head * 64, qt * 128, b);
          tma_store_commit();
          tma_store_wait_read<0>();
          mbar_arrive(&q_free[pend_k & 1]);
        }
      }
      ++nfin;
      lap(4);
    };

    for (int k = 0; k < n_my; ++k) {
      const int item = item_of(k);
      const int qt = item % nqt;
      const bool warp_valid = qt * 128 + q * 32 < args.N;  // rows >= N: scores of zero-filled queries, never stored
      for (int j = team; j < nkb; j += 2) {
        const int g = k * nkb + j, buf = g % 3;
        const bool first = j == team;
        const int kn = keys_n_of(j);
        const int n32 = kn >> 5;              // full 32-column groups; a 16-column tail only in the last block
        const bool tail16 = (kn & 16) != 0;
        const uint32_t s_addr = tmem_base + lane_off + buf * 128;
        lap(5);
        mbar_wait(&s_ready[buf], (g / 3) & 1);
        tcgen05_fence_after();
        lap(0);
        // ---- pass 1: row max of the raw scores (zero-filled pad keys score exactly 0: harmless in the max). The
        // thread owns the whole row, so the only latency to hide is the TMEM load's: two 32-column loads in flight
        float mx = -INFINITY;
        auto max32 = [&](const uint32_t (&v)[32]) {
          float m0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
          float m1 = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
#pragma unroll
          for (int i = 4; i < 32; i += 4) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
          }
          mx = fmaxf(mx, fmaxf(m0, m1));
        };
        if (warp_valid) {
          uint32_t va[32], vb[32];
          for (int c = 0; c < n32; c += 2) {
            tmem_ld_32x32b_x32(s_addr + c * 32, va);
            if (c + 1 < n32) tmem_ld_32x32b_x32(s_addr + (c + 1) * 32, vb);
            tmem_ld_wait();
            max32(va);
            if (c + 1 < n32) max32(vb);
          }
          if (tail16) {
            uint32_t t[16];
            tmem_ld_32x32b_x16(s_addr + n32 * 32, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(t[i]));
          }
        }
        // ---- the P.V result of this team's previous block is certainly there by now: fold it in; at an item
        // boundary that completes the previous item
        if (pend) {
          resolve();
          if (first) finalize();
        }
        lap(1);
        if (first) {
          m_run = -INFINITY;
          l_run = 0.f;
#pragma unroll
          for (int i = 0; i < 64; ++i) o_acc[i] = 0.f;
        }
        // ---- pass 2: p = 2^(s * c - m) -> bf16 pairs over the consumed columns of the same S buffer. The exp2 passes
        // take turns in block order: the MUFU pipe is what both teams compete for, and two passes running side by side
        // at half speed lock the teams in phase (both then sit in their MUFU-free phases together, the pipe idle)
        if (FS_EXP_TURNS && g > 0) mbar_wait(exp_turn, (g - 1) & 1);
        lap(1);
        if (warp_valid) {
          const float m_new = fmaxf(m_run, mx * args.scale_log2);
          const float alpha = m_run == -INFINITY ? 0.f : ex2_approx(m_run - m_new);
          float sum0 = 0.f, sum1 = 0.f;
          auto exp16 = [&](const uint32_t (&v)[16], uint32_t col) {   // 16 scores -> 8 packed columns of P at `col`
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), args.scale_log2, -m_new));
              const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), args.scale_log2, -m_new));
              sum0 += p0;
              sum1 += p1;
              pk[i] = pack_bf16x2(p0, p1);
            }
            tmem_st_32x32b_x8(col, pk);
          };
          const int nch = kn >> 4;
          uint32_t wa[16], wb[16];
          tmem_ld_32x32b_x16(s_addr, wa);
          for (int c = 0; c < nch; c += 2) {
            tmem_ld_wait();
            if (c + 1 < nch) tmem_ld_32x32b_x16(s_addr + (c + 1) * 16, wb);
            exp16(wa, s_addr + c * 8);
            if (c + 1 < nch) {
              tmem_ld_wait();
              if (c + 2 < nch) tmem_ld_32x32b_x16(s_addr + (c + 2) * 16, wa);
              exp16(wb, s_addr + (c + 1) * 8);
            }
          }
          float sum = sum0 + sum1;
          sum -= static_cast<float>(kn - keys_of(j)) * ex2_approx(-m_new);  // the zero-filled pad keys of the last block
          l_run = fmaf(l_run, alpha, sum);
          m_run = m_new;
          alpha_pend = alpha;
          if (FS_EXP_TURNS) { __syncwarp(); if (lane == 0) mbar_arrive(exp_turn); }
          tmem_st_wait();
        } else if (FS_EXP_TURNS) {
          __syncwarp();
          if (lane == 0) mbar_arrive(exp_turn);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[buf]);
        pend = true;
        pend_k = k;
        pend_valid = warp_valid;
        if (prof_on) atomicAdd(args.prof + team * 8 + 6, 1ull);
      }
    }
    lap(5);
    if (pend) {
      resolve();
      finalize();
    }
    if (prof_on) atomicAdd(args.prof + team * 8 + 7, static_cast<unsigned long long>(clock64() - tp_begin));
    if (warp == 0 && lane == 0) tma_store_wait_all<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

static unsigned long long* g_attn_prof = nullptr;
// forward kernel choice: 0 = streaming kernel for N > 256 only (two-tile kernel for 129..256), 1 = streaming kernel for
// every N > 128, -1 = never (N > 256 falls back to the block decomposition with an lse merge; developer A/B)
static int g_attn_stream = 0;
static int g_attn_dephase = 0;  // two-tile forward: 1 = event-driven MMA issue, slots half a period apart (measured
                                // slower: 107 vs 102 us at B = 512, N = 197); 0 = round-locked issue; > 1 = 1 with that many ns of back-off

// Nq / Nk: query and key tokens of this launch (equal, and equal to the stride Ns, on the main path).
static int setup_args(AttnArgs& a, int B, int Nq, int Nk, int Ns, int H, float scale, bool blocked, int& nt, int& groups) {
  a.prof = g_attn_prof;
  a.dephase = g_attn_dephase;
  B200SSL_CHECK(Nq >= 1 && Nq <= 256 && Nk >= 1 && Nk <= 256, -2, "attention: block of %d x %d tokens unsupported (1..256)",
                Nq, Nk);
  a.B = B; a.N = Nq; a.Nk = Nk; a.Ns = Ns; a.H = H;
  a.acc_dq = 0; a.acc_dkv = 0;
  a.nblk = 1;
  // blocks of a long sequence always take the two-tile kernels (no packing): their zero-padding rules cover any
  // Nq, Nk <= 256, including Nq != Nk
  nt = (!blocked && Nq <= 128) ? 1 : 2;
  a.G = nt == 1 ? 128 / Nq : 1;
  if (a.G > B) a.G = B;
  a.rows = nt == 1 ? a.G * Nq : Nq;
  a.keys_n = ((nt == 1 ? a.rows : Nk) + 15) / 16 * 16;
  a.scale = scale;
  a.scale_log2 = scale * 1.4426950408889634f;
  groups = (B + a.G - 1) / a.G;
  return 0;
}

// 3-D map (cols, rows of this block, B) over a [B, Ns, cols] bf16 tensor whose base already points at the block's
// first row; TMA clips rows >= n_rows and batches >= B.
static int make_bnd_map(CUtensorMap* tm, const void* base, int cols, int n_rows, int Ns, int B, int nt, int G) {
  uint64_t dims[3] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(n_rows), static_cast<uint64_t>(B)};
  uint64_t strides[3] = {2, static_cast<uint64_t>(cols) * 2, static_cast<uint64_t>(cols) * 2 * Ns};
  uint32_t box[3] = {64, static_cast<uint32_t>(nt == 1 ? n_rows : 128), static_cast<uint32_t>(nt == 1 ? G : 1)};
  return make_tensor_map(tm, base, 2, 3, dims, strides, box, 128);
}

// one launch: queries [q0, q0 + Nq) against keys [k0, k0 + Nk) of every (batch, head)
static int attention_fwd_block(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse2, int B, int Ns, int H, int q0, int Nq,
                               int k0, int Nk, float scale, bool blocked, cudaStream_t stream) {
  AttnArgs a{};
  int nt, groups;
  if (int rc = setup_args(a, B, Nq, Nk, Ns, H, scale, blocked, nt, groups)) return rc;
  a.lse2 = lse2 + q0;
  const long long row = 3LL * H * 64;
  CUtensorMap tq, tkv, to;
  if (int rc = make_bnd_map(&tq, qkv + q0 * row, 3 * H * 64, Nq, Ns, B, nt, a.G)) return rc;
  if (int rc = make_bnd_map(&tkv, qkv + k0 * row, 3 * H * 64, Nk, Ns, B, nt, a.G)) return rc;
  if (int rc = make_bnd_map(&to, out + static_cast<long long>(q0) * H * 64, H * 64, Nq, Ns, B, nt, a.G)) return rc;
  const int num_items = groups * H;
  const int num_rounds = nt == 2 ? num_items : (num_items + 1) / 2;
  const int grid = num_rounds < sm_count() ? num_rounds : sm_count();
  const int smem = 2 * FWD_BUF_BYTES + 8192 + 1024;
  if (nt == 1) {
    static bool cfg = false;
    if (!cfg) {
      B200SSL_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cfg = true;
    }
    B200SSL_CUDA(launch_pdl(attention_fwd_kernel<1>, dim3(grid), dim3(FWD_THREADS), smem, stream, 1, tq, tkv, to, a, num_items));
  } else {
    static bool cfg = false;
    if (!cfg) {
      B200SSL_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cfg = true;
    }
    B200SSL_CUDA(launch_pdl(attention_fwd_kernel<2>, dim3(grid), dim3(FWD_THREADS), smem, stream, 1, tq, tkv, to, a, num_items));
  }
  return 0;
}

// streaming forward (any N > 128): one launch for the whole sequence, no workspace
static int attention_fwd_stream(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse2, int B, int N, int H, float scale,
                                cudaStream_t stream) {
  AttnArgs a{};
  a.prof = g_attn_prof;
  a.B = B; a.N = N; a.Nk = N; a.Ns = N; a.H = H;
  a.G = 1; a.rows = N; a.keys_n = 0;
  a.scale = scale;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.lse2 = lse2;
  CUtensorMap tq, to;
  if (int rc = make_bnd_map(&tq, qkv, 3 * H * 64, N, N, B, 2, 1)) return rc;
  if (int rc = make_bnd_map(&to, out, H * 64, N, N, B, 2, 1)) return rc;
  const int nqt = (N + 127) / 128, nkb = (N + 127) / 128;
  const long long items = static_cast<long long>(B) * H * nqt;
  B200SSL_CHECK(items * nkb < (1LL << 30), -2, "attention: problem too large for the streaming forward (B*H*tiles = %lld)", items);
  const int grid = items < sm_count() ? static_cast<int>(items) : sm_count();
  static bool cfg = false;
  if (!cfg) {
    B200SSL_CUDA(cudaFuncSetAttribute(attention_fwd_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM_BYTES));
    cfg = true;
  }
  B200SSL_CUDA(launch_pdl(attention_fwd_stream_kernel, dim3(grid), dim3(FS_THREADS), FS_SMEM_BYTES, stream, 1, tq, to, a,
                          static_cast<int>(items), nqt, nkb));
  return 0;
}

static int attention_bwd_block(const __nv_bfloat16* qkv, const __nv_bfloat16* out, const __nv_bfloat16* dout, const float* lse2,
                               __nv_bfloat16* dqkv, int B, int Ns, int H, int q0, int Nq, int k0, int Nk, float scale,
                               bool blocked, bool acc_dq, bool acc_dkv, cudaStream_t stream) {
  AttnArgs a{};
  int nt, groups;
  if (int rc = setup_args(a, B, Nq, Nk, Ns, H, scale, blocked, nt, groups)) return rc;
  a.acc_dq = acc_dq; a.acc_dkv = acc_dkv;
  a.lse2 = const_cast<float*>(lse2) + q0;
  a.out = out + static_cast<long long>(q0) * H * 64;
  a.dout = dout + static_cast<long long>(q0) * H * 64;
  const long long row = 3LL * H * 64;
  a.qkv_q = qkv + q0 * row;
  a.qkv_k = qkv + k0 * row;
  a.dq = dqkv + q0 * row;
  a.dkv = dqkv + k0 * row;
  a.dq_rows = Nq;
  a.dkv_rows = Nk;
  CUtensorMap tq, tkv, tdo;
  if (int rc = make_bnd_map(&tq, qkv + q0 * row, 3 * H * 64, Nq, Ns, B, nt, a.G)) return rc;
  if (int rc = make_bnd_map(&tkv, qkv + k0 * row, 3 * H * 64, Nk, Ns, B, nt, a.G)) return rc;
  if (int rc = make_bnd_map(&tdo, a.dout, H * 64, Nq, Ns, B, nt, a.G)) return rc;
  const int num_items = groups * H;
  const int grid = num_items < sm_count() ? num_items : sm_count();
  if (nt == 1) {
    const int smem = 12 * TILE_BYTES + 1024 /*align*/ + 128 /*barriers*/ + 2 * 128 * 2 * 4 + 512 /*row table*/ + 256;
    static bool cfg = false;
    if (!cfg) {
      B200SSL_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cfg = true;
    }
    B200SSL_CUDA(launch_pdl(attention_bwd_kernel<1>, dim3(grid), dim3(BWD_THREADS), smem, stream, 1, tq, tkv, tdo, a, num_items));
  } else {
    const int smem = 12 * TILE_BYTES + 1024 /*align*/ + 128 /*barriers*/ + 2 * 256 * 2 * 4 + 256;
    static bool cfg = false;
    if (!cfg) {
      B200SSL_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cfg = true;
    }
    B200SSL_CUDA(launch_pdl(attention_bwd_kernel<2>, dim3(grid), dim3(BWD_THREADS), smem, stream, 1, tq, tkv, tdo, a, num_items));
  }
  return 0;
}

// backward for N > 256 in ONE launch: every (sequence, head, query block, key block) pair of 256 x 256 tokens is an item
// of the persistent two-tile kernel; each pair is exact given the row's global lse and delta, dQ accumulates over key
// blocks and dK / dV over query blocks through TMA reduce-add into a zeroed dqkv (sixteen launches per layer at 785
// tokens before: each with its own ramp and a last, partly filled round of items)
static int attention_bwd_paired(const __nv_bfloat16* qkv, const __nv_bfloat16* out, const __nv_bfloat16* dout, const float* lse2,
                                __nv_bfloat16* dqkv, int B, int N, int H, float scale, cudaStream_t stream) {
  AttnArgs a{};
  int nt, groups;
  if (int rc = setup_args(a, B, 256, 256, N, H, scale, true, nt, groups)) return rc;
  a.nblk = (N + 255) / 256;
  a.acc_dq = 1; a.acc_dkv = 1;
  a.lse2 = const_cast<float*>(lse2);
  a.out = out;
  a.dout = dout;
  a.qkv_q = qkv;
  a.qkv_k = qkv;
  a.dq = dqkv;
  a.dkv = dqkv;
  a.dq_rows = N;
  a.dkv_rows = N;
  CUtensorMap tq, tdo;
  if (int rc = make_bnd_map(&tq, qkv, 3 * H * 64, N, N, B, 2, 1)) return rc;
  if (int rc = make_bnd_map(&tdo, dout, H * 64, N, N, B, 2, 1)) return rc;
  B200SSL_CUDA(cudaMemsetAsync(dqkv, 0, static_cast<size_t>(B) * N * 3 * H * 64 * sizeof(__nv_bfloat16), stream));
  const long long items = static_cast<long long>(B) * H * a.nblk * a.nblk;
  B200SSL_CHECK(items < (1LL << 30), -2, "attention: problem too large for the paired backward (%lld items)", items);
  const int grid = items < sm_count() ? static_cast<int>(items) : sm_count();
  const int smem = 12 * TILE_BYTES + 1024 /*align*/ + 128 /*barriers*/ + 2 * 256 * 2 * 4 + 256;
  static bool cfg = false;
  if (!cfg) {
    B200SSL_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cfg = true;
  }
  B200SSL_CUDA(launch_pdl(attention_bwd_kernel<2>, dim3(grid), dim3(BWD_THREADS), smem, stream, 1, tq, tq, tdo, a,
                          static_cast<int>(items)));
  return 0;
}

// ---- long sequences (N > 256): key / query blocks of <= 256 tokens, partial softmax results merged by lse ----
// out[b, n, h, :] = sum_j 2^(lse_j - L) O_j[b, n, h, :],  L = log2 sum_j 2^lse_j ; one warp per (b, n, h)
__global__ void __launch_bounds__(256)
attention_combine_kernel(const __nv_bfloat16* __restrict__ o_part, const float* __restrict__ lse_part,
                         __nv_bfloat16* __restrict__ out, float* __restrict__ lse2, int B, int N, int H, int nblk) {
  const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(B) * N * H;
  if (w >= total) return;
  const int h = static_cast<int>(w % H);
  const long long bn = w / H;
  const int n = static_cast<int>(bn % N);
  const long long b = bn / N;
  const long long lse_idx = (b * H + h) * N + n;
  const long long o_idx = (bn * H + h) * 64 + 2 * lane;
  const long long o_stride = static_cast<long long>(B) * N * H * 64, l_stride = static_cast<long long>(B) * H * N;
  float m = -INFINITY;
  for (int j = 0; j < nblk; ++j) m = fmaxf(m, lse_part[j * l_stride + lse_idx]);
  float den = 0.f, a0 = 0.f, a1 = 0.f;
  for (int j = 0; j < nblk; ++j) {
    const float wj = ex2_approx(lse_part[j * l_stride + lse_idx] - m);
    const float2 o = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(o_part + j * o_stride + o_idx));
    den += wj;
    a0 += wj * o.x;
    a1 += wj * o.y;
  }
  const float inv = 1.f / den;
  *reinterpret_cast<uint32_t*>(out + o_idx) = pack_bf16x2(a0 * inv, a1 * inv);
  if (lane == 0) lse2[lse_idx] = m + log2f(den);
}

static int num_blocks_for(int N) { return (N + 255) / 256; }
static void block_range(int N, int nblk, int i, int& start, int& len) {  // near-equal blocks
  const int base = N / nblk, rem = N % nblk;
  start = i * base + (i < rem ? i : rem);
  len = base + (i < rem ? 1 : 0);
}

}  // namespace b200ssl

using namespace b200ssl;

// Developer instrumentation: device buffer of 16 uint64 counters (2 slots x 8) the forward kernel's first softmax
// thread of each slot adds to: cycles in [0] wait for S, [1] max pass, [2] barrier, [3] exp pass, [4] barrier + lse,
// [5] wait for O, [6] epilogue + store, [7] tiles. NULL = off. The backward kernel's leader math thread adds:
// [0] wait row constants, [1] wait S/dP, [2] P/dS math, [3] wait previous MMAs + smem writes, [7] item turnaround,
// [8] items ([4..6] were the math warps' dK/dV waits and stores before the auxiliary group took the drain over).
extern "C" int b200ssl_set_attn_prof(void* counters) {
  b200ssl::g_attn_prof = static_cast<unsigned long long*>(counters);
  return 0;
}

// bytes of scratch b200ssl_attention_fwd_ws needs (0 for N <= 256): per key block a partial output and lse
extern "C" int b200ssl_set_attn_dephase(int on) {
  b200ssl::g_attn_dephase = on;
  return 0;
}

extern "C" int b200ssl_set_attn_stream(int mode) {
  B200SSL_CHECK(mode >= -1 && mode <= 1, -2, "attention stream mode must be -1, 0 or 1");
  b200ssl::g_attn_stream = mode;
  return 0;
}

extern "C" long long b200ssl_attention_fwd_workspace_bytes(int B, int N, int H) {
  if (N <= 256 || b200ssl::g_attn_stream >= 0) return 0;
  const long long nblk = num_blocks_for(N);
  return nblk * (static_cast<long long>(B) * N * H * 64 * 2 + static_cast<long long>(B) * H * N * 4);
}

extern "C" int b200ssl_attention_fwd_ws(const void* qkv_, void* out_, float* lse2, int B, int N, int H, int head_dim,
                                        float scale, void* workspace, long long workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200SSL_CHECK(head_dim == 64, -2, "attention: head_dim %d unsupported (64 only)", head_dim);
  B200SSL_CHECK(B > 0 && H > 0 && N > 0, -2, "attention: empty problem");
  const __nv_bfloat16* qkv = static_cast<const __nv_bfloat16*>(qkv_);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(out_);
  if (N > 128 && (g_attn_stream == 1 || (g_attn_stream == 0 && N > 256))) {
    B200SSL_CHECK(N <= 65536, -2, "attention: sequence length %d unsupported (1..65536)", N);
    return attention_fwd_stream(qkv, out, lse2, B, N, H, scale, stream);
  }
  if (N <= 256) return attention_fwd_block(qkv, out, lse2, B, N, H, 0, N, 0, N, scale, false, stream);
  B200SSL_CHECK(N <= 4096, -2, "attention: sequence length %d unsupported (1..4096)", N);
  const long long need = b200ssl_attention_fwd_workspace_bytes(B, N, H);
  B200SSL_CHECK(workspace != nullptr && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 127) == 0, -2,
                "attention: N=%d needs a 128B-aligned workspace of %lld bytes (b200ssl_attention_fwd_workspace_bytes)", N, need);
  const int nblk = num_blocks_for(N);
  const long long o_elems = static_cast<long long>(B) * N * H * 64, l_elems = static_cast<long long>(B) * H * N;
  __nv_bfloat16* o_part = static_cast<__nv_bfloat16*>(workspace);
  float* l_part = reinterpret_cast<float*>(o_part + nblk * o_elems);
  for (int j = 0; j < nblk; ++j) {
    int k0, nk;
    block_range(N, nblk, j, k0, nk);
    for (int i = 0; i < nblk; ++i) {
      int q0, nq;
      block_range(N, nblk, i, q0, nq);
      if (int rc = attention_fwd_block(qkv, o_part + j * o_elems, l_part + j * l_elems, B, N, H, q0, nq, k0, nk, scale, true,
                                       stream))
        return rc;
    }
  }
  const long long warps = static_cast<long long>(B) * N * H;
  attention_combine_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, stream>>>(o_part, l_part, out, lse2, B, N,
                                                                                               H, nblk);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200ssl_attention_fwd(const void* qkv, void* out, float* lse2, int B, int N, int H, int head_dim,
                                     float scale, void* stream) {
  return b200ssl_attention_fwd_ws(qkv, out, lse2, B, N, H, head_dim, scale, nullptr, 0, stream);
}

extern "C" int b200ssl_attention_bwd(const void* qkv_, const void* out_, const void* dout_, const float* lse2,
                                     void* dqkv_, int B, int N, int H, int head_dim, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200SSL_CHECK(head_dim == 64, -2, "attention: head_dim %d unsupported (64 only)", head_dim);
  B200SSL_CHECK(B > 0 && H > 0 && N > 0, -2, "attention: empty problem");
  const __nv_bfloat16* qkv = static_cast<const __nv_bfloat16*>(qkv_);
  const __nv_bfloat16* out = static_cast<const __nv_bfloat16*>(out_);
  const __nv_bfloat16* dout = static_cast<const __nv_bfloat16*>(dout_);
  __nv_bfloat16* dqkv = static_cast<__nv_bfloat16*>(dqkv_);
  // the drain warps write whole 128-byte rows with 256-bit stores
  B200SSL_CHECK((reinterpret_cast<uintptr_t>(dqkv) & 31) == 0, -2, "attention_bwd: dqkv must be 32-byte aligned");
  if (N <= 256) return attention_bwd_block(qkv, out, dout, lse2, dqkv, B, N, H, 0, N, 0, N, scale, false, false, false, stream);
  B200SSL_CHECK(N <= 4096, -2, "attention: sequence length %d unsupported (1..4096)", N);
  // three or more blocks per dimension: ONE launch over all block pairs (N = 785, B = 64: 723 -> 565 us). With two blocks
  // the four launches of near-equal blocks stay ahead (N = 257, B = 256: 476 vs 610 us -- the second 256-token block
  // would hold a single token, and the paired mode pays a memset plus reduce-adds where the first block stores)
  if (g_attn_stream >= 0 && N > 512) return attention_bwd_paired(qkv, out, dout, lse2, dqkv, B, N, H, scale, stream);
  // one launch per block pair (also the developer A/B path, b200ssl_set_attn_stream(-1))
  // every (query block, key block) pair is exact given the row's global lse and delta; dQ accumulates over key
  // blocks, dK / dV over query blocks (TMA reduce-add in bf16 after the first, plain store)
  const int nblk = num_blocks_for(N);
  for (int i = 0; i < nblk; ++i) {
    int q0, nq;
    block_range(N, nblk, i, q0, nq);
    for (int j = 0; j < nblk; ++j) {
      int k0, nk;
      block_range(N, nblk, j, k0, nk);
      if (int rc = attention_bwd_block(qkv, out, dout, lse2, dqkv, B, N, H, q0, nq, k0, nk, scale, true, j > 0, i > 0, stream))
        return rc;
    }
  }
  return 0;
}
