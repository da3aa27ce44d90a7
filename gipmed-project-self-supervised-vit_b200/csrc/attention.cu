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
// forward: one 128-row query tile per CTA, two CTAs per SM
// ------------------------------------------------------------------------------------------------
// smem : Q tile (16K, re-used as the output staging tile) | K (NT x 16K) | V (NT x 16K) | barriers
// TMEM : S fp32 in columns [0, keys_n); after the softmax has read it, P (bf16 pairs, the A operand of
//        the second MMA, read straight from TMEM) overwrites columns [0, keys_n/2) and O accumulates in
//        the last 64 columns of the allocation, which lie in the dead tail of S.
// warps: 0 = control (TMA + tcgen05.mma issue), 1..8 = softmax/epilogue, two threads per query row.
constexpr int FWD_THREADS = 32 + 256;

template <int NT>
__global__ void __launch_bounds__(FWD_THREADS, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                     const AttnArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;
  uint8_t* sV = sK + NT * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NT * TILE_BYTES);
  uint64_t* bar_load = bars;
  uint64_t* bar_s = bars + 1;
  uint64_t* bar_p = bars + 2;
  uint64_t* bar_o = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* xchg = reinterpret_cast<float*>(bars + 6);  // [2][128]
  constexpr int TMEM_COLS = 128 * NT;
  constexpr int O_COL = TMEM_COLS - 64;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = NT == 2 ? blockIdx.x >> 1 : blockIdx.x;
  const int t = NT == 2 ? (blockIdx.x & 1) : 0;  // query tile within the sequence
  const int head = item % args.H;
  const int b0 = (item / args.H) * args.G;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV);
      tma_prefetch_desc(&tmO);
      mbar_init(bar_load, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 256);
      mbar_init(bar_o, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  if (NT == 1) {
    // V rows the TMA box never writes must not feed NaN bit patterns into P(=0) x V
    const int first = args.rows * 128, last = args.keys_n * 128;
    for (int i = first + threadIdx.x * 16; i < last; i += FWD_THREADS * 16)
      *reinterpret_cast<uint4*>(sV + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int cq = head * 64, ck = (args.H + head) * 64, cv = (2 * args.H + head) * 64;
      if (NT == 1) {
        mbar_expect_tx(bar_load, 3 * args.rows * 128);
        tma_load_3d(sQ, &tmQKV, bar_load, cq, 0, b0);
        tma_load_3d(sK, &tmQKV, bar_load, ck, 0, b0);
        tma_load_3d(sV, &tmQKV, bar_load, cv, 0, b0);
      } else {
        mbar_expect_tx(bar_load, (1 + 2 * NT) * TILE_BYTES);
        tma_load_3d(sQ, &tmQKV, bar_load, cq, t * 128, b0);
        for (int u = 0; u < NT; ++u) {
          tma_load_3d(sK + u * TILE_BYTES, &tmQKV, bar_load, ck, u * 128, b0);
          tma_load_3d(sV + u * TILE_BYTES, &tmQKV, bar_load, cv, u * 128, b0);
        }
      }
      mbar_wait(bar_load, 0);
      tcgen05_fence_after();
      const uint32_t idesc_s = make_idesc_bf16(128, args.keys_n, false, false);
      const uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);
      const uint32_t a0 = smem_u32(sQ), bk = smem_u32(sK), v0 = smem_u32(sV);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_ss(tmem_base, make_smem_desc_sw128(a0 + k * 32, 16, 1024),
                     make_smem_desc_sw128(bk + k * 32, 16, 1024), idesc_s, k > 0);
      umma_commit(bar_s);
      mbar_wait(bar_p, 0);
      tcgen05_fence_after();
      const int ksteps = args.keys_n / 16;
      for (int j = 0; j < ksteps; ++j)
        umma_bf16_ts(tmem_base + O_COL, tmem_base + j * 8, make_smem_desc_sw128(v0 + j * 2048, 8192, 1024),
                     idesc_o, j > 0);
      umma_commit(bar_o);
    }
  } else {
    const int q = warp & 3;
    const int hf = (warp - 1) >> 2;
    const int r = q * 32 + lane;
    const int nchunks = args.keys_n / 16;
    const int c_begin = hf == 0 ? 0 : (nchunks + 1) / 2;
    const int c_end = hf == 0 ? (nchunks + 1) / 2 : nchunks;
    int lo, hi;
    bool row_valid;
    key_range(args, NT, NT == 1 ? r : t * 128 + r, lo, hi, row_valid);
    if (!row_valid) hi = lo;  // no valid keys: the row is all padding
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    mbar_wait(bar_s, 0);
    tcgen05_fence_after();
    // ---- pass 1: row max of the raw scores over this thread's column range
    float mx = -INFINITY;
    for (int c = c_begin; c < c_end; ++c) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(trow + c * 16, v);
      tmem_ld_wait();
      const int col0 = c * 16;
      if (col0 >= lo && col0 + 16 <= hi) {
#pragma unroll
        for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col0 + j >= lo && col0 + j < hi) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
    }
    xchg[hf * 128 + r] = mx;
    named_bar_sync(1, 256);
    mx = fmaxf(mx, xchg[(hf ^ 1) * 128 + r]);
    const float m2 = mx == -INFINITY ? 0.f : mx * args.scale_log2;
    // ---- pass 2: p = 2^(s*c - m), packed to bf16 in registers (P aliases S, so no TMEM write yet)
    uint32_t pk[8][8];
    float sum = 0.f;
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) {
      const int c = c_begin + ci;
      if (c < c_end) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(trow + c * 16, v);
        tmem_ld_wait();
        const int col0 = c * 16;
        const bool full = col0 >= lo && col0 + 16 <= hi;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float x0 = fmaf(__uint_as_float(v[2 * j]), args.scale_log2, -m2);
          float x1 = fmaf(__uint_as_float(v[2 * j + 1]), args.scale_log2, -m2);
          if (!full) {
            if (!(col0 + 2 * j >= lo && col0 + 2 * j < hi)) x0 = -INFINITY;
            if (!(col0 + 2 * j + 1 >= lo && col0 + 2 * j + 1 < hi)) x1 = -INFINITY;
          }
          const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
          sum += p0 + p1;
          pk[ci][j] = pack_bf16x2(p0, p1);
        }
      }
    }
    xchg[256 + hf * 128 + r] = sum;
    tcgen05_fence_before();
    named_bar_sync(1, 256);  // every thread has finished reading S: P may now overwrite it
    tcgen05_fence_after();
    sum += xchg[256 + (hf ^ 1) * 128 + r];
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) {
      const int c = c_begin + ci;
      if (c < c_end) tmem_st_32x32b_x8(trow + c * 8, pk[ci]);
    }
    tmem_st_wait();
    tcgen05_fence_before();
    mbar_arrive(bar_p);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    if (hf == 0 && row_valid) {
      const int rr = NT == 1 ? r : t * 128 + r;
      const int b = b0 + (NT == 1 ? rr / args.N : 0);
      const int n = NT == 1 ? rr % args.N : rr;
      if (b < args.B) args.lse2[(static_cast<long long>(b) * args.H + head) * args.N + n] = m2 + log2f(sum);
    }
    // ---- epilogue: O / rowsum -> bf16 -> swizzled staging (the dead Q tile) -> TMA store
    mbar_wait(bar_o, 0);
    tcgen05_fence_after();
    uint32_t v[32];
    tmem_ld_32x32b_x32(trow + O_COL + hf * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 o;
      o.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]) * inv, __uint_as_float(v[8 * j + 1]) * inv);
      o.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv);
      o.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv);
      o.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv);
      *reinterpret_cast<uint4*>(sQ + sw128_offset(r, hf * 4 + j)) = o;
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 256);
    if (threadIdx.x == 32) {
      tma_store_3d(&tmO, sQ, head * 64, NT == 1 ? 0 : t * 128, b0);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// backward: one (sequence | packed group, head) per CTA
// ------------------------------------------------------------------------------------------------
// smem: Q (NT) | dO (NT) | K (NT) | V (NT) | P (2 chunks) | dS (2 chunks) | staging | barriers | row consts
// TMEM: S [0,128) | dP [128,256) | dQ_t [256+64t) | dK [384,448) | dV [448,512)
// warps: 0 = control (TMA + MMA issue); 4..19 = math/epilogue, four threads per query row (32 key
// columns each). Loops over (key tile u, query tile t) are fully unrolled so per-tile row constants
// stay in registers.
constexpr int BWD_MATH_THREADS = 512;
constexpr int BWD_THREADS = 128 + BWD_MATH_THREADS;

template <int NT>
__global__ void __launch_bounds__(BWD_THREADS, 1)
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
  uint64_t* bar_sdp_free = bars + 2;  // math threads done reading S/dP
  uint64_t* bar_pds = bars + 3;       // P and dS written to smem
  uint64_t* bar_mma = bars + 4;       // dQ/dK/dV MMAs of the pair finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* rowc = reinterpret_cast<float*>(bars + 8);  // [NT][128][2]: delta, lse2

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
    mbar_init(bar_sdp_free, BWD_MATH_THREADS);
    mbar_init(bar_pds, BWD_MATH_THREADS);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  if (NT == 1) {
    // rows never written by the TMA boxes: zero them so 0 x garbage cannot become NaN
    const int first = args.rows * 128, last = 128 * 128;
    for (int i = first + threadIdx.x * 16; i < last; i += BWD_THREADS * 16) {
      *reinterpret_cast<uint4*>(sQ + i) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sdO + i) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sK + i) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sV + i) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  // per-row constants: delta = sum_d dO*O and lse2; invalid rows get lse2 = +inf so that P = 0
  if (warp >= 4 && threadIdx.x - 128 < NT * 128) {
    const int idx = threadIdx.x - 128;
    const int rr = NT == 1 ? idx : idx;  // row within the sequence / packed group
    int lo_, hi_;
    bool row_valid;
    key_range(args, NT, rr, lo_, hi_, row_valid);
    const int b = b0 + (NT == 1 ? rr / args.N : 0);
    const int n = NT == 1 ? rr % args.N : rr;
    float delta = 0.f, l2 = INFINITY;
    if (row_valid && b < args.B) {
      const long long off = ((static_cast<long long>(b) * args.N + n) * args.H + head) * 64;
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
      l2 = args.lse2[(static_cast<long long>(b) * args.H + head) * args.N + n];
    }
    rowc[idx * 2] = delta;
    rowc[idx * 2 + 1] = l2;
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
    const int q = warp & 3;           // TMEM lane quarter
    const int qc = (warp - 4) >> 2;   // which 32 key columns of the 128-wide key tile
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const bool leader = threadIdx.x == 128;

    float delta[NT], lse2[NT];
    int lo[NT], hi[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      bool row_valid;
      key_range(args, NT, NT == 1 ? r : t * 128 + r, lo[t], hi[t], row_valid);
      delta[t] = rowc[(t * 128 + r) * 2];
      lse2[t] = rowc[(t * 128 + r) * 2 + 1];
    }

    // 128x64 fp32 accumulator at TMEM column `tcol` -> bf16 -> staging -> TMA store at (col, row0, b0)
    auto store_tile = [&](uint32_t tcol, int gcol, int row0) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(tcol + lane_off + qc * 16, v);
      tmem_ld_wait();
      if (leader) tma_store_wait_read<0>();
      named_bar_sync(1, BWD_MATH_THREADS);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint4 pk;
        pk.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
        pk.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
        pk.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
        pk.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
        *reinterpret_cast<uint4*>(stg + sw128_offset(r, qc * 2 + j)) = pk;
      }
      fence_proxy_async_smem();
      named_bar_sync(1, BWD_MATH_THREADS);
      if (leader) {
        tma_store_3d(&tmDQKV, stg, gcol, row0, b0);
        tma_store_commit();
      }
    };

#pragma unroll
    for (int u = 0; u < NT; ++u) {
      const int ku = min(128, args.keys_n - u * 128);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        constexpr int kPairsPerU = NT;
        const int pair = u * kPairsPerU + t;
        mbar_wait(bar_sdp, pair & 1);
        tcgen05_fence_after();
        const int col0 = qc * 32;             // first key column (within the tile) of this thread
        const bool active = col0 < ku;
        uint32_t pp[16], dd[16];              // 32 columns of P and dS, packed bf16 pairs
        if (active) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t sv[16], dv[16];
            tmem_ld_32x32b_x16(T_S + lane_off + col0 + h * 16, sv);
            tmem_ld_32x32b_x16(T_DP + lane_off + col0 + h * 16, dv);
            tmem_ld_wait();
            const int gcol = u * 128 + col0 + h * 16;
            const bool full = gcol >= lo[t] && gcol + 16 <= hi[t];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * e]), args.scale_log2, -lse2[t]));
              float p1 = ex2_approx(fmaf(__uint_as_float(sv[2 * e + 1]), args.scale_log2, -lse2[t]));
              if (!full) {
                if (!(gcol + 2 * e >= lo[t] && gcol + 2 * e < hi[t])) p0 = 0.f;
                if (!(gcol + 2 * e + 1 >= lo[t] && gcol + 2 * e + 1 < hi[t])) p1 = 0.f;
              }
              const float d0 = p0 * (__uint_as_float(dv[2 * e]) - delta[t]) * args.scale;
              const float d1 = p1 * (__uint_as_float(dv[2 * e + 1]) - delta[t]) * args.scale;
              pp[h * 8 + e] = pack_bf16x2(p0, p1);
              dd[h * 8 + e] = pack_bf16x2(d0, d1);
            }
          }
        }
        tcgen05_fence_before();
        mbar_arrive(bar_sdp_free);
        if (pair > 0) mbar_wait(bar_mma, (pair - 1) & 1);  // previous MMAs done with sP / sdS
        if (active) {
          uint8_t* pc = sP + (qc >> 1) * TILE_BYTES;
          uint8_t* dc = sdS + (qc >> 1) * TILE_BYTES;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t off = sw128_offset(r, (qc & 1) * 4 + j);
            *reinterpret_cast<uint4*>(pc + off) = make_uint4(pp[4 * j], pp[4 * j + 1], pp[4 * j + 2], pp[4 * j + 3]);
            *reinterpret_cast<uint4*>(dc + off) = make_uint4(dd[4 * j], dd[4 * j + 1], dd[4 * j + 2], dd[4 * j + 3]);
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(bar_pds);

        if (t == NT - 1) {
          // dK_u and dV_u are complete once this pair's MMAs retire
          mbar_wait(bar_mma, pair & 1);
          tcgen05_fence_after();
          store_tile(T_DK, (args.H + head) * 64, NT == 1 ? 0 : u * 128);
          store_tile(T_DV, (2 * args.H + head) * 64, NT == 1 ? 0 : u * 128);
          // the dK/dV accumulators are re-used by the next key tile: order these reads before its MMAs
          tcgen05_fence_before();
        }
      }
    }
    // dQ tiles (complete after the last pair; bar_mma already waited on above)
#pragma unroll
    for (int t = 0; t < NT; ++t) store_tile(T_DQ + t * 64, head * 64, NT == 1 ? 0 : t * 128);
    if (leader) tma_store_wait_all<0>();
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
  if (nt == 1) {
    const int smem = 3 * TILE_BYTES + 2048 + 1024 + 1024;
    static bool cfg = false;
    if (!cfg) {
      B200SSL_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cfg = true;
    }
    attention_fwd_kernel<1><<<groups * H, FWD_THREADS, smem, stream>>>(tq, to, a);
  } else {
    const int smem = 5 * TILE_BYTES + 2048 + 1024 + 1024;
    static bool cfg = false;
    if (!cfg) {
      B200SSL_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cfg = true;
    }
    attention_fwd_kernel<2><<<groups * H * 2, FWD_THREADS, smem, stream>>>(tq, to, a);
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
    const int smem = 9 * TILE_BYTES + 1024 + 2048 + 1024;
    static bool cfg = false;
    if (!cfg) {
      B200SSL_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cfg = true;
    }
    attention_bwd_kernel<1><<<grid, BWD_THREADS, smem, stream>>>(tq, tdo, tdq, a);
  } else {
    const int smem = 13 * TILE_BYTES + 1024 + 2048 + 1024;
    static bool cfg = false;
    if (!cfg) {
      B200SSL_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cfg = true;
    }
    attention_bwd_kernel<2><<<grid, BWD_THREADS, smem, stream>>>(tq, tdo, tdq, a);
  }
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}
