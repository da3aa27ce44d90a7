// K2: LayerNorm forward / backward (+ residual-gradient add), HBM-bound.
//
// Replaces ATen layer_norm at every norm site of the reference encoder (Block.norm1/norm2
// VT.pyc@L138,142,147,151; VisionTransformer.norm @L195,252; eps 1e-6 via the vit_* factories
// @L278,285,292). bf16 activations, fp32 statistics, 128-bit accesses, LPR lanes per row with
// CPL 16-byte chunks per lane (D = LPR * CPL * 8), warp-shuffle reductions inside the LPR group.
#include "common.cuh"

namespace b200ssl {

// load 8 consecutive elements (one "chunk") of a bf16 or fp32 row as floats
template <typename T>
__device__ __forceinline__ void load_chunk8(const T* row_ptr, int chunk, float (&out)[8]);
template <>
__device__ __forceinline__ void load_chunk8<__nv_bfloat16>(const __nv_bfloat16* row_ptr, int chunk, float (&out)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(row_ptr) + chunk);
  const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = unpack_bf16x2(uw[e]);
    out[2 * e] = f.x;
    out[2 * e + 1] = f.y;
  }
}
template <>
__device__ __forceinline__ void load_chunk8<float>(const float* row_ptr, int chunk, float (&out)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(row_ptr) + 2 * chunk);
  const float4 b = __ldg(reinterpret_cast<const float4*>(row_ptr) + 2 * chunk + 1);
  out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
  out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
}

template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// gamma / beta are read from shared memory rather than held in 48 registers: <= 64 registers per thread, so four CTAs
// (1024 threads) are resident per SM and twice as many row loads are in flight (measured: 46.3 -> 42.2 us at 100,864
// rows, 78.9 -> 72.8 us = 6.2 TB/s at 195,584 rows).
template <int LPR, int CPL, typename XT>
__global__ void __launch_bounds__(256, 4)
ln_fwd_kernel(const XT* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                long long rows, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int D = LPR * CPL * 8;
  constexpr int RPW = 32 / LPR;
  __shared__ __align__(16) float s_w[D];
  __shared__ __align__(16) float s_b[D];
  for (int i = threadIdx.x; i < D; i += 256) {
    s_w[i] = __ldg(w + i);
    s_b[i] = __ldg(b + i);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long base = warp_global * RPW; base < rows; base += nwarps * RPW) {
    const bool valid = base + lane / LPR < rows;
    const long long row = valid ? base + lane / LPR : rows - 1;
    const XT* px = x + row * D;
    float v[CPL * 8];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      float t8[8];
      load_chunk8<XT>(px, c * LPR + sub, t8);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[c * 8 + e] = t8[e];
        s += t8[e];
      }
    }
    const float mean = group_sum<LPR>(s) * (1.f / D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < CPL * 8; ++i) {
      const float d = v[i] - mean;
      sq += d * d;
    }
    const float rstd = rsqrtf(group_sum<LPR>(sq) * (1.f / D) + eps);
    if (!valid) continue;
    uint4* py = reinterpret_cast<uint4*>(y + row * D);
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int col = (c * LPR + sub) * 8;
      const float4 w0 = *reinterpret_cast<const float4*>(s_w + col), w1 = *reinterpret_cast<const float4*>(s_w + col + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(s_b + col), b1 = *reinterpret_cast<const float4*>(s_b + col + 4);
      const float gw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float gb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a0 = (v[c * 8 + 2 * e] - mean) * rstd * gw[2 * e] + gb[2 * e];
        const float a1 = (v[c * 8 + 2 * e + 1] - mean) * rstd * gw[2 * e + 1] + gb[2 * e + 1];
        o[e] = pack_bf16x2(a0, a1);
      }
      py[c * LPR + sub] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (sub == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) + dres, with g = dy * gamma;
// dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy   (fp32 atomics, one per column per CTA)
template <int LPR, int CPL, typename XT>
__global__ void __launch_bounds__(256, 2)
ln_bwd_kernel(const XT* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
              const float* __restrict__ w, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
              const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx, float* __restrict__ dw,
              float* __restrict__ db, long long rows) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int D = LPR * CPL * 8;
  constexpr int RPW = 32 / LPR;
  __shared__ float s_dw[D];
  __shared__ float s_db[D];
  for (int i = threadIdx.x; i < D; i += blockDim.x) { s_dw[i] = 0.f; s_db[i] = 0.f; }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;

  // dgamma / dbeta partials live in registers for the whole kernel; gamma is re-read per row (L1-resident) so
  // that two 256-thread blocks fit per SM (the kernel is latency bound at one)
  float adw[CPL * 8], adb[CPL * 8];
#pragma unroll
  for (int i = 0; i < CPL * 8; ++i) { adw[i] = 0.f; adb[i] = 0.f; }
  for (long long base = warp_global * RPW; base < rows; base += nwarps * RPW) {
    const bool valid = base + lane / LPR < rows;
    const long long row = valid ? base + lane / LPR : rows - 1;
    const float vmask = valid ? 1.f : 0.f;  // tail lanes re-read the last row but contribute nothing
    const XT* px = x + row * D;
    const uint4* pdy = reinterpret_cast<const uint4*>(dy + row * D);
    const float mean = mean_in[row], rstd = rstd_in[row];
    float xh[CPL * 8], g[CPL * 8];
    float s1 = 0.f, s2 = 0.f;
    // the residual-branch gradient is only needed at the end: fetch it now so its latency overlaps the
    // reductions instead of being exposed after them
    const uint4* pr = dres ? reinterpret_cast<const uint4*>(dres + row * D) : nullptr;
    uint4 r4[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) r4[c] = pr ? __ldg(pr + c * LPR + sub) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      float x8[8], gw[8];
      load_chunk8<XT>(px, c * LPR + sub, x8);
      load_chunk8<float>(w, c * LPR + sub, gw);
      const uint4 ud = __ldg(pdy + c * LPR + sub);
      const uint32_t dw4[4] = {ud.x, ud.y, ud.z, ud.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fx = make_float2(x8[2 * e], x8[2 * e + 1]);
        float2 fd = unpack_bf16x2(dw4[e]);
        fd.x *= vmask;
        fd.y *= vmask;
        const int i = c * 8 + 2 * e;
        xh[i] = (fx.x - mean) * rstd;
        xh[i + 1] = (fx.y - mean) * rstd;
        adw[i] += fd.x * xh[i];
        adw[i + 1] += fd.y * xh[i + 1];
        adb[i] += fd.x;
        adb[i + 1] += fd.y;
        g[i] = fd.x * gw[2 * e];
        g[i + 1] = fd.y * gw[2 * e + 1];
        s1 += g[i] + g[i + 1];
        s2 += g[i] * xh[i] + g[i + 1] * xh[i + 1];
      }
    }
    const float m1 = group_sum<LPR>(s1) * (1.f / D);
    const float m2 = group_sum<LPR>(s2) * (1.f / D);
    if (!valid) continue;
    uint4* pdx = reinterpret_cast<uint4*>(dx + row * D);
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const uint32_t rw[4] = {r4[c].x, r4[c].y, r4[c].z, r4[c].w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = c * 8 + 2 * e;
        const float2 fr = unpack_bf16x2(rw[e]);
        const float a0 = rstd * (g[i] - m1 - xh[i] * m2) + fr.x;
        const float a1 = rstd * (g[i + 1] - m1 - xh[i + 1] * m2) + fr.y;
        o[e] = pack_bf16x2(a0, a1);
      }
      pdx[c * LPR + sub] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  // fold the per-thread column partials: smem atomics per CTA, then one global atomic per column
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int col = (c * LPR + sub) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(&s_dw[col + e], adw[c * 8 + e]);
      atomicAdd(&s_db[col + e], adb[c * 8 + e]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dw + i, s_dw[i]);
    atomicAdd(db + i, s_db[i]);
  }
}

// Backward with the rows staged through shared memory by the bulk-copy engine: a CTA works on tiles of 8 * RPW
// consecutive rows; while it reduces tile k, one thread has the x / dy / dres rows of tile k + 1 (48 KB at D = 384)
// in flight into the other stage. The register-resident version above is latency bound: nothing of the next rows is
// requested before the current ones are finished, and 128 registers leave no room for a register prefetch.
template <int LPR, int CPL, typename XT>
struct LnBwdStaged {
  static constexpr int D = LPR * CPL * 8;
  static constexpr int RPW = 32 / LPR;
  static constexpr int TILE_ROWS = 8 * RPW;
  static constexpr int XB = TILE_ROWS * D * static_cast<int>(sizeof(XT));
  static constexpr int YB = TILE_ROWS * D * 2;
  static constexpr int STAGE = XB + 2 * YB;
  static constexpr int SMEM = 2 * STAGE + 2 * D * 4 + 16;
};

template <int LPR, int CPL, typename XT>
__global__ void __launch_bounds__(256, 2)
ln_bwd_staged_kernel(const XT* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                     const float* __restrict__ w, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                     const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx, float* __restrict__ dw,
                     float* __restrict__ db, long long rows) {
  pdl_launch_dependents();
  pdl_wait();
  using L = LnBwdStaged<LPR, CPL, XT>;
  constexpr int D = L::D, RPW = L::RPW, TILE_ROWS = L::TILE_ROWS;
  extern __shared__ __align__(128) uint8_t ln_smem[];
  float* s_dw = reinterpret_cast<float*>(ln_smem + 2 * L::STAGE);
  float* s_db = s_dw + D;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_db + D);
  for (int i = threadIdx.x; i < D; i += 256) { s_dw[i] = 0.f; s_db[i] = 0.f; }
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % LPR;
  const int row_in_tile = warp * RPW + lane / LPR;
  const long long ntiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
  auto issue = [&](long long tile, int stage) {
    const long long r0 = tile * TILE_ROWS;
    const int nr = static_cast<int>(rows - r0 < TILE_ROWS ? rows - r0 : TILE_ROWS);
    uint8_t* st = ln_smem + stage * L::STAGE;
    const uint32_t xb = nr * D * static_cast<uint32_t>(sizeof(XT)), yb = nr * D * 2u;
    mbar_expect_tx(&bars[stage], xb + yb + (dres ? yb : 0u));
    bulk_load_1d(st, x + r0 * D, xb, &bars[stage]);
    bulk_load_1d(st + L::XB, dy + r0 * D, yb, &bars[stage]);
    if (dres) bulk_load_1d(st + L::XB + L::YB, dres + r0 * D, yb, &bars[stage]);
  };

  float adw[CPL * 8], adb[CPL * 8];
#pragma unroll
  for (int i = 0; i < CPL * 8; ++i) { adw[i] = 0.f; adb[i] = 0.f; }
  if (threadIdx.x == 0 && blockIdx.x < ntiles) issue(blockIdx.x, 0);
  int k = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++k) {
    const int stage = k & 1;
    // stage ^ 1 was last read in iteration k - 1, which ended with a block barrier
    if (threadIdx.x == 0 && tile + gridDim.x < ntiles) issue(tile + gridDim.x, stage ^ 1);
    const long long row_raw = tile * TILE_ROWS + row_in_tile;
    const bool valid = row_raw < rows;
    const long long row = valid ? row_raw : rows - 1;
    const float mean = __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
    const uint8_t* st = ln_smem + stage * L::STAGE;
    const XT* px = reinterpret_cast<const XT*>(st) + row_in_tile * D;
    const uint4* pdy = reinterpret_cast<const uint4*>(st + L::XB) + row_in_tile * (D / 8);
    const uint4* pr = reinterpret_cast<const uint4*>(st + L::XB + L::YB) + row_in_tile * (D / 8);
    mbar_wait(&bars[stage], (k >> 1) & 1);
    float xh[CPL * 8], g[CPL * 8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      float x8[8], gw[8];
      if constexpr (sizeof(XT) == 4) {
        const float4 a = reinterpret_cast<const float4*>(px)[2 * (c * LPR + sub)];
        const float4 b = reinterpret_cast<const float4*>(px)[2 * (c * LPR + sub) + 1];
        x8[0] = a.x; x8[1] = a.y; x8[2] = a.z; x8[3] = a.w; x8[4] = b.x; x8[5] = b.y; x8[6] = b.z; x8[7] = b.w;
      } else {
        const uint4 u = reinterpret_cast<const uint4*>(px)[c * LPR + sub];
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16x2(uu[e]);
          x8[2 * e] = f.x;
          x8[2 * e + 1] = f.y;
        }
      }
      load_chunk8<float>(w, c * LPR + sub, gw);
      uint4 ud = pdy[c * LPR + sub];
      if (!valid) {  // rows past the end: the stage holds stale bytes there
        ud = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int e = 0; e < 8; ++e) x8[e] = mean;
      }
      const uint32_t dw4[4] = {ud.x, ud.y, ud.z, ud.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fd = unpack_bf16x2(dw4[e]);
        const int i = c * 8 + 2 * e;
        xh[i] = (x8[2 * e] - mean) * rstd;
        xh[i + 1] = (x8[2 * e + 1] - mean) * rstd;
        adw[i] += fd.x * xh[i];
        adw[i + 1] += fd.y * xh[i + 1];
        adb[i] += fd.x;
        adb[i + 1] += fd.y;
        g[i] = fd.x * gw[2 * e];
        g[i + 1] = fd.y * gw[2 * e + 1];
        s1 += g[i] + g[i + 1];
        s2 += g[i] * xh[i] + g[i + 1] * xh[i + 1];
      }
    }
    const float m1 = group_sum<LPR>(s1) * (1.f / D);
    const float m2 = group_sum<LPR>(s2) * (1.f / D);
    if (valid) {
      uint4* pdx = reinterpret_cast<uint4*>(dx + row * D);
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const uint4 r4 = dres ? pr[c * LPR + sub] : make_uint4(0, 0, 0, 0);
        const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = c * 8 + 2 * e;
          const float2 fr = unpack_bf16x2(rw[e]);
          const float a0 = rstd * (g[i] - m1 - xh[i] * m2) + fr.x;
          const float a1 = rstd * (g[i + 1] - m1 - xh[i + 1] * m2) + fr.y;
          o[e] = pack_bf16x2(a0, a1);
        }
        pdx[c * LPR + sub] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    __syncthreads();  // every read of this stage is done before the next iteration refills it
  }
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int col = (c * LPR + sub) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(&s_dw[col + e], adw[c * 8 + e]);
      atomicAdd(&s_db[col + e], adb[c * 8 + e]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += 256) {
    atomicAdd(dw + i, s_dw[i]);
    atomicAdd(db + i, s_db[i]);
  }
}

static int g_ln_bwd_staged = 1;

template <int LPR, int CPL>
static int launch_ln_fwd(const void* x, int x_f32, const float* w, const float* b, void* y, float* mean, float* rstd,
                         long long rows, float eps, cudaStream_t s) {
  constexpr int RPW = 32 / LPR;
  const long long warps_needed = (rows + RPW - 1) / RPW;
  long long blocks = (warps_needed + 7) / 8;
  const long long cap = static_cast<long long>(sm_count()) * 4;  // exactly the four resident blocks per SM: one wave (41.9 -> 39.8 us at 100,864 rows vs 8 per SM)
  if (blocks > cap) blocks = cap;
  if (x_f32)
    B200SSL_CUDA(launch_pdl(ln_fwd_kernel<LPR, CPL, float>, dim3(static_cast<int>(blocks)), dim3(256), 0, s, 1,
                            static_cast<const float*>(x), w, b, static_cast<__nv_bfloat16*>(y), mean, rstd, rows, eps));
  else
    B200SSL_CUDA(launch_pdl(ln_fwd_kernel<LPR, CPL, __nv_bfloat16>, dim3(static_cast<int>(blocks)), dim3(256), 0, s, 1,
                            static_cast<const __nv_bfloat16*>(x), w, b, static_cast<__nv_bfloat16*>(y), mean, rstd, rows,
                            eps));
  return 0;
}
template <int LPR, int CPL>
static int launch_ln_bwd(const void* x, int x_f32, const void* dy, const float* w, const float* mean, const float* rstd,
                         const void* dres, void* dx, float* dw, float* db, long long rows, cudaStream_t s) {
  constexpr int RPW = 32 / LPR;
  const long long warps_needed = (rows + RPW - 1) / RPW;
  long long blocks = (warps_needed + 7) / 8;
  const long long cap = static_cast<long long>(sm_count()) * 2;  // two resident blocks per SM, one wave
  if (blocks > cap) blocks = cap;
  if constexpr (LnBwdStaged<LPR, CPL, float>::SMEM <= 110 * 1024) {   // two CTAs per SM must fit
    if (g_ln_bwd_staged) {
      static bool attr_done = false;
      if (!attr_done) {
        B200SSL_CUDA(cudaFuncSetAttribute(ln_bwd_staged_kernel<LPR, CPL, float>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, LnBwdStaged<LPR, CPL, float>::SMEM));
        B200SSL_CUDA(cudaFuncSetAttribute(ln_bwd_staged_kernel<LPR, CPL, __nv_bfloat16>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          LnBwdStaged<LPR, CPL, __nv_bfloat16>::SMEM));
        attr_done = true;
      }
      const long long tiles = (rows + 8 * RPW - 1) / (8 * RPW);
      const int grid = static_cast<int>(tiles < cap ? tiles : cap);
      if (x_f32)
        B200SSL_CUDA(launch_pdl(ln_bwd_staged_kernel<LPR, CPL, float>, dim3(grid), dim3(256),
                                LnBwdStaged<LPR, CPL, float>::SMEM, s, 1, static_cast<const float*>(x),
                                static_cast<const __nv_bfloat16*>(dy), w, mean, rstd,
                                static_cast<const __nv_bfloat16*>(dres), static_cast<__nv_bfloat16*>(dx), dw, db, rows));
      else
        B200SSL_CUDA(launch_pdl(ln_bwd_staged_kernel<LPR, CPL, __nv_bfloat16>, dim3(grid), dim3(256),
                                LnBwdStaged<LPR, CPL, __nv_bfloat16>::SMEM, s, 1, static_cast<const __nv_bfloat16*>(x),
                                static_cast<const __nv_bfloat16*>(dy), w, mean, rstd,
                                static_cast<const __nv_bfloat16*>(dres), static_cast<__nv_bfloat16*>(dx), dw, db, rows));
      return 0;
    }
  }
  if (x_f32)
    B200SSL_CUDA(launch_pdl(ln_bwd_kernel<LPR, CPL, float>, dim3(static_cast<int>(blocks)), dim3(256), 0, s, 1,
                            static_cast<const float*>(x), static_cast<const __nv_bfloat16*>(dy), w, mean, rstd,
                            static_cast<const __nv_bfloat16*>(dres), static_cast<__nv_bfloat16*>(dx), dw, db, rows));
  else
    B200SSL_CUDA(launch_pdl(ln_bwd_kernel<LPR, CPL, __nv_bfloat16>, dim3(static_cast<int>(blocks)), dim3(256), 0, s, 1,
                            static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), w, mean, rstd,
                            static_cast<const __nv_bfloat16*>(dres), static_cast<__nv_bfloat16*>(dx), dw, db, rows));
  return 0;
}

}  // namespace b200ssl

using namespace b200ssl;

#define LN_DISPATCH(D, CALL)                                       \
  switch (D) {                                                     \
    case 192: return CALL(8, 3);                                   \
    case 256: return CALL(32, 1);                                  \
    case 384: return CALL(16, 3);                                  \
    case 512: return CALL(32, 2);                                  \
    case 768: return CALL(32, 3);                                  \
    case 1024: return CALL(32, 4);                                 \
    default:                                                       \
      set_last_error("layernorm: width %d unsupported (192/256/384/512/768/1024)", D); \
      return -2;                                                   \
  }

// x is bf16 (x_f32 == 0) or fp32 (the residual stream, x_f32 == 1); y is bf16.
extern "C" int b200ssl_layernorm_fwd(const void* x, int x_f32, const float* w, const float* b, void* y, float* mean,
                                     float* rstd, long long rows, int D, float eps, void* stream) {
  B200SSL_CHECK(rows > 0, -2, "layernorm: no rows");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define CALL_FWD(L, C) launch_ln_fwd<L, C>(x, x_f32, w, b, y, mean, rstd, rows, eps, s)
  LN_DISPATCH(D, CALL_FWD)
#undef CALL_FWD
}

// dw / db are ACCUMULATED into (fp32); the caller zeroes them when starting a fresh gradient.
extern "C" int b200ssl_layernorm_bwd(const void* x, int x_f32, const void* dy, const float* w, const float* mean,
                                     const float* rstd, const void* dres, void* dx, float* dw, float* db,
                                     long long rows, int D, void* stream) {
  B200SSL_CHECK(rows > 0, -2, "layernorm: no rows");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define CALL_BWD(L, C) launch_ln_bwd<L, C>(x, x_f32, dy, w, mean, rstd, dres, dx, dw, db, rows, s)
  LN_DISPATCH(D, CALL_BWD)
#undef CALL_BWD
}

// developer A/B switch: 0 = register-resident LayerNorm backward, 1 (default) = rows staged by bulk copies
extern "C" int b200ssl_set_ln_bwd_staged(int on) {
  b200ssl::g_ln_bwd_staged = on ? 1 : 0;
  return 0;
}
