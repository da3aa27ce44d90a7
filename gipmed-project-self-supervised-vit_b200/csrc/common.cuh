// Shared device/host helpers for the b200ssl kernels (sm_100a only).
//
// Thin inline-PTX wrappers over the Blackwell primitives the hot path uses:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and
// the shared-memory + instruction descriptors that tcgen05.mma consumes.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200ssl {

// ----------------------------------------------------------------------------------------------
// error plumbing (C-ABI never throws; see include/b200ssl.h)
// ----------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);

#define B200SSL_CHECK(cond, code, ...)                 \
  do {                                                 \
    if (!(cond)) {                                     \
      ::b200ssl::set_last_error(__VA_ARGS__);          \
      return (code);                                   \
    }                                                  \
  } while (0)

#define B200SSL_CUDA(call)                                                              \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      ::b200ssl::set_last_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                                __FILE__, __LINE__);                                    \
      return -3;                                                                        \
    }                                                                                   \
  } while (0)

int sm_count();
bool pdl_enabled();  // programmatic dependent launch (b200ssl_set_pdl)

// Host: encode a bf16/fp32 tiled tensor map. dims/strides innermost-first, strides in BYTES for
// dims 1..rank-1 (dim 0 is contiguous). Returns 0 on success.
int make_tensor_map(CUtensorMap* out, const void* base, int elem_bytes, int rank,
                    const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                    int swizzle_bytes /* 0, 64 or 128 */);

#ifdef __CUDACC__

// ----------------------------------------------------------------------------------------------
// programmatic dependent launch: a kernel launched with the attribute may start (block scheduling, barrier /
// TMEM set-up, descriptor prefetch) while its predecessor in the stream drains; it must execute pdl_wait()
// before its first access to global memory, and pdl_launch_dependents() lets ITS successor start early.
// Both are no-ops for a kernel launched without the attribute.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ----------------------------------------------------------------------------------------------
// small utilities
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 %%rx;\n"
      ".reg .pred %%px;\n"
      "     elect.sync %%rx|%%px, %1;\n"
      "@%%px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exact-erf GELU (nn.GELU() default in the reference: VT.pyc@L89) and its derivative, sharing one
// erf evaluation: erf via Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, i.e. fp32-exact for a bf16
// result) with the two transcendentals on the MUFU pipe (rcp, ex2); exp(-x^2/2) is re-used for the pdf.
__device__ __forceinline__ void gelu_and_grad(float x, float& h, float& g) {
  float t, e;
  const float az = fabsf(x) * 0.70710678118654752f;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, az, 1.0f)));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(az * az * -1.4426950408889634f));
  const float erf_abs = fmaf(-poly, e, 1.0f);                 // erf(|x|/sqrt2)
  const float cdf = fmaf(copysignf(0.5f, x), erf_abs, 0.5f);  // Phi(x)
  h = x * cdf;
  g = fmaf(x * e, 0.39894228040143268f, cdf);                 // Phi(x) + x * phi(x)
}

// ----------------------------------------------------------------------------------------------
// explicit shared-space accesses (32-bit shared addresses; generic pointers into dynamic smem lose their
// address space after the alignment cast and would compile to generic LD/ST)
// ----------------------------------------------------------------------------------------------
// 16-byte global load that bypasses L1 (data another proxy / SM wrote moments ago: read it where it is coherent, in L2)
__device__ __forceinline__ float4 ld_global_cg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// ----------------------------------------------------------------------------------------------
// packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 — two fp32 lanes per issue slot). The epilogues
// are issue-bound, not FP32-pipe-bound, so halving the instruction count is what counts.
// ----------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack_f32x2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 splat_f32x2(float c) { return pack_f32x2(c, c); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// gelu_and_grad on a pair of pre-activations: same A&S 7.1.26 evaluation as above, all FMA-class
// operations packed; the four transcendentals (2 x rcp, 2 x ex2) stay scalar on the MUFU pipe.
__device__ __forceinline__ void gelu_and_grad2(float x0, float x1, f32x2& h, f32x2& g) {
  const float az0 = fabsf(x0) * 0.70710678118654752f, az1 = fabsf(x1) * 0.70710678118654752f;
  const f32x2 az = pack_f32x2(az0, az1);
  const f32x2 x = pack_f32x2(x0, x1);
  float d0, d1;
  unpack_f32x2(ffma2(splat_f32x2(0.3275911f), az, splat_f32x2(1.0f)), d0, d1);
  const f32x2 t = pack_f32x2(rcp_approx(d0), rcp_approx(d1));
  // npoly = -(a1 t + a2 t^2 + ... + a5 t^5)  (coefficients negated so that erf = 1 + npoly * e)
  f32x2 np = ffma2(t, splat_f32x2(-1.061405429f), splat_f32x2(1.453152027f));
  np = ffma2(np, t, splat_f32x2(-1.421413741f));
  np = ffma2(np, t, splat_f32x2(0.284496736f));
  np = ffma2(np, t, splat_f32x2(-0.254829592f));
  np = fmul2(np, t);
  float a0, a1;
  unpack_f32x2(fmul2(fmul2(az, splat_f32x2(-1.4426950408889634f)), az), a0, a1);
  const float e0 = ex2_approx(a0), e1 = ex2_approx(a1);
  const f32x2 e = pack_f32x2(e0, e1);
  const f32x2 erf_abs = ffma2(np, e, splat_f32x2(1.0f));                       // erf(|x|/sqrt2)
  const f32x2 half_s = pack_f32x2(copysignf(0.5f, x0), copysignf(0.5f, x1));
  const f32x2 cdf = ffma2(half_s, erf_abs, splat_f32x2(0.5f));                 // Phi(x)
  h = fmul2(x, cdf);
  g = ffma2(fmul2(x, e), splat_f32x2(0.39894228040143268f), cdf);              // Phi(x) + x phi(x)
}

// gelu only (inference / no-grad forward: the teacher), packed like gelu_and_grad2
__device__ __forceinline__ f32x2 gelu2(float x0, float x1) {
  const float az0 = fabsf(x0) * 0.70710678118654752f, az1 = fabsf(x1) * 0.70710678118654752f;
  const f32x2 az = pack_f32x2(az0, az1);
  float d0, d1;
  unpack_f32x2(ffma2(splat_f32x2(0.3275911f), az, splat_f32x2(1.0f)), d0, d1);
  const f32x2 t = pack_f32x2(rcp_approx(d0), rcp_approx(d1));
  f32x2 np = ffma2(t, splat_f32x2(-1.061405429f), splat_f32x2(1.453152027f));
  np = ffma2(np, t, splat_f32x2(-1.421413741f));
  np = ffma2(np, t, splat_f32x2(0.284496736f));
  np = ffma2(np, t, splat_f32x2(-0.254829592f));
  np = fmul2(np, t);
  float a0, a1;
  unpack_f32x2(fmul2(fmul2(az, splat_f32x2(-1.4426950408889634f)), az), a0, a1);
  const f32x2 e = pack_f32x2(ex2_approx(a0), ex2_approx(a1));
  const f32x2 erf_abs = ffma2(np, e, splat_f32x2(1.0f));
  const f32x2 half_s = pack_f32x2(copysignf(0.5f, x0), copysignf(0.5f, x1));
  return fmul2(pack_f32x2(x0, x1), ffma2(half_s, erf_abs, splat_f32x2(0.5f)));
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (never parks the warp): for a thread that polls SEVERAL barriers and acts on whichever is ready
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with an explicit suspend-time hint: the warp may be parked by the hardware for up to `ns` nanoseconds
// waiting for the phase, instead of coming back to re-issue the probe (spinning warps steal issue slots from the
// math warps: in the first profiles more than a third of all executed instructions were wait loops).
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must trap (launch failure the host can report), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity, 2000u)) {
    if (++spins > (1u << 22)) __trap();   // >= 4 M probes of up to 2 us each: seconds, i.e. a dead pipeline
  }
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];"
      ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5}], [%2];"
      ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// multicast load: the box lands at the same smem offset in every CTA of `mask`, and completes on the
// mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, "
      "{%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// ---- CTA-pair (cta_group::2) variants. In a 2-CTA cluster the shared-window address of CTA rank 1 has bit
// 24 set; clearing it names the same offset in the leader (rank 0) CTA.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t leader_smem_addr(const void* p) { return smem_u32(p) & kPeerBitMask; }
// TMA load issued by either CTA of the pair; completes (complete_tx) on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];"
      ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(leader_smem_addr(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the leader CTA's copy of `bar` (local arrive when executed by the leader itself)
// 1-D bulk copy global -> shared (TMA engine, no tensor map): bytes % 16 == 0, both addresses 16-byte aligned;
// completion is signalled on the mbarrier as transaction bytes
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(leader_smem_addr(bar)) : "memory");
}
// cluster-scope release/acquire pair for data one CTA of a pair writes to its own shared memory and the other
// CTA's MMA issuer must observe (the LayerNorm-fused GEMM's A panels)
__device__ __forceinline__ void mbar_arrive_leader_release(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
// same, source given as a 32-bit shared address
__device__ __forceinline__ void tma_store_2d_s(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, const void* src, int c0, int c1,
                                             int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// same box, but ADDED to global memory (element type taken from the tensor map): cp.reduce.async.bulk.tensor
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS)
               : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate, one CTA.
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A (M x 16 bf16) lives in TMEM, lane = row, two bf16 per column.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// same, but the arrive is multicast to the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// CTA-pair MMA: D[256 x N] split over the two CTAs' TMEM (128 rows each); each CTA supplies its 128 rows of A
// and its HALF of B (N/2 rows) at the same smem offsets. Issued by one thread of the leader CTA.
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (base_lane + t).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// thread t of the warp writes 8 consecutive 32-bit columns of lane (base_lane + t)
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor for tcgen05.mma, SWIZZLE_128B, descriptor version 1 (sm_100).
//   start address  bits [0,14)  (>>4)      leading byte offset bits [16,30) (>>4)
//   stride byte offset bits [32,46) (>>4)  version bits [46,48) = 1      layout bits [61,64) = 2
// K-major tile  : rows of 128 B (64 bf16 of K), 8-row groups SBO = 1024 B apart, LBO unused.
// MN-major tile : rows of 128 B (64 bf16 of M/N) per k, 8-k groups SBO = 1024 B apart,
//                 64-wide M/N chunks LBO bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// The same descriptor for an operand `byte_off` bytes behind a base whose address field `lo_base` =
// (smem address & 0x3FFFF) >> 4 the issuing thread computed once. With `byte_off` and the strides compile-time
// constants this is ONE integer add per descriptor: the MMA-issuing thread runs alone on the uniform datapath, where
// the mask / shift / or chain of make_smem_desc_sw128 costs ~8 dependent instructions per descriptor -- for MMAs
// of N = 64 (32 cycles of tensor work) the issue loop, not the tensor pipe, was the bound (attention.cu).
__device__ __forceinline__ uint64_t sw128_desc_at(uint32_t lo_base, uint32_t byte_off, uint32_t lbo_bytes,
                                                  uint32_t sbo_bytes) {
  const uint32_t lo = lo_base + ((byte_off >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16));
  const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32.
//   c_format [4,6)=1 (F32); a_format [7,10)=1 (BF16); b_format [10,13)=1; a_major bit 15;
//   b_major bit 16 (0 = K-major, 1 = MN-major); n_dim [17,23) = N>>3; m_dim [24,29) = M>>4.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) |
         (static_cast<uint32_t>(b_mn) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 16-byte chunk address inside a 128B-swizzled tile whose rows are 128 B (row r, chunk c in 0..7).
__device__ __forceinline__ uint32_t sw128_offset(int r, int c) {
  return static_cast<uint32_t>(r * 128 + ((c ^ (r & 7)) << 4));
}

// 16-byte chunk address inside a 64B-swizzled tile whose rows are 64 B (row r, chunk c in 0..3).
__device__ __forceinline__ uint32_t sw64_offset(int r, int c) {
  return static_cast<uint32_t>(r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
}

#endif  // __CUDACC__

}  // namespace b200ssl
