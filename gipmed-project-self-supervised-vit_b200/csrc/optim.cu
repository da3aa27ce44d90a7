// K8: multi-tensor EMA (teacher update), fused AdamW(+clip, +EMA, +bf16 shadow) and the
// gradient sum-of-squares, each ONE launch over every parameter tensor. HBM-bound, 128-bit access.
//
// Reference sites: `model_ema.update(model)` (train.py:1081; timm ModelEmaV2 loops over ~150
// state-dict tensors with 2-3 tiny kernels each), `loss_scaler(loss, optimizer, clip_grad=...)` /
// `optimizer.step()` (train.py:1063-1078).
//
// A launch is described by a device-resident chunk table (int64 rows), built once per model by the
// host and cached: each row describes <= CHUNK elements of one tensor, so one CTA handles one row.
#include "common.cuh"

namespace b200ssl {

constexpr int OPT_THREADS = 256;

struct EmaRow {  // int64 x 4
  float* dst;
  const float* src;
  long long n;
  __nv_bfloat16* shadow;  // optional bf16 copy of dst the GEMMs read (nullptr: none)
};

// dst = m * dst + (1 - m) * src (+ refresh of dst's bf16 shadow in the same pass); the momentum is read from
// device memory so that a captured CUDA graph can be replayed with a new value every step (cosine schedule)
__global__ void __launch_bounds__(OPT_THREADS)
ema_kernel(const EmaRow* __restrict__ table, const float* __restrict__ momentum) {
  const EmaRow row = table[blockIdx.x];
  const float m = __ldg(momentum);
  const float om = 1.f - m;
  const bool vec = ((reinterpret_cast<uintptr_t>(row.dst) | reinterpret_cast<uintptr_t>(row.src) |
                     (reinterpret_cast<uintptr_t>(row.shadow) << 1)) & 15) == 0;
  if (vec) {
    const long long n4 = row.n / 4;
    float4* d = reinterpret_cast<float4*>(row.dst);
    const float4* s = reinterpret_cast<const float4*>(row.src);
    for (long long i = threadIdx.x; i < n4; i += OPT_THREADS) {
      float4 a = d[i];
      const float4 b = __ldg(s + i);
      a.x = m * a.x + om * b.x; a.y = m * a.y + om * b.y; a.z = m * a.z + om * b.z; a.w = m * a.w + om * b.w;
      d[i] = a;
      if (row.shadow) reinterpret_cast<uint2*>(row.shadow)[i] = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
    }
    for (long long i = n4 * 4 + threadIdx.x; i < row.n; i += OPT_THREADS) {
      const float v = m * row.dst[i] + om * row.src[i];
      row.dst[i] = v;
      if (row.shadow) row.shadow[i] = __float2bfloat16_rn(v);
    }
  } else {
    for (long long i = threadIdx.x; i < row.n; i += OPT_THREADS) {
      const float v = m * row.dst[i] + om * row.src[i];
      row.dst[i] = v;
      if (row.shadow) row.shadow[i] = __float2bfloat16_rn(v);
    }
  }
}

struct SumsqRow {  // int64 x 2
  const float* g;
  long long n;
};

__global__ void __launch_bounds__(OPT_THREADS)
sumsq_kernel(const SumsqRow* __restrict__ table, float* __restrict__ out) {
  const SumsqRow row = table[blockIdx.x];
  float acc = 0.f;
  if ((reinterpret_cast<uintptr_t>(row.g) & 15) == 0) {
    const long long n4 = row.n / 4;
    const float4* g = reinterpret_cast<const float4*>(row.g);
    for (long long i = threadIdx.x; i < n4; i += OPT_THREADS) {
      const float4 a = __ldg(g + i);
      acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    for (long long i = n4 * 4 + threadIdx.x; i < row.n; i += OPT_THREADS) acc += row.g[i] * row.g[i];
  } else {
    for (long long i = threadIdx.x; i < row.n; i += OPT_THREADS) acc += row.g[i] * row.g[i];
  }
  __shared__ float red[OPT_THREADS / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < OPT_THREADS / 32 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[1 + blockIdx.x] = v;  // per-chunk partial; summed in a fixed order below
  }
}

// out[0] = sum of the n partials in out[1..n], in an order that does not depend on scheduling: data-parallel
// replicas must compute bit-identical clipping coefficients or they drift apart
__global__ void __launch_bounds__(1024)
sumsq_final_kernel(float* __restrict__ out, int n) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) acc += out[1 + i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = red[threadIdx.x];
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = v;
  }
}

struct AdamRow {  // int64 x 8
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
  long long decay;        // 1: apply weight decay to this tensor
  float* ema;             // optional teacher copy (nullptr: skip)
  __nv_bfloat16* shadow;  // optional bf16 copy of the updated weights (nullptr: skip)
};

struct AdamHyper {
  float lr, beta1, beta2, eps, wd, max_norm, bc1, bc2, ema_m;
};
// per-step scalars live in device memory (graph replay): {lr, weight_decay, 1-beta1^t, 1-beta2^t, ema_m}
constexpr int ADAM_DEV_HYPER = 5;

__device__ __forceinline__ float adam_one(float p, float g, float& m, float& v, const AdamHyper& h, float clip,
                                          bool decay) {
  g *= clip;
  if (decay) p *= 1.f - h.lr * h.wd;
  m = h.beta1 * m + (1.f - h.beta1) * g;
  v = h.beta2 * v + (1.f - h.beta2) * g * g;
  const float denom = sqrtf(v) / sqrtf(h.bc2) + h.eps;
  return p - (h.lr / h.bc1) * (m / denom);
}

// torch.optim.AdamW semantics (decoupled decay first, bias-corrected moments); gradient clipping by
// global norm as torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (norm + 1e-6)).
__global__ void __launch_bounds__(OPT_THREADS)
adamw_kernel(const AdamRow* __restrict__ table, const float* __restrict__ gnorm_sq,
             const float* __restrict__ dev_hyper, AdamHyper h) {
  const AdamRow row = table[blockIdx.x];
  h.lr = __ldg(dev_hyper + 0);
  h.wd = __ldg(dev_hyper + 1);
  h.bc1 = __ldg(dev_hyper + 2);
  h.bc2 = __ldg(dev_hyper + 3);
  h.ema_m = __ldg(dev_hyper + 4);
  float clip = 1.f;
  if (gnorm_sq != nullptr && h.max_norm > 0.f) {
    const float norm = sqrtf(__ldg(gnorm_sq));
    clip = fminf(1.f, h.max_norm / (norm + 1e-6f));
  }
  const bool decay = row.decay != 0;
  const float om = 1.f - h.ema_m;
  const uintptr_t al = reinterpret_cast<uintptr_t>(row.p) | reinterpret_cast<uintptr_t>(row.g) |
                       reinterpret_cast<uintptr_t>(row.m) | reinterpret_cast<uintptr_t>(row.v) |
                       reinterpret_cast<uintptr_t>(row.ema) | (reinterpret_cast<uintptr_t>(row.shadow) << 1);
  const long long n4 = (al & 15) == 0 ? row.n / 4 : 0;
  for (long long i = threadIdx.x; i < n4; i += OPT_THREADS) {
    float4 p = reinterpret_cast<float4*>(row.p)[i];
    const float4 g = __ldg(reinterpret_cast<const float4*>(row.g) + i);
    float4 m = reinterpret_cast<float4*>(row.m)[i];
    float4 v = reinterpret_cast<float4*>(row.v)[i];
    p.x = adam_one(p.x, g.x, m.x, v.x, h, clip, decay);
    p.y = adam_one(p.y, g.y, m.y, v.y, h, clip, decay);
    p.z = adam_one(p.z, g.z, m.z, v.z, h, clip, decay);
    p.w = adam_one(p.w, g.w, m.w, v.w, h, clip, decay);
    reinterpret_cast<float4*>(row.p)[i] = p;
    reinterpret_cast<float4*>(row.m)[i] = m;
    reinterpret_cast<float4*>(row.v)[i] = v;
    if (row.ema) {
      float4 e = reinterpret_cast<float4*>(row.ema)[i];
      e.x = h.ema_m * e.x + om * p.x; e.y = h.ema_m * e.y + om * p.y;
      e.z = h.ema_m * e.z + om * p.z; e.w = h.ema_m * e.w + om * p.w;
      reinterpret_cast<float4*>(row.ema)[i] = e;
    }
    if (row.shadow)
      reinterpret_cast<uint2*>(row.shadow)[i] = make_uint2(pack_bf16x2(p.x, p.y), pack_bf16x2(p.z, p.w));
  }
  for (long long i = n4 * 4 + threadIdx.x; i < row.n; i += OPT_THREADS) {
    float m = row.m[i], v = row.v[i];
    const float p = adam_one(row.p[i], row.g[i], m, v, h, clip, decay);
    row.p[i] = p; row.m[i] = m; row.v[i] = v;
    if (row.ema) row.ema[i] = h.ema_m * row.ema[i] + om * p;
    if (row.shadow) row.shadow[i] = __float2bfloat16(p);
  }
}

}  // namespace b200ssl

using namespace b200ssl;

extern "C" int b200ssl_ema_multi_tensor(const void* table, int n_rows, const float* momentum_dev, void* stream) {
  if (n_rows <= 0) return 0;
  B200SSL_CHECK(momentum_dev != nullptr, -2, "ema: momentum must be a device pointer");
  ema_kernel<<<n_rows, OPT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const EmaRow*>(table),
                                                                           momentum_dev);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

// out is a float[1 + n_rows] workspace: out[0] receives sum over all rows of g^2 (deterministic order),
// out[1..] the per-row partials.
extern "C" int b200ssl_sumsq_multi_tensor(const void* table, int n_rows, float* out, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n_rows <= 0) {
    B200SSL_CUDA(cudaMemsetAsync(out, 0, sizeof(float), s));
    return 0;
  }
  sumsq_kernel<<<n_rows, OPT_THREADS, 0, s>>>(static_cast<const SumsqRow*>(table), out);
  sumsq_final_kernel<<<1, 1024, 0, s>>>(out, n_rows);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

// dev_hyper: device float[5] = {lr, weight_decay, 1-beta1^t, 1-beta2^t, ema_momentum} for THIS step.
extern "C" int b200ssl_adamw_multi_tensor(const void* table, int n_rows, const float* gnorm_sq,
                                          const float* dev_hyper, float beta1, float beta2, float eps,
                                          float max_norm, void* stream) {
  if (n_rows <= 0) return 0;
  B200SSL_CHECK(dev_hyper != nullptr, -2, "adamw: dev_hyper must be a device pointer");
  AdamHyper h{0.f, beta1, beta2, eps, 0.f, max_norm, 1.f, 1.f, 0.f};
  adamw_kernel<<<n_rows, OPT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const AdamRow*>(table),
                                                                             gnorm_sq, dev_hyper, h);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}
