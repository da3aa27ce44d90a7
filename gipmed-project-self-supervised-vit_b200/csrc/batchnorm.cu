// BatchNorm1d + GELU for DINOHead(use_bn=True) (VT.pyc@L304-305,309-310: a BatchNorm1d(hidden_dim) between every
// hidden Linear and its GELU). A non-default constructor option of the reference; small ([rows, 2048]) and HBM bound.
//
//   training: mean / biased var over the rows of this rank's batch; z = (x - mean) * invstd * w + b; h = gelu(z);
//             running_mean = (1 - mom) * running_mean + mom * mean, running_var likewise with the UNBIASED variance,
//             num_batches_tracked += 1  (torch.nn.BatchNorm1d semantics);
//   eval    : statistics = running_mean / running_var (the teacher).
//   backward: dz = dh * gelu'(z) (z recomputed from x), db = sum dz, dw = sum dz * xhat,
//             dx = w * invstd * (dz - mean(dz) - xhat * mean(dz * xhat))        (eval: dx = w * invstd * dz)
//
// One CTA owns 64 columns (32 lanes x 2 bf16) and walks all rows with 8 row lanes: 128-byte coalesced row
// segments, fp32 statistics, a shared-memory tree over the row lanes.
#include "common.cuh"

namespace b200ssl {

constexpr int BN_ROWLANES = 8;
constexpr int BN_THREADS = 32 * BN_ROWLANES;

__device__ __forceinline__ float2 bn_reduce2(float2 v, float2 (*red)[32], int rl, int cl) {
  red[rl][cl] = v;
  __syncthreads();
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < BN_ROWLANES; ++k) {
    s.x += red[k][cl].x;
    s.y += red[k][cl].y;
  }
  __syncthreads();
  return s;
}

__global__ void __launch_bounds__(BN_THREADS)
bn_gelu_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                   float* __restrict__ running_mean, float* __restrict__ running_var,
                   long long* __restrict__ num_batches, __nv_bfloat16* __restrict__ h, float* __restrict__ save_mean,
                   float* __restrict__ save_invstd, long long rows, int C, float momentum, float eps, int training,
                   int apply_gelu) {
  __shared__ float2 red[BN_ROWLANES][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + cl * 2;
  const bool live = c < C;
  float2 mean = make_float2(0.f, 0.f), invstd = make_float2(1.f, 1.f);
  if (training) {
    float2 s = make_float2(0.f, 0.f);
    if (live)
      for (long long r = rl; r < rows; r += BN_ROWLANES) {
        const float2 v = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(x + r * C + c)));
        s.x += v.x;
        s.y += v.y;
      }
    s = bn_reduce2(s, red, rl, cl);
    mean = make_float2(s.x / rows, s.y / rows);
    float2 q = make_float2(0.f, 0.f);
    if (live)
      for (long long r = rl; r < rows; r += BN_ROWLANES) {
        const float2 v = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(x + r * C + c)));
        q.x += (v.x - mean.x) * (v.x - mean.x);
        q.y += (v.y - mean.y) * (v.y - mean.y);
      }
    q = bn_reduce2(q, red, rl, cl);
    const float2 var = make_float2(q.x / rows, q.y / rows);
    invstd = make_float2(rsqrtf(var.x + eps), rsqrtf(var.y + eps));
    if (live && rl == 0) {
      const float unb = rows > 1 ? static_cast<float>(rows) / static_cast<float>(rows - 1) : 1.f;
      if (running_mean != nullptr) {
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean.x;
        running_mean[c + 1] = (1.f - momentum) * running_mean[c + 1] + momentum * mean.y;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * var.x * unb;
        running_var[c + 1] = (1.f - momentum) * running_var[c + 1] + momentum * var.y * unb;
      }
      save_mean[c] = mean.x; save_mean[c + 1] = mean.y;
      save_invstd[c] = invstd.x; save_invstd[c + 1] = invstd.y;
    }
    if (num_batches != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *num_batches += 1;
  } else if (live) {
    mean = make_float2(running_mean[c], running_mean[c + 1]);
    invstd = make_float2(rsqrtf(running_var[c] + eps), rsqrtf(running_var[c + 1] + eps));
    if (rl == 0 && save_mean != nullptr) {
      save_mean[c] = mean.x; save_mean[c + 1] = mean.y;
      save_invstd[c] = invstd.x; save_invstd[c + 1] = invstd.y;
    }
  }
  if (!live) return;
  const float2 g = make_float2(w ? w[c] : 1.f, w ? w[c + 1] : 1.f);
  const float2 bb = make_float2(b ? b[c] : 0.f, b ? b[c + 1] : 0.f);
  for (long long r = rl; r < rows; r += BN_ROWLANES) {
    const float2 v = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(x + r * C + c)));
    float z0 = (v.x - mean.x) * invstd.x * g.x + bb.x, z1 = (v.y - mean.y) * invstd.y * g.y + bb.y;
    if (apply_gelu) {
      float d0, d1;
      gelu_and_grad(z0, z0, d0);
      gelu_and_grad(z1, z1, d1);
    }
    *reinterpret_cast<uint32_t*>(h + r * C + c) = pack_bf16x2(z0, z1);
  }
}

__global__ void __launch_bounds__(BN_THREADS)
bn_gelu_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dh,
                   const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ save_mean,
                   const float* __restrict__ save_invstd, __nv_bfloat16* __restrict__ dx, float* __restrict__ dw,
                   float* __restrict__ db, long long rows, int C, int training, int apply_gelu) {
  __shared__ float2 red[BN_ROWLANES][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + cl * 2;
  const bool live = c < C;
  float2 mean = make_float2(0.f, 0.f), invstd = make_float2(1.f, 1.f), g = make_float2(1.f, 1.f),
         bb = make_float2(0.f, 0.f);
  if (live) {
    mean = make_float2(save_mean[c], save_mean[c + 1]);
    invstd = make_float2(save_invstd[c], save_invstd[c + 1]);
    if (w) g = make_float2(w[c], w[c + 1]);
    if (b) bb = make_float2(b[c], b[c + 1]);
  }
  auto dz_of = [&](long long r, float2& xh) {
    const float2 v = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(x + r * C + c)));
    float2 d = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(dh + r * C + c)));
    xh = make_float2((v.x - mean.x) * invstd.x, (v.y - mean.y) * invstd.y);
    if (apply_gelu) {
      float h0, h1, g0, g1;
      gelu_and_grad(xh.x * g.x + bb.x, h0, g0);
      gelu_and_grad(xh.y * g.y + bb.y, h1, g1);
      d.x *= g0;
      d.y *= g1;
    }
    return d;
  };
  float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
  if (live)
    for (long long r = rl; r < rows; r += BN_ROWLANES) {
      float2 xh;
      const float2 d = dz_of(r, xh);
      s1.x += d.x; s1.y += d.y;
      s2.x += d.x * xh.x; s2.y += d.y * xh.y;
    }
  s1 = bn_reduce2(s1, red, rl, cl);
  s2 = bn_reduce2(s2, red, rl, cl);
  if (!live) return;
  if (rl == 0) {
    if (db) { db[c] += s1.x; db[c + 1] += s1.y; }
    if (dw) { dw[c] += s2.x; dw[c + 1] += s2.y; }
  }
  const float inv_rows = 1.f / static_cast<float>(rows);
  for (long long r = rl; r < rows; r += BN_ROWLANES) {
    float2 xh;
    const float2 d = dz_of(r, xh);
    float o0, o1;
    if (training) {
      o0 = g.x * invstd.x * (d.x - s1.x * inv_rows - xh.x * s2.x * inv_rows);
      o1 = g.y * invstd.y * (d.y - s1.y * inv_rows - xh.y * s2.y * inv_rows);
    } else {
      o0 = g.x * invstd.x * d.x;
      o1 = g.y * invstd.y * d.y;
    }
    *reinterpret_cast<uint32_t*>(dx + r * C + c) = pack_bf16x2(o0, o1);
  }
}

}  // namespace b200ssl

using namespace b200ssl;

// x, h: bf16 [rows, C]; w, b (nullable), running_mean / running_var (nullable in training), save_mean / save_invstd:
// fp32 [C]; num_batches: int64 device scalar (nullable). apply_gelu = 1 fuses the exact-erf GELU that follows the norm.
extern "C" int b200ssl_bn_gelu_fwd(const void* x, const float* w, const float* b, float* running_mean,
                                   float* running_var, long long* num_batches, void* h, float* save_mean,
                                   float* save_invstd, long long rows, int C, float momentum, float eps, int training,
                                   int apply_gelu, void* stream) {
  B200SSL_CHECK(rows > 0 && C > 0 && C % 2 == 0, -2, "bn_gelu_fwd: rows=%lld C=%d (C must be even)", rows, C);
  B200SSL_CHECK(training || (running_mean != nullptr && running_var != nullptr), -2,
                "bn_gelu_fwd: eval mode needs running statistics");
  B200SSL_CHECK(!training || (save_mean != nullptr && save_invstd != nullptr), -2,
                "bn_gelu_fwd: training mode needs save_mean / save_invstd");
  bn_gelu_fwd_kernel<<<(C + 63) / 64, BN_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), w, b, running_mean, running_var, num_batches,
      static_cast<__nv_bfloat16*>(h), save_mean, save_invstd, rows, C, momentum, eps, training, apply_gelu);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

// dw / db (nullable) are ACCUMULATED into (gradient sinks); dx bf16 [rows, C].
extern "C" int b200ssl_bn_gelu_bwd(const void* x, const void* dh, const float* w, const float* b,
                                   const float* save_mean, const float* save_invstd, void* dx, float* dw, float* db,
                                   long long rows, int C, int training, int apply_gelu, void* stream) {
  B200SSL_CHECK(rows > 0 && C > 0 && C % 2 == 0, -2, "bn_gelu_bwd: rows=%lld C=%d (C must be even)", rows, C);
  bn_gelu_bwd_kernel<<<(C + 63) / 64, BN_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dh), w, b, save_mean, save_invstd,
      static_cast<__nv_bfloat16*>(dx), dw, db, rows, C, training, apply_gelu);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}
