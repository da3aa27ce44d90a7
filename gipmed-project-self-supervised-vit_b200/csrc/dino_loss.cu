// K7: DINO loss — teacher centering + sharpening, cross-entropy over (global teacher view,
// student crop) pairs, its gradient, and the running-centre update. HBM-bound.
//
// Not in the reference; occupies the `loss = loss_fn(output, target)` slot
// (train.py:1053). Semantics: Caron et al. 2021 Alg. 1, restated in oracle/dino.py.
//
//   t_iq = softmax((T_iq - c) / tau_t)            iq in {0,1}  (teacher, fp32 stats)
//   loss = mean_{iq, v != iq} mean_b  sum_k -t_iq[k] * log_softmax(S_v / tau_s)[k]
//        = mean_{pairs,b} ( lse_v - <t_iq, S_v / tau_s> )
//   dS_v = g / (tau_s * B * n_pairs) * ( n_v * softmax(S_v / tau_s) - sum_{iq != v} t_iq )
//
// Forward is ONE pass over the 2 + ncrops logit rows of a sample (online max / sum / dot with
// rescaling), so logits are read once forward and once backward; gradients are written once.
// Rows are crop-major: student row = v * B + b, teacher row = iq * B + b.
#include "common.cuh"

namespace b200ssl {

constexpr int MAXC = 12;       // max student crops handled by the unrolled loops
constexpr int LOSS_THREADS = 256;

struct OnlineLSE {
  float m, z;  // running max and sum exp(x - m)
};

// exp(a - b) with the convention exp(-inf - anything) = 0 (an empty partial state)
__device__ __forceinline__ float exp_diff(float a, float b) { return a == -INFINITY ? 0.f : __expf(a - b); }

__device__ __forceinline__ void lse_merge(float& m, float& z, float m2, float z2) {
  const float mn = fmaxf(m, m2);
  z = z * exp_diff(m, mn) + z2 * exp_diff(m2, mn);
  m = mn;
}

template <int NC>
__global__ void __launch_bounds__(LOSS_THREADS)
dino_loss_fwd_kernel(const __nv_bfloat16* __restrict__ student, const __nv_bfloat16* __restrict__ teacher,
                     const float* __restrict__ center, float* __restrict__ loss, float* __restrict__ s_lse,
                     float* __restrict__ t_lse, int B, int K, float inv_ts, float inv_tt) {
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  float tm[2] = {-INFINITY, -INFINITY}, tz[2] = {0.f, 0.f};
  float sm[NC], sz[NC], u0[NC], u1[NC];
#pragma unroll
  for (int v = 0; v < NC; ++v) { sm[v] = -INFINITY; sz[v] = 0.f; u0[v] = 0.f; u1[v] = 0.f; }

  for (int k8 = tid; k8 < K / 8; k8 += LOSS_THREADS) {
    float w[2][8];
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(center) + 2 * k8);
    const float4 c1 = __ldg(reinterpret_cast<const float4*>(center) + 2 * k8 + 1);
    const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
    for (int iq = 0; iq < 2; ++iq) {
      const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(teacher + (static_cast<long long>(iq) * B + b) * K) + k8);
      const uint32_t tw[4] = {t4.x, t4.y, t4.z, t4.w};
      float a[8];
      float mx = -INFINITY;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(tw[e]);
        a[2 * e] = (f.x - cc[2 * e]) * inv_tt;
        a[2 * e + 1] = (f.y - cc[2 * e + 1]) * inv_tt;
        mx = fmaxf(mx, fmaxf(a[2 * e], a[2 * e + 1]));
      }
      if (mx > tm[iq]) {  // rescale the running sums to the new max (rare after the first chunks)
        const float r = __expf(tm[iq] - mx);
        tz[iq] *= r;
        if (iq == 0) {
#pragma unroll
          for (int v = 0; v < NC; ++v) u0[v] *= r;
        } else {
#pragma unroll
          for (int v = 0; v < NC; ++v) u1[v] *= r;
        }
        tm[iq] = mx;
      }
      float zs = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        w[iq][e] = __expf(a[e] - tm[iq]);
        zs += w[iq][e];
      }
      tz[iq] += zs;
    }
#pragma unroll
    for (int v = 0; v < NC; ++v) {
      const uint4 s4 = __ldg(reinterpret_cast<const uint4*>(student + (static_cast<long long>(v) * B + b) * K) + k8);
      const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w};
      float x[8];
      float mx = -INFINITY;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(sw[e]);
        x[2 * e] = f.x * inv_ts;
        x[2 * e + 1] = f.y * inv_ts;
        mx = fmaxf(mx, fmaxf(x[2 * e], x[2 * e + 1]));
      }
      if (mx > sm[v]) {
        sz[v] *= __expf(sm[v] - mx);
        sm[v] = mx;
      }
      float zs = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        zs += __expf(x[e] - sm[v]);
        d0 += w[0][e] * x[e];
        d1 += w[1][e] * x[e];
      }
      sz[v] += zs;
      u0[v] += d0;
      u1[v] += d1;
    }
  }

  // ---- block reduction of the online states (warp shuffles, then one smem round) ----
  __shared__ float red[LOSS_THREADS / 32][4 + 4 * NC];
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int iq = 0; iq < 2; ++iq) {
      const float m2 = __shfl_xor_sync(0xffffffffu, tm[iq], o);
      const float z2 = __shfl_xor_sync(0xffffffffu, tz[iq], o);
      const float mn = fmaxf(tm[iq], m2);
      const float r1 = exp_diff(tm[iq], mn), r2 = exp_diff(m2, mn);
      tz[iq] = tz[iq] * r1 + z2 * r2;
#pragma unroll
      for (int v = 0; v < NC; ++v) {
        float& u = iq == 0 ? u0[v] : u1[v];
        const float uo = __shfl_xor_sync(0xffffffffu, u, o);
        u = u * r1 + uo * r2;
      }
      tm[iq] = mn;
    }
#pragma unroll
    for (int v = 0; v < NC; ++v) {
      const float m2 = __shfl_xor_sync(0xffffffffu, sm[v], o);
      const float z2 = __shfl_xor_sync(0xffffffffu, sz[v], o);
      lse_merge(sm[v], sz[v], m2, z2);
    }
  }
  if (lane == 0) {
    red[warp][0] = tm[0]; red[warp][1] = tz[0]; red[warp][2] = tm[1]; red[warp][3] = tz[1];
#pragma unroll
    for (int v = 0; v < NC; ++v) {
      red[warp][4 + 4 * v] = sm[v];
      red[warp][5 + 4 * v] = sz[v];
      red[warp][6 + 4 * v] = u0[v];
      red[warp][7 + 4 * v] = u1[v];
    }
  }
  __syncthreads();
  if (tid == 0) {
    constexpr int NW = LOSS_THREADS / 32;
    float M0 = -INFINITY, M1 = -INFINITY;
    for (int w = 0; w < NW; ++w) { M0 = fmaxf(M0, red[w][0]); M1 = fmaxf(M1, red[w][2]); }
    float Z0 = 0.f, Z1 = 0.f;
    for (int w = 0; w < NW; ++w) { Z0 += red[w][1] * exp_diff(red[w][0], M0); Z1 += red[w][3] * exp_diff(red[w][2], M1); }
    t_lse[b] = M0 + logf(Z0);
    t_lse[B + b] = M1 + logf(Z1);
    float total = 0.f;
    for (int v = 0; v < NC; ++v) {
      float M = -INFINITY;
      for (int w = 0; w < NW; ++w) M = fmaxf(M, red[w][4 + 4 * v]);
      float Z = 0.f, U0 = 0.f, U1 = 0.f;
      for (int w = 0; w < NW; ++w) {
        Z += red[w][5 + 4 * v] * exp_diff(red[w][4 + 4 * v], M);
        U0 += red[w][6 + 4 * v] * exp_diff(red[w][0], M0);
        U1 += red[w][7 + 4 * v] * exp_diff(red[w][2], M1);
      }
      const float lse = M + logf(Z);
      s_lse[static_cast<long long>(v) * B + b] = lse;
      if (v != 0) total += lse - U0 / Z0;
      if (v != 1) total += lse - U1 / Z1;
    }
    const int n_pairs = 2 * NC - 2;
    atomicAdd(loss, total / (static_cast<float>(B) * n_pairs));
  }
}

template <int NC>
__global__ void __launch_bounds__(256)
dino_loss_bwd_kernel(const __nv_bfloat16* __restrict__ student, const __nv_bfloat16* __restrict__ teacher,
                     const float* __restrict__ center, const float* __restrict__ s_lse,
                     const float* __restrict__ t_lse, const float* __restrict__ gout,
                     __nv_bfloat16* __restrict__ dstudent, int B, int K, float inv_ts, float inv_tt) {
  const int b = blockIdx.y;
  const int k8 = blockIdx.x * blockDim.x + threadIdx.x;
  if (k8 >= K / 8) return;
  const float coef = (gout ? __ldg(gout) : 1.f) * inv_ts / (static_cast<float>(B) * (2 * NC - 2));
  const float4 c0 = __ldg(reinterpret_cast<const float4*>(center) + 2 * k8);
  const float4 c1 = __ldg(reinterpret_cast<const float4*>(center) + 2 * k8 + 1);
  const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
  float t[2][8];
#pragma unroll
  for (int iq = 0; iq < 2; ++iq) {
    const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(teacher + (static_cast<long long>(iq) * B + b) * K) + k8);
    const uint32_t tw[4] = {t4.x, t4.y, t4.z, t4.w};
    const float l = __ldg(t_lse + iq * B + b);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_bf16x2(tw[e]);
      t[iq][2 * e] = __expf((f.x - cc[2 * e]) * inv_tt - l);
      t[iq][2 * e + 1] = __expf((f.y - cc[2 * e + 1]) * inv_tt - l);
    }
  }
#pragma unroll
  for (int v = 0; v < NC; ++v) {
    const long long row = static_cast<long long>(v) * B + b;
    const uint4 s4 = __ldg(reinterpret_cast<const uint4*>(student + row * K) + k8);
    const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w};
    const float l = __ldg(s_lse + row);
    const float nv = (v < 2) ? 1.f : 2.f;
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_bf16x2(sw[e]);
      const float p0 = __expf(f.x * inv_ts - l), p1 = __expf(f.y * inv_ts - l);
      const float tt0 = (v != 0 ? t[0][2 * e] : 0.f) + (v != 1 ? t[1][2 * e] : 0.f);
      const float tt1 = (v != 0 ? t[0][2 * e + 1] : 0.f) + (v != 1 ? t[1][2 * e + 1] : 0.f);
      o[e] = pack_bf16x2(coef * (nv * p0 - tt0), coef * (nv * p1 - tt1));
    }
    reinterpret_cast<uint4*>(dstudent + row * K)[k8] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// center = center * m + batch_sum * (1 - m) / total_rows   (batch_sum already all-reduced)
__global__ void center_update_kernel(float* __restrict__ center, const float* __restrict__ batch_sum, int K,
                                     float scale, float m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K) center[i] = center[i] * m + batch_sum[i] * scale;
}

}  // namespace b200ssl

using namespace b200ssl;

#define NC_DISPATCH(NC, CALL)            \
  switch (NC) {                          \
    case 2: CALL(2); break;              \
    case 3: CALL(3); break;              \
    case 4: CALL(4); break;              \
    case 5: CALL(5); break;              \
    case 6: CALL(6); break;              \
    case 7: CALL(7); break;              \
    case 8: CALL(8); break;              \
    case 9: CALL(9); break;              \
    case 10: CALL(10); break;            \
    case 11: CALL(11); break;            \
    case 12: CALL(12); break;            \
    default:                             \
      set_last_error("dino_loss: ncrops=%d unsupported (2..12)", NC); \
      return -2;                         \
  }

// loss[0] is overwritten with the mean loss; s_lse [ncrops*B] and t_lse [2*B] are saved for backward.
extern "C" int b200ssl_dino_loss_fwd(const void* student, const void* teacher, const float* center, float* loss,
                                     float* s_lse, float* t_lse, int B, int ncrops, int K, float student_temp,
                                     float teacher_temp, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  B200SSL_CHECK(B > 0 && K > 0 && K % 8 == 0, -2, "dino_loss: B=%d K=%d (K must be a multiple of 8)", B, K);
  B200SSL_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), s));
#define CALL_F(NC)                                                                                          \
  dino_loss_fwd_kernel<NC><<<B, LOSS_THREADS, 0, s>>>(static_cast<const __nv_bfloat16*>(student),           \
                                                      static_cast<const __nv_bfloat16*>(teacher), center,  \
                                                      loss, s_lse, t_lse, B, K, 1.f / student_temp,         \
                                                      1.f / teacher_temp)
  NC_DISPATCH(ncrops, CALL_F)
#undef CALL_F
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

// gout: device pointer to the upstream scalar gradient (nullptr == 1).
extern "C" int b200ssl_dino_loss_bwd(const void* student, const void* teacher, const float* center,
                                     const float* s_lse, const float* t_lse, const float* gout, void* dstudent,
                                     int B, int ncrops, int K, float student_temp, float teacher_temp,
                                     void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  B200SSL_CHECK(B > 0 && K > 0 && K % 8 == 0, -2, "dino_loss: B=%d K=%d (K must be a multiple of 8)", B, K);
  dim3 grid((K / 8 + 255) / 256, B);
#define CALL_B(NC)                                                                                          \
  dino_loss_bwd_kernel<NC><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(student),                 \
                                                static_cast<const __nv_bfloat16*>(teacher), center, s_lse,  \
                                                t_lse, gout, static_cast<__nv_bfloat16*>(dstudent), B, K,   \
                                                1.f / student_temp, 1.f / teacher_temp)
  NC_DISPATCH(ncrops, CALL_B)
#undef CALL_B
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200ssl_center_update(float* center, const float* batch_sum, int K, long long total_rows,
                                     float momentum, void* stream) {
  B200SSL_CHECK(total_rows > 0, -2, "center_update: total_rows must be positive");
  center_update_kernel<<<(K + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      center, batch_sum, K, (1.f - momentum) / static_cast<float>(total_rows), momentum);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}
