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
#include <cstdlib>

namespace b200ssl {

constexpr int MAXC = 12;       // max student crops handled by the unrolled loops
constexpr int LOSS_THREADS = 256;

struct OnlineLSE {
  float m, z;  // running max and sum exp(x - m)
};

// 2^(a - b) with the convention 2^(-inf - anything) = 0 (an empty partial state)
__device__ __forceinline__ float exp2_diff(float a, float b) { return a == -INFINITY ? 0.f : ex2_approx(a - b); }

// distributed shared memory: read a float of CTA `rank` of this cluster at the address `p` has in the caller's own smem
__device__ __forceinline__ float ld_cluster_f32(const float* p, uint32_t rank) {
  uint32_t ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(p)), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}

// Merge `n` partial online states for crop v. A part is [tm0, tz0, tm1, tz1, (sm, sz, u0, u1) x NC]; u0 / u1 of a part are
// relative to that part's OWN teacher maxima, so they are rescaled to the merged maxima M0 / M1. All in the log2 domain.
template <typename Get>
__device__ __forceinline__ void merge_parts(Get get, int n, int v, float& M0, float& Z0, float& M1, float& Z1, float& M,
                                            float& Z, float& U0, float& U1) {
  M0 = -INFINITY; M1 = -INFINITY; M = -INFINITY;
  for (int p = 0; p < n; ++p) {
    M0 = fmaxf(M0, get(p, 0));
    M1 = fmaxf(M1, get(p, 2));
    M = fmaxf(M, get(p, 4 + 4 * v));
  }
  Z0 = 0.f; Z1 = 0.f; Z = 0.f; U0 = 0.f; U1 = 0.f;
  for (int p = 0; p < n; ++p) {
    const float r0 = exp2_diff(get(p, 0), M0), r1 = exp2_diff(get(p, 2), M1);
    Z0 += get(p, 1) * r0;
    Z1 += get(p, 3) * r1;
    Z += get(p, 5 + 4 * v) * exp2_diff(get(p, 4 + 4 * v), M);
    U0 += get(p, 6 + 4 * v) * r0;
    U1 += get(p, 7 + 4 * v) * r1;
  }
}

// One CTA per sample -- or, as a launch attribute, a CLUSTER of 2..8 CTAs per sample: CTA `rank` streams its slice of the K
// prototypes of the sample's 2 + NC logit rows and the partial online states meet in CTA 0 through distributed shared
// memory (measured slower, see the launcher). Arithmetic in the log2 domain (one FMUL less per ex2).
template <int NC>
__global__ void __launch_bounds__(LOSS_THREADS)
dino_loss_fwd_kernel(const __nv_bfloat16* __restrict__ student, const __nv_bfloat16* __restrict__ teacher,
                     const float* __restrict__ center, float* __restrict__ loss, float* __restrict__ s_lse,
                     float* __restrict__ t_lse, int B, int K, float inv_ts, float inv_tt) {
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  pdl_wait();
  const uint32_t cs = cluster_nctarank(), rank = cluster_ctarank();
  const int b = blockIdx.x / cs;
  const int tid = threadIdx.x;
  const int n8 = K / 8, per = (n8 + static_cast<int>(cs) - 1) / static_cast<int>(cs);
  const int k_lo = static_cast<int>(rank) * per, k_hi = min(n8, k_lo + per);
  const float ts2 = inv_ts * kLog2e, tt2 = inv_tt * kLog2e;
  float tm[2] = {-INFINITY, -INFINITY}, tz[2] = {0.f, 0.f};
  float sm[NC], sz[NC], u0[NC], u1[NC];
#pragma unroll
  for (int v = 0; v < NC; ++v) { sm[v] = -INFINITY; sz[v] = 0.f; u0[v] = 0.f; u1[v] = 0.f; }

  for (int k8 = k_lo + tid; k8 < k_hi; k8 += LOSS_THREADS) {
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
        a[2 * e] = (f.x - cc[2 * e]) * tt2;
        a[2 * e + 1] = (f.y - cc[2 * e + 1]) * tt2;
        mx = fmaxf(mx, fmaxf(a[2 * e], a[2 * e + 1]));
      }
      if (mx > tm[iq]) {  // rescale the running sums to the new max (rare after the first chunks)
        const float r = ex2_approx(tm[iq] - mx);
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
        w[iq][e] = ex2_approx(a[e] - tm[iq]);
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
        x[2 * e] = f.x * ts2;
        x[2 * e + 1] = f.y * ts2;
        mx = fmaxf(mx, fmaxf(x[2 * e], x[2 * e + 1]));
      }
      if (mx > sm[v]) {
        sz[v] *= ex2_approx(sm[v] - mx);
        sm[v] = mx;
      }
      float zs = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        zs += ex2_approx(x[e] - sm[v]);
        d0 = fmaf(w[0][e], x[e], d0);
        d1 = fmaf(w[1][e], x[e], d1);
      }
      sz[v] += zs;
      u0[v] += d0;
      u1[v] += d1;
    }
  }

  // ---- reduction of the online states: warp shuffles, one smem round per CTA, one DSMEM round per cluster ----
  constexpr int NW = LOSS_THREADS / 32;
  constexpr int PART = 4 + 4 * NC;
  __shared__ float red[NW][PART];
  __shared__ float part[PART];
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int iq = 0; iq < 2; ++iq) {
      const float m2 = __shfl_xor_sync(0xffffffffu, tm[iq], o);
      const float z2 = __shfl_xor_sync(0xffffffffu, tz[iq], o);
      const float mn = fmaxf(tm[iq], m2);
      const float r1 = exp2_diff(tm[iq], mn), r2 = exp2_diff(m2, mn);
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
      const float mn = fmaxf(sm[v], m2);
      sz[v] = sz[v] * exp2_diff(sm[v], mn) + z2 * exp2_diff(m2, mn);
      sm[v] = mn;
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
  float M0, Z0, M1, Z1, M, Z, U0, U1;
  if (tid < NC) {  // thread v merges crop v over the warps of this CTA
    merge_parts([&](int p, int i) { return red[p][i]; }, NW, tid, M0, Z0, M1, Z1, M, Z, U0, U1);
    if (cs > 1) {
      if (tid == 0) { part[0] = M0; part[1] = Z0; part[2] = M1; part[3] = Z1; }
      part[4 + 4 * tid] = M; part[5 + 4 * tid] = Z; part[6 + 4 * tid] = U0; part[7 + 4 * tid] = U1;
    }
  }
  if (cs > 1) {
    cluster_sync_all();  // every CTA's `part` is written (release / acquire at cluster scope)
    if (rank == 0 && tid < NC)
      merge_parts([&](int p, int i) { return ld_cluster_f32(&part[i], static_cast<uint32_t>(p)); }, static_cast<int>(cs), tid,
                  M0, Z0, M1, Z1, M, Z, U0, U1);
  }
  if (rank == 0 && tid < 32) {
    float total = 0.f;
    if (tid < NC) {
      const float lse2 = M + log2f(Z);
      s_lse[static_cast<long long>(tid) * B + b] = lse2 * kLn2;
      // sum_k t_iq[k] * x[k] = U / Z_t (in log2 units, like lse2)
      if (tid != 0) total += lse2 - U0 / Z0;
      if (tid != 1) total += lse2 - U1 / Z1;
      if (tid == 0) {
        t_lse[b] = (M0 + log2f(Z0)) * kLn2;
        t_lse[B + b] = (M1 + log2f(Z1)) * kLn2;
      }
    }
    total = warp_sum(total);
    const int n_pairs = 2 * NC - 2;
    if (tid == 0) atomicAdd(loss, total * kLn2 / (static_cast<float>(B) * n_pairs));
  }
  if (cs > 1) cluster_sync_all();  // the other CTAs' shared memory must outlive CTA 0's reads
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
  // One CTA per sample by default. B200SSL_LOSS_CLUSTER=2|4|8 (developer A/B, read per call) spreads a sample over a
  // cluster; measured at B = 256, 12 crops, K = 65,536: 142 us (1) / 150 (2) / 162 (4) / 190 (8) -- the kernel is bound by
  // issue + MUFU (112 ex2 and ~900 instructions per thread and iteration), not by the 256-on-148 placement, and every
  // extra CTA repeats the 52-value butterfly reduction.
  const char* env_cs = getenv("B200SSL_LOSS_CLUSTER");
  const int max_cs = env_cs ? atoi(env_cs) : 1;
  int cs = 1;
  while (cs < max_cs && cs < 8 && (K / 8) / (cs * 2) >= 2 * LOSS_THREADS) cs *= 2;
#define CALL_F(NC)                                                                                          \
  B200SSL_CUDA(launch_pdl(dino_loss_fwd_kernel<NC>, dim3(B * cs), dim3(LOSS_THREADS), 0, s, cs,             \
                          static_cast<const __nv_bfloat16*>(student), static_cast<const __nv_bfloat16*>(teacher), \
                          center, loss, s_lse, t_lse, B, K, 1.f / student_temp, 1.f / teacher_temp))
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
