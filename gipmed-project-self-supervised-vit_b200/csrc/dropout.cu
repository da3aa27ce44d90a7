// Element dropout on the encoder path: nn.Dropout(drop_rate) of the reference's VisionTransformer -- pos_drop on the token
// stream (VT.pyc@L196,245), Attention.proj_drop on the attention branch (@L117,130) and Mlp.drop behind the activation and
// behind fc2 (@L96,101-104); train.py exposes the rate as --drop (train.py:283, :487). Default 0 in the reference, in which
// case none of this is launched.
//
// The mask is a pure function of (seed, site, element index): a counter-based generator, so backward (and an activation
// recompute) regenerates it instead of storing it. One splitmix64 word per group of FOUR consecutive elements, 16 random
// bits per element; an element is dropped when its 16 bits are below thr = round(p * 65536) and scaled by 1 / (1 - p)
// otherwise (torch semantics). `seed` is a DEVICE pointer to one 64-bit word (drawn on the device by the host module,
// so a captured CUDA graph sees a fresh seed on every replay); `site` names the dropout instance inside the model.
// oracle/dropout.py restates the generator so that a PyTorch oracle can replay the identical mask.
#include "common.cuh"

namespace b200ssl {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111C5ull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned long long drop_key(const unsigned long long* seed, unsigned site) {
  return __ldg(seed) ^ (static_cast<unsigned long long>(site) + 1ull) * 0xA0761D6478BD642Full;
}
// the 4 x 16 random bits of element group g (elements 4g .. 4g+3)
__device__ __forceinline__ unsigned long long drop_word(unsigned long long key, unsigned long long g) {
  return splitmix64(key + g * 0xD1342543DE82EF95ull);
}
// multiplier of element e (0..3) of a group: 0 when dropped, `scale` when kept
__device__ __forceinline__ float drop_mul(unsigned long long word, int e, unsigned thr, float scale) {
  return (static_cast<unsigned>(word >> (16 * e)) & 0xFFFFu) >= thr ? scale : 0.f;
}

// dst = src * m (bf16 or fp32, 8 elements per thread; src == dst allowed); `second` (bf16, optional, in place) gets the
// same mask: the saved gelu'(pre) next to gelu(pre), so that the fused dgrad epilogue multiplies by gelu' * m / (1 - p)
template <bool F32>
__global__ void __launch_bounds__(256)
dropout_kernel(const void* src_, void* dst_, uint4* __restrict__ second, long long n8,   // src_ may be dst_
               unsigned thr, float scale, const unsigned long long* __restrict__ seed, unsigned site) {
  const unsigned long long key = drop_key(seed, site);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned long long w0 = drop_word(key, 2ull * i), w1 = drop_word(key, 2ull * i + 1ull);
    float m[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      m[e] = drop_mul(w0, e, thr, scale);
      m[4 + e] = drop_mul(w1, e, thr, scale);
    }
    if constexpr (F32) {
      const float4* src = static_cast<const float4*>(src_);
      float4* dst = static_cast<float4*>(dst_);
      float4 a = src[2 * i], b = src[2 * i + 1];
      a.x *= m[0]; a.y *= m[1]; a.z *= m[2]; a.w *= m[3];
      b.x *= m[4]; b.y *= m[5]; b.z *= m[6]; b.w *= m[7];
      dst[2 * i] = a;
      dst[2 * i + 1] = b;
    } else {
      const uint4 v = static_cast<const uint4*>(src_)[i];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(w[e]);
        o[e] = pack_bf16x2(f.x * m[2 * e], f.y * m[2 * e + 1]);
      }
      static_cast<uint4*>(dst_)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (second != nullptr) {
      const uint4 v = second[i];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(w[e]);
        o[e] = pack_bf16x2(f.x * m[2 * e], f.y * m[2 * e + 1]);
      }
      second[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// y = residual + rowscale[row] * dropout(branch): the residual add of a Block's branch (x + drop_path(drop(branch)),
// VT.pyc@L150-151 with @L130 / @L104 inside) on the fp32 stream; branch bf16 [rows, D], residual / y fp32
__global__ void __launch_bounds__(256)
dropout_residual_kernel(const uint4* __restrict__ branch, const float4* __restrict__ residual,
                        const float* __restrict__ rowscale, float4* __restrict__ y, long long n8, int vec_per_row,
                        unsigned thr, float scale, const unsigned long long* __restrict__ seed, unsigned site) {
  const unsigned long long key = drop_key(seed, site);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned long long w0 = drop_word(key, 2ull * i), w1 = drop_word(key, 2ull * i + 1ull);
    const float rs = rowscale != nullptr ? __ldg(rowscale + i / vec_per_row) : 1.f;
    const uint4 v = __ldg(branch + i);
    const float2 f0 = unpack_bf16x2(v.x), f1 = unpack_bf16x2(v.y), f2 = unpack_bf16x2(v.z), f3 = unpack_bf16x2(v.w);
    float4 a = __ldg(residual + 2 * i), b = __ldg(residual + 2 * i + 1);
    a.x += rs * (f0.x * drop_mul(w0, 0, thr, scale));
    a.y += rs * (f0.y * drop_mul(w0, 1, thr, scale));
    a.z += rs * (f1.x * drop_mul(w0, 2, thr, scale));
    a.w += rs * (f1.y * drop_mul(w0, 3, thr, scale));
    b.x += rs * (f2.x * drop_mul(w1, 0, thr, scale));
    b.y += rs * (f2.y * drop_mul(w1, 1, thr, scale));
    b.z += rs * (f3.x * drop_mul(w1, 2, thr, scale));
    b.w += rs * (f3.y * drop_mul(w1, 3, thr, scale));
    y[2 * i] = a;
    y[2 * i + 1] = b;
  }
}

static int drop_params(float p, unsigned* thr, float* scale) {
  if (!(p >= 0.f && p <= 1.f)) return -1;
  long t = lrintf(p * 65536.f);
  *thr = static_cast<unsigned>(t < 0 ? 0 : (t > 65536 ? 65536 : t));
  *scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  return 0;
}
static int drop_grid(long long n8) {
  long long blocks = (n8 + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  return static_cast<int>(blocks > cap ? cap : blocks);
}

}  // namespace b200ssl

using namespace b200ssl;

extern "C" int b200ssl_dropout(const void* src, void* dst, void* second, long long n, int is_f32, float p,
                               const void* seed, unsigned site, void* stream) {
  if (n <= 0) return 0;
  unsigned thr;
  float scale;
  B200SSL_CHECK(drop_params(p, &thr, &scale) == 0, -2, "dropout: p=%f outside [0, 1]", static_cast<double>(p));
  B200SSL_CHECK(n % 8 == 0, -2, "dropout: element count %lld must be a multiple of 8", n);
  B200SSL_CHECK(seed != nullptr && (reinterpret_cast<uintptr_t>(seed) & 7) == 0, -2,
                "dropout: seed must be a device pointer to one 64-bit word");
  B200SSL_CHECK(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) |
                  reinterpret_cast<uintptr_t>(second)) & 15) == 0, -2, "dropout: operands must be 16-byte aligned");
  B200SSL_CHECK(!(is_f32 && second != nullptr), -2, "dropout: the second (bf16) tensor goes with a bf16 first one");
  const long long n8 = n / 8;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (is_f32)
    dropout_kernel<true><<<drop_grid(n8), 256, 0, s>>>(src, dst, nullptr, n8, thr, scale,
                                                       static_cast<const unsigned long long*>(seed), site);
  else
    dropout_kernel<false><<<drop_grid(n8), 256, 0, s>>>(src, dst, static_cast<uint4*>(second), n8, thr, scale,
                                                        static_cast<const unsigned long long*>(seed), site);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200ssl_dropout_residual(const void* branch, const float* residual, const float* rowscale, float* y,
                                        long long rows, int D, float p, const void* seed, unsigned site, void* stream) {
  if (rows <= 0) return 0;
  unsigned thr;
  float scale;
  B200SSL_CHECK(drop_params(p, &thr, &scale) == 0, -2, "dropout_residual: p=%f outside [0, 1]", static_cast<double>(p));
  B200SSL_CHECK(D > 0 && D % 8 == 0, -2, "dropout_residual: D=%d must be a multiple of 8", D);
  B200SSL_CHECK(seed != nullptr && (reinterpret_cast<uintptr_t>(seed) & 7) == 0, -2,
                "dropout_residual: seed must be a device pointer to one 64-bit word");
  B200SSL_CHECK(((reinterpret_cast<uintptr_t>(branch) | reinterpret_cast<uintptr_t>(residual) |
                  reinterpret_cast<uintptr_t>(y)) & 15) == 0, -2, "dropout_residual: operands must be 16-byte aligned");
  const long long n8 = rows * (D / 8);
  dropout_residual_kernel<<<drop_grid(n8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(branch), reinterpret_cast<const float4*>(residual), rowscale,
      reinterpret_cast<float4*>(y), n8, D / 8, thr, scale, static_cast<const unsigned long long*>(seed), site);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}
