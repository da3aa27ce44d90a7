// Small HBM-bound kernels around the GEMMs: patch gathering, token assembly (+CLS, +pos),
// column sums (bias gradients), fp32->bf16 weight casts, F.normalize and weight-norm rows.
//
// Reference sites: PatchEmbed.forward VT.pyc@L167-170 and prepare_tokens @L235-246 (patchify +
// assemble_tokens), DINOHead.forward @L326-330 (l2norm = F.normalize(dim=-1,p=2), weight-normed
// last layer @L315-318).
#include "common.cuh"

namespace b200ssl {

// img [B,C,H,W] bf16 -> A [B*(H/P)*(W/P), C*P*P] bf16, K index = (c, i, j) as in Conv2d weight.flatten(1).
// One thread moves one 16-byte piece (8 pixels of one patch row).
// IdxT = unsigned (whenever the piece count fits 32 bits: 64-bit div / mod per 16-byte piece made the kernel issue
// bound) or long long.
template <typename IdxT>
__global__ void patchify_kernel(const __nv_bfloat16* __restrict__ img, __nv_bfloat16* __restrict__ out, int B,
                                int C, int H, int W, int P) {
  const IdxT pw = W / P, ph = H / P, w8 = W / 8, uH = H, uC = C, uP = P;
  const IdxT total = static_cast<IdxT>(B) * uC * uH * w8;
  const IdxT stride = static_cast<IdxT>(gridDim.x) * blockDim.x;
  const long long row_elems = static_cast<long long>(C) * P * P;
  for (IdxT idx = static_cast<IdxT>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    // idx enumerates the image in memory order so global reads are fully coalesced
    const IdxT x8 = idx % w8;
    IdxT r = idx / w8;
    const IdxT y = r % uH;
    r /= uH;
    const IdxT c = r % uC;
    const IdxT b = r / uC;
    const IdxT px = (x8 * 8) / uP, j = (x8 * 8) % uP;
    const IdxT py = y / uP, i = y % uP;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(img) + idx);
    const long long row = static_cast<long long>((b * ph + py) * pw + px);
    const long long col = static_cast<long long>((c * uP + i) * uP + j);
    *reinterpret_cast<uint4*>(out + row * row_elems + col) = v;
  }
}

// x[b,0,:] = cls + pos[0] ; x[b,1+p,:] = y[b*Np+p,:] + pos[1+p]   (pos, cls fp32; y bf16; x = fp32 stream)
template <typename IdxT>
__global__ void assemble_tokens_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ cls,
                                       const float* __restrict__ pos, float* __restrict__ x, int B, int Np,
                                       int D) {
  const IdxT N = Np + 1;
  const IdxT d8 = D / 8;
  const IdxT total = static_cast<IdxT>(B) * N * d8;
  const IdxT stride = static_cast<IdxT>(gridDim.x) * blockDim.x;
  for (IdxT idx = static_cast<IdxT>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const IdxT c8 = idx % d8;
    const IdxT rn = idx / d8;
    const IdxT n = rn % N;
    const IdxT b = rn / N;
    float v[8];
    if (n == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __ldg(cls + c8 * 8 + e);
    } else {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(y + static_cast<long long>(b * Np + n - 1) * D) + c8);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(w[e]);
        v[2 * e] = f.x;
        v[2 * e + 1] = f.y;
      }
    }
    const float4* pp = reinterpret_cast<const float4*>(pos + static_cast<long long>(n) * D + c8 * 8);
    const float4 p0 = __ldg(pp), p1 = __ldg(pp + 1);
    float4* px = reinterpret_cast<float4*>(x + static_cast<long long>(rn) * D) + 2 * c8;
    px[0] = make_float4(v[0] + p0.x, v[1] + p0.y, v[2] + p0.z, v[3] + p0.w);
    px[1] = make_float4(v[4] + p1.x, v[5] + p1.y, v[6] + p1.z, v[7] + p1.w);
  }
}

// dy[b*Np+p,:] = dx[b,1+p,:] ; dpos[n,:] += sum_b dx[b,n,:]   (block = (n, batch segment))
__global__ void assemble_tokens_bwd_kernel(const __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dy,
                                           float* __restrict__ dpos, int B, int Np, int D, int bseg) {
  const int N = Np + 1;
  const int n = blockIdx.x;
  const int b_lo = blockIdx.y * bseg, b_hi = min(B, b_lo + bseg);
  for (int c8 = threadIdx.x; c8 < D / 8; c8 += blockDim.x) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = b_lo; b < b_hi; ++b) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(dx + (static_cast<long long>(b) * N + n) * D) + c8);
      if (n > 0) reinterpret_cast<uint4*>(dy + (static_cast<long long>(b) * Np + n - 1) * D)[c8] = u;
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(w[e]);
        acc[2 * e] += f.x;
        acc[2 * e + 1] += f.y;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(dpos + static_cast<long long>(n) * D + c8 * 8 + e, acc[e]);
  }
}

// out[c] (+)= sum_r x[r, c]   (bf16 in, fp32 atomics out); block = 32 column-chunks x 8 row lanes
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, float* __restrict__ out, long long rows, int ncols,
              int rows_per_block) {
  const int c8 = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  __shared__ float red[8][32][8];
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  if (c8 * 8 < ncols) {
    for (long long r = r0 + rl; r < r1; r += 8) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + r * ldx) + c8);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(w[e]);
        acc[2 * e] += f.x;
        acc[2 * e + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rl][threadIdx.x & 31][e] = acc[e];
  __syncthreads();
  if (rl == 0 && c8 * 8 < ncols) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x][e];
      atomicAdd(out + c8 * 8 + e, s);
    }
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long n4 = n / 4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
  for (long long i = n4 * 4 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __float2bfloat16(src[i]);
}

__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long n4 = n / 4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(src) + i);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    reinterpret_cast<float4*>(dst)[i] = make_float4(a.x, a.y, b.x, b.y);
  }
  for (long long i = n4 * 4 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __bfloat162float(src[i]);
}

// y = x / max(||x||_2, eps) per row (F.normalize, eps 1e-12). One warp per row.
__global__ void l2norm_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                  float* __restrict__ norm_out, long long rows, int D, float eps) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const __nv_bfloat16* px = x + row * D;
  float ss = 0.f;
  for (int i = lane * 2; i < D; i += 64) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(px + i));
    ss += f.x * f.x + f.y * f.y;
  }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  const float inv = 1.f / fmaxf(nrm, eps);
  for (int i = lane * 2; i < D; i += 64) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(px + i));
    *reinterpret_cast<__nv_bfloat162*>(y + row * D + i) = __floats2bfloat162_rn(f.x * inv, f.y * inv);
  }
  if (lane == 0) norm_out[row] = nrm;
}

// dx = (dy - y * <y, dy>) / max(norm, eps)
__global__ void l2norm_bwd_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dy,
                                  const float* __restrict__ norm, __nv_bfloat16* __restrict__ dx, long long rows,
                                  int D, float eps) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float dot = 0.f;
  for (int i = lane * 2; i < D; i += 64) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(y + row * D + i));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dy + row * D + i));
    dot += a.x * b.x + a.y * b.y;
  }
  dot = warp_sum(dot);
  const float inv = 1.f / fmaxf(norm[row], eps);
  for (int i = lane * 2; i < D; i += 64) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(y + row * D + i));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dy + row * D + i));
    *reinterpret_cast<__nv_bfloat162*>(dx + row * D + i) =
        __floats2bfloat162_rn((b.x - a.x * dot) * inv, (b.y - a.y * dot) * inv);
  }
}

// w[r,:] = g[r] * v[r,:] / ||v[r,:]||  (fp32 v, g -> bf16 w); g == nullptr means g = 1
__global__ void weightnorm_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                      __nv_bfloat16* __restrict__ w, float* __restrict__ norm_out, long long rows,
                                      int D) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* pv = v + row * D;
  float ss = 0.f;
  for (int i = lane * 4; i < D; i += 128) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(pv + i));
    ss += f.x * f.x + f.y * f.y + f.z * f.z + f.w * f.w;
  }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  const float s = (g ? g[row] : 1.f) / nrm;
  for (int i = lane * 4; i < D; i += 128) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(pv + i));
    *reinterpret_cast<uint2*>(w + row * D + i) = make_uint2(pack_bf16x2(f.x * s, f.y * s), pack_bf16x2(f.z * s, f.w * s));
  }
  if (lane == 0) norm_out[row] = nrm;
}

// dg[r] = <dw, v>/||v|| ; dv = g/||v|| * dw - g*<dw,v>/||v||^3 * v
__global__ void weightnorm_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                      const float* __restrict__ norm, const float* __restrict__ dw,
                                      float* __restrict__ dv, float* __restrict__ dg, long long rows, int D,
                                      int accumulate) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float dot = 0.f;
  for (int i = lane * 4; i < D; i += 128) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(v + row * D + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(dw + row * D + i));
    dot += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  }
  dot = warp_sum(dot);
  const float nrm = norm[row];
  const float gg = g ? g[row] : 1.f;
  const float c1 = gg / nrm, c2 = gg * dot / (nrm * nrm * nrm);
  for (int i = lane * 4; i < D; i += 128) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(v + row * D + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(dw + row * D + i));
    float4 o = make_float4(c1 * b.x - c2 * a.x, c1 * b.y - c2 * a.y, c1 * b.z - c2 * a.z, c1 * b.w - c2 * a.w);
    if (accumulate) {  // gradient sink: dv is the parameter's .grad (zeroed at the start of the step)
      const float4 p = *reinterpret_cast<const float4*>(dv + row * D + i);
      o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
    }
    *reinterpret_cast<float4*>(dv + row * D + i) = o;
  }
  if (dg && lane == 0) dg[row] = (accumulate ? dg[row] : 0.f) + dot / nrm;
}

static int grid_for(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace b200ssl

using namespace b200ssl;

extern "C" int b200ssl_patchify(const void* img, void* out, int B, int C, int H, int W, int P, void* stream) {
  B200SSL_CHECK(P % 8 == 0 && H % P == 0 && W % P == 0, -2, "patchify: H=%d W=%d must divide by P=%d (P %% 8 == 0)", H, W, P);
  const long long total = static_cast<long long>(B) * C * H * (W / 8);
  if (total < (1ll << 31))
    patchify_kernel<unsigned><<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(img), static_cast<__nv_bfloat16*>(out), B, C, H, W, P);
  else
    patchify_kernel<long long><<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(img), static_cast<__nv_bfloat16*>(out), B, C, H, W, P);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200ssl_assemble_tokens(const void* y, const float* cls, const float* pos, void* x, int B, int Np,
                                       int D, void* stream) {
  B200SSL_CHECK(D % 8 == 0, -2, "assemble_tokens: D=%d must be a multiple of 8", D);
  const long long total = static_cast<long long>(B) * (Np + 1) * (D / 8);
  if (total < (1ll << 31))
    assemble_tokens_kernel<unsigned><<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(y), cls, pos, static_cast<float*>(x), B, Np, D);
  else
    assemble_tokens_kernel<long long><<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(y), cls, pos, static_cast<float*>(x), B, Np, D);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

// dpos [Np+1, D] fp32 is zeroed here; dcls [D] receives dpos[0] (x[b,0] = cls + pos[0]).
extern "C" int b200ssl_assemble_tokens_bwd(const void* dx, void* dy, float* dpos, float* dcls, int B, int Np, int D,
                                           void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  B200SSL_CHECK(D % 8 == 0, -2, "assemble_tokens_bwd: D=%d must be a multiple of 8", D);
  B200SSL_CUDA(cudaMemsetAsync(dpos, 0, sizeof(float) * static_cast<size_t>(Np + 1) * D, s));
  const int bseg = 32;
  dim3 grid(Np + 1, (B + bseg - 1) / bseg);
  assemble_tokens_bwd_kernel<<<grid, 64, 0, s>>>(static_cast<const __nv_bfloat16*>(dx),
                                                  static_cast<__nv_bfloat16*>(dy), dpos, B, Np, D, bseg);
  B200SSL_CUDA(cudaGetLastError());
  if (dcls) B200SSL_CUDA(cudaMemcpyAsync(dcls, dpos, sizeof(float) * D, cudaMemcpyDeviceToDevice, s));
  return 0;
}

extern "C" int b200ssl_colsum(const void* x, long long ldx, float* out, long long rows, int ncols, int accumulate,
                              void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  B200SSL_CHECK(ncols % 8 == 0 && ldx % 8 == 0, -2, "colsum: ncols/ldx must be multiples of 8");
  if (!accumulate) B200SSL_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * ncols, s));
  const int col_blocks = (ncols / 8 + 31) / 32;
  int row_blocks = (sm_count() * 4 + col_blocks - 1) / col_blocks;
  long long rpb = (rows + row_blocks - 1) / row_blocks;
  if (rpb < 64) rpb = 64;
  row_blocks = static_cast<int>((rows + rpb - 1) / rpb);
  dim3 grid(col_blocks, row_blocks);
  colsum_kernel<<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), ldx, out, rows, ncols,
                                     static_cast<int>(rpb));
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200ssl_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream) {
  B200SSL_CHECK((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0, -2,
                "cast: pointers must be 16B/8B aligned");
  cast_f32_bf16_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), n);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200ssl_cast_bf16_to_f32(const void* src, float* dst, long long n, void* stream) {
  B200SSL_CHECK((reinterpret_cast<uintptr_t>(src) & 7) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, -2,
                "cast: pointers must be 8B/16B aligned");
  cast_bf16_f32_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), dst, n);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200ssl_l2norm_fwd(const void* x, void* y, float* norm, long long rows, int D, float eps,
                                  void* stream) {
  B200SSL_CHECK(D % 2 == 0, -2, "l2norm: D must be even");
  const long long threads = rows * 32;
  l2norm_fwd_kernel<<<static_cast<int>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), norm, rows, D, eps);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200ssl_l2norm_bwd(const void* y, const void* dy, const float* norm, void* dx, long long rows, int D,
                                  float eps, void* stream) {
  const long long threads = rows * 32;
  l2norm_bwd_kernel<<<static_cast<int>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(y), static_cast<const __nv_bfloat16*>(dy), norm,
      static_cast<__nv_bfloat16*>(dx), rows, D, eps);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200ssl_weightnorm_fwd(const float* v, const float* g, void* w, float* norm, long long rows, int D,
                                      void* stream) {
  B200SSL_CHECK(D % 4 == 0, -2, "weightnorm: D must be a multiple of 4");
  const long long threads = rows * 32;
  weightnorm_fwd_kernel<<<static_cast<int>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      v, g, static_cast<__nv_bfloat16*>(w), norm, rows, D);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200ssl_weightnorm_bwd(const float* v, const float* g, const float* norm, const float* dw, float* dv,
                                      float* dg, long long rows, int D, int accumulate, void* stream) {
  B200SSL_CHECK(D % 4 == 0, -2, "weightnorm: D must be a multiple of 4");
  const long long threads = rows * 32;
  weightnorm_bwd_kernel<<<static_cast<int>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      v, g, norm, dw, dv, dg, rows, D, accumulate);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

// y[r, :] = x[r, :] * scale[r]   (bf16 rows, fp32 per-row scale): the branch gradient under stochastic depth
// (drop_path VT.pyc@L66-74: the kept samples' branch is scaled by 1/keep, the dropped ones' by 0)
namespace b200ssl {
__global__ void __launch_bounds__(256)
scale_rows_kernel(const uint4* __restrict__ x, const float* __restrict__ scale, uint4* __restrict__ y, long long rows,
                  int vec_per_row) {
  const long long total = rows * vec_per_row;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float s = __ldg(scale + i / vec_per_row);
    const uint4 v = __ldg(x + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_bf16x2(w[e]);
      o[e] = pack_bf16x2(f.x * s, f.y * s);
    }
    y[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
}  // namespace b200ssl

extern "C" int b200ssl_scale_rows(const void* x, const float* scale, void* y, long long rows, int D, void* stream) {
  B200SSL_CHECK(D % 8 == 0 && rows > 0, -2, "scale_rows: D=%d must be a multiple of 8", D);
  B200SSL_CHECK(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, -2,
                "scale_rows: operands must be 16-byte aligned");
  const long long total = rows * (D / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(b200ssl::sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  b200ssl::scale_rows_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), scale, static_cast<uint4*>(y), rows, D / 8);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Plumbing that used to be PyTorch glue inside the captured step (torch.zeros / clone / slicing + cat / indexed
// assignment / tiny matmuls): memset and copy nodes, a strided row copy, an fp32 accumulate and the position-table
// resize as a fixed linear map.
// ------------------------------------------------------------------------------------------------
extern "C" int b200ssl_zero_bytes(void* p, long long nbytes, void* stream) {
  if (nbytes <= 0) return 0;
  B200SSL_CUDA(cudaMemsetAsync(p, 0, static_cast<size_t>(nbytes), static_cast<cudaStream_t>(stream)));
  return 0;
}

extern "C" int b200ssl_copy_bytes(void* dst, const void* src, long long nbytes, void* stream) {
  if (nbytes <= 0) return 0;
  B200SSL_CUDA(cudaMemcpyAsync(dst, src, static_cast<size_t>(nbytes), cudaMemcpyDeviceToDevice,
                               static_cast<cudaStream_t>(stream)));
  return 0;
}

namespace b200ssl {
// dst[r * dst_stride + :] = src[r * src_stride + :] in 16-byte pieces (strides counted in 16-byte units)
__global__ void __launch_bounds__(256)
copy_rows_kernel(const uint4* __restrict__ src, long long src_stride, uint4* __restrict__ dst, long long dst_stride,
                 long long rows, int vec_per_row) {
  const long long total = rows * vec_per_row;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / vec_per_row;
    const int c = static_cast<int>(i - r * vec_per_row);
    dst[r * dst_stride + c] = __ldg(src + r * src_stride + c);
  }
}

__global__ void __launch_bounds__(256)
add_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] += __ldg(src + i);
}

// Position-table resize (VisionTransformer.interpolate_pos_encoding, VT.pyc@L213-233) as the fixed linear map it is:
//   forward : out[0, :] = pos[0, :] ;  out[1 + i, :] = sum_j mat[i, j] * pos[1 + j, :]          (mat [Mo, Ki], fp32)
//   backward: dpos[0, :] += dout[0, :] ;  dpos[1 + j, :] += sum_i mat[i, j] * dout[1 + i, :]
// one thread per output element; Mo, Ki <= a few hundred, D <= 768: microseconds.
__global__ void __launch_bounds__(128)
pos_interp_fwd_kernel(const float* __restrict__ mat, const float* __restrict__ pos, float* __restrict__ out, int Mo,
                      int Ki, int D) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;  // output row, 0 = class position
  if (d >= D) return;
  if (i == 0) {
    out[d] = __ldg(pos + d);
    return;
  }
  const float* m = mat + static_cast<long long>(i - 1) * Ki;
  // four independent accumulators, unrolled: the loop is a chain of dependent L2 loads otherwise (40 us measured)
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int j = 0;
#pragma unroll 4
  for (; j + 4 <= Ki; j += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
      acc[u] = fmaf(__ldg(m + j + u), __ldg(pos + static_cast<long long>(1 + j + u) * D + d), acc[u]);
  }
  for (; j < Ki; ++j) acc[0] = fmaf(__ldg(m + j), __ldg(pos + static_cast<long long>(1 + j) * D + d), acc[0]);
  out[static_cast<long long>(i) * D + d] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

__global__ void __launch_bounds__(128)
pos_interp_bwd_kernel(const float* __restrict__ mat, const float* __restrict__ dout, float* __restrict__ dpos, int Mo,
                      int Ki, int D) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;  // input row, 0 = class position
  if (d >= D) return;
  if (j == 0) {
    dpos[d] += __ldg(dout + d);
    return;
  }
  float acc = 0.f;
  for (int i = 0; i < Mo; ++i)
    acc = fmaf(__ldg(mat + static_cast<long long>(i) * Ki + (j - 1)), __ldg(dout + static_cast<long long>(1 + i) * D + d), acc);
  dpos[static_cast<long long>(j) * D + d] += acc;
}
}  // namespace b200ssl

// rows x row_bytes from a strided source to a strided destination (gather of the CLS rows, scatter of their gradient)
extern "C" int b200ssl_copy_rows(const void* src, long long src_stride_bytes, void* dst, long long dst_stride_bytes,
                                 long long rows, int row_bytes, void* stream) {
  if (rows <= 0) return 0;
  B200SSL_CHECK(row_bytes > 0 && row_bytes % 16 == 0 && src_stride_bytes % 16 == 0 && dst_stride_bytes % 16 == 0 &&
                    ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0,
                -2, "copy_rows: pointers, strides and row size must be multiples of 16 bytes");
  const long long total = rows * (row_bytes / 16);
  b200ssl::copy_rows_kernel<<<b200ssl::grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(src), src_stride_bytes / 16, static_cast<uint4*>(dst), dst_stride_bytes / 16, rows,
      row_bytes / 16);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

// dst[i] += src[i], fp32
extern "C" int b200ssl_add_f32(float* dst, const float* src, long long n, void* stream) {
  if (n <= 0) return 0;
  b200ssl::add_f32_kernel<<<b200ssl::grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dst, src, n);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}

// transpose = 0: out[1 + Mo, D] = resize(pos[1 + Ki, D]); transpose = 1: pos (as dpos) += resize^T(out (as dout))
extern "C" int b200ssl_pos_interp(const float* mat, const float* in, float* out, int Mo, int Ki, int D, int transpose,
                                  void* stream) {
  B200SSL_CHECK(Mo > 0 && Ki > 0 && D > 0, -2, "pos_interp: empty problem");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!transpose) {
    dim3 grid((D + 127) / 128, Mo + 1);
    b200ssl::pos_interp_fwd_kernel<<<grid, 128, 0, s>>>(mat, in, out, Mo, Ki, D);
  } else {
    dim3 grid((D + 127) / 128, Ki + 1);
    b200ssl::pos_interp_bwd_kernel<<<grid, 128, 0, s>>>(mat, in, out, Mo, Ki, D);
  }
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}
