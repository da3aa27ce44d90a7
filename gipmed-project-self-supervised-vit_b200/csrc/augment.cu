// Multi-crop tile augmentation on the GPU: uint8 RGB tiles -> normalised bf16 crops.
//
// Replaces the reference's per-tile host pipeline (datasets.py:498-502 -> transformations.py:103-208: PIL ColorJitter,
// GaussianBlur(3, sigma <= 0.1), MyGaussianNoiseTransform, RandomVerticalFlip, MyRotation{0,90,180,270}, RandomAffine
// scale, ToTensor, Normalize(mean / std of `norm_type`)), which runs on 2 CPU cores per job in the reference
// (sbatch-ssl.sh:20), and adds the DINO multi-crop geometry (random-resized crops at two resolutions) in the same pass.
//
// One CTA per source tile. The whole 256 x 256 x 3 uint8 tile (192 KB) is staged in shared memory ONCE; every crop of
// the tile (2 global + n local) is then produced from it: HBM traffic = 1 tile read + the crop writes, nothing else.
// Per crop (parameters come from a host-sampled table, so a CPU oracle can replay exactly the same draw):
//   geometry : integer crop box -> bilinear resize (align_corners = false, 2 taps, torch F.interpolate arithmetic) ->
//              horizontal / vertical flip -> rotation by k * 90 degrees (torch.rot90); a zoom about the centre
//              (RandomAffine(scale)) is the same thing as a smaller crop box and is folded into the box by the sampler;
//   colour   : brightness / contrast / saturation / hue in the drawn order, torchvision tensor semantics (contrast blends
//              with the mean grey level of the image AT THAT POINT of the pipeline: a first pass over the crop computes it);
//   noise    : x + sigma * n, n ~ N(0,1) from a counter-based generator (PCG hash of (seed, element pair), Box-Muller),
//              clamped to [0,1] (skimage.util.random_noise(mode='gaussian', clip=True));
//   output   : (x - mean[c]) / std[c] -> bf16, CHW;
//   cutout   : the reference's Cutout (transformations.py:10-45, appended behind Normalize at :206-207) multiplies the
//              NORMALISED image by a mask that is 0 inside one square hole: the hole's pixels are written as 0.
// GaussianBlur(3, sigma in (1e-7, 0.1)) is the identity to 2e-22 (off-centre tap weight exp(-1 / (2 * 0.1^2))) and is
// skipped; intermediate uint8 re-quantisation of the PIL pipeline is not reproduced (float pipeline, documented).
#include "common.cuh"

namespace b200ssl {

constexpr int AUG_TILE = 256;                          // source tile side (train.py:415)
constexpr int AUG_TILE_BYTES = AUG_TILE * AUG_TILE * 3;
// shared-memory row pitch: 768 B of pixels + 4 B of padding = 193 words, so that a COLUMN walk (crops rotated by 90 or
// 270 degrees read the tile down its columns) steps through all 32 banks instead of hammering one (768 B = 192 words
// = a multiple of 32 banks: 2.9e8 bank conflicts per launch in the first ncu capture, issue slots 48 % used)
constexpr int AUG_ROW_PITCH = AUG_TILE * 3 + 4;
constexpr int AUG_SMEM_BYTES = AUG_TILE * AUG_ROW_PITCH;
constexpr int AUG_THREADS = 1024;
constexpr int AUG_PARAM_WORDS = 16;

// one row of the parameter table (16 x 32 bit) per (tile, crop)
struct AugParams {
  int top, left, h, w;   // crop box in source pixels
  unsigned flags;        // bit 0 hflip, bit 1 vflip, bits 2-3 rot90 k, bits 4-11 colour-op order (4 x 2 bit), bit 12 jitter on
  float brightness, contrast, saturation, hue;
  float sigma;           // noise standard deviation (0 = none)
  unsigned seed;
  unsigned cut_y, cut_x;  // Cutout hole in output-frame pixels: y1 | y2 << 16, x1 | x2 << 16 (empty when y2 <= y1 or x2 <= x1)
  unsigned pad[3];
};
static_assert(sizeof(AugParams) == AUG_PARAM_WORDS * 4, "parameter row is 16 words");

__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }
__device__ __forceinline__ float gray_of(float r, float g, float b) { return 0.2989f * r + 0.587f * g + 0.114f * b; }

// torchvision.transforms._functional_tensor.adjust_hue: _rgb2hsv -> h = (h + f) % 1 -> _hsv2rgb
__device__ __forceinline__ void adjust_hue(float& r, float& g, float& b, float f) {
  const float maxc = fmaxf(r, fmaxf(g, b)), minc = fminf(r, fminf(g, b));
  const bool eqc = maxc == minc;
  const float cr = maxc - minc;
  const float s = __fdividef(cr, eqc ? 1.f : maxc);
  const float idiv = __fdividef(1.f, eqc ? 1.f : cr);
  const float rc = (maxc - r) * idiv, gc = (maxc - g) * idiv, bc = (maxc - b) * idiv;
  const float hr = (maxc == r) ? (bc - gc) : 0.f;
  const float hg = ((maxc == g) && (maxc != r)) ? (2.f + rc - bc) : 0.f;
  const float hb = ((maxc != g) && (maxc != r)) ? (4.f + gc - rc) : 0.f;
  float h = (hr + hg + hb) * (1.f / 6.f) + 1.f;   // in [5/6, 2): fmod(h, 1) == h - floor(h)
  h = h - floorf(h);
  h = h + f;
  h = h - floorf(h);  // python-style % 1.0
  const float v = maxc;
  const float h6 = h * 6.f;
  const float fi = floorf(h6);
  const float fr = h6 - fi;
  int i = static_cast<int>(fi) % 6;
  const float p = clamp01(v * (1.f - s)), q = clamp01(v * (1.f - fr * s)), t = clamp01(v * (1.f - (1.f - fr) * s));
  switch (i) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

__device__ __forceinline__ unsigned pcg_out(unsigned x) {
  const unsigned word = ((x >> ((x >> 28) + 4u)) ^ x) * 277803737u;
  return (word >> 22) ^ word;
}
// two standard normals from (seed, element-pair index): two PCG outputs -> Box-Muller; the even element of the pair takes
// the cosine branch, the odd one the sine branch
__device__ __forceinline__ void normal_pair(unsigned seed, unsigned pair, float& n_even, float& n_odd) {
  const unsigned x1 = pair * 747796405u + seed * 2891336453u + 12345u;
  const unsigned x2 = x1 * 747796405u + 2891336453u;
  const float u1 = (static_cast<float>(pcg_out(x1) >> 8) + 1.f) * (1.f / 16777216.f);  // (0, 1]
  const float u2 = static_cast<float>(pcg_out(x2) >> 8) * (1.f / 16777216.f);           // [0, 1)
  const float rad = sqrtf(-2.f * __logf(u1));
  float sn, cs;
  __sincosf(6.283185307179586f * u2, &sn, &cs);
  n_even = rad * cs;
  n_odd = rad * sn;
}

// the resized crop's pixel (v, u) before flips / rotation: bilinear taps from the staged tile, torch arithmetic
__device__ __forceinline__ void sample_rgb(const uint8_t* tile, const AugParams& p, float sch, float scw, int v, int u,
                                           float& r, float& g, float& b) {
  float sy = sch * (static_cast<float>(v) + 0.5f) - 0.5f;
  float sx = scw * (static_cast<float>(u) + 0.5f) - 0.5f;
  sy = fmaxf(sy, 0.f);
  sx = fmaxf(sx, 0.f);
  const int y0 = min(static_cast<int>(sy), p.h - 1), x0 = min(static_cast<int>(sx), p.w - 1);
  const int y1 = min(y0 + 1, p.h - 1), x1 = min(x0 + 1, p.w - 1);
  const float ly = sy - static_cast<float>(y0), lx = sx - static_cast<float>(x0);
  const float hy = 1.f - ly, hx = 1.f - lx;
  const uint8_t* p00 = tile + (p.top + y0) * AUG_ROW_PITCH + (p.left + x0) * 3;
  const uint8_t* p01 = tile + (p.top + y0) * AUG_ROW_PITCH + (p.left + x1) * 3;
  const uint8_t* p10 = tile + (p.top + y1) * AUG_ROW_PITCH + (p.left + x0) * 3;
  const uint8_t* p11 = tile + (p.top + y1) * AUG_ROW_PITCH + (p.left + x1) * 3;
  // uint8 -> float through the 2^23 trick on the ALU / FMA pipes (I2F runs on the quarter-rate XU pipe: with twelve
  // taps per pixel it was the kernel's bound, 68 % XU utilisation in the first ncu capture)
  auto f8 = [](uint8_t v) { return __uint_as_float(0x4B000000u | static_cast<unsigned>(v)) - 8388608.f; };
  const float k = 1.f / 255.f;
  r = (hy * (hx * f8(p00[0]) + lx * f8(p01[0])) + ly * (hx * f8(p10[0]) + lx * f8(p11[0]))) * k;
  g = (hy * (hx * f8(p00[1]) + lx * f8(p01[1])) + ly * (hx * f8(p10[1]) + lx * f8(p11[1]))) * k;
  b = (hy * (hx * f8(p00[2]) + lx * f8(p01[2])) + ly * (hx * f8(p10[2]) + lx * f8(p11[2]))) * k;
}

// colour ops [first, last) of the drawn order; `mean_gray` is what contrast blends with
__device__ __forceinline__ void colour_ops(const AugParams& p, int first, int last, float mean_gray, float& r, float& g,
                                           float& b) {
  for (int s = first; s < last; ++s) {
    const int op = (p.flags >> (4 + 2 * s)) & 3;
    if (op == 0) {
      r = clamp01(r * p.brightness); g = clamp01(g * p.brightness); b = clamp01(b * p.brightness);
    } else if (op == 1) {
      const float f = p.contrast, m = (1.f - f) * mean_gray;
      r = clamp01(f * r + m); g = clamp01(f * g + m); b = clamp01(f * b + m);
    } else if (op == 2) {
      const float f = p.saturation, m = (1.f - f) * gray_of(r, g, b);
      r = clamp01(f * r + m); g = clamp01(f * g + m); b = clamp01(f * b + m);
    } else {
      adjust_hue(r, g, b, p.hue);
    }
  }
}

// output pixel (oy, ox) -> pixel (v, u) of the resized crop: undo torch.rot90(k), then the flips
__device__ __forceinline__ void unmap(const AugParams& p, int S, int oy, int ox, int& v, int& u) {
  const int k = (p.flags >> 2) & 3;
  if (k == 0) { v = oy; u = ox; }
  else if (k == 1) { v = ox; u = S - 1 - oy; }
  else if (k == 2) { v = S - 1 - oy; u = S - 1 - ox; }
  else { v = S - 1 - ox; u = oy; }
  if (p.flags & 2u) v = S - 1 - v;
  if (p.flags & 1u) u = S - 1 - u;
}

__global__ void __launch_bounds__(AUG_THREADS, 1)
multicrop_augment_kernel(const uint8_t* __restrict__ tiles, const AugParams* __restrict__ params,
                         __nv_bfloat16* __restrict__ out_global, __nv_bfloat16* __restrict__ out_local, int B, int n_global,
                         int n_local, int Sg, int Sl, float mean0, float mean1, float mean2, float istd0, float istd1,
                         float istd2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tile = smem_raw;
  __shared__ float red[32];
  __shared__ float mean_sh;
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  {
    constexpr int kRowWords = AUG_TILE * 3 / 4;  // 192
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tiles + static_cast<long long>(b) * AUG_TILE_BYTES);
    uint32_t* dst = reinterpret_cast<uint32_t*>(tile);
    for (int i = tid; i < AUG_TILE * kRowWords; i += AUG_THREADS)
      dst[(i / kRowWords) * (AUG_ROW_PITCH / 4) + i % kRowWords] = __ldg(src + i);
  }
  __syncthreads();
  const int ncrops = n_global + n_local;
  for (int c = 0; c < ncrops; ++c) {
    AugParams p = params[static_cast<long long>(b) * ncrops + c];
    // the table is caller data: a box outside the tile must not become an out-of-bounds shared-memory read
    p.top = min(max(p.top, 0), AUG_TILE - 1);
    p.left = min(max(p.left, 0), AUG_TILE - 1);
    p.h = min(max(p.h, 1), AUG_TILE - p.top);
    p.w = min(max(p.w, 1), AUG_TILE - p.left);
    const bool is_g = c < n_global;
    const int S = is_g ? Sg : Sl;
    const float sch = static_cast<float>(p.h) / static_cast<float>(S), scw = static_cast<float>(p.w) / static_cast<float>(S);
    const bool jitter = (p.flags >> 12) & 1u;
    // position of the contrast op in the drawn order (4 = absent)
    int cpos = 4;
    if (jitter) {
      for (int s = 0; s < 4; ++s)
        if (((p.flags >> (4 + 2 * s)) & 3) == 1) cpos = s;
    }
    float mean_gray = 0.f;
    if (cpos < 4) {
      // pass A: mean grey level of the crop after the ops that precede contrast (the mean is invariant under the
      // flips and rotations, so the crop is walked in its own pixel order)
      float acc = 0.f;
      for (int v = tid >> 5; v < S; v += AUG_THREADS / 32) {      // a warp per row: no per-pixel division by S
        for (int u = tid & 31; u < S; u += 32) {
          float r, g, bl;
          sample_rgb(tile, p, sch, scw, v, u, r, g, bl);
          colour_ops(p, 0, cpos, 0.f, r, g, bl);
          acc += gray_of(r, g, bl);
        }
      }
      acc = warp_sum(acc);
      if ((tid & 31) == 0) red[tid >> 5] = acc;
      __syncthreads();
      if (tid < 32) {
        float t = red[tid];
        t = warp_sum(t);
        if (tid == 0) mean_sh = t / static_cast<float>(S * S);
      }
      __syncthreads();
      mean_gray = mean_sh;
    }
    // pass B: two horizontally adjacent output pixels per thread, one 4-byte store per channel
    __nv_bfloat16* out = is_g ? out_global + (static_cast<long long>(c) * B + b) * 3 * Sg * Sg
                              : out_local + (static_cast<long long>(c - n_global) * B + b) * 3 * Sl * Sl;
    const int half = S / 2;
    const int cut_y1 = p.cut_y & 0xFFFFu, cut_y2 = p.cut_y >> 16, cut_x1 = p.cut_x & 0xFFFFu, cut_x2 = p.cut_x >> 16;
    for (int oy = tid >> 5; oy < S; oy += AUG_THREADS / 32)
    for (int ox2 = tid & 31; ox2 < half; ox2 += 32) {
      const int ox = ox2 * 2;
      float px[2][3];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        int v, u;
        unmap(p, S, oy, ox + e, v, u);
        sample_rgb(tile, p, sch, scw, v, u, px[e][0], px[e][1], px[e][2]);
        if (jitter) colour_ops(p, 0, 4, mean_gray, px[e][0], px[e][1], px[e][2]);
      }
      if (p.sigma > 0.f) {
        // element index (ch, oy, ox) of the output crop; ox is even here, so (ox, ox + 1) is one Box-Muller pair
        const unsigned pair0 = static_cast<unsigned>(oy * S + ox) >> 1, plane = static_cast<unsigned>(S * S) >> 1;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          float n0, n1;
          normal_pair(p.seed, pair0 + ch * plane, n0, n1);
          px[0][ch] = clamp01(px[0][ch] + p.sigma * n0);
          px[1][ch] = clamp01(px[1][ch] + p.sigma * n1);
        }
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        px[e][0] = (px[e][0] - mean0) * istd0;
        px[e][1] = (px[e][1] - mean1) * istd1;
        px[e][2] = (px[e][2] - mean2) * istd2;
      }
      if (oy >= cut_y1 && oy < cut_y2) {
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if (ox + e >= cut_x1 && ox + e < cut_x2) px[e][0] = px[e][1] = px[e][2] = 0.f;
      }
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
        *reinterpret_cast<uint32_t*>(out + (static_cast<long long>(ch) * S + oy) * S + ox) = pack_bf16x2(px[0][ch], px[1][ch]);
    }
    __syncthreads();  // mean_sh / red are re-used by the next crop
  }
}

}  // namespace b200ssl

using namespace b200ssl;

// tiles uint8 [B, 256, 256, 3] (HWC, RGB); params int32/float32 [B, n_global + n_local, 16] (AugParams rows);
// out_global bf16 [n_global, B, 3, Sg, Sg], out_local bf16 [n_local, B, 3, Sl, Sl] (crop-major, the layout
// GraphedDinoStep's static inputs use); mean / std: the 3 channel statistics of Normalize.
extern "C" int b200ssl_multicrop_augment(const void* tiles, const void* params, void* out_global, void* out_local, int B,
                                         int n_global, int n_local, int size_global, int size_local, const float* mean,
                                         const float* std, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200SSL_CHECK(B > 0 && n_global >= 0 && n_local >= 0 && n_global + n_local > 0, -2, "augment: empty problem");
  B200SSL_CHECK((n_global == 0 || (size_global >= 2 && size_global % 2 == 0)) &&
                    (n_local == 0 || (size_local >= 2 && size_local % 2 == 0)),
                -2, "augment: crop sizes must be even (got %d, %d)", size_global, size_local);
  B200SSL_CHECK((reinterpret_cast<uintptr_t>(tiles) & 15) == 0 && (reinterpret_cast<uintptr_t>(params) & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(out_global) & 3) == 0 && (reinterpret_cast<uintptr_t>(out_local) & 3) == 0,
                -2, "augment: tiles must be 16-byte aligned, params / outputs 4-byte aligned");
  B200SSL_CHECK(mean != nullptr && std != nullptr && std[0] > 0.f && std[1] > 0.f && std[2] > 0.f, -2,
                "augment: mean / std (3 host floats each, std > 0) are required");
  B200SSL_CHECK((n_global == 0 || out_global != nullptr) && (n_local == 0 || out_local != nullptr), -2,
                "augment: output buffer missing");
  static bool cfg = false;
  if (!cfg) {
    B200SSL_CUDA(cudaFuncSetAttribute(multicrop_augment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AUG_SMEM_BYTES));
    cfg = true;
  }
  multicrop_augment_kernel<<<B, AUG_THREADS, AUG_SMEM_BYTES, stream>>>(
      static_cast<const uint8_t*>(tiles), static_cast<const AugParams*>(params), static_cast<__nv_bfloat16*>(out_global),
      static_cast<__nv_bfloat16*>(out_local), B, n_global, n_local, size_global, size_local, mean[0], mean[1], mean[2],
      1.f / std[0], 1.f / std[1], 1.f / std[2]);
  B200SSL_CUDA(cudaGetLastError());
  return 0;
}
