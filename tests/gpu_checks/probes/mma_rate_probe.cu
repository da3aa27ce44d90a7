// Developer probe (GPU): cycles per tcgen05.mma (cta_group::1, bf16, M = 128, K = 16) for the operand layouts the
// attention kernels use, issued back to back by ONE elected thread with nothing else running on the SM.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I ../../../gipmed-project-self-supervised-vit_b200/csrc \
//        -I ../../../include -o mma_rate_probe mma_rate_probe.cu ../../../gipmed-project-self-supervised-vit_b200/csrc/api.cu
#include "common.cuh"
#include <cstdio>
using namespace b200ssl;

constexpr int TILE = 128 * 128;  // 128 rows x 64 bf16

// type: 0 S-type   A K-major,  B K-major,  N = 128      1 dQ-type  A K-major,  B MN-major, N = 64
//       2 dV-type  A MN-major, B MN-major, N = 64       3 PV-type  A in TMEM,  B MN-major, N = 64
//       4          A K-major,  B K-major,  N = 64       5 dV + dK alternating accumulators (type 2 operands)
//       6          A MN-major, B MN-major, N = 128      7          A K-major,  B MN-major, N = 128
//       8 dQ, dV, dK round-robin (the backward's issue order)
template <int type>
__global__ void __launch_bounds__(128, 1) probe(int reps, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x * 16; i < 12 * TILE; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<512>(&tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tb = __shfl_sync(0xffffffffu, tmem_slot, 0);
  if (warp == 0 && elect_one_sync()) {
    const uint32_t lo0 = (smem_u32(smem) & 0x3FFFFu) >> 4;
    const uint32_t i_s = make_idesc_bf16(128, 128, false, false), i_q = make_idesc_bf16(128, 64, false, true);
    const uint32_t i_kv = make_idesc_bf16(128, 64, true, true), i_k64 = make_idesc_bf16(128, 64, false, false);
    const uint32_t i_kv128 = make_idesc_bf16(128, 128, true, true), i_q128 = make_idesc_bf16(128, 128, false, true);
    // operand homes: A tiles at 0 / TILE, B tiles at 4 TILE / 5 TILE, P / dS chunk pairs at 8 TILE / 10 TILE
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        switch (type) {
          case 0: umma_bf16_ss(tb, sw128_desc_at(lo0, (j & 3) * 32, 16, 1024), sw128_desc_at(lo0, 4 * TILE + (j & 3) * 32, 16, 1024), i_s, 1); break;
          case 1: umma_bf16_ss(tb + 256, sw128_desc_at(lo0, 10 * TILE + (j >> 2) * TILE + (j & 3) * 32, 16, 1024),
                               sw128_desc_at(lo0, 4 * TILE + j * 2048, 8192, 1024), i_q, 1); break;
          case 2: umma_bf16_ss(tb + 448, sw128_desc_at(lo0, 8 * TILE + j * 2048, TILE, 1024),
                               sw128_desc_at(lo0, TILE + j * 2048, 8192, 1024), i_kv, 1); break;
          case 3: umma_bf16_ts(tb + 192, tb + 8 * j, sw128_desc_at(lo0, 4 * TILE + j * 2048, 8192, 1024), i_q, 1); break;
          case 4: umma_bf16_ss(tb, sw128_desc_at(lo0, (j & 3) * 32, 16, 1024), sw128_desc_at(lo0, 4 * TILE + (j & 3) * 32, 16, 1024), i_k64, 1); break;
          case 5:
            if (j & 1) umma_bf16_ss(tb + 448, sw128_desc_at(lo0, 8 * TILE + (j >> 1) * 2048, TILE, 1024),
                                    sw128_desc_at(lo0, TILE + (j >> 1) * 2048, 8192, 1024), i_kv, 1);
            else umma_bf16_ss(tb + 384, sw128_desc_at(lo0, 10 * TILE + (j >> 1) * 2048, TILE, 1024),
                              sw128_desc_at(lo0, (j >> 1) * 2048, 8192, 1024), i_kv, 1);
            break;
          case 6: umma_bf16_ss(tb, sw128_desc_at(lo0, 8 * TILE + j * 2048, TILE, 1024),
                               sw128_desc_at(lo0, 4 * TILE + j * 2048, TILE, 1024), i_kv128, 1); break;
          case 7: umma_bf16_ss(tb, sw128_desc_at(lo0, 10 * TILE + (j >> 2) * TILE + (j & 3) * 32, 16, 1024),
                               sw128_desc_at(lo0, 4 * TILE + j * 2048, TILE, 1024), i_q128, 1); break;
          default:
            umma_bf16_ss(tb + 256, sw128_desc_at(lo0, 10 * TILE + (j >> 2) * TILE + (j & 3) * 32, 16, 1024),
                         sw128_desc_at(lo0, 4 * TILE + j * 2048, 8192, 1024), i_q, 1);
            umma_bf16_ss(tb + 448, sw128_desc_at(lo0, 8 * TILE + j * 2048, TILE, 1024),
                         sw128_desc_at(lo0, TILE + j * 2048, 8192, 1024), i_kv, 1);
            umma_bf16_ss(tb + 384, sw128_desc_at(lo0, 10 * TILE + j * 2048, TILE, 1024),
                         sw128_desc_at(lo0, j * 2048, 8192, 1024), i_kv, 1);
            break;
        }
      }
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tb);
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 16);
  const int smem = 12 * TILE + 1024;
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"S-type  A K-major  x B K-major   N=128", "dQ-type A K-major  x B MN-major  N=64 ", "dV-type A MN-major x B MN-major  N=64 ",
                         "PV-type A in TMEM  x B MN-major  N=64 ", "        A K-major  x B K-major   N=64 ", "dV,dK alternating accumulators   N=64 ",
                         "        A MN-major x B MN-major  N=128", "        A K-major  x B MN-major  N=128", "dQ,dV,dK round-robin (x3 MMAs)   N=64 "};
  for (int grid : {1, 148})
    for (int type = 0; type < 9; ++type) {
      const int reps = 32;
      for (int it = 0; it < 2; ++it) {   // warm, then measured
        switch (type) {
          case 0: probe<0><<<grid, 128, smem>>>(reps, d); break;
          case 1: probe<1><<<grid, 128, smem>>>(reps, d); break;
          case 2: probe<2><<<grid, 128, smem>>>(reps, d); break;
          case 3: probe<3><<<grid, 128, smem>>>(reps, d); break;
          case 4: probe<4><<<grid, 128, smem>>>(reps, d); break;
          case 5: probe<5><<<grid, 128, smem>>>(reps, d); break;
          case 6: probe<6><<<grid, 128, smem>>>(reps, d); break;
          case 7: probe<7><<<grid, 128, smem>>>(reps, d); break;
          default: probe<8><<<grid, 128, smem>>>(reps, d); break;
        }
      }
      unsigned long long h[2];
      cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("type %d: %s\n", type, cudaGetErrorString(e)); return 1; }
      const int n = reps * 8 * (type == 8 ? 3 : 1);
      printf("grid %3d  %s : %6.1f clk / MMA issued, %6.1f clk / MMA until the last one retired (%d MMAs)\n", grid, names[type],
             double(h[0]) / n, double(h[1]) / n, n);
    }
  return 0;
}
