// Developer probe (GPU): write bandwidth of 16 warps per SM that each push private smem slabs to global memory --
// with TMA tensor stores of different box shapes, or by reading the slab back and storing it with coalesced st.global.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I ../../../gipmed-project-self-supervised-vit_b200/csrc \
//        -I ../../../include -o tma_store_probe tma_store_probe.cu ../../../gipmed-project-self-supervised-vit_b200/csrc/api.cu
#include "common.cuh"
#include <cstdio>
using namespace b200ssl;
extern "C" const char* b200ssl_last_error(void);

// mode 0: TMA store per slab; mode 1: lds + st.global (lane -> row l / (ROWB/16), piece l % (ROWB/16))
template <int ROWS, int ROWB, int MODE>
__global__ void __launch_bounds__(512, 1) probe(const __grid_constant__ CUtensorMap tm, uint8_t* out, long long rows, long long ld_bytes,
                                                int tw) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int SLAB = ROWS * ROWB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t slab0 = smem_u32(smem) + warp * 2 * SLAB;
  const int ntile_n = ld_bytes / tw;
  const long long ntiles = (rows / 128) * ntile_n;
  // a tile = 128 rows x tw bytes = (128 / ROWS) x (tw / ROWB) slabs, dealt round-robin to the 16 warps
  const int slabs_r = 128 / ROWS, slabs_c = tw / ROWB;
  uint32_t ring = 0;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long m = t / ntile_n, nb = t % ntile_n;
    for (int u = warp; u < slabs_r * slabs_c; u += 16) {
      const int sr = u % slabs_r, sc = u / slabs_r;
      const uint32_t slab = slab0 + (ring & 1) * SLAB;
      if (MODE == 0) {
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
      }
      // fill: every lane writes ROWS * ROWB / 32 bytes
#pragma unroll
      for (int i = 0; i < SLAB / 512; ++i) sts128(slab + (i * 32 + lane) * 16, make_uint4(lane, warp, i, 7));
      const long long row0 = m * 128 + sr * ROWS;
      const long long colb = nb * tw + sc * ROWB;
      if (MODE == 0) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d_s(&tm, slab, static_cast<int>(colb / 2), static_cast<int>(row0));
          tma_store_commit();
        }
      } else {
        __syncwarp();
        constexpr int LPR = ROWB / 16;        // lanes per row
        constexpr int RPI = 32 / LPR;         // rows per instruction
#pragma unroll
        for (int i = 0; i < ROWS / RPI; ++i) {
          const int r = i * RPI + lane / LPR;
          const uint4 v = lds128(slab + r * ROWB + (lane % LPR) * 16);
          *reinterpret_cast<uint4*>(out + (row0 + r) * ld_bytes + colb + (lane % LPR) * 16) = v;
        }
        __syncwarp();
      }
      ++ring;
    }
  }
  if (MODE == 0 && lane == 0) tma_store_wait_all<0>();
}

template <int ROWS, int ROWB, int MODE>
void run(uint8_t* buf, long long rows, long long ld, int tw, int sms) {
  CUtensorMap tm;
  uint64_t dims[2] = {static_cast<uint64_t>(ld / 2), static_cast<uint64_t>(rows)};
  uint64_t strides[2] = {2, static_cast<uint64_t>(ld)};
  uint32_t box[2] = {ROWB / 2, ROWS};
  if (make_tensor_map(&tm, buf, 2, 2, dims, strides, box, 0)) { printf("tensor map failed: %s\n", b200ssl_last_error()); return; }
  const int smem = 16 * 2 * ROWS * ROWB;
  cudaFuncSetAttribute(probe<ROWS, ROWB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) probe<ROWS, ROWB, MODE><<<sms, 512, smem>>>(tm, buf, rows, ld, tw);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) probe<ROWS, ROWB, MODE><<<sms, 512, smem>>>(tm, buf, rows, ld, tw);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("%s  slab %3d rows x %4d B, tile width %4d B : %.2f TB/s   (%s)\n", MODE == 0 ? "TMA store      " : "lds + st.global", ROWS, ROWB,
         tw, rows * ld * 5 / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const long long rows = 195584, ld = 3072;
  uint8_t* buf;
  cudaMalloc(&buf, rows * ld);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int tw : {512, 3072}) {
    run<32, 64, 0>(buf, rows, ld, tw, sms);
    run<32, 128, 0>(buf, rows, ld, tw, sms);
    run<32, 256, 0>(buf, rows, ld, tw, sms);
    run<16, 512, 0>(buf, rows, ld, tw, sms);
    run<64, 64, 0>(buf, rows, ld, tw, sms);
    run<128, 64, 0>(buf, rows, ld, tw, sms);
    run<32, 64, 1>(buf, rows, ld, tw, sms);
    run<32, 128, 1>(buf, rows, ld, tw, sms);
    run<32, 256, 1>(buf, rows, ld, tw, sms);
  }
  return 0;
}
