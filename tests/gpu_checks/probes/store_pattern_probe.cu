// Developer probe (GPU): write-only bandwidth of ONE persistent 512/1024-thread CTA per SM for different store patterns
// over a [rows, ld] bf16 matrix -- what bounds the GEMM epilogues' output path (4.4 - 4.6 TB/s measured there)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_pattern_probe store_pattern_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// pattern 0: linear fill, CTA-strided 16-byte stores
// pattern 1: tiles of 128 rows x `tw` bytes; a warp instruction writes 512 contiguous bytes of one row (tw >= 512)
// pattern 2: tiles of 128 rows x `tw` bytes; a warp instruction writes 8 rows x 64 bytes (lane -> row l/4, piece l%4)
// pattern 3: tiles of 128 rows x `tw` bytes; a warp instruction writes 4 rows x 128 bytes
// pattern 4: tiles of 128 rows x `tw` bytes; a warp instruction writes 2 rows x 256 bytes
__global__ void __launch_bounds__(1024, 1) probe(uint4* out, long long rows, long long ld_bytes, int tw, int pattern) {
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  if (pattern == 0) {
    const long long n = rows * ld_bytes / 16;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = v;
    return;
  }
  const int ntile_n = ld_bytes / tw;
  const long long ntiles = (rows / 128) * ntile_n;
  const int seg = pattern == 1 ? 512 : pattern == 2 ? 64 : pattern == 3 ? 128 : 256;   // contiguous bytes per row per instruction
  const int rows_per_instr = 512 / seg;
  const int lanes_per_row = seg / 16;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    // n fastest across CTAs at a given time: all column tiles of a row block are written at about the same time
    const long long m = t / ntile_n, nb = t % ntile_n;
    char* base = reinterpret_cast<char*>(out) + m * 128 * ld_bytes + nb * tw;
    const int col_steps = tw / seg;
    const int units = (128 / rows_per_instr) * col_steps;   // warp instructions per tile
    for (int u = warp; u < units; u += nwarps) {
      const int rblk = u / col_steps, cs = u % col_steps;
      const int r = rblk * rows_per_instr + lane / lanes_per_row;
      *reinterpret_cast<uint4*>(base + (long long)r * ld_bytes + cs * seg + (lane % lanes_per_row) * 16) = v;
    }
  }
}

int main() {
  const long long rows = 195584, ld = 3072;
  uint4* buf;
  cudaMalloc(&buf, rows * ld);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int pats[] = {0, 1, 4, 3, 2};
  for (int threads : {512, 1024})
    for (int grid_mult : {1, 2})
      for (int pi = 0; pi < 5; ++pi)
        for (int tw : {512, 3072}) {
          const int pat = pats[pi];
          if (pat == 0 && tw != 512) continue;
          if (grid_mult == 2 && threads == 1024) continue;
          for (int i = 0; i < 2; ++i) probe<<<sms * grid_mult, threads>>>(buf, rows, ld, tw, pat);
          cudaEventRecord(e0);
          for (int i = 0; i < 5; ++i) probe<<<sms * grid_mult, threads>>>(buf, rows, ld, tw, pat);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          printf("threads %4d x %d CTA/SM  pattern %d  tile width %4d B : %.2f TB/s\n", threads, grid_mult, pat, tw,
                 rows * ld * 5 / (ms * 1e-3) / 1e12);
        }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
