"""Developer probe (GPU): what this B200 sustains for write-only, read-only and copy traffic with library kernels
(cudaMemset, ATen fill / copy / sum) -- the reference points for the GEMM epilogues' store-path ceiling."""
import torch


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


n = 2 * 1024 ** 3
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
af, bf = a.view(torch.float32), b.view(torch.float32)
print(f"cudaMemset      write {n / timed(lambda: a.zero_()) / 1e12:.2f} TB/s")
print(f"ATen fill fp32  write {n / timed(lambda: af.fill_(1.5)) / 1e12:.2f} TB/s")
print(f"ATen copy       r+w   {2 * n / timed(lambda: bf.copy_(af)) / 1e12:.2f} TB/s")
print(f"ATen sum fp32   read  {n / timed(lambda: af.sum()) / 1e12:.2f} TB/s")
# strided tile-like writes: [rows, 1536] bf16, write a 32-column (64-byte) slice of every row at a time
rows = 195584
t = torch.empty(rows, 1536, dtype=torch.bfloat16, device="cuda")
for w in (32, 64, 256, 1536):
    def go():
        for c in range(0, 1536, w):
            t[:, c:c + w].fill_(1.0)
    print(f"ATen fill of {w:4d}-column slices of [195584, 1536] bf16: {t.numel() * 2 / timed(go, 3) / 1e12:.2f} TB/s")
