"""Developer probe: write bandwidth of the GEMM epilogue's store path (K = 64: the mainloop is negligible, the kernel
is a 128 x BN tile writer) vs a plain full-row streaming kernel, to see what the 32-row x 64-byte TMA-store boxes cost."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from b200ssl import ops

rows = 100864


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


x = torch.randn(rows, 64, device="cuda").bfloat16()
for N in (1536, 384):
    w = torch.randn(N, 64, device="cuda").bfloat16()
    y = torch.empty(rows, N, device="cuda", dtype=torch.bfloat16)
    for bn in (256, 192, 128) if N == 1536 else (192, 128):
        t = timeit(lambda: ops.gemm(x, w, y, rows, N, 64, epi=ops.EPI_BIAS, block_n=bn))
        print(f"bf16 out N={N} bn={bn}: {t:6.1f} us  {rows*N*2/t/1e3:6.0f} GB/s written")
    res = torch.randn(rows, N, device="cuda")
    y32 = torch.empty(rows, N, device="cuda")
    t = timeit(lambda: ops.gemm(x, w, y32, rows, N, 64, epi=ops.EPI_BIAS_RES_F32, aux=res))
    print(f"fp32 residual in/out N={N}: {t:6.1f} us  {rows*N*8/t/1e3:6.0f} GB/s read+written")
src = torch.randn(rows, 1536, device="cuda")
dst = torch.empty(rows, 1536, device="cuda", dtype=torch.bfloat16)
t = timeit(lambda: ops.cast_f32_to_bf16(src, dst))
print(f"streaming cast fp32->bf16 (full rows): {t:6.1f} us  {rows*1536*6/t/1e3:6.0f} GB/s")
a, b = torch.empty(rows, 1536, device="cuda"), torch.empty(rows, 1536, device="cuda")
t = timeit(lambda: b.copy_(a))
print(f"torch copy fp32: {t:6.1f} us  {rows*1536*8/t/1e3:6.0f} GB/s")

# per-chunk phases of one epilogue thread in the tile-writer setting
import b200ssl
lib = b200ssl._lib.lib()
names = ["wait slab", "TMEM load", "bias/aux", "pack+sts", "proxy fence", "store issue"]
for N, bn, epi, tag in ((1536, 256, ops.EPI_BIAS, "bf16 out"), (384, 192, ops.EPI_BIAS_RES_F32, "fp32 res")):
    w = torch.randn(N, 64, device="cuda").bfloat16()
    y = torch.empty(rows, N, device="cuda", dtype=torch.float32 if epi == ops.EPI_BIAS_RES_F32 else torch.bfloat16)
    res = torch.randn(rows, N, device="cuda") if epi == ops.EPI_BIAS_RES_F32 else None
    bias = torch.randn(N, device="cuda")
    fn = lambda: ops.gemm(x, w, y, rows, N, 64, epi=epi, aux=res, bias=bias, block_n=bn)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    prof = torch.zeros(16, dtype=torch.int64, device="cuda")
    lib.b200ssl_set_gemm_prof(prof.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.b200ssl_set_gemm_prof(None)
    p = prof.tolist()
    n = max(p[15], 1)
    print(f"{tag} bn={bn}: {sum(p[8:14])/n:6.0f} clk/chunk  " + "  ".join(f"{nm} {p[8+i]/n:5.0f}" for i, nm in enumerate(names)))

# is a pure write stream limited below the copy bandwidth?
buf = torch.empty(rows * 1536, device="cuda", dtype=torch.bfloat16)
t = timeit(lambda: buf.fill_(1.0))
print(f"torch fill bf16 (pure write, {buf.numel()*2/1e6:.0f} MB): {t:6.1f} us  {buf.numel()*2/t/1e3:6.0f} GB/s")
big = torch.empty(rows * 1536 * 2, device="cuda", dtype=torch.float32)
t = timeit(lambda: big.fill_(1.0))
print(f"torch fill fp32 (pure write, {big.numel()*4/1e6:.0f} MB): {t:6.1f} us  {big.numel()*4/t/1e3:6.0f} GB/s")
t = timeit(lambda: big.sum())
print(f"torch sum fp32 (pure read, {big.numel()*4/1e6:.0f} MB): {t:6.1f} us  {big.numel()*4/t/1e3:6.0f} GB/s")
