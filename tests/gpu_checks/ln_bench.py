"""Developer bench: LayerNorm fwd/bwd at config-2 row counts vs the HBM roofline."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from b200ssl import ops

rows, D = 100864, 384
x = torch.randn(rows, D, device="cuda")
dy = torch.randn(rows, D, device="cuda").bfloat16()
dres = torch.randn(rows, D, device="cuda").bfloat16()
w, b = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
y, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-6)


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


import b200ssl
lib = b200ssl._lib.lib()
for rows_ in (100864, 195584, 94720 + 7):
    x_ = torch.randn(rows_, D, device="cuda")
    dy_, dres_ = torch.randn(rows_, D, device="cuda").bfloat16(), torch.randn(rows_, D, device="cuda").bfloat16()
    _, mean_, rstd_ = ops.layernorm_fwd(x_, w, b, 1e-6)
    outs = {}
    for v in (0, 1, 0, 1):
        lib.b200ssl_set_ln_bwd_staged(v)
        outs[v] = ops.layernorm_bwd(x_, dy_, w, mean_, rstd_, dres=dres_)
        t = timeit(lambda: ops.layernorm_bwd(x_, dy_, w, mean_, rstd_, dres=dres_))
        print(f"rows {rows_} staged {v}: ln bwd {t:.1f} us  {rows_*D*10/t/1e3:.0f} GB/s")
    print("   dx equal:", torch.equal(outs[0][0], outs[1][0]), " dgamma rel:",
          ((outs[0][1] - outs[1][1]).norm() / outs[0][1].norm()).item())
lib.b200ssl_set_ln_bwd_staged(1)
t_f = timeit(lambda: ops.layernorm_fwd(x, w, b, 1e-6))
t_b = timeit(lambda: ops.layernorm_bwd(x, dy, w, mean, rstd, dres=dres))
print(f"ln fwd {t_f:.1f} us  {rows*D*6/t_f/1e3:.0f} GB/s | ln bwd {t_b:.1f} us  {rows*D*10/t_b/1e3:.0f} GB/s (algorithmic bytes)")
