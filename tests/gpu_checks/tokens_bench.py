"""Developer bench: patch gather + token assembly at the config-2 crop shapes (HBM-bound byte movers)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import b200ssl
from b200ssl import ops

lib = b200ssl._lib.lib()
D, P = 384, 16


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


s = torch.cuda.current_stream().cuda_stream
for B, S in ((512, 224), (2560, 96)):
    Np = (S // P) ** 2
    img = torch.randn(B, 3, S, S, device="cuda").bfloat16()
    cols = torch.empty(B * Np, 3 * P * P, device="cuda", dtype=torch.bfloat16)
    t = timeit(lambda: lib.b200ssl_patchify(img.data_ptr(), cols.data_ptr(), B, 3, S, S, P, s))
    ref = img.view(B, 3, S // P, P, S // P, P).permute(0, 2, 4, 1, 3, 5).reshape(B * Np, -1)
    ok = torch.equal(ref, cols)
    print(f"patchify B={B} {S}^2: {t:.1f} us  {2 * img.numel() * 2 / t / 1e3:.0f} GB/s  exact={ok}")
    y = torch.randn(B * Np, D, device="cuda").bfloat16()
    cls, pos = torch.randn(D, device="cuda"), torch.randn(Np + 1, D, device="cuda")
    x = torch.empty(B * (Np + 1), D, device="cuda")
    t = timeit(lambda: lib.b200ssl_assemble_tokens(y.data_ptr(), cls.data_ptr(), pos.data_ptr(), x.data_ptr(), B, Np, D, s))
    refx = torch.cat(((cls + pos[0]).expand(B, 1, D), y.float().view(B, Np, D) + pos[1:]), 1).reshape(-1, D)
    print(f"assemble B={B} N={Np + 1}: {t:.1f} us  {(y.numel() * 2 + x.numel() * 4) / t / 1e3:.0f} GB/s  "
          f"exact={torch.equal(refx, x)}")
