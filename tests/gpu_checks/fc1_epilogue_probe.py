"""Developer probe (GPU): what the GELU arithmetic costs in the fc1 epilogue. Same GEMM (rows x 384 -> 1536, B-stationary
pair tiles) with three epilogues -- plain bias (one bf16 output), GELU forward only (teacher; one output), GELU + GELU'
(student; two outputs) -- at the student and teacher row counts, with the per-chunk phase counters of one epilogue thread."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import b200ssl
from b200ssl import ops

lib = b200ssl._lib.lib()
names = ["wait slab", "TMEM load", "bias/aux", "math+pack+sts", "proxy fence", "store issue"]


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


for rows in (195584, 100864):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(rows, 384, device="cuda", generator=g).bfloat16()
    w = (torch.randn(1536, 384, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(1536, device="cuda", generator=g)
    cases = {
        "plain bias      ": lambda: ops.linear_fwd(x, w, b),
        "GELU fwd only   ": lambda: ops.linear_fwd(x, w, b, gelu="fwd_only"),
        "GELU + GELU'    ": lambda: ops.linear_fwd(x, w, b, gelu=True),
    }
    for name, fn in cases.items():
        t = timeit(fn)
        prof = torch.zeros(16, dtype=torch.int64, device="cuda")
        lib.b200ssl_set_gemm_prof(prof.data_ptr())
        fn()
        torch.cuda.synchronize()
        lib.b200ssl_set_gemm_prof(None)
        p = prof.tolist()
        n = max(p[15], 1)
        nt = max(p[3], 1)
        print(f"rows {rows} {name} {t:6.1f} us {2*rows*384*1536/t/1e6:6.0f} TF/s | {sum(p[8:14])/n:6.0f} clk/chunk  " +
              "  ".join(f"{nm} {p[8+i]/n:5.0f}" for i, nm in enumerate(names)) +
              f" | epilogue waited for the accumulator {p[4]/nt:6.0f} clk/tile, issuer: operand wait {100*p[0]/max(p[2],1):4.1f}% accumulator wait {100*p[1]/max(p[2],1):4.1f}%")
