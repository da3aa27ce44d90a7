"""Developer bench: every GEMM of one ViT-S block (fprop / dgrad / wgrad) at config-2 row counts,
B tile streamed vs stationary, through the same ops-level entry points the model uses."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import b200ssl
from b200ssl import ops

lib = b200ssl._lib.lib()
D, H4 = 384, 1536


def timeit(fn, iters=10):
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


def run(rows):
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    x, x4 = r(rows, D).bfloat16(), r(rows, H4).bfloat16()
    x3 = r(rows, 3 * D).bfloat16()
    wq, wp, w1, w2 = (r(3 * D, D) * .05).bfloat16(), (r(D, D) * .05).bfloat16(), (r(H4, D) * .05).bfloat16(), (r(D, H4) * .05).bfloat16()
    bq, bp, b1, b2 = r(3 * D), r(D), r(H4), r(D)
    res = r(rows, D)
    y32 = torch.empty(rows, D, device="cuda")
    cases = {
        "qkv fwd": (lambda: ops.linear_fwd(x, wq, bq), 2 * rows * D * 3 * D),
        "proj fwd+res32": (lambda: ops.linear_fwd(x, wp, bp, residual=res), 2 * rows * D * D),
        "fc1 fwd+gelu": (lambda: ops.linear_fwd(x, w1, b1, gelu=True), 2 * rows * D * H4),
        "fc2 fwd+res32": (lambda: ops.linear_fwd(x4, w2, b2, residual=res), 2 * rows * D * H4),
        "fc2 fwd+res32/384": (lambda: ops.gemm(x4, w2, y32, rows, D, H4, epi=ops.EPI_BIAS_RES_F32, bias=b2, aux=res,
                                                block_n=384), 2 * rows * D * H4),
        "proj dgrad": (lambda: ops.linear_dgrad(x, wp), 2 * rows * D * D),
        "qkv dgrad": (lambda: ops.linear_dgrad(x3, wq), 2 * rows * D * 3 * D),
        "fc2 dgrad*aux": (lambda: ops.linear_dgrad(x, w2, dgelu_of=x4), 2 * rows * D * H4),
        "fc1 dgrad": (lambda: ops.linear_dgrad(x4, w1), 2 * rows * D * H4),
        "proj wgrad": (lambda: ops.linear_wgrad(x, x), 2 * rows * D * D),
        "qkv wgrad": (lambda: ops.linear_wgrad(x3, x), 2 * rows * D * 3 * D),
        "fc1 wgrad": (lambda: ops.linear_wgrad(x4, x), 2 * rows * D * H4),
        "fc2 wgrad": (lambda: ops.linear_wgrad(x, x4), 2 * rows * D * H4),
    }
    modes = {"narrow": 0, "wide": 1}   # without / with the automatic 256 x 384 CTA-pair tiles
    tot = {k: 0.0 for k in modes}
    for name, (fn, fl) in cases.items():
        ts = {}
        for mode, bs in modes.items():
            lib.b200ssl_set_gemm_wide(bs)
            ts[mode] = timeit(fn)
            tot[mode] += ts[mode]
        print(f"rows={rows:7d} {name:16s} " + "   ".join(f"{m} {ts[m]:7.1f} us ({fl/ts[m]/1e6:5.0f} TF)" for m in modes))
    print(f"rows={rows:7d} TOTAL " + "   ".join(f"{m} {tot[m]:.0f} us" for m in modes))
    lib.b200ssl_set_gemm_wide(1)


run(100864)
run(94720)
