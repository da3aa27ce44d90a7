"""Developer bench: what would running the student's global-crop rows (100,864) and local-crop rows (94,720) through
ONE launch per row-wise kernel (GEMMs, LayerNorm) save over today's two launches? The op sequence of one ViT-S block
(forward + backward, attention left out: it stays per crop group) is captured in a CUDA graph, once with the two row
groups in separate launches and once on the concatenated rows, and replayed."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import b200ssl  # noqa: F401
from b200ssl import ops

D, H4 = 384, 1536
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda", generator=g)
wq, wp, w1, w2 = (r(3 * D, D) * .05).bfloat16(), (r(D, D) * .05).bfloat16(), (r(H4, D) * .05).bfloat16(), (r(D, H4) * .05).bfloat16()
bq, bp, b1, b2 = r(3 * D), r(D), r(H4), r(D)
lw, lb = r(D), r(D)


def make(rows):
    return dict(rows=rows, x32=r(rows, D), x=r(rows, D).bfloat16(), x3=r(rows, 3 * D).bfloat16(),
                x4=r(rows, H4).bfloat16(), dy=r(rows, D).bfloat16())


def block(t):
    """row-wise kernels of one block: fwd (LN, qkv, proj+res, LN, fc1+gelu, fc2+res) and bwd (dgrads, wgrads, LN bwd)"""
    y, mean, rstd = ops.layernorm_fwd(t["x32"], lw, lb, 1e-6)
    ops.linear_fwd(y, wq, bq)
    ops.linear_fwd(t["x"], wp, bp, residual=t["x32"])
    y2, mean2, rstd2 = ops.layernorm_fwd(t["x32"], lw, lb, 1e-6)
    ops.linear_fwd(y2, w1, b1, gelu=True)
    ops.linear_fwd(t["x4"], w2, b2, residual=t["x32"])
    # backward
    ops.linear_dgrad(t["dy"], w2, dgelu_of=t["x4"])
    ops.linear_wgrad(t["dy"], t["x4"])
    ops.linear_dgrad(t["x4"], w1)
    ops.linear_wgrad(t["x4"], t["x"])
    ops.layernorm_bwd(t["x32"], t["dy"], lw, mean2, rstd2, dres=t["dy"])
    ops.linear_dgrad(t["dy"], wp)
    ops.linear_wgrad(t["dy"], t["x"])
    ops.linear_dgrad(t["x3"], wq)
    ops.linear_wgrad(t["x3"], t["x"])
    ops.layernorm_bwd(t["x32"], t["dy"], lw, mean, rstd, dres=t["dy"])


def graph_time(fn, reps=12, iters=5):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    for _ in range(2):
        gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / reps


a, b, ab = make(100864), make(94720), make(100864 + 94720)
t_split = graph_time(lambda: (block(a), block(b)))
t_merged = graph_time(lambda: block(ab))
print(f"one block, row-wise kernels: split {t_split*1e3:.1f} us   merged {t_merged*1e3:.1f} us   "
      f"saving {(t_split - t_merged)*1e3:.1f} us/block = {(t_split - t_merged) * 12:.3f} ms per 12-block student pass")
