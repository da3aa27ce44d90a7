"""Developer A/B: tile choices for the small (N = 384, K = 384) projection GEMMs at the packed student row count."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from b200ssl import ops

D = 384


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


for rows in (195584, 100864):
    g = torch.Generator(device="cuda").manual_seed(0)
    dy = torch.randn(rows, D, device="cuda", generator=g).bfloat16()
    x = torch.randn(rows, D, device="cuda", generator=g).bfloat16()
    w = (torch.randn(D, D, device="cuda", generator=g) * 0.05).bfloat16()
    dx = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(D, D, device="cuda")
    db = torch.zeros(D, device="cuda")
    ref_dx = ref_dw = None
    for bn in (0, 128, 192, 384):
        try:
            t = timeit(lambda: ops.gemm(dy, w, dx, rows, D, D, b_mn=True, epi=ops.EPI_BIAS, block_n=bn))
            if ref_dx is None:
                ref_dx = dx.clone()
            err = ((dx.float() - ref_dx.float()).norm() / ref_dx.float().norm()).item()
            print(f"rows={rows} proj dgrad block_n={bn:3d}: {t:7.1f} us  ({2*rows*D*D/t/1e6:5.0f} TF, {2*rows*D*2/t/1e3:5.0f} GB/s)  diff {err:.1e}")
        except RuntimeError as e:
            print(f"rows={rows} proj dgrad block_n={bn}: {e}")
    for bn in (0, 128, 192, 384):
        try:
            def run():
                ops.gemm(dy, x, dw, D, D, rows, a_mn=True, b_mn=True, epi=ops.EPI_ATOMIC_F32, split_k=0, bias=db, block_n=bn)
            t = timeit(run)
            dw.zero_(); db.zero_(); run(); torch.cuda.synchronize()
            if ref_dw is None:
                ref_dw = dw.clone()
            err = ((dw - ref_dw).norm() / ref_dw.norm()).item()
            print(f"rows={rows} proj wgrad block_n={bn:3d}: {t:7.1f} us  ({2*rows*D*D/t/1e6:5.0f} TF)  diff {err:.1e}")
        except RuntimeError as e:
            print(f"rows={rows} proj wgrad block_n={bn}: {e}")
