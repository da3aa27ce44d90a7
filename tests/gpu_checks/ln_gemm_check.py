"""Developer check (GPU): the LayerNorm-fused GEMM (b200ssl_ln_gemm) vs LayerNorm + matmul in fp32 torch, and
its timing against the separate LayerNorm kernel + GEMM it replaces."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F

from b200ssl import ops


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


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


def check(rows, N, gelu, keep, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(rows, 384, device="cuda", generator=g) * 2.0 + torch.randn(rows, 1, device="cuda", generator=g)
    gw = 1.0 + 0.1 * torch.randn(384, device="cuda", generator=g)
    gb = 0.1 * torch.randn(384, device="cuda", generator=g)
    w = (torch.randn(N, 384, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    out, ln, mean, rstd = ops.ln_linear_fwd(x, gw, gb, 1e-6, w, b, gelu=gelu, keep=keep)
    torch.cuda.synchronize()
    ln_ref = F.layer_norm(x, (384,), gw, gb, 1e-6)
    pre_ref = ln_ref.bfloat16().float() @ w.float().t() + b
    errs = {}
    if gelu:
        pr = pre_ref.clone().requires_grad_(True)
        F.gelu(pr).sum().backward()
        errs["gelu"] = rel(out[1], F.gelu(pre_ref))
        if gelu is True:
            errs["gelu'"] = rel(out[0], pr.grad)
    else:
        errs["y"] = rel(out, pre_ref)
    if keep:
        errs["ln"] = rel(ln, ln_ref)
        errs["mean"] = rel(mean, x.mean(1))
        errs["rstd"] = rel(rstd, 1.0 / torch.sqrt(x.var(1, unbiased=False) + 1e-6))
    nan = sum(int(torch.isnan(t.float()).sum()) for t in (out if isinstance(out, tuple) else (out,)) if t is not None)
    ok = max(errs.values()) < 1e-2 and nan == 0
    print(f"[{'ok' if ok else 'FAIL'}] rows={rows} N={N} gelu={gelu} keep={keep}: " +
          " ".join(f"{k} {v:.2e}" for k, v in errs.items()) + f" nan={nan}")
    return ok


ok = True
ok &= check(300, 1152, False, True, 1)
ok &= check(129, 1536, True, True, 2)
ok &= check(128 * 9 + 77, 1152, False, False, 3)
ok &= check(128 * 160 + 5, 1536, "fwd_only", False, 4)
ok &= check(128 * 301, 1536, True, True, 5)
ok &= check(100864, 1152, False, True, 6)
print("ALL OK" if ok else "SOME FAILED")
if ok and "--bench" in sys.argv:
    for rows in (100864, 94720):
        x = torch.randn(rows, 384, device="cuda")
        gw, gb = torch.ones(384, device="cuda"), torch.zeros(384, device="cuda")
        for name, N, gelu in (("qkv", 1152, False), ("fc1", 1536, True), ("fc1 (no-grad)", 1536, "fwd_only")):
            w = (torch.randn(N, 384, device="cuda") * 0.05).bfloat16()
            b = torch.randn(N, device="cuda")
            keep = gelu != "fwd_only"
            t_f = timeit(lambda: ops.ln_linear_fwd(x, gw, gb, 1e-6, w, b, gelu=gelu, keep=keep))
            t_ln = timeit(lambda: ops.layernorm_fwd(x, gw, gb, 1e-6))
            ln = ops.layernorm_fwd(x, gw, gb, 1e-6)[0]
            t_g = timeit(lambda: ops.linear_fwd(ln, w, b, gelu=gelu))
            print(f"rows={rows} {name:14s}: fused {t_f:6.1f} us | LayerNorm {t_ln:5.1f} + GEMM {t_g:6.1f} = {t_ln + t_g:6.1f} us")
sys.exit(0 if ok else 1)
