"""Developer probe: where the GEMM's MMA-issuer thread spends its time (operand wait / accumulator wait / issue),
per layer shape, from the in-kernel cycle counters (b200ssl_set_gemm_prof)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import b200ssl
from b200ssl import ops

lib = b200ssl._lib.lib()
D, H4 = 384, 1536
rows = 100864
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda", generator=g)
x, x4, x3 = r(rows, D).bfloat16(), r(rows, H4).bfloat16(), r(rows, 3 * D).bfloat16()
wq, wp, w1, w2 = (r(3 * D, D) * .05).bfloat16(), (r(D, D) * .05).bfloat16(), (r(H4, D) * .05).bfloat16(), (r(D, H4) * .05).bfloat16()
bq, bp, b1, b2 = r(3 * D), r(D), r(H4), r(D)
res = r(rows, D)
cases = {
    "qkv fwd": lambda: ops.linear_fwd(x, wq, bq),
    "proj fwd+res32": lambda: ops.linear_fwd(x, wp, bp, residual=res),
    "fc1 fwd+gelu": lambda: ops.linear_fwd(x, w1, b1, gelu=True),
    "fc2 fwd+res32": lambda: ops.linear_fwd(x4, w2, b2, residual=res),
    "qkv dgrad": lambda: ops.linear_dgrad(x3, wq),
    "fc2 dgrad*aux": lambda: ops.linear_dgrad(x, w2, dgelu_of=x4),
    "fc1 dgrad": lambda: ops.linear_dgrad(x4, w1),
    "qkv wgrad": lambda: ops.linear_wgrad(x3, x),
    "fc1 wgrad": lambda: ops.linear_wgrad(x4, x),
    "fc2 wgrad": lambda: ops.linear_wgrad(x, x4),
}
prof = torch.zeros(16, dtype=torch.int64, device="cuda")
for name, fn in cases.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    prof.zero_()
    lib.b200ssl_set_gemm_prof(prof.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    lib.b200ssl_set_gemm_prof(None)
    p = prof.tolist()
    n = max(p[3], 1)
    tot = p[2] / n
    print(f"{name:16s} {e0.elapsed_time(e1)*1e3:7.1f} us | issuer loop {tot:9.0f} clk/CTA: operand wait {100*p[0]/max(p[2],1):5.1f}%  "
          f"accumulator wait {100*p[1]/max(p[2],1):5.1f}%  issue {100*(p[2]-p[0]-p[1])/max(p[2],1):5.1f}% | "
          f"epilogue warp waited {p[4]/max(n,1):9.0f} clk ({100*p[4]/max(p[2],1):5.1f}% of loop) | CTA lifetime "
          f"{p[5]/max(p[7],1):9.0f} clk, wgrad store phase {p[6]/max(p[7],1):7.0f} clk")
