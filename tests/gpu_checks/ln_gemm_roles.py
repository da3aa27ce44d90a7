"""Developer probe: role timing of the LayerNorm-fused GEMM from the in-kernel cycle counters."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import b200ssl
from b200ssl import ops

lib = b200ssl._lib.lib()
rows = 100864
x = torch.randn(rows, 384, device="cuda")
gw, gb = torch.ones(384, device="cuda"), torch.zeros(384, device="cuda")
prof = torch.zeros(16, dtype=torch.int64, device="cuda")
for name, N, gelu in (("qkv", 1152, False), ("fc1", 1536, True)):
    w = (torch.randn(N, 384, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda")
    for _ in range(3):
        ops.ln_linear_fwd(x, gw, gb, 1e-6, w, b, gelu=gelu)
    torch.cuda.synchronize()
    prof.zero_()
    lib.b200ssl_set_gemm_prof(prof.data_ptr())
    ops.ln_linear_fwd(x, gw, gb, 1e-6, w, b, gelu=gelu)
    torch.cuda.synchronize()
    lib.b200ssl_set_gemm_prof(None)
    p = prof.tolist()
    n = max(p[3], 1)          # issuing (leader) CTAs
    n_cta = 2 * n
    print(f"{name}: issuer loop {p[2]/n:9.0f} clk: operand wait {100*p[0]/p[2]:5.1f}% accumulator wait {100*p[1]/p[2]:5.1f}% "
          f"panel wait {100*p[6]/p[2]:5.1f}% issue {100*(p[2]-p[0]-p[1]-p[6])/p[2]:5.1f}% | per CTA: transform {p[5]/n_cta:8.0f} clk, "
          f"waiting for the panel {p[7]/n_cta:8.0f} clk, epilogue warp waited for accumulators {p[4]/n_cta:8.0f} clk")
