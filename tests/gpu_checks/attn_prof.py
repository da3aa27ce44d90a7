"""Developer helper: run the fused attention fwd/bwd once warm + once measured (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from b200ssl import ops

B, N, H = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (512, 197, 6)
qkv = torch.randn(B * N, 3 * H * 64, device="cuda").bfloat16()
dout = torch.randn(B * N, H * 64, device="cuda").bfloat16()
for _ in range(2):
    out, lse2 = ops.attention_fwd(qkv, B, N, H, 0.125)
    dqkv = ops.attention_bwd(qkv, out, dout, lse2, B, N, H, 0.125)
torch.cuda.synchronize()
print("done")
