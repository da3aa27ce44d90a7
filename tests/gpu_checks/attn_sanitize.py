"""Developer helper for compute-sanitizer: one forward + backward of the fused attention at a few small shapes that cover
the packed (N <= 64), single-tile, two-tile, block-pair and streaming / paired paths."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from b200ssl import ops

for B, N, H in ((2, 128, 1), (7, 37, 3), (3, 64, 2), (2, 197, 2), (3, 257, 2), (2, 785, 1), (5, 100, 2)):
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda").bfloat16()
    dout = torch.randn(B * N, H * 64, device="cuda").bfloat16()
    out, lse2 = ops.attention_fwd(qkv, B, N, H, 0.125)
    dqkv = ops.attention_bwd(qkv, out, dout, lse2, B, N, H, 0.125)
    torch.cuda.synchronize()
    assert torch.isfinite(dqkv.float()).all() and torch.isfinite(out.float()).all(), (B, N, H)
    print("ok", B, N, H)
print("done")
