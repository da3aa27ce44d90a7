"""Developer helper for ncu: one warm + one measured launch of the main GEMM variants at config-2 shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from b200ssl import ops

rows = 100864
g = torch.Generator(device="cuda").manual_seed(0)
x384 = torch.randn(rows, 384, device="cuda", generator=g).bfloat16()
x1536 = torch.randn(rows, 1536, device="cuda", generator=g).bfloat16()
w_fc1 = (torch.randn(1536, 384, device="cuda", generator=g) * 0.05).bfloat16()
w_qkv = (torch.randn(1152, 384, device="cuda", generator=g) * 0.05).bfloat16()
w_fc2 = (torch.randn(384, 1536, device="cuda", generator=g) * 0.05).bfloat16()
w_proj = (torch.randn(384, 384, device="cuda", generator=g) * 0.05).bfloat16()
b1536 = torch.randn(1536, device="cuda", generator=g)
b1152 = torch.randn(1152, device="cuda", generator=g)
b384 = torch.randn(384, device="cuda", generator=g)
res32 = torch.randn(rows, 384, device="cuda", generator=g)
for _ in range(2):
    ops.linear_fwd(x384, w_fc1, b1536, gelu=True)          # EPI 1
    ops.linear_fwd(x384, w_qkv, b1152)                     # EPI 0
    ops.linear_fwd(x1536, w_fc2, b384, residual=res32)     # EPI 5
    ops.linear_dgrad(x384, w_fc2, dgelu_of=x1536)          # EPI 3 (dy [rows,384] @ W2 [384,1536]) * aux
    ops.linear_wgrad(x1536, x384)                          # EPI 4
    ops.linear_fwd(x384, w_proj, b384, residual=res32)     # EPI 5, K = 384 (proj)
    ops.linear_wgrad(x384, x384)                           # EPI 4, proj wgrad (384 x 384 output)
torch.cuda.synchronize()
print("done")
