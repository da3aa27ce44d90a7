"""Developer bench (GPU): DINO loss forward / backward at the config-2 shape (B = 256, 12 crops, K = 65,536)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import b200ssl
from b200ssl import ops

B, NC, K = 256, 12, 65536
g = torch.Generator(device="cuda").manual_seed(0)
s = torch.randn(NC * B, K, device="cuda", generator=g).bfloat16().requires_grad_(True)
t = torch.randn(2 * B, K, device="cuda", generator=g).bfloat16()
c = torch.randn(K, device="cuda", generator=g) * 0.1
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timed(fn, reps=10):
    tot = 0.0
    for i in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


loss = ops.DinoLossFn.apply(s, t, c, NC, 0.1, 0.04)
print("loss", loss.item())
us = timed(lambda: ops.DinoLossFn.apply(s, t, c, NC, 0.1, 0.04))
print(f"forward  {us:7.1f} us  {(NC + 2) * B * K * 2 / us / 1e6:6.2f} TB/s (470 MB)")
loss = ops.DinoLossFn.apply(s, t, c, NC, 0.1, 0.04)
us = timed(lambda: torch.autograd.grad(loss, s, retain_graph=True))
print(f"backward {us:7.1f} us  {(2 * NC + 2) * B * K * 2 / us / 1e6:6.2f} TB/s (872 MB)")
