"""Developer A/B (GPU): two-tile attention forward with round-locked vs event-driven (de-phased) MMA issue
(b200ssl_set_attn_dephase). Same bits expected from both; time per launch and the phase counters of slot 0 / slot 1."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import b200ssl
from b200ssl import ops

lib = b200ssl._lib.lib()
MODES = (0, 1, 40, 100, 200)
names = ["wait S", "max pass", "barrier", "exp pass", "barrier+lse", "wait O", "epilogue"]


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


shapes = [(512, 197, 6), (2560, 37, 6), (256, 197, 6), (5, 37, 3), (7, 197, 1), (64, 256, 6), (33, 129, 2), (300, 100, 6),
          (1, 197, 1), (3, 64, 2), (128, 197, 12)]
for B, N, H in shapes:
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + N)
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda", generator=g).bfloat16()
    res = {}
    for mode in MODES:
        lib.b200ssl_set_attn_dephase(mode)
        out, lse2 = ops.attention_fwd(qkv, B, N, H, 0.125)
        torch.cuda.synchronize()
        us = timed(lambda: ops.attention_fwd(qkv, B, N, H, 0.125))
        prof = torch.zeros(16, dtype=torch.int64, device="cuda")
        lib.b200ssl_set_attn_prof(prof.data_ptr())
        ops.attention_fwd(qkv, B, N, H, 0.125)
        torch.cuda.synchronize()
        lib.b200ssl_set_attn_prof(None)
        res[mode] = (out.clone(), lse2.clone(), us, prof.view(2, 8).tolist())
    same = all(torch.equal(res[0][0], res[m][0]) and torch.equal(res[0][1], res[m][1]) for m in MODES)
    print(f"[{'ok' if same else 'FAIL'}] B={B} N={N} H={H}: " + ", ".join(f"mode {m} {res[m][2]:.1f} us" for m in MODES))
    if B >= 512:
        for mode in MODES:
            p = res[mode][3]
            for s in range(2):
                n = max(p[s][7], 1)
                print(f"    mode {mode} slot {s}: {sum(p[s][:7]) / n:7.0f} clk/tile  " +
                      "  ".join(f"{nm} {p[s][i] / n:6.0f}" for i, nm in enumerate(names)))
lib.b200ssl_set_attn_dephase(1)
