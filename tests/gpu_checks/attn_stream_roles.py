"""Developer probe: where the first thread of each softmax team of the STREAMING attention forward spends its cycles."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import b200ssl
from b200ssl import ops

lib = b200ssl._lib.lib()
lib.b200ssl_set_attn_stream(1)
names = ["wait S", "max pass", "wait O", "fold O", "finalize", "exp pass"]
shapes = [tuple(int(a) for a in sys.argv[1:4])] if len(sys.argv) >= 4 else [(64, 785, 6), (512, 197, 6), (256, 257, 6)]
for B, N, H in shapes:
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda").bfloat16()
    for _ in range(3):
        ops.attention_fwd(qkv, B, N, H, 0.125)
    torch.cuda.synchronize()
    prof = torch.zeros(16, dtype=torch.int64, device="cuda")
    lib.b200ssl_set_attn_prof(prof.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.attention_fwd(qkv, B, N, H, 0.125)
    e1.record()
    torch.cuda.synchronize()
    lib.b200ssl_set_attn_prof(None)
    p = prof.view(2, 8).tolist()
    print(f"B={B} N={N} H={H}: {e0.elapsed_time(e1)*1e3:.1f} us")
    for t in range(2):
        n = max(p[t][6], 1)
        print(f"  team {t}: {p[t][6]} blocks, {p[t][7]/n:7.0f} clk/block  " +
              "  ".join(f"{nm} {p[t][i]/n:6.0f}" for i, nm in enumerate(names)))
