"""Developer probe: where one softmax thread of the attention-forward kernel spends its cycles, per phase."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import b200ssl
from b200ssl import ops

lib = b200ssl._lib.lib()
names = ["wait S", "max pass", "barrier", "exp pass", "barrier+lse", "wait O", "epilogue"]
for B, N, H in ((512, 197, 6), (2560, 37, 6)):
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
    for s in range(2):
        n = max(p[s][7], 1)
        tot = sum(p[s][:7])
        print(f"  slot {s}: {tot/n:7.0f} clk/tile  " + "  ".join(f"{nm} {p[s][i]/n:6.0f}" for i, nm in enumerate(names)))

# [4..6] were the math warps' waits for the dK/dV MMAs and their dK/dV/dQ stores: the auxiliary group drains the accumulators now
bnames = ["wait rowc", "wait S/dP", "math", "wait prev MMA + smem", "-", "-", "-", "turnaround"]
for B, N, H in ((512, 197, 6), (2560, 37, 6)):
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda").bfloat16()
    dout = torch.randn(B * N, H * 64, device="cuda").bfloat16()
    out, lse2 = ops.attention_fwd(qkv, B, N, H, 0.125)
    for _ in range(3):
        ops.attention_bwd(qkv, out, dout, lse2, B, N, H, 0.125)
    torch.cuda.synchronize()
    prof = torch.zeros(128, dtype=torch.int64, device="cuda")
    lib.b200ssl_set_attn_prof(prof.data_ptr())
    ops.attention_bwd(qkv, out, dout, lse2, B, N, H, 0.125)
    torch.cuda.synchronize()
    lib.b200ssl_set_attn_prof(None)
    p = prof.tolist()
    n = max(p[8], 1)
    # math-side counters need a library built with -DB200SSL_ATTN_BWD_PROF (they cost the math warps registers); zeros otherwise
    print(f"bwd B={B} N={N} H={H}: {sum(p[:8])/n:7.0f} clk/item  " + "  ".join(f"{nm} {p[i]/n:6.0f}" for i, nm in enumerate(bnames) if nm != "-"))
    cn = ["wait S/dP taken", "wait loads", "wait P/dS", "wait drain", "issue S/dP", "wait last MMAs", "issue dQ/dK/dV"]
    # the issuing thread of EVERY CTA adds to these (the math counters come from one thread per CTA as well): per item
    print(f"    issuer: {sum(p[9:16])/n:7.0f} clk/item  " + "  ".join(f"{nm} {p[9+i]/n:6.0f}" for i, nm in enumerate(cn)))
