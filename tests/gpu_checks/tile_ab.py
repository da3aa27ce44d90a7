"""Developer A/B: alternative tile widths for the K = 384 GEMMs of a ViT-S block (same problem, explicit block_n)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from b200ssl import ops

rows, D, H4 = 100864, 384, 1536
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda", generator=g)
x, x4 = r(rows, D).bfloat16(), r(rows, H4).bfloat16()
wq, w1, w2 = (r(3 * D, D) * .05).bfloat16(), (r(H4, D) * .05).bfloat16(), (r(D, H4) * .05).bfloat16()
bq, b1 = r(3 * D), r(H4)
y3 = torch.empty(rows, 3 * D, device="cuda", dtype=torch.bfloat16)
y4a, y4b = torch.empty(rows, H4, device="cuda", dtype=torch.bfloat16), torch.empty(rows, H4, device="cuda", dtype=torch.bfloat16)


def timeit(fn, iters=20):
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


cases = []
for bn in (192, 384, 128):
    cases.append((f"qkv fwd bn={bn}", lambda bn=bn: ops.gemm(x, wq, y3, rows, 3 * D, D, epi=ops.EPI_BIAS, bias=bq, block_n=bn)))
for bn in (256, 192, 128):
    cases.append((f"fc1 fwd+gelu bn={bn}", lambda bn=bn: ops.gemm(x, w1, y4a, rows, H4, D, epi=ops.EPI_BIAS_GELU, D2=y4b, bias=b1, block_n=bn)))
    cases.append((f"fc1 fwd gelu-only bn={bn}", lambda bn=bn: ops.gemm(x, w1, y4a, rows, H4, D, epi=ops.EPI_BIAS_GELU_FWD, bias=b1, block_n=bn)))
    cases.append((f"fc2 dgrad*aux bn={bn}", lambda bn=bn: ops.gemm(x, w2, y4a, rows, H4, D, b_mn=True, epi=ops.EPI_MUL_AUX, aux=x4, block_n=bn)))
for rep in range(2):
    for name, fn in cases:
        print(f"{name:28s} {timeit(fn):7.1f} us")

# proj wgrad (384 x 384 output, K = rows): single-CTA 128x192 split-K vs the 256x384 pair tile
dy3 = r(rows, D).bfloat16()
dw = torch.zeros(D, D, device="cuda")
db = torch.zeros(D, device="cuda")
for bn in (192, 384, 128):
    t = timeit(lambda bn=bn: ops.gemm(dy3, x, dw, D, D, rows, a_mn=True, b_mn=True, epi=ops.EPI_ATOMIC_F32, split_k=0, bias=db, block_n=bn))
    print(f"proj wgrad bn={bn:3d}          {t:7.1f} us")
