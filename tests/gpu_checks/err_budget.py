"""Developer check: where does the bf16 error of the logits come from? Compares (a) b200ssl and
(b) the oracle under torch bf16 autocast against the fp32 oracle at several points of the student."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import b200ssl
from oracle import dino as odino
from oracle import vision_transformer as ovt

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm()).item()


def run(name, out_dim, hidden, bott, B=4):
    torch.manual_seed(0)
    factory = getattr(ovt, name)
    D = {"vit_tiny": 192, "vit_small": 384}[name]
    ref = odino.MultiCropWrapper(factory(), ovt.DINOHead(D, out_dim, hidden_dim=hidden, bottleneck_dim=bott)).cuda()
    with torch.no_grad():
        for p in ref.parameters():
            if p.ndim == 1:
                p.add_(torch.randn_like(p) * 0.02)
    mine = b200ssl.MultiCropWrapper(getattr(b200ssl, name)(), b200ssl.DINOHead(D, out_dim, hidden_dim=hidden, bottleneck_dim=bott)).cuda()
    mine.load_state_dict(ref.state_dict())
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, 3, 224, 224, device="cuda", generator=g)
    xs = torch.randn(B, 3, 96, 96, device="cuda", generator=g)
    with torch.no_grad():
        for inp, tag in ((x, "224"), (xs, "96")):
            f_ref = ref.backbone(inp)
            f_mine = mine.backbone(inp.bfloat16())
            with torch.autocast("cuda", dtype=torch.bfloat16):
                f_amp = ref.backbone(inp)
            l_ref = ref.head(f_ref)
            l_mine = mine.head(f_mine)
            l_mine_exactfeat = mine.head(f_ref.bfloat16())
            with torch.autocast("cuda", dtype=torch.bfloat16):
                l_amp = ref.head(f_amp)
                l_amp_exactfeat = ref.head(f_ref)
            print(f"{name} {tag} hidden={hidden} bott={bott} K={out_dim}: feat mine {rel(f_mine, f_ref):.2e} amp {rel(f_amp, f_ref):.2e} | "
                  f"logits mine {rel(l_mine, l_ref):.2e} amp {rel(l_amp, l_ref):.2e} | head-only mine {rel(l_mine_exactfeat, l_ref):.2e} "
                  f"amp {rel(l_amp_exactfeat, l_ref):.2e}")


run("vit_tiny", 2048, 256, 64)
run("vit_tiny", 4096, 2048, 256)
run("vit_small", 8192, 2048, 256)
