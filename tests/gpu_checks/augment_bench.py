"""Developer bench: multi-crop augmentation kernel at the config-2 batch (256 tiles -> 2 x 224^2 + 10 x 96^2 crops)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import b200ssl

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
tiles = torch.randint(0, 256, (B, 256, 256, 3), dtype=torch.uint8, device="cuda")
for ttype in ("none", "flip", "cbnfr", "pcbnfrs"):
    a = b200ssl.MultiCropAugment(ttype)
    p = a.sample_params(B, torch.Generator().manual_seed(0)).cuda()
    out = a.alloc_outputs(B, tiles.device)
    for _ in range(3):
        a(tiles, params=p, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        a(tiles, params=p, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    bytes_ = tiles.numel() + sum(o.numel() * 2 for o in out)
    print(f"{ttype:8s} B={B}: {ms*1e3:8.1f} us  {bytes_/ms/1e6:7.0f} GB/s (tile read + crop writes = {bytes_/1e6:.0f} MB)")
