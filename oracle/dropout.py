"""ORACLE (test infrastructure, not product code): the counter-based dropout mask of b200ssl's encoder path, restated
with numpy uint64 arithmetic, and an ``nn.Dropout`` stand-in that replays it inside the oracle model.

Only tests/ may import this. What it follows: ``nn.Dropout(p)`` of the reference's VisionTransformer -- ``pos_drop``
(VT.pyc@L196,245), ``Attention.proj_drop`` (@L117,130), ``Mlp.drop`` used twice per forward (@L96,101-104) -- i.e.
``x * mask / (1 - p)`` in training mode, the identity in eval mode. torch draws the mask from its global Philox stream,
which no other implementation can reproduce; to compare element for element the product's generator (csrc/dropout.cu:
one splitmix64 word per four consecutive elements, 16 bits each, dropped when below round(p * 65536)) is restated here
and the oracle's Dropout modules are swapped for ``ReplayDropout`` carrying the seed the product drew.

Site numbering of the product (vision_transformer.py): 0 = pos_drop; block i: 1 + 3 i = attn.proj_drop,
2 + 3 i = mlp.drop behind the activation, 3 + 3 i = mlp.drop behind fc2.
"""
import numpy as np
import torch
import torch.nn as nn

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z):
    z = z + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111C5)
    return z ^ (z >> np.uint64(31))


def keep_mask(seed: int, site: int, n: int, p: float) -> torch.Tensor:
    """bool [n]: True where element i (flat, row-major) is KEPT."""
    with np.errstate(over="ignore"):
        key = np.array([seed & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64) ^ \
            (np.array([site + 1], dtype=np.uint64) * np.uint64(0xA0761D6478BD642F))
        idx = np.arange(n, dtype=np.uint64)
        word = _splitmix64(key + (idx >> np.uint64(2)) * np.uint64(0xD1342543DE82EF95))
        bits = (word >> (np.uint64(16) * (idx & np.uint64(3)))) & np.uint64(0xFFFF)
    thr = min(max(int(np.rint(np.float32(p) * np.float32(65536.0))), 0), 65536)
    return torch.from_numpy(bits >= np.uint64(thr))


class ReplayDropout(nn.Module):
    """nn.Dropout(p) with the mask of the given site(s); a module called several times per forward (Mlp.drop) walks
    through its list of sites call by call."""

    def __init__(self, p, seed, sites):
        super().__init__()
        self.p, self.seed, self.sites, self.calls = float(p), int(seed), list(sites), 0

    def forward(self, x):
        if not self.training or self.p == 0.0:
            return x
        site = self.sites[self.calls % len(self.sites)]
        self.calls += 1
        keep = keep_mask(self.seed, site, x.numel(), self.p).to(x.device).view(x.shape)
        scale = 1.0 / (1.0 - self.p) if self.p < 1.0 else 0.0
        return x * keep.to(x.dtype) * scale


def replay_in(model, seed):
    """Swap the Dropout modules of an oracle VisionTransformer for ReplayDropout with the product's site numbering."""
    model.pos_drop = ReplayDropout(model.pos_drop.p, seed, [0])
    for i, blk in enumerate(model.blocks):
        blk.attn.proj_drop = ReplayDropout(blk.attn.proj_drop.p, seed, [1 + 3 * i])
        blk.mlp.drop = ReplayDropout(blk.mlp.drop.p, seed, [2 + 3 * i, 3 + 3 * i])
    return model
