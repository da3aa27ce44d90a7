"""GPU multi-crop tile augmentation behind the reference's transform vocabulary.

The reference builds a per-tile host pipeline with ``define_transformations(transform_type, train, tile_size,
color_param, norm_type)`` (transformations.py:103-208, called from datasets.py:423 and applied per tile at
datasets.py:498-502): ColorJitter, GaussianBlur(3), Gaussian noise, vertical flip, rotation by a multiple of 90
degrees, RandomAffine(scale), ToTensor, Normalize -- PIL on 2 CPU cores per job (sbatch-ssl.sh:20). Here the same
operators (and the DINO multi-crop random-resized crops the training step consumes) run in ONE kernel launch per batch
from uint8 256 x 256 tiles: ``MultiCropAugment`` samples the per-(tile, crop) parameters on the host (so a CPU oracle
can replay the identical draw), ``__call__`` launches ``b200ssl_multicrop_augment`` and returns the crops in the
crop-major layout ``GraphedDinoStep`` / ``MultiCropWrapper`` consume without a concatenation copy.
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import _lib, ops

# transformations.py:104-116
MEAN = {"TCGA": [58.2069073 / 255, 96.22645279 / 255, 70.26442606 / 255],
        "HEROHE": [224.46091564 / 255, 190.67338568 / 255, 218.47883547 / 255],
        "Ron": [0.8998, 0.8253, 0.9357],
        "Imagenet": [0.485, 0.456, 0.406],
        "Amir": [0.9357, 0.8253, 0.8998]}
STD = {"TCGA": [40.40400300279664 / 255, 58.90625962739444 / 255, 45.09334057330417 / 255],
       "HEROHE": [math.sqrt(1110.25292532) / 255, math.sqrt(2950.9804851) / 255, math.sqrt(1027.10911208) / 255],
       "Ron": [0.1125, 0.1751, 0.0787],
       "Imagenet": [0.229, 0.224, 0.225],
       "Amir": [0.0787, 0.1751, 0.1125]}

PARAM_WORDS = 16
TILE = 256

# transform_type letters of define_transformations (transformations.py:131-205): which operator groups are on
_TYPES = {
    "none": "", "flip": "f", "rvf": "fr", "frs": "frs",
    "cbnfr": "cbnfr", "cbnfrs": "cbnfrs", "cbnfrsc": "cbnfrs",
    "pcbnfrs": "pcbnfrs", "pcbnfrsc": "pcbnfrs", "bnfrs": "bnfrs", "bnfrsc": "bnfrs",
}
# the types define_transformations appends Cutout(n_holes=1, length=100) to, behind Normalize (transformations.py:206-207;
# 'c_0_05_bnfrsc' of that list never gets there: it leaves through the "custom transform" return at :201-202)
_CUTOUT_TYPES = ("cbnfrsc", "bnfrsc", "pcbnfrsc")
CUTOUT_LENGTH = 100      # pixels of the reference's 256 x 256 output frame (transformations.py:207)


def cutout_box(y, x, length, size):
    """The hole of the reference's ``Cutout.__call__`` (transformations.py:35-40) for a centre draw (y, x) on a
    ``size`` x ``size`` image: ``[y1, y2) x [x1, x2)`` with ``clip(c -+ length // 2, 0, size)``. Tensors or ints."""
    half = length // 2
    lo = lambda c: torch.clamp(torch.as_tensor(c) - half, 0, size)    # noqa: E731
    hi = lambda c: torch.clamp(torch.as_tensor(c) + half, 0, size)    # noqa: E731
    return lo(y), hi(y), lo(x), hi(x)


class MultiCropAugment:
    """``MultiCropAugment(transform_type='pcbnfrs', color_param=0.1, norm_type='Ron')(tiles_u8) -> [crop_0 .. crop_11]``.

    tiles: uint8 ``[B, 256, 256, 3]`` CUDA tensor (HWC RGB, as the slide reader delivers them). Returns ``n_global``
    tensors ``[B, 3, Sg, Sg]`` followed by ``n_local`` tensors ``[B, 3, Sl, Sl]`` (bf16, normalised); equal-size crops
    are views of one allocation, back to back.

    Operator semantics follow torchvision's tensor implementations with the parameter ranges of the reference's
    ``define_transformations``: 'c' ColorJitter(brightness=(0.85,1.15), contrast=(0.75,1.25), saturation=0.1,
    hue=(-0.1,0.1)); 'pc' ColorJitter(color_param, 2*color_param, color_param, color_param); 'b' GaussianBlur(3,
    sigma<=0.1) (the identity to 2e-22: skipped); 'n' Gaussian noise sigma ~ U(0, 0.05); 'f' vertical flip p=0.5 (and
    horizontal flip p=0.5 for 'flip'); 'r' rotation by a uniform multiple of 90 degrees; 's' zoom 1..1.2 about the
    centre (folded into the crop box); trailing 'c' (``cbnfrsc``, ``bnfrsc``, ``pcbnfrsc``) Cutout(n_holes=1): one
    square hole, centre uniform over the output frame, clipped at the borders, set to 0 AFTER Normalize like the
    reference does. The reference cuts 100 pixels out of its 256 x 256 output; a crop of side S gets
    ``cutout_length * S // 256`` (the same fraction of the frame), ``cutout_length`` defaulting to 100.
    Random-resized crops: area fraction ~ U(scale), log-uniform aspect ratio in (3/4, 4/3), torchvision's
    ``RandomResizedCrop.get_params`` procedure."""

    def __init__(self, transform_type="pcbnfrs", color_param=0.1, norm_type="Ron", global_size=224, local_size=96,
                 n_global=2, n_local=10, global_scale=(0.4, 1.0), local_scale=(0.05, 0.4), train=True,
                 cutout_length=CUTOUT_LENGTH):
        if transform_type not in _TYPES:
            raise ValueError(f"unknown transform_type {transform_type!r}")
        self.ops = _TYPES[transform_type] if train else ""
        # define_transformations appends the Cutout outside its `train` branch (transformations.py:206-207)
        self.cutout = transform_type in _CUTOUT_TYPES
        self.cutout_length = int(cutout_length)
        if self.cutout and not 0 <= self.cutout_length < 65536:
            raise ValueError(f"cutout_length {cutout_length!r} out of range")
        self.transform_type, self.color_param, self.norm_type = transform_type, color_param, norm_type
        self.global_size, self.local_size, self.n_global, self.n_local = global_size, local_size, n_global, n_local
        self.global_scale, self.local_scale = global_scale, local_scale
        self.mean = (ctypes.c_float * 3)(*MEAN[norm_type])
        self.std = (ctypes.c_float * 3)(*STD[norm_type])

    # ------------------------------------------------------------------ parameter sampling (host, vectorised)
    def _boxes(self, n, scale, g):
        """torchvision RandomResizedCrop.get_params on a TILE x TILE image for n crops at once (ten attempts, then the
        whole tile), plus the 's' zoom about the centre folded into the box. -> int64 [n, 4] (top, left, h, w)."""
        area = float(TILE * TILE)
        target = area * torch.empty(n, 10).uniform_(scale[0], scale[1], generator=g)
        ar = torch.exp(torch.empty(n, 10).uniform_(math.log(3 / 4), math.log(4 / 3), generator=g))
        w = torch.round(torch.sqrt(target * ar)).long()
        h = torch.round(torch.sqrt(target / ar)).long()
        ok = (w > 0) & (w <= TILE) & (h > 0) & (h <= TILE)
        first = torch.where(ok.any(1), ok.float().argmax(1), torch.zeros(n, dtype=torch.long))
        idx = torch.arange(n)
        w = torch.where(ok.any(1), w[idx, first], torch.full((n,), TILE))
        h = torch.where(ok.any(1), h[idx, first], torch.full((n,), TILE))
        top = (torch.rand(n, generator=g) * (TILE - h + 1).float()).long().clamp_(max=TILE - 1)
        left = (torch.rand(n, generator=g) * (TILE - w + 1).float()).long().clamp_(max=TILE - 1)
        top, left = torch.minimum(top, TILE - h), torch.minimum(left, TILE - w)
        if "s" in self.ops:
            z = torch.empty(n).uniform_(1.0, 1.2, generator=g)
            nh = torch.round(h.float() / z).long().clamp_(min=1)
            nw = torch.round(w.float() / z).long().clamp_(min=1)
            top, left, h, w = top + (h - nh) // 2, left + (w - nw) // 2, nh, nw
        return torch.stack((top, left, h, w), 1)

    def sample_params(self, B, generator=None):
        """int32 tensor [B, ncrops, 16] (CPU): one AugParams row per (tile, crop); float fields bit-cast."""
        g = generator
        ncrops = self.n_global + self.n_local
        n = B * ncrops
        p = torch.zeros(B, ncrops, PARAM_WORDS, dtype=torch.int32)
        fp = p.view(torch.float32)
        boxes = torch.empty(B, ncrops, 4, dtype=torch.long)
        if self.n_global:
            boxes[:, :self.n_global] = self._boxes(B * self.n_global, self.global_scale, g).view(B, self.n_global, 4)
        if self.n_local:
            boxes[:, self.n_global:] = self._boxes(B * self.n_local, self.local_scale, g).view(B, self.n_local, 4)
        p[..., 0:4] = boxes.int()
        u = lambda lo, hi: torch.empty(B, ncrops).uniform_(lo, hi, generator=g)   # noqa: E731
        flags = torch.zeros(B, ncrops, dtype=torch.long)
        if "f" in self.ops:
            flags |= (torch.rand(B, ncrops, generator=g) < 0.5).long() << 1                 # vertical flip
            if self.transform_type == "flip":
                flags |= (torch.rand(B, ncrops, generator=g) < 0.5).long()                  # horizontal flip
        if "r" in self.ops:
            flags |= torch.randint(0, 4, (B, ncrops), generator=g) << 2
        if "c" in self.ops:
            order = torch.rand(n, 4, generator=g).argsort(1).view(B, ncrops, 4)              # uniform random permutations
            for s_ in range(4):
                flags |= order[..., s_] << (4 + 2 * s_)
            flags |= 1 << 12
            if "pc" in self.ops:
                cp = self.color_param
                fp[..., 5], fp[..., 6] = u(max(0.0, 1 - cp), 1 + cp), u(max(0.0, 1 - 2 * cp), 1 + 2 * cp)
                fp[..., 7], fp[..., 8] = u(max(0.0, 1 - cp), 1 + cp), u(-cp, cp)
            else:
                fp[..., 5], fp[..., 6], fp[..., 7], fp[..., 8] = u(0.85, 1.15), u(0.75, 1.25), u(0.9, 1.1), u(-0.1, 0.1)
        else:
            fp[..., 5:8] = 1.0
        if "n" in self.ops:
            fp[..., 9] = u(0.0, 0.05)
        p[..., 4] = flags.int()
        p[..., 10] = torch.randint(0, 2 ** 31 - 1, (B, ncrops), generator=g).int()
        if self.cutout:
            # words 11 / 12: the hole in OUTPUT-frame pixels, y1 | y2 << 16 and x1 | x2 << 16 (0 = no hole)
            for lo, hi, S in ((0, self.n_global, self.global_size), (self.n_global, ncrops, self.local_size)):
                if hi > lo:
                    y = torch.randint(0, S, (B, hi - lo), generator=g)
                    x = torch.randint(0, S, (B, hi - lo), generator=g)
                    y1, y2, x1, x2 = cutout_box(y, x, self.cutout_length * S // TILE, S)
                    p[:, lo:hi, 11] = (y1 | (y2 << 16)).int()
                    p[:, lo:hi, 12] = (x1 | (x2 << 16)).int()
        return p

    # ------------------------------------------------------------------ launch
    def alloc_outputs(self, B, device):
        g = torch.empty(self.n_global, B, 3, self.global_size, self.global_size, dtype=torch.bfloat16, device=device)
        l = torch.empty(self.n_local, B, 3, self.local_size, self.local_size, dtype=torch.bfloat16, device=device)
        return g, l

    def __call__(self, tiles, params=None, generator=None, out=None):
        ops.require_cuda(tiles, "MultiCropAugment")
        if tiles.dtype != torch.uint8 or tiles.dim() != 4 or tuple(tiles.shape[1:]) != (TILE, TILE, 3):
            raise RuntimeError(f"MultiCropAugment expects uint8 [B, {TILE}, {TILE}, 3] tiles, got {tuple(tiles.shape)} "
                               f"{tiles.dtype}")
        tiles = tiles.contiguous()
        B = tiles.shape[0]
        if params is None:
            params = self.sample_params(B, generator)
        if not params.is_cuda:
            params = params.pin_memory().to(tiles.device, non_blocking=True)
        params = params.contiguous()
        g, l = out if out is not None else self.alloc_outputs(B, tiles.device)
        ops._call("b200ssl_multicrop_augment", tiles.data_ptr(), params.data_ptr(), g.data_ptr(), l.data_ptr(), B,
                  self.n_global, self.n_local, self.global_size, self.local_size,
                  ctypes.cast(self.mean, ctypes.c_void_p), ctypes.cast(self.std, ctypes.c_void_p), ops._stream())
        return [g[i] for i in range(self.n_global)] + [l[i] for i in range(self.n_local)]
