"""ORACLE (test infrastructure, not product code): plain-PyTorch restatement of the multi-crop tile augmentation.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product path
(b200ssl.augment -> b200ssl_multicrop_augment) never does.

What it follows: the reference's per-tile transform pipeline, transformations.py:103-208 (``define_transformations``:
ColorJitter :142-143/:153-154, GaussianBlur :144, MyGaussianNoiseTransform :71-87, RandomVerticalFlip :146,
MyRotation :47-55, RandomAffine(scale) :148, ToTensor + Normalize with the MEAN / STD tables :104-128, Cutout :10-45
appended behind Normalize :206-207), applied per
tile at datasets.py:498-502 -- with torchvision's own TENSOR implementations of the operators
(torchvision.transforms.functional.adjust_brightness / adjust_contrast / adjust_saturation / adjust_hue on float
images, torch.rot90, flips, F.interpolate bilinear) -- plus the DINO multi-crop random-resized crops the training step
consumes (public DINO recipe; NOT in the reference: parity unpinned for the crop geometry, as for the loss).
Stated differences from the reference's PIL pipeline: no uint8 re-quantisation between operators; the resize is
bilinear without antialiasing; GaussianBlur(3, sigma <= 0.1) is the identity to 2e-22 and omitted; the noise comes from
a counter-based generator (below) instead of numpy's global stream, so that a GPU kernel can reproduce it.

The per-(tile, crop) parameters are an explicit table (b200ssl.augment.MultiCropAugment.sample_params), so product and
oracle replay the identical draw.
"""
import math

import torch
import torch.nn.functional as F


def _pcg_out(x):
    word = (((x >> ((x >> 28) + 4)) ^ x) * 277803737) & 0xFFFFFFFF
    return ((word >> 22) ^ word) & 0xFFFFFFFF


def normal_of(seed: int, idx: torch.Tensor) -> torch.Tensor:
    """Standard normal per element index: two PCG-hash outputs of the element PAIR (idx >> 1) -> Box-Muller, the even
    element takes the cosine branch and the odd one the sine branch (uint32 arithmetic emulated in int64)."""
    idx = idx.to(torch.int64)
    pair = idx >> 1
    x1 = (pair * 747796405 + (seed * 2891336453 + 12345)) & 0xFFFFFFFF
    x2 = (x1 * 747796405 + 2891336453) & 0xFFFFFFFF
    u1 = ((_pcg_out(x1) >> 8).to(torch.float32) + 1.0) * (1.0 / 16777216.0)
    u2 = (_pcg_out(x2) >> 8).to(torch.float32) * (1.0 / 16777216.0)
    rad, theta = torch.sqrt(-2.0 * torch.log(u1)), 6.283185307179586 * u2
    return rad * torch.where((idx & 1) == 1, torch.sin(theta), torch.cos(theta))


def augment_one(tile_u8: torch.Tensor, row: torch.Tensor, size: int, mean, std) -> torch.Tensor:
    """tile uint8 [256,256,3]; row int32 [16] (AugParams); -> float32 [3,size,size] normalised crop."""
    import torchvision.transforms.functional as TF
    top, left, h, w, flags = (int(v) for v in row[:5])
    bri, con, sat, hue, sigma = (float(v) for v in row.view(torch.float32)[5:10])
    seed = int(row[10]) & 0xFFFFFFFF
    crop = tile_u8[top:top + h, left:left + w].permute(2, 0, 1).float() / 255.0
    img = F.interpolate(crop[None], size=(size, size), mode="bilinear", align_corners=False, antialias=False)[0]
    if flags & 1:
        img = img.flip(-1)
    if flags & 2:
        img = img.flip(-2)
    img = torch.rot90(img, (flags >> 2) & 3, dims=(-2, -1))
    if (flags >> 12) & 1:
        for s in range(4):
            op = (flags >> (4 + 2 * s)) & 3
            if op == 0:
                img = TF.adjust_brightness(img, bri)
            elif op == 1:
                img = TF.adjust_contrast(img, con)
            elif op == 2:
                img = TF.adjust_saturation(img, sat)
            else:
                img = TF.adjust_hue(img, hue)
    if sigma > 0:
        idx = torch.arange(3 * size * size, device=img.device).view(3, size, size)
        img = (img + sigma * normal_of(seed, idx)).clamp(0.0, 1.0)
    m = torch.tensor(mean, dtype=torch.float32, device=img.device).view(3, 1, 1)
    s_ = torch.tensor(std, dtype=torch.float32, device=img.device).view(3, 1, 1)
    return cutout((img - m) / s_, int(row[11]), int(row[12]))


def cutout(img: torch.Tensor, word_y: int, word_x: int) -> torch.Tensor:
    """Cutout.__call__ of the reference (transformations.py:21-45) for a hole already drawn: ``img * mask`` with the
    mask 0 on ``[y1, y2) x [x1, x2)`` and 1 elsewhere, applied to the NORMALISED image (define_transformations appends it
    behind Normalize, :206-207). The words hold ``y1 | y2 << 16`` / ``x1 | x2 << 16`` (0 = no hole)."""
    y1, y2, x1, x2 = word_y & 0xFFFF, (word_y >> 16) & 0xFFFF, word_x & 0xFFFF, (word_x >> 16) & 0xFFFF
    mask = torch.ones(img.shape[-2:], dtype=img.dtype, device=img.device)
    mask[y1:y2, x1:x2] = 0.0
    return img * mask.expand_as(img)


def multicrop_augment(tiles_u8: torch.Tensor, params: torch.Tensor, n_global: int, n_local: int, size_global: int,
                      size_local: int, mean, std):
    """tiles uint8 [B,256,256,3], params int32 [B, n_global+n_local, 16] -> list of float32 [B,3,S,S] per crop."""
    B = tiles_u8.shape[0]
    out = []
    for c in range(n_global + n_local):
        size = size_global if c < n_global else size_local
        out.append(torch.stack([augment_one(tiles_u8[b], params[b, c], size, mean, std) for b in range(B)]))
    return out


def reference_style_pipeline(tile_u8: torch.Tensor, row: torch.Tensor, size: int, mean, std) -> torch.Tensor:
    """The same draw through torchvision's high-level tensor transforms in the REFERENCE's operator order
    (transformations.py:153-160: ColorJitter ops, noise, vertical flip, rotation -- then Normalize), starting from the
    resized crop: a cross-check that augment_one's re-ordering (geometry first, colour after) is the same function."""
    import torchvision.transforms.functional as TF
    top, left, h, w, flags = (int(v) for v in row[:5])
    bri, con, sat, hue, sigma = (float(v) for v in row.view(torch.float32)[5:10])
    img = TF.resized_crop(tile_u8.permute(2, 0, 1).float() / 255.0, top, left, h, w, [size, size],
                          interpolation=TF.InterpolationMode.BILINEAR, antialias=False)
    if (flags >> 12) & 1:
        for s in range(4):
            op = (flags >> (4 + 2 * s)) & 3
            img = (TF.adjust_brightness(img, bri) if op == 0 else TF.adjust_contrast(img, con) if op == 1 else
                   TF.adjust_saturation(img, sat) if op == 2 else TF.adjust_hue(img, hue))
    if flags & 1:
        img = TF.hflip(img)
    if flags & 2:
        img = TF.vflip(img)
    k = (flags >> 2) & 3
    if k:
        img = TF.rotate(img, 90.0 * k)     # MyRotation: transforms.functional.rotate(x, angle), counter-clockwise
    return cutout(TF.normalize(img, list(mean), list(std)), int(row[11]), int(row[12]))
