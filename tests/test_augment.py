"""Multi-crop tile augmentation: oracle self-checks on the CPU, kernel parity on the GPU.

Reference operators: transformations.py:103-208 (define_transformations), applied per tile at datasets.py:498-502."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200ssl  # noqa: E402
from b200ssl import augment as aug  # noqa: E402
from oracle import augment as oaug  # noqa: E402


def _tiles(B, seed=0):
    g = torch.Generator().manual_seed(seed)
    # smooth structure + texture, in the colour range of H&E tiles (bright, low saturation)
    base = torch.rand(B, 16, 16, 3, generator=g).permute(0, 3, 1, 2)
    img = torch.nn.functional.interpolate(base, size=(256, 256), mode="bicubic", align_corners=False).clamp(0, 1)
    img = 0.55 + 0.45 * img + 0.04 * torch.randn(B, 3, 256, 256, generator=g)
    return (img.clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def test_sampler_produces_valid_tables():
    a = b200ssl.MultiCropAugment("pcbnfrs", color_param=0.1)
    p = a.sample_params(64, torch.Generator().manual_seed(1))
    assert p.shape == (64, 12, aug.PARAM_WORDS) and p.dtype == torch.int32
    top, left, h, w = (p[..., i].long() for i in range(4))
    assert (h >= 1).all() and (w >= 1).all() and (top >= 0).all() and (left >= 0).all()
    assert (top + h <= 256).all() and (left + w <= 256).all()
    area = (h * w).float() / 65536
    assert area[:, :2].min() > 0.4 / 1.2 ** 2 * 0.9 and area[:, 2:].max() < 0.4 * 1.1      # scale ranges (with the zoom)
    fp = p.view(torch.float32)
    assert ((fp[..., 5] >= 0.9) & (fp[..., 5] <= 1.1)).all() and ((fp[..., 6] >= 0.8) & (fp[..., 6] <= 1.2)).all()
    assert (fp[..., 8].abs() <= 0.1).all() and ((fp[..., 9] >= 0) & (fp[..., 9] <= 0.05)).all()
    order = torch.stack([(p[..., 4] >> (4 + 2 * s)) & 3 for s in range(4)], -1)
    assert (order.sort(-1).values == torch.arange(4, dtype=torch.int32)).all()              # a permutation per crop
    assert ((p[..., 4] >> 12) & 1).all()
    none = b200ssl.MultiCropAugment("none").sample_params(4, torch.Generator().manual_seed(1))
    assert (none[..., 4] == 0).all() and (none.view(torch.float32)[..., 9] == 0).all()
    with pytest.raises(ValueError):
        b200ssl.MultiCropAugment("no_such_type")


def test_normalisation_tables_match_the_reference_constants():
    # transformations.py:104-116
    assert aug.MEAN["Ron"] == [0.8998, 0.8253, 0.9357] and aug.STD["Ron"] == [0.1125, 0.1751, 0.0787]
    assert aug.MEAN["Imagenet"] == [0.485, 0.456, 0.406] and aug.STD["Imagenet"] == [0.229, 0.224, 0.225]
    assert abs(aug.MEAN["TCGA"][0] - 58.2069073 / 255) < 1e-12 and aug.MEAN["Amir"] == aug.MEAN["Ron"][::-1]


def test_oracle_counter_based_noise_is_standard_normal():
    n = oaug.normal_of(12345, torch.arange(400000))
    assert abs(float(n.mean())) < 5e-3 and abs(float(n.std()) - 1.0) < 5e-3
    assert torch.equal(n[:100], oaug.normal_of(12345, torch.arange(100)))                   # a pure function of (seed, index)
    assert not torch.equal(n[:100], oaug.normal_of(12346, torch.arange(100)))


def test_oracle_matches_torchvision_pipeline_in_the_reference_operator_order():
    """augment_one applies geometry first and the colour operators after it; the reference order is colour, flips,
    rotation (transformations.py:153-160). Same function: every colour operator is pointwise except contrast, whose
    mean grey level is invariant under flips and rotations."""
    tiles = _tiles(3)
    a = b200ssl.MultiCropAugment("pcbnfrs", color_param=0.1)
    p = a.sample_params(3, torch.Generator().manual_seed(2))
    p.view(torch.float32)[..., 9] = 0.0           # the noise is defined on the output frame: excluded from this comparison
    mean, std = aug.MEAN["Ron"], aug.STD["Ron"]
    for b in range(3):
        for c in (0, 1, 2, 5, 11):
            size = 224 if c < 2 else 96
            x = oaug.augment_one(tiles[b], p[b, c], size, mean, std)
            y = oaug.reference_style_pipeline(tiles[b], p[b, c], size, mean, std)
            assert x.shape == (3, size, size)
            assert float((x - y).abs().max()) < 1e-4, (b, c, float((x - y).abs().max()))


def test_float_pipeline_stays_within_quantisation_of_the_pil_pipeline():
    """The reference applies its operators to PIL images (uint8 between operators, hue rotated in 8-bit HSV); the
    oracle / kernel keep floats throughout -- a stated difference (oracle/augment.py). This bounds it: the same draw
    through torchvision's PIL implementations (what transformations.py:131-160 runs per tile) against augment_one, in
    grey levels of 255. Per operator the two differ by PIL's truncation to uint8 (< 1 level; hue <= 7); the four colour
    operators in sequence: mean < 2, 99th percentile < 7, maximum < 16 levels. Flips and rotations are exact."""
    from PIL import Image
    import torchvision.transforms.functional as TF
    tiles = _tiles(3, seed=11)
    ft = tiles[0].permute(2, 0, 1).float() / 255
    pil = Image.fromarray(tiles[0].numpy())
    for fn, factors, bound in ((TF.adjust_brightness, (0.9, 1.1), 1.0), (TF.adjust_contrast, (0.8, 1.2), 1.5),
                               (TF.adjust_saturation, (0.9, 1.1), 1.0), (TF.adjust_hue, (-0.1, 0.05, 0.1), 8.0)):
        for f in factors:
            d = (fn(ft, f) - TF.to_tensor(fn(pil, f))).abs() * 255
            assert float(d.max()) < bound and float(d.mean()) < 0.8, (fn.__name__, f, float(d.max()), float(d.mean()))
    for k in (1, 2, 3):
        assert torch.equal(torch.rot90(ft, k, dims=(-2, -1)), TF.to_tensor(TF.rotate(pil, 90 * k)))
    assert torch.equal(ft.flip(-2), TF.to_tensor(TF.vflip(pil)))
    a = b200ssl.MultiCropAugment("pcbnfrs", color_param=0.1, global_size=256, local_size=256, n_global=1, n_local=1)
    p = a.sample_params(3, torch.Generator().manual_seed(3))
    p[..., 0:2] = 0
    p[..., 2:4] = 256                              # whole tile at its own size: no resampling in this comparison
    p.view(torch.float32)[..., 9] = 0.0            # no noise (the generators differ by construction)
    mean, std = aug.MEAN["Ron"], aug.STD["Ron"]
    s = torch.tensor(std).view(3, 1, 1)
    for b in range(3):
        for c in range(2):
            row = p[b, c]
            flags, fp = int(row[4]), row.view(torch.float32)
            img = Image.fromarray(tiles[b].numpy())
            for k in range(4):                       # ColorJitter: the drawn order, PIL implementations
                op = (flags >> (4 + 2 * k)) & 3
                img = (TF.adjust_brightness, TF.adjust_contrast, TF.adjust_saturation, TF.adjust_hue)[op](img, float(fp[5 + op]))
            if flags & 2:
                img = TF.vflip(img)
            if (flags >> 2) & 3:
                img = TF.rotate(img, 90 * ((flags >> 2) & 3))
            ref = TF.normalize(TF.to_tensor(img), mean, std)
            d = ((oaug.augment_one(tiles[b], row, 256, mean, std) - ref) * s).abs().flatten() * 255
            p99 = float(d.kthvalue(int(0.99 * d.numel())).values)
            assert float(d.mean()) < 2.0 and p99 < 7.0 and float(d.max()) < 16.0, (b, c, float(d.mean()), p99, float(d.max()))


def _gold():
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "transform_goldens.json")) as fh:
        return json.load(fh)


def test_vocabulary_matches_the_reference_pipelines():
    """tests/golden/transform_goldens.json is what the reference's own define_transformations builds for every
    transform_type (generated by tests/golden/make_transform_goldens.py from the unmodified module): the operator
    groups, sampling ranges and normalisation constants of MultiCropAugment must be those."""
    gold = _gold()
    for nt, t in gold["norm"].items():
        assert aug.MEAN[nt] == pytest.approx(t["mean"], rel=0, abs=1e-15) and aug.STD[nt] == pytest.approx(t["std"], abs=1e-15)
    letter_of = {"ColorJitter": "c", "GaussianBlur": "b", "MyGaussianNoiseTransform": "n", "RandomVerticalFlip": "f",
                 "MyRotation": "r", "RandomAffine": "s"}
    cp = 0.1                                                    # the color_param the goldens were generated with
    for ttype in aug._TYPES:
        for mode in ("train", "eval"):
            ops_ = gold["pipelines"][f"{ttype}/{mode}"]
            names = [o["op"] for o in ops_]
            a = b200ssl.MultiCropAugment(ttype, color_param=cp, train=(mode == "train"))
            want = {letter_of[n] for n in names if n in letter_of}
            have = set(a.ops) - {"p"}
            assert have == want, (ttype, mode, have, want)
            assert a.cutout == ("Cutout" in names), (ttype, mode)
            if a.cutout:
                assert names[-1] == "Cutout" and names[-2] == "Normalize"          # the hole is cut from the NORMALISED image
                cut = ops_[-1]
                assert cut["n_holes"] == 1 and cut["length"] == aug.CUTOUT_LENGTH == a.cutout_length
            assert ("RandomHorizontalFlip" in names) == (ttype == "flip" and mode == "train")
            if mode == "eval":
                continue
            # sampled ranges: 4,000 draws stay inside the reference's interval and cover most of it
            p = a.sample_params(400, torch.Generator().manual_seed(7))[:, :10]
            fp = p.view(torch.float32)
            for o in ops_:
                if o["op"] == "ColorJitter":
                    for word, key in ((5, "brightness"), (6, "contrast"), (7, "saturation"), (8, "hue")):
                        lo, hi = o[key]
                        v = fp[..., word]
                        assert float(v.min()) >= lo - 1e-6 and float(v.max()) <= hi + 1e-6, (ttype, key)
                        assert float(v.min()) < lo + 0.02 * (hi - lo) and float(v.max()) > hi - 0.02 * (hi - lo), (ttype, key)
                    assert ("p" in a.ops) == (ttype.startswith("p"))
                elif o["op"] == "MyGaussianNoiseTransform":
                    lo, hi = o["sigma"]
                    assert float(fp[..., 9].min()) >= lo and float(fp[..., 9].max()) <= hi
                    assert float(fp[..., 9].max()) > 0.98 * hi
                elif o["op"] == "MyRotation":
                    assert o["angles"] == [0, 90, 180, 270]
                    assert sorted(((p[..., 4] >> 2) & 3).unique().tolist()) == [0, 1, 2, 3]
                elif o["op"] == "RandomAffine":
                    assert o["degrees"] == [0.0, 0.0] and o["scale"] == [1.0, 1.2] and o["translate"] is None and o["shear"] is None
                elif o["op"] == "RandomVerticalFlip":
                    assert o["p"] == 0.5 and 0.4 < float(((p[..., 4] >> 1) & 1).float().mean()) < 0.6
                elif o["op"] == "GaussianBlur":
                    assert o["kernel_size"] == [3, 3] and o["sigma"][1] <= 0.1   # identity to 2e-22: omitted
    # MyRotation(angle) == torch.rot90(k = angle / 90): the direction the kernel and the oracle rotate in
    assert gold["rotation"] == {"0": 0, "90": 1, "180": 2, "270": 3}
    assert gold["custom_transform_returns_its_argument"]       # 'c_0_05_bnfrsc' never reaches the Cutout append


def test_cutout_follows_the_reference_class():
    """Hole geometry against Cutout.__call__ of the reference run on seeded draws (transform_goldens.json 'cutout'), the
    oracle's masking against the same, and the sampler's packed words."""
    gold = _gold()["cutout"]
    assert len(gold) >= 40
    for case in gold:
        h, w, L = case["h"], case["w"], case["length"]
        if h == w:
            y1, y2, x1, x2 = (int(v) for v in aug.cutout_box(case["y"], case["x"], L, h))
        else:
            y1, y2, _, _ = (int(v) for v in aug.cutout_box(case["y"], case["x"], L, h))
            _, _, x1, x2 = (int(v) for v in aug.cutout_box(case["y"], case["x"], L, w))
        assert [y1, y2, x1, x2] == case["box"], case
        img = torch.full((3, h, w), -1.5)
        out = oaug.cutout(img, y1 | (y2 << 16), x1 | (x2 << 16))
        want = img.clone()
        want[:, case["box"][0]:case["box"][1], case["box"][2]:case["box"][3]] = 0.0
        assert torch.equal(out, want)
    assert torch.equal(oaug.cutout(img, 0, 0), img)                                   # no hole drawn: untouched
    a = b200ssl.MultiCropAugment("pcbnfrsc")
    p = a.sample_params(200, torch.Generator().manual_seed(3))
    for lo, hi, S in ((0, 2, 224), (2, 12, 96)):
        L = aug.CUTOUT_LENGTH * S // 256
        wy, wx = p[:, lo:hi, 11].long(), p[:, lo:hi, 12].long()
        for wv in (wy, wx):
            c1, c2 = wv & 0xFFFF, wv >> 16
            assert (c1 >= 0).all() and (c2 <= S).all() and (c2 > c1).all()
            assert ((c2 - c1) <= 2 * (L // 2)).all() and ((c2 - c1) >= L // 2).all()
            assert int((c2 - c1).max()) == 2 * (L // 2)                               # interior holes are full size
            assert bool(((c1 == 0) | (c2 == S)).any())                                # and some are clipped at a border
    assert (b200ssl.MultiCropAugment("pcbnfrs").sample_params(8)[..., 11:13] == 0).all()
    # eval mode keeps the hole (the reference appends it outside its `train` branch) and nothing else
    e = b200ssl.MultiCropAugment("cbnfrsc", train=False)
    pe = e.sample_params(8, torch.Generator().manual_seed(1))
    assert e.ops == "" and (pe[..., 11] != 0).all() and (pe[..., 4] == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("ttype", ["pcbnfrs", "cbnfr", "flip", "none", "pcbnfrsc", "bnfrsc"])
def test_kernel_matches_oracle(ttype):
    dev = torch.device("cuda")
    B = 5
    tiles = _tiles(B, seed=3)
    a = b200ssl.MultiCropAugment(ttype, color_param=0.1)
    p = a.sample_params(B, torch.Generator().manual_seed(4))
    crops = a(tiles.to(dev), params=p)
    torch.cuda.synchronize()
    ref = oaug.multicrop_augment(tiles, p, 2, 10, 224, 96, aug.MEAN["Ron"], aug.STD["Ron"])
    assert len(crops) == 12
    for c, (x, y) in enumerate(zip(crops, ref)):
        assert x.dtype == torch.bfloat16 and tuple(x.shape) == tuple(y.shape)
        x = x.float().cpu()
        rel = float((x - y).norm() / y.norm())
        assert rel < 1e-2, (ttype, c, rel)                                   # bf16 output: 2^-9 per element
        # beyond the bf16 rounding of the output nothing may differ: compare against the oracle rounded the same way
        exact = float((x - y.bfloat16().float()).abs().max())
        assert exact < 0.08, (ttype, c, exact)                               # <= a few bf16 ulps at |x| <= 8
        if a.cutout:                                                         # the hole is exactly zero, and it is there
            y1, y2, x1, x2 = (int(p[0, c, 11]) & 0xFFFF, int(p[0, c, 11]) >> 16, int(p[0, c, 12]) & 0xFFFF,
                              int(p[0, c, 12]) >> 16)
            assert y2 > y1 and x2 > x1 and (x[0, :, y1:y2, x1:x2] == 0).all()
            assert int((x[0] == 0).all(0).sum()) == (y2 - y1) * (x2 - x1)
    # crop-major layout: equal-size crops are views of one allocation, back to back (no concatenation copy downstream)
    assert crops[1].data_ptr() == crops[0].data_ptr() + crops[0].numel() * 2
    assert crops[3].data_ptr() == crops[2].data_ptr() + crops[2].numel() * 2


@pytest.mark.gpu
def test_identity_parameters_reproduce_the_tile():
    """Whole-tile box at the tile's own size, no operators: the output is the normalised tile, exactly (bf16-rounded)."""
    dev = torch.device("cuda")
    tiles = _tiles(2, seed=5)
    a = b200ssl.MultiCropAugment("none", global_size=256, local_size=256, n_global=1, n_local=1)
    p = torch.zeros(2, 2, aug.PARAM_WORDS, dtype=torch.int32)
    p[..., 2] = 256
    p[..., 3] = 256
    p.view(torch.float32)[..., 5:8] = 1.0
    out = a(tiles.to(dev), params=p)
    m = torch.tensor(aug.MEAN["Ron"]).view(1, 3, 1, 1)
    s = torch.tensor(aug.STD["Ron"]).view(1, 3, 1, 1)
    want = ((tiles.permute(0, 3, 1, 2).float() / 255 - m) / s)
    for x in out:
        assert float((x.float().cpu() - want.bfloat16().float()).abs().max()) <= 0.0625   # one bf16 ulp at |x| in [4, 8)
        assert float((x.float().cpu() - want).norm() / want.norm()) < 4e-3


@pytest.mark.gpu
def test_augment_feeds_the_training_step():
    """uint8 tiles -> MultiCropAugment -> MultiCropWrapper forward: the crops arrive in the layout the packed
    multi-crop pass consumes (global crops back to back, local crops back to back)."""
    dev = torch.device("cuda")
    tiles = _tiles(4, seed=6).to(dev)
    a = b200ssl.MultiCropAugment("pcbnfrs")
    crops = a(tiles, generator=torch.Generator().manual_seed(7))
    model = b200ssl.MultiCropWrapper(b200ssl.vit_tiny(), b200ssl.DINOHead(192, 512, hidden_dim=128, bottleneck_dim=64)).to(dev)
    with torch.no_grad():
        out = model(crops)
    assert tuple(out.shape) == (12 * 4, 512) and torch.isfinite(out.float()).all()


@pytest.mark.gpu
def test_augment_writes_into_the_graphed_step_inputs():
    """GraphedDinoStep.crop_blocks() exposes the static inputs per resolution group; the augmentation kernel fills them
    in place (no intermediate crops, no copy) and the replay consumes them."""
    dev = torch.device("cuda")
    B = 2
    tiles = _tiles(B, seed=8).to(dev)
    a = b200ssl.MultiCropAugment("pcbnfrs", n_local=2)
    model = b200ssl.MultiCropWrapper(b200ssl.vit_tiny(), b200ssl.DINOHead(192, 512, hidden_dim=128, bottleneck_dim=64)).to(dev)
    teacher = b200ssl.ModelEma(model)
    loss_fn = b200ssl.DINOLoss(512, 4, 0.04, 0.04, 0, 10).to(dev)
    opt = b200ssl.FusedAdamW(b200ssl.param_groups_wd(model, 0.04), lr=1e-4)
    example = [torch.zeros(B, 3, 224, 224, device=dev, dtype=torch.bfloat16) for _ in range(2)] + \
              [torch.zeros(B, 3, 96, 96, device=dev, dtype=torch.bfloat16) for _ in range(2)]
    step = b200ssl.GraphedDinoStep(model, teacher, loss_fn, opt, example)
    blocks = step.crop_blocks()
    assert [tuple(b.shape) for b in blocks] == [(2, B, 3, 224, 224), (2, B, 3, 96, 96)]
    assert blocks[0][1].data_ptr() == step.static_crops[1].data_ptr()
    assert blocks[1][0].data_ptr() == step.static_crops[2].data_ptr()
    p = a.sample_params(B, torch.Generator().manual_seed(9))
    crops = a(tiles, params=p, out=blocks)
    ref = a(tiles, params=p)
    for c in range(4):
        assert crops[c].data_ptr() == step.static_crops[c].data_ptr()
        assert torch.equal(step.static_crops[c], ref[c])
    loss = step(None, epoch=0, momentum=0.99)
    assert torch.isfinite(loss).all()
    step.release()


@pytest.mark.gpu
def test_augment_rejects_bad_input():
    a = b200ssl.MultiCropAugment()
    with pytest.raises(RuntimeError, match="CUDA|cuda"):
        a(_tiles(1))
    with pytest.raises(RuntimeError, match="uint8"):
        a(torch.zeros(1, 256, 256, 3, device="cuda"))
