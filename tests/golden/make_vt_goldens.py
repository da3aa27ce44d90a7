"""Generates tests/golden/vt_goldens.pt by EXECUTING the reference's own encoder / head code.

The reference's ``nn_encoder_arch/vision_transformer`` exists only as CPython-3.7 bytecode
(/root/reference/nn_encoder_arch/__pycache__/vision_transformer.cpython-37.pyc); ``py37vm`` runs that bytecode
unmodified (its classes become real ``nn.Module`` subclasses on this container's torch), so every tensor below is an
output of the reference itself, not of any restatement. tests/test_oracle_golden.py then pins ``oracle/`` (and the
initialisation of the drop-in modules) to these vectors.

Run in the build container only (needs /root/reference):   python tests/golden/make_vt_goldens.py
The fixture holds, all in fp32 on the CPU:
  small    a 2-block encoder (patch 8, dim 64) : full state_dict, a non-square input that exercises the position-table
           interpolation, eval outputs (forward, get_last_selfattention, get_intermediate_layers, prepare_tokens,
           interpolate_pos_encoding), a train-mode forward under a fixed seed (stochastic depth) and every gradient
  heads    three DINOHead configurations: state_dict, input, output, requires_grad flags
  factory  vit_tiny / vit_small / vit_base at seed 0: per-tensor fingerprints of the initial weights and, for
           tiny / small, the CLS output of a seeded 2x3x32x48 input
  fns      trunc_normal_ and drop_path on seeded inputs
  hot      the bench's encoder + head (vit_small/16, DINOHead 384 -> 1024) on 2 global 224x224 and 3 local 96x96
           seeded tiles: CLS features, head logits and the gradients of a fixed linear functional of the logits for a
           subset of parameters (all 1-D tensors of blocks 0/5/11, cls_token, final norm, patch-embed bias, head
           biases, leading rows of four weight matrices). Weights are re-drawn from the seed by the consumer (the
           `factory` fingerprints prove they are the same), so the fixture stays small. tests/test_gpu_model.py
           compares the CUDA path with these directly.
"""
import os
import sys
import warnings
from functools import partial

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import py37vm  # noqa: E402

PYC = "/root/reference/nn_encoder_arch/__pycache__/vision_transformer.cpython-37.pyc"

SMALL_KW = dict(img_size=[32], patch_size=8, embed_dim=64, depth=2, num_heads=2, mlp_ratio=2.0, qkv_bias=True,
                drop_path_rate=0.2)
HEAD_CASES = [
    ("default3", 64, 48, dict(hidden_dim=96, bottleneck_dim=32)),
    ("bn2_free_last", 64, 40, dict(use_bn=True, nlayers=2, hidden_dim=80, bottleneck_dim=24, norm_last_layer=False)),
    ("single", 64, 33, dict(nlayers=1, bottleneck_dim=16)),
]


def fingerprint(sd):
    """name -> (shape, sum, abs-sum in fp64, first four values)."""
    out = {}
    for k, v in sd.items():
        f = v.detach().double().flatten()
        out[k] = (tuple(v.shape), f.sum().item(), f.abs().sum().item(), f[:4].float().clone())
    return out


def small_case(vt, seed=0):
    """Everything recorded for the small encoder; ``vt`` is a module namespace exposing VisionTransformer.
    Shared with the test so that the oracle is driven through exactly the same calls."""
    torch.manual_seed(seed)
    model = vt.VisionTransformer(norm_layer=partial(nn.LayerNorm, eps=1e-6), **SMALL_KW)
    gen = torch.Generator().manual_seed(100)
    x = torch.randn(3, 3, 40, 24, generator=gen)
    rec = {"state_dict": {k: v.clone() for k, v in model.state_dict().items()}, "x": x}
    model.eval()
    with torch.no_grad():
        rec["forward"] = model(x)
        rec["last_selfattention"] = model.get_last_selfattention(x)
        rec["intermediate"] = list(model.get_intermediate_layers(x, 2))
        rec["tokens"] = model.prepare_tokens(x)
        rec["pos_40x24"] = model.interpolate_pos_encoding(rec["tokens"], 40, 24)
        rec["pos_native"] = model.interpolate_pos_encoding(torch.zeros(1, 17, 64), 32, 32)
    model.train()
    torch.manual_seed(7)
    y = model(x)
    (y * torch.linspace(-1, 1, y.numel()).view_as(y)).sum().backward()
    rec["train_forward"] = y.detach()
    rec["grads"] = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    return rec, model


def head_case(vt, in_dim, out_dim, kw, seed=1):
    torch.manual_seed(seed)
    head = vt.DINOHead(in_dim, out_dim, **kw)
    gen = torch.Generator().manual_seed(200)
    x = torch.randn(6, in_dim, generator=gen)
    y = head(x)
    y.square().sum().backward()
    return {"state_dict": {k: v.clone() for k, v in head.state_dict().items()}, "x": x, "y": y.detach(),
            "requires_grad": {k: p.requires_grad for k, p in head.named_parameters()},
            "grads": {k: p.grad.clone() for k, p in head.named_parameters() if p.grad is not None}}


def factory_case(vt, name, seed=0, with_output=True):
    torch.manual_seed(seed)
    model = getattr(vt, name)(drop_path_rate=0.1)
    rec = {"fingerprint": fingerprint(model.state_dict())}
    if with_output:
        gen = torch.Generator().manual_seed(300)
        x = torch.randn(2, 3, 32, 48, generator=gen)
        model.eval()
        with torch.no_grad():
            rec["x"], rec["forward"] = x, model(x)
    return rec


HOT_OUT_DIM = 1024
HOT_GRAD_ROWS = ("backbone.blocks.0.attn.qkv.weight", "backbone.blocks.5.attn.proj.weight",
                 "backbone.blocks.11.mlp.fc1.weight", "backbone.blocks.11.mlp.fc2.weight", "head.mlp.0.weight",
                 "head.last_layer.weight_v")


def hot_modules(vt, seed=0):
    """(backbone, head) exactly as the hot-path golden was produced; shared with the tests."""
    torch.manual_seed(seed)
    backbone = vt.vit_small()
    head = vt.DINOHead(384, HOT_OUT_DIM)
    return backbone, head


def hot_inputs():
    xg = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(400))
    xl = torch.randn(3, 3, 96, 96, generator=torch.Generator().manual_seed(401))
    w = torch.randn(5, HOT_OUT_DIM, generator=torch.Generator().manual_seed(402))
    return xg, xl, w


def hot_grad_selected(name):
    if name in HOT_GRAD_ROWS:
        return True
    if any(name.startswith(f"backbone.blocks.{i}.") for i in (0, 5, 11)) and (name.endswith("bias") or ".norm" in name):
        return True
    return name in ("backbone.cls_token", "backbone.norm.weight", "backbone.norm.bias",
                    "backbone.patch_embed.proj.bias", "head.mlp.0.bias", "head.mlp.2.bias", "head.mlp.4.bias")


def hot_case(vt):
    backbone, head = hot_modules(vt)
    xg, xl, w = hot_inputs()
    backbone.train()
    head.train()
    feats = torch.cat((backbone(xg), backbone(xl)))
    logits = head(feats)
    (logits * w).sum().backward()
    grads = {}
    for prefix, mod in (("backbone.", backbone), ("head.", head)):
        for k, p in mod.named_parameters():
            name = prefix + k
            if p.grad is not None and hot_grad_selected(name):
                grads[name] = (p.grad[:8] if name in HOT_GRAD_ROWS else p.grad).clone()
    return {"features": feats.detach(), "logits": logits.detach(), "grads": grads}


def fn_cases(vt):
    out = {}
    torch.manual_seed(11)
    out["trunc_default"] = vt.trunc_normal_(torch.empty(257), std=0.02).clone()
    torch.manual_seed(12)
    out["trunc_shifted"] = vt.trunc_normal_(torch.empty(64, 5), mean=0.5, std=1.0, a=-1.0, b=2.0).clone()
    gen = torch.Generator().manual_seed(13)
    x = torch.randn(8, 5, 4, generator=gen)
    torch.manual_seed(14)
    out["drop_path_x"] = x
    out["drop_path_train"] = vt.drop_path(x, 0.3, True)
    out["drop_path_eval"] = vt.drop_path(x, 0.3, False)
    out["drop_path_zero"] = vt.drop_path(x, 0.0, True)
    return out


class _NS:
    def __init__(self, d):
        self.__dict__.update(d)


def generate():
    ref = _NS(py37vm.load_module(PYC, "reference_vision_transformer"))
    gold = {"source": PYC, "torch": torch.__version__, "small_kw": SMALL_KW}
    gold["small"], _ = small_case(ref)
    gold["heads"] = {name: head_case(ref, i, o, kw) for name, i, o, kw in HEAD_CASES}
    gold["factory"] = {n: factory_case(ref, n, with_output=(n != "vit_base")) for n in ("vit_tiny", "vit_small", "vit_base")}
    gold["fns"] = fn_cases(ref)
    gold["hot"] = hot_case(ref)
    return gold


if __name__ == "__main__":
    warnings.filterwarnings("ignore")
    g = generate()
    dst = os.path.join(HERE, "vt_goldens.pt")
    torch.save(g, dst)
    print(f"wrote {dst}: {os.path.getsize(dst) / 1024:.0f} KiB")
