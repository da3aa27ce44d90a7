"""Pins ``oracle/vision_transformer.py`` (and the initialisation of the drop-in modules) to outputs of the
reference's OWN code.

``tests/golden/vt_goldens.pt`` was produced by executing the reference's CPython-3.7 bytecode
(nn_encoder_arch/__pycache__/vision_transformer.cpython-37.pyc) with ``tests/golden/py37vm.py`` — see
``tests/golden/make_vt_goldens.py``. The oracle is driven through exactly the same calls and seeds here.

In the build container (where /root/reference exists) the comparison is also made live, bit for bit; against the
committed fixture a tolerance of a few fp32 ulps absorbs CPU-kernel differences between hosts.
"""
import os
import sys
import warnings

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, os.path.dirname(HERE))

import make_vt_goldens as mk  # noqa: E402
from oracle import vision_transformer as oracle_vt  # noqa: E402

GOLD_PATH = os.path.join(HERE, "golden", "vt_goldens.pt")
RTOL, ATOL = 2e-5, 2e-6     # fp32 on a different host CPU; the live comparison below is exact


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD_PATH, weights_only=False)


def close(a, b, what):
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs().max().item() if a.numel() else 0.0
    assert torch.allclose(a, b, rtol=RTOL, atol=ATOL), f"{what}: max |diff| {err:.3e}"


def compare_small(rec, ref, cmp):
    assert list(rec["state_dict"]) == list(ref["state_dict"])
    for k in ref["state_dict"]:
        cmp(rec["state_dict"][k], ref["state_dict"][k], f"init {k}")
    for k in ("forward", "last_selfattention", "tokens", "pos_40x24", "pos_native", "train_forward"):
        cmp(rec[k], ref[k], k)
    assert len(rec["intermediate"]) == len(ref["intermediate"]) == 2
    for i, (a, b) in enumerate(zip(rec["intermediate"], ref["intermediate"])):
        cmp(a, b, f"intermediate[{i}]")
    assert sorted(rec["grads"]) == sorted(ref["grads"])
    for k in ref["grads"]:
        cmp(rec["grads"][k], ref["grads"][k], f"grad {k}")


def compare_head(rec, ref, cmp):
    assert list(rec["state_dict"]) == list(ref["state_dict"])
    for k in ref["state_dict"]:
        cmp(rec["state_dict"][k], ref["state_dict"][k], f"head init {k}")
    assert rec["requires_grad"] == ref["requires_grad"]
    cmp(rec["y"], ref["y"], "head output")
    assert sorted(rec["grads"]) == sorted(ref["grads"])
    for k in ref["grads"]:
        cmp(rec["grads"][k], ref["grads"][k], f"head grad {k}")


def compare_fingerprint(fp, ref):
    assert list(fp) == list(ref)
    for k in ref:
        shape, s, a, head = fp[k]
        rshape, rs, ra, rhead = ref[k]
        assert shape == rshape, k
        assert abs(s - rs) <= 1e-6 * max(1.0, ra), f"{k}: sum {s} vs {rs}"
        assert abs(a - ra) <= 1e-6 * max(1.0, ra), f"{k}: abs-sum {a} vs {ra}"
        close(head, rhead, f"{k}[:4]")


# ---- committed fixture (runs anywhere) -------------------------------------------------------------------------
def test_small_encoder_matches_reference_outputs(gold):
    rec, _ = mk.small_case(oracle_vt)
    compare_small(rec, gold["small"], close)
    # the non-square input really went through the bicubic position-table resize, and stochastic depth was live
    assert gold["small"]["pos_40x24"].shape == (1, 1 + 5 * 3, 64)
    assert not torch.allclose(gold["small"]["train_forward"], gold["small"]["forward"], atol=1e-4)


def test_small_encoder_from_reference_weights(gold):
    """Loading the reference's weights (rather than re-drawing them) isolates the forward arithmetic from init."""
    from functools import partial
    model = oracle_vt.VisionTransformer(norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **gold["small_kw"])
    model.load_state_dict(gold["small"]["state_dict"], strict=True)
    model.eval()
    with torch.no_grad():
        close(model(gold["small"]["x"]), gold["small"]["forward"], "forward")
        close(model.get_last_selfattention(gold["small"]["x"]), gold["small"]["last_selfattention"], "attention")


@pytest.mark.parametrize("case", [c[0] for c in mk.HEAD_CASES])
def test_head_matches_reference_outputs(gold, case):
    _, i, o, kw = next(c for c in mk.HEAD_CASES if c[0] == case)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rec = mk.head_case(oracle_vt, i, o, kw)
    compare_head(rec, gold["heads"][case], close)


@pytest.mark.parametrize("name", ["vit_tiny", "vit_small", "vit_base"])
def test_factories_match_reference_initialisation(gold, name):
    ref = gold["factory"][name]
    rec = mk.factory_case(oracle_vt, name, with_output="forward" in ref)
    compare_fingerprint(rec["fingerprint"], ref["fingerprint"])
    if "forward" in ref:
        close(rec["forward"], ref["forward"], f"{name} forward")


def test_functions_match_reference_outputs(gold):
    rec = mk.fn_cases(oracle_vt)
    assert sorted(rec) == sorted(gold["fns"])
    for k in gold["fns"]:
        close(rec[k], gold["fns"][k], k)


def test_hot_path_shapes_match_reference_outputs(gold):
    """ViT-S/16 + head on 224x224 and 96x96 tiles (the bench's crop shapes): the vectors the CUDA path is compared
    with in tests/test_gpu_model.py::test_cuda_path_matches_reference_golden."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rec = mk.hot_case(oracle_vt)
    ref = gold["hot"]
    close(rec["features"], ref["features"], "features")
    close(rec["logits"], ref["logits"], "logits")
    assert sorted(rec["grads"]) == sorted(ref["grads"]) and len(ref["grads"]) >= 30
    for k in ref["grads"]:
        a, b = rec["grads"][k], ref["grads"][k]
        assert a.shape == b.shape, k
        assert (a - b).norm() <= 1e-4 * b.norm() + 1e-7, f"grad {k}: {(a - b).norm().item():.3e} of {b.norm().item():.3e}"


@pytest.mark.parametrize("name", ["vit_tiny", "vit_small"])
def test_dropin_modules_initialise_like_the_reference(gold, name):
    """The product-side modules (host construction only, no kernels) draw the same initial weights, under the
    same names, as the reference for the same seed — a checkpoint of either loads into the other."""
    import b200ssl  # noqa: F401
    from b200ssl import vision_transformer as prod
    torch.manual_seed(0)
    model = getattr(prod, name)(drop_path_rate=0.1)
    compare_fingerprint(mk.fingerprint(model.state_dict()), gold["factory"][name]["fingerprint"])


def test_dropin_head_initialises_like_the_reference(gold):
    import b200ssl  # noqa: F401
    from b200ssl import vision_transformer as prod
    _, i, o, kw = mk.HEAD_CASES[0]
    torch.manual_seed(1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        head = prod.DINOHead(i, o, **kw)
    ref = gold["heads"]["default3"]
    assert list(head.state_dict()) == list(ref["state_dict"])
    for k, v in head.state_dict().items():
        close(v, ref["state_dict"][k], k)
    assert {k: p.requires_grad for k, p in head.named_parameters()} == ref["requires_grad"]


# ---- live, bit-exact (build container only) --------------------------------------------------------------------
@pytest.mark.skipif(not os.path.exists(mk.PYC), reason="the reference tree is only present in the build container")
def test_oracle_is_bit_exact_against_the_running_reference():
    def exact(a, b, what):
        assert torch.equal(a, b), f"{what}: max |diff| {(a - b).abs().max().item():.3e}"

    import py37vm
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = mk._NS(py37vm.load_module(mk.PYC, "reference_vision_transformer_live"))
        compare_small(mk.small_case(oracle_vt)[0], mk.small_case(ref)[0], exact)
        for _, i, o, kw in mk.HEAD_CASES:
            compare_head(mk.head_case(oracle_vt, i, o, kw), mk.head_case(ref, i, o, kw), exact)
        a, b = mk.fn_cases(oracle_vt), mk.fn_cases(ref)
        for k in b:
            exact(a[k], b[k], k)
        # the default factory at the bench's encoder size, training mode with stochastic depth, gradients included
        outs = []
        for vt in (ref, oracle_vt):
            torch.manual_seed(0)
            m = vt.vit_tiny(drop_path_rate=0.1)
            m.train()
            torch.manual_seed(5)
            x = torch.randn(2, 3, 64, 48, generator=torch.Generator().manual_seed(9))
            y = m(x)
            y.sum().backward()
            outs.append((y.detach(), {k: p.grad for k, p in m.named_parameters() if p.grad is not None}))
        exact(outs[0][0], outs[1][0], "vit_tiny train forward")
        assert sorted(outs[0][1]) == sorted(outs[1][1])
        for k in outs[0][1]:
            exact(outs[0][1][k], outs[1][1][k], f"vit_tiny grad {k}")
