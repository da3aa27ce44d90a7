"""Element dropout on the encoder path: ``drop_rate`` > 0 (train.py --drop, train.py:283,487) -> ``pos_drop``
(VT.pyc@L196,245), ``Attention.proj_drop`` (@L117,130) and ``Mlp.drop`` behind the activation and behind fc2
(@L96,101-104).

torch draws its dropout masks from the global Philox stream, which another implementation cannot reproduce, so the
comparison is made the other way round: the product's counter-based mask generator (csrc/dropout.cu) is restated in
oracle/dropout.py, the oracle model's ``nn.Dropout`` modules are swapped for ``ReplayDropout`` carrying the seed the
product drew, and the two models are then compared element for element at the usual gate (output rel <= 1e-2, every
parameter-gradient cosine >= 0.999)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dropout as odrop  # noqa: E402


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item()


def test_oracle_mask_is_a_pure_function_with_the_right_rate():
    n = 1 << 18
    for p in (0.05, 0.1, 0.5, 0.9):
        m = odrop.keep_mask(12345, 4, n, p)
        assert abs(float(m.float().mean()) - (1 - p)) < 4e-3, p
        assert torch.equal(m, odrop.keep_mask(12345, 4, n, p))
        assert torch.equal(m[:1000], odrop.keep_mask(12345, 4, 1000, p))            # index i does not depend on n
        assert not torch.equal(m, odrop.keep_mask(12345, 5, n, p))                  # another site, another mask
        assert not torch.equal(m, odrop.keep_mask(12346, 4, n, p))                  # another seed, another mask
    assert odrop.keep_mask(1, 0, 64, 0.0).all() and not odrop.keep_mask(1, 0, 64, 1.0).any()
    # neighbouring elements are independent: the four 16-bit fields of one word do not move together
    m = odrop.keep_mask(99, 0, n, 0.5).float().view(-1, 4)
    c = torch.corrcoef(m.t())
    assert float((c - torch.eye(4)).abs().max()) < 0.02
    # the nn.Dropout stand-in: identity in eval mode, x * mask / (1 - p) in training mode, sites walked call by call
    d = odrop.ReplayDropout(0.25, 7, [2, 3])
    x = torch.ones(2, 8, 16)
    d.eval()
    assert d(x) is x
    d.train()
    y1, y2 = d(x), d(x)
    assert torch.equal(y1 != 0, odrop.keep_mask(7, 2, 256, 0.25).view(2, 8, 16))
    assert torch.equal(y2 != 0, odrop.keep_mask(7, 3, 256, 0.25).view(2, 8, 16))
    assert float(y1.max()) == pytest.approx(1 / 0.75)


@pytest.mark.gpu
def test_dropout_kernels_match_the_oracle_mask(cuda_device):
    import b200ssl
    ops = b200ssl.ops
    seed = torch.tensor([0x1234_5678_9ABC_DEF0 >> 2], dtype=torch.int64, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(3)
    rows, D, p, site = 77, 384, 0.3, 5
    keep = odrop.keep_mask(int(seed.item()), site, rows * D, p).view(rows, D).cuda()
    scale = 1.0 / (1.0 - p)
    x32 = torch.randn(rows, D, device="cuda", generator=g)
    x16 = x32.bfloat16()
    second = torch.randn(rows, D, device="cuda", generator=g).bfloat16()
    # fp32, out of place
    y32 = ops.dropout(x32, (p, seed), site)
    assert torch.equal(y32 != 0, keep & (x32 != 0))
    assert torch.allclose(y32, x32 * keep * scale, rtol=1e-6, atol=0)
    # bf16 with the second tensor, in place
    a, b = x16.clone(), second.clone()
    out = ops.dropout(a, (p, seed), site, second=b, out=a)
    assert out.data_ptr() == a.data_ptr()
    for got, src in ((a, x16), (b, second)):                 # exact mask, values to the bf16 rounding of x * 1 / (1 - p)
        assert torch.equal(got != 0, keep & (src != 0))
        assert torch.allclose(got.float(), src.float() * keep * scale, rtol=4e-3, atol=0)
    # residual add with and without the stochastic-depth row scale
    rs = (torch.rand(rows, device="cuda", generator=g) > 0.3).float() / 0.7
    for r in (None, rs):
        y = ops.dropout_residual(x16, x32, r, (p, seed), site)
        want = x32 + (1.0 if r is None else r[:, None]) * (x16.float() * keep * scale)
        assert torch.allclose(y, want, rtol=1e-6, atol=1e-6)
    # another site / another seed: another mask; p = 0 keeps everything
    assert not torch.equal(ops.dropout(x32, (p, seed), site + 1) != 0, y32 != 0)
    assert not torch.equal(ops.dropout(x32, (p, seed + 1), site) != 0, y32 != 0)
    assert torch.equal(ops.dropout(x32, (0.0, seed), site), x32)
    with pytest.raises(RuntimeError):
        ops.dropout(x32[:, :7].contiguous(), (p, seed), site)       # 77 * 7 elements: not a multiple of 8
    with pytest.raises(RuntimeError):
        ops.dropout(x32, (1.5, seed), site)


@pytest.mark.gpu
@pytest.mark.parametrize("drop_path,size,B,ckpt", [(0.0, 224, 3, False), (0.2, 96, 8, False), (0.0, 96, 4, True)])
def test_vit_with_drop_rate_matches_oracle_replaying_the_mask(cuda_device, drop_path, size, B, ckpt):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    import b200ssl
    from oracle import vision_transformer as ovt
    p = 0.1
    torch.manual_seed(0)
    ref = ovt.vit_tiny(drop_rate=p, drop_path_rate=drop_path).cuda()
    with torch.no_grad():
        for q in ref.parameters():
            if q.ndim == 1:
                q.add_(torch.randn_like(q) * 0.02)
    mine = b200ssl.vit_tiny(drop_rate=p, drop_path_rate=drop_path).cuda()
    mine.load_state_dict(ref.state_dict())
    mine.set_grad_checkpointing(ckpt)
    ref.train(), mine.train()
    x = torch.randn(B, 3, size, size, device="cuda", generator=torch.Generator(device="cuda").manual_seed(size))
    w = torch.randn(B, 192, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    torch.manual_seed(11)
    out = mine(x.bfloat16())
    (out.float() * w).sum().backward()
    seed = int(mine._last_dropout_seed.item())
    odrop.replay_in(ref, seed)
    torch.manual_seed(11)                       # the stochastic-depth masks come from torch's stream in both models
    out_ref = ref(x)
    (out_ref * w).sum().backward()
    assert rel(out, out_ref) < 1e-2, rel(out, out_ref)
    bad = [(n, cos(q.grad, r.grad)) for (n, r), (_, q) in zip(ref.named_parameters(), mine.named_parameters())
           if cos(q.grad, r.grad) < 0.999]
    assert not bad, bad
    # the masks really dropped something: without replaying them the oracle lands somewhere else ...
    plain = ovt.vit_tiny(drop_rate=0.0, drop_path_rate=0.0).cuda().train()
    plain.load_state_dict(ref.state_dict())
    assert rel(out, plain(x)) > 3e-2
    # ... a call from another generator state draws another seed ...
    torch.manual_seed(12)
    out2 = mine(x.bfloat16())
    assert int(mine._last_dropout_seed.item()) != seed and rel(out2, out) > 1e-2
    # ... and eval mode is mask free
    mine.eval(), plain.eval()
    assert rel(mine(x.bfloat16()), plain(x)) < 1e-2


@pytest.mark.gpu
def test_dropout_on_the_other_entry_points(cuda_device):
    """forward_multi (the packed multi-crop pass), get_intermediate_layers / prepare_tokens (pos_drop as its own node,
    blocks one by one), Block.forward on its own, and what still refuses: attn_drop."""
    import b200ssl
    torch.manual_seed(0)
    m = b200ssl.vit_tiny(drop_rate=0.2).cuda().train()
    g = torch.Generator(device="cuda").manual_seed(2)
    xg = torch.randn(2, 3, 224, 224, device="cuda", generator=g).bfloat16()
    xl = torch.randn(4, 3, 96, 96, device="cuda", generator=g).bfloat16()
    out = m.forward_multi([xg, xl])
    assert out.shape == (6, 192) and torch.isfinite(out.float()).all()
    out.float().sum().backward()
    assert all(q.grad is not None and torch.isfinite(q.grad).all() for q in m.parameters())
    # seeded calls reproduce: same torch seed -> same dropout seed -> identical output
    torch.manual_seed(5)
    a = m(xg)
    torch.manual_seed(5)
    assert torch.equal(a, m(xg))
    tok = m.prepare_tokens(xg.float())
    assert tok.shape == (2, 197, 192)
    frac = float((tok == 0).float().mean())
    assert 0.17 < frac < 0.23, frac                                   # pos_drop really masks 20 % of the token stream
    layers = m.get_intermediate_layers(xg, n=2)
    assert len(layers) == 2 and layers[0].shape == (2, 197, 192)
    blk = m.blocks[3]
    y = blk(tok)
    assert y.shape == tok.shape and torch.isfinite(y.float()).all()
    attn = m.get_last_selfattention(xg)                               # inspection path: works in training mode too
    assert attn.shape == (2, 3, 197, 197)
    m.eval()
    assert torch.isfinite(blk.mlp(tok).float()).all()                 # eval mode: nn.Dropout is the identity
    m.train()
    with pytest.raises(NotImplementedError):
        b200ssl.vit_tiny(attn_drop_rate=0.1).cuda().train()(xg)


@pytest.mark.gpu
def test_graphed_step_draws_a_fresh_mask_on_every_replay(cuda_device):
    """The whole DINO step with drop_rate > 0 as a CUDA-graph replay: the seed is drawn on the device inside the
    captured graph (torch's CUDA generator is a graph input), so every replay sees a new mask; the teacher (eval mode)
    is mask free. Eager and graphed steps from the same generator state agree."""
    import b200ssl
    out_dim, ncrops, B = 1024, 4, 4
    g = torch.Generator(device="cuda").manual_seed(5)
    crops = [torch.randn(B, 3, 224, 224, device="cuda", generator=g).bfloat16() for _ in range(2)] + \
            [torch.randn(B, 3, 96, 96, device="cuda", generator=g).bfloat16() for _ in range(ncrops - 2)]

    def build():
        torch.manual_seed(0)
        mod = b200ssl.MultiCropWrapper(b200ssl.vit_tiny(drop_rate=0.1, drop_path_rate=0.1),
                                       b200ssl.DINOHead(192, out_dim, hidden_dim=256, bottleneck_dim=64)).cuda()
        teacher = b200ssl.ModelEma(mod)
        loss_fn = b200ssl.DINOLoss(out_dim, ncrops, 0.04, 0.04, 0, 10).cuda()
        opt = b200ssl.FusedAdamW(b200ssl.param_groups_wd(mod, 0.04), lr=1e-3)
        return mod, teacher, loss_fn, opt

    mod, teacher, loss_fn, opt = build()
    step = b200ssl.GraphedDinoStep(b200ssl.GradBucketDataParallel(mod), teacher, loss_fn, opt, crops)
    seeds, losses = [], []
    for i in range(4):
        loss = step(crops if i == 0 else None, momentum=0.99)
        torch.cuda.synchronize()
        seeds.append(int(mod.backbone._last_dropout_seed.item()))
        losses.append(float(loss))
    assert len(set(seeds)) == 4, seeds
    assert all(torch.isfinite(torch.tensor(losses))), losses
    assert all(torch.isfinite(q).all() for q in mod.parameters())
    step.release()
    # eager step 0 from the same generator state as the graph's step 0: same masks, same loss (to bf16 / atomics noise)
    mod2, teacher2, loss_fn2, opt2 = build()
    l2, _, _ = b200ssl.dino_step(mod2, teacher2, loss_fn2, opt2, crops, momentum=0.99)
    assert int(mod2.backbone._last_dropout_seed.item()) == seeds[0]
    assert abs(float(l2) - losses[0]) / abs(losses[0]) < 2e-3, (float(l2), losses[0])


@pytest.mark.gpu
def test_bare_mlp_and_attention_modules_with_dropout(cuda_device):
    """Mlp(drop=p) and Attention(proj_drop=p) called on their own in training mode (VT.pyc@L98-104, @L119-131): against
    the oracle modules replaying the mask (Mlp.drop: sites 0 and 1; proj_drop: site 0)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    import b200ssl
    from oracle import vision_transformer as ovt
    p = 0.25
    g = torch.Generator(device="cuda").manual_seed(8)
    x = torch.randn(5, 37, 192, device="cuda", generator=g)
    w = torch.randn(5, 37, 192, device="cuda", generator=g)
    for kind in ("mlp", "attn"):
        torch.manual_seed(0)
        if kind == "mlp":
            ref, mine = ovt.Mlp(192, 768, drop=p).cuda().train(), b200ssl.Mlp(192, 768, drop=p).cuda().train()
        else:
            ref = ovt.Attention(192, num_heads=3, qkv_bias=True, proj_drop=p).cuda().train()
            mine = b200ssl.Attention(192, num_heads=3, qkv_bias=True, proj_drop=p).cuda().train()
        with torch.no_grad():
            for q in ref.parameters():
                if q.ndim == 1:
                    q.add_(torch.randn_like(q) * 0.1)
        mine.load_state_dict(ref.state_dict())
        xm = x.clone().requires_grad_(True)
        xr = x.clone().requires_grad_(True)
        out = mine(xm)
        out = out[0] if kind == "attn" else out
        (out.float() * w).sum().backward()
        seed = int(mine._last_dropout_seed.item())
        if kind == "mlp":
            ref.drop = odrop.ReplayDropout(p, seed, [0, 1])
        else:
            ref.proj_drop = odrop.ReplayDropout(p, seed, [0])
        out_ref = ref(xr)
        out_ref = out_ref[0] if kind == "attn" else out_ref
        (out_ref * w).sum().backward()
        assert 0.2 < float((out == 0).float().mean()) < 0.3, kind              # a quarter of the output is masked
        assert rel(out, out_ref) < 1e-2, (kind, rel(out, out_ref))
        assert cos(xm.grad, xr.grad) > 0.999, (kind, cos(xm.grad, xr.grad))
        bad = [(n, cos(q.grad, r.grad)) for (n, r), (_, q) in zip(ref.named_parameters(), mine.named_parameters())
               if cos(q.grad, r.grad) < 0.999]
        assert not bad, (kind, bad)
