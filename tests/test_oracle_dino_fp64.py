"""Second, independent formulation of the pieces in oracle/dino.py (which is "parity unpinned": the reference holds
no DINO loss / EMA to compare with). Here the loss, its gradient and the centre / EMA updates are re-derived in closed
form from Caron et al. 2021 (Alg. 1) and evaluated in numpy float64 with explicit loops -- no torch, no autograd, no
log_softmax -- and the oracle (which computes in fp32: it casts its inputs with .float(), as upstream does)
must agree to fp32 rounding (4e-6 relative). Closed forms:

    t_iq   = softmax((T_iq - c) / tau_t)                                   iq in {0, 1}  (global crops of the teacher)
    loss   = 1/(n_terms B) * sum_b sum_iq sum_{v != iq} [ LSE(S_v / tau_s) - <t_iq, S_v / tau_s> ],  n_terms = 2 (V - 1)
    dS_v   = ( n_v softmax(S_v / tau_s) - sum_{iq != v} t_iq ) / (tau_s B n_terms),   n_v = #{iq != v}
    c'     = m c + (1 - m) mean_rows(T)          (raw teacher logits, all ranks)
    theta' = mu theta_t + (1 - mu) theta_s,      mu_it = 1 - (1 - mu_0) (cos(pi it / T) + 1) / 2
"""
import math

import numpy as np
import pytest
import torch

from oracle import dino as odino


def _closed_form(S, T, c, V, tau_s, tau_t):
    rows, K = S.shape
    B = rows // V
    t = []
    for iq in range(2):
        z = (T[iq * B:(iq + 1) * B] - c) / tau_t
        z = z - z.max(axis=1, keepdims=True)
        e = np.exp(z)
        t.append(e / e.sum(axis=1, keepdims=True))
    n_terms = 2 * (V - 1)
    loss = 0.0
    dS = np.zeros_like(S)
    for v in range(V):
        x = S[v * B:(v + 1) * B] / tau_s
        m = x.max(axis=1, keepdims=True)
        lse = m[:, 0] + np.log(np.exp(x - m).sum(axis=1))
        p = np.exp(x - lse[:, None])
        n_v, tsum = 0, np.zeros_like(x)
        for iq in range(2):
            if iq == v:
                continue
            for b in range(B):
                loss += lse[b] - float(np.dot(t[iq][b], x[b]))
            n_v += 1
            tsum += t[iq]
        dS[v * B:(v + 1) * B] = (n_v * p - tsum) / (tau_s * B * n_terms)
    return loss / (n_terms * B), dS


@pytest.mark.parametrize("V,B,K,epoch", [(2, 3, 17, 0), (4, 2, 33, 0), (12, 2, 64, 3), (5, 4, 29, 9)])
def test_oracle_dino_loss_and_gradient_match_closed_form_fp64(V, B, K, epoch):
    rng = np.random.default_rng(V * 1000 + K)
    S = rng.normal(size=(V * B, K)) * 2.0
    T = rng.normal(size=(2 * B, K)) * 2.0
    c = rng.normal(size=(1, K)) * 0.3
    loss_fn = odino.DINOLoss(K, V, 0.04, 0.07, 5, 10).double()
    loss_fn.center = torch.from_numpy(c.copy())
    tau_t = float(loss_fn.teacher_temp_schedule[epoch])
    assert tau_t == pytest.approx(0.04 + (0.07 - 0.04) * min(epoch, 4) / 4 if epoch < 5 else 0.07)
    s = torch.from_numpy(S.copy()).requires_grad_(True)
    loss = loss_fn(s, torch.from_numpy(T.copy()), epoch)
    loss.backward()
    want, dS = _closed_form(S, T, c, V, 0.1, tau_t)
    assert abs(loss.item() - want) < 4e-6 * max(1.0, abs(want))
    assert np.abs(s.grad.numpy() - dS).max() < 4e-6 * np.abs(dS).max()
    # the centre update ran inside forward, from the RAW teacher logits
    want_c = 0.9 * c + 0.1 * T.mean(axis=0, keepdims=True)
    assert np.abs(loss_fn.center.numpy() - want_c).max() < 1e-6


def test_oracle_loss_is_shift_invariant_and_bounded_below_by_teacher_entropy():
    """Two properties no implementation detail can fake: adding a per-row constant to the student logits changes
    nothing, and a student equal to the (sharpened, centred) teacher attains the teacher's entropy, the minimum."""
    rng = np.random.default_rng(7)
    V, B, K = 2, 5, 41
    T = rng.normal(size=(2 * B, K))
    loss_fn = odino.DINOLoss(K, V, 0.04, 0.04, 0, 10).double()
    S = rng.normal(size=(V * B, K))
    a = loss_fn(torch.from_numpy(S), torch.from_numpy(T), 0).item()
    loss_fn.center.zero_()
    b = loss_fn(torch.from_numpy(S + rng.normal(size=(V * B, 1))), torch.from_numpy(T), 0).item()
    assert abs(a - b) < 1e-5 * abs(a)
    loss_fn.center.zero_()
    z = T / 0.04
    t = np.exp(z - z.max(1, keepdims=True))
    t /= t.sum(1, keepdims=True)
    entropy = -(t * np.log(np.maximum(t, 1e-300))).sum(1)
    # student crop v must match teacher crop 1 - v: logits tau_s * log t (any shift) reproduce t exactly
    S_opt = 0.1 * np.log(np.maximum(np.concatenate((t[B:], t[:B])), 1e-300))
    best = loss_fn(torch.from_numpy(S_opt), torch.from_numpy(T), 0).item()
    assert abs(best - entropy.mean()) < 1e-5 * max(1.0, entropy.mean())
    assert a > best


def test_oracle_ema_and_momentum_schedule_closed_form():
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.LayerNorm(7)).double()
    ema = odino.ModelEma(model)
    before = [p.detach().numpy().copy() for p in ema.module.parameters()]
    with torch.no_grad():
        for p in model.parameters():
            p.add_(torch.randn_like(p))
    for it, total in [(0, 10), (3, 10), (10, 10)]:
        mu = odino.cosine_momentum(it, total, base=0.996)
        assert mu == pytest.approx(1.0 - (1.0 - 0.996) * (math.cos(math.pi * it / total) + 1) / 2, abs=1e-15)
    assert odino.cosine_momentum(0, 10) == pytest.approx(0.996) and odino.cosine_momentum(10, 10) == pytest.approx(1.0)
    ema.update(model, momentum=0.9)
    for b, e, s in zip(before, ema.module.parameters(), model.parameters()):
        assert np.abs(e.detach().numpy() - (0.9 * b + 0.1 * s.detach().numpy())).max() < 1e-15
    assert all(not p.requires_grad for p in ema.module.parameters())
