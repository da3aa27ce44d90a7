"""Developer check: ONE full DINO step at the benchmarked shape (ViT-S/16, 2 x 224^2 + 10 x 96^2, out_dim 65536) against
the fp32 oracle -- loss, logits, centre, EMA and EVERY parameter gradient (cosine); eager and CUDA-graph replay.
Prints the per-tensor numbers the -m gpu test (tests/test_gpu_model.py::test_full_step_at_bench_shape) asserts on.
usage: python tests/gpu_checks/full_step_check.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
import b200ssl
from oracle import dino as odino
from oracle import vision_transformer as ovt

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
out_dim, ncrops = 65536, 12


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-300)).item()


torch.manual_seed(0)
ref = odino.MultiCropWrapper(ovt.vit_small(), ovt.DINOHead(384, out_dim)).cuda()
with torch.no_grad():
    for p in ref.parameters():
        if p.ndim == 1:
            p.add_(torch.randn_like(p) * 0.02)
g = torch.Generator(device="cuda").manual_seed(1234)
crops = [torch.randn(B, 3, 224, 224, device="cuda", generator=g) for _ in range(2)] + \
        [torch.randn(B, 3, 96, 96, device="cuda", generator=g) for _ in range(ncrops - 2)]
ref_t = odino.ModelEma(ref)
ref_l = odino.DINOLoss(out_dim, ncrops, 0.04, 0.04, 0, 10).cuda()
ref_o = torch.optim.AdamW(b200ssl.param_groups_wd(ref, 0.04), lr=5e-4)
state0 = {k: v.clone() for k, v in ref.state_dict().items()}
l_ref, s_ref, t_ref = odino.dino_step(ref, ref_t, ref_l, ref_o, crops, momentum=0.9)
g_ref = {n: p.grad.clone() for n, p in ref.named_parameters() if p.grad is not None}
# the same step under torch's bf16 autocast: the noise floor of "a bf16 implementation of this step"
ac = odino.MultiCropWrapper(ovt.vit_small(), ovt.DINOHead(384, out_dim)).cuda()
ac.load_state_dict(state0)
ac_t = odino.ModelEma(ac)
ac_l = odino.DINOLoss(out_dim, ncrops, 0.04, 0.04, 0, 10).cuda()
ac_o = torch.optim.AdamW(b200ssl.param_groups_wd(ac, 0.04), lr=5e-4)
l_ac, s_ac, _ = odino.dino_step(ac, ac_t, ac_l, ac_o, crops, momentum=0.9, autocast_dtype=torch.bfloat16)
g_ac = {n: p.grad.clone() for n, p in ac.named_parameters() if p.grad is not None}
print(f"autocast oracle: loss rel {abs(l_ac.item() - l_ref.item()) / abs(l_ref.item()):.2e} logits rel {rel(s_ac, s_ref):.2e}")

for mode in ("eager", "graph"):
    mine = b200ssl.MultiCropWrapper(b200ssl.vit_small(), b200ssl.DINOHead(384, out_dim)).cuda()
    mine.load_state_dict(state0)
    mine_t = b200ssl.ModelEma(mine)
    mine_l = b200ssl.DINOLoss(out_dim, ncrops, 0.04, 0.04, 0, 10).cuda()
    mine_o = b200ssl.FusedAdamW(b200ssl.param_groups_wd(mine, 0.04), lr=5e-4)
    c16 = [c.bfloat16() for c in crops]
    if mode == "eager":
        l, s, t = b200ssl.dino_step(mine, mine_t, mine_l, mine_o, c16, momentum=0.9)
    else:
        ddp = b200ssl.GradBucketDataParallel(mine)
        step = b200ssl.GraphedDinoStep(ddp, mine_t, mine_l, mine_o, c16)
        l = step(c16, momentum=0.9)
        s, t = step.student_out, step.teacher_out
    torch.cuda.synchronize()
    mine_l.finish_center_update()
    print(f"[{mode}] loss {l.item():.6f} oracle {l_ref.item():.6f} rel {abs(l.item() - l_ref.item()) / abs(l_ref.item()):.2e}; "
          f"student logits rel {rel(s, s_ref):.2e} teacher logits rel {rel(t, t_ref):.2e} centre rel "
          f"{rel(mine_l.center, ref_l.center):.2e}")
    rows = []
    for n, p in mine.named_parameters():
        if p.grad is None:
            continue
        rows.append((cos(p.grad, g_ref[n]), n, cos(g_ac[n], g_ref[n]), g_ref[n].norm().item()))
    rows.sort()
    print(f"[{mode}] {len(rows)} gradient tensors; worst 12 (cosine ours, name, cosine autocast-oracle, |g|):")
    for c, n, ca, gn in rows[:12]:
        print(f"   {c:.6f}  {n:45s} {ca:.6f}  {gn:.3e}")
    print(f"[{mode}] tensors below 0.999: {sum(1 for r in rows if r[0] < 0.999)}; autocast oracle below 0.999: "
          f"{sum(1 for r in rows if r[2] < 0.999)}")
    ema = max(rel(q, p) for (n, p), (_, q) in zip(ref_t.module.named_parameters(), mine_t.module.named_parameters()))
    upd = max(rel(q, p) for (n, p), (_, q) in zip(ref.named_parameters(), mine.named_parameters()))
    print(f"[{mode}] EMA max rel {ema:.2e}; updated student weights max rel {upd:.2e}")
