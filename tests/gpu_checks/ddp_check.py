"""Developer check (N GPUs, torchrun): data-parallel DINO step.
(a) replicas stay identical after eager and graphed steps; (b) the all-reduced gradients / centre equal a
single-process step on the concatenated global batch; (c) eager (overlapped NCCL) == two-graph step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist

import b200ssl

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
out_dim, ncrops, B = 1024, 4, 4


def build():
    torch.manual_seed(0)
    m = b200ssl.MultiCropWrapper(b200ssl.vit_tiny(), b200ssl.DINOHead(192, out_dim, hidden_dim=256, bottleneck_dim=64)).cuda()
    with torch.no_grad():
        for p in m.parameters():
            if p.ndim == 1:
                p.add_(torch.randn_like(p) * 0.02)
    return m


def crops_for(r):
    g = torch.Generator(device="cuda").manual_seed(100 + r)
    return [torch.randn(B, 3, 224, 224, device="cuda", generator=g).bfloat16() for _ in range(2)] + \
           [torch.randn(B, 3, 96, 96, device="cuda", generator=g).bfloat16() for _ in range(ncrops - 2)]


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


def cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item()


mine = crops_for(rank)
# ---- distributed eager step
mod = build()
ddp = b200ssl.GradBucketDataParallel(mod, bucket_mb=1.0)
teacher = b200ssl.ModelEma(mod)
loss_fn = b200ssl.DINOLoss(out_dim, ncrops, 0.04, 0.04, 0, 10).cuda()
opt = b200ssl.FusedAdamW(b200ssl.param_groups_wd(mod, 0.04), lr=5e-4)
loss0 = b200ssl.dino_step(ddp, teacher, loss_fn, opt, mine, momentum=0.9)[0]
grads = {n: p.grad.clone() for n, p in mod.named_parameters() if p.grad is not None}
center = loss_fn.center.clone()
# ---- single-process reference on the concatenated global batch (every rank computes it)
ref = build()
ref_teacher = b200ssl.ModelEma(ref)
ref_loss = b200ssl.DINOLoss(out_dim, ncrops, 0.04, 0.04, 0, 10).cuda()
ref_opt = b200ssl.FusedAdamW(b200ssl.param_groups_wd(ref, 0.04), lr=5e-4)
allc = [crops_for(r) for r in range(world)]
glob = [torch.cat([allc[r][v] for r in range(world)]) for v in range(ncrops)]
was = dist.is_initialized  # the reference must not all-reduce its centre: run it with a world-1 view
ref_loss.defer_comm = True
dist_world = dist.get_world_size
dist.get_world_size = lambda *a, **k: 1
try:
    l_ref = b200ssl.dino_step(ref, ref_teacher, ref_loss, ref_opt, glob, momentum=0.9)[0]
    ref_grads = {n: p.grad.clone() for n, p in ref.named_parameters() if p.grad is not None}
    ref_center = ref_loss.center.clone()
    ref_more = [b200ssl.dino_step(ref, ref_teacher, ref_loss, ref_opt, glob, momentum=0.9)[0].item() for _ in range(2)]
finally:
    dist.get_world_size = dist_world
ok = True
lsum = loss0.clone()
dist.all_reduce(lsum)
ok &= abs(lsum.item() / world - l_ref.item()) / abs(l_ref.item()) < 2e-3
worst = min(cos(grads[n], g) for n, g in ref_grads.items() if n in grads)
ok &= worst > 0.999
ok &= rel(center, ref_center) < 1e-3
# ---- replicas identical
for n, p in mod.named_parameters():
    t = p.detach().clone()
    dist.broadcast(t, src=0)
    ok &= bool(torch.equal(t, p.detach()))
# ---- two-graph step continues identically to eager on a twin model
twin = build()
twin_ddp = b200ssl.GradBucketDataParallel(twin, bucket_mb=1.0)
twin_teacher = b200ssl.ModelEma(twin)
twin_loss = b200ssl.DINOLoss(out_dim, ncrops, 0.04, 0.04, 0, 10).cuda()
twin_opt = b200ssl.FusedAdamW(b200ssl.param_groups_wd(twin, 0.04), lr=5e-4)
step = b200ssl.GraphedDinoStep(twin_ddp, twin_teacher, twin_loss, twin_opt, mine)  # warm-up leaves no trace: replay i == step i
eager = [b200ssl.dino_step(ddp, teacher, loss_fn, opt, mine, momentum=0.9)[0].item() for _ in range(2)]
graph_all = [step(mine, momentum=0.9).item() for _ in range(3)]
ok &= abs(graph_all[0] - loss0.item()) / abs(loss0.item()) < 1e-4   # first replay == eager step 0 on the same weights
graph = graph_all[1:]
ok &= max(abs(a - b) / abs(a) for a, b in zip(eager, graph)) < 1e-2
ok &= rel(twin_loss.center, loss_fn.center) < 1e-2
for n, p in twin.named_parameters():
    t = p.detach().clone()
    dist.broadcast(t, src=0)
    ok &= bool(torch.equal(t, p.detach()))
def gmean(vals):
    t = torch.tensor(vals, device="cuda")
    dist.all_reduce(t)
    return [round(v, 4) for v in (t / world).tolist()]


eager_g, graph_g = gmean(eager), gmean(graph)   # global (mean over ranks) losses of steps 1, 2
ok &= max(abs(a - b) / abs(b) for a, b in zip(eager_g, ref_more)) < 1e-2
ok &= max(abs(a - b) / abs(b) for a, b in zip(graph_g, ref_more)) < 1e-2
print(f"rank {rank}: loss {loss0.item():.5f} (global ref {l_ref.item():.5f}) worst grad cos {worst:.6f} "
      f"center rel {rel(center, ref_center):.2e} | steps 1-2 global loss: ref {ref_more} eager {eager_g} graph {graph_g} "
      f"-> {'OK' if ok else 'FAIL'}")
# the captured step holds NCCL kernels: release it before the communicator goes away (destroy_process_group hangs
# otherwise), then leave without the interpreter's teardown
step.release()
del step
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
sys.stderr.flush()
os._exit(0 if ok else 1)
