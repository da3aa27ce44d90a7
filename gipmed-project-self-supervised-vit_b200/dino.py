"""DINO-side pieces of the hot path behind the reference's call slots (train.py:1045 ``model(input)``,
:1053 ``loss_fn(output, target)``, :1063-1078 backward/clip/step, :1081 ``model_ema.update(model)``,
:634 DDP): fused loss with running centre, multi-crop wrapper, teacher EMA, fused AdamW, the
bucketed gradient all-reduce and the step that strings them together.
"""
from __future__ import annotations

import copy
import math

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops


# ------------------------------------------------------------------------------------------------
# loss
# ------------------------------------------------------------------------------------------------
class DINOLoss(nn.Module):
    """``forward(student_output, teacher_output, epoch)`` — rows crop-major ([ncrops*B, K] / [2*B, K]).

    One fused forward kernel (single pass over the logits, online softmax statistics) and one fused
    backward kernel; the centre update is a column-sum kernel, one all-reduce of [K] fp32 when a
    process group is initialised, and an in-place EMA kernel. Two-positional-argument compatible with
    the reference's ``loss_fn(output, target)`` call."""

    def __init__(self, out_dim, ncrops, warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs, nepochs,
                 student_temp=0.1, center_momentum=0.9):
        super().__init__()
        self.student_temp = student_temp
        self.center_momentum = center_momentum
        self.ncrops = ncrops
        self.register_buffer("center", torch.zeros(1, out_dim))
        self.teacher_temp_schedule = np.concatenate((
            np.linspace(warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs),
            np.ones(nepochs - warmup_teacher_temp_epochs) * teacher_temp))
        self._pending = None  # (work handle, local column sum, rows) of an in-flight centre all-reduce
        self.defer_comm = False  # True: forward only computes the local column sums; the caller all-reduces

    def forward(self, student_output, teacher_output, epoch=0):
        ops.require_cuda(student_output, "DINOLoss")
        self.finish_center_update()  # the centre this step reads must include the previous step's update
        temp = float(self.teacher_temp_schedule[epoch])
        s = student_output if student_output.dtype == torch.bfloat16 else student_output.to(torch.bfloat16)
        t = teacher_output.detach()
        t = t if t.dtype == torch.bfloat16 else t.to(torch.bfloat16)
        center = self.center.view(-1)
        if center.dtype != torch.float32:
            raise RuntimeError("DINOLoss.center must stay fp32")
        loss = ops.DinoLossFn.apply(s.contiguous(), t.contiguous(), center, self.ncrops, self.student_temp, temp)
        self.update_center(t)
        return loss

    @torch.no_grad()
    def update_center(self, teacher_output):
        """center <- m*center + (1-m)*mean_rows(teacher_output) over all ranks. The all-reduce is
        launched asynchronously and only waited for when the centre is next read."""
        t = teacher_output if teacher_output.dtype == torch.bfloat16 else teacher_output.to(torch.bfloat16)
        batch_sum = ops.teacher_colsum(t.contiguous())
        rows, world, work = t.shape[0], 1, None
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size()
            if world > 1 and not self.defer_comm:
                work = dist.all_reduce(batch_sum, async_op=True)
        self._pending = (work, batch_sum, rows * world)
        if work is None and not (self.defer_comm and world > 1):
            self.finish_center_update()

    @torch.no_grad()
    def allreduce_center_now(self):
        """Deferred mode (two-graph step): sum the pending local column sums over ranks, eagerly."""
        if self._pending is not None and self._pending[0] is None and dist.is_initialized() \
                and dist.get_world_size() > 1:
            dist.all_reduce(self._pending[1])

    @torch.no_grad()
    def finish_center_update(self):
        if self._pending is None:
            return
        work, batch_sum, total_rows = self._pending
        self._pending = None
        if work is not None:
            work.wait()
        ops.center_update(self.center.view(-1), batch_sum, total_rows, self.center_momentum)


# ------------------------------------------------------------------------------------------------
# multi-crop wrapper
# ------------------------------------------------------------------------------------------------
class MultiCropWrapper(nn.Module):
    """Runs the backbone once per run of equal-resolution crops and the head once on all rows."""

    def __init__(self, backbone, head):
        super().__init__()
        backbone.fc, backbone.head = nn.Identity(), nn.Identity()
        self.backbone = backbone
        self.head = head

    def forward(self, x):
        if not isinstance(x, (list, tuple)):
            x = [x]
        sizes = [int(inp.shape[-1]) for inp in x]
        bounds = [i + 1 for i in range(len(sizes)) if i + 1 == len(sizes) or sizes[i + 1] != sizes[i]]
        start, groups = 0, []
        for end in bounds:
            groups.append(x[start] if end - start == 1 else _cat_or_view(list(x[start:end])))
            start = end
        if len(groups) > 1 and MERGE_CROP_GROUPS["on"] and hasattr(self.backbone, "forward_multi"):
            # all resolutions in one pass over the packed token rows (VisionTransformer.forward_multi)
            return self.head(self.backbone.forward_multi(groups))
        outs = []
        for group in groups:
            out = self.backbone(group)
            if isinstance(out, tuple):
                out = out[0]
            outs.append(out)
        feats = outs[0] if len(outs) == 1 else torch.cat(outs)
        return self.head(feats)


# Run the crop groups of different resolution through the backbone in ONE pass (packed token rows) instead of one
# pass per resolution. Same results; off only for A/B measurements (bench.py --no-merge-crops).
MERGE_CROP_GROUPS = {"on": True}


def _cat_or_view(tensors):
    """torch.cat along dim 0 -- or, when the crops already lie back to back in one allocation (GraphedDinoStep's
    static inputs, ``alloc_crop_buffers``), a zero-copy view of that memory."""
    t0 = tensors[0]
    adjacent = all(t.is_contiguous() and t.dtype == t0.dtype and t.shape == t0.shape and not t.requires_grad
                   and t.untyped_storage().data_ptr() == t0.untyped_storage().data_ptr() for t in tensors)
    if adjacent:
        for a, b in zip(tensors[:-1], tensors[1:]):
            if b.storage_offset() != a.storage_offset() + a.numel():
                adjacent = False
                break
    if not adjacent:
        return torch.cat(tensors)
    return torch.as_strided(t0, (len(tensors) * t0.shape[0],) + tuple(t0.shape[1:]), t0.stride(), t0.storage_offset())


def alloc_crop_buffers(example_crops):
    """Device buffers shaped like ``example_crops`` in which consecutive crops of equal resolution share one
    allocation, so the multi-crop wrapper can feed each resolution group to the backbone without a concatenation
    copy. Returns the list of per-crop views."""
    out, i = [], 0
    while i < len(example_crops):
        j = i
        while j + 1 < len(example_crops) and example_crops[j + 1].shape == example_crops[i].shape \
                and example_crops[j + 1].dtype == example_crops[i].dtype:
            j += 1
        c = example_crops[i]
        block = torch.empty((j - i + 1,) + tuple(c.shape), dtype=c.dtype, device=c.device)
        out += [block[k] for k in range(j - i + 1)]
        i = j + 1
    return out


def cosine_momentum(it: int, total_iters: int, base: float = 0.996, final: float = 1.0) -> float:
    return final - (final - base) * (math.cos(math.pi * it / max(total_iters, 1)) + 1) / 2


def cosine_scheduler(base_value, final_value, epochs, niter_per_ep, warmup_epochs=0, start_warmup_value=0.0):
    """Per-iteration cosine schedule with linear warm-up (learning rate, weight decay and teacher momentum in DINO
    training; the reference builds its LR schedule with timm's ``create_scheduler_v2``, train.py:883-898). The fused
    optimiser reads lr / weight_decay from its param groups before every step -- also under CUDA-graph replay."""
    warmup_iters = int(warmup_epochs * niter_per_ep)
    warm = np.linspace(start_warmup_value, base_value, warmup_iters) if warmup_iters > 0 else np.array([])
    iters = np.arange(epochs * niter_per_ep - warmup_iters)
    sched = final_value + 0.5 * (base_value - final_value) * (1 + np.cos(np.pi * iters / max(len(iters), 1)))
    out = np.concatenate((warm, sched))
    assert len(out) == epochs * niter_per_ep
    return out


def apply_schedules(optimizer, it, lr_schedule=None, wd_schedule=None):
    """Write iteration ``it`` of the schedules into the param groups (weight decay only where it is non-zero)."""
    for group in optimizer.param_groups:
        if lr_schedule is not None:
            group["lr"] = float(lr_schedule[it])
        if wd_schedule is not None and group.get("weight_decay", 0.0) > 0:
            group["weight_decay"] = float(wd_schedule[it])


def cancel_gradients_last_layer(epoch, model, freeze_last_layer, optimizer):
    """DINO's first-epoch stabiliser: the prototype layer is not updated while ``epoch < freeze_last_layer``.
    Upstream sets ``p.grad = None`` so the optimiser skips those tensors (no step, no decay, moments untouched);
    gradients live in persistent flat buckets here, so the same effect is obtained by excluding the tensors from
    the fused optimiser's chunk tables (``FusedAdamW.set_frozen``). Call once per epoch, before the step."""
    frozen = [p for n, p in model.named_parameters() if "last_layer" in n] if epoch < freeze_last_layer else []
    optimizer.set_frozen(frozen)


# ------------------------------------------------------------------------------------------------
# chunk tables for the multi-tensor kernels
# ------------------------------------------------------------------------------------------------
_CHUNK = 65536


class _ScalarStager:
    """Per-step scalars (lr, weight decay, Adam bias corrections, EMA momentum) travel host -> device through a RING
    of pinned slots. The copy is asynchronous and a whole step is enqueued in well under a millisecond (one graph
    launch) while the GPU needs tens of milliseconds to run it, so a single reused pinned buffer would be
    overwritten for step k+1 before step k's copy has executed. Each slot carries an event recorded behind its
    last copy; the host waits on it before reusing the slot, i.e. it can run at most ``slots`` steps ahead."""

    def __init__(self, shape, device, slots=8):
        self.host = torch.zeros((slots,) + tuple(shape), dtype=torch.float32).pin_memory()
        self.dev = torch.zeros(tuple(shape), dtype=torch.float32, device=device)
        self._events = [None] * slots
        self._i = 0

    def slot(self):
        """The next pinned slot, safe to overwrite (its previous host-to-device copy has executed)."""
        k = self._i % len(self._events)
        if self._events[k] is not None:
            self._events[k].synchronize()
        return self.host[k]

    def push(self):
        """Enqueue the copy of the slot handed out by ``slot()`` on the current stream."""
        k = self._i % len(self._events)
        self.dev.copy_(self.host[k], non_blocking=True)
        if self._events[k] is None:
            self._events[k] = torch.cuda.Event()
        self._events[k].record()
        self._i += 1


def _chunk_rows(tensors_per_row, numels):
    """int64 table: one row per <=_CHUNK elements; columns = per-tensor byte addresses + count (+extras)."""
    rows = []
    for tens, n in zip(tensors_per_row, numels):
        for off in range(0, n, _CHUNK):
            cnt = min(_CHUNK, n - off)
            rows.append([(t.data_ptr() + off * t.element_size()) if isinstance(t, torch.Tensor) else t for t in tens]
                        + [cnt])
    return rows


def _detach_weight_norm(model):
    """Old-style nn.utils.weight_norm leaves a non-leaf ``weight`` attribute that deepcopy refuses;
    the forward pre-hook recomputes it anyway, so park a detached copy before cloning the module."""
    from torch.nn.utils.weight_norm import WeightNorm
    for m in model.modules():
        for hook in m._forward_pre_hooks.values():
            if isinstance(hook, WeightNorm) and isinstance(getattr(m, hook.name, None), torch.Tensor):
                setattr(m, hook.name, getattr(m, hook.name).detach())


class ModelEma(nn.Module):
    """Teacher as an EMA of the student with timm's ``ModelEmaV2`` call convention (train.py:619-620,
    :948, :1081): ``.module`` is the averaged copy, ``update(model)`` folds every floating-point
    state-dict tensor with ONE multi-tensor kernel launch. ``momentum`` overrides ``decay`` per step."""

    def __init__(self, model, decay=0.9998, device=None):
        super().__init__()
        _detach_weight_norm(model)
        self.module = copy.deepcopy(model)
        self.module.eval()
        for p in self.module.parameters():
            p.requires_grad_(False)
        self.decay = decay
        self.device = device
        if device is not None:
            self.module.to(device=device)
        self._table = None
        self._table_key = None
        self._m_stage = None     # _ScalarStager for the per-step momentum
        self._m_value = decay
        self._pairs, self._others, self._shadowed = [], [], []

    def _build(self, model):
        pairs, others = [], []
        for e, s in zip(self.module.state_dict().values(), model.state_dict().values()):
            if e.dtype == torch.float32 and s.dtype == torch.float32 and e.is_contiguous() and s.is_contiguous():
                pairs.append((e, s))
            else:
                others.append((e, s))
        key = tuple((e.data_ptr(), s.data_ptr(), e.numel()) for e, s in pairs)
        if key != self._table_key:
            # matrices the GEMMs read through a bf16 shadow get it refreshed by the same kernel pass
            by_ptr = {p.data_ptr(): p for p in self.module.parameters() if p.ndim >= 2}
            rows, shadowed = [], []
            for e, s in pairs:
                param = by_ptr.get(e.data_ptr())
                if param is not None and id(param) not in ops.shadows._d:
                    param = None   # never read through a bf16 shadow (position table, weight-norm direction, ...)
                sh = ops.shadows.shadow_for(param).data_ptr() if param is not None else 0
                if param is not None:
                    shadowed.append(param)
                for off in range(0, e.numel(), _CHUNK):
                    rows.append([e.data_ptr() + 4 * off, s.data_ptr() + 4 * off, min(_CHUNK, e.numel() - off),
                                 sh + 2 * off if sh else 0])
            self._table = torch.tensor(rows, dtype=torch.int64).to(pairs[0][0].device) if rows else None
            self._table_key = key
            self._pairs, self._others, self._shadowed = pairs, others, shadowed
        return self._table

    def prepare(self, model, momentum=None):
        """Host-side part of an update: (re)build the chunk table if storages moved and stage this step's
        momentum in device memory (pinned host -> device, async). Must run OUTSIDE a CUDA-graph capture."""
        m = self.decay if momentum is None else momentum
        model = getattr(model, "module", model) if isinstance(model, GradBucketDataParallel) else model
        table = self._build(model)
        if table is not None:
            ops.require_cuda(table, "ModelEma")
            if self._m_stage is None or self._m_stage.dev.device != table.device:
                self._m_stage = _ScalarStager((1,), table.device)
            self._m_stage.slot()[0] = float(m)
            self._m_stage.push()
        self._m_value = float(m)

    @torch.no_grad()
    def launch(self):
        """Device-side part: one multi-tensor kernel over every fp32 state-dict tensor (graph-capturable)."""
        if self._table is not None:
            ops._call("b200ssl_ema_multi_tensor", self._table.data_ptr(), self._table.shape[0],
                      self._m_stage.dev.data_ptr(), ops._stream())
        m = self._m_value
        for e, s in self._others:  # integer buffers etc.: plain copy like timm
            if e.dtype.is_floating_point:
                e.mul_(m).add_(s.to(e.dtype), alpha=1.0 - m)
            else:
                e.copy_(s)
        fresh = {id(p) for p in self._shadowed}
        for p in self.module.parameters():
            if id(p) in fresh:
                ops.shadows.mark_fresh(p)   # the EMA kernel rewrote the bf16 shadow together with the fp32 value
            else:
                ops.shadows.invalidate(p)

    @torch.no_grad()
    def update(self, model, momentum=None):
        self.prepare(model, momentum)
        self.launch()


# ------------------------------------------------------------------------------------------------
# optimiser
# ------------------------------------------------------------------------------------------------
class FusedAdamW(torch.optim.Optimizer):
    """AdamW + global-norm gradient clipping in two launches (sum of squares, fused update) over a
    device-resident chunk table; also refreshes the bf16 weight shadows the GEMMs read. Follows
    ``torch.optim.AdamW`` update rules; ``step(max_grad_norm=...)`` folds in ``clip_grad_norm_``."""

    def __init__(self, params, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.04):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}
        self._hyper = None           # _ScalarStager for {lr, weight_decay, 1-beta1^t, 1-beta2^t} per group
        self.last_grad_norm_sq = None
        self._gall = None            # (key, concatenated gradient chunk table of all groups, workspace)
        self._frozen = frozenset()   # ids of parameters the step skips entirely (set_frozen)
        self.table_version = 0       # bumped whenever the chunk tables change shape (captured graphs must follow)

    def set_frozen(self, params):
        """Parameters to leave completely untouched by ``step`` (no update, no decay, no moment update, not part
        of the clipping norm) until the set changes again."""
        new = frozenset(id(p) for p in params)
        if new != self._frozen:
            self._frozen = new
            self._tables.clear()
            self._gall = None
            self.table_version += 1

    def _table(self, gi, group):
        params = [p for p in group["params"] if p.grad is not None and id(p) not in self._frozen]

        def _moments(p):  # load_state_dict replaces the moment tensors: the table must follow them
            st = self.state.get(p, {})
            return (st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()) if "exp_avg" in st else (0, 0)

        key = tuple((p.data_ptr(), p.grad.data_ptr()) + _moments(p) for p in params)
        ent = self._tables.get(gi)
        if ent is not None and ent["key"] == key:
            return ent
        rows, grows = [], []
        for p in params:
            if p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                raise RuntimeError("FusedAdamW needs contiguous fp32 parameters and gradients")
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p)
                st["exp_avg_sq"] = torch.zeros_like(p)
            sh = ops.shadows.shadow_for(p)
            n = p.numel()
            for off in range(0, n, _CHUNK):
                cnt = min(_CHUNK, n - off)
                rows.append([p.data_ptr() + 4 * off, p.grad.data_ptr() + 4 * off, st["exp_avg"].data_ptr() + 4 * off,
                             st["exp_avg_sq"].data_ptr() + 4 * off, cnt, 1 if group["weight_decay"] > 0 else 0, 0,
                             sh.data_ptr() + 2 * off])
                grows.append([p.grad.data_ptr() + 4 * off, cnt])
        dev = params[0].device
        key = tuple((p.data_ptr(), p.grad.data_ptr()) + _moments(p) for p in params)   # moments exist now
        ent = {"key": key, "params": params, "table": torch.tensor(rows, dtype=torch.int64).to(dev),
               "gtable": torch.tensor(grows, dtype=torch.int64).to(dev)}
        self._tables[gi] = ent
        return ent

    def _entries(self):
        ents = []
        for gi, group in enumerate(self.param_groups):
            if any(p.grad is not None and id(p) not in self._frozen for p in group["params"]):
                ents.append((gi, group, self._table(gi, group)))
        if ents:
            self._grad_table_all(ents)   # built here (host side, never inside a graph capture)
        return ents

    def _grad_table_all(self, ents):
        gkey = tuple(id(ent["gtable"]) for _, _, ent in ents)
        if self._gall is None or self._gall[0] != gkey:
            gt = torch.cat([ent["gtable"] for _, _, ent in ents]) if len(ents) > 1 else ents[0][2]["gtable"]
            self._gall = (gkey, gt, torch.zeros(gt.shape[0] + 1, dtype=torch.float32, device=gt.device))
        return self._gall

    def prepare_step(self):
        """Host-side part of a step: advance the step counters and stage {lr, weight_decay, 1-beta1^t,
        1-beta2^t, 0} per group in device memory (pinned -> device, async). Runs OUTSIDE graph capture, so
        schedulers may change lr / weight_decay between replays."""
        ents = self._entries()
        if not ents:
            return ents
        dev = ents[0][2]["table"].device
        if self._hyper is None or self._hyper.dev.device != dev:
            n = len(self.param_groups)
            self._hyper = _ScalarStager((n, 8), dev)
        host = self._hyper.slot()
        for gi, group, _ in ents:
            group["step"] = group.get("step", 0) + 1
            t = group["step"]
            b1, b2 = group["betas"]
            host[gi, 0] = float(group["lr"])
            host[gi, 1] = float(group["weight_decay"])
            host[gi, 2] = 1.0 - b1 ** t
            host[gi, 3] = 1.0 - b2 ** t
            host[gi, 4] = 0.0
        self._hyper.push()
        return ents

    @torch.no_grad()
    def launch_step(self, max_grad_norm: float = 0.0, ents=None):
        """Device-side part: grad sum-of-squares (when clipping) + one fused AdamW launch per group."""
        ents = self._entries() if ents is None else ents
        if not ents:
            return
        gnorm_ptr = None
        if max_grad_norm and max_grad_norm > 0:
            # ONE sum-of-squares pass over the gradients of all groups (a concatenated chunk table): the global norm
            _, gt, ws = self._grad_table_all(ents)
            ops._call("b200ssl_sumsq_multi_tensor", gt.data_ptr(), gt.shape[0], ws.data_ptr(), ops._stream(), launches=2)
            total = ws[0:1]
            self.last_grad_norm_sq = total
            gnorm_ptr = total.data_ptr()
        for gi, group, ent in ents:
            b1, b2 = group["betas"]
            ops._call("b200ssl_adamw_multi_tensor", ent["table"].data_ptr(), ent["table"].shape[0], gnorm_ptr,
                      self._hyper.dev[gi].data_ptr(), float(b1), float(b2), float(group["eps"]),
                      float(max_grad_norm or 0.0), ops._stream())
            for p in ent["params"]:
                ops.shadows.mark_fresh(p)

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm: float = 0.0):
        self.launch_step(max_grad_norm, self.prepare_step())
        return None

    def zero_grad(self, set_to_none: bool = False):
        """Defaults to ``set_to_none=False`` (torch's default is True): gradients are persistent buffers here --
        views into GradBucketDataParallel's flat buckets that the kernels accumulate into and the cached chunk
        tables point at. Freeing them would silently detach the parameters from their buckets."""
        return super().zero_grad(set_to_none=set_to_none)


def param_groups_wd(model, weight_decay):
    """DINO's grouping: no weight decay on biases and 1-D (norm) parameters."""
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (no_decay if name.endswith(".bias") or p.ndim == 1 else decay).append(p)
    return [{"params": decay, "weight_decay": weight_decay}, {"params": no_decay, "weight_decay": 0.0}]


# ------------------------------------------------------------------------------------------------
# data parallel
# ------------------------------------------------------------------------------------------------
class GradBucketDataParallel(nn.Module):
    """Data-parallel wrapper (reference: ``NativeDDP(model, device_ids=[device])`` train.py:634).

    Gradients live as views into flat fp32 buckets (reverse parameter order, ``bucket_mb`` each). A
    post-accumulate hook per parameter counts arrivals; when every parameter of a bucket has received
    all its contributions for this backward (the backbone is used once per resolution group), the
    bucket is all-reduced (AVG) asynchronously on NCCL's stream, overlapping the rest of backward.
    ``finish()`` (called by ``dino_step`` before the optimiser) waits for the outstanding handles.
    The first backward runs without overlap to learn the per-parameter contribution counts."""

    def __init__(self, module, bucket_mb: float = 25.0, process_group=None, compress: str | None = None):
        super().__init__()
        self.module = module
        self.pg = process_group
        # compress="bf16": every bucket travels as bf16 (cast -> all-reduce -> cast back into the fp32 bucket): half
        # the bytes on the wire (88 MB instead of 176 MB for ViT-S + head); the optimiser still sees fp32 gradients
        if compress not in (None, "none", "bf16"):
            raise ValueError(f"unknown gradient compression {compress!r}")
        self.compress = "bf16" if compress == "bf16" else None
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._avg_native = dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        params = [p for p in module.parameters() if p.requires_grad]
        self._params = params
        cap = int(bucket_mb * 1024 * 1024 / 4)
        self.buckets, cur, cur_n = [], [], 0
        for p in reversed(params):
            if cur and cur_n + p.numel() > cap:
                self.buckets.append(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            self.buckets.append(cur)
        self._flat, self._bucket_of, self._view_of = [], {}, {}
        sizes = [sum((p.numel() + 3) // 4 * 4 for p in bucket) for bucket in self.buckets]
        # all buckets are slices of ONE allocation: zeroing the gradients is a single memset node per step
        self._all = torch.zeros(sum(sizes), dtype=torch.float32, device=params[0].device)
        self._all16 = torch.zeros(sum(sizes), dtype=torch.bfloat16, device=params[0].device) \
            if self.compress and self.world > 1 else None
        self._flat16 = []
        start = 0
        for bi, (bucket, n) in enumerate(zip(self.buckets, sizes)):
            flat = self._all[start:start + n]
            if self._all16 is not None:
                self._flat16.append(self._all16[start:start + n])
            start += n
            off = 0
            for p in bucket:
                p.grad = flat[off:off + p.numel()].view_as(p)
                self._view_of[id(p)] = p.grad
                off += (p.numel() + 3) // 4 * 4
                self._bucket_of[id(p)] = bi
            self._flat.append(flat)
        self.defer_comm = False        # True: hooks / finish() launch nothing; the caller runs allreduce_now()
        self._expected = None          # id(p) -> hook firings per backward (learned on the first step)
        self._seen = {}
        self._pending = [0] * len(self.buckets)
        self._handles = []
        self._launched = [False] * len(self.buckets)
        if self.world > 1:
            for p in params:  # replicas must start identical (DDP broadcasts rank 0's state)
                dist.broadcast(p.data, src=0, group=self.pg)
            for b in module.buffers():
                dist.broadcast(b.data, src=0, group=self.pg)
        for p in params:
            p.register_post_accumulate_grad_hook(self._hook)
            # our wgrad / LayerNorm-backward kernels add straight into these bucket views (no autograd
            # accumulation); they call the same arrival hook once per contribution
            ops.register_grad_sink(p, self._hook)
        self._arm()

    def _arm(self):
        self._seen = {}
        self._launched = [False] * len(self.buckets)
        if self._expected is not None:
            self._pending = [sum(self._expected.get(id(p), 1) for p in b) for b in self.buckets]

    def _adopt(self, p):
        """``optimizer.zero_grad()`` with torch's default ``set_to_none=True`` (the reference loop, train.py:1061)
        makes autograd allocate a fresh ``p.grad`` outside the flat bucket; move it into the bucket view and point
        ``p.grad`` back at the view, so what is all-reduced is what backward produced."""
        view = self._view_of[id(p)]
        g = p.grad
        if g is None:
            view.zero_()
        elif g.data_ptr() != view.data_ptr():
            view.copy_(g)
        p.grad = view

    def _check_views(self):
        for p in self._params:
            g = p.grad
            if g is None or g.data_ptr() != self._view_of[id(p)].data_ptr():
                if p.is_cuda and torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("GradBucketDataParallel: a parameter's .grad left its flat bucket during a "
                                       "CUDA-graph capture; use ddp.zero_grad() / zero_grad(set_to_none=False)")
                self._adopt(p)

    def _hook(self, p):
        g = p.grad
        if g is not None and g.data_ptr() != self._view_of[id(p)].data_ptr():
            self._adopt(p)
        self._seen[id(p)] = self._seen.get(id(p), 0) + 1
        if self._expected is None or self.world == 1 or self.defer_comm:
            return
        bi = self._bucket_of[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0 and not self._launched[bi]:
            self._launch(bi)

    def _wire(self, bi):
        """The tensor that goes on the wire for bucket bi (the bf16 copy under compression, filled here)."""
        if self._all16 is None:
            return self._flat[bi]
        if self._flat[bi].is_cuda:
            ops.cast_f32_to_bf16(self._flat[bi], self._flat16[bi])
        else:
            self._flat16[bi].copy_(self._flat[bi])
        return self._flat16[bi]

    def _unwire(self, bi):
        """After the all-reduce: bring a compressed bucket back into its fp32 home."""
        if self._all16 is None:
            return
        if self._flat[bi].is_cuda:
            ops.cast_bf16_to_f32(self._flat16[bi], self._flat[bi])
        else:
            self._flat[bi].copy_(self._flat16[bi])

    def _launch(self, bi):
        self._launched[bi] = True
        if self._avg_native:
            op = dist.ReduceOp.AVG
        else:  # gloo (CPU tests) has no AVG: pre-scale, then SUM
            op = dist.ReduceOp.SUM
            self._flat[bi].div_(self.world)
        self._handles.append((bi, dist.all_reduce(self._wire(bi), op=op, group=self.pg, async_op=True)))

    def zero_grad(self, set_to_none: bool = False):
        """Zero the flat buckets (gradients stay views into them): one memset over the single allocation."""
        if self._all.is_cuda:
            ops._call("b200ssl_zero_bytes", self._all.data_ptr(), self._all.numel() * 4, ops._stream(), launches=0)
        else:
            self._all.zero_()
        for p in self._params:
            view = self._view_of[id(p)]
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                p.grad = view

    def allreduce_now(self):
        """Deferred mode (two-graph step): average every flat bucket over ranks, eagerly, in bucket order."""
        self._check_views()
        if self.world > 1:
            for bi in range(len(self.buckets)):
                if self._avg_native:
                    dist.all_reduce(self._wire(bi), op=dist.ReduceOp.AVG, group=self.pg)
                else:
                    self._flat[bi].div_(self.world)
                    dist.all_reduce(self._wire(bi), op=dist.ReduceOp.SUM, group=self.pg)
                self._unwire(bi)

    def finish(self):
        """Wait for the bucket all-reduces of this backward; re-arm the counters for the next one."""
        if self.defer_comm:
            return
        self._check_views()
        if self.world > 1:
            for bi in range(len(self.buckets)):
                if not self._launched[bi]:
                    self._launch(bi)
            for bi, h in self._handles:
                h.wait()
                self._unwire(bi)
        self._handles = []
        if self._expected is None:
            self._expected = dict(self._seen)
        self._arm()

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


# ------------------------------------------------------------------------------------------------
# the step
# ------------------------------------------------------------------------------------------------
def _step_prepare(student, teacher_ema, optimizer, momentum):
    """Host-only work of a step (tables, per-step scalars -> device). Never captured in a graph."""
    ddp = student if isinstance(student, GradBucketDataParallel) else None
    ents = optimizer.prepare_step() if isinstance(optimizer, FusedAdamW) else None
    teacher_ema.prepare(ddp.module if ddp is not None else student, momentum)
    return ents


# Measured on B200 (bench.py, same box, alternating): 54.59 / 54.78 ms per step without, 54.19 / 53.88 ms with.
TEACHER_STREAM = {"on": True}
# N > 1: capture the NCCL all-reduces inside the step graph (overlapped with backward). Off = the two-graph step with
# eager, serial all-reduces in between (round-1 behaviour; kept as the automatic fallback and for A/B runs).
NCCL_IN_GRAPH = {"on": True}
_SIDE_STREAMS = {}


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def _step_compute(student, teacher_ema, loss_fn, optimizer, crops, epoch):
    """Forward (teacher + student), loss, backward. Kernel launches (and, unless deferred, NCCL calls) only."""
    if TEACHER_STREAM["on"]:
        # teacher forward on a second stream (a parallel branch of the captured graph): it is independent of the
        # student forward until the loss, so its CTAs can fill the tails of the student's persistent kernels
        main = torch.cuda.current_stream()
        side = _side_stream(main.device)
        side.wait_stream(main)
        with torch.cuda.stream(side), torch.no_grad():
            teacher_out = teacher_ema.module(list(crops[:2]))
        student_out = student(list(crops))
        main.wait_stream(side)
        teacher_out.record_stream(main)
    else:
        with torch.no_grad():
            teacher_out = teacher_ema.module(list(crops[:2]))
        student_out = student(list(crops))
    loss = loss_fn(student_out, teacher_out, epoch)
    ddp = student if isinstance(student, GradBucketDataParallel) else None
    if ddp is not None:
        ddp.zero_grad()
    else:
        optimizer.zero_grad(set_to_none=False)  # keep .grad storage stable: the fused optimiser caches its table
    loss.backward()
    if ddp is not None:
        ddp.finish()
    return loss.detach(), student_out.detach(), teacher_out.detach()


def _step_update(student, teacher_ema, loss_fn, optimizer, clip_grad, ents):
    """Gradient clipping + AdamW, teacher EMA, centre update. Kernel launches only."""
    if isinstance(optimizer, FusedAdamW):
        optimizer.launch_step(clip_grad or 0.0, ents)
    else:
        if clip_grad:
            torch.nn.utils.clip_grad_norm_([p for p in student.parameters() if p.requires_grad], clip_grad)
        optimizer.step()
    teacher_ema.launch()
    # the centre all-reduce was launched right after the loss forward and has overlapped with backward and
    # the optimiser; close it inside the step so a step is self-contained
    if hasattr(loss_fn, "finish_center_update"):
        loss_fn.finish_center_update()


def _step_launch(student, teacher_ema, loss_fn, optimizer, crops, epoch, clip_grad, ents):
    """Device work of a whole step on the current stream."""
    out = _step_compute(student, teacher_ema, loss_fn, optimizer, crops, epoch)
    _step_update(student, teacher_ema, loss_fn, optimizer, clip_grad, ents)
    return out


def dino_step(student, teacher_ema, loss_fn, optimizer, crops, epoch=0, momentum=0.996, clip_grad=3.0):
    """One optimisation step: teacher forward on the 2 global crops, student forward on
    all crops, fused loss (+ centre update), backward (bucketed all-reduce overlapped when ``student`` is
    a ``GradBucketDataParallel``), clip + AdamW, teacher EMA. Returns (loss, student_out, teacher_out)."""
    if isinstance(optimizer, FusedAdamW) and optimizer._hyper is None:
        # very first step: gradients (hence the optimiser tables) do not exist before the first backward
        out = _step_launch_first(student, teacher_ema, loss_fn, optimizer, crops, epoch, momentum, clip_grad)
        return out
    ents = _step_prepare(student, teacher_ema, optimizer, momentum)
    return _step_launch(student, teacher_ema, loss_fn, optimizer, crops, epoch, clip_grad, ents)


def _step_launch_first(student, teacher_ema, loss_fn, optimizer, crops, epoch, momentum, clip_grad):
    ddp = student if isinstance(student, GradBucketDataParallel) else None
    with torch.no_grad():
        teacher_out = teacher_ema.module(list(crops[:2]))
    student_out = student(list(crops))
    loss = loss_fn(student_out, teacher_out, epoch)
    if ddp is not None:
        ddp.zero_grad()
    else:
        optimizer.zero_grad(set_to_none=False)
    loss.backward()
    if ddp is not None:
        ddp.finish()
    optimizer.step(max_grad_norm=clip_grad or 0.0)
    teacher_ema.update(ddp.module if ddp is not None else student, momentum=momentum)
    if hasattr(loss_fn, "finish_center_update"):
        loss_fn.finish_center_update()
    return loss.detach(), student_out.detach(), teacher_out.detach()


class GraphedDinoStep:
    """The whole step captured ONCE as a CUDA graph and replayed: one graph launch per step instead of
    ~1,000 kernel launches from Python (the eager step needs ~55 ms of host time to enqueue, more than the
    GPU needs to execute it). Per-step scalars (lr, weight decay, Adam bias corrections, EMA momentum) are
    staged in device memory before each replay, so schedules keep working; a change of teacher temperature
    (epoch warm-up) triggers a re-capture.

        step = GraphedDinoStep(student, teacher, loss_fn, optimizer, example_crops)
        loss = step(crops, epoch=e, momentum=m)          # crops are copied into the graph's static inputs

    ``student`` may be a GradBucketDataParallel (its NCCL all-reduces are captured with the graph)."""

    def __init__(self, student, teacher_ema, loss_fn, optimizer, example_crops, clip_grad=3.0, warmup=3):
        if not isinstance(optimizer, FusedAdamW):
            raise TypeError("GraphedDinoStep needs b200ssl.FusedAdamW (device-resident step scalars)")
        self.student, self.teacher, self.loss_fn, self.opt = student, teacher_ema, loss_fn, optimizer
        self.clip_grad = clip_grad
        self.static_crops = alloc_crop_buffers(example_crops)   # equal-resolution crops back to back: no cat copy
        for dst, src in zip(self.static_crops, example_crops):
            dst.copy_(src)
        self._warmup = warmup
        self._graph = self._graph_update = None
        self._multi = False
        self.comm_mode = None
        self._pending_static = None
        self._temp = None
        self._opt_version = -1
        self.loss = self.student_out = self.teacher_out = None

    def _set_defer(self, flag):
        if isinstance(self.student, GradBucketDataParallel):
            self.student.defer_comm = flag
        if hasattr(self.loss_fn, "defer_comm"):
            self.loss_fn.defer_comm = flag

    def _comm(self):
        """Eager collectives between the two graphs (N > 1): gradient buckets and the centre column sums."""
        if isinstance(self.student, GradBucketDataParallel):
            self.student.allreduce_now()
        if hasattr(self.loss_fn, "allreduce_center_now"):
            self.loss_fn.allreduce_center_now()

    def _snapshot(self):
        """Everything a step mutates: student parameters / buffers, AdamW moments and step counters, the teacher,
        the centre and the CUDA RNG stream (stochastic depth)."""
        mod = self.student.module if isinstance(self.student, GradBucketDataParallel) else self.student
        snap = {"student": [(t, t.detach().clone()) for t in mod.state_dict().values()],
                "teacher": [(t, t.detach().clone()) for t in self.teacher.module.state_dict().values()],
                "center": self.loss_fn.center.detach().clone() if hasattr(self.loss_fn, "center") else None,
                "steps": [g.get("step", 0) for g in self.opt.param_groups],
                "moments": {}, "rng": torch.cuda.get_rng_state()}
        for g in self.opt.param_groups:
            for p in g["params"]:
                st = self.opt.state.get(p, {})
                snap["moments"][p] = (st["exp_avg"].clone(), st["exp_avg_sq"].clone()) if "exp_avg" in st else None
        return snap

    @torch.no_grad()
    def _restore(self, snap):
        """Put the state back IN PLACE (storages, hence chunk tables and bucket views, are kept)."""
        for t, saved in snap["student"] + snap["teacher"]:
            t.copy_(saved)
        if snap["center"] is not None:
            self.loss_fn.center.copy_(snap["center"])
        for g, n in zip(self.opt.param_groups, snap["steps"]):
            g["step"] = n
        for p, saved in snap["moments"].items():
            st = self.opt.state.get(p, {})
            if "exp_avg" in st:
                if saved is None:
                    st["exp_avg"].zero_()
                    st["exp_avg_sq"].zero_()
                else:
                    st["exp_avg"].copy_(saved[0])
                    st["exp_avg_sq"].copy_(saved[1])
        torch.cuda.set_rng_state(snap["rng"])
        # the bf16 weight shadows were written for the warm-up's weights: refresh them now, outside the capture
        mod = self.student.module if isinstance(self.student, GradBucketDataParallel) else self.student
        for p in list(mod.parameters()) + list(self.teacher.module.parameters()):
            if id(p) in ops.shadows._d:
                ops.shadows.invalidate(p)
                ops.bf16_of(p)

    def _capture(self, epoch, momentum):
        self._multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        torch.cuda.synchronize()
        if self._graph is None and self._warmup > 0:
            # Warm-up (first capture only): eager steps on a side stream teach the allocator its working set, build
            # the optimiser / EMA tables and let the data-parallel wrapper count gradient arrivals. They are steps
            # on the live model, so everything they changed is restored afterwards: the first replay is step 0 of
            # training, exactly as in an eager run. Re-captures (new teacher temperature, frozen-set change) need
            # none of this and skip the warm-up.
            snap = self._snapshot()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(self._warmup):
                    dino_step(self.student, self.teacher, self.loss_fn, self.opt, self.static_crops, epoch, momentum,
                              self.clip_grad)
            torch.cuda.current_stream().wait_stream(side)
            self._restore(snap)
            torch.cuda.synchronize()
        # capture records launches only: it needs the optimiser's chunk tables but must NOT advance the step
        # counter / restage scalars (that is _step_prepare's job, once per replay)
        ents = self.opt._entries()
        self._graph_update = None
        if not self._multi or NCCL_IN_GRAPH["on"]:
            # ONE graph for the whole step. With several ranks the NCCL calls are captured too: every gradient
            # bucket's all-reduce is issued (async, on NCCL's own stream = a forked branch of the graph) by the
            # arrival hook of its last gradient, i.e. it overlaps the rest of backward exactly as in the eager step
            # (reference: NativeDDP's bucketed, overlapped all-reduce, train.py:634); the centre all-reduce forks off
            # right after the loss forward and is joined after the optimiser. Nothing runs between replays.
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.loss, self.student_out, self.teacher_out = _step_launch(
                        self.student, self.teacher, self.loss_fn, self.opt, self.static_crops, epoch, self.clip_grad,
                        ents)
                self._graph = g
                self.comm_mode = "nccl all-reduces captured in the step graph, overlapped with backward" \
                    if self._multi else "single rank"
            except Exception as e:  # pragma: no cover - depends on the NCCL / driver combination
                if not self._multi:
                    raise
                import warnings
                warnings.warn(f"capturing the NCCL collectives in the step graph failed ({e!r}); falling back to "
                              "compute graph -> eager all-reduces -> update graph")
                NCCL_IN_GRAPH["on"] = False
                if isinstance(self.student, GradBucketDataParallel):
                    self.student._handles = []
                    self.student._arm()
                self.loss_fn._pending = None
                torch.cuda.synchronize()
        if self._multi and not NCCL_IN_GRAPH["on"]:
            # fallback: NCCL stays out of the graphs: compute graph -> eager all-reduces -> update graph
            self._set_defer(True)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self.loss, self.student_out, self.teacher_out = _step_compute(
                    self.student, self.teacher, self.loss_fn, self.opt, self.static_crops, epoch)
            # capture executed nothing; loss_fn._pending holds the static (graph-pool) column-sum buffer that
            # every replay of the compute graph refills
            self._pending_static = self.loss_fn._pending
            self._graph_update = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_update, pool=self._graph.pool()):
                _step_update(self.student, self.teacher, self.loss_fn, self.opt, self.clip_grad, ents)
            self._set_defer(False)   # the flags only matter while capturing; eager steps stay self-contained
            self.comm_mode = "eager nccl all-reduces between a compute graph and an update graph (not overlapped)"
        self._temp = float(self.loss_fn.teacher_temp_schedule[epoch])
        self._opt_version = self.opt.table_version

    def release(self):
        """Drop the captured graphs (their outputs stay valid tensors). MUST run before
        ``dist.destroy_process_group()`` when the step was captured with NCCL collectives inside: NCCL keeps a
        communicator alive -- and ``ncclCommDestroy`` waits -- for as long as a CUDA graph holding its kernels exists
        (measured: the process hangs in destroy_process_group otherwise)."""
        torch.cuda.synchronize()
        self._graph = self._graph_update = None
        self._pending_static = None
        torch.cuda.synchronize()

    def crop_blocks(self):
        """The static input buffers as one tensor per resolution group, ``[n_crops_of_that_size, B, 3, S, S]`` each
        (the crops of a group lie back to back): ``MultiCropAugment(...)(tiles, out=step.crop_blocks())`` writes the
        augmented crops straight into the graph's inputs, and ``step(None, ...)`` replays on them."""
        blocks, i = [], 0
        sc = self.static_crops
        while i < len(sc):
            j = i
            while j + 1 < len(sc) and sc[j + 1].shape == sc[i].shape and \
                    sc[j + 1].data_ptr() == sc[j].data_ptr() + sc[j].numel() * sc[j].element_size():
                j += 1
            blocks.append(torch.as_strided(sc[i], (j - i + 1,) + tuple(sc[i].shape), (sc[i].numel(),) + tuple(sc[i].stride()),
                                           sc[i].storage_offset()))
            i = j + 1
        return tuple(blocks)

    def load(self, crops, non_blocking=True):
        """Copy a batch of crops into the graph's static input buffers (same stream as the replay)."""
        for dst, src in zip(self.static_crops, crops):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=non_blocking)

    def __call__(self, crops=None, epoch=0, momentum=0.996):
        if crops is not None:
            self.load(crops)
        if (self._graph is None or float(self.loss_fn.teacher_temp_schedule[epoch]) != self._temp
                or self.opt.table_version != self._opt_version):
            self._capture(epoch, momentum)      # the capture itself executes nothing: fall through to replay
        _step_prepare(self.student, self.teacher, self.opt, momentum)
        self._graph.replay()
        if self._graph_update is not None:
            self.loss_fn._pending = self._pending_static   # the compute graph refilled this buffer
            self._comm()
            self._graph_update.replay()
            self.loss_fn._pending = None
        return self.loss
