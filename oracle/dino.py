"""ORACLE (test infrastructure only — never imported by the product path).

PyTorch restatement of the DINO pieces the north star names but the reference does not contain
(SURVEY.md §0.3, §8a rows L1-L3 / M1): the teacher-centred cross-entropy over global/local crops,
the running-centre update, the multi-crop wrapper, the teacher EMA with its cosine momentum
schedule, and the full step (§3.3) that occupies the reference's call slots
``output = model(input)`` / ``loss = loss_fn(output, target)`` / ``model_ema.update(model)``
(train.py:1045, :1053, :1081).  Semantics follow Caron et al. 2021, Alg. 1.

PARITY UNPINNED for this file: nothing in /root/reference implements or tests these pieces (the encoder and
head they wrap ARE pinned to the reference's own outputs, see vision_transformer.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F


class DINOLoss(nn.Module):
    """§8a L1/L2. ``forward(student_output, teacher_output, epoch)``; rows are crop-major."""

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

    def forward(self, student_output, teacher_output, epoch=0):
        student_out = (student_output.float() / self.student_temp).chunk(self.ncrops)
        temp = float(self.teacher_temp_schedule[epoch])
        teacher_out = F.softmax((teacher_output.float() - self.center) / temp, dim=-1).detach().chunk(2)
        total, n_terms = 0.0, 0
        for iq, q in enumerate(teacher_out):
            for v in range(len(student_out)):
                if v == iq:
                    continue
                term = torch.sum(-q * F.log_softmax(student_out[v], dim=-1), dim=-1)
                total = total + term.mean()
                n_terms += 1
        total = total / n_terms
        self.update_center(teacher_output)
        return total

    @torch.no_grad()
    def update_center(self, teacher_output):
        batch_center = torch.sum(teacher_output.float(), dim=0, keepdim=True)
        world = 1
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(batch_center)
            world = dist.get_world_size()
        batch_center = batch_center / (len(teacher_output) * world)
        self.center = self.center * self.center_momentum + batch_center * (1 - self.center_momentum)


class MultiCropWrapper(nn.Module):
    """§8a L3: one backbone call per run of equal-resolution crops, then the head on all rows."""

    def __init__(self, backbone, head):
        super().__init__()
        backbone.fc, backbone.head = nn.Identity(), nn.Identity()
        self.backbone = backbone
        self.head = head

    def forward(self, x):
        if not isinstance(x, (list, tuple)):
            x = [x]
        sizes = torch.tensor([inp.shape[-1] for inp in x])
        idx_crops = torch.cumsum(torch.unique_consecutive(sizes, return_counts=True)[1], 0)
        start, outs = 0, []
        for end in idx_crops.tolist():
            out = self.backbone(torch.cat(list(x[start:end])))
            if isinstance(out, tuple):
                out = out[0]
            outs.append(out)
            start = end
        return self.head(torch.cat(outs))


def cosine_momentum(it: int, total_iters: int, base: float = 0.996, final: float = 1.0) -> float:
    """§8a M1: m_it = final - (final - base) * (cos(pi*it/T) + 1) / 2."""
    return final - (final - base) * (math.cos(math.pi * it / max(total_iters, 1)) + 1) / 2


def _detach_weight_norm(model):
    """Old-style nn.utils.weight_norm leaves a non-leaf ``weight`` attribute that deepcopy refuses;
    the forward pre-hook recomputes it anyway, so park a detached copy before cloning the module."""
    from torch.nn.utils.weight_norm import WeightNorm
    for m in model.modules():
        for hook in m._forward_pre_hooks.values():
            if isinstance(hook, WeightNorm) and isinstance(getattr(m, hook.name, None), torch.Tensor):
                setattr(m, hook.name, getattr(m, hook.name).detach())


class ModelEma(nn.Module):
    """timm ``ModelEmaV2`` call convention (train.py:619-620, :1081): ``.module`` is the EMA copy and
    ``update(model)`` folds every state-dict tensor; ``momentum`` overrides the fixed decay per step."""

    def __init__(self, model, decay=0.9998, device=None):
        super().__init__()
        import copy
        _detach_weight_norm(model)
        self.module = copy.deepcopy(model)
        self.module.eval()
        for p in self.module.parameters():
            p.requires_grad_(False)
        self.decay = decay
        if device is not None:
            self.module.to(device)

    @torch.no_grad()
    def update(self, model, momentum=None):
        m = self.decay if momentum is None else momentum
        for e, s in zip(self.module.state_dict().values(), model.state_dict().values()):
            if e.dtype.is_floating_point:
                e.copy_(m * e + (1.0 - m) * s.to(e.dtype))
            else:
                e.copy_(s)


def clip_grad_norm_(params, max_norm):
    return torch.nn.utils.clip_grad_norm_(params, max_norm)


def dino_step(student, teacher_ema, loss_fn, optimizer, crops, epoch=0, momentum=0.996, clip_grad=3.0,
              autocast_dtype=None):
    """One optimisation step (§3.3). ``crops`` = 2 global tensors followed by the local ones."""
    device_type = crops[0].device.type
    ctx = (torch.autocast(device_type=device_type, dtype=autocast_dtype) if autocast_dtype is not None
           else torch.autocast(device_type=device_type, enabled=False))
    with ctx:
        with torch.no_grad():
            teacher_out = teacher_ema.module(list(crops[:2]))
        student_out = student(list(crops))
        loss = loss_fn(student_out, teacher_out, epoch)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    if clip_grad:
        clip_grad_norm_([p for p in student.parameters() if p.requires_grad], clip_grad)
    optimizer.step()
    teacher_ema.update(student, momentum=momentum)
    return loss.detach(), student_out.detach(), teacher_out.detach()
