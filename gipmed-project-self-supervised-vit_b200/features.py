"""Frozen-encoder embedding of WSI tiles and the two on-disk formats the reference hands to its MIL stage
(BASELINE config 5, second half).

Reference flow: ``train.py --extract_features`` (train.py:530-533) swaps the classifier for Identity, runs every
tile of a slide through the encoder under ``torch.no_grad()`` in batches (train.py:1202-1282) and saves
``<slide>_features.pt`` = ``np.zeros((1, D))`` followed by one row per tile (train.py:1203, 1264, 1282). The MIL
data sets (datasets.py:1043-1092) read pickled tuples ``(labels, targets, scores, patch_scores, slide_names,
features[, batch_number[, tile_location]])`` with ``features`` of shape ``[num_slides, 1, max_tiles, D]``, NaN
padded behind each slide's last tile (the first NaN of feature 0 marks the tile count, datasets.py:1088-1092).

The encoder forward is the b200ssl hot path (no-grad mode: nothing is saved for backward, GELU-only epilogue);
everything else here is host-side format code.
"""
from __future__ import annotations

import pickle

import numpy as np
import torch

from . import ops


@torch.no_grad()
def embed_tiles(backbone, tiles: torch.Tensor, batch_size: int = 512) -> torch.Tensor:
    """CLS embeddings of ``tiles`` ([n, 3, H, W], host or device, any float dtype) -> fp32 ``[n, D]`` on the host.

    Batches are staged through two pinned host buffers and copied on a side stream while the previous batch
    computes; the encoder runs in eval / no-grad mode in bf16."""
    if tiles.ndim != 4:
        raise ValueError("embed_tiles expects [n, 3, H, W]")
    dev = next(backbone.parameters()).device
    ops.require_cuda(next(backbone.parameters()), "embed_tiles")
    was_training = backbone.training
    backbone.eval()
    n = tiles.shape[0]
    feats = None
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    host_side = not tiles.is_cuda
    stage = [None, None]
    dev_buf = [None, None]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    copied = [None, None]   # recorded behind each slot's host-to-device copy; the HOST waits on it before reuse

    def upload(i, slot):
        chunk = tiles[i:i + batch_size]
        if not host_side:
            dev_buf[slot] = chunk.to(device=dev, dtype=torch.bfloat16)
            return
        if stage[slot] is None or stage[slot].shape[0] < chunk.shape[0]:
            stage[slot] = torch.empty((batch_size,) + tuple(tiles.shape[1:]), dtype=torch.bfloat16).pin_memory()
            dev_buf[slot] = torch.empty((batch_size,) + tuple(tiles.shape[1:]), dtype=torch.bfloat16, device=dev)
        # the pinned slot is written by the CPU right now, not in stream order: wait until the asynchronous copy that
        # last read it has executed (a device-side wait alone would let the host run two batches ahead and overwrite it)
        if copied[slot] is not None:
            copied[slot].synchronize()
        stage[slot][:chunk.shape[0]].copy_(chunk)        # host-side cast to bf16 into pinned memory
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[slot])           # the batch that last used this slot's device buffer is consumed
            dev_buf[slot][:chunk.shape[0]].copy_(stage[slot][:chunk.shape[0]], non_blocking=True)
            ready[slot].record(copy_stream)
            if copied[slot] is None:
                copied[slot] = torch.cuda.Event()
            copied[slot].record(copy_stream)

    for s in range(2):
        done[s].record(main)
    starts = list(range(0, n, batch_size))
    if starts:
        upload(starts[0], 0)
    for bi, i in enumerate(starts):
        slot = bi & 1
        if bi + 1 < len(starts):
            upload(starts[bi + 1], slot ^ 1)
        m = min(batch_size, n - i)
        if host_side:
            main.wait_event(ready[slot])
        out = backbone(dev_buf[slot][:m]).float()
        done[slot].record(main)
        if feats is None:
            feats = torch.empty(n, out.shape[1], dtype=torch.float32).pin_memory()
        feats[i:i + m].copy_(out, non_blocking=True)
    torch.cuda.synchronize(dev)
    if was_training:
        backbone.train()
    return feats if feats is not None else torch.empty(0, 0)


def save_slide_features(path: str, feats) -> np.ndarray:
    """``<slide>_features.pt`` exactly as the reference writes it (train.py:1203,1264,1282): a float64 numpy array
    whose first row is zeros, followed by one row per tile, stored with ``torch.save``."""
    f = feats.detach().cpu().numpy() if isinstance(feats, torch.Tensor) else np.asarray(feats)
    arr = np.concatenate((np.zeros((1, f.shape[1])), f.astype(np.float64)), axis=0)
    torch.save(arr, path)
    return arr


def pack_mil_inference_file(path: str, slide_names, per_slide_feats, targets, labels=None, scores=None,
                            patch_scores=None, tile_locations=None) -> tuple:
    """One pickled inference tuple in the layout ``datasets.py:1043-1060`` reads:
    ``(labels, targets, scores, patch_scores, slide_names, features[, batch_number, tile_location])`` with
    ``features [num_slides, 1, max_tiles, D]`` / ``patch_scores [num_slides, max_tiles]`` NaN padded."""
    num = len(slide_names)
    if num == 0 or len(per_slide_feats) != num:
        raise ValueError("pack_mil_inference_file: one feature matrix per slide is required")
    mats = [f.detach().cpu().numpy() if isinstance(f, torch.Tensor) else np.asarray(f) for f in per_slide_feats]
    D = mats[0].shape[1]
    max_tiles = max(m.shape[0] for m in mats)
    features = np.full((num, 1, max_tiles, D), np.nan, dtype=np.float32)
    ps = np.full((num, max_tiles), np.nan, dtype=np.float32)
    for i, m in enumerate(mats):
        features[i, 0, :m.shape[0]] = m
        if patch_scores is not None:
            ps[i, :m.shape[0]] = np.asarray(patch_scores[i], dtype=np.float32)
    targets = np.asarray(targets).reshape(num)
    labels = np.asarray(labels).reshape(num) if labels is not None else np.zeros(num)
    scores = np.asarray(scores).reshape(num) if scores is not None else np.zeros(num)
    names = np.asarray(list(slide_names))
    if tile_locations is None:
        payload = (labels, targets, scores, ps, names, features)
    else:
        loc = np.full((num, max_tiles, 2), np.nan, dtype=np.float32)
        for i, t in enumerate(tile_locations):
            t = np.asarray(t, dtype=np.float32)
            loc[i, :t.shape[0]] = t
        payload = (labels, targets, scores, ps, names, features, 0, loc)
    with open(path, "wb") as fh:
        pickle.dump(payload, fh)
    return payload
