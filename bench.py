#!/usr/bin/env python
"""Headline benchmark: DINO self-supervised ViT training step, images/s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|4|5a|5b]

`--config` picks one of BASELINE.json's configurations (default 2 = configs[1], the one the metric is quoted on):
  1   ViT-Tiny/16, batch 8, 2 global 224^2 crops only              (configs[0], the reference's CPU-runnable case)
  2   ViT-S/16, batch 256/GPU, 2 x 224^2 + 10 x 96^2 crops         (configs[1]; with --gpus N: configs[2])
  4   ViT-B/16, batch 128/GPU, multi-crop                          (configs[3])
  5a  ViT-S/8 (785-token sequences), batch 64/GPU, multi-crop      (configs[4], training half)
  5b  ViT-S/8 frozen-encoder embedding of 4096 tiles per slide     (configs[4], embedding half; a step = one slide)

One process per GPU (torchrun sets RANK / LOCAL_RANK / WORLD_SIZE for N > 1). A training "step" is the whole
hot path over one batch of synthetic crops (bf16): teacher forward, student forward, fused DINO loss + centre
update, backward (bucketed NCCL all-reduce for N > 1), gradient clipping + AdamW, teacher EMA. Prints ONE JSON
line on rank 0 (contract in the task brief).

`value`      : device-resident inputs, K steps bracketed by barrier + synchronize, CUDA events, max over ranks.
`e2e`        : same step driven from PINNED HOST crops (H2D copy every step, prefetched on a copy stream)
               plus a D2H read of the loss (5b: of the feature matrix) every step.
`roofline`   : the tcgen05 GEMM kernel family (dominant kernel): algorithmic FLOPs / summed per-launch
               CUDA-event durations over one fully instrumented step, vs the measured bf16 peak.
`cpu_baseline`: the PyTorch oracle (oracle/) timed on the host cores on a bounded sample of the workload.
`--impl reference`: the reference's CPU path for this step == the oracle restatement (the reference's own
               sources for the path are Python-3.7 bytecode + un-vendored timm), all host threads,
               EXACTLY K timed steps after W warm-up steps, each step a bounded sample (``--cpu-batch`` images) of
               the same workload. The restatement's encoder and head are bit-identical to that bytecode executed by
               tests/golden/py37vm.py (tests/test_oracle_golden.py); /root/reference does not exist on the GPU box.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

MODELS = {
    "vit_tiny": dict(D=192, depth=12, heads=3),
    "vit_small": dict(D=384, depth=12, heads=6),
    "vit_base": dict(D=768, depth=12, heads=12),
}

CONFIGS = {
    "1": dict(model="vit_tiny", patch=16, batch=8, local_crops=0, kind="train", cpu_batch=8,
              label="BASELINE.json configs[0]"),
    "2": dict(model="vit_small", patch=16, batch=256, local_crops=10, kind="train", cpu_batch=4,
              label="BASELINE.json configs[1]"),
    "4": dict(model="vit_base", patch=16, batch=128, local_crops=10, kind="train", cpu_batch=2,
              label="BASELINE.json configs[3]"),
    "5a": dict(model="vit_small", patch=8, batch=64, local_crops=10, kind="train", cpu_batch=1,
               label="BASELINE.json configs[4], training half"),
    "5b": dict(model="vit_small", patch=8, batch=4096, local_crops=0, kind="embed", cpu_batch=4,
               label="BASELINE.json configs[4], embedding half"),
}

METRIC = "SSL train images/sec (ViT-S/16 DINO multi-crop)"
SUMMARY_JSON = os.path.join("profiles", "r2_step_summary.json")


def fwd_flops(model: str, S: int, patch: int):
    """(GEMM, attention) FLOPs of one encoder forward on one S x S image."""
    cfg = MODELS[model]
    D, depth = cfg["D"], cfg["depth"]
    N = (S // patch) ** 2 + 1
    return 2 * (N - 1) * 3 * patch * patch * D + depth * 24 * N * D * D, depth * 4 * N * N * D


def flops_per_sample(model: str, out_dim: int, n_local: int, patch: int = 16):
    """multiply-add = 2, backward = 2x forward, no recompute / padding credit."""
    D = MODELS[model]["D"]
    g224, a224 = fwd_flops(model, 224, patch)
    g96, a96 = fwd_flops(model, 96, patch)
    f_head = 2 * (D * 2048 + 2048 * 2048 + 2048 * 256 + 256 * out_dim)
    ncrops = 2 + n_local
    gemm = 3 * (2 * g224 + n_local * g96) + 2 * g224 + (3 * ncrops + 2) * f_head
    attn = 3 * (2 * a224 + n_local * a96) + 2 * a224
    return gemm + attn, gemm, attn


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p.get("bf16_tflops", 1590.0), p.get("bf16_tflops_sustained", 1400.0), p.get("hbm_gbs", 6650.0), "measured"
    except Exception:
        return 1590.0, 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi SM clocks and throttle reasons during the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                parts = [x.strip() for x in out.split(",")]
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for n, v in zip(names, parts[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t.start()

    def stop(self):
        self._stop.set()
        self._t.join(timeout=6)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_crops(B, n_local, device, dtype, seed, pin=False):
    g = torch.Generator(device="cpu" if pin else device).manual_seed(seed)
    dev = "cpu" if pin else device
    crops = [torch.randn(B, 3, 224, 224, generator=g, device=dev).to(dtype) for _ in range(2)]
    crops += [torch.randn(B, 3, 96, 96, generator=g, device=dev).to(dtype) for _ in range(n_local)]
    if pin:
        crops = [c.pin_memory() for c in crops]
    return crops


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle restatement on the host cores
# ------------------------------------------------------------------------------------------------
def run_cpu_oracle(cfg, out_dim, batch, steps, warmup):
    """-> (images/s, seconds per step, threads). Training configs: the whole DINO step; 5b: the frozen-encoder
    forward (train.py:1229-1230) on ``batch`` tiles."""
    from oracle import dino as odino
    from oracle import vision_transformer as ovt
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    name, patch, n_local = cfg["model"], cfg["patch"], cfg["local_crops"]
    backbone = getattr(ovt, name)(patch_size=patch)
    if cfg["kind"] == "embed":
        backbone.eval()
        tiles = torch.randn(batch, 3, 224, 224, generator=torch.Generator().manual_seed(1234))

        def one():
            with torch.no_grad():
                backbone(tiles)
    else:
        student = odino.MultiCropWrapper(backbone, ovt.DINOHead(MODELS[name]["D"], out_dim))
        teacher = odino.ModelEma(student)
        loss_fn = odino.DINOLoss(out_dim, 2 + n_local, 0.04, 0.04, 0, 10)
        opt = torch.optim.AdamW([p for p in student.parameters() if p.requires_grad], lr=5e-4, weight_decay=0.04)
        crops = make_crops(batch, n_local, "cpu", torch.float32, 1234)

        def one():
            odino.dino_step(student, teacher, loss_fn, opt, crops)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return batch / dt, dt, torch.get_num_threads()


def workload_text(cfg, args):
    m, p = cfg["model"], cfg["patch"]
    if cfg["kind"] == "embed":
        return (f"{m}/{p} frozen-encoder embedding (train.py --extract_features flow), {args.batch} tiles of 3x224x224 per "
                f"slide, forward only, bf16 ({cfg['label']})")
    extra = f", NON-DEFAULT drop_rate {args.drop_rate}" if getattr(args, "drop_rate", 0.0) > 0 else ""
    return (f"{m}/{p} DINO multi-crop 2x224+{args.local_crops}x96, batch {args.batch}/GPU, head out_dim "
            f"{args.out_dim}, bf16 ({cfg['label']}){extra}")


def make_config(cfg, args, world, total_f):
    """The `config` object of the JSON line: identical for both arms (the reference arm times a bounded sample of
    this very workload and says so under `reference_sample`)."""
    B = args.batch
    return {"workload": workload_text(cfg, args), "global_batch": B * world, "parallelism": f"dp{world}",
            "l2_policy": "inputs + activations per step (GBs) far exceed the 126 MB L2",
            "flop_per_image": total_f,
            "reference_sample": f"--impl reference and cpu_baseline time the same step on the host cores at "
                                f"{args.cpu_batch} image(s) per step (a bounded sample of the batch above), fp32"}


def run_reference(cfg, args, world, total_f):
    steps, warm = max(1, args.steps), max(0, args.warmup)
    ips, dt, threads = run_cpu_oracle(cfg, args.out_dim, args.cpu_batch, steps, warm)
    what = ("frozen-encoder forward" if cfg["kind"] == "embed"
            else f"whole DINO step (all {2 + args.local_crops} crops, loss, backward, AdamW, EMA)")
    sample = (f"{steps} timed step(s) after {warm} warm-up; each step = the {what} on a bounded sample of "
              f"{args.cpu_batch} image(s) of the workload's batch of {args.batch}, fp32, PyTorch oracle restatement "
              f"(encoder + head bit-identical to the reference's bytecode, tests/test_oracle_golden.py) on {threads} "
              f"host threads")
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": make_config(cfg, args, world, total_f),
            "sample_images_per_step": args.cpu_batch,
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline_leg(cfg, args):
    """Bounded CPU sample of this run's workload; for the default config also BASELINE config 1 IN FULL
    (ViT-Tiny/16, batch 8, 2 global crops, 5 timed steps), the case the reference plumbing is quoted on."""
    out = None
    try:
        ips, dt, threads = run_cpu_oracle(cfg, args.out_dim, args.cpu_batch, 2, 1)
        out = {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"2 timed steps after 1 warm-up of the same step at {args.cpu_batch} image(s) per step, "
                         f"fp32 oracle, {dt:.2f} s/step"}
        if args.config == "2":
            c1 = CONFIGS["1"]
            ips1, dt1, _ = run_cpu_oracle(c1, args.out_dim, c1["batch"], 5, 1)
            out["config1_full"] = {"value": ips1, "unit": "images/s", "cores": threads,
                                   "sample": f"BASELINE config 1 in full: vit_tiny/16, batch 8, 2 global 224^2 crops, "
                                             f"out_dim {args.out_dim}, 5 timed steps after 1 warm-up, {dt1:.2f} s/step"}
    except Exception as e:  # the baseline is informative; never lose the GPU line over it
        out = {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="2", choices=list(CONFIGS))
    ap.add_argument("--model", default=None, choices=list(MODELS))
    ap.add_argument("--patch", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (5b: tiles per slide)")
    ap.add_argument("--out-dim", type=int, default=65536)
    ap.add_argument("--local-crops", type=int, default=None)
    ap.add_argument("--cpu-batch", type=int, default=None, help="images per step of the bounded CPU sample")
    ap.add_argument("--embed-batch", type=int, default=512, help="5b: tiles per encoder forward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-tiles", action="store_true", help="end-to-end leg fed with uint8 256x256 tiles from pinned "
                    "host memory (H2D of the tiles + GPU multi-crop augmentation every step) instead of ready-made bf16 "
                    "crops; training configs with 224/96 crops only")
    ap.add_argument("--no-teacher-stream", action="store_true",
                    help="teacher forward on the main stream instead of a second captured stream (A/B switch)")
    ap.add_argument("--no-merge-crops", action="store_true",
                    help="one backbone pass per crop resolution instead of one pass over the packed rows (A/B switch)")
    ap.add_argument("--no-nccl-graph", action="store_true", help="N > 1: keep the NCCL all-reduces out of the step "
                    "graph (compute graph -> eager all-reduces -> update graph; A/B switch)")
    ap.add_argument("--bucket-mb", type=float, default=None, help="N > 1: gradient bucket size in MiB (default: the "
                    "wrapper's; a huge value = ONE all-reduce after the last gradient)")
    ap.add_argument("--grad-compress", default=None, choices=["none", "bf16"], help="N > 1: gradient buckets travel as "
                    "bf16 (cast, all-reduce, cast back) instead of fp32")
    ap.add_argument("--ncu-step", action="store_true", help="profiling aid: after the warm-up run ONE steady-state eager "
                    "step between cudaProfilerStart/Stop and exit (use with ncu --profile-from-start off)")
    ap.add_argument("--ln-tail", action="store_true", help="LayerNorm after each residual add in the tail of the proj / fc2 "
                    "GEMM instead of its own kernel (A/B switch, measured slower; ViT-S only)")
    ap.add_argument("--pdl", action="store_true", help="enable programmatic dependent launch (A/B switch; off by default)")
    ap.add_argument("--drop-rate", type=float, default=0.0, help="developer switch: element dropout in the student "
                    "(the reference's --drop; default 0 = BASELINE.json's configuration). A non-zero value is recorded in "
                    "config.workload: such a line is not the headline metric")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from Python instead of replaying "
                    "the captured CUDA graph of the step")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    for k_arg, k_cfg in (("model", "model"), ("patch", "patch"), ("batch", "batch"), ("local_crops", "local_crops"),
                         ("cpu_batch", "cpu_batch")):
        if getattr(args, k_arg) is None:
            setattr(args, k_arg, cfg[k_cfg])
        else:
            cfg[k_cfg] = getattr(args, k_arg)

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    n_local = args.local_crops
    if cfg["kind"] == "embed":
        g, a = fwd_flops(args.model, 224, args.patch)
        total_f, gemm_f, attn_f = g + a, g, a
    else:
        total_f, gemm_f, attn_f = flops_per_sample(args.model, args.out_dim, n_local, args.patch)

    if args.impl == "reference":
        if rank == 0:
            run_reference(cfg, args, world, total_f)
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a B200; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    import b200ssl
    from b200ssl import ops
    if args.pdl:
        b200ssl._lib.lib().b200ssl_set_pdl(1)
    if args.no_teacher_stream:
        b200ssl.dino.TEACHER_STREAM["on"] = False
    if args.no_merge_crops:
        b200ssl.dino.MERGE_CROP_GROUPS["on"] = False
    if args.no_nccl_graph:
        b200ssl.dino.NCCL_IN_GRAPH["on"] = False
    if args.ln_tail:
        ops._LN_TAIL["on"] = True

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    torch.manual_seed(0)
    D = MODELS[args.model]["D"]
    B = args.batch
    total_steps = args.steps + args.warmup
    sampler = ClockSampler(local_rank)
    peak_burst, peak_sus, hbm_peak, peak_src = read_peaks()
    gemm_events = []
    real_gemm, real_gemm_res_ln = ops.gemm, ops.gemm_res_ln

    def _timed(fn):
        def wrapper(*a, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(*a, **kw)
            e1.record()
            gemm_events.append((e0, e1))
        return wrapper

    timed_gemm, timed_gemm_res_ln = _timed(real_gemm), _timed(real_gemm_res_ln)   # the latter: GEMM + LayerNorm tail

    if cfg["kind"] == "embed":
        # ------------------------------------------------------------------ 5b: frozen-encoder embedding
        backbone = getattr(b200ssl, args.model)(patch_size=args.patch).to(device).eval()
        tiles_dev = torch.randn(B, 3, 224, 224, device=device,
                                generator=torch.Generator(device=device).manual_seed(1234 + rank)).bfloat16()
        feats_dev = torch.empty(B, D, dtype=torch.float32, device=device)

        def slide_resident(i):
            with torch.no_grad():
                for s0 in range(0, B, args.embed_batch):
                    feats_dev[s0:s0 + args.embed_batch] = backbone(tiles_dev[s0:s0 + args.embed_batch]).float()

        for i in range(args.warmup):
            slide_resident(i)
        if rank == 0:
            sampler.start()
        l0 = ops.launch_count()
        ms_step = timed(slide_resident, args.steps) / args.steps
        launches = ops.launch_count() - l0
        clocks = sampler.stop() if rank == 0 else None
        value = B * world / (ms_step / 1e3)
        e2e = None
        if not args.no_e2e:
            tiles_host = tiles_dev.cpu().pin_memory()
            b200ssl.embed_tiles(backbone, tiles_host, batch_size=args.embed_batch)

            def slide_e2e(i):
                b200ssl.embed_tiles(backbone, tiles_host, batch_size=args.embed_batch)   # returns host features

            ms_e2e = timed(slide_e2e, args.steps) / args.steps
            e2e = {"value": B * world / (ms_e2e / 1e3), "unit": "images/s",
                   "h2d_bytes_per_step": tiles_host.numel() * 2, "d2h_bytes_per_step": B * D * 4, "ms_per_step": ms_e2e}
        ops.gemm, ops.gemm_res_ln = timed_gemm, timed_gemm_res_ln
        slide_resident(0)
        torch.cuda.synchronize()
        ops.gemm, ops.gemm_res_ln = real_gemm, real_gemm_res_ln
        step_api = "b200ssl.embed_tiles / VisionTransformer.forward under no_grad (eager launches)"
        host_ms = None
    else:
        # ------------------------------------------------------------------ training step
        vit_kw = {"drop_rate": args.drop_rate} if args.drop_rate > 0 else {}
        student = b200ssl.MultiCropWrapper(getattr(b200ssl, args.model)(patch_size=args.patch, **vit_kw),
                                           b200ssl.DINOHead(D, args.out_dim)).to(device)
        teacher = b200ssl.ModelEma(student)
        ddp_kw = {}
        if args.bucket_mb is not None:
            ddp_kw["bucket_mb"] = args.bucket_mb
        if args.grad_compress is not None:
            ddp_kw["compress"] = args.grad_compress
        ddp = b200ssl.GradBucketDataParallel(student, **ddp_kw)
        loss_fn = b200ssl.DINOLoss(args.out_dim, 2 + n_local, 0.04, 0.04, 0, 10).to(device)
        opt = b200ssl.FusedAdamW(b200ssl.param_groups_wd(student, 0.04), lr=5e-4 * args.batch * world / 256.0)
        crops = make_crops(B, n_local, device, torch.bfloat16, 1234 + 1000 * rank)

        def eager_step(cr, it):
            m = b200ssl.cosine_momentum(it, max(total_steps, 1))
            return b200ssl.dino_step(ddp, teacher, loss_fn, opt, cr, epoch=0, momentum=m, clip_grad=3.0)

        if args.ncu_step:
            for i in range(max(args.warmup, 2)):
                eager_step(crops, i)
            torch.cuda.synchronize()
            b200ssl.dino.TEACHER_STREAM["on"] = False     # one stream: serialised per-kernel times
            torch.cuda.cudart().cudaProfilerStart()
            eager_step(crops, args.warmup)
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
            return 0
        use_graph = not args.no_graph
        graphed = b200ssl.GraphedDinoStep(ddp, teacher, loss_fn, opt, crops, clip_grad=3.0) if use_graph else None

        def step(cr, it):
            """The public step API: GraphedDinoStep (whole step = one CUDA-graph replay) or eager dino_step.
            cr=None replays on the crops already resident in the graph's input buffers (the `value` measurement:
            inputs in HBM before the timed region); the e2e path passes host crops and pays the copy every step."""
            m = b200ssl.cosine_momentum(it, max(total_steps, 1))
            if graphed is not None:
                return graphed(cr, epoch=0, momentum=m), None, None
            return b200ssl.dino_step(ddp, teacher, loss_fn, opt, cr, epoch=0, momentum=m, clip_grad=3.0)

        for i in range(args.warmup):
            step(crops, i)
        if rank == 0:
            sampler.start()
        host_t = [0.0]
        resident = None if graphed is not None else crops   # graph mode: the crops sit in the static inputs

        def step_host_timed(i):
            t0 = time.perf_counter()
            step(resident, args.warmup + i)
            host_t[0] += time.perf_counter() - t0

        l0 = ops.launch_count()
        ms_total = timed(step_host_timed, args.steps)
        launches = ops.launch_count() - l0
        host_ms = host_t[0] * 1e3 / args.steps   # host time to ENQUEUE one step (GPU runs asynchronously)
        if graphed is not None:
            # replays launch the captured kernels without going through Python: count them on one eager step
            l0 = ops.launch_count()
            eager_step(crops, total_steps)
            torch.cuda.synchronize()
            launches = (ops.launch_count() - l0) * args.steps
        clocks = sampler.stop() if rank == 0 else None
        ms_step = ms_total / args.steps
        value = B * world / (ms_step / 1e3)

        # ---- end-to-end: pinned host crops -> H2D (prefetched on a copy stream) -> step -> D2H loss
        e2e = None
        if not args.no_e2e and args.e2e_tiles and graphed is not None:
            # production input path: uint8 tiles + the per-crop parameter table travel host -> device (prefetched on a
            # copy stream, two slots), MultiCropAugment writes the crops straight into the graph's static inputs
            aug = b200ssl.MultiCropAugment("pcbnfrs", n_local=n_local)
            gen = torch.Generator().manual_seed(4321 + rank)
            tiles_host = torch.randint(0, 256, (B, 256, 256, 3), dtype=torch.uint8, generator=gen).pin_memory()
            params_host = aug.sample_params(B, gen).pin_memory()
            h2d_bytes = tiles_host.numel() + params_host.numel() * 4
            copy_stream = torch.cuda.Stream()
            t_bufs = [torch.empty_like(tiles_host, device=device) for _ in range(2)]
            p_bufs = [torch.empty_like(params_host, device=device) for _ in range(2)]
            ready = [torch.cuda.Event() for _ in range(2)]
            consumed = [torch.cuda.Event() for _ in range(2)]
            blocks = graphed.crop_blocks()
            if n_local == 0:
                blocks = (blocks[0], torch.empty(0, dtype=torch.bfloat16, device=device))

            def prefetch(slot):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[slot])
                    t_bufs[slot].copy_(tiles_host, non_blocking=True)
                    p_bufs[slot].copy_(params_host, non_blocking=True)
                    ready[slot].record(copy_stream)

            losses = []

            def e2e_step(i):
                slot = i & 1
                torch.cuda.current_stream().wait_event(ready[slot])
                aug(t_bufs[slot], params=p_bufs[slot], out=blocks)
                consumed[slot].record(torch.cuda.current_stream())
                loss, _, _ = step(None, args.warmup + i)
                prefetch(slot)
                losses.append(loss.item())

            for s_ in range(2):
                consumed[s_].record(torch.cuda.current_stream())
                prefetch(s_)
            for i in range(2):
                e2e_step(i)
            ms_e2e = timed(e2e_step, args.steps) / args.steps
            e2e = {"value": B * world / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": h2d_bytes,
                   "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e,
                   "input": "uint8 256x256x3 tiles + per-crop parameter table; GPU multi-crop augmentation "
                            "(b200ssl.MultiCropAugment 'pcbnfrs') inside the timed step"}
        elif not args.no_e2e:
            host = make_crops(B, n_local, device, torch.bfloat16, 4321 + rank, pin=True)
            h2d_bytes = sum(c.numel() * c.element_size() for c in host)
            copy_stream = torch.cuda.Stream()
            bufs = [[torch.empty_like(c, device=device) for c in host] for _ in range(2)]
            ready = [torch.cuda.Event() for _ in range(2)]
            consumed = [torch.cuda.Event() for _ in range(2)]

            def prefetch(slot):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[slot])
                    for dst, src in zip(bufs[slot], host):
                        dst.copy_(src, non_blocking=True)
                    ready[slot].record(copy_stream)

            losses = []

            def e2e_step(i):
                slot = i & 1
                torch.cuda.current_stream().wait_event(ready[slot])
                loss, _, _ = step(bufs[slot], args.warmup + i)
                consumed[slot].record(torch.cuda.current_stream())
                prefetch(slot)                 # refill this slot for step i+2 while step i+1 computes
                losses.append(loss.item())     # D2H read of the step's result (4 bytes) every step

            for s in range(2):
                consumed[s].record(torch.cuda.current_stream())
                prefetch(s)
            for i in range(2):
                e2e_step(i)
            ms_e2e = timed(e2e_step, args.steps) / args.steps
            e2e = {"value": B * world / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": h2d_bytes,
                   "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e}

        # ---- roofline of the dominant kernel family (tcgen05 GEMM): one instrumented step on ONE stream (with the
        # teacher forward running concurrently the event-bracketed durations would include the other's share)
        ops.gemm, ops.gemm_res_ln = timed_gemm, timed_gemm_res_ln
        teacher_branch = b200ssl.dino.TEACHER_STREAM["on"]
        b200ssl.dino.TEACHER_STREAM["on"] = False
        eager_step(crops, total_steps)
        torch.cuda.synchronize()
        b200ssl.dino.TEACHER_STREAM["on"] = teacher_branch
        ops.gemm, ops.gemm_res_ln = real_gemm, real_gemm_res_ln
        step_api = "b200ssl.GraphedDinoStep (CUDA graph replay)" if use_graph else "b200ssl.dino_step (eager)"
        if world > 1:
            step_api += "; " + (graphed.comm_mode if graphed is not None else
                                "nccl bucket all-reduces launched from backward hooks (overlapped)")
            step_api += f"; {len(ddp.buckets)} gradient bucket(s), {'bf16' if ddp.compress else 'fp32'} on the wire"

    gemm_ms = sum(a.elapsed_time(b) for a, b in gemm_events)
    gemm_tflops = gemm_f * B / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    # DRAM traffic of the same kernel family from the committed ncu capture of one config-2 step (profiles/): per
    # launch, like `achieved` (family total per step / launches per step)
    traffic, traffic_src = None, None
    if args.config == "2" and args.batch == 256:
        try:
            with open(os.path.join(ROOT, SUMMARY_JSON)) as f:
                summ = json.load(f)
            gk = [k for k in summ["kernels"] if k["kernel"].startswith("gemm_kernel")]
            n_l = sum(k["launches_per_step"] for k in gk)
            traffic = sum(k["dram_read_mb_per_step"] + k["dram_write_mb_per_step"] for k in gk) * 1e6 / max(n_l, 1)
            traffic_src = (f"{SUMMARY_JSON}: dram__bytes_read.sum + dram__bytes_write.sum summed over the "
                           f"{n_l:.0f} gemm_kernel launches of one step, divided by the launch count")
        except Exception:
            pass
    roofline = {"bound": "tensor", "kernel": "gemm_kernel<BN,EPI> (tcgen05, all Linear fprop/dgrad/wgrad)",
                "achieved": gemm_tflops, "peak": peak_sus, "unit": "TFLOP/s", "frac": gemm_tflops / peak_sus,
                "traffic": traffic, "traffic_source": traffic_src,
                "flops_per_launch": gemm_f * B / max(len(gemm_events), 1),
                "peak_source": f"{peak_src} (sustained bf16; burst {peak_burst})",
                "launches_per_step": len(gemm_events), "gemm_ms_per_step": gemm_ms,
                "gemm_share_of_step": gemm_ms / ms_step,
                "step_tflops": total_f * B / (ms_step / 1e3) / 1e12,
                "step_frac_of_peak": total_f * B / (ms_step / 1e3) / 1e12 / peak_sus,
                "step_frac_of_burst_peak": total_f * B / (ms_step / 1e3) / 1e12 / peak_burst}

    def shutdown():
        """N > 1: drop the captured step graph BEFORE the process group (NCCL does not let go of a communicator that a
        live CUDA graph still references: destroy_process_group would hang), then leave without running the
        interpreter's teardown (which would destroy the communicator in an arbitrary order)."""
        if world > 1:
            if cfg["kind"] != "embed" and graphed is not None:
                graphed.release()
            torch.cuda.synchronize()
            dist.barrier()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    # N > 1: the timed value is the max over ranks, and the ranks run in lock-step through the all-reduces, so the
    # slowest GPU of the box sets the pace. Each rank's own GEMM time (event-bracketed launches of its instrumented
    # step, untouched by the collectives) shows how far the 8 power-capped GPUs of one box are apart.
    if world > 1:
        mine = torch.tensor([gemm_ms], device=device)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        roofline["rank_gemm_ms_per_step"] = [round(float(t.item()), 3) for t in allr]

    if rank != 0:
        shutdown()
        return 0

    line = {"metric": METRIC, "value": value, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": make_config(cfg, args, world, total_f), "step_api": step_api,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "host_enqueue_ms_per_step": host_ms,
            "roofline": roofline}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_leg(cfg, args)
    print(json.dumps(line), flush=True)
    shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
