"""CPU tests of the host side: the C-ABI library loads and exports exactly what include/b200ssl.h
declares, the product path refuses to run without CUDA (no fallback), bench.py's FLOP accounting, and
the data-parallel bucket logic on a 2-rank gloo group."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "b200ssl.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200ssl_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import b200ssl
    h = b200ssl._lib.lib()                     # raises if the .so is missing: there is no fallback
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(h, name), f"{name} declared in include/b200ssl.h but not exported"
    assert sorted(b200ssl._lib.exported_symbols()) == declared
    assert h.b200ssl_version() >= 100
    nm = subprocess.run(["nm", "-D", "--defined-only", b200ssl._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (b200ssl_\w+)", nm)))
    assert exported == declared, "library exports symbols the header does not declare (or vice versa)"


def test_library_holds_blackwell_kernels():
    """The shipped cubin must contain tcgen05 MMA / TMEM / TMA instructions (SASS mnemonics)."""
    import b200ssl
    r = subprocess.run(["cuobjdump", "-sass", b200ssl._lib.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    sass = r.stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG"):
        assert mnemonic in sass, mnemonic


def test_mma_issue_code_is_straight_line():
    """Regression guard for the round-2 issue-path finding (profiles/README.md item 10): every tcgen05.mma must be issued
    straight -- by a thread chosen with elect.sync, with uniform operands -- not inside the per-active-lane
    ELECT / R2UR.BROADCAST / BRA.U.ANY loop the compiler wraps around uniform-datapath instructions under `lane == 0`."""
    import b200ssl
    r = subprocess.run(["cuobjdump", "-sass", b200ssl._lib.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "R2UR.BROADCAST" not in r.stdout
    looped = {}
    for fn in re.split(r"\n\s*Function : ", r.stdout)[1:]:
        lines = fn.splitlines()
        n = sum(1 for i, l in enumerate(lines) if "UTCHMMA" in l
                and any("BRA.U.ANY" in w or "ELECT" in w for w in lines[max(0, i - 6):i + 6]))
        if n:
            looped[lines[0][:80]] = n
    assert not looped, f"tcgen05.mma issued inside per-lane loops: {looped}"


def test_gemm_tile_coordinates_carried_incrementally_match_the_closed_form():
    """csrc/gemm.cu TileIter (default mode) carries (n_blk, m_unit, ks) along with additions instead of dividing the
    linear tile index for every tile; this is the same update in Python against t % n, (t / n) % m, t / (m * n)."""
    import itertools
    for step, nnb, nmu, ksplits in itertools.product((1, 2, 37, 74, 148), (1, 2, 3, 6, 8), (1, 3, 37, 394, 764), (1, 4, 13)):
        total = nmu * nnb * ksplits
        for start in {0, 1, step - 1}:
            t = start
            if t >= total:
                continue
            step_n, step_m = step % nnb, step // nnb
            n_blk, m_lin = t % nnb, t // nnb
            ks = m_lin // nmu
            m_unit = m_lin - ks * nmu
            while t < total:
                assert (n_blk, m_unit, ks) == (t % nnb, (t // nnb) % nmu, t // (nmu * nnb)), (step, nnb, nmu, ksplits, t)
                t += step
                n_blk += step_n
                m_unit += step_m
                if n_blk >= nnb:
                    n_blk -= nnb
                    m_unit += 1
                while m_unit >= nmu:
                    m_unit -= nmu
                    ks += 1


def test_ops_refuse_cpu_tensors():
    import b200ssl
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b200ssl.vit_tiny()(torch.randn(1, 3, 224, 224))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b200ssl.DINOHead(192, 256)(torch.randn(2, 192))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b200ssl.DINOLoss(256, 2, 0.04, 0.04, 0, 1)(torch.randn(4, 256), torch.randn(4, 256))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b200ssl.Mlp(64)(torch.randn(2, 64))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gipmed-project-self-supervised-vit_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_flop_accounting_matches_baseline_md():
    sys.path.insert(0, ROOT)
    import bench
    tot, gemm, attn = bench.flops_per_sample("vit_small", 65536, 10)
    assert abs(tot / 1e9 - 123.80) < 0.05
    assert abs(bench.flops_per_sample("vit_base", 65536, 10)[0] / 1e9 - 474.11) < 0.1
    assert abs(bench.flops_per_sample("vit_tiny", 65536, 10)[0] / 1e9 - 34.20) < 0.05
    assert gemm + attn == tot and attn / tot < 0.1


def test_teacher_temp_schedule_and_param_groups():
    import b200ssl
    from oracle import dino as odino
    a = b200ssl.DINOLoss(64, 4, 0.04, 0.07, 3, 10).teacher_temp_schedule
    b = odino.DINOLoss(64, 4, 0.04, 0.07, 3, 10).teacher_temp_schedule
    assert list(a) == list(b) and len(a) == 10
    groups = b200ssl.param_groups_wd(b200ssl.vit_tiny(), 0.04)
    assert groups[1]["weight_decay"] == 0.0 and all(p.ndim == 1 for p in groups[1]["params"])
    assert all(p.ndim > 1 for p in groups[0]["params"])


def test_position_table_resize_matches_oracle():
    """interpolate_pos_encoding applies the bicubic resize as a cached matrix; it must equal the oracle's
    F.interpolate call (VT.pyc@L213-233, incl. the +0.1 fudge) for every crop size the step uses."""
    import b200ssl
    from oracle import vision_transformer as ovt
    torch.manual_seed(0)
    ours, ref = b200ssl.vit_tiny(), ovt.vit_tiny()
    ref.load_state_dict(ours.state_dict())
    for size in (96, 224, 256, 160):
        n = (size // 16) ** 2 + 1
        x = torch.zeros(1, n, 192)
        a = ours.interpolate_pos_encoding(x, size, size)
        b = ref.interpolate_pos_encoding(x, size, size)
        assert a.shape == b.shape and torch.allclose(a, b, atol=2e-6), size
    # gradients flow back to pos_embed through the matrix form
    ours.interpolate_pos_encoding(torch.zeros(1, 37, 192), 96, 96).sum().backward()
    assert ours.pos_embed.grad is not None and ours.pos_embed.grad.abs().sum() > 0


def test_crop_buffers_group_without_copy():
    """alloc_crop_buffers lays equal-resolution crops back to back; the multi-crop wrapper then views each group
    instead of concatenating it, and falls back to torch.cat for ordinary (separate) tensors."""
    from b200ssl import dino
    ex = [torch.zeros(4, 3, 8, 8) for _ in range(2)] + [torch.zeros(4, 3, 4, 4) for _ in range(3)]
    bufs = dino.alloc_crop_buffers(ex)
    assert [tuple(b.shape) for b in bufs] == [tuple(e.shape) for e in ex]
    for i, b in enumerate(bufs):
        b.fill_(float(i))
    g = dino._cat_or_view(bufs[:2])
    assert g.shape == (8, 3, 8, 8) and g.data_ptr() == bufs[0].data_ptr()          # a view, not a copy
    assert torch.equal(g, torch.cat(bufs[:2]))
    l = dino._cat_or_view(bufs[2:])
    assert l.shape == (12, 3, 4, 4) and l.data_ptr() == bufs[2].data_ptr() and torch.equal(l, torch.cat(bufs[2:]))
    sep = [torch.randn(4, 3, 4, 4) for _ in range(3)]
    c = dino._cat_or_view(sep)
    assert torch.equal(c, torch.cat(sep)) and c.data_ptr() != sep[0].data_ptr()
    assert torch.equal(dino._cat_or_view([bufs[3], bufs[2]]), torch.cat([bufs[3], bufs[2]]))   # wrong order -> cat


def test_packed_row_segments():
    """ops.segments: row ranges of crop groups packed back to back (the merged multi-crop pass)."""
    from b200ssl import ops
    assert ops.segments(4, 197) == [(4, 197, 0)]
    assert ops.segments((512, 2560), (197, 37)) == [(512, 197, 0), (2560, 37, 512 * 197)]
    assert ops.segments([2, 3, 1], [5, 7, 11]) == [(2, 5, 0), (3, 7, 10), (1, 11, 31)]


def test_cosine_scheduler_and_apply():
    import b200ssl
    lr = b200ssl.cosine_scheduler(1e-3, 1e-6, epochs=10, niter_per_ep=7, warmup_epochs=2)
    assert len(lr) == 70 and lr[0] == 0.0 and abs(lr[14] - 1e-3) < 1e-12 and abs(lr[-1] - 1e-6) < 1e-5
    assert all(lr[i] <= lr[i + 1] for i in range(13)) and all(lr[i] >= lr[i + 1] for i in range(14, 69))
    wd = b200ssl.cosine_scheduler(0.04, 0.4, epochs=10, niter_per_ep=7)
    model = torch.nn.Linear(4, 4)
    opt = torch.optim.AdamW(b200ssl.param_groups_wd(model, 0.04), lr=1.0)
    b200ssl.apply_schedules(opt, 20, lr, wd)
    assert opt.param_groups[0]["lr"] == float(lr[20]) and opt.param_groups[0]["weight_decay"] == float(wd[20])
    assert opt.param_groups[1]["weight_decay"] == 0.0          # biases / norms stay undecayed


def test_feature_file_formats(tmp_path):
    """The two on-disk formats of the frozen-encoder embedding flow, read back the way the reference reads them
    (train.py:1203,1282 for <slide>_features.pt; datasets.py:1043-1092 for the MIL inference pickles)."""
    import pickle
    import numpy as np
    import b200ssl
    feats = torch.randn(5, 384)
    arr = b200ssl.save_slide_features(str(tmp_path / "s_features.pt"), feats)
    back = torch.load(str(tmp_path / "s_features.pt"), weights_only=False)
    assert back.shape == (6, 384) and (back[0] == 0).all() and np.allclose(back[1:], feats.numpy())
    assert np.array_equal(arr, back)
    per_slide = [torch.randn(3, 384), torch.randn(7, 384)]
    b200ssl.pack_mil_inference_file(str(tmp_path / "inf.data"), ["a.mrxs", "b.mrxs"], per_slide, targets=[1, 0])
    with open(tmp_path / "inf.data", "rb") as fh:
        data = pickle.load(fh)
    assert len(data) == 6
    labels, targets, scores, patch_scores, slide_names, features = data
    assert features.shape == (2, 1, 7, 384) and patch_scores.shape == (2, 7) and list(slide_names) == ["a.mrxs", "b.mrxs"]
    # the reference finds the tile count of a slide as the first NaN of feature 0 (datasets.py:1088-1092)
    for i, m in enumerate(per_slide):
        nan_idx = np.argwhere(np.isnan(features[i, :, :, 0])).tolist()
        first_nan = nan_idx[0][1] if nan_idx else features.shape[2]
        assert first_nan == m.shape[0] and np.allclose(features[i, 0, :first_nan], m.numpy())
    data8 = b200ssl.pack_mil_inference_file(str(tmp_path / "inf8.data"), ["a.mrxs"], per_slide[:1], targets=[1],
                                            tile_locations=[[(0, 0), (256, 0), (0, 256)]])
    assert len(data8) == 8 and data8[7].shape == (1, 3, 2)


_DDP_SCRIPT = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import b200ssl
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
torch.manual_seed(1234 + rank)                      # different init per rank: the wrapper must broadcast rank 0
model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 8))
compress = sys.argv[2] if len(sys.argv) > 2 else None
ddp = b200ssl.GradBucketDataParallel(model, bucket_mb=0.001, compress=compress)   # tiny buckets -> several all-reduces
assert len(ddp.buckets) > 1
tol = 1e-5 if compress is None else 2e-2            # bf16 on the wire: 2^-9 relative per element
ref = [p.detach().clone() for p in model.parameters()]
gathered = [torch.zeros_like(ref[0]) for _ in range(world)]
dist.all_gather(gathered, ref[0])
assert all(torch.equal(g, gathered[0]) for g in gathered), "parameters were not broadcast from rank 0"
for step in range(3):                               # step 0 learns the arrival counts, steps 1-2 overlap
    torch.manual_seed(100 * step + rank)
    x = torch.randn(4, 16)
    ddp.zero_grad()
    ddp(x).pow(2).sum().backward()
    ddp.finish()
    local = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 8))
    local.load_state_dict(model.state_dict())
    total = [torch.zeros_like(p) for p in local.parameters()]
    for r in range(world):                          # serial reference: mean over every rank's batch
        torch.manual_seed(100 * step + r)
        xr = torch.randn(4, 16)
        local.zero_grad()
        local(xr).pow(2).sum().backward()
        for t, p in zip(total, local.parameters()):
            t += p.grad / world
    for t, p in zip(total, model.parameters()):
        assert torch.allclose(p.grad, t, atol=tol, rtol=tol), (step, (p.grad - t).abs().max())
        others = [torch.zeros_like(p.grad) for _ in range(world)]
        dist.all_gather(others, p.grad.contiguous())
        assert all(torch.equal(o, others[0]) for o in others), "replicas hold different gradients"
        assert p.grad.data_ptr() != 0 and p.grad.is_contiguous()
dist.destroy_process_group()
print("ok", rank)
'''


@pytest.mark.parametrize("compress", [None, "bf16"])
def test_grad_bucket_data_parallel_gloo_world2(tmp_path, compress):
    script = tmp_path / "ddp_check.py"
    script.write_text(_DDP_SCRIPT)
    port = "29533" if compress is None else "29534"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", port, str(script), ROOT]
                       + ([compress] if compress else []),
                       capture_output=True, text=True, env=env, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok") == 2


def test_c_abi_rejects_bad_arguments_before_touching_the_device():
    """Every entry point validates shapes / alignment / ranges BEFORE its first CUDA call and reports through the
    return code + b200ssl_last_error() (never throws across the C boundary; SURVEY.md §8b "Errors"): checked here,
    without a GPU, by calling the real library with arguments that must be refused."""
    import ctypes
    from b200ssl import _lib
    h = _lib.lib()
    buf = (ctypes.c_char * 4096)()
    p = ctypes.addressof(buf)
    p = (p + 255) & ~255                     # a 256-byte aligned host address: never dereferenced by the checks
    f3 = (ctypes.c_float * 3)(0.5, 0.5, 0.5)
    bad_std = (ctypes.c_float * 3)(0.5, 0.0, 0.5)
    cases = [
        ("b200ssl_gemm", (p, 64, 0, p, 64, 0, p, 64, None, None, None, 0, 128, 100, 64, 0, 1, 0, None), "multiple of 64"),
        ("b200ssl_gemm", (p, 64, 0, p, 64, 0, p, 64, None, None, None, 0, 0, 64, 64, 0, 1, 0, None), "empty problem"),
        ("b200ssl_gemm", (p + 2, 64, 0, p, 64, 0, p, 64, None, None, None, 0, 128, 64, 64, 0, 1, 0, None), "16-byte aligned"),
        ("b200ssl_gemm", (p, 64, 0, p, 64, 0, p, 64, None, None, None, 0, 128, 64, 64, 9, 1, 0, None), "unknown epilogue"),
        ("b200ssl_gemm", (p, 64, 0, p, 64, 0, p, 64, None, None, None, 0, 128, 64, 64, 1, 1, 0, None), "needs D2"),
        ("b200ssl_attention_fwd", (p, p, p, 2, 197, 6, 32, 0.125, None), "head_dim 32 unsupported"),
        ("b200ssl_attention_bwd", (p, p, p, p, p, 2, 197, 6, 128, 0.125, None), "head_dim 128 unsupported"),
        ("b200ssl_layernorm_fwd", (p, 1, p, p, p, p, p, 0, 384, 1e-6, None), "no rows"),
        ("b200ssl_patchify", (p, p, 2, 3, 224, 224, 12, None), "must divide"),
        ("b200ssl_assemble_tokens", (p, p, p, p, 2, 196, 100, None), "multiple of 8"),
        ("b200ssl_scale_rows", (p, p, p, 16, 100, None), "multiple of 8"),
        ("b200ssl_colsum", (p, 100, p, 16, 100, 0, None), "multiples of 8"),
        ("b200ssl_center_update", (p, p, 1024, 0, 0.9, None), "total_rows"),
        ("b200ssl_dino_loss_fwd", (p, p, p, p, p, p, 4, 4, 1001, 0.1, 0.04, None), "multiple of 8"),
        ("b200ssl_multicrop_augment", (p, p, p, p, 2, 2, 10, 223, 96, f3, f3, None), "must be even"),
        ("b200ssl_multicrop_augment", (p, p, p, p, 0, 2, 10, 224, 96, f3, f3, None), "empty problem"),
        ("b200ssl_multicrop_augment", (p, p, p, p, 2, 2, 10, 224, 96, f3, bad_std, None), "std > 0"),
        ("b200ssl_dropout", (p, p, None, 1001, 0, 0.1, p, 0, None), "multiple of 8"),
        ("b200ssl_dropout", (p, p, None, 1024, 0, 1.5, p, 0, None), "outside [0, 1]"),
        ("b200ssl_dropout", (p, p, None, 1024, 0, 0.1, None, 0, None), "seed"),
        ("b200ssl_dropout", (p, p, p, 1024, 1, 0.1, p, 0, None), "second"),
        ("b200ssl_dropout_residual", (p, p, None, p, 16, 100, 0.1, p, 0, None), "multiple of 8"),
        ("b200ssl_set_attn_stream", (7,), "stream mode"),
        ("b200ssl_set_gemm_cluster", (3,), "cluster size"),
    ]
    for name, args, needle in cases:
        args = tuple(ctypes.cast(a, ctypes.c_void_p) if isinstance(a, ctypes.Array) else a for a in args)
        rc = getattr(h, name)(*args)
        assert rc == -2, (name, rc)
        assert needle in _lib.last_error(), (name, needle, _lib.last_error())
    # a no-op size is accepted without a device, too (nothing to launch)
    assert h.b200ssl_dropout(p, p, None, 0, 0, 0.1, p, 0, None) == 0
    assert h.b200ssl_zero_bytes(p, 0, None) == 0 and h.b200ssl_copy_rows(p, 16, p, 16, 0, 16, None) == 0


def test_header_is_plain_c_and_the_integration_example_type_checks(tmp_path):
    """include/b200ssl.h must be consumable from C (the drop-in boundary is a C ABI: no C++ / torch types), and the C
    calls INTEGRATION.md shows must match the declared prototypes: the ```c block of INTEGRATION.md is wrapped in a
    function and compiled with gcc -std=c99 -pedantic -Werror (syntax + types only, nothing is linked or run)."""
    import re
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    with open(os.path.join(ROOT, "INTEGRATION.md"), encoding="utf-8") as fh:
        blocks = re.findall(r"```c\n(.*?)```", fh.read(), flags=re.S)
    assert blocks, "INTEGRATION.md lost its C example"
    body = "\n".join(line for line in blocks[0].splitlines() if not line.startswith("#include"))
    src = ('#include <stdio.h>\n#include "b200ssl.h"\n'
           "int example(const void* tiles_u8, const void* params, void* crops224, void* crops96, const void* x,\n"
           "            const void* W1, const void* W2, void* dact, void* h, const float* b1, const float* b2,\n"
           "            float* x1_f32, const float* x0_f32, const void* qkv, void* out, float* lse2, int rows, int B,\n"
           "            int N, int H, void* stream) {\n" + body + "\nreturn 0;\n}\n")
    path = tmp_path / "example.c"
    path.write_text(src)
    r = subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter",
                        "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # and a C translation unit that names every declared entry point (their addresses type-check against `int (*)()` etc.)
    with open(os.path.join(ROOT, "include", "b200ssl.h")) as fh:
        names = sorted(set(re.findall(r"\b(b200ssl_[a-z0-9_]+)\s*\(", fh.read())))
    from b200ssl import _lib
    assert set(names) == set(_lib.exported_symbols())
    tu = '#include "b200ssl.h"\nconst void* table[] = {\n' + "".join(f"  (const void*)&{n},\n" for n in names) + "};\n"
    path2 = tmp_path / "table.c"
    path2.write_text(tu)
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        str(path2)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
