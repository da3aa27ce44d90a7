"""Developer check (GPU): tcgen05 GEMM vs torch fp32 matmul over layouts / tiles / epilogues.

Run under gpurun:  python tests/gpu_checks/gemm_check.py [--bench]
Prints one line per case and a block-error map on failure so a descriptor/layout bug can be
diagnosed from the log alone. Not collected by pytest (no test_ prefix).
"""
import ctypes
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "gipmed-project-self-supervised-vit_b200", "libb200ssl.so")
h = ctypes.CDLL(LIB)
h.b200ssl_last_error.restype = ctypes.c_char_p
P, I, L = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
h.b200ssl_gemm.argtypes = [P, L, I, P, L, I, P, L, P, P, P, L, I, I, I, I, I, I, P]
h.b200ssl_gemm.restype = I


def gemm(A, a_mn, B, b_mn, D, M, N, K, epi=0, D2=None, bias=None, aux=None, split_k=1, bn=0):
    s = torch.cuda.current_stream().cuda_stream
    rc = h.b200ssl_gemm(A.data_ptr(), A.stride(0), a_mn, B.data_ptr(), B.stride(0), b_mn,
                        D.data_ptr(), D.stride(0), D2.data_ptr() if D2 is not None else None,
                        bias.data_ptr() if bias is not None else None,
                        aux.data_ptr() if aux is not None else None,
                        aux.stride(0) if aux is not None else 0, M, N, K, epi, split_k, bn, s)
    if rc != 0:
        raise RuntimeError(h.b200ssl_last_error().decode())


def errmap(got, ref, bm=32, bn=32):
    d = (got.float() - ref.float()).abs()
    M, N = d.shape
    Mb, Nb = (M + bm - 1) // bm, (N + bn - 1) // bn
    pad = torch.zeros(Mb * bm, Nb * bn, device=d.device)
    pad[:M, :N] = d
    blk = pad.view(Mb, bm, Nb, bn).amax(dim=(1, 3))
    return blk


def run_case(name, M, N, K, a_mn, b_mn, epi, bn, split_k=1, results=None):
    g = torch.Generator(device="cuda").manual_seed(hash(name) & 0xffff)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    A_in = A.t().contiguous() if a_mn else A
    B_in = B.t().contiguous() if b_mn else B
    ref = A.float() @ B.float().t()
    bias = torch.randn(N, device="cuda", generator=g) if epi in (0, 1, 2, 5, 7) else None
    if epi == 4 and a_mn:
        bias = torch.zeros(M, device="cuda")  # wgrad: receives db = column sums of A^T (i.e. A.sum over K)
    if epi == 6:
        bias = torch.zeros(N, device="cuda")  # transposed wgrad: receives db = B.sum over K
    aux = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16) if epi in (2, 3) else None
    if epi == 5:
        aux = torch.randn(M, N, device="cuda", generator=g) * 3.0   # fp32 residual stream
    D2 = None
    if epi == 4:
        D = torch.zeros(M, N, device="cuda", dtype=torch.float32)
    elif epi == 6:
        D = torch.zeros(N, M, device="cuda", dtype=torch.float32)   # stored transposed
    elif epi == 5:
        D = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    else:
        D = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    if epi in (0, 1, 2, 5, 7):
        ref = ref + bias
    if epi == 7:
        ref = torch.nn.functional.gelu(ref)
    ref2 = None
    if epi == 1:
        D2 = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
        ref2 = torch.nn.functional.gelu(ref)
        xr = ref.clone().requires_grad_(True)
        torch.nn.functional.gelu(xr).sum().backward()
        ref = xr.grad
    if epi in (2, 5):
        ref = ref + aux.float()
    if epi == 3:
        ref = ref * aux.float()
    try:
        gemm(A_in, a_mn, B_in, b_mn, D, M, N, K, epi, D2, bias, aux, split_k, bn)
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        print(f"[FAIL] {name}: exception {e}")
        if results is not None:
            results.append({"name": name, "ok": False, "err": str(e)})
        return False
    if epi == 6:
        D = D.t()
    scale = ref.abs().max().item() + 1e-6
    err = (D.float() - ref).abs().max().item() / scale
    nan = int(torch.isnan(D.float()).sum().item())
    ok = err < 1.5e-2 and nan == 0
    if epi == 6:
        dbref = B.float().sum(1)
        errb = (bias - dbref).abs().max().item() / (dbref.abs().max().item() + 1e-6)
        ok = ok and errb < 1e-3
        err = max(err, errb)
    if epi == 4 and a_mn:
        dbref = A.float().sum(1)
        errb = (bias - dbref).abs().max().item() / (dbref.abs().max().item() + 1e-6)
        ok = ok and errb < 1e-3
        err = max(err, errb)
    if ref2 is not None:
        err2 = (D2.float() - ref2).abs().max().item() / (ref2.abs().max().item() + 1e-6)
        ok = ok and err2 < 1.5e-2
        err = max(err, err2)
    print(f"[{'ok' if ok else 'FAIL'}] {name}: M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} epi={epi} bn={bn} "
          f"split={split_k} relerr={err:.3e} nan={nan}")
    if not ok:
        em = errmap(torch.nan_to_num(D.float(), nan=1e9), ref)
        torch.set_printoptions(linewidth=250, precision=1, sci_mode=False)
        print("  block (32x32) max-abs-error map, first 8x16 blocks:")
        print(em[:8, :16].cpu())
        print("  got[0,:8]", D[0, :8].float().cpu().tolist())
        print("  ref[0,:8]", ref[0, :8].cpu().tolist())
    if results is not None:
        results.append({"name": name, "ok": bool(ok), "relerr": err, "nan": nan})
    return ok


def bench_case(M, N, K, a_mn, b_mn, epi, bn, split_k=1, iters=20):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    A_in = A.t().contiguous() if a_mn else A
    B_in = B.t().contiguous() if b_mn else B
    bias = torch.randn(N, device="cuda") if epi not in (4, 6) else torch.zeros(max(M, N), device="cuda")
    aux = torch.randn(M, N, device="cuda").to(torch.bfloat16) if epi in (2, 3) else None
    if epi == 5:
        aux = torch.randn(M, N, device="cuda")
    D = torch.zeros(M, N, device="cuda", dtype=torch.float32 if epi in (4, 5, 6) else torch.bfloat16)
    if epi == 6:
        D = torch.zeros(N, M, device="cuda", dtype=torch.float32)
    D2 = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16) if epi == 1 else None
    for _ in range(3):
        gemm(A_in, a_mn, B_in, b_mn, D, M, N, K, epi, D2, bias, aux, split_k, bn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        gemm(A_in, a_mn, B_in, b_mn, D, M, N, K, epi, D2, bias, aux, split_k, bn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS reference for the same shape
    Wt = B
    for _ in range(3):
        torch.matmul(A, Wt.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(A, Wt.t())
    e1.record()
    torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / iters
    print(f"[bench] M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} epi={epi} bn={bn} split={split_k}: "
          f"{ms*1e3:.1f} us  {tf:.0f} TFLOP/s   (torch.matmul {ms_ref*1e3:.1f} us {2.0*M*N*K/ms_ref/1e9:.0f} TF)")
    return {"M": M, "N": N, "K": K, "epi": epi, "bn": bn, "us": ms * 1e3, "tflops": tf,
            "cublas_us": ms_ref * 1e3}


def main():
    assert torch.cuda.is_available()
    if "--cluster" in sys.argv:
        n = int(sys.argv[sys.argv.index("--cluster") + 1])
        h.b200ssl_set_gemm_cluster.argtypes = [I]
        assert h.b200ssl_set_gemm_cluster(n) == 0
        print("gemm cluster size", n)
    results = []
    allok = True
    # 1. smallest K-major case first: one tile, one k-block
    allok &= run_case("kk_1tile_k64", 128, 64, 64, 0, 0, 0, 64, results=results)
    allok &= run_case("kk_1tile_k128", 128, 128, 128, 0, 0, 0, 128, results=results)
    for bn in (64, 128, 192, 256):
        allok &= run_case(f"kk_bn{bn}", 384, 768, 384, 0, 0, 0, bn, results=results)
    allok &= run_case("kk_ragged_m", 1000, 384, 384, 0, 0, 0, 192, results=results)
    allok &= run_case("kk_many_tiles", 128 * 333 + 17, 1152, 384, 0, 0, 0, 192, results=results)
    # epilogues
    allok &= run_case("kk_gelu", 1000, 1536, 384, 0, 0, 1, 256, results=results)
    allok &= run_case("kk_res", 1000, 384, 1536, 0, 0, 2, 192, results=results)
    allok &= run_case("kk_dgelu", 1000, 1536, 384, 0, 0, 3, 256, results=results)
    allok &= run_case("kk_gelu_fwd_only", 128 * 70 + 9, 1536, 384, 0, 0, 7, 256, results=results)
    allok &= run_case("kk_res32_k384", 1000, 384, 384, 0, 0, 5, 192, results=results)
    allok &= run_case("kk_res32_k1536", 5000, 384, 1536, 0, 0, 5, 192, results=results)
    allok &= run_case("kk_res32_bn128", 700, 768, 384, 0, 0, 5, 128, results=results)
    allok &= run_case("kk_gelu_many", 128 * 300 + 5, 1536, 384, 0, 0, 1, 256, results=results)
    allok &= run_case("kk_bn64_many", 128 * 40 + 77, 256, 2048, 0, 0, 0, 64, results=results)
    # dgrad layout: B MN-major
    allok &= run_case("kmn_small", 128, 64, 64, 0, 1, 0, 64, results=results)
    allok &= run_case("kmn", 1000, 384, 1536, 0, 1, 0, 192, results=results)
    allok &= run_case("kmn_dgelu", 1000, 1536, 384, 0, 1, 3, 256, results=results)
    # wgrad layout: both MN-major, fp32 atomics, split-K
    allok &= run_case("mnmn_small", 128, 64, 64, 1, 1, 4, 64, results=results)
    allok &= run_case("mnmn_small_bf16", 128, 128, 128, 1, 1, 0, 128, results=results)
    allok &= run_case("mnmn_wgrad", 1536, 384, 5000, 1, 1, 4, 128, split_k=7, results=results)
    allok &= run_case("mnmn_wgrad_256", 384, 1536, 3152, 1, 1, 4, 256, split_k=5, results=results)
    # 256 x 384 CTA-pair tiles (one accumulator, MMAs of N = 256 + 128)
    allok &= run_case("kk_bn384", 1000, 768, 1536, 0, 0, 0, 384, results=results)
    allok &= run_case("kmn_bn384", 128 * 37 + 5, 384, 1152, 0, 1, 0, 384, results=results)
    allok &= run_case("kk_res32_bn384", 128 * 21 + 3, 384, 1536, 0, 0, 5, 384, results=results)
    allok &= run_case("mnmn_wgrad_bn384", 1536, 384, 5000, 1, 1, 4, 384, split_k=7, results=results)
    allok &= run_case("mnmn_wgrad_bn384_odd", 384, 384, 3000, 1, 1, 4, 384, split_k=0, results=results)
    allok &= run_case("mnmn_wgrad_T", 1536, 384, 5000, 1, 1, 6, 384, split_k=0, results=results)
    allok &= run_case("mnmn_wgrad_T_768", 1000, 768, 3152, 1, 1, 6, 384, split_k=3, results=results)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = {"all_ok": bool(allok), "cases": results}
    if "--bench" in sys.argv and allok:
        b = []
        rows = 100864
        b.append(bench_case(rows, 1152, 384, 0, 0, 0, 192))
        b.append(bench_case(rows, 1152, 384, 0, 0, 0, 128))
        b.append(bench_case(rows, 384, 384, 0, 0, 5, 192))
        b.append(bench_case(rows, 1536, 384, 0, 0, 1, 256))
        b.append(bench_case(rows, 1536, 384, 0, 0, 0, 256))
        b.append(bench_case(rows, 384, 1536, 0, 0, 5, 192))
        b.append(bench_case(rows, 384, 1536, 0, 1, 0, 192))
        b.append(bench_case(rows, 1536, 384, 0, 1, 3, 256))
        b.append(bench_case(1536, 384, rows, 1, 1, 4, 128, split_k=8))
        b.append(bench_case(1536, 384, rows, 1, 1, 4, 192, split_k=12))
        b.append(bench_case(8192, 8192, 8192, 0, 0, 0, 256))
        b.append(bench_case(3072, 65536, 256, 0, 0, 0, 256))
        out["bench"] = b
    with open(os.path.join(ROOT, "gpurun_out", "gemm_check.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("ALL OK" if allok else "SOME FAILED")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
