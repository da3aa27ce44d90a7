"""Developer check (GPU): fused tcgen05 attention fwd/bwd vs a torch fp32 reference. Not a pytest file."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# B200SSL_LIB: another build of the library (A/B against a previous kernel version)
h = ctypes.CDLL(os.environ.get("B200SSL_LIB") or os.path.join(ROOT, "gipmed-project-self-supervised-vit_b200", "libb200ssl.so"))
h.b200ssl_last_error.restype = ctypes.c_char_p
P, I, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
h.b200ssl_attention_fwd.argtypes = [P, P, P, I, I, I, I, F, P]
h.b200ssl_attention_bwd.argtypes = [P, P, P, P, P, I, I, I, I, F, P]
LL = ctypes.c_longlong
h.b200ssl_attention_fwd_ws.argtypes = [P, P, P, I, I, I, I, F, P, LL, P]
h.b200ssl_attention_fwd_workspace_bytes.argtypes = [I, I, I]
h.b200ssl_attention_fwd_workspace_bytes.restype = LL


def ck(rc):
    if rc != 0:
        raise RuntimeError(h.b200ssl_last_error().decode())


def ref_attn(qkv, B, N, H, scale):
    q, k, v = qkv.float().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    a = (q @ k.transpose(-2, -1)) * scale
    p = a.softmax(-1)
    o = (p @ v).transpose(1, 2).reshape(B, N, H * 64)
    return o, torch.logsumexp(a, -1)


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def run(B, N, H, bench=False):
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + N)
    scale = 0.125
    qkv = (torch.randn(B, N, 3 * H * 64, device="cuda", generator=g) * 1.0).to(torch.bfloat16)
    out = torch.full((B, N, H * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    lse2 = torch.zeros(B, H, N, device="cuda", dtype=torch.float32)
    s = torch.cuda.current_stream().cuda_stream
    nws = h.b200ssl_attention_fwd_workspace_bytes(B, N, H)
    ws = torch.empty(max(nws, 16), dtype=torch.uint8, device="cuda")
    ck(h.b200ssl_attention_fwd_ws(qkv.data_ptr(), out.data_ptr(), lse2.data_ptr(), B, N, H, 64, scale, ws.data_ptr(), nws, s))
    torch.cuda.synchronize()
    qkv_ref = qkv.float().requires_grad_(True)
    o_ref, lse_ref = ref_attn(qkv_ref, B, N, H, scale)
    e_o = rel(out, o_ref)
    e_l = rel(lse2 * 0.6931471805599453, lse_ref)
    nan = int(torch.isnan(out.float()).sum())
    dout = torch.randn(B, N, H * 64, device="cuda", generator=g).to(torch.bfloat16)
    o_ref.backward(dout.float())
    dqkv = torch.full_like(qkv, float("nan"))
    ck(h.b200ssl_attention_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse2.data_ptr(), dqkv.data_ptr(),
                               B, N, H, 64, scale, s))
    torch.cuda.synchronize()
    gref = qkv_ref.grad.view(B, N, 3, H, 64)
    gg = dqkv.view(B, N, 3, H, 64)
    e_q, e_k, e_v = rel(gg[:, :, 0], gref[:, :, 0]), rel(gg[:, :, 1], gref[:, :, 1]), rel(gg[:, :, 2], gref[:, :, 2])
    nan_g = int(torch.isnan(dqkv.float()).sum())
    ok = max(e_o, e_l, e_q, e_k, e_v) < 2e-2 and nan == 0 and nan_g == 0
    print(f"[{'ok' if ok else 'FAIL'}] B={B} N={N} H={H}: out {e_o:.2e} lse {e_l:.2e} dq {e_q:.2e} dk {e_k:.2e} "
          f"dv {e_v:.2e} nan={nan},{nan_g}")
    if not ok:
        d = (out.float() - o_ref).abs().view(B, N, H, 64).amax(-1)
        print("  out err per (b0, n, h0):", d[0, :, 0].cpu().tolist()[:40])
        for nm, i in (("dq", 0), ("dk", 1), ("dv", 2)):
            dd = (gg[:, :, i].float() - gref[:, :, i]).abs().amax(-1)
            print(f"  {nm} err per n (b0,h0):", [round(x, 3) for x in dd[0, :, 0].cpu().tolist()[:40]])
    if bench:
        for name, fn in (("fwd", lambda: h.b200ssl_attention_fwd_ws(qkv.data_ptr(), out.data_ptr(), lse2.data_ptr(), B, N, H, 64, scale, ws.data_ptr(), nws, s)),
                         ("bwd", lambda: h.b200ssl_attention_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse2.data_ptr(), dqkv.data_ptr(), B, N, H, 64, scale, s))):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            fl = 4.0 * B * H * N * N * 64 * (1 if name == "fwd" else 2.5)
            print(f"  [bench] {name}: {ms*1e3:.1f} us, {fl/ms/1e9:.0f} TFLOP/s (algorithmic)")
    return ok


def main():
    ok = True
    if "--stream" in sys.argv:   # forward kernel choice for N > 128: -1 block decomposition, 0 default, 1 always streaming
        h.b200ssl_set_attn_stream.argtypes = [I]
        ck(h.b200ssl_set_attn_stream(int(sys.argv[sys.argv.index("--stream") + 1])))
    ok &= run(2, 128, 1)
    ok &= run(3, 64, 2)
    ok &= run(7, 37, 3)
    ok &= run(5, 100, 2)
    ok &= run(2, 197, 2)
    ok &= run(3, 256, 1)
    ok &= run(4, 145, 3)
    ok &= run(64, 197, 6)     # 384 items: persistent CTAs take 2-3 items each
    ok &= run(700, 37, 6)     # 1404 packed items: ~9.5 per CTA, two operand sets in flight
    ok &= run(160, 100, 3)    # N <= 128 without packing
    ok &= run(3, 257, 2)      # long sequences: native 256^2 tiles -> 257 tokens (2 x 2 blocks of 129 / 128)
    ok &= run(2, 785, 3)      # ViT-S/8 at 224^2: 785 tokens (4 x 4 blocks)
    ok &= run(40, 325, 6)     # 288^2 tiles, many items
    ok &= run(2, 129, 1)      # one key past the first block
    ok &= run(1, 1025, 2)     # ViT-S/8 at 256^2: 9 key blocks, the last of one key
    ok &= run(300, 200, 2)    # many items per CTA on the streaming path (mode 1)
    if "--bench" in sys.argv and ok:
        run(512, 197, 6, bench=True)
        run(2560, 37, 6, bench=True)
        run(256, 257, 6, bench=True)
        run(64, 785, 6, bench=True)
    print("ALL OK" if ok else "SOME FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
