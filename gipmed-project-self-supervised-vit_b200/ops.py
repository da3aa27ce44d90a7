"""Thin host layer over the C-ABI kernels: raw-pointer wrappers and the autograd Functions the
drop-in modules are built from.

Conventions: activations are bf16 row-major 2-D ``[rows, features]`` (rows = B*N tokens); parameters
stay fp32 ``nn.Parameter`` (state-dict compatible with the reference) and are shadowed in bf16 by
``bf16_of`` for the tensor-core operands; gradients of parameters are produced in fp32.
Nothing here computes on the CPU or through PyTorch math kernels: every op is a launch of a kernel
from ``libb200ssl.so`` on the current CUDA stream (graph-capturable, no host syncs).
"""
from __future__ import annotations

import torch

from . import _lib

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RES, EPI_MUL_AUX, EPI_ATOMIC_F32, EPI_BIAS_RES_F32, EPI_ATOMIC_F32_T, EPI_BIAS_GELU_FWD = 0, 1, 2, 3, 4, 5, 6, 7

_BF16 = torch.bfloat16
_counters = {"launches": 0}


def launch_count() -> int:
    """Number of b200ssl kernel launches issued by this process (bench.py reports the delta)."""
    return _counters["launches"]


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _call(name, *args, launches=1):
    rc = getattr(_lib.lib(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (code {rc}): {_lib.last_error()}")
    _counters["launches"] += launches


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: b200ssl ops run on sm_100a CUDA tensors only (got {t.device}); "
                           "there is no CPU fallback")


# ------------------------------------------------------------------------------------------------
# bf16 shadows of fp32 parameters
# ------------------------------------------------------------------------------------------------
class _ShadowCache:
    """bf16 copies of fp32 parameters keyed by the parameter object; refreshed when the parameter's
    version counter or storage changes. Writers that update parameters through raw pointers (the fused
    optimizer / EMA kernels) refresh the shadow themselves and call ``mark_fresh``."""

    def __init__(self):
        self._d = {}

    def _entry(self, p):
        key = id(p)
        ent = self._d.get(key)
        if ent is None or ent["t"].shape != p.shape or ent["t"].device != p.device or ent["ref"]() is not p:
            import weakref
            ent = {"t": torch.empty(p.shape, dtype=_BF16, device=p.device), "ver": -1, "ptr": 0,
                   "ref": weakref.ref(p)}
            self._d[key] = ent
        return ent

    def get(self, p: torch.Tensor) -> torch.Tensor:
        if p.dtype == _BF16:
            return p.detach()
        ent = self._entry(p)
        if ent["ver"] != p._version or ent["ptr"] != p.data_ptr():
            src = p.detach()
            if src.dtype != torch.float32 or not src.is_contiguous():
                src = src.float().contiguous()
            cast_f32_to_bf16(src, ent["t"])
            ent["ver"], ent["ptr"] = p._version, p.data_ptr()
        return ent["t"]

    def shadow_for(self, p: torch.Tensor) -> torch.Tensor:
        """Storage the raw-pointer writers fill; call mark_fresh(p) after the write is enqueued."""
        return self._entry(p)["t"]

    def mark_fresh(self, p: torch.Tensor):
        ent = self._entry(p)
        ent["ver"], ent["ptr"] = p._version, p.data_ptr()

    def invalidate(self, p: torch.Tensor):
        ent = self._d.get(id(p))
        if ent is not None:
            ent["ver"] = -1


shadows = _ShadowCache()


def bf16_of(p: torch.Tensor) -> torch.Tensor:
    return shadows.get(p)


def to_bf16_2d(x: torch.Tensor) -> torch.Tensor:
    x2 = x.reshape(-1, x.shape[-1])
    if x2.dtype != _BF16:
        x2 = x2.to(_BF16)
    return x2.contiguous()


def to_stream_2d(x: torch.Tensor) -> torch.Tensor:
    """The residual stream is fp32 [rows, D] (as under torch autocast); only MMA operands are bf16."""
    x2 = x.reshape(-1, x.shape[-1])
    if x2.dtype != torch.float32:
        x2 = x2.float()
    return x2.contiguous()


# ------------------------------------------------------------------------------------------------
# gradient sinks: accumulate parameter gradients straight into a persistent fp32 buffer
# ------------------------------------------------------------------------------------------------
# The wgrad GEMM (fp32 atomics), its folded bias gradient and the LayerNorm dgamma/dbeta reductions all
# ACCUMULATE into their output. When a parameter is registered here (GradBucketDataParallel does it for
# the views into its flat buckets), backward adds into ``param.grad`` directly and hands autograd ``None``:
# no per-step zero-filled temporaries and no autograd accumulation kernels (~470 launches per step).
_grad_sinks = {}


def register_grad_sink(param, callback=None):
    import weakref
    _grad_sinks[id(param)] = (weakref.ref(param), callback)


def unregister_grad_sink(param):
    _grad_sinks.pop(id(param), None)


def _sink(param):
    """The buffer to accumulate ``param``'s gradient into, or None (then the gradient goes through autograd)."""
    if param is None:
        return None
    ent = _grad_sinks.get(id(param))
    if ent is None or ent[0]() is not param:
        return None
    g = param.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != param.shape:
        return None
    return g


def _sunk(param):
    """Tell the sink's owner that one more contribution to ``param.grad`` has been enqueued."""
    cb = _grad_sinks[id(param)][1]
    if cb is not None:
        cb(param)


# ------------------------------------------------------------------------------------------------
# raw wrappers
# ------------------------------------------------------------------------------------------------
def cast_f32_to_bf16(src: torch.Tensor, dst: torch.Tensor):
    _call("b200ssl_cast_f32_to_bf16", src.data_ptr(), dst.data_ptr(), src.numel(), _stream())


def cast_bf16_to_f32(src: torch.Tensor, dst: torch.Tensor):
    _call("b200ssl_cast_bf16_to_f32", src.data_ptr(), dst.data_ptr(), src.numel(), _stream())


def zeros(shape, dtype, device):
    """torch.empty + one memset node on the current stream (instead of a PyTorch fill kernel)."""
    t = torch.empty(shape, dtype=dtype, device=device)
    _call("b200ssl_zero_bytes", t.data_ptr(), t.numel() * t.element_size(), _stream(), launches=0)
    return t


def clone(t):
    out = torch.empty_like(t, memory_format=torch.contiguous_format)
    _call("b200ssl_copy_bytes", out.data_ptr(), t.contiguous().data_ptr(), t.numel() * t.element_size(), _stream(),
          launches=0)
    return out


def copy_rows(src, src_stride_bytes, dst, dst_stride_bytes, rows, row_bytes):
    """dst[r * dst_stride ...] = src[r * src_stride ...] for r < rows (src / dst: tensors giving the base pointers)."""
    _call("b200ssl_copy_rows", src.data_ptr(), src_stride_bytes, dst.data_ptr(), dst_stride_bytes, rows, row_bytes,
          _stream())


def add_f32(dst, src):
    _call("b200ssl_add_f32", dst.data_ptr(), src.data_ptr(), dst.numel(), _stream())


def pos_table(pos_embed32, mat):
    """[1 + Mo, D] fp32 position table for one resolution: the stored table itself (mat None) or its bicubic
    resize applied as the cached linear map ``mat`` [Mo, Ki] (VT.pyc@L213-233)."""
    if mat is None:
        return pos_embed32
    Mo, Ki = mat.shape
    D = pos_embed32.shape[1]
    out = torch.empty(Mo + 1, D, dtype=torch.float32, device=pos_embed32.device)
    _call("b200ssl_pos_interp", mat.data_ptr(), pos_embed32.data_ptr(), out.data_ptr(), Mo, Ki, D, 0, _stream())
    return out


def pos_table_bwd(dtable, mat, dpos_embed):
    """dpos_embed (fp32 [1 + Ki, D]) += resize^T(dtable)."""
    if mat is None:
        add_f32(dpos_embed, dtable)
    else:
        Mo, Ki = mat.shape
        _call("b200ssl_pos_interp", mat.data_ptr(), dtable.data_ptr(), dpos_embed.data_ptr(), Mo, Ki,
              dtable.shape[1], 1, _stream())


def gemm(A, B, D, M, N, K, *, a_mn=False, b_mn=False, epi=EPI_BIAS, D2=None, bias=None, aux=None, split_k=1,
         block_n=0):
    _call("b200ssl_gemm", A.data_ptr(), A.stride(0), int(a_mn), B.data_ptr(), B.stride(0), int(b_mn),
          D.data_ptr(), D.stride(0), _ptr(D2), _ptr(bias), _ptr(aux), aux.stride(0) if aux is not None else 0,
          M, N, K, epi, split_k, block_n, _stream())


def scale_rows(x, rs):
    """y[r, :] = x[r, :] * rs[r] (bf16 rows, fp32 per-row scale)."""
    y = torch.empty_like(x)
    _call("b200ssl_scale_rows", x.data_ptr(), rs.data_ptr(), y.data_ptr(), x.shape[0], x.shape[1], _stream())
    return y


def dropout(x, drop, site, second=None, out=None):
    """nn.Dropout on the encoder path: ``out = x * mask / (1 - p)`` (bf16 or fp32, contiguous; ``out=x`` for in place,
    default a new tensor). ``drop`` = (p, seed): seed an int64 CUDA tensor of one element; the mask is a pure function
    of (seed, site, flat element index), so backward calls this again instead of saving it. ``second`` (bf16, optional)
    receives the same mask in place."""
    p, seed = drop
    out = torch.empty_like(x) if out is None else out
    if not x.is_contiguous() or not out.is_contiguous():
        raise RuntimeError("dropout: contiguous tensors only")
    if x.dtype not in (torch.float32, _BF16) or out.dtype != x.dtype or out.numel() != x.numel():
        raise RuntimeError(f"dropout: fp32 or bf16 in, the same out (got {x.dtype} -> {out.dtype})")
    if second is not None and (second.dtype != _BF16 or second.numel() != x.numel() or not second.is_contiguous()):
        raise RuntimeError("dropout: the second tensor must be contiguous bf16 of the same size")
    _call("b200ssl_dropout", x.data_ptr(), out.data_ptr(), _ptr(second), x.numel(), int(x.dtype == torch.float32),
          float(p), seed.data_ptr(), int(site), _stream())
    return out


def dropout_residual(branch, residual, rowscale, drop, site):
    """``residual + rowscale[row] * dropout(branch)`` on the fp32 stream (x + drop_path(drop(branch)), VT.pyc@L150-151)."""
    p, seed = drop
    if branch.dtype != _BF16 or residual.dtype != torch.float32 or branch.shape != residual.shape or \
            not (branch.is_contiguous() and residual.is_contiguous()):
        raise RuntimeError("dropout_residual: contiguous bf16 branch and fp32 stream of the same shape")
    y = torch.empty_like(residual)
    _call("b200ssl_dropout_residual", branch.data_ptr(), residual.data_ptr(), _ptr(rowscale), y.data_ptr(),
          branch.shape[0], branch.shape[1], float(p), seed.data_ptr(), int(site), _stream())
    return y


def linear_fwd(x, w16, bias=None, residual=None, gelu=False, rowscale=None):
    """y = x @ w16^T (+bias) (+residual); with gelu=True returns (gelu'(pre), gelu(pre)) — the derivative is
    what backward needs, so it is saved instead of the pre-activation (one erf evaluation serves both).
    An fp32 ``residual`` selects the fp32-stream epilogue (fp32 output); bf16 residual -> bf16 output."""
    M, K = x.shape
    N = w16.shape[0]
    if residual is not None and residual.dtype == torch.float32:
        # rowscale (fp32 [M], optional): y = residual + rowscale[row] * (x w^T + bias) -- stochastic depth
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        gemm(x, w16, y, M, N, K, epi=EPI_BIAS_RES_F32, bias=bias, aux=residual, D2=rowscale)
        return y
    if rowscale is not None:
        raise RuntimeError("linear_fwd: rowscale needs the fp32 residual epilogue")
    y = torch.empty(M, N, dtype=_BF16, device=x.device)
    if gelu == "fwd_only":
        # no-grad forward (teacher): gelu(pre) only, nothing saved for backward
        gemm(x, w16, y, M, N, K, epi=EPI_BIAS_GELU_FWD, bias=bias)
        return None, y
    if gelu:
        h = torch.empty(M, N, dtype=_BF16, device=x.device)
        gemm(x, w16, y, M, N, K, epi=EPI_BIAS_GELU, D2=h, bias=bias)
        return y, h
    if residual is not None:
        gemm(x, w16, y, M, N, K, epi=EPI_BIAS_RES, bias=bias, aux=residual)
    else:
        gemm(x, w16, y, M, N, K, epi=EPI_BIAS, bias=bias)
    return y


def linear_dgrad(dy, w16, dgelu_of=None):
    """dx = dy @ w16 ; optionally multiplied in the epilogue by ``dgelu_of`` = the saved gelu'(pre)."""
    M, N = dy.shape
    K = w16.shape[1]
    if dgelu_of is None and N >= 8192 and (M // 128) * ((K + 255) // 256) < 64:
        # a very long reduction with too few output tiles to fill the SMs (the DINO head's last layer: 65,536
        # prototypes, 3072 x 256 output): split K across the machine into an fp32 buffer, then round once
        acc = zeros((M, K), torch.float32, dy.device)
        gemm(dy, w16, acc, M, K, N, b_mn=True, epi=EPI_ATOMIC_F32, split_k=0)
        dx = torch.empty(M, K, dtype=_BF16, device=dy.device)
        cast_f32_to_bf16(acc, dx)
        return dx
    dx = torch.empty(M, K, dtype=_BF16, device=dy.device)
    if dgelu_of is not None:
        gemm(dy, w16, dx, M, K, N, b_mn=True, epi=EPI_MUL_AUX, aux=dgelu_of)
    else:
        gemm(dy, w16, dx, M, K, N, b_mn=True, epi=EPI_BIAS)
    return dx


def linear_wgrad(dy, x, need_bias=True, weight=None, bias=None):
    """dW[N,K] = dy^T @ x (fp32, split-K atomics), db[N] = column sums of dy (folded by idle warps of the
    wgrad kernel from its smem stages). With ``weight`` / ``bias`` parameters that have a registered
    gradient sink the result is accumulated into ``.grad`` in place and ``None`` is returned for it."""
    M, N = dy.shape
    K = x.shape[1]
    w_sink = _sink(weight)
    b_sink = _sink(bias) if need_bias else None

    def run(dw, db):
        if N % 384 == 0 and K > N:
            # more inputs than outputs (fc2): compute dW^T = x^T dy with the long dimension as GEMM-M so the
            # 256 x 384 CTA-pair tiles are full, and store it transposed; db folds from the B operand (dy)
            gemm(x, dy, dw, K, N, M, a_mn=True, b_mn=True, epi=EPI_ATOMIC_F32_T, split_k=0, bias=db)
        else:
            gemm(dy, x, dw, N, K, M, a_mn=True, b_mn=True, epi=EPI_ATOMIC_F32, split_k=0, bias=db)

    if w_sink is not None and (not need_bias or b_sink is not None):
        run(w_sink.view(N, K), b_sink)
        _sunk(weight)
        if b_sink is not None:
            _sunk(bias)
        return None, None
    buf = zeros((N * K + (N if need_bias else 0),), torch.float32, dy.device)
    dw = buf[:N * K].view(N, K)
    db = buf[N * K:] if need_bias else None
    run(dw, db)
    return dw, db


def layernorm_fwd(x, w, b, eps):
    """x: bf16 or fp32 (residual stream) [rows, D] -> bf16 y, fp32 mean / rstd."""
    rows, D = x.shape
    y = torch.empty(rows, D, dtype=_BF16, device=x.device)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    _call("b200ssl_layernorm_fwd", x.data_ptr(), int(x.dtype == torch.float32), w.data_ptr(), b.data_ptr(),
          y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, D, float(eps), _stream())
    return y, mean, rstd


# The LayerNorm-fused GEMM is correct (tests/gpu_checks/ln_gemm_check.py) but measured SLOWER than the separate
# LayerNorm kernel + B-stationary GEMM (qkv: 175 vs 135 us, fc1: 237 vs 210 us at 100,864 rows): with one 96 KB A
# panel per CTA the in-kernel normalisation (17 k cycles per 128 rows) cannot overlap the MMAs. Off by default.
_LN_GEMM = {"on": False}


def ln_gemm_ok(x, w16):
    """The LayerNorm-fused GEMM covers the fp32 residual stream of a 384-wide encoder (ViT-S) in CTA-pair tiles."""
    N = w16.shape[0]
    return (_LN_GEMM["on"] and x.dtype == torch.float32 and x.shape[1] == 384 and x.shape[0] > 128
            and x.is_contiguous() and (N % 192 == 0 or N % 256 == 0))


def ln_linear_fwd(x, ln_w, ln_b, eps, w16, bias=None, gelu=False, keep=True):
    """y = epilogue(LN(x) @ w16^T + bias) in ONE kernel (b200ssl_ln_gemm): the normalised rows never make a round
    trip through HBM on their way into the GEMM. Returns (y | (gelu', gelu) | (None, gelu), ln, mean, rstd); with
    keep=False (no-grad forward) ln / mean / rstd are not produced."""
    M, K = x.shape
    N = w16.shape[0]
    dev = x.device
    ln = torch.empty(M, K, dtype=_BF16, device=dev) if keep else None
    mean = torch.empty(M, dtype=torch.float32, device=dev) if keep else None
    rstd = torch.empty(M, dtype=torch.float32, device=dev) if keep else None
    y = torch.empty(M, N, dtype=_BF16, device=dev)
    h = torch.empty(M, N, dtype=_BF16, device=dev) if gelu is True else None
    epi = EPI_BIAS if not gelu else (EPI_BIAS_GELU if gelu is True else EPI_BIAS_GELU_FWD)
    _call("b200ssl_ln_gemm", x.data_ptr(), x.stride(0), ln_w.data_ptr(), ln_b.data_ptr(), float(eps), _ptr(ln),
          _ptr(mean), _ptr(rstd), w16.data_ptr(), w16.stride(0), y.data_ptr(), y.stride(0), _ptr(h), _ptr(bias),
          M, N, K, epi, _stream())
    if gelu is True:
        return (y, h), ln, mean, rstd
    if gelu:
        return (None, y), ln, mean, rstd
    return y, ln, mean, rstd


# LayerNorm TAIL of the residual GEMMs (b200ssl_gemm_res_ln): attn.proj / mlp.fc2 write the fp32 stream AND the bf16
# LayerNorm of it (Block.norm2 / the next Block.norm1) from one kernel; the standalone LayerNorm-forward launches of a
# 384-wide encoder disappear. Correct (tests/test_gpu_kernels.py::test_residual_gemm_with_layernorm_tail, the model test
# with the switch on) but measured SLOWER on the config-2 step, in both placements of the tail: inside the epilogue warps
# (wait for the stores, re-read the rows from L2) 53.0 vs 52.7 ms per step; on the two idle warps one m unit behind the
# epilogue 60.5 vs 54.3 ms (64 rows per warp with twelve 16-byte L2 loads in flight per lane is latency bound and
# back-pressures the epilogue). The standalone LayerNorm kernel streams at 5.5 TB/s; the 300 MB of HBM reads the tail
# saves per call are worth less than what it costs the warps these HBM-bound GEMMs are limited by. Off by default
# (bench.py --ln-tail).
_LN_TAIL = {"on": False}


def ln_tail_ok(rows, n_out):
    return _LN_TAIL["on"] and n_out == 384 and rows > 128


def gemm_res_ln(x, w16, y, bias, residual, rowscale, ln_w, ln_b, eps, ln, mean, rstd):
    M, K = x.shape
    _call("b200ssl_gemm_res_ln", x.data_ptr(), x.stride(0), w16.data_ptr(), w16.stride(0), y.data_ptr(), y.stride(0),
          _ptr(rowscale), _ptr(bias), residual.data_ptr(), residual.stride(0), M, w16.shape[0], K, ln_w.data_ptr(),
          ln_b.data_ptr(), float(eps), ln.data_ptr(), _ptr(mean), _ptr(rstd), _stream())


def linear_res_ln_fwd(x, w16, bias, residual, rowscale, ln_w, ln_b, eps, keep=True):
    """y = residual + rowscale * (x @ w16^T + bias) on the fp32 stream, and LayerNorm(y) from the same kernel.
    -> (y fp32, (ln bf16, mean, rstd)); mean / rstd are None with keep=False (no-grad forward)."""
    M, N = x.shape[0], w16.shape[0]
    dev = x.device
    y = torch.empty(M, N, dtype=torch.float32, device=dev)
    ln = torch.empty(M, N, dtype=_BF16, device=dev)
    mean = torch.empty(M, dtype=torch.float32, device=dev) if keep else None
    rstd = torch.empty(M, dtype=torch.float32, device=dev) if keep else None
    gemm_res_ln(x, w16, y, bias, residual, rowscale, ln_w, ln_b, eps, ln, mean, rstd)
    return y, (ln, mean, rstd)


def layernorm_bwd(x, dy, w, mean, rstd, dres=None, weight=None, bias=None):
    """Gradients travel in bf16: dy, dres, dx are bf16 regardless of the stream dtype of x. dgamma / dbeta
    are accumulated into the parameters' gradient sinks when registered (then returned as None)."""
    rows, D = x.shape
    dx = torch.empty(rows, D, dtype=_BF16, device=x.device)
    w_sink, b_sink = _sink(weight), _sink(bias)
    if w_sink is not None and b_sink is not None:
        dw, db, ret = w_sink, b_sink, (None, None)
    else:
        dwb = zeros((2, D), torch.float32, x.device)
        dw, db = dwb[0], dwb[1]
        ret = (dw, db)
    _call("b200ssl_layernorm_bwd", x.data_ptr(), int(x.dtype == torch.float32), dy.data_ptr(), w.data_ptr(),
          mean.data_ptr(), rstd.data_ptr(), _ptr(dres), dx.data_ptr(), dw.data_ptr(), db.data_ptr(), rows, D,
          _stream())
    if ret[0] is None:
        _sunk(weight)
        _sunk(bias)
    return dx, ret[0], ret[1]


def segments(B, N):
    """[(B_i, N_i, first_row_i)] of a packed token stream. ``B`` / ``N`` are ints (one crop group) or equally long
    tuples: several groups of different sequence length stored back to back (global crops, then local crops), so
    that every row-wise kernel (LayerNorm, the GEMMs) runs ONCE over all rows and only attention is per group."""
    if not isinstance(B, (tuple, list)):
        return [(int(B), int(N), 0)]
    out, r0 = [], 0
    for b, n in zip(B, N):
        out.append((int(b), int(n), r0))
        r0 += int(b) * int(n)
    return out


def _attention_fwd_one(qkv, out, B, N, H, scale):
    lse2 = torch.empty(B, H, N, dtype=torch.float32, device=qkv.device)
    if N > 256:
        # long sequences (native 256^2 tiles: 257 tokens; ViT-S/8: 785): the streaming kernel (one launch, no scratch);
        # only the developer A/B mode b200ssl_set_attn_stream(-1) takes the block decomposition and needs scratch
        nbytes = int(_lib.lib().b200ssl_attention_fwd_workspace_bytes(B, N, H))
        if nbytes > 0:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=qkv.device)
            _call("b200ssl_attention_fwd_ws", qkv.data_ptr(), out.data_ptr(), lse2.data_ptr(), B, N, H, 64,
                  float(scale), ws.data_ptr(), nbytes, _stream(), launches=((N + 255) // 256) ** 2 + 1)
            return lse2
    _call("b200ssl_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse2.data_ptr(), B, N, H, 64, float(scale),
          _stream())
    return lse2


def attention_fwd(qkv, B, N, H, scale):
    """-> (out [rows, H*64] bf16, lse2): lse2 is one [B, H, N] tensor, or a tuple of them for several groups."""
    out = torch.empty(qkv.shape[0], H * 64, dtype=_BF16, device=qkv.device)
    segs = segments(B, N)
    lses = tuple(_attention_fwd_one(qkv[r0:r0 + b * n], out[r0:r0 + b * n], b, n, H, scale) for b, n, r0 in segs)
    return out, (lses if isinstance(B, (tuple, list)) else lses[0])


def attention_bwd(qkv, out, dout, lse2, B, N, H, scale):
    dqkv = torch.empty_like(qkv)
    lses = lse2 if isinstance(B, (tuple, list)) else (lse2,)
    for (b, n, r0), lse in zip(segments(B, N), lses):
        r1 = r0 + b * n
        _call("b200ssl_attention_bwd", qkv[r0:r1].data_ptr(), out[r0:r1].data_ptr(), dout[r0:r1].data_ptr(),
              lse.data_ptr(), dqkv[r0:r1].data_ptr(), b, n, H, 64, float(scale), _stream(),
              launches=4 if 256 < n <= 512 else 1)   # 257..512 tokens: four block-pair launches; else one
    return dqkv


class GradRelay:
    """Carries the bf16 token-stream gradient from EncoderFn.backward straight to (Multi)TokensFn.backward. The
    stream is fp32 in the forward, so autograd would demand an fp32 gradient for it: 450 MB of bf16 -> fp32 -> bf16
    round trip per step at config 2 (0.26 ms). With a relay the autograd edge carries a stride-0 zero placeholder of
    the right shape and dtype, and the real gradient travels here. Only for token tensors private to one forward."""
    __slots__ = ("grad",)
    _zeros = {}

    def __init__(self):
        self.grad = None

    def put(self, dx, shape):
        self.grad = dx
        key = (dx.device.type, dx.device.index)
        if key not in GradRelay._zeros:
            GradRelay._zeros[key] = torch.zeros((), dtype=torch.float32, device=dx.device)
        return GradRelay._zeros[key].expand(shape)

    def take(self, placeholder):
        if self.grad is None:          # the producer did not use the relay (e.g. a plain autograd consumer)
            return _g16(placeholder)
        dx, self.grad = self.grad, None
        return dx


def _g16(g):
    """Incoming gradients are carried in bf16 (autograd may hand us fp32 when the forward output was fp32)."""
    return g.to(_BF16).contiguous() if g.dtype != _BF16 else g.contiguous()


def _f32(p):
    t = p.detach()
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.float().contiguous()
    return t


# ------------------------------------------------------------------------------------------------
# autograd Functions
# ------------------------------------------------------------------------------------------------
class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        y, mean, rstd = layernorm_fwd(x, _f32(weight), _f32(bias), eps)
        ctx.save_for_backward(x, weight, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean, rstd = ctx.saved_tensors
        dx, dw, db = layernorm_bwd(x, _g16(dy), _f32(weight), mean, rstd)
        return dx.to(x.dtype), dw.to(weight.dtype), db.to(weight.dtype), None


class DropoutFn(torch.autograd.Function):
    """nn.Dropout as its own autograd node (standalone prepare_tokens / get_intermediate_layers: pos_drop)."""

    @staticmethod
    def forward(ctx, x, p, seed, site):
        ctx.drop, ctx.site = (p, seed), site
        return dropout(x.contiguous(), ctx.drop, site)

    @staticmethod
    def backward(ctx, dy):
        return dropout(dy.contiguous(), ctx.drop, ctx.site), None, None, None


class LinearFn(torch.autograd.Function):
    """y = x W^T + b (+ residual)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual):
        w16 = bf16_of(weight)
        y = linear_fwd(x, w16, _f32(bias) if bias is not None else None, residual)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        dx = linear_dgrad(dy, bf16_of(weight)) if ctx.needs_input_grad[0] else None
        dw, db = (None, None)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw, db = linear_wgrad(dy, x, ctx.has_bias)
        return dx, dw, db, (dy if ctx.has_res else None)


class MlpChainFn(torch.autograd.Function):
    """y = L_n(gelu(... gelu(L_1(x)))) (+ residual): GELU fused in each GEMM epilogue forward, its
    derivative fused in the dgrad epilogues backward. Args: x, residual, w1, b1, ..., wn, bn."""

    @staticmethod
    def forward(ctx, x, residual, *wb):
        n = len(wb) // 2
        saved = [x]
        h = x
        keep = any(ctx.needs_input_grad)  # False under no_grad (teacher head): gelu' is not produced
        for i in range(n):
            w, b = wb[2 * i], wb[2 * i + 1]
            b32 = _f32(b) if b is not None else None
            if i < n - 1:
                pre, h = linear_fwd(h, bf16_of(w), b32, gelu=True if keep else "fwd_only")
                saved += [pre, h] if keep else []
            else:
                h = linear_fwd(h, bf16_of(w), b32, residual=residual)
        ctx.n = n
        ctx.has_res = residual is not None
        ctx.has_bias = [wb[2 * i + 1] is not None for i in range(n)]
        ctx.biases = [wb[2 * i + 1] for i in range(n)]
        ctx.save_for_backward(*saved, *[wb[2 * i] for i in range(n)])
        return h

    @staticmethod
    def backward(ctx, dy):
        n = ctx.n
        saved = ctx.saved_tensors
        acts, weights = saved[: 1 + 2 * (n - 1)], saved[1 + 2 * (n - 1):]
        dy = dy.contiguous()
        dres = dy if ctx.has_res else None
        grads = [None] * (2 * n)
        d = dy
        for i in range(n - 1, -1, -1):
            inp = acts[0] if i == 0 else acts[2 * i]            # input of layer i (x or gelu output)
            dw, db = linear_wgrad(d, inp, ctx.has_bias[i], weights[i], ctx.biases[i])
            grads[2 * i], grads[2 * i + 1] = dw, db
            if i > 0:
                d = linear_dgrad(d, bf16_of(weights[i]), dgelu_of=acts[2 * i - 1])
            elif ctx.needs_input_grad[0]:
                d = linear_dgrad(d, bf16_of(weights[0]))
            else:
                d = None
        return (d, dres, *grads)


class MlpDropFn(torch.autograd.Function):
    """The bare Mlp module with dropout (VT.pyc@L98-104: fc1 -> GELU -> drop -> fc2 -> drop) as one autograd node;
    mask sites 0 (behind the activation) and 1 (behind fc2) of ``seed``. The Block / VisionTransformer paths do not
    come through here (mlp_half_fwd fuses the second dropout with the residual add)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, p, seed):
        drop = (p, seed)
        keep = any(ctx.needs_input_grad)
        pre, h = linear_fwd(x, bf16_of(w1), _f32(b1) if b1 is not None else None, gelu=True if keep else "fwd_only")
        dropout(h, drop, 0, second=pre, out=h)
        y = linear_fwd(h, bf16_of(w2), _f32(b2) if b2 is not None else None)
        dropout(y, drop, 1, out=y)
        ctx.drop, ctx.biases = drop, (b1, b2)
        if keep:
            ctx.save_for_backward(x, pre, h, w1, w2)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, pre, h, w1, w2 = ctx.saved_tensors
        b1, b2 = ctx.biases
        dyb = dropout(_g16(dy).contiguous(), ctx.drop, 1)
        d_w2, d_b2 = linear_wgrad(dyb, h, b2 is not None, w2, b2)
        d_pre = linear_dgrad(dyb, bf16_of(w2), dgelu_of=pre)
        d_w1, d_b1 = linear_wgrad(d_pre, x, b1 is not None, w1, b1)
        dx = linear_dgrad(d_pre, bf16_of(w1)) if ctx.needs_input_grad[0] else None
        return dx, d_w1, d_b1, d_w2, d_b2, None, None


class AttentionCoreFn(torch.autograd.Function):
    """softmax(q k^T * scale) v on the packed QKV-GEMM output [B*N, 3*H*64] -> [B*N, H*64]."""

    @staticmethod
    def forward(ctx, qkv, B, N, H, scale):
        out, lse2 = attention_fwd(qkv, B, N, H, scale)
        ctx.save_for_backward(qkv, out, lse2)
        ctx.dims = (B, N, H, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse2 = ctx.saved_tensors
        B, N, H, scale = ctx.dims
        return attention_bwd(qkv, out, dout.contiguous(), lse2, B, N, H, scale), None, None, None, None


# ---- residual half-blocks: plain functions shared by the per-block and whole-encoder autograd nodes ----
def attn_half_fwd(x, ln_w, ln_b, qkv_w, qkv_b, proj_w, proj_b, eps, B, N, H, scale, keep=True, rs=None, pre=None,
                  nxt=None, drop=None, site=0):
    """x + proj(attention(qkv(LN(x)))) (VT.pyc@L147,150). Returns (y, saved-for-backward, next LayerNorm state).
    ``pre`` = (ln, mean, rstd) of THIS half's LayerNorm when the previous residual GEMM already produced it;
    ``nxt`` = (ln_w, ln_b, eps) of the LayerNorm that follows (Block.norm2): produced by the proj GEMM's tail when the
    shape allows, else the third return value is None. ``drop`` = (p, seed) turns on Attention.proj_drop
    (VT.pyc@L130) with the mask of ``site``: the branch leaves the proj GEMM in bf16 and a second kernel masks it and
    adds it to the stream."""
    wq16 = bf16_of(qkv_w)
    qb32 = _f32(qkv_b) if qkv_b is not None else None
    if pre is not None:
        ln, mean, rstd = pre
        qkv = linear_fwd(ln, wq16, qb32)
    elif ln_gemm_ok(x, wq16):
        qkv, ln, mean, rstd = ln_linear_fwd(x, _f32(ln_w), _f32(ln_b), eps, wq16, qb32, keep=keep)
    else:
        ln, mean, rstd = layernorm_fwd(x, _f32(ln_w), _f32(ln_b), eps)
        qkv = linear_fwd(ln, wq16, qb32)
    att, lse2 = attention_fwd(qkv, B, N, H, scale)
    pb32 = _f32(proj_b) if proj_b is not None else None
    nxt_state = None
    if drop is not None:
        y = dropout_residual(linear_fwd(att, bf16_of(proj_w), pb32), x, rs, drop, site)
    elif nxt is not None and ln_tail_ok(att.shape[0], proj_w.shape[0]):
        y, nxt_state = linear_res_ln_fwd(att, bf16_of(proj_w), pb32, x, rs, _f32(nxt[0]), _f32(nxt[1]), nxt[2], keep=keep)
    else:
        y = linear_fwd(att, bf16_of(proj_w), pb32, residual=x, rowscale=rs)
    return y, (x, mean, rstd, ln, qkv, att, lse2, rs), nxt_state


def attn_half_bwd(dy, saved, ln_w, qkv_w, proj_w, has_qb, has_pb, B, N, H, scale, ln_b=None, qkv_b=None,
                  proj_b=None, drop=None, site=0):
    """dy: bf16 gradient of the half-block output. Returns (dx bf16, d_ln_w, d_ln_b, d_qkv_w, d_qkv_b,
    d_proj_w, d_proj_b); the residual gradient is added inside the LayerNorm-backward kernel."""
    x, mean, rstd, ln, qkv, att, lse2, rs = saved
    dyb = dy if rs is None else scale_rows(dy, rs)   # stochastic depth: the branch sees the scaled gradient
    if drop is not None:
        dyb = dropout(dyb, drop, site)   # proj_drop: forward's mask, regenerated; dy itself stays whole for the residual
    d_att = linear_dgrad(dyb, bf16_of(proj_w))
    d_pw, d_pb = linear_wgrad(dyb, att, has_pb, proj_w, proj_b)
    d_qkv = attention_bwd(qkv, att, d_att, lse2, B, N, H, scale)
    d_ln = linear_dgrad(d_qkv, bf16_of(qkv_w))
    d_qw, d_qb = linear_wgrad(d_qkv, ln, has_qb, qkv_w, qkv_b)
    dx, d_lw, d_lb = layernorm_bwd(x, d_ln, _f32(ln_w), mean, rstd, dres=dy, weight=ln_w, bias=ln_b)
    return dx, d_lw, d_lb, d_qw, d_qb, d_pw, d_pb


def mlp_half_fwd(x, ln_w, ln_b, w1, b1, w2, b2, eps, keep=True, rs=None, need_out=True, pre=None, nxt=None, drop=None,
                 site=0):
    """x + fc2(gelu(fc1(LN(x)))) (VT.pyc@L151). keep=False (no-grad forward): gelu' is not produced.
    need_out=False (activation recompute in backward): the fc2 GEMM is skipped, only the saved tensors are rebuilt.
    ``pre`` / ``nxt`` as in attn_half_fwd (``nxt`` = the next Block's norm1). Returns (y, saved, next LayerNorm state).
    ``drop`` = (p, seed): Mlp.drop behind the activation (mask ``site``: gelu(pre) and the saved gelu'(pre) are masked in
    place together, so fc2's wgrad and the fused gelu' dgrad epilogue see the dropped activation without knowing) and
    behind fc2 (mask ``site + 1``, applied with the residual add)."""
    w116 = bf16_of(w1)
    b132 = _f32(b1) if b1 is not None else None
    if pre is not None:
        ln, mean, rstd = pre
        pre_act, h = linear_fwd(ln, w116, b132, gelu=True if keep else "fwd_only")
    elif ln_gemm_ok(x, w116):
        (pre_act, h), ln, mean, rstd = ln_linear_fwd(x, _f32(ln_w), _f32(ln_b), eps, w116, b132,
                                                     gelu=True if keep else "fwd_only", keep=keep)
    else:
        ln, mean, rstd = layernorm_fwd(x, _f32(ln_w), _f32(ln_b), eps)
        pre_act, h = linear_fwd(ln, w116, b132, gelu=True if keep else "fwd_only")
    if drop is not None:
        dropout(h, drop, site, second=pre_act, out=h)
    b232 = _f32(b2) if b2 is not None else None
    y, nxt_state = None, None
    if need_out:
        if drop is not None:
            y = dropout_residual(linear_fwd(h, bf16_of(w2), b232), x, rs, drop, site + 1)
        elif nxt is not None and ln_tail_ok(h.shape[0], w2.shape[0]):
            y, nxt_state = linear_res_ln_fwd(h, bf16_of(w2), b232, x, rs, _f32(nxt[0]), _f32(nxt[1]), nxt[2], keep=keep)
        else:
            y = linear_fwd(h, bf16_of(w2), b232, residual=x, rowscale=rs)
    return y, (x, mean, rstd, ln, pre_act, h, rs), nxt_state


def mlp_half_bwd(dy, saved, ln_w, w1, w2, has_b1, has_b2, ln_b=None, b1=None, b2=None, drop=None, site=0):
    x, mean, rstd, ln, pre, h, rs = saved
    dyb = dy if rs is None else scale_rows(dy, rs)
    if drop is not None:
        dyb = dropout(dyb, drop, site + 1)           # Mlp.drop behind fc2; the one behind the activation lives in pre / h
    d_pre = linear_dgrad(dyb, bf16_of(w2), dgelu_of=pre)
    d_w2, d_b2 = linear_wgrad(dyb, h, has_b2, w2, b2)
    d_ln = linear_dgrad(d_pre, bf16_of(w1))
    d_w1, d_b1 = linear_wgrad(d_pre, ln, has_b1, w1, b1)
    dx, d_lw, d_lb = layernorm_bwd(x, d_ln, _f32(ln_w), mean, rstd, dres=dy, weight=ln_w, bias=ln_b)
    return dx, d_lw, d_lb, d_w1, d_b1, d_w2, d_b2


class AttnHalfFn(torch.autograd.Function):
    """First residual branch of a Block as one autograd node (standalone Block.forward path)."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, qkv_w, qkv_b, proj_w, proj_b, eps, B, N, H, scale, rs=None, drop=None, site=0):
        y, saved, _ = attn_half_fwd(x, ln_w, ln_b, qkv_w, qkv_b, proj_w, proj_b, eps, B, N, H, scale, rs=rs, drop=drop,
                                    site=site)
        ctx.save_for_backward(*saved, ln_w, qkv_w, proj_w)
        ctx.meta = (B, N, H, scale, qkv_b is not None, proj_b is not None)
        ctx.drop = (drop, site)
        return y

    @staticmethod
    def backward(ctx, dy):
        *saved, ln_w, qkv_w, proj_w = ctx.saved_tensors
        B, N, H, scale, has_qb, has_pb = ctx.meta
        dx, *g = attn_half_bwd(_g16(dy), saved, ln_w, qkv_w, proj_w, has_qb, has_pb, B, N, H, scale, drop=ctx.drop[0],
                               site=ctx.drop[1])
        return (dx.to(dy.dtype), *g, None, None, None, None, None, None, None, None)


class MlpHalfFn(torch.autograd.Function):
    """Second residual branch of a Block as one autograd node (standalone Block.forward path)."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, w1, b1, w2, b2, eps, rs=None, drop=None, site=0):
        y, saved, _ = mlp_half_fwd(x, ln_w, ln_b, w1, b1, w2, b2, eps, rs=rs, drop=drop, site=site)
        ctx.save_for_backward(*saved, ln_w, w1, w2)
        ctx.meta = (b1 is not None, b2 is not None)
        ctx.drop = (drop, site)
        return y

    @staticmethod
    def backward(ctx, dy):
        *saved, ln_w, w1, w2 = ctx.saved_tensors
        has_b1, has_b2 = ctx.meta
        dx, *g = mlp_half_bwd(_g16(dy), saved, ln_w, w1, w2, has_b1, has_b2, drop=ctx.drop[0], site=ctx.drop[1])
        return (dx.to(dy.dtype), *g, None, None, None, None)


BLOCK_PARAMS = 12  # ln1 w,b | qkv w,b | proj w,b | ln2 w,b | fc1 w,b | fc2 w,b


class EncoderFn(torch.autograd.Function):
    """All transformer blocks + the final LayerNorm on the CLS rows as ONE autograd node
    (VisionTransformer.forward VT.pyc@L248-253). The residual stream is fp32 forward; gradients flow in
    bf16 between blocks without touching autograd (no per-block dtype casts, ~2*depth fewer graph nodes).

    apply(tokens[B*N, D] fp32, meta, *params) with params = depth * BLOCK_PARAMS tensors (None for absent
    biases) followed by (norm_w, norm_b); meta = (B, N, H, scale, [eps1, eps2] per block, norm_eps).
    B and N may be tuples (``segments``): several crop groups of different resolution packed back to back run
    through every row-wise kernel in one launch; only attention is launched per group.
    Returns the normalised CLS embedding [sum(B), D] bf16."""

    @staticmethod
    def forward(ctx, tok, meta, *params):
        B, N, H, scale, eps_list, norm_eps = meta[:6]
        rs_list = meta[6] if len(meta) > 6 else None   # stochastic depth: per block (rs_attn, rs_mlp) row scales or None
        depth = (len(params) - 2) // BLOCK_PARAMS
        keep = any(ctx.needs_input_grad)  # False under no_grad (teacher): nothing is retained
        # activation recompute (the reference's --grad-checkpointing, train.py:146,509-510): keep only every block's
        # input (the fp32 stream) and rebuild the block's activations in backward
        recompute = keep and len(meta) > 8 and bool(meta[8])
        # element dropout (drop_rate > 0 in training mode): (p, seed) or None. Sites: 0 = pos_drop, block i:
        # 1 + 3 i = attn.proj_drop, 2 + 3 i / 3 + 3 i = mlp.drop behind the activation / behind fc2
        drop = meta[9] if len(meta) > 9 else None
        saved = []
        x = tok
        if drop is not None:
            dropout(tok, drop, 0, out=tok)   # pos_drop (VT.pyc@L245), in place: nothing else reads the token stream
        ln_state = None
        for i in range(depth):
            ln1w, ln1b, qw, qb, pw, pb, ln2w, ln2b, w1, b1, w2, b2 = params[i * BLOCK_PARAMS:(i + 1) * BLOCK_PARAMS]
            rs1, rs2 = rs_list[i] if rs_list is not None else (None, None)
            x_in = x
            # the LayerNorm after each residual add rides in the tail of the GEMM that produces the sum (when the shape
            # allows): norm2 with attn.proj, the NEXT block's norm1 with mlp.fc2. Not under activation recompute.
            nxt1 = None if recompute else (ln2w, ln2b, eps_list[i][1])
            nxt2 = None
            if not recompute and i + 1 < depth:
                nb = params[(i + 1) * BLOCK_PARAMS:(i + 1) * BLOCK_PARAMS + 2]
                nxt2 = (nb[0], nb[1], eps_list[i + 1][0])
            x, s1, ln_state = attn_half_fwd(x, ln1w, ln1b, qw, qb, pw, pb, eps_list[i][0], B, N, H, scale,
                                            keep=keep and not recompute, rs=rs1, pre=ln_state, nxt=nxt1, drop=drop,
                                            site=1 + 3 * i)
            x, s2, ln_state = mlp_half_fwd(x, ln2w, ln2b, w1, b1, w2, b2, eps_list[i][1], keep=keep and not recompute,
                                           rs=rs2, pre=ln_state, nxt=nxt2, drop=drop, site=2 + 3 * i)
            if recompute:
                saved.append(x_in)
            elif keep:
                saved.append((s1, s2))
        D = x.shape[1]
        segs = segments(B, N)
        # x[:, 0] of every group gathered by one strided row copy each (VT.pyc@L253)
        cls = torch.empty(sum(b for b, _, _ in segs), D, dtype=x.dtype, device=x.device)
        es, b0 = x.element_size(), 0
        for b, n, r0 in segs:
            copy_rows(x[r0], n * D * es, cls[b0], D * es, b, D * es)
            b0 += b
        norm_w, norm_b = params[-2], params[-1]
        out, mean, rstd = layernorm_fwd(cls, _f32(norm_w), _f32(norm_b), norm_eps)
        if keep:
            ctx.saved = saved          # activations live exactly as long as the graph node
            ctx.final = (cls, mean, rstd)
            ctx.params = params
            ctx.meta = meta
        return out

    @staticmethod
    def backward(ctx, dout):
        B, N, H, scale, eps_list, norm_eps = ctx.meta[:6]
        params = ctx.params
        depth = (len(params) - 2) // BLOCK_PARAMS
        drop = ctx.meta[9] if len(ctx.meta) > 9 else None
        cls, mean, rstd = ctx.final
        d_cls, d_nw, d_nb = layernorm_bwd(cls, _g16(dout), _f32(params[-2]), mean, rstd, weight=params[-2],
                                          bias=params[-1])
        D = cls.shape[1]
        segs = segments(B, N)
        dx = zeros((sum(b * n for b, n, _ in segs), D), _BF16, cls.device)
        b0 = 0
        for b, n, r0 in segs:
            copy_rows(d_cls[b0], D * 2, dx[r0], n * D * 2, b, D * 2)   # only the CLS rows carry gradient
            b0 += b
        grads = [None] * len(params)
        for i in range(depth - 1, -1, -1):
            ln1w, ln1b, qw, qb, pw, pb, ln2w, ln2b, w1, b1, w2, b2 = params[i * BLOCK_PARAMS:(i + 1) * BLOCK_PARAMS]
            if torch.is_tensor(ctx.saved[i]):   # activation recompute: rebuild this block's saved tensors from its input
                rs1, rs2 = ctx.meta[6][i] if ctx.meta[6] is not None else (None, None)
                x1, s1, _ = attn_half_fwd(ctx.saved[i], ln1w, ln1b, qw, qb, pw, pb, eps_list[i][0], B, N, H, scale, rs=rs1,
                                          drop=drop, site=1 + 3 * i)
                _, s2, _ = mlp_half_fwd(x1, ln2w, ln2b, w1, b1, w2, b2, eps_list[i][1], rs=rs2, need_out=False, drop=drop,
                                        site=2 + 3 * i)
            else:
                s1, s2 = ctx.saved[i]
            dx, g_l2w, g_l2b, g_w1, g_b1, g_w2, g_b2 = mlp_half_bwd(dx, s2, ln2w, w1, w2, b1 is not None, b2 is not None,
                                                                    ln2b, b1, b2, drop=drop, site=2 + 3 * i)
            dx, g_l1w, g_l1b, g_qw, g_qb, g_pw, g_pb = attn_half_bwd(dx, s1, ln1w, qw, pw, qb is not None,
                                                                      pb is not None, B, N, H, scale, ln1b, qb, pb,
                                                                      drop=drop, site=1 + 3 * i)
            grads[i * BLOCK_PARAMS:(i + 1) * BLOCK_PARAMS] = [g_l1w, g_l1b, g_qw, g_qb, g_pw, g_pb, g_l2w, g_l2b,
                                                              g_w1, g_b1, g_w2, g_b2]
            ctx.saved[i] = None  # free this block's activations as soon as they are consumed
        grads[-2], grads[-1] = d_nw, d_nb
        if drop is not None and ctx.needs_input_grad[0]:
            dropout(dx, drop, 0, out=dx)     # pos_drop backward: the token gradient through the same mask
        relay = ctx.meta[7] if len(ctx.meta) > 7 else None
        if not ctx.needs_input_grad[0]:
            d_tok = None
        elif relay is not None:
            d_tok = relay.put(dx, dx.shape)
        else:
            d_tok = dx.float()
        return (d_tok, None, *grads)


def _tokens_param_grads(cls_token, pos_embed, dcls_parts, dtables, mats):
    """Gradients of cls_token / pos_embed from the per-group token-table gradients. With registered gradient sinks
    the contributions are accumulated straight into ``.grad`` (and None is returned for autograd); otherwise into
    fresh zero-filled buffers. No PyTorch kernels either way."""
    D = dtables[0].shape[1]
    out = []
    for param, parts in ((cls_token, dcls_parts), (pos_embed, None)):
        if param is None or not param.requires_grad:
            out.append(None)
            continue
        sink = _sink(param)
        buf = sink if sink is not None else zeros(tuple(param.shape), torch.float32, dtables[0].device)
        flat = buf.view(-1, D)
        if parts is not None:
            for d in parts:
                add_f32(flat[0], d)
        else:
            for dt, mat in zip(dtables, mats):
                pos_table_bwd(dt, mat, flat)
        if sink is not None:
            _sunk(param)
            out.append(None)
        else:
            out.append(buf if param.dtype == torch.float32 else buf.to(param.dtype))
    return out[0], out[1]


class TokensFn(torch.autograd.Function):
    """prepare_tokens (VT.pyc@L235-246): patch-embed conv as GEMM, prepend CLS, add position table.
    img [B,C,H,W] bf16 -> tokens [B*(Np+1), D] fp32. ``pos_embed`` is the stored table [1, 1+Ki, D]; ``mat`` (None at
    the native resolution) the cached linear map of its bicubic resize, applied here forward and transposed backward."""

    @staticmethod
    def forward(ctx, img, proj_w, proj_b, cls_token, pos_embed, patch, relay=None, mat=None):
        ctx.relay = relay
        B, C, H, W = img.shape
        Np = (H // patch) * (W // patch)
        D = proj_w.shape[0]
        Kp = C * patch * patch
        cols = torch.empty(B * Np, Kp, dtype=_BF16, device=img.device)
        _call("b200ssl_patchify", img.data_ptr(), cols.data_ptr(), B, C, H, W, patch, _stream())
        w16 = bf16_of(proj_w).view(D, Kp)
        y = linear_fwd(cols, w16, _f32(proj_b) if proj_b is not None else None)
        x = torch.empty(B * (Np + 1), D, dtype=torch.float32, device=img.device)
        table = pos_table(_f32(pos_embed).view(-1, D), mat)
        if table.shape[0] != Np + 1:
            raise RuntimeError(f"position table has {table.shape[0]} rows, the image needs {Np + 1}")
        _call("b200ssl_assemble_tokens", y.data_ptr(), _f32(cls_token).data_ptr(), table.data_ptr(),
              x.data_ptr(), B, Np, D, _stream())
        ctx.save_for_backward(cols, proj_w)
        ctx.params = (proj_b, cls_token, pos_embed, mat)
        ctx.meta = (B, Np, D, proj_b is not None)
        return x

    @staticmethod
    def backward(ctx, dx):
        cols, proj_w = ctx.saved_tensors
        proj_b, cls_token, pos_embed, mat = ctx.params
        B, Np, D, has_bias = ctx.meta
        dx = ctx.relay.take(dx) if ctx.relay is not None else _g16(dx)
        dy = torch.empty(B * Np, D, dtype=_BF16, device=dx.device)
        dtab = torch.empty(Np + 1, D, dtype=torch.float32, device=dx.device)
        dcls = torch.empty(D, dtype=torch.float32, device=dx.device)
        _call("b200ssl_assemble_tokens_bwd", dx.data_ptr(), dy.data_ptr(), dtab.data_ptr(), dcls.data_ptr(), B, Np,
              D, _stream(), launches=1)
        dw, db = linear_wgrad(dy, cols, has_bias, proj_w, proj_b)
        g_cls, g_pos = _tokens_param_grads(cls_token, pos_embed, [dcls], [dtab], [mat])
        return None, (dw.view(proj_w.shape) if dw is not None else None), db, g_cls, g_pos, None, None, None


class MultiTokensFn(torch.autograd.Function):
    """prepare_tokens for several crop groups of different resolution in one node: per-group patch gather into one
    column buffer, ONE patch-embed GEMM over all patches, per-group CLS / position assembly into one packed fp32
    token stream (group after group). apply(patch, n, relay, mats, proj_w, proj_b, cls_token, pos_embed, img_1..img_n);
    ``relay`` is a GradRelay or None, ``mats`` the per-group resize maps (None = native resolution)."""

    @staticmethod
    def forward(ctx, patch, n, relay, mats, proj_w, proj_b, cls_token, pos_embed, *imgs):
        ctx.relay = relay
        D = proj_w.shape[0]
        dims = []
        for img in imgs:
            B, C, H, W = img.shape
            dims.append((B, (H // patch) * (W // patch), C, H, W))
        Kp = dims[0][2] * patch * patch
        dev = imgs[0].device
        cols = torch.empty(sum(B * Np for B, Np, *_ in dims), Kp, dtype=_BF16, device=dev)
        c0 = 0
        for img, (B, Np, C, H, W) in zip(imgs, dims):
            _call("b200ssl_patchify", img.data_ptr(), cols[c0:c0 + B * Np].data_ptr(), B, C, H, W, patch, _stream())
            c0 += B * Np
        y = linear_fwd(cols, bf16_of(proj_w).view(D, Kp), _f32(proj_b) if proj_b is not None else None)
        x = torch.empty(sum(B * (Np + 1) for B, Np, *_ in dims), D, dtype=torch.float32, device=dev)
        pos32, cls32 = _f32(pos_embed).view(-1, D), _f32(cls_token)
        c0 = r0 = 0
        for mat, (B, Np, *_) in zip(mats, dims):
            table = pos_table(pos32, mat)
            if table.shape[0] != Np + 1:
                raise RuntimeError(f"position table has {table.shape[0]} rows, the crop needs {Np + 1}")
            _call("b200ssl_assemble_tokens", y[c0:c0 + B * Np].data_ptr(), cls32.data_ptr(), table.data_ptr(),
                  x[r0:r0 + B * (Np + 1)].data_ptr(), B, Np, D, _stream())
            c0 += B * Np
            r0 += B * (Np + 1)
        ctx.save_for_backward(cols, proj_w)
        ctx.params = (proj_b, cls_token, pos_embed, mats)
        ctx.meta = (n, dims, D, proj_b is not None)
        return x

    @staticmethod
    def backward(ctx, dx):
        cols, proj_w = ctx.saved_tensors
        proj_b, cls_token, pos_embed, mats = ctx.params
        n, dims, D, has_bias = ctx.meta
        dx = ctx.relay.take(dx) if ctx.relay is not None else _g16(dx)
        dy = torch.empty(cols.shape[0], D, dtype=_BF16, device=dx.device)
        dcls = torch.empty(n, D, dtype=torch.float32, device=dx.device)
        dtabs = []
        c0 = r0 = 0
        for i, (B, Np, *_) in enumerate(dims):
            dtab = torch.empty(Np + 1, D, dtype=torch.float32, device=dx.device)
            _call("b200ssl_assemble_tokens_bwd", dx[r0:r0 + B * (Np + 1)].data_ptr(), dy[c0:c0 + B * Np].data_ptr(),
                  dtab.data_ptr(), dcls[i].data_ptr(), B, Np, D, _stream(), launches=1)
            dtabs.append(dtab)
            c0 += B * Np
            r0 += B * (Np + 1)
        dw, db = linear_wgrad(dy, cols, has_bias, proj_w, proj_b)
        g_cls, g_pos = _tokens_param_grads(cls_token, pos_embed, [dcls[i] for i in range(n)], dtabs, mats)
        return (None, None, None, None, (dw.view(proj_w.shape) if dw is not None else None), db, g_cls, g_pos,
                *([None] * n))


class L2NormFn(torch.autograd.Function):
    """F.normalize(x, dim=-1, p=2) with the reference's eps=1e-12 clamp (VT.pyc@L328)."""

    @staticmethod
    def forward(ctx, x, eps):
        rows, D = x.shape
        y = torch.empty_like(x)
        norm = torch.empty(rows, dtype=torch.float32, device=x.device)
        _call("b200ssl_l2norm_fwd", x.data_ptr(), y.data_ptr(), norm.data_ptr(), rows, D, float(eps), _stream())
        ctx.save_for_backward(y, norm)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, dy):
        y, norm = ctx.saved_tensors
        dx = torch.empty_like(y)
        _call("b200ssl_l2norm_bwd", y.data_ptr(), dy.contiguous().data_ptr(), norm.data_ptr(), dx.data_ptr(),
              y.shape[0], y.shape[1], float(ctx.eps), _stream())
        return dx, None


class WeightNormLinearFn(torch.autograd.Function):
    """y = x (g * v / ||v||_row)^T — the weight-normed, bias-free last layer (VT.pyc@L315-318,329)."""

    @staticmethod
    def forward(ctx, x, weight_v, weight_g):
        Kout, D = weight_v.shape
        v32 = _f32(weight_v)
        g32 = _f32(weight_g).view(-1) if weight_g is not None else None
        w16 = torch.empty(Kout, D, dtype=_BF16, device=x.device)
        norm = torch.empty(Kout, dtype=torch.float32, device=x.device)
        _call("b200ssl_weightnorm_fwd", v32.data_ptr(), _ptr(g32), w16.data_ptr(), norm.data_ptr(), Kout, D,
              _stream())
        y = linear_fwd(x, w16)
        ctx.save_for_backward(x, weight_v, weight_g, w16, norm)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight_v, weight_g, w16, norm = ctx.saved_tensors
        dy = dy.contiguous()
        dx = linear_dgrad(dy, w16) if ctx.needs_input_grad[0] else None
        dv = dg = None
        if ctx.needs_input_grad[1] or (weight_g is not None and ctx.needs_input_grad[2]):
            dw, _ = linear_wgrad(dy, x, need_bias=False)
            v32 = _f32(weight_v)
            g32 = _f32(weight_g).view(-1) if weight_g is not None else None
            want_g = weight_g is not None and ctx.needs_input_grad[2]
            v_sink = _sink(weight_v) if ctx.needs_input_grad[1] else None
            g_sink = _sink(weight_g) if want_g else None
            # with gradient sinks the result is added straight into .grad (no autograd accumulation kernel)
            sunk = v_sink is not None and (not want_g or g_sink is not None)
            dv = v_sink if sunk else torch.empty_like(v32)
            dg = (g_sink.view(-1) if sunk else torch.empty(v32.shape[0], dtype=torch.float32, device=x.device)) \
                if want_g else None
            _call("b200ssl_weightnorm_bwd", v32.data_ptr(), _ptr(g32), norm.data_ptr(), dw.data_ptr(),
                  dv.data_ptr(), _ptr(dg), v32.shape[0], v32.shape[1], int(sunk), _stream())
            if sunk:
                _sunk(weight_v)
                if want_g:
                    _sunk(weight_g)
                dv = dg = None
            elif dg is not None:
                dg = dg.view(weight_g.shape)
        return dx, dv, dg


class BatchNormGeluFn(torch.autograd.Function):
    """gelu(BatchNorm1d(x)) for DINOHead(use_bn=True) (VT.pyc@L304-305,309-310): batch statistics in training mode
    (running statistics and num_batches_tracked updated in place, torch semantics), running statistics in eval."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, num_batches, momentum, eps, training, gelu=True):
        rows, C = x.shape
        use_batch = bool(training or running_mean is None)
        h = torch.empty_like(x)
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        invstd = torch.empty(C, dtype=torch.float32, device=x.device)
        if momentum is None:
            # nn.BatchNorm1d(momentum=None): cumulative moving average, factor 1 / num_batches_tracked (after increment)
            raise NotImplementedError("BatchNorm1d(momentum=None) is not on the b200ssl path")
        w32 = _f32(weight) if weight is not None else None
        b32 = _f32(bias) if bias is not None else None
        _call("b200ssl_bn_gelu_fwd", x.data_ptr(), _ptr(w32), _ptr(b32), _ptr(running_mean), _ptr(running_var),
              _ptr(num_batches) if use_batch and training else None, h.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
              rows, C, float(momentum), float(eps), int(use_batch), int(gelu), _stream())
        ctx.save_for_backward(x, mean, invstd)
        ctx.params = (weight, bias)
        ctx.flags = (use_batch, gelu)
        return h

    @staticmethod
    def backward(ctx, dh):
        x, mean, invstd = ctx.saved_tensors
        weight, bias = ctx.params
        use_batch, gelu = ctx.flags
        rows, C = x.shape
        dx = torch.empty_like(x)
        w_sink, b_sink = _sink(weight), _sink(bias)
        sunk = weight is not None and bias is not None and w_sink is not None and b_sink is not None
        if sunk:
            dw, db = w_sink, b_sink
        else:
            dwb = zeros((2, C), torch.float32, x.device)
            dw, db = dwb[0], dwb[1]
        _call("b200ssl_bn_gelu_bwd", x.data_ptr(), _g16(dh).data_ptr(), _ptr(_f32(weight) if weight is not None else None),
              _ptr(_f32(bias) if bias is not None else None), mean.data_ptr(), invstd.data_ptr(), dx.data_ptr(),
              dw.data_ptr(), db.data_ptr(), rows, C, int(use_batch), int(gelu), _stream())
        if sunk:
            _sunk(weight)
            _sunk(bias)
            dw = db = None
        return (dx, dw if weight is not None else None, db if bias is not None else None, None, None, None, None, None,
                None, None)


class DinoLossFn(torch.autograd.Function):
    """Fused teacher-centred cross-entropy (oracle/dino.py::DINOLoss.forward without the centre update)."""

    @staticmethod
    def forward(ctx, student, teacher, center, ncrops, student_temp, teacher_temp):
        rows, K = student.shape
        B = rows // ncrops
        loss = torch.empty(1, dtype=torch.float32, device=student.device)
        s_lse = torch.empty(rows, dtype=torch.float32, device=student.device)
        t_lse = torch.empty(2 * B, dtype=torch.float32, device=student.device)
        _call("b200ssl_dino_loss_fwd", student.data_ptr(), teacher.data_ptr(), center.data_ptr(), loss.data_ptr(),
              s_lse.data_ptr(), t_lse.data_ptr(), B, ncrops, K, float(student_temp), float(teacher_temp), _stream(),
              launches=2)
        ctx.save_for_backward(student, teacher, clone(center), s_lse, t_lse)   # the centre moves before backward runs
        ctx.meta = (B, ncrops, K, float(student_temp), float(teacher_temp))
        return loss.view(())

    @staticmethod
    def backward(ctx, gout):
        student, teacher, center, s_lse, t_lse = ctx.saved_tensors
        B, ncrops, K, ts, tt = ctx.meta
        g = gout.detach().to(torch.float32).contiguous()
        ds = torch.empty_like(student)
        _call("b200ssl_dino_loss_bwd", student.data_ptr(), teacher.data_ptr(), center.data_ptr(), s_lse.data_ptr(),
              t_lse.data_ptr(), g.data_ptr(), ds.data_ptr(), B, ncrops, K, ts, tt, _stream())
        return ds, None, None, None, None, None


def teacher_colsum(teacher: torch.Tensor) -> torch.Tensor:
    """fp32 column sums of the raw teacher logits (the local part of the centre update)."""
    rows, K = teacher.shape
    out = torch.empty(K, dtype=torch.float32, device=teacher.device)
    _call("b200ssl_colsum", teacher.data_ptr(), teacher.stride(0), out.data_ptr(), rows, K, 0, _stream(), launches=2)
    return out


def center_update(center: torch.Tensor, batch_sum: torch.Tensor, total_rows: int, momentum: float):
    _call("b200ssl_center_update", center.data_ptr(), batch_sum.data_ptr(), center.numel(), int(total_rows),
          float(momentum), _stream())
