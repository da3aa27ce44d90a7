"""Drop-in replacements for the reference's ``nn_encoder_arch.vision_transformer`` modules, computing
through the b200ssl sm_100a kernels.

Same class names, constructor keyword arguments / defaults, attribute names (hence ``state_dict``
keys) and forward signatures as the reference bytecode
(``nn_encoder_arch/__pycache__/vision_transformer.cpython-37.pyc``; SURVEY.md §8a E1-E11, H1;
``VT.pyc@Lnn`` = original source line). Parameters are fp32 ``nn.Parameter``; arithmetic is bf16 on
the tensor cores with fp32 accumulation and fp32 LayerNorm / softmax statistics. Inputs must be CUDA
tensors on an sm_100 device — there is no CPU or PyTorch-math fallback on the compute path.
"""
from __future__ import annotations

import math
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

__all__ = ["trunc_normal_", "drop_path", "DropPath", "Mlp", "Attention", "Block", "PatchEmbed",
           "VisionTransformer", "vit_tiny", "vit_small", "vit_base", "DINOHead"]


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    """Truncated-normal init via uniform -> erfinv (VT.pyc@L25-63); equivalent to nn.init.trunc_normal_."""
    lo = (1.0 + math.erf((a - mean) / std / math.sqrt(2.0))) / 2.0
    hi = (1.0 + math.erf((b - mean) / std / math.sqrt(2.0))) / 2.0
    with torch.no_grad():
        tensor.uniform_(2 * lo - 1, 2 * hi - 1).erfinv_()
        tensor.mul_(std * math.sqrt(2.0)).add_(mean).clamp_(min=a, max=b)
    return tensor


def _no_grad_trunc_normal_(tensor, mean, std, a, b):
    """VT.pyc@L25-57: the worker trunc_normal_ wraps (same routine, positional arguments)."""
    return trunc_normal_(tensor, mean, std, a, b)


def drop_path(x, drop_prob: float = 0.0, training: bool = False):
    """Per-sample stochastic depth (VT.pyc@L66-74). Identity at the constructor default (0.0)."""
    if drop_prob == 0.0 or not training:
        return x
    keep = 1 - drop_prob
    mask = (keep + torch.rand((x.shape[0],) + (1,) * (x.ndim - 1), dtype=x.dtype, device=x.device)).floor_()
    return x.div(keep) * mask


class DropPath(nn.Module):
    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training)


def _is_plain_layernorm(m) -> bool:
    return isinstance(m, nn.LayerNorm) and m.elementwise_affine and m.bias is not None


def _layer_norm(mod: nn.LayerNorm, x2d: torch.Tensor) -> torch.Tensor:
    return ops.LayerNormFn.apply(x2d, mod.weight, mod.bias, mod.eps)


def draw_dropout_seed(device):
    """One 64-bit seed for the dropout masks of a model call, drawn ON the device from torch's CUDA generator: seeded
    runs reproduce, and a captured CUDA graph draws a fresh one on every replay (the generator's offset is a graph
    input). The kernels derive every mask from (seed, site, element index) -- ops.dropout."""
    return torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=device)


class Mlp(nn.Module):
    """fc1 -> GELU(erf) -> drop -> fc2 -> drop (VT.pyc@L88-104)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def _fusable(self):
        return isinstance(self.act, nn.GELU) and getattr(self.act, "approximate", "none") == "none"

    def forward(self, x):
        ops.require_cuda(x, "Mlp")
        if not self._fusable():
            raise NotImplementedError("b200ssl Mlp supports the exact-erf GELU (reference default)")
        if self.training and self.drop.p > 0:
            # the bare module (Block / VisionTransformer fuse their dropouts into the residual path instead)
            shape, dtype = x.shape, x.dtype
            self._last_dropout_seed = draw_dropout_seed(x.device)
            y = ops.MlpDropFn.apply(ops.to_bf16_2d(x), self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias,
                                    self.drop.p, self._last_dropout_seed)
            return y.view(*shape[:-1], -1).to(dtype)
        shape, dtype = x.shape, x.dtype
        y = ops.MlpChainFn.apply(ops.to_bf16_2d(x), None, self.fc1.weight, self.fc1.bias, self.fc2.weight,
                                 self.fc2.bias)
        return y.view(*shape[:-1], -1).to(dtype)


class Attention(nn.Module):
    """Multi-head self-attention (VT.pyc@L107-131). ``forward`` returns ``(x, attn)`` like the
    reference; the probability map is materialised only on this standalone path (inspection /
    get_last_selfattention) — Block's training path uses the fused kernel and never forms it."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def _check(self, C):
        if C // self.num_heads != 64 or C % self.num_heads:
            raise NotImplementedError("b200ssl attention kernels are built for head_dim 64 (vit_tiny/small/base)")
        if self.training and self.attn_drop.p > 0:
            raise NotImplementedError("dropout on the attention probabilities (attn_drop_rate > 0) is not on the b200ssl "
                                      "hot path: the fused kernel never forms them (reference default 0; train.py has "
                                      "no flag for it)")

    def forward(self, x):
        return self._forward(x)

    def _forward(self, x, map_only=False):
        ops.require_cuda(x, "Attention")
        B, N, C = x.shape
        self._check(C)
        dtype = x.dtype
        x2 = ops.to_bf16_2d(x)
        qkv = ops.LinearFn.apply(x2, self.qkv.weight, self.qkv.bias, None)
        out = ops.AttentionCoreFn.apply(qkv, B, N, self.num_heads, self.scale)
        y = ops.LinearFn.apply(out, self.proj.weight, self.proj.bias, None)
        if self.training and self.proj_drop.p > 0 and not map_only:   # the bare module: proj_drop as its own node
            self._last_dropout_seed = draw_dropout_seed(x.device)
            y = ops.DropoutFn.apply(y, self.proj_drop.p, self._last_dropout_seed, 0)
        with torch.no_grad():  # probability map for callers that ask for it (not the training path)
            q, k = qkv.view(B, N, 3, self.num_heads, 64)[:, :, :2].permute(2, 0, 3, 1, 4).float()
            attn = ((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1).to(dtype)
        return y.view(B, N, C).to(dtype), attn


class Block(nn.Module):
    """Pre-LN transformer block (VT.pyc@L134-152)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        mlp_hidden_dim = int(dim * mlp_ratio)
        self.mlp = Mlp(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, drop=drop)

    def _fused_ok(self):
        return _is_plain_layernorm(self.norm1) and _is_plain_layernorm(self.norm2) and self.mlp._fusable()

    def _drop_path_scales(self, B, N, device):
        """Stochastic depth (VT.pyc@L66-85,150-151) for the fused path: one fp32 scale per token row and residual
        branch -- mask[b] / keep, drawn exactly like drop_path() does (torch.rand of shape [B,1,1], attention branch
        first) so that a seeded run matches the reference -- or (None, None) when the block drops nothing."""
        dp = self.drop_path
        if isinstance(dp, nn.Identity) or not self.training or not dp.drop_prob:
            return None, None
        keep = 1.0 - dp.drop_prob
        out = []
        for _ in range(2):
            mask = (keep + torch.rand((B, 1, 1), dtype=torch.float32, device=device)).floor_()
            out.append((mask.view(B) / keep).repeat_interleave(N).contiguous())
        return out[0], out[1]

    def _param_list(self):
        a, m = self.attn, self.mlp
        return [self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias,
                self.norm2.weight, self.norm2.bias, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias]

    def forward_tokens(self, x2, B, N, seed=None, index=0):
        """Fused path on the packed fp32 token stream [B*N, C]: two autograd nodes per block. ``seed`` / ``index``:
        the dropout seed of the enclosing model call and this block's position in it (mask sites 1 + 3 index ...);
        a Block called on its own draws a seed per call."""
        a = self.attn
        a._check(x2.shape[1])
        if not self._fused_ok():
            raise NotImplementedError("b200ssl Block: custom norm or activation layers are not on the accelerated "
                                      "path (SURVEY.md §8f)")
        rs1, rs2 = self._drop_path_scales(B, N, x2.device)
        drop_a = drop_m = None
        if self.training and (a.proj_drop.p > 0 or self.mlp.drop.p > 0):
            seed = draw_dropout_seed(x2.device) if seed is None else seed
            drop_a = (a.proj_drop.p, seed) if a.proj_drop.p > 0 else None
            drop_m = (self.mlp.drop.p, seed) if self.mlp.drop.p > 0 else None
        x2 = ops.AttnHalfFn.apply(x2, self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias, a.proj.weight,
                                  a.proj.bias, self.norm1.eps, B, N, a.num_heads, a.scale, rs1, drop_a, 1 + 3 * index)
        m = self.mlp
        return ops.MlpHalfFn.apply(x2, self.norm2.weight, self.norm2.bias, m.fc1.weight, m.fc1.bias, m.fc2.weight,
                                   m.fc2.bias, self.norm2.eps, rs2, drop_m, 2 + 3 * index)

    def forward(self, x, return_attention=False):
        ops.require_cuda(x, "Block")
        B, N, C = x.shape
        if return_attention:
            y, attn = self.attn._forward(_layer_norm(self.norm1, ops.to_stream_2d(x)).view(B, N, C), map_only=True)
            return attn.to(x.dtype)
        dtype = x.dtype
        return self.forward_tokens(ops.to_stream_2d(x), B, N).view(B, N, C).to(dtype)


class PatchEmbed(nn.Module):
    """Image to patch embedding (VT.pyc@L155-170): Conv2d(k=s=P) expressed as a GEMM over gathered patches."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        num_patches = (img_size // patch_size) * (img_size // patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.num_patches = num_patches
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        ops.require_cuda(x, "PatchEmbed")
        B, C, H, W = x.shape
        D = self.proj.weight.shape[0]
        Np = (H // self.patch_size) * (W // self.patch_size)
        zeros_cls = torch.zeros(D, device=x.device)
        zeros_pos = torch.zeros(Np + 1, D, device=x.device)
        tok = ops.TokensFn.apply(x.to(torch.bfloat16).contiguous(), self.proj.weight, self.proj.bias, zeros_cls,
                                 zeros_pos, self.patch_size)
        return tok.view(B, Np + 1, D)[:, 1:].to(x.dtype)  # fp32 stream -> caller's dtype


class VisionTransformer(nn.Module):
    """Vision Transformer (VT.pyc@L173-272); ``forward`` returns the normalised CLS embedding."""

    def __init__(self, img_size=[224], patch_size=16, in_chans=3, num_classes=0, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0,
                 drop_path_rate=0.0, norm_layer=nn.LayerNorm, **kwargs):
        super().__init__()
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size=img_size[0], patch_size=patch_size, in_chans=in_chans,
                                      embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer)
            for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        trunc_normal_(self.pos_embed, std=0.02)
        trunc_normal_(self.cls_token, std=0.02)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def interpolate_pos_encoding(self, x, w, h):
        """Bicubic resize of the patch position table for non-native resolutions (VT.pyc@L213-233).
        A [Np+1, D]-sized host-side table op (PyTorch), differentiable back to ``pos_embed``."""
        npatch = x.shape[1] - 1
        N = self.pos_embed.shape[1] - 1
        if npatch == N and w == h:
            return self.pos_embed
        class_pos_embed = self.pos_embed[:, 0]
        patch_pos_embed = self.pos_embed[:, 1:]
        dim = x.shape[-1]
        w0 = w // self.patch_embed.patch_size
        h0 = h // self.patch_embed.patch_size
        w0, h0 = w0 + 0.1, h0 + 0.1
        # The resize is a fixed linear map of the table: applied as a cached [Np_out, Np_in] matrix (_pos_matrix).
        # This public method keeps plain differentiable torch ops; the training path applies the same matrix with
        # its own kernel (ops.pos_table).
        mat = self._pos_matrix(w, h)
        patch_pos_embed = torch.matmul(mat, patch_pos_embed[0].to(torch.float32)).to(patch_pos_embed.dtype).unsqueeze(0)
        return torch.cat((class_pos_embed.unsqueeze(0), patch_pos_embed), dim=1)

    def _pos_matrix(self, w, h):
        """The resize of ``interpolate_pos_encoding`` for a w x h input as a cached [Np_out, Np_in] fp32 matrix, or None
        at the native resolution (``npatch == N and w == h``, VT.pyc@L216-217). Built once per resolution by pushing
        the identity through the very same ``F.interpolate(scale_factor=..., mode='bicubic')`` call (incl. the +0.1
        fudge and the output-size assertion, @L225-231); the hot path applies it with b200ssl_pos_interp."""
        P = self.patch_embed.patch_size
        npatch = (w // P) * (h // P)
        N = self.pos_embed.shape[1] - 1
        if npatch == N and w == h:
            return None
        key = (w, h, N, str(self.pos_embed.device))
        cache = self.__dict__.setdefault("_interp_cache", {})
        mat = cache.get(key)
        if mat is None:
            w0, h0 = w // P + 0.1, h // P + 0.1
            side = int(math.sqrt(N))
            with torch.no_grad():
                eye = torch.eye(N, device=self.pos_embed.device, dtype=torch.float32)
                out = F.interpolate(eye.reshape(1, side, side, N).permute(0, 3, 1, 2),
                                    scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode="bicubic")
                assert int(w0) == out.shape[-2] and int(h0) == out.shape[-1]
                mat = out.permute(0, 2, 3, 1).reshape(-1, N).contiguous()   # [Np_out, Np_in]
            cache[key] = mat
        return mat

    def _tokens(self, x, relay=None):
        """prepare_tokens on the packed layout: returns ([B*N, D] fp32 stream, B, N)."""
        ops.require_cuda(x, "VisionTransformer")
        B, nc, w, h = x.shape
        P = self.patch_embed.patch_size
        tok = ops.TokensFn.apply(x.to(torch.bfloat16).contiguous(), self.patch_embed.proj.weight,
                                 self.patch_embed.proj.bias, self.cls_token, self.pos_embed, P, relay,
                                 self._pos_matrix(w, h))
        return tok, B, (w // P) * (h // P) + 1

    def _dropout(self, device):
        """(p, seed) for this call when element dropout is on (training mode, drop_rate > 0), else None. The whole
        encoder runs as one node with one rate: every block must carry the rate of ``pos_drop`` (the constructor's
        ``drop_rate``, VT.pyc@L196-200)."""
        p = self.pos_drop.p
        rates = {p} | {b.attn.proj_drop.p for b in self.blocks} | {b.mlp.drop.p for b in self.blocks}
        if not self.training or rates == {0.0}:
            return None
        if len(rates) != 1:
            raise NotImplementedError(f"b200ssl VisionTransformer: one dropout rate for pos_drop / proj_drop / mlp.drop "
                                      f"(the constructor's drop_rate), got {sorted(rates)}")
        self._last_dropout_seed = draw_dropout_seed(device)
        return (p, self._last_dropout_seed)

    def _pos_drop(self, tok, drop):
        """pos_drop as its own autograd node, for the callers that walk the blocks themselves."""
        return tok if drop is None else ops.DropoutFn.apply(tok, drop[0], drop[1], 0)

    def prepare_tokens(self, x):
        tok, B, N = self._tokens(x)
        tok = self._pos_drop(tok, self._dropout(tok.device))
        return tok.view(B, N, -1).to(x.dtype)

    def set_grad_checkpointing(self, enable=True):
        """timm's API behind the reference's ``--grad-checkpointing`` (train.py:146,509-510): keep only each block's
        input in the forward and recompute the block's activations in backward (≈ 1/3 more forward work, the per-block
        activation memory drops from ~23 to 4 bytes per token feature)."""
        self.grad_checkpointing = bool(enable)

    def _encode(self, tok, B, N, rs_list, relay=None, drop=None):
        """All blocks + the final norm on the CLS rows as one autograd node. ``B`` / ``N`` are ints, or tuples when
        ``tok`` packs several crop groups (``forward_multi``)."""
        blk0 = self.blocks[0]
        blk0.attn._check(tok.shape[1])
        if not (all(b._fused_ok() for b in self.blocks) and _is_plain_layernorm(self.norm)):
            raise NotImplementedError("b200ssl VisionTransformer: custom norm or activation layers are not on the "
                                      "accelerated path (SURVEY.md §8f)")
        params = []
        for b in self.blocks:
            params += b._param_list()
        params += [self.norm.weight, self.norm.bias]
        if all(r[0] is None for r in rs_list):
            rs_list = None
        meta = (B, N, blk0.attn.num_heads, blk0.attn.scale, [(b.norm1.eps, b.norm2.eps) for b in self.blocks],
                self.norm.eps, rs_list, relay, getattr(self, "grad_checkpointing", False), drop)
        # the reference normalises every token then keeps row 0 (@L252-253); only CLS rows are normalised here
        return ops.EncoderFn.apply(tok, meta, *params)

    def forward(self, x):
        relay = ops.GradRelay()   # the token stream is private to this call: its gradient travels in bf16
        tok, B, N = self._tokens(x, relay)
        rs_list = [b._drop_path_scales(B, N, tok.device) for b in self.blocks]
        return self._encode(tok, B, N, rs_list, relay, self._dropout(tok.device)).to(x.dtype)

    def forward_multi(self, xs):
        """``torch.cat([self(x) for x in xs])`` for image batches of DIFFERENT resolution (the multi-crop student:
        global 224^2 crops, then local 96^2 crops) in ONE pass: the token rows of all groups are packed back to back,
        so every row-wise kernel (patch-embed GEMM, LayerNorm, qkv / proj / fc1 / fc2 and their dgrad / wgrad) is
        launched once over all rows instead of once per group; attention runs per group on its row range. Same
        arithmetic per row as ``forward`` -- measured 1.5 ms (2.7 %) per ViT-S/16 step from fewer kernel tails
        (tests/gpu_checks/merge_rows_bench.py). Stochastic-depth masks are drawn in the reference's order (group
        by group, block by block), so seeded runs match the unmerged path exactly."""
        xs = list(xs)
        if len(xs) == 1:
            return self.forward(xs[0])
        for x in xs:
            ops.require_cuda(x, "VisionTransformer")
        P = self.patch_embed.patch_size
        mats = tuple(self._pos_matrix(x.shape[2], x.shape[3]) for x in xs)
        relay = ops.GradRelay()
        tok = ops.MultiTokensFn.apply(P, len(xs), relay, mats, self.patch_embed.proj.weight,
                                      self.patch_embed.proj.bias, self.cls_token, self.pos_embed,
                                      *[x.to(torch.bfloat16).contiguous() for x in xs])
        Bs = tuple(int(x.shape[0]) for x in xs)
        Ns = tuple((x.shape[2] // P) * (x.shape[3] // P) + 1 for x in xs)
        per_group = [[b._drop_path_scales(B, N, tok.device) for b in self.blocks] for B, N in zip(Bs, Ns)]
        rs_list = []
        for i in range(len(self.blocks)):
            pair = tuple(None if per_group[0][i][j] is None else torch.cat([g[i][j] for g in per_group])
                         for j in range(2))
            rs_list.append(pair)
        return self._encode(tok, Bs, Ns, rs_list, relay, self._dropout(tok.device)).to(xs[0].dtype)

    def get_last_selfattention(self, x):
        tok, B, N = self._tokens(x)
        drop = self._dropout(tok.device)
        seed = drop[1] if drop is not None else None
        tok = self._pos_drop(tok, drop)
        for i, blk in enumerate(self.blocks):
            if i < len(self.blocks) - 1:
                tok = blk.forward_tokens(tok, B, N, seed, i)
            else:
                return blk(tok.view(B, N, -1), return_attention=True).to(x.dtype)

    def get_intermediate_layers(self, x, n=1):
        tok, B, N = self._tokens(x)
        drop = self._dropout(tok.device)
        seed = drop[1] if drop is not None else None
        tok = self._pos_drop(tok, drop)
        output = []
        for i, blk in enumerate(self.blocks):
            tok = blk.forward_tokens(tok, B, N, seed, i)
            if len(self.blocks) - i <= n:
                output.append(_layer_norm(self.norm, tok).view(B, N, -1).to(x.dtype))
        return output


def vit_tiny(patch_size=16, **kwargs):
    return VisionTransformer(patch_size=patch_size, embed_dim=192, depth=12, num_heads=3, mlp_ratio=4,
                             qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def vit_small(patch_size=16, **kwargs):
    return VisionTransformer(patch_size=patch_size, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4,
                             qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def vit_base(patch_size=16, **kwargs):
    return VisionTransformer(patch_size=patch_size, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4,
                             qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


class DINOHead(nn.Module):
    """Projection head (VT.pyc@L296-330): MLP -> L2 normalise -> weight-normed linear, no bias."""

    def __init__(self, in_dim, out_dim, use_bn=False, norm_last_layer=True, nlayers=3, hidden_dim=2048,
                 bottleneck_dim=256):
        super().__init__()
        nlayers = max(nlayers, 1)
        if nlayers == 1:
            self.mlp = nn.Linear(in_dim, bottleneck_dim)
        else:
            layers = [nn.Linear(in_dim, hidden_dim)]
            if use_bn:
                layers.append(nn.BatchNorm1d(hidden_dim))
            layers.append(nn.GELU())
            for _ in range(nlayers - 2):
                layers.append(nn.Linear(hidden_dim, hidden_dim))
                if use_bn:
                    layers.append(nn.BatchNorm1d(hidden_dim))
                layers.append(nn.GELU())
            layers.append(nn.Linear(hidden_dim, bottleneck_dim))
            self.mlp = nn.Sequential(*layers)
        self.apply(self._init_weights)
        self.last_layer = nn.utils.weight_norm(nn.Linear(bottleneck_dim, out_dim, bias=False))
        self.last_layer.weight_g.data.fill_(1)
        if norm_last_layer:
            self.last_layer.weight_g.requires_grad = False
        self.use_bn = use_bn

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        ops.require_cuda(x, "DINOHead")
        dtype = x.dtype
        x2 = ops.to_bf16_2d(x)
        if self.use_bn and not isinstance(self.mlp, nn.Linear):
            # Linear -> BatchNorm1d -> GELU per hidden layer (VT.pyc@L304-305,309-310): plain GEMM, then ONE kernel for
            # the batch statistics, the normalisation and the GELU
            mods = list(self.mlp)
            i = 0
            while i < len(mods):
                lin = mods[i]
                x2 = ops.LinearFn.apply(x2, lin.weight, lin.bias, None)
                i += 1
                if i < len(mods) and isinstance(mods[i], nn.BatchNorm1d):
                    bn = mods[i]
                    x2 = ops.BatchNormGeluFn.apply(x2, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                                   bn.num_batches_tracked, bn.momentum, bn.eps,
                                                   bn.training or not bn.track_running_stats, True)
                    i += 2   # the GELU that follows is fused
        else:
            linears = [self.mlp] if isinstance(self.mlp, nn.Linear) else [m for m in self.mlp if isinstance(m, nn.Linear)]
            wb = []
            for lin in linears:
                wb += [lin.weight, lin.bias]
            x2 = ops.MlpChainFn.apply(x2, None, *wb)
        x2 = ops.L2NormFn.apply(x2, 1e-12)
        ll = self.last_layer
        y = ops.WeightNormLinearFn.apply(x2, ll.weight_v, ll.weight_g)
        return y if dtype == torch.bfloat16 else y.to(dtype)
