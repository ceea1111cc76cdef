"""Drop-in ``DiT`` (UNet_DS_Diff/DiT_models.py:145-272, configs :334-378) on libb200diff.

Same constructor keywords, ``forward(x, t, y=None, cond=None)`` and ``state_dict`` keys / shapes as the reference
(including the timm sub-module names ``x_embedder.proj``, ``attn.qkv`` / ``attn.proj``, ``mlp.fc1`` / ``mlp.fc2``), so a
reference checkpoint loads with ``load_state_dict(strict=True)``.  The ``nn`` modules are parameter containers; the engine
below runs the network as tensor-core GEMMs with every pointwise step folded into an epilogue or a single fused pass:

* patch embedding = ``b200_patchify`` (NCHW fp32 -> token matrix) + GEMM with bias and the sin-cos ``pos_embed`` as
  the epilogue residual (DiT_models.py:243);
* ``modulate(norm(x), shift, scale)`` (:19-20,120-121,140) = one LayerNorm pass (``b200_layernorm`` with adaLN rows);
* attention (timm ``Attention``: fused qkv Linear, heads are channel slices) = the tcgen05 flash kernel reading q/k/v
  straight out of the qkv GEMM output;
* ``x + gate * proj(...)`` / ``x + gate * fc2(...)`` (:120-121) = GEMM epilogue (bias -> per-sample gate -> residual);
  GELU(tanh) rides in fc1's epilogue;
* the twelve ``adaLN_modulation`` Linears + the final one are ONE batched fp32 GEMV per forward (:118,139);
* final Linear -> fp32 tokens -> ``b200_unpatchify`` (:220-232).

No CPU / eager fallback.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import B200Error
from .transformer import _dpad, _pad_heads_in, _pad_heads_out, linear_tokens
from .unet import default_act_dtype, param_fingerprint

Tensor = torch.Tensor


# ---------------------------------------------------------------------------------------------
# parameter containers (names mirror timm / the reference so state_dict keys are identical)
# ---------------------------------------------------------------------------------------------
class PatchEmbed(nn.Module):
    """timm.models.vision_transformer.PatchEmbed as used at DiT_models.py:192 (no norm layer)."""

    def __init__(self, img_size, patch_size, in_chans, embed_dim, bias=True):
        super().__init__()
        self.img_size, self.patch_size = (img_size, img_size), (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)


class Attention(nn.Module):
    """timm Attention(dim, num_heads, qkv_bias=True) parameters (DiT_models.py:108)."""

    def __init__(self, dim, num_heads=8, qkv_bias=True):
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)


class TimestepEmbedder(nn.Module):
    """DiT_models.py:27-64."""

    def __init__(self, hidden_size, frequency_embedding_size=256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size, bias=True), nn.SiLU(),
                                 nn.Linear(hidden_size, hidden_size, bias=True))
        self.frequency_embedding_size = frequency_embedding_size


class LabelEmbedder(nn.Module):
    """DiT_models.py:67-95 (inference: no label dropout unless force_drop_ids is given)."""

    def __init__(self, num_classes, hidden_size, dropout_prob):
        super().__init__()
        use_cfg_embedding = dropout_prob > 0
        self.embedding_table = nn.Embedding(num_classes + use_cfg_embedding, hidden_size)
        self.num_classes, self.dropout_prob = num_classes, dropout_prob


class DiTBlock(nn.Module):
    """DiT_models.py:102-122; norm1 / norm2 are affine-free LayerNorms (eps 1e-6) and hold no parameters."""

    def __init__(self, hidden_size, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.attn = Attention(hidden_size, num_heads=num_heads, qkv_bias=True)
        self.mlp = Mlp(hidden_size, int(hidden_size * mlp_ratio))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 6 * hidden_size, bias=True))


class FinalLayer(nn.Module):
    """DiT_models.py:125-142."""

    def __init__(self, hidden_size, patch_size, out_channels):
        super().__init__()
        self.linear = nn.Linear(hidden_size, patch_size * patch_size * out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 2 * hidden_size, bias=True))


def get_1d_sincos_pos_embed_from_grid(embed_dim, pos):
    """DiT_models.py:307-326."""
    assert embed_dim % 2 == 0
    omega = 1.0 / 10000 ** (np.arange(embed_dim // 2, dtype=np.float64) / (embed_dim / 2.0))
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def get_2d_sincos_pos_embed(embed_dim, grid_size):
    """DiT_models.py:276-304 (w goes first in the meshgrid; first half of the channels encodes grid[0])."""
    grid_h = np.arange(grid_size, dtype=np.float32)
    grid_w = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(grid_w, grid_h), axis=0).reshape([2, 1, grid_size, grid_size])
    return np.concatenate([get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[0]),
                           get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[1])], axis=1)


def _f(p: Tensor) -> Tensor:
    t = p.detach()
    return t if t.dtype == torch.float32 and t.is_contiguous() else t.float().contiguous()


class _Packed:
    """Kernel-layout weights of one DiT (rebuilt after load_state_dict / .to())."""

    def __init__(self, m: "DiT", dtype: torch.dtype):
        self.dtype = dtype
        pk = lambda w: ops.pack_conv_weight(_f(w), dtype)
        wpe = _f(m.x_embedder.proj.weight)                       # [d, C, p, p] -> one K row of C*p*p (+ zero pad)
        d, k = wpe.shape[0], wpe[0].numel()
        kpad = (k + 7) // 8 * 8
        flat = torch.zeros((d, kpad), device=wpe.device)
        flat[:, :k] = wpe.reshape(d, k)
        self.patch = pk(flat)
        # heads whose width is not a multiple of 16 (DiT-XL: 72) are zero-padded INSIDE the packed qkv / proj weights
        h, dh = m.num_heads, m.hidden_size // m.num_heads
        self.dp = _dpad(dh)

        def qkv_w(b):
            return torch.cat([_pad_heads_out(part, h, dh) for part in _f(b.attn.qkv.weight).chunk(3, 0)], 0)

        def qkv_b(b):
            return torch.cat([_pad_heads_out(part[:, None], h, dh)[:, 0] for part in _f(b.attn.qkv.bias).chunk(3, 0)], 0).contiguous()

        self.blocks = [dict(qkv=pk(qkv_w(b)), qkv_b=qkv_b(b), proj=pk(_pad_heads_in(_f(b.attn.proj.weight), h, dh)),
                            fc1=pk(b.mlp.fc1.weight), fc2=pk(b.mlp.fc2.weight)) for b in m.blocks]
        # final Linear: Cout = p*p*c may be tiny -> zero-pad the rows to a multiple of 16 (one N tile)
        wf, bf = _f(m.final_layer.linear.weight), _f(m.final_layer.linear.bias)
        self.final_cols = wf.shape[0]
        cp = (wf.shape[0] + 15) // 16 * 16
        wfp = torch.zeros((cp, wf.shape[1]), device=wf.device)
        wfp[:wf.shape[0]] = wf
        self.final_b = torch.zeros(cp, device=wf.device)
        self.final_b[:wf.shape[0]] = bf
        self.final = pk(wfp)
        # all adaLN projections as one [sum, d] fp32 matrix: 6d per block, then 2d for the final layer
        self.mod_w = torch.cat([_f(b.adaLN_modulation[1].weight) for b in m.blocks]
                               + [_f(m.final_layer.adaLN_modulation[1].weight)], 0).contiguous()
        self.mod_b = torch.cat([_f(b.adaLN_modulation[1].bias) for b in m.blocks]
                               + [_f(m.final_layer.adaLN_modulation[1].bias)], 0).contiguous()
        self.pos16 = m.pos_embed.detach().to(dtype).contiguous()      # [1, T, d]
        self.pos_cache = {}
        half = m.t_embedder.frequency_embedding_size // 2
        # the reference's own fp32 expression (DiT_models.py:52-54), evaluated on the CPU once
        self.freqs = torch.exp(-math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(wpe.device)

    def pos(self, n: int) -> Tensor:
        if n not in self.pos_cache:
            self.pos_cache[n] = self.pos16.expand(n, -1, -1).contiguous()
        return self.pos_cache[n]


class DiT(nn.Module):
    """Diffusion model with a Transformer backbone (DiT_models.py:145-272)."""
    _b200_engine = True       # the samplers' step engine drives this network (samplers._engine_model_fn)

    def __init__(self, input_size=32, patch_size=2, in_channels=4, hidden_size=1152, depth=28, num_heads=16,
                 mlp_ratio=4.0, class_dropout_prob=0.1, num_classes=1000, learn_sigma=True,
                 act_dtype: Optional[torch.dtype] = None):
        super().__init__()
        self.learn_sigma = learn_sigma
        self.in_channels = in_channels
        self.out_channels = in_channels // 3 * 2 if learn_sigma else in_channels      # (sic) DiT_models.py:186
        self.patch_size = patch_size
        self.num_heads = num_heads
        self.hidden_size = hidden_size
        self.act_dtype = act_dtype or default_act_dtype()
        if hidden_size % 8 or (hidden_size // num_heads) % 8:
            raise B200Error("B200 engine: DiT hidden_size and head width must be multiples of 8")
        self.x_embedder = PatchEmbed(input_size, patch_size, in_channels, hidden_size, bias=True)
        self.t_embedder = TimestepEmbedder(hidden_size)
        self.y_embedder = LabelEmbedder(num_classes, hidden_size, class_dropout_prob)
        num_patches = self.x_embedder.num_patches
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, hidden_size), requires_grad=False)
        self.blocks = nn.ModuleList([DiTBlock(hidden_size, num_heads, mlp_ratio=mlp_ratio) for _ in range(depth)])
        self.final_layer = FinalLayer(hidden_size, patch_size, self.out_channels)
        self.initialize_weights()
        self._pw: Optional[_Packed] = None
        self.register_load_state_dict_post_hook(lambda mod, keys: mod.invalidate())

    def initialize_weights(self):
        """DiT_models.py:203-237 (xavier Linears, fixed sin-cos pos_embed, adaLN-Zero)."""
        def _basic_init(module):
            if isinstance(module, nn.Linear):
                torch.nn.init.xavier_uniform_(module.weight)
                if module.bias is not None:
                    nn.init.constant_(module.bias, 0)
        self.apply(_basic_init)
        pos = get_2d_sincos_pos_embed(self.pos_embed.shape[-1], int(self.x_embedder.num_patches ** 0.5))
        self.pos_embed.data.copy_(torch.from_numpy(pos).float().unsqueeze(0))
        w = self.x_embedder.proj.weight.data
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        nn.init.constant_(self.x_embedder.proj.bias, 0)
        nn.init.normal_(self.y_embedder.embedding_table.weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
        for block in self.blocks:
            nn.init.constant_(block.adaLN_modulation[-1].weight, 0)
            nn.init.constant_(block.adaLN_modulation[-1].bias, 0)
        nn.init.constant_(self.final_layer.adaLN_modulation[-1].weight, 0)
        nn.init.constant_(self.final_layer.adaLN_modulation[-1].bias, 0)
        nn.init.constant_(self.final_layer.linear.weight, 0)
        nn.init.constant_(self.final_layer.linear.bias, 0)

    # -- engine ------------------------------------------------------------------------------
    def invalidate(self):
        self._pw = None

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def packed(self) -> _Packed:
        dev = self.pos_embed.device
        if dev.type != "cuda":
            raise B200Error("the B200 engine has no CPU path: move the module to a CUDA device")
        fp = param_fingerprint(self)
        if self._pw is not None and fp != getattr(self, "_pw_fp", None):      # in-place weight update since the last pack
            self.invalidate()
        if self._pw is None or self._pw.dtype != self.act_dtype:
            with torch.no_grad():
                self._pw = _Packed(self, self.act_dtype)
            self._pw_fp = fp
        return self._pw

    @torch.no_grad()
    def forward(self, x, t, y=None, cond=None, step_ptr: Optional[Tensor] = None, force_drop_ids=None):
        """x [N, C, H, W] fp32, t [N] (int or float) diffusion timesteps, y [N] class labels (optional),
        cond [N, C', H, W] concatenated on the channel axis (DiT_models.py:234-254) -> [N, out_channels, H, W].

        ``step_ptr`` (engine extension): device int32 index into ``t`` used as a table of per-step timesteps, so a
        captured CUDA graph can be replayed for every sampling step."""
        if cond is not None:
            x = torch.cat([x, cond], dim=1)
        pw = self.packed()
        n, _, hh, ww = x.shape
        p, d = self.patch_size, self.hidden_size
        gh, gw = hh // p, ww // p
        if gh * gw != self.x_embedder.num_patches:
            raise B200Error(f"DiT: input {hh}x{ww} does not match input_size {self.x_embedder.img_size}")
        dt = self.act_dtype
        # c = t_embedder(t) (+ y_embedder(y)); then every adaLN projection of the network in one GEMV
        if step_ptr is None:
            t_table = t.reshape(-1).to(torch.float32)
            if t_table.numel() == 1 and n > 1:
                t_table = t_table.expand(n)
            t_table = t_table.contiguous()
        else:
            t_table = t.to(torch.float32).contiguous()
        te = self.t_embedder
        t_freq = ops.timestep_embedding(t_table, pw.freqs, n, te.frequency_embedding_size, step_ptr)
        hid = ops.linear_f32(t_freq, _f(te.mlp[0].weight), _f(te.mlp[0].bias), out_silu=True)
        c = ops.linear_f32(hid, _f(te.mlp[2].weight), _f(te.mlp[2].bias))
        if y is not None:
            labels = y
            if force_drop_ids is not None:                      # LabelEmbedder.token_drop, DiT_models.py:78-87
                labels = torch.where(force_drop_ids == 1, self.y_embedder.num_classes, labels)
            c = ops.add_f32(c, _f(self.y_embedder.embedding_table.weight)[labels].contiguous())   # row gather: no arithmetic
        mod = ops.linear_f32(c, pw.mod_w, pw.mod_b, in_silu=True)                                    # [N, depth*6d + 2d]

        tok = ops.patchify(x.float(), p, dt)
        h = linear_tokens(tok, pw.patch, bias=_f(self.x_embedder.proj.bias), residual=pw.pos(n))
        heads, inner = self.num_heads, self.num_heads * pw.dp
        for i, blk in enumerate(self.blocks):
            w = pw.blocks[i]
            m = mod[:, i * 6 * d:(i + 1) * 6 * d]
            sl = lambda j: m[:, j * d:(j + 1) * d]
            a = ops.layer_norm(h, None, None, 1e-6, shift=sl(0), scale=sl(1))
            qkv = linear_tokens(a, w["qkv"], bias=w["qkv_b"])
            att = ops.attention(qkv, qkv, qkv, heads, pw.dp, blk.attn.scale, q_off=0, k_off=inner, v_off=2 * inner)
            h = linear_tokens(att, w["proj"], bias=_f(blk.attn.proj.bias), residual=h, rowscale=sl(2))
            a = ops.layer_norm(h, None, None, 1e-6, shift=sl(3), scale=sl(4))
            u = linear_tokens(a, w["fc1"], bias=_f(blk.mlp.fc1.bias), act=1)
            h = linear_tokens(u, w["fc2"], bias=_f(blk.mlp.fc2.bias), residual=h, rowscale=sl(5))
        fm = mod[:, len(self.blocks) * 6 * d:]
        a = ops.layer_norm(h, None, None, 1e-6, shift=fm[:, :d], scale=fm[:, d:])
        out = linear_tokens(a, pw.final, bias=pw.final_b, out_f32=True)                  # fp32 [N, T, >= p*p*c]
        res = ops.unpatchify(out, self.out_channels, gh, gw, p).to(x.dtype)
        if step_ptr is None and not torch.cuda.is_current_stream_capturing():      # fp16 range guard (ops.check_range)
            ops.check_range(x.device, "DiT.forward")
        return res

    def forward_with_cfg(self, x, t, y, cfg_scale):
        """DiT_models.py:256-272: classifier-free guidance on the first three channels (sic)."""
        half = x[: len(x) // 2]
        combined = torch.cat([half, half], dim=0)
        model_out = self.forward(combined, t, y)
        eps, rest = model_out[:, :3], model_out[:, 3:]
        cond_eps, uncond_eps = torch.split(eps, len(eps) // 2, dim=0)
        half_eps = uncond_eps + cfg_scale * (cond_eps - uncond_eps)      # host-side glue on [N, 3, H, W]; not a hot path
        eps = torch.cat([half_eps, half_eps], dim=0)
        return torch.cat([eps, rest], dim=1)


def DiT_XL_2(**kw): return DiT(depth=28, hidden_size=1152, patch_size=2, num_heads=16, **kw)
def DiT_XL_4(**kw): return DiT(depth=28, hidden_size=1152, patch_size=4, num_heads=16, **kw)
def DiT_XL_8(**kw): return DiT(depth=28, hidden_size=1152, patch_size=8, num_heads=16, **kw)
def DiT_L_2(**kw): return DiT(depth=24, hidden_size=1024, patch_size=2, num_heads=16, **kw)
def DiT_L_4(**kw): return DiT(depth=24, hidden_size=1024, patch_size=4, num_heads=16, **kw)
def DiT_L_8(**kw): return DiT(depth=24, hidden_size=1024, patch_size=8, num_heads=16, **kw)
def DiT_B_2(**kw): return DiT(depth=12, hidden_size=768, patch_size=2, num_heads=12, **kw)
def DiT_B_4(**kw): return DiT(depth=12, hidden_size=768, patch_size=4, num_heads=12, **kw)
def DiT_B_8(**kw): return DiT(depth=12, hidden_size=768, patch_size=8, num_heads=12, **kw)
def DiT_S_2(**kw): return DiT(depth=12, hidden_size=384, patch_size=2, num_heads=6, **kw)
def DiT_S_4(**kw): return DiT(depth=12, hidden_size=384, patch_size=4, num_heads=6, **kw)
def DiT_S_8(**kw): return DiT(depth=12, hidden_size=384, patch_size=8, num_heads=6, **kw)


DiT_models = {
    'DiT-XL/2': DiT_XL_2, 'DiT-XL/4': DiT_XL_4, 'DiT-XL/8': DiT_XL_8,
    'DiT-L/2': DiT_L_2, 'DiT-L/4': DiT_L_4, 'DiT-L/8': DiT_L_8,
    'DiT-B/2': DiT_B_2, 'DiT-B/4': DiT_B_4, 'DiT-B/8': DiT_B_8,
    'DiT-S/2': DiT_S_2, 'DiT-S/4': DiT_S_4, 'DiT-S/8': DiT_S_8,
}
