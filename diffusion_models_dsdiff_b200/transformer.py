"""SpatialTransformer / BasicTransformerBlock / CrossAttention / GEGLU feed-forward (ldm/modules/attention.py:49-76,
145-194, 302-331, 366-428) on libb200diff — parameter containers with the reference's state_dict names plus the engine
that runs them.

Every Linear / 1x1 conv is the tcgen05 GEMM (``ops.conv2d`` over the token tensor viewed as NHWC) with bias, residual
add and head re-layout folded into packing / epilogues; LayerNorm and GEGLU are single fused passes; attention is the
tcgen05 flash kernel (``ops.attention``) for self- and cross-attention alike (Lk = context length).  Head widths that
are not a multiple of 16 (SD-v1's 40) are zero-padded to the next multiple INSIDE the packed projection weights, so
q/k/v come out of the GEMM already in a tensor-core friendly layout and the padding never touches HBM activations'
real channels (zero columns of to_out undo it).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import B200Error

Tensor = torch.Tensor


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.):
        super().__init__()
        inner_dim = int(dim * mult)
        dim_out = dim_out or dim
        project_in = nn.Sequential(nn.Linear(dim, inner_dim), nn.GELU()) if not glu else GEGLU(dim, inner_dim)
        self.net = nn.Sequential(project_in, nn.Dropout(dropout), nn.Linear(inner_dim, dim_out))


class CrossAttention(nn.Module):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner_dim = dim_head * heads
        context_dim = context_dim or query_dim
        self.scale = dim_head ** -0.5
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, query_dim), nn.Dropout(dropout))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, n_heads, d_head, dropout=0., context_dim=None, gated_ff=True, checkpoint=True,
                 disable_self_attn=False):
        super().__init__()
        self.disable_self_attn = disable_self_attn
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout,
                                    context_dim=context_dim if disable_self_attn else None)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head,
                                    dropout=dropout)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)


class SpatialTransformer(nn.Module):
    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0., context_dim=None, disable_self_attn=False,
                 use_linear=False, use_checkpoint=True):
        super().__init__()
        if context_dim is not None and not isinstance(context_dim, (list, tuple)):
            context_dim = [context_dim] * depth
        if context_dim is None:
            context_dim = [None] * depth
        self.in_channels = in_channels
        inner_dim = n_heads * d_head
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner_dim) if use_linear else nn.Conv2d(in_channels, inner_dim, 1)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=context_dim[d],
                                  disable_self_attn=disable_self_attn) for d in range(depth)])
        # (sic) the reference builds the linear proj_out as Linear(in_channels, inner_dim): attention.py:406
        self.proj_out = nn.Linear(in_channels, inner_dim) if use_linear else nn.Conv2d(inner_dim, in_channels, 1)
        self.use_linear = use_linear


# ---------------------------------------------------------------------------------------------
# engine
# ---------------------------------------------------------------------------------------------
def _dpad(d: int) -> int:
    return (d + 15) // 16 * 16


def _pad_heads_out(w: Tensor, heads: int, d: int) -> Tensor:
    """[heads*d, in] -> [heads*dpad, in] with zero rows: the projection emits zero-padded heads."""
    dp = _dpad(d)
    if dp == d:
        return w
    out = torch.zeros((heads, dp, w.shape[1]), device=w.device, dtype=w.dtype)
    out[:, :d] = w.view(heads, d, -1)
    return out.view(heads * dp, -1)


def _pad_heads_in(w: Tensor, heads: int, d: int) -> Tensor:
    """[out, heads*d] -> [out, heads*dpad] with zero columns: consumes zero-padded heads."""
    dp = _dpad(d)
    if dp == d:
        return w
    out = torch.zeros((w.shape[0], heads, dp), device=w.device, dtype=w.dtype)
    out[:, :, :d] = w.view(w.shape[0], heads, d)
    return out.view(w.shape[0], heads * dp)


def _packed(pw, key, build):
    if key not in pw.conv:
        with torch.no_grad():
            pw.conv[key] = build()
    return pw.conv[key]


def _f(p: Tensor) -> Tensor:
    t = p.detach()
    return t if t.dtype == torch.float32 and t.is_contiguous() else t.float().contiguous()


def _tok_img(t: Tensor) -> Tensor:
    """tokens [B, L, C] -> an NHWC view [B, H, W, C] for the GEMM (any H*W = L; wide rows tile best)."""
    b, l, c = t.shape
    w = 32 if l % 32 == 0 else (16 if l % 16 == 0 else (8 if l % 8 == 0 else l))
    return t.view(b, l // w, w, c)


def linear_tokens(t: Tensor, wp: Tensor, bias: Optional[Tensor] = None, residual: Optional[Tensor] = None, **kw) -> Tensor:
    """y = t W^T (+ bias) (+ residual) on the tensor cores; t [B, L, Cin] 16-bit, wp packed [1, Cout, Cin]."""
    b, l, _ = t.shape
    res = None if residual is None else _tok_img(residual)
    y = ops.conv2d([(_tok_img(t), wp, 0, 1, 0)], bias=bias, residual=res, **kw)
    return y.view(b, l, wp.shape[1])


def run_cross_attention(pw, m: CrossAttention, x: Tensor, context: Optional[Tensor], residual: Tensor) -> Tensor:
    """CrossAttention.forward(x, context) + residual (attention.py:163-194); context None = self-attention."""
    dt = x.dtype
    h, d = m.heads, m.dim_head
    dp = _dpad(d)
    inner_p = h * dp
    if context is None:
        wqkv = _packed(pw, ("xa_qkv", id(m)), lambda: ops.pack_conv_weight(
            torch.cat([_pad_heads_out(_f(w), h, d) for w in (m.to_q.weight, m.to_k.weight, m.to_v.weight)], 0), dt))
        qkv = linear_tokens(x, wqkv)
        att = ops.attention(qkv, qkv, qkv, h, dp, m.scale, q_off=0, k_off=inner_p, v_off=2 * inner_p)
    else:
        wq = _packed(pw, ("xa_q", id(m)), lambda: ops.pack_conv_weight(_pad_heads_out(_f(m.to_q.weight), h, d), dt))
        wkv = _packed(pw, ("xa_kv", id(m)), lambda: ops.pack_conv_weight(
            torch.cat([_pad_heads_out(_f(w), h, d) for w in (m.to_k.weight, m.to_v.weight)], 0), dt))
        q = linear_tokens(x, wq)
        kv = linear_tokens(context, wkv)
        att = ops.attention(q, kv, kv, h, dp, m.scale, q_off=0, k_off=0, v_off=inner_p)
    lin = m.to_out[0]
    wo = _packed(pw, ("xa_o", id(m)), lambda: ops.pack_conv_weight(_pad_heads_in(_f(lin.weight), h, d), dt))
    return linear_tokens(att, wo, bias=_f(lin.bias), residual=residual)


def run_basic_block(pw, m: BasicTransformerBlock, t: Tensor, context: Optional[Tensor]) -> Tensor:
    """BasicTransformerBlock._forward (attention.py:327-331)."""
    n1, n2, n3 = m.norm1, m.norm2, m.norm3
    t = run_cross_attention(pw, m.attn1, ops.layer_norm(t, _f(n1.weight), _f(n1.bias), n1.eps),
                            context if m.disable_self_attn else None, t)
    t = run_cross_attention(pw, m.attn2, ops.layer_norm(t, _f(n2.weight), _f(n2.bias), n2.eps), context, t)
    proj_in, lin_out = m.ff.net[0], m.ff.net[2]
    if not isinstance(proj_in, GEGLU):
        raise NotImplementedError("B200 engine: FeedForward(glu=False) is not built")
    dt = t.dtype
    w1 = _packed(pw, ("ff1", id(m)), lambda: ops.pack_conv_weight(_f(proj_in.proj.weight), dt))
    w2 = _packed(pw, ("ff2", id(m)), lambda: ops.pack_conv_weight(_f(lin_out.weight), dt))
    hid = ops.geglu(linear_tokens(ops.layer_norm(t, _f(n3.weight), _f(n3.bias), n3.eps), w1, bias=_f(proj_in.proj.bias)))
    return linear_tokens(hid, w2, bias=_f(lin_out.bias), residual=t)


def context_tokens(context, dtype: torch.dtype) -> Optional[Tensor]:
    """fp32 [B, Lk, D] conditioning -> 16-bit tokens (once per forward)."""
    if context is None:
        return None
    if isinstance(context, (list, tuple)):
        if len(context) != 1:
            raise NotImplementedError("B200 engine: one context tensor per forward")
        context = context[0]
    b, lk, dctx = context.shape
    if dctx % 8:
        raise B200Error("B200 engine: context_dim must be a multiple of 8")
    return ops.nchw_to_nhwc(context.float().reshape(b * lk, dctx, 1, 1).contiguous(), dctx, dtype).view(b, lk, dctx)
