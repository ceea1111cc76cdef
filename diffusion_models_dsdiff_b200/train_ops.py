"""Tensor-level wrappers over the training-step entry points of the C ABI (include/b200diff.h, csrc/train.cu).

Same rules as ``ops``: torch is device memory + streams; every arithmetic op is a kernel of libb200diff.so on torch's
current stream; no CPU / eager fallback.  The three GEMM flavours of a linear layer all run on the tcgen05 implicit-GEMM
kernel (``ops.conv2d``): forward ``y = x W^T``, dgrad ``dx = dy W`` (a linear with the transposed packed weight) and wgrad
``dW = dy^T x`` (both token-major operands transposed to K-major, fp32 result, forced split-K over the token axis)."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import _lib as L
from . import ops
from .ops import _dev, _f32, _p, _stream, _timed, act_code

Tensor = torch.Tensor
_NO_MN_WGRAD = os.environ.get("B200DIFF_WGRAD_TRANSPOSE", "0") == "1"


def transpose16(x: Tensor, colsum_out: Optional[Tensor] = None) -> Tensor:
    """[R, C] 16-bit -> [C, R]; ``colsum_out`` (fp32 [C], overwritten): the column sums of ``x`` from the same pass."""
    _dev(x)
    r, c = x.shape
    out = torch.empty((c, r), device=x.device, dtype=x.dtype)
    if colsum_out is not None:
        colsum_out.zero_()
    _timed("transpose16", 0.0, 4.0 * x.numel(),
           lambda: L.check(L.lib().b200_transpose16(_p(x), _p(out), _p(_f32(colsum_out)), r, c, act_code(x.dtype), _stream()),
                           "transpose16"))
    return out


def colsum(x: Tensor, out: Tensor) -> Tensor:
    """out[c] = sum over rows of x [R, C] (bias gradient); ``out`` fp32 [C] is overwritten."""
    _dev(x)
    out.zero_()
    r, c = x.shape
    _timed("colsum16", 0.0, 2.0 * x.numel(),
           lambda: L.check(L.lib().b200_colsum16(_p(x), _p(_f32(out)), r, c, act_code(x.dtype), _stream()), "colsum16"))
    return out


def gelu_fwd(pre: Tensor) -> Tensor:
    _dev(pre)
    out = torch.empty_like(pre)
    _timed("gelu", 0.0, 4.0 * pre.numel(),
           lambda: L.check(L.lib().b200_gelu_fwd(_p(pre), _p(out), pre.numel(), act_code(pre.dtype), _stream()), "gelu_fwd"))
    return out


def gelu_bwd(pre: Tensor, du: Tensor) -> Tensor:
    _dev(pre)
    out = torch.empty_like(pre)
    _timed("gelu_bwd", 0.0, 6.0 * pre.numel(),
           lambda: L.check(L.lib().b200_gelu_bwd(_p(pre), _p(du), _p(out), pre.numel(), act_code(pre.dtype), _stream()),
                           "gelu_bwd"))
    return out


def _rows(t: Tensor, name: str) -> Tuple[int, int]:
    """fp32 [N, C] row view (a column slice of a wider matrix is fine): (data_ptr, pitch)."""
    if t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1 or t.data_ptr() % 16:
        raise L.B200Error(f"{name} must be fp32 [N, C] rows with unit channel stride, 16-byte aligned")
    return t.data_ptr(), t.stride(0)


def gated_residual(h_in: Tensor, f: Tensor, gate: Tensor) -> Tensor:
    """h_in + gate[n] * f over tokens [N, L, C]."""
    _dev(h_in)
    n, l, c = h_in.shape
    gp, gs = _rows(gate, "gate")
    out = torch.empty_like(h_in)
    _timed("gated_residual", 0.0, 6.0 * h_in.numel(),
           lambda: L.check(L.lib().b200_gated_residual(_p(h_in), _p(f), gp, gs, _p(out), n * l, l, c, act_code(h_in.dtype),
                                                       _stream()), "gated_residual"))
    return out


def gate_bwd(dh: Tensor, f: Tensor, gate: Tensor, dgate: Tensor) -> Tensor:
    """df = gate[n] * dh ; dgate[n, c] += sum_t dh * f   (dgate: zeroed fp32 [N, C] rows)."""
    _dev(dh)
    n, l, c = dh.shape
    gp, gs = _rows(gate, "gate")
    dp, ds = _rows(dgate, "dgate")
    out = torch.empty_like(dh)
    _timed("gate_bwd", 0.0, 6.0 * dh.numel(),
           lambda: L.check(L.lib().b200_gate_bwd(_p(dh), _p(f), gp, gs, _p(out), dp, ds, n, l, c, act_code(dh.dtype),
                                                 _stream()), "gate_bwd"))
    return out


def ln_mod_bwd(x: Tensor, da: Tensor, scale: Tensor, dres: Optional[Tensor], dshift: Tensor, dscale: Tensor,
               eps: float = 1e-6) -> Tensor:
    """Backward of LayerNorm(x) * (1 + scale[n]) + shift[n]: returns dx (+ dres); accumulates dshift / dscale rows."""
    _dev(x)
    n, l, c = x.shape
    sp, ss = _rows(scale, "scale")
    hp, hs = _rows(dshift, "dshift")
    cp, cs = _rows(dscale, "dscale")
    if hs != cs:
        raise L.B200Error("ln_mod_bwd: dshift / dscale must share one row pitch")
    out = torch.empty_like(x)
    _timed("ln_mod_bwd", 0.0, 2.0 * x.numel() * (3 + (dres is not None)),
           lambda: L.check(L.lib().b200_ln_mod_bwd(_p(x), _p(da), sp, ss, _p(dres), _p(out), hp, cp, hs, n, l, c, eps,
                                                   act_code(x.dtype), _stream()), "ln_mod_bwd"))
    return out


def attention_bwd(qkv: Tensor, att: Tensor, datt: Tensor, lse: Tensor, heads: int, d: int, scale: float) -> Tensor:
    """Self-attention backward over the packed projection qkv [B, L, 3*heads*d] (q | k | v blocks): returns dqkv."""
    _dev(qkv)
    b, l, ld = qkv.shape
    inner = heads * d
    dqkv = torch.empty_like(qkv)
    dsum = torch.empty((b, l, heads), device=qkv.device, dtype=torch.float32)
    dq = torch.empty((b, l, inner), device=qkv.device, dtype=torch.float32)
    _timed("attention_bwd", 10.0 * b * heads * l * l * d, 2.0 * (2 * qkv.numel() + 2 * att.numel()),
           lambda: L.check(L.lib().b200_attention_bwd(_p(qkv), ld, 0, inner, 2 * inner, _p(att), _p(datt), _p(_f32(lse)), _p(dsum),
                                                      _p(dq), _p(dqkv), b, l, heads, d, scale, act_code(qkv.dtype), _stream()),
                           "attention_bwd"), detail=f"B{b} L{l} h{heads} d{d}")
    return dqkv


def unpatchify_bwd(dimg: Tensor, p: int, ld: int, dtype: torch.dtype) -> Tensor:
    _dev(dimg)
    n, c, h, w = dimg.shape
    gh, gw = h // p, w // p
    out = torch.empty((n, gh * gw, ld), device=dimg.device, dtype=dtype)
    L.check(L.lib().b200_unpatchify_bwd(_p(_f32(dimg.contiguous())), _p(out), n, c, gh, gw, p, ld, act_code(dtype), _stream()),
            "unpatchify_bwd")
    return out


def q_sample(x0: Tensor, noise: Tensor, coef: Tensor) -> Tensor:
    _dev(x0)
    n = x0.shape[0]
    out = torch.empty_like(x0)
    L.check(L.lib().b200_q_sample(_p(_f32(x0.contiguous())), _p(_f32(noise.contiguous())), _p(_f32(coef)), _p(out), n,
                                  x0.numel() // n, _stream()), "q_sample")
    return out


def diffusion_loss(model_out: Tensor, x_start: Tensor, x_t: Tensor, noise: Tensor, coef: Tensor, param: int,
                   learn_sigma: bool) -> Tuple[Tensor, Tensor]:
    """(terms [N, 2] = (mse, vb) per sample, d_out = gradient of mean_n(mse + vb) w.r.t. model_out)."""
    _dev(model_out)
    n, c = x_start.shape[0], x_start.shape[1]
    hw = x_start.numel() // (n * c)
    if model_out.shape[1] != (2 * c if learn_sigma else c):
        raise L.B200Error("diffusion_loss: model_out channel count does not match learn_sigma")
    terms = torch.empty((n, 2), device=model_out.device, dtype=torch.float32)
    d_out = torch.empty_like(model_out)
    _timed("diffusion_loss", 0.0, 4.0 * (2 * model_out.numel() + 3 * x_start.numel()),
           lambda: L.check(L.lib().b200_diffusion_loss(_p(_f32(model_out.contiguous())), _p(_f32(x_start.contiguous())),
                                                       _p(_f32(x_t.contiguous())), _p(_f32(noise.contiguous())), _p(_f32(coef)),
                                                       param, int(learn_sigma), _p(terms), _p(d_out), n, c, hw, _stream()),
                           "diffusion_loss"))
    return terms, d_out


def linear_f32_wgrad(dy: Tensor, x: Tensor, dW: Tensor, db: Optional[Tensor], in_silu: bool = False,
                     accumulate: bool = False) -> None:
    _dev(dy)
    yp, ys = _rows(dy, "dy")
    if x.dtype != torch.float32 or x.stride(1) != 1:
        raise L.B200Error("linear_f32_wgrad: x must be fp32 rows")
    n, o = dy.shape
    i = x.shape[1]
    if tuple(dW.shape) != (o, i) or not dW.is_contiguous():
        raise L.B200Error("linear_f32_wgrad: dW must be a contiguous [O, I] tensor")
    L.check(L.lib().b200_linear_f32_wgrad(yp, ys, x.data_ptr(), x.stride(0), int(in_silu), _p(_f32(dW)), _p(db), n, o, i,
                                          int(accumulate), _stream()), "linear_f32_wgrad")


def linear_f32_dgrad(dy: Tensor, w: Tensor, x_pre: Optional[Tensor] = None) -> Tensor:
    """dx = dy W  (* silu'(x_pre) when the linear's input was silu(x_pre))."""
    _dev(dy)
    yp, ys = _rows(dy, "dy")
    n, o = dy.shape
    i = w.shape[1]
    dx = torch.empty((n, i), device=dy.device, dtype=torch.float32)
    L.check(L.lib().b200_linear_f32_dgrad(yp, ys, _p(_f32(w)), _p(x_pre), _p(dx), n, o, i, _stream()), "linear_f32_dgrad")
    return dx


def adamw(p: Tensor, g: Tensor, m: Tensor, v: Tensor, lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
          weight_decay: float = 1e-2, step: int = 1, grad_scale: float = 1.0, hyper: Optional[Tensor] = None) -> None:
    """``hyper``: device fp32 [3] = (lr, 1 - beta1^step, sqrt(1 - beta2^step)) read instead of ``lr`` / ``step`` (graph replay)."""
    _dev(p)
    _timed("adamw", 0.0, 4.0 * 7 * p.numel(),
           lambda: L.check(L.lib().b200_adamw(_p(_f32(p)), _p(_f32(g)), _p(_f32(m)), _p(_f32(v)), p.numel(), lr, betas[0],
                                              betas[1], eps, weight_decay, step, grad_scale, _p(hyper), _stream()), "adamw"))


def adamw_hyper(lr: float, betas, step: int) -> Tuple[float, float, float]:
    return lr, 1.0 - betas[0] ** step, (1.0 - betas[1] ** step) ** 0.5


# ---------------------------------------------------------------------------------------------
# the three GEMMs of a linear layer on the tensor cores
# ---------------------------------------------------------------------------------------------
def _img(t2d: Tensor) -> Tensor:
    """[R, C] -> an NHWC view [1, R/w, w, C] for the implicit-GEMM kernel (wide rows tile best)."""
    r, c = t2d.shape
    w = 32 if r % 32 == 0 else (16 if r % 16 == 0 else (8 if r % 8 == 0 else r))
    return t2d.view(1, r // w, w, c)


def wgrad_split(o: int, i: int, t: int) -> int:
    """K-split of the weight-gradient GEMM dW[O, I] = dY^T X over T tokens: enough (tile, split) items for every SM pair,
    at least 8 K-chunks of 64 tokens per split.  Geometry only."""
    tiles = -(-o // 256) * -(-i // 256)
    s = max(1, min(16, round(74 / tiles)))
    return max(1, min(s, t // 512))


def linear_wgrad(dy: Tensor, x: Tensor, db: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
    """dW [O, I] fp32 = dy^T x for token-major dy [T, O], x [T, I] (16-bit).  T, O, I multiples of 8.
    ``db`` (fp32 [O], overwritten): the bias gradient sum_t dy[t].

    O and I multiples of 64 (every Linear of a DiT block): the MN-major tcgen05 GEMM reads dy and x as they are
    (b200_linear_wgrad) and the bias gradient is one column-sum pass.  Otherwise (patch embedding, final layer: 16-wide
    operands) both operands are transposed to K-major — the bias gradient comes out of dy's transpose pass — and the
    implicit-GEMM kernel runs with a forced K split."""
    t, o = dy.shape
    i = x.shape[1]
    if o % 64 == 0 and i % 64 == 0 and not _NO_MN_WGRAD:
        _dev(dy)
        if out is None:
            out = torch.empty((o, i), device=dy.device, dtype=torch.float32)
        elif out.dtype != torch.float32 or out.numel() != o * i or not out.is_contiguous():
            raise L.B200Error("linear_wgrad: out must be a contiguous fp32 [O, I] tensor")
        splits = int(L.lib().b200_wgrad_splits(t, o, i))
        ws = ops._workspace(dy.device, splits * o * i * 4) if splits > 1 else None
        _timed("wgrad_tn", 2.0 * t * o * i, 2.0 * (dy.numel() + x.numel()) + 4.0 * o * i,
               lambda: L.check(L.lib().b200_linear_wgrad(_p(dy), _p(x), _p(out), _p(ws), t, o, i, act_code(dy.dtype), _stream()),
                               "linear_wgrad"), detail=f"T{t} {o}x{i} S{splits}")
        if db is not None:
            colsum(dy, db)
        return out.view(o, i)
    dyt, xt = transpose16(dy, db), transpose16(x)               # K-major rows: [O, T], [I, T]
    res = ops.conv2d([(_img(dyt), xt.view(1, i, t), 0, 1, 0)], out_f32=True, force_ksplit=wgrad_split(o, i, t), out=out)
    return res.view(o, i)
