"""Tensor-level wrappers over the C ABI (include/b200diff.h).

torch is used for device memory and streams only; every arithmetic op below is a kernel of
libb200diff.so launched on torch's current stream.  Internal activations are NHWC 16-bit tensors
(``[N, H, W, C]``, ``torch.float16`` by default or ``torch.bfloat16``); the API edge is NCHW fp32
like the reference.  Nothing here has a CPU or eager fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import torch

from . import _lib as L

Tensor = torch.Tensor


class KernelTimer:
    """Optional per-launch CUDA-event timing (bench.py's roofline pass).  Events are recorded on the launching
    stream around each kernel of libb200diff; nothing is recorded unless a timer is installed."""

    def __init__(self):
        self.records = []      # (name, algorithmic flops, algorithmic bytes, start event, end event, detail)

    def summary(self) -> dict:
        torch.cuda.synchronize()
        out = {}
        for name, fl, by, a, b, _ in self.records:
            d = out.setdefault(name, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += a.elapsed_time(b)
            d["flops"] += fl
            d["bytes"] += by
        return out

    def by_detail(self) -> dict:
        """The same totals keyed by (kernel, shape string) — tools/step_breakdown.py."""
        torch.cuda.synchronize()
        out = {}
        for name, fl, by, a, b, det in self.records:
            d = out.setdefault((name, det or ""), {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += a.elapsed_time(b)
            d["flops"] += fl
            d["bytes"] += by
        return out


_timer: Optional[KernelTimer] = None


def set_timer(t: Optional[KernelTimer]) -> None:
    global _timer
    _timer = t


def _timed(name: str, flops: float, nbytes: float, launch, detail: Optional[str] = None):
    if _timer is None:
        return launch()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = launch()
    b.record()
    _timer.records.append((name, flops, nbytes, a, b, detail))
    return r


def act_code(dtype: torch.dtype) -> int:
    if dtype == torch.float16:
        return L.F16
    if dtype == torch.bfloat16:
        return L.BF16
    raise L.B200Error(f"activation dtype must be float16 or bfloat16, got {dtype}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dev(t: Tensor) -> None:
    if not t.is_cuda:
        raise L.B200Error("libb200diff has no CPU path: tensor is not on a CUDA device")
    cur = torch.cuda.current_device()
    idx = t.device.index if t.device.index is not None else cur
    if idx != cur:      # kernels launch on the CURRENT device's stream: one process per GPU (torch.cuda.set_device first)
        raise L.B200Error(f"tensor lives on cuda:{idx} but the current device is cuda:{cur}")
    L.ensure_init(idx)


def _p(t: Optional[Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_contiguous():
        raise L.B200Error("tensor must be contiguous")
    return t.data_ptr()


def _f32(t: Optional[Tensor]) -> Optional[Tensor]:
    if t is not None and t.dtype != torch.float32:
        raise L.B200Error(f"expected float32, got {t.dtype}")
    return t


# ---------------------------------------------------------------------------------------------
# layout / weights
# ---------------------------------------------------------------------------------------------
def nchw_to_nhwc(x: Tensor, cpad: Optional[int] = None, dtype=torch.float16) -> Tensor:
    _dev(x)
    n, c, h, w = x.shape
    cpad = cpad or ((c + 7) // 8) * 8
    out = torch.empty((n, h, w, cpad), device=x.device, dtype=dtype)
    L.check(L.lib().b200_nchw_f32_to_nhwc(_p(_f32(x.contiguous())), _p(out), n, c, h, w, cpad, act_code(dtype),
                                          _stream()), "nchw_f32_to_nhwc")
    return out


def nhwc_to_nchw(x: Tensor, c: Optional[int] = None) -> Tensor:
    _dev(x)
    n, h, w, cs = x.shape
    c = c or cs
    out = torch.empty((n, c, h, w), device=x.device, dtype=torch.float32)
    L.check(L.lib().b200_nhwc_to_nchw_f32(_p(x), _p(out), n, c, h, w, cs, act_code(x.dtype), _stream()),
            "nhwc_to_nchw_f32")
    return out


def pack_conv_weight(w_oihw: Tensor, dtype=torch.float16, cin_pad: Optional[int] = None) -> Tensor:
    """OIHW fp32 -> [KH*KW, Cout, CinPad] 16-bit (K-major rows for the tcgen05 B operand)."""
    _dev(w_oihw)
    if w_oihw.dim() == 3:      # Conv1d k=1 weights [O, I, 1]
        w_oihw = w_oihw[..., None]
    if w_oihw.dim() == 2:      # Linear weights [O, I]
        w_oihw = w_oihw[..., None, None]
    co, ci, kh, kw = w_oihw.shape
    cin_pad = cin_pad or ((ci + 7) // 8) * 8
    out = torch.empty((kh * kw, co, cin_pad), device=w_oihw.device, dtype=dtype)
    L.check(L.lib().b200_pack_conv_weight(_p(_f32(w_oihw.contiguous())), _p(out), co, ci, kh, kw, cin_pad,
                                          act_code(dtype), _stream()), "pack_conv_weight")
    return out


# ---------------------------------------------------------------------------------------------
# convolutions
# ---------------------------------------------------------------------------------------------
Seg = Tuple[Tensor, Tensor, int, int, int]  # (src NHWC, packed weight, coff, ksize, pad)


def conv2d(segs: Sequence[Seg], stride: int = 1, bias: Optional[Tensor] = None,
           rowbias: Optional[Tensor] = None, residual: Optional[Tensor] = None, res_scale: float = 1.0,
           out: Optional[Tensor] = None, stats: Optional[Tensor] = None, head_channels: int = 0,
           out_hw: Optional[Tuple[int, int]] = None, out_f32: bool = False, rowscale: Optional[Tensor] = None,
           act: int = 0, wsets: int = 1, rowbias_gstride: int = 0, force_ksplit: int = 0,
           out_sub: Optional[Tuple[int, int, int, int]] = None, out_lo=None,
           residual_lo: Optional[Tensor] = None):
    """tcgen05 implicit-GEMM conv over up to 3 K-segments (see b200_conv2d_igemm).

    ``wsets`` > 1: grouped along N — the batch is ``wsets`` equal groups of images, group g convolves with its own weight
    set (``weight`` is the stack [wsets * k * k, Cout, Cin]), ``bias`` is [wsets, Cout] and ``rowbias`` advances by
    ``rowbias_gstride`` floats per group.

    ``out_sub`` = (sy, sx, oy, ox): the result is one parity class of a [N, Hout*sy, Wout*sx, Cout] tensor ``out`` (required),
    stored at (h*sy + oy, w*sx + ox) — the sub-pixel form of Upsample -> conv3x3; a segment's ``pad`` may then be a
    (pad_top, pad_left) pair (2x2 kernels).

    ``out_lo``: True (allocate) or a pre-allocated tensor of ``out``'s shape / dtype — the LOW plane of the result
    (b200_conv_desc.out_lo: ``out + out_lo`` carries the fp32 value); the call then returns ``(out, out_lo)``.
    ``residual_lo``: low plane of ``residual``."""
    src0, w0, _, k0, p0 = segs[0]
    if isinstance(p0, tuple):
        p0 = p0[0]
    _dev(src0)
    n, hin, win, _ = src0.shape
    cout = w0.shape[1]
    hout = (hin + 2 * p0 - k0) // stride + 1
    wout = (win + 2 * p0 - k0) // stride + 1
    if out_hw is not None:     # asymmetric (right/bottom) zero padding: the KL-autoencoder Downsample, model.py:80-85
        hout, wout = out_hw
    d = L.ConvDesc()
    d.N, d.Hin, d.Win, d.Hout, d.Wout, d.Cout, d.stride, d.nseg = n, hin, win, hout, wout, cout, stride, len(segs)
    for i, (src, wt, coff, ks, pad) in enumerate(segs):
        pad_w1 = 0
        if isinstance(pad, tuple):
            pad, pad_w1 = pad[0], pad[1] + 1
        if src.shape[:3] != src0.shape[:3] or src.dtype != src0.dtype or wt.dtype != src0.dtype:
            raise L.B200Error("conv2d: segments must share N,H,W and dtype")
        if ((out_hw is None and (hin + 2 * pad - ks) // stride + 1 != hout) or wt.shape[1] != cout
                or wt.shape[0] != ks * ks * wsets):
            raise L.B200Error("conv2d: segment geometry mismatch")
        s = d.seg[i]
        s.src, s.weight, s.cin, s.wcin, s.coff, s.ksize, s.pad = _p(src), _p(wt), src.shape[3], wt.shape[2], coff, ks, pad
        s.pad_w1 = pad_w1
    if head_channels:      # network head: fp32 NCHW output of the first head_channels (zero-padded N tile)
        out = torch.empty((n, head_channels, hout, wout), device=src0.device, dtype=torch.float32)
        d.out_nchw_f32, d.cout_real = _p(out), head_channels
    elif out_f32:          # fp32 NHWC result (attention logits; weight gradients)
        if out is None:
            out = torch.empty((n, hout, wout, cout), device=src0.device, dtype=torch.float32)
        elif out.dtype != torch.float32 or out.numel() != n * hout * wout * cout:
            raise L.B200Error("conv2d: fp32 out shape/dtype mismatch")
        d.out_nhwc_f32 = _p(out)
    elif out is None:
        out = torch.empty((n, hout, wout, cout), device=src0.device, dtype=src0.dtype)
    elif out_sub is not None:
        if tuple(out.shape) != (n, hout * out_sub[0], wout * out_sub[1], cout) or out.dtype != src0.dtype:
            raise L.B200Error("conv2d: strided out shape/dtype mismatch")
    elif tuple(out.shape) != (n, hout, wout, cout) or out.dtype != src0.dtype:
        raise L.B200Error("conv2d: out shape/dtype mismatch")
    if out_sub is not None:
        if out is None or head_channels or out_f32 or residual is not None:
            raise L.B200Error("conv2d: out_sub needs a pre-allocated 16-bit NHWC out and no residual")
        d.out_sy, d.out_sx, d.out_oy, d.out_ox = out_sub
    d.bias = _p(_f32(bias))
    d.rowbias, d.rowbias_stride = None, 0
    d.wsets, d.rowbias_gstride, d.rowscale_gstride = wsets, rowbias_gstride, 0
    d.force_ksplit = int(force_ksplit)
    npg = n // max(wsets, 1)
    if wsets > 1 and (n % wsets or rowscale is not None or (bias is not None and bias.numel() != wsets * cout)):
        raise L.B200Error("conv2d: grouped launch needs N % wsets == 0, a [wsets, Cout] bias and no rowscale")
    if rowbias is not None:       # may be a column slice of the batched emb projection
        if (_f32(rowbias).stride(-1) != 1 or rowbias.shape[1] != cout or rowbias.shape[0] not in (1, npg)
                or rowbias.data_ptr() % 16 or (wsets > 1 and rowbias_gstride % 4)):
            raise L.B200Error("conv2d: rowbias must be [N or 1, Cout] fp32 with dense, 16B-aligned rows")
        d.rowbias = rowbias.data_ptr()
        d.rowbias_stride = rowbias.stride(0) if rowbias.shape[0] == npg and npg > 1 else 0   # [1, Cout]: shared by the batch
    d.rowscale, d.rowscale_stride, d.epi_act = None, 0, int(act)
    if rowscale is not None:      # adaLN-Zero gate, [N or 1, Cout] fp32 (may be a column slice)
        if (_f32(rowscale).stride(-1) != 1 or rowscale.shape[1] != cout or rowscale.shape[0] not in (1, n)
                or rowscale.data_ptr() % 16):
            raise L.B200Error("conv2d: rowscale must be [N or 1, Cout] fp32 with dense, 16B-aligned rows")
        d.rowscale = rowscale.data_ptr()
        d.rowscale_stride = rowscale.stride(0) if rowscale.shape[0] == n and n > 1 else 0
    if residual is not None and (head_channels or out_f32 or residual.shape != out.shape or residual.dtype != out.dtype):
        raise L.B200Error("conv2d: residual shape/dtype mismatch")
    d.residual = _p(residual)
    if residual_lo is not None and (residual is None or residual_lo.shape != residual.shape
                                    or residual_lo.dtype != residual.dtype):
        raise L.B200Error("conv2d: residual_lo must match residual")
    d.residual_lo = _p(residual_lo)
    if out_lo is not None and out_lo is not False:
        if head_channels or out_f32:
            raise L.B200Error("conv2d: out_lo needs a 16-bit NHWC output")
        if out_lo is True:
            out_lo = torch.empty_like(out)
        elif out_lo.shape != out.shape or out_lo.dtype != out.dtype:
            raise L.B200Error("conv2d: out_lo shape/dtype mismatch")
        d.out_lo = _p(out_lo)
    else:
        out_lo = None
    d.res_scale = res_scale
    d.out = None if (head_channels or out_f32) else _p(out)
    if stats is not None and (stats.dtype != torch.float64 or stats.numel() != n * cout * 2):
        raise L.B200Error("conv2d: stats must be a zeroed float64 [N, Cout, 2] tensor")
    d.gn_sums = _p(stats)
    d.act_dtype = act_code(src0.dtype)
    d.range_events = range_counter(src0.device).data_ptr() if src0.dtype == torch.float16 and _range_check else None
    ks = L.lib().b200_conv_ksplit(C.byref(d))
    if ks > 1 and (out_lo is not None or residual_lo is not None):
        # split-K layers (small maps) finish in a separate kernel that has no low planes: they are best-effort extra
        # precision, so the result simply keeps 16 bits there
        d.out_lo, d.residual_lo, out_lo, residual_lo = None, None, None, None
    if ks > 1:          # small maps run split-K: fp32 partial planes in a per-stream scratch buffer (never shared by
        need = ks * out.numel() * 4          # two launches that may overlap)
        ws = _workspace(src0.device, need)
        d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    # ALGORITHMIC work (the reference's count): a 2x2 sub-pixel launch stands for its share of a 3x3 convolution on the
    # upsampled map (9 taps per output pixel) although it executes 4
    kk = sum(sg[0].shape[3] * (9 if (out_sub is not None and sg[3] == 2) else sg[3] * sg[3]) for sg in segs)
    flops = 2.0 * n * hout * wout * cout * kk
    nbytes = 2.0 * (sum(sg[0].numel() for sg in segs) + out.numel() + cout * kk
                    + (residual.numel() if residual is not None else 0)
                    + (out_lo.numel() if out_lo is not None else 0)
                    + (residual_lo.numel() if residual_lo is not None else 0))
    _timed("conv_igemm", flops, nbytes,
           lambda: L.check(L.lib().b200_conv2d_igemm(C.byref(d), _stream()), "conv2d_igemm"),
           None if _timer is None else
           f"{n}x{hout}x{wout} s{stride} " + "+".join(f"k{sg[3]}c{sg[0].shape[3]}" for sg in segs) + f" ->{cout}"
           + (" +lo" if out_lo is not None else "") + (" +rlo" if residual_lo is not None else ""))
    return out if out_lo is None else (out, out_lo)


_workspaces = {}

# ---------------------------------------------------------------------------------------------
# fp16 range guard
# ---------------------------------------------------------------------------------------------
# fp16 activations saturate at +-65504 where the reference's fp32 path keeps going.  Every convolution launch counts the
# output tiles in which it had to saturate a result into a per-device counter (b200_conv_desc.range_events); the sampling
# loops (and direct network calls) read it once per run and RAISE instead of returning slices computed from clipped
# activations.  B200DIFF_RANGE_CHECK=0 disables counting and checking.
_range_check = os.environ.get("B200DIFF_RANGE_CHECK", "1") != "0"
_range_counters = {}


def range_counter(device) -> Tensor:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    t = _range_counters.get(idx)
    if t is None:
        t = _range_counters[idx] = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", idx))
    return t


def check_range(device, what: str = "the network") -> None:
    """Raise if any fp16 activation was saturated since the last check (one 4-byte device->host read: synchronises)."""
    if not _range_check:
        return
    idx = device.index if device.index is not None else torch.cuda.current_device()
    t = _range_counters.get(idx)
    if t is None:
        return
    n = int(t.item())
    if n:
        t.zero_()
        raise L.B200Error(
            f"fp16 range overflow in {what}: {n} convolution output tile(s) held values beyond +-65504 and were "
            "saturated, so the result would differ from the reference's fp32 arithmetic.  Run this checkpoint with bf16 "
            "activations (act_dtype=torch.bfloat16 or B200DIFF_ACT_DTYPE=bf16).")



def _workspace(device, nbytes: int) -> Tensor:
    """Split-K scratch, one buffer per (device, stream): captured CUDA graphs keep raw pointers, so a buffer that was
    ever handed out is never freed or shrunk — a larger request retires it into the keep-alive list."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    cur = _workspaces.get(key)
    if cur is None or cur[0].numel() < nbytes:
        size = max(nbytes, 32 << 20)
        buf = torch.empty(size, dtype=torch.uint8, device=device)
        _workspaces[key] = (buf, (cur[1] + [cur[0]]) if cur else [])
        cur = _workspaces[key]
    return cur[0]


def conv3x3_in(x_nchw: Tensor, w_oihw: Tensor, bias: Optional[Tensor], dtype=torch.float16,
               stats: Optional[Tensor] = None, want_lo: bool = False):
    """``want_lo``: also write the LOW plane of the output (see conv2d) and return ``(out, out_lo)``."""
    _dev(x_nchw)
    n, ci, h, w = x_nchw.shape
    co = w_oihw.shape[0]
    out = torch.empty((n, h, w, co), device=x_nchw.device, dtype=dtype)
    lo = torch.empty_like(out) if want_lo else None
    _timed("conv3x3_in", 2.0 * n * h * w * co * ci * 9, 4.0 * x_nchw.numel() + 2.0 * out.numel() * (2 if want_lo else 1),
           lambda: L.check(L.lib().b200_conv3x3_in(_p(_f32(x_nchw)), _p(_f32(w_oihw)), _p(_f32(bias)), _p(out), n, ci,
                                                   h, w, co, _p(stats), _p(lo), act_code(dtype), _stream()), "conv3x3_in"))
    return (out, lo) if want_lo else out


def conv3x3_out(x: Tensor, w_oihw: Tensor, bias: Optional[Tensor]) -> Tensor:
    _dev(x)
    n, h, w, ci = x.shape
    co = w_oihw.shape[0]
    out = torch.empty((n, co, h, w), device=x.device, dtype=torch.float32)
    _timed("conv3x3_out", 2.0 * n * h * w * co * ci * 9, 2.0 * x.numel() + 4.0 * out.numel(),
           lambda: L.check(L.lib().b200_conv3x3_out(_p(x), _p(_f32(w_oihw)), _p(_f32(bias)), _p(out), n, ci, h, w, co,
                                                    act_code(x.dtype), _stream()), "conv3x3_out"))
    return out


# ---------------------------------------------------------------------------------------------
# normalisation
# ---------------------------------------------------------------------------------------------
class StatsArena:
    """Bump allocator for the per-(image, channel) fp64 statistics buffers of one network forward: ONE memset at
    ``begin`` instead of one per GroupNorm.  Buffers that were ever handed out stay alive (captured CUDA graphs keep
    raw pointers to them).

    The arena is only ever (re)allocated and zeroed in ``begin`` — on the caller's stream, BEFORE the network forks its
    side streams (dsunet.py) — sized from the high-water mark of earlier forwards.  A forward that needs more than the
    arena holds (the first one of a new, larger shape) gets private zero-filled buffers for the overflow, each zeroed on
    the stream its producer kernel runs on, so no stream can ever see a slot being cleared under it; the next ``begin``
    grows the arena."""

    MIN = 1 << 18

    def __init__(self):
        self.buf: Optional[Tensor] = None
        self.off = 0
        self.need = 0          # doubles requested in the current forward
        self.high = 0          # high-water mark over all forwards
        self.overflows = 0     # take() calls served outside the arena (diagnostics / tests)
        self.retired = []

    def begin(self, device) -> None:
        want = max(self.MIN, -(-int(self.high * 1.25) // 4096) * 4096)
        if self.buf is None or self.buf.device != device or self.buf.numel() < want:
            if self.buf is not None:
                self.retired.append(self.buf)
            self.buf = torch.zeros(want, dtype=torch.float64, device=device)
        else:
            self.buf.zero_()
        self.off = 0
        self.need = 0

    def take(self, n: int, c: int) -> Tensor:
        k = n * c * 2
        self.need += k
        self.high = max(self.high, self.need)
        if self.off + k > self.buf.numel():
            self.overflows += 1
            return torch.zeros((n, c, 2), dtype=torch.float64, device=self.buf.device)
        v = self.buf[self.off:self.off + k].view(n, c, 2)
        self.off += k
        return v


def channel_stats(x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """sums[N, C, 2] fp64 = per-(image, channel) (sum, sum of squares) of an NHWC activation."""
    _dev(x)
    n, c = x.shape[0], x.shape[-1]
    hw = x.numel() // (n * c)
    if out is None:
        out = torch.zeros((n, c, 2), device=x.device, dtype=torch.float64)
    _timed("channel_stats", 0.0, 2.0 * x.numel(),
           lambda: L.check(L.lib().b200_channel_stats(_p(x), n, hw, c, _p(out), act_code(x.dtype), _stream()),
                           "channel_stats"))
    return out


def group_norm(x0: Tensor, gamma: Tensor, beta: Tensor, x1: Optional[Tensor] = None, groups: int = 32,
               eps: float = 1e-5, silu: bool = True, film: Optional[Tensor] = None,
               stats: Optional[Tensor] = None, cs0: Optional[Tensor] = None, cs1: Optional[Tensor] = None,
               preshift: Optional[Tensor] = None, wsets: int = 1, film_gstride: int = 0,
               preshift_gstride: int = 0, x0_lo: Optional[Tensor] = None, x1_lo: Optional[Tensor] = None) -> Tensor:
    """GroupNorm(+FiLM)(+SiLU) of the channel concatenation [x0 | x1]; NHWC in, NHWC out.

    With ``cs0`` (and ``cs1`` when ``x1`` is given) — the producers' per-channel sums — this is ONE read + ONE write
    pass; otherwise a statistics pass runs first (``stats``: optional pre-zeroed fp64 workspace [N, groups, 2])."""
    _dev(x0)
    n = x0.shape[0]
    hw = x0.numel() // (n * x0.shape[-1])
    c0 = x0.shape[-1]
    c1 = x1.shape[-1] if x1 is not None else 0
    code = act_code(x0.dtype)
    lib = L.lib()
    out = torch.empty(tuple(x0.shape[:-1]) + (c0 + c1,), device=x0.device, dtype=x0.dtype)
    if wsets > 1 and (cs0 is None or (x1 is not None and cs1 is None) or n % wsets or gamma.numel() != wsets * (c0 + c1)):
        raise L.B200Error("group_norm: grouped launch needs producer sums, N % wsets == 0 and [wsets, C] gamma / beta")
    film_p = None if film is None else film.data_ptr()
    film_s = 0 if film is None or film.shape[0] == 1 else film.stride(0)   # [1, 2C]: shared by the batch
    ps_p, ps_s = None, 0
    if preshift is not None:      # x0 + preshift[n, c] is what gets normalised (cached-conv1 ResBlocks); fp32 [N or 1, C0]
        npg = n // max(wsets, 1)
        if cs0 is None or _f32(preshift).stride(-1) != 1 or preshift.shape[-1] != c0 or preshift.shape[0] not in (1, npg):
            raise L.B200Error("group_norm: preshift must be fp32 [N or 1, C0] and needs the producer's channel sums")
        ps_p = preshift.data_ptr()
        ps_s = preshift.stride(0) if preshift.shape[0] == npg and npg > 1 else 0
    for lo_t, src_t in ((x0_lo, x0), (x1_lo, x1)):
        if lo_t is not None and (src_t is None or lo_t.shape != src_t.shape or lo_t.dtype != src_t.dtype):
            raise L.B200Error("group_norm: a low plane must match its source")
    if (x0_lo is not None or x1_lo is not None) and not (cs0 is not None and (x1 is None or cs1 is not None)):
        raise L.B200Error("group_norm: low planes need the producers' channel sums (single-pass form)")
    if cs0 is not None and (x1 is None or cs1 is not None):
        lo_elems = (x0_lo.numel() if x0_lo is not None else 0) + (x1_lo.numel() if x1_lo is not None else 0)
        _timed("groupnorm_silu", 0.0, 2.0 * (2.0 * out.numel() + lo_elems),
               detail=f"{n}x{hw} c{c0}+{c1}" + (" +lo" if lo_elems else ""), launch=
               lambda: L.check(lib.b200_groupnorm_apply_cs(_p(x0), c0, _p(cs0), _p(x1), c1, _p(cs1), n, hw, groups, eps,
                                                           _p(_f32(gamma)), _p(_f32(beta)), film_p, film_s, int(silu),
                                                           _p(out), ps_p, ps_s, wsets, film_gstride, preshift_gstride,
                                                           _p(x0_lo), _p(x1_lo), code, _stream()), "groupnorm_apply_cs"))
        return out
    if stats is None:
        stats = torch.zeros((n, groups, 2), device=x0.device, dtype=torch.float64)

    def launch():
        L.check(lib.b200_groupnorm_stats(_p(x0), c0, _p(x1), c1, n, hw, groups, _p(stats), code, _stream()),
                "groupnorm_stats")
        L.check(lib.b200_groupnorm_apply(_p(x0), c0, _p(x1), c1, n, hw, groups, eps, _p(stats), _p(_f32(gamma)),
                                         _p(_f32(beta)), film_p, film_s, int(silu), _p(out), code, _stream()),
                "groupnorm_apply")

    # algorithmic bytes: read the input once, write the output once (16-bit each)
    _timed("groupnorm_silu", 0.0, 2.0 * 2.0 * out.numel(), launch)
    return out


# ---------------------------------------------------------------------------------------------
# attention / timestep path / elementwise
# ---------------------------------------------------------------------------------------------
def attention_qkv(qkv: Tensor, heads: int, new_order: bool) -> Tensor:
    """qkv: [B, L, 3*heads*ch] -> [B, L, heads*ch];  softmax((q*s)(k*s)) v with s = ch^-1/4."""
    _dev(qkv)
    b, l, c3 = qkv.shape
    ch = c3 // (3 * heads)
    out = torch.empty((b, l, heads * ch), device=qkv.device, dtype=qkv.dtype)
    _timed("attention", 4.0 * b * heads * l * l * ch, 2.0 * (qkv.numel() + out.numel()),
           detail=f"B{b} L{l} h{heads} d{ch}", launch=lambda: L.check(L.lib().b200_attention_qkv(_p(qkv), _p(out), b, l, heads, ch, int(new_order),
                                                      float(ch) ** -0.5, act_code(qkv.dtype), _stream()),
                           "attention_qkv"))
    return out


def attention(q: Tensor, k: Tensor, v: Tensor, heads: int, d: int, scale: float, q_off: int = 0, k_off: int = 0,
              v_off: int = 0, q_hstride: Optional[int] = None, k_hstride: Optional[int] = None,
              v_hstride: Optional[int] = None, force_simt: bool = False, lse: Optional[Tensor] = None) -> Tensor:
    """General fused attention over token tensors q [B,Lq,ldq], k/v [B,Lk,ld*] (heads are channel slices)."""
    _dev(q)
    b, lq, ldq = q.shape
    lk = k.shape[1]
    out = torch.empty((b, lq, heads * d), device=q.device, dtype=q.dtype)
    a = L.AttnDesc()
    a.q, a.k, a.v, a.out = _p(q), _p(k), _p(v), _p(out)
    a.B, a.Lq, a.Lk, a.heads, a.d = b, lq, lk, heads, d
    a.ldq, a.ldk, a.ldv, a.ldo = ldq, k.shape[2], v.shape[2], heads * d
    a.q_off, a.k_off, a.v_off = q_off, k_off, v_off
    a.q_hstride = d if q_hstride is None else q_hstride
    a.k_hstride = d if k_hstride is None else k_hstride
    a.v_hstride = d if v_hstride is None else v_hstride
    a.scale, a.act_dtype, a.force_simt = scale, act_code(q.dtype), int(force_simt)
    if lse is not None and (lse.dtype != torch.float32 or lse.numel() != b * heads * lq):
        raise L.B200Error("attention: lse must be fp32 [B, heads, Lq]")
    a.lse = _p(lse)
    _timed("attention", 4.0 * b * heads * lq * lk * d, 2.0 * (b * heads * d * (2 * lq + 2 * lk)),
           lambda: L.check(L.lib().b200_attention(C.byref(a), _stream()), "attention"))
    return out


def layer_norm(x: Tensor, gamma: Optional[Tensor], beta: Optional[Tensor], eps: float = 1e-5,
               shift: Optional[Tensor] = None, scale: Optional[Tensor] = None) -> Tensor:
    """LayerNorm over the last dim of tokens [B, L, C] (16-bit); optional adaLN modulation with fp32 shift / scale [B, C]
    (column slices of a wider matrix are fine): y * (1 + scale[b]) + shift[b]."""
    _dev(x)
    b, l, c = x.shape
    out = torch.empty_like(x)
    ms = 0
    if shift is not None:
        if shift.stride(-1) != 1 or scale.stride(-1) != 1 or shift.stride(0) != scale.stride(0):
            raise L.B200Error("layer_norm: shift / scale must be fp32 rows with the same pitch")
        ms = shift.stride(0)
    _timed("layernorm", 0.0, 4.0 * x.numel(),
           lambda: L.check(L.lib().b200_layernorm(_p(x), _p(_f32(gamma)), _p(_f32(beta)),
                                                  None if shift is None else _f32(shift).data_ptr(),
                                                  None if scale is None else _f32(scale).data_ptr(), ms, l, b * l, c, eps,
                                                  _p(out), act_code(x.dtype), _stream()), "layernorm"))
    return out


def geglu(x: Tensor) -> Tensor:
    """tokens [B, L, 2I] -> [B, L, I] = x[..., :I] * gelu(x[..., I:])."""
    _dev(x)
    b, l, c2 = x.shape
    out = torch.empty((b, l, c2 // 2), device=x.device, dtype=x.dtype)
    _timed("geglu", 0.0, 2.0 * (x.numel() + out.numel()),
           lambda: L.check(L.lib().b200_geglu(_p(x), _p(out), b * l, c2 // 2, act_code(x.dtype), _stream()), "geglu"))
    return out


def patchify(x: Tensor, p: int, dtype: torch.dtype) -> Tensor:
    """NCHW fp32 [N, C, H, W] -> patch tokens [N, (H/p)*(W/p), Kpad] (column order of PatchEmbed.proj.weight.view(d, -1))."""
    _dev(x)
    n, c, h, w = x.shape
    kpad = (c * p * p + 7) // 8 * 8
    out = torch.empty((n, (h // p) * (w // p), kpad), device=x.device, dtype=dtype)
    _timed("patchify", 0.0, 4.0 * x.numel() + 2.0 * out.numel(),
           lambda: L.check(L.lib().b200_patchify(_p(_f32(x.contiguous())), _p(out), n, c, h, w, p, kpad, act_code(dtype),
                                                 _stream()), "patchify"))
    return out


def unpatchify(tok: Tensor, cout: int, gh: int, gw: int, p: int) -> Tensor:
    """fp32 tokens [N, gh*gw, ld] -> NCHW fp32 [N, cout, gh*p, gw*p] (DiT.unpatchify)."""
    _dev(tok)
    n, _, ld = tok.shape
    out = torch.empty((n, cout, gh * p, gw * p), device=tok.device, dtype=torch.float32)
    L.check(L.lib().b200_unpatchify(_p(_f32(tok)), _p(out), n, cout, gh, gw, p, ld, _stream()), "unpatchify")
    return out


def add_f32(a: Tensor, b: Tensor) -> Tensor:
    _dev(a)
    if a.shape != b.shape:
        raise L.B200Error("add_f32: shape mismatch")
    out = torch.empty_like(a)
    L.check(L.lib().b200_add_f32(_p(_f32(a)), _p(_f32(b)), _p(out), a.numel(), _stream()), "add_f32")
    return out


def cfg_combine(uncond: Tensor, cond: Tensor, scale: float) -> Tensor:
    """uncond + scale * (cond - uncond) (classifier-free guidance), fp32."""
    _dev(uncond)
    if uncond.shape != cond.shape:
        raise L.B200Error("cfg_combine: shape mismatch")
    out = torch.empty_like(uncond)
    L.check(L.lib().b200_cfg_combine(_p(_f32(uncond)), _p(_f32(cond)), float(scale), _p(out), uncond.numel(), _stream()),
            "cfg_combine")
    return out


def softmax_rows(s: Tensor, scale: float, dtype: torch.dtype) -> Tensor:
    """softmax(scale * s, dim=-1): fp32 logits [..., cols] -> 16-bit probabilities."""
    _dev(s)
    cols = s.shape[-1]
    out = torch.empty(s.shape, device=s.device, dtype=dtype)
    _timed("softmax_rows", 0.0, 4.0 * s.numel() + 2.0 * s.numel(),
           lambda: L.check(L.lib().b200_softmax_rows(_p(_f32(s)), _p(out), s.numel() // cols, cols, scale, act_code(dtype),
                                                     _stream()), "softmax_rows"))
    return out


def conv1x1_f32(x: Tensor, w: Tensor, bias: Optional[Tensor]) -> Tensor:
    """fp32 NCHW pointwise conv with <= 16 channels (quant_conv / post_quant_conv)."""
    _dev(x)
    n, ci, h, wd = x.shape
    co = w.shape[0]
    out = torch.empty((n, co, h, wd), device=x.device, dtype=torch.float32)
    L.check(L.lib().b200_conv1x1_f32_nchw(_p(_f32(x.contiguous())), _p(_f32(w.reshape(co, ci).contiguous())),
                                          _p(_f32(bias)), _p(out), n, ci, co, h * wd, _stream()), "conv1x1_f32_nchw")
    return out


def gaussian_sample(moments: Tensor, noise: Optional[Tensor]) -> Tensor:
    """mean + exp(0.5 * clamp(logvar, -30, 20)) * noise over moments [N, 2C, H, W]; noise None -> the mode."""
    _dev(moments)
    n, c2, h, w = moments.shape
    out = torch.empty((n, c2 // 2, h, w), device=moments.device, dtype=torch.float32)
    L.check(L.lib().b200_gaussian_sample(_p(_f32(moments.contiguous())), _p(_f32(noise)), _p(out), n, c2 // 2, h * w,
                                         _stream()), "gaussian_sample")
    return out


def timestep_embedding(t_table: Tensor, freqs: Tensor, batch: int, dim: int,
                       step_ptr: Optional[Tensor] = None) -> Tensor:
    _dev(t_table)
    out = torch.empty((batch, dim), device=t_table.device, dtype=torch.float32)
    L.check(L.lib().b200_timestep_embedding(_p(_f32(t_table)), _p(step_ptr), _p(_f32(freqs)), _p(out), batch, dim,
                                            _stream()), "timestep_embedding")
    return out


def linear_f32(x: Tensor, w: Tensor, bias: Optional[Tensor], in_silu: bool = False, out_silu: bool = False) -> Tensor:
    _dev(x)
    b, k = x.shape
    o = w.shape[0]
    if in_silu and b > 8:      # the GEMV kernel activates its input per output row: for big batches do it once up front
        act = torch.empty_like(x, memory_format=torch.contiguous_format)
        L.check(L.lib().b200_silu_f32(_p(_f32(x.contiguous())), _p(act), x.numel(), _stream()), "silu_f32")
        x, in_silu = act, False
    out = torch.empty((b, o), device=x.device, dtype=torch.float32)
    L.check(L.lib().b200_linear_f32(_p(_f32(x)), _p(_f32(w)), _p(_f32(bias)), _p(out), b, k, o, int(in_silu),
                                    int(out_silu), _stream()), "linear_f32")
    return out


def gather_row(table: Tensor, step_ptr: Tensor) -> Tensor:
    """table [S, W] fp32, device int step -> [1, W] (the row of the current sampling step)."""
    _dev(table)
    out = torch.empty((1, table.shape[1]), device=table.device, dtype=torch.float32)
    L.check(L.lib().b200_gather_row_f32(_p(_f32(table)), _p(step_ptr), _p(out), table.shape[1], _stream()), "gather_row")
    return out


def upsample_nearest2x(x: Tensor) -> Tensor:
    _dev(x)
    n, h, w, c = x.shape
    out = torch.empty((n, 2 * h, 2 * w, c), device=x.device, dtype=x.dtype)
    _timed("upsample2x", 0.0, 2.0 * (x.numel() + out.numel()),
           lambda: L.check(L.lib().b200_upsample_nearest2x(_p(x), _p(out), n, h, w, c, act_code(x.dtype), _stream()),
                           "upsample"))
    return out


def head_gather9(taps: Tensor, bias: Optional[Tensor], channels: int) -> Tensor:
    """Second half of a tap-split 3x3 network head (``unet.run_out_head``): ``taps`` is the fp32 NHWC result of the 1x1
    convolution to 9 * channels tap channels (rows may be padded); returns fp32 NCHW [N, channels, H, W]."""
    _dev(taps)
    n, h, w, ldp = taps.shape
    if taps.dtype != torch.float32 or not taps.is_contiguous() or not 1 <= channels <= 3 or ldp < 9 * channels:
        raise L.B200Error("head_gather9: needs contiguous fp32 NHWC taps of >= 9 * channels columns, 1..3 channels")
    out = torch.empty((n, channels, h, w), device=taps.device, dtype=torch.float32)
    _timed("head_gather9", 0.0, 4.0 * (n * h * w * 9 * channels + out.numel()),
           lambda: L.check(L.lib().b200_head_gather9(_p(taps), _p(_f32(bias)) if bias is not None else None, _p(out), n, h, w,
                                                     channels, ldp, _stream()), "head_gather9"))
    return out


def avg_pool2x(x: Tensor) -> Tensor:
    """nn.AvgPool2d(2, 2) on an NHWC activation."""
    _dev(x)
    n, h, w, c = x.shape
    out = torch.empty((n, h // 2, w // 2, c), device=x.device, dtype=x.dtype)
    _timed("avgpool2x", 0.0, 2.0 * (x.numel() + out.numel()),
           lambda: L.check(L.lib().b200_avgpool2x(_p(x), _p(out), n, h, w, c, act_code(x.dtype), _stream()), "avgpool2x"))
    return out


def add4_scale(a: Tensor, b: Optional[Tensor] = None, c: Optional[Tensor] = None, d: Optional[Tensor] = None,
               scale: float = 1.0, silu: bool = False, out: Optional[Tensor] = None) -> Tensor:
    _dev(a)
    if out is None:
        out = torch.empty_like(a)
    elif out.shape != a.shape or out.dtype != a.dtype:
        raise L.B200Error("add4_scale: out shape/dtype mismatch")
    nin = 1 + (b is not None) + (c is not None) + (d is not None)
    _timed("add4_scale", 0.0, 2.0 * a.numel() * (nin + 1),
           lambda: L.check(L.lib().b200_add4_scale(_p(a), _p(b), _p(c), _p(d), scale, _p(out), a.numel(), int(silu),
                                                   act_code(a.dtype), _stream()), "add4_scale"))
    return out


def add4_scale_stats(a: Tensor, b: Optional[Tensor], c: Optional[Tensor], d: Optional[Tensor], scale: float,
                     stats: Tensor, los: Sequence[Optional[Tensor]] = (), want_lo: bool = False):
    """scale * (a + b + c + d) over NHWC tensors, accumulating the result's per-(image, channel) sums into ``stats``.
    ``los``: low planes of (a, b, c, d) (None where a source has none), added in; ``want_lo``: also write the low plane of
    the result and return ``(out, out_lo)``."""
    _dev(a)
    n, ch = a.shape[0], a.shape[-1]
    hw = a.numel() // (n * ch)
    out = torch.empty_like(a)
    nin = 1 + (b is not None) + (c is not None) + (d is not None)
    los = list(los) + [None] * (4 - len(los))
    for lo_t in los:
        if lo_t is not None and (lo_t.shape != a.shape or lo_t.dtype != a.dtype):
            raise L.B200Error("add4_scale_stats: a low plane must match its source")
    out_lo = torch.empty_like(a) if want_lo else None
    planes = nin + 1 + sum(t is not None for t in los) + (1 if want_lo else 0)
    _timed("add4_scale", 0.0, 2.0 * a.numel() * planes,
           lambda: L.check(L.lib().b200_add4_scale_stats(_p(a), _p(b), _p(c), _p(d), scale, _p(out), n, hw, ch, _p(stats),
                                                         _p(los[0]), _p(los[1]), _p(los[2]), _p(los[3]), _p(out_lo),
                                                         act_code(a.dtype), _stream()), "add4_scale_stats"))
    return (out, out_lo) if want_lo else out


def se_attention(x: Tensor, w1: Tensor, w2: Tensor, out: Optional[Tensor] = None) -> Tensor:
    _dev(x)
    n, h, w, c = x.shape
    cr = w1.shape[0]
    scratch = torch.empty((2, n, c), device=x.device, dtype=torch.float32)
    if out is None:
        out = torch.empty_like(x)
    elif out.shape != x.shape or out.dtype != x.dtype:
        raise L.B200Error("se_attention: out shape/dtype mismatch")
    L.check(L.lib().b200_se_attention(_p(x), _p(_f32(w1)), _p(_f32(w2)), _p(out), _p(scratch), n, h * w, c, cr,
                                      act_code(x.dtype), _stream()), "se_attention")
    return out


# ---------------------------------------------------------------------------------------------
# sampler update
# ---------------------------------------------------------------------------------------------
def sampler_update(mode: int, param: int, x: Tensor, model_out: Tensor, noise: Optional[Tensor], coef: Tensor,
                   step: int = 0, step_ptr: Optional[Tensor] = None, clip_denoised: bool = True,
                   learned_range: bool = False, want_x0: bool = True, x_prev: Optional[Tensor] = None,
                   pred_x0: Optional[Tensor] = None):
    _dev(x)
    n, c = x.shape[0], x.shape[1]
    hw = x.numel() // (n * c)
    mc = model_out.shape[1]
    if mc != (2 * c if learned_range else c):
        raise L.B200Error(f"sampler_update: model_out has {mc} channels, expected {2 * c if learned_range else c}")
    if x_prev is None:
        x_prev = torch.empty_like(x)
    if pred_x0 is None and want_x0:
        pred_x0 = torch.empty_like(x)
    d = L.SamplerDesc()
    d.mode, d.param, d.clip_denoised, d.learned_range = mode, param, int(clip_denoised), int(learned_range)
    d.N, d.C, d.HW = n, c, hw
    d.x, d.model_out, d.noise, d.coef = _p(_f32(x)), _p(_f32(model_out)), _p(_f32(noise)), _p(_f32(coef))
    d.step_ptr = _p(step_ptr)
    d.step_host = step
    d.x_prev, d.pred_x0 = _p(x_prev), _p(pred_x0)
    nb = 4.0 * (x.numel() + model_out.numel() + (noise.numel() if noise is not None else 0) + x_prev.numel()
                + (pred_x0.numel() if pred_x0 is not None else 0))
    _timed("sampler_update", 0.0, nb, lambda: L.check(L.lib().b200_sampler_update(C.byref(d), _stream()),
                                                       "sampler_update"))
    return x_prev, pred_x0


def dpm_solver_update(param: int, x: Tensor, model_out: Optional[Tensor], coef: Tensor, hist: Optional[Tensor],
                      step: int = 0, step_ptr: Optional[Tensor] = None, phase: int = 0, pred_x0: Optional[Tensor] = None,
                      thresh: Optional[Tensor] = None, x_next: Optional[Tensor] = None) -> Optional[Tensor]:
    """One DPM-Solver++ (2M) step (b200_sampler_update, mode DPMPP_2M): phase 0 fused; phase 1 writes the raw x0 into
    ``pred_x0``; phase 2 thresholds ``pred_x0`` with the per-sample ``thresh`` and updates.  ``hist`` is the previous x0."""
    _dev(x)
    n, c = x.shape[0], x.shape[1]
    d = L.SamplerDesc()
    d.mode, d.param, d.clip_denoised, d.learned_range = L.SAMPLER_DPMPP_2M, param, 0, 0
    d.N, d.C, d.HW = n, c, x.numel() // (n * c)
    if phase != 1 and x_next is None:
        x_next = torch.empty_like(x)
    d.x, d.model_out, d.noise, d.coef = _p(_f32(x)), _p(_f32(model_out)), None, _p(_f32(coef))
    d.step_ptr, d.step_host = _p(step_ptr), step
    d.x_prev, d.pred_x0, d.hist, d.thresh, d.phase = _p(x_next), _p(pred_x0), _p(hist), _p(_f32(thresh)), phase
    nb = 4.0 * x.numel() * (4 if phase == 0 else 3 if phase == 1 else 5)
    _timed("sampler_update", 0.0, nb, lambda: L.check(L.lib().b200_sampler_update(C.byref(d), _stream()),
                                                       "sampler_update(dpm-solver++)"))
    return x_next


def abs_quantile(x: Tensor, q: float, floor_val: float, out: Optional[Tensor] = None) -> Tensor:
    """out[n] = max(quantile(|x[n]|, q), floor_val) per sample (dynamic thresholding scale)."""
    _dev(x)
    n = x.shape[0]
    if out is None:
        out = torch.empty((n,), device=x.device, dtype=torch.float32)
    L.check(L.lib().b200_abs_quantile(_p(_f32(x)), n, x.numel() // n, float(q), float(floor_val), _p(out), _stream()),
            "abs_quantile")
    return out


def lincomb2(x: Tensor, y: Tensor, c1: Tensor, c2: Tensor, subtract: bool = False) -> Tensor:
    """c1[n] * x[n] (+|-) c2[n] * y[n] with per-sample fp32 coefficients [N] (schedule-table gathers)."""
    _dev(x)
    if x.shape != y.shape or c1.numel() != x.shape[0] or c2.numel() != x.shape[0]:
        raise L.B200Error("lincomb2: shape mismatch")
    xf, yf = _f32(x.contiguous()), _f32(y.contiguous())
    out = torch.empty_like(xf)
    n = x.shape[0]
    L.check(L.lib().b200_lincomb2(_p(xf), _p(yf), _p(_f32(c1.contiguous())), _p(_f32(c2.contiguous())), _p(out), n,
                                  x.numel() // n, int(subtract), _stream()), "lincomb2")
    return out


def advance_step(step_ptr: Tensor, delta: int = 1) -> None:
    _dev(step_ptr)
    L.check(L.lib().b200_advance_step(_p(step_ptr), delta, _stream()), "advance_step")
