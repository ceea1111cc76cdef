"""GPU parity: every kernel of libb200diff.so against a plain fp32 torch-CPU statement of the same op
(operands pre-rounded to the 16-bit activation type, so the tolerance only has to cover accumulation
order and the final 16-bit store), called through the C ABI wrappers in ``ops``.

Tolerances: fp16 store = 2^-11 relative; bf16 store = 2^-8 relative; fp32 paths 1e-5.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import dsdiff_oracle as O

pytestmark = pytest.mark.gpu

DTYPES = [torch.float16, torch.bfloat16]
STORE_EPS = {torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7}


def _ops():
    from diffusion_models_dsdiff_b200 import ops
    return ops


def rnd(seed, *shape, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def q(x, dt):
    return x.to(dt).float()


def nhwc(x, dt):           # NCHW fp32 cpu -> NHWC 16-bit cuda
    return x.permute(0, 2, 3, 1).contiguous().to(dt).cuda()


def nchw(y):               # NHWC cuda -> NCHW fp32 cpu
    return y.float().cpu().permute(0, 3, 1, 2).contiguous()


def close(got, ref, dt, what=""):
    err = (got - ref).abs().max().item()
    tol = STORE_EPS[dt] * max(1.0, ref.abs().max().item()) * 2.0
    assert err <= tol, f"{what}: max abs err {err:.3e} > tol {tol:.3e} (ref max {ref.abs().max():.3f})"


# -------------------------------------------------------------------------------------------------
# implicit-GEMM convolution
# -------------------------------------------------------------------------------------------------
CONV_CASES = [
    # N, H, W, Cin, Cout, k, stride
    (2, 16, 16, 32, 32, 3, 1),
    (1, 32, 32, 96, 96, 3, 1),       # K tail chunk of 32
    (2, 16, 16, 160, 160, 3, 1),     # BN=160, K tail
    (1, 24, 24, 64, 320, 3, 1),      # ragged tiles, 2 N-tiles
    (3, 8, 8, 64, 64, 3, 1),         # box spans 2 images, ragged in N
    (2, 32, 32, 64, 64, 3, 2),       # stride-2 (Downsample)
    (1, 34, 30, 32, 48, 3, 2),       # stride-2, ragged
    (2, 16, 16, 96, 288, 1, 1),      # 1x1 (qkv-style), BN=144
    (1, 128, 128, 64, 64, 3, 1),     # 128 tiles/CTA loop is not reached, many tiles
    (4, 128, 128, 32, 32, 3, 1),     # 512 tiles > 148 CTAs: persistent loop + TMEM double buffer
    (1, 16, 16, 1280, 640, 3, 1),    # long K loop (180 chunks): pipeline phase wrap
]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_igemm(case, dt):
    ops = _ops()
    n, h, w, ci, co, k, stride = case
    x = q(rnd(1, n, ci, h, w), dt)
    wt = q(rnd(2, co, ci, k, k, scale=1.0 / math.sqrt(ci * k * k)), dt)
    b = rnd(3, co, scale=0.1)
    ref = F.conv2d(x, wt, b, stride=stride, padding=k // 2)
    wp = ops.pack_conv_weight(wt.cuda(), dt)
    y = ops.conv2d([(nhwc(x, dt), wp, 0, k, k // 2)], stride=stride, bias=b.cuda())
    torch.cuda.synchronize()
    close(nchw(y), ref, dt, f"conv {case}")


@pytest.mark.parametrize("dt", DTYPES)
def test_conv_epilogue_rowbias_residual(dt):
    ops = _ops()
    n, h, w, ci, co = 2, 16, 16, 64, 96
    x = q(rnd(1, n, ci, h, w), dt)
    wt = q(rnd(2, co, ci, 3, 3, scale=0.05), dt)
    b = rnd(3, co, scale=0.1)
    rb_all = rnd(4, n, 3 * co)                       # column slice of a wider matrix
    res = q(rnd(5, n, co, h, w), dt)
    ref = F.conv2d(x, wt, b, padding=1) + rb_all[:, co:2 * co, None, None] + res
    rbc = rb_all.cuda()
    y = ops.conv2d([(nhwc(x, dt), ops.pack_conv_weight(wt.cuda(), dt), 0, 3, 1)], bias=b.cuda(),
                   rowbias=rbc[:, co:2 * co], residual=nhwc(res, dt))
    close(nchw(y), ref, dt, "conv epilogue")


@pytest.mark.parametrize("dt", DTYPES)
def test_conv_multi_segment_concat_and_fused_skip(dt):
    """conv3x3(a2) + conv1x1(cat[x0,x1]) in one launch (ResBlock out_layers + skip_connection)."""
    ops = _ops()
    n, h, w, c0, c1, co = 2, 16, 16, 96, 32, 64
    a2 = q(rnd(1, n, co, h, w), dt)
    x0, x1 = q(rnd(2, n, c0, h, w), dt), q(rnd(3, n, c1, h, w), dt)
    w2 = q(rnd(4, co, co, 3, 3, scale=0.04), dt)
    ws = q(rnd(5, co, c0 + c1, 1, 1, scale=0.08), dt)
    b = rnd(6, co, scale=0.1)
    ref = F.conv2d(a2, w2, b, padding=1) + F.conv2d(torch.cat([x0, x1], 1), ws)
    wsp = ops.pack_conv_weight(ws.cuda(), dt)
    y = ops.conv2d([(nhwc(a2, dt), ops.pack_conv_weight(w2.cuda(), dt), 0, 3, 1),
                    (nhwc(x0, dt), wsp, 0, 1, 0), (nhwc(x1, dt), wsp, c0, 1, 0)], bias=b.cuda())
    close(nchw(y), ref, dt, "multi-seg conv")


@pytest.mark.parametrize("dt", DTYPES)
def test_conv_in_out(dt):
    ops = _ops()
    x = rnd(1, 2, 4, 32, 32)
    w_in, b_in = rnd(2, 64, 4, 3, 3, scale=0.2), rnd(3, 64, scale=0.1)
    y = ops.conv3x3_in(x.cuda(), w_in.cuda(), b_in.cuda(), dt)
    close(nchw(y), F.conv2d(x, w_in, b_in, padding=1), dt, "conv3x3_in")
    x1 = rnd(9, 1, 1, 16, 16)
    y1 = ops.conv3x3_in(x1.cuda(), w_in[:, :1].contiguous().cuda(), None, dt)
    close(nchw(y1), F.conv2d(x1, w_in[:, :1], None, padding=1), dt, "conv3x3_in Cin=1")
    h = q(rnd(4, 2, 64, 32, 32), dt)
    for co in (1, 2):
        w_out, b_out = rnd(5, co, 64, 3, 3, scale=0.05), rnd(6, co, scale=0.1)
        o = ops.conv3x3_out(nhwc(h, dt), w_out.cuda(), b_out.cuda()).cpu()
        ref = F.conv2d(h, w_out, b_out, padding=1)
        assert (o - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())


# -------------------------------------------------------------------------------------------------
# GroupNorm (+SiLU, +FiLM, concat input)
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(2, 32, 8, 8), (2, 160, 32, 32), (1, 96, 64, 64), (3, 320, 16, 16), (1, 1920, 8, 8)])
def test_groupnorm_silu(shape, dt):
    ops = _ops()
    x = q(rnd(1, *shape) * 1.5 + 0.3, dt)
    g, b = 1 + 0.1 * rnd(2, shape[1]), 0.1 * rnd(3, shape[1])
    ref = F.silu(F.group_norm(x, 32, g, b, 1e-5))
    y = ops.group_norm(nhwc(x, dt), g.cuda(), b.cuda(), eps=1e-5, silu=True)
    close(nchw(y), ref, dt, f"gn+silu {shape}")
    ref2 = F.group_norm(x, 32, g, b, 1e-6)
    y2 = ops.group_norm(nhwc(x, dt), g.cuda(), b.cuda(), eps=1e-6, silu=False)
    close(nchw(y2), ref2, dt, f"gn {shape}")


@pytest.mark.parametrize("dt", DTYPES)
def test_groupnorm_concat_and_film(dt):
    ops = _ops()
    n, c0, c1, h, w = 2, 320, 160, 16, 16
    x0, x1 = q(rnd(1, n, c0, h, w), dt), q(rnd(2, n, c1, h, w) * 2 + 1, dt)
    c = c0 + c1
    g, b = 1 + 0.1 * rnd(3, c), 0.1 * rnd(4, c)
    ref = F.silu(F.group_norm(torch.cat([x0, x1], 1), 32, g, b, 1e-5))
    y = ops.group_norm(nhwc(x0, dt), g.cuda(), b.cuda(), nhwc(x1, dt), silu=True)
    close(nchw(y), ref, dt, "gn concat")
    film_all = rnd(5, n, 4 * c0 + 8, scale=0.5)
    film = film_all[:, 8:8 + 2 * c0]
    g0, b0 = g[:c0], b[:c0]
    scale, shift = film[:, :c0, None, None], film[:, c0:, None, None]
    ref = F.silu(F.group_norm(x0, 32, g0, b0, 1e-5) * (1 + scale) + shift)
    fc = film_all.cuda()
    y = ops.group_norm(nhwc(x0, dt), g0.contiguous().cuda(), b0.contiguous().cuda(), silu=True, film=fc[:, 8:8 + 2 * c0])
    close(nchw(y), ref, dt, "gn film")


def test_groupnorm_wide_concat_2560_channels():
    """SD-v1's decoder normalises cat([h, skip]) with 1280 + 1280 channels (> 2048: several column passes), statistics from
    the producers' per-channel sums."""
    ops = _ops()
    dt = torch.float16
    n, c0, c1, h, w = 2, 1280, 1280, 8, 8
    x0, x1 = q(rnd(1, n, c0, h, w), dt), q(rnd(2, n, c1, h, w) * 2 + 1, dt)
    g, b = 1 + 0.1 * rnd(3, c0 + c1), 0.1 * rnd(4, c0 + c1)
    ref = F.silu(F.group_norm(torch.cat([x0, x1], 1), 32, g, b, 1e-5))
    a0, a1 = nhwc(x0, dt), nhwc(x1, dt)
    y = ops.group_norm(a0, g.cuda(), b.cuda(), a1, silu=True, cs0=ops.channel_stats(a0), cs1=ops.channel_stats(a1))
    close(nchw(y), ref, dt, "gn 2560-channel concat")


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", [(1, 32, 32, 64, 96, 1), (2, 40, 24, 32, 320, 1), (3, 8, 8, 64, 64, 1), (2, 32, 32, 64, 64, 2)])
def test_conv_fused_channel_stats_and_single_pass_groupnorm(case, dt):
    """The conv epilogue's per-(image, channel) sums (tiles inside one image: fused; tiles spanning images: the
    standalone kernel behind the same ABI field) feed b200_groupnorm_apply_cs = GroupNorm in one read + one write."""
    ops = _ops()
    n, h, w, ci, co, stride = case
    x = q(rnd(1, n, ci, h, w), dt)
    wt = q(rnd(2, co, ci, 3, 3, scale=1.0 / math.sqrt(ci * 9)), dt)
    b = rnd(3, co, scale=0.1)
    ref = F.conv2d(x, wt, b, stride=stride, padding=1)
    st = torch.zeros(n, co, 2, dtype=torch.float64).cuda()
    y = ops.conv2d([(nhwc(x, dt), ops.pack_conv_weight(wt.cuda(), dt), 0, 3, 1)], stride=stride, bias=b.cuda(), stats=st)
    sums = torch.stack([ref.double().sum((2, 3)), (ref.double() ** 2).sum((2, 3))], -1)
    cnt = ref.shape[2] * ref.shape[3]
    st2 = ops.channel_stats(y)                                      # standalone kernel on the rounded tensor
    assert ((st2.cpu() - sums).abs() / cnt).max().item() < STORE_EPS[dt] * 4
    # fused epilogue sums: of the STORED 16-bit values on the fast path (== the standalone kernel up to fp32 summation
    # order), of the fp32 accumulators on the generic path; both within the 16-bit store noise of the exact sums
    yd = nchw(y).double()
    stored = torch.stack([yd.sum((2, 3)), (yd ** 2).sum((2, 3))], -1)
    err_stored = ((st.cpu() - stored).abs() / cnt).max().item()
    err_exact = ((st.cpu() - sums).abs() / cnt).max().item()
    assert min(err_stored, err_exact) < 1e-5 and err_exact < STORE_EPS[dt] * 4
    if co % 32 == 0:
        g, be = 1 + 0.1 * rnd(4, co), 0.1 * rnd(5, co)
        yq = nchw(y)
        refn = F.silu(F.group_norm(yq, 32, g, be, 1e-5))
        got = ops.group_norm(y, g.cuda(), be.cuda(), silu=True, cs0=st)
        close(nchw(got), refn, dt, f"single-pass gn {case}")


@pytest.mark.parametrize("dt", DTYPES)
def test_groupnorm_concat_from_channel_sums(dt):
    ops = _ops()
    n, c0, c1, h, w = 2, 320, 160, 16, 16
    x0, x1 = q(rnd(1, n, c0, h, w), dt), q(rnd(2, n, c1, h, w) * 2 + 1, dt)
    c = c0 + c1
    g, b = 1 + 0.1 * rnd(3, c), 0.1 * rnd(4, c)
    ref = F.silu(F.group_norm(torch.cat([x0, x1], 1), 32, g, b, 1e-5))
    a0, a1 = nhwc(x0, dt), nhwc(x1, dt)
    y = ops.group_norm(a0, g.cuda(), b.cuda(), a1, silu=True, cs0=ops.channel_stats(a0), cs1=ops.channel_stats(a1))
    close(nchw(y), ref, dt, "gn concat from channel sums")


# -------------------------------------------------------------------------------------------------
# attention
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("cfg", [(2, 64, 2, 16, True), (1, 256, 4, 32, True), (2, 100, 3, 48, False), (1, 1024, 2, 32, False)])
def test_attention(cfg, dt):
    ops = _ops()
    b, l, heads, ch, new_order = cfg
    qkv = q(rnd(1, b, 3 * heads * ch, l), dt)            # reference layout [B, 3C, L]
    scale = 1 / math.sqrt(math.sqrt(ch))
    if new_order:
        qq, kk, vv = qkv.chunk(3, dim=1)
        qq = (qq * scale).reshape(b * heads, ch, l); kk = (kk * scale).reshape(b * heads, ch, l)
        vv = vv.reshape(b * heads, ch, l)
    else:
        qq, kk, vv = qkv.reshape(b * heads, ch * 3, l).split(ch, dim=1)
        qq, kk = qq * scale, kk * scale
    wgt = torch.softmax(torch.einsum("bct,bcs->bts", qq, kk), dim=-1)
    ref = torch.einsum("bts,bcs->bct", wgt, vv).reshape(b, -1, l)      # [B, C, L]
    y = ops.attention_qkv(qkv.permute(0, 2, 1).contiguous().to(dt).cuda(), heads, new_order)
    got = y.float().cpu().permute(0, 2, 1)
    close(got, ref, dt, f"attention {cfg}")


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("cfg", [(1, 128, 128, 1, 32), (2, 200, 77, 8, 64), (1, 256, 300, 2, 160), (1, 1024, 77, 8, 80),
                                 (2, 320, 320, 3, 48), (1, 64, 64, 2, 40), (3, 130, 257, 2, 96)])
def test_attention_general(cfg, dt):
    """b200_attention with separate q / k / v token tensors (cross-attention shapes, Lq != Lk, ragged tiles).
    Head dims that are multiples of 16 take the tcgen05 flash kernel, 40 the CUDA-core kernel."""
    ops = _ops()
    b, lq, lk, heads, d = cfg
    qq, kk, vv = q(rnd(1, b, lq, heads * d), dt), q(rnd(2, b, lk, heads * d), dt), q(rnd(3, b, lk, heads * d), dt)
    scale = d ** -0.5
    sp = lambda t, l: t.view(b, l, heads, d).permute(0, 2, 1, 3)
    wgt = torch.softmax(sp(qq, lq) @ sp(kk, lk).transpose(-1, -2) * scale, dim=-1)
    ref = (wgt @ sp(vv, lk)).permute(0, 2, 1, 3).reshape(b, lq, heads * d)
    y = ops.attention(qq.to(dt).cuda(), kk.to(dt).cuda(), vv.to(dt).cuda(), heads, d, scale)
    close(y.float().cpu(), ref, dt, f"attention_general {cfg}")
    if d % 16 == 0 and d <= 64:      # both kernels agree
        y2 = ops.attention(qq.to(dt).cuda(), kk.to(dt).cuda(), vv.to(dt).cuda(), heads, d, scale, force_simt=True)
        close(y2.float().cpu(), ref, dt, f"attention_general simt {cfg}")


# -------------------------------------------------------------------------------------------------
# timestep path, elementwise
# -------------------------------------------------------------------------------------------------
def test_timestep_embedding_and_linear():
    ops = _ops()
    dim = 160
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(0, half, dtype=torch.float32) / half)
    t = torch.tensor([1.0, 41.0, 1961.0, 999.0, 530.5])
    ref = O.timestep_embedding(t, dim)
    got = ops.timestep_embedding(t.cuda(), freqs.cuda(), 5, dim).cpu()
    assert (got - ref).abs().max().item() < 2e-6
    step = torch.tensor([2], dtype=torch.int32).cuda()
    got2 = ops.timestep_embedding(t.cuda(), freqs.cuda(), 3, dim, step).cpu()
    assert (got2 - ref[2:3].expand(3, -1)).abs().max().item() < 2e-6
    w, b = rnd(1, 640, dim, scale=0.08), rnd(2, 640, scale=0.1)
    x = rnd(3, 5, dim)
    ref = F.linear(F.silu(x), w, b)
    got = ops.linear_f32(x.cuda(), w.cuda(), b.cuda(), in_silu=True).cpu()
    assert (got - ref).abs().max().item() < 1e-5
    ref = F.silu(F.linear(x, w, b))
    got = ops.linear_f32(x.cuda(), w.cuda(), b.cuda(), out_silu=True).cpu()
    assert (got - ref).abs().max().item() < 1e-5
    x11 = rnd(4, 11, dim)
    got = ops.linear_f32(x11.cuda(), w.cuda(), None).cpu()
    assert (got - F.linear(x11, w)).abs().max().item() < 1e-5


@pytest.mark.parametrize("dt", DTYPES)
def test_elementwise(dt):
    ops = _ops()
    x = q(rnd(1, 2, 32, 8, 8), dt)
    up = ops.upsample_nearest2x(nhwc(x, dt))
    assert torch.equal(nchw(up), F.interpolate(x, scale_factor=2, mode="nearest"))
    a, b, c, d = (q(rnd(i, 2, 32, 8, 8), dt) for i in range(4))
    y = ops.add4_scale(nhwc(a, dt), nhwc(b, dt), nhwc(c, dt), nhwc(d, dt), scale=0.25)
    close(nchw(y), (a + b + c + d) / 4, dt, "mean4")
    y = ops.add4_scale(nhwc(a, dt), silu=True)
    close(nchw(y), F.silu(a), dt, "silu")
    w1, w2 = rnd(5, 4, 32, scale=0.3), rnd(6, 32, 4, scale=0.3)
    gate = torch.sigmoid(F.linear(torch.relu(F.linear(x.mean(dim=(2, 3)), w1)), w2))
    y = ops.se_attention(nhwc(x, dt), w1.cuda(), w2.cuda())
    close(nchw(y), x * gate[:, :, None, None], dt, "se_attention")
    back = ops.nhwc_to_nchw(ops.nchw_to_nhwc(x.cuda(), 40, dt), 32).cpu()
    assert torch.equal(back, x)


@pytest.mark.parametrize("batch_rows", [True, False])
def test_groupnorm_preshift_matches_shifted_input(batch_rows):
    """GroupNorm of (x0 + per-(n, c) constant) with the statistics corrected analytically from x0's channel sums."""
    ops = _ops()
    dt = torch.float16
    n, c, h, w = 2, 96, 16, 16
    x = q(rnd(1, n, c, h, w) * 1.3, dt)
    s = rnd(2, n if batch_rows else 1, c) * 0.8
    g, b = 1 + 0.1 * rnd(3, c), 0.1 * rnd(4, c)
    ref = F.silu(F.group_norm(x + s[:, :, None, None], 32, g, b, 1e-5))
    a = nhwc(x, dt)
    wide = torch.zeros(s.shape[0], 3 * c + 8)          # preshift as a column slice of a wider matrix (like emb_all)
    wide[:, 8:8 + c] = s
    wide = wide.cuda()
    y = ops.group_norm(a, g.cuda(), b.cuda(), cs0=ops.channel_stats(a), preshift=wide[:, 8:8 + c])
    close(nchw(y), ref, dt, "gn preshift")
    film = rnd(5, n, 2 * c, scale=0.5).cuda()
    ref2 = F.silu(F.group_norm(x + s[:, :, None, None], 32, g, b, 1e-5) * (1 + film.cpu()[:, :c, None, None]) + film.cpu()[:, c:, None, None])
    y2 = ops.group_norm(a, g.cuda(), b.cuda(), cs0=ops.channel_stats(a), preshift=wide[:, 8:8 + c], film=film)
    close(nchw(y2), ref2, dt, "gn preshift + film")


@pytest.mark.parametrize("dt", DTYPES)
def test_avg_pool2x(dt):
    ops = _ops()
    x = q(rnd(1, 2, 24, 12, 20), dt)
    y = ops.avg_pool2x(nhwc(x, dt))
    close(nchw(y), F.avg_pool2d(x, 2, 2), dt, "avgpool 2x2")


def test_empty_and_cpu_inputs_fail_loudly():
    """No silent fallbacks: an empty batch or a CPU tensor is an error, reported through b200_last_error()."""
    import diffusion_models_dsdiff_b200 as P
    ops = _ops()
    wt = ops.pack_conv_weight(torch.randn(16, 16, 3, 3, device="cuda") * 0.1, torch.float16)
    with pytest.raises(P.B200Error):
        ops.conv2d([(torch.empty(0, 8, 8, 16, device="cuda", dtype=torch.float16), wt, 0, 3, 1)])
    with pytest.raises(P.B200Error):
        ops.conv2d([(torch.zeros(1, 8, 8, 16, dtype=torch.float16), wt, 0, 3, 1)])
    with pytest.raises(P.B200Error):
        ops.group_norm(torch.zeros(1, 8, 8, 32, dtype=torch.float16), torch.ones(32), torch.zeros(32))
