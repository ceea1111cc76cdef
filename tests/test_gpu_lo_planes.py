"""GPU tests of the hi + lo activation planes (b200_conv_desc.out_lo / residual_lo, the low-plane inputs of
b200_groupnorm_apply_cs and b200_add4_scale_stats, unet.LoPolicy): selected stored activations keep the fp32 value the
reference keeps (a 16-bit plane plus the 16-bit remainder) while tensor-core operand reads keep using the 16-bit plane."""
import math

import pytest
import torch
import torch.nn.functional as F

import dsdiff_oracle as O

pytestmark = pytest.mark.gpu

DTYPES = [torch.float16, torch.bfloat16]
EPS = {torch.float16: 2.0 ** -11, torch.bfloat16: 2.0 ** -8}       # unit roundoff of one stored plane
EPS_LO = {torch.float16: 2.0 ** -21, torch.bfloat16: 2.0 ** -15}   # of a hi + lo pair


def _ops():
    from diffusion_models_dsdiff_b200 import ops
    return ops


def rnd(seed, *shape, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def q(x, dt):
    return x.to(dt).float()


def nhwc(x, dt):
    return x.permute(0, 2, 3, 1).contiguous().to(dt).cuda()


def nchw(y):
    return y.float().cpu().permute(0, 3, 1, 2).contiguous()


def split(x, dt):
    """fp32 NCHW -> (hi, lo) NHWC planes on the GPU and the fp32 value they carry."""
    hi = x.to(dt)
    lo = (x - hi.float()).to(dt)
    return nhwc(hi.float(), dt), nhwc(lo.float(), dt), hi.float() + lo.float()


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", [
    # N, H, W, Cin, Cout, k, residual
    (2, 64, 64, 64, 64, 3, True),       # fast epilogue (Cout % 16 == 0, full tiles), CTA pairs
    (1, 64, 64, 32, 320, 3, True),      # two N tiles per pixel tile
    (2, 48, 40, 64, 24, 3, True),       # generic epilogue (ragged tiles, Cout % 16 != 0)
    (2, 64, 64, 64, 64, 1, False),      # 1x1, no residual
])
def test_conv_low_planes_carry_the_fp32_result(case, dt):
    ops = _ops()
    n, h, w, ci, co, k, with_res = case
    x = q(rnd(1, n, ci, h, w), dt)
    wt = q(rnd(2, co, ci, k, k, scale=1.0 / math.sqrt(k * k * ci)), dt)
    bias = rnd(3, co, scale=0.1)
    ref = F.conv2d(x.double(), wt.double(), bias.double(), padding=k // 2)
    kw = {}
    if with_res:
        r_hi, r_lo, r_val = split(rnd(4, n, co, h, w, scale=3.0), dt)
        ref = ref + r_val.double()
        kw = {"residual": r_hi, "residual_lo": r_lo}
    wp = ops.pack_conv_weight(wt.cuda(), dt)
    st = torch.zeros((n, co, 2), dtype=torch.float64, device="cuda")
    hi, lo = ops.conv2d([(nhwc(x, dt), wp, 0, k, k // 2)], bias=bias.cuda(), stats=st, out_lo=True, **kw)
    plain = ops.conv2d([(nhwc(x, dt), wp, 0, k, k // 2)], bias=bias.cuda(), **kw)
    assert torch.equal(hi, plain), "the 16-bit plane must not depend on whether the low plane is written"
    scale = max(1.0, ref.abs().max().item())
    err_hi = (nchw(hi).double() - ref).abs().max().item()
    err_pair = (nchw(hi).double() + nchw(lo).double() - ref).abs().max().item()
    # fp32 accumulation over K = k*k*ci terms + the pair's own rounding
    assert err_pair <= (EPS_LO[dt] + 2e-6) * scale * 4, (err_pair, err_hi)
    assert err_pair < err_hi / 20, (err_pair, err_hi)
    # GroupNorm statistics stay those of the 16-bit plane
    yf = hi.double()
    exact = torch.stack([yf.sum((1, 2)), (yf ** 2).sum((1, 2))], -1)
    assert torch.allclose(st, exact, rtol=1e-5, atol=1e-3)
    if with_res:        # without the residual's low plane the result moves by about one rounding of the residual
        hi2, lo2 = ops.conv2d([(nhwc(x, dt), wp, 0, k, k // 2)], bias=bias.cuda(), out_lo=True, residual=r_hi)
        err2 = (nchw(hi2).double() + nchw(lo2).double() - ref).abs().max().item()
        assert err2 > 20 * err_pair


@pytest.mark.parametrize("dt", DTYPES)
def test_first_conv_low_plane(dt):
    ops = _ops()
    x = rnd(5, 2, 1, 64, 64)
    wt = rnd(6, 32, 1, 3, 3, scale=0.3)
    bias = rnd(7, 32, scale=0.1)
    ref = F.conv2d(x.double(), wt.double(), bias.double(), padding=1)
    hi, lo = ops.conv3x3_in(x.cuda(), wt.cuda(), bias.cuda(), dt, want_lo=True)
    assert torch.equal(hi, ops.conv3x3_in(x.cuda(), wt.cuda(), bias.cuda(), dt))
    scale = max(1.0, ref.abs().max().item())
    assert (nchw(hi).double() + nchw(lo).double() - ref).abs().max().item() <= (EPS_LO[dt] + 1e-6) * scale * 4


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("two", [False, True])
def test_groupnorm_reads_the_low_planes(dt, two):
    """Values near 1000 with sigma 0.5: the 16-bit plane alone is quantised to steps of 0.5 (fp16) / 4 (bf16), so the
    normalised result is only right when the low plane is added."""
    ops = _ops()
    n, c, h, w = 2, 64, 32, 32
    x0_hi, x0_lo, x0 = split(1000.0 + rnd(8, n, c, h, w, scale=0.5), dt)
    srcs, los, vals = [x0_hi], [x0_lo], [x0]
    if two:
        x1_hi, x1_lo, x1 = split(-500.0 + rnd(9, n, c, h, w, scale=0.25), dt)
        srcs.append(x1_hi); los.append(x1_lo); vals.append(x1)
    gam, bet = 1.0 + 0.1 * rnd(10, c * len(srcs)), 0.1 * rnd(11, c * len(srcs))
    # statistics of the 16-bit planes (what the producers' epilogues give), summed exactly here: the kernels' fp32 partial
    # sums cannot resolve a variance 4e6 times smaller than mean^2, which is not what this test is about
    def exact_cs(s):
        d = s.double()
        return torch.stack([d.sum((1, 2)), (d * d).sum((1, 2))], -1).contiguous()
    cs = [exact_cs(s) for s in srcs]
    kw = dict(x1=srcs[1], cs1=cs[1]) if two else {}
    got = ops.group_norm(srcs[0], gam.cuda(), bet.cuda(), groups=32, cs0=cs[0], x0_lo=los[0],
                         x1_lo=los[1] if two else None, **kw)
    no_lo = ops.group_norm(srcs[0], gam.cuda(), bet.cuda(), groups=32, cs0=cs[0], **kw)
    val = torch.cat(vals, 1).double()
    hi_only = torch.cat([nchw(s) for s in srcs], 1).double()
    g = hi_only.reshape(n, 32, -1)
    mean, var = g.mean(-1), g.var(-1, unbiased=False)
    cpg = val.shape[1] // 32
    mean_c, rstd_c = mean.repeat_interleave(cpg, 1)[:, :, None, None], (var + 1e-5).rsqrt().repeat_interleave(cpg, 1)[:, :, None, None]
    ref = F.silu((val - mean_c) * rstd_c * gam.double()[None, :, None, None] + bet.double()[None, :, None, None])
    err = (nchw(got).double() - ref).abs().max().item()
    err_no = (nchw(no_lo).double() - ref).abs().max().item()
    assert err <= 4 * EPS[dt] * max(1.0, ref.abs().max().item()) + 2e-3, err
    assert err_no > 10 * err, (err_no, err)


@pytest.mark.parametrize("dt", DTYPES)
def test_skip_mean_with_low_planes(dt):
    ops = _ops()
    n, c, h, w = 2, 64, 32, 32
    parts = [split(rnd(20 + i, n, c, h, w, scale=2.0), dt) for i in range(4)]
    ref = sum(p[2].double() for p in parts) / 4
    st = torch.zeros((n, c, 2), dtype=torch.float64, device="cuda")
    # two sources with a low plane, two without (the x stream's first conv vs a plain tensor)
    los = [parts[0][1], None, parts[2][1], None]
    ref_mixed = (parts[0][2].double() + nchw(parts[1][0]).double() + parts[2][2].double() + nchw(parts[3][0]).double()) / 4
    hi, lo = ops.add4_scale_stats(*(p[0] for p in parts), 0.25, st, los=los, want_lo=True)
    err = (nchw(hi).double() + nchw(lo).double() - ref_mixed).abs().max().item()
    assert err <= (EPS_LO[dt] + 1e-6) * 8, err
    yf = hi.double()
    assert torch.allclose(st, torch.stack([yf.sum((1, 2)), (yf ** 2).sum((1, 2))], -1), rtol=1e-5, atol=1e-3)
    st2 = torch.zeros_like(st)
    hi_all = ops.add4_scale_stats(*(p[0] for p in parts), 0.25, st2, los=[p[1] for p in parts])
    assert (nchw(hi_all).double() - ref).abs().max().item() <= 2 * EPS[dt] * max(1.0, ref.abs().max().item())
    st3 = torch.zeros_like(st)
    plain = ops.add4_scale_stats(*(p[0] for p in parts), 0.25, st3)
    assert plain.shape == hi.shape


@pytest.mark.parametrize("kinds", [("in", "dec_rb", "dec_conv1"),
                                   ("in", "enc_rb", "enc_conv1", "dec_rb", "dec_conv1", "mean", "up")])
def test_dsunet_with_low_planes_matches_the_oracle(kinds):
    """The small DS-Diff fixture on a 64x64 input with the low planes switched on for every map of >= 32x32 pixels (the
    shipped default only takes 256x256 maps): same network function, held to the fp32 oracle at the parity bound; the
    low-plane launches really happen; hoisted conditioning and the CUDA-graph step engine see the same planes."""
    import diffusion_models_dsdiff_b200 as P
    from diffusion_models_dsdiff_b200 import ops
    from diffusion_models_dsdiff_b200.unet import LoPolicy
    from _util import rel_tol, synth_sd
    hp, sd = synth_sd("ds_small")
    net = P.DSUnetModel(**hp).cuda().eval()
    net.load_state_dict(sd, strict=True)
    assert net.lo_policy is not None and net.lo_policy.min_pixels == 256 * 256      # the shipped default
    x = rnd(30, 2, 4, 64, 64)
    t = torch.tensor([37, 412])
    with torch.no_grad():
        ref = O.dsunet_forward(sd, hp, x, t)[0]
    net.lo_policy = None
    y0, _ = net(x.cuda(), t.cuda())
    net.lo_policy = LoPolicy(kinds, min_pixels=32 * 32)
    timer = ops.KernelTimer()
    ops.set_timer(timer)
    try:
        y1, _ = net(x.cuda(), t.cuda())
    finally:
        ops.set_timer(None)
    details = [d for _, _, _, _, _, d in timer.records if d]
    assert any(d.endswith("+lo") for d in details), "no launch wrote / read a low plane"
    e0, e1 = O.rel_err(y0.cpu(), ref), O.rel_err(y1.cpu(), ref)
    assert e1 < rel_tol("ds_small.lo_planes") and e0 < rel_tol("ds_small.lo_planes"), (e0, e1)   # measured x 1.2
    assert e1 <= e0 * 1.1, (e0, e1)         # extra stored precision never costs accuracy
    # hoisted conditioning prefix (persistent first-conv planes) gives the same function
    cond = x[:, 1:].cuda()
    assert net.prepare_conditioning([cond])
    y2, _ = net(x.cuda(), t.cuda(), cond_cached=True)
    # the cached conv1 output is rounded BEFORE the bias + timestep term is added (inside the next GroupNorm): a different
    # rounding point, the same bound (tests/test_gpu_model.py::test_hoisted_conditioning_prefix...)
    assert O.rel_err(y2.cpu(), ref) < rel_tol("ds_small.lo_planes_hoisted") and O.rel_err(y2.cpu(), y1.cpu()) < rel_tol("ds_small.lo_planes_hoisted")
