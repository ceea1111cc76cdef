"""GPU tests of the round-2 engine features: weight-set ("grouped along N") convolution / GroupNorm launches, the fp16
range guard, the statistics arena under forked streams, and re-packing after in-place weight updates."""
import math

import pytest
import torch
import torch.nn.functional as F

import dsdiff_oracle as O
from _util import rel_tol, synth_sd

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


def nhwc(x, dt):
    return x.permute(0, 2, 3, 1).contiguous().to(dt).cuda()


def nchw(y):
    return y.float().cpu().permute(0, 3, 1, 2).contiguous()


def close(got, ref, dt, what=""):
    err = (got - ref).abs().max().item()
    tol = STORE_EPS[dt] * max(1.0, ref.abs().max().item()) * 2.0
    assert err <= tol, f"{what}: max abs err {err:.3e} > tol {tol:.3e}"


# -------------------------------------------------------------------------------------------------
# weight sets: one launch for G same-shape problems with different weights (DS-Diff's four encoders)
# -------------------------------------------------------------------------------------------------
GROUPED_CASES = [
    # G, images per set, H, W, Cin, Cout, k, stride
    (4, 2, 8, 8, 64, 64, 3, 1),        # 8x8 maps: boxes span 2 images, CTA pairs
    (4, 8, 8, 8, 960, 960, 3, 1),      # the cfg2 long-K layer: split-K + finalize with per-set bias
    (4, 2, 16, 16, 96, 160, 3, 1),     # halo path, BN = 160
    (4, 1, 16, 16, 64, 64, 3, 2),      # stride 2, 8x8 output, ONE image per set (odd pixel tiles per set -> single CTAs)
    (4, 2, 16, 16, 64, 192, 1, 1),     # 1x1 (qkv-style)
    (3, 2, 32, 32, 32, 32, 3, 1),      # three sets
]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", GROUPED_CASES)
def test_grouped_conv_matches_separate_launches(case, dt):
    ops = _ops()
    G, npg, h, w, ci, co, k, stride = case
    xs = [q(rnd(10 + g, npg, ci, h, w), dt) for g in range(G)]
    ws = [q(rnd(20 + g, co, ci, k, k, scale=1.0 / math.sqrt(ci * k * k)), dt) for g in range(G)]
    bs = [rnd(30 + g, co, scale=0.1) for g in range(G)]
    emb = rnd(40, 1, G * co + 24).cuda()                  # per-set timestep rows live at a constant column pitch
    res = [q(rnd(50 + g, npg, co, h // stride, w // stride), dt) for g in range(G)]
    packed = [ops.pack_conv_weight(wt.cuda(), dt) for wt in ws]
    x_all = torch.cat([nhwc(x, dt) for x in xs], 0)
    r_all = torch.cat([nhwc(r, dt) for r in res], 0)
    st = torch.zeros(G * npg, co, 2, dtype=torch.float64, device="cuda")
    y = ops.conv2d([(x_all, torch.cat(packed, 0), 0, k, k // 2)], stride=stride, bias=torch.stack(bs).cuda(),
                   rowbias=emb[:, 8:8 + co], rowbias_gstride=co, residual=r_all, stats=st, wsets=G)
    torch.cuda.synchronize()
    for g in range(G):
        ref = (F.conv2d(xs[g], ws[g], bs[g], stride=stride, padding=k // 2)
               + emb[0, 8 + g * co:8 + (g + 1) * co].cpu()[None, :, None, None] + res[g])
        close(nchw(y[g * npg:(g + 1) * npg]), ref, dt, f"grouped conv {case} set {g}")
        # and bit-identical to the ordinary single-set launch of the same problem (same tiles, same K order)
        st1 = torch.zeros(npg, co, 2, dtype=torch.float64, device="cuda")
        y1 = ops.conv2d([(nhwc(xs[g], dt), packed[g], 0, k, k // 2)], stride=stride, bias=bs[g].cuda(),
                        rowbias=emb[:, 8 + g * co:8 + (g + 1) * co], residual=nhwc(res[g], dt), stats=st1)
        if ci * k * k < 8000:      # long-K 8x8 layers run split-K with a set-count-dependent split: equal to rounding only
            assert torch.equal(y1, y[g * npg:(g + 1) * npg]), f"grouped != separate launch, set {g}"
            # statistics of the same stored values; the two launches may take different epilogue flavours (fast / generic),
            # whose fp32 column sums add in a different order
            assert torch.allclose(st1, st[g * npg:(g + 1) * npg], rtol=1e-5, atol=1e-4)
        else:
            close(nchw(y1), nchw(y[g * npg:(g + 1) * npg]), dt, "grouped vs separate (split-K)")
            assert torch.allclose(st1, st[g * npg:(g + 1) * npg], rtol=2e-2, atol=1e-2 if dt == torch.float16 else 1e-1)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("film", [False, True])
def test_grouped_groupnorm_matches_separate_launches(dt, film):
    ops = _ops()
    G, npg, h, w, c = 4, 2, 8, 8, 96
    xs = [nhwc(q(rnd(60 + g, npg, c, h, w, scale=2.0), dt), dt) for g in range(G)]
    gam, bet = rnd(70, G, c).cuda(), rnd(71, G, c).cuda()
    emb = (rnd(72, 1, G * 2 * c + 8) * 0.3).cuda()
    x_all = torch.cat(xs, 0)
    cs_all = ops.channel_stats(x_all)
    kw = dict(film=emb[:, 4:4 + 2 * c], film_gstride=2 * c) if film else {}
    y = ops.group_norm(x_all, gam, bet, cs0=cs_all, wsets=G, **kw)
    for g in range(G):
        kw1 = dict(film=emb[:, 4 + g * 2 * c:4 + (g + 1) * 2 * c]) if film else {}
        y1 = ops.group_norm(xs[g], gam[g].contiguous(), bet[g].contiguous(), cs0=ops.channel_stats(xs[g]), **kw1)
        assert torch.equal(y1, y[g * npg:(g + 1) * npg])
        xr = xs[g].float().cpu().permute(0, 3, 1, 2)
        ref = F.group_norm(xr, 32, gam[g].cpu(), bet[g].cpu(), eps=1e-5)
        if film:
            row = emb[0, 4 + g * 2 * c:4 + (g + 1) * 2 * c].cpu()
            ref = ref * (1 + row[:c][None, :, None, None]) + row[c:][None, :, None, None]
        close(nchw(y1), F.silu(ref), dt, "grouped GN")


def test_silu_shared_reciprocal_accuracy():
    """SiLU with one reciprocal per four values (csrc/common.cuh silu8) over the whole range a GroupNorm + FiLM can
    produce, including the clamped tail (v < -20.8) and large positive values: within the 16-bit store rounding."""
    ops = _ops()
    c, dt = 64, torch.float16
    x = nhwc(q(rnd(80, 8, c, 16, 16), dt), dt)
    gam = torch.linspace(-60.0, 60.0, c)                      # normalised values * gamma reach +-200
    bet = torch.linspace(-30.0, 30.0, c)
    y = ops.group_norm(x, gam.cuda(), bet.cuda(), cs0=ops.channel_stats(x))
    xr = x.float().cpu().permute(0, 3, 1, 2).double()
    ref = F.silu(F.group_norm(xr, 32, gam.double(), bet.double(), eps=1e-5)).float()
    got = nchw(y)
    assert torch.isfinite(got).all()
    err = (got - ref).abs()
    assert bool((err <= 2.0 ** -10 * ref.abs() + 1e-6).all()), float((err / (ref.abs() + 1e-3)).max())
    assert float(ref.min()) < -0.2 and float(ref.max()) > 150 and float((ref.abs() < 1e-7).float().mean()) > 0.05


# -------------------------------------------------------------------------------------------------
# fp16 range guard
# -------------------------------------------------------------------------------------------------
def test_fp16_range_guard_trips_on_saturated_conv_and_not_otherwise():
    import diffusion_models_dsdiff_b200 as P
    ops = _ops()
    dt = torch.float16
    x = q(rnd(1, 2, 64, 16, 16), dt)
    wt = q(rnd(2, 64, 64, 3, 3, scale=0.05), dt)
    wp = ops.pack_conv_weight(wt.cuda(), dt)
    ops.range_counter(torch.device("cuda", 0)).zero_()                         # clean start
    y = ops.conv2d([(nhwc(x, dt), wp, 0, 3, 1)], bias=torch.full((64,), 65000.0, device="cuda"))
    ops.check_range(torch.device("cuda", 0))                                   # large but in range: no error
    assert 64000.0 < float(y.float().abs().max()) < 65504.0
    big = ops.conv2d([(nhwc(x, dt), wp, 0, 3, 1)], bias=torch.full((64,), 70000.0, device="cuda"))
    assert float(big.float().abs().max()) == 65504.0                           # the store saturated ...
    with pytest.raises(P.B200Error, match="fp16 range overflow"):
        ops.check_range(torch.device("cuda", 0))                               # ... and the guard says so
    ops.check_range(torch.device("cuda", 0))                                   # counter was reset
    # split-K + finalize path (8x8 map, long K)
    xl = q(rnd(3, 8, 960, 8, 8), dt)
    wl = ops.pack_conv_weight(q(rnd(4, 960, 960, 3, 3, scale=0.02), dt).cuda(), dt)
    ops.conv2d([(nhwc(xl, dt), wl, 0, 3, 1)], bias=torch.full((960,), 70000.0, device="cuda"))
    with pytest.raises(P.B200Error, match="fp16 range overflow"):
        ops.check_range(torch.device("cuda", 0))
    # bf16 has fp32's exponent range: nothing to count
    ops.conv2d([(nhwc(x, torch.bfloat16), ops.pack_conv_weight(wt.cuda(), torch.bfloat16), 0, 3, 1)],
               bias=torch.full((64,), 70000.0, device="cuda"))
    ops.check_range(torch.device("cuda", 0))


def test_sampling_a_checkpoint_that_overflows_fp16_raises_and_bf16_runs():
    """A checkpoint whose residual stream leaves the fp16 range must not produce silently clipped slices."""
    import diffusion_models_dsdiff_b200 as P
    from diffusion_models_dsdiff_b200 import samplers as S
    hp, sd = synth_sd("unet_small")
    sd = {k: v.clone() for k, v in sd.items()}
    sd["input_blocks.0.0.weight"] *= 3.0e5            # first conv output ~1e5: beyond +-65504
    cond = torch.clamp(rnd(2024, 2, 3, 32, 32) * 0.5, -1, 1).cuda()
    x_T = rnd(2025, 2, 1, 32, 32).cuda()
    outs = {}
    for dt in (torch.float16, torch.bfloat16):
        net = P.UNetModel(**hp, act_dtype=dt)
        net.load_state_dict(sd, strict=True)
        model = S.DDPM(net.cuda().eval(), timesteps=2000, linear_start=1e-4, linear_end=2e-2, conditioning_key="concat",
                       parameterization="v").cuda()
        run = lambda: S.DDIMSampler(model).sample(4, 2, (1, 32, 32), cond, eta=0.0, x_T=x_T, verbose=False)[0]
        if dt == torch.float16:
            with pytest.raises(P.B200Error, match="act_dtype=torch.bfloat16"):
                run()
        else:
            outs[dt] = run()
    assert torch.isfinite(outs[torch.bfloat16]).all()


# -------------------------------------------------------------------------------------------------
# statistics arena under forked encoder streams; stale packed weights
# -------------------------------------------------------------------------------------------------
def test_stats_arena_growth_with_forked_streams_is_race_free(monkeypatch):
    """The first forwards of a batch large enough to outgrow the statistics arena used to re-zero it on a side stream
    while other streams were accumulating GroupNorm sums into it.  Shrink the arena so a small DS-Diff forward overflows
    it, with the encoders on forked streams, and compare with the single-stream result bit for bit."""
    import diffusion_models_dsdiff_b200 as P
    from diffusion_models_dsdiff_b200 import ops
    hp, sd = synth_sd("ds_small")
    x = rnd(5, 4, 4, 32, 32).cuda()
    t = torch.tensor([37, 1503, 5, 900]).cuda()
    ref_net = P.DSUnetModel(**hp)
    ref_net.load_state_dict(sd, strict=True)
    ref_net = ref_net.cuda().eval()
    ref_net.encoder_streams = False
    ref = ref_net(x, t)[0]
    monkeypatch.setattr(ops.StatsArena, "MIN", 1 << 9)           # 512 doubles: overflows within the first block
    net = P.DSUnetModel(**hp)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    assert net.encoder_streams
    for i in range(3):                                           # overflow -> grow at begin() -> steady state
        out = net(x, t)[0]
        torch.cuda.synchronize()
        assert torch.equal(out, ref), f"forward {i} differs from the single-stream result"
        if i == 0:
            assert net._arena.overflows > 0                      # the shrunken arena really overflowed
    grown = net._arena.overflows
    net(x, t)
    assert net._arena.overflows == grown                         # the arena now holds a whole forward


def test_inplace_weight_update_repacks_and_recaptures():
    """EMA copy_to / optimizer steps change parameters in place: the next call must run on the new weights."""
    import diffusion_models_dsdiff_b200 as P
    from diffusion_models_dsdiff_b200 import samplers as S
    hp, sd = synth_sd("unet_small")
    net = P.UNetModel(**hp)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    x, t = rnd(11, 2, 4, 32, 32).cuda(), torch.tensor([37, 1503]).cuda()
    y0 = net(x, t)
    pw0 = net.packed()
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(1.25)                                          # in place, like LitEma.copy_to / an optimizer step
    y1 = net(x, t)
    assert net.packed() is not pw0 and not torch.equal(y0, y1)
    fresh = P.UNetModel(**hp)
    fresh.load_state_dict({k: v * 1.25 if v.is_floating_point() else v for k, v in sd.items()}, strict=True)
    assert torch.equal(fresh.cuda().eval()(x, t), y1)
    # the sampler's graph cache is keyed on the packed weights: a replayed graph must not survive the update
    model = S.DDPM(net, timesteps=2000, linear_start=1e-4, linear_end=2e-2, conditioning_key="concat",
                   parameterization="v").cuda()
    cond, x_T = rnd(12, 2, 3, 32, 32).cuda(), rnd(13, 2, 1, 32, 32).cuda()
    run = lambda: S.DDIMSampler(model).sample(4, 2, (1, 32, 32), cond, eta=0.0, x_T=x_T, verbose=False)[0]
    a = run()
    assert torch.equal(a, run())
    with torch.no_grad():
        net.out[2].weight.mul_(0.5)
    assert not torch.equal(a, run())


# -------------------------------------------------------------------------------------------------
# DS-Diff: the four encoders' small-map suffix and the FeatureDisentangle section as grouped launches
# -------------------------------------------------------------------------------------------------
def _ds_net(dtype=torch.float16):
    import diffusion_models_dsdiff_b200 as P
    from _util import synth_sd
    hp, sd = synth_sd("ds_small")
    net = P.DSUnetModel(**hp, act_dtype=dtype)
    net.load_state_dict(sd, strict=True)
    return net.cuda().eval()


@pytest.mark.parametrize("streams", [True, False])
@pytest.mark.parametrize("n", [1, 2, 3])
def test_grouped_encoders_match_separate_launches(n, streams):
    """Grouping only changes which launch computes a layer: same tiles, same K order, same stored values.  Where both
    launches also take the same epilogue flavour the network output is bit-identical (tools/group_debug.py: n = 1, 3, 4);
    otherwise the per-channel fp32 partial sums behind the GroupNorm statistics add in a different order (1e-7 relative,
    tools/epi_debug.py), which this fixture's tiny 4x4 maps (mean^2 >> variance) amplify to the level of any other
    re-ordering of 16-bit roundings — so both are held to the fp32 oracle at the parity bound and to each other at
    twice that."""
    import dsdiff_oracle as O
    from _util import golden, synth_sd
    net = _ds_net()
    net.encoder_streams = streams
    g = golden("ds_small")
    x = torch.cat([g["x"][:1] * (1.0 + 0.1 * i) for i in range(n)])
    t = g["t"][:1].repeat(n)
    assert net._group_start(x.shape[2], x.shape[3]) < len(net.input_blocks), "fixture does not exercise grouping"
    hp, sd = synth_sd("ds_small")
    with torch.no_grad():
        ref = O.dsunet_forward(sd, hp, x, t)[0]
    y_g, f_g = net(x.cuda(), t.cuda())
    net.group_max_pixels = 0
    y_s, f_s = net(x.cuda(), t.cuda())
    tol = rel_tol("ds_small.grouped")            # measured x 1.2 (tests/_util.py MEASURED_REL)
    assert O.rel_err(y_g.cpu(), ref) < tol and O.rel_err(y_s.cpu(), ref) < tol
    assert O.rel_err(y_g.cpu(), y_s.cpu()) < rel_tol("ds_small.grouped_vs_separate")
    for k in f_g:
        for a, b in zip(f_g[k], f_s[k]):
            assert a.shape == b.shape and O.rel_err(a.cpu(), b.cpu()) < rel_tol("ds_small.grouped_features_vs_separate"), k


def test_grouped_encoders_are_batch_independent():
    """A slice's result must not depend on the batch it is sampled in (per-slice reproducibility across world sizes)."""
    from _util import golden
    net = _ds_net()
    g = golden("ds_small")
    x = torch.cat([g["x"][:1] * (1.0 + 0.1 * i) for i in range(3)]).cuda()
    t = g["t"][:1].repeat(3).cuda()
    y3, _ = net(x, t)
    y1, _ = net(x[1:2].contiguous(), t[1:2])
    assert torch.equal(y3[1:2], y1)


# -------------------------------------------------------------------------------------------------
# Upsample folded into the convolution: four 2x2 sub-pixel convolutions, strided output
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 96), (1, 24, 40, 32, 32), (3, 8, 8, 128, 64)])
def test_folded_upsample_conv_matches_interpolate_then_conv(shape, dt):
    """nearest x2 -> conv3x3 pad 1 (openaimodel.py:113-120) without the upsampled intermediate; GroupNorm statistics of the
    interleaved output included."""
    import torch.nn as nn
    from diffusion_models_dsdiff_b200 import unet as U
    ops = _ops()
    n, h, w, ci, co = shape
    x = q(rnd(70, n, ci, h, w), dt)
    conv = nn.Conv2d(ci, co, 3, padding=1)
    with torch.no_grad():
        conv.weight.copy_(rnd(71, co, ci, 3, 3, scale=1.0 / math.sqrt(9 * ci)))
        conv.bias.copy_(rnd(72, co, scale=0.1))
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), conv.weight, conv.bias, padding=1)
    conv = conv.cuda()

    class PW:
        conv = {}
    arena = ops.StatsArena()
    arena.begin(torch.device("cuda"))
    cx = U.Ctx(PW(), None, arena)
    got = U.run_upsample_conv(cx, conv, U.Act(nhwc(x, dt)))
    assert tuple(got.t.shape) == (n, 2 * h, 2 * w, co)
    # one rounding of the summed taps instead of one per tap: the same operand-rounding class as the unfolded form
    close(nchw(got.t), ref, dt, f"folded upsample conv {shape}")
    yf = got.t.float()
    exact = torch.stack([yf.double().sum((1, 2)), (yf.double() ** 2).sum((1, 2))], -1)
    assert torch.allclose(got.cs, exact, rtol=1e-5, atol=1e-3)


# -------------------------------------------------------------------------------------------------
# 3x3 network head with 1..3 output channels = 1x1 convolution to 9*C tap channels + 9-point gather (b200_head_gather9)
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 1), (1, 24, 40, 96, 2), (3, 8, 8, 32, 3), (1, 64, 64, 320, 1)])
def test_split_head_matches_conv3x3_and_the_unsplit_path(shape, dt):
    """self.out = GroupNorm, SiLU, conv3x3 (openaimodel.py:898-902): the tap-split form equals F.conv2d on the same 16-bit
    operand (fp32 sums in another order) and the padded-N-tile 3x3 form it replaces."""
    import torch.nn as nn
    from diffusion_models_dsdiff_b200 import unet as U
    ops = _ops()
    n, h, w, ci, co = shape
    x = rnd(80, n, ci, h, w)
    head = nn.Sequential(nn.GroupNorm(32, ci), nn.SiLU(), nn.Conv2d(ci, co, 3, padding=1))
    with torch.no_grad():
        head[0].weight.copy_(1.0 + rnd(81, ci, scale=0.2))
        head[0].bias.copy_(rnd(82, ci, scale=0.2))
        head[2].weight.copy_(rnd(83, co, ci, 3, 3, scale=1.0 / math.sqrt(9 * ci)))
        head[2].bias.copy_(rnd(84, co, scale=0.1))
    head = head.cuda()

    def run(split):
        class PW:
            conv = {}
        arena = ops.StatsArena()
        arena.begin(torch.device("cuda"))
        old, U._SPLIT_HEAD = U._SPLIT_HEAD, split
        try:
            return U.run_out_head(U.Ctx(PW(), None, arena), head, U.Act(nhwc(x, dt)))
        finally:
            U._SPLIT_HEAD = old

    n0 = ops.L.lib().b200_launch_count()
    got = run(True)
    assert ops.L.lib().b200_launch_count() - n0 >= 3          # GroupNorm (+ statistics), 1x1 convolution, gather
    assert got.shape == (n, co, h, w) and got.dtype == torch.float32
    unsplit = run(False)
    # the conv operand both forms read: the 16-bit GroupNorm + SiLU output
    a = ops.group_norm(nhwc(x, dt), head[0].weight.float(), head[0].bias.float())
    ref = F.conv2d(nchw(a), q(head[2].weight.detach().cpu(), dt), head[2].bias.detach().cpu(), padding=1)
    scale = max(1.0, ref.abs().max().item())
    # same GroupNorm launch, same operands, fp32 sums in another order: the two forms agree to fp32 roundoff
    assert (got.cpu() - unsplit.cpu()).abs().max().item() < 2e-5 * scale
    # vs torch on the operand of a separately launched GroupNorm (its statistics pass may round an element the other way)
    close(got.cpu(), ref, dt, f"split head {shape}")
