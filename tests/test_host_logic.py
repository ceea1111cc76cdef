"""CPU tests of the host side: state_dict compatibility, schedule / coefficient tables (checked by
replaying the fused kernel's arithmetic in numpy fp32 against the oracle's per-step functions and the
reference-generated goldens), respacing, config plumbing, and the C-ABI surface of the built library.
No CUDA compute is called here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import dsdiff_oracle as O
from _util import golden
from _util import manifest, maxabs, synth_sd

import diffusion_models_dsdiff_b200 as P
from diffusion_models_dsdiff_b200 import _lib as L
from diffusion_models_dsdiff_b200 import samplers as S

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# -------------------------------------------------------------------------------------------------
# C ABI surface
# -------------------------------------------------------------------------------------------------
def _header_symbols():
    txt = open(os.path.join(REPO, "include", "b200diff.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    syms = _header_symbols()
    assert len(syms) >= 18
    assert os.path.exists(L.LIB_PATH), "libb200diff.so not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(L.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200diff.h but not exported"
    assert set(L.SIGNATURES) == set(syms), "ctypes table and header disagree"
    assert L.lib().b200_abi_version() == L.ABI_VERSION == 16
    assert L.lib().b200_launch_count() == 0 or True


def test_error_path_reports_text_without_gpu():
    d = L.ConvDesc()
    d.nseg = 0
    rc = L.lib().b200_conv2d_igemm(ctypes.byref(d), None)
    assert rc != 0 and b"nseg" in L.lib().b200_last_error()
    with pytest.raises(L.B200Error):
        L.check(rc, "conv")


# -------------------------------------------------------------------------------------------------
# drop-in modules: constructor keywords + state_dict keys/shapes of the reference
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["unet_small", "unet_legacy_film", "unet_updown", "unet_seq_label", "xattn_small",
                                  "unet_codebook"])
def test_unet_state_dict_matches_reference(case):
    hp, sd = synth_sd(case)
    net = P.UNetModel(**hp)
    mine = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    ref = {k: tuple(v) for k, v in manifest()[case]["shapes"].items()}
    assert mine == ref
    net.load_state_dict(sd, strict=True)
    assert all(v.dtype == torch.float32 for v in net.state_dict().values())


@pytest.mark.parametrize("case", ["ds_small", "ds_film2"])
def test_dsunet_state_dict_matches_reference(case):
    hp, sd = synth_sd(case)
    net = P.DSUnetModel(**hp)
    mine = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    ref = {k: tuple(v) for k, v in manifest()[case]["shapes"].items()}
    assert mine == ref
    net.load_state_dict(sd, strict=True)
    # the three conditioning encoders are separate parameters, not views of the first
    assert net.input_blocks_a[1][0].in_layers[2].weight.data_ptr() != net.input_blocks[1][0].in_layers[2].weight.data_ptr()


def test_instantiate_from_config_maps_reference_targets():
    hp, _ = synth_sd("unet_small")
    net = S.instantiate_from_config({"target": "ldm.modules.diffusionmodules.openaimodel.UNetModel", "params": hp})
    assert isinstance(net, P.UNetModel)
    hp2, _ = synth_sd("ds_small")
    assert isinstance(S.instantiate_from_config({"target": "UNet_DS_Diff.model.DSUnetModel", "params": hp2}),
                      P.DSUnetModel)
    with pytest.raises(KeyError):
        S.instantiate_from_config({"params": {}})


# -------------------------------------------------------------------------------------------------
# the fused update kernel's arithmetic, replayed in numpy fp32 on the host-built coefficient rows
# -------------------------------------------------------------------------------------------------
f32 = np.float32


def emulate(mode, param, c, x, m, s, z, clip=True, learned=False):
    """sampler.cu::sampler_elem, op for op, in numpy fp32 (no FMA)."""
    x, m, z = x.astype(f32), m.astype(f32), z.astype(f32)
    c = c.astype(f32)
    cl = (lambda v: np.clip(v, f32(-1), f32(1))) if clip else (lambda v: v)
    if mode == L.SAMPLER_LDM_DDIM:
        if param == L.PARAM_V:
            e = c[0] * m + c[1] * x
            x0 = c[0] * x - c[1] * m
        else:
            e = m
            x0 = (x - c[3] * e) / c[2]
        x0 = cl(x0)
        return (c[4] * x0 + c[5] * e) + (c[6] * z) * c[7], x0
    x0 = m.copy() if param == L.PARAM_X0 else c[0] * x - c[1] * m
    x0 = cl(x0)
    if mode == L.SAMPLER_LDM_DDPM:
        return (c[2] * x0 + c[3] * x) + c[4] * z, x0
    if mode == L.SAMPLER_GD_DDIM:
        eps = (c[6] * x - x0) / c[7]
        return (x0 * c[8] + c[9] * eps) + c[10] * z, x0
    mean = c[2] * x0 + c[3] * x
    if learned:
        frac = (s.astype(f32) + f32(1)) / f32(2)
        logvar = frac * c[5] + (f32(1) - frac) * c[4]
    else:
        logvar = np.full_like(x, c[4])
    return mean + (c[11] * np.exp(f32(0.5) * logvar)) * z, x0


class _Stub(torch.nn.Module):
    def forward(self, *a, **k):
        raise AssertionError("network must not run in a host-logic test")


def _ddpm(param, T=2000):
    return S.DDPM(_Stub(), timesteps=T, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2,
                  conditioning_key="concat", parameterization=param)


def test_ldm_schedule_buffers_bit_exact_vs_reference():
    g = golden("tables_ldm_T2000_S50")
    m = _ddpm("v")
    for name, ref in (("betas", "betas"), ("alphas_cumprod", "alphas_cumprod"), ("sqrt_alphas_cumprod", "sqrt_ac"),
                      ("sqrt_one_minus_alphas_cumprod", "sqrt_1mac"), ("posterior_log_variance_clipped", "post_logvar"),
                      ("posterior_mean_coef1", "coef1"), ("posterior_mean_coef2", "coef2")):
        assert torch.equal(getattr(m, name), g[ref]), name
    smp = S.DDIMSampler(m)
    smp.make_schedule(50, ddim_eta=0.3, verbose=False)
    assert np.array_equal(smp.ddim_timesteps, g["ddim_timesteps"].numpy())
    assert np.array_equal(np.asarray(smp.ddim_alphas), g["ddim_alphas"].numpy())
    assert np.array_equal(np.asarray(smp.ddim_alphas_prev), g["ddim_alphas_prev"].numpy())
    assert np.array_equal(np.asarray(smp.ddim_sigmas), g["ddim_sigmas"].numpy())
    assert np.array_equal(np.asarray(smp.ddim_sqrt_one_minus_alphas), g["ddim_s1m"].numpy())


@pytest.mark.parametrize("param", ["v", "eps"])
@pytest.mark.parametrize("eta", [0.0, 0.7])
def test_v1_coef_rows_reproduce_reference_steps(param, eta):
    g = golden(f"v1_ddim_{param}_eta{eta}")
    m = _ddpm(param)
    smp = S.DDIMSampler(m)
    n = g["S"]
    smp.make_schedule(n, ddim_eta=eta, verbose=False)
    coef, tsteps = smp._coef(False, 1.0)
    assert coef.shape == (n, L.SAMPLER_NCOEF) and tsteps == sorted(tsteps, reverse=True)
    x = g["x_T"].numpy()
    for i in range(n):
        xp, x0 = emulate(L.SAMPLER_LDM_DDIM, S._PARAM[param], coef[i].numpy(), x, g["model_out"][i].numpy(), None,
                         g["noises"][i].numpy())
        assert np.abs(xp - g["x_inter"][i].numpy()).max() < 2e-6, i
        assert np.abs(x0 - g["pred_x0"][i].numpy()).max() < 2e-6, i
        x = g["x_inter"][i].numpy()


def test_v1_original_steps_rows():
    m = _ddpm("v", T=40)
    smp = S.DDIMSampler(m)
    smp.make_schedule(8, ddim_eta=0.5, verbose=False)
    coef, tsteps = smp._coef(True, 1.0)
    assert tsteps == list(range(39, -1, -1)) and coef.shape[0] == 40
    assert torch.isfinite(coef[:-1]).all()


@pytest.mark.parametrize("param", ["v", "eps", "x0"])
def test_v2_coef_rows_match_oracle_step(param):
    T = 12
    m = _ddpm(param, T=T)
    sch = O.LdmSchedule(T, "linear", 1e-4, 2e-2)
    seq = list(reversed(range(T)))
    coef = m._ddpm_coef(seq)
    rs = np.random.RandomState(0)
    for i, t in enumerate(seq):
        x, mo, z = (torch.from_numpy(rs.standard_normal((2, 1, 8, 8)).astype(f32)) for _ in range(3))
        ref, ref0 = O.ldm_ddpm_step(sch, x, mo, t, z, param, True)
        xp, x0 = emulate(L.SAMPLER_LDM_DDPM, S._PARAM[param], coef[i].numpy(), x.numpy(), mo.numpy(), None, z.numpy())
        assert np.abs(xp - ref.numpy()).max() < 2e-6 and np.abs(x0 - ref0.numpy()).max() < 2e-6


@pytest.mark.parametrize("mode", ["ddim", "ddpm"])
@pytest.mark.parametrize("learn_sigma", [True, False])
@pytest.mark.parametrize("eta", [0.0, 0.5])
def test_v3_v4_coef_rows_match_oracle_step(mode, learn_sigma, eta):
    d = S.create_gaussian_diffusion(steps=1000, learn_sigma=learn_sigma, timestep_respacing="5",
                                    rescale_timesteps=True, parameterization="v")
    sch = O.GdSchedule(1000, "linear", "5", True)
    assert d.timestep_map == sch.timestep_map
    assert [d._model_timestep(t) for t in range(5)] == [sch.model_timestep(t) for t in range(5)]
    coef = d._coef(eta)
    rs = np.random.RandomState(1)
    kmode = L.SAMPLER_GD_DDIM if mode == "ddim" else L.SAMPLER_GD_DDPM
    for i, t in enumerate(reversed(range(5))):
        x, z = (torch.from_numpy(rs.standard_normal((2, 1, 8, 8)).astype(f32)) for _ in range(2))
        mo = torch.from_numpy(rs.standard_normal((2, 2 if learn_sigma else 1, 8, 8)).astype(f32))
        if mode == "ddim":
            ref, ref0 = O.gd_ddim_step(sch, x, mo, t, z, eta=eta, parameterization="v", learn_sigma=learn_sigma)
        else:
            ref, ref0 = O.gd_ddpm_step(sch, x, mo, t, z, parameterization="v", learn_sigma=learn_sigma)
        mean_ch = mo[:, :1].numpy()
        s_ch = mo[:, 1:].numpy() if learn_sigma else None
        xp, x0 = emulate(kmode, L.PARAM_V, coef[i].numpy(), x.numpy(), mean_ch, s_ch, z.numpy(), learned=learn_sigma)
        assert np.abs(xp - ref.numpy()).max() < 3e-6, (i, t)
        assert np.abs(x0 - ref0.numpy()).max() < 2e-6


@pytest.mark.parametrize("resp", ["50", "20"])
def test_gd_respacing_tables_bit_exact_vs_reference(resp):
    g = golden(f"tables_gd_T1000_S{resp}")
    d = S.create_gaussian_diffusion(steps=1000, learn_sigma=True, timestep_respacing=resp, rescale_timesteps=True,
                                    parameterization="v")
    assert d.timestep_map == g["timestep_map"].tolist()
    assert np.array_equal(d.betas, g["betas"].numpy())
    assert np.array_equal(d.alphas_cumprod, g["alphas_cumprod"].numpy())
    assert np.array_equal(d.posterior_log_variance_clipped, g["post_logvar"].numpy())
    assert np.array_equal(d.posterior_mean_coef1, g["coef1"].numpy())
    assert np.array_equal(d.posterior_mean_coef2, g["coef2"].numpy())
    assert np.array_equal(d.sqrt_recipm1_alphas_cumprod, g["sqrt_recipm1"].numpy())


def test_space_timesteps_edge_cases():
    assert S.space_timesteps(1000, "ddim50") == set(range(0, 1000, 20))
    assert sorted(S.space_timesteps(10, [10])) == list(range(10))
    assert sorted(S.space_timesteps(300, "10,5")) == sorted(O.gd_space_timesteps(300, "10,5"))
    with pytest.raises(ValueError):
        S.space_timesteps(10, [11])


def test_unsupported_paths_fail_loudly():
    m = _ddpm("v", T=20)
    smp = S.DDIMSampler(m)
    with pytest.raises(NotImplementedError):
        smp.sample(5, 1, (1, 8, 8), None, mask=torch.ones(1), verbose=False)
    hp, _ = synth_sd("unet_small")
    with pytest.raises(NotImplementedError):
        P.UNetModel(**{**hp, "dims": 3})               # 3D UNet: out of scope
    with pytest.raises(AssertionError):             # the reference's own ctor check (openaimodel.py:630-637)
        P.UNetModel(**{**hp, "use_spatial_transformer": True})


# -------------------------------------------------------------------------------------------------
# DPM-Solver++ host tables (product code) vs the oracle restatement and the reference-generated goldens
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("steps,skip,lof", [(6, "time_uniform", True), (12, "time_uniform", True), (20, "logSNR", False)])
def test_dpm_solver_tables_match_oracle(steps, skip, lof):
    import torch
    import dsdiff_oracle as O
    from diffusion_models_dsdiff_b200 import dpm_solver as D
    sch = O.LdmSchedule(timesteps=2000, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2)
    betas = torch.from_numpy(sch.betas)
    ns = D.NoiseScheduleVP("discrete", betas=betas)
    ref = O.DpmSolverPP2M(betas, steps, skip_type=skip, lower_order_final=lof)
    assert ns.total_N == ref.ns.total_N < 2000          # the -5.1 logSNR clip drops the tail of the T=2000 schedule
    coef, t_in, orders = D.multistep_tables(ns, steps, skip, 2, lof)
    assert orders == ref.orders
    for k, r in enumerate(ref.rows):
        got = coef[k]
        want = [r["alpha_s"], r["sigma_s"], r["A"], r["B"], r["inv_r0"], r["halfB"]]
        for j, w in enumerate(want):
            assert abs(float(got[j]) - float(w)) <= 1e-6 * max(1.0, abs(float(w))), (k, j)
        assert float(got[6]) == r["order"]
        assert abs(float(t_in[k]) - float(r["t_input"])) < 1e-3


def test_dpm_solver_model_timesteps_match_reference_golden():
    import torch
    import dsdiff_oracle as O
    from _util import golden
    from diffusion_models_dsdiff_b200 import dpm_solver as D
    sch = O.LdmSchedule(timesteps=2000, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2)
    for S in (6, 12):
        g = golden(f"dpmpp_v_S{S}")
        ns = D.NoiseScheduleVP("discrete", betas=torch.from_numpy(sch.betas))
        assert ns.total_N == g["total_N"]
        _, t_in, _ = D.multistep_tables(ns, S)
        assert float((t_in - g["t_input"]).abs().max()) < 1e-3


# -------------------------------------------------------------------------------------------------
# evaluation metrics restated for the parity gates (oracle/metrics.py): closed-form checks
# -------------------------------------------------------------------------------------------------
def test_metrics_closed_forms():
    import numpy as np
    import metrics as M
    rs = np.random.RandomState(0)
    gt = rs.rand(3, 192, 192)
    assert abs(M.psnr(gt, gt + 0.01) - 10 * np.log10((gt.max() - gt.min()) ** 2 / 1e-4)) < 1e-9
    assert abs(M.ms_ssim(gt.copy(), gt.copy()) - 1.0) < 1e-12                 # identical volumes
    noisy = gt + 0.2 * rs.randn(*gt.shape)
    v = M.ms_ssim(gt.copy(), noisy)
    assert 0.0 < v < 0.999
    assert abs(M.ms_ssim(gt.copy(), noisy * 3.0 + 7.0) - v) < 1e-9            # scale12bit removes affine intensity maps
    s = M.scale12bit(gt)
    assert abs(s.mean() - 2048.0) < 1e-6 and abs(s.std() - 400.0) < 1e-6
    x = np.array([-2.0, -1.0, 0.0, 1.0], dtype=np.float32)
    assert np.allclose(M.inverse_normalize(x, 10.0, 30.0), [10.0, 10.0, 20.0, 30.0])
    suv = M.pet_to_suv(np.array([1000.0]), weight_kg=70.0, total_dose_bq=3.7e8, half_life_s=6586.2, decay_time_s=6586.2)
    assert abs(float(suv[0]) - 1000.0 * 1000.0 * 70.0 / (3.7e8 / 2)) < 1e-3


@pytest.mark.parametrize("case", ["dit_small", "dit_p2"])
def test_dit_state_dict_matches_reference(case):
    """Key names / shapes of the reference's DiT (with timm's sub-module names) so its checkpoints load strictly."""
    import diffusion_models_dsdiff_b200 as P
    from _util import manifest
    m = manifest()[case]
    net = P.DiT(**m["hp"])
    got = {k: list(v.shape) for k, v in net.state_dict().items() if k != "pos_embed"}
    assert got == {k: list(v) for k, v in m["shapes"].items()}
    assert "pos_embed" in net.state_dict()
    with pytest.raises(P.B200Error):
        net(torch.zeros(1, m["hp"]["in_channels"], m["hp"]["input_size"], m["hp"]["input_size"]), torch.zeros(1))


def test_diffusion_wrapper_dispatches_every_conditioning_key():
    """DiffusionWrapper.forward (ddpm.py:1328-1365): which tensors reach the denoiser as x / context / y for each key."""
    class Probe(torch.nn.Module):
        def forward(self, x, t, context=None, y=None):
            self.seen = (x, context, y)
            return x[:, :1]

    x, t = torch.zeros(2, 1, 4, 4), torch.zeros(2)
    cc, ca = [torch.ones(2, 3, 4, 4)], [torch.ones(2, 5, 8), 2 * torch.ones(2, 2, 8)]
    adm = torch.full((2, 6), 3.0)
    for key, kw, xch, ctx_len, y in (
            (None, {}, 1, None, None),
            ("concat", {"c_concat": cc}, 4, None, None),
            ("crossattn", {"c_crossattn": ca}, 1, 7, None),
            ("hybrid", {"c_concat": cc, "c_crossattn": ca}, 4, 7, None),
            ("hybrid-adm", {"c_concat": cc, "c_crossattn": ca, "c_adm": adm}, 4, 7, adm),
            ("crossattn-adm", {"c_crossattn": ca, "c_adm": adm}, 1, 7, adm),
            ("crossattn-adm", {"c_crossattn": [None], "c_adm": adm}, 1, None, adm),
            ("adm", {"c_crossattn": [adm]}, 1, None, adm)):
        probe = Probe()
        out = S.DiffusionWrapper(probe, key)(x, t, **kw)
        sx, sc, sy = probe.seen
        assert out.shape == (2, 1, 4, 4) and sx.shape[1] == xch, key
        assert (sc is None) if ctx_len is None else (sc.shape == (2, ctx_len, 8)), key
        assert (sy is None) if y is None else torch.equal(sy, y), key
    with pytest.raises(AssertionError):
        S.DiffusionWrapper(Probe(), "hybrid-adm")(x, t, c_concat=cc, c_crossattn=ca)
    with pytest.raises(AssertionError):
        S.DiffusionWrapper(Probe(), "bogus")


def test_every_runtime_switch_is_documented():
    """Each B200DIFF_* environment variable read by the library or the package is described in DESIGN.md."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "diffusion_models_dsdiff_b200")
    found = set()
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".py", ".h")):
                found |= set(re.findall(r"B200DIFF_[A-Z0-9_]+", open(os.path.join(d, f), errors="ignore").read()))
    found.discard("B200DIFF_DEBUG_TOGGLES")          # compile-time macro, described next to B200DIFF_SKIP_EPI
    design = open(os.path.join(root, "DESIGN.md")).read()
    missing = sorted(v for v in found if v not in design)
    assert not missing, missing


def test_split_head_weight_layout_reproduces_conv3x3():
    """unet.split_head_weight + the 9-point gather b200_head_gather9 performs (restated in torch) == F.conv2d(pad 1)."""
    import torch.nn.functional as F
    from diffusion_models_dsdiff_b200.unet import split_head_weight
    g = torch.Generator().manual_seed(4)
    for c, cin, h, w in ((1, 8, 5, 7), (2, 16, 6, 6), (3, 8, 4, 9)):
        x = torch.randn(2, cin, h, w, generator=g, dtype=torch.float64)
        wt = torch.randn(c, cin, 3, 3, generator=g, dtype=torch.float64)
        b = torch.randn(c, generator=g, dtype=torch.float64)
        taps = F.conv2d(x, split_head_weight(wt)[:, :, None, None])               # [N, 9C, H, W]: the 1x1 convolution
        tp = F.pad(taps, (1, 1, 1, 1))
        out = b.view(1, c, 1, 1).expand(2, c, h, w).clone()
        for r in range(3):
            for s in range(3):
                t = r * 3 + s
                out += tp[:, t * c:(t + 1) * c, r:r + h, s:s + w]                # P[n, h + r - 1, w + s - 1, tap * C + co]
        assert torch.allclose(out, F.conv2d(x, wt, b, padding=1), atol=1e-12)
