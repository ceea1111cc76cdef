"""Pin the oracle restatement against outputs of the UNMODIFIED reference (tests/golden, made by
oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import dsdiff_oracle as O
from _util import golden, synth_sd, maxabs

TOL = 2e-5      # fp32 CPU, different op fusion/order inside torch kernels only


def test_unet_small_matches_reference():
    hp, sd = synth_sd("unet_small")
    g = golden("unet_small")
    y = O.unet_forward(sd, hp, g["x"], g["t"])
    assert y.shape == g["y"].shape
    assert g["y"].abs().max() > 0.05            # fixture is not degenerate
    assert maxabs(y, g["y"]) < TOL * max(1.0, float(g["y"].abs().max()))


def test_unet_legacy_attention_film_float_t():
    hp, sd = synth_sd("unet_legacy_film")
    g = golden("unet_legacy_film")
    y = O.unet_forward(sd, hp, g["x"], g["t"])
    assert maxabs(y, g["y"]) < TOL * max(1.0, float(g["y"].abs().max()))


def test_dsunet_matches_reference_including_feature_dict():
    hp, sd = synth_sd("ds_small")
    g = golden("ds_small")
    y, f = O.dsunet_forward(sd, hp, g["x"], g["t"])
    s = max(1.0, float(g["y"].abs().max()))
    assert maxabs(y, g["y"]) < TOL * s
    assert maxabs(f["style"][0], g["style0"]) < TOL * s
    assert maxabs(f["content"][2], g["content2"]) < TOL * s
    assert maxabs(f["anatomy"][1], g["anatomy1"]) < TOL * s
    assert maxabs(f["lesion"][0], g["lesion0"]) < TOL * s
    for i in range(4):
        assert maxabs(f["n_style_content"][i], g[f"nsc{i}"]) < TOL * s


def test_dsunet_two_channel_input_zero_fills_streams():
    hp, sd = synth_sd("ds_small")
    g = golden("ds_small_2ch")
    y, _ = O.dsunet_forward(sd, hp, g["x"], g["t"])
    assert maxabs(y, g["y"]) < TOL * max(1.0, float(g["y"].abs().max()))


def test_ldm_tables_bit_exact():
    g = golden("tables_ldm_T2000_S50")
    sch = O.LdmSchedule(2000, "linear", 1e-4, 2e-2)
    for name, ref in (("betas", "betas"), ("alphas_cumprod", "alphas_cumprod"),
                      ("sqrt_alphas_cumprod", "sqrt_ac"), ("sqrt_one_minus_alphas_cumprod", "sqrt_1mac"),
                      ("posterior_log_variance_clipped", "post_logvar"),
                      ("posterior_mean_coef1", "coef1"), ("posterior_mean_coef2", "coef2")):
        assert np.array_equal(getattr(sch, name), g[ref].numpy()), name
    tab = O.LdmDdimTables(sch, 50, eta=0.3)
    assert np.array_equal(tab.timesteps, g["ddim_timesteps"].numpy())
    assert tab.timesteps[0] == 1 and tab.timesteps[-1] == 1961
    assert np.array_equal(tab.alphas, g["ddim_alphas"].numpy())
    assert np.array_equal(tab.alphas_prev, g["ddim_alphas_prev"].numpy())
    assert np.array_equal(tab.sqrt_one_minus_alphas, g["ddim_s1m"].numpy())
    assert np.array_equal(tab.sigmas, g["ddim_sigmas"].numpy())


@pytest.mark.parametrize("resp", ["50", "20"])
def test_gd_tables_bit_exact(resp):
    g = golden(f"tables_gd_T1000_S{resp}")
    sch = O.GdSchedule(1000, "linear", resp, True)
    assert sch.timestep_map == g["timestep_map"].tolist()
    assert np.array_equal(sch.betas, g["betas"].numpy())
    assert np.array_equal(sch.alphas_cumprod, g["alphas_cumprod"].numpy())
    assert np.array_equal(sch.posterior_log_variance_clipped, g["post_logvar"].numpy())
    assert np.array_equal(sch.posterior_mean_coef1, g["coef1"].numpy())
    assert np.array_equal(sch.posterior_mean_coef2, g["coef2"].numpy())
    assert np.array_equal(sch.sqrt_recipm1_alphas_cumprod, g["sqrt_recipm1"].numpy())


@pytest.mark.parametrize("param", ["v", "eps"])
@pytest.mark.parametrize("eta", [0.0, 0.7])
def test_v1_ldm_ddim_steps_and_loop(param, eta):
    g = golden(f"v1_ddim_{param}_eta{eta}")
    sch = O.LdmSchedule(2000, "linear", 1e-4, 2e-2)
    S = g["S"]
    tab = O.LdmDdimTables(sch, S, eta)
    # (1) update kernel alone: feed the reference's own model outputs, compare every step
    x = g["x_T"]
    for i in range(S):
        index = S - 1 - i
        x_prev, x0 = O.ldm_ddim_step(sch, tab, x, g["model_out"][i], index, g["noises"][i], param, True)
        assert maxabs(x_prev, g["x_inter"][i]) < 2e-6
        assert maxabs(x0, g["pred_x0"][i]) < 2e-6
        x = g["x_inter"][i]
    # (2) full loop with the oracle network
    hp, sd = synth_sd("unet_small")
    rec = []
    final = O.ldm_ddim_loop(lambda xc, t: O.unet_forward(sd, hp, xc, t), sch, S, g["x_T"], g["cond"],
                            g["noises"], eta, param, True, record=rec)
    assert maxabs(final, g["final"]) < 1e-4
    assert maxabs(rec[0]["model_out"], g["model_out"][0]) < 1e-4


@pytest.mark.parametrize("param", ["v", "eps", "x0"])
def test_v2_ldm_ancestral(param):
    g = golden(f"v2_ddpm_{param}")
    T = g["T"]
    sch = O.LdmSchedule(T, "linear", 1e-4, 2e-2)
    hp, sd = synth_sd("unet_small")
    rec = []
    final = O.ldm_ddpm_loop(lambda xc, t: O.unet_forward(sd, hp, xc, t), sch, g["x_T"], g["cond"],
                            g["noises"], param, True, record=rec)
    for i in range(T):
        assert maxabs(rec[i]["x_prev"], g["x_inter"][i]) < 1e-4, i
    assert maxabs(final, g["final"]) < 1e-4


@pytest.mark.parametrize("mode,eta", [("ddim", 0.0), ("ddim", 0.5), ("ddpm", 0.0)])
@pytest.mark.parametrize("tag", ["ls", "fixed"])
def test_v3_v4_guided_diffusion(mode, eta, tag):
    g = golden(f"v34_{mode}_{tag}_eta{eta}")
    sch = O.GdSchedule(1000, "linear", "5", True)
    assert sch.timestep_map == g["timestep_map"].tolist()
    hp, sd = synth_sd("ds_film2" if tag == "ls" else "ds_small")
    rec = []
    final = O.gd_loop(lambda xc, t: O.dsunet_forward(sd, hp, xc, t), sch, g["x_T"], g["cond"], g["noises"],
                      mode=mode, eta=eta, record=rec, parameterization="v", learn_sigma=(tag == "ls"),
                      clip_denoised=True)
    for i in range(g["S"]):
        assert maxabs(rec[i]["pred_x0"], g["pred_x0"][i]) < 1e-4, i
        assert maxabs(rec[i]["x_prev"], g["x_inter"][i]) < 1e-4, i
    assert maxabs(final, g["final"]) < 1e-4


# -------------------------------------------------------------------------------------------------
# KL autoencoder: oracle restatement vs the reference's AutoencoderKL outputs
# -------------------------------------------------------------------------------------------------
def test_autoencoder_oracle_matches_reference():
    hp, sd = synth_sd("ae_small")
    g = golden("ae_small")
    mom = O.ae_encode_moments(sd, g["x"])
    assert maxabs(mom, g["moments"]) < 2e-5
    z = O.ae_posterior_sample(g["moments"], g["noise"])
    assert maxabs(z, g["z"]) < 1e-6
    assert maxabs(O.ae_posterior_sample(g["moments"], None), g["mode"]) == 0.0
    dec = O.ae_decode(sd, g["z"])
    assert maxabs(dec, g["dec"]) < 2e-5


# -------------------------------------------------------------------------------------------------
# SpatialTransformer / cross-attention UNet
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["xattn_small", "xattn_d40"])
def test_cross_attention_unet_oracle_matches_reference(case):
    hp, sd = synth_sd(case)
    g = golden(case)
    y = O.unet_forward(sd, hp, g["x"], g["t"], context=g["context"])
    assert g["y"].abs().max() > 0.05
    assert maxabs(y, g["y"]) < TOL * max(1.0, float(g["y"].abs().max()))


# -------------------------------------------------------------------------------------------------
# DiT: the reference's DiT_models.py glue run on a restatement of timm's PatchEmbed / Attention /
# Mlp (timm itself is absent and unpinned -> that part stays "parity unpinned", see oracle/dsdiff_oracle.py)
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["dit_small", "dit_p2"])
def test_dit_oracle_matches_reference(case):
    hp, sd = synth_sd(case)
    g = golden(case)
    grid = hp["input_size"] // hp["patch_size"]
    pos = O.dit_pos_embed(hp["hidden_size"], grid)
    assert maxabs(pos, g["pos_embed"]) < 1e-6          # fixed sin-cos table, DiT_models.py:196-197, 276-326
    sd["pos_embed"] = g["pos_embed"]
    y = O.dit_forward(sd, hp, g["x"], g["t"], y=g.get("labels"), cond=g.get("cond"))
    assert y.shape == g["y_out"].shape
    assert g["y_out"].abs().max() > 0.05
    assert maxabs(y, g["y_out"]) < TOL * max(1.0, float(g["y_out"].abs().max()))


# -------------------------------------------------------------------------------------------------
# DPM-Solver++ 2M: oracle restatement vs the reference's DPMSolverSampler / DPM_Solver outputs
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case,param", [("dpmpp_v_S6", "v"), ("dpmpp_v_S12", "v"), ("dpmpp_eps_S6", "eps"),
                                        ("dpmpp_eps_S12", "eps")])
def test_dpm_solver_pp_oracle_matches_reference(case, param):
    hp, sd = synth_sd("unet_small")
    g = golden(case)
    sch = O.LdmSchedule(timesteps=2000, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2)
    solver = O.DpmSolverPP2M(torch.from_numpy(sch.betas), g["S"])
    assert solver.ns.total_N == g["total_N"]                 # numerical_clip_alpha drops the same tail
    t_in = torch.stack([r["t_input"][0] for r in solver.rows])
    assert maxabs(t_in, g["t_input"]) < 1e-3                  # on a 0..1000 scale
    rec = []
    fn = lambda x, t: O.unet_forward(sd, hp, torch.cat([x, g["cond"]], 1), t)
    out = solver.sample(fn, g["x_T"], "noise" if param == "eps" else "v", record=rec)
    for i in range(g["S"]):      # eps-parameterised random networks grow to |x| ~ 400: tolerance relative to the scale
        assert maxabs(rec[i]["x"], g["x_inter"][i]) < 2e-5 * max(1.0, float(g["x_inter"][i].abs().max())), i
    assert maxabs(out, g["final"]) < 2e-5 * max(1.0, float(g["final"].abs().max()))


def test_dpm_solver_pp_thresholding_logsnr_oracle_matches_reference():
    hp, sd = synth_sd("ds_small")
    g = golden("dpmpp_gd_thresh")
    solver = O.DpmSolverPP2M(g["betas"], g["S"], skip_type="logSNR", lower_order_final=False, thresholding=True)
    assert solver.ns.total_N == g["total_N"]
    rec = []
    fn = lambda x, t: O.dsunet_forward(sd, hp, torch.cat([x, g["cond"]], 1), t)[0]
    out = solver.sample(fn, g["x_T"], "v", record=rec)
    for i in range(g["S"]):
        assert maxabs(rec[i]["x"], g["x_inter"][i]) < 2e-4, i
    assert maxabs(out, g["final"]) < 2e-4


# -------------------------------------------------------------------------------------------------
# Where north_star's "per-step eps within 1e-3 relative" sits: the reference itself runs its convolutions / matmuls in
# TF32 on the GPU (inference/inference_2d_main.py:17 + cuDNN default), i.e. with 10-bit-mantissa operands.  Evaluated with
# the oracle (everything else exact fp32), that operand rounding ALONE moves the output by ~1e-3 relative — the same as
# the fp16 operands of the B200 path.  The GPU parity tests therefore assert 2e-3 against the pure-fp32 reference.
# -------------------------------------------------------------------------------------------------
def _tf32(v):
    i = v.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("case", ["unet_small", "ds_small"])
def test_ten_bit_mantissa_operand_floor(case):
    hp, sd = synth_sd(case)
    g = golden(case)
    run = (lambda q: O.unet_forward(sd, hp, g["x"], g["t"], opq=q)) if case == "unet_small" else \
        (lambda q: O.dsunet_forward(sd, hp, g["x"], g["t"], opq=q)[0])
    tf32 = O.rel_err(run(_tf32), g["y"])
    fp16 = O.rel_err(run(lambda v: v.half().float()), g["y"])
    assert 0.7e-3 < tf32 < 2e-3 and 0.7e-3 < fp16 < 2e-3      # measured: 0.98e-3 / 0.93e-3 (UNet), 1.24e-3 / 1.27e-3 (DS-Diff)
    assert abs(tf32 - fp16) < 0.3 * tf32                      # same mantissa width -> same floor


def test_ddim_classifier_free_guidance_oracle_matches_reference():
    """ddim.py:194-219: the network runs on [x; x] with [uncond; cond], out = u + s (c - u)."""
    hp, sd = synth_sd("unet_small")
    g = golden("v1_ddim_cfg3")
    sch = O.LdmSchedule(timesteps=2000, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2)
    S = g["S"]
    tab = O.LdmDdimTables(sch, S, 0.0)
    x = g["x_T"]
    for i in range(S):
        index = S - 1 - i
        ts = torch.full((4,), int(tab.timesteps[index]), dtype=torch.long)
        xin = torch.cat([torch.cat([x, x]), torch.cat([g["uncond"], g["cond"]])], 1)
        out2 = O.unet_forward(sd, hp, xin, ts)
        assert maxabs(out2, g["model_out2"][i]) < 2e-5 * max(1.0, float(g["model_out2"][i].abs().max()))
        u, c = out2.chunk(2)
        x, _ = O.ldm_ddim_step(sch, tab, x, u + g["scale"] * (c - u), index, g["noises"][i], "v", True)
        assert maxabs(x, g["x_inter"][i]) < 1e-4, i
    assert maxabs(x, g["final"]) < 1e-4


def test_codebook_id_head_oracle_matches_reference():
    """predict_codebook_ids: id_predictor = GroupNorm32 + 1x1 conv (openaimodel.py:903-907, 955-956)."""
    hp, sd = synth_sd("unet_codebook")
    g = golden("unet_codebook")
    y = O.unet_forward(sd, hp, g["x"], g["t"])
    assert y.shape == (2, 40, 16, 16) and g["y"].abs().max() > 0.05
    assert maxabs(y, g["y"]) < TOL * max(1.0, float(g["y"].abs().max()))


@pytest.mark.parametrize("case", ["unet_updown", "unet_seq_label"])
def test_resblock_updown_and_label_embedding_oracle_matches_reference(case):
    hp, sd = synth_sd(case)
    g = golden(case)
    y = O.unet_forward(sd, hp, g["x"], g["t"], y=g["labels"])
    assert g["y"].abs().max() > 0.05
    assert maxabs(y, g["y"]) < TOL * max(1.0, float(g["y"].abs().max()))
