"""SURVEY.md §8 f4 groundwork: the training-objective oracle (oracle/training_oracle.py) against outputs of the unmodified
reference's GaussianDiffusion.training_losses (tests/golden/train_*.npz, made by oracle/make_golden_train.py).  CPU only;
there is no CUDA training step yet — these goldens are what it will be held to."""
import numpy as np
import pytest
import torch

import dsdiff_oracle as O
import training_oracle as TO
from _util import golden, maxabs, synth_sd


def _close(a, b, rel=2e-5):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max()) <= rel * max(1e-6, float(b.abs().max()))


@pytest.mark.parametrize("tag,rescaled", [("mse", False), ("rescaled", True)])
def test_training_losses_oracle_matches_reference(tag, rescaled):
    g = golden(f"train_ds_film2_{tag}")
    hp, sd = synth_sd("ds_film2")
    sch = O.GdSchedule(steps=1000, noise_schedule="linear", timestep_respacing="", rescale_timesteps=True)
    model = lambda xc, tm: O.dsunet_forward(sd, hp, xc, tm)
    assert maxabs(TO.q_sample(sch, g["x0"], g["t"], g["noise"]), g["x_t"]) == 0.0
    with torch.no_grad():
        terms = TO.training_losses(sch, model, g["x0"], g["t"], g["noise"], list(g["cond"]), parameterization="v",
                                   learn_sigma=True, rescaled_mse=rescaled, disentangle="eu")
    for k in ("mse", "vb", "loss"):
        assert _close(terms[k], g[k]), (k, terms[k], g[k])
    for k in ("disen_c_s_loss", "disen_s_a_l_loss"):
        assert _close(terms[k], g[k]), (k, float(terms[k]), g[k])
    assert float(g["vb"][0]) != float(g["vb"][1])          # the t == 0 sample took the decoder-NLL branch


def test_vb_term_closed_forms():
    """KL of identical Gaussians is 0; the discretised likelihood integrates to 1 over the 256 bins."""
    sch = O.GdSchedule(steps=100, noise_schedule="linear", timestep_respacing="", rescale_timesteps=False)
    m = torch.randn(2, 1, 4, 4)
    lv = torch.randn(2, 1, 4, 4)
    assert float(TO.normal_kl(m, lv, m, lv).abs().max()) < 1e-6
    centers = torch.linspace(-1, 1, 256).view(256, 1)
    means = torch.tensor([[0.1, -0.7]]).expand(256, 2)
    logp = TO.discretized_gaussian_log_likelihood(centers.expand(256, 2), means, torch.full((256, 2), -1.5))
    assert torch.allclose(logp.exp().sum(0), torch.ones(2), atol=2e-3)
    # a model that predicts the true posterior mean and variance has zero KL for t > 0
    x0, nz = torch.rand(2, 1, 8, 8) * 2 - 1, torch.randn(2, 1, 8, 8)
    t = torch.tensor([10, 99])
    x_t = TO.q_sample(sch, x0, t, nz)
    v = TO.get_v(sch, x0, nz, t)
    frac_for_min = -torch.ones_like(v)                      # frac = 0 -> log-variance = posterior_log_variance_clipped
    vb = TO.vb_terms_bpd(sch, x0, x_t, t, torch.cat([v, frac_for_min], 1), "v")
    assert float(vb.abs().max()) < 1e-4


def test_disentangle_loss_labels_and_value():
    c_s, s_a_l = TO.disentangle_labels(2, 3, 3, 2, 2)
    assert c_s.tolist() == [[0, 0, 0, -1, -2, -3], [1, 1, 1, -1, -2, -3]]              # gaussian_diffusion.py:939
    assert s_a_l.tolist() == [[-1, -2, -3, 0, 0, 1, 1], [-1, -2, -3, 2, 2, 3, 3]]      # :953
    # two tight clusters far apart: same-label distances << different-label distances
    f = torch.zeros(1, 4, 2, 1, 1)
    f[0, 2:] = 10.0
    f[0, 1] += 0.1
    f[0, 3] += 0.1
    loss, logits, perfect = TO.disentangle_loss_eu(f, torch.tensor([[0, 0, 1, 1]]))
    d_same, d_diff = np.sqrt(2 * 0.1 ** 2) / 2, None
    same = 2 * 2 * d_same                                     # (0,1),(1,0),(2,3),(3,2)
    diff = sum(np.sqrt(2) * abs(a - b) / 2 for a in (0.0, 0.1) for b in (10.0, 10.1)) * 2
    assert abs(float(loss) - same / diff) < 1e-6
    assert perfect.tolist()[0] == [1, 1, -1, -1] and logits.shape == (4, 4)


def test_training_gradients_oracle_matches_reference():
    """Backward of the trainer's total loss (trainer_use_gaussian_diff.py:464-478) through the oracle network (torch autograd
    over the restated forward) vs the reference's own .backward(): per-parameter gradient norms for all 384 parameter
    tensors plus five full gradient tensors.  This is the target the CUDA dgrad / wgrad / GroupNorm-backward kernels will be
    held to."""
    import os
    from _util import GOLD
    g = golden("train_ds_film2_grads")
    gi = golden("train_ds_film2_mse")
    names = open(os.path.join(GOLD, "train_ds_film2_grads.names.txt")).read().split()
    hp, sd = synth_sd("ds_film2")
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    sch = O.GdSchedule(steps=1000, noise_schedule="linear", timestep_respacing="", rescale_timesteps=True)
    terms = TO.training_losses(sch, lambda xc, tm: O.dsunet_forward(sd, hp, xc, tm), gi["x0"], gi["t"], gi["noise"],
                               list(gi["cond"]), parameterization="v", learn_sigma=True, disentangle="eu")
    total = (g["lam"] * (terms["disen_c_s_loss"] + terms["disen_s_a_l_loss"]) + terms["loss"]).mean()
    assert _close(total.detach(), g["total"])
    total.backward()
    assert len(names) == g["grad_norms"].numel() == 384
    worst = 0.0
    # a bias in front of a one-channel-per-group GroupNorm has an exactly zero gradient; both sides hold ~1e-9 of
    # roundoff there, so the relative check carries an absolute floor tied to the largest gradient in the network
    floor = 1e-4 * float(g["grad_norms"].max())
    for name, ref in zip(names, g["grad_norms"].tolist()):
        got = sd[name].grad
        if ref < 0:                      # parameter unused by the forward in the reference
            assert got is None or float(got.abs().max()) == 0.0, name
            continue
        assert got is not None, name
        worst = max(worst, abs(float(got.double().norm()) - ref) / max(ref, floor))
    assert worst < 2e-4, worst
    for k in g:
        if k.startswith("g:"):
            ref = g[k]
            assert maxabs(sd[k[2:]].grad, ref) <= 2e-4 * max(floor, float(ref.abs().max())), k


@pytest.mark.parametrize("name", ["train_dit_v", "train_dit_eps_rescaled"])
def test_dit_training_step_oracle_matches_reference(name):
    """cfg5 groundwork: loss terms and gradients of ONE DiT training step — the oracle's restated forward (dit_forward) and
    objective under torch autograd vs the reference's own DiT + training_losses + .backward()
    (oracle/make_golden_train_dit.py)."""
    import os
    from _util import GOLD, manifest
    g = golden(name)
    man = manifest()[name]
    hp, sd = synth_sd(name)
    sd["pos_embed"] = g["pos_embed"]
    names = open(os.path.join(GOLD, name + ".names.txt")).read().split()
    sd = {k: (v.clone().requires_grad_(True) if k != "pos_embed" else v) for k, v in sd.items()}
    sch = O.GdSchedule(steps=1000, noise_schedule="linear", timestep_respacing="", rescale_timesteps=True)
    terms = TO.training_losses(sch, lambda xc, tm: O.dit_forward(sd, hp, xc, tm), g["x0"], g["t"], g["noise"], list(g["cond"]),
                               parameterization=man["parameterization"], learn_sigma=True, rescaled_mse=man["rescaled"])
    for k in ("mse", "vb", "loss"):
        assert _close(terms[k].detach(), g[k]), (k, terms[k], g[k])
    terms["loss"].mean().backward()
    floor = 1e-4 * float(g["grad_norms"].max())
    for nm, ref in zip(names, g["grad_norms"].tolist()):
        got = sd[nm].grad
        if ref < 0:
            assert got is None or float(got.abs().max()) == 0.0, nm
            continue
        assert abs(float(got.double().norm()) - ref) / max(ref, floor) < 2e-4, nm
    for k in g:
        if k.startswith("g:"):
            assert maxabs(sd[k[2:]].grad.reshape(g[k].shape), g[k]) <= 2e-4 * max(floor, float(g[k].abs().max())), k
