"""The sampler seam against the REFERENCE's own host objects.

The reference's trainers build ``DSDiffModel(DDPM)`` / ``TryTrainerDiffusion`` (holding the reference's own
``DiffusionWrapper``) and hand them to ``DDIMSampler(self)`` / ``DPMSolverSampler(self)`` / ``sample_fn(self.model, ...)``
(trainers/trainer_ds_diff.py:650-755, trainer_use_gaussian_diff.py:586-623).  With the yaml ``target`` pointing at the B200
network and the sampler imports swapped, those UNMODIFIED trainer classes must drive the B200 path:

* CPU (``-m "not gpu"``): up to the first kernel call, which has to fail loudly (no CPU path);
* GPU: end to end, against the same trainer class running the reference's own network + samplers in strict fp32 on the same
  seeds — the reference itself is the oracle here (byte-compiled ``oracle/_ref`` on the GPU box).
"""
import contextlib
import io

import pytest
import torch

import dsdiff_oracle as O
import _refhost as R
import diffusion_models_dsdiff_b200 as P
from diffusion_models_dsdiff_b200 import samplers as S

needs_ref = pytest.mark.skipif(not R.available(), reason="reference not importable (run oracle/build_ref.py)")


def _load_synth(net):
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(O.synth_state_dict(shapes), strict=True)


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


# ---------------------------------------------------------------------------------------------
# CPU: the reference's trainers reach the first kernel call of the B200 path
# ---------------------------------------------------------------------------------------------
@needs_ref
@pytest.mark.parametrize("sampler", ["ddim", "dpm", "ddpm"])
def test_reference_dsdiffmodel_reaches_the_b200_kernels(sampler):
    m = R.ds_diff_model(R.B200_TARGETS["ds"], R.DS_SMALL, sampler=sampler)
    assert type(m).__module__ == "trainers.trainer_ds_diff"
    assert type(m.model).__module__ == "ldm.models.diffusion.ddpm"          # the REFERENCE's DiffusionWrapper
    assert isinstance(m.model.diffusion_model, P.DSUnetModel)
    if sampler == "ddpm":
        # the ancestral loop is a method of the reference class itself (trainer_ds_diff.py:698-738); it calls the network
        # through apply_model -> DiffusionWrapper.forward, which reaches the engine's first kernel
        m.num_timesteps = 3
    with _quiet(), pytest.raises(P.B200Error, match="no CPU path"):
        m.predict_step(R.batch(), 0)


@needs_ref
@pytest.mark.parametrize("sampler", ["ddim", "dpm", "ddpm"])
def test_reference_gaussian_trainer_reaches_the_b200_kernels(sampler):
    g = R.gaussian_trainer(R.B200_TARGETS["ds"], R.DS_FILM2, sampler=sampler)
    assert type(g).__module__ == "trainers.trainer_use_gaussian_diff"
    assert type(g.model).__module__ == "ldm.models.diffusion.ddpm"
    assert isinstance(g.diffusion_for_val, S.SpacedDiffusion) and g.diffusion_for_val.num_timesteps == 5
    with _quiet(), pytest.raises(P.B200Error, match="no CPU path"):
        g.predict_step(R.batch(), 0)


@needs_ref
def test_reference_network_is_rejected_loudly():
    """A host whose denoiser is still the reference's PyTorch network must not be sampled through some eager fallback."""
    m = R.ds_diff_model(R.REF_TARGETS["ds"], R.DS_SMALL)
    with _quiet(), pytest.raises(P.B200Error, match="not a B200 engine network"):
        m.predict_step(R.batch(), 0)
    g = R.gaussian_trainer(R.REF_TARGETS["ds"], R.DS_FILM2)
    with _quiet(), pytest.raises(P.B200Error, match="not a B200 engine network"):
        g.predict_step(R.batch(), 0)


class DuckHost:
    """Exactly the attribute set ``DDIMSampler`` reads from its model (ldm/models/diffusion/ddim.py:12-23, 25-55, 188-261):
    schedule buffers, ``parameterization``, ``device``, ``apply_model`` and ``model`` (a DiffusionWrapper-shaped object:
    ``diffusion_model`` + ``conditioning_key``).  Not an nn.Module, none of this package's private helpers."""

    class Wrapper:
        def __init__(self, net, key):
            self.diffusion_model, self.conditioning_key = net, key

    def __init__(self, ddpm: "S.DDPM"):
        for name in ("num_timesteps", "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                     "sqrt_one_minus_alphas_cumprod", "parameterization"):
            setattr(self, name, getattr(ddpm, name))
        self.device = ddpm.betas.device
        self.model = DuckHost.Wrapper(ddpm.model.diffusion_model, ddpm.model.conditioning_key)

    def apply_model(self, x, t, cond):
        xc = torch.cat([x] + (cond if isinstance(cond, list) else [cond]), dim=1)
        out = self.model.diffusion_model(xc, t)
        return out[0] if isinstance(out, tuple) else out


def test_duck_typed_host_reaches_the_b200_kernels():
    net = P.UNetModel(**R.UNET_SMALL)
    ddpm = S.DDPM(net, timesteps=2000, linear_start=1e-4, linear_end=2e-2, conditioning_key="concat", parameterization="v")
    host = DuckHost(ddpm)
    assert not any(hasattr(host, n) for n in ("_cond_list", "_context", "_model_fn", "_graph_cache_args"))
    with _quiet(), pytest.raises(P.B200Error, match="no CPU path"):
        S.DDIMSampler(host).sample(5, 2, (1, 32, 32), torch.zeros(2, 3, 32, 32), eta=0.0, verbose=False)
    with _quiet(), pytest.raises(P.B200Error, match="no CPU path"):
        P.DPMSolverSampler(host).sample(5, 2, (1, 32, 32), torch.zeros(2, 3, 32, 32))


# ---------------------------------------------------------------------------------------------
# GPU: end to end
# ---------------------------------------------------------------------------------------------
def _strict_fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")


def _cuda_batch():
    b = R.batch()
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in b.items()}


@pytest.mark.gpu
def test_duck_typed_host_matches_package_host_bitwise():
    net = P.UNetModel(**R.UNET_SMALL)
    _load_synth(net)
    ddpm = S.DDPM(net.cuda().eval(), timesteps=2000, linear_start=1e-4, linear_end=2e-2, conditioning_key="concat",
                  parameterization="v").cuda()
    cond = _cuda_batch()["image"]
    x_T = torch.randn(2, 1, 32, 32, generator=torch.Generator().manual_seed(3)).cuda()
    with _quiet():
        a, ia = S.DDIMSampler(ddpm).sample(5, 2, (1, 32, 32), cond, eta=0.0, x_T=x_T, verbose=False)
        b, ib = S.DDIMSampler(DuckHost(ddpm)).sample(5, 2, (1, 32, 32), cond, eta=0.0, x_T=x_T, verbose=False)
        c, _ = P.DPMSolverSampler(ddpm).sample(5, 2, (1, 32, 32), cond, x_T=x_T)
        d, _ = P.DPMSolverSampler(DuckHost(ddpm)).sample(5, 2, (1, 32, 32), cond, x_T=x_T)
    assert torch.equal(a, b) and torch.equal(c, d)
    assert len(ia["x_inter"]) == len(ib["x_inter"])


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("sampler,tol", [("ddim", 1e-2), ("dpm", 1e-2), ("ddpm", 1e-2)])
def test_reference_dsdiffmodel_on_b200_path_vs_reference_itself(sampler, tol):
    """predict_step of the reference's DSDiffModel: B200 network + B200 samplers vs the reference's own network + samplers
    (strict fp32 eager on the same GPU), same weights / batch / RNG seed.  north_star gate: final slices <= 1e-2 max-abs."""
    _strict_fp32()
    steps = 5
    hosts = {}
    for name, targets in (("b200", R.B200_TARGETS), ("ref", R.REF_TARGETS)):
        m = R.ds_diff_model(targets["ds"], R.DS_SMALL, sampler=sampler, steps=steps, b200_samplers=name == "b200")
        _load_synth(m.model.diffusion_model)
        if sampler == "ddpm":
            m.register_schedule(timesteps=8, linear_start=1e-4, linear_end=2e-2)      # an 8-step ancestral chain
        hosts[name] = m.cuda().eval()
    outs = {}
    for name, m in hosts.items():
        R.use_samplers(name == "b200")
        torch.manual_seed(11)
        with _quiet(), torch.no_grad():
            ids, out, slices = m.predict_step(_cuda_batch(), 0)
        assert ids == ["id0", "id0"] and slices == [0, 1]
        outs[name] = out.float().cpu()
    assert outs["b200"].shape == outs["ref"].shape == (2, 1, 32, 32)
    err = float((outs["b200"] - outs["ref"]).abs().max())
    print(f"DSDiffModel[{sampler}] B200 path vs reference itself: final max-abs {err:.2e}")
    assert err <= tol


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("sampler", ["ddim", "ddpm", "dpm"])
def test_reference_gaussian_trainer_on_b200_path_vs_reference_itself(sampler):
    """TryTrainerDiffusion.predict_step (learned-sigma 2-channel DS-Diff, respaced guided-diffusion process, rescaled float
    timesteps): B200 path vs the reference's own modules in strict fp32."""
    _strict_fp32()
    outs = {}
    for name, targets in (("b200", R.B200_TARGETS), ("ref", R.REF_TARGETS)):
        g = R.gaussian_trainer(targets["ds"], R.DS_FILM2, sampler=sampler, steps=6, b200_samplers=name == "b200")
        _load_synth(g.model.diffusion_model)
        g = g.cuda().eval()
        torch.manual_seed(12)
        with _quiet(), torch.no_grad():
            _, out, _ = g.predict_step(_cuda_batch(), 0)
        outs[name] = out.float().cpu()
    err = float((outs["b200"] - outs["ref"]).abs().max())
    print(f"TryTrainerDiffusion[{sampler}] B200 path vs reference itself: final max-abs {err:.2e}")
    assert err <= 1e-2
