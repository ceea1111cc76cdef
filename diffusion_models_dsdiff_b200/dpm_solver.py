"""DPM-Solver++ (2M) sampling on libb200diff — drop-in for ``ldm.models.diffusion.dpm_solver_new.DPMSolverSampler``
(dpm_solver_new/sampler.py:19-103) and the engine behind ``GaussianDiffusion.dpm_solver_sample_loop``
(training_project/utils/gaussian_diffusion.py:468-523).

The reference evaluates NoiseScheduleVP / DPM_Solver scalar expressions (dpm_solver_pytorch.py) with one torch launch
per operator per step and three to six broadcast tensor ops for every update.  Here the per-step scalars are evaluated
ONCE on the host with the reference's own fp32 expressions into a coefficient table [S, 16]; each step is then the
network + ONE fused elementwise kernel (``b200_sampler_update`` mode DPMPP_2M: data prediction, first / second order
multistep update, history rotation) selected by a device step counter, so the whole step replays as one CUDA graph.
Dynamic thresholding (the guided-diffusion caller) adds an exact per-sample radix-select quantile kernel.

Only what the reference's callers use is built: schedule 'discrete', algorithm 'dpmsolver++', method 'multistep',
order <= 2, solver_type 'dpmsolver', skip types 'time_uniform' / 'logSNR' / 'time_quadratic'.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib as L
from .samplers import (NCOEF, StepEngine, _engine_model_fn, _graph_cache_args, _PARAM, DiffusionWrapper, gd_conditioning,  # noqa: F401
                       host_cond_list, host_context, host_denoiser, host_graph_cache)

Tensor = torch.Tensor
MODEL_TYPES = {"eps": "noise", "v": "v"}        # dpm_solver_new/sampler.py:14-17


def interpolate_fn(x: Tensor, xp: Tensor, yp: Tensor) -> Tensor:
    """Piecewise-linear y(x) through the keypoints (xp increasing), extended linearly beyond both ends: the function
    dpm_solver_pytorch.py:1255-1294 computes (C = 1).  x [N], xp / yp [K] -> [N]."""
    k = xp.numel()
    j = (torch.searchsorted(xp.contiguous(), x.contiguous()) - 1).clamp(0, k - 2)
    return yp[j] + (x - xp[j]) * (yp[j + 1] - yp[j]) / (xp[j + 1] - xp[j])


class NoiseScheduleVP:
    """NoiseScheduleVP('discrete', betas= | alphas_cumprod=) (dpm_solver_pytorch.py:8-170), host (CPU fp32) tensors."""

    def __init__(self, schedule="discrete", betas=None, alphas_cumprod=None, dtype=torch.float32):
        if schedule != "discrete":
            raise ValueError(f"Unsupported noise schedule {schedule}. The B200 path builds the 'discrete' schedule only")
        if betas is not None:
            log_alphas = 0.5 * torch.log(1 - betas.detach().float().cpu()).cumsum(dim=0)
        else:
            assert alphas_cumprod is not None
            log_alphas = 0.5 * torch.log(alphas_cumprod.detach().float().cpu())
        self.schedule = schedule
        self.T = 1.
        self.log_alpha_array = self.numerical_clip_alpha(log_alphas).to(dtype)
        self.total_N = int(self.log_alpha_array.numel())
        self.t_array = torch.linspace(0., 1., self.total_N + 1)[1:].to(dtype)

    @staticmethod
    def numerical_clip_alpha(log_alphas, clipped_lambda=-5.1):
        """Drop the schedule tail whose half-logSNR is below ``clipped_lambda`` (:116-127)."""
        log_sigmas = 0.5 * torch.log(1. - torch.exp(2. * log_alphas))
        lambs = log_alphas - log_sigmas
        idx = int(torch.searchsorted(torch.flip(lambs, [0]), torch.tensor(clipped_lambda)))
        return log_alphas[:-idx] if idx > 0 else log_alphas

    def marginal_log_mean_coeff(self, t):
        return interpolate_fn(t.reshape(-1), self.t_array, self.log_alpha_array)

    def marginal_alpha(self, t):
        return torch.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t):
        return torch.sqrt(1. - torch.exp(2. * self.marginal_log_mean_coeff(t)))

    def marginal_lambda(self, t):
        log_mean_coeff = self.marginal_log_mean_coeff(t)
        return log_mean_coeff - 0.5 * torch.log(1. - torch.exp(2. * log_mean_coeff))

    def inverse_lambda(self, lamb):
        log_alpha = -0.5 * torch.logaddexp(torch.zeros((1,)), -2. * lamb.reshape(-1))
        return interpolate_fn(log_alpha, torch.flip(self.log_alpha_array, [0]), torch.flip(self.t_array, [0]))


def get_time_steps(ns: NoiseScheduleVP, skip_type: str, t_T: float, t_0: float, N: int) -> Tensor:
    """DPM_Solver.get_time_steps (:455-482)."""
    if skip_type == "logSNR":
        lambda_T, lambda_0 = ns.marginal_lambda(torch.tensor(t_T)), ns.marginal_lambda(torch.tensor(t_0))
        return ns.inverse_lambda(torch.linspace(lambda_T.item(), lambda_0.item(), N + 1))
    if skip_type == "time_uniform":
        return torch.linspace(t_T, t_0, N + 1)
    if skip_type == "time_quadratic":
        return torch.linspace(t_T ** 0.5, t_0 ** 0.5, N + 1).pow(2)
    raise ValueError(f"Unsupported skip_type {skip_type}, need to be 'logSNR' or 'time_uniform' or 'time_quadratic'")


def multistep_tables(ns: NoiseScheduleVP, steps: int, skip_type: str = "time_uniform", order: int = 2,
                     lower_order_final: bool = True, t_start=None, t_end=None):
    """Per-update scalars of DPM_Solver.sample(method='multistep') (:1160-1215) for algorithm 'dpmsolver++':
    returns (coef [steps, NCOEF] fp32, model timesteps [steps] fp32, orders).

    coef row k (update from timesteps[k] to timesteps[k+1]):
      0 alpha_s, 1 sigma_s           data_prediction_fn at timesteps[k]            (:435-444)
      2 sigma_t / sigma_s, 3 alpha_t * expm1(-h)                                    (:566-575, :826-833)
      4 1 / r0, 5 0.5 * alpha_t * expm1(-h)   second-order term                     (:822-833)
      6 order of the update (1 for the first step and, when steps < 10, the last)   (:1184-1203)"""
    if order not in (1, 2):
        raise NotImplementedError("B200 path: multistep DPM-Solver++ of order 1 or 2")
    assert steps >= order
    t_0 = 1. / ns.total_N if t_end is None else t_end
    t_T = ns.T if t_start is None else t_start
    assert t_0 > 0 and t_T > 0, "Time range needs to be greater than 0."
    ts = get_time_steps(ns, skip_type, t_T, t_0, steps)
    assert ts.shape[0] - 1 == steps
    coef = torch.zeros((steps, NCOEF), dtype=torch.float32)
    t_input, orders = [], []
    for k in range(steps):
        step = k + 1
        if step < order:
            o = step
        elif lower_order_final and steps < 10:
            o = min(order, steps + 1 - step)
        else:
            o = order
        s, t = ts[k].reshape(1), ts[k + 1].reshape(1)
        lambda_s, lambda_t = ns.marginal_lambda(s), ns.marginal_lambda(t)
        h = lambda_t - lambda_s
        sigma_s, sigma_t = ns.marginal_std(s), ns.marginal_std(t)
        alpha_t = torch.exp(ns.marginal_log_mean_coeff(t))
        phi_1 = torch.expm1(-h)
        coef[k, 0], coef[k, 1] = ns.marginal_alpha(s), sigma_s
        coef[k, 2], coef[k, 3] = sigma_t / sigma_s, alpha_t * phi_1
        if o == 2:
            r0 = (lambda_s - ns.marginal_lambda(ts[k - 1].reshape(1))) / h
            coef[k, 4] = 1. / r0
        coef[k, 5] = 0.5 * (alpha_t * phi_1)
        coef[k, 6] = float(o)
        t_input.append((s - 1. / ns.total_N) * 1000.)            # get_model_input_time (:273-282)
        orders.append(o)
    return coef, torch.cat(t_input).to(torch.float32), orders


def run_multistep(net, ns: NoiseScheduleVP, x_T: Tensor, cond: Optional[List[Tensor]], steps: int, model_type: str,
                  skip_type: str, lower_order_final: bool, thresholding, cache: dict, context=None, use_graph=True,
                  record=None, cfg: Optional[dict] = None):
    coef, t_input, _ = multistep_tables(ns, steps, skip_type, 2, lower_order_final)
    dev = x_T.device
    if model_type not in ("noise", "v"):
        raise NotImplementedError(f"B200 path: DPM-Solver++ model_type {model_type!r}")
    eng = StepEngine(_engine_model_fn(net), L.SAMPLER_DPMPP_2M, _PARAM["eps" if model_type == "noise" else "v"],
                     coef.to(dev), t_input.to(dev), False, context=context, thresholding=thresholding, **(cfg or {}),
                     **_graph_cache_args(cache, net))
    out = eng.run(x_T, cond, steps, use_graph=use_graph, want_x0=record is not None, record=record)
    return out, eng


class DPMSolverSampler(object):
    """dpm_solver_new/sampler.py:19-103: ``sample(S, batch_size, shape, conditioning, ..., x_T=None)`` -> (x, None)."""

    def __init__(self, model, device=None, **kwargs):
        super().__init__()
        self.model = model
        self.device = device if device is not None else model.betas.device
        to_torch = lambda x: x.clone().detach().to(torch.float32)
        self.alphas_cumprod = to_torch(model.alphas_cumprod)
        self.betas = to_torch(model.betas)
        self.use_graph = getattr(model, "use_graph", True)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, **kwargs):
        if conditioning is not None:
            ctmp = conditioning
            if isinstance(ctmp, dict):
                ctmp = ctmp[list(ctmp.keys())[0]]
            while isinstance(ctmp, list):
                ctmp = ctmp[0]
            if isinstance(ctmp, torch.Tensor) and ctmp.shape[0] != batch_size:
                print(f"Warning: Got {ctmp.shape[0]} conditionings but batch-size is {batch_size}")
        C, H, W = shape
        size = (batch_size, C, H, W)
        print(f'Data shape for DPM-Solver sampling is {size}, sampling steps {S}')
        cfg = None
        if unconditional_conditioning is not None and unconditional_guidance_scale != 1.:     # model_wrapper :323-331
            cfg = dict(cfg_scale=unconditional_guidance_scale,
                       uncond=host_cond_list(self.model, unconditional_conditioning),
                       uncond_context=host_context(self.model, unconditional_conditioning))
        device = self.model.betas.device
        img = torch.randn(size, device=device) if x_T is None else x_T.to(device)
        ns = NoiseScheduleVP('discrete', betas=self.betas)
        m = self.model
        x, eng = run_multistep(host_denoiser(m), ns, img, host_cond_list(m, conditioning), S,
                               MODEL_TYPES[m.parameterization], "time_uniform", True, None,
                               host_graph_cache(m), context=host_context(m, conditioning),
                               use_graph=self.use_graph, record=kwargs.get("record"), cfg=cfg)
        self.last_engine = eng
        m.__dict__["last_engine"] = eng
        return x.to(device), None


def gaussian_dpm_solver_sample_loop(diffusion, model, shape, clip_denoised=None, model_kwargs=None, noise=None,
                                    record=None):
    """GaussianDiffusion.dpm_solver_sample_loop (gaussian_diffusion.py:468-523): the RESPACED betas are the discrete
    schedule, model_type 'v', logSNR steps, dynamic thresholding (ratio 0.995, max 1), lower_order_final=False; the model
    is called with the solver's own timestep (t - 1/N) * 1000, not through the respacing wrapper."""
    device = next(model.parameters()).device
    x = torch.randn(*shape, device=device) if noise is None else noise.to(device)
    print(x.shape)
    ns = NoiseScheduleVP(schedule='discrete', betas=torch.from_numpy(diffusion.betas).float())
    cond, context = gd_conditioning(model, model_kwargs)
    net = host_denoiser(model)
    out, eng = run_multistep(net, ns, x, cond, diffusion.num_timesteps, "v", "logSNR", False, (0.995, 1.0),
                             diffusion.__dict__.setdefault("_graphs", {}), use_graph=getattr(diffusion, "use_graph", True),
                             record=record, context=context)
    diffusion.last_engine = eng
    return out
