"""Drop-in diffusion processes / samplers on the fused B200 update kernel.

Mirrors the reference's sampler seam — same class / method names, argument meaning
and error behaviour:

* ``DDPM`` + ``DiffusionWrapper``  — ldm/models/diffusion/ddpm.py:46-192, 284-356, 1319-1365 with the
  conditional / v-prediction sampling overrides the reference's trainers install
  (trainers/trainer_ds_diff.py:698-755);
* ``DDIMSampler``                  — ldm/models/diffusion/ddim.py:11-261;
* ``GaussianDiffusion`` / ``SpacedDiffusion`` / ``create_gaussian_diffusion`` / ``space_timesteps`` —
  training_project/utils/gaussian_diffusion.py:113-789, respace.py:7-128, script_util.py:129-169.

Schedule tables are built on the HOST with the reference's own float64 -> float32 expressions (so
every per-step scalar is bit-identical to what ``torch.full((b,1,1,1), table[index])`` would hold),
packed into one ``[steps, 16]`` fp32 coefficient table on the device, and consumed by ONE fused
kernel per step (``ops.sampler_update`` -> ``b200_sampler_update``).  The step index lives in a
device int, so a whole denoising step (network + update) is captured once as a CUDA graph and
replayed; no per-step host->device scalar traffic or ``.item()`` syncs remain.

Coefficient row layout (``B200_SAMPLER_NCOEF`` = 16 floats, unused = 0):
  LDM_DDIM : 0 sqrt_ac[t]  1 sqrt_1m_ac[t]  2 sqrt(a_t)  3 sqrt(1-a_t)  4 sqrt(a_prev)
             5 sqrt(1-a_prev-sigma^2)  6 sigma  7 temperature
  LDM_DDPM : 0 ca  1 cb  (x0 = ca*x - cb*out)  2 post_mean_coef1  3 post_mean_coef2
             4 [t!=0]*exp(0.5*post_log_var_clipped)
  GD_*     : 0 ca  1 cb  2 coef1  3 coef2  4 min_log (or the fixed log-variance)  5 max_log
             6 sqrt_recip_ac  7 sqrt_recipm1_ac  8 sqrt(ac_prev)  9 sqrt(1-ac_prev-sigma^2)
             10 [t!=0]*sigma  11 [t!=0]
"""
from __future__ import annotations

import importlib
import math
import warnings
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops

Tensor = torch.Tensor
NCOEF = L.SAMPLER_NCOEF
_PARAM = {"eps": L.PARAM_EPS, "x0": L.PARAM_X0, "v": L.PARAM_V}


# ---------------------------------------------------------------------------------------------
# config plumbing (ldm/util.py:82-97)
# ---------------------------------------------------------------------------------------------
# reference dotted paths -> the B200 drop-in classes
TARGET_ALIASES = {
    "ldm.modules.diffusionmodules.openaimodel.UNetModel": "diffusion_models_dsdiff_b200.unet.UNetModel",
    "UNet_DS_Diff.model.DSUnetModel": "diffusion_models_dsdiff_b200.dsunet.DSUnetModel",
    "ldm.models.autoencoder.AutoencoderKL": "diffusion_models_dsdiff_b200.autoencoder.AutoencoderKL",
}


def get_obj_from_str(string: str):
    string = TARGET_ALIASES.get(string, string)
    module, cls = string.rsplit(".", 1)
    return getattr(importlib.import_module(module, package=None), cls)


def instantiate_from_config(config):
    if "target" not in config:
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**config.get("params", dict()))


# ---------------------------------------------------------------------------------------------
# schedules (ldm/modules/diffusionmodules/util.py:21-81)
# ---------------------------------------------------------------------------------------------
def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    if schedule == "linear":     # linear in sqrt(beta)
        betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    elif schedule == "cosine":
        timesteps = torch.arange(n_timestep + 1, dtype=torch.float64) / n_timestep + cosine_s
        alphas = torch.cos(timesteps / (1 + cosine_s) * np.pi / 2).pow(2)
        alphas = alphas / alphas[0]
        betas = np.clip(1 - alphas[1:] / alphas[:-1], a_min=0, a_max=0.999)
    elif schedule == "sqrt_linear":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64)
    elif schedule == "sqrt":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64) ** 0.5
    else:
        raise ValueError(f"schedule '{schedule}' unknown.")
    return betas.numpy() if isinstance(betas, torch.Tensor) else np.asarray(betas)


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    if ddim_discr_method == "uniform":
        c = num_ddpm_timesteps // num_ddim_timesteps
        ddim_timesteps = np.asarray(list(range(0, num_ddpm_timesteps, c)))
    elif ddim_discr_method == "quad":
        ddim_timesteps = ((np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps)) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    steps_out = ddim_timesteps + 1
    if verbose:
        print(f"Selected timesteps for ddim sampler: {steps_out}")
    return steps_out


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    """alphacums: CPU fp32 tensor.  Keeps the reference's fp32-tensor / float64-ndarray type mix
    (util.py:70-81) so sigmas round exactly like the reference's."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    if verbose:
        print(f"Selected alphas for ddim sampler: a_t: {alphas}; a_(t-1): {alphas_prev}")
        print(f"For the chosen value of eta, which is {eta}, "
              f"this results in the following sigma_t schedule for ddim sampler {sigmas}")
    return sigmas, alphas, alphas_prev


def extract_into_tensor(a, t, x_shape):
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


def noise_like(shape, device, repeat=False):
    if repeat:
        return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
    return torch.randn(shape, device=device)


def _f32(v) -> Tensor:
    """fp32 CPU tensor holding what ``torch.full(..., v)`` would hold (float64 -> float32 rounding)."""
    if isinstance(v, torch.Tensor):
        return v.detach().to("cpu", torch.float64).to(torch.float32)
    return torch.from_numpy(np.asarray(v, dtype=np.float64)).to(torch.float32)


def _coef_table(cols: Dict[int, Tensor], steps: int) -> Tensor:
    tab = torch.zeros((steps, NCOEF), dtype=torch.float32)
    for j, col in cols.items():
        tab[:, j] = col
    return tab


# ---------------------------------------------------------------------------------------------
# the step engine: one denoising step = network + fused update, optionally a replayed CUDA graph
# ---------------------------------------------------------------------------------------------
class StepEngine:
    """Runs ``steps`` iterations of  x <- update(x, model(cat[x, cond], t_i), z_i)  on the device.

    ``model_fn(xc, t_table, step_ptr)`` must read its timestep as ``t_table[*step_ptr]``.
    ``use_graph``: capture iteration 0 as a CUDA graph (after one eager warm-up on a scratch copy)
    and replay it; noise is drawn by ``torch.randn`` INSIDE the graph so the consumption order of
    torch's CUDA generator is the reference's (x_T first, then one draw per step, even when
    sigma == 0).  With ``noises`` given (parity tests) the loop runs eagerly with those tensors.
    """

    def __init__(self, model_fn: Callable, mode: int, param: int, coef: Tensor, t_table: Tensor,
                 clip_denoised: bool, learned_range: bool = False, cache: Optional[dict] = None, cache_key=None,
                 keepalive=None, prepare: Optional[Callable] = None, prepare_cond: Optional[Callable] = None,
                 context: Optional[Tensor] = None,
                 thresholding: Optional[Tuple[float, float]] = None, cfg_scale: float = 1.0,
                 uncond: Optional[Sequence[Tensor]] = None, uncond_context: Optional[Tensor] = None):
        self.model_fn, self.mode, self.param = model_fn, mode, param
        # classifier-free guidance (ddim.py:194-219): the network runs on [x; x] with [unconditional; conditional]
        # conditioning and the two halves are combined before the update
        self.cfg_scale = float(cfg_scale)
        self.uncond = None if uncond is None else [u.float().contiguous() for u in uncond]
        self.uncond_context = uncond_context
        self.cfg = self.cfg_scale != 1.0 and (uncond is not None or uncond_context is not None)
        self.thresholding = thresholding    # DPM-Solver++ dynamic thresholding (ratio, max_val) or None
        self._dpm: Optional[dict] = None    # DPM-Solver++ static buffers: previous x0, raw x0, per-sample thresholds
        self.coef, self.t_table = coef, t_table
        self.clip, self.learned_range = clip_denoised, learned_range
        # captured graphs are reused across calls (the reference builds a new DDIMSampler per batch);
        # cache_key must change whenever the network's packed weights do, keepalive pins them
        self.cache, self.cache_key, self.keepalive = cache, cache_key, keepalive
        self.prepare = prepare
        self.prepare_cond = prepare_cond  # loop-invariant conditioning prefix of the network, once per run
        self.model_kwargs: dict = {}
        self.context = context          # cross-attention conditioning [B, Lk, D], constant over the steps
        self.stats: dict = {}

    def _one(self, x: Tensor, cond: Optional[Sequence[Tensor]], step_ptr: Tensor, noise: Optional[Tensor],
             pred_x0: Optional[Tensor]):
        if self.cfg:
            n = x.shape[0]
            x2 = torch.cat([x, x], dim=0)
            if cond:
                unc = self.uncond if self.uncond is not None else cond
                x2 = torch.cat([x2] + [torch.cat([u, c], dim=0) for u, c in zip(unc, cond)], dim=1)
            kw = {}
            if self.context is not None:
                uctx = self.uncond_context if self.uncond_context is not None else self.context
                kw["context"] = torch.cat([uctx, self.context], dim=0)
            out = self.model_fn(x2, self.t_table, step_ptr, **kw, **self.model_kwargs)
            if isinstance(out, tuple):
                out = out[0]
            out = ops.cfg_combine(out[:n].contiguous(), out[n:].contiguous(), self.cfg_scale)
        else:
            xc = x if not cond else torch.cat([x] + list(cond), dim=1)
            if self.context is not None:
                out = self.model_fn(xc, self.t_table, step_ptr, context=self.context, **self.model_kwargs)
            else:
                out = self.model_fn(xc, self.t_table, step_ptr, **self.model_kwargs)
            if isinstance(out, tuple):
                out = out[0]
        if self.mode == L.SAMPLER_DPMPP_2M:      # deterministic ODE solver: no noise is drawn (dpm_solver_pytorch.py)
            b = self._dpm
            if out.shape[1] == 2 * x.shape[1]:   # learn-sigma networks: mean half only (gaussian_diffusion.py:484-487)
                out = out[:, :x.shape[1]]
            out = out.contiguous()
            if self.thresholding is None:
                ops.dpm_solver_update(self.param, x, out, self.coef, b["hist"], step_ptr=step_ptr, pred_x0=pred_x0, x_next=x)
            else:
                ops.dpm_solver_update(self.param, x, out, self.coef, None, step_ptr=step_ptr, phase=1, pred_x0=b["x0"])
                ops.abs_quantile(b["x0"], self.thresholding[0], self.thresholding[1], out=b["thr"])
                ops.dpm_solver_update(self.param, x, None, self.coef, b["hist"], step_ptr=step_ptr, phase=2,
                                      pred_x0=b["x0"], thresh=b["thr"], x_next=x)
                if pred_x0 is not None:
                    pred_x0.copy_(b["x0"])
            ops.advance_step(step_ptr, 1)
            return out
        z = noise if noise is not None else torch.randn_like(x)
        ops.sampler_update(self.mode, self.param, x, out.contiguous(), z, self.coef, step_ptr=step_ptr,
                           clip_denoised=self.clip, learned_range=self.learned_range, x_prev=x,
                           pred_x0=pred_x0, want_x0=pred_x0 is not None)
        ops.advance_step(step_ptr, 1)
        return out

    def run(self, x_T: Tensor, cond, steps: int, noises: Optional[Sequence[Tensor]] = None, use_graph: bool = True,
            on_step: Optional[Callable] = None, want_x0: bool = True, record: Optional[list] = None) -> Tensor:
        dev = x_T.device
        x = x_T.clone().float().contiguous()
        pred_x0 = torch.empty_like(x) if want_x0 else None
        step_ptr = torch.zeros(1, dtype=torch.int32, device=dev)
        if self.prepare is not None:       # timestep MLP + emb projections for ALL steps, once per loop
            self.prepare(self.t_table)
        self.model_kwargs = {}
        if self.prepare_cond is not None and cond:
            cc = list(cond)
            if self.cfg:                   # the network runs on [uncond; cond] (see _one)
                unc = self.uncond if self.uncond is not None else cc
                cc = [torch.cat([u.to(dev), c.to(dev)], dim=0) for u, c in zip(unc, cc)]
            if self.prepare_cond(cc):
                self.model_kwargs = {"cond_cached": True}
        if self.mode == L.SAMPLER_DPMPP_2M:
            self._dpm = {"hist": torch.zeros_like(x), "x0": torch.empty_like(x),
                         "thr": torch.ones(x.shape[0], dtype=torch.float32, device=dev)}
        if noises is not None or record is not None or not use_graph or steps < 3:
            for i in range(steps):
                out = self._one(x, cond, step_ptr, None if noises is None else noises[i].to(dev).float().contiguous(),
                                pred_x0)
                if record is not None:
                    record.append({"model_out": out.clone(), "x_prev": x.clone(),
                                   "pred_x0": None if pred_x0 is None else pred_x0.clone()})
                if on_step:
                    on_step(i, x, pred_x0)
            ops.check_range(dev, "the sampling loop")
            return x
        # ---- graph path ----
        key = None
        if self.cache is not None:
            key = (self.cache_key, self.mode, self.param, self.clip, self.learned_range, self.thresholding, self.cfg,
                   bool(self.model_kwargs),
                   self.cfg_scale, tuple(x.shape),
                   tuple(tuple(c.shape) for c in (cond or [])), steps, want_x0,
                   None if self.context is None else tuple(self.context.shape))
        entry = self.cache.get(key) if key is not None else None
        if entry is None:
            entry = self._capture(x, cond, pred_x0, step_ptr, key)
        else:
            entry["coef"].copy_(self.coef)
            entry["t_table"].copy_(self.t_table)
            entry["x"].copy_(x)
            for dst, src in zip(entry["cond"], cond or []):
                dst.copy_(src)
            if self.context is not None:
                entry["context"].copy_(self.context)
            if self.uncond is not None:
                for dst, src in zip(entry["uncond"], self.uncond):
                    dst.copy_(src)
            if self.uncond_context is not None:
                entry["uncond_context"].copy_(self.uncond_context)
            entry["step_ptr"].zero_()
            self._dpm = entry.get("dpm")
            self.stats = {"graph_launches_per_step": entry["launches"], "graph_reused": True}
        xs, x0s = entry["x"], entry["pred_x0"]
        for i in range(steps):
            entry["graph"].replay()
            if on_step:
                on_step(i, xs, x0s)
        self.stats["replays"] = steps
        ops.check_range(dev, "the sampling loop")
        return xs.clone() if key is not None else xs

    def _capture(self, x, cond, pred_x0, step_ptr, key):
        """Warm-up one step on scratch state (restoring the RNG state so the caller-visible random stream is
        untouched), then capture one step on static buffers."""
        dev = x.device
        coef, t_table = self.coef.clone(), self.t_table.clone()
        self.coef, self.t_table = coef, t_table           # the captured kernels read these static tables
        cond = [c.clone().contiguous() for c in (cond or [])]
        if self.context is not None:
            self.context = self.context.clone().contiguous()      # static buffer read by the captured kernels
        if self.uncond is not None:
            self.uncond = [u.clone().contiguous() for u in self.uncond]
        if self.uncond_context is not None:
            self.uncond_context = self.uncond_context.clone().contiguous()
        gen_state = torch.cuda.get_rng_state(dev)
        scratch = x.clone()
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            self._one(scratch, cond, step_ptr, None, pred_x0)
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        torch.cuda.set_rng_state(gen_state, dev)
        step_ptr.zero_()
        graph = torch.cuda.CUDAGraph()
        n0 = L.launch_count()
        with torch.cuda.graph(graph, stream=s):
            self._one(x, cond, step_ptr, None, pred_x0)
        launches = L.launch_count() - n0
        # capture does not execute: x / step_ptr / RNG are untouched
        entry = {"graph": graph, "x": x, "cond": cond, "pred_x0": pred_x0, "step_ptr": step_ptr, "coef": coef,
                 "t_table": t_table, "launches": launches, "keepalive": self.keepalive, "context": self.context,
                 "dpm": self._dpm, "uncond": self.uncond, "uncond_context": self.uncond_context}
        self.stats = {"graph_launches_per_step": launches, "graph_reused": False}
        if key is not None:
            if len(self.cache) >= 4:            # bounded: each entry pins a private activation pool
                self.cache.pop(next(iter(self.cache)))
            self.cache[key] = entry
        return entry


def _engine_model_fn(net):
    """model_fn(xc, t_table, step_ptr) for a B200 engine network; the DS-Diff feature dict (discarded by
    every sampler, trainer_ds_diff.py:752-755 / gaussian_diffusion.py:278-279) is not materialised."""
    if not getattr(net, "_b200_engine", False):
        raise L.B200Error(
            f"{type(net).__module__}.{type(net).__name__} is not a B200 engine network: point the yaml "
            "`unet_config.target` at diffusion_models_dsdiff_b200.unet.UNetModel / .dsunet.DSUnetModel / .dit.DiT "
            "(INTEGRATION.md §1).  The samplers have no eager / PyTorch fallback.")
    extra = {"return_features": False} if getattr(net, "_returns_features", False) else {}
    return lambda xc, t_table, step_ptr, **kw: net(xc, t_table, step_ptr=step_ptr, **extra, **kw)


def _graph_cache_args(cache: dict, net) -> dict:
    pw = net.packed() if hasattr(net, "packed") else None
    return {"cache": cache, "cache_key": (id(net), id(pw)), "keepalive": pw,
            "prepare": getattr(net, "prepare_steps", None), "prepare_cond": getattr(net, "prepare_conditioning", None)}


# ---------------------------------------------------------------------------------------------
# the sampler seam: everything the samplers need from the HOST object is read through these helpers,
# from PUBLIC attributes only, so the host may be this package's DDPM / DiffusionWrapper or the REFERENCE's own
# (``trainers.trainer_ds_diff.DSDiffModel(DDPM)``: ``DDIMSampler(self)`` at trainer_ds_diff.py:672;
# ``TryTrainerDiffusion.model = DiffusionWrapper(...)``: ``sample_fn(self.model, ...)`` at
# trainer_use_gaussian_diff.py:86,615) holding a B200 network instantiated from the swapped yaml ``target``.
# ---------------------------------------------------------------------------------------------
def host_denoiser(model_or_wrapper):
    """DDPM-like host (``.model`` is a DiffusionWrapper), a DiffusionWrapper (``.diffusion_model``) or a bare network
    -> the network (ddpm.py:1320-1326)."""
    m = model_or_wrapper
    inner = getattr(m, "model", None)
    if inner is not None and hasattr(inner, "diffusion_model"):
        m = inner
    return getattr(m, "diffusion_model", m)


def host_conditioning_key(model_or_wrapper):
    m = model_or_wrapper
    inner = getattr(m, "model", None)
    if inner is not None and hasattr(inner, "diffusion_model"):
        m = inner
    return getattr(m, "conditioning_key", None)


def host_cond_list(model, cond):
    """The tensors DiffusionWrapper concatenates to x on the channel axis (ddpm.py:1331-1333), from whatever the caller
    passes as conditioning: a tensor, a list (apply_model wraps it under the wrapper's key, trainer_ds_diff.py:740-748)
    or the hybrid dict."""
    key = host_conditioning_key(model)
    if cond is None or key is None or key in ("crossattn", "adm", "crossattn-adm"):
        return None
    if isinstance(cond, dict):
        cond = cond.get("c_concat")
        if cond is None:
            return None
    return list(cond) if isinstance(cond, (list, tuple)) else [cond]


def host_context(model, cond) -> Optional[Tensor]:
    """The cross-attention context of the 'crossattn' / 'hybrid' keys: cat(c_crossattn, 1) (ddpm.py:1334-1347)."""
    key = host_conditioning_key(model)
    if cond is None or key not in ("crossattn", "hybrid"):
        return None
    if isinstance(cond, dict):
        cond = cond.get("c_crossattn")
        if cond is None:
            return None
    elif key == "hybrid":
        return None
    cond = list(cond) if isinstance(cond, (list, tuple)) else [cond]
    return torch.cat(cond, 1).float().contiguous()


def host_graph_cache(host) -> dict:
    """Captured step graphs live on the host object (the reference builds a fresh DDIMSampler per batch,
    trainer_ds_diff.py:672); ``__dict__`` so a reference nn.Module host does not try to register it."""
    return host.__dict__.setdefault("_b200_graphs", {})


def host_engine_args(host, net=None) -> dict:
    net = host_denoiser(host) if net is None else net
    return _graph_cache_args(host_graph_cache(host), net)


# ---------------------------------------------------------------------------------------------
# ldm: DiffusionWrapper + DDPM
# ---------------------------------------------------------------------------------------------
class DiffusionWrapper(nn.Module):
    """ldm/models/diffusion/ddpm.py:1319-1365: every conditioning key of the reference (None, concat, crossattn, hybrid,
    adm, hybrid-adm, crossattn-adm)."""

    def __init__(self, diff_model_config, conditioning_key):
        super().__init__()
        if isinstance(diff_model_config, nn.Module):
            self.diffusion_model = diff_model_config
        else:
            self.sequential_cross_attn = diff_model_config.pop("sequential_crossattn", False)
            self.diffusion_model = instantiate_from_config(diff_model_config)
        self.conditioning_key = conditioning_key
        assert self.conditioning_key in [None, "concat", "crossattn", "hybrid", "adm", "hybrid-adm", "crossattn-adm"]

    def forward(self, x, t, c_concat: list = None, c_crossattn: list = None, c_adm=None, **kw):
        if self.conditioning_key is None:
            return self.diffusion_model(x, t, **kw)
        if self.conditioning_key == "concat":
            xc = torch.cat([x] + list(c_concat), dim=1)
            return self.diffusion_model(xc, t, **kw)
        if self.conditioning_key == "crossattn":
            return self.diffusion_model(x, t, context=torch.cat(list(c_crossattn), 1), **kw)
        if self.conditioning_key == "hybrid":
            xc = torch.cat([x] + list(c_concat), dim=1)
            return self.diffusion_model(xc, t, context=torch.cat(list(c_crossattn), 1), **kw)
        if self.conditioning_key == "hybrid-adm":          # ddpm.py:1350-1354
            assert c_adm is not None
            xc = torch.cat([x] + list(c_concat), dim=1)
            return self.diffusion_model(xc, t, context=torch.cat(list(c_crossattn), 1), y=c_adm, **kw)
        if self.conditioning_key == "crossattn-adm":       # ddpm.py:1355-1358
            assert c_adm is not None
            cc = torch.cat(list(c_crossattn), 1) if c_crossattn[0] is not None else None
            return self.diffusion_model(x, t, context=cc, y=c_adm, **kw)
        if self.conditioning_key == "adm":                 # ddpm.py:1359-1361
            return self.diffusion_model(x, t, y=c_crossattn[0], **kw)
        raise NotImplementedError()


class DDPM(nn.Module):
    """Sampling half of ldm DDPM with the trainers' conditional / v-prediction overrides."""

    def __init__(self, unet_config, timesteps=1000, beta_schedule="linear", loss_type="l2", ckpt_path=None,
                 ignore_keys=(), load_only_unet=False, monitor="val/loss", use_ema=False, first_stage_key="image",
                 image_size=256, channels=3, log_every_t=100, clip_denoised=True, linear_start=1e-4,
                 linear_end=2e-2, cosine_s=8e-3, given_betas=None, original_elbo_weight=0., v_posterior=0.,
                 l_simple_weight=1., conditioning_key=None, parameterization="eps", scheduler_config=None,
                 use_positional_encodings=False, learn_logvar=False, logvar_init=0., make_it_fit=False,
                 ucg_training=None, reset_ema=False, reset_num_ema_updates=False, **ignored):
        super().__init__()
        assert parameterization in ["eps", "x0", "v"], 'currently only supporting "eps" and "x0" and "v"'
        self.parameterization = parameterization
        self.clip_denoised = clip_denoised
        self.log_every_t = log_every_t
        self.first_stage_key = first_stage_key
        self.image_size = image_size
        self.channels = channels
        self.model = DiffusionWrapper(unet_config, conditioning_key)
        self.use_ema = False
        self.v_posterior = v_posterior
        self.use_graph = True
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys, only_model=load_only_unet)
        self.register_schedule(given_betas=given_betas, beta_schedule=beta_schedule, timesteps=timesteps,
                               linear_start=linear_start, linear_end=linear_end, cosine_s=cosine_s)

    @property
    def device(self):
        return self.betas.device

    def init_from_ckpt(self, path, ignore_keys=(), only_model=False):
        sd = torch.load(path, map_location="cpu")
        if "state_dict" in list(sd.keys()):
            sd = sd["state_dict"]
        for k in list(sd.keys()):
            for ik in ignore_keys:
                if k.startswith(ik):
                    del sd[k]
        missing, unexpected = (self.load_state_dict(sd, strict=False) if not only_model
                               else self.model.load_state_dict(sd, strict=False))
        print(f"Restored from {path} with {len(missing)} missing and {len(unexpected)} unexpected keys")

    def register_schedule(self, given_betas=None, beta_schedule="linear", timesteps=1000, linear_start=1e-4,
                          linear_end=2e-2, cosine_s=8e-3):
        """float64 numpy tables -> fp32 buffers with the reference's names (ddpm.py:138-192)."""
        betas = given_betas if given_betas is not None else make_beta_schedule(
            beta_schedule, timesteps, linear_start=linear_start, linear_end=linear_end, cosine_s=cosine_s)
        alphas = 1. - betas
        alphas_cumprod = np.cumprod(alphas, axis=0)
        alphas_cumprod_prev = np.append(1., alphas_cumprod[:-1])
        timesteps, = betas.shape
        self.num_timesteps = int(timesteps)
        self.linear_start, self.linear_end = linear_start, linear_end
        to_torch = lambda a: torch.tensor(a, dtype=torch.float32)
        reg = lambda name, a: self.register_buffer(name, to_torch(a))
        reg("betas", betas)
        reg("alphas_cumprod", alphas_cumprod)
        reg("alphas_cumprod_prev", alphas_cumprod_prev)
        reg("sqrt_alphas_cumprod", np.sqrt(alphas_cumprod))
        reg("sqrt_one_minus_alphas_cumprod", np.sqrt(1. - alphas_cumprod))
        reg("log_one_minus_alphas_cumprod", np.log(1. - alphas_cumprod))
        reg("sqrt_recip_alphas_cumprod", np.sqrt(1. / alphas_cumprod))
        reg("sqrt_recipm1_alphas_cumprod", np.sqrt(1. / alphas_cumprod - 1))
        posterior_variance = (1 - self.v_posterior) * betas * (1. - alphas_cumprod_prev) / (
            1. - alphas_cumprod) + self.v_posterior * betas
        reg("posterior_variance", posterior_variance)
        reg("posterior_log_variance_clipped", np.log(np.maximum(posterior_variance, 1e-20)))
        reg("posterior_mean_coef1", betas * np.sqrt(alphas_cumprod_prev) / (1. - alphas_cumprod))
        reg("posterior_mean_coef2", (1. - alphas_cumprod_prev) * np.sqrt(alphas) / (1. - alphas_cumprod))

    # ---- the reference's elementwise helpers (API parity; the sampling loops use the fused kernel) ----
    # ---- forward-process / parameterisation helpers (ddpm.py:284-311, 349-356): one fused kernel each; the per-sample
    # coefficients are row gathers of the schedule buffers (no arithmetic on the host side) ----
    def _comb(self, c1_name, x, c2_name, y, t, subtract):
        t = t.reshape(-1).long()
        c1 = getattr(self, c1_name).to(x.device)[t]
        c2 = getattr(self, c2_name).to(x.device)[t]
        return ops.lincomb2(x.float(), y.float(), c1, c2, subtract)

    def predict_start_from_noise(self, x_t, t, noise):
        return self._comb("sqrt_recip_alphas_cumprod", x_t, "sqrt_recipm1_alphas_cumprod", noise, t, True)

    def predict_start_from_z_and_v(self, x_t, t, v):
        return self._comb("sqrt_alphas_cumprod", x_t, "sqrt_one_minus_alphas_cumprod", v, t, True)

    def predict_eps_from_z_and_v(self, x_t, t, v):
        return self._comb("sqrt_alphas_cumprod", v, "sqrt_one_minus_alphas_cumprod", x_t, t, False)

    def get_v(self, x, noise, t):
        """ddpm.py:358-362: v = sqrt(ac) * noise - sqrt(1 - ac) * x."""
        return self._comb("sqrt_alphas_cumprod", noise, "sqrt_one_minus_alphas_cumprod", x, t, True)

    def q_posterior(self, x_start, x_t, t):
        mean = self._comb("posterior_mean_coef1", x_start, "posterior_mean_coef2", x_t, t, False)
        return (mean, extract_into_tensor(self.posterior_variance, t, x_t.shape),
                extract_into_tensor(self.posterior_log_variance_clipped, t, x_t.shape))

    def q_sample(self, x_start, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        return self._comb("sqrt_alphas_cumprod", x_start, "sqrt_one_minus_alphas_cumprod", noise, t, False)

    # ---- batch -> network input (ddpm.py:423-429; trainers/trainer_ds_diff.py:532-546 with cond_k) ----
    def get_input(self, batch, k, cond_k=None, **kwargs):
        x = batch[k]
        if len(x.shape) == 3:
            x = x[..., None]
        x = x.to(memory_format=torch.contiguous_format).float()
        return x if cond_k is None else (x, batch[cond_k])

    # ---- model call (trainers/trainer_ds_diff.py:740-755) ----
    def apply_model(self, x_noisy, t, cond, return_ids=False, **kw):
        if not isinstance(cond, dict):
            if not isinstance(cond, list):
                cond = [cond]
            key = "c_concat" if self.model.conditioning_key == "concat" else "c_crossattn"
            cond = {key: cond}
        if self.model.conditioning_key is None:
            cond = {}
        x_recon = self.model(x_noisy, t, **cond, **kw)
        if isinstance(x_recon, tuple) and not return_ids:
            return x_recon[0]
        return x_recon

    # ---- ancestral sampling (V2) ----
    def _ddpm_coef(self, steps: Sequence[int]) -> Tensor:
        t = torch.as_tensor(list(steps), dtype=torch.long)
        g = lambda name: getattr(self, name).detach().cpu()[t]
        if self.parameterization == "eps":
            ca, cb = g("sqrt_recip_alphas_cumprod"), g("sqrt_recipm1_alphas_cumprod")
        else:
            ca, cb = g("sqrt_alphas_cumprod"), g("sqrt_one_minus_alphas_cumprod")
        std = (0.5 * g("posterior_log_variance_clipped")).exp() * (t != 0).float()
        return _coef_table({0: ca, 1: cb, 2: g("posterior_mean_coef1"), 3: g("posterior_mean_coef2"), 4: std}, len(t))

    def _cond_list(self, cond):
        return host_cond_list(self, cond)

    def _context(self, cond) -> Optional[Tensor]:
        return host_context(self, cond)

    def _model_fn(self):
        return _engine_model_fn(host_denoiser(self))

    def _graph_cache_args(self):
        return host_engine_args(self)

    @torch.no_grad()
    def p_sample(self, x, t, cond=None, clip_denoised=True, repeat_noise=False, noise=None, **kw):
        """One ancestral step at per-batch timesteps ``t`` (all equal, as every reference loop passes)."""
        tv = t.reshape(-1)
        tt = int(tv[0].item())
        if not bool((tv == tt).all()):
            raise NotImplementedError("B200 path: one timestep per p_sample call (every reference loop passes "
                                      "torch.full((b,), i))")
        coef = self._ddpm_coef([tt]).to(x.device)
        out = self.apply_model(x, t, cond)
        z = noise if noise is not None else noise_like(x.shape, x.device, repeat_noise)
        x_prev, _ = ops.sampler_update(L.SAMPLER_LDM_DDPM, _PARAM[self.parameterization], x.float().contiguous(),
                                       out.contiguous(), z.contiguous(), coef, step=0, clip_denoised=clip_denoised)
        return x_prev

    @torch.no_grad()
    def p_sample_loop(self, shape, cond=None, return_intermediates=False, x_T=None, noises=None, steps=None,
                      record=None, *args, **kwargs):
        device = self.betas.device
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device)
        seq = list(reversed(range(0, self.num_timesteps))) if steps is None else list(steps)
        coef = self._ddpm_coef(seq).to(device)
        t_table = torch.tensor(seq, dtype=torch.float32, device=device)
        eng = StepEngine(self._model_fn(), L.SAMPLER_LDM_DDPM, _PARAM[self.parameterization], coef, t_table,
                         self.clip_denoised, context=self._context(cond), **self._graph_cache_args())
        self.last_engine = eng
        intermediates = [img]

        def on_step(i, x, _x0):
            t = seq[i]
            if t % self.log_every_t == 0 or t == self.num_timesteps - 1:
                intermediates.append(x.clone())

        out = eng.run(img, self._cond_list(cond), len(seq), noises=noises, use_graph=self.use_graph,
                      on_step=on_step if return_intermediates else None, want_x0=record is not None, record=record)
        return (out, intermediates) if return_intermediates else out

    @torch.no_grad()
    def sample(self, batch_size=16, cond=None, shape=None, return_intermediates=False, *args, **kwargs):
        return self.p_sample_loop((batch_size, shape[0], shape[1], shape[2]), cond,
                                  return_intermediates=return_intermediates, **kwargs)


class LatentDiffusion(DDPM):
    """Sampling half of ldm LatentDiffusion (ldm/models/diffusion/ddpm.py:527-1128): a DDPM over first-stage latents.
    ``first_stage_config`` is an AutoencoderKL module or an ``instantiate_from_config`` dict; conditioning arrives
    already encoded (``c_concat`` latents and / or ``c_crossattn`` context), like the trainers pass it."""

    def __init__(self, first_stage_config, cond_stage_config=None, num_timesteps_cond=None, cond_stage_key="image",
                 cond_stage_trainable=False, concat_mode=True, cond_stage_forward=None, conditioning_key=None,
                 scale_factor=1.0, scale_by_std=False, force_null_conditioning=False, *args, **kwargs):
        if conditioning_key is None:
            conditioning_key = "concat" if concat_mode else "crossattn"
        if cond_stage_config == "__is_unconditional__" and not force_null_conditioning:
            conditioning_key = None
        super().__init__(*args, conditioning_key=conditioning_key, **kwargs)
        self.scale_factor = scale_factor
        self.first_stage_model = (first_stage_config if isinstance(first_stage_config, nn.Module)
                                  else instantiate_from_config(first_stage_config)).eval()
        for prm in self.first_stage_model.parameters():
            prm.requires_grad = False
        self.clip_denoised = False                      # ddpm.py:572

    @torch.no_grad()
    def decode_first_stage(self, z, predict_cids=False, force_not_quantize=False):
        if predict_cids:
            raise NotImplementedError("decode_first_stage(predict_cids=True) is not built in the B200 path")
        return self.first_stage_model.decode(self._scaled(z, 1. / float(self.scale_factor)))

    @torch.no_grad()
    def encode_first_stage(self, x):
        return self.first_stage_model.encode(x)

    def get_first_stage_encoding(self, encoder_posterior):
        z = encoder_posterior.sample() if hasattr(encoder_posterior, "sample") else encoder_posterior
        return self._scaled(z, float(self.scale_factor))

    def _posterior_sample(self, posterior, noise=None):
        """encode -> sample of either first-stage flavour: the in-tree KL returns the distribution itself
        (ldm/models/autoencoder.py:138-142), the diffusers one wraps it in ``.latent_dist``
        (trainers/trainer_latent_diffusion.py:185-186, ddpm.py:779-780)."""
        posterior = getattr(posterior, "latent_dist", posterior)
        if noise is not None:
            return posterior.sample(noise)
        return posterior.sample()

    @torch.no_grad()
    def get_input(self, batch, k, return_first_stage_outputs=False, force_c_encode=False, cond_key=None,
                  return_original_cond=False, bs=None, return_x=False, noise=None):
        """LatentDiffusion.get_input (ddpm.py:771-824): target image -> first-stage posterior sample -> scaled latent z, plus
        the conditioning ``c`` (the un-encoded ``batch[cond_key]``: the trainers' cond stage is the identity / the concat
        glue below).  ``noise``: optional N(0,1) tensor for the posterior sample (reproducible fixtures)."""
        x = DDPM.get_input(self, batch, k)
        if bs is not None:
            x = x[:bs]
        x = x.to(self.betas.device)
        z = self.get_first_stage_encoding(self._posterior_sample(self.encode_first_stage(x), noise)).detach()
        c = xc = None
        if self.model.conditioning_key is not None and not getattr(self, "force_null_conditioning", False):
            cond_key = cond_key or getattr(self, "cond_stage_key", "image")
            xc = x if cond_key == k else batch[cond_key]
            if torch.is_tensor(xc):
                xc = xc.to(self.betas.device)
            c = xc
            if bs is not None and torch.is_tensor(c):
                c = c[:bs]
        out = [z, c]
        if return_first_stage_outputs:
            out.extend([x, self.decode_first_stage(z)])
        if return_x:
            out.append(x)
        if return_original_cond:
            out.append(xc)
        return out

    @torch.no_grad()
    def get_concat_conditioning(self, img, n_keys: int = 1, bs=None, noises=None):
        """The latent-conditioning glue of LatentDiffusionModel.get_input (trainers/trainer_latent_diffusion.py:153-191,
        conditioning_key 'concat'): the conditioning image holds ``n_keys`` contrasts stacked along C; EVERY contrast is
        pushed through the first-stage encoder separately, its posterior sampled and scaled, and the latents are
        concatenated along C -> ``{"c_concat": [c_cat]}``.  Here the ``n_keys`` encodes run as ONE batched encoder pass
        ([n_keys * B, C/n_keys, H, W]) on the same kernels as ``encode``; ``noises``: optional per-contrast N(0,1)
        tensors."""
        if bs is not None:
            img = img[:bs]
        img = img.to(self.betas.device).float()
        if img.shape[1] > 3 and n_keys > 1:            # :157-158
            parts = list(torch.chunk(img, n_keys, 1))
        else:
            parts = [img]
        b = img.shape[0]
        stacked = torch.cat(parts, 0).contiguous() if len(parts) > 1 else parts[0].contiguous()
        noise = None if noises is None else torch.cat([nz.to(img.device) for nz in noises], 0)
        z = self.get_first_stage_encoding(self._posterior_sample(self.encode_first_stage(stacked), noise))
        c_cat = torch.cat([z[i * b:(i + 1) * b] for i in range(len(parts))], dim=1) if len(parts) > 1 else z
        return {"c_concat": [c_cat]}

    @staticmethod
    def _scaled(z, s: float):
        """s * z (ddpm.py:808, 830): one fp32 multiply per element in the fused kernel; the identity when s == 1."""
        if s == 1.0:
            return z
        n = z.shape[0]
        c = torch.full((n,), s, dtype=torch.float32, device=z.device)
        return ops.lincomb2(z, z, c, torch.zeros_like(c))

    @torch.no_grad()
    def p_sample_loop(self, cond, shape, return_intermediates=False, x_T=None, verbose=True, callback=None,
                      timesteps=None, quantize_denoised=False, mask=None, x0=None, img_callback=None, start_T=None,
                      log_every_t=None, noises=None, record=None):
        if self.parameterization == "v":
            raise NotImplementedError()                 # like LatentDiffusion.p_mean_variance (ddpm.py:941-946)
        if quantize_denoised or mask is not None:
            raise NotImplementedError("quantize_denoised / mask are not built in the B200 path")
        n = self.num_timesteps if timesteps is None else timesteps
        if start_T is not None:
            n = min(n, start_T)
        steps = list(reversed(range(0, n)))
        if log_every_t:
            self.log_every_t = log_every_t
        return DDPM.p_sample_loop(self, shape, cond, return_intermediates=return_intermediates, x_T=x_T, noises=noises,
                                  steps=steps, record=record)

    @torch.no_grad()
    def sample(self, cond, batch_size=16, return_intermediates=False, x_T=None, verbose=True, timesteps=None,
               quantize_denoised=False, mask=None, x0=None, shape=None, **kwargs):
        if shape is None:
            shape = (batch_size, self.channels, self.image_size, self.image_size)
        return self.p_sample_loop(cond, shape, return_intermediates=return_intermediates, x_T=x_T, verbose=verbose,
                                  timesteps=timesteps, quantize_denoised=quantize_denoised, mask=mask, x0=x0, **kwargs)


# ---------------------------------------------------------------------------------------------
# ldm: DDIMSampler (V1)
# ---------------------------------------------------------------------------------------------
class DDIMSampler(object):
    def __init__(self, model, schedule="linear", device=None, **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self.device = device if device is not None else model.betas.device
        self.use_graph = getattr(model, "use_graph", True)

    def register_buffer(self, name, attr):
        if type(attr) == torch.Tensor and attr.device != self.device:
            attr = attr.to(self.device)
        setattr(self, name, attr)

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        # The tables are evaluated with numpy ufuncs on torch tensors, like the reference: numpy's fp32 sqrt / log and its
        # float64-ndarray x fp32-tensor promotion decide the last bit (torch.sqrt differs by 1 ulp on some entries —
        # measured), and tests/test_host_logic.py pins them bit-exactly.  NumPy 2 flags Tensor.__array_wrap__'s signature.
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", DeprecationWarning)
            self._make_schedule(ddim_num_steps, ddim_discretize, ddim_eta, verbose)

    def _make_schedule(self, ddim_num_steps, ddim_discretize, ddim_eta, verbose):
        self.ddim_timesteps = make_ddim_timesteps(ddim_discretize, ddim_num_steps, self.ddpm_num_timesteps, verbose)
        ac = self.model.alphas_cumprod
        assert ac.shape[0] == self.ddpm_num_timesteps, "alphas have to be defined for each timestep"
        to_torch = lambda x: x.clone().detach().to(torch.float32).to(self.model.device)
        self.register_buffer("betas", to_torch(self.model.betas))
        self.register_buffer("alphas_cumprod", to_torch(ac))
        self.register_buffer("alphas_cumprod_prev", to_torch(self.model.alphas_cumprod_prev))
        self.register_buffer("sqrt_alphas_cumprod", to_torch(np.sqrt(ac.cpu())))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", to_torch(np.sqrt(1. - ac.cpu())))
        self.register_buffer("log_one_minus_alphas_cumprod", to_torch(np.log(1. - ac.cpu())))
        self.register_buffer("sqrt_recip_alphas_cumprod", to_torch(np.sqrt(1. / ac.cpu())))
        self.register_buffer("sqrt_recipm1_alphas_cumprod", to_torch(np.sqrt(1. / ac.cpu() - 1)))
        sig, a, a_prev = make_ddim_sampling_parameters(ac.cpu(), self.ddim_timesteps, ddim_eta, verbose)
        self.register_buffer("ddim_sigmas", sig)
        self.register_buffer("ddim_alphas", a)
        self.register_buffer("ddim_alphas_prev", a_prev)
        self.register_buffer("ddim_sqrt_one_minus_alphas", np.sqrt(1. - a))
        acp, acc = self.model.alphas_cumprod_prev.cpu(), ac.cpu()
        self.register_buffer("ddim_sigmas_for_original_num_steps",
                             ddim_eta * torch.sqrt((1 - acp) / (1 - acc) * (1 - acc / acp)))

    def _coef(self, use_original_steps: bool, temperature: float) -> (Tensor, List[int]):
        """Rows in ITERATION order (row i <-> index = total-1-i), ddim.py:229-258."""
        m = self.model
        if use_original_steps:
            index = torch.arange(self.ddpm_num_timesteps - 1, -1, -1)
            tsteps = index.clone()
            a_t = _f32(m.alphas_cumprod)[index]
            a_prev = _f32(m.alphas_cumprod_prev)[index]
            s1m = _f32(m.sqrt_one_minus_alphas_cumprod)[index]
            sigma = _f32(self.ddim_sigmas_for_original_num_steps)[index]
        else:
            total = self.ddim_timesteps.shape[0]
            index = torch.arange(total - 1, -1, -1)
            tsteps = torch.from_numpy(np.ascontiguousarray(self.ddim_timesteps))[index]
            a_t = _f32(self.ddim_alphas)[index]
            a_prev = _f32(self.ddim_alphas_prev)[index]
            s1m = _f32(self.ddim_sqrt_one_minus_alphas)[index]
            sigma = _f32(self.ddim_sigmas)[index]
        sa = _f32(m.sqrt_alphas_cumprod)[tsteps]
        s1 = _f32(m.sqrt_one_minus_alphas_cumprod)[tsteps]
        cols = {0: sa, 1: s1, 2: a_t.sqrt(), 3: s1m, 4: a_prev.sqrt(), 5: (1. - a_prev - sigma ** 2).sqrt(),
                6: sigma, 7: torch.full_like(sigma, float(temperature))}
        return _coef_table(cols, len(index)), [int(v) for v in tsteps]

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None,
               img_callback=None, clip_denoised=True, quantize_x0=False, eta=0., mask=None, x0=None, temperature=1.,
               noise_dropout=0., score_corrector=None, corrector_kwargs=None, verbose=True, x_T=None,
               log_every_t=100, unconditional_guidance_scale=1., unconditional_conditioning=None,
               dynamic_threshold=None, ucg_schedule=None, **kwargs):
        if conditioning is not None:
            ctmp = conditioning
            if isinstance(ctmp, dict):
                ctmp = ctmp[list(ctmp.keys())[-1]]
            while isinstance(ctmp, list):
                ctmp = ctmp[0]
            if ctmp.shape[0] != batch_size:
                print(f"Warning: Got {ctmp.shape[0]} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose,
                           ddim_discretize=kwargs.get("ddim_discretize", "uniform"))
        C, H, W = shape
        size = (batch_size, C, H, W)
        print(f"Data shape for DDIM sampling is {size}, eta {eta}")
        for name, val in (("mask", mask), ("score_corrector", score_corrector), ("dynamic_threshold", dynamic_threshold),
                          ("ucg_schedule", ucg_schedule)):
            if val is not None:
                raise NotImplementedError(f"DDIMSampler.sample({name}=...) is not built in the B200 path")
        if quantize_x0 or noise_dropout > 0.:
            raise NotImplementedError("quantize_x0 / noise_dropout are not built in the B200 path")
        return self.ddim_sampling(conditioning, size, callback=callback, img_callback=img_callback,
                                  clip_denoised=clip_denoised,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning,
                                  ddim_use_original_steps=kwargs.get("ddim_use_original_steps", False),
                                  temperature=temperature, x_T=x_T, log_every_t=log_every_t,
                                  noises=kwargs.get("noises"), record=kwargs.get("record"))

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      clip_denoised=True, img_callback=None, log_every_t=100, temperature=1., noises=None,
                      record=None, unconditional_guidance_scale=1., unconditional_conditioning=None, **unused):
        device = self.model.betas.device
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device)
        coef, tsteps = self._coef(ddim_use_original_steps, temperature)
        if timesteps is not None:       # run only the LAST (lowest-noise) steps of the schedule, ddim.py:143-147
            if ddim_use_original_steps:
                keep = len(range(0, int(timesteps)))
            else:
                n = self.ddim_timesteps.shape[0]
                keep = len(range(n)[:int(min(timesteps / n, 1) * n) - 1])
            coef, tsteps = coef[len(tsteps) - keep:].contiguous(), tsteps[len(tsteps) - keep:]
        total_steps = len(tsteps)
        print(f"Running DDIM Sampling with {total_steps} timesteps")
        t_table = torch.tensor(tsteps, dtype=torch.float32, device=device)
        cfg = {}
        m = self.model
        if unconditional_conditioning is not None and unconditional_guidance_scale != 1.:
            cfg = dict(cfg_scale=unconditional_guidance_scale, uncond=host_cond_list(m, unconditional_conditioning),
                       uncond_context=host_context(m, unconditional_conditioning))
        eng = StepEngine(_engine_model_fn(host_denoiser(m)), L.SAMPLER_LDM_DDIM, _PARAM[m.parameterization],
                         coef.to(device), t_table, clip_denoised, context=host_context(m, cond), **cfg,
                         **host_engine_args(m))
        self.last_engine = eng
        m.__dict__["last_engine"] = eng
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        if total_steps == 0:
            return img, intermediates

        def on_step(i, x, x0):
            index = total_steps - i - 1
            if callback:
                callback(i)
            if img_callback:
                img_callback(x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(x.clone())
                intermediates["pred_x0"].append(x0.clone())

        out = eng.run(img, host_cond_list(m, cond), total_steps, noises=noises, use_graph=self.use_graph,
                      on_step=on_step, want_x0=True, record=record)
        return out, intermediates

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, clip_denoised=True,
                      dynamic_threshold=None, noise=None):
        """Single step (ddim.py:188-261): network call(s) + fused update at table row ``index``."""
        for name, val in (("score_corrector", score_corrector), ("dynamic_threshold", dynamic_threshold)):
            if val is not None:
                raise NotImplementedError(f"DDIMSampler.p_sample_ddim({name}=...) is not built in the B200 path")
        if quantize_denoised or noise_dropout > 0.:
            raise NotImplementedError("quantize_denoised / noise_dropout are not built in the B200 path")
        coef, _ = self._coef(use_original_steps, temperature)
        total = coef.shape[0]
        if unconditional_conditioning is None or unconditional_guidance_scale == 1.:
            out = self.model.apply_model(x, t, c)
        else:                           # classifier-free guidance, ddim.py:197-219: [uncond; cond] in one batch
            x_in, t_in = torch.cat([x] * 2), torch.cat([t] * 2)
            if isinstance(c, dict):
                assert isinstance(unconditional_conditioning, dict)
                c_in = {}
                for k in c:
                    if isinstance(c[k], list):
                        c_in[k] = [torch.cat([unconditional_conditioning[k][i], c[k][i]]) for i in range(len(c[k]))]
                    else:
                        c_in[k] = torch.cat([unconditional_conditioning[k], c[k]])
            elif isinstance(c, list):
                assert isinstance(unconditional_conditioning, list)
                c_in = [torch.cat([unconditional_conditioning[i], c[i]]) for i in range(len(c))]
            else:
                c_in = torch.cat([unconditional_conditioning, c])
            both = self.model.apply_model(x_in, t_in, c_in)
            n = x.shape[0]
            out = ops.cfg_combine(both[:n].contiguous(), both[n:].contiguous(), float(unconditional_guidance_scale))
        z = noise if noise is not None else noise_like(x.shape, x.device, repeat_noise)
        return ops.sampler_update(L.SAMPLER_LDM_DDIM, _PARAM[self.model.parameterization], x.float().contiguous(),
                                  out.contiguous(), z.contiguous(), coef.to(x.device), step=total - 1 - index,
                                  clip_denoised=clip_denoised)


# ---------------------------------------------------------------------------------------------
# guided-diffusion process (V3 / V4)
# ---------------------------------------------------------------------------------------------
def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    """training_project/utils/gaussian_diffusion.py:33-57."""
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        return np.array([min(1 - f((i + 1) / num_diffusion_timesteps) / f(i / num_diffusion_timesteps), 0.999)
                         for i in range(num_diffusion_timesteps)])
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def space_timesteps(num_timesteps, section_counts):
    """training_project/utils/respace.py:7-60."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired_count = int(section_counts[len("ddim"):])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == desired_count:
                    return set(range(0, num_timesteps, i))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per = num_timesteps // len(section_counts)
    extra = num_timesteps % len(section_counts)
    start_idx, all_steps = 0, []
    for i, section_count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < section_count:
            raise ValueError(f"cannot divide section of {size} steps into {section_count}")
        frac_stride = 1 if section_count <= 1 else (size - 1) / (section_count - 1)
        cur_idx = 0.0
        for _ in range(section_count):
            all_steps.append(start_idx + round(cur_idx))
            cur_idx += frac_stride
        start_idx += size
    return set(all_steps)


def gd_conditioning(model, model_kwargs):
    """What ``model(x, t, **model_kwargs)`` (gaussian_diffusion.py:276) makes of the caller's ``model_kwargs`` when
    ``model`` is a DiffusionWrapper (ours or the reference's; trainer_use_gaussian_diff.py:603-620 passes
    ``dict(c_concat=[images(, edge)])``): (tensors concatenated to x, cross-attention context)."""
    if not model_kwargs:
        return None, None
    unknown = set(model_kwargs) - {"c_concat", "c_crossattn"}
    if unknown:
        raise NotImplementedError(f"B200 path: model_kwargs {sorted(unknown)} are not built")
    key = host_conditioning_key(model) if hasattr(model, "diffusion_model") else "hybrid"
    cond = context = None
    if model_kwargs.get("c_concat") is not None and key in ("concat", "hybrid"):
        cond = list(model_kwargs["c_concat"])
    if model_kwargs.get("c_crossattn") is not None and key in ("crossattn", "hybrid"):
        context = torch.cat(list(model_kwargs["c_crossattn"]), 1).float().contiguous()
    return cond, context


class GaussianDiffusion:
    """Sampling half of the modified guided-diffusion process.  ``model_mean_type`` in
    {"eps", "x0"}; ``parameterization="v"`` overrides it like the reference (:338-342);
    ``model_var_type`` in {"learned_range", "fixed_large", "fixed_small"}."""

    def __init__(self, *, betas, model_mean_type="eps", model_var_type="fixed_large", loss_type=None,
                 rescale_timesteps=False, parameterization="eps"):
        self.model_mean_type, self.model_var_type = model_mean_type, model_var_type
        self.loss_type, self.rescale_timesteps, self.parameterization = loss_type, rescale_timesteps, parameterization
        betas = np.array(betas, dtype=np.float64)
        self.betas = betas
        assert len(betas.shape) == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        self.use_graph = True

    # model timestep fed to the network for process index t (identity here; SpacedDiffusion remaps)
    def _model_timestep(self, t: int) -> float:
        if self.rescale_timesteps:
            return float(np.float32(t) * np.float32(1000.0 / self.num_timesteps))
        return float(t)

    def _param(self) -> int:
        if self.parameterization == "v":
            return L.PARAM_V
        return L.PARAM_X0 if self.model_mean_type == "x0" else L.PARAM_EPS

    def _coef(self, eta: float) -> Tensor:
        """Rows in ITERATION order (row i <-> t = T-1-i); every scalar is float64 -> float32 exactly like
        _extract_into_tensor (:1097-1109), then combined with fp32 torch ops like the reference."""
        T = self.num_timesteps
        order = torch.arange(T - 1, -1, -1)
        g = lambda arr: _f32(arr)[order]
        if self._param() == L.PARAM_V:
            ca, cb = g(self.sqrt_alphas_cumprod), g(self.sqrt_one_minus_alphas_cumprod)
        else:
            ca, cb = g(self.sqrt_recip_alphas_cumprod), g(self.sqrt_recipm1_alphas_cumprod)
        if self.model_var_type == "learned_range":
            min_log, max_log = g(self.posterior_log_variance_clipped), g(np.log(self.betas))
        else:
            lv = (np.log(np.append(self.posterior_variance[1], self.betas[1:]))
                  if self.model_var_type == "fixed_large" else self.posterior_log_variance_clipped)
            min_log, max_log = g(lv), torch.zeros(T)
        ab, abp = g(self.alphas_cumprod), g(self.alphas_cumprod_prev)
        sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
        nonzero = (order != 0).float()
        cols = {0: ca, 1: cb, 2: g(self.posterior_mean_coef1), 3: g(self.posterior_mean_coef2), 4: min_log, 5: max_log,
                6: g(self.sqrt_recip_alphas_cumprod), 7: g(self.sqrt_recipm1_alphas_cumprod), 8: torch.sqrt(abp),
                9: torch.sqrt(1 - abp - sigma ** 2), 10: nonzero * sigma, 11: nonzero}
        return _coef_table(cols, T)

    def _loop(self, mode, model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device, progress,
              eta, noises=None, record=None):
        if denoised_fn is not None or cond_fn is not None:
            raise NotImplementedError("denoised_fn / cond_fn are not built in the B200 path")
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        img = noise.to(device) if noise is not None else torch.randn(*shape, device=device)
        T = self.num_timesteps
        t_table = torch.tensor([self._model_timestep(t) for t in range(T - 1, -1, -1)], dtype=torch.float32,
                               device=device)
        cond, context = gd_conditioning(model, model_kwargs)
        net = host_denoiser(model)
        eng = StepEngine(_engine_model_fn(net), mode, self._param(), self._coef(eta).to(device), t_table, clip_denoised,
                         learned_range=self.model_var_type == "learned_range", context=context,
                         **_graph_cache_args(self.__dict__.setdefault("_graphs", {}), net))
        self.last_engine = eng
        return eng.run(img, cond, T, noises=noises, use_graph=self.use_graph, want_x0=record is not None,
                       record=record)

    # ---- forward process + single reverse steps (the pieces the loops above fuse) ----
    def _gather(self, arr, t: Tensor) -> Tensor:
        """_extract_into_tensor (:1097-1109) without the broadcast: float64 table -> float32 -> [t]."""
        return _f32(arr).to(t.device)[t.reshape(-1).long()]

    @torch.no_grad()
    def q_sample(self, x_start, t, noise=None):
        """sqrt(abar_t) x_0 + sqrt(1 - abar_t) noise (:199-215), per-sample t."""
        noise = torch.randn_like(x_start) if noise is None else noise
        assert noise.shape == x_start.shape
        return ops.lincomb2(x_start, noise, self._gather(self.sqrt_alphas_cumprod, t),
                            self._gather(self.sqrt_one_minus_alphas_cumprod, t))

    @torch.no_grad()
    def get_v(self, x, noise, t):
        """sqrt(abar_t) noise - sqrt(1 - abar_t) x (:376-380)."""
        return ops.lincomb2(noise, x, self._gather(self.sqrt_alphas_cumprod, t),
                            self._gather(self.sqrt_one_minus_alphas_cumprod, t), subtract=True)

    @torch.no_grad()
    def predict_start_from_z_and_v(self, x_t, t, v):
        """sqrt(abar_t) x_t - sqrt(1 - abar_t) v (:239-245)."""
        return ops.lincomb2(x_t, v, self._gather(self.sqrt_alphas_cumprod, t),
                            self._gather(self.sqrt_one_minus_alphas_cumprod, t), subtract=True)

    def _single_step(self, mode, model, x, t, clip_denoised, denoised_fn, cond_fn, model_kwargs, eta, noise=None):
        if denoised_fn is not None or cond_fn is not None:
            raise NotImplementedError("denoised_fn / cond_fn are not built in the B200 path")
        tt = t.reshape(-1)
        ti = int(tt[0])
        if not bool((tt == ti).all()):
            raise NotImplementedError("B200 path: one timestep per p_sample / ddim_sample call")
        net = host_denoiser(model)
        _engine_model_fn(net)           # loud rejection of a non-engine denoiser
        cond, context = gd_conditioning(model, model_kwargs)
        xc = x.float()
        if cond:
            xc = torch.cat([xc] + [c.to(x.device).float() for c in cond], dim=1)
        tm = torch.full((x.shape[0],), self._model_timestep(ti), dtype=torch.float32, device=x.device)
        extra = {"return_features": False} if getattr(net, "_returns_features", False) else {}
        if context is not None:
            extra["context"] = context.to(x.device)
        out = net(xc.contiguous(), tm, **extra)
        out = out[0] if isinstance(out, tuple) else out
        cache = self.__dict__.setdefault("_coef_cache", {})
        key = (float(eta), str(x.device))
        if key not in cache:
            cache[key] = self._coef(eta).to(x.device)
        z = torch.randn_like(x) if noise is None else noise.to(x.device)     # drawn even at t == 0, like :455 / :653
        x_prev, x0 = ops.sampler_update(mode, self._param(), x.float().contiguous(), out.contiguous(),
                                        z.float().contiguous(), cache[key], step=self.num_timesteps - 1 - ti,
                                        clip_denoised=clip_denoised, learned_range=self.model_var_type == "learned_range")
        return {"sample": x_prev, "pred_xstart": x0}

    @torch.no_grad()
    def p_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None, noise=None):
        """One ancestral step (:423-466): {'sample', 'pred_xstart'}."""
        return self._single_step(L.SAMPLER_GD_DDPM, model, x, t, clip_denoised, denoised_fn, cond_fn, model_kwargs, 0.0,
                                 noise)

    @torch.no_grad()
    def ddim_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None, eta=0.0,
                    noise=None):
        """One DDIM step (:620-667): {'sample', 'pred_xstart'}."""
        return self._single_step(L.SAMPLER_GD_DDIM, model, x, t, clip_denoised, denoised_fn, cond_fn, model_kwargs, eta,
                                 noise)

    def _progressive(self, step_fn, model, shape, noise, device, **kw):
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        img = noise.to(device) if noise is not None else torch.randn(*shape, device=device)
        for i in range(self.num_timesteps - 1, -1, -1):
            t = torch.tensor([i] * shape[0], device=device)
            out = step_fn(model, img, t, **kw)
            yield out
            img = out["sample"]

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                  model_kwargs=None, device=None, progress=False):
        """Generator over the per-step dicts (:570-618); eager single steps (the graph-replayed loop is p_sample_loop)."""
        return self._progressive(self.p_sample, model, shape, noise, device, clip_denoised=clip_denoised,
                                 denoised_fn=denoised_fn, cond_fn=cond_fn, model_kwargs=model_kwargs)

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                     model_kwargs=None, device=None, progress=False, eta=0.0):
        """:742-789."""
        return self._progressive(self.ddim_sample, model, shape, noise, device, clip_denoised=clip_denoised,
                                 denoised_fn=denoised_fn, cond_fn=cond_fn, model_kwargs=model_kwargs, eta=eta)

    @torch.no_grad()
    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                      model_kwargs=None, device=None, progress=False, **kw):
        return self._loop(L.SAMPLER_GD_DDPM, model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs,
                          device, progress, 0.0, **kw)

    @torch.no_grad()
    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, eta=0.0, **kw):
        return self._loop(L.SAMPLER_GD_DDIM, model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs,
                          device, progress, eta, **kw)


def _dpm_solver_sample_loop(self, model, shape, clip_denoised=None, model_kwargs=None, noise=None, record=None):
    """training_project/utils/gaussian_diffusion.py:468-523 (DPM-Solver++ 2M, logSNR steps, dynamic thresholding)."""
    from .dpm_solver import gaussian_dpm_solver_sample_loop
    return gaussian_dpm_solver_sample_loop(self, model, shape, clip_denoised, model_kwargs, noise=noise, record=record)


GaussianDiffusion.dpm_solver_sample_loop = torch.no_grad()(_dpm_solver_sample_loop)


class SpacedDiffusion(GaussianDiffusion):
    """respace.py:63-128: keep ``use_timesteps`` of a base process; the model sees the ORIGINAL timestep."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.timestep_map = []
        self.original_num_steps = len(kwargs["betas"])
        base = GaussianDiffusion(**kwargs)
        last, new_betas = 1.0, []
        for i, ac in enumerate(base.alphas_cumprod):
            if i in self.use_timesteps:
                new_betas.append(1 - ac / last)
                last = ac
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)

    def _model_timestep(self, t: int) -> float:
        v = self.timestep_map[t]
        if self.rescale_timesteps:
            return float(np.float32(v) * np.float32(1000.0 / self.original_num_steps))
        return float(v)


def create_gaussian_diffusion(*, steps=1000, learn_sigma=False, sigma_small=False, noise_schedule="linear",
                              use_kl=False, predict_xstart=False, rescale_timesteps=False,
                              rescale_learned_sigmas=False, timestep_respacing="", parameterization="eps"):
    """script_util.py:129-169."""
    betas = get_named_beta_schedule(noise_schedule, steps)
    if not timestep_respacing:
        timestep_respacing = [steps]
    var = "learned_range" if learn_sigma else ("fixed_small" if sigma_small else "fixed_large")
    return SpacedDiffusion(use_timesteps=space_timesteps(steps, timestep_respacing), betas=betas,
                           model_mean_type="x0" if predict_xstart else "eps", model_var_type=var,
                           loss_type="rescaled_kl" if use_kl else ("rescaled_mse" if rescale_learned_sigmas else "mse"),
                           rescale_timesteps=rescale_timesteps, parameterization=parameterization)


def sr_create_model_and_diffusion(args):
    """script_util.py:57-86: the trainers call this with their config object and keep only the diffusion
    (trainer_use_gaussian_diff.py:70-76, 591-596); the network comes from ``unet_config`` through DiffusionWrapper."""
    diffusion = create_gaussian_diffusion(
        steps=args.diffusion_steps, learn_sigma=args.learn_sigma, noise_schedule=args.noise_schedule,
        use_kl=args.use_kl, predict_xstart=args.predict_xstart, rescale_timesteps=args.rescale_timesteps,
        rescale_learned_sigmas=args.rescale_learned_sigmas, timestep_respacing=args.timestep_respacing,
        parameterization=args.parameterization)
    return None, diffusion
