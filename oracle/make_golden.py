"""Generate tests/golden/*.npz by running the UNMODIFIED reference (test infrastructure).

Run in the authoring container only (needs /root/reference):

    python oracle/make_golden.py

The reference modules are imported from /root/reference behind the inert stubs of
``oracle/_ref_import.py``; every parameter is overwritten with the key-seeded tensors of
``dsdiff_oracle.synth_tensor`` (a fresh reference network outputs exactly 0 because of
zero_module), the reference's own forward / sampler code is executed on CPU fp32, and inputs +
outputs are written as small fixtures.  ``tests/test_oracle_golden.py`` replays them against the
oracle restatement; the GPU parity tests replay them against the CUDA path.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import  # noqa: E402
import dsdiff_oracle as O  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

UNET_SMALL = dict(image_size=32, in_channels=4, model_channels=32, out_channels=1, num_res_blocks=1,
                  attention_resolutions=[2, 4], channel_mult=[1, 2, 4], num_head_channels=16,
                  use_new_attention_order=True, use_checkpoint=True, legacy=False)
UNET_LEGACY_FILM = dict(image_size=32, in_channels=3, model_channels=64, out_channels=2, num_res_blocks=2,
                        attention_resolutions=[2], channel_mult=[1, 2], num_heads=2,
                        use_new_attention_order=False, use_scale_shift_norm=True, legacy=False)
DS_SMALL = dict(image_size=32, in_channels=1, model_channels=32, out_channels=1, num_res_blocks=1,
                attention_resolutions=[2, 4], channel_mult=[1, 1, 2], num_head_channels=16,
                use_new_attention_order=True, use_checkpoint=True, legacy=False)
DS_FILM2 = dict(image_size=32, in_channels=1, model_channels=32, out_channels=2, num_res_blocks=1,
                attention_resolutions=[2, 4], channel_mult=[1, 1, 2], num_head_channels=16,
                use_new_attention_order=True, use_scale_shift_norm=True, legacy=False)


def load_synth(module: torch.nn.Module) -> dict:
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    module.load_state_dict(O.synth_state_dict(shapes), strict=True)
    module.eval()
    return shapes


def rnd(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def save(name, **arrs):
    os.makedirs(GOLD, exist_ok=True)
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in out.items()})


def main():
    _ref_import.install()
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    from UNet_DS_Diff.model import DSUnetModel
    import ldm.models.diffusion.ddpm as ddpm
    from ldm.models.diffusion.ddim import DDIMSampler
    from training_project.utils.script_util import create_gaussian_diffusion
    from trainers.trainer_ds_diff import DSDiffModel

    manifest = {}
    torch.manual_seed(0)

    # ---- network forwards ---------------------------------------------------------------
    with torch.no_grad():
        m = UNetModel(**UNET_SMALL)
        manifest["unet_small"] = {"hp": UNET_SMALL, "shapes": load_synth(m)}
        x = rnd(11, 2, 4, 32, 32)
        t = torch.tensor([37, 1503])
        save("unet_small", x=x, t=t, y=m(x, t))

        m2 = UNetModel(**UNET_LEGACY_FILM)
        manifest["unet_legacy_film"] = {"hp": UNET_LEGACY_FILM, "shapes": load_synth(m2)}
        x = rnd(12, 2, 3, 16, 16)
        t = torch.tensor([12.5, 947.25])            # float (rescaled) timesteps
        save("unet_legacy_film", x=x, t=t, y=m2(x, t))

        ds = DSUnetModel(**DS_SMALL)
        manifest["ds_small"] = {"hp": DS_SMALL, "shapes": load_synth(ds)}
        x = rnd(13, 2, 4, 32, 32)
        t = torch.tensor([1961, 3])
        y, feats = ds(x, t)
        save("ds_small", x=x, t=t, y=y, style0=feats["style"][0], content2=feats["content"][2],
             anatomy1=feats["anatomy"][1], lesion0=feats["lesion"][0],
             nsc0=feats["n_style_content"][0], nsc1=feats["n_style_content"][1],
             nsc2=feats["n_style_content"][2], nsc3=feats["n_style_content"][3])
        x2 = rnd(14, 1, 2, 32, 32)                   # 2-channel input => zero-filled streams
        save("ds_small_2ch", x=x2, t=torch.tensor([500]), y=ds(x2, torch.tensor([500]))[0])

        ds2 = DSUnetModel(**DS_FILM2)
        manifest["ds_film2"] = {"hp": DS_FILM2, "shapes": load_synth(ds2)}

    # ---- V1 / V2: ldm DDPM + DDIMSampler, through the trainer overrides ----------------------
    # DSDiffModel.__init__ needs a full training config; the sampling overrides only touch
    # DDPM state, so graft them onto a plain ldm DDPM built from the yaml fields.
    def make_ddpm(unet_target, unet_params, parameterization, timesteps):
        cfg = _ref_import.to_attr({"target": unet_target, "params": dict(unet_params)})
        model = ddpm.DDPM(unet_config=cfg, timesteps=timesteps, beta_schedule="linear",
                          linear_start=1e-4, linear_end=2e-2, parameterization=parameterization,
                          conditioning_key="concat", use_ema=False, clip_denoised=True,
                          log_every_t=1, image_size=32, channels=1, loss_type="l2")
        for name in ("apply_model", "p_mean_variance", "p_sample", "p_sample_loop", "sample"):
            setattr(model, name, getattr(DSDiffModel, name).__get__(model))
        load_synth(model.model.diffusion_model)
        model.eval()
        return model

    for param in ("v", "eps"):
        model = make_ddpm("ldm.modules.diffusionmodules.openaimodel.UNetModel", UNET_SMALL, param, 2000)
        cond = torch.clamp(rnd(2024, 2, 3, 32, 32) * 0.5, -1, 1)
        x_T = rnd(2025, 2, 1, 32, 32)
        for eta in (0.0, 0.7):
            S = 5
            outs = []
            orig = model.apply_model
            model.apply_model = lambda x, t, c, _o=orig: (outs.append(_o(x, t, c)) or outs[-1])
            torch.manual_seed(2026)
            sampler = DDIMSampler(model, device=torch.device("cpu"))
            with torch.no_grad():
                img, inter = sampler.sample(S, 2, (1, 32, 32), cond, eta=eta, x_T=x_T, verbose=False,
                                            log_every_t=1)
            model.apply_model = orig
            torch.manual_seed(2026)
            noises = torch.stack([torch.randn(2, 1, 32, 32) for _ in range(S)])
            save(f"v1_ddim_{param}_eta{eta}", cond=cond, x_T=x_T, noises=noises, final=img,
                 model_out=torch.stack(outs), x_inter=torch.stack(inter["x_inter"][1:]),
                 pred_x0=torch.stack(inter["pred_x0"][1:]), S=np.int64(S), eta=np.float64(eta))

    for param in ("v", "eps", "x0"):
        T = 6
        model = make_ddpm("ldm.modules.diffusionmodules.openaimodel.UNetModel", UNET_SMALL, param, T)
        cond = torch.clamp(rnd(2024, 2, 3, 32, 32) * 0.5, -1, 1)
        torch.manual_seed(77)
        with torch.no_grad():
            img, inter = model.sample(batch_size=2, cond=cond, shape=(1, 32, 32), return_intermediates=True)
        torch.manual_seed(77)
        x_T = torch.randn(2, 1, 32, 32)
        noises = torch.stack([torch.randn(2, 1, 32, 32) for _ in range(T)])
        save(f"v2_ddpm_{param}", cond=cond, x_T=x_T, noises=noises, final=img,
             x_inter=torch.stack(inter[1:]), T=np.int64(T))

    # ---- V3 / V4: guided-diffusion SpacedDiffusion driving the DS-Diff backbone ---------------
    class Wrap(torch.nn.Module):      # DiffusionWrapper 'concat' (ddpm.py:1331-1333)
        def __init__(self, net):
            super().__init__()
            self.diffusion_model = net

        def forward(self, x, t, c_concat=None):
            return self.diffusion_model(torch.cat([x] + c_concat, dim=1), t)

    for learn_sigma, net_hp, tag in ((True, DS_FILM2, "ls"), (False, DS_SMALL, "fixed")):
        with torch.no_grad():
            net = DSUnetModel(**net_hp)
            load_synth(net)
        wrapped = Wrap(net)
        for respacing, steps in (("5", 1000),):
            diff = create_gaussian_diffusion(steps=steps, learn_sigma=learn_sigma, noise_schedule="linear",
                                             rescale_timesteps=True, timestep_respacing=respacing,
                                             parameterization="v")
            cond = torch.clamp(rnd(2024, 1, 3, 32, 32) * 0.5, -1, 1)
            x_T = rnd(2025, 1, 1, 32, 32)
            S = diff.num_timesteps
            for mode, eta in (("ddim", 0.0), ("ddim", 0.5), ("ddpm", 0.0)):
                torch.manual_seed(2026)
                recs = []
                if mode == "ddim":
                    gen = diff.ddim_sample_loop_progressive(wrapped, (1, 1, 32, 32), noise=x_T,
                                                            model_kwargs={"c_concat": [cond]}, eta=eta)
                else:
                    gen = diff.p_sample_loop_progressive(wrapped, (1, 1, 32, 32), noise=x_T,
                                                         model_kwargs={"c_concat": [cond]})
                for out in gen:
                    recs.append(out)
                torch.manual_seed(2026)
                noises = torch.stack([torch.randn(1, 1, 32, 32) for _ in range(S)])
                save(f"v34_{mode}_{tag}_eta{eta}", cond=cond, x_T=x_T, noises=noises,
                     final=recs[-1]["sample"], x_inter=torch.stack([r["sample"] for r in recs]),
                     pred_x0=torch.stack([r["pred_xstart"] for r in recs]),
                     timestep_map=np.asarray(diff.timestep_map), S=np.int64(S), eta=np.float64(eta))

    # ---- schedule tables at the BASELINE configs (size-independent pins) -------------------
    model = make_ddpm("ldm.modules.diffusionmodules.openaimodel.UNetModel", UNET_SMALL, "v", 2000)
    sampler = DDIMSampler(model, device=torch.device("cpu"))
    sampler.make_schedule(50, ddim_eta=0.3, verbose=False)
    save("tables_ldm_T2000_S50", betas=model.betas, alphas_cumprod=model.alphas_cumprod,
         sqrt_ac=model.sqrt_alphas_cumprod, sqrt_1mac=model.sqrt_one_minus_alphas_cumprod,
         post_logvar=model.posterior_log_variance_clipped, coef1=model.posterior_mean_coef1,
         coef2=model.posterior_mean_coef2, ddim_timesteps=sampler.ddim_timesteps,
         ddim_alphas=np.asarray(sampler.ddim_alphas), ddim_alphas_prev=np.asarray(sampler.ddim_alphas_prev),
         ddim_sigmas=np.asarray(sampler.ddim_sigmas),
         ddim_s1m=np.asarray(sampler.ddim_sqrt_one_minus_alphas))
    for resp in ("50", "20"):
        diff = create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear",
                                         rescale_timesteps=True, timestep_respacing=resp,
                                         parameterization="v")
        save(f"tables_gd_T1000_S{resp}", timestep_map=np.asarray(diff.timestep_map), betas=diff.betas,
             alphas_cumprod=diff.alphas_cumprod, post_logvar=diff.posterior_log_variance_clipped,
             coef1=diff.posterior_mean_coef1, coef2=diff.posterior_mean_coef2,
             sqrt_recipm1=diff.sqrt_recipm1_alphas_cumprod)

    with open(os.path.join(GOLD, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=0, sort_keys=True)
    print("manifest written")


if __name__ == "__main__":
    main()
