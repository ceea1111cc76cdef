"""Builders for the REFERENCE's own host objects (test / baseline infrastructure; never imported by the product).

``trainers.trainer_ds_diff.DSDiffModel`` (an ``ldm`` ``DDPM`` subclass; ``DDIMSampler(self)`` at trainer_ds_diff.py:672) and
``trainers.trainer_use_gaussian_diff.TryTrainerDiffusion`` (``self.model = DiffusionWrapper(...)``, ``sample_fn(self.model,
...)`` at :86, :615) are constructed through ``oracle/_ref_import`` from the unmodified reference — ``/root/reference`` in
the authoring container, the byte-compiled ``oracle/_ref`` on the GPU box — with the two edits INTEGRATION.md §1 asks a
maintainer to make: the yaml ``unet_config.target`` points at the B200 network and the sampler names imported by the trainer
module are the B200 ones.
"""
import contextlib
import io
import warnings

import torch

import _ref_import

A = _ref_import.to_attr

DS_SMALL = dict(image_size=32, in_channels=1, model_channels=32, out_channels=1, num_res_blocks=1,
                attention_resolutions=[2, 4], channel_mult=[1, 1, 2], num_head_channels=16,
                use_new_attention_order=True, use_checkpoint=True, legacy=False)
DS_FILM2 = dict(image_size=32, in_channels=1, model_channels=32, out_channels=2, num_res_blocks=1,
                attention_resolutions=[2, 4], channel_mult=[1, 1, 2], num_head_channels=16,
                use_new_attention_order=True, use_scale_shift_norm=True, legacy=False)
UNET_SMALL = dict(image_size=32, in_channels=4, model_channels=32, out_channels=1, num_res_blocks=1,
                  attention_resolutions=[2, 4], channel_mult=[1, 2, 4], num_head_channels=16,
                  use_new_attention_order=True, use_checkpoint=True, legacy=False)

REF_TARGETS = {"ds": "UNet_DS_Diff.model.DSUnetModel", "unet": "ldm.modules.diffusionmodules.openaimodel.UNetModel"}
B200_TARGETS = {"ds": "diffusion_models_dsdiff_b200.dsunet.DSUnetModel", "unet": "diffusion_models_dsdiff_b200.unet.UNetModel"}


def available() -> bool:
    return _ref_import.available()


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def _trainer_fields(sampler="ddim", steps=5):
    """The training-side fields the trainers' constructors read; none touches the sampling path."""
    return dict(sampler_setting=dict(sampler=sampler, sample_steps=steps, ddim_use_original_steps=False, ddim_eta=0.0),
                num_workers=0, train_batch_size=2, val_batch_size=2, dataset_type="h5", fold_K=5, fold_idx=0,
                train_keys=["image", "pet"], contrast_lambda=0.1, beta1=0.9, beta2=0.999, lr=1e-4, lr_low=1e-6, seed=2024,
                augmentation_prob=0.5, num_epochs=1, lr_warm_epoch=0, lr_cos_epoch=1, h5_2d_img_dir="/tmp/none",
                data_name="PET", filepath_img="/tmp/none", root_dir="/tmp/none")


def install():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with _quiet():
            _ref_import.install()
            import trainers.trainer_ds_diff as tds
            import trainers.trainer_use_gaussian_diff as tgd
    return tds, tgd


def use_samplers(b200: bool):
    """The import swap of INTEGRATION.md §1: the sampler names the trainer modules look up at call time are the B200
    drop-ins (``b200``) or the reference's own."""
    tds, tgd = install()
    import ldm.models.diffusion.ddim as ref_ddim
    import ldm.models.diffusion.dpm_solver_new.sampler as ref_dpm
    import training_project.utils.script_util as ref_su
    from diffusion_models_dsdiff_b200 import DDIMSampler, DPMSolverSampler
    from diffusion_models_dsdiff_b200.samplers import sr_create_model_and_diffusion
    tds.DDIMSampler = DDIMSampler if b200 else ref_ddim.DDIMSampler
    tds.DPMSolverSampler = DPMSolverSampler if b200 else ref_dpm.DPMSolverSampler
    tgd.sr_create_model_and_diffusion = sr_create_model_and_diffusion if b200 else ref_su.sr_create_model_and_diffusion
    return tds, tgd


def ds_diff_model(target: str, params: dict, parameterization="v", sampler="ddim", steps=5, b200_samplers=True):
    """The reference's DSDiffModel (ldm DDPM + Lightning trainer) around ``target``; with ``b200_samplers`` the trainer
    module's DDIMSampler / DPMSolverSampler names are the B200 drop-ins (the documented import swap)."""
    tds, _ = use_samplers(b200_samplers)
    cfg = A(_trainer_fields(sampler, steps))
    with _quiet():
        return tds.DSDiffModel(cfg, unet_config=A({"target": target, "params": dict(params)}), timesteps=2000,
                               beta_schedule="linear", linear_start=1e-4, linear_end=2e-2,
                               parameterization=parameterization, conditioning_key="concat", use_ema=False,
                               clip_denoised=True, log_every_t=100, image_size=32, channels=1, loss_type="l2").eval()


def gaussian_trainer(target: str, params: dict, sampler="ddim", steps=5, b200_samplers=True):
    """The reference's TryTrainerDiffusion (guided-diffusion process + reference DiffusionWrapper) around ``target``."""
    _, tgd = use_samplers(b200_samplers)
    cfg = dict(_trainer_fields(sampler, steps))
    cfg.update(learn_sigma=True, diffusion_steps=1000, noise_schedule="linear", use_kl=False, predict_xstart=False,
               rescale_timesteps=False, rescale_learned_sigmas=False, timestep_respacing="", parameterization="v",
               image_size=32, conditioning_key="concat", schedule_sampler="uniform", clip_denoised=True,
               disentangle_distance="l2", use_ema=False, model_type="unet",
               unet_config={"target": target, "params": dict(params)})
    with _quiet():
        g = tgd.TryTrainerDiffusion(A(cfg)).eval()
    # on_predict_start (trainer_use_gaussian_diff.py:586-600) minus its directory listing
    d = g.diffusion_for_val
    g.sample_fn = (d.ddim_sample_loop if sampler == "ddim" else
                   d.dpm_solver_sample_loop if sampler == "dpm" else d.p_sample_loop)
    return g


def batch(n=2, size=32, seed=7):
    g = torch.Generator().manual_seed(seed)
    return {"image": (torch.randn(n, 3, size, size, generator=g) * 0.5).clamp_(-1, 1),
            "pet": torch.randn(n, 1, size, size, generator=g),
            "path": [f"/data/images_ts_256/id{i // 2}/slice_{i}.h5" for i in range(n)]}


# ---------------------------------------------------------------------------------------------
# the reference's own path for bench.py's reference arm / cpu_baseline / gpu_eager_reference legs
# ---------------------------------------------------------------------------------------------
def reference_network(cls_name: str, hp: dict):
    """The reference's own denoiser class for a bench workload ("DSUnetModel" / "UNetModel"), random-init."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with _quiet():
            _ref_import.install()
            if cls_name == "DSUnetModel":
                from UNet_DS_Diff.model import DSUnetModel as Ref
            else:
                from ldm.modules.diffusionmodules.openaimodel import UNetModel as Ref
            return Ref(**hp).eval()


def reference_ddim_host(cls_name: str, hp: dict, state_dict: dict, device, image_size=256):
    """(DSDiffModel, DDIMSampler) of the UNMODIFIED reference — its trainer class (ldm DDPM + the conditional / v-prediction
    overrides, trainers/trainer_ds_diff.py:698-755), its own network and its own DDIMSampler — for the ldm DDIM-50 v-param
    T=2000 workload of bench.py, with ``state_dict`` loaded into the network."""
    key = "ds" if cls_name == "DSUnetModel" else "unet"
    m = ds_diff_model(REF_TARGETS[key], hp, parameterization="v", sampler="ddim", steps=50, b200_samplers=False)
    m.model.diffusion_model.load_state_dict(state_dict, strict=True)
    m = m.to(device).eval()
    import ldm.models.diffusion.ddim as ref_ddim
    with _quiet():
        sampler = ref_ddim.DDIMSampler(m, device=torch.device(device))
        sampler.make_schedule(ddim_num_steps=50, ddim_eta=0.0, verbose=False)
    return m, sampler
