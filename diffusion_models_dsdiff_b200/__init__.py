"""B200-native (sm_100a) sampling hot path — and the DiT training step — for larrybb626/diffusion_models_dsdiff.

Drop-in operator surface (same names, arguments and state_dict keys as the reference):
``UNetModel``, ``DSUnetModel``, ``DiT``, ``DDPM`` / ``DiffusionWrapper``, ``DDIMSampler``,
``GaussianDiffusion`` / ``SpacedDiffusion`` / ``create_gaussian_diffusion``.
All arithmetic runs in libb200diff.so (include/b200diff.h); there is no CPU fallback.
"""
from ._lib import B200Error, LIB_PATH, launch_count  # noqa: F401
from .unet import UNetModel  # noqa: F401
from .dsunet import DSUnetModel  # noqa: F401
from .dit import DiT, DiT_models  # noqa: F401
from .autoencoder import AutoencoderKL, DiagonalGaussianDistribution  # noqa: F401
from .samplers import (DDPM, DDIMSampler, DiffusionWrapper, GaussianDiffusion, LatentDiffusion, SpacedDiffusion,  # noqa: F401
                       create_gaussian_diffusion, instantiate_from_config, space_timesteps)

from .dpm_solver import DPMSolverSampler, NoiseScheduleVP  # noqa: F401,E402
from .dit_train import DiTTrainer  # noqa: F401,E402
from .h5_io import H5File, H5SliceLoader, read_h5  # noqa: F401,E402

__all__ = ["DiTTrainer", "H5File", "H5SliceLoader", "read_h5", "DPMSolverSampler", "NoiseScheduleVP", "B200Error", "LIB_PATH", "launch_count", "UNetModel", "DSUnetModel", "DiT", "DiT_models", "AutoencoderKL", "DiagonalGaussianDistribution", "DDPM", "LatentDiffusion", "DDIMSampler",
           "DiffusionWrapper", "GaussianDiffusion", "SpacedDiffusion", "create_gaussian_diffusion",
           "instantiate_from_config", "space_timesteps"]
