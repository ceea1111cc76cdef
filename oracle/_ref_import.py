"""Import shim for the UNMODIFIED reference tree (test infrastructure only).

Used ONLY by ``oracle/make_golden.py`` (run in the authoring container, where
``/root/reference`` exists) to pin the oracle restatement against the real
reference code.  Nothing in the product package, ``bench.py`` or the ``-m gpu``
tests imports this file: ``/root/reference`` does not exist on the GPU box.

The reference imports a number of third-party packages that are absent from
this image (pytorch_lightning, lightning, omegaconf, monai, SimpleITK,
diffusers, timm, ...).  None of them does arithmetic on the sampling hot path,
so they are replaced with inert ``sys.modules`` stubs.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types

import torch.nn as nn

REFERENCE_ROOT = os.environ.get("DSDIFF_REFERENCE_ROOT", "/root/reference")


class _Anything:
    """Callable/attribute sink: every attribute is again a sink."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = f"{self.__name__}.{name}"
        if full in sys.modules:
            return sys.modules[full]
        # a distinct class per name, so `class X(Randomizable, MapTransform)` works
        cls = type(name, (_Anything,), {})
        setattr(self, name, cls)
        return cls


def _stub(name: str) -> types.ModuleType:
    if name in sys.modules:
        return sys.modules[name]
    m = _StubModule(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__path__ = []  # behave like a package so sub-imports resolve
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(_stub(parent), child, m)
    return m


class _LightningModule(nn.Module):
    """pl.LightningModule := nn.Module + the few attributes the reference touches."""

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            import torch

            return torch.device("cpu")

    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass


def install(reference_root: str = REFERENCE_ROOT) -> str:
    """Put the reference on sys.path behind inert stubs; returns the root."""
    if not os.path.isdir(reference_root):
        raise FileNotFoundError(
            f"reference tree not found at {reference_root}; golden generation only runs "
            "in the authoring container")
    for name in [
        "pytorch_lightning", "pytorch_lightning.utilities", "pytorch_lightning.callbacks",
        "pytorch_lightning.loggers", "pytorch_lightning.utilities.distributed",
        "lightning", "lightning.pytorch", "lightning.pytorch.utilities",
        "omegaconf", "omegaconf.listconfig",
        "monai", "monai.data", "monai.transforms", "monai.metrics", "monai.utils", "monai.losses",
        "monai.networks", "monai.networks.nets", "monai.networks.blocks", "monai.inferers",
        "monai.config", "monai.networks.layers", "monai.networks.layers.utils",
        "SimpleITK", "diffusers", "diffusers.models", "h5py", "deepspeed", "deepspeed.ops",
        "deepspeed.ops.adam", "matplotlib", "matplotlib.pyplot", "skimage", "skimage.metrics",
        "lpips", "torchmetrics", "torchmetrics.image", "kornia", "open_clip", "xformers_absent",
        "timm", "timm.models", "timm.models.vision_transformer",
    ]:
        _stub(name)
    pl = sys.modules["pytorch_lightning"]
    pl.LightningModule = _LightningModule
    sys.modules["lightning"].LightningModule = _LightningModule
    sys.modules["lightning.pytorch"].LightningModule = _LightningModule
    sys.modules["pytorch_lightning.utilities"].rank_zero_only = lambda f: f
    sys.modules["pytorch_lightning.utilities.distributed"].rank_zero_only = lambda f: f
    sys.modules["lightning.pytorch.utilities"].rank_zero_only = lambda f: f
    sys.modules["omegaconf"].ListConfig = list
    sys.modules["omegaconf.listconfig"].ListConfig = list
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    return reference_root


class AttrDict(dict):
    """yaml dict with attribute access and .pop — what DiffusionWrapper expects of a config."""

    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(obj):
    if isinstance(obj, dict):
        return AttrDict({k: to_attr(v) for k, v in obj.items()})
    if isinstance(obj, list):
        return [to_attr(v) for v in obj]
    return obj
