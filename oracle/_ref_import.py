"""Import shim for the UNMODIFIED reference tree (test infrastructure only).

Used by ``oracle/make_golden*.py`` (run in the authoring container, where
``/root/reference`` exists) to pin the oracle restatement against the real
reference code, and — through the byte-compiled copy ``oracle/_ref/`` that
``oracle/build_ref.py`` produces and ``gpurun`` ships — by ``bench.py``'s
reference arm / ``gpu_eager_reference`` leg and the reference-deviation tests on
the GPU box, where ``/root/reference`` does not exist.  Nothing in the product
package imports this file.

The reference imports a number of third-party packages that are absent from
this image (pytorch_lightning, lightning, omegaconf, monai, SimpleITK,
diffusers, timm, ...).  None of them does arithmetic on the sampling hot path,
so they are replaced with inert ``sys.modules`` stubs.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types

import torch.nn as nn

COMPILED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _default_root() -> str:
    env = os.environ.get("DSDIFF_REFERENCE_ROOT")
    if env:
        return env
    return "/root/reference" if os.path.isdir("/root/reference") else COMPILED_ROOT


REFERENCE_ROOT = _default_root()


def available() -> bool:
    """True when the reference (sources here, or the byte-compiled oracle/_ref on the GPU box) can be imported."""
    root = _default_root()
    return any(os.path.isfile(os.path.join(root, "UNet_DS_Diff", "model" + ext)) for ext in (".py", ".refc"))


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


# third-party packages the reference imports that are absent from the image; none does arithmetic on the hot path
ABSENT_TOP = {"pytorch_lightning", "lightning", "omegaconf", "monai", "SimpleITK", "diffusers", "h5py", "deepspeed",
              "matplotlib", "skimage", "lpips", "torchmetrics", "kornia", "open_clip", "timm", "openpyxl", "nibabel",
              "seaborn", "thop", "fvcore", "clip", "taming", "pytorch_msssim", "medpy", "natsort", "wandb", "tensorboard",
              "torchsummary", "albumentations", "imageio", "ants", "torchio", "einops_exts", "mpi4py", "blobfile"}


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Last-resort finder: any (sub)module of an ABSENT_TOP package that no real finder provides becomes an inert stub."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in ABSENT_TOP:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        parent, _, child = module.__name__.rpartition(".")
        if parent and parent in sys.modules:
            setattr(sys.modules[parent], child, module)


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


def install(reference_root: str = None) -> str:
    """Put the reference on sys.path behind inert stubs; returns the root (/root/reference in the authoring container,
    else the byte-compiled oracle/_ref)."""
    reference_root = reference_root or _default_root()
    if not any(os.path.isfile(os.path.join(reference_root, "UNet_DS_Diff", "model" + ext)) for ext in (".py", ".refc")):
        raise FileNotFoundError(
            f"reference tree not found at {reference_root}: run oracle/build_ref.py in the authoring container "
            "(it byte-compiles /root/reference into oracle/_ref, which gpurun ships to the GPU box)")
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())
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
    if os.path.isfile(os.path.join(reference_root, "BUILD_INFO.txt")):     # compiled tree: .refc modules
        _register_compiled_root(reference_root)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    return reference_root


_COMPILED_EXT = ".refc"      # oracle/build_ref.py


def _compiled_path_hook(path: str):
    """sys.path_hooks entry: directories under a compiled reference tree are searched for ``.refc`` modules (the .pyc
    format under an extension that snapshot tools do not treat as cache files)."""
    if not os.path.isdir(path) or _compiled_root_of(path) is None:
        raise ImportError("not a compiled reference directory")
    return importlib.machinery.FileFinder(path, (importlib.machinery.SourcelessFileLoader, [_COMPILED_EXT]))


def _compiled_root_of(path: str):
    path = os.path.abspath(path)
    for root in _compiled_roots:
        if path == root or path.startswith(root + os.sep):
            return root
    return None


_compiled_roots = []


def _register_compiled_root(root: str) -> None:
    root = os.path.abspath(root)
    if root not in _compiled_roots:
        _compiled_roots.append(root)
    if _compiled_path_hook not in sys.path_hooks:
        sys.path_hooks.insert(0, _compiled_path_hook)
    sys.path_importer_cache.clear()


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
