"""Predict-step glue on the OUTPUT side of the sampling path: finished 2D slices -> 3D volumes.

Mirrors what the reference's Lightning hooks do around the sampler (trainers/trainer_ds_diff.py:826-875):

* ``predict_step``           parses ``id_num`` / ``slice_idx`` out of each slice path (:828-829)   -> ``parse_slice_path``
* ``on_predict_batch_end``   one ``img.cpu().numpy()`` PER SLICE into ``pred_dict[id][slice]`` (:847-852) -> ``PredictionAssembler.add_batch``
                             (here: ONE device->pinned-host copy per batch on the sampling stream, no per-slice sync)
* ``on_predict_end``         ``pred_array = zeros_like(template)``; ``pred_array[slice] = img``; copy the template's geometry;
                             write ``{id}_pred.nii.gz`` (:854-875)                                 -> ``PredictionAssembler.finalize``

The reference goes through SimpleITK, which is not part of this image; the NIfTI-1 container is restated here from the
format's public definition (348-byte header, x-fastest voxel order == the (z, y, x) array ``sitk.GetArrayFromImage``
returns).  Host-side byte shuffling only: no arithmetic of the sampling path lives in this file.
"""
from __future__ import annotations

import gzip
import os
import struct
from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor

_HDR = 348
_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
           768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


def parse_slice_path(path: str) -> Tuple[str, int]:
    """``.../<id_num>/<name>_<slice>.h5`` -> (id_num, slice): trainer_ds_diff.py:828-829."""
    parts = path.split("/")
    if len(parts) < 2:
        raise ValueError(f"slice path has no patient directory: {path!r}")
    return parts[-2], int(os.path.basename(path).split(".")[0].split("_")[-1])


class NiftiHeader:
    """The 348 raw header bytes of a NIfTI-1 file plus the few fields the assembly needs."""

    def __init__(self, raw: bytes):
        if len(raw) < _HDR:
            raise ValueError("NIfTI header is shorter than 348 bytes")
        self.raw = bytes(raw[:_HDR])
        for order in ("<", ">"):
            if struct.unpack_from(order + "i", self.raw, 0)[0] == _HDR:
                self.order = order
                break
        else:
            raise ValueError("not a NIfTI-1 header (sizeof_hdr != 348)")
        if self.raw[344:347] not in (b"n+1", b"ni1"):
            raise ValueError("not a NIfTI-1 header (bad magic)")
        u = lambda fmt, off: struct.unpack_from(self.order + fmt, self.raw, off)
        self.dim = u("8h", 40)
        self.datatype = u("h", 70)[0]
        self.pixdim = u("8f", 76)
        self.vox_offset = int(u("f", 108)[0])
        self.scl_slope, self.scl_inter = u("2f", 112)
        if not 1 <= self.dim[0] <= 7:
            raise ValueError(f"bad NIfTI dim[0] = {self.dim[0]}")
        if self.datatype not in _DTYPES:
            raise ValueError(f"unsupported NIfTI datatype code {self.datatype}")

    @property
    def shape(self) -> Tuple[int, ...]:
        """Array shape in the (z, y, x) order SimpleITK's GetArrayFromImage uses (x is the fastest axis on disk)."""
        return tuple(reversed(self.dim[1:1 + self.dim[0]]))

    @property
    def dtype(self) -> np.dtype:
        return np.dtype(_DTYPES[self.datatype]).newbyteorder(self.order)

    def with_dtype(self, dtype) -> bytes:
        """Header bytes for an array of ``dtype`` on this geometry (what CopyInformation + WriteImage produce): same
        dim / pixdim / qform / sform, data type of the new array, no intensity rescale, single-file layout."""
        dt = np.dtype(dtype).newbyteorder("=")
        if dt not in _CODES:
            raise ValueError(f"dtype {dt} has no NIfTI-1 code")
        b = bytearray(self.raw)
        struct.pack_into(self.order + "2h", b, 70, _CODES[dt], dt.itemsize * 8)
        struct.pack_into(self.order + "f", b, 108, 352.0)
        struct.pack_into(self.order + "2f", b, 112, 1.0, 0.0)
        b[344:348] = b"n+1\0"
        return bytes(b)


def _open(path: str, mode: str, gz: Optional[bool] = None):
    gz = path.endswith(".gz") if gz is None else gz
    return gzip.open(path, mode) if gz else open(path, mode)


def read_nifti(path: str) -> Tuple[np.ndarray, NiftiHeader]:
    """-> (array in (z, y, x) order, header).  A header with an intensity rescale (scl_slope not in {0, 1} or
    scl_inter != 0) yields the rescaled float32 array, as ITK's reader does."""
    with _open(path, "rb") as f:
        blob = f.read()
    hdr = NiftiHeader(blob)
    n = int(np.prod(hdr.shape))
    off = max(hdr.vox_offset, 352) if hdr.raw[344:347] == b"n+1" else None
    if off is None:
        raise ValueError("two-file NIfTI (.hdr/.img) is not supported")
    if len(blob) < off + n * hdr.dtype.itemsize:
        raise ValueError(f"{path}: truncated voxel data")
    arr = np.frombuffer(blob, dtype=hdr.dtype, count=n, offset=off).reshape(hdr.shape)
    arr = arr.astype(hdr.dtype.newbyteorder("="))
    if hdr.scl_slope not in (0.0, 1.0) or (hdr.scl_slope != 0.0 and hdr.scl_inter != 0.0):
        arr = (arr.astype(np.float64) * hdr.scl_slope + hdr.scl_inter).astype(np.float32)
    return arr, hdr


def write_nifti_like(path: str, array: np.ndarray, template: NiftiHeader) -> None:
    """Write ``array`` ((z, y, x) order) with the geometry of ``template`` (origin, spacing, direction: the header's
    pixdim / qform / sform bytes are carried over untouched)."""
    if tuple(array.shape) != template.shape:
        raise ValueError(f"array shape {tuple(array.shape)} does not match the template volume {template.shape}")
    data = np.ascontiguousarray(array, dtype=np.dtype(array.dtype).newbyteorder(template.order))
    tmp = path + ".part"
    with _open(tmp, "wb", gz=path.endswith(".gz")) as f:
        f.write(template.with_dtype(array.dtype))
        f.write(b"\0\0\0\0")                       # no header extensions; voxels start at byte 352
        f.write(data.tobytes())
    os.replace(tmp, path)                           # a crashed run never leaves a half-written volume behind


def assemble_volume(slices: Dict[int, np.ndarray], template_array: np.ndarray) -> np.ndarray:
    """``pred_array = zeros_like(template); pred_array[slice_idx] = img`` (trainer_ds_diff.py:864-866): slices that were
    never predicted stay zero, and values are cast to the TEMPLATE's dtype by the assignment."""
    pred = np.zeros_like(template_array)
    for idx, img in slices.items():
        i = int(idx)
        if not 0 <= i < pred.shape[0]:
            raise IndexError(f"slice {i} outside the template volume ({pred.shape[0]} slices)")
        img = np.asarray(img)
        while img.ndim > pred.ndim - 1 and img.shape[0] == 1:       # [1, H, W] network output -> [H, W]
            img = img[0]
        pred[i] = img
    return pred


class PredictionAssembler:
    """``pred_dict`` of the reference's predict hooks (trainer_ds_diff.py:821-875) with batched device->host traffic.

    ``add_batch`` enqueues ONE copy of the whole batch into a pinned staging buffer on the current stream and returns
    without synchronising (the reference does ``img.cpu().numpy()`` per slice: B blocking copies per batch);
    ``finalize`` synchronises once, then assembles and writes one ``{id}_pred.nii.gz`` per patient."""

    def __init__(self, ids: Optional[Iterable[str]] = None):
        self.pred: Dict[str, Dict[int, np.ndarray]] = {i: {} for i in (ids or [])}
        self._pending = []       # (host tensor, event, ids, slice indices)

    def add_batch(self, ids: Sequence[str], slice_idx: Sequence[int], samples: Tensor) -> None:
        if not (len(ids) == len(slice_idx) == samples.shape[0]):
            raise ValueError("add_batch: ids, slice_idx and samples disagree on the batch size")
        if samples.is_cuda:
            host = torch.empty(samples.shape, dtype=samples.dtype, pin_memory=True)
            host.copy_(samples, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        else:
            host, ev = samples.detach().clone(), None
        self._pending.append((host, ev, list(ids), [int(s) for s in slice_idx]))

    def add_paths(self, paths: Sequence[str], samples: Tensor) -> None:
        parsed = [parse_slice_path(p) for p in paths]
        self.add_batch([p[0] for p in parsed], [p[1] for p in parsed], samples)

    def _drain(self) -> None:
        for host, ev, ids, idx in self._pending:
            if ev is not None:
                ev.synchronize()
            arr = host.numpy()
            for j, (pid, s) in enumerate(zip(ids, idx)):
                self.pred.setdefault(pid, {})[s] = arr[j]
        self._pending = []

    def volumes(self, templates: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
        self._drain()
        return {pid: assemble_volume(sl, templates[pid]) for pid, sl in self.pred.items()}

    def finalize(self, template_path: str, out_dir: str, template_name: str = "T1CE.nii.gz") -> Dict[str, str]:
        """For every patient: read ``<template_path>/<id>/<template_name>``, assemble, write ``<out_dir>/<id>_pred.nii.gz``
        with the template's geometry.  Returns {id: written path}."""
        self._drain()
        os.makedirs(out_dir, exist_ok=True)
        written = {}
        for pid, sl in self.pred.items():
            tarr, thdr = read_nifti(os.path.join(template_path, pid, template_name))
            out = os.path.join(out_dir, f"{pid}_pred.nii.gz")
            write_nifti_like(out, assemble_volume(sl, tarr), thdr)
            written[pid] = out
        return written
