"""Predict-step output glue: slice-path parsing, 2D -> 3D assembly, NIfTI-1 container round trips.
CPU only; mirrors trainers/trainer_ds_diff.py:826-875 (the reference's SimpleITK is not in this image, so the container
is checked against hand-packed NIfTI-1 bytes and its own round trip)."""
import gzip
import os
import struct

import numpy as np
import pytest
import torch

from diffusion_models_dsdiff_b200 import volume_io as V


def _header(shape_zyx, code, bitpix, order="<", slope=0.0, inter=0.0, pixdim=(1.0, 0.9, 0.8, 2.5)):
    b = bytearray(348)
    struct.pack_into(order + "i", b, 0, 348)
    dims = [len(shape_zyx)] + list(reversed(shape_zyx)) + [1] * (7 - len(shape_zyx))
    struct.pack_into(order + "8h", b, 40, *dims)
    struct.pack_into(order + "2h", b, 70, code, bitpix)
    struct.pack_into(order + "8f", b, 76, *(list(pixdim) + [0.0] * (8 - len(pixdim))))
    struct.pack_into(order + "f", b, 108, 352.0)
    struct.pack_into(order + "2f", b, 112, slope, inter)
    struct.pack_into(order + "h", b, 252, 1)                       # qform_code
    struct.pack_into(order + "3f", b, 268, -90.0, 126.0, -72.0)    # qoffset
    b[344:348] = b"n+1\0"
    return bytes(b)


def _write_template(path, arr, order="<", **kw):
    code = {np.dtype("int16"): 4, np.dtype("float32"): 16, np.dtype("uint8"): 2}[arr.dtype]
    blob = _header(arr.shape, code, arr.dtype.itemsize * 8, order, **kw) + b"\0" * 4 + \
        arr.astype(arr.dtype.newbyteorder(order)).tobytes()
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with (gzip.open(path, "wb") if path.endswith(".gz") else open(path, "wb")) as f:
        f.write(blob)


def test_parse_slice_path_follows_reference_convention():
    assert V.parse_slice_path("/data/test/BraTS_0042/BraTS_0042_77.h5") == ("BraTS_0042", 77)
    assert V.parse_slice_path("p/id7/slice_003.nii.gz") == ("id7", 3)
    with pytest.raises(ValueError):
        V.parse_slice_path("nofolder_3.h5")
    with pytest.raises(ValueError):
        V.parse_slice_path("a/b/noindex.h5")


@pytest.mark.parametrize("order", ["<", ">"])
@pytest.mark.parametrize("gz", [True, False])
def test_read_nifti_layout_and_byte_order(tmp_path, order, gz):
    arr = (np.arange(4 * 5 * 6, dtype=np.int16) - 30).reshape(4, 5, 6)         # (z, y, x): x fastest on disk
    p = str(tmp_path / "id" / ("T1CE.nii.gz" if gz else "T1CE.nii"))
    _write_template(p, arr, order)
    got, hdr = V.read_nifti(p)
    assert hdr.shape == (4, 5, 6) and got.dtype == np.int16 and hdr.order == order
    assert np.array_equal(got, arr)
    assert hdr.pixdim[1:4] == pytest.approx((0.9, 0.8, 2.5))


def test_read_nifti_applies_intensity_rescale_like_itk(tmp_path):
    arr = np.arange(24, dtype=np.int16).reshape(2, 3, 4)
    p = str(tmp_path / "a" / "x.nii.gz")
    _write_template(p, arr, slope=0.5, inter=-3.0)
    got, _ = V.read_nifti(p)
    assert got.dtype == np.float32 and np.array_equal(got, arr.astype(np.float32) * 0.5 - 3.0)
    _write_template(p, arr, slope=1.0, inter=0.0)
    assert V.read_nifti(p)[0].dtype == np.int16
    _write_template(p, arr, slope=0.0, inter=9.0)                               # slope 0: "no scaling", inter ignored
    assert V.read_nifti(p)[0].dtype == np.int16


def test_bad_files_are_rejected(tmp_path):
    p = str(tmp_path / "bad.nii")
    open(p, "wb").write(b"\0" * 400)
    with pytest.raises(ValueError):
        V.read_nifti(p)
    arr = np.zeros((2, 3, 4), np.float32)
    q = str(tmp_path / "t" / "trunc.nii")
    _write_template(q, arr)
    blob = open(q, "rb").read()
    open(q, "wb").write(blob[:-8])
    with pytest.raises(ValueError, match="truncated"):
        V.read_nifti(q)
    _, hdr = (lambda: (_write_template(q, arr), V.read_nifti(q))[1])()
    with pytest.raises(ValueError, match="shape"):
        V.write_nifti_like(str(tmp_path / "o.nii.gz"), np.zeros((2, 3, 5), np.float32), hdr)


def test_write_keeps_template_geometry_and_round_trips(tmp_path):
    tarr = np.zeros((3, 8, 8), np.int16)
    p = str(tmp_path / "id" / "T1CE.nii.gz")
    _write_template(p, tarr, slope=2.0, inter=1.0)
    _, hdr = V.read_nifti(p)
    pred = np.random.default_rng(0).standard_normal((3, 8, 8)).astype(np.float32)
    out = str(tmp_path / "pred.nii.gz")
    V.write_nifti_like(out, pred, hdr)
    got, h2 = V.read_nifti(out)
    assert got.dtype == np.float32 and np.array_equal(got, pred)
    assert h2.raw[76:108] == hdr.raw[76:108] and h2.raw[252:344] == hdr.raw[252:344]     # pixdim, qform, sform
    assert (h2.scl_slope, h2.scl_inter) == (1.0, 0.0) and h2.datatype == 16 and h2.vox_offset == 352
    assert not os.path.exists(out + ".part")


def test_assemble_volume_matches_reference_semantics():
    """pred_array = zeros_like(template); pred_array[i] = img: template dtype wins, missing slices stay zero."""
    t = np.ones((5, 4, 4), np.int16)
    sl = {3: np.full((1, 4, 4), 7.9, np.float32), "1": np.full((4, 4), -2.2, np.float32)}
    got = V.assemble_volume(sl, t)
    ref = np.zeros_like(t)
    ref[3] = sl[3]
    ref[1] = sl["1"]
    assert got.dtype == np.int16 and np.array_equal(got, ref)
    assert got[0].max() == 0 and got[3, 0, 0] == 7 and got[1, 0, 0] == -2
    with pytest.raises(IndexError):
        V.assemble_volume({5: np.zeros((4, 4))}, t)


def test_prediction_assembler_end_to_end(tmp_path):
    rng = np.random.default_rng(1)
    tdir, odir = str(tmp_path / "templates"), str(tmp_path / "out")
    for pid, nz in (("P01", 4), ("P02", 3)):
        _write_template(os.path.join(tdir, pid, "T1CE.nii.gz"), rng.standard_normal((nz, 16, 16)).astype(np.float32))
    asm = V.PredictionAssembler(["P01", "P02"])
    b1 = torch.randn(3, 1, 16, 16)
    b2 = torch.randn(3, 1, 16, 16)
    asm.add_paths(["/d/P01/P01_0.h5", "/d/P01/P01_2.h5", "/d/P02/P02_1.h5"], b1)
    asm.add_batch(["P02", "P01", "P02"], [0, 3, 2], b2)
    with pytest.raises(ValueError):
        asm.add_batch(["P01"], [0, 1], b1)
    written = asm.finalize(tdir, odir)
    assert sorted(written) == ["P01", "P02"] and written["P01"].endswith("P01_pred.nii.gz")
    v1, _ = V.read_nifti(written["P01"])
    v2, _ = V.read_nifti(written["P02"])
    assert np.array_equal(v1[0], b1[0, 0].numpy()) and np.array_equal(v1[2], b1[1, 0].numpy())
    assert np.array_equal(v1[3], b2[1, 0].numpy()) and not v1[1].any()                 # slice 1 never predicted
    assert np.array_equal(v2[1], b1[2, 0].numpy()) and np.array_equal(v2[0], b2[0, 0].numpy())
    assert np.array_equal(v2[2], b2[2, 0].numpy())


@pytest.mark.gpu
def test_prediction_assembler_takes_device_batches_without_per_slice_sync():
    asm = V.PredictionAssembler()
    batches = [torch.randn(4, 1, 32, 32, device="cuda") for _ in range(3)]
    for k, b in enumerate(batches):
        asm.add_batch(["A"] * 4, range(4 * k, 4 * k + 4), b)
    vols = asm.volumes({"A": np.zeros((12, 32, 32), np.float32)})
    assert np.array_equal(vols["A"], torch.cat(batches)[:, 0].cpu().numpy())
