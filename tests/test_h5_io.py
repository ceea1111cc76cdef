"""HDF5 slice-file reader / batch builder (input half): the container format is restated from the HDF5
specification (no libhdf5 / h5py in the image: parity with libhdf5 is unpinned), so the reader is checked against
structures packed BY HAND here, byte by byte, and against the writer's round trip."""
import struct
import zlib

import numpy as np
import pytest
import torch

from diffusion_models_dsdiff_b200 import h5_io as H

U = H.UNDEF


def _slice_arrays(seed=0, hw=(12, 10)):
    r = np.random.default_rng(seed)
    return {"F_Data1": r.standard_normal(hw).astype(np.float32), "F_Data2": r.standard_normal(hw).astype(np.float64),
            "S_Data1": r.integers(-1000, 1000, hw).astype(np.int16), "S_Data2": r.integers(0, 60000, hw).astype(np.uint16)}


def test_round_trip_all_dtypes(tmp_path):
    arrs = _slice_arrays()
    arrs["big"] = np.arange(24, dtype=">f4").reshape(2, 3, 4)
    p = tmp_path / "layer_0.h5"
    H.write_h5(p, arrs)
    f = H.H5File(p)
    assert f.keys() == sorted(arrs)
    for k, v in arrs.items():
        got = f[k]
        assert got.shape == v.shape and got.dtype == v.dtype and np.array_equal(got, v), k
    assert "F_Data1" in f and "nope" not in f
    with pytest.raises(KeyError):
        f["nope"]


def test_written_file_has_the_documented_structure(tmp_path):
    """Byte-level checks of what write_h5 emits against the HDF5 specification's field offsets."""
    p = tmp_path / "a.h5"
    H.write_h5(p, {"x": np.arange(6, dtype="<f4").reshape(2, 3)})
    b = p.read_bytes()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0                      # signature, superblock version 0
    assert b[13] == 8 and b[14] == 8                                        # sizes of offsets / lengths
    assert struct.unpack_from("<HH", b, 16) == (4, 16)                      # group leaf / internal node K
    base, _free, eof, _drv = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and eof == len(b)
    _noff, root, cache = struct.unpack_from("<QQI", b, 56)
    assert cache == 1 and b[root] == 1                                      # cached symbol table, v1 object header
    btree, heap = struct.unpack_from("<QQ", b, 56 + 24)
    assert b[btree:btree + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP"
    mtype, msize = struct.unpack_from("<HH", b, root + 16)
    assert mtype == 0x11 and struct.unpack_from("<QQ", b, root + 24) == (btree, heap)
    snod = struct.unpack_from("<Q", b, btree + 24 + 8)[0]
    assert b[snod:snod + 4] == b"SNOD" and struct.unpack_from("<H", b, snod + 6)[0] == 1
    # the dataset's raw data: six little-endian floats, contiguous, 8-byte aligned
    raw = np.arange(6, dtype="<f4").tobytes()
    assert b.count(raw) == 1 and b.index(raw) % 8 == 0


def _hand_packed_chunked(gzip: bool, shuffle: bool):
    """A file assembled by hand: superblock 1, root group, ONE chunked int32 dataset 5 x 6 with 4 x 4 chunks stored through a
    two-level version-1 chunk B-tree, optional shuffle + deflate filters, object header with a continuation block."""
    data = np.arange(30, dtype="<i4").reshape(5, 6)
    chunks = {}
    for r0 in (0, 4):
        for c0 in (0, 4):
            blk = np.zeros((4, 4), "<i4")
            sub = data[r0:r0 + 4, c0:c0 + 4]
            blk[:sub.shape[0], :sub.shape[1]] = sub
            raw = blk.tobytes()
            if shuffle:
                raw = np.frombuffer(raw, np.uint8).reshape(-1, 4).T.tobytes()
            if gzip:
                raw = zlib.compress(raw, 6)
            chunks[(r0, c0)] = raw
    out = bytearray(4096)
    pos = [1024]

    def put(blob, align=8):
        pos[0] += -pos[0] % align
        a = pos[0]
        out[a:a + len(blob)] = blob
        pos[0] += len(blob)
        return a

    caddr = {k: put(v) for k, v in chunks.items()}

    def key(size, r, c):
        return struct.pack("<IIQQQ", size, 0, r, c, 0)

    def leaf(items):
        node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(items), U, U)
        for (r, c) in items:
            node += key(len(chunks[(r, c)]), r, c) + struct.pack("<Q", caddr[(r, c)])
        return put(node + key(0, 8, 8))

    l0, l1 = leaf([(0, 0), (0, 4)]), leaf([(4, 0), (4, 4)])
    top = put(b"TREE" + struct.pack("<BBHQQ", 1, 1, 2, U, U) + key(0, 0, 0) + struct.pack("<Q", l0) + key(0, 4, 0)
              + struct.pack("<Q", l1) + key(0, 8, 8))
    space = struct.pack("<BBB5xQQ", 1, 2, 0, 5, 6)
    dtype = struct.pack("<BBBBIHH", 0x10, 0x08, 0, 0, 4, 0, 32)
    layout = struct.pack("<BBBQIII", 3, 2, 3, top, 4, 4, 4)
    filt = b""
    nf = int(shuffle) + int(gzip)
    if nf:
        filt = struct.pack("<BB6x", 1, nf)
        if shuffle:
            filt += struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<I4x", 4)
        if gzip:
            filt += struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<I4x", 6)
    cont_msgs = H._msg(0x08, layout) + (H._msg(0x0B, filt) if nf else b"")
    cont = put(cont_msgs)
    first = H._msg(0x01, space) + H._msg(0x03, dtype, 1) + H._msg(0x10, struct.pack("<QQ", cont, len(cont_msgs)))
    dset = put(struct.pack("<BxHII4x", 1, 3 + (2 if nf else 1), 1, len(first)) + first)
    heap_data = put(b"\x00" * 8 + b"vol\x00" + bytes(4) + bytes(16))
    heap = put(b"HEAP" + struct.pack("<B3xQQQ", 0, 32, U, heap_data))
    snod = put(b"SNOD" + struct.pack("<BxH", 1, 1) + struct.pack("<QQII16x", 8, dset, 0, 0))
    gtree = put(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, U, U) + struct.pack("<QQQ", 0, snod, 8))
    root = put(H._object_header([H._msg(0x11, struct.pack("<QQ", gtree, heap))]))
    sb = H.SIGNATURE + struct.pack("<BBBBBBBBHHI", 1, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0) + struct.pack("<HH", 32, 0)
    sb += struct.pack("<QQQQ", 0, U, len(out), U) + struct.pack("<QQII16x", 0, root, 0, 0)
    out[0:len(sb)] = sb
    return bytes(out), data


@pytest.mark.parametrize("gzip,shuffle", [(False, False), (True, False), (True, True)])
def test_hand_packed_chunked_dataset(gzip, shuffle):
    blob, data = _hand_packed_chunked(gzip, shuffle)
    f = H.H5File(blob)
    assert f.keys() == ["vol"]
    got = f["vol"]
    assert got.dtype == np.dtype("<i4") and np.array_equal(got, data)


def test_hand_packed_v2_superblock_compact_group_and_compact_dataset():
    """libver='latest' flavour: superblock 2, OHDR object headers, a link message, a compact (in-header) dataset."""
    vals = np.array([1.5, -2.25, 3.0], "<f8")
    out = bytearray(1024)

    def ohdr(msgs):
        body = b"".join(struct.pack("<BHB", t, len(m), 0) + m for t, m in msgs)
        return b"OHDR" + struct.pack("<BBB", 2, 0, len(body)) + body + b"\x00\x00\x00\x00"

    space = struct.pack("<BBBBQ", 2, 1, 0, 1, 3)
    dtype = struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
    layout = struct.pack("<BBH", 3, 0, vals.nbytes) + vals.tobytes()
    d = ohdr([(0x01, space), (0x03, dtype), (0x08, layout)])
    a_d = 256
    out[a_d:a_d + len(d)] = d
    link = struct.pack("<BBB", 1, 0, 3) + b"pet" + struct.pack("<Q", a_d)
    g = ohdr([(0x02, struct.pack("<BBQQ", 0, 0, U, U)), (0x06, link)])
    a_g = 128
    out[a_g:a_g + len(g)] = g
    sb = H.SIGNATURE + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<QQQQ", 0, U, len(out), a_g) + b"\x00" * 4
    out[:len(sb)] = sb
    f = H.H5File(bytes(out))
    assert f.keys() == ["pet"] and np.array_equal(f["pet"], vals)


def test_rejects_garbage():
    with pytest.raises(H.H5FormatError):
        H.H5File(b"not an hdf5 file" * 10)


def test_slice_loader_builds_padded_pinned_batches(tmp_path):
    """LoadH5 + collate (my_transform.py:142-153): keys stacked along C, 240 -> 256 centre padding, ragged last batch."""
    paths, want = [], []
    for i in range(5):
        arrs = _slice_arrays(seed=i, hw=(24, 20))
        p = tmp_path / f"layer_{i}.h5"
        H.write_h5(p, arrs)
        paths.append(str(p))
        want.append(np.stack([arrs[k].astype(np.float32) for k in ("F_Data1", "S_Data1", "S_Data2")]))
    loader = H.H5SliceLoader(paths, ("F_Data1", "S_Data1", "S_Data2"), batch_size=2, pad_to=(32, 32), pin=False)
    assert len(loader) == 3
    seen = 0
    for names, batch in loader:
        assert batch.dtype == torch.float32 and tuple(batch.shape[1:]) == (3, 32, 32) and batch.shape[0] == len(names)
        for j in range(len(names)):
            assert np.array_equal(batch[j, :, 4:28, 6:26].numpy(), want[seen + j])
            assert float(batch[j, :, :4].abs().sum()) == 0.0 and float(batch[j, :, :, 26:].abs().sum()) == 0.0
        seen += len(names)
    assert seen == 5
