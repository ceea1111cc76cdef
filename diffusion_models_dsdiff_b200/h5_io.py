"""Input half of the predict glue: HDF5 slice files -> pinned host batches.

The reference stores every 2-D slice as ``layer_{i}.h5`` with one dataset per contrast, written with
``f[key] = array`` (preprocess/to_h5.py:40-50) and read back per sample with ``h5py.File(path)[key][()]``
(``LoadH5``, training_project/utils/my_transform.py:142-153) inside a MONAI dataset; the batches then reach the sampler
through ``get_input`` (trainers/trainer_ds_diff.py:543-557).  ``h5py`` / libhdf5 are not in this image, so the container
format is restated here from the HDF5 File Format Specification (version 3.0) for exactly what those files hold:

* superblock version 0 / 1 (h5py's default ``libver='earliest'``) and 2 / 3 (``libver='latest'``);
* "old style" groups: symbol-table message -> version-1 B-tree -> ``SNOD`` symbol nodes + local heap, and
  "new style" compact groups (link messages in a version-2 object header; dense groups raise);
* object headers version 1 (with continuation blocks) and version 2 (``OHDR`` / ``OCHK``);
* dataspace (simple), datatype (fixed-point and IEEE floating-point, either byte order), data layout version 1-3
  (compact, contiguous, chunked through the version-1 chunk B-tree) and layout version 4 compact / contiguous /
  single-chunk; filters deflate (zlib), shuffle and fletcher32.

Host byte shuffling only — no arithmetic.  Parity with libhdf5 is UNPINNED (no HDF5 library exists here to write or read
a fixture): ``tests/test_h5_io.py`` checks the reader against hand-packed structures and the round trip through
``write_h5`` below, which emits the layout h5py produces by default (superblock 0, symbol-table root group, contiguous
little-endian datasets).

``H5SliceLoader`` is the batch builder: it loads the keys of ``batch_size`` slice files straight into ONE pinned
``[B, C, H, W]`` fp32 buffer per batch (the reference builds per-sample tensors, collates them, and copies the collated
batch from pageable memory), ready for a single asynchronous host->device copy.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


# ---------------------------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------------------------
class H5File:
    """Read-only view of an HDF5 file: ``H5File(path)[key]`` -> numpy array (``h5py.File(path)[key][()]``).
    Nested groups are addressed with ``"a/b"`` paths; ``keys()`` lists the root group."""

    def __init__(self, path_or_bytes):
        if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
            self.buf = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, "rb") as fh:
                self.buf = fh.read()
        self._parse_superblock()
        self._root = self._group_links(self.root_addr)

    # -- low level -----------------------------------------------------------------------------
    def _u(self, off: int, size: int) -> int:
        if off + size > len(self.buf):
            raise H5FormatError("read past the end of the file")
        return int.from_bytes(self.buf[off:off + size], "little")

    def _addr(self, off: int) -> int:
        return self._u(off, self.so)

    def _len(self, off: int) -> int:
        return self._u(off, self.sl)

    def _parse_superblock(self) -> None:
        base = -1
        off = 0
        while off < len(self.buf):          # the superblock may sit at 0, 512, 1024, ... (user block)
            if self.buf[off:off + 8] == SIGNATURE:
                base = off
                break
            off = 512 if off == 0 else off * 2
        if base < 0:
            raise H5FormatError("not an HDF5 file (signature not found)")
        ver = self.buf[base + 8]
        self.sb_version = ver
        if ver in (0, 1):
            self.so, self.sl = self.buf[base + 13], self.buf[base + 14]
            p = base + 24 + (4 if ver == 1 else 0)
            self.base_addr = self._addr(p)
            p += 4 * self.so                 # base, free-space, end-of-file, driver-info addresses
            # root group symbol table entry: name offset, object header address, cache type, reserved, scratch
            self.root_addr = self._addr(p + self.so)
        elif ver in (2, 3):
            self.so, self.sl = self.buf[base + 9], self.buf[base + 10]
            p = base + 12
            self.base_addr = self._addr(p)
            self.root_addr = self._addr(p + 3 * self.so)
        else:
            raise H5FormatError(f"unsupported superblock version {ver}")
        if self.so not in (4, 8) or self.sl not in (4, 8):
            raise H5FormatError("unsupported offset / length size")
        self.base_addr = self.base_addr + 0 if self.base_addr != self._undef() else 0

    def _undef(self) -> int:
        return (1 << (8 * self.so)) - 1

    # -- object headers ------------------------------------------------------------------------
    def _messages(self, addr: int) -> List[Tuple[int, int, int]]:
        """(type, data offset, data size) of every header message of the object at ``addr``."""
        addr += self.base_addr
        out: List[Tuple[int, int, int]] = []
        if self.buf[addr:addr + 4] == b"OHDR":
            if self.buf[addr + 4] != 2:
                raise H5FormatError("bad version-2 object header")
            flags = self.buf[addr + 5]
            p = addr + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            csz = 1 << (flags & 3)
            chunk0 = self._u(p, csz)
            p += csz
            blocks = [(p, p + chunk0)]
            track_order = bool(flags & 0x04)
            while blocks:
                p, end = blocks.pop(0)
                while p + 4 <= end:
                    mtype = self.buf[p]
                    msize = self._u(p + 1, 2)
                    p += 4 + (2 if track_order else 0)
                    if p + msize > end:
                        break
                    if mtype == 0x10:          # continuation: OCHK block (signature + messages + checksum)
                        caddr, clen = self._addr(p) + self.base_addr, self._len(p + self.so)
                        if self.buf[caddr:caddr + 4] != b"OCHK":
                            raise H5FormatError("bad object header continuation block")
                        blocks.append((caddr + 4, caddr + clen - 4))
                    elif mtype != 0:
                        out.append((mtype, p, msize))
                    p += msize
            return out
        if self.buf[addr] != 1:
            raise H5FormatError(f"unsupported object header version {self.buf[addr]} at {addr}")
        nmsg = self._u(addr + 2, 2)
        hsize = self._u(addr + 8, 4)
        blocks = [(addr + 16, addr + 16 + hsize)]
        while blocks and len(out) < nmsg + 64:
            p, end = blocks.pop(0)
            while p + 8 <= end:
                mtype, msize = self._u(p, 2), self._u(p + 2, 2)
                p += 8
                if mtype == 0x10:
                    blocks.append((self._addr(p) + self.base_addr, self._addr(p) + self.base_addr + self._len(p + self.so)))
                elif mtype != 0:
                    out.append((mtype, p, msize))
                p += msize
        return out

    # -- groups --------------------------------------------------------------------------------
    def _group_links(self, addr: int) -> Dict[str, int]:
        links: Dict[str, int] = {}
        for mtype, p, size in self._messages(addr):
            if mtype == 0x11:              # symbol table message: B-tree + local heap
                btree, heap = self._addr(p), self._addr(p + self.so)
                heap += self.base_addr
                if self.buf[heap:heap + 4] != b"HEAP":
                    raise H5FormatError("bad local heap")
                data = self._addr(heap + 8 + 2 * self.sl) + self.base_addr
                self._walk_group_btree(btree, data, links)
            elif mtype == 0x06:            # link message (compact new-style group)
                if self.buf[p] != 1:
                    raise H5FormatError("bad link message")
                flags = self.buf[p + 1]
                q = p + 2
                ltype = 0
                if flags & 0x08:
                    ltype = self.buf[q]
                    q += 1
                if flags & 0x04:
                    q += 8
                if flags & 0x10:
                    q += 1
                nsz = 1 << (flags & 3)
                nlen = self._u(q, nsz)
                q += nsz
                name = self.buf[q:q + nlen].decode("utf-8")
                q += nlen
                if ltype == 0:
                    links[name] = self._addr(q)
            elif mtype == 0x02:            # link info: dense storage when a fractal heap address is present
                q = p + 2 + (8 if self.buf[p + 1] & 1 else 0)
                if self._addr(q) != self._undef():
                    raise H5FormatError("dense (fractal-heap) groups are not supported")
        return links

    def _walk_group_btree(self, addr: int, heap_data: int, links: Dict[str, int]) -> None:
        addr += self.base_addr
        if self.buf[addr:addr + 4] != b"TREE" or self.buf[addr + 4] != 0:
            raise H5FormatError("bad group B-tree node")
        level, used = self.buf[addr + 5], self._u(addr + 6, 2)
        p = addr + 8 + 2 * self.so
        for i in range(used):
            child = self._addr(p + self.sl + i * (self.sl + self.so))
            if level > 0:
                self._walk_group_btree(child, heap_data, links)
                continue
            node = child + self.base_addr
            if self.buf[node:node + 4] != b"SNOD":
                raise H5FormatError("bad symbol table node")
            nsym = self._u(node + 6, 2)
            q = node + 8
            for _ in range(nsym):
                noff, oaddr = self._addr(q), self._addr(q + self.so)
                end = self.buf.index(b"\x00", heap_data + noff)
                links[self.buf[heap_data + noff:end].decode("utf-8")] = oaddr
                q += 2 * self.so + 4 + 4 + 16

    # -- datasets ------------------------------------------------------------------------------
    def keys(self) -> List[str]:
        return sorted(self._root)

    def __contains__(self, key: str) -> bool:
        try:
            self._resolve(key)
            return True
        except KeyError:
            return False

    def _resolve(self, key: str) -> int:
        links = self._root
        parts = [s for s in key.split("/") if s]
        for i, part in enumerate(parts):
            if part not in links:
                raise KeyError(key)
            addr = links[part]
            if i + 1 < len(parts):
                links = self._group_links(addr)
        return addr

    def __getitem__(self, key: str) -> np.ndarray:
        return self._read_dataset(self._resolve(key))

    def _read_dataset(self, addr: int) -> np.ndarray:
        shape = dtype = layout = None
        filters: List[Tuple[int, List[int]]] = []
        for mtype, p, size in self._messages(addr):
            if mtype == 0x01:
                shape = self._dataspace(p)
            elif mtype == 0x03:
                dtype = self._datatype(p)
            elif mtype == 0x08:
                layout = (p, size)
            elif mtype == 0x0B:
                filters = self._filters(p)
        if shape is None or dtype is None or layout is None:
            raise H5FormatError("object is not a dataset")
        count = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        raw = self._layout_bytes(layout[0], shape, dtype.itemsize, count * dtype.itemsize, filters)
        return np.frombuffer(raw, dtype=dtype, count=count).reshape(shape).copy()

    def _dataspace(self, p: int) -> Tuple[int, ...]:
        ver, rank = self.buf[p], self.buf[p + 1]
        if ver == 1:
            q = p + 8
        elif ver == 2:
            if self.buf[p + 3] == 2:
                raise H5FormatError("null dataspace")
            q = p + 4
        else:
            raise H5FormatError(f"unsupported dataspace version {ver}")
        return tuple(self._len(q + i * self.sl) for i in range(rank))

    def _datatype(self, p: int) -> np.dtype:
        cls, ver = self.buf[p] & 0x0F, self.buf[p] >> 4
        bits0 = self.buf[p + 1]
        size = self._u(p + 4, 4)
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{size}")
        if cls == 1:
            if size not in (2, 4, 8):
                raise H5FormatError(f"unsupported float size {size}")
            return np.dtype(f"{order}f{size}")
        raise H5FormatError(f"unsupported datatype class {cls} (version {ver})")

    def _filters(self, p: int) -> List[Tuple[int, List[int]]]:
        ver, n = self.buf[p], self.buf[p + 1]
        q = p + (8 if ver == 1 else 2)
        out = []
        for _ in range(n):
            fid = self._u(q, 2)
            q += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = self._u(q, 2)
                q += 2
            q += 2                      # flags
            ncd = self._u(q, 2)
            q += 2
            q += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            cd = [self._u(q + 4 * i, 4) for i in range(ncd)]
            q += 4 * ncd
            if ver == 1 and ncd % 2:
                q += 4
            out.append((fid, cd))
        return out

    @staticmethod
    def _unfilter(chunk: bytes, filters, mask: int, itemsize: int) -> bytes:
        for i in reversed(range(len(filters))):
            if mask & (1 << i):
                continue
            fid, _cd = filters[i]
            if fid == 1:
                chunk = zlib.decompress(chunk)
            elif fid == 2:
                n = len(chunk) // itemsize
                chunk = np.frombuffer(chunk[:n * itemsize], np.uint8).reshape(itemsize, n).T.tobytes() + chunk[n * itemsize:]
            elif fid == 3:
                chunk = chunk[:-4]
            else:
                raise H5FormatError(f"unsupported filter id {fid}")
        return chunk

    def _layout_bytes(self, p: int, shape, itemsize: int, nbytes: int, filters) -> bytes:
        ver = self.buf[p]
        if ver in (1, 2):
            ndim, cls = self.buf[p + 1], self.buf[p + 2]
            q = p + 8
            addr = None
            if cls != 0:
                addr = self._addr(q)
                q += self.so
            dims = [self._u(q + 4 * i, 4) for i in range(ndim)]
            q += 4 * ndim
            if cls == 0:
                size = self._u(q, 4)
                return self.buf[q + 4:q + 4 + size]
            if cls == 1:
                return self._contiguous(addr, nbytes)
            return self._chunked_v1(addr, shape, dims[:-1], itemsize, filters)
        if ver in (3, 4):
            cls = self.buf[p + 1]
            if cls == 0:
                size = self._u(p + 2, 2)
                return self.buf[p + 4:p + 4 + size]
            if cls == 1:
                return self._contiguous(self._addr(p + 2), nbytes)
            if cls == 2 and ver == 3:
                ndim = self.buf[p + 2]
                addr = self._addr(p + 3)
                dims = [self._u(p + 3 + self.so + 4 * i, 4) for i in range(ndim)]
                return self._chunked_v1(addr, shape, dims[:-1], itemsize, filters)
            if cls == 2:                   # version 4: only the single-chunk index is restated
                flags, ndim, enc = self.buf[p + 2], self.buf[p + 3], self.buf[p + 4]
                q = p + 5 + ndim * enc
                if self.buf[q] != 1:
                    raise H5FormatError("chunk index type other than 'single chunk' is not supported")
                q += 1
                size, mask = nbytes, 0
                if flags & 0x02:
                    size, mask = self._len(q), self._u(q + self.sl, 4)
                    q += self.sl + 4
                addr = self._addr(q) + self.base_addr
                return self._unfilter(self.buf[addr:addr + size], filters, mask, itemsize)
        raise H5FormatError(f"unsupported data layout (version {ver})")

    def _contiguous(self, addr: int, nbytes: int) -> bytes:
        if addr == self._undef():          # never written: fill value (zeros)
            return bytes(nbytes)
        addr += self.base_addr
        if addr + nbytes > len(self.buf):
            raise H5FormatError("dataset extends past the end of the file")
        return self.buf[addr:addr + nbytes]

    def _chunked_v1(self, btree: int, shape, chunk, itemsize: int, filters) -> bytes:
        full = np.zeros(tuple(shape) + (itemsize,), dtype=np.uint8)      # unwritten chunks read as the fill value (0)
        if btree != self._undef():
            self._walk_chunk_btree(btree, full, tuple(chunk), itemsize, filters)
        return full.tobytes()

    def _walk_chunk_btree(self, addr: int, full: np.ndarray, chunk, itemsize: int, filters) -> None:
        addr += self.base_addr
        if self.buf[addr:addr + 4] != b"TREE" or self.buf[addr + 4] != 1:
            raise H5FormatError("bad chunk B-tree node")
        level, used = self.buf[addr + 5], self._u(addr + 6, 2)
        rank = len(chunk)
        ksize = 8 + 8 * (rank + 1)
        p = addr + 8 + 2 * self.so
        for _ in range(used):
            csize, mask = self._u(p, 4), self._u(p + 4, 4)
            offs = [self._u(p + 8 + 8 * i, 8) for i in range(rank)]
            child = self._addr(p + ksize)
            p += ksize + self.so
            if level > 0:
                self._walk_chunk_btree(child, full, chunk, itemsize, filters)
                continue
            child += self.base_addr
            raw = self._unfilter(self.buf[child:child + csize], filters, mask, itemsize)
            block = np.frombuffer(raw, np.uint8, count=int(np.prod(chunk)) * itemsize).reshape(tuple(chunk) + (itemsize,))
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, full.shape))
            full[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]


def read_h5(path, keys: Optional[Iterable[str]] = None) -> Dict[str, np.ndarray]:
    """``{key: h5py.File(path)[key][()]}`` for the given (default: all root) datasets."""
    f = H5File(path)
    return {k: f[k] for k in (list(keys) if keys is not None else f.keys())}


# ---------------------------------------------------------------------------------------------
# writer (synthetic fixtures / tests): the layout h5py emits for ``f[key] = array`` with default settings
# ---------------------------------------------------------------------------------------------
def _dtype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    big = dt.byteorder == ">"
    if dt.kind in "iu":
        bits = (1 if big else 0) | (0x08 if dt.kind == "i" else 0)
        return struct.pack("<BBBBIHH", 0x10, bits, 0, 0, dt.itemsize, 0, dt.itemsize * 8)
    if dt.kind == "f":
        spec = {2: (15, 10, 5, 0, 10, 15), 4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}[dt.itemsize]
        sign, epos, esize, mpos, msize, bias = spec
        bits0 = (1 if big else 0) | 0x20          # mantissa normalisation: msb implied
        return struct.pack("<BBBBIHHBBBBI", 0x11, bits0, sign, 0, dt.itemsize, 0, dt.itemsize * 8, epos, esize, mpos,
                           msize, bias)
    raise H5FormatError(f"unsupported dtype {dt}")


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = body + bytes(-len(body) % 8)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages: Sequence[bytes]) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body


def write_h5(path, arrays: Dict[str, np.ndarray]) -> None:
    """Minimal HDF5 writer: superblock version 0, root group with a symbol table (one B-tree leaf, one symbol node, a
    local heap), one contiguous dataset per entry (at most 2 * 4 = 8 entries per symbol node with the default K = 4)."""
    names = sorted(arrays)
    if not 0 < len(names) <= 8:
        raise H5FormatError("write_h5 holds 1..8 datasets")
    # local heap data: "" at 0 (root name), then the link names, 8-byte aligned
    heap = bytearray(b"\x00" * 8)
    name_off = {}
    for n in names:
        name_off[n] = len(heap)
        heap += n.encode("utf-8") + b"\x00"
        heap += bytes(-len(heap) % 8)
    heap_size = max(88, len(heap) + 16)
    heap_size += -heap_size % 8
    free_off = len(heap)
    heap += struct.pack("<QQ", 1, heap_size - free_off) + bytes(heap_size - free_off - 16)   # one free block to the end

    SB = 96                                   # superblock 0 with 8-byte offsets incl. the root symbol table entry
    root_hdr = _object_header([_msg(0x11, struct.pack("<QQ", 0, 0))])           # patched below
    a_root = SB
    a_btree = a_root + len(root_hdr)
    btree_size = 8 + 16 + (2 * 4 + 1) * 8 + 2 * 4 * 8            # header + keys (2K+1) + children (2K), K = 4
    a_heap = a_btree + btree_size
    a_heapdata = a_heap + 32
    a_snod = a_heapdata + heap_size
    snod_size = 8 + 8 * 40
    a_next = a_snod + snod_size
    # datasets
    hdrs, blobs, a_hdr, a_data = {}, {}, {}, {}
    for n in names:
        arr = np.ascontiguousarray(arrays[n])
        dims = b"".join(struct.pack("<Q", d) for d in arr.shape)
        space = struct.pack("<BBB5x", 1, arr.ndim, 0) + dims
        fill = struct.pack("<BBBB", 2, 2, 2, 0)                   # version 2, late allocation, write if set, undefined
        a_hdr[n] = a_next
        layout_stub = struct.pack("<BBQQ", 3, 1, 0, arr.nbytes)
        hdr = _object_header([_msg(0x01, space), _msg(0x03, _dtype_message(arr.dtype), 1), _msg(0x05, fill),
                              _msg(0x08, layout_stub)])
        a_data[n] = a_next + len(hdr)
        layout = struct.pack("<BBQQ", 3, 1, a_data[n], arr.nbytes)
        hdrs[n] = _object_header([_msg(0x01, space), _msg(0x03, _dtype_message(arr.dtype), 1), _msg(0x05, fill),
                                  _msg(0x08, layout)])
        blobs[n] = arr.tobytes()
        a_next = a_data[n] + len(blobs[n])
        a_next += -a_next % 8
    eof = a_next
    root_hdr = _object_header([_msg(0x11, struct.pack("<QQ", a_btree, a_heap))])
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, a_root, 1, 0) + struct.pack("<QQ", a_btree, a_heap)       # cached symbol table entry
    assert len(sb) == SB
    btree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF)
    btree += struct.pack("<QQQ", 0, a_snod, name_off[names[-1]])                              # key0, child0, key1
    btree += bytes(btree_size - len(btree))
    heap_hdr = b"HEAP" + struct.pack("<B3xQQQ", 0, heap_size, free_off, a_heapdata)
    snod = b"SNOD" + struct.pack("<BxH", 1, len(names))
    for n in names:
        snod += struct.pack("<QQII16x", name_off[n], a_hdr[n], 0, 0)
    snod += bytes(snod_size - len(snod))
    out = bytearray(sb + root_hdr + btree + heap_hdr + bytes(heap) + snod)
    for n in names:
        assert len(out) == a_hdr[n]
        out += hdrs[n] + blobs[n]
        out += bytes(-len(out) % 8)
    with open(path, "wb") as fh:
        fh.write(bytes(out))


# ---------------------------------------------------------------------------------------------
# batches
# ---------------------------------------------------------------------------------------------
class H5SliceLoader:
    """``layer_{i}.h5`` slice files -> pinned ``[B, C, H, W]`` fp32 batches.

    ``keys``: datasets stacked along C in this order (the reference's ``train_keys`` + target key,
    trainers/trainer_ds_diff.py:256-265, 543-557).  ``pad_to``: optional (H, W) to zero-pad (top-left anchored centre
    padding like ``DivisiblePadd``: the extra rows / columns are split evenly, the odd one goes to the end) so a 6-level
    UNet gets H, W % 32 == 0 (SURVEY §8 gotcha 8).  Yields ``(paths, batch)``; ``batch`` is a view of one of two pinned
    buffers used alternately, so the caller may issue ``batch.cuda(non_blocking=True)`` and keep loading the next one."""

    def __init__(self, paths: Sequence[str], keys: Sequence[str], batch_size: int, pad_to: Optional[Tuple[int, int]] = None,
                 pin: bool = True):
        self.paths, self.keys, self.batch_size, self.pad_to, self.pin = list(paths), list(keys), int(batch_size), pad_to, pin
        self._bufs = [None, None]

    def __len__(self) -> int:
        return -(-len(self.paths) // self.batch_size)

    def _buffer(self, i: int, shape):
        import torch
        b = self._bufs[i]
        if b is None or tuple(b.shape) != tuple(shape):
            b = torch.zeros(shape, dtype=torch.float32)
            if self.pin and torch.cuda.is_available():
                b = b.pin_memory()
            self._bufs[i] = b
        return b

    def __iter__(self) -> Iterator[Tuple[List[str], "object"]]:
        flip = 0
        for s in range(0, len(self.paths), self.batch_size):
            chunk = self.paths[s:s + self.batch_size]
            first = H5File(chunk[0])
            h, w = first[self.keys[0]].shape[-2:]
            H, W = self.pad_to or (h, w)
            if H < h or W < w:
                raise H5FormatError(f"pad_to {self.pad_to} is smaller than the slices ({h}, {w})")
            top, left = (H - h) // 2, (W - w) // 2
            buf = self._buffer(flip, (self.batch_size, len(self.keys), H, W))
            out = buf[:len(chunk)]
            if (H, W) != (h, w):
                out.zero_()
            dst = out.numpy()
            for i, path in enumerate(chunk):
                f = first if i == 0 else H5File(path)
                for c, key in enumerate(self.keys):
                    a = f[key]
                    if a.shape[-2:] != (h, w):
                        raise H5FormatError(f"{path}:{key} has shape {a.shape}, expected (..., {h}, {w})")
                    dst[i, c, top:top + h, left:left + w] = a.reshape(h, w)
            flip ^= 1
            yield chunk, out
