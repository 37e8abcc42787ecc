"""Feature-file container: a small reader / writer of the HDF5 subset STAMP's feature files use.

The reference stores every slide's features as an HDF5 file through h5py (writer:
src/stamp/preprocessing/__init__.py:342-366 -- datasets ``coords`` and ``feats`` plus scalar root
attributes; reader: src/stamp/modeling/data.py:584-655 and ``get_coords`` :741-808).  h5py / libhdf5
are not part of this image, so the container is written here directly from the HDF5 file-format
specification, restricted to what those call sites touch:

* writer: "earliest" file format as h5py's default ``File(path, "w")`` produces it -- version-0
  superblock, root group as symbol table (local heap + version-1 B-tree + one symbol node), version-1
  object headers, contiguous little-endian datasets, scalar / array attributes, Python ``str``
  attributes as variable-length UTF-8 strings in a global heap collection (so that
  ``attrs["unit"] == "um"`` holds for an h5py reader exactly as it does for h5py-written files);
* reader: superblock versions 0-3, version-1 and version-2 object headers (with continuation
  blocks), symbol-table groups and compact link messages, contiguous / compact / chunked
  (version-1 chunk B-tree; deflate, shuffle, fletcher32) layouts, fixed-point, floating-point,
  fixed and variable-length string datatypes, attribute messages versions 1-3.

The API is the slice of h5py's that the reference uses (``File`` as context manager, ``in``,
``f[name]``, ``f[name] = array``, ``ds[()]`` / ``ds[:]`` / ``ds.shape`` / ``ds.dtype``, ``f.attrs``).
Pinned by ``tests/test_h5lite_cpu.py``: the reader against a file written by libhdf5 itself
(``tests/golden/libhdf5_matlab73.mat``), the writer against the reader and against the byte layout
of that file's structures.
"""

from __future__ import annotations

import os
import struct
import zlib
from collections.abc import Iterator, Mapping
from typing import Any, BinaryIO

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(OSError):
    """Malformed or unsupported file content (h5py raises OSError for the same situations)."""


def _pad8(n: int) -> int:
    return (n + 7) & ~7


_PARSE_ERRORS = (struct.error, IndexError, ValueError, TypeError, OverflowError, MemoryError, RecursionError,
                 UnicodeDecodeError, zlib.error)


class _guard:
    """Turns whatever a corrupt structure makes the parser raise into H5Error (an OSError, like h5py's)."""

    def __init__(self, what: str) -> None:
        self.what = what

    def __enter__(self) -> None:
        return None

    def __exit__(self, exc_type, exc, _tb) -> bool:  # noqa: ANN001
        if exc_type is not None and issubclass(exc_type, _PARSE_ERRORS) and not issubclass(exc_type, H5Error):
            raise H5Error(f"{self.what}: truncated or corrupt HDF5 structure ({exc_type.__name__}: {exc})") from exc
        return False


# --------------------------------------------------------------------------------------------------
# datatypes <-> numpy
# --------------------------------------------------------------------------------------------------

_VLEN_STR = "vlen-str"


def _encode_dtype(dt: np.dtype) -> bytes:
    """Datatype message (version 1) for a numpy dtype."""
    dt = np.dtype(dt)
    if dt.byteorder == ">":
        raise H5Error("big-endian arrays are not written; convert first")
    size = dt.itemsize
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBIHH", 0x10, bits0, 0, 0, size, 0, size * 8)
    if dt.kind == "f":
        try:
            exp_bits, man_bits, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[size]
        except KeyError:
            raise H5Error(f"unsupported float width {size}") from None
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, size * 8 - 1, 0, size, 0, size * 8, man_bits, exp_bits, 0,
                           man_bits, bias)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, size)  # null-terminated, ASCII
    raise H5Error(f"dtype {dt} has no mapping to an HDF5 datatype here")


def _vlen_str_dtype(utf8: bool = True) -> bytes:
    # class 9 (variable length), type = string, null-terminated, character set; base type: 1-byte unsigned
    base = struct.pack("<BBBBIHH", 0x10, 0, 0, 0, 1, 0, 8)
    return struct.pack("<BBBBI", 0x19, 0x01, 0x01 if utf8 else 0x00, 0, 16) + base


def _decode_dtype(buf: bytes) -> tuple[Any, int]:
    """-> (numpy dtype | _VLEN_STR, element size in the file)."""
    cls_ver, b0, b1, _b2, size = struct.unpack_from("<BBBBI", buf, 0)
    cls, ver = cls_ver & 0x0F, cls_ver >> 4
    if ver not in (1, 2, 3):
        raise H5Error(f"datatype message version {ver}")
    order = ">" if b0 & 1 else "<"
    if cls == 0:
        if size not in (1, 2, 4, 8):
            raise H5Error(f"fixed-point type of {size} bytes")
        return np.dtype(f"{order}{'i' if b0 & 0x08 else 'u'}{size}"), size
    if cls == 1:
        if size not in (2, 4, 8):
            raise H5Error(f"floating-point type of {size} bytes")
        return np.dtype(f"{order}f{size}"), size
    if cls == 3:
        return np.dtype(f"S{size}"), size
    if cls == 9:
        if b0 & 0x0F != 1:
            raise H5Error("variable-length sequences are not supported (only strings)")
        return _VLEN_STR, size
    if cls == 8:  # enumeration: h5py stores numpy bool as an enum over int8
        base, bsize = _decode_dtype(buf[8:])
        return base, bsize
    raise H5Error(f"datatype class {cls} is not supported")


# --------------------------------------------------------------------------------------------------
# writer
# --------------------------------------------------------------------------------------------------


class _GlobalHeapWriter:
    """One global heap collection holding the variable-length strings of the attributes."""

    def __init__(self) -> None:
        self.objects: list[bytes] = []

    def add(self, data: bytes) -> int:
        self.objects.append(data)
        return len(self.objects)  # heap object indices start at 1 (0 is the free-space object)

    def encode(self) -> bytes:
        body = b""
        for i, data in enumerate(self.objects, start=1):
            body += struct.pack("<HHIQ", i, 1, 0, len(data)) + data.ljust(_pad8(len(data)), b"\0")
        size = max(4096, _pad8(16 + len(body) + 16))
        free = size - 16 - len(body)
        body += struct.pack("<HHIQ", 0, 0, 0, free)  # object 0: the remaining free space (size includes its header)
        return (b"GCOL" + struct.pack("<B3xQ", 1, size) + body).ljust(size, b"\0")


def _dataspace_msg(shape: tuple[int, ...]) -> bytes:
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = body.ljust(_pad8(len(body)), b"\0")
    if len(body) > 0xFFF8:
        raise H5Error("object header message larger than 64 KiB (attribute too large)")
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages: list[bytes]) -> bytes:
    body = b"".join(messages)
    # 12-byte prefix + 4 bytes of padding so that messages start 8-aligned
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body


class _AttrsWriter(dict):
    """Python values waiting to be encoded at close."""


def _attr_message(name: str, value: Any, gheap: _GlobalHeapWriter, gheap_addr_slot: list[int]) -> bytes:
    """Attribute message, version 1.  Variable-length strings reference the global heap collection whose
    address is only known at layout time: the 8 address bytes are written as a placeholder and their
    offsets inside the returned message recorded in ``gheap_addr_slot``."""
    name_b = name.encode("utf-8") + b"\0"
    slots: list[int] = []
    if isinstance(value, (str, bytes)) and not isinstance(value, np.bytes_):  # np.bytes_ -> fixed-length, as in h5py
        raw = value.encode("utf-8") if isinstance(value, str) else value
        dt = _vlen_str_dtype(utf8=isinstance(value, str))
        ds = struct.pack("<BBB5x", 1, 0, 0)  # scalar
        idx = gheap.add(raw)
        data = struct.pack("<I", len(raw)) + b"\0" * 8 + struct.pack("<I", idx)
        slots.append(4)
    else:
        arr = np.asarray(value)
        if arr.dtype == object or arr.dtype.kind == "U":
            raise H5Error(f"attribute {name!r}: arrays of Python objects / unicode are not supported")
        if arr.dtype.kind == "b":
            arr = arr.astype(np.int8)
        arr = np.ascontiguousarray(arr.astype(arr.dtype.newbyteorder("<"), copy=False)).reshape(arr.shape)
        dt = _encode_dtype(arr.dtype)
        ds = _dataspace_msg(arr.shape)
        data = arr.tobytes()
    head = struct.pack("<BxHHH", 1, len(name_b), len(dt), len(ds))
    parts = head + name_b.ljust(_pad8(len(name_b)), b"\0") + dt.ljust(_pad8(len(dt)), b"\0") + ds.ljust(
        _pad8(len(ds)), b"\0")
    msg = _message(0x000C, parts + data)
    for s in slots:
        gheap_addr_slot.append(8 + len(parts) + s)
    return msg


class _PendingDataset:
    def __init__(self, data: np.ndarray) -> None:
        self.data = data
        self.attrs = _AttrsWriter()

    @property
    def shape(self) -> tuple[int, ...]:
        return self.data.shape

    @property
    def dtype(self) -> np.dtype:
        return self.data.dtype


def _write_file(fp: BinaryIO, datasets: dict[str, _PendingDataset], root_attrs: Mapping[str, Any]) -> None:
    names = sorted(datasets, key=lambda s: s.encode("utf-8"))  # symbol nodes are ordered by strcmp
    for n in names:
        if "/" in n or not n:
            raise H5Error(f"dataset name {n!r}: only root-level datasets are written")
    leaf_k = max(4, (len(names) + 1) // 2)
    internal_k = 16
    gheap = _GlobalHeapWriter()

    # ---- local heap data segment: "" at offset 0, then the link names
    heap_data = bytearray(8)
    name_off = {}
    for n in names:
        name_off[n] = len(heap_data)
        nb = n.encode("utf-8") + b"\0"
        heap_data += nb.ljust(_pad8(len(nb)), b"\0")
    heap_size = max(_pad8(len(heap_data)) + 16, 88)
    free_off = len(heap_data)
    heap_data += struct.pack("<QQ", 1, heap_size - free_off)  # one free block: (next = 1: none, size)
    heap_data = bytes(heap_data).ljust(heap_size, b"\0")

    # ---- messages whose sizes fix the layout (addresses patched afterwards)
    root_slots: list[int] = []
    root_attr_msgs = []
    for k, v in root_attrs.items():
        slots: list[int] = []
        m = _attr_message(k, v, gheap, slots)
        root_attr_msgs.append((m, slots))
    ds_attr_msgs: dict[str, list[tuple[bytes, list[int]]]] = {}
    for n in names:
        lst = []
        for k, v in datasets[n].attrs.items():
            slots = []
            lst.append((_attr_message(k, v, gheap, slots), slots))
        ds_attr_msgs[n] = lst
    del root_slots

    sb_size = 56 + 40
    root_ohdr_addr = sb_size
    stab_msg_len = 8 + 16
    root_ohdr_len = 16 + stab_msg_len + sum(len(m) for m, _ in root_attr_msgs)
    heap_addr = _pad8(root_ohdr_addr + root_ohdr_len)
    heap_seg_addr = heap_addr + 32
    btree_addr = heap_seg_addr + heap_size
    btree_len = 24 + (2 * internal_k + 1) * 8 + 2 * internal_k * 8
    snod_addr = btree_addr + btree_len
    snod_len = 8 + 2 * leaf_k * 40
    cursor = snod_addr + snod_len

    ds_ohdr_addr: dict[str, int] = {}
    ds_ohdr_len: dict[str, int] = {}
    for n in names:
        d = datasets[n]
        fixed = (8 + _pad8(len(_dataspace_msg(d.shape)))) + (8 + _pad8(len(_encode_dtype(d.dtype)))) + (8 + 8) + (
            8 + 24)
        ds_ohdr_addr[n] = cursor
        ds_ohdr_len[n] = 16 + fixed + sum(len(m) for m, _ in ds_attr_msgs[n])
        cursor = _pad8(cursor + ds_ohdr_len[n])
    gheap_addr = cursor if gheap.objects else _UNDEF
    gheap_bytes = gheap.encode() if gheap.objects else b""
    cursor += len(gheap_bytes)
    data_addr: dict[str, int] = {}
    for n in names:
        nbytes = datasets[n].data.nbytes
        if nbytes == 0:
            data_addr[n] = _UNDEF
            continue
        cursor = (cursor + 63) & ~63
        data_addr[n] = cursor
        cursor += nbytes
    eof = cursor

    def patch(msg: bytes, slots: list[int]) -> bytes:
        if not slots:
            return msg
        b = bytearray(msg)
        for s in slots:
            b[s:s + 8] = struct.pack("<Q", gheap_addr)
        return bytes(b)

    # ---- superblock (version 0) with the root group's symbol table entry
    out = bytearray()
    out += _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, leaf_k, internal_k, 0)
    out += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
    out += struct.pack("<QQII", 0, root_ohdr_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)
    assert len(out) == sb_size
    # ---- root object header
    stab = _message(0x0011, struct.pack("<QQ", btree_addr, heap_addr))
    out += _object_header([stab] + [patch(m, s) for m, s in root_attr_msgs])
    out = out.ljust(heap_addr, b"\0")
    # ---- local heap
    out += b"HEAP" + struct.pack("<B3xQQQ", 0, heap_size, free_off, heap_seg_addr) + heap_data
    # ---- B-tree node (group node, leaf level) with one child
    keys_children = struct.pack("<QQQ", 0, snod_addr, name_off[names[-1]]) if names else b""
    out += (b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, _UNDEF, _UNDEF) + keys_children).ljust(
        btree_len, b"\0")
    # ---- symbol node
    snod = b"SNOD" + struct.pack("<BxH", 1, len(names))
    for n in names:
        snod += struct.pack("<QQII16x", name_off[n], ds_ohdr_addr[n], 0, 0)
    out += snod.ljust(snod_len, b"\0")
    # ---- dataset object headers
    for n in names:
        d = datasets[n]
        msgs = [
            _message(0x0001, _dataspace_msg(d.shape)),
            _message(0x0003, _encode_dtype(d.dtype), flags=1),
            _message(0x0005, struct.pack("<BBBB", 2, 2, 2, 0), flags=1),  # fill value v2: late alloc, write if set, undefined
            _message(0x0008, struct.pack("<BBQQ", 3, 1, data_addr[n], d.data.nbytes)),
        ] + [patch(m, s) for m, s in ds_attr_msgs[n]]
        out = out.ljust(ds_ohdr_addr[n], b"\0")
        hdr = _object_header(msgs)
        assert len(hdr) == ds_ohdr_len[n], (len(hdr), ds_ohdr_len[n])
        out += hdr
    if gheap.objects:
        out = out.ljust(gheap_addr, b"\0")
        out += gheap_bytes
    fp.write(bytes(out))
    pos = len(out)
    for n in names:
        if data_addr[n] == _UNDEF:
            continue
        fp.write(b"\0" * (data_addr[n] - pos))
        fp.write(memoryview(datasets[n].data).cast("B"))
        pos = data_addr[n] + datasets[n].data.nbytes
    assert pos == eof or not names or all(a == _UNDEF for a in data_addr.values())
    if pos < eof:
        fp.write(b"\0" * (eof - pos))


# --------------------------------------------------------------------------------------------------
# reader
# --------------------------------------------------------------------------------------------------


class _Reader:
    def __init__(self, fp: BinaryIO) -> None:
        self.fp = fp
        fp.seek(0, os.SEEK_END)
        self.file_size = fp.tell()
        self.base = 0
        self._gheap_cache: dict[int, dict[int, bytes]] = {}
        self._superblock()

    def read(self, addr: int, n: int) -> bytes:
        if addr == _UNDEF or addr + self.base + n > self.file_size:
            raise H5Error(f"read of {n} bytes at address {addr:#x} runs past the end of the file")
        self.fp.seek(self.base + addr)
        return self.fp.read(n)

    def _superblock(self) -> None:
        off = 0
        while True:
            self.fp.seek(off)
            if self.fp.read(8) == _SIG:
                break
            off = 512 if off == 0 else off * 2
            if off >= self.file_size:
                raise H5Error("not an HDF5 file (no superblock signature)")
        self.fp.seek(off)
        sb = self.fp.read(128)
        ver = sb[8]
        if ver in (0, 1):
            if sb[13] != 8 or sb[14] != 8:
                raise H5Error("only 8-byte offsets and lengths are supported")
            p = 24 + (4 if ver == 1 else 0)
            base, _fs, _eof, _drv = struct.unpack_from("<QQQQ", sb, p)
            _name, ohdr, cache, _r = struct.unpack_from("<QQII", sb, p + 32)
            self.base = base
            self.root_ohdr = ohdr
        elif ver in (2, 3):
            if sb[9] != 8 or sb[10] != 8:
                raise H5Error("only 8-byte offsets and lengths are supported")
            base, _ext, _eof, ohdr = struct.unpack_from("<QQQQ", sb, 12)
            self.base = base
            self.root_ohdr = ohdr
        else:
            raise H5Error(f"superblock version {ver}")
        if self.base == 0 and off:
            self.base = off

    # ---- object headers --------------------------------------------------------------------------
    def messages(self, addr: int) -> list[tuple[int, bytes]]:
        head = self.read(addr, 16)
        out: list[tuple[int, bytes]] = []
        if head[:4] == b"OHDR":
            if head[4] != 2:
                raise H5Error(f"object header version {head[4]}")
            flags = head[5]
            p = 6 + (16 if flags & 0x20 else 0) + (4 if flags & 0x10 else 0)
            width = 1 << (flags & 3)
            head = self.read(addr, p + width)
            size0 = int.from_bytes(head[p:p + width], "little")
            chunks = [(addr + p + width, size0)]
            track_order = bool(flags & 0x04)
            while chunks:
                caddr, clen = chunks.pop(0)
                buf = self.read(caddr, clen)
                q = 0
                while q + 4 <= clen:
                    mtype, msize, _mflags = struct.unpack_from("<BHB", buf, q)
                    q += 4 + (2 if track_order else 0)
                    body = buf[q:q + msize]
                    q += msize
                    if mtype == 0x10:
                        coff, clen2 = struct.unpack_from("<QQ", body, 0)
                        chunks.append((coff + 4, clen2 - 8))  # skip "OCHK", drop the checksum
                    elif mtype != 0:
                        out.append((mtype, body))
            return out
        ver, _r, nmsg, _ref, hsize = struct.unpack_from("<BBHII", head, 0)
        if ver != 1:
            raise H5Error(f"object header version {ver} at {addr:#x}")
        chunks = [(addr + 16, hsize)]
        while chunks and len(out) < nmsg + 64:
            caddr, clen = chunks.pop(0)
            buf = self.read(caddr, clen)
            q = 0
            while q + 8 <= clen:
                mtype, msize, _mflags = struct.unpack_from("<HHB", buf, q)
                body = buf[q + 8:q + 8 + msize]
                if len(body) != msize:
                    raise H5Error("object header message overruns its chunk")
                q += 8 + msize
                if mtype == 0x10:
                    chunks.append(struct.unpack_from("<QQ", body, 0))
                elif mtype != 0:
                    out.append((mtype, body))
        return out

    # ---- groups ------------------------------------------------------------------------------------
    def links(self, msgs: list[tuple[int, bytes]]) -> dict[str, int] | None:
        """name -> object header address, or None when the object is not a group."""
        found = None
        for mtype, body in msgs:
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", body, 0)
                found = dict(self._symbol_table(btree, heap))
            elif mtype == 0x06:
                found = found if found is not None else {}
                ver, flags = body[0], body[1]
                if ver != 1:
                    raise H5Error(f"link message version {ver}")
                p = 2
                ltype = 0
                if flags & 0x08:
                    ltype = body[p]
                    p += 1
                if flags & 0x04:
                    p += 8
                if flags & 0x10:
                    p += 1
                w = 1 << (flags & 3)
                nlen = int.from_bytes(body[p:p + w], "little")
                p += w
                name = body[p:p + nlen].decode("utf-8")
                p += nlen
                if ltype == 0:
                    found[name] = struct.unpack_from("<Q", body, p)[0]
            elif mtype == 0x02:
                found = found if found is not None else {}
                flags = body[1]
                p = 2 + (8 if flags & 1 else 0)
                fheap = struct.unpack_from("<Q", body, p)[0]
                if fheap != _UNDEF:
                    raise H5Error("groups with dense link storage (fractal heap) are not supported")
        return found

    def _heap_string(self, heap_seg: int, seg_size: int, off: int) -> str:
        raw = self.read(heap_seg + off, min(1024, seg_size - off))
        return raw[:raw.index(b"\0")].decode("utf-8")

    def _symbol_table(self, btree: int, heap: int) -> Iterator[tuple[str, int]]:
        h = self.read(heap, 32)
        if h[:4] != b"HEAP":
            raise H5Error("local heap signature missing")
        seg_size, _free, seg = struct.unpack_from("<QQQ", h, 8)
        yield from self._group_node(btree, seg, seg_size)

    def _group_node(self, addr: int, seg: int, seg_size: int) -> Iterator[tuple[str, int]]:
        sig = self.read(addr, 8)
        if sig[:4] == b"TREE":
            ntype, level, used = struct.unpack_from("<BBH", sig, 4)
            if ntype != 0:
                raise H5Error("group B-tree expected")
            body = self.read(addr + 24, (2 * used + 1) * 8)
            for i in range(used):
                child = struct.unpack_from("<Q", body, 8 + 16 * i)[0]
                yield from self._group_node(child, seg, seg_size)
            del level
        elif sig[:4] == b"SNOD":
            n = struct.unpack_from("<H", sig, 6)[0]
            body = self.read(addr + 8, n * 40)
            for i in range(n):
                noff, ohdr, _cache = struct.unpack_from("<QQI", body, 40 * i)
                yield self._heap_string(seg, seg_size, noff), ohdr
        else:
            raise H5Error(f"unexpected block {sig[:4]!r} in a group index")

    # ---- attributes and data -------------------------------------------------------------------------
    def global_heap_object(self, coll: int, index: int) -> bytes:
        if coll not in self._gheap_cache:
            head = self.read(coll, 16)
            if head[:4] != b"GCOL":
                raise H5Error("global heap collection signature missing")
            size = struct.unpack_from("<Q", head, 8)[0]
            buf = self.read(coll, size)
            objs: dict[int, bytes] = {}
            q = 16
            while q + 16 <= size:
                idx, _ref, _r, osize = struct.unpack_from("<HHIQ", buf, q)
                if idx == 0:
                    break
                objs[idx] = buf[q + 16:q + 16 + osize]
                q += 16 + _pad8(osize)
            self._gheap_cache[coll] = objs
        try:
            return self._gheap_cache[coll][index]
        except KeyError:
            raise H5Error(f"global heap object {index} missing") from None

    def decode_elements(self, dtype: Any, shape: tuple[int, ...], raw: bytes) -> Any:
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if dtype is _VLEN_STR:
            vals = []
            for i in range(count):
                n, coll, idx = struct.unpack_from("<IQI", raw, 16 * i)
                vals.append(self.global_heap_object(coll, idx)[:n].decode("utf-8") if n else "")
            if not shape:
                return vals[0]
            return np.array(vals, dtype=object).reshape(shape)
        arr = np.frombuffer(raw, dtype=dtype, count=count).reshape(shape)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        return arr[()] if not shape else arr.copy()

    @staticmethod
    def parse_dataspace(body: bytes) -> tuple[int, ...] | None:
        ver, rank, flags = body[0], body[1], body[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            if body[3] == 2:
                return None  # null dataspace
            p = 4
        else:
            raise H5Error(f"dataspace message version {ver}")
        del flags
        return tuple(struct.unpack_from("<Q", body, p + 8 * i)[0] for i in range(rank))

    def attributes(self, msgs: list[tuple[int, bytes]]) -> dict[str, Any]:
        out: dict[str, Any] = {}
        for mtype, body in msgs:
            if mtype == 0x15:
                fheap = struct.unpack_from("<Q", body, 2 + (2 if body[1] & 1 else 0))[0]
                if fheap != _UNDEF:
                    raise H5Error("dense attribute storage (fractal heap) is not supported")
            if mtype != 0x0C:
                continue
            ver = body[0]
            nlen, dlen, slen = struct.unpack_from("<HHH", body, 2)
            if ver == 1:
                p = 8
                rnd = _pad8
            elif ver in (2, 3):
                if body[1] & 3:
                    raise H5Error("shared attribute datatypes / dataspaces are not supported")
                p = 8 + (1 if ver == 3 else 0)
                rnd = lambda n: n  # noqa: E731
            else:
                raise H5Error(f"attribute message version {ver}")
            name = body[p:p + nlen].split(b"\0")[0].decode("utf-8")
            p += rnd(nlen)
            dtype, esize = _decode_dtype(body[p:p + dlen])
            p += rnd(dlen)
            shape = self.parse_dataspace(body[p:p + slen])
            p += rnd(slen)
            if shape is None:
                out[name] = None
                continue
            count = int(np.prod(shape, dtype=np.int64)) if shape else 1
            out[name] = self.decode_elements(dtype, shape, body[p:p + count * esize])
        return out

    def read_dataset(self, msgs: list[tuple[int, bytes]], shape: tuple[int, ...], dtype: Any, esize: int) -> Any:
        layout = next((b for t, b in msgs if t == 0x08), None)
        if layout is None:
            raise H5Error("dataset without a layout message")
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        nbytes = count * esize
        ver = layout[0]
        if ver in (1, 2):
            ndim, cls = layout[1], layout[2]
            if cls == 1:
                addr = struct.unpack_from("<Q", layout, 8)[0]
                raw = self.read(addr, nbytes) if nbytes else b""
            elif cls == 0:
                size = struct.unpack_from("<I", layout, 8 + 4 * ndim)[0]
                raw = layout[12 + 4 * ndim:12 + 4 * ndim + size]
            else:
                addr = struct.unpack_from("<Q", layout, 8)[0]
                chunk = struct.unpack_from(f"<{ndim}I", layout, 16)
                raw = self._read_chunked(msgs, addr, chunk[:-1], shape, esize)
        elif ver == 3:
            cls = layout[1]
            if cls == 1:
                addr, size = struct.unpack_from("<QQ", layout, 2)
                if nbytes and addr == _UNDEF:
                    raw = bytes(nbytes)  # never written: fill value (zeros)
                else:
                    if size < nbytes:
                        raise H5Error("contiguous dataset smaller than its dataspace")
                    raw = self.read(addr, nbytes) if nbytes else b""
            elif cls == 0:
                size = struct.unpack_from("<H", layout, 2)[0]
                raw = layout[4:4 + size]
            elif cls == 2:
                ndim = layout[2]
                addr = struct.unpack_from("<Q", layout, 3)[0]
                chunk = struct.unpack_from(f"<{ndim}I", layout, 11)
                raw = self._read_chunked(msgs, addr, chunk[:-1], shape, esize)
            else:
                raise H5Error(f"layout class {cls}")
        else:
            raise H5Error(f"data layout message version {ver} (written with libver='latest'?)")
        if len(raw) < nbytes:
            raise H5Error("dataset storage shorter than its dataspace")
        return self.decode_elements(dtype, shape, raw[:nbytes])

    def _read_chunked(self, msgs, btree: int, chunk: tuple[int, ...], shape: tuple[int, ...], esize: int) -> bytes:
        filters: list[tuple[int, tuple[int, ...]]] = []
        for t, b in msgs:
            if t != 0x0B:
                continue
            ver, nf = b[0], b[1]
            p = 8 if ver == 1 else 2
            for _ in range(nf):
                fid, = struct.unpack_from("<H", b, p)
                if ver == 1 or fid >= 256:
                    nlen, _fl, ncd = struct.unpack_from("<HHH", b, p + 2)
                    p += 8 + (_pad8(nlen) if ver == 1 else nlen)
                else:
                    _fl, ncd = struct.unpack_from("<HH", b, p + 2)
                    p += 6
                cd = struct.unpack_from(f"<{ncd}I", b, p)
                p += 4 * ncd + (4 if ver == 1 and ncd % 2 else 0)
                filters.append((fid, cd))
        out = np.zeros(shape, dtype=f"V{esize}")
        if btree == _UNDEF or out.size == 0:
            return out.tobytes()
        rank = len(shape)

        def walk(addr: int) -> None:
            head = self.read(addr, 24)
            if head[:4] != b"TREE" or head[4] != 1:
                raise H5Error("chunk B-tree node expected")
            level, used = head[5], struct.unpack_from("<H", head, 6)[0]
            ksize = 8 + 8 * (rank + 1)
            body = self.read(addr + 24, used * (ksize + 8) + ksize)
            for i in range(used):
                q = i * (ksize + 8)
                csize, mask = struct.unpack_from("<II", body, q)
                offs = struct.unpack_from(f"<{rank}Q", body, q + 8)
                child = struct.unpack_from("<Q", body, q + ksize)[0]
                if level:
                    walk(child)
                    continue
                raw = self.read(child, csize)
                for n, (fid, cd) in reversed(list(enumerate(filters))):
                    if mask & (1 << n):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        w = cd[0] if cd else esize
                        raw = np.frombuffer(raw, np.uint8).reshape(w, -1).T.tobytes() if w > 1 else raw
                    elif fid == 3:
                        raw = raw[:-4]
                    else:
                        raise H5Error(f"filter {fid} is not supported")
                block = np.frombuffer(raw, dtype=f"V{esize}", count=int(np.prod(chunk))).reshape(chunk)
                sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, shape))
                src = tuple(slice(0, s.stop - s.start) for s in sel)
                out[sel] = block[src]

        walk(btree)
        return out.tobytes()


class Dataset:
    """Read-side handle: shape / dtype come from the object header, data is read on indexing."""

    def __init__(self, reader: _Reader, name: str, msgs: list[tuple[int, bytes]]) -> None:
        self._r, self.name, self._msgs = reader, name, msgs
        space = next((b for t, b in msgs if t == 0x01), None)
        dtype = next((b for t, b in msgs if t == 0x03), None)
        if space is None or dtype is None:
            raise H5Error(f"{name}: not a dataset")
        with _guard(name):
            shape = reader.parse_dataspace(space)
            self.shape: tuple[int, ...] = shape if shape is not None else ()
            self._dtype, self._esize = _decode_dtype(dtype)
        self._attrs: dict[str, Any] | None = None

    @property
    def dtype(self) -> np.dtype:
        return np.dtype(object) if self._dtype is _VLEN_STR else self._dtype

    @property
    def attrs(self) -> dict[str, Any]:
        if self._attrs is None:
            with _guard(self.name):
                self._attrs = self._r.attributes(self._msgs)
        return self._attrs

    def __len__(self) -> int:
        if not self.shape:
            raise TypeError("scalar dataset has no length")
        return self.shape[0]

    def __getitem__(self, key: Any) -> Any:
        with _guard(self.name):
            arr = self._read_all()
        if key is Ellipsis or (isinstance(key, tuple) and len(key) == 0):
            return arr
        return arr[key]

    def _read_all(self) -> Any:
        if self._dtype is not _VLEN_STR and int(np.prod(self.shape, dtype=np.float64)) * self._esize > 64 * max(self._r.file_size, 1 << 20):
            raise H5Error(f"{self.name}: dataspace {self.shape} is larger than the file could hold")
        if self.shape and self._dtype is not _VLEN_STR and self._dtype.byteorder != ">" and int(np.prod(self.shape)) > 0:
            arr = np.empty(self.shape, dtype=self._dtype)      # contiguous storage: one read straight into the result
            self.read_direct(arr)
        else:
            arr = self._r.read_dataset(self._msgs, self.shape, self._dtype, self._esize)
        return arr

    def __array__(self, dtype=None, copy=None):  # noqa: ANN001
        arr = self[()]
        return arr if dtype is None else arr.astype(dtype)

    def read_direct(self, dest: np.ndarray) -> None:
        """Read the whole dataset into ``dest`` (same shape and dtype, C-contiguous; e.g. a view of a pinned staging
        buffer).  Contiguous little-endian storage goes from the file into ``dest`` without an intermediate copy."""
        if tuple(dest.shape) != self.shape or dest.dtype != self.dtype or not dest.flags.c_contiguous:
            raise TypeError(f"read_direct needs a C-contiguous {self.dtype} array of shape {self.shape}")
        layout = next((b for t, b in self._msgs if t == 0x08), b"")
        if len(layout) >= 18 and layout[0] == 3 and layout[1] == 1 and self._dtype is not _VLEN_STR \
                and self._dtype.byteorder != ">":
            addr, size = struct.unpack_from("<QQ", layout, 2)
            if addr != _UNDEF and size >= dest.nbytes and self._r.base + addr + dest.nbytes <= self._r.file_size:
                self._r.fp.seek(self._r.base + addr)
                if self._r.fp.readinto(memoryview(dest).cast("B")) != dest.nbytes:
                    raise H5Error(f"{self.name}: short read")
                return
        dest[...] = self._r.read_dataset(self._msgs, self.shape, self._dtype, self._esize)


class Group:
    def __init__(self, reader: _Reader, name: str, msgs: list[tuple[int, bytes]], links: dict[str, int]) -> None:
        self._r, self.name, self._msgs, self._links = reader, name, msgs, links
        self._attrs: dict[str, Any] | None = None

    @property
    def attrs(self) -> dict[str, Any]:
        if self._attrs is None:
            with _guard(self.name):
                self._attrs = self._r.attributes(self._msgs)
        return self._attrs

    def keys(self):
        return self._links.keys()

    def __iter__(self) -> Iterator[str]:
        return iter(self._links)

    def __len__(self) -> int:
        return len(self._links)

    def __contains__(self, name: object) -> bool:
        try:
            self[str(name)]
        except KeyError:
            return False
        return True

    def __getitem__(self, name: str) -> "Dataset | Group":
        node: Dataset | Group = self
        for part in [p for p in name.split("/") if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError(f"Unable to open object (object {part!r} doesn't exist)")
            path = f"{node.name.rstrip('/')}/{part}"
            with _guard(path):
                msgs = self._r.messages(node._links[part])
                links = self._r.links(msgs)
            node = Group(self._r, path, msgs, links) if links is not None else Dataset(self._r, path, msgs)
        return node


class File:
    """``h5py.File`` look-alike for STAMP feature files.  Modes: "r" (read) and "w" (create / truncate; the file is
    laid out and written when it is closed).  Accepts a path or an open binary file object (the reference hands
    h5py a ``NamedTemporaryFile``, src/stamp/preprocessing/__init__.py:343-346)."""

    def __init__(self, name: str | os.PathLike | BinaryIO, mode: str = "r", **_ignored: Any) -> None:
        if mode not in ("r", "w"):
            raise ValueError(f"mode {mode!r}: only 'r' and 'w' are supported")
        self.mode = mode
        self._owns = not hasattr(name, "read") and not hasattr(name, "write")
        self.filename = os.fspath(name) if self._owns else getattr(name, "name", "<file object>")
        self._fp: BinaryIO | None = open(name, "rb" if mode == "r" else "wb") if self._owns else name  # type: ignore[arg-type]
        self._root: Group | None = None
        self._pending: dict[str, _PendingDataset] = {}
        self._wattrs = _AttrsWriter()
        if mode == "r":
            try:
                with _guard(self.filename):
                    reader = _Reader(self._fp)
                    msgs = reader.messages(reader.root_ohdr)
                    links = reader.links(msgs)
                if links is None:
                    raise H5Error("root object is not a group")
                self._root = Group(reader, "/", msgs, links)
            except Exception:
                self.close()
                raise

    # ---- shared
    def __enter__(self) -> "File":
        return self

    def __exit__(self, exc_type, *_exc) -> None:  # noqa: ANN001
        self.close(_discard=exc_type is not None)

    def close(self, _discard: bool = False) -> None:
        fp, self._fp = self._fp, None
        if fp is None:
            return
        try:
            if self.mode == "w" and not _discard:
                fp.seek(0)
                fp.truncate()
                _write_file(fp, self._pending, self._wattrs)
                fp.flush()
        finally:
            if self._owns:
                fp.close()

    @property
    def attrs(self):
        return self._wattrs if self.mode == "w" else self._group().attrs

    def _group(self) -> Group:
        if self._fp is None or self._root is None:
            raise ValueError("file is closed or not open for reading")
        return self._root

    # ---- read side
    def keys(self):
        return self._pending.keys() if self.mode == "w" else self._group().keys()

    def __iter__(self) -> Iterator[str]:
        return iter(self.keys())

    def __len__(self) -> int:
        return len(self.keys())

    def __contains__(self, name: object) -> bool:
        return name in self._pending if self.mode == "w" else name in self._group()

    def __getitem__(self, name: str):
        return self._pending[name] if self.mode == "w" else self._group()[name]

    # ---- write side
    def __setitem__(self, name: str, value: Any) -> None:
        self.create_dataset(name, data=value)

    def create_dataset(self, name: str, data: Any = None, shape=None, dtype=None) -> _PendingDataset:  # noqa: ANN001
        if self.mode != "w" or self._fp is None:
            raise ValueError("file is not open for writing")
        if name in self._pending:
            raise ValueError(f"Unable to create dataset (name {name!r} already exists)")
        if hasattr(data, "detach"):  # torch tensor
            data = data.detach().cpu().numpy()
        if data is None:
            data = np.zeros(shape, dtype=dtype)
        arr = np.asarray(data, dtype=dtype)
        if arr.dtype.kind not in "iufS":
            raise H5Error(f"dataset {name!r}: dtype {arr.dtype} is not supported")
        arr = np.ascontiguousarray(arr.astype(arr.dtype.newbyteorder("<"), copy=False)).reshape(arr.shape)
        self._pending[name] = _PendingDataset(arr)
        return self._pending[name]
