"""Minimal HDF5 writer / reader for the `waveforms.h5` schema of generate-waveforms.

The reference writes its output with h5py (tqdne/generate_waveforms.py:176-193: five 1-D float64 feature datasets and
`waveforms (N, 3, 4064) float32` in the root group); its consumers open the file with h5py
(scripts/write_to_seisbench.py:64-68).  h5py / libhdf5 are not in this image, so this module writes the same file with
the oldest, simplest on-disk structures of the HDF5 File Format Specification (version 0 superblock, version 1 object
headers, a version 1 group B-tree with symbol-table nodes and a local heap, contiguous little-endian datasets), which
every HDF5 library release reads.  `read()` parses those structures back by following the addresses in the file (it
does not assume this writer's layout); it is what the tests use, since no libhdf5 is available here to cross-check:
files from this writer have NOT been opened with libhdf5 in this environment.

Supported: root-group datasets of dtype float32 / float64 / int32 / int64 / uint8, any rank, up to 256 datasets.
"""

from __future__ import annotations

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K, INTERNAL_K = 4, 16          # symbols per SNOD <= 2 * LEAF_K, children per TREE node <= 2 * INTERNAL_K
HEAP_FREE_NULL = 1                  # "no free block" marker of a local heap's free list

_DTYPES = {
    # numpy dtype -> datatype message body (class/version, class bit field, size, properties)
    "<f4": bytes([0x11, 0x20, 31, 0]) + struct.pack("<I", 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127),
    "<f8": bytes([0x11, 0x20, 63, 0]) + struct.pack("<I", 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023),
    "<i4": bytes([0x10, 0x08, 0, 0]) + struct.pack("<I", 4) + struct.pack("<HH", 0, 32),
    "<i8": bytes([0x10, 0x08, 0, 0]) + struct.pack("<I", 8) + struct.pack("<HH", 0, 64),
    "|u1": bytes([0x10, 0x00, 0, 0]) + struct.pack("<I", 1) + struct.pack("<HH", 0, 8),
}


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages: list[bytes]) -> bytes:
    data = b"".join(messages)
    # version 1 prefix: version, reserved, message count, reference count, header data size, 4 pad bytes
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(data)) + data


def _dataset_header(arr: np.ndarray, data_addr: int) -> bytes:
    key = arr.dtype.newbyteorder("<").str if arr.dtype.byteorder != "|" else arr.dtype.str
    if key not in _DTYPES:
        raise TypeError(f"hdf5_min: unsupported dtype {arr.dtype}")
    dataspace = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
    fill = struct.pack("<BBBB", 2, 2, 2, 0)   # version 2, late allocation, fill-time "if set", no fill value defined
    layout = struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes)   # version 3, contiguous storage
    return _object_header([_message(0x0001, dataspace), _message(0x0003, _DTYPES[key], 1), _message(0x0005, fill, 1),
                           _message(0x0008, layout)])


def write(path, datasets: dict) -> None:
    """Write {name: array} as contiguous datasets of the root group."""
    items = sorted(((str(k), np.ascontiguousarray(v)) for k, v in datasets.items()), key=lambda kv: kv[0].encode())
    if not items or len(items) > 2 * LEAF_K * 2 * INTERNAL_K:
        raise ValueError("hdf5_min: between 1 and 256 datasets")
    for name, _ in items:
        if "/" in name or not name:
            raise ValueError(f"hdf5_min: bad dataset name {name!r}")
    items = [(n, a.astype(a.dtype.newbyteorder("<")) if a.dtype.byteorder == ">" else a) for n, a in items]

    # local heap data segment: the empty root name at offset 0, then the link names, each padded to 8 bytes
    heap_data, name_off = bytearray(8), {}
    for name, _ in items:
        name_off[name] = len(heap_data)
        heap_data += _pad8(name.encode() + b"\0")
    nodes = [items[i:i + 2 * LEAF_K] for i in range(0, len(items), 2 * LEAF_K)]

    # ---- file layout (every structure 8-byte aligned)
    pos = 96                                              # after the superblock
    root_hdr_addr = pos; pos += 16 + 24                   # object header prefix + one symbol-table message
    btree_addr = pos; pos += 24 + 2 * INTERNAL_K * 16 + 8
    heap_addr = pos; pos += 32
    heap_data_addr = pos; pos += len(heap_data)
    snod_addr = []
    for _ in nodes:
        snod_addr.append(pos); pos += 8 + 2 * LEAF_K * 40
    hdr_addr, data_addr = {}, {}
    for name, arr in items:
        hdr_addr[name] = pos
        pos += len(_dataset_header(arr, 0))
    for name, arr in items:
        data_addr[name] = pos if arr.nbytes else UNDEF    # no storage is allocated for an empty dataset
        pos += arr.nbytes + (-arr.nbytes % 8)
    eof = pos

    out = bytearray()
    out += SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", LEAF_K, INTERNAL_K, 0)
    out += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    out += struct.pack("<QQII", 0, root_hdr_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)   # root symbol-table entry
    assert len(out) == 96
    out += _object_header([_message(0x0011, struct.pack("<QQ", btree_addr, heap_addr))])
    # group B-tree: one leaf-level node; key[i] < names of child i <= key[i + 1] (keys are heap offsets of names)
    tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(nodes), UNDEF, UNDEF) + struct.pack("<Q", 0)
    for node, addr in zip(nodes, snod_addr):
        tree += struct.pack("<QQ", addr, name_off[node[-1][0]])
    out += tree + b"\0" * (24 + 2 * INTERNAL_K * 16 + 8 - len(tree))
    out += b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), HEAP_FREE_NULL, heap_data_addr)
    out += heap_data
    for node in nodes:
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(node))
        for name, _ in node:
            snod += struct.pack("<QQII16x", name_off[name], hdr_addr[name], 0, 0)
        out += snod + b"\0" * (8 + 2 * LEAF_K * 40 - len(snod))
    for name, arr in items:
        assert len(out) == hdr_addr[name]
        out += _dataset_header(arr, data_addr[name])
    with open(path, "wb") as f:
        f.write(out)
        for name, arr in items:
            assert f.tell() == data_addr[name] or not arr.nbytes
            f.write(arr.tobytes())
            f.write(b"\0" * (-arr.nbytes % 8))
        assert f.tell() == eof


# ------------------------------------------------------------------------------------------------------------------
def _parse_datatype(body: bytes) -> np.dtype:
    cls, ver = body[0] & 0x0F, body[0] >> 4
    size = struct.unpack_from("<I", body, 4)[0]
    order = ">" if body[1] & 1 else "<"
    if ver not in (1, 2, 3):
        raise ValueError(f"hdf5_min: datatype version {ver}")
    if cls == 1:
        return np.dtype(f"{order}f{size}")
    if cls == 0:
        return np.dtype(f"{order}{'i' if body[1] & 0x08 else 'u'}{size}")
    raise TypeError(f"hdf5_min: unsupported datatype class {cls}")


def _read_object_header(buf, addr):
    ver, _, nmsg, _refs, size = struct.unpack_from("<BBHII", buf, addr)
    if ver != 1:
        raise ValueError(f"hdf5_min: object header version {ver} (only version 1 is read)")
    p, end, msgs = addr + 16, addr + 16 + size, []
    while p < end and len(msgs) < nmsg:
        mtype, msize, _flags = struct.unpack_from("<HHB", buf, p)
        msgs.append((mtype, bytes(buf[p + 8:p + 8 + msize])))
        p += 8 + msize
    return msgs


def read(path) -> dict:
    """Parse a file of the structure classes `write` emits (any layout / ordering of them) -> {name: array}."""
    buf = memoryview(open(path, "rb").read())
    if bytes(buf[:8]) != SIGNATURE:
        raise ValueError("hdf5_min: not an HDF5 file")
    if buf[8] != 0 or buf[13] != 8 or buf[14] != 8:
        raise ValueError("hdf5_min: only version 0 superblocks with 8-byte offsets / lengths are read")
    eof = struct.unpack_from("<Q", buf, 40)[0]
    if eof != len(buf):
        raise ValueError(f"hdf5_min: end-of-file address {eof} != file size {len(buf)}")
    root_hdr = struct.unpack_from("<Q", buf, 64)[0]
    stab = [b for t, b in _read_object_header(buf, root_hdr) if t == 0x0011]
    if not stab:
        raise ValueError("hdf5_min: root group has no symbol-table message")
    btree, heap = struct.unpack_from("<QQ", stab[0])
    if bytes(buf[heap:heap + 4]) != b"HEAP":
        raise ValueError("hdf5_min: bad local heap")
    heap_size, _free, heap_data = struct.unpack_from("<QQQ", buf, heap + 8)

    def name_at(off):
        s = bytes(buf[heap_data + off:heap_data + heap_size])
        return s[:s.index(b"\0")].decode()

    out = {}

    def walk(addr):
        if bytes(buf[addr:addr + 4]) == b"TREE":
            ntype, level, used = struct.unpack_from("<BBH", buf, addr + 4)
            assert ntype == 0, "not a group B-tree"
            for i in range(used):
                walk(struct.unpack_from("<Q", buf, addr + 24 + 8 + 16 * i)[0])
            return
        assert bytes(buf[addr:addr + 4]) == b"SNOD", "bad symbol-table node"
        nsym = struct.unpack_from("<H", buf, addr + 6)[0]
        for i in range(nsym):
            noff, ohdr = struct.unpack_from("<QQ", buf, addr + 8 + 40 * i)
            msgs = dict(_read_object_header(buf, ohdr))
            space, layout = msgs[0x0001], msgs[0x0008]
            rank = space[1]
            shape = struct.unpack_from(f"<{rank}Q", space, 8)
            dt = _parse_datatype(msgs[0x0003])
            lver, lclass, daddr, dsize = struct.unpack_from("<BBQQ", layout)
            assert lver == 3 and lclass == 1, "only contiguous version-3 layouts are read"
            assert dsize == int(np.prod(shape)) * dt.itemsize
            count = int(np.prod(shape))
            out[name_at(noff)] = (np.frombuffer(buf, dtype=dt, count=count, offset=daddr).reshape(shape).copy() if count
                                  else np.zeros(shape, dt))

    walk(btree)
    return out
