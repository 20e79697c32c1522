"""TF-free reader/writer for the TensorFlow TensorBundle checkpoint format.

The reference stores its weights as a SavedModel `variables/` bundle
(`variables.index` + `variables.data-00000-of-00001`), written by
`tf.saved_model.save` (reference bfcnn/export_model.py:136) and expected by
`bfcnn.load_model` under `<model>/saved_model/` (bfcnn/__init__.py:71, setup.py:59-73) or
`<model>/denoiser/` (export_model.py:117).

Format (verified on the reference's shipped unet_laplacian_v5.6 bundle):
`variables.index` is a LevelDB-style SSTable -- prefix-compressed key/value blocks, each
followed by a 1-byte compression tag and a masked CRC32C, an index block, and a 48-byte footer
ending in the magic 0xdb4775248b80fb57.  Key "" maps to a BundleHeaderProto; every other key
maps to a BundleEntryProto {1: dtype, 2: shape, 3: shard_id, 4: offset, 5: size, 6: crc32c}
locating raw little-endian tensor bytes in the data shard.
"""
from __future__ import annotations

import os
import re
import struct
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
DT_FLOAT = 1
DT_INT64 = 9   # tensorflow/core/framework/types.proto (the step / epoch counters of tf.train.Checkpoint)
_NP_OF_DT = {DT_FLOAT: "<f4", DT_INT64: "<i8"}
_KEY_RE = re.compile(r"variables/(\d+)/\.ATTRIBUTES/VARIABLE_VALUE$")

# ---------------------------------------------------------------------------- crc32c
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        poly = 0x82F63B78
        tbl = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ poly if c & 1 else c >> 1
            tbl[i] = c
        _CRC_TABLE = [int(x) for x in tbl]
    return _CRC_TABLE


def crc32c(data: bytes, crc: int = 0) -> int:
    tbl = _crc_table()
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c = tbl[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------- varints / protos
def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf: bytes) -> Dict[int, list]:
    """Minimal protobuf wire parser: {field: [values]} (varint -> int, len-delimited -> bytes,
    fixed32 -> int)."""
    out: Dict[int, list] = {}
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


def _parse_shape(buf: bytes) -> Tuple[int, ...]:
    dims = []
    for d in _parse_proto(buf).get(2, []):       # TensorShapeProto.dim
        dims.append(_parse_proto(d).get(1, [0])[0])  # Dim.size
    return tuple(int(x) for x in dims)


# ---------------------------------------------------------------------------- table reader
def _read_block(buf: bytes, offset: int, size: int, verify: bool) -> List[Tuple[bytes, bytes]]:
    raw = buf[offset:offset + size]
    trailer = buf[offset + size:offset + size + 5]
    if trailer[0] != 0:
        raise ValueError("compressed SSTable blocks are not supported (TensorBundle writes none)")
    if verify:
        want = struct.unpack("<I", trailer[1:5])[0]
        if masked_crc32c(raw + trailer[:1]) != want:
            raise ValueError("SSTable block checksum mismatch")
    num_restarts = struct.unpack_from("<I", raw, len(raw) - 4)[0]
    end = len(raw) - 4 - 4 * num_restarts
    out, pos, key = [], 0, b""
    while pos < end:
        shared, pos = _get_varint(raw, pos)
        unshared, pos = _get_varint(raw, pos)
        vlen, pos = _get_varint(raw, pos)
        key = key[:shared] + raw[pos:pos + unshared]
        pos += unshared
        out.append((key, raw[pos:pos + vlen]))
        pos += vlen
    return out


def read_index(index_path: str, verify: bool = True) -> Dict[str, dict]:
    buf = Path(index_path).read_bytes()
    if len(buf) < 48 or struct.unpack("<Q", buf[-8:])[0] != TABLE_MAGIC:
        raise ValueError(f"{index_path} is not a TensorBundle index (bad magic)")
    footer = buf[-48:]
    pos = 0
    _, pos = _get_varint(footer, pos)   # metaindex offset
    _, pos = _get_varint(footer, pos)   # metaindex size
    ioff, pos = _get_varint(footer, pos)
    isize, pos = _get_varint(footer, pos)
    entries: Dict[str, dict] = {}
    for _, handle in _read_block(buf, ioff, isize, verify):
        boff, p = _get_varint(handle, 0)
        bsize, p = _get_varint(handle, p)
        for k, v in _read_block(buf, boff, bsize, verify):
            pr = _parse_proto(v)
            if k == b"":
                entries[""] = {"num_shards": pr.get(1, [1])[0], "endianness": pr.get(2, [0])[0]}
                continue
            entries[k.decode("utf-8")] = {
                "dtype": pr.get(1, [0])[0],
                "shape": _parse_shape(pr[2][0]) if 2 in pr else (),
                "shard_id": pr.get(3, [0])[0],
                "offset": pr.get(4, [0])[0],
                "size": pr.get(5, [0])[0],
                "crc32c": pr.get(6, [None])[0],
            }
    return entries


def read_bundle(prefix: str, verify_crc: bool = True) -> Dict[str, np.ndarray]:
    """Read every float32 / int64 tensor of the bundle `<prefix>.index` / `<prefix>.data-*`."""
    entries = read_index(prefix + ".index")
    hdr = entries.pop("", {"num_shards": 1, "endianness": 0})
    if hdr.get("endianness", 0) != 0:
        raise ValueError("big-endian bundles are not supported")
    nsh = hdr.get("num_shards", 1)
    shards = {}
    out: Dict[str, np.ndarray] = {}
    for key, e in entries.items():
        if e["dtype"] not in _NP_OF_DT:
            continue
        npdt = np.dtype(_NP_OF_DT[e["dtype"]])
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = Path(f"{prefix}.data-{sid:05d}-of-{nsh:05d}").read_bytes()
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        n = int(np.prod(e["shape"])) if e["shape"] else 1
        if len(raw) != npdt.itemsize * n:
            raise ValueError(f"tensor {key}: {len(raw)} bytes for shape {e['shape']}")
        if verify_crc and e["crc32c"] is not None and masked_crc32c(raw) != e["crc32c"]:
            raise ValueError(f"tensor {key}: crc32c mismatch")
        out[key] = np.frombuffer(raw, dtype=npdt).reshape(e["shape"]).copy()
    return out


def read_model_variables(variables_dir: str, verify_crc: bool = True) -> List[np.ndarray]:
    """`hydra.variables` in integer order from a SavedModel `variables/` directory.

    Keys are `<root>/variables/<i>/.ATTRIBUTES/VARIABLE_VALUE` and sort lexicographically in
    the table ("10" < "2"), so they are re-sorted by the integer <i>."""
    tensors = read_bundle(os.path.join(variables_dir, "variables"), verify_crc)
    idx = []
    for k, v in tensors.items():
        m = _KEY_RE.search(k)
        if m:
            idx.append((int(m.group(1)), v))
    idx.sort(key=lambda t: t[0])
    if [i for i, _ in idx] != list(range(len(idx))):
        raise ValueError("variable indices in the bundle are not contiguous")
    return [v for _, v in idx]


# ---------------------------------------------------------------------------- writer
def _block(entries: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    out = bytearray()
    restarts = []
    prev = b""
    for i, (k, v) in enumerate(entries):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            for a, b in zip(prev, k):
                if a != b:
                    break
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v))
        out += k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _field_varint(field: int, v: int) -> bytes:
    return _put_varint(field << 3) + _put_varint(v)


def _field_bytes(field: int, b: bytes) -> bytes:
    return _put_varint((field << 3) | 2) + _put_varint(len(b)) + b


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray]) -> None:
    """Write float32 (and int64: step / epoch counters) tensors as a one-shard TensorBundle readable by
    `tf.train.load_checkpoint(prefix)` and by `read_bundle`."""
    keys = sorted(tensors.keys(), key=lambda s: s.encode("utf-8"))
    data = bytearray()
    entries: List[Tuple[bytes, bytes]] = []
    header = _field_varint(1, 1) + _field_varint(2, 0) + _field_bytes(3, _field_varint(1, 1))
    entries.append((b"", header))
    for k in keys:
        src = np.asarray(tensors[k])
        dt = DT_INT64 if src.dtype.kind in "iu" else DT_FLOAT
        a = np.ascontiguousarray(src, dtype=_NP_OF_DT[dt])
        raw = a.tobytes()
        shape = b"".join(_field_bytes(2, _field_varint(1, int(d))) for d in a.shape)
        e = _field_varint(1, dt) + _field_bytes(2, shape)
        if len(data):
            e += _field_varint(4, len(data))
        e += _field_varint(5, len(raw))
        e += _put_varint((6 << 3) | 5) + struct.pack("<I", masked_crc32c(raw))
        entries.append((k.encode("utf-8"), e))
        data += raw
    os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
    Path(prefix + ".data-00000-of-00001").write_bytes(bytes(data))

    def with_trailer(b: bytes) -> bytes:
        return b + b"\x00" + struct.pack("<I", masked_crc32c(b + b"\x00"))

    out = bytearray()
    dblock = _block(entries)
    d_off, d_size = 0, len(dblock)
    out += with_trailer(dblock)
    mblock = _block([])
    m_off, m_size = len(out), len(mblock)
    out += with_trailer(mblock)
    # index key: any key >= last data key
    iblock = _block([(entries[-1][0] + b"\x00", _put_varint(d_off) + _put_varint(d_size))], 1)
    i_off, i_size = len(out), len(iblock)
    out += with_trailer(iblock)
    footer = _put_varint(m_off) + _put_varint(m_size) + _put_varint(i_off) + _put_varint(i_size)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out += footer
    Path(prefix + ".index").write_bytes(bytes(out))


def write_model_variables(variables_dir: str, variables: List[np.ndarray],
                          root: str = "_model_hydra") -> None:
    """Inverse of `read_model_variables`."""
    tensors = {f"{root}/variables/{i}/.ATTRIBUTES/VARIABLE_VALUE": v for i, v in enumerate(variables)}
    write_bundle(os.path.join(variables_dir, "variables"), tensors)
