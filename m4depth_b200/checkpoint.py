"""Pure-Python reader (and a minimal writer) for the TensorFlow tensor-bundle checkpoints the reference ships and saves.

The reference stores weights with ``model.save_weights`` / ``tf.train.Checkpoint`` (callbacks.py:119-129): a pair of files
``<prefix>.index`` + ``<prefix>.data-00000-of-00001``.  The index is a leveldb *table* (sorted string table) whose values
are ``BundleEntryProto`` messages {dtype, shape, shard_id, offset, size, crc32c}; the data file is the raw little-endian
tensor bytes at those offsets.  TensorFlow is not needed to read either, so the shipped
Mid-Air / KITTI weights (``pretrained_weights.zip``) load straight into ``M4Depth.load_weights``:

    w = load_reference_weights("pretrained_weights.zip", "midair")      # or a ".../cp-0071.ckpt" prefix on disk
    model.load_weights(w)

Keys: the object-graph names of the checkpoint (``encoder/conv_layers_s1/0/kernel/.ATTRIBUTES/VARIABLE_VALUE``) map 1:1
onto this package's names by dropping the ``/.ATTRIBUTES/VARIABLE_VALUE`` suffix; optimizer slots and bookkeeping entries
are skipped.

Format notes (leveldb ``table_format.md``, TF ``tensor_bundle.proto``): blocks are ``contents | type(1) | crc32c(4)``;
the bundle writer never compresses (type 0); block entries are prefix-compressed ``shared | non_shared | value_len`` varints
with a restart array at the end; the 48-byte footer holds the metaindex and index block handles and the magic number.
"""
import io
import os
import struct
import zipfile

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"), 6: np.dtype("i1"),
           9: np.dtype("<i8"), 10: np.dtype("?"), 17: np.dtype("<u2"), 19: np.dtype("<f2"), 22: np.dtype("<u4"), 23: np.dtype("<u8")}
_DTYPE_CODES = {v: k for k, v in _DTYPES.items()}


class CheckpointError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------ crc32c (Castagnoli)
def _make_crc_table():
    tbl = np.zeros(256, dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tbl[i] = c
    return tbl


_CRC_TABLE = _make_crc_table()


def crc32c(data):
    """CRC-32C of a bytes-like object (table driven; a few MB/s - only used for verification on request and by the writer)."""
    tbl = _CRC_TABLE.tolist()
    c = 0xFFFFFFFF
    for b in bytes(data):
        c = tbl[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _mask_crc(c):
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ varints / protobuf
def _varint(buf, pos):
    r, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise CheckpointError("truncated varint")
        b = buf[pos]
        pos += 1
        r |= (b & 0x7F) << shift
        if not b & 0x80:
            return r, pos
        shift += 7
        if shift > 63:
            raise CheckpointError("varint too long")


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _proto_fields(buf):
    """Yield (field number, wire type, value) of one protobuf message; value is an int or a bytes slice."""
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise CheckpointError(f"unsupported protobuf wire type {wt}")
        yield field, wt, v


def _parse_entry(buf):
    """BundleEntryProto -> dict(dtype, shape, shard_id, offset, size, crc32c)."""
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    for field, _, v in _proto_fields(buf):
        if field == 1:
            e["dtype"] = v
        elif field == 2:                                   # TensorShapeProto: repeated Dim dim = 2 {int64 size = 1}
            for f2, _, dim in _proto_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, dv in _proto_fields(dim):
                        if f3 == 1:
                            size = dv
                    e["shape"].append(size)
        elif field == 3:
            e["shard_id"] = v
        elif field == 4:
            e["offset"] = v
        elif field == 5:
            e["size"] = v
        elif field == 6:
            e["crc32c"] = v
        elif field == 7:
            e["sliced"] = True
    return e


# ------------------------------------------------------------------------------------------------ leveldb table
def _read_block(index_bytes, offset, size, verify):
    end = offset + size
    if end + 5 > len(index_bytes):
        raise CheckpointError("block handle points past the end of the index file")
    contents, ctype = index_bytes[offset:end], index_bytes[end]
    if ctype != 0:
        raise CheckpointError("compressed table block (type %d): tensor bundles are written uncompressed" % ctype)
    if verify:
        stored = struct.unpack_from("<I", index_bytes, end + 1)[0]
        if _mask_crc(crc32c(index_bytes[offset:end + 1])) != stored:
            raise CheckpointError("table block checksum mismatch")
    return contents


def _block_entries(block):
    if len(block) < 4:
        raise CheckpointError("table block too small")
    nrestarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * nrestarts
    if limit < 0:
        raise CheckpointError("bad restart array")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > limit:
            raise CheckpointError("corrupt table entry")
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_index(index_bytes, verify=False):
    """``.index`` file contents -> (header dict, {tensor name: entry dict})."""
    if len(index_bytes) < 48:
        raise CheckpointError("index file shorter than a table footer")
    footer = index_bytes[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != _MAGIC:
        raise CheckpointError("not a leveldb table (bad magic number)")
    pos = 0
    _, pos = _varint(footer, pos)          # metaindex handle (unused)
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)
    isize, pos = _varint(footer, pos)
    entries, header = {}, {}
    for _, handle in _block_entries(_read_block(index_bytes, ioff, isize, verify)):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        for key, value in _block_entries(_read_block(index_bytes, boff, bsize, verify)):
            if key == b"":
                for field, _, v in _proto_fields(value):   # BundleHeaderProto: num_shards = 1, endianness = 2
                    if field == 1:
                        header["num_shards"] = v
                    elif field == 2:
                        header["endianness"] = v
                continue
            entries[key.decode("utf-8")] = _parse_entry(value)
    if header.get("endianness", 0) != 0:
        raise CheckpointError("big-endian bundles are not supported")
    return header, entries


class _Files:
    """Byte access to ``<prefix>.index`` / ``<prefix>.data-*`` on disk or inside a zip archive."""

    def __init__(self, source, prefix):
        self.zip = zipfile.ZipFile(source) if (isinstance(source, (str, os.PathLike)) and zipfile.is_zipfile(source)) else None
        self.prefix = prefix if self.zip else (prefix or source)

    def read(self, suffix):
        name = self.prefix + suffix
        if self.zip:
            try:
                return self.zip.read(name)
            except KeyError:
                raise CheckpointError(f"{name} not found in the archive")
        if not os.path.exists(name):
            raise CheckpointError(f"{name} not found")
        with open(name, "rb") as f:
            return f.read()

    def names(self):
        return self.zip.namelist() if self.zip else []


def read_bundle(source, prefix=None, verify=False, keep=None):
    """Read a tensor bundle -> {name: numpy array}.

    ``source``: a checkpoint prefix on disk (``.../cp-0071.ckpt``) or a zip archive, in which case ``prefix`` names the
    checkpoint inside it.  ``keep(name) -> bool`` filters entries before their bytes are touched; ``verify`` checks the
    table-block and per-tensor CRC-32C values (slow in pure Python: ~1 s per MB)."""
    files = _Files(source, prefix)
    header, entries = read_index(files.read(".index"), verify)
    nshards = header.get("num_shards", 1)
    shards = {}
    out = {}
    for name, e in entries.items():
        if keep is not None and not keep(name):
            continue
        if e["sliced"]:
            raise CheckpointError(f"{name}: partitioned (sliced) variables are not supported")
        if e["dtype"] not in _DTYPES:
            continue                                       # strings / variants (the object graph proto, save counters' names)
        dt = _DTYPES[e["dtype"]]
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = files.read(".data-%05d-of-%05d" % (sid, nshards))
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        n = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if len(raw) != e["size"] or n * dt.itemsize != e["size"]:
            raise CheckpointError(f"{name}: size {e['size']} does not match shape {e['shape']} of {dt}")
        if verify and e["crc32c"] is not None and _mask_crc(crc32c(raw)) != e["crc32c"]:
            raise CheckpointError(f"{name}: tensor checksum mismatch")
        out[name] = np.frombuffer(raw, dtype=dt).reshape(e["shape"]).copy()
    return out


def find_checkpoints(zip_path):
    """Checkpoint prefixes inside an archive (every ``*.index`` member)."""
    with zipfile.ZipFile(zip_path) as z:
        return sorted(n[:-len(".index")] for n in z.namelist() if n.endswith(".index"))


def to_model_weights(bundle):
    """Object-graph checkpoint names -> the names ``M4Depth.load_weights`` expects (model variables only)."""
    w = {}
    for name, arr in bundle.items():
        if not name.endswith(_SUFFIX):
            continue
        key = name[:-len(_SUFFIX)]
        if "/.OPTIMIZER_SLOT" in key or key.startswith(("optimizer", "save_counter", "_CHECKPOINTABLE")):
            continue
        if key.startswith(("encoder/", "d_estimator/")) and arr.dtype == np.float32:
            w[key] = arr
    return w


def load_reference_weights(source, which=None):
    """Weights of a reference checkpoint as {name: float32 numpy array}.

    ``source`` is a checkpoint prefix or a zip such as the reference's ``pretrained_weights.zip``; ``which`` selects the
    checkpoint inside the archive by substring ("midair", "kitti")."""
    prefix = None
    if isinstance(source, (str, os.PathLike)) and os.path.isfile(source) and zipfile.is_zipfile(source):
        cands = [p for p in find_checkpoints(source) if which is None or which in p]
        if len(cands) != 1:
            raise CheckpointError(f"{len(cands)} checkpoints match {which!r} in {source}: {cands}")
        prefix = cands[0]
    bundle = read_bundle(source, prefix, keep=lambda n: n.endswith(_SUFFIX) and "/.OPTIMIZER_SLOT" not in n)
    w = to_model_weights(bundle)
    if not w:
        raise CheckpointError("no model variables found in the checkpoint")
    return w


# ------------------------------------------------------------------------------------------------ minimal writer
def write_bundle(prefix, tensors):
    """Write {name: numpy array} as a single-shard tensor bundle that TensorFlow's (and this module's) reader accepts:
    one data block, one index block, uncompressed, with all checksums.  Used to save weights in the reference's format
    and by the tests (the shipped checkpoints cannot travel with the repository)."""
    names = sorted(tensors)
    data = io.BytesIO()
    items = [(b"", b"\x08\x01\x1a\x02\x08\x01")]          # BundleHeaderProto{num_shards: 1, version{producer: 1}}
    for name in names:
        arr = np.asarray(tensors[name], order="C")               # (ascontiguousarray would turn scalars into [1])
        dt = arr.dtype.newbyteorder("<") if arr.dtype.byteorder == ">" else arr.dtype
        if np.dtype(dt) not in _DTYPE_CODES:
            raise CheckpointError(f"{name}: dtype {arr.dtype} is not supported")
        raw = arr.astype(dt, copy=False).tobytes()
        shape = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in arr.shape))
        msg = b"\x08" + _put_varint(_DTYPE_CODES[np.dtype(dt)]) + b"\x12" + _put_varint(len(shape)) + shape
        if data.tell():
            msg += b"\x20" + _put_varint(data.tell())
        msg += b"\x28" + _put_varint(len(raw)) + b"\x35" + struct.pack("<I", _mask_crc(crc32c(raw)))
        items.append((name.encode("utf-8"), msg))
        data.write(raw)

    def block(entries):
        body = b"".join(b"\x00" + _put_varint(len(k)) + _put_varint(len(v)) + k + v for k, v in entries)
        restarts, pos = [], 0
        for k, v in entries:                                # every entry is a restart point (no prefix sharing)
            restarts.append(pos)
            pos += 1 + len(_put_varint(len(k))) + len(_put_varint(len(v))) + len(k) + len(v)
        body += b"".join(struct.pack("<I", r) for r in (restarts or [0])) + struct.pack("<I", max(len(restarts), 1))
        return body

    def framed(body):
        return body + b"\x00" + struct.pack("<I", _mask_crc(crc32c(body + b"\x00")))

    out = io.BytesIO()
    data_block = block(items)
    out.write(framed(data_block))
    meta_off = out.tell()
    meta_block = block([])
    out.write(framed(meta_block))
    index_off = out.tell()
    index_block = block([(items[-1][0] + b"\x00", _put_varint(0) + _put_varint(len(data_block)))])
    out.write(framed(index_block))
    footer = _put_varint(meta_off) + _put_varint(len(meta_block)) + _put_varint(index_off) + _put_varint(len(index_block))
    out.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC))
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + ".index", "wb") as f:
        f.write(out.getvalue())
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(data.getvalue())


def save_reference_weights(prefix, weights):
    """Inverse of ``load_reference_weights``: {package name: array} -> a bundle with the reference's object-graph names."""
    write_bundle(prefix, {k + _SUFFIX: np.asarray(v.detach().cpu() if hasattr(v, "detach") else v, dtype=np.float32)
                          for k, v in weights.items()})
