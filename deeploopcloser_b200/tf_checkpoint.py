"""Reader / writer for TensorFlow "tensor bundle" (V2) checkpoints - the files `tf.train.Saver.save` produces for the
reference (`SDAV._saver.save(sess, checkpoint_file, global_step)`, src/sdav/network/SDAV.py:228-240, 273-275;
`DA._saver`, src/sdav/network/DenoisingAutoencoderVariant.py:160-174) - so that weights trained with the reference go
into the B200 encoder unchanged, and weights trained here can be restored by the reference.

A checkpoint `prefix` is three kinds of file:
    <prefix>.index                  an SSTable (LevelDB table format): key "" -> BundleHeaderProto, key <variable name>
                                    -> BundleEntryProto {dtype, shape, shard_id, offset, size, crc32c}
    <prefix>.data-00000-of-00001    the raw little-endian tensor bytes, back to back
    checkpoint                      text proto naming the latest prefix (`tf.train.latest_checkpoint`)

Only what the reference's checkpoints need is implemented: uncompressed table blocks (TensorFlow writes the index
uncompressed), single-shard bundles, dense tensors of the dtypes below. TensorFlow itself is not required (it is not
installable next to this code); the format is restated from its public specification: `tensor_bundle.proto`,
`table_format.md`. Pure host code - nothing here touches the GPU.

Variable names of the reference graphs:
    SDAV (SDAV.py:188-217, creation order): layer l -> `Variable_{3l}` = encoder weights [in, out], `Variable_{3l+1}`
        = encoder biases [out], `Variable_{3l+2}` = decoder biases [in] (index 0 is spelled `Variable`); `global_step`.
    DA (DenoisingAutoencoderVariant.py:92-101): `encoder_variables/encoder_weights`, `encoder_variables/encoder_biases`,
        `decoder_variables/decoder_biases`, `global_step`.
"""
import os
import re
import struct

import numpy as np

_TABLE_MAGIC = 0xdb4775248b80fb57
_DT = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8"), 4: np.dtype("u1"),
       6: np.dtype("i1"), 19: np.dtype("<f2"), 10: np.dtype("bool")}                       # types.proto DataType
_DT_INV = {v: k for k, v in _DT.items()}


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


def crc32c(data, crc=0):
    crc ^= 0xFFFFFFFF
    tbl = _CRC_TABLE
    for b in bytes(data):
        crc = int(tbl[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ varints / protobuf
def _get_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """Wire-format fields of one message -> list of (field number, wire type, value)."""
    pos, out = 0, []
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((field, wt, v))
    return out


def _field(tag, wt, payload):
    return _put_varint((tag << 3) | wt) + payload


def _parse_entry(buf):
    """BundleEntryProto -> dict (tensor_bundle.proto: dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6)."""
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    for field, wt, v in _parse_proto(buf):
        if field == 1:
            e["dtype"] = v
        elif field == 2:                                    # TensorShapeProto: repeated Dim dim = 2 {int64 size = 1}
            for f2, _, dim in _parse_proto(v):
                if f2 == 2:
                    size = 0
                    for f3, _, dv in _parse_proto(dim):
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


def _encode_entry(dtype_code, shape, offset, size, crc):
    dims = b"".join(_field(2, 2, _put_varint(len(d)) + d) for d in (_field(1, 0, _put_varint(s)) for s in shape))
    out = _field(1, 0, _put_varint(dtype_code))
    out += _field(2, 2, _put_varint(len(dims)) + dims)
    if offset:
        out += _field(4, 0, _put_varint(offset))
    out += _field(5, 0, _put_varint(size))
    out += _field(6, 5, struct.pack("<I", crc))
    return out


# ------------------------------------------------------------------------------------------------ SSTable (.index)
def _read_block(buf, offset, size, verify_crc=True):
    """One table block -> list of (key, value). Trailer: 1 byte compression type + 4 bytes masked crc32c of the block
    contents and the type byte (table_format.md); keys are prefix-compressed against their predecessor (`shared`
    bytes) except at restart points."""
    if offset + size + 5 > len(buf):
        raise ValueError("table block [%d, +%d) runs past the end of the file" % (offset, size))
    ctype = buf[offset + size]
    if ctype != 0:
        raise ValueError("compressed table block (type %d): only uncompressed checkpoint indices are supported" % ctype)
    block = buf[offset:offset + size]
    if verify_crc:
        stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
        if masked_crc32c(bytes(block) + b"\x00") != stored:
            raise ValueError("table block at offset %d: crc32c mismatch (corrupt checkpoint index)" % offset)
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(block[pos:pos + vlen])))
        pos += vlen
    return out


def _read_table(path):
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != _TABLE_MAGIC:
        raise ValueError("%s is not a TensorFlow checkpoint index (bad table magic)" % path)
    footer = buf[len(buf) - 48:]
    pos = 0
    _, pos = _get_varint(footer, pos)           # metaindex handle
    _, pos = _get_varint(footer, pos)
    idx_off, pos = _get_varint(footer, pos)     # index handle
    idx_size, pos = _get_varint(footer, pos)
    entries = {}
    for _, handle in _read_block(buf, idx_off, idx_size):
        off, p = _get_varint(handle, 0)
        size, _ = _get_varint(handle, p)
        for k, v in _read_block(buf, off, size):
            entries[k] = v
    return entries


def _build_block(items, restart_interval=16):
    """Uncompressed block as LevelDB's BlockBuilder (and therefore TensorFlow's checkpoint index) lays it out: every
    key drops the prefix it shares with its predecessor, a restart point (full key) every `restart_interval` entries,
    then the restart offsets, their count, and the 5-byte trailer. restart_interval = 1 gives no prefix sharing."""
    body, restarts = bytearray(), []
    last, counter = b"", 0
    for k, v in items:
        shared = 0
        if counter < restart_interval and restarts:
            n = min(len(last), len(k))
            while shared < n and last[shared] == k[shared]:
                shared += 1
        else:
            restarts.append(len(body))
            counter = 0
        body += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        last = k
        counter += 1
    if not restarts:
        restarts = [0]
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    trailer = b"\x00" + struct.pack("<I", masked_crc32c(bytes(body) + b"\x00"))
    return bytes(body), trailer


def _write_table(path, items, block_size=4096, restart_interval=16):
    """Sorted (key, value) items -> table file: data blocks of about `block_size` bytes (prefix-compressed keys, restart
    interval 16 - LevelDB's defaults, which TensorFlow's table builder keeps), an empty metaindex block, the index
    block (one entry per data block: its last key -> block handle; restart interval 1 as in LevelDB) and the footer."""
    items = sorted(items)
    out = bytearray()
    index_items = []
    cur, cur_bytes = [], 0
    chunks = []
    for k, v in items:
        cur.append((k, v))
        cur_bytes += len(k) + len(v) + 3
        if cur_bytes >= block_size:
            chunks.append(cur)
            cur, cur_bytes = [], 0
    if cur or not chunks:
        chunks.append(cur)
    for chunk in chunks:
        data, tr = _build_block(chunk, restart_interval)
        index_items.append((chunk[-1][0] if chunk else b"", _put_varint(len(out)) + _put_varint(len(data))))
        out += data + tr
    meta_off = len(out)
    meta, tr = _build_block([])
    out += meta + tr
    index_off = len(out)
    index, tr = _build_block(index_items, 1)
    out += index + tr
    footer = _put_varint(meta_off) + _put_varint(len(meta)) + _put_varint(index_off) + _put_varint(len(index))
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _TABLE_MAGIC)
    out += footer
    with open(path, "wb") as f:
        f.write(bytes(out))


# ------------------------------------------------------------------------------------------------ public API
def latest_checkpoint(directory):
    """`tf.train.latest_checkpoint`: the prefix named by `<directory>/checkpoint`, or None."""
    state = os.path.join(directory, "checkpoint")
    if not os.path.exists(state):
        return None
    with open(state) as f:
        m = re.search(r'^model_checkpoint_path:\s*"([^"]*)"', f.read(), re.M)
    if not m:
        return None
    p = m.group(1)
    return p if os.path.isabs(p) else os.path.join(directory, p)


def list_variables(prefix):
    """[(name, shape, numpy dtype)] of a checkpoint prefix."""
    out = []
    for k, v in sorted(_read_table(prefix + ".index").items()):
        if k == b"":
            continue
        e = _parse_entry(v)
        out.append((k.decode(), tuple(e["shape"]), _DT.get(e["dtype"])))
    return out


def load_checkpoint(prefix, verify_crc=True):
    """{variable name: ndarray} of a checkpoint prefix; every tensor is checked against its stored crc32c."""
    entries = _read_table(prefix + ".index")
    header = _parse_proto(entries.get(b"", b""))
    num_shards = next((v for f, _, v in header if f == 1), 1)
    out = {}
    files = {}
    try:
        for k, v in sorted(entries.items()):
            if k == b"":
                continue
            e = _parse_entry(v)
            name = k.decode()
            if e["sliced"]:
                raise ValueError("%s: partitioned variables are not supported" % name)
            if e["dtype"] not in _DT:
                raise ValueError("%s: unsupported dtype code %d" % (name, e["dtype"]))
            shard = "%s.data-%05d-of-%05d" % (prefix, e["shard_id"], num_shards)
            if shard not in files:
                files[shard] = open(shard, "rb")
            f = files[shard]
            f.seek(e["offset"])
            raw = f.read(e["size"])
            if len(raw) != e["size"]:
                raise ValueError("%s: data file truncated" % name)
            if verify_crc and e["crc32c"] is not None and _fast_masked_crc(raw) != e["crc32c"]:
                raise ValueError("%s: crc32c mismatch" % name)
            out[name] = np.frombuffer(raw, dtype=_DT[e["dtype"]]).reshape(e["shape"]).copy()
    finally:
        for f in files.values():
            f.close()
    return out


def save_checkpoint(prefix, tensors, update_state=True, block_size=4096, restart_interval=16):
    """Write {name: ndarray} as a single-shard V2 checkpoint at `prefix` (+ the `checkpoint` state file). The index
    is laid out like TensorFlow's (prefix-compressed keys, restart interval 16; see _write_table)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = []
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(tensors):
            a = np.asarray(tensors[name])
            a = np.ascontiguousarray(a) if a.ndim else a      # ascontiguousarray would turn a scalar into shape (1,)
            dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
            a = a.astype(dt, copy=False)
            if np.dtype(dt) not in _DT_INV:
                raise ValueError("%s: unsupported dtype %s" % (name, a.dtype))
            raw = a.tobytes()
            f.write(raw)
            # the masked crc32c of the tensor bytes is mandatory: restore verifies it
            items.append((name.encode(), _encode_entry(_DT_INV[np.dtype(dt)], a.shape, offset, len(raw), _fast_masked_crc(raw))))
            offset += len(raw)
    header = _field(1, 0, _put_varint(1)) + _field(2, 0, _put_varint(0)) + _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1)))
    _write_table(prefix + ".index", [(b"", header)] + items, block_size, restart_interval)
    if update_state:
        directory = os.path.dirname(os.path.abspath(prefix))
        rel = os.path.basename(prefix)
        with open(os.path.join(directory, "checkpoint"), "w") as f:
            f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (rel, rel))
    return prefix


def _fast_masked_crc(raw):
    c = _crc32c_np(raw)
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF


def _crc32c_np(raw, lanes=1 << 14):
    """crc32c of a large buffer with NumPy: the buffer is cut into `lanes` equal segments whose CRC registers advance
    in lock-step (one vectorised table step per byte position), then the registers are chained with the linear
    'advance by one segment of zero bytes' operator (CRC is linear over GF(2))."""
    n = len(raw)
    if n < 64 * lanes:
        return crc32c(raw)
    seg = n // lanes
    main = seg * lanes
    tbl = _CRC_TABLE
    cols = np.ascontiguousarray(np.frombuffer(raw, dtype=np.uint8, count=main).reshape(lanes, seg).T)
    reg = np.zeros(lanes, dtype=np.uint32)
    reg[0] = 0xFFFFFFFF
    # the same recurrence applied to the 32 basis states gives the 'advance by `seg` zero bytes' operator
    basis = (np.uint32(1) << np.arange(32, dtype=np.uint32))
    for i in range(seg):
        reg = tbl[(reg ^ cols[i]) & np.uint32(0xFF)] ^ (reg >> np.uint32(8))
        basis = tbl[basis & np.uint32(0xFF)] ^ (basis >> np.uint32(8))
    op = []                                   # 4 byte-indexed tables of the operator
    for byte in range(4):
        t = np.zeros(256, dtype=np.uint32)
        for v in range(256):
            acc = 0
            for bit in range(8):
                if v >> bit & 1:
                    acc ^= int(basis[8 * byte + bit])
            t[v] = acc
        op.append(t.tolist())
    o0, o1, o2, o3 = op
    regs = reg.tolist()
    r = regs[0]
    for s in range(1, lanes):
        r = o0[r & 0xFF] ^ o1[(r >> 8) & 0xFF] ^ o2[(r >> 16) & 0xFF] ^ o3[r >> 24] ^ regs[s]
    r ^= 0xFFFFFFFF
    return crc32c(raw[main:], r) if main < n else r


# ------------------------------------------------------------------------------------------------ reference graphs
def sdav_variable_names(n_layers=5):
    """(weights, encoder biases, decoder biases) variable names of the SDAV graph, per layer."""
    def nm(i):
        return "Variable" if i == 0 else "Variable_%d" % i
    return [(nm(3 * l), nm(3 * l + 1), nm(3 * l + 2)) for l in range(n_layers)]


DA_VARIABLE_NAMES = ("encoder_variables/encoder_weights", "encoder_variables/encoder_biases",
                     "decoder_variables/decoder_biases")
