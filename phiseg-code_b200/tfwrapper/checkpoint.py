"""TensorFlow-1 checkpoint bundles (`model.ckpt-N.index` + `model.ckpt-N.data-00000-of-00001`) without TensorFlow.

The reference saves and restores its weights with tf.train.Saver (phiseg_model.py:144-148,505-525) and reads them back
with pywrap_tensorflow.NewCheckpointReader (tfwrapper/utils.py:182-187).  TensorFlow cannot be installed here, so this is
a restatement of the published on-disk format (tensorflow/core/util/tensor_bundle: an SSTable "index" in the LevelDB table
format whose values are BundleHeaderProto / BundleEntryProto messages, plus raw little-endian tensor bytes in the data
shards): `read_bundle(prefix)` -> {variable name: numpy array} lets trained PHiSeg weights be loaded into the engine
(phiseg.load_weights), `write_bundle(prefix, tensors)` writes the same format (masked CRC32C checksums included).
No TensorFlow build was available to cross-check the files: the writer / reader pair is tested against each other and
against hand-assembled blocks (tests/test_tf_checkpoint.py)."""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_}
DTYPE_IDS = {np.dtype(v): k for k, v in DTYPES.items()}


# ---- CRC32C (Castagnoli), masked as LevelDB / TensorFlow store it -------------------------------------------------
def _crc_table():
    t = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t.append(c)
    return np.array(t, dtype=np.uint32)


_TABLE = _crc_table()


def crc32c(data, crc=0):
    """CRC32C of bytes.  Uses the C helper of libphiseg_sm100.so when the library is built (tensor payloads are tens of MB),
    a table-driven Python loop otherwise."""
    data = bytes(data)
    try:
        from .. import lib as L
        import ctypes
        h = L.load()
        h.phs_crc32c.restype = ctypes.c_uint32
        h.phs_crc32c.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint32]
        return int(h.phs_crc32c(data, len(data), crc))
    except Exception:      # noqa: BLE001 - library not built: slow path
        c = crc ^ 0xFFFFFFFF
        tab = _TABLE
        for b in data:
            c = int(tab[(c ^ b) & 0xFF]) ^ (c >> 8)
        return c ^ 0xFFFFFFFF


def mask_crc(c):
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF


# ---- varints / minimal protobuf -----------------------------------------------------------------------------------
def _get_varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


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


def _parse_proto(buf):
    """{field number: [values]}; varint -> int, length-delimited -> bytes, fixed32/64 -> int"""
    out, pos = {}, 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        out.setdefault(field, []).append(v)
    return out


def _field(num, wt, payload):
    return _put_varint((num << 3) | wt) + payload


# ---- LevelDB table format ---------------------------------------------------------------------------------------------
def _snappy_decompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], 'little')
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = struct.unpack_from('<H', buf, pos)[0]
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        for _ in range(ln):
            out.append(out[-off])
    if len(out) != n:
        raise ValueError('snappy: length mismatch')
    return bytes(out)


def _read_block(f, offset, size, verify=True):
    f.seek(offset)
    raw = f.read(size + 5)
    body, ctype, crc = raw[:size], raw[size], struct.unpack_from('<I', raw, size + 1)[0]
    if verify and mask_crc(crc32c(raw[:size + 1])) != crc:
        raise ValueError('checkpoint index: block checksum mismatch at offset %d' % offset)
    if ctype == 1:
        body = _snappy_decompress(body)
    elif ctype != 0:
        raise ValueError('checkpoint index: unknown block compression %d' % ctype)
    return body


def _block_entries(block):
    nrestarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = _get_varint(block, pos)
        nons, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + nons])
        pos += nons
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _handle(buf, pos=0):
    off, pos = _get_varint(buf, pos)
    size, pos = _get_varint(buf, pos)
    return off, size, pos


def read_index(path, verify=True):
    """[(key bytes, value bytes)] of an SSTable file"""
    with open(path, 'rb') as f:
        f.seek(0, os.SEEK_END)
        n = f.tell()
        if n < 48:
            raise ValueError('%s is too short to be a checkpoint index' % path)
        f.seek(n - 48)
        footer = f.read(48)
        if struct.unpack_from('<Q', footer, 40)[0] != TABLE_MAGIC:
            raise ValueError('%s is not a TensorFlow checkpoint index (bad table magic)' % path)
        _, _, pos = _handle(footer, 0)                 # metaindex block (unused)
        ioff, isize, _ = _handle(footer, pos)
        out = []
        for _, hv in _block_entries(_read_block(f, ioff, isize, verify)):
            boff, bsize, _ = _handle(hv)
            out.extend(_block_entries(_read_block(f, boff, bsize, verify)))
        return out


def read_bundle(prefix, verify=True):
    """{variable name: numpy array} of the bundle `prefix`.index / `prefix`.data-*-of-* (tfwrapper/utils.py:182-187)."""
    entries = read_index(prefix + '.index', verify)
    header = None
    tensors = {}
    files = {}
    try:
        for key, val in entries:
            msg = _parse_proto(val)
            if key == b'':
                header = msg
                if msg.get(2, [0])[0] != 0:
                    raise ValueError('big-endian checkpoints are not supported')
                continue
            dtype = DTYPES.get(msg.get(1, [0])[0])
            if dtype is None:
                raise ValueError('variable %s has unsupported dtype %s' % (key.decode(), msg.get(1)))
            if 7 in msg:
                raise ValueError('variable %s is stored in slices (partitioned variables are not supported)' % key.decode())
            shape = []
            if 2 in msg:
                for d in _parse_proto(msg[2][0]).get(2, []):
                    shape.append(_parse_proto(d).get(1, [0])[0])
            shard, off, size = msg.get(3, [0])[0], msg.get(4, [0])[0], msg.get(5, [0])[0]
            nsh = header.get(1, [1])[0] if header else 1
            if shard not in files:
                files[shard] = open('%s.data-%05d-of-%05d' % (prefix, shard, nsh), 'rb')
            files[shard].seek(off)
            raw = files[shard].read(size)
            if verify and 6 in msg and mask_crc(crc32c(raw)) != msg[6][0]:
                raise ValueError('variable %s: data checksum mismatch' % key.decode())
            tensors[key.decode()] = np.frombuffer(raw, dtype=dtype).reshape(shape).copy()
    finally:
        for fh in files.values():
            fh.close()
    if header is None:
        raise ValueError('%s.index has no bundle header' % prefix)
    return tensors


# ---- writer ---------------------------------------------------------------------------------------------------------
def _build_block(items, restart_interval=16):
    out, restarts, last = bytearray(), [], b''
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def _emit_block(f, block):
    off = f.tell()
    f.write(block)
    f.write(b'\x00')
    f.write(struct.pack('<I', mask_crc(crc32c(block + b'\x00'))))
    return off, len(block)


def write_bundle(prefix, tensors):
    """Writes {name: array} as a one-shard V2 bundle readable by read_bundle (and laid out as tf.train.Saver does)."""
    names = sorted(tensors)
    entries = [(b'', _field(1, 0, _put_varint(1)) + _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1))))]
    with open(prefix + '.data-00000-of-00001', 'wb') as df:
        for n in names:
            a = np.asarray(tensors[n])
            if not a.flags.c_contiguous:
                a = a.copy()            # (np.ascontiguousarray would turn a scalar into a 1-d array)
            if a.dtype not in DTYPE_IDS:
                raise ValueError('variable %s: dtype %s cannot be stored' % (n, a.dtype))
            raw = a.tobytes()
            off = df.tell()
            df.write(raw)
            dims = b''.join(_field(2, 2, (lambda m: _put_varint(len(m)) + m)(_field(1, 0, _put_varint(int(d))))) for d in a.shape)
            msg = _field(1, 0, _put_varint(DTYPE_IDS[a.dtype])) + _field(2, 2, _put_varint(len(dims)) + dims)
            msg += _field(4, 0, _put_varint(off)) + _field(5, 0, _put_varint(len(raw)))
            msg += _field(6, 5, struct.pack('<I', mask_crc(crc32c(raw))))
            entries.append((n.encode(), msg))
    with open(prefix + '.index', 'wb') as f:
        index_items = []
        for i in range(0, len(entries), 64):
            chunk = entries[i:i + 64]
            off, size = _emit_block(f, _build_block(chunk))
            index_items.append((chunk[-1][0] + b'\x00', _put_varint(off) + _put_varint(size)))
        moff, msize = _emit_block(f, _build_block([]))
        ioff, isize = _emit_block(f, _build_block(index_items, restart_interval=1))
        footer = _put_varint(moff) + _put_varint(msize) + _put_varint(ioff) + _put_varint(isize)
        f.write(footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC))
    return prefix
