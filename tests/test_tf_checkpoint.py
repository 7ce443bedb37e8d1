"""TensorFlow-1 checkpoint bundles without TensorFlow (phiseg_model.py:144-148,505-525,
tfwrapper/utils.py:182-210): the LevelDB-table / protobuf reader against hand-assembled bytes, the writer / reader pair
against each other (checksums verified), corruption detection, and - with the engine - loading a bundle into the flat
parameter buffer by the reference's variable names."""
import importlib
import os
import struct

import numpy as np
import pytest


@pytest.fixture(scope='module')
def ck(pkg):
    return importlib.import_module('phiseg_code_b200.tfwrapper.checkpoint')


def test_crc32c_and_varints(ck):
    # published CRC32C check values (RFC 3720 appendix B.4): 32 zero bytes, 32 0xFF bytes, and "123456789"
    assert ck.crc32c(b'\x00' * 32) == 0x8A9136AA
    assert ck.crc32c(b'\xff' * 32) == 0x62A8AB43
    assert ck.crc32c(b'123456789') == 0xE3069283
    assert ck.crc32c(b'6789', ck.crc32c(b'12345')) == 0xE3069283          # incremental
    for v in (0, 1, 127, 128, 300, 2 ** 35 + 7):
        assert ck._get_varint(ck._put_varint(v), 0) == (v, len(ck._put_varint(v)))
    # hand-assembled protobuf: field 1 varint 150, field 2 bytes "ab", field 6 fixed32 0x01020304
    msg = bytes([0x08, 0x96, 0x01, 0x12, 0x02, 0x61, 0x62, 0x35, 0x04, 0x03, 0x02, 0x01])
    assert ck._parse_proto(msg) == {1: [150], 2: [b'ab'], 6: [0x01020304]}


def test_block_prefix_compression_by_hand(ck):
    # LevelDB block: entries (shared, non_shared, value_len, key delta, value), restart array, restart count
    blk = bytes([0, 5, 1]) + b'apple' + b'1' + bytes([3, 3, 1]) + b'ric' + b'2' + struct.pack('<II', 0, 1)
    assert list(ck._block_entries(blk)) == [(b'apple', b'1'), (b'appric', b'2')]
    assert list(ck._block_entries(ck._build_block([(b'apple', b'1'), (b'appric', b'2')]))) == [(b'apple', b'1'), (b'appric', b'2')]
    # snappy: literal "abcd" then a copy of 4 bytes at offset 4  -> "abcdabcd"
    assert ck._snappy_decompress(bytes([8, 0x0C]) + b'abcd' + bytes([0x01, 0x04])) == b'abcdabcd'


def test_bundle_round_trip_and_corruption(ck, tmp_path):
    rng = np.random.default_rng(0)
    tensors = {'posterior/z0_pre_1/W': rng.standard_normal((3, 3, 3, 32)).astype(np.float32),
               'posterior/z0_pre_1/batch_norm/BatchNorm/moving_variance': np.ones(32, np.float32),
               'global_step': np.asarray(12000, np.int64)}
    for i in range(150):                       # several data blocks in the index
        tensors['likelihood/filler_%03d/W' % i] = rng.standard_normal((i % 5 + 1, 4)).astype(np.float32)
    prefix = str(tmp_path / 'model.ckpt-12000')
    ck.write_bundle(prefix, tensors)
    assert os.path.exists(prefix + '.index') and os.path.exists(prefix + '.data-00000-of-00001')
    back = ck.read_bundle(prefix)
    assert sorted(back) == sorted(tensors)
    for k, v in tensors.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape and np.array_equal(back[k], v), k
    # a flipped byte in the data shard is caught by the per-variable checksum, a truncated index by the table magic
    raw = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    raw[10] ^= 0xFF
    open(prefix + '.data-00000-of-00001', 'wb').write(raw)
    with pytest.raises(ValueError):
        ck.read_bundle(prefix)
    open(prefix + '.index', 'wb').write(open(prefix + '.index', 'rb').read()[:-9])
    with pytest.raises(ValueError):
        ck.read_bundle(prefix, verify=False)


def test_checkpoint_listing_sees_tf_bundles(pkg, tmp_path):
    pm = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model')
    for f in ('model.ckpt-500.npz', 'model.ckpt-900.index', 'model.ckpt-900.data-00000-of-00001', 'model.ckpt-900.meta'):
        (tmp_path / f).write_bytes(b'')
    assert pm._latest_checkpoint(str(tmp_path), 'model.ckpt').endswith('model.ckpt-900.index')


@pytest.mark.gpu
def test_engine_loads_a_tf_bundle(pkg, ck, tmp_path):
    pm = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model')
    ex = importlib.import_module('phiseg_code_b200.phiseg.experiments')
    exp = ex.load_experiment(ex.experiment_path('phiseg_7_5'))
    exp.image_size = (64, 64, 1)
    a = pm.phiseg(exp, mode='parity', use_cuda_graph=False, seed=3)
    w = a.get_weights()
    w['global_step'] = np.asarray(4321, np.int64)
    w['posterior/z0_pre_1/W/Adam'] = np.zeros_like(w['posterior/z0_pre_1/W'])      # optimizer slots of a real TF file: ignored
    ck.write_bundle(str(tmp_path / 'model.ckpt-4321'), w)
    b = pm.phiseg(exp, mode='parity', use_cuda_graph=False, seed=99)
    path = b.load_weights(str(tmp_path), 'latest')
    assert path.endswith('model.ckpt-4321.index') and b.params.step == 4321
    for k, v in a.get_weights().items():
        assert np.array_equal(b.get_weights()[k], v), k
