"""TensorFlow checkpoint reader/writer, CPU only.  No TF-written file exists
offline, so every primitive is pinned by a published known answer or by an independent
implementation (tensorboard's CRC-32C), and the file formats by round trips and by tables
assembled by hand in this test."""
import os
import struct

import numpy as np
import pytest

import oracle
from squeezedet_b200.utils import checkpoint as ckpt
from squeezedet_b200.utils import synth
from squeezedet_b200.utils import tf_checkpoint as tfc


def test_crc32c_known_answers():
  # RFC 3720 B.4 test vectors
  assert tfc.crc32c(b'123456789') == 0xE3069283
  assert tfc.crc32c(bytes(32)) == 0x8A9136AA
  assert tfc.crc32c(b'\xff' * 32) == 0x62A8AB43
  assert tfc.crc32c(bytes(range(32))) == 0x46DD794E
  # incremental == one shot
  assert tfc.crc32c(b'6789', tfc.crc32c(b'12345')) == 0xE3069283


def test_crc32c_and_mask_match_tensorboards_independent_implementation():
  tb = pytest.importorskip('tensorboard.compat.tensorflow_stub.pywrap_tensorflow')
  rng = np.random.default_rng(0)
  for n in (0, 1, 7, 64, 1000):
    data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    assert tfc.crc32c(data) == tb.crc32c(data)
    assert tfc.mask_crc(tfc.crc32c(data)) == tb.masked_crc32c(data)
    assert tfc.unmask_crc(tb.masked_crc32c(data)) == tb.crc32c(data)


def test_varint_and_proto_walk():
  for v in (0, 1, 127, 128, 300, 2 ** 32, 2 ** 63):
    enc = tfc.put_varint(v)
    assert tfc.get_varint(enc, 0) == (v, len(enc))
  assert tfc.put_varint(300) == b'\xac\x02'                      # protobuf docs' example
  # message {1: varint 150, 2: "testing", 6: fixed32 7}
  msg = b'\x08\x96\x01' + b'\x12\x07testing' + b'\x35' + struct.pack('<I', 7)
  assert list(tfc.parse_proto(msg)) == [(1, 0, 150), (2, 2, b'testing'), (6, 5, 7)]
  assert tfc.parse_shape(tfc.encode_shape((3, 3, 64, 16))) == (3, 3, 64, 16)
  assert tfc.parse_shape(b'') == ()
  with pytest.raises(tfc.CheckpointError):
    tfc.get_varint(b'\x80', 0)


def test_snappy_known_vectors():
  # literal only: length 5, tag (5-1)<<2
  assert tfc.snappy_decompress(b'\x05' + bytes([4 << 2]) + b'hello') == b'hello'
  # "abababab": literal "ab" + copy(len 6, offset 2) with a 1-byte-offset tag (overlapping)
  copy1 = bytes([((6 - 4) << 2) | 1, 2])
  assert tfc.snappy_decompress(b'\x08' + bytes([1 << 2]) + b'ab' + copy1) == b'abababab'
  # 2-byte-offset copy
  copy2 = bytes([((4 - 1) << 2) | 2, 4, 0])
  assert tfc.snappy_decompress(b'\x08' + bytes([3 << 2]) + b'wxyz' + copy2) == b'wxyzwxyz'
  # long literal (length byte follows the tag when len-1 >= 60)
  body = bytes(range(100))
  assert tfc.snappy_decompress(b'\x64' + bytes([60 << 2, 99]) + body) == body
  with pytest.raises(tfc.CheckpointError):
    tfc.snappy_decompress(b'\x05' + bytes([1 << 2]) + b'ab' + bytes([1, 9]))    # offset past start


def test_table_round_trip_and_layout(tmp_path):
  items = [(b'', b'header')] + [(('k%04d' % i).encode(), os.urandom(i % 97)) for i in range(500)]
  p = str(tmp_path / 't.index')
  tfc.write_table(p, items, block_size=512)
  assert tfc.read_table(p) == items
  raw = open(p, 'rb').read()
  assert raw[-8:] == bytes.fromhex('57fb808b247547db')           # LevelDB magic, little-endian
  assert len(raw) > 48
  # flip one payload byte: the block checksum must catch it
  bad = bytearray(raw)
  bad[10] ^= 0x40
  open(p, 'wb').write(bad)
  with pytest.raises(tfc.CheckpointError, match='checksum'):
    tfc.read_table(p)


def test_table_reader_accepts_snappy_blocks_and_prefix_compression(tmp_path):
  """A table assembled by hand the way LevelDB's builder would: shared-prefix keys and a
  snappy-typed (literal-only) data block."""
  entries = [(b'conv1/biases', b'B'), (b'conv1/kernels', b'K'), (b'conv12/biases', b'b')]
  block = tfc._build_block(entries, restart_interval=16)
  assert block.count(b'conv1/') == 1                             # prefix compression really used
  comp = tfc.put_varint(len(block)) + bytes([60 << 2, len(block) - 1]) + block   # one long literal
  out = bytearray()

  def emit(contents, btype):
    off = len(out)
    out.extend(contents)
    out.append(btype)
    out.extend(struct.pack('<I', tfc.mask_crc(tfc.crc32c(bytes(contents) + bytes([btype])))))
    return tfc.put_varint(off) + tfc.put_varint(len(contents))

  h_data = emit(comp, 1)
  h_meta = emit(tfc._build_block([]), 0)
  h_index = emit(tfc._build_block([(b'conv12/biases', h_data)], 1), 0)
  footer = h_meta + h_index
  out.extend(footer + b'\0' * (40 - len(footer)) + struct.pack('<Q', tfc.TABLE_MAGIC))
  p = str(tmp_path / 'snappy.tbl')
  open(p, 'wb').write(out)
  assert tfc.read_table(p) == entries


@pytest.mark.parametrize('net', ['squeezeDet', 'resnet50'])
def test_v2_round_trip_of_model_parameters(net, tmp_path):
  specs = oracle.param_specs(net)
  weights = synth.synthetic_weights(specs, seed=4)
  extra = dict(weights)
  extra['global_step'] = np.array(87000, dtype=np.int64)          # scalar, int64, like a real Saver file
  extra['iou'] = np.zeros((0,), np.float32)                       # empty tensor
  prefix = str(tmp_path / 'model.ckpt-87000')
  tfc.write_v2(prefix, extra)
  assert tfc.checkpoint_kind(prefix) == 'v2'
  back = tfc.read_checkpoint(prefix)
  assert sorted(back) == sorted(extra)
  for k, v in extra.items():
    assert back[k].dtype == np.asarray(v).dtype and back[k].shape == np.asarray(v).shape
    np.testing.assert_array_equal(back[k], v)
  # names= restricts what is read; load_weights_file picks the format from the path
  some = [n for n, _ in specs][:3]
  assert sorted(tfc.read_checkpoint(prefix, names=set(some))) == sorted(some)
  via = ckpt.load_weights_file(prefix)
  np.testing.assert_array_equal(via['conv1/kernels'], weights['conv1/kernels'])
  # corrupt one tensor byte in the data shard -> per-tensor checksum
  data = prefix + '.data-00000-of-00001'
  raw = bytearray(open(data, 'rb').read())
  raw[len(raw) // 2] ^= 1
  open(data, 'wb').write(raw)
  with pytest.raises(tfc.CheckpointError, match='checksum'):
    tfc.read_checkpoint(prefix)


def test_v2_index_bytes_follow_the_bundle_protos(tmp_path):
  prefix = str(tmp_path / 'm')
  tfc.write_v2(prefix, {'b': np.arange(6, dtype=np.float32).reshape(2, 3),
                        'a': np.array([1, 2], np.int32)})
  items = tfc.read_table(prefix + '.index')
  assert [k for k, _ in items] == [b'', b'a', b'b']               # header first, names sorted
  assert list(tfc.parse_proto(items[0][1])) == [(1, 0, 1), (3, 2, b'\x08\x01')]   # 1 shard, producer 1
  ent = dict((f, v) for f, _, v in tfc.parse_proto(items[2][1]))
  assert ent[1] == 1 and ent[4] == 8 and ent[5] == 24             # DT_FLOAT, after a's 8 bytes, 24 bytes
  assert tfc.parse_shape(ent[2]) == (2, 3)
  raw = open(prefix + '.data-00000-of-00001', 'rb').read()
  assert raw[8:] == np.arange(6, dtype='<f4').tobytes()
  assert tfc.unmask_crc(ent[6]) == tfc.crc32c(raw[8:])


def test_v1_reader_on_a_hand_built_file(tmp_path):
  """V1: SavedTensorSlices{data=2: SavedSlice{name=1, slice=2, data=3: TensorProto}} per key."""
  w = np.arange(24, dtype=np.float32).reshape(1, 1, 6, 4)
  b = np.array([0.5, -1.5, 2.0, 7.0], np.float32)

  def tensor_proto(a, packed_vals):
    t = tfc._field(1, 0, 1) + tfc._field(2, 2, tfc.encode_shape(a.shape))
    if packed_vals:
      return t + tfc._field(5, 2, a.astype('<f4').tobytes())      # float_val, packed
    return t + tfc._field(4, 2, a.astype('<f4').tobytes())        # tensor_content

  def slices(name, a, packed_vals):
    sl = tfc._field(1, 2, name.encode()) + tfc._field(2, 2, b'') + tfc._field(3, 2, tensor_proto(a, packed_vals))
    return tfc._field(2, 2, sl)

  items = [(b'', tfc._field(1, 2, b'')),
           (b'\x00conv1/biases', slices('conv1/biases', b, True)),
           (b'\x00conv1/kernels', slices('conv1/kernels', w, False))]
  p = str(tmp_path / 'model.ckpt-1')
  tfc.write_table(p, items)
  assert tfc.checkpoint_kind(p) == 'v1'
  got = tfc.read_checkpoint(p)
  np.testing.assert_array_equal(got['conv1/kernels'], w)
  np.testing.assert_array_equal(got['conv1/biases'], b)


def test_not_a_checkpoint(tmp_path):
  p = str(tmp_path / 'junk')
  open(p, 'wb').write(b'x' * 100)
  assert tfc.checkpoint_kind(p) is None
  with pytest.raises(tfc.CheckpointError):
    tfc.read_checkpoint(p)
  with pytest.raises(tfc.CheckpointError):
    tfc.read_checkpoint(str(tmp_path / 'missing'))


def test_load_weights_file_dispatch(tmp_path):
  w = {'conv1/kernels': np.ones((3, 3, 3, 4), np.float32), 'conv1/biases': np.zeros(4, np.float32)}
  npz = str(tmp_path / 'w.npz')
  ckpt.save_npz(npz, w)
  np.testing.assert_array_equal(ckpt.load_weights_file(npz)['conv1/kernels'], w['conv1/kernels'])
  np.testing.assert_array_equal(ckpt.load_weights_file(npz[:-4])['conv1/biases'], w['conv1/biases'])
  prefix = str(tmp_path / 'model.ckpt-5')
  ckpt.save_tf_checkpoint(prefix, w)
  assert sorted(ckpt.load_weights_file(prefix)) == sorted(w)
  assert sorted(ckpt.load_weights_file(prefix + '.index')) == sorted(w)
  with pytest.raises(FileNotFoundError):
    ckpt.load_weights_file(str(tmp_path / 'nothing-here'))
