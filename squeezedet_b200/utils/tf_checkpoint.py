"""TensorFlow checkpoint files without TensorFlow.

The reference restores its weights with ``tf.train.Saver(model.model_params).restore(sess,
FLAGS.checkpoint)`` (reference src/demo.py:181-184, src/eval.py:165-169): the checkpoint maps
the variable names (`conv1/kernels`, `fire2/squeeze1x1/biases`, BN `.../gamma|beta|mean|var`,
src/nn_skeleton.py:425-438) to tensors.  This module reads those files so that the published
``model.ckpt-87000`` can be fed to the engine, and writes them so that weights produced here
can be restored by the reference's unmodified Saver.

Formats (restated from TensorFlow's published sources, tensorflow/core/util/tensor_bundle and
tensorflow/core/lib/io/{table,block,format}.cc, which are LevelDB's table format):

* **V2 "tensor bundle"** (the Saver default since TF 0.12): ``<prefix>.index`` is an SSTable
  whose key "" holds a BundleHeaderProto and whose other keys are variable names holding
  BundleEntryProto {dtype, shape, shard_id, offset, size, crc32c}; the bytes live in
  ``<prefix>.data-0000S-of-0000N`` at [offset, offset+size), little-endian.
* **V1**: one SSTable ``<prefix>``; key "" holds SavedTensorSlices{meta}, every other key a
  SavedTensorSlices{data: SavedSlice{name, slice, TensorProto}}.  Read-only here, full-tensor
  slices only (what a Saver over unpartitioned variables writes).

SSTable: [blocks][metaindex block][index block][48-byte footer: 2 BlockHandles (varint
offset, varint size), zero padding to 40 bytes, magic 0xdb4775248b80fb57 little-endian].
Block = entries (varint shared, varint non_shared, varint value_len, key suffix, value),
uint32 restart offsets, uint32 restart count; followed by 1 type byte (0 raw, 1 snappy) and
the masked CRC-32C of contents+type.

No TensorFlow-written file is available offline, so the reader is pinned by known-answer
tests of every primitive (CRC-32C, mask, varint, snappy, footer magic) and by round trips
through the writer (tests/test_tf_checkpoint.py); it has not yet met a real checkpoint.
"""
from __future__ import annotations

import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
FOOTER_LEN = 48
BLOCK_TRAILER = 5

# tensorflow/core/framework/types.proto
DTYPES = {1: np.dtype('<f4'), 2: np.dtype('<f8'), 3: np.dtype('<i4'), 4: np.dtype('u1'),
          5: np.dtype('<i2'), 6: np.dtype('i1'), 9: np.dtype('<i8'), 10: np.dtype('?'),
          17: np.dtype('<u2'), 19: np.dtype('<f2'), 22: np.dtype('<u4'), 23: np.dtype('<u8')}
DTYPE_ENUM = {v: k for k, v in DTYPES.items()}


class CheckpointError(ValueError):
  pass


# ------------------------------------------------------------------------------------------
# CRC-32C (Castagnoli), TF's masking (lib/hash/crc32c.h)
def _make_table():
  tbl = []
  for n in range(256):
    c = n
    for _ in range(8):
      c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
    tbl.append(c)
  return tbl


_CRC_TABLE = _make_table()


def crc32c(data, crc=0):
  """CRC-32C of a bytes-like object (bytewise table walk; ~5 MB/s, checkpoints are read once)."""
  c = crc ^ 0xFFFFFFFF
  tbl = _CRC_TABLE
  for b in bytes(data):
    c = tbl[(c ^ b) & 0xFF] ^ (c >> 8)
  return c ^ 0xFFFFFFFF


def mask_crc(crc):
  return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


def unmask_crc(masked):
  rot = (masked - 0xa282ead8) & 0xFFFFFFFF
  return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------
# varints and the three protobuf wire types the checkpoint protos use
def put_varint(n):
  out = bytearray()
  n &= (1 << 64) - 1
  while n >= 0x80:
    out.append((n & 0x7F) | 0x80)
    n >>= 7
  out.append(n)
  return bytes(out)


def get_varint(buf, pos):
  shift = 0
  val = 0
  while True:
    if pos >= len(buf):
      raise CheckpointError('truncated varint')
    b = buf[pos]
    pos += 1
    val |= (b & 0x7F) << shift
    if not b & 0x80:
      return val, pos
    shift += 7
    if shift > 63:
      raise CheckpointError('varint longer than 64 bits')


def parse_proto(buf):
  """Flat protobuf walk: yields (field number, wire type, value); value is an int for varint /
  fixed32 / fixed64 and a bytes slice for length-delimited fields."""
  pos = 0
  buf = bytes(buf)
  while pos < len(buf):
    key, pos = get_varint(buf, pos)
    field, wt = key >> 3, key & 7
    if wt == 0:
      val, pos = get_varint(buf, pos)
    elif wt == 1:
      val = struct.unpack_from('<Q', buf, pos)[0]
      pos += 8
    elif wt == 2:
      n, pos = get_varint(buf, pos)
      if pos + n > len(buf):
        raise CheckpointError('truncated length-delimited field')
      val = buf[pos:pos + n]
      pos += n
    elif wt == 5:
      val = struct.unpack_from('<I', buf, pos)[0]
      pos += 4
    else:
      raise CheckpointError('unsupported protobuf wire type %d' % wt)
    yield field, wt, val


def _field(field, wt, payload):
  key = put_varint((field << 3) | wt)
  if wt == 0:
    return key + put_varint(payload)
  if wt == 2:
    return key + put_varint(len(payload)) + payload
  if wt == 5:
    return key + struct.pack('<I', payload)
  raise ValueError(wt)


def _signed64(v):
  return v - (1 << 64) if v >= (1 << 63) else v


def parse_shape(buf):
  """TensorShapeProto {repeated Dim dim = 2 {int64 size = 1}; bool unknown_rank = 3}."""
  dims = []
  for f, _, v in parse_proto(buf):
    if f == 2:
      size = 0
      for f2, _, v2 in parse_proto(v):
        if f2 == 1:
          size = _signed64(v2)
      dims.append(size)
    elif f == 3 and v:
      raise CheckpointError('tensor of unknown rank in a checkpoint')
  return tuple(dims)


def encode_shape(shape):
  return b''.join(_field(2, 2, _field(1, 0, int(d))) for d in shape)


# ------------------------------------------------------------------------------------------
# snappy block decompression (format_description.txt of google/snappy)
def snappy_decompress(buf):
  buf = bytes(buf)
  n, pos = get_varint(buf, 0)
  out = bytearray()
  while pos < len(buf):
    tag = buf[pos]
    pos += 1
    kind = tag & 3
    if kind == 0:                                   # literal
      ln = tag >> 2
      if ln >= 60:
        nb = ln - 59
        ln = int.from_bytes(buf[pos:pos + nb], 'little')
        pos += nb
      ln += 1
      if pos + ln > len(buf):
        raise CheckpointError('snappy: truncated literal')
      out += buf[pos:pos + ln]
      pos += ln
      continue
    if kind == 1:                                   # copy, 1-byte offset
      ln = ((tag >> 2) & 7) + 4
      off = ((tag >> 5) << 8) | buf[pos]
      pos += 1
    elif kind == 2:                                 # copy, 2-byte offset
      ln = (tag >> 2) + 1
      off = buf[pos] | (buf[pos + 1] << 8)
      pos += 2
    else:                                           # copy, 4-byte offset
      ln = (tag >> 2) + 1
      off = int.from_bytes(buf[pos:pos + 4], 'little')
      pos += 4
    if off == 0 or off > len(out):
      raise CheckpointError('snappy: bad copy offset')
    for _ in range(ln):                             # overlapping copies are byte-serial
      out.append(out[-off])
  if len(out) != n:
    raise CheckpointError('snappy: length mismatch (%d != %d)' % (len(out), n))
  return bytes(out)


# ------------------------------------------------------------------------------------------
# SSTable
def _read_block(buf, offset, size, verify):
  end = offset + size
  if end + BLOCK_TRAILER > len(buf):
    raise CheckpointError('block handle past the end of the table')
  contents = buf[offset:end]
  btype = buf[end]
  if verify:
    want = unmask_crc(struct.unpack_from('<I', buf, end + 1)[0])
    if crc32c(buf[offset:end + 1]) != want:
      raise CheckpointError('table block checksum mismatch at offset %d' % offset)
  if btype == 1:
    contents = snappy_decompress(contents)
  elif btype != 0:
    raise CheckpointError('unknown block compression type %d' % btype)
  return contents


def _block_entries(block):
  if len(block) < 4:
    raise CheckpointError('table block too small')
  nrestart = struct.unpack_from('<I', block, len(block) - 4)[0]
  limit = len(block) - 4 - 4 * nrestart
  if limit < 0:
    raise CheckpointError('corrupt restart array')
  pos = 0
  key = b''
  while pos < limit:
    shared, pos = get_varint(block, pos)
    non_shared, pos = get_varint(block, pos)
    vlen, pos = get_varint(block, pos)
    if shared > len(key) or pos + non_shared + vlen > limit:
      raise CheckpointError('corrupt table entry')
    key = key[:shared] + block[pos:pos + non_shared]
    pos += non_shared
    yield key, block[pos:pos + vlen]
    pos += vlen


def read_table(path, verify=True):
  """All (key, value) pairs of an SSTable file, in key order."""
  with open(path, 'rb') as f:
    buf = f.read()
  if len(buf) < FOOTER_LEN:
    raise CheckpointError('%s: too short to be a table' % path)
  footer = buf[-FOOTER_LEN:]
  if struct.unpack('<Q', footer[40:])[0] != TABLE_MAGIC:
    raise CheckpointError('%s: not an SSTable (bad magic number)' % path)
  _, pos = get_varint(footer, 0)                   # metaindex handle (unused)
  _, pos = get_varint(footer, pos)
  ioff, pos = get_varint(footer, pos)
  isize, pos = get_varint(footer, pos)
  out = []
  for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
    boff, p2 = get_varint(handle, 0)
    bsize, _ = get_varint(handle, p2)
    out.extend(_block_entries(_read_block(buf, boff, bsize, verify)))
  return out


def _build_block(entries, restart_interval=16):
  out = bytearray()
  restarts = []
  last = b''
  for i, (key, value) in enumerate(entries):
    shared = 0
    if i % restart_interval == 0:
      restarts.append(len(out))
    else:
      m = min(len(last), len(key))
      while shared < m and last[shared] == key[shared]:
        shared += 1
    out += put_varint(shared) + put_varint(len(key) - shared) + put_varint(len(value))
    out += key[shared:] + value
    last = key
  if not restarts:
    restarts = [0]
  for r in restarts:
    out += struct.pack('<I', r)
  out += struct.pack('<I', len(restarts))
  return bytes(out)


def write_table(path, items, block_size=4096):
  """Uncompressed SSTable of the (key, value) pairs (keys must be sorted and unique)."""
  keys = [k for k, _ in items]
  if keys != sorted(set(keys)):
    raise ValueError('table keys must be sorted and unique')
  out = bytearray()
  index = []

  def emit(block):
    off = len(out)
    out.extend(block)
    out.append(0)                                              # kNoCompression
    out.extend(struct.pack('<I', mask_crc(crc32c(block + b'\x00'))))
    return put_varint(off) + put_varint(len(block))

  pending, size = [], 0
  for key, value in items:
    pending.append((key, value))
    size += len(key) + len(value) + 8
    if size >= block_size:
      index.append((pending[-1][0], emit(_build_block(pending))))
      pending, size = [], 0
  if pending:
    index.append((pending[-1][0], emit(_build_block(pending))))
  meta_handle = emit(_build_block([]))
  index_handle = emit(_build_block(index, restart_interval=1))
  footer = meta_handle + index_handle
  footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
  out.extend(footer)
  with open(path, 'wb') as f:
    f.write(out)


# ------------------------------------------------------------------------------------------
# V2 tensor bundle
def _parse_entry(buf):
  e = {'dtype': 0, 'shape': (), 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': None,
       'slices': 0}
  for f, _, v in parse_proto(buf):
    if f == 1: e['dtype'] = v
    elif f == 2: e['shape'] = parse_shape(v)
    elif f == 3: e['shard_id'] = v
    elif f == 4: e['offset'] = v
    elif f == 5: e['size'] = v
    elif f == 6: e['crc32c'] = v
    elif f == 7: e['slices'] += 1
  return e


def read_v2(prefix, names=None, verify=True):
  items = read_table(prefix + '.index', verify)
  if not items or items[0][0] != b'':
    raise CheckpointError('%s.index: missing bundle header' % prefix)
  num_shards, endian = 1, 0
  for f, _, v in parse_proto(items[0][1]):
    if f == 1: num_shards = v
    elif f == 2: endian = v
  if endian != 0:
    raise CheckpointError('big-endian tensor bundles are not supported')
  shards = {}
  out = {}
  for key, value in items[1:]:
    name = key.decode('utf-8')
    if names is not None and name not in names:
      continue
    e = _parse_entry(value)
    if e['slices']:
      raise CheckpointError('%s: partitioned (sliced) variables are not supported' % name)
    if e['dtype'] not in DTYPES:
      raise CheckpointError('%s: unsupported dtype enum %d' % (name, e['dtype']))
    dt = DTYPES[e['dtype']]
    count = int(np.prod(e['shape'], dtype=np.int64)) if e['shape'] else 1
    if count * dt.itemsize != e['size']:
      raise CheckpointError('%s: %d bytes stored for shape %s of %s' % (name, e['size'], e['shape'], dt))
    sid = e['shard_id']
    if sid not in shards:
      shards[sid] = np.memmap('%s.data-%05d-of-%05d' % (prefix, sid, num_shards), dtype=np.uint8,
                              mode='r')
    raw = shards[sid][e['offset']:e['offset'] + e['size']]
    if len(raw) != e['size']:
      raise CheckpointError('%s: data shard is truncated' % name)
    if verify and e['crc32c'] is not None and crc32c(raw) != unmask_crc(e['crc32c']):
      raise CheckpointError('%s: tensor checksum mismatch' % name)
    out[name] = np.frombuffer(bytes(raw), dtype=dt).reshape(e['shape']).copy()
  return out


def write_v2(prefix, tensors):
  """Write ``<prefix>.index`` + ``<prefix>.data-00000-of-00001`` that tf.train.Saver (V2)
  restores: entries sorted by name, raw little-endian bytes, masked CRC-32C per tensor."""
  os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
  header = _field(1, 0, 1) + _field(3, 2, _field(1, 0, 1))     # num_shards = 1, version.producer = 1
  items = [(b'', header)]
  offset = 0
  with open(prefix + '.data-00000-of-00001', 'wb') as f:
    for name in sorted(tensors, key=lambda s: s.encode('utf-8')):
      a = np.asarray(tensors[name])
      if a.ndim:
        a = np.ascontiguousarray(a)                  # (ascontiguousarray would make a scalar 1-D)
      dt = a.dtype.newbyteorder('<') if a.dtype.byteorder == '>' else a.dtype
      if np.dtype(dt) not in DTYPE_ENUM:
        raise CheckpointError('%s: dtype %s has no TensorFlow enum here' % (name, a.dtype))
      raw = a.astype(dt, copy=False).tobytes()
      f.write(raw)
      entry = _field(1, 0, DTYPE_ENUM[np.dtype(dt)])
      entry += _field(2, 2, encode_shape(a.shape))
      if offset:
        entry += _field(4, 0, offset)
      entry += _field(5, 0, len(raw))
      entry += _field(6, 5, mask_crc(crc32c(raw)))
      items.append((name.encode('utf-8'), entry))
      offset += len(raw)
  write_table(prefix + '.index', items)


# ------------------------------------------------------------------------------------------
# V1 (read-only)
def _parse_tensor_proto(buf, name):
  dtype, shape, content = 0, (), None
  packed = {}
  for f, wt, v in parse_proto(buf):
    if f == 1: dtype = v
    elif f == 2: shape = parse_shape(v)
    elif f == 4: content = v
    elif f in (5, 6, 7, 10):                        # float_val, double_val, int_val, int64_val
      packed.setdefault(f, []).append((wt, v))
  if dtype not in DTYPES:
    raise CheckpointError('%s: unsupported dtype enum %d' % (name, dtype))
  dt = DTYPES[dtype]
  count = int(np.prod(shape, dtype=np.int64)) if shape else 1
  if content is not None:
    arr = np.frombuffer(content, dtype=dt)
  else:
    fld = {1: 5, 2: 6, 3: 7, 9: 10}.get(dtype)
    vals = []
    for wt, v in packed.get(fld, []):
      if wt == 2:                                   # packed repeated
        if fld == 5: vals.append(np.frombuffer(v, dtype='<f4'))
        elif fld == 6: vals.append(np.frombuffer(v, dtype='<f8'))
        else:
          p, tmp = 0, []
          while p < len(v):
            x, p = get_varint(v, p)
            tmp.append(_signed64(x))
          vals.append(np.array(tmp, dtype=np.int64))
      elif wt == 5: vals.append(np.array([v], dtype='<u4').view('<f4'))
      elif wt == 1: vals.append(np.array([v], dtype='<u8').view('<f8'))
      else: vals.append(np.array([_signed64(v)], dtype=np.int64))
    arr = np.concatenate(vals).astype(dt) if vals else np.zeros(0, dt)
    if arr.size == 1 and count > 1:                 # TensorProto's "repeat the last value" rule
      arr = np.full(count, arr[0], dt)
  if arr.size != count:
    raise CheckpointError('%s: %d values for shape %s' % (name, arr.size, shape))
  return arr.reshape(shape).copy()


def read_v1(path, names=None, verify=True):
  out = {}
  for key, value in read_table(path, verify):
    if key == b'':
      continue                                       # SavedTensorSliceMeta: shapes only
    for f, _, v in parse_proto(value):
      if f != 2:
        continue                                     # SavedTensorSlices.data = 2
      name, tensor = None, None
      for f2, _, v2 in parse_proto(v):
        if f2 == 1: name = v2.decode('utf-8')
        elif f2 == 3: tensor = v2
      if name is None or tensor is None or (names is not None and name not in names):
        continue
      if name in out:
        raise CheckpointError('%s: partitioned (sliced) variables are not supported' % name)
      out[name] = _parse_tensor_proto(tensor, name)
  return out


# ------------------------------------------------------------------------------------------
def checkpoint_kind(path):
  """'v2' / 'v1' / None for a Saver path such as .../model.ckpt-87000."""
  if os.path.exists(path + '.index'):
    return 'v2'
  if path.endswith('.index') and os.path.exists(path):
    return 'v2'
  if os.path.isfile(path):
    with open(path, 'rb') as f:
      f.seek(0, 2)
      if f.tell() >= FOOTER_LEN:
        f.seek(-8, 2)
        if struct.unpack('<Q', f.read(8))[0] == TABLE_MAGIC:
          return 'v1'
  return None


def read_checkpoint(path, names=None, verify=True):
  """{variable name: ndarray} of a TF checkpoint given the path the reference passes to
  ``saver.restore`` (no extension)."""
  kind = checkpoint_kind(path)
  if kind == 'v2':
    return read_v2(path[:-6] if path.endswith('.index') else path, names, verify)
  if kind == 'v1':
    return read_v1(path, names, verify)
  raise CheckpointError('%s: neither a V2 (.index/.data) nor a V1 TensorFlow checkpoint' % path)
