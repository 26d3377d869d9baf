"""TensorFlow checkpoints (V2 "tensor bundles" and V1 single files) read without TensorFlow.

The reference restores ``zoo/inception_v2_2016_08_28/inception_v2.ckpt`` (a V1 single file) into the two
feature-extractor scopes (models/utils.py:179-186) and a text-classifier checkpoint (V2) into ``text_classifier/``
(models/label_extractor.py:456-458).  TensorFlow is not installable here, so this module restates the published
container formats.  A V1 checkpoint is ONE table file whose values embed the tensors (see ``load_v1_variables``);
a V2 checkpoint ``<prefix>`` is made of:

* ``<prefix>.index`` - a LevelDB-format sorted table (tensorflow/core/lib/io/table*): data blocks of prefix-compressed
  ``key -> value`` entries with a restart array, each block followed by a 1-byte compression type (0 raw, 1 snappy)
  and a masked CRC-32C; an index block mapping separator keys to block handles; a 48-byte footer holding the
  metaindex / index handles and the magic ``0xdb4775248b80fb57``.  Key ``""`` holds ``BundleHeaderProto``
  (num_shards, endianness, version), every other key is a tensor name holding ``BundleEntryProto``
  (dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6, slices=7).
* ``<prefix>.data-%05d-of-%05d`` - the raw little-endian tensor bytes at ``offset`` / ``size`` of shard ``shard_id``.

**Parity unpinned**: no TensorFlow checkpoint exists in this environment to read; the readers are checked against
``write_checkpoint`` / ``write_v1_checkpoint`` below (same format description, raw and snappy blocks) and
hand-built blocks.  Partitioned
variables (``slices``) and string tensors are refused.
"""
import os
import struct

import numpy as np

from cap2det_b200.tfrecord import _fields, _varint, _enc_varint, masked_crc32c

_MAGIC = 0xdb4775248b80fb57
_FOOTER = 48
# tensorflow/core/framework/types.proto
_DTYPES = {1: '<f4', 2: '<f8', 3: '<i4', 4: 'u1', 5: '<i2', 6: 'i1', 9: '<i8', 10: '?', 17: '<u2', 19: '<f2',
           22: '<u4', 23: '<u8'}
_DT_BFLOAT16, _DT_STRING = 14, 7


def snappy_decompress(data):
  """Snappy raw format: varint length, then literal / copy elements (tag low bits 00 literal, 01 / 10 / 11 copies
  with 1 / 2 / 4 offset bytes)."""
  data = bytes(data)
  n, pos = _varint(data, 0)
  out = bytearray()
  while pos < len(data):
    tag = data[pos]
    pos += 1
    kind = tag & 3
    if kind == 0:
      ln = tag >> 2
      if ln >= 60:
        nb = ln - 59
        ln = int.from_bytes(data[pos:pos + nb], 'little')
        pos += nb
      ln += 1
      out += data[pos:pos + ln]
      pos += ln
      continue
    if kind == 1:
      ln = ((tag >> 2) & 7) + 4
      off = ((tag >> 5) << 8) | data[pos]
      pos += 1
    elif kind == 2:
      ln = (tag >> 2) + 1
      off = int.from_bytes(data[pos:pos + 2], 'little')
      pos += 2
    else:
      ln = (tag >> 2) + 1
      off = int.from_bytes(data[pos:pos + 4], 'little')
      pos += 4
    if off == 0 or off > len(out):
      raise ValueError('snappy: bad copy offset')
    start = len(out) - off
    for i in range(ln):                       # copies may overlap their own output
      out.append(out[start + i])
  if len(out) != n:
    raise ValueError('snappy: length mismatch (%d != %d)' % (len(out), n))
  return bytes(out)


def _handle(buf, pos):
  offset, pos = _varint(buf, pos)
  size, pos = _varint(buf, pos)
  return offset, size, pos


def _read_block(fid, offset, size, verify=True):
  fid.seek(offset)
  raw = fid.read(size + 5)
  if len(raw) != size + 5:
    raise IOError('truncated table block')
  body, kind = raw[:size], raw[size]
  if verify and struct.unpack('<I', raw[size + 1:])[0] != masked_crc32c(raw[:size + 1]):
    raise IOError('table block checksum mismatch')
  if kind == 0:
    return body
  if kind == 1:
    return snappy_decompress(body)
  raise IOError('unknown block compression type %d' % kind)


def _block_entries(block):
  """Entries of one table block in key order (restart points only matter for seeking)."""
  num_restarts, = struct.unpack('<I', block[-4:])
  end = len(block) - 4 - 4 * num_restarts
  pos, key = 0, b''
  while pos < end:
    shared, pos = _varint(block, pos)
    non_shared, pos = _varint(block, pos)
    value_len, pos = _varint(block, pos)
    key = key[:shared] + block[pos:pos + non_shared]
    pos += non_shared
    yield key, block[pos:pos + value_len]
    pos += value_len


def _parse_entry(value):
  """BundleEntryProto -> dict(dtype, shape, shard_id, offset, size, crc32c, sliced)."""
  e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
  for num, wt, v in _fields(memoryview(value)):
    if num == 1:
      e['dtype'] = v
    elif num == 2:                                        # TensorShapeProto
      for n2, _, dim in _fields(v):
        if n2 == 2:                                       # Dim { size = 1 }
          size = 0
          for n3, _, x in _fields(dim):
            if n3 == 1:
              size = x
          e['shape'].append(size)
    elif num == 3:
      e['shard_id'] = v
    elif num == 4:
      e['offset'] = v
    elif num == 5:
      e['size'] = v
    elif num == 6:
      e['crc32c'] = struct.unpack('<I', v)[0]
    elif num == 7:
      e['sliced'] = True
  return e


def _table_entries(path):
  """All (key, value) pairs of a LevelDB-format table file, in key order."""
  total = os.path.getsize(path)
  if total < _FOOTER:
    raise IOError('%s: too short for a table' % path)
  with open(path, 'rb') as fid:
    fid.seek(total - _FOOTER)
    footer = fid.read(_FOOTER)
    if struct.unpack('<Q', footer[40:])[0] != _MAGIC:
      raise IOError('%s: not a TensorFlow checkpoint table (bad magic)' % path)
    _, _, pos = _handle(footer, 0)                        # metaindex: unused
    index_offset, index_size, _ = _handle(footer, pos)
    for _, handle in _block_entries(_read_block(fid, index_offset, index_size)):
      offset, size, _ = _handle(handle, 0)
      for key, value in _block_entries(_read_block(fid, offset, size)):
        yield key, value


def read_index(prefix):
  """``<prefix>.index`` -> (header dict(num_shards, endianness), {tensor name: entry dict})."""
  path = prefix + '.index'
  entries, header = {}, dict(num_shards=1, endianness=0)
  for key, value in _table_entries(path):
    if key == b'':
      for num, _, v in _fields(memoryview(value)):
        if num == 1:
          header['num_shards'] = v
        elif num == 2:
          header['endianness'] = v
    else:
      entries[key.decode('utf-8')] = _parse_entry(value)
  if header['endianness'] != 0:
    raise IOError('%s: big-endian checkpoints are not supported' % path)
  return header, entries


# ---- V1 checkpoints: ONE table file (tensorflow/core/util/tensor_slice_writer, saved_tensor_slice.proto) -------
# key "" -> SavedTensorSlices{meta=1: SavedTensorSliceMeta{tensor=1: SavedSliceMeta{name=1, shape=2, type=3, slice=4}}};
# every other key -> SavedTensorSlices{data=2: SavedSlice{name=1, slice=2, data=3: TensorProto}} where the TensorProto
# carries the values in its typed repeated field (float_val=5, double_val=6, int_val=7, int64_val=10, bool_val=11,
# half_val=13) or in tensor_content=4.  This is the format of slim's inception_v2_2016_08_28/inception_v2.ckpt.
_V1_VALUE_FIELDS = {5: ('<f4', 1), 6: ('<f8', 2), 7: (None, 3), 10: (None, 9), 11: (None, 10), 13: (None, 19)}


def is_v1_checkpoint(path):
  return os.path.isfile(path) and not os.path.exists(path + '.index')


def _shape_dims(shape_msg):
  dims = []
  for n2, _, dim in _fields(shape_msg):
    if n2 == 2:
      size = 0
      for n3, _, x in _fields(dim):
        if n3 == 1:
          size = x
      dims.append(size)
  return dims


def _slice_is_full(slice_msg):
  """TensorSliceProto{extent=1: Extent{start=1, length=2}}: an extent without fields covers its whole dimension."""
  for n, _, extent in _fields(slice_msg):
    if n == 1 and len(extent) > 0:
      return False
  return True


def _tensor_proto_values(proto, dtype_code):
  """Flat values of a TensorProto written by the V1 saver."""
  floats, ints, content = [], [], None
  for num, wt, v in _fields(proto):
    if num == 4:
      content = bytes(v)
    elif num in (5, 6):                                   # packed (wire type 2) or single fixed-width values
      floats.append(bytes(v))
    elif num in (7, 10, 11, 13):
      if wt == 0:
        ints.append(v)
      else:
        pos = 0
        while pos < len(v):
          x, pos = _varint(v, pos)
          ints.append(x)
  if dtype_code == _DT_BFLOAT16:
    raise NotImplementedError('bfloat16 in V1 checkpoints')
  np_dtype = _DTYPES[dtype_code]
  if content is not None:
    return np.frombuffer(content, np_dtype).copy()
  if dtype_code in (1, 2):
    return np.frombuffer(b''.join(floats), np_dtype).copy()
  if dtype_code == 19:                                    # half_val: uint16 bit patterns as varints
    return np.array(ints, np.uint16).view(np.float16)
  vals = np.array([x - (1 << 64) if x >= (1 << 63) else x for x in ints], np.int64)
  return vals.astype(np_dtype)


def load_v1_variables(path, names=None):
  """{name: array} of a V1 (single table file) checkpoint."""
  shapes, types, out = {}, {}, {}
  wanted = None if names is None else set(names)
  for key, value in _table_entries(path):
    for num, _, msg in _fields(memoryview(value)):
      if key == b'' and num == 1:                         # meta
        for n2, _, tensor in _fields(msg):
          if n2 != 1:
            continue
          name, dims, code = None, [], 0
          for n3, _, x in _fields(tensor):
            if n3 == 1:
              name = bytes(x).decode('utf-8')
            elif n3 == 2:
              dims = _shape_dims(x)
            elif n3 == 3:
              code = x
          shapes[name], types[name] = dims, code
      elif key != b'' and num == 2:                       # data
        name, full, proto = None, True, None
        for n3, _, x in _fields(msg):
          if n3 == 1:
            name = bytes(x).decode('utf-8')
          elif n3 == 2:
            full = _slice_is_full(x)
          elif n3 == 3:
            proto = x
        if wanted is not None and name not in wanted:
          continue
        if not full:
          raise NotImplementedError('%s is a partitioned variable' % name)
        code = types.get(name, 1)
        if code == _DT_STRING or (code not in _DTYPES and code != _DT_BFLOAT16):
          if names is None:
            continue
          raise NotImplementedError('%s has unsupported dtype %d' % (name, code))
        out[name] = _tensor_proto_values(proto, code).reshape(shapes[name])
  if wanted is not None:
    for name in names:
      if name not in out:
        raise KeyError('checkpoint %s lacks variable %s' % (path, name))
  return out


def list_variables(prefix):
  """[(name, shape)] like tf.train.list_variables."""
  if is_v1_checkpoint(prefix):
    return sorted((k, list(v.shape)) for k, v in load_v1_variables(prefix).items())
  return sorted((k, list(e['shape'])) for k, e in read_index(prefix)[1].items())


def load_variables(prefix, names=None, verify_crc=False):
  """{name: array} of the checkpoint (all tensors, or ``names``; a missing name raises KeyError).  ``prefix`` is a
  V2 prefix (``<prefix>.index`` exists) or the path of a V1 single-file checkpoint."""
  if is_v1_checkpoint(prefix):
    return load_v1_variables(prefix, names)
  header, entries = read_index(prefix)
  wanted = sorted(entries) if names is None else list(names)
  out, shards = {}, {}
  try:
    for name in wanted:
      if name not in entries:
        raise KeyError('checkpoint %s lacks variable %s' % (prefix, name))
      e = entries[name]
      if e['sliced']:
        raise NotImplementedError('%s is a partitioned variable' % name)
      if e['dtype'] == _DT_STRING or (e['dtype'] not in _DTYPES and e['dtype'] != _DT_BFLOAT16):
        if names is None:
          continue                                        # e.g. the saver's string bookkeeping tensors
        raise NotImplementedError('%s has unsupported dtype %d' % (name, e['dtype']))
      sid = e['shard_id']
      if sid not in shards:
        shards[sid] = open('%s.data-%05d-of-%05d' % (prefix, sid, header['num_shards']), 'rb')
      fid = shards[sid]
      fid.seek(e['offset'])
      raw = fid.read(e['size'])
      if len(raw) != e['size']:
        raise IOError('%s: truncated data for %s' % (prefix, name))
      if verify_crc and e['crc32c'] is not None and masked_crc32c(raw) != e['crc32c']:
        raise IOError('%s: checksum mismatch for %s' % (prefix, name))
      if e['dtype'] == _DT_BFLOAT16:
        arr = (np.frombuffer(raw, '<u2').astype(np.uint32) << 16).view(np.float32)
      else:
        arr = np.frombuffer(raw, _DTYPES[e['dtype']]).copy()
      out[name] = arr.reshape(e['shape'])
  finally:
    for fid in shards.values():
      fid.close()
  return out


# ---- writer (tests and exchange in the other direction) -------------------------------------------------
def _enc_tag(num, wt):
  return _enc_varint((num << 3) | wt)


def _enc_bytes(num, payload):
  return _enc_tag(num, 2) + _enc_varint(len(payload)) + payload


def _block(entries, restart_interval=16):
  out, restarts, prev = bytearray(), [], b''
  for i, (key, value) in enumerate(entries):
    shared = 0
    if i % restart_interval == 0:
      restarts.append(len(out))
    else:
      while shared < min(len(prev), len(key)) and prev[shared] == key[shared]:
        shared += 1
    out += _enc_varint(shared) + _enc_varint(len(key) - shared) + _enc_varint(len(value)) + key[shared:] + value
    prev = key
  if not restarts:
    restarts = [0]
  for r in restarts:
    out += struct.pack('<I', r)
  out += struct.pack('<I', len(restarts))
  return bytes(out)


def _snappy_literals(data):
  """A valid (incompressible-style) snappy stream: the length followed by literal elements only."""
  out = bytearray(_enc_varint(len(data)))
  for i in range(0, len(data), 65536):
    chunk = data[i:i + 65536]
    n = len(chunk) - 1
    if n < 60:
      out.append(n << 2)
    else:
      nb = (n.bit_length() + 7) // 8
      out.append((59 + nb) << 2)
      out += n.to_bytes(nb, 'little')
    out += chunk
  return bytes(out)


def write_checkpoint(prefix, variables, entries_per_block=8, snappy=False):
  """Writes ``variables`` ({name: array}) as a one-shard V2 checkpoint ``<prefix>.index`` + ``.data-00000-of-00001``."""
  rev = {np.dtype(v).str.lstrip('|'): k for k, v in _DTYPES.items()}
  data, items = bytearray(), []
  for name in sorted(variables, key=lambda s: s.encode('utf-8')):
    arr = np.asarray(variables[name])               # (ascontiguousarray would turn a scalar into shape [1])
    code = rev.get(arr.dtype.newbyteorder('<').str.lstrip('|'))
    if code is None:
      raise ValueError('unsupported dtype %s for %s' % (arr.dtype, name))
    raw = arr.astype(arr.dtype.newbyteorder('<')).tobytes()
    shape = b''.join(_enc_bytes(2, _enc_tag(1, 0) + _enc_varint(int(d))) for d in arr.shape)
    entry = (_enc_tag(1, 0) + _enc_varint(code) + _enc_bytes(2, shape) + _enc_tag(4, 0) + _enc_varint(len(data)) +
             _enc_tag(5, 0) + _enc_varint(len(raw)) + _enc_tag(6, 5) + struct.pack('<I', masked_crc32c(raw)))
    items.append((name.encode('utf-8'), entry))
    data += raw
  header = _enc_tag(1, 0) + _enc_varint(1) + _enc_bytes(3, _enc_tag(1, 0) + _enc_varint(1))
  items = [(b'', header)] + items
  with open(prefix + '.data-00000-of-00001', 'wb') as fid:
    fid.write(bytes(data))
  _write_table(prefix + '.index', items, entries_per_block, snappy)
  return prefix


def _write_table(path, items, entries_per_block, snappy):
  with open(path, 'wb') as fid:
    def put(block):
      body, kind = (_snappy_literals(block), 1) if snappy else (block, 0)
      offset = fid.tell()
      fid.write(body + bytes([kind]) + struct.pack('<I', masked_crc32c(body + bytes([kind]))))
      return _enc_varint(offset) + _enc_varint(len(body))
    index = []
    for i in range(0, len(items), entries_per_block):
      chunk = items[i:i + entries_per_block]
      index.append((chunk[-1][0], put(_block(chunk))))
    meta = put(_block([]))
    idx = put(_block(index, restart_interval=1))
    footer = meta + idx
    fid.write(footer + b'\0' * (40 - len(footer)) + struct.pack('<Q', _MAGIC))


def write_v1_checkpoint(path, variables, entries_per_block=4, snappy=False):
  """Writes ``variables`` as a V1 single-file checkpoint (float / double values in float_val / double_val, integers
  in int_val / int64_val).  Keys only need to sort after "" and be unique: the reader takes names from the values."""
  metas, items = b'', []
  for i, name in enumerate(sorted(variables, key=lambda s: s.encode('utf-8'))):
    arr = np.asarray(variables[name])
    kind = arr.dtype.kind
    if arr.dtype == np.float32:
      code, values = 1, _enc_bytes(5, arr.astype('<f4').tobytes())
    elif arr.dtype == np.float64:
      code, values = 2, _enc_bytes(6, arr.astype('<f8').tobytes())
    elif kind in 'iu' and arr.dtype.itemsize <= 4:
      code, values = 3, _enc_bytes(7, b''.join(_enc_varint(int(x)) for x in arr.reshape(-1)))
    elif arr.dtype == np.int64:
      code, values = 9, _enc_bytes(10, b''.join(_enc_varint(int(x)) for x in arr.reshape(-1)))
    else:
      raise ValueError('unsupported dtype %s for %s' % (arr.dtype, name))
    shape = b''.join(_enc_bytes(2, _enc_tag(1, 0) + _enc_varint(int(d))) for d in arr.shape)
    full_slice = b''.join(_enc_bytes(1, b'') for _ in arr.shape)
    metas += _enc_bytes(1, _enc_bytes(1, name.encode('utf-8')) + _enc_bytes(2, shape) + _enc_tag(3, 0) +
                        _enc_varint(code) + _enc_bytes(4, full_slice))
    tensor = _enc_tag(1, 0) + _enc_varint(code) + values
    saved = _enc_bytes(1, name.encode('utf-8')) + _enc_bytes(2, full_slice) + _enc_bytes(3, tensor)
    items.append((b'\x00' + name.encode('utf-8') + struct.pack('>I', i), _enc_bytes(2, saved)))
  items = [(b'', _enc_bytes(1, metas))] + items
  _write_table(path, items, entries_per_block, snappy)
  return path
