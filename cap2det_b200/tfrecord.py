"""On-disk input format of the reference: TFRecord files of tf.Example protos, as written
by ``dataset-tools/create_pascal_tf_record.py:140-193`` and parsed by ``readers/cap2det_reader.py:30-140``.

TensorFlow is not installable here, so the two container formats are read directly:

* TFRecord framing: ``uint64 length | uint32 masked_crc32c(length) | data | uint32 masked_crc32c(data)``
  (little endian; CRC-32C = Castagnoli, mask = rotate right 15 + 0xa282ead8);
* tf.Example = ``Example{features=1: Features{feature=1: map<string, Feature>}}`` with
  ``Feature{bytes_list=1 | float_list=2 | int64_list=3}``, each list ``{value=1}`` (floats / ints packed or not).

``decode_example`` mirrors ``_parse_fn``: JPEG decode (PIL), caption ``parse_texts``, proposal / object boxes,
object texts -- the host-side dict that ``cap2det_b200.reader.parse_example`` / ``make_batch`` consume.
"""
import io
import struct

import numpy as np

from cap2det_b200.standard_fields import InputDataFields as F

# core/standard_fields.py:35-60 (TFExampleDataFields)
IMAGE_ID = 'image/source_id'
IMAGE_ENCODED = 'image/encoded'
CAPTION_STRING = 'image/caption/string'
CAPTION_OFFSET = 'image/caption/offset'
CAPTION_LENGTH = 'image/caption/length'
PROPOSAL_BOX = 'image/proposal/bbox'
OBJECT_BOX = 'image/object/bbox'
OBJECT_TEXT = 'image/object/class/text'
OBJECT_LABEL = 'image/object/class/label'

# ---- CRC-32C (Castagnoli), table driven -------------------------------------------------------------
_POLY = 0x82F63B78
_TABLE = []
for _i in range(256):
  _c = _i
  for _ in range(8):
    _c = (_c >> 1) ^ _POLY if _c & 1 else _c >> 1
  _TABLE.append(_c)
_TABLE = np.array(_TABLE, np.uint32)


def _crc_scalar(crc, data):
  tab = _TABLE
  for b in data:
    crc = int(tab[(crc ^ b) & 0xFF]) ^ (crc >> 8)
  return crc


_CHUNK = 4096
_ZERO_SHIFT = []          # 4 x 256 table: the register after _CHUNK zero bytes, per byte of the starting register


def _zero_shift_tables():
  if not _ZERO_SHIFT:
    state = (np.arange(256, dtype=np.uint32)[None, :] << (8 * np.arange(4, dtype=np.uint32))[:, None]).reshape(-1)
    for _ in range(_CHUNK):
      state = _TABLE[state & 0xFF] ^ (state >> 8)
    _ZERO_SHIFT.append(state.reshape(4, 256))
  return _ZERO_SHIFT[0]


def crc32c(data):
  """CRC-32C.  The register update is linear over GF(2): for long inputs the 4 KiB chunks are run through the
  byte table side by side (NumPy, all chunks at once, from register 0) and then chained with the precomputed
  'advance by 4 KiB of zeros' map; short inputs take the plain byte loop."""
  data = bytes(data)
  crc = 0xFFFFFFFF
  n_chunks = len(data) // _CHUNK
  if n_chunks >= 16:
    body = np.frombuffer(data, np.uint8, n_chunks * _CHUNK).reshape(n_chunks, _CHUNK)
    state = np.zeros(n_chunks, np.uint32)
    for j in range(_CHUNK):
      state = _TABLE[(state ^ body[:, j]) & 0xFF] ^ (state >> 8)
    z = _zero_shift_tables()
    for c in state.tolist():
      crc = int(z[0][crc & 0xFF] ^ z[1][(crc >> 8) & 0xFF] ^ z[2][(crc >> 16) & 0xFF] ^ z[3][crc >> 24]) ^ c
    data = data[n_chunks * _CHUNK:]
  return _crc_scalar(crc, data) ^ 0xFFFFFFFF


def masked_crc32c(data):
  crc = crc32c(data)
  return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- TFRecord framing -----------------------------------------------------------------------------------
def read_records(path, verify_length_crc=True, verify_data_crc=False):
  """Yields the raw records of a TFRecord file.  The length CRC is cheap and always worth checking; the data CRC
  of a JPEG-sized record costs milliseconds in pure Python, so it is optional."""
  with open(path, 'rb') as fid:
    while True:
      header = fid.read(12)
      if not header:
        return
      if len(header) != 12:
        raise IOError('%s: truncated record header' % path)
      length, = struct.unpack('<Q', header[:8])
      if verify_length_crc and struct.unpack('<I', header[8:])[0] != masked_crc32c(header[:8]):
        raise IOError('%s: corrupted record length' % path)
      data = fid.read(length)
      footer = fid.read(4)
      if len(data) != length or len(footer) != 4:
        raise IOError('%s: truncated record' % path)
      if verify_data_crc and struct.unpack('<I', footer)[0] != masked_crc32c(data):
        raise IOError('%s: corrupted record data' % path)
      yield data


def write_records(path, records):
  with open(path, 'wb') as fid:
    for data in records:
      head = struct.pack('<Q', len(data))
      fid.write(head + struct.pack('<I', masked_crc32c(head)) + data + struct.pack('<I', masked_crc32c(data)))


# ---- protobuf wire format, just enough for tf.Example ---------------------------------------------------
def _varint(buf, pos):
  result, shift = 0, 0
  while True:
    b = buf[pos]
    pos += 1
    result |= (b & 0x7F) << shift
    if not b & 0x80:
      return result, pos
    shift += 7


def _fields(buf):
  """Yields (field number, wire type, value) of one message; value = int (varint / fixed) or memoryview (bytes)."""
  pos, n = 0, len(buf)
  while pos < n:
    key, pos = _varint(buf, pos)
    num, wt = key >> 3, key & 7
    if wt == 0:
      val, pos = _varint(buf, pos)
    elif wt == 1:
      val = bytes(buf[pos:pos + 8]); pos += 8
    elif wt == 2:
      ln, pos = _varint(buf, pos)
      val = buf[pos:pos + ln]; pos += ln
    elif wt == 5:
      val = bytes(buf[pos:pos + 4]); pos += 4
    else:
      raise ValueError('unsupported protobuf wire type %d' % wt)
    yield num, wt, val


def _signed64(v):
  return v - (1 << 64) if v >= (1 << 63) else v


def parse_example(data):
  """tf.Example bytes -> {feature name: list of bytes | np.float32 array | np.int64 array}."""
  out = {}
  buf = memoryview(data)
  for num, wt, features in _fields(buf):
    if num != 1 or wt != 2:
      continue
    for num2, wt2, entry in _fields(features):            # map<string, Feature> entries
      if num2 != 1 or wt2 != 2:
        continue
      key, feature = None, None
      for num3, wt3, v in _fields(entry):
        if num3 == 1:
          key = bytes(v).decode('utf-8')
        elif num3 == 2:
          feature = v
      value = []
      if feature is not None:
        for kind, wtk, lst in _fields(feature):
          if kind == 1:                                    # BytesList
            value = [bytes(v) for n4, _, v in _fields(lst) if n4 == 1]
          elif kind == 2:                                  # FloatList: packed (wire type 2) or repeated fixed32
            vals = []
            for n4, w4, v in _fields(lst):
              if n4 == 1:
                vals.append(bytes(v))
            value = np.frombuffer(b''.join(vals), '<f4').copy()
          elif kind == 3:                                  # Int64List: packed varints or repeated varints
            vals = []
            for n4, w4, v in _fields(lst):
              if n4 != 1:
                continue
              if w4 == 0:
                vals.append(_signed64(v))
              else:
                pos, mv = 0, v
                while pos < len(mv):
                  x, pos = _varint(mv, pos)
                  vals.append(_signed64(x))
            value = np.array(vals, np.int64)
      out[key] = value
  return out


def _enc_varint(v):
  v &= (1 << 64) - 1
  out = bytearray()
  while True:
    b = v & 0x7F
    v >>= 7
    out.append(b | (0x80 if v else 0))
    if not v:
      return bytes(out)


def _enc_field(num, payload):
  return _enc_varint((num << 3) | 2) + _enc_varint(len(payload)) + payload


def encode_example(features):
  """{name: list of bytes/str | float sequence | int sequence} -> tf.Example bytes (lists are written packed, as
  TensorFlow does)."""
  entries = b''
  for key in sorted(features):
    v = features[key]
    if isinstance(v, (bytes, str)):
      v = [v]
    v = list(v) if not isinstance(v, np.ndarray) else v
    if len(v) == 0 and not isinstance(v, np.ndarray):
      feature = b''                              # no kind set: TensorFlow reads it as an empty list of any type
    elif len(v) > 0 and isinstance(v[0], (bytes, str)):
      lst = b''.join(_enc_field(1, x.encode('utf-8') if isinstance(x, str) else x) for x in v)
      feature = _enc_field(1, lst)
    elif isinstance(v, np.ndarray) and v.dtype.kind == 'f' or (len(v) > 0 and isinstance(v[0], float)):
      feature = _enc_field(2, _enc_field(1, np.asarray(v, '<f4').tobytes()))
    else:
      feature = _enc_field(3, _enc_field(1, b''.join(_enc_varint(int(x)) for x in v)))
    entries += _enc_field(1, _enc_field(1, key.encode('utf-8')) + _enc_field(2, feature))
  return _enc_field(1, entries)


# ---- readers/cap2det_reader.py:30-140 (_parse_fn), host side ---------------------------------------------
def _boxes(parsed, prefix):
  """slim tfexample_decoder.BoundingBox(prefix): [n, 4] = (ymin, xmin, ymax, xmax)."""
  cols = [np.asarray(parsed.get(prefix + '/' + k, []), np.float32) for k in ('ymin', 'xmin', 'ymax', 'xmax')]
  n = len(cols[0])
  if any(len(c) != n for c in cols):
    raise ValueError('inconsistent bounding box lists under %r' % prefix)
  return np.stack(cols, axis=-1).reshape(n, 4) if n else np.zeros((0, 4), np.float32)


def decode_jpeg(encoded):
  """tf.image.decode_jpeg(channels=3) -> uint8 [H, W, 3] (RGB)."""
  from PIL import Image
  with Image.open(io.BytesIO(encoded)) as im:
    return np.array(im.convert('RGB'), np.uint8)           # own, writable buffer


def decode_example(data, decode_image=True):
  """One tf.Example record -> the host-side feature dict of ``_parse_fn`` (before flip / truncation, which
  ``cap2det_b200.reader.parse_example`` applies): image uint8 [H,W,3], proposals / object boxes fp32 [n,4],
  object texts, caption tokens with per-caption offsets / lengths."""
  parsed = parse_example(data)
  tokens = [t.decode('utf-8') for t in parsed.get(CAPTION_STRING, [])]
  offsets = [int(x) for x in parsed.get(CAPTION_OFFSET, [])]
  lengths = [int(x) for x in parsed.get(CAPTION_LENGTH, [])]
  if len(offsets) != len(lengths):
    raise ValueError('Not equal: num_offsets and num_lengths')        # core/preprocess.py:168-170
  max_len = max(lengths) if lengths else 0
  caption_strings = [tokens[o:o + l] + [''] * (max_len - l) for o, l in zip(offsets, lengths)]
  out = {
      F.image_id: parsed[IMAGE_ID][0].decode('utf-8') if parsed.get(IMAGE_ID) else '',
      F.num_captions: len(offsets),
      F.caption_strings: caption_strings,
      F.caption_lengths: lengths,
      F.concat_caption_string: tokens,
      F.concat_caption_length: len(tokens),
      F.proposals: _boxes(parsed, PROPOSAL_BOX),
      F.object_boxes: _boxes(parsed, OBJECT_BOX),
      F.object_texts: [t.decode('utf-8') for t in parsed.get(OBJECT_TEXT, [])],
  }
  if decode_image:
    out[F.image] = decode_jpeg(parsed[IMAGE_ENCODED][0])
  return out


def read_examples(paths, decode_image=True):
  """All examples of the given TFRecord files, in file order."""
  for path in ([paths] if isinstance(paths, str) else paths):
    for record in read_records(path):
      yield decode_example(record, decode_image=decode_image)


# ---- dataset-tools/create_pascal_tf_record.py:80-193 (dict_to_tf_example) ----------------------------------------
def pascal_example(encoded_jpg, filename, objects, proposals, label_map_dict, ignore_difficult_instances=False):
  """The tf.Example the reference writes per VOC image: ``objects`` = the annotation's object dicts (name, bndbox
  in pixels, difficult, truncated, pose), ``proposals`` = the ``<image_id>.npy`` array [n,4] of normalised
  (ymin, xmin, ymax, xmax) boxes, or its path.  Boxes are normalised by the decoded image size; the class names
  double as the caption (offset [0], length [#objects]), as in :172-177.  Returns the serialized record."""
  import hashlib
  from PIL import Image
  with Image.open(io.BytesIO(encoded_jpg)) as image:
    if image.format != 'JPEG':
      raise ValueError('Image format not JPEG')
    height, width = image.height, image.width
  if isinstance(proposals, str):
    with open(proposals, 'rb') as fid:
      proposals = np.load(fid)
  proposals = np.asarray(proposals, np.float32).reshape(-1, 4)
  xmin, ymin, xmax, ymax, classes, classes_text, truncated, poses, difficult_obj = [], [], [], [], [], [], [], [], []
  for obj in objects:
    difficult = bool(int(obj.get('difficult', 0)))
    if ignore_difficult_instances and difficult:
      continue
    difficult_obj.append(int(difficult))
    xmin.append(float(obj['bndbox']['xmin']) / width)
    ymin.append(float(obj['bndbox']['ymin']) / height)
    xmax.append(float(obj['bndbox']['xmax']) / width)
    ymax.append(float(obj['bndbox']['ymax']) / height)
    classes_text.append(obj['name'])
    classes.append(int(label_map_dict[obj['name']]))
    truncated.append(int(obj.get('truncated', 0)))
    poses.append(obj.get('pose', 'Unspecified'))
  f32 = lambda v: np.asarray(v, np.float32)
  i64 = lambda v: np.asarray(v, np.int64)
  return encode_example({
      'image/height': i64([height]), 'image/width': i64([width]), 'image/filename': [filename],
      IMAGE_ID: [filename], 'image/key/sha256': [hashlib.sha256(encoded_jpg).hexdigest()],
      IMAGE_ENCODED: [encoded_jpg], 'image/format': ['jpeg'],
      OBJECT_BOX + '/xmin': f32(xmin), OBJECT_BOX + '/xmax': f32(xmax), OBJECT_BOX + '/ymin': f32(ymin),
      OBJECT_BOX + '/ymax': f32(ymax), OBJECT_TEXT: classes_text, OBJECT_LABEL: i64(classes),
      'image/object/difficult': i64(difficult_obj), 'image/object/truncated': i64(truncated),
      'image/object/view': poses, CAPTION_STRING: classes_text, CAPTION_OFFSET: i64([0]),
      CAPTION_LENGTH: i64([len(classes_text)]),
      PROPOSAL_BOX + '/ymin': proposals[:, 0], PROPOSAL_BOX + '/xmin': proposals[:, 1],
      PROPOSAL_BOX + '/ymax': proposals[:, 2], PROPOSAL_BOX + '/xmax': proposals[:, 3]})
