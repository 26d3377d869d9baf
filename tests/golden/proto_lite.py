"""Test infrastructure: builds the reference's `protos.*_pb2` modules WITHOUT protoc.

The reference's configuration schema is a set of proto2 files (`/root/reference/protos/*.proto`); `build.sh:5`
compiles them with protoc, which this image does not have.  The protobuf PYTHON runtime is installed, so this
module parses the (small) proto2 subset those files use -- messages, nested enums, oneofs, `extend`, imports,
defaults -- into FileDescriptorProtos, registers them in a private descriptor pool and exposes the generated
message classes as modules `protos.<name>_pb2`, which is all the unmodified reference code imports.
Only `tests/golden/make_reference_outputs.py` uses it (it needs /root/reference, which exists in the build
container only).
"""
import os
import re
import sys
import types

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_SCALARS = {
    'double': 1, 'float': 2, 'int64': 3, 'uint64': 4, 'int32': 5, 'fixed64': 6, 'fixed32': 7, 'bool': 8,
    'string': 9, 'bytes': 12, 'uint32': 13, 'sfixed32': 15, 'sfixed64': 16, 'sint32': 17, 'sint64': 18,
}
_LABELS = {'optional': 1, 'required': 2, 'repeated': 3}
_TOKEN = re.compile(r'"(?:[^"\\]|\\.)*"|[A-Za-z_][A-Za-z0-9_.]*|-?[0-9][0-9.eE+-]*|[{}\[\]=;,()<>]')


def _tokens(text):
  text = re.sub(r'//[^\n]*', '', text)
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return _TOKEN.findall(text)


class _Parser(object):

  def __init__(self, text, name):
    self.t, self.i = _tokens(text), 0
    self.fd = descriptor_pb2.FileDescriptorProto(name=name, syntax='proto2')
    self.pending = []        # (field proto, type name as written, scope) resolved after all files are read

  def peek(self):
    return self.t[self.i] if self.i < len(self.t) else None

  def next(self):
    self.i += 1
    return self.t[self.i - 1]

  def expect(self, tok):
    got = self.next()
    assert got == tok, 'expected %r, got %r in %s' % (tok, got, self.fd.name)

  def parse(self):
    while self.peek() is not None:
      tok = self.next()
      if tok == 'syntax':
        self.expect('='); self.next(); self.expect(';')
      elif tok == 'import':
        self.fd.dependency.append(self.next().strip('"')); self.expect(';')
      elif tok == 'package':
        self.fd.package = self.next(); self.expect(';')
      elif tok == 'message':
        self.message(self.fd.message_type.add(), '')
      elif tok == 'enum':
        self.enum(self.fd.enum_type.add())
      elif tok == 'extend':
        self.extend(self.fd.extension, '')
      elif tok == ';':
        pass
      else:
        raise ValueError('unexpected %r in %s' % (tok, self.fd.name))
    return self

  def enum(self, ed):
    ed.name = self.next()
    self.expect('{')
    while self.peek() != '}':
      name = self.next(); self.expect('='); num = int(self.next())
      if self.peek() == '[':
        while self.next() != ']':
          pass
      self.expect(';')
      ed.value.add(name=name, number=num)
    self.expect('}')

  def options(self, fp):
    if self.peek() != '[':
      return
    self.next()
    while True:
      key = self.next(); self.expect('='); val = self.next()
      if key == 'default':
        fp.default_value = val.strip('"') if val.startswith('"') else val
      if self.next() == ']':
        break

  def field(self, fp, scope, label=None):
    if label is not None:
      fp.label = _LABELS[label]
    typ = self.next()
    fp.name = self.next()
    self.expect('=')
    fp.number = int(self.next())
    self.options(fp)
    self.expect(';')
    if typ in _SCALARS:
      fp.type = _SCALARS[typ]
    else:
      self.pending.append((fp, typ, scope))

  def extend(self, container, scope):
    extendee = self.next()
    self.expect('{')
    while self.peek() != '}':
      fp = container.add()
      fp.extendee = '.' + extendee
      self.field(fp, scope, self.next())
    self.expect('}')

  def message(self, md, scope):
    md.name = self.next()
    inner = (scope + '.' if scope else '') + md.name
    self.expect('{')
    while self.peek() != '}':
      tok = self.next()
      if tok in _LABELS:
        self.field(md.field.add(), inner, tok)
      elif tok == 'oneof':
        md.oneof_decl.add(name=self.next())
        index = len(md.oneof_decl) - 1
        self.expect('{')
        while self.peek() != '}':
          self.i -= 0
          fp = md.field.add()
          fp.label = 1
          fp.oneof_index = index
          self.field(fp, inner)
        self.expect('}')
      elif tok == 'message':
        self.message(md.nested_type.add(), inner)
      elif tok == 'enum':
        self.enum(md.enum_type.add())
      elif tok == 'extend':
        self.extend(md.extension, inner)
      elif tok == 'extensions':
        lo = int(self.next()); self.expect('to'); hi = self.next(); self.expect(';')
        md.extension_range.add(start=lo, end=536870912 if hi == 'max' else int(hi) + 1)
      elif tok == 'reserved':
        while self.next() != ';':
          pass
      elif tok == ';':
        pass
      else:
        raise ValueError('unexpected %r in message %s of %s' % (tok, md.name, self.fd.name))
    self.expect('}')


def _collect(prefix, messages, enums, table):
  for e in enums:
    table[prefix + e.name] = 14
  for m in messages:
    table[prefix + m.name] = 11
    _collect(prefix + m.name + '.', m.nested_type, m.enum_type, table)


def load(proto_dir, package='protos'):
  """Parses every .proto under `proto_dir` and installs `<package>.<name>_pb2` modules in sys.modules."""
  parsers = {}
  for f in sorted(os.listdir(proto_dir)):
    if f.endswith('.proto'):
      with open(os.path.join(proto_dir, f)) as fid:
        parsers[package + '/' + f] = _Parser(fid.read(), package + '/' + f).parse()
  table = {}
  for p in parsers.values():
    _collect('', p.fd.message_type, p.fd.enum_type, table)
  for p in parsers.values():
    for fp, typ, scope in p.pending:
      parts = scope.split('.') if scope else []
      full = None
      while True:                                # innermost scope first, like protoc
        cand = '.'.join(parts + [typ])
        if cand in table:
          full = cand
          break
        if not parts:
          break
        parts.pop()
      assert full is not None, 'unknown type %s in %s' % (typ, p.fd.name)
      fp.type = table[full]
      fp.type_name = '.' + full
  pool = descriptor_pool.DescriptorPool()
  done = set()

  def add(name):
    if name in done:
      return
    for dep in parsers[name].fd.dependency:
      add(dep)
    pool.Add(parsers[name].fd)
    done.add(name)

  for name in parsers:
    add(name)
  pkg = sys.modules.get(package) or types.ModuleType(package)
  pkg.__path__ = getattr(pkg, '__path__', [])
  sys.modules[package] = pkg
  for name, p in parsers.items():
    mod_name = os.path.basename(name)[:-len('.proto')] + '_pb2'
    mod = types.ModuleType(package + '.' + mod_name)
    fdesc = pool.FindFileByName(name)
    mod.DESCRIPTOR = fdesc
    for mname, mdesc in fdesc.message_types_by_name.items():
      setattr(mod, mname, message_factory.GetMessageClass(mdesc))
    for ename, edesc in fdesc.enum_types_by_name.items():
      for v in edesc.values:
        setattr(mod, v.name, v.number)
    sys.modules[package + '.' + mod_name] = mod
    setattr(pkg, mod_name, mod)
  return pool
