"""Builds cap2det_b200/lib/libcap2det_b200.so (sm_100a only) with nvcc, in-tree.

The library is a plain C-ABI shared object (include/cap2det_b200.h); it is loaded with ctypes
by cap2det_b200/capi.py.  Nothing here depends on torch.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, 'csrc')
LIB_DIR = os.path.join(ROOT, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libcap2det_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']

def _sources():
  return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _fingerprint():
  h = hashlib.sha256()
  for d, names in ((CSRC, sorted(os.listdir(CSRC))),
                   (os.path.join(ROOT, '..', 'include'), sorted(os.listdir(os.path.join(ROOT, '..', 'include'))))):
    for f in names:
      if f.endswith(('.cu', '.cuh', '.h')):
        h.update(f.encode())
        with open(os.path.join(d, f), 'rb') as fid:
          h.update(fid.read())
  h.update(' '.join(FLAGS).encode())
  return h.hexdigest()


def _up_to_date(fp):
  stamp = os.path.join(LIB_DIR, 'build.stamp')
  if os.path.exists(LIB_PATH) and os.path.exists(stamp):
    with open(stamp) as fid:
      return fid.read().strip() == fp
  return False


def build_library(force=False, verbose=False):
  """Compiles every .cu under csrc/ and links the shared library.  Returns its path.

  Safe under torchrun: an exclusive file lock serialises the ranks of one node (the first one builds, the others find
  the stamp up to date), objects are compiled into a private directory and the library and its stamp are moved into
  place with os.replace(), so no process can dlopen a half-written file."""
  import fcntl
  import shutil
  import tempfile
  os.makedirs(LIB_DIR, exist_ok=True)
  fp = _fingerprint()
  if not force and _up_to_date(fp):
    return LIB_PATH
  if not os.path.exists(NVCC):
    raise RuntimeError('nvcc not found at %s and no up-to-date %s' % (NVCC, LIB_PATH))
  with open(os.path.join(LIB_DIR, '.build.lock'), 'w') as lock:
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
      if not force and _up_to_date(fp):          # another rank built it while this one waited for the lock
        return LIB_PATH
      tmp = tempfile.mkdtemp(prefix='.build-', dir=LIB_DIR)
      try:
        def compile_one(src):
          obj = os.path.join(tmp, src[:-3] + '.o')
          cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
          r = subprocess.run(cmd, capture_output=True, text=True)
          if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
          if verbose:
            sys.stderr.write(r.stderr)
          return obj

        with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
          objs = list(ex.map(compile_one, _sources()))
        tmp_lib = os.path.join(tmp, 'libcap2det_b200.so')
        r = subprocess.run([NVCC, '-shared', '-o', tmp_lib] + objs + ['-lcudart', '-Xlinker', '--no-undefined'],
                           capture_output=True, text=True)      # an unresolved c2d_* symbol fails the build, not the first dlopen
        if r.returncode != 0:
          raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
        tmp_stamp = os.path.join(tmp, 'build.stamp')
        with open(tmp_stamp, 'w') as fid:
          fid.write(fp)
        os.replace(tmp_lib, LIB_PATH)
        os.replace(tmp_stamp, os.path.join(LIB_DIR, 'build.stamp'))
      finally:
        shutil.rmtree(tmp, ignore_errors=True)
    finally:
      fcntl.flock(lock, fcntl.LOCK_UN)
  return LIB_PATH


if __name__ == '__main__':
  print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
