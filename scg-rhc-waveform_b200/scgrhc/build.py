"""Builds libscgrhc.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(HERE, 'libscgrhc.so')
SOURCES = [os.path.join(CSRC, 'api.cu')]
HEADERS = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))) + [os.path.join(ROOT, 'include', 'scgrhc.h')]


def nvcc_path():
  return shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'


def needs_build():
  if not os.path.exists(LIB):
    return True
  t = os.path.getmtime(LIB)
  return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
  """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -shared -> scgrhc/libscgrhc.so"""
  if not force and not needs_build():
    return LIB
  cmd = [nvcc_path(), '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '-shared', '-I' + os.path.join(ROOT, 'include'), '-I' + CSRC]
  if verbose:
    cmd += ['-Xptxas', '-v']
  cmd += SOURCES + ['-o', LIB, '-ldl']
  subprocess.run(cmd, check=True)
  return LIB


if __name__ == '__main__':
  import sys
  print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
