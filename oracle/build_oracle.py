"""Compiles the C restatement (oracle/scgrhc_oracle.c) into oracle/_build/liboracle.so with gcc.
The reference itself is Python, so there is no oracle/_ref to compile."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'scgrhc_oracle.c')
OUT_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(OUT_DIR, 'liboracle.so')


def build(force=False):
  if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
    return LIB
  os.makedirs(OUT_DIR, exist_ok=True)
  # -ffp-contract=off: no fused multiply-add, every fp64 operation rounds like numpy's
  subprocess.run(['gcc', '-O2', '-ffp-contract=off', '-fopenmp', '-fPIC', '-shared', SRC, '-o', LIB, '-lm'], check=True)
  return LIB


if __name__ == '__main__':
  print(build(force=True))
