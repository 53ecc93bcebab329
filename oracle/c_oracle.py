"""TEST / BASELINE INFRASTRUCTURE — ctypes wrapper of oracle/_build/liboracle.so."""
import ctypes as C
import os

import numpy as np

from . import build_oracle

_lib = None


def lib():
  global _lib
  if _lib is None:
    path = build_oracle.LIB
    if not os.path.exists(path):
      build_oracle.build()
    _lib = C.CDLL(path)
    _lib.oracle_process_windows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int64,
                                            C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_int]
  return _lib


def process_windows(arena, W, cols, rcol, row_start, min_rhc, thr=1e-3, write=True, threads=None):
  """keep, reason, minmax, scg (n, C, W) f32, rhc (n, 1, W) f32 for windows starting at ``row_start``."""
  arena = np.ascontiguousarray(arena, dtype=np.float64)
  row_start = np.ascontiguousarray(row_start, dtype=np.int64)
  n, Cn = len(row_start), len(cols)
  cols_a = np.ascontiguousarray(cols, dtype=np.int32)
  keep = np.zeros(n, np.uint8); reason = np.zeros(n, np.uint8); mm = np.zeros((n, 4))
  scg = np.zeros((n, Cn, W), np.float32) if write else None
  rhc = np.zeros((n, 1, W), np.float32) if write else None
  if threads is not None:
    os.environ['OMP_NUM_THREADS'] = str(threads)
  rc = lib().oracle_process_windows(arena.ctypes.data, arena.shape[1], W, Cn, cols_a.ctypes.data, rcol, n,
                                    row_start.ctypes.data, float(min_rhc), thr, keep.ctypes.data, reason.ctypes.data,
                                    mm.ctypes.data, scg.ctypes.data if write else None,
                                    rhc.ctypes.data if write else None, 0)
  if rc:
    raise RuntimeError('oracle_process_windows failed')
  return keep.astype(bool), reason, mm, scg, rhc
