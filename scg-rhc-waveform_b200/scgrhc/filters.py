"""Extension (named by the project brief, absent from the reference; default off): zero-phase band-pass filtering of
record columns on the device with scipy.signal.sosfiltfilt semantics.  Host part: initial conditions and pad length,
computed with the same operations scipy uses (scipy/signal/_signaltools.py: lfilter_zi, sosfilt_zi, sosfiltfilt), so
the device result is bit-identical to scipy's.  scipy itself is only needed to *design* a filter from corner
frequencies (``butter_sos``); explicit ``sos`` coefficients need no scipy.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as N
from . import ops


def lfilter_zi_biquad(b, a):
  """scipy.signal.lfilter_zi for a second-order section (a[0] == 1), operation for operation (scipy 1.18
  _signaltools.py: y_inf = sum(b)/sum(a); zi = flip(cumsum(flip(b - y_inf*a)))[1:])."""
  b, a = np.asarray(b, dtype=np.float64), np.asarray(a, dtype=np.float64)
  if a[0] != 1:
    b, a = b / a[0], a / a[0]
  y_inf = np.sum(b) / np.sum(a)
  c = b - y_inf * a
  return np.array([c[2] + c[1], c[2]])


def sosfilt_zi(sos):
  sos = np.asarray(sos, dtype=np.float64)
  zi = np.empty((sos.shape[0], 2))
  scale = 1.0
  for s in range(sos.shape[0]):
    b, a = sos[s, :3], sos[s, 3:]
    zi[s] = scale * lfilter_zi_biquad(b, a)
    scale *= b.sum() / a.sum()
  return zi


def pad_length(sos):
  """scipy's default padlen for sosfiltfilt: 3 * ntaps, ntaps = 2 n + 1 - min(#(b2 == 0), #(a2 == 0))."""
  sos = np.asarray(sos, dtype=np.float64)
  ntaps = 2 * sos.shape[0] + 1 - min(int((sos[:, 2] == 0).sum()), int((sos[:, 5] == 0).sum()))
  return 3 * ntaps


def butter_sos(low_hz, high_hz, fs, order=4):
  """Butterworth band-pass as second-order sections (filter *design* is host-side configuration; needs scipy)."""
  from scipy.signal import butter
  return butter(order, [low_hz, high_hz], btype='bandpass', fs=fs, output='sos')


def sosfiltfilt(arena, record_rows, sos, columns):
  """Filter ``columns`` of every record of ``arena`` ((rows, ncols) fp64 CUDA; records back to back, ``record_rows``
  rows each) forward-backward; returns a new arena whose other columns are copied unchanged."""
  if not arena.is_cuda:
    raise RuntimeError('sosfiltfilt needs a CUDA arena (no CPU fallback)')
  sos = np.ascontiguousarray(sos, dtype=np.float64)
  if sos.ndim != 2 or sos.shape[1] != 6:
    raise ValueError('sos must be shape (n_sections, 6)')
  zi = np.ascontiguousarray(sosfilt_zi(sos))
  edge = pad_length(sos)
  rows = np.asarray(record_rows, dtype=np.int64)
  row0 = np.ascontiguousarray(np.concatenate([[0], np.cumsum(rows)]).astype(np.int64))
  if row0[-1] != arena.shape[0]:
    raise ValueError('record_rows do not add up to the arena')
  cols = (C.c_int32 * len(columns))(*columns)
  dev = arena.device
  out = arena.clone()
  tmp = torch.empty((int(row0[-1]) + 2 * edge * len(rows)) * len(columns), dtype=torch.float64, device=dev)
  row0_dev = torch.from_numpy(row0).to(dev)
  c = ops.ctx(dev.index)
  N.check(c, N.lib().scgrhc_sosfiltfilt(c, ops._ptr(arena), ops._ptr(out), ops._ptr(tmp), ops._ptr(row0_dev),
                                        row0.ctypes.data_as(C.POINTER(C.c_int64)), len(rows), arena.shape[1], cols,
                                        len(columns), sos.ctypes.data_as(C.POINTER(C.c_double)),
                                        zi.ctypes.data_as(C.POINTER(C.c_double)), sos.shape[0], edge,
                                        ops._stream(dev.index)))
  return out
