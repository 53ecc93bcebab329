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


def sosfiltfilt(arena, record_rows, sos, columns, exact=True, chunk=0, nbuf=0, inplace=False):
  """Filter ``columns`` of every record of ``arena`` ((rows, ncols) fp64 CUDA; records back to back, ``record_rows``
  rows each) forward-backward; returns a new arena whose other columns are copied unchanged.

  ``exact=True``: serial-in-time systolic kernel, bit-identical to scipy.  ``exact=False``: time-parallel scan (one warp
  per record, lanes filter chunks of ``chunk`` rows concurrently, chunk-boundary states through a scan over the lanes),
  an order of magnitude faster, equal to scipy within rounding noise (<= 1e-10 of full scale); up to 4 sections, 4
  columns per launch (more columns are looped, the later groups in place).  ``inplace=True`` (scan only) filters
  ``arena`` itself and returns it."""
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
  if not arena.is_contiguous() or arena.dtype != torch.float64:
    raise ValueError('arena must be a contiguous float64 tensor')
  dev = arena.device
  row0_dev = torch.from_numpy(row0).to(dev)
  c = ops.ctx(dev.index)
  p_sos, p_zi = sos.ctypes.data_as(C.POINTER(C.c_double)), zi.ctypes.data_as(C.POINTER(C.c_double))
  p_row0 = row0.ctypes.data_as(C.POINTER(C.c_int64))
  if exact or sos.shape[0] > 4:
    if inplace:
      raise ValueError('inplace filtering needs the time-parallel kernel (exact=False, at most 4 sections)')
    out = arena.clone()
    cols = (C.c_int32 * len(columns))(*columns)
    tmp = torch.empty((int(row0[-1]) + 2 * edge * len(rows)) * len(columns), dtype=torch.float64, device=dev)
    N.check(c, N.lib().scgrhc_sosfiltfilt(c, ops._ptr(arena), ops._ptr(out), ops._ptr(tmp), ops._ptr(row0_dev), p_row0,
                                          len(rows), arena.shape[1], cols, len(columns), p_sos, p_zi, sos.shape[0], edge,
                                          ops._stream(dev.index)))
    return out
  out = arena if inplace else torch.empty_like(arena)
  src = arena
  for g0 in range(0, len(columns), 4):
    grp = list(columns[g0:g0 + 4])
    cols = (C.c_int32 * len(grp))(*grp)
    N.check(c, N.lib().scgrhc_sosfiltfilt_scan(c, ops._ptr(src), ops._ptr(out), ops._ptr(row0_dev), p_row0, len(rows),
                                               arena.shape[1], cols, len(grp), p_sos, p_zi, sos.shape[0], edge, int(chunk),
                                               int(nbuf), ops._stream(dev.index)))
    src = out                                              # the first launch copied every other column through
  return out


def _output_len(len_h, in_len, up, down):
  """scipy.signal._upfirdn_apply._output_len."""
  in_len_copy = in_len + (len_h + (-len_h % up)) // up - 1
  nt = in_len_copy * up
  return nt // down + (1 if nt % down else 0)


def resample_design(n_in, up, down, window=('kaiser', 5.0)):
  """Everything scipy.signal.resample_poly derives before calling upfirdn, for an input of n_in samples:
  (up, down) reduced, n_out, transposed-flipped taps, taps per phase, n_pre_remove (needs scipy for firwin)."""
  import math
  from scipy.signal import firwin
  g = math.gcd(int(up), int(down))
  up, down = int(up) // g, int(down) // g
  n_out = n_in * up
  n_out = n_out // down + bool(n_out % down)
  max_rate = max(up, down)
  half_len = 10 * max_rate
  h = firwin(2 * half_len + 1, 1.0 / max_rate, window=window).astype(np.float64)
  h *= up
  n_pre_pad = down - half_len % down
  n_post_pad = 0
  n_pre_remove = (half_len + n_pre_pad) // down
  while _output_len(len(h) + n_pre_pad + n_post_pad, n_in, up, down) < n_out + n_pre_remove:
    n_post_pad += 1
  h = np.concatenate([np.zeros(n_pre_pad), h, np.zeros(n_post_pad)])
  padlen = len(h) + (-len(h) % up)                       # _upfirdn.py:_pad_h
  h_full = np.zeros(padlen)
  h_full[:len(h)] = h
  taps = np.ascontiguousarray(h_full.reshape(-1, up).T[:, ::-1]).ravel()
  return up, down, n_out, taps, padlen // up, n_pre_remove, n_post_pad


def resample_poly(arena, record_rows, up, down, window=('kaiser', 5.0), exact=True):
  """Resample every record of ``arena`` ((rows, ncols) fp64 CUDA) by up/down; returns (new_arena, new_record_rows).
  ``up == down`` after reduction returns a copy, as scipy does.  ``exact=True``: bit-identical to scipy; ``exact=False``
  (integer decimation only): one FMA per tap, ~1.7x faster, within 1e-14 of scipy."""
  import math
  if not arena.is_cuda:
    raise RuntimeError('resample_poly needs a CUDA arena (no CPU fallback)')
  rows = [int(r) for r in record_rows]
  g = math.gcd(int(up), int(down))
  if up // g == down // g == 1:
    return arena.clone(), rows
  lengths = sorted(set(rows))
  designs = dict(zip(lengths, [resample_design(n, up, down, window) for n in lengths]))
  first = designs[lengths[0]]
  if any(d[6] != first[6] for d in designs.values()):
    # the FIR is post-padded only when the output would otherwise come out short; scipy notes "we should rarely need
    # to do this given our filter lengths" and it has not been observed with the kaiser design used here
    raise NotImplementedError('records whose lengths need different FIR post-padding in one cohort')
  out_rows = [designs[n][2] for n in rows]
  dev = arena.device
  in0 = torch.from_numpy(np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)).to(dev)
  out0_h = np.concatenate([[0], np.cumsum(out_rows)]).astype(np.int64)
  out0 = torch.from_numpy(out0_h).to(dev)
  out = torch.empty((int(out0_h[-1]), arena.shape[1]), dtype=torch.float64, device=dev)
  taps = torch.from_numpy(first[3]).to(dev)
  c = ops.ctx(dev.index)
  for lo in range(0, len(rows), 65535):
    n = min(65535, len(rows) - lo)
    N.check(c, N.lib().scgrhc_resample_poly(c, ops._ptr(arena), ops._ptr(out), ops._ptr(taps), ops._ptr(in0[lo:]), ops._ptr(out0[lo:]),
                                            n, max(out_rows[lo:lo + n]), arena.shape[1], first[0], first[1], first[4], first[5],
                                            0 if exact else 1, ops._stream(dev.index)))
  return out, out_rows


class DecimSpec:
  """What the decimating front end of the window kernel needs (``engine.prepare_windows(..., decim=...)``): the FIR of
  scipy.signal.resample_poly for an integer decimation and the native / model-rate row counts of every record of the
  arena.  ``DecimSpec.design(record_rows, rate, native)`` derives it exactly as ``resample_poly`` does."""

  def __init__(self, taps, per_phase, down, n_pre_remove, in_rows, out_rows, fused=False):
    self.fused = bool(fused)        # one FMA per tap (~1e-15 of scipy) instead of the bit-identical multiply + add
    self.taps, self.per_phase, self.down, self.n_pre_remove = np.ascontiguousarray(taps, dtype=np.float64), int(per_phase), int(down), int(n_pre_remove)
    self.in_rows, self.out_rows = [int(v) for v in in_rows], [int(v) for v in out_rows]

  @classmethod
  def design(cls, record_rows, up, down, window=('kaiser', 5.0), fused=False):
    import math
    rows = [int(r) for r in record_rows]
    g = math.gcd(int(up), int(down))
    if up // g != 1 or down // g < 2:
      raise ValueError('the decimating front end does integer decimation (up/down = 1/k, k >= 2); got %d/%d' % (up, down))
    lengths = sorted(set(rows))
    designs = {n: resample_design(n, up, down, window) for n in lengths}
    first = designs[lengths[0]]
    if any(d[6] != first[6] for d in designs.values()):
      raise NotImplementedError('records whose lengths need different FIR post-padding in one cohort')
    if first[4] > 128:
      raise NotImplementedError('more than 128 taps per output')
    return cls(first[3], first[4], first[1], first[5], rows, [designs[n][2] for n in rows], fused)

  def tables(self, plan, device):
    """Per interval of ``plan`` (planned on the model-rate rows): arena row of the record's first native-rate row, its
    native-rate length, and the interval's first model-rate row relative to the record — int64 CUDA tensors."""
    out_base = np.concatenate([[0], np.cumsum(self.out_rows)]).astype(np.int64)
    in_base = np.concatenate([[0], np.cumsum(self.in_rows)]).astype(np.int64)
    row0 = plan.intervals['row0'].astype(np.int64)
    r = np.searchsorted(out_base, row0, side='right') - 1
    mk = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(device)
    return mk(in_base[r]), mk(np.asarray(self.in_rows, dtype=np.int64)[r]), mk(row0 - out_base[r])
