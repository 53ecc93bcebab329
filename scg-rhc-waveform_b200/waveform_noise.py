"""Drop-in for the reference's ``waveform_noise`` (waveform_noise.py:6-49): the three predicates that
reject a candidate RHC window, evaluated on the GPU (fp64, bit-compatible decisions).

Signatures are the reference's.  Single-waveform calls exist for API parity; the hot path evaluates
the same predicates for every candidate window inside the fused window kernel
(``scgrhc.prepare_windows``).  ``has_noise_batch`` is the batched form.
"""
import numpy as np
import torch

from scgrhc import _native as N
from scgrhc import engine, ops

_DEV = None


def _device():
  global _DEV
  if _DEV is None:
    if not torch.cuda.is_available():
      raise RuntimeError('waveform_noise needs a CUDA device: the predicates run in libscgrhc (no CPU fallback)')
    _DEV = torch.device('cuda', torch.cuda.current_device())
  return _DEV


def _upload(waveform):
  y = np.ascontiguousarray(np.asarray(waveform, dtype=np.float64).reshape(-1))
  return y, torch.from_numpy(y).to(_device())


def get_flat_lines(waveform, threshold=1e-3, min_duration=0.1, sampling_rate=500):
  """Flat segments as the reference reports them, quirk included (waveform_noise.py:6-26): the list is
  non-empty iff at least two window positions have a rolling range below ``threshold``.  The rolling
  max-min comparison runs on the device; the (start, end) bookkeeping over the flagged indices is
  integer logic on the host."""
  m = int(min_duration * sampling_rate)
  y, d = _upload(waveform)
  flags = torch.empty(len(y), dtype=torch.uint8, device=d.device)
  if len(y):
    ops.rolling_range_lt(d, m, float(threshold), flags)
  idx = np.nonzero(flags.cpu().numpy())[0].tolist()
  segs, start = [], None
  for i in range(len(idx) - 1):
    if start is None:
      start = idx[i]
    if idx[i + 1] != idx[i] + 1:
      segs.append((start, idx[i]))
      start = None
    if start is not None:
      segs.append((start, idx[-1]))
  return segs


def _stats(y, d, min_rhc):
  st = torch.empty((1, 6), dtype=torch.float64, device=d.device)
  ops.waveform_stats(d.reshape(1, -1), float(min_rhc), st)
  return st.cpu().numpy()[0]


def is_straight_line(waveform):
  """R^2 of the OLS line through (0..n-1, y) > 0.8 (waveform_noise.py:29-34).  Non-finite input raises
  ValueError as sklearn does.  An exactly constant waveform is decided the way the reference's rounding
  noise decides it: R^2 is 1.0 when ``np.mean(y) == y[0]`` exactly, else 0.0 (see oracle/scgrhc_oracle.py)."""
  y, d = _upload(waveform)
  r2, mn, mx, _, nonfinite, _ = _stats(y, d, float('-inf'))
  if nonfinite:
    raise ValueError('Input y contains NaN.' if np.isnan(y).any() else 'Input y contains infinity or a value too large for dtype(\'float64\').')
  if mn == mx:
    return bool(np.mean(y) == y[0])
  return bool(r2 > 0.8)


def in_rhc_range(params, waveform):
  """False iff a sample is strictly below ``params.min_RHC`` (waveform_noise.py:37-41)."""
  y, d = _upload(waveform)
  if len(y) == 0:
    return True
  return not bool(_stats(y, d, params.min_RHC)[3])


def has_noise(params, waveform):
  """flat OR straight OR below-floor (waveform_noise.py:44-49).  Windows of 2..1024 samples go through
  the fused window kernel exactly as the hot path does; other lengths use the standalone kernels."""
  y, d = _upload(waveform)
  L = len(y)
  if 2 <= L <= 1024:
    keep, reason = has_noise_batch(params, d.reshape(1, L), _raw=True)
    r = int(reason[0])
    if r & N.REASON_FLAT:
      return True
    if r & N.REASON_NONFINITE:
      raise ValueError('Input y contains NaN.')
    if y.max() == y.min():
      return bool(np.mean(y) == y[0]) or bool(r & N.REASON_FLOOR)
    return bool(r & (N.REASON_STRAIGHT | N.REASON_FLOOR))
  return (len(get_flat_lines(y)) > 0 or is_straight_line(y) or not in_rhc_range(params, y))


def has_noise_batch(params, windows, _raw=False):
  """has_noise for every row of ``windows`` (n, L) at once (numpy array or CUDA tensor).  Returns a bool
  numpy array; raises ValueError if a non-finite window reaches the regression, as the reference would
  on the first such window."""
  if isinstance(windows, torch.Tensor):
    w = windows.to(_device(), torch.float64).contiguous()
  else:
    w = torch.from_numpy(np.ascontiguousarray(windows, dtype=np.float64)).to(_device())
  n, L = w.shape
  plan = engine.Plan(np.array([(0, 0, n, 0)], dtype=engine.INTERVAL_DTYPE), n, L)
  st = engine.prepare_windows(w.reshape(-1, 1), plan, [0], 0, params.min_RHC, predicates_only=True, check=not _raw)
  keep, reason = st.keep.cpu().numpy().astype(bool), st.reason.cpu().numpy()
  if _raw:
    return keep, reason
  return ~keep
