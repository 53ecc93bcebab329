"""TEST INFRASTRUCTURE — numpy restatement of the reference's window-preparation path.

This is the parity oracle for the CUDA path.  It is NOT shipped behaviour: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may import it.  The product (``scg-rhc-waveform_b200/``) never does and fails loudly without its
CUDA library.

Pinning: the reference ships no golden vectors, so this oracle is pinned against
outputs of the *unmodified reference itself*, produced in the build container by
``tests/golden/make_golden.py`` through ``oracle/ref_harness.py`` and committed under
``tests/golden/`` (checked by ``tests/test_oracle.py``).  Installed third-party versions that
generated them: numpy 2.3.5, pandas 3.0.2, scikit-learn 1.9.0, scipy 1.18.1, torch 2.11.0
(the reference pins none).

Every function cites the reference lines it restates (paths relative to ``/root/reference``).
"""
from datetime import datetime

import numpy as np

SAMPLE_FREQ = 500            # recordutil.py:19
FLAT_THRESHOLD = 1e-3        # waveform_noise.py:6 (default used by has_noise, :46)
FLAT_MIN_SAMPLES = 50        # int(0.1 * 500), waveform_noise.py:7
R2_THRESHOLD = 0.8           # waveform_noise.py:34
NORM_EPS = 0.0001            # recordutil.py:46
RHC_NAME = 'RHC_pressure'    # recordutil.py:140


# ----------------------------------------------------------------------------------------
# a4  get_chamber_intervals  (recordutil.py:93-110)
# ----------------------------------------------------------------------------------------
def chamber_intervals(meta, chamber):
  """Sample ranges [(a, b), ...] of every event whose key prefix (before '_') is ``chamber``.

  recordutil.py:100-101 parse only the HH:MM:SS token; :103 non-dict events -> [];
  :104 'END' = elapsed seconds; :105 stable sort by time; :107-109 every event but the last,
  bounds int(t*500) (truncation toward zero)."""
  t0 = datetime.strptime(meta['MacStTime'].split()[1], '%H:%M:%S')
  t1 = datetime.strptime(meta['MacEndTime'].split()[1], '%H:%M:%S')
  events = meta['ChamEvents_in_s']
  if not isinstance(events, dict):
    return []
  events = dict(events)
  events['END'] = (t1 - t0).total_seconds()
  order = sorted(events.items(), key=lambda kv: kv[1])
  out = []
  for i in range(len(order) - 1):
    if order[i][0].split('_')[0] == chamber:
      out.append((int(order[i][1] * SAMPLE_FREQ), int(order[i + 1][1] * SAMPLE_FREQ)))
  return out


# ----------------------------------------------------------------------------------------
# a5/a6  get_channels + window enumeration  (recordutil.py:113-119, 136-146)
# ----------------------------------------------------------------------------------------
def candidate_windows(intervals, T, W, stride=None):
  """Candidate windows of one record, in the reference's order (interval order, then i).

  ``p_signal[a:b]`` (recordutil.py:118) follows Python slice semantics, so the bounds are
  clamped (and negative values wrap) exactly as ``slice(a, b).indices(T)`` does; the number
  of windows is ``L // W`` (:141) and ``start_idx = i*W`` is relative to the interval (:143).
  ``stride`` (extension, not in the reference): rows between window starts; None = W.
  Returns int64 arrays (abs_start, rel_start, interval_index)."""
  abs_start, rel_start, which = [], [], []
  for k, (a, b) in enumerate(intervals):
    lo, hi, _ = slice(a, b).indices(T)
    L, st = max(0, hi - lo), (W if stride is None else stride)
    n = L // W if st == W else ((L - W) // st + 1 if L >= W else 0)
    for i in range(n):
      abs_start.append(lo + i * st)
      rel_start.append(i * st)
      which.append(k)
  return (np.asarray(abs_start, dtype=np.int64), np.asarray(rel_start, dtype=np.int64),
          np.asarray(which, dtype=np.int32))


# ----------------------------------------------------------------------------------------
# a7  get_flat_lines  (waveform_noise.py:6-26)
# ----------------------------------------------------------------------------------------
def rolling_range(y, m=FLAT_MIN_SAMPLES):
  """fl(rolling max - rolling min) over ``m`` samples for positions m-1 .. len-1
  (waveform_noise.py:10-11; the first m-1 pandas outputs are NaN and never compare true)."""
  y = np.asarray(y, dtype=np.float64)
  if y.shape[-1] < m:
    return np.empty(y.shape[:-1] + (0,), dtype=np.float64)
  v = np.lib.stride_tricks.sliding_window_view(y, m, axis=-1)
  return v.max(axis=-1) - v.min(axis=-1)       # NaN in a window -> NaN -> compares false, as pandas


def flat_count(y, threshold=FLAT_THRESHOLD, m=FLAT_MIN_SAMPLES):
  """Number of window positions whose rolling range is < threshold (strict, :13)."""
  with np.errstate(invalid='ignore'):
    return (rolling_range(y, m) < threshold).sum(axis=-1)


def flat_segments(y, threshold=FLAT_THRESHOLD, min_duration=0.1, sampling_rate=500):
  """The reference's return value, quirk included (waveform_noise.py:13-26): the final
  ``if start is not None: append`` sits inside the loop, so a (start, last) tuple is appended
  on every iteration in which a run is open.  Non-empty iff count(d < thr) >= 2."""
  m = int(min_duration * sampling_rate)
  y = np.asarray(y, dtype=np.float64)
  with np.errstate(invalid='ignore'):
    idx = (np.nonzero(rolling_range(y, m) < threshold)[0] + (m - 1)).tolist()
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


# ----------------------------------------------------------------------------------------
# a8  is_straight_line  (waveform_noise.py:29-34)
# ----------------------------------------------------------------------------------------
def r_squared(y):
  """R^2 of the OLS line through (0..n-1, y) with intercept == Sxy^2 / (Sxx * Syy).

  The reference goes sklearn LinearRegression -> scipy lstsq -> r2_score (1 - SSres/SStot);
  the closed form agrees to ~6e-16, so decisions can differ only when R^2 is
  within ~1e-15 of 0.8.  Evaluated in long double with centred sums to sit well inside that.
  Exactly constant y is decided by rounding noise in the reference: sklearn centres y by
  ``np.average(y)``; if that mean is exact, SSres == SStot == 0 and r2_score returns 1.0
  (flagged), otherwise SSres == SStot != 0 and R^2 == 0.0 (golden cases 'constant' vs
  'run750@0').  Restated literally; it can only matter for windows shorter than 51 samples,
  because longer constant windows are already rejected as flat lines."""
  y64 = np.asarray(y, dtype=np.float64)
  y = y64.astype(np.longdouble)
  n = y.shape[-1]
  x = np.arange(n, dtype=np.longdouble) - np.longdouble(n - 1) / 2
  yc = y - y.mean(axis=-1, keepdims=True)
  sxx = (x * x).sum()
  sxy = (x * yc).sum(axis=-1)
  syy = (yc * yc).sum(axis=-1)
  const = (y64.max(axis=-1) == y64.min(axis=-1))
  const_r2 = np.where(y64.mean(axis=-1) == y64[..., 0], 1.0, 0.0)
  with np.errstate(invalid='ignore', divide='ignore'):
    r2 = np.where(const, const_r2, (sxy * sxy / (sxx * syy)).astype(np.float64))
  return r2.astype(np.float64)


def is_straight_line(y):
  return r_squared(y) > R2_THRESHOLD


# ----------------------------------------------------------------------------------------
# a9/a10  in_rhc_range, has_noise  (waveform_noise.py:37-49)
# ----------------------------------------------------------------------------------------
def below_floor(y, min_rhc):
  """True iff any sample is strictly below ``min_rhc`` (NaN never is; :38-40)."""
  with np.errstate(invalid='ignore'):
    return (np.asarray(y, dtype=np.float64) < min_rhc).any(axis=-1)


def has_noise(y, min_rhc):
  """flat OR straight OR below-floor (waveform_noise.py:44-49).  Raises ValueError exactly
  where the reference does: a non-finite sample reaching sklearn (i.e. not already rejected
  as flat by the short-circuit ``or``)."""
  y = np.asarray(y, dtype=np.float64)
  flat = flat_count(y) >= 2
  if flat:
    return True
  if not np.isfinite(y).all():
    raise ValueError('Input y contains NaN.' if np.isnan(y).any() else 'Input y contains infinity')
  return bool(is_straight_line(y)) or bool(below_floor(y, min_rhc))


# ----------------------------------------------------------------------------------------
# a12-a14  per-window min/max, normalise, transpose, cast  (recordutil.py:41-66)
# ----------------------------------------------------------------------------------------
def minmax_norm(x, mn, mx):
  """(x - mn) / (mx - mn + 0.0001) in fp64 (recordutil.py:45-46)."""
  return (x - mn) / (mx - mn + NORM_EPS)


def window_minmax(scg, rhc):
  """(scg_min, scg_max, rhc_min, rhc_max): one pair jointly over all SCG channels
  (recordutil.py:58-59).  scg: (..., W, C), rhc: (..., W)."""
  return (scg.min(axis=(-2, -1)), scg.max(axis=(-2, -1)), rhc.min(axis=-1), rhc.max(axis=-1))


# ----------------------------------------------------------------------------------------
# whole-record restatement of get_segments + SCGDataset.init_segments
# ----------------------------------------------------------------------------------------
class RecordWindows:
  """All candidate windows of one record plus the reference's verdict on each."""
  __slots__ = ('abs_start', 'rel_start', 'interval', 'keep', 'flat_count', 'r2', 'floor',
               'nonfinite', 'minmax', 'W', 'C')


def scan_record(p_signal, sig_name, meta, in_channels, chamber, segment_size, min_rhc, stride=None):
  """get_segments for one record (recordutil.py:133-149) without materialising windows.
  ``list.index`` raises ValueError for a missing channel (:117), as here."""
  W = int(segment_size * SAMPLE_FREQ)
  cols = [list(sig_name).index(n) for n in in_channels]
  rcol = list(sig_name).index(RHC_NAME)
  T = p_signal.shape[0]
  intervals = chamber_intervals(meta, chamber)
  out = RecordWindows()
  out.W, out.C = W, len(cols)
  out.abs_start, out.rel_start, out.interval = candidate_windows(intervals, T, W, stride)
  n = len(out.abs_start)
  idx = out.abs_start[:, None] + np.arange(W)[None, :]
  rhc = p_signal[:, rcol][idx] if n else np.empty((0, W))
  out.flat_count = flat_count(rhc).astype(np.int64) if n else np.zeros(0, np.int64)
  out.nonfinite = ~np.isfinite(rhc).all(axis=-1)
  flat = out.flat_count >= 2
  bad = out.nonfinite & ~flat
  if bad.any():
    raise ValueError('Input y contains NaN.')
  with np.errstate(invalid='ignore'):
    out.r2 = r_squared(rhc) if n else np.zeros(0)
  out.floor = below_floor(rhc, min_rhc) if n else np.zeros(0, bool)
  out.keep = ~(flat | (out.r2 > R2_THRESHOLD) | out.floor)
  scg = p_signal[:, cols][idx] if n else np.empty((0, W, len(cols)))
  mm = window_minmax(scg, rhc) if n else [np.zeros(0)] * 4
  out.minmax = np.stack(mm, axis=-1).astype(np.float64) if n else np.zeros((0, 4))
  return out


def normalise_record(p_signal, sig_name, in_channels, rw, global_minmax=None, out_dtype=np.float32):
  """SCGDataset.init_segments for the kept windows of ``rw`` (recordutil.py:55-66):
  returns scg (n_kept, C, W), rhc (n_kept, 1, W) of ``out_dtype`` and the (n_kept, 4) pairs used."""
  cols = [list(sig_name).index(n) for n in in_channels]
  rcol = list(sig_name).index(RHC_NAME)
  k = np.nonzero(rw.keep)[0]
  idx = rw.abs_start[k][:, None] + np.arange(rw.W)[None, :]
  scg = p_signal[:, cols][idx]                  # (n, W, C)
  rhc = p_signal[:, rcol][idx]                  # (n, W)
  mm = rw.minmax[k].copy()
  if global_minmax is not None:
    mm[:] = np.asarray(global_minmax, dtype=np.float64)[None, :]
  s = minmax_norm(scg, mm[:, 0, None, None], mm[:, 1, None, None])
  r = minmax_norm(rhc, mm[:, 2, None], mm[:, 3, None])
  return (np.ascontiguousarray(s.transpose(0, 2, 1)).astype(out_dtype),
          r[:, None, :].astype(out_dtype), mm)


def zscore_record(p_signal, sig_name, in_channels, rw, out_dtype=np.float32):
  """Extension (ABSENT from the reference — parity unpinned; plain numpy is the definition): per-window z-score of the kept
  windows, (x - mean) / (std + 0.0001) with mean / population std taken jointly over the (W, C) SCG block, as the
  reference takes its min/max (recordutil.py:58), and over the RHC window.  Returns scg, rhc and (n_kept, 4) rows
  {scg_mean, scg_std, rhc_mean, rhc_std}."""
  cols = [list(sig_name).index(n) for n in in_channels]
  rcol = list(sig_name).index(RHC_NAME)
  k = np.nonzero(rw.keep)[0]
  idx = rw.abs_start[k][:, None] + np.arange(rw.W)[None, :]
  scg = p_signal[:, cols][idx]
  rhc = p_signal[:, rcol][idx]
  ms = np.stack([scg.mean(axis=(1, 2)), scg.std(axis=(1, 2)), rhc.mean(axis=1), rhc.std(axis=1)], axis=1)
  s = (scg - ms[:, 0, None, None]) / (ms[:, 1, None, None] + 0.0001)
  r = (rhc - ms[:, 2, None]) / (ms[:, 3, None] + 0.0001)
  return (np.ascontiguousarray(s.transpose(0, 2, 1)).astype(out_dtype), r[:, None, :].astype(out_dtype), ms)


def global_minmax(minmax_rows):
  """get_global_minmax_vals (recordutil.py:152-169) over the kept windows' (n,4) rows."""
  m = np.asarray(minmax_rows, dtype=np.float64)
  return np.array([m[:, 0].min(), m[:, 1].max(), m[:, 2].min(), m[:, 3].max()])
