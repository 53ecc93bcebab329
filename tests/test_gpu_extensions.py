"""GPU: north-star extensions that the reference does not have (default off).  Oracle = scipy itself; the reference
cannot pin these ("parity unpinned by the reference")."""
import numpy as np
import pytest
import torch
from scipy import signal

from oracle import synth_ref
from tests import helpers as H

pytestmark = pytest.mark.gpu

from scgrhc import filters  # noqa: E402

DEV = 'cuda:0'


@pytest.mark.parametrize('order,band,btype', [(4, (1.0, 40.0), 'bandpass'), (2, (0.5, 20.0), 'bandpass'), (4, 30.0, 'low'),
                                              (3, 2.0, 'high'), (8, (0.8, 45.0), 'bandpass')])
def test_sosfiltfilt_matches_scipy(order, band, btype):
  sos = signal.butter(order, band, btype=btype, fs=500, output='sos')
  assert (filters.sosfilt_zi(sos) == signal.sosfilt_zi(sos)).all()
  sig = synth_ref.SIG_NAMES_5
  rows = [9000, 311, 20001]
  recs = [synth_ref.gen_record(H.SEED, 50 + r, T, kinds=synth_ref.kinds_for(sig)) for r, T in enumerate(rows)]
  arena = torch.from_numpy(np.concatenate(recs)).to(DEV)
  cols = [0, 2, 4, 3]
  out = filters.sosfiltfilt(arena, rows, sos, cols).cpu().numpy()
  at = 0
  worst = 0.0
  for p in recs:
    want = p.copy()
    want[:, cols] = signal.sosfiltfilt(sos, p[:, cols], axis=0)
    got = out[at:at + len(p)]
    assert (got[:, 1] == p[:, 1]).all()                                    # untouched column copied through
    scale = np.abs(want[:, cols]).max(axis=0)
    worst = max(worst, float((np.abs(got[:, cols] - want[:, cols]) / scale).max()))
    at += len(p)
  assert worst <= 1e-10, worst                                             # BASELINE north_star: fp64 mode within 1e-10
  print('sosfiltfilt max scaled error', worst)


def test_sosfiltfilt_rejects_short_records_like_scipy():
  sos = signal.butter(4, (1.0, 40.0), btype='bandpass', fs=500, output='sos')
  arena = torch.zeros((20, 2), dtype=torch.float64, device=DEV)
  from scgrhc._native import ScgrhcError
  with pytest.raises(ScgrhcError, match='greater than padlen'):
    filters.sosfiltfilt(arena, [20], sos, [0])


@pytest.mark.parametrize('up,down', [(1, 2), (250, 500), (2, 5), (3, 2), (125, 500), (500, 500)])
def test_resample_poly_matches_scipy(up, down):
  sig = synth_ref.DEFAULT_SIG_NAMES
  rows = [5000, 1234, 777]
  recs = [synth_ref.gen_record(H.SEED, 70 + r, T, kinds=synth_ref.kinds_for(sig)) for r, T in enumerate(rows)]
  arena = torch.from_numpy(np.concatenate(recs)).to(DEV)
  out, out_rows = filters.resample_poly(arena, rows, up, down)
  out = out.cpu().numpy()
  at = 0
  for p, n in zip(recs, out_rows):
    want = signal.resample_poly(p, up, down, axis=0)
    assert want.shape == (n, 4)
    got = out[at:at + n]
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()
    assert got.tobytes() == want.tobytes()                # in fact bit-identical: same taps, same summation order
    at += n
  assert at == out.shape[0]


def test_filter_resample_window_pipeline_against_scipy_plus_oracle():
  """All extensions chained the way recordutil does when the optional params keys are present: band-pass the SCG
  columns, resample every column 500 -> 250 Hz, plan the chamber intervals at the new rate, window + normalise."""
  import scgrhc
  from oracle import scgrhc_oracle as orc
  sig = synth_ref.DEFAULT_SIG_NAMES
  T = 60000
  p = synth_ref.gen_record(H.SEED, 90, T, kinds=synth_ref.kinds_for(sig))
  meta = synth_ref.record_meta(120, events={'RA_1': 0, 'PA_1': 20.3, 'RV_1': 95})
  sos = signal.butter(4, (1.0, 40.0), btype='bandpass', fs=500, output='sos')
  arena = torch.from_numpy(p).to(DEV)
  f = filters.sosfiltfilt(arena, [T], sos, [0, 1, 2])
  r, rows = filters.resample_poly(f, [T], 250, 500)
  fs, W = 250, int(1.5 * 250)
  plan = scgrhc.plan_cohort([meta], 'PA', rows, W, fs=fs)
  st = scgrhc.prepare_windows(r, plan, [0, 1, 2], 3, -50.0)
  # the same chain with scipy + the oracle's per-window arithmetic
  q = p.copy()
  q[:, :3] = signal.sosfiltfilt(sos, p[:, :3], axis=0)
  q = signal.resample_poly(q, 250, 500, axis=0)
  a, b = int(20.3 * fs), int(95 * fs)
  n = (b - a) // W
  starts = a + np.arange(n) * W
  idx = starts[:, None] + np.arange(W)[None, :]
  rhc = q[:, 3][idx]
  keep = ~((orc.flat_count(rhc) >= 2) | (orc.r_squared(rhc) > 0.8) | orc.below_floor(rhc, -50.0))
  assert plan.n_cand == n and st.keep.cpu().numpy().astype(bool).tolist() == keep.tolist() and 0 < keep.sum() < n
  scg = q[:, :3][idx][keep]
  mm = scg.min(axis=(1, 2)), scg.max(axis=(1, 2))
  want = ((scg - mm[0][:, None, None]) / (mm[1] - mm[0] + 0.0001)[:, None, None]).transpose(0, 2, 1).astype(np.float32)
  assert st.materialise()[0].cpu().numpy().tobytes() == np.ascontiguousarray(want).tobytes()
