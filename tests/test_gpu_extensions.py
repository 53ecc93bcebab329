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
