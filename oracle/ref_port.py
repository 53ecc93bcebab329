"""TEST / BASELINE INFRASTRUCTURE — per-window CPU port that keeps the reference's cost profile.

``bench.py --impl reference`` and the ``cpu_baseline`` leg time THIS module on the GPU box's host
cores, because ``/root/reference`` itself cannot travel there.  Unlike ``scgrhc_oracle.py`` (which is
vectorised over windows and therefore much faster than the reference), this port deliberately does
what the reference does per candidate window — a pandas rolling max/min, an sklearn
``LinearRegression`` fit + score, a Python loop over samples, numpy min/max + normalise + a torch
fp32 tensor per kept window — so its windows/s is the reference's (≈390/s/core).
Checked against the golden fixtures in tests/test_oracle.py.  Never imported by the product.
"""
import numpy as np
import pandas as pd
import torch
from sklearn.linear_model import LinearRegression

from . import scgrhc_oracle as orc


def window_has_noise(y, min_rhc):
  """waveform_noise.py:44-49 with the same third-party calls (pandas rolling, sklearn OLS)."""
  s = pd.Series(y)
  d = s.rolling(window=orc.FLAT_MIN_SAMPLES).max() - s.rolling(window=orc.FLAT_MIN_SAMPLES).min()
  if int((d < orc.FLAT_THRESHOLD).sum()) >= 2:          # == len(get_flat_lines(y)) > 0 (quirk, :13-26)
    return True
  x = np.arange(len(y)).reshape(-1, 1)
  if LinearRegression().fit(x, y).score(x, y) > orc.R2_THRESHOLD:
    return True
  for v in y:                                            # waveform_noise.py:38-40
    if v < min_rhc:
      return True
  return False


def prepare_record(p_signal, sig_name, meta, in_channels, chamber, segment_size, min_rhc):
  """get_segments + SCGDataset.init_segments for one record (recordutil.py:133-149,55-66).
  Returns the list of (scg f32 (C,W), rhc f32 (1,W), start, stop, (smin,smax), (rmin,rmax))."""
  W = int(segment_size * orc.SAMPLE_FREQ)
  cols = [list(sig_name).index(n) for n in in_channels]
  rcol = list(sig_name).index(orc.RHC_NAME)
  out = []
  n_cand = 0
  for a, b in orc.chamber_intervals(meta, chamber):
    scg_sig = p_signal[a:b, cols]
    rhc_sig = p_signal[a:b, [rcol]]
    for i in range(scg_sig.shape[0] // W):
      n_cand += 1
      scg = scg_sig[i * W:(i + 1) * W]
      rhc = rhc_sig[i * W:(i + 1) * W]
      if window_has_noise(rhc[:, 0], min_rhc):
        continue
      mm_s = (np.min(scg), np.max(scg))
      mm_r = (np.min(rhc), np.max(rhc))
      s = torch.tensor(orc.minmax_norm(scg, *mm_s).T, dtype=torch.float32)
      r = torch.tensor(orc.minmax_norm(rhc, *mm_r).T, dtype=torch.float32)
      out.append((s, r, i * W, i * W + W, mm_s, mm_r))
  return out, n_cand
