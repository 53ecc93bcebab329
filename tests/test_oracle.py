"""CPU: the numpy oracle vs fixtures produced by the unmodified reference (tests/golden)."""
import json
import os

import numpy as np
import pytest

from oracle import scgrhc_oracle as orc
from tests import helpers as H


def test_intervals_match_reference():
  cases = H.load_json('intervals.json')
  for name, case in cases.items():
    for chamber, want in case['intervals'].items():
      got = orc.chamber_intervals(case['meta'], chamber)
      assert [list(x) for x in got] == want, (name, chamber)


def test_predicates_match_reference():
  g = np.load(os.path.join(H.GOLDEN, 'predicates.npz'))
  names, ys = H.predicate_inputs()
  assert H.sha(ys) == str(g['inputs_sha']), 'synthetic generator drifted from the fixture inputs'
  assert ((orc.flat_count(ys) >= 2) == g['flat']).all()
  assert (orc.is_straight_line(ys) == g['straight']).all()
  assert (~orc.below_floor(ys, -50) == g['in_range']).all()
  assert (np.array([orc.has_noise(y, -50) for y in ys]) == g['has_noise']).all()
  # the closed form tracks sklearn's R^2 far inside the decision margin
  ok = g['r2'] > 1e-6
  assert np.abs(orc.r_squared(ys)[ok] - g['r2'][ok]).max() < 1e-12
  # API parity of the quirky segment list
  segs = json.loads(str(g['adv_segments']))
  for want, y in zip(segs, ys):
    assert [tuple(s) for s in want] == orc.flat_segments(y)


def test_short_windows_and_constant_windows_match_reference():
  """Windows of 2..50 samples: no flat-line test can fire, so exactly constant windows are decided by sklearn's rounding
  noise (R^2 = 1.0 iff np.mean(y) is exact) — the oracle restates that literally."""
  g = H.load_json('short_windows.json')
  ys = H.short_windows()
  assert H.sha(np.concatenate(ys)) == g['inputs_sha'] and [len(y) for y in ys] == g['n']
  assert [bool(orc.is_straight_line(y)) for y in ys] == g['straight']
  assert [bool(orc.has_noise(y, -50)) for y in ys] == g['has_noise']
  assert any(s for y, s in zip(ys, g['straight']) if y.max() == y.min()) and any(not s for y, s in zip(ys, g['straight']) if y.max() == y.min())


def test_ambiguity_band_fixture():
  """R^2 planted at 0.8 +- 1e-13 .. 1e-3: the closed form tracks sklearn's score to ~1e-15, so on these inputs it
  still decides every window the way the reference does."""
  g = np.load(os.path.join(H.GOLDEN, 'ambiguity.npz'))
  assert (orc.is_straight_line(g['ys']) == g['straight']).all()
  assert np.abs(orc.r_squared(g['ys']) - g['r2']).max() < 5e-15


def test_flat_quirk_needs_two_positions():
  names, ys = H.predicate_inputs()
  by = dict(zip(names, ys))
  assert orc.flat_count(by['run50@350']) == 1 and not (orc.flat_count(by['run50@350']) >= 2)
  assert orc.flat_count(by['run51@350']) == 2
  assert orc.flat_count(by['run49@350']) == 0


def test_nonfinite_raises_like_reference():
  y = np.linspace(0, 30, 750) ** 1.5 + np.sin(np.arange(750))
  y[10] = np.nan
  with pytest.raises(ValueError):
    orc.has_noise(y, -50)
  y[100:200] = 3.0          # flat line short-circuits before sklearn sees the NaN
  assert orc.has_noise(y, -50) is True


@pytest.mark.parametrize('cfg', ['waveform_06', 'waveform_10', 'waveform_11', 'waveform_23', 'waveform_19', 'waveform_15'])
def test_record_small_bit_exact(cfg):
  g = np.load(os.path.join(H.GOLDEN, 'record_small.npz'))
  sig, p, meta = H.small_record()
  c = H.effective_config(cfg)
  rw = orc.scan_record(p, sig, meta, c['in_channels'], c['chamber'], c['segment_size'], c['min_RHC'])
  k = np.nonzero(rw.keep)[0]
  assert (rw.rel_start[k] == g[cfg + '.start']).all()
  assert (rw.rel_start[k] + rw.W == g[cfg + '.stop']).all()
  scg, rhc, mm = orc.normalise_record(p, sig, c['in_channels'], rw)
  assert (mm == g[cfg + '.minmax']).all()
  assert scg.tobytes() == g[cfg + '.scg'].tobytes()
  assert rhc.tobytes() == g[cfg + '.rhc'].tobytes()


def test_records_full_all_configs():
  full = H.load_json('records_full.json')
  table = H.configs()
  recs = {('rec%d' % r): H.full_record(r) for r in (0, 1)}
  for name, (sig, p, meta) in recs.items():
    assert H.sha(p) == full['record_sha'][name]
  assert len(full['configs']) == 36
  for cfg, entry in full['configs'].items():
    c = H.effective_config(cfg, table)
    scans = {n: orc.scan_record(p, sig, meta, c['in_channels'], c['chamber'], c['segment_size'], c['min_RHC'])
             for n, (sig, p, meta) in recs.items()}
    gmm = None
    if c['use_global_min_max']:
      gmm = orc.global_minmax(np.concatenate([rw.minmax[rw.keep] for rw in scans.values()]))
      assert [float(v).hex() for v in gmm] == entry['global_minmax_hex']
    for n, (sig, p, meta) in recs.items():
      rw, want = scans[n], entry['records'][n]
      k = np.nonzero(rw.keep)[0]
      assert rw.rel_start[k].tolist() == want['start'], (cfg, n)
      assert (rw.rel_start[k] + rw.W).tolist() == want['stop'], (cfg, n)
      scg, rhc, mm = orc.normalise_record(p, sig, c['in_channels'], rw, gmm)
      assert H.sha(mm) == want['minmax_sha'], (cfg, n)
      assert H.sha(scg) == want['scg_sha'], (cfg, n)
      assert H.sha(rhc) == want['rhc_sha'], (cfg, n)


def test_philox_known_answer_vectors():
  """Random123 kat_vectors for philox4x32_10 pin the host generator the device stream is compared with."""
  from oracle import philox_ref
  kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
         ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
         ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
  for ctr, key, want in kat:
    got = philox_ref.philox4x32_10(np.array([ctr], dtype=np.uint64), key)[0]
    assert tuple(int(v) for v in got) == want
  z = philox_ref.normals(7, 3, 200000)
  assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
