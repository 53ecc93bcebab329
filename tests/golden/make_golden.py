"""Generates the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Inputs are produced by ``oracle/synth_ref.py`` (integer hashing + exact fp64 ops, so they are
reproducible from code); only the reference's OUTPUTS are stored:

  intervals.json      get_chamber_intervals on hand-written JSON side-cars      (recordutil.py:93-110)
  predicates.npz      get_flat_lines / is_straight_line / in_rhc_range / has_noise
                      on adversarial + synthetic windows                        (waveform_noise.py:6-49)
  record_small.npz    get_segments + SCGDataset on a short record, full tensors (recordutil.py:55-66,122-149)
  configs.json        path-relevant keys of the 37 params.json + the reference loader's verdict
  short_windows.json  is_straight_line / has_noise on windows of 2..50 samples (exact constants included)
  ambiguity.npz       windows with R^2 = 0.8 +- 1e-13..1e-3: inputs, the reference's verdict and score
  records_full.json   the same on a 10-min record for all 37 waveform_NN configs:
                      ordered (start, stop) lists + sha256 of the fp32 tensors and min/max pairs
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import synth_ref  # noqa: E402
from oracle.ref_harness import ReferenceHarness  # noqa: E402

SEED = 0x5C6
LEGACY_DEFAULTS = dict(min_RHC=float('-inf'), use_global_min_max=False)


def sha(a):
  return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---- hand-written side-cars for the interval maths ------------------------------------
INTERVAL_CASES = {
  'oracle_layout': synth_ref.record_meta(600),
  'unsorted': {'MacStTime': '3/4/2021 09:15:10', 'MacEndTime': '3/4/2021 09:31:55',
               'ChamEvents_in_s': {'PA_1': 400.5, 'RA_1': 12.25, 'RV_1': 180, 'PA_2': 700.004, 'PCW_1': 650.1}},
  'duplicates_ties': {'MacStTime': 'x 00:00:00', 'MacEndTime': 'x 00:20:00',
                      'ChamEvents_in_s': {'PA_1': 100, 'PA_2': 100, 'RV_1': 100, 'PA_3': 300.0, 'RA_1': 1199.999}},
  'non_integral': {'MacStTime': 'd 10:00:00', 'MacEndTime': 'd 10:10:01',
                   'ChamEvents_in_s': {'RA_1': 0.001, 'RV_1': 0.0029999, 'PA_1': 33.3333333, 'PCW_1': 77.7777777,
                                       'PA_2': 123.456789, 'RV_2': 599.9999999}},
  'event_after_end': {'MacStTime': 'd 10:00:00', 'MacEndTime': 'd 10:05:00',
                      'ChamEvents_in_s': {'RA_1': 10, 'PA_1': 200, 'RV_1': 450}},
  'not_a_dict': {'MacStTime': 'd 10:00:00', 'MacEndTime': 'd 10:05:00', 'ChamEvents_in_s': []},
  'nan_string': {'MacStTime': 'd 10:00:00', 'MacEndTime': 'd 10:05:00', 'ChamEvents_in_s': 'NaN'},
  'end_before_start': {'MacStTime': 'd 23:58:00', 'MacEndTime': 'd 00:04:00',
                       'ChamEvents_in_s': {'RA_1': 0, 'PA_1': 100}},
  'no_prefix_match': {'MacStTime': 'd 10:00:00', 'MacEndTime': 'd 10:05:00',
                      'ChamEvents_in_s': {'PAX_1': 5, 'pa_1': 50, 'PA': 100, 'PA_': 200}},
}
CHAMBERS = ['RA', 'RV', 'PA', 'PCW']


def adversarial_windows():
  """Deterministic 750-sample RHC windows that sit on every decision boundary."""
  key, finc = synth_ref.record_params(SEED, 12345)
  t = np.arange(750, dtype=np.int64)
  base = synth_ref._rhc_base(key, finc, t)
  out, names = [], []

  def add(name, y):
    names.append(name)
    out.append(np.asarray(y, dtype=np.float64).copy())

  add('clean', base)
  for L in (48, 49, 50, 51, 52, 100, 701, 750):
    for a in (0, 1, 350, 750 - L):
      if a + L > 750:
        continue
      y = base.copy()
      y[a:a + L] = y[a]
      add('run%d@%d' % (L, a), y)
  # two separate 50-runs: count == 2 from non-adjacent positions
  y = base.copy(); y[10:60] = y[10]; y[400:450] = y[400]; add('two_runs50', y)
  # range exactly at / one ulp either side of the threshold over 51 samples
  for name, top in (('range_eq', 1e-3), ('range_below', np.nextafter(1e-3, 0)), ('range_above', np.nextafter(1e-3, 1))):
    y = base.copy(); y[200:251] = 20.0; y[225] = 20.0 + top; add(name, y)
  # pressure floor: exactly -50, one ulp below, one ulp above; NaN-free
  for name, v in (('floor_eq', -50.0), ('floor_below', np.nextafter(-50.0, -np.inf)), ('floor_above', np.nextafter(-50.0, 0))):
    y = base.copy(); y[375] = v; add(name, y)
  # straight lines: exact, constant, and noisy lines with R^2 straddling 0.8
  add('line_exact', 3.0 + 0.01 * t)
  add('constant', np.full(750, 17.25))
  add('two_level', np.where(t < 375, 10.0, 11.0))
  nz = synth_ref.noise(key, 3, t)
  for A in (1.0, 1.8, 2.0, 2.1, 2.15, 2.2, 2.25, 2.3, 2.5, 3.0, 4.0):
    add('noisy_line_A%g' % A, (5.0 + 0.02 * t) + A * nz)
  add('decreasing', 40.0 - 0.03 * t + 0.5 * nz)
  return names, np.stack(out)


def run_predicates(h):
  wn = h.waveform_noise
  names, adv = adversarial_windows()
  # plus every grid window of one synthetic record (defects planted)
  rec = synth_ref.gen_record(SEED, 7, 150000, kinds=(3,))[:, 0]
  syn = rec.reshape(-1, 750)
  ys = np.concatenate([adv, syn])
  p50 = h.params('waveform_06')
  flat = np.array([len(wn.get_flat_lines(y)) > 0 for y in ys])
  nseg = np.array([len(wn.get_flat_lines(y)) for y in ys])
  straight = np.array([bool(wn.is_straight_line(y)) for y in ys])
  inrange = np.array([bool(wn.in_rhc_range(p50, y)) for y in ys])
  noisy = np.array([bool(wn.has_noise(p50, y)) for y in ys])
  # the R^2 sklearn actually produced (for the ambiguity-band test)
  from sklearn.linear_model import LinearRegression
  x = np.arange(750).reshape(-1, 1)
  r2 = np.array([LinearRegression().fit(x, y).score(x, y) for y in ys])
  # flat-segment tuples for the adversarial set (API parity of get_flat_lines)
  segs = [[(int(a), int(b)) for a, b in wn.get_flat_lines(y)] for y in adv]
  np.savez_compressed(os.path.join(HERE, 'predicates.npz'),
                      names=np.array(names), n_adv=len(adv), inputs_sha=sha(ys),
                      flat=flat, n_flat_segments=nseg, straight=straight, in_range=inrange,
                      has_noise=noisy, r2=r2, adv_segments=json.dumps(segs))
  print('predicates: %d windows, flat=%d straight=%d below=%d noisy=%d' %
        (len(ys), flat.sum(), straight.sum(), (~inrange).sum(), noisy.sum()))


def run_intervals(h):
  ru = h.recordutil
  res = {}
  for name, meta in INTERVAL_CASES.items():
    h.add_record(name, ['RHC_pressure'], np.zeros((1, 1)), meta)
    res[name] = {'meta': meta, 'intervals': {c: ru.get_chamber_intervals(name, c) for c in CHAMBERS}}
  with open(os.path.join(HERE, 'intervals.json'), 'w') as f:
    json.dump(res, f, indent=1)
  print('intervals: %d cases' % len(res))


def dataset_arrays(h, params, segments, mm_scg=None, mm_rhc=None):
  ds = h.recordutil.SCGDataset(list(segments), params.segment_size, mm_scg, mm_rhc)
  n = len(ds)
  C = len(params.in_channels)
  scg = np.stack([ds[i][0].numpy() for i in range(n)]) if n else np.zeros((0, C, 750), np.float32)
  rhc = np.stack([ds[i][1].numpy() for i in range(n)]) if n else np.zeros((0, 1, 750), np.float32)
  mm = np.array([[ds[i][5][0], ds[i][5][1], ds[i][6][0], ds[i][6][1]] for i in range(n)], dtype=np.float64).reshape(n, 4)
  return scg, rhc, mm


def run_record_small(h):
  """Short record (T=45,000 rows, 5 signals), events squeezed so every chamber has windows;
  one interval is not window-aligned and one runs past the end of the record."""
  T = 45000
  sig = synth_ref.SIG_NAMES_5
  p = synth_ref.gen_record(SEED, 3, T, kinds=synth_ref.kinds_for(sig))
  meta = synth_ref.record_meta(100, events={'RA_1': 0.5, 'RV_1': 15.001, 'PA_1': 30, 'PCW_1': 55.25, 'PA_2': 61.7})
  h.add_record('small', sig, p, meta)
  out = {}
  for cfg in ('waveform_06', 'waveform_10', 'waveform_11', 'waveform_23', 'waveform_19', 'waveform_15'):
    params = h.params(cfg)
    segs = h.recordutil.get_segments(params, record_name='small')
    starts = np.array([s[3] for s in segs], dtype=np.int64)
    stops = np.array([s[4] for s in segs], dtype=np.int64)
    scg, rhc, mm = dataset_arrays(h, params, segs)
    out[cfg + '.start'] = starts
    out[cfg + '.stop'] = stops
    out[cfg + '.scg'] = scg
    out[cfg + '.rhc'] = rhc
    out[cfg + '.minmax'] = mm
    print('record_small %s: kept %d' % (cfg, len(starts)))
  np.savez_compressed(os.path.join(HERE, 'record_small.npz'), **out)


def run_records_full(h):
  T = 300000
  sig = synth_ref.SIG_NAMES_5
  recs = {}
  for r in (0, 1):
    p = synth_ref.gen_record(SEED, r, T, kinds=synth_ref.kinds_for(sig))
    h.add_record('rec%d' % r, sig, p, synth_ref.record_meta(600))
    recs['rec%d' % r] = sha(p)
  res = {'seed': SEED, 'T': T, 'sig_name': sig, 'record_sha': recs, 'configs': {}}
  for nn in range(1, 38):
    cfg = 'waveform_%02d' % nn
    legacy = nn <= 5
    over = {}
    if legacy:
      over = dict(LEGACY_DEFAULTS)
      if nn == 1:
        continue          # no `chamber` key and no code in the reference that could run it
      if nn == 4:
        over = dict(min_RHC=0, use_global_min_max=True)       # project_log.txt:19-21
    params = h.params(cfg, **over)
    entry = {'legacy_defaults': {k: (str(v) if isinstance(v, float) else v) for k, v in over.items()}, 'records': {}}
    all_segs = []
    for name in ('rec0', 'rec1'):
      segs = h.recordutil.get_segments(params, record_name=name)
      all_segs.append(segs)
    gmm = None
    if params.use_global_min_max:
      mm_scg, mm_rhc = h.recordutil.get_global_minmax_vals(all_segs[0] + all_segs[1])
      gmm = [float(mm_scg[0]), float(mm_scg[1]), float(mm_rhc[0]), float(mm_rhc[1])]
      entry['global_minmax_hex'] = [float(v).hex() for v in gmm]
    for name, segs in zip(('rec0', 'rec1'), all_segs):
      starts = [int(s[3]) for s in segs]
      stops = [int(s[4]) for s in segs]
      if gmm is None:
        scg, rhc, mm = dataset_arrays(h, params, segs)
      else:
        scg, rhc, mm = dataset_arrays(h, params, segs, (gmm[0], gmm[1]), (gmm[2], gmm[3]))
      entry['records'][name] = {'start': starts, 'stop': stops, 'scg_sha': sha(scg), 'rhc_sha': sha(rhc),
                                'minmax_sha': sha(mm), 'n': len(starts)}
    res['configs'][cfg] = entry
    print('records_full %s: kept %s' % (cfg, [entry['records'][n]['n'] for n in ('rec0', 'rec1')]))
  with open(os.path.join(HERE, 'records_full.json'), 'w') as f:
    json.dump(res, f)


def run_configs(h):
  """Path-relevant keys of all 37 params.json + whether the reference's own Params loads them
  (SURVEY.md §5a).  A condensed table, not a copy of the files."""
  keys = ('in_channels', 'chamber', 'segment_size', 'batch_size', 'min_RHC', 'use_global_min_max')
  res = {}
  for nn in range(1, 38):
    cfg = 'waveform_%02d' % nn
    path = os.path.join(h.paramutil.__file__.rsplit(os.sep, 1)[0], cfg, 'params.json')
    with open(path) as f:
      data = json.load(f)
    try:
      h.paramutil.Params(path)
      err = None
    except KeyError as e:
      err = 'KeyError %s' % e
    res[cfg] = {k: data[k] for k in keys if k in data}
    res[cfg]['all_keys'] = sorted(data.keys())
    res[cfg]['reference_params_error'] = err
  with open(os.path.join(HERE, 'configs.json'), 'w') as f:
    json.dump(res, f, indent=1)
  print('configs: %d (%d load in the reference)' % (len(res), sum(v['reference_params_error'] is None for v in res.values())))


def short_windows():
  """Windows shorter than 51 samples (no flat-line test can fire): exact constants, ramps, noise."""
  key, finc = synth_ref.record_params(SEED, 4321)
  out = []
  for n in (2, 3, 7, 8, 9, 16, 30, 49, 50):
    t = np.arange(n, dtype=np.int64)
    nz = synth_ref.noise(key, 3, t)
    for v in (0.1, 0.3, 17.25, 1.0 / 3.0, -2.7, 1e-3, 25.123456789, -50.0, -50.000001, 0.0):
      out.append(np.full(n, v))
    out.append(3.0 + 0.01 * t)
    out.append(20.0 + nz)
    out.append(5.0 + 0.05 * t + 0.3 * nz)
    out.append(5.0 + 0.05 * t + 1.0 * nz)
  return out


def run_short_windows(h):
  """is_straight_line / has_noise of the reference on windows of 2..50 samples (segment_size < 0.102 s): exactly constant
  windows are decided by sklearn's rounding noise there (R^2 = 1.0 iff the mean is exact), nothing else rejects them."""
  wn = h.waveform_noise
  p50 = h.params('waveform_06')
  ys = short_windows()
  res = {'inputs_sha': sha(np.concatenate(ys)), 'n': [len(y) for y in ys],
         'straight': [bool(wn.is_straight_line(y)) for y in ys], 'has_noise': [bool(wn.has_noise(p50, y)) for y in ys]}
  with open(os.path.join(HERE, 'short_windows.json'), 'w') as f:
    json.dump(res, f)
  print('short_windows: %d windows, straight=%d noisy=%d' % (len(ys), sum(res['straight']), sum(res['has_noise'])))


def run_ambiguity(h):
  """750-sample windows whose R^2 sits at 0.8 +- {1e-13 .. 1e-3}: the reference's own verdict (sklearn lstsq + r2_score)
  and score, with the inputs stored (they come out of dot products, which are not bit-reproducible across BLAS builds)."""
  from sklearn.linear_model import LinearRegression
  wn = h.waveform_noise
  key, _ = synth_ref.record_params(SEED, 999)
  t = np.arange(750, dtype=np.float64)
  x = t - t.mean()
  xh = x / np.sqrt((x * x).sum())
  e = synth_ref.noise(key, 3, np.arange(750, dtype=np.int64))
  e = e - e.mean()
  e = e - (e * xh).sum() * xh
  eh = e / np.sqrt((e * e).sum())
  ys, deltas = [], []
  for d in (0.0, 1e-13, -1e-13, 1e-12, -1e-12, 1e-11, -1e-11, 1e-10, -1e-10, 3e-9, -3e-9, 1e-6, -1e-6, 1e-3, -1e-3):
    r2 = 0.8 + d
    ys.append(12.5 + 40.0 * (np.sqrt(r2) * xh + np.sqrt(1.0 - r2) * eh))
    deltas.append(d)
  ys = np.stack(ys)
  X = np.arange(750).reshape(-1, 1)
  np.savez_compressed(os.path.join(HERE, 'ambiguity.npz'), ys=ys, deltas=np.array(deltas),
                      straight=np.array([bool(wn.is_straight_line(y)) for y in ys]),
                      r2=np.array([LinearRegression().fit(X, y).score(X, y) for y in ys]))
  print('ambiguity: %d windows' % len(ys))


def run_reference_pickle(h):
  """A DataLoader pickled by the reference itself (recordutil.py:198-209): the drop-in's load_dataloader must read it."""
  import pickle
  from torch.utils.data import DataLoader
  sig = synth_ref.SIG_NAMES_5
  p = synth_ref.gen_record(SEED, 3, 45000, kinds=synth_ref.kinds_for(sig))
  meta = synth_ref.record_meta(100, events={'RA_1': 0.5, 'RV_1': 15.001, 'PA_1': 30, 'PCW_1': 55.25, 'PA_2': 61.7})
  h.add_record('small', sig, p, meta)
  params = h.params('waveform_19')
  segs = h.recordutil.get_segments(params, record_name='small')
  ds = h.recordutil.SCGDataset(segs, params.segment_size, None, None)
  with open(os.path.join(HERE, 'reference_loader.pickle'), 'wb') as f:
    pickle.dump(DataLoader(ds, batch_size=2, shuffle=True), f)
  print('reference_loader.pickle: %d items' % len(ds))


if __name__ == '__main__':
  if len(sys.argv) > 1 and sys.argv[1] == 'pickle':
    with ReferenceHarness() as h:
      run_reference_pickle(h)
    sys.exit(0)
  if len(sys.argv) > 1 and sys.argv[1] == 'short':
    with ReferenceHarness() as h:
      run_short_windows(h)
      run_ambiguity(h)
    sys.exit(0)
  with ReferenceHarness() as h:
    run_reference_pickle(h)
    run_configs(h)
    run_intervals(h)
    run_predicates(h)
    run_record_small(h)
    run_records_full(h)
    run_short_windows(h)
    run_ambiguity(h)
