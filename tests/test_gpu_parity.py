"""GPU parity tests: the CUDA path (through the torch ops -> C ABI) vs golden fixtures produced by the
unmodified reference and vs the numpy oracle on the same seeded inputs.

Bar: keep/reject decisions, window indices, min/max pairs and fp32/fp64 samples are all BIT-EXACT
(the device evaluates the reference's fp64 operations with the same roundings), which is stricter
than the rel 1e-5 (fp32) / 1e-10 (fp64) tolerance BASELINE.json asks for.
"""
import os

import numpy as np
import pytest
import torch

from oracle import scgrhc_oracle as orc
from oracle import synth_ref
from tests import helpers as H

pytestmark = pytest.mark.gpu

import scgrhc  # noqa: E402
from scgrhc import _native as N  # noqa: E402
from scgrhc import ops  # noqa: E402

DEV = 'cuda:0'


def run_record(p, sig, meta, c, out_dtype=torch.float32, global_mm=False, keep_all=False):
  W = int(c['segment_size'] * 500)
  cols, rcol = scgrhc.resolve_columns(sig, c['in_channels'])
  plan = scgrhc.plan_cohort([meta], c['chamber'], [p.shape[0]], W)
  arena = torch.from_numpy(np.ascontiguousarray(p)).to(DEV)
  return plan, scgrhc.prepare_windows(arena, plan, cols, rcol, c['min_RHC'], use_global_min_max=global_mm,
                                      out_dtype=out_dtype, keep_all=keep_all)


def test_synth_matches_numpy_twin_bit_for_bit():
  for sig in (synth_ref.DEFAULT_SIG_NAMES, synth_ref.SIG_NAMES_5, ['a', 'RHC_pressure', 'b']):
    kinds = synth_ref.kinds_for(sig)
    T, n_rec = 6000, 3
    out = torch.empty((n_rec * T, len(kinds)), dtype=torch.float64, device=DEV)
    ops.synth_records(out, H.SEED, 5, n_rec, T, list(kinds), 16, 750)
    got = out.cpu().numpy()
    want = np.concatenate([synth_ref.gen_record(H.SEED, 5 + r, T, kinds=kinds) for r in range(n_rec)])
    assert got.tobytes() == want.tobytes()


def test_division_by_reciprocal_is_correctly_rounded():
  for mode in (0, 1):
    bad64, bad32, n_fast = ops.selftest_div(0, 1234 + mode, 1 << 28, mode)
    assert bad64 == 0 and bad32 == 0, (mode, bad64, bad32)
    # mode 0 is shaped like real windows and must stay on the reciprocal path; mode 1 mixes in the IEEE loop
    assert n_fast > (0.99 if mode == 0 else 0.01) * (1 << 28), (mode, n_fast)
  # fp32 tier: every element the integer check lets through must round to the reference's float;
  # elements planted on float rounding boundaries must be flagged (and only about that many)
  ineligible, bad32, flagged = ops.selftest_div(0, 77, 1 << 28, 2)
  assert bad32 == 0 and ineligible == 0, (bad32, ineligible)
  assert 1e-3 * (1 << 28) < flagged < 0.55 * (1 << 28), flagged   # planted boundaries survive rounding of x only in part
  # the same with minima down to 2^-86 and samples a few ulp above the minimum (the smallest non-zero quotients): eligibility
  # is decided by the exponents of min and range, and most of these windows still qualify
  ineligible, bad32, flagged = ops.selftest_div(0, 78, 1 << 28, 3)
  assert bad32 == 0 and ineligible < 0.35 * (1 << 28), (bad32, ineligible)


def test_predicates_match_reference_golden():
  g = np.load(os.path.join(H.GOLDEN, 'predicates.npz'))
  names, ys = H.predicate_inputs()
  assert H.sha(ys) == str(g['inputs_sha'])
  n = len(ys)
  arena = torch.from_numpy(ys.reshape(-1, 1).copy()).to(DEV)      # nsig = 1: the RHC column is also the SCG column
  plan = scgrhc.Plan(np.array([(0, 0, n, 0)], dtype=scgrhc.engine.INTERVAL_DTYPE), n, 750)
  st = scgrhc.prepare_windows(arena, plan, [0], 0, -50.0, predicates_only=True, check=False)
  reason = st.reason.cpu().numpy()
  keep = st.keep.cpu().numpy().astype(bool)
  assert (((reason & N.REASON_FLAT) != 0) == g['flat']).all()
  # exactly-constant windows included: sklearn's R^2 is rounding noise there (1.0 iff np.mean(y) is exact)
  assert (((reason & N.REASON_STRAIGHT) != 0) == g['straight']).all()
  assert (((reason & N.REASON_FLOOR) == 0) == g['in_range']).all()
  assert (keep == ~g['has_noise']).all()
  assert (reason & N.REASON_AMBIGUOUS).sum() == 0
  mm = st.minmax.cpu().numpy()
  assert (mm[:, 2] == ys.min(axis=1)).all() and (mm[:, 3] == ys.max(axis=1)).all()


def test_short_and_constant_windows_match_reference_golden():
  """Windows of 2..50 samples (no flat-line test can fire): the fused kernel decides exactly constant windows the way
  sklearn's rounding noise does (R^2 = 1.0 iff np.mean(y) is exact, numpy's pairwise summation order), so get_segments
  and the standalone is_straight_line()/has_noise() cannot disagree for small segment_size."""
  import waveform_noise
  import types
  g = H.load_json('short_windows.json')
  ys = H.short_windows()
  assert H.sha(np.concatenate(ys)) == g['inputs_sha']
  got_s, got_n = {}, {}
  for L in sorted(set(g['n'])):
    idx = [i for i, n in enumerate(g['n']) if n == L]
    block = np.concatenate([ys[i] for i in idx]).reshape(-1, 1)
    plan = scgrhc.Plan(np.array([(0, 0, len(idx), 0)], dtype=scgrhc.engine.INTERVAL_DTYPE), len(idx), L)
    st = scgrhc.prepare_windows(torch.from_numpy(block.copy()).to(DEV), plan, [0], 0, -50.0, predicates_only=True)
    reason, keep = st.reason.cpu().numpy(), st.keep.cpu().numpy().astype(bool)
    for j, i in enumerate(idx):
      got_s[i], got_n[i] = bool(reason[j] & N.REASON_STRAIGHT), not keep[j]
  assert [got_s[i] for i in range(len(ys))] == g['straight']
  assert [got_n[i] for i in range(len(ys))] == g['has_noise']
  p = types.SimpleNamespace(min_RHC=-50)
  for i in range(0, len(ys), 5):                    # the standalone API gives the same answers
    assert waveform_noise.is_straight_line(ys[i]) == g['straight'][i] and waveform_noise.has_noise(p, ys[i]) == g['has_noise'][i]


def test_ambiguous_windows_are_counted():
  """R^2 planted at 0.8 +- 1e-13 .. 1e-3 (inputs and the reference's verdicts stored in the fixture): windows within 1e-9
  of the threshold carry REASON_AMBIGUOUS and are counted on the WindowStore; here the closed form still agrees with
  sklearn on every one of them."""
  g = np.load(os.path.join(H.GOLDEN, 'ambiguity.npz'))
  ys, n = g['ys'], len(g['ys'])
  plan = scgrhc.Plan(np.array([(0, 0, n, 0)], dtype=scgrhc.engine.INTERVAL_DTYPE), n, 750)
  st = scgrhc.prepare_windows(torch.from_numpy(ys.reshape(-1, 1).copy()).to(DEV), plan, [0], 0, -50.0, predicates_only=True)
  reason = st.reason.cpu().numpy()
  assert (((reason & N.REASON_STRAIGHT) != 0) == g['straight']).all()
  amb = (reason & N.REASON_AMBIGUOUS) != 0
  assert (amb == (np.abs(g['deltas']) < 5e-10)).all()
  assert st.n_ambiguous == int(amb.sum()) == 9
  clean = scgrhc.prepare_windows(torch.from_numpy(ys[-2:].reshape(-1, 1).copy()).to(DEV),
                                 scgrhc.Plan(np.array([(0, 0, 2, 0)], dtype=scgrhc.engine.INTERVAL_DTYPE), 2, 750), [0], 0, -50.0)
  assert clean.n_ambiguous == 0


@pytest.mark.parametrize('cfg', ['waveform_06', 'waveform_10', 'waveform_11', 'waveform_23', 'waveform_19', 'waveform_15'])
def test_record_small_golden_bit_exact(cfg):
  g = np.load(os.path.join(H.GOLDEN, 'record_small.npz'))
  sig, p, meta = H.small_record()
  c = H.effective_config(cfg)
  plan, st = run_record(p, sig, meta, c)
  assert st.start_idx.cpu().numpy().tolist() == g[cfg + '.start'].tolist()
  assert st.stop_idx.cpu().numpy().tolist() == g[cfg + '.stop'].tolist()
  assert (st.kept_minmax().cpu().numpy() == g[cfg + '.minmax']).all()
  scg, rhc = st.materialise()
  assert scg.cpu().numpy().tobytes() == g[cfg + '.scg'].tobytes()
  assert rhc.cpu().numpy().tobytes() == g[cfg + '.rhc'].tobytes()


@pytest.mark.parametrize('fan_out', [True, False])
def test_records_full_golden_all_configs(fan_out):
  """All 36 runnable configs as ONE sweep job over a device-resident cohort (BASELINE configs[4]): with the fan-out pass
  (one predicate pass + one normalisation pass per chamber for its 8 channel subsets) and with one fused pass per config."""
  import types
  from scgrhc import sweep
  full = H.load_json('records_full.json')
  table = H.configs()
  recs = {('rec%d' % r): H.full_record(r) for r in (0, 1)}
  sig = recs['rec0'][0]
  metas = [recs['rec0'][2], recs['rec1'][2]]
  arena = torch.from_numpy(np.concatenate([recs['rec0'][1], recs['rec1'][1]])).to(DEV)
  configs = {cfg: types.SimpleNamespace(**H.effective_config(cfg, table)) for cfg in full['configs']}
  assert len(configs) == 36
  seen = 0
  dense = 0
  for cfg, st in sweep.iter_sweep(arena, sig, metas, [300000, 300000], configs, buffers={}, fan_out=fan_out):
    entry, c = full['configs'][cfg], configs[cfg]
    dense += bool(st.minmax_dense)
    if c.use_global_min_max:
      assert [float(v).hex() for v in st.global_minmax.cpu().tolist()] == entry['global_minmax_hex']
    scg, rhc = st.materialise()
    scg, rhc = scg.cpu().numpy(), rhc.cpu().numpy()
    rec_id = st.rec_id.cpu().numpy()
    start, stop = st.start_idx.cpu().numpy(), st.stop_idx.cpu().numpy()
    mm = st.kept_minmax().cpu().numpy()
    for r, name in enumerate(('rec0', 'rec1')):
      want = entry['records'][name]
      m = rec_id == r
      assert start[m].tolist() == want['start'], (cfg, name)
      assert stop[m].tolist() == want['stop'], (cfg, name)
      assert H.sha(mm[m]) == want['minmax_sha'], (cfg, name)
      assert H.sha(scg[m]) == want['scg_sha'], (cfg, name)
      assert H.sha(rhc[m]) == want['rhc_sha'], (cfg, name)
    seen += 1
  assert seen == 36
  # 4 chambers x 8 channel subsets + the legacy 02/03/05 (same channels, no pressure floor) went through the fan-out pass;
  # waveform_04 (dataset-level min/max) takes its own two passes
  assert dense == (35 if fan_out else 0)


@pytest.mark.parametrize('nsig,out_dtype', [(4, torch.float32), (4, torch.float64), (5, torch.float32), (7, torch.float64)])
def test_cohort_vs_oracle(nsig, out_dtype):
  """20 seeded records: device cohort generated on the GPU, oracle on the numpy twin."""
  sig = (synth_ref.SIG_NAMES_5 + ['x5', 'x6'])[:nsig] if nsig != 4 else synth_ref.DEFAULT_SIG_NAMES
  kinds = synth_ref.kinds_for(sig)
  n_rec, T = 20, 60000
  meta = synth_ref.record_meta(120, events={'RA_1': 0, 'PA_1': 10.002, 'RV_1': 50, 'PA_2': 70.5})
  chans = ['patch_ACC_dv', 'patch_ACC_lat'] if nsig != 4 else ['patch_ACC_lat', 'patch_ACC_hf', 'patch_ACC_dv']
  cols, rcol = scgrhc.resolve_columns(sig, chans)
  arena = torch.empty((n_rec * T, nsig), dtype=torch.float64, device=DEV)
  ops.synth_records(arena, 99, 0, n_rec, T, list(kinds), 16, 750)
  plan = scgrhc.plan_uniform(meta, 'PA', T, 750, n_rec)
  st = scgrhc.prepare_windows(arena, plan, cols, rcol, -50.0, out_dtype=out_dtype)
  scg, rhc = st.materialise()
  scg, rhc = scg.cpu().numpy(), rhc.cpu().numpy()
  rec_id, start = st.rec_id.cpu().numpy(), st.start_idx.cpu().numpy()
  mm = st.kept_minmax().cpu().numpy()
  npdt = np.float64 if out_dtype == torch.float64 else np.float32
  total = 0
  for r in range(n_rec):
    p = synth_ref.gen_record(99, r, T, kinds=kinds)
    rw = orc.scan_record(p, sig, meta, chans, 'PA', 1.5, -50.0)
    k = np.nonzero(rw.keep)[0]
    m = rec_id == r
    assert start[m].tolist() == rw.rel_start[k].tolist(), r
    s_o, r_o, mm_o = orc.normalise_record(p, sig, chans, rw, out_dtype=npdt)
    assert (mm[m] == mm_o).all()
    assert scg[m].tobytes() == s_o.tobytes(), r
    assert rhc[m].tobytes() == r_o.tobytes(), r
    total += len(k)
  assert total == st.n_kept and 0 < total < plan.n_cand


def _same_store(a, b, check_windows=True):
  """Two WindowStores hold the same decisions, kept list, pairs and (bit for bit) windows."""
  assert a.n_kept == b.n_kept and torch.equal(a.keep, b.keep) and torch.equal(a.reason, b.reason)
  assert torch.equal(a.kept_idx, b.kept_idx) and torch.equal(a.start_idx, b.start_idx) and torch.equal(a.rec_id, b.rec_id)
  ka, kb = a.kept_minmax().cpu().numpy(), b.kept_minmax().cpu().numpy()
  assert ka.tobytes() == kb.tobytes() or (np.array_equal(np.isnan(ka), np.isnan(kb)) and np.array_equal(ka[~np.isnan(ka)], kb[~np.isnan(kb)]))
  if not a.dense:
    assert torch.equal(a.minmax[:, 2:], b.minmax[:, 2:])                     # the RHC pair of every candidate
  if check_windows and a.scg is not None:
    xa, xb = a.materialise(), b.materialise()
    assert xa[0].cpu().numpy().tobytes() == xb[0].cpu().numpy().tobytes()
    assert xa[1].cpu().numpy().tobytes() == xb[1].cpu().numpy().tobytes()


@pytest.mark.parametrize('nsig,chans,W,out_dtype', [
    (4, ['patch_ACC_lat', 'patch_ACC_hf', 'patch_ACC_dv'], 750, torch.float32),
    (4, ['patch_ACC_lat', 'patch_ACC_hf', 'patch_ACC_dv'], 750, torch.float64),
    (5, ['patch_ACC_dv'], 750, torch.float32), (5, ['patch_ACC_hf', 'patch_ECG'], 375, torch.float32),
    (5, ['patch_ACC_lat', 'patch_ACC_hf', 'patch_ACC_dv', 'patch_ECG'], 333, torch.float32),
    (7, ['patch_ACC_lat', 'patch_ACC_hf', 'patch_ACC_dv', 'patch_ECG'], 1024, torch.float64),
    (4, ['patch_ACC_dv', 'patch_ACC_lat'], 1000, torch.float32), (4, ['patch_ACC_hf'], 52, torch.float32), (4, ['patch_ACC_hf'], 7, torch.float32)])
def test_planar_arena_equals_interleaved(nsig, chans, W, out_dtype):
  """ARENA_PLANAR (one plane per signal; RHC plane first, SCG planes of kept windows only; pair ownership) gives exactly
  what the interleaved kernel gives: decisions, reasons, kept list, pairs, windows — for every template shape, odd row
  offsets (leading element), odd window lengths, planted NaN / Inf, dataset-level pairs (two passes) and KEEP_ALL."""
  sig = (synth_ref.SIG_NAMES_5 + ['x5', 'x6'])[:nsig] if nsig != 4 else synth_ref.DEFAULT_SIG_NAMES
  kinds = synth_ref.kinds_for(sig)
  n_rec, T = 6, 30011
  metas = [synth_ref.record_meta(60, events=e) for e in ({'PA_1': 0.002, 'RV_1': 31}, {'PA_1': 0}, {'RA_1': 0}, {'RV_1': 1, 'PA_1': 20.005, 'RA_1': 40, 'PA_2': 44.444},
                                                         {'PA_1': 0.5}, {'PA_1': 3.3})]
  cols, rcol = scgrhc.resolve_columns(sig, chans)
  arena = torch.empty((n_rec * T, nsig), dtype=torch.float64, device=DEV)
  ops.synth_records(arena, 4242, 0, n_rec, T, list(kinds), 16, W)
  arena[1000:1003, cols[0]] = float('nan')                 # NaN in an SCG column of a kept window
  arena[T + 5000, cols[-1]] = float('inf')                 # Inf in an SCG column: min/max stay finite-or-inf, not NaN
  planes = arena.t().contiguous()
  plan = scgrhc.plan_cohort(metas, 'PA', [T] * n_rec, W, stride=0)
  assert plan.n_cand > 20 and (plan.intervals['row0'] % 2 == 1).any() and (plan.intervals['row0'] % 2 == 0).any()
  # the device generator writes the planar layout itself
  gen = torch.empty((nsig, n_rec * T), dtype=torch.float64, device=DEV)
  ops.synth_records(gen, 4242, 0, n_rec, T, list(kinds), 16, W, n_rec * T)
  clean = torch.empty_like(arena)
  ops.synth_records(clean, 4242, 0, n_rec, T, list(kinds), 16, W)
  assert torch.equal(gen, clean.t().contiguous())
  for kw in (dict(), dict(use_global_min_max=True), dict(keep_all=True), dict(predicates_only=True)):
    if kw.get('use_global_min_max'):
      a2, p2 = clean, clean.t().contiguous()                # NaN pairs would poison the dataset-level min/max
    else:
      a2, p2 = arena, planes
    a = scgrhc.prepare_windows(a2, plan, cols, rcol, -50.0, out_dtype=out_dtype, **kw)
    b = scgrhc.prepare_windows(p2, plan, cols, rcol, -50.0, out_dtype=out_dtype, planar=True, **kw)
    _same_store(a, b)
    assert 0 < b.n_kept and (kw.get('keep_all') or b.n_kept < plan.n_cand)
    if kw.get('use_global_min_max'):
      assert torch.equal(a.global_minmax, b.global_minmax)
    elif not kw:
      rej = (~b.keep.bool()).cpu().numpy()
      assert np.isnan(b.minmax.cpu().numpy()[rej][:, :2]).all()      # SCG planes of rejected windows were never read
  # non-finite RHC reaches the regression -> ValueError, as in the interleaved kernel
  bad = planes.clone(); bad[rcol, plan.intervals['row0'][0] + 10] = float('nan')
  with pytest.raises(ValueError):
    scgrhc.prepare_windows(bad, plan, cols, rcol, -50.0, planar=True)
  # capacity edge: the last window of the last plane ends exactly at the end of an odd-sized arena
  tail_rows = (planes.shape[1] // W) * W - (1 - (planes.shape[1] // W * W) % 2)
  if W > 50:
    odd = planes[:, :tail_rows].contiguous()
    last = scgrhc.Plan(np.array([(tail_rows - 2 * W, 0, 2, 0)], dtype=scgrhc.engine.INTERVAL_DTYPE), 2, W)
    a = scgrhc.prepare_windows(odd.t().contiguous(), last, cols, rcol, -50.0, out_dtype=out_dtype, keep_all=True)
    b = scgrhc.prepare_windows(odd, last, cols, rcol, -50.0, out_dtype=out_dtype, keep_all=True, planar=True)
    _same_store(a, b)


def test_planar_digital_ingest_equals_interleaved_ingest():
  """HostIngest decodes format-16 chunks into planar arenas (the default for digital cohorts); the interleaved decode + kernel
  (planar=False) gives the same store."""
  from scgrhc.engine import HostIngest
  sig = synth_ref.DEFAULT_SIG_NAMES
  kinds = synth_ref.kinds_for(sig)
  rows = [30011, 752, 15000, 40001, 1500, 8000, 22222]
  metas = [synth_ref.record_meta(90, events=e) for e in
           ({'PA_1': 0.2, 'RV_1': 50}, {'PA_1': 0}, {'RA_1': 0}, {'RV_1': 1, 'PA_1': 20, 'RA_1': 60, 'PA_2': 70}, {'PA_1': 0.5}, {'PA_1': 10}, {'PA_1': 3.3})]
  recs = [synth_ref.gen_record(H.SEED, 500 + r, T, kinds=kinds) for r, T in enumerate(rows)]
  gains = [[1e5 + 10 * r, 2e5, 1.5e5, 400.0 + r] for r in range(len(rows))]
  bases = [[3.0 * r, -7.0, 0.0, 100.0 - r] for r in range(len(rows))]
  d = [np.clip(np.round(p * np.array(g) + np.array(b)), -32767, 32767).astype(np.int16) for p, g, b in zip(recs, gains, bases)]
  hostd = torch.from_numpy(np.concatenate(d)).pin_memory()
  plan = scgrhc.plan_cohort(metas, 'PA', rows, 750)
  for chunk in (1, 3, 100):
    for kw in (dict(), dict(use_global_min_max=True)):
      a = HostIngest(plan, rows, 4, DEV, chunk_records=chunk, digital_nsig=4, planar=False).run(hostd, [0, 1, 2], 3, -50.0, decode=([0, 1, 2, 3], gains, bases), **kw)
      ing = HostIngest(plan, rows, 4, DEV, chunk_records=chunk, digital_nsig=4, planar=True)
      assert ing.planar and HostIngest(plan, rows, 4, DEV, chunk_records=chunk, digital_nsig=4).planar and \
          not HostIngest(plan, rows, 4, DEV, chunk_records=chunk).planar          # default: on for digital cohorts only
      b = ing.run(hostd, [0, 1, 2], 3, -50.0, decode=([0, 1, 2, 3], gains, bases), **kw)
      _same_store(a, b)
      assert b.n_kept > 0
  # one calibration for the whole cohort (the other decode kernel)
  a = HostIngest(plan, rows, 4, DEV, chunk_records=2, digital_nsig=4, planar=False).run(hostd, [0, 1, 2], 3, -50.0, decode=([0, 1, 2, 3], gains[0], bases[0]))
  b = HostIngest(plan, rows, 4, DEV, chunk_records=2, digital_nsig=4, planar=True).run(hostd, [0, 1, 2], 3, -50.0, decode=([0, 1, 2, 3], gains[0], bases[0]))
  _same_store(a, b)


def test_capacity_edge_uses_fallback_loads():
  """Odd number of arena elements: the last window cannot be bulk-copied with 16-byte granularity."""
  sig = synth_ref.SIG_NAMES_5
  T = 7501
  p = synth_ref.gen_record(H.SEED, 11, T, kinds=synth_ref.kinds_for(sig), defect_scale=0)
  meta = synth_ref.record_meta(20, events={'PA_1': 0.002})
  c = dict(in_channels=['patch_ACC_lat', 'patch_ECG'], chamber='PA', segment_size=1.5, min_RHC=-50)
  plan, st = run_record(p, sig, meta, c)
  assert plan.n_cand == 10 and plan.intervals['row0'][0] == 1
  rw = orc.scan_record(p, sig, meta, c['in_channels'], 'PA', 1.5, -50)
  s_o, r_o, _ = orc.normalise_record(p, sig, c['in_channels'], rw)
  scg, rhc = st.materialise()
  assert scg.cpu().numpy().tobytes() == s_o.tobytes() and rhc.cpu().numpy().tobytes() == r_o.tobytes()


def test_nonfinite_rhc_raises_like_reference_unless_flat():
  sig = synth_ref.DEFAULT_SIG_NAMES
  p = synth_ref.gen_record(H.SEED, 21, 3000, kinds=synth_ref.kinds_for(sig), defect_scale=0)
  meta = synth_ref.record_meta(6, events={'PA_1': 0})
  c = dict(in_channels=sig[:3], chamber='PA', segment_size=1.5, min_RHC=-50)
  q = p.copy(); q[800, 3] = np.nan
  with pytest.raises(ValueError):
    run_record(q, sig, meta, c)
  q[900:1000, 3] = 5.0                                  # flat line short-circuits before the regression
  plan, st = run_record(q, sig, meta, c)
  assert st.keep.cpu().tolist()[1] == 0
  assert st.reason.cpu().tolist()[1] & N.REASON_FLAT
  q = p.copy(); q[100, 0] = np.nan                      # NaN in an SCG channel: np.min/np.max propagate it
  plan, st = run_record(q, sig, meta, c)
  mm = st.minmax.cpu().numpy()
  assert np.isnan(mm[0, 0]) and np.isnan(mm[0, 1]) and st.keep.cpu().tolist()[0] == 1
  assert np.isnan(st.materialise()[0][0].cpu().numpy()).all()


def test_keep_all_normalises_rejected_windows_too():
  sig, p, meta = H.small_record()
  c = H.effective_config('waveform_06')
  plan, st = run_record(p, sig, meta, c, keep_all=True)
  assert st.n_kept == plan.n_cand
  rw = orc.scan_record(p, sig, meta, c['in_channels'], c['chamber'], 1.5, c['min_RHC'])
  want = rw.keep.copy()
  rw.keep[:] = True
  s_o, r_o, _ = orc.normalise_record(p, sig, c['in_channels'], rw)
  scg, rhc = st.materialise()
  assert scg.cpu().numpy().tobytes() == s_o.tobytes()
  reason = st.reason.cpu().numpy()
  assert ((reason & 15) == 0).tolist() == want.tolist()


def test_empty_and_ragged_inputs():
  sig = synth_ref.DEFAULT_SIG_NAMES
  p = synth_ref.gen_record(H.SEED, 2, 2000, kinds=synth_ref.kinds_for(sig))
  c = dict(in_channels=sig[:3], chamber='PCW', segment_size=1.5, min_RHC=-50)
  plan, st = run_record(p, sig, synth_ref.record_meta(4, events={'PA_1': 0}), c)       # chamber never visited
  assert plan.n_cand == 0 and st.n_kept == 0
  c['chamber'] = 'PA'
  plan, st = run_record(p, sig, synth_ref.record_meta(4, events={'PA_1': 3.5}), c)     # interval shorter than a window
  assert plan.n_cand == 0 and st.n_kept == 0
  plan, st = run_record(p, sig, synth_ref.record_meta(60, events={'PA_1': 0.1}), c)    # interval runs past the record: clamped
  assert plan.n_cand == (2000 - 50) // 750
  with pytest.raises(ValueError):
    scgrhc.resolve_columns(sig, ['patch_ECG'])


def test_full_size_properties_1000_records():
  """BASELINE config 2 shape (1,000 x 10-min records, waveform_06): size-independent properties."""
  n_rec, T = 1000, 300000
  kinds = synth_ref.kinds_for(synth_ref.DEFAULT_SIG_NAMES)
  arena = torch.empty((n_rec * T, 4), dtype=torch.float64, device=DEV)
  ops.synth_records(arena, H.SEED, 0, n_rec, T, list(kinds), 16, 750)
  meta = synth_ref.record_meta(600)
  plan = scgrhc.plan_uniform(meta, 'PA', T, 750, n_rec)
  assert plan.n_cand == 200 * n_rec
  st = scgrhc.prepare_windows(arena, plan, [0, 1, 2], 3, -50.0)
  keep = st.keep.bool()
  assert int(keep.sum()) == st.n_kept
  k = st.kept_idx
  assert bool((k[1:] > k[:-1]).all()) and bool(keep[k].all())
  # records 0 and 1 of this cohort are the golden 5-signal records minus the ECG column
  full = H.load_json('records_full.json')['configs']['waveform_06']['records']
  for r in (0, 1):
    m = st.rec_id == r
    assert st.start_idx[m].cpu().tolist() == full['rec%d' % r]['start']
  # every kept window: min sample normalises to exactly 0, max to (mx-mn)/(mx-mn+1e-4) < 1
  pos = torch.arange(0, st.n_kept, 37, device=DEV)
  scg, rhc = st.gather(pos)
  assert float(scg.amin()) == 0.0 and float(rhc.amin()) == 0.0
  assert bool((scg.amin(dim=(1, 2)) == 0).all()) and bool((rhc.amin(dim=(1, 2)) == 0).all())
  mm = st.kept_minmax()[pos]
  top = ((mm[:, 1] - mm[:, 0]) / (mm[:, 1] - mm[:, 0] + 0.0001)).float()
  assert bool((scg.amax(dim=(1, 2)) == top).all())
  # idempotence: a second run is byte-identical
  st2 = scgrhc.prepare_windows(arena, plan, [0, 1, 2], 3, -50.0)
  assert torch.equal(st2.keep, st.keep) and torch.equal(st2.minmax, st.minmax)
  scg2, rhc2 = st2.gather(pos)
  assert torch.equal(scg2, scg) and torch.equal(rhc2, rhc)
  # shard invariance: the second half of the cohort processed alone gives the same windows
  half = scgrhc.plan_uniform(meta, 'PA', T, 750, n_rec // 2, rec0=n_rec // 2)
  sth = scgrhc.prepare_windows(arena[(n_rec // 2) * T:], half, [0, 1, 2], 3, -50.0)
  m = st.rec_id >= n_rec // 2
  assert torch.equal(sth.start_idx, st.start_idx[m]) and torch.equal(sth.rec_id, st.rec_id[m])
  assert torch.equal(sth.kept_minmax(), st.kept_minmax()[m])


@pytest.mark.parametrize('segment_size,nsig', [(1.0, 4), (2.0, 4), (0.128, 5), (0.1, 4), (0.004, 4), (1.7, 12), (2.048, 3)])
def test_other_window_lengths_and_wide_records(segment_size, nsig):
  """The runtime-W kernel (every length 2..1024 the library accepts) and records with many signals."""
  names = (synth_ref.SIG_NAMES_5 + ['s%d' % i for i in range(5, 12)])[:nsig] if nsig != 3 else ['patch_ACC_hf', 'RHC_pressure', 'patch_ACC_dv']
  if nsig == 4:
    names = synth_ref.DEFAULT_SIG_NAMES
  kinds = synth_ref.kinds_for(names)
  W = int(segment_size * 500)
  T = 20011
  chans = [n for n in names if n.startswith('patch_ACC')][::-1]
  meta = synth_ref.record_meta(40, events={'PA_1': 0.302, 'RA_1': 25.4, 'PA_2': 31.001})
  arena = torch.empty((2 * T, nsig), dtype=torch.float64, device=DEV)
  ops.synth_records(arena, 5, 0, 2, T, list(kinds), 16, 750)
  cols, rcol = scgrhc.resolve_columns(names, chans)
  plan = scgrhc.plan_uniform(meta, 'PA', T, W, 2)
  st = scgrhc.prepare_windows(arena, plan, cols, rcol, -50.0, out_dtype=torch.float64)
  scg, rhc = st.materialise()
  scg, rhc, rec_id = scg.cpu().numpy(), rhc.cpu().numpy(), st.rec_id.cpu().numpy()
  for r in range(2):
    p = synth_ref.gen_record(5, r, T, kinds=kinds)
    rw = orc.scan_record(p, names, meta, chans, 'PA', segment_size, -50.0)
    # exactly-constant windows shorter than 51 samples are excluded: the reference's R^2 on them is rounding
    # noise of np.average (1.0 or 0.0); longer ones are flat-line rejects on both sides
    const = (rw.minmax[:, 2] == rw.minmax[:, 3]) & (W < 51)
    m = rec_id == r
    mine = st.start_idx.cpu().numpy()[m]
    cand_of = {int(a): i for i, a in enumerate(rw.abs_start)}
    abs_mine = st.kept_idx.cpu().numpy()[m] - (plan.n_cand // 2) * r
    ok_mine = ~const[abs_mine]
    k = np.nonzero(rw.keep & ~const)[0]
    assert abs_mine[ok_mine].tolist() == k.tolist(), (segment_size, r)
    assert mine[ok_mine].tolist() == rw.rel_start[k].tolist()
    rw.keep = rw.keep & ~const
    s_o, r_o, mm_o = orc.normalise_record(p, names, chans, rw, out_dtype=np.float64)
    assert scg[m][ok_mine].tobytes() == s_o.tobytes() and rhc[m][ok_mine].tobytes() == r_o.tobytes()
  assert plan.n_cand > 0


def test_unsupported_shapes_fail_loudly():
  arena = torch.zeros((5000, 4), dtype=torch.float64, device=DEV)
  meta = synth_ref.record_meta(10, events={'PA_1': 0})
  plan = scgrhc.plan_uniform(meta, 'PA', 5000, 1025, 1)
  with pytest.raises(N.ScgrhcError, match='window of 1025 samples'):
    scgrhc.prepare_windows(arena, plan, [0, 1, 2], 3, -50.0)
  plan = scgrhc.plan_uniform(meta, 'PA', 5000, 750, 1)
  with pytest.raises(N.ScgrhcError):
    scgrhc.prepare_windows(arena, plan, [0, 1, 2, 3, 0], 3, -50.0)
  with pytest.raises(N.ScgrhcError):
    scgrhc.prepare_windows(arena, plan, [0, 1, 7], 3, -50.0)
  with pytest.raises(RuntimeError):
    scgrhc.prepare_windows(arena.cpu(), plan, [0, 1, 2], 3, -50.0)


@pytest.mark.parametrize('stride', [750, 250, 1, 1000])
def test_strided_windows_extension(stride):
  """Overlapping / gapped windows (extension; stride == W is the reference): same per-window arithmetic."""
  sig = synth_ref.DEFAULT_SIG_NAMES
  T = 9000
  p = synth_ref.gen_record(H.SEED, 31, T, kinds=synth_ref.kinds_for(sig))
  meta = synth_ref.record_meta(18, events={'PA_1': 0.5, 'RV_1': 9.3, 'PA_2': 11})
  plan = scgrhc.plan_cohort([meta], 'PA', [T], 750, stride=stride)
  arena = torch.from_numpy(p).to(DEV)
  st = scgrhc.prepare_windows(arena, plan, [0, 1, 2], 3, -50.0)
  rw = orc.scan_record(p, sig, meta, sig[:3], 'PA', 1.5, -50.0, stride=stride)
  assert plan.n_cand == len(rw.abs_start) > 0
  k = np.nonzero(rw.keep)[0]
  assert st.start_idx.cpu().tolist() == rw.rel_start[k].tolist() and st.stop_idx.cpu().tolist() == (rw.rel_start[k] + 750).tolist()
  s_o, r_o, mm_o = orc.normalise_record(p, sig, sig[:3], rw)
  scg, rhc = st.materialise()
  assert scg.cpu().numpy().tobytes() == s_o.tobytes() and rhc.cpu().numpy().tobytes() == r_o.tobytes()


def test_host_ingest_chunked_equals_resident(tmp_path):
  """Ragged record lengths, chunk boundaries that split the cohort unevenly, records without windows: the
  double-buffered host pipeline (fp64 and int16/format-16 inputs) yields exactly the resident result."""
  from scgrhc.engine import HostIngest
  sig = synth_ref.DEFAULT_SIG_NAMES
  kinds = synth_ref.kinds_for(sig)
  rows = [30011, 752, 15000, 40000, 1500, 8000, 22222, 9000]
  metas = [synth_ref.record_meta(90, events=e) for e in
           ({'PA_1': 0.2, 'RV_1': 50}, {'PA_1': 0}, {'RA_1': 0}, {'RV_1': 1, 'PA_1': 20, 'RA_1': 60, 'PA_2': 70}, {'PA_1': 0.5},
            {'PA_1': 10}, {'PA_1': 3.3}, {'RV_1': 0})]       # records 2 and 7 (the last) have no PA interval: empty chunks
  recs = [synth_ref.gen_record(H.SEED, 200 + r, T, kinds=kinds) for r, T in enumerate(rows)]
  host = torch.from_numpy(np.concatenate(recs)).pin_memory()
  plan = scgrhc.plan_cohort(metas, 'PA', rows, 750)
  ref = scgrhc.prepare_windows(host.to(DEV), plan, [0, 1, 2], 3, -50.0)
  for chunk in (1, 2, 3, 100):
    st = HostIngest(plan, rows, 4, DEV, chunk_records=chunk).run(host, [0, 1, 2], 3, -50.0)
    assert st.n_kept == ref.n_kept > 0 and torch.equal(st.kept_idx, ref.kept_idx) and torch.equal(st.rec_id, ref.rec_id)
    assert torch.equal(st.start_idx, ref.start_idx) and torch.equal(st.minmax[st.kept_idx], ref.minmax[ref.kept_idx])
    a, b = st.materialise(), ref.materialise()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
  # dataset-global pairs over a streamed cohort (two passes over the host data) == the resident two-pass result
  refg = scgrhc.prepare_windows(host.to(DEV), plan, [0, 1, 2], 3, -50.0, use_global_min_max=True)
  for chunk in (1, 2, 3):          # chunk_records=1: chunks without a single candidate window, mid-cohort and trailing
    stg = HostIngest(plan, rows, 4, DEV, chunk_records=chunk).run(host, [0, 1, 2], 3, -50.0, use_global_min_max=True,
                                                                  buffers={'scg': torch.full((refg.n_kept, 3, 750), float('nan'), device=DEV),
                                                                           'rhc': torch.full((refg.n_kept, 1, 750), float('nan'), device=DEV)})
    assert stg.dense and stg.n_kept == refg.n_kept and torch.equal(stg.global_minmax, refg.global_minmax)
    assert torch.equal(stg.scg[:stg.n_kept], refg.scg[:refg.n_kept]) and torch.equal(stg.rhc[:stg.n_kept], refg.rhc[:refg.n_kept])
  # format-16 digital frames with per-record calibration
  gains = [[1e5 + 10 * r, 2e5, 1.5e5, 400.0 + r] for r in range(len(rows))]
  bases = [[3.0 * r, -7.0, 0.0, 100.0 - r] for r in range(len(rows))]
  d = [np.clip(np.round(p * np.array(g) + np.array(b)), -32767, 32767).astype(np.int16) for p, g, b in zip(recs, gains, bases)]
  phys = [(x.astype(np.float64) - np.array(b)) / np.array(g) for x, g, b in zip(d, gains, bases)]
  ref2 = scgrhc.prepare_windows(torch.from_numpy(np.concatenate(phys)).to(DEV), plan, [0, 1, 2], 3, -50.0)
  hostd = torch.from_numpy(np.concatenate(d)).pin_memory()
  st = HostIngest(plan, rows, 4, DEV, chunk_records=2, digital_nsig=4).run(hostd, [0, 1, 2], 3, -50.0,
                                                                          decode=([0, 1, 2, 3], gains, bases))
  assert st.n_kept == ref2.n_kept > 0 and torch.equal(st.kept_idx, ref2.kept_idx)
  a, b = st.materialise(), ref2.materialise()
  assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


@pytest.mark.parametrize('nsig,out_dtype,W', [(5, torch.float32, 750), (4, torch.float64, 750), (7, torch.float32, 333), (5, torch.float32, 1024)])
def test_subset_fan_out_equals_one_pass_per_subset(nsig, out_dtype, W):
  """scgrhc_normalize_subsets (the window read once, written for every channel subset) against scgrhc_process_windows per
  subset, bit for bit — NaN and Inf planted in SCG columns (np.min/np.max poisoning per subset), a constant column
  (IEEE-division tier), 9 subsets (two launches)."""
  import scgrhc
  sig = (synth_ref.SIG_NAMES_5 + ['x5', 'x6'])[:nsig] if nsig != 4 else synth_ref.DEFAULT_SIG_NAMES
  kinds = synth_ref.kinds_for(sig)
  T, n_rec = 40000, 3
  meta = synth_ref.record_meta(80, events={'RA_1': 0, 'PA_1': 3.3, 'RV_1': 71})
  recs = [synth_ref.gen_record(H.SEED, 300 + r, T, kinds=kinds) for r in range(n_rec)]
  recs[0][5000, 1] = np.nan
  recs[1][9000:9010, 0] = np.inf
  recs[2][12000:16000, 2] = 0.0
  arena = torch.from_numpy(np.concatenate(recs)).to(DEV)
  plan = scgrhc.plan_cohort([meta] * n_rec, 'PA', [T] * n_rec, W)
  rcol = sig.index('RHC_pressure')
  sup = [c for c in range(nsig) if c != rcol][:4]
  subsets = [sup, sup[:2], sup[1:3], [sup[0], sup[2]], [sup[0]], [sup[1]], [sup[2]], sup[:3], [sup[-1]]]
  stores = scgrhc.prepare_subsets(arena, plan, sup, rcol, -50.0, subsets, out_dtype=out_dtype)
  assert len(stores) == len(subsets)
  for sub, st in zip(subsets, stores):
    ref = scgrhc.prepare_windows(arena, plan, sub, rcol, -50.0, out_dtype=out_dtype)
    assert st.n_kept == ref.n_kept and 0 < st.n_kept < plan.n_cand
    assert torch.equal(st.kept_idx, ref.kept_idx) and torch.equal(st.start_idx, ref.start_idx)
    a, b = st.materialise(), ref.materialise()
    assert a[0].cpu().numpy().tobytes() == b[0].cpu().numpy().tobytes(), sub
    assert a[1].cpu().numpy().tobytes() == b[1].cpu().numpy().tobytes(), sub
    assert st.kept_minmax().cpu().numpy().tobytes() == ref.kept_minmax().cpu().numpy().tobytes(), sub
  assert stores[0].rhc.data_ptr() == stores[-1].rhc.data_ptr()            # one RHC tensor for every subset
  with pytest.raises(ValueError):
    scgrhc.prepare_subsets(arena, plan, sup, rcol, -50.0, [[sup[1], sup[0]]])


def test_nonfinite_scg_does_not_change_the_rhc_predicates():
  """has_noise() looks at the RHC channel only (waveform_noise.py:44-49): NaN / Inf in an SCG column must leave keep and
  reason untouched (a straight-line window stays rejected), and only poison that window's joint min/max as np.min does."""
  import scgrhc
  from scgrhc import _native as N
  sig = synth_ref.DEFAULT_SIG_NAMES
  T = 60000
  p = synth_ref.gen_record(H.SEED, 411, T, kinds=synth_ref.kinds_for(sig))
  meta = synth_ref.record_meta(120, events={'PA_1': 0.0})
  plan = scgrhc.plan_cohort([meta], 'PA', [T], 750)
  base = scgrhc.prepare_windows(torch.from_numpy(p).to(DEV), plan, [0, 1, 2], 3, -50.0)
  reason = base.reason.cpu().numpy()
  straight = np.nonzero(reason == N.REASON_STRAIGHT)[0]
  kept = np.nonzero(reason == 0)[0]
  assert len(straight) >= 2 and len(kept) >= 2
  q = p.copy()
  q[straight[0] * 750 + 10, 1] = np.nan
  q[straight[1] * 750 + 700, 0] = np.inf
  q[kept[0] * 750 + 5, 2] = np.nan
  q[kept[1] * 750 + 5, 2] = -np.inf
  for cols in ([0, 1, 2], [2, 0], [1]):
    st = scgrhc.prepare_windows(torch.from_numpy(q).to(DEV), plan, cols, 3, -50.0)
    assert st.reason.cpu().numpy().tolist() == reason.tolist(), cols
    assert torch.equal(st.keep, base.keep)
  st = scgrhc.prepare_windows(torch.from_numpy(q).to(DEV), plan, [0, 1, 2], 3, -50.0)
  mm = st.minmax.cpu().numpy()
  assert np.isnan(mm[kept[0], :2]).all() and mm[kept[1], 0] == -np.inf and np.isfinite(mm[kept[1], 1])
  assert (mm[:, 2:] == base.minmax.cpu().numpy()[:, 2:]).all()


@pytest.mark.parametrize('seed', list(range(16)))
def test_fuzz_shapes_both_layouts_against_the_oracle(seed):
  """Seeded random jobs — signal count, channel subset and order, window length, ragged records, side-car events, floor,
  output type — through the interleaved kernel, the planar kernel and the numpy oracle (scan_record / normalise_record, the
  restatement pinned to the unmodified reference): decisions, indices, pairs and samples bit for bit."""
  rng = np.random.default_rng(1000 + seed)
  nsig = int(rng.integers(2, 9))
  acc = ['patch_ACC_lat', 'patch_ACC_hf', 'patch_ACC_dv']
  rest = ['patch_ECG', 'x5', 'x6', 'x7']
  if seed % 2 == 0:     # the three axes first (so that two-, three-channel subsets occur), else one axis guaranteed and the rest at random
    others = [str(s) for s in rng.permutation(acc)] + [str(s) for s in rng.permutation(rest)]
  else:
    first = str(rng.choice(acc))
    others = [first] + [str(s) for s in rng.permutation([a for a in acc if a != first] + rest)]
  sig = [str(s) for s in rng.permutation(others[:nsig - 1] + ['RHC_pressure'])]        # columns in any order, at least one SCG axis
  if nsig == 4 and seed % 2:
    sig = list(synth_ref.DEFAULT_SIG_NAMES)                                             # wfdb's usual layout: the IDENT instantiations
  scg_pool = [s for s in sig if s.startswith('patch_ACC')]
  C = int(rng.integers(1, min(4, len(scg_pool)) + 1))
  chans = [str(s) for s in rng.permutation(scg_pool)[:C]]
  seg = float(rng.choice([1.5, 1.5, 1.0, 0.75, 0.25, 2.0, 0.102, 1.234]))
  W = int(seg * 500)
  out_dtype = torch.float64 if seed % 5 == 0 else torch.float32
  floor = float(rng.choice([-50.0, -50.0, 0.0, 12.5]))
  n_rec = int(rng.integers(2, 6))
  rows = [int(rng.integers(3 * W + 1, 30 * W + 7)) for _ in range(n_rec)]
  chamber = str(rng.choice(['PA', 'RV']))
  metas = []
  for T in rows:
    dur = T / 500.0
    ts = np.sort(rng.uniform(0, dur, size=int(rng.integers(1, 5))))
    ev = {'%s_%d' % (rng.choice(['PA', 'RV', 'RA']), k + 1): float(np.round(t, int(rng.integers(0, 4)))) for k, t in enumerate(ts)}
    ev[chamber + '_9'] = 0.002 * float(rng.integers(0, 3))
    metas.append(synth_ref.record_meta(int(dur) + 1, events=ev))
  kinds = synth_ref.kinds_for(sig)
  recs = [synth_ref.gen_record(7000 + seed, r, T, kinds=kinds) for r, T in enumerate(rows)]
  cols, rcol = scgrhc.resolve_columns(sig, chans)
  host = np.concatenate(recs)
  arena = torch.from_numpy(host).to(DEV)
  planes = arena.t().contiguous()
  plan = scgrhc.plan_cohort(metas, chamber, rows, W)
  a = scgrhc.prepare_windows(arena, plan, cols, rcol, floor, out_dtype=out_dtype)
  b = scgrhc.prepare_windows(planes, plan, cols, rcol, floor, out_dtype=out_dtype, planar=True)
  _same_store(a, b)
  scg, rhc = b.materialise()
  scg, rhc = scg.cpu().numpy(), rhc.cpu().numpy()
  rec_id, start, mm = b.rec_id.cpu().numpy(), b.start_idx.cpu().numpy(), b.kept_minmax().cpu().numpy()
  npdt = np.float64 if out_dtype == torch.float64 else np.float32
  total = 0
  for r, p in enumerate(recs):
    rw = orc.scan_record(p, sig, metas[r], chans, chamber, seg, floor)
    k = np.nonzero(rw.keep)[0]
    m = rec_id == r
    assert start[m].tolist() == rw.rel_start[k].tolist(), (seed, r)
    s_o, r_o, mm_o = orc.normalise_record(p, sig, chans, rw, out_dtype=npdt)
    assert (mm[m] == mm_o).all(), (seed, r)
    assert scg[m].tobytes() == s_o.tobytes() and rhc[m].tobytes() == r_o.tobytes(), (seed, r)
    total += len(k)
  assert total == b.n_kept
