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


@pytest.mark.parametrize('order,band,btype,chunk,nbuf,tol', [(4, (1.0, 40.0), 'bandpass', 0, 0, 1e-10), (2, (0.5, 20.0), 'bandpass', 24, 1, 1e-10),
                                                             (4, 30.0, 'low', 7, 2, 1e-10), (3, 2.0, 'high', 32, 1, 1e-10),
                                                             (1, 5.0, 'low', 3, 2, 1e-10), (4, (1.0, 40.0), 'bandpass', 12, 2, 1e-10),
                                                             (3, (2.0, 35.0), 'bandpass', 0, 0, 1e-10), (6, 45.0, 'low', 24, 1, 1e-10)])
def test_sosfiltfilt_time_parallel_scan_within_1e_10(order, band, btype, chunk, nbuf, tol):
  """The time-parallel variant (BASELINE north_star: fp64 device mode within 1e-10) against scipy and the exact kernel:
  5 signals per row (40-byte rows: plain warp copies, two column groups, the second in place) and 4 signals per row
  (32-byte rows: bulk async copies); records shorter than one span, of exactly one span, and of many spans + a ragged
  last chunk."""
  sos = signal.butter(order, band, btype=btype, fs=500, output='sos')
  L = chunk or 24                                                          # the library default for 32-byte rows
  for sig, cols in ((synth_ref.SIG_NAMES_5, [0, 1, 2, 3, 4]), (synth_ref.DEFAULT_SIG_NAMES, [0, 1, 2]), (synth_ref.DEFAULT_SIG_NAMES, [3, 1])):
    rows = [30000, 311, 20001, 32 * L, 64 * L + 1, 4096]
    recs = [synth_ref.gen_record(H.SEED, 50 + r, T, kinds=synth_ref.kinds_for(sig)) for r, T in enumerate(rows)]
    arena = torch.from_numpy(np.concatenate(recs)).to(DEV)
    fast = filters.sosfiltfilt(arena, rows, sos, cols, exact=False, chunk=chunk, nbuf=nbuf).cpu().numpy()
    exact = filters.sosfiltfilt(arena, rows, sos, cols, exact=True).cpu().numpy()
    inpl = filters.sosfiltfilt(arena.clone(), rows, sos, cols, exact=False, chunk=chunk, nbuf=nbuf, inplace=True).cpu().numpy()
    assert inpl.tobytes() == fast.tobytes()                                 # in place == out of place
    at, worst = 0, 0.0
    for p in recs:
      want = p.copy()
      want[:, cols] = signal.sosfiltfilt(sos, p[:, cols], axis=0)
      assert exact[at:at + len(p)].tobytes() == want.tobytes()
      other = [c for c in range(p.shape[1]) if c not in cols]
      assert (fast[at:at + len(p)][:, other] == p[:, other]).all()          # untouched columns copied through
      scale = np.abs(want[:, cols]).max(axis=0)
      worst = max(worst, float((np.abs(fast[at:at + len(p)][:, cols] - want[:, cols]) / scale).max()))
      at += len(p)
    # 1e-10 of full scale (BASELINE north_star); measured <= 2e-12 for these designs, the 0.5 Hz corner included
    assert worst <= tol, worst
    print('scan max scaled error', len(sig), cols, worst)


def test_sosfiltfilt_rejects_short_records_like_scipy():
  sos = signal.butter(4, (1.0, 40.0), btype='bandpass', fs=500, output='sos')
  arena = torch.zeros((20, 2), dtype=torch.float64, device=DEV)
  from scgrhc._native import ScgrhcError
  with pytest.raises(ScgrhcError, match='greater than padlen'):
    filters.sosfiltfilt(arena, [20], sos, [0])


@pytest.mark.parametrize('ncols', [4, 5, 3, 1, 2, 7])
@pytest.mark.parametrize('up,down', [(1, 2), (250, 500), (2, 5), (3, 2), (125, 500), (500, 500), (100, 500)])
def test_resample_poly_matches_scipy(up, down, ncols):
  """Bit-identical to scipy for integer decimation (register-blocked kernel: every column count it is instantiated for,
  even and odd, and > 5 columns -> the general kernel) and for other ratios (general polyphase kernel)."""
  if ncols not in (4, 5) and (up, down) not in ((1, 2), (125, 500), (2, 5)):
    pytest.skip('column-count sweep only on a few ratios')
  sig = (synth_ref.SIG_NAMES_5 + ['x5', 'x6'])[:ncols] if ncols != 4 else synth_ref.DEFAULT_SIG_NAMES
  if ncols < 4:
    sig = synth_ref.DEFAULT_SIG_NAMES
  rows = [5000, 1234, 777, 12345]
  recs = [synth_ref.gen_record(H.SEED, 70 + r, T, kinds=synth_ref.kinds_for(sig))[:, :ncols].copy() for r, T in enumerate(rows)]
  arena = torch.from_numpy(np.concatenate(recs)).to(DEV)
  out, out_rows = filters.resample_poly(arena, rows, up, down)
  out = out.cpu().numpy()
  at = 0
  for p, n in zip(recs, out_rows):
    want = signal.resample_poly(p, up, down, axis=0)
    assert want.shape == (n, ncols)
    got = out[at:at + n]
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()
    assert got.tobytes() == want.tobytes()                # in fact bit-identical: same taps, same summation order
    at += n
  assert at == out.shape[0]
  fast, _ = filters.resample_poly(arena, rows, up, down, exact=False)      # one FMA per tap (integer decimation only)
  scale = np.abs(out).max(axis=0)
  assert (np.abs(fast.cpu().numpy() - out) / scale).max() <= 1e-14


@pytest.mark.parametrize('down,chans,norm', [(2, ['patch_ACC_lat', 'patch_ACC_hf', 'patch_ACC_dv'], 'minmax'), (2, ['patch_ACC_lat', 'patch_ACC_hf', 'patch_ACC_dv'], 'zscore'),
                                             (2, ['patch_ACC_dv'], 'minmax'), (4, ['patch_ACC_hf', 'patch_ACC_lat'], 'minmax'), (5, ['patch_ACC_lat', 'patch_ACC_hf', 'patch_ACC_dv'], 'minmax')])
def test_decimating_window_kernel_equals_resample_then_window(down, chans, norm):
  """scgrhc_process_windows_decim: native-rate arena -> model-rate windows INSIDE the window kernel (the polyphase FIR of
  scipy.signal.resample_poly, separately rounded taps) == resample_poly (bit-identical to scipy) followed by the window
  kernel, bit for bit — decisions, pairs, windows; ragged records, intervals that start at the very first / end at the
  very last row of a record (the FIR's zero padding), several decimation factors, sub-sets of columns, z-score."""
  import scgrhc
  sig = synth_ref.DEFAULT_SIG_NAMES
  kinds = synth_ref.kinds_for(sig)
  rows = [60000, 20011, 30000, 45007]
  fs = 500 // down
  W = int(1.5 * fs)
  metas = [synth_ref.record_meta(120, events=e) for e in ({'PA_1': 0, 'RV_1': 95}, {'RV_1': 0, 'PA_1': 1.234}, {'RA_1': 0}, {'PA_1': 0.002, 'RV_1': 30, 'PA_2': 50.5})]
  recs = [synth_ref.gen_record(H.SEED, 610 + r, T, kinds=kinds) for r, T in enumerate(rows)]
  arena = torch.from_numpy(np.concatenate(recs)).to(DEV)
  cols, rcol = scgrhc.resolve_columns(sig, chans)
  spec = filters.DecimSpec.design(rows, fs, 500)
  r, out_rows = filters.resample_poly(arena, rows, fs, 500)
  assert out_rows == spec.out_rows
  plan = scgrhc.plan_cohort(metas, 'PA', out_rows, W, fs=float(fs))
  assert plan.n_cand > 50 and (plan.intervals['row0'] == np.concatenate([[0], np.cumsum(out_rows)])[[0]]).any()
  for kw in (dict(), dict(keep_all=True), dict(predicates_only=True)):
    a = scgrhc.prepare_windows(r, plan, cols, rcol, -50.0, normalisation=norm, **kw)
    b = scgrhc.prepare_windows(arena, plan, cols, rcol, -50.0, normalisation=norm, decim=spec, **kw)
    assert a.n_kept == b.n_kept > 0 and torch.equal(a.keep, b.keep) and torch.equal(a.reason, b.reason)
    assert torch.equal(a.kept_idx, b.kept_idx) and torch.equal(a.start_idx, b.start_idx) and torch.equal(a.rec_id, b.rec_id)
    assert a.minmax.cpu().numpy().tobytes() == b.minmax.cpu().numpy().tobytes()
    if not kw.get('predicates_only'):
      x, y = a.materialise(), b.materialise()
      assert x[0].cpu().numpy().tobytes() == y[0].cpu().numpy().tobytes() and x[1].cpu().numpy().tobytes() == y[1].cpu().numpy().tobytes()
  with pytest.raises(ValueError):
    scgrhc.prepare_windows(arena, plan, cols, rcol, -50.0, decim=spec, use_global_min_max=True)
  # one FMA per tap instead of multiply + add: not bit-identical, within 1e-5 of the exact windows after normalisation
  fast = scgrhc.prepare_windows(arena, plan, cols, rcol, -50.0, normalisation=norm, decim=filters.DecimSpec.design(rows, fs, 500, fused=True))
  exact = scgrhc.prepare_windows(arena, plan, cols, rcol, -50.0, normalisation=norm, decim=spec)
  assert torch.equal(fast.keep, exact.keep)
  assert (fast.materialise()[0] - exact.materialise()[0]).abs().max().item() <= 1e-5


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
  want64 = ((scg - mm[0][:, None, None]) / (mm[1] - mm[0] + 0.0001)[:, None, None]).transpose(0, 2, 1)
  want = want64.astype(np.float32)
  assert st.materialise()[0].cpu().numpy().tobytes() == np.ascontiguousarray(want).tobytes()
  # BASELINE configs[1]: fp32 vs fp64 tolerance check of the whole chain — the fp64 device mode against the fp64 host chain
  # within 1e-10 (here: equal), the fp32 mode against it within rel 1e-5; and the fast modes of every stage (time-parallel
  # band-pass, fused-multiply-add decimator) within 1e-10 of full scale before the cast
  st64 = scgrhc.prepare_windows(r, plan, [0, 1, 2], 3, -50.0, out_dtype=torch.float64)
  got64 = st64.materialise()[0].cpu().numpy()
  assert np.abs(got64 - want64).max() <= 1e-10 and got64.tobytes() == np.ascontiguousarray(want64).tobytes()
  assert np.abs(st.materialise()[0].cpu().numpy().astype(np.float64) - want64).max() <= 1e-5
  ff = filters.sosfiltfilt(arena, [T], sos, [0, 1, 2], exact=False)
  rf, _ = filters.resample_poly(ff, [T], 250, 500, exact=False)
  assert (np.abs(rf.cpu().numpy() - q) / np.abs(q).max(axis=0)).max() <= 1e-10
  stf = scgrhc.prepare_windows(rf, plan, [0, 1, 2], 3, -50.0, out_dtype=torch.float64)
  assert torch.equal(stf.keep, st64.keep)
  assert np.abs(stf.materialise()[0].cpu().numpy() - want64).max() <= 1e-9     # normalised samples in [0, 1]; 1e-10 x window scale


def test_recordutil_optional_keys_drive_the_extension_stages(tmp_path, monkeypatch):
  """params.json with bandpass / resample_rate / segment_stride keys goes through filter -> resample -> windows."""
  import json
  import types
  import recordutil
  import scgrhc
  from scgrhc import wfdbio
  root = tmp_path / 'data'; root.mkdir()
  monkeypatch.setattr(recordutil, 'PROCESSED_DATA_PATH', str(root))
  monkeypatch.setattr(recordutil, 'wfdb', wfdbio)
  sig = synth_ref.DEFAULT_SIG_NAMES
  meta = synth_ref.record_meta(80, events={'PA_1': 2.0, 'RV_1': 70})
  p = synth_ref.gen_record(H.SEED, 95, 40000, kinds=synth_ref.kinds_for(sig))
  wfdbio.wrsamp('rec0', 500, ['g', 'g', 'g', 'mmHg'], sig, p, write_dir=str(root))
  (root / 'rec0.json').write_text(json.dumps(meta))
  base = dict(in_channels=sig[:3], chamber='PA', segment_size=1.5, min_RHC=-50, use_global_min_max=False)
  plain, _ = recordutil.prepare_cohort(types.SimpleNamespace(**base))
  ext, _ = recordutil.prepare_cohort(types.SimpleNamespace(bandpass=[1.0, 40.0], resample_rate=250, segment_stride=0.75, **base))
  q = wfdbio.rdrecord(str(root / 'rec0')).p_signal
  sos = signal.butter(4, (1.0, 40.0), btype='bandpass', fs=500, output='sos')
  q[:, :3] = signal.sosfiltfilt(sos, q[:, :3], axis=0)
  q = signal.resample_poly(q, 250, 500, axis=0)
  a, b, W, S = int(2.0 * 250), int(70 * 250), 375, 187
  n = (b - a - W) // S + 1
  assert ext.n_cand == n and plain.n_cand == (35000 - 1000) // 750
  assert ext.start_idx.cpu().tolist() == (np.nonzero(ext.keep.cpu().numpy())[0] * S).tolist()
  k0 = int(ext.kept_idx[0])
  win = q[a + k0 * S: a + k0 * S + W, :3]
  want = ((win - win.min()) / (win.max() - win.min() + 0.0001)).T.astype(np.float32)
  assert ext.materialise()[0][0].cpu().numpy().tobytes() == np.ascontiguousarray(want).tobytes()
  # bandpass_mode 'scan' = the time-parallel filter (in place on the staged cohort): same windows, samples within fp32 noise
  fast, _ = recordutil.prepare_cohort(types.SimpleNamespace(bandpass=[1.0, 40.0], bandpass_mode='scan', resample_rate=250,
                                                            segment_stride=0.75, **base))
  assert fast.kept_idx.cpu().tolist() == ext.kept_idx.cpu().tolist()
  a, b = fast.materialise()[0].cpu().numpy(), ext.materialise()[0].cpu().numpy()
  assert np.abs(a - b).max() <= 1e-5                                       # normalised samples lie in [0, 1]


@pytest.mark.parametrize('sig,chans,out_dtype,W', [(synth_ref.DEFAULT_SIG_NAMES, [0, 1, 2], torch.float32, 750),
                                                   (synth_ref.DEFAULT_SIG_NAMES, [2, 0], torch.float64, 750),
                                                   (synth_ref.SIG_NAMES_5, [0, 1, 2, 4], torch.float32, 300)])
def test_zscore_normalisation_matches_numpy(sig, chans, out_dtype, W):
  """normalisation='zscore' (the brief's mean/std normalisation; the reference only has min-max): the same keep list as
  the default mode, samples against plain numpy — fp64 outputs within 1e-10, fp32 within rel 1e-5 (BASELINE north_star)."""
  import scgrhc
  from oracle import scgrhc_oracle as orc
  T = 60000
  p = synth_ref.gen_record(H.SEED, 97, T, kinds=synth_ref.kinds_for(sig))
  meta = synth_ref.record_meta(120, events={'RA_1': 0, 'PA_1': 4.1, 'RV_1': 110})
  names = [sig[c] for c in chans]
  arena = torch.from_numpy(p).to(DEV)
  plan = scgrhc.plan_cohort([meta], 'PA', [T], W)
  rcol = sig.index('RHC_pressure')
  z = scgrhc.prepare_windows(arena, plan, chans, rcol, -50.0, out_dtype=out_dtype, normalisation='zscore')
  d = scgrhc.prepare_windows(arena, plan, chans, rcol, -50.0, out_dtype=out_dtype)
  assert z.kept_idx.cpu().tolist() == d.kept_idx.cpu().tolist() and 0 < z.n_kept < plan.n_cand
  rw = orc.scan_record(p, sig, meta, names, 'PA', W / 500.0, -50.0)
  s_o, r_o, ms = orc.zscore_record(p, sig, names, rw, out_dtype=np.float64)
  scg, rhc = z.materialise()
  scg, rhc = scg.cpu().numpy().astype(np.float64), rhc.cpu().numpy().astype(np.float64)
  got_ms = z.kept_minmax().cpu().numpy()
  assert np.abs(got_ms - ms).max() <= 1e-12 * np.abs(ms).max()
  tol = 1e-10 if out_dtype == torch.float64 else 1e-5
  scale = max(np.abs(s_o).max(), np.abs(r_o).max())                        # z-scores: a few units
  assert np.abs(scg - s_o).max() <= tol * scale and np.abs(rhc - r_o).max() <= tol * scale
  assert abs(float(scg.mean())) < 1e-2 and abs(float(scg.reshape(len(scg), -1).std(axis=1).mean()) - 1.0) < 5e-2
  with pytest.raises(ValueError):
    scgrhc.prepare_windows(arena, plan, chans, rcol, -50.0, use_global_min_max=True, normalisation='zscore')


def test_extension_stages_stream_per_chunk(tmp_path, monkeypatch):
  """The optional stages run per chunk inside the streamed host ingest: a cohort of ragged records gives the same windows
  whatever the chunking (one record per chunk vs the whole cohort in one chunk), for physical and digital host records,
  local and dataset-global min-max — and equals scipy + the oracle arithmetic on one record."""
  import json
  import types
  import recordutil
  from scgrhc import wfdbio
  root = tmp_path / 'data'; root.mkdir()
  monkeypatch.setattr(recordutil, 'PROCESSED_DATA_PATH', str(root))
  monkeypatch.setattr(recordutil, 'wfdb', wfdbio)
  sig = synth_ref.DEFAULT_SIG_NAMES
  names = []
  for r, T in enumerate((40000, 23456, 31001, 9000)):
    meta = synth_ref.record_meta(T // 500, events={'PA_1': 1.0 + r, 'RV_1': T // 500 - 3})
    p = synth_ref.gen_record(H.SEED, 120 + r, T, kinds=synth_ref.kinds_for(sig))
    wfdbio.wrsamp('rec%d' % r, 500, ['g', 'g', 'g', 'mmHg'], sig, p, write_dir=str(root))
    (root / ('rec%d.json' % r)).write_text(json.dumps(meta))
    names.append('rec%d' % r)
  base = dict(in_channels=sig[:3], chamber='PA', segment_size=1.5, min_RHC=-50)
  for extra in (dict(bandpass=[1.0, 40.0], bandpass_mode='scan', resample_rate=250, use_global_min_max=False),
                dict(bandpass=[0.8, 30.0], resample_rate=125, segment_stride=0.5, use_global_min_max=True),
                dict(resample_rate=250, resample_mode='fused', normalisation='zscore', use_global_min_max=False)):
    params = types.SimpleNamespace(**base, **extra)
    one, _ = recordutil.prepare_cohort(params, names, chunk_records=1)
    whole, _ = recordutil.prepare_cohort(params, names, chunk_records=32)
    assert one.n_kept == whole.n_kept > 0 and one.n_cand == whole.n_cand
    assert torch.equal(one.kept_idx, whole.kept_idx) and torch.equal(one.rec_id, whole.rec_id)
    a, b = one.materialise(), whole.materialise()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert torch.equal(one.kept_minmax(), whole.kept_minmax())
    # integer decimation runs inside the window kernel when it applies (bit-identical taps, fp32, per-window pairs); the
    # separate resample stage gives the same windows
    monkeypatch.setenv('SCGRHC_FUSE_DECIM', '0')
    sep, _ = recordutil.prepare_cohort(params, names, chunk_records=2)
    monkeypatch.delenv('SCGRHC_FUSE_DECIM')
    assert sep.n_kept == whole.n_kept and torch.equal(sep.kept_idx, whole.kept_idx) and torch.equal(sep.kept_minmax(), whole.kept_minmax())
    c = sep.materialise()
    assert torch.equal(c[0], b[0]) and torch.equal(c[1], b[1])


def _guarded(shape, dtype, pad=4096):
  """A tensor carved out of the middle of a larger buffer whose margins hold a sentinel: (view, check) — check() asserts
  that no kernel wrote outside the view (compute-sanitizer is closed on this GPU pool)."""
  n = int(np.prod(shape))
  sentinel = {torch.float64: 1.2345e300, torch.float32: 1.2345e30, torch.int16: 12345, torch.uint8: 123,
              torch.int32: 123456789, torch.int64: 1234567890123}[dtype]
  buf = torch.full((n + 2 * pad,), sentinel, dtype=dtype, device=DEV)
  view = buf[pad:pad + n].view(*shape)
  def check():
    assert bool((buf[:pad] == sentinel).all()) and bool((buf[pad + n:] == sentinel).all()), 'write outside the output buffer'
  return view, check


def test_no_kernel_writes_outside_its_outputs():
  """Every kernel added for the optional stages / the sweep writes only inside its output buffers: outputs are carved
  out of sentinel-filled buffers, ragged shapes chosen to end mid-tile / mid-span / mid-chunk."""
  import scgrhc
  from scgrhc import ops
  sig = synth_ref.SIG_NAMES_5
  rows = [777 + 32 * 24, 5001, 311, 32 * 24 * 3]
  recs = [synth_ref.gen_record(H.SEED, 500 + r, T, kinds=synth_ref.kinds_for(sig)) for r, T in enumerate(rows)]
  for ncols in (4, 5):
    src = torch.from_numpy(np.concatenate([p[:, :ncols] for p in recs])).to(DEV)
    sos = signal.butter(4, (1.0, 40.0), btype='bandpass', fs=500, output='sos')
    # band-pass, time-parallel: in place on a guarded arena
    arena, chk = _guarded(tuple(src.shape), torch.float64)
    arena.copy_(src)
    got = filters.sosfiltfilt(arena, rows, sos, [0, 1, 2], exact=False, inplace=True)
    chk()
    want = filters.sosfiltfilt(src, rows, sos, [0, 1, 2], exact=False)
    assert torch.equal(got, want)
    # decimation by 2 and by 5 and a general ratio into guarded outputs (through the C ABI: the wrapper allocates its own)
    for up, down in ((1, 2), (1, 5), (2, 3)):
      ref, out_rows = filters.resample_poly(src, rows, up, down)
      import ctypes as C
      from scgrhc import _native as N
      d = filters.resample_design(rows[0], up, down)
      designs = [filters.resample_design(n, up, down) for n in rows]
      out, chk = _guarded((sum(out_rows), ncols), torch.float64)
      in0 = torch.from_numpy(np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)).to(DEV)
      out0 = torch.from_numpy(np.concatenate([[0], np.cumsum(out_rows)]).astype(np.int64)).to(DEV)
      taps = torch.from_numpy(d[3]).to(DEV)
      c = ops.ctx(0)
      N.check(c, N.lib().scgrhc_resample_poly(c, ops._ptr(src), ops._ptr(out), ops._ptr(taps), ops._ptr(in0), ops._ptr(out0), len(rows),
                                              max(out_rows), ncols, d[0], d[1], d[4], d[5], 0, ops._stream(0)))
      chk()
      assert torch.equal(out, ref)
  # sweep fan-out: every output of every subset guarded
  src = torch.from_numpy(np.concatenate(recs)).to(DEV)
  metas = [synth_ref.record_meta(max(2, T // 500), events={'PA_1': 0.2}) for T in rows]
  plan = scgrhc.plan_cohort(metas, 'PA', rows, 333)
  pred = scgrhc.prepare_windows(src, plan, [0], 3, -50.0, predicates_only=True)
  n = pred.n_kept
  assert n > 3
  subsets, members, counts = [[0, 1, 2, 4], [1], [0, 4]], [0, 1, 2, 3, 1, 0, 3], [4, 1, 2]
  scgs, mms, checks = [], [], []
  for cnt in counts:
    t, chk = _guarded((n, cnt, 333), torch.float32); scgs.append(t); checks.append(chk)
    t, chk = _guarded((n, 4), torch.float64); mms.append(t); checks.append(chk)
  rhc, chk = _guarded((n, 1, 333), torch.float32); checks.append(chk)
  ops.normalize_subsets(src, plan.device_intervals(src.device), 333, 0, [0, 1, 2, 4], 3, pred.kept_idx, n, members, counts, False, scgs, mms, rhc)
  for chk in checks:
    chk()
  ref = scgrhc.prepare_windows(src, plan, [0, 4], 3, -50.0)
  assert torch.equal(scgs[2], ref.materialise()[0]) and torch.equal(rhc, ref.materialise()[1])
  # format-16 decode, 3 of 5 columns
  d16 = torch.randint(-32768, 32767, (4099, 5), dtype=torch.int16, device=DEV)
  out, chk = _guarded((4099, 3), torch.float64)
  ops.decode_fmt16(d16, [4, 0, 2], [200.0, 3.0, 0.5], [0.0, 1.0, -2.0], out)
  chk()
