"""GPU: the drop-in modules (waveform_noise, recordutil) behind the reference's own signatures, checked
against fixtures produced by the unmodified reference and against the oracle."""
import json
import os
import pickle
import types

import numpy as np
import pytest
import torch

from oracle import scgrhc_oracle as orc
from oracle import synth_ref
from oracle.ref_harness import FakeRecord
from tests import helpers as H

pytestmark = pytest.mark.gpu

import recordutil  # noqa: E402
import waveform_noise  # noqa: E402
from scgrhc import wfdbio  # noqa: E402

P50 = types.SimpleNamespace(min_RHC=-50)


def test_waveform_noise_api_matches_reference():
  g = np.load(os.path.join(H.GOLDEN, 'predicates.npz'))
  names, ys = H.predicate_inputs()
  n_adv = int(g['n_adv'])
  segs = json.loads(str(g['adv_segments']))
  for i in range(n_adv):
    y = ys[i]
    assert waveform_noise.get_flat_lines(y) == [tuple(s) for s in segs[i]], names[i]
    assert waveform_noise.is_straight_line(y) == bool(g['straight'][i]), names[i]
    assert waveform_noise.in_rhc_range(P50, y) == bool(g['in_range'][i]), names[i]
    assert waveform_noise.has_noise(P50, y) == bool(g['has_noise'][i]), names[i]
  assert (waveform_noise.has_noise_batch(P50, ys) == g['has_noise']).all()
  # non-default arguments and other lengths go through the standalone kernels
  y = ys[0][:300]
  assert waveform_noise.get_flat_lines(y, threshold=0.5, min_duration=0.02, sampling_rate=250) == \
      orc.flat_segments(y, threshold=0.5, min_duration=0.02, sampling_rate=250)
  long = np.concatenate([ys[0], ys[1], ys[2]])
  assert waveform_noise.has_noise(P50, long) == orc.has_noise(long, -50)
  assert waveform_noise.is_straight_line(0.5 * np.arange(2000) + 3) is True
  with pytest.raises(ValueError):
    bad = ys[0].copy(); bad[5] = np.nan
    waveform_noise.has_noise(P50, bad)


@pytest.fixture
def data_root(tmp_path, monkeypatch):
  records = {}
  monkeypatch.setattr(recordutil, 'PROCESSED_DATA_PATH', str(tmp_path))
  monkeypatch.setattr(recordutil.wfdb, 'rdrecord', lambda path: records[os.path.basename(path)])

  def add(name, sig, p, meta):
    records[name] = FakeRecord(sig, p)
    (tmp_path / (name + '.json')).write_text(json.dumps(meta))
    (tmp_path / (name + '.hea')).write_text('')
  return add


def cfg_params(cfg, **over):
  c = H.effective_config(cfg)
  c.update(over)
  return types.SimpleNamespace(**c)


def test_recordutil_api_matches_reference(data_root):
  g = np.load(os.path.join(H.GOLDEN, 'record_small.npz'))
  sig, p, meta = H.small_record()
  data_root('small', sig, p, meta)
  assert recordutil.get_record_names() == ['small']
  for name, case in H.load_json('intervals.json').items():
    data_root(name, ['RHC_pressure'], np.zeros((1, 1)), case['meta'])
    for chamber, want in case['intervals'].items():
      assert [list(x) for x in recordutil.get_chamber_intervals(name, chamber)] == want
  for cfg in ('waveform_06', 'waveform_10', 'waveform_19'):
    params = cfg_params(cfg)
    segs = recordutil.get_segments(params, record_name='small')
    assert [s[3] for s in segs] == g[cfg + '.start'].tolist() and [s[4] for s in segs] == g[cfg + '.stop'].tolist()
    assert all(s[0].shape == (750, len(params.in_channels)) and s[1].shape == (750, 1) and s[2] == 'small' for s in segs)
    gm = recordutil.get_global_minmax_vals(segs)
    want = orc.global_minmax(g[cfg + '.minmax'])
    assert [gm[0][0], gm[0][1], gm[1][0], gm[1][1]] == want.tolist()
    raw = list(segs)
    ds = recordutil.SCGDataset(segs, params.segment_size, None, None)
    assert ds.segments is segs and len(ds) == len(raw) and ds.segment_size == 750
    for i in (0, len(ds) - 1):
      item = ds[i]
      assert item is segs[i] and len(item) == 7
      assert item[0].dtype == torch.float32 and not item[0].is_cuda and tuple(item[0].shape) == (len(params.in_channels), 750)
      assert item[0].numpy().tobytes() == g[cfg + '.scg'][i].tobytes()
      assert item[1].numpy().tobytes() == g[cfg + '.rhc'][i].tobytes()
      assert (item[2], item[3], item[4]) == ('small', raw[i][3], raw[i][4])
      assert [item[5][0], item[5][1], item[6][0], item[6][1]] == g[cfg + '.minmax'][i].tolist()
      assert isinstance(item[5][0], np.float64)
    # dataset-global pairs (use_global_min_max): same arithmetic with the given pairs
    ds2 = recordutil.SCGDataset(list(raw), params.segment_size, gm[0], gm[1])
    s0 = ((raw[0][0] - gm[0][0]) / (gm[0][1] - gm[0][0] + 0.0001)).T.astype(np.float32)
    assert ds2[0][0].numpy().tobytes() == s0.tobytes() and ds2[0][5] == gm[0]
    # helper methods kept from the reference
    mn = ds.minmax_norm(raw[0][0], (raw[0][0].min(), raw[0][0].max()))
    assert mn.tobytes() == orc.minmax_norm(raw[0][0], raw[0][0].min(), raw[0][0].max()).tobytes()
    assert ds.pad(torch.zeros(2, 700)).shape[-1] == 750
    with pytest.raises(IndexError):
      ds.pad(torch.zeros(2, 800))
  with pytest.raises(ValueError):
    recordutil.get_segments(cfg_params('waveform_06', in_channels=['nope']), record_name='small')


def test_save_dataloaders_end_to_end(tmp_path, monkeypatch):
  """Records on disk (WFDB format 16) -> save_dataloaders -> pickles -> what the consumers read."""
  root = tmp_path / 'data'; root.mkdir()
  exp = tmp_path / 'waveform_06'; exp.mkdir()
  monkeypatch.setattr(recordutil, 'PROCESSED_DATA_PATH', str(root))
  monkeypatch.setattr(recordutil, 'wfdb', wfdbio)
  sig = synth_ref.SIG_NAMES_5
  meta = synth_ref.record_meta(120, events={'RA_1': 0, 'PA_1': 20, 'RV_1': 100})
  want = {}
  for r in range(3):
    p = synth_ref.gen_record(H.SEED, 40 + r, 60000, kinds=synth_ref.kinds_for(sig))
    wfdbio.wrsamp('rec%d' % r, 500, ['g', 'g', 'g', 'mmHg', 'mV'], sig, p, write_dir=str(root))
    (root / ('rec%d.json' % r)).write_text(json.dumps(meta))
    q = wfdbio.rdrecord(str(root / ('rec%d' % r))).p_signal          # what the pipeline sees after quantisation
    rw = orc.scan_record(q, sig, meta, sig[:3], 'PA', 1.5, -50)
    s, rr, mm = orc.normalise_record(q, sig, sig[:3], rw)
    for k, st in enumerate(rw.rel_start[rw.keep]):
      want[('rec%d' % r, int(st))] = (s[k], rr[k], mm[k])
  c = H.effective_config('waveform_06')
  params = types.SimpleNamespace(dir_path=str(exp), train_path=str(exp / 'loader_train.pickle'),
                                 valid_path=str(exp / 'loader_valid.pickle'), test_path=str(exp / 'loader_test.pickle'),
                                 batch_size=16, split_seed=3, **{k: c[k] for k in ('in_channels', 'chamber', 'segment_size', 'min_RHC', 'use_global_min_max')})
  recordutil.run(params)
  with pytest.raises(Exception, match='Train file already exists!'):
    recordutil.save_dataloaders(params)
  log = (exp / 'record_log.txt').read_text().splitlines()
  counts = {ln.split(':')[0]: int(ln.split(':')[1]) for ln in log[1:]}
  n = len(want)
  assert counts['All segments'] == n and counts['Train segments'] == int(np.floor(0.9 * n))
  assert counts['Valid segments'] + counts['Test segments'] + counts['Train segments'] == n
  train = recordutil.load_dataloader(params.train_path)
  valid = recordutil.load_dataloader(params.valid_path)
  with open(params.test_path, 'rb') as f:
    test = pickle.load(f)                                   # waveform_test.py:115-116
  seen = set()
  # waveform_train.py:357-362: iterate batches, take [0] and [1], .to(device)
  assert len(train) == -(-counts['Train segments'] // 16)
  nb = 0
  for i, segment in enumerate(train):
    scg, rhc = segment[0], segment[1]
    assert scg.is_cuda and rhc.is_cuda and scg.dtype == torch.float32
    assert scg.to('cuda') is scg
    assert scg.shape[1:] == (3, 750) and rhc.shape[1:] == (1, 750) and len(segment) == 7
    for b in range(scg.shape[0]):
      key = (segment[2][b], int(segment[3][b]))
      s, r, mm = want[key]
      assert scg[b].cpu().numpy().tobytes() == s.tobytes() and rhc[b].cpu().numpy().tobytes() == r.tobytes()
      assert int(segment[4][b]) == key[1] + 750 and float(segment[5][0][b]) == mm[0] and float(segment[6][1][b]) == mm[3]
      seen.add(key)
    nb += 1
  assert nb == len(train)
  # waveform_test.py:58-67: iterate loader.dataset, .unsqueeze(0), .detach().numpy(), unpack item[6]
  for loader in (valid, test):
    for segment in loader.dataset:
      scg = segment[0].unsqueeze(0)
      real = segment[1].detach().numpy()[0, :]
      min_rhc, max_rhc = segment[6]
      key = (segment[2], int(segment[3]))
      s, r, mm = want[key]
      assert scg.shape == (1, 3, 750) and real.tobytes() == r[0].tobytes() and (min_rhc, max_rhc) == (mm[2], mm[3])
      assert int(segment[4]) == key[1] + 750
      seen.add(key)
  assert seen == set(want)                                  # every kept window lands in exactly one split


def test_streamed_ingest_equals_eager_ingest(tmp_path, monkeypatch):
  """prepare_cohort streams format-16 cohorts chunk by chunk (parse / plan / read / copy / kernel overlapped,
  engine.LazyDiskIngest); parsing everything first (_prepare_eager) gives the same store — ragged records, records
  without windows, dataset-level pairs; a cohort whose records list their signals differently falls back by itself."""
  root = tmp_path / 'data'; root.mkdir()
  monkeypatch.setattr(recordutil, 'PROCESSED_DATA_PATH', str(root))
  monkeypatch.setattr(recordutil, 'wfdb', wfdbio)
  sig = synth_ref.SIG_NAMES_5
  events = [{'RA_1': 0, 'PA_1': 12, 'RV_1': 70}, {'PA_1': 0}, {'RV_1': 0}, {'PA_1': 3.2, 'PCW_1': 40, 'PA_2': 55}, {'PA_1': 1}, {'RA_1': 3},
            {'PA_1': 0.001}]
  for r, ev in enumerate(events):
    p = synth_ref.gen_record(H.SEED, 700 + r, 30000 + 1111 * r, kinds=synth_ref.kinds_for(sig))
    wfdbio.wrsamp('rec%d' % r, 500, ['g', 'g', 'g', 'mmHg', 'mV'], sig, p, write_dir=str(root))
    (root / ('rec%d.json' % r)).write_text(json.dumps(synth_ref.record_meta(100, events=ev)))
  names = sorted(recordutil.get_record_names())
  dev = torch.device('cuda', torch.cuda.current_device())
  for cfg in ('waveform_06', 'waveform_04', 'waveform_25'):
    c = H.effective_config(cfg)
    params = types.SimpleNamespace(**{k: c[k] for k in ('in_channels', 'chamber', 'segment_size', 'min_RHC', 'use_global_min_max')})
    C = len(params.in_channels)
    for chunk in (1, 2, 3, 50):
      a = recordutil._prepare_streamed(params, names, 0, C, dev, chunk, None)
      b = recordutil._prepare_eager(params, names, names, 0, C, dev, chunk, None)
      assert a is not None and a.n_kept == b.n_kept > 0 and a.n_cand == b.n_cand
      assert torch.equal(a.kept_idx, b.kept_idx) and torch.equal(a.rec_id, b.rec_id) and torch.equal(a.start_idx, b.start_idx)
      assert torch.equal(a.keep, b.keep) and torch.equal(a.kept_minmax(), b.kept_minmax())
      x, y = a.materialise(), b.materialise()
      assert torch.equal(x[0], y[0]) and torch.equal(x[1], y[1])
    monkeypatch.setenv('SCGRHC_LAZY_MIN_RECORDS', '1' if cfg == 'waveform_06' else '4096')   # both ingests through the public entry
    st, _ = recordutil.prepare_cohort(params, record_names=names)
    assert st.n_kept == b.n_kept and torch.equal(st.materialise()[0], y[0])
  # the native header / side-car scanner (scgrhc.hostscan) against the Python parsers, end to end: same store with the scanner
  # off; and with files outside its common shape (a comment line in a header, an event time given as a string) mixed in
  hea = (root / 'rec1.hea').read_text()
  (root / 'rec1.hea').write_text('# exported by a tool that comments\n' + hea)
  (root / 'rec3.json').write_text(json.dumps(synth_ref.record_meta(100, events={'PA_1': '3.2', 'PCW_1': 40, 'PA_2': 55})))
  for streamed in (False, True):
    got = []
    for native in (True, False):
      monkeypatch.setattr(wfdbio, 'NATIVE_SCAN', native)
      got.append(recordutil._prepare_streamed(params, names, 0, C, dev, 3, None) if streamed
                 else recordutil._prepare_eager(params, names, names, 0, C, dev, 3, None))
    a, b2 = got
    assert a.n_kept == b2.n_kept == b.n_kept and torch.equal(a.kept_idx, b2.kept_idx) and torch.equal(a.rec_id, b2.rec_id)
    assert torch.equal(a.kept_minmax(), b2.kept_minmax()) and torch.equal(a.materialise()[0], b2.materialise()[0])
    assert torch.equal(a.materialise()[0], y[0])
  monkeypatch.setattr(wfdbio, 'NATIVE_SCAN', True)
  # a rank whose block of records is empty (more ranks than records) still returns a well-formed, empty store
  empty = recordutil._prepare_eager(params, [], names, len(names), C, dev, 2, None)
  assert empty.n_kept == 0 and empty.n_cand == 0 and empty.kept_idx.numel() == 0 and empty.materialise()[0].shape[0] == 0
  # a record that lists its signals in another order: not one frame layout -> the streamed ingest declines
  p = synth_ref.gen_record(H.SEED, 799, 20000, kinds=synth_ref.kinds_for(sig))
  order = [3, 0, 1, 2, 4]
  wfdbio.wrsamp('rec9', 500, ['mmHg', 'g', 'g', 'g', 'mV'], [sig[i] for i in order], p[:, order], write_dir=str(root))
  (root / 'rec9.json').write_text(json.dumps(synth_ref.record_meta(40, events={'PA_1': 0})))
  names = sorted(recordutil.get_record_names())
  assert recordutil._prepare_streamed(params, names, 0, C, dev, 2, None) is None
  monkeypatch.setenv('SCGRHC_LAZY_MIN_RECORDS', '1')
  st, _ = recordutil.prepare_cohort(params, record_names=names)
  assert st.n_kept > b.n_kept and int(st.rec_id.max()) == len(names) - 1


def test_save_dataloaders_sweep_equals_one_job_per_config(tmp_path, monkeypatch):
  """`recordutil.py prepare d1 d2 ...`: one read + one upload of the cohort, one predicate pass and one fan-out pass
  per chamber — and every config directory ends up with the loaders a separate save_dataloaders job writes."""
  root = tmp_path / 'data'; root.mkdir()
  monkeypatch.setattr(recordutil, 'PROCESSED_DATA_PATH', str(root))
  monkeypatch.setattr(recordutil, 'wfdb', wfdbio)
  sig = synth_ref.SIG_NAMES_5
  for r in range(3):
    meta = synth_ref.record_meta(120, events={'RA_1': 0, 'PA_1': 20 + r, 'RV_1': 100, 'PCW_1': 110})
    p = synth_ref.gen_record(H.SEED, 140 + r, 60000 + 500 * r, kinds=synth_ref.kinds_for(sig))
    wfdbio.wrsamp('rec%d' % r, 500, ['g', 'g', 'g', 'mmHg', 'mV'], sig, p, write_dir=str(root))
    (root / ('rec%d.json' % r)).write_text(json.dumps(meta))
  table = H.configs()
  cfgs = ['waveform_06', 'waveform_07', 'waveform_09', 'waveform_10', 'waveform_25', 'waveform_11', 'waveform_04']
  dirs = {}
  for tag in ('sweep', 'single'):
    for cfg in cfgs:
      d = tmp_path / tag / cfg; d.mkdir(parents=True)
      e = H.effective_config(cfg, table)
      c = dict(in_channels=e['in_channels'], chamber=e['chamber'], segment_size=e['segment_size'], batch_size=e['batch_size'],
               min_RHC=e['min_RHC'], use_global_min_max=e['use_global_min_max'],
               dir_path=str(d), train_path='loader_train.pickle', valid_path='loader_valid.pickle', test_path='loader_test.pickle',
               split_seed=11, comparison_dir_path='comparisons', checkpoint_dir_path='checkpoints',
               pred_top_dir_path='pred_top', pred_rand_dir_path='pred_rand', alpha=1e-4, beta1=0.5, beta2=0.999, n_critic=2,
               lambda_gp=10, lambda_aux=1, total_epochs=1)
      (d / 'params.json').write_text(json.dumps(c))
      dirs[(tag, cfg)] = d
  names = sorted(recordutil.get_record_names())
  monkeypatch.setattr(recordutil, 'get_record_names', lambda: names)     # the reference's order is hash-random (list(set))
  counts = recordutil.prepare_all([str(dirs[('sweep', c)]) for c in cfgs])
  assert all(v for v in counts.values()) and len(counts) == len(cfgs)
  from paramutil import Params
  for cfg in cfgs:
    recordutil.run(Params(str(dirs[('single', cfg)] / 'params.json')))
    for which in ('train', 'valid', 'test'):
      a = recordutil.load_dataloader(str(dirs[('sweep', cfg)] / ('loader_%s.pickle' % which))).dataset
      b = recordutil.load_dataloader(str(dirs[('single', cfg)] / ('loader_%s.pickle' % which))).dataset
      assert len(a) == len(b) > 0, (cfg, which)
      for x, y in zip(a, b):
        assert torch.equal(x[0].cpu(), y[0].cpu()) and torch.equal(x[1].cpu(), y[1].cpu()), (cfg, which)
        assert (x[2], int(x[3]), int(x[4])) == (y[2], int(y[3]), int(y[4]))
        assert tuple(x[5]) == tuple(y[5]) and tuple(x[6]) == tuple(y[6])
    la = (dirs[('sweep', cfg)] / 'record_log.txt').read_text().splitlines()[1:]
    assert la == (dirs[('single', cfg)] / 'record_log.txt').read_text().splitlines()[1:]
  # a second run reports the existing loaders and touches nothing, like the reference's guard
  assert recordutil.prepare_all([str(dirs[('sweep', c)]) for c in cfgs[:2]]) == {}


def test_device_decode_matches_host_dac_and_digital_ingest(tmp_path, monkeypatch):
  """Format-16 frames decoded on the device == wfdb's host-side (d - baseline) / gain, bit for bit; the digital
  ingest path (int16 over PCIe) gives the same windows as the physical one, also with dataset-global pairs."""
  from scgrhc import ops
  root = tmp_path / 'data'; root.mkdir()
  monkeypatch.setattr(recordutil, 'PROCESSED_DATA_PATH', str(root))
  monkeypatch.setattr(recordutil, 'wfdb', wfdbio)
  sig = synth_ref.SIG_NAMES_5
  meta = synth_ref.record_meta(80, events={'PA_1': 1.0, 'RV_1': 70})
  for r in range(3):
    p = synth_ref.gen_record(H.SEED, 80 + r, 40000 + 37 * r, kinds=synth_ref.kinds_for(sig))
    if r == 1:
      p[1234, 3] = np.nan                                  # stored as the invalid code -32768 -> NaN -> flat-free window raises
    wfdbio.wrsamp('rec%d' % r, 500, ['g', 'g', 'g', 'mmHg', 'mV'], sig, p, write_dir=str(root))
    (root / ('rec%d.json' % r)).write_text(json.dumps(meta))
  rec = wfdbio.rdrecord(str(root / 'rec1'))
  d = torch.from_numpy(rec.d_signal.copy()).cuda()
  out = torch.empty((d.shape[0], 3), dtype=torch.float64, device='cuda')
  sel = [4, 0, 3]
  ops.decode_fmt16(d, sel, [rec.adc_gain[j] for j in sel], [float(rec.baseline[j]) for j in sel], out)
  want = rec.p_signal[:, sel]
  got = out.cpu().numpy()
  assert np.isnan(got[1234, 2]) and np.array_equal(np.isnan(got), np.isnan(want))
  assert got[~np.isnan(got)].tobytes() == want[~np.isnan(want)].tobytes()
  c = H.effective_config('waveform_08')                     # lat + dv: a non-trivial column selection
  params = types.SimpleNamespace(in_channels=c['in_channels'], chamber='PA', segment_size=1.5, min_RHC=-50, use_global_min_max=False)
  with pytest.raises(ValueError):                           # the NaN reaches the regression, as in the reference
    recordutil.prepare_cohort(params, ['rec0', 'rec1', 'rec2'])
  for use_global in (False, True):
    params.use_global_min_max = use_global
    dig, _ = recordutil.prepare_cohort(params, ['rec0', 'rec2'], chunk_records=1)
    # same cohort through the physical (fp64 p_signal) path: hide the digital frames from the reader
    monkeypatch.setattr(recordutil, 'wfdb', types.SimpleNamespace(
      rdrecord=lambda path: FakeRecord(wfdbio.rdrecord(path).sig_name, wfdbio.rdrecord(path).p_signal)))
    phy, _ = recordutil.prepare_cohort(params, ['rec0', 'rec2'], chunk_records=1)
    monkeypatch.setattr(recordutil, 'wfdb', wfdbio)
    assert dig.n_kept == phy.n_kept > 0
    assert torch.equal(dig.start_idx, phy.start_idx) and torch.equal(dig.rec_id, phy.rec_id)
    assert torch.equal(dig.kept_minmax(), phy.kept_minmax())
    a, b = dig.materialise(), phy.materialise()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


@pytest.mark.parametrize('nsig_in,sel', [(4, [0, 1, 2, 3]), (4, [3, 1]), (5, [4, 0, 3]), (5, [0, 1, 2, 3, 4]), (3, [1]), (8, [7, 2, 5, 0])])
def test_device_decode_every_shape_is_bit_exact(nsig_in, sel):
  """scgrhc_decode_fmt16 == numpy's (d - baseline) / gain bit for bit: every output column count, 4-sample frames (8-byte
  loads) and others, awkward gains (reciprocal + two FMA corrections must give the IEEE quotient), every int16 code
  incl. the invalid-sample code; a fractional baseline and a subnormal gain take the full-division kernel."""
  from scgrhc import ops
  rng = np.random.default_rng(nsig_in * 31 + len(sel))
  T = 70001
  d = rng.integers(-32768, 32768, size=(T, nsig_in), dtype=np.int64).astype(np.int16)
  d[:65536, sel[0]] = np.arange(-32768, 32768, dtype=np.int64).astype(np.int16)     # every code once
  for gains, bases in (([200000.0, 3.0, 0.1, 1e-3 / 3, 497.3][:len(sel)], [0.0, -7.0, 1024.0, 32767.0, -32768.0][:len(sel)]),
                       ([1e5 / 7.0] * len(sel), [0.5] * len(sel)),
                       ([5e-324 * 4] + [2.0] * (len(sel) - 1), [0.0] * len(sel))):
    out = torch.empty((T, len(sel)), dtype=torch.float64, device='cuda')
    ops.decode_fmt16(torch.from_numpy(d).cuda(), sel, gains, bases, out)
    x = d[:, sel].astype(np.float64)
    with np.errstate(over='ignore'):
      want = (x - np.asarray(bases)) / np.asarray(gains)
    want[d[:, sel] == -32768] = np.nan
    got = out.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert got[~np.isnan(got)].tobytes() == want[~np.isnan(want)].tobytes(), (gains, bases)


@pytest.mark.parametrize('nsig_in,sel', [(4, [0, 1, 2, 3]), (5, [4, 0, 3]), (3, [1]), (8, [7, 2, 5, 0, 1])])
def test_device_decode_per_record_tables_one_launch(nsig_in, sel):
  """scgrhc_decode_fmt16_records: a chunk of ragged records, every record with its own gain / baseline (device tables),
  ONE launch — bit for bit numpy's (d - baseline) / gain per record; both the reciprocal path and (a fractional baseline
  in the table) the IEEE-division path."""
  from scgrhc import ops
  rng = np.random.default_rng(nsig_in * 7 + len(sel))
  rows = [7001, 1, 300, 25000, 64]
  off = np.concatenate([[0], np.cumsum(rows)])
  d = rng.integers(-32768, 32768, size=(off[-1], nsig_in), dtype=np.int64).astype(np.int16)
  for frac in (False, True):
    gain = rng.uniform(0.5, 3e5, size=(len(rows), len(sel)))
    gain[1] = 1e-3 / 3
    base = np.round(rng.uniform(-3e4, 3e4, size=(len(rows), len(sel))))
    if frac:
      base[3, 0] += 0.5
    out = torch.full((off[-1] + 1, len(sel)), -1.0, dtype=torch.float64, device='cuda')
    ops.decode_fmt16_records(torch.from_numpy(d).cuda(), torch.from_numpy(off).cuda(), max(rows), sel, torch.from_numpy(gain).cuda(),
                             torch.from_numpy(base).cuda(), not frac, out[:off[-1]])
    got = out.cpu().numpy()
    assert (got[-1] == -1.0).all()                                    # nothing written past the chunk
    for r in range(len(rows)):
      x = d[off[r]:off[r + 1]][:, sel]
      want = (x.astype(np.float64) - base[r]) / gain[r]
      want[x == -32768] = np.nan
      g = got[off[r]:off[r + 1]]
      assert np.array_equal(np.isnan(g), np.isnan(want)) and g[~np.isnan(g)].tobytes() == want[~np.isnan(want)].tobytes(), (r, frac)


def test_one_launch_collate_equals_the_gathers():
  """scgrhc_collate_batch (one launch for the SCG and RHC batch, plain ctypes) == the two generic gathers, with the noise
  extension == scgrhc_gather_windows_noise; the loader's ring of reused batch buffers and its lazy metadata."""
  from scgrhc import ops
  g = torch.Generator().manual_seed(3)
  for C, W, n_store, nb in ((3, 750, 700, 256), (1, 750, 50, 7), (4, 375, 300, 300), (2, 333, 40, 33)):
    scg_store = torch.rand((n_store, C, W), generator=g).cuda()
    rhc_store = torch.rand((n_store, 1, W), generator=g).cuda()
    slots = torch.randint(0, n_store, (nb + 5,), generator=g).cuda()
    col = ops.BatchCollator(scg_store, rhc_store, slots)
    a, b = torch.full((nb + 1, C, W), -2.0, device='cuda'), torch.full((nb + 1, 1, W), -2.0, device='cuda')
    col(3, nb, a, b)
    assert torch.equal(a[:nb], scg_store[slots[3:3 + nb]]) and torch.equal(b[:nb], rhc_store[slots[3:3 + nb]])
    assert (a[nb] == -2).all() and (b[nb] == -2).all()
    want = torch.empty((nb, C, W), device='cuda')
    ops.gather_windows_noise(scg_store, slots[3:3 + nb].contiguous(), want, 0.07, 1234, 9)
    col(3, nb, a, b, 0.07, 1234, 9)
    assert torch.equal(a[:nb], want) and torch.equal(b[:nb], rhc_store[slots[3:3 + nb]]) and (a[nb] == -2).all()
    with pytest.raises(ValueError):
      col(3, nb + 3, a, b)
  n, C, W = 70, 3, 750
  ds = recordutil.SCGDataset.from_arrays(torch.rand((n, C, W), generator=g).cuda(), torch.rand((n, 1, W), generator=g).cuda(),
                                         ['r%d' % i for i in range(n)], np.arange(n) * W, np.arange(n) * W + W, np.arange(4 * n, dtype=np.float64).reshape(n, 4), 1.5)
  fresh = list(recordutil.WindowLoader(ds, batch_size=32))
  assert [len(x[2]) for x in fresh] == [32, 32, 6] and torch.equal(torch.cat([x[0] for x in fresh]), ds.scg)
  assert fresh[2][2] == ('r64', 'r65', 'r66', 'r67', 'r68', 'r69') and fresh[1][3].tolist() == (np.arange(32, 64) * W).tolist()
  assert float(fresh[2][6][1][5]) == 4 * 69 + 3 and len(fresh[0]) == 7 and pickle.loads(pickle.dumps(fresh[0]))[2][0] == 'r0'
  ring = recordutil.WindowLoader(ds, batch_size=32, reuse_buffers=2)
  ptrs = []
  for k, batch in enumerate(ring):
    assert torch.equal(batch[0], ds.scg[32 * k:32 * k + 32]) and torch.equal(batch[1], ds.rhc[32 * k:32 * k + 32])
    ptrs.append(batch[0].data_ptr())
  assert ptrs[0] == ptrs[2] != ptrs[1]                               # batch k + R reuses the buffer of batch k
  assert len(pickle.dumps(ring)) < len(pickle.dumps(ds)) + 4096      # the ring is not pickled


def test_noise_injection_extension_philox():
  """Extension (absent from the reference): seed-exact Philox4x32-10 stream, Box-Muller within fp32 tolerance,
  distribution checks, and the loader wiring (SCG inputs only, a fresh stream per batch, off by default)."""
  from oracle import philox_ref
  from scgrhc import ops
  for seed, offset in ((0, 0), (0x5C6, 7), (0xDEADBEEFCAFEF00D, 0x1_0000_0003)):
    got = ops.philox_words(0, seed, offset, 4096)
    assert (got == philox_ref.words(seed, offset, 4096).astype(np.int64)).all()
  n, C, W = 300, 3, 750
  g = torch.Generator().manual_seed(1)
  store = torch.rand((n, C, W), generator=g).cuda()
  slots = torch.randperm(n, generator=g)[:256].cuda()
  out = torch.empty((256, C, W), dtype=torch.float32, device='cuda')
  sigma = 0.05
  ops.gather_windows_noise(store, slots, out, sigma, 99, 5)
  z = philox_ref.normals(99, 5, 256 * C * W).reshape(256, C, W)
  want = store[slots].cpu().numpy() + np.float32(sigma) * z
  assert np.abs(out.cpu().numpy() - want).max() < 2e-6
  resid = ((out - store[slots]) / sigma).double().flatten().cpu().numpy()
  assert abs(resid.mean()) < 5e-3 and abs(resid.std() - 1) < 5e-3
  from scipy import stats
  assert stats.kstest(resid[::7], 'norm').pvalue > 1e-3
  assert abs(stats.skew(resid)) < 0.02 and abs(stats.kurtosis(resid)) < 0.05
  # loader wiring
  ds = recordutil.SCGDataset.from_arrays(store, torch.rand((n, 1, W)).cuda(), ['r'] * n, np.arange(n) * W, np.arange(n) * W + W,
                                         np.zeros((n, 4)), 1.5)
  plain = list(recordutil.WindowLoader(ds, batch_size=128))
  noisy = recordutil.WindowLoader(ds, batch_size=128, noise_std=0.1, noise_seed=4)
  a, b = list(noisy), list(noisy)
  assert torch.equal(plain[0][0], store[:128]) and torch.equal(a[0][1], plain[0][1])          # targets untouched
  d0 = (a[0][0] - plain[0][0]).std().item()
  assert 0.09 < d0 < 0.11 and not torch.equal(a[0][0], b[0][0]) and not torch.equal(a[0][0][:44], a[1][0][:44])


def test_evaluation_metrics_on_device_match_consumer_formulas():
  """waveform_test.py:21-50 (reverse_minmax, scipy pearsonr, sqrt(sklearn MSE)) per window, on the device."""
  from scipy.stats import pearsonr
  from sklearn.metrics import mean_squared_error
  from scgrhc import ops
  g = torch.Generator().manual_seed(3)
  n, W = 64, 750
  real = torch.rand((n, 1, W), generator=g)
  pred = (real + 0.1 * torch.randn((n, 1, W), generator=g)).clamp(0, 1)
  pred[5] = real[5]                                    # r == 1 exactly
  mm = np.stack([np.linspace(-5, 20, n), np.linspace(30, 90, n)], axis=1)
  out = torch.empty((n, 2), dtype=torch.float64, device='cuda')
  ops.window_metrics(real.cuda(), pred.cuda(), torch.from_numpy(mm).cuda(), out)
  got = out.cpu().numpy()
  for i in range(n):
    x = real[i].numpy()[0, :] * (np.float64(mm[i, 1]) - np.float64(mm[i, 0])) + np.float64(mm[i, 0])   # reverse_minmax
    y = pred[i].numpy()[0, :] * (np.float64(mm[i, 1]) - np.float64(mm[i, 0])) + np.float64(mm[i, 0])
    assert abs(got[i, 0] - pearsonr(x, y).statistic) < 1e-12
    assert abs(got[i, 1] - np.sqrt(mean_squared_error(x, y))) < 1e-12 * max(1.0, got[i, 1])
