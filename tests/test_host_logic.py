"""CPU tests: C-ABI exports, host planner, params loader, WFDB I/O, split arithmetic, gloo all-reduce."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import scgrhc_oracle as orc
from oracle import synth_ref
from tests import helpers as H

import scgrhc  # noqa: E402
from scgrhc import _native as N  # noqa: E402
from scgrhc import wfdbio  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
  with open(os.path.join(ROOT, 'include', 'scgrhc.h')) as f:
    declared = set(re.findall(r'^(?:int|void|int64_t|const char\*)\s+(scgrhc_[a-z0-9_]+)\s*\(', f.read(), flags=re.M))
  assert declared == set(N.SYMBOLS), declared ^ set(N.SYMBOLS)
  lib = ctypes.CDLL(scgrhc._native._build.LIB)
  for name in declared:
    assert hasattr(lib, name), name
  assert lib.scgrhc_abi_version() == 1


def test_no_cpu_fallback_without_a_device():
  if torch.cuda.is_available():
    pytest.skip('a CUDA device is present')
  lib = N.lib()
  h = ctypes.c_void_p()
  assert lib.scgrhc_ctx_create(0, ctypes.byref(h)) == N.ERR_NO_DEVICE
  assert b'no CPU fallback' in lib.scgrhc_last_error(None)
  with pytest.raises(NotImplementedError):
    torch.ops.scgrhc.rolling_range_lt(torch.zeros(8, dtype=torch.float64), 4, 1e-3, torch.zeros(8, dtype=torch.uint8))
  import waveform_noise
  with pytest.raises(RuntimeError):
    waveform_noise.get_flat_lines(np.zeros(100))


def test_planner_matches_reference_intervals_and_slicing():
  cases = H.load_json('intervals.json')
  for name, case in cases.items():
    for chamber, want in case['intervals'].items():
      iv, n, bounds = scgrhc.plan_record(case['meta'], chamber, 0, 750)
      assert [list(b) for b in bounds] == want, (name, chamber)
      for T in (0, 1, 749, 750, 100000, 299999, 600500):
        iv, n, _ = scgrhc.plan_record(case['meta'], chamber, T, 750, rec_base_row=123, rec_id=4, cand_base=9)
        a, r, w = orc.candidate_windows([tuple(b) for b in want], T, 750)
        mine = np.concatenate([i['row0'] - 123 + np.arange(i['n_win']) * 750 for i in iv]) if len(iv) else np.zeros(0)
        assert n == len(a) and (mine == a).all(), (name, chamber, T)
        if len(iv):
          assert iv['cand0'][0] == 9 and (np.diff(iv['cand0']) == iv['n_win'][:-1]).all() and (iv['rec_id'] == 4).all()


def test_plan_uniform_equals_plan_cohort():
  meta = synth_ref.record_meta(600)
  a = scgrhc.plan_uniform(meta, 'PA', 300000, 750, 5, rec0=3)
  b = scgrhc.plan_cohort([meta] * 5, 'PA', [300000] * 5, 750)
  assert a.n_cand == b.n_cand == 1000
  assert (a.intervals['row0'] == b.intervals['row0']).all() and (a.intervals['cand0'] == b.intervals['cand0']).all()
  assert (a.intervals['rec_id'] == b.intervals['rec_id'] + 3).all()


def test_plan_cohort_one_call_equals_the_per_record_planner():
  """scgrhc_plan_cohort (one C call per cohort) == scgrhc_plan_record record by record, on a ragged cohort: the golden
  side-cars (unsorted events, duplicate chambers, a non-dict ChamEvents_in_s -> no intervals), records of different
  lengths, strides and sampling rates, chamber '*'."""
  cases = H.load_json('intervals.json')
  metas = [c['meta'] for c in cases.values()] * 2
  rows = [1000 + 37111 * (i % 9) for i in range(len(metas))]
  for chamber, W, stride, fs in (('PA', 750, 0, 0.0), ('RV', 375, 187, 250.0), ('*', 500, 0, 0.0), ('PCW', 100, 250, 0.0)):
    plan = scgrhc.plan_cohort(metas, chamber, rows, W, stride=stride, fs=fs)
    ivs, base, cand = [], 0, 0
    for r, meta in enumerate(metas):
      iv, n, _ = scgrhc.plan_record(meta, chamber, rows[r], W, base, r, cand, stride, fs)
      ivs.append(iv); base += rows[r]; cand += n
    want = np.concatenate(ivs)
    assert plan.n_cand == cand and plan.intervals.tobytes() == want.tobytes(), (chamber, W)
  assert scgrhc.plan_cohort([], 'PA', [], 750).n_cand == 0


def test_params_loads_all_37_configs(tmp_path):
  from paramutil import Params
  table = H.configs()
  for cfg, c in table.items():
    data = {k: c[k] for k in ('in_channels', 'chamber', 'segment_size', 'batch_size', 'min_RHC', 'use_global_min_max') if k in c}
    for k in c['all_keys']:
      data.setdefault(k, 1 if k not in ('dir_path', 'train_path', 'valid_path', 'test_path', 'checkpoint_dir_path',
                                         'comparison_dir_path', 'pred_top_dir_path', 'pred_rand_dir_path') else k)
    data['dir_path'] = cfg
    p = tmp_path / (cfg + '.json')
    p.write_text(json.dumps(data))
    params = Params(str(p))
    assert params.in_channels == c['in_channels'] and params.segment_size == 1.5
    assert params.train_path == os.path.join(cfg, 'train_path')
    if c['reference_params_error'] is None:
      assert params.chamber == c['chamber'] and params.min_RHC == -50 and params.use_global_min_max is False
      Params(str(p), strict=True)
    else:
      with pytest.raises(KeyError):                 # the reference's own loader rejects these five
        Params(str(p), strict=True)
      assert params.min_RHC == float('-inf') and params.use_global_min_max is False
      assert params.chamber == c.get('chamber', '*')


def test_wfdbio_roundtrip(tmp_path):
  sig = synth_ref.SIG_NAMES_5
  p = synth_ref.gen_record(H.SEED, 9, 5000, kinds=synth_ref.kinds_for(sig))
  d, gain, base = wfdbio.wrsamp('r9', 500, ['g'] * 3 + ['mmHg', 'mV'], sig, p, write_dir=str(tmp_path))
  rec = wfdbio.rdrecord(str(tmp_path / 'r9'))
  assert rec.sig_name == sig and rec.fs == 500 and rec.p_signal.shape == p.shape
  assert (rec.d_signal == d).all()
  want = (d.astype(np.float64) - np.array(base, dtype=np.float64)) / np.array(gain)
  assert rec.p_signal.tobytes() == want.tobytes()
  assert np.abs(rec.p_signal - p).max() <= 0.5 / min(gain) * 1.0001


def test_split_arithmetic_matches_sklearn():
  import runpy
  runpy.run_path(os.path.join(ROOT, 'tests', 'recordutil_split_probe.py'))


def test_gloo_minmax_allreduce_and_sharding():
  script = os.path.join(ROOT, 'tests', 'dist_worker.py')
  from tests import helpers as H
  port = str(H.free_port())
  env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT=port)
  out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', port, script],
                       capture_output=True, text=True, env=env, timeout=300)
  assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
  assert 'DIST_OK' in out.stdout


def test_planner_property_random_sidecars():
  """Randomised side-cars (unsorted, duplicated, negative, fractional, past-the-end times; random T / W / stride / rate):
  the C planner enumerates exactly the windows the oracle's restatement of recordutil.py:93-110,138-146 does."""
  from hypothesis import given, settings, strategies as st

  times = st.one_of(st.integers(-50, 700), st.floats(-20, 700, allow_nan=False, width=32).map(lambda v: round(v, 4)))
  keys = st.sampled_from(['PA_1', 'PA_2', 'PA_3', 'RV_1', 'RA_1', 'RA_2', 'PCW_1', 'PAX_1', 'pa_1', 'PA', 'X'])

  @settings(max_examples=300, deadline=None)
  @given(st.dictionaries(keys, times, max_size=7), st.integers(0, 86399), st.integers(0, 86399), st.integers(0, 400000),
         st.sampled_from([1, 2, 50, 375, 750, 1000]), st.sampled_from([0, 1, 100, 750, 1500]), st.sampled_from([0.0, 250.0, 500.0]),
         st.sampled_from(['PA', 'RA', 'RV', 'PCW']))
  def check(events, t_start, t_end, T, W, stride, fs, chamber):
    fmt = lambda s: 'd %02d:%02d:%02d' % (s // 3600, (s // 60) % 60, s % 60)
    meta = {'MacStTime': fmt(t_start), 'MacEndTime': fmt(t_end), 'ChamEvents_in_s': events}
    rate = fs or 500.0
    orc_fs = orc.SAMPLE_FREQ
    try:
      orc.SAMPLE_FREQ = rate if rate != 500.0 else 500
      ivals = orc.chamber_intervals(meta, chamber)
    finally:
      orc.SAMPLE_FREQ = orc_fs
    a, r, w = orc.candidate_windows(ivals, T, W, stride or None)
    iv, n, bounds = scgrhc.plan_record(meta, chamber, T, W, rec_base_row=7, rec_id=3, cand_base=11, stride=stride, fs=fs)
    assert [tuple(b) for b in bounds] == ivals
    st_ = stride or W
    mine = np.concatenate([i['row0'] - 7 + np.arange(i['n_win']) * st_ for i in iv]) if len(iv) else np.zeros(0, np.int64)
    assert n == len(a) and (mine == a).all()
    assert (iv['n_win'] > 0).all() and (len(iv) == 0 or iv['cand0'][0] == 11)

  check()


def test_params_optional_extension_keys_default_off(tmp_path):
  from paramutil import Params
  base = dict(in_channels=['patch_ACC_lat'], chamber='PA', segment_size=1.5, batch_size=8, dir_path='w', train_path='a',
              valid_path='b', test_path='c', checkpoint_dir_path='d', comparison_dir_path='e', pred_top_dir_path='f',
              pred_rand_dir_path='g', alpha=1, beta1=1, beta2=1, n_critic=1, lambda_gp=1, lambda_aux=1, total_epochs=1,
              min_RHC=-50, use_global_min_max=False)
  p = tmp_path / 'p.json'
  p.write_text(json.dumps(base))
  q = Params(str(p), strict=True)
  for k in ('split_seed', 'segment_stride', 'noise_std', 'noise_seed', 'bandpass', 'bandpass_order', 'bandpass_sos',
            'bandpass_mode', 'resample_rate', 'resample_mode', 'normalisation'):
    assert getattr(q, k) is None
  p.write_text(json.dumps(dict(base, bandpass=[1, 40], resample_rate=250, segment_stride=0.5, noise_std=0.01)))
  q = Params(str(p))
  assert q.bandpass == [1, 40] and q.resample_rate == 250 and q.segment_stride == 0.5 and q.noise_std == 0.01


def test_load_dataloader_reads_pickles_written_by_the_reference():
  """A user switching over has loader pickles made by the reference's own recordutil: they must keep loading."""
  for name in ('recordutil', 'waveform_noise'):
    sys.modules.pop(name, None)
  import recordutil
  assert os.path.dirname(recordutil.__file__).endswith('scg-rhc-waveform_b200')
  loader = recordutil.load_dataloader(os.path.join(H.GOLDEN, 'reference_loader.pickle'))
  g = np.load(os.path.join(H.GOLDEN, 'record_small.npz'))
  ds = loader.dataset
  assert isinstance(ds, recordutil.SCGDataset) and len(ds) == len(g['waveform_19.start']) == 3
  for i, item in enumerate(ds):
    assert item[0].numpy().tobytes() == g['waveform_19.scg'][i].tobytes() and item[3] == g['waveform_19.start'][i]
    assert [item[5][0], item[5][1], item[6][0], item[6][1]] == g['waveform_19.minmax'][i].tolist()
  batches = list(loader)                                   # the reference's torch DataLoader object, our dataset class
  assert sum(b[0].shape[0] for b in batches) == 3 and batches[0][0].shape[1:] == (3, 750)
  col = ds.collate(torch.tensor([2, 0]))
  assert col[0].shape == (2, 3, 750) and col[3].tolist() == [int(g['waveform_19.start'][2]), int(g['waveform_19.start'][0])]
