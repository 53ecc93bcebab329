"""Multi-GPU product path: 2-way record sharding through the drop-in gives the same windows, in the
same order, as 1-way, the same dataset-level min/max bits, and the same pickled loaders.

The ranks are real processes under torch.distributed.run (tests/dist_gpu_worker.py): NCCL with one GPU per rank when the
box has two GPUs, else both ranks on cuda:0 with gloo carrying the (tiny) collectives."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import synth_ref
from tests import helpers as H

pytestmark = pytest.mark.gpu

import recordutil  # noqa: E402
from paramutil import Params  # noqa: E402
from scgrhc import wfdbio  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORLD = 2


def _params_dir(base, tag, cfg, **extra):
  d = base / tag
  d.mkdir(parents=True)
  e = H.effective_config(cfg)
  c = dict(in_channels=e['in_channels'], chamber=e['chamber'], segment_size=e['segment_size'], batch_size=e['batch_size'],
           min_RHC=e['min_RHC'], use_global_min_max=e['use_global_min_max'], dir_path=str(d), train_path='loader_train.pickle',
           valid_path='loader_valid.pickle', test_path='loader_test.pickle', split_seed=5, comparison_dir_path='comparisons',
           checkpoint_dir_path='checkpoints', pred_top_dir_path='pred_top', pred_rand_dir_path='pred_rand', alpha=1e-4,
           beta1=0.5, beta2=0.999, n_critic=2, lambda_gp=10, lambda_aux=1, total_epochs=1)
  c.update(extra)
  (d / 'params.json').write_text(json.dumps(c))
  return str(d)


def _same_loader(a_path, b_path):
  a, b = recordutil.load_dataloader(a_path).dataset, recordutil.load_dataloader(b_path).dataset
  assert len(a) == len(b) > 0
  for x, y in zip(a, b):
    assert torch.equal(x[0].cpu(), y[0].cpu()) and torch.equal(x[1].cpu(), y[1].cpu())
    assert (x[2], int(x[3]), int(x[4])) == (y[2], int(y[3]), int(y[4])) and tuple(x[5]) == tuple(y[5]) and tuple(x[6]) == tuple(y[6])


def test_two_rank_sharding_equals_one_rank(tmp_path, monkeypatch):
  root = tmp_path / 'data'; root.mkdir()
  out = tmp_path / 'out'; out.mkdir()
  sig = synth_ref.SIG_NAMES_5
  events = [{'RA_1': 0, 'PA_1': 12, 'RV_1': 70}, {'PA_1': 0}, {'RV_1': 0}, {'PA_1': 3.2, 'PCW_1': 40, 'PA_2': 55}, {'PA_1': 1}]
  for r, ev in enumerate(events):                       # ragged lengths; record 2 has no PA interval at all
    p = synth_ref.gen_record(H.SEED, 300 + r, 45000 + 1250 * r, kinds=synth_ref.kinds_for(sig))
    wfdbio.wrsamp('rec%d' % r, 500, ['g', 'g', 'g', 'mmHg', 'mV'], sig, p, write_dir=str(root))
    (root / ('rec%d.json' % r)).write_text(json.dumps(synth_ref.record_meta(100, events=ev)))
  jobs = {'prepare': {'local': _params_dir(tmp_path, 'p06', 'waveform_06'), 'global': _params_dir(tmp_path, 'p04', 'waveform_04')},
          'save': [_params_dir(tmp_path, 'multi/s06', 'waveform_06'), _params_dir(tmp_path, 'multi/s04', 'waveform_04'),
                   _params_dir(tmp_path, 'multi/s06_sharded', 'waveform_06', train_layout='sharded')],
          'sweep': [_params_dir(tmp_path, 'multi/w%s' % c[-2:], c) for c in ('waveform_06', 'waveform_07', 'waveform_25', 'waveform_04')]}
  (out / 'jobs.json').write_text(json.dumps(jobs))
  port = str(H.free_port())
  env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT=port)
  run = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=%d' % WORLD,
                        '--master-addr', '127.0.0.1', '--master-port', port, os.path.join(ROOT, 'tests', 'dist_gpu_worker.py'),
                        str(root), str(out)], capture_output=True, text=True, env=env, timeout=900)
  assert run.returncode == 0 and 'MULTI_OK' in run.stdout, run.stdout[-3000:] + run.stderr[-3000:]

  monkeypatch.setattr(recordutil, 'PROCESSED_DATA_PATH', str(root))
  monkeypatch.setattr(recordutil, 'wfdb', wfdbio)
  # ---- prepare_cohort: the ranks' stores, concatenated in rank order, ARE the single-process store
  for tag, pdir in jobs['prepare'].items():
    store, names = recordutil.prepare_cohort(Params(os.path.join(pdir, 'params.json')), chunk_records=2)
    scg, rhc = store.materialise()
    parts = [np.load(out / ('prep_%s_rank%d.npz' % (tag, r))) for r in range(WORLD)]
    assert [int(p['offset']) for p in parts] == [0, int(parts[0]['counts'][0])] and int(parts[0]['total']) == store.n_kept > 0
    assert all(len(p['rec_id']) == int(p['counts'][r]) for r, p in enumerate(parts))
    cat = lambda k: np.concatenate([p[k] for p in parts])
    assert (cat('rec_id') == store.rec_id.cpu().numpy()).all() and (cat('start') == store.start_idx.cpu().numpy()).all()
    assert cat('scg').tobytes() == scg.cpu().numpy().tobytes() and cat('rhc').tobytes() == rhc.cpu().numpy().tobytes()
    assert cat('mm').tobytes() == store.kept_minmax().cpu().numpy().tobytes()
    if tag == 'global':
      assert all(p['gmm'].tobytes() == store.global_minmax.cpu().numpy().tobytes() for p in parts)
  # ---- save_dataloaders: the pickles rank 0 wrote == the pickles of a single-process job with the same split_seed
  singles = {'s06': _params_dir(tmp_path, 'single/s06', 'waveform_06'), 's04': _params_dir(tmp_path, 'single/s04', 'waveform_04')}
  for tag, pdir in singles.items():
    recordutil.run(Params(os.path.join(pdir, 'params.json')))
    for which in ('train', 'valid', 'test'):
      _same_loader(os.path.join(str(tmp_path / 'multi' / tag), 'loader_%s.pickle' % which), os.path.join(pdir, 'loader_%s.pickle' % which))
    assert (tmp_path / 'multi' / tag / 'record_log.txt').read_text().splitlines()[1:] == \
        open(os.path.join(pdir, 'record_log.txt')).read().splitlines()[1:]
  # sharded train layout: per-rank loaders next to a manifest; a single process sees the same SET of train windows
  sh_dir = tmp_path / 'multi' / 's06_sharded'
  assert all((sh_dir / ('loader_train.pickle.rank%d-of-%d' % (r, WORLD))).exists() for r in range(WORLD))
  a = recordutil.load_dataloader(str(sh_dir / 'loader_train.pickle')).dataset
  b = recordutil.load_dataloader(os.path.join(singles['s06'], 'loader_train.pickle')).dataset
  key = lambda it: (it[2], int(it[3]), H.sha(it[0].cpu().numpy()), H.sha(it[1].cpu().numpy()))
  assert sorted(key(it) for it in a) == sorted(key(it) for it in b) and len(a) > 0
  for which in ('valid', 'test'):
    _same_loader(str(sh_dir / ('loader_%s.pickle' % which)), os.path.join(singles['s06'], 'loader_%s.pickle' % which))
  # ---- the sweep, records sharded: every config's loaders == a single-process sweep's
  sw = [_params_dir(tmp_path, 'single/w%s' % c[-2:], c) for c in ('waveform_06', 'waveform_07', 'waveform_25', 'waveform_04')]
  counts = recordutil.prepare_all(sw)
  multi_counts = json.loads((out / 'sweep_counts.json').read_text())
  assert sorted(counts.values()) == sorted(multi_counts.values()) and all(v for v in counts.values())
  for m_dir, s_dir in zip(jobs['sweep'], sw):
    for which in ('train', 'valid', 'test'):
      _same_loader(os.path.join(m_dir, 'loader_%s.pickle' % which), os.path.join(s_dir, 'loader_%s.pickle' % which))
  # ---- dataset-level min/max over two ranks == the unmodified reference's value (fixture), and so are the windows
  gold = H.load_json('records_full.json')['configs']['waveform_04']
  for r in range(WORLD):
    got = json.loads((out / ('golden04_rank%d.json' % r)).read_text())
    want = gold['records']['rec%d' % r]
    assert got['gmm_hex'] == gold['global_minmax_hex']
    assert got['start'] == want['start'] and got['scg_sha'] == want['scg_sha'] and got['rhc_sha'] == want['rhc_sha']
