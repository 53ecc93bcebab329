"""CPU, build container only: the reference's UNMODIFIED orchestrator and consumers (waveform_pipeline.run ->
waveform_train.run, waveform_test.run, waveform_checkpoint.run — SURVEY.md §8b) run on loaders that this repo's GPU
recordutil produced (tests/golden/loaders/*.pickle, made on a B200 by tools/make_loader_fixture.py)."""
import os
import shutil
import sys
import types

import pytest
import torch

from oracle.ref_harness import REFERENCE_PATH, reference_available

HERE = os.path.dirname(os.path.abspath(__file__))
LOADERS = os.path.join(HERE, 'golden', 'loaders')


@pytest.mark.skipif(not reference_available(), reason='/root/reference is not present on this machine')
def test_reference_train_test_checkpoint_on_our_loaders(tmp_path, monkeypatch):
  exp = tmp_path / 'waveform_06'
  exp.mkdir()
  for f in ('loader_train.pickle', 'loader_valid.pickle', 'loader_test.pickle'):
    shutil.copy(os.path.join(LOADERS, f), exp / f)
  # matplotlib is absent from the image and imported at module scope by the reference's scripts
  plt = types.ModuleType('matplotlib.pyplot')
  for fn in ('plot', 'title', 'xlabel', 'ylabel', 'ylim', 'legend', 'savefig', 'close'):
    setattr(plt, fn, lambda *a, **k: None)
  mpl = types.ModuleType('matplotlib'); mpl.pyplot = plt
  monkeypatch.setitem(sys.modules, 'matplotlib', mpl)
  monkeypatch.setitem(sys.modules, 'matplotlib.pyplot', plt)
  # our drop-in directory first (conftest), the reference second: exactly the INTEGRATION.md layout
  monkeypatch.syspath_prepend(REFERENCE_PATH)
  pkg = os.path.join(os.path.dirname(HERE), 'scg-rhc-waveform_b200')
  monkeypatch.syspath_prepend(pkg)
  for name in ('recordutil', 'paramutil', 'waveform_noise', 'pathutil', 'timelog', 'waveform_train', 'waveform_test', 'waveform_checkpoint',
               'waveform_pipeline'):
    sys.modules.pop(name, None)
  import recordutil
  assert os.path.dirname(recordutil.__file__) == pkg
  import waveform_train, waveform_test, waveform_checkpoint
  assert os.path.dirname(waveform_train.__file__) == REFERENCE_PATH

  params = types.SimpleNamespace(
    dir_path=str(exp), in_channels=['patch_ACC_lat', 'patch_ACC_hf', 'patch_ACC_dv'], chamber='PA', segment_size=1.5,
    batch_size=16, train_path=str(exp / 'loader_train.pickle'), valid_path=str(exp / 'loader_valid.pickle'),
    test_path=str(exp / 'loader_test.pickle'), checkpoint_dir_path=str(exp / 'checkpoints'),
    comparison_dir_path=str(exp / 'comparisons'), pred_top_dir_path=str(exp / 'pred_top'),
    pred_rand_dir_path=str(exp / 'pred_rand'), alpha=1e-4, beta1=0.5, beta2=0.999, n_critic=1, lambda_gp=10,
    lambda_aux=100, total_epochs=1, min_RHC=-50, use_global_min_max=False)
  torch.manual_seed(0)
  train = recordutil.load_dataloader(params.train_path)
  assert len(train.dataset) == 62 and len(train) == 4
  # the reference's own orchestrator (this repo ships no copy): its `from recordutil import run` resolves to OUR module;
  # the loaders exist, so data preparation reports "Train file already exists!" (the benign branch of
  # waveform_pipeline.py:12-15) and the reference's train -> validate -> select -> test sequence runs on our batches
  import waveform_pipeline
  assert os.path.dirname(waveform_pipeline.__file__) == REFERENCE_PATH and waveform_pipeline.recordutil is recordutil.run
  waveform_pipeline.run(params)
  assert os.path.exists(exp / 'checkpoints' / '000.checkpoint')
  best = (exp / 'checkpoint_best.txt').read_text().splitlines()[0].split()[1]
  assert best == '000.checkpoint' or best.startswith('000')
  import pandas as pd
  df = pd.read_csv(exp / 'comparisons' / 'test' / '000.csv')
  assert len(df) == 4 and set(df['filename']) <= {'rec0', 'rec1'} and (df['stop_idx'] - df['start_idx'] == 750).all()
