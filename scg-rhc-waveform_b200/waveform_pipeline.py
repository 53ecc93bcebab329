"""Drop-in for the reference's ``waveform_pipeline`` (waveform_pipeline.py:1-40): prepare -> train ->
validate every checkpoint -> score -> test the best.  Pure orchestration, no arithmetic: data preparation
is this repo's GPU ``recordutil.run``; training / evaluation / model selection are the reference's own,
unmodified ``waveform_train`` / ``waveform_test`` / ``waveform_checkpoint`` modules, imported from wherever
the user keeps them (put this directory *before* the reference on PYTHONPATH — INTEGRATION.md).
"""
import os
import sys

from paramutil import Params
from recordutil import run as recordutil, SCGDataset  # noqa: F401  (SCGDataset must be importable to unpickle loaders)


def _consumers():
  """The reference's trainer / evaluator / selector (waveform_pipeline.py:5-7), resolved at call time so
  that data preparation alone does not need them."""
  from waveform_train import run as waveform_train
  from waveform_test import run as waveform_test
  from waveform_checkpoint import run as waveform_checkpoint
  return waveform_train, waveform_test, waveform_checkpoint


def run(params):
  waveform_train, waveform_test, waveform_checkpoint = _consumers()

  try:
    recordutil(params)
  except Exception as e:        # "Train file already exists!" is benign here (waveform_pipeline.py:12-15)
    print(e)

  waveform_train(params)

  try:
    waveform_test(params, 'valid', 'all')
  except Exception as e:
    print(e)

  waveform_checkpoint(params)

  with open(os.path.join(params.dir_path, 'checkpoint_best.txt'), 'r') as f:
    best_checkpoint = f.read().splitlines()[0].split()[1]
    waveform_test(params, 'test', best_checkpoint)


def prepare_all(dir_names):
  """Extension: data preparation only, for several experiment directories (the 37-config sweep of BASELINE.json
  configs[4]).  The cohort is read and uploaded once; configs that differ only in their channel subset share one
  predicate pass and one fan-out pass (recordutil.save_dataloaders_sweep)."""
  from recordutil import save_dataloaders_sweep
  return save_dataloaders_sweep([Params(os.path.join(d, 'params.json')) for d in dir_names])


if __name__ == '__main__':
  dir_name = sys.argv[1]
  if dir_name == 'all':
    for i in range(6, 34):      # the reference's own sweep range (waveform_pipeline.py:34)
      dir_name = f'waveform_{i:02d}'
      params = Params(os.path.join(dir_name, 'params.json'))
      run(params)
  elif dir_name == 'prepare':
    prepare_all(sys.argv[2:])
  else:
    params = Params(os.path.join(dir_name, 'params.json'))
    run(params)
