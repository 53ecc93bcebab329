"""Drop-in for the reference's ``pathutil`` (pathutil.py:1-19).  The data root is the one seam the
window-preparation path needs; ``SCG_RHC_DATA_PATH`` overrides the author's absolute path."""
import os
import shutil

DATA_PATH = os.environ.get('SCG_RHC_DATA_PATH', os.path.join('/', 'home', 'jesse', 'scg-rhc-database'))

PROCESSED_DATA_PATH = os.path.join(DATA_PATH, 'processed_data')


def clear(paths):
  """Empty each existing directory (pathutil.py:9-14)."""
  for path in paths:
    if os.path.exists(path):
      shutil.rmtree(path)
      os.makedirs(path)
      print(f'Cleared {path}')


def clear_comparisons_valid():
  """pathutil.py:17-19."""
  clear([os.path.join(p, 'comparisons', 'valid') for p in sorted(os.listdir(os.getcwd()))])
