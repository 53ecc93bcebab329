"""Imported by test_host_logic: recordutil's native split arithmetic vs sklearn.train_test_split under the
same numpy global RandomState (the reference is unseeded, recordutil.py:191-192)."""
import ast
import os

import numpy as np
from sklearn.model_selection import train_test_split

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# recordutil imports waveform_noise -> CUDA at call time only; the split helper is pure numpy, so lift it
# out of the module source without importing the GPU-dependent module graph.
src = open(os.path.join(ROOT, 'scg-rhc-waveform_b200', 'recordutil.py')).read()
fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'train_valid_test_split')
ns = {'np': np}
exec(compile(ast.Module(body=[fn], type_ignores=[]), 'recordutil.py', 'exec'), ns)
split = ns['train_valid_test_split']

for n in (20, 21, 37, 1000, 4321):
  items = list(range(n))
  np.random.seed(n)
  tr, rest = train_test_split(items, train_size=0.9)
  va, te = train_test_split(rest, train_size=0.5)
  np.random.seed(n)
  a, b, c = split(n)
  assert a.tolist() == tr and b.tolist() == va and c.tolist() == te, n
  a2, b2, c2 = split(n, seed=5)
  a3, b3, c3 = split(n, seed=5)
  assert a2.tolist() == a3.tolist() and sorted(a2.tolist() + b2.tolist() + c2.tolist()) == items
