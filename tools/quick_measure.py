"""Builder-side probe (not part of bench.py): synth generator speed, predicate-only pass, collate per batch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import numpy as np, torch
import scgrhc, recordutil
from scgrhc import ops, _native as N

dev = torch.device('cuda:0')
def timed(fn, reps=5):
  fn(); torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps): fn()
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b) / reps

T, n_rec = 300000, 1000
arena = torch.empty((n_rec * T, 4), dtype=torch.float64, device=dev)
print('synth 1000 records ms', timed(lambda: ops.synth_records(arena, 0x5C6, 0, n_rec, T, [0, 1, 2, 3], 16, 750)))
meta = {'MacStTime': '1/1/2020 10:00:00', 'MacEndTime': '1/1/2020 10:10:00', 'ChamEvents_in_s': {'PA_1': 0}}
plan = scgrhc.plan_uniform(meta, 'PA', T, 750, n_rec)
bufs = {}
print('pred-only pass ms', timed(lambda: scgrhc.prepare_windows(arena, plan, [0, 1, 2], 3, -50.0, predicates_only=True, buffers=bufs, check=False)))
st = scgrhc.prepare_windows(arena, plan, [0, 1, 2], 3, -50.0, buffers=bufs)
print('kept', st.n_kept)
scg, rhc = st.materialise()
n = st.n_kept
ds = recordutil.SCGDataset.from_arrays(scg, rhc, ['r'] * n, np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros((n, 4)), 1.5)
for reuse in (0, 4):
  ld = recordutil.WindowLoader(ds, batch_size=256, shuffle=True, reuse_buffers=reuse)
  it = iter(ld)
  for _ in range(20): next(it)
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0 = time.perf_counter(); a.record()
  k = 0
  for batch in it:
    x = batch[0]; k += 1
    if k == 1000: break
  b.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
  print('collate reuse=%d: %.2f us/batch (events) %.2f us/batch (host wall)' % (reuse, a.elapsed_time(b) * 1e3 / k, (t1 - t0) * 1e6 / k))
