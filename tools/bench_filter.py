"""Tuning sweep of the time-parallel sosfiltfilt (chunk rows per lane x staging buffers) on a synthetic cohort — GPU box.
usage: python tools/bench_filter.py [n_rec] [chunk:nbuf ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import numpy as np, torch
import bench
from scgrhc import ops, filters
from scipy import signal

n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
shapes = [tuple(int(v) for v in a.split(":")) for a in sys.argv[2:]] or [(0, 0), (24, 1), (16, 1), (12, 2), (20, 1), (28, 1)]
dev = torch.device('cuda', 0)
arena = torch.empty((n_rec * bench.T_ROWS, 4), dtype=torch.float64, device=dev)
ops.synth_records(arena, bench.SEED, 0, n_rec, bench.T_ROWS, bench.KINDS, 16, bench.W)
rows = [bench.T_ROWS] * n_rec
out = torch.empty_like(arena)
sos = signal.butter(4, (1.0, 40.0), btype='bandpass', fs=500, output='sos')
gb = arena.numel() * 8 / 1e9

def timed(fn, reps=5):
  fn(); torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps): fn()
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b) / reps

import ctypes as C
from scgrhc import _native as N
zi = np.ascontiguousarray(filters.sosfilt_zi(sos)); edge = filters.pad_length(sos)
row0 = np.arange(n_rec + 1, dtype=np.int64) * bench.T_ROWS
row0_dev = torch.from_numpy(row0).to(dev)
c = ops.ctx(0)
def call(ncf, chunk, nbuf, s=sos):
  cols = (C.c_int32 * ncf)(*range(ncf))
  z = np.ascontiguousarray(filters.sosfilt_zi(s))
  N.check(c, N.lib().scgrhc_sosfiltfilt_scan(c, ops._ptr(arena), ops._ptr(out), ops._ptr(row0_dev), row0.ctypes.data_as(C.POINTER(C.c_int64)),
                                             n_rec, 4, cols, ncf, s.ctypes.data_as(C.POINTER(C.c_double)), z.ctypes.data_as(C.POINTER(C.c_double)),
                                             s.shape[0], filters.pad_length(s), chunk, nbuf, ops._stream(0)))
for chunk, nbuf in shapes:
  ms = timed(lambda: call(3, chunk, nbuf))
  print(json.dumps(dict(stage='sosfiltfilt_scan 4 sections, 3 of 4 columns', records=n_rec, chunk=chunk, nbuf=nbuf, ms=round(ms, 3),
                        hbm_gbs=round(4 * gb / ms * 1e3, 1), frac_of_peak=round(4 * gb / ms * 1e3 / 6547.2, 3))), flush=True)
sos2 = signal.butter(2, (1.0, 40.0), btype='bandpass', fs=500, output='sos')
for ncf, s, name in ((1, sos, '4 sections, 1 column'), (4, sos, '4 sections, 4 columns'), (3, sos2, '2 sections, 3 columns')):
  ms = timed(lambda: call(ncf, 0, 0, s))
  print(json.dumps(dict(stage='sosfiltfilt_scan ' + name, records=n_rec, ms=round(ms, 3), hbm_gbs=round(4 * gb / ms * 1e3, 1))), flush=True)
ms = timed(lambda: torch.cuda.synchronize() or out.copy_(arena))
print(json.dumps(dict(stage='copy arena (roofline reference: 2 of the 4 crossings)', ms=round(ms, 3), gbs=round(2 * gb / ms * 1e3, 1))))
