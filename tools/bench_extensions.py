"""Timing of the optional stages (filter, resample, decode, noise gather, metrics) on a synthetic cohort — GPU box."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import numpy as np, torch
import bench, scgrhc
from scgrhc import ops, filters
from scipy import signal

n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device('cuda', 0)
arena = torch.empty((n_rec * bench.T_ROWS, 4), dtype=torch.float64, device=dev)
ops.synth_records(arena, bench.SEED, 0, n_rec, bench.T_ROWS, bench.KINDS, 16, bench.W)
rows = [bench.T_ROWS] * n_rec

def timed(fn, reps=3):
  fn(); torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  out = None
  for _ in range(reps):
    out = None                       # release the previous result first: the caching allocator then reuses its block
    out = fn()
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b) / reps, out

gb = arena.numel() * 8 / 1e9
sos = signal.butter(4, (1.0, 40.0), btype='bandpass', fs=500, output='sos')
ms, f = timed(lambda: filters.sosfiltfilt(arena, rows, sos, [0, 1, 2]))
print(json.dumps(dict(stage='sosfiltfilt order-4 bandpass (4 sections), 3 of 4 columns', records=n_rec, ms=ms, records_per_s=n_rec / ms * 1e3, input_gb=gb)))
ms2, _ = timed(lambda: filters.sosfiltfilt(arena, rows, sos, [0, 1, 2], exact=False))
print(json.dumps(dict(stage='sosfiltfilt time-parallel scan (CTA per record, warp per column), same filter', records=n_rec, ms=ms2, records_per_s=n_rec / ms2 * 1e3)))
ms, (r, rrows) = timed(lambda: filters.resample_poly(f, rows, 250, 500))
print(json.dumps(dict(stage='resample_poly 500->250 Hz, 4 columns', records=n_rec, ms=ms, records_per_s=n_rec / ms * 1e3, gbs=(gb * 1.5) / ms * 1e3)))
ms, _ = timed(lambda: filters.resample_poly(f, rows, 250, 500, exact=False))
print(json.dumps(dict(stage='resample_poly 500->250 Hz, 4 columns, fused multiply-add', records=n_rec, ms=ms, gbs=(gb * 1.5) / ms * 1e3)))
d = torch.clamp(torch.round(arena * torch.tensor([2e5, 2e5, 2e5, 500.0], device=dev, dtype=torch.float64)), -32767, 32767).to(torch.int16)
out = torch.empty_like(arena)
ms, _ = timed(lambda: ops.decode_fmt16(d, [0, 1, 2, 3], [2e5, 2e5, 2e5, 500.0], [0.0] * 4, out))
print(json.dumps(dict(stage='decode_fmt16 4 columns', records=n_rec, ms=ms, gbs=(gb * 1.25) / ms * 1e3)))
