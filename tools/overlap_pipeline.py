"""Does running the band-pass of chunk k+1 beside the decimating window kernel of chunk k (two streams) beat running the
two stages back to back?  Both are bound by dependent fp64 latency at ~50 % pipe utilisation when alone."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import numpy as np, torch
import bench, scgrhc
from scgrhc import ops, filters
from scipy import signal as _sig

dev = torch.device('cuda', 0)
n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
T, SIG = bench.T_ROWS, bench.SIG
cols, rcol = scgrhc.resolve_columns(SIG, bench.IN_CHANNELS)
arena = torch.empty((n_rec * T, len(SIG)), dtype=torch.float64, device=dev)
ops.synth_records(arena, bench.SEED, 0, n_rec, T, bench.KINDS, 16, bench.W)
sos = _sig.butter(4, (1.0, 40.0), btype='bandpass', fs=500, output='sos')
fs2, W2 = 250, 375

def chunks(K):
  b = [n_rec * i // K for i in range(K + 1)]
  out = []
  for i in range(K):
    n = b[i + 1] - b[i]
    out.append(dict(n=n, x=arena[b[i] * T:b[i + 1] * T], rows=[T] * n, plan=scgrhc.plan_uniform(bench.meta(), 'PA', T // 2, W2, n, rec0=b[i]),
                    spec=filters.DecimSpec.design([T] * n, fs2, 500), bufs={}, f=torch.empty((n * T, len(SIG)), dtype=torch.float64, device=dev)))
  return out

def bp(c):
  c['f'] = filters.sosfiltfilt(c['x'], c['rows'], sos, cols, exact=False)
def win(c):
  return scgrhc.prepare_windows(c['f'], c['plan'], cols, rcol, bench.MIN_RHC, normalisation='zscore', buffers=c['bufs'], check=False, decim=c['spec'])

def run(K, overlap, ctas=0):
  cs = chunks(K)
  sA, sB = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
  def once():
    kept = 0
    if not overlap:
      for c in cs:
        bp(c); win(c)
      return
    evs = []
    for i, c in enumerate(cs):
      with torch.cuda.stream(sA):
        bp(c)
        e = torch.cuda.Event(); e.record(sA); evs.append(e)
      with torch.cuda.stream(sB):
        sB.wait_event(evs[i])
        win(c)
  main = torch.cuda.current_stream(dev)
  for _ in range(2):
    sA.wait_stream(main); sB.wait_stream(main); once(); main.wait_stream(sA); main.wait_stream(sB)
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  reps = 3
  a.record()
  for _ in range(reps):
    sA.wait_stream(main); sB.wait_stream(main); once(); main.wait_stream(sA); main.wait_stream(sB)
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b) / reps

for K in (1, 2, 4, 8):
  for overlap in (False, True):
    if K == 1 and overlap: continue
    print(json.dumps(dict(chunks=K, overlap=overlap, ms=round(run(K, overlap), 3))), flush=True)
