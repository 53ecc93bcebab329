"""Builder-side probe: planar vs interleaved window kernel (burst + sustained), several shapes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import numpy as np, torch
import scgrhc
from scgrhc import ops, _native as N

dev = torch.device('cuda:0')
T, n_rec = 300000, int(sys.argv[1]) if len(sys.argv) > 1 else 1000
PEAK = 6547.2
meta = {'MacStTime': '1/1/2020 10:00:00', 'MacEndTime': '1/1/2020 10:10:00', 'ChamEvents_in_s': {'PA_1': 0}}

def run(C, W, planar, sustained=False, T_rows=T):
  nsig = C + 1
  kinds = [0, 1, 2, 4][:C] + [3]
  if planar:
    arena = torch.empty((nsig, n_rec * T_rows), dtype=torch.float64, device=dev)
    ops.synth_records(arena, 0x5C6, 0, n_rec, T_rows, kinds, 16, W, n_rec * T_rows)
  else:
    arena = torch.empty((n_rec * T_rows, nsig), dtype=torch.float64, device=dev)
    ops.synth_records(arena, 0x5C6, 0, n_rec, T_rows, kinds, 16, W)
  plan = scgrhc.plan_uniform(meta, 'PA', T_rows, W, n_rec)
  n = plan.n_cand
  iv = plan.device_intervals(dev)
  scg = torch.empty((n, C, W), dtype=torch.float32, device=dev); rhc = torch.empty((n, 1, W), dtype=torch.float32, device=dev)
  mm = torch.empty((n, 4), dtype=torch.float64, device=dev); keep = torch.empty(n, dtype=torch.uint8, device=dev); reason = torch.empty_like(keep)
  cw = torch.empty(n, dtype=torch.int32, device=dev); cr = torch.empty_like(cw)
  flags = N.ARENA_PLANAR if planar else 0
  def step():
    ops.process_windows(arena, iv, n, W, 0, list(range(C)), C, -50.0, 1e-3, flags, [0.0] * 4, None, 0, scg, rhc, mm, keep, reason, cw, cr)
  for _ in range(3): step()
  torch.cuda.synchronize()
  reps = 10
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps): step()
  b.record(); torch.cuda.synchronize()
  ms = a.elapsed_time(b) / reps
  nk = int(keep.sum())
  alg = n * (W * 8 + 1) + nk * (W * C * 8 + W * (C + 1) * 4 + 52)
  out = dict(C=C, W=W, planar=planar, ms=round(ms, 4), frac=round(alg / ms / 1e6 / PEAK, 4), kept=nk, cand=n)
  if sustained:
    reps = int(2500 / ms)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps // 2)]
    for k in range(reps):
      if k >= reps - len(ev): ev[k - (reps - len(ev))][0].record()
      step()
      if k >= reps - len(ev): ev[k - (reps - len(ev))][1].record()
    torch.cuda.synchronize()
    ms2 = sum(x.elapsed_time(y) for x, y in ev) / len(ev)
    out.update(sustained_ms=round(ms2, 4), sustained_frac=round(alg / ms2 / 1e6 / PEAK, 4))
    time.sleep(3)
  return out

import json
for C, W, Tr in ((3, 750, T), (1, 750, T), (2, 750, T), (4, 750, T), (3, 375, T // 2)):
  for planar in (False, True):
    print(json.dumps(run(C, W, planar, sustained=(C == 3 and W == 750), T_rows=Tr)), flush=True)
    torch.cuda.empty_cache()
