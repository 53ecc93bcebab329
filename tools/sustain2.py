"""Burst -> sustained: per-quarter mean launch time, SM clock and board power of the window kernel, both arena layouts."""
import json, os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import torch, bench, scgrhc
from scgrhc import ops, _native as N

n_rec, iters = int(sys.argv[1]), int(sys.argv[2])
layouts = sys.argv[3].split(',') if len(sys.argv) > 3 else ['interleaved', 'planar']
dev = torch.device('cuda', 0)
plan = scgrhc.plan_uniform(bench.meta(), 'PA', bench.T_ROWS, bench.W, n_rec)
n, W, C = plan.n_cand, bench.W, 3
iv = plan.device_intervals(dev)
scg = torch.empty((n, C, W), dtype=torch.float32, device=dev); rhc = torch.empty((n, 1, W), dtype=torch.float32, device=dev)
minmax = torch.empty((n, 4), dtype=torch.float64, device=dev); keep = torch.empty(n, dtype=torch.uint8, device=dev)
reason = torch.empty(n, dtype=torch.uint8, device=dev); cw = torch.empty(n, dtype=torch.int32, device=dev); cr = torch.empty(n, dtype=torch.int32, device=dev)
for layout in layouts:
  planar = layout == 'planar'
  rows = n_rec * bench.T_ROWS
  if planar:
    arena = torch.empty((4, rows), dtype=torch.float64, device=dev)
    ops.synth_records(arena, bench.SEED, 0, n_rec, bench.T_ROWS, bench.KINDS, 16, bench.W, rows)
  else:
    arena = torch.empty((rows, 4), dtype=torch.float64, device=dev)
    ops.synth_records(arena, bench.SEED, 0, n_rec, bench.T_ROWS, bench.KINDS, 16, bench.W)
  flags = N.ARENA_PLANAR if planar else 0
  samples = []
  p = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu', '--format=csv,noheader,nounits', '-lms', '20', '-i', '0'], stdout=subprocess.PIPE, text=True)
  def rd():
    for line in p.stdout: samples.append((time.time(), line.strip()))
  threading.Thread(target=rd, daemon=True).start()
  def step():
    ops.process_windows(arena, iv, n, W, 0, [0, 1, 2], 3, -50.0, 1e-3, flags, [0.0] * 4, None, 0, scg, rhc, minmax, keep, reason, cw, cr)
  for _ in range(3): step()
  torch.cuda.synchronize()
  time.sleep(3.0)
  evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
  t0 = time.time()
  for a, b in evs:
    a.record(); step(); b.record()
  torch.cuda.synchronize()
  t1 = time.time()
  p.terminate()
  ms = [a.elapsed_time(b) for a, b in evs]
  sm = [s for t, s in samples if t0 <= t <= t1]
  q = iters // 8
  out = {'layout': layout, 'ms_by_eighth': [round(sum(ms[i * q:(i + 1) * q]) / q, 4) for i in range(8)], 'first10': round(sum(ms[:10]) / 10, 4)}
  k = max(1, len(sm) // 8)
  def col(j, part): 
    v = [float(x.split(',')[j]) for x in part]
    return round(sum(v) / len(v), 1) if v else None
  out['sm_mhz_by_eighth'] = [col(0, sm[i * k:(i + 1) * k]) for i in range(8)]
  out['watts_by_eighth'] = [col(2, sm[i * k:(i + 1) * k]) for i in range(8)]
  out['temp'] = [col(3, sm[:k]), col(3, sm[-k:])]
  print(json.dumps(out), flush=True)
  del arena
  torch.cuda.empty_cache()
  time.sleep(5.0)
