"""Does the window kernel slow down under sustained load?  Per-iteration times over a long loop + nvidia-smi clocks."""
import json, os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import torch, bench, scgrhc
from scgrhc import ops

n_rec, iters = int(sys.argv[1]), int(sys.argv[2])
ctas, stages = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, 0)
dev = torch.device('cuda', 0)
ops.set_tuning(0, ctas, stages)
arena = torch.empty((n_rec * bench.T_ROWS, 4), dtype=torch.float64, device=dev)
ops.synth_records(arena, bench.SEED, 0, n_rec, bench.T_ROWS, bench.KINDS, 16, bench.W)
plan = scgrhc.plan_uniform(bench.meta(), 'PA', bench.T_ROWS, bench.W, n_rec)
n, W, C = plan.n_cand, bench.W, 3
iv = plan.device_intervals(dev)
scg = torch.empty((n, C, W), dtype=torch.float32, device=dev); rhc = torch.empty((n, 1, W), dtype=torch.float32, device=dev)
minmax = torch.empty((n, 4), dtype=torch.float64, device=dev); keep = torch.empty(n, dtype=torch.uint8, device=dev)
reason = torch.empty(n, dtype=torch.uint8, device=dev); cw = torch.empty(n, dtype=torch.int32, device=dev); cr = torch.empty(n, dtype=torch.int32, device=dev)
samples = []
p = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,temperature.gpu', '--format=csv,noheader,nounits', '-lms', '20', '-i', '0'], stdout=subprocess.PIPE, text=True)
def rd():
  for line in p.stdout: samples.append((time.time(), line.strip()))
threading.Thread(target=rd, daemon=True).start()
def step():
  ops.process_windows(arena, iv, n, W, 0, [0, 1, 2], 3, -50.0, 1e-3, 0, [0.0] * 4, None, 0, scg, rhc, minmax, keep, reason, cw, cr)
for _ in range(3): step()
torch.cuda.synchronize()
time.sleep(0.5)
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
t0 = time.time()
for a, b in evs:
  a.record(); step(); b.record()
torch.cuda.synchronize()
t1 = time.time()
p.terminate()
ms = [a.elapsed_time(b) for a, b in evs]
print('ctas %d stages %d: first10 %.3f ms  last50%% mean %.3f ms' % (ctas, stages, sum(ms[:10]) / 10, sum(ms[iters // 2:]) / (iters - iters // 2)))
print('under load (last):', [s for t, s in samples if t0 <= t <= t1][-2:])
