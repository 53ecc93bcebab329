"""Sustained vs burst HBM copy bandwidth (the roofline denominator is a burst copy) — GPU box."""
import subprocess, sys, threading, time
import torch
n = 1 << 30                                        # 1 Gi bf16 elements = 2 GiB per buffer, as MEASURED_PEAKS.json
a = torch.empty(n, dtype=torch.bfloat16, device='cuda'); b = torch.empty_like(a)
a.normal_()
samples = []
p = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap', '--format=csv,noheader,nounits', '-lms', '50', '-i', '0'], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [samples.append((time.time(), l.strip())) for l in p.stdout], daemon=True).start()
for _ in range(3): b.copy_(a)
torch.cuda.synchronize(); time.sleep(0.5)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 600
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
t0 = time.time()
for x, y in evs:
  x.record(); b.copy_(a); y.record()
torch.cuda.synchronize(); t1 = time.time()
p.terminate()
ms = [x.elapsed_time(y) for x, y in evs]
gbs = [2 * n * 2 / m / 1e6 for m in ms]
print('burst best of first 10: %.0f GB/s; mean of first 10: %.0f; mean of last half: %.0f; min over run: %.0f' %
      (max(gbs[:10]), sum(gbs[:10]) / 10, sum(gbs[iters // 2:]) / (iters - iters // 2), min(gbs)))
print('clocks under load:', [s for t, s in samples if t0 <= t <= t1][::6][:12])
