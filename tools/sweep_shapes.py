"""Window kernel throughput for the arena shapes the drop-in produces: C SCG channels + RHC, identity columns — GPU box."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import torch, bench, scgrhc
from scgrhc import ops

n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 500
dev = torch.device('cuda', 0)
peak = bench.measured_peak()[0]
for C, kinds in ((1, [0, 3]), (2, [0, 1, 3]), (3, [0, 1, 2, 3]), (4, [0, 1, 2, 4, 3])):
  nsig = C + 1
  arena = torch.empty((n_rec * bench.T_ROWS, nsig), dtype=torch.float64, device=dev)
  ops.synth_records(arena, bench.SEED, 0, n_rec, bench.T_ROWS, kinds, 16, bench.W)
  plan = scgrhc.plan_uniform(bench.meta(), 'PA', bench.T_ROWS, bench.W, n_rec)
  n, W = plan.n_cand, bench.W
  iv = plan.device_intervals(dev)
  scg = torch.empty((n, C, W), dtype=torch.float32, device=dev); rhc = torch.empty((n, 1, W), dtype=torch.float32, device=dev)
  mm = torch.empty((n, 4), dtype=torch.float64, device=dev); keep = torch.empty(n, dtype=torch.uint8, device=dev)
  reason = torch.empty(n, dtype=torch.uint8, device=dev); cw = torch.empty(n, dtype=torch.int32, device=dev); cr = torch.empty(n, dtype=torch.int32, device=dev)
  def step():
    ops.process_windows(arena, iv, n, W, 0, list(range(C)), C, -50.0, 1e-3, 0, [0.0] * 4, None, 0, scg, rhc, mm, keep, reason, cw, cr)
  for _ in range(3): step()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(10): step()
  b.record(); torch.cuda.synchronize()
  ms = a.elapsed_time(b) / 10
  nk = int(keep.sum())
  alg = n * (W * 8 + 1) + nk * (W * C * 8 + W * (C + 1) * 4 + 52)
  print(json.dumps(dict(C=C, nsig=nsig, ms=ms, mcand_s=n / ms / 1e3, alg_gbs=alg / ms / 1e6, frac=alg / ms / 1e6 / peak)), flush=True)
  del arena, scg, rhc
  torch.cuda.empty_cache()
