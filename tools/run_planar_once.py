"""ncu target: a few launches of the planar window kernel (C=3, W=750)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import torch, scgrhc
from scgrhc import ops, _native as N
dev = torch.device('cuda:0')
T, n_rec, C, W = 300000, int(sys.argv[1]) if len(sys.argv) > 1 else 500, 3, 750
planar = (sys.argv[2] != 'interleaved') if len(sys.argv) > 2 else True
meta = {'MacStTime': '1/1/2020 10:00:00', 'MacEndTime': '1/1/2020 10:10:00', 'ChamEvents_in_s': {'PA_1': 0}}
if planar:
  arena = torch.empty((4, n_rec * T), dtype=torch.float64, device=dev)
  ops.synth_records(arena, 0x5C6, 0, n_rec, T, [0, 1, 2, 3], 16, W, n_rec * T)
else:
  arena = torch.empty((n_rec * T, 4), dtype=torch.float64, device=dev)
  ops.synth_records(arena, 0x5C6, 0, n_rec, T, [0, 1, 2, 3], 16, W)
plan = scgrhc.plan_uniform(meta, 'PA', T, W, n_rec)
n = plan.n_cand
iv = plan.device_intervals(dev)
scg = torch.empty((n, C, W), dtype=torch.float32, device=dev); rhc = torch.empty((n, 1, W), dtype=torch.float32, device=dev)
mm = torch.empty((n, 4), dtype=torch.float64, device=dev); keep = torch.empty(n, dtype=torch.uint8, device=dev); reason = torch.empty_like(keep)
cw = torch.empty(n, dtype=torch.int32, device=dev); cr = torch.empty_like(cw)
for _ in range(4):
  ops.process_windows(arena, iv, n, W, 0, [0, 1, 2], 3, -50.0, 1e-3, N.ARENA_PLANAR if planar else 0, [0.0] * 4, None, 0, scg, rhc, mm, keep, reason, cw, cr)
torch.cuda.synchronize()
print('ok', int(keep.sum()))
