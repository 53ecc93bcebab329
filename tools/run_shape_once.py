"""ncu target: the window kernel for an arena of C SCG columns + RHC (the shapes the drop-in uploads): C W [planar]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import torch, scgrhc
from scgrhc import ops, _native as N
dev = torch.device('cuda:0')
C, W = int(sys.argv[1]), int(sys.argv[2])
planar = len(sys.argv) > 3 and sys.argv[3] == 'planar'
n_rec, T = 500, 300000 * W // 750
nsig = C + 1
kinds = [0, 1, 2, 4][:C] + [3]
meta = {'MacStTime': '1/1/2020 10:00:00', 'MacEndTime': '1/1/2020 10:10:00', 'ChamEvents_in_s': {'PA_1': 0}}
if planar:
  arena = torch.empty((nsig, n_rec * T), dtype=torch.float64, device=dev)
  ops.synth_records(arena, 0x5C6, 0, n_rec, T, kinds, 16, W, n_rec * T)
else:
  arena = torch.empty((n_rec * T, nsig), dtype=torch.float64, device=dev)
  ops.synth_records(arena, 0x5C6, 0, n_rec, T, kinds, 16, W)
plan = scgrhc.Plan(scgrhc.plan_uniform(meta, 'PA', 300000, 750, n_rec).intervals.copy(), 0, W)
iv_np = plan.intervals
iv_np['row0'] = (iv_np['row0'] // 300000) * T
n = int(iv_np['n_win'].sum())
plan.n_cand = n
iv = plan.device_intervals(dev)
scg = torch.empty((n, C, W), dtype=torch.float32, device=dev); rhc = torch.empty((n, 1, W), dtype=torch.float32, device=dev)
mm = torch.empty((n, 4), dtype=torch.float64, device=dev); keep = torch.empty(n, dtype=torch.uint8, device=dev); reason = torch.empty_like(keep)
cw = torch.empty(n, dtype=torch.int32, device=dev); cr = torch.empty_like(cw)
for _ in range(4):
  ops.process_windows(arena, iv, n, W, 0, list(range(C)), C, -50.0, 1e-3, N.ARENA_PLANAR if planar else 0, [0.0] * 4, None, 0, scg, rhc, mm, keep, reason, cw, cr)
torch.cuda.synchronize()
print('ok', n, int(keep.sum()))
