"""ncu target: the decimating window kernel (500 -> 250 Hz inside the window kernel, 375-sample windows, z-score)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import torch, scgrhc
from scgrhc import ops, filters
dev = torch.device('cuda:0')
T, n_rec = 300000, int(sys.argv[1]) if len(sys.argv) > 1 else 300
meta = {'MacStTime': '1/1/2020 10:00:00', 'MacEndTime': '1/1/2020 10:10:00', 'ChamEvents_in_s': {'PA_1': 0}}
arena = torch.empty((n_rec * T, 4), dtype=torch.float64, device=dev)
ops.synth_records(arena, 0x5C6, 0, n_rec, T, [0, 1, 2, 3], 16, 750)
spec = filters.DecimSpec.design([T] * n_rec, 250, 500)
plan = scgrhc.plan_uniform(meta, 'PA', T // 2, 375, n_rec)
bufs = {}
for _ in range(3):
  st = scgrhc.prepare_windows(arena, plan, [0, 1, 2], 3, -50.0, normalisation='zscore', buffers=bufs, check=False, decim=spec)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
  st = scgrhc.prepare_windows(arena, plan, [0, 1, 2], 3, -50.0, normalisation='zscore', buffers=bufs, check=False, decim=spec)
b.record(); torch.cuda.synchronize()
print('ok', st.n_kept, 'ms per call', a.elapsed_time(b) / 5)
