"""Builder-side probe: time recordutil.prepare_cohort over format-16 files on tmpfs, streamed vs eager."""
import os, sys, time, json, types, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import numpy as np, torch
import recordutil
from scgrhc import wfdbio, ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
T = 300000
SIG = ['patch_ACC_lat', 'patch_ACC_hf', 'patch_ACC_dv', 'RHC_pressure']
root = '/dev/shm/scgrhc_probe'
shutil.rmtree(root, ignore_errors=True); os.makedirs(root)
dev = torch.device('cuda:0')
arena = torch.empty((n * T, 4), dtype=torch.float64, device=dev)
ops.synth_records(arena, 0x5C6, 0, n, T, [0, 1, 2, 3], 16, 750)
g = torch.tensor([2e5, 2e5, 2e5, 500.0], dtype=torch.float64, device=dev)
frames = torch.clamp(torch.round(arena * g), -32767, 32767).to(torch.int16).cpu().numpy()
del arena
meta = json.dumps({'MacStTime': '1/1/2020 10:00:00', 'MacEndTime': '1/1/2020 10:10:00', 'ChamEvents_in_s': {'PA_1': 0}})
for r in range(n):
  name = 'rec%05d' % r
  frames[r * T:(r + 1) * T].tofile(os.path.join(root, name + '.dat'))
  with open(os.path.join(root, name + '.hea'), 'w') as f:
    f.write('%s 4 500 %d\n' % (name, T))
    for k, gg in enumerate([2e5, 2e5, 2e5, 500.0]):
      f.write('%s.dat 16 %.17g(0)/g 16 0 0 0 0 %s\n' % (name, gg, SIG[k]))
  open(os.path.join(root, name + '.json'), 'w').write(meta)
recordutil.PROCESSED_DATA_PATH, recordutil.wfdb = root, wfdbio
params = types.SimpleNamespace(in_channels=SIG[:3], chamber='PA', segment_size=1.5, min_RHC=-50.0, use_global_min_max=False)
names = recordutil.get_record_names()
for mode in ('streamed', 'eager'):
  for chunk in (16, 32, 64):
    ts = []
    for rep in range(4):
      torch.cuda.synchronize(); t0 = time.perf_counter()
      if mode == 'streamed':
        st = recordutil._prepare_streamed(params, names, 0, 3, dev, chunk, None)
      else:
        st = recordutil._prepare_eager(params, names, names, 0, 3, dev, chunk, None)
      k = st.kept_idx.cpu(); torch.cuda.synchronize()
      ts.append(time.perf_counter() - t0)
      nk = st.n_kept
      del st
    print(mode, chunk, 'ms', [round(t * 1e3, 1) for t in ts], 'M windows/s', round(nk / min(ts[1:]) / 1e6, 2), flush=True)
shutil.rmtree(root, ignore_errors=True)
