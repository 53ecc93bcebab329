"""Two (or more) ranks of the sharded drop-in on real GPUs: launched by tests/test_gpu_multi.py under torch.distributed.run.

Backend: NCCL with one GPU per rank when the box has at least as many GPUs as ranks; otherwise every rank drives cuda:0
and the (tiny) collectives run over gloo — the CUDA path, the sharding and the ordering logic are the same.
Writes what each rank produced to ``out_dir`` for the single-process test to compare with its own 1-rank run."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))

data_root, out_dir = sys.argv[1], sys.argv[2]
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
multi = torch.cuda.device_count() >= world
torch.cuda.set_device(local if multi else 0)
dev = torch.device('cuda', torch.cuda.current_device())
if multi:
  dist.init_process_group('nccl', device_id=dev)
else:
  dist.init_process_group('gloo')

import recordutil  # noqa: E402
import scgrhc  # noqa: E402
from paramutil import Params  # noqa: E402
from scgrhc import wfdbio  # noqa: E402
from tests import helpers as H  # noqa: E402

recordutil.PROCESSED_DATA_PATH = data_root
recordutil.wfdb = wfdbio

with open(os.path.join(out_dir, 'jobs.json')) as f:
  jobs = json.load(f)

for tag, pdir in jobs['prepare'].items():               # ---- prepare_cohort: this rank's block of the ordered list
  params = Params(os.path.join(pdir, 'params.json'))
  store, names = recordutil.prepare_cohort(params, chunk_records=2)
  scg, rhc = store.materialise()
  np.savez(os.path.join(out_dir, 'prep_%s_rank%d.npz' % (tag, rank)), scg=scg.cpu().numpy(), rhc=rhc.cpu().numpy(),
           rec_id=store.rec_id.cpu().numpy(), start=store.start_idx.cpu().numpy(), mm=store.kept_minmax().cpu().numpy(),
           offset=store.shard.offset, total=store.shard.total, counts=np.array(store.shard.counts),
           gmm=store.global_minmax.cpu().numpy() if store.global_minmax is not None else np.zeros(0))

for pdir in jobs['save']:                               # ---- save_dataloaders: rank 0 writes the reference's pickles
  recordutil.save_dataloaders(Params(os.path.join(pdir, 'params.json')))

counts = recordutil.prepare_all(jobs['sweep'])          # ---- the sweep, records sharded
if rank == 0:
  with open(os.path.join(out_dir, 'sweep_counts.json'), 'w') as f:
    json.dump(counts, f)

# ---- dataset-level min/max across ranks against the fixture of the unmodified reference (waveform_04 on two 10-min
#      records, tests/golden/records_full.json): rank r owns record r
if world == 2:
  c = H.effective_config('waveform_04')
  sig, p, meta = H.full_record(rank)
  cols, rcol = scgrhc.resolve_columns(sig, c['in_channels'])
  plan = scgrhc.plan_cohort([meta], c['chamber'], [p.shape[0]], 750, rec0=rank)
  st = scgrhc.prepare_windows(torch.from_numpy(np.ascontiguousarray(p)).to(dev), plan, cols, rcol, c['min_RHC'], use_global_min_max=True)
  with open(os.path.join(out_dir, 'golden04_rank%d.json' % rank), 'w') as f:
    json.dump({'gmm_hex': [float(v).hex() for v in st.global_minmax.cpu().numpy()], 'start': st.start_idx.cpu().tolist(),
               'scg_sha': H.sha(st.scg[:st.n_kept].cpu().numpy()), 'rhc_sha': H.sha(st.rhc[:st.n_kept].cpu().numpy())}, f)

dist.barrier()
if rank == 0:
  print('MULTI_OK backend=%s' % dist.get_backend())
dist.destroy_process_group()
