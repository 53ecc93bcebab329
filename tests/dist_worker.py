"""world_size-2 gloo worker: record sharding + the {min,-max} MIN all-reduce used for use_global_min_max."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import scgrhc  # noqa: E402
from oracle import scgrhc_oracle as orc, synth_ref  # noqa: E402

dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
n_rec = 7
lo, hi = scgrhc.shard_records(n_rec, rank, world)
spans = [scgrhc.shard_records(n_rec, r, world) for r in range(world)]
assert spans[0][0] == 0 and spans[-1][1] == n_rec and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))

# per-shard statistics from the oracle (the device reduction is tested on the GPU); the collective is the product's
sig = synth_ref.DEFAULT_SIG_NAMES
meta = synth_ref.record_meta(60, events={'PA_1': 0})
rows = []
for r in range(n_rec):
  p = synth_ref.gen_record(77, r, 30000, kinds=synth_ref.kinds_for(sig))
  rw = orc.scan_record(p, sig, meta, sig[:3], 'PA', 1.5, -50.0)
  rows.append(rw.minmax[rw.keep])
want = orc.global_minmax(np.concatenate(rows))
mine = rows[lo:hi]
local = torch.from_numpy(orc.global_minmax(np.concatenate(mine))) if len(mine) else \
    torch.tensor([np.inf, -np.inf, np.inf, -np.inf], dtype=torch.float64)
got = scgrhc.allreduce_minmax(local.clone())
assert got.numpy().tobytes() == want.tobytes(), (got, want)

# the plan of a shard is the global plan restricted to its records
full = scgrhc.plan_uniform(meta, 'PA', 30000, 750, n_rec)
part = scgrhc.plan_uniform(meta, 'PA', 30000, 750, hi - lo, rec0=lo)
m = (full.intervals['rec_id'] >= lo) & (full.intervals['rec_id'] < hi)
assert (part.intervals['rec_id'] == full.intervals['rec_id'][m]).all()
assert (part.intervals['n_win'] == full.intervals['n_win'][m]).all()
counts = torch.tensor([part.n_cand], dtype=torch.int64)
dist.all_reduce(counts)
assert int(counts) == full.n_cand
# the plumbing of the sharded drop-in (scgrhc/dist.py): kept-count exchange, one split draw for the job, rows to rank 0
from scgrhc import dist as sdist  # noqa: E402
n_local = 5 + 3 * rank
sh = sdist.exchange_counts(n_local, torch.device('cpu'))
assert sh.world == world and sh.counts == tuple(5 + 3 * r for r in range(world)) and sh.offset == sum(sh.counts[:rank])
assert sh.total == sum(sh.counts)
idx = sdist.broadcast_index(torch.arange(7) * (rank + 1), torch.device('cpu'))
assert idx.tolist() == list(range(7))
rows_t = (torch.arange(n_local * 3, dtype=torch.float32).reshape(n_local, 3) + 1000 * rank)
got = sdist.gather_rows(rows_t, list(sh.counts))
if rank == 0:
  want_rows = torch.cat([torch.arange((5 + 3 * r) * 3, dtype=torch.float32).reshape(-1, 3) + 1000 * r for r in range(world)])
  assert torch.equal(got, want_rows)
else:
  assert got is None
empty = sdist.gather_rows(torch.zeros((0, 2)), [0] * world)           # nothing to send anywhere is fine
assert empty is None or empty.shape == (0, 2)
dist.barrier()
if rank == 0:
  print('DIST_OK')
dist.destroy_process_group()
