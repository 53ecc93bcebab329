import json, os, sys, time, types
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import torch, scgrhc
from scgrhc import ops, sweep, engine
from oracle import synth_ref
from tests import helpers as H
n_rec = int(sys.argv[1]); T=300000
sig = synth_ref.SIG_NAMES_5
dev = torch.device('cuda', 0)
arena = torch.empty((n_rec * T, 5), dtype=torch.float64, device=dev)
ops.synth_records(arena, 0x5C6, 0, n_rec, T, list(synth_ref.kinds_for(sig)), 16, 750)
metas = [synth_ref.record_meta(600)] * n_rec
def sync(): torch.cuda.synchronize(); return time.time()
for rep in range(3):
  t0=sync()
  plan = engine.plan_cohort(metas, 'PA', [T]*n_rec, 750)
  t1=sync()
  pred = engine.prepare_windows(arena, plan, [0], 3, -50.0, predicates_only=True)
  t2=sync()
  subsets=[[0,1,2],[0,1],[0,2],[1,2],[0],[1],[2],[0,1,2,4]]
  sts = engine.prepare_subsets(arena, plan, [0,1,2,4], 3, -50.0, subsets)
  t3=sync()
  one = engine.prepare_windows(arena, plan, [0,1,2], 3, -50.0)
  t4=sync()
  a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
  n_kept=pred.n_kept
  rhc=sts[0].rhc; scgs=[s.scg for s in sts]; mms=[s.minmax for s in sts]
  a.record()
  ops.normalize_subsets(arena, plan.device_intervals(dev), 750, 0, [0,1,2,4], 3, pred.kept_idx, n_kept, [0,1,2,0,1,0,2,1,2,0,1,2,0,1,2,3],[3,2,2,2,1,1,1,4], False, scgs, mms, rhc)
  b.record(); torch.cuda.synchronize()
  print(dict(plan_ms=(t1-t0)*1e3, pred_ms=(t2-t1)*1e3, subsets_ms=(t3-t2)*1e3, one_config_ms=(t4-t3)*1e3, fanout_kernel_ms=a.elapsed_time(b), n_cand=plan.n_cand, n_kept=n_kept))
  del sts, scgs, mms, rhc, one
