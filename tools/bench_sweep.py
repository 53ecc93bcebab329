"""BASELINE configs[4]: all runnable waveform_NN configs as one sweep job over a device-resident 5-signal cohort — GPU box."""
import json, os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import torch
import scgrhc
from scgrhc import ops, sweep
from oracle import synth_ref                      # side-car layout + signal names only (tool, not product)
from tests import helpers as H

n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 200
T = 300000
sig = synth_ref.SIG_NAMES_5
dev = torch.device('cuda', 0)
arena = torch.empty((n_rec * T, 5), dtype=torch.float64, device=dev)
ops.synth_records(arena, 0x5C6, 0, n_rec, T, list(synth_ref.kinds_for(sig)), 16, 750)
table = H.configs()
configs = {c: types.SimpleNamespace(**H.effective_config(c, table)) for c in table if c != 'waveform_01'}
metas = [synth_ref.record_meta(600)] * n_rec
bufs = {}
for rep in range(2):
  torch.cuda.synchronize(); t0 = time.time()
  kept = cand = 0
  for name, st in sweep.iter_sweep(arena, sig, metas, [T] * n_rec, configs, buffers=bufs):
    kept += st.n_kept; cand += st.n_cand
  torch.cuda.synchronize(); dt = time.time() - t0
print(json.dumps(dict(configs=len(configs), records=n_rec, seconds=dt, candidate_windows=cand, kept_windows=kept,
                      kept_windows_per_s=kept / dt, configs_per_s=len(configs) / dt)))
