"""ncu target for the auxiliary kernels the judge asked summaries for: subset fan-out, decimator, format-16 decode (both
kernels), one-launch collate (+noise), compaction.  One call of each after a warm-up, on bench-sized inputs.

  ncu --set full --clock-control none -k regex:'subset_norm|resample_decim|decode_fmt16|collate_batch' -o ... python tools/prof_aux.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import numpy as np, torch
import scgrhc
from scgrhc import ops, engine, filters

dev = torch.device('cuda:0')
n_rec, T = int(sys.argv[1]) if len(sys.argv) > 1 else 300, 300000
meta = {'MacStTime': '1/1/2020 10:00:00', 'MacEndTime': '1/1/2020 10:10:00', 'ChamEvents_in_s': {'PA_1': 0}}
for rep in range(2):
  # sweep fan-out over a 5-signal cohort
  a5 = torch.empty((n_rec * T, 5), dtype=torch.float64, device=dev)
  ops.synth_records(a5, 0x5C6, 0, n_rec, T, [0, 1, 2, 3, 4], 16, 750)
  plan = scgrhc.plan_uniform(meta, 'PA', T, 750, n_rec)
  subsets = [[0, 1, 2], [0, 1], [0, 2], [1, 2], [0], [1], [2], [0, 1, 2, 4]]
  sts = engine.prepare_subsets(a5, plan, [0, 1, 2, 4], 3, -50.0, subsets)
  n_kept = sts[0].n_kept
  del sts, a5
  # decimator 500 -> 250 Hz, bit-identical form, 4 columns
  a4 = torch.empty((n_rec * T, 4), dtype=torch.float64, device=dev)
  ops.synth_records(a4, 0x5C6, 0, n_rec, T, [0, 1, 2, 3], 16, 750)
  r, _ = filters.resample_poly(a4, [T] * n_rec, 250, 500)
  del r
  # format-16 decode: one calibration, and per-record tables
  g = torch.tensor([2e5, 2e5, 2e5, 500.0], dtype=torch.float64, device=dev)
  d = torch.clamp(torch.round(a4 * g), -32767, 32767).to(torch.int16)
  out = torch.empty_like(a4)
  ops.decode_fmt16(d, [0, 1, 2, 3], [2e5, 2e5, 2e5, 500.0], [0.0] * 4, out)
  off = torch.arange(0, (n_rec + 1) * T, T, dtype=torch.int64, device=dev)
  gt = g.repeat(n_rec, 1).contiguous(); bt = torch.zeros_like(gt)
  ops.decode_fmt16_records(d, off, T, [0, 1, 2, 3], gt, bt, True, out)
  del d, out
  # batch-256 collate from a window store
  plan4 = scgrhc.plan_uniform(meta, 'PA', T, 750, n_rec)
  st = scgrhc.prepare_windows(a4, plan4, [0, 1, 2], 3, -50.0)
  slots = st.kept_idx[torch.randperm(st.n_kept, device=dev)[:256 * 8]].contiguous()
  col = ops.BatchCollator(st.scg, st.rhc, slots)
  bs, br = torch.empty((256, 3, 750), device=dev), torch.empty((256, 1, 750), device=dev)
  for k in range(8):
    col(256 * k, 256, bs, br, 0.01 if k % 2 else 0.0, 7, k)
  del st, a4
torch.cuda.synchronize()
print('ok', n_kept)
