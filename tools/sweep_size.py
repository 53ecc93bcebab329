"""Kernel-only timing of the window kernel vs cohort size (footprint effects) — run on the GPU box."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import torch  # noqa: E402
import bench  # noqa: E402
import scgrhc  # noqa: E402
from scgrhc import ops  # noqa: E402


def run(n_rec, iters=10, flags=0, quantised=False):
  dev = torch.device('cuda', 0)
  arena = torch.empty((n_rec * bench.T_ROWS, 4), dtype=torch.float64, device=dev)
  ops.synth_records(arena, bench.SEED, 0, n_rec, bench.T_ROWS, bench.KINDS, 16, bench.W)
  if quantised:   # what a format-16 record decodes to: RHC on a 2e-3 mmHg grid (coarser than the 1e-3 flat-line threshold)
    g = torch.tensor([2e5, 2e5, 2e5, 500.0], dtype=torch.float64, device=dev)
    for r0 in range(0, arena.shape[0], 50 * bench.T_ROWS):
      arena[r0:r0 + 50 * bench.T_ROWS] = torch.round(arena[r0:r0 + 50 * bench.T_ROWS] * g) / g
  plan = scgrhc.plan_uniform(bench.meta(), 'PA', bench.T_ROWS, bench.W, n_rec)
  n, W, C = plan.n_cand, bench.W, 3
  iv = plan.device_intervals(dev)
  scg = torch.empty((n, C, W), dtype=torch.float32, device=dev)
  rhc = torch.empty((n, 1, W), dtype=torch.float32, device=dev)
  minmax = torch.empty((n, 4), dtype=torch.float64, device=dev)
  keep = torch.empty(n, dtype=torch.uint8, device=dev)
  reason = torch.empty(n, dtype=torch.uint8, device=dev)
  cw = torch.empty(n, dtype=torch.int32, device=dev)
  cr = torch.empty(n, dtype=torch.int32, device=dev)

  def step():
    ops.process_windows(arena, iv, n, W, 0, [0, 1, 2], 3, -50.0, 1e-3, flags, [0.0] * 4, None, 0, scg, rhc, minmax, keep, reason, cw, cr)
  for _ in range(3):
    step()
  torch.cuda.synchronize()
  evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
  for a, b in evs:
    a.record(); step(); b.record()
  torch.cuda.synchronize()
  ms = sorted(a.elapsed_time(b) for a, b in evs)
  nk = int(keep.sum())
  alg = bench.algorithmic_bytes(n, nk, C, 4)
  peak = bench.measured_peak()[0]
  r = dict(quantised=quantised, kept=nk, n_rec=n_rec, ms_min=ms[0], ms_med=ms[len(ms) // 2], ms_max=ms[-1], frac_med=alg / ms[len(ms) // 2] / 1e6 / peak,
           mcand_s=n / ms[len(ms) // 2] / 1e3)
  print(json.dumps(r), flush=True)
  del arena, scg, rhc
  torch.cuda.empty_cache()


if __name__ == '__main__':
  q = '--quantised' in sys.argv
  for n_rec in [int(a) for a in sys.argv[1:] if not a.startswith('--')] or [125, 250, 500, 1000, 1500]:
    run(n_rec, quantised=q)
