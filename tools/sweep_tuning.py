"""Kernel-only timing of the window kernel over (CTAs/SM, stages) — run on the GPU box."""
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


def main():
  n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 500
  dev = torch.device('cuda', 0)
  arena = torch.empty((n_rec * bench.T_ROWS, 4), dtype=torch.float64, device=dev)
  ops.synth_records(arena, bench.SEED, 0, n_rec, bench.T_ROWS, bench.KINDS, 16, bench.W)
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
  peak = bench.measured_peak()[0]
  res = []
  for flags, label in ((0, 'full'), (2, 'pred_only')):
    for stages in (2, 3, 4):
      for ctas in (1, 2, 3, 4):
        try:
          ops.set_tuning(0, ctas, stages)
          def step():
            ops.process_windows(arena, iv, n, W, 0, [0, 1, 2], 3, -50.0, 1e-3, flags, [0.0] * 4, None, 0,
                                scg, rhc, minmax, keep, reason, cw, cr)
          for _ in range(3):
            step()
          torch.cuda.synchronize()
          a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          a.record()
          for _ in range(10):
            step()
          b.record()
          torch.cuda.synchronize()
          ms = a.elapsed_time(b) / 10
          nk = int(keep.sum())
          alg = bench.algorithmic_bytes(n, nk if flags == 0 else 0, C, 4) if flags == 0 else n * 4 * W * 8
          r = dict(mode=label, stages=stages, ctas_per_sm=ctas, ms=ms, gbs=alg / ms / 1e6, frac=alg / ms / 1e6 / peak,
                   mwin_s=n / ms / 1e3)
        except Exception as e:
          r = dict(mode=label, stages=stages, ctas_per_sm=ctas, error=str(e)[:100])
        res.append(r)
        print(json.dumps(r), flush=True)


if __name__ == '__main__':
  main()
