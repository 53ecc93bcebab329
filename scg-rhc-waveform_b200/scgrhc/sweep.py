"""Sweep scheduler (BASELINE configs[4]): many waveform_NN configs over ONE device-resident cohort.

The reference runs its sweep as 28 independent end-to-end jobs (`waveform_pipeline.py:33-37`), re-reading every
record from disk per config.  Here the cohort is uploaded once; configs are grouped by (chamber, window length) so the
interval plan is built once per group (the 32 loadable configs are 4 chambers x 8 channel subsets), and each
config is one fused kernel pass.  Records shard across ranks exactly as for a single config; the only collective is the
min/max all-reduce of configs with ``use_global_min_max``.
"""
import torch

from . import engine

SAMPLE_FREQ = 500


def iter_sweep(arena, sig_name, metas, record_rows, configs, out_dtype=torch.float32, group=None, buffers=None):
  """Yield ``(name, WindowStore)`` for every config in ``configs`` (name -> object with ``in_channels``, ``chamber``,
  ``segment_size``, ``min_RHC``, ``use_global_min_max``).  Stores of successive configs reuse ``buffers`` (pass a dict)
  so that a long sweep does not hold 37 cohorts of windows in HBM: consume each store before advancing."""
  plans = {}
  order = sorted(configs, key=lambda k: (str(configs[k].chamber), float(configs[k].segment_size), k))
  for name in order:
    c = configs[name]
    W = int(c.segment_size * SAMPLE_FREQ)
    key = (c.chamber, W)
    if key not in plans:
      plans[key] = engine.plan_cohort(metas, c.chamber, record_rows, W)
    cols, rcol = engine.resolve_columns(sig_name, c.in_channels)
    store = engine.prepare_windows(arena, plans[key], cols, rcol, c.min_RHC,
                                   use_global_min_max=bool(c.use_global_min_max), out_dtype=out_dtype, group=group,
                                   buffers=buffers)
    yield name, store
