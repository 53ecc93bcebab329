"""Sweep scheduler (BASELINE configs[4]): many waveform_NN configs over ONE device-resident cohort.

The reference runs its sweep as 28 independent end-to-end jobs (`waveform_pipeline.py:33-37`), re-reading every
record from disk per config.  Here the cohort is uploaded once; configs are grouped by (chamber, window length) so the
interval plan is built once per group (the 32 loadable configs are 4 chambers x 8 channel subsets); the
configs of one chamber share one predicate pass (has_noise looks at the RHC channel only) and one fan-out pass that reads
every kept window once and writes it for all 8 subsets.  Records shard across ranks exactly as for a single config; the only collective is the
min/max all-reduce of configs with ``use_global_min_max``.
"""
import torch

from . import engine

SAMPLE_FREQ = 500


def iter_sweep(arena, sig_name, metas, record_rows, configs, out_dtype=torch.float32, group=None, buffers=None, fan_out=True,
               rec0=0):
  """Yield ``(name, WindowStore)`` for every config in ``configs`` (name -> object with ``in_channels``, ``chamber``,
  ``segment_size``, ``min_RHC``, ``use_global_min_max``).

  ``fan_out`` (default): configs that differ only in their channel subset (same chamber, window length and ``min_RHC``,
  local min-max, at most 4 distinct channels between them, each list in signal order) share ONE predicate pass and ONE
  normalisation pass that reads every kept window once and writes it per subset (`engine.prepare_subsets`); their
  stores are dense and share the RHC tensor.  Everything else takes one fused pass per config; those stores reuse
  ``buffers`` (pass a dict) so that a long sweep does not hold 37 cohorts of windows in HBM: consume each store before
  advancing.  ``rec0``: record number of the first record of ``arena`` (a rank's shard of a larger cohort); with a
  process ``group`` every rank must iterate the same configs in the same order (they do: the order is a function of
  ``configs`` alone), because configs with ``use_global_min_max`` all-reduce inside."""
  plans = {}
  tabs = engine.event_tabs(metas)          # event times of every side-car, once for all chambers
  order = sorted(configs, key=lambda k: (str(configs[k].chamber), float(configs[k].segment_size), k))

  def plan_for(c):
    W = int(c.segment_size * SAMPLE_FREQ)
    key = (c.chamber, W)
    if key not in plans:
      plans[key] = engine.plan_cohort(metas, c.chamber, record_rows, W, rec0=rec0, tabs=tabs)
    return plans[key]

  groups = {}
  if fan_out:
    for name in order:
      c = configs[name]
      if bool(c.use_global_min_max) or getattr(c, 'normalisation', None) not in (None, 'minmax'):
        continue
      cols, rcol = engine.resolve_columns(sig_name, c.in_channels)
      if cols != sorted(cols) or len(set(cols)) != len(cols):
        continue
      groups.setdefault((c.chamber, int(c.segment_size * SAMPLE_FREQ), float(c.min_RHC), rcol), []).append((name, cols))
  done = set()
  for key, members in groups.items():
    sup = sorted({c for _, cols in members for c in cols})
    if len(members) < 2 or len(sup) > 4:
      continue
    first = configs[members[0][0]]
    stores = engine.prepare_subsets(arena, plan_for(first), sup, key[3], first.min_RHC, [cols for _, cols in members],
                                    out_dtype=out_dtype)
    for (name, _), store in zip(members, stores):
      done.add(name)
      yield name, store
    del stores
  for name in order:
    if name in done:
      continue
    c = configs[name]
    cols, rcol = engine.resolve_columns(sig_name, c.in_channels)
    store = engine.prepare_windows(arena, plan_for(c), cols, rcol, c.min_RHC,
                                   use_global_min_max=bool(c.use_global_min_max), out_dtype=out_dtype, group=group,
                                   buffers=buffers)
    yield name, store
