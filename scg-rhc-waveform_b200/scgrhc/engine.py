"""Host side of the hot path: planning (which windows exist) and driving the device ops.

  plan_record / plan_cohort   recordutil.get_chamber_intervals + window enumeration of get_segments
                              (recordutil.py:93-110,138-146), via the C planner scgrhc_plan_record
  prepare_windows             has_noise + SCGDataset.init_segments for every candidate window of a
                              device-resident cohort (recordutil.py:141-148,55-66,152-169)
"""
import ctypes as C
import os
from dataclasses import dataclass, field
from datetime import datetime
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _native as N
from . import ops

SAMPLE_FREQ = 500
FLAT_THRESHOLD = 1e-3
RHC_NAME = 'RHC_pressure'

INTERVAL_DTYPE = np.dtype([('row0', '<i8'), ('cand0', '<i8'), ('n_win', '<i4'), ('rec_id', '<i4')])


_duration_cache = {}


def _record_seconds(start, end):
  """(MacEndTime - MacStTime).total_seconds() with the date ignored (recordutil.py:100-101,104); strptime is the slow
  part of planning a large cohort, so the result is cached by the two strings."""
  key = (start, end)
  v = _duration_cache.get(key)
  if v is None:
    t0 = datetime.strptime(start.split()[1], '%H:%M:%S')
    t1 = datetime.strptime(end.split()[1], '%H:%M:%S')
    v = (t1 - t0).total_seconds()
    if len(_duration_cache) > 100000:
      _duration_cache.clear()
    _duration_cache[key] = v
  return v


def event_times(meta):
  """(keys, times) of a side-car: event times in dict order with 'END' appended (recordutil.py:100-104); None when
  ChamEvents_in_s is not a dict (:103 -> no intervals)."""
  end = _record_seconds(meta['MacStTime'], meta['MacEndTime'])
  events = meta['ChamEvents_in_s']
  if not isinstance(events, dict):
    return None
  events = dict(events)
  events['END'] = end
  keys = list(events.keys())
  return keys, np.array([float(events[k]) for k in keys], dtype=np.float64)


def event_table(meta, chamber):
  """Event times in dict order with 'END' appended (recordutil.py:100-104) and the per-event
  chamber match ``key.split('_')[0] == chamber`` (:108).  None when ChamEvents_in_s is not a dict
  (:103 -> no intervals)."""
  kt = event_times(meta)
  if kt is None:
    return None
  keys, times = kt
  # '*' (extension, waveform_01 legacy default): every chamber event, i.e. no chamber segmentation
  match = np.array([(k != 'END') if chamber == '*' else (k.split('_')[0] == chamber) for k in keys], dtype=np.uint8)
  return times, match


def plan_record(meta, chamber, T, W, rec_base_row=0, rec_id=0, cand_base=0, stride=0, fs=0.0):
  """(intervals structured array, n_cand, bounds) for one record via the C planner."""
  tab = event_table(meta, chamber)
  if tab is None or len(tab[0]) < 2:
    return np.zeros(0, dtype=INTERVAL_DTYPE), 0, []
  times, match = tab
  n = len(times)
  out = (N.Interval * n)()
  bounds = (C.c_int64 * (2 * n))()
  n_out, n_b, n_cand = C.c_int(0), C.c_int(0), C.c_int64(0)
  rc = N.lib().scgrhc_plan_record(times.ctypes.data_as(C.POINTER(C.c_double)),
                                  match.ctypes.data_as(C.POINTER(C.c_uint8)), n, int(T), int(W), int(stride), float(fs),
                                  int(rec_base_row), int(rec_id), int(cand_base), out, n,
                                  C.byref(n_out), C.byref(n_cand), bounds, n, C.byref(n_b))
  if rc != N.OK:
    raise N.ScgrhcError(rc, 'scgrhc_plan_record failed')
  iv = np.frombuffer(out, dtype=INTERVAL_DTYPE, count=n_out.value).copy()
  b = [(bounds[2 * k], bounds[2 * k + 1]) for k in range(n_b.value)]
  return iv, n_cand.value, b


@dataclass
class Plan:
  """Candidate windows of a cohort laid out in one (rows, nsig) arena."""
  intervals: np.ndarray            # INTERVAL_DTYPE, cand0 ascending
  n_cand: int
  W: int
  record_names: List[str] = field(default_factory=list)
  stride: int = 0                  # 0 = W (non-overlapping, the reference); extension: rows between window starts
  _dev: Optional[torch.Tensor] = None

  def device_intervals(self, device):
    if self._dev is None or self._dev.device != torch.device(device):
      host = torch.from_numpy(self.intervals.view(np.int64).reshape(-1, 3).copy()) if len(self.intervals) else \
          torch.zeros((0, 3), dtype=torch.int64)
      self._dev = host.to(device)
    return self._dev


class EventTabs:
  """The chamber-independent part of planning a cohort, once: event times of every side-car flattened into one array with
  per-record offsets, and the chamber prefix ``key.split('_')[0]`` of every event (recordutil.py:100-108).  A sweep plans
  the same cohort once per chamber; with this table a plan is one vectorised string compare + one C call."""

  def __init__(self, metas):
    times, prefix, is_end, off = [], [], [], [0]
    for meta in metas:
      kt = event_times(meta)
      if kt is not None and len(kt[1]) >= 2:
        times.append(kt[1])
        prefix.extend(k.split('_')[0] for k in kt[0])
        is_end.extend(k == 'END' for k in kt[0])
        off.append(off[-1] + len(kt[0]))
      else:
        off.append(off[-1])
    self.n_rec = len(metas)
    self.times = np.ascontiguousarray(np.concatenate(times), dtype=np.float64) if times else np.zeros(0, dtype=np.float64)
    self.prefix = np.array(prefix, dtype=object) if prefix else np.zeros(0, dtype=object)
    self.is_end = np.array(is_end, dtype=bool) if is_end else np.zeros(0, dtype=bool)
    self.off = np.asarray(off, dtype=np.int64)

  def match(self, chamber):
    """uint8 per event: does it open an interval of ``chamber`` ('*' = every chamber event, the waveform_01 extension)."""
    if chamber == '*':
      return np.ascontiguousarray(~self.is_end, dtype=np.uint8)
    if self.prefix.dtype.kind == 'S':          # tables of the native scanner (scgrhc.hostscan): fixed-width bytes
      return np.ascontiguousarray(self.prefix == chamber.encode(), dtype=np.uint8)
    return np.ascontiguousarray(self.prefix == chamber, dtype=np.uint8)


def event_tabs(metas):
  return getattr(metas, 'tabs', None) or EventTabs(metas)


def plan_cohort(metas, chamber, T_rows, W, record_names=None, stride=0, fs=0.0, rec0=0, tabs=None):
  """Plan for records stored back to back in the arena; ``T_rows[r]`` rows each.  One C call for the whole cohort
  (`scgrhc_plan_cohort`).  ``rec0``: record number of the first record (a rank's shard of a larger cohort reports global
  record ids).  ``tabs``: an ``EventTabs`` of the same side-cars, when several plans are made of one cohort."""
  if tabs is None:
    tabs = getattr(metas, 'tabs', None) or EventTabs(metas)      # hostscan.ScannedMetas carries its table
  t, m, off = tabs.times, tabs.match(chamber), tabs.off
  n_rec = len(metas)
  names = list(record_names) if record_names is not None else []
  if off[-1] == 0:
    return Plan(np.zeros(0, dtype=INTERVAL_DTYPE), 0, W, names, stride)
  o = off
  rows = np.ascontiguousarray(np.asarray(T_rows[:n_rec], dtype=np.int64))
  cap = int(m.sum())
  out = np.zeros(max(cap, 1), dtype=INTERVAL_DTYPE)
  n_out, n_cand = C.c_int64(0), C.c_int64(0)
  rc = N.lib().scgrhc_plan_cohort(t.ctypes.data_as(C.POINTER(C.c_double)), m.ctypes.data_as(C.POINTER(C.c_uint8)),
                                  o.ctypes.data_as(C.POINTER(C.c_int64)), rows.ctypes.data_as(C.POINTER(C.c_int64)), n_rec,
                                  int(W), int(stride), float(fs), int(rec0), out.ctypes.data_as(C.POINTER(N.Interval)), cap,
                                  C.byref(n_out), C.byref(n_cand))
  if rc != N.OK:
    raise N.ScgrhcError(rc, 'scgrhc_plan_cohort failed')
  return Plan(out[:n_out.value].copy(), n_cand.value, W, names, stride)


def plan_uniform(meta, chamber, T, W, n_rec, rec0=0, stride=0):
  """Every record shares one side-car (synthetic cohorts): plan record 0 in C, replicate with offsets."""
  iv0, n0, _ = plan_record(meta, chamber, T, W, 0, 0, 0, stride)
  k = len(iv0)
  iv = np.tile(iv0, n_rec)
  r = np.repeat(np.arange(n_rec, dtype=np.int64), k)
  iv['row0'] += r * T
  iv['cand0'] += r * n0
  iv['rec_id'] = (r + rec0).astype(np.int32)
  return Plan(iv, n0 * n_rec, W, [], stride)


@dataclass
class WindowStore:
  """Device-resident result of the hot path.  Window tensors are *slot*-addressed: slot == candidate
  index when ``dense`` is False (rejected candidates leave unwritten holes), list position when True.
  ``kept_idx`` lists the kept candidates in the reference's order."""
  scg: Optional[torch.Tensor]        # (slots, C, W)
  rhc: Optional[torch.Tensor]        # (slots, 1, W)
  minmax: torch.Tensor               # (n_cand, 4) fp64: scg_min, scg_max, rhc_min, rhc_max
  keep: torch.Tensor                 # (n_cand,) uint8
  reason: torch.Tensor               # (n_cand,) uint8
  kept_idx: torch.Tensor             # (n_kept,) int64
  start_idx: torch.Tensor            # (n_kept,) int64, relative to the interval (recordutil.py:143)
  stop_idx: torch.Tensor             # (n_kept,) int64
  rec_id: torch.Tensor               # (n_kept,) int32
  n_kept: int
  n_cand: int
  dense: bool
  global_minmax: Optional[torch.Tensor] = None   # (4,) fp64 when use_global_min_max
  minmax_dense: bool = False                     # minmax rows are in list order (sweep fan-out), not per candidate
  shard: Optional[object] = None                 # scgrhc.dist.Shard: this rank's offset / the total in the cohort-wide ordered list
  n_ambiguous: int = 0                           # candidates whose R^2 lies within 1e-9 of 0.8 (REASON_AMBIGUOUS): the only
                                                 # windows the closed form could decide differently from sklearn's lstsq

  def slots(self):
    return torch.arange(self.n_kept, device=self.kept_idx.device) if self.dense else self.kept_idx

  def kept_minmax(self):
    if self.global_minmax is not None:
      return self.global_minmax.unsqueeze(0).expand(self.n_kept, 4)
    if self.minmax_dense:
      return self.minmax[:self.n_kept]
    return self.minmax[self.kept_idx]

  def gather(self, positions):
    """(scg, rhc) of the kept windows at list positions ``positions`` (int64 device tensor)."""
    slots = positions if self.dense else self.kept_idx[positions]
    n = slots.numel()
    scg = torch.empty((n,) + tuple(self.scg.shape[1:]), dtype=self.scg.dtype, device=self.scg.device)
    rhc = torch.empty((n,) + tuple(self.rhc.shape[1:]), dtype=self.rhc.dtype, device=self.rhc.device)
    if n:
      ops.gather_windows(self.scg, slots.contiguous(), scg)
      ops.gather_windows(self.rhc, slots.contiguous(), rhc)
    return scg, rhc

  def materialise(self):
    """Dense (n_kept, C, W) / (n_kept, 1, W) tensors in the reference's order."""
    if self.dense:
      return self.scg[:self.n_kept], self.rhc[:self.n_kept]
    return self.gather(torch.arange(self.n_kept, device=self.kept_idx.device))


def resolve_columns(sig_name, in_channels):
  """`sig_name.index(name)` for every channel (recordutil.py:117); ValueError if absent, as there."""
  sig_name = list(sig_name)
  return [sig_name.index(n) for n in in_channels], sig_name.index(RHC_NAME)


def prepare_windows(arena, plan, scg_cols, rhc_col, min_rhc, use_global_min_max=False, out_dtype=torch.float32,
                    predicates_only=False, keep_all=False, flat_threshold=FLAT_THRESHOLD, group=None,
                    check=True, buffers=None, normalisation='minmax', planar=False, decim=None):
  """Run the hot path over every candidate window of ``plan``.

  Local normalisation (default): ONE fused kernel pass — predicates, min/max, normalise, transpose,
  cast — then the ordered compaction of the keep flags.  ``use_global_min_max``: pass A (predicates
  + per-window pairs), device reduction (+ MIN all-reduce of {min, -max} over ``group`` when
  torch.distributed is initialised, recordutil.py:152-169 seen across shards), pass B over the kept
  list writing dense outputs.

  ``normalisation='zscore'`` (extension, absent from the reference): (x - mean) / (std + 1e-4) per window, mean and
  population std taken jointly over the SCG block and over the RHC window; the store's ``minmax`` then holds
  (scg_mean, scg_std, rhc_mean, rhc_std).
  """
  if not arena.is_cuda:
    raise RuntimeError('prepare_windows needs a CUDA arena (no CPU fallback)')
  zflag = _norm_flag(normalisation, use_global_min_max)
  dev = arena.device
  n, W, Cn = plan.n_cand, plan.W, len(scg_cols)
  iv = plan.device_intervals(dev)
  f64 = out_dtype == torch.float64
  base_flags = (N.OUT_F64 if f64 else 0) | (N.KEEP_ALL if keep_all else 0) | zflag | (N.ARENA_PLANAR if planar else 0)
  b = buffers if buffers is not None else {}

  def buf(name, shape, dtype):
    t = b.get(name)
    if t is None or t.shape != torch.Size(shape) or t.dtype != dtype:
      t = torch.empty(shape, dtype=dtype, device=dev)
      if buffers is not None:
        buffers[name] = t
    return t

  minmax = buf('minmax', (n, 4), torch.float64)
  keep = buf('keep', (n,), torch.uint8)
  reason = buf('reason', (n,), torch.uint8)
  cand_win = buf('cand_win', (n,), torch.int32)
  cand_rec = buf('cand_rec', (n,), torch.int32)
  kept_idx = buf('kept_idx', (n,), torch.int64)
  start_idx = buf('start_idx', (n,), torch.int64)
  stop_idx = buf('stop_idx', (n,), torch.int64)
  rec_id = buf('rec_id', (n,), torch.int32)
  n_kept_t = buf('n_kept', (1,), torch.int64)
  two_pass = use_global_min_max and not predicates_only
  scg = rhc = None
  if not predicates_only and not two_pass:
    scg = buf('scg', (n, Cn, W), out_dtype)
    rhc = buf('rhc', (n, 1, W), out_dtype)
  flags = base_flags | (N.PREDICATES_ONLY if (predicates_only or two_pass) else 0)
  if decim is not None:
    if two_pass or f64 or planar:
      raise ValueError('decim: fp32 outputs with per-window pairs on an interleaved arena only')
    t_in0, t_len, t_rel = decim.tables(plan, dev)
    ops.process_windows_decim(arena, iv, n, W, plan.stride, list(scg_cols), rhc_col, float(min_rhc), float(flat_threshold), flags,
                              decim.taps, decim.per_phase, decim.down, decim.n_pre_remove, t_in0, t_len, t_rel,
                              scg, rhc, minmax, keep, reason, cand_win, cand_rec, fused=decim.fused)
  else:
    ops.process_windows(arena, iv, n, W, plan.stride, list(scg_cols), rhc_col, float(min_rhc), float(flat_threshold), flags,
                        [0.0] * 4, None, 0, scg, rhc, minmax, keep, reason, cand_win, cand_rec)
  ops.compact_kept(keep, cand_win, cand_rec, n, W, plan.stride, kept_idx, start_idx, stop_idx, rec_id, n_kept_t)
  gmm = None
  if use_global_min_max:
    gmm = buf('gmm', (4,), torch.float64)
    ops.global_minmax(minmax, keep, n, gmm)
    gmm = allreduce_minmax(gmm, group)
  n_amb = ops.check_errors(dev.index) if check else 0
  n_kept = int(n_kept_t.item())
  dense = False
  if two_pass:
    scg = buf('scg', (n_kept, Cn, W), out_dtype)
    rhc = buf('rhc', (n_kept, 1, W), out_dtype)
    if n_kept:
      ops.process_windows(arena, iv, n, W, plan.stride, list(scg_cols), rhc_col, float(min_rhc), float(flat_threshold),
                          base_flags | N.USE_KEPT_LIST | N.NORM_GLOBAL, gmm.cpu().tolist(), kept_idx, n_kept,
                          scg, rhc, minmax, None, None, None, None)
    dense = True
  return WindowStore(scg, rhc, minmax, keep, reason, kept_idx[:n_kept], start_idx[:n_kept], stop_idx[:n_kept],
                     rec_id[:n_kept], n_kept, n, dense, gmm, n_ambiguous=n_amb)


def prepare_subsets(arena, plan, sup_cols, rhc_col, min_rhc, subsets, out_dtype=torch.float32,
                    flat_threshold=FLAT_THRESHOLD, check=True):
  """Sweep fan-out: several channel subsets of one cohort / chamber in two passes instead of one fused
  pass per subset.  ``sup_cols``: the superset of SCG columns (<= 4, arena column indices); ``subsets``: lists of arena
  column indices, each a subsequence of ``sup_cols`` (channel order = that order).  has_noise() looks at the RHC
  channel only (waveform_noise.py:44-49), so ONE predicate pass serves every subset; then every kept window is read
  once and written per subset (`scgrhc_normalize_subsets`).  Returns one dense WindowStore per subset; they share the
  RHC tensor, the keep flags and the kept list.  Outputs are bit-identical to ``prepare_windows`` per subset."""
  if not arena.is_cuda:
    raise RuntimeError('prepare_subsets needs a CUDA arena (no CPU fallback)')
  sup = list(sup_cols)
  members, counts = [], []
  for sub in subsets:
    idx = [sup.index(c) for c in sub]
    if idx != sorted(idx) or len(set(idx)) != len(idx) or not idx:
      raise ValueError('subset %r is not a subsequence of the superset %r' % (list(sub), sup))
    members += idx
    counts.append(len(idx))
  dev = arena.device
  pred = prepare_windows(arena, plan, sup[:1], rhc_col, min_rhc, predicates_only=True, flat_threshold=flat_threshold,
                         check=check)
  n_kept, W = pred.n_kept, plan.W
  rhc = torch.empty((n_kept, 1, W), dtype=out_dtype, device=dev)
  scgs = [torch.empty((n_kept, c, W), dtype=out_dtype, device=dev) for c in counts]
  mms = [torch.empty((n_kept, 4), dtype=torch.float64, device=dev) for _ in counts]
  stores = []
  for g0 in range(0, len(counts), N.MAX_SUBSETS):
    g1 = min(len(counts), g0 + N.MAX_SUBSETS)
    if n_kept:
      m0 = sum(counts[:g0])
      ops.normalize_subsets(arena, plan.device_intervals(dev), W, plan.stride, sup, rhc_col, pred.kept_idx, n_kept,
                            members[m0:m0 + sum(counts[g0:g1])], counts[g0:g1], out_dtype == torch.float64,
                            scgs[g0:g1], mms[g0:g1], rhc)
  for scg, mm in zip(scgs, mms):
    stores.append(WindowStore(scg, rhc, mm, pred.keep, pred.reason, pred.kept_idx, pred.start_idx, pred.stop_idx,
                              pred.rec_id, n_kept, pred.n_cand, True, None, True, pred.n_ambiguous))
  return stores


def _norm_flag(normalisation, use_global_min_max):
  if normalisation in (None, 'minmax'):
    return 0
  if normalisation != 'zscore':
    raise ValueError("normalisation must be 'minmax' (the reference) or 'zscore', got %r" % (normalisation,))
  if use_global_min_max:
    raise ValueError("normalisation='zscore' is per window; it cannot be combined with use_global_min_max")
  return N.NORM_ZSCORE


def allreduce_minmax(gmm, group=None):
  """Dataset-level (scg_min, scg_max, rhc_min, rhc_max) across record shards: one MIN all-reduce of
  {min, -max} (exact and order independent, so bit-identical to the unsharded reference value).
  A rank whose shard kept nothing contributes (+inf, -inf)."""
  import torch.distributed as dist
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
    return gmm
  sign = torch.tensor([1.0, -1.0, 1.0, -1.0], dtype=gmm.dtype, device=gmm.device)
  v = gmm * sign
  if dist.get_backend(group) != 'nccl' and v.is_cuda:     # gloo (CPU tests, 2 ranks sharing one GPU): host copy of 32 bytes
    h = v.cpu()
    dist.all_reduce(h, op=dist.ReduceOp.MIN, group=group)
    v = h.to(gmm.device)
  else:
    dist.all_reduce(v, op=dist.ReduceOp.MIN, group=group)
  return v * sign


def shard_records(n_rec, rank, world):
  """Contiguous block of record indices owned by ``rank`` (records are independent, recordutil.py:131-132)."""
  lo = n_rec * rank // world
  hi = n_rec * (rank + 1) // world
  return lo, hi


class PinnedArenaSource:
  """Chunk source: the cohort already sits in (pinned) host memory, records back to back."""

  def __init__(self, host_arena):
    self.host = host_arena

  def begin(self, ingest):
    pass

  def enqueue(self, k, chunk, stage):        # called with the copy stream current
    stage.copy_(self.host[chunk[0]:chunk[1]], non_blocking=True)

  def end(self):
    pass


class SynthSource:
  """Chunk source: records generated on the device (scgrhc_synth_records) — cohorts that never exist on the host, e.g.
  BASELINE configs[3]'s 100k records = 960 GB of fp64 ("must be generated on device in chunks")."""

  def __init__(self, seed, T, kinds, defect_scale=16, grid=750, rec0=0):
    self.seed, self.T, self.kinds, self.defect_scale, self.grid, self.rec0 = seed, int(T), list(kinds), defect_scale, grid, rec0

  def begin(self, ingest):
    self.planar = bool(getattr(ingest, 'planar_run', False))

  def enqueue(self, k, chunk, stage):
    r0, r1 = chunk[5], chunk[6]
    ops.synth_records(stage, self.seed, self.rec0 + r0, r1 - r0, self.T, self.kinds, self.defect_scale, self.grid,
                      plane_stride=(r1 - r0) * self.T if self.planar else 0)

  def end(self):
    pass


class DiskSource:
  """Chunk source: WFDB format-16 ``.dat`` files -> a ring of pinned chunk buffers -> HBM (the streaming replacement of the
  reference's per-record ``wfdb.rdrecord``, recordutil.py:137).  A small pool of reader threads fills ring slot k % R with
  the frames of chunk k (``readinto`` straight into pinned memory: no intermediate host copy, the GIL is released during
  the read); the H2D copy of chunk k is enqueued as soon as its reads have landed, while chunks k+1 .. k+R-1 are still
  being read and chunk k-1 is in the window kernel.  Host memory: R chunks, whatever the cohort size."""

  def __init__(self, dat_paths, record_rows, nsig_file, ring=3, workers=None, byte_offsets=None):
    self.paths, self.rows, self.nsig = list(dat_paths), [int(r) for r in record_rows], int(nsig_file)
    if workers is None:           # SCGRHC_READERS, else three quarters of the host's cores shared between the ranks of this box
      # (16 cores, one rank, 500 records on tmpfs: 4 readers 60 ms, 8: 41 ms, 12: 35 ms, 16: 40 ms, 24: 46 ms)
      workers = int(os.environ.get('SCGRHC_READERS', 0)) or \
          max(4, min(16, 3 * (os.cpu_count() or 8) // (4 * max(1, int(os.environ.get('LOCAL_WORLD_SIZE', 1))))))
    self.ring, self.workers = max(2, int(ring)), max(1, int(workers))
    self.offsets = list(byte_offsets) if byte_offsets is not None else [0] * len(self.paths)
    self.slots = None
    self.bytes_read = 0

  def _read(self, r, view):
    want = view.nbytes
    with open(self.paths[r], 'rb', buffering=0) as f:
      if self.offsets[r]:
        f.seek(self.offsets[r])
      got, mv = 0, memoryview(view).cast('B')
      while got < want:
        n = f.readinto(mv[got:])
        if not n:
          raise IOError('%s: %d bytes short of the %d frames its header announces' % (self.paths[r], want - got, self.rows[r]))
        got += n
    return want

  def _submit(self, j):
    r0, r1 = self.ingest.chunks[j][5], self.ingest.chunks[j][6]
    buf = self.slots[j % self.ring]
    at, futs = 0, []
    for r in range(r0, r1):
      futs.append(self.pool.submit(self._read, r, buf[at:at + self.rows[r]]))
      at += self.rows[r]
    self.futures[j] = futs

  def begin(self, ingest):
    from concurrent.futures import ThreadPoolExecutor
    self.ingest = ingest
    max_rows = ingest.max_chunk_rows
    if self.slots is None or self.slots[0].shape[0] < max_rows:
      self.slots = [torch.empty((max_rows, self.nsig), dtype=torch.int16, pin_memory=True).numpy() for _ in range(self.ring)]
      self.pinned = [torch.from_numpy(a) for a in self.slots]
    self.pool = ThreadPoolExecutor(self.workers)
    self.futures, self.copied = {}, {}
    for j in range(self.ring - 1):
      if ingest._ensure_chunk(j):
        self._submit(j)

  def enqueue(self, k, chunk, stage):
    nxt = k + self.ring - 1
    if self.ingest._ensure_chunk(nxt):            # lazy ingests parse + plan that chunk now, while earlier ones are in flight
      if k >= 1:
        self.copied.pop(k - 1).synchronize()      # slot (k-1) % R is free once chunk k-1 has left host memory
      self._submit(nxt)
    for f in self.futures.pop(k):
      self.bytes_read += f.result()
    stage.copy_(self.pinned[k % self.ring][:chunk[1] - chunk[0]], non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(stage.device))
    self.copied[k] = ev

  def end(self):
    self.pool.shutdown(wait=True)
    self.copied.clear()
    self.ingest = None                            # no ingest <-> source cycle: the pinned ring goes back to the allocator at once


class HostIngest:
  """Host-resident (or disk-resident, or device-generated) cohort -> HBM -> hot path, chunked at record boundaries and
  double buffered: the pinned-host -> device copy of chunk k+1 (copy stream) overlaps the window kernel of chunk k
  (compute stream).  This is the end-to-end path a caller with records outside HBM uses (the reference reads each record
  from disk into host numpy arrays, recordutil.py:137)."""

  def __init__(self, plan, record_rows, nsig, device, chunk_records=64, digital_nsig=None, stages=None, planar=None):
    """``nsig``: columns of the fp64 arena the window kernel reads.  ``digital_nsig``: the host cohort is WFDB
    format-16 int16 frames with that many signals per frame; they are copied as int16 (4x fewer PCIe bytes than
    fp64 physical samples) and converted on the device (scgrhc_decode_fmt16).

    ``stages`` (extension, default off): the optional per-record stages run on every chunk between the copy and the
    window kernel, so they overlap the next chunk's copy and the cohort never has to be resident:
    ``{'sos': (n, 6) sections, 'filter_cols': [...], 'filter_exact': bool, 'resample': (up, down), 'out_rows': [...]}``;
    with ``resample`` the plan's rows are rows AFTER resampling (``out_rows`` per record).

    ``planar`` (default: on for digital cohorts of three or more SCG channels without optional stages): the fp64 chunk
    arenas on the device are PLANAR (one plane per signal, csrc/window_planar_kernel.cuh): the device decode (digital
    cohorts) or generator (SynthSource) writes that layout at no extra cost, a rejected window then costs 6 KB of DRAM
    traffic instead of 24 KB (traffic == the algorithmic bytes) and the kernel runs at 0.92 of the HBM peak against 0.84-0.89
    for the interleaved one.  With one or two SCG channels the interleaved kernel is the faster one (its
    per-window overhead is smaller: 0.79 / 0.89 against 0.65 / 0.83), so those cohorts keep interleaved rows.  Not for fp64
    host cohorts (they arrive interleaved over PCIe), nor with the optional stages (the filters read interleaved rows), nor
    with z-score normalisation (those runs decode to interleaved rows)."""
    self.plan, self.nsig, self.device = plan, nsig, torch.device(device)
    self.stages = stages or None
    self.planar = (digital_nsig is not None and not stages and nsig >= 4) if planar is None else bool(planar)
    if self.planar and stages:
      raise ValueError('planar chunk arenas cannot feed the optional filter / resample stages (they read interleaved rows)')
    self.planar_run = False
    rows = np.asarray(record_rows, dtype=np.int64)
    self.record_rows = rows
    base = np.concatenate([[0], np.cumsum(rows)])
    self.record_base = base
    self.total_rows = int(base[-1])
    plan_base = base
    if self.stages and self.stages.get('resample'):
      plan_base = np.concatenate([[0], np.cumsum(np.asarray(self.stages['out_rows'], dtype=np.int64))])
    iv = plan.intervals
    self.chunks = []
    max_rows = 0
    for r0 in range(0, len(rows), chunk_records):
      r1 = min(len(rows), r0 + chunk_records)
      lo, hi = int(base[r0]), int(base[r1])
      a, b = np.searchsorted(iv['row0'], [int(plan_base[r0]), int(plan_base[r1])], side='left')
      sub = iv[a:b].copy()
      # a chunk without candidate windows still needs the running candidate prefix: pass B of use_global_min_max cuts the
      # kept list at these values, so they must be monotonic
      cand_lo = int(iv['cand0'][a]) if a < len(iv) else int(plan.n_cand)
      n = int(sub['n_win'].sum())
      sub['row0'] -= int(plan_base[r0])
      sub['cand0'] -= cand_lo
      t = torch.from_numpy(sub.view(np.int64).reshape(-1, 3).copy()).to(self.device) if len(sub) else None
      self.chunks.append((lo, hi, cand_lo, n, t, r0, r1))
      max_rows = max(max_rows, hi - lo)
    self.max_chunk_rows = max_rows
    self.W, self.stride, self.n_alloc, self.n_total = plan.W, plan.stride, plan.n_cand, plan.n_cand
    # Integer decimation with bit-identical taps can run INSIDE the window kernel (scgrhc_process_windows_decim): the
    # resampled chunk is never written.  Per chunk: the native-rate position / length of every interval's record.
    self.decim, self._decim_tabs, self._fuse_decim = None, None, False
    if self.stages and self.stages.get('resample') and self.stages.get('fuse_decim', os.environ.get('SCGRHC_FUSE_DECIM', '1') != '0') \
        and nsig == 4 and plan.W <= 384:
      import math
      from . import filters
      up, down = self.stages['resample']
      g = math.gcd(int(up), int(down))
      if up // g == 1 and down // g >= 2:
        try:
          self.decim = filters.DecimSpec.design(rows, up, down, fused=not self.stages.get('resample_exact', True))
        except NotImplementedError:
          self.decim = None
      if self.decim is not None:
        self._decim_tabs = []
        for (lo, hi, cand_lo, n, t, r0, r1) in self.chunks:
          a, b = np.searchsorted(iv['row0'], [int(plan_base[r0]), int(plan_base[r1])], side='left')
          row0 = iv['row0'][a:b].astype(np.int64)
          r = np.searchsorted(plan_base, row0, side='right') - 1                    # record of every interval
          mk = lambda v: torch.from_numpy(np.ascontiguousarray(v, dtype=np.int64)).to(self.device)
          self._decim_tabs.append((mk(base[r] - lo), mk(rows[r]), mk(row0 - plan_base[r])) if len(row0) else None)
    self.bufs = [torch.empty((max_rows, nsig), dtype=torch.float64, device=self.device) for _ in range(2)]
    self.digital_nsig = digital_nsig
    self.dbufs = [torch.empty((max_rows, digital_nsig), dtype=torch.int16, device=self.device) for _ in range(2)] \
        if digital_nsig else None
    self.copy_stream = torch.cuda.Stream(self.device)
    self.h2d_bytes = self.total_rows * (digital_nsig * 2 if digital_nsig else nsig * 8)
    self._tables = None

  def _ensure_chunk(self, k):
    """True iff chunk k exists (lazy ingests build it on demand)."""
    return k < len(self.chunks)

  def _chunk_decode(self, k, decode):
    """(record offsets, longest record, gain table, baseline table, recip) of chunk k for the one-launch decode."""
    gain_t, base_t, recip, offs, longest = self._decode_tables(decode)
    r0, r1 = self.chunks[k][5], self.chunks[k][6]
    return offs[k], longest[k], gain_t[r0:r1], base_t[r0:r1], recip

  def _decode_tables(self, decode):
    """Per-record calibration as device tables (gain, baseline: (n_rec, ncols) fp64) + per-chunk record offsets, so that
    the decode of a chunk is ONE launch whatever the number of records in it."""
    cols, gain, baseline = decode
    key = id(gain), id(baseline)
    if self._tables is not None and self._tables[0] == key:
      return self._tables[1:]
    g = np.ascontiguousarray(np.asarray(gain, dtype=np.float64).reshape(len(self.record_rows), len(cols)))
    b = np.ascontiguousarray(np.asarray(baseline, dtype=np.float64).reshape(len(self.record_rows), len(cols)))
    ag = np.abs(g)
    recip = bool(g.size and (ag >= 2.0 ** -40).all() and (ag <= 2.0 ** 60).all() and (np.abs(b) < 65536.0).all() and (b == np.floor(b)).all())
    offs = [torch.from_numpy(np.ascontiguousarray(self.record_base[c[5]:c[6] + 1] - c[0])).to(self.device) for c in self.chunks]
    longest = [int(self.record_rows[c[5]:c[6]].max()) if c[6] > c[5] else 0 for c in self.chunks]
    t = (torch.from_numpy(g).to(self.device), torch.from_numpy(b).to(self.device), recip, offs, longest)
    self._tables = (key,) + t
    return t

  def _stream(self, source, decode, body, planar=False):
    """Copy (and, for digital cohorts, decode) chunk after chunk, double buffered, and hand each resident chunk to
    ``body(dst, chunk)`` on the compute stream.  ``source``: a CPU tensor (the whole cohort in host memory) or a chunk
    source (PinnedArenaSource / DiskSource / SynthSource)."""
    dev = self.device
    if isinstance(source, torch.Tensor):
      source = PinnedArenaSource(source)
    compute = torch.cuda.current_stream(dev)
    done = [None, None]
    self.copy_stream.wait_stream(compute)
    digital = self.digital_nsig is not None
    if digital and decode is None:
      raise ValueError('digital cohort: decode=(cols, gain, baseline) is required')
    per_record = digital and (decode[1] is None or (len(decode[1]) and isinstance(decode[1][0], (list, tuple, np.ndarray))))
    self.planar_run = bool(planar)
    source.begin(self)
    try:
      k = -1
      while self._ensure_chunk(k + 1):
        k += 1
        chunk = self.chunks[k]
        lo, hi, cand_lo, nc, iv, r0, r1 = chunk[:7]
        dst = self.bufs[k & 1][:hi - lo]
        if planar:                  # (nsig, rows of this chunk): one contiguous plane per signal
          dst = self.bufs[k & 1].view(-1)[:(hi - lo) * self.nsig].view(self.nsig, hi - lo)
        stage = self.dbufs[k & 1][:hi - lo] if digital else dst
        plane = (hi - lo) if planar else 0
        with torch.cuda.stream(self.copy_stream):
          if done[k & 1] is not None:
            self.copy_stream.wait_event(done[k & 1])
          source.enqueue(k, chunk, stage)
          ready = torch.cuda.Event()
          ready.record(self.copy_stream)
        compute.wait_event(ready)
        if digital and nc:
          if per_record:       # every record has its own gain / baseline (WFDB headers): device tables, one launch per chunk
            off_k, longest_k, gain_k, base_k, recip = self._chunk_decode(k, decode)
            ops.decode_fmt16_records(stage, off_k, longest_k, list(decode[0]), gain_k, base_k, recip, dst, plane)
          else:
            ops.decode_fmt16(stage, list(decode[0]), [float(v) for v in decode[1]], [float(v) for v in decode[2]], dst, plane)
        if nc:
          body(self._run_stages(dst, r0, r1), chunk, k)
        done[k & 1] = torch.cuda.Event()
        done[k & 1].record(compute)
    finally:
      source.end()

  def _run_stages(self, dst, r0, r1):
    """Optional per-record stages on one resident chunk (records r0..r1): zero-phase band-pass, then resampling."""
    st = self.stages
    if not st:
      return dst
    from . import filters
    rows = [int(v) for v in self.record_rows[r0:r1]]
    if st.get('sos') is not None:
      exact = bool(st.get('filter_exact', True)) or len(st['sos']) > 4
      dst = filters.sosfiltfilt(dst, rows, st['sos'], list(st['filter_cols']), exact=exact, inplace=not exact)
    if st.get('resample') and not self._fuse_decim:
      up, down = st['resample']
      dst, _ = filters.resample_poly(dst, rows, up, down, exact=bool(st.get('resample_exact', True)))
    return dst

  def run(self, host_arena, scg_cols, rhc_col, min_rhc, out_dtype=torch.float32, flat_threshold=FLAT_THRESHOLD,
          buffers=None, decode=None, use_global_min_max=False, group=None, normalisation='minmax', sink=None):
    """``host_arena``: (total_rows, nsig) fp64 CPU tensor (pinned for an asynchronous copy), or — with
    ``digital_nsig`` — (total_rows, digital_nsig) int16 frames plus ``decode = (cols, gain, baseline)`` where
    gain/baseline are one list per selected column (all records) or one such list per record; or a chunk source
    (``DiskSource``: format-16 files streamed through a pinned ring; ``SynthSource``: generated on the device).

    ``use_global_min_max`` (recordutil.py:185-186) streams the cohort twice — predicates + per-window pairs, the
    dataset-level reduction (+ all-reduce over ``group``), then normalisation of the kept windows with the global pairs
    into dense outputs — so cohorts larger than HBM (BASELINE configs[3]: 100k records) never have to be resident.

    ``sink(k, scg, rhc, info)``: streamed OUTPUTS for cohorts whose windows do not fit in HBM either (100k records are
    480 GB of fp32 windows): the window tensors of chunk k are written into a two-slot ring and handed to ``sink`` on
    the compute stream right after the chunk's kernel (``info``: candidate range, and in global mode the kept-list
    range — the tensors are then dense); the consumer must enqueue its reads on the current stream.  The returned
    store then carries no window tensors, only the per-candidate metadata and the ordered kept list."""
    dev = self.device
    n, W, Cn, stride = self.n_alloc, self.W, len(scg_cols), self.stride      # n: allocation (an upper bound for lazy ingests)
    b = buffers if buffers is not None else {}

    def buf(name, shape, dtype):
      t = b.get(name)
      if t is None or t.shape != torch.Size(shape) or t.dtype != dtype:
        t = b[name] = torch.empty(shape, dtype=dtype, device=dev)
      return t

    minmax, keep, reason = buf('minmax', (n, 4), torch.float64), buf('keep', (n,), torch.uint8), buf('reason', (n,), torch.uint8)
    cand_win, cand_rec = buf('cand_win', (n,), torch.int32), buf('cand_rec', (n,), torch.int32)
    kept_idx, start_idx, stop_idx = (buf(k, (n,), torch.int64) for k in ('kept_idx', 'start_idx', 'stop_idx'))
    rec_id, n_kept_t = buf('rec_id', (n,), torch.int32), buf('n_kept', (1,), torch.int64)
    planar = self.planar and normalisation in (None, 'minmax')
    if planar and self.digital_nsig is None and isinstance(host_arena, (torch.Tensor, PinnedArenaSource)):
      raise ValueError('planar=True needs a source that writes planes (digital cohort or SynthSource); fp64 host rows are interleaved')
    base_flags = (N.OUT_F64 if out_dtype == torch.float64 else 0) | _norm_flag(normalisation, use_global_min_max) | \
        (N.ARENA_PLANAR if planar else 0)
    # decimation inside the window kernel when it applies (fp32 outputs, per-window pairs); else the resample stage runs
    self._fuse_decim = fuse = getattr(self, 'decim', None) is not None and not use_global_min_max and out_dtype == torch.float32 \
        and len(scg_cols) <= 3
    scg = rhc = None
    ring = None
    if sink is not None:
      cmax = getattr(self, 'max_chunk_cand', None) or max((c[3] for c in self.chunks), default=0)
      ring = [(buf('scg_ring%d' % i, (cmax, Cn, W), out_dtype), buf('rhc_ring%d' % i, (cmax, 1, W), out_dtype)) for i in range(2)]
    elif not use_global_min_max:
      scg, rhc = buf('scg', (n, Cn, W), out_dtype), buf('rhc', (n, 1, W), out_dtype)
    launched = [0]

    def pass_a(dst, chunk, k):
      lo, hi, cand_lo, nc, iv, r0, r1 = chunk[:7]
      flags = base_flags | (N.PREDICATES_ONLY if use_global_min_max else 0) | (N.KEEP_ERRORS if launched[0] else 0)
      so = ro = None
      if ring is not None and not use_global_min_max:
        so, ro = ring[k & 1]
      elif scg is not None:
        so, ro = scg[cand_lo:], rhc[cand_lo:]
      if fuse:
        dc, (t_in0, t_len, t_rel) = self.decim, self._decim_tabs[k]
        ops.process_windows_decim(dst, iv, nc, W, stride, list(scg_cols), rhc_col, float(min_rhc), float(flat_threshold), flags,
                                  dc.taps, dc.per_phase, dc.down, dc.n_pre_remove, t_in0, t_len, t_rel, so, ro,
                                  minmax[cand_lo:], keep[cand_lo:], reason[cand_lo:], cand_win[cand_lo:], cand_rec[cand_lo:], fused=dc.fused)
      else:
        ops.process_windows(dst, iv, nc, W, stride, list(scg_cols), rhc_col, float(min_rhc), float(flat_threshold), flags,
                            [0.0] * 4, None, 0, so, ro,
                            minmax[cand_lo:], keep[cand_lo:], reason[cand_lo:], cand_win[cand_lo:], cand_rec[cand_lo:])
      launched[0] += 1
      if ring is not None and not use_global_min_max:
        sink(k, so[:nc], ro[:nc], {'cand_lo': cand_lo, 'n_cand': nc, 'dense': False, 'keep': keep[cand_lo:cand_lo + nc]})

    self._stream(host_arena, decode, pass_a, planar)
    n = self.n_total                      # lazy ingests know the candidate count only now
    minmax, keep, reason = minmax[:n], keep[:n], reason[:n]
    ops.compact_kept(keep, cand_win, cand_rec, n, W, stride, kept_idx, start_idx, stop_idx, rec_id, n_kept_t)
    gmm = None
    if use_global_min_max:
      gmm = buf('gmm', (4,), torch.float64)
      ops.global_minmax(minmax, keep, n, gmm)
      gmm = allreduce_minmax(gmm, group)
    n_amb = 0
    if n and launched[0]:
      n_amb = ops.check_errors(dev.index)      # raises ValueError like the reference if a non-finite RHC window reached the regression
    n_kept = int(n_kept_t.item())      # device -> host read of the step's result
    if use_global_min_max:
      if ring is None:
        scg, rhc = buf('scg', (n_kept, Cn, W), out_dtype), buf('rhc', (n_kept, 1, W), out_dtype)
      kept = kept_idx[:n_kept]
      edges = torch.tensor([c[2] for c in self.chunks] + [n], dtype=torch.int64, device=dev)
      pos = torch.searchsorted(kept, edges).cpu().tolist()        # kept-list range of every chunk (plumbing, not arithmetic)
      gm = gmm.cpu().tolist()

      def pass_b(dst, chunk, k):
        lo, hi, cand_lo, nc, iv, r0, r1 = chunk[:7]
        a, e = pos[k], pos[k + 1]
        if e > a:
          so, ro = (ring[k & 1]) if ring is not None else (scg[a:], rhc[a:])
          ops.process_windows(dst, iv, nc, W, stride, list(scg_cols), rhc_col, float(min_rhc), float(flat_threshold),
                              base_flags | N.USE_KEPT_LIST | N.NORM_GLOBAL, gm, (kept[a:e] - cand_lo).contiguous(), e - a,
                              so, ro, None, None, None, None, None)
          if ring is not None:
            sink(k, so[:e - a], ro[:e - a], {'cand_lo': cand_lo, 'n_cand': nc, 'dense': True, 'kept_lo': a, 'kept_hi': e})

      if n_kept:
        self._stream(host_arena, decode, pass_b, planar)
    return WindowStore(scg, rhc, minmax, keep, reason, kept_idx[:n_kept], start_idx[:n_kept], stop_idx[:n_kept],
                       rec_id[:n_kept], n_kept, n, bool(use_global_min_max), gmm, n_ambiguous=n_amb)


class LazyDiskIngest(HostIngest):
  """HostIngest for cohorts of format-16 FILES that are parsed, planned, read, copied and processed chunk by chunk: the
  per-record host work of the reference's loop (side-car JSON, header, interval maths — recordutil.py:96-109,137) for chunk
  k+2 runs on the CPU while chunk k+1 is being read into the pinned ring and chunk k is in the window kernel, instead of
  all of it up front.  Only the file SIZES are needed before the first byte moves (buffer capacities); nothing about the
  cohort is held in host memory beyond the ring.

  ``rows_est[r]``: upper bound of the frames of record r (file size / frame bytes); ``parse_chunk(r0, r1)`` ->
  ``(dat_paths, rows, gains, baselines, metas)`` of records r0..r1; ``plan_chunk(metas, rows, r0)`` -> engine.Plan of the
  chunk alone (rows and candidates chunk-local, record ids global)."""

  def __init__(self, rows_est, nsig, device, W, nsig_file, parse_chunk, plan_chunk, chunk_records=32, planar=None):
    self.plan, self.nsig, self.device = None, nsig, torch.device(device)
    self.stages, self.planar, self.planar_run = None, (nsig >= 4 if planar is None else bool(planar)), False
    est = [int(v) for v in rows_est]
    self.n_records, self.chunk_records = len(est), int(chunk_records)
    self.record_rows = np.zeros(len(est), dtype=np.int64)
    self.bounds = [(r0, min(len(est), r0 + self.chunk_records)) for r0 in range(0, len(est), self.chunk_records)]
    self.max_chunk_rows = max((sum(est[a:b]) for a, b in self.bounds), default=0)
    self.max_chunk_cand = max((sum(v // W for v in est[a:b]) for a, b in self.bounds), default=0)
    self.W, self.stride = int(W), 0
    self.n_alloc, self.n_total = sum(v // W for v in est), 0
    self.parse_chunk, self.plan_chunk = parse_chunk, plan_chunk
    self.chunks, self._dec = [], []
    self.total_rows = 0
    self.bufs = [torch.empty((self.max_chunk_rows, nsig), dtype=torch.float64, device=self.device) for _ in range(2)]
    self.digital_nsig = int(nsig_file)
    self.dbufs = [torch.empty((self.max_chunk_rows, self.digital_nsig), dtype=torch.int16, device=self.device) for _ in range(2)]
    self.copy_stream = torch.cuda.Stream(self.device)
    self.h2d_bytes = 0
    self.source = DiskSource([], [], self.digital_nsig)

  def _ensure_chunk(self, k):
    while len(self.chunks) <= k and len(self.chunks) < len(self.bounds):
      r0, r1 = self.bounds[len(self.chunks)]
      paths, rows, gains, bases, metas = self.parse_chunk(r0, r1)
      rows = [int(v) for v in rows]
      plan = self.plan_chunk(metas, rows, r0)
      if sum(rows) > self.max_chunk_rows or plan.n_cand > self.max_chunk_cand or plan.W != self.W or plan.stride not in (0, self.W):
        raise RuntimeError('record files changed under the ingest (chunk %d larger than its size estimate)' % len(self.chunks))
      self.record_rows[r0:r1] = rows
      self.source.paths.extend(paths); self.source.rows.extend(rows); self.source.offsets.extend([0] * len(paths))
      iv = plan.device_intervals(self.device) if plan.n_cand else None
      g = np.ascontiguousarray(np.asarray(gains, dtype=np.float64).reshape(len(rows), -1))
      b = np.ascontiguousarray(np.asarray(bases, dtype=np.float64).reshape(len(rows), -1))
      ag = np.abs(g)
      recip = bool(g.size and (ag >= 2.0 ** -40).all() and (ag <= 2.0 ** 60).all() and (np.abs(b) < 65536.0).all() and (b == np.floor(b)).all())
      off = torch.from_numpy(np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)).to(self.device)
      self._dec.append((off, max(rows) if rows else 0, torch.from_numpy(g).to(self.device), torch.from_numpy(b).to(self.device), recip))
      lo = self.total_rows
      self.total_rows += sum(rows)
      self.chunks.append((lo, self.total_rows, self.n_total, plan.n_cand, iv, r0, r1))
      self.n_total += plan.n_cand
      self.h2d_bytes = self.total_rows * self.digital_nsig * 2
    return k < len(self.chunks)

  def _chunk_decode(self, k, decode):
    return self._dec[k]

  def run_files(self, scg_cols, rhc_col, min_rhc, sel_cols, **kw):
    """``run`` over the files: ``sel_cols`` = the file's signal columns that become arena columns 0..nsig-1."""
    return self.run(self.source, scg_cols, rhc_col, min_rhc, decode=(list(sel_cols), None, None), **kw)
