"""PyTorch custom ops (namespace ``scgrhc``) over the C ABI of libscgrhc.so.

Only CUDA implementations are registered: calling an op on CPU tensors raises, there is no
fallback.  torch supplies device memory and the current stream; all arithmetic is in the library.
"""
import atexit
import ctypes as C
from typing import List, Optional, Sequence

import torch
from torch import Tensor

from . import _native as N

_ctx = {}


def ctx(device_index):
  """One library context per CUDA device (created on first use)."""
  c = _ctx.get(device_index)
  if c is None:
    L = N.lib()
    h = C.c_void_p()
    rc = L.scgrhc_ctx_create(int(device_index), C.byref(h))
    if rc != N.OK:
      raise N.ScgrhcError(rc, (L.scgrhc_last_error(None) or b'').decode())
    c = _ctx[device_index] = h
  return c


@atexit.register
def _destroy_contexts():
  while _ctx:
    _, h = _ctx.popitem()
    N.lib().scgrhc_ctx_destroy(h)


def set_tuning(device_index, ctas_per_sm=0, stages=0):
  N.check(ctx(device_index), N.lib().scgrhc_ctx_set_tuning(ctx(device_index), int(ctas_per_sm), int(stages)))


def sm_count(device_index):
  return N.lib().scgrhc_ctx_sm_count(ctx(device_index))


def _dev(t):
  if not t.is_cuda:
    raise RuntimeError('scgrhc ops need CUDA tensors (no CPU fallback)')
  return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _stream(dev):
  return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ptr(t):
  return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class _output_planes:
  """Planar output layout for the next decode / synth call (sticky per context in the C ABI: always reset here)."""

  def __init__(self, c, plane_stride):
    self.c, self.p = c, int(plane_stride)

  def __enter__(self):
    if self.p:
      N.check(self.c, N.lib().scgrhc_ctx_set_output_planes(self.c, self.p))

  def __exit__(self, *exc):
    if self.p:
      N.lib().scgrhc_ctx_set_output_planes(self.c, 0)


def _contig(t, dtype, name):
  if t is None:
    return
  if t.dtype != dtype or not t.is_contiguous():
    raise ValueError('%s must be a contiguous %s tensor' % (name, dtype))


@torch.library.custom_op('scgrhc::process_windows',
                         mutates_args=('scg_out', 'rhc_out', 'minmax', 'keep', 'reason', 'cand_win', 'cand_rec'),
                         device_types='cuda')
def process_windows(arena: Tensor, intervals: Tensor, n_cand: int, W: int, stride: int, scg_cols: Sequence[int], rhc_col: int,
                    min_rhc: float, flat_threshold: float, flags: int, global_minmax: Sequence[float],
                    kept_list: Optional[Tensor], n_items: int, scg_out: Optional[Tensor], rhc_out: Optional[Tensor],
                    minmax: Optional[Tensor], keep: Optional[Tensor], reason: Optional[Tensor],
                    cand_win: Optional[Tensor], cand_rec: Optional[Tensor]) -> None:
  """has_noise + SCGDataset.init_segments fused (recordutil.py:141-148,55-66; waveform_noise.py:6-49)."""
  dev = _dev(arena)
  if arena.dim() != 2:
    raise ValueError('arena must be (rows, nsig), or (nsig, rows) with ARENA_PLANAR')
  _contig(arena, torch.float64, 'arena')
  planar = bool(flags & N.ARENA_PLANAR)
  _contig(intervals, torch.int64, 'intervals')
  out_dtype = torch.float64 if flags & N.OUT_F64 else torch.float32
  _contig(scg_out, out_dtype, 'scg_out'); _contig(rhc_out, out_dtype, 'rhc_out')
  _contig(minmax, torch.float64, 'minmax'); _contig(keep, torch.uint8, 'keep'); _contig(reason, torch.uint8, 'reason')
  _contig(cand_win, torch.int32, 'cand_win'); _contig(cand_rec, torch.int32, 'cand_rec')
  _contig(kept_list, torch.int64, 'kept_list')
  if len(scg_cols) > N.MAX_C:
    raise N.ScgrhcError(N.ERR_UNSUPPORTED, 'at most %d SCG channels' % N.MAX_C)
  j = N.Job()
  j.arena = arena.data_ptr()
  j.arena_rows = arena.shape[1 if planar else 0]
  j.arena_capacity_bytes = arena.numel() * 8
  j.nsig = arena.shape[0 if planar else 1]
  j.W = W
  j.stride = stride
  j.C = len(scg_cols)
  for i, c in enumerate(scg_cols):
    j.scg_cols[i] = c
  j.rhc_col = rhc_col
  j.flags = flags
  j.intervals = intervals.data_ptr()
  j.n_intervals = intervals.shape[0]
  j.n_cand = n_cand
  j.min_rhc = min_rhc
  j.flat_threshold = flat_threshold
  for i in range(4):
    j.global_minmax[i] = global_minmax[i] if len(global_minmax) == 4 else 0.0
  j.kept_list = kept_list.data_ptr() if kept_list is not None else None
  j.n_items = n_items
  slots = n_items if flags & N.USE_KEPT_LIST else n_cand
  if not flags & N.PREDICATES_ONLY and slots:
    if scg_out is None or rhc_out is None or scg_out.numel() < slots * j.C * W or rhc_out.numel() < slots * W:
      raise ValueError('scg_out/rhc_out too small for %d slots' % slots)
  for t, nm in ((minmax, 4), (keep, 1), (reason, 1), (cand_win, 1), (cand_rec, 1)):
    if t is not None and t.numel() < n_cand * nm:
      raise ValueError('per-candidate output too small')
  o = N.Outputs(_ptr(scg_out), _ptr(rhc_out), _ptr(minmax), _ptr(keep), _ptr(reason), _ptr(cand_win), _ptr(cand_rec))
  c = ctx(dev)
  N.check(c, N.lib().scgrhc_process_windows(c, C.byref(j), C.byref(o), _stream(dev)))


def process_windows_decim(arena, intervals, n_cand, W, stride, scg_cols, rhc_col, min_rhc, flat_threshold, flags, taps, per_phase, down,
                          n_pre_remove, iv_in0, iv_len, iv_rel, scg_out, rhc_out, minmax, keep, reason, cand_win, cand_rec, fused=False):
  """Extension: process_windows with the decimating front end (scgrhc_process_windows_decim): ``arena`` holds native-rate
  (rows, 4) records, the windows are W rows at the model rate; ``taps`` = numpy fp64 (per_phase,), the three ``iv_*`` =
  int64 CUDA tensors (n_intervals,)."""
  dev = _dev(arena)
  _contig(arena, torch.float64, 'arena'); _contig(intervals, torch.int64, 'intervals')
  for t, nm in ((iv_in0, 'iv_in0'), (iv_len, 'iv_len'), (iv_rel, 'iv_rel')):
    _contig(t, torch.int64, nm)
    if t.numel() < intervals.shape[0]:
      raise ValueError('%s: one entry per interval' % nm)
  _contig(scg_out, torch.float32, 'scg_out'); _contig(rhc_out, torch.float32, 'rhc_out')
  _contig(minmax, torch.float64, 'minmax'); _contig(keep, torch.uint8, 'keep'); _contig(reason, torch.uint8, 'reason')
  _contig(cand_win, torch.int32, 'cand_win'); _contig(cand_rec, torch.int32, 'cand_rec')
  if arena.dim() != 2 or arena.shape[1] != 4:
    raise N.ScgrhcError(N.ERR_UNSUPPORTED, 'the decimating front end reads (rows, 4) arenas')
  j = N.Job()
  j.arena = arena.data_ptr(); j.arena_rows = arena.shape[0]; j.arena_capacity_bytes = arena.numel() * 8; j.nsig = 4
  j.W = W; j.stride = stride; j.C = len(scg_cols)
  for i, c in enumerate(scg_cols):
    j.scg_cols[i] = c
  j.rhc_col = rhc_col; j.flags = flags
  j.intervals = intervals.data_ptr(); j.n_intervals = intervals.shape[0]; j.n_cand = n_cand
  j.min_rhc = min_rhc; j.flat_threshold = flat_threshold
  if not flags & N.PREDICATES_ONLY and n_cand:
    if scg_out is None or rhc_out is None or scg_out.numel() < n_cand * j.C * W or rhc_out.numel() < n_cand * W:
      raise ValueError('scg_out/rhc_out too small for %d slots' % n_cand)
  for t, nm in ((minmax, 4), (keep, 1), (reason, 1), (cand_win, 1), (cand_rec, 1)):
    if t is None or t.numel() < n_cand * nm:
      raise ValueError('per-candidate output too small')
  import numpy as np
  tp = np.ascontiguousarray(taps, dtype=np.float64)
  d = N.Decim(tp.ctypes.data, int(per_phase), int(down), int(n_pre_remove), 1 if fused else 0, iv_in0.data_ptr(), iv_len.data_ptr(), iv_rel.data_ptr())
  o = N.Outputs(_ptr(scg_out), _ptr(rhc_out), _ptr(minmax), _ptr(keep), _ptr(reason), _ptr(cand_win), _ptr(cand_rec))
  c = ctx(dev)
  N.check(c, N.lib().scgrhc_process_windows_decim(c, C.byref(j), C.byref(o), C.byref(d), _stream(dev)))


@torch.library.custom_op('scgrhc::compact_kept', mutates_args=('kept_idx', 'start_idx', 'stop_idx', 'rec_id', 'n_kept'),
                         device_types='cuda')
def compact_kept(keep: Tensor, cand_win: Tensor, cand_rec: Tensor, n_cand: int, W: int, stride: int, kept_idx: Tensor,
                 start_idx: Tensor, stop_idx: Tensor, rec_id: Tensor, n_kept: Tensor) -> None:
  """Ordered list of kept windows = the order of the list get_segments returns (recordutil.py:148)."""
  dev = _dev(keep)
  _contig(keep, torch.uint8, 'keep'); _contig(cand_win, torch.int32, 'cand_win'); _contig(cand_rec, torch.int32, 'cand_rec')
  for t, nm in ((kept_idx, 'kept_idx'), (start_idx, 'start_idx'), (stop_idx, 'stop_idx'), (n_kept, 'n_kept')):
    _contig(t, torch.int64, nm)
  _contig(rec_id, torch.int32, 'rec_id')
  co = N.Compact(_ptr(kept_idx), _ptr(start_idx), _ptr(stop_idx), _ptr(rec_id), _ptr(n_kept), stride, 0)
  c = ctx(dev)
  N.check(c, N.lib().scgrhc_compact_kept(c, _ptr(keep), _ptr(cand_win), _ptr(cand_rec), n_cand, W, C.byref(co), _stream(dev)))


@torch.library.custom_op('scgrhc::normalize_subsets', mutates_args=('scg_outs', 'minmaxes', 'rhc_out'), device_types='cuda')
def normalize_subsets(arena: Tensor, intervals: Tensor, W: int, stride: int, sup_cols: Sequence[int], rhc_col: int,
                      kept_list: Tensor, n_items: int, members: Sequence[int], member_counts: Sequence[int], out_f64: bool,
                      scg_outs: List[Tensor], minmaxes: List[Tensor], rhc_out: Tensor) -> None:
  """Sweep fan-out: SCGDataset.init_segments (recordutil.py:55-66) for several channel subsets of one superset over the
  kept windows of one predicate pass; subset k = members[sum(counts[:k]) : sum(counts[:k+1])] (ascending indices into
  ``sup_cols``).  Dense outputs in list order; bit-identical to process_windows run per subset."""
  dev = _dev(arena)
  _contig(arena, torch.float64, 'arena'); _contig(intervals, torch.int64, 'intervals'); _contig(kept_list, torch.int64, 'kept_list')
  out_dtype = torch.float64 if out_f64 else torch.float32
  _contig(rhc_out, out_dtype, 'rhc_out')
  if len(member_counts) != len(scg_outs) or len(scg_outs) != len(minmaxes) or not 1 <= len(scg_outs) <= N.MAX_SUBSETS:
    raise ValueError('1..%d subsets, one scg_out and one minmax each' % N.MAX_SUBSETS)
  if len(sup_cols) > N.MAX_C:
    raise N.ScgrhcError(N.ERR_UNSUPPORTED, 'at most %d superset channels' % N.MAX_C)
  j = N.Job()
  j.arena = arena.data_ptr()
  j.arena_rows = arena.shape[0]
  j.arena_capacity_bytes = arena.numel() * 8
  j.nsig = arena.shape[1]
  j.W = W
  j.stride = stride
  j.C = len(sup_cols)
  for i, c in enumerate(sup_cols):
    j.scg_cols[i] = c
  j.rhc_col = rhc_col
  j.flags = N.OUT_F64 if out_f64 else 0
  j.intervals = intervals.data_ptr()
  j.n_intervals = intervals.shape[0]
  j.kept_list = kept_list.data_ptr()
  j.n_items = n_items
  if kept_list.numel() < n_items or rhc_out.numel() < n_items * W:
    raise ValueError('kept_list / rhc_out too small for %d windows' % n_items)
  subs = (N.Subset * len(scg_outs))()
  at = 0
  for k, cnt in enumerate(member_counts):
    _contig(scg_outs[k], out_dtype, 'scg_out'); _contig(minmaxes[k], torch.float64, 'minmax')
    if scg_outs[k].numel() < n_items * cnt * W or minmaxes[k].numel() < n_items * 4:
      raise ValueError('subset %d outputs too small' % k)
    subs[k].scg_out = scg_outs[k].data_ptr(); subs[k].minmax = minmaxes[k].data_ptr(); subs[k].C = cnt
    for i in range(cnt):
      subs[k].member[i] = members[at + i]
    at += cnt
  c = ctx(dev)
  N.check(c, N.lib().scgrhc_normalize_subsets(c, C.byref(j), subs, len(scg_outs), _ptr(rhc_out), _stream(dev)))


@torch.library.custom_op('scgrhc::global_minmax', mutates_args=('mm_out',), device_types='cuda')
def global_minmax(minmax: Tensor, keep: Tensor, n_cand: int, mm_out: Tensor) -> None:
  """get_global_minmax_vals over the kept windows (recordutil.py:152-169)."""
  dev = _dev(minmax)
  _contig(minmax, torch.float64, 'minmax'); _contig(keep, torch.uint8, 'keep'); _contig(mm_out, torch.float64, 'mm_out')
  c = ctx(dev)
  N.check(c, N.lib().scgrhc_global_minmax(c, _ptr(minmax), _ptr(keep), n_cand, _ptr(mm_out), _stream(dev)))


@torch.library.custom_op('scgrhc::gather_windows', mutates_args=('out',), device_types='cuda')
def gather_windows(store: Tensor, slots: Tensor, out: Tensor) -> None:
  """Batch collate (default_collate at recordutil.py:198): out[b] = store[slots[b]]."""
  dev = _dev(store)
  _contig(slots, torch.int64, 'slots')
  if not store.is_contiguous() or not out.is_contiguous() or store.dtype != out.dtype:
    raise ValueError('store/out must be contiguous and of one dtype')
  wb = store[0].numel() * store.element_size()
  if out.numel() * out.element_size() < wb * slots.numel():
    raise ValueError('out too small')
  c = ctx(dev)
  N.check(c, N.lib().scgrhc_gather_windows(c, _ptr(store), _ptr(slots), slots.numel(), wb, _ptr(out), _stream(dev)))


@torch.library.custom_op('scgrhc::gather_windows_noise', mutates_args=('out',), device_types='cuda')
def gather_windows_noise(store: Tensor, slots: Tensor, out: Tensor, sigma: float, seed: int, offset: int) -> None:
  """Extension (absent from the reference): batch gather + sigma * N(0,1) from Philox4x32-10 / Box-Muller."""
  dev = _dev(store)
  _contig(slots, torch.int64, 'slots'); _contig(store, torch.float32, 'store'); _contig(out, torch.float32, 'out')
  E = store[0].numel()
  if out.numel() < E * slots.numel():
    raise ValueError('out too small')
  c = ctx(dev)
  N.check(c, N.lib().scgrhc_gather_windows_noise(c, _ptr(store), _ptr(slots), slots.numel(), E, _ptr(out), C.c_float(sigma),
                                                 C.c_uint64(seed & (2 ** 64 - 1)), C.c_uint64(offset & (2 ** 64 - 1)), _stream(dev)))


class BatchCollator:
  """The train loop's per-batch call (default_collate of recordutil.py:198 for the tensors waveform_train.py:358-359
  reads): ONE launch of scgrhc_collate_batch writes the SCG batch (+ Philox noise) and the RHC batch from one slot list.
  Bound through plain ctypes with everything resolved once per epoch — the torch.library dispatch of the generic ops
  costs more than the kernel at batch sizes of a few hundred windows."""

  def __init__(self, scg_store, rhc_store, slots_dev):
    for t, nm in ((scg_store, 'scg_store'), (rhc_store, 'rhc_store')):
      _contig(t, torch.float32, nm)
    _contig(slots_dev, torch.int64, 'slots')
    self.dev = _dev(scg_store)
    if not (rhc_store.is_cuda and slots_dev.is_cuda) or rhc_store.shape[0] != scg_store.shape[0]:
      raise ValueError('collate: stores and slots must live on one CUDA device and hold the same windows')
    self.scg_store, self.rhc_store, self.slots = scg_store, rhc_store, slots_dev      # keep them alive
    self.E_s, self.E_r = scg_store[0].numel() if scg_store.shape[0] else 1, rhc_store[0].numel() if rhc_store.shape[0] else 1
    self.ctx, self.fn = ctx(self.dev), N.lib().scgrhc_collate_batch
    self.p_s, self.p_r, self.p_i = scg_store.data_ptr(), rhc_store.data_ptr(), slots_dev.data_ptr()

  def __call__(self, first, n, scg_out, rhc_out, sigma=0.0, seed=0, offset=0):
    """scg_out[b], rhc_out[b] = stores[slots[first + b]] for b < n (outputs: contiguous fp32 CUDA tensors, large enough)."""
    if first < 0 or first + n > self.slots.numel() or scg_out.numel() < n * self.E_s or rhc_out.numel() < n * self.E_r:
      raise ValueError('collate: slot range or output buffers out of bounds')
    stream = torch.cuda.current_stream(self.dev).cuda_stream
    args = (self.ctx, self.p_s, self.p_r, self.p_i + 8 * first, n, self.E_s, self.E_r, scg_out.data_ptr(), rhc_out.data_ptr(),
            sigma, seed & 0xFFFFFFFFFFFFFFFF, offset & 0xFFFFFFFFFFFFFFFF, stream)
    if torch.cuda.current_device() != self.dev:
      with torch.cuda.device(self.dev):
        rc = self.fn(*args)
    else:
      rc = self.fn(*args)
    if rc:
      N.check(self.ctx, rc)


def philox_words(device_index, seed, offset, nquads):
  """Raw Philox4x32-10 blocks (nquads, 4) uint32 as int64 numpy array, for seed-exact checks."""
  out = torch.empty((nquads, 4), dtype=torch.int32, device='cuda:%d' % device_index)
  c = ctx(device_index)
  N.check(c, N.lib().scgrhc_philox_words(c, C.c_uint64(seed), C.c_uint64(offset), nquads, _ptr(out), _stream(device_index)))
  return out.cpu().numpy().astype('int64') & 0xFFFFFFFF


@torch.library.custom_op('scgrhc::window_metrics', mutates_args=('out',), device_types='cuda')
def window_metrics(real: Tensor, pred: Tensor, minmax: Tensor, out: Tensor) -> None:
  """Per window: de-normalise real/pred (n, W) fp32 with (rhc_min, rhc_max) and return Pearson r and RMSE
  (waveform_test.py:21-50,66-70: reverse_minmax, pearsonr, sqrt(mean_squared_error))."""
  dev = _dev(real)
  _contig(real, torch.float32, 'real'); _contig(pred, torch.float32, 'pred')
  _contig(minmax, torch.float64, 'minmax'); _contig(out, torch.float64, 'out')
  n, W = real.shape[0], real[0].numel()
  if pred.numel() != real.numel() or minmax.numel() < 2 * n or out.numel() < 2 * n:
    raise ValueError('window_metrics: shape mismatch')
  c = ctx(dev)
  N.check(c, N.lib().scgrhc_window_metrics(c, _ptr(real), _ptr(pred), _ptr(minmax), n, W, _ptr(out), _stream(dev)))


@torch.library.custom_op('scgrhc::rolling_range_lt', mutates_args=('flags',), device_types='cuda')
def rolling_range_lt(y: Tensor, m: int, threshold: float, flags: Tensor) -> None:
  """flags[p] = rolling(m).max - rolling(m).min < threshold (waveform_noise.py:10-13)."""
  dev = _dev(y)
  _contig(y, torch.float64, 'y'); _contig(flags, torch.uint8, 'flags')
  c = ctx(dev)
  N.check(c, N.lib().scgrhc_rolling_range_lt(c, _ptr(y), y.numel(), m, threshold, _ptr(flags), _stream(dev)))


@torch.library.custom_op('scgrhc::decode_fmt16', mutates_args=('out',), device_types='cuda')
def decode_fmt16(d: Tensor, cols: Sequence[int], gain: Sequence[float], baseline: Sequence[float], out: Tensor,
                 plane_stride: int = 0) -> None:
  """WFDB format-16 frames (T, nsig) int16 -> physical fp64 (T, len(cols)): (d - baseline) / gain, -32768 -> NaN
  (the host-side dac of wfdb.rdrecord, recordutil.py:137, moved onto the device)."""
  dev = _dev(d)
  _contig(d, torch.int16, 'd'); _contig(out, torch.float64, 'out')
  n = len(cols)
  if d.dim() != 2 or out.numel() < d.shape[0] * n or len(gain) != n or len(baseline) != n:
    raise ValueError('decode_fmt16: d must be (T, nsig), out (T, len(cols)); one gain and baseline per column')
  c = ctx(dev)
  with _output_planes(c, plane_stride):        # plane_stride > 0: planar output, column j of row t at out[j * plane_stride + t]
    N.check(c, N.lib().scgrhc_decode_fmt16(c, _ptr(d), d.shape[0], d.shape[1], (C.c_int32 * n)(*cols), n,
                                           (C.c_double * n)(*gain), (C.c_double * n)(*baseline), _ptr(out), _stream(dev)))


@torch.library.custom_op('scgrhc::decode_fmt16_records', mutates_args=('out',), device_types='cuda')
def decode_fmt16_records(d: Tensor, rec_row0: Tensor, max_rec_rows: int, cols: Sequence[int], gain: Tensor, baseline: Tensor,
                         recip: bool, out: Tensor, plane_stride: int = 0) -> None:
  """decode_fmt16 for a chunk of records with per-record calibration in ONE launch: ``rec_row0`` (n_rec+1,) int64 device,
  ``gain`` / ``baseline`` (n_rec, len(cols)) fp64 device tables (every WFDB header carries its own pair)."""
  dev = _dev(d)
  _contig(d, torch.int16, 'd'); _contig(out, torch.float64, 'out'); _contig(rec_row0, torch.int64, 'rec_row0')
  _contig(gain, torch.float64, 'gain'); _contig(baseline, torch.float64, 'baseline')
  n, n_rec = len(cols), rec_row0.numel() - 1
  if d.dim() != 2 or out.numel() < d.shape[0] * n or gain.numel() < n_rec * n or baseline.numel() < n_rec * n:
    raise ValueError('decode_fmt16_records: d must be (T, nsig), out (T, len(cols)), tables (n_rec, len(cols))')
  c = ctx(dev)
  with _output_planes(c, plane_stride):
    N.check(c, N.lib().scgrhc_decode_fmt16_records(c, _ptr(d), _ptr(rec_row0), n_rec, int(max_rec_rows), d.shape[1],
                                                   (C.c_int32 * n)(*cols), n, _ptr(gain), _ptr(baseline), 1 if recip else 0,
                                                   _ptr(out), _stream(dev)))


@torch.library.custom_op('scgrhc::waveform_stats', mutates_args=('stats',), device_types='cuda')
def waveform_stats(y: Tensor, min_rhc: float, stats: Tensor) -> None:
  """Per row of y (n_wave, L): R^2 of the OLS line, min, max, below-floor, non-finite, sum
  (waveform_noise.py:29-41 for waveforms of any length)."""
  dev = _dev(y)
  _contig(y, torch.float64, 'y'); _contig(stats, torch.float64, 'stats')
  if y.dim() != 2 or stats.numel() < 6 * y.shape[0]:
    raise ValueError('y must be (n_wave, L) and stats (n_wave, 6)')
  c = ctx(dev)
  N.check(c, N.lib().scgrhc_waveform_stats(c, _ptr(y), y.shape[0], y.shape[1], min_rhc, _ptr(stats), _stream(dev)))


@torch.library.custom_op('scgrhc::synth_records', mutates_args=('out',), device_types='cuda')
def synth_records(out: Tensor, seed: int, rec0: int, n_rec: int, T: int, kinds: Sequence[int], defect_scale: int,
                  grid: int, plane_stride: int = 0) -> None:
  """Synthetic cohort generator, bit-identical to oracle/synth_ref.py."""
  dev = _dev(out)
  _contig(out, torch.float64, 'out')
  nsig = len(kinds)
  if out.numel() < n_rec * T * nsig:
    raise ValueError('out too small')
  arr = (C.c_int32 * nsig)(*kinds)
  c = ctx(dev)
  with _output_planes(c, plane_stride):
    N.check(c, N.lib().scgrhc_synth_records(c, C.c_uint64(seed), rec0, n_rec, T, nsig, arr, defect_scale, grid,
                                            _ptr(out), _stream(dev)))


def check_errors(device_index):
  """Synchronises the current stream; raises ValueError like the reference when a non-finite RHC
  sample reached the regression (waveform_noise.py:32 via sklearn).  Returns the number of candidate windows the
  covered process calls flagged REASON_AMBIGUOUS (R^2 within 1e-9 of 0.8)."""
  c = ctx(device_index)
  bad = C.c_int64(-1)
  rc = N.lib().scgrhc_check_errors(c, _stream(device_index), C.byref(bad))
  if rc == N.ERR_NONFINITE_RHC:
    raise ValueError('Input y contains NaN.')
  N.check(c, rc)
  return int(N.lib().scgrhc_ambiguous_count(c))


def selftest_div(device_index, seed, n, mode):
  counts = torch.zeros(3, dtype=torch.int64, device='cuda:%d' % device_index)
  c = ctx(device_index)
  N.check(c, N.lib().scgrhc_selftest_div(c, C.c_uint64(seed), n, mode, _ptr(counts), _stream(device_index)))
  return counts.cpu().tolist()
