"""Drop-in for the reference's ``recordutil`` (recordutil.py:1-236): same public names and signatures,
window preparation on the GPU.

  reference call                              here
  ------------------------------------------  -----------------------------------------------------------
  get_chamber_intervals(record, chamber)      C planner scgrhc_plan_record (fp64 int(t*500) truncation)
  get_segments(params[, record_name])         keep/reject of every candidate window by the fused CUDA
                                              kernel (predicates only); returns the reference's tuples
  SCGDataset(segments, size, mm_scg, mm_rhc)  min/max + normalise + transpose + fp32 cast on the GPU
  get_global_minmax_vals(segments)            device reduction of per-window pairs
  save_dataloaders(params)                    THE HOT PATH: records -> HBM -> one fused kernel pass ->
                                              device-resident train windows, CPU valid/test windows
  load_dataloader(path)                       unpickle; train windows go back to the GPU

What consumers receive is unchanged: ``waveform_train.run`` iterates the train loader and reads batch[0],
batch[1] ((B,C,750) and (B,1,750) fp32, already CUDA so its ``.to(device)`` is a no-op);
``waveform_test.run`` iterates ``loader.dataset`` and calls ``.numpy()`` on item[1], so valid/test items are
CPU tensors.  Without a CUDA device or without libscgrhc.so every entry point that needs
arithmetic raises — there is no CPU fallback.
"""
import json
import os
import pickle
import sys
from datetime import datetime
from pathlib import Path
from time import time

import numpy as np
import torch
from torch.utils.data import Dataset

try:
  import wfdb
except ImportError:  # not in this image: own format-16 reader (scgrhc/wfdbio.py)
  from scgrhc import wfdbio as wfdb

from paramutil import Params
from pathutil import PROCESSED_DATA_PATH
from timelog import timelog
from waveform_noise import has_noise  # noqa: F401  (re-exported like the reference, recordutil.py:17)

from scgrhc import _native as N
from scgrhc import engine, filters, hostscan, ops

SAMPLE_FREQ = 500


def _device():
  if not torch.cuda.is_available():
    raise RuntimeError('recordutil needs a CUDA device: window preparation runs in libscgrhc (no CPU fallback)')
  return torch.device('cuda', torch.cuda.current_device())


class SCGDataset(Dataset):
  """
  Container dataset class SCG and RHC segments (recordutil.py:22-79).

  ``segments`` is the reference's list of ``(scg (L,C) f64, rhc (L,1) f64, record_name, start_idx,
  stop_idx)`` tuples.  Items are ``(scg f32 (C,L), rhc f32 (1,L), record_name, start_idx, stop_idx,
  (scg_min, scg_max), (rhc_min, rhc_max))`` as in the reference (:65).  Window tensors are stored batched
  (``.scg`` (n,C,L), ``.rhc`` (n,1,L)); ``.segments`` materialises the per-item tuples on first use.
  """

  def __init__(self, segments, segment_size, minmax_scg, minmax_rhc, device='cpu'):
    self.segment_size = int(segment_size * SAMPLE_FREQ)
    self.device = torch.device(device)
    self._segments = None
    self.segments = self.init_segments(segments, minmax_scg, minmax_rhc)

  # -- reference helpers (kept for API parity; the arithmetic they describe runs in the kernel) --------
  def pad(self, tensor):
    """Right-pad the last dim with zeros up to segment_size (recordutil.py:30-39).  The reference's
    '>' branch indexes three dims of a 2-D tensor and raises IndexError; so does this."""
    if tensor.shape[-1] < self.segment_size:
      tensor = torch.nn.functional.pad(tensor, (0, self.segment_size - tensor.shape[-1]))
    elif tensor.shape[-1] > self.segment_size:
      tensor = tensor[:, :, :self.segment_size]
    return tensor

  def minmax_norm(self, tensor, minmax_vals):
    """(x - min) / (max - min + 0.0001) on the GPU in fp64 (recordutil.py:41-47); returns a numpy array."""
    x = np.ascontiguousarray(np.asarray(tensor, dtype=np.float64))
    L = 512                                   # element-wise with a given pair: any window split will do
    n = max(1, -(-x.size // L))
    flat = np.zeros((n * L, 1), dtype=np.float64)
    flat[:x.size, 0] = x.reshape(-1)
    mm = (float(minmax_vals[0]), float(minmax_vals[1]))
    out = _normalise_block(flat, n, mm, mm, torch.float64, L=L)[0]
    return out.cpu().numpy().reshape(-1)[:x.size].reshape(x.shape)

  def invert(self, tensor):
    """Transpose + fp32 cast (recordutil.py:49-53)."""
    return torch.tensor(np.asarray(tensor).T, dtype=torch.float32)

  def init_segments(self, segments, minmax_scg, minmax_rhc):
    """Normalise every segment (recordutil.py:55-66); mutates and returns the passed list like the
    reference.  Segments are grouped by length so that each group is one kernel launch."""
    n = len(segments)
    self._names = [s[2] for s in segments]
    self._start = np.array([int(s[3]) for s in segments], dtype=np.int64)
    self._stop = np.array([int(s[4]) for s in segments], dtype=np.int64)
    self._mm = np.zeros((n, 4), dtype=np.float64)
    C = segments[0][0].shape[1] if n else 0
    scg = torch.zeros((n, C, self.segment_size), dtype=torch.float32)
    rhc = torch.zeros((n, 1, self.segment_size), dtype=torch.float32)
    by_len = {}
    for i, s in enumerate(segments):
      by_len.setdefault(s[0].shape[0], []).append(i)
    for L, idx in by_len.items():
      if L > self.segment_size:
        raise IndexError('too many indices for tensor of dimension 2')   # the reference's pad() on a long segment
      block = np.concatenate([np.concatenate([segments[i][0], segments[i][1]], axis=1) for i in idx]).astype(np.float64)
      s_t, r_t, mm = _normalise_block(block, len(idx), minmax_scg, minmax_rhc, torch.float32, L=L)
      ii = torch.as_tensor(idx)
      scg[ii, :, :L] = s_t.cpu()
      rhc[ii, :, :L] = r_t.cpu()
      self._mm[idx] = mm
    self.scg, self.rhc = scg.to(self.device), rhc.to(self.device)
    for i in range(n):
      segments[i] = self._item(i)
    self._segments = segments
    return segments

  @classmethod
  def from_arrays(cls, scg, rhc, names, start, stop, minmax, segment_size):
    """Hot-path constructor: batched tensors straight from the window kernel, no per-window Python."""
    self = cls.__new__(cls)
    self.segment_size = int(segment_size * SAMPLE_FREQ)
    self.device = scg.device
    self.scg, self.rhc = scg, rhc
    self._names, self._start, self._stop, self._mm = list(names), np.asarray(start), np.asarray(stop), np.asarray(minmax)
    self._segments = None
    return self

  def _item(self, i):
    mm = self._mm[i]
    return (self.scg[i], self.rhc[i], self._names[i], int(self._start[i]), int(self._stop[i]),
            (np.float64(mm[0]), np.float64(mm[1])), (np.float64(mm[2]), np.float64(mm[3])))

  @property
  def segments(self):
    if self._segments is None:
      self._segments = [self._item(i) for i in range(len(self._names))]
    return self._segments

  @segments.setter
  def segments(self, value):
    self._segments = value

  def to(self, device):
    self.device = torch.device(device)
    self.scg, self.rhc = self.scg.to(self.device), self.rhc.to(self.device)
    self._segments = None
    return self

  def collate_meta(self, h):
    """Elements 2..6 of a collated batch for the items ``h`` (numpy indices): names, start, stop, the two min/max pairs."""
    mm = torch.from_numpy(self._mm[h])
    return [tuple(self._names[i] for i in h), torch.from_numpy(self._start[h]), torch.from_numpy(self._stop[h]),
            [mm[:, 0], mm[:, 1]], [mm[:, 2], mm[:, 3]]]

  def collate(self, idx, noise=None):
    """What default_collate makes of items ``idx`` (recordutil.py:198): [scg (B,C,L), rhc (B,1,L), names,
    start (B,), stop (B,), [scg_min (B,), scg_max (B,)], [rhc_min (B,), rhc_max (B,)]]."""
    if self.scg.is_cuda:
      ii = idx.to(self.scg.device, torch.int64).contiguous()
      scg = torch.empty((ii.numel(),) + tuple(self.scg.shape[1:]), dtype=self.scg.dtype, device=self.scg.device)
      rhc = torch.empty((ii.numel(),) + tuple(self.rhc.shape[1:]), dtype=self.rhc.dtype, device=self.rhc.device)
      if self.scg.dtype == torch.float32 and ii.numel():
        sigma = float(noise[0]) if noise is not None and noise[0] > 0 else 0.0             # extension: SCG inputs only
        ops.BatchCollator(self.scg, self.rhc, ii)(0, ii.numel(), scg, rhc, sigma, int(noise[1]) if sigma else 0, int(noise[2]) if sigma else 0)
      elif ii.numel():
        ops.gather_windows(self.scg, ii, scg)
        ops.gather_windows(self.rhc, ii, rhc)
    else:
      scg, rhc = self.scg[idx], self.rhc[idx]
    return [scg, rhc] + self.collate_meta(idx.cpu().numpy())

  def __getstate__(self):
    st = dict(self.__dict__)
    st['scg'], st['rhc'] = self.scg.cpu(), self.rhc.cpu()
    st['_segments'] = None
    st['device'] = str(self.device)
    return st

  def __setstate__(self, st):
    if 'scg' not in st:      # a pickle written by the reference's own SCGDataset: just `segment_size` and `segments`
      segs = st['segments']
      self.segment_size = st['segment_size']
      self.device = torch.device('cpu')
      self._names = [s[2] for s in segs]
      self._start = np.array([int(s[3]) for s in segs], dtype=np.int64)
      self._stop = np.array([int(s[4]) for s in segs], dtype=np.int64)
      self._mm = np.array([[s[5][0], s[5][1], s[6][0], s[6][1]] for s in segs], dtype=np.float64).reshape(len(segs), 4)
      C = segs[0][0].shape[0] if segs else 0
      self.scg = torch.stack([s[0].contiguous() for s in segs]) if segs else torch.zeros((0, C, self.segment_size))
      self.rhc = torch.stack([s[1].contiguous() for s in segs]) if segs else torch.zeros((0, 1, self.segment_size))
      self._segments = segs
      return
    self.__dict__.update(st)
    want = torch.device(st['device'])
    self.device = want if (want.type == 'cpu' or torch.cuda.is_available()) else torch.device('cpu')
    if self.device.type == 'cuda':
      self.device = torch.device('cuda', torch.cuda.current_device())
    self.scg, self.rhc = self.scg.to(self.device), self.rhc.to(self.device)

  def __len__(self):
    """
    Get number of segments.
    """
    return len(self._names)

  def __getitem__(self, index):
    """
    Iterate through segments.
    """
    if self._segments is not None:
      return self._segments[index]
    if index < 0:
      index += len(self)
    if not 0 <= index < len(self):
      raise IndexError(index)
    return self._item(index)


class _Batch(list):
  """A collated batch ``[scg, rhc, names, start, stop, [scg_min, scg_max], [rhc_min, rhc_max]]`` whose metadata elements
  are assembled on first access: the trainer reads elements 0 and 1 only (waveform_train.py:358-359), and the tuple of
  256 names + four small tensors per batch would cost more host time than the gather kernel takes."""
  __slots__ = ('_fill',)

  def __init__(self, scg, rhc, fill):
    super().__init__((scg, rhc, None, None, None, None, None))
    self._fill = fill

  def _materialise(self):
    if self._fill is not None:
      fill, self._fill = self._fill, None
      list.__setitem__(self, slice(2, 7), fill())

  def __getitem__(self, i):
    if self._fill is not None and not (type(i) is int and 0 <= i < 2):
      self._materialise()
    return list.__getitem__(self, i)

  def __iter__(self):
    self._materialise()
    return list.__iter__(self)

  def __reduce__(self):
    self._materialise()
    return (list, (list(list.__iter__(self)),))

  def __repr__(self):
    self._materialise()
    return list.__repr__(self)

  def __eq__(self, other):
    self._materialise()
    return list.__eq__(self, other)

  __hash__ = None


class WindowLoader:
  """The slice of ``torch.utils.data.DataLoader`` the consumers use (``len``, iteration, ``.dataset``,
  ``.batch_size``) with device-side batch assembly: a shuffled index permutation, uploaded once per epoch, + ONE gather
  launch per batch instead of per-item collation (recordutil.py:198-200: shuffle=True, drop_last=False).

  ``reuse_buffers = R`` (default 0 = fresh tensors per batch, like DataLoader): batches are written into a ring of R
  preallocated (batch_size, C, L) / (batch_size, 1, L) buffers, so batch k's tensors are overwritten by batch k + R — for
  train loops that are done with a batch before asking for the R-th next one (the reference's is).

  Noise extension (default off): every batch draws from the Philox stream ``(noise_seed, epoch * len(loader) + batch)``;
  ``epoch`` counts the passes made through THIS object and is not persisted — a trainer that resumes from a checkpoint
  calls ``set_epoch(e)`` (as with DistributedSampler) to continue the noise sequence instead of replaying it."""

  noise_std, noise_seed, epoch, reuse_buffers = 0.0, 0, 0, 0      # defaults for loaders pickled before these existed

  def __init__(self, dataset, batch_size=1, shuffle=False, generator=None, noise_std=0.0, noise_seed=0, reuse_buffers=0):
    self.dataset, self.batch_size, self.shuffle, self.generator = dataset, int(batch_size), shuffle, generator
    # extension (absent from the reference, default off): Gaussian noise on the SCG inputs of every batch, drawn from a
    # counter-based Philox stream (seed, epoch, batch) inside the gather kernel
    self.noise_std, self.noise_seed, self.epoch = float(noise_std), int(noise_seed), 0
    self.reuse_buffers = int(reuse_buffers)

  def set_epoch(self, epoch):
    self.epoch = int(epoch)

  def __len__(self):
    return (len(self.dataset) + self.batch_size - 1) // self.batch_size

  def __getstate__(self):
    st = dict(self.__dict__)
    st.pop('_ring', None)
    return st

  def __iter__(self):
    ds = self.dataset
    n, bs = len(ds), self.batch_size
    order = torch.randperm(n, generator=self.generator) if self.shuffle else torch.arange(n)
    base = self.epoch * len(self)
    self.epoch += 1
    sigma, seed = (self.noise_std, self.noise_seed) if self.noise_std > 0 else (0.0, 0)
    if not (ds.scg.is_cuda and ds.scg.dtype == torch.float32 and n):
      for k, b in enumerate(range(0, n, bs)):
        yield ds.collate(order[b:b + bs], (sigma, seed, base + k) if sigma else None)
      return
    dev = ds.scg.device
    collate = ops.BatchCollator(ds.scg, ds.rhc, order.to(dev))
    tail_s, tail_r = tuple(ds.scg.shape[1:]), tuple(ds.rhc.shape[1:])
    ring = None
    if self.reuse_buffers > 0:
      ring = getattr(self, '_ring', None)
      if ring is None or len(ring) != self.reuse_buffers or ring[0][0].shape != (bs,) + tail_s or ring[0][0].device != dev:
        ring = self._ring = [(torch.empty((bs,) + tail_s, dtype=torch.float32, device=dev),
                              torch.empty((bs,) + tail_r, dtype=torch.float32, device=dev)) for _ in range(self.reuse_buffers)]
    order_np = order.numpy()
    for k, b in enumerate(range(0, n, bs)):
      m = min(bs, n - b)
      if ring is not None:
        scg, rhc = ring[k % len(ring)]
        scg, rhc = scg[:m], rhc[:m]
      else:
        scg = torch.empty((m,) + tail_s, dtype=torch.float32, device=dev)
        rhc = torch.empty((m,) + tail_r, dtype=torch.float32, device=dev)
      collate(b, m, scg, rhc, sigma, seed, base + k)
      yield _Batch(scg, rhc, lambda h=order_np[b:b + m]: ds.collate_meta(h))


def _normalise_block(block, n, minmax_scg, minmax_rhc, out_dtype, L=None):
  """``block``: (n*L, C+1) fp64, SCG columns then RHC.  Local pairs where a ``minmax_*`` is None, the
  given pair otherwise (recordutil.py:58-59).  Returns (scg (n,C,L), rhc (n,1,L), minmax (n,4) numpy)."""
  dev = _device()
  Cn = block.shape[1] - 1
  if Cn == 0:                         # minmax_norm() helper: a bare column, treated as its own "RHC"
    block = np.concatenate([block, block], axis=1)
    Cn = 1
  L = L if L is not None else block.shape[0] // n
  arena = torch.from_numpy(np.ascontiguousarray(block)).to(dev)
  plan = engine.Plan(np.array([(0, 0, n, 0)], dtype=engine.INTERVAL_DTYPE), n, L)
  st = engine.prepare_windows(arena, plan, list(range(Cn)), Cn, float('-inf'), predicates_only=True, keep_all=True,
                              check=False)
  mm = st.minmax.clone()
  if minmax_scg is not None:
    mm[:, 0], mm[:, 1] = float(minmax_scg[0]), float(minmax_scg[1])
  if minmax_rhc is not None:
    mm[:, 2], mm[:, 3] = float(minmax_rhc[0]), float(minmax_rhc[1])
  scg = torch.empty((n, Cn, L), dtype=out_dtype, device=dev)
  rhc = torch.empty((n, 1, L), dtype=out_dtype, device=dev)
  flags = N.USE_KEPT_LIST | (N.OUT_F64 if out_dtype == torch.float64 else 0)
  ops.process_windows(arena, plan.device_intervals(dev), n, L, 0, list(range(Cn)), Cn, float('-inf'), 1e-3, flags,
                      [0.0] * 4, torch.arange(n, device=dev), n, scg, rhc, mm, None, None, None, None)
  return scg, rhc, mm.cpu().numpy()


def get_record_names():
  """
  Get record names in a given directory (recordutil.py:82-90; order is arbitrary there, sorted here).
  """
  names = set()
  for filename in os.listdir(PROCESSED_DATA_PATH):
    if filename.endswith('.dat') or filename.endswith('.hea'):
      names.add(Path(filename).stem)
  return sorted(names)


def _read_meta(record_name):
  with open(os.path.join(PROCESSED_DATA_PATH, f'{record_name}.json'), 'rb') as f:
    return json.loads(f.read())


def get_chamber_intervals(record_name, chamber):
  """
  Get sample intervals for when cath was in a particular chamber (recordutil.py:93-110).
  """
  _, _, bounds = engine.plan_record(_read_meta(record_name), chamber, 0, 1)
  return [(int(a), int(b)) for a, b in bounds]


def get_channels(record, channel_names, start_idx, stop_idx):
  """
  Get specific channels from record by channel name (recordutil.py:113-119).  A host-side copy; raises
  ValueError for a missing channel exactly like ``list.index``.
  """
  indexes = [record.sig_name.index(name) for name in channel_names]
  return record.p_signal[start_idx:stop_idx, indexes]


def _scan_record(params, record_name):
  """Keep flags of every candidate window of one record, decided on the GPU."""
  record = wfdb.rdrecord(os.path.join(PROCESSED_DATA_PATH, record_name))
  W = int(params.segment_size * SAMPLE_FREQ)
  cols, rcol = engine.resolve_columns(record.sig_name, params.in_channels)
  p = np.ascontiguousarray(record.p_signal[:, cols + [rcol]], dtype=np.float64)
  plan = engine.plan_cohort([_read_meta(record_name)], params.chamber, [p.shape[0]], W)
  arena = torch.from_numpy(p).to(_device())
  st = engine.prepare_windows(arena, plan, list(range(len(cols))), len(cols), params.min_RHC, predicates_only=True)
  return record, plan, st.keep.cpu().numpy().astype(bool)


def get_segments(params, record_name=None):
  """
  Get segments of a given size with the specified SCG channels (recordutil.py:122-149): the reference's
  list of (scg (W,C) view, rhc (W,1) view, record_name, start_idx, stop_idx) for windows without noise.
  """
  if record_name is None:
    segments = []
    for record_name in get_record_names():
      segments.extend(get_segments(params, record_name=record_name))
    return segments
  record, plan, keep = _scan_record(params, record_name)
  W = plan.W
  segments, cand = [], 0
  for interval in get_chamber_intervals(record_name, params.chamber):
    scg_signal = get_channels(record, params.in_channels, interval[0], interval[1])
    rhc_signal = get_channels(record, ['RHC_pressure'], interval[0], interval[1])
    for i in range(scg_signal.shape[0] // W):
      if keep[cand]:
        segments.append((scg_signal[i * W:(i + 1) * W], rhc_signal[i * W:(i + 1) * W], record_name, i * W, i * W + W))
      cand += 1
  return segments


def get_global_minmax_vals(segments):
  """
  Get min and max values for normalization (recordutil.py:152-169): ((scg_min, scg_max), (rhc_min, rhc_max))
  over all given segments, reduced on the GPU.
  """
  if len(segments) == 0:
    return (None, None), (None, None)
  dev = _device()
  by_len = {}
  for s in segments:
    by_len.setdefault(s[0].shape[0], []).append(s)
  parts = []
  for L, group in by_len.items():
    block = np.concatenate([np.concatenate([s[0], s[1]], axis=1) for s in group]).astype(np.float64)
    Cn = block.shape[1] - 1
    arena = torch.from_numpy(np.ascontiguousarray(block)).to(dev)
    plan = engine.Plan(np.array([(0, 0, len(group), 0)], dtype=engine.INTERVAL_DTYPE), len(group), L)
    st = engine.prepare_windows(arena, plan, list(range(Cn)), Cn, float('-inf'), predicates_only=True, keep_all=True,
                                check=False, use_global_min_max=True)
    parts.append(st.global_minmax.cpu().numpy())
  g = np.stack(parts)
  return ((np.float64(g[:, 0].min()), np.float64(g[:, 1].max())), (np.float64(g[:, 2].min()), np.float64(g[:, 3].max())))


def train_valid_test_split(n, seed=None):
  """Index split with sklearn's ``train_test_split(train_size=0.9)`` then ``(train_size=0.5)`` arithmetic
  (recordutil.py:191-192): n_train = floor(0.9 n), the rest is split floor(0.5 m) / remainder; indices come
  from one permutation per call, test part first, exactly as ShuffleSplit draws them.  ``seed=None`` uses
  numpy's global RandomState like the reference (unseeded there)."""
  rng = np.random.mtrand._rand if seed is None else np.random.RandomState(seed)

  def split(idx, frac):
    m = len(idx)
    n_train = int(np.floor(frac * m))
    n_test = m - n_train
    if n_train == 0:                  # sklearn raises for an empty cohort too: the reference then writes no loaders
      raise ValueError('With n_samples=%d, test_size=None and train_size=%s, the resulting train set will be empty. '
                       'Adjust any of the aforementioned parameters.' % (m, frac))
    perm = rng.permutation(m)
    return idx[perm[n_test:n_test + n_train]], idx[perm[:n_test]]

  train, rest = split(np.arange(n), 0.9)
  valid, test = split(rest, 0.5)
  return train, valid, test


def _stage_spec(params, C, rows, metas, names, rec0):
  """Optional per-record stages (band-pass, resample; absent from the reference, default off) as HostIngest's ``stages``
  dict, plus the plan of the cohort AFTER them."""
  W = int(params.segment_size * SAMPLE_FREQ)
  stride_s = getattr(params, 'segment_stride', None)
  plan = engine.plan_cohort(metas, params.chamber, rows, W, names, stride=int(stride_s * SAMPLE_FREQ) if stride_s else 0, rec0=rec0)
  sos = _bandpass_sos(params)
  rate = getattr(params, 'resample_rate', None)
  stages = None
  if sos is not None or (rate and int(rate) != SAMPLE_FREQ):
    # they run per chunk inside the same streamed ingest: copy of chunk k+1 overlaps decode / band-pass / resample /
    # window kernel of chunk k, and the cohort never has to be resident
    stages = {}
    if sos is not None:                    # zero-phase band-pass of the SCG channels (scipy sosfiltfilt semantics)
      stages.update(sos=sos, filter_cols=list(range(C)), filter_exact=getattr(params, 'bandpass_mode', None) != 'scan')
    if rate and int(rate) != SAMPLE_FREQ:  # every channel to the model rate (scipy resample_poly semantics)
      import math
      g = math.gcd(int(rate), SAMPLE_FREQ)
      up, down = int(rate) // g, SAMPLE_FREQ // g
      out_rows = [-(-int(n) * up // down) for n in rows]
      stages.update(resample=(int(rate), SAMPLE_FREQ), out_rows=out_rows,
                    resample_exact=getattr(params, 'resample_mode', None) != 'fused')
      W = int(params.segment_size * int(rate))
      plan = engine.plan_cohort(metas, params.chamber, out_rows, W, names, stride=int(stride_s * int(rate)) if stride_s else 0,
                                fs=float(int(rate)), rec0=rec0)
  return plan, stages


class _Heterogeneous(Exception):
  """The records of a cohort do not share one frame layout: the streamed ingest hands over to the eager one."""


def _prepare_streamed(params, names, rec0, C, dev, chunk_records, group):
  """Format-16 files through ``engine.LazyDiskIngest``: only the file sizes are looked at before the first chunk moves;
  side-cars and headers are parsed, and intervals planned, chunk by chunk while earlier chunks are being read / copied /
  processed.  Returns None when the cohort does not qualify (other formats, records with different frame layouts)."""
  W = int(params.segment_size * SAMPLE_FREQ)
  try:
    first = wfdb.read_header(os.path.join(PROCESSED_DATA_PATH, names[0]))
    cols0, rcol0 = engine.resolve_columns(first[0], params.in_channels)
    nsig_file = len(first[0])
    sizes = [os.path.getsize(os.path.join(PROCESSED_DATA_PATH, n + '.dat')) // (2 * nsig_file) for n in names]
  except (NotImplementedError, OSError):
    return None
  sel = cols0 + [rcol0]

  native = getattr(wfdb, 'NATIVE_SCAN', False)

  def parse_chunk(r0, r1):
    if native:                    # headers + side-cars of the chunk in one native call (scgrhc.hostscan)
      try:
        rows, g, b, same, metas = hostscan.scan(PROCESSED_DATA_PATH, names[r0:r1], first[0], wfdb.read_header, _read_meta)
      except hostscan.Unscannable:
        rows = None
      if rows is not None:
        if not same.all():
          raise _Heterogeneous(names[r0 + int(np.argmin(same))])
        return ([os.path.join(PROCESSED_DATA_PATH, n + '.dat') for n in names[r0:r1]], [int(v) for v in rows],
                np.ascontiguousarray(g[:, sel]), np.ascontiguousarray(b[:, sel]), metas)
    paths, rows, gains, bases, metas = [], [], [], [], []
    for name in names[r0:r1]:
      h = wfdb.read_header(os.path.join(PROCESSED_DATA_PATH, name))
      if len(h[0]) != nsig_file or engine.resolve_columns(h[0], params.in_channels) != (cols0, rcol0):
        raise _Heterogeneous(name)
      paths.append(h[5]); rows.append(h[2])
      gains.append([float(h[3][j]) for j in sel]); bases.append([float(h[4][j]) for j in sel])
      metas.append(_read_meta(name))
    return paths, rows, gains, bases, metas

  def plan_chunk(metas, rows, r0):
    return engine.plan_cohort(metas, params.chamber, rows, W, rec0=rec0 + r0)

  ing = engine.LazyDiskIngest(sizes, C + 1, dev, W, nsig_file, parse_chunk, plan_chunk, chunk_records=chunk_records or 32)
  try:
    return ing.run_files(list(range(C)), C, params.min_RHC, sel, use_global_min_max=bool(params.use_global_min_max), group=group,
                         normalisation=getattr(params, 'normalisation', None))
  except (_Heterogeneous, NotImplementedError):
    return None


def _prepare_eager(params, names, names_all, lo, C, dev, chunk_records, group):
  """Every side-car and header (or, for readers without digital access, every record) is read first, then one plan for the
  whole shard and a chunked ingest: the path of the optional stages and of readers other than scgrhc.wfdbio."""
  plan = stages = source = decode = host = metas = None
  rows = []
  if getattr(wfdb, 'NATIVE_SCAN', False) and names:
    try:                          # headers + side-cars of the shard in one native call (scgrhc.hostscan)
      first = wfdb.read_header(os.path.join(PROCESSED_DATA_PATH, names[0]))
      cols0, rcol0 = engine.resolve_columns(first[0], params.in_channels)
      nrows, g, b, same, scanned = hostscan.scan(PROCESSED_DATA_PATH, names, first[0], wfdb.read_header, _read_meta)
      if same.all():
        sel = cols0 + [rcol0]
        rows, metas = [int(v) for v in nrows], scanned
        source = engine.DiskSource([os.path.join(PROCESSED_DATA_PATH, n + '.dat') for n in names], rows, len(first[0]))
        decode = (sel, np.ascontiguousarray(g[:, sel]), np.ascontiguousarray(b[:, sel]))
    except (NotImplementedError, hostscan.Unscannable):
      source, rows, metas = None, [], None
  if metas is None:
    metas = [_read_meta(name) for name in names]
  if source is None and hasattr(wfdb, 'read_header'):
    try:
      heads = [wfdb.read_header(os.path.join(PROCESSED_DATA_PATH, name)) for name in names]
      sels = [engine.resolve_columns(h[0], params.in_channels) for h in heads]
      if heads and all(len(h[0]) == len(heads[0][0]) and s == sels[0] for h, s in zip(heads, sels)):
        rows = [h[2] for h in heads]
        sel = sels[0][0] + [sels[0][1]]
        source = engine.DiskSource([h[5] for h in heads], rows, len(heads[0][0]))
        decode = (sel, [[float(h[3][j]) for j in sel] for h in heads], [[float(h[4][j]) for j in sel] for h in heads])
    except NotImplementedError:
      source = None
  if source is None:
    blocks, digital = [], True
    for name in names:
      record = wfdb.rdrecord(os.path.join(PROCESSED_DATA_PATH, name))
      cols, rcol = engine.resolve_columns(record.sig_name, params.in_channels)
      digital = digital and getattr(record, 'd_signal', None) is not None and getattr(record, 'adc_gain', None) is not None \
          and getattr(record, 'baseline', None) is not None and record.d_signal.dtype == np.int16
      blocks.append((record, cols + [rcol]))
      rows.append(record.d_signal.shape[0] if digital else record.p_signal.shape[0])
    host = torch.empty((int(sum(rows)), C + 1), dtype=torch.int16 if digital else torch.float64, pin_memory=True)
    at, gains, bases = 0, [], []
    for (record, sel), n in zip(blocks, rows):
      if digital:
        host[at:at + n] = torch.from_numpy(np.ascontiguousarray(record.d_signal[:, sel]))
        gains.append([float(record.adc_gain[j]) for j in sel])
        bases.append([float(record.baseline[j]) for j in sel])
      else:
        host[at:at + n] = torch.from_numpy(np.ascontiguousarray(record.p_signal[:, sel], dtype=np.float64))
      at += n
    source, digital_nsig = host, (C + 1) if digital else None
    decode = (list(range(C + 1)), gains, bases) if digital else None
  else:
    digital_nsig = source.nsig
  plan, stages = _stage_spec(params, C, rows, metas, names_all, lo)
  if chunk_records is None:
    # the time-parallel band-pass runs one CTA per record (8 per SM): a chunk of a few hundred records costs what 32 do
    chunk_records = 256 if (stages and stages.get('sos') is not None) else 32
  ing = engine.HostIngest(plan, rows, C + 1, dev, chunk_records=chunk_records, digital_nsig=digital_nsig, stages=stages)
  store = ing.run(source, list(range(C)), C, params.min_RHC, decode=decode, use_global_min_max=bool(params.use_global_min_max),
                  group=group, normalisation=getattr(params, 'normalisation', None))
  return store


def prepare_cohort(params, record_names=None, chunk_records=None, group=None):
  """Records -> HBM -> fused window kernel.  Returns (WindowStore, record_names): every kept window of this rank's
  records, device resident, in the reference's order (records in ``record_names`` order).

  Multi-GPU (one process per GPU under torchrun): the per-record loop of get_segments
  (recordutil.py:131-132) shards by contiguous blocks of ``record_names`` (``engine.shard_records``); every rank runs
  the same pipeline on its block, ``use_global_min_max`` (recordutil.py:152-169,185-189) goes through ONE MIN all-reduce
  of {min, -max}, and one all-gather of the kept counts gives ``store.shard`` (this rank's offset in the cohort-wide
  ordered list), so the concatenation of the ranks' stores in rank order IS the single-GPU store.  ``rec_id`` indexes
  the full ``record_names`` list on every rank.

  Ingest: the SIGNAL DATA of format-16 records (scgrhc.wfdbio) is always streamed: a reader pool fills a ring of pinned
  chunk buffers straight from the ``.dat`` files (``engine.DiskSource``) while the previous chunk crosses PCIe and the
  one before is in the window kernel; the int16 frames travel as they are (4x fewer bytes than fp64) and
  ``(d - baseline) / gain`` runs on the device with the per-record calibration in a device table (what wfdb.rdrecord does
  on the host at recordutil.py:137).  Host memory holds a few chunks whatever the cohort size.  The per-record METADATA
  (side-car JSON, header, interval maths: ~60 us of Python per record) is parsed up front for shards below
  ``SCGRHC_LAZY_MIN_RECORDS`` (4096) records and chunk by chunk, overlapped with the stream, above
  (``engine.LazyDiskIngest``).  Readers without digital access (the real ``wfdb`` package, other formats) upload
  ``p_signal`` of every record of the shard from one pinned buffer."""
  from scgrhc import dist as sdist
  names_all = list(record_names) if record_names is not None else get_record_names()
  rank, world = sdist.current(group)
  lo, hi = engine.shard_records(len(names_all), rank, world)
  names = names_all[lo:hi]
  C = len(params.in_channels)
  dev = _device()
  store = None
  plain = not (_bandpass_sos(params) is not None or getattr(params, 'resample_rate', None) or getattr(params, 'segment_stride', None))
  # parse-as-you-go pays off on large shards (100k records: ~8 s of side-car / header parsing hidden behind the stream);
  # on small ones the host parse of the whole shard is a few ms and the simpler eager ingest is just as fast (measured)
  lazy_from = int(os.environ.get('SCGRHC_LAZY_MIN_RECORDS', '4096'))
  if hasattr(wfdb, 'read_header') and plain and len(names) >= lazy_from:
    store = _prepare_streamed(params, names, lo, C, dev, chunk_records, group)
  if store is None:
    store = _prepare_eager(params, names, names_all, lo, C, dev, chunk_records, group)
  store.shard = sdist.exchange_counts(store.n_kept, dev, group)
  return store, names_all


def _bandpass_sos(params):
  """Optional params keys (absent from all shipped params.json): ``bandpass_sos`` = explicit second-order sections, or
  ``bandpass`` = [low_hz, high_hz] (+ ``bandpass_order``, default 4) designed as a Butterworth band-pass at 500 Hz."""
  sos = getattr(params, 'bandpass_sos', None)
  if sos is not None:
    return np.asarray(sos, dtype=np.float64)
  band = getattr(params, 'bandpass', None)
  if band:
    return filters.butter_sos(float(band[0]), float(band[1]), SAMPLE_FREQ, int(getattr(params, 'bandpass_order', None) or 4))
  return None


def save_dataloaders(params):
  """
  Get training and test segments, then save as loader objects (recordutil.py:172-216).  Under torchrun every rank
  prepares its block of records and rank 0 writes the reference's three pickles (see ``_write_loaders``).
  """
  from scgrhc import dist as sdist
  _check_no_loaders(params)
  sdist.barrier()                        # every rank has looked before any rank writes
  store, names = prepare_cohort(params)
  _write_loaders(params, store, names)


def _check_no_loaders(params):
  """The reference's guards, same messages (recordutil.py:176-181; waveform_pipeline.py:12-15 relies on them)."""
  if os.path.exists(params.train_path):
    raise Exception('Train file already exists!')
  elif os.path.exists(params.valid_path):
    raise Exception('Valid file already exists!')
  elif os.path.exists(params.test_path):
    raise Exception('Test file already exists!')


class ShardedTrainLoader:
  """What ``train_path`` holds when the train windows of a multi-GPU job stay on their ranks (``train_layout: "sharded"``,
  or automatically when they exceed ``SCGRHC_GATHER_LIMIT_GB``): rank r's ``WindowLoader`` is pickled next to it as
  ``<train_path>.rank<r>-of-<N>``.  ``load_dataloader(train_path)`` resolves it: under torchrun with the same world size
  every rank gets its own shard (data-parallel training); a single process gets all shards concatenated."""

  def __init__(self, world, counts, batch_size):
    self.world, self.counts, self.batch_size = int(world), list(counts), int(batch_size)

  @staticmethod
  def shard_path(train_path, rank, world):
    return '%s.rank%d-of-%d' % (train_path, rank, world)

  def resolve(self, train_path):
    from scgrhc import dist as sdist
    rank, world = sdist.current()
    if world == self.world and world > 1:
      with open(self.shard_path(train_path, rank, world), 'rb') as f:
        return pickle.load(f)
    parts = []
    for r in range(self.world):
      with open(self.shard_path(train_path, r, self.world), 'rb') as f:
        parts.append(pickle.load(f))
    ds = [p.dataset for p in parts]
    dev = ds[0].scg.device
    merged = SCGDataset.from_arrays(torch.cat([d.scg.to(dev) for d in ds]), torch.cat([d.rhc.to(dev) for d in ds]),
                                    sum((d._names for d in ds), []), np.concatenate([d._start for d in ds]),
                                    np.concatenate([d._stop for d in ds]), np.concatenate([d._mm for d in ds]), 1.0)
    merged.segment_size = ds[0].segment_size
    first = parts[0]
    return WindowLoader(merged, batch_size=first.batch_size, shuffle=first.shuffle, noise_std=first.noise_std,
                        noise_seed=first.noise_seed)


def _write_loaders(params, store, names, shard=None):
  """Split 90/5/5, build the three loaders, pickle them and write record_log.txt (recordutil.py:191-216).

  Multi-GPU: ``store`` holds this rank's kept windows, ``shard`` (default ``store.shard``) its place in the cohort-wide
  ordered list.  The split is ONE draw over the global list (the same on every rank); each rank gathers its members of
  the three subsets on its own GPU; valid / test windows (5 % each) travel to rank 0, which writes the reference's
  pickles — identical to a single-GPU run with the same ``split_seed``.  Train windows are gathered to rank 0 as well
  (``train_layout: "gathered"``, the drop-in default: the reference's trainer is one process) unless they exceed
  ``SCGRHC_GATHER_LIMIT_GB`` (default 16) or ``train_layout: "sharded"`` asks for per-rank loaders (``ShardedTrainLoader``)."""
  from scgrhc import dist as sdist
  sh = shard if shard is not None else getattr(store, 'shard', None)
  if sh is None:
    sh = sdist.Shard(0, 1, None, 0, store.n_kept, (store.n_kept,))
  dev = store.kept_idx.device
  n_all = sh.total
  n_amb = store.n_ambiguous
  if n_amb:
    import warnings
    warnings.warn('%d candidate window(s) have an R^2 within 1e-9 of the 0.8 straight-line threshold (waveform_noise.py:34); '
                  'they were decided by the closed-form R^2, which sklearn\'s lstsq-based score could round the other way'
                  % n_amb)
  seed = getattr(params, 'split_seed', None)
  if sh.active and seed is None:        # unseeded like the reference, but one draw for the whole job: rank 0's
    seed = int(sdist.broadcast_index(torch.tensor([np.random.randint(0, 2 ** 31 - 1)]), dev, sh.group)[0])
  train_idx, valid_idx, test_idx = train_valid_test_split(n_all, seed)

  rec_id = store.rec_id.cpu().numpy()
  start, stop = store.start_idx.cpu().numpy(), store.stop_idx.cpu().numpy()
  mm = store.kept_minmax().cpu().numpy()
  C, W = (store.scg.shape[1], store.scg.shape[2]) if store.scg is not None else (len(params.in_channels), int(params.segment_size * SAMPLE_FREQ))
  bounds = np.concatenate([[0], np.cumsum(sh.counts)])

  def make(idx, device, gather=True):
    """SCGDataset over the global list positions ``idx`` (in that order).  Single GPU: a device gather.  Multi-GPU: every
    rank gathers its members; with ``gather`` they are sent to rank 0 (None elsewhere), else each rank keeps its own."""
    owner = np.searchsorted(bounds, idx, side='right') - 1            # rank holding each member
    mine = owner == sh.rank
    local = (idx[mine] - sh.offset).astype(np.int64)
    pos = torch.as_tensor(local, dtype=torch.int64, device=dev)
    scg, rhc = store.gather(pos) if store.scg is not None else (torch.empty((0, C, W), device=dev), torch.empty((0, 1, W), device=dev))
    meta = np.concatenate([rec_id[local, None].astype(np.float64), start[local, None].astype(np.float64),
                           stop[local, None].astype(np.float64), mm[local].reshape(-1, 4)], axis=1)
    if sh.active and gather:
      counts = [int((owner == r).sum()) for r in range(sh.world)]
      scg = sdist.gather_rows(scg, counts, sh.group)
      rhc = sdist.gather_rows(rhc, counts, sh.group)
      meta_t = sdist.gather_rows(torch.from_numpy(meta).to(dev), counts, sh.group)
      if sh.rank != 0:
        return None
      # rows arrived rank by rank; put them back in the order of ``idx``
      back = torch.as_tensor(np.argsort(np.argsort(owner, kind='stable'), kind='stable'), dtype=torch.int64, device=dev)
      scg, rhc, meta = scg[back], rhc[back], meta_t[back].cpu().numpy()
    return SCGDataset.from_arrays(scg.to(device), rhc.to(device), [names[int(r)] for r in meta[:, 0]], meta[:, 1].astype(np.int64),
                                  meta[:, 2].astype(np.int64), meta[:, 3:7], params.segment_size)

  layout = getattr(params, 'train_layout', None)
  if sh.active and layout is None:
    limit = float(os.environ.get('SCGRHC_GATHER_LIMIT_GB', '16')) * 1e9
    layout = 'sharded' if len(train_idx) * (C + 1) * W * 4 > limit else 'gathered'
  sharded = sh.active and layout == 'sharded'

  noise = dict(noise_std=getattr(params, 'noise_std', None) or 0.0, noise_seed=getattr(params, 'noise_seed', None) or 0)
  train_set = make(train_idx, dev, gather=not sharded)   # stays in HBM for waveform_train
  valid_set = make(valid_idx, 'cpu')                     # waveform_test calls .numpy() on items
  test_set = make(test_idx, 'cpu')

  if sharded:
    with open(ShardedTrainLoader.shard_path(params.train_path, sh.rank, sh.world), 'wb') as f:
      pickle.dump(WindowLoader(train_set, batch_size=params.batch_size, shuffle=True, **noise), f)
  if sh.rank == 0:
    if sharded:
      owner = np.searchsorted(bounds, train_idx, side='right') - 1
      train_loader = ShardedTrainLoader(sh.world, [int((owner == r).sum()) for r in range(sh.world)], params.batch_size)
    else:
      train_loader = WindowLoader(train_set, batch_size=params.batch_size, shuffle=True, **noise)
    valid_loader = WindowLoader(valid_set, batch_size=1, shuffle=True)
    test_loader = WindowLoader(test_set, batch_size=1, shuffle=True)

    with open(params.train_path, 'wb') as f:
      pickle.dump(train_loader, f)

    with open(params.valid_path, 'wb') as f:
      pickle.dump(valid_loader, f)

    with open(params.test_path, 'wb') as f:
      pickle.dump(test_loader, f)

    with open(os.path.join(params.dir_path, 'record_log.txt'), 'w') as f:
      f.write(f'Dataset created: {datetime.now()}\n')
      f.write(f'All segments: {n_all}\n')
      f.write(f'Valid segments: {len(valid_idx)}\n')
      f.write(f'Train segments: {len(train_idx)}\n')
      f.write(f'Test segments: {len(test_idx)}\n')
  sdist.barrier(sh.group)                # the files exist when any rank returns


def save_dataloaders_sweep(params_list, record_names=None, group=None):
  """Extension (BASELINE configs[4]; the reference runs its `all` sweep as independent jobs, waveform_pipeline.py:33-37):
  data preparation for SEVERAL experiment configs over one cohort.  Every record is read and uploaded once (the union
  of the configs' channels + RHC); configs that differ only in their channel subset share one predicate pass and one
  fan-out pass (scgrhc.sweep.iter_sweep); each config then gets its three pickled loaders and record_log.txt exactly as
  save_dataloaders writes them.  Configs whose loaders already exist are reported and skipped, like the reference's
  guard; configs with optional-stage keys (band-pass, resample, stride, z-score) take the single-config path.
  Under torchrun the records shard across the ranks exactly as for one config (``prepare_cohort``); every config's
  loaders are assembled on rank 0.  Returns {dir_path: n_kept}."""
  from scgrhc import dist as sdist
  from scgrhc import sweep
  todo, single, out = [], [], {}
  for params in params_list:
    try:
      _check_no_loaders(params)
    except Exception as e:
      print(e)
      continue
    ext = any(getattr(params, k, None) for k in ('bandpass', 'bandpass_sos', 'resample_rate', 'segment_stride')) or \
        getattr(params, 'normalisation', None) not in (None, 'minmax')
    (single if ext else todo).append(params)
  sdist.barrier(group)                   # every rank has looked before any rank writes
  for params in single:
    save_dataloaders(params)
    out[params.dir_path] = None
  if not todo:
    return out
  names_all = list(record_names) if record_names is not None else get_record_names()
  rank, world = sdist.current(group)
  lo, hi = engine.shard_records(len(names_all), rank, world)
  wanted = []
  for params in todo:
    for ch in params.in_channels:
      if ch not in wanted:
        wanted.append(ch)
  records, metas, rows, sig = [], [], [], None
  for name in names_all[lo:hi]:
    record = wfdb.rdrecord(os.path.join(PROCESSED_DATA_PATH, name))
    order = sorted(wanted, key=record.sig_name.index)          # ValueError for a missing channel, as list.index does
    cols, rcol = engine.resolve_columns(record.sig_name, order)
    this_sig = order + [engine.RHC_NAME]
    if sig is None:
      sig = this_sig
    elif sig != this_sig:
      raise ValueError('records of one sweep must list the selected signals in the same order (%s vs %s)' % (sig, this_sig))
    records.append(np.ascontiguousarray(record.p_signal[:, cols + [rcol]], dtype=np.float64))
    metas.append(_read_meta(name))
    rows.append(records[-1].shape[0])
  if sig is None:                        # a rank without records still takes part in every collective
    sig = list(wanted) + [engine.RHC_NAME]
  dev = _device()
  host = torch.empty((int(sum(rows)), len(sig)), dtype=torch.float64, pin_memory=True)
  at = 0
  for block in records:
    host[at:at + len(block)] = torch.from_numpy(block)
    at += len(block)
  arena = host.to(dev, non_blocking=True)
  configs = {str(i): p for i, p in enumerate(todo)}
  for key, store in sweep.iter_sweep(arena, sig, metas, rows, configs, buffers={}, group=group, rec0=lo):
    params = configs[key]
    shard = sdist.exchange_counts(store.n_kept, dev, group)
    _write_loaders(params, store, names_all, shard)
    out[params.dir_path] = shard.total
  return out


def load_dataloader(path):
  """
  Load prior loader object (recordutil.py:219-224).
  """
  with open(path, 'rb') as f:
    loader = pickle.load(f)
  if isinstance(loader, ShardedTrainLoader):      # a multi-GPU job left the train windows on their ranks
    loader = loader.resolve(path)
  return loader


def run(params):
  start_time = time()
  print(timelog(f'Run recordutil for {params.dir_path}', start_time))
  save_dataloaders(params)


def prepare_all(dir_names):
  """Extension (BASELINE configs[4]): data preparation only, for several experiment directories in one job —
  ``python recordutil.py prepare waveform_06 waveform_07 ...``.  The cohort is read and uploaded once; configs that
  differ only in their channel subset share one predicate pass and one fan-out pass (save_dataloaders_sweep).  The
  reference's own ``waveform_pipeline.py`` (used as is) then finds the loaders in place and goes
  straight to training (its "already exists" branch, waveform_pipeline.py:12-15)."""
  return save_dataloaders_sweep([Params(os.path.join(d, 'params.json')) for d in dir_names])


if __name__ == '__main__':
  if sys.argv[1] == 'prepare':
    prepare_all(sys.argv[2:])
  else:
    dir_path = sys.argv[1]
    params = Params(os.path.join(dir_path, 'params.json'))
    run(params)
