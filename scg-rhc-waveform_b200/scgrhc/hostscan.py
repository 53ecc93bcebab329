"""Headers and side-cars of many records in one native call (csrc/host_scan.h, `scgrhc_scan_records`).

The reference opens `<name>.hea` (inside wfdb.rdrecord, recordutil.py:137) and `<name>.json` (recordutil.py:97-98) one
record at a time; in Python that is ~55 us per record — more than the record's frames take to cross PCIe.  The native
scanner parses the common shape of both files on a few threads and flags every record it does not recognise; those go
through the general Python parsers (``wfdbio.read_header``, ``json``) here, so the result is the same either way
(``tests/test_hostscan.py`` compares the two on generated cohorts, odd files included).  Host-only: no device is touched.
"""
import ctypes as C
import os

import numpy as np

from . import _native as N
from . import engine

MAX_EVENTS = 64
PREFIX_BYTES = 16


class Unscannable(Exception):
  """A record whose side-car does not fit the fixed-size tables (more than MAX_EVENTS events, a chamber prefix of
  PREFIX_BYTES characters or more, a key literally named END): the caller takes its general path for the chunk."""


class ScannedMetas:
  """Stands in for the list of side-car dicts wherever a cohort is planned: ``len()`` and a ready ``EventTabs``."""

  def __init__(self, n, tabs):
    self.n, self.tabs = int(n), tabs

  def __len__(self):
    return self.n


def _blob(strings):
  return b''.join(s.encode() + b'\0' for s in strings)


def tabs_from_arrays(n_events, duration, ev_time, ev_prefix):
  """``engine.EventTabs`` of a chunk from the scanner's fixed-size tables: per record its events in file order plus the
  appended END = record length in seconds (recordutil.py:100-104); records with no event (or no event object) own none."""
  n, M = ev_time.shape
  n_ev = np.asarray(n_events, dtype=np.int64)
  valid = n_ev >= 1
  cnt = np.where(valid, n_ev + 1, 0)
  ext_t = np.zeros((n, M + 1), dtype=np.float64)
  ext_t[:, :M] = ev_time
  ext_p = np.zeros((n, M + 1), dtype='S%d' % PREFIX_BYTES)
  ext_p[:, :M] = ev_prefix
  rows = np.nonzero(valid)[0]
  ext_t[rows, n_ev[rows]] = np.asarray(duration, dtype=np.float64)[rows]
  ext_p[rows, n_ev[rows]] = b'END'
  pos = np.arange(M + 1)[None, :]
  mask = pos < cnt[:, None]
  tabs = engine.EventTabs.__new__(engine.EventTabs)
  tabs.n_rec = n
  tabs.times = np.ascontiguousarray(ext_t[mask])
  tabs.prefix = ext_p[mask]
  tabs.is_end = (pos == n_ev[:, None])[mask]
  tabs.off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
  return tabs


def scan(root, names, sig_expect, read_header, read_meta, threads=0, stats=None):
  """Scan ``names`` under ``root``.  ``sig_expect``: the signal descriptions every record must carry (record 0's).
  Returns ``(rows, gains, baselines, names_match, metas)``: int64 (n,), float64 (n, nsig), float64 (n, nsig), bool (n,),
  ``ScannedMetas``.  Records the native parser hands back are read with ``read_header(path)`` / ``read_meta(name)`` (the
  general parsers; their exceptions propagate as they would have); ``stats['fallback']`` counts them."""
  n, nsig = len(names), len(sig_expect)
  out = (N.RecordScan * max(n, 1))()
  gains = np.zeros((n, nsig), dtype=np.float64)
  bases32 = np.zeros((n, nsig), dtype=np.int32)
  ev_time = np.zeros((n, MAX_EVENTS), dtype=np.float64)
  ev_prefix = np.zeros((n, MAX_EVENTS), dtype='S%d' % PREFIX_BYTES)
  rc = N.lib().scgrhc_scan_records(os.fsencode(root), _blob(names), n, _blob(sig_expect), nsig, MAX_EVENTS, int(threads),
                                   C.cast(out, C.c_void_p), gains.ctypes.data, bases32.ctypes.data, ev_time.ctypes.data,
                                   ev_prefix.ctypes.data)
  if rc != N.OK:
    raise N.ScgrhcError(rc, 'scgrhc_scan_records failed')
  rec = np.frombuffer(out, dtype=np.dtype([('status', '<i4'), ('nsig', '<i4'), ('rows', '<i8'), ('fs', '<f8'), ('duration_s', '<f8'),
                                           ('n_events', '<i4'), ('names_match', '<i4')], align=True), count=n)
  rows = rec['rows'].astype(np.int64)
  n_ev = rec['n_events'].astype(np.int64)
  dur = rec['duration_s'].astype(np.float64)
  match = rec['names_match'] != 0
  bases = bases32.astype(np.float64)
  if stats is not None:
    stats['fallback'] = int((rec['status'] != 0).sum())
  for r in np.nonzero(rec['status'] != 0)[0]:          # the general parsers, for whatever the scanner did not recognise
    r = int(r)
    h = read_header(os.path.join(root, names[r]))
    match[r] = list(h[0]) == list(sig_expect)
    rows[r] = int(h[2])
    if match[r]:
      gains[r] = [float(g) for g in h[3]]
      bases[r] = [float(b) for b in h[4]]
      if os.path.basename(h[5]) != names[r] + '.dat' or os.path.dirname(h[5]) != os.path.dirname(os.path.join(root, names[r])):
        raise Unscannable(names[r])
    kt = engine.event_times(read_meta(names[r]))
    if kt is None:
      n_ev[r] = -1
      continue
    keys, times = kt[0][:-1], kt[1]
    if len(keys) > MAX_EVENTS or 'END' in keys:
      raise Unscannable(names[r])
    pre = [k.split('_')[0].encode() for k in keys]
    if any(len(p) >= PREFIX_BYTES for p in pre):
      raise Unscannable(names[r])
    n_ev[r], dur[r] = len(keys), times[-1]
    ev_time[r, :len(keys)] = times[:-1]
    ev_prefix[r, :len(keys)] = pre
  return rows, gains, bases, match, ScannedMetas(n, tabs_from_arrays(n_ev, dur, ev_time, ev_prefix))
