"""Minimal WFDB record I/O used when the ``wfdb`` package is absent (it is not in this image).

The reference obtains ``record.sig_name`` / ``record.p_signal`` from ``wfdb.rdrecord`` (recordutil.py:137).
This module restates the published WFDB header + format-16 signal file layout (physionet.org WFDB
header(5) / signal(5) pages; third-party, not part of the reference, so *parity unpinned*):

  header line 1   <name> <nsig> <fs>[/<cfs>] <nsamp> ...
  signal lines    <file> <fmt> <gain>[(<baseline>)][/<units>] <adcres> <adczero> <initval> <checksum> <blocksize> <description>
  format 16       little-endian int16, samples of all signals of a frame interleaved
  physical value  (digital - baseline) / gain   (fp64; digital -32768 = invalid sample -> NaN)

Only what the window-preparation path needs: single-segment records, one signal file, format 16.
``read_digital`` returns the int16 frames so that the GPU can do the conversion (scgrhc ops.decode).
"""
import os

import numpy as np

INVALID_16 = -32768
NATIVE_SCAN = True      # scgrhc.hostscan parses the same header shape as _fast_header below, natively and in bulk


class Record:
  def __init__(self, record_name, sig_name, fs, d_signal, adc_gain, baseline, units):
    self.record_name = record_name
    self.sig_name = list(sig_name)
    self.fs = fs
    self.d_signal = d_signal
    self.adc_gain = list(adc_gain)
    self.baseline = list(baseline)
    self.units = list(units)
    self.n_sig = len(self.sig_name)
    self.sig_len = d_signal.shape[0]
    self._p = None

  @property
  def p_signal(self):
    """(sig_len, n_sig) float64 physical samples, as wfdb's dac(): (d - baseline) / gain, invalid -> NaN."""
    if self._p is None:
      p = self.d_signal.astype(np.float64)
      bad = self.d_signal == INVALID_16
      p = (p - np.asarray(self.baseline, dtype=np.float64)) / np.asarray(self.adc_gain, dtype=np.float64)
      p[bad] = np.nan
      self._p = p
    return self._p


def _parse_header(path):
  with open(path + '.hea') as f:
    lines = [ln.strip() for ln in f if ln.strip() and not ln.startswith('#')]
  head = lines[0].split()
  name, nsig = head[0], int(head[1])
  fs = float(head[2].split('/')[0]) if len(head) > 2 else 250.0
  nsamp = int(head[3]) if len(head) > 3 else None
  sigs = []
  for ln in lines[1:1 + nsig]:
    tok = ln.split()
    fname, fmt = tok[0], tok[1].split('x')[0].split(':')[0].split('+')[0]
    gain, baseline, units = 200.0, None, 'mV'
    if len(tok) > 2:
      g = tok[2]
      if '/' in g:
        g, units = g.split('/', 1)
      if '(' in g:
        g, b = g.split('(')
        baseline = int(b.rstrip(')'))
      gain = float(g) if g else 200.0
      if gain == 0:
        gain = 200.0
    adczero = int(tok[4]) if len(tok) > 4 else 0
    if baseline is None:
      baseline = adczero
    desc = ' '.join(tok[8:]) if len(tok) > 8 else 'sig%d' % len(sigs)
    sigs.append(dict(file=fname, fmt=fmt, gain=gain, baseline=baseline, units=units, name=desc))
  return name, fs, nsamp, sigs


def _fast_header(path):
  """The common shape of a header — no comments, every signal line `<file> 16 <gain>(<baseline>)/<units> <res> <zero>
  <init> <checksum> <blocksize> <description>` — parsed without the general machinery (the streamed ingest reads one
  header per record: 100,000 of them for BASELINE configs[3]).  None when the header is anything else."""
  with open(path + '.hea', 'rb') as f:
    lines = f.read().decode().split('\n')
  head = lines[0].split()
  if len(head) < 4 or '#' in lines[0]:
    return None
  nsig = int(head[1])
  names, gains, bases, fname = [], [], [], None
  for ln in lines[1:1 + nsig]:
    tok = ln.split(None, 8)
    if len(tok) < 9 or tok[1] != '16' or '(' not in tok[2] or ln[0] == '#':
      return None
    g, rest = tok[2].split('(', 1)
    b = rest.split(')', 1)[0]
    gain = float(g)
    if gain == 0 or (fname is not None and tok[0] != fname):
      return None
    fname = tok[0]
    names.append(tok[8].strip()); gains.append(gain); bases.append(int(b))
  if len(names) != nsig:
    return None
  return names, float(head[2].split('/')[0]), int(head[3]), gains, bases, fname


def read_header(path):
  """Everything the streaming ingest needs WITHOUT touching the signal file: (sig_name, fs, n_frames, gain, baseline,
  dat_path) of a single-file format-16 record; raises NotImplementedError for other layouts.  The frame count comes from
  the header, else from the size of the signal file."""
  fast = _fast_header(path)
  if fast is not None:
    names, fs, nsamp, gains, bases, fname = fast
    dat = os.path.join(os.path.dirname(path), fname)
    on_disk = os.path.getsize(dat) // (2 * len(names))
    return names, fs, min(nsamp, on_disk), gains, bases, dat
  name, fs, nsamp, sigs = _parse_header(path)
  if not sigs:
    raise ValueError('record %s has no signals' % path)
  if any(s['fmt'] != '16' for s in sigs) or len({s['file'] for s in sigs}) != 1:
    raise NotImplementedError('wfdbio reads single-file format-16 records only (install wfdb for others)')
  dat = os.path.join(os.path.dirname(path), sigs[0]['file'])
  on_disk = os.path.getsize(dat) // (2 * len(sigs))
  T = on_disk if nsamp is None else min(nsamp, on_disk)
  return [s['name'] for s in sigs], fs, T, [s['gain'] for s in sigs], [s['baseline'] for s in sigs], dat


def read_digital(path):
  """(name, sig_name, fs, int16 frames (T, nsig), gain, baseline, units) for the record at ``path`` (no extension)."""
  name, fs, nsamp, sigs = _parse_header(path)
  if not sigs:
    raise ValueError('record %s has no signals' % path)
  if any(s['fmt'] != '16' for s in sigs) or len({s['file'] for s in sigs}) != 1:
    raise NotImplementedError('wfdbio reads single-file format-16 records only (install wfdb for others)')
  raw = np.fromfile(os.path.join(os.path.dirname(path), sigs[0]['file']), dtype='<i2')
  nsig = len(sigs)
  T = len(raw) // nsig if nsamp is None else min(nsamp, len(raw) // nsig)
  d = raw[:T * nsig].reshape(T, nsig)
  return (name, [s['name'] for s in sigs], fs, d, [s['gain'] for s in sigs], [s['baseline'] for s in sigs],
          [s['units'] for s in sigs])


def rdrecord(path):
  """Drop-in for the two attributes the reference reads from ``wfdb.rdrecord(path)``."""
  return Record(*read_digital(path))


def wrsamp(record_name, fs, units, sig_name, p_signal, write_dir='', adc_gain=None, baseline=None):
  """Write a format-16 record (synthetic cohorts / tests).  Quantises ``p_signal`` with the given or an
  automatic per-signal gain and baseline; returns the (d_signal, adc_gain, baseline) actually stored."""
  p = np.asarray(p_signal, dtype=np.float64)
  T, nsig = p.shape
  if adc_gain is None or baseline is None:
    lo, hi = np.nanmin(p, axis=0), np.nanmax(p, axis=0)
    span = np.where(hi > lo, hi - lo, 1.0)
    adc_gain = (65000.0 / span).tolist()          # leaves -32768 free for "invalid"
    baseline = np.round(-32500.0 - lo * np.asarray(adc_gain)).astype(np.int64).tolist()
  g, b = np.asarray(adc_gain, dtype=np.float64), np.asarray(baseline, dtype=np.float64)
  d = np.round(p * g + b)
  d = np.where(np.isnan(p), INVALID_16, np.clip(d, -32767, 32767)).astype('<i2')
  d.tofile(os.path.join(write_dir, record_name + '.dat'))
  with open(os.path.join(write_dir, record_name + '.hea'), 'w') as f:
    f.write('%s %d %g %d\n' % (record_name, nsig, fs, T))
    for k in range(nsig):
      f.write('%s.dat 16 %.17g(%d)/%s 16 0 %d 0 0 %s\n' % (record_name, adc_gain[k], int(baseline[k]), units[k],
                                                          int(d[0, k]), sig_name[k]))
  return d, list(adc_gain), [int(v) for v in baseline]
