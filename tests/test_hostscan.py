"""The native header / side-car scanner (csrc/host_scan.h, scgrhc.hostscan) against the general Python parsers
(scgrhc.wfdbio.read_header, json + engine.event_times) on generated cohorts, odd files included.  Host-only."""
import json
import os

import numpy as np
import pytest

from tests import helpers as H  # noqa: F401  (puts the package on sys.path)
from scgrhc import engine, hostscan, wfdbio

SIG = ['patch_ACC_lat', 'patch_ACC_hf', 'patch_ACC_dv', 'RHC_pressure']


def _write(root, name, header_lines, sidecar_text, frames=10):
  with open(os.path.join(root, name + '.hea'), 'w') as f:
    f.write('\n'.join(header_lines) + '\n')
  with open(os.path.join(root, name + '.json'), 'w') as f:
    f.write(sidecar_text)
  np.zeros((frames, 4), dtype='<i2').tofile(os.path.join(root, name + '.dat'))


def _sig_lines(name, gains=('200000(0)/g', '2e5(-3)/g', '200000.5(12)/g', '500(100)/mmHg'), sig=SIG, fname=None):
  return ['%s 16 %s 16 0 0 0 0 %s' % (fname or name + '.dat', g, s) for g, s in zip(gains, sig)]


def _meta(events, st='1/1/2020 10:00:00', en='1/1/2020 10:10:00', **extra):
  d = dict(extra)
  d.update({'MacStTime': st, 'ChamEvents_in_s': events, 'MacEndTime': en})
  return json.dumps(d)


def _read_meta(root):
  return lambda name: json.loads(open(os.path.join(root, name + '.json'), 'rb').read())


def _cohort(root):
  """name -> expected to be parsed natively?"""
  native = {}
  def add(name, lines, side, ok, frames=10):
    _write(root, name, lines, side, frames)
    native[name] = ok
  add('r00', ['r00 4 500 10'] + _sig_lines('r00'), _meta({'RA_1': 0, 'RV_1': 0.004, 'PA_1': 8e-3, 'PCW_1': 0.012, 'PA_2': 0.016}), True)
  add('r01', ['r01 4 500/250 7'] + _sig_lines('r01'), _meta({'PA_1': 1, 'RV_2': -0.5}, st='x 9:5:3', en='y 23:59:59'), True)          # nsamp < on disk
  add('r02', ['r02 4 500 300000'] + _sig_lines('r02'), _meta({}), True)                                                           # nsamp > on disk, no events
  add('r03', ['r03 4 500 10'] + _sig_lines('r03'), _meta([1, 2, 3]), True)                                                        # not an object
  add('r04', ['r04 4 500 10'] + _sig_lines('r04'), _meta({'PA_1': 1.5e-2, 'PA': 3}, patient={'id': 'é\\u00e9"x\\"', 'list': [1, [2, {'a': None}], True, False]}, n=-1.5E+3), True)
  add('r05', ['# a comment', 'r05 4 500 10'] + _sig_lines('r05'), _meta({'PA_1': 0}), False)                                      # comment line
  add('r06', ['r06 4 500 10'] + _sig_lines('r06'), _meta({'PA_1': '0.004', 'RV_1': 0}), False)                                    # numeric string
  add('r07', ['r07 4 500 10'] + _sig_lines('r07'), '{"MacStTime": "d 10:00:00", "MacEndTime": "d 10:10:00", "ChamEvents_in_s": {"PA_1": 1, "RV_1": 2, "PA_1": 0.002}}', False)   # repeated key
  add('r08', ['r08 4 500 10'] + _sig_lines('r08'), '{"MacStTime": "d 10:00:00", "MacEndTime": "d 10:10:00", "ChamEvents_in_s": {"P\\u0041_1": 0.002}}', False)              # escaped key
  add('r09', ['r09 4 500 10'] + _sig_lines('r09', gains=('200', '200(5)', '1(0)/g', '2(0)/x')), _meta({'PA_1': 0}), False)        # gain without baseline
  add('r10', ['r10 4 500 10'] + _sig_lines('r10'), '\n {"ChamEvents_in_s": {"RV_9": 7}, "MacStTime": "d 00:00:00", "ChamEvents_in_s": {"PA_1": 0.01}, "MacEndTime": "d 00:00:59"}\n', True)  # last one wins
  add('r11', ['r11 4 500 10'] + _sig_lines('r11'), _meta({'PA_1': 0, 'X' * 20 + '_1': 0.004}), False)                             # long prefix
  add('r12', ['r12 4 500 10'] + _sig_lines('r12'), _meta({'PA_1': float('nan')}), False)                                          # NaN literal
  add('r13', ['r13 4 500 10', ''] + _sig_lines('r13'), _meta({'PA_1': 0}), False)                                                 # blank line inside
  add('r14', ['r14 4 500 10'] + _sig_lines('r14'), _meta({'PA_1': 0.002, 'PA_2': 12345678901234567890, 'RV_1': 1e400 if False else 3}), True)
  return native


def _plan_equal(a, b):
  assert a.n_cand == b.n_cand
  assert np.array_equal(a.intervals, b.intervals)


def test_scan_matches_the_python_parsers(tmp_path):
  root = str(tmp_path)
  native = _cohort(root)
  names = sorted(native)
  # r11's prefix does not fit the fixed-width table: the scan refuses the chunk instead of truncating
  with pytest.raises(hostscan.Unscannable):
    hostscan.scan(root, names, SIG, wfdbio.read_header, _read_meta(root))
  names.remove('r11')
  stats = {}
  rows, gains, bases, same, metas = hostscan.scan(root, names, SIG, wfdbio.read_header, _read_meta(root), stats=stats)
  assert stats['fallback'] == sum(1 for n in names if not native[n])
  heads = [wfdbio.read_header(os.path.join(root, n)) for n in names]
  assert same.all()
  assert rows.tolist() == [h[2] for h in heads]
  assert gains.tolist() == [[float(g) for g in h[3]] for h in heads]
  assert bases.tolist() == [[float(b) for b in h[4]] for h in heads]
  py_metas = [_read_meta(root)(n) for n in names]
  T = [2000] * len(names)
  for chamber in ('PA', 'RV', 'RA', 'PCW', '*', 'nope'):
    for W in (1, 3):
      _plan_equal(engine.plan_cohort(metas, chamber, T, W), engine.plan_cohort(py_metas, chamber, T, W))
  # the tables themselves
  ref = engine.EventTabs(py_metas)
  assert np.array_equal(metas.tabs.times, ref.times, equal_nan=True) and np.array_equal(metas.tabs.off, ref.off)
  assert np.array_equal(metas.tabs.is_end, ref.is_end)
  assert [p.decode() for p in metas.tabs.prefix] == list(ref.prefix)


def test_scan_flags_other_layouts_and_missing_files(tmp_path):
  root = str(tmp_path)
  _write(root, 'a', ['a 4 500 10'] + _sig_lines('a'), _meta({'PA_1': 0}))
  _write(root, 'b', ['b 4 500 10'] + _sig_lines('b', sig=SIG[:3] + ['other']), _meta({'PA_1': 0}))
  _write(root, 'c', ['c 3 500 10'] + _sig_lines('c')[:3], _meta({'PA_1': 0}))
  rows, gains, bases, same, metas = hostscan.scan(root, ['a', 'b', 'c'], SIG, wfdbio.read_header, _read_meta(root))
  assert same.tolist() == [True, False, False]
  os.remove(os.path.join(root, 'a.json'))
  with pytest.raises(FileNotFoundError):
    hostscan.scan(root, ['a'], SIG, wfdbio.read_header, _read_meta(root))
  _write(root, 'd', ['d 4 500 10'] + _sig_lines('d', fname='elsewhere.dat'), _meta({'PA_1': 0}))
  with pytest.raises((FileNotFoundError, hostscan.Unscannable)):
    hostscan.scan(root, ['d'], SIG, wfdbio.read_header, _read_meta(root))


def test_scan_many_records_threaded(tmp_path):
  root = str(tmp_path)
  rng = np.random.default_rng(5)
  names = ['rec%04d' % r for r in range(300)]
  for r, n in enumerate(names):
    ev = {'%s_%d' % (rng.choice(['RA', 'RV', 'PA', 'PCW']), k): float(rng.uniform(0, 500)) for k in range(int(rng.integers(0, 9)))}
    _write(root, n, ['%s 4 500 %d' % (n, 5 + r % 7)] + _sig_lines(n), _meta(ev, en='1/1/2020 10:%02d:%02d' % (r % 60, (7 * r) % 60)), frames=8)
  stats = {}
  rows, gains, bases, same, metas = hostscan.scan(root, names, SIG, wfdbio.read_header, _read_meta(root), threads=4, stats=stats)
  assert stats['fallback'] == 0 and same.all()
  assert rows.tolist() == [min(5 + r % 7, 8) for r in range(300)]
  py = [_read_meta(root)(n) for n in names]
  for chamber in ('PA', 'RV'):
    _plan_equal(engine.plan_cohort(metas, chamber, [300000] * 300, 750), engine.plan_cohort(py, chamber, [300000] * 300, 750))
