#!/usr/bin/env python
"""bench.py — throughput of the SCG/RHC window-preparation hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K --warmup W

A *step* is one pass of the hot path over one synthetic cohort per GPU: BASELINE.json configs[1],
1,000 synthetic 10-min records x 4 signals (3-axis SCG + RHC pressure) with the waveform_06 params
(waveform_01, which BASELINE names, has no `chamber` key and cannot be run by the reference either —
SURVEY.md §0; waveform_06 is the first config the reference loads and has the same 3-axis shape).

  value      kept windows/s, records resident in HBM (planar layout: one fp64 plane per signal) when the timed region starts
             (whole job, all GPUs); roofline.interleaved_variant = the same step on wfdb's interleaved p_signal rows
  e2e        the same metric through the public host API (pinned host records -> H2D -> hot path ->
             D2H of the kept-window list); windows stay device-resident by design (the trainer reads them there).
             Sub-legs: fmt16 (int16 frames as stored on disk), dropin (format-16 files on tmpfs ->
             recordutil.prepare_cohort, wall clock), h2d_ceiling_gbs (bare concurrent pinned copies at this N)
  roofline   algorithmic bytes of the window kernel / its CUDA-event duration vs MEASURED_PEAKS.json; `frac` from the
             K timed steps, `sustained_frac` from a loop of >= 2 s (the board reaches its power cap there)
  legs       the other BASELINE configs at this N:  global_minmax (configs[3] mechanics on the resident cohort: pass A,
             device reduction, MIN all-reduce, pass B — all inside the timed region), config4_100k (configs[3] itself:
             100,000 records sharded over the N ranks, generated on the device chunk by chunk, two passes),
             sweep36 (configs[4]: 36 configs over one 5-signal cohort per rank, fan-out passes)
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, staged by oracle/make_ref.py) on host cores, bounded sample;
             the reference arm (--impl reference) runs it on all cores.  Falls back to the port (oracle/ref_port.py)
             only when oracle/_ref is absent.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'scg-rhc-waveform_b200')
for _p in (ROOT, PKG):
  if _p not in sys.path:
    sys.path.insert(0, _p)

SEED = 0x5C6
T_ROWS = 300000                      # 10 min @ 500 Hz
SIG = ['patch_ACC_lat', 'patch_ACC_hf', 'patch_ACC_dv', 'RHC_pressure']
KINDS = [0, 1, 2, 3]
IN_CHANNELS = SIG[:3]                # waveform_06
W = 750
MIN_RHC = -50.0
EVENTS = {'PA_1': 0}                 # one full-length PA interval -> 400 candidate windows / record
METRIC = 'preprocessed windows/sec'
UNIT = 'windows/s'

# BASELINE configs[4]: the 36 runnable waveform_NN configs = 4 chambers x 8 channel subsets + the legacy
# 02..05 (lat,hf / PA; 04 with min_RHC=0 and dataset-level min/max per project_log.txt:19-21); 01 has no chamber
SIG5 = ['patch_ACC_lat', 'patch_ACC_hf', 'patch_ACC_dv', 'RHC_pressure', 'patch_ECG']
KINDS5 = [0, 1, 2, 3, 4]
EVENTS5 = {'RA_1': 0, 'RV_1': 120, 'PA_1': 240, 'PCW_1': 420, 'PA_2': 480}
_L, _H, _D, _E = SIG5[0], SIG5[1], SIG5[2], SIG5[4]
SUBSETS = [[_L, _H, _D], [_L, _H], [_L, _D], [_H, _D], [_L], [_H], [_D], [_L, _H, _D, _E]]


def sweep_configs():
  import types
  cfg = {}
  for chamber in ('PA', 'RV', 'RA', 'PCW'):
    for k, sub in enumerate(SUBSETS):
      cfg['%s_%d' % (chamber, k)] = types.SimpleNamespace(in_channels=sub, chamber=chamber, segment_size=1.5, min_RHC=-50,
                                                          use_global_min_max=False)
  for nn in ('02', '03', '05'):
    cfg['legacy_' + nn] = types.SimpleNamespace(in_channels=[_L, _H], chamber='PA', segment_size=1.5, min_RHC=float('-inf'),
                                                use_global_min_max=False)
  cfg['legacy_04'] = types.SimpleNamespace(in_channels=[_L, _H], chamber='PA', segment_size=1.5, min_RHC=0, use_global_min_max=True)
  return cfg


def meta(events=None):
  return {'MacStTime': '1/1/2020 10:00:00', 'MacEndTime': '1/1/2020 10:10:00', 'ChamEvents_in_s': dict(events or EVENTS)}


def workload_name(n_rec):
  return ('%d synthetic 10-min records x 4 signals (3-axis SCG + RHC, fp64) x waveform_06 params '
          '(PA, C=3, W=750, min_RHC=-50, local min-max) -> fp32 windows; 400 candidate windows/record' % n_rec)


def config_dict(n_rec, gpus, out_f64=False):
  """The SAME dict from both arms (the driver compares them): only quantities that follow from the command line."""
  return {'workload': workload_name(n_rec), 'records_per_gpu': n_rec, 'candidate_windows_per_step': n_rec * 400 * gpus,
          'out_dtype': 'f64' if out_f64 else 'f32', 'params': 'waveform_06',
          'hbm_layout': 'planar (one fp64 plane per signal, as the device decode of the on-disk format writes it); roofline.interleaved_variant = wfdb p_signal rows',
          'l2': 'inputs %.1f GB + outputs up to %.1f GB per step per GPU, far larger than the 126 MB L2: no flush needed'
                % (n_rec * T_ROWS * 4 * 8 / 1e9, n_rec * 400 * W * 4 * (8 if out_f64 else 4) / 1e9)}


def algorithmic_bytes(n_cand, n_kept, C, out_bytes, w=W):
  """candidate = RHC read 6000 B + 1 B flag; kept additionally SCG read, outputs, 52 B metadata."""
  return n_cand * (w * 8 + 1) + n_kept * (w * C * 8 + w * (C + 1) * out_bytes + 52)


class ClockSampler:
  """SM clock, board power and throttle reasons sampled while the GPU is under load: NVML polled every 2 ms from a thread
  (the device-resident timed region is ~40 ms: nvidia-smi's own loop, 50 ms at best, would see it once or not at all), with
  `nvidia-smi -lms` as the fallback when pynvml is missing.  Both read the same driver counters."""
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')
  NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

  def __init__(self, device):
    self.samples, self.proc, self.device, self.source, self._stop = [], None, device, None, False
    self.period = 0.002          # the main thread relaxes it to 20 ms once the short device-resident regions are over

  def _start_nvml(self):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
    sm_max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
    reasons_fn = getattr(pynvml, 'nvmlDeviceGetCurrentClocksEventReasons', None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
    def bit(name_new, name_old):
      return getattr(pynvml, name_new, None) or getattr(pynvml, name_old)
    masks = [bit('nvmlClocksEventReasonHwSlowdown', 'nvmlClocksThrottleReasonHwSlowdown'),
             bit('nvmlClocksEventReasonHwThermalSlowdown', 'nvmlClocksThrottleReasonHwThermalSlowdown'),
             bit('nvmlClocksEventReasonSwThermalSlowdown', 'nvmlClocksThrottleReasonSwThermalSlowdown'),
             bit('nvmlClocksEventReasonSwPowerCap', 'nvmlClocksThrottleReasonSwPowerCap')]
    pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); reasons_fn(h)                 # fail here, not in the thread
    def loop():
      while not self._stop:
        try:
          r = reasons_fn(h)
          f = [str(self.device), str(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), str(sm_max),
               '%.2f' % (pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)] + ['Active' if r & m else 'Not Active' for m in masks]
          self.samples.append((time.time(), f))
        except Exception:
          pass
        time.sleep(self.period)
    threading.Thread(target=loop, daemon=True).start()
    self.source = 'nvml, 2 ms'

  def start(self):
    try:
      self._start_nvml()
      return
    except Exception:
      pass
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                    '-lms', '20', '-i', str(self.device)], stdout=subprocess.PIPE,
                                   stderr=subprocess.DEVNULL, text=True)
    except OSError:
      return
    self.source = 'nvidia-smi -lms 20'
    def reader():
      for line in self.proc.stdout:
        f = [x.strip() for x in line.split(',')]
        if len(f) >= 8:
          self.samples.append((time.time(), f))
    threading.Thread(target=reader, daemon=True).start()

  def stop(self):
    self._stop = True
    if self.proc:
      self.proc.terminate()

  def summary(self, windows):
    """windows: list of (t0, t1) host-time intervals during which the GPU was under our load.  A window too short to hold a
    sample takes the samples within 60 ms of it and says so."""
    sel = [f for (t, f) in self.samples if any(a <= t <= b for a, b in windows)]
    widened = False
    if not sel:
      sel, widened = [f for (t, f) in self.samples if any(a - 0.06 <= t <= b + 0.06 for a, b in windows)], True
    if not sel:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0, 'source': self.source}
    def num(x):
      try:
        return float(x)
      except ValueError:
        return None
    sm = [num(f[1]) for f in sel if num(f[1]) is not None]
    reasons = [n for i, n in enumerate(self.NAMES) if any(f[4 + i].lower().startswith('active') for f in sel)]
    pw = [num(f[3]) for f in sel if num(f[3]) is not None]
    out = {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': num(sel[0][2]), 'reasons': reasons,
           'samples': len(sel), 'power_w_max': max(pw) if pw else None, 'source': self.source}
    if widened:
      out['window'] = 'no sample inside the timed region: samples within 60 ms of it'
    return out


def bind_to_gpu_numa(local):
  """Pin this rank to the CPUs the driver reports as local to its GPU, so that the pinned host staging buffers are
  first-touched on the GPU's NUMA node (matters for the end-to-end leg when 8 ranks share one host)."""
  try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    ncpu = os.cpu_count() or 1
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
    want = {i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1}
    cpus = sorted(want & os.sched_getaffinity(0))
    if cpus:
      os.sched_setaffinity(0, cpus)
      return len(cpus)
  except Exception:
    pass
  return None


def measured_peak():
  try:
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
      return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)'
  except Exception:
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ------------------------------------------------------------------------------------------
# CPU arms: the unmodified reference (oracle/_ref through oracle/ref_harness.py's stub shim), else its port
# ------------------------------------------------------------------------------------------
_records = {}          # record number -> (T_ROWS, 4) fp64, generated by oracle/synth_ref.py (bit-identical to the device's)
_ref = {}              # harness + params of the staged reference, created in the parent before the pool forks


def _cpu_worker_init():
  try:
    from threadpoolctl import threadpool_limits
    threadpool_limits(1)
  except Exception:
    pass
  import torch
  torch.set_num_threads(1)


def _cpu_gen_return(rec):
  from oracle import synth_ref
  return rec, synth_ref.gen_record(SEED, rec, T_ROWS, kinds=tuple(KINDS))


def _generate(recs, pool_ctx=None):
  recs = [r for r in recs if r not in _records]
  if not recs:
    return
  if pool_ctx is None or len(recs) < 4:
    for r in recs:
      _records[r] = _cpu_gen_return(r)[1]
    return
  with pool_ctx.Pool(min(os.cpu_count() or 1, 32, len(recs)), initializer=_cpu_worker_init) as pool:
    for rec, arr in pool.imap_unordered(_cpu_gen_return, recs, chunksize=1):
      _records[rec] = arr


def reference_setup():
  """'reference' when the staged copy of the reference's own modules is present (oracle/make_ref.py), else 'port'."""
  if 'kind' in _ref:
    return _ref['kind']
  from oracle import ref_harness
  path = ref_harness.STAGED_PATH if ref_harness.staged_available() else \
      (ref_harness.REFERENCE_PATH if ref_harness.reference_available() else None)
  if path is None:
    _ref['kind'] = 'port'
    return 'port'
  h = ref_harness.ReferenceHarness(path)
  _ref.update(kind='reference', h=h, params=h.params('waveform_06'), path=path)
  return 'reference'


UNIT_WINDOWS = 50                      # reference-arm work unit: 50 candidate windows = 75 s of one record
UNIT_ROWS = UNIT_WINDOWS * W
UNITS_PER_RECORD = T_ROWS // UNIT_ROWS


def _unit_meta():
  return {'MacStTime': '1/1/2020 10:00:00', 'MacEndTime': '1/1/2020 10:01:15', 'ChamEvents_in_s': dict(EVENTS)}


def _register_units(n_rec):
  """Name every 50-window piece of the first n_rec records for the reference's wfdb.rdrecord stub (views, no copies)."""
  if _ref.get('kind') != 'reference':
    return
  h = _ref['h']
  for rec in range(n_rec):
    for u in range(UNITS_PER_RECORD):
      name = 'u%d' % (rec * UNITS_PER_RECORD + u)
      if name not in h.records:
        h.add_record(name, SIG, _records[rec][u * UNIT_ROWS:(u + 1) * UNIT_ROWS], _unit_meta())
    if ('r%d' % rec) not in h.records:
      h.add_record('r%d' % rec, SIG, _records[rec], meta())


def _reference_prepare(name):
  """get_segments + SCGDataset of the UNMODIFIED reference for one named record (recordutil.py:122-149,55-66)."""
  ru, params = _ref['h'].recordutil, _ref['params']
  segments = ru.get_segments(params, record_name=name)
  ds = ru.SCGDataset(segments, params.segment_size, None, None)
  return len(ds)


def _cpu_run_unit(unit):
  if _ref.get('kind') == 'reference':
    return _reference_prepare('u%d' % unit), UNIT_WINDOWS
  from oracle import ref_port
  rec, u = divmod(unit, UNITS_PER_RECORD)
  out, n_cand = ref_port.prepare_record(_records[rec][u * UNIT_ROWS:(u + 1) * UNIT_ROWS], SIG, _unit_meta(), IN_CHANNELS, 'PA', 1.5, MIN_RHC)
  return len(out), n_cand


def _cpu_run_record(rec):
  if _ref.get('kind') == 'reference':
    return _reference_prepare('r%d' % rec), T_ROWS // W
  from oracle import ref_port
  out, n_cand = ref_port.prepare_record(_records[rec], SIG, meta(), IN_CHANNELS, 'PA', 1.5, MIN_RHC)
  return len(out), n_cand


def _ref_what(kind):
  return ('the unmodified reference (oracle/_ref: recordutil.get_segments + waveform_noise.has_noise + SCGDataset, imported '
          'through oracle/ref_harness.py; wfdb.rdrecord stubbed with in-memory records)') if kind == 'reference' else \
      'oracle/ref_port.py: per-window pandas rolling + sklearn OLS + numpy/torch normalise (oracle/_ref not staged)'


def cpu_single_core(n_rec):
  """The reference as shipped (one process, no parallelism) over n_rec whole records, generated outside the timed region."""
  _cpu_worker_init()
  kind = reference_setup()
  _generate(range(n_rec))
  _register_units(n_rec)
  _cpu_run_unit(0)                                                  # imports, first-call overheads
  t0 = time.perf_counter()
  kept = cand = 0
  for r in range(n_rec):
    k, c = _cpu_run_record(r)
    kept += k; cand += c
  dt = time.perf_counter() - t0
  return kept, cand, dt, kind


def cpu_fast(n_rec, threads):
  """The C restatement (oracle/_build/liboracle.so, OpenMP) — the best host implementation we have."""
  import numpy as np
  from oracle import c_oracle
  _generate(range(n_rec))
  arena = np.concatenate([_records[r] for r in range(n_rec)])
  rs = np.arange(0, n_rec * T_ROWS, W, dtype=np.int64)
  c_oracle.process_windows(arena[:T_ROWS], W, [0, 1, 2], 3, rs[:400], MIN_RHC, threads=threads)   # warm
  t0 = time.perf_counter()
  keep = c_oracle.process_windows(arena, W, [0, 1, 2], 3, rs, MIN_RHC, threads=threads)[0]
  dt = time.perf_counter() - t0
  return int(keep.sum()), len(rs), dt


def _dropin_reference_record(path_name):
  """The reference's per-record path on a format-16 record ON DISK: wfdb.rdrecord (our reader stands in for the absent
  wfdb package: host-side dac to fp64) -> get_segments -> SCGDataset."""
  return _reference_prepare(path_name)


def write_fmt16_cohort(root, n_rec, rec0, arrays=None, device_arena=None):
  """n_rec synthetic records as WFDB format-16 files + JSON side-cars under ``root`` (the layout recordutil reads)."""
  import numpy as np
  from scgrhc import wfdbio
  os.makedirs(root, exist_ok=True)
  gains, bases = [2.0e5, 2.0e5, 2.0e5, 500.0], [0, 0, 0, 0]
  for r in range(n_rec):
    name = 'rec%05d' % (rec0 + r)
    if device_arena is not None:
      p = device_arena[r * T_ROWS:(r + 1) * T_ROWS].cpu().numpy()
    else:
      p = arrays[r]
    wfdbio.wrsamp(name, 500, ['g', 'g', 'g', 'mmHg'], SIG, p, write_dir=root, adc_gain=gains, baseline=bases)
    with open(os.path.join(root, name + '.json'), 'w') as f:
      json.dump(meta(), f)


def run_reference(args):
  """Reference arm: the reference's own CPU implementation of the path on all host cores, sharded by record.  Each step is a
  bounded sample of the workload (whole 50-window pieces of records), sized after one calibration pass so that the
  K timed steps take about a minute whatever K is."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  import multiprocessing as mp
  cores = os.cpu_count() or 1
  ctx = mp.get_context('fork')
  kind = reference_setup()

  n_rec = max(2, -(-cores // UNITS_PER_RECORD))
  _generate(range(n_rec), ctx)
  _register_units(n_rec)
  with ctx.Pool(cores, initializer=_cpu_worker_init) as pool:          # calibration: one unit per core, second (warm) pass
    pool.map(_cpu_run_unit, list(range(cores)), chunksize=1)
    t0 = time.perf_counter()
    pool.map(_cpu_run_unit, list(range(cores)), chunksize=1)
    t_unit = time.perf_counter() - t0
  # steps of >= 1 s keep the pool efficient (the arm must not be handicapped); the whole run is capped near 4 minutes
  k = max(args.steps, 1)
  target_step = min(max(1.0, 60.0 / k), 240.0 / k)
  per_core = min(64, max(1, int(round(target_step / max(t_unit, 1e-3)))))
  units_per_step = cores * per_core
  need = -(-units_per_step // UNITS_PER_RECORD)
  if need > n_rec:
    _generate(range(n_rec, need), ctx)
    n_rec = need
  _register_units(n_rec)
  units = list(range(units_per_step))
  with ctx.Pool(cores, initializer=_cpu_worker_init) as pool:          # forked after generation: workers share the records
    for _ in range(args.warmup):
      pool.map(_cpu_run_unit, units, chunksize=1)
    t0 = time.perf_counter()
    kept = cand = 0
    for _ in range(args.steps):
      for k, c in pool.map(_cpu_run_unit, units, chunksize=1):
        kept += k; cand += c
    dt = time.perf_counter() - t0
  value = kept / dt
  sample = ('%d candidate windows/step (%d pieces of %d windows from %d of the %d records), %d steps'
            % (units_per_step * UNIT_WINDOWS, units_per_step, UNIT_WINDOWS, n_rec, args.records, args.steps))

  # ---- the drop-in leg on the reference side: format-16 records on tmpfs -> the reference's per-record path ----
  dropin = None
  if kind == 'reference' and not args.no_dropin:
    try:
      import shutil
      import tempfile
      from scgrhc import wfdbio
      n_d = max(2, min(cores, 32))
      root = tempfile.mkdtemp(prefix='scgrhc_ref_dropin_', dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
      _generate(range(n_d), ctx)
      write_fmt16_cohort(root, n_d, 0, arrays=[_records[r] for r in range(n_d)])
      h = _ref['h']
      h.recordutil.PROCESSED_DATA_PATH = root
      sys.modules['wfdb'].rdrecord = wfdbio.rdrecord           # the absent wfdb package: our format-16 reader, host dac
      names = sorted(h.recordutil.get_record_names())
      with ctx.Pool(min(cores, n_d), initializer=_cpu_worker_init) as pool:
        t0 = time.perf_counter()
        kept_d = sum(pool.map(_dropin_reference_record, names, chunksize=1))
        dt_d = time.perf_counter() - t0
      dropin = {'value': kept_d / dt_d, 'unit': UNIT, 'records': n_d, 'seconds': dt_d, 'cores': min(cores, n_d),
                'what': 'format-16 records on tmpfs -> the reference\'s get_segments + SCGDataset per record, one process per record'}
      shutil.rmtree(root, ignore_errors=True)
    except Exception as exc:
      dropin = {'error': str(exc)[:300]}

  line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
          'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
          'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
          'config': config_dict(args.records, args.gpus),
          'note': 'bounded sample per step; host records resident in RAM',
          'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample, 'what': _ref_what(kind)},
          'candidate_windows_per_s': cand / dt,
          'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0, 'dropin': dropin},
          'gpu_launches': 0}
  print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def run_b200(args):
  import torch
  import torch.distributed as dist
  import scgrhc
  from scgrhc import ops, _native as N
  from scgrhc.engine import HostIngest, SynthSource

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  # stdout carries exactly ONE line (the JSON): libraries that print there (NCCL's version banner) go to stderr
  sys.stdout.flush()
  json_fd = os.dup(1)
  os.dup2(2, 1)
  if not torch.cuda.is_available():
    raise RuntimeError('bench.py needs a CUDA device: the hot path has no CPU fallback')
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  numa_cpus = bind_to_gpu_numa(local) if world > 1 else None
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  if args.ctas_per_sm or args.stages:
    ops.set_tuning(local, args.ctas_per_sm, args.stages)
  launches = [0]

  n_rec = args.records
  C = len(IN_CHANNELS)
  cols, rcol = scgrhc.resolve_columns(SIG, IN_CHANNELS)
  lo = rank * n_rec                                           # weak scaling: every rank owns n_rec records
  # the cohort resident in HBM twice: `planes` = one plane per signal (SCGRHC_ARENA_PLANAR, the layout the device decode of the
  # on-disk format writes and the headline step reads), `arena` = wfdb's interleaved (rows, nsig) p_signal layout (what fp64
  # host arrays look like; read by the optional stages, the fp64 end-to-end leg and the interleaved comparison step)
  planes = torch.empty((len(SIG), n_rec * T_ROWS), dtype=torch.float64, device=dev)
  ops.synth_records(planes, SEED, lo, n_rec, T_ROWS, KINDS, 16, W, n_rec * T_ROWS)
  arena = torch.empty((n_rec * T_ROWS, len(SIG)), dtype=torch.float64, device=dev)
  ops.synth_records(arena, SEED, lo, n_rec, T_ROWS, KINDS, 16, W)
  plan = scgrhc.plan_uniform(meta(), 'PA', T_ROWS, W, n_rec, rec0=lo)
  n = plan.n_cand
  iv = plan.device_intervals(dev)
  out_dtype = torch.float64 if args.out_f64 else torch.float32
  out_bytes = 8 if args.out_f64 else 4
  flags = N.OUT_F64 if args.out_f64 else 0
  scg = torch.empty((n, C, W), dtype=out_dtype, device=dev)
  rhc = torch.empty((n, 1, W), dtype=out_dtype, device=dev)
  minmax = torch.empty((n, 4), dtype=torch.float64, device=dev)
  keep = torch.empty(n, dtype=torch.uint8, device=dev)
  reason = torch.empty(n, dtype=torch.uint8, device=dev)
  cand_win = torch.empty(n, dtype=torch.int32, device=dev)
  cand_rec = torch.empty(n, dtype=torch.int32, device=dev)
  kept_idx, start_idx, stop_idx = (torch.empty(n, dtype=torch.int64, device=dev) for _ in range(3))
  rec_id = torch.empty(n, dtype=torch.int32, device=dev)
  n_kept_t = torch.zeros(1, dtype=torch.int64, device=dev)

  def kernel_step():
    ops.process_windows(planes, iv, n, W, 0, cols, rcol, MIN_RHC, 1e-3, flags | N.ARENA_PLANAR, [0.0] * 4, None, 0,
                        scg, rhc, minmax, keep, reason, cand_win, cand_rec)

  def interleaved_step():
    ops.process_windows(arena, iv, n, W, 0, cols, rcol, MIN_RHC, 1e-3, flags, [0.0] * 4, None, 0,
                        scg, rhc, minmax, keep, reason, cand_win, cand_rec)

  def tail_step():
    ops.compact_kept(keep, cand_win, cand_rec, n, W, 0, kept_idx, start_idx, stop_idx, rec_id, n_kept_t)

  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def over_ranks(values, op):
    """MAX or SUM over ranks of a list of floats."""
    if world == 1:
      return [float(v) for v in values]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == 'max' else dist.ReduceOp.SUM)
    return t.cpu().tolist()

  for _ in range(max(args.warmup, 3)):
    kernel_step(); tail_step()
  barrier()
  ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t_host0 = time.time()
  e0.record()
  for k in range(args.steps):
    ev[k][0].record()
    kernel_step()
    ev[k][1].record()
    tail_step()
  e1.record()
  barrier()
  t_host1 = time.time()
  ms_total = e0.elapsed_time(e1)
  ms_kernel = sum(a.elapsed_time(b) for a, b in ev) / args.steps
  n_kept = int(n_kept_t.item())
  ms_total, ms_kernel_max = over_ranks([ms_total, ms_kernel], 'max')
  kept_all, cand_all = over_ranks([n_kept, n], 'sum')
  ms_step = ms_total / args.steps
  value = kept_all / (ms_step * 1e-3)

  # ---- sustained: the same step back to back for >= 2 s (the board reaches its power cap) ----
  sustained = None
  if not args.no_sustained:
    n_s = max(args.steps, int(2200.0 / max(ms_step, 1e-3)))
    barrier()
    sa, sb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ka = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_s // 2, n_s)]
    ts0 = time.time()
    sa.record()
    for k in range(n_s):
      if k >= n_s // 2:
        ka[k - n_s // 2][0].record()
      kernel_step()
      if k >= n_s // 2:
        ka[k - n_s // 2][1].record()
      tail_step()
    sb.record()
    barrier()
    ts1 = time.time()
    ms_s = sa.elapsed_time(sb) / n_s
    ms_k_s = sum(a.elapsed_time(b) for a, b in ka) / len(ka)                 # second half of the loop: clocks have settled
    ms_s_max, = over_ranks([ms_s], 'max')
    sustained = {'steps': n_s, 'seconds': (ts1 - ts0), 'ms_per_step': ms_s_max, 'value': kept_all / (ms_s_max * 1e-3),
                 'kernel_ms_second_half': ms_k_s, 'window': (ts0, ts1)}

  # ---- roofline of the dominant kernel (this rank's window kernel) ----
  peak, peak_src = measured_peak()
  alg = algorithmic_bytes(n, n_kept, C, out_bytes)
  achieved = alg / (ms_kernel * 1e-3) / 1e9
  traffic, traffic_src = None, None
  try:
    with open(os.path.join(ROOT, 'profiles', 'window_planar_kernel_traffic.json')) as f:
      tj = json.load(f)
      traffic, traffic_src = tj.get('dram_bytes_per_launch'), tj.get('source')
  except Exception:
    pass
  roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
              'traffic': traffic,
              'traffic_source': 'NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` '
                                'capture of the same launch, stored in profiles/window_planar_kernel_traffic.json (%s)' % traffic_src,
              'kernel': 'scgrhc::window_planar_kernel<C=3,128,3,%s,W=750,PLAIN>' % ('double' if args.out_f64 else 'float'), 'kernel_ms': ms_kernel,
              'layout': 'planar arena (one plane per signal): a candidate costs its RHC plane, a kept window additionally its SCG planes',
              'algorithmic_bytes_per_launch': alg, 'peak_source': peak_src,
              'bytes_per_kept_window': W * C * 8 + W * 8 + W * (C + 1) * out_bytes + 53}
  if sustained:
    roofline['sustained_frac'] = alg / (sustained['kernel_ms_second_half'] * 1e-3) / 1e9 / peak
    roofline['sustained_kernel_ms'] = sustained['kernel_ms_second_half']
    roofline['sustained_loop_s'] = sustained['seconds']

  # ---- the same cohort in wfdb's INTERLEAVED (rows, nsig) layout (window_kernel): what the step costs when the records arrive as
  #      fp64 p_signal arrays; a rejected window then drags its SCG columns through DRAM by sector (traffic 1.089 x algorithmic) ----
  if not args.no_sustained:
    try:
      def inter_step():
        interleaved_step()
      for _ in range(3):
        inter_step()
      barrier()
      pe = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
      for a, b in pe:
        a.record(); inter_step(); b.record(); tail_step()
      torch.cuda.synchronize()
      ms_p = sum(a.elapsed_time(b) for a, b in pe) / len(pe)
      n_kept_p = int(n_kept_t.item())
      n_s = int(1500.0 / max(ms_p, 1e-3))
      pk = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_s // 2, n_s)]
      for k in range(n_s):
        if k >= n_s // 2:
          pk[k - n_s // 2][0].record()
        inter_step()
        if k >= n_s // 2:
          pk[k - n_s // 2][1].record()
        tail_step()
      torch.cuda.synchronize()
      ms_ps = sum(a.elapsed_time(b) for a, b in pk) / len(pk)
      launches[0] += (args.steps + n_s + 3) * 4
      roofline['interleaved_variant'] = {
          'kernel': 'scgrhc::window_kernel<C=3,NSIG4,IDENT,%s,W=750,PLAIN>' % ('double' if args.out_f64 else 'float'),
          'kernel_ms': ms_p, 'frac': alg / (ms_p * 1e-3) / 1e9 / peak, 'sustained_kernel_ms': ms_ps,
          'sustained_frac': alg / (ms_ps * 1e-3) / 1e9 / peak, 'same_kept_windows': n_kept_p == n_kept,
          'note': 'measured right after the sustained loop, i.e. with the board already at its power cap (its 20-step figure is therefore not a cold burst); '
                  'same cohort, same outputs bit for bit, wfdb\'s interleaved (rows, nsig) layout: the kernel of fp64 host cohorts and of the optional stages'}
      kernel_step(); tail_step()
      torch.cuda.synchronize()
    except Exception as exc:
      roofline['interleaved_variant'] = {'error': str(exc)[:300]}

  # ---- BASELINE configs[2]: batches of 256 kept windows for the trainer, through the loader the drop-in pickles
  #      (recordutil.WindowLoader: one scgrhc_collate_batch launch per batch, shuffled slots uploaded once per epoch) ----
  sampler.period = 0.02          # the host-side legs below are long; a 2 ms poll would only take GIL slices from them
  batch256 = None
  if not args.out_f64:
    import numpy as np
    import recordutil
    view = recordutil.SCGDataset.from_arrays(scg, rhc, ['r'] * n, np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros((n, 4)), 1.5)
    res = {}
    for tag, reuse, sigma in (('fresh_tensors', 0, 0.0), ('ring4', 4, 0.0), ('ring4_noise', 4, 0.01)):
      ld = recordutil.WindowLoader(view, batch_size=256, shuffle=True, reuse_buffers=reuse, noise_std=sigma, noise_seed=SEED)
      it = iter(ld)
      for _ in range(16):
        next(it)
      torch.cuda.synchronize()
      ca, cb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      nb, th = 0, time.perf_counter()
      ca.record()
      for batch in it:
        x, y = batch[0], batch[1]                     # waveform_train.py:358-359
        nb += 1
        if nb == 1000:
          break
      cb.record()
      torch.cuda.synchronize()
      res[tag] = {'us_per_batch': ca.elapsed_time(cb) * 1e3 / nb, 'host_us_per_batch': (time.perf_counter() - th) * 1e6 / nb, 'batches': nb}
      launches[0] += nb + 16
    us = res['ring4']['us_per_batch']
    batch256 = {'us_per_batch': us, 'windows_per_s': 256 / (us * 1e-6), 'gbs': 2 * 256 * (C + 1) * W * 4 / (us * 1e-6) / 1e9,
                'variants': res,
                'what': 'shuffled batches of 256 windows (slots anywhere in the %d-slot store) -> (256,%d,750)+(256,1,750) fp32 device '
                        'tensors through recordutil.WindowLoader: ONE scgrhc_collate_batch launch per batch; headline = ring of 4 '
                        'reused batch buffers, fresh_tensors = two torch.empty per batch like DataLoader' % (n, C)}
    del view

  # ---- the other BASELINE configs at this N --------------------------------------------------------------------
  legs = {}
  windows = [(t_host0, t_host1)]

  def leg_global_minmax():
    """configs[3] mechanics on the resident cohort: pass A (predicates + per-window pairs), ordered compaction, device
    reduction, MIN all-reduce of {min,-max} over the ranks, D2H of the kept count, pass B (normalise the kept windows
    with the dataset-level pairs into dense outputs) — every part inside the timed region."""
    bufs = dict(minmax=minmax, keep=keep, reason=reason, cand_win=cand_win, cand_rec=cand_rec, kept_idx=kept_idx,
                start_idx=start_idx, stop_idx=stop_idx, rec_id=rec_id, n_kept=n_kept_t)
    k_g = max(1, min(args.steps, 10))
    planar_legs = args.legs_layout == 'planar'
    src_arena = planes if planar_legs else arena
    for _ in range(2):
      st = scgrhc.prepare_windows(src_arena, plan, cols, rcol, MIN_RHC, use_global_min_max=True, out_dtype=out_dtype, buffers=bufs, planar=planar_legs)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k_g):
      st = scgrhc.prepare_windows(src_arena, plan, cols, rcol, MIN_RHC, use_global_min_max=True, out_dtype=out_dtype, buffers=bufs, planar=planar_legs)
    b.record()
    barrier()
    ms, = over_ranks([a.elapsed_time(b) / k_g], 'max')
    kept_g, = over_ranks([st.n_kept], 'sum')
    gm = st.global_minmax.cpu().tolist()
    launches[0] += (k_g + 2) * 7
    if planar_legs:      # pass A: the RHC plane of every candidate + the SCG planes of the kept ones (for the pairs); pass B: kept windows in, out
      bytes_ = n * (W * 8 + 33) + st.n_kept * (C * W * 8) + st.n_kept * ((C + 1) * W * 8 + (C + 1) * W * out_bytes)
    else:                # both passes read whole rows
      bytes_ = n * (4 * W * 8 + 33) + st.n_kept * (4 * W * 8 + (C + 1) * W * out_bytes)
    legs['global_minmax'] = {'value': kept_g / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'steps': k_g, 'kept_windows_per_step': int(kept_g),
                             'global_minmax': gm, 'allreduce': 'MIN of {min,-max}, 4 doubles, %s' % ('NCCL over %d ranks' % world if world > 1 else 'single rank: no-op'),
                             'frac_of_hbm_peak_this_rank': bytes_ / (ms * 1e-3) / 1e9 / peak,
                             'hbm_layout': args.legs_layout,
                             'what': 'scgrhc.prepare_windows(use_global_min_max=True) on the resident %d-record cohort per rank: pass A + compaction + '
                                     'device reduction + all-reduce + pass B, all timed' % n_rec}
    del bufs

  def leg_config4():
    """BASELINE configs[3]: 100,000 records sharded by record over the N ranks (strong scaling), with dataset-level
    min/max.  960 GB of fp64 records exist nowhere: every rank generates its block on the device chunk by chunk
    (SynthSource), twice — pass A and pass B — and the 480 GB of windows stream through a ring to a sink."""
    total = args.config4_records
    r_lo, r_hi = scgrhc.shard_records(total, rank, world)
    mine = r_hi - r_lo
    chunk = args.config4_chunk
    plan4 = scgrhc.plan_uniform(meta(), 'PA', T_ROWS, W, mine, rec0=r_lo)
    ing = HostIngest(plan4, [T_ROWS] * mine, len(SIG), dev, chunk_records=chunk, planar=(args.legs_layout == 'planar'))
    src = SynthSource(SEED, T_ROWS, KINDS, 16, W, rec0=r_lo)
    seen = [0]

    def sink(k, s, r, info):            # the consumer (a trainer, a shard writer) would read s, r on this stream here
      seen[0] += info['kept_hi'] - info['kept_lo']

    class Replay:                       # the same launch sequence without the generator: chunk buffers keep their last content
      def begin(self, ingest): pass
      def enqueue(self, k, chunk, stage): pass
      def end(self): pass

    bufs4 = {}
    res = {}
    for tag, source in (('with_generation', src), ('pipeline_only', Replay())):
      barrier()
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      th = time.time()
      a.record()
      seen[0] = 0
      st = ing.run(source, cols, rcol, MIN_RHC, out_dtype=out_dtype, buffers=bufs4, use_global_min_max=True, sink=sink)
      b.record()
      barrier()
      windows.append((th, time.time()))
      ms, = over_ranks([a.elapsed_time(b)], 'max')
      kept4, cand4 = over_ranks([st.n_kept, st.n_cand], 'sum')
      assert seen[0] == st.n_kept
      res[tag] = {'value': kept4 / (ms * 1e-3), 'seconds': ms * 1e-3, 'kept_windows': int(kept4), 'candidate_windows': int(cand4)}
      if tag == 'with_generation':
        gm = st.global_minmax.cpu().tolist()
    launches[0] += 2 * (len(ing.chunks) * 4 + 5)
    legs['config4_100k'] = {'value': res['with_generation']['value'], 'unit': UNIT, 'scaling': 'strong', 'records_total': total,
                            'records_this_rank': mine, 'chunk_records': chunk, 'hbm_layout': args.legs_layout, 'global_minmax': gm, **{k: v for k, v in res.items()},
                            'what': '%d records sharded over %d rank(s); per rank: chunks of %d records generated on the device (17 ms per 1,000 '
                                    'records, ALU bound — stands in for the disk/host feed), pass A (predicates + pairs) -> reduction -> MIN '
                                    'all-reduce -> pass B (regenerate, normalise kept windows with the dataset-level pairs, fp32 windows '
                                    'streamed through a 2-slot ring to a no-op consumer); pipeline_only = the same launches without the generator'
                                    % (total, world, chunk)}
    del bufs4, ing

  def leg_sweep():
    """BASELINE configs[4]: the 36 runnable configs as ONE job over a 5-signal cohort per rank (records sharded like any
    other job): one predicate pass + one fan-out pass per chamber (scgrhc.sweep.iter_sweep)."""
    from scgrhc import sweep
    n5 = args.sweep_records
    a5 = torch.empty((n5 * T_ROWS, len(SIG5)), dtype=torch.float64, device=dev)
    ops.synth_records(a5, SEED, rank * n5, n5, T_ROWS, KINDS5, 16, W)
    cfgs = sweep_configs()
    metas = [meta(EVENTS5)] * n5
    rows = [T_ROWS] * n5

    def once():
      kept = cand = 0
      bufs = {}
      for name, st in sweep.iter_sweep(a5, SIG5, metas, rows, cfgs, buffers=bufs, rec0=rank * n5):
        kept += st.n_kept; cand += st.n_cand
      return kept, cand

    once()
    barrier()
    reps = 3
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    th = time.time()
    a.record()
    for _ in range(reps):
      kept, cand = once()
    b.record()
    barrier()
    windows.append((th, time.time()))
    ms, = over_ranks([a.elapsed_time(b) / reps], 'max')
    kept_s, cand_s = over_ranks([kept, cand], 'sum')
    launches[0] += (reps + 1) * (4 * 6 + 4 * 8)
    legs['sweep36'] = {'value': kept_s / (ms * 1e-3), 'unit': UNIT, 'ms_per_sweep': ms, 'configs': len(cfgs), 'records_per_gpu': n5,
                       'kept_windows_per_sweep': int(kept_s), 'candidate_windows_per_sweep': int(cand_s), 'configs_per_s': len(cfgs) / (ms * 1e-3),
                       'what': '%d configs (4 chambers x 8 channel subsets + legacy 02..05; 04 with dataset-level min/max and its all-reduce) over %d '
                               '5-signal records per rank, cohort resident: per chamber ONE predicate pass + ONE fan-out pass writing all 8 subsets' % (len(cfgs), n5)}
    del a5

  # ---- the brief's full pipeline (extension stages ON; the reference has none of them: DESIGN.md §9) on the same cohort:
  #      zero-phase band-pass of the SCG columns -> 500 -> 250 Hz -> 1.5 s windows (z-score) -> noisy batch of 256 ----
  pipeline = None
  def run_pipeline():
    nonlocal pipeline
    from scgrhc import filters
    from scipy import signal as _sig
    sos = _sig.butter(4, (1.0, 40.0), btype='bandpass', fs=500, output='sos')
    rows = [T_ROWS] * n_rec
    fs2, W2 = 250, int(1.5 * 250)
    plan2 = scgrhc.plan_uniform(meta(), 'PA', T_ROWS // 2, W2, n_rec, rec0=lo)
    def timed(fn, reps=3):
      out = fn(); torch.cuda.synchronize()
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      for _ in range(reps):
        out = None
        out = fn()
      b.record(); torch.cuda.synchronize()
      return a.elapsed_time(b) / reps, out
    ms_f, f = timed(lambda: filters.sosfiltfilt(arena, rows, sos, cols, exact=False))
    ms_rx, (r, rrows) = timed(lambda: filters.resample_poly(f, rows, fs2, 500))             # bit-identical to scipy
    ms_r, _ = timed(lambda: filters.resample_poly(f, rows, fs2, 500, exact=False))            # one FMA per tap (1e-14)
    bufs2 = {}
    ms_w, st2 = timed(lambda: scgrhc.prepare_windows(r, plan2, cols, rcol, MIN_RHC, normalisation='zscore', buffers=bufs2, check=False))
    # decimation fused into the window kernel (scgrhc_process_windows_decim): the resampled cohort never exists in HBM
    spec = filters.DecimSpec.design(rows, fs2, 500)
    bufs3 = {}
    ms_dw, st3 = timed(lambda: scgrhc.prepare_windows(f, plan2, cols, rcol, MIN_RHC, normalisation='zscore', buffers=bufs3, check=False, decim=spec))
    same = st3.n_kept == st2.n_kept and torch.equal(st3.kept_idx, st2.kept_idx) and \
        torch.equal(st3.scg[st3.kept_idx[:2048]], st2.scg[st2.kept_idx[:2048]])
    spec_fma = filters.DecimSpec.design(rows, fs2, 500, fused=True)
    ms_dwf, st4 = timed(lambda: scgrhc.prepare_windows(f, plan2, cols, rcol, MIN_RHC, normalisation='zscore', buffers=bufs3, check=False, decim=spec_fma))
    del bufs3, st3, st4
    sl = st2.kept_idx[torch.randperm(st2.n_kept, device=dev)[:256]].contiguous()
    nb_scg = torch.empty((256, C, W2), dtype=torch.float32, device=dev)
    ms_n, _ = timed(lambda: ops.gather_windows_noise(st2.scg, sl, nb_scg, 0.01, SEED, 0), reps=20)
    gb = arena.numel() * 8 / 1e9
    alg_f = 4 * gb * C / len(SIG)                                          # filtered columns: x read, tmp written, tmp read, y written
    alg_r = 1.5 * gb
    alg_w = (plan2.n_cand * (W2 * 8 + 1) + st2.n_kept * (W2 * C * 8 + W2 * (C + 1) * 4 + 52)) / 1e9
    del f
    tot_sep, tot_fma = ms_f + ms_rx + ms_w, ms_f + ms_r + ms_w
    tot_ms = ms_f + ms_dw
    launches[0] += 4 * 16
    pipeline = {'what': 'extension stages ON (absent from the reference): sosfiltfilt order-4 1-40 Hz band-pass of the 3 SCG columns '
                        '(time-parallel kernel, <= 2e-12 vs scipy) -> resample_poly 500->250 Hz (all 4 columns, BIT-IDENTICAL to scipy) -> 375-sample '
                        'windows, z-score -> Philox noise fused into the batch-256 gather.  Primary figure (total_prepare): the decimation runs '
                        'INSIDE the window kernel (scgrhc_process_windows_decim, same separately rounded taps, bit-identical windows: '
                        'decim_windows_equal_separate_stages); the separate resample + window kernels and the fused-multiply-add resampler beside it',
                'records': n_rec, 'kept_windows': st2.n_kept, 'candidate_windows': plan2.n_cand,
                'decim_windows_equal_separate_stages': bool(same),
                'ms': {'bandpass': ms_f, 'decimating_windows': ms_dw, 'decimating_windows_fma': ms_dwf, 'total_prepare_fma_taps': ms_f + ms_dwf, 'resample': ms_rx, 'resample_fused_fma': ms_r, 'windows': ms_w,
                       'noise_batch256': ms_n, 'total_prepare': tot_ms, 'total_prepare_separate_stages': tot_sep,
                       'total_prepare_separate_fma_resampler': tot_fma},
                'algorithmic_gb': {'bandpass': alg_f, 'resample': alg_r, 'windows': alg_w,
                                   'note': 'per stage as if each ran alone (round-1 accounting, kept so that rounds compare): band-pass 4 crossings of the '
                                           'filtered columns, resample 1.5 x the cohort, windows per SURVEY 8d at 375 samples'},
                'frac_of_hbm_peak': {'bandpass': alg_f / ms_f * 1e3 / peak, 'decimating_windows': (alg_r + alg_w) / ms_dw * 1e3 / peak,
                                     'resample': alg_r / ms_rx * 1e3 / peak,
                                     'resample_fused_fma': alg_r / ms_r * 1e3 / peak,
                                     'windows': alg_w / ms_w * 1e3 / peak,
                                     'total_prepare': (alg_f + alg_r + alg_w) / tot_ms * 1e3 / peak,
                                     'total_prepare_fma_taps': (alg_f + alg_r + alg_w) / (ms_f + ms_dwf) * 1e3 / peak,
                                     'total_prepare_separate_stages': (alg_f + alg_r + alg_w) / tot_sep * 1e3 / peak,
                                     'total_prepare_separate_fma_resampler': (alg_f + alg_r + alg_w) / tot_fma * 1e3 / peak,
                                     'strict_minimum_traffic': (gb + (plan2.n_cand + st2.n_kept * (W2 * (C + 1) * 4 + 52)) / 1e9) / tot_ms * 1e3 / peak},
                'strict_minimum_traffic_note': 'records read once + kept windows written once, nothing else: the floor a single fused pass over the '
                                               'cohort would have; the zero-phase filter alone needs a forward and a backward pass over every record',
                'kept_windows_per_s': st2.n_kept / (tot_ms * 1e-3),
                'note': 'timed through the Python stage API (allocations included); windows = fused kernel + compaction'}
  if world == 1 and not args.no_pipeline:
    try:
      run_pipeline()
    except Exception as exc:   # the extension leg must never take the headline line down
      pipeline = {'error': str(exc)[:300]}
    torch.cuda.empty_cache()

  # ---- end to end through the host API: pinned host records -> H2D -> hot path -> D2H result ----
  e2e = None
  def run_e2e():
    nonlocal e2e
    host = torch.empty(arena.shape, dtype=torch.float64, pin_memory=True)
    host.copy_(arena)
    torch.cuda.synchronize()
    # the ceiling of this leg: bare pinned -> device copies of the same bytes, all ranks at once (no kernels)
    scratch = torch.empty_like(arena[:50 * T_ROWS])
    step_rows = scratch.shape[0]
    def bare():
      for r0 in range(0, arena.shape[0], step_rows):
        scratch[:min(step_rows, arena.shape[0] - r0)].copy_(host[r0:r0 + step_rows], non_blocking=True)
    bare()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(2):
      bare()
    b.record()
    barrier()
    ms_bare, = over_ranks([a.elapsed_time(b) / 2], 'max')
    h2d_ceiling = arena.numel() * 8 / (ms_bare * 1e-3) / 1e9
    del scratch
    ing = HostIngest(plan, [T_ROWS] * n_rec, len(SIG), dev, chunk_records=args.chunk_records)
    bufs = dict(scg=scg, rhc=rhc, minmax=minmax, keep=keep, reason=reason, cand_win=cand_win, cand_rec=cand_rec,
                kept_idx=kept_idx, start_idx=start_idx, stop_idx=stop_idx, rec_id=rec_id, n_kept=n_kept_t)
    k_e2e = max(1, min(args.steps, args.e2e_steps))
    for _ in range(2):
      st = ing.run(host, cols, rcol, MIN_RHC, out_dtype=out_dtype, buffers=bufs)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    th0 = time.time()
    a.record()
    d2h = 0
    for _ in range(k_e2e):
      st = ing.run(host, cols, rcol, MIN_RHC, out_dtype=out_dtype, buffers=bufs)
      meta_host = (st.kept_idx.cpu(), st.start_idx.cpu(), st.rec_id.cpu())     # the step's result crosses to the host
      d2h = 8 + sum(t.numel() * t.element_size() for t in meta_host)
    b.record()
    barrier()
    th1 = time.time()
    windows.append((th0, th1))
    ms_e2e, = over_ranks([a.elapsed_time(b) / k_e2e], 'max')
    kept_e2e, = over_ranks([st.n_kept], 'sum')
    assert st.n_kept == n_kept, 'host-ingest path kept a different number of windows'
    launches[0] += (k_e2e + 2) * (len(ing.chunks) + 3)
    e2e = {'value': kept_e2e / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': ing.h2d_bytes,
           'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e, 'steps': k_e2e,
           'api': 'scgrhc.engine.HostIngest.run (pinned host fp64 records, %d-record chunks, copy/compute overlap)' % args.chunk_records,
           'h2d_gbs': ing.h2d_bytes / (ms_e2e * 1e-3) / 1e9,
           'h2d_ceiling_gbs': h2d_ceiling, 'h2d_ceiling_gbs_all_ranks': h2d_ceiling * world,
           'frac_of_h2d_ceiling': ms_bare / ms_e2e,
           'note': 'per rank; h2d_ceiling_gbs = bare pinned->device copies of the same %.1f GB in the same chunks with all %d rank(s) copying at once '
                   '(max over ranks); D2H per step is the kept-window list only: the windows stay in HBM for the trainer by design '
                   '(north_star: "receive device-resident window tensors"); fp64 host arrays are the reference\'s in-memory format '
                   '(p_signal), e2e.fmt16 is what is on disk' % (arena.numel() * 8 / 1e9, world)}
    # ---- the brief's full pipeline end to end: the same pinned host records, every optional stage ON per chunk
    #      (band-pass -> 500->250 Hz -> z-score windows) between the copy and the window kernel ----
    if world == 1 and not args.no_pipeline:
      try:
        from scipy import signal as _sig
        sos = _sig.butter(4, (1.0, 40.0), btype='bandpass', fs=500, output='sos')
        out_rows = [T_ROWS // 2] * n_rec
        plan2 = scgrhc.plan_uniform(meta(), 'PA', T_ROWS // 2, 375, n_rec, rec0=lo)
        ingp = HostIngest(plan2, [T_ROWS] * n_rec, len(SIG), dev, chunk_records=200,   # one filter CTA per record: big chunks
                          stages=dict(sos=sos, filter_cols=cols, filter_exact=False, resample=(250, 500), out_rows=out_rows,
                                      resample_exact=True))
        bufp = {}
        for _ in range(2):
          stp = ingp.run(host, cols, rcol, MIN_RHC, buffers=bufp, normalisation='zscore')
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k_e2e):
          stp = ingp.run(host, cols, rcol, MIN_RHC, buffers=bufp, normalisation='zscore')
          meta_host = (stp.kept_idx.cpu(), stp.start_idx.cpu(), stp.rec_id.cpu())
        b.record()
        barrier()
        ms_p = a.elapsed_time(b) / k_e2e
        e2e['north_star_pipeline'] = {'value': stp.n_kept / (ms_p * 1e-3), 'unit': UNIT, 'ms_per_step': ms_p,
                                      'h2d_bytes_per_step': ingp.h2d_bytes, 'h2d_gbs': ingp.h2d_bytes / (ms_p * 1e-3) / 1e9,
                                      'kept_windows_per_step': stp.n_kept, 'scope': 'this rank',
                                      'note': 'every optional stage ON (band-pass, bit-identical resample, z-score) per 200-record chunk, '
                                              'overlapping the PCIe copy of the next chunk'}
        del bufp, stp
      except Exception as exc:
        e2e['north_star_pipeline'] = {'error': str(exc)[:300]}
    del host
    # ---- the same, from WFDB format-16 digital frames (what is on disk): int16 over PCIe, decode on the device ----
    if not args.no_fmt16:
      gains, bases = [2.0e5, 2.0e5, 2.0e5, 500.0], [0.0, 0.0, 0.0, 0.0]
      g = torch.tensor(gains, dtype=torch.float64, device=dev)
      hostd = torch.empty(arena.shape, dtype=torch.int16, pin_memory=True)
      step_rows = 50 * T_ROWS
      for r0 in range(0, arena.shape[0], step_rows):                       # quantise the cohort once (untimed input preparation)
        hostd[r0:r0 + step_rows].copy_(torch.clamp(torch.round(arena[r0:r0 + step_rows] * g), -32767, 32767).to(torch.int16))
      torch.cuda.synchronize()
      ingd = HostIngest(plan, [T_ROWS] * n_rec, len(SIG), dev, chunk_records=args.chunk_records, digital_nsig=len(SIG))
      dec = (list(range(len(SIG))), gains, bases)
      for _ in range(2):
        std = ingd.run(hostd, cols, rcol, MIN_RHC, out_dtype=out_dtype, buffers=bufs, decode=dec)
      barrier()
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      for _ in range(k_e2e):
        std = ingd.run(hostd, cols, rcol, MIN_RHC, out_dtype=out_dtype, buffers=bufs, decode=dec)
        meta_host = (std.kept_idx.cpu(), std.start_idx.cpu(), std.rec_id.cpu())
      b.record()
      barrier()
      ms_d, = over_ranks([a.elapsed_time(b) / k_e2e], 'max')
      kept_d, = over_ranks([std.n_kept], 'sum')
      launches[0] += (k_e2e + 2) * (2 * len(ingd.chunks) + 3)
      d2h_d = 8 + sum(t.numel() * t.element_size() for t in meta_host)
      fmt16 = {'value': kept_d / (ms_d * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': ingd.h2d_bytes, 'd2h_bytes_per_step': d2h_d,
               'ms_per_step': ms_d, 'steps': k_e2e,
               'kept_windows_per_step': int(kept_d), 'h2d_gbs': ingd.h2d_bytes / (ms_d * 1e-3) / 1e9,
               'api': 'scgrhc.engine.HostIngest.run(digital_nsig=4): pinned host int16 frames, %d-record chunks, copy / device decode / window kernel overlapped' % args.chunk_records,
               'note': 'host buffers are the records as stored on disk (WFDB format 16, int16 frames — what wfdb.rdrecord reads at recordutil.py:137); '
                       '(d - baseline) / gain runs on the device (scgrhc_decode_fmt16); cohort quantised with gains %s; per rank; D2H per step is the '
                       'kept-window list only: the windows stay in HBM for the trainer by design (north_star: "receive device-resident window tensors")' % gains}
      # headline e2e = the on-disk record format; the fp64 p_signal variant (the reference's in-memory format, 4x the PCIe bytes) beside it
      fp64_leg = {k: e2e[k] for k in ('value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step', 'ms_per_step', 'steps', 'api', 'h2d_gbs', 'frac_of_h2d_ceiling', 'note')}
      rest = {k: v for k, v in e2e.items() if k not in fp64_leg}
      e2e = dict(fmt16, fp64_p_signal=fp64_leg, **rest)
      e2e['frac_of_h2d_ceiling'] = e2e['h2d_gbs'] / e2e['h2d_ceiling_gbs']
      if not args.no_dropin:
        try:
          e2e['dropin'] = run_dropin(hostd)
        except Exception as exc:
          e2e['dropin'] = {'error': str(exc)[:300]}
      del hostd

  def run_dropin(hostd):
    """Format-16 records + JSON side-cars on tmpfs -> recordutil.prepare_cohort(params) (the drop-in's own public entry:
    header parse, plan, reader pool -> pinned ring -> H2D -> device decode with per-record tables -> fused window kernel),
    wall clock, sharded by record over the ranks like any multi-GPU job."""
    import shutil
    import tempfile
    import types
    import recordutil
    from scgrhc import wfdbio
    n_d = min(args.dropin_records, n_rec)
    base = '/dev/shm' if os.path.isdir('/dev/shm') else None
    root = os.path.join(base or tempfile.gettempdir(), 'scgrhc_bench_dropin_%s' % os.environ.get('MASTER_PORT', 'solo'))
    if rank == 0:
      shutil.rmtree(root, ignore_errors=True)
    barrier()
    os.makedirs(root, exist_ok=True)
    frames = hostd.numpy()
    side = json.dumps(meta())
    for r in range(n_d):                                                     # every rank writes the (quantised) records it generated
      name = 'rec%05d' % (rank * n_d + r)
      frames[r * T_ROWS:(r + 1) * T_ROWS].tofile(os.path.join(root, name + '.dat'))
      with open(os.path.join(root, name + '.hea'), 'w') as f:
        f.write('%s %d 500 %d\n' % (name, len(SIG), T_ROWS))
        for k2, g in enumerate([2.0e5, 2.0e5, 2.0e5, 500.0]):
          f.write('%s.dat 16 %.17g(0)/%s 16 0 %d 0 0 %s\n' % (name, g, 'mmHg' if k2 == 3 else 'g', int(frames[r * T_ROWS, k2]), SIG[k2]))
      with open(os.path.join(root, name + '.json'), 'w') as f:
        f.write(side)
    barrier()
    saved = recordutil.PROCESSED_DATA_PATH, recordutil.wfdb
    recordutil.PROCESSED_DATA_PATH, recordutil.wfdb = root, wfdbio
    try:
      params = types.SimpleNamespace(in_channels=IN_CHANNELS, chamber='PA', segment_size=1.5, min_RHC=MIN_RHC, use_global_min_max=False)
      names = recordutil.get_record_names()
      store, _ = recordutil.prepare_cohort(params, record_names=names, chunk_records=args.dropin_chunk)   # warm: page cache, allocations
      import gc
      reps, times = 5, []
      for _ in range(reps):
        del store
        gc.collect()
        barrier()
        t0 = time.perf_counter()
        store, _ = recordutil.prepare_cohort(params, record_names=names, chunk_records=args.dropin_chunk)
        host_list = store.kept_idx.cpu()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
      dt, = over_ranks([statistics.median(times)], 'max')
      launches[0] += (reps + 1) * (2 * (n_d // args.dropin_chunk + 1) + 3)
      return {'value': store.shard.total / dt, 'unit': UNIT, 'seconds': dt, 'seconds_each_rep_this_rank': times, 'records': n_d * world, 'kept_windows': store.shard.total,
              'bytes_read_per_rank': n_d * T_ROWS * len(SIG) * 2, 'read_gbs_per_rank': n_d * T_ROWS * len(SIG) * 2 / dt / 1e9,
              'what': 'wall clock (median of 5 calls, max over ranks) of recordutil.prepare_cohort(params) over %d format-16 records on tmpfs (%d per rank): JSON side-cars + headers of the shard parsed in one native call (scgrhc_scan_records), '
                      'C planner, reader pool -> ring of pinned chunks -> H2D -> scgrhc_decode_fmt16_records -> fused window kernel -> ordered kept list '
                      'on the host; includes every Python-side cost of the public entry point' % (n_d * world, n_d)}
    finally:
      recordutil.PROCESSED_DATA_PATH, recordutil.wfdb = saved
      barrier()
      if rank == 0:
        shutil.rmtree(root, ignore_errors=True)

  if not args.no_legs:
    try:
      leg_global_minmax()
    except Exception as exc:
      legs['global_minmax'] = {'error': str(exc)[:300]}
  if not args.no_e2e:
    try:
      run_e2e()
    except (RuntimeError, MemoryError) as exc:      # e.g. pinned host memory exhausted with 8 ranks on one host
      e2e = dict(e2e or {}, error=str(exc)[:300])

  # the big buffers of the headline leg are no longer needed
  del arena, planes, scg, rhc, minmax, keep, reason, cand_win, cand_rec, kept_idx, start_idx, stop_idx, rec_id
  torch.cuda.empty_cache()
  if not args.no_legs:
    for fn, key in ((leg_sweep, 'sweep36'), (leg_config4, 'config4_100k')):
      try:
        fn()
      except Exception as exc:
        legs[key] = {'error': str(exc)[:300]}
      torch.cuda.empty_cache()

  if rank == 0:
    sampler.stop()
  clocks = sampler.summary(windows[:1]) if rank == 0 else None          # the device-resident timed region
  clocks_sustained = sampler.summary([sustained['window']]) if rank == 0 and sustained else None
  clocks_other = sampler.summary(windows[1:]) if rank == 0 and len(windows) > 1 else None
  if sustained:
    sustained.pop('window')

  cpu = cpu_fast_d = None
  if rank == 0 and world == 1 and not args.no_cpu:
    kept_c, cand_c, dt, kind = cpu_single_core(args.cpu_records)
    cpu = {'value': kept_c / dt, 'unit': UNIT, 'cores': 1, 'kind': kind,
           'sample': '%d of the %d records (%d candidate windows), %.1f s' % (args.cpu_records, n_rec, cand_c, dt),
           'what': _ref_what(kind) + '; single process = the reference as shipped (it has no parallelism)',
           'host_cores_available': os.cpu_count()}
    try:
      th = os.cpu_count() or 1
      kf, cf, dtf = cpu_fast(min(args.cpu_records, n_rec), th)
      cpu_fast_d = {'value': kf / dtf, 'unit': UNIT, 'cores': th, 'kind': 'port',
                    'what': 'oracle/scgrhc_oracle.c (O(W) C restatement, OpenMP) on %d records' % min(args.cpu_records, n_rec)}
    except Exception as e:  # the C oracle is optional
      cpu_fast_d = {'error': str(e)[:200]}

  if rank == 0:
    n_timed = args.steps * 4
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': config_dict(n_rec, world, args.out_f64),
            'kept_windows_per_step': int(kept_all),
            'candidate_windows_per_s': cand_all / (ms_step * 1e-3),
            'roofline': roofline, 'sustained': sustained, 'cpu_baseline': cpu, 'cpu_baseline_fast': cpu_fast_d, 'e2e': e2e,
            'gpu_launches': n_timed,
            'gpu_launches_all_legs': n_timed + launches[0] + (sustained['steps'] * 4 if sustained else 0),
            'launches_per_step': 'window_planar_kernel + count_kept + scan_blocks + scatter_kept',
            'legs': legs, 'batch256': batch256, 'north_star_pipeline': pipeline, 'numa_bound_cpus': numa_cpus,
            'clocks': clocks, 'clocks_sustained': clocks_sustained, 'clocks_other_legs': clocks_other}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + '\n').encode())
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=20)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--records', type=int, default=1000, help='records per GPU (BASELINE configs[1]: 1,000)')
  ap.add_argument('--out-f64', action='store_true')
  ap.add_argument('--no-e2e', action='store_true')
  ap.add_argument('--no-cpu', action='store_true')
  ap.add_argument('--no-fmt16', action='store_true')
  ap.add_argument('--no-pipeline', action='store_true')
  ap.add_argument('--no-legs', action='store_true')
  ap.add_argument('--no-dropin', action='store_true')
  ap.add_argument('--no-sustained', action='store_true')
  ap.add_argument('--legs-layout', default='planar', choices=['planar', 'interleaved'],
                  help='HBM layout of the cohort in legs.global_minmax / config4_100k (two-pass job: 3.7 ms per step on planes, 3.9 on interleaved rows)')
  ap.add_argument('--e2e-steps', type=int, default=5)
  ap.add_argument('--chunk-records', type=int, default=50)
  ap.add_argument('--cpu-records', type=int, default=16)
  ap.add_argument('--config4-records', type=int, default=100000, help='BASELINE configs[3]: records of the whole cohort (all ranks)')
  ap.add_argument('--config4-chunk', type=int, default=500)
  ap.add_argument('--sweep-records', type=int, default=1000, help='BASELINE configs[4]: 5-signal records per GPU')
  ap.add_argument('--dropin-records', type=int, default=500, help='format-16 records per GPU written to tmpfs for e2e.dropin')
  ap.add_argument('--dropin-chunk', type=int, default=25)
  ap.add_argument('--ctas-per-sm', type=int, default=0)
  ap.add_argument('--stages', type=int, default=0)
  args = ap.parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_b200(args)


if __name__ == '__main__':
  main()
