#!/usr/bin/env python
"""bench.py — throughput of the SCG/RHC window-preparation hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K --warmup W

A *step* is one pass of the hot path over one synthetic cohort per GPU: BASELINE.json configs[1],
1,000 synthetic 10-min records x 4 signals (3-axis SCG + RHC pressure) with the waveform_06 params
(waveform_01, which BASELINE names, has no `chamber` key and cannot be run by the reference either —
SURVEY.md §0; waveform_06 is the first config the reference loads and has the same 3-axis shape).

  value      kept windows/s, records resident in HBM when the timed region starts (whole job, all GPUs)
  e2e        the same metric through the public host API (pinned host records -> H2D -> hot path ->
             D2H of the kept-window count/indices); windows stay device-resident by design
  roofline   algorithmic bytes of the window kernel / its CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  oracle/ref_port.py (per-window pandas + sklearn, the reference's cost profile) on
             host cores, bounded sample; the reference arm (--impl reference) runs it on all cores
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


def meta():
  return {'MacStTime': '1/1/2020 10:00:00', 'MacEndTime': '1/1/2020 10:10:00', 'ChamEvents_in_s': dict(EVENTS)}


def workload_name(n_rec):
  return ('%d synthetic 10-min records x 4 signals (3-axis SCG + RHC, fp64) x waveform_06 params '
          '(PA, C=3, W=750, min_RHC=-50, local min-max) -> fp32 windows; 400 candidate windows/record' % n_rec)


def algorithmic_bytes(n_cand, n_kept, C, out_bytes):
  """candidate = RHC read 6000 B + 1 B flag; kept additionally SCG read, outputs, 52 B metadata."""
  return n_cand * (W * 8 + 1) + n_kept * (W * C * 8 + W * (C + 1) * out_bytes + 52)


class ClockSampler:
  """nvidia-smi clocks / throttle reasons sampled while the GPU is under load."""
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, device):
    self.samples, self.proc, self.device = [], None, device

  def start(self):
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                    '-lms', '50', '-i', str(self.device)], stdout=subprocess.PIPE,
                                   stderr=subprocess.DEVNULL, text=True)
    except OSError:
      return
    def reader():
      for line in self.proc.stdout:
        f = [x.strip() for x in line.split(',')]
        if len(f) >= 8:
          self.samples.append((time.time(), f))
    threading.Thread(target=reader, daemon=True).start()

  def stop(self):
    if self.proc:
      self.proc.terminate()

  def summary(self, windows):
    """windows: list of (t0, t1) host-time intervals during which the GPU was under our load."""
    sel = [f for (t, f) in self.samples if any(a <= t <= b for a, b in windows)] or [f for _, f in self.samples]
    if not sel:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
    def num(x):
      try:
        return float(x)
      except ValueError:
        return None
    sm = [num(f[1]) for f in sel if num(f[1]) is not None]
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    reasons = [n for i, n in enumerate(names) if any(f[4 + i].lower().startswith('active') for f in sel)]
    pw = [num(f[3]) for f in sel if num(f[3]) is not None]
    return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': num(sel[0][2]), 'reasons': reasons,
            'samples': len(sel), 'power_w_max': max(pw) if pw else None}


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
# CPU arms (oracle port)
# ------------------------------------------------------------------------------------------
_worker_cache = {}


def _cpu_worker_init():
  try:
    from threadpoolctl import threadpool_limits
    threadpool_limits(1)
  except Exception:
    pass
  import torch
  torch.set_num_threads(1)


def _cpu_gen(rec):
  from oracle import synth_ref
  if rec not in _worker_cache:
    _worker_cache[rec] = synth_ref.gen_record(SEED, rec, T_ROWS, kinds=tuple(KINDS))
  return rec


def _cpu_gen_return(rec):
  from oracle import synth_ref
  return rec, synth_ref.gen_record(SEED, rec, T_ROWS, kinds=tuple(KINDS))


def _cpu_run(rec):
  from oracle import ref_port
  p = _worker_cache.get(rec)
  if p is None:
    _cpu_gen(rec)
    p = _worker_cache[rec]
  out, n_cand = ref_port.prepare_record(p, SIG, meta(), IN_CHANNELS, 'PA', 1.5, MIN_RHC)
  return len(out), n_cand


def cpu_single_core(n_rec):
  """oracle/ref_port.py on one core over n_rec records (records generated outside the timed region)."""
  _cpu_worker_init()
  for r in range(n_rec):
    _cpu_gen(r)
  t0 = time.perf_counter()
  kept = cand = 0
  for r in range(n_rec):
    k, c = _cpu_run(r)
    kept += k; cand += c
  dt = time.perf_counter() - t0
  return kept, cand, dt


def cpu_fast(n_rec, threads):
  """The C restatement (oracle/_build/liboracle.so, OpenMP) — the best host implementation we have."""
  import numpy as np
  from oracle import c_oracle, synth_ref
  arena = np.concatenate([synth_ref.gen_record(SEED, r, T_ROWS, kinds=tuple(KINDS)) for r in range(n_rec)])
  rs = np.arange(0, n_rec * T_ROWS, W, dtype=np.int64)
  c_oracle.process_windows(arena[:T_ROWS], W, [0, 1, 2], 3, rs[:400], MIN_RHC, threads=threads)   # warm
  t0 = time.perf_counter()
  keep = c_oracle.process_windows(arena, W, [0, 1, 2], 3, rs, MIN_RHC, threads=threads)[0]
  dt = time.perf_counter() - t0
  return int(keep.sum()), len(rs), dt


UNIT_WINDOWS = 50                      # reference-arm work unit: 50 candidate windows = 75 s of one record
UNIT_ROWS = UNIT_WINDOWS * W
UNITS_PER_RECORD = T_ROWS // UNIT_ROWS


def _unit_meta():
  return {'MacStTime': '1/1/2020 10:00:00', 'MacEndTime': '1/1/2020 10:01:15', 'ChamEvents_in_s': dict(EVENTS)}


def _cpu_run_unit(unit):
  from oracle import ref_port
  rec, u = divmod(unit, UNITS_PER_RECORD)
  p = _worker_cache[rec][u * UNIT_ROWS:(u + 1) * UNIT_ROWS]
  out, n_cand = ref_port.prepare_record(p, SIG, _unit_meta(), IN_CHANNELS, 'PA', 1.5, MIN_RHC)
  return len(out), n_cand


def run_reference(args):
  """Reference arm: the reference's CPU implementation of the path (its port, oracle/ref_port.py —
  /root/reference is Python and cannot travel to the GPU box) on all host cores, sharded by record.  Each step is a
  bounded sample of the workload (whole 50-window pieces of records), sized after one calibration pass so that the
  K timed steps take about a minute whatever K is."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  import multiprocessing as mp
  cores = os.cpu_count() or 1
  ctx = mp.get_context('fork')

  def generate(recs):
    with ctx.Pool(min(cores, 32), initializer=_cpu_worker_init) as pool:
      for rec, arr in pool.imap_unordered(_cpu_gen_return, recs, chunksize=1):
        _worker_cache[rec] = arr

  n_rec = max(2, -(-cores // UNITS_PER_RECORD))
  generate(range(n_rec))
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
    generate(range(n_rec, need))
    n_rec = need
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
  line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
          'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
          'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
          'config': {'workload': workload_name(args.records), 'note': 'bounded sample per step; host records resident in RAM'},
          'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample,
                           'what': 'oracle/ref_port.py: per-window pandas rolling + sklearn OLS + numpy/torch normalise, multiprocessing by record'},
          'candidate_windows_per_s': cand / dt,
          'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
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
  from scgrhc.engine import HostIngest

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

  n_rec = args.records
  C = len(IN_CHANNELS)
  cols, rcol = scgrhc.resolve_columns(SIG, IN_CHANNELS)
  lo = rank * n_rec                                           # weak scaling: every rank owns n_rec records
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
  gmm = torch.empty(4, dtype=torch.float64, device=dev)

  def kernel_step():
    ops.process_windows(arena, iv, n, W, 0, cols, rcol, MIN_RHC, 1e-3, flags, [0.0] * 4, None, 0,
                        scg, rhc, minmax, keep, reason, cand_win, cand_rec)

  def tail_step():
    ops.compact_kept(keep, cand_win, cand_rec, n, W, 0, kept_idx, start_idx, stop_idx, rec_id, n_kept_t)
    if args.global_minmax:   # BASELINE configs[3]: dataset-level min/max statistics + all-reduce
      ops.global_minmax(minmax, keep, n, gmm)
      scgrhc.allreduce_minmax(gmm)

  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

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
  tt = torch.tensor([ms_total, float(n_kept), float(n), ms_kernel], dtype=torch.float64, device=dev)
  if world > 1:
    mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    ms_total, kept_all, cand_all, ms_kernel_max = float(mx[0]), float(sm[1]), float(sm[2]), float(mx[3])
  else:
    kept_all, cand_all, ms_kernel_max = float(n_kept), float(n), ms_kernel
  ms_step = ms_total / args.steps
  value = kept_all / (ms_step * 1e-3)

  # ---- roofline of the dominant kernel (this rank's window kernel) ----
  peak, peak_src = measured_peak()
  alg = algorithmic_bytes(n, n_kept, C, out_bytes)
  achieved = alg / (ms_kernel * 1e-3) / 1e9
  traffic = None
  try:
    with open(os.path.join(ROOT, 'profiles', 'window_kernel_traffic.json')) as f:
      traffic = json.load(f).get('dram_bytes_per_launch')
  except Exception:
    pass
  roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
              'traffic': traffic, 'kernel': 'scgrhc::window_kernel<C=3,NSIG4,IDENT,%s,W=750>' % ('double' if args.out_f64 else 'float'), 'kernel_ms': ms_kernel,
              'algorithmic_bytes_per_launch': alg, 'peak_source': peak_src,
              'bytes_per_kept_window': W * C * 8 + W * 8 + W * (C + 1) * out_bytes + 53}

  # ---- BASELINE configs[2]: batches of 256 kept windows gathered on the device for the trainer ----
  kept_pos = kept_idx[:n_kept]
  perm = kept_pos[torch.randperm(n_kept, device=dev)[:256 * 64]].contiguous()
  b_scg = torch.empty((256, C, W), dtype=out_dtype, device=dev)
  b_rhc = torch.empty((256, 1, W), dtype=out_dtype, device=dev)
  nb = perm.numel() // 256
  def collate(i):
    ops.gather_windows(scg, perm[i * 256:(i + 1) * 256], b_scg)
    ops.gather_windows(rhc, perm[i * 256:(i + 1) * 256], b_rhc)
  for i in range(min(4, nb)):
    collate(i)
  ca, cb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  ca.record()
  for i in range(nb):
    collate(i)
  cb.record()
  torch.cuda.synchronize()
  ms_b = ca.elapsed_time(cb) / max(nb, 1)
  batch256 = {'us_per_batch': ms_b * 1e3, 'windows_per_s': 256 / (ms_b * 1e-3) if nb else None,
              'gbs': 2 * 256 * (C + 1) * W * out_bytes / (ms_b * 1e-3) / 1e9 if nb else None,
              'what': 'shuffled batch of 256 kept windows -> (256,%d,750)+(256,1,750) device tensors, 2 gather launches' % C}

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
    ms_rx, _ = timed(lambda: filters.resample_poly(f, rows, fs2, 500))                  # bit-identical to scipy
    ms_r, (r, rrows) = timed(lambda: filters.resample_poly(f, rows, fs2, 500, exact=False))   # one FMA per tap (1e-14)
    del f
    bufs2 = {}
    ms_w, st2 = timed(lambda: scgrhc.prepare_windows(r, plan2, cols, rcol, MIN_RHC, normalisation='zscore', buffers=bufs2, check=False))
    sl = st2.kept_idx[torch.randperm(st2.n_kept, device=dev)[:256]].contiguous()
    nb_scg = torch.empty((256, C, W2), dtype=torch.float32, device=dev)
    ms_n, _ = timed(lambda: ops.gather_windows_noise(st2.scg, sl, nb_scg, 0.01, SEED, 0), reps=20)
    gb = arena.numel() * 8 / 1e9
    alg_f = 4 * gb * C / len(SIG)                                          # filtered columns: x read, tmp written, tmp read, y written
    alg_r = 1.5 * gb
    alg_w = (plan2.n_cand * (W2 * 8 + 1) + st2.n_kept * (W2 * C * 8 + W2 * (C + 1) * 4 + 52)) / 1e9
    tot_ms = ms_f + ms_r + ms_w
    pipeline = {'what': 'extension stages ON (absent from the reference): sosfiltfilt order-4 1-40 Hz band-pass of the 3 SCG columns '
                        '(time-parallel kernel, <= 2e-12 vs scipy) -> resample_poly 500->250 Hz (all 4 columns, fused multiply-add form, <= 1e-14 vs scipy) '
                        '-> 375-sample windows, z-score -> '
                        'Philox noise fused into the batch-256 gather',
                'records': n_rec, 'kept_windows': st2.n_kept, 'candidate_windows': plan2.n_cand,
                'ms': {'bandpass': ms_f, 'resample': ms_r, 'resample_bit_identical_to_scipy': ms_rx, 'windows': ms_w,
                       'noise_batch256': ms_n, 'total_prepare': tot_ms},
                'algorithmic_gb': {'bandpass': alg_f, 'resample': alg_r, 'windows': alg_w},
                'frac_of_hbm_peak': {'bandpass': alg_f / ms_f * 1e3 / peak, 'resample': alg_r / ms_r * 1e3 / peak,
                                     'windows': alg_w / ms_w * 1e3 / peak,
                                     'total_prepare': (alg_f + alg_r + alg_w) / tot_ms * 1e3 / peak},
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
  windows = [(t_host0, t_host1)]
  def run_e2e():
    nonlocal e2e
    host = torch.empty(arena.shape, dtype=torch.float64, pin_memory=True)
    host.copy_(arena)
    torch.cuda.synchronize()
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
    ms_e2e = a.elapsed_time(b) / k_e2e
    t2 = torch.tensor([ms_e2e, float(st.n_kept)], dtype=torch.float64, device=dev)
    if world > 1:
      m2 = t2.clone(); dist.all_reduce(m2, op=dist.ReduceOp.MAX)
      s2 = t2.clone(); dist.all_reduce(s2, op=dist.ReduceOp.SUM)
      ms_e2e, kept_e2e = float(m2[0]), float(s2[1])
    else:
      kept_e2e = float(st.n_kept)
    assert st.n_kept == n_kept, 'host-ingest path kept a different number of windows'
    e2e = {'value': kept_e2e / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': ing.h2d_bytes,
           'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e, 'steps': k_e2e,
           'api': 'scgrhc.engine.HostIngest.run (pinned host fp64 records, %d-record chunks, copy/compute overlap)' % args.chunk_records,
           'h2d_gbs': ing.h2d_bytes / (ms_e2e * 1e-3) / 1e9}
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
                                      resample_exact=False))
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
                                      'note': 'every optional stage ON (band-pass, resample, z-score) per 200-record chunk, '
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
      ms_d = a.elapsed_time(b) / k_e2e
      t3 = torch.tensor([ms_d, float(std.n_kept)], dtype=torch.float64, device=dev)
      if world > 1:
        m3 = t3.clone(); dist.all_reduce(m3, op=dist.ReduceOp.MAX)
        s3 = t3.clone(); dist.all_reduce(s3, op=dist.ReduceOp.SUM)
        ms_d, kept_d = float(m3[0]), float(s3[1])
      else:
        kept_d = float(std.n_kept)
      e2e['fmt16'] = {'value': kept_d / (ms_d * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': ingd.h2d_bytes, 'ms_per_step': ms_d,
                      'kept_windows_per_step': int(kept_d), 'h2d_gbs': ingd.h2d_bytes / (ms_d * 1e-3) / 1e9,
                      'note': 'host buffers are the records as stored on disk (WFDB format 16, int16 frames); '
                              '(d - baseline) / gain runs on the device (scgrhc_decode_fmt16); cohort quantised with gains %s' % gains}
      del hostd


  if not args.no_e2e:
    try:
      run_e2e()
    except (RuntimeError, MemoryError) as exc:      # e.g. pinned host memory exhausted with 8 ranks on one host
      e2e = dict(e2e or {}, error=str(exc)[:300])

  if rank == 0:
    sampler.stop()
  clocks = sampler.summary(windows[:1]) if rank == 0 else None          # the device-resident timed region
  clocks_e2e = sampler.summary(windows[1:]) if rank == 0 and len(windows) > 1 else None

  cpu = cpu_fast_d = None
  if rank == 0 and world == 1 and not args.no_cpu:
    kept_c, cand_c, dt = cpu_single_core(args.cpu_records)
    cpu = {'value': kept_c / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
           'sample': '%d of the %d records (%d candidate windows), %.1f s' % (args.cpu_records, n_rec, cand_c, dt),
           'what': 'oracle/ref_port.py single process = the reference as shipped (no parallelism in the reference)',
           'host_cores_available': os.cpu_count()}
    try:
      th = os.cpu_count() or 1
      kf, cf, dtf = cpu_fast(min(64, n_rec), th)
      cpu_fast_d = {'value': kf / dtf, 'unit': UNIT, 'cores': th, 'kind': 'port',
                    'what': 'oracle/scgrhc_oracle.c (O(W) C restatement, OpenMP) on %d records' % min(64, n_rec)}
    except Exception as e:  # the C oracle is optional
      cpu_fast_d = {'error': str(e)[:200]}

  if rank == 0:
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': workload_name(n_rec), 'records_per_gpu': n_rec, 'candidate_windows_per_step': int(cand_all),
                       'kept_windows_per_step': int(kept_all), 'out_dtype': 'f64' if args.out_f64 else 'f32',
                       'global_minmax_allreduce': bool(args.global_minmax),
                       'l2': 'inputs %.1f GB + outputs %.1f GB per step per GPU, far larger than the 126 MB L2: no flush needed'
                             % (arena.numel() * 8 / 1e9, (scg.numel() + rhc.numel()) * out_bytes / 1e9)},
            'candidate_windows_per_s': cand_all / (ms_step * 1e-3),
            'roofline': roofline, 'cpu_baseline': cpu, 'cpu_baseline_fast': cpu_fast_d, 'e2e': e2e,
            'gpu_launches': args.steps * (4 + (2 if args.global_minmax else 0)),
            'launches_per_step': 'window_kernel + count_kept + scan_blocks + scatter_kept',
            'batch256': batch256, 'north_star_pipeline': pipeline, 'numa_bound_cpus': numa_cpus, 'clocks': clocks, 'clocks_e2e': clocks_e2e}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + '\n').encode())
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=200)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--records', type=int, default=1000, help='records per GPU (BASELINE configs[1]: 1,000)')
  ap.add_argument('--out-f64', action='store_true')
  ap.add_argument('--global-minmax', action='store_true')
  ap.add_argument('--no-e2e', action='store_true')
  ap.add_argument('--no-cpu', action='store_true')
  ap.add_argument('--no-fmt16', action='store_true')
  ap.add_argument('--no-pipeline', action='store_true')
  ap.add_argument('--e2e-steps', type=int, default=5)
  ap.add_argument('--chunk-records', type=int, default=50)
  ap.add_argument('--cpu-records', type=int, default=24)
  ap.add_argument('--ctas-per-sm', type=int, default=0)
  ap.add_argument('--stages', type=int, default=0)
  args = ap.parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_b200(args)


if __name__ == '__main__':
  main()
