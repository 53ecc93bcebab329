"""TEST INFRASTRUCTURE — numpy twin of the device synthetic-record generator.

This file is part of ``oracle/``: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
generator is the CUDA kernel ``scgrhc_synth_records`` (``scg-rhc-waveform_b200/csrc/synth.cu``);
both follow the same specification and must agree **bit for bit**, which is why the
specification uses only integer hashing plus individually rounded IEEE-754 fp64 add/mul
(no sin/cos, no fused multiply-add).

Specification (SURVEY.md §8(d) "Synthetic inputs", adapted so host == device exactly)
-------------------------------------------------------------------------------------
* record key      k_r   = mix(mix(seed) ^ (rec * 0xD1342543DE82EF95))
* per-item hash   h(k_r, stream, idx) = mix(k_r ^ (stream << 48) ^ idx)          (mix = splitmix64)
* heart-rate step finc  = 7730941 + h(k_r,0,0) % 7730942      (0.9 .. 1.8 Hz as a 32-bit phase step @500 Hz)
* channel phases  ph_c  = h(k_r,0,1+kind) & 0xFFFFFFFF
* shape           s(p)  = 4 * (u * (1 - |u|)),  u = 2p - 1,  p = phase * 2^-32   (a parabolic "sine")
* noise           n     = (sum of the four 16-bit fields of h(k_r,16+kind,t) - 131070) * NOISE_K   (~N(0,1))
* kinds: 0,1,2 = patch_ACC_lat/hf/dv (0.02*s + 0.005*n, phase step finc*{9,13,17});
         3 = RHC_pressure ((25 + 12*s1) + 3*s2) + 0.3*n with planted defects (below);
         4 = patch_ECG (0.8*(s*|s|) + 0.02*n, step finc); >=5 filler (0.1*n).
* defects act on the RHC channel per grid window j = t // grid (grid = 750), t' = t - j*grid,
  d = h(k_r,2,j), r = d % 10000, thresholds scaled by defect_scale/16:
    [0,500)      flat run of L in {49,50,51,200} equal samples starting at a = 100 + (d>>16)%400
    [500,1000)   linear ramp  (10 + 0.04 t') + 0.05 n            (R^2 ~ 1  -> straight line)
    [1000,1500)  dip to -60.0 on t' in [300,310)                   (below min_RHC = -50)
    [1500,2000)  one sample exactly -50.0 at t' = 375              (== min_RHC passes)
    [2000,2300)  noisy line   (5 + 0.02 t') + A n, A in [1.5,3)   (R^2 straddles 0.8)
    [2300,2600)  120-sample jitter run v_a + amp*u, amp in [0.6e-3,1.4e-3)  (range straddles 1e-3)
"""
import numpy as np

NOISE_K = 2.6429e-05           # ~ 1/std of the sum of four 16-bit uniforms
GRID = 750
MASK32 = np.uint64(0xFFFFFFFF)
SIG_KIND = {'patch_ACC_lat': 0, 'patch_ACC_hf': 1, 'patch_ACC_dv': 2, 'RHC_pressure': 3, 'patch_ECG': 4}
DEFAULT_SIG_NAMES = ['patch_ACC_lat', 'patch_ACC_hf', 'patch_ACC_dv', 'RHC_pressure']
SIG_NAMES_5 = DEFAULT_SIG_NAMES + ['patch_ECG']
_SCG_MULT = (9, 13, 17)
_DEFECT_CUM = (500, 1000, 1500, 2000, 2300, 2600)
_FLAT_LEN = np.array([49, 50, 51, 200], dtype=np.int64)


def _u64(x):
  return np.asarray(x, dtype=np.uint64)


def mix(x):
  """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
  with np.errstate(over='ignore'):
    x = _u64(x) + np.uint64(0x9E3779B97F4A7C15)
    z = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def record_key(seed, rec):
  with np.errstate(over='ignore'):
    return mix(mix(_u64(seed)) ^ (_u64(rec) * np.uint64(0xD1342543DE82EF95)))


def h(key, stream, idx):
  return mix(key ^ (np.uint64(stream) << np.uint64(48)) ^ _u64(idx))


def u01(hh):
  return (hh >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def noise(key, kind, t):
  hh = h(key, 16 + kind, t)
  s = ((hh & np.uint64(0xFFFF)) + ((hh >> np.uint64(16)) & np.uint64(0xFFFF)) +
       ((hh >> np.uint64(32)) & np.uint64(0xFFFF)) + (hh >> np.uint64(48))).astype(np.int64)
  return (s - 131070).astype(np.float64) * NOISE_K


def shape(phase):
  """phase: uint64 array holding 32-bit phases."""
  p = phase.astype(np.float64) * (2.0 ** -32)
  u = 2.0 * p - 1.0
  return 4.0 * (u * (1.0 - np.abs(u)))


def _phase(ph0, t, step):
  with np.errstate(over='ignore'):
    return (ph0 + _u64(t) * np.uint64(step)) & MASK32


def _rhc_base(key, finc, t):
  ph1 = _phase(h(key, 0, 1 + 3) & MASK32, t, finc)
  with np.errstate(over='ignore'):
    ph2 = (ph1 * np.uint64(2) + np.uint64(0x14000000)) & MASK32
  s1 = shape(ph1)
  s2 = shape(ph2)
  n = noise(key, 3, t)
  return ((25.0 + 12.0 * s1) + 3.0 * s2) + 0.3 * n


def gen_channel(key, finc, kind, t, defect_scale=16, grid=GRID):
  """One channel of one record at sample indices ``t`` (int64 array)."""
  t = np.asarray(t, dtype=np.int64)
  if kind in (0, 1, 2):
    ph = _phase(h(key, 0, 1 + kind) & MASK32, t, int(finc) * _SCG_MULT[kind])
    return 0.02 * shape(ph) + 0.005 * noise(key, kind, t)
  if kind == 4:
    ph = _phase(h(key, 0, 1 + 4) & MASK32, t, finc)
    s = shape(ph)
    return 0.8 * (s * np.abs(s)) + 0.02 * noise(key, 4, t)
  if kind != 3:
    return 0.1 * noise(key, kind, t)
  # RHC with planted defects
  y = _rhc_base(key, finc, t)
  if defect_scale <= 0:
    return y
  j = t // grid
  tp = t - j * grid
  d = h(key, 2, j)
  r = (d % np.uint64(10000)).astype(np.int64)
  cum = [c * defect_scale // 16 for c in _DEFECT_CUM]
  sel = ((d >> np.uint64(8)) & np.uint64(3)).astype(np.int64)
  a = 100 + ((d >> np.uint64(16)) % np.uint64(400)).astype(np.int64)
  amp_u = u01(h(key, 3, j))
  n = noise(key, 3, t)
  tpf = tp.astype(np.float64)
  # flat run
  L = _FLAT_LEN[sel]
  m = (r < cum[0]) & (tp >= a) & (tp < a + L)
  if m.any():
    y = np.where(m, _rhc_base(key, finc, j * grid + a), y)
  m = (r >= cum[0]) & (r < cum[1])
  if m.any():
    y = np.where(m, (10.0 + 0.04 * tpf) + 0.05 * n, y)
  m = (r >= cum[1]) & (r < cum[2]) & (tp >= 300) & (tp < 310)
  y = np.where(m, -60.0, y)
  m = (r >= cum[2]) & (r < cum[3]) & (tp == 375)
  y = np.where(m, -50.0, y)
  m = (r >= cum[3]) & (r < cum[4])
  if m.any():
    A = 1.5 + 1.5 * amp_u
    y = np.where(m, (5.0 + 0.02 * tpf) + A * n, y)
  m = (r >= cum[4]) & (r < cum[5]) & (tp >= a) & (tp < a + 120)
  if m.any():
    amp = 0.6e-3 + 0.8e-3 * amp_u
    y = np.where(m, _rhc_base(key, finc, j * grid + a) + amp * u01(h(key, 4, t)), y)
  return y


def record_params(seed, rec):
  key = record_key(seed, rec)
  finc = int(7730941 + int(h(key, 0, 0) % np.uint64(7730942)))
  return key, finc


def gen_record(seed, rec, T, kinds=(0, 1, 2, 3), defect_scale=16, grid=GRID, t0=0):
  """(T, len(kinds)) float64 row-major record ``rec`` of the cohort ``seed``."""
  key, finc = record_params(seed, rec)
  t = np.arange(t0, t0 + T, dtype=np.int64)
  out = np.empty((T, len(kinds)), dtype=np.float64)
  for c, kind in enumerate(kinds):
    out[:, c] = gen_channel(key, finc, int(kind), t, defect_scale, grid)
  return out


def kinds_for(sig_names):
  return tuple(SIG_KIND.get(n, 5 + i) for i, n in enumerate(sig_names))


# Chamber-event layouts used by tests / bench (seconds).
ORACLE_EVENTS = {"RA_1": 0, "RV_1": 120, "PA_1": 240, "PCW_1": 420, "PA_2": 480}


def record_meta(duration_s=600, events=None, start='1/1/2020 10:00:00'):
  """JSON side-car in the reference's format (`recordutil.py:97-104` reads these keys)."""
  hh, mm, ss = 10, 0, 0
  tot = hh * 3600 + mm * 60 + ss + int(duration_s)
  end = '1/1/2020 %02d:%02d:%02d' % (tot // 3600, (tot // 60) % 60, tot % 60)
  return {'MacStTime': start, 'MacEndTime': end,
          'ChamEvents_in_s': dict(ORACLE_EVENTS if events is None else events)}
