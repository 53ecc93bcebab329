"""ctypes binding of libscgrhc.so (C ABI: include/scgrhc.h).  No CPU fallback: a missing library or a
missing CUDA device raises."""
import ctypes as C
import os

from . import build as _build

MAX_C = 4
MAX_NSIG = 64
MAX_SUBSETS = 8

OK = 0
ERR_BAD_ARG, ERR_MISSING_CHANNEL, ERR_NONFINITE_RHC, ERR_CUDA, ERR_UNSUPPORTED, ERR_NO_DEVICE = -1, -2, -3, -4, -5, -6

REASON_FLAT, REASON_STRAIGHT, REASON_FLOOR, REASON_NONFINITE, REASON_AMBIGUOUS = 1, 2, 4, 8, 16
OUT_F64, PREDICATES_ONLY, USE_KEPT_LIST, NORM_GLOBAL, KEEP_ALL, KEEP_ERRORS, NORM_ZSCORE, ARENA_PLANAR = 1, 2, 4, 8, 16, 32, 64, 128


class Interval(C.Structure):
  _fields_ = [('row0', C.c_int64), ('cand0', C.c_int64), ('n_win', C.c_int32), ('rec_id', C.c_int32)]


class Job(C.Structure):
  _fields_ = [('arena', C.c_void_p), ('arena_rows', C.c_int64), ('arena_capacity_bytes', C.c_int64),
              ('nsig', C.c_int32), ('W', C.c_int32), ('C', C.c_int32), ('scg_cols', C.c_int32 * MAX_C),
              ('rhc_col', C.c_int32), ('flags', C.c_uint32), ('intervals', C.c_void_p),
              ('n_intervals', C.c_int32), ('stride', C.c_int32), ('n_cand', C.c_int64),
              ('min_rhc', C.c_double), ('flat_threshold', C.c_double), ('global_minmax', C.c_double * 4),
              ('kept_list', C.c_void_p), ('n_items', C.c_int64)]


class Outputs(C.Structure):
  _fields_ = [('scg_out', C.c_void_p), ('rhc_out', C.c_void_p), ('minmax', C.c_void_p), ('keep', C.c_void_p),
              ('reason', C.c_void_p), ('cand_win', C.c_void_p), ('cand_rec', C.c_void_p)]


class Subset(C.Structure):
  _fields_ = [('scg_out', C.c_void_p), ('minmax', C.c_void_p), ('C', C.c_int32), ('member', C.c_int32 * MAX_C)]


class Decim(C.Structure):
  _fields_ = [('taps', C.c_void_p), ('per_phase', C.c_int32), ('down', C.c_int32), ('n_pre_remove', C.c_int32), ('fused', C.c_int32),
              ('iv_in0', C.c_void_p), ('iv_len', C.c_void_p), ('iv_rel', C.c_void_p)]


class RecordScan(C.Structure):
  _fields_ = [('status', C.c_int32), ('nsig', C.c_int32), ('rows', C.c_int64), ('fs', C.c_double), ('duration_s', C.c_double),
              ('n_events', C.c_int32), ('names_match', C.c_int32)]


class Compact(C.Structure):
  _fields_ = [('kept_idx', C.c_void_p), ('start_idx', C.c_void_p), ('stop_idx', C.c_void_p),
              ('rec_id', C.c_void_p), ('n_kept', C.c_void_p), ('stride', C.c_int32), ('reserved', C.c_int32)]


class ScgrhcError(RuntimeError):
  def __init__(self, code, message):
    super().__init__('libscgrhc error %d: %s' % (code, message))
    self.code = code
    self.message = message


_lib = None

# every symbol include/scgrhc.h declares (tests check the library exports all of them)
SYMBOLS = ['scgrhc_abi_version', 'scgrhc_ctx_create', 'scgrhc_ctx_destroy', 'scgrhc_last_error',
           'scgrhc_ctx_set_tuning', 'scgrhc_ctx_set_output_planes', 'scgrhc_ctx_sm_count', 'scgrhc_plan_record', 'scgrhc_plan_cohort', 'scgrhc_scan_records', 'scgrhc_process_windows', 'scgrhc_process_windows_decim',
           'scgrhc_compact_kept', 'scgrhc_normalize_subsets', 'scgrhc_global_minmax', 'scgrhc_check_errors', 'scgrhc_ambiguous_count', 'scgrhc_gather_windows',
           'scgrhc_window_metrics', 'scgrhc_sosfiltfilt', 'scgrhc_sosfiltfilt_scan', 'scgrhc_resample_poly', 'scgrhc_gather_windows_noise', 'scgrhc_collate_batch', 'scgrhc_philox_words', 'scgrhc_rolling_range_lt', 'scgrhc_decode_fmt16', 'scgrhc_decode_fmt16_records', 'scgrhc_waveform_stats', 'scgrhc_synth_records', 'scgrhc_selftest_div']


def lib():
  """The loaded library; raises (never falls back) when it has not been built."""
  global _lib
  if _lib is not None:
    return _lib
  path = os.environ.get('SCGRHC_LIB', _build.LIB)
  if not os.path.exists(path):
    raise RuntimeError('libscgrhc.so not found at %s — build it first (python __graft_entry__.py build); '
                       'there is no CPU fallback' % path)
  L = C.CDLL(path)
  vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
  L.scgrhc_abi_version.restype = C.c_int
  L.scgrhc_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
  L.scgrhc_ctx_destroy.argtypes = [vp]
  L.scgrhc_ctx_destroy.restype = None
  L.scgrhc_last_error.argtypes = [vp]
  L.scgrhc_last_error.restype = C.c_char_p
  L.scgrhc_ctx_set_tuning.argtypes = [vp, C.c_int, C.c_int]
  L.scgrhc_ctx_sm_count.argtypes = [vp]
  L.scgrhc_ctx_set_output_planes.argtypes = [vp, i64]
  L.scgrhc_plan_record.argtypes = [C.POINTER(dbl), C.POINTER(C.c_uint8), C.c_int, i64, i32, i32, dbl, i64, i32, i64,
                                   C.POINTER(Interval), C.c_int, C.POINTER(C.c_int), C.POINTER(i64),
                                   C.POINTER(i64), C.c_int, C.POINTER(C.c_int)]
  L.scgrhc_plan_cohort.argtypes = [C.POINTER(dbl), C.POINTER(C.c_uint8), C.POINTER(i64), C.POINTER(i64), i64, i32, i32, dbl, i32,
                                   C.POINTER(Interval), i64, C.POINTER(i64), C.POINTER(i64)]
  L.scgrhc_scan_records.argtypes = [C.c_char_p, C.c_char_p, i64, C.c_char_p, i32, i32, i32, vp, vp, vp, vp, vp]
  L.scgrhc_process_windows.argtypes = [vp, C.POINTER(Job), C.POINTER(Outputs), vp]
  L.scgrhc_process_windows_decim.argtypes = [vp, C.POINTER(Job), C.POINTER(Outputs), C.POINTER(Decim), vp]
  L.scgrhc_normalize_subsets.argtypes = [vp, C.POINTER(Job), C.POINTER(Subset), i32, vp, vp]
  L.scgrhc_compact_kept.argtypes = [vp, vp, vp, vp, i64, i32, C.POINTER(Compact), vp]
  L.scgrhc_global_minmax.argtypes = [vp, vp, vp, i64, vp, vp]
  L.scgrhc_check_errors.argtypes = [vp, vp, C.POINTER(i64)]
  L.scgrhc_ambiguous_count.argtypes = [vp]
  L.scgrhc_ambiguous_count.restype = i64
  L.scgrhc_gather_windows.argtypes = [vp, vp, vp, i64, i64, vp, vp]
  L.scgrhc_gather_windows_noise.argtypes = [vp, vp, vp, i64, i64, vp, C.c_float, u64, u64, vp]
  L.scgrhc_collate_batch.argtypes = [vp, vp, vp, vp, i64, i32, i32, vp, vp, C.c_float, u64, u64, vp]
  L.scgrhc_philox_words.argtypes = [vp, u64, u64, i64, vp, vp]
  L.scgrhc_window_metrics.argtypes = [vp, vp, vp, vp, i64, i32, vp, vp]
  L.scgrhc_sosfiltfilt.argtypes = [vp, vp, vp, vp, vp, C.POINTER(i64), i32, i32, C.POINTER(i32), i32, C.POINTER(dbl), C.POINTER(dbl), i32, i32, vp]
  L.scgrhc_sosfiltfilt_scan.argtypes = [vp, vp, vp, vp, C.POINTER(i64), i32, i32, C.POINTER(i32), i32, C.POINTER(dbl), C.POINTER(dbl),
                                        i32, i32, i32, i32, vp]
  L.scgrhc_resample_poly.argtypes = [vp, vp, vp, vp, vp, vp, i32, i64, i32, i32, i32, i32, i32, i32, vp]
  L.scgrhc_rolling_range_lt.argtypes = [vp, vp, i64, i32, dbl, vp, vp]
  L.scgrhc_decode_fmt16.argtypes = [vp, vp, i64, i32, C.POINTER(i32), i32, C.POINTER(dbl), C.POINTER(dbl), vp, vp]
  L.scgrhc_decode_fmt16_records.argtypes = [vp, vp, vp, i32, i64, i32, C.POINTER(i32), i32, vp, vp, i32, vp, vp]
  L.scgrhc_waveform_stats.argtypes = [vp, vp, i64, i64, dbl, vp, vp]
  L.scgrhc_synth_records.argtypes = [vp, u64, i64, i64, i64, i32, C.POINTER(i32), i32, i32, vp, vp]
  L.scgrhc_selftest_div.argtypes = [vp, u64, i64, i32, vp, vp]
  for name in SYMBOLS:
    getattr(L, name)
  if L.scgrhc_abi_version() != 1:
    raise RuntimeError('libscgrhc ABI version mismatch')
  _lib = L
  return L


def check(ctx, rc):
  if rc != OK:
    msg = lib().scgrhc_last_error(ctx)
    raise ScgrhcError(rc, msg.decode() if msg else '')
