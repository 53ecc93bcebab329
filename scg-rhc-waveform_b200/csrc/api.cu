// C ABI of libscgrhc (declared in include/scgrhc.h): context, host planner, kernel launchers.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "aux_kernels.cuh"
#include "filter_kernels.cuh"
#include "filter_scan_kernel.cuh"
#include "window_kernel.cuh"
#include "window_planar_kernel.cuh"
#include "subset_kernel.cuh"
#include "host_scan.h"

using namespace scgrhc;

struct scgrhc_ctx {
  int device = 0;
  int sm_count = 0;
  int ctas_per_sm = 0;  // 0 = from occupancy
  int stages = 0;       // 0 = default
  unsigned long long* err_dev = nullptr;  // 3 words: flags, first bad candidate, ambiguous windows
  long long out_plane = 0;                // scgrhc_ctx_set_output_planes
  long long last_ambiguous = 0;           // word 2 as read by the last scgrhc_check_errors
  int* block_counts = nullptr;
  long long* block_offsets = nullptr;
  long long scan_cap = 0;
  double* mm_partial = nullptr;
  std::string last_error;
};

static std::string g_create_error;

// NVTX range per C-ABI call: shows up in Nsight timelines, costs nothing when no tool is attached
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

static int fail(scgrhc_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->last_error = buf; else g_create_error = buf;
  return code;
}
#define CUDA_TRY(ctx, expr)                                                                         \
  do {                                                                                              \
    cudaError_t e_ = (expr);                                                                        \
    if (e_ != cudaSuccess) return fail(ctx, SCGRHC_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

extern "C" int scgrhc_abi_version(void) { return SCGRHC_ABI_VERSION; }

extern "C" const char* scgrhc_last_error(const scgrhc_ctx* ctx) {
  return ctx ? ctx->last_error.c_str() : g_create_error.c_str();
}

extern "C" int scgrhc_ctx_create(int device, scgrhc_ctx** out) {
  if (!out) return fail(nullptr, SCGRHC_ERR_BAD_ARG, "scgrhc_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(nullptr, SCGRHC_ERR_NO_DEVICE, "no CUDA device available (libscgrhc has no CPU fallback)");
  }
  if (device < 0 || device >= n) return fail(nullptr, SCGRHC_ERR_BAD_ARG, "device %d out of range (0..%d)", device, n - 1);
  scgrhc_ctx* ctx = new (std::nothrow) scgrhc_ctx();
  if (!ctx) return fail(nullptr, SCGRHC_ERR_CUDA, "out of host memory");
  ctx->device = device;
  CUDA_TRY(nullptr, cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    delete ctx;
    return fail(nullptr, SCGRHC_ERR_UNSUPPORTED, "device %d is sm_%d%d; libscgrhc is built for sm_100a only", device, prop.major, prop.minor);
  }
  ctx->sm_count = prop.multiProcessorCount;
  CUDA_TRY(nullptr, cudaMalloc(&ctx->err_dev, 4 * sizeof(unsigned long long)));
  CUDA_TRY(nullptr, cudaMalloc(&ctx->mm_partial, GMM_BLOCKS * 4 * sizeof(double)));
  *out = ctx;
  return SCGRHC_OK;
}

extern "C" void scgrhc_ctx_destroy(scgrhc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaFree(ctx->err_dev);
  cudaFree(ctx->mm_partial);
  cudaFree(ctx->block_counts);
  cudaFree(ctx->block_offsets);
  delete ctx;
}

extern "C" int scgrhc_ctx_set_tuning(scgrhc_ctx* ctx, int ctas_per_sm, int stages) {
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (ctas_per_sm < 0 || ctas_per_sm > 32 || stages < 0 || stages > kMaxStages)
    return fail(ctx, SCGRHC_ERR_BAD_ARG, "tuning out of range: ctas_per_sm=%d stages=%d", ctas_per_sm, stages);
  ctx->ctas_per_sm = ctas_per_sm;
  ctx->stages = stages;
  return SCGRHC_OK;
}

extern "C" int scgrhc_ctx_set_output_planes(scgrhc_ctx* ctx, int64_t plane_stride) {
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (plane_stride < 0) return fail(ctx, SCGRHC_ERR_BAD_ARG, "plane_stride must be >= 0");
  ctx->out_plane = plane_stride;
  return SCGRHC_OK;
}

extern "C" int scgrhc_ctx_sm_count(const scgrhc_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

// ---- host planner (recordutil.py:104-109,118,141) -------------------------------------------------
static int64_t py_trunc(double v) { return (int64_t)std::trunc(v); }

// Python's slice(a, b).indices(T) for step 1.
static void slice_indices(int64_t a, int64_t b, int64_t T, int64_t* lo, int64_t* hi) {
  auto clamp = [T](int64_t v) {
    if (v < 0) { v += T; if (v < 0) v = 0; }
    else if (v > T) v = T;
    return v;
  };
  *lo = clamp(a);
  *hi = clamp(b);
}

extern "C" int scgrhc_plan_record(const double* event_time, const uint8_t* event_match, int n_events, int64_t T,
                                  int32_t W, int32_t stride, double fs, int64_t rec_base_row, int32_t rec_id, int64_t cand_base,
                                  scgrhc_interval* out, int out_cap, int* n_out, int64_t* n_cand, int64_t* bounds,
                                  int bounds_cap, int* n_bounds) {
  if (stride == 0) stride = W;
  if (!(fs > 0.0)) fs = (double)SCGRHC_SAMPLE_FREQ;
  if (!n_out || !n_cand || W <= 0 || stride < 0 || T < 0 || n_events < 0 || (n_events && (!event_time || !event_match)))
    return SCGRHC_ERR_BAD_ARG;
  std::vector<int> order(n_events);
  for (int i = 0; i < n_events; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return event_time[a] < event_time[b]; });
  int k = 0, nb = 0;
  int64_t cand = cand_base;
  for (int i = 0; i + 1 < n_events; ++i) {
    const int e = order[i];
    if (!event_match[e]) continue;
    const int64_t a = py_trunc(event_time[e] * fs);
    const int64_t b = py_trunc(event_time[order[i + 1]] * fs);
    if (bounds && nb < bounds_cap) { bounds[2 * nb] = a; bounds[2 * nb + 1] = b; }
    ++nb;
    int64_t lo, hi;
    slice_indices(a, b, T, &lo, &hi);
    const int64_t L = hi > lo ? hi - lo : 0;
    const int64_t nw = stride == W ? L / W : (L >= W ? (L - W) / stride + 1 : 0);
    if (nw <= 0) continue;
    if (nw > INT32_MAX) return SCGRHC_ERR_BAD_ARG;
    if (out && k < out_cap) {
      out[k].row0 = rec_base_row + lo;
      out[k].cand0 = cand;
      out[k].n_win = (int32_t)nw;
      out[k].rec_id = rec_id;
    }
    ++k;
    cand += nw;
  }
  *n_out = k;
  *n_cand = cand - cand_base;
  if (n_bounds) *n_bounds = nb;
  if ((out && k > out_cap) || (bounds && nb > bounds_cap)) return SCGRHC_ERR_BAD_ARG;
  return SCGRHC_OK;
}

// The same for a whole cohort in one call: record r owns events ev_off[r] .. ev_off[r+1] and T_rows[r] arena rows, records
// back to back; rec_id = rec0 + r.  out_cap must hold one interval per matching event.
extern "C" int scgrhc_plan_cohort(const double* event_time, const uint8_t* event_match, const int64_t* ev_off, const int64_t* T_rows,
                                  int64_t n_rec, int32_t W, int32_t stride, double fs, int32_t rec0, scgrhc_interval* out,
                                  int64_t out_cap, int64_t* n_out, int64_t* n_cand) {
  if (!n_out || !n_cand || n_rec < 0 || (n_rec && (!ev_off || !T_rows))) return SCGRHC_ERR_BAD_ARG;
  int64_t k = 0, cand = 0, base = 0;
  for (int64_t r = 0; r < n_rec; ++r) {
    const int64_t e0 = ev_off[r], ne = ev_off[r + 1] - e0;
    if (ne < 0 || ne > INT32_MAX || out_cap - k > INT32_MAX) return SCGRHC_ERR_BAD_ARG;
    int nk = 0;
    int64_t nc = 0;
    const int rc = scgrhc_plan_record(event_time + e0, event_match + e0, (int)ne, T_rows[r], W, stride, fs, base, (int32_t)(rec0 + r), cand,
                                      out ? out + k : nullptr, (int)(out_cap - k), &nk, &nc, nullptr, 0, nullptr);
    if (rc != SCGRHC_OK) return rc;
    k += nk; cand += nc; base += T_rows[r];
  }
  *n_out = k;
  *n_cand = cand;
  return SCGRHC_OK;
}

extern "C" int scgrhc_scan_records(const char* dir, const char* names_blob, int64_t n, const char* expect_sig_blob, int32_t nsig_expect,
                                   int32_t max_events, int32_t threads, scgrhc_record_scan* out, double* gains, int32_t* baselines,
                                   double* ev_time, char* ev_prefix) {
  try {
    return hostscan::scan_records(dir, names_blob, n, expect_sig_blob, nsig_expect, max_events, threads, out, gains, baselines, ev_time, ev_prefix);
  } catch (const std::bad_alloc&) {
    return SCGRHC_ERR_BAD_ARG;
  }
}

// ---- hot path launcher -------------------------------------------------------------------------------
template <int C, bool NSIG4, bool IDENT, typename OutT, int WCT, bool PLAIN = false>
static int launch_window(scgrhc_ctx* ctx, const KParams& P, long long items, cudaStream_t st) {
  auto kern = window_kernel<C, NSIG4, IDENT, OutT, WCT, false, 0, 0, false, PLAIN>;
  const size_t smem = ((sizeof(Scratch) + 127) & ~size_t(127)) + (size_t)P.stages * P.stage_elems * sizeof(double);
  CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
  if (occ < 1) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "window of %d samples x %d signals does not fit in shared memory (%zu B/CTA)", P.job.W, P.job.nsig, smem);
  if (ctx->ctas_per_sm > 0) occ = std::min(occ, ctx->ctas_per_sm);
  long long grid = std::min<long long>(items, (long long)ctx->sm_count * occ);
  kern<<<(unsigned)grid, NT, smem, st>>>(P);
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}

template <int C, bool NSIG4, bool IDENT, typename OutT>
static int dispatch_w(scgrhc_ctx* ctx, const KParams& P, long long items, cudaStream_t st) {
  if (P.job.W == 750) {                                                                       // int(1.5 * 500): all 37 configs
    constexpr unsigned kModes = SCGRHC_USE_KEPT_LIST | SCGRHC_PREDICATES_ONLY | SCGRHC_NORM_GLOBAL | SCGRHC_KEEP_ALL | SCGRHC_NORM_ZSCORE;
    if constexpr (IDENT && sizeof(OutT) == 4) {          // what save_dataloaders launches: no mode flag, fp32, the uploaded column layout
      if ((P.job.flags & kModes) == 0) return launch_window<C, NSIG4, IDENT, OutT, 750, true>(ctx, P, items, st);
    }
    return launch_window<C, NSIG4, IDENT, OutT, 750>(ctx, P, items, st);
  }
  if constexpr (NSIG4 && sizeof(OutT) == 4) {
    if (P.job.W == 375) return launch_window<C, NSIG4, IDENT, OutT, 375>(ctx, P, items, st);    // 1.5 s at 250 Hz: compile-time length
    if (P.job.W <= 3 * NT) return launch_window<C, NSIG4, IDENT, OutT, -3>(ctx, P, items, st);  // resampled cohorts (1.5 s at <= 250 Hz)
  }
  return launch_window<C, NSIG4, IDENT, OutT, 0>(ctx, P, items, st);
}
template <int C, bool NSIG4, bool IDENT>
static int dispatch_out(scgrhc_ctx* ctx, const KParams& P, long long items, cudaStream_t st) {
  if (P.job.flags & SCGRHC_OUT_F64) return dispatch_w<C, NSIG4, IDENT, double>(ctx, P, items, st);
  return dispatch_w<C, NSIG4, IDENT, float>(ctx, P, items, st);
}
template <int C>
static int dispatch_nsig(scgrhc_ctx* ctx, const KParams& P, long long items, cudaStream_t st) {
  const scgrhc_job& J = P.job;
  if (J.nsig == 4) {
    if constexpr (C == 3) {
      if (J.scg_cols[0] == 0 && J.scg_cols[1] == 1 && J.scg_cols[2] == 2 && J.rhc_col == 3)
        return dispatch_out<3, true, true>(ctx, P, items, st);
    }
    return dispatch_out<C, true, false>(ctx, P, items, st);
  }
  if constexpr (C != 3) {                 // the drop-in uploads exactly the selected columns, in order, then RHC: compile-time row layout
    bool ident = J.nsig == C + 1 && J.rhc_col == C;
    for (int c = 0; c < C; ++c) ident = ident && J.scg_cols[c] == c;
    if (ident) return dispatch_out<C, false, true>(ctx, P, items, st);
  }
  return dispatch_out<C, false, false>(ctx, P, items, st);
}

template <int C, int NTH, int PR, typename OutT, int WCT, bool PLAIN = false>
static int launch_planar(scgrhc_ctx* ctx, const KParams& P, long long items, cudaStream_t st) {
  auto kern = window_planar_kernel<C, NTH, PR, OutT, WCT, PLAIN>;
  const size_t smem = ((sizeof(PScratch<NTH>) + 127) & ~size_t(127)) + (size_t)(PNR + 2 * C) * P.stage_elems * sizeof(double);
  CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NTH, smem));
  if (occ < 1) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "planar window of %d samples x %d channels does not fit in shared memory (%zu B/CTA)", P.job.W, C, smem);
  if (ctx->ctas_per_sm > 0) occ = std::min(occ, ctx->ctas_per_sm);
  const long long grid = std::min<long long>(items, (long long)ctx->sm_count * occ);
  kern<<<(unsigned)grid, NTH, smem, st>>>(P);
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}
template <int C, typename OutT>
static int dispatch_planar_w(scgrhc_ctx* ctx, const KParams& P, long long items, cudaStream_t st) {
  if (P.job.W == 750) {                                                                  // int(1.5 * 500): all 37 configs
    constexpr unsigned kModes = SCGRHC_USE_KEPT_LIST | SCGRHC_PREDICATES_ONLY | SCGRHC_NORM_GLOBAL | SCGRHC_KEEP_ALL;
    if constexpr (sizeof(OutT) == 4) {
      if ((P.job.flags & kModes) == 0) return launch_planar<C, 128, 3, OutT, 750, true>(ctx, P, items, st);
    }
    return launch_planar<C, 128, 3, OutT, 750>(ctx, P, items, st);
  }
  if (P.job.W <= 384) return launch_planar<C, 64, 3, OutT, 0>(ctx, P, items, st);      // resampled cohorts: 375 samples, 2 warps per window
  if (P.job.W <= 768) return launch_planar<C, 128, 3, OutT, 0>(ctx, P, items, st);
  return launch_planar<C, 128, 4, OutT, 0>(ctx, P, items, st);
}
template <int C>
static int dispatch_planar(scgrhc_ctx* ctx, const KParams& P, long long items, cudaStream_t st) {
  if (P.job.flags & SCGRHC_OUT_F64) return dispatch_planar_w<C, double>(ctx, P, items, st);
  return dispatch_planar_w<C, float>(ctx, P, items, st);
}

extern "C" int scgrhc_process_windows(scgrhc_ctx* ctx, const scgrhc_job* job, const scgrhc_outputs* out, void* stream) {
  NvtxRange nvtx_range("scgrhc_process_windows");
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (!job || !out) return fail(ctx, SCGRHC_ERR_BAD_ARG, "job/out is NULL");
  const scgrhc_job& J = *job;
  const bool use_list = J.flags & SCGRHC_USE_KEPT_LIST;
  const bool pred_only = J.flags & SCGRHC_PREDICATES_ONLY;
  if (J.C < 1 || J.C > SCGRHC_MAX_C) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "C=%d SCG channels (supported 1..%d)", J.C, SCGRHC_MAX_C);
  if (J.nsig < 1 || J.nsig > SCGRHC_MAX_NSIG) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "nsig=%d (supported 1..%d)", J.nsig, SCGRHC_MAX_NSIG);
  if (J.W < 2 || J.W > RMAX * NT) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "window of %d samples (supported 2..%d)", J.W, RMAX * NT);
  if (J.rhc_col < 0 || J.rhc_col >= J.nsig) return fail(ctx, SCGRHC_ERR_MISSING_CHANNEL, "RHC column %d outside 0..%d", J.rhc_col, J.nsig - 1);
  for (int c = 0; c < J.C; ++c)
    if (J.scg_cols[c] < 0 || J.scg_cols[c] >= J.nsig) return fail(ctx, SCGRHC_ERR_MISSING_CHANNEL, "SCG column %d outside 0..%d", J.scg_cols[c], J.nsig - 1);
  if (J.n_cand < 0 || J.n_intervals < 0 || J.stride < 0) return fail(ctx, SCGRHC_ERR_BAD_ARG, "negative counts");
  if (J.arena_capacity_bytes < J.arena_rows * (int64_t)J.nsig * 8) return fail(ctx, SCGRHC_ERR_BAD_ARG, "arena_capacity_bytes smaller than the arena");
  if ((reinterpret_cast<uintptr_t>(J.arena) & 15) != 0) return fail(ctx, SCGRHC_ERR_BAD_ARG, "arena must be 16-byte aligned");
  const long long items = use_list ? J.n_items : J.n_cand;
  if (use_list && pred_only) return fail(ctx, SCGRHC_ERR_BAD_ARG, "USE_KEPT_LIST and PREDICATES_ONLY are exclusive");
  if ((J.flags & SCGRHC_NORM_ZSCORE) && (use_list || (J.flags & SCGRHC_NORM_GLOBAL)))
    return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "NORM_ZSCORE is a per-window mode: not combinable with NORM_GLOBAL / USE_KEPT_LIST");
  if (use_list && !J.kept_list && items) return fail(ctx, SCGRHC_ERR_BAD_ARG, "kept_list is NULL");
  if (!use_list && (!out->keep || !out->reason || !out->minmax || !out->cand_win || !out->cand_rec) && items)
    return fail(ctx, SCGRHC_ERR_BAD_ARG, "keep/reason/minmax/cand_win/cand_rec outputs are required");
  if (use_list && !(J.flags & SCGRHC_NORM_GLOBAL) && !out->minmax && items) return fail(ctx, SCGRHC_ERR_BAD_ARG, "minmax input required");
  if (!pred_only && (!out->scg_out || !out->rhc_out) && items) return fail(ctx, SCGRHC_ERR_BAD_ARG, "scg_out/rhc_out are required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (!(J.flags & SCGRHC_KEEP_ERRORS)) {
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->err_dev, 0, sizeof(unsigned long long), st));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->err_dev + 1, 0xFF, sizeof(unsigned long long), st));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->err_dev + 2, 0, sizeof(unsigned long long), st));
  }
  if (items == 0) return SCGRHC_OK;
  if (!J.intervals || J.n_intervals == 0) return fail(ctx, SCGRHC_ERR_BAD_ARG, "candidates without intervals");

  KParams P;
  P.job = J;
  P.out = *out;
  P.err = ctx->err_dev;
  P.stages = ctx->stages > 0 ? ctx->stages : 2;
  P.stage_elems = (int)(((long long)J.W * J.nsig + J.nsig + 2 + 1) & ~1LL);  // + one padded row (next-row loads) + lead
  P.arena_elems_cap = J.arena_capacity_bytes / 8;
  if (J.flags & SCGRHC_ARENA_PLANAR) {
    if (J.flags & SCGRHC_NORM_ZSCORE) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "ARENA_PLANAR is not combinable with NORM_ZSCORE");
    if ((reinterpret_cast<uintptr_t>(out->scg_out) & 7) || (reinterpret_cast<uintptr_t>(out->rhc_out) & 7))
      return fail(ctx, SCGRHC_ERR_BAD_ARG, "scg_out/rhc_out must be 8-byte aligned");
    P.stage_elems = (J.W + 3 + 1) & ~1;    // lead + window + the next-sample pad, even: plane buffers stay 16-byte aligned
    switch (J.C) {
      case 1: return dispatch_planar<1>(ctx, P, items, st);
      case 2: return dispatch_planar<2>(ctx, P, items, st);
      case 3: return dispatch_planar<3>(ctx, P, items, st);
      default: return dispatch_planar<4>(ctx, P, items, st);
    }
  }
  switch (J.C) {
    case 1: return dispatch_nsig<1>(ctx, P, items, st);
    case 2: return dispatch_nsig<2>(ctx, P, items, st);
    case 3: return dispatch_nsig<3>(ctx, P, items, st);
    default: return dispatch_nsig<4>(ctx, P, items, st);
  }
}

// ---- the same with the decimating front end (extension): native-rate arena -> model-rate windows inside the kernel --------
template <int C, bool IDENT, int DDOWN, int DPP, bool DFMA, int WCT = -3>
static int launch_window_decim_t(scgrhc_ctx* ctx, const KParamsDecim& P, long long items, cudaStream_t st) {
  auto kern = window_kernel<C, true, IDENT, float, WCT, true, DDOWN, DPP, DFMA>;
  const size_t smem = ((sizeof(Scratch) + 127) & ~size_t(127)) +
                      ((size_t)P.k.stage_elems + (size_t)(P.k.job.W + 1) * 4 + kDecimMaxTaps) * sizeof(double);
  CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
  if (occ < 1) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "decimating window of %d samples (x%d, %d taps) does not fit in shared memory (%zu B/CTA)", P.k.job.W, P.d.down, P.d.pp, smem);
  if (ctx->ctas_per_sm > 0) occ = std::min(occ, ctx->ctas_per_sm);
  const long long grid = std::min<long long>(items, (long long)ctx->sm_count * occ);
  kern<<<(unsigned)grid, NT, smem, st>>>(P);
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}

template <int C, bool IDENT>
static int launch_window_decim(scgrhc_ctx* ctx, const KParamsDecim& P, long long items, bool fma, cudaStream_t st) {
  if (P.d.down == 2 && P.d.pp == 43) {                   // 500 -> 250 Hz, scipy's kaiser-5 design: unrolled, taps as constant operands
    if (P.k.job.W == 375)                                // ... and 1.5 s windows: compile-time length
      return fma ? launch_window_decim_t<C, IDENT, 2, 43, true, 375>(ctx, P, items, st) : launch_window_decim_t<C, IDENT, 2, 43, false, 375>(ctx, P, items, st);
    return fma ? launch_window_decim_t<C, IDENT, 2, 43, true>(ctx, P, items, st) : launch_window_decim_t<C, IDENT, 2, 43, false>(ctx, P, items, st);
  }
  return fma ? launch_window_decim_t<C, IDENT, 0, 0, true>(ctx, P, items, st) : launch_window_decim_t<C, IDENT, 0, 0, false>(ctx, P, items, st);
}

extern "C" int scgrhc_process_windows_decim(scgrhc_ctx* ctx, const scgrhc_job* job, const scgrhc_outputs* out, const scgrhc_decim* dec,
                                            void* stream) {
  NvtxRange nvtx_range("scgrhc_process_windows_decim");
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (!job || !out || !dec) return fail(ctx, SCGRHC_ERR_BAD_ARG, "job/out/dec is NULL");
  const scgrhc_job& J = *job;
  if (J.nsig != 4) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "the decimating front end reads 4-signal records (nsig=%d)", J.nsig);
  if (J.C < 1 || J.C > 3) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "C=%d SCG channels (supported 1..3 of the 4 signals)", J.C);
  if (J.W < 2 || J.W > kDecimR * NT) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "window of %d samples (supported 2..%d)", J.W, kDecimR * NT);
  if (J.flags & (SCGRHC_USE_KEPT_LIST | SCGRHC_NORM_GLOBAL | SCGRHC_OUT_F64 | SCGRHC_ARENA_PLANAR))
    return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "process_windows_decim: fp32 outputs, per-window pairs, interleaved arena only");
  if (dec->per_phase < 1 || dec->per_phase > kDecimMaxTaps || dec->down < 1 || dec->down > 16 || dec->n_pre_remove < 0 || !dec->taps)
    return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "process_windows_decim: 1..%d taps per output, down 1..16", kDecimMaxTaps);
  if (J.rhc_col < 0 || J.rhc_col >= 4) return fail(ctx, SCGRHC_ERR_MISSING_CHANNEL, "RHC column %d outside 0..3", J.rhc_col);
  for (int c = 0; c < J.C; ++c)
    if (J.scg_cols[c] < 0 || J.scg_cols[c] >= 4) return fail(ctx, SCGRHC_ERR_MISSING_CHANNEL, "SCG column %d outside 0..3", J.scg_cols[c]);
  if (J.n_cand < 0 || J.n_intervals < 0 || J.stride < 0) return fail(ctx, SCGRHC_ERR_BAD_ARG, "negative counts");
  if ((reinterpret_cast<uintptr_t>(J.arena) & 15) != 0) return fail(ctx, SCGRHC_ERR_BAD_ARG, "arena must be 16-byte aligned");
  const bool pred_only = J.flags & SCGRHC_PREDICATES_ONLY;
  if (J.n_cand && (!out->keep || !out->reason || !out->minmax || !out->cand_win || !out->cand_rec))
    return fail(ctx, SCGRHC_ERR_BAD_ARG, "keep/reason/minmax/cand_win/cand_rec outputs are required");
  if (!pred_only && (!out->scg_out || !out->rhc_out) && J.n_cand) return fail(ctx, SCGRHC_ERR_BAD_ARG, "scg_out/rhc_out are required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (!(J.flags & SCGRHC_KEEP_ERRORS)) {
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->err_dev, 0, sizeof(unsigned long long), st));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->err_dev + 1, 0xFF, sizeof(unsigned long long), st));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->err_dev + 2, 0, sizeof(unsigned long long), st));
  }
  if (J.n_cand == 0) return SCGRHC_OK;
  if (!J.intervals || J.n_intervals == 0 || !dec->iv_in0 || !dec->iv_len || !dec->iv_rel)
    return fail(ctx, SCGRHC_ERR_BAD_ARG, "candidates without intervals / record tables");
  KParamsDecim P;
  memset(&P, 0, sizeof P);
  P.k.job = J;
  P.k.out = *out;
  P.k.err = ctx->err_dev;
  P.k.stages = 1;
  const int PB = kDecimR * dec->down;
  const int rows_tot = ((J.W + kDecimR - 1) / kDecimR) * PB + dec->per_phase;      // rows the last thread's sliding pass may touch
  P.k.stage_elems = (rows_tot + rows_tot / PB + 2) * 4;
  P.k.arena_elems_cap = J.arena_capacity_bytes / 8;
  for (int i = 0; i < dec->per_phase; ++i) P.d.taps[i] = dec->taps[i];
  P.d.pp = dec->per_phase; P.d.down = dec->down; P.d.npr = dec->n_pre_remove;
  P.d.rows_in = (J.W - 1) * dec->down + dec->per_phase;
  P.d.pb_magic = (unsigned)(((1ULL << 32) + PB - 1) / PB);
  P.d.iv_in0 = reinterpret_cast<const long long*>(dec->iv_in0);
  P.d.iv_len = reinterpret_cast<const long long*>(dec->iv_len);
  P.d.iv_rel = reinterpret_cast<const long long*>(dec->iv_rel);
  const bool ident = J.C == 3 && J.scg_cols[0] == 0 && J.scg_cols[1] == 1 && J.scg_cols[2] == 2 && J.rhc_col == 3;
  const bool fma = dec->fused != 0;
  switch (J.C) {
    case 1: return launch_window_decim<1, false>(ctx, P, J.n_cand, fma, st);
    case 2: return launch_window_decim<2, false>(ctx, P, J.n_cand, fma, st);
    default: return ident ? launch_window_decim<3, true>(ctx, P, J.n_cand, fma, st) : launch_window_decim<3, false>(ctx, P, J.n_cand, fma, st);
  }
}

extern "C" int scgrhc_normalize_subsets(scgrhc_ctx* ctx, const scgrhc_job* job, const scgrhc_subset* subsets, int32_t n_subsets,
                                        void* rhc_out, void* stream) {
  NvtxRange nvtx_range("scgrhc_normalize_subsets");
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (!job || !subsets || n_subsets < 1 || n_subsets > SCGRHC_MAX_SUBSETS) return fail(ctx, SCGRHC_ERR_BAD_ARG, "normalize_subsets: 1..%d subsets", SCGRHC_MAX_SUBSETS);
  const scgrhc_job& J = *job;
  if (J.C < 1 || J.C > SCGRHC_MAX_C) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "C=%d superset channels (supported 1..%d)", J.C, SCGRHC_MAX_C);
  if (J.nsig < 1 || J.nsig > SCGRHC_MAX_NSIG) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "nsig=%d (supported 1..%d)", J.nsig, SCGRHC_MAX_NSIG);
  if (J.W < 2 || J.W > RMAX * NT) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "window of %d samples (supported 2..%d)", J.W, RMAX * NT);
  if (J.rhc_col < 0 || J.rhc_col >= J.nsig) return fail(ctx, SCGRHC_ERR_MISSING_CHANNEL, "RHC column %d outside 0..%d", J.rhc_col, J.nsig - 1);
  for (int c = 0; c < J.C; ++c)
    if (J.scg_cols[c] < 0 || J.scg_cols[c] >= J.nsig) return fail(ctx, SCGRHC_ERR_MISSING_CHANNEL, "SCG column %d outside 0..%d", J.scg_cols[c], J.nsig - 1);
  if (J.n_items < 0 || J.n_intervals < 0 || J.stride < 0) return fail(ctx, SCGRHC_ERR_BAD_ARG, "negative counts");
  if ((reinterpret_cast<uintptr_t>(J.arena) & 15) != 0) return fail(ctx, SCGRHC_ERR_BAD_ARG, "arena must be 16-byte aligned");
  if (J.arena_capacity_bytes < J.arena_rows * (int64_t)J.nsig * 8) return fail(ctx, SCGRHC_ERR_BAD_ARG, "arena_capacity_bytes smaller than the arena");
  if (J.n_items == 0) return SCGRHC_OK;
  if (!J.kept_list || !J.intervals || J.n_intervals == 0 || !rhc_out) return fail(ctx, SCGRHC_ERR_BAD_ARG, "normalize_subsets: kept_list, intervals and rhc_out are required");
  SubsetParams P;
  P.job = J;
  P.rhc_out = rhc_out;
  P.n_sub = n_subsets;
  for (int k = 0; k < n_subsets; ++k) {
    const scgrhc_subset& U = subsets[k];
    if (U.C < 1 || U.C > J.C || !U.scg_out || !U.minmax) return fail(ctx, SCGRHC_ERR_BAD_ARG, "normalize_subsets: subset %d is malformed", k);
    int mask = 0;
    for (int i = 0; i < U.C; ++i) {
      if (U.member[i] < 0 || U.member[i] >= J.C || (i && U.member[i] <= U.member[i - 1]))
        return fail(ctx, SCGRHC_ERR_BAD_ARG, "normalize_subsets: subset %d members must be strictly ascending superset indices", k);
      mask |= 1 << U.member[i];
    }
    P.sub[k].scg_out = U.scg_out; P.sub[k].minmax = U.minmax; P.sub[k].mask = mask; P.sub[k].C = U.C;
  }
  P.stage_elems = (int)(((long long)J.W * J.nsig + J.nsig + 2 + 1) & ~1LL);
  P.arena_elems_cap = J.arena_capacity_bytes / 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t smem = ((sizeof(SubsetScratch) + 127) & ~size_t(127)) + (size_t)2 * P.stage_elems * sizeof(double);
  auto launch = [&](auto kern) -> int {
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    if (occ < 1) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "window of %d samples x %d signals does not fit in shared memory (%zu B/CTA)", J.W, J.nsig, smem);
    const long long grid = std::min<long long>(J.n_items, (long long)ctx->sm_count * occ);
    kern<<<(unsigned)grid, NT, smem, st>>>(P);
    CUDA_TRY(ctx, cudaGetLastError());
    return SCGRHC_OK;
  };
  if (J.W <= 6 * NT) {                                       // 750-sample windows: 6 rows per thread
    if (J.flags & SCGRHC_OUT_F64) return launch(subset_norm_kernel<double, 6>);
    return launch(subset_norm_kernel<float, 6>);
  }
  if (J.flags & SCGRHC_OUT_F64) return launch(subset_norm_kernel<double, RMAX>);
  return launch(subset_norm_kernel<float, RMAX>);
}

extern "C" int scgrhc_check_errors(scgrhc_ctx* ctx, void* stream, int64_t* first_bad_cand) {
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long h[3] = {0, 0, 0};
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->err_dev, sizeof h, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  ctx->last_ambiguous = (long long)h[2];
  if (first_bad_cand) *first_bad_cand = (h[0] & 1ull) ? (int64_t)h[1] : -1;
  if (h[0] & 1ull)
    return fail(ctx, SCGRHC_ERR_NONFINITE_RHC, "Input y contains NaN. (candidate window %lld; waveform_noise.py:32)", (long long)h[1]);
  return SCGRHC_OK;
}

extern "C" int64_t scgrhc_ambiguous_count(const scgrhc_ctx* ctx) { return ctx ? ctx->last_ambiguous : -1; }

// ---- ordered compaction ----------------------------------------------------------------------------
static int ensure_scan(scgrhc_ctx* ctx, long long nblocks) {
  if (nblocks <= ctx->scan_cap) return SCGRHC_OK;
  cudaFree(ctx->block_counts);
  cudaFree(ctx->block_offsets);
  ctx->block_counts = nullptr; ctx->block_offsets = nullptr; ctx->scan_cap = 0;
  const long long cap = std::max<long long>(nblocks, 1024);
  CUDA_TRY(ctx, cudaMalloc(&ctx->block_counts, cap * sizeof(int)));
  CUDA_TRY(ctx, cudaMalloc(&ctx->block_offsets, cap * sizeof(long long)));
  ctx->scan_cap = cap;
  return SCGRHC_OK;
}

extern "C" int scgrhc_compact_kept(scgrhc_ctx* ctx, const uint8_t* keep, const int32_t* cand_win, const int32_t* cand_rec,
                                   int64_t n_cand, int32_t W, const scgrhc_compact* out, void* stream) {
  NvtxRange nvtx_range("scgrhc_compact_kept");
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (!out || !out->n_kept || n_cand < 0) return fail(ctx, SCGRHC_ERR_BAD_ARG, "compact: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (n_cand == 0) { CUDA_TRY(ctx, cudaMemsetAsync(out->n_kept, 0, sizeof(int64_t), st)); return SCGRHC_OK; }
  if (!keep || !out->kept_idx || (out->start_idx && (!cand_win || !out->stop_idx)) || (out->rec_id && !cand_rec))
    return fail(ctx, SCGRHC_ERR_BAD_ARG, "compact: NULL array");
  const long long nblocks = (n_cand + CTILE - 1) / CTILE;
  if (nblocks > INT32_MAX) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "too many candidates");
  int rc = ensure_scan(ctx, nblocks);
  if (rc) return rc;
  count_kept_kernel<<<(unsigned)nblocks, CB, 0, st>>>(keep, n_cand, ctx->block_counts);
  scan_blocks_kernel<<<1, 1024, 0, st>>>(ctx->block_counts, (int)nblocks, ctx->block_offsets,
                                         reinterpret_cast<long long*>(out->n_kept));
  scatter_kept_kernel<<<(unsigned)nblocks, CB, 0, st>>>(keep, cand_win, cand_rec, n_cand, W, out->stride > 0 ? out->stride : W,
                                                        ctx->block_offsets, *out);
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}

extern "C" int scgrhc_global_minmax(scgrhc_ctx* ctx, const double* minmax, const uint8_t* keep, int64_t n_cand,
                                    double* mm_out, void* stream) {
  NvtxRange nvtx_range("scgrhc_global_minmax");
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (!mm_out || n_cand < 0 || (n_cand && (!minmax || !keep))) return fail(ctx, SCGRHC_ERR_BAD_ARG, "global_minmax: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int blocks = (int)std::min<long long>(GMM_BLOCKS, std::max<long long>(1, (n_cand + 255) / 256));
  minmax_partial_kernel<<<blocks, 256, 0, st>>>(minmax, keep, n_cand, ctx->mm_partial);
  minmax_final_kernel<<<1, 32, 0, st>>>(ctx->mm_partial, blocks, mm_out);
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}

extern "C" int scgrhc_gather_windows(scgrhc_ctx* ctx, const void* store, const int64_t* slots, int64_t n,
                                     int64_t window_bytes, void* out, void* stream) {
  NvtxRange nvtx_range("scgrhc_gather_windows");
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (n < 0 || window_bytes <= 0 || (window_bytes & 3) || (n && (!store || !slots || !out)))
    return fail(ctx, SCGRHC_ERR_BAD_ARG, "gather: bad arguments (window_bytes must be a positive multiple of 4)");
  if (n == 0) return SCGRHC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const unsigned grid = (unsigned)std::min<long long>(n, (long long)ctx->sm_count * 8);
  const bool w8 = (window_bytes & 7) == 0 && (reinterpret_cast<uintptr_t>(store) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0;
  if (w8)
    gather_windows_kernel<uint2><<<grid, 256, 0, st>>>(static_cast<const unsigned char*>(store), reinterpret_cast<const long long*>(slots),
                                                       n, window_bytes, static_cast<unsigned char*>(out));
  else
    gather_windows_kernel<unsigned int><<<grid, 256, 0, st>>>(static_cast<const unsigned char*>(store), reinterpret_cast<const long long*>(slots),
                                                              n, window_bytes, static_cast<unsigned char*>(out));
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}

extern "C" int scgrhc_gather_windows_noise(scgrhc_ctx* ctx, const float* store, const int64_t* slots, int64_t n,
                                           int64_t window_elems, float* out, float sigma, uint64_t seed, uint64_t offset,
                                           void* stream) {
  NvtxRange nvtx_range("scgrhc_gather_windows_noise");
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (n < 0 || window_elems <= 0 || (n && (!store || !slots || !out)))
    return fail(ctx, SCGRHC_ERR_BAD_ARG, "gather_windows_noise: bad arguments");
  if (n == 0) return SCGRHC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const long long quads = (n * window_elems + 3) / 4;
  const unsigned grid = (unsigned)std::min<long long>((quads + 255) / 256, (long long)ctx->sm_count * 16);
  gather_noise_kernel<<<grid, 256, 0, st>>>(store, reinterpret_cast<const long long*>(slots), n, window_elems, out, sigma, seed, offset);
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}

extern "C" int scgrhc_collate_batch(scgrhc_ctx* ctx, const float* scg_store, const float* rhc_store, const int64_t* slots, int64_t n,
                                    int32_t scg_elems, int32_t rhc_elems, float* scg_out, float* rhc_out, float sigma,
                                    uint64_t seed, uint64_t offset, void* stream) {
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (n < 0 || scg_elems <= 0 || rhc_elems <= 0 || n > INT32_MAX || (n && (!scg_store || !rhc_store || !slots || !scg_out || !rhc_out)))
    return fail(ctx, SCGRHC_ERR_BAD_ARG, "collate_batch: bad arguments");
  if (n == 0) return SCGRHC_OK;
  CollateParams P;
  P.scg_store = scg_store; P.rhc_store = rhc_store; P.slots = reinterpret_cast<const long long*>(slots);
  P.scg_out = scg_out; P.rhc_out = rhc_out; P.n = n; P.scg_elems = scg_elems; P.rhc_elems = rhc_elems;
  P.sigma = sigma; P.seed = seed; P.offset = offset;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // no cudaSetDevice here: this is the per-batch call of the train loop; the launch goes to the stream's device
  const int split = (int)std::min<long long>(8, std::max<long long>(1, (long long)ctx->sm_count * 4 / n));
  dim3 grid((unsigned)n, (unsigned)split);
  if (sigma > 0.0f) collate_batch_kernel<true><<<grid, 256, 0, st>>>(P);
  else collate_batch_kernel<false><<<grid, 256, 0, st>>>(P);
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}

extern "C" int scgrhc_philox_words(scgrhc_ctx* ctx, uint64_t seed, uint64_t offset, int64_t nquads, uint32_t* out, void* stream) {
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (nquads < 0 || (nquads && !out)) return fail(ctx, SCGRHC_ERR_BAD_ARG, "philox_words: bad arguments");
  if (nquads == 0) return SCGRHC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  philox_words_kernel<<<(unsigned)((nquads + 255) / 256), 256, 0, st>>>(seed, offset, nquads, reinterpret_cast<uint4*>(out));
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}

extern "C" int scgrhc_window_metrics(scgrhc_ctx* ctx, const float* real, const float* pred, const double* minmax, int64_t n,
                                     int32_t W, double* out, void* stream) {
  NvtxRange nvtx_range("scgrhc_window_metrics");
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (n < 0 || W < 2 || (n && (!real || !pred || !minmax || !out))) return fail(ctx, SCGRHC_ERR_BAD_ARG, "window_metrics: bad arguments");
  if (n == 0) return SCGRHC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const unsigned grid = (unsigned)std::min<long long>((n + 7) / 8, (long long)ctx->sm_count * 8);
  window_metrics_kernel<<<grid, 256, 0, st>>>(real, pred, minmax, n, W, out);
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}

extern "C" int scgrhc_sosfiltfilt(scgrhc_ctx* ctx, const double* x, double* y, double* tmp, const int64_t* row0_dev,
                                  const int64_t* row0_host, int32_t n_rec, int32_t ncols, const int32_t* fcols, int32_t ncf,
                                  const double* sos, const double* zi, int32_t nsec, int32_t edge, void* stream) {
  NvtxRange nvtx_range("scgrhc_sosfiltfilt");
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (n_rec < 0 || ncols < 1 || ncf < 1 || ncf > kMaxFilterCols || nsec < 1 || nsec > kMaxSections || edge < 0 || !sos || !zi || !fcols ||
      (n_rec && (!x || !y || !tmp || !row0_dev || !row0_host)))
    return fail(ctx, SCGRHC_ERR_BAD_ARG, "sosfiltfilt: bad arguments (1..%d sections, 1..%d filtered columns)", kMaxSections, kMaxFilterCols);
  if (n_rec == 0) return SCGRHC_OK;
  SosParams P;
  P.x = x; P.y = y; P.tmp = tmp; P.row0 = reinterpret_cast<const long long*>(row0_dev);
  P.n_rec = n_rec; P.ncols = ncols; P.nsec = nsec; P.edge = edge; P.ncf = ncf;
  for (int r = 0; r < n_rec; ++r)
    if (row0_host[r + 1] - row0_host[r] <= edge)   // scipy: "The length of the input vector x must be greater than padlen"
      return fail(ctx, SCGRHC_ERR_BAD_ARG, "The length of the input vector x must be greater than padlen, which is %d.", edge);
  for (int j = 0; j < ncf; ++j) {
    if (fcols[j] < 0 || fcols[j] >= ncols) return fail(ctx, SCGRHC_ERR_MISSING_CHANNEL, "sosfiltfilt: column %d outside 0..%d", fcols[j], ncols - 1);
    P.fcols[j] = fcols[j];
  }
  for (int s = 0; s < nsec; ++s) {
    if (sos[6 * s + 3] != 1.0) return fail(ctx, SCGRHC_ERR_BAD_ARG, "sos[:, 3] should be all ones");
    for (int k = 0; k < 6; ++k) P.sos[s][k] = sos[6 * s + k];
    P.zi[s][0] = zi[2 * s]; P.zi[s][1] = zi[2 * s + 1];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  // A warp's recurrence needs ~45 cycles per sample (4 dependent fp64 ops) and occupies the fp64 pipe for 18 whatever
  // the number of live lanes: beyond ~2.5 warps per sub-partition more warps only queue on the pipe.  So: as few
  // columns per warp as keeps the warp count under 10 per SM.
  int cpw = std::min(32 / nsec, (int)ncf);
  while (cpw > 1 && (long long)n_rec * ((ncf + cpw - 2) / (cpw - 1)) <= (long long)ctx->sm_count * 10) --cpw;
  P.cpw = cpw;
  for (int r = 0; r < n_rec; ++r)
    if (row0_host[r + 1] - row0_host[r] + 2LL * edge >= INT32_MAX) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "sosfiltfilt: record too long");
  const int groups = (ncf + cpw - 1) / cpw;
  const long long warps = (long long)n_rec * groups;
  const unsigned grid = (unsigned)((warps + 3) / 4);
  const size_t smem = (size_t)4 * (kPrefetch + nsec + 1) * cpw * (kBlock + 1) * sizeof(double);   // 4 warps per CTA
  CUDA_TRY(ctx, cudaFuncSetAttribute(sosfilt_pass_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_TRY(ctx, cudaFuncSetAttribute(sosfilt_pass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sosfilt_pass_kernel<0><<<grid, 128, smem, st>>>(P);
  sosfilt_pass_kernel<1><<<grid, 128, smem, st>>>(P);
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}

// host tables of the time-parallel filter: A = the cascade's state matrix (column i = one zero-input step from the
// i-th unit state), B = one unit-input step from the zero state; G[i] = A^(L-1-i) B, Mp[k] = A^(L 2^k).  long double
// products, rounded once.
static void tp_tables(const double* sos, int nsec, int L, SosTpParams& P) {
  typedef long double ld;
  const int D = 2 * nsec;
  auto step = [&](ld* z, ld u) {
    for (int s = 0; s < nsec; ++s) {
      const ld b0 = sos[6 * s], b1 = sos[6 * s + 1], b2 = sos[6 * s + 2], a1 = sos[6 * s + 4], a2 = sos[6 * s + 5];
      const ld xn = b0 * u + z[2 * s];
      z[2 * s] = b1 * u - a1 * xn + z[2 * s + 1];
      z[2 * s + 1] = b2 * u - a2 * xn;
      u = xn;
    }
  };
  ld A[8][8] = {}, B[8] = {}, z[8];
  for (int i = 0; i < D; ++i) {
    for (int a = 0; a < D; ++a) z[a] = a == i ? 1.0L : 0.0L;
    step(z, 0.0L);
    for (int a = 0; a < D; ++a) A[a][i] = z[a];
  }
  for (int a = 0; a < D; ++a) z[a] = 0.0L;
  step(z, 1.0L);
  for (int a = 0; a < D; ++a) B[a] = z[a];
  auto matmul = [&](const ld X[8][8], const ld Y[8][8], ld Z[8][8]) {
    ld R[8][8];
    for (int a = 0; a < D; ++a)
      for (int b = 0; b < D; ++b) {
        ld acc = 0.0L;
        for (int k = 0; k < D; ++k) acc += X[a][k] * Y[k][b];
        R[a][b] = acc;
      }
    for (int a = 0; a < D; ++a)
      for (int b = 0; b < D; ++b) Z[a][b] = R[a][b];
  };
  memset(P.G, 0, sizeof P.G);
  memset(P.Mp, 0, sizeof P.Mp);
  ld g[8];
  for (int a = 0; a < D; ++a) g[a] = B[a];
  for (int i = L - 1; i >= 0; --i) {                       // G[L-1] = B, G[i] = A G[i+1]
    for (int a = 0; a < D; ++a) P.G[i][a] = (double)g[a];
    ld n[8];
    for (int a = 0; a < D; ++a) {
      ld acc = 0.0L;
      for (int k = 0; k < D; ++k) acc += A[a][k] * g[k];
      n[a] = acc;
    }
    for (int a = 0; a < D; ++a) g[a] = n[a];
  }
  ld M[8][8] = {};
  for (int a = 0; a < D; ++a) M[a][a] = 1.0L;
  for (int i = 0; i < L; ++i) matmul(A, M, M);            // A^L
  for (int k = 0; k < 5; ++k) {
    for (int a = 0; a < D; ++a)
      for (int b = 0; b < D; ++b) P.Mp[k][a][b] = (double)M[a][b];
    matmul(M, M, M);
  }
}

template <int NSEC, int LT>
static int sos_tp_launch_l(scgrhc_ctx* ctx, const SosTpParams& P, int ncf, size_t smem, cudaStream_t st) {
  CUDA_TRY(ctx, cudaFuncSetAttribute(sosfilt_tp_kernel<NSEC, LT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sosfilt_tp_kernel<NSEC, LT><<<(unsigned)P.n_rec, 32 * ncf, smem, st>>>(P);
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}
template <int NSEC>
static int sos_tp_launch(scgrhc_ctx* ctx, const SosTpParams& P, int ncf, size_t smem, cudaStream_t st) {
  if (P.L == 24) return sos_tp_launch_l<NSEC, 24>(ctx, P, ncf, smem, st);   // the default for 32-byte rows: unrolled
  return sos_tp_launch_l<NSEC, 0>(ctx, P, ncf, smem + kTpMaxL * 2 * NSEC * 8, st);
}

extern "C" int scgrhc_sosfiltfilt_scan(scgrhc_ctx* ctx, const double* x, double* y, const int64_t* row0_dev,
                                       const int64_t* row0_host, int32_t n_rec, int32_t ncols, const int32_t* fcols,
                                       int32_t ncf, const double* sos, const double* zi, int32_t nsec, int32_t edge,
                                       int32_t chunk, int32_t nbuf, void* stream) {
  NvtxRange nvtx_range("scgrhc_sosfiltfilt_scan");
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (n_rec < 0 || ncols < 1 || ncf < 1 || ncf > kTpMaxCols || nsec < 1 || nsec > kTpMaxSec || edge < 1 || edge > kTpMaxEdge ||
      chunk < 0 || chunk > kTpMaxL || nbuf < 0 || nbuf > 2 || !sos || !zi || !fcols || (n_rec && (!x || !y || !row0_dev || !row0_host)))
    return fail(ctx, SCGRHC_ERR_BAD_ARG, "sosfiltfilt_scan: bad arguments (1..%d sections, 1..%d filtered columns, edge 1..%d, chunk 0..%d)",
                kTpMaxSec, kTpMaxCols, kTpMaxEdge, kTpMaxL);
  if (n_rec == 0) return SCGRHC_OK;
  if (ncols > SCGRHC_MAX_NSIG) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "sosfiltfilt_scan: at most %d signals per row", SCGRHC_MAX_NSIG);
  SosTpParams P;
  memset(&P, 0, sizeof P);
  P.x = x; P.y = y; P.row0 = reinterpret_cast<const long long*>(row0_dev);
  P.n_rec = n_rec; P.ncols = ncols; P.edge = edge;
  for (int r = 0; r < n_rec; ++r) {
    const long long T = row0_host[r + 1] - row0_host[r];
    if (T <= edge) return fail(ctx, SCGRHC_ERR_BAD_ARG, "The length of the input vector x must be greater than padlen, which is %d.", edge);
    if (T + 2LL * edge >= INT32_MAX) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "sosfiltfilt_scan: record too long");
  }
  for (int j = 0; j < ncf; ++j) {
    if (fcols[j] < 0 || fcols[j] >= ncols) return fail(ctx, SCGRHC_ERR_MISSING_CHANNEL, "sosfiltfilt_scan: column %d outside 0..%d", fcols[j], ncols - 1);
    P.fcols[j] = fcols[j];
  }
  for (int s = 0; s < nsec; ++s) {
    if (sos[6 * s + 3] != 1.0) return fail(ctx, SCGRHC_ERR_BAD_ARG, "sos[:, 3] should be all ones");
    P.c[s][0] = sos[6 * s]; P.c[s][1] = sos[6 * s + 1]; P.c[s][2] = sos[6 * s + 2];
    P.c[s][3] = -sos[6 * s + 4]; P.c[s][4] = -sos[6 * s + 5];
    P.zi[s][0] = zi[2 * s]; P.zi[s][1] = zi[2 * s + 1];
  }
  // Shape of the staging: one buffer per CTA and as many rows per lane as keep it near 25 KB of shared memory, i.e. 8
  // CTAs per SM (32-byte rows: L = 24); measured faster than two buffers of half the chunk length
  const int rb = ncols * 8;
  P.nbuf = nbuf ? nbuf : 1;
  int L = chunk;
  if (!L) {
    L = (25600 / (P.nbuf * 32) - 16) / rb;
    if (L >= 4) L &= ~3;                                   // L * row bytes a multiple of 128: conflict-free lane pitch
    L = std::max(2, std::min(kTpMaxL, L));
  }
  P.L = L;
  P.bulk = (ncols % 2 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0);
  const size_t smem = 16 + 5 * (2 * nsec) * (2 * nsec) * 8 + (size_t)P.nbuf * 32 * ((size_t)L * rb + 16);
  if (smem + kTpMaxL * 2 * kTpMaxSec * 8 > 220 * 1024) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "sosfiltfilt_scan: chunk %d x %d signals does not fit in shared memory", L, ncols);
  tp_tables(sos, nsec, L, P);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  switch (nsec) {
    case 1: return sos_tp_launch<1>(ctx, P, ncf, smem, st);
    case 2: return sos_tp_launch<2>(ctx, P, ncf, smem, st);
    case 3: return sos_tp_launch<3>(ctx, P, ncf, smem, st);
    default: return sos_tp_launch<4>(ctx, P, ncf, smem, st);
  }
}

extern "C" int scgrhc_resample_poly(scgrhc_ctx* ctx, const double* x, double* y, const double* taps_dev, const int64_t* in0_dev,
                                    const int64_t* out0_dev, int32_t n_rec, int64_t max_out_rows, int32_t ncols, int32_t up, int32_t down,
                                    int32_t per_phase, int32_t n_pre_remove, int32_t fused, void* stream) {
  NvtxRange nvtx_range("scgrhc_resample_poly");
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (n_rec < 0 || ncols < 1 || up < 1 || down < 1 || per_phase < 1 || n_pre_remove < 0 || max_out_rows < 0 ||
      (n_rec && (!x || !y || !taps_dev || !in0_dev || !out0_dev)))
    return fail(ctx, SCGRHC_ERR_BAD_ARG, "resample_poly: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (up == 1 && ncols <= 5 && n_rec > 0 && n_rec <= 65535 && max_out_rows > 0) {   // integer decimation: register-blocked kernel
    const int PB = kDecR * down;
    const int rows = kDecNT * PB + per_phase;                 // input rows a tile of kDecNT * kDecR outputs touches
    const size_t dsmem = ((((size_t)per_phase + 1) & ~size_t(1)) + (size_t)(rows + rows / PB + 2) * ncols) * sizeof(double);
    if (dsmem <= 200 * 1024 && (long long)rows * PB < (1LL << 31)) {
      ResampleParams P;
      P.x = x; P.y = y; P.taps = taps_dev; P.in0 = reinterpret_cast<const long long*>(in0_dev);
      P.out0 = reinterpret_cast<const long long*>(out0_dev);
      P.n_rec = n_rec; P.ncols = ncols; P.up = up; P.down = down; P.per_phase = per_phase; P.n_pre_remove = n_pre_remove;
      P.tile_rows = rows;
      const unsigned magic = (unsigned)(((1ULL << 32) + PB - 1) / PB);   // r / PB == umulhi(r, magic) while r * PB < 2^32
      CUDA_TRY(ctx, cudaSetDevice(ctx->device));
      const int tile_out = kDecNT * kDecR;
      dim3 grid((unsigned)std::min<long long>((max_out_rows + tile_out - 1) / tile_out, 1024), (unsigned)n_rec);
      auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem);
        if (e != cudaSuccess) return e;
        kern<<<grid, kDecNT, dsmem, st>>>(P, magic);
        return cudaSuccess;
      };
      if (fused) {
        switch (ncols) {
          case 1: CUDA_TRY(ctx, launch(resample_decim_kernel<1, true>)); break;
          case 2: CUDA_TRY(ctx, launch(resample_decim_kernel<2, true>)); break;
          case 3: CUDA_TRY(ctx, launch(resample_decim_kernel<3, true>)); break;
          case 4: CUDA_TRY(ctx, launch(resample_decim_kernel<4, true>)); break;
          default: CUDA_TRY(ctx, launch(resample_decim_kernel<5, true>)); break;
        }
      } else {
        switch (ncols) {
          case 1: CUDA_TRY(ctx, launch(resample_decim_kernel<1, false>)); break;
          case 2: CUDA_TRY(ctx, launch(resample_decim_kernel<2, false>)); break;
          case 3: CUDA_TRY(ctx, launch(resample_decim_kernel<3, false>)); break;
          case 4: CUDA_TRY(ctx, launch(resample_decim_kernel<4, false>)); break;
          default: CUDA_TRY(ctx, launch(resample_decim_kernel<5, false>)); break;
        }
      }
      CUDA_TRY(ctx, cudaGetLastError());
      return SCGRHC_OK;
    }
  }
  const int tile_rows = (int)(((long long)(kResTile - 1) * down) / up) + 2 + per_phase;
  const size_t smem = ((((size_t)up * per_phase + 1) & ~size_t(1)) + (size_t)(tile_rows | 1) * ncols) * sizeof(double);
  if (smem > 200 * 1024) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "resample_poly: %zu bytes of taps + staged rows do not fit in shared memory", smem);
  if (n_rec == 0 || max_out_rows == 0) return SCGRHC_OK;
  if (n_rec > 65535) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "resample_poly: at most 65535 records per call");
  ResampleParams P;
  P.x = x; P.y = y; P.taps = taps_dev; P.in0 = reinterpret_cast<const long long*>(in0_dev);
  P.out0 = reinterpret_cast<const long long*>(out0_dev);
  P.n_rec = n_rec; P.ncols = ncols; P.up = up; P.down = down; P.per_phase = per_phase; P.n_pre_remove = n_pre_remove;
  P.tile_rows = tile_rows;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  dim3 grid((unsigned)std::min<long long>((max_out_rows + kResTile - 1) / kResTile, 1024), (unsigned)n_rec);
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kResTile, smem, st>>>(P);
    return cudaSuccess;
  };
  switch (ncols) {
    case 1: CUDA_TRY(ctx, launch(resample_poly_kernel<1>)); break;
    case 2: CUDA_TRY(ctx, launch(resample_poly_kernel<2>)); break;
    case 3: CUDA_TRY(ctx, launch(resample_poly_kernel<3>)); break;
    case 4: CUDA_TRY(ctx, launch(resample_poly_kernel<4>)); break;
    case 5: CUDA_TRY(ctx, launch(resample_poly_kernel<5>)); break;
    default: CUDA_TRY(ctx, launch(resample_poly_kernel<0>)); break;
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}

extern "C" int scgrhc_rolling_range_lt(scgrhc_ctx* ctx, const double* y, int64_t n, int32_t m, double threshold,
                                       uint8_t* flags, void* stream) {
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (n < 0 || m < 1 || (n && (!y || !flags))) return fail(ctx, SCGRHC_ERR_BAD_ARG, "rolling_range_lt: bad arguments");
  if (n == 0) return SCGRHC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  rolling_range_lt_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(y, n, m, threshold, flags);
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}

extern "C" int scgrhc_decode_fmt16(scgrhc_ctx* ctx, const int16_t* d, int64_t T, int32_t nsig_in, const int32_t* cols,
                                   int32_t ncols, const double* gain, const double* baseline, double* out, void* stream) {
  NvtxRange nvtx_range("scgrhc_decode_fmt16");
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (T < 0 || nsig_in < 1 || ncols < 1 || ncols > SCGRHC_MAX_C + 1 || !cols || !gain || !baseline || (T && (!d || !out)))
    return fail(ctx, SCGRHC_ERR_BAD_ARG, "decode_fmt16: bad arguments (1..%d output columns)", SCGRHC_MAX_C + 1);
  DecodeParams P;
  P.d = d; P.out = out; P.T = T; P.nsig_in = nsig_in; P.ncols = ncols; P.plane = ctx->out_plane;
  if (P.plane && P.plane < T) return fail(ctx, SCGRHC_ERR_BAD_ARG, "decode_fmt16: output plane stride %lld < %lld rows", (long long)P.plane, (long long)T);
  for (int j = 0; j < ncols; ++j) {
    if (cols[j] < 0 || cols[j] >= nsig_in) return fail(ctx, SCGRHC_ERR_MISSING_CHANNEL, "decode_fmt16: column %d outside 0..%d", cols[j], nsig_in - 1);
    if (gain[j] == 0.0) return fail(ctx, SCGRHC_ERR_BAD_ARG, "decode_fmt16: zero gain");
    P.cols[j] = cols[j]; P.gain[j] = gain[j]; P.baseline[j] = baseline[j];
  }
  if (T == 0) return SCGRHC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  bool fast = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  for (int j = 0; j < ncols; ++j)
    fast = fast && std::fabs(gain[j]) >= 0x1p-40 && std::fabs(gain[j]) <= 0x1p60 && std::fabs(baseline[j]) < 65536.0 &&
           baseline[j] == std::floor(baseline[j]);             // wfdb baselines are integers: d - baseline is exact
  if (fast) {
    const unsigned rgrid = (unsigned)std::min<long long>((T + 255) / 256, (long long)ctx->sm_count * 16);
    switch (ncols) {
      case 1: decode_fmt16_rows_kernel<1><<<rgrid, 256, 0, st>>>(P); break;
      case 2: decode_fmt16_rows_kernel<2><<<rgrid, 256, 0, st>>>(P); break;
      case 3: decode_fmt16_rows_kernel<3><<<rgrid, 256, 0, st>>>(P); break;
      case 4: decode_fmt16_rows_kernel<4><<<rgrid, 256, 0, st>>>(P); break;
      default: decode_fmt16_rows_kernel<5><<<rgrid, 256, 0, st>>>(P); break;
    }
  } else {
    const long long total = T * ncols;
    const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 16);
    decode_fmt16_kernel<<<grid, 256, 0, st>>>(P);
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}

extern "C" int scgrhc_decode_fmt16_records(scgrhc_ctx* ctx, const int16_t* d, const int64_t* rec_row0_dev, int32_t n_rec,
                                           int64_t max_rec_rows, int32_t nsig_in, const int32_t* cols, int32_t ncols,
                                           const double* gain_dev, const double* baseline_dev, int32_t recip, double* out,
                                           void* stream) {
  NvtxRange nvtx_range("scgrhc_decode_fmt16_records");
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (n_rec < 0 || max_rec_rows < 0 || nsig_in < 1 || ncols < 1 || ncols > SCGRHC_MAX_C + 1 || !cols ||
      (n_rec && (!d || !out || !rec_row0_dev || !gain_dev || !baseline_dev)))
    return fail(ctx, SCGRHC_ERR_BAD_ARG, "decode_fmt16_records: bad arguments (1..%d output columns)", SCGRHC_MAX_C + 1);
  if (n_rec > 65535) return fail(ctx, SCGRHC_ERR_UNSUPPORTED, "decode_fmt16_records: at most 65535 records per call");
  DecodeRecParams P;
  P.d = d; P.out = out; P.rec_row0 = reinterpret_cast<const long long*>(rec_row0_dev); P.gain = gain_dev; P.baseline = baseline_dev;
  P.n_rec = n_rec; P.nsig_in = nsig_in; P.ncols = ncols; P.recip = recip ? 1 : 0; P.plane = ctx->out_plane;
  for (int j = 0; j < ncols; ++j) {
    if (cols[j] < 0 || cols[j] >= nsig_in) return fail(ctx, SCGRHC_ERR_MISSING_CHANNEL, "decode_fmt16_records: column %d outside 0..%d", cols[j], nsig_in - 1);
    P.cols[j] = cols[j];
  }
  if (n_rec == 0 || max_rec_rows == 0) return SCGRHC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  // enough CTAs per record to fill the GPU when the chunk holds few records, no more than the record has frames for
  const long long per_rec = std::max<long long>(1, ((long long)ctx->sm_count * 16 + n_rec - 1) / n_rec);
  dim3 grid((unsigned)std::min<long long>((max_rec_rows + 255) / 256, per_rec), (unsigned)n_rec);
  switch (ncols) {
    case 1: decode_fmt16_records_kernel<1><<<grid, 256, 0, st>>>(P); break;
    case 2: decode_fmt16_records_kernel<2><<<grid, 256, 0, st>>>(P); break;
    case 3: decode_fmt16_records_kernel<3><<<grid, 256, 0, st>>>(P); break;
    case 4: decode_fmt16_records_kernel<4><<<grid, 256, 0, st>>>(P); break;
    default: decode_fmt16_records_kernel<5><<<grid, 256, 0, st>>>(P); break;
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}

extern "C" int scgrhc_waveform_stats(scgrhc_ctx* ctx, const double* y, int64_t n_wave, int64_t L, double min_rhc,
                                     double* stats, void* stream) {
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (n_wave < 0 || L < 1 || (n_wave && (!y || !stats))) return fail(ctx, SCGRHC_ERR_BAD_ARG, "waveform_stats: bad arguments");
  if (n_wave == 0) return SCGRHC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const unsigned grid = (unsigned)std::min<long long>(n_wave, (long long)ctx->sm_count * 8);
  waveform_stats_kernel<<<grid, 256, 0, st>>>(y, n_wave, L, min_rhc, stats);
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}

extern "C" int scgrhc_synth_records(scgrhc_ctx* ctx, uint64_t seed, int64_t rec0, int64_t n_rec, int64_t T, int32_t nsig,
                                    const int32_t* kinds, int32_t defect_scale, int32_t grid, double* out, void* stream) {
  NvtxRange nvtx_range("scgrhc_synth_records");
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (n_rec < 0 || T < 0 || nsig < 1 || nsig > SCGRHC_MAX_NSIG || !kinds || grid < 1 || (n_rec * T && !out))
    return fail(ctx, SCGRHC_ERR_BAD_ARG, "synth: bad arguments");
  if (n_rec * T == 0) return SCGRHC_OK;
  SynthParams P;
  P.seed = seed; P.rec0 = rec0; P.n_rec = n_rec; P.T = T; P.nsig = nsig; P.defect_scale = defect_scale; P.grid = grid;
  for (int i = 0; i < SCGRHC_MAX_NSIG; ++i) P.kinds[i] = i < nsig ? kinds[i] : 0;
  P.out = out;
  P.plane = ctx->out_plane;
  if (P.plane && P.plane < n_rec * T) return fail(ctx, SCGRHC_ERR_BAD_ARG, "synth: output plane stride smaller than the cohort");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const long long total = n_rec * T * nsig;
  const unsigned gridDim = (unsigned)std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 16);
  synth_kernel<<<gridDim, 256, 0, st>>>(P);
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}

extern "C" int scgrhc_selftest_div(scgrhc_ctx* ctx, uint64_t seed, int64_t n, int32_t mode, uint64_t* counts, void* stream) {
  if (!ctx) return SCGRHC_ERR_BAD_ARG;
  if (n < 0 || !counts) return fail(ctx, SCGRHC_ERR_BAD_ARG, "selftest_div: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CUDA_TRY(ctx, cudaMemsetAsync(counts, 0, 3 * sizeof(uint64_t), st));
  if (n == 0) return SCGRHC_OK;
  selftest_div_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(seed, n, mode, reinterpret_cast<unsigned long long*>(counts));
  CUDA_TRY(ctx, cudaGetLastError());
  return SCGRHC_OK;
}
