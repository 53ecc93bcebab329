/*
 * scgrhc.h — C ABI of the B200-native SCG/RHC window-preparation library (libscgrhc.so).
 *
 * The reference (jwang6174/scg-rhc-waveform) is pure Python and has no FFI; its boundary is the
 * set of module-level Python functions other scripts import.  Each entry point
 * below names the reference interface it replaces (paths relative to the reference root).  The
 * Python drop-in modules in scg-rhc-waveform_b200/ (recordutil.py, waveform_noise.py, …) bind
 * these symbols with ctypes and keep the reference's signatures; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns an scgrhc_status (0 = ok, negative = error); the message of the last
 *     error on a context is available from scgrhc_last_error().  No C++ exception crosses the ABI.
 *   - a context is bound to one CUDA device and is not thread-safe; distinct contexts are independent.
 *   - "device" pointers are CUDA device pointers on the context's device; the library never owns
 *     record or output memory (the caller allocates, e.g. with torch.empty).
 *   - all device work is enqueued on the cudaStream_t passed as `void* stream` (NULL = legacy
 *     default stream) and is asynchronous to the host unless stated otherwise.
 */
#ifndef SCGRHC_H_
#define SCGRHC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCGRHC_ABI_VERSION 1
#define SCGRHC_MAX_C 4          /* SCG channels per window (reference configs use 1..4) */
#define SCGRHC_MAX_NSIG 64      /* signals per record row */
#define SCGRHC_SAMPLE_FREQ 500  /* recordutil.py:19 */
#define SCGRHC_FLAT_WIN 50      /* int(0.1*500), waveform_noise.py:7 */

typedef enum {
  SCGRHC_OK = 0,
  SCGRHC_ERR_BAD_ARG = -1,
  SCGRHC_ERR_MISSING_CHANNEL = -2, /* maps to ValueError from list.index, recordutil.py:117 */
  SCGRHC_ERR_NONFINITE_RHC = -3,   /* maps to sklearn's ValueError, waveform_noise.py:32 */
  SCGRHC_ERR_CUDA = -4,
  SCGRHC_ERR_UNSUPPORTED = -5,
  SCGRHC_ERR_NO_DEVICE = -6
} scgrhc_status;

/* reason bits written per candidate window (why has_noise() was true, waveform_noise.py:44-49) */
#define SCGRHC_REASON_FLAT      1u  /* get_flat_lines non-empty  (>= 2 positions, waveform_noise.py:13-26) */
#define SCGRHC_REASON_STRAIGHT  2u  /* R^2 > 0.8                 (waveform_noise.py:34) */
#define SCGRHC_REASON_FLOOR     4u  /* a sample < min_RHC        (waveform_noise.py:38-40) */
#define SCGRHC_REASON_NONFINITE 8u  /* NaN/Inf reached the regression -> reference raises */
#define SCGRHC_REASON_AMBIGUOUS 16u /* |R^2 - 0.8| < 1e-9 * R^2-scale (relative): closed form vs sklearn's lstsq could disagree in
                                       the last bits; decided by the closed form, counted by scgrhc_ambiguous_count */

/* job flags */
#define SCGRHC_OUT_F64        1u  /* write fp64 windows instead of fp32 (torch.float32 at recordutil.py:53) */
#define SCGRHC_PREDICATES_ONLY 2u /* keep/reason/minmax only, no window tensors (get_segments, pass A of global mode) */
#define SCGRHC_USE_KEPT_LIST  4u  /* iterate kept_idx[0..n_items) instead of all candidates; skip predicates;
                                     slot = list position (dense output).  Pass B of use_global_min_max. */
#define SCGRHC_NORM_GLOBAL    8u  /* normalise with job->global_minmax instead of per-window pairs (recordutil.py:58-59) */
#define SCGRHC_KEEP_ERRORS   32u  /* do not clear the context's error word first: several launches (chunks of one cohort)
                                     accumulate into one scgrhc_check_errors() */
#define SCGRHC_NORM_ZSCORE   64u  /* extension (named by the project brief, ABSENT from the reference; default off): per-window
                                     z-score instead of min-max: (x - mean) / (std + 0.0001), mean and population std taken
                                     jointly over the (W, C) SCG block and over the RHC window; the minmax output then holds
                                     {scg_mean, scg_std, rhc_mean, rhc_std}.  Not combinable with NORM_GLOBAL / USE_KEPT_LIST. */
#define SCGRHC_ARENA_PLANAR 128u /* the arena is PLANAR: column c of arena row r at arena[c * arena_rows + r] (one plane per signal over the
                                     whole arena) instead of wfdb's interleaved (rows, nsig).  The kernel then judges a candidate on its RHC
                                     plane alone and fetches the SCG planes of KEPT windows only (a rejected window costs 6 KB of DRAM traffic,
                                     not 24 KB); scgrhc_decode_fmt16* and scgrhc_synth_records can write this layout directly.  The minmax SCG
                                     pair of a rejected candidate is NaN.  Not combinable with NORM_ZSCORE; W <= 1024. */
#define SCGRHC_KEEP_ALL      16u  /* evaluate the predicates (reason bits) but keep and normalise every window:
                                     SCGDataset(segments, ...) on caller-chosen segments, no has_noise, no error */

/* One chamber interval of one record, in arena coordinates (recordutil.py:107-109,138-141). */
typedef struct {
  int64_t row0;   /* first arena row of the (clamped) interval: record base row + slice start */
  int64_t cand0;  /* index of its first candidate window = exclusive prefix sum of n_win */
  int32_t n_win;  /* L // W, recordutil.py:141 (stride extension: (L - W) // stride + 1) */
  int32_t rec_id; /* caller's record number, reported back per kept window */
} scgrhc_interval;

typedef struct {
  /* input records: (arena_rows, nsig) row-major fp64, like wfdb's p_signal (recordutil.py:118) */
  const double* arena;          /* device */
  int64_t arena_rows;
  int64_t arena_capacity_bytes; /* bytes readable from `arena` (>= arena_rows*nsig*8) */
  int32_t nsig;
  int32_t W;                    /* window length in samples = int(segment_size*500), recordutil.py:136 */
  int32_t C;                    /* number of SCG channels, len(params.in_channels) */
  int32_t scg_cols[SCGRHC_MAX_C]; /* sig_name.index(name) per in_channel, recordutil.py:117 */
  int32_t rhc_col;              /* sig_name.index('RHC_pressure'), recordutil.py:140 */
  uint32_t flags;
  const scgrhc_interval* intervals; /* device, sorted by cand0 */
  int32_t n_intervals;
  int32_t stride;               /* rows between consecutive windows of an interval; 0 = W (the reference: non-overlapping,
                                   recordutil.py:143).  Extension: overlapping windows, start_idx = i*stride */
  int64_t n_cand;               /* total candidate windows = sum n_win */
  double min_rhc;               /* params.min_RHC, waveform_noise.py:39 */
  double flat_threshold;        /* 1e-3, waveform_noise.py:6 */
  double global_minmax[4];      /* scg_min, scg_max, rhc_min, rhc_max when SCGRHC_NORM_GLOBAL */
  const int64_t* kept_list;     /* device, when SCGRHC_USE_KEPT_LIST */
  int64_t n_items;              /* length of kept_list when SCGRHC_USE_KEPT_LIST */
} scgrhc_job;

typedef struct {
  void* scg_out;     /* device (slots, C, W) fp32|fp64; slot = candidate index (or list position) */
  void* rhc_out;     /* device (slots, 1, W) */
  double* minmax;    /* device (n_cand, 4): scg_min, scg_max, rhc_min, rhc_max  (recordutil.py:58-59) */
  uint8_t* keep;     /* device (n_cand): 1 = not has_noise */
  uint8_t* reason;   /* device (n_cand): SCGRHC_REASON_* bits */
  int32_t* cand_win; /* device (n_cand): i, the window number inside its interval (start_idx = i*W, recordutil.py:143) */
  int32_t* cand_rec; /* device (n_cand): rec_id */
} scgrhc_outputs;

typedef struct {
  int64_t* kept_idx;  /* device (n_cand): candidate indices of kept windows, ascending = reference order */
  int64_t* start_idx; /* device (n_cand): per kept window, relative to its interval (recordutil.py:143) */
  int64_t* stop_idx;  /* device (n_cand): start_idx + W (recordutil.py:144) */
  int32_t* rec_id;    /* device (n_cand) */
  int64_t* n_kept;    /* device (1) */
  int32_t stride;     /* 0 = W; start_idx = i*stride */
  int32_t reserved;
} scgrhc_compact;

typedef struct scgrhc_ctx scgrhc_ctx;

/* ---- context ---------------------------------------------------------------------------- */
int scgrhc_abi_version(void);
int scgrhc_ctx_create(int device, scgrhc_ctx** out);
void scgrhc_ctx_destroy(scgrhc_ctx* ctx);
const char* scgrhc_last_error(const scgrhc_ctx* ctx); /* ctx may be NULL: last create error */
/* tuning knobs (0 = library default): CTAs per SM and bulk-copy stages per CTA of the window kernel */
int scgrhc_ctx_set_tuning(scgrhc_ctx* ctx, int ctas_per_sm, int stages);
int scgrhc_ctx_sm_count(const scgrhc_ctx* ctx);

/* ---- host planner: recordutil.get_chamber_intervals + the window count of get_segments ------
 * recordutil.py:104-109 — events in dict order with 'END' last, stable sort by time, every event
 * but the last whose key prefix matched emits (int(t*500), int(t_next*500)), truncation toward
 * zero; recordutil.py:118,141 — Python slice clamping against T rows, n_win = L // W.
 * event_time[i] in seconds (double), event_match[i] != 0 iff key.split('_')[0] == chamber.
 * Writes up to out_cap intervals (also the empty ones are skipped) and the raw (a,b) sample
 * bounds of every matching event to bounds[2*k] (may be NULL).  cand_base seeds cand0. */
int scgrhc_plan_record(const double* event_time, const uint8_t* event_match, int n_events,
                       int64_t T, int32_t W, int32_t stride /* 0 = W */, double fs /* <= 0: 500 Hz, recordutil.py:19 */,
                       int64_t rec_base_row, int32_t rec_id, int64_t cand_base,
                       scgrhc_interval* out, int out_cap, int* n_out, int64_t* n_cand,
                       int64_t* bounds, int bounds_cap, int* n_bounds);

/* The same for a whole cohort in one call (the per-record loop of get_segments, recordutil.py:131-132): record r owns
 * events ev_off[r] .. ev_off[r+1] and T_rows[r] arena rows, records back to back, rec_id = rec0 + r. */
int scgrhc_plan_cohort(const double* event_time, const uint8_t* event_match, const int64_t* ev_off, const int64_t* T_rows,
                       int64_t n_rec, int32_t W, int32_t stride, double fs, int32_t rec0, scgrhc_interval* out,
                       int64_t out_cap, int64_t* n_out, int64_t* n_cand);

/* ---- host side of the ingest: headers and side-cars of a chunk of records in one call (no device work) -------------
 * Replaces, for the common shape of both files, the per-record Python of wfdb.rdrecord's header read (recordutil.py:137)
 * and json.load + strptime of the side-car (recordutil.py:97-101).  status 0: fields valid; 1: not the common shape, send
 * this record through the general parsers; 2: a file is missing / unreadable.  names_match: the signal descriptions
 * equal the expected list (the caller's "same layout as record 0" test).  n_events -1: ChamEvents_in_s is not an object
 * (no intervals, recordutil.py:103). */
typedef struct scgrhc_record_scan {
  int32_t status;
  int32_t nsig;
  int64_t rows;        /* frames: min(header count, signal file size / (2 nsig)) */
  double fs;
  double duration_s;   /* MacEndTime - MacStTime with the date ignored */
  int32_t n_events;
  int32_t names_match;
} scgrhc_record_scan;

/* names_blob: n NUL-terminated record names back to back; expect_sig_blob: nsig_expect NUL-terminated descriptions.
 * Per record r: gains/baselines[r * nsig_expect + k], ev_time[r * max_events + i], ev_prefix[(r * max_events + i) * 16]
 * (key.split('_')[0], NUL padded), events in file order.  threads <= 0: up to 8. */
int scgrhc_scan_records(const char* dir, const char* names_blob, int64_t n, const char* expect_sig_blob, int32_t nsig_expect,
                        int32_t max_events, int32_t threads, scgrhc_record_scan* out, double* gains, int32_t* baselines,
                        double* ev_time, char* ev_prefix);

/* ---- the hot path: has_noise + SCGDataset.init_segments fused (recordutil.py:141-148,55-66;
 *      waveform_noise.py:6-49).  Asynchronous. */
int scgrhc_process_windows(scgrhc_ctx* ctx, const scgrhc_job* job, const scgrhc_outputs* out, void* stream);

/* ---- extension (named by the project brief, ABSENT from the reference; default off): the hot path with a DECIMATING front
 *      end.  job->arena holds the records at their native rate, (rows, 4) fp64; every candidate window of job->W rows at the
 *      MODEL rate (native / down) is produced inside the kernel by the polyphase FIR of scipy.signal.resample_poly (up == 1;
 *      taps = scipy's flipped table, every tap a separately rounded multiply and add, oldest sample first; zero padding
 *      outside the record) — bit-identical to scgrhc_resample_poly followed by scgrhc_process_windows, without the
 *      resampled cohort ever existing in HBM.  job->intervals / n_cand / stride are in model-rate rows as planned on the
 *      resampled cohort (row0 is not used); per interval: iv_in0 = arena row of its record's first native-rate row,
 *      iv_len = native-rate rows of that record, iv_rel = the interval's first model-rate row relative to its record.
 *      fp32 outputs, per-window normalisation (min-max or NORM_ZSCORE), W <= 384, 1..3 SCG channels. */
typedef struct {
  const double* taps;      /* host, per_phase doubles */
  int32_t per_phase, down, n_pre_remove;
  int32_t fused;           /* != 0: one FMA per tap (within ~1e-15 of scipy) instead of the bit-identical multiply + add */
  const int64_t* iv_in0;   /* device (n_intervals) */
  const int64_t* iv_len;   /* device (n_intervals) */
  const int64_t* iv_rel;   /* device (n_intervals) */
} scgrhc_decim;
int scgrhc_process_windows_decim(scgrhc_ctx* ctx, const scgrhc_job* job, const scgrhc_outputs* out, const scgrhc_decim* dec,
                                 void* stream);

/* ---- sweep fan-out (the loadable waveform_NN configs are 4 chambers x 8 channel subsets; for one
 *      chamber has_noise() looks at the RHC channel only, waveform_noise.py:44-49): after ONE predicate pass
 *      (SCGRHC_PREDICATES_ONLY) and scgrhc_compact_kept, normalise every kept window for up to SCGRHC_MAX_SUBSETS channel
 *      subsets in one pass: SCGDataset.init_segments (recordutil.py:55-66) per subset, the window read once.
 *      job: scg_cols[0..C) = the superset of channels (C <= SCGRHC_MAX_C), kept_list / n_items = the kept candidates,
 *      flags: only SCGRHC_OUT_F64 is looked at.  Subset k takes the superset columns member[0..C) (ascending) in that
 *      channel order; scg_out (n_items, C, W), minmax (n_items, 4) dense in list order; rhc_out (n_items, 1, W) is
 *      written once for all subsets.  Outputs are bit-identical to scgrhc_process_windows run per subset. */
#define SCGRHC_MAX_SUBSETS 8
typedef struct {
  void* scg_out;                  /* device (n_items, C, W) fp32|fp64 */
  double* minmax;                 /* device (n_items, 4) */
  int32_t C;
  int32_t member[SCGRHC_MAX_C];   /* indices into job->scg_cols, strictly ascending */
} scgrhc_subset;
int scgrhc_normalize_subsets(scgrhc_ctx* ctx, const scgrhc_job* job, const scgrhc_subset* subsets, int32_t n_subsets,
                             void* rhc_out, void* stream);

/* ---- ordered list of kept windows (the order of the list get_segments returns, recordutil.py:148) */
int scgrhc_compact_kept(scgrhc_ctx* ctx, const uint8_t* keep, const int32_t* cand_win, const int32_t* cand_rec,
                        int64_t n_cand, int32_t W, const scgrhc_compact* out, void* stream);

/* ---- get_global_minmax_vals over kept windows (recordutil.py:152-169): device reduction of the
 *      (n_cand,4) pairs to mm_out[4] (device).  Cross-GPU: allreduce MIN of {min,-max} by the caller. */
int scgrhc_global_minmax(scgrhc_ctx* ctx, const double* minmax, const uint8_t* keep, int64_t n_cand,
                         double* mm_out, void* stream);

/* ---- error word of the last process call: synchronises `stream`; returns SCGRHC_ERR_NONFINITE_RHC
 *      and the first offending candidate when a non-finite RHC sample reached the regression. */
int scgrhc_check_errors(scgrhc_ctx* ctx, void* stream, int64_t* first_bad_cand);
/* Candidate windows flagged SCGRHC_REASON_AMBIGUOUS by the process calls the last scgrhc_check_errors covered (no
 * synchronisation of its own; -1 for a NULL context).  SURVEY.md §7: the closed-form R^2 and the reference's
 * sklearn/LAPACK score agree to ~6e-16, so only these windows could be decided differently; callers warn when > 0. */
int64_t scgrhc_ambiguous_count(const scgrhc_ctx* ctx);

/* ---- batch collate (default_collate of recordutil.py:198): out[b] = store[slot[b]] ---------- */
int scgrhc_gather_windows(scgrhc_ctx* ctx, const void* store, const int64_t* slots, int64_t n,
                          int64_t window_bytes, void* out, void* stream);

/* ---- extension (named by the project brief, ABSENT from the reference; default off): train-time noise injection
 *      fused into the batch gather of fp32 windows: out[b][e] = store[slots[b]][e] + sigma * N(0,1).  Normals come from
 *      the counter-based Philox4x32-10 generator (key = seed, counter = (quad index, offset)) through Box-Muller;
 *      `offset` selects an independent stream per batch.  scgrhc_philox_words exposes the raw 4x32-bit blocks
 *      (nquads * 4 words, device, 16-byte aligned) for seed-exact checks against a host implementation. */
int scgrhc_gather_windows_noise(scgrhc_ctx* ctx, const float* store, const int64_t* slots, int64_t n,
                                int64_t window_elems, float* out, float sigma, uint64_t seed, uint64_t offset, void* stream);
/* ---- the train loop's per-batch call in ONE launch (default_collate of recordutil.py:198 for the two tensors
 *      waveform_train.py:358-359 reads): scg_out[b] = scg_store[slots[b]] (+ sigma * N(0,1) when sigma > 0, the same
 *      Philox stream as scgrhc_gather_windows_noise) and rhc_out[b] = rhc_store[slots[b]]; fp32 windows of scg_elems
 *      (= C*W) and rhc_elems (= W) floats; outputs are caller-provided (e.g. a ring of batch buffers).  The caller must
 *      have the context's device current (no cudaSetDevice per batch). */
int scgrhc_collate_batch(scgrhc_ctx* ctx, const float* scg_store, const float* rhc_store, const int64_t* slots, int64_t n,
                         int32_t scg_elems, int32_t rhc_elems, float* scg_out, float* rhc_out, float sigma, uint64_t seed,
                         uint64_t offset, void* stream);
int scgrhc_philox_words(scgrhc_ctx* ctx, uint64_t seed, uint64_t offset, int64_t nquads, uint32_t* out, void* stream);

/* ---- evaluation metrics of the consumer (waveform_test.py:21-50,66-70), per window: both fp32 waveforms (n, W) are
 *      de-normalised with the window's RHC pair minmax (n, 2) = {rhc_min, rhc_max} as reverse_minmax does
 *      (v * (max - min) + min in fp64, no epsilon); out (n, 2) = {Pearson r, RMSE}.  Confidence intervals stay on the host. */
int scgrhc_window_metrics(scgrhc_ctx* ctx, const float* real, const float* pred, const double* minmax, int64_t n,
                          int32_t W, double* out, void* stream);

/* ---- extension (named by the project brief, ABSENT from the reference; default off): zero-phase IIR filtering of
 *      the columns fcols[0..ncf) of every record of the arena, scipy.signal.sosfiltfilt semantics (odd extension by
 *      `edge` samples, lfilter_zi initial state scaled by the first sample, forward pass, reversed second pass, trim).
 *      x, y: (rows, ncols) device (y may alias nothing of x; unfiltered columns of y are not written); tmp: device,
 *      (rows + 2*edge*n_rec) * ncf doubles; row0: n_rec+1 record boundaries, on the device AND on the host; sos
 *      (nsec, 6) with a0 == 1, zi (nsec, 2) = scipy.signal.sosfilt_zi(sos), both host.  Bit-identical to scipy. */
int scgrhc_sosfiltfilt(scgrhc_ctx* ctx, const double* x, double* y, double* tmp, const int64_t* row0_dev,
                       const int64_t* row0_host, int32_t n_rec, int32_t ncols, const int32_t* fcols, int32_t ncf,
                       const double* sos, const double* zi, int32_t nsec, int32_t edge, void* stream);

/* ---- the same filter, time-parallel (the brief's "warp-level parallel linear-recurrence scan over biquad state"): one
 *      CTA per record (a warp per filtered column) walks it in spans of 32 * chunk rows; every lane filters its own chunk, the chunk-boundary
 *      states follow S_(t+1) = A^chunk S_t + f_t through a Kogge-Stone scan over the lanes (A = state matrix of the
 *      cascade, tables built on the host from sos).  Forward pass x -> y (whole rows, so the other columns are copied
 *      through), backward pass over y in place: no scratch.  y may be x itself (in place); otherwise they must not
 *      overlap.  chunk = rows per lane per span (0 = library default, <= 32), nbuf = staging buffers per CTA (0 =
 *      default 1; 2 prefetches one span ahead at half the chunk length).  Up to 4 sections and 4 filtered columns per call, edge <= 32.  Not bit-identical to scipy (FMA,
 *      different rounding order at chunk boundaries): within 1e-10 of full scale for SCG/RHC pass bands at 500 Hz. */
int scgrhc_sosfiltfilt_scan(scgrhc_ctx* ctx, const double* x, double* y, const int64_t* row0_dev,
                            const int64_t* row0_host, int32_t n_rec, int32_t ncols, const int32_t* fcols, int32_t ncf,
                            const double* sos, const double* zi, int32_t nsec, int32_t edge, int32_t chunk, int32_t nbuf,
                            void* stream);

/* ---- extension (named by the project brief, ABSENT from the reference; default off): rational resampling of every
 *      record of the arena, scipy.signal.resample_poly semantics (polyphase upfirdn, zero padding).  taps: device,
 *      scipy's transposed-flipped table (up x per_phase, scipy/signal/_upfirdn.py:_pad_h) of the up-scaled, pre-padded
 *      FIR; in0/out0: device record boundaries (n_rec+1) of x (rows_in, ncols) and y (rows_out, ncols);
 *      out rows per record = ceil(n_in*up/down); n_pre_remove as in resample_poly.  Bit-identical to scipy (every tap a
 *      separately rounded multiply and add, oldest sample first).  fused != 0 (integer decimation only; ignored otherwise):
 *      one FMA per tap instead — half the fp64 instructions, within 1e-14 of scipy instead of bit-identical. */
int scgrhc_resample_poly(scgrhc_ctx* ctx, const double* x, double* y, const double* taps_dev, const int64_t* in0_dev,
                         const int64_t* out0_dev, int32_t n_rec, int64_t max_out_rows, int32_t ncols, int32_t up, int32_t down,
                         int32_t per_phase, int32_t n_pre_remove, int32_t fused, void* stream);

/* ---- standalone predicate helpers for API parity of waveform_noise.get_flat_lines with
 *      non-default arguments: flags[p] = (rolling range over m samples ending at p) < threshold */
int scgrhc_rolling_range_lt(scgrhc_ctx* ctx, const double* y, int64_t n, int32_t m, double threshold,
                            uint8_t* flags, void* stream);

/* ---- record ingest (what wfdb.rdrecord does on the host at recordutil.py:137): WFDB format-16 digital frames
 *      d (T, nsig_in) int16, device -> physical fp64 samples out (T, ncols), device, for the selected columns:
 *      (d - baseline) / gain, the invalid code -32768 -> NaN.  gain/baseline/cols are host arrays of ncols entries. */
int scgrhc_decode_fmt16(scgrhc_ctx* ctx, const int16_t* d, int64_t T, int32_t nsig_in, const int32_t* cols,
                        int32_t ncols, const double* gain, const double* baseline, double* out, void* stream);
/* Output layout of the two decode entry points and of scgrhc_synth_records from now on: plane_stride = 0 (default) is
 * interleaved (T, ncols); plane_stride = P > 0 writes column j of row t to out[j * P + t] (SCGRHC_ARENA_PLANAR with
 * arena_rows = P).  Sticky per context; the calls themselves are unchanged. */
int scgrhc_ctx_set_output_planes(scgrhc_ctx* ctx, int64_t plane_stride);

/* The same for a chunk of n_rec records back to back, each with its own calibration (one launch per chunk instead of one
 * per record): rec_row0 (n_rec+1, device) = first frame of every record inside d / out; gain, baseline (n_rec, ncols)
 * device tables; max_rec_rows = the longest record of the chunk; recip != 0 selects the reciprocal-based quotient
 * (bit-identical to IEEE division; the caller has checked 2^-40 <= |gain| <= 2^60 and integer |baseline| < 2^16 for the
 * whole table), 0 = IEEE division. */
int scgrhc_decode_fmt16_records(scgrhc_ctx* ctx, const int16_t* d, const int64_t* rec_row0_dev, int32_t n_rec,
                                int64_t max_rec_rows, int32_t nsig_in, const int32_t* cols, int32_t ncols,
                                const double* gain_dev, const double* baseline_dev, int32_t recip, double* out, void* stream);

/* ---- is_straight_line / in_rhc_range on waveforms of any length (waveform_noise.py:29-41), one
 *      waveform per row of y (n_wave, L); stats (n_wave, 6) = {R^2, min, max, below_floor, nonfinite, sum} */
int scgrhc_waveform_stats(scgrhc_ctx* ctx, const double* y, int64_t n_wave, int64_t L, double min_rhc,
                          double* stats, void* stream);

/* ---- synthetic cohort generator: records rec0..rec0+n_rec-1 of cohort `seed`,
 *      each (T, nsig) fp64 row-major, written back to back into `out` (device). */
int scgrhc_synth_records(scgrhc_ctx* ctx, uint64_t seed, int64_t rec0, int64_t n_rec, int64_t T,
                         int32_t nsig, const int32_t* kinds, int32_t defect_scale, int32_t grid,
                         double* out, void* stream);

/* ---- diagnostics: compares the kernel's reciprocal-based division (Markstein correction) with
 *      IEEE division on n hashed operand pairs; counts (device, 3 x uint64): fp64 mismatches,
 *      mismatches after the fp32 cast, pairs evaluated.  mode 0: operands shaped like the
 *      normalisation (min <= x <= max, ranges 2^-14..2^10); mode 1: exponents up to +-1000, zeros (selects the
 *      IEEE loop); mode 2: the integer-checked fp32 tier, half of the quotients planted within 8 ulp64 of float
 *      rounding boundaries — counts = {operands not eligible, fp32 mismatches among unflagged, flagged}; mode 3: the
 *      same with minima down to 2^-86 and a third of the samples a few ulp above the minimum. */
int scgrhc_selftest_div(scgrhc_ctx* ctx, uint64_t seed, int64_t n, int32_t mode, uint64_t* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCGRHC_H_ */
