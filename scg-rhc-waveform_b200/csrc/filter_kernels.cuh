// Extension (named by the project brief, ABSENT from the reference; default off): zero-phase IIR filtering with scipy's
// sosfiltfilt semantics (scipy/signal/_signaltools.py: sosfiltfilt, _validate_pad, sosfilt_zi; _arraytools.py: odd_ext;
// _sosfilt.pyx: direct-form-II-transposed cascade).  Oracle = scipy itself ("parity unpinned by the reference").
//
// Every (record, column) is an independent sequence.  The recurrence is serial in time, so parallelism comes from
//   * sequences: one warp per (record, group of columns),
//   * sections: a systolic pipeline across lanes — lane (column c, section s) handles sample it-s at iteration `it`
//     and hands its output to lane s+1 with one shuffle,
// and every operation is rounded exactly as scipy's compiled loop does (separate multiply/add, no FMA), which makes the
// result bit-identical to scipy instead of merely close.  A chunked parallel scan over time would be faster but changes
// the rounding order.
#pragma once
#include "common.cuh"

namespace scgrhc {

constexpr int kMaxSections = 8;
constexpr int kMaxFilterCols = 8;
constexpr int kPrefetch = 8;

struct SosParams {
  const double* x;        // (rows, ncols) arena
  double* y;              // (rows, ncols) output arena (unfiltered columns are the caller's business)
  double* tmp;            // forward-pass output: per record (T + 2 edge) rows x ncf columns
  const long long* row0;  // device, n_rec + 1 record boundaries (rows)
  int n_rec, ncols, nsec, edge, ncf;
  int fcols[kMaxFilterCols];
  double sos[kMaxSections][6];
  double zi[kMaxSections][2];
};

template <int PASS>  // 0: forward over the odd extension, x -> tmp; 1: backward over reversed tmp -> y (trimmed)
__global__ void __launch_bounds__(128) sosfilt_pass_kernel(const __grid_constant__ SosParams P) {
  const int lane = threadIdx.x & 31;
  const int cpw = 32 / P.nsec;                              // columns per warp
  const int groups = (P.ncf + cpw - 1) / cpw;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= (long long)P.n_rec * groups) return;
  const int rec = (int)(warp / groups), grp = (int)(warp % groups);
  const int ci = lane / P.nsec, s = lane - ci * P.nsec;
  const int j = grp * cpw + ci;                             // filtered-column slot
  const bool live = ci < cpw && j < P.ncf;
  const int col = live ? P.fcols[j] : 0;
  const long long r0 = P.row0[rec], T = P.row0[rec + 1] - r0;
  const long long Lext = T + 2LL * P.edge;
  const long long tbase = (r0 + 2LL * P.edge * rec) * P.ncf;   // this record's rows in tmp
  const double b0 = P.sos[s][0], b1 = P.sos[s][1], b2 = P.sos[s][2], a1 = P.sos[s][4], a2 = P.sos[s][5];

  auto X = [&](long long t) { return P.x[(r0 + t) * P.ncols + col]; };
  auto input = [&](long long e) -> double {                  // sample e of this pass's input sequence
    if (PASS == 0) {
      if (e < P.edge) return __dsub_rn(2.0 * X(0), X(P.edge - e));                      // 2*x[0] - x[edge:0:-1]
      if (e < P.edge + T) return X(e - P.edge);
      return __dsub_rn(2.0 * X(T - 1), X(T - 2 - (e - P.edge - T)));                    // 2*x[-1] - x[-2:-(edge+2):-1]
    }
    return P.tmp[tbase + (Lext - 1 - e) * P.ncf + j];
  };

  double z0 = 0.0, z1 = 0.0;
  if (live) {
    const double first = input(0);                           // zi * x_0 (forward) / zi * y_0 (backward)
    z0 = __dmul_rn(P.zi[s][0], first);
    z1 = __dmul_rn(P.zi[s][1], first);
  }
  double cur[kPrefetch], nxt[kPrefetch];
  const bool feeder = live && s == 0;
#pragma unroll
  for (int u = 0; u < kPrefetch; ++u) cur[u] = (feeder && u < Lext) ? input(u) : 0.0;
  double x_new = 0.0;
  const long long iters = Lext + P.nsec - 1;
  for (long long base = 0; base < iters; base += kPrefetch) {
#pragma unroll
    for (int u = 0; u < kPrefetch; ++u) {
      const long long e = base + kPrefetch + u;
      nxt[u] = (feeder && e < Lext) ? input(e) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kPrefetch; ++u) {
      const long long it = base + u;
      const double from_prev = __shfl_up_sync(kFull, x_new, 1);
      const long long n = it - s;
      if (live && n >= 0 && n < Lext) {
        const double x_cur = s == 0 ? cur[u] : from_prev;
        x_new = __dadd_rn(__dmul_rn(b0, x_cur), z0);
        z0 = __dadd_rn(__dsub_rn(__dmul_rn(b1, x_cur), __dmul_rn(a1, x_new)), z1);
        z1 = __dsub_rn(__dmul_rn(b2, x_cur), __dmul_rn(a2, x_new));
        if (s == P.nsec - 1) {
          if (PASS == 0) {
            P.tmp[tbase + n * P.ncf + j] = x_new;
          } else {
            const long long t = Lext - 1 - P.edge - n;       // reverse again and trim the padding
            if (t >= 0 && t < T) P.y[(r0 + t) * P.ncols + col] = x_new;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kPrefetch; ++u) cur[u] = nxt[u];
  }
}

}  // namespace scgrhc

namespace scgrhc {

// Extension: polyphase rational resampling with scipy.signal.resample_poly / upfirdn semantics (zero padding).  The
// host passes scipy's transposed-flipped coefficient table h_trans_flip (up phases x per_phase taps, _upfirdn.py:_pad_h);
// output sample m of a record is  sum_k x[x_idx - per_phase + 1 + k] * h_trans_flip[t*per_phase + k]  with
// x_idx = (m*down)/up, t = (m*down)%up, accumulated oldest-sample-first with separately rounded multiply and add,
// exactly like the compiled loop in _upfirdn_apply.pyx, so the result is bit-identical to scipy.  Taps are staged
// in shared memory; every thread produces one output row (all columns are resampled: the rate change must keep the
// channels aligned).
struct ResampleParams {
  const double* x;         // (rows_in, ncols)
  double* y;               // (rows_out, ncols)
  const double* taps;      // device, up * per_phase
  const long long* in0;    // device, n_rec + 1
  const long long* out0;   // device, n_rec + 1
  int n_rec, ncols, up, down, per_phase, n_pre_remove;
};

__global__ void __launch_bounds__(256) resample_poly_kernel(const __grid_constant__ ResampleParams P) {
  extern __shared__ double s_taps[];
  const int ntaps = P.up * P.per_phase;
  for (int i = threadIdx.x; i < ntaps; i += blockDim.x) s_taps[i] = P.taps[i];
  __syncthreads();
  const int rec = blockIdx.y;
  const long long i0 = P.in0[rec], len_x = P.in0[rec + 1] - i0;
  const long long o0 = P.out0[rec], n_out = P.out0[rec + 1] - o0;
  for (long long mo = (long long)blockIdx.x * blockDim.x + threadIdx.x; mo < n_out; mo += (long long)gridDim.x * blockDim.x) {
    const long long m = mo + P.n_pre_remove;
    const long long x_idx = (m * P.down) / P.up;
    const int t = (int)((m * P.down) % P.up);
    const double* h = s_taps + t * P.per_phase;
    for (int c = 0; c < P.ncols; ++c) {
      double acc = 0.0;
      for (int k = 0; k < P.per_phase; ++k) {
        const long long xi = x_idx - P.per_phase + 1 + k;
        if (xi >= 0 && xi < len_x) acc = __dadd_rn(acc, __dmul_rn(P.x[(i0 + xi) * P.ncols + c], h[k]));
      }
      P.y[(o0 + mo) * P.ncols + c] = acc;
    }
  }
}

}  // namespace scgrhc
