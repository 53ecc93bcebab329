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
constexpr int kTileRows = 32;   // samples per staged tile
constexpr int kRing = 8;        // tiles in flight per warp (256 samples ahead)

struct SosParams {
  const double* x;        // (rows, ncols) arena
  double* y;              // (rows, ncols) output arena (unfiltered columns are the caller's business)
  double* tmp;            // forward-pass output: per record (T + 2 edge) rows x ncf columns
  const long long* row0;  // device, n_rec + 1 record boundaries (rows)
  int n_rec, ncols, nsec, edge, ncf;
  int cpw;                // filtered columns per warp (1 .. 32 / nsec): fewer columns per warp = more warps in flight
  int fcols[kMaxFilterCols];
  double sos[kMaxSections][6];
  double zi[kMaxSections][2];
};

template <int PASS>  // 0: forward over the odd extension, x -> tmp; 1: backward over reversed tmp -> y (trimmed)
__global__ void __launch_bounds__(128) sosfilt_pass_kernel(const __grid_constant__ SosParams P) {
  const int lane = threadIdx.x & 31;
  const int cpw = P.cpw;
  const int groups = (P.ncf + cpw - 1) / cpw;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= (long long)P.n_rec * groups) return;
  const int rec = (int)(warp / groups), grp = (int)(warp % groups);
  const int ci = lane / P.nsec, s = lane - ci * P.nsec;
  const int j = grp * cpw + ci;                             // filtered-column slot
  const bool live = ci < cpw && j < P.ncf;
  const int col = live ? P.fcols[j] : 0;
  const long long r0 = P.row0[rec];
  const int T = (int)(P.row0[rec + 1] - r0);                // host guarantees T + 2 edge < 2^31
  const int Lext = T + 2 * P.edge;
  const long long tbase = (r0 + 2LL * P.edge * rec) * P.ncf;   // this record's rows in tmp
  const double b0 = P.sos[s][0], b1 = P.sos[s][1], b2 = P.sos[s][2], a1 = P.sos[s][4], a2 = P.sos[s][5];

  auto input = [&](int e, int slot, int column) -> double {   // sample e of this pass's input sequence
    if (PASS == 0) {
      const double* xc = P.x + r0 * P.ncols + column;
      if (e < P.edge) return __dsub_rn(2.0 * xc[0], xc[(long long)(P.edge - e) * P.ncols]);          // 2*x[0] - x[edge:0:-1]
      if (e < P.edge + T) return xc[(long long)(e - P.edge) * P.ncols];
      return __dsub_rn(2.0 * xc[(long long)(T - 1) * P.ncols], xc[(long long)(T - 2 - (e - P.edge - T)) * P.ncols]);
    }
    return P.tmp[tbase + (long long)(Lext - 1 - e) * P.ncf + slot];
  };

  double z0 = 0.0, z1 = 0.0;
  if (live) {
    const double first = input(0, j, col);                   // zi * x_0 (forward) / zi * y_0 (backward)
    z0 = __dmul_rn(P.zi[s][0], first);
    z1 = __dmul_rn(P.zi[s][1], first);
  }
  // Input staging: the whole warp loads tiles of kTileRows samples x (this warp's columns) into a shared-memory ring,
  // kRing tiles ahead of the sample being filtered, so the serial recurrence never waits on a global load.
  extern __shared__ double s_ring[];
  const int wcols = min(cpw, P.ncf - grp * cpw);            // filtered columns handled by this warp
  double* ring = s_ring + (size_t)(threadIdx.x >> 5) * kRing * kTileRows * cpw;
  auto load_tile = [&](int tile) {
    double* dst = ring + (size_t)(tile % kRing) * kTileRows * cpw;
    double v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {                            // issue the loads first, then the stores
      const int q = lane + 32 * i;
      const int row = q / wcols, jj = q - row * wcols, e = tile * kTileRows + row;
      v[i] = (q < kTileRows * wcols && e < Lext) ? input(e, grp * cpw + jj, P.fcols[grp * cpw + jj]) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = lane + 32 * i;
      const int row = q / wcols, jj = q - row * wcols;
      if (q < kTileRows * wcols) dst[row * cpw + jj] = v[i];
    }
    for (int q = lane + 128; q < kTileRows * wcols; q += 32) {   // more than 4 columns per warp: plain loop
      const int row = q / wcols, jj = q - row * wcols, e = tile * kTileRows + row;
      dst[row * cpw + jj] = e < Lext ? input(e, grp * cpw + jj, P.fcols[grp * cpw + jj]) : 0.0;
    }
  };
  const int ntiles = (Lext + kTileRows - 1) / kTileRows;
  for (int tl = 0; tl < kRing - 1 && tl < ntiles; ++tl) load_tile(tl);
  __syncwarp();
  double x_new = 0.0;
  const int iters = Lext + P.nsec - 1;
  const bool last = live && s == P.nsec - 1;
  // output cursor of the last section: forward -> tmp row n; backward -> y row Lext-1-edge-n (reverse + trim)
  double* outp = PASS == 0 ? P.tmp + tbase + j - (long long)s * P.ncf
                           : P.y + (r0 + (long long)(Lext - 1 - P.edge + s)) * P.ncols + col;
  const long long ostep = PASS == 0 ? (long long)P.ncf : -(long long)P.ncols;
  for (int tile = 0; tile * kTileRows < iters; ++tile) {
    if (tile + kRing - 1 < ntiles) load_tile(tile + kRing - 1);   // slot (tile-1) % kRing: its reads finished last round
    const double* src = ring + (size_t)(tile % kRing) * kTileRows * cpw + (live ? ci : 0);   // idle lanes stay in bounds
#pragma unroll 8
    for (int u = 0; u < kTileRows; ++u) {
      const int n = tile * kTileRows + u - s;
      const double from_prev = __shfl_up_sync(kFull, x_new, 1);
      const double x_cur = s == 0 ? src[u * cpw] : from_prev;
      if (live && (unsigned)n < (unsigned)Lext) {
        x_new = __dadd_rn(__dmul_rn(b0, x_cur), z0);
        z0 = __dadd_rn(__dsub_rn(__dmul_rn(b1, x_cur), __dmul_rn(a1, x_new)), z1);
        z1 = __dsub_rn(__dmul_rn(b2, x_cur), __dmul_rn(a2, x_new));
        if (last) {
          if (PASS == 0) *outp = x_new;
          else if (n >= P.edge && n < P.edge + T) *outp = x_new;
        }
      }
      outp += ostep;
    }
    __syncwarp();                                            // the tile just consumed may be overwritten next round
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
  int tile_rows;           // input rows staged per CTA: (255 * down) / up + 1 + per_phase
};

constexpr int kResTile = 256;   // output rows per CTA tile (one per thread)

// NC = compile-time column count (0: runtime, processed in predicated chunks of 8)
template <int NC>
__global__ void __launch_bounds__(kResTile) resample_poly_kernel(const __grid_constant__ ResampleParams P) {
  extern __shared__ double s_res[];
  double* s_taps = s_res;
  const int ntaps = P.up * P.per_phase;
  const int ncols = NC ? NC : P.ncols;
  const int pitch = P.tile_rows | 1;                        // column-major staging, odd pitch
  double* s_x = s_res + ((ntaps + 1) & ~1);                 // s_x[c * pitch + r]
  for (int i = threadIdx.x; i < ntaps; i += blockDim.x) s_taps[i] = P.taps[i];
  const int rec = blockIdx.y;
  const long long i0 = P.in0[rec], len_x = P.in0[rec + 1] - i0;
  const long long o0 = P.out0[rec], n_out = P.out0[rec + 1] - o0;
  for (long long tile0 = (long long)blockIdx.x * kResTile; tile0 < n_out; tile0 += (long long)gridDim.x * kResTile) {
    // input rows this tile touches: [first, first + tile_rows), zero outside the record (upfirdn's zero padding).
    // Staged column-major so that neighbouring threads (neighbouring output rows) read words `down` apart instead of
    // `down * ncols` apart: at most a `down`-way bank conflict instead of 16-way.
    const long long first = ((tile0 + P.n_pre_remove) * P.down) / P.up - P.per_phase + 1;
    __syncthreads();
    for (int q = threadIdx.x; q < P.tile_rows * ncols; q += blockDim.x) {
      const int r = q / ncols, c = q - r * ncols;
      const long long xi = first + r;
      s_x[c * pitch + r] = (xi >= 0 && xi < len_x) ? P.x[(i0 + xi) * ncols + c] : 0.0;
    }
    __syncthreads();
    const long long mo = tile0 + threadIdx.x;
    if (mo < n_out) {
      const long long m = mo + P.n_pre_remove;
      const int base = (int)((m * P.down) / P.up - P.per_phase + 1 - first);
      const double* h = s_taps + (int)((m * P.down) % P.up) * P.per_phase;
      // each column is summed oldest-sample-first with separately rounded multiply and add, as scipy's compiled loop
      // does (zero-padded rows add an exact +0); columns are independent chains and advance together
      if constexpr (NC > 0) {
        double acc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = 0.0;
        const double* col0 = s_x + base;
#pragma unroll 4
        for (int k = 0; k < P.per_phase; ++k) {
          const double hk = h[k];
#pragma unroll
          for (int c = 0; c < NC; ++c) acc[c] = __dadd_rn(acc[c], __dmul_rn(col0[c * pitch + k], hk));
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) P.y[(o0 + mo) * NC + c] = acc[c];
      } else {
        for (int c0 = 0; c0 < ncols; c0 += 8) {
          double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          const int nc = min(8, ncols - c0);
          for (int k = 0; k < P.per_phase; ++k) {
            const double hk = h[k];
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (c < nc) acc[c] = __dadd_rn(acc[c], __dmul_rn(s_x[(c0 + c) * pitch + base + k], hk));
          }
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c < nc) P.y[(o0 + mo) * ncols + c0 + c] = acc[c];
        }
      }
    }
  }
}

}  // namespace scgrhc

namespace scgrhc {

// ---- time-parallel variant of the same filter (the brief's "parallel linear-recurrence scan over biquad state") ----
// The cascade is one linear system  S' = A S + B x  (S = the 2*nsec delay elements).  A sequence is cut into chunks of
// `chunk` samples; thread (record, chunk) filters its chunk for ALL filtered columns
//   pass A: from a zero state, keeping only the final state f_k                       (sosfilt_chunk_kernel<.,0>)
//   scan  : S_{k+1} = M S_k + f_k with M = A^chunk (host, fp64), S_0 = zi * x_0        (sosfilt_scan_kernel)
//   pass B: again from the true S_k, writing the outputs                               (sosfilt_chunk_kernel<.,1>)
// Work doubles, but parallelism is records x chunks instead of records, and a thread consumes whole 32-byte rows.
// Inside a chunk every operation is rounded as in the exact kernel; only the chunk-boundary states come from the
// matrix recurrence, so the result differs from scipy by rounding noise (measured <= 1e-13 of full scale, bar 1e-10).
struct SosScanParams {
  const double* x; double* y; double* tmp;
  const long long* row0;
  double* fstate;           // (total_chunks, ncf, 2*nsec): final zero-state of pass A, then S_k for pass B (in place)
  const long long* chunk0;  // device, n_rec + 1: first chunk index of every record
  const double* M;          // device, (2*nsec)^2 row-major
  int n_rec, ncols, nsec, edge, ncf, chunk;
  int fcols[kMaxFilterCols];
  double sos[kMaxSections][6];
  double zi[kMaxSections][2];
};

template <int PASS>
__device__ __forceinline__ double sos_input(const SosScanParams& P, long long r0, int T, int Lext, long long tbase, int e, int j) {
  if (PASS == 0) {
    const double* xc = P.x + r0 * P.ncols + P.fcols[j];
    if (e < P.edge) return __dsub_rn(2.0 * xc[0], xc[(long long)(P.edge - e) * P.ncols]);
    if (e < P.edge + T) return xc[(long long)(e - P.edge) * P.ncols];
    return __dsub_rn(2.0 * xc[(long long)(T - 1) * P.ncols], xc[(long long)(T - 2 - (e - P.edge - T)) * P.ncols]);
  }
  return P.tmp[tbase + (long long)(Lext - 1 - e) * P.ncf + j];
}

// PASS: 0 forward / 1 backward (as in the exact kernel);  PHASE: 0 zero-state (final state only) / 1 true state + outputs
template <int PASS, int PHASE, int NSEC, int NCF>
__global__ void __launch_bounds__(128) sosfilt_chunk_kernel(const __grid_constant__ SosScanParams P, long long total_chunks) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total_chunks) return;
  int lo = 0, hi = P.n_rec - 1;                            // record of this chunk: last r with chunk0[r] <= gid
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (P.chunk0[mid] <= gid) lo = mid; else hi = mid - 1; }
  const int rec = lo;
  const int k = (int)(gid - P.chunk0[rec]);
  const long long r0 = P.row0[rec];
  const int T = (int)(P.row0[rec + 1] - r0), Lext = T + 2 * P.edge;
  const long long tbase = (r0 + 2LL * P.edge * rec) * P.ncf;
  const int e0 = k * P.chunk, e1 = min(Lext, e0 + P.chunk);
  double z[NCF][NSEC][2];
  double* st = P.fstate + gid * (long long)(NCF * NSEC * 2);
#pragma unroll
  for (int j = 0; j < NCF; ++j)
#pragma unroll
    for (int s = 0; s < NSEC; ++s) {
      z[j][s][0] = PHASE ? st[(j * NSEC + s) * 2] : 0.0;
      z[j][s][1] = PHASE ? st[(j * NSEC + s) * 2 + 1] : 0.0;
    }
  for (int e = e0; e < e1; ++e) {
#pragma unroll
    for (int j = 0; j < NCF; ++j) {
      double v = sos_input<PASS>(P, r0, T, Lext, tbase, e, j);
#pragma unroll
      for (int s = 0; s < NSEC; ++s) {
        const double xn = __dadd_rn(__dmul_rn(P.sos[s][0], v), z[j][s][0]);
        z[j][s][0] = __dadd_rn(__dsub_rn(__dmul_rn(P.sos[s][1], v), __dmul_rn(P.sos[s][4], xn)), z[j][s][1]);
        z[j][s][1] = __dsub_rn(__dmul_rn(P.sos[s][2], v), __dmul_rn(P.sos[s][5], xn));
        v = xn;
      }
      if (PHASE) {
        if (PASS == 0) {
          P.tmp[tbase + (long long)e * P.ncf + j] = v;
        } else {
          const int t = Lext - 1 - P.edge - e;
          if (t >= 0 && t < T) P.y[(r0 + t) * P.ncols + P.fcols[j]] = v;
        }
      }
    }
  }
  if (!PHASE) {
#pragma unroll
    for (int j = 0; j < NCF; ++j)
#pragma unroll
      for (int s = 0; s < NSEC; ++s) { st[(j * NSEC + s) * 2] = z[j][s][0]; st[(j * NSEC + s) * 2 + 1] = z[j][s][1]; }
  }
}

// one thread per (record, column): S_{k+1} = M S_k + f_k over the record's chunks, in place (f_k -> S_k)
template <int PASS, int NSEC>
__global__ void sosfilt_scan_kernel(const __grid_constant__ SosScanParams P) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= P.n_rec * P.ncf) return;
  const int rec = id / P.ncf, j = id - rec * P.ncf;
  const long long r0 = P.row0[rec];
  const int T = (int)(P.row0[rec + 1] - r0), Lext = T + 2 * P.edge;
  const long long tbase = (r0 + 2LL * P.edge * rec) * P.ncf;
  constexpr int D = 2 * NSEC;
  double S[D], Mr[D][D];
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = 0; b < D; ++b) Mr[a][b] = P.M[a * D + b];
  const double first = sos_input<PASS>(P, r0, T, Lext, tbase, 0, j);
#pragma unroll
  for (int s = 0; s < NSEC; ++s) { S[2 * s] = __dmul_rn(P.zi[s][0], first); S[2 * s + 1] = __dmul_rn(P.zi[s][1], first); }
  const long long c0 = P.chunk0[rec], c1 = P.chunk0[rec + 1];
  for (long long c = c0; c < c1; ++c) {
    double* st = P.fstate + (c * P.ncf + j) * D;
    double f[D], nx[D];
#pragma unroll
    for (int a = 0; a < D; ++a) { f[a] = st[a]; st[a] = S[a]; }
#pragma unroll
    for (int a = 0; a < D; ++a) {
      double acc = f[a];
#pragma unroll
      for (int b = 0; b < D; ++b) acc = __fma_rn(Mr[a][b], S[b], acc);
      nx[a] = acc;
    }
#pragma unroll
    for (int a = 0; a < D; ++a) S[a] = nx[a];
  }
}

}  // namespace scgrhc
