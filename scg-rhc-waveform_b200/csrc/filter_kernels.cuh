// Extension (named by the project brief, ABSENT from the reference; default off): zero-phase IIR filtering with scipy's
// sosfiltfilt semantics (scipy/signal/_signaltools.py: sosfiltfilt, _validate_pad, sosfilt_zi; _arraytools.py: odd_ext;
// _sosfilt.pyx: direct-form-II-transposed cascade).  Oracle = scipy itself ("parity unpinned by the reference").
//
// Every (record, column) is an independent sequence.  The recurrence is serial in time, so parallelism comes from
//   * sequences: one warp per (record, group of columns),
//   * sections: a block pipeline across lanes — lane (column c, section s) filters block t - s (kBlock samples) at step t,
//     in place in a shared-memory ring, and the next section picks the block up one step later (one __syncwarp per
//     block instead of one shuffle per sample),
// and every operation is rounded exactly as scipy's compiled loop does (separate multiply/add, no FMA), which makes the
// result bit-identical to scipy instead of merely close.  The time-parallel kernel (filter_scan_kernel.cuh) is an order
// of magnitude faster but changes the rounding order.
#pragma once
#include "common.cuh"

namespace scgrhc {

constexpr int kMaxSections = 8;
constexpr int kMaxFilterCols = 8;
constexpr int kBlock = 32;      // samples per block
constexpr int kPrefetch = 6;    // blocks loaded ahead of the first section

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
  const int cpw = P.cpw, nsec = P.nsec;
  const int groups = (P.ncf + cpw - 1) / cpw;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= (long long)P.n_rec * groups) return;
  const int rec = (int)(warp / groups), grp = (int)(warp % groups);
  const int ci = lane / nsec, s = lane - ci * nsec;
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
  // Ring of D = prefetch + sections + 1 blocks per column, pitch kBlock + 1 doubles (lanes of different sections /
  // columns then fall into different banks).  Step t: the block that left the last section one step ago is stored,
  // block t + prefetch is loaded, lane (c, s) filters block t - s in place.
  extern __shared__ double s_ring[];
  const int wcols = min(cpw, P.ncf - grp * cpw);            // filtered columns handled by this warp
  const int D = kPrefetch + nsec + 1;
  constexpr int pitch = kBlock + 1;
  double* ring = s_ring + (size_t)(threadIdx.x >> 5) * D * cpw * pitch;
  const int nblocks = (Lext + kBlock - 1) / kBlock;
  auto load_block = [&](int b) {
    double* dst = ring + (size_t)(b % D) * cpw * pitch;
    for (int q0 = 0; q0 < kBlock * wcols; q0 += 128) {       // loads first, then the stores
      double v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int q = q0 + lane + 32 * i;
        const int u = q / wcols, jj = q - u * wcols, e = b * kBlock + u;
        v[i] = (q < kBlock * wcols && e < Lext) ? input(e, grp * cpw + jj, P.fcols[grp * cpw + jj]) : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int q = q0 + lane + 32 * i;
        const int u = q / wcols, jj = q - u * wcols;
        if (q < kBlock * wcols) dst[jj * pitch + u] = v[i];
      }
    }
  };
  auto store_block = [&](int b) {
    const double* src = ring + (size_t)(b % D) * cpw * pitch;
    for (int q = lane; q < kBlock * wcols; q += 32) {
      const int u = q / wcols, jj = q - u * wcols, n = b * kBlock + u;
      if (n >= Lext) continue;
      const double v = src[jj * pitch + u];
      if (PASS == 0) {
        P.tmp[tbase + (long long)n * P.ncf + grp * cpw + jj] = v;
      } else if (n >= P.edge && n < P.edge + T) {            // reverse + trim
        P.y[(r0 + (long long)(Lext - 1 - P.edge - n)) * P.ncols + P.fcols[grp * cpw + jj]] = v;
      }
    }
  };
  for (int b = 0; b < kPrefetch && b < nblocks; ++b) load_block(b);
  __syncwarp();
  for (int t = 0; t < nblocks + nsec; ++t) {
    if (t - nsec >= 0) store_block(t - nsec);               // left the last section at step t - 1
    if (t + kPrefetch < nblocks) load_block(t + kPrefetch); // its slot held block t - nsec - 1, stored at step t - 1
    const int b = t - s;
    if (live && b >= 0 && b < nblocks) {
      double* buf = ring + ((size_t)(b % D) * cpw + ci) * pitch;
      const int cnt = min(kBlock, Lext - b * kBlock);
#pragma unroll 4
      for (int u = 0; u < cnt; ++u) {
        const double x_cur = buf[u];
        const double x_new = __dadd_rn(__dmul_rn(b0, x_cur), z0);
        z0 = __dadd_rn(__dsub_rn(__dmul_rn(b1, x_cur), __dmul_rn(a1, x_new)), z1);
        z1 = __dsub_rn(__dmul_rn(b2, x_cur), __dmul_rn(a2, x_new));
        buf[u] = x_new;
      }
    }
    __syncwarp();
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

// ---- integer decimation (up == 1 after reduction: 500 -> 250 / 125 / 100 Hz), the same arithmetic, register blocked ----
// In the general kernel every output reads its per_phase input rows from shared memory again (per_phase * ncols * 8 bytes
// per output: shared-memory bandwidth bound, 13.7 ms per 1,000 records).  Consecutive outputs of a decimator share all
// but `down` of their input rows, so here a thread produces kDecR consecutive outputs from ONE sliding pass over
// per_phase + (kDecR - 1) * down rows: each row is loaded once and feeds up to kDecR accumulators, each accumulator
// still sums its own taps oldest-sample-first with separately rounded multiply and add => still bit-identical to scipy.
// Threads start kDecR * down rows apart; one pad row per thread span makes that stride 32 bytes mod 128, and odd
// groups of four lanes fetch the two halves of a 32-byte row in the opposite order: conflict-free 16-byte loads.
constexpr int kDecR = 4;      // outputs per thread
constexpr int kDecNT = 128;   // threads per CTA: a tile is 512 outputs

// FUSED: multiply-add in one rounding (faster: half the fp64 instructions; within ~1e-15 of scipy instead of bit-identical)
template <int NC, bool FUSED>
__global__ void __launch_bounds__(kDecNT) resample_decim_kernel(const __grid_constant__ ResampleParams P, unsigned pb_magic) {
  extern __shared__ __align__(16) double s_dec[];
  double* s_taps = s_dec;
  double* s_x = s_dec + ((P.per_phase + 1) & ~1);            // padded rows of NC doubles
  const int down = P.down, pp = P.per_phase;
  const int PB = kDecR * down;                                // rows between the starts of neighbouring threads
  for (int i = threadIdx.x; i < pp; i += kDecNT) s_taps[i] = P.taps[i];
  const int rec = blockIdx.y;
  const long long i0 = P.in0[rec], len_x = P.in0[rec + 1] - i0;
  const long long o0 = P.out0[rec], n_out = P.out0[rec + 1] - o0;
  const int tid = threadIdx.x;
  const int swap = (tid >> 2) & 1;
  for (long long tile0 = (long long)blockIdx.x * (kDecNT * kDecR); tile0 < n_out; tile0 += (long long)gridDim.x * (kDecNT * kDecR)) {
    const long long first = (tile0 + P.n_pre_remove) * down - pp + 1;   // input row of staged row 0
    __syncthreads();
    // stage tile_rows input rows (zero outside the record = upfirdn's zero padding); staged row r sits at r + r / PB
    if constexpr (NC % 2 == 0) {
      constexpr int CH = NC / 2;                                       // 16-byte chunks per row
      const bool aligned = ((i0 * NC) & 1) == 0;                       // always for even NC
      // 16-byte asynchronous copies (LDGSTS) with zero fill outside the record: every copy of the tile is in flight
      // before the first one is waited for
      for (int q = tid; q < P.tile_rows * CH; q += kDecNT) {
        const int r = q / CH, h = q - r * CH;
        const long long xi = first + r;
        const bool in = xi >= 0 && xi < len_x && aligned;
        const double* src = P.x + (i0 + (in ? xi : 0)) * NC + 2 * h;
        const int pr = r + (int)__umulhi((unsigned)r, pb_magic);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(s_x + (size_t)pr * NC + 2 * h)), "l"(src),
                     "r"(in ? 16 : 0) : "memory");
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
      for (int q = tid; q < P.tile_rows * NC; q += kDecNT) {
        const int r = q / NC, c = q - r * NC;
        const long long xi = first + r;
        const int pr = r + (int)__umulhi((unsigned)r, pb_magic);
        s_x[(size_t)pr * NC + c] = (xi >= 0 && xi < len_x) ? P.x[(i0 + xi) * NC + c] : 0.0;
      }
    }
    __syncthreads();
    const long long mo = tile0 + (long long)tid * kDecR;
    if (mo < n_out) {
      double acc[kDecR][NC];
#pragma unroll
      for (int r = 0; r < kDecR; ++r)
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[r][c] = 0.0;
      // staged row j of this thread: tid * PB + j, at padded position tid * (PB + 1) + j + j / PB.  NC == 4: lanes whose
      // (tid >> 2) is odd fetch the two 16-byte halves of a row in the opposite order (conflict-free) and simply keep their
      // accumulators in that order; the columns are put back once, at the store.
      const double* base = s_x + (size_t)tid * (PB + 1) * NC;
      const int J = pp + (kDecR - 1) * down, ramp = (kDecR - 1) * down;
      auto load_row = [&](const double* p, double (&xv)[NC]) {
        if constexpr (NC == 4) {
          const double2 a = *reinterpret_cast<const double2*>(p + (swap ? 2 : 0));
          const double2 b = *reinterpret_cast<const double2*>(p + (swap ? 0 : 2));
          xv[0] = a.x; xv[1] = a.y; xv[2] = b.x; xv[3] = b.y;
        } else {
#pragma unroll
          for (int c = 0; c < NC; ++c) xv[c] = p[c];
        }
      };
      auto some = [&](int j, const double (&xv)[NC]) {        // ramp-up / ramp-down: output r takes rows r*down .. r*down + pp - 1
#pragma unroll
        for (int r = 0; r < kDecR; ++r) {
          const int k = j - r * down;
          if (k >= 0 && k < pp) {
            const double hk = s_taps[k];
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[r][c] = FUSED ? __fma_rn(xv[c], hk, acc[r][c]) : __dadd_rn(acc[r][c], __dmul_rn(xv[c], hk));
          }
        }
      };
      // Row j sits at pb + j * NC while it belongs to the current block of PB rows (one pad row per block).  Three
      // loops, so that the steady one is branch-free and the compiler can run the loads of the next rows ahead.
      const double* pb = base;
      int j = 0, bend = PB;
      for (; j < ramp; ++j) {                                 // ramp < PB: all in block 0
        double xv[NC];
        load_row(pb + (size_t)j * NC, xv);
        some(j, xv);
      }
      while (j < pp) {                                        // steady state: every output takes the row
        const int je = min(pp, bend);
#pragma unroll 4
        for (; j < je; ++j) {
          double xv[NC];
          load_row(pb + (size_t)j * NC, xv);
#pragma unroll
          for (int r = 0; r < kDecR; ++r) {
            const double hk = s_taps[j - r * down];
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[r][c] = FUSED ? __fma_rn(xv[c], hk, acc[r][c]) : __dadd_rn(acc[r][c], __dmul_rn(xv[c], hk));
          }
        }
        if (j == bend) { bend += PB; pb += NC; }
      }
      while (j < J) {
        const int je = min(J, bend);
        for (; j < je; ++j) {
          double xv[NC];
          load_row(pb + (size_t)j * NC, xv);
          some(j, xv);
        }
        if (j == bend) { bend += PB; pb += NC; }
      }
      if constexpr (NC == 4) {
#pragma unroll
        for (int r = 0; r < kDecR; ++r) {
          const double t0 = acc[r][0], t1 = acc[r][1];
          acc[r][0] = swap ? acc[r][2] : t0; acc[r][1] = swap ? acc[r][3] : t1;
          acc[r][2] = swap ? t0 : acc[r][2]; acc[r][3] = swap ? t1 : acc[r][3];
        }
      }
#pragma unroll
      for (int r = 0; r < kDecR; ++r) {
        if (mo + r < n_out) {
          double* o = P.y + (o0 + mo + r) * NC;
          if constexpr (NC % 2 == 0) {
#pragma unroll
            for (int c = 0; c < NC; c += 2) __stcs(reinterpret_cast<double2*>(o + c), make_double2(acc[r][c], acc[r][c + 1]));
          } else {
#pragma unroll
            for (int c = 0; c < NC; ++c) __stcs(o + c, acc[r][c]);
          }
        }
      }
    }
  }
}

}  // namespace scgrhc

