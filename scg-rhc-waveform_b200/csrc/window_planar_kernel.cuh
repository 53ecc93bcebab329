// The hot-path kernel for PLANAR arenas (SCGRHC_ARENA_PLANAR): column c of arena row r lives at arena[c * arena_rows + r].
//
// Same arithmetic and the same bit-exact contract as window_kernel.cuh (which reads wfdb's interleaved (rows, nsig)
// layout); what changes is the data movement the layout allows:
//   * a candidate window is first judged on its RHC plane alone — ONE contiguous 6,000-byte bulk copy
//     (has_noise() looks at the RHC channel only, waveform_noise.py:44-49) — and the C SCG planes are fetched only if it is
//     kept: a rejected window costs 6 KB of DRAM traffic instead of the 24 KB the interleaved rows drag in by sector;
//   * every thread owns PAIRS of consecutive samples: 16-byte shared loads without bank conflicts, 8-byte global stores
//     (two fp32 outputs), no column selects and no transposition — the output (C, W) layout IS planar.
//
// Software pipeline per CTA (one candidate per iteration j, one __syncthreads per iteration):
//   phase A, item a = lo + j      RHC window (bulk copy issued two iterations earlier) -> registers -> predicates ->
//                                 keep / reject -> RHC normalised and stored; if kept, the SCG bulk copies are issued
//   phase B, item b = lo + j - 2  its SCG planes have landed -> registers -> joint min/max -> normalise -> store
// Both phases put their per-warp partials into one shared exchange before the barrier of the iteration.
#pragma once
#include "window_kernel.cuh"

namespace scgrhc {

#ifndef SCGRHC_PNR
#define SCGRHC_PNR 3
#endif
constexpr int PNR = SCGRHC_PNR;      // RHC windows in flight per CTA (6 KB each); 3 keeps 4 CTAs/SM at C = 3 (4 slots: 3 CTAs, slower)
constexpr int PNRED = 12;   // phase A: -ymin, ymax | s1, s2, sxy, dense ; phase B: -smin, smax | nan accumulator

struct PMeta {      // what every thread reads per item: 16 bytes
  long long cand;   // candidate index (the output slot too, unless the job walks a kept list)
  long long rowf;   // first arena row of the window; bit 62: bulk copy not possible (capacity edge) -> cooperative plain loads
};
constexpr long long kPFallback = 1LL << 62;

template <int NTH>
struct PScratch {
  uint64_t rfull[PNR], sfull[2];
  PMeta rmeta[PNR];                       // item whose RHC window sits in RHC slot s
  PMeta bmeta[2];                       // kept item whose SCG planes sit (or are landing) in SCG slot s
  int2 raux[PNR];                       // (window number inside its interval, record id) of the item in RHC slot s: thread 0 only
  double red[2][NTH / 32][PNRED];
  uint32_t cmask[2][8 * (NTH / 32)];
  uint32_t a24[8 * (NTH / 32)];
  int slow_cnt;
  // the RHC producer's interval cursor (thread 0 only): shared memory instead of seven registers in every thread
  long long p_cand0, p_row0;
  int p_nwin, p_rec, p_iv;
};

// Window of W samples; NTH threads; every thread owns PR pairs (samples 2p, 2p+1 for p = tid + k*NTH): W <= 2*PR*NTH.
// WCT > 0: compile-time window length (750 = int(1.5 * 500), all 37 configs): pair validity folds away; 0: runtime length.
// PLAIN: no mode flag besides the layout (see window_kernel): the mode tests fold away.
template <int C, int NTH, int PR, typename OutT, int WCT, bool PLAIN = false>
__global__ void __launch_bounds__(NTH, NTH == 64 ? 8 : 4) window_planar_kernel(const __grid_constant__ KParams P) {
  static_assert(WCT == 0 || (WCT + 1) / 2 <= PR * NTH, "window does not fit the pair grid");
  static_assert(SCGRHC_FLAT_WIN == 50, "run detection below is hard-wired to 24 whole pairs inside a flat run of 50 samples");
  constexpr int NW = NTH / 32;
  constexpr int NWORDS = PR * NW;
  static_assert(NWORDS <= 32, "one mask word per lane");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  PScratch<NTH>& S = *reinterpret_cast<PScratch<NTH>*>(smem_raw);
  double* const rbase = reinterpret_cast<double*>(smem_raw + ((sizeof(PScratch<NTH>) + 127) & ~size_t(127)));
  const int wpad = P.stage_elems;                      // doubles per plane window in shared memory (even, >= W + 3)
  double* const sbase = rbase + PNR * wpad;

  const scgrhc_job& J = P.job;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = WCT > 0 ? WCT : J.W;
  const long long rows = J.arena_rows;
  const int wstride = J.stride > 0 ? J.stride : W;
  const bool use_list = !PLAIN && (J.flags & SCGRHC_USE_KEPT_LIST) != 0;
  const bool pred_only = !PLAIN && (J.flags & SCGRHC_PREDICATES_ONLY) != 0;
  const bool norm_global = !PLAIN && (J.flags & SCGRHC_NORM_GLOBAL) != 0;
  const bool keep_all = !PLAIN && (J.flags & SCGRHC_KEEP_ALL) != 0;
  const double thr = J.flat_threshold, min_rhc = J.min_rhc;
  // plane bases are recomputed where they are used (the producers and the capacity-edge loads): job fields are constant-bank
  // operands, pointers held across the loop would be eight more live registers
  auto yplane_of = [&]() { return J.arena + (long long)J.rhc_col * rows; };
  auto xplane_of = [&](int c) { return J.arena + (long long)J.scg_cols[c] * rows; };

  const long long items = use_list ? J.n_items : J.n_cand;
  const long long lo = items * (long long)blockIdx.x / gridDim.x;
  const long long hi = items * (long long)(blockIdx.x + 1) / gridDim.x;
  if (lo >= hi) return;
  const int cnt = (int)(hi - lo);               // a CTA's share of the items: far below 2^31

  if (tid == 0) {
    for (int i = 0; i < PNR; ++i) mbar_init(&S.rfull[i], 1);
    mbar_init(&S.sfull[0], 1); mbar_init(&S.sfull[1], 1);
    S.slow_cnt = 0;
    fence_barrier_init();
  }
  __syncthreads();

  // parity of the plane base c * rows (lead = (parity ^ row) & 1): one AND of two job fields where it is needed, not a register each
  auto ppar_of = [&](int c) { return (int)(J.scg_cols[c] & (int)rows & 1); };
  auto ypar_of = [&]() { return (int)(J.rhc_col & (int)rows & 1); };

  // ---- producer state (thread 0 only) -----------------------------------------------------------
  auto load_iv = [&](int iv) {
    const scgrhc_interval I = J.intervals[iv];
    S.p_cand0 = I.cand0; S.p_row0 = I.row0; S.p_nwin = I.n_win; S.p_rec = I.rec_id; S.p_iv = iv;
  };
  // one plane window -> shared memory: 16-byte aligned start (the element before when the offset is odd), even length.
  // A window within two rows of the end of the planes takes cooperative plain loads instead (the copy of the LAST plane
  // would read past the arena; for the other planes the one-element overshoot lands in the next plane, harmlessly).
  const uint32_t bytes_even = (uint32_t)(((W + 1) & ~1) * 8), bytes_odd = (uint32_t)(((W + 2) & ~1) * 8);
  const bool tail_ok = (long long)J.arena_capacity_bytes >= (long long)J.nsig * rows * 8 + 16;   // room behind the last plane
  auto issue_rhc = [&](long long item, int s) {
    const long long cand = use_list ? J.kept_list[item] : item;
    while (cand >= S.p_cand0 + S.p_nwin) load_iv(S.p_iv + 1);
    const int i = (int)(cand - S.p_cand0);
    const long long row = S.p_row0 + (long long)i * wstride;
    const int p_rec = S.p_rec;
    const bool fb = !tail_ok && row + W + 2 > rows;
    PMeta m;
    m.cand = cand; m.rowf = row | (fb ? kPFallback : 0);
    S.rmeta[s] = m;
    S.raux[s] = make_int2(i, p_rec);
    if (fb) {
      mbar_arrive(&S.rfull[s]);
    } else {
      const int lead = (ypar_of() ^ (int)row) & 1;
      const uint32_t bytes = lead ? bytes_odd : bytes_even;
      mbar_arrive_expect_tx(&S.rfull[s], bytes);
      bulk_g2s(rbase + (size_t)s * wpad, yplane_of() + (row - lead), bytes, &S.rfull[s]);
    }
  };
  auto issue_scg = [&](const PMeta& m, int s) {      // the SCG planes of a kept window; plain loads in phase B if the item is at the edge
    S.bmeta[s] = m;
    if (m.rowf & kPFallback) return;
    const long long row = m.rowf;
    int odd = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) odd += (ppar_of(c) ^ (int)row) & 1;
    mbar_arrive_expect_tx(&S.sfull[s], (uint32_t)C * bytes_even + (uint32_t)odd * (bytes_odd - bytes_even));
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int lead = (ppar_of(c) ^ (int)row) & 1;
      bulk_g2s(sbase + ((size_t)s * C + c) * wpad, xplane_of(c) + (row - lead), lead ? bytes_odd : bytes_even, &S.sfull[s]);
    }
  };
  if (tid == 0) {
    const long long first = use_list ? J.kept_list[lo] : lo;
    int a = 0, b = J.n_intervals - 1;  // last interval with cand0 <= first
    while (a < b) {
      const int mid = (a + b + 1) >> 1;
      if (J.intervals[mid].cand0 <= first) a = mid; else b = mid - 1;
    }
    load_iv(a);
    for (int i = 0; i < PNR && i < cnt; ++i) issue_rhc(lo + i, i);
  }

  const double xbar = 0.5 * (double)(W - 1);
  const double sxx = (double)W * ((double)W * (double)W - 1.0) / 12.0;  // sum (t - xbar)^2, exact here
  const double tx = (double)(2 * tid) - xbar;
  const double inv_w = 1.0 / (double)W;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  const int npairs = (W + 1) >> 1;
  // pair p = tid + k*NTH exists / has its second sample: compile-time for all but the last k when the length is known
  auto pair_ok = [&](int k, int p) { return WCT > 0 ? ((k + 1) * NTH <= (WCT + 1) / 2 || p < (WCT + 1) / 2) : (p < npairs); };
  auto has_second = [&](int k, int p) { return WCT > 0 ? (WCT % 2 == 0 || (k + 1) * NTH <= WCT / 2 || 2 * p + 1 < WCT) : (2 * p + 1 < W); };

  uint32_t spar[2] = {0u, 0u};
  uint32_t khist = 0u;        // bit k: the item of iteration j - 1 - k was kept (its SCG planes were requested)
  int rs = 0;                 // RHC slot of item j: j % PNR, parity (j / PNR) & 1
  uint32_t rparity = 0u;
  for (int j = 0; j < cnt + 2; ++j) {
    const int s = (int)(j & 1);
    const bool doA = j < cnt;
    const bool doB = ((khist >> 1) & 1u) != 0;         // the (CTA-uniform) decision of iteration j - 2

    // ================= phase A, before the barrier: RHC window -> registers, per-thread statistics ==================
    PMeta MA;
    double y0[PR], y1[PR];
    double K = 0.0;
    if (doA) {
      mbar_wait(&S.rfull[rs], rparity);
      MA = S.rmeta[rs];
      double* buf = rbase + (size_t)rs * wpad;
      const long long rowA = MA.rowf & ~kPFallback;
      const int lead = (ypar_of() ^ (int)rowA) & 1;
      if (MA.rowf & kPFallback) {
        for (int e = tid; e < W; e += NTH) buf[lead + e] = yplane_of()[rowA + e];
        __syncthreads();
      }
      const double* win = buf + lead;
      K = win[0];
      double a_ymin = CUDART_INF, a_ymax = -CUDART_INF, s1 = 0.0, s2 = 0.0, sB = 0.0, sC = 0.0;
      bool dense_word = false;
#pragma unroll
      for (int k = 0; k < PR; ++k) {
        const int p = tid + k * NTH;
        const int i0 = 2 * p;
        y0[k] = 0.0; y1[k] = 0.0;
        bool full = false;
        if (pair_ok(k, p)) {
          const bool has1 = has_second(k, p);
          double v0, v1;
          if (lead == 0) {               // CTA-uniform: 16-byte loads when the window starts on an even element of the buffer
            const double2 q = *reinterpret_cast<const double2*>(win + i0);
            v0 = q.x; v1 = q.y;
          } else {
            v0 = win[i0]; v1 = win[i0 + 1];
          }
          y0[k] = v0; y1[k] = v1;
          a_ymin = v0 < a_ymin ? v0 : a_ymin;
          a_ymax = v0 > a_ymax ? v0 : a_ymax;
          const double d0 = __dsub_rn(v0, K);
          s2 = __fma_rn(d0, d0, s2);
          double dd = d0;
          if (has1) {
            a_ymin = v1 < a_ymin ? v1 : a_ymin;
            a_ymax = v1 > a_ymax ? v1 : a_ymax;
            const double d1 = __dsub_rn(v1, K);
            s2 = __fma_rn(d1, d1, s2);
            sC = __dadd_rn(sC, d1);
            dd = __dadd_rn(d0, d1);
            // fl(|y[2p+1] - y[2p]|) < thr is necessary for any flat 50-window that holds the pair (rounding is monotone and
            // the step cannot exceed the window's range); the steps BETWEEN pairs are left to the exact recheck
            full = fabs(__dsub_rn(v1, v0)) < thr;
          }
          s1 = __dadd_rn(s1, dd);         // the pair's sum serves both the plain and the k-weighted sum
          sB = __fma_rn((double)k, dd, sB);
        }
        const uint32_t word = __ballot_sync(kFull, full);
        if (lane == 0) S.cmask[s][k * NW + warp] = word;
        // a flat run of 50 samples holds 24 consecutive whole pairs, which span at most 2 mask words
        dense_word |= __popc(word) >= 12;
      }
      // sum_i (i - xbar) dy_i over this thread's samples i = 2(tid + NTH k) + b
      double sxy = __fma_rn(tx, s1, __fma_rn((double)(2 * NTH), sB, sC));
      a_ymin = warp_min(a_ymin); a_ymax = warp_max(a_ymax);
      const double s3 = warp_sum3(s1, s2, sxy, lane);        // lane 0: s1, lane 8: s2, lane 16: sxy
      double* r = S.red[s][warp];
      if ((lane & 7) == 0 && lane < 24) r[4 + (lane >> 3)] = s3;
      if (lane == 0) { r[0] = -a_ymin; r[1] = a_ymax; r[7] = dense_word ? 1.0 : 0.0; }
    }

    // ================= phase B, before the barrier: SCG planes -> registers, joint min/max ==========================
    PMeta MB;
    double x0[PR][C], x1[PR][C];
    if (doB) {
      MB = S.bmeta[s];
      double* buf = sbase + (size_t)s * C * wpad;
      const long long rowB = MB.rowf & ~kPFallback;
      if (!(MB.rowf & kPFallback)) {
        mbar_wait(&S.sfull[s], spar[s]);
        spar[s] ^= 1u;
      } else {                                   // capacity edge: plain loads
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const int lead = (ppar_of(c) ^ (int)rowB) & 1;
          for (int e = tid; e < W; e += NTH) buf[(size_t)c * wpad + lead + e] = xplane_of(c)[rowB + e];
        }
        __syncthreads();
      }
      double a_smin = CUDART_INF, a_smax = -CUDART_INF, nanacc = 0.0;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int lead = (ppar_of(c) ^ (int)rowB) & 1;
        const double* win = buf + (size_t)c * wpad + lead;
#pragma unroll
        for (int k = 0; k < PR; ++k) {
          const int p = tid + k * NTH;
          const int i0 = 2 * p;
          x0[k][c] = 0.0; x1[k][c] = 0.0;
          if (pair_ok(k, p)) {
            double v0, v1;
            if (lead == 0) {
              const double2 q = *reinterpret_cast<const double2*>(win + i0);
              v0 = q.x; v1 = q.y;
            } else {
              v0 = win[i0]; v1 = win[i0 + 1];
            }
            x0[k][c] = v0; x1[k][c] = v1;
            if (has_second(k, p)) {
              // joint min / max of a pair with three comparisons instead of four; a NaN can hide its partner from one
              // of the two extrema, but a window with a NaN gets NaN extrema below whatever these say
              const bool lt = v0 < v1;
              const double lo2 = lt ? v0 : v1, hi2 = lt ? v1 : v0;
              a_smin = lo2 < a_smin ? lo2 : a_smin;
              a_smax = hi2 > a_smax ? hi2 : a_smax;
              // one FMA per pair: the accumulator ends non-finite if v0 or v1 is NaN or Inf (Inf * 0 = NaN, Inf * v = Inf,
              // Inf - Inf = NaN); a finite overflow of the products only sends the window through the exact recheck
              nanacc = __fma_rn(v0, v1, nanacc);
            } else {
              a_smin = v0 < a_smin ? v0 : a_smin;
              a_smax = v0 > a_smax ? v0 : a_smax;
              nanacc = __fma_rn(v0, 0.0, nanacc);
            }
          }
        }
      }
      a_smin = warp_min(a_smin); a_smax = warp_max(a_smax);
      const bool odd_b = __any_sync(kFull, !(fabs(nanacc) <= DBL_MAX));      // a vote instead of a sum tree
      if (lane == 0) {
        double* r = S.red[s][warp];
        r[2] = -a_smin; r[3] = a_smax; r[8] = odd_b ? qnan : 0.0;
      }
    }
    __syncthreads();  // the barrier of the iteration: partials visible, RHC slot s and SCG slot s are in registers

    // ---- block combine, lane parallel: lane v folds slot v over the warps, results broadcast ---------------------------
    double ymin, ymax, smin, smax, s1, s2, sxy, dense, nanB;
    {
      const int v = lane < PNRED ? lane : 0;
      double a = S.red[s][0][v];
#pragma unroll
      for (int w = 1; w < NW; ++w) {
        const double b = S.red[s][w][v];
        const double mx = b > a ? b : a;
        const double sm = __dadd_rn(a, b);
        a = v < 4 ? mx : sm;
      }
      ymin = -__shfl_sync(kFull, a, 0); ymax = __shfl_sync(kFull, a, 1);
      smin = -__shfl_sync(kFull, a, 2); smax = __shfl_sync(kFull, a, 3);
      s1 = __shfl_sync(kFull, a, 4); s2 = __shfl_sync(kFull, a, 5); sxy = __shfl_sync(kFull, a, 6);
      dense = __shfl_sync(kFull, a, 7); nanB = __shfl_sync(kFull, a, 8);
    }

    // ================= phase A, after the barrier: decide, publish, normalise the RHC window ============================
    if (doA) {
      uint32_t reason = 0;
      bool keep = true;
      if (!use_list) {
        uint32_t a24 = 0u;
        bool run = false;
        if (dense != 0.0) {
          const uint32_t a1 = lane < NWORDS ? S.cmask[s][lane] : 0u;
          auto down = [&](uint32_t v, int d) {  // word (lane + d) of the mask, 0 past the end
            const uint32_t o = __shfl_down_sync(kFull, v, d);
            return (lane + d < 32) ? o : 0u;
          };
          auto shr = [&](uint32_t v, int sft) { return __funnelshift_r(v, down(v, 1), sft); };
          const uint32_t a2 = a1 & shr(a1, 1);
          const uint32_t a4 = a2 & shr(a2, 2);
          const uint32_t a8 = a4 & shr(a4, 4);
          const uint32_t a16 = a8 & shr(a8, 8);
          a24 = a16 & shr(a8, 16);          // bit p: pairs p .. p+23 are all full
          run = __any_sync(kFull, a24 != 0u);
        }
        const bool s2_bad = !(s2 <= DBL_MAX);   // non-finite RHC sample, or the sum of squares overflowed: recheck exactly
        int flat_cnt = 0;
        bool nonfinite = false;
        if (run || s2_bad) {                    // CTA-uniform and rare: exact work on the RHC window still in shared memory
          if (warp == 0 && lane < NWORDS) S.a24[lane] = a24;
          __syncthreads();
          const double* win = rbase + (size_t)rs * wpad + ((ypar_of() ^ (int)MA.rowf) & 1);
          int c = 0, fl = 0;
          if (run) {
            for (int k = 0; k < PR; ++k) {
              const int p = tid + k * NTH;
              if (pair_ok(k, p) && ((S.a24[p >> 5] >> (p & 31)) & 1u)) {
                for (int q = 2 * p - 1; q <= 2 * p; ++q) {      // a flat 50-window can start at the pair or one sample before it
                  if (q < 0 || q + SCGRHC_FLAT_WIN > W) continue;
                  double mx = -CUDART_INF, mn = CUDART_INF;
                  for (int i = 0; i < SCGRHC_FLAT_WIN; ++i) {
                    const double v = win[q + i];
                    mx = v > mx ? v : mx; mn = v < mn ? v : mn;
                  }
                  c += (__dsub_rn(mx, mn) < thr) ? 1 : 0;
                }
              }
            }
          }
#pragma unroll
          for (int k = 0; k < PR; ++k) {
            const int p = tid + k * NTH;
            if (pair_ok(k, p)) {
              if (!(fabs(y0[k]) <= DBL_MAX)) fl = 1;
              if (has_second(k, p) && !(fabs(y1[k]) <= DBL_MAX)) fl = 1;
            }
          }
          if (c) atomicAdd(&S.slow_cnt, c);
          nonfinite = __syncthreads_or(fl) != 0;
          flat_cnt = S.slow_cnt;
          __syncthreads();
          if (tid == 0) S.slow_cnt = 0;
        }
        // Syy = sum (y-K)^2 - (sum (y-K))^2 / n ; Sxy is shift invariant because sum (t - xbar) = 0.
        // R^2 > 0.8  <=>  Sxy^2 > 0.8 Sxx Syy (Syy > 0): no division; see window_kernel.cuh for the ambiguity band.
        const double syy = __dsub_rn(s2, __dmul_rn(__dmul_rn(s1, s1), inv_w));
        const double lhs = __dmul_rn(sxy, sxy), den = __dmul_rn(sxx, syy);
        if (flat_cnt >= 2) reason |= SCGRHC_REASON_FLAT;
        if (syy > 0.0 && lhs > __dmul_rn(0.8, den)) reason |= SCGRHC_REASON_STRAIGHT;
        if (syy > 0.0 && fabs(__fma_rn(-0.8, den, lhs)) < __dmul_rn(1e-9, den)) reason |= SCGRHC_REASON_AMBIGUOUS;
        if (ymin == ymax && !nonfinite && np_mean_of_const_is_exact(ymin, W)) reason |= SCGRHC_REASON_STRAIGHT;
        if (ymin < min_rhc) reason |= SCGRHC_REASON_FLOOR;
        if (nonfinite) reason |= SCGRHC_REASON_NONFINITE;
        keep = keep_all ||
               (reason & (SCGRHC_REASON_FLAT | SCGRHC_REASON_STRAIGHT | SCGRHC_REASON_FLOOR | SCGRHC_REASON_NONFINITE)) == 0;
        if (tid == 0) {
          P.out.keep[MA.cand] = keep ? 1 : 0;
          P.out.reason[MA.cand] = (uint8_t)reason;
          double2* mm = reinterpret_cast<double2*>(P.out.minmax + 4 * MA.cand);
          if (!keep) mm[0] = make_double2(qnan, qnan);   // the SCG planes of a rejected window are never read
          mm[1] = make_double2(ymin, ymax);
          const int2 aux = S.raux[rs];
          P.out.cand_win[MA.cand] = aux.x;
          P.out.cand_rec[MA.cand] = aux.y;
          if (!keep_all && (reason & SCGRHC_REASON_NONFINITE) && !(reason & SCGRHC_REASON_FLAT)) {
            atomicOr(P.err, 1ull);
            atomicMin(P.err + 1, (unsigned long long)MA.cand);
          }
          if (reason & SCGRHC_REASON_AMBIGUOUS) atomicAdd(P.err + 2, 1ull);
        }
      } else if (!norm_global) {   // dense re-materialisation with the per-window pairs of an earlier pass
        const double2 b = reinterpret_cast<const double2*>(P.out.minmax + 4 * MA.cand)[1];
        ymin = b.x; ymax = b.y;
      }
      // two producers, in different warps, so that neither serial section sits alone on the iteration's critical path:
      // thread 0 refills the RHC slot (it owns the interval cursor), thread 32 issues the SCG planes of the kept window
      if (tid == 0 && j + PNR < cnt) issue_rhc(lo + j + PNR, rs);    // RHC slot rs is in registers everywhere
      if (tid == 32 && keep) issue_scg(MA, s);
      khist = (khist << 1) | (keep ? 1u : 0u);
      if (keep && !pred_only) {
        if (norm_global) { ymin = J.global_minmax[2]; ymax = J.global_minmax[3]; }
        Normaliser nr;
        nr.init(ymin, ymax);
        const long long slotA = use_list ? lo + j : MA.cand;
        OutT* ro = reinterpret_cast<OutT*>(P.out.rhc_out) + (size_t)slotA * W;
        const bool vec = (WCT > 0 && WCT % 2 == 0) || (((size_t)slotA * W) & 1) == 0;     // 8-byte aligned pair stores
        bool redo = true;
        if constexpr (sizeof(OutT) == 4) {
          if (nr.quick) {         // tier 1, see window_kernel.cuh (whatever the pairs' origin: this window, an earlier pass, the dataset)
            uint32_t acc = 0xffffffffu;
#pragma unroll
            for (int k = 0; k < PR; ++k) {
              const int p = tid + k * NTH;
              if (pair_ok(k, p)) {
                const double q0 = __dmul_rn(__dsub_rn(y0[k], nr.mn), nr.inv), q1 = __dmul_rn(__dsub_rn(y1[k], nr.mn), nr.inv);
                acc = min(acc, tier1_key(q0));
                if (has_second(k, p)) {
                  acc = min(acc, tier1_key(q1));
                  if (vec) __stcs(reinterpret_cast<float2*>(ro + 2 * p), make_float2(__double2float_rn(q0), __double2float_rn(q1)));
                  else { st_cs(ro + 2 * p, __double2float_rn(q0)); st_cs(ro + 2 * p + 1, __double2float_rn(q1)); }
                } else {
                  st_cs(ro + 2 * p, __double2float_rn(q0));
                }
              }
            }
            redo = acc <= kTier1Risky;
          }
        }
        if (redo) {
#pragma unroll 1
          for (int k = 0; k < PR; ++k) {
            const int p = tid + k * NTH;
            if (pair_ok(k, p)) {
              double v0 = 0.0, v1 = 0.0;
#pragma unroll
              for (int kk = 0; kk < PR; ++kk) if (kk == k) { v0 = y0[kk]; v1 = y1[kk]; }
              OutT o;
              cvt_out(o, nr.slow ? nr.exact(v0) : nr.fast(v0)); st_cs(ro + 2 * p, o);
              if (has_second(k, p)) { cvt_out(o, nr.slow ? nr.exact(v1) : nr.fast(v1)); st_cs(ro + 2 * p + 1, o); }
            }
          }
        }
      }
    } else {
      khist <<= 1;                                           // drain iterations: nothing enters the pipeline any more
    }

    // ================= phase B, after the barrier: normalise the SCG block from registers, store =======================
    if (doB) {
      if (!use_list) {
        if (nanB != nanB) {     // some SCG sample is NaN or Inf (rare, CTA-uniform): np.min / np.max give NaN iff one is NaN
          int fl = 0;
#pragma unroll
          for (int k = 0; k < PR; ++k) {
            const int p = tid + k * NTH;
            if (pair_ok(k, p)) {
#pragma unroll
              for (int c = 0; c < C; ++c) {
                if (x0[k][c] != x0[k][c]) fl = 1;
                if (has_second(k, p) && x1[k][c] != x1[k][c]) fl = 1;
              }
            }
          }
          if (__syncthreads_or(fl)) { smin = qnan; smax = qnan; }
        }
        if (tid == 0) reinterpret_cast<double2*>(P.out.minmax + 4 * MB.cand)[0] = make_double2(smin, smax);
      } else if (!norm_global) {
        const double2 a = reinterpret_cast<const double2*>(P.out.minmax + 4 * MB.cand)[0];
        smin = a.x; smax = a.y;
      }
      if (!pred_only) {
        if (norm_global) { smin = J.global_minmax[0]; smax = J.global_minmax[1]; }
        Normaliser ns;
        ns.init(smin, smax);
        const long long slotB = use_list ? lo + j - 2 : MB.cand;
        OutT* so = reinterpret_cast<OutT*>(P.out.scg_out) + (size_t)slotB * C * W;
        bool redo = true;
        if constexpr (sizeof(OutT) == 4) {
          if (ns.quick) {
            uint32_t acc = 0xffffffffu;
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const bool vec = (WCT > 0 && WCT % 2 == 0) || ((((size_t)slotB * C + c) * W) & 1) == 0;
#pragma unroll
              for (int k = 0; k < PR; ++k) {
                const int p = tid + k * NTH;
                if (pair_ok(k, p)) {
                  const double q0 = __dmul_rn(__dsub_rn(x0[k][c], ns.mn), ns.inv), q1 = __dmul_rn(__dsub_rn(x1[k][c], ns.mn), ns.inv);
                  acc = min(acc, tier1_key(q0));
                  float* o = so + (size_t)c * W + 2 * p;
                  if (has_second(k, p)) {
                    acc = min(acc, tier1_key(q1));
                    if (vec) __stcs(reinterpret_cast<float2*>(o), make_float2(__double2float_rn(q0), __double2float_rn(q1)));
                    else { st_cs(o, __double2float_rn(q0)); st_cs(o + 1, __double2float_rn(q1)); }
                  } else {
                    st_cs(o, __double2float_rn(q0));
                  }
                }
              }
            }
            redo = acc <= kTier1Risky;
          }
        }
        if (redo) {   // tiers 2 / 3: exact quotients; out of line (not unrolled) so the common path stays small
#pragma unroll 1
          for (int e = 0; e < PR * C; ++e) {
            const int k = e / C, c = e - k * C;
            const int p = tid + k * NTH;
            if (pair_ok(k, p)) {
              double v0 = 0.0, v1 = 0.0;
#pragma unroll
              for (int kk = 0; kk < PR; ++kk) {
#pragma unroll
                for (int cc = 0; cc < C; ++cc) if (kk == k && cc == c) { v0 = x0[kk][cc]; v1 = x1[kk][cc]; }
              }
              OutT o;
              OutT* dst = so + (size_t)c * W + 2 * p;
              cvt_out(o, ns.slow ? ns.exact(v0) : ns.fast(v0)); st_cs(dst, o);
              if (has_second(k, p)) { cvt_out(o, ns.slow ? ns.exact(v1) : ns.fast(v1)); st_cs(dst + 1, o); }
            }
          }
        }
      }
    }
    if (++rs == PNR) { rs = 0; rparity ^= 1u; }
  }
}

}  // namespace scgrhc
