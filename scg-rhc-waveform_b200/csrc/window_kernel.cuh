// The fused hot-path kernel: one candidate window per CTA iteration.
//
// Restates, per candidate window (all arithmetic that decides keep/reject is fp64, bit-compatible
// with the reference's numpy/pandas path):
//   get_flat_lines   waveform_noise.py:6-26   (rolling 50-sample range < 1e-3 at >= 2 positions)
//   is_straight_line waveform_noise.py:29-34  (OLS R^2 > 0.8, closed form Sxy^2/(Sxx*Syy))
//   in_rhc_range     waveform_noise.py:37-41  (any sample < min_RHC)
//   init_segments    recordutil.py:55-66      (joint SCG min/max, RHC min/max, (x-mn)/(mx-mn+1e-4),
//                                              transpose to (C,W), cast fp32)
//
// Data movement: a window is W consecutive rows of the (rows, nsig) fp64 arena, i.e. ONE contiguous
// chunk of W*nsig*8 bytes.  An elected thread streams it into shared memory with a 1-D bulk async
// copy (cp.async.bulk -> SASS UBLKCP) completing on an mbarrier; `stages` windows are in flight per
// CTA, several CTAs per SM.  Each thread then owns rows t = tid + k*NT in registers, so the window is
// read from shared memory once; reductions are warp shuffles + one small shared exchange.
#pragma once
#include <float.h>
#include <math_constants.h>

#include "common.cuh"

namespace scgrhc {

constexpr int NT = 128;          // threads per CTA
constexpr int NWARP = NT / kWarp;
constexpr int kMaxStages = 8;

struct StageMeta {
  long long cand;   // candidate index
  long long slot;   // output slot
  long long elem0;  // arena element offset of the window's first sample (row * nsig)
  int win;          // window number inside its interval
  int rec;          // record id
  int lead;         // 0/1 padding doubles in front of the window inside the stage buffer
  int fallback;     // 1: bulk copy not possible (capacity edge) -> cooperative plain loads
};

struct KParams {
  scgrhc_job job;
  scgrhc_outputs out;
  unsigned long long* err;  // [0] error flags, [1] first offending candidate (atomicMin)
  int stages;
  int stage_elems;          // doubles per stage buffer (even)
  long long arena_elems_cap;
};

template <int R>
struct Scratch {
  uint64_t full[kMaxStages];
  StageMeta meta[kMaxStages];
  double red1[NWARP][5];
  double red2[NWARP][2];
  uint32_t a49[R * NWARP];
  uint32_t cmask[R * NWARP];
  uint32_t wflags[NWARP];
  int need_slow;
  int slow_cnt[2];
};

// Correctly rounded a/d from a correctly rounded reciprocal (Markstein): two FMA residual
// corrections.  Validated against __ddiv_rn in tests (scgrhc_selftest_div).
__device__ __forceinline__ double div_by_recip(double a, double d, double inv) {
  double q = __dmul_rn(a, inv);
  double r = __fma_rn(-d, q, a);
  q = __fma_rn(r, inv, q);
  r = __fma_rn(-d, q, a);
  q = __fma_rn(r, inv, q);
  if (fabs(q) < 0x1p-900 && a != 0.0) q = __ddiv_rn(a, d);  // residuals could underflow
  return q;
}

struct Normaliser {
  double mn, d, inv;
  bool slow;
  __device__ __forceinline__ void init(double mn_, double mx_) {
    mn = mn_;
    d = __dadd_rn(__dsub_rn(mx_, mn_), 0.0001);  // (max - min + 0.0001), recordutil.py:46
    inv = __drcp_rn(d);
    slow = !(d < 0x1p1000 && d > 0x1p-1000);     // also catches NaN
  }
  __device__ __forceinline__ double operator()(double x) const {
    double a = __dsub_rn(x, mn);
    return slow ? __ddiv_rn(a, d) : div_by_recip(a, d, inv);
  }
};

__device__ __forceinline__ void cvt_out(float& o, double q) { o = __double2float_rn(q); }
__device__ __forceinline__ void cvt_out(double& o, double q) { o = q; }

__device__ __forceinline__ double sel4(int col, double v0, double v1, double v2, double v3) {
  double lo = (col & 1) ? v1 : v0;
  double hi = (col & 1) ? v3 : v2;
  return (col & 2) ? hi : lo;
}

// ---- diagnostics: div_by_recip vs IEEE division on hashed operands --------------------------------
__device__ __forceinline__ unsigned long long st_mix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  unsigned long long z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256) selftest_div_kernel(unsigned long long seed, long long n, int mode,
                                                           unsigned long long* counts) {
  unsigned long long bad64 = 0, bad32 = 0, cnt = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long h1 = st_mix64(seed ^ (unsigned long long)(2 * i));
    const unsigned long long h2 = st_mix64(seed ^ (unsigned long long)(2 * i + 1));
    const double m1 = 1.0 + (double)(h1 >> 12) * 0x1p-52, m2 = (double)(h2 >> 11) * 0x1p-53;
    double d, a;
    if (mode == 0) {
      d = ldexp(m1, (int)(h1 & 0xFFF) % 24 - 14);  // 2^-14 .. 2^10
      a = d * m2;                                  // 0 <= a < d, rounded product: arbitrary mantissa
    } else {
      d = ldexp(m1, (int)(h1 & 0xFFF) % 801 - 400);
      a = ldexp(1.0 + m2, (int)(h2 & 0xFFF) % 801 - 400);
      if (h2 & 0x1000) a = -a;
    }
    Normaliser nz;
    nz.mn = 0.0; nz.d = d; nz.inv = __drcp_rn(d); nz.slow = !(d < 0x1p1000 && d > 0x1p-1000);
    const double q = nz.slow ? __ddiv_rn(a, d) : div_by_recip(a, d, nz.inv);
    const double ref = __ddiv_rn(a, d);
    bad64 += (__double_as_longlong(q) != __double_as_longlong(ref));
    bad32 += (__float_as_int(__double2float_rn(q)) != __float_as_int(__double2float_rn(ref)));
    ++cnt;
  }
  for (int m = 16; m; m >>= 1) {
    bad64 += __shfl_xor_sync(kFull, bad64, m);
    bad32 += __shfl_xor_sync(kFull, bad32, m);
    cnt += __shfl_xor_sync(kFull, cnt, m);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(counts, bad64); atomicAdd(counts + 1, bad32); atomicAdd(counts + 2, cnt);
  }
}

template <int C, bool NSIG4, typename OutT, int R>
__global__ void __launch_bounds__(NT, 4) window_kernel(const __grid_constant__ KParams P) {
  static_assert(SCGRHC_FLAT_WIN == 50, "run detection below is hard-wired to 49 = 32 + 16 + 1 pairs");
  static_assert(R * NWARP <= 32, "one mask word per lane of warp 0");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Scratch<R>& S = *reinterpret_cast<Scratch<R>*>(smem_raw);
  double* stage_base = reinterpret_cast<double*>(smem_raw + ((sizeof(Scratch<R>) + 127) & ~size_t(127)));

  const scgrhc_job& J = P.job;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = J.W, nsig = J.nsig, nstage = P.stages;
  const bool use_list = (J.flags & SCGRHC_USE_KEPT_LIST) != 0;
  const bool pred_only = (J.flags & SCGRHC_PREDICATES_ONLY) != 0;
  const bool norm_global = (J.flags & SCGRHC_NORM_GLOBAL) != 0;
  const bool keep_all = (J.flags & SCGRHC_KEEP_ALL) != 0;
  const double thr = J.flat_threshold, min_rhc = J.min_rhc;
  const int rcol = J.rhc_col;
  int col[C];
#pragma unroll
  for (int c = 0; c < C; ++c) col[c] = J.scg_cols[c];

  const long long items = use_list ? J.n_items : J.n_cand;
  const long long lo = items * (long long)blockIdx.x / gridDim.x;
  const long long hi = items * (long long)(blockIdx.x + 1) / gridDim.x;
  if (lo >= hi) return;

  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) mbar_init(&S.full[s], 1);
    S.slow_cnt[0] = S.slow_cnt[1] = 0;
    fence_barrier_init();
  }
  __syncthreads();

  // ---- producer state (thread 0 only) -----------------------------------------------------------
  int p_iv = 0;
  long long p_cand0 = 0, p_row0 = 0;
  int p_nwin = 0, p_rec = 0;
  auto load_iv = [&](int iv) {
    const scgrhc_interval I = J.intervals[iv];
    p_cand0 = I.cand0; p_row0 = I.row0; p_nwin = I.n_win; p_rec = I.rec_id;
  };
  auto issue = [&](long long item, int s) {
    const long long cand = use_list ? J.kept_list[item] : item;
    while (cand >= p_cand0 + p_nwin) load_iv(++p_iv);
    const int i = (int)(cand - p_cand0);
    const long long elem0 = (p_row0 + (long long)i * W) * nsig;
    const int lead = (int)(elem0 & 1);
    const long long n_even = ((long long)W * nsig + lead + 1) & ~1LL;
    StageMeta m;
    m.cand = cand; m.slot = use_list ? item : cand; m.elem0 = elem0; m.win = i; m.rec = p_rec; m.lead = lead;
    m.fallback = (elem0 - lead + n_even > P.arena_elems_cap) ? 1 : 0;
    S.meta[s] = m;
    if (m.fallback) {
      mbar_arrive(&S.full[s]);
    } else {
      const uint32_t bytes = (uint32_t)(n_even * 8);
      mbar_arrive_expect_tx(&S.full[s], bytes);
      bulk_g2s(stage_base + (size_t)s * P.stage_elems, J.arena + (elem0 - lead), bytes, &S.full[s]);
    }
  };
  if (tid == 0) {
    const long long first = use_list ? J.kept_list[lo] : lo;
    int a = 0, b = J.n_intervals - 1;  // last interval with cand0 <= first
    while (a < b) {
      const int mid = (a + b + 1) >> 1;
      if (J.intervals[mid].cand0 <= first) a = mid; else b = mid - 1;
    }
    p_iv = a;
    load_iv(a);
    for (int s = 0; s < nstage && lo + s < hi; ++s) issue(lo + s, s);
  }

  const double xbar = 0.5 * (double)(W - 1);
  const double sxx = (double)W * ((double)W * (double)W - 1.0) / 12.0;  // sum (t - xbar)^2, exact here
  const double inv_w = 1.0 / (double)W;

  for (long long n = 0; n < hi - lo; ++n) {
    const int s = (int)(n % nstage);
    const uint32_t parity = (uint32_t)((n / nstage) & 1);
    mbar_wait(&S.full[s], parity);
    const StageMeta M = S.meta[s];
    double* sbuf = stage_base + (size_t)s * P.stage_elems;
    if (M.fallback) {  // capacity edge: plain loads, rare
      const long long ne = (long long)W * nsig;
      for (long long e = tid; e < ne; e += NT) sbuf[M.lead + e] = J.arena[M.elem0 + e];
      __syncthreads();
    }
    const double* win = sbuf + M.lead;

    // ---- rows -> registers ----------------------------------------------------------------------
    double x[R][C], y[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int t = tid + k * NT;
      y[k] = 0.0;
#pragma unroll
      for (int c = 0; c < C; ++c) x[k][c] = 0.0;
      if (t < W) {
        if constexpr (NSIG4) {
          const double2 a = *reinterpret_cast<const double2*>(win + 4 * t);
          const double2 b = *reinterpret_cast<const double2*>(win + 4 * t + 2);
#pragma unroll
          for (int c = 0; c < C; ++c) x[k][c] = sel4(col[c], a.x, a.y, b.x, b.y);
          y[k] = sel4(rcol, a.x, a.y, b.x, b.y);
        } else {
          const double* row = win + (size_t)t * nsig;
#pragma unroll
          for (int c = 0; c < C; ++c) x[k][c] = row[col[c]];
          y[k] = row[rcol];
        }
      }
    }

    double smin = 0, smax = 0, ymin = 0, ymax = 0;
    uint32_t reason = 0;
    bool keep = true;

    if (!use_list) {
      // ---- pass 1: min/max, sum(y), floor / non-finite flags, small-step bits --------------------
      double a_smin = CUDART_INF, a_smax = -CUDART_INF, a_ymin = CUDART_INF, a_ymax = -CUDART_INF, a_sum = 0.0;
      uint32_t f = 0;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int t = tid + k * NT;
        if (t < W) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const double v = x[k][c];
            a_smin = fmin(a_smin, v);
            a_smax = fmax(a_smax, v);
            if (v != v) f |= 4u;
          }
          const double v = y[k];
          a_ymin = fmin(a_ymin, v);
          a_ymax = fmax(a_ymax, v);
          a_sum = __dadd_rn(a_sum, v);
          if (v < min_rhc) f |= 1u;
          if (!(fabs(v) <= DBL_MAX)) f |= 2u;
        }
        // c[t] = fl(|y[t+1] - y[t]|) < thr: necessary for any flat 50-window covering (t, t+1)
        double yn = __shfl_down_sync(kFull, y[k], 1);
        if (lane == 31 && t + 1 < W) yn = win[(size_t)(t + 1) * nsig + rcol];
        const bool cb = (t + 1 < W) && (fabs(__dsub_rn(yn, y[k])) < thr);
        const uint32_t word = __ballot_sync(kFull, cb);
        if (lane == 0) S.cmask[k * NWARP + warp] = word;
      }
      a_smin = warp_min(a_smin); a_smax = warp_max(a_smax);
      a_ymin = warp_min(a_ymin); a_ymax = warp_max(a_ymax);
      a_sum = warp_sum(a_sum);
      f = __reduce_or_sync(kFull, f);
      if (lane == 0) {
        S.red1[warp][0] = a_smin; S.red1[warp][1] = a_smax; S.red1[warp][2] = a_ymin;
        S.red1[warp][3] = a_ymax; S.red1[warp][4] = a_sum;
        S.wflags[warp] = f;
      }
      __syncthreads();  // #1
      if (tid == 0) S.slow_cnt[(n + 1) & 1] = 0;  // everyone is past iteration n-1, which used this slot

      smin = S.red1[0][0]; smax = S.red1[0][1]; ymin = S.red1[0][2]; ymax = S.red1[0][3];
      double ysum = S.red1[0][4];
      f = S.wflags[0];
#pragma unroll
      for (int w = 1; w < NWARP; ++w) {
        smin = fmin(smin, S.red1[w][0]); smax = fmax(smax, S.red1[w][1]);
        ymin = fmin(ymin, S.red1[w][2]); ymax = fmax(ymax, S.red1[w][3]);
        ysum = __dadd_rn(ysum, S.red1[w][4]);
        f |= S.wflags[w];
      }
      if (f & 4u) { smin = smax = __longlong_as_double(0x7ff8000000000000LL); }  // np.min/np.max propagate NaN

      // ---- pass 2: centred sums for R^2; warp 0 also looks for >= 49 consecutive small steps -----
      const double ybar = __dmul_rn(ysum, inv_w);
      double sxy = 0.0, syy = 0.0;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int t = tid + k * NT;
        if (t < W) {
          const double dy = __dsub_rn(y[k], ybar);
          sxy = __fma_rn((double)t - xbar, dy, sxy);
          syy = __fma_rn(dy, dy, syy);
        }
      }
      sxy = warp_sum(sxy); syy = warp_sum(syy);
      if (lane == 0) { S.red2[warp][0] = sxy; S.red2[warp][1] = syy; }
      if (warp == 0) {
        constexpr int NWORDS = R * NWARP;
        const uint32_t a1 = lane < NWORDS ? S.cmask[lane] : 0u;
        auto down = [&](uint32_t v, int d) {  // word (lane + d) of the mask, 0 past the end
          const uint32_t o = __shfl_down_sync(kFull, v, d);
          return (lane + d < 32) ? o : 0u;
        };
        auto shr = [&](uint32_t v, int sft) { return __funnelshift_r(v, down(v, 1), sft); };
        const uint32_t a2 = a1 & shr(a1, 1);
        const uint32_t a4 = a2 & shr(a2, 2);
        const uint32_t a8 = a4 & shr(a4, 4);
        const uint32_t a16 = a8 & shr(a8, 8);
        const uint32_t a32 = a16 & shr(a16, 16);
        const uint32_t a48 = a32 & down(a16, 1);
        const uint32_t a49 = a48 & __funnelshift_r(down(a1, 1), down(a1, 2), 16);
        if (lane < NWORDS) S.a49[lane] = a49;
        const int any = __any_sync(kFull, a49 != 0u);
        if (lane == 0) S.need_slow = any;
      }
      __syncthreads();  // #2

      sxy = S.red2[0][0]; syy = S.red2[0][1];
#pragma unroll
      for (int w = 1; w < NWARP; ++w) { sxy = __dadd_rn(sxy, S.red2[w][0]); syy = __dadd_rn(syy, S.red2[w][1]); }
      const double r2 = __ddiv_rn(__dmul_rn(sxy, sxy), __dmul_rn(sxx, syy));

      int flat_cnt = 0;
      if (S.need_slow) {  // exact rolling range, only where 49 consecutive small steps allow a flat window
        int cnt = 0;
        for (int k = 0; k < R; ++k) {
          const int p = tid + k * NT;
          if (p + SCGRHC_FLAT_WIN <= W && ((S.a49[p >> 5] >> (p & 31)) & 1u)) {
            double mx = -CUDART_INF, mn = CUDART_INF;
            for (int i = 0; i < SCGRHC_FLAT_WIN; ++i) {
              const double v = win[(size_t)(p + i) * nsig + rcol];
              mx = fmax(mx, v); mn = fmin(mn, v);
            }
            cnt += (__dsub_rn(mx, mn) < thr) ? 1 : 0;
          }
        }
        if (cnt) atomicAdd(&S.slow_cnt[n & 1], cnt);
        __syncthreads();  // #3 (rare)
        flat_cnt = S.slow_cnt[n & 1];
      }
      if (flat_cnt >= 2) reason |= SCGRHC_REASON_FLAT;
      if (r2 > 0.8) reason |= SCGRHC_REASON_STRAIGHT;
      if (fabs(r2 - 0.8) < 1e-12) reason |= SCGRHC_REASON_AMBIGUOUS;
      if (f & 1u) reason |= SCGRHC_REASON_FLOOR;
      if (f & 2u) reason |= SCGRHC_REASON_NONFINITE;
      keep = keep_all ||
             (reason & (SCGRHC_REASON_FLAT | SCGRHC_REASON_STRAIGHT | SCGRHC_REASON_FLOOR | SCGRHC_REASON_NONFINITE)) == 0;

      if (tid == 0) {
        P.out.keep[M.cand] = keep ? 1 : 0;
        P.out.reason[M.cand] = (uint8_t)reason;
        double2* mm = reinterpret_cast<double2*>(P.out.minmax + 4 * M.cand);
        mm[0] = make_double2(smin, smax);
        mm[1] = make_double2(ymin, ymax);
        P.out.cand_win[M.cand] = M.win;
        P.out.cand_rec[M.cand] = M.rec;
        if (!keep_all && (reason & SCGRHC_REASON_NONFINITE) && !(reason & SCGRHC_REASON_FLAT)) {
          atomicOr(P.err, 1ull);
          atomicMin(P.err + 1, (unsigned long long)M.cand);
        }
      }
    } else {
      __syncthreads();  // all rows are in registers before the stage is refilled
      if (!norm_global) {  // dense re-materialisation with the per-window pairs of an earlier pass
        const double2* mm = reinterpret_cast<const double2*>(P.out.minmax + 4 * M.cand);
        const double2 a = mm[0], b = mm[1];
        smin = a.x; smax = a.y; ymin = b.x; ymax = b.y;
      }
    }

    // ---- the stage buffer is dead: refill it with window n + stages ---------------------------
    if (tid == 0 && lo + n + nstage < hi) issue(lo + n + nstage, s);

    // ---- normalise from registers, transpose, cast, store ---------------------------------------
    if (keep && !pred_only) {
      if (norm_global) { smin = J.global_minmax[0]; smax = J.global_minmax[1]; ymin = J.global_minmax[2]; ymax = J.global_minmax[3]; }
      Normaliser ns, nr;
      ns.init(smin, smax);
      nr.init(ymin, ymax);
      OutT* so = reinterpret_cast<OutT*>(P.out.scg_out) + (size_t)M.slot * C * W;
      OutT* ro = reinterpret_cast<OutT*>(P.out.rhc_out) + (size_t)M.slot * W;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int t = tid + k * NT;
        if (t < W) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            OutT o;
            cvt_out(o, ns(x[k][c]));
            st_cs(so + (size_t)c * W + t, o);
          }
          OutT o;
          cvt_out(o, nr(y[k]));
          st_cs(ro + t, o);
        }
      }
    }
  }
}

}  // namespace scgrhc
