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
// read from shared memory once; reductions are warp shuffles + one small shared exchange and ONE
// __syncthreads per window.
//
// Instruction economy (the kernel is issue-bound before it is HBM-bound; see profiles/):
//   * no per-element division: one correctly rounded reciprocal per window + two FMA residual
//     corrections (Markstein) give the correctly rounded quotient; windows whose scale could make the
//     residuals underflow take a separate, non-unrolled IEEE-division loop (CTA-uniform branch);
//   * below-floor is decided from the window minimum, non-finite RHC from the sum of squares, NaN in
//     SCG from a 0*v accumulator — no per-element flag logic;
//   * R^2 from one pass of sums shifted by K = y[0] (cancellation bounded by n because K is a sample);
//   * the flat-line test first asks for >= 49 consecutive small steps |y[t+1]-y[t]| < thr (a necessary
//     condition, one bit per sample, checked with shifts); only then the exact rolling range runs.
#pragma once
#include <float.h>
#include <math_constants.h>

#include "common.cuh"

namespace scgrhc {

constexpr int NT = 128;          // threads per CTA
constexpr int NWARP = NT / kWarp;
constexpr int kMaxStages = 8;
constexpr int RMAX = 8;          // rows per thread, generic kernel: W <= RMAX * NT

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
  unsigned long long* err;  // [0] error flags, [1] first offending candidate (atomicMin), [2] windows flagged AMBIGUOUS
  int stages;
  int stage_elems;          // doubles per stage buffer (even, >= W*nsig + nsig + 2)
  long long arena_elems_cap;
};

// Decimating front end (extension): the arena holds the records at their NATIVE rate; every candidate window
// of W rows at the model rate is produced on the fly by the polyphase FIR of scipy.signal.resample_poly (up == 1, integer
// `down`; every tap a separately rounded multiply and add, oldest sample first: bit-identical to the standalone
// decimator and hence to scipy), straight from the staged native-rate rows into the shared-memory window the rest of the
// kernel reads.  The resampled cohort never exists in HBM.
constexpr int kDecimMaxTaps = 128;
constexpr int kDecimR = 3;           // consecutive outputs per thread: W <= kDecimR * NT
struct DecimK {
  double taps[kDecimMaxTaps];        // scipy's flipped polyphase table for up == 1 (resample_design)
  int pp, down, npr;                 // taps per output, decimation factor, n_pre_remove
  int rows_in;                       // native-rate rows one window needs: (W - 1) * down + pp
  unsigned pb_magic;                 // r / (kDecimR * down) == umulhi(r, pb_magic)
  const long long* iv_in0;           // device (n_intervals): arena row of the first native-rate row of the interval's record
  const long long* iv_len;           // device: native-rate rows of that record
  const long long* iv_rel;           // device: first model-rate row of the interval, relative to its record
};
struct KParamsDecim {
  KParams k;
  DecimK d;
};
template <bool DECIM> struct KParamsOf { typedef KParams type; };
template <> struct KParamsOf<true> { typedef KParamsDecim type; };
__device__ __forceinline__ const KParams& kparams_base(const KParams& p) { return p; }
__device__ __forceinline__ const KParams& kparams_base(const KParamsDecim& p) { return p.k; }

// Conflict-free row loads in the identity-column variant (see the kernel).  Measured in round 2 and NOT adopted: shared-memory
// bank conflicts 154 M -> 21 M per launch, but the selects that put the halves back cost +16 % instructions (3,331 -> 3,875 per
// window): burst 2.22 -> 2.26 ms, sustained 2.62 -> 2.56 ms, i.e. within noise of each other.
constexpr bool kSwizzleLoads = false;
constexpr int NRED = 8;  // smin, smax, ymin, ymax, s1, s2 (+ 0*v NaN accumulator), sxy, run-candidate flag

struct Scratch {
  uint64_t full[kMaxStages];
  StageMeta meta[kMaxStages];
  double red[2][NWARP][NRED];        // double buffered by window parity: one __syncthreads per window
  uint32_t cmask[2][RMAX * NWARP];
  uint32_t a49[RMAX * NWARP];
  int slow_cnt;
  int slow_flag;
  double zred[NWARP][2];             // extension: shifted SCG sums of the z-score mode
};

struct Normaliser {
  double mn, d, inv;
  bool slow, quick;
  __device__ __forceinline__ void init(double mn_, double mx_) {
    mn = mn_;
    d = __dadd_rn(__dsub_rn(mx_, mn_), 0.0001);  // (max - min + 0.0001), recordutil.py:46
    inv = __drcp_rn(d);
    // The fast path needs every non-zero numerator a = x - mn to satisfy |a| >= 2^-960 (then the
    // residuals and the quotient stay normal): true when |mn| >= 2^-900 (a is then either >= |mn|/2
    // or a multiple of ulp(mn)/2) and d is within [2^-14, 2^60].  Anything else (mn == 0, tiny, NaN,
    // Inf, huge ranges) takes the IEEE-division loop.
    // The three range tests on the exponent fields (integer pipe; every thread of the CTA evaluates them for every window):
    // 2^-14 <= d < 2^60 (negative, NaN and Inf fall outside: the subtraction wraps), |mn| >= 2^-900, e_mn - e_d >= -70.  A NaN /
    // Inf minimum passes the |mn| test but makes d NaN / Inf, which the first test rejects.
    const uint32_t hd = (uint32_t)__double2hiint(d), hm = (uint32_t)__double2hiint(mn_) & 0x7fffffffu;
    slow = !((hd - 0x3f100000u) < (0x43b00000u - 0x3f100000u) && hm >= 0x07b00000u);
    // Tier 1 (below) needs every non-zero quotient to be a NORMAL float with room to spare (>= 2^-124): a non-zero numerator is
    // at least ulp(mn) / 2 = 2^(e_mn - 53), the quotient therefore at least 2^(e_mn - e_d - 54), so e_mn - e_d >= -70 is enough
    // (|mn| >= 2^-10 with d <= 2^60, the earlier form of this test, is the special case at the edge of the range).
    quick = !slow && (hm >> 20) + 70u >= (hd >> 20);
  }
  // extension (absent from the reference): (x - mean) / (std + 0.0001); same tiers, same validity conditions
  __device__ __forceinline__ void init_z(double mean, double sd) {
    mn = mean;
    d = __dadd_rn(sd, 0.0001);
    inv = __drcp_rn(d);
    slow = !(d >= 0x1p-14 && d <= 0x1p60 && fabs(mean) >= 0x1p-900);
    quick = !slow && fabs(mean) >= 0x1p-10;
  }
  __device__ __forceinline__ double fast(double x) const { return div_by_recip(__dsub_rn(x, mn), d, inv); }
  __device__ __forceinline__ double exact(double x) const { return __ddiv_rn(__dsub_rn(x, mn), d); }
};

__device__ __forceinline__ void cvt_out(float& o, double q) { o = __double2float_rn(q); }
__device__ __forceinline__ void cvt_out(double& o, double q) { o = q; }

// np.add.reduce over n copies of v in numpy's pairwise order (numpy/_core/src/umath/loops_utils.h.src, *_pairwise_sum:
// < 8 elements serial from 0; <= 128 eight strided accumulators combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then the
// tail; else split at n/2 rounded down to a multiple of 8).  For a constant vector the eight accumulators are equal.
// Used for exactly constant RHC windows only: sklearn's R^2 of a constant y is rounding noise, 1.0 iff np.mean(y) is
// exact (oracle/scgrhc_oracle.py: r_squared).  D bounds the recursion statically: n <= 128 * 2^D.
template <int D>
__device__ __forceinline__ double np_sum_const(double v, int n) {
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = __dadd_rn(r, v);
    return r;
  }
  if (n <= 128 || D == 0) {
    double r = v;
    for (int i = 8; i < n - (n % 8); i += 8) r = __dadd_rn(r, v);
    r = __dadd_rn(r, r); r = __dadd_rn(r, r); r = __dadd_rn(r, r);
    for (int i = n - (n % 8); i < n; ++i) r = __dadd_rn(r, v);
    return r;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __dadd_rn(np_sum_const<(D > 0 ? D - 1 : 0)>(v, n2), np_sum_const<(D > 0 ? D - 1 : 0)>(v, n - n2));
}
__device__ __noinline__ bool np_mean_of_const_is_exact(double v, int n) {
  return __ddiv_rn(np_sum_const<3>(v, n), (double)n) == v;      // W <= RMAX * NT = 1024 = 128 * 2^3
}

__device__ __forceinline__ double sel4(int col, double v0, double v1, double v2, double v3) {
  double lo = (col & 1) ? v1 : v0;
  double hi = (col & 1) ? v3 : v2;
  return (col & 2) ? hi : lo;
}

// ---- diagnostics: the normaliser's fast path vs IEEE division on hashed operands -------------------
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
    double mn, mx, x;
    if (mode == 0) {         // shaped like the normalisation: mn <= x <= mx, ranges 2^-14 .. 2^10
      const double range = ldexp(m1, (int)(h1 & 0xFFF) % 24 - 14);
      mn = ldexp((double)((h2 >> 3) & 0xFFFFF) - 524288.5, (int)(h1 >> 52) % 30 - 25);
      mx = mn + range;
      x = mn + range * m2;
      if (x > mx) x = mx;
    } else {                 // wide exponents, both signs, zeros: exercises the slow-path selection
      mn = ldexp(m1, (int)(h1 & 0xFFF) % 2001 - 1000);
      if (h1 & 0x1000) mn = -mn;
      mx = mn + ldexp(1.0 + m2, (int)(h2 & 0xFFF) % 2001 - 1000);
      x = mn + (mx - mn) * m2;
      if (h2 & 0x2000) mn = 0.0;
    }
    if (mode >= 2) {         // tier 1 of the fp32 normalisation, half of the quotients planted next to float midpoints
      const double range = ldexp(m1, (int)(h1 & 0xFFF) % 24 - 14);
      // mode 2: |mn| from 2^-16 up; mode 3: minima down to 2^-86 (tier 1's eligibility compares the exponents of mn and d) and
      // a third of the samples a few ulp(mn) above the minimum: the smallest non-zero quotients there are
      mn = mode == 2 ? ldexp((double)((h2 >> 3) & 0xFFFFF) - 524288.5, (int)(h1 >> 52) % 20 - 15)
                     : ldexp((double)((h2 >> 3) & 0xFFFFF) - 524288.5, (int)(h1 >> 52) % 70 - 85);
      mx = mn + range;
      const double d = __dadd_rn(__dsub_rn(mx, mn), 0.0001);
      if (h2 & 4) {
        const float f = (float)m2;
        const double mid = 0.5 * ((double)f + (double)nextafterf(f, 2.0f));
        x = mn + mid * d * (1.0 + ((double)((h1 >> 40) & 15) - 8.0) * 0x1p-53);
      } else if (mode == 3 && (h2 & 3) == 1) {
        x = mn + fabs(mn) * 0x1p-53 * (double)(1 + ((h1 >> 40) & 1023));
      } else {
        x = mn + range * m2;
      }
      if (x < mn) x = mn;
    }
    Normaliser nz;
    nz.init(mn, mx);
    const double ref = __ddiv_rn(__dsub_rn(x, mn), __dadd_rn(__dsub_rn(mx, mn), 0.0001));
    if (mode >= 2) {
      if (nz.quick) {
        const double q0 = __dmul_rn(__dsub_rn(x, nz.mn), nz.inv);
        if (tier1_key(q0) <= kTier1Risky) ++cnt;      // risky: the kernel recomputes these exactly
        else bad32 += (__float_as_int(__double2float_rn(q0)) != __float_as_int(__double2float_rn(ref)));
      } else {
        ++bad64;                     // operands of this mode must all qualify for tier 1
      }
      continue;
    }
    const double q = nz.slow ? nz.exact(x) : nz.fast(x);
    const bool both_nan = (q != q) && (ref != ref);
    bad64 += (__double_as_longlong(q) != __double_as_longlong(ref)) && !both_nan;
    bad32 += (__float_as_int(__double2float_rn(q)) != __float_as_int(__double2float_rn(ref))) && !both_nan;
    cnt += nz.slow ? 0 : 1;
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

// C      SCG channels;  NSIG4  rows are 4 doubles (16-byte shared loads);
// IDENT  columns are (0..C-1 | C) in order (no selects);  WCT  > 0: compile-time window length; 0: runtime length, RMAX rows per thread;
//        < 0: runtime length <= -WCT * NT with -WCT rows per thread (resampled cohorts: 375 samples = 3 rows)
// DDOWN / DPP > 0 (DECIM only): compile-time decimation factor and taps per output — the FIR is then fully unrolled, the
// taps are immediate constant-bank operands and the pad-row bookkeeping folds away (500 -> 250 Hz: down 2, 43 taps).
// DFMA (DECIM only): one fused multiply-add per tap instead of a separately rounded multiply and add — half the fp64
// instructions, within ~1e-15 of scipy instead of bit-identical (the `fused` mode of scgrhc_resample_poly).
// PLAIN: the job has no mode flag set (not USE_KEPT_LIST / PREDICATES_ONLY / NORM_GLOBAL / KEEP_ALL / NORM_ZSCORE): the one-pass
// local-min-max job of save_dataloaders.  The mode tests then fold away at compile time instead of being uniform branches in
// the per-window loop.
template <int C, bool NSIG4, bool IDENT, typename OutT, int WCT, bool DECIM = false, int DDOWN = 0, int DPP = 0, bool DFMA = false,
          bool PLAIN = false>
__global__ void __launch_bounds__(NT, DECIM ? 5 : 4) window_kernel(const __grid_constant__ typename KParamsOf<DECIM>::type PP) {
  const KParams& P = kparams_base(PP);
  static_assert(!DECIM || (NSIG4 && ((WCT < 0 && -WCT <= kDecimR) || (WCT > 0 && (WCT + NT - 1) / NT <= kDecimR))),
                "the decimating front end feeds the 4-signal, <= 3 rows per thread variants");
  static_assert(SCGRHC_FLAT_WIN == 50, "run detection below is hard-wired to 49 = 32 + 16 + 1 pairs");
  constexpr int R = WCT > 0 ? (WCT + NT - 1) / NT : (WCT < 0 ? -WCT : RMAX);
  static_assert(R * NWARP <= 32 && R <= RMAX, "one mask word per lane");
  static_assert(!IDENT || !NSIG4 || C == 3, "identity mapping: rows are exactly the C SCG columns followed by RHC (nsig == C + 1)");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Scratch& S = *reinterpret_cast<Scratch*>(smem_raw);
  double* stage_base = reinterpret_cast<double*>(smem_raw + ((sizeof(Scratch) + 127) & ~size_t(127)));

  const scgrhc_job& J = P.job;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = WCT > 0 ? WCT : J.W;
  const int nsig = NSIG4 ? 4 : (IDENT ? C + 1 : J.nsig);   // IDENT: what the drop-in uploads (the selected columns, in order, then RHC)
  const int nstage = P.stages;
  const int wstride = J.stride > 0 ? J.stride : W;
  const bool use_list = !PLAIN && (J.flags & SCGRHC_USE_KEPT_LIST) != 0;
  const bool pred_only = !PLAIN && (J.flags & SCGRHC_PREDICATES_ONLY) != 0;
  const bool norm_global = !PLAIN && (J.flags & SCGRHC_NORM_GLOBAL) != 0;
  const bool keep_all = !PLAIN && (J.flags & SCGRHC_KEEP_ALL) != 0;
  const bool zscore = !PLAIN && (J.flags & SCGRHC_NORM_ZSCORE) != 0;
  const double thr = J.flat_threshold, min_rhc = J.min_rhc;
  const int rcol = IDENT ? C : J.rhc_col;
  int col[C];
#pragma unroll
  for (int c = 0; c < C; ++c) col[c] = IDENT ? c : J.scg_cols[c];

  const long long items = use_list ? J.n_items : J.n_cand;
  const long long lo = items * (long long)blockIdx.x / gridDim.x;
  const long long hi = items * (long long)(blockIdx.x + 1) / gridDim.x;
  if (lo >= hi) return;

  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) mbar_init(&S.full[s], 1);
    S.slow_cnt = 0;
    S.slow_flag = 0;
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
    const long long elem0 = (p_row0 + (long long)i * wstride) * nsig;
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
  // ---- decimating front end: every thread walks the intervals itself (two cursors: prefetch and consume) and stages its
  //      share of the native-rate rows with 16-byte cp.async copies (zero fill outside the record = upfirdn's zero padding)
  double* wbuf = nullptr;               // the model-rate window the FIR writes: (W + 1) rows of 4 doubles
  double* s_taps = nullptr;
  int c_iv = 0;                         // consumer cursor
  long long c_cand0 = 0, c_in0 = 0, c_len = 0, c_rel = 0;
  int c_nwin = 0, c_rec = 0;
  long long q_in0 = 0, q_len = 0, q_rel = 0;   // producer cursor shares p_iv / p_cand0 / p_nwin
  auto stage_in = [&](long long item, int st) {
    if constexpr (DECIM) {
      const DecimK& D = PP.d;
      while (item >= p_cand0 + p_nwin) {
        ++p_iv;
        const scgrhc_interval I = J.intervals[p_iv];
        p_cand0 = I.cand0; p_nwin = I.n_win;
        q_in0 = D.iv_in0[p_iv]; q_len = D.iv_len[p_iv]; q_rel = D.iv_rel[p_iv];
      }
      const long long m0 = q_rel + (item - p_cand0) * wstride;            // first model-rate row of the window in its record
      const long long first = (m0 + D.npr) * D.down - D.pp + 1;          // native-rate row of staged row 0 (may be < 0)
      double* dst = stage_base + (size_t)st * P.stage_elems;
      // chunk q = tid + i * NT of the window's 16-byte chunks: row r = (tid >> 1) + 64 i, half h = tid & 1 (fixed per thread)
      const int h = tid & 1;
      const double* src0 = J.arena + q_in0 * 4 + 2 * h;
      for (int r = tid >> 1; r < D.rows_in; r += NT / 2) {
        const long long xi = first + r;
        const bool in = xi >= 0 && xi < q_len;
        const int pr = r + (int)__umulhi((unsigned)r, D.pb_magic);       // one pad row per thread span: conflict-free FIR loads
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst + (size_t)pr * 4 + 2 * h)),
                     "l"(src0 + (in ? xi : 0) * 4), "r"(in ? 16 : 0) : "memory");
      }
    }
  };
  if constexpr (DECIM) {
    const DecimK& D = PP.d;
    wbuf = stage_base + (size_t)P.stage_elems;
    s_taps = wbuf + (size_t)(W + 1) * 4;
    for (int i = tid; i < D.pp; i += NT) s_taps[i] = D.taps[i];
    int a = 0, b = J.n_intervals - 1;  // last interval with cand0 <= lo
    while (a < b) {
      const int mid = (a + b + 1) >> 1;
      if (J.intervals[mid].cand0 <= lo) a = mid; else b = mid - 1;
    }
    const scgrhc_interval I = J.intervals[a];
    p_iv = c_iv = a;
    p_cand0 = c_cand0 = I.cand0; p_nwin = c_nwin = I.n_win; c_rec = I.rec_id;
    q_in0 = c_in0 = D.iv_in0[a]; q_len = c_len = D.iv_len[a]; q_rel = c_rel = D.iv_rel[a];
    // ONE raw stage per CTA (5 CTAs per SM instead of 3): the copy of window n+1 is issued as soon as the FIR of window n has
    // drained the stage and overlaps the statistics / normalisation / stores of window n; the other CTAs cover the rest
    stage_in(lo, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
  } else if (tid == 0) {
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
  const double tx = (double)tid - xbar;
  const double inv_w = 1.0 / (double)W;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);

  int s = 0;
  uint32_t parity = 0, par2 = 0;
  for (long long n = 0; n < hi - lo; ++n) {
    StageMeta M;
    const double* win;
    if constexpr (DECIM) {
      const DecimK& D = PP.d;
      asm volatile("cp.async.wait_group 0;" ::: "memory");      // this thread's copies of window n have landed
      __syncthreads();                                            // ... and everybody else's; the previous window is out of wbuf
      {
        // polyphase FIR, kDecimR consecutive outputs per thread from ONE sliding pass over pp + (kDecimR - 1) * down staged
        // rows (the arithmetic and the conflict-free staging of resample_decim_kernel, filter_kernels.cuh)
        const int down = D.down, pp = D.pp, PB = kDecimR * down;
        const int swap = (tid >> 2) & 1;
        const double* base = stage_base + (size_t)tid * (PB + 1) * 4;
        double acc[kDecimR][4];
#pragma unroll
        for (int r = 0; r < kDecimR; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
        if constexpr (DPP > 0) {
          if (tid * kDecimR < W) {
            constexpr int CPB = kDecimR * DDOWN, CJ = DPP + (kDecimR - 1) * DDOWN;
#pragma unroll
            for (int j = 0; j < CJ; ++j) {
              const double* prow = base + (size_t)(j + j / CPB) * 4;       // compile-time offset: one pad row per CPB rows
              const double2 a = *reinterpret_cast<const double2*>(prow + (swap ? 2 : 0));
              const double2 b = *reinterpret_cast<const double2*>(prow + (swap ? 0 : 2));
#pragma unroll
              for (int r = 0; r < kDecimR; ++r) {
                if (j - r * DDOWN >= 0 && j - r * DDOWN < DPP) {
                  const double hk = D.taps[j - r * DDOWN >= 0 && j - r * DDOWN < DPP ? j - r * DDOWN : 0];   // constant-bank operand
                  if constexpr (DFMA) {
                    acc[r][0] = __fma_rn(a.x, hk, acc[r][0]); acc[r][1] = __fma_rn(a.y, hk, acc[r][1]);
                    acc[r][2] = __fma_rn(b.x, hk, acc[r][2]); acc[r][3] = __fma_rn(b.y, hk, acc[r][3]);
                  } else {
                    acc[r][0] = __dadd_rn(acc[r][0], __dmul_rn(a.x, hk)); acc[r][1] = __dadd_rn(acc[r][1], __dmul_rn(a.y, hk));
                    acc[r][2] = __dadd_rn(acc[r][2], __dmul_rn(b.x, hk)); acc[r][3] = __dadd_rn(acc[r][3], __dmul_rn(b.y, hk));
                  }
                }
              }
            }
          }
        } else if (tid * kDecimR < W) {
          const int Jn = pp + (kDecimR - 1) * down, ramp = (kDecimR - 1) * down;
          auto load_row = [&](const double* p, double (&xv)[4]) {
            const double2 a = *reinterpret_cast<const double2*>(p + (swap ? 2 : 0));
            const double2 b = *reinterpret_cast<const double2*>(p + (swap ? 0 : 2));
            xv[0] = a.x; xv[1] = a.y; xv[2] = b.x; xv[3] = b.y;
          };
          auto some = [&](int j, const double (&xv)[4]) {   // ramp-up / ramp-down: output r takes rows r*down .. r*down + pp - 1
#pragma unroll
            for (int r = 0; r < kDecimR; ++r) {
              const int k = j - r * down;
              if (k >= 0 && k < pp) {
                const double hk = s_taps[k];
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = DFMA ? __fma_rn(xv[c], hk, acc[r][c]) : __dadd_rn(acc[r][c], __dmul_rn(xv[c], hk));
              }
            }
          };
          const double* pb = base;
          int j = 0, bend = PB;
          for (; j < ramp; ++j) {                             // ramp < PB: all in block 0
            double xv[4];
            load_row(pb + (size_t)j * 4, xv);
            some(j, xv);
          }
          while (j < pp) {                                    // steady state: every output takes the row
            const int je = min(pp, bend);
#pragma unroll 4
            for (; j < je; ++j) {
              double xv[4];
              load_row(pb + (size_t)j * 4, xv);
#pragma unroll
              for (int r = 0; r < kDecimR; ++r) {
                const double hk = s_taps[j - r * down];
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = DFMA ? __fma_rn(xv[c], hk, acc[r][c]) : __dadd_rn(acc[r][c], __dmul_rn(xv[c], hk));
              }
            }
            if (j == bend) { bend += PB; pb += 4; }
          }
          while (j < Jn) {
            const int je = min(Jn, bend);
            for (; j < je; ++j) {
              double xv[4];
              load_row(pb + (size_t)j * 4, xv);
              some(j, xv);
            }
            if (j == bend) { bend += PB; pb += 4; }
          }
        }
#pragma unroll
        for (int r = 0; r < kDecimR; ++r) {
          const int o = tid * kDecimR + r;
          if (o < W) {
            double2* dst = reinterpret_cast<double2*>(wbuf + (size_t)o * 4);
            dst[swap ? 1 : 0] = make_double2(acc[r][0], acc[r][1]);     // odd lane groups hold the two halves swapped
            dst[swap ? 0 : 1] = make_double2(acc[r][2], acc[r][3]);
          }
        }
      }
      __syncthreads();                                            // the window is in wbuf; stage s is free
      if (lo + n + 1 < hi) stage_in(lo + n + 1, 0);
      asm volatile("cp.async.commit_group;" ::: "memory");
      const long long item = lo + n;
      while (item >= c_cand0 + c_nwin) {
        ++c_iv;
        const scgrhc_interval I = J.intervals[c_iv];
        c_cand0 = I.cand0; c_nwin = I.n_win; c_rec = I.rec_id;
      }
      M.cand = item; M.slot = item; M.elem0 = 0; M.win = (int)(item - c_cand0); M.rec = c_rec; M.lead = 0; M.fallback = 0;
      win = wbuf;
    } else {
      mbar_wait(&S.full[s], parity);
      M = S.meta[s];
      double* sbuf = stage_base + (size_t)s * P.stage_elems;
      if (M.fallback) {  // capacity edge: plain loads, rare
        const long long ne = (long long)W * nsig;
        for (long long e = tid; e < ne; e += NT) sbuf[M.lead + e] = J.arena[M.elem0 + e];
        __syncthreads();
      }
      win = sbuf + M.lead;
    }

    // ---- rows -> registers (+ the next row's RHC sample for the small-step bit) --------------------
    double x[R][C], y[R], yn[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int t = tid + k * NT;
      const bool valid = WCT > 0 ? ((k + 1) * NT <= WCT || t < WCT) : (t < W);
      y[k] = 0.0; yn[k] = 0.0;
#pragma unroll
      for (int c = 0; c < C; ++c) x[k][c] = 0.0;
      if constexpr (NSIG4 && IDENT && kSwizzleLoads) {
        // Rows are 32 bytes: consecutive lanes reading the same half of their rows hit every other 16-byte bank group twice
        // (2-way conflict, 153.8 M per launch in round 1).  Odd groups of four lanes fetch the two halves in the opposite
        // order instead: each quarter-warp then covers all 32 banks.  The RHC sample of the next row comes from the next lane.
        if (valid) {
          const int sw = (tid >> 2) & 1;
          const double2 d0 = *reinterpret_cast<const double2*>(win + 4 * t + (sw ? 2 : 0));
          const double2 d1 = *reinterpret_cast<const double2*>(win + 4 * t + (sw ? 0 : 2));
          x[k][0] = sw ? d1.x : d0.x; x[k][1] = sw ? d1.y : d0.y; x[k][2] = sw ? d0.x : d1.x; y[k] = sw ? d0.y : d1.y;
        }
        const double up = __shfl_down_sync(kFull, y[k], 1);
        yn[k] = up;
        if (lane == 31 && valid) yn[k] = win[4 * (t + 1) + 3];      // row t+1 lives in the next warp (the stage is padded)
      } else if (valid) {
        if constexpr (NSIG4) {
          const double2 a = *reinterpret_cast<const double2*>(win + 4 * t);
          const double2 b = *reinterpret_cast<const double2*>(win + 4 * t + 2);
          const double2 nb = *reinterpret_cast<const double2*>(win + 4 * t + 6);  // row t+1, columns 2,3 (stage is padded)
          if constexpr (IDENT) {
            x[k][0] = a.x; x[k][1] = a.y; x[k][2] = b.x; y[k] = b.y; yn[k] = nb.y;
          } else {
            const double2 na = *reinterpret_cast<const double2*>(win + 4 * t + 4);
#pragma unroll
            for (int c = 0; c < C; ++c) x[k][c] = sel4(col[c], a.x, a.y, b.x, b.y);
            y[k] = sel4(rcol, a.x, a.y, b.x, b.y);
            yn[k] = sel4(rcol, na.x, na.y, nb.x, nb.y);
          }
        } else {
          const double* row = win + (size_t)t * nsig;
          if constexpr (IDENT && C == 1) {           // 16-byte rows: one load for (x, y)
            const double2 xy = *reinterpret_cast<const double2*>(row);
            x[k][0] = xy.x; y[k] = xy.y; yn[k] = row[3];
          } else {
#pragma unroll
            for (int c = 0; c < C; ++c) x[k][c] = row[col[c]];
            y[k] = row[rcol];
            yn[k] = row[nsig + rcol];
          }
        }
      }
    }

    double smin = 0, smax = 0, ymin = 0, ymax = 0;
    uint32_t reason = 0;
    bool keep = true;

    if (!use_list) {
      // ---- one pass of per-thread statistics --------------------------------------------------------
      const double K = win[rcol];  // shift for the sums: a sample of the window, so cancellation is bounded by n
      double a_smin = CUDART_INF, a_smax = -CUDART_INF, a_ymin = CUDART_INF, a_ymax = -CUDART_INF;
      double s1 = 0.0, s2 = 0.0, sB = 0.0, nanacc = 0.0;
      bool dense_word = false;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int t = tid + k * NT;
        const bool valid = WCT > 0 ? ((k + 1) * NT <= WCT || t < WCT) : (t < W);
        const bool has_next = WCT > 0 ? ((k + 1) * NT < WCT || t + 1 < WCT) : (t + 1 < W);
        bool cb = false;
        if (valid) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const double v = x[k][c];
            a_smin = v < a_smin ? v : a_smin;
            a_smax = v > a_smax ? v : a_smax;
            nanacc = __fma_rn(v, 0.0, nanacc);  // NaN iff some v is NaN or Inf
          }
          const double v = y[k];
          a_ymin = v < a_ymin ? v : a_ymin;
          a_ymax = v > a_ymax ? v : a_ymax;
          const double dy = __dsub_rn(v, K);
          s1 = __dadd_rn(s1, dy);
          s2 = __fma_rn(dy, dy, s2);
          sB = __fma_rn((double)k, dy, sB);
          // c[t] = fl(|y[t+1] - y[t]|) < thr: necessary for any flat 50-window covering (t, t+1)
          cb = has_next && (fabs(__dsub_rn(yn[k], v)) < thr);
        }
        const uint32_t word = __ballot_sync(kFull, cb);
        if (lane == 0) S.cmask[par2][k * NWARP + warp] = word;
        // 49 consecutive ones span at most 3 mask words, so one of them holds >= 17 of them
        dense_word |= __popc(word) >= 17;
      }
      // sum_k (tid + NT k - xbar) dy_k = (tid - xbar) sum dy + NT sum k dy
      double sxy = __fma_rn(tx, s1, __dmul_rn((double)NT, sB));
      a_smin = warp_min(a_smin); a_smax = warp_max(a_smax);
      a_ymin = warp_min(a_ymin); a_ymax = warp_max(a_ymax);
      // 0*v is +-0 for finite v, so adding the accumulator leaves s2 unchanged unless some v is NaN/Inf
      const double s3 = warp_sum3(s1, __dadd_rn(s2, nanacc), sxy, lane);   // lane 0: s1, lane 8: s2, lane 16: sxy
      double* r = S.red[par2][warp];
      if ((lane & 7) == 0 && lane < 24) r[4 + (lane >> 3)] = s3;
      if (lane == 0) {
        r[0] = -a_smin; r[1] = a_smax; r[2] = -a_ymin; r[3] = a_ymax;   // minima negated: slots 0..3 all take a max
        r[7] = dense_word ? 1.0 : 0.0;
      }
      __syncthreads();  // the only block barrier of the common path; every thread is also done with the stage

      // block combine, lane parallel: lane v & 7 folds slot v over the warps (same order as a serial fold: ties keep the
      // earlier warp, sums add warp 0 + 1 + 2 + 3), then the eight results are broadcast — ~40 instructions per thread
      // instead of ~90 for every thread folding all eight slots itself
      double dense;
      {
        const int v = lane & 7;
        double a = S.red[par2][0][v];
#pragma unroll
        for (int w = 1; w < NWARP; ++w) {
          const double b = S.red[par2][w][v];
          const double mx = b > a ? b : a;
          const double sm = __dadd_rn(a, b);
          a = v < 4 ? mx : sm;
        }
        smin = -__shfl_sync(kFull, a, 0); smax = __shfl_sync(kFull, a, 1);
        ymin = -__shfl_sync(kFull, a, 2); ymax = __shfl_sync(kFull, a, 3);
        s1 = __shfl_sync(kFull, a, 4); s2 = __shfl_sync(kFull, a, 5); sxy = __shfl_sync(kFull, a, 6);
        dense = __shfl_sync(kFull, a, 7);
      }
      // `dense` counts the warps that saw a dense mask word; only then scan the small-step mask for
      // >= 49 consecutive ones (every warp does, on the same shared words: no second barrier)
      constexpr int NWORDS = R * NWARP;
      uint32_t a49 = 0u;
      bool run49 = false;
      if (dense != 0.0) {
        const uint32_t a1 = lane < NWORDS ? S.cmask[par2][lane] : 0u;
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
        a49 = a48 & __funnelshift_r(down(a1, 1), down(a1, 2), 16);
        run49 = __any_sync(kFull, a49 != 0u);
      }
      // non-finite RHC or NaN/Inf SCG <=> the shifted sum of squares (+ the 0*v accumulator) is not finite
      // (or it overflowed): recheck exactly.  NaN in SCG must poison the joint min/max as np.min/np.max do.
      const bool s2_bad = !(s2 <= DBL_MAX);
      const bool scg_bad = s2_bad;

      int flat_cnt = 0;
      bool nonfinite = false;
      if (run49 || s2_bad || scg_bad) {  // CTA-uniform and rare: exact work on the stage buffer / registers
        if (warp == 0 && lane < NWORDS) S.a49[lane] = a49;
        __syncthreads();
        int cnt = 0, fl = 0;
        if (run49) {
          for (int k = 0; k < R; ++k) {
            const int p = tid + k * NT;
            if (p + SCGRHC_FLAT_WIN <= W && ((S.a49[p >> 5] >> (p & 31)) & 1u)) {
              double mx = -CUDART_INF, mn = CUDART_INF;
              for (int i = 0; i < SCGRHC_FLAT_WIN; ++i) {
                const double v = win[(size_t)(p + i) * nsig + rcol];
                mx = v > mx ? v : mx; mn = v < mn ? v : mn;
              }
              cnt += (__dsub_rn(mx, mn) < thr) ? 1 : 0;
            }
          }
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int t = tid + k * NT;
          if (t < W) {
            if (!(fabs(y[k]) <= DBL_MAX)) fl |= 1;
#pragma unroll
            for (int c = 0; c < C; ++c) if (x[k][c] != x[k][c]) fl |= 2;
          }
        }
        if (s2_bad) {
          // s2 carries the 0*v terms of the SCG columns, so a NaN/Inf SCG sample has poisoned it; has_noise() looks at the
          // RHC channel only (waveform_noise.py:44-49): redo the RHC sum of squares alone, in the order of the common path
          double t2 = 0.0;
#pragma unroll
          for (int k = 0; k < R; ++k) {
            if (tid + k * NT < W) {
              const double dy = __dsub_rn(y[k], K);
              t2 = __fma_rn(dy, dy, t2);
            }
          }
          t2 = warp_sum(t2);
          if (lane == 0) S.zred[warp][0] = t2;
        }
        if (cnt) atomicAdd(&S.slow_cnt, cnt);
        if (fl) atomicOr(&S.slow_flag, fl);
        __syncthreads();
        flat_cnt = S.slow_cnt;
        nonfinite = (S.slow_flag & 1) != 0;
        if (S.slow_flag & 2) { smin = qnan; smax = qnan; }
        if (s2_bad) {
          s2 = S.zred[0][0];
#pragma unroll
          for (int w = 1; w < NWARP; ++w) s2 = __dadd_rn(s2, S.zred[w][0]);
        }
        __syncthreads();
        if (tid == 0) { S.slow_cnt = 0; S.slow_flag = 0; }
      }

      // Syy = sum (y-K)^2 - (sum (y-K))^2 / n ; Sxy is shift invariant because sum (t - xbar) = 0.
      // R^2 > 0.8  <=>  Sxy^2 > 0.8 Sxx Syy (Syy > 0): no division; the two forms can only differ inside
      // the ambiguity band that is flagged below.
      const double syy = __dsub_rn(s2, __dmul_rn(__dmul_rn(s1, s1), inv_w));
      const double lhs = __dmul_rn(sxy, sxy), den = __dmul_rn(sxx, syy);
      if (flat_cnt >= 2) reason |= SCGRHC_REASON_FLAT;
      if (syy > 0.0 && lhs > __dmul_rn(0.8, den)) reason |= SCGRHC_REASON_STRAIGHT;
      if (syy > 0.0 && fabs(__fma_rn(-0.8, den, lhs)) < __dmul_rn(1e-9, den)) reason |= SCGRHC_REASON_AMBIGUOUS;
      // exactly constant window (CTA-uniform, rare): the reference's R^2 is rounding noise — is_straight_line() is True
      // iff np.mean(y) == y[0] exactly.  Only decides windows shorter than 51 samples (longer ones are flat lines).
      if (ymin == ymax && !nonfinite && np_mean_of_const_is_exact(ymin, W)) reason |= SCGRHC_REASON_STRAIGHT;
      if (ymin < min_rhc) reason |= SCGRHC_REASON_FLOOR;  // some sample < floor <=> the minimum is
      if (nonfinite) reason |= SCGRHC_REASON_NONFINITE;
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
        if (reason & SCGRHC_REASON_AMBIGUOUS) atomicAdd(P.err + 2, 1ull);   // surfaced by scgrhc_ambiguous_count
      }
      if (zscore && keep && !pred_only) {
        // extension: joint mean / population std of the SCG block (as the reference takes its min/max jointly) and of
        // the RHC window; sums shifted by a sample of the window, so the cancellation is bounded by the element count.
        // CTA-uniform and outside the unrolled statistics pass: the default mode pays one predicate.
        const double Ks = win[col[0]];
        double t1 = 0.0, t2 = 0.0;
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int t = tid + k * NT;
          const bool valid = WCT > 0 ? ((k + 1) * NT <= WCT || t < WCT) : (t < W);
          if (valid) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const double dv = __dsub_rn(x[k][c], Ks);
              t1 = __dadd_rn(t1, dv);
              t2 = __fma_rn(dv, dv, t2);
            }
          }
        }
        t1 = warp_sum(t1); t2 = warp_sum(t2);
        if (lane == 0) { S.zred[warp][0] = t1; S.zred[warp][1] = t2; }
        __syncthreads();
        t1 = S.zred[0][0]; t2 = S.zred[0][1];
#pragma unroll
        for (int w = 1; w < NWARP; ++w) { t1 = __dadd_rn(t1, S.zred[w][0]); t2 = __dadd_rn(t2, S.zred[w][1]); }
        __syncthreads();                                  // zred is free again before the next window writes it
        const double inv_n = 1.0 / ((double)W * (double)C);
        const double var_s = __dmul_rn(__dsub_rn(t2, __dmul_rn(__dmul_rn(t1, t1), inv_n)), inv_n);
        const double var_y = __dmul_rn(syy, inv_w);
        smin = __fma_rn(t1, inv_n, Ks);                   // from here on the four scalars are mean, std, mean, std
        smax = sqrt(var_s > 0.0 ? var_s : 0.0);
        ymin = __fma_rn(s1, inv_w, K);
        ymax = sqrt(var_y > 0.0 ? var_y : 0.0);
        if (tid == 0) {
          double2* mm = reinterpret_cast<double2*>(P.out.minmax + 4 * M.cand);
          mm[0] = make_double2(smin, smax);
          mm[1] = make_double2(ymin, ymax);
        }
      }
      par2 ^= 1;
    } else {
      __syncthreads();  // all rows are in registers before the stage is refilled
      if (!norm_global) {  // dense re-materialisation with the per-window pairs of an earlier pass
        const double2* mm = reinterpret_cast<const double2*>(P.out.minmax + 4 * M.cand);
        const double2 a = mm[0], b = mm[1];
        smin = a.x; smax = a.y; ymin = b.x; ymax = b.y;
      }
    }

    // ---- the stage buffer is dead: refill it with window n + stages ---------------------------
    if constexpr (!DECIM) {
      if (tid == 0 && lo + n + nstage < hi) issue(lo + n + nstage, s);
    }
    if (++s == nstage) { s = 0; parity ^= 1; }

    // ---- normalise from registers, transpose, cast, store ---------------------------------------
    if (keep && !pred_only) {
      if (norm_global) { smin = J.global_minmax[0]; smax = J.global_minmax[1]; ymin = J.global_minmax[2]; ymax = J.global_minmax[3]; }
      Normaliser ns, nr;
      if (zscore) { ns.init_z(smin, smax); nr.init_z(ymin, ymax); }
      else { ns.init(smin, smax); nr.init(ymin, ymax); }
      OutT* so = reinterpret_cast<OutT*>(P.out.scg_out) + (size_t)M.slot * C * W + tid;
      OutT* ro = reinterpret_cast<OutT*>(P.out.rhc_out) + (size_t)M.slot * W + tid;
      // Tier 1 (fp32 output): q0 = RN(a * RN(1/d)) is within 2.5 ulp64 of the correctly
      // rounded quotient, so both round to the same float unless q0 sits within a few ulp64 of a float
      // rounding boundary (low 29 mantissa bits == 0x10000000).  Track the smallest distance to that
      // pattern with integer ops; a thread that saw a risky element (p ~ 3e-8 per element) falls through
      // to tier 2 and rewrites its elements.  Normaliser::quick keeps every non-zero quotient >= 2^-124, above
      // the float subnormal range where the boundary pattern differs.
      bool redo = true;
      if (zscore && !(ns.slow || nr.slow)) {
        // extension: no bit-exact contract here; RN(a * RN(1/d)) is within 2.5 ulp64 of the quotient (1e-10 / 1e-5 bars)
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int t = tid + k * NT;
          const bool valid = WCT > 0 ? ((k + 1) * NT <= WCT || t < WCT) : (t < W);
          if (valid) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
              OutT o;
              cvt_out(o, __dmul_rn(__dsub_rn(x[k][c], ns.mn), ns.inv));
              st_cs(so + (size_t)c * W + k * NT, o);
            }
            OutT o;
            cvt_out(o, __dmul_rn(__dsub_rn(y[k], nr.mn), nr.inv));
            st_cs(ro + k * NT, o);
          }
        }
        redo = false;
      }
      if constexpr (sizeof(OutT) == 4) {
        if (!zscore && ns.quick && nr.quick) {      // the pairs may be this window's, an earlier pass's or the dataset's: tier 1 only looks at mn and d
          uint32_t acc = 0xffffffffu;
#pragma unroll
          for (int k = 0; k < R; ++k) {
            const int t = tid + k * NT;
            const bool valid = WCT > 0 ? ((k + 1) * NT <= WCT || t < WCT) : (t < W);
            if (valid) {
#pragma unroll
              for (int c = 0; c < C; ++c) {
                const double q0 = __dmul_rn(__dsub_rn(x[k][c], ns.mn), ns.inv);
                acc = min(acc, tier1_key(q0));
                st_cs(so + (size_t)c * W + k * NT, __double2float_rn(q0));
              }
              const double q0 = __dmul_rn(__dsub_rn(y[k], nr.mn), nr.inv);
              acc = min(acc, tier1_key(q0));
              st_cs(ro + k * NT, __double2float_rn(q0));
            }
          }
          redo = acc <= kTier1Risky;
        }
      }
      if (!redo) {
      } else if (!(ns.slow || nr.slow)) {  // tier 2: exact quotient from the reciprocal (Markstein)
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int t = tid + k * NT;
          const bool valid = WCT > 0 ? ((k + 1) * NT <= WCT || t < WCT) : (t < W);
          if (valid) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
              OutT o;
              cvt_out(o, ns.fast(x[k][c]));
              st_cs(so + (size_t)c * W + k * NT, o);
            }
            OutT o;
            cvt_out(o, nr.fast(y[k]));
            st_cs(ro + k * NT, o);
          }
        }
      } else {  // tier 3: IEEE division; kept out of line (not unrolled) so the common path stays small
#pragma unroll 1
        for (int e = 0; e < R * (C + 1); ++e) {
          const int k = e / (C + 1), c = e - k * (C + 1);
          const int t = tid + k * NT;
          if (t < W) {
            double v = 0.0;
#pragma unroll
            for (int kk = 0; kk < R; ++kk) {
              if (kk == k) {
                v = y[kk];
#pragma unroll
                for (int cc = 0; cc < C; ++cc) if (cc == c) v = x[kk][cc];
              }
            }
            OutT o;
            if (c < C) { cvt_out(o, ns.exact(v)); st_cs(so + (size_t)c * W + k * NT, o); }
            else { cvt_out(o, nr.exact(v)); st_cs(ro + k * NT, o); }
          }
        }
      }
    }
  }
}

}  // namespace scgrhc
