/*
 * TEST / BASELINE INFRASTRUCTURE — plain-C restatement of the reference's per-window arithmetic.
 *
 * Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may
 * load this (oracle/_build/liboracle.so).  It is pinned against the fixtures the unmodified reference
 * produced (tests/golden, via tests/test_oracle.py) and is the *fast* CPU baseline: one thread per
 * window with OpenMP, O(W) rolling min/max — i.e. the best a host implementation of the same
 * algorithm does, as opposed to oracle/ref_port.py which keeps the reference's pandas/sklearn cost.
 *
 * Restates (paths relative to the reference):
 *   waveform_noise.py:6-26   get_flat_lines: rolling(50) max-min < 1e-3 at >= 2 positions
 *   waveform_noise.py:29-34  is_straight_line: OLS R^2 > 0.8   (closed form Sxy^2/(Sxx Syy), long double)
 *   waveform_noise.py:37-41  in_rhc_range: any sample < min_RHC
 *   recordutil.py:41-66      min/max (SCG joint over channels), (x-mn)/(mx-mn+1e-4), transpose, fp32
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FLAT_WIN 50

/* number of positions p >= FLAT_WIN-1 with fl(max - min over y[p-49..p]) < thr; NaN never compares */
static int flat_count(const double* y, int W, double thr) {
  if (W < FLAT_WIN) return 0;
  int cnt = 0;
  /* van Herk / Gil-Werman: prefix and suffix extrema per block of FLAT_WIN samples */
  double pmax[4096], pmin[4096], smax[4096], smin[4096];
  unsigned char bad_p[4096], bad_s[4096];
  if (W > 4096) return -1;
  for (int b = 0; b < W; b += FLAT_WIN) {
    int e = b + FLAT_WIN < W ? b + FLAT_WIN : W;
    double mx = -INFINITY, mn = INFINITY; unsigned char bad = 0;
    for (int i = b; i < e; ++i) {
      if (y[i] != y[i]) bad = 1; else { if (y[i] > mx) mx = y[i]; if (y[i] < mn) mn = y[i]; }
      pmax[i] = mx; pmin[i] = mn; bad_p[i] = bad;
    }
    mx = -INFINITY; mn = INFINITY; bad = 0;
    for (int i = e - 1; i >= b; --i) {
      if (y[i] != y[i]) bad = 1; else { if (y[i] > mx) mx = y[i]; if (y[i] < mn) mn = y[i]; }
      smax[i] = mx; smin[i] = mn; bad_s[i] = bad;
    }
  }
  for (int s = 0; s + FLAT_WIN <= W; ++s) {
    int e = s + FLAT_WIN - 1;
    double mx, mn; unsigned char bad;
    if (s % FLAT_WIN == 0) { mx = smax[s]; mn = smin[s]; bad = bad_s[s]; }
    else {
      mx = smax[s] > pmax[e] ? smax[s] : pmax[e];
      mn = smin[s] < pmin[e] ? smin[s] : pmin[e];
      bad = bad_s[s] | bad_p[e];
    }
    if (!bad && (mx - mn) < thr) ++cnt;
  }
  return cnt;
}

static double r_squared(const double* y, int W, int* is_const) {
  long double sum = 0;
  double mn = y[0], mx = y[0];
  for (int i = 0; i < W; ++i) { sum += y[i]; if (y[i] < mn) mn = y[i]; if (y[i] > mx) mx = y[i]; }
  *is_const = (mn == mx);
  const long double ybar = sum / W, xbar = (long double)(W - 1) / 2;
  long double sxy = 0, syy = 0, sxx = 0;
  for (int i = 0; i < W; ++i) {
    const long double dx = i - xbar, dy = y[i] - ybar;
    sxy += dx * dy; syy += dy * dy; sxx += dx * dx;
  }
  return (double)(sxy * sxy / (sxx * syy));
}

/* reason bits as in include/scgrhc.h */
int oracle_process_windows(const double* arena, int nsig, int W, int C, const int* cols, int rcol,
                           int64_t n_win, const int64_t* row_start, double min_rhc, double thr,
                           uint8_t* keep, uint8_t* reason, double* minmax, float* scg_out, float* rhc_out,
                           int write_all_slots) {
  int err = 0;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t w = 0; w < n_win; ++w) {
    const double* base = arena + row_start[w] * nsig;
    double y[4096];
    if (W > 4096) { err = 1; continue; }
    int below = 0, nonfin = 0;
    double rmin = INFINITY, rmax = -INFINITY;
    for (int t = 0; t < W; ++t) {
      const double v = base[(int64_t)t * nsig + rcol];
      y[t] = v;
      if (v < min_rhc) below = 1;
      if (!(fabs(v) <= 1.7976931348623157e308)) nonfin = 1;
      if (v < rmin) rmin = v;
      if (v > rmax) rmax = v;
    }
    unsigned r = 0;
    if (flat_count(y, W, thr) >= 2) r |= 1u;
    int is_const = 0;
    const double r2 = r_squared(y, W, &is_const);
    if (!is_const && r2 > 0.8) r |= 2u;
    if (below) r |= 4u;
    if (nonfin) r |= 8u;
    double smin = INFINITY, smax = -INFINITY; int snan = 0;
    for (int t = 0; t < W; ++t)
      for (int c = 0; c < C; ++c) {
        const double v = base[(int64_t)t * nsig + cols[c]];
        if (v != v) snan = 1;
        if (v < smin) smin = v;
        if (v > smax) smax = v;
      }
    if (snan) smin = smax = NAN;
    keep[w] = r == 0;
    reason[w] = (uint8_t)r;
    minmax[4 * w] = smin; minmax[4 * w + 1] = smax; minmax[4 * w + 2] = rmin; minmax[4 * w + 3] = rmax;
    if (scg_out && (keep[w] || write_all_slots)) {
      const double ds = smax - smin + 0.0001, dr = rmax - rmin + 0.0001;
      float* so = scg_out + (int64_t)w * C * W;
      float* ro = rhc_out + (int64_t)w * W;
      for (int t = 0; t < W; ++t) {
        for (int c = 0; c < C; ++c) so[(int64_t)c * W + t] = (float)((base[(int64_t)t * nsig + cols[c]] - smin) / ds);
        ro[t] = (float)((y[t] - rmin) / dr);
      }
    }
  }
  return err;
}
