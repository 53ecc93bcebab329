// Sweep fan-out (BASELINE configs[4]): the 32 loadable waveform_NN configs are 4 chambers x 8 channel
// subsets, and for one chamber the keep/reject mask depends on the RHC channel only.  So a sweep runs the predicates ONCE
// per chamber (window_kernel, SCGRHC_PREDICATES_ONLY) and then this kernel ONCE over the kept windows: every kept
// window is read once, per-column min/max are reduced once, and the window is normalised and written for every channel
// subset (recordutil.py:55-66 per config: joint min/max over the subset's columns, (x - mn) / (mx - mn + 1e-4) in fp64,
// transpose, cast) plus the RHC window once (it is the same tensor for all subsets).  The reference re-reads and
// re-screens every record per config (waveform_pipeline.py:33-37).
//
// Same staging as the window kernel (1-D bulk async copy of the window's contiguous rows into a 2-stage ring, thread 0
// walks the interval table), same Normaliser (correctly rounded quotient from the reciprocal, IEEE loop for scales that
// could underflow), so every output is bit-identical to what the per-config pass writes.
#pragma once
#include "window_kernel.cuh"

namespace scgrhc {

constexpr int kMaxSubsets = 8;

struct SubsetDesc {
  void* scg_out;     // (n_items, C, W)
  double* minmax;    // (n_items, 4): scg_min, scg_max, rhc_min, rhc_max of this subset
  int mask;          // bit c: superset column c belongs to the subset (channel order = ascending c)
  int C;
};

struct SubsetParams {
  scgrhc_job job;    // superset columns in scg_cols[0..C); kept_list / n_items; intervals; W, stride, nsig
  void* rhc_out;     // (n_items, 1, W)
  int n_sub;
  SubsetDesc sub[kMaxSubsets];
  int stage_elems;
  long long arena_elems_cap;
};

struct SubsetScratch {
  uint64_t full[2];
  StageMeta meta[2];
  double red[NWARP][2 * (SCGRHC_MAX_C + 1)];
  int nanmask[NWARP];
};

// one column of one window: normalise the thread's rows and store them; the IEEE-division loop is kept out of line
// (not unrolled, CTA-uniform branch) so that the common path stays small, as in the window kernel
template <typename OutT, int R>
__device__ __forceinline__ void norm_store(const Normaliser& nz, const double (&v)[R], OutT* base, int tid, int W) {
  if (!nz.slow) {
    // the window kernel's tiers: fp32 outputs first take RN(a * RN(1/d)) with an integer check of the distance to a float
    // rounding boundary; a thread that saw a risky element (p ~ 3e-8) rewrites its elements with the exact quotient
    bool redo = true;
    if constexpr (sizeof(OutT) == 4) {
      if (nz.quick) {
        uint32_t acc = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < R; ++k) {
          if (tid + k * NT < W) {
            const double q0 = __dmul_rn(__dsub_rn(v[k], nz.mn), nz.inv);
            acc = min(acc, tier1_key(q0));
            st_cs(base + k * NT, __double2float_rn(q0));
          }
        }
        redo = acc <= kTier1Risky;
      }
    }
    if (redo) {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        if (tid + k * NT < W) {
          OutT o;
          cvt_out(o, nz.fast(v[k]));
          st_cs(base + k * NT, o);
        }
      }
    }
  } else {
#pragma unroll 1
    for (int k = 0; k < R; ++k) {
      double a = v[0];
#pragma unroll
      for (int kk = 1; kk < R; ++kk) a = kk == k ? v[kk] : a;
      if (tid + k * NT < W) {
        OutT o;
        cvt_out(o, nz.exact(a));
        st_cs(base + k * NT, o);
      }
    }
  }
}

// R: rows per thread (windows of up to R * NT samples)
template <typename OutT, int R>
__global__ void __launch_bounds__(NT, 4) subset_norm_kernel(const __grid_constant__ SubsetParams P) {
  constexpr int CS = SCGRHC_MAX_C;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SubsetScratch& S = *reinterpret_cast<SubsetScratch*>(smem_raw);
  double* stage_base = reinterpret_cast<double*>(smem_raw + ((sizeof(SubsetScratch) + 127) & ~size_t(127)));
  const scgrhc_job& J = P.job;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = J.W, nsig = J.nsig, Csup = J.C, rcol = J.rhc_col;
  const int wstride = J.stride > 0 ? J.stride : W;
  int col[CS];
#pragma unroll
  for (int c = 0; c < CS; ++c) col[c] = c < Csup ? J.scg_cols[c] : 0;

  const long long lo = J.n_items * (long long)blockIdx.x / gridDim.x;
  const long long hi = J.n_items * (long long)(blockIdx.x + 1) / gridDim.x;
  if (lo >= hi) return;
  if (tid == 0) {
    mbar_init(&S.full[0], 1);
    mbar_init(&S.full[1], 1);
    fence_barrier_init();
  }
  __syncthreads();

  // producer (thread 0): kept candidate -> interval -> window rows, as in the window kernel
  int p_iv = 0, p_nwin = 0, p_rec = 0;
  long long p_cand0 = 0, p_row0 = 0;
  auto load_iv = [&](int iv) {
    const scgrhc_interval I = J.intervals[iv];
    p_cand0 = I.cand0; p_row0 = I.row0; p_nwin = I.n_win; p_rec = I.rec_id;
  };
  auto issue = [&](long long item, int s) {
    const long long cand = J.kept_list[item];
    while (cand >= p_cand0 + p_nwin) load_iv(++p_iv);
    const int i = (int)(cand - p_cand0);
    const long long elem0 = (p_row0 + (long long)i * wstride) * nsig;
    const int lead = (int)(elem0 & 1);
    const long long n_even = ((long long)W * nsig + lead + 1) & ~1LL;
    StageMeta m;
    m.cand = cand; m.slot = item; m.elem0 = elem0; m.win = i; m.rec = p_rec; m.lead = lead;
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
    const long long first = J.kept_list[lo];
    int a = 0, b = J.n_intervals - 1;
    while (a < b) {
      const int mid = (a + b + 1) >> 1;
      if (J.intervals[mid].cand0 <= first) a = mid; else b = mid - 1;
    }
    p_iv = a;
    load_iv(a);
    for (int s = 0; s < 2 && lo + s < hi; ++s) issue(lo + s, s);
  }

  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  int s = 0;
  uint32_t parity = 0;
  for (long long n = 0; n < hi - lo; ++n) {
    mbar_wait(&S.full[s], parity);
    const StageMeta M = S.meta[s];
    double* sbuf = stage_base + (size_t)s * P.stage_elems;
    if (M.fallback) {
      const long long ne = (long long)W * nsig;
      for (long long e = tid; e < ne; e += NT) sbuf[M.lead + e] = J.arena[M.elem0 + e];
      __syncthreads();
    }
    const double* win = sbuf + M.lead;

    // rows -> registers; per-column extrema (NaN never wins a comparison; it is tracked per column in a bit mask)
    double x[CS][R], y[R];
    double cmin[CS + 1], cmax[CS + 1];
    unsigned tn = 0u;
#pragma unroll
    for (int c = 0; c <= CS; ++c) { cmin[c] = CUDART_INF; cmax[c] = -CUDART_INF; }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int t = tid + k * NT;
      y[k] = 0.0;
#pragma unroll
      for (int c = 0; c < CS; ++c) x[c][k] = 0.0;
      if (t < W) {
        const double* row = win + (size_t)t * nsig;
#pragma unroll
        for (int c = 0; c < CS; ++c) {
          if (c < Csup) {
            const double v = row[col[c]];
            x[c][k] = v;
            cmin[c] = v < cmin[c] ? v : cmin[c];
            cmax[c] = v > cmax[c] ? v : cmax[c];
            tn |= v != v ? 1u << c : 0u;
          }
        }
        const double v = row[rcol];
        y[k] = v;
        cmin[CS] = v < cmin[CS] ? v : cmin[CS];
        cmax[CS] = v > cmax[CS] ? v : cmax[CS];
        tn |= v != v ? 1u << CS : 0u;
      }
    }
    int nm = (int)__reduce_or_sync(kFull, tn);
#pragma unroll
    for (int c = 0; c <= CS; ++c) {
      cmin[c] = warp_min(cmin[c]);
      cmax[c] = warp_max(cmax[c]);
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c <= CS; ++c) { S.red[warp][2 * c] = cmin[c]; S.red[warp][2 * c + 1] = cmax[c]; }
      S.nanmask[warp] = nm;
    }
    __syncthreads();   // every row is in registers: the stage may be refilled
#pragma unroll
    for (int c = 0; c <= CS; ++c) {
      cmin[c] = S.red[0][2 * c]; cmax[c] = S.red[0][2 * c + 1];
#pragma unroll
      for (int w = 1; w < NWARP; ++w) {
        const double a = S.red[w][2 * c], b = S.red[w][2 * c + 1];
        cmin[c] = a < cmin[c] ? a : cmin[c];
        cmax[c] = b > cmax[c] ? b : cmax[c];
      }
    }
    nm = S.nanmask[0];
#pragma unroll
    for (int w = 1; w < NWARP; ++w) nm |= S.nanmask[w];
    __syncthreads();   // red / nanmask are free for the next window
    if (tid == 0 && lo + n + 2 < hi) issue(lo + n + 2, s);
    if (++s == 2) { s = 0; parity ^= 1; }

    const double ymin = (nm >> CS) & 1 ? qnan : cmin[CS], ymax = (nm >> CS) & 1 ? qnan : cmax[CS];
    // the RHC window: once, whatever the subset
    {
      Normaliser nr;
      nr.init(ymin, ymax);
      OutT* ro = reinterpret_cast<OutT*>(P.rhc_out) + (size_t)M.slot * W + tid;
      norm_store<OutT, R>(nr, y, ro, tid, W);
    }
#pragma unroll 1
    for (int q = 0; q < P.n_sub; ++q) {
      const int mask = P.sub[q].mask;
      double smin = CUDART_INF, smax = -CUDART_INF;
#pragma unroll
      for (int c = 0; c < CS; ++c) {
        if ((mask >> c) & 1) {
          smin = cmin[c] < smin ? cmin[c] : smin;
          smax = cmax[c] > smax ? cmax[c] : smax;
        }
      }
      if (nm & mask) { smin = qnan; smax = qnan; }   // np.min / np.max over a block holding a NaN
      if (tid == 0) {
        double2* mm = reinterpret_cast<double2*>(P.sub[q].minmax + 4 * M.slot);
        mm[0] = make_double2(smin, smax);
        mm[1] = make_double2(ymin, ymax);
      }
      Normaliser ns;
      ns.init(smin, smax);
      OutT* so = reinterpret_cast<OutT*>(P.sub[q].scg_out) + (size_t)M.slot * P.sub[q].C * W + tid;
      int pos = 0;
#pragma unroll
      for (int c = 0; c < CS; ++c) {
        if ((mask >> c) & 1) {
          norm_store<OutT, R>(ns, x[c], so + (size_t)pos * W, tid, W);
          ++pos;
        }
      }
    }
  }
}

}  // namespace scgrhc
