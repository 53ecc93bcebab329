// Small kernels around the hot path: ordered compaction of kept windows, dataset-global min/max,
// batch gather, the standalone rolling-range predicate and the synthetic cohort generator.
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace scgrhc {

// ---- ordered compaction (the order of the list get_segments returns, recordutil.py:148) ---------
constexpr int CB = 256;              // threads per block
constexpr int CITEMS = 8;            // flags per thread
constexpr int CTILE = CB * CITEMS;   // flags per block

__global__ void __launch_bounds__(CB) count_kept_kernel(const uint8_t* __restrict__ keep, long long n,
                                                        int* __restrict__ block_counts) {
  const long long base = (long long)blockIdx.x * CTILE;
  int c = 0;
#pragma unroll
  for (int i = 0; i < CITEMS; ++i) {
    const long long j = base + (long long)i * CB + threadIdx.x;
    if (j < n) c += keep[j] ? 1 : 0;
  }
  __shared__ int ws[CB / 32];
  for (int m = 16; m; m >>= 1) c += __shfl_xor_sync(kFull, c, m);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < CB / 32; ++w) t += ws[w];
    block_counts[blockIdx.x] = t;
  }
}

// single block: exclusive scan of block_counts -> block_offsets, total -> n_kept
__global__ void __launch_bounds__(1024) scan_blocks_kernel(const int* __restrict__ block_counts, int nblocks,
                                                           long long* __restrict__ block_offsets,
                                                           long long* __restrict__ n_kept) {
  __shared__ long long ws[32];
  __shared__ long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nblocks; base += 1024) {
    const int j = base + threadIdx.x;
    const long long v = j < nblocks ? block_counts[j] : 0;
    long long incl = v;
    for (int d = 1; d < 32; d <<= 1) {
      const long long o = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      long long w = ws[lane];
      for (int d = 1; d < 32; d <<= 1) {
        const long long o = __shfl_up_sync(kFull, w, d);
        if (lane >= d) w += o;
      }
      ws[lane] = w;
    }
    __syncthreads();
    const long long carry = carry_s;
    const long long warp_off = warp ? ws[warp - 1] : 0;
    if (j < nblocks) block_offsets[j] = carry + warp_off + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_off + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_kept = carry_s;
}

__global__ void __launch_bounds__(CB) scatter_kept_kernel(const uint8_t* __restrict__ keep,
                                                          const int32_t* __restrict__ cand_win,
                                                          const int32_t* __restrict__ cand_rec, long long n, int W, int stride,
                                                          const long long* __restrict__ block_offsets,
                                                          scgrhc_compact out) {
  // thread t owns CITEMS consecutive flags so that ranks follow candidate order
  const long long base = (long long)blockIdx.x * CTILE + (long long)threadIdx.x * CITEMS;
  int flags[CITEMS], c = 0;
#pragma unroll
  for (int i = 0; i < CITEMS; ++i) {
    const long long j = base + i;
    flags[i] = (j < n && keep[j]) ? 1 : 0;
    c += flags[i];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = c;
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(kFull, incl, d);
    if (lane >= d) incl += o;
  }
  __shared__ int ws[CB / 32];
  if (lane == 31) ws[warp] = incl;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += ws[w];
  long long r = block_offsets[blockIdx.x] + woff + incl - c;
#pragma unroll
  for (int i = 0; i < CITEMS; ++i) {
    if (flags[i]) {
      const long long j = base + i;
      out.kept_idx[r] = j;
      if (out.start_idx) {
        const long long st = (long long)cand_win[j] * stride;
        out.start_idx[r] = st;
        out.stop_idx[r] = st + W;
      }
      if (out.rec_id) out.rec_id[r] = cand_rec[j];
      ++r;
    }
  }
}

// ---- get_global_minmax_vals (recordutil.py:152-169): min/max over kept windows' pairs -----------
constexpr int GMM_BLOCKS = 296;
__global__ void __launch_bounds__(256) minmax_partial_kernel(const double* __restrict__ minmax,
                                                             const uint8_t* __restrict__ keep, long long n,
                                                             double* __restrict__ partial /* (grid,4) */) {
  double v0 = CUDART_INF, v1 = -CUDART_INF, v2 = CUDART_INF, v3 = -CUDART_INF;
  bool nan_s = false;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    if (keep[j]) {
      const double2 a = reinterpret_cast<const double2*>(minmax + 4 * j)[0];
      const double2 b = reinterpret_cast<const double2*>(minmax + 4 * j)[1];
      nan_s |= (a.x != a.x);
      v0 = fmin(v0, a.x); v1 = fmax(v1, a.y); v2 = fmin(v2, b.x); v3 = fmax(v3, b.y);
    }
  }
  v0 = warp_min(v0); v1 = warp_max(v1); v2 = warp_min(v2); v3 = warp_max(v3);
  const int any_nan = __syncthreads_or(nan_s ? 1 : 0);
  __shared__ double ws[8][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { ws[warp][0] = v0; ws[warp][1] = v1; ws[warp][2] = v2; ws[warp][3] = v3; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      v0 = fmin(v0, ws[w][0]); v1 = fmax(v1, ws[w][1]); v2 = fmin(v2, ws[w][2]); v3 = fmax(v3, ws[w][3]);
    }
    if (any_nan) v0 = v1 = __longlong_as_double(0x7ff8000000000000LL);
    partial[4 * blockIdx.x + 0] = v0; partial[4 * blockIdx.x + 1] = v1;
    partial[4 * blockIdx.x + 2] = v2; partial[4 * blockIdx.x + 3] = v3;
  }
}
__global__ void minmax_final_kernel(const double* __restrict__ partial, int nparts, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double v0 = CUDART_INF, v1 = -CUDART_INF, v2 = CUDART_INF, v3 = -CUDART_INF;
  bool nan_s = false;
  for (int p = 0; p < nparts; ++p) {
    const double a = partial[4 * p];
    nan_s |= (a != a);
    v0 = fmin(v0, a); v1 = fmax(v1, partial[4 * p + 1]);
    v2 = fmin(v2, partial[4 * p + 2]); v3 = fmax(v3, partial[4 * p + 3]);
  }
  if (nan_s) v0 = v1 = __longlong_as_double(0x7ff8000000000000LL);
  out[0] = v0; out[1] = v1; out[2] = v2; out[3] = v3;
}

// ---- batch collate: out[b] = store[slot[b]]; WordT = 8-byte words when window_bytes allows, else 4-byte -------
template <typename WordT>
__global__ void __launch_bounds__(256) gather_windows_kernel(const unsigned char* __restrict__ store,
                                                             const long long* __restrict__ slots, long long n,
                                                             long long window_bytes, unsigned char* __restrict__ out) {
  const long long words = window_bytes / (long long)sizeof(WordT);
  for (long long b = blockIdx.x; b < n; b += gridDim.x) {
    const WordT* src = reinterpret_cast<const WordT*>(store + slots[b] * window_bytes);
    WordT* dst = reinterpret_cast<WordT*>(out + b * window_bytes);
    for (long long i = threadIdx.x; i < words; i += blockDim.x) dst[i] = __ldcs(src + i);
  }
}

// ---- record ingest: WFDB format-16 digital frames -> physical fp64 samples on the device ------------------
// wfdb's dac(): p = (d - baseline) / gain in fp64, the invalid-sample code -32768 -> NaN.  Only the selected
// columns are produced, so the arena the window kernel reads is (rows, ncols) with the identity column map.
struct DecodeParams {
  const short* d;          // (T, nsig_in) interleaved frames
  double* out;             // (T, ncols)
  long long T;
  long long plane;         // 0: out is interleaved (T, ncols); P > 0: column j of row t at out[j * P + t]
  int nsig_in, ncols;
  int cols[SCGRHC_MAX_C + 1];
  double gain[SCGRHC_MAX_C + 1], baseline[SCGRHC_MAX_C + 1];
};
__global__ void __launch_bounds__(256) decode_fmt16_kernel(const __grid_constant__ DecodeParams P) {
  const long long total = P.T * P.ncols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long t = e / P.ncols;
    const int j = (int)(e - t * P.ncols);
    const short d = P.d[t * P.nsig_in + P.cols[j]];
    const double v = __ddiv_rn(__dsub_rn((double)d, P.baseline[j]), P.gain[j]);
    P.out[P.plane ? j * P.plane + t : e] = d == -32768 ? __longlong_as_double(0x7ff8000000000000LL) : v;
  }
}

// One frame per thread: the frame's samples come in with one 8-byte load when it is 4 x int16 (else per column), every
// output row goes out with 16-byte stores, and the IEEE quotient comes from the column's correctly rounded reciprocal
// with two FMA residual corrections (div_by_recip, validated against __ddiv_rn in scgrhc_selftest_div) instead of a full
// division per sample.  The host selects it when every gain is a normal number in [2^-40, 2^60] (then no intermediate
// can underflow: |d - baseline| is 0 or >= 2^-36 for |baseline| < 2^16).
template <int NC>
__global__ void __launch_bounds__(256) decode_fmt16_rows_kernel(const __grid_constant__ DecodeParams P) {
  double inv[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) inv[j] = __drcp_rn(P.gain[j]);
  const bool frame4 = P.nsig_in == 4 && (reinterpret_cast<uintptr_t>(P.d) & 7) == 0;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < P.T; t += (long long)gridDim.x * blockDim.x) {
    short dv[NC];
    if (frame4) {
      const short4 f = __ldcs(reinterpret_cast<const short4*>(P.d) + t);
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int c = P.cols[j];
        dv[j] = c == 0 ? f.x : (c == 1 ? f.y : (c == 2 ? f.z : f.w));
      }
    } else {
      const short* row = P.d + t * P.nsig_in;
#pragma unroll
      for (int j = 0; j < NC; ++j) dv[j] = row[P.cols[j]];
    }
    double v[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const double q = div_by_recip(__dsub_rn((double)dv[j], P.baseline[j]), P.gain[j], inv[j]);
      v[j] = dv[j] == -32768 ? qnan : q;
    }
    if (P.plane) {                     // planar arena: consecutive threads write consecutive rows of each plane
#pragma unroll
      for (int j = 0; j < NC; ++j) __stcs(P.out + j * P.plane + t, v[j]);
      continue;
    }
    double* o = P.out + t * NC;
    if constexpr (NC % 2 == 0) {
#pragma unroll
      for (int j = 0; j < NC; j += 2) __stcs(reinterpret_cast<double2*>(o + j), make_double2(v[j], v[j + 1]));
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j) __stcs(o + j, v[j]);
    }
  }
}

// The same for a CHUNK of records with per-record calibration (every WFDB header carries its own gain/baseline): one
// launch per chunk instead of one per record.  blockIdx.y = record of the chunk; rec_row0[r] .. rec_row0[r+1] = its
// frames inside d / out; gain / baseline are (n_rec, NC) device tables.  `recip` != 0: the host has checked that every
// gain of the table satisfies the conditions of the reciprocal path above; else IEEE division.
struct DecodeRecParams {
  const short* d;            // (T_chunk, nsig_in) interleaved frames, records back to back
  double* out;               // (T_chunk, ncols)
  const long long* rec_row0; // device (n_rec + 1)
  const double* gain;        // device (n_rec, ncols)
  const double* baseline;    // device (n_rec, ncols)
  long long plane;           // 0: interleaved output; P > 0: planar, column j of row t at out[j * P + t]
  int n_rec, nsig_in, ncols, recip;
  int cols[SCGRHC_MAX_C + 1];
};
template <int NC>
__global__ void __launch_bounds__(256) decode_fmt16_records_kernel(const __grid_constant__ DecodeRecParams P) {
  const int r = blockIdx.y;
  const long long t0 = P.rec_row0[r], t1 = P.rec_row0[r + 1];
  double g[NC], b[NC], inv[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    g[j] = P.gain[(size_t)r * NC + j];
    b[j] = P.baseline[(size_t)r * NC + j];
    inv[j] = __drcp_rn(g[j]);
  }
  const bool frame4 = P.nsig_in == 4 && (reinterpret_cast<uintptr_t>(P.d) & 7) == 0;
  const bool out16 = (reinterpret_cast<uintptr_t>(P.out) & 15) == 0;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  for (long long t = t0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; t < t1; t += (long long)gridDim.x * blockDim.x) {
    short dv[NC];
    if (frame4) {
      const short4 f = __ldcs(reinterpret_cast<const short4*>(P.d) + t);
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int c = P.cols[j];
        dv[j] = c == 0 ? f.x : (c == 1 ? f.y : (c == 2 ? f.z : f.w));
      }
    } else {
      const short* row = P.d + t * P.nsig_in;
#pragma unroll
      for (int j = 0; j < NC; ++j) dv[j] = row[P.cols[j]];
    }
    double v[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const double a = __dsub_rn((double)dv[j], b[j]);
      const double q = P.recip ? div_by_recip(a, g[j], inv[j]) : __ddiv_rn(a, g[j]);
      v[j] = dv[j] == -32768 ? qnan : q;
    }
    if (P.plane) {
#pragma unroll
      for (int j = 0; j < NC; ++j) __stcs(P.out + j * P.plane + t, v[j]);
      continue;
    }
    double* o = P.out + t * NC;
    if (NC % 2 == 0 && out16) {
#pragma unroll
      for (int j = 0; j + 1 < NC; j += 2) __stcs(reinterpret_cast<double2*>(o + j), make_double2(v[j], v[j + 1]));
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j) __stcs(o + j, v[j]);
    }
  }
}

// ---- extension (north star, absent from the reference): train-time noise injection fused into the batch gather.
// Counter-based Philox4x32-10 (Salmon et al., Random123; the cuRAND-style 4x32 variant, NOT numpy's 4x64):
// key = seed, counter = (block index of the element quad, stream offset).  Element j of the batch takes word j&3
// of block j>>2; words (0,1) and (2,3) feed one Box-Muller pair each.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long offset, unsigned long long q) {
  const uint4 u = philox4x32_10(make_uint4((unsigned)q, (unsigned)(q >> 32), (unsigned)offset, (unsigned)(offset >> 32)),
                                make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  const float k = 2.3283064365386963e-10f;  // 2^-32
  const float u0 = ((float)u.x + 0.5f) * k, u1 = ((float)u.y + 0.5f) * k, u2 = ((float)u.z + 0.5f) * k, u3 = ((float)u.w + 0.5f) * k;
  const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincosf(6.2831853071795865f * u1, &s0, &c0);
  sincosf(6.2831853071795865f * u3, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}
__global__ void __launch_bounds__(256) gather_noise_kernel(const float* __restrict__ store, const long long* __restrict__ slots,
                                                           long long n, long long E, float* __restrict__ out, float sigma,
                                                           unsigned long long seed, unsigned long long offset) {
  // E (elements per window) is a multiple of 2, so a quad never straddles more than two windows; handle per element
  const long long quads = (n * E + 3) >> 2;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
    const float4 z = philox_normal4(seed, offset, (unsigned long long)q);
    const float zs[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const long long j = 4 * q + r;
      if (j < n * E) {
        const long long b = j / E, e = j - b * E;
        out[j] = __fmaf_rn(sigma, zs[r], __ldcs(store + slots[b] * E + e));
      }
    }
  }
}
// ---- one-launch batch collate (default_collate of recordutil.py:198 for the two tensors the trainer reads,
//      waveform_train.py:358-359): scg_out[b] = scg_store[slot[b]] (+ sigma * N(0,1), the noise extension, element j of
//      the SCG batch drawing word j & 3 of Philox block j >> 2 exactly as gather_noise_kernel does) and
//      rhc_out[b] = rhc_store[slot[b]], fp32.  blockIdx.x = batch item; blockIdx.y splits the window when the batch is
//      too small to fill the GPU.
struct CollateParams {
  const float* scg_store; const float* rhc_store;
  const long long* slots;
  float* scg_out; float* rhc_out;
  long long n;
  int scg_elems, rhc_elems;    // floats per window: C*W and W
  float sigma;
  unsigned long long seed, offset;
};
template <bool NOISE>
__global__ void __launch_bounds__(256) collate_batch_kernel(const __grid_constant__ CollateParams P) {
  const long long b = blockIdx.x;
  const long long slot = P.slots[b];
  const int part = blockIdx.y, nparts = gridDim.y;
  const float* ssrc = P.scg_store + slot * P.scg_elems;
  const float* rsrc = P.rhc_store + slot * P.rhc_elems;
  float* sdst = P.scg_out + b * P.scg_elems;
  float* rdst = P.rhc_out + b * P.rhc_elems;
  const int tid = part * blockDim.x + threadIdx.x, nthr = nparts * blockDim.x;
  // RHC window (and the SCG window without noise): 8-byte words when the window offsets allow, else 4-byte
  const bool r8 = (P.rhc_elems & 1) == 0 && ((reinterpret_cast<uintptr_t>(P.rhc_store) | reinterpret_cast<uintptr_t>(P.rhc_out)) & 7) == 0;
  if (r8) {
    for (int i = tid; i < (P.rhc_elems >> 1); i += nthr) reinterpret_cast<float2*>(rdst)[i] = __ldcs(reinterpret_cast<const float2*>(rsrc) + i);
  } else {
    for (int i = tid; i < P.rhc_elems; i += nthr) rdst[i] = __ldcs(rsrc + i);
  }
  if constexpr (!NOISE) {
    const bool s8 = (P.scg_elems & 1) == 0 && ((reinterpret_cast<uintptr_t>(P.scg_store) | reinterpret_cast<uintptr_t>(P.scg_out)) & 7) == 0;
    if (s8) {
      for (int i = tid; i < (P.scg_elems >> 1); i += nthr) reinterpret_cast<float2*>(sdst)[i] = __ldcs(reinterpret_cast<const float2*>(ssrc) + i);
    } else {
      for (int i = tid; i < P.scg_elems; i += nthr) sdst[i] = __ldcs(ssrc + i);
    }
  } else {
    // Philox blocks are aligned to the element index of the whole batch: walk the quads that overlap this window (the
    // two at its ends are shared with the neighbours and computed by both)
    const long long j0 = b * P.scg_elems, j1 = j0 + P.scg_elems;
    for (long long q = (j0 >> 2) + tid; q < ((j1 + 3) >> 2); q += nthr) {
      const float4 z = philox_normal4(P.seed, P.offset, (unsigned long long)q);
      const float zs[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const long long j = 4 * q + r;
        if (j >= j0 && j < j1) sdst[j - j0] = __fmaf_rn(P.sigma, zs[r], __ldcs(ssrc + (j - j0)));
      }
    }
  }
}

__global__ void philox_words_kernel(unsigned long long seed, unsigned long long offset, long long nquads, uint4* out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nquads)
    out[q] = philox4x32_10(make_uint4((unsigned)q, (unsigned)(q >> 32), (unsigned)offset, (unsigned)(offset >> 32)),
                           make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
}

// ---- evaluation metrics of waveform_test.py:21-50,66-70 on the device: de-normalise both waveforms with the window's
//      RHC pair (reverse_minmax: v*(max-min)+min, no epsilon), Pearson r (centred, as scipy.stats.pearsonr) and RMSE.
//      One warp per window; out[w] = {pcc_r, rmse}.
__global__ void __launch_bounds__(256) window_metrics_kernel(const float* __restrict__ real, const float* __restrict__ pred,
                                                             const double* __restrict__ minmax, long long n, int W,
                                                             double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long w = warp0; w < n; w += nwarps) {
    const double mn = minmax[2 * w], range = __dsub_rn(minmax[2 * w + 1], mn);
    const float* a = real + w * W;
    const float* b = pred + w * W;
    double sx = 0, sy = 0, sd = 0;
    for (int i = lane; i < W; i += 32) {
      const double x = __dadd_rn(__dmul_rn((double)a[i], range), mn), y = __dadd_rn(__dmul_rn((double)b[i], range), mn);
      sx = __dadd_rn(sx, x); sy = __dadd_rn(sy, y);
      const double d = __dsub_rn(x, y);
      sd = __fma_rn(d, d, sd);
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sd = warp_sum(sd);
    const double mx = sx / W, my = sy / W;
    double sxx = 0, syy = 0, sxy = 0;
    for (int i = lane; i < W; i += 32) {
      const double x = __dsub_rn(__dadd_rn(__dmul_rn((double)a[i], range), mn), mx);
      const double y = __dsub_rn(__dadd_rn(__dmul_rn((double)b[i], range), mn), my);
      sxx = __fma_rn(x, x, sxx); syy = __fma_rn(y, y, syy); sxy = __fma_rn(x, y, sxy);
    }
    sxx = warp_sum(sxx); syy = warp_sum(syy); sxy = warp_sum(sxy);
    if (lane == 0) {
      double r = sxy / (sqrt(sxx) * sqrt(syy));
      r = r > 1.0 ? 1.0 : (r < -1.0 ? -1.0 : r);
      out[2 * w] = r;
      out[2 * w + 1] = sqrt(sd / W);
    }
  }
}

// ---- standalone rolling range (API parity of get_flat_lines with non-default arguments) ------------
__global__ void rolling_range_lt_kernel(const double* __restrict__ y, long long n, int m, double thr,
                                        uint8_t* __restrict__ flags) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint8_t f = 0;
  if (p >= m - 1) {
    double mx = -CUDART_INF, mn = CUDART_INF;
    bool bad = false;
    for (int i = 0; i < m; ++i) {
      const double v = y[p - i];
      bad |= (v != v);
      mx = fmax(mx, v); mn = fmin(mn, v);
    }
    f = (!bad && (__dsub_rn(mx, mn) < thr)) ? 1 : 0;
  }
  flags[p] = f;
}

// ---- per-waveform statistics of arbitrary length (API parity of is_straight_line / in_rhc_range on
//      waveforms that are not 2..1024 samples long; waveform_noise.py:29-41).  One block per waveform.
//      stats[w] = {R^2, min, max, below_floor, nonfinite, sum}
__global__ void __launch_bounds__(256) waveform_stats_kernel(const double* __restrict__ y, long long n_wave, long long L,
                                                             double min_rhc, double* __restrict__ stats) {
  __shared__ double red[8][4];
  __shared__ int flags_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long w = blockIdx.x; w < n_wave; w += gridDim.x) {
    const double* v = y + w * L;
    double mn = CUDART_INF, mx = -CUDART_INF, sum = 0.0;
    int f = 0;
    if (threadIdx.x == 0) flags_s = 0;
    __syncthreads();
    for (long long i = threadIdx.x; i < L; i += blockDim.x) {
      const double a = v[i];
      mn = fmin(mn, a); mx = fmax(mx, a); sum = __dadd_rn(sum, a);
      if (a < min_rhc) f |= 1;
      if (!(fabs(a) <= 1.7976931348623157e308)) f |= 2;
    }
    mn = warp_min(mn); mx = warp_max(mx); sum = warp_sum(sum);
    f = __reduce_or_sync(kFull, f);
    if (lane == 0) { red[warp][0] = mn; red[warp][1] = mx; red[warp][2] = sum; if (f) atomicOr(&flags_s, f); }
    __syncthreads();
    mn = red[0][0]; mx = red[0][1]; sum = red[0][2];
    for (int k = 1; k < 8; ++k) { mn = fmin(mn, red[k][0]); mx = fmax(mx, red[k][1]); sum = __dadd_rn(sum, red[k][2]); }
    const int fl = flags_s;
    __syncthreads();
    const double ybar = __ddiv_rn(sum, (double)L), xbar = 0.5 * (double)(L - 1);
    double sxy = 0.0, syy = 0.0;
    for (long long i = threadIdx.x; i < L; i += blockDim.x) {
      const double dy = __dsub_rn(v[i], ybar);
      sxy = __fma_rn((double)i - xbar, dy, sxy);
      syy = __fma_rn(dy, dy, syy);
    }
    sxy = warp_sum(sxy); syy = warp_sum(syy);
    if (lane == 0) { red[warp][0] = sxy; red[warp][1] = syy; }
    __syncthreads();
    if (threadIdx.x == 0) {
      sxy = red[0][0]; syy = red[0][1];
      for (int k = 1; k < 8; ++k) { sxy = __dadd_rn(sxy, red[k][0]); syy = __dadd_rn(syy, red[k][1]); }
      const double Ld = (double)L;
      const double sxx = Ld * (Ld * Ld - 1.0) / 12.0;
      double* o = stats + 6 * w;
      o[0] = __ddiv_rn(__dmul_rn(sxy, sxy), __dmul_rn(sxx, syy));
      o[1] = mn; o[2] = mx; o[3] = (fl & 1) ? 1.0 : 0.0; o[4] = (fl & 2) ? 1.0 : 0.0; o[5] = sum;
    }
    __syncthreads();
  }
}

// ---- Markstein division self-test --------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  unsigned long long z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// ---- synthetic cohort generator; bit-identical twin of oracle/synth_ref.py ------------------------
struct SynthParams {
  unsigned long long seed;
  long long rec0, n_rec, T;
  int nsig, defect_scale, grid;
  int kinds[SCGRHC_MAX_NSIG];
  double* out;
  long long plane;           // 0: interleaved (rows, nsig); P > 0: planar, column ch of row r at out[ch * P + r]
};

__device__ __forceinline__ unsigned long long syn_h(unsigned long long key, unsigned stream, unsigned long long idx) {
  return mix64(key ^ ((unsigned long long)stream << 48) ^ idx);
}
__device__ __forceinline__ double syn_u01(unsigned long long h) { return (double)(h >> 11) * 0x1p-53; }
__device__ __forceinline__ double syn_noise(unsigned long long key, int kind, long long t) {
  const unsigned long long h = syn_h(key, 16 + kind, (unsigned long long)t);
  const long long s = (long long)((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + (h >> 48));
  return __dmul_rn((double)(s - 131070), 2.6429e-05);
}
__device__ __forceinline__ double syn_shape(unsigned long long phase) {
  const double p = (double)phase * 0x1p-32;
  const double u = __dadd_rn(__dmul_rn(2.0, p), -1.0);
  return __dmul_rn(4.0, __dmul_rn(u, __dadd_rn(1.0, -fabs(u))));
}
__device__ __forceinline__ unsigned long long syn_phase(unsigned long long ph0, long long t, unsigned long long step) {
  return (ph0 + (unsigned long long)t * step) & 0xFFFFFFFFull;
}
__device__ __forceinline__ double syn_rhc_base(unsigned long long key, unsigned long long finc, long long t) {
  const unsigned long long ph1 = syn_phase(syn_h(key, 0, 1 + 3) & 0xFFFFFFFFull, t, finc);
  const unsigned long long ph2 = (ph1 * 2ull + 0x14000000ull) & 0xFFFFFFFFull;
  const double s1 = syn_shape(ph1), s2 = syn_shape(ph2), n = syn_noise(key, 3, t);
  return __dadd_rn(__dadd_rn(__dadd_rn(25.0, __dmul_rn(12.0, s1)), __dmul_rn(3.0, s2)), __dmul_rn(0.3, n));
}

__device__ double syn_channel(unsigned long long key, unsigned long long finc, int kind, long long t, int defect_scale,
                              int grid) {
  if (kind >= 0 && kind <= 2) {
    const unsigned long long mult = kind == 0 ? 9 : (kind == 1 ? 13 : 17);
    const unsigned long long ph = syn_phase(syn_h(key, 0, 1 + kind) & 0xFFFFFFFFull, t, finc * mult);
    return __dadd_rn(__dmul_rn(0.02, syn_shape(ph)), __dmul_rn(0.005, syn_noise(key, kind, t)));
  }
  if (kind == 4) {
    const unsigned long long ph = syn_phase(syn_h(key, 0, 1 + 4) & 0xFFFFFFFFull, t, finc);
    const double s = syn_shape(ph);
    return __dadd_rn(__dmul_rn(0.8, __dmul_rn(s, fabs(s))), __dmul_rn(0.02, syn_noise(key, 4, t)));
  }
  if (kind != 3) return __dmul_rn(0.1, syn_noise(key, kind, t));
  double y = syn_rhc_base(key, finc, t);
  if (defect_scale <= 0) return y;
  const long long j = t / grid;
  const long long tp = t - j * grid;
  const unsigned long long d = syn_h(key, 2, (unsigned long long)j);
  const int r = (int)(d % 10000ull);
  const int c0 = 500 * defect_scale / 16, c1 = 1000 * defect_scale / 16, c2 = 1500 * defect_scale / 16,
            c3 = 2000 * defect_scale / 16, c4 = 2300 * defect_scale / 16, c5 = 2600 * defect_scale / 16;
  if (r >= c5) return y;
  const int sel = (int)((d >> 8) & 3ull);
  const long long a = 100 + (long long)((d >> 16) % 400ull);
  const double tpf = (double)tp;
  if (r < c0) {
    const int L = sel == 0 ? 49 : (sel == 1 ? 50 : (sel == 2 ? 51 : 200));
    if (tp >= a && tp < a + L) y = syn_rhc_base(key, finc, j * grid + a);
  } else if (r < c1) {
    y = __dadd_rn(__dadd_rn(10.0, __dmul_rn(0.04, tpf)), __dmul_rn(0.05, syn_noise(key, 3, t)));
  } else if (r < c2) {
    if (tp >= 300 && tp < 310) y = -60.0;
  } else if (r < c3) {
    if (tp == 375) y = -50.0;
  } else if (r < c4) {
    const double A = __dadd_rn(1.5, __dmul_rn(1.5, syn_u01(syn_h(key, 3, (unsigned long long)j))));
    y = __dadd_rn(__dadd_rn(5.0, __dmul_rn(0.02, tpf)), __dmul_rn(A, syn_noise(key, 3, t)));
  } else {
    if (tp >= a && tp < a + 120) {
      const double amp = __dadd_rn(0.6e-3, __dmul_rn(0.8e-3, syn_u01(syn_h(key, 3, (unsigned long long)j))));
      y = __dadd_rn(syn_rhc_base(key, finc, j * grid + a), __dmul_rn(amp, syn_u01(syn_h(key, 4, (unsigned long long)t))));
    }
  }
  return y;
}

__global__ void __launch_bounds__(256) synth_kernel(const __grid_constant__ SynthParams P) {
  const long long total = P.n_rec * P.T * P.nsig;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(e % P.nsig);
    const long long row = e / P.nsig;
    const long long rec = row / P.T;
    const long long t = row - rec * P.T;
    const unsigned long long key = mix64(mix64(P.seed) ^ ((unsigned long long)(P.rec0 + rec) * 0xD1342543DE82EF95ull));
    const unsigned long long finc = 7730941ull + syn_h(key, 0, 0) % 7730942ull;
    P.out[P.plane ? ch * P.plane + row : e] = syn_channel(key, finc, P.kinds[ch], t, P.defect_scale, P.grid);
  }
}

}  // namespace scgrhc
