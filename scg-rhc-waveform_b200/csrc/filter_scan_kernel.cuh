// Extension (named by the project brief, ABSENT from the reference; default off): the time-parallel form of the
// zero-phase IIR filter — the brief's "warp-level parallel linear-recurrence scan over biquad state".  Same semantics
// as filter_kernels.cuh (scipy.signal.sosfiltfilt: odd extension, sosfilt_zi * first sample, forward pass, reversed
// second pass, trim), different schedule.  Oracle = scipy ("parity unpinned by the reference"), bar 1e-10 of full scale.
//
// The cascade of NSEC direct-form-II-transposed biquads is one linear system  S' = A S + B u  over its D = 2 NSEC delay
// elements.  ONE CTA owns one record and walks it in spans of 32 L rows, one warp per filtered column (so an SM holds
// ~24 independent recurrences to hide the FMA latency of each);
// lane t owns the L consecutive rows [t L, (t+1) L) of the span:
//   pass A   f_t = sum_i A^(L-1-i) B u_i      the state a zero-state chunk ends in: D FMAs per sample against a host table
//   scan     S_(t+1) = A^L S_t + f_t          Kogge-Stone over the lanes with A^(L 2^k), k = 0..4; lane 0 is seeded with
//                                             the state the previous span ended in (registers, no global traffic)
//   pass B   the chunk again from its true state S_t, outputs written over the inputs in shared memory
// The forward pass writes whole rows (pass-through columns included) to y; the same CTA then runs the backward pass
// over y in place, so there is no scratch arena: the data cross HBM four times (x read, y written, y read, y written).
//
// Staging: the lanes of warp 0 move one chunk each with 1-D bulk async copies (cp.async.bulk: global -> shared completing
// on an mbarrier every warp waits on, shared -> global as a bulk group after a CTA barrier).  nbuf = 2 prefetches span
// k+1 while span k is filtered; nbuf = 1 spends the shared memory on longer chunks instead (fewer scan steps per row)
// and lets the other resident CTAs cover the copy.  Chunks sit `pitch = L * row_bytes + 16` apart so that the lanes'
// row accesses spread over the banks.  Rows that are not 16-byte aligned (odd signal counts) take plain copies.
#pragma once
#include "common.cuh"

namespace scgrhc {

constexpr int kTpMaxL = 32;      // rows per lane per span
constexpr int kTpMaxSec = 4;     // sections (an order-8 band-pass); longer cascades use the exact kernel
constexpr int kTpMaxCols = 4;    // filtered columns per launch
constexpr int kTpMaxEdge = 32;   // odd-extension length (3 * ntaps <= 27 for 4 sections); one lane per edge sample

struct SosTpParams {
  const double* x;        // (rows, ncols) input arena
  double* y;              // (rows, ncols) output arena; may be x itself (in place)
  const long long* row0;  // device, n_rec + 1 record boundaries (rows)
  int n_rec, ncols, edge;
  int L;                  // rows per lane per span
  int nbuf;               // staging buffers per warp: 2 = prefetch one span ahead, 1 = none (longer chunks fit instead)
  int bulk;               // 1: bulk async copies; 0: plain warp copies (rows not 16-byte aligned)
  int fcols[kTpMaxCols];
  double c[kTpMaxSec][5];                       // b0, b1, b2, -a1, -a2
  double zi[kTpMaxSec][2];
  double G[kTpMaxL][2 * kTpMaxSec];             // G[i] = A^(L-1-i) B
  double Mp[5][2 * kTpMaxSec][2 * kTpMaxSec];   // Mp[k] = A^(L 2^k)
};

__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int NSEC, int LT>   // LT: compile-time chunk length (0 = P.L): unrolled pass A, table as immediate operands
struct TpWarp {
  static constexpr int D = 2 * NSEC;
  static constexpr int CPW = 1;   // columns per warp (more than one was measured slower: fewer recurrences in flight)
  const SosTpParams& P;
  const int lane, warp, T, nc, L, rb, pitch;
  uint64_t* bars;
  const double* sMp;       // shared copies of P.Mp and P.G: indexed constant-bank loads (LDC) in the loops were the top stall
  const double* sG;
  unsigned char* bufs;
  int col[CPW];            // this warp's filtered columns
  uint32_t uses0, uses1;   // completed phases of each buffer's mbarrier (they run on across the two passes)
  double z[CPW][NSEC][2];  // lane 0: the state the sequence has reached; other lanes: scratch

  __device__ __forceinline__ TpWarp(const SosTpParams& p, int T_, unsigned char* smem)
      : P(p), lane(threadIdx.x & 31), warp(threadIdx.x >> 5), T(T_), nc(p.ncols), L(LT ? LT : p.L), rb(p.ncols * 8),
        pitch((LT ? LT : p.L) * p.ncols * 8 + 16), bars(reinterpret_cast<uint64_t*>(smem)),
        sMp(reinterpret_cast<const double*>(smem + 16)), sG(reinterpret_cast<const double*>(smem + 16 + 5 * D * D * 8)),
        bufs(smem + 16 + 5 * D * D * 8 + (LT ? 0 : kTpMaxL * D * 8)) {
    uses0 = uses1 = 0;
#pragma unroll
    for (int j = 0; j < CPW; ++j) col[j] = p.fcols[warp * CPW + j];
  }

  // one sample of every column of this warp through the cascade (v: inputs in, outputs out)
  __device__ __forceinline__ void step(double (&v)[CPW]) {
#pragma unroll
    for (int j = 0; j < CPW; ++j) {
      double u = v[j];
#pragma unroll
      for (int s = 0; s < NSEC; ++s) {
        const double xn = __fma_rn(P.c[s][0], u, z[j][s][0]);
        const double t = __fma_rn(P.c[s][1], u, z[j][s][1]);
        z[j][s][0] = __fma_rn(P.c[s][3], xn, t);
        z[j][s][1] = __fma_rn(P.c[s][4], xn, __dmul_rn(P.c[s][2], u));
        u = xn;
      }
      v[j] = u;
    }
  }

  __device__ __forceinline__ void init_state(const double (&first)[CPW]) {
#pragma unroll
    for (int j = 0; j < CPW; ++j)
#pragma unroll
      for (int s = 0; s < NSEC; ++s) {
        z[j][s][0] = __dmul_rn(P.zi[s][0], first[j]);
        z[j][s][1] = __dmul_rn(P.zi[s][1], first[j]);
      }
  }

  // rows of span sp that lane t owns: l of them, from row `start` of the record (PASS 1 walks the record backwards)
  template <int PASS>
  __device__ __forceinline__ void geom(int sp, int t, int& start, int& l) const {
    const int e0 = sp * 32 * L;
    const int rows = min(32 * L, T - e0);
    l = max(0, min(L, rows - t * L));
    start = PASS == 0 ? e0 + t * L : T - (e0 + t * L + l);
  }

  // global <-> shared for a whole span.  Bulk: warp 0, one copy per lane.  Plain: every thread of the CTA, between
  // CTA barriers.
  template <int PASS, bool TO_SMEM>
  __device__ __forceinline__ void plain_copy(double* g, int sp) {
    unsigned char* buf = bufs + (size_t)(sp % P.nbuf) * 32 * pitch;
    for (int t = 0; t < 32; ++t) {
      int start, l;
      geom<PASS>(sp, t, start, l);
      if (l == 0) break;
      double* s = reinterpret_cast<double*>(buf + (size_t)t * pitch);
      double* d = g + (long long)start * nc;
      for (int q = threadIdx.x; q < l * nc; q += blockDim.x) {
        if (TO_SMEM) s[q] = d[q]; else d[q] = s[q];
      }
    }
  }

  template <int PASS>
  __device__ __forceinline__ void load(const double* src, int sp) {
    const int b = sp % P.nbuf;
    if (P.bulk) {
      if (warp == 0) {
        bulk_wait_read_all();                           // this lane's last store has left shared memory
        const int rows = min(32 * L, T - sp * 32 * L);
        if (lane == 0) mbar_arrive_expect_tx(&bars[b], (uint32_t)rows * rb);
        __syncwarp();
        int start, l;
        geom<PASS>(sp, lane, start, l);
        if (l > 0) bulk_g2s(bufs + (size_t)b * 32 * pitch + (size_t)lane * pitch, src + (long long)start * nc, (uint32_t)(l * rb), &bars[b]);
      }
    } else {
      plain_copy<PASS, true>(const_cast<double*>(src), sp);
      __syncthreads();
    }
  }

  template <int PASS>
  __device__ __forceinline__ void store(double* dst, int sp) {
    if (P.bulk) {
      fence_async_smem();                               // this thread's generic-proxy writes -> visible to the copy engine
      __syncthreads();                                  // every column of the span is filtered
      if (warp == 0) {
        int start, l;
        geom<PASS>(sp, lane, start, l);
        if (l > 0) bulk_s2g(dst + (long long)start * nc, bufs + (size_t)(sp % P.nbuf) * 32 * pitch + (size_t)lane * pitch, (uint32_t)(l * rb));
        bulk_commit();
      }
    } else {
      __syncthreads();
      plain_copy<PASS, false>(dst, sp);
      __syncthreads();
    }
  }

  // one direction over the whole record: src -> dst (dst rows are complete rows; src == dst is fine, a span is staged
  // completely before it is written back)
  template <int PASS>
  __device__ __forceinline__ void run_pass(const double* src, double* dst) {
    const int nspans = (T + 32 * L - 1) / (32 * L);
    const int nbuf = P.nbuf;
    if (nbuf == 2) load<PASS>(src, 0);
    for (int sp = 0; sp < nspans; ++sp) {
      if (nbuf == 2) {
        if (sp + 1 < nspans) load<PASS>(src, sp + 1);   // its buffer was released by the barrier of span sp - 1
      } else {
        load<PASS>(src, sp);
      }
      const int b = sp % nbuf;
      if (P.bulk) {
        mbar_wait(&bars[b], (b ? uses1 : uses0) & 1u);
        if (b) ++uses1; else ++uses0;
      }
      int start, l;
      geom<PASS>(sp, lane, start, l);
      double* chunk = reinterpret_cast<double*>(bufs + (size_t)b * 32 * pitch + (size_t)lane * pitch);

      // pass A: where a zero-state chunk ends
      double f[CPW][D];
#pragma unroll
      for (int j = 0; j < CPW; ++j)
#pragma unroll
        for (int a = 0; a < D; ++a) f[j][a] = 0.0;
      if (l == L) {
        auto row_a = [&](int i) {
          const double* row = chunk + (PASS == 0 ? i : L - 1 - i) * nc;
#pragma unroll
          for (int j = 0; j < CPW; ++j) {
            const double u = row[col[j]];
#pragma unroll
            for (int a = 0; a < D; ++a) f[j][a] = __fma_rn(LT ? P.G[i][a] : sG[i * D + a], u, f[j][a]);
          }
        };
        if constexpr (LT > 0) {
#pragma unroll
          for (int i = 0; i < LT; ++i) row_a(i);
        } else {
#pragma unroll 4
          for (int i = 0; i < L; ++i) row_a(i);
        }
      }
      if (lane == 0) {                                  // seed: the state this span starts in, pushed through chunk 0
#pragma unroll
        for (int j = 0; j < CPW; ++j)
#pragma unroll
          for (int a = 0; a < D; ++a)
#pragma unroll
            for (int bb = 0; bb < D; ++bb)
              if ((bb >> 1) <= (a >> 1)) f[j][a] = __fma_rn(P.Mp[0][a][bb], z[j][bb >> 1][bb & 1], f[j][a]);
      }
      // inclusive scan over the lanes: f_t <- state at the END of chunk t (A is block lower triangular: a section's
      // delay elements depend on the sections before it only)
#pragma unroll
      for (int k = 0; k < 5; ++k) {                     // unrolled: the next step's matrix loads overlap this step's FMAs
        double g[CPW][D];
#pragma unroll
        for (int j = 0; j < CPW; ++j)
#pragma unroll
          for (int a = 0; a < D; ++a) {
            const double up = __shfl_up_sync(kFull, f[j][a], 1 << k);
            g[j][a] = lane >= (1 << k) ? up : 0.0;
          }
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
          for (int bb = 0; bb < D; ++bb)
            if ((bb >> 1) <= (a >> 1)) {
              const double m = sMp[(k * D + a) * D + bb];
#pragma unroll
              for (int j = 0; j < CPW; ++j) f[j][a] = __fma_rn(m, g[j][bb], f[j][a]);
            }
      }
      // exclusive: lane t starts where lane t-1 ended; lane 0 keeps the carried state
#pragma unroll
      for (int j = 0; j < CPW; ++j)
#pragma unroll
        for (int a = 0; a < D; ++a) {
          const double up = __shfl_up_sync(kFull, f[j][a], 1);
          if (lane > 0) z[j][a >> 1][a & 1] = up;
        }
      // pass B: the chunk from its true state, outputs over the inputs
#pragma unroll 4
      for (int i = 0; i < l; ++i) {
        double* row = chunk + (PASS == 0 ? i : l - 1 - i) * nc;
        double v[CPW];
#pragma unroll
        for (int j = 0; j < CPW; ++j) v[j] = row[col[j]];
        step(v);
#pragma unroll
        for (int j = 0; j < CPW; ++j) row[col[j]] = v[j];
      }
      // carry: the state after the span's last row lives in the lane that owns it
      const int rows = min(32 * L, T - sp * 32 * L);
      const int tl = (rows - 1) / L;
#pragma unroll
      for (int j = 0; j < CPW; ++j)
#pragma unroll
        for (int s = 0; s < NSEC; ++s) {
          z[j][s][0] = __shfl_sync(kFull, z[j][s][0], tl);
          z[j][s][1] = __shfl_sync(kFull, z[j][s][1], tl);
        }
      store<PASS>(dst, sp);
    }
  }
};

// grid = records, block = 32 * filtered columns
template <int NSEC, int LT>
__global__ void __launch_bounds__(32 * kTpMaxCols, 6) sosfilt_tp_kernel(const __grid_constant__ SosTpParams P) {
  extern __shared__ __align__(16) unsigned char tp_smem[];
  const int rec = blockIdx.x;
  const long long r0 = P.row0[rec];
  const int T = (int)(P.row0[rec + 1] - r0);              // host guarantees edge < T < 2^31
  const int nc = P.ncols, edge = P.edge;
  TpWarp<NSEC, LT> W(P, T, tp_smem);
  constexpr int CPW = 1;
  const int lane = W.lane;
  if (P.bulk) {
    if (threadIdx.x == 0) {
      for (int b = 0; b < P.nbuf; ++b) mbar_init(&W.bars[b], 1);
    }
    fence_barrier_init();
  }
  {
    constexpr int D = 2 * NSEC;
    double* mp = reinterpret_cast<double*>(tp_smem + 16);
    for (int q = threadIdx.x; q < 5 * D * D; q += blockDim.x) mp[q] = P.Mp[q / (D * D)][(q / D) % D][q % D];
    if (LT == 0)
      for (int q = threadIdx.x; q < P.L * D; q += blockDim.x) mp[5 * D * D + q] = P.G[q / D][q % D];
  }
  __syncthreads();
  const double* xr = P.x + r0 * nc;
  double* yr = P.y + r0 * nc;

  // Everything the odd extension needs is read before the first span is written (y may be x): lane e holds
  // x[edge - e] for the head (x_ext[e] = 2 x[0] - x[edge - e]) and x[T - 2 - e] for the tail.
  double x0[CPW], xl[CPW], hx[CPW], tx[CPW], tail[CPW];
#pragma unroll
  for (int j = 0; j < CPW; ++j) {
    const int col = W.col[j];
    x0[j] = xr[col];
    xl[j] = xr[(long long)(T - 1) * nc + col];
    hx[j] = lane < edge ? xr[(long long)(edge - lane) * nc + col] : 0.0;
    tx[j] = lane < edge ? xr[(long long)(T - 2 - lane) * nc + col] : 0.0;
    tail[j] = 0.0;
  }
  // forward: head extension serially (every lane computes it; its outputs only feed samples that get trimmed)
  double v[CPW];
#pragma unroll
  for (int j = 0; j < CPW; ++j) v[j] = __dsub_rn(2.0 * x0[j], __shfl_sync(kFull, hx[j], 0));
  W.init_state(v);
  for (int e = 0; e < edge; ++e) {
#pragma unroll
    for (int j = 0; j < CPW; ++j) v[j] = __dsub_rn(2.0 * x0[j], __shfl_sync(kFull, hx[j], e));
    W.step(v);
  }
  W.template run_pass<0>(xr, yr);
  // tail extension: its outputs are the first inputs of the backward pass; lane k keeps output k
  for (int k = 0; k < edge; ++k) {
#pragma unroll
    for (int j = 0; j < CPW; ++j) v[j] = __dsub_rn(2.0 * xl[j], __shfl_sync(kFull, tx[j], k));
    W.step(v);
    if (lane == k) {
#pragma unroll
      for (int j = 0; j < CPW; ++j) tail[j] = v[j];
    }
  }
  // backward: starts from zi * (last forward output), runs over the reversed tail, then over y in place
#pragma unroll
  for (int j = 0; j < CPW; ++j) v[j] = __shfl_sync(kFull, tail[j], edge - 1);
  W.init_state(v);
  for (int k = edge - 1; k >= 0; --k) {
#pragma unroll
    for (int j = 0; j < CPW; ++j) v[j] = __shfl_sync(kFull, tail[j], k);
    W.step(v);
  }
  if (P.bulk && W.warp == 0) {
    bulk_wait_all();                                     // the forward pass's rows have landed in y
    __syncwarp();
  }
  W.template run_pass<1>(yr, yr);
  if (P.bulk && W.warp == 0) bulk_wait_read_all();       // shared memory must outlive the last stores' reads
}

}  // namespace scgrhc
