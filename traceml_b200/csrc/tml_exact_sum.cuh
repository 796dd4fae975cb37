// tml_exact_sum.cuh -- K3e: the seven per-rank window sums in the REFERENCE'S summation order,
// bit for bit, for any window size (math and rationale: tml_exact_sum.h).  Included by
// tml_engine.cu; replaces round 1's single-thread dependency chain (k_seq_sums, <= 2^17 rows).
//
//   X1 k_xs_partial   chunk (256 rows) -> 7 approximate sums          HBM read, 64 B/row
//   X2 k_xs_plan      prefix of the chunk sums -> exponent per (chunk, chain) or UNSAFE / ZERO
//   X3 k_xs_compose   chunk -> 7 (c0, c1) maps under the planned exponent   HBM read, 64 B/row
//   X3b k_xs_groups   32 chunks -> one map (warp-ordered composition)
//   X4 k_xs_walk      one warp per chain walks groups / chunks / 32-row tiles; crossings and the
//                     start-up from 0 are redone with real dependent adds on that tile only
//
// Rows are read twice (X1, X3): 128 B/row of local HBM traffic.  At R > 1 that hides under the
// NVLink-bound K4 on a side stream; the R = 1 bulk path does not need reference-order sums at all
// (no second rank to break a tie against) and skips K3e.
#pragma once
#include "tml_exact_sum.h"

struct XsSrc {
  const tml_window_row* rows;
  const u8* flags;      // window mode: only rows with (flags & need) == need count; others add +0.0
  u32 need;
  long long first, last;  // inclusive; summation order is last, last - 1, ..., first (newest row first)
  int aligned;            // 1: addend 4 is max(0, traced) (alignment.py:72), 0: the step wall (model.py:265)
  const tml_window_row* xrows;  // aligned mode: where the aligned rows live is decided on the device
  const u32* noncontig;
  const u32* sel_rows;
  long long dense_first;
};

__device__ __forceinline__ const tml_window_row* xs_rows(const XsSrc& s) {
  if (!s.aligned) return s.rows;
  if (s.dense_first >= 0) return s.rows + s.dense_first;
  return (*s.noncontig) ? s.xrows : (s.rows + s.sel_rows[0]);
}

// the seven addends of summation position p (p = 0 is the newest row); zeros past the end
__device__ __forceinline__ void xs_addends(const tml_window_row* __restrict__ rows, const XsSrc& s, long long p,
                                           long long n, double (&o)[7]) {
#pragma unroll
  for (int k = 0; k < 7; ++k) o[k] = 0.0;
  if (p >= n) return;
  const long long i = s.last - p;
  if (s.flags && ((s.flags[i] & s.need) != s.need)) return;
  const uint4* src = reinterpret_cast<const uint4*>(rows + i);
  const uint4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
  const double dl = __longlong_as_double((long long)((u64)a.x | ((u64)a.y << 32)));
  const double fwd = __longlong_as_double((long long)((u64)b.x | ((u64)b.y << 32)));
  const double bwd = __longlong_as_double((long long)((u64)b.z | ((u64)b.w << 32)));
  const double opt = __longlong_as_double((long long)((u64)c.x | ((u64)c.y << 32)));
  const double wall = __longlong_as_double((long long)((u64)c.z | ((u64)c.w << 32)));
  const double compute = (fwd + bwd) + opt;
  const double traced = fmax(wall, compute);
  o[0] = dl; o[1] = fwd; o[2] = bwd; o[3] = opt;
  o[4] = s.aligned ? fmax(0.0, traced) : wall;
  o[5] = traced;
  o[6] = dl + traced;
}

__device__ __forceinline__ u64 shfl_down_u64(u64 v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ u64 shfl_idx_u64(u64 v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// ordered composition over the warp: lane 0 receives f[0] o f[1] o ... o f[31] (f[0] applied first)
__device__ __forceinline__ XsFn xs_warp_compose(XsFn f, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    XsFn g;
    g.c0 = shfl_down_u64(f.c0, d);
    g.c1 = shfl_down_u64(f.c1, d);
    if ((lane & (2 * d - 1)) == 0) f = xs_compose(f, g);
  }
  return f;
}

// ---- X1
__global__ void __launch_bounds__(XS_CHUNK) k_xs_partial(const XsSrc s, long long n, long long nchunks,
                                                         double* __restrict__ csum /* [nchunks][8] */) {
  __shared__ double s_part[XS_CHUNK / 32][7];
  const tml_window_row* rows = xs_rows(s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    double o[7];
    xs_addends(rows, s, ch * XS_CHUNK + threadIdx.x, n, o);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      double x = o[k];
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) x += shfl_xor_f64(x, m);
      if (lane == 0) s_part[warp][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
      double x = 0.0;
#pragma unroll
      for (int w = 0; w < XS_CHUNK / 32; ++w) x += s_part[w][threadIdx.x];
      csum[ch * 8 + threadIdx.x] = x;
    }
    __syncthreads();
  }
}

// ---- X2: one CTA; thread t owns a contiguous run of chunks
__global__ void __launch_bounds__(1024) k_xs_plan(const double* __restrict__ csum, long long nchunks,
                                                  int* __restrict__ plan /* [nchunks][8] */) {
  __shared__ double s_warp[32][7];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const long long per = (nchunks + 1023) / 1024;
  const long long lo = (long long)t * per, hi = (lo + per < nchunks) ? lo + per : nchunks;
  double tot[7] = {0, 0, 0, 0, 0, 0, 0};
  for (long long c = lo; c < hi; ++c)
#pragma unroll
    for (int k = 0; k < 7; ++k) tot[k] += csum[c * 8 + k];
  double excl[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) {  // block-wide exclusive scan of tot[k]
    double incl = tot[k];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      double y = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += y;
    }
    if (lane == 31) s_warp[warp][k] = incl;
    excl[k] = incl - tot[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      double w = s_warp[lane][k], incl = w;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        double y = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += y;
      }
      s_warp[lane][k] = incl - w;
    }
  }
  __syncthreads();
  double run[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) run[k] = s_warp[warp][k] + excl[k];
  for (long long c = lo; c < hi; ++c) {
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const double a = run[k], b = a + csum[c * 8 + k];
      plan[c * 8 + k] = xs_plan(a, b);
      run[k] = b;
    }
  }
}

// ---- X3
__global__ void __launch_bounds__(XS_CHUNK) k_xs_compose(const XsSrc s, long long n, long long nchunks,
                                                         const int* __restrict__ plan,
                                                         XsFn* __restrict__ fn /* [nchunks][7] */) {
  __shared__ XsFn s_fn[XS_CHUNK / 32][7];
  const tml_window_row* rows = xs_rows(s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long long p = ch * XS_CHUNK + threadIdx.x;
    double o[7];
    xs_addends(rows, s, p, n, o);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int e = plan[ch * 8 + k];
      XsFn f = (e >= 1 && p < n) ? xs_elem(o[k], e) : xs_identity();
      f = xs_warp_compose(f, lane);
      if (lane == 0) s_fn[warp][k] = f;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
      const int e = plan[ch * 8 + threadIdx.x];
      XsFn f = s_fn[0][threadIdx.x];
#pragma unroll
      for (int w = 1; w < XS_CHUNK / 32; ++w) f = xs_compose(f, s_fn[w][threadIdx.x]);
      if (e == XS_PLAN_ZERO) f = xs_identity();
      else if (e < 1) f = xs_invalid();
      fn[ch * 7 + threadIdx.x] = f;
    }
    __syncthreads();
  }
}

// ---- X3b: one warp per (group, chain)
__global__ void __launch_bounds__(256) k_xs_groups(const XsFn* __restrict__ fn, const int* __restrict__ plan,
                                                   long long nchunks, long long ngroups,
                                                   XsFn* __restrict__ gfn /* [ngroups][7] */,
                                                   int* __restrict__ gplan /* [ngroups][8] */) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wid >= ngroups * 7) return;
  const long long g = wid / 7;
  const int k = (int)(wid % 7);
  const long long ch = g * XS_GROUP + lane;
  int e = XS_PLAN_ZERO;
  XsFn f = xs_identity();
  if (ch < nchunks) { e = plan[ch * 8 + k]; f = fn[ch * 7 + k]; }
  // one exponent for every non-zero chunk of the group, else the group is walked chunk by chunk
  int emax = e;
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) { int y = __shfl_xor_sync(0xffffffffu, emax, m); emax = y > emax ? y : emax; }
  const bool ok = (e == emax) || (e == XS_PLAN_ZERO);
  const bool all_ok = __all_sync(0xffffffffu, ok) && emax >= 1;
  f = xs_warp_compose(f, lane);
  if (lane == 0) {
    gfn[g * 7 + k] = all_ok ? f : xs_invalid();
    gplan[g * 8 + k] = (emax == XS_PLAN_ZERO) ? XS_PLAN_ZERO : (all_ok ? emax : XS_PLAN_UNSAFE);
  }
}

// ---- X4: warp k walks chain k.  Every lane carries the same running sum (the updates are
// deterministic functions of broadcast values), so nothing has to be re-broadcast.
__global__ void __launch_bounds__(7 * 32) k_xs_walk(const XsSrc s, long long n, long long nchunks, long long ngroups,
                                                    const XsFn* __restrict__ fn, const int* __restrict__ plan,
                                                    const XsFn* __restrict__ gfn, const int* __restrict__ gplan,
                                                    int planned, double* __restrict__ out,
                                                    unsigned long long* __restrict__ stats /* [7]: rows added one by one */) {
  const tml_window_row* rows = xs_rows(s);
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double sum = 0.0;
  unsigned long long slow_rows = 0;
  for (long long gb = 0; gb < ngroups; gb += 32) {
    XsFn gf = xs_invalid();
    int ge = XS_PLAN_UNSAFE;
    if (planned && gb + lane < ngroups) { gf = gfn[(gb + lane) * 7 + k]; ge = gplan[(gb + lane) * 8 + k]; }
    const int gcount = (int)((ngroups - gb) < 32 ? (ngroups - gb) : 32);
    for (int j = 0; j < gcount; ++j) {
      XsFn F;
      F.c0 = shfl_idx_u64(gf.c0, j); F.c1 = shfl_idx_u64(gf.c1, j);
      const int E = __shfl_sync(0xffffffffu, ge, j);
      if (E == XS_PLAN_ZERO) continue;
      if (E >= 1 && xs_apply(&sum, F, E)) continue;
      // ---- the group, chunk by chunk
      const long long g = gb + j;
      const long long c0 = g * XS_GROUP;
      XsFn cf = xs_invalid();
      int ce = XS_PLAN_UNSAFE;
      if (planned && c0 + lane < nchunks) { cf = fn[(c0 + lane) * 7 + k]; ce = plan[(c0 + lane) * 8 + k]; }
      const int ccount = (int)((nchunks - c0) < XS_GROUP ? (nchunks - c0) : XS_GROUP);
      for (int q = 0; q < ccount; ++q) {
        XsFn C;
        C.c0 = shfl_idx_u64(cf.c0, q); C.c1 = shfl_idx_u64(cf.c1, q);
        const int CE = __shfl_sync(0xffffffffu, ce, q);
        if (CE == XS_PLAN_ZERO) continue;
        if (CE >= 1 && xs_apply(&sum, C, CE)) continue;
        // ---- the chunk, 32-row tile by tile: compose under the TRUE exponent, else add one by one
        const long long p0 = (c0 + q) * XS_CHUNK;
        double x[XS_CHUNK / 32];
#pragma unroll
        for (int t = 0; t < XS_CHUNK / 32; ++t) {
          double o[7];
          xs_addends(rows, s, p0 + t * 32 + lane, n, o);
          x[t] = o[0];
#pragma unroll
          for (int m = 1; m < 7; ++m) x[t] = (k == m) ? o[m] : x[t];
        }
#pragma unroll
        for (int t = 0; t < XS_CHUNK / 32; ++t) {
          if (p0 + t * 32 >= n) break;
          const int eb = xs_exp(sum);
          bool done = false;
          if (eb >= 1 && eb < 0x7ff && sum > 0.0) {
            XsFn f = xs_warp_compose(xs_elem(x[t], eb), lane);
            f.c0 = shfl_idx_u64(f.c0, 0); f.c1 = shfl_idx_u64(f.c1, 0);
            done = xs_apply(&sum, f, eb);
          }
          if (!done) {
            for (int r = 0; r < 32; ++r) sum += __shfl_sync(0xffffffffu, x[t], r);  // rows past n are +0.0
            slow_rows += 32;
          }
        }
      }
    }
  }
  if (lane == 0) { out[k] = sum; if (stats) stats[k] = slow_rows; }
}
