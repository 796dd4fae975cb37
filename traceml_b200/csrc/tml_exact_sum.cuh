// tml_exact_sum.cuh -- K3e: the seven per-rank window sums in the REFERENCE'S summation order,
// bit for bit, for any window size (math and rationale: tml_exact_sum.h).  Included by
// tml_engine.cu; replaces round 1's single-thread dependency chain (k_seq_sums, <= 2^17 rows).
//
// Work unit: a CHUNK of 256 consecutive rows (in summation order) = one warp, lane l owning the 8
// consecutive rows 8l .. 8l+7, so the ordered composition is 8 serial steps per lane plus ONE
// 5-level shuffle tree per chunk and chain.  A CTA owns a contiguous run of chunks.
//
//   X1 k_xs_partial   chunk -> 7 approximate sums + their running prefix inside the CTA's run
//   X2 k_xs_bscan     exclusive scan of the CTA totals (one small CTA)
//   X3 k_xs_compose   chunk -> exponent plan + 7 (c0, c1) maps; for a chunk whose running sum
//                     changes binade: the eight 32-row TILE maps under both candidate exponents
//   X3b k_xs_groups   32 chunks -> one map (warp-ordered composition)
//   X4 k_xs_walk      one warp per chain: groups -> chunks -> tiles; only the tile that contains
//                     a binade crossing (and the start-up from 0) is redone with real adds
//
// Rows are read twice (X1, X3): 128 B/row of local HBM traffic.  At R > 1 that hides under the
// NVLink-bound K4 on a side stream; the R = 1 bulk path does not need reference-order sums at all
// (no second rank to break a tie against) and skips K3e.
#pragma once
#include "tml_exact_sum.h"

#define XS_WARPS 8                       // chunks per CTA trip
#define XS_ROWS_PER_LANE (XS_CHUNK / 32)  // 8
#define XS_TILES (XS_CHUNK / 32)          // 32-row tiles per chunk (= 4 lanes each)
#define XS_SLOT_CAP 4096                  // (chunk, chain) pairs that may carry tile maps
#define XS_PLAN_SLOT0 (-3)                // plan <= XS_PLAN_SLOT0: unsafe, tile maps in slot -(plan) - 3

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

struct XsTileMaps {  // one unsafe (chunk, chain): tile maps under exponent ea (h = 0) and ea + 1 (h = 1)
  XsFn f[2][XS_TILES];
};

struct XsWork {  // device workspace of one launch
  double* csum;   // [nchunks][8] approximate chunk sums
  double* cpre;   // [nchunks][8] their exclusive prefix inside the owning CTA's run
  double* btot;   // [nblocks][8]
  double* bpre;   // [nblocks][8]
  int* plan;      // [nchunks][8]
  int* ea;        // [nchunks][8] candidate exponent of an unsafe chunk (tile maps: ea, ea + 1)
  XsFn* fn;       // [nchunks][7]
  XsFn* gfn;      // [ngroups][7]
  int* gplan;     // [ngroups][8]
  XsTileMaps* tiles;  // [XS_SLOT_CAP]
  unsigned int* nslots;
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

// ordered composition over the warp: lane 0 receives f[0] o f[1] o ... o f[31] (f[0] applied first);
// after the levels d = 1, 2 lanes 0, 4, 8, ... hold the maps of their 4-lane group (one 32-row tile)
__device__ __forceinline__ XsFn xs_warp_compose(XsFn f, int lane, XsFn* tile_out = nullptr) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    XsFn g;
    g.c0 = shfl_down_u64(f.c0, d);
    g.c1 = shfl_down_u64(f.c1, d);
    if ((lane & (2 * d - 1)) == 0) f = xs_compose(f, g);
    if (d == 2 && tile_out) *tile_out = f;
  }
  return f;
}

// same tree on unsealed maps (callers seal lane 0's result)
__device__ __forceinline__ XsFn xs_warp_compose_raw(XsFn f, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    XsFn g;
    g.c0 = shfl_down_u64(f.c0, d);
    g.c1 = shfl_down_u64(f.c1, d);
    if ((lane & (2 * d - 1)) == 0) f = xs_compose_raw(f, g);
  }
  return f;
}

__device__ __forceinline__ void xs_block_range(long long nchunks, long long* lo, long long* hi) {
  long long per = (nchunks + gridDim.x - 1) / gridDim.x;
  per = (per + XS_WARPS - 1) / XS_WARPS * XS_WARPS;
  *lo = (long long)blockIdx.x * per;
  *hi = (*lo + per < nchunks) ? *lo + per : nchunks;
}

// ---- X1
__global__ void __launch_bounds__(XS_WARPS * 32) k_xs_partial(const XsSrc s, long long n, long long nchunks, XsWork w) {
  __shared__ double s_part[XS_WARPS][7];
  const tml_window_row* rows = xs_rows(s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long c_lo, c_hi;
  xs_block_range(nchunks, &c_lo, &c_hi);
  double run = 0.0;  // threads 0..6: running prefix of chain `threadIdx.x` inside this CTA's run
  for (long long base = c_lo; base < c_hi; base += XS_WARPS) {
    const long long ch = base + warp;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    if (ch < c_hi) {
#pragma unroll
      for (int j = 0; j < XS_ROWS_PER_LANE; ++j) {
        double o[7];
        xs_addends(rows, s, ch * XS_CHUNK + lane * XS_ROWS_PER_LANE + j, n, o);
#pragma unroll
        for (int k = 0; k < 7; ++k) acc[k] += o[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      double x = acc[k];
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) x += shfl_xor_f64(x, m);
      if (lane == 0) s_part[warp][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
#pragma unroll
      for (int q = 0; q < XS_WARPS; ++q) {
        if (base + q < c_hi) {
          const double v = s_part[q][threadIdx.x];
          w.cpre[(base + q) * 8 + threadIdx.x] = run;
          w.csum[(base + q) * 8 + threadIdx.x] = v;
          run += v;
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x < 7) w.btot[(long long)blockIdx.x * 8 + threadIdx.x] = run;
}

// ---- X2: exclusive scan of the CTA totals, nblocks <= 1024
__global__ void __launch_bounds__(1024) k_xs_bscan(XsWork w, int nblocks) {
  __shared__ double s_warp[32][7];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  double v[7], excl[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    v[k] = (t < nblocks) ? w.btot[(long long)t * 8 + k] : 0.0;
    double incl = v[k];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      double y = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += y;
    }
    if (lane == 31) s_warp[warp][k] = incl;
    excl[k] = incl - v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      double x = s_warp[lane][k], incl = x;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        double y = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += y;
      }
      s_warp[lane][k] = incl - x;
    }
  }
  __syncthreads();
  if (t < nblocks) {
#pragma unroll
    for (int k = 0; k < 7; ++k) w.bpre[(long long)t * 8 + k] = s_warp[warp][k] + excl[k];
  }
  if (t == 0) *w.nslots = 0u;
}

// ---- X3
__global__ void __launch_bounds__(XS_WARPS * 32) k_xs_compose(const XsSrc s, long long n, long long nchunks, XsWork w) {
  const tml_window_row* rows = xs_rows(s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long c_lo, c_hi;
  xs_block_range(nchunks, &c_lo, &c_hi);
  for (long long ch = c_lo + warp; ch < c_hi; ch += XS_WARPS) {
    int e[7];
    double lo[7];
    bool any_unsafe = false;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      lo[k] = w.bpre[(long long)blockIdx.x * 8 + k] + w.cpre[ch * 8 + k];
      e[k] = xs_plan(lo[k], lo[k] + w.csum[ch * 8 + k]);
      any_unsafe = any_unsafe || (e[k] == XS_PLAN_UNSAFE && lo[k] > 0.0);
    }
    XsFn f[7];
    unsigned bad = 0u;  // bit k: some element of chain k does not fit under the planned exponent
#pragma unroll
    for (int k = 0; k < 7; ++k) f[k] = xs_identity();
    const long long p0 = ch * XS_CHUNK + lane * XS_ROWS_PER_LANE;
#pragma unroll
    for (int j = 0; j < XS_ROWS_PER_LANE; ++j) {
      double o[7];
      xs_addends(rows, s, p0 + j, n, o);
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        if (e[k] >= 1) {
          XsFn g;
          if (!xs_elem_raw(o[k], e[k], &g)) bad |= 1u << k;
          f[k] = xs_compose_raw(f[k], g);
        }
      }
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      XsFn r = xs_seal(xs_warp_compose_raw(f[k], lane), ((bad >> k) & 1u) == 0u);  // 256 maps < 2^53: < 2^61
      if (e[k] == XS_PLAN_ZERO) r = xs_identity();
      else if (e[k] < 1) r = xs_invalid();
      if (lane == 0) w.fn[ch * 7 + k] = r;
    }
    // a chunk whose running sum changes binade: tile maps under both candidate exponents, so the
    // walk redoes only the ONE tile that holds the crossing (rare: ~ one chunk per binade and chain)
    if (any_unsafe) {
      for (int k = 0; k < 7; ++k) {
        if (!(e[k] == XS_PLAN_UNSAFE && lo[k] > 0.0)) continue;
        const int ea = xs_exp(lo[k] * (1.0 - 1.0e-6));
        unsigned int slot = 0;
        if (lane == 0) slot = atomicAdd(w.nslots, 1u);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= XS_SLOT_CAP || ea < 1 || ea >= 0x7fd) continue;  // plain UNSAFE: the walk adds row by row
        XsFn fa = xs_identity(), fb = xs_identity();
        for (int j = 0; j < XS_ROWS_PER_LANE; ++j) {
          double o[7];
          xs_addends(rows, s, p0 + j, n, o);
          double x = o[0];
#pragma unroll
          for (int m = 1; m < 7; ++m) x = (k == m) ? o[m] : x;
          fa = xs_compose(fa, xs_elem(x, ea));
          fb = xs_compose(fb, xs_elem(x, ea + 1));
        }
        XsFn ta, tb;
        xs_warp_compose(fa, lane, &ta);
        xs_warp_compose(fb, lane, &tb);
        if ((lane & 3) == 0) {
          w.tiles[slot].f[0][lane >> 2] = ta;
          w.tiles[slot].f[1][lane >> 2] = tb;
        }
        e[k] = XS_PLAN_SLOT0 - (int)slot;
        if (lane == 0) w.ea[ch * 8 + k] = ea;
      }
    }
    if (lane < 7) {
      int ek = e[0];
#pragma unroll
      for (int m = 1; m < 7; ++m) ek = (lane == m) ? e[m] : ek;
      w.plan[ch * 8 + lane] = ek;
    }
  }
}

// ---- X3b: one warp per (group, chain)
__global__ void __launch_bounds__(256) k_xs_groups(XsWork w, long long nchunks, long long ngroups) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wid >= ngroups * 7) return;
  const long long g = wid / 7;
  const int k = (int)(wid % 7);
  const long long ch = g * XS_GROUP + lane;
  int e = XS_PLAN_ZERO;
  XsFn f = xs_identity();
  if (ch < nchunks) { e = w.plan[ch * 8 + k]; f = w.fn[ch * 7 + k]; }
  if (e <= XS_PLAN_SLOT0) e = XS_PLAN_UNSAFE;
  // one exponent for every non-zero chunk of the group, else the group is walked chunk by chunk
  int emax = e;
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) { int y = __shfl_xor_sync(0xffffffffu, emax, m); emax = y > emax ? y : emax; }
  const bool ok = (e == emax) || (e == XS_PLAN_ZERO);
  const bool all_ok = __all_sync(0xffffffffu, ok) && emax >= 1;
  f = xs_warp_compose(f, lane);
  if (lane == 0) {
    w.gfn[g * 7 + k] = all_ok ? f : xs_invalid();
    w.gplan[g * 8 + k] = (emax == XS_PLAN_ZERO) ? XS_PLAN_ZERO : (all_ok ? emax : XS_PLAN_UNSAFE);
  }
}

// ---- X4: warp k walks chain k.  Every lane carries the same running sum (the updates are
// deterministic functions of broadcast values), so nothing has to be re-broadcast.
__device__ __forceinline__ void xs_seq_tile(const tml_window_row* rows, const XsSrc& s, long long p0, long long n,
                                            int k, int lane, double* sum) {
  double o[7];
  xs_addends(rows, s, p0 + lane, n, o);
  double x = o[0];
#pragma unroll
  for (int m = 1; m < 7; ++m) x = (k == m) ? o[m] : x;
  double acc = *sum;
#pragma unroll 8
  for (int r = 0; r < 32; ++r) acc += __shfl_sync(0xffffffffu, x, r);  // rows past n are +0.0
  *sum = acc;
}

__global__ void __launch_bounds__(7 * 32) k_xs_walk(const XsSrc s, long long n, long long nchunks, long long ngroups,
                                                    XsWork w, int planned, double* __restrict__ out,
                                                    unsigned long long* __restrict__ stats /* [7]: rows added one by one */) {
  __shared__ XsFn s_g[7][32];
  __shared__ int s_ge[7][32];
  __shared__ XsFn s_c[7][32];
  __shared__ int s_ce[7][32];
  __shared__ XsTileMaps s_t[7];
  const tml_window_row* rows = xs_rows(s);
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double sum = 0.0;
  unsigned long long slow_rows = 0;
  for (long long gb = 0; gb < ngroups; gb += 32) {
    {
      XsFn gf = xs_invalid();
      int ge = XS_PLAN_UNSAFE;
      if (planned && gb + lane < ngroups) { gf = w.gfn[(gb + lane) * 7 + k]; ge = w.gplan[(gb + lane) * 8 + k]; }
      s_g[k][lane] = gf; s_ge[k][lane] = ge;
    }
    __syncwarp();
    const int gcount = (int)((ngroups - gb) < 32 ? (ngroups - gb) : 32);
    for (int j = 0; j < gcount; ++j) {
      const int E = s_ge[k][j];
      if (E == XS_PLAN_ZERO) continue;
      if (E >= 1 && xs_apply(&sum, s_g[k][j], E)) continue;
      // ---- the group, chunk by chunk
      const long long c0 = (gb + j) * XS_GROUP;
      __syncwarp();
      {
        XsFn cf = xs_invalid();
        int ce = XS_PLAN_UNSAFE;
        if (planned && c0 + lane < nchunks) { cf = w.fn[(c0 + lane) * 7 + k]; ce = w.plan[(c0 + lane) * 8 + k]; }
        s_c[k][lane] = cf; s_ce[k][lane] = ce;
      }
      __syncwarp();
      const int ccount = (int)((nchunks - c0) < XS_GROUP ? (nchunks - c0) : XS_GROUP);
      for (int q = 0; q < ccount; ++q) {
        const int CE = s_ce[k][q];
        if (CE == XS_PLAN_ZERO) continue;
        if (CE >= 1 && xs_apply(&sum, s_c[k][q], CE)) continue;
        const long long p0 = (c0 + q) * XS_CHUNK;
        if (CE <= XS_PLAN_SLOT0) {
          // ---- tile maps under ea / ea + 1; the tile that crosses is added row by row
          const int slot = XS_PLAN_SLOT0 - CE;
          // one round trip: candidate exponent, tile maps and this chain's addends of all 8 tiles
          const int ea = w.ea[(c0 + q) * 8 + k];
          XsFn tm = xs_invalid();
          if (lane < 2 * XS_TILES) tm = (&w.tiles[slot].f[0][0])[lane];
          double x[XS_TILES];
#pragma unroll
          for (int t = 0; t < XS_TILES; ++t) {
            double o[7];
            xs_addends(rows, s, p0 + t * 32 + lane, n, o);
            x[t] = o[0];
#pragma unroll
            for (int m = 1; m < 7; ++m) x[t] = (k == m) ? o[m] : x[t];
          }
          __syncwarp();
          if (lane < 2 * XS_TILES) (&s_t[k].f[0][0])[lane] = tm;
          __syncwarp();
#pragma unroll
          for (int t = 0; t < XS_TILES; ++t) {
            if (p0 + t * 32 < n) {
              const int eb = xs_exp(sum);
              const int h = eb - ea;
              if (!((h == 0 || h == 1) && xs_apply(&sum, s_t[k].f[h][t], eb))) {
                double acc = sum;
#pragma unroll 8
                for (int r = 0; r < 32; ++r) acc += __shfl_sync(0xffffffffu, x[t], r);
                sum = acc;
                slow_rows += 32;
              }
            }
          }
          continue;
        }
        // ---- no maps (start-up from 0, or the slot table is full): every row is a real add
#pragma unroll 1
        for (int t = 0; t < XS_TILES; ++t) {
          if (p0 + t * 32 >= n) break;
          const int eb = xs_exp(sum);
          bool done = false;
          if (sum > 0.0 && eb >= 1 && eb < 0x7ff && p0 > 0) {  // compose this tile under the true exponent
            double o[7];
            xs_addends(rows, s, p0 + t * 32 + lane, n, o);
            double x = o[0];
#pragma unroll
            for (int m = 1; m < 7; ++m) x = (k == m) ? o[m] : x;
            XsFn f = xs_warp_compose(xs_elem(x, eb), lane);
            f.c0 = __shfl_sync(0xffffffffu, f.c0, 0); f.c1 = __shfl_sync(0xffffffffu, f.c1, 0);
            done = xs_apply(&sum, f, eb);
          }
          if (!done) { xs_seq_tile(rows, s, p0 + t * 32, n, k, lane, &sum); slow_rows += 32; }
        }
      }
    }
    __syncwarp();
  }
  if (lane == 0) { out[k] = sum; if (stats) stats[k] = slow_rows; }
}
