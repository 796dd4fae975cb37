// tml_exact_sum.h -- the reference's sequential double-precision sums, bit for bit, in parallel.
//
// Why.  The reference accumulates each rank's window with a plain Python loop, ``s += x`` one row
// after another (reporting/sections/step_time/model.py:262-268, alignment.py:59-75).  Its
// rank-level tie-breaks (closest_rank_to_median: model.py:77-105; argmax: adapters.py:142-197)
// are decided on the LAST ULP of those sums whenever the rank count is even, so "the same sum up
// to rounding" is not enough for bit-exact rank indices: the rounding history itself has to be
// reproduced.  A dependent chain of W double adds is 10 ns per row on one thread -- 40 ms at
// W = 4*10^6 -- so round 1 only did it below 2^17 rows.
//
// How.  All addends are >= 0, so the running sum s is monotone.  While s stays inside one binade
// [2^e, 2^(e+1)) its ulp u = 2^(e-52) is fixed, s = S*u with an integer S in [2^52, 2^53), and
//     RN(s + x) = (S + rne(x / u)) * u,     rne = round to nearest integer, ties to EVEN RESULT.
// x/u = q + f (integer q, 0 <= f < 1).  For f != 1/2 the increment is a constant (q or q+1); for
// the tie f == 1/2 it depends only on the parity of S + q.  So every add is a map
//     S -> S + c[S & 1]                              (two integers c[0], c[1])
// and such maps are closed under composition:
//     (g o f)[p] = f[p] + g[(p + f[p]) & 1].
// That is an associative monoid: a chunk of rows collapses to one (c0, c1) pair by an ordered
// tree reduction, chunks collapse to groups, and a single short walk applies them to S -- integer
// adds only.  Binade crossings (at most ~60 per sum: the exponent only grows) and the start-up
// from s = 0 fall back to real dependent DADDs on just the 32-row tile that contains them.  The
// exponent a chunk was composed under comes from an approximate prefix sum and is VERIFIED against
// the true running sum when the chunk is applied; a wrong guess costs time, never correctness.
//
// The memory columns need none of this: they are integer byte counts and the reference averages
// them with CPython >= 3.12's compensated sum() (step_memory/model.py:224-246), i.e. the exactly
// rounded integer sum -- the kernels add them as u64.
//
// This header is shared by the CUDA kernels (tml_engine.cu) and by a host emulation that runs
// the same plan / compose / walk steps serially (tml_xs_host_sum, tests only: the CPU suite
// fuzzes the arithmetic here, without a GPU, against a plain sequential loop).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define XS_HD __host__ __device__ __forceinline__
#else
#define XS_HD inline
#endif

#define XS_CHUNK 256           // rows per composed chunk (one CTA pass)
#define XS_GROUP 32            // chunks per composed group (one warp pass)
#define XS_PLAN_UNSAFE (-1)    // exponent not constant across the chunk (or unknown)
#define XS_PLAN_ZERO (-2)      // every addend of the chunk is +0.0: identity

struct XsFn {
  unsigned long long c0, c1;   // S -> S + (S & 1 ? c1 : c0); c0 == ~0ull: invalid (crossing)
};

XS_HD XsFn xs_identity() { XsFn f; f.c0 = 0ull; f.c1 = 0ull; return f; }
XS_HD XsFn xs_invalid() { XsFn f; f.c0 = ~0ull; f.c1 = ~0ull; return f; }
XS_HD bool xs_valid(const XsFn& f) { return f.c0 != ~0ull; }

XS_HD unsigned long long xs_bits(double x) {
  unsigned long long b;
#if defined(__CUDA_ARCH__)
  b = (unsigned long long)__double_as_longlong(x);
#else
  memcpy(&b, &x, 8);
#endif
  return b;
}
XS_HD double xs_from_bits(unsigned long long b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  double x;
  memcpy(&x, &b, 8);
  return x;
#endif
}

// biased exponent of a positive normal double (0: zero / subnormal)
XS_HD int xs_exp(double s) { return (int)((xs_bits(s) >> 52) & 0x7ffull); }

// The map of "s += x" while s has biased exponent eb (1..2046).  x >= 0, finite.
XS_HD XsFn xs_elem(double x, int eb) {
  const unsigned long long b = xs_bits(x);
  if ((b << 1) == 0ull) return xs_identity();               // +-0
  int ex = (int)((b >> 52) & 0x7ffull);
  unsigned long long m = b & 0x000fffffffffffffull;
  if (ex == 0) ex = 1; else m |= 0x0010000000000000ull;     // subnormal: no implicit bit
  if ((b >> 63) || ex == 0x7ff || ex > eb) return xs_invalid();  // negative / inf / nan / x >= 2^(e+1)
  const int sh = eb - ex;                                   // x / u = m * 2^-sh
  if (sh == 0) { XsFn f; f.c0 = m; f.c1 = m; return f; }
  if (sh >= 54) return xs_identity();                       // x < u / 2
  const unsigned long long q = (sh <= 52) ? (m >> sh) : 0ull;
  const unsigned long long rem = (sh <= 52) ? (m & ((1ull << sh) - 1ull)) : m;
  const unsigned long long half = 1ull << (sh - 1);
  XsFn f;
  if (rem > half) { f.c0 = q + 1ull; f.c1 = q + 1ull; }
  else if (rem < half) { f.c0 = q; f.c1 = q; }
  else { f.c0 = q + (q & 1ull); f.c1 = q + ((q + 1ull) & 1ull); }  // tie: the RESULT S + c must be even
  return f;
}

// f first, then g
XS_HD XsFn xs_compose(const XsFn& f, const XsFn& g) {
  if (!xs_valid(f) || !xs_valid(g)) return xs_invalid();
  XsFn h;
  h.c0 = f.c0 + (((0ull + f.c0) & 1ull) ? g.c1 : g.c0);
  h.c1 = f.c1 + (((1ull + f.c1) & 1ull) ? g.c1 : g.c0);
  // a map that can leave the binade is of no use to anybody: keep the integers small
  if (h.c0 >= (1ull << 53) || h.c1 >= (1ull << 53)) return xs_invalid();
  return h;
}

// Hot-loop variants: no validity sentinel, no range check.  The caller tracks "some element was
// invalid" separately (xs_elem_raw's return value) and bounds the magnitudes itself: a map of one
// element is < 2^53, so 2^10 composed elements stay < 2^63.  Non-tie maps (c0 == c1, all but
// ~2^-20 of the elements) compose by a plain add.
XS_HD bool xs_elem_raw(double x, int eb, XsFn* out) {
  const XsFn f = xs_elem(x, eb);
  if (!xs_valid(f)) { *out = xs_identity(); return false; }
  *out = f;
  return true;
}
XS_HD XsFn xs_compose_raw(const XsFn& f, const XsFn& g) {
  XsFn h;
  if (g.c0 == g.c1) { h.c0 = f.c0 + g.c0; h.c1 = f.c1 + g.c0; return h; }
  h.c0 = f.c0 + ((f.c0 & 1ull) ? g.c1 : g.c0);
  h.c1 = f.c1 + (((1ull + f.c1) & 1ull) ? g.c1 : g.c0);
  return h;
}
// The same map through the FPU (hot loops): t = x / u is an exact scaling by a power of two,
// floor(t) and t - floor(t) are exact for t < 2^53, so q, "above / below / at one half" come out of
// one multiply, one floor, one subtract and two compares instead of a page of 64-bit shifts.
// ``scale`` = 2^(1075 - eb) from xs_scale (0.0: eb too small for a normal scale -> integer path).
XS_HD double xs_scale(int eb) {
  const int f = 2098 - eb;  // biased exponent of 2^(1075 - eb)
  return (f >= 1 && f <= 2046) ? xs_from_bits((unsigned long long)f << 52) : 0.0;
}
XS_HD bool xs_elem_fp(double x, int eb, double scale, XsFn* out) {
  if (scale == 0.0) return xs_elem_raw(x, eb, out);
  const double t = x * scale;
  if (!(t >= 0.0 && t < 9007199254740992.0)) { *out = xs_identity(); return false; }  // < 0, nan, >= 2^53
#if defined(__CUDA_ARCH__)
  const double fl = floor(t);
#else
  const double fl = __builtin_floor(t);
#endif
  const unsigned long long q = (unsigned long long)fl;
  const double fr = t - fl;
  const unsigned long long up = fr > 0.5 ? 1ull : 0ull;
  const unsigned long long tie = fr == 0.5 ? 1ull : 0ull;
  out->c0 = q + (up | (tie & q));           // tie: S + c must be even; S even -> c even
  out->c1 = q + (up | (tie & (q + 1ull)));  //                          S odd  -> c odd
  return true;
}

// Branch-free forms for the compose kernel's inner loop (scale != 0 guaranteed by the caller):
// *bad accumulates "this element does not fit under the exponent" (negative, nan, >= 2^(e+1));
// the garbage map such an element yields is discarded when the chunk map is sealed.
XS_HD XsFn xs_elem_fp_nb(double x, double scale, bool* bad) {
  const double t = x * scale;
  *bad = *bad || !(t >= 0.0 && t < 9007199254740992.0);
#if defined(__CUDA_ARCH__)
  const double fl = floor(t);
#else
  const double fl = __builtin_floor(t);
#endif
  const unsigned long long q = (unsigned long long)fl;
  const double fr = t - fl;
  const unsigned long long up = fr > 0.5 ? 1ull : 0ull;
  const unsigned long long tie = fr == 0.5 ? 1ull : 0ull;
  XsFn f;
  f.c0 = q + (up | (tie & q));
  f.c1 = q + (up | (tie & (q + 1ull)));
  return f;
}
XS_HD XsFn xs_compose_nb(const XsFn& f, const XsFn& g) {
  XsFn h;
  h.c0 = f.c0 + ((f.c0 & 1ull) ? g.c1 : g.c0);
  h.c1 = f.c1 + ((f.c1 & 1ull) ? g.c0 : g.c1);
  return h;
}

// close a raw composition: anything that could leave the binade (or saw an invalid element) is invalid
XS_HD XsFn xs_seal(const XsFn& f, bool ok) {
  if (!ok || f.c0 >= (1ull << 53) || f.c1 >= (1ull << 53)) return xs_invalid();
  return f;
}

// Apply f (composed under biased exponent eb) to s.  False -- s untouched -- if s is not in that
// binade, f is invalid, or the result would reach 2^(e+1) (then some add inside crossed, or would
// round at a coarser ulp: the caller redoes those rows with real adds).
XS_HD bool xs_apply(double* s, const XsFn& f, int eb) {
  if (!xs_valid(f)) return false;
  const unsigned long long b = xs_bits(*s);
  if ((int)((b >> 52) & 0x7ffull) != eb || (b >> 63) || eb <= 0 || eb >= 0x7ff) return false;
  const unsigned long long S = (b & 0x000fffffffffffffull) | 0x0010000000000000ull;
  const unsigned long long S2 = S + ((S & 1ull) ? f.c1 : f.c0);
  if (S2 >= (1ull << 53)) return false;
  *s = xs_from_bits(((unsigned long long)eb << 52) | (S2 & 0x000fffffffffffffull));
  return true;
}

// Walk state as integers: (eb, S) with s = S * 2^(eb - 1075); eb == 0: s is zero / subnormal.
XS_HD void xs_unpack(double s, int* eb, unsigned long long* S) {
  const unsigned long long b = xs_bits(s);
  *eb = (int)((b >> 52) & 0x7ffull);
  *S = (b & 0x000fffffffffffffull) | 0x0010000000000000ull;
  if (b >> 63) *eb = 0;  // negative: never applies
}
XS_HD double xs_pack(int eb, unsigned long long S) {
  return xs_from_bits(((unsigned long long)eb << 52) | (S & 0x000fffffffffffffull));
}
XS_HD bool xs_apply_s(int eb, unsigned long long* S, const XsFn& f, int E) {
  if (E != eb || E <= 0 || E >= 0x7ff || f.c0 == ~0ull) return false;
  const unsigned long long S2 = *S + ((*S & 1ull) ? f.c1 : f.c0);
  if (S2 >> 53) return false;
  *S = S2;
  return true;
}

// Plan: may rows [start, end] of approximate running sums all share one binade?  ``lo`` is the
// approximate sum before the chunk, ``hi`` after it (both from tree sums, relative error far
// below the 1e-6 margin used here).  Returns the biased exponent, XS_PLAN_ZERO or XS_PLAN_UNSAFE.
XS_HD int xs_plan(double lo, double hi) {
  if (hi == 0.0) return XS_PLAN_ZERO;
  if (!(lo > 0.0)) return XS_PLAN_UNSAFE;
  const double a = lo * (1.0 - 1.0e-6), b = hi * (1.0 + 1.0e-6);
  const int ea = xs_exp(a), ebb = xs_exp(b);
  if (ea != ebb || ea <= 0 || ea >= 0x7fe) return XS_PLAN_UNSAFE;
  return ea;
}
