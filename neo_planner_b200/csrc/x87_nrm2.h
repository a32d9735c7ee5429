// x87_nrm2.h -- the Euclidean norm exactly as OpenBLAS's x86-64 dnrm2 kernel returns it (scipy's L-BFGS-B calls it
// for |d| and |y|, see lbfgsb_tile.cuh): every square and the running sum are kept in the x87's 80-bit extended format
// (64-bit significand, round to nearest even), the square root is taken in that format (fsqrt is correctly rounded) and
// only the final value is rounded to double. The device has no extended format, so the three operations are carried
// out on (64-bit significand, exponent) pairs with integer arithmetic. Host-compilable: tests/test_x87_nrm2.py checks
// it against the real kernel (scipy.linalg.blas.dnrm2) and against `long double` arithmetic on the build machine.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define X87_FN __host__ __device__ __forceinline__
#define X87_NOINLINE __host__ __device__ __noinline__
#else
#define X87_FN static inline
#define X87_NOINLINE static __attribute__((noinline))
#endif

namespace neo {

struct Ext {            // value = m * 2^e, m = 0 or 2^63 <= m < 2^64
    uint64_t m;
    int e;
};

X87_FN int x87_clz64(uint64_t v)
{
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}

// round a 128-bit magnitude (hi, lo) * 2^e with hi's top bit set to a 64-bit significand; sticky: bits already lost below lo
X87_FN Ext x87_round(uint64_t hi, uint64_t lo, int e, bool sticky)
{
    const uint64_t half = 0x8000000000000000ull;
    const bool up = lo > half || (lo == half && (sticky || (hi & 1ull)));
    Ext r;
    r.m = hi; r.e = e + 64;
    if (up) {
        r.m = hi + 1ull;
        if (r.m == 0ull) { r.m = half; r.e += 1; }
    }
    return r;
}

X87_FN Ext x87_from_double(double d)       // |d|, finite
{
    union { double d; uint64_t u; } c;
    c.d = d;
    const uint64_t frac = c.u & 0x000fffffffffffffull;
    const int ex = (int)((c.u >> 52) & 0x7ff);
    Ext r;
    if (ex == 0) {
        if (frac == 0) { r.m = 0; r.e = 0; return r; }
        const int s = x87_clz64(frac);
        r.m = frac << s; r.e = -1074 - s;
        return r;
    }
    r.m = (frac | 0x0010000000000000ull) << 11;
    r.e = ex - 1075 - 11;
    return r;
}

X87_FN Ext x87_sqr(Ext a)
{
    if (a.m == 0) return a;
    const unsigned __int128 p = (unsigned __int128)a.m * a.m;          // top bit at 127 or 126
    uint64_t hi = (uint64_t)(p >> 64), lo = (uint64_t)p;
    int e = 2 * a.e;
    if (!(hi >> 63)) { hi = (hi << 1) | (lo >> 63); lo <<= 1; e -= 1; }
    return x87_round(hi, lo, e, false);
}

X87_FN Ext x87_add(Ext a, Ext b)           // both >= 0
{
    if (a.m == 0) return b;
    if (b.m == 0) return a;
    if (a.e < b.e) { const Ext t = a; a = b; b = t; }
    const int dlt = a.e - b.e;
    if (dlt > 66) return a;
    // 128-bit lanes with one bit of headroom: A = a.m << 63, B = (b.m << 63) >> dlt
    const unsigned __int128 A = (unsigned __int128)a.m << 63;
    unsigned __int128 B = (unsigned __int128)b.m << 63;
    bool sticky = false;
    if (dlt > 0) {
        if (dlt < 128) { sticky = (B & (((unsigned __int128)1 << dlt) - 1)) != 0; B >>= dlt; }
        else { sticky = true; B = 0; }
    }
    const unsigned __int128 S = A + B;                                  // top bit at 127 or 126
    if (S >> 127) return x87_round((uint64_t)(S >> 64), (uint64_t)S, a.e - 63, sticky);
    return x87_round((uint64_t)(S >> 63), (uint64_t)(S << 1), a.e - 64, sticky);
}

X87_FN Ext x87_sqrt(Ext a)
{
    if (a.m == 0) return a;
    // N = m * 2^64 (even exponent) or m * 2^63 (odd): sqrt(N) has its top bit at 63
    int e = a.e;
    unsigned __int128 N;
    if (e & 1) { N = (unsigned __int128)a.m << 63; e -= 63; }
    else { N = (unsigned __int128)a.m << 64; e -= 64; }
    // integer square root of a 128-bit number: double estimate, two Newton steps, exact fix-up
    double est;
    {
        const double hi = (double)(uint64_t)(N >> 64), lo = (double)(uint64_t)N;
        est = hi * 18446744073709551616.0 + lo;
    }
#if defined(__CUDA_ARCH__)
    double rs = sqrt(est);
#else
    double rs = __builtin_sqrt(est);
#endif
    uint64_t r = rs >= 18446744073709551615.0 ? 0xffffffffffffffffull : (uint64_t)rs;
    {   // one Newton step with the residual N - r^2 formed exactly (|residual| < 2^78) and divided in floating point
        const unsigned __int128 sq = (unsigned __int128)r * r;
        const bool neg = sq > N;
        const unsigned __int128 mag = neg ? sq - N : N - sq;
        const double magd = (double)(uint64_t)(mag >> 64) * 18446744073709551616.0 + (double)(uint64_t)mag;
        const double stepd = magd / (2.0 * (double)r);
        const uint64_t step = (uint64_t)(stepd + 0.5);
        if (neg) r -= step;
        else r = (0xffffffffffffffffull - r < step) ? 0xffffffffffffffffull : r + step;
    }
    while ((unsigned __int128)r * r > N) r--;
    while (r != 0xffffffffffffffffull && (unsigned __int128)(r + 1) * (r + 1) <= N) r++;
    const unsigned __int128 rem = N - (unsigned __int128)r * r;        // N > (r + 1/2)^2  <=>  rem > r
    Ext o;
    o.m = r; o.e = e / 2;
    if (rem > (unsigned __int128)r) {
        o.m = r + 1ull;
        if (o.m == 0ull) { o.m = 0x8000000000000000ull; o.e += 1; }
    }
    return o;
}

X87_FN double x87_to_double(Ext a)
{
    if (a.m == 0) return 0.0;
    uint64_t m = a.m >> 11;
    const uint64_t rem = a.m & 0x7ffull;
    int e = a.e + 11;
    if (rem > 0x400ull || (rem == 0x400ull && (m & 1ull))) {
        m += 1ull;
        if (m >> 53) { m >>= 1; e += 1; }
    }
    // m in [2^52, 2^53): value m * 2^e (normal range assumed: norms of finite optimizer vectors)
    union { double d; uint64_t u; } c;
    const int ex = e + 1075;
    if (ex <= 0 || ex >= 0x7ff) {
#if defined(__CUDA_ARCH__)
        return ldexp((double)m, e);
#else
        return __builtin_ldexp((double)m, e);
#endif
    }
    c.u = ((uint64_t)ex << 52) | (m & 0x000fffffffffffffull);
    return c.d;
}

// dnrm2 of v[0..n), elements taken in index order -- the operations exactly as the x87 performs them
// The summation order is the kernel's (OpenBLAS nrm2.S, disassembled from scipy's bundled library): four accumulators,
// element i of the leading blocks of 8 goes to accumulator i mod 4, the remaining n mod 8 elements to accumulator 0,
// total = D + ((C + A) + B). For n < 8 that is the plain sequential sum.
X87_NOINLINE double x87_nrm2_exact(int n, const double *v)
{
    Ext acc[4];
    for (int k = 0; k < 4; k++) { acc[k].m = 0; acc[k].e = 0; }
    const int nb8 = n & ~7;
    for (int i = 0; i < nb8; i++) acc[i & 3] = x87_add(acc[i & 3], x87_sqr(x87_from_double(v[i])));
    for (int i = nb8; i < n; i++) acc[0] = x87_add(acc[0], x87_sqr(x87_from_double(v[i])));
    const Ext s = x87_add(acc[3], x87_add(x87_add(acc[2], acc[0]), acc[1]));
    return x87_to_double(x87_sqrt(s));
}

X87_FN double x87_fma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

// Same value, usually without the emulation. The x87 result differs from the exact norm by less than (n + 2) * 2^-64
// relative before its final rounding to double, i.e. by less than 2^-6 ulp of the result for n <= 60. The norm is
// therefore computed to ~100 bits in double-double arithmetic; whenever it lies further than 2^-6 ulp from every
// rounding boundary, the correctly rounded double IS the x87's answer. Only the remaining ~3 % of the calls (and
// denormal / huge magnitudes) go through x87_nrm2_exact.
X87_FN double x87_nrm2(int n, const double *v)
{
    double hi = 0.0, lo = 0.0;
    for (int i = 0; i < n; i++) {
        const double a = v[i];
        const double p = a * a, pe = x87_fma(a, a, -p);            // a^2 = p + pe exactly
        const double t = hi + p, bb = t - hi;                       // two-sum
        const double err = (hi - (t - bb)) + (p - bb);
        hi = t; lo += err + pe;
    }
    {   // renormalise
        const double t = hi + lo;
        lo = lo - (t - hi); hi = t;
    }
    union { double d; uint64_t u; } c;
    c.d = hi;
    const int ex = (int)((c.u >> 52) & 0x7ff);
    if (ex < 120 || ex > 1900) return x87_nrm2_exact(n, v);         // zero, denormal squares, overflow range
#if defined(__CUDA_ARCH__)
    const double r = __dsqrt_rn(hi);
#else
    const double r = __builtin_sqrt(hi);
#endif
    const double res = x87_fma(-r, r, hi) + lo;                      // hi - r^2 is exact for a correctly rounded root
    const double rlo = res / (2.0 * r);
    c.d = r;
    if ((c.u & 0x000fffffffffffffull) == 0) return x87_nrm2_exact(n, v);    // power of two: uneven neighbour spacing
    c.u &= 0x7ff0000000000000ull;
    const double ulp = c.d * 2.220446049250313e-16;                  // 2^(e - 52)
    const double q = rlo / ulp;
    const double aq = q < 0.0 ? -q : q;
    if (aq < 0.5 - 0.015625) return r;
    if (aq > 0.5 + 0.015625 && aq < 1.0) return q > 0.0 ? r + ulp : r - ulp;
    return x87_nrm2_exact(n, v);
}

}  // namespace neo
