/* crmath.h -- correctly-rounded FP64 leaf functions for the RRT hot path, identical op-for-op on
 * the GPU (nvcc, device code) and on the CPU (gcc, used by the oracle's "cr" mode).
 *
 * Why: the reference (pure Python, FP64) takes thousands of decisions per query whose margin is
 * a few ulp -- `floor(d / path_resolution)` and `d <= path_resolution` in steer (rrt_04:1099-1107)
 * hit exact multiples of the resolution because nodes are built by accumulating resolution-sized
 * steps.  A tree only stays bit-identical if hypot/atan2/cos/sin are reproducible to the last bit.
 * libm results are platform specific (glibc's sin/cos/atan2 are not correctly rounded in ~0.1 % of
 * calls and depend on the CPU's FMA ifunc variant), so this path defines its arithmetic as the
 * IEEE-754 ideal: every leaf function returns the correctly rounded result.
 *   - crm_hypot    : CPython's math.hypot algorithm (Modules/mathmodule.c vector_norm, n = 2),
 *                    bit-identical to the reference's `math.hypot` (rrt_04:1235).
 *   - crm_atan2_sincos : theta = RN(atan2(dy, dx)), RN(sin(theta)), RN(cos(theta)) -- the three
 *                    calls of rrt_04:1236 and :1100-1101 -- from one double-double sin/cos
 *                    evaluation plus a Newton correction.
 *   - crm_sin / crm_cos / crm_atan2 : stand-alone versions (Dubins path, tests).
 * Double-double accuracy is ~2^-98 relative, so a result is mis-rounded only when the true value
 * lies within 2^-45 ulp of a rounding midpoint.  Checked against mpmath in tests/test_crmath.py.
 *
 * Requirements: no FMA contraction (nvcc -fmad=false, gcc -ffp-contract=off); fma() is used only
 * where written.  Domain: finite inputs, |x| < 2^20 * pi/2 for sin/cos.
 */
#ifndef RRTK_CRMATH_H
#define RRTK_CRMATH_H

#ifdef __CUDACC__
#define CRM_FN static __device__ __forceinline__
#define CRM_NOINLINE static __device__ __noinline__
#define CRM_CONST static __device__ __constant__ const  /* uniform index: constant cache */
#define CRM_TABLE static __device__ const               /* per-lane index: global memory / L1 */
#define CRM_ROLLED _Pragma("unroll 1")
/* double-double primitives: inlined (out-of-line calls measured 2 % slower on B200) */
#define CRM_DDOP static __device__ __forceinline__
CRM_FN long long crm_d2ll(double x) { return __double_as_longlong(x); }
CRM_FN double crm_ll2d(long long v) { return __longlong_as_double(v); }
#else
#include <float.h>
#include <math.h>
#include <string.h>
#define CRM_FN static inline
#define CRM_NOINLINE static
#define CRM_CONST static const
#define CRM_TABLE static const
#define CRM_ROLLED
#define CRM_DDOP static inline
CRM_FN long long crm_d2ll(double x) { long long v; memcpy(&v, &x, 8); return v; }
CRM_FN double crm_ll2d(long long v) { double x; memcpy(&x, &v, 8); return x; }
#endif

#include "crmath_consts.h"

typedef struct { double hi, lo; } crm_dd;

CRM_FN crm_dd crm_mk(double hi, double lo) { crm_dd r; r.hi = hi; r.lo = lo; return r; }

/* error-free transforms */
CRM_FN crm_dd crm_two_sum(double a, double b) {
    double s = a + b;
    double bb = s - a;
    double e = (a - (s - bb)) + (b - bb);
    return crm_mk(s, e);
}
CRM_FN crm_dd crm_fast_two_sum(double a, double b) { /* |a| >= |b| or a == 0 */
    double s = a + b;
    double e = b - (s - a);
    return crm_mk(s, e);
}
CRM_FN crm_dd crm_two_prod(double a, double b) {
    double p = a * b;
    double e = fma(a, b, -p);
    return crm_mk(p, e);
}

/* double-double arithmetic (Dekker / Hida-Li-Bailey "accurate" variants) */
CRM_DDOP crm_dd crm_add(crm_dd a, crm_dd b) {
    crm_dd s = crm_two_sum(a.hi, b.hi);
    crm_dd t = crm_two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = crm_fast_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return crm_fast_two_sum(s.hi, s.lo);
}
CRM_DDOP crm_dd crm_add_d(crm_dd a, double b) {
    crm_dd s = crm_two_sum(a.hi, b);
    s.lo += a.lo;
    return crm_fast_two_sum(s.hi, s.lo);
}
CRM_FN crm_dd crm_neg(crm_dd a) { return crm_mk(-a.hi, -a.lo); }
CRM_FN crm_dd crm_sub(crm_dd a, crm_dd b) { return crm_add(a, crm_neg(b)); }
CRM_DDOP crm_dd crm_mul(crm_dd a, crm_dd b) {
    crm_dd p = crm_two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return crm_fast_two_sum(p.hi, p.lo);
}
CRM_DDOP crm_dd crm_mul_d(crm_dd a, double b) {
    crm_dd p = crm_two_prod(a.hi, b);
    p.lo += a.lo * b;
    return crm_fast_two_sum(p.hi, p.lo);
}
CRM_DDOP crm_dd crm_div(crm_dd a, crm_dd b) {
    /* quotient digits from one reciprocal; each digit is corrected by the exact residual */
    double inv = 1.0 / b.hi;
    double q1 = a.hi * inv;
    crm_dd r = crm_sub(a, crm_mul_d(b, q1));
    double q2 = r.hi * inv;
    r = crm_sub(r, crm_mul_d(b, q2));
    double q3 = r.hi * inv;
    crm_dd q = crm_fast_two_sum(q1, q2);
    return crm_add_d(q, q3);
}

/* ---- math.hypot of CPython (correctly rounded in practice; bit-identical to the reference) ---- */
CRM_NOINLINE double crm_hypot(double a, double b) {
    double v0 = fabs(a), v1 = fabs(b);
    double mx = v0 > v1 ? v0 : v1;
    if (mx == 0.0) return mx;
    /* frexp(mx) -> max_e, scale = 2^-max_e, and the final h / scale = h * 2^max_e, by exponent-field
     * arithmetic when mx is a normal number away from the range ends (bit-identical to frexp/ldexp) */
    int be = (int)((crm_d2ll(mx) >> 52) & 0x7ff);
    double scale, unscale, post = 1.0;
    int fast = be >= 2 && be <= 2040;
    if (fast) {
        scale = crm_ll2d((long long)(2045 - be) << 52);
        unscale = crm_ll2d((long long)(be + 1) << 52);
    } else {
        int max_e;
        (void)frexp(mx, &max_e);
        if (max_e < -1023) { /* subnormal range: rescale first, as CPython does */
            const double dmin = 2.2250738585072014e-308;
            v0 /= dmin; v1 /= dmin; mx /= dmin; post = dmin;
            (void)frexp(mx, &max_e);
        }
        scale = ldexp(1.0, -max_e);
        unscale = 0.0;
    }
    double csum = 1.0, frac1 = 0.0, frac2 = 0.0;
    double x = v0 * scale;
    crm_dd pr = crm_two_prod(x, x);
    crm_dd sm = crm_fast_two_sum(csum, pr.hi);
    csum = sm.hi; frac1 += pr.lo; frac2 += sm.lo;
    x = v1 * scale;
    pr = crm_two_prod(x, x);
    sm = crm_fast_two_sum(csum, pr.hi);
    csum = sm.hi; frac1 += pr.lo; frac2 += sm.lo;
    double h = sqrt(csum - 1.0 + (frac1 + frac2));
    pr = crm_two_prod(-h, h);
    sm = crm_fast_two_sum(csum, pr.hi);
    csum = sm.hi; frac1 += pr.lo; frac2 += sm.lo;
    x = csum - 1.0 + (frac1 + frac2);
    h += x / (2.0 * h);
    return fast ? h * unscale : post * (h / scale);
}

/* ---- double-double sin and cos of a double ---- */
CRM_NOINLINE void crm_sincos_dd(double x, crm_dd *s_out, crm_dd *c_out) {
    /* Cody-Waite reduction with pi/2 = P1 + P2 + P3 (33 + 33 + 106 bits): k*P1, k*P2 exact */
    double kf = rint(x * CRM_2OPI);
    crm_dd r;
    if (kf == 0.0) {
        r = crm_mk(x, 0.0);
    } else {
        double t = x - kf * CRM_PIO2_1;              /* exact (Sterbenz) */
        r = crm_two_sum(t, -(kf * CRM_PIO2_2));      /* exact product, error-free sum */
        crm_dd p3 = crm_two_prod(kf, CRM_PIO2_3H);
        p3.lo += kf * CRM_PIO2_3L;
        r = crm_sub(r, p3);
    }
    /* second reduction: |r| = i/128 + h, |h| <= 2^-8; sin/cos(i/128) from the table */
    int neg = r.hi < 0.0;
    if (neg) r = crm_neg(r);
    int ti = (int)rint(r.hi * 128.0);
    double xi = (double)ti * 0.0078125;
    crm_dd h = crm_fast_two_sum(r.hi - xi, r.lo); /* r.hi - xi is exact (Sterbenz) */
    crm_dd h2 = crm_mul(h, h);
    /* sin(h) = h * (1 + h2 * S(h2)),  cos(h) = 1 + h2 * C(h2): Taylor / Horner in double-double */
    /* |h| <= 2^-8: the terms from h^7/7! (sin) and h^6/6! (cos) on are below 2^-57 of the result, so
     * their polynomial runs in plain double (error < 2^-109); the two leading coefficients of each
     * series are applied in double-double */
    double z = h2.hi;
    double ts = crm_sin_c[CRM_NSIN - 1][0], tc = crm_cos_c[CRM_NCOS - 1][0];
    tc = fma(tc, z, crm_cos_c[CRM_NCOS - 2][0]);
    CRM_ROLLED
    for (int k = CRM_NSIN - 2; k >= 2; k--) { /* CRM_NCOS == CRM_NSIN + 1 */
        ts = fma(ts, z, crm_sin_c[k][0]);
        tc = fma(tc, z, crm_cos_c[k][0]);
    }
    crm_dd ps = crm_add(crm_mul_d(h2, ts), crm_mk(crm_sin_c[1][0], crm_sin_c[1][1]));
    crm_dd pc = crm_add(crm_mul_d(h2, tc), crm_mk(crm_cos_c[1][0], crm_cos_c[1][1]));
    ps = crm_add(crm_mul(ps, h2), crm_mk(crm_sin_c[0][0], crm_sin_c[0][1]));
    pc = crm_add(crm_mul(pc, h2), crm_mk(crm_cos_c[0][0], crm_cos_c[0][1]));
    crm_dd sh = crm_add(h, crm_mul(h, crm_mul(ps, h2)));
    crm_dd ch = crm_add_d(crm_mul(pc, h2), 1.0);
    crm_dd sr, cr;
    if (ti == 0) {
        sr = sh; cr = ch;
    } else { /* angle addition with the table entry */
        crm_dd si = crm_mk(crm_sincos_tab[ti][0], crm_sincos_tab[ti][1]);
        crm_dd ci = crm_mk(crm_sincos_tab[ti][2], crm_sincos_tab[ti][3]);
        sr = crm_add(crm_mul(si, ch), crm_mul(ci, sh));
        cr = crm_sub(crm_mul(ci, ch), crm_mul(si, sh));
    }
    if (neg) sr = crm_neg(sr);
    int q = ((int)kf) & 3;
    if (q == 0) { *s_out = sr; *c_out = cr; }
    else if (q == 1) { *s_out = cr; *c_out = crm_neg(sr); }
    else if (q == 2) { *s_out = crm_neg(sr); *c_out = crm_neg(cr); }
    else { *s_out = crm_neg(cr); *c_out = sr; }
}

CRM_FN double crm_sin(double x) {
    if (x == 0.0) return x;
    crm_dd s, c;
    crm_sincos_dd(x, &s, &c);
    return s.hi;
}
CRM_FN double crm_cos(double x) {
    crm_dd s, c;
    crm_sincos_dd(x, &s, &c);
    return c.hi;
}

/* crude atan2 (|error| < 1e-10), plain double, same op sequence everywhere */
CRM_FN double crm_atan2_guess(double y, double x) {
    double ax = fabs(x), ay = fabs(y);
    double mx = ax > ay ? ax : ay, mn = ax > ay ? ay : ax;
    double t = mn / mx;
    double off = 0.0;
    if (t > 0.41421356237309503) { t = (t - 1.0) / (t + 1.0); off = CRM_PIO4_H; }
    double z = t * t;
    double p = crm_atan_c[CRM_NATAN - 1];
    CRM_ROLLED
    for (int k = CRM_NATAN - 2; k >= 0; k--) p = p * z + crm_atan_c[k];
    double a = off + (t + t * (z * p));
    if (ay > ax) a = CRM_PIO2_H - a;
    if (x < 0.0) a = CRM_PI_H - a;
    return y < 0.0 ? -a : a;
}

/* theta = RN(atan2(y, x)); *s = RN(sin(theta)); *c = RN(cos(theta)).
 * Newton: theta = t0 + atan(u), u = (y cos t0 - x sin t0) / (x cos t0 + y sin t0), all double-double;
 * then sin/cos(theta) by rotating (sin t0, cos t0) through eps = theta - t0 (exact, |eps| < 1e-9). */
CRM_FN double crm_atan2_sincos(double y, double x, double *s, double *c) {
    if (y == 0.0) { /* includes (0, 0): atan2 = +-0 or +-pi */
        int neg = (x < 0.0) || (x == 0.0 && signbit(x));
        if (!neg) { *s = y; *c = 1.0; return y; }
        double th = copysign(CRM_PI_H, y);
        *s = copysign(CRM_PI_L, y); /* RN(sin(RN(pi))) = RN(pi - RN(pi)) */
        *c = -1.0;
        return th;
    }
    double t0 = crm_atan2_guess(y, x);
    crm_dd s0, c0;
    crm_sincos_dd(t0, &s0, &c0);
    crm_dd num = crm_sub(crm_mul_d(c0, y), crm_mul_d(s0, x));
    crm_dd den = crm_add(crm_mul_d(c0, x), crm_mul_d(s0, y));
    crm_dd u = crm_div(num, den);
    crm_dd u3 = crm_mul(crm_mul(u, u), u);
    crm_dd del = crm_sub(u, crm_mul(u3, crm_mk(CRM_THIRD_H, CRM_THIRD_L)));
    crm_dd th = crm_add_d(del, t0);
    double theta = th.hi;
    double eps = theta - t0; /* exact */
    /* sin(eps) = eps - eps^3/6, cos(eps) = 1 - eps^2/2 (next terms < 1e-37) */
    crm_dd e2 = crm_two_prod(eps, eps);
    crm_dd se = crm_sub(crm_mk(eps, 0.0), crm_mul_d(crm_mul_d(e2, eps), 1.0 / 6.0));
    crm_dd ce = crm_add_d(crm_mul_d(e2, -0.5), 1.0);
    crm_dd st = crm_add(crm_mul(s0, ce), crm_mul(c0, se));
    crm_dd ct = crm_sub(crm_mul(c0, ce), crm_mul(s0, se));
    *s = st.hi;
    *c = ct.hi;
    return theta;
}

CRM_FN double crm_atan2(double y, double x) {
    double s, c;
    return crm_atan2_sincos(y, x, &s, &c);
}

/* RN(acos(x)) for |x| <= 1 (NaN outside): acos(x) = atan2(sqrt((1 - x)(1 + x)), x) with the square root
 * carried in double-double and the same Newton step as crm_atan2_sincos. */
CRM_FN double crm_acos(double x) {
    if (!(fabs(x) <= 1.0)) return x - x == 0.0 ? (x - x) / (x - x) : x + x; /* NaN */
    if (x == 1.0) return 0.0;
    if (x == -1.0) return CRM_PI_H;
    crm_dd a = crm_two_sum(1.0, -x), b = crm_two_sum(1.0, x); /* exact */
    crm_dd p = crm_mul(a, b);
    double s0 = sqrt(p.hi);
    crm_dd e = crm_sub(p, crm_two_prod(s0, s0));
    crm_dd y = crm_fast_two_sum(s0, e.hi / (2.0 * s0)); /* sqrt(p) to ~2^-104 */
    double t0 = crm_atan2_guess(y.hi, x);
    crm_dd s0d, c0d;
    crm_sincos_dd(t0, &s0d, &c0d);
    crm_dd num = crm_sub(crm_mul(c0d, y), crm_mul_d(s0d, x));
    crm_dd den = crm_add(crm_mul_d(c0d, x), crm_mul(s0d, y));
    crm_dd u = crm_div(num, den);
    crm_dd u3 = crm_mul(crm_mul(u, u), u);
    crm_dd del = crm_sub(u, crm_mul(u3, crm_mk(CRM_THIRD_H, CRM_THIRD_L)));
    crm_dd th = crm_add_d(del, t0);
    return th.hi;
}

#endif /* RRTK_CRMATH_H */
