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
/* double-double primitives: inlined by default; -DCRM_DDOP_OUTLINE makes them calls (smaller code: the planner
 * kernels are instruction-fetch sensitive, L1.5 I-cache = 32 KB) */
#ifdef CRM_DDOP_OUTLINE
#define CRM_DDOP static __device__ __noinline__
#else
#define CRM_DDOP static __device__ __forceinline__
#endif
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
/* vector_norm's core for n = 2 given the power-of-two scale */
CRM_FN double crm_hypot_core(double v0, double v1, double scale) {
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
    return h;
}
/* range ends (subnormal / near-overflow maxima): frexp / ldexp exactly as CPython does; kept out of line so the
 * common path stays small (the planner kernels are instruction-fetch sensitive) */
CRM_NOINLINE double crm_hypot_edge(double v0, double v1, double mx) {
    int max_e;
    double post = 1.0;
    (void)frexp(mx, &max_e);
    if (max_e < -1023) { /* subnormal range: rescale first, as CPython does */
        const double dmin = 2.2250738585072014e-308;
        v0 /= dmin; v1 /= dmin; mx /= dmin; post = dmin;
        (void)frexp(mx, &max_e);
    }
    const double scale = ldexp(1.0, -max_e);
    return post * (crm_hypot_core(v0, v1, scale) / scale);
}
CRM_NOINLINE double crm_hypot(double a, double b) {
    double v0 = fabs(a), v1 = fabs(b);
    double mx = v0 > v1 ? v0 : v1;
    if (mx == 0.0) return mx;
    /* frexp(mx) -> max_e, scale = 2^-max_e, and the final h / scale = h * 2^max_e, by exponent-field
     * arithmetic when mx is a normal number away from the range ends (bit-identical to frexp/ldexp) */
    int be = (int)((crm_d2ll(mx) >> 52) & 0x7ff);
    if (!(be >= 2 && be <= 2040)) return crm_hypot_edge(v0, v1, mx);
    const double scale = crm_ll2d((long long)(2045 - be) << 52);
    const double unscale = crm_ll2d((long long)(be + 1) << 52);
    return crm_hypot_core(v0, v1, scale) * unscale;
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

#ifdef CRM_FAST
/* ---- first phase of Ziv's strategy (GPU builds, -DCRM_FAST) ----
 * The double-double evaluations above cost ~350 dependent FP64 operations; the planners' steering code (Dubins /
 * Reeds-Shepp words and course points) is bound by exactly that latency.  These first phases compute the same value as
 * hi + lo with an error bound E (a few 2^-66, derived below) in ~60 operations and return RN(hi + lo) when hi + (lo - E)
 * and hi + (lo + E) round to the same double -- then that double IS the correctly rounded result, the one the
 * double-double path returns.  Otherwise (about 1 call in 100 .. 1000) the caller runs the double-double path.  The CPU
 * oracle is built without CRM_FAST, so every parity test also checks the bounds; tests/test_crmath.py runs both forms
 * over 10^7 arguments.
 * sin / cos: r = x - k pi/2 = (rh, rl) (error < 2^-97), |r| = i/128 + h, h = (h, hl), |h| <= 2^-8:
 *   sin h = h + sl,  sl = h z ps + hl        z = h^2, ps = -1/6 + z/120 - z^2/5040 + z^3/362880   (|sl| < 2^-26 |h|)
 *   cos h = 1 + cm,  cm = z pc - h hl        pc = -1/2 + z/24 - z^2/720 + z^3/40320                (|cm| < 2^-17)
 *   sin(i/128 + h) = S + [C h] + (Sl + S cm + C sl + Cl h),  cos(..) = C - [S h] + (Cl + C cm - S sl - Sl h)
 * with [.] an exact product (two_prod) and (S, Sl), (C, Cl) the table entry.  Rounding errors: every term inside (..) is
 * below 2^-16 and is formed by <= 3 roundings, their sum by 4 more: < 8 * 2^-69; dropped terms (Sl cm, hl z / 2, ...)
 * < 2^-69.  E = 2^-65 absolute for i >= 1 (results >= 2^-8.1), 2^-62 relative for i == 0. */
CRM_FN int crm_sincos_fast(double x, double *s_out, double *c_out) {
    double kf = rint(x * CRM_2OPI);
    double rh, rl;
    if (kf == 0.0) {
        rh = x; rl = 0.0;
    } else {
        double t = x - kf * CRM_PIO2_1;                       /* exact (Sterbenz) */
        crm_dd r = crm_two_sum(t, -(kf * CRM_PIO2_2));        /* exact product, error-free sum */
        double l = r.lo - (kf * CRM_PIO2_3H + kf * CRM_PIO2_3L);
        r = crm_two_sum(r.hi, l);
        rh = r.hi; rl = r.lo;
    }
    int neg = rh < 0.0;
    if (neg) { rh = -rh; rl = -rl; }
    /* x within |k| 2^-50 of k pi/2: the reduction above (error ~|k| 2^-121) is no longer good to 2^-64 RELATIVE to r */
    if (rh < fabs(kf) * 8.8817841970012523e-16) return 0;
    int ti = (int)rint(rh * 128.0);
    double xi = (double)ti * 0.0078125;
    crm_dd hh = crm_fast_two_sum(rh - xi, rl);                /* rh - xi exact (Sterbenz); |rh - xi| >= |rl| or == 0 */
    double h = hh.hi, hl = hh.lo;
    double z = h * h;
    double ps = fma(fma(fma(2.7557319223985893e-06, z, -1.9841269841269841e-04), z, 8.3333333333333332e-03), z, -1.6666666666666666e-01);
    double pc = fma(fma(fma(2.4801587301587302e-05, z, -1.3888888888888889e-03), z, 4.1666666666666664e-02), z, -0.5);
    double sl = h * (z * ps) + hl;
    double cm = z * pc - h * hl;
    double shi, slo, chi, clo, es, ec;
    if (ti == 0) {
        crm_dd a = crm_fast_two_sum(h, sl);
        shi = a.hi; slo = a.lo; es = fabs(a.hi) * 2.168404344971009e-19;   /* 2^-62 relative */
        a = crm_fast_two_sum(1.0, cm);
        chi = a.hi; clo = a.lo; ec = 2.710505431213761e-20;                /* 2^-65 */
    } else {
        double S = crm_sincos_tab[ti][0], Sl = crm_sincos_tab[ti][1], C = crm_sincos_tab[ti][2], Cl = crm_sincos_tab[ti][3];
        crm_dd p = crm_two_prod(C, h);
        double rest = p.lo + (Sl + (S * cm + (C * sl + Cl * h)));
        crm_dd a = crm_fast_two_sum(S, p.hi);                 /* S >= 2^-7.1 > |C h| */
        a = crm_fast_two_sum(a.hi, a.lo + rest);
        shi = a.hi; slo = a.lo;
        p = crm_two_prod(-S, h);
        rest = p.lo + (Cl + (C * cm - (S * sl + Sl * h)));
        a = crm_fast_two_sum(C, p.hi);
        a = crm_fast_two_sum(a.hi, a.lo + rest);
        chi = a.hi; clo = a.lo;
        es = ec = 2.710505431213761e-20;                      /* 2^-65 */
    }
    double s1 = shi + (slo + es), s2 = shi + (slo - es);
    double c1 = chi + (clo + ec), c2 = chi + (clo - ec);
    if (s1 != s2 || c1 != c2) return 0;
    if (neg) s1 = -s1;
    int q = ((int)kf) & 3;
    if (q == 0) { *s_out = s1; *c_out = c1; }
    else if (q == 1) { *s_out = c1; *c_out = -s1; }
    else if (q == 2) { *s_out = -s1; *c_out = -c1; }
    else { *s_out = -c1; *c_out = s1; }
    return 1;
}
#endif

/* RN(sin(x)), RN(cos(x)) together.  Out of line on the GPU (one copy per kernel: the planner kernels are bound by
 * instruction fetch, and this body was being inlined at ~20 call sites of the steering code); the pair comes back BY
 * VALUE (registers) -- results written through pointers would go through local memory. */
typedef struct { double s, c; } crm_sc;
CRM_NOINLINE crm_sc crm_sincos_v(double x) {
    crm_sc r;
#ifdef CRM_FAST
    if (crm_sincos_fast(x, &r.s, &r.c)) return r;
#endif
    crm_dd sd, cd;
    crm_sincos_dd(x, &sd, &cd);
    r.s = sd.hi; r.c = cd.hi;
    return r;
}
CRM_FN void crm_sincos(double x, double *s, double *c) {
    const crm_sc r = crm_sincos_v(x);
    *s = r.s; *c = r.c;
}

CRM_FN double crm_sin(double x) {
    if (x == 0.0) return x;
    double s, c;
    crm_sincos(x, &s, &c);
    return s;
}
CRM_FN double crm_cos(double x) {
    double s, c;
    crm_sincos(x, &s, &c);
    return c;
}

/* RN(tan(x)): the double-double quotient of crm_sincos_dd (relative error ~2^-100 before the final rounding) */
CRM_FN double crm_tan(double x) {
    if (x == 0.0) return x;
    crm_dd s, c;
    crm_sincos_dd(x, &s, &c);
    return crm_div(s, c).hi;
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

/* Double-double atan2 of (my, mx) given as double-doubles with 0 <= my <= mx (first octant):
 * t = my / mx, table entry atan(i/128), u = (t - i/128) / (1 + t i/128), |u| <= 2^-8, and
 * atan(u) = u - u^3/3 + u^5/5 - ... (two leading corrections in double-double, the tail in double). */
CRM_NOINLINE crm_dd crm_atan_octant_dd(crm_dd my, crm_dd mx) {
    crm_dd t = crm_div(my, mx);
    int i = (int)rint(t.hi * 128.0);
    double ti = (double)i * 0.0078125;
    crm_dd num = crm_fast_two_sum(t.hi - ti, t.lo);          /* t.hi - ti exact (Sterbenz) */
    crm_dd den = crm_add_d(crm_mul_d(t, ti), 1.0);
    crm_dd u = i == 0 ? t : crm_div(num, den);
    crm_dd z = crm_mul(u, u);
    double tail = -1.0 / 7.0 + z.hi * (1.0 / 9.0 - z.hi * (1.0 / 11.0));
    crm_dd p = crm_add(crm_mul_d(z, tail), crm_mk(CRM_FIFTH_H, CRM_FIFTH_L));
    p = crm_sub(crm_mul(z, p), crm_mk(CRM_THIRD_H, CRM_THIRD_L));
    crm_dd at = crm_add(u, crm_mul(u, crm_mul(z, p)));
    return i == 0 ? at : crm_add(crm_mk(crm_atan_tab[i][0], crm_atan_tab[i][1]), at);
}

/* full-circle double-double atan2 for finite (y, x), y != 0 */
CRM_FN crm_dd crm_atan2_dd(double y, double x) {
    double ax = fabs(x), ay = fabs(y);
    int swap = ay > ax;
    crm_dd a = crm_atan_octant_dd(crm_mk(swap ? ax : ay, 0.0), crm_mk(swap ? ay : ax, 0.0));
    if (swap) a = crm_sub(crm_mk(CRM_PIO2_H, CRM_PIO2_L), a);
    if (x < 0.0) a = crm_sub(crm_mk(CRM_PI_H, CRM_PI_L), a);
    return y < 0.0 ? crm_neg(a) : a;
}

#ifdef CRM_FAST
/* First phase for atan2 (see crm_sincos_fast): t = mn / mx as t0 + tl (quotient + exact-residual correction, relative
 * error < 2^-104), table entry atan(i/128), u = (t - i/128) / (1 + t i/128) as u0 + ul (relative error < 2^-100),
 * atan(u) = u0 + al, al = u0 z pa + ul, pa = -1/3 + z/5 - z^2/7 + z^3/9 (|u| <= 2^-8: |al| < 2^-17 |u0|, error < 2^-68 |u0|;
 * the dropped u^11/11 < 2^-83 |u0|), then the octant reflections with pi/2, pi as (H, L) pairs, each an error-free sum
 * of the leading parts plus one rounded sum of the tails (< 2^-105).  E = 2^-62 relative to the result. */
/* atan(mn / mx) for 0 <= mn <= mx given as (hi, lo) pairs, as a normalised (hi, lo) pair.  One reciprocal per quotient:
 * q0 = n * RN(1 / d) is within 2 ulp of the quotient, the residual fma(-q0, d, n) (rounded: 2^-53 of a 2^-51 term) times the
 * reciprocal restores it to < 2^-100 relative. */
CRM_NOINLINE crm_dd crm_atan_octant_fast(double mnh, double mnl, double mxh, double mxl) {
    double inv = 1.0 / mxh;
    double t0 = mnh * inv;
    double tl = ((fma(-t0, mxh, mnh) + mnl) - t0 * mxl) * inv;
    int i = (int)rint(t0 * 128.0);
    double u0, ul;
    if (i == 0) {
        u0 = t0; ul = tl;
    } else {
        double ti = (double)i * 0.0078125;
        crm_dd n = crm_fast_two_sum(t0 - ti, tl);             /* t0 - ti exact (Sterbenz) */
        crm_dd p = crm_two_prod(t0, ti);
        crm_dd d = crm_fast_two_sum(1.0, p.hi);
        double dl = d.lo + (p.lo + tl * ti);
        double invd = 1.0 / d.hi;
        u0 = n.hi * invd;
        ul = ((fma(-u0, d.hi, n.hi) + n.lo) - u0 * dl) * invd;
    }
    double z = u0 * u0;
    double pa = fma(fma(fma(1.1111111111111110e-01, z, -1.4285714285714285e-01), z, 0.2), z, -3.3333333333333331e-01);
    double al = u0 * (z * pa) + ul;
    crm_dd a;
    if (i == 0) {
        a = crm_fast_two_sum(u0, al);
    } else {
        a = crm_fast_two_sum(crm_atan_tab[i][0], u0);         /* atan(i/128) >= 2^-7.01 > |u0| */
        a = crm_fast_two_sum(a.hi, a.lo + (crm_atan_tab[i][1] + al));
    }
    return a;
}
/* pi/2 - a, pi - a for a normalised pair a in [0, pi/2] */
CRM_FN crm_dd crm_reflect_fast(double ch, double cl, crm_dd a) {
    crm_dd b = crm_two_sum(ch, -a.hi);
    return crm_fast_two_sum(b.hi, b.lo + (cl - a.lo));
}
CRM_FN int crm_round_fast(crm_dd a, double *out) {           /* E = 2^-62 relative */
    double e = fabs(a.hi) * 2.168404344971009e-19;
    double r1 = a.hi + (a.lo + e), r2 = a.hi + (a.lo - e);
    *out = r1;
    return r1 == r2;
}
CRM_FN int crm_atan2_fast(double y, double x, double *out) {
    double ax = fabs(x), ay = fabs(y);
    int swap = ay > ax;
    double mn = swap ? ax : ay, mx = swap ? ay : ax;
    if (!(mx < 1e150 && mn > 1e-150)) return 0;               /* residuals would leave the normal range */
    crm_dd a = crm_atan_octant_fast(mn, 0.0, mx, 0.0);
    if (swap) a = crm_reflect_fast(CRM_PIO2_H, CRM_PIO2_L, a);
    if (x < 0.0) a = crm_reflect_fast(CRM_PI_H, CRM_PI_L, a);
    double r;
    if (!crm_round_fast(a, &r)) return 0;
    *out = y < 0.0 ? -r : r;
    return 1;
}
/* sqrt((1 - x)(1 + x)) for |x| < 1 as a (hi, lo) pair, relative error < 2^-100: 1 -+ x are exact two-term sums, their
 * product in double-double, the root corrected once by its exact residual */
CRM_FN crm_dd crm_sqrt1mx2_fast(double x) {
    crm_dd a = crm_two_sum(1.0, -x), b = crm_two_sum(1.0, x);
    crm_dd p = crm_two_prod(a.hi, b.hi);
    double pl = p.lo + (a.hi * b.lo + a.lo * b.hi);
    double s0 = sqrt(p.hi);
    double sl = (fma(-s0, s0, p.hi) + pl) / (2.0 * s0);
    return crm_fast_two_sum(s0, sl);
}
/* first phases of acos / asin for 0 < |x| < 1 (the Dubins RLR / LRL words, every Reeds-Shepp word family with an arc
 * in the middle): the same octant core on (sqrt(1 - x^2), |x|) */
CRM_FN int crm_acos_fast(double x, double *out) {
    crm_dd y = crm_sqrt1mx2_fast(x);
    double ax = fabs(x);
    if (!(y.hi > 1e-150 && ax > 1e-150)) return 0;
    int swap = y.hi > ax;
    crm_dd a = swap ? crm_atan_octant_fast(ax, 0.0, y.hi, y.lo) : crm_atan_octant_fast(y.hi, y.lo, ax, 0.0);
    if (swap) a = crm_reflect_fast(CRM_PIO2_H, CRM_PIO2_L, a);
    if (x < 0.0) a = crm_reflect_fast(CRM_PI_H, CRM_PI_L, a);
    return crm_round_fast(a, out);
}
CRM_FN int crm_asin_fast(double x, double *out) {
    crm_dd c = crm_sqrt1mx2_fast(x);
    double ax = fabs(x);
    if (!(c.hi > 1e-150 && ax > 1e-150)) return 0;
    int swap = ax > c.hi;
    crm_dd a = swap ? crm_atan_octant_fast(c.hi, c.lo, ax, 0.0) : crm_atan_octant_fast(ax, 0.0, c.hi, c.lo);
    if (swap) a = crm_reflect_fast(CRM_PIO2_H, CRM_PIO2_L, a);
    double r;
    if (!crm_round_fast(a, &r)) return 0;
    *out = x < 0.0 ? -r : r;
    return 1;
}
#endif

/* theta = RN(atan2(y, x)); *s = RN(sin(theta)); *c = RN(cos(theta)) -- the three calls of
 * rrt_04:1236 and :1100-1101.  With A the exact angle and e = theta - A (|e| <= ulp(theta) / 2):
 *   cos(theta) = x/d - e * y/d,  sin(theta) = y/d + e * x/d,  d = sqrt(x^2 + y^2)
 * (e^2 < 2^-104 is dropped); x/d and y/d are formed in double-double, no sin/cos series needed.
 * Near the axes cos or sin of theta is itself of the order of e, so e must be known to ~2^-53
 * relative, not absolute: |A| = k pi/2 + sa * a with a = atan(min/max) in [0, pi/4] from the octant core
 * (relative error 2^-100), and e = (|theta| - k PIO2_H) - sa a.hi - k PIO2_L - sa a.lo - k PIO2_LL is
 * summed in double-double (the first difference is exact by Sterbenz). */
typedef struct { double th, s, c; } crm_tsc;
CRM_FN double crm_atan2_sincos_body(double y, double x, double *s, double *c);
/* by value (registers) across the call; see crm_sincos_v */
CRM_NOINLINE crm_tsc crm_atan2_sincos_v(double y, double x) {
    crm_tsc r;
    r.th = crm_atan2_sincos_body(y, x, &r.s, &r.c);
    return r;
}
CRM_FN double crm_atan2_sincos(double y, double x, double *s, double *c) {
    const crm_tsc r = crm_atan2_sincos_v(y, x);
    *s = r.s; *c = r.c;
    return r.th;
}
CRM_FN double crm_atan2_sincos_body(double y, double x, double *s, double *c) {
    if (y == 0.0) { /* includes (0, 0): atan2 = +-0 or +-pi */
        int neg = (x < 0.0) || (x == 0.0 && signbit(x));
        if (!neg) { *s = y; *c = 1.0; return y; }
        double th = copysign(CRM_PI_H, y);
        *s = copysign(CRM_PI_L, y); /* RN(sin(RN(pi))) = RN(pi - RN(pi)) */
        *c = -1.0;
        return th;
    }
#ifdef CRM_FAST
    {   /* first phases: RN(atan2), then RN(sin), RN(cos) of that double -- the three values by definition */
        double th;
        if (crm_atan2_fast(y, x, &th) && crm_sincos_fast(th, s, c)) return th;
    }
#endif
    double ax = fabs(x), ay = fabs(y);
    int swap = ay > ax, xneg = x < 0.0;
    crm_dd a = crm_atan_octant_dd(crm_mk(swap ? ax : ay, 0.0), crm_mk(swap ? ay : ax, 0.0));
    if (swap != xneg) a = crm_neg(a);                 /* sa * a */
    double ch = swap ? CRM_PIO2_H : CRM_PI_H, cl = swap ? CRM_PIO2_L : CRM_PI_L;
    double cll = swap ? CRM_PIO2_LL : CRM_PI_LL;
    double tabs;
    crm_dd e;
    if (swap || xneg) {
        tabs = crm_add(crm_mk(ch, cl), a).hi;
        e = crm_add(crm_two_sum(tabs - ch, -a.hi), crm_two_sum(-cl, -a.lo));
        e = crm_add_d(e, -cll);
    } else {
        tabs = a.hi;
        e = crm_mk(-a.lo, 0.0);
    }
    if (y < 0.0) e = crm_neg(e);
    /* d = sqrt(x*x + y*y) and 1/d in double-double */
    crm_dd d2 = crm_add(crm_two_prod(x, x), crm_two_prod(y, y));
    double d0 = sqrt(d2.hi);
    crm_dd r = crm_sub(d2, crm_two_prod(d0, d0));
    crm_dd d = crm_fast_two_sum(d0, r.hi / (2.0 * d0));
    double i0 = 1.0 / d.hi;
    r = crm_sub(crm_mk(1.0, 0.0), crm_mul_d(d, i0));
    crm_dd inv = crm_fast_two_sum(i0, r.hi * i0);
    crm_dd cd = crm_mul_d(inv, x), sd = crm_mul_d(inv, y);
    *c = crm_sub(cd, crm_mul(e, sd)).hi;
    *s = crm_add(sd, crm_mul(e, cd)).hi;
    return y < 0.0 ? -tabs : tabs;
}


CRM_NOINLINE double crm_atan2(double y, double x) {
    if (y == 0.0) {
        int neg = (x < 0.0) || (x == 0.0 && signbit(x));
        return neg ? copysign(CRM_PI_H, y) : y;
    }
#ifdef CRM_FAST
    double r_;
    if (crm_atan2_fast(y, x, &r_)) return r_;
#endif
    return crm_atan2_dd(y, x).hi;
}

/* RN(acos(x)) for |x| <= 1 (NaN outside): acos(x) = atan2(sqrt((1 - x)(1 + x)), x) with the square root
 * carried in double-double into the same octant core as crm_atan2. */
CRM_NOINLINE double crm_acos(double x) {
    if (!(fabs(x) <= 1.0)) return x - x == 0.0 ? (x - x) / (x - x) : x + x; /* NaN */
    if (x == 1.0) return 0.0;
    if (x == -1.0) return CRM_PI_H;
#ifdef CRM_FAST
    { double r_; if (crm_acos_fast(x, &r_)) return r_; }
#endif
    crm_dd a = crm_two_sum(1.0, -x), b = crm_two_sum(1.0, x); /* exact */
    crm_dd p = crm_mul(a, b);
    double s0 = sqrt(p.hi);
    crm_dd e = crm_sub(p, crm_two_prod(s0, s0));
    crm_dd y = crm_fast_two_sum(s0, e.hi / (2.0 * s0)); /* sqrt(p) to ~2^-104 */
    crm_dd ax = crm_mk(fabs(x), 0.0);
    int swap = y.hi > ax.hi;
    crm_dd at = crm_atan_octant_dd(swap ? ax : y, swap ? y : ax);
    if (swap) at = crm_sub(crm_mk(CRM_PIO2_H, CRM_PIO2_L), at);
    if (x < 0.0) at = crm_sub(crm_mk(CRM_PI_H, CRM_PI_L), at);
    return at.hi;
}

/* RN(asin(x)) for |x| <= 1 (NaN outside): asin(x) = atan2(x, sqrt((1 - x)(1 + x))), same double-double core as acos. */
CRM_NOINLINE double crm_asin(double x) {
    if (!(fabs(x) <= 1.0)) return x - x == 0.0 ? (x - x) / (x - x) : x + x; /* NaN */
    if (x == 0.0) return x;
    if (fabs(x) == 1.0) return copysign(CRM_PIO2_H, x);
#ifdef CRM_FAST
    { double r_; if (crm_asin_fast(x, &r_)) return r_; }
#endif
    crm_dd a = crm_two_sum(1.0, -x), b = crm_two_sum(1.0, x); /* exact */
    crm_dd p = crm_mul(a, b);
    double s0 = sqrt(p.hi);
    crm_dd e = crm_sub(p, crm_two_prod(s0, s0));
    crm_dd c = crm_fast_two_sum(s0, e.hi / (2.0 * s0)); /* sqrt(1 - x^2) to ~2^-104 */
    crm_dd ax = crm_mk(fabs(x), 0.0);
    int swap = ax.hi > c.hi;                               /* |x| > cos: angle above pi/4 */
    crm_dd at = crm_atan_octant_dd(swap ? c : ax, swap ? ax : c);
    if (swap) at = crm_sub(crm_mk(CRM_PIO2_H, CRM_PIO2_L), at);
    return x < 0.0 ? -at.hi : at.hi;
}

#endif /* RRTK_CRMATH_H */
