/* crmath_fastcheck.c -- TEST INFRASTRUCTURE.  The first-phase (Ziv) evaluations of crmath.h (-DCRM_FAST, what the GPU
 * build uses) against the double-double path they stand in for: whenever the fast form accepts, its result must be the
 * double-double path's, bit for bit.  Prints "<name> n accepted mismatches" per function; exit code 1 on any mismatch.
 *   gcc -O2 -ffp-contract=off -march=native -DCRM_FAST crmath_fastcheck.c -lm -o _build/crmath_fastcheck && ./crmath_fastcheck [n] */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../robotics-path-planning_b200/csrc/crmath.h"

/* the double-double acos / asin alone (what crm_acos / crm_asin compute when their first phase declines) */
static double acos_dd(double x) {
    crm_dd a = crm_two_sum(1.0, -x), b = crm_two_sum(1.0, x);
    crm_dd p = crm_mul(a, b);
    double s0 = sqrt(p.hi);
    crm_dd e = crm_sub(p, crm_two_prod(s0, s0));
    crm_dd y = crm_fast_two_sum(s0, e.hi / (2.0 * s0));
    crm_dd ax = crm_mk(fabs(x), 0.0);
    int swap = y.hi > ax.hi;
    crm_dd at = crm_atan_octant_dd(swap ? ax : y, swap ? y : ax);
    if (swap) at = crm_sub(crm_mk(CRM_PIO2_H, CRM_PIO2_L), at);
    if (x < 0.0) at = crm_sub(crm_mk(CRM_PI_H, CRM_PI_L), at);
    return at.hi;
}
static double asin_dd(double x) {
    crm_dd a = crm_two_sum(1.0, -x), b = crm_two_sum(1.0, x);
    crm_dd p = crm_mul(a, b);
    double s0 = sqrt(p.hi);
    crm_dd e = crm_sub(p, crm_two_prod(s0, s0));
    crm_dd c = crm_fast_two_sum(s0, e.hi / (2.0 * s0));
    crm_dd ax = crm_mk(fabs(x), 0.0);
    int swap = ax.hi > c.hi;
    crm_dd at = crm_atan_octant_dd(swap ? c : ax, swap ? ax : c);
    if (swap) at = crm_sub(crm_mk(CRM_PIO2_H, CRM_PIO2_L), at);
    return x < 0.0 ? -at.hi : at.hi;
}

static uint64_t st = 0x9E3779B97F4A7C15ull;
static uint64_t rnd(void) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; }
static double u01(void) { return (double)(rnd() >> 11) * (1.0 / 9007199254740992.0); }
static double nudge(double x, int k) { int64_t b; memcpy(&b, &x, 8); b += k; memcpy(&x, &b, 8); return x; }

static long n_sc, ok_sc, bad_sc, n_at, ok_at, bad_at, n_as, bad_as, n_ac, ok_ac, bad_ac;

static void check_sincos(double x) {
    double s, c;
    crm_dd sd, cd;
    if (!(fabs(x) < 1.6e6)) return;          /* crmath.h's domain: finite, |x| < 2^20 pi/2 */
    n_sc++;
    if (!crm_sincos_fast(x, &s, &c)) return;
    ok_sc++;
    crm_sincos_dd(x, &sd, &cd);
    if (s != sd.hi || c != cd.hi) {
        if (bad_sc < 10) fprintf(stderr, "sincos mismatch x=%a: fast (%a, %a) dd (%a, %a)\n", x, s, c, sd.hi, cd.hi);
        bad_sc++;
    }
}
static void check_atan2(double y, double x) {
    double r;
    if (!(fabs(x) < 1e300 && fabs(y) < 1e300)) return;
    n_at++;
    if (y == 0.0 || !crm_atan2_fast(y, x, &r)) return;
    ok_at++;
    double d = crm_atan2_dd(y, x).hi;
    if (r != d) {
        if (bad_at < 10) fprintf(stderr, "atan2 mismatch y=%a x=%a: fast %a dd %a\n", y, x, r, d);
        bad_at++;
    }
}

/* crm_atan2_sincos as compiled here (first phases + fallback) against its double-double body alone: the three values are
 * RN(atan2), RN(sin(that)), RN(cos(that)) */
static void check_atan2_sincos(double y, double x) {
    if (!(fabs(x) < 1e300 && fabs(y) < 1e300) || y == 0.0) return;
    double s, c, th = crm_atan2_sincos(y, x, &s, &c);
    double td = crm_atan2_dd(y, x).hi;
    crm_dd sd, cd;
    crm_sincos_dd(td, &sd, &cd);
    n_as++;
    if (th != td || s != sd.hi || c != cd.hi) {
        if (bad_as < 10) fprintf(stderr, "atan2_sincos mismatch y=%a x=%a: (%a %a %a) vs (%a %a %a)\n", y, x, th, s, c, td, sd.hi, cd.hi);
        bad_as++;
    }
}

static void check_acos_asin(double x) {
    double r;
    if (!(fabs(x) < 1.0) || x == 0.0) return;
    n_ac += 2;
    if (crm_acos_fast(x, &r)) { ok_ac++; if (r != acos_dd(x)) { if (bad_ac < 10) fprintf(stderr, "acos mismatch x=%a: %a vs %a\n", x, r, acos_dd(x)); bad_ac++; } }
    if (crm_asin_fast(x, &r)) { ok_ac++; if (r != asin_dd(x)) { if (bad_ac < 10) fprintf(stderr, "asin mismatch x=%a: %a vs %a\n", x, r, asin_dd(x)); bad_ac++; } }
}

int main(int argc, char **argv) {
    long n = argc > 1 ? atol(argv[1]) : 2000000;
    for (long i = 0; i < n; i++) {
        /* angles of the planners (yaws, word lengths, their sums), wide range, tiny values */
        check_sincos((u01() * 2.0 - 1.0) * 6.5);
        check_sincos((u01() * 2.0 - 1.0) * 40.0);
        check_sincos((u01() * 2.0 - 1.0) * 1e5);
        check_sincos(ldexp(u01() * 2.0 - 1.0, -(int)(rnd() % 60)));
        /* next to multiples of pi/2 and next to the table points i/128 */
        double k = (double)((long)(rnd() % 4001) - 2000);
        check_sincos(nudge(k * 1.5707963267948966, (int)(rnd() % 2001) - 1000));
        check_sincos(k * 1.5707963267948966 + ldexp(u01() - 0.5, -(int)(rnd() % 50)));
        check_sincos(nudge((double)(rnd() % 104) * 0.0078125 + 1.5707963267948966 * (double)(rnd() % 8), (int)(rnd() % 65) - 32));
        check_sincos(nudge((double)(rnd() % 830) * 0.00390625, (int)(rnd() % 65) - 32));
        /* atan2: planner-sized coordinates, ratios next to table points, extreme ratios, all quadrants */
        double sx = (rnd() & 1) ? 1.0 : -1.0, sy = (rnd() & 2) ? 1.0 : -1.0;
        check_atan2(sy * u01() * 30.0, sx * u01() * 30.0);
        check_atan2(sy * u01(), sx * ldexp(u01(), (int)(rnd() % 80) - 40));
        double m = 0.5 + u01();
        check_atan2(sy * nudge(m * (double)(rnd() % 129) * 0.0078125, (int)(rnd() % 65) - 32), sx * m);
        check_atan2(sy * m, sx * nudge(m * (double)(rnd() % 129) * 0.0078125, (int)(rnd() % 65) - 32));
        check_atan2(sy * nudge(m, (int)(rnd() % 9) - 4), sx * m);
        check_atan2(sy * 2.0, sx * u01() * 40.0);
        check_acos_asin(sx * u01());
        check_acos_asin(sx * (1.0 - ldexp(u01(), -(int)(rnd() % 50))));
        check_acos_asin(sx * ldexp(u01(), -(int)(rnd() % 60)));
        check_acos_asin(sx * nudge(0.70710678118654752, (int)(rnd() % 2001) - 1000));
        check_atan2_sincos(sy * u01() * 17.0, sx * u01() * 17.0);
        check_atan2_sincos(sy * ldexp(u01(), -(int)(rnd() % 40)), sx * u01());
    }
    printf("sincos %ld %ld %ld\natan2 %ld %ld %ld\natan2_sincos %ld %ld %ld\nacos_asin %ld %ld %ld\n", n_sc, ok_sc, bad_sc, n_at, ok_at, bad_at,
           n_as, n_as, bad_as, n_ac, ok_ac, bad_ac);
    return (bad_sc || bad_at || bad_as || bad_ac) ? 1 : 0;
}
