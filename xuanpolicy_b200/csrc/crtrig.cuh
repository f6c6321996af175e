// crtrig.cuh — correctly-rounded fp64 sin/cos for the environment kernels.
//
// gym computes sin/cos through the host libm, whose results are neither correctly rounded nor stable across
// CPUs, so "bit-exact with gym" is defined against a platform-independent target:
// the correctly-rounded value.  This routine evaluates sin/cos in double-double arithmetic (~2^-100 relative
// error) and rounds once; the result equals the correctly-rounded value unless the true value lies within
// ~2^-100 of a rounding boundary (never observed; the parity tests compare 10^7 points with libquadmath).
//
// Everything is built from individually-rounded + and * plus the exact fused multiply-add (fma is correctly
// rounded on both the GPU and an IEEE host), so the translation unit may be compiled with -fmad=false and the
// same source gives bit-identical results when compiled for the host (tests/host_crtrig.cpp does that).
//
// Valid for |x| <= 2^19 (Pendulum's |theta| stays below ~84 rad; CartPole's below 0.5 rad).
//
// Two phases (Ziv's strategy).  Phase 1 evaluates the tails of both series (terms z^4 and up) in plain double and only the
// four leading Horner steps in double-double: relative error < 2^-72, a third of the dependent-instruction chain of the full
// evaluation.  Its result (hi, lo) is accepted when rounding hi + (lo - E) and hi + (lo + E), E = 2^-68 |hi|, gives the same
// double — that double is then the correctly-rounded value, i.e. exactly what phase 2 returns.  Otherwise (expected once in
// ~2^14 calls) phase 2, the full double-double series, runs.  Both phases therefore return identical bits wherever phase 2
// is correctly rounded; tests/test_trig_fast_path.py checks that and the fallback rate on the host build.
#pragma once
#include <math.h>

#include "crtrig_consts.inc"

#if defined(__CUDACC__)
#define XB_HD __host__ __device__ __forceinline__
#else
#define XB_HD static inline
#endif

namespace xb {

struct dd {
    double hi, lo;
};

XB_HD dd two_sum(double a, double b) {
    double s = a + b;
    double bb = s - a;
    double e = (a - (s - bb)) + (b - bb);
    return dd{s, e};
}
XB_HD dd quick_two_sum(double a, double b) {  // requires |a| >= |b| (or a == 0)
    double s = a + b;
    return dd{s, b - (s - a)};
}
XB_HD dd two_prod(double a, double b) {
    double p = a * b;
    return dd{p, fma(a, b, -p)};
}
XB_HD dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    dd t = two_sum(a.lo, b.lo);
    dd u = quick_two_sum(s.hi, s.lo + t.hi);
    return quick_two_sum(u.hi, u.lo + t.lo);
}
XB_HD dd dd_add_d(dd a, double b) {
    dd s = two_sum(a.hi, b);
    return quick_two_sum(s.hi, s.lo + a.lo);
}
XB_HD dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    double cross = fma(a.hi, b.lo, a.lo * b.hi);
    return quick_two_sum(p.hi, p.lo + cross);
}

// r = x - k*pi/2 as a double-double, k = nearest integer to x*2/pi.  pi/2 is split into three 33-bit pieces
// (k*piece is exact for |k| < 2^20) plus a 53-bit tail: 152 bits of pi/2 in total.
XB_HD dd reduce_pio2(double x, int* quadrant) {
    double k = rint(x * XB_TWO_OVER_PI);
    *quadrant = (int)((long long)k & 3);
    double t = x - k * XB_PIO2_1;  // exact (Sterbenz) for k != 0
    dd r = two_sum(t, -(k * XB_PIO2_2));
    r = dd_add_d(r, -(k * XB_PIO2_3));
    dd p4 = two_prod(k, XB_PIO2_4);
    r = dd_add(r, dd{-p4.hi, -p4.lo});
    return r;
}

// sin and cos of a double-double r, |r| <= pi/4 (+ rounding slack), by Taylor series in double-double Horner form.
XB_HD void sincos_reduced(dd r, double* s, double* c) {
    const dd S[XB_SIN_TERMS] = XB_SIN_COEFFS;
    const dd C[XB_COS_TERMS] = XB_COS_COEFFS;
    dd z = dd_mul(r, r);
    dd ps = S[XB_SIN_TERMS - 1];
#pragma unroll
    for (int j = XB_SIN_TERMS - 2; j >= 0; --j) ps = dd_add(dd_mul(ps, z), S[j]);
    dd pc = C[XB_COS_TERMS - 1];
#pragma unroll
    for (int j = XB_COS_TERMS - 2; j >= 0; --j) pc = dd_add(dd_mul(pc, z), C[j]);
    dd rs = dd_add(r, dd_mul(r, dd_mul(z, ps)));  // r + r*z*S(z)
    dd rc = dd_add_d(dd_mul(z, pc), 1.0);         // 1 + z*C(z)
    *s = rs.hi;
    *c = rc.hi;
}

// Phase 1.  Returns false when either rounding test fails (the caller then runs the full evaluation).
XB_HD dd dd_add_nocancel(dd a, dd b) {  // a + b where no heavy cancellation occurs: ~2^-104 relative
    dd t = two_sum(a.hi, b.hi);
    return quick_two_sum(t.hi, t.lo + (a.lo + b.lo));
}
XB_HD bool round_test(dd v, double* out) {
    const double e = fabs(v.hi) * 0x1p-68;
    const double a = v.hi + (v.lo + e), b = v.hi + (v.lo - e);
    *out = a;
    return a == b;
}
XB_HD bool sincos_reduced_fast(dd r, double* s, double* c) {
    const dd S[XB_SIN_TERMS] = XB_SIN_COEFFS;
    const dd C[XB_COS_TERMS] = XB_COS_COEFFS;
    constexpr int kDD = 4;  // leading Horner steps carried in double-double
    if (fabs(r.hi) < 0x1p-20) return false;  // (a tiny reduced argument: its own relative accuracy is the limit)
    dd z = dd_mul(r, r);
    double ts = S[XB_SIN_TERMS - 1].hi, tc = C[XB_COS_TERMS - 1].hi;
#pragma unroll
    for (int j = XB_SIN_TERMS - 2; j >= kDD; --j) ts = fma(ts, z.hi, S[j].hi);
#pragma unroll
    for (int j = XB_COS_TERMS - 2; j >= kDD; --j) tc = fma(tc, z.hi, C[j].hi);
    dd ps = two_prod(z.hi, ts), pc = two_prod(z.hi, tc);
    ps = dd_add_nocancel(ps, S[kDD - 1]);
    pc = dd_add_nocancel(pc, C[kDD - 1]);
#pragma unroll
    for (int j = kDD - 2; j >= 0; --j) {
        ps = dd_add_nocancel(dd_mul(ps, z), S[j]);
        pc = dd_add_nocancel(dd_mul(pc, z), C[j]);
    }
    dd rs = dd_add(r, dd_mul(r, dd_mul(z, ps)));  // r + r*z*S(z)
    dd rc = dd_add_d(dd_mul(z, pc), 1.0);         // 1 + z*C(z)
    const bool ok_s = round_test(rs, s), ok_c = round_test(rc, c);
    return ok_s && ok_c;
}

#ifndef XB_TRIG_FAST
#define XB_TRIG_FAST 1
#endif

XB_HD void sincos_cr(double x, double* s, double* c, int* slow_path = nullptr) {
    double ax = fabs(x);
    if (ax < 0x1p-27) {  // sin x rounds to x and cos x to 1 (also keeps the sign of zero)
        *s = x;
        *c = 1.0;
        return;
    }
    int q = 0;
    dd r = dd{x, 0.0};
    if (ax > 0.78539816339744828) r = reduce_pio2(x, &q);
    double sr, cr;
    if (!XB_TRIG_FAST || !sincos_reduced_fast(r, &sr, &cr)) {
        sincos_reduced(r, &sr, &cr);
        if (slow_path) *slow_path += 1;
    }
    switch (q) {
        case 0: *s = sr; *c = cr; break;
        case 1: *s = cr; *c = -sr; break;
        case 2: *s = -sr; *c = -cr; break;
        default: *s = -cr; *c = sr; break;
    }
}

}  // namespace xb
