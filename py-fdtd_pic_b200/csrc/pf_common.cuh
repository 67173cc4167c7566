// pf_common.cuh -- shared device/host helpers for libpyfdtd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/pyfdtd_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libpyfdtd_b200 is written for sm_100a (B200) only"
#endif

namespace pf {

extern thread_local char g_err[512];               // pf_last_error(): per host thread
extern std::atomic<unsigned long long> g_launches; // pf_launch_count()

int set_err(int code, const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);

// Optional per-launch timing (pf_profile_enable / pf_profile_report): while enabled, a ProfScope brackets a kernel launch
// with CUDA events on the launch's own stream and files them under the kernel's name.  Defined in pf_probe.cu.
struct ProfScope {
    cudaStream_t st;
    const char *name;
    cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(cudaStream_t s, const char *kernel_name);
    ~ProfScope();
};

#define PF_CUDA(call)                                        \
    do {                                                     \
        int _rc = ::pf::check_cuda((call), #call);           \
        if (_rc) return _rc;                                 \
    } while (0)

#define PF_LAUNCH_CHECK(name)                                \
    do {                                                     \
        ++::pf::g_launches;                                  \
        int _rc = ::pf::check_cuda(cudaGetLastError(), name);\
        if (_rc) return _rc;                                 \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Arithmetic policy.  The reference's loops are compiled by numba/LLVM without contraction, so
// bit-identical results need separately rounded multiplies and adds: Exact uses the _rn
// intrinsics, which nvcc never fuses.  Fused lets a*b+c become one DFMA (PF_F_FMA).
// ---------------------------------------------------------------------------------------------
struct Exact {
    using real = double;
    static constexpr bool newton = false;
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    // a*b + c with two roundings
    static __device__ __forceinline__ double mad(double a, double b, double c) { return __dadd_rn(__dmul_rn(a, b), c); }
    // c - a*b with two roundings
    static __device__ __forceinline__ double nmad(double a, double b, double c) { return __dsub_rn(c, __dmul_rn(a, b)); }
};
struct Fused {
    using real = double;
    static constexpr bool newton = false;
    static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
    static __device__ __forceinline__ double add(double a, double b) { return a + b; }
    static __device__ __forceinline__ double sub(double a, double b) { return a - b; }
    static __device__ __forceinline__ double mad(double a, double b, double c) { return fma(a, b, c); }
    static __device__ __forceinline__ double nmad(double a, double b, double c) { return fma(-a, b, c); }
};

// PF_F_FP32: the on-chip state is advanced in single precision with contraction.  Not a parity mode:
// reported against a stated 1e-5 tolerance (fields relative to the peak of the trace).
// PF_F_NEWTON in the tile engine: Exact arithmetic everywhere except the cubic material law, which is the
// inlined Newton iteration (cubic_root0_newton) with reciprocal-multiply divisions.
struct ExactNewton : Exact {
    static constexpr bool newton = true;
};
struct Fast32 {
    using real = float;
    static constexpr bool newton = false;
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
    static __device__ __forceinline__ float mad(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ float nmad(float a, float b, float c) { return fmaf(-a, b, c); }
};

// Correctly rounded x / d for a loop-invariant divisor d with r = RN(1/d) precomputed
// (Markstein: q = RN(x*r); e = x - q*d exactly by FMA; RN(q + e*r) == RN(x/d)).
// The two-FMA correction is exact whenever e is representable; for |x| so small that the
// residual could underflow (or non-finite x) the IEEE divide is used instead.  This is an
// implementation of the division the reference performs, not a contraction of its arithmetic.
static __device__ __noinline__ double div_const_slow(double x, double d) { return __ddiv_rn(x, d); }

__device__ __forceinline__ double div_const(double x, double d, double r)
{
    double q = __dmul_rn(x, r);
    double e = __fma_rn(-q, d, x);
    double q2 = __fma_rn(e, r, q);
    // Guard on the exponent field with integer ops (keeps the test off the FP64 pipe): the two-FMA
    // correction is exact for 2^-823 <= |x| < 2^825.  Outside that range: zero (by far the most
    // common case -- cells the wave has not reached) already gave q2 = +-0 up to the sign, which q
    // carries correctly; denormal-range, huge, inf and nan inputs take the IEEE divide, kept out of
    // line so that it is never evaluated speculatively.
    const int hi = __double2hiint(x);
    const unsigned ebits = (unsigned)hi & 0x7ff00000u;
    if (ebits - 0x0c800000u >= 0x67000000u) {
        q2 = q;
        if ((((unsigned)hi & 0x7fffffffu) | (unsigned)__double2loint(x)) != 0u) q2 = div_const_slow(x, d);
    }
    return q2;
}

// The same division split in two for loops over several cells: div_const_fast() is the unguarded
// two-FMA path and folds the guard quantity into `key` (max over cells; integer pipe only);
// div_const_in_range(key) tells whether every cell of the loop was inside the exact range, and
// div_const_fix() redoes one cell exactly when it was not.
constexpr unsigned DIVC_LO = 0x0c800000u, DIVC_RANGE = 0x67000000u;
__device__ __forceinline__ double div_const_fast(double x, double d, double r, unsigned &key)
{
    const double q = __dmul_rn(x, r);
    const double e = __fma_rn(-q, d, x);
    key = max(key, ((unsigned)__double2hiint(x) & 0x7ff00000u) - DIVC_LO);
    return __fma_rn(e, r, q);
}
__device__ __forceinline__ bool div_const_in_range(unsigned key) { return key < DIVC_RANGE; }
__device__ __forceinline__ double div_const_fix(double x, double d, double r, double fast)
{
    const int hi = __double2hiint(x);
    if ((((unsigned)hi & 0x7ff00000u) - DIVC_LO) < DIVC_RANGE) return fast;
    if ((((unsigned)hi & 0x7fffffffu) | (unsigned)__double2loint(x)) == 0u) return __dmul_rn(x, r);   // +-0
    return div_const_slow(x, d);
}

// ---------------------------------------------------------------------------------------------
// First root of a x^3 + b x^2 + c x + d as the reference's solver returns it
// (CubicEquationSolver.py:29-105; only root[0] is consumed, BaseFDTD11.py:838).
// Per-run constants (functions of a,b,c only) are hoisted into CubicConsts on the device once per
// thread; per-cell work is the d-dependent part.  Arithmetic order follows the reference.
// ---------------------------------------------------------------------------------------------
struct CubicConsts {
    double a, b, c;
    double inv_a;        // RN(1/a) for div_const
    double f;            // findF(a,b,c)
    double g_ab;         // (2 b^3)/a^3 - (9 b c)/a^2   (first two terms of findG)
    double f3_27;        // f^3/27
    double b_3a;         // b/(3a)
    double inv_c;        // RN(1/c): starting slope of the Newton variant
    int newton;          // PF_F_NEWTON requested and a, b >= 0, c > 0 (see cubic_root0_newton)
    int pad;
};

// Device-side constants (generic pf_cubic_root0 entry, where every polynomial has its own a,b,c).
// Grid runs use cubic_consts_host() instead, which calls the same libm pow() as CPython does.
__device__ __forceinline__ CubicConsts cubic_consts_dev(double a, double b, double c)
{
    CubicConsts k;
    k.a = a; k.b = b; k.c = c;
    k.inv_a = 1.0 / a;
    double a2 = a * a;               // pow(a,2.0) is correctly rounded == a*a
    double b2 = b * b;
    // pow(x,3.0) in glibc is correctly rounded to < 1ulp of x^3; x*x*x can differ in the last bit,
    // so cube through an exact double-double product: x^3 = RN(x2_hi*x + x2_lo*x).
    double b2lo = __fma_rn(b, b, -b2), a2lo = __fma_rn(a, a, -a2);
    double b3 = __fma_rn(b2, b, b2lo * b);
    double a3 = __fma_rn(a2, a, a2lo * a);
    k.f = __ddiv_rn(__dsub_rn(__ddiv_rn(__dmul_rn(3.0, c), a), __ddiv_rn(b2, a2)), 3.0);
    k.g_ab = __dsub_rn(__ddiv_rn(__dmul_rn(2.0, b3), a3), __ddiv_rn(__dmul_rn(__dmul_rn(9.0, b), c), a2));
    double f2 = k.f * k.f, f2lo = __fma_rn(k.f, k.f, -f2);
    double f3 = __fma_rn(f2, k.f, f2lo * k.f);
    k.f3_27 = __ddiv_rn(f3, 27.0);
    k.b_3a = __ddiv_rn(b, __dmul_rn(3.0, a));
    k.inv_c = 1.0 / c;
    k.newton = 0;
    k.pad = 0;
    return k;
}

// v ** (1/3.0) on |v| with the sign restored (CubicEquationSolver.py:76-85).  The reference's
// exponent is the double nearest to 1/3, c = 1/3 - 1.85e-17, so pow(x, c) = cbrt(x) * x^(c - 1/3)
// = cbrt(x) * (1 + (c - 1/3) ln x + O(1e-31)).  cbrt() is ~5x cheaper than CUDA's pow() and the
// correction is below 1e-15, so single-precision ln is ample.  Accuracy ~1.5 ulp -- the same class as
// CUDA pow() (2 ulp) against glibc pow(), which is what the 1e-10 tolerance on Acubic covers.
__device__ __forceinline__ double pow_third(double x)   // x >= 0
{
    const double c_minus_third = -1.850371707708594e-17;   // double(1/3.0) - 1/3, exactly
    const double s = cbrt(x);
    const double lnx = (double)__logf((float)x);
    return (x > 0.0) ? __fma_rn(s, c_minus_third * lnx, s) : s;
}

__device__ __forceinline__ double signed_cbrt_pow(double v)
{
    return copysign(pow_third(fabs(v)), v);
}

__device__ __forceinline__ double cubic_root0(const CubicConsts &k, double d)
{
    if (k.a == 0.0) {  // linear / quadratic fallbacks (CubicEquationSolver.py:30-46); never hit on the NL path
        if (k.b == 0.0) return __ddiv_rn(-d, k.c);
        double D = __dsub_rn(__dmul_rn(k.c, k.c), __dmul_rn(__dmul_rn(4.0, k.b), d));
        double twob = __dmul_rn(2.0, k.b);
        return (D >= 0.0) ? __ddiv_rn(__dadd_rn(-k.c, sqrt(D)), twob) : __ddiv_rn(-k.c, twob);
    }
    double t3 = div_const(__dmul_rn(27.0, d), k.a, k.inv_a);                 // 27*d/a
    double g = div_const(__dadd_rn(k.g_ab, t3), 27.0, 1.0 / 27.0);           // findG
    double gg4 = __dmul_rn(__dmul_rn(g, g), 0.25);                           // g**2/4
    double h = __dadd_rn(gg4, k.f3_27);                                      // findH
    double ghalf = __dmul_rn(g, 0.5);
    if (h > 0.0) {  // one real root -- the only branch the NL path takes
        double sh = sqrt(h);
        double S = signed_cbrt_pow(__dadd_rn(-ghalf, sh));
        double U = signed_cbrt_pow(__dsub_rn(-ghalf, sh));
        return __dsub_rn(__dadd_rn(S, U), k.b_3a);
    }
    if (k.f == 0.0 && g == 0.0 && h == 0.0) {  // triple root
        double da = __ddiv_rn(d, k.a);
        return (da >= 0.0) ? -pow(da, 1.0 / 3.0) : pow(-da, 1.0 / 3.0);
    }
    // three real roots: x1 = 2 j cos(k/3) - b/(3a)
    double i = sqrt(__dsub_rn(gg4, h));
    double j = pow(i, 1.0 / 3.0);
    double kk = acos(-__ddiv_rn(g, __dmul_rn(2.0, i)));
    return __dsub_rn(__dmul_rn(__dmul_rn(2.0, j), cos(__ddiv_rn(kk, 3.0))), k.b_3a);
}

// PF_F_NEWTON: the same root by Newton iteration.  For a, b >= 0, c > 0 and d = -q^2 < 0 the polynomial
// p(x) = a x^3 + b x^2 + c x - q^2 is increasing and convex on x > 0 with exactly one positive root -- the
// root[0] the closed form returns in every branch (largest real root) -- and x0 = q^2/c >= root, so the
// iteration descends monotonically.  1/p'(x) is carried along by its own Newton step (no division).
// The result is the correctly converged root (|p| at rounding level); the reference's closed form loses
// ~4 digits to cancellation in (S+U) - b/3a, so the two differ by up to ~1e-11 absolute, inside the
// 1e-10 absolute tolerance Acubic is held to (its effect on Ex is below 1e-14 relative).
// General starting point and true divisions: used when x0 = q^2/c is far from the root (p'(x0) > 1.5 c), where the
// carried reciprocal of the fast path would lag behind.  Out of line: rare on the grid path.
static __device__ __noinline__ double cubic_root0_newton_far(double a, double b, double c, double q2)
{
    double x = q2 / c;                                   // each single term bounds the root from above
    if (b > 0.0) x = fmin(x, sqrt(q2 / b));
    if (a > 0.0) x = fmin(x, cbrt(q2 / a));
    const double a3 = 3.0 * a, b2 = 2.0 * b;
    for (int it = 0; it < 100; ++it) {
        const double p = fma(fma(fma(a, x, b), x, c), x, -q2);
        const double dp = fma(fma(a3, x, b2), x, c);
        const double step = p / dp;
        x -= step;
        if (fabs(step) <= 1e-10 * x) {                   // one more step from here is converged to rounding
            const double p2 = fma(fma(fma(a, x, b), x, c), x, -q2);
            x -= p2 / fma(fma(a3, x, b2), x, c);
            break;
        }
    }
    return x;
}

__device__ __forceinline__ double cubic_root0_newton(double a, double b, double c, double inv_c, double q2)
{
    const double a3 = 3.0 * a, b2 = 2.0 * b;
    double x = q2 * inv_c, r = inv_c, step;
    if (fma(fma(a3, x, b2), x, c) * inv_c > 1.5) return cubic_root0_newton_far(a, b, c, q2);
#ifndef PF_NEWTON_PLAIN_START
    // first-order start: with t = b x0/c the root is x0 (1 - t + O(t^2)) and 1/p' is (1 - 2t + O(t^2))/c, which saves an
    // iteration in the weakly nonlinear regime (t ~ 5e-4 |E|^2 on the sweeps).  The iteration no longer starts above
    // the root; from below its first step lands above it (p convex), from where it descends as before.
    const double t = (b * inv_c) * x;
    x = fma(-t, x, x);
    r = fma(-2.0 * t, inv_c, inv_c);
    constexpr int MIN_IT = 2;
#else
    constexpr int MIN_IT = 3;
#endif
    int it = 0;
    do {
        const double p = fma(fma(fma(a, x, b), x, c), x, -q2);
        const double dp = fma(fma(a3, x, b2), x, c);
        r = fma(r, fma(-dp, r, 1.0), r);                 // 0 < r dp < 2: p'(x) <= p'(x0) <= 1.5 c and r <= 1/c
        step = p * r;
        x -= step;
        ++it;
        // superlinear: once a step is below 1e-10 x the error left after it is far below one ulp
    } while (it < MIN_IT || (it < 24 && fabs(step) > 1e-10 * x));
    return x;
}
__device__ __forceinline__ double cubic_root0_newton(const CubicConsts &k, double q2)
{
    return cubic_root0_newton(k.a, k.b, k.c, k.inv_c, q2);
}

// 1/x to ~1 ulp for normal positive x (MUFU seed + two Newton steps); PF_F_NEWTON's material law only.
__device__ __forceinline__ double rcp_newton(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

// The nonlinear material law of one cell under PF_F_NEWTON, small enough to be inlined: Acubic by Newton
// iteration, |Dn/eps0| and Dn/(den0 + den1 A) by reciprocal multiplication (each within ~1 ulp of the
// reference's divisions; the mode's tolerance is 1e-10).
struct NlNewtonConsts {
    double a, b, c, inv_c;
};
#ifdef PF_NEWTON_NOINLINE   // tuning experiment: the law as a call
#define PF_NEWTON_LAW_QUAL static __device__ __noinline__
#else
#define PF_NEWTON_LAW_QUAL __device__ __forceinline__
#endif
PF_NEWTON_LAW_QUAL void nl_material_law_newton(const NlNewtonConsts &k, double dn, double inv_eps0, double den0,
                                                       double den1, double &acub, double &e)
{
    const double q = dn * inv_eps0;
    const double q2 = q * q;
    double x = 0.0;
    if (q2 > 1e-8) x = cubic_root0_newton(k.a, k.b, k.c, k.inv_c, q2);
    const double den = fma(den1, x, den0);
    const double r = rcp_newton(den);
    const double e0 = dn * r;
    acub = x;
    e = fma(fma(-den, e0, dn), r, e0);
}

// Acubic of one cell: root0 of [cub, qua, one, -|Dx/eps0|^2] if |d| > 1e-8 else 0
// (BaseFDTD11.py:793-853).
__device__ __forceinline__ double acubic_cell(const CubicConsts &k, double dx, double eps0, double inv_eps0)
{
    double q = fabs(div_const(dx, eps0, inv_eps0));
    double d = -__dmul_rn(q, q);
    if (!(fabs(d) > 1e-8)) return 0.0;
    return k.newton ? cubic_root0_newton(k, -d) : cubic_root0(k, d);
}

// The whole nonlinear material law of one cell (AcubicFinder + NonLinExUpdate, BaseFDTD11.py:793-877),
// deliberately OUT OF LINE: inlined C times into an unrolled time-step body it overflows the
// instruction cache (ncu: stall_no_instruction 8.3 per issue), and at ~200 instructions per call the
// call overhead is noise.  Constants are read through the pointer (L1-resident).
struct NlResult {
    double a, e;
};
static __device__ __noinline__ NlResult nl_material_law(const CubicConsts *__restrict__ kc, double dx, double eps0,
                                                       double inv_eps0, double den0, double den1)
{
    NlResult r;
    r.a = acubic_cell(*kc, dx, eps0, inv_eps0);
    r.e = __ddiv_rn(dx, __dadd_rn(den0, __dmul_rn(den1, r.a)));
    return r;
}

// C cells at once (all of a thread's cells lie in the slab): one out-of-line copy whose C independent
// dependency chains the scheduler can interleave.
template <int C>
struct NlVec {
    double v[C];
};
template <int C>
struct NlResultVec {
    double a[C], e[C];
};
template <int C>
static __device__ __noinline__ NlResultVec<C> nl_material_law_vec(const CubicConsts *__restrict__ kc, NlVec<C> dx, double eps0,
                                                                double inv_eps0, double den0, double den1)
{
    NlResultVec<C> r;
    const CubicConsts k = *kc;
#pragma unroll
    for (int j = 0; j < C; ++j) {
        r.a[j] = acubic_cell(k, dx.v[j], eps0, inv_eps0);
        r.e[j] = __ddiv_rn(dx.v[j], __dadd_rn(den0, __dmul_rn(den1, r.a[j])));
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// The closed-form cubic material law of one cell, FAST PATH (tile engine, default cubic mode).  Same algorithm and
// operation order as acubic_cell() + cubic_root0() above -- AcubicFinder / CubicEquationSolver.solve / NonLinExUpdate,
// including the cancellation in T = -g/2 - sqrt(h) and in (S + U) - b/3a that parity has to reproduce -- but written
// for the one-real-root branch only and without per-operation special cases:
//   * the three divisions by run constants are the unguarded two-FMA Markstein form (exact for every operand the law
//     can produce on this branch);
//   * v ** (1/3.0) is cbrt by ONE Halley step from a single-precision seed ex2(lg2(v)/3) (the construction CUDA's cbrt
//     uses, minus its denormal / zero / infinity handling), and the same lg2 feeds the exponent correction
//     (1 + (double(1/3) - 1/3) ln v) that turns cbrt into the reference's pow(v, 1/3.0);
//   * Dn / (den0 + den1 A) is a reciprocal refined by two Newton steps with a final residual correction (<= 1 ulp).
// Anything outside the branch (a == 0, h <= 0, an operand outside the single-precision exponent range of the seed, a
// non-finite result) returns ok = false and the caller redoes the cell with nl_material_law() -- never taken on the
// nonlinear sweeps.  Accuracy class: ~1 ulp per elementary function, like the out-of-line law (whose cbrt / sqrt / div
// are CUDA's); the mode's tolerances (Acubic 1e-10 absolute, Ex 1e-10 relative) are tested against the reference goldens.
// ---------------------------------------------------------------------------------------------
struct NlFastConsts {
    double a, inv_a, g_ab, f3_27, b_3a;
};
__device__ __forceinline__ double pow_third_fast(double v, bool &ok)   // v != 0
{
    const double c_minus_third = -1.850371707708594e-17;   // double(1/3.0) - 1/3
    const double x = fabs(v);
    const float xf = (float)x;
    ok = ok && (xf > 1e-30f) && (xf < 1e30f);
    float lg, yf;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(xf));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"(lg * 0.333333343f));
    const double y = (double)yf;
    const double y2 = y * y;
    const double den = fma(y, y2 + y2, x);      // 2 y^3 + x
    const double num = fma(-y, y2, x);          // x - y^3
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    double e = fma(-den, r, 1.0);
    e = fma(e, e, e);
    r = fma(r, e, r);
    double s = fma(y, num * r, y);              // Halley step: cbrt(x) to rounding
    const double lnx = (double)(lg * 0.693147182f);
    s = fma(s, c_minus_third * lnx, s);         // -> x ** double(1/3)
    return copysign(s, v);
}
__device__ __forceinline__ bool nl_material_law_fast(const NlFastConsts &k, double dn, double eps0, double inv_eps0,
                                                     double den0, double den1, double &acub, double &e_out)
{
    bool ok = true;          // (the caller has checked a != 0 once per launch)
    // AcubicFinder: q = |Dn/eps0|, d = -q^2
    double q = __dmul_rn(dn, inv_eps0);
    q = fabs(__fma_rn(__fma_rn(-q, eps0, dn), inv_eps0, q));
    const double qq = __dmul_rn(q, q);
    const double d = -qq;
    double A = 0.0;
    if (qq > 1e-8) {         // warp-divergent only while the wave front passes
        const double x3 = __dmul_rn(27.0, d);
        double t3 = __dmul_rn(x3, k.inv_a);                                // 27 d / a
        t3 = __fma_rn(__fma_rn(-t3, k.a, x3), k.inv_a, t3);
        const double xg = __dadd_rn(k.g_ab, t3);
        double g = __dmul_rn(xg, 1.0 / 27.0);                              // findG
        g = __fma_rn(__fma_rn(-g, 27.0, xg), 1.0 / 27.0, g);
        const double h = __dadd_rn(__dmul_rn(__dmul_rn(g, g), 0.25), k.f3_27);   // findH
        const double ghalf = __dmul_rn(g, 0.5);
        ok = ok && (h > 0.0);
        const double sh = sqrt(h);
        const double S = pow_third_fast(__dadd_rn(-ghalf, sh), ok);
        const double U = pow_third_fast(__dsub_rn(-ghalf, sh), ok);
        A = __dsub_rn(__dadd_rn(S, U), k.b_3a);
    }
    const double den = __dadd_rn(den0, __dmul_rn(den1, A));
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    r = fma(r, fma(-den, r, 1.0), r);
    r = fma(r, fma(-den, r, 1.0), r);
    const double e0 = dn * r;
    const double e = fma(fma(-den, e0, dn), r, e0);
    // validity on the exponent fields (integer pipe): A and e finite (also catches NaN); Dn zero or inside the range in
    // which the two-FMA divisions above are exact
    const unsigned ea = (unsigned)__double2hiint(A) & 0x7ff00000u, ee = (unsigned)__double2hiint(e) & 0x7ff00000u;
    const unsigned hd = (unsigned)__double2hiint(dn);
    const bool dn_zero = ((hd << 1) | (unsigned)__double2loint(dn)) == 0u;
    ok = ok && ea != 0x7ff00000u && ee != 0x7ff00000u && (dn_zero || ((hd & 0x7ff00000u) - DIVC_LO) < DIVC_RANGE);
    acub = A;
    e_out = e;
    return ok;
}

// Host-side constants, evaluated exactly as CubicEquationSolver.findF/findG/findH do in CPython
// (float ** float -> libm pow).  Defined in pf_host.cu.
CubicConsts cubic_consts_host(double a, double b, double c);

// What the kernels see for one grid: the caller's descriptor plus host-derived constants.
struct GridDev {
    PfGrid g;
    CubicConsts k;
    double inv_eps0;
};
GridDev make_grid_dev(const PfGrid &g);

}  // namespace pf
