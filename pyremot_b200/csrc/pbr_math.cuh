// pbr_math.cuh — FP64 elementary functions for the rate expressions, written so that every coefficient is a
// constant-bank operand of the FMA that uses it.
//
// Why not libdevice's exp(): inlined, it costs ~40 instructions plus ~25 instructions that only materialise its
// 64-bit coefficients in registers (sm_100a has no 64-bit immediates; ncu/SASS of the first kernels showed 26 % of all
// issued instructions were such moves), and the RHS kernels are instruction-issue bound.  The methanol/DME kinetics
// evaluate 10 exponentials per node.  A __constant__ table that is NOT const-qualified cannot be folded into immediates
// by the compiler, so `fma(p, r, PBR_EXPC[k])` becomes one DFMA with a c[bank][offset] operand.
//
// pbr_exp: x = k ln2 + r, |r| <= ln2/2, exp(r) by the degree-13 Taylor polynomial (truncation 4e-18), 2^k applied in two
// exact halves so that results in the subnormal range and near overflow are formed without a branch.  Measured against
// libm over [-700, 700]: max error 1.0 ulp.  NaN propagates; x > 709.78 gives +inf, x < -745.2 gives 0.
#pragma once

PBR_CTABLE double PBR_EXPC[20] = {
    1.0, 1.0, 0.5, 1.0/6.0, 1.0/24.0, 1.0/120.0, 1.0/720.0, 1.0/5040.0, 1.0/40320.0, 1.0/362880.0, 1.0/3628800.0,
    1.0/39916800.0, 1.0/479001600.0, 1.0/6227020800.0,
    1.4426950408889634,            // [14] log2(e)
    6.93147180369123816490e-01,    // [15] ln2 high part (trailing bits zero: k*LN2_HI is exact for |k| < 2^20)
    1.90821492927058770002e-10,    // [16] ln2 low part
    6755399441055744.0,            // [17] 1.5 * 2^52: adding it rounds to the nearest integer
    2.302585092994046,             // [18] ln(10)
    0.0};

// PBR_EXP_MODE selects the exponential:
//   0  libdevice exp()/exp10()                      (default)
//   1  Taylor-13 on |r| <= ln2/2 with constant-bank coefficients (round 1: +2 % on K2, -7 % on K3: register pressure)
//   2  table-driven (round 2; measured in one gpurun call against 0 / 1: K2 8.62 vs 8.40 / 8.33 ms, K3 0.364 vs 0.358 / 0.358 ms,
//      K4b 156.5 vs 158.6 / 167.6 ms -- the kernels are latency-bound, not instruction-bound, so it does not pay): x = (32k + j) ln2/32 + r, |r| <= ln2/64, exp(x) = 2^k * T[j] * (1 + q(r)) with a 32-entry table
//      of 2^(j/32) in global memory (L1-resident; the index differs per lane, so the constant bank would serialise) and a
//      degree-6 polynomial: 12 FP64 instructions instead of ~21, no 64-bit immediates.  |x| >= 708, inf and NaN take the
//      library function in an out-of-line slow path (one predicated branch on the fast path).  Max error 1 ulp against libm
//      (host replica of the algorithm over [-700, 700], tests/test_gpu_math.py checks the device version).
#ifndef PBR_EXP_MODE
#ifdef PBR_CUSTOM_EXP
#define PBR_EXP_MODE (PBR_CUSTOM_EXP ? 1 : 0)
#else
#define PBR_EXP_MODE 0
#endif
#endif
#if defined(PBR_HOST) || PBR_EXP_MODE == 0
PBR_DEV double pbr_exp(double x) { return exp(x); }
#ifdef PBR_HOST
PBR_DEV double pbr_exp10(double x) { return pow(10.0, x); }
#else
PBR_DEV double pbr_exp10(double x) { return exp10(x); }
#endif
#elif PBR_EXP_MODE == 2
__device__ double PBR_EXPT[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
    0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
    0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
    0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};
PBR_CTABLE double PBR_EXPC2[12] = {
    46.166241308446828,             // [0] 32/ln2
    6755399441055744.0,             // [1] 1.5 * 2^52: adding it rounds to the nearest integer
    -6.93147180369123816490e-01/32, // [2] -(ln2 high part)/32 (trailing bits zero: n*this is exact for |n| < 2^20)
    -1.90821492927058770002e-10/32, // [3] -(ln2 low part)/32
    1.0/6.0, 1.0/24.0, 1.0/120.0, 1.0/720.0,    // [4..7]
    2.302585092994046,          // [8] ln(10) high
    -2.1707562233822494e-16,       // [9] ln(10) low
    0.0, 0.0};
__device__ __noinline__ double pbr_exp_slow(double x) { return exp(x); }
PBR_DEV double pbr_exp(double x) {
    const double t = fma(x, PBR_EXPC2[0], PBR_EXPC2[1]);
    const double nf = t - PBR_EXPC2[1];
    double r = fma(nf, PBR_EXPC2[2], x);
    r = fma(nf, PBR_EXPC2[3], r);
    const int n = __double2loint(t);                              // 32k + j in the low mantissa bits (two's complement)
    const double Tj = __ldg(&PBR_EXPT[n & 31]);
    double q = fma(r, PBR_EXPC2[7], PBR_EXPC2[6]);
    q = fma(q, r, PBR_EXPC2[5]);
    q = fma(q, r, PBR_EXPC2[4]);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = q*r;                                                      // exp(r) - 1
    double y = fma(Tj, q, Tj);                                    // in [1, 2): scaling by 2^k is an add on the exponent field
    y = __hiloint2double(__double2hiint(y) + ((n << 15) & 0xfff00000), __double2loint(y));      // normal result for |x| < 708
    if ((__double2hiint(x) & 0x7fffffff) >= 0x40862000) y = pbr_exp_slow(x);     // |x| >= 708, inf, NaN
    return y;
}
// 10^x = exp(x ln10): the product is formed in two parts so that the argument's rounding error (|x ln10| ulps) does not reach the result
PBR_DEV double pbr_exp10(double x) {
    const double p = x*PBR_EXPC2[8];
    const double e = fma(x, PBR_EXPC2[8], -p) + x*PBR_EXPC2[9];
    const double y = pbr_exp(p);
    return fma(y, e, y);
}
#else
PBR_DEV double pbr_exp(double x) {
    const double xc = fmin(fmax(x, -746.0), 710.0);              // keeps k in range; NaN passes through fmin/fmax below
    const double t = fma(xc, PBR_EXPC[14], PBR_EXPC[17]);
    const double kf = t - PBR_EXPC[17];
    double r = fma(kf, -PBR_EXPC[15], xc);
    r = fma(kf, -PBR_EXPC[16], r);
    double p = PBR_EXPC[13];
#pragma unroll
    for (int i = 12; i >= 0; --i) p = fma(p, r, PBR_EXPC[i]);
    // k sits in the low mantissa bits of t (two's complement): no double->int conversion (a slow-pipe instruction)
#ifdef PBR_HOST
    union { double d; long long i; } tb; tb.d = t;
    const int k = (int)(unsigned)(tb.i & 0xffffffffll);
    const int k1 = k >> 1, k2 = k - k1;
    union { long long i; double d; } s1, s2;
    s1.i = (long long)(k1 + 1023) << 52; s2.i = (long long)(k2 + 1023) << 52;
    const double r1 = s1.d, r2 = s2.d;
#else
    const int k = __double2loint(t);
    const int k1 = k >> 1, k2 = k - k1;
    const double r1 = __hiloint2double((k1 + 1023) << 20, 0);
    const double r2 = __hiloint2double((k2 + 1023) << 20, 0);
#endif
    const double y = (p*r1)*r2;
    return (x != x) ? x : y;                                     // fmin/fmax drop a NaN argument: put it back
}

PBR_DEV double pbr_exp10(double x) { return pbr_exp(x*PBR_EXPC[18]); }
#endif

// pbr_rcp: 1/x without the IEEE division routine.  A double-precision divide compiles to MUFU.RCP64H + Newton steps +
// range check + a branch to an out-of-line fix-up (~15 issued instructions and a basic-block split per divide; the RHS
// evaluates ~12-25 of them).  Here: rcp.approx.ftz.f64 (about 20 good bits) and two Newton steps in FMA form -- 6
// instructions, branch-free, result within 1 ulp for normal x.  x = 0 or subnormal gives NaN/inf garbage (the reference
// raises ZeroDivisionError there); NaN propagates.  PBR_FAST_RCP=0 falls back to the IEEE divide.
#ifndef PBR_FAST_RCP
#define PBR_FAST_RCP 1
#endif
PBR_DEV double pbr_rcp(double x) {
#if defined(PBR_HOST) || !PBR_FAST_RCP
    return 1.0/x;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
#endif
}

// Reduced-precision reciprocals for quantities that only steer the iteration, never the solution:
//   pbr_rcp_20bit: the bare rcp.approx.ftz.f64 (relative error ~1e-6) -- error-norm weights 1/(atol + rtol|y|)
//   pbr_rcp_40bit: one Newton step (~1e-12) -- LU pivots and Jacobian-only factors of the Newton matrix
PBR_DEV double pbr_rcp_20bit(double x) {
#if defined(PBR_HOST) || !PBR_FAST_RCP
    return 1.0/x;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
#endif
}
PBR_DEV double pbr_rcp_40bit(double x) {
#if defined(PBR_HOST) || !PBR_FAST_RCP
    return 1.0/x;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return fma(fma(-x, r, 1.0), r, r);
#endif
}

// pbr_log / pbr_sqrt: branch-free replacements for libdevice's log() / sqrt() (each of those carries a range check and a
// branch to an out-of-line slow path).  PBR_FAST_LOG=0 falls back to the library versions.
//   log:  x = m 2^e with m in [sqrt(1/2), sqrt(2)), s = (m-1)/(m+1), log m = 2s(1 + z/3 + z^2/5 + ... + z^10/21), z = s^2
//         (|s| <= 0.1716, truncation 2e-17); e ln2 added in two parts.  Normal positive x: <= 1.5 ulp (host test).
//         x = 0 -> -inf, x < 0 or NaN -> NaN, +inf -> +inf by selects; subnormal x is treated as 0.
//   sqrt: rsqrt.approx.ftz.f64 + two Newton steps + one correction of the product; x = 0 -> 0.
#ifndef PBR_FAST_LOG
#define PBR_FAST_LOG 1
#endif
PBR_CTABLE double PBR_LOGC[14] = {
    2.0/3.0, 2.0/5.0, 2.0/7.0, 2.0/9.0, 2.0/11.0, 2.0/13.0, 2.0/15.0, 2.0/17.0, 2.0/19.0, 2.0/21.0,
    6.93147180369123816490e-01, 1.90821492927058770002e-10,      // [10] ln2 hi, [11] ln2 lo
    0.4342944819032518,                                          // [12] log10(e)
    0.0};

PBR_DEV double pbr_log(double x) {
#if defined(PBR_HOST) || !PBR_FAST_LOG
    return log(x);
#else
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;                         // m in [1, 2)
    const int big = (hi >= 0x3ff6a09e) ? 1 : 0;                  // m >= ~sqrt(2): halve it
    hi -= big << 20; e += big;
    const double m = __hiloint2double(hi, lo);
    const double f = m - 1.0;
    const double s = f*pbr_rcp(2.0 + f);
    const double z = s*s;
    double pz = PBR_LOGC[9];
#pragma unroll
    for (int k = 8; k >= 0; --k) pz = fma(pz, z, PBR_LOGC[k]);
    const double ef = (double)e;
    double r = fma(s*z, pz, ef*PBR_LOGC[11]);                    // s z P(z) + e ln2_lo
    r = r + 2.0*s;
    r = fma(ef, PBR_LOGC[10], r);
    r = (x > 1.7976931348623157e308) ? x : r;                    // +inf
    r = (x < 2.2250738585072014e-308) ? ((x < 0.0) ? __longlong_as_double(0x7ff8000000000000ll) : -__longlong_as_double(0x7ff0000000000000ll)) : r;
    return (x != x) ? x : r;
#endif
}
PBR_DEV double pbr_log10(double x) {
#if defined(PBR_HOST) || !PBR_FAST_LOG
    return log10(x);
#else
    return pbr_log(x)*PBR_LOGC[12];
#endif
}
PBR_DEV double pbr_sqrt(double x) {
#if defined(PBR_HOST) || !PBR_FAST_LOG
    return sqrt(x);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = y*fma(-0.5*x, y*y, 1.5);
    y = y*fma(-0.5*x, y*y, 1.5);
    double r = x*y;
    r = fma(fma(-r, r, x), 0.5*y, r);
    return (x == 0.0) ? 0.0 : r;
#endif
}

// ---- single-precision helpers of the mixed-precision Jacobian (pbr_rates_jacf): the partial derivatives only build the Newton
//      matrix, so the hardware approximations (MUFU, ~2 ulp) are what is wanted.
#ifdef PBR_HOST
PBR_DEV float pbr_rcpf(float x) { return 1.0f/x; }
PBR_DEV float pbr_expf(float x) { return expf(x); }
PBR_DEV float pbr_logf(float x) { return logf(x); }
PBR_DEV float pbr_log10f(float x) { return log10f(x); }
PBR_DEV float pbr_sqrtf(float x) { return sqrtf(x); }
PBR_DEV float pbr_powf(float x, float y) { return powf(x, y); }
#else
PBR_DEV float pbr_rcpf(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
PBR_DEV float pbr_expf(float x) { return __expf(x); }
PBR_DEV float pbr_logf(float x) { return __logf(x); }
PBR_DEV float pbr_log10f(float x) { return __log10f(x); }
PBR_DEV float pbr_sqrtf(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
PBR_DEV float pbr_powf(float x, float y) { return __powf(x, y); }
#endif
