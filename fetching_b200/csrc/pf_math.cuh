// pf_math.cuh -- transcendental kernels sized for the FP64 pipe of sm_100a.
//
// The mixture log-likelihood (pymixture.py:86-92) costs K exps and one log per (row, node).  With
// the CUDA math library those are ~35-45 FP64 instructions each, every call carries a divergent
// special-case region, and the kernel ends up ISSUE / latency bound.  These versions are
// branch-free on the hot path, keep <= ~2 ulp accuracy (far inside the 1e-10 budget) with 11 / ~24
// FP64 instructions and short dependency chains, do exponent surgery and range checks on the integer pipe, and take their
// coefficients from the constant bank (a DFMA source operand, no materialising moves).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pf {

// 2^(j/32), j = 0..31 (copied into shared memory by the kernels: per-lane indexed loads)
__constant__ double kExp2Tab[32] = {
    1.0,
    1.0218971486541166,
    1.0442737824274138,
    1.0671404006768237,
    1.0905077326652577,
    1.1143867425958924,
    1.1387886347566916,
    1.1637248587775775,
    1.189207115002721,
    1.215247359980469,
    1.241857812073484,
    1.2690509571917332,
    1.2968395546510096,
    1.3252366431597413,
    1.3542555469368927,
    1.383909881963832,
    1.4142135623730951,
    1.4451808069770467,
    1.4768261459394993,
    1.5091644275934228,
    1.5422108254079407,
    1.5759808451078865,
    1.6104903319492543,
    1.645755478153965,
    1.681792830507429,
    1.718619298122478,
    1.7562521603732995,
    1.7947090750031072,
    1.8340080864093424,
    1.8741676341103,
    1.9152065613971474,
    1.9571441241754002,
};
// ln2^i / i!, i = 1..6: Taylor coefficients of 2^g - 1 (truncation 3.5e-18 relative on |g| <= 1/64)
__constant__ double kExp2C[6] = {0.6931471805599453, 0.2402265069591007, 0.055504108664821576, 0.009618129107628477, 0.0013333558146428441, 0.00015403530393381606};
// 1/(2i+1), i = 1..10: odd atanh series
__constant__ double kAtanhC[10] = {1.0 / 3.0,  1.0 / 5.0,  1.0 / 7.0,  1.0 / 9.0,  1.0 / 11.0,
                                   1.0 / 13.0, 1.0 / 15.0, 1.0 / 17.0, 1.0 / 19.0, 1.0 / 21.0};
__constant__ double kMathC[4] = {-0.72134752044448170368 /* -log2(e)/2 */,
                                 211106232532992.0 /* 1.5 * 2^47: rint(32 z) lands in the low word */,
                                 6.93147180369123816490e-01 /* ln2 hi */, 1.90821492927058770002e-10 /* ln2 lo */};

// exp(-e/2) for e >= 0, BRANCH FREE, 11 FP64 instructions.
//   2^z,  z = -e*log2(e)/2 = k + j/32 + g,  |g| <= 1/64:
//   2^z = 2^k * tab[j] * (1 + g*(c1 + g*(c2 + ... + c6 g^5)))
// k and j come out of one magic-number add (integer pipe afterwards), tab[] is a 32-entry
// shared-memory table, 2^k goes through the exponent field.  |z| >= 1021 (result would be
// subnormal), inf and nan are flushed to 0: callers re-do such rows on the slow path.
__device__ __forceinline__ double exp_neg_half_fast(double e, const double* tab) {
    const double z = e * kMathC[0];
    const double t = z + kMathC[1];
    const int ki = __double2loint(t);  // rint(32 z) = 32 k + j
    const double g = z - (t - kMathC[1]);
    double q = kExp2C[5];
#pragma unroll
    for (int i = 4; i >= 0; --i) q = fma(q, g, kExp2C[i]);
    const double tj = tab[ki & 31];
    const double p = fma(q * g, tj, tj);  // tab[j] * (1 + g q)
    const bool flush = (unsigned) (__double2hiint(z) & 0x7FFFFFFF) >= 0x408FE800u;  // integer pipe
    const int hi = flush ? 0 : __double2hiint(p) + (int) ((unsigned) (ki >> 5) << 20);
    const int lo = flush ? 0 : __double2loint(p);
    return __hiloint2double(hi, lo);
}

// G independent exps, written step-interleaved so the FP64 pipe always has independent work.
template <int G>
__device__ __forceinline__ void exp_neg_half_fast_batch(const double (&e)[G], double (&out)[G], const double* tab) {
    double z[G], g[G], q[G], tj[G];
    int ki[G];
#pragma unroll
    for (int c = 0; c < G; ++c) {
        z[c] = e[c] * kMathC[0];
        const double t = z[c] + kMathC[1];
        ki[c] = __double2loint(t);
        g[c] = z[c] - (t - kMathC[1]);
        tj[c] = tab[ki[c] & 31];
        q[c] = kExp2C[5];
    }
#pragma unroll
    for (int i = 4; i >= 0; --i)
#pragma unroll
        for (int c = 0; c < G; ++c) q[c] = fma(q[c], g[c], kExp2C[i]);
#pragma unroll
    for (int c = 0; c < G; ++c) {
        const double p = fma(q[c] * g[c], tj[c], tj[c]);
        const bool flush = (unsigned) (__double2hiint(z[c]) & 0x7FFFFFFF) >= 0x408FE800u;
        const int hi = flush ? 0 : __double2hiint(p) + (int) ((unsigned) (ki[c] >> 5) << 20);
        const int lo = flush ? 0 : __double2loint(p);
        out[c] = __hiloint2double(hi, lo);
    }
}

template <int G>
__device__ __forceinline__ void exp_neg_half_fast_batch(const float (&e)[G], float (&out)[G], const double*) {
#pragma unroll
    for (int c = 0; c < G; ++c) out[c] = exp2f(e[c] * -0.72134752f);
}

__device__ __forceinline__ float exp_neg_half_fast(float e, const double*) {
    return exp2f(e * -0.72134752f);  // MUFU.EX2; underflows to 0 by itself
}

// true when a mixture row sum needs the slow path: it is tiny (< 2^-959: flushed terms could
// matter), zero, negative, inf or nan.
__device__ __forceinline__ bool needs_slow_path(double lp) {
    return (unsigned) (__double2hiint(lp) - 0x04000000) >= 0x7BF00000u;
}
__device__ __forceinline__ bool needs_slow_path(float lp) { return !(lp >= 1e-30f && lp < 3e38f); }

// log(x) for x > 0 normal (callers route everything else through needs_slow_path / log()).
// x = 2^n m, m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716,
// odd series through s^21 (truncation < 1e-18 relative).
__device__ __forceinline__ double log_pos_fast(double x) {
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int n = (hi >> 20) - 1023;
    hi = (hi & 0x000FFFFF) | 0x3FF00000;  // m in [1, 2)
    const bool big = hi >= 0x3FF6A09F;    // m >= ~sqrt(2): halve
    hi = big ? hi - 0x00100000 : hi;
    n = big ? n + 1 : n;
    const double m = __hiloint2double(hi, lo);
    const double num = m - 1.0, den = m + 1.0;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));  // ~2^-23; two Newton steps -> full precision
    double err = fma(-den, r, 1.0);
    r = fma(r, err, r);
    err = fma(-den, r, 1.0);
    r = fma(r, err, r);
    const double s = num * r;
    const double s2 = s * s;
    // q(s2) = sum_{i<10} s2^i / (2i+3), evaluated as A(s4) + s2 B(s4): two independent Horner chains
    const double s4 = s2 * s2;
    double qa = kAtanhC[8], qb = kAtanhC[9];
#pragma unroll
    for (int i = 6; i >= 0; i -= 2) {
        qa = fma(qa, s4, kAtanhC[i]);
        qb = fma(qb, s4, kAtanhC[i + 1]);
    }
    const double q = fma(qb, s2, qa);
    const double s3 = s * s2;
    const double logm = fma(s3 + s3, q, s + s);  // 2s + 2 s^3 (1/3 + s^2/5 + ...)
    const double nd = (double) n;
    return fma(nd, kMathC[2], fma(nd, kMathC[3], logm));  // n ln2 (hi + lo)
}

__device__ __forceinline__ float log_pos_fast(float x) {
    return __logf(x);  // MUFU.LG2 * ln2
}

// full-range wrappers (diagnostics and slow paths)
__device__ __forceinline__ double exp_neg_half(double e, const double* tab) {
    const double z = e * -0.72134752044448170368;
    if ((unsigned) (__double2hiint(z) & 0x7FFFFFFF) >= 0x408FE800u) return exp(-0.5 * e);
    return exp_neg_half_fast(e, tab);
}
__device__ __forceinline__ float exp_neg_half(float e, const double* tab) { return exp_neg_half_fast(e, tab); }
__device__ __forceinline__ double log_pos(double x) {
    if ((unsigned) (__double2hiint(x) - 0x00100000) >= 0x7FE00000u) return log(x);  // 0, subnormal, inf, nan, < 0
    return log_pos_fast(x);
}
__device__ __forceinline__ float log_pos(float x) { return log_pos_fast(x); }

}  // namespace pf
