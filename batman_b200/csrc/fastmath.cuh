// Branch-free FP64 sqrt and exp for the covariance kernels.
//
// CUDA's sqrt()/exp() are accurate but carry slow-path branches; a branch ends the basic block, which stops
// ptxas from interleaving the (long, serially dependent) chains of independent test points / design blocks, and
// ncu shows K1/K3 stalled on fixed-latency "wait" with the FP64 pipe at 60 %.  These versions are straight-line:
//   bk_sqrt     MUFU.RSQ64H seed + 2 Newton steps + Markstein correction: correctly rounded (bit-identical to
//               IEEE sqrt over the tested range, tests/test_gpu_fastmath.py), 10 DP instructions.
//   bk_exp_neg  x <= 0 only (every kernel argument is -distance): Cody-Waite reduction with FMA, degree-12
//               near-minimax polynomial (Chebyshev interpolation, approximation error 0.003 ulp), exponent
//               insertion; <= 1 ulp (measured in the same test), 18 DP instructions.  Results below
//               2^-1021 flush to 0.
#pragma once

namespace bk {

__device__ __forceinline__ double bk_sqrt(double s) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));   // ~2^-20 relative (upper 32 bits of s only)
    double g = __dmul_rn(s, y);
    double h = __dmul_rn(y, 0.5);
    double r = __fma_rn(-g, h, 0.5);
    g = __fma_rn(g, r, g);
    h = __fma_rn(h, r, h);
    r = __fma_rn(-g, h, 0.5);
    g = __fma_rn(g, r, g);
    h = __fma_rn(h, r, h);
    const double d = __fma_rn(-g, g, s);      // exact residual s - g^2
    g = __fma_rn(d, h, g);                    // Markstein: correctly rounded
    return s == 0.0 ? 0.0 : g;                // rsqrt(0) = inf -> NaN above; select, not branch
}

__device__ __forceinline__ double bk_exp_neg(double x) {
    const double t = __fma_rn(x, 1.4426950408889634, 6755399441055744.0);   // x*log2(e) + 1.5*2^52
    const int k = __double2loint(t);                                        // round(x*log2(e))
    const double kf = __dadd_rn(t, -6755399441055744.0);
    double r = __fma_rn(kf, -6.93147180559945286e-01, x);                   // ln2 hi
    r = __fma_rn(kf, -2.31904681384629956e-17, r);                          // ln2 lo
    double p = 2.0914717341716378e-09;
    p = __fma_rn(p, r, 2.5105259538448494e-08);
    p = __fma_rn(p, r, 2.755727357010214e-07);
    p = __fma_rn(p, r, 2.7557255297969643e-06);
    p = __fma_rn(p, r, 2.4801587325605324e-05);
    p = __fma_rn(p, r, 0.0001984126987490126);
    p = __fma_rn(p, r, 0.001388888888888373);
    p = __fma_rn(p, r, 0.008333333333326112);
    p = __fma_rn(p, r, 0.04166666666666667);
    p = __fma_rn(p, r, 0.1666666666666667);
    p = __fma_rn(p, r, 0.5);
    p = __fma_rn(p, r, 1.0);
    p = __fma_rn(p, r, 1.0);
    const double res = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));   // p * 2^k
    return x < -708.0 ? 0.0 : res;
}

}  // namespace bk
