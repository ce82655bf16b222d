// K6 + K1 + K3: Saltelli design -> surrogate mean -> compensated Sobol' sums, fused.
//
// Replaces, for every base row k of the experiment,
//   ot.SobolIndicesExperiment(dist, N, True).generate()        (uq/uq.py:331-332)
//   output_design = UQ.func(input_design)  [N(2p+2) predictions] (uq/uq.py:333, 192-203)
//   the dot products of ot.SaltelliSensitivityAlgorithm          (uq/uq.py:342-349, 367-375)
//   output_design.var(axis=0)                                    (uq/uq.py:403)
// One thread owns one base row k: it holds A_k and B_k (read, or generated
// in-kernel with Philox4x32-10), forms the per-dimension squared terms
// ta_c = ((A_kc - X_jc)/l_c)^2 and tb_c once per training point and re-sums
// them per block (A, B, E^i = A with column i from B, C^i = B with column i
// from A) in scipy-cdist order, so 2p+2 distances cost 4P + (2p+2)P DP
// instructions instead of (2p+2)*3P.  Outputs never reach HBM: per row they
// are parked in shared memory, multiplied into the (5 + 6p [+ 2p + p(p-1)/2])
// Saltelli/Jansen products (plus the row count, the last sum), tree-reduced across the warp with shuffles and
// accumulated as (hi, lo) double-double sums; each CTA owns one partial slot
// (no atomics, bit-reproducible for a fixed grid).
#pragma once
#include "common.cuh"
#include "covariance.cuh"
#include "model.h"

namespace bk {

constexpr int SOB_THREADS = 128;
constexpr int SOB_TJ = 128;       // training rows per stage
constexpr int SOB_CH = 6;         // design blocks evaluated per pass over the training set
constexpr int SOB_SLOTS = 148 * 4;

__host__ __device__ inline int sobol_nsums(int p, int second) {
    // ... + 1: the number of base rows behind the sums (last entry; the delete-a-group jackknife needs it per slot)
    return 5 + 6 * p + ((second && p > 2) ? 2 * p + p * (p - 1) / 2 : 0) + 1;
}
__host__ __device__ inline int sobol_nblocks(int p, int second) { return (second && p > 2) ? 2 * p + 2 : p + 2; }

// ---------------------------------------------------------------- Philox4x32-10 (oracle/philox.py is the CPU twin)
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double unit_from_words(uint32_t hi, uint32_t lo) {
    const uint64_t k = ((uint64_t)(hi >> 6) << 26) | (uint64_t)(lo >> 6);
    return __dmul_rn(__dadd_rn((double)k, 0.5), 0x1p-52);
}
// ---------------------------------------------------------------- inverse CDFs of the marginals (bk_marginal_kind)
// log Beta(a, b) prefactor and continued fraction of the regularised incomplete beta function (modified Lentz,
// DLMF 8.17.22); converges in O(sqrt(max(a, b))) terms for x < (a + 1) / (a + b + 2).
static __device__ __noinline__ double beta_cf(double a, double b, double x) {
    const double tiny = 1e-300, qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 600; ++m) {
        const double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return h;
}
static __device__ __noinline__ double beta_inc(double a, double b, double x, double lbeta) {   // I_x(a, b)
    if (x <= 0.0) return 0.0;
    if (x >= 1.0) return 1.0;
    const double bt = exp(a * log(x) + b * log1p(-x) - lbeta);
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * beta_cf(a, b, x) / a;
    return 1.0 - bt * beta_cf(b, a, 1.0 - x) / b;
}
// x with I_x(a, b) = u: AS 109 starting value, then safeguarded Halley steps on I_x - u (cubic convergence; the
// bracket [lo, hi] keeps every iterate inside (0, 1))
static __device__ __noinline__ double beta_inc_inv(double a, double b, double u) {
    const double lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
    double x;
    if (a >= 1.0 && b >= 1.0) {
        const double pp = u < 0.5 ? u : 1.0 - u;
        const double t = sqrt(-2.0 * log(pp));
        double z = (2.30753 + t * 0.27061) / (1.0 + t * (0.99229 + t * 0.04481)) - t;
        if (u < 0.5) z = -z;
        const double al = (z * z - 3.0) / 6.0;
        const double h = 2.0 / (1.0 / (2.0 * a - 1.0) + 1.0 / (2.0 * b - 1.0));
        const double w = z * sqrt(al + h) / h - (1.0 / (2.0 * b - 1.0) - 1.0 / (2.0 * a - 1.0)) * (al + 5.0 / 6.0 - 2.0 / (3.0 * h));
        x = a / (a + b * exp(2.0 * w));
    } else {
        const double lna = log(a / (a + b)), lnb = log(b / (a + b));
        const double t = exp(a * lna) / a, t1 = exp(b * lnb) / b, w = t + t1;
        x = u < t / w ? pow(a * w * u, 1.0 / a) : 1.0 - pow(b * w * (1.0 - u), 1.0 / b);
    }
    double lo = 0.0, hi = 1.0;
    if (!(x > 0.0 && x < 1.0)) x = 0.5;
    for (int it = 0; it < 60; ++it) {
        const double err = beta_inc(a, b, x, lbeta) - u;
        if (err > 0.0) hi = x; else lo = x;
        const double pdf = exp((a - 1.0) * log(x) + (b - 1.0) * log1p(-x) - lbeta);
        double xn;
        if (pdf > 0.0 && isfinite(pdf)) {
            const double st = err / pdf;
            double corr = 0.5 * st * ((a - 1.0) / x - (b - 1.0) / (1.0 - x));
            if (corr > 0.5) corr = 0.5;
            if (corr < -0.5) corr = -0.5;
            xn = x - st / (1.0 - corr);
        } else {
            xn = 0.5 * (lo + hi);
        }
        if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);              // left the bracket: bisect
        const double dx = fabs(xn - x);
        x = xn;
        if (dx <= 4e-16 * x || hi - lo <= 4e-16 * x) break;
    }
    return x;
}

// par = the BK_MARGINAL_NPAR parameters of the marginal (include/batman_b200.h)
__device__ __forceinline__ double marginal(int kind, double a, double b, double c, double d, double u) {
    if (kind == BK_UNIFORM) return __dadd_rn(a, __dmul_rn(__dadd_rn(b, -a), u));   // a + (b - a) * u
    if (kind == BK_NORMAL) return __dadd_rn(a, __dmul_rn(b, normcdfinv(u)));       // mu + sigma * ndtri(u)
    switch (kind) {
        case BK_LOGNORMAL: return c + exp(a + b * normcdfinv(u));
        case BK_TRIANGULAR: {
            const double fm = (b - a) / (c - a);
            return u < fm ? a + sqrt(u * (c - a) * (b - a)) : c - sqrt((1.0 - u) * (c - a) * (c - b));
        }
        case BK_TRUNCNORMAL: return a + b * normcdfinv(c + u * (d - c));
        case BK_BETA: return c + (d - c) * beta_inc_inv(a, b, u);
        case BK_GUMBEL: return b - a * log(-log(u));
        case BK_EXPONENTIAL: return b - log1p(-u) / a;
        case BK_LOGUNIFORM: return exp(a + (b - a) * u);
        case BK_WEIBULLMIN: return c + a * pow(-log1p(-u), 1.0 / b);
        case BK_LOGISTIC: return a + b * (log(u) - log1p(-u));
        case BK_RAYLEIGH: return b + a * sqrt(-2.0 * log1p(-u));
    }
    return u;
}
// marginal parameters as the kernels carry them: par[k][column]
template <int P>
struct MargPar {
    int kind[P];
    double par[BK_MARGINAL_NPAR][P];
};
template <int P>
__host__ inline void marg_fill(MargPar<P>& mp, int p, const int* mkind, const double* mpar) {
    for (int c = 0; c < P; ++c) {
        mp.kind[c] = (c < p && mkind) ? mkind[c] : 0;
        for (int k = 0; k < BK_MARGINAL_NPAR; ++k) mp.par[k][c] = (c < p && mpar) ? mpar[BK_MARGINAL_NPAR * c + k] : 0.0;
    }
}
template <int P>
__device__ __forceinline__ double marginal_col(const MargPar<P>& mp, int c, double u) {
    return marginal(mp.kind[c], mp.par[0][c], mp.par[1][c], mp.par[2][c], mp.par[3][c], u);
}
template <int P>
__device__ __forceinline__ void design_row(uint64_t seed, uint64_t row, int p, int stream, const MargPar<P>& mp,
                                           double x[P]) {
#pragma unroll
    for (int cp = 0; cp < P / 2; ++cp) {
        if (2 * cp < p) {
            uint32_t w[4];
            philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)cp, (uint32_t)stream, (uint32_t)seed,
                          (uint32_t)(seed >> 32), w);
            x[2 * cp] = marginal_col<P>(mp, 2 * cp, unit_from_words(w[0], w[1]));
            x[2 * cp + 1] = (2 * cp + 1 < p) ? marginal_col<P>(mp, 2 * cp + 1, unit_from_words(w[2], w[3])) : 0.0;
        } else {
            x[2 * cp] = 0.0;
            x[2 * cp + 1] = 0.0;
        }
    }
}

// ---------------------------------------------------------------- warp tree reduction + lane-owned dd accumulation
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sum index r is owned by lane r%32 of each warp: no races, no atomics
__device__ __forceinline__ void emit(double v, int& r, int lane, double* wsum_warp) {
    v = warp_sum(v);
    if (lane == (r & 31)) {
        dd a{wsum_warp[2 * r], wsum_warp[2 * r + 1]};
        dd_add_d(a, v);
        wsum_warp[2 * r] = a.hi;
        wsum_warp[2 * r + 1] = a.lo;
    }
    ++r;
}

// products of one base row; ys(b) returns the shifted output of block b for this thread's row; one = 1.0 for a
// real row, 0.0 for a padding thread
template <typename YS>
__device__ __forceinline__ void saltelli_products(YS ys, int p, int second, int lane, double* wsum_warp, double one) {
    int r = 0;
    const double yA = ys(0), yB = ys(1);
    emit(yA, r, lane, wsum_warp);
    emit(yB, r, lane, wsum_warp);
    emit(__dmul_rn(yA, yA), r, lane, wsum_warp);
    emit(__dmul_rn(yB, yB), r, lane, wsum_warp);
    emit(__dmul_rn(yA, yB), r, lane, wsum_warp);
    for (int i = 0; i < p; ++i) {
        const double yE = ys(2 + i);
        emit(yE, r, lane, wsum_warp);
        emit(__dmul_rn(yE, yE), r, lane, wsum_warp);
        emit(__dmul_rn(yB, yE), r, lane, wsum_warp);
        emit(__dmul_rn(yA, yE), r, lane, wsum_warp);
        const double dB = __dadd_rn(yB, -yE), dA = __dadd_rn(yA, -yE);
        emit(__dmul_rn(dB, dB), r, lane, wsum_warp);      // Jansen first order
        emit(__dmul_rn(dA, dA), r, lane, wsum_warp);      // Jansen total order
    }
    if (second) {
        if (p > 2) {
            for (int j = 0; j < p; ++j) {
                const double yC = ys(2 + p + j);
                emit(yC, r, lane, wsum_warp);
                emit(__dmul_rn(yC, yC), r, lane, wsum_warp);
            }
            for (int i = 0; i < p; ++i)
                for (int j = 0; j < i; ++j) emit(__dmul_rn(ys(2 + i), ys(2 + p + j)), r, lane, wsum_warp);
        }
        // p == 2: C^0 == E^1 and C^1 == E^0, the product sum(yE^1 yC^0) is sum(yE^1 yE^1): no extra sums
    }
    emit(one, r, lane, wsum_warp);                        // row count
}

// CTA epilogue: fixed-order sum of the warps' (hi, lo) sums, ADDED into this CTA's slot
__device__ __forceinline__ void flush_slot(const double* wsum, int nwarps, int nsums, double* slot) {
    for (int r = threadIdx.x; r < nsums; r += blockDim.x) {
        dd t{slot[2 * r], slot[2 * r + 1]};
        for (int w = 0; w < nwarps; ++w) t = dd_add(t, dd{wsum[(w * nsums + r) * 2], wsum[(w * nsums + r) * 2 + 1]});
        slot[2 * r] = t.hi;
        slot[2 * r + 1] = t.lo;
    }
}

template <int P>
struct SobolParams {
    const double* a;       // count x p explicit base samples, or null -> Philox
    const double* b;
    uint64_t seed;
    MargPar<P> marg;
    long row_lo, count;
    int p, second, nb, nsums;
    int n_pad;
    const double* xs;
    const double* alpha;
    double scale[P], mn[P], ls[P];
    double c0, c1, y_mean, y_std, shift;
    double* partial;       // slot 0 of this output; slots are slot_stride doubles apart
    long slot_stride;
    double* yout;          // if set: write the (unshifted) outputs [nb][count][m] instead of reducing them
    int m, q;
};

// statically scheduled variant (sobol_static.cu); -1 = not instantiated for this (kind, P)
template <int P>
int launch_sobol_static(int kind, const SobolParams<P>& prm, size_t smem, cudaStream_t st);

}  // namespace bk
