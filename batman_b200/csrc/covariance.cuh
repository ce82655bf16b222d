// Stationary covariance factors, rounded exactly like the numpy expressions of
// sklearn:kernels.py:1716-1730 (Matern) and 1557-1563 (RBF): explicit _rn
// intrinsics so that nvcc cannot contract a mul+add into an FMA where numpy
// rounds twice.  The squared scaled distance `s` is accumulated by the caller
// in scipy cdist order (sequential, non-fused s += d*d).
#pragma once
#include "../../include/batman_b200.h"
#include "fastmath.cuh"
#include "model.h"

namespace bk {

constexpr double SQRT3 = 1.7320508075688772;   // math.sqrt(3)
constexpr double SQRT5 = 2.23606797749979;     // math.sqrt(5)

template <int KIND>
__device__ __forceinline__ double stationary(double s) {
    if constexpr (KIND == BK_MATERN12) {
        return bk_exp_neg(-bk_sqrt(s));                                       // np.exp(-dists)
    } else if constexpr (KIND == BK_MATERN32) {
        double t = __dmul_rn(bk_sqrt(s), SQRT3);                       // K = dists * math.sqrt(3)
        return __dmul_rn(__dadd_rn(1.0, t), bk_exp_neg(-t));        // (1.0 + K) * np.exp(-K)
    } else if constexpr (KIND == BK_MATERN52) {
        double t = __dmul_rn(bk_sqrt(s), SQRT5);                       // K = dists * math.sqrt(5)
        double q = __ddiv_rn(__dmul_rn(t, t), 3.0);                 // K**2 / 3.0
        return __dmul_rn(__dadd_rn(__dadd_rn(1.0, t), q), bk_exp_neg(-t)); // (1.0 + K + K**2/3.0) * np.exp(-K)
    } else if constexpr (KIND == BK_RBF) {
        return bk_exp_neg(__dmul_rn(-0.5, s));                      // np.exp(-0.5 * sqeuclidean)
    } else {
        double d = bk_sqrt(s);
        return bk_exp_neg(-__dmul_rn(__dmul_rn(d, d), 0.5));               // np.exp(-(dists**2) / 2.0)
    }
}

// k = c0 + c1 * f : ConstantKernel * stat (c0 = 0), stat alone (c1 = 1), ConstantKernel + stat (c1 = 1)
__device__ __forceinline__ double affine(double c0, double c1, double f) {
    return __dadd_rn(c0, __dmul_rn(c1, f));
}

// runtime kind (warp-uniform branch): the general kernels and the training-covariance builder
__device__ __forceinline__ double stationary_rt(int kind, double s) {
    switch (kind) {
        case BK_MATERN12: return stationary<BK_MATERN12>(s);
        case BK_MATERN32: return stationary<BK_MATERN32>(s);
        case BK_MATERN52: return stationary<BK_MATERN52>(s);
        case BK_RBF: return stationary<BK_RBF>(s);
        default: return stationary<BK_MATERN_INF>(s);
    }
}

// k = sum_t coef_t * prod_f fv[f]^texp[t][f]: Sum adds, Product multiplies, every operation rounded like the numpy
// expression of the tree evaluated left to right (sklearn:kernels.py:870, 968).  A term starts from its first
// operand -- 1.0 * f is exact, so a unit coefficient costs no rounding either way.
__device__ __forceinline__ double general_cov(const GenSpec& g, const double (&fv)[GEN_MAX_F]) {
    double k = 0.0;
#pragma unroll
    for (int t = 0; t < GEN_MAX_T; ++t) {
        if (t < g.nt) {
            double term = g.tcoef[t];
#pragma unroll
            for (int f = 0; f < GEN_MAX_F; ++f)
                if (f < g.nf)
                    for (int e = 0; e < g.texp[t][f]; ++e) term = __dmul_rn(term, fv[f]);
            k = t == 0 ? term : __dadd_rn(k, term);
        }
    }
    return k;
}

}  // namespace bk
