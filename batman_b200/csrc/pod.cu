// K4: POD mode -> field step on the FP64 tensor cores, alone and fused with the Saltelli sums.
//
// Replaces  pred = mean_snapshot + (U @ modes.T).T        (batman/pod/pod.py:226-228)
// as applied by SurrogateModel.__call__ to the mean AND to sigma (surrogate_model.py:191-196), and, in
// UQ.sobol with a POD surrogate, the N(2p+2) x F output design that the reference materialises before
// ot.SaltelliSensitivityAlgorithm reads it (uq.py:333-343).  The GEMM is (rows x m) . (m x F) with
// m <= ~20 modes and F ~ 400 features: each warp owns one 8-feature tile (U fragment in registers) and
// walks 8-row tiles with mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).
//
// k_pod_inverse : stores the field (HBM-bound on the k x F output).
// k_pod_sobol   : rows are Saltelli groups; per group and feature the DMMA produces the outputs of all
//                 design blocks in the SAME lane (one accumulator pair per block), so the Saltelli/Jansen
//                 products are formed in registers, tree-reduced across the 8 groups of the fragment with
//                 shuffles and accumulated as (hi, lo) sums in shared memory.  The field never exists.
//                 'block' indices (uq.py:205-218) are the same kernel with F = 1 and U replaced by the
//                 trapezoid weights applied to U (the integral of the field is linear in the modes).
#include "common.cuh"
#include "model.h"

namespace bk {

int sobol_nsums_host(int p, int second);
int sobol_nslots_host();

constexpr int POD_MAX_KSTEPS = 16;   // m <= 64 modes: the U fragment of a feature tile lives in registers; beyond that
                                     // (no limit in the reference, pod.py:226-228) it is re-read from L1/L2 per row tile

// ---------------------------------------------------------------- field = mean + modes . U^T
__global__ void __launch_bounds__(128) k_pod_inverse(const double* __restrict__ modes, long k, int m,
                                                     const double* __restrict__ U, const double* __restrict__ mean,
                                                     int F, double* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tq = lane & 3;
    const int ftiles = (F + 7) / 8, ksteps = (m + 3) / 4;
    const long rtiles = (k + 7) / 8;
    for (int ft = blockIdx.y * 4 + warp; ft < ftiles; ft += gridDim.y * 4) {
        const int fb = ft * 8 + g;                // feature whose U row this lane holds as B fragment
        double b[POD_MAX_KSTEPS];
#pragma unroll
        for (int ks = 0; ks < POD_MAX_KSTEPS; ++ks) {
            const int q = ks * 4 + tq;
            b[ks] = (ks < ksteps && fb < F && q < m) ? U[(size_t)fb * m + q] : 0.0;
        }
        const int f0 = ft * 8 + 2 * tq;           // features of this lane's accumulators
        const double m0 = f0 < F ? mean[f0] : 0.0, m1 = f0 + 1 < F ? mean[f0 + 1] : 0.0;
        for (long rt = blockIdx.x; rt < rtiles; rt += gridDim.x) {
            const long row = rt * 8 + g;
            double c0 = m0, c1 = m1;
#pragma unroll
            for (int ks = 0; ks < POD_MAX_KSTEPS; ++ks) {
                if (ks < ksteps) {
                    const int q = ks * 4 + tq;
                    const double a = (row < k && q < m) ? modes[(size_t)row * m + q] : 0.0;
                    dmma884(c0, c1, a, b[ks]);
                }
            }
            for (int ks = POD_MAX_KSTEPS; ks < ksteps; ++ks) {          // m > 64: same order, U from the cache
                const int q = ks * 4 + tq;
                const double a = (row < k && q < m) ? modes[(size_t)row * m + q] : 0.0;
                dmma884(c0, c1, a, (fb < F && q < m) ? U[(size_t)fb * m + q] : 0.0);
            }
            if (row < k) {
                if (f0 + 1 < F && (F & 1) == 0) {
                    *reinterpret_cast<double2*>(out + (size_t)row * F + f0) = make_double2(c0, c1);
                } else {
                    if (f0 < F) out[(size_t)row * F + f0] = c0;
                    if (f0 + 1 < F) out[(size_t)row * F + f0 + 1] = c1;
                }
            }
        }
    }
}

int launch_pod_inverse(const double* modes, long k, int m, const double* U, const double* mean, int F, double* out,
                       cudaStream_t st) {
    if (k <= 0) return 0;
    const int ftiles = (F + 7) / 8;
    const int gy = (ftiles + 3) / 4 < 8 ? (ftiles + 3) / 4 : 8;
    long rtiles = (k + 7) / 8;
    const int gx = (int)(rtiles < 148 * 8 / gy ? rtiles : 148 * 8 / gy);
    ProfScope prof(PROF_POD, st);
    k_pod_inverse<<<dim3(gx < 1 ? 1 : gx, gy), 128, 0, st>>>(modes, k, m, U, mean, F, out);
    BK_LAUNCH_CHECK("k_pod_inverse");
    return 0;
}

// ---------------------------------------------------------------- fused field + Saltelli sums
// ymodes: [nb][count][m] (block-major, as bk_sobol_modes writes them).  One warp = one 8-feature tile over the
// CTA's slab of groups; shared accumulators acc[warp][nsums][8 features][2].
__global__ void __launch_bounds__(128) k_pod_sobol(const double* __restrict__ ymodes, long count, int p, int m,
                                                   int second, int nb, int nsums, const double* __restrict__ U,
                                                   const double* __restrict__ mean, int F,
                                                   const double* __restrict__ shift, double* __restrict__ partials) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tq = lane & 3;
    double* acc = sm + (size_t)warp * nsums * 16;                         // [nsums][8][2]
    double* ys = sm + (size_t)4 * nsums * 16 + (size_t)warp * nb * 64;    // [nb][32 lanes][2 features]
    const int ftiles = (F + 7) / 8, ksteps = (m + 3) / 4;
    const long gtiles = (count + 7) / 8;
    const int slot = blockIdx.x;                                          // group slab = partial slot
    for (int ft = blockIdx.y * 4 + warp; ft < ftiles; ft += gridDim.y * 4) {
        for (int i = lane; i < nsums * 16; i += 32) acc[i] = 0.0;
        __syncwarp();
        const int fb = ft * 8 + g;
        double b[POD_MAX_KSTEPS];
#pragma unroll
        for (int ks = 0; ks < POD_MAX_KSTEPS; ++ks) {
            const int q = ks * 4 + tq;
            b[ks] = (ks < ksteps && fb < F && q < m) ? U[(size_t)fb * m + q] : 0.0;
        }
        const int f0 = ft * 8 + 2 * tq;
        const double m0 = f0 < F ? mean[f0] - shift[f0] : 0.0, m1 = f0 + 1 < F ? mean[f0 + 1] - shift[f0 + 1] : 0.0;
        for (long gt = blockIdx.x; gt < gtiles; gt += gridDim.x) {
            const long grp = gt * 8 + g;
            const bool valid = grp < count;
            for (int blk = 0; blk < nb; ++blk) {
                double c0 = m0, c1 = m1;
                const double* src = ymodes + ((size_t)blk * count + (valid ? grp : 0)) * m;
#pragma unroll
                for (int ks = 0; ks < POD_MAX_KSTEPS; ++ks) {
                    if (ks < ksteps) {
                        const int q = ks * 4 + tq;
                        const double a = (valid && q < m) ? src[q] : 0.0;
                        dmma884(c0, c1, a, b[ks]);
                    }
                }
                for (int ks = POD_MAX_KSTEPS; ks < ksteps; ++ks) {      // m > 64: same order, U from the cache
                    const int q = ks * 4 + tq;
                    const double a = (valid && q < m) ? src[q] : 0.0;
                    dmma884(c0, c1, a, (fb < F && q < m) ? U[(size_t)fb * m + q] : 0.0);
                }
                // shifted outputs of (group g, features f0, f0+1); rows beyond count contribute exact zeros
                ys[(blk * 32 + lane) * 2 + 0] = valid ? c0 : 0.0;
                ys[(blk * 32 + lane) * 2 + 1] = valid ? c1 : 0.0;
            }
            __syncwarp();
            // products for both features, reduced over the 8 groups of the fragment (lanes differing in g)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int r = 0;
                auto Y = [&](int blk) { return ys[(blk * 32 + lane) * 2 + e]; };
                auto emit8 = [&](double v) {
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if (g == (r & 7)) {                       // sum r of this feature is owned by lane (g = r%8, tq)
                        double* a = acc + ((size_t)r * 8 + 2 * tq + e) * 2;
                        dd t{a[0], a[1]};
                        dd_add_d(t, v);
                        a[0] = t.hi; a[1] = t.lo;
                    }
                    ++r;
                };
                const double yA = Y(0), yB = Y(1);
                emit8(yA); emit8(yB); emit8(__dmul_rn(yA, yA)); emit8(__dmul_rn(yB, yB)); emit8(__dmul_rn(yA, yB));
                for (int i = 0; i < p; ++i) {
                    const double yE = Y(2 + i);
                    emit8(yE); emit8(__dmul_rn(yE, yE)); emit8(__dmul_rn(yB, yE)); emit8(__dmul_rn(yA, yE));
                    const double dB = __dadd_rn(yB, -yE), dA = __dadd_rn(yA, -yE);
                    emit8(__dmul_rn(dB, dB)); emit8(__dmul_rn(dA, dA));
                }
                if (second && p > 2) {
                    for (int j = 0; j < p; ++j) { const double yC = Y(2 + p + j); emit8(yC); emit8(__dmul_rn(yC, yC)); }
                    for (int i = 0; i < p; ++i)
                        for (int j = 0; j < i; ++j) emit8(__dmul_rn(Y(2 + i), Y(2 + p + j)));
                }
                emit8(valid ? 1.0 : 0.0);                 // row count
            }
            __syncwarp();
        }
        // flush this feature tile into slot `slot` (ADD: chunks accumulate)
        __syncwarp();
        for (int i = lane; i < nsums * 8; i += 32) {
            const int r = i / 8, fl = i % 8, f = ft * 8 + fl;
            if (f < F) {
                double* dst = partials + (((size_t)slot * F + f) * nsums + r) * 2;
                const dd t = dd_add(dd{dst[0], dst[1]}, dd{acc[((size_t)r * 8 + fl) * 2], acc[((size_t)r * 8 + fl) * 2 + 1]});
                dst[0] = t.hi; dst[1] = t.lo;
            }
        }
        __syncwarp();
    }
}

int launch_pod_sobol(const double* ymodes, long count, int p, int m, int second, const double* U, const double* mean,
                     int F, const double* shift, double* partials, cudaStream_t st) {
    if (count <= 0) return 0;
    const int nsums = sobol_nsums_host(p, second);
    const int nb = (second && p > 2) ? 2 * p + 2 : p + 2;
    const size_t smem = ((size_t)4 * nsums * 16 + (size_t)4 * nb * 64) * 8;
    if (smem > 220 * 1024) return set_error("bk_pod_sobol_accumulate: p = %d needs %zu B of shared memory", p, smem);
    const int ftiles = (F + 7) / 8;
    int gy = (ftiles + 3) / 4;                       // every feature tile gets its own warp where possible
    if (gy > 64) gy = 64;
    int gx = sobol_nslots_host() / gy;               // group slabs = partial slots actually used
    if (gx < 1) gx = 1;
    const long gtiles = (count + 7) / 8;
    if (gx > gtiles) gx = (int)gtiles;
    BK_CUDA(cudaFuncSetAttribute(k_pod_sobol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope prof(PROF_POD, st);
    k_pod_sobol<<<dim3(gx, gy), 128, smem, st>>>(ymodes, count, p, m, second, nb, nsums, U, mean, F, shift, partials);
    BK_LAUNCH_CHECK("k_pod_sobol");
    return 0;
}

}  // namespace bk
