// K1 for general kernel trees (kind == BK_GENERAL): cross-covariance x alpha with several stationary leaves.
//
// Same contract and K* output formats as k_predict (gp_predict.cu); the covariance is the expanded Sum/Product tree
//     k(x, x') = sum_t coef_t * prod_f stat_f(||(x - x')/ls_f||)^e_tf             (covariance.cuh: general_cov)
// of sklearn:kernels.py:838-871, 936-990 -- e.g. `C * RBF * Matern`, `C1 * RBF + C2 * Matern + White`, which the
// reference accepts through `Kriging(kernel=...)` (kriging.py:80-96, driver.py:181-190).  Every leaf has its own
// length scales, so a training row is staged as nf scaled copies [X_j/ls_0 | X_j/ls_1 | ...] and the test points'
// nf scaled copies live in shared memory ([f][c][e][tid]: conflict-free 8-byte columns) instead of registers.
// This path trades speed for coverage (2 LDS per difference instead of 1 broadcast LDS); single-leaf affine trees
// -- every kernel the reference's tests and docs use -- never come here.
#include "common.cuh"
#include "covariance.cuh"
#include "kstar_writers.cuh"
#include "model.h"

namespace bk {

constexpr int PG_THREADS = 128;
constexpr int PG_TP = 2;
constexpr int PG_TJ = 128;

template <int P>
struct PredictGenParams {
    const double* pts;
    long k;
    int p, n_pad;
    const double* xs;      // n_pad x (nf * P)
    const double* alpha;
    double scale[P], mn[P];
    double ls[GEN_MAX_F][P];
    GenSpec gen;
    double y_mean, y_std;
    double* mean;
    int ld, q;
    double* kstar;
    int8_t* kq;
    double inv_kappa;
};

template <int P, int WRITE_K>
__global__ void __launch_bounds__(PG_THREADS) k_predict_gen(const PredictGenParams<P> prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int nf = prm.gen.nf, FP = nf * P;
    double* xs_s = reinterpret_cast<double*>(smem_raw);                 // [2][TJ * FP]
    double* al_s = xs_s + 2 * PG_TJ * FP;                               // [2][TJ]
    double* u_s = al_s + 2 * PG_TJ;                                     // [FP][TP][128]
    uint64_t* bar = reinterpret_cast<uint64_t*>(u_s + (size_t)FP * PG_TP * PG_THREADS);

    const int tid = threadIdx.x;
    const int ntiles = prm.n_pad / PG_TJ;
    const uint32_t XS_BYTES = PG_TJ * FP * 8, AL_BYTES = PG_TJ * 8;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < 2 && s < ntiles; ++s) {
            mbar_expect_tx(&bar[s], XS_BYTES + AL_BYTES);
            tma_load_1d(xs_s + s * PG_TJ * FP, prm.xs + (size_t)s * PG_TJ * FP, XS_BYTES, &bar[s]);
            tma_load_1d(al_s + s * PG_TJ, prm.alpha + (size_t)s * PG_TJ, AL_BYTES, &bar[s]);
        }
    }

    const long base = (long)blockIdx.x * (PG_THREADS * PG_TP);
    long t_idx[PG_TP];
#pragma unroll
    for (int e = 0; e < PG_TP; ++e) {
        const long t = base + e * PG_THREADS + tid;
        t_idx[e] = t;
        const long tc = t < prm.k ? t : prm.k - 1;
#pragma unroll
        for (int c = 0; c < P; ++c) {
            const double x = c < prm.p ? prm.pts[tc * prm.p + c] : 0.0;
            const double xs = __dadd_rn(__dmul_rn(x, prm.scale[c]), prm.mn[c]);      // MinMaxScaler.transform
            for (int f = 0; f < nf; ++f)
                u_s[((size_t)(f * P + c) * PG_TP + e) * PG_THREADS + tid] = __ddiv_rn(xs, prm.ls[f][c]);   // X / length_scale
        }
    }
    // u_s entries are private to their thread: no barrier needed before use

    double acc[PG_TP], comp[PG_TP];
    double* kdst[PG_TP];
    int8_t* qdst[PG_TP];
#pragma unroll
    for (int e = 0; e < PG_TP; ++e) {
        acc[e] = 0.0; comp[e] = 0.0;
        kdst[e] = WRITE_K == 1 ? kstar_tile_base(prm.kstar, t_idx[e], prm.n_pad) : nullptr;
        qdst[e] = WRITE_K >= 2 ? kq_tile_base(prm.kq, t_idx[e], prm.n_pad) : nullptr;
    }

    for (int tile = 0; tile < ntiles; ++tile) {
        const int s = tile & 1;
        mbar_wait(&bar[s], (tile >> 1) & 1);
        const double* xrow = xs_s + s * PG_TJ * FP;
        const double* arow = al_s + s * PG_TJ;
#pragma unroll 1
        for (int j8 = 0; j8 < PG_TJ; j8 += 8) {
            double kbuf[PG_TP][8];
#pragma unroll 1
            for (int jj = 0; jj < 8; ++jj) {
                const int j = j8 + jj;
                double fv[PG_TP][GEN_MAX_F];
#pragma unroll
                for (int f = 0; f < GEN_MAX_F; ++f) {
                    if (f < nf) {
                        double ssum[PG_TP];
#pragma unroll
                        for (int e = 0; e < PG_TP; ++e) ssum[e] = 0.0;
                        const double* vr = xrow + j * FP + f * P;
                        const double* ur = u_s + (size_t)(f * P) * PG_TP * PG_THREADS + tid;
#pragma unroll
                        for (int c = 0; c < P; ++c) {
                            const double v = vr[c];
#pragma unroll
                            for (int e = 0; e < PG_TP; ++e) {
                                const double d = __dadd_rn(ur[(c * PG_TP + e) * PG_THREADS], -v);
                                ssum[e] = __dadd_rn(ssum[e], __dmul_rn(d, d));       // cdist order: sequential, non-fused
                            }
                        }
#pragma unroll
                        for (int e = 0; e < PG_TP; ++e) fv[e][f] = stationary_rt(prm.gen.fkind[f], ssum[e]);
                    } else {
#pragma unroll
                        for (int e = 0; e < PG_TP; ++e) fv[e][f] = 1.0;
                    }
                }
                const double a = arow[j];
#pragma unroll
                for (int e = 0; e < PG_TP; ++e) {
                    const double kv = general_cov(prm.gen, fv[e]);
                    // kbuf is indexed with a loop-carried jj: keep it in registers with a uniform select chain
#pragma unroll
                    for (int w = 0; w < 8; ++w)
                        if (w == jj) kbuf[e][w] = kv;
                    const double y = __fma_rn(kv, a, -comp[e]);                      // Kahan, as in k_predict
                    const double tsum = __dadd_rn(acc[e], y);
                    comp[e] = __dadd_rn(__dadd_rn(tsum, -acc[e]), -y);
                    acc[e] = tsum;
                }
            }
            if constexpr (WRITE_K != 0) {
                const int jg = tile * PG_TJ + j8;
#pragma unroll
                for (int e = 0; e < PG_TP; ++e) kstar_write8<WRITE_K>(kbuf[e], jg, kdst[e], qdst[e], prm.inv_kappa);
            }
        }
        __syncthreads();
        if (tid == 0 && tile + 2 < ntiles) {
            mbar_expect_tx(&bar[s], XS_BYTES + AL_BYTES);
            tma_load_1d(xs_s + s * PG_TJ * FP, prm.xs + (size_t)(tile + 2) * PG_TJ * FP, XS_BYTES, &bar[s]);
            tma_load_1d(al_s + s * PG_TJ, prm.alpha + (size_t)(tile + 2) * PG_TJ, AL_BYTES, &bar[s]);
        }
    }

#pragma unroll
    for (int e = 0; e < PG_TP; ++e)
        if (t_idx[e] < prm.k)
            prm.mean[t_idx[e] * prm.ld + prm.q] = __dadd_rn(__dmul_rn(prm.y_std, acc[e]), prm.y_mean);
}

template <int P, int WRITE_K>
static int launch_gen_pw(const PredictGenParams<P>& prm, size_t smem, unsigned grid, cudaStream_t st) {
    BK_CUDA(cudaFuncSetAttribute(k_predict_gen<P, WRITE_K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_predict_gen<P, WRITE_K><<<grid, PG_THREADS, smem, st>>>(prm);
    BK_LAUNCH_CHECK("k_predict_gen");
    return 0;
}

template <int P>
static int launch_gen_p(const bk_gp_s* h, int q, const double* pts, long k, double* mean, int ld, int col,
                        double* kstar, int8_t* kq, int kq_scheme, double kappa, cudaStream_t st) {
    const GpOutput& o = h->out[q];
    if (o.gen.nf * P > GEN_MAX_FP)
        return set_error("general kernel tree: %d stationary leaves x padded dimension %d exceeds %d", o.gen.nf, P, GEN_MAX_FP);
    PredictGenParams<P> prm;
    prm.pts = pts; prm.k = k; prm.p = h->p; prm.n_pad = h->n_pad;
    prm.xs = o.xs; prm.alpha = o.alpha;
    for (int c = 0; c < P; ++c) {
        prm.scale[c] = c < h->p ? h->scale[c] : 0.0;
        prm.mn[c] = c < h->p ? h->mn[c] : 0.0;
        for (int f = 0; f < GEN_MAX_F; ++f) prm.ls[f][c] = (c < h->p && f < o.gen.nf) ? o.gls[f][c] : 1.0;
    }
    prm.gen = o.gen; prm.y_mean = o.y_mean; prm.y_std = o.y_std;
    prm.mean = mean; prm.ld = ld; prm.q = col; prm.kstar = kstar; prm.kq = kq; prm.inv_kappa = 1.0 / kappa;
    const int FP = o.gen.nf * P;
    const size_t smem = ((size_t)2 * PG_TJ * FP + 2 * PG_TJ + (size_t)FP * PG_TP * PG_THREADS) * 8 + 16;
    const long per_cta = PG_THREADS * PG_TP;
    const unsigned grid = (unsigned)((k + per_cta - 1) / per_cta);
    ProfScope prof(PROF_PREDICT, st);
    if (kq && kq_scheme == 1) return launch_gen_pw<P, 3>(prm, smem, grid, st);
    if (kq) return launch_gen_pw<P, 2>(prm, smem, grid, st);
    if (kstar) return launch_gen_pw<P, 1>(prm, smem, grid, st);
    return launch_gen_pw<P, 0>(prm, smem, grid, st);
}

int launch_predict_general(const bk_gp_s* h, int q, const double* pts, long k, double* mean, int ld, int col,
                           double* kstar, int8_t* kq, int kq_scheme, double kappa, cudaStream_t st) {
#define BK_CASE(PP) case PP: return launch_gen_p<PP>(h, q, pts, k, mean, ld, col, kstar, kq, kq_scheme, kappa, st);
    switch (h->P) {
        BK_CASE(2) BK_CASE(4) BK_CASE(6) BK_CASE(8) BK_CASE(10) BK_CASE(12) BK_CASE(16) BK_CASE(20) BK_CASE(24)
    }
#undef BK_CASE
    return set_error("general kernel tree: padded dimension %d is not instantiated (max 24)", h->P);
}

}  // namespace bk
