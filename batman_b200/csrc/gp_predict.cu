// K1 `gp_kstar_mean`: fused anisotropic Matern/RBF cross-covariance x alpha.
//
// Replaces  K_trans = kernel_(X, X_train_);  y_mean = K_trans @ alpha_   (sklearn:_gpr.py:446-450)
// for every output of Kriging.evaluate (kriging.py:209-212).  K* is produced in
// registers from TMA-staged training points and contracted with alpha on the
// fly; it only goes to HBM (already in the DMMA B-fragment order K2 wants) when
// sigma is requested.
//
// Work split: 128 threads per CTA, 2 test points per thread (coordinates in
// registers, P is a template parameter), training rows + alpha streamed through
// a 2-stage shared-memory ring with cp.async.bulk + mbarrier (SASS: UBLKCP).
// Every train row read from shared memory is a warp-wide broadcast LDS.128.
// Bound: FP64 pipe (3P + sqrt + exp + ~12 DP instructions per pair); bytes are
// negligible (8P/256 B of training data per pair from L2).
#include "common.cuh"
#include "covariance.cuh"
#include "kstar_writers.cuh"
#include "model.h"

namespace bk {

constexpr int PRED_THREADS = 128;
constexpr int PRED_TP = 2;                         // test points per thread
constexpr int PRED_TJ = 128;                       // training rows per stage

template <int P>
struct PredictParams {
    const double* pts;     // k x p raw points
    long k;
    int p;
    int n_pad;             // training rows, multiple of 128, zero padded
    const double* xs;      // n_pad x P
    const double* alpha;   // n_pad
    double scale[P], mn[P], ls[P];
    double c0, c1, y_mean, y_std;
    double* mean;          // k x ld, column q
    int ld, q;
    double* kstar;         // packed tiles [T][n_pad/16][2048] or null
    int8_t* kq;            // int8 slice tiles [T][n_pad/32][8][4096] (WRITE_K == 2)
    double inv_kappa;      // 1 / (power of two >= max |k|)
};

// WRITE_K: 0 = mean only, 1 = K* as FP64 DMMA fragment tiles (gp_var.cu), 2 = K* as 8 signed 7-bit int8 slices,
// 3 = K* as 7 unsigned byte slices (var_ozaki.cu schemes 0 / 1)
template <int KIND, int P, int WRITE_K>
__global__ void __launch_bounds__(PRED_THREADS) k_predict(const PredictParams<P> prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* xs_s = reinterpret_cast<double*>(smem_raw);                 // [2][TJ*P]
    double* al_s = xs_s + 2 * PRED_TJ * P;                              // [2][TJ]
    uint64_t* bar = reinterpret_cast<uint64_t*>(al_s + 2 * PRED_TJ);    // [2]

    const int tid = threadIdx.x;
    const int ntiles = prm.n_pad / PRED_TJ;
    constexpr uint32_t XS_BYTES = PRED_TJ * P * 8, AL_BYTES = PRED_TJ * 8;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < 2 && s < ntiles; ++s) {
            mbar_expect_tx(&bar[s], XS_BYTES + AL_BYTES);
            tma_load_1d(xs_s + s * PRED_TJ * P, prm.xs + (size_t)s * PRED_TJ * P, XS_BYTES, &bar[s]);
            tma_load_1d(al_s + s * PRED_TJ, prm.alpha + (size_t)s * PRED_TJ, AL_BYTES, &bar[s]);
        }
    }

    // this thread's test points: MinMax scaling then division by the length scale,
    // rounded as  X*scale_ + min_  (sklearn MinMaxScaler.transform) and  X / length_scale
    const long base = (long)blockIdx.x * (PRED_THREADS * PRED_TP);
    double u[PRED_TP][P];
    long t_idx[PRED_TP];
#pragma unroll
    for (int e = 0; e < PRED_TP; ++e) {
        long t = base + e * PRED_THREADS + tid;
        t_idx[e] = t;
        long tc = t < prm.k ? t : prm.k - 1;
#pragma unroll
        for (int c = 0; c < P; ++c) {
            double x = c < prm.p ? prm.pts[tc * prm.p + c] : 0.0;
            u[e][c] = __ddiv_rn(__dadd_rn(__dmul_rn(x, prm.scale[c]), prm.mn[c]), prm.ls[c]);
        }
    }

    double acc[PRED_TP], comp[PRED_TP];
#pragma unroll
    for (int e = 0; e < PRED_TP; ++e) { acc[e] = 0.0; comp[e] = 0.0; }

    // K* destination of this thread (see gp_var.cu / var_ozaki.cu for the tile layouts)
    double* kdst[PRED_TP];
    int8_t* qdst[PRED_TP];
#pragma unroll
    for (int e = 0; e < PRED_TP; ++e) {
        kdst[e] = WRITE_K == 1 ? kstar_tile_base(prm.kstar, t_idx[e], prm.n_pad) : nullptr;
        qdst[e] = WRITE_K >= 2 ? kq_tile_base(prm.kq, t_idx[e], prm.n_pad) : nullptr;
    }

    for (int tile = 0; tile < ntiles; ++tile) {
        const int s = tile & 1;
        mbar_wait(&bar[s], (tile >> 1) & 1);
        const double* xrow = xs_s + s * PRED_TJ * P;
        const double* arow = al_s + s * PRED_TJ;
#pragma unroll 1
        for (int j8 = 0; j8 < PRED_TJ; j8 += 8) {
            double kbuf[PRED_TP][8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = j8 + jj;
                double v[P];
#pragma unroll
                for (int c = 0; c < P; c += 2) {
                    double2 vv = *reinterpret_cast<const double2*>(xrow + j * P + c);
                    v[c] = vv.x;
                    v[c + 1] = vv.y;
                }
                const double a = arow[j];
#pragma unroll
                for (int e = 0; e < PRED_TP; ++e) {
                    double ssum = 0.0;
#pragma unroll
                    for (int c = 0; c < P; ++c) {
                        double d = __dadd_rn(u[e][c], -v[c]);
                        ssum = __dadd_rn(ssum, __dmul_rn(d, d));
                    }
                    double kv = affine(prm.c0, prm.c1, stationary<KIND>(ssum));
                    kbuf[e][jj] = kv;
                    // Kahan: the sum is sign-alternating and badly conditioned
                    double y = __fma_rn(kv, a, -comp[e]);       // product unrounded: the summation order is ours, not sklearn's
                    double tsum = __dadd_rn(acc[e], y);
                    comp[e] = __dadd_rn(__dadd_rn(tsum, -acc[e]), -y);
                    acc[e] = tsum;
                }
            }
            if constexpr (WRITE_K != 0) {
                const int jg = tile * PRED_TJ + j8;          // global train index of kbuf[.][0]
#pragma unroll
                for (int e = 0; e < PRED_TP; ++e) kstar_write8<WRITE_K>(kbuf[e], jg, kdst[e], qdst[e], prm.inv_kappa);
            }
        }
        __syncthreads();   // everyone is done reading stage s
        if (tid == 0 && tile + 2 < ntiles) {
            mbar_expect_tx(&bar[s], XS_BYTES + AL_BYTES);
            tma_load_1d(xs_s + s * PRED_TJ * P, prm.xs + (size_t)(tile + 2) * PRED_TJ * P, XS_BYTES, &bar[s]);
            tma_load_1d(al_s + s * PRED_TJ, prm.alpha + (size_t)(tile + 2) * PRED_TJ, AL_BYTES, &bar[s]);
        }
    }

#pragma unroll
    for (int e = 0; e < PRED_TP; ++e) {
        if (t_idx[e] < prm.k) {
            // y_mean = _y_train_std * y_mean + _y_train_mean   (sklearn:_gpr.py:450)
            prm.mean[t_idx[e] * prm.ld + prm.q] = __dadd_rn(__dmul_rn(prm.y_std, acc[e]), prm.y_mean);
        }
    }
}

template <int KIND, int P>
static int launch_predict_kp(const bk_gp_s* h, int q, const double* pts, long k, double* mean, double* kstar,
                             int8_t* kq, int kq_scheme, double kappa, cudaStream_t st) {
    const GpOutput& o = h->out[q];
    PredictParams<P> prm;
    prm.pts = pts; prm.k = k; prm.p = h->p; prm.n_pad = h->n_pad;
    prm.xs = o.xs; prm.alpha = o.alpha;
    for (int c = 0; c < P; ++c) {
        prm.scale[c] = c < h->p ? h->scale[c] : 0.0;
        prm.mn[c] = c < h->p ? h->mn[c] : 0.0;
        prm.ls[c] = c < h->p ? o.ls[c] : 1.0;
    }
    prm.c0 = o.c0; prm.c1 = o.c1; prm.y_mean = o.y_mean; prm.y_std = o.y_std;
    prm.mean = mean; prm.ld = h->m; prm.q = q; prm.kstar = kstar; prm.kq = kq; prm.inv_kappa = 1.0 / kappa;
    const size_t smem = (size_t)2 * PRED_TJ * P * 8 + 2 * PRED_TJ * 8 + 16;
    const long per_cta = PRED_THREADS * PRED_TP;
    // with K* output every 128-point tile must be fully written: round the grid up to whole tiles
    const unsigned grid = (unsigned)((k + per_cta - 1) / per_cta);
    ProfScope prof(PROF_PREDICT, st);
    if (kq && kq_scheme == 1) {
        BK_CUDA(cudaFuncSetAttribute(k_predict<KIND, P, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_predict<KIND, P, 3><<<grid, PRED_THREADS, smem, st>>>(prm);
    } else if (kq) {
        BK_CUDA(cudaFuncSetAttribute(k_predict<KIND, P, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_predict<KIND, P, 2><<<grid, PRED_THREADS, smem, st>>>(prm);
    } else if (kstar) {
        BK_CUDA(cudaFuncSetAttribute(k_predict<KIND, P, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_predict<KIND, P, 1><<<grid, PRED_THREADS, smem, st>>>(prm);
    } else {
        BK_CUDA(cudaFuncSetAttribute(k_predict<KIND, P, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_predict<KIND, P, 0><<<grid, PRED_THREADS, smem, st>>>(prm);
    }
    BK_LAUNCH_CHECK("k_predict");
    return 0;
}

template <int KIND>
static int launch_predict_k(const bk_gp_s* h, int q, const double* pts, long k, double* mean, double* kstar,
                            int8_t* kq, int kq_scheme, double kappa, cudaStream_t st) {
    switch (h->P) {
        case 2: return launch_predict_kp<KIND, 2>(h, q, pts, k, mean, kstar, kq, kq_scheme, kappa, st);
        case 4: return launch_predict_kp<KIND, 4>(h, q, pts, k, mean, kstar, kq, kq_scheme, kappa, st);
        case 6: return launch_predict_kp<KIND, 6>(h, q, pts, k, mean, kstar, kq, kq_scheme, kappa, st);
        case 8: return launch_predict_kp<KIND, 8>(h, q, pts, k, mean, kstar, kq, kq_scheme, kappa, st);
        case 10: return launch_predict_kp<KIND, 10>(h, q, pts, k, mean, kstar, kq, kq_scheme, kappa, st);
        case 12: return launch_predict_kp<KIND, 12>(h, q, pts, k, mean, kstar, kq, kq_scheme, kappa, st);
        case 16: return launch_predict_kp<KIND, 16>(h, q, pts, k, mean, kstar, kq, kq_scheme, kappa, st);
        case 20: return launch_predict_kp<KIND, 20>(h, q, pts, k, mean, kstar, kq, kq_scheme, kappa, st);
        case 24: return launch_predict_kp<KIND, 24>(h, q, pts, k, mean, kstar, kq, kq_scheme, kappa, st);
        case 32: return launch_predict_kp<KIND, 32>(h, q, pts, k, mean, kstar, kq, kq_scheme, kappa, st);
    }
    return set_error("unsupported padded dimension %d", h->P);
}

__global__ void k_debug_fastmath(const double* __restrict__ in, long n, double* __restrict__ o_sqrt,
                                 double* __restrict__ o_exp) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = in[i];
    o_sqrt[i] = bk_sqrt(v);
    o_exp[i] = bk_exp_neg(-v);
}
int launch_debug_fastmath(const double* in, long n, double* o_sqrt, double* o_exp, cudaStream_t st) {
    if (n <= 0) return 0;
    k_debug_fastmath<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, n, o_sqrt, o_exp);
    BK_LAUNCH_CHECK("k_debug_fastmath");
    return 0;
}

int launch_predict_general(const bk_gp_s* h, int q, const double* pts, long k, double* mean, int ld, int col,
                           double* kstar, int8_t* kq, int kq_scheme, double kappa, cudaStream_t st);   // gp_predict_gen.cu

// mean of output q at k points; if kstar != null also the packed K* tiles for gp_var
int launch_predict(const bk_gp_s* h, int q, const double* pts, long k, double* mean, double* kstar, int8_t* kq,
                   int kq_scheme, double kappa, cudaStream_t st) {
    switch (h->out[q].kind) {
        case BK_MATERN12: return launch_predict_k<BK_MATERN12>(h, q, pts, k, mean, kstar, kq, kq_scheme, kappa, st);
        case BK_MATERN32: return launch_predict_k<BK_MATERN32>(h, q, pts, k, mean, kstar, kq, kq_scheme, kappa, st);
        case BK_MATERN52: return launch_predict_k<BK_MATERN52>(h, q, pts, k, mean, kstar, kq, kq_scheme, kappa, st);
        case BK_RBF: return launch_predict_k<BK_RBF>(h, q, pts, k, mean, kstar, kq, kq_scheme, kappa, st);
        case BK_MATERN_INF: return launch_predict_k<BK_MATERN_INF>(h, q, pts, k, mean, kstar, kq, kq_scheme, kappa, st);
        case BK_GENERAL: return launch_predict_general(h, q, pts, k, mean, h->m, q, kstar, kq, kq_scheme, kappa, st);
    }
    return set_error("output %d has no kernel set", q);
}

}  // namespace bk
