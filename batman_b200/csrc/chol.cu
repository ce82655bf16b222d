// K5: batched log-marginal-likelihood (build K, blocked Cholesky, forward solve) and the
// triangular inverse W = L^-1 that K2 consumes -- all FP64, DMMA for the O(n^3) parts.
//
// Replaces, per hyper-parameter candidate theta_r (kriging.py:164-171 -> sklearn:_gpr.py:584-617):
//     K = kernel(X_train) ; K[diag] += 1e-10 ; L = cholesky(K) ; alpha = cho_solve(L, y)
//     lml = -1/2 y.alpha - sum(log diag L) - n/2 log(2 pi)            (-inf if the Cholesky fails)
// The reference evaluates one theta per call in forked processes; here R candidates are factorised
// side by side (grid.y = candidate).  y.alpha = ||L^-1 y||^2, so only the forward solve is needed for
// the LML; z = L^-1 y rides along with the factorisation, block column by block column.
//
// Blocked left-looking Cholesky, NB = 128, row-major n_pad x n_pad per candidate (lower triangle):
//   for block column k:   T_ik  = A_ik - sum_{j<k} L_ij L_kj^T      (k_gemm_nt: DMMA, i >= k)
//                         L_kk  = chol(T_kk), Winv_k = L_kk^-1       (k_chol_diag: one CTA, shared memory)
//                         z_k   = Winv_k (y_k - L_k,<k z_<k)         (same kernel)
//                         L_ik  = T_ik Winv_k^T                      (k_gemm_nt, i > k)
// The same NT GEMM drives the blocked triangular inversion (computed transposed, WT = L^-T):
//   for block column i:   S_ki  = sum_{k<=j<i} WT_kj L_ij^T ;  WT_ki = -S_ki Winv_i^T ;  WT_ii = Winv_i^T
// Padded rows/columns carry an identity block, so log-det, z and W are unaffected.
#include <math.h>

#include <vector>

#include "common.cuh"
#include "covariance.cuh"
#include "model.h"

namespace bk {

constexpr int NB = 128;
constexpr int SPLIT_MAX = 8;       // split-K slices of the diagonal block-row update
constexpr double JITTER = 1e-10;   // GaussianProcessRegressor(alpha=1e-10), sklearn:_gpr.py:215

// ---------------------------------------------------------------- candidate parameters (device copy)
struct Cand {
    int kind;
    int target;                       // row of the (n_targets x n) target matrix this candidate is scored on
    double c0, c1, kdiag;
    double ls[MAX_P];
    GenSpec gen;                      // kind == BK_GENERAL
    double gls[GEN_MAX_F][MAX_P];
};

// ---------------------------------------------------------------- K(theta_r): lower block tiles
__global__ void __launch_bounds__(256) k_build_K(const double* __restrict__ X, int n, int p, int n_pad,
                                                 const Cand* __restrict__ cands, double* __restrict__ M,
                                                 long batch_stride) {
    extern __shared__ __align__(16) double sm[];
    // lower-triangular tile index -> (bi, bj), bj <= bi
    int t = blockIdx.x, bi = 0;
    while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
    const int bj = t - bi * (bi + 1) / 2;
    const Cand& cd = cands[blockIdx.y];
    double* Mr = M + (size_t)blockIdx.y * batch_stride;
    const int kind = cd.kind;
    const int nf = kind == BK_GENERAL ? cd.gen.nf : 1;
    double* xr = sm;                      // [nf][128][p]   rows of this tile, scaled per leaf
    double* xcT = sm + nf * NB * p;       // [nf][p][128]   cols of this tile, scaled, transposed
    for (int idx = threadIdx.x; idx < nf * NB * p; idx += blockDim.x) {
        const int f = idx / (NB * p), r = (idx / p) % NB, c = idx % p;
        const int gr = bi * NB + r, gc = bj * NB + r;
        const double l = kind == BK_GENERAL ? cd.gls[f][c] : cd.ls[c];
        xr[(f * NB + r) * p + c] = gr < n ? __ddiv_rn(X[(size_t)gr * p + c], l) : 0.0;     // X / length_scale
        xcT[(f * p + c) * NB + r] = gc < n ? __ddiv_rn(X[(size_t)gc * p + c], l) : 0.0;
    }
    __syncthreads();
    for (int e = 0; e < NB * NB / 256; ++e) {
        const int idx = e * 256 + threadIdx.x;
        const int r = idx >> 7, c = idx & 127;
        const int gr = bi * NB + r, gc = bj * NB + c;
        double v;
        if (gr >= n || gc >= n) {
            v = gr == gc ? 1.0 : 0.0;                                   // identity padding
        } else if (gr == gc) {
            v = __dadd_rn(cd.kdiag, JITTER);                            // kernel_.diag + alpha (sklearn:_gpr.py:589)
        } else {
            double fv[GEN_MAX_F] = {1.0, 1.0, 1.0};
            for (int f = 0; f < nf; ++f) {
                double s = 0.0;
                for (int d = 0; d < p; ++d) {                           // pdist order: sequential, non-fused
                    const double df = __dadd_rn(xr[(f * NB + r) * p + d], -xcT[(f * p + d) * NB + c]);
                    s = __dadd_rn(s, __dmul_rn(df, df));
                }
                const double val = stationary_rt(kind == BK_GENERAL ? cd.gen.fkind[f] : kind, s);
                if (f == 0) fv[0] = val; else if (f == 1) fv[1] = val; else fv[2] = val;
            }
            v = kind == BK_GENERAL ? general_cov(cd.gen, fv) : affine(cd.c0, cd.c1, fv[0]);
        }
        Mr[(size_t)gr * n_pad + gc] = v;
    }
}

// ---------------------------------------------------------------- C = beta*C + alpha * A * B^T  (row-major, DMMA)
struct GemmArgs {
    const double* A; long a_batch, a_tile; int lda;
    const double* B; long b_batch, b_tile; int ldb;
    double* C; long c_batch, c_tile; int ldc;
    int K0, k_tile;          // K = K0 + blockIdx.x * k_tile   (multiple of 16)
    double alpha, beta;
    int split_k_total = 0;   // > 0: split-K, tile x covers columns [x*K0, min((x+1)*K0, split_k_total)) (a_tile = b_tile = K0)
};
constexpr int G_BK = 16, G_PITCH = 20, G_STAGES = 3;     // pitch 20 doubles: conflict-free fragment loads
constexpr size_t G_SMEM = (size_t)G_STAGES * 2 * NB * G_PITCH * 8;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(256, 1) k_gemm_nt(const GemmArgs g) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3, gq = lane >> 2, tq = lane & 3;
    const double* A = g.A + (size_t)blockIdx.y * g.a_batch + (size_t)blockIdx.x * g.a_tile;
    const double* B = g.B + (size_t)blockIdx.y * g.b_batch + (size_t)blockIdx.x * g.b_tile;
    double* C = g.C + (size_t)blockIdx.y * g.c_batch + (size_t)blockIdx.x * g.c_tile;
    int K = g.K0 + (int)blockIdx.x * g.k_tile;
    if (g.split_k_total > 0 && (int)(blockIdx.x + 1) * g.K0 > g.split_k_total) K = g.split_k_total - (int)blockIdx.x * g.K0;
    const int nk = K / G_BK;

    auto load_stage = [&](int s, int kb) {
        double* As = sm + (size_t)s * 2 * NB * G_PITCH;
        double* Bs = As + NB * G_PITCH;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int id = tid + 256 * i, row = id >> 3, ch = id & 7;
            cp_async16(As + row * G_PITCH + ch * 2, A + (size_t)row * g.lda + kb * G_BK + ch * 2);
            cp_async16(Bs + row * G_PITCH + ch * 2, B + (size_t)row * g.ldb + kb * G_BK + ch * 2);
        }
    };
    for (int s = 0; s < G_STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    double c[8][4][2];
#pragma unroll
    for (int mt = 0; mt < 8; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) c[mt][nt][0] = c[mt][nt][1] = 0.0;

    for (int kb = 0; kb < nk; ++kb) {
        cp_async_wait<G_STAGES - 2>();
        __syncthreads();                       // stage kb landed for everyone; stage kb-1 fully consumed
        if (kb + G_STAGES - 1 < nk) load_stage((kb + G_STAGES - 1) % G_STAGES, kb + G_STAGES - 1);
        cp_async_commit();
        const double* As = sm + (size_t)(kb % G_STAGES) * 2 * NB * G_PITCH;
        const double* Bs = As + NB * G_PITCH;
#pragma unroll
        for (int k4 = 0; k4 < G_BK / 4; ++k4) {
            double a[8], b[4];
#pragma unroll
            for (int mt = 0; mt < 8; ++mt) a[mt] = As[(wm * 64 + mt * 8 + gq) * G_PITCH + k4 * 4 + tq];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) b[nt] = Bs[(wn * 32 + nt * 8 + gq) * G_PITCH + k4 * 4 + tq];
#pragma unroll
            for (int mt = 0; mt < 8; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma884(c[mt][nt][0], c[mt][nt][1], a[mt], b[nt]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();      // in-place use (C aliases A): every read of A is complete before any write
#pragma unroll
    for (int mt = 0; mt < 8; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int r = wm * 64 + mt * 8 + gq, cc = wn * 32 + nt * 8 + 2 * tq;
            double2* dst = reinterpret_cast<double2*>(C + (size_t)r * g.ldc + cc);
            double2 v = make_double2(g.alpha * c[mt][nt][0], g.alpha * c[mt][nt][1]);
            if (g.beta != 0.0) {
                const double2 o = *dst;
                v.x = fma(g.beta, o.x, v.x);
                v.y = fma(g.beta, o.y, v.y);
            }
            *dst = v;
        }
}

static int gemm_nt(const GemmArgs& g, int tiles, int batch, cudaStream_t st) {
    if (tiles <= 0 || batch <= 0) return 0;
    static OncePerDevice attr;
    if (attr.first()) {
        BK_CUDA(cudaFuncSetAttribute(k_gemm_nt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM));
    }
    ProfScope prof(PROF_CHOL, st);
    k_gemm_nt<<<dim3(tiles, batch), 256, G_SMEM, st>>>(g);
    BK_LAUNCH_CHECK("k_gemm_nt");
    return 0;
}

// ---------------------------------------------------------------- s = y_k - L[k, <k] z_<k, one warp per row
// (16 CTAs per candidate instead of a serial prologue inside the single-CTA diagonal-block kernel)
__global__ void __launch_bounds__(256) k_zdot(const double* __restrict__ M, long m_batch, int ld, int k,
                                              const Cand* __restrict__ cands, const double* __restrict__ ypad,
                                              const double* __restrict__ z, long z_batch, double* __restrict__ sv) {
    const int r = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.y * 8 + warp;
    const double* row = M + (size_t)r * m_batch + (size_t)(k * NB + i) * ld;
    const double* zr = z + (size_t)r * z_batch;
    const double* y = ypad + (size_t)cands[r].target * ld;      // this candidate's (normalised, zero padded) targets
    double part = 0.0;
    for (int j = lane; j < k * NB; j += 32) part = fma(row[j], zr[j], part);     // lanes stride the columns (coalesced)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) sv[(size_t)r * NB + i] = y[k * NB + i] - part;
}

// ---------------------------------------------------------------- diagonal block: factor, invert, forward-solve
// One CTA per candidate works on the 128 x 128 diagonal block in shared memory, blocked by 32 so that the serial
// part (a 32 x 32 factorisation / inversion in the registers of ONE warp, shuffles instead of block barriers) is short
// and everything else is a small matrix product over all 256 threads.  The Cholesky performs, element by element,
// the same operations in the same order as the unblocked right-looking algorithm (scale by the reciprocal pivot,
// rank-1 updates in column order), so its results do not depend on the blocking.
constexpr int DP = NB + 1;   // shared pitch
constexpr int DB = 32;       // inner block
constexpr size_t DIAG_SMEM = ((size_t)NB * DP + 3 * NB + DB * (DB + 1)) * 8;

// Cholesky of the 32 x 32 diagonal block at (j0, j0) by one warp: lane = row, the row lives in registers.
// Writes L (zeros above the diagonal) back, the reciprocal pivots to rinv[j0 ..], returns the first non-positive
// pivot (1-based, local) or 0.
__device__ __forceinline__ int chol32_warp(double* a, int j0, int lane, double* rinv) {
    double r[DB];
#pragma unroll
    for (int c = 0; c < DB; ++c) r[c] = a[(j0 + lane) * DP + j0 + c];       // the strict upper triangle of a is zero
    int bad = 0;
#pragma unroll
    for (int j = 0; j < DB; ++j) {
        const double d = __shfl_sync(0xffffffffu, r[j], j);
        if (!(d > 0.0) && bad == 0) bad = j + 1;
        const double piv = sqrt(d), inv = 1.0 / piv;
        const double lj = lane > j ? r[j] * inv : (lane == j ? piv : r[j]);
        r[j] = lj;
        if (lane == j) rinv[j0 + j] = inv;
#pragma unroll
        for (int c = j + 1; c < DB; ++c) {
            const double lc = __shfl_sync(0xffffffffu, lj, c);
            r[c] = lane >= c ? fma(-lj, lc, r[c]) : r[c];
        }
    }
#pragma unroll
    for (int c = 0; c < DB; ++c) a[(j0 + lane) * DP + j0 + c] = c <= lane ? r[c] : 0.0;
    return bad;
}

// Inverse of the lower-triangular 32 x 32 block at (j0, j0) by one warp: lane = column k, forward substitution of
// L x = e_k with x in registers; L entries are shared-memory broadcasts.  out[i][k], pitch DB + 1.
__device__ __forceinline__ void inv32_warp(const double* a, int j0, int lane, double* out) {
    double x[DB];
#pragma unroll
    for (int i = 0; i < DB; ++i) {
        const double rd = 1.0 / a[(j0 + i) * DP + j0 + i];
        double sum = 0.0;
#pragma unroll
        for (int c = 0; c < i; ++c) sum = fma(a[(j0 + i) * DP + j0 + c], x[c], sum);     // x[c] = 0 for c < k
        x[i] = i < lane ? 0.0 : (i == lane ? rd : -sum * rd);
        out[i * (DB + 1) + lane] = x[i];
    }
}

// blocked Cholesky of a[NB][DP] (lower), 4 panels of 32 columns; returns through bad_s the first bad pivot (1-based)
__device__ void chol_block_smem(double* a, double* rinv, double* scratch, int* bad_s) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j0 = 0; j0 < NB; j0 += DB) {
        const int m = NB - DB - j0;                  // rows below the diagonal block
        __syncthreads();
        if (warp == 0) {
            const int b = chol32_warp(a, j0, lane, rinv);
            if (lane == 0 && b != 0 && *bad_s == 0) *bad_s = j0 + b;
        }
        __syncthreads();
        if (tid < m) {
            // panel rows: x = a[i, panel] L11^-T by substitution, the order of the unblocked algorithm
            const int i = j0 + DB + tid;
            double x[DB];
#pragma unroll
            for (int c = 0; c < DB; ++c) x[c] = a[i * DP + j0 + c];
#pragma unroll
            for (int t = 0; t < DB; ++t) {
                x[t] *= rinv[j0 + t];
#pragma unroll
                for (int c = t + 1; c < DB; ++c) x[c] = fma(-x[t], a[(j0 + c) * DP + j0 + t], x[c]);
            }
#pragma unroll
            for (int c = 0; c < DB; ++c) a[i * DP + j0 + c] = x[c];
        }
        __syncthreads();
        // trailing update (lower triangle): a[i][c] -= sum_t x[i][t] x[c][t], t ascending.  lane = column inside a
        // 32-wide column block, a warp walks rows: x[i][.] is a broadcast, x[c][.] is conflict free (pitch 129).
        const int r0 = j0 + DB;
        for (int cb = 0; cb < m; cb += DB) {
            const int c = r0 + cb + lane;
            for (int i = r0 + cb + warp; i < NB; i += 8) {
                if (c <= i) {
                    double acc = a[i * DP + c];
#pragma unroll 8
                    for (int t = 0; t < DB; ++t) acc = fma(-a[i * DP + j0 + t], a[c * DP + j0 + t], acc);
                    a[i * DP + c] = acc;
                }
            }
        }
    }
    __syncthreads();
    (void)scratch;
}

// in-place inverse of the lower-triangular matrix in a[NB][DP], block columns from the right (LAPACK dtrtri):
//   W21 = -W22 L21 L11^-1,  W11 = L11^-1   with W22 the already inverted trailing block
__device__ void tri_invert_smem(double* a, double* dsc /* [DB][DB+1] */) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j0 = NB - DB; j0 >= 0; j0 -= DB) {
        const int m = NB - DB - j0, r0 = j0 + DB;
        __syncthreads();
        if (warp == 0) inv32_warp(a, j0, lane, dsc);
        double acc[(NB - DB) / 8];
        if (m > 0) {
            // T = W22 L21: thread = (column lane, rows warp + 8 k)
#pragma unroll
            for (int kk = 0; kk < (NB - DB) / 8; ++kk) acc[kk] = 0.0;
            for (int t = 0; t < m; ++t) {
                const double l = a[(r0 + t) * DP + j0 + lane];
#pragma unroll
                for (int kk = 0; kk < (NB - DB) / 8; ++kk) {
                    const int r = warp + 8 * kk;
                    if (r < m && r >= t) acc[kk] = fma(a[(r0 + r) * DP + r0 + t], l, acc[kk]);
                }
            }
        }
        __syncthreads();                       // every read of L21 and W22 done; inv32 finished
        if (m > 0) {
#pragma unroll
            for (int kk = 0; kk < (NB - DB) / 8; ++kk) {
                const int r = warp + 8 * kk;
                if (r < m) a[(r0 + r) * DP + j0 + lane] = acc[kk];
            }
        }
        __syncthreads();
        if (m > 0) {
            // W21 = -T L11^-1: out[r][c] = -sum_{t >= c} T[r][t] Dinv[t][c]
#pragma unroll
            for (int kk = 0; kk < (NB - DB) / 8; ++kk) {
                const int r = warp + 8 * kk;
                double sum = 0.0;
                if (r < m)
                    for (int t = 0; t < DB; ++t) sum = fma(a[(r0 + r) * DP + j0 + t], dsc[t * (DB + 1) + lane], sum);   // Dinv[t][c] = 0 for t < c
                acc[kk] = -sum;
            }
        }
        __syncthreads();
        if (m > 0) {
#pragma unroll
            for (int kk = 0; kk < (NB - DB) / 8; ++kk) {
                const int r = warp + 8 * kk;
                if (r < m) a[(r0 + r) * DP + j0 + lane] = acc[kk];
            }
        }
        for (int e = tid; e < DB * DB; e += 256) a[(j0 + (e >> 5)) * DP + j0 + (e & 31)] = dsc[(e >> 5) * (DB + 1) + (e & 31)];
    }
    __syncthreads();
}

// FACTOR = true : block k of the Cholesky (factor T_kk in place, Winv_k, z_k, log-det, info)
// FACTOR = false: only Winv_k = L_kk^-1 of an already factorised matrix
template <bool FACTOR>
__global__ void __launch_bounds__(256, 1)
k_diag_block(double* __restrict__ M, long m_batch, int ld, int k, double* __restrict__ Winv, long w_batch,
             double* __restrict__ WT, long wt_batch, const double* __restrict__ sv_in /* [R][NB] from k_zdot */,
             double* __restrict__ z, long z_batch, double* __restrict__ acc /* [R][2]: logdet, zz */,
             int* __restrict__ info, const double* __restrict__ part /* [R][SPLIT_MAX][NB*NB] */, int nsplit) {
    extern __shared__ __align__(16) double sm[];
    double* a = sm;                  // [NB][DP]
    double* sv = sm + NB * DP;       // [NB] rhs of the forward solve
    double* zk = sv + NB;            // [NB]
    double* rinv = zk + NB;          // [NB] reciprocal pivots
    double* dsc = rinv + NB;         // [DB][DB+1] inverse of the current 32 x 32 diagonal block
    __shared__ int bad_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = blockIdx.x;
    double* Mr = M + (size_t)r * m_batch;
    double* blk = Mr + (size_t)k * NB * ld + (size_t)k * NB;
    if (tid == 0) bad_s = 0;
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int i = idx >> 7, cidx = idx & 127;
        double v = cidx <= i ? blk[(size_t)i * ld + cidx] : 0.0;
        if (cidx <= i)      // T_kk = A_kk - sum_s (L[k, slice s] L[k, slice s]^T), fixed order
            for (int sp = 0; sp < nsplit; ++sp) v -= part[((size_t)r * SPLIT_MAX + sp) * NB * NB + idx];
        a[i * DP + cidx] = v;
    }
    if constexpr (FACTOR) {
        if (tid < NB) sv[tid] = sv_in[(size_t)r * NB + tid];       // s = y_k - L[k, <k] z_<k (k_zdot)
        chol_block_smem(a, rinv, dsc, &bad_s);
        if (tid == 0 && bad_s != 0 && info[r] == 0) info[r] = k * NB + bad_s;   // not positive definite (sklearn:_gpr.py:591-593)
        for (int idx = tid; idx < NB * NB; idx += 256) {
            const int i = idx >> 7, cidx = idx & 127;
            blk[(size_t)i * ld + cidx] = a[i * DP + cidx];          // L_kk (upper part zero)
        }
        // log-det contribution
        double lg = tid < NB ? log(a[tid * DP + tid]) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
        if (lane == 0) zk[warp] = lg;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < 4; ++w) t += zk[w];
            acc[2 * r] += t;
        }
    }
    tri_invert_smem(a, dsc);
    double* Wr = Winv + (size_t)r * w_batch + (size_t)k * NB * NB;
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int i = idx >> 7, cidx = idx & 127;
        Wr[idx] = a[i * DP + cidx];
    }
    if (WT) {   // WT_kk = Winv_k^T into the transposed inverse
        double* wt = WT + (size_t)r * wt_batch + (size_t)k * NB * ld + (size_t)k * NB;
        for (int idx = tid; idx < NB * NB; idx += 256) {
            const int i = idx >> 7, cidx = idx & 127;
            wt[(size_t)i * ld + cidx] = a[cidx * DP + i];
        }
    }
    if constexpr (FACTOR) {
        // z_k = Winv_k s
        if (tid < NB) {
            double t = 0.0;
            for (int j = 0; j <= tid; ++j) t = fma(a[tid * DP + j], sv[j], t);
            zk[tid] = t;
            z[(size_t)r * z_batch + k * NB + tid] = t;
        }
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int j = 0; j < NB; ++j) t = fma(zk[j], zk[j], t);
            acc[2 * r + 1] += t;
        }
    }
}

__global__ void k_lml_final(const double* __restrict__ acc, const int* __restrict__ info, int n, int R,
                            double* __restrict__ lml) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    // -0.5 y.alpha - sum(log diag L) - n/2 log(2 pi)      (sklearn:_gpr.py:613-615)
    const double v = -0.5 * acc[2 * r + 1] - acc[2 * r] - 0.5 * n * 1.8378770664093453;
    lml[r] = (info[r] != 0 || !(v == v)) ? -INFINITY : v;
}

// alpha = L^-T z = WT z : one warp per row
__global__ void k_alpha_from_wt(const double* __restrict__ WT, int ld, const double* __restrict__ z, int n_pad,
                                double* __restrict__ alpha) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n_pad) return;
    double t = 0.0;
    for (int j = row + lane; j < n_pad; j += 32) t = fma(WT[(size_t)row * ld + j], z[j], t);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) alpha[row] = t;
}

__global__ void k_pad_copy_lower(const double* __restrict__ L, int n, int n_pad, double* __restrict__ M) {
    const size_t total = (size_t)n_pad * n_pad;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx / n_pad), c = (int)(idx % n_pad);
        M[idx] = (r < n && c < n) ? (c <= r ? L[(size_t)r * n + c] : 0.0) : (r == c ? 1.0 : 0.0);
    }
}

// ---------------------------------------------------------------- host orchestration
struct LmlWorkspace {
    Cand* cands; double* M; double* Winv; double* z; double* acc; int* info; double* ypad; double* sv; double* part;
};
static size_t align_up(size_t v) { return (v + 255) / 256 * 256; }

size_t lml_workspace_bytes(int n, int R, bool with_wt) {
    const size_t n_pad = padded_n(n), nP = n_pad / NB;
    size_t b = 0;
    b += align_up(sizeof(Cand) * R);
    b += align_up(sizeof(double) * n_pad * n_pad * R);        // M
    b += align_up(sizeof(double) * nP * NB * NB * R);         // Winv
    b += align_up(sizeof(double) * n_pad * R);                // z
    b += align_up(sizeof(double) * 2 * R);                    // acc
    b += align_up(sizeof(int) * R);                           // info
    b += align_up(sizeof(double) * n_pad * R);                // targets, padded (at most one row per candidate)
    b += align_up(sizeof(double) * NB * R);                   // right-hand side of the block forward solve
    b += align_up(sizeof(double) * SPLIT_MAX * NB * NB * R);  // split-K partials of the diagonal block-row update
    if (with_wt) b += align_up(sizeof(double) * n_pad * n_pad * R);
    return b;
}
static LmlWorkspace carve(void* ws, int n, int R, double** wt_out) {
    const size_t n_pad = padded_n(n), nP = n_pad / NB;
    char* p = (char*)ws;
    LmlWorkspace w;
    w.cands = (Cand*)p; p += align_up(sizeof(Cand) * R);
    w.M = (double*)p; p += align_up(sizeof(double) * n_pad * n_pad * R);
    w.Winv = (double*)p; p += align_up(sizeof(double) * nP * NB * NB * R);
    w.z = (double*)p; p += align_up(sizeof(double) * n_pad * R);
    w.acc = (double*)p; p += align_up(sizeof(double) * 2 * R);
    w.info = (int*)p; p += align_up(sizeof(int) * R);
    w.ypad = (double*)p; p += align_up(sizeof(double) * n_pad * R);
    w.sv = (double*)p; p += align_up(sizeof(double) * NB * R);
    w.part = (double*)p; p += align_up(sizeof(double) * SPLIT_MAX * NB * NB * R);
    if (wt_out) *wt_out = (double*)p;
    return w;
}

static int upload_cands(const LmlWorkspace& w, const std::vector<Cand>& h, cudaStream_t st) {
    BK_CUDA(cudaMemcpyAsync(w.cands, h.data(), sizeof(Cand) * h.size(), cudaMemcpyHostToDevice, st));
    BK_CUDA(cudaStreamSynchronize(st));     // h belongs to the caller's frame
    return 0;
}
static int affine_cands(std::vector<Cand>& h, int R, int p, const int* kinds, const double* c0, const double* c1,
                        const double* kdiag, const double* ls) {
    h.assign(R, Cand());
    for (int r = 0; r < R; ++r) {
        if (kinds[r] < BK_MATERN12 || kinds[r] > BK_MATERN_INF) return set_error("candidate %d: unknown kernel kind %d", r, kinds[r]);
        h[r].kind = kinds[r]; h[r].target = 0; h[r].c0 = c0[r]; h[r].c1 = c1[r]; h[r].kdiag = kdiag[r];
        for (int c = 0; c < MAX_P; ++c) h[r].ls[c] = c < p ? ls[(size_t)r * p + c] : 1.0;
    }
    return 0;
}
// general trees: one shape (nf, fkind, nt, texp) for all candidates; tcoef R x nt, fls R x nf x p, kdiag R
static int general_cands(std::vector<Cand>& h, int R, int p, int nf, const int* fkind, int nt, const int* texp,
                         const double* tcoef, const double* fls, const double* kdiag) {
    h.assign(R, Cand());
    for (int r = 0; r < R; ++r) {
        Cand& cd = h[r];
        cd.kind = BK_GENERAL; cd.target = 0; cd.c0 = 0.0; cd.c1 = 0.0; cd.kdiag = kdiag[r];
        for (int c = 0; c < MAX_P; ++c) cd.ls[c] = 1.0;
        cd.gen.nf = nf; cd.gen.nt = nt;
        for (int f = 0; f < nf; ++f) {
            cd.gen.fkind[f] = fkind[f];
            for (int c = 0; c < MAX_P; ++c) cd.gls[f][c] = c < p ? fls[((size_t)r * nf + f) * p + c] : 1.0;
        }
        for (int t = 0; t < nt; ++t) {
            cd.gen.tcoef[t] = tcoef[(size_t)r * nt + t];
            for (int f = 0; f < nf; ++f) cd.gen.texp[t][f] = (unsigned char)texp[t * nf + f];
        }
    }
    return 0;
}

// Side stream (highest priority) for the latency-bound single-CTA diagonal-block kernel: it runs while the main
// stream updates the block rows below the diagonal (look-ahead), so only the row-k update sits on the critical path.
// Streams of the factorisation.  The diagonal-block kernel (one latency-bound CTA per candidate, 140 KiB of shared
// memory, 146 registers) cannot share an SM with the 200-register GEMM CTAs, so with >= 64 candidates the batch is
// split into two groups that walk the block columns half a step apart, each on its own (main, side) stream pair:
// while one group's diagonal blocks occupy half of the SMs, the other group's GEMMs fill the rest.
struct SideStream {
    cudaStream_t s = nullptr;        // diagonal block + forward-solve step (high priority)
    cudaStream_t main2 = nullptr;    // main stream of the second candidate group
    cudaEvent_t e_main = nullptr, e_side = nullptr, e_fork = nullptr, e_join = nullptr;
};
static int side_streams(SideStream** out) {
    static SideStream per_dev[64][2];
    int dev = 0;
    BK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return set_error("device ordinal %d out of range", dev);
    for (int g = 0; g < 2; ++g) {
        SideStream& ss = per_dev[dev][g];
        if (!ss.s) {
            int lo = 0, hi = 0;
            BK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            BK_CUDA(cudaStreamCreateWithPriority(&ss.s, cudaStreamNonBlocking, hi));
            BK_CUDA(cudaStreamCreateWithFlags(&ss.main2, cudaStreamNonBlocking));
            BK_CUDA(cudaEventCreateWithFlags(&ss.e_main, cudaEventDisableTiming));
            BK_CUDA(cudaEventCreateWithFlags(&ss.e_side, cudaEventDisableTiming));
            BK_CUDA(cudaEventCreateWithFlags(&ss.e_fork, cudaEventDisableTiming));
            BK_CUDA(cudaEventCreateWithFlags(&ss.e_join, cudaEventDisableTiming));
        }
    }
    *out = per_dev[dev];
    return 0;
}

// candidates [r0, r0 + R) of a workspace as a workspace of their own
static LmlWorkspace group_view(const LmlWorkspace& w, int n, int r0) {
    const size_t n_pad = padded_n(n), nP = n_pad / NB;
    LmlWorkspace v = w;
    v.cands = w.cands + r0;
    v.M = w.M + (size_t)r0 * n_pad * n_pad;
    v.Winv = w.Winv + (size_t)r0 * nP * NB * NB;
    v.z = w.z + (size_t)r0 * n_pad;
    v.acc = w.acc + 2 * (size_t)r0;
    v.info = w.info + r0;
    v.sv = w.sv + (size_t)r0 * NB;
    v.part = w.part + (size_t)r0 * SPLIT_MAX * NB * NB;
    return v;                        // ypad is indexed by the candidate's absolute target row
}

// block column k of the Cholesky of the R matrices in w.M (in place), z and the accumulators alongside
static int cholesky_column(const LmlWorkspace& w, int n, int R, int k, double* WT, cudaStream_t st, SideStream* ss) {
    const int n_pad = padded_n(n), nP = n_pad / NB;
    const long mb = (long)n_pad * n_pad, wb = (long)nP * NB * NB;
    int nsplit = 0;
    if (k > 0) {
        // the diagonal block first, split-K over up to SPLIT_MAX CTAs per candidate: partial products go to
        // scratch and are subtracted by the diagonal-block kernel when it loads T_kk
        const int per = (k + SPLIT_MAX - 1) / SPLIT_MAX;               // 128-column blocks per slice
        nsplit = (k + per - 1) / per;
        GemmArgs g;
        g.A = w.M + (size_t)k * NB * n_pad; g.a_batch = mb; g.a_tile = (long)per * NB; g.lda = n_pad;
        g.B = g.A; g.b_batch = mb; g.b_tile = (long)per * NB; g.ldb = n_pad;
        g.C = w.part; g.c_batch = (long)SPLIT_MAX * NB * NB; g.c_tile = (long)NB * NB; g.ldc = NB;
        g.K0 = per * NB; g.k_tile = 0; g.alpha = 1.0; g.beta = 0.0; g.split_k_total = k * NB;
        if (int rc = gemm_nt(g, nsplit, R, st)) return rc;
    }
    BK_CUDA(cudaEventRecord(ss->e_main, st));
    BK_CUDA(cudaStreamWaitEvent(ss->s, ss->e_main, 0));
    {
        ProfScope prof(PROF_CHOL, ss->s);
        k_zdot<<<dim3(R, NB / 8), 256, 0, ss->s>>>(w.M, mb, n_pad, k, w.cands, w.ypad, w.z, n_pad, w.sv);
        BK_LAUNCH_CHECK("k_zdot");
        k_diag_block<true><<<R, 256, DIAG_SMEM, ss->s>>>(w.M, mb, n_pad, k, w.Winv, wb, WT, mb, w.sv, w.z, n_pad, w.acc, w.info,
                                                         w.part, nsplit);
        BK_LAUNCH_CHECK("k_diag_block<true>");
    }
    BK_CUDA(cudaEventRecord(ss->e_side, ss->s));
    if (k > 0 && k + 1 < nP) {
        // T_ik = A_ik - L[i, <k] L[k, <k]^T for the block rows below the diagonal: overlaps the diagonal block
        GemmArgs g;
        g.A = w.M + (size_t)(k + 1) * NB * n_pad; g.a_batch = mb; g.a_tile = (long)NB * n_pad; g.lda = n_pad;
        g.B = w.M + (size_t)k * NB * n_pad; g.b_batch = mb; g.b_tile = 0; g.ldb = n_pad;
        g.C = w.M + (size_t)(k + 1) * NB * n_pad + (size_t)k * NB; g.c_batch = mb; g.c_tile = (long)NB * n_pad; g.ldc = n_pad;
        g.K0 = k * NB; g.k_tile = 0; g.alpha = -1.0; g.beta = 1.0;
        if (int rc = gemm_nt(g, nP - k - 1, R, st)) return rc;
    }
    BK_CUDA(cudaStreamWaitEvent(st, ss->e_side, 0));
    if (k + 1 < nP) {   // L_ik = T_ik Winv_k^T, i = k+1..nP-1 (in place)
        GemmArgs g;
        g.A = w.M + (size_t)(k + 1) * NB * n_pad + (size_t)k * NB; g.a_batch = mb; g.a_tile = (long)NB * n_pad; g.lda = n_pad;
        g.B = w.Winv + (size_t)k * NB * NB; g.b_batch = wb; g.b_tile = 0; g.ldb = NB;
        g.C = const_cast<double*>(g.A); g.c_batch = mb; g.c_tile = (long)NB * n_pad; g.ldc = n_pad;
        g.K0 = NB; g.k_tile = 0; g.alpha = 1.0; g.beta = 0.0;
        if (int rc = gemm_nt(g, nP - k - 1, R, st)) return rc;
    }
    return 0;
}

// blocked Cholesky of the R matrices in w.M (in place)
static int cholesky_sweep(const LmlWorkspace& w, int n, int R, double* WT, cudaStream_t st) {
    const int nP = padded_n(n) / NB;
    static OncePerDevice attr;
    if (attr.first()) {
        BK_CUDA(cudaFuncSetAttribute(k_diag_block<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
        BK_CUDA(cudaFuncSetAttribute(k_diag_block<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
    }
    SideStream* ss = nullptr;
    if (int rc = side_streams(&ss)) return rc;
    BK_CUDA(cudaMemsetAsync(w.acc, 0, sizeof(double) * 2 * R, st));
    BK_CUDA(cudaMemsetAsync(w.info, 0, sizeof(int) * R, st));
    if (R < 64 || WT != nullptr) {
        for (int k = 0; k < nP; ++k)
            if (int rc = cholesky_column(w, n, R, k, WT, st, &ss[0])) return rc;
        return 0;
    }
    const int R0 = (R + 1) / 2;
    const LmlWorkspace w1 = group_view(w, n, R0);
    cudaStream_t st1 = ss[1].main2;
    BK_CUDA(cudaEventRecord(ss[1].e_fork, st));
    BK_CUDA(cudaStreamWaitEvent(st1, ss[1].e_fork, 0));
    for (int k = 0; k < nP; ++k) {
        if (int rc = cholesky_column(w, n, R0, k, nullptr, st, &ss[0])) return rc;
        // stagger: the second group starts when the first group's first diagonal block is done, so that from then on
        // one group's diagonal blocks run beside the other group's GEMMs instead of beside its own twin
        if (k == 0) BK_CUDA(cudaStreamWaitEvent(st1, ss[0].e_side, 0));
        if (int rc = cholesky_column(w1, n, R - R0, k, nullptr, st1, &ss[1])) return rc;
    }
    BK_CUDA(cudaEventRecord(ss[1].e_join, st1));
    BK_CUDA(cudaStreamWaitEvent(st, ss[1].e_join, 0));
    return 0;
}

// WT = L^-T (upper triangular, row-major) from L in M and the inverted diagonal blocks in Winv (WT_ii already set)
static int trtri_sweep(const double* M, const double* Winv, double* WT, int n, cudaStream_t st) {
    const int n_pad = padded_n(n), nP = n_pad / NB;
    for (int i = 1; i < nP; ++i) {
        GemmArgs g;   // S_ki = sum_{j=k}^{i-1} WT_kj L_ij^T  for k = 0..i-1   (K shrinks with k)
        g.A = WT; g.a_batch = 0; g.a_tile = (long)NB * n_pad + NB; g.lda = n_pad;
        g.B = M + (size_t)i * NB * n_pad; g.b_batch = 0; g.b_tile = NB; g.ldb = n_pad;
        g.C = WT + (size_t)i * NB; g.c_batch = 0; g.c_tile = (long)NB * n_pad; g.ldc = n_pad;
        g.K0 = i * NB; g.k_tile = -NB; g.alpha = 1.0; g.beta = 0.0;
        if (int rc = gemm_nt(g, i, 1, st)) return rc;
        GemmArgs h;   // WT_ki = -S_ki Winv_i^T (in place)
        h.A = WT + (size_t)i * NB; h.a_batch = 0; h.a_tile = (long)NB * n_pad; h.lda = n_pad;
        h.B = Winv + (size_t)i * NB * NB; h.b_batch = 0; h.b_tile = 0; h.ldb = NB;
        h.C = WT + (size_t)i * NB; h.c_batch = 0; h.c_tile = (long)NB * n_pad; h.ldc = n_pad;
        h.K0 = NB; h.k_tile = 0; h.alpha = -1.0; h.beta = 0.0;
        if (int rc = gemm_nt(h, i, 1, st)) return rc;
    }
    return 0;
}

int launch_pack_w_ex(const double* w, int n_rows_valid, int ld, int transposed, double* out, cudaStream_t st);
int launch_slice_w(const double* src, int n, int ld, int transposed, int scheme, int8_t* wq, double* wsig, cudaStream_t st);

static int build_K(const LmlWorkspace& w, const double* X, int n, int p, int R, int nf, cudaStream_t st) {
    const int n_pad = padded_n(n), nP = n_pad / NB;
    const size_t smem = (size_t)2 * nf * NB * p * 8;
    if (smem > 48 * 1024) BK_CUDA(cudaFuncSetAttribute(k_build_K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope prof(PROF_CHOL, st);
    k_build_K<<<dim3(nP * (nP + 1) / 2, R), 256, smem, st>>>(X, n, p, n_pad, w.cands, w.M, (long)n_pad * n_pad);
    BK_LAUNCH_CHECK("k_build_K");
    return 0;
}

// y: n_targets x n row-major -> ypad: n_targets x n_pad, zero padded
static int stage_y(const LmlWorkspace& w, const double* y, int n, int n_targets, cudaStream_t st) {
    const int n_pad = padded_n(n);
    BK_CUDA(cudaMemsetAsync(w.ypad, 0, sizeof(double) * n_pad * n_targets, st));
    BK_CUDA(cudaMemcpy2DAsync(w.ypad, sizeof(double) * n_pad, y, sizeof(double) * n, sizeof(double) * n, n_targets,
                              cudaMemcpyDeviceToDevice, st));
    return 0;
}

static int lml_cands(const double* X, const double* y, int n_targets, const int* target, int n, int p,
                     std::vector<Cand>& cands, int nf, double* lml_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    const int R = (int)cands.size();
    if (n_targets < 1 || n_targets > R) return set_error("bk_lml_batched: n_targets = %d must be in [1, n_candidates = %d]", n_targets, R);
    for (int r = 0; r < R; ++r) {
        const int t = target ? target[r] : 0;
        if (t < 0 || t >= n_targets) return set_error("bk_lml_batched: candidate %d scores target %d of %d", r, t, n_targets);
        cands[r].target = t;
    }
    if (p > MAX_P) return set_error("p = %d exceeds %d", p, MAX_P);
    if (ws_bytes < lml_workspace_bytes(n, R, false))
        return set_error("bk_lml_batched: workspace of %zu bytes needed (got %zu)", lml_workspace_bytes(n, R, false), ws_bytes);
    LmlWorkspace w = carve(ws, n, R, nullptr);
    if (int rc = upload_cands(w, cands, st)) return rc;
    if (int rc = stage_y(w, y, n, n_targets, st)) return rc;
    if (int rc = build_K(w, X, n, p, R, nf, st)) return rc;
    if (int rc = cholesky_sweep(w, n, R, nullptr, st)) return rc;
    k_lml_final<<<(R + 127) / 128, 128, 0, st>>>(w.acc, w.info, n, R, lml_out);
    BK_LAUNCH_CHECK("k_lml_final");
    return 0;
}

int lml_batched(const double* X, const double* y, int n_targets, const int* target, int n, int p, int R,
                const int* kinds, const double* c0, const double* c1, const double* kdiag, const double* ls,
                double* lml_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    std::vector<Cand> cands;
    if (int rc = affine_cands(cands, R, p, kinds, c0, c1, kdiag, ls)) return rc;
    return lml_cands(X, y, n_targets, target, n, p, cands, 1, lml_out, ws, ws_bytes, st);
}

int lml_batched_general(const double* X, const double* y, int n_targets, const int* target, int n, int p, int R,
                        int nf, const int* fkind, int nt, const int* texp, const double* tcoef, const double* fls,
                        const double* kdiag, double* lml_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    std::vector<Cand> cands;
    if (int rc = general_cands(cands, R, p, nf, fkind, nt, texp, tcoef, fls, kdiag)) return rc;
    return lml_cands(X, y, n_targets, target, n, p, cands, nf, lml_out, ws, ws_bytes, st);
}

// final state at the chosen theta: L (n_pad x n_pad, row-major lower), alpha (n_pad), packed W tiles, lml, info
static int factorise_cand(const double* X, const double* y, int n, int p, const std::vector<Cand>& cand, int nf,
                          double* L_out, double* alpha_out, double* wt_tiles_out, int8_t* wq_out, double* wsig_out,
                          int wq_scheme, double* lml_out, int* info_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (p > MAX_P) return set_error("p = %d exceeds %d", p, MAX_P);
    if (ws_bytes < lml_workspace_bytes(n, 1, true))
        return set_error("bk_gp_factorise: workspace of %zu bytes needed (got %zu)", lml_workspace_bytes(n, 1, true), ws_bytes);
    const int n_pad = padded_n(n);
    double* WT = nullptr;
    LmlWorkspace w = carve(ws, n, 1, &WT);
    if (int rc = upload_cands(w, cand, st)) return rc;
    if (int rc = stage_y(w, y, n, 1, st)) return rc;
    if (int rc = build_K(w, X, n, p, 1, nf, st)) return rc;
    BK_CUDA(cudaMemsetAsync(WT, 0, sizeof(double) * n_pad * n_pad, st));
    if (int rc = cholesky_sweep(w, n, 1, WT, st)) return rc;
    k_lml_final<<<1, 128, 0, st>>>(w.acc, w.info, n, 1, lml_out);
    BK_LAUNCH_CHECK("k_lml_final");
    BK_CUDA(cudaMemcpyAsync(info_out, w.info, sizeof(int), cudaMemcpyDeviceToDevice, st));
    if (int rc = trtri_sweep(w.M, w.Winv, WT, n, st)) return rc;
    k_alpha_from_wt<<<(n_pad + 7) / 8, 256, 0, st>>>(WT, n_pad, w.z, n_pad, alpha_out);
    BK_LAUNCH_CHECK("k_alpha_from_wt");
    BK_CUDA(cudaMemcpyAsync(L_out, w.M, sizeof(double) * n_pad * n_pad, cudaMemcpyDeviceToDevice, st));
    if (wt_tiles_out)
        if (int rc = launch_pack_w_ex(WT, n, n_pad, 1, wt_tiles_out, st)) return rc;   // padded rows/cols -> 0
    if (wq_out && wsig_out)
        if (int rc = launch_slice_w(WT, n, n_pad, 1, wq_scheme, wq_out, wsig_out, st)) return rc;
    return 0;
}

int gp_factorise(const double* X, const double* y, int n, int p, int kind, double c0, double c1, double kdiag,
                 const double* ls, double* L_out, double* alpha_out, double* wt_tiles_out, int8_t* wq_out,
                 double* wsig_out, int wq_scheme, double* lml_out, int* info_out, void* ws, size_t ws_bytes,
                 cudaStream_t st) {
    std::vector<Cand> cand;
    if (int rc = affine_cands(cand, 1, p, &kind, &c0, &c1, &kdiag, ls)) return rc;
    return factorise_cand(X, y, n, p, cand, 1, L_out, alpha_out, wt_tiles_out, wq_out, wsig_out, wq_scheme, lml_out,
                          info_out, ws, ws_bytes, st);
}

int gp_factorise_general(const double* X, const double* y, int n, int p, int nf, const int* fkind, int nt,
                         const int* texp, const double* tcoef, const double* fls, double kdiag, double* L_out,
                         double* alpha_out, double* wt_tiles_out, int8_t* wq_out, double* wsig_out, int wq_scheme,
                         double* lml_out, int* info_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    std::vector<Cand> cand;
    if (int rc = general_cands(cand, 1, p, nf, fkind, nt, texp, tcoef, fls, &kdiag)) return rc;
    return factorise_cand(X, y, n, p, cand, nf, L_out, alpha_out, wt_tiles_out, wq_out, wsig_out, wq_scheme, lml_out,
                          info_out, ws, ws_bytes, st);
}

// W tiles from a given Cholesky factor (models adopted from scikit-learn state)
int tri_inverse_pack(const double* L, int n, double* wt_tiles_out, int8_t* wq_out, double* wsig_out, int wq_scheme,
                     void* ws, size_t ws_bytes, cudaStream_t st) {
    if (ws_bytes < lml_workspace_bytes(n, 1, true))
        return set_error("bk_tri_inverse_pack: workspace of %zu bytes needed (got %zu)", lml_workspace_bytes(n, 1, true), ws_bytes);
    const int n_pad = padded_n(n), nP = n_pad / NB;
    double* WT = nullptr;
    LmlWorkspace w = carve(ws, n, 1, &WT);
    static OncePerDevice attr;
    if (attr.first()) {
        BK_CUDA(cudaFuncSetAttribute(k_diag_block<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
    }
    {
        ProfScope prof(PROF_CHOL, st);
        k_pad_copy_lower<<<1184, 256, 0, st>>>(L, n, n_pad, w.M);
        BK_LAUNCH_CHECK("k_pad_copy_lower");
    }
    BK_CUDA(cudaMemsetAsync(WT, 0, sizeof(double) * n_pad * n_pad, st));
    for (int k = 0; k < nP; ++k) {
        ProfScope prof(PROF_CHOL, st);
        k_diag_block<false><<<1, 256, DIAG_SMEM, st>>>(w.M, 0, n_pad, k, w.Winv, 0, WT, 0, nullptr, nullptr, 0, nullptr, nullptr,
                                                       nullptr, 0);
        BK_LAUNCH_CHECK("k_diag_block<false>");
    }
    if (int rc = trtri_sweep(w.M, w.Winv, WT, n, st)) return rc;
    if (wq_out && wsig_out)
        if (int rc = launch_slice_w(WT, n, n_pad, 1, wq_scheme, wq_out, wsig_out, st)) return rc;
    return launch_pack_w_ex(WT, n, n_pad, 1, wt_tiles_out, st);   // padded rows/cols -> 0
}

}  // namespace bk
