// Statically scheduled fused Saltelli kernel (P <= 12).
//
// Same contract as k_sobol (sobol.cu) but the design-block structure is compile time: logical block lb is
//   0 = A, 1 = B, 2+i = E^i (A with column i from B), 2+P+i = C^i (B with column i from A),  i < P,
// so which of ta_c / tb_c enters each distance sum is a constant and the re-summation in cdist order is a pure
// DADD chain -- no 64-bit selects on the ALU pipe (29 % ALU and "wait" stalls in the ncu capture of the generic
// kernel), identical prefixes (ta_0 + ... + ta_{i-1}) are common sub-expressions, and the sqrt / exp chains of the
// blocks of a pass sit in one basic block so ptxas can interleave them.  Logical blocks with i >= p (input
// dimension padded up to P) are evaluated and dropped: at most 2 of 2P+2.
#include <utility>

#include "sobol_common.cuh"

namespace bk {

constexpr int SS_CH = 6;   // logical blocks per pass over the training set

template <int P>
struct StaticCtx {
    const SobolParams<P>* prm;
    double* xs_s; double* al_s; double* y_s; uint64_t* bar;
    long tctr, total_loads;
    int ntiles, tid;
    bool valid;
    long g;
};

template <int P>
__device__ __forceinline__ void ss_issue(const StaticCtx<P>& cx, long t) {
    constexpr uint32_t XS_BYTES = SOB_TJ * P * 8, AL_BYTES = SOB_TJ * 8;
    const int s = (int)(t & 1);
    const int tile = (int)(t % cx.ntiles);
    mbar_expect_tx(&cx.bar[s], XS_BYTES + AL_BYTES);
    tma_load_1d(cx.xs_s + s * SOB_TJ * P, cx.prm->xs + (size_t)tile * SOB_TJ * P, XS_BYTES, &cx.bar[s]);
    tma_load_1d(cx.al_s + s * SOB_TJ, cx.prm->alpha + (size_t)tile * SOB_TJ, AL_BYTES, &cx.bar[s]);
}

// squared scaled distance of logical block LB in scipy-cdist order (sequential, non-fused)
template <int P, int LB>
__device__ __forceinline__ double ss_block_sum(const double (&ta)[P], const double (&tb)[P]) {
    constexpr bool fromB = (LB == 1) || (LB >= 2 + P);
    constexpr int swp = LB < 2 ? -1 : (LB < 2 + P ? LB - 2 : LB - 2 - P);
    double s = ((0 == swp) != fromB) ? tb[0] : ta[0];          // 0.0 + x == x exactly
#pragma unroll
    for (int c = 1; c < P; ++c) s = __dadd_rn(s, ((c == swp) != fromB) ? tb[c] : ta[c]);
    return s;
}

template <int P, int B0, int... BB>
__device__ __forceinline__ void ss_all_sums(double* kv, const double (&ta)[P], const double (&tb)[P],
                                            std::integer_sequence<int, BB...>) {
    ((kv[BB] = ss_block_sum<P, B0 + BB>(ta, tb)), ...);
}

template <int KIND, int P, int B0, int CNT>
__device__ __forceinline__ void ss_pass(StaticCtx<P>& cx, const double (&ua)[P], const double (&ub)[P]) {
    const SobolParams<P>& prm = *cx.prm;
    double acc[CNT], comp[CNT];
#pragma unroll
    for (int bb = 0; bb < CNT; ++bb) { acc[bb] = 0.0; comp[bb] = 0.0; }
    for (int tile = 0; tile < cx.ntiles; ++tile, ++cx.tctr) {
        const int s = (int)(cx.tctr & 1);
        mbar_wait(&cx.bar[s], (uint32_t)((cx.tctr >> 1) & 1));
        const double* xrow = cx.xs_s + s * SOB_TJ * P;
        const double* arow = cx.al_s + s * SOB_TJ;
#pragma unroll 2   // two training rows in flight: 2 x CNT independent sqrt/exp chains for ptxas to interleave
        for (int j = 0; j < SOB_TJ; ++j) {
            double ta[P], tb[P];
#pragma unroll
            for (int c = 0; c < P; c += 2) {
                const double2 vv = *reinterpret_cast<const double2*>(xrow + j * P + c);
                double d;
                d = __dadd_rn(ua[c], -vv.x); ta[c] = __dmul_rn(d, d);
                d = __dadd_rn(ub[c], -vv.x); tb[c] = __dmul_rn(d, d);
                d = __dadd_rn(ua[c + 1], -vv.y); ta[c + 1] = __dmul_rn(d, d);
                d = __dadd_rn(ub[c + 1], -vv.y); tb[c + 1] = __dmul_rn(d, d);
            }
            const double al = arow[j];
            double kv[CNT];
            // stage-wise over the blocks of the pass: independent chains side by side
            ss_all_sums<P, B0>(kv, ta, tb, std::make_integer_sequence<int, CNT>{});
#pragma unroll
            for (int bb = 0; bb < CNT; ++bb) kv[bb] = affine(prm.c0, prm.c1, stationary<KIND>(kv[bb]));
#pragma unroll
            for (int bb = 0; bb < CNT; ++bb) {
                const double y = __fma_rn(kv[bb], al, -comp[bb]);
                const double tsum = __dadd_rn(acc[bb], y);
                comp[bb] = __dadd_rn(__dadd_rn(tsum, -acc[bb]), -y);
                acc[bb] = tsum;
            }
        }
        __syncthreads();
        if (cx.tid == 0 && cx.tctr + 2 < cx.total_loads) ss_issue<P>(cx, cx.tctr + 2);
    }
#pragma unroll
    for (int bb = 0; bb < CNT; ++bb) {
        const int lb = B0 + bb;
        const int i = lb < 2 ? -1 : (lb < 2 + P ? lb - 2 : lb - 2 - P);
        if (i < prm.p) {                                     // drop blocks of padded dimensions
            const int b = lb < 2 + P ? (lb < 2 ? lb : 2 + i) : 2 + prm.p + i;
            const double y = __dadd_rn(__dmul_rn(prm.y_std, acc[bb]), prm.y_mean);
            if (prm.yout && cx.valid) prm.yout[((size_t)b * prm.count + cx.g) * prm.m + prm.q] = y;
            if (prm.partial) cx.y_s[b * SOB_THREADS + cx.tid] = cx.valid ? __dadd_rn(y, -prm.shift) : 0.0;
        }
    }
}

template <int KIND, int P, int NLB, int B0>
__device__ __forceinline__ void ss_passes(StaticCtx<P>& cx, const double (&ua)[P], const double (&ub)[P]) {
    if constexpr (B0 < NLB) {
        constexpr int CNT = (NLB - B0) < SS_CH ? (NLB - B0) : SS_CH;
        ss_pass<KIND, P, B0, CNT>(cx, ua, ub);
        ss_passes<KIND, P, NLB, B0 + SS_CH>(cx, ua, ub);
    }
}

template <int KIND, int P, bool WITH_C>
__global__ void __launch_bounds__(SOB_THREADS) k_sobol_static(const SobolParams<P> prm) {
    constexpr int NLB = WITH_C ? 2 * P + 2 : P + 2;
    constexpr int NPASS = (NLB + SS_CH - 1) / SS_CH;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* xs_s = reinterpret_cast<double*>(smem_raw);               // [2][TJ*P]
    double* al_s = xs_s + 2 * SOB_TJ * P;                             // [2][TJ]
    double* y_s = al_s + 2 * SOB_TJ;                                  // [nb][128]
    double* wsum = y_s + (size_t)prm.nb * SOB_THREADS;                // [4][nsums][2]
    uint64_t* bar = reinterpret_cast<uint64_t*>(wsum + 4 * prm.nsums * 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long n_rowtiles = (prm.count + SOB_THREADS - 1) / SOB_THREADS;
    const long my_rowtiles = blockIdx.x < n_rowtiles ? (n_rowtiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    StaticCtx<P> cx;
    cx.prm = &prm; cx.xs_s = xs_s; cx.al_s = al_s; cx.y_s = y_s; cx.bar = bar;
    cx.ntiles = prm.n_pad / SOB_TJ; cx.tid = tid; cx.tctr = 0;
    cx.total_loads = my_rowtiles * NPASS * cx.ntiles;

    for (int i = tid; i < 4 * prm.nsums * 2; i += SOB_THREADS) wsum[i] = 0.0;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0)
        for (long t = 0; t < 2 && t < cx.total_loads; ++t) ss_issue<P>(cx, t);

    for (long rt = blockIdx.x; rt < n_rowtiles; rt += gridDim.x) {
        const long g = rt * SOB_THREADS + tid;
        cx.valid = g < prm.count;
        cx.g = g;
        const long gc = cx.valid ? g : prm.count - 1;
        double ua[P], ub[P];
        {
            double xa[P], xb[P];
            if (prm.a) {
#pragma unroll
                for (int c = 0; c < P; ++c) {
                    xa[c] = c < prm.p ? prm.a[gc * prm.p + c] : 0.0;
                    xb[c] = c < prm.p ? prm.b[gc * prm.p + c] : 0.0;
                }
            } else {
                design_row<P>(prm.seed, (uint64_t)(prm.row_lo + gc), prm.p, 0, prm.marg, xa);
                design_row<P>(prm.seed, (uint64_t)(prm.row_lo + gc), prm.p, 1, prm.marg, xb);
            }
#pragma unroll
            for (int c = 0; c < P; ++c) {
                ua[c] = __ddiv_rn(__dadd_rn(__dmul_rn(xa[c], prm.scale[c]), prm.mn[c]), prm.ls[c]);
                ub[c] = __ddiv_rn(__dadd_rn(__dmul_rn(xb[c], prm.scale[c]), prm.mn[c]), prm.ls[c]);
            }
        }
        ss_passes<KIND, P, NLB, 0>(cx, ua, ub);
        if (prm.partial)
            saltelli_products([&](int b) { return y_s[b * SOB_THREADS + tid]; }, prm.p, prm.second, lane,
                              wsum + (size_t)warp * prm.nsums * 2, cx.valid ? 1.0 : 0.0);
    }
    __syncthreads();
    if (prm.partial)
        flush_slot(wsum, SOB_THREADS / 32, prm.nsums, prm.partial + (size_t)blockIdx.x * prm.slot_stride);
}

template <int KIND, int P>
static int launch_static_kp(const SobolParams<P>& prm, size_t smem, cudaStream_t st) {
    const bool with_c = prm.nb > prm.p + 2;
    ProfScope prof(PROF_SOBOL, st);
    if (with_c) {
        BK_CUDA(cudaFuncSetAttribute(k_sobol_static<KIND, P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_sobol_static<KIND, P, true><<<SOB_SLOTS, SOB_THREADS, smem, st>>>(prm);
    } else {
        BK_CUDA(cudaFuncSetAttribute(k_sobol_static<KIND, P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_sobol_static<KIND, P, false><<<SOB_SLOTS, SOB_THREADS, smem, st>>>(prm);
    }
    BK_LAUNCH_CHECK("k_sobol_static");
    return 0;
}

// returns -1 when (kind, P) has no static instantiation (caller falls through to the generic kernel)
template <int P>
int launch_sobol_static(int kind, const SobolParams<P>& prm, size_t smem, cudaStream_t st) {
    if constexpr (P <= 12) {
        switch (kind) {
            case BK_MATERN12: return launch_static_kp<BK_MATERN12, P>(prm, smem, st);
            case BK_MATERN32: return launch_static_kp<BK_MATERN32, P>(prm, smem, st);
            case BK_MATERN52: return launch_static_kp<BK_MATERN52, P>(prm, smem, st);
            case BK_RBF: return launch_static_kp<BK_RBF, P>(prm, smem, st);
            case BK_MATERN_INF: return launch_static_kp<BK_MATERN_INF, P>(prm, smem, st);
        }
    }
    return -1;
}

template int launch_sobol_static<2>(int, const SobolParams<2>&, size_t, cudaStream_t);
template int launch_sobol_static<4>(int, const SobolParams<4>&, size_t, cudaStream_t);
template int launch_sobol_static<6>(int, const SobolParams<6>&, size_t, cudaStream_t);
template int launch_sobol_static<8>(int, const SobolParams<8>&, size_t, cudaStream_t);
template int launch_sobol_static<10>(int, const SobolParams<10>&, size_t, cudaStream_t);
template int launch_sobol_static<12>(int, const SobolParams<12>&, size_t, cudaStream_t);

}  // namespace bk
