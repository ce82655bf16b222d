// K2 `gp_var_trmm`: predictive variance from the packed inverse Cholesky factor.
//
// Replaces  V = solve_triangular(L_, K_trans.T, lower=True);
//           y_var = kernel_.diag(X) - einsum("ij,ji->i", V.T, V); clamp; * y_std^2; sqrt
// (sklearn:_gpr.py:460-500).  The per-point forward substitution (O(n^2), BLAS-2,
// sequential in the row index) is re-cast as a triangular matrix product with
// W = L^-1 (inverted once at upload):  V = W K*^T, and only the column sums of
// squares of V are kept -- V never exists in memory.  oracle experiments put
// the W-form within 1e-14 (normalised variance) of the long-double solve even
// at cond(K) = 5e12 ("Why W = L^-1").
//
// One CTA owns a tile of 128 test points and sweeps ALL row panels of W (so
// every CTA does identical work and no cross-CTA reduction exists).  Per row
// panel it runs a 128x128 output tile over k = 0..(panel+1)*128 with
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4): 8 warps as 2(M) x 4(N), warp tile
// 64x32 = 32 accumulator fragments.  Operands are staged by cp.async.bulk
// (UBLKCP) through a 6-stage full/empty mbarrier ring (no block-wide barrier in the sweep; lane 0 of warp 0
// is the producer); both W and K* live in HBM already
// in fragment-major order, so each 16 KiB stage is ONE contiguous bulk copy and
// every fragment load is a conflict-free LDS.128.
//
// Packed tile layout (shared by W row panels and K* test tiles), element (r, j)
// of a 128 x 16 tile, r = row of W / test point, j = train index:
//     off = (((r/8)*2 + (j/8)) * 8 + r%8) * 8 + (j%4)*2 + (j%8)/4
// i.e. [m-tile 16][k8 2][lane = (r%8)*4 + j%4][2], lane.x -> j%8 < 4, lane.y -> j%8 >= 4.
// Tile (panel, jb) sits at ((panel * n_pad/16) + jb) * 2048 doubles.
//
// Bound: FP64 tensor pipe.  Algorithmic flops per prediction: n^2 + 2n.  Executed: the strictly
// lower 128x128 blocks in full, the diagonal blocks only for the 8x8 sub-blocks on or below the diagonal
// (72 of 128 sub-steps on the busier warp row).
#include "common.cuh"
#include "model.h"

namespace bk {

constexpr int VAR_CONSUMERS = 8;                       // DMMA warps (2 per SM sub-partition: 16 K registers each)
constexpr int VAR_THREADS = VAR_CONSUMERS * 32;        // lane 0 of warp 0 doubles as the TMA producer
constexpr int VAR_STAGES = 6;
constexpr int VAR_LAG = 2;                             // a stage is refilled 2 steps after warp 0 left it
constexpr size_t VAR_STAGE_BYTES = 2 * TILE_DOUBLES * 8;   // A + B
constexpr size_t VAR_SMEM = VAR_STAGES * VAR_STAGE_BYTES + 2 * TILE_T * 8 + 2 * VAR_STAGES * 8;

// one k8 step (two DMMA k4 steps) of a warp tile, m-tiles MT_LO..7 only
template <int MT_LO>
__device__ __forceinline__ void var_k8_step(double (&c)[8][4][2], const double2* __restrict__ As,
                                            const double2* __restrict__ Bs, int wm, int wn, int k8, int lane) {
    double2 a[8], b[4];
#pragma unroll
    for (int mt = MT_LO; mt < 8; ++mt) a[mt] = As[(((mt * 2 + wm) * 2 + k8) * 32) + lane];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) b[nt] = Bs[(((wn * 4 + nt) * 2 + k8) * 32) + lane];
#pragma unroll
    for (int mt = MT_LO; mt < 8; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            dmma884(c[mt][nt][0], c[mt][nt][1], a[mt].x, b[nt].x);
            dmma884(c[mt][nt][0], c[mt][nt][1], a[mt].y, b[nt].y);
        }
}

__global__ void __launch_bounds__(VAR_THREADS, 1)
k_var_trmm(const double* __restrict__ wt, const double* __restrict__ kstar, int n_pad, double kdiag, double y_std,
           double* __restrict__ std_out, long k, int ld, int q) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage = reinterpret_cast<double*>(smem_raw);
    double* red = stage + VAR_STAGES * 2 * TILE_DOUBLES;                 // [2][128]
    uint64_t* full = reinterpret_cast<uint64_t*>(red + 2 * TILE_T);      // [STAGES] TMA -> consumers
    uint64_t* empty = full + VAR_STAGES;                                 // [STAGES] consumers -> TMA

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int n_panels = n_pad / TILE_M;
    const int njb = n_pad / TILE_K;
    const int total_steps = (TILE_M / TILE_K) * n_panels * (n_panels + 1) / 2;
    const double* kt = kstar + (size_t)blockIdx.x * njb * TILE_DOUBLES;

    if (tid == 0) {
        for (int s = 0; s < VAR_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], VAR_CONSUMERS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    double colsum[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) colsum[nt][0] = colsum[nt][1] = 0.0;

    // producer cursor (lane 0 of warp 0): streams (W tile, K* tile) pairs through the ring
    int p_step = 0, p_panel = 0, p_jb = 0;
    auto produce = [&]() {
        const int s = p_step % VAR_STAGES;
        if (p_step >= VAR_STAGES) mbar_wait(&empty[s], ((p_step / VAR_STAGES) - 1) & 1);
        double* dst = stage + (size_t)s * 2 * TILE_DOUBLES;
        mbar_expect_tx(&full[s], (uint32_t)VAR_STAGE_BYTES);
        tma_load_1d(dst, wt + ((size_t)p_panel * njb + p_jb) * TILE_DOUBLES, TILE_DOUBLES * 8, &full[s]);
        tma_load_1d(dst + TILE_DOUBLES, kt + (size_t)p_jb * TILE_DOUBLES, TILE_DOUBLES * 8, &full[s]);
        ++p_step;
        if (++p_jb == (p_panel + 1) * (TILE_M / TILE_K)) { p_jb = 0; ++p_panel; }
    };
    if (tid == 0)
        for (int s = 0; s < VAR_STAGES - VAR_LAG && p_step < total_steps; ++s) produce();

    {
        // ===== no block-wide barrier inside the sweep: each warp frees its stage itself (empty[s]) and the
        // producer lane refills a stage VAR_LAG steps after its own warp left it, so it practically never waits =====
        int step = 0;
        for (int panel = 0; panel < n_panels; ++panel) {
            double c[8][4][2];
#pragma unroll
            for (int mt = 0; mt < 8; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) c[mt][nt][0] = c[mt][nt][1] = 0.0;

            const int nsteps = (panel + 1) * (TILE_M / TILE_K);
            for (int jb = 0; jb < nsteps; ++jb, ++step) {
                const int s = step % VAR_STAGES;
                if (tid == 0 && p_step < total_steps) produce();
                __syncwarp();
                mbar_wait(&full[s], (step / VAR_STAGES) & 1);
                const double2* As = reinterpret_cast<const double2*>(stage + (size_t)s * 2 * TILE_DOUBLES);
                const double2* Bs = As + TILE_DOUBLES / 2;
                // Diagonal block of the panel: W is lower triangular, so the 8x8 sub-block (m-tile, k8-step) is
                // identically zero when its rows lie above its columns.  M-tiles are dealt to the two warp rows
                // round-robin (m-tile = 2*mt + wm) so both carry the same share of the surviving sub-blocks.  The
                // skip is a warp-uniform BRANCH to a shorter unrolled body (a predicated-off DMMA still costs).
                const int kk8_base = (jb - panel * (TILE_M / TILE_K)) * 2;   // >= 0 only inside the diagonal block
                if (kk8_base + 1 <= wm) {
                    var_k8_step<0>(c, As, Bs, wm, wn, 0, lane);
                    var_k8_step<0>(c, As, Bs, wm, wn, 1, lane);
                } else {
#pragma unroll
                    for (int k8 = 0; k8 < 2; ++k8) {
                        const int kk8 = kk8_base + k8;
                        const int mt_lo = kk8 > wm ? (kk8 - wm + 1) >> 1 : 0;    // first mt with 2*mt + wm >= kk8
                        switch (mt_lo) {
                            case 0: var_k8_step<0>(c, As, Bs, wm, wn, k8, lane); break;
                            case 1: var_k8_step<1>(c, As, Bs, wm, wn, k8, lane); break;
                            case 2: var_k8_step<2>(c, As, Bs, wm, wn, k8, lane); break;
                            case 3: var_k8_step<3>(c, As, Bs, wm, wn, k8, lane); break;
                            case 4: var_k8_step<4>(c, As, Bs, wm, wn, k8, lane); break;
                            case 5: var_k8_step<5>(c, As, Bs, wm, wn, k8, lane); break;
                            case 6: var_k8_step<6>(c, As, Bs, wm, wn, k8, lane); break;
                            case 7: var_k8_step<7>(c, As, Bs, wm, wn, k8, lane); break;
                            default: break;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);   // this warp's fragment loads of stage s are done
            }
            // fold this row panel into the running column sums of squares
#pragma unroll
            for (int mt = 0; mt < 8; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    colsum[nt][0] = fma(c[mt][nt][0], c[mt][nt][0], colsum[nt][0]);
                    colsum[nt][1] = fma(c[mt][nt][1], c[mt][nt][1], colsum[nt][1]);
                }
        }
    }

    // reduce over the 8 row groups of the fragment (lanes differing in g = lane/4), then over wm
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double v = colsum[nt][e];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            colsum[nt][e] = v;
        }
    if (lane < 4) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            red[wm * TILE_T + (wn * 4 + nt) * 8 + 2 * lane + 0] = colsum[nt][0];
            red[wm * TILE_T + (wn * 4 + nt) * 8 + 2 * lane + 1] = colsum[nt][1];
        }
    }
    __syncthreads();
    if (tid < TILE_T) {
        const long t = (long)blockIdx.x * TILE_T + tid;
        if (t < k) {
            double var = __dadd_rn(kdiag, -__dadd_rn(red[tid], red[TILE_T + tid]));   // _gpr.py:480-481
            if (var < 0.0) var = 0.0;                                                 // _gpr.py:485-491
            std_out[t * ld + q] = sqrt(__dmul_rn(var, __dmul_rn(y_std, y_std)));      // _gpr.py:494,500
        }
    }
}

int launch_var(const bk_gp_s* h, int q, const double* kstar, long k, double* std_out, cudaStream_t st) {
    const GpOutput& o = h->out[q];
    if (!o.wt) return set_error("sigma requested but output %d has no packed L^-1 (w_tiles_dev was NULL)", q);
    static OncePerDevice attr_set;
    if (attr_set.first()) {
        BK_CUDA(cudaFuncSetAttribute(k_var_trmm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VAR_SMEM));
    }
    const unsigned grid = (unsigned)((k + TILE_T - 1) / TILE_T);
    ProfScope prof(PROF_VAR, st);
    k_var_trmm<<<grid, VAR_THREADS, VAR_SMEM, st>>>(o.wt, kstar, h->n_pad, o.kdiag, o.y_std, std_out, k, h->m, q);
    BK_LAUNCH_CHECK("k_var_trmm");
    return 0;
}

// ---------------------------------------------------------------- packing of W
// src is row-major with leading dimension ld; element W[r][j] is src[r*ld + j], or src[j*ld + r] when the
// source holds the transpose (WT = L^-T from the blocked inversion).  Entries outside the valid lower
// triangle are written as zeros.
__global__ void k_pack_w(const double* __restrict__ w, int n, int ld, int transposed, int n_pad,
                         double* __restrict__ out) {
    const int njb = n_pad / TILE_K;
    const size_t total = (size_t)n_pad * n_pad;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        // idx enumerates the OUTPUT so that writes are coalesced
        const int within = (int)(idx % TILE_DOUBLES);
        const size_t tile = idx / TILE_DOUBLES;
        const int jb = (int)(tile % njb), panel = (int)(tile / njb);
        const int h = within & 1, tq = (within >> 1) & 3, g = (within >> 3) & 7, k8 = (within >> 6) & 1, mt = within >> 7;
        const int r = panel * TILE_M + mt * 8 + g;
        const int j = jb * TILE_K + k8 * 8 + h * 4 + tq;
        double val = 0.0;
        if (r < n && j < n && j <= r) val = transposed ? w[(size_t)j * ld + r] : w[(size_t)r * ld + j];
        out[idx] = val;
    }
}

int launch_pack_w_ex(const double* w, int n_rows_valid, int ld, int transposed, double* out, cudaStream_t st) {
    const int n_pad = padded_n(n_rows_valid);
    ProfScope prof(PROF_OTHER, st);
    k_pack_w<<<1184, 256, 0, st>>>(w, n_rows_valid, ld, transposed, n_pad, out);
    BK_LAUNCH_CHECK("k_pack_w");
    return 0;
}

int launch_pack_w(const double* w, int n, double* out, cudaStream_t st) {
    return launch_pack_w_ex(w, n, n, 0, out, st);
}

}  // namespace bk
