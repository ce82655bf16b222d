// The steps of UQ right before and right after the Saltelli pass:
//   k_lhs_rows     the N-point randomised LHS of UQ.__init__ (ot.LHSExperiment(dist, N, True, True), uq.py:159-160):
//                  cell permutation supplied, in-cell shift from Philox, inverse CDF of the marginal;
//   k_moments      per-output sum, sum of squares (compensated), min and max of a stored output sample: the moments
//                  UQ.error_propagation reports (uq.py:501-529) without bringing the sample to the host;
//   k_sobol_asym   the row-wise products behind the asymptotic (delta-method) confidence intervals of the
//                  aggregated indices (sobol.getFirstOrderIndicesInterval(), uq.py:341, 425-426), from the stored
//                  per-block outputs of the Saltelli pass and the exact centring known after it.
// All three are HBM-bound streaming kernels: every input value is read once, coalesced.
#include "sobol_common.cuh"

namespace bk {

// ---------------------------------------------------------------- LHS rows
struct LhsParams {
    uint64_t seed;
    long count;
    int p;
    MargPar<MAX_P> marg;
    const long* perm;      // count x p: column c is a permutation of 0 .. count-1 (the cell of row i in dimension c)
    double* out;           // count x p
};
__global__ void k_lhs_rows(const LhsParams prm) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.count) return;
    const double inv_n = 1.0 / (double)prm.count;
    for (int cp = 0; 2 * cp < prm.p; ++cp) {
        uint32_t w[4];
        philox4x32_10((uint32_t)i, (uint32_t)((uint64_t)i >> 32), (uint32_t)cp, 2u, (uint32_t)prm.seed,
                      (uint32_t)(prm.seed >> 32), w);                       // stream 2: the in-cell shifts
        for (int h = 0; h < 2; ++h) {
            const int c = 2 * cp + h;
            if (c >= prm.p) break;
            const double shift = unit_from_words(w[2 * h], w[2 * h + 1]);   // in (0, 1)
            const double u = __dmul_rn(__dadd_rn((double)prm.perm[i * prm.p + c], shift), inv_n);
            prm.out[i * prm.p + c] = marginal_col<MAX_P>(prm.marg, c, u);
        }
    }
}
int launch_lhs_rows(uint64_t seed, long count, int p, const int* mkind, const double* mpar, const long* perm,
                    double* out, cudaStream_t st) {
    if (p > MAX_P) return set_error("p = %d exceeds %d", p, MAX_P);
    if (count <= 0) return 0;
    LhsParams prm;
    prm.seed = seed; prm.count = count; prm.p = p; prm.perm = perm; prm.out = out;
    marg_fill<MAX_P>(prm.marg, p, mkind, mpar);
    ProfScope prof(PROF_OTHER, st);
    k_lhs_rows<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(prm);
    BK_LAUNCH_CHECK("k_lhs_rows");
    return 0;
}

// ---------------------------------------------------------------- moments of a stored sample
// y: rows x F row-major.  CTA = (feature tile of 32, row slab); lane = feature (coalesced 256 B per row), warps stride
// the rows.  part: [slabs][F][6] = {sum hi, sum lo, sumsq hi, sumsq lo, min, max} of y - shift.
constexpr int MOM_WARPS = 8;
__global__ void __launch_bounds__(MOM_WARPS * 32) k_moments(const double* __restrict__ y, long rows, int F,
                                                            const double* __restrict__ shift, int nslabs,
                                                            double* __restrict__ part) {
    __shared__ double sh[MOM_WARPS][32][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ftiles = (F + 31) / 32;
    const int ft = blockIdx.x % ftiles, slab = blockIdx.x / ftiles;
    const int f = ft * 32 + lane;
    dd s{0.0, 0.0}, q{0.0, 0.0};
    double mn = INFINITY, mx = -INFINITY;
    if (f < F) {
        const double sv = shift ? shift[f] : 0.0;
        for (long r = (long)slab * MOM_WARPS + warp; r < rows; r += (long)nslabs * MOM_WARPS) {
            const double v = __dadd_rn(y[(size_t)r * F + f], -sv);
            dd_add_d(s, v);
            double pr, er;
            two_prod(v, v, pr, er);
            dd_add_d(q, pr);
            q.lo = __dadd_rn(q.lo, er);
            mn = fmin(mn, v);
            mx = fmax(mx, v);
        }
    }
    sh[warp][lane][0] = s.hi; sh[warp][lane][1] = s.lo; sh[warp][lane][2] = q.hi; sh[warp][lane][3] = q.lo;
    sh[warp][lane][4] = mn; sh[warp][lane][5] = mx;
    __syncthreads();
    if (warp == 0 && f < F) {
        dd ts{0.0, 0.0}, tq{0.0, 0.0};
        double tmn = INFINITY, tmx = -INFINITY;
        for (int w = 0; w < MOM_WARPS; ++w) {                                   // fixed order: reproducible
            ts = dd_add(ts, dd{sh[w][lane][0], sh[w][lane][1]});
            tq = dd_add(tq, dd{sh[w][lane][2], sh[w][lane][3]});
            tmn = fmin(tmn, sh[w][lane][4]);
            tmx = fmax(tmx, sh[w][lane][5]);
        }
        double* o = part + ((size_t)slab * F + f) * 6;
        o[0] = ts.hi; o[1] = ts.lo; o[2] = tq.hi; o[3] = tq.lo; o[4] = tmn; o[5] = tmx;
    }
}
__global__ void k_moments_reduce(const double* __restrict__ part, int nslabs, int F, const double* __restrict__ shift,
                                 double* __restrict__ out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    dd ts{0.0, 0.0}, tq{0.0, 0.0};
    double mn = INFINITY, mx = -INFINITY;
    for (int s = 0; s < nslabs; ++s) {
        const double* o = part + ((size_t)s * F + f) * 6;
        ts = dd_add(ts, dd{o[0], o[1]});
        tq = dd_add(tq, dd{o[2], o[3]});
        mn = fmin(mn, o[4]);
        mx = fmax(mx, o[5]);
    }
    const double sv = shift ? shift[f] : 0.0;
    double* o = out + (size_t)f * 6;
    o[0] = ts.hi; o[1] = ts.lo; o[2] = tq.hi; o[3] = tq.lo; o[4] = mn + sv; o[5] = mx + sv;   // min / max of y itself
}
int launch_moments(const double* y, long rows, int F, const double* shift, double* out, cudaStream_t st) {
    if (rows <= 0) return set_error("bk_moments: no rows");
    const int ftiles = (F + 31) / 32;
    long want = (rows + MOM_WARPS - 1) / MOM_WARPS;
    int nslabs = 148 * 8 / ftiles;
    if (nslabs < 1) nslabs = 1;
    if (nslabs > want) nslabs = (int)want;
    double* part = nullptr;
    BK_CUDA(cudaMallocAsync(&part, sizeof(double) * 6 * (size_t)nslabs * F, st));
    ProfScope prof(PROF_OTHER, st);
    k_moments<<<nslabs * ftiles, MOM_WARPS * 32, 0, st>>>(y, rows, F, shift, nslabs, part);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) {
        k_moments_reduce<<<(F + 127) / 128, 128, 0, st>>>(part, nslabs, F, shift, out);
        e = cudaGetLastError();
    }
    cudaFreeAsync(part, st);
    if (e != cudaSuccess) return cuda_fail(e, "k_moments");
    return 0;
}

// ---------------------------------------------------------------- products behind the asymptotic intervals
// z: [block][ld_block rows][r] stored outputs (blocks 0 = A, 1 = B, 2+i = E^i), centred here by `center` (r values: the
// mean over ALL rows of the experiment, known after the Saltelli pass).  Per base row k, with <.,.> the dot product
// over the r outputs (OpenTURNS aggregates the outputs inside the row):
//     b = <zA, zA>,   a1_i = <zB, zE^i>,   a2_i = <zA, zE^i>
// sums (2 + 6p): b, b^2, then per i: a1, a1^2, a1 b, a2, a2^2, a2 b.  A thread owns base rows k, k + stride, ...; its
// sums stay in (local-memory) doubles -- at most a few hundred terms each -- and are combined across threads, CTAs
// and calls as (hi, lo) pairs.
constexpr int ASYM_THREADS = 128;
constexpr int ASYM_MAXQ = 2 + 6 * MAX_P;
__global__ void __launch_bounds__(ASYM_THREADS) k_sobol_asym(const double* __restrict__ z, long ld_block, long count,
                                                             int p, int r, const double* __restrict__ center,
                                                             double* __restrict__ part) {
    __shared__ double sh[ASYM_THREADS / 32][ASYM_MAXQ];
    const int nq = 2 + 6 * p;
    double acc[ASYM_MAXQ];
    for (int q = 0; q < nq; ++q) acc[q] = 0.0;
    for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (long)gridDim.x * blockDim.x) {
        const double* zA = z + (size_t)k * r;
        const double* zB = z + ((size_t)ld_block + k) * r;
        double b = 0.0;
        for (int c = 0; c < r; ++c) { const double v = zA[c] - center[c]; b = fma(v, v, b); }
        acc[0] += b;
        acc[1] = fma(b, b, acc[1]);
        for (int i = 0; i < p; ++i) {
            const double* zE = z + ((size_t)(2 + i) * ld_block + k) * r;
            double a1 = 0.0, a2 = 0.0;
            for (int c = 0; c < r; ++c) {
                const double e = zE[c] - center[c];
                a1 = fma(zB[c] - center[c], e, a1);
                a2 = fma(zA[c] - center[c], e, a2);
            }
            double* o = acc + 2 + 6 * i;
            o[0] += a1; o[1] = fma(a1, a1, o[1]); o[2] = fma(a1, b, o[2]);
            o[3] += a2; o[4] = fma(a2, a2, o[4]); o[5] = fma(a2, b, o[5]);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = 0; q < nq; ++q) {
        const double v = warp_sum(acc[q]);
        if (lane == 0) sh[warp][q] = v;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < nq; q += blockDim.x) {
        dd t{0.0, 0.0};
        for (int w = 0; w < ASYM_THREADS / 32; ++w) dd_add_d(t, sh[w][q]);
        part[((size_t)blockIdx.x * nq + q) * 2] = t.hi;
        part[((size_t)blockIdx.x * nq + q) * 2 + 1] = t.lo;
    }
}
__global__ void k_sobol_asym_reduce(const double* __restrict__ part, int nctas, int nq, double* __restrict__ sums) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    dd t{sums[2 * q], sums[2 * q + 1]};                                       // ADD: chunks of base rows accumulate
    for (int c = 0; c < nctas; ++c) t = dd_add(t, dd{part[((size_t)c * nq + q) * 2], part[((size_t)c * nq + q) * 2 + 1]});
    sums[2 * q] = t.hi;
    sums[2 * q + 1] = t.lo;
}
int launch_sobol_asym(const double* z, long ld_block, long count, int p, int r, const double* center, double* sums,
                      cudaStream_t st) {
    if (p > MAX_P) return set_error("p = %d exceeds %d", p, MAX_P);
    if (count <= 0) return 0;
    const int nq = 2 + 6 * p;
    long want = (count + ASYM_THREADS - 1) / ASYM_THREADS;
    const int nctas = (int)(want < 148 * 8 ? want : 148 * 8);
    double* part = nullptr;
    BK_CUDA(cudaMallocAsync(&part, sizeof(double) * 2 * (size_t)nctas * nq, st));
    ProfScope prof(PROF_OTHER, st);
    k_sobol_asym<<<nctas, ASYM_THREADS, 0, st>>>(z, ld_block, count, p, r, center, part);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) {
        k_sobol_asym_reduce<<<(nq + 127) / 128, 128, 0, st>>>(part, nctas, nq, sums);
        e = cudaGetLastError();
    }
    cudaFreeAsync(part, st);
    if (e != cudaSuccess) return cuda_fail(e, "k_sobol_asym");
    return 0;
}

}  // namespace bk
