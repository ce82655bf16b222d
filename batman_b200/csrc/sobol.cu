// Generic Saltelli kernel (any instantiated P; runtime block selection), the reducer over stored outputs, the slot
// reduction and the estimator finalisation.  Shared device code: sobol_common.cuh; the statically scheduled
// variant for P <= 12: sobol_static.cu.
#include "sobol_common.cuh"

namespace bk {

struct DesignParams {
    uint64_t seed;
    long row_lo, count;
    int p, stream;
    MargPar<MAX_P> marg;
    double* out;
};
__global__ void k_design_rows(const DesignParams prm) {
    const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= prm.count) return;
    double x[MAX_P];
    design_row<MAX_P>(prm.seed, (uint64_t)(prm.row_lo + g), prm.p, prm.stream, prm.marg, x);
    for (int c = 0; c < prm.p; ++c) prm.out[g * prm.p + c] = x[c];
}

template <int KIND, int P>
__global__ void __launch_bounds__(SOB_THREADS) k_sobol(const SobolParams<P> prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* xs_s = reinterpret_cast<double*>(smem_raw);               // [2][TJ*P]
    double* al_s = xs_s + 2 * SOB_TJ * P;                             // [2][TJ]
    double* y_s = al_s + 2 * SOB_TJ;                                  // [nb][128]
    double* wsum = y_s + (size_t)prm.nb * SOB_THREADS;                // [4][nsums][2]
    uint64_t* bar = reinterpret_cast<uint64_t*>(wsum + 4 * prm.nsums * 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntiles = prm.n_pad / SOB_TJ;
    const int npass = (prm.nb + SOB_CH - 1) / SOB_CH;
    const long n_rowtiles = (prm.count + SOB_THREADS - 1) / SOB_THREADS;
    const long my_rowtiles = blockIdx.x < n_rowtiles ? (n_rowtiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long total_loads = my_rowtiles * npass * ntiles;
    constexpr uint32_t XS_BYTES = SOB_TJ * P * 8, AL_BYTES = SOB_TJ * 8;

    for (int i = tid; i < 4 * prm.nsums * 2; i += SOB_THREADS) wsum[i] = 0.0;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](long t) {   // t = running index of the training tile stream (cyclic over passes)
        const int s = (int)(t & 1);
        const int tile = (int)(t % ntiles);
        mbar_expect_tx(&bar[s], XS_BYTES + AL_BYTES);
        tma_load_1d(xs_s + s * SOB_TJ * P, prm.xs + (size_t)tile * SOB_TJ * P, XS_BYTES, &bar[s]);
        tma_load_1d(al_s + s * SOB_TJ, prm.alpha + (size_t)tile * SOB_TJ, AL_BYTES, &bar[s]);
    };
    if (tid == 0)
        for (long t = 0; t < 2 && t < total_loads; ++t) issue(t);

    long tctr = 0;
    for (long rt = blockIdx.x; rt < n_rowtiles; rt += gridDim.x) {
        const long g = rt * SOB_THREADS + tid;
        const bool valid = g < prm.count;
        const long gc = valid ? g : prm.count - 1;
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

        for (int pass = 0; pass < npass; ++pass) {
            const int b0 = pass * SOB_CH;
            // block bb of this pass: base row (A or B) and the swapped column (or -1)
            bool fromB[SOB_CH];
            int swp[SOB_CH];
#pragma unroll
            for (int bb = 0; bb < SOB_CH; ++bb) {
                const int b = b0 + bb;
                fromB[bb] = (b == 1) || (b >= 2 + prm.p);
                swp[bb] = b < 2 ? -1 : (b < 2 + prm.p ? b - 2 : b - 2 - prm.p);
            }
            double acc[SOB_CH], comp[SOB_CH];
#pragma unroll
            for (int bb = 0; bb < SOB_CH; ++bb) { acc[bb] = 0.0; comp[bb] = 0.0; }

            for (int tile = 0; tile < ntiles; ++tile, ++tctr) {
                const int s = (int)(tctr & 1);
                mbar_wait(&bar[s], (uint32_t)((tctr >> 1) & 1));
                const double* xrow = xs_s + s * SOB_TJ * P;
                const double* arow = al_s + s * SOB_TJ;
#pragma unroll 1
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
#pragma unroll
                    for (int bb = 0; bb < SOB_CH; ++bb) {
                        if (b0 + bb < prm.nb) {   // warp-uniform
                            double ssum = 0.0;
#pragma unroll
                            for (int c = 0; c < P; ++c)
                                ssum = __dadd_rn(ssum, ((c == swp[bb]) != fromB[bb]) ? tb[c] : ta[c]);
                            const double kv = affine(prm.c0, prm.c1, stationary<KIND>(ssum));
                            const double y = __fma_rn(kv, al, -comp[bb]);
                            const double tsum = __dadd_rn(acc[bb], y);
                            comp[bb] = __dadd_rn(__dadd_rn(tsum, -acc[bb]), -y);
                            acc[bb] = tsum;
                        }
                    }
                }
                __syncthreads();
                if (tid == 0 && tctr + 2 < total_loads) issue(tctr + 2);
            }
#pragma unroll
            for (int bb = 0; bb < SOB_CH; ++bb) {
                if (b0 + bb < prm.nb) {
                    const double y = __dadd_rn(__dmul_rn(prm.y_std, acc[bb]), prm.y_mean);
                    if (prm.yout && valid) prm.yout[((size_t)(b0 + bb) * prm.count + g) * prm.m + prm.q] = y;
                    if (prm.partial) y_s[(b0 + bb) * SOB_THREADS + tid] = valid ? __dadd_rn(y, -prm.shift) : 0.0;
                }
            }
        }
        if (prm.partial)
            saltelli_products([&](int b) { return y_s[b * SOB_THREADS + tid]; }, prm.p, prm.second, lane,
                              wsum + (size_t)warp * prm.nsums * 2, valid ? 1.0 : 0.0);
    }
    __syncthreads();
    if (prm.partial)
        flush_slot(wsum, SOB_THREADS / 32, prm.nsums, prm.partial + (size_t)blockIdx.x * prm.slot_stride);
}

template <int KIND, int P>
static int launch_sobol_kp(const bk_gp_s* h, int q, const double* a, const double* b, uint64_t seed,
                           const int* mkind, const double* mpar, long row_lo, long count, int second,
                           double shift, double* partials, double* yout, cudaStream_t st) {
    const GpOutput& o = h->out[q];
    SobolParams<P> prm;
    prm.a = a; prm.b = b; prm.seed = seed;
    marg_fill<P>(prm.marg, h->p, mkind, mpar);
    for (int c = 0; c < P; ++c) {
        prm.scale[c] = c < h->p ? h->scale[c] : 0.0;
        prm.mn[c] = c < h->p ? h->mn[c] : 0.0;
        prm.ls[c] = c < h->p ? o.ls[c] : 1.0;
    }
    prm.row_lo = row_lo; prm.count = count; prm.p = h->p; prm.second = second;
    prm.nb = sobol_nblocks(h->p, second); prm.nsums = sobol_nsums(h->p, second);
    prm.n_pad = h->n_pad; prm.xs = o.xs; prm.alpha = o.alpha;
    prm.c0 = o.c0; prm.c1 = o.c1; prm.y_mean = o.y_mean; prm.y_std = o.y_std; prm.shift = shift;
    prm.slot_stride = (long)h->m * prm.nsums * 2;
    prm.partial = partials ? partials + (size_t)q * prm.nsums * 2 : nullptr;
    prm.yout = yout; prm.m = h->m; prm.q = q;
    const size_t smem = ((size_t)2 * SOB_TJ * P + 2 * SOB_TJ + (size_t)prm.nb * SOB_THREADS + 4 * prm.nsums * 2) * 8 + 16;
    if constexpr (P <= 12) {
        const int rc = launch_sobol_static<P>(KIND, prm, smem, st);
        if (rc >= 0) return rc;
    }
    BK_CUDA(cudaFuncSetAttribute(k_sobol<KIND, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope prof(PROF_SOBOL, st);
    k_sobol<KIND, P><<<SOB_SLOTS, SOB_THREADS, smem, st>>>(prm);
    BK_LAUNCH_CHECK("k_sobol");
    return 0;
}

template <int KIND>
static int launch_sobol_k(const bk_gp_s* h, int q, const double* a, const double* b, uint64_t seed, const int* mkind,
                          const double* mpar, long row_lo, long count, int second, double shift, double* partials,
                          double* yout, cudaStream_t st) {
#define BK_CASE(PP) case PP: return launch_sobol_kp<KIND, PP>(h, q, a, b, seed, mkind, mpar, row_lo, count, second, shift, partials, yout, st);
    switch (h->P) {
        BK_CASE(2) BK_CASE(4) BK_CASE(6) BK_CASE(8) BK_CASE(10) BK_CASE(12) BK_CASE(16) BK_CASE(20) BK_CASE(24) BK_CASE(32)
    }
#undef BK_CASE
    return set_error("unsupported padded dimension %d", h->P);
}

static int launch_sobol_general(const bk_gp_s* h, int q, const double* a, const double* b, uint64_t seed,
                                const int* mkind, const double* mpar, long row_lo, long count, int second,
                                double shift, double* partials, double* yout, cudaStream_t st);

int launch_sobol(const bk_gp_s* h, int q, const double* a, const double* b, uint64_t seed, const int* mkind,
                 const double* mpar, long row_lo, long count, int second, double shift, double* partials,
                 double* yout, cudaStream_t st) {
    switch (h->out[q].kind) {
        case BK_MATERN12: return launch_sobol_k<BK_MATERN12>(h, q, a, b, seed, mkind, mpar, row_lo, count, second, shift, partials, yout, st);
        case BK_MATERN32: return launch_sobol_k<BK_MATERN32>(h, q, a, b, seed, mkind, mpar, row_lo, count, second, shift, partials, yout, st);
        case BK_MATERN52: return launch_sobol_k<BK_MATERN52>(h, q, a, b, seed, mkind, mpar, row_lo, count, second, shift, partials, yout, st);
        case BK_RBF: return launch_sobol_k<BK_RBF>(h, q, a, b, seed, mkind, mpar, row_lo, count, second, shift, partials, yout, st);
        case BK_MATERN_INF: return launch_sobol_k<BK_MATERN_INF>(h, q, a, b, seed, mkind, mpar, row_lo, count, second, shift, partials, yout, st);
        case BK_GENERAL: return launch_sobol_general(h, q, a, b, seed, mkind, mpar, row_lo, count, second, shift, partials, yout, st);
    }
    return set_error("output %d has no kernel set", q);
}

int launch_design_rows(uint64_t seed, long row_lo, long count, int p, const int* mkind, const double* mpar,
                       int stream_id, double* out, cudaStream_t st) {
    if (p > MAX_P) return set_error("p = %d exceeds %d", p, MAX_P);
    DesignParams prm;
    prm.seed = seed; prm.row_lo = row_lo; prm.count = count; prm.p = p; prm.stream = stream_id; prm.out = out;
    marg_fill<MAX_P>(prm.marg, p, mkind, mpar);
    if (count <= 0) return 0;
    k_design_rows<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(prm);
    BK_LAUNCH_CHECK("k_design_rows");
    return 0;
}

// ---------------------------------------------------------------- ensemble mode: sums from a user-supplied output sample
// (uq/uq.py:336-339).  HBM-bound: every y is read exactly once.  Warp w of the CTA walks base rows
// k = w, w+nw, ...; lane = feature (coalesced 256 B per block row); accumulators live in shared memory.
__global__ void k_sobol_y(const double* __restrict__ y, long ld_block, long count, int p, int F, int second, int nb,
                          int nsums, const double* __restrict__ shift, double* __restrict__ partials, long slot_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* acc = reinterpret_cast<double*>(smem_raw);   // [nwarps][32][nsums][2]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int ftiles = (F + 31) / 32;
    const int ft = blockIdx.x % ftiles;                  // feature tile of this CTA
    const int kslot = blockIdx.x / ftiles, nkslots = gridDim.x / ftiles;
    const int f = ft * 32 + lane;
    double* my = acc + ((size_t)(warp * 32 + lane) * nsums) * 2;
    for (int r = 0; r < 2 * nsums; ++r) my[r] = 0.0;
    if (kslot < nkslots && f < F) {
        const double sh = shift[f];
        for (long k = (long)kslot * nw + warp; k < count; k += (long)nkslots * nw) {
            auto ys = [&](int b) { return __dadd_rn(y[((size_t)b * ld_block + k) * F + f], -sh); };
            int r = 0;
            auto add = [&](double v) {
                dd a{my[2 * r], my[2 * r + 1]};
                dd_add_d(a, v);
                my[2 * r] = a.hi; my[2 * r + 1] = a.lo;
                ++r;
            };
            const double yA = ys(0), yB = ys(1);
            add(yA); add(yB); add(__dmul_rn(yA, yA)); add(__dmul_rn(yB, yB)); add(__dmul_rn(yA, yB));
            for (int i = 0; i < p; ++i) {
                const double yE = ys(2 + i);
                add(yE); add(__dmul_rn(yE, yE)); add(__dmul_rn(yB, yE)); add(__dmul_rn(yA, yE));
                const double dB = __dadd_rn(yB, -yE), dA = __dadd_rn(yA, -yE);
                add(__dmul_rn(dB, dB)); add(__dmul_rn(dA, dA));
            }
            if (second) {
                if (p > 2) {
                    for (int j = 0; j < p; ++j) { const double yC = ys(2 + p + j); add(yC); add(__dmul_rn(yC, yC)); }
                    for (int i = 0; i < p; ++i)
                        for (int j = 0; j < i; ++j) add(__dmul_rn(ys(2 + i), ys(2 + p + j)));
                }
            }
            add(1.0);                                    // row count
        }
    }
    __syncthreads();
    // combine the warps (fixed order) and ADD into slot kslot of feature f
    if (warp == 0 && f < F && kslot < nkslots) {
        double* slot = partials + (size_t)kslot * slot_stride + (size_t)f * nsums * 2;
        for (int r = 0; r < nsums; ++r) {
            dd t{slot[2 * r], slot[2 * r + 1]};
            for (int w = 0; w < nw; ++w) {
                const double* o = acc + ((size_t)(w * 32 + lane) * nsums) * 2;
                t = dd_add(t, dd{o[2 * r], o[2 * r + 1]});
            }
            slot[2 * r] = t.hi; slot[2 * r + 1] = t.lo;
        }
    }
}

// slot_stride: doubles between consecutive partial slots (0 = dense [slot][F][nsums][2])
int launch_sobol_y_ex(const double* y, long ld_block, long count, int p, int F, int second, const double* shift,
                      double* partials, long slot_stride, cudaStream_t st) {
    const int nsums = sobol_nsums(p, second), nb = sobol_nblocks(p, second);
    int nw = 4;
    while (nw > 1 && (size_t)nw * 32 * nsums * 16 > 200 * 1024) nw >>= 1;
    const size_t smem = (size_t)nw * 32 * nsums * 16;
    if (smem > 220 * 1024) return set_error("p = %d needs %zu B of shared accumulators (> 220 KiB)", p, smem);
    const int ftiles = (F + 31) / 32;
    int nkslots = SOB_SLOTS / ftiles;
    if (nkslots < 1) nkslots = 1;   // F > 32*SOB_SLOTS would need more slots than the partial buffer has
    if ((long)ftiles > SOB_SLOTS) return set_error("F = %d too large for %d partial slots", F, SOB_SLOTS);
    BK_CUDA(cudaFuncSetAttribute(k_sobol_y, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope prof(PROF_SOBOL_Y, st);
    if (slot_stride <= 0) slot_stride = (long)F * nsums * 2;
    k_sobol_y<<<nkslots * ftiles, nw * 32, smem, st>>>(y, ld_block, count, p, F, second, nb, nsums, shift, partials, slot_stride);
    BK_LAUNCH_CHECK("k_sobol_y");
    return 0;
}

int launch_sobol_y(const double* y, long ld_block, long count, int p, int F, int second, const double* shift,
                   double* partials, cudaStream_t st) {
    return launch_sobol_y_ex(y, ld_block, count, p, F, second, shift, partials, 0, st);
}

// ---------------------------------------------------------------- general kernel trees: unfused composition
// Outputs with kind == BK_GENERAL have no fused Saltelli kernel: block b of the design is materialised for a bounded
// slab of base rows (A / B read or drawn with Philox, one column swapped), predicted with k_predict_gen (mean only)
// and, unless the caller wants the outputs themselves, reduced with the reducer over stored outputs (k_sobol_y).
// Same sums, same slots; scratch comes from the stream-ordered allocator and is private to the call.
struct BlockParams {
    const double* a;
    const double* b;
    uint64_t seed;
    long row_lo, count;
    int p, block;          // block: 0 = A, 1 = B, 2+i = E^i, 2+p+i = C^i
    MargPar<MAX_P> marg;
    double* out;           // count x p
};
__global__ void k_design_block(const BlockParams prm) {
    const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= prm.count) return;
    const int b = prm.block, p = prm.p;
    const bool fromB = (b == 1) || (b >= 2 + p);
    const int swp = b < 2 ? -1 : (b < 2 + p ? b - 2 : b - 2 - p);
    double xa[MAX_P], xb[MAX_P];
    if (prm.a) {
        for (int c = 0; c < p; ++c) { xa[c] = prm.a[g * p + c]; xb[c] = prm.b[g * p + c]; }
    } else {
        design_row<MAX_P>(prm.seed, (uint64_t)(prm.row_lo + g), p, 0, prm.marg, xa);
        design_row<MAX_P>(prm.seed, (uint64_t)(prm.row_lo + g), p, 1, prm.marg, xb);
    }
    for (int c = 0; c < p; ++c) prm.out[g * p + c] = ((c == swp) != fromB) ? xb[c] : xa[c];
}

int launch_predict_general(const bk_gp_s* h, int q, const double* pts, long k, double* mean, int ld, int col,
                           double* kstar, int8_t* kq, int kq_scheme, double kappa, cudaStream_t st);   // gp_predict_gen.cu

static int launch_sobol_general(const bk_gp_s* h, int q, const double* a, const double* b, uint64_t seed,
                                const int* mkind, const double* mpar, long row_lo, long count, int second,
                                double shift, double* partials, double* yout, cudaStream_t st) {
    const int p = h->p, nb = sobol_nblocks(p, second), nsums = sobol_nsums(p, second);
    const long SLAB = 1L << 18;                       // base rows per slab: scratch <= 2^18 * (p + nb + 1) doubles
    const long slab_max = count < SLAB ? count : SLAB;
    struct Scratch {                                   // stream-ordered scratch, released on every exit path
        cudaStream_t st;
        double* design = nullptr;
        double* ybuf = nullptr;
        double* shift_dev = nullptr;
        ~Scratch() {
            if (design) cudaFreeAsync(design, st);
            if (ybuf) cudaFreeAsync(ybuf, st);
            if (shift_dev) cudaFreeAsync(shift_dev, st);
        }
    } sc{st};
    BK_CUDA(cudaMallocAsync(&sc.design, sizeof(double) * slab_max * p, st));
    if (partials) {
        BK_CUDA(cudaMallocAsync(&sc.ybuf, sizeof(double) * slab_max * nb, st));
        BK_CUDA(cudaMallocAsync(&sc.shift_dev, sizeof(double), st));
        BK_CUDA(cudaMemcpyAsync(sc.shift_dev, &shift, sizeof(double), cudaMemcpyHostToDevice, st));
        BK_CUDA(cudaStreamSynchronize(st));            // `shift` is a stack value
    }
    double* design = sc.design;
    double* ybuf = sc.ybuf;
    double* shift_dev = sc.shift_dev;
    BlockParams bp;
    bp.seed = seed; bp.p = p;
    marg_fill<MAX_P>(bp.marg, p, mkind, mpar);
    int rc = 0;
    for (long lo = 0; lo < count && rc == 0; lo += SLAB) {
        const long cnt = (count - lo) < SLAB ? (count - lo) : SLAB;
        bp.a = a ? a + lo * p : nullptr;
        bp.b = b ? b + lo * p : nullptr;
        bp.row_lo = row_lo + lo; bp.count = cnt; bp.out = design;
        for (int blk = 0; blk < nb && rc == 0; ++blk) {
            bp.block = blk;
            k_design_block<<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(bp);
            if (cudaGetLastError() != cudaSuccess) { rc = set_error("k_design_block launch failed"); break; }
            if (!partials) {     // yout[(blk * count + lo + k) * m + q]
                rc = launch_predict_general(h, q, design, cnt, yout + ((size_t)blk * count + lo) * h->m, h->m, q, nullptr,
                                            nullptr, 0, 1.0, st);
            } else {             // ybuf[blk * cnt + k] (+ a strided copy when the caller keeps the outputs as well)
                rc = launch_predict_general(h, q, design, cnt, ybuf + (size_t)blk * cnt, 1, 0, nullptr, nullptr, 0, 1.0, st);
                if (rc == 0 && yout)
                    BK_CUDA(cudaMemcpy2DAsync(yout + ((size_t)blk * count + lo) * h->m + q, sizeof(double) * h->m,
                                              ybuf + (size_t)blk * cnt, sizeof(double), sizeof(double), cnt,
                                              cudaMemcpyDeviceToDevice, st));
            }
        }
        if (rc == 0 && partials)
            rc = launch_sobol_y_ex(ybuf, cnt, cnt, p, 1, second, shift_dev, partials + (size_t)q * nsums * 2,
                                   (long)h->m * nsums * 2, st);
    }
    return rc;
}

// ---------------------------------------------------------------- slots -> totals -> indices
__global__ void k_reduce_slots(const double* __restrict__ partials, int nslots, long per_slot, double* __restrict__ totals) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;   // (feature, sum) pair
    if (i >= per_slot) return;
    dd t{0.0, 0.0};
    for (int s = 0; s < nslots; ++s) t = dd_add(t, dd{partials[((size_t)s * per_slot + i) * 2], partials[((size_t)s * per_slot + i) * 2 + 1]});
    totals[2 * i] = t.hi;
    totals[2 * i + 1] = t.lo;
}

int launch_reduce_slots(const double* partials, int F, int nsums, double* totals, cudaStream_t st) {
    const long per_slot = (long)F * nsums;
    ProfScope prof(PROF_OTHER, st);
    k_reduce_slots<<<(unsigned)((per_slot + 127) / 128), 128, 0, st>>>(partials, SOB_SLOTS, per_slot, totals);
    BK_LAUNCH_CHECK("k_reduce_slots");
    return 0;
}

// Estimators (conventions frozen in oracle/saltelli.py; OT 1.15 SaltelliSensitivityAlgorithm::computeIndices,
// uq/uq.py:400-436 for the aggregation).  With y' = y - shift the centred output is y' - delta, delta = mean
// over ALL R rows of y'; every centred product follows algebraically, evaluated in double-double.
struct FinalizeLayout {
    size_t S1, ST, S2, V1, VT, var, outvar, J1, JT, S1agg, STagg, S2agg;
    size_t M1, MT, K1, KT;                                  // Martinez, Mauntz-Kucherenko (per feature)
    size_t J1agg, JTagg, M1agg, MTagg, K1agg, KTagg, mean;  // aggregated variants; mean of every output over all rows
    size_t total;
};
__host__ __device__ inline FinalizeLayout finalize_layout(int p, int F) {
    FinalizeLayout L;
    size_t o = 0;
    L.S1 = o; o += (size_t)F * p;
    L.ST = o; o += (size_t)F * p;
    L.S2 = o; o += (size_t)F * p * p;
    L.V1 = o; o += (size_t)F * p;
    L.VT = o; o += (size_t)F * p;
    L.var = o; o += F;
    L.outvar = o; o += F;
    L.J1 = o; o += (size_t)F * p;
    L.JT = o; o += (size_t)F * p;
    L.S1agg = o; o += p;
    L.STagg = o; o += p;
    L.S2agg = o; o += (size_t)p * p;
    L.M1 = o; o += (size_t)F * p;
    L.MT = o; o += (size_t)F * p;
    L.K1 = o; o += (size_t)F * p;
    L.KT = o; o += (size_t)F * p;
    L.J1agg = o; o += p;
    L.JTagg = o; o += p;
    L.M1agg = o; o += p;
    L.MTagg = o; o += p;
    L.K1agg = o; o += p;
    L.KTagg = o; o += p;
    L.mean = o; o += F;
    L.total = o;
    return L;
}

__global__ void k_finalize_feature(const double* __restrict__ totals, long N, int p, int F, int second, int nsums,
                                   double* __restrict__ out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const FinalizeLayout L = finalize_layout(p, F);
    const double* T = totals + (size_t)f * nsums * 2;
    auto S = [&](int r) { return dd{T[2 * r], T[2 * r + 1]}; };
    const int nb = sobol_nblocks(p, second);
    const dd dN = dd_from((double)N), dN1 = dd_from((double)(N - 1)), dR = dd_from((double)N * nb);
    const int base2 = 5 + 6 * p;
    // delta = mean of y' over all rows; sum of squares over all rows
    dd sum_all = dd_add(S(0), S(1)), sq_all = dd_add(S(2), S(3));
    for (int i = 0; i < p; ++i) { sum_all = dd_add(sum_all, S(5 + 6 * i)); sq_all = dd_add(sq_all, S(5 + 6 * i + 1)); }
    if (second && p > 2)
        for (int j = 0; j < p; ++j) { sum_all = dd_add(sum_all, S(base2 + 2 * j)); sq_all = dd_add(sq_all, S(base2 + 2 * j + 1)); }
    const dd delta = dd_div(sum_all, dR);
    const dd nd2 = dd_mul(dN, dd_mul(delta, delta));
    // centred dot product of two blocks: sum(xy) - delta (sum x + sum y) + N delta^2
    auto cdot = [&](dd sxy, dd sx, dd sy) { return dd_add(dd_sub(sxy, dd_mul(delta, dd_add(sx, sy))), nd2); };
    const dd SA = S(0), SB = S(1);
    const dd var = dd_div(dd_sub(S(2), dd_div(dd_mul(SA, SA), dN)), dN1);           // var(yA, ddof=1)
    const dd f02 = dd_div(cdot(S(4), SA, SB), dN);                                  // sum(yA yB)/N
    const dd muA = dd_sub(dd_div(SA, dN), delta);
    const dd outvar = dd_div(dd_sub(sq_all, dd_mul(dR, dd_mul(delta, delta))), dR); // output_design.var(axis=0)
    out[L.var + f] = dd_to_double(var);
    out[L.outvar + f] = dd_to_double(outvar);
    out[L.mean + f] = dd_to_double(delta);                 // mean over all rows of y - shift (the host adds the shift)
    const dd d2N1 = dd_from((double)(2 * N - 1));
    // other estimators from the same sums (the comment of uq.py:340 names them; conventions: oracle/saltelli.py).
    // Martinez: S_i = corr(yB, yE^i), ST_i = 1 - corr(yA, yE^i) (shift invariant: no centring needed);
    // Mauntz-Kucherenko: V_i = sum yB (yE^i - yA) / (N - 1), VT_i = sum yA (yA - yE^i) / (N - 1) on the centred output.
    const dd ssA = dd_sub(S(2), dd_div(dd_mul(SA, SA), dN));                         // sum (yA - mean yA)^2
    const dd ssB = dd_sub(S(3), dd_div(dd_mul(SB, SB), dN));
    const dd cAA = cdot(S(2), SA, SA), cAB = cdot(S(4), SA, SB);
    for (int i = 0; i < p; ++i) {
        const int b = 5 + 6 * i;
        const dd SE = S(b);
        const dd V1 = dd_sub(dd_div(cdot(S(b + 2), SB, SE), dN1), f02);
        const dd VT = dd_add(dd_sub(var, dd_div(cdot(S(b + 3), SA, SE), dN1)), dd_mul(muA, muA));
        out[L.V1 + (size_t)f * p + i] = dd_to_double(V1);
        out[L.VT + (size_t)f * p + i] = dd_to_double(VT);
        out[L.S1 + (size_t)f * p + i] = dd_to_double(dd_div(V1, var));
        out[L.ST + (size_t)f * p + i] = dd_to_double(dd_div(VT, var));
        out[L.J1 + (size_t)f * p + i] = dd_to_double(dd_div(dd_sub(var, dd_div(S(b + 4), d2N1)), var));
        out[L.JT + (size_t)f * p + i] = dd_to_double(dd_div(dd_div(S(b + 5), d2N1), var));
        const dd ssE = dd_sub(S(b + 1), dd_div(dd_mul(SE, SE), dN));
        const double covBE = dd_to_double(dd_sub(S(b + 2), dd_div(dd_mul(SB, SE), dN)));
        const double covAE = dd_to_double(dd_sub(S(b + 3), dd_div(dd_mul(SA, SE), dN)));
        out[L.M1 + (size_t)f * p + i] = covBE / sqrt(dd_to_double(ssB) * dd_to_double(ssE));
        out[L.MT + (size_t)f * p + i] = 1.0 - covAE / sqrt(dd_to_double(ssA) * dd_to_double(ssE));
        const dd cBE = cdot(S(b + 2), SB, SE), cAE = cdot(S(b + 3), SA, SE);
        out[L.K1 + (size_t)f * p + i] = dd_to_double(dd_div(dd_div(dd_sub(cBE, cAB), dN1), var));
        out[L.KT + (size_t)f * p + i] = dd_to_double(dd_div(dd_div(dd_sub(cAA, cAE), dN1), var));
    }
    for (int i = 0; i < p; ++i)
        for (int j = 0; j < p; ++j) out[L.S2 + ((size_t)f * p + i) * p + j] = 0.0;
    if (second) {
        for (int i = 0; i < p; ++i)
            for (int j = 0; j < i; ++j) {
                dd sEC, SC;
                if (p > 2) { sEC = S(base2 + 2 * p + i * (i - 1) / 2 + j); SC = S(base2 + 2 * j); }
                else { sEC = S(5 + 6 * i + 1); SC = S(5 + 6 * i); }   // p == 2: C^j == E^(1-j) == E^i
                const dd SE = S(5 + 6 * i);
                const dd s1i = dd_div(dd_sub(dd_div(cdot(S(5 + 6 * i + 2), SB, SE), dN1), f02), var);
                const dd SEj = S(5 + 6 * j);
                const dd s1j = dd_div(dd_sub(dd_div(cdot(S(5 + 6 * j + 2), SB, SEj), dN1), f02), var);
                const dd s2 = dd_sub(dd_sub(dd_div(dd_sub(dd_div(cdot(sEC, SE, SC), dN1), f02), var), s1i), s1j);
                out[L.S2 + ((size_t)f * p + i) * p + j] = dd_to_double(s2);
                out[L.S2 + ((size_t)f * p + j) * p + i] = dd_to_double(s2);
            }
    }
}

// aggregated indices (uq.py:412-426): S1agg = sum_q V1 / sum_q var; S2agg = sum_q outvar_q nan_to_num(S2_q) / sum outvar
__global__ void k_finalize_aggregate(int p, int F, double* __restrict__ out) {
    const FinalizeLayout L = finalize_layout(p, F);
    const int t = threadIdx.x;
    if (t < p) {
        dd sv{0, 0}, s1{0, 0}, st{0, 0};
        for (int f = 0; f < F; ++f) {
            dd_add_d(sv, out[L.var + f]);
            dd_add_d(s1, out[L.V1 + (size_t)f * p + t]);
            dd_add_d(st, out[L.VT + (size_t)f * p + t]);
        }
        out[L.S1agg + t] = dd_to_double(dd_div(s1, sv));
        out[L.STagg + t] = dd_to_double(dd_div(st, sv));
        // every estimator aggregates the same way: sum_q (index_q * Var_q) / sum_q Var_q
        const size_t src[6] = {L.J1, L.JT, L.M1, L.MT, L.K1, L.KT};
        const size_t dst[6] = {L.J1agg, L.JTagg, L.M1agg, L.MTagg, L.K1agg, L.KTagg};
        for (int e = 0; e < 6; ++e) {
            dd acc{0, 0};
            for (int f = 0; f < F; ++f) {
                double p_, e_;
                two_prod(out[src[e] + (size_t)f * p + t], out[L.var + f], p_, e_);
                acc = dd_add(acc, dd{p_, e_});
            }
            out[dst[e] + t] = dd_to_double(dd_div(acc, sv));
        }
    }
    for (int e = t; e < p * p; e += blockDim.x) {
        dd so{0, 0}, s2{0, 0};
        for (int f = 0; f < F; ++f) {
            double v = out[L.S2 + (size_t)f * p * p + e];
            if (isnan(v)) v = 0.0;                                   // np.nan_to_num
            else if (isinf(v)) v = v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
            const double ov = out[L.outvar + f];
            dd_add_d(so, ov);
            dd_add_d(s2, __dmul_rn(ov, v));
        }
        out[L.S2agg + e] = dd_to_double(dd_div(s2, so));
    }
}

// ---------------------------------------------------------------- delete-a-group jackknife of the aggregated indices
// The partial slots are disjoint groups of base rows, so "the estimator without group g" is the estimator of
// (totals - slot g) with N - n_g rows, n_g = the slot's row count (last sum).  One CTA per slot: threads stride the
// features, form Var, V_i, VT_i of every feature from the leave-one-group-out sums (same algebra as
// k_finalize_feature, double-double) and reduce them to the aggregated S_i = sum_f V_i / sum_f Var (uq.py:423-424).
// out[g] = {active, n_g, S1agg(-g)[p], STagg(-g)[p]}; the host turns these into the jackknife variance.
__global__ void __launch_bounds__(128) k_sobol_jackknife(const double* __restrict__ partials, const double* __restrict__ totals,
                                                         long N, int p, int F, int second, int nsums,
                                                         double* __restrict__ out) {
    __shared__ double red[4][1 + 2 * MAX_P];
    const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int width = 2 + 2 * p;
    const double* P0 = partials + (size_t)g * F * nsums * 2;
    const double n_g = P0[2 * (nsums - 1)] + P0[2 * (nsums - 1) + 1];
    const double Nl = (double)N - n_g;                                   // rows left
    if (!(n_g > 0.0) || Nl < 2.0) {                                      // empty slot (or nothing left): not a group
        if (tid == 0) { out[(size_t)g * width] = 0.0; out[(size_t)g * width + 1] = n_g; }
        return;
    }
    const int nb = sobol_nblocks(p, second);
    const dd dN = dd_from(Nl), dN1 = dd_from(Nl - 1.0), dR = dd_from(Nl * nb);
    const int base2 = 5 + 6 * p;
    double acc[1 + 2 * MAX_P];
    for (int i = 0; i < 1 + 2 * p; ++i) acc[i] = 0.0;
    for (int f = tid; f < F; f += blockDim.x) {
        const double* T = totals + (size_t)f * nsums * 2;
        const double* Pg = P0 + (size_t)f * nsums * 2;
        auto S = [&](int r) { return dd_sub(dd{T[2 * r], T[2 * r + 1]}, dd{Pg[2 * r], Pg[2 * r + 1]}); };
        dd sum_all = dd_add(S(0), S(1));
        for (int i = 0; i < p; ++i) sum_all = dd_add(sum_all, S(5 + 6 * i));
        if (second && p > 2)
            for (int j = 0; j < p; ++j) sum_all = dd_add(sum_all, S(base2 + 2 * j));
        const dd delta = dd_div(sum_all, dR);
        const dd nd2 = dd_mul(dN, dd_mul(delta, delta));
        auto cdot = [&](dd sxy, dd sx, dd sy) { return dd_add(dd_sub(sxy, dd_mul(delta, dd_add(sx, sy))), nd2); };
        const dd SA = S(0), SB = S(1);
        const dd var = dd_div(dd_sub(S(2), dd_div(dd_mul(SA, SA), dN)), dN1);
        const dd f02 = dd_div(cdot(S(4), SA, SB), dN);
        const dd muA = dd_sub(dd_div(SA, dN), delta);
        acc[0] += dd_to_double(var);
        for (int i = 0; i < p; ++i) {
            const int b = 5 + 6 * i;
            const dd SE = S(b);
            const dd V1 = dd_sub(dd_div(cdot(S(b + 2), SB, SE), dN1), f02);
            const dd VT = dd_add(dd_sub(var, dd_div(cdot(S(b + 3), SA, SE), dN1)), dd_mul(muA, muA));
            acc[1 + i] += dd_to_double(V1);
            acc[1 + p + i] += dd_to_double(VT);
        }
    }
    for (int i = 0; i < 1 + 2 * p; ++i) {
        double v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (tid < 2 * p) {
        const double sv = red[0][0] + red[1][0] + red[2][0] + red[3][0];
        const double sn = red[0][1 + tid] + red[1][1 + tid] + red[2][1 + tid] + red[3][1 + tid];
        out[(size_t)g * width + 2 + tid] = sn / sv;
    }
    if (tid == 0) { out[(size_t)g * width] = 1.0; out[(size_t)g * width + 1] = n_g; }
}

int launch_sobol_jackknife(const double* partials, const double* totals, long N, int p, int F, int second, double* out,
                           cudaStream_t st) {
    if (p > MAX_P) return set_error("p = %d exceeds %d", p, MAX_P);
    const int nsums = sobol_nsums(p, second);
    ProfScope prof(PROF_OTHER, st);
    k_sobol_jackknife<<<SOB_SLOTS, 128, 0, st>>>(partials, totals, N, p, F, second, nsums, out);
    BK_LAUNCH_CHECK("k_sobol_jackknife");
    return 0;
}

size_t sobol_result_len(int p, int F) { return finalize_layout(p, F).total; }

int launch_finalize(const double* totals, long N, int p, int F, int second, double* out, cudaStream_t st) {
    const int nsums = sobol_nsums(p, second);
    ProfScope prof(PROF_OTHER, st);
    k_finalize_feature<<<(F + 63) / 64, 64, 0, st>>>(totals, N, p, F, second, nsums, out);
    BK_LAUNCH_CHECK("k_finalize_feature");
    k_finalize_aggregate<<<1, 128, 0, st>>>(p, F, out);
    BK_LAUNCH_CHECK("k_finalize_aggregate");
    return 0;
}

}  // namespace bk

namespace bk {
int sobol_nsums_host(int p, int second) { return sobol_nsums(p, second); }
int sobol_nslots_host() { return SOB_SLOTS; }
}  // namespace bk
