// K2' `k_var_ozaki`: the variance TRMM on the 5th-generation tensor cores (tcgen05, int8, TMEM accumulators).
//
// tcgen05 has no FP64 kind, so `V = W K*` is evaluated EXACTLY from int8 slices (Ozaki scheme I):
//     W_ij  = sigma_i * sum_{a=1..8} q_a(i,j) 2^-(7a-1)      sigma_i = power of two >= max_j |W_ij|
//     K*_jt = kappa   * sum_{b=1..8} p_b(j,t) 2^-(7b-1)      kappa   = power of two >= max |k|,   |q|,|p| <= 64
//     V_it  = sigma_i kappa * sum_{g=2..9} 2^-(7g-2) D_g(i,t),   D_g = sum_{a+b=g} sum_j q_a p_b
// D_g is an int32 accumulation of int8 products (exact for n <= 16384); the 36 slice pairs with a+b <= 9
// reproduce float64 accuracy (profiles/r01_ozaki_feasibility.md).  Same contract as k_var_trmm (gp_var.cu).
//
// A GANG of G CTAs = 128 test points (TMEM lanes, M side = K* slices) against 64-row panels of W (N side).  The panels
// are dealt to the G CTAs in snake order (CTA c takes the panels c and 2G-1-c of every run of 2G): the same number
// of k slabs each, because W is lower triangular and panel P runs 2 (P + 1) slabs.  The CTAs of a gang work at the
// same time and re-read the same K* tile (n_pad * 896 bytes: 0.9 MB at n = 1024, 3.7 MB at n = 4096) once per
// panel, so only (SM count / G) tiles are live at a time; G grows with n_train so that they fit in the L2 and the
// re-reads never reach DRAM.  The CTA that finishes last adds the G partial sums of v^2 in a fixed order.  ONE pass over
// K per panel: all accumulator groups D_2..D_9 of the 128 x 64 tile are resident in TMEM at once, group g in columns
// [64 (g-2), 64 (g-1)) -- 448 or 512 of the 512 columns.  The slices of a W panel are stacked along N in shared
// memory ([slice][64 rows][32 B], 8-row core-matrix groups 256 B apart), so ONE tcgen05.mma of K* slice b against
// W slices a..a+3 (N = 256) accumulates into the four adjacent groups a+b .. a+b+3: the instruction shape that
// reaches the int8 peak (N = 256; an M128 instruction costs >= 94 cycles whatever N) without the 256 TMEM columns
// per group a conventional N = 256 tile needs.  Per 32-wide k slab: 10 MMAs for the 28 pairs of the 7 x 8-bit
// scheme (4 x N256, 2 x N192, 2 x N128, 2 x N64), 12 for the 36 pairs of the 8 x 7-bit scheme.
//   warp 5 lane 0 : TMA producer  (cp.async.bulk of pre-packed slice tiles, 4-slot full/empty mbarrier ring;
//                                  a slot = one k slab = the K* slices (S x 4 KiB) and the W slices (S x 2 KiB))
//   warp 4 lane 0 : MMA issuer    (tcgen05.mma.cta_group::1.kind::i8, M128 K32, SMEM descriptors, commit)
//   warps 0..3, 6..9 : epilogue   (tcgen05.ld -> exact int64 combine -> FP64 -> sum of v^2; two threads per test
//                                  point, each draining 32 of the panel's 64 W rows from every group,
//                                  software-pipelined TMEM reads, no cross-thread reduction)
// Operand tiles are packed on the producer side (W at upload, K* by K1) in the non-swizzled K-major UMMA
// layout: element (row r, byte b) of an R x 32 B tile at (r/8)*256 + (b/16)*128 + (r%8)*16 + b%16.
#include <stdlib.h>

#include "common.cuh"
#include "model.h"

namespace bk {

constexpr int OZ_S = 8;                       // slice stride of both tile arrays (the 7 x 8-bit scheme leaves slot 8 unused)
constexpr int OZ_KTILE = 4096;                // bytes of one K* slice tile: 128 test points x 32 B
constexpr int OZ_PANEL = 64;                  // W rows per panel
constexpr int OZ_WTILE = OZ_PANEL * 32;       // bytes of one W slice tile: 64 rows x 32 B
constexpr int OZ_SLOTS = 4;                   // ring of 48 KiB slots: [K* slices 1..8 | W slices 1..8] of one k slab
constexpr int OZ_SLOT_W = OZ_S * OZ_KTILE;    // offset of the W slices inside a slot
constexpr int OZ_SLOT_BYTES = OZ_S * (OZ_KTILE + OZ_WTILE);
constexpr int OZ_THREADS = 320;               // warps 0-3 and 6-9: epilogue (two column halves); warp 4: MMA; warp 5: TMA
constexpr size_t OZ_SMEM = (size_t)OZ_SLOTS * OZ_SLOT_BYTES + 2 * OZ_PANEL * 8 + 2 * 128 * 8 + (2 * OZ_SLOTS + 2) * 8 + 1024;

__device__ __forceinline__ uint64_t oz_desc(uint32_t smem_addr) {
    // SMEM matrix descriptor, no swizzle, K-major: LBO = 128 B (core matrices along K), SBO = 256 B (8-row groups)
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(128 >> 4) << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;                   // descriptor version for sm_100
    return d;
}
__device__ __forceinline__ void oz_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void oz_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bulk copy with an L2 eviction-priority hint.  The K* tiles of the gangs in flight are the re-read set (once per W
// panel): evict-last.  W is re-read once per tile by every gang: normal priority while it fits beside the K* tiles,
// evict-first when it is far larger than the L2 (n_train > ~5000: it streams, and must not push the K* tiles out).
__device__ __forceinline__ void tma_load_1d_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
// bounded wait: a mis-programmed pipeline must not hang the GPU box.  A time-out (or a fault recorded by another
// role in `fail_s`) makes every later wait of the CTA return at once; callers skip their work but keep walking
// through their named barriers, so nobody is left blocked.
__device__ __forceinline__ bool oz_wait(uint64_t* bar, uint32_t parity, int* fail, volatile int* fail_s) {
    if (*fail) return false;
    for (long spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1L << 24) || ((spin & 1023) == 1023 && *fail_s)) { *fail = 1; *fail_s = 1; return false; }
    return true;
}

// Two error-free slicings (SCHEME):
//   0: 8 x 7-bit signed digits for both operands; group g = a + b has weight 2^-(7g-2); 36 pairs with g <= 9;
//      int32 exact for n <= 16384.
//   1: 7 x 8-bit digits, W signed (|w| < 1/2, X = rint(w 2^55)), K* unsigned bytes (0 <= x < 1, X = rint(x 2^56));
//      group g has weight 2^(1-8g); 28 pairs with g <= 8; int32 exact for n <= 8192; needs k >= 0 (c0, c1 >= 0).
//      22 % fewer MMAs, byte-extraction slicing in K1.
template <int SCHEME> struct OzTraits;
template <> struct OzTraits<0> { static constexpr int S = 8, BITS = 7, G_END = 10; };   // groups 2 .. G_END-1
template <> struct OzTraits<1> { static constexpr int S = 7, BITS = 8, G_END = 9; };

// instruction descriptor: S32 accumulate; A = K* slices (signed, or unsigned bytes in scheme 1), B = W slices
// (signed); K-major both; M = 128, N = 64 * (stacked W slices)
template <int SCHEME>
__device__ __forceinline__ constexpr uint32_t oz_idesc(int n) {
    return (2u << 4) | ((SCHEME == 0 ? 1u : 0u) << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

// All slice pairs of one k slab, fully unrolled: K* slice b against the stacked W slices a .. a+3 lands in the
// adjacent accumulator groups a+b .. a+b+3.  Per MMA the issuing thread only adds a tile offset to two pre-built
// descriptors (start-address field, 16-B units).  The two b = 1 instructions touch every group first: they
// overwrite on the first slab of a panel (`not_first` = 0), everything else accumulates.
template <int SCHEME>
__device__ __forceinline__ void oz_issue_slab(uint32_t tmem_base, uint64_t kdesc0, uint64_t wdesc0, uint32_t not_first) {
    constexpr int S = OzTraits<SCHEME>::S, G_END = OzTraits<SCHEME>::G_END;
#pragma unroll
    for (int b = 1; b <= S; ++b) {
        const int a_max = G_END - 1 - b < S ? G_END - 1 - b : S;
#pragma unroll
        for (int a = 1; a <= a_max; a += 4) {
            const int cnt = a_max - a + 1 < 4 ? a_max - a + 1 : 4;
            oz_mma(tmem_base + (uint32_t)((a + b - 2) * OZ_PANEL), kdesc0 + (uint64_t)((b - 1) * (OZ_KTILE >> 4)),
                   wdesc0 + (uint64_t)((a - 1) * (OZ_WTILE >> 4)), oz_idesc<SCHEME>(cnt * OZ_PANEL), b == 1 ? not_first : 1u);
        }
    }
}

__device__ __forceinline__ double oz_i64_to_double(long long h) {    // exact for |h| < 2^51; no I2F.F64
    return __longlong_as_double(h + 0x4338000000000000ll) - 6755399441055744.0;
}

template <int NG>
__device__ __forceinline__ void oz_load_chunk(uint32_t (&d)[NG][8], uint32_t taddr) {
#pragma unroll
    for (int g = 0; g < NG; ++g)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(d[g][0]), "=r"(d[g][1]), "=r"(d[g][2]), "=r"(d[g][3]), "=r"(d[g][4]), "=r"(d[g][5]),
                       "=r"(d[g][6]), "=r"(d[g][7])
                     : "r"(taddr + (uint32_t)(g * OZ_PANEL)));
}

template <int SCHEME>
__global__ void __launch_bounds__(OZ_THREADS, 1)
k_var_ozaki(const int8_t* __restrict__ wq, const double* __restrict__ wsig, const int8_t* __restrict__ kq,
            int n_pad, double kappa, double kdiag, double y_std, double* __restrict__ std_out, long k, int ld, int q,
            double* __restrict__ partial /* [tiles][G][128] */, int* __restrict__ arrived /* [tiles], zeroed */,
            int G, int w_streams, int* __restrict__ fail_flag) {
    constexpr int S = OzTraits<SCHEME>::S;
    constexpr int NG = OzTraits<SCHEME>::G_END - 2;                  // accumulator groups (8 or 7)
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* stages = smem;
    double* sigs = reinterpret_cast<double*>(smem + (size_t)OZ_SLOTS * OZ_SLOT_BYTES);        // [2][64] sigma_i * kappa
    double* colred = sigs + 2 * OZ_PANEL;                                                     // [2][128]
    uint64_t* full = reinterpret_cast<uint64_t*>(colred + 2 * 128);                           // [slots]
    uint64_t* empty = full + OZ_SLOTS;                                                        // [slots]
    uint64_t* acc_full = empty + OZ_SLOTS;
    uint64_t* acc_empty = acc_full + 1;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int fail_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nks = n_pad / 32;
    // Persistent CTAs: the grid is a whole number of gangs (<= the SM count); CTA b = (gang b / G, member b % G) walks
    // the tiles gang, gang + gangs, ... and always takes the same share of their panels, so the members of a gang
    // stay on the same tile.  The mbarrier ring, the accumulator hand-shake and TMEM live across tiles: the producer
    // is already filling the first slabs of the next tile while the epilogue drains the last panel of this one.
    const int member = blockIdx.x % G, tile0 = blockIdx.x / G, tile_step = gridDim.x / G;
    const int n_tiles = (int)((k + 127) / 128);
    // panels of this CTA in snake order: i-th one = 2G (i / 2) + (member | 2G - 1 - member), increasing in i
    auto my_panel = [&](int i) { return 2 * G * (i >> 1) + ((i & 1) ? 2 * G - 1 - member : member); };
    int n_mine = 0;
    while (my_panel(n_mine) < n_pad / OZ_PANEL) ++n_mine;
    __shared__ int last_s;

    if (tid == 0) {
        for (int s = 0; s < OZ_SLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 8);
        mbar_fence_init();
        fail_s = 0;
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 5) {
        // ================= TMA producer: W is lower triangular, panel P needs the k slabs below 2 (P + 1) =================
        if (lane == 0) {
            int fail = 0;
            long u = 0;                                                      // running slot-use counter
            uint64_t pol_keep, pol_stream;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
            for (int tile = tile0; tile < n_tiles && !fail; tile += tile_step) {
              const int8_t* kt = kq + (size_t)tile * nks * OZ_S * OZ_KTILE;
              for (int pi = 0; pi < n_mine && !fail; ++pi) {
                const int panel = my_panel(pi);
                for (int ks = 0; ks < (panel + 1) * 2; ++ks, ++u) {
                    const int s = (int)(u % OZ_SLOTS);
                    if (u >= OZ_SLOTS && !oz_wait(&empty[s], (uint32_t)(((u / OZ_SLOTS) - 1) & 1), &fail, &fail_s)) break;
                    unsigned char* dst = stages + (size_t)s * OZ_SLOT_BYTES;
                    mbar_expect_tx(&full[s], (uint32_t)(S * (OZ_KTILE + OZ_WTILE)));
                    tma_load_1d_hint(dst, kt + (size_t)ks * OZ_S * OZ_KTILE, S * OZ_KTILE, &full[s], pol_keep);
                    const int8_t* w_src = wq + ((size_t)panel * nks + ks) * OZ_S * OZ_WTILE;
                    if (w_streams) tma_load_1d_hint(dst + OZ_SLOT_W, w_src, S * OZ_WTILE, &full[s], pol_stream);
                    else tma_load_1d(dst + OZ_SLOT_W, w_src, S * OZ_WTILE, &full[s]);
                }
              }
            }
            if (fail) fail_s = 1;
        }
    } else if (warp == 4) {
        // ================= MMA issuer =================
        if (lane == 0) {
            int fail = 0;
            long u = 0;
            int rnd = 0;                                                     // panels issued so far, over all tiles
            for (int tile = tile0; tile < n_tiles && !fail; tile += tile_step)
              for (int pi = 0; pi < n_mine && !fail; ++pi, ++rnd) {
                const int panel = my_panel(pi);
                if (rnd > 0 && !oz_wait(acc_empty, (uint32_t)((rnd - 1) & 1), &fail, &fail_s)) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int ks = 0; ks < (panel + 1) * 2; ++ks, ++u) {
                    const int s = (int)(u % OZ_SLOTS);
                    if (!oz_wait(&full[s], (uint32_t)((u / OZ_SLOTS) & 1), &fail, &fail_s)) break;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t k_base = smem_u32(stages + (size_t)s * OZ_SLOT_BYTES);
                    oz_issue_slab<SCHEME>(tmem_base, oz_desc(k_base), oz_desc(k_base + OZ_SLOT_W), ks > 0 ? 1u : 0u);
                    oz_commit(&empty[s]);                                    // slot free once these MMAs retire
                }
                oz_commit(acc_full);                                         // accumulators of this panel complete
            }
            if (fail) fail_s = 1;
        }
    } else {
        // ================= epilogue: warps 0..3 drain W rows 0..31 of the panel from every group, warps 6..9 rows
        // 32..63.  A warp may only touch TMEM lanes 32*(warp%4)..+31, so thread (warp%4, lane) owns test point
        // 32*(warp%4)+lane.
        int fail = 0;
        double vsum = 0.0;                                  // sum over this thread's W rows of v^2 for its test point
        const int half = warp < 4 ? 0 : 1;
        const int tp = (warp & 3) * 32 + lane;
        const uint32_t col_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(half * 32);
        auto process_chunk = [&](const uint32_t (&d)[NG][8], const double* sg) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                // Combine the groups EXACTLY in int64 (weights 2^-BITS apart) in two parts, then integer add + DADD
                // converts to FP64 (I2F.F64 is slow).
                double hi, lo;
                if constexpr (SCHEME == 0) {             // |D| < 2^29: 21 + 29 < 51 bits; groups 5 / 9: 2^-33 / 2^-61
                    hi = oz_i64_to_double(((long long)(int)d[0][j] << 21) + ((long long)(int)d[1][j] << 14) +
                                          ((long long)(int)d[2][j] << 7) + (long long)(int)d[3][j]) * 0x1p-33;
                    lo = oz_i64_to_double(((long long)(int)d[4][j] << 21) + ((long long)(int)d[5][j] << 14) +
                                          ((long long)(int)d[6][j] << 7) + (long long)(int)d[7][j]) * 0x1p-61;
                } else {                                 // |D| < 2^31: groups 2..5 as two exact halves 2^16 apart
                    const double ha = oz_i64_to_double(((long long)(int)d[0][j] << 8) + (long long)(int)d[1][j]);
                    const double hb = oz_i64_to_double(((long long)(int)d[2][j] << 8) + (long long)(int)d[3][j]);
                    hi = fma(ha, 0x1p16, hb) * 0x1p-39;  // group 5 weight 2^(1-40)
                    lo = oz_i64_to_double(((long long)(int)d[4][j] << 16) + ((long long)(int)d[5][j] << 8) +
                                          (long long)(int)d[6][j]) * 0x1p-63;     // groups 6..8: 16 + 31 < 51 bits
                }
                const double v = sg[j] * (lo + hi);
                vsum = fma(v, v, vsum);
            }
        };
        // After a fault the loop keeps its shape (every epilogue warp meets the named barrier the same number of
        // times) and only the work is skipped.
        int rnd = 0;                                        // panels drained so far, over all tiles
        for (int tile = tile0; tile < n_tiles; tile += tile_step) {
          for (int pi = 0; pi < n_mine; ++pi, ++rnd) {
            const int panel = my_panel(pi);
            // row scales of this panel, double buffered: one named barrier per panel among the 8 epilogue warps
            double* sg = sigs + (rnd & 1) * OZ_PANEL;
            if (half == 0 && tp < OZ_PANEL) sg[tp] = wsig[panel * OZ_PANEL + tp] * kappa;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (!oz_wait(acc_full, (uint32_t)(rnd & 1), &fail, &fail_s)) continue;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // 4 chunks of 8 W rows, software pipelined: the TMEM read of chunk c+1 flies under the arithmetic of
            // chunk c (tcgen05.wait::ld waits for everything issued so far)
            const double* sgh = sg + half * 32;
            uint32_t da[NG][8], db[NG][8];
            oz_load_chunk<NG>(da, col_base);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            oz_load_chunk<NG>(db, col_base + 8);
            process_chunk(da, sgh);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            oz_load_chunk<NG>(da, col_base + 16);
            process_chunk(db, sgh + 8);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            oz_load_chunk<NG>(db, col_base + 24);
            process_chunk(da, sgh + 16);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // every TMEM read of this panel is complete: release the accumulators before the last arithmetic
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
            process_chunk(db, sgh + 24);
          }
          // ---- end of this CTA's share of the tile: the two column halves meet in shared memory, the share goes to
          // global memory, and the member of the gang that arrives last finishes the tile in a fixed order
          colred[half * 128 + tp] = vsum;
          vsum = 0.0;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (half == 0) partial[((size_t)tile * G + member) * 128 + tp] = colred[tp] + colred[128 + tp];
          __threadfence();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (half == 0 && tp == 0) last_s = atomicAdd(&arrived[tile], 1);
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (last_s == G - 1 && half == 0) {
              __threadfence();
              const long t = (long)tile * 128 + tp;
              if (t < k) {
                  const volatile double* pt = partial + (size_t)tile * G * 128;
                  double vv = pt[tp];
                  for (int c = 1; c < G; ++c) vv += pt[c * 128 + tp];                         // fixed order: member 0, 1, ...
                  double var = __dadd_rn(kdiag, -vv);                                        // sklearn:_gpr.py:480-481
                  if (var < 0.0) var = 0.0;                                                  // :485-491
                  std_out[t * ld + q] = sqrt(__dmul_rn(var, __dmul_rn(y_std, y_std)));       // :494,500
              }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");      // colred and last_s are free for the next tile
        }
        if (fail) fail_s = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    if (tid == 0 && fail_s) *fail_flag = 1;
}

// ---------------------------------------------------------------- W -> row scales + int8 slices (upload time)
// src: row-major ld-strided lower-triangular W (transposed = 1: src holds W^T).  One warp per row.
__global__ void k_slice_w(const double* __restrict__ src, int n, int ld, int transposed, int n_pad, int scheme,
                          int8_t* __restrict__ wq, double* __restrict__ wsig) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n_pad) return;
    const int nks = n_pad / 32;
    double mx = 0.0;
    if (row < n)
        for (int j = lane; j <= row; j += 32) {
            const double w = transposed ? src[(size_t)j * ld + row] : src[(size_t)row * ld + j];
            mx = fmax(mx, fabs(w));
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    int e = 0;
    double sig = 1.0;
    if (mx > 0.0) { frexp(mx, &e); sig = ldexp(1.0, e + (scheme == 1 ? 1 : 0)); }   // |w|/sig < 1 (scheme 0) or < 1/2 (scheme 1)
    if (lane == 0) wsig[row] = sig;
    const double inv = 1.0 / sig;                           // power of two: exact
    const int panel = row / OZ_PANEL, r = row % OZ_PANEL;      // tile (panel, k slab, slice): 64 rows x 32 B
    for (int j = lane; j < n_pad; j += 32) {
        double w = 0.0;
        if (row < n && j <= row) w = (transposed ? src[(size_t)j * ld + row] : src[(size_t)row * ld + j]) * inv;
        const int ks = j >> 5, b = j & 31;
        int8_t* base = wq + ((size_t)panel * nks + ks) * OZ_S * OZ_WTILE + (r >> 3) * 256 + (b >> 4) * 128 + (r & 7) * 16 + (b & 15);
        if (scheme == 0) {                                  // 8 x 7-bit digits, greedy from the top (all exact in FP64)
            double y = w * 64.0;
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const double qd = rint(y);
                base[(size_t)a * OZ_WTILE] = (int8_t)(int)qd;
                y = (y - qd) * 128.0;                       // exact: |y - qd| <= 1/2
            }
        } else {                                            // 7 signed base-256 digits of X = rint(w 2^55), |X| < 2^54
            long long X = llrint(w * 0x1p55);
#pragma unroll
            for (int a = 6; a >= 1; --a) {
                const long long dg = ((X + 128) & 255) - 128;
                base[(size_t)a * OZ_WTILE] = (int8_t)dg;
                X = (X - dg) >> 8;
            }
            base[0] = (int8_t)X;                            // |X| <= 64
            base[(size_t)7 * OZ_WTILE] = 0;
        }
    }
}

int launch_slice_w(const double* src, int n, int ld, int transposed, int scheme, int8_t* wq, double* wsig, cudaStream_t st) {
    const int n_pad = padded_n(n);
    ProfScope prof(PROF_OTHER, st);
    k_slice_w<<<(n_pad + 7) / 8, 256, 0, st>>>(src, n, ld, transposed, n_pad, scheme, wq, wsig);
    BK_LAUNCH_CHECK("k_slice_w");
    return 0;
}

// CTAs per tile of test points: the live K* tiles ((SMs / G) x n_pad x 896 bytes) should fit in about half of the L2
// (126 MB on B200); candidates are the sizes that leave at most four SMs without a gang on 148 SMs.
int oz_gang_size(int n_pad, int sms) {
    const int cand[] = {2, 4, 8, 16, 37};
    const double budget = 72e6, tile_bytes = (double)n_pad * 128.0 * 7.0;
    for (int g : cand)
        if (g <= sms && (double)(sms / g) * tile_bytes <= budget) return g;
    return sms >= 37 ? 37 : 2;
}

int launch_var_ozaki(const bk_gp_s* h, int q, int scheme, const int8_t* kq, double kappa, long k, double* std_out,
                     double* scratch, int* fail_flag, cudaStream_t st) {
    const GpOutput& o = h->out[q];
    if (!o.wq || !o.wsig) return set_error("int8 variance path requested but output %d has no sliced L^-1", q);
    static OncePerDevice attr_set;
    if (attr_set.first()) {
        BK_CUDA(cudaFuncSetAttribute(k_var_ozaki<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM));
        BK_CUDA(cudaFuncSetAttribute(k_var_ozaki<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM));
    }
    // scratch: [tiles][G][128] partial sums, then [tiles] arrival counters
    const unsigned tiles = (unsigned)((k + 127) / 128);
    static int sm_count[64] = {};
    int dev = 0;
    BK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return set_error("device ordinal %d out of range", dev);
    if (!sm_count[dev]) BK_CUDA(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
    const int G = oz_gang_size(h->n_pad, sm_count[dev]);
    unsigned gangs = (unsigned)(sm_count[dev] / G);                   // one CTA per SM: persistent gangs
    if (gangs > tiles) gangs = tiles;
    const unsigned grid = gangs * (unsigned)G;
    double* partial = scratch;
    int* arrived = reinterpret_cast<int*>(scratch + (size_t)tiles * G * 128);
    BK_CUDA(cudaMemsetAsync(arrived, 0, sizeof(int) * tiles, st));
    const int w_streams = (double)h->n_pad * h->n_pad * 3.5 > 100e6 ? 1 : 0;       // bytes of the lower triangle's slices
    ProfScope prof(PROF_VAR, st);
    if (scheme == 1)
        k_var_ozaki<1><<<grid, OZ_THREADS, OZ_SMEM, st>>>(o.wq, o.wsig, kq, h->n_pad, kappa, o.kdiag, o.y_std,
                                                          std_out, k, h->m, q, partial, arrived, G, w_streams, fail_flag);
    else
        k_var_ozaki<0><<<grid, OZ_THREADS, OZ_SMEM, st>>>(o.wq, o.wsig, kq, h->n_pad, kappa, o.kdiag, o.y_std,
                                                          std_out, k, h->m, q, partial, arrived, G, w_streams, fail_flag);
    BK_LAUNCH_CHECK("k_var_ozaki");
    return 0;
}

}  // namespace bk
