// K* writers of K1 (gp_predict.cu, gp_predict_gen.cu): 8 consecutive covariances of one test point go to HBM in the
// operand layout the sigma kernel wants -- FP64 DMMA B fragments (gp_var.cu) or error-free int8 slices (var_ozaki.cu).
#pragma once
#include "common.cuh"
#include "model.h"

namespace bk {

// destination bases of test point t (tile t / 128, row t % 128 inside the tile)
__device__ __forceinline__ double* kstar_tile_base(double* kstar, long t, int n_pad) {
    const long T = t / TILE_T;
    const int tl = (int)(t % TILE_T);
    return kstar + (size_t)T * (n_pad / TILE_K) * TILE_DOUBLES + (tl / 8) * 128 + (tl % 8) * 8;
}
__device__ __forceinline__ int8_t* kq_tile_base(int8_t* kq, long t, int n_pad) {
    const int r = (int)(t % TILE_T);
    return kq + (size_t)(t / TILE_T) * (n_pad / 32) * 8 * 4096 + (r >> 3) * 256 + (r & 7) * 16;
}

// WRITE_K: 1 = FP64 fragment tiles, 2 = 8 signed 7-bit slices, 3 = 7 unsigned byte slices.
// kb[0..7] = k(x_t, X_jg .. X_jg+7), jg a multiple of 8.
template <int WRITE_K>
__device__ __forceinline__ void kstar_write8(const double (&kb)[8], int jg, double* kdst, int8_t* qdst, double inv_kappa) {
    if constexpr (WRITE_K == 2) {
        // error-free int8 slicing: k / kappa = sum_a q_a 2^-(7a-1), |q_a| <= 64 (var_ozaki.cu)
        unsigned int pk[8][2];
#pragma unroll
        for (int a = 0; a < 8; ++a) pk[a][0] = pk[a][1] = 0u;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            // X = rint(k/kappa * 2^55) = hi * 2^28 + lo from two magic-number roundings (no F2I), then
            // signed base-128 digits in [-64, 63] with 32-bit integer ops: X = sum_a q_a 2^(56-7a)
            const double x27 = __dmul_rn(__dmul_rn(kb[jj], inv_kappa), 0x1p27);
            const double hm = __dadd_rn(x27, 6755399441055744.0);
            int hi = __double2loint(hm);                                  // rint(x * 2^27), |hi| <= 2^27
            const double r28 = __dmul_rn(__dadd_rn(x27, -__dadd_rn(hm, -6755399441055744.0)), 0x1p28);
            int lo = __double2loint(__dadd_rn(r28, 6755399441055744.0));  // |lo| <= 2^27
            const int sh = 8 * (jj & 3), w = jj >> 2;
#pragma unroll
            for (int a = 7; a >= 4; --a) {                                // slices 8..5 from lo
                const int dgt = ((lo + 64) & 127) - 64;
                pk[a][w] |= (unsigned int)(dgt & 255) << sh;
                lo = (lo - dgt) >> 7;
            }
            hi += lo;                                                     // carry out of the low half
#pragma unroll
            for (int a = 3; a >= 1; --a) {                                // slices 4..2 from hi
                const int dgt = ((hi + 64) & 127) - 64;
                pk[a][w] |= (unsigned int)(dgt & 255) << sh;
                hi = (hi - dgt) >> 7;
            }
            pk[0][w] |= (unsigned int)(hi & 255) << sh;                   // slice 1: |hi| <= 64
        }
        int8_t* d = qdst + (size_t)(jg >> 5) * 8 * 4096 + ((jg & 31) >> 4) * 128 + (jg & 15);
#pragma unroll
        for (int a = 0; a < 8; ++a) *reinterpret_cast<uint2*>(d + (size_t)a * 4096) = make_uint2(pk[a][0], pk[a][1]);
    }
    if constexpr (WRITE_K == 3) {
        // unsigned byte slicing: X = rint(k/kappa * 2^56) in [0, 2^56) = hi * 2^32 + lo (hi 24 bits, lo 32 bits, both
        // byte aligned); slice b = byte 6-b of X.  Two magic-number roundings (no F2I), then the bytes of four
        // consecutive covariances are gathered with byte permutes (3 PRMT per slice word instead of 16 shift/mask ops:
        // the kernel is close to issue-slot bound, so integer instructions cost as much as DP ones).
        unsigned int hi[8], lo[8];
        const double scale24 = __dmul_rn(inv_kappa, 0x1p24);                              // power of two: exact
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const double x24 = __dmul_rn(kb[jj], scale24);
            const double hm = __dadd_rn(x24, 6755399441055744.0);
            int h = __double2loint(hm);                                                   // rint(x * 2^24) in [0, 2^24]
            const double r32 = __dmul_rn(__dadd_rn(x24, -__dadd_rn(hm, -6755399441055744.0)), 0x1p32);   // [-2^31, 2^31]
            int l = __double2loint(__dadd_rn(r32, 6755399441055744.0));                   // rint(r32) mod 2^32
            if (r32 < 0.0) h -= 1;                                                        // borrow: lo = r32 + 2^32
            if (h < 0) { h = 0; l = 0; }                                                  // k slightly below 0 by rounding: clamp
            hi[jj] = (unsigned int)h;
            lo[jj] = (unsigned int)l;
        }
        unsigned int pk[7][2];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const unsigned int *H = hi + 4 * w, *L = lo + 4 * w;
            // byte B of four words -> one word [v0.B | v1.B << 8 | v2.B << 16 | v3.B << 24]
#define BK_GATHER(V, B) __byte_perm(__byte_perm(V[0], V[1], (B) | ((4 + (B)) << 4)), \
                                    __byte_perm(V[2], V[3], (B) | ((4 + (B)) << 4)), 0x5410)
            pk[0][w] = BK_GATHER(H, 2);                                   // bits 48..55
            pk[1][w] = BK_GATHER(H, 1);
            pk[2][w] = BK_GATHER(H, 0);                                   // bits 32..39
            pk[3][w] = BK_GATHER(L, 3);                                   // bits 24..31
            pk[4][w] = BK_GATHER(L, 2);
            pk[5][w] = BK_GATHER(L, 1);
            pk[6][w] = BK_GATHER(L, 0);                                   // bits 0..7
#undef BK_GATHER
        }
        int8_t* d = qdst + (size_t)(jg >> 5) * 8 * 4096 + ((jg & 31) >> 4) * 128 + (jg & 15);
#pragma unroll
        for (int a = 0; a < 7; ++a) *reinterpret_cast<uint2*>(d + (size_t)a * 4096) = make_uint2(pk[a][0], pk[a][1]);
    }
    if constexpr (WRITE_K == 1) {
        double* d = kdst + (size_t)(jg / TILE_K) * TILE_DOUBLES + ((jg % TILE_K) / 8) * 64;
        // lane order inside an 8-wide k group: (tq, h) -> j = tq + 4h
        reinterpret_cast<double2*>(d)[0] = make_double2(kb[0], kb[4]);
        reinterpret_cast<double2*>(d)[1] = make_double2(kb[1], kb[5]);
        reinterpret_cast<double2*>(d)[2] = make_double2(kb[2], kb[6]);
        reinterpret_cast<double2*>(d)[3] = make_double2(kb[3], kb[7]);
    }
}

}  // namespace bk
