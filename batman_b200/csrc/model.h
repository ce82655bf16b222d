// Host-side model handle behind the C ABI (include/batman_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/batman_b200.h"

namespace bk {

constexpr int MAX_P = 32;        // largest padded input dimension with an instantiated kernel
constexpr int TILE_T = 128;      // test points per DMMA column tile
constexpr int TILE_M = 128;      // rows of W per DMMA row panel
constexpr int TILE_K = 16;       // train points per staged K slab
constexpr int TILE_DOUBLES = TILE_M * TILE_K;   // 2048 doubles = 16 KiB per packed tile

// General covariance (kind == BK_GENERAL): k = sum_t coef_t * prod_f stat_{fkind[f]}(||(x - x')/ls_f||)^texp[t][f],
// the expanded form of a Sum/Product tree over Constant/White/RBF/Matern leaves (kernel_spec.py).
constexpr int GEN_MAX_F = 3;     // stationary leaves
constexpr int GEN_MAX_T = 6;     // product terms
constexpr int GEN_MAX_FP = 48;   // nf * P that fits the staging buffers of the general kernels
struct GenSpec {
    int nf = 0, nt = 0;
    int fkind[GEN_MAX_F] = {0, 0, 0};
    double tcoef[GEN_MAX_T] = {0, 0, 0, 0, 0, 0};
    unsigned char texp[GEN_MAX_T][GEN_MAX_F] = {};
};

struct GpOutput {
    int kind = -1;
    GenSpec gen;                     // kind == BK_GENERAL: xs is n_pad x (nf * P), factor-major inside a row
    double gls[GEN_MAX_F][MAX_P];    //   length scales of the factors
    double kmax = 1;                 // bound on |k(x, x')|: |c0| + |c1|, or sum |coef_t|
    bool nonneg = true;              // every coefficient >= 0 (K* can be sliced into unsigned bytes)
    double c0 = 0, c1 = 1, kdiag = 1, y_mean = 0, y_std = 1;
    double ls[MAX_P];
    const double* xs = nullptr;      // n_pad x P   (X_train / ls, zero padded)
    const double* alpha = nullptr;   // n_pad       (zero padded)
    const double* wt = nullptr;      // packed L^-1 tiles, or null
    const int8_t* wq = nullptr;      // int8 slices of L^-1 (Ozaki path), or null
    const double* wsig = nullptr;    // per-row power-of-two scales of the slices
};

}  // namespace bk

// Two internal streams of a sigma prediction: K1 of unit u+1 (FP64 pipe) runs beside the sigma kernel of unit u
// (tensor pipe) on double-buffered K* workspaces.  Created on first use, on the device current at that time.
struct PredictLanes {
    cudaStream_t k1 = nullptr, k2 = nullptr;
    cudaEvent_t fork = nullptr, k1done[2] = {nullptr, nullptr}, k2done[2] = {nullptr, nullptr};
    int device = -1;
};

struct bk_gp_s {
    int n = 0, p = 0, P = 0, m = 0, n_pad = 0;
    int var_mode = 0;                // 0 = FP64 DMMA (k_var_trmm), 1 / 2 = int8 tcgen05 (k_var_ozaki, scheme 0 / 1)
    PredictLanes* lanes = nullptr;
    bool overlap = false;            // bk_gp_set_overlap: K1 of unit u+1 beside the sigma kernel of unit u
    double scale[bk::MAX_P];
    double mn[bk::MAX_P];
    std::vector<bk::GpOutput> out;
};

namespace bk {
inline int padded_dim(int p) {
    const int sizes[] = {2, 4, 6, 8, 10, 12, 16, 20, 24, 32};
    for (int s : sizes)
        if (p <= s) return s;
    return -1;
}
inline int padded_n(int n) { return (n + TILE_M - 1) / TILE_M * TILE_M; }
}  // namespace bk
