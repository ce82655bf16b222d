// Shared device/host helpers for the batman_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace bk {

// ---------------------------------------------------------------- errors
extern thread_local char g_err[512];
int set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

// ---------------------------------------------------------------- optional per-kernel CUDA-event timing
// (bk_profile_enable / bk_profile_read in the C ABI; used by bench.py for the roofline numbers)
enum ProfCat { PROF_PREDICT = 0, PROF_VAR = 1, PROF_SOBOL = 2, PROF_SOBOL_Y = 3, PROF_OTHER = 4, PROF_CHOL = 5, PROF_POD = 6, PROF_NCAT = 8 };
struct ProfScope {
    cudaStream_t st;
    int idx;
    ProfScope(int cat, cudaStream_t s);
    ~ProfScope();
};

// "once per device" latch: function attributes (dynamic shared-memory size) are per device, and one process may
// drive several GPUs
struct OncePerDevice {
    bool done[64] = {};
    bool first() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};

#define BK_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return bk::cuda_fail(e__, #call); } while (0)
#define BK_LAUNCH_CHECK(name) do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return bk::cuda_fail(e__, name); } while (0)

// ---------------------------------------------------------------- mbarrier + bulk async copy (TMA 1-D)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// global -> shared bulk copy, completion counted in bytes on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- FP64 tensor core (SASS: DMMA.8x8x4)
// D(8x8) += A(8x4, row) * B(4x8, col).  lane = g*4 + t:
//   a = A[g][t], b = B[t][g], c0/c1 = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---------------------------------------------------------------- compensated arithmetic
struct dd { double hi, lo; };

__device__ __forceinline__ void two_sum(double a, double b, double& s, double& e) {
    s = __dadd_rn(a, b);
    double bb = __dadd_rn(s, -a);
    e = __dadd_rn(__dadd_rn(a, -__dadd_rn(s, -bb)), __dadd_rn(b, -bb));
}
__device__ __forceinline__ void two_prod(double a, double b, double& p, double& e) {
    p = __dmul_rn(a, b);
    e = __fma_rn(a, b, -p);
}
// acc += x  (acc is an unevaluated hi+lo pair)
__device__ __forceinline__ void dd_add_d(dd& acc, double x) {
    double s, e;
    two_sum(acc.hi, x, s, e);
    acc.hi = s;
    acc.lo = __dadd_rn(acc.lo, e);
}
__device__ __forceinline__ dd dd_norm(dd a) {
    double s, e;
    two_sum(a.hi, a.lo, s, e);
    return dd{s, e};
}
__device__ __forceinline__ dd dd_add(dd a, dd b) {
    double s, e;
    two_sum(a.hi, b.hi, s, e);
    e = __dadd_rn(e, __dadd_rn(a.lo, b.lo));
    double hi, lo;
    two_sum(s, e, hi, lo);
    return dd{hi, lo};
}
__device__ __forceinline__ dd dd_neg(dd a) { return dd{-a.hi, -a.lo}; }
__device__ __forceinline__ dd dd_sub(dd a, dd b) { return dd_add(a, dd_neg(b)); }
__device__ __forceinline__ dd dd_mul(dd a, dd b) {
    double p, e;
    two_prod(a.hi, b.hi, p, e);
    e = __dadd_rn(e, __dadd_rn(__dmul_rn(a.hi, b.lo), __dmul_rn(a.lo, b.hi)));
    double hi, lo;
    two_sum(p, e, hi, lo);
    return dd{hi, lo};
}
__device__ __forceinline__ dd dd_from(double a) { return dd{a, 0.0}; }
__device__ __forceinline__ dd dd_div(dd a, dd b) {
    double q1 = a.hi / b.hi;
    dd r = dd_sub(a, dd_mul(dd_from(q1), b));
    double q2 = r.hi / b.hi;
    r = dd_sub(r, dd_mul(dd_from(q2), b));
    double q3 = r.hi / b.hi;
    double hi, lo;
    two_sum(q1, q2, hi, lo);
    return dd_add(dd{hi, lo}, dd_from(q3));
}
__device__ __forceinline__ double dd_to_double(dd a) { return __dadd_rn(a.hi, a.lo); }

}  // namespace bk
