// FP64 pipe micro-benchmarks for B200 (sm_100a).
// MEASURED_PEAKS.json has no FP64 entry; the hot path (Kriging predict + L^-1 k*) is
// FP64-bound, so these numbers are the roofline denominators quoted in DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_dadd_dmul(double* out, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = __dmul_rn(x[i], a); x[i] = __dadd_rn(x[i], b); }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

// m8n8k4: A 1 reg, B 1 reg, C 2 regs
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma884(double* out, double a, double b) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

// m16n8k4: A 2 regs, B 1 reg, C 4 regs
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma1684(double* out, double a, double b) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

// m16n8k8: A 4 regs, B 2 regs, C 4 regs
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, double a, double b) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

// m16n8k16: A 8 regs, B 4 regs, C 4 regs
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma16816(double* out, double a, double b) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                         "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b),
                           "d"(b), "d"(a), "d"(b), "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

// DMMA and DFMA interleaved in the same warp: do they share the FP64 datapath?
__global__ void __launch_bounds__(256) k_mixed(double* out, double a, double b) {
    double c[4][4];
    double x[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

// transcendental throughput as the mean kernel uses them
__global__ void __launch_bounds__(256) k_exp(double* out, double a, double b) {
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = -(threadIdx.x * 1e-3 + i);
    double s = 0;
    for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { s += exp(x[i]); x[i] = x[i] * a - b; }
    }
    if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) k_sqrt(double* out, double a, double b) {
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = (threadIdx.x * 1e-3 + i + 1);
    double s = 0;
    for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { s += sqrt(x[i]); x[i] = x[i] * a + b; }
    }
    if (s == 12345.678) out[0] = s;
}

template <typename F>
static double timeit(F launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best * 1e-3;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    double* out; CK(cudaMalloc(&out, 8));
    const double a = 0.999999, b = 1e-9;
    for (int bps : {1, 2, 4, 8}) {
        int grid = sms * bps;
        double thr = (double)grid * 256;
        double t;
        t = timeit([&] { k_dfma<<<grid, 256>>>(out, a, b); });
        printf("{\"kernel\": \"dfma\", \"blocks_per_sm\": %d, \"tflops\": %.3f}\n", bps, thr * ITERS * 8 * 2 / t * 1e-12);
        t = timeit([&] { k_dadd_dmul<<<grid, 256>>>(out, a, b); });
        printf("{\"kernel\": \"dmul+dadd\", \"blocks_per_sm\": %d, \"tinstr_per_s\": %.3f}\n", bps, thr * ITERS * 16 / t * 1e-12);
        double warps = (double)grid * 8;
        t = timeit([&] { k_dmma884<8><<<grid, 256>>>(out, a, b); });
        printf("{\"kernel\": \"dmma.m8n8k4 x8acc\", \"blocks_per_sm\": %d, \"tflops\": %.3f}\n", bps, warps * ITERS * 8 * (8 * 8 * 4 * 2.0) / t * 1e-12);
        t = timeit([&] { k_dmma1684<8><<<grid, 256>>>(out, a, b); });
        printf("{\"kernel\": \"dmma.m16n8k4 x8acc\", \"blocks_per_sm\": %d, \"tflops\": %.3f}\n", bps, warps * ITERS * 8 * (16 * 8 * 4 * 2.0) / t * 1e-12);
        t = timeit([&] { k_dmma1688<8><<<grid, 256>>>(out, a, b); });
        printf("{\"kernel\": \"dmma.m16n8k8 x8acc\", \"blocks_per_sm\": %d, \"tflops\": %.3f}\n", bps, warps * ITERS * 8 * (16 * 8 * 8 * 2.0) / t * 1e-12);
        t = timeit([&] { k_dmma1688<2><<<grid, 256>>>(out, a, b); });
        printf("{\"kernel\": \"dmma.m16n8k8 x2acc\", \"blocks_per_sm\": %d, \"tflops\": %.3f}\n", bps, warps * ITERS * 2 * (16 * 8 * 8 * 2.0) / t * 1e-12);
        t = timeit([&] { k_dmma16816<8><<<grid, 256>>>(out, a, b); });
        printf("{\"kernel\": \"dmma.m16n8k16 x8acc\", \"blocks_per_sm\": %d, \"tflops\": %.3f}\n", bps, warps * ITERS * 8 * (16 * 8 * 16 * 2.0) / t * 1e-12);
        t = timeit([&] { k_mixed<<<grid, 256>>>(out, a, b); });
        printf("{\"kernel\": \"mixed 4xdmma.m16n8k8 + 8xdfma\", \"blocks_per_sm\": %d, \"tflops_dmma\": %.3f, \"tflops_dfma\": %.3f}\n", bps,
               warps * ITERS * 4 * (16 * 8 * 8 * 2.0) / t * 1e-12, thr * ITERS * 8 * 2 / t * 1e-12);
        t = timeit([&] { k_exp<<<grid, 256>>>(out, a, b); });
        printf("{\"kernel\": \"exp(double)\", \"blocks_per_sm\": %d, \"gexp_per_s\": %.2f}\n", bps, thr * (ITERS / 8) * 4 / t * 1e-9);
        t = timeit([&] { k_sqrt<<<grid, 256>>>(out, a, b); });
        printf("{\"kernel\": \"sqrt(double)\", \"blocks_per_sm\": %d, \"gsqrt_per_s\": %.2f}\n", bps, thr * (ITERS / 8) * 4 / t * 1e-9);
    }
    return 0;
}
