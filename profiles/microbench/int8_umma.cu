// Feasibility probe for an Ozaki-scheme (error-free int8 slicing) variant of the variance TRMM:
// sustained tcgen05.mma kind::i8 throughput with SMEM-resident operands in a hand-packed, non-swizzled
// K-major layout (core matrices of 8 rows x 16 bytes; LBO = 128 B along K, SBO = 256 B along M/N),
// plus a correctness check of the descriptor / TMEM-load plumbing on known data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o int8_umma int8_umma.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: element (row r, byte b of a 32-byte K slab):
//   off = (r/8)*256 + (b/16)*128 + (r%8)*16 + b%16
__host__ __device__ inline int packed_off(int r, int b) { return (r >> 3) * 256 + (b >> 4) * 128 + (r & 7) * 16 + (b & 15); }

__device__ __forceinline__ uint64_t make_desc(const void* smem, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)(lbo_bytes >> 4) << 16;                     // leading byte offset, bits [16,30)
    d |= (uint64_t)(sbo_bytes >> 4) << 32;                     // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                    // descriptor version (sm_100)
    return d;                                                  // layout_type = 0 (no swizzle), base_offset = 0
}
__host__ __device__ inline uint32_t make_idesc_i8(int M, int N) {
    uint32_t d = 0;
    d |= 2u << 4;                   // c_format = S32
    d |= 1u << 7;                   // a_format = signed 8 bit
    d |= 1u << 10;                  // b_format = signed 8 bit
    d |= (uint32_t)(N >> 3) << 17;  // n_dim
    d |= (uint32_t)(M >> 4) << 24;  // m_dim
    return d;                       // K-major A and B, dense, no saturate
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

template <int N, int NSLICE>
__global__ void __launch_bounds__(128, 1) k_umma_probe(int iters, int* out /* [gridDim][128][N] of group 0 */, int check) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* As = smem;                                 // NSLICE x (128 x 32 B)
    unsigned char* Bs = smem + NSLICE * 4096;                 // NSLICE x (N x 32 B)
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // A slice a: A[r][b] = (r % 7) - 3 + a ;  B slice s: B[c][b] = (c % 5) - 2 - s   (small signed ints)
    for (int i = tid; i < NSLICE * 128 * 32; i += 128) {
        const int a = i / 4096, r = (i % 4096) / 32, b = i % 32;
        As[a * 4096 + packed_off(r, b)] = (unsigned char)(signed char)((r % 7) - 3 + a);
    }
    for (int i = tid; i < NSLICE * N * 32; i += 128) {
        const int s = i / (N * 32), c = (i % (N * 32)) / 32, b = i % 32;
        Bs[s * N * 32 + packed_off(c, b)] = (unsigned char)(signed char)((c % 5) - 2 - s);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (UMMA) reads
    __syncthreads();
    constexpr int GROUPS = 512 / N;                                // accumulator groups that fit in TMEM
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = make_idesc_i8(128, N);
    if (tid == 0) {
        uint32_t seen = 0;
        for (int it = 0; it < iters; ++it) {
            // every (a, b) slice pair into accumulator group (a + b) % GROUPS: NSLICE^2 MMAs per iteration
#pragma unroll 1
            for (int a = 0; a < NSLICE; ++a)
#pragma unroll 1
                for (int b = 0; b < NSLICE; ++b) {
                    const int g = (a + b) % GROUPS;
                    umma_i8(tmem_base + g * N, make_desc(As + a * 4096, 128, 256), make_desc(Bs + b * N * 32, 128, 256),
                            idesc, (seen >> g) & 1u);
                    seen |= 1u << g;
                }
        }
        umma_commit(&bar);
    }
    for (long spin = 0; !mbar_try_wait(&bar, 0); ++spin)
        if (spin > (1L << 22)) { if (tid == 0) printf("TIMEOUT waiting for UMMA commit (block %d)\n", blockIdx.x); break; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (check) {
        // thread (warp w, lane l) owns TMEM lane 32w + l = output row; read group 0, N columns, 8 at a time
        const int row = warp * 32 + lane;
        for (int c0 = 0; c0 < N; c0 += 8) {
            uint32_t v[8];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 8; ++j) out[((size_t)blockIdx.x * 128 + row) * N + c0 + j] = (int)v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

template <int N, int NSLICE>
static void run(int sms) {
    const size_t smem = (size_t)NSLICE * 4096 + (size_t)NSLICE * N * 32 + 1024;
    CK(cudaFuncSetAttribute(k_umma_probe<N, NSLICE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int* out; CK(cudaMalloc(&out, (size_t)sms * 128 * N * sizeof(int)));
    // correctness: 1 iteration, group 0 gets the pairs with (a + b) % GROUPS == 0
    k_umma_probe<N, NSLICE><<<sms, 128, smem>>>(1, out, 1);
    CK(cudaDeviceSynchronize());
    int* h = (int*)malloc((size_t)128 * N * sizeof(int));
    CK(cudaMemcpy(h, out, (size_t)128 * N * sizeof(int), cudaMemcpyDeviceToHost));
    constexpr int GROUPS = 512 / N;
    long bad = 0;
    for (int r = 0; r < 128; ++r)
        for (int c = 0; c < N; ++c) {
            long want = 0;
            for (int a = 0; a < NSLICE; ++a)
                for (int b = 0; b < NSLICE; ++b)
                    if ((a + b) % GROUPS == 0) want += 32L * ((r % 7) - 3 + a) * ((c % 5) - 2 - b);
            if (h[r * N + c] != (int)want) { if (bad < 5) printf("  mismatch r=%d c=%d got %d want %ld\n", r, c, h[r * N + c], want); ++bad; }
        }
    // throughput
    const int iters = 2000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_umma_probe<N, NSLICE><<<sms, 128, smem>>>(iters, out, 0);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        k_umma_probe<N, NSLICE><<<sms, 128, smem>>>(iters, out, 0);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double macs = (double)sms * iters * NSLICE * NSLICE * 128.0 * N * 32.0;
    printf("{\"kernel\": \"tcgen05.mma kind::i8 M128 N%d K32, SS no-swizzle, %d x %d slice pairs\", \"mismatches\": %ld, "
           "\"tops\": %.1f, \"ms\": %.3f}\n", N, NSLICE, NSLICE, bad, 2.0 * macs / (best * 1e-3) * 1e-12, best);
    free(h); CK(cudaFree(out));
}

// The same kernel launched back to back for `seconds`: the rate a long step can sustain under the board's power cap
// (MEASURED_PEAKS.json distinguishes burst and sustained bf16 the same way); measured over the last 3/4 of the run.
template <int N, int NSLICE>
static void run_sustained(int sms, double seconds) {
    const size_t smem = (size_t)NSLICE * 4096 + (size_t)NSLICE * N * 32 + 1024;
    CK(cudaFuncSetAttribute(k_umma_probe<N, NSLICE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int* out; CK(cudaMalloc(&out, (size_t)sms * 128 * N * sizeof(int)));
    const int iters = 2000;
    cudaEvent_t e0, e1, ew; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&ew));
    CK(cudaEventRecord(ew));
    long launches = 0, counted = 0;
    bool started = false;
    for (;;) {
        k_umma_probe<N, NSLICE><<<sms, 128, smem>>>(iters, out, 0);
        ++launches;
        if (started) ++counted;
        if (launches % 16 == 0) {
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, ew, e1));
            if (!started && ms > 250.0 * seconds) { CK(cudaEventRecord(e0)); started = true; }
            if (ms > 1000.0 * seconds) break;
        }
    }
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double macs = (double)sms * iters * NSLICE * NSLICE * 128.0 * N * 32.0 * counted;
    printf("{\"kernel\": \"tcgen05.mma kind::i8 M128 N%d K32 sustained for %.0f s (back-to-back launches)\", "
           "\"tops\": %.1f, \"launches\": %ld}\n", N, seconds, 2.0 * macs / (ms * 1e-3) * 1e-12, counted);
    CK(cudaFree(out));
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    run<64, 8>(sms);
    run<128, 8>(sms);
    run<192, 6>(sms);
    run<256, 6>(sms);
    run_sustained<256, 6>(sms, 4.0);
    run_sustained<128, 8>(sms, 4.0);
    return 0;
}
