// pipe_bench3: prototype of a carry-free Montgomery product: 9 limbs of 29 bits (R = 2^261), 64-bit
// column accumulators fed by PLAIN IMAD.WIDE.U32 (no carry predicate), carries extracted by shifts.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "fq.cuh"
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

// p in radix 2^29
__device__ __constant__ const uint32_t P29c[9] = {0};
#define M29 0x1fffffffu

__device__ __forceinline__ void mac(uint64_t& c, uint32_t a, uint32_t b) {
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c) : "r"(a), "r"(b));
}
template <bool SQR>
__device__ __forceinline__ void mul29(uint32_t (&r)[9], const uint32_t (&a)[9], const uint32_t (&b)[9]) {
    // p = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47 in 29-bit limbs
    const uint32_t P[9] = {0x187cfd47u, 0x10460b6u, 0x1c72a34fu, 0x2d522d0u, 0x1585d978u, 0x2db40c0u, 0xa6e141u, 0xe5c2634u, 0x30644eu};
    const uint32_t N0 = 0x4866389u;  // -p^-1 mod 2^29
    uint64_t c[18];
#pragma unroll
    for (int k = 0; k < 18; k++) c[k] = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
#pragma unroll
        for (int j = 0; j < 9; j++) c[i + j] += (uint64_t)a[j] * b[i];
        uint32_t m; asm("mul.lo.u32 %0, %1, %2;" : "=r"(m) : "r"((uint32_t)c[i]), "r"(N0)); m &= M29;
#pragma unroll
        for (int j = 0; j < 9; j++) c[i + j] += (uint64_t)m * P[j];
        c[i + 1] += c[i] >> 29;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        r[k] = (uint32_t)c[9 + k] & M29;
        c[10 + k] += c[9 + k] >> 29;
    }
    r[8] = (uint32_t)c[17];
}

__global__ void k_mul29(uint32_t* out, int iters) {
    uint32_t x[9], y[9];
    for (int k = 0; k < 9; k++) { x[k] = (threadIdx.x * 977u + k * 13u) & M29; y[k] = (blockIdx.x * 31u + threadIdx.x * 7u + k + 5u) & M29; }
    for (int it = 0; it < iters; it++) mul29<false>(x, x, y);
    uint32_t s = 0;
    for (int k = 0; k < 9; k++) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mul29x2(uint32_t* out, int iters) {
    uint32_t x[9], y[9], z[9];
    for (int k = 0; k < 9; k++) { x[k] = (threadIdx.x * 977u + k * 13u) & M29; y[k] = (blockIdx.x * 31u + threadIdx.x * 7u + k + 5u) & M29; z[k] = x[k] ^ 5; }
    for (int it = 0; it < iters; it++) { mul29<false>(x, x, y); mul29<false>(z, z, y); }
    uint32_t s = 0;
    for (int k = 0; k < 9; k++) s ^= x[k] ^ z[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mul32(uint32_t* out, int iters) {
    fq x, y;
    for (int k = 0; k < 8; k++) { x.v[k] = threadIdx.x * 977u + k * 13u; y.v[k] = blockIdx.x * 31u + threadIdx.x * 7u + k + 5u; }
    x.v[7] &= 0x0fffffffu; y.v[7] &= 0x0fffffffu;
    for (int it = 0; it < iters; it++) x = fq_mul_inline(x, y);
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s ^= x.v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float time_it(F launch) {
    cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    launch(); launch(); CHECK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CHECK(cudaEventRecord(e0)); launch(); CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
        float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}
int main(int argc, char** argv) {
    FILE* f = argc > 1 ? fopen(argv[1], "w") : stdout; if (!f) f = stdout;
    cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount; void* out; CHECK(cudaMalloc(&out, (size_t)sms * 8 * 1024 * 8));
    const int iters = 1024;
    for (int bps : {1, 2, 4, 8}) {
        for (int threads : {128, 256}) {
            int blocks = sms * bps; double T = (double)blocks * threads;
            float a = time_it([&] { k_mul32<<<blocks, threads>>>((uint32_t*)out, iters); });
            float b = time_it([&] { k_mul29<<<blocks, threads>>>((uint32_t*)out, iters); });
            float c2 = time_it([&] { k_mul29x2<<<blocks, threads>>>((uint32_t*)out, iters); });
            fprintf(f, "{\"blocks_per_sm\": %d, \"threads\": %d, \"mul32_per_s\": %.4e, \"mul29_per_s\": %.4e, \"mul29x2_per_s\": %.4e}\n", bps, threads,
                    T * iters / (a * 1e-3), T * iters / (b * 1e-3), 2 * T * iters / (c2 * 1e-3));
            fflush(f);
        }
    }
    return 0;
}
