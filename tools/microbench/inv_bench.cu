// inv_bench: price of one Fq inversion on B200 in units of the production field multiplication, stand-alone and
// hidden under multiply work (the situation inside the batched-affine accumulation: one inversion per lane per batch of
// k additions, 6 multiplications each).  Prints one JSON object per test.
//   fermat      fq_inv     a^(p-2), ~380 multiplications on the integer-multiply pipe
//   safegcd     fq_inv_by  Bernstein-Yang divsteps, mostly ALU-pipe work
//   mul         fq_mul chain (the unit)
//   mix_k<K>    per thread and iteration: one safegcd inversion + 6*K multiplications (two independent chains)
//   mulonly_k<K> the same without the inversion
#include <cstdio>
#include <cstdlib>

#include "g1.cuh"
#include "fq_inv.cuh"

#define CHECK(x)                                                                    \
    do {                                                                            \
        cudaError_t e = (x);                                                        \
        if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } \
    } while (0)

__device__ __forceinline__ fq seed_fq(uint32_t s) {
    fq x;
    for (int k = 0; k < 8; k++) x.v[k] = s * 977u + k * 0x9e3779b9u + 17u;
    x.v[7] &= 0x0fffffffu;
    return x;
}
__global__ void __launch_bounds__(128) k_fermat(uint32_t* out, int iters) {
    fq x = seed_fq(blockIdx.x * blockDim.x + threadIdx.x);
#pragma unroll 1
    for (int it = 0; it < iters; it++) x = fq_add(fq_inv(x), fq_one());
    out[blockIdx.x * blockDim.x + threadIdx.x] = x.v[0] ^ x.v[7];
}
__global__ void __launch_bounds__(128) k_safegcd(uint32_t* out, int iters) {
    fq x = seed_fq(blockIdx.x * blockDim.x + threadIdx.x);
#pragma unroll 1
    for (int it = 0; it < iters; it++) x = fq_add(fq_inv_by(x), fq_one());
    out[blockIdx.x * blockDim.x + threadIdx.x] = x.v[0] ^ x.v[7];
}
__global__ void __launch_bounds__(128) k_mul(uint32_t* out, int iters) {
    fq x = seed_fq(blockIdx.x * blockDim.x + threadIdx.x), y = seed_fq(threadIdx.x + 99), z = seed_fq(threadIdx.x + 7);
#pragma unroll 1
    for (int it = 0; it < iters; it++) { x = fq_mul(x, y); z = fq_mul(z, y); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x.v[0] ^ z.v[7];
}
template <int K, bool INV>
__global__ void __launch_bounds__(128) k_mix(uint32_t* out, int iters) {
    fq x = seed_fq(blockIdx.x * blockDim.x + threadIdx.x), y = seed_fq(threadIdx.x + 99), z = seed_fq(threadIdx.x + 7);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (INV) x = fq_inv_by(x);
#pragma unroll 1
        for (int k = 0; k < 3 * K; k++) { x = fq_mul(x, y); z = fq_mul(z, x); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x.v[0] ^ z.v[7];
}

template <typename F>
float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    CHECK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        CHECK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    FILE* f = argc > 1 ? fopen(argv[1], "w") : stdout;
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    uint32_t* out;
    CHECK(cudaMalloc(&out, (size_t)sms * 8 * 128 * 4));
    for (int bps : {2, 4}) {
        const int blocks = sms * bps, threads = 128;
        const double lanes = (double)blocks * threads;
        float ms = time_ms([&] { k_mul<<<blocks, threads>>>(out, 2048); });
        const double mul_rate = lanes * 4096 / (ms * 1e-3);
        fprintf(f, "{\"test\": \"mul\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"fqmul_per_s\": %.4e}\n", bps, ms, mul_rate);
        ms = time_ms([&] { k_fermat<<<blocks, threads>>>(out, 8); });
        double rate = lanes * 8 / (ms * 1e-3);
        fprintf(f, "{\"test\": \"fermat\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"inv_per_s\": %.4e, \"fqmul_equivalents\": %.1f}\n", bps, ms, rate, mul_rate / rate);
        ms = time_ms([&] { k_safegcd<<<blocks, threads>>>(out, 32); });
        rate = lanes * 32 / (ms * 1e-3);
        fprintf(f, "{\"test\": \"safegcd\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"inv_per_s\": %.4e, \"fqmul_equivalents\": %.1f}\n", bps, ms, rate, mul_rate / rate);
#define MIX(K)                                                                                                           \
    {                                                                                                                    \
        float m1 = time_ms([&] { k_mix<K, true><<<blocks, threads>>>(out, 16); });                                       \
        float m0 = time_ms([&] { k_mix<K, false><<<blocks, threads>>>(out, 16); });                                      \
        fprintf(f, "{\"test\": \"mix_k%d\", \"blocks_per_sm\": %d, \"ms_with_inv\": %.4f, \"ms_mul_only\": %.4f, "       \
                   "\"muls_per_add_equiv\": %.3f, \"hidden_inv_fqmul_equivalents\": %.1f}\n",                            \
                K, bps, m1, m0, 6.0 * m1 / m0, (m1 - m0) / m0 * 6.0 * K);                                                \
    }
        MIX(16) MIX(32) MIX(64) MIX(128)
        fflush(f);
    }
    return 0;
}
