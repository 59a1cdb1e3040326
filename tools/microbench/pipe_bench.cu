// pipe_bench: measures the integer-multiply roofline of the GPU it runs on, the denominator of
// every `roofline.frac` this repo reports for the point-arithmetic kernels (the
// IMAD peak "is not in MEASURED_PEAKS.json ... measure it, do not trust a datasheet number").
//
// Prints one JSON object per test:  ops/s, ops/clk/SM (using the SM clock observed through
// clock64 during the run), for
//   imad_lo        32-bit IMAD (mad.lo.u32), 8 independent chains / thread
//   imad_wide      IMAD.WIDE.U32 (mad.wide.u32), 8 independent chains / thread
//   imad_wide_imm  same with an immediate multiplier (the m*p half of the Montgomery product)
//   iadd3          3-input integer add (ALU pipe), for the carry-fixup budget
//   fq_mul         the production Montgomery multiply, 1 and 2 independent chains / thread
//   xyzz_madd      the production mixed addition with operands in registers (no memory): the
//                  compute ceiling of k_accumulate
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "g1.cuh"

#define CHECK(x)                                                                       \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                    \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

__global__ void k_imad_lo(uint32_t* out, int iters, uint32_t a, uint32_t b, long long* clk) {
    uint32_t x[8];
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

__global__ void k_imad_wide(uint64_t* out, int iters, uint32_t a, uint32_t b, long long* clk) {
    uint64_t x[8];
    uint32_t m[8];
    for (int k = 0; k < 8; k++) { x[k] = threadIdx.x + k; m[k] = a + k * b + threadIdx.x; }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[k]) : "r"(m[k]), "r"(b));
    }
    long long t1 = clock64();
    uint64_t s = 0;
    for (int k = 0; k < 8; k++) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

__global__ void k_imad_wide_imm(uint64_t* out, int iters, uint32_t a, uint32_t b, long long* clk) {
    uint64_t x[8];
    uint32_t m[8];
    for (int k = 0; k < 8; k++) { x[k] = threadIdx.x + k; m[k] = a + k * b + threadIdx.x; }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, 0x3c208c16, %0;" : "+l"(x[k]) : "r"(m[k]));
    }
    long long t1 = clock64();
    uint64_t s = 0;
    for (int k = 0; k < 8; k++) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

__global__ void k_iadd3(uint32_t* out, int iters, uint32_t a, uint32_t b, long long* clk) {
    uint32_t x[8];
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x[k]) : "r"(a), "r"(b));
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int CHAINS>
__global__ void k_fq_mul(uint32_t* out, int iters, long long* clk) {
    fq x[CHAINS], y;
    for (int c = 0; c < CHAINS; c++)
        for (int k = 0; k < 8; k++) x[c].v[k] = threadIdx.x * 977u + k * 13u + c;
    for (int k = 0; k < 8; k++) y.v[k] = blockIdx.x * 31u + k + 5u;
    x[0].v[7] &= 0x0fffffffu; y.v[7] &= 0x0fffffffu;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) x[c] = fq_mul_inline(x[c], y);
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int c = 0; c < CHAINS; c++)
        for (int k = 0; k < 8; k++) s ^= x[c].v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

__global__ void k_fq_sqr(uint32_t* out, int iters, long long* clk) {
    fq x;
    for (int k = 0; k < 8; k++) x.v[k] = threadIdx.x * 977u + k * 13u + blockIdx.x;
    x.v[7] &= 0x0fffffffu;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) x = fq_sqr(x);
    long long t1 = clock64();
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s ^= x.v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

__global__ void __launch_bounds__(128) k_madd(uint32_t* out, int iters, long long* clk) {
    // acc = G, then acc += P repeatedly with P = 2G (in registers).  Never hits the special cases.
    affine_t G;
    G.x = fq_one();
    G.y = fq_dbl(fq_one());
    xyzz_t acc = xyzz_dbl_affine(G);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) xyzz_madd(acc, G);
    long long t1 = clock64();
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s ^= acc.x.v[k] ^ acc.y.v[k] ^ acc.zz.v[k] ^ acc.zzz.v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

__global__ void __launch_bounds__(128) k_madd_nc(uint32_t* out, int iters, long long* clk) {
    affine_t G;
    G.x = fq_one();
    G.y = fq_dbl(fq_one());
    xyzz_t acc = xyzz_dbl_affine(G);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) xyzz_madd(acc, G);
    long long t1 = clock64();
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s ^= acc.x.v[k] ^ acc.y.v[k] ^ acc.zz.v[k] ^ acc.zzz.v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

struct Result {
    double ms;
    long long clk;
};

template <typename F>
Result time_it(F launch, long long* d_clk) {
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    launch();
    launch();
    CHECK(cudaDeviceSynchronize());
    float best = 1e30f;
    long long clk = 0;
    for (int rep = 0; rep < 5; rep++) {
        CHECK(cudaEventRecord(e0));
        launch();
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        float ms;
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) {
            best = ms;
            CHECK(cudaMemcpy(&clk, d_clk, 8, cudaMemcpyDeviceToHost));
        }
    }
    return {best, clk};
}

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : nullptr;
    FILE* f = path ? fopen(path, "w") : stdout;
    if (!f) f = stdout;
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    void* out;
    long long* d_clk;
    CHECK(cudaMalloc(&out, (size_t)sms * 16 * 1024 * 8));
    CHECK(cudaMalloc(&d_clk, 8));
    fprintf(f, "{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    const int iters = 4096;
    auto report = [&](const char* name, Result r, double ops_per_thread, int blocks, int threads, const char* unit) {
        double total = ops_per_thread * blocks * threads;
        double per_s = total / (r.ms * 1e-3);
        // SM clock during the run: block 0's cycle count over (approximately) the kernel time
        double mhz = r.clk / (r.ms * 1e-3) / 1e6;
        double per_clk_sm = total / sms / (double)r.clk;
        fprintf(f, "{\"test\": \"%s\", \"blocks_per_sm\": %d, \"threads\": %d, \"ms\": %.4f, \"%s_per_s\": %.4e, \"per_clk_per_sm\": %.2f, \"approx_sm_mhz\": %.0f}\n",
                name, blocks / sms, threads, r.ms, unit, per_s, per_clk_sm, mhz);
        fflush(f);
    };
    for (int bps : {1, 2, 4}) {
        int blocks = sms * bps, threads = 256;
        report("imad_lo", time_it([&] { k_imad_lo<<<blocks, threads>>>((uint32_t*)out, iters, 3, 7, d_clk); }, d_clk), 32.0 * iters, blocks, threads, "imad");
        report("imad_wide", time_it([&] { k_imad_wide<<<blocks, threads>>>((uint64_t*)out, iters, 3, 7, d_clk); }, d_clk), 32.0 * iters, blocks, threads, "imad");
        report("imad_wide_imm", time_it([&] { k_imad_wide_imm<<<blocks, threads>>>((uint64_t*)out, iters, 3, 7, d_clk); }, d_clk), 32.0 * iters, blocks, threads, "imad");
        report("iadd3", time_it([&] { k_iadd3<<<blocks, threads>>>((uint32_t*)out, iters, 3, 7, d_clk); }, d_clk), 64.0 * iters, blocks, threads, "iadd");
        report("fq_mul_x1", time_it([&] { k_fq_mul<1><<<blocks, threads>>>((uint32_t*)out, 1024, d_clk); }, d_clk), 1024.0, blocks, threads, "fqmul");
        report("fq_sqr", time_it([&] { k_fq_sqr<<<blocks, threads>>>((uint32_t*)out, 1024, d_clk); }, d_clk), 1024.0, blocks, threads, "fqsqr");
        report("fq_mul_x2", time_it([&] { k_fq_mul<2><<<blocks, threads>>>((uint32_t*)out, 1024, d_clk); }, d_clk), 2048.0, blocks, threads, "fqmul");
    }
    for (int bps : {1, 2, 3, 4}) {
        int blocks = sms * bps, threads = 128;
        report("xyzz_madd", time_it([&] { k_madd<<<blocks, threads>>>((uint32_t*)out, 512, d_clk); }, d_clk), 512.0, blocks, threads, "madd");
        report("xyzz_madd_nc", time_it([&] { k_madd_nc<<<blocks, threads>>>((uint32_t*)out, 512, d_clk); }, d_clk), 512.0, blocks, threads, "madd");
    }
    if (f != stdout) fclose(f);
    return 0;
}
