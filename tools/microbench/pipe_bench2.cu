// pipe_bench2: what limits the IMAD.WIDE.U32.X carry chains?  Candidate instruction mixes for the
// 256-bit Montgomery product, each measured as 32x32->64 multiply-accumulates per second.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cstdint>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

// (a) rows of 4 carry-linked wide MACs (+ addc), 4 independent rows per iteration: the shape inside fq_mul
__global__ void k_wide_x_chain(uint32_t* out, int iters, uint32_t b0) {
    uint32_t x[4][9], a[8];
    for (int r = 0; r < 4; r++) for (int k = 0; k < 9; k++) x[r][k] = threadIdx.x + r * 9 + k;
    for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 77 + k;
    uint32_t b = b0 + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            asm volatile(
                "mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
                "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
                "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
                "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\t"
                "addc.u32 %8, %8, 0;"
                : "+r"(x[r][0]), "+r"(x[r][1]), "+r"(x[r][2]), "+r"(x[r][3]), "+r"(x[r][4]), "+r"(x[r][5]), "+r"(x[r][6]), "+r"(x[r][7]), "+r"(x[r][8])
                : "r"(a[r]), "r"(a[r + 1]), "r"(a[r + 2]), "r"(a[r + 3]), "r"(b));
        }
    }
    uint32_t s = 0;
    for (int r = 0; r < 4; r++) for (int k = 0; k < 9; k++) s ^= x[r][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// (b) carry-out only: each wide MAC starts its own chain; the carry is consumed by an addc
__global__ void k_wide_cout(uint32_t* out, int iters, uint32_t b0) {
    uint32_t x[8][3], a[8];
    for (int r = 0; r < 8; r++) for (int k = 0; k < 3; k++) x[r][k] = threadIdx.x + r * 3 + k;
    for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 77 + k;
    uint32_t b = b0 + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                         : "+r"(x[r][0]), "+r"(x[r][1]), "+r"(x[r][2]) : "r"(a[r]), "r"(b));
        }
    }
    uint32_t s = 0;
    for (int r = 0; r < 8; r++) for (int k = 0; k < 3; k++) s ^= x[r][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// (c) plain wide multiply on the FMA pipe + 64-bit accumulate with add.cc/addc on the ALU pipe
__global__ void k_mulwide_iadd(uint32_t* out, int iters, uint32_t b0) {
    uint32_t x[8][3], a[8];
    for (int r = 0; r < 8; r++) for (int k = 0; k < 3; k++) x[r][k] = threadIdx.x + r * 3 + k;
    for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 77 + k;
    uint32_t b = b0 + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            asm volatile("{.reg .u64 p; .reg .u32 pl, ph;\n\tmul.wide.u32 p, %1, %4;\n\tmov.b64 {pl, ph}, p;\n\t"
                         "add.cc.u32 %0, %0, pl;\n\taddc.cc.u32 %1, %1, ph;\n\taddc.u32 %2, %2, 0;}"
                         : "+r"(x[r][0]), "+r"(x[r][1]), "+r"(x[r][2]) : "r"(a[r]), "r"(b));
        }
    }
    uint32_t s = 0;
    for (int r = 0; r < 8; r++) for (int k = 0; k < 3; k++) s ^= x[r][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// (d) 64-bit accumulators without carries (mad.wide.u32 into 8 independent u64): the ceiling
__global__ void k_wide_plain(uint64_t* out, int iters, uint32_t b0) {
    uint64_t x[8]; uint32_t a[8];
    for (int k = 0; k < 8; k++) { x[k] = threadIdx.x + k; a[k] = threadIdx.x * 77 + k; }
    uint32_t b = b0 + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[k]) : "r"(a[k]), "r"(b));
    }
    uint64_t s = 0;
    for (int k = 0; k < 8; k++) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// (e) DFMA rate, and (f) DFMA interleaved with IMAD.WIDE: do the FP64 and FMA pipes overlap?
__global__ void k_dfma(double* out, int iters, double b) {
    double x[8];
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(x[k]) : "d"(b));
    }
    double s = 0;
    for (int k = 0; k < 8; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma_imad(double* out, int iters, double b, uint32_t bi) {
    double x[8]; uint64_t y[8]; uint32_t a[8];
    for (int k = 0; k < 8; k++) { x[k] = threadIdx.x + k; y[k] = threadIdx.x + k; a[k] = threadIdx.x * 77 + k; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(x[k]) : "d"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y[k]) : "r"(a[k]), "r"(bi));
            }
    }
    double s = 0;
    for (int k = 0; k < 8; k++) s += x[k] + (double)y[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// (g) IMAD.WIDE interleaved 1:1 with IADD3 (different pipes): total issue rate
__global__ void k_imad_iadd(uint64_t* out, int iters, uint32_t bi) {
    uint64_t y[8]; uint32_t a[8], z[8];
    for (int k = 0; k < 8; k++) { y[k] = threadIdx.x + k; a[k] = threadIdx.x * 77 + k; z[k] = k; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y[k]) : "r"(a[k]), "r"(bi));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(z[k]) : "r"(bi));
            }
    }
    uint64_t s = 0;
    for (int k = 0; k < 8; k++) s ^= y[k] ^ z[k];
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
    const int iters = 4096;
    for (int bps : {1, 2, 4}) {
        int blocks = sms * bps, threads = 256; double T = (double)blocks * threads;
        auto rep = [&](const char* name, float ms, double ops, const char* what) {
            fprintf(f, "{\"test\": \"%s\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"%s_per_s\": %.4e}\n", name, bps, ms, what, ops * T / (ms * 1e-3)); fflush(f);
        };
        rep("wide_x_chain4", time_it([&] { k_wide_x_chain<<<blocks, threads>>>((uint32_t*)out, iters, 7); }), 16.0 * iters, "mac");
        rep("wide_cout_only", time_it([&] { k_wide_cout<<<blocks, threads>>>((uint32_t*)out, iters, 7); }), 8.0 * iters, "mac");
        rep("mulwide_iadd", time_it([&] { k_mulwide_iadd<<<blocks, threads>>>((uint32_t*)out, iters, 7); }), 8.0 * iters, "mac");
        rep("wide_plain", time_it([&] { k_wide_plain<<<blocks, threads>>>((uint64_t*)out, iters, 7); }), 32.0 * iters, "mac");
        rep("dfma", time_it([&] { k_dfma<<<blocks, threads>>>((double*)out, iters, 1.000001); }), 32.0 * iters, "dfma");
        rep("dfma_plus_imad", time_it([&] { k_dfma_imad<<<blocks, threads>>>((double*)out, iters, 1.000001, 7); }), 32.0 * iters, "pair");
        rep("imad_plus_iadd", time_it([&] { k_imad_iadd<<<blocks, threads>>>((uint64_t*)out, iters, 7); }), 32.0 * iters, "pair");
    }
    return 0;
}
