// pipe_bench4: issue cost of each SASS form that a 256-bit Montgomery product can be built from.
// Every kernel: 8 independent dependency chains per thread, forms verified with cuobjdump.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
#define REP4(X) X X X X
#define BODY(NAME, DECL, INIT, ASM, FIN)                                             \
    __global__ void NAME(uint32_t* out, int iters, uint32_t b, uint32_t c) {          \
        DECL; INIT;                                                                  \
        for (int it = 0; it < iters; it++) {                                         \
            _Pragma("unroll") for (int r = 0; r < 4; r++)                            \
            _Pragma("unroll") for (int k = 0; k < 8; k++) { ASM; }                   \
        }                                                                            \
        uint32_t s = 0; FIN; out[blockIdx.x * blockDim.x + threadIdx.x] = s;         \
    }
// A: wide multiply, no addend (a = low word of previous product keeps the chain alive)
BODY(k_A_mulwide_rz, uint64_t x[8], for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k + 1,
     asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mul.wide.u32 %0, lo, %1;}" : "+l"(x[k]) : "r"(b)),
     for (int k = 0; k < 8; k++) s ^= (uint32_t)x[k] ^ (uint32_t)(x[k] >> 32))
// B: wide multiply-accumulate with a live 64-bit addend
BODY(k_B_madwide_acc, uint64_t x[8], for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k + 1,
     asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(x[k]) : "r"(b)),
     for (int k = 0; k < 8; k++) s ^= (uint32_t)x[k] ^ (uint32_t)(x[k] >> 32))
// C: 32-bit multiply-add low
BODY(k_C_mad_lo, uint32_t x[8], for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k + 1,
     asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(b), "r"(c)),
     for (int k = 0; k < 8; k++) s ^= x[k])
// D: 32-bit multiply-add high
BODY(k_D_mad_hi, uint32_t x[8], for (int k = 0; k < 8; k++) x[k] = threadIdx.x * 0x9e3779b9u + k + 1,
     asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(b), "r"(c)),
     for (int k = 0; k < 8; k++) s ^= x[k])
// E: three-input add
BODY(k_E_iadd3, uint32_t x[8], for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k + 1,
     asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x[k]) : "r"(b), "r"(c)),
     for (int k = 0; k < 8; k++) s ^= x[k])
// F: 64-bit add = add.cc + addc (IADD3 P-out, IADD3.X P-in)
BODY(k_F_add64, uint64_t x[8], for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k + 1,
     asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; add.cc.u32 lo, lo, %1; addc.u32 hi, hi, %2; mov.b64 %0, {lo, hi};}" : "+l"(x[k]) : "r"(b), "r"(c)),
     for (int k = 0; k < 8; k++) s ^= (uint32_t)x[k] ^ (uint32_t)(x[k] >> 32))
// G: carry chain of four (add.cc, addc.cc, addc.cc, addc): 128-bit add
BODY(k_G_add128, uint32_t x[8][4], for (int k = 0; k < 8; k++) for (int j = 0; j < 4; j++) x[k][j] = threadIdx.x + k + j,
     asm volatile("add.cc.u32 %0, %0, %4; addc.cc.u32 %1, %1, %5; addc.cc.u32 %2, %2, %4; addc.u32 %3, %3, %5;" : "+r"(x[k][0]), "+r"(x[k][1]), "+r"(x[k][2]), "+r"(x[k][3]) : "r"(b), "r"(c)),
     for (int k = 0; k < 8; k++) for (int j = 0; j < 4; j++) s ^= x[k][j])
// H: wide MAC with carry-out only, carry consumed by addc (IMAD.WIDE P-out + IADD3.X)
BODY(k_H_madwide_cout, uint32_t x[8][3], for (int k = 0; k < 8; k++) for (int j = 0; j < 3; j++) x[k][j] = threadIdx.x + k + j,
     asm volatile("mad.lo.cc.u32 %0, %0, %3, %0; madc.hi.cc.u32 %1, %0, %3, %1; addc.u32 %2, %2, 0;" : "+r"(x[k][0]), "+r"(x[k][1]), "+r"(x[k][2]) : "r"(b)),
     for (int k = 0; k < 8; k++) for (int j = 0; j < 3; j++) s ^= x[k][j])
// I: shift-right funnel (SHF) + mask (LOP3): the carry extraction of an unsaturated-limb design
BODY(k_I_shf_lop, uint32_t x[8], for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k + 1,
     asm volatile("{.reg .u32 t; shf.r.wrap.b32 t, %0, %1, 29; and.b32 %0, t, 0x1fffffff; add.u32 %0, %0, %2;}" : "+r"(x[k]) : "r"(b), "r"(c)),
     for (int k = 0; k < 8; k++) s ^= x[k])
// J: FP64 FMA
__global__ void k_J_dfma(uint32_t* out, int iters, uint32_t b, uint32_t c) {
    double x[8]; double bb = 1.0 + b * 1e-9, cc = c * 1e-9;
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[k]) : "d"(bb), "d"(cc));
    }
    double s = 0; for (int k = 0; k < 8; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)s;
}
// K: A and E interleaved 1:1 (FMA-heavy pipe + ALU pipe)
BODY(k_K_mulwide_plus_iadd3, uint64_t x[8]; uint32_t z[8], for (int k = 0; k < 8; k++) { x[k] = threadIdx.x + k + 1; z[k] = k; },
     asm volatile("{.reg .u32 lo, hi, t; mov.b64 {lo, hi}, %0; mul.wide.u32 %0, lo, %2; add.u32 t, %1, %2; add.u32 %1, t, %3;}" : "+l"(x[k]), "+r"(z[k]) : "r"(b), "r"(c)),
     for (int k = 0; k < 8; k++) s ^= (uint32_t)x[k] ^ z[k])
// L: A + two-carry 3-input 64-bit accumulate: the "multiply on FMA pipe, accumulate on ALU pipe" MAC
BODY(k_L_mulwide_add64, uint64_t x[8]; uint32_t a0[8]; uint32_t a1[8]; uint32_t a2[8],
     for (int k = 0; k < 8; k++) { x[k] = threadIdx.x + k + 1; a0[k] = k; a1[k] = k + 1; a2[k] = 0; },
     asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mul.wide.u32 %0, lo, %4; mov.b64 {lo, hi}, %0; add.cc.u32 %1, %1, lo; addc.cc.u32 %2, %2, hi; addc.u32 %3, %3, 0;}" : "+l"(x[k]), "+r"(a0[k]), "+r"(a1[k]), "+r"(a2[k]) : "r"(b)),
     for (int k = 0; k < 8; k++) s ^= (uint32_t)x[k] ^ a0[k] ^ a1[k] ^ a2[k])

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
    const int iters = 4096; int blocks = sms * 4, threads = 256;
    double T = (double)blocks * threads * 32.0 * iters;
    // cycles per warp-instruction-group per SMSP, assuming the SM clock printed by pipe_bench (1.92 GHz)
    auto rep = [&](const char* name, float ms, double groups_scale) {
        double per_s = T * groups_scale / (ms * 1e-3);
        double cyc = 1.0 / (per_s / 32.0 / (sms * 4) / 1.92e9);
        fprintf(f, "{\"test\": \"%s\", \"ms\": %.4f, \"ops_per_s\": %.4e, \"cycles_per_warp_op_per_smsp@1.92GHz\": %.2f}\n", name, ms, per_s, cyc); fflush(f);
    };
    rep("A_mulwide_rz", time_it([&] { k_A_mulwide_rz<<<blocks, threads>>>((uint32_t*)out, iters, 7, 9); }), 1);
    rep("B_madwide_acc64", time_it([&] { k_B_madwide_acc<<<blocks, threads>>>((uint32_t*)out, iters, 7, 9); }), 1);
    rep("C_mad_lo", time_it([&] { k_C_mad_lo<<<blocks, threads>>>((uint32_t*)out, iters, 7, 9); }), 1);
    rep("D_mad_hi", time_it([&] { k_D_mad_hi<<<blocks, threads>>>((uint32_t*)out, iters, 7, 9); }), 1);
    rep("E_iadd3", time_it([&] { k_E_iadd3<<<blocks, threads>>>((uint32_t*)out, iters, 7, 9); }), 1);
    rep("F_add64(2 instr)", time_it([&] { k_F_add64<<<blocks, threads>>>((uint32_t*)out, iters, 7, 9); }), 1);
    rep("G_add128(4 instr)", time_it([&] { k_G_add128<<<blocks, threads>>>((uint32_t*)out, iters, 7, 9); }), 1);
    rep("H_madwide_cout+addc", time_it([&] { k_H_madwide_cout<<<blocks, threads>>>((uint32_t*)out, iters, 7, 9); }), 1);
    rep("I_shf+lop+add(3 instr)", time_it([&] { k_I_shf_lop<<<blocks, threads>>>((uint32_t*)out, iters, 7, 9); }), 1);
    rep("J_dfma", time_it([&] { k_J_dfma<<<blocks, threads>>>((uint32_t*)out, iters, 7, 9); }), 1);
    rep("K_mulwide+iadd3", time_it([&] { k_K_mulwide_plus_iadd3<<<blocks, threads>>>((uint32_t*)out, iters, 7, 9); }), 1);
    rep("L_mulwide+add64c(4 instr)", time_it([&] { k_L_mulwide_add64<<<blocks, threads>>>((uint32_t*)out, iters, 7, 9); }), 1);
    return 0;
}
