// FP64 pipe vs ALU pipe issue overlap on sm_100a, with instructions ptxas cannot move to another pipe:
// DFMA (fp64), LOP3 (alu), SHF (alu), IADD3 with carry-out + IADD3.X (alu), IMAD.WIDE.U32.X carry chain (fmaheavy).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int ND, int NL, int NC, int NM>
__global__ void k(double* out, double x, double y, uint32_t v, uint32_t w, int iters) {
    double d[8];
    uint32_t u[8], lo[4], hi[4];
    uint32_t m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { d[i] = threadIdx.x + i; u[i] = threadIdx.x * 3 + i; m[i] = threadIdx.x + i; }
#pragma unroll
    for (int i = 0; i < 4; i++) { lo[i] = threadIdx.x + i; hi[i] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < ND; i++) d[(r * ND + i) & 7] = __fma_rz(d[(r * ND + i) & 7], x, y);
#pragma unroll
            for (int i = 0; i < NL; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[(r * NL + i) & 7]) : "r"(v), "r"(w));
#pragma unroll
            for (int i = 0; i < NC; i++)
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(lo[(r * NC + i) & 3]), "+r"(hi[(r * NC + i) & 3]) : "r"(v), "r"(w));
            if (NM) {  // one 8-limb carry chain row: 4 mad.lo.cc/madc.hi.cc pairs -> 4 IMAD.WIDE.U32.X
                asm volatile(
                    "mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1;"
                    "madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3;"
                    "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5;"
                    "madc.lo.cc.u32 %6, %8, %10, %6; madc.hi.u32 %7, %8, %10, %7;"
                    : "+r"(m[0]), "+r"(m[1]), "+r"(m[2]), "+r"(m[3]), "+r"(m[4]), "+r"(m[5]), "+r"(m[6]), "+r"(m[7])
                    : "r"(v + r), "r"(w), "r"(v));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += d[i] + u[i] + m[i];
#pragma unroll
    for (int i = 0; i < 4; i++) s += lo[i] + hi[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ND, int NL, int NC, int NM>
void run(double* out, int sms, double ghz) {
    for (int warps_per_smsp = 4; warps_per_smsp <= 16; warps_per_smsp *= 2) {
        const int iters = 4000, threads = 128, ctas = sms * warps_per_smsp;
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a);
            k<ND, NL, NC, NM><<<ctas, threads>>>(out, 1.0000001, 3.0, 12345u, 777u, iters);
            cudaEventRecord(b); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        const double steps = (double)iters * 4 * warps_per_smsp;
        printf("{\"dfma\": %d, \"lop3\": %d, \"add64_pairs\": %d, \"imad_wide_x4\": %d, \"warps_per_smsp\": %d, \"cycles_per_step\": %.2f}\n", ND, NL, NC, NM,
               warps_per_smsp, best * 1e-3 * ghz * 1e9 / steps);
    }
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz * 1e-6;
    double* out; cudaMalloc(&out, (size_t)sms * 16 * 128 * 8);
    run<1, 0, 0, 0>(out, sms, ghz); run<0, 1, 0, 0>(out, sms, ghz); run<0, 0, 1, 0>(out, sms, ghz); run<0, 0, 0, 1>(out, sms, ghz);
    run<1, 1, 0, 0>(out, sms, ghz); run<2, 2, 0, 0>(out, sms, ghz); run<1, 2, 0, 0>(out, sms, ghz); run<1, 0, 1, 0>(out, sms, ghz);
    run<2, 0, 0, 1>(out, sms, ghz); run<4, 0, 0, 1>(out, sms, ghz); run<4, 4, 0, 1>(out, sms, ghz); run<6, 5, 0, 1>(out, sms, ghz);
    run<0, 4, 0, 1>(out, sms, ghz);
    return 0;
}
