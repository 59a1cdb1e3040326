// Issue-overlap probe for the FP64 pipe against the integer ALU and the IMAD pipe on sm_100a:
// ND double FMAs, NI 32-bit IADD3 and NM IMAD.WIDE per unrolled step, all independent chains.
// Prints cycles per step per warp per SM sub-partition at several occupancies.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int ND, int NI, int NM>
__global__ void k(double* out, double x, double y, uint32_t v, uint32_t w, int iters) {
    double d[8];
    uint32_t u[8];
    uint64_t m[4];
#pragma unroll
    for (int i = 0; i < 8; i++) { d[i] = threadIdx.x + i; u[i] = threadIdx.x * 3 + i; }
#pragma unroll
    for (int i = 0; i < 4; i++) m[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < ND; i++) d[(r * ND + i) & 7] = __fma_rz(d[(r * ND + i) & 7], x, y);
#pragma unroll
            for (int i = 0; i < NI; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[(r * NI + i) & 7]) : "r"(v));
#pragma unroll
            for (int i = 0; i < NM; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(m[(r * NM + i) & 3]) : "r"(v), "r"(w));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += d[i] + u[i];
#pragma unroll
    for (int i = 0; i < 4; i++) s += (double)m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ND, int NI, int NM>
void run(double* out, int sms, double ghz) {
    for (int warps_per_smsp = 2; warps_per_smsp <= 16; warps_per_smsp *= 2) {
        const int iters = 4000, threads = 128, ctas = sms * warps_per_smsp;  // 4 warps per CTA, one per sub-partition
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a);
            k<ND, NI, NM><<<ctas, threads>>>(out, 1.0000001, 3.0, 12345u, 777u, iters);
            cudaEventRecord(b); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        const double steps = (double)iters * 4 * warps_per_smsp;  // per sub-partition
        printf("{\"dfma\": %d, \"iadd\": %d, \"imad_wide\": %d, \"warps_per_smsp\": %d, \"cycles_per_step\": %.2f}\n", ND, NI, NM,
               warps_per_smsp, best * 1e-3 * ghz * 1e9 / steps);
    }
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz * 1e-6;
    double* out; cudaMalloc(&out, (size_t)sms * 16 * 128 * 8);
    run<1, 0, 0>(out, sms, ghz); run<0, 1, 0>(out, sms, ghz); run<0, 0, 1>(out, sms, ghz);
    run<1, 1, 0>(out, sms, ghz); run<1, 2, 0>(out, sms, ghz); run<2, 2, 0>(out, sms, ghz);
    run<1, 0, 1>(out, sms, ghz); run<2, 0, 1>(out, sms, ghz); run<1, 1, 1>(out, sms, ghz); run<2, 2, 1>(out, sms, ghz);
    run<0, 2, 1>(out, sms, ghz);
    return 0;
}
