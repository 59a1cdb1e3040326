// Register-operand forms: DFMA with three distinct register operands, IADD3 pairs with three register operands and two
// carry-outs (the 3-input 64-bit add of fd.cuh), alone and together.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int ND, int NA>
__global__ void k(const double* in, double* out, int iters) {
    double d[8], a[8], b[8];
    uint64_t u[4], v[4], w[4];
#pragma unroll
    for (int i = 0; i < 8; i++) { d[i] = in[threadIdx.x + i]; a[i] = in[threadIdx.x + 8 + i]; b[i] = in[threadIdx.x + 16 + i]; }
#pragma unroll
    for (int i = 0; i < 4; i++) { u[i] = (uint64_t)in[i + 32]; v[i] = (uint64_t)in[threadIdx.x + 40 + i]; w[i] = (uint64_t)in[threadIdx.x + 50 + i]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < ND; i++) d[(r * ND + i) & 7] = __fma_rz(a[(r * ND + i) & 7], b[(r * ND + i + 3) & 7], d[(r * ND + i) & 7]);
#pragma unroll
            for (int i = 0; i < NA; i++) {
                const int c = (r * NA + i) & 3;
                u[c] = u[c] + v[c] + w[(c + 1) & 3];
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += d[i];
#pragma unroll
    for (int i = 0; i < 4; i++) s += (double)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ND, int NA>
void run(const double* in, double* out, int sms, double ghz) {
    for (int warps_per_smsp = 4; warps_per_smsp <= 8; warps_per_smsp *= 2) {
        const int iters = 4000, threads = 128, ctas = sms * warps_per_smsp;
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a);
            k<ND, NA><<<ctas, threads>>>(in, out, iters);
            cudaEventRecord(b); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        const double steps = (double)iters * 4 * warps_per_smsp;
        printf("{\"dfma_rrr\": %d, \"add64_3input_pairs\": %d, \"warps_per_smsp\": %d, \"cycles_per_step\": %.2f}\n", ND, NA, warps_per_smsp,
               best * 1e-3 * ghz * 1e9 / steps);
    }
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz * 1e-6;
    double *in, *out; cudaMalloc(&in, 4096 * 8); cudaMemset(in, 0, 4096 * 8); cudaMalloc(&out, (size_t)sms * 16 * 128 * 8);
    run<1, 0>(in, out, sms, ghz); run<2, 0>(in, out, sms, ghz); run<0, 1>(in, out, sms, ghz); run<0, 2>(in, out, sms, ghz);
    run<1, 1>(in, out, sms, ghz); run<2, 1>(in, out, sms, ghz); run<3, 1>(in, out, sms, ghz); run<3, 2>(in, out, sms, ghz); run<4, 2>(in, out, sms, ghz);
    return 0;
}
