// Throughput and correctness probe for the FP64-pipe field product (csrc/fd.cuh) against the integer product
// (csrc/fq.cuh), alone and side by side on the same SM sub-partitions.
//   fd_bench check            -> compares fd results with fq_mul on random operands (exit code 1 on mismatch)
//   fd_bench time             -> JSON lines: products/s of imad-only, fd-only and hybrid warps
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "fd.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(2); } } while (0)

__global__ void k_check(const fq* a, const fq* b, fq* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fq x = a[i], y = b[i];
    fd X = fd_from_fq(x), Y = fd_from_fq(y);
    out[6 * i + 0] = fq_mul(x, y);
    out[6 * i + 1] = fd_to_fq(fd_mul(X, Y));
    out[6 * i + 2] = fd_to_fq(fd_mul(fd_from_fq_x16(x), Y));
    out[6 * i + 3] = fd_to_fq(fd_sqr(X));
    out[6 * i + 4] = fq_mul(x, x);
    // x*y - y*y + p*R', then minus x, plus 2p: (x*y - y*y - x) in the field
    uint64_t c[10];
    fd_cols_init<0, 1>(c);
    fd_cols_mac(c, X, Y);
    fd_cols_msub(c, Y, Y);
    fd_cols_redc(c);
    fd_cols_sub(c, X);
    fd_cols_addp<2>(c);
    out[6 * i + 5] = fd_to_fq(fd_cols_norm(c));
}

// mode 0: every warp runs integer products; 1: every warp FP64 products; 2: warps 0-3 of each CTA integer, 4-7 FP64
template <int CHAINS>
__global__ void __launch_bounds__(256, 2) k_time(const fq* a, fq* out, int iters, int mode, unsigned long long* counts) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int warp = threadIdx.x >> 5;
    const bool use_fd = mode == 1 || (mode == 2 && warp >= 4);
    fq x[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) x[k] = a[(tid + k * 7) & 1023];
    if (use_fd) {
        fd X[CHAINS], Y = fd_from_fq(a[(tid + 99) & 1023]);
#pragma unroll
        for (int k = 0; k < CHAINS; k++) X[k] = fd_from_fq(x[k]);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int k = 0; k < CHAINS; k++) X[k] = fd_mul(X[k], Y);
        }
#pragma unroll
        for (int k = 0; k < CHAINS; k++) x[k] = fd_to_fq(X[k]);
    } else {
        fq y = a[(tid + 99) & 1023];
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int k = 0; k < CHAINS; k++) x[k] = fq_mul(x[k], y);
        }
    }
    fq acc = x[0];
#pragma unroll
    for (int k = 1; k < CHAINS; k++) acc = fq_add(acc, x[k]);
    out[tid] = acc;
    if ((threadIdx.x & 31) == 0) atomicAdd(counts + (use_fd ? 1 : 0), (unsigned long long)iters * CHAINS * 32);
}

static uint64_t rnd_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { rnd_state ^= rnd_state << 13; rnd_state ^= rnd_state >> 7; rnd_state ^= rnd_state << 17; return rnd_state; }

int main(int argc, char** argv) {
    const char* cmd = argc > 1 ? argv[1] : "time";
    const int n = 1 << 16;
    std::vector<fq> ha(n), hb(n);
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < 8; k++) { ha[i].v[k] = (uint32_t)rnd(); hb[i].v[k] = (uint32_t)rnd(); }
        ha[i].v[7] &= 0x1fffffffu; hb[i].v[7] &= 0x1fffffffu;  // below p
    }
    // edge operands
    memset(&ha[0], 0, sizeof(fq)); memset(&hb[1], 0, sizeof(fq));
    const uint32_t pm1[8] = {0xd87cfd46u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    memcpy(&ha[2], pm1, 32); memcpy(&hb[2], pm1, 32); memcpy(&ha[3], pm1, 32);
    fq *da, *db, *dout;
    CK(cudaMalloc(&da, n * sizeof(fq))); CK(cudaMalloc(&db, n * sizeof(fq))); CK(cudaMalloc(&dout, (size_t)n * 6 * sizeof(fq)));
    CK(cudaMemcpy(da, ha.data(), n * sizeof(fq), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), n * sizeof(fq), cudaMemcpyHostToDevice));
    if (!strcmp(cmd, "check")) {
        k_check<<<n / 128, 128>>>(da, db, dout, n);
        CK(cudaDeviceSynchronize());
        std::vector<fq> ho((size_t)n * 6);
        CK(cudaMemcpy(ho.data(), dout, ho.size() * sizeof(fq), cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int i = 0; i < n; i++) {
            if (memcmp(&ho[6 * i], &ho[6 * i + 1], 32)) bad++;
            if (memcmp(&ho[6 * i], &ho[6 * i + 2], 32)) bad++;
            if (memcmp(&ho[6 * i + 3], &ho[6 * i + 4], 32)) bad++;
        }
        // dump a few (x, y, x*y - y*y - x) triples for the host-side big-int check
        FILE* f = fopen(argc > 2 ? argv[2] : "fd_check.bin", "wb");
        for (int i = 0; i < 256; i++) { fwrite(&ha[i], 32, 1, f); fwrite(&hb[i], 32, 1, f); fwrite(&ho[6 * i + 5], 32, 1, f); }
        fclose(f);
        printf("{\"check\": \"fd vs fq\", \"operands\": %d, \"mismatches\": %d}\n", n, bad);
        return bad ? 1 : 0;
    }
    int dev_sms = 0;
    CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0));
    unsigned long long* dcounts;
    CK(cudaMalloc(&dcounts, 16));
    const int iters = 2000;
    for (int chains = 1; chains <= 2; chains++)
        for (int mode = 0; mode < 3; mode++) {
            float best = 1e30f;
            unsigned long long hc[2];
            for (int rep = 0; rep < 3; rep++) {
                CK(cudaMemset(dcounts, 0, 16));
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0);
                if (chains == 1) k_time<1><<<dev_sms * 2, 256>>>(da, dout, iters, mode, dcounts);
                else k_time<2><<<dev_sms * 2, 256>>>(da, dout, iters, mode, dcounts);
                cudaEventRecord(e1);
                CK(cudaDeviceSynchronize());
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
                CK(cudaMemcpy(hc, dcounts, 16, cudaMemcpyDeviceToHost));
            }
            printf("{\"mode\": \"%s\", \"chains\": %d, \"ms\": %.3f, \"imad_products\": %llu, \"fd_products\": %llu, \"products_per_s\": %.4g}\n",
                   mode == 0 ? "imad" : mode == 1 ? "fd" : "hybrid", chains, best, hc[0], hc[1], (double)(hc[0] + hc[1]) / (best * 1e-3));
        }
    return 0;
}
