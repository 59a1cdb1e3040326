// Example user of the device math library (include/b200math.cuh): a kernel of its own that walks k -> k*G for
// k = 1..N with the library's XYZZ arithmetic, normalises every point with the safegcd inversion and checks the curve
// equation y^2 = x^3 + 3 with the field ops -- nothing here goes through libb200msm.so's kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o example_math gpu-acceleration_b200/cpp/example_math.cu
#include <cstdio>
#include <cstdlib>

#include "b200math.cuh"

__global__ void k_walk(int n, int* bad, uint32_t* last_x) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    affine_t G;
    G.x = fq_one();
    G.y = fq_dbl(fq_one());
    const fq three = fq_add(fq_dbl(fq_one()), fq_one());
    xyzz_t acc = xyzz_inf();
    int errors = 0;
    for (int k = 1; k <= n; k++) {
        xyzz_madd(acc, G);                                   // k = 1: from infinity; k = 2: P + P (the doubling path)
        const fq izz = fq_inv_by(acc.zz), izzz = fq_inv_by(acc.zzz);
        const fq x = fq_mul(acc.x, izz), y = fq_mul(acc.y, izzz);
        const fq lhs = fq_sqr(y), rhs = fq_add(fq_mul(fq_sqr(x), x), three);
        if (!fq_eq(lhs, rhs)) errors++;
        if (!fq_eq(fq_inv(acc.zz), izz)) errors++;           // Fermat and safegcd agree
        if (k == n)
            for (int j = 0; j < 8; j++) last_x[j] = x.v[j];
    }
    xyzz_t neg = xyzz_neg(acc);
    xyzz_add(acc, neg);                                      // P + (-P) = infinity
    if (!xyzz_is_inf(acc)) errors++;
    *bad = errors;
}

int main() {
    int *d_bad, h_bad = -1;
    uint32_t *d_x, h_x[8];
    if (cudaMalloc(&d_bad, 4) != cudaSuccess || cudaMalloc(&d_x, 32) != cudaSuccess) {
        fprintf(stderr, "no CUDA device\n");
        return 2;
    }
    k_walk<<<1, 32>>>(64, d_bad, d_x);
    if (cudaMemcpy(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return 2;
    cudaMemcpy(h_x, d_x, 32, cudaMemcpyDeviceToHost);
    printf("errors=%d x64G_mont_limb0=%08x\n", h_bad, h_x[0]);
    return h_bad == 0 ? 0 : 1;
}
