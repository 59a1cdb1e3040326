// Test-kit kernels: synthetic input generation and element-wise wrappers around the production
// device functions (the role of the reference's single-thread `test_*` kernels, SURVEY §2.2, and of
// `test_utils::generate_random_bases_and_scalars`, metal_msm.rs:698-731).  Not on the MSM path.
#pragma once
#include "g1.cuh"

__host__ __device__ __forceinline__ uint64_t tk_mix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

// uniform 256-bit words below r (rejection on 254-bit candidates); out as 4 LE u64
__host__ __device__ inline void tk_random_below_r(uint64_t seed, uint64_t index, uint64_t out[4]) {
    const uint64_t r[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    for (uint64_t attempt = 0; attempt < 64; attempt++) {
        for (int k = 0; k < 4; k++) out[k] = tk_mix64(seed ^ tk_mix64((index * 4 + k) + (attempt << 44)));
        out[3] &= 0x3fffffffffffffffull;
        bool lt = false;
        for (int k = 3; k >= 0; k--) {
            if (out[k] != r[k]) { lt = out[k] < r[k]; break; }
        }
        if (lt) return;
    }
    out[0] = 1; out[1] = out[2] = out[3] = 0;
}

__global__ void k_tk_gen_scalars(uint64_t seed, uint32_t n, uint64_t* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t v[4];
    tk_random_below_r(seed, i, v);
    for (int k = 0; k < 4; k++) out[(size_t)i * 4 + k] = v[k];
}

__device__ __noinline__ fq tk_fq_inv(const fq& a) {
    // a^(p-2), MSB-first square-and-multiply
    const uint32_t e[8] = {0xd87cfd45u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    fq r = fq_one();
    for (int bit = 253; bit >= 0; bit--) {
        r = fq_sqr(r);
        if ((e[bit >> 5] >> (bit & 31)) & 1) r = fq_mul(r, a);
    }
    return r;
}

__device__ __noinline__ affine_t tk_xyzz_to_affine(const xyzz_t& a) {
    affine_t r;
    fq i = tk_fq_inv(fq_mul(a.zz, a.zzz));
    r.x = fq_mul(a.x, fq_mul(i, a.zzz));
    r.y = fq_mul(a.y, fq_mul(i, a.zz));
    return r;
}

// table[k] = dlog[k] * G  (dlogs canonical LE, 8 x u32 each, non-zero)
__global__ void k_tk_gen_table(const uint32_t* __restrict__ dlogs, uint32_t count, affine_t* __restrict__ out) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    affine_t G;
    G.x = fq_one();
    G.y = fq_dbl(fq_one());
    xyzz_t acc = xyzz_inf();
    for (int bit = 253; bit >= 0; bit--) {
        xyzz_dbl_inplace(acc);
        if ((dlogs[(size_t)k * 8 + (bit >> 5)] >> (bit & 31)) & 1) xyzz_madd(acc, G);
    }
    affine_t r = tk_xyzz_to_affine(acc);
    char* o = reinterpret_cast<char*>(out + k);
    fq_store(o, r.x); fq_store(o + 32, r.y);
}

// base[i] = T1[i mod 4096] + T2[i / 4096]
__global__ void k_tk_gen_bases(const affine_t* __restrict__ T1, const affine_t* __restrict__ T2, uint32_t n,
                               affine_t* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine_t a = affine_load_nc(T1 + (i & 4095)), b = affine_load_nc(T2 + (i >> 12));
    xyzz_t acc = xyzz_from_affine(a);
    xyzz_madd(acc, b);
    affine_t r = tk_xyzz_to_affine(acc);
    char* o = reinterpret_cast<char*>(out + i);
    fq_store(o, r.x); fq_store(o + 32, r.y);
}

__global__ void k_tk_op(int op, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint8_t* __restrict__ out, uint32_t count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    if (op < 10) {
        fq x = fq_load(a + (size_t)i * 32), y = fq_zero();
        if (b) y = fq_load(b + (size_t)i * 32);
        fq r;
        switch (op) {
            case 0: r = fq_mul(x, y); break;
            case 1: r = fq_add(x, y); break;
            case 2: r = fq_sub(x, y); break;
            default: r = fq_sqr(x); break;
        }
        fq_store(out + (size_t)i * 32, r);
    } else if (op == 10) {
        xyzz_t acc = xyzz_load(a + (size_t)i * 128);
        affine_t p;
        p.x = fq_load(b + (size_t)i * 64); p.y = fq_load(b + (size_t)i * 64 + 32);
        xyzz_madd(acc, p);
        xyzz_store(out + (size_t)i * 128, acc);
    } else if (op == 11) {
        xyzz_t acc = xyzz_load(a + (size_t)i * 128), v = xyzz_load(b + (size_t)i * 128);
        xyzz_add(acc, v);
        xyzz_store(out + (size_t)i * 128, acc);
    } else if (op == 12) {
        xyzz_t acc = xyzz_load(a + (size_t)i * 128);
        xyzz_dbl_inplace(acc);
        xyzz_store(out + (size_t)i * 128, acc);
    } else if (op == 13) {
        xyzz_t acc = xyzz_load(a + (size_t)i * 128);
        jac_t r = xyzz_to_jacobian(acc);
        uint8_t* o = out + (size_t)i * 96;
        fq_store(o, r.x); fq_store(o + 32, r.y); fq_store(o + 64, r.z);
    } else if (op == 20) {
        const uint4* q = reinterpret_cast<const uint4*>(a + (size_t)i * 32);
        uint4 lo = q[0], hi = q[1];
        uint32_t t[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        fr_from_mont(t);
        uint4* o = reinterpret_cast<uint4*>(out + (size_t)i * 32);
        o[0] = make_uint4(t[0], t[1], t[2], t[3]);
        o[1] = make_uint4(t[4], t[5], t[6], t[7]);
    }
}

__global__ void k_tk_imad_wide(uint64_t* out, int iters, uint32_t a, uint32_t b) {
    uint64_t x[8];
    uint32_t m[8];
    for (int k = 0; k < 8; k++) { x[k] = threadIdx.x + k; m[k] = a + k * b + threadIdx.x; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[k]) : "r"(m[k]), "r"(b));
    }
    uint64_t s = 0;
    for (int k = 0; k < 8; k++) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
