// BN254 base-field (Fq) arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form, R = 2^256.
//
// The in-memory representation is exactly arkworks' `Fp<MontBackend<FqConfig,4>,4>`
// (four little-endian u64 = eight little-endian u32 holding a*R mod p), so bases are
// consumed without any conversion.  Replaces the reference's 16x16-bit-limb math library
// (/root/reference/mopro-msm/src/msm/metal_msm/shader/{bigint,field,mont_backend}/*.metal):
//   mont_mul_cios   mont_backend/mont.metal:105-181  -> fq_mul (IMAD.WIDE.U32.X carry chains)
//   ff_add/ff_sub   field/ff.metal:9-35              -> fq_add / fq_sub
//   bigint_*        bigint/bigint.metal:7-178        -> folded into the asm bodies
// All values are kept fully reduced in [0, p).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "fq_asm.inc"

struct fq {
    uint32_t v[8];
};

__device__ __constant__ const uint32_t FQ_P[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                                  0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
// R mod p (Montgomery one)
#define FQ_ONE_INIT {{0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}}

__device__ __forceinline__ fq fq_zero() {
    fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
__device__ __forceinline__ fq fq_one() {
    fq r = FQ_ONE_INIT;
    return r;
}
// The product is inlined at every call site.  An out-of-line shared body (fq_mul_call) was tried because the
// inlined mixed addition is ~35 KB of straight-line code: +10% in an isolated register-only loop
// (pipe_bench, 5.76e9 vs 5.24e9 madd/s) but SLOWER inside k_accumulate (3.08 vs 2.79 ms at 2^20, with either
// a 128-register cap + spills or 142 registers at 3 CTAs/SM) and in the reduce kernels, so it is not used.
__device__ __forceinline__ fq fq_mul(const fq& a, const fq& b) {
    fq r;
    fq_mul_asm(r.v, a.v, b.v);
    return r;
}
__device__ __forceinline__ fq fq_mul_inline(const fq& a, const fq& b) { return fq_mul(a, b); }
// a*b - c*d with ONE Montgomery reduction (fq_mulsub_asm: 200 wide MACs instead of 272 for two products + a subtraction):
// the Y3 = R (Q - X3) - Y1 PPP step of every XYZZ addition and doubling.
__device__ __forceinline__ fq fq_mulsub(const fq& a, const fq& b, const fq& c, const fq& d) {
    fq r;
    fq_mulsub_asm(r.v, a.v, b.v, c.v, d.v);
    return r;
}
// a*b + c*d with ONE Montgomery reduction (fq_muladd_asm, 200 wide MACs): with fq_mulsub the two halves of an Fq2 product.
__device__ __forceinline__ fq fq_muladd(const fq& a, const fq& b, const fq& c, const fq& d) {
    fq r;
    fq_muladd_asm(r.v, a.v, b.v, c.v, d.v);
    return r;
}
// A dedicated squaring (fq_sqr_asm: 108 wide MACs instead of 136, verified in tests/test_ptx_arith.py) was
// MEASURED no faster on B200 (6.8e10 vs 6.6e10 /s stand-alone; k_accumulate 2.96 vs 2.79 ms at 2^20): its
// ~100 extra carry-propagation IADD3.X cancel the 28 saved multiplies.  The plain product is used.
__device__ __forceinline__ fq fq_sqr(const fq& a) { return fq_mul(a, a); }
__device__ __forceinline__ fq fq_add(const fq& a, const fq& b) {
    fq r;
    fq_add_asm(r.v, a.v, b.v);
    return r;
}
__device__ __forceinline__ fq fq_sub(const fq& a, const fq& b) {
    fq r;
    fq_sub_asm(r.v, a.v, b.v);
    return r;
}
__device__ __forceinline__ fq fq_dbl(const fq& a) { return fq_add(a, a); }
__device__ __forceinline__ bool fq_is_zero(const fq& a) {
    uint32_t o = a.v[0];
#pragma unroll
    for (int i = 1; i < 8; i++) o |= a.v[i];
    return o == 0;
}
__device__ __forceinline__ bool fq_eq(const fq& a, const fq& b) {
    uint32_t o = a.v[0] ^ b.v[0];
#pragma unroll
    for (int i = 1; i < 8; i++) o |= a.v[i] ^ b.v[i];
    return o == 0;
}
__device__ __forceinline__ fq fq_neg(const fq& a) {
    // p - a, with 0 -> 0 (jacobian_neg semantics, curve/jacobian.metal:195-210)
    return fq_sub(fq_zero(), a);
}
// conditional negate without a branch on the sign bit
__device__ __forceinline__ fq fq_cneg(const fq& a, bool neg) {
    fq n = fq_neg(a);
    fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = neg ? n.v[i] : a.v[i];
    return r;
}

// a^(p-2) mod p, MSB-first square-and-multiply (one-time set-up work only).
__device__ __noinline__ fq fq_inv(const fq& a) {
    const uint32_t e[8] = {0xd87cfd45u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    fq r = fq_one();
    for (int bit = 253; bit >= 0; bit--) {
        r = fq_sqr(r);
        if ((e[bit >> 5] >> (bit & 31)) & 1) r = fq_mul(r, a);
    }
    return r;
}

// 32-byte aligned vector access (two 16-byte transactions)
__device__ __forceinline__ fq fq_load(const void* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = q[0], hi = q[1];
    fq r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
__device__ __forceinline__ fq fq_load_nc(const void* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = __ldg(q), hi = __ldg(q + 1);
    fq r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
__device__ __forceinline__ void fq_store(void* p, const fq& a) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

// ---------------------------------------------------------------------------------------
// Scalar field: only Montgomery -> canonical is needed (arkworks `Fr` memory holds s*R mod r;
// the reference does this on the CPU with `into_bigint()`, utils/limbs_conversion.rs:311-378).
__device__ __constant__ const uint32_t FR_R[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                                  0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
#define FR_N0 0xefffffffu

// t <- t * R^-1 mod r, fully reduced.  8 rounds of (m = t0 * n0'; t += m*r; t >>= 32).
__device__ __forceinline__ void fr_from_mont(uint32_t (&t)[8]) {
    const uint32_t r[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                           0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t m = t[0] * FR_N0;
        uint64_t c = (uint64_t)m * r[0] + t[0];
        c >>= 32;
#pragma unroll
        for (int j = 1; j < 8; j++) {
            c += (uint64_t)m * r[j] + t[j];
            t[j - 1] = (uint32_t)c;
            c >>= 32;
        }
        t[7] = (uint32_t)c;
    }
    // conditional subtract r
    uint32_t s[8];
    uint64_t b = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint64_t d = (uint64_t)t[j] - r[j] - b;
        s[j] = (uint32_t)d;
        b = (d >> 63) & 1;
    }
    if (!b) {
#pragma unroll
        for (int j = 0; j < 8; j++) t[j] = s[j];
    }
}
