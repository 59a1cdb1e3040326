// Fast modular inversion in Fq for the batched-affine bucket accumulation (msm_ba_kernels.cuh).
//
// Fermat's a^(p-2) costs ~380 field multiplications, all on the integer-multiply pipe that bounds the whole MSM; a
// batched-affine addition amortises ONE inversion over the additions of a lane's batch, so the inversion price decides
// how small that batch may be.  This is the Bernstein-Yang "safegcd" inversion (divsteps with delta = 1/2, the
// formulation of the public-domain modinv32 design: signed 30-bit limbs, 20 batches of 30 divsteps = 600 >= the proven
// bound of 590 for 256-bit inputs): every lane executes the SAME instruction sequence whatever its operand, which is
// what SIMT needs (no divergence, fixed trip counts), and most of the work is ALU-pipe shifts / adds / logic
// (600 x ~14 instructions) plus ~1800 multiply-accumulates for the matrix updates -- about 45 field-multiplication
// times, of which only ~13 on the multiply pipe.
//
// The reference has no inversion on its hot path (it stays in Jacobian coordinates throughout:
// shader/curve/jacobian.metal:46-100); its only modular inverses are the host-side parameter derivations
// (utils/mont_params.rs:31-88, egcd over BigUint).
#pragma once
#include "fq.cuh"

#define BY_M30 0x3fffffff
// p in signed-30 limbs and p^-1 mod 2^30
#define BY_P0 0x187cfd47
#define BY_P1 0x3082305b
#define BY_P2 0x071ca8d3
#define BY_P3 0x205aa45a
#define BY_P4 0x01585d97
#define BY_P5 0x0116da06
#define BY_P6 0x1a029b85
#define BY_P7 0x139cb84c
#define BY_P8 0x00003064
#define BY_PINV30 0x1b799c77u
// R^3 mod p (R = 2^256), 8 x 32-bit little-endian
#define BY_R3_0 0xda1530dfu
#define BY_R3_1 0xb1cd6dafu
#define BY_R3_2 0xa7283db6u
#define BY_R3_3 0x62f210e6u
#define BY_R3_4 0x0ada0afbu
#define BY_R3_5 0xef7f0b0cu
#define BY_R3_6 0x2d592544u
#define BY_R3_7 0x20fd6e90u

struct by_mat {
    int32_t u, v, q, r;
};

// 30 divsteps on the low words; returns the new zeta.  2^30 * [f'; g'] = t * [f; g].
__device__ __forceinline__ int32_t by_divsteps_30(int32_t zeta, uint32_t f, uint32_t g, by_mat& t) {
    uint32_t u = 1, v = 0, q = 0, r = 1;
#pragma unroll 6
    for (int i = 0; i < 30; i++) {
        uint32_t c1 = (uint32_t)(zeta >> 31);
        const uint32_t c2 = 0u - (g & 1u);
        const uint32_t x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;
        g += x & c2;
        q += y & c2;
        r += z & c2;
        c1 &= c2;
        zeta = (int32_t)((uint32_t)zeta ^ c1) - 1;
        f += g & c1;
        u += q & c1;
        v += r & c1;
        g >>= 1;
        u <<= 1;
        v <<= 1;
    }
    t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
    return zeta;
}

__device__ __forceinline__ void by_update_fg(int32_t (&f)[9], int32_t (&g)[9], const by_mat& t) {
    int64_t cf = (int64_t)t.u * f[0] + (int64_t)t.v * g[0];
    int64_t cg = (int64_t)t.q * f[0] + (int64_t)t.r * g[0];
    cf >>= 30;
    cg >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        cf += (int64_t)t.u * f[i] + (int64_t)t.v * g[i];
        cg += (int64_t)t.q * f[i] + (int64_t)t.r * g[i];
        f[i - 1] = (int32_t)cf & BY_M30;
        g[i - 1] = (int32_t)cg & BY_M30;
        cf >>= 30;
        cg >>= 30;
    }
    f[8] = (int32_t)cf;
    g[8] = (int32_t)cg;
}

__device__ __forceinline__ void by_update_de(int32_t (&d)[9], int32_t (&e)[9], const by_mat& t, uint32_t pinv30) {
    const int32_t P[9] = {BY_P0, BY_P1, BY_P2, BY_P3, BY_P4, BY_P5, BY_P6, BY_P7, BY_P8};
    const int32_t sd = d[8] >> 31, se = e[8] >> 31;
    int32_t md = (t.u & sd) + (t.v & se);
    int32_t me = (t.q & sd) + (t.r & se);
    int64_t cd = (int64_t)t.u * d[0] + (int64_t)t.v * e[0];
    int64_t ce = (int64_t)t.q * d[0] + (int64_t)t.r * e[0];
    md -= (int32_t)((pinv30 * (uint32_t)cd + (uint32_t)md) & BY_M30);
    me -= (int32_t)((pinv30 * (uint32_t)ce + (uint32_t)me) & BY_M30);
    cd += (int64_t)P[0] * md;
    ce += (int64_t)P[0] * me;
    cd >>= 30;
    ce >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        cd += (int64_t)t.u * d[i] + (int64_t)t.v * e[i] + (int64_t)P[i] * md;
        ce += (int64_t)t.q * d[i] + (int64_t)t.r * e[i] + (int64_t)P[i] * me;
        d[i - 1] = (int32_t)cd & BY_M30;
        e[i - 1] = (int32_t)ce & BY_M30;
        cd >>= 30;
        ce >>= 30;
    }
    d[8] = (int32_t)cd;
    e[8] = (int32_t)ce;
}

// a (Montgomery form, a*R) -> a^-1 in Montgomery form (a^-1 * R); 0 -> 0.
__device__ __noinline__ fq fq_inv_by(const fq a) {
    const int32_t P[9] = {BY_P0, BY_P1, BY_P2, BY_P3, BY_P4, BY_P5, BY_P6, BY_P7, BY_P8};
    int32_t f[9], g[9], d[9], e[9];
#pragma unroll
    for (int i = 0; i < 9; i++) { f[i] = P[i]; d[i] = 0; e[i] = 0; }
    e[0] = 1;
    // 8 x 32-bit -> 9 x 30-bit
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
        uint32_t lo = w < 8 ? a.v[w] >> sh : 0u;
        if (sh > 2 && w + 1 < 8) lo |= a.v[w + 1] << (32 - sh);
        g[i] = (int32_t)(lo & BY_M30);
    }
    int32_t zeta = -1;
    for (int it = 0; it < 20; it++) {
        by_mat t;
        zeta = by_divsteps_30(zeta, (uint32_t)f[0], (uint32_t)g[0], t);
        by_update_de(d, e, t, BY_PINV30);
        by_update_fg(f, g, t);
    }
    // g == 0, f == +-1 (or f == +-p when a == 0): d = +- a^-1, in (-2p, p)
    {
        int32_t ca = d[8] >> 31;
#pragma unroll
        for (int i = 0; i < 9; i++) d[i] += P[i] & ca;
        const int32_t cn = f[8] >> 31;
#pragma unroll
        for (int i = 0; i < 9; i++) d[i] = (d[i] ^ cn) - cn;
#pragma unroll
        for (int i = 0; i < 8; i++) { d[i + 1] += d[i] >> 30; d[i] &= BY_M30; }
        ca = d[8] >> 31;
#pragma unroll
        for (int i = 0; i < 9; i++) d[i] += P[i] & ca;
#pragma unroll
        for (int i = 0; i < 8; i++) { d[i + 1] += d[i] >> 30; d[i] &= BY_M30; }
    }
    // 9 x 30-bit -> 8 x 32-bit
    fq r;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const int bit = 32 * w, i = bit / 30, sh = bit - 30 * i;
        uint32_t lo = (uint32_t)d[i] >> sh;
        lo |= (uint32_t)d[i + 1] << (30 - sh);
        if (60 - sh < 32 && i + 2 < 9) lo |= (uint32_t)d[i + 2] << (60 - sh);
        r.v[w] = lo;
    }
    // r = (aR)^-1 = a^-1 R^-1; one Montgomery product with R^3 mod p gives a^-1 R
    const fq R3 = {{BY_R3_0, BY_R3_1, BY_R3_2, BY_R3_3, BY_R3_4, BY_R3_5, BY_R3_6, BY_R3_7}};
    return fq_mul(r, R3);
}
