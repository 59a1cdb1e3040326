// BN254 base-field arithmetic on the FP64 pipe of sm_100a ("fd": field elements held in doubles).
//
// Why: on B200 a 32x32->64 multiply-accumulate (IMAD.WIDE.U32.X) takes TWO passes of the 16-lane fmaheavy pipe
// (4 cycles per warp instruction), while DFMA issues at full rate (2 cycles) on a pipe of
// its own.  A 52x52-bit limb product costs 2 DFMA + 1 DADD (6 FP64-pipe cycles for 2704 bit products) against
// 4 cycles per 1024 bit products on the integer pipe, and the pipes are separate -- on paper 1.7x alone and ~2x with
// FP64 warps beside IMAD warps.
//
// MEASURED NEGATIVE (B200, tools/microbench/fd_bench.cu, profiles/r02q_*): bit-exact against fq_mul on 65 536
// operands, but 6.1e10 products/s against 6.2e10 for the integer product, and 6.95e10 (+12 %) for FP64 and IMAD
// warps side by side.  Why (pipe_bench5/6/7): (1) a DFMA with three REGISTER operands issues every 3.15 cycles, not 2.07
// (register-file bandwidth: 6 operand registers per thread), and the product needs 75 of them; (2) an IMAD.WIDE.U32.X
// holds the sub-partition's issue port for ~4 cycles, so FP64 warps only get the leftover slots -- the pipes are
// separate, the dispatch is not.  ncu: FP64 pipe 59 % + ALU 50 % busy, issue 56 %, flat from 4 to 8 warps per
// sub-partition.  Kept here as evidence; the product path does not use it.
//
// Representation: five radix-2^52 limbs held as doubles with exact integer values in [0, 2^52) ("tight").
// Montgomery factor R' = 2^260 (five word-serial rounds of 52 bits).  Values are lazily reduced: any operand
// below 2^260 is legal and a product is < a*b/2^260 + p.  The memory format of the rest of the engine (arkworks
// Montgomery words, R = 2^256) is bridged without extra multiplications on the hot path:
//   * a base coordinate x~ = x*2^256 enters its product as the integer 16*x~ == x*2^260 (mod p); the factor 16 is
//     only a different bit offset when the 256-bit word is cut into limbs (fd_from_fq_x16);
//   * a result leaves through one product with 2^256 mod p and a canonicalisation (fd_to_fq), once per bucket.
// The limb product follows the published double-precision technique (Emmart, Zheng, Weems, "Faster modular
// exponentiation using double precision floating point arithmetic on the GPU", ARITH 2018): with c1 = 2^104,
// hi = fma_rz(a, b, c1) carries floor(a*b / 2^52) in its mantissa and lo = fma_rz(a, b, (c1 + 2^52) - hi)
// carries a*b mod 2^52; the raw IEEE bit patterns are summed as 64-bit integers (3-input IADD3 pairs) and the
// exponent fields are cancelled by constants folded into the column initialisers.  Subtractions are folded into
// the integer column stage of the product that feeds them (fd_cols_sub), where signed carries are free.
// Every step is restated with exact integer arithmetic in tools/microbench/fd_model.py and compared with the
// integer path on the device (fd_bench check).  Replaces, like fq.cuh, the reference's mont_mul_cios
// (/root/reference/mopro-msm/src/msm/metal_msm/shader/mont_backend/mont.metal:105-181).
#pragma once
#include "fq.cuh"

struct fd {
    double v[5];
};

#define FD_C1 0x1p104
#define FD_C2 (0x1p104 + 0x1p52)
#define FD_TWO52 0x1p52
#define FD_MASK52 0xFFFFFFFFFFFFFull
#define FD_BH 0x4670000000000000ull  // raw bits of 2^104: exponent field of every "hi" term
#define FD_BL 0x4330000000000000ull  // raw bits of 2^52:  exponent field of every "lo" term

// p in radix 2^52, n0' = -p^-1 mod 2^52
#define FD_P0 154029749239111.0
#define FD_P1 2558044347618242.0
#define FD_P2 423691504025962.0
#define FD_P3 2817616741948264.0
#define FD_P4 53207371014449.0
#define FD_N0 571208714576777.0
__device__ __forceinline__ double fd_p(int j) {
    return j == 0 ? FD_P0 : j == 1 ? FD_P1 : j == 2 ? FD_P2 : j == 3 ? FD_P3 : FD_P4;
}
__device__ __forceinline__ uint64_t fd_pi(int j) {  // the same limbs as integers
    return j == 0 ? 154029749239111ull : j == 1 ? 2558044347618242ull : j == 2 ? 423691504025962ull
         : j == 3 ? 2817616741948264ull : 53207371014449ull;
}

__device__ __forceinline__ constexpr uint32_t fq_p_word(int i) {
    return i == 0 ? 0xd87cfd47u : i == 1 ? 0x3c208c16u : i == 2 ? 0x6871ca8du : i == 3 ? 0x97816a91u
         : i == 4 ? 0x8181585du : i == 5 ? 0xb85045b6u : i == 6 ? 0xe131a029u : 0x30644e72u;
}

// number of limb products that land in column k of a 5x5 product
__device__ __forceinline__ constexpr int fd_cnt(int k) { return (k < 0 || k > 8) ? 0 : (k < 4 ? k : 8 - k) + 1; }
// what `n` products (n may be 0 or negative: subtracted products cancel added ones) plus one reduction leave in
// the exponent fields of column k, negated: the column initialiser
__device__ __forceinline__ constexpr uint64_t fd_init(int k, int n) {
    return 0ull - ((uint64_t)((n + 1) * fd_cnt(k)) * FD_BL + (uint64_t)((n + 1) * fd_cnt(k - 1)) * FD_BH);
}

// raw (hi, lo) bit patterns of the 104-bit product of two tight limbs
__device__ __forceinline__ void fd_split(double a, double b, uint64_t& hi, uint64_t& lo) {
    const double h = __fma_rz(a, b, FD_C1);
    const double s = FD_C2 - h;
    const double l = __fma_rz(a, b, s);
    hi = (uint64_t)__double_as_longlong(h);
    lo = (uint64_t)__double_as_longlong(l);
}

// columns for `n` net products and one reduction; `kp` adds kp * p * 2^260 (keeps a*b - c*d non-negative)
template <int N, int KP>
__device__ __forceinline__ void fd_cols_init(uint64_t (&c)[10]) {
#pragma unroll
    for (int k = 0; k < 10; k++) c[k] = fd_init(k, N) + (k >= 5 ? (uint64_t)KP * fd_pi(k - 5) : 0ull);
}
// c += a*b (25 limb products)
__device__ __forceinline__ void fd_cols_mac(uint64_t (&c)[10], const fd& a, const fd& b) {
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) {
            uint64_t hi, lo;
            fd_split(a.v[i], b.v[j], hi, lo);
            c[i + j] += lo;
            c[i + j + 1] += hi;
        }
}
// c -= a*b
__device__ __forceinline__ void fd_cols_msub(uint64_t (&c)[10], const fd& a, const fd& b) {
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) {
            uint64_t hi, lo;
            fd_split(a.v[i], b.v[j], hi, lo);
            c[i + j] -= lo;
            c[i + j + 1] -= hi;
        }
}
// c += a*a (15 limb products, the off-diagonal ones counted twice)
__device__ __forceinline__ void fd_cols_sqr(uint64_t (&c)[10], const fd& a) {
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = i; j < 5; j++) {
            uint64_t hi, lo;
            fd_split(a.v[i], a.v[j], hi, lo);
            if (i == j) {
                c[i + j] += lo;
                c[i + j + 1] += hi;
            } else {
                c[i + j] += lo + lo;
                c[i + j + 1] += hi + hi;
            }
        }
}
// five Montgomery rounds: afterwards c[5..9] hold (T + m*p) / 2^260 as true (small, possibly signed) integers
__device__ __forceinline__ void fd_cols_redc(uint64_t (&c)[10]) {
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const double xd = __longlong_as_double((long long)((c[k] & FD_MASK52) | FD_BL)) - FD_TWO52;
        const double h = __fma_rz(xd, FD_N0, FD_C1);
        const double s = FD_C2 - h;
        const double qb = __fma_rz(xd, FD_N0, s);  // 2^52 + (x * n0' mod 2^52)
        const double q = qb - FD_TWO52;
#pragma unroll
        for (int j = 0; j < 5; j++) {
            uint64_t hi, lo;
            fd_split(q, fd_p(j), hi, lo);
            c[k + j] += lo;
            c[k + j + 1] += hi;
        }
        c[k + 1] += (uint64_t)((long long)c[k] >> 52);  // c[k] is now a multiple of 2^52
    }
}
// c[5..9] -= s (limb-wise, before the carries are resolved); s tight
__device__ __forceinline__ void fd_cols_sub(uint64_t (&c)[10], const fd& s) {
#pragma unroll
    for (int i = 0; i < 5; i++) c[5 + i] -= (uint64_t)__double_as_longlong(s.v[i] + FD_TWO52) - FD_BL;
}
__device__ __forceinline__ void fd_cols_add(uint64_t (&c)[10], const fd& s) {
#pragma unroll
    for (int i = 0; i < 5; i++) c[5 + i] += (uint64_t)__double_as_longlong(s.v[i] + FD_TWO52) - FD_BL;
}
// c[5..9] += k * p
template <int K>
__device__ __forceinline__ void fd_cols_addp(uint64_t (&c)[10]) {
#pragma unroll
    for (int i = 0; i < 5; i++) c[5 + i] += (uint64_t)K * fd_pi(i);
}
// resolve the (signed) carries of c[5..9] and hand the value back as tight doubles; the value must be in [0, 2^260)
__device__ __forceinline__ fd fd_cols_norm(const uint64_t (&c)[10]) {
    fd r;
    uint64_t t = c[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
        r.v[i] = __longlong_as_double((long long)((t & FD_MASK52) | FD_BL)) - FD_TWO52;
        if (i < 4) t = c[6 + i] + (uint64_t)((long long)t >> 52);
    }
    return r;
}
// the same with the columns negated first (value = -(columns), must again be in [0, 2^260))
__device__ __forceinline__ fd fd_cols_norm_neg(const uint64_t (&c)[10]) {
    fd r;
    uint64_t t = 0ull - c[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
        r.v[i] = __longlong_as_double((long long)((t & FD_MASK52) | FD_BL)) - FD_TWO52;
        if (i < 4) t = (uint64_t)((long long)t >> 52) - c[6 + i];
    }
    return r;
}

__device__ __forceinline__ fd fd_mul(const fd& a, const fd& b) {
    uint64_t c[10];
    fd_cols_init<1, 0>(c);
    fd_cols_mac(c, a, b);
    fd_cols_redc(c);
    return fd_cols_norm(c);
}
__device__ __forceinline__ fd fd_sqr(const fd& a) {
    uint64_t c[10];
    fd_cols_init<1, 0>(c);
    fd_cols_sqr(c, a);
    fd_cols_redc(c);
    return fd_cols_norm(c);
}

// ---- bridges to the 8 x 32-bit Montgomery words (R = 2^256)
// the 256-bit integer a, times 2^SH (SH = 0 or 4), cut into limbs
template <int SH>
__device__ __forceinline__ fd fd_from_words(const fq& a) {
    uint64_t w[5];
    w[0] = (uint64_t)a.v[0] | ((uint64_t)a.v[1] << 32);
    w[1] = (uint64_t)a.v[2] | ((uint64_t)a.v[3] << 32);
    w[2] = (uint64_t)a.v[4] | ((uint64_t)a.v[5] << 32);
    w[3] = (uint64_t)a.v[6] | ((uint64_t)a.v[7] << 32);
    w[4] = 0;
    fd r;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const int lo = 52 * i - SH;  // first bit of limb i inside a
        uint64_t t;
        if (lo < 0) t = w[0] << (-lo);
        else {
            const int q = lo >> 6, s = lo & 63;
            t = w[q] >> s;
            if (s > 12) t |= w[q + 1] << (64 - s);
        }
        r.v[i] = __longlong_as_double((long long)((t & FD_MASK52) | FD_BL)) - FD_TWO52;
    }
    return r;
}
__device__ __forceinline__ fd fd_from_fq_x16(const fq& a) { return fd_from_words<4>(a); }  // x~ -> 16 x~ (lazy, < 16p)
__device__ __forceinline__ fd fd_const(double a0, double a1, double a2, double a3, double a4) {
    fd r;
    r.v[0] = a0; r.v[1] = a1; r.v[2] = a2; r.v[3] = a3; r.v[4] = a4;
    return r;
}
__device__ __forceinline__ fd fd_one() {  // 2^260 mod p
    return fd_const(572299946026164.0, 1297056913851477.0, 438710680783112.0, 2010973926982104.0, 34180462156727.0);
}
// x*2^256 (words) -> x*2^260, below 2p
__device__ __forceinline__ fd fd_from_fq(const fq& a) {
    // times 2^264 mod p, divided by 2^260 by the reduction
    return fd_mul(fd_from_words<0>(a),
                  fd_const(3112902016657018.0, 4179666400182205.0, 2782455852270170.0, 3999415412231024.0, 14813684363142.0));
}
// x*2^260 (any lazy value) -> canonical x*2^256 words
__device__ __forceinline__ fq fd_to_fq(const fd& a) {
    uint64_t c[10];
    fd_cols_init<1, 0>(c);
    // times 2^256 mod p
    fd_cols_mac(c, a, fd_const(3733450881174941.0, 720577144020277.0, 2385142107240683.0, 3926314799740663.0, 15438121638407.0));
    fd_cols_redc(c);
    // integer limbs of the (< 2p) result
    uint64_t l[5];
    uint64_t t = c[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
        l[i] = t & FD_MASK52;
        if (i < 4) t = c[6 + i] + (uint64_t)((long long)t >> 52);
    }
    uint64_t w0 = l[0] | (l[1] << 52);
    uint64_t w1 = (l[1] >> 12) | (l[2] << 40);
    uint64_t w2 = (l[2] >> 24) | (l[3] << 28);
    uint64_t w3 = (l[3] >> 36) | (l[4] << 16);
    fq r;
    r.v[0] = (uint32_t)w0; r.v[1] = (uint32_t)(w0 >> 32); r.v[2] = (uint32_t)w1; r.v[3] = (uint32_t)(w1 >> 32);
    r.v[4] = (uint32_t)w2; r.v[5] = (uint32_t)(w2 >> 32); r.v[6] = (uint32_t)w3; r.v[7] = (uint32_t)(w3 >> 32);
    // r < 2p < 2^255: one conditional subtraction
    uint32_t s[8];
    uint32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint64_t d = (uint64_t)r.v[i] - fq_p_word(i) - borrow;
        s[i] = (uint32_t)d;
        borrow = (uint32_t)(d >> 63);
    }
    if (!borrow) {
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = s[i];
    }
    return r;
}
