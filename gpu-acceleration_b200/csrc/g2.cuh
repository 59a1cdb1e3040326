// BN254 G2 (the sextic twist y^2 = x^3 + 3/(9+u) over Fq2 = Fq[u]/(u^2 + 1)) for the B2 MSM of a Groth16 prover.
// SURVEY.md 8(f) rank 4; the reference has no G2 code (its roadmap names it, README.md:199-200).
//
// Same design as g1.cuh one level up: XYZZ accumulators (infinity <=> ZZ == 0), complete operations (infinity operands,
// P + P and P + (-P) detected by value: every Fq2 coefficient is kept canonical in [0, p)).  An Fq2 product is three Fq
// products (Karatsuba), a square two; the point operations are out of line because one XYZZ accumulator alone is
// 64 registers.  Memory format: arkworks Fq2 = {c0, c1}, each the Montgomery integer in four LE u64.
#pragma once
#include "fq.cuh"

struct fq2 {
    fq c0, c1;
};
struct g2_affine_t {  // 128 bytes: x.c0 || x.c1 || y.c0 || y.c1; all-zero = the device encoding of infinity
    fq2 x, y;
};
struct g2_xyzz_t {  // 256 bytes
    fq2 x, y, zz, zzz;
};
struct g2_jac_t {  // 192 bytes: arkworks G2Projective memory order (x, y, z)
    fq2 x, y, z;
};

__device__ __forceinline__ fq2 fq2_zero() { fq2 r; r.c0 = fq_zero(); r.c1 = fq_zero(); return r; }
__device__ __forceinline__ fq2 fq2_one() { fq2 r; r.c0 = fq_one(); r.c1 = fq_zero(); return r; }
__device__ __forceinline__ fq2 fq2_add(const fq2& a, const fq2& b) { fq2 r; r.c0 = fq_add(a.c0, b.c0); r.c1 = fq_add(a.c1, b.c1); return r; }
__device__ __forceinline__ fq2 fq2_sub(const fq2& a, const fq2& b) { fq2 r; r.c0 = fq_sub(a.c0, b.c0); r.c1 = fq_sub(a.c1, b.c1); return r; }
__device__ __forceinline__ fq2 fq2_dbl(const fq2& a) { return fq2_add(a, a); }
__device__ __forceinline__ fq2 fq2_neg(const fq2& a) { fq2 r; r.c0 = fq_neg(a.c0); r.c1 = fq_neg(a.c1); return r; }
__device__ __forceinline__ fq2 fq2_cneg(const fq2& a, bool neg) { fq2 r; r.c0 = fq_cneg(a.c0, neg); r.c1 = fq_cneg(a.c1, neg); return r; }
__device__ __forceinline__ bool fq2_is_zero(const fq2& a) { return fq_is_zero(a.c0) && fq_is_zero(a.c1); }

// (a0 + a1 u)(b0 + b1 u) = (a0 b0 - a1 b1) + (a0 b1 + a1 b0) u: each half is two limb products under ONE Montgomery reduction
// (fq_mulsub / fq_muladd, 200 wide MACs each) -- 400 wide MACs and no field additions, against 408 + five additions /
// subtractions (~125 ALU instructions that the IMAD chains do not hide) for Karatsuba.
__device__ __noinline__ fq2 fq2_mul(const fq2& a, const fq2& b) {
    fq2 r;
    r.c0 = fq_mulsub(a.c0, b.c0, a.c1, b.c1);
    r.c1 = fq_muladd(a.c0, b.c1, a.c1, b.c0);
    return r;
}
// (a0 + a1 u)^2 = (a0 + a1)(a0 - a1) + 2 a0 a1 u
__device__ __noinline__ fq2 fq2_sqr(const fq2& a) {
    fq2 r;
    r.c0 = fq_mul(fq_add(a.c0, a.c1), fq_sub(a.c0, a.c1));
    r.c1 = fq_dbl(fq_mul(a.c0, a.c1));
    return r;
}

// 1 / (a0 + a1 u) = (a0 - a1 u) / (a0^2 + a1^2) -- set-up work only
__device__ __noinline__ fq2 fq2_inv(const fq2& a) {
    fq nrm = fq_add(fq_sqr(a.c0), fq_sqr(a.c1));
    fq ni = fq_inv(nrm);
    fq2 r;
    r.c0 = fq_mul(a.c0, ni);
    r.c1 = fq_neg(fq_mul(a.c1, ni));
    return r;
}

__device__ __forceinline__ fq2 fq2_load(const void* p) {
    const char* c = reinterpret_cast<const char*>(p);
    fq2 r; r.c0 = fq_load(c); r.c1 = fq_load(c + 32);
    return r;
}
__device__ __forceinline__ fq2 fq2_load_nc(const void* p) {
    const char* c = reinterpret_cast<const char*>(p);
    fq2 r; r.c0 = fq_load_nc(c); r.c1 = fq_load_nc(c + 32);
    return r;
}
__device__ __forceinline__ void fq2_store(void* p, const fq2& a) {
    char* c = reinterpret_cast<char*>(p);
    fq_store(c, a.c0); fq_store(c + 32, a.c1);
}

__device__ __forceinline__ g2_xyzz_t g2_inf() {
    g2_xyzz_t r;
    r.x = fq2_zero(); r.y = fq2_zero(); r.zz = fq2_zero(); r.zzz = fq2_zero();
    return r;
}
__device__ __forceinline__ bool g2_is_inf(const g2_xyzz_t& a) { return fq2_is_zero(a.zz); }
__device__ __forceinline__ bool g2_affine_is_inf(const g2_affine_t& p) { return fq2_is_zero(p.x) && fq2_is_zero(p.y); }

__device__ __forceinline__ g2_xyzz_t g2_load(const void* p) {
    const char* c = reinterpret_cast<const char*>(p);
    g2_xyzz_t r;
    r.x = fq2_load(c); r.y = fq2_load(c + 64); r.zz = fq2_load(c + 128); r.zzz = fq2_load(c + 192);
    return r;
}
__device__ __forceinline__ void g2_store(void* p, const g2_xyzz_t& a) {
    char* c = reinterpret_cast<char*>(p);
    fq2_store(c, a.x); fq2_store(c + 64, a.y); fq2_store(c + 128, a.zz); fq2_store(c + 192, a.zzz);
}
__device__ __forceinline__ g2_affine_t g2_affine_load_nc(const void* p) {
    const char* c = reinterpret_cast<const char*>(p);
    g2_affine_t r;
    r.x = fq2_load_nc(c); r.y = fq2_load_nc(c + 64);
    return r;
}

// dbl-2008-s-1 (a = 0) over Fq2
__device__ __noinline__ void g2_dbl_impl(g2_xyzz_t& a) {
    if (g2_is_inf(a)) return;
    fq2 U = fq2_dbl(a.y);
    fq2 V = fq2_sqr(U);
    fq2 W = fq2_mul(U, V);
    fq2 S = fq2_mul(a.x, V);
    fq2 M = fq2_sqr(a.x);
    M = fq2_add(fq2_dbl(M), M);
    fq2 X3 = fq2_sub(fq2_sub(fq2_sqr(M), S), S);
    fq2 Y3 = fq2_sub(fq2_mul(M, fq2_sub(S, X3)), fq2_mul(W, a.y));
    a.x = X3; a.y = Y3;
    a.zz = fq2_mul(V, a.zz);
    a.zzz = fq2_mul(W, a.zzz);
}

// acc += p (p affine, not the infinity marker).  madd-2008-s over Fq2, complete.
__device__ __noinline__ void g2_madd_impl(g2_xyzz_t& acc, const g2_affine_t& p) {
    if (g2_is_inf(acc)) {
        acc.x = p.x; acc.y = p.y; acc.zz = fq2_one(); acc.zzz = fq2_one();
        return;
    }
    fq2 U2 = fq2_mul(p.x, acc.zz);
    fq2 S2 = fq2_mul(p.y, acc.zzz);
    fq2 P = fq2_sub(U2, acc.x);
    fq2 R = fq2_sub(S2, acc.y);
    if (fq2_is_zero(P)) {
        if (fq2_is_zero(R)) {
            acc.x = p.x; acc.y = p.y; acc.zz = fq2_one(); acc.zzz = fq2_one();
            g2_dbl_impl(acc);
        } else {
            acc = g2_inf();
        }
        return;
    }
    fq2 PP = fq2_sqr(P);
    fq2 PPP = fq2_mul(P, PP);
    fq2 Q = fq2_mul(acc.x, PP);
    fq2 X3 = fq2_sub(fq2_sub(fq2_sub(fq2_sqr(R), PPP), Q), Q);
    fq2 Y3 = fq2_sub(fq2_mul(R, fq2_sub(Q, X3)), fq2_mul(acc.y, PPP));
    acc.x = X3; acc.y = Y3;
    acc.zz = fq2_mul(acc.zz, PP);
    acc.zzz = fq2_mul(acc.zzz, PPP);
}

// acc += b (both XYZZ).  add-2008-s over Fq2, complete.
__device__ __noinline__ void g2_add_impl(g2_xyzz_t& acc, const g2_xyzz_t& b) {
    if (g2_is_inf(b)) return;
    if (g2_is_inf(acc)) { acc = b; return; }
    fq2 U1 = fq2_mul(acc.x, b.zz);
    fq2 U2 = fq2_mul(b.x, acc.zz);
    fq2 S1 = fq2_mul(acc.y, b.zzz);
    fq2 S2 = fq2_mul(b.y, acc.zzz);
    fq2 P = fq2_sub(U2, U1);
    fq2 R = fq2_sub(S2, S1);
    if (fq2_is_zero(P)) {
        if (fq2_is_zero(R)) g2_dbl_impl(acc);
        else acc = g2_inf();
        return;
    }
    fq2 PP = fq2_sqr(P);
    fq2 PPP = fq2_mul(P, PP);
    fq2 Q = fq2_mul(U1, PP);
    fq2 X3 = fq2_sub(fq2_sub(fq2_sub(fq2_sqr(R), PPP), Q), Q);
    fq2 Y3 = fq2_sub(fq2_mul(R, fq2_sub(Q, X3)), fq2_mul(S1, PPP));
    acc.x = X3; acc.y = Y3;
    acc.zz = fq2_mul(fq2_mul(acc.zz, b.zz), PP);
    acc.zzz = fq2_mul(fq2_mul(acc.zzz, b.zzz), PPP);
}

// Wrappers with the compiler barrier g1.cuh documents (cicc 12.9 otherwise drops the reload of the accumulator's first
// word after an out-of-line call).
__device__ __forceinline__ void g2_madd(g2_xyzz_t& acc, const g2_affine_t& p) {
    g2_madd_impl(acc, p);
    asm volatile("" ::: "memory");
}
__device__ __forceinline__ void g2_add(g2_xyzz_t& acc, const g2_xyzz_t& b) {
    g2_add_impl(acc, b);
    asm volatile("" ::: "memory");
}
__device__ __forceinline__ void g2_dbl(g2_xyzz_t& a) {
    g2_dbl_impl(a);
    asm volatile("" ::: "memory");
}

// (X ZZ^2, Y ZZ^3, ZZZ) is a Jacobian representative; infinity -> (1, 1, 0) like arkworks' Projective::zero()
__device__ __forceinline__ g2_jac_t g2_to_jacobian(const g2_xyzz_t& a) {
    g2_jac_t r;
    if (g2_is_inf(a)) {
        r.x = fq2_one(); r.y = fq2_one(); r.z = fq2_zero();
        return r;
    }
    fq2 zz2 = fq2_sqr(a.zz);
    r.x = fq2_mul(a.x, zz2);
    r.y = fq2_mul(a.y, fq2_mul(zz2, a.zz));
    r.z = a.zzz;
    return r;
}
