// BN254 G1 (y^2 = x^3 + 3) point arithmetic for the bucket kernels.
//
// Accumulators use XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; infinity <=> ZZ == 0):
// a mixed addition is 8M + 2S (EFD madd-2008-s) against the 16 multiplications of the
// reference's general Jacobian add with z2 = R
// (/root/reference/mopro-msm/src/msm/metal_msm/shader/curve/jacobian.metal:46-100, called from
// shader/cuzk/smvp.metal:61-71).  Every operation is COMPLETE: infinity operands, P + P and
// P + (-P) are detected by value (not by limb-equality of one representation, the gap noted in
// SURVEY.md §2.3 item 2), because all field values are kept canonical in [0, p).
#pragma once
#include "fq.cuh"

struct affine_t {  // 64 bytes, the device-resident base format: x || y, Montgomery LE; (0,0) never occurs on-curve
    fq x, y;
};
struct xyzz_t {  // 128 bytes
    fq x, y, zz, zzz;
};
struct jac_t {  // 96 bytes: arkworks G1Projective memory order (x, y, z)
    fq x, y, z;
};

__device__ __forceinline__ xyzz_t xyzz_inf() {
    xyzz_t r;
    r.x = fq_zero(); r.y = fq_zero(); r.zz = fq_zero(); r.zzz = fq_zero();
    return r;
}
__device__ __forceinline__ bool xyzz_is_inf(const xyzz_t& a) { return fq_is_zero(a.zz); }

__device__ __forceinline__ xyzz_t xyzz_from_affine(const affine_t& p) {
    xyzz_t r;
    r.x = p.x; r.y = p.y; r.zz = fq_one(); r.zzz = fq_one();
    return r;
}

// EFD dbl-2008-s-1 with a = 0: 6M + 3S
__device__ __noinline__ void xyzz_dbl_inplace_impl(xyzz_t& a) {
    if (xyzz_is_inf(a)) return;
    fq U = fq_dbl(a.y);
    fq V = fq_sqr(U);
    fq W = fq_mul(U, V);
    fq S = fq_mul(a.x, V);
    fq M = fq_sqr(a.x);
    M = fq_add(fq_dbl(M), M);
    fq X3 = fq_sub(fq_sub(fq_sqr(M), S), S);
    fq Y3 = fq_mulsub(M, fq_sub(S, X3), W, a.y);
    a.x = X3; a.y = Y3;
    a.zz = fq_mul(V, a.zz);
    a.zzz = fq_mul(W, a.zzz);
}

// mdbl-2008-s-1: doubling of an affine point into XYZZ (4M... here 3M + 3S + W*y)
__device__ __noinline__ xyzz_t xyzz_dbl_affine(const affine_t& p) {
    xyzz_t r;
    fq U = fq_dbl(p.y);
    r.zz = fq_sqr(U);
    r.zzz = fq_mul(U, r.zz);
    fq S = fq_mul(p.x, r.zz);
    fq M = fq_sqr(p.x);
    M = fq_add(fq_dbl(M), M);
    r.x = fq_sub(fq_sub(fq_sqr(M), S), S);
    r.y = fq_mulsub(M, fq_sub(S, r.x), r.zzz, p.y);
    return r;
}

// acc += p  (p affine, never infinity).  EFD madd-2008-s, 8M + 2S, complete.
__device__ __forceinline__ void xyzz_madd(xyzz_t& acc, const affine_t& p) {
    if (xyzz_is_inf(acc)) {
        acc = xyzz_from_affine(p);
        return;
    }
    fq U2 = fq_mul(p.x, acc.zz);
    fq S2 = fq_mul(p.y, acc.zzz);
    fq P = fq_sub(U2, acc.x);
    fq R = fq_sub(S2, acc.y);
    if (fq_is_zero(P)) {  // same x: doubling or cancellation (rare; kept out of line)
        if (fq_is_zero(R)) acc = xyzz_dbl_affine(p);
        else acc = xyzz_inf();
        return;
    }
    fq PP = fq_sqr(P);
    fq PPP = fq_mul(P, PP);
    fq Q = fq_mul(acc.x, PP);
    fq X3 = fq_sub(fq_sub(fq_sub(fq_sqr(R), PPP), Q), Q);
    fq Y3 = fq_mulsub(R, fq_sub(Q, X3), acc.y, PPP);
    acc.x = X3; acc.y = Y3;
    acc.zz = fq_mul(acc.zz, PP);
    acc.zzz = fq_mul(acc.zzz, PPP);
}

// acc += b (both XYZZ).  EFD add-2008-s, 12M + 2S, complete.
__device__ __noinline__ void xyzz_add_impl(xyzz_t& acc, const xyzz_t& b) {
    if (xyzz_is_inf(b)) return;
    if (xyzz_is_inf(acc)) { acc = b; return; }
    fq U1 = fq_mul(acc.x, b.zz);
    fq U2 = fq_mul(b.x, acc.zz);
    fq S1 = fq_mul(acc.y, b.zzz);
    fq S2 = fq_mul(b.y, acc.zzz);
    fq P = fq_sub(U2, U1);
    fq R = fq_sub(S2, S1);
    if (fq_is_zero(P)) {
        if (fq_is_zero(R)) xyzz_dbl_inplace_impl(acc);
        else acc = xyzz_inf();
        return;
    }
    fq PP = fq_sqr(P);
    fq PPP = fq_mul(P, PP);
    fq Q = fq_mul(U1, PP);
    fq X3 = fq_sub(fq_sub(fq_sub(fq_sqr(R), PPP), Q), Q);
    fq Y3 = fq_mulsub(R, fq_sub(Q, X3), S1, PPP);
    acc.x = X3; acc.y = Y3;
    acc.zz = fq_mul(fq_mul(acc.zz, b.zz), PP);
    acc.zzz = fq_mul(fq_mul(acc.zzz, b.zzz), PPP);
}

// Out-of-line bodies are reached through these wrappers.  The empty asm with a "memory" clobber is
// required: without it cicc (CUDA 12.9) drops the reload of word 0 of the accumulator after the call
// (observed: `st.shared.v4 {%undef, ...}` in k_bucket_reduce), i.e. it treats the callee as not writing it.
__device__ __forceinline__ void xyzz_add(xyzz_t& acc, const xyzz_t& b) {
    xyzz_add_impl(acc, b);
    asm volatile("" ::: "memory");
}
__device__ __forceinline__ void xyzz_dbl_inplace(xyzz_t& a) {
    xyzz_dbl_inplace_impl(a);
    asm volatile("" ::: "memory");
}

__device__ __forceinline__ xyzz_t xyzz_neg(const xyzz_t& a) {
    xyzz_t r = a;
    r.y = fq_neg(a.y);
    return r;
}

// (X*ZZ^2, Y*ZZ^3, ZZZ) is a Jacobian representative of the same point; infinity -> (1, 1, 0)
// as arkworks' `Projective::zero()` builds it.
__device__ __forceinline__ jac_t xyzz_to_jacobian(const xyzz_t& a) {
    jac_t r;
    if (xyzz_is_inf(a)) {
        r.x = fq_one(); r.y = fq_one(); r.z = fq_zero();
        return r;
    }
    fq zz2 = fq_sqr(a.zz);
    r.x = fq_mul(a.x, zz2);
    r.y = fq_mul(a.y, fq_mul(zz2, a.zz));
    r.z = a.zzz;
    return r;
}
__device__ __forceinline__ xyzz_t xyzz_from_jacobian(const jac_t& a) {
    xyzz_t r;
    if (fq_is_zero(a.z)) return xyzz_inf();
    r.x = a.x; r.y = a.y;
    r.zz = fq_sqr(a.z);
    r.zzz = fq_mul(r.zz, a.z);
    return r;
}

// EFD dbl-2009-l (a = 0) on a non-infinity Jacobian point: 2M + 5S, dependency depth 3.
__device__ __forceinline__ void jac_dbl_inplace(jac_t& p) {
    fq A = fq_sqr(p.x);
    fq B = fq_sqr(p.y);
    fq Z3 = fq_mul(p.y, p.z);
    fq C = fq_sqr(B);
    fq t = fq_sqr(fq_add(p.x, B));
    fq E = fq_add(fq_dbl(A), A);
    fq F = fq_sqr(E);
    fq D = fq_dbl(fq_sub(fq_sub(t, A), C));
    fq X3 = fq_sub(fq_sub(F, D), D);
    fq C8 = fq_dbl(fq_dbl(fq_dbl(C)));
    p.y = fq_sub(fq_mul(E, fq_sub(D, X3)), C8);
    p.x = X3;
    p.z = fq_dbl(Z3);
}

__device__ __forceinline__ xyzz_t xyzz_load(const void* p) {
    const char* c = reinterpret_cast<const char*>(p);
    xyzz_t r;
    r.x = fq_load(c); r.y = fq_load(c + 32); r.zz = fq_load(c + 64); r.zzz = fq_load(c + 96);
    return r;
}
__device__ __forceinline__ void xyzz_store(void* p, const xyzz_t& a) {
    char* c = reinterpret_cast<char*>(p);
    fq_store(c, a.x); fq_store(c + 32, a.y); fq_store(c + 64, a.zz); fq_store(c + 96, a.zzz);
}
__device__ __forceinline__ affine_t affine_load_nc(const void* p) {
    const char* c = reinterpret_cast<const char*>(p);
    affine_t r;
    r.x = fq_load_nc(c); r.y = fq_load_nc(c + 32);
    return r;
}
