/* b200math.h -- the stand-alone BN254 math library of libb200msm.so.
 *
 * The reference keeps a device math library next to its MSM kernels
 *     /root/reference/mopro-msm/src/msm/metal_msm/shader/{bigint,field,mont_backend,curve,misc}/   (16 x 16-bit limbs)
 * and tests every level of it with single-purpose kernels
 *     /root/reference/mopro-msm/src/msm/metal_msm/tests/{bigint,field,mont_backend,curve}/         (one thread, one op)
 * (its README names "a crypto math library" as a goal of its own, README.md:199-200).  This header is the same thing for
 * the B200 engine, in two forms:
 *
 *   1. DEVICE side: `#include "b200math.cuh"` in your own .cu file (compile for sm_100a) and call the functions the MSM
 *      kernels are built from -- F_q (8 x 32-bit limbs, Montgomery, R = 2^256: arkworks' in-memory `Fq`), F_r decode,
 *      G1 in XYZZ / Jacobian coordinates, F_q2 and G2.  Example: gpu-acceleration_b200/cpp/example_math.cu.
 *   2. HOST side: b200math_apply() runs one operation element-wise over host arrays through those same device functions
 *      (one thread per element) -- what the reference's `test_*` kernels do, and what tests/test_gpu_pyramid.py and
 *      tests/test_gpu_g2.py use to check every level bit-exactly against the oracle.
 *
 * All field elements are 4 little-endian u64 in Montgomery form, fully reduced (< p), exactly arkworks' memory.
 */
#ifndef B200MATH_H
#define B200MATH_H

#include "b200msm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* record sizes: Fq / Fr 32 B; Fq2 64 B (c0 | c1); G1 affine 64 B (x | y); G1 XYZZ 128 B (X | Y | ZZ | ZZZ, x = X/ZZ,
 * y = Y/ZZZ, infinity <=> ZZ = 0); G1 Jacobian 96 B (X | Y | Z: arkworks `G1Projective`); G2 affine 128 B; G2 XYZZ 256 B */
typedef enum b200math_op {
    /* ---- F_q          replaces mont_backend/mont.metal:105-181 (mont_mul_cios), field/ff.metal:9-35, bigint/bigint.metal:7-178 */
    B200MATH_FQ_MUL = 0,         /* out = a * b                          (a, b, out: 32 B)                                */
    B200MATH_FQ_ADD = 1,         /* out = a + b                          ff_add                                          */
    B200MATH_FQ_SUB = 2,         /* out = a - b                          ff_sub                                          */
    B200MATH_FQ_SQR = 3,         /* out = a^2                                                                            */
    B200MATH_FQ_NEG = 4,         /* out = p - a, 0 -> 0                  jacobian_neg's y <- p - y (curve/jacobian.metal:195-210) */
    B200MATH_FQ_INV = 5,         /* out = a^(p-2) (Fermat), 0 -> 0                                                       */
    B200MATH_FQ_DBL = 6,         /* out = 2a                                                                             */
    B200MATH_FQ_INV_SAFEGCD = 7, /* out = a^-1 by Bernstein-Yang divsteps (fq_inv.cuh), 0 -> 0                            */
    B200MATH_FQ_MULSUB = 8,      /* out = a*b - c*d with one reduction; a: 64 B [a | c], b: 64 B [b | d], out: 32 B       */
    /* ---- G1           replaces curve/jacobian.metal:11-226 (dbl-2009-l, add-2007-bl, madd-2007-bl, neg, scalar_mul)     */
    B200MATH_G1_XYZZ_MADD = 10,  /* out = a + b, a: XYZZ 128 B, b: affine 64 B (complete: P+P, P+(-P), infinity)          */
    B200MATH_G1_XYZZ_ADD = 11,   /* out = a + b, both XYZZ                                                               */
    B200MATH_G1_XYZZ_DBL = 12,   /* out = 2a                                                                             */
    B200MATH_G1_XYZZ_TO_JACOBIAN = 13, /* a: 128 B -> out: 96 B; infinity -> (R, R, 0) as arkworks builds it             */
    B200MATH_G1_JAC_DBL = 14,    /* dbl-2009-l on a finite Jacobian point (a, out: 96 B)                                  */
    B200MATH_G1_SCALAR_MUL_U32 = 15, /* out = k * a, a: affine 64 B, b: 8 B holding a 32-bit k, out: XYZZ                 */
    /* ---- F_r          replaces the CPU `into_bigint()` of utils/limbs_conversion.rs:311-378                            */
    B200MATH_FR_FROM_MONT = 20,  /* Montgomery -> canonical scalar (a, out: 32 B)                                         */
    /* ---- F_q2 / G2    (nothing to replace: the reference has no G2)                                                    */
    B200MATH_FQ2_MUL = 30,       /* a, b, out: 64 B                                                                       */
    B200MATH_FQ2_SQR = 31,
    B200MATH_G2_XYZZ_MADD = 32,  /* a: 256 B XYZZ over Fq2, b: 128 B affine                                               */
    B200MATH_G2_XYZZ_ADD = 33,
    B200MATH_G2_XYZZ_DBL = 34
} b200math_op;

/* out[i] = op(a[i], b[i]) for i < count, on the context's first device; b may be NULL for unary operations.
 * Blocking; host arrays (any memory kind).  Returns 0 or a negative B200MSM_E* code (b200msm_last_error()).           */
int b200math_apply(b200msm_ctx* ctx, b200math_op op, const void* a, const void* b, void* out, size_t count);

#ifdef __cplusplus
}
#endif
#endif /* B200MATH_H */
