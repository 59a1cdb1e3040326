// b200math.cuh -- DEVICE-side BN254 math library (see b200math.h).  Header-only: include it in a .cu file compiled with
//     nvcc -gencode arch=compute_100a,code=sm_100a -I<repo>/include ...
// and call the functions below from your own kernels.  They are the ones libb200msm.so's MSM kernels are built from.
//
//   F_q   (fq.cuh)      struct fq {uint32_t v[8]}: Montgomery form, R = 2^256, always fully reduced (< p)
//         fq_mul  fq_sqr  fq_add  fq_sub  fq_dbl  fq_neg  fq_cneg  fq_is_zero  fq_eq  fq_zero  fq_one
//         fq_mulsub(a, b, c, d) = a*b - c*d, fq_muladd(a, b, c, d) = a*b + c*d, each with ONE Montgomery reduction
//         fq_inv (Fermat)   fq_inv_by (safegcd, fq_inv.cuh: ~6x cheaper)
//         fq_load / fq_load_nc / fq_store   (32-byte records, 16-byte aligned)
//   F_r   (fq.cuh)      fr_from_mont(uint32_t (&t)[8])    Montgomery -> canonical scalar
//   G1    (g1.cuh)      affine_t {x, y} (64 B; (0,0) = infinity marker), xyzz_t {x, y, zz, zzz} (128 B), jac_t {x, y, z} (96 B)
//         xyzz_madd(acc, p)  xyzz_add(acc, b)  xyzz_dbl_inplace(a)  xyzz_dbl_affine(p)  xyzz_neg  xyzz_inf  xyzz_is_inf
//         xyzz_from_affine  xyzz_to_jacobian  xyzz_from_jacobian  jac_dbl_inplace   -- all complete (P+P, P+(-P), infinity)
//   F_q2 / G2 (g2.cuh)  fq2 {c0, c1}; fq2_mul  fq2_sqr  fq2_add  fq2_sub  fq2_neg  fq2_inv;
//         g2_affine_t (128 B), g2_xyzz_t (256 B), g2_jac_t (192 B); g2_madd  g2_add  g2_dbl  g2_to_jacobian
//
// Replaces the reference's shader/{bigint,field,mont_backend,curve,misc}/*.metal (16 x 16-bit limbs, Jacobian only).
#pragma once
#include "../gpu-acceleration_b200/csrc/fq.cuh"
#include "../gpu-acceleration_b200/csrc/fq_inv.cuh"
#include "../gpu-acceleration_b200/csrc/g1.cuh"
#include "../gpu-acceleration_b200/csrc/g2.cuh"
