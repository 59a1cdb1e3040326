// BN254 G2 MSM kernels: the G1 pipeline's skeleton (msm_kernels.cuh) with Fq2 points.
//   K1 / K2 (scalars -> sorted entries) are shared with G1 unchanged: they never look at a point.
//   K3  k_g2_accumulate + k_g2_fixup   fixed chunks of L sorted entries per thread, XYZZ over Fq2 (madd = 8M + 2S in Fq2 =
//                                     28 Fq products), 128-byte gathers
//   K4  k_g2_reduce_level             the lane-parallel cooperative engine of the G1 reduce over Fq2 (latency-bound sizes), or
//       k_g2_bucket_reduce + k_g2_window_finish   thread-per-segment running sums + shared-memory suffix scan / tree sums
//   K5  k_g2_combine                  Horner over the windows (cooperative Jacobian doublings on four warps)
// The scalar split of K1 (GLV) applies unchanged: phi acts on G2 through beta^2.
#pragma once
#include "g2.cuh"

#define G2_ACC_THREADS 128
#define G2_RED_THREADS 64

// 16 threads per point, one u64 each: raw caller records -> 128-byte device records; infinity -> all zero
__global__ void __launch_bounds__(256) k_g2_repack(const uint8_t* __restrict__ raw, size_t stride, size_t x_off, size_t y_off,
                                                   size_t inf_off, uint32_t n, uint64_t* __restrict__ out) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t i = gid >> 4;
    if (i >= n) return;
    const uint32_t k = (uint32_t)(gid & 15);
    const uint8_t* rec = raw + i * stride;
    uint64_t v = *reinterpret_cast<const uint64_t*>(rec + (k < 8 ? x_off + 8 * k : y_off + 8 * (k - 8)));
    if (inf_off != (size_t)-1 && rec[inf_off] != 0) v = 0;
    out[i * 16 + k] = v;
}

// GLV on G2: the curve y^2 = x^3 + b' has j = 0 too, so (x, y) -> (beta' x, y) with beta' a cube root of unity of Fq is an
// endomorphism; on the order-r subgroup it acts as multiplication by lambda (the eigenvalue the scalar split of K1 uses)
// for beta' = beta^2, beta being G1's constant (pinned by tests/test_oracle_g2.py).  Pseudo-point index >= n means
// phi(P_{i-n}): two extra Fq products on the gathered x, no second base array.
__global__ void __launch_bounds__(G2_ACC_THREADS) k_g2_accumulate(const g2_affine_t* __restrict__ bases, uint32_t n,
                                                                 const uint32_t* __restrict__ entries,
                                                                 const uint32_t* __restrict__ ends, uint32_t G, uint32_t L,
                                                                 g2_xyzz_t* __restrict__ buckets, g2_xyzz_t* __restrict__ head,
                                                                 g2_xyzz_t* __restrict__ tail, int into) {
    // into: slices 1.. of a sliced MSM add into the bucket array of the earlier slices (see k_accumulate)
    const uint32_t P1 = ends[G - 1];
    const uint64_t t64 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t clo64 = t64 * L;
    if (clo64 >= P1) return;
    const uint32_t t = (uint32_t)t64;
    const uint32_t clo = (uint32_t)clo64;
    const uint32_t chi = (uint32_t)min((uint64_t)0xffffffffu, clo64 + L);
    const uint32_t hi = min(chi, P1);
    // smallest g with ends[g] > clo
    uint32_t a = 0, b = G - 1;
    while (a < b) {
        uint32_t mid = (a + b) >> 1;
        if (__ldg(ends + mid) > clo) b = mid; else a = mid + 1;
    }
    uint32_t g = a;
    uint32_t bstart = g ? __ldg(ends + g - 1) : 0;
    uint32_t bend = __ldg(ends + g);
    g2_xyzz_t acc = g2_inf();
    if (into && bstart >= clo) acc = g2_load(buckets + g);
    uint32_t e_next = __ldg(entries + clo);
    for (uint32_t pos = clo; pos < hi; pos++) {
        const uint32_t e = e_next;
        if (pos + 1 < hi) e_next = __ldg(entries + pos + 1);
        if (pos >= bend) {
            g2_store((bstart >= clo) ? buckets + g : head + t, acc);
            do { g++; bstart = bend; bend = __ldg(ends + g); } while (bend <= pos);
            acc = g2_inf();
            if (into) acc = g2_load(buckets + g);
        }
        const uint32_t idx = e & 0x7fffffffu;
        const bool endo = idx >= n;
        g2_affine_t p = g2_affine_load_nc(bases + (endo ? idx - n : idx));
        if (!g2_affine_is_inf(p)) {
            if (endo) {
                const fq beta2 = {{0x13e80b9cu, 0x3350c88eu, 0xdb5e56b9u, 0x7dce557cu, 0xb615564au, 0x6001b4b8u, 0x020217e0u, 0x2682e617u}};
                p.x.c0 = fq_mul(p.x.c0, beta2);
                p.x.c1 = fq_mul(p.x.c1, beta2);
            }
            p.y = fq2_cneg(p.y, (e >> 31) != 0);
            g2_madd(acc, p);
        }
    }
    g2_xyzz_t* dst;
    if (bstart < clo) dst = head + t;
    else if (bend > chi) dst = tail + t;
    else dst = buckets + g;
    g2_store(dst, acc);
}

// One thread per global bucket: empty -> infinity; straddling -> tail[first chunk] + head[following chunks].  Buckets that
// span more than FIX_LONG chunks (skewed scalars; a narrow top window) are queued for k_g2_fixup_long.
__global__ void __launch_bounds__(128) k_g2_fixup(const uint32_t* __restrict__ ends, uint32_t G, uint32_t L,
                                                  g2_xyzz_t* __restrict__ buckets, const g2_xyzz_t* __restrict__ head,
                                                  const g2_xyzz_t* __restrict__ tail, uint32_t* __restrict__ long_count,
                                                  uint32_t* __restrict__ long_list, int keep_empty) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const uint32_t start = g ? ends[g - 1] : 0, end = ends[g];
    if (start == end) {
        if (!keep_empty) g2_store(buckets + g, g2_inf());
        return;
    }
    const uint32_t t0 = start / L, t1 = (end - 1) / L;
    if (t0 == t1) return;
    if (t1 - t0 > FIX_LONG) {
        long_list[atomicAdd(long_count, 1u)] = g;
        return;
    }
    g2_xyzz_t acc = g2_load(tail + t0);
    for (uint32_t t = t0 + 1; t <= t1; t++) {
        g2_xyzz_t h = g2_load(head + t);
        g2_add(acc, h);
    }
    g2_store(buckets + g, acc);
}

// A whole CTA per queued bucket: strided partial sums, then a shared-memory tree.
#define G2_FIXL_THREADS 128
__global__ void __launch_bounds__(G2_FIXL_THREADS) k_g2_fixup_long(const uint32_t* __restrict__ ends, uint32_t L,
                                                                  g2_xyzz_t* __restrict__ buckets,
                                                                  const g2_xyzz_t* __restrict__ head,
                                                                  const g2_xyzz_t* __restrict__ tail,
                                                                  const uint32_t* __restrict__ long_count,
                                                                  const uint32_t* __restrict__ long_list) {
    __shared__ uint4 sm[G2_FIXL_THREADS * 16];
    g2_xyzz_t* s = reinterpret_cast<g2_xyzz_t*>(sm);
    const uint32_t count = *long_count;
    for (uint32_t item = blockIdx.x; item < count; item += gridDim.x) {
        const uint32_t g = long_list[item];
        const uint32_t start = g ? ends[g - 1] : 0, end = ends[g];
        const uint32_t t0 = start / L, t1 = (end - 1) / L;
        g2_xyzz_t acc = g2_inf();
        for (uint32_t t = t0 + 1 + threadIdx.x; t <= t1; t += G2_FIXL_THREADS) {
            g2_xyzz_t h = g2_load(head + t);
            g2_add(acc, h);
        }
        if (threadIdx.x == 0) {
            g2_xyzz_t h = g2_load(tail + t0);
            g2_add(acc, h);
        }
        g2_store(s + threadIdx.x, acc);
        __syncthreads();
        for (int stride = G2_FIXL_THREADS / 2; stride > 0; stride >>= 1) {
            if (threadIdx.x < stride) {
                g2_xyzz_t x = g2_load(s + threadIdx.x), y = g2_load(s + threadIdx.x + stride);
                g2_add(x, y);
                g2_store(s + threadIdx.x, x);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) g2_store(buckets + g, g2_load(s));
        __syncthreads();
    }
}

// Sliced host-input G2 MSM: bucket g of the MSM is the sum over the slices' bucket arrays (see k_merge_buckets).
struct g2_merge_srcs {
    const g2_xyzz_t* p[7];
};
__global__ void __launch_bounds__(128) k_g2_merge_buckets(g2_xyzz_t* __restrict__ dst, g2_merge_srcs src, int count, uint32_t G) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    g2_xyzz_t acc = g2_load(dst + g);
    for (int k = 0; k < count; k++) {
        g2_xyzz_t b = g2_load(src.p[k] + g);
        g2_add(acc, b);
    }
    g2_store(dst + g, acc);
}

// In: per-thread (run, tot) BY VALUE (see g1.cuh on by-reference accumulators).  Out: sA[0] = sum run,
// sB[0] = sum tot + 2^log2w * sum_t t * run_t.  sA, sB: N slots, sC: 1 slot.  All N threads must call.
template <int N>
__device__ __forceinline__ void g2_block_weighted_sum(g2_xyzz_t* sA, g2_xyzz_t* sB, g2_xyzz_t* sC, g2_xyzz_t run, g2_xyzz_t tot,
                                                      int log2w) {
    const int t = threadIdx.x;
    g2_store(sA + t, run);
    g2_store(sB + t, tot);
    __syncthreads();
    for (int d = 1; d < N; d <<= 1) {   // inclusive suffix scan of sA
        g2_xyzz_t v = g2_load(sA + t);
        if (t + d < N) {
            g2_xyzz_t u = g2_load(sA + t + d);
            g2_add(v, u);
        }
        __syncthreads();
        g2_store(sA + t, v);
        __syncthreads();
    }
    // every thread weights its own share first -- V_t = tot_t + 2^log2w * suffix[t + 1], and sum_t suffix[t + 1] is
    // sum_t t * run_t -- so that ONE tree sum finishes the job (the doublings run in parallel across the threads)
    {
        g2_xyzz_t q = g2_inf();
        if (t + 1 < N) q = g2_load(sA + t + 1);
        for (int k = 0; k < log2w; k++) g2_dbl(q);
        g2_xyzz_t v = g2_load(sB + t);
        g2_add(v, q);
        __syncthreads();
        g2_store(sB + t, v);
    }
    __syncthreads();
    for (int stride = N / 2; stride > 0; stride >>= 1) {
        if (t < stride) {
            g2_xyzz_t x = g2_load(sB + t), y = g2_load(sB + t + stride);
            g2_add(x, y);
            g2_store(sB + t, x);
        }
        __syncthreads();
    }
    (void)sC;
}

__global__ void __launch_bounds__(G2_RED_THREADS) k_g2_bucket_reduce(const g2_xyzz_t* __restrict__ buckets, uint32_t nb,
                                                                    uint32_t log2Bsz, uint32_t blocks_per_window,
                                                                    g2_xyzz_t* __restrict__ wpartR,
                                                                    g2_xyzz_t* __restrict__ wpartT) {
    __shared__ uint4 smA[G2_RED_THREADS * 16];
    __shared__ uint4 smB[G2_RED_THREADS * 16];
    __shared__ uint4 smC[16];
    const uint32_t half = nb - 1;
    const uint32_t Bsz = 1u << log2Bsz;
    const uint32_t w = blockIdx.x / blocks_per_window;
    const uint32_t j = (blockIdx.x % blocks_per_window) * G2_RED_THREADS + threadIdx.x;
    const uint64_t lo64 = (uint64_t)j << log2Bsz;
    g2_xyzz_t tot = g2_inf(), run = g2_inf();
    if (lo64 < half) {
        const uint32_t lo = (uint32_t)lo64;
        const uint32_t top = min(half, lo + Bsz);
        const g2_xyzz_t* B = buckets + (size_t)w * nb;
        for (uint32_t m = top; m > lo; m--) {
            g2_xyzz_t v = g2_load(B + m);
            g2_add(run, v);
            g2_add(tot, run);
        }
    }
    g2_block_weighted_sum<G2_RED_THREADS>(reinterpret_cast<g2_xyzz_t*>(smA), reinterpret_cast<g2_xyzz_t*>(smB),
                                          reinterpret_cast<g2_xyzz_t*>(smC), run, tot, (int)log2Bsz);
    if (threadIdx.x == 0) {
        g2_store(wpartR + blockIdx.x, g2_load(smA));
        g2_store(wpartT + blockIdx.x, g2_load(smB));
    }
}

// One 64-thread CTA per window: thread b holds CTA b's (R_b, T_b) (blocks_per_window <= 64)
__global__ void __launch_bounds__(G2_RED_THREADS) k_g2_window_finish(const g2_xyzz_t* __restrict__ wpartR,
                                                                    const g2_xyzz_t* __restrict__ wpartT,
                                                                    uint32_t blocks_per_window, uint32_t log2weight,
                                                                    g2_xyzz_t* __restrict__ wsum) {
    __shared__ uint4 smA[G2_RED_THREADS * 16];
    __shared__ uint4 smB[G2_RED_THREADS * 16];
    __shared__ uint4 smC[16];
    const uint32_t w = blockIdx.x;
    g2_xyzz_t run = g2_inf(), tot = g2_inf();
    if (threadIdx.x < blocks_per_window) {
        run = g2_load(wpartR + (size_t)w * blocks_per_window + threadIdx.x);
        tot = g2_load(wpartT + (size_t)w * blocks_per_window + threadIdx.x);
    }
    g2_block_weighted_sum<G2_RED_THREADS>(reinterpret_cast<g2_xyzz_t*>(smA), reinterpret_cast<g2_xyzz_t*>(smB),
                                          reinterpret_cast<g2_xyzz_t*>(smC), run, tot, (int)log2weight);
    if (threadIdx.x == 0) g2_store(wsum + w, g2_load(smB));
}

// ------------------------------------------------------------------------------------------ K4 (cooperative, Fq2)
// The lane-parallel cooperative engine of msm_kernels.cuh (k_reduce_level) over Fq2: a CTA of 4 warps serves 32 independent
// chains (one per lane); in each phase warp w computes ONE of the (up to four) independent Fq2 products of the operation for
// all 32 lanes, values exchanged through shared memory laid out [slot][limb][lane] (16 limbs per Fq2, conflict-free).  An
// XYZZ addition is 4 phases of one Fq2 product each instead of 14 in a row on a lone thread.  Same recursion as G1
// (see reduce_level_cta): level l turns `cnt` items per window into ceil(cnt / (32 * 2^lb)) items, until one CTA is left.
// 26 slots x 16 limbs x 32 lanes x 4 B = 52 KB of dynamic shared memory: 4 CTAs per SM (the G1 engine's CL_SAVE group is
// folded into the scratch group, which is free by then).
#define G2CL_THREADS 128
#define G2CL_SLOTS 26
#define G2CL_SMEM_BYTES (G2CL_SLOTS * 16 * 32 * 4 + 32 * 4)
enum { G2CL_RUN = 0, G2CL_TOT = 4, G2CL_XS = 8, G2CL_B = 12, G2CL_T = 16 /* 10 temporaries */ };

__device__ __forceinline__ fq2 g2cl_ld(const uint32_t* sm, int slot, int lane) {
    fq2 r;
#pragma unroll
    for (int k = 0; k < 8; k++) r.c0.v[k] = sm[(slot * 16 + k) * 32 + lane];
#pragma unroll
    for (int k = 0; k < 8; k++) r.c1.v[k] = sm[(slot * 16 + 8 + k) * 32 + lane];
    return r;
}
__device__ __forceinline__ void g2cl_st(uint32_t* sm, int slot, int lane, const fq2& v) {
#pragma unroll
    for (int k = 0; k < 8; k++) sm[(slot * 16 + k) * 32 + lane] = v.c0.v[k];
#pragma unroll
    for (int k = 0; k < 8; k++) sm[(slot * 16 + 8 + k) * 32 + lane] = v.c1.v[k];
}
__device__ __forceinline__ bool g2cl_zero(const uint32_t* sm, int slot, int lane) {
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) o |= sm[(slot * 16 + k) * 32 + lane];
    return o == 0;
}

// acc(A) <- 2 * acc(A) for lanes with `on` (dbl-2008-s-1, a = 0).  All G2CL_THREADS threads call.
__device__ __noinline__ void g2cl_dbl(uint32_t* sm, int A, int warp, int lane, bool on) {
    on = on && !g2cl_zero(sm, A + 2, lane);
    if (on && warp == 0) { fq2 U = fq2_dbl(g2cl_ld(sm, A + 1, lane)); g2cl_st(sm, G2CL_T + 0, lane, U); g2cl_st(sm, G2CL_T + 1, lane, fq2_sqr(U)); }
    if (on && warp == 1) { fq2 a = fq2_sqr(g2cl_ld(sm, A + 0, lane)); g2cl_st(sm, G2CL_T + 2, lane, fq2_add(fq2_dbl(a), a)); }
    __syncthreads();
    if (on && warp == 0) g2cl_st(sm, G2CL_T + 3, lane, fq2_mul(g2cl_ld(sm, G2CL_T + 0, lane), g2cl_ld(sm, G2CL_T + 1, lane)));   // W = U V
    if (on && warp == 1) g2cl_st(sm, G2CL_T + 4, lane, fq2_mul(g2cl_ld(sm, A + 0, lane), g2cl_ld(sm, G2CL_T + 1, lane)));        // S = X V
    if (on && warp == 2) g2cl_st(sm, G2CL_T + 5, lane, fq2_sqr(g2cl_ld(sm, G2CL_T + 2, lane)));                                   // M^2
    __syncthreads();
    if (on && warp == 0) {
        fq2 S = g2cl_ld(sm, G2CL_T + 4, lane);
        fq2 X3 = fq2_sub(fq2_sub(g2cl_ld(sm, G2CL_T + 5, lane), S), S);
        g2cl_st(sm, G2CL_T + 6, lane, fq2_mul(g2cl_ld(sm, G2CL_T + 2, lane), fq2_sub(S, X3)));
        g2cl_st(sm, A + 0, lane, X3);
    }
    if (on && warp == 1) g2cl_st(sm, G2CL_T + 7, lane, fq2_mul(g2cl_ld(sm, G2CL_T + 3, lane), g2cl_ld(sm, A + 1, lane)));        // W Y
    if (on && warp == 2) g2cl_st(sm, A + 2, lane, fq2_mul(g2cl_ld(sm, G2CL_T + 1, lane), g2cl_ld(sm, A + 2, lane)));
    if (on && warp == 3) g2cl_st(sm, A + 3, lane, fq2_mul(g2cl_ld(sm, G2CL_T + 3, lane), g2cl_ld(sm, A + 3, lane)));
    __syncthreads();
    if (on && warp == 0) g2cl_st(sm, A + 1, lane, fq2_sub(g2cl_ld(sm, G2CL_T + 6, lane), g2cl_ld(sm, G2CL_T + 7, lane)));
    __syncthreads();
}

// acc(A) <- acc(A) + b(B) per lane (add-2008-s), complete: infinity operands, P + P, P + (-P).  All threads call.
__device__ __noinline__ void g2cl_add(uint32_t* sm, uint32_t* flags, int A, int B, int warp, int lane) {
    const bool b_inf = g2cl_zero(sm, B + 2, lane), a_inf = g2cl_zero(sm, A + 2, lane);
    const bool go = !b_inf && !a_inf;
    // phase 1: U1 = X1 ZZ2, U2 = X2 ZZ1, S1 = Y1 ZZZ2, S2 = Y2 ZZZ1
    if (go) {
        if (warp == 0) g2cl_st(sm, G2CL_T + 0, lane, fq2_mul(g2cl_ld(sm, A + 0, lane), g2cl_ld(sm, B + 2, lane)));
        if (warp == 1) g2cl_st(sm, G2CL_T + 1, lane, fq2_mul(g2cl_ld(sm, B + 0, lane), g2cl_ld(sm, A + 2, lane)));
        if (warp == 2) g2cl_st(sm, G2CL_T + 2, lane, fq2_mul(g2cl_ld(sm, A + 1, lane), g2cl_ld(sm, B + 3, lane)));
        if (warp == 3) g2cl_st(sm, G2CL_T + 3, lane, fq2_mul(g2cl_ld(sm, B + 1, lane), g2cl_ld(sm, A + 3, lane)));
    }
    __syncthreads();
    // phase 2: P, R, classification, PP, RR, ZZ1 ZZ2, ZZZ1 ZZZ2
    if (warp == 0) {
        uint32_t f = b_inf ? 1u : a_inf ? 2u : 0u;   // 1 keep acc, 2 copy b, 3 result infinity, 4 double
        if (go) {
            fq2 P = fq2_sub(g2cl_ld(sm, G2CL_T + 1, lane), g2cl_ld(sm, G2CL_T + 0, lane));
            if (fq2_is_zero(P)) {
                fq2 R = fq2_sub(g2cl_ld(sm, G2CL_T + 3, lane), g2cl_ld(sm, G2CL_T + 2, lane));
                f = fq2_is_zero(R) ? 4u : 3u;
            } else {
                g2cl_st(sm, G2CL_T + 4, lane, P);
                g2cl_st(sm, G2CL_T + 6, lane, fq2_sqr(P));
            }
        }
        flags[lane] = f;
    }
    if (go && warp == 1) {
        fq2 R = fq2_sub(g2cl_ld(sm, G2CL_T + 3, lane), g2cl_ld(sm, G2CL_T + 2, lane));
        g2cl_st(sm, G2CL_T + 5, lane, R);
        g2cl_st(sm, G2CL_T + 7, lane, fq2_sqr(R));
    }
    if (go && warp == 2) g2cl_st(sm, G2CL_T + 8, lane, fq2_mul(g2cl_ld(sm, A + 2, lane), g2cl_ld(sm, B + 2, lane)));
    if (go && warp == 3) g2cl_st(sm, G2CL_T + 9, lane, fq2_mul(g2cl_ld(sm, A + 3, lane), g2cl_ld(sm, B + 3, lane)));
    __syncthreads();
    const uint32_t f = flags[lane];
    // phase 3: PPP = P PP (-> T1), Q = U1 PP (-> T3), ZZ' = ZZ12 PP
    if (f == 0) {
        if (warp == 0) g2cl_st(sm, G2CL_T + 1, lane, fq2_mul(g2cl_ld(sm, G2CL_T + 4, lane), g2cl_ld(sm, G2CL_T + 6, lane)));
        if (warp == 1) g2cl_st(sm, G2CL_T + 3, lane, fq2_mul(g2cl_ld(sm, G2CL_T + 0, lane), g2cl_ld(sm, G2CL_T + 6, lane)));
        if (warp == 2) g2cl_st(sm, G2CL_T + 8, lane, fq2_mul(g2cl_ld(sm, G2CL_T + 8, lane), g2cl_ld(sm, G2CL_T + 6, lane)));
    }
    __syncthreads();
    // phase 4: X3 = RR - PPP - 2Q, t = R (Q - X3) (-> T4), u = S1 PPP (-> T6), ZZZ' = ZZZ12 PPP
    if (f == 0) {
        if (warp == 0) {
            fq2 Q = g2cl_ld(sm, G2CL_T + 3, lane);
            fq2 X3 = fq2_sub(fq2_sub(fq2_sub(g2cl_ld(sm, G2CL_T + 7, lane), g2cl_ld(sm, G2CL_T + 1, lane)), Q), Q);
            g2cl_st(sm, G2CL_T + 4, lane, fq2_mul(g2cl_ld(sm, G2CL_T + 5, lane), fq2_sub(Q, X3)));
            g2cl_st(sm, A + 0, lane, X3);
        }
        if (warp == 1) g2cl_st(sm, G2CL_T + 6, lane, fq2_mul(g2cl_ld(sm, G2CL_T + 2, lane), g2cl_ld(sm, G2CL_T + 1, lane)));
        if (warp == 2) g2cl_st(sm, G2CL_T + 9, lane, fq2_mul(g2cl_ld(sm, G2CL_T + 9, lane), g2cl_ld(sm, G2CL_T + 1, lane)));
    }
    __syncthreads();
    // commit: each warp owns one coordinate
    if (f == 0) {
        if (warp == 1) g2cl_st(sm, A + 1, lane, fq2_sub(g2cl_ld(sm, G2CL_T + 4, lane), g2cl_ld(sm, G2CL_T + 6, lane)));
        if (warp == 2) g2cl_st(sm, A + 2, lane, g2cl_ld(sm, G2CL_T + 8, lane));
        if (warp == 3) g2cl_st(sm, A + 3, lane, g2cl_ld(sm, G2CL_T + 9, lane));
    } else if (f == 2) {
        g2cl_st(sm, A + warp, lane, g2cl_ld(sm, B + warp, lane));
    } else if (f == 3) {
        g2cl_st(sm, A + warp, lane, fq2_zero());
    }
    const int any_dbl = __syncthreads_or(f == 4);
    if (any_dbl) g2cl_dbl(sm, A, warp, lane, f == 4);
}

// copy group S -> group D with a lane shift: D[lane] = S[lane + shift] (infinity beyond lane 31 or when !take)
__device__ __noinline__ void g2cl_shift_copy(uint32_t* sm, int D, int S, int shift, bool take, int warp, int lane) {
    fq2 v = fq2_zero();
    if (take && lane + shift < 32) v = g2cl_ld(sm, S + warp, lane + shift);
    g2cl_st(sm, D + warp, lane, v);
    __syncthreads();
}

// One level of the recursive weighted sum over G2 buckets; parameters as in k_reduce_level (msm_kernels.cuh).
__global__ void __launch_bounds__(G2CL_THREADS) k_g2_reduce_level(const g2_xyzz_t* __restrict__ A_in, const g2_xyzz_t* __restrict__ X_in,
                                                                  uint32_t in_stride, uint32_t in_off, uint32_t cnt, uint32_t lb,
                                                                  uint32_t log2u, uint32_t delta, uint32_t ctas_per_window,
                                                                  g2_xyzz_t* __restrict__ A_out, g2_xyzz_t* __restrict__ X_out) {
    extern __shared__ __align__(16) uint32_t g2cl_smem[];
    uint32_t* sm = g2cl_smem;
    uint32_t* flags = g2cl_smem + G2CL_SLOTS * 16 * 32;
    const uint32_t w = blockIdx.x / ctas_per_window;
    const uint32_t b = blockIdx.x % ctas_per_window;
    const g2_xyzz_t* Ain = A_in + (size_t)w * in_stride + in_off;
    const g2_xyzz_t* Xin = X_in ? X_in + (size_t)w * in_stride + in_off : nullptr;
    const size_t o = (size_t)w * ctas_per_window + b;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t Bsz = 1u << lb;
    const uint64_t first = ((uint64_t)b * 32 + lane) << lb;   // first item of this lane's chain
    for (int g = 0; g < 3; g++) g2cl_st(sm, g * 4 + warp, lane, fq2_zero());   // run = tot = xs = infinity
    __syncthreads();
    for (uint32_t k = Bsz; k-- > 0;) {
        const uint64_t i = first + k;
        const bool in = i < cnt;
        fq2 c = fq2_zero();
        if (in) c = fq2_load(reinterpret_cast<const char*>(Ain + i) + warp * 64);
        g2cl_st(sm, G2CL_B + warp, lane, c);
        __syncthreads();
        g2cl_add(sm, flags, G2CL_RUN, G2CL_B, warp, lane);      // run += item
        g2cl_add(sm, flags, G2CL_TOT, G2CL_RUN, warp, lane);    // tot += run
        if (Xin) {
            fq2 x = fq2_zero();
            if (in) x = fq2_load(reinterpret_cast<const char*>(Xin + i) + warp * 64);
            g2cl_st(sm, G2CL_B + warp, lane, x);
            __syncthreads();
            g2cl_add(sm, flags, G2CL_XS, G2CL_B, warp, lane);   // xs += side term
        }
    }
    // suffix scan of run over the lanes: run[l] = sum_{j >= l} run_j
    for (int d = 1; d < 32; d <<= 1) {
        g2cl_shift_copy(sm, G2CL_B, G2CL_RUN, d, true, warp, lane);
        g2cl_add(sm, flags, G2CL_RUN, G2CL_B, warp, lane);
    }
    //   V_l = tot_l + 2^lb * suffix[l+1]   (sum_l suffix[l+1] = sum_l l * run_l);   V_0 -= delta * R,  R = suffix[0]
    //   W_l = xs_l + 2^log2u * V_l   ->  X_out = sum_l W_l
    g2cl_shift_copy(sm, G2CL_B, G2CL_RUN, 1, true, warp, lane);   // B[l] = suffix[l + 1]
    for (uint32_t k = 0; k < lb; k++) g2cl_dbl(sm, G2CL_B, warp, lane, true);
    g2cl_add(sm, flags, G2CL_TOT, G2CL_B, warp, lane);
    if (delta) {
        fq2 c = fq2_zero();                                       // lane 0: -R (negate y); other lanes: infinity
        if (lane == 0) {
            c = g2cl_ld(sm, G2CL_RUN + warp, 0);
            if (warp == 1) c = fq2_neg(c);
        }
        __syncthreads();                                          // every warp has read its group-B operands of the add above
        g2cl_st(sm, G2CL_B + warp, lane, c);
        __syncthreads();
        g2cl_add(sm, flags, G2CL_TOT, G2CL_B, warp, lane);
    }
    for (uint32_t k = 0; k < log2u; k++) g2cl_dbl(sm, G2CL_TOT, warp, lane, true);
    g2cl_add(sm, flags, G2CL_XS, G2CL_TOT, warp, lane);
    for (int d = 16; d >= 1; d >>= 1) {
        g2cl_shift_copy(sm, G2CL_B, G2CL_XS, d, lane < d, warp, lane);
        g2cl_add(sm, flags, G2CL_XS, G2CL_B, warp, lane);
    }
    if (lane == 0) {
        fq2_store(reinterpret_cast<char*>(A_out + o) + warp * 64, g2cl_ld(sm, G2CL_RUN + warp, 0));
        fq2_store(reinterpret_cast<char*>(X_out + o) + warp * 64, g2cl_ld(sm, G2CL_XS + warp, 0));
    }
}

// Horner over the window sums, top window first: result = sum_w 2^(c w) G_w, written as arkworks G2Projective words.
// The c doublings between two windows are the critical path (c (W-1) of them, one after the other).  They run in Jacobian
// coordinates, cooperatively (below).  The one addition per window stays on thread 0 (XYZZ).
// Fq-granular cooperative engine: an Fq2 product is three independent Fq products (Karatsuba), a square two, so every
// phase of an XYZZ operation is spread over up to TWELVE warps, the lead lane of each computing ONE Fq product; the raw
// products go to shared memory and every consumer rebuilds the Fq2 values it needs with a few additions.  A doubling
// (dbl-2008-s-1) is three phases of one lone-warp Fq product (835 cycles) -- 4 / 11 / 9 products -- and an addition
// (add-2008-s) four -- 12 / 10 / 9 / 9 -- instead of 9 and 14 Fq2 products in a row on one thread, and the accumulator
// never leaves XYZZ (no Jacobian round trip per window).  Slots hold single Fq values (8 words).
enum { GF_X = 0, GF_Y = 2, GF_ZZ = 4, GF_ZZZ = 6,             // accumulator, plain Fq2 (c0, c1)
       GF_QX = 8, GF_QY = 10, GF_QZZ = 12, GF_QZZZ = 14,      // the addend, plain
       GF_A = 16, GF_B = 19, GF_C = 22, GF_D = 25,            // raw products: a product takes 3 slots (v0, v1, s), a square 2 (S, P)
       GF_E = 28, GF_F = 31, GF_G = 34, GF_H = 37, GF_I = 40, GF_J = 43, GF_K = 46, GF_L = 49,
       GF_SLOTS = 52 };
__device__ __forceinline__ fq gf_ld(const uint32_t* sm, int slot) { return fq_load(sm + slot * 8); }
__device__ __forceinline__ void gf_st(uint32_t* sm, int slot, const fq& v) { fq_store(sm + slot * 8, v); }
__device__ __forceinline__ fq2 gf_plain(const uint32_t* sm, int slot) { fq2 r; r.c0 = gf_ld(sm, slot); r.c1 = gf_ld(sm, slot + 1); return r; }
__device__ __forceinline__ void gf_put(uint32_t* sm, int slot, const fq2& v) { gf_st(sm, slot, v.c0); gf_st(sm, slot + 1, v.c1); }
// value of a product from its raw parts: (v0 - v1, s - v0 - v1); of a square: (S, 2P)
__device__ __forceinline__ fq2 gf_mulv(const uint32_t* sm, int slot) {
    fq v0 = gf_ld(sm, slot), v1 = gf_ld(sm, slot + 1);
    fq2 r; r.c0 = fq_sub(v0, v1); r.c1 = fq_sub(fq_sub(gf_ld(sm, slot + 2), v0), v1);
    return r;
}
__device__ __forceinline__ fq2 gf_sqrv(const uint32_t* sm, int slot) { fq2 r; r.c0 = gf_ld(sm, slot); r.c1 = fq_dbl(gf_ld(sm, slot + 1)); return r; }
// part 0..2 of the product a * b / part 0..1 of the square a^2 into the raw slots at `slot`
__device__ __forceinline__ void gf_mul_part(uint32_t* sm, int slot, int part, const fq2& a, const fq2& b) {
    if (part == 0) gf_st(sm, slot, fq_mul(a.c0, b.c0));
    else if (part == 1) gf_st(sm, slot + 1, fq_mul(a.c1, b.c1));
    else gf_st(sm, slot + 2, fq_mul(fq_add(a.c0, a.c1), fq_add(b.c0, b.c1)));
}
__device__ __forceinline__ void gf_sqr_part(uint32_t* sm, int slot, int part, const fq2& a) {
    if (part == 0) gf_st(sm, slot, fq_mul(fq_add(a.c0, a.c1), fq_sub(a.c0, a.c1)));
    else gf_st(sm, slot + 1, fq_mul(a.c0, a.c1));
}
__device__ __forceinline__ bool gf_zero2(const uint32_t* sm, int slot) {
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) o |= sm[slot * 8 + k];
    return o == 0;
}

#define G2_CMB_THREADS 384
// acc <- 2 acc (acc finite).  All threads call; only the lead lanes of the twelve warps work.
__device__ __forceinline__ void g2f_dbl(uint32_t* sm, int warp, bool lead) {
    // phase 1: V = (2Y)^2 -> A (w0, w1), XX = X^2 -> B (w2, w3)
    if (lead && warp < 2) gf_sqr_part(sm, GF_A, warp, fq2_dbl(gf_plain(sm, GF_Y)));
    if (lead && warp >= 2 && warp < 4) gf_sqr_part(sm, GF_B, warp - 2, gf_plain(sm, GF_X));
    __syncthreads();
    // phase 2: W = U V -> C (w0-2), S = X V -> D (w3-5), MM = M^2 -> E (w6, w7), ZZ' = V ZZ -> F (w8-10);  U = 2Y, M = 3 XX
    if (lead && warp < 11) {
        const fq2 V = gf_sqrv(sm, GF_A);
        if (warp < 3) gf_mul_part(sm, GF_C, warp, fq2_dbl(gf_plain(sm, GF_Y)), V);
        else if (warp < 6) gf_mul_part(sm, GF_D, warp - 3, gf_plain(sm, GF_X), V);
        else if (warp < 8) { fq2 xx = gf_sqrv(sm, GF_B); gf_sqr_part(sm, GF_E, warp - 6, fq2_add(fq2_dbl(xx), xx)); }
        else gf_mul_part(sm, GF_F, warp - 8, V, gf_plain(sm, GF_ZZ));
    }
    __syncthreads();
    // phase 3: T1 = M (S - X3) -> G (w0-2), T2 = W Y -> H (w3-5), ZZZ' = W ZZZ -> I (w6-8);  X3 = MM - 2S
    if (lead && warp < 9) {
        if (warp < 3) {
            fq2 xx = gf_sqrv(sm, GF_B), S = gf_mulv(sm, GF_D);
            fq2 X3 = fq2_sub(fq2_sub(gf_sqrv(sm, GF_E), S), S);
            gf_mul_part(sm, GF_G, warp, fq2_add(fq2_dbl(xx), xx), fq2_sub(S, X3));
        } else if (warp < 6) gf_mul_part(sm, GF_H, warp - 3, gf_mulv(sm, GF_C), gf_plain(sm, GF_Y));
        else gf_mul_part(sm, GF_I, warp - 6, gf_mulv(sm, GF_C), gf_plain(sm, GF_ZZZ));
    }
    __syncthreads();
    // commit (nothing below reads the old accumulator)
    if (lead && warp == 0) { fq2 S = gf_mulv(sm, GF_D); gf_put(sm, GF_X, fq2_sub(fq2_sub(gf_sqrv(sm, GF_E), S), S)); }
    if (lead && warp == 1) gf_put(sm, GF_Y, fq2_sub(gf_mulv(sm, GF_G), gf_mulv(sm, GF_H)));
    if (lead && warp == 2) gf_put(sm, GF_ZZ, gf_mulv(sm, GF_F));
    if (lead && warp == 3) gf_put(sm, GF_ZZZ, gf_mulv(sm, GF_I));
    __syncthreads();
}

// acc <- acc + Q, both finite (slots GF_X.. and GF_QX..).  Returns false -- with acc untouched -- when the two share their x
// (P + P or P + (-P)): the caller then takes the complete single-thread addition.  All threads call.
__device__ __forceinline__ bool g2f_add(uint32_t* sm, int* same_x, int warp, bool lead) {
    // phase 1: U1 = X1 ZZ2 -> A, U2 = X2 ZZ1 -> B, S1 = Y1 ZZZ2 -> C, S2 = Y2 ZZZ1 -> D  (twelve products)
    if (lead) {
        if (warp < 3) gf_mul_part(sm, GF_A, warp, gf_plain(sm, GF_X), gf_plain(sm, GF_QZZ));
        else if (warp < 6) gf_mul_part(sm, GF_B, warp - 3, gf_plain(sm, GF_QX), gf_plain(sm, GF_ZZ));
        else if (warp < 9) gf_mul_part(sm, GF_C, warp - 6, gf_plain(sm, GF_Y), gf_plain(sm, GF_QZZZ));
        else gf_mul_part(sm, GF_D, warp - 9, gf_plain(sm, GF_QY), gf_plain(sm, GF_ZZZ));
    }
    __syncthreads();
    // phase 2: P = U2 - U1, R = S2 - S1;  PP = P^2 -> E (w0, w1), RR = R^2 -> F (w2, w3), ZZ12 -> G (w4-6), ZZZ12 -> H (w7-9)
    if (lead && warp < 2) {
        fq2 P = fq2_sub(gf_mulv(sm, GF_B), gf_mulv(sm, GF_A));
        if (warp == 0) *same_x = fq2_is_zero(P) ? 1 : 0;
        gf_sqr_part(sm, GF_E, warp, P);
    }
    if (lead && warp >= 2 && warp < 4) gf_sqr_part(sm, GF_F, warp - 2, fq2_sub(gf_mulv(sm, GF_D), gf_mulv(sm, GF_C)));
    if (lead && warp >= 4 && warp < 7) gf_mul_part(sm, GF_G, warp - 4, gf_plain(sm, GF_ZZ), gf_plain(sm, GF_QZZ));
    if (lead && warp >= 7 && warp < 10) gf_mul_part(sm, GF_H, warp - 7, gf_plain(sm, GF_ZZZ), gf_plain(sm, GF_QZZZ));
    __syncthreads();
    if (*same_x) return false;   // uniform: read after the barrier; nothing has touched the accumulator
    // phase 3: PPP = P PP -> I (w0-2), Q = U1 PP -> J (w3-5), ZZ' = ZZ12 PP -> K (w6-8)
    if (lead && warp < 9) {
        const fq2 PP = gf_sqrv(sm, GF_E);
        if (warp < 3) gf_mul_part(sm, GF_I, warp, fq2_sub(gf_mulv(sm, GF_B), gf_mulv(sm, GF_A)), PP);
        else if (warp < 6) gf_mul_part(sm, GF_J, warp - 3, gf_mulv(sm, GF_A), PP);
        else gf_mul_part(sm, GF_K, warp - 6, gf_mulv(sm, GF_G), PP);
    }
    __syncthreads();
    // phase 4: X3 = RR - PPP - 2Q;  T1 = R (Q - X3) -> L (w0-2), T2 = S1 PPP -> B (w3-5), ZZZ' = ZZZ12 PPP -> E (w6-8).
    // B (U2) and E (PP) were last read in phase 3, before the barrier above, so they are free to be rewritten here.
    if (lead && warp < 9) {
        const fq2 PPP = gf_mulv(sm, GF_I);
        if (warp < 3) {
            fq2 Q = gf_mulv(sm, GF_J);
            fq2 X3 = fq2_sub(fq2_sub(fq2_sub(gf_sqrv(sm, GF_F), PPP), Q), Q);
            gf_mul_part(sm, GF_L, warp, fq2_sub(gf_mulv(sm, GF_D), gf_mulv(sm, GF_C)), fq2_sub(Q, X3));
        } else if (warp < 6) gf_mul_part(sm, GF_B, warp - 3, gf_mulv(sm, GF_C), PPP);          // T2 -> B (U2 is dead after phase 3)
        else gf_mul_part(sm, GF_E, warp - 6, gf_mulv(sm, GF_H), PPP);                           // ZZZ' -> E (PP is dead after phase 3)
    }
    __syncthreads();
    // commit
    if (lead && warp == 0) {
        fq2 Q = gf_mulv(sm, GF_J);
        gf_put(sm, GF_X, fq2_sub(fq2_sub(fq2_sub(gf_sqrv(sm, GF_F), gf_mulv(sm, GF_I)), Q), Q));
    }
    if (lead && warp == 1) gf_put(sm, GF_Y, fq2_sub(gf_mulv(sm, GF_L), gf_mulv(sm, GF_B)));
    if (lead && warp == 2) gf_put(sm, GF_ZZ, gf_mulv(sm, GF_K));
    if (lead && warp == 3) gf_put(sm, GF_ZZZ, gf_mulv(sm, GF_E));
    __syncthreads();
    return true;
}

__global__ void __launch_bounds__(G2_CMB_THREADS) k_g2_combine(const g2_xyzz_t* __restrict__ wsum, int W, int c,
                                                              g2_jac_t* __restrict__ out) {
    __shared__ __align__(16) uint32_t sm[GF_SLOTS * 8];
    __shared__ int same_x;
    const int warp = threadIdx.x >> 5;
    const bool lead = (threadIdx.x & 31) == 0;
    if (threadIdx.x < 64) sm[threadIdx.x] = 0;   // acc = infinity
    __syncthreads();
    for (int w = W - 1; w >= 0; w--) {
        if (!gf_zero2(sm, GF_ZZ))                 // uniform: every thread reads the same words after a barrier
            for (int k = 0; k < c; k++) g2f_dbl(sm, warp, lead);
        if (threadIdx.x < 64) sm[GF_QX * 8 + threadIdx.x] = reinterpret_cast<const uint32_t*>(wsum + w)[threadIdx.x];
        __syncthreads();
        const bool a_inf = gf_zero2(sm, GF_ZZ), q_inf = gf_zero2(sm, GF_QZZ);
        __syncthreads();                          // every thread has classified before the accumulator is rewritten below
        if (!q_inf) {
            if (a_inf) {
                if (threadIdx.x < 64) sm[threadIdx.x] = sm[GF_QX * 8 + threadIdx.x];
            } else if (!g2f_add(sm, &same_x, warp, lead)) {
                if (threadIdx.x == 0) {           // P + P or P + (-P): the complete addition, on one thread
                    g2_xyzz_t a, q;
                    a.x = gf_plain(sm, GF_X); a.y = gf_plain(sm, GF_Y); a.zz = gf_plain(sm, GF_ZZ); a.zzz = gf_plain(sm, GF_ZZZ);
                    q.x = gf_plain(sm, GF_QX); q.y = gf_plain(sm, GF_QY); q.zz = gf_plain(sm, GF_QZZ); q.zzz = gf_plain(sm, GF_QZZZ);
                    g2_add(a, q);
                    gf_put(sm, GF_X, a.x); gf_put(sm, GF_Y, a.y); gf_put(sm, GF_ZZ, a.zz); gf_put(sm, GF_ZZZ, a.zzz);
                }
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        g2_xyzz_t a;
        a.x = gf_plain(sm, GF_X); a.y = gf_plain(sm, GF_Y); a.zz = gf_plain(sm, GF_ZZ); a.zzz = gf_plain(sm, GF_ZZZ);
        g2_jac_t r = g2_to_jacobian(a);
        char* o = reinterpret_cast<char*>(out);
        fq2_store(o, r.x); fq2_store(o + 64, r.y); fq2_store(o + 128, r.z);
    }
}

// Multi-GPU combine: out = sum of `count` G2 Jacobian partials (192 B each), the G2 counterpart of k_sum_partials.
__global__ void __launch_bounds__(32) k_g2_sum_partials(const g2_jac_t* __restrict__ parts, int count, g2_jac_t* __restrict__ out) {
    if (threadIdx.x != 0) return;
    g2_xyzz_t acc = g2_inf();
    for (int k = 0; k < count; k++) {
        const char* p = reinterpret_cast<const char*>(parts + k);
        const fq2 x = fq2_load(p), y = fq2_load(p + 64), z = fq2_load(p + 128);
        if (fq2_is_zero(z)) continue;
        g2_xyzz_t v;
        v.x = x; v.y = y;
        v.zz = fq2_sqr(z);
        v.zzz = fq2_mul(v.zz, z);
        g2_add(acc, v);
    }
    g2_jac_t r = g2_to_jacobian(acc);
    char* o = reinterpret_cast<char*>(out);
    fq2_store(o, r.x); fq2_store(o + 64, r.y); fq2_store(o + 128, r.z);
}

// Precomputed window table for a registered G2 base set: table[w][i] = 2^(c w) * P_i in affine form (see k_build_table).
// One thread per point; (ZZ, ZZZ, running product) of every window parked in local memory, ONE Fq2 inversion per point.
#define G2_TBL_MAXW 33
__global__ void __launch_bounds__(64) k_g2_build_table(uint32_t n, size_t tstride, int c, int W, g2_affine_t* __restrict__ table) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    g2_affine_t p;
    p.x = fq2_load(reinterpret_cast<const char*>(table + i));
    p.y = fq2_load(reinterpret_cast<const char*>(table + i) + 64);
    if (g2_affine_is_inf(p)) {
        for (int w = 1; w < W; w++) {
            char* o = reinterpret_cast<char*>(table + (size_t)w * tstride + i);
            fq2_store(o, p.x); fq2_store(o + 64, p.y);
        }
        return;
    }
    fq2 zz[G2_TBL_MAXW], zzz[G2_TBL_MAXW], pref[G2_TBL_MAXW];
    g2_xyzz_t a;
    a.x = p.x; a.y = p.y; a.zz = fq2_one(); a.zzz = fq2_one();
    fq2 run = fq2_one();
    for (int w = 1; w < W; w++) {
        for (int k = 0; k < c; k++) g2_dbl(a);
        char* o = reinterpret_cast<char*>(table + (size_t)w * tstride + i);
        fq2_store(o, a.x); fq2_store(o + 64, a.y);
        zz[w] = a.zz;
        zzz[w] = a.zzz;
        run = fq2_mul(run, fq2_mul(a.zz, a.zzz));
        pref[w] = run;
    }
    fq2 inv = fq2_inv(run);
    for (int w = W - 1; w >= 1; w--) {
        fq2 dinv = w > 1 ? fq2_mul(inv, pref[w - 1]) : inv;       // 1 / (ZZ_w ZZZ_w)
        inv = fq2_mul(inv, fq2_mul(zz[w], zzz[w]));
        char* o = reinterpret_cast<char*>(table + (size_t)w * tstride + i);
        fq2 X = fq2_load(o), Y = fq2_load(o + 64);
        fq2_store(o, fq2_mul(X, fq2_mul(dinv, zzz[w])));          // X / ZZ
        fq2_store(o + 64, fq2_mul(Y, fq2_mul(dinv, zz[w])));      // Y / ZZZ
    }
}

// Test kit: element-wise Fq2 / G2 operations (op codes 30..) through the production device functions
__global__ void k_g2_tk_op(int op, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint8_t* __restrict__ out,
                           uint32_t count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    if (op == 30 || op == 31) {
        fq2 x = fq2_load(a + (size_t)i * 64);
        fq2 r = op == 30 ? fq2_mul(x, fq2_load(b + (size_t)i * 64)) : fq2_sqr(x);
        fq2_store(out + (size_t)i * 64, r);
    } else if (op == 32) {
        g2_xyzz_t acc = g2_load(a + (size_t)i * 256);
        g2_affine_t p;
        p.x = fq2_load(b + (size_t)i * 128); p.y = fq2_load(b + (size_t)i * 128 + 64);
        g2_madd(acc, p);
        g2_store(out + (size_t)i * 256, acc);
    } else if (op == 33) {
        g2_xyzz_t acc = g2_load(a + (size_t)i * 256), v = g2_load(b + (size_t)i * 256);
        g2_add(acc, v);
        g2_store(out + (size_t)i * 256, acc);
    } else if (op == 34) {
        g2_xyzz_t acc = g2_load(a + (size_t)i * 256);
        g2_dbl(acc);
        g2_store(out + (size_t)i * 256, acc);
    }
}
