// Device kernels of the B200 MSM pipeline (sparse-matrix Pippenger, cuZK-style content,
// B200-first structure).  Stage map against the reference:
//
//   K1 k_decompose        <- convert_point_coords_and_decompose_scalars.metal:16-122 (scalar half only:
//                            arkworks memory is already Montgomery, so there is no coordinate
//                            conversion and no Barrett) + the histogram half of transpose.metal:27-32
//   K2 k_scan_* / k_scatter <- transpose.metal:8-65 (serial per-window counting sort -> parallel)
//   K3 k_accumulate / k_fixup <- smvp.metal:14-107 (thread-per-bucket Jacobian adds ->
//                            load-balanced fixed-size chunks of the sorted entry list, XYZZ madd)
//   K4 k_bucket_reduce    <- pbpr.metal:33-148 (bpr_stage_1 + bpr_stage_2 fused)
//   K5 k_window_combine   <- final_reduction, metal_msm.rs:204-262 (CPU Horner -> device)
//
// Data layout in HBM (n points, window c, half = 2^(c-1), nb = half+1 counters per window, W windows):
//   bases    n x 64 B      x||y Montgomery LE (AoS: one gather = one 64-byte segment)
//   scalars  n x 32 B      Fr Montgomery LE
//   digits   W x n         int16 (c <= 16) or int32, window-major: signed digit d in [-half, half)
//   ends     W*nb u32      after K2: ends[g] = exclusive end of global bucket g = w*nb + |d|
//   entries  <= W*n u32    point index | sign<<31, grouped by global bucket (zero digits dropped)
//   buckets  W*nb x 128 B  XYZZ bucket sums (slot w*nb + m holds magnitude m; slot m = 0 unused)
//   head/tail nchunks x 128 B each: partial sums of buckets that straddle a chunk boundary
#pragma once
#include "g1.cuh"

#define MSM_FULL_MASK 0xffffffffu

// ------------------------------------------------------------------------------------------ K1
// GLV split (engine-internal): BN254 G1 has phi(x, y) = (beta*x, y) = lambda*(x, y).  Every scalar is written
// s = k1 + k2*lambda (mod r) with |k1|, |k2| < 2^127, turning the 254-bit MSM over n points into a 127-bit MSM
// over 2n pseudo-points (i -> P_i with k1_i, n+i -> phi(P_i) with k2_i): the same number of bucket additions,
// but HALF the windows for the latency-bound reduce stage (K4) and half the doublings of the Horner chain (K5).
// Exact integer procedure (restated by the test oracle (glv_decompose) and compared digit-for-digit in the tests):
//   c1 = (s*G1 + 2^255) >> 256,  c2 = (s*G2 + 2^255) >> 256            (G = floor(2^256 * b / r))
//   k1 = s - c1*A1 - c2*A2,      k2 = c1*|B1| - c2*B2                  (mod 2^256, sign = bit 255)
typedef unsigned __int128 u128_t;
__device__ __forceinline__ void glv_mul(const uint64_t* a, int na, const uint64_t* b, int nb, uint64_t* out, int nout) {
    for (int k = 0; k < nout; k++) out[k] = 0;
    for (int i = 0; i < na; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < nb && i + j < nout; j++) {
            u128_t t = (u128_t)a[i] * b[j] + out[i + j] + carry;
            out[i + j] = (uint64_t)t;
            carry = (uint64_t)(t >> 64);
        }
        for (int k = i + nb; k < nout && carry; k++) {
            u128_t t = (u128_t)out[k] + carry;
            out[k] = (uint64_t)t;
            carry = (uint64_t)(t >> 64);
        }
    }
}
__device__ __forceinline__ void glv_sub4(uint64_t* r, const uint64_t* a, const uint64_t* b) {  // r = a - b mod 2^256
    uint64_t borrow = 0;
    for (int k = 0; k < 4; k++) {
        u128_t t = (u128_t)a[k] - b[k] - borrow;
        r[k] = (uint64_t)t;
        borrow = (uint64_t)(t >> 64) & 1;
    }
}
// s: canonical scalar, 8 x u32.  Outputs: |k1|, |k2| as 4 x u32 each and their signs.
__device__ __forceinline__ void glv_decompose(const uint32_t (&t)[8], uint32_t (&k1)[4], bool& neg1, uint32_t (&k2)[4], bool& neg2) {
    const uint64_t G1[2] = {0xd91d232ec7e0b3d7ull, 0x0000000000000002ull};
    const uint64_t G2[3] = {0x7a7bd9d4391eb18dull, 0x4ccef014a773d2cfull, 0x0000000000000002ull};
    const uint64_t A1[1] = {0x89d3256894d213e3ull};
    const uint64_t A2[2] = {0x0be4e1541221250bull, 0x6f4d8248eeb859fdull};
    const uint64_t NB1[2] = {0x8211bbeb7d4f1128ull, 0x6f4d8248eeb859fcull};
    const uint64_t B2[1] = {0x89d3256894d213e3ull};
    uint64_t s[4];
    for (int k = 0; k < 4; k++) s[k] = (uint64_t)t[2 * k] | ((uint64_t)t[2 * k + 1] << 32);
    uint64_t p1[6], p2[7], c1[2], c2[3];
    glv_mul(s, 4, G1, 2, p1, 6);
    glv_mul(s, 4, G2, 3, p2, 7);
    {   // + 2^255, then >> 256
        u128_t x = (u128_t)p1[3] + 0x8000000000000000ull;
        uint64_t cy = (uint64_t)(x >> 64);
        x = (u128_t)p1[4] + cy; c1[0] = (uint64_t)x; cy = (uint64_t)(x >> 64);
        c1[1] = p1[5] + cy;
        x = (u128_t)p2[3] + 0x8000000000000000ull;
        cy = (uint64_t)(x >> 64);
        x = (u128_t)p2[4] + cy; c2[0] = (uint64_t)x; cy = (uint64_t)(x >> 64);
        x = (u128_t)p2[5] + cy; c2[1] = (uint64_t)x; cy = (uint64_t)(x >> 64);
        c2[2] = p2[6] + cy;
    }
    uint64_t m1[4], m2[4], r1[4], r2[4];
    glv_mul(c1, 2, A1, 1, m1, 4);
    glv_mul(c2, 3, A2, 2, m2, 4);
    glv_sub4(r1, s, m1);
    glv_sub4(r1, r1, m2);
    glv_mul(c1, 2, NB1, 2, m1, 4);
    glv_mul(c2, 3, B2, 1, m2, 4);
    glv_sub4(r2, m1, m2);
    neg1 = (r1[3] >> 63) != 0;
    neg2 = (r2[3] >> 63) != 0;
    if (neg1) { const uint64_t z[4] = {0, 0, 0, 0}; glv_sub4(r1, z, r1); }
    if (neg2) { const uint64_t z[4] = {0, 0, 0, 0}; glv_sub4(r2, z, r2); }
    k1[0] = (uint32_t)r1[0]; k1[1] = (uint32_t)(r1[0] >> 32); k1[2] = (uint32_t)r1[1]; k1[3] = (uint32_t)(r1[1] >> 32);
    k2[0] = (uint32_t)r2[0]; k2[1] = (uint32_t)(r2[0] >> 32); k2[2] = (uint32_t)r2[1]; k2[3] = (uint32_t)(r2[1] >> 32);
}

// Scalar -> signed digits, shared by K1 (k_decompose) and the partitioned sort's K1 (k_decompose_count,
// msm_psort_kernels.cuh): Montgomery reduction mod r, optional GLV split, signed-digit recoding (v >= half -> v - 2^c,
// carry; the reference's rule, convert kernel :108-116), digits stored window-major ([W][n_eff], n_eff = n or 2n).
// `on_digit(w, col, d, mag, live)` is called for every (window, pseudo-point), warp-uniformly (live = the digit is a
// non-zero digit of a valid point).
__device__ __forceinline__ bool load_scalar_canonical(const uint4* __restrict__ scalars, const uint8_t* __restrict__ inf_mask, uint32_t n,
                                                      uint32_t i, uint32_t (&t)[8]) {
    const bool valid = i < n;
    if (valid) {
        uint4 lo = __ldg(scalars + 2 * (size_t)i), hi = __ldg(scalars + 2 * (size_t)i + 1);
        t[0] = lo.x; t[1] = lo.y; t[2] = lo.z; t[3] = lo.w;
        t[4] = hi.x; t[5] = hi.y; t[6] = hi.z; t[7] = hi.w;
        fr_from_mont(t);
        if (inf_mask != nullptr && inf_mask[i]) {
#pragma unroll
            for (int k = 0; k < 8; k++) t[k] = 0;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = 0;
    }
    return valid;
}
template <typename DigitT, bool GLV, typename F>
__device__ __forceinline__ void decompose_scalar(const uint32_t (&t)[8], bool valid, uint32_t i, uint32_t n, int c, int W,
                                                 DigitT* __restrict__ digits, F&& on_digit) {
    const uint32_t half = 1u << (c - 1);
    const uint32_t cmask = (1u << c) - 1;
    const size_t n_eff = GLV ? 2 * (size_t)n : (size_t)n;
    // One pass of the digit extractor over NL limbs for pseudo-point `col`, digits negated when `neg`.
    // Streaming bit buffer: every limb index is a compile-time constant, so the scalar stays in registers
    // (a dynamic `t[bit >> 5]` would push it to local memory).  have <= c - 1 + 32 < 64.
    auto run = [&](const uint32_t* limbs, const int NL, size_t col, bool neg) {
        uint32_t carry = 0;
        uint64_t buf = 0;
        int have = 0, w = 0;
        auto emit = [&](uint32_t raw) {
            uint32_t v = raw + carry;
            int d;
            // negative scalars recode |k| with the mirrored rule (v > half wraps), so that after the sign flip
            // every digit is again in [-half, half - 1] and fits the int16 digit array at c = 16
            if (neg ? (v > half) : (v >= half)) { d = (int)v - (int)(cmask + 1); carry = 1; }
            else { d = (int)v; carry = 0; }
            if (neg) d = -d;
            if (valid) digits[(size_t)w * n_eff + col] = (DigitT)d;
            const uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
            on_digit(w, col, d, mag, valid && mag != 0);
            w++;
        };
#pragma unroll
        for (int k = 0; k < NL; k++) {
            buf |= (uint64_t)limbs[k] << have;
            have += 32;
            while (have >= c && w < W) {   // trip count depends only on (k, c): uniform across the warp
                emit((uint32_t)buf & cmask);
                buf >>= c;
                have -= c;
            }
        }
        while (w < W) {  // top window(s): the remaining high bits, then zeros
            emit((uint32_t)buf & cmask);
            buf >>= c;
        }
    };
    if (GLV) {
        uint32_t k1[4], k2[4];
        bool neg1, neg2;
        glv_decompose(t, k1, neg1, k2, neg2);
        run(k1, 4, (size_t)i, neg1);
        run(k2, 4, (size_t)n + i, neg2);
    } else {
        run(t, 8, (size_t)i, false);
    }
}

// One thread per scalar: 2 x 16-byte coalesced loads, digits (decompose_scalar), and a warp-aggregated histogram of |d|
// per window.
// RANK: the histogram atomic also hands every digit its rank inside its bucket (ranks[w][col]), so that the scatter pass
// needs no second round of atomics: position = bucket start + rank.
template <typename DigitT, bool GLV, bool RANK>
__global__ void __launch_bounds__(256) k_decompose(const uint4* __restrict__ scalars, const uint8_t* __restrict__ inf_mask,
                                                   uint32_t n, int c, int W, uint32_t wstride, DigitT* __restrict__ digits,
                                                   uint32_t* __restrict__ hist, uint32_t* __restrict__ ranks) {
    // wstride = 2^(c-1) + 1: one bucket set per window.  wstride = 0 (precomputed-table mode): all windows share one
    // bucket set, because window w of point i is served by the table point 2^(c*w) * P_i.
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t t[8];
    const bool valid = load_scalar_canonical(scalars, inf_mask, n, i, t);
    const unsigned lane = threadIdx.x & 31;
    const size_t n_eff = GLV ? 2 * (size_t)n : (size_t)n;
    decompose_scalar<DigitT, GLV>(t, valid, i, n, c, W, digits, [&](int w, size_t col, int d, uint32_t mag, bool live) {
        (void)d;
        const uint32_t key = live ? (uint32_t)w * wstride + mag : 0xffffffffu;
        // warp aggregation: one atomic per distinct key in the warp (skewed scalars -- many equal
        // small values, as witness vectors have -- would otherwise serialise on one L2 address)
        const unsigned peers = __match_any_sync(MSM_FULL_MASK, key);
        const unsigned leader = (unsigned)(__ffs(peers) - 1);
        uint32_t base = 0;
        if (key != 0xffffffffu && lane == leader) base = atomicAdd(hist + key, __popc(peers));
        if (RANK) {
            base = __shfl_sync(MSM_FULL_MASK, base, leader);
            if (key != 0xffffffffu) ranks[(size_t)w * n_eff + col] = base + __popc(peers & ((1u << lane) - 1));
        }
    });
}

// Precomputed-table mode (registered bases): table[w][i] = 2^(c*w) * P_i in affine form, w < W.
// One thread per point walks the doubling chain in XYZZ, parks the unnormalised (X, Y) in the table slot and
// (ZZ, ZZZ, running product of ZZ*ZZZ) in local memory, inverts the last running product once (Montgomery's trick
// across the windows of one point) and normalises backwards.  G1 has prime order, so no multiple of a finite point is
// infinity; the (0,0) infinity marker is copied to every window.
#define TBL_MAXW 33
__global__ void __launch_bounds__(128) k_build_table(uint32_t n, size_t tstride, int c, int W, affine_t* __restrict__ table) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // window 0 of the table IS the base array (the caller uploads the bases straight into it)
    affine_t p;
    p.x = fq_load(reinterpret_cast<const char*>(table + i));
    p.y = fq_load(reinterpret_cast<const char*>(table + i) + 32);
    if (fq_is_zero(p.x) && fq_is_zero(p.y)) {
        for (int w = 1; w < W; w++) {
            char* o = reinterpret_cast<char*>(table + (size_t)w * tstride + i);
            fq_store(o, p.x); fq_store(o + 32, p.y);
        }
        return;
    }
    fq zz[TBL_MAXW], zzz[TBL_MAXW], pref[TBL_MAXW];
    xyzz_t a = xyzz_from_affine(p);
    fq run = fq_one();
    for (int w = 1; w < W; w++) {
        for (int k = 0; k < c; k++) xyzz_dbl_inplace(a);
        char* o = reinterpret_cast<char*>(table + (size_t)w * tstride + i);
        fq_store(o, a.x); fq_store(o + 32, a.y);
        zz[w] = a.zz;
        zzz[w] = a.zzz;
        run = fq_mul(run, fq_mul(a.zz, a.zzz));
        pref[w] = run;
    }
    fq inv = fq_inv(run);
    for (int w = W - 1; w >= 1; w--) {
        fq dinv = w > 1 ? fq_mul(inv, pref[w - 1]) : inv;        // 1 / (ZZ_w * ZZZ_w)
        inv = fq_mul(inv, fq_mul(zz[w], zzz[w]));
        char* o = reinterpret_cast<char*>(table + (size_t)w * tstride + i);
        fq X = fq_load(o), Y = fq_load(o + 32);
        fq_store(o, fq_mul(X, fq_mul(dinv, zzz[w])));            // X / ZZ
        fq_store(o + 32, fq_mul(Y, fq_mul(dinv, zz[w])));        // Y / ZZZ
    }
}

// xb[i] = beta * x_i: the x coordinates of phi(P_i) (y is shared with P_i).  (0,0) stays (0,0).
__global__ void __launch_bounds__(256) k_endo_x(const affine_t* __restrict__ bases, uint32_t n, fq* __restrict__ xb) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fq beta = {{0xd782e155u, 0x71930c11u, 0xffbe3323u, 0xa6bb947cu, 0xd4741444u, 0xaa303344u, 0x26594943u, 0x2c3b3f0du}};
    fq x = fq_load_nc(bases + i);
    fq_store(xb + i, fq_mul(x, beta));
}

// ------------------------------------------------------------------------------------------ K2
// Two-level exclusive scan of the flat counter array (all windows back to back, `total` counters): CTA s scans segment
// [s*seg, min(total, (s+1)*seg)) with 1024 threads and emits the segment total; k_add_window_base adds the bases.
// inclusive = 0: exclusive offsets (the cursors k_scatter advances to the bucket ends); 1: bucket ends directly (ranked sort).
__global__ void __launch_bounds__(1024) k_scan_windows(uint32_t* __restrict__ hist, uint32_t seg, uint32_t total,
                                                       uint32_t* __restrict__ wtotal, int inclusive) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t running;
    const size_t seg_lo = (size_t)blockIdx.x * seg;
    uint32_t* h = hist + seg_lo;
    const uint32_t nb = seg_lo >= total ? 0u : (uint32_t)min((size_t)seg, (size_t)total - seg_lo);
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 4096) {
        uint32_t idx = base + tid * 4;
        uint32_t v0 = idx < nb ? h[idx] : 0, v1 = idx + 1 < nb ? h[idx + 1] : 0;
        uint32_t v2 = idx + 2 < nb ? h[idx + 2] : 0, v3 = idx + 3 < nb ? h[idx + 3] : 0;
        uint32_t s = v0 + v1 + v2 + v3;
        uint32_t inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(MSM_FULL_MASK, inc, o);
            if (lane >= (unsigned)o) inc += y;
        }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint32_t ws = warp_sums[lane];
            uint32_t winc = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(MSM_FULL_MASK, winc, o);
                if (lane >= (unsigned)o) winc += y;
            }
            warp_sums[lane] = winc - ws;  // exclusive
        }
        __syncthreads();
        uint32_t ex = running + warp_sums[wid] + inc - s;
        const uint32_t sh = inclusive ? v0 : 0;   // inclusive: everything moves up by one element
        if (idx < nb) h[idx] = ex + sh;
        if (idx + 1 < nb) h[idx + 1] = ex + v0 + (inclusive ? v1 : 0);
        if (idx + 2 < nb) h[idx + 2] = ex + v0 + v1 + (inclusive ? v2 : 0);
        if (idx + 3 < nb) h[idx + 3] = ex + v0 + v1 + v2 + (inclusive ? v3 : 0);
        __syncthreads();
        if (tid == 1023) running = ex + s;
        __syncthreads();
    }
    if (tid == 0) wtotal[blockIdx.x] = running;
}

// Add the segment base (sum of earlier segments' totals) so offsets index the global entry list.
__global__ void __launch_bounds__(256) k_add_window_base(uint32_t* __restrict__ hist, uint32_t seg, uint32_t total,
                                                         const uint32_t* __restrict__ wtotal) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= total) return;
    int w = (int)(j / seg);
    uint32_t base = 0;
    for (int v = 0; v < w; v++) base += wtotal[v];
    hist[j] += base;
}

// Flat, window-major pass over the digits: entries[cursor[key]++] = index | sign<<31.
// Afterwards cursor[g] is the exclusive END of bucket g (= start of g+1).
template <typename DigitT>
__global__ void __launch_bounds__(256) k_scatter(const DigitT* __restrict__ digits, uint32_t n, int W, uint32_t wstride,
                                                 uint32_t istride, uint32_t* __restrict__ cursor,
                                                 uint32_t* __restrict__ entries) {
    // (wstride, istride) = (buckets per window, 0) normally; (0, table stride) in precomputed-table mode, where the
    // entry is the index of 2^(c*w) * P_i in the [W][istride] table and every window feeds the same bucket set.
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)W * n;
    const unsigned lane = threadIdx.x & 31;
    int d = 0;
    uint32_t w = 0, i = 0;
    if (j < total) {
        w = (uint32_t)(j / n);
        i = (uint32_t)(j - (size_t)w * n);
        d = (int)digits[j];
    }
    uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
    uint32_t key = mag != 0 ? w * wstride + mag : 0xffffffffu;
    unsigned peers = __match_any_sync(MSM_FULL_MASK, key);
    unsigned leader = (unsigned)(__ffs(peers) - 1);
    uint32_t base = 0;
    if (key != 0xffffffffu && lane == leader) base = atomicAdd(cursor + key, __popc(peers));
    base = __shfl_sync(MSM_FULL_MASK, base, leader);
    if (key != 0xffffffffu) {
        uint32_t pos = base + __popc(peers & ((1u << lane) - 1));
        entries[pos] = (i + w * istride) | (d < 0 ? 0x80000000u : 0u);
    }
}

// Ranked scatter: no atomics.  ends[] already holds the bucket ends (inclusive scan); a digit of bucket `key` with rank r
// goes to position ends[key - 1] + r (key >= 1 always: magnitude 0 is never an entry).
template <typename DigitT>
__global__ void __launch_bounds__(256) k_scatter_ranked(const DigitT* __restrict__ digits, const uint32_t* __restrict__ ranks,
                                                        uint32_t n, int W, uint32_t wstride, uint32_t istride,
                                                        const uint32_t* __restrict__ ends, uint32_t* __restrict__ entries) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (size_t)W * n) return;
    const int d = (int)digits[j];
    if (d == 0) return;
    const uint32_t w = (uint32_t)(j / n);
    const uint32_t i = (uint32_t)(j - (size_t)w * n);
    const uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
    const uint32_t key = w * wstride + mag;
    entries[__ldg(ends + key - 1) + __ldg(ranks + j)] = (i + w * istride) | (d < 0 ? 0x80000000u : 0u);
}

// ------------------------------------------------------------------------------------------ K3
// Load-balanced bucket accumulation.  The sorted entry list is cut into fixed chunks of L
// entries, one thread per chunk, whatever the bucket sizes are: every lane of every warp does
// the same number of mixed additions.  A bucket that lies inside one chunk is written straight
// to `buckets`; the (at most two) buckets that straddle the chunk's edges leave partial sums in
// head[chunk] / tail[chunk], which k_fixup folds.  The next point is prefetched (4 x LDG.128)
// while the current addition runs.
#define ACC_THREADS 128
// Processes the global buckets [g_lo, g_hi) (a group of whole windows).  Chunk t always means the
// absolute positions [t*L, (t+1)*L) of the entry list, so head/tail slots and k_fixup's arithmetic do not
// depend on how the windows are grouped; a chunk cut by a group boundary is shared by two launches that
// touch disjoint slots (head[t] can only come from its first part, tail[t] only from its last).
// Pseudo-point gather: index < n is P_i (64 contiguous bytes); index >= n is phi(P_{i-n}) = (xb[i-n], y_{i-n}).
__device__ __forceinline__ affine_t load_pseudo_point(const affine_t* __restrict__ bases, const fq* __restrict__ xb, uint32_t n,
                                                      uint32_t idx) {
    const bool endo = idx >= n;
    const uint32_t base = endo ? idx - n : idx;
    const char* b = reinterpret_cast<const char*>(bases + base);
    affine_t p;
    p.x = fq_load_nc(endo ? reinterpret_cast<const char*>(xb + base) : b);
    p.y = fq_load_nc(b + 32);
    return p;
}

__global__ void __launch_bounds__(ACC_THREADS, 4) k_accumulate(const affine_t* __restrict__ bases, const fq* __restrict__ xb,
                                                            uint32_t n, const uint32_t* __restrict__ entries,
                                                            const uint32_t* __restrict__ ends, uint32_t g_lo, uint32_t g_hi,
                                                            uint32_t L, xyzz_t* __restrict__ buckets,
                                                            xyzz_t* __restrict__ head, xyzz_t* __restrict__ tail,
                                                            uint32_t* __restrict__ chunk_g, int into, int tail_group) {
    // into != 0 (slices 1.. of a sliced MSM): `buckets` already holds the sums of the earlier slices, and the accumulator of a
    // bucket whose FIRST entry lies in this chunk starts from that value instead of infinity -- the slices add up in place, with
    // no extra addition and no merge pass (a bucket's later chunks still start from infinity; the fix-up sums tail + heads).
    const uint32_t P0 = g_lo ? ends[g_lo - 1] : 0;
    const uint32_t P1 = ends[g_hi - 1];
    const uint64_t t64 = (uint64_t)(P0 / L) + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t clo64 = t64 * L;
    if (clo64 >= P1) return;
    const uint32_t t = (uint32_t)t64;
    const uint32_t clo = (uint32_t)clo64;                                  // chunk bounds
    const uint32_t chi = (uint32_t)min((uint64_t)0xffffffffu, clo64 + L);
    const uint32_t lo = max(clo, P0);                                      // part of the chunk inside this group
    const uint32_t hi = min(chi, P1);
    if (lo >= hi) return;
    // smallest g in [g_lo, g_hi) with ends[g] > lo
    uint32_t a = g_lo, b = g_hi - 1;
    while (a < b) {
        uint32_t mid = (a + b) >> 1;
        if (__ldg(ends + mid) > lo) b = mid; else a = mid + 1;
    }
    uint32_t g = a;
    uint32_t bstart = g ? __ldg(ends + g - 1) : 0;
    uint32_t bend = __ldg(ends + g);
    xyzz_t acc = xyzz_inf();
    if (into && bstart >= lo) acc = xyzz_load(buckets + g);
    uint32_t e_next = __ldg(entries + lo);
    affine_t p_next = load_pseudo_point(bases, xb, n, e_next & 0x7fffffffu);
    for (uint32_t pos = lo; pos < hi; pos++) {
        uint32_t e = e_next;
        affine_t p = p_next;
        if (pos + 1 < hi) {
            e_next = __ldg(entries + pos + 1);
            p_next = load_pseudo_point(bases, xb, n, e_next & 0x7fffffffu);
        }
        if (pos >= bend) {
            xyzz_t* dst = (bstart >= clo) ? buckets + g : head + t;  // bend <= pos < hi here
            xyzz_store(dst, acc);
            do { g++; bstart = bend; bend = __ldg(ends + g); } while (bend <= pos);
            acc = xyzz_inf();
            if (into) acc = xyzz_load(buckets + g);
        }
        // (0,0) is not on the curve: it is the device encoding of an infinity base (k_repack_bases) and adds nothing
        if (!(fq_is_zero(p.x) && fq_is_zero(p.y))) {
            p.y = fq_cneg(p.y, (e >> 31) != 0);
            xyzz_madd(acc, p);
        }
    }
    xyzz_t* dst;
    if (bstart < clo) dst = head + t;
    else if (bend > chi) dst = tail + t;
    else dst = buckets + g;
    xyzz_store(dst, acc);
    // for k_fixup_chunks: the bucket that runs on into the next chunk.  Only the launch whose window group covers the chunk's END
    // writes it (a chunk cut by a group boundary is touched by two launches, the higher group first); the very last chunk of the
    // entry list belongs to the group that holds the top window (tail_group)
    if (chunk_g != nullptr && (hi == chi || tail_group)) chunk_g[t] = bend > chi ? g : 0xffffffffu;
}

// One thread per global bucket: empty -> infinity; straddling -> tail[first chunk] + head[...].
// Buckets that span more than FIX_LONG chunks (skewed scalars; the narrow top window) are queued
// for k_fixup_long, which gives each one a whole CTA.
#define FIX_LONG 24
// ... and more than FIX_GIANT chunks (a bucket that holds a large share of ALL points: witness vectors full of ones) for
// k_fixup_giant, where every CTA of the grid sums a segment of FIX_SEG chunks.
#define FIX_GIANT 8192
#define FIX_SEG 4096
__global__ void __launch_bounds__(128) k_fixup(const uint32_t* __restrict__ ends, uint32_t g_lo, uint32_t g_hi, uint32_t L,
                                               xyzz_t* __restrict__ buckets, const xyzz_t* __restrict__ head,
                                               const xyzz_t* __restrict__ tail, uint32_t* __restrict__ long_count,
                                               uint32_t* __restrict__ long_list, int keep_empty,
                                               uint32_t* __restrict__ giant_count, uint32_t* __restrict__ giant_list) {
    uint32_t g = g_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= g_hi) return;
    uint32_t start = g ? ends[g - 1] : 0, end = ends[g];
    if (start == end) {
        if (!keep_empty) xyzz_store(buckets + g, xyzz_inf());   // keep_empty: the bucket holds the earlier slices' sum
        return;
    }
    uint32_t t0 = start / L, t1 = (end - 1) / L;
    if (t0 == t1) return;
    if (t1 - t0 > FIX_GIANT) {
        giant_list[atomicAdd(giant_count, 1u)] = g;
        return;
    }
    if (t1 - t0 > FIX_LONG) {
        long_list[atomicAdd(long_count, 1u)] = g;
        return;
    }
    xyzz_t acc = xyzz_load(tail + t0);
    for (uint32_t t = t0 + 1; t <= t1; t++) {
        xyzz_t h = xyzz_load(head + t);
        xyzz_add(acc, h);
    }
    xyzz_store(buckets + g, acc);
}

// The same fix-up with full warps, for big inputs (k_fixup runs 16-19 of 32 lanes at 2^24 / c = 20: one thread per bucket, and
// only the buckets that straddle a chunk edge add anything).  k_fixup_empty only writes the infinity of empty buckets;
// k_fixup_chunks gives one thread to every CHUNK whose last bucket runs on into the next chunk (k_accumulate leaves its index
// in chunk_g): a bucket that ends in the next chunk costs that thread exactly ONE addition (uniform lanes); a bucket that spans
// three or more chunks is queued -- up to FIX_LONG chunks for k_fixup_medium (one thread per queued bucket: neighbours in the
// queue come from the same window and have similar lengths), beyond that for k_fixup_long (one CTA per bucket).
__global__ void __launch_bounds__(256) k_fixup_empty(const uint32_t* __restrict__ ends, uint32_t g_lo, uint32_t g_hi,
                                                     xyzz_t* __restrict__ buckets) {
    const uint32_t g = g_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= g_hi) return;
    const uint32_t start = g ? ends[g - 1] : 0, end = ends[g];
    if (start == end) xyzz_store(buckets + g, xyzz_inf());
}
__global__ void __launch_bounds__(128) k_fixup_chunks(const uint32_t* __restrict__ ends, uint32_t g_lo, uint32_t g_hi, uint32_t L,
                                                      xyzz_t* __restrict__ buckets, const xyzz_t* __restrict__ head,
                                                      const xyzz_t* __restrict__ tail, const uint32_t* __restrict__ chunk_g,
                                                      uint32_t* __restrict__ long_count, uint32_t* __restrict__ long_list,
                                                      uint32_t* __restrict__ medium_count, uint32_t* __restrict__ medium_list,
                                                      uint32_t* __restrict__ giant_count, uint32_t* __restrict__ giant_list) {
    const uint32_t P0 = g_lo ? ends[g_lo - 1] : 0;
    const uint32_t P1 = ends[g_hi - 1];
    const uint64_t t64 = (uint64_t)(P0 / L) + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t64 * L >= P1) return;
    const uint32_t t = (uint32_t)t64;
    const uint32_t g = chunk_g[t];
    if (g == 0xffffffffu || g < g_lo || g >= g_hi) return;
    const uint32_t start = g ? ends[g - 1] : 0, end = ends[g];
    if (start / L != t) return;              // a middle chunk of a longer bucket: its first chunk's thread does the work
    const uint32_t t1 = (end - 1) / L;
    if (t1 - t > FIX_GIANT) {
        giant_list[atomicAdd(giant_count, 1u)] = g;
        return;
    }
    if (t1 - t > FIX_LONG) {
        long_list[atomicAdd(long_count, 1u)] = g;
        return;
    }
    if (t1 - t >= 2) {
        medium_list[atomicAdd(medium_count, 1u)] = g;
        return;
    }
    xyzz_t acc = xyzz_load(tail + t);
    xyzz_t h = xyzz_load(head + t + 1);
    xyzz_add(acc, h);
    xyzz_store(buckets + g, acc);
}
// Queued buckets (3 .. FIX_LONG + 1 chunks).  Many of them (throughput regime, e.g. 2^22 points at c = 16: every bucket is
// queued): one THREAD per bucket, neighbours have similar lengths.  Few of them (latency regime, small inputs: a bucket of a
// hundred entries over 8-entry chunks): one WARP per bucket -- lane l takes partial l, then a tree of ceil(log2(k)) additions
// instead of k sequential ones.
__global__ void __launch_bounds__(128) k_fixup_medium(const uint32_t* __restrict__ ends, uint32_t L, xyzz_t* __restrict__ buckets,
                                                      const xyzz_t* __restrict__ head, const xyzz_t* __restrict__ tail,
                                                      const uint32_t* __restrict__ medium_count, const uint32_t* __restrict__ medium_list) {
    __shared__ uint4 smT[4 * 16 * 8];
    const uint32_t count = *medium_count;
    const uint32_t total_warps = gridDim.x * (blockDim.x >> 5);
    if (count > total_warps / 2) {   // throughput regime (more queued buckets than half the warps of the grid)
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
            const uint32_t g = medium_list[i];
            const uint32_t start = g ? ends[g - 1] : 0, end = ends[g];
            const uint32_t t0 = start / L, t1 = (end - 1) / L;
            xyzz_t acc = xyzz_load(tail + t0);
            for (uint32_t t = t0 + 1; t <= t1; t++) {
                xyzz_t h = xyzz_load(head + t);
                xyzz_add(acc, h);
            }
            xyzz_store(buckets + g, acc);
        }
        return;
    }
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    xyzz_t* T = reinterpret_cast<xyzz_t*>(smT) + warp * 16;
    for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + warp; i < count; i += total_warps) {   // warp-uniform loop
        const uint32_t g = medium_list[i];
        const uint32_t start = g ? ends[g - 1] : 0, end = ends[g];
        const uint32_t t0 = start / L, t1 = (end - 1) / L;
        const uint32_t k = t1 - t0 + 1;                    // partials: tail[t0], head[t0 + 1 .. t1]; k <= FIX_LONG + 1 <= 32
        xyzz_t acc = xyzz_inf();
        if (lane == 0) acc = xyzz_load(tail + t0);
        else if (lane < k) acc = xyzz_load(head + t0 + lane);
        for (uint32_t half = 16; half >= 1; half >>= 1) {
            if (half >= k) continue;                        // nothing in the upper half (warp-uniform: k is)
            if (lane >= half && lane < 2 * half) xyzz_store(T + (lane - half), acc);
            __syncwarp();
            if (lane < half && lane + half < k) {
                xyzz_t v = xyzz_load(T + lane);
                xyzz_add(acc, v);
            }
            __syncwarp();
        }
        if (lane == 0) xyzz_store(buckets + g, acc);
    }
}

#define FIXL_THREADS 256
__global__ void __launch_bounds__(FIXL_THREADS) k_fixup_long(const uint32_t* __restrict__ ends, uint32_t L,
                                                             xyzz_t* __restrict__ buckets, const xyzz_t* __restrict__ head,
                                                             const xyzz_t* __restrict__ tail,
                                                             const uint32_t* __restrict__ long_count,
                                                             const uint32_t* __restrict__ long_list) {
    __shared__ uint4 sm[FIXL_THREADS * 8];
    xyzz_t* s = reinterpret_cast<xyzz_t*>(sm);
    const uint32_t count = *long_count;
    for (uint32_t item = blockIdx.x; item < count; item += gridDim.x) {
        const uint32_t g = long_list[item];
        const uint32_t start = g ? ends[g - 1] : 0, end = ends[g];
        const uint32_t t0 = start / L, t1 = (end - 1) / L;
        xyzz_t acc = xyzz_inf();
        for (uint32_t t = t0 + 1 + threadIdx.x; t <= t1; t += FIXL_THREADS) {
            xyzz_t h = xyzz_load(head + t);
            xyzz_add(acc, h);
        }
        if (threadIdx.x == 0) {
            xyzz_t h = xyzz_load(tail + t0);
            xyzz_add(acc, h);
        }
        xyzz_store(s + threadIdx.x, acc);
        __syncthreads();
        for (int stride = FIXL_THREADS / 2; stride > 0; stride >>= 1) {
            if (threadIdx.x < stride) {
                xyzz_t x = xyzz_load(s + threadIdx.x), y = xyzz_load(s + threadIdx.x + stride);
                xyzz_add(x, y);
                xyzz_store(s + threadIdx.x, x);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) xyzz_store(buckets + g, xyzz_load(s));
        __syncthreads();
    }
}

// Buckets that span more than FIX_GIANT chunks are cut into segments of FIX_SEG chunk partials; the (bucket, segment) pairs are
// dealt round-robin to the CTAs of the grid (strided per-thread sums, then a shared-memory tree, result to gpart); the CTA that
// completes a bucket's LAST segment (per-bucket ticket counter) adds the bucket's segment sums and its tail partial.  One launch,
// no host round trip.
__device__ __forceinline__ xyzz_t xyzz_load_cg(const xyzz_t* p) {   // L2 (coherent) loads: the partials were written by other SMs
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = __ldcg(q + k);
    xyzz_t r;
    uint32_t* o = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int k = 0; k < 8; k++) { o[4 * k] = v[k].x; o[4 * k + 1] = v[k].y; o[4 * k + 2] = v[k].z; o[4 * k + 3] = v[k].w; }
    return r;
}
__device__ __forceinline__ xyzz_t fixl_block_sum(xyzz_t* s, xyzz_t acc) {   // result valid in thread 0
    xyzz_store(s + threadIdx.x, acc);
    __syncthreads();
    for (int stride = FIXL_THREADS / 2; stride > 0; stride >>= 1) {
        if (threadIdx.x < stride) {
            xyzz_t x = xyzz_load(s + threadIdx.x), y = xyzz_load(s + threadIdx.x + stride);
            xyzz_add(x, y);
            xyzz_store(s + threadIdx.x, x);
        }
        __syncthreads();
    }
    xyzz_t r = xyzz_load(s);
    __syncthreads();
    return r;
}
__global__ void __launch_bounds__(FIXL_THREADS) k_fixup_giant(const uint32_t* __restrict__ ends, uint32_t L, xyzz_t* __restrict__ buckets,
                                                              const xyzz_t* __restrict__ head, const xyzz_t* __restrict__ tail,
                                                              const uint32_t* __restrict__ giant_count,
                                                              const uint32_t* __restrict__ giant_list, xyzz_t* __restrict__ gpart,
                                                              uint32_t* __restrict__ gdone) {
    __shared__ uint4 sm[FIXL_THREADS * 8];
    __shared__ uint32_t ticket;
    xyzz_t* s = reinterpret_cast<xyzz_t*>(sm);
    const uint32_t count = *giant_count;
    uint32_t off = 0;                                                     // index of the bucket's first segment in gpart
    for (uint32_t j = 0; j < count; j++) {
        const uint32_t g = giant_list[j];
        const uint32_t start = g ? ends[g - 1] : 0, end = ends[g];
        const uint32_t t0 = start / L, t1 = (end - 1) / L;
        const uint32_t nseg = (t1 - t0 + FIX_SEG - 1) / FIX_SEG;         // heads t0 + 1 .. t1
        for (uint32_t sg = 0; sg < nseg; sg++) {
            if ((off + sg) % gridDim.x != blockIdx.x) continue;          // (bucket, segment) pairs dealt round-robin to the CTAs
            const uint32_t lo = t0 + 1 + sg * FIX_SEG, hi = min(t1 + 1, lo + FIX_SEG);
            xyzz_t acc = xyzz_inf();
            for (uint32_t t = lo + threadIdx.x; t < hi; t += FIXL_THREADS) {
                xyzz_t h = xyzz_load(head + t);
                xyzz_add(acc, h);
            }
            xyzz_t r = fixl_block_sum(s, acc);
            if (threadIdx.x == 0) {
                xyzz_store(gpart + off + sg, r);
                __threadfence();
                ticket = atomicAdd(gdone + j, 1u);
            }
            __syncthreads();
            if (ticket == nseg - 1) {                                     // this CTA finished the bucket's last segment: final sum
                __threadfence();
                xyzz_t fin = xyzz_inf();
                for (uint32_t k = threadIdx.x; k < nseg; k += FIXL_THREADS) {
                    xyzz_t h = xyzz_load_cg(gpart + off + k);
                    xyzz_add(fin, h);
                }
                if (threadIdx.x == 0) {
                    xyzz_t h = xyzz_load(tail + t0);
                    xyzz_add(fin, h);
                }
                xyzz_t tot = fixl_block_sum(s, fin);
                if (threadIdx.x == 0) xyzz_store(buckets + g, tot);
            }
            __syncthreads();
        }
        off += nseg;
    }
}

// Sliced host-input MSM (b200msm.cu: enqueue_sliced): every slice of the point range has run its own sort,
// accumulation and fix-up into its own bucket array; bucket g of the MSM is the sum over slices.
#define MERGE_MAX 7
struct merge_srcs {
    const xyzz_t* p[MERGE_MAX];
};
__global__ void __launch_bounds__(128) k_merge_buckets(xyzz_t* __restrict__ dst, merge_srcs src, int count, uint32_t G) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    xyzz_t acc = xyzz_load(dst + g);
    for (int k = 0; k < count; k++) {
        xyzz_t b = xyzz_load(src.p[k] + g);
        xyzz_add(acc, b);
    }
    xyzz_store(dst + g, acc);
}

// ------------------------------------------------------------------------------------------ K4
// sum_m m * B[m] per window, without any per-thread scalar multiplication.
// Thread j of a window owns the Bsz (a power of two) magnitudes (j*Bsz, (j+1)*Bsz]; a descending
// running sum gives run_j = sum B[m] and tot_j = sum (m - j*Bsz) B[m].  Then
//     S = sum_j tot_j + Bsz * sum_j j*run_j,      j = 128*b + t  (CTA b, thread t)
// and sum_t t*run_t is the sum of the suffix sums of run_t over t >= 1: one Hillis-Steele suffix scan and
// two tree sums in shared memory.  CTA b emits R_b = sum_t run_t and T_b = sum_t tot_t + Bsz*sum_t t*run_t;
// k_window_finish repeats the same step over the CTAs of a window:  G_w = sum_b T_b + 128*Bsz * sum_b b*R_b.
#define RED_THREADS 128

// In: per-thread (run, tot).  Out: sA[0] = R = sum run, sB[0] = T = sum tot + 2^log2w * sum_t t*run_t
// (valid for every thread after the final barrier).  sA, sB: N slots each, sC: 1 slot.  All N threads must call.
template <int N>
__device__ __forceinline__ void block_weighted_sum(xyzz_t* sA, xyzz_t* sB, xyzz_t* sC, xyzz_t run, xyzz_t tot,
                                                   int log2w) {
    const int t = threadIdx.x;
    xyzz_store(sA + t, run);
    xyzz_store(sB + t, tot);
    __syncthreads();
    // inclusive suffix scan of sA
    for (int d = 1; d < N; d <<= 1) {
        xyzz_t v = xyzz_load(sA + t);
        if (t + d < N) {
            xyzz_t u = xyzz_load(sA + t + d);
            xyzz_add(v, u);
        }
        __syncthreads();
        xyzz_store(sA + t, v);
        __syncthreads();
    }
    // tree sum of sB -> sB[0], parked in sC
    for (int stride = N / 2; stride > 0; stride >>= 1) {
        if (t < stride) {
            xyzz_t x = xyzz_load(sB + t), y = xyzz_load(sB + t + stride);
            xyzz_add(x, y);
            xyzz_store(sB + t, x);
        }
        __syncthreads();
    }
    if (t == 0) xyzz_store(sC, xyzz_load(sB));
    __syncthreads();
    // Q = sum_{t>=1} suffix[t]: reuse sB
    {
        xyzz_t q = xyzz_inf();
        if (t >= 1) q = xyzz_load(sA + t);
        xyzz_store(sB + t, q);
    }
    __syncthreads();
    for (int stride = N / 2; stride > 0; stride >>= 1) {
        if (t < stride) {
            xyzz_t x = xyzz_load(sB + t), y = xyzz_load(sB + t + stride);
            xyzz_add(x, y);
            xyzz_store(sB + t, x);
        }
        __syncthreads();
    }
    if (t == 0) {
        xyzz_t q = xyzz_load(sB);
        for (int k = 0; k < log2w; k++) xyzz_dbl_inplace(q);
        xyzz_t ts = xyzz_load(sC);
        xyzz_add(ts, q);
        xyzz_store(sB, ts);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(RED_THREADS) k_bucket_reduce(const xyzz_t* __restrict__ buckets, uint32_t nb,
                                                               uint32_t log2Bsz, uint32_t blocks_per_window, uint32_t w_lo,
                                                               xyzz_t* __restrict__ wpartR, xyzz_t* __restrict__ wpartT) {
    __shared__ uint4 smA[RED_THREADS * 8];
    __shared__ uint4 smB[RED_THREADS * 8];
    __shared__ uint4 smC[8];
    const uint32_t half = nb - 1;
    const uint32_t Bsz = 1u << log2Bsz;
    const uint32_t w = w_lo + blockIdx.x / blocks_per_window;
    const uint32_t slot = w * blocks_per_window + blockIdx.x % blocks_per_window;
    const uint32_t j = (blockIdx.x % blocks_per_window) * RED_THREADS + threadIdx.x;
    const uint64_t lo64 = (uint64_t)j << log2Bsz;
    xyzz_t tot = xyzz_inf(), run = xyzz_inf();
    if (lo64 < half) {
        const uint32_t lo = (uint32_t)lo64;
        const uint32_t top = min(half, lo + Bsz);
        const xyzz_t* B = buckets + (size_t)w * nb;
        for (uint32_t m = top; m > lo; m--) {
            xyzz_t v = xyzz_load(B + m);
            xyzz_add(run, v);
            xyzz_add(tot, run);
        }
    }
    block_weighted_sum<RED_THREADS>(reinterpret_cast<xyzz_t*>(smA), reinterpret_cast<xyzz_t*>(smB),
                                    reinterpret_cast<xyzz_t*>(smC), run, tot, (int)log2Bsz);
    if (threadIdx.x == 0) {
        xyzz_store(wpartR + slot, xyzz_load(smA));
        xyzz_store(wpartT + slot, xyzz_load(smB));
    }
}

// One N-thread CTA per window (N = 32 or 128): thread b holds CTA b's (R_b, T_b) (blocks_per_window <= N).
template <int N>
__global__ void __launch_bounds__(N) k_window_finish(const xyzz_t* __restrict__ wpartR, const xyzz_t* __restrict__ wpartT,
                                                     uint32_t blocks_per_window, uint32_t log2weight, uint32_t w_lo,
                                                     xyzz_t* __restrict__ wsum) {
    __shared__ uint4 smA[N * 8];
    __shared__ uint4 smB[N * 8];
    __shared__ uint4 smC[8];
    const uint32_t w = w_lo + blockIdx.x;
    xyzz_t run = xyzz_inf(), tot = xyzz_inf();
    if (threadIdx.x < blocks_per_window) {
        run = xyzz_load(wpartR + (size_t)w * blocks_per_window + threadIdx.x);
        tot = xyzz_load(wpartT + (size_t)w * blocks_per_window + threadIdx.x);
    }
    block_weighted_sum<N>(reinterpret_cast<xyzz_t*>(smA), reinterpret_cast<xyzz_t*>(smB), reinterpret_cast<xyzz_t*>(smC), run,
                          tot, (int)log2weight);
    if (threadIdx.x == 0) xyzz_store(wsum + w, xyzz_load(smB));
}

// ------------------------------------------------------------------------------------------ K4 (cooperative)
// Lane-parallel cooperative EC engine.  A lone warp cannot issue IMAD.WIDE.X faster than ~1 per 6 cycles
// (835 cycles per field multiplication whatever the ILP: profiles/r01_pipe_bench.jsonl), so the latency of the
// bucket-reduce chains is cut by splitting every EC operation ACROSS warps instead: a CTA of 4 warps serves 32
// independent chains (one per lane); in each phase warp w computes ONE of the (up to four) independent
// multiplications of the operation for all 32 lanes, and values are exchanged through shared memory laid out
// [slot][limb][lane] (conflict-free).  An addition is 4 multiply phases (~2 us) instead of 14 serial
// multiplications (~6.5 us), and all 32 lanes of every warp are busy.
#define CL_THREADS 128
#define CL_SLOTS 30
__device__ __forceinline__ fq cl_ld(const uint32_t* sm, int slot, int lane) {
    fq r;
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[k] = sm[(slot * 8 + k) * 32 + lane];
    return r;
}
__device__ __forceinline__ void cl_st(uint32_t* sm, int slot, int lane, const fq& v) {
#pragma unroll
    for (int k = 0; k < 8; k++) sm[(slot * 8 + k) * 32 + lane] = v.v[k];
}
__device__ __forceinline__ bool cl_zero(const uint32_t* sm, int slot, int lane) {
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) o |= sm[(slot * 8 + k) * 32 + lane];
    return o == 0;
}
// slot groups (4 slots each = one XYZZ point per lane)
enum { CL_RUN = 0, CL_TOT = 4, CL_XS = 8, CL_B = 12, CL_T = 16 /* 10 temporaries: CL_T .. CL_T+9 */, CL_SAVE = 26 };

// acc(A) <- 2*acc(A) for lanes with `on` (others untouched).  All CL_THREADS threads call.
__device__ __noinline__ void cl_dbl(uint32_t* sm, int A, int warp, int lane, bool on) {
    on = on && !cl_zero(sm, A + 2, lane);
    if (on && warp == 0) { fq U = fq_dbl(cl_ld(sm, A + 1, lane)); cl_st(sm, CL_T + 0, lane, U); cl_st(sm, CL_T + 1, lane, fq_sqr(U)); }
    if (on && warp == 1) { fq a = fq_sqr(cl_ld(sm, A + 0, lane)); cl_st(sm, CL_T + 2, lane, fq_add(fq_dbl(a), a)); }
    __syncthreads();
    if (on && warp == 0) cl_st(sm, CL_T + 3, lane, fq_mul(cl_ld(sm, CL_T + 0, lane), cl_ld(sm, CL_T + 1, lane)));   // W = U V
    if (on && warp == 1) cl_st(sm, CL_T + 4, lane, fq_mul(cl_ld(sm, A + 0, lane), cl_ld(sm, CL_T + 1, lane)));      // S = X V
    if (on && warp == 2) cl_st(sm, CL_T + 5, lane, fq_sqr(cl_ld(sm, CL_T + 2, lane)));                               // MM
    __syncthreads();
    if (on && warp == 0) {
        fq S = cl_ld(sm, CL_T + 4, lane);
        fq X3 = fq_sub(fq_sub(cl_ld(sm, CL_T + 5, lane), S), S);
        cl_st(sm, CL_T + 6, lane, fq_mul(cl_ld(sm, CL_T + 2, lane), fq_sub(S, X3)));
        cl_st(sm, A + 0, lane, X3);
    }
    if (on && warp == 1) cl_st(sm, CL_T + 7, lane, fq_mul(cl_ld(sm, CL_T + 3, lane), cl_ld(sm, A + 1, lane)));      // W Y
    if (on && warp == 2) cl_st(sm, A + 2, lane, fq_mul(cl_ld(sm, CL_T + 1, lane), cl_ld(sm, A + 2, lane)));
    if (on && warp == 3) cl_st(sm, A + 3, lane, fq_mul(cl_ld(sm, CL_T + 3, lane), cl_ld(sm, A + 3, lane)));
    __syncthreads();
    if (on && warp == 0) cl_st(sm, A + 1, lane, fq_sub(cl_ld(sm, CL_T + 6, lane), cl_ld(sm, CL_T + 7, lane)));
    __syncthreads();
}

// acc(A) <- acc(A) + b(B) per lane, complete (infinity operands, P + P, P + (-P)).  All threads call.
__device__ __noinline__ void cl_add(uint32_t* sm, uint32_t* flags, int A, int B, int warp, int lane) {
    const bool b_inf = cl_zero(sm, B + 2, lane), a_inf = cl_zero(sm, A + 2, lane);
    const bool go = !b_inf && !a_inf;
    // phase 1: U1 = X1 ZZ2, U2 = X2 ZZ1, S1 = Y1 ZZZ2, S2 = Y2 ZZZ1
    if (go) {
        if (warp == 0) cl_st(sm, CL_T + 0, lane, fq_mul(cl_ld(sm, A + 0, lane), cl_ld(sm, B + 2, lane)));
        if (warp == 1) cl_st(sm, CL_T + 1, lane, fq_mul(cl_ld(sm, B + 0, lane), cl_ld(sm, A + 2, lane)));
        if (warp == 2) cl_st(sm, CL_T + 2, lane, fq_mul(cl_ld(sm, A + 1, lane), cl_ld(sm, B + 3, lane)));
        if (warp == 3) cl_st(sm, CL_T + 3, lane, fq_mul(cl_ld(sm, B + 1, lane), cl_ld(sm, A + 3, lane)));
    }
    __syncthreads();
    // phase 2: P, R, classification, PP, RR, ZZ1 ZZ2, ZZZ1 ZZZ2
    if (warp == 0) {
        uint32_t f = b_inf ? 1u : a_inf ? 2u : 0u;   // 1 keep acc, 2 copy b, 3 result infinity, 4 double
        if (go) {
            fq P = fq_sub(cl_ld(sm, CL_T + 1, lane), cl_ld(sm, CL_T + 0, lane));
            if (fq_is_zero(P)) {
                fq R = fq_sub(cl_ld(sm, CL_T + 3, lane), cl_ld(sm, CL_T + 2, lane));
                f = fq_is_zero(R) ? 4u : 3u;
            } else {
                cl_st(sm, CL_T + 4, lane, P);
                cl_st(sm, CL_T + 6, lane, fq_sqr(P));
            }
        }
        flags[lane] = f;
    }
    if (go && warp == 1) {
        fq R = fq_sub(cl_ld(sm, CL_T + 3, lane), cl_ld(sm, CL_T + 2, lane));
        cl_st(sm, CL_T + 5, lane, R);
        cl_st(sm, CL_T + 7, lane, fq_sqr(R));
    }
    if (go && warp == 2) cl_st(sm, CL_T + 8, lane, fq_mul(cl_ld(sm, A + 2, lane), cl_ld(sm, B + 2, lane)));
    if (go && warp == 3) cl_st(sm, CL_T + 9, lane, fq_mul(cl_ld(sm, A + 3, lane), cl_ld(sm, B + 3, lane)));
    __syncthreads();
    const uint32_t f = flags[lane];
    // phase 3: PPP = P PP (-> T1), Q = U1 PP (-> T3), ZZ' = ZZ12 PP
    if (f == 0) {
        if (warp == 0) cl_st(sm, CL_T + 1, lane, fq_mul(cl_ld(sm, CL_T + 4, lane), cl_ld(sm, CL_T + 6, lane)));
        if (warp == 1) cl_st(sm, CL_T + 3, lane, fq_mul(cl_ld(sm, CL_T + 0, lane), cl_ld(sm, CL_T + 6, lane)));
        if (warp == 2) cl_st(sm, CL_T + 8, lane, fq_mul(cl_ld(sm, CL_T + 8, lane), cl_ld(sm, CL_T + 6, lane)));
    }
    __syncthreads();
    // phase 4: X3 = RR - PPP - 2Q, t = R (Q - X3) (-> T4), u = S1 PPP (-> T6), ZZZ' = ZZZ12 PPP
    if (f == 0) {
        if (warp == 0) {
            fq Q = cl_ld(sm, CL_T + 3, lane);
            fq X3 = fq_sub(fq_sub(fq_sub(cl_ld(sm, CL_T + 7, lane), cl_ld(sm, CL_T + 1, lane)), Q), Q);
            cl_st(sm, CL_T + 4, lane, fq_mul(cl_ld(sm, CL_T + 5, lane), fq_sub(Q, X3)));
            cl_st(sm, A + 0, lane, X3);
        }
        if (warp == 1) cl_st(sm, CL_T + 6, lane, fq_mul(cl_ld(sm, CL_T + 2, lane), cl_ld(sm, CL_T + 1, lane)));
        if (warp == 2) cl_st(sm, CL_T + 9, lane, fq_mul(cl_ld(sm, CL_T + 9, lane), cl_ld(sm, CL_T + 1, lane)));
    }
    __syncthreads();
    // commit: each warp owns one coordinate
    if (f == 0) {
        if (warp == 1) cl_st(sm, A + 1, lane, fq_sub(cl_ld(sm, CL_T + 4, lane), cl_ld(sm, CL_T + 6, lane)));
        if (warp == 2) cl_st(sm, A + 2, lane, cl_ld(sm, CL_T + 8, lane));
        if (warp == 3) cl_st(sm, A + 3, lane, cl_ld(sm, CL_T + 9, lane));
    } else if (f == 2) {
        cl_st(sm, A + warp, lane, cl_ld(sm, B + warp, lane));
    } else if (f == 3) {
        cl_st(sm, A + warp, lane, fq_zero());
    }
    const int any_dbl = __syncthreads_or(f == 4);
    if (any_dbl) cl_dbl(sm, A, warp, lane, f == 4);
}

// copy group S -> group D with a lane shift: D[lane] = S[lane + shift] (infinity beyond lane 31 or when !take)
__device__ __noinline__ void cl_shift_copy(uint32_t* sm, int D, int S, int shift, bool take, int warp, int lane) {
    fq v = fq_zero();
    if (take && lane + shift < 32) v = cl_ld(sm, S + warp, lane + shift);
    cl_st(sm, D + warp, lane, v);
    __syncthreads();
}

// One level of the recursive weighted sum.  Per window: `cnt` input items A_in[i] and (optionally) side terms
// X_in[i]; the window's answer is   sum_i X_in[i] + u * sum_i w(i) A_in[i],   w(i) = i + 1 (delta = 0, level 0:
// item i is the bucket of magnitude i + 1) or w(i) = i (delta = 1, higher levels).  Lane-chain t owns the 2^lb items
// [t 2^lb, (t+1) 2^lb); CTA b owns chains 32 b .. 32 b + 31 and emits
//     A_out[b] = R_b = sum of its items,     X_out[b] = XS_b + u * (TOT_b + 2^lb Q_b - delta R_b)
// so that the answer becomes  sum_b X_out[b] + (u * 32 * 2^lb) * sum_b b * A_out[b]: the same problem, 32 * 2^lb
// times smaller, with delta = 1.  When one CTA is left its X_out is the window sum.
// Body of one CTA: items Ain[0 .. cnt) (and side terms Xin), chains of 2^lb items, CTA index b inside its problem; results
// to *A_res / *X_res.
__device__ __forceinline__ void reduce_level_cta(const xyzz_t* __restrict__ Ain, const xyzz_t* __restrict__ Xin, uint32_t cnt,
                                                 uint32_t lb, uint32_t log2u, uint32_t delta, uint32_t b,
                                                 xyzz_t* __restrict__ A_res, xyzz_t* __restrict__ X_res) {
    __shared__ uint32_t sm[CL_SLOTS * 8 * 32];
    __shared__ uint32_t flags[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t Bsz = 1u << lb;
    const uint64_t first = ((uint64_t)b * 32 + lane) << lb;   // first item of this lane's chain
    // run = tot = xs = infinity
    for (int g = 0; g < 3; g++) cl_st(sm, g * 4 + warp, lane, fq_zero());
    __syncthreads();
    for (uint32_t k = Bsz; k-- > 0;) {
        const uint64_t i = first + k;
        const bool in = i < cnt;
        fq c = fq_zero();
        if (in) c = fq_load(reinterpret_cast<const char*>(Ain + i) + warp * 32);
        cl_st(sm, CL_B + warp, lane, c);
        __syncthreads();
        cl_add(sm, flags, CL_RUN, CL_B, warp, lane);      // run += item
        cl_add(sm, flags, CL_TOT, CL_RUN, warp, lane);    // tot += run
        if (Xin) {
            fq x = fq_zero();
            if (in) x = fq_load(reinterpret_cast<const char*>(Xin + i) + warp * 32);
            cl_st(sm, CL_B + warp, lane, x);
            __syncthreads();
            cl_add(sm, flags, CL_XS, CL_B, warp, lane);   // xs += side term
        }
    }
    // lane-level combine.  suffix scan of run: run[l] = sum_{j >= l} run_j
    for (int d = 1; d < 32; d <<= 1) {
        cl_shift_copy(sm, CL_B, CL_RUN, d, true, warp, lane);
        cl_add(sm, flags, CL_RUN, CL_B, warp, lane);
    }
    // Per lane, BEFORE any tree sum (every lane runs the same cooperative operation anyway, so weighting all 32
    // partials costs what weighting one does, and one tree sum replaces three):
    //   V_l = tot_l + 2^lb * suffix[l+1]            (sum_l suffix[l+1] = sum_l l * run_l)
    //   V_0 -= delta * R,  R = suffix[0]
    //   W_l = xs_l + u * V_l,  u = 2^log2u          ->  X_out = sum_l W_l
    cl_shift_copy(sm, CL_SAVE, CL_RUN, 1, true, warp, lane);   // SAVE[l] = suffix[l + 1]
    for (uint32_t k = 0; k < lb; k++) cl_dbl(sm, CL_SAVE, warp, lane, true);
    cl_add(sm, flags, CL_TOT, CL_SAVE, warp, lane);
    if (delta) {
        fq c = fq_zero();                                      // lane 0: -R (negate y); other lanes: infinity
        if (lane == 0) {
            c = cl_ld(sm, CL_RUN + warp, 0);
            if (warp == 1) c = fq_neg(c);
        }
        cl_st(sm, CL_B + warp, lane, c);
        __syncthreads();
        cl_add(sm, flags, CL_TOT, CL_B, warp, lane);
    }
    for (uint32_t k = 0; k < log2u; k++) cl_dbl(sm, CL_TOT, warp, lane, true);
    cl_add(sm, flags, CL_XS, CL_TOT, warp, lane);
    for (int d = 16; d >= 1; d >>= 1) {
        cl_shift_copy(sm, CL_B, CL_XS, d, lane < d, warp, lane);
        cl_add(sm, flags, CL_XS, CL_B, warp, lane);
    }
    if (lane == 0) {
        fq_store(reinterpret_cast<char*>(A_res) + warp * 32, cl_ld(sm, CL_RUN + warp, 0));
        fq_store(reinterpret_cast<char*>(X_res) + warp * 32, cl_ld(sm, CL_XS + warp, 0));
    }
}
__global__ void __launch_bounds__(CL_THREADS) k_reduce_level(const xyzz_t* __restrict__ A_in, const xyzz_t* __restrict__ X_in,
                                                              uint32_t in_stride, uint32_t in_off, uint32_t cnt, uint32_t lb,
                                                              uint32_t log2u, uint32_t delta, uint32_t ctas_per_window,
                                                              uint32_t w_lo, xyzz_t* __restrict__ A_out,
                                                              xyzz_t* __restrict__ X_out) {
    const uint32_t w = w_lo + blockIdx.x / ctas_per_window;
    const uint32_t b = blockIdx.x % ctas_per_window;
    const xyzz_t* Ain = A_in + (size_t)w * in_stride + in_off;
    const xyzz_t* Xin = X_in ? X_in + (size_t)w * in_stride + in_off : nullptr;
    const size_t o = (size_t)w * ctas_per_window + b;
    reduce_level_cta(Ain, Xin, cnt, lb, log2u, delta, b, A_out + o, X_out + o);
}

// ------------------------------------------------------------------------------------------ K4 (row / column sums)
// The latency-bound middle regime (2^10 .. 2^16 buckets per window) without running sums over all buckets: write the
// magnitude as m = hi * C + lo + 1 (C = 2^cb columns, R = 2^ra rows, ra + cb = c - 1); then
//     sum_m m B_m  =  C * sum_hi hi * ROW_hi  +  sum_lo (lo + 1) * COL_lo,      ROW_hi = sum_lo B[hi][lo],  COL_lo = sum_hi B[hi][lo]
// i.e. two PLAIN sums per bucket (as many additions as the running-sum method, but every one of them independent: one warp
// per row / column, a few serial additions per lane and a 5-level tree), followed by the weighted sums of only R + C items
// per window, which the cooperative engine finishes in ONE level with two CTAs per window (k_reduce_rowcol).
#define RC_WARPS 4
__global__ void __launch_bounds__(RC_WARPS * 32) k_rowcol_sums(const xyzz_t* __restrict__ buckets, uint32_t nb, uint32_t ra, uint32_t cb,
                                                               uint32_t w_lo, uint32_t problems, xyzz_t* __restrict__ rc) {
    __shared__ uint4 smT[RC_WARPS * 16 * 8];
    const uint32_t R = 1u << ra, C = 1u << cb;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t pid = blockIdx.x * RC_WARPS + warp;
    if (pid >= problems) return;                       // whole warps leave; the tree below only uses __syncwarp
    const uint32_t w = w_lo + pid / (R + C);
    const uint32_t prob = pid % (R + C);
    const xyzz_t* B = buckets + (size_t)w * nb + 1;    // bucket of magnitude m at B[m - 1]
    const bool row = prob < R;
    const uint32_t count = row ? C : R;
    const uint32_t base = row ? prob * C : prob - R;
    const uint32_t stride = row ? 1u : C;
    xyzz_t acc = xyzz_inf();
    for (uint32_t j = lane; j < count; j += 32) {
        xyzz_t v = xyzz_load(B + base + (size_t)j * stride);
        xyzz_add(acc, v);
    }
    xyzz_t* T = reinterpret_cast<xyzz_t*>(smT) + warp * 16;
    for (uint32_t half = 16; half >= 1; half >>= 1) {
        if (lane >= half && lane < 2 * half) xyzz_store(T + (lane - half), acc);
        __syncwarp();
        if (lane < half) {
            xyzz_t v = xyzz_load(T + lane);
            xyzz_add(acc, v);
        }
        __syncwarp();
    }
    if (lane == 0) xyzz_store(rc + (size_t)w * (R + C) + prob, acc);
}
// a[i] += b[i] (test-kit probe of the per-window sums under the row / column reduce)
__global__ void __launch_bounds__(32) k_add_into(xyzz_t* __restrict__ a, const xyzz_t* __restrict__ b, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xyzz_t x = xyzz_load(a + i), y = xyzz_load(b + i);
    xyzz_add(x, y);
    xyzz_store(a + i, x);
}
// Two CTAs per window: CTA 0 the rows (weights hi, times C), CTA 1 the columns (weights lo + 1).  Results to wsum_rows[w] /
// wsum_cols[w]; k_window_combine adds both.
__global__ void __launch_bounds__(CL_THREADS) k_reduce_rowcol(const xyzz_t* __restrict__ rc, uint32_t ra, uint32_t cb, uint32_t w_lo,
                                                               xyzz_t* __restrict__ scratch, xyzz_t* __restrict__ wsum_rows,
                                                               xyzz_t* __restrict__ wsum_cols) {
    const uint32_t R = 1u << ra, C = 1u << cb;
    const uint32_t w = w_lo + blockIdx.x / 2;
    const bool rows = (blockIdx.x & 1) == 0;
    const uint32_t cnt = rows ? R : C;
    uint32_t lb = 0;
    while ((32u << lb) < cnt) lb++;
    const xyzz_t* A = rc + (size_t)w * (R + C) + (rows ? 0 : R);
    reduce_level_cta(A, nullptr, cnt, lb, rows ? cb : 0u, rows ? 1u : 0u, 0u, scratch + (size_t)w * 2 + (rows ? 0 : 1),
                     (rows ? wsum_rows : wsum_cols) + w);
}

// ------------------------------------------------------------------------------------------ K5
// Horner from the top window down:  acc = 2^c * acc + G_w.  The chain of c*(W-1) doublings is
// inherently serial, and ONE warp is issue-bound at >= 544 cycles per field multiplication (each
// IMAD.WIDE.X holds the fmaheavy pipe of its SM sub-partition for 4 cycles whatever the lane count), so
// the independent multiplications inside each doubling / addition are spread over the FOUR warps of the
// CTA (= four sub-partitions), exchanging 32-byte values through shared memory:
//   doubling (dbl-2008-s-1, 9 mul) = 3 multiply phases,  addition (add-2008-s, 14 mul) = 4 multiply phases.
// Lane 0 of each warp computes; every thread takes the barriers.  Result: Jacobian (arkworks memory).
#define CMB_THREADS 128
enum { S_X = 0, S_Y, S_ZZ, S_ZZZ, S_BX, S_BY, S_BZZ, S_BZZZ, S_T0, S_T1, S_T2, S_T3, S_T4, S_T5, S_T6, S_T7, S_SLOTS };

__device__ __forceinline__ fq sm_ld(const uint32_t* sm, int slot) { return fq_load(sm + slot * 8); }
__device__ __forceinline__ void sm_st(uint32_t* sm, int slot, const fq& v) { fq_store(sm + slot * 8, v); }
__device__ __forceinline__ bool sm_is_zero(const uint32_t* sm, int slot) {
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) o |= sm[slot * 8 + k];
    return o == 0;
}

// acc (slots S_X..S_ZZZ) <- 2*acc.  All CMB_THREADS threads call; acc must not be infinity.
__device__ __forceinline__ void coop_dbl(uint32_t* sm, int warp, bool lead) {
    // phase 1:  w0: U = 2Y, V = U^2      w1: M = 3 X^2
    if (lead && warp == 0) { fq U = fq_dbl(sm_ld(sm, S_Y)); sm_st(sm, S_T0, U); sm_st(sm, S_T1, fq_sqr(U)); }
    if (lead && warp == 1) { fq A = fq_sqr(sm_ld(sm, S_X)); sm_st(sm, S_T2, fq_add(fq_dbl(A), A)); }
    __syncthreads();
    // phase 2:  w0: W = U V     w1: S = X V     w2: MM = M^2
    if (lead && warp == 0) sm_st(sm, S_T3, fq_mul(sm_ld(sm, S_T0), sm_ld(sm, S_T1)));
    if (lead && warp == 1) sm_st(sm, S_T4, fq_mul(sm_ld(sm, S_X), sm_ld(sm, S_T1)));
    if (lead && warp == 2) sm_st(sm, S_T5, fq_sqr(sm_ld(sm, S_T2)));
    __syncthreads();
    // phase 3:  w0: X3 = MM - 2S, t = M (S - X3)     w1: u = W Y     w2: ZZ *= V     w3: ZZZ *= W
    if (lead && warp == 0) {
        fq S = sm_ld(sm, S_T4);
        fq X3 = fq_sub(fq_sub(sm_ld(sm, S_T5), S), S);
        sm_st(sm, S_T6, fq_mul(sm_ld(sm, S_T2), fq_sub(S, X3)));
        sm_st(sm, S_X, X3);
    }
    if (lead && warp == 1) sm_st(sm, S_T7, fq_mul(sm_ld(sm, S_T3), sm_ld(sm, S_Y)));
    if (lead && warp == 2) sm_st(sm, S_ZZ, fq_mul(sm_ld(sm, S_T1), sm_ld(sm, S_ZZ)));
    if (lead && warp == 3) sm_st(sm, S_ZZZ, fq_mul(sm_ld(sm, S_T3), sm_ld(sm, S_ZZZ)));
    __syncthreads();
    if (lead && warp == 0) sm_st(sm, S_Y, fq_sub(sm_ld(sm, S_T6), sm_ld(sm, S_T7)));
    __syncthreads();
}

// acc (S_X..S_ZZZ) <- acc + b (S_BX..S_BZZZ), complete.  All threads call.
__device__ __forceinline__ void coop_add(uint32_t* sm, int warp, bool lead) {
    if (sm_is_zero(sm, S_BZZ)) {  // uniform: every thread reads the same shared words
        __syncthreads();          // ... and all reads finish before the caller refills the operand slots
        return;
    }
    if (sm_is_zero(sm, S_ZZ)) {
        __syncthreads();
        if (threadIdx.x < 32) sm[S_X * 8 + threadIdx.x] = sm[S_BX * 8 + threadIdx.x];
        __syncthreads();
        return;
    }
    // phase 1:  U1 = X1 ZZ2, U2 = X2 ZZ1, S1 = Y1 ZZZ2, S2 = Y2 ZZZ1
    if (lead && warp == 0) sm_st(sm, S_T0, fq_mul(sm_ld(sm, S_X), sm_ld(sm, S_BZZ)));
    if (lead && warp == 1) sm_st(sm, S_T1, fq_mul(sm_ld(sm, S_BX), sm_ld(sm, S_ZZ)));
    if (lead && warp == 2) sm_st(sm, S_T2, fq_mul(sm_ld(sm, S_Y), sm_ld(sm, S_BZZZ)));
    if (lead && warp == 3) sm_st(sm, S_T3, fq_mul(sm_ld(sm, S_BY), sm_ld(sm, S_ZZZ)));
    __syncthreads();
    // P = U2 - U1, R = S2 - S1 (by thread 0), then a uniform decision on the special cases
    if (threadIdx.x == 0) {
        sm_st(sm, S_T4, fq_sub(sm_ld(sm, S_T1), sm_ld(sm, S_T0)));
        sm_st(sm, S_T5, fq_sub(sm_ld(sm, S_T3), sm_ld(sm, S_T2)));
    }
    __syncthreads();
    if (sm_is_zero(sm, S_T4)) {
        const bool same = sm_is_zero(sm, S_T5);
        __syncthreads();
        if (same) {
            coop_dbl(sm, warp, lead);
        } else {
            if (threadIdx.x < 32) sm[S_X * 8 + threadIdx.x] = 0;
            __syncthreads();
        }
        return;
    }
    // phase 2:  PP = P^2, RR = R^2, ZZ12 = ZZ1 ZZ2, ZZZ12 = ZZZ1 ZZZ2
    if (lead && warp == 0) sm_st(sm, S_T6, fq_sqr(sm_ld(sm, S_T4)));
    if (lead && warp == 1) sm_st(sm, S_T7, fq_sqr(sm_ld(sm, S_T5)));
    if (lead && warp == 2) sm_st(sm, S_ZZ, fq_mul(sm_ld(sm, S_ZZ), sm_ld(sm, S_BZZ)));
    if (lead && warp == 3) sm_st(sm, S_ZZZ, fq_mul(sm_ld(sm, S_ZZZ), sm_ld(sm, S_BZZZ)));
    __syncthreads();
    // phase 3:  PPP = P PP (-> T1), Q = U1 PP (-> T3), ZZ = ZZ12 PP
    if (lead && warp == 0) sm_st(sm, S_T1, fq_mul(sm_ld(sm, S_T4), sm_ld(sm, S_T6)));
    if (lead && warp == 1) sm_st(sm, S_T3, fq_mul(sm_ld(sm, S_T0), sm_ld(sm, S_T6)));
    if (lead && warp == 2) sm_st(sm, S_ZZ, fq_mul(sm_ld(sm, S_ZZ), sm_ld(sm, S_T6)));
    __syncthreads();
    // phase 4:  X3 = RR - PPP - 2Q, t = R (Q - X3) (-> T0), u = S1 PPP (-> T4), ZZZ = ZZZ12 PPP
    if (lead && warp == 0) {
        fq Q = sm_ld(sm, S_T3);
        fq X3 = fq_sub(fq_sub(fq_sub(sm_ld(sm, S_T7), sm_ld(sm, S_T1)), Q), Q);
        sm_st(sm, S_T0, fq_mul(sm_ld(sm, S_T5), fq_sub(Q, X3)));
        sm_st(sm, S_X, X3);
    }
    if (lead && warp == 1) sm_st(sm, S_T4, fq_mul(sm_ld(sm, S_T2), sm_ld(sm, S_T1)));
    if (lead && warp == 2) sm_st(sm, S_ZZZ, fq_mul(sm_ld(sm, S_ZZZ), sm_ld(sm, S_T1)));
    __syncthreads();
    if (threadIdx.x == 0) sm_st(sm, S_Y, fq_sub(sm_ld(sm, S_T0), sm_ld(sm, S_T4)));
    __syncthreads();
}

// Windows [w_lo, w_hi) of the chain; `state` carries the accumulator between the launches of successive
// window groups (first: start from infinity; last: emit the Jacobian result).
__global__ void __launch_bounds__(CMB_THREADS) k_window_combine(const xyzz_t* __restrict__ wsum, const xyzz_t* __restrict__ wsum2,
                                                                int w_lo, int w_hi, int c,
                                                                uint32_t* __restrict__ state, int first, int last,
                                                                jac_t* __restrict__ out) {
    __shared__ __align__(16) uint32_t sm[S_SLOTS * 8];
    const int warp = threadIdx.x >> 5;
    const bool lead = (threadIdx.x & 31) == 0;
    if (threadIdx.x < 32) sm[S_X * 8 + threadIdx.x] = first ? 0u : state[threadIdx.x];
    __syncthreads();
    for (int w = w_hi - 1; w >= w_lo; w--) {
        if (!sm_is_zero(sm, S_ZZ)) {
            for (int k = 0; k < c; k++) coop_dbl(sm, warp, lead);
        }
        if (threadIdx.x < 32) sm[S_BX * 8 + threadIdx.x] = reinterpret_cast<const uint32_t*>(wsum + w)[threadIdx.x];
        __syncthreads();
        coop_add(sm, warp, lead);
        if (wsum2 != nullptr) {   // row / column reduce: the window sum arrives as two terms
            if (threadIdx.x < 32) sm[S_BX * 8 + threadIdx.x] = reinterpret_cast<const uint32_t*>(wsum2 + w)[threadIdx.x];
            __syncthreads();
            coop_add(sm, warp, lead);
        }
    }
    if (threadIdx.x < 32) state[threadIdx.x] = sm[S_X * 8 + threadIdx.x];
    if (last && threadIdx.x == 0) {
        xyzz_t a;
        a.x = sm_ld(sm, S_X); a.y = sm_ld(sm, S_Y); a.zz = sm_ld(sm, S_ZZ); a.zzz = sm_ld(sm, S_ZZZ);
        jac_t r = xyzz_to_jacobian(a);
        char* o = reinterpret_cast<char*>(out);
        fq_store(o, r.x); fq_store(o + 32, r.y); fq_store(o + 64, r.z);
    }
}

// Multi-GPU combine: out = sum of `count` Jacobian partials (96 B each).
__global__ void __launch_bounds__(32) k_sum_partials(const jac_t* __restrict__ parts, int count, jac_t* __restrict__ out) {
    if (threadIdx.x != 0) return;
    xyzz_t acc = xyzz_inf();
    for (int k = 0; k < count; k++) {
        const char* p = reinterpret_cast<const char*>(parts + k);
        jac_t j;
        j.x = fq_load(p); j.y = fq_load(p + 32); j.z = fq_load(p + 64);
        xyzz_t v = xyzz_from_jacobian(j);
        xyzz_add(acc, v);
    }
    jac_t r = xyzz_to_jacobian(acc);
    char* o = reinterpret_cast<char*>(out);
    fq_store(o, r.x); fq_store(o + 32, r.y); fq_store(o + 64, r.z);
}

// ------------------------------------------------------------------------------------------ K0
// Repack arkworks records (stride / offsets from the Rust shim) into the 64-byte device format
// and an infinity byte mask.  Replaces the CPU `pack_affine_and_scalars`
// (utils/limbs_conversion.rs:311-378) -- which also dropped the infinity flag.
__global__ void __launch_bounds__(256) k_repack_bases(const uint8_t* __restrict__ raw, size_t stride, size_t x_off,
                                                      size_t y_off, size_t inf_off, uint32_t n,
                                                      uint64_t* __restrict__ out, uint8_t* __restrict__ inf_mask) {
    // 8 threads per point, one u64 each: coalesced 64-byte stores
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t i = gid >> 3;
    unsigned k = gid & 7;
    if (i >= n) return;
    const uint8_t* rec = raw + i * stride;
    const uint64_t* src = reinterpret_cast<const uint64_t*>(rec + (k < 4 ? x_off : y_off)) + (k & 3);
    bool inf = inf_off != (size_t)-1 && rec[inf_off] != 0;
    out[i * 8 + k] = inf ? 0ull : *src;
    if (k == 0 && inf_mask != nullptr) inf_mask[i] = inf ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_repack_scalars(const uint8_t* __restrict__ raw, size_t stride, uint32_t n,
                                                        uint64_t* __restrict__ out) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t i = gid >> 2;
    unsigned k = gid & 3;
    if (i >= n) return;
    out[i * 4 + k] = reinterpret_cast<const uint64_t*>(raw + i * stride)[k];
}
