// K1 + K2 as a SHARED-MEMORY RADIX PARTITION (the sort of the sparse-matrix transpose, for large inputs).
//
// Replaces the reference's serial per-window counting sort (/root/reference/mopro-msm/src/msm/metal_msm/shader/cuzk/
// transpose.metal:8-65: one thread per window walks all columns three times) and, from 2^22 digits, this engine's own
// first version (k_decompose's global histogram atomics with returned ranks + k_scatter_ranked), whose cost was one L2
// atomic WITH RETURN per digit and one scattered 4-byte store per digit into an entry list far larger than L2.
//
// The global key of a digit is g = w * wstride + |d| (window-major bucket slots, msm_kernels.cuh).  Each window's
// magnitudes are cut into partitions of 2^shift consecutive buckets (<= 4096 per window, <= 16384 in total, ~16 K digits each):
//   K1  k_decompose_count  a CTA walks a tile of scalars, writes the digits and histograms them over all partitions in
//                          SHARED memory; one global reduction (no return value) per (CTA, non-empty partition).
//   K2a k_pscan            one CTA: exclusive scan of the partition sizes -> partition bases and write cursors; list of
//                          (partition, slice) work items of the heavy partitions.
//   K2b k_partition        a CTA takes 8192 digits of ONE window (coalesced), counting-sorts them by partition inside shared
//                          memory, claims a run in every partition with ONE global atomic per (CTA, partition) and writes
//                          (entry, key inside the partition) to the staging arrays as contiguous runs.
//   K2c k_place            a CTA owns a partition (<= 4096 buckets): histogram of the keys and block scan in shared memory ->
//                          bucket ends; up to 20480 digits are placed into a shared-memory image of the output and written as
//                          one contiguous block, larger partitions directly into their L2-resident output range.
//       k_place_heavy      partitions far above their share (skewed scalars): one CTA per slice of 32768 digits, positions
//                          from global per-bucket counts and cursors.
// No `ranks` array, no global atomic per digit, no scattered store outside an L2-resident range.  HBM traffic per digit:
// digit 2-4 B written + read, staging 6 B written + read, entry 4 B written (~22-26 B against 16 B + 218 M atomics).
// Any scalar distribution is handled: sizes are counted, never assumed.
#pragma once
#include "msm_kernels.cuh"

#define PSORT_TILE 8192        // digits per k_partition CTA (512 threads x 16)
#define PART_THREADS 512
#define PSORT_MAX_NP 16384     // partitions in total (64 KB of shared-memory counters in K1)
#define PSORT_MAX_NPW 4096     // partitions per window (k_partition's shared-memory counters)
#define PSORT_MAX_SLOTS 4096   // buckets per partition (2^shift)
#define PLACE_THREADS 1024
#define PLACE_CAP 20480        // digits a k_place CTA sorts inside shared memory (80 KB); larger partitions are placed in HBM
// Skewed scalars (witness vectors: many zeros, ones, repeated values) put a large share of a window into one partition --
// often into ONE bucket, which no finer partition grid can split.  A partition above max(PSORT_HEAVY, 4 x the mean) digits is not given
// to a single CTA: k_partition additionally counts its digits per BUCKET with (warp-aggregated) global atomics, and
// k_place_heavy cuts it into slices of PSORT_SLICE digits, one CTA each: shared-memory histogram of the slice, one global
// atomic per (slice, non-empty bucket) on the bucket's cursor, placement with shared-memory cursors.
#define PSORT_HEAVY 65536
#define PSORT_SLICE 32768
#define PSORT_MAX_W 64

// Partition grid.  Windows 0 .. W-2 are cut into npw partitions of 2^shift magnitudes.  The TOP window is narrower (its
// digit has bits - c (W-1) bits, e.g. 14 of 20 at 254 / 20): all its digits fall on magnitudes <= top_used, so it gets
// its own, finer shift_top and npw_top partitions over [1, npw_top << shift_top] -- otherwise a handful of partitions
// would hold 2^(c-1) / top_used times their share.  Magnitudes above that range cannot occur in the top window (the
// scalar is canonical, < r < 2^254, resp. |k| < 2^127 after the GLV split); they are clamped into the last partition so
// that every access stays in bounds.  In precomputed-table mode all windows feed ONE bucket set (shared_set).
struct psort_shape {
    uint32_t shift, npw;          // log2 buckets per partition and partitions per window
    uint32_t shift_top, npw_top;  // the same for the top window
    uint32_t np;                  // partitions in total
    uint32_t top;                 // index of the top window (W - 1), or 0xffffffff when it is not special (shared_set)
    uint32_t shared_set;
    uint32_t heavy;               // a partition with more digits than this is cut into slices (k_place_heavy): max(PSORT_HEAVY, 4 x the mean)
};
// (window, magnitude - 1) -> partition index; key = bucket index inside the partition
__device__ __forceinline__ uint32_t psort_map(const psort_shape& ps, uint32_t w, uint32_t m1, uint32_t& key) {
    if (w == ps.top) {
        uint32_t q = m1 >> ps.shift_top;
        key = m1 & ((1u << ps.shift_top) - 1);
        if (q >= ps.npw_top) { q = ps.npw_top - 1; key = (1u << ps.shift_top) - 1; }
        return ps.top * ps.npw + q;
    }
    key = m1 & ((1u << ps.shift) - 1);
    return (ps.shared_set ? 0u : w * ps.npw) + (m1 >> ps.shift);
}

// exclusive scan over THREADS * PER values, PER consecutive ones per thread; returns the thread's exclusive prefix
template <int THREADS>
__device__ __forceinline__ uint32_t psort_block_scan(uint32_t mine, uint32_t* warp_sums, uint32_t* total) {
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(MSM_FULL_MASK, inc, o);
        if (lane >= (unsigned)o) inc += y;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = lane < THREADS / 32 ? warp_sums[lane] : 0, winc = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(MSM_FULL_MASK, winc, o);
            if (lane >= (unsigned)o) winc += y;
        }
        if (lane < THREADS / 32) warp_sums[lane] = winc - ws;
        if (lane == 31) warp_sums[32] = winc;
    }
    __syncthreads();
    if (total) *total = warp_sums[32];
    return warp_sums[wid] + inc - mine;
}

// K1: tile of `tile_pts` scalars per CTA (256 threads), digits out, partition histogram in shared memory.
template <typename DigitT, bool GLV>
__global__ void __launch_bounds__(256) k_decompose_count(const uint4* __restrict__ scalars, const uint8_t* __restrict__ inf_mask,
                                                         uint32_t n, int c, int W, uint32_t tile_pts, psort_shape ps,
                                                         DigitT* __restrict__ digits, uint32_t* __restrict__ part_count) {
    extern __shared__ uint32_t sm_cnt[];
    for (uint32_t p = threadIdx.x; p < ps.np; p += blockDim.x) sm_cnt[p] = 0;
    __syncthreads();
    const uint32_t lo = blockIdx.x * tile_pts;
    // the trip count is uniform across the CTA (decompose_scalar runs warp-collectively for the old path's sake)
    for (uint32_t off = 0; off < tile_pts; off += blockDim.x) {
        const uint32_t i = lo + off + threadIdx.x;
        uint32_t t[8];
        const bool valid = load_scalar_canonical(scalars, inf_mask, n, i, t);
        decompose_scalar<DigitT, GLV>(t, valid, i, n, c, W, digits, [&](int w, size_t col, int d, uint32_t mag, bool live) {
            (void)col; (void)d;
            uint32_t key;
            if (live) atomicAdd(&sm_cnt[psort_map(ps, (uint32_t)w, mag - 1, key)], 1u);
        });
    }
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < ps.np; p += blockDim.x) {
        const uint32_t v = sm_cnt[p];
        if (v) atomicAdd(part_count + p, v);
    }
}

// K2a: part_base[p] = exclusive prefix of part_count (part_base[np] = total), part_cursor = part_base.  One CTA.
// heavy[0] = number of (partition, slice) work items of heavy partitions, heavy[1 + w] = window w has a heavy partition
__global__ void __launch_bounds__(1024) k_pscan(const uint32_t* __restrict__ part_count, uint32_t np, uint32_t npw, uint32_t top,
                                                uint32_t heavy_min,
                                                uint32_t* __restrict__ part_base, uint32_t* __restrict__ part_cursor,
                                                uint32_t* __restrict__ heavy, uint2* __restrict__ heavy_items) {
    __shared__ uint32_t warp_sums[33];
    if (threadIdx.x <= PSORT_MAX_W) heavy[threadIdx.x] = 0;
    __syncthreads();
    const unsigned tid = threadIdx.x;
    uint32_t v[PSORT_MAX_NP / 1024];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < PSORT_MAX_NP / 1024; k++) {
        const uint32_t p = tid * (PSORT_MAX_NP / 1024) + k;
        v[k] = p < np ? part_count[p] : 0;
        s += v[k];
    }
    uint32_t total;
    uint32_t ex = psort_block_scan<1024>(s, warp_sums, &total);
#pragma unroll
    for (int k = 0; k < PSORT_MAX_NP / 1024; k++) {
        const uint32_t p = tid * (PSORT_MAX_NP / 1024) + k;
        if (p < np) { part_base[p] = ex; part_cursor[p] = ex; }
        ex += v[k];
    }
    if (tid == 0) part_base[np] = total;
    // work items (partition, slice) of the heavy partitions
    uint32_t ns = 0;
#pragma unroll
    for (int k = 0; k < PSORT_MAX_NP / 1024; k++)
        if (v[k] > heavy_min) ns += (v[k] + PSORT_SLICE - 1) / PSORT_SLICE;
    __syncthreads();   // warp_sums is reused
    uint32_t nitems;
    uint32_t at = psort_block_scan<1024>(ns, warp_sums, &nitems);
#pragma unroll
    for (int k = 0; k < PSORT_MAX_NP / 1024; k++) {
        if (v[k] > heavy_min) {
            const uint32_t p = tid * (PSORT_MAX_NP / 1024) + k;
            const uint32_t cnt = (v[k] + PSORT_SLICE - 1) / PSORT_SLICE;
            for (uint32_t sl = 0; sl < cnt; sl++) heavy_items[at + sl] = make_uint2(p, sl);
            at += cnt;
            heavy[1 + min(min(p / npw, top), (uint32_t)PSORT_MAX_W - 1)] = 1;   // the top window's (finer) partitions start at top * npw
        }
    }
    if (tid == 0) heavy[0] = nitems;
}

// K2b: one CTA = PSORT_TILE consecutive digits of one window.  The tile is sorted by partition INSIDE shared memory, so
// that the staging stores of a warp cover a few contiguous runs instead of 32 different sectors.
// dynamic shared memory: off[npw_w] | delta[npw_w] | tile entries (u32) | tile keys (u16) | tile partition ids (u16)
template <typename DigitT>
__global__ void __launch_bounds__(PART_THREADS, 2) k_partition(const DigitT* __restrict__ digits, uint32_t n_eff, uint32_t tiles_per_window,
                                                            uint32_t istride, psort_shape ps, uint32_t* __restrict__ part_cursor,
                                                            uint32_t* __restrict__ stage_e, uint16_t* __restrict__ stage_k,
                                                            const uint32_t* __restrict__ part_base, const uint32_t* __restrict__ heavy,
                                                            uint32_t wstride, uint32_t* __restrict__ bcount) {
    extern __shared__ uint32_t sm_dyn[];
    __shared__ uint32_t warp_sums[33];
    constexpr int PER_T = PSORT_TILE / PART_THREADS;          // digits per thread
    constexpr int PER_Q = PSORT_MAX_NPW / PART_THREADS;       // partition counters per thread (scan)
    const uint32_t w = blockIdx.x / tiles_per_window;
    const uint32_t tile = blockIdx.x - w * tiles_per_window;
    const uint32_t col0 = tile * PSORT_TILE;
    const uint32_t npw = w == ps.top ? ps.npw_top : ps.npw;
    uint32_t dummy;
    const uint32_t pbase = psort_map(ps, w, 0, dummy);
    const uint32_t npw_pad = (npw + PART_THREADS - 1) & ~(uint32_t)(PART_THREADS - 1);
    uint32_t* s_off = sm_dyn;
    uint32_t* s_delta = s_off + npw_pad;
    uint32_t* s_e = s_delta + npw_pad;
    uint16_t* s_k = reinterpret_cast<uint16_t*>(s_e + PSORT_TILE);
    uint16_t* s_q = s_k + PSORT_TILE;
    // heavy_window (uniform across the CTA): some partition of this window is heavy; then s_delta holds the per-partition flags
    // until the scan below overwrites it, and the shared-memory atomics are aggregated per warp (one bucket may hold every digit)
    const bool heavy_window = heavy[1 + (ps.shared_set ? 0u : min(w, (uint32_t)PSORT_MAX_W - 1))] != 0;
    for (uint32_t q = threadIdx.x; q < npw_pad; q += PART_THREADS) {
        s_off[q] = 0;
        if (heavy_window) s_delta[q] = (q < npw && part_base[pbase + q + 1] - part_base[pbase + q] > ps.heavy) ? 1u : 0u;
    }
    __syncthreads();
    int d[PER_T];
    const DigitT* src = digits + (size_t)w * n_eff;
#pragma unroll
    for (int k = 0; k < PER_T; k++) {
        const uint32_t col = col0 + k * PART_THREADS + threadIdx.x;
        d[k] = col < n_eff ? (int)src[col] : 0;
    }
    if (!heavy_window) {
#pragma unroll
        for (int k = 0; k < PER_T; k++) {
            if (d[k] != 0) {
                uint32_t key;
                atomicAdd(&s_off[psort_map(ps, w, (uint32_t)(d[k] < 0 ? -d[k] : d[k]) - 1, key) - pbase], 1u);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < PER_T; k++) {
            uint32_t q = 0xffffffffu, slot = 0xffffffffu;
            if (d[k] != 0) {
                uint32_t key;
                const uint32_t mag = (uint32_t)(d[k] < 0 ? -d[k] : d[k]);
                q = psort_map(ps, w, mag - 1, key) - pbase;
                slot = w * wstride + mag;
            }
            // one atomic per distinct bucket in the warp, on the partition counter and (heavy partitions) on the global bucket counter
            const unsigned peers = __match_any_sync(MSM_FULL_MASK, slot);
            if (slot != 0xffffffffu && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) {
                atomicAdd(&s_off[q], (uint32_t)__popc(peers));
                if (s_delta[q]) atomicAdd(bcount + slot, (uint32_t)__popc(peers));
            }
        }
    }
    __syncthreads();
    // exclusive scan of the counters (npw_pad / PART_THREADS consecutive ones per thread); one global atomic per non-empty
    // partition, all of a thread's atomics in flight together
    const uint32_t per = npw_pad / PART_THREADS;   // <= PER_Q
    uint32_t cnt[PER_Q], got[PER_Q];
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < PER_Q; k++) {
        cnt[k] = (uint32_t)k < per ? s_off[threadIdx.x * per + k] : 0;
        mine += cnt[k];
    }
#pragma unroll
    for (int k = 0; k < PER_Q; k++) got[k] = cnt[k] ? atomicAdd(part_cursor + pbase + threadIdx.x * per + k, cnt[k]) : 0;
    uint32_t total;
    uint32_t ex = psort_block_scan<PART_THREADS>(mine, warp_sums, &total);
#pragma unroll
    for (int k = 0; k < PER_Q; k++) {
        if ((uint32_t)k < per) {
            const uint32_t q = threadIdx.x * per + k;
            s_off[q] = ex;
            s_delta[q] = got[k] - ex;
            ex += cnt[k];
        }
    }
    __syncthreads();
    if (!heavy_window) {
#pragma unroll
        for (int k = 0; k < PER_T; k++) {
            if (d[k] == 0) continue;
            const uint32_t col = col0 + k * PART_THREADS + threadIdx.x;
            uint32_t key;
            const uint32_t q = psort_map(ps, w, (uint32_t)(d[k] < 0 ? -d[k] : d[k]) - 1, key) - pbase;
            const uint32_t pos = atomicAdd(&s_off[q], 1u);
            s_e[pos] = (col + w * istride) | (d[k] < 0 ? 0x80000000u : 0u);
            s_k[pos] = (uint16_t)key;
            s_q[pos] = (uint16_t)q;
        }
    } else {
        const unsigned lane = threadIdx.x & 31;
#pragma unroll
        for (int k = 0; k < PER_T; k++) {
            const uint32_t col = col0 + k * PART_THREADS + threadIdx.x;
            uint32_t key = 0, q = 0xffffffffu;
            if (d[k] != 0) q = psort_map(ps, w, (uint32_t)(d[k] < 0 ? -d[k] : d[k]) - 1, key) - pbase;
            const unsigned peers = __match_any_sync(MSM_FULL_MASK, q);   // one cursor atomic per distinct partition in the warp
            const unsigned leader = (unsigned)(__ffs(peers) - 1);
            uint32_t got = 0;
            if (q != 0xffffffffu && lane == leader) got = atomicAdd(&s_off[q], (uint32_t)__popc(peers));
            got = __shfl_sync(MSM_FULL_MASK, got, leader);
            if (q != 0xffffffffu) {
                const uint32_t pos = got + __popc(peers & ((1u << lane) - 1));
                s_e[pos] = (col + w * istride) | (d[k] < 0 ? 0x80000000u : 0u);
                s_k[pos] = (uint16_t)key;
                s_q[pos] = (uint16_t)q;
            }
        }
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < total; j += PART_THREADS) {
        const uint32_t dst = s_delta[s_q[j]] + j;
        stage_e[dst] = s_e[j];
        stage_k[dst] = s_k[j];
    }
}

// K2c: one CTA per partition (issued from the last to the first).  Partitions of up to PLACE_CAP digits are sorted inside
// shared memory and written out as one contiguous block; larger ones (skewed scalars) are placed directly in HBM.
// dynamic shared memory: cursors[PSORT_MAX_SLOTS] | sorted entries[PLACE_CAP]
__global__ void __launch_bounds__(PLACE_THREADS) k_place(const uint32_t* __restrict__ stage_e, const uint16_t* __restrict__ stage_k,
                                                         const uint32_t* __restrict__ part_base, psort_shape ps, uint32_t half,
                                                         uint32_t wstride, uint32_t* __restrict__ ends, uint32_t* __restrict__ entries) {
    extern __shared__ uint32_t sm_dyn[];
    __shared__ uint32_t warp_sums[33];
    uint32_t* sm_cur = sm_dyn;
    uint32_t* sm_out = sm_dyn + PSORT_MAX_SLOTS;
    const uint32_t p = ps.np - 1 - blockIdx.x;
    uint32_t w, q, shift;
    if (ps.shared_set) { w = 0; q = p; shift = ps.shift; }
    else if (p >= ps.top * ps.npw) { w = ps.top; q = p - ps.top * ps.npw; shift = ps.shift_top; }
    else { w = p / ps.npw; q = p - w * ps.npw; shift = ps.shift; }
    const uint32_t m0 = (q << shift) + 1;                             // first magnitude of the partition
    const uint32_t slots = min(1u << shift, half - (m0 - 1));         // buckets it really has
    const uint32_t base = part_base[p], cnt = part_base[p + 1] - base;
    const unsigned tid = threadIdx.x;
    if (cnt > ps.heavy) return;       // k_place_heavy's slices do the work (and write the bucket ends)
    for (uint32_t k = tid; k < PSORT_MAX_SLOTS; k += PLACE_THREADS) sm_cur[k] = 0;
    __syncthreads();
    const uint16_t* keys = stage_k + base;
    const uint32_t* ents = stage_e + base;
    for (uint32_t i0 = 0; i0 < cnt; i0 += 8 * PLACE_THREADS) {       // 8 independent loads in flight per thread
        uint32_t k8[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t i = i0 + u * PLACE_THREADS + tid;
            k8[u] = i < cnt ? keys[i] : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (k8[u] != 0xffffffffu) atomicAdd(&sm_cur[k8[u]], 1u);
    }
    __syncthreads();
    uint32_t v[PSORT_MAX_SLOTS / PLACE_THREADS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < PSORT_MAX_SLOTS / PLACE_THREADS; k++) {
        v[k] = sm_cur[tid * (PSORT_MAX_SLOTS / PLACE_THREADS) + k];
        s += v[k];
    }
    uint32_t ex = psort_block_scan<PLACE_THREADS>(s, warp_sums, nullptr);
    uint32_t* e = ends + (size_t)w * wstride + m0;
#pragma unroll
    for (int k = 0; k < PSORT_MAX_SLOTS / PLACE_THREADS; k++) {
        const uint32_t idx = tid * (PSORT_MAX_SLOTS / PLACE_THREADS) + k;
        sm_cur[idx] = ex;            // cursor of bucket idx inside the partition
        ex += v[k];
        if (idx < slots) e[idx] = base + ex;   // exclusive end of the bucket in the entry list
    }
    if (q == 0 && tid == 0) ends[(size_t)w * wstride] = base;   // slot 0 (magnitude 0 never occurs): start of the window
    if (w == ps.top && q + 1 == ps.npw_top) {
        // magnitudes the top window cannot reach: empty buckets at the very end of the entry list
        for (uint32_t m = m0 + slots + tid; m <= half; m += PLACE_THREADS) ends[(size_t)w * wstride + m] = base + cnt;
    }
    __syncthreads();
    const bool in_smem = cnt <= PLACE_CAP;
    for (uint32_t i0 = 0; i0 < cnt; i0 += 8 * PLACE_THREADS) {
        uint32_t k8[8], e8[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t i = i0 + u * PLACE_THREADS + tid;
            k8[u] = i < cnt ? keys[i] : 0xffffffffu;
            e8[u] = i < cnt ? ents[i] : 0;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (k8[u] == 0xffffffffu) continue;
            const uint32_t pos = atomicAdd(&sm_cur[k8[u]], 1u);
            if (in_smem) sm_out[pos] = e8[u];
            else entries[base + pos] = e8[u];
        }
    }
    if (in_smem) {
        __syncthreads();
        for (uint32_t i = tid; i < cnt; i += PLACE_THREADS) entries[base + i] = sm_out[i];
    }
}

// K2c for heavy partitions: CTA i takes work item i = (partition, slice) of k_pscan's list.  Bucket starts come from the global
// per-bucket counts (k_partition), every slice claims its range inside each bucket with ONE atomic on the bucket's cursor and
// places its digits with shared-memory cursors; all shared-memory atomics are aggregated per warp (a single bucket may hold the
// whole slice).  Slice 0 also writes the partition's bucket ends.
// dynamic shared memory: bucket starts[PSORT_MAX_SLOTS] | slice counts / cursors[PSORT_MAX_SLOTS]
__global__ void __launch_bounds__(PLACE_THREADS) k_place_heavy(const uint32_t* __restrict__ stage_e, const uint16_t* __restrict__ stage_k,
                                                               const uint32_t* __restrict__ part_base, psort_shape ps, uint32_t half,
                                                               uint32_t wstride, uint32_t* __restrict__ ends, uint32_t* __restrict__ entries,
                                                               const uint32_t* __restrict__ heavy, const uint2* __restrict__ heavy_items,
                                                               const uint32_t* __restrict__ bcount, uint32_t* __restrict__ bcursor) {
    extern __shared__ uint32_t sm_dyn[];
    __shared__ uint32_t warp_sums[33];
    uint32_t* sm_start = sm_dyn;
    uint32_t* sm_cur = sm_dyn + PSORT_MAX_SLOTS;
    const uint32_t nitems = heavy[0];
    const unsigned tid = threadIdx.x, lane = tid & 31;
    for (uint32_t item = blockIdx.x; item < nitems; item += gridDim.x) {   // uniform across the CTA
        const uint2 it = heavy_items[item];
        const uint32_t p = it.x, slice = it.y;
        uint32_t w, q, shift;
        if (ps.shared_set) { w = 0; q = p; shift = ps.shift; }
        else if (p >= ps.top * ps.npw) { w = ps.top; q = p - ps.top * ps.npw; shift = ps.shift_top; }
        else { w = p / ps.npw; q = p - w * ps.npw; shift = ps.shift; }
        const uint32_t m0 = (q << shift) + 1;
        const uint32_t slots = min(1u << shift, half - (m0 - 1));
        const uint32_t base = part_base[p], cnt = part_base[p + 1] - base;
        const uint32_t slot0 = w * wstride + m0;
        const uint16_t* keys = stage_k + base;
        const uint32_t* ents = stage_e + base;
        // bucket starts of the partition from the global counts
        uint32_t v[PSORT_MAX_SLOTS / PLACE_THREADS];
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < PSORT_MAX_SLOTS / PLACE_THREADS; k++) {
            const uint32_t idx = tid * (PSORT_MAX_SLOTS / PLACE_THREADS) + k;
            v[k] = idx < slots ? bcount[slot0 + idx] : 0;
            s += v[k];
            sm_cur[idx] = 0;
        }
        uint32_t ex = psort_block_scan<PLACE_THREADS>(s, warp_sums, nullptr);
#pragma unroll
        for (int k = 0; k < PSORT_MAX_SLOTS / PLACE_THREADS; k++) {
            const uint32_t idx = tid * (PSORT_MAX_SLOTS / PLACE_THREADS) + k;
            sm_start[idx] = ex;
            ex += v[k];
            if (slice == 0 && idx < slots) ends[(size_t)w * wstride + m0 + idx] = base + ex;
        }
        if (slice == 0 && q == 0 && tid == 0) ends[(size_t)w * wstride] = base;
        if (slice == 0 && w == ps.top && q + 1 == ps.npw_top)
            for (uint32_t m = m0 + slots + tid; m <= half; m += PLACE_THREADS) ends[(size_t)w * wstride + m] = base + cnt;
        __syncthreads();
        const uint32_t lo = slice * PSORT_SLICE, hi = min(cnt, lo + PSORT_SLICE);
        // pass 1: histogram of the slice
        for (uint32_t i0 = lo; i0 < hi; i0 += 8 * PLACE_THREADS) {
            uint32_t k8[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t i = i0 + u * PLACE_THREADS + tid;
                k8[u] = i < hi ? keys[i] : 0xffffffffu;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const unsigned peers = __match_any_sync(MSM_FULL_MASK, k8[u]);
                if (k8[u] != 0xffffffffu && lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(&sm_cur[k8[u]], (uint32_t)__popc(peers));
            }
        }
        __syncthreads();
        // claim the slice's range inside every non-empty bucket: cursor = bucket start + range start
        for (uint32_t k = tid; k < slots; k += PLACE_THREADS) {
            const uint32_t c = sm_cur[k];
            sm_cur[k] = sm_start[k] + (c ? atomicAdd(bcursor + slot0 + k, c) : 0u);
        }
        __syncthreads();
        // pass 2: placement
        for (uint32_t i0 = lo; i0 < hi; i0 += 8 * PLACE_THREADS) {
            uint32_t k8[8], e8[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t i = i0 + u * PLACE_THREADS + tid;
                k8[u] = i < hi ? keys[i] : 0xffffffffu;
                e8[u] = i < hi ? ents[i] : 0;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const unsigned peers = __match_any_sync(MSM_FULL_MASK, k8[u]);
                const unsigned leader = (unsigned)(__ffs(peers) - 1);
                uint32_t got = 0;
                if (k8[u] != 0xffffffffu && lane == leader) got = atomicAdd(&sm_cur[k8[u]], (uint32_t)__popc(peers));
                got = __shfl_sync(MSM_FULL_MASK, got, leader);
                if (k8[u] != 0xffffffffu) entries[base + got + __popc(peers & ((1u << lane) - 1))] = e8[u];
            }
        }
        __syncthreads();
    }
}
