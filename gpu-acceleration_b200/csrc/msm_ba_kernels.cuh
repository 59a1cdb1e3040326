// K3, large-n variant: bucket accumulation with BATCHED AFFINE additions.
//
// Replaces the same reference stage as k_accumulate (shader/cuzk/smvp.metal:14-107: one thread per bucket, a full
// Jacobian add-2007-bl = 16 multiplications per point).  k_accumulate brought that to 10 (XYZZ madd); here an addition
// of two AFFINE points costs
//     lambda = (y2 - y1) / (x2 - x1),  x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1        3 multiplications
// plus its share of ONE field inversion per batch through Montgomery's trick (3 multiplications per element), i.e.
// 6 + I/k per addition for a batch of k independent additions per lane, I = 65 multiplication times for the safegcd
// inversion of fq_inv.cuh (measured, profiles/r02b_inv_bench.jsonl; Fermat: 380).
//
// Where k independent additions per LANE come from: every thread owns one chunk of L consecutive entries of the sorted
// entry list (exactly k_accumulate's chunks, so head/tail/fix-up are unchanged) and reduces it as a TREE: round 0 adds
// neighbouring entries of the same bucket pairwise (~L/2 independent additions, one inversion), round 1 the results
// (~L/4), ... until fewer than `min_pairs` pairs remain; the handful of survivors goes through the XYZZ accumulator as
// before.  Intermediate points and the prefix products live in a per-thread slice of a global scratch buffer (they cannot
// fit on chip: k >= 64 needs kilobytes per lane), addressed lane-contiguously in 64-byte records like the base gathers.
//
// Special cases keep the operation complete without leaving the batch: an infinity operand (the (0,0) marker of
// k_repack_bases) passes the other operand through, P + P becomes a doubling whose denominator 2y joins the batch, and
// P + (-P) yields the marker; their denominators are replaced by 1.
#pragma once
#include "fq_inv.cuh"
#include "msm_kernels.cuh"

#define BA_THREADS 128
#ifndef BA_MIN_CTAS
#define BA_MIN_CTAS 3
#endif
#define BA_MAXL 512
#define BA_MASK_WORDS (BA_MAXL / 32)
#ifndef BA_PF
#define BA_PF 8   // L2 prefetch distance, in items
#endif

// Per-thread bit vectors in shared memory, layout [word][thread] (conflict-free: a thread only touches its column).
struct ba_bits {
    uint32_t* base;   // &smem[0 * BA_THREADS + threadIdx.x]
    __device__ __forceinline__ uint32_t word(int w) const { return base[w * BA_THREADS]; }
    __device__ __forceinline__ void set_word(int w, uint32_t v) const { base[w * BA_THREADS] = v; }
    __device__ __forceinline__ bool get(uint32_t i) const { return (base[(i >> 5) * BA_THREADS] >> (i & 31)) & 1u; }
    __device__ __forceinline__ void set(uint32_t i) const { base[(i >> 5) * BA_THREADS] |= 1u << (i & 31); }
};

struct ba_item_src {
    // round 0: gather through the entry list; later rounds: the thread's scratch records
    const affine_t* bases;
    const fq* xb;
    uint32_t n;
    const uint32_t* entries;   // entries + lo (gather mode) or nullptr
    const affine_t* buf;       // scratch records (buffer mode)
    __device__ __forceinline__ fq load_x(uint32_t i) const {
        if (entries) {
            const uint32_t idx = __ldg(entries + i) & 0x7fffffffu;
            const bool endo = idx >= n;
            const uint32_t b = endo ? idx - n : idx;
            return fq_load_nc(endo ? reinterpret_cast<const char*>(xb + b) : reinterpret_cast<const char*>(bases + b));
        }
        return fq_load(reinterpret_cast<const char*>(buf + i));
    }
    // L2 prefetch of item i (no destination register: the request only warms L2 for the loads a few iterations later)
    __device__ __forceinline__ void prefetch(uint32_t i, bool want_y) const {
        const char* px;
        const char* py;
        if (entries) {
            const uint32_t idx = __ldg(entries + i) & 0x7fffffffu;
            const bool endo = idx >= n;
            const uint32_t b = endo ? idx - n : idx;
            py = reinterpret_cast<const char*>(bases + b) + 32;
            px = endo ? reinterpret_cast<const char*>(xb + b) : py - 32;
        } else {
            px = reinterpret_cast<const char*>(buf + i);
            py = px + 32;
        }
        asm volatile("prefetch.global.L2 [%0];" ::"l"(px));
        if (want_y) asm volatile("prefetch.global.L2 [%0];" ::"l"(py));
    }
    __device__ __forceinline__ affine_t load(uint32_t i) const {
        affine_t p;
        if (entries) {
            const uint32_t e = __ldg(entries + i);
            p = load_pseudo_point(bases, xb, n, e & 0x7fffffffu);
            p.y = fq_cneg(p.y, (e >> 31) != 0);
        } else {
            const char* c = reinterpret_cast<const char*>(buf + i);
            p.x = fq_load(c);
            p.y = fq_load(c + 32);
        }
        return p;
    }
};

// The tree rounds call ONE shared multiplication body: the kernel runs warps in four different phases (pairing, forward,
// inversion, backward, survivors) and with every product inlined (~5 KB each) its code no longer fits the instruction
// caches (ncu: 17 % of the stall samples were "no instruction").
#ifdef BA_INLINE_MUL
__device__ __forceinline__ fq ba_mul(const fq& a, const fq& b) { return fq_mul(a, b); }
#else
__device__ __noinline__ fq ba_mul(const fq a, const fq b) { return fq_mul(a, b); }
#endif
__device__ __forceinline__ bool ba_is_inf(const affine_t& p) { return fq_is_zero(p.x) && fq_is_zero(p.y); }

// Denominator of the addition a + b inside a batch (1 for the cases that need no division).
// kind: 0 generic, 1 a is infinity (result b), 2 b is infinity (result a), 3 doubling, 4 cancellation (result infinity)
__device__ __forceinline__ int ba_special_den(const affine_t& a, const affine_t& b, fq& den) {
    den = fq_one();
    if (ba_is_inf(a)) return 1;
    if (ba_is_inf(b)) return 2;
    if (fq_eq(a.y, b.y) && !fq_is_zero(a.y)) {
        den = fq_dbl(a.y);
        return 3;
    }
    return 4;
}

// One tree round over m items of `src` with segment-start bits `bin`: writes the pairwise sums / passed-through items to
// dst[0 .. m_out) and their segment-start bits to `bout`.  Returns m_out.
// Three passes: (1) bits only -- greedy pairing inside each segment, output segment bits; (2) forward over the pairs:
// denominators and exclusive prefix products, the operands of pair q+1 requested before the product of pair q;
// (3) after the inversion, backward over the outputs: running inverse, slopes, results, again one pair ahead.
struct ba_pair_ops {
    affine_t a, b;
    fq pre;
};
__device__ __forceinline__ uint32_t ba_round(const ba_item_src& src, uint32_t m, const ba_bits& bin, const ba_bits& bout,
                                             const ba_bits& pairm, affine_t* __restrict__ dst, fq* __restrict__ prefix) {
    // ---- (1) pairing decisions
    uint32_t m_out = 0;
    {
        uint32_t o = 0, pw = 0, bw = 0, i = 0;
        while (i < m) {
            const bool pair = (i + 1 < m) && !bin.get(i + 1);
            if (i == 0 || bin.get(i)) bw |= 1u << (o & 31);
            if (pair) pw |= 1u << (o & 31);
            i += pair ? 2 : 1;
            o++;
            if ((o & 31) == 0) {
                pairm.set_word((o >> 5) - 1, pw);
                bout.set_word((o >> 5) - 1, bw);
                pw = 0;
                bw = 0;
            }
        }
        if (o & 31) {
            pairm.set_word(o >> 5, pw);
            bout.set_word(o >> 5, bw);
        }
        m_out = o;
    }
    // ---- (2) forward: denominators and exclusive prefix products
    fq prod = fq_one();
    uint32_t pc = 0;
    {
        uint32_t o = 0, i = 0;
        while (o < m_out && !pairm.get(o)) { o++; i++; }
        fq x1n = fq_zero(), x2n = fq_zero();
        if (o < m_out) { x1n = src.load_x(i); x2n = src.load_x(i + 1); }
        while (o < m_out) {
            const fq x1 = x1n, x2 = x2n;
            const uint32_t ci = i;
            o++;
            i += 2;
            while (o < m_out && !pairm.get(o)) { o++; i++; }
            if (o < m_out) { x1n = src.load_x(i); x2n = src.load_x(i + 1); }
            fq den = fq_sub(x2, x1);
            // x1 == x2: doubling or cancellation; x == 0: possibly the (0,0) infinity marker
            if (fq_is_zero(den) || fq_is_zero(x1) || fq_is_zero(x2)) {
                affine_t a = src.load(ci), b = src.load(ci + 1);
                if (ba_is_inf(a) || ba_is_inf(b) || fq_is_zero(den)) ba_special_den(a, b, den);
            }
            fq_store(prefix + pc, prod);
            prod = ba_mul(prod, den);
            pc++;
        }
    }
    fq inv = fq_inv_by(prod);
    // ---- (3) backward: running inverse, results
    {
        uint32_t i = m, oo = m_out;
        // walk down to the next pair, copying passed-through items on the way; returns false when none is left
        auto next_pair = [&](ba_pair_ops& ops) -> bool {
            while (oo > 0) {
                oo--;
                if (pairm.get(oo)) {
                    i -= 2;
                    pc--;
                    ops.a = src.load(i);
                    ops.b = src.load(i + 1);
                    ops.pre = fq_load(prefix + pc);
                    return true;
                }
                i -= 1;
                const affine_t a = src.load(i);
                char* c = reinterpret_cast<char*>(dst + oo);
                fq_store(c, a.x);
                fq_store(c + 32, a.y);
            }
            return false;
        };
        ba_pair_ops nxt;
        bool have = next_pair(nxt);
        while (have) {
            const ba_pair_ops cur = nxt;
            const uint32_t co = oo;
            have = next_pair(nxt);
            fq den = fq_sub(cur.b.x, cur.a.x);
            int kind = 0;
            if (fq_is_zero(den) || fq_is_zero(cur.a.x) || fq_is_zero(cur.b.x)) {
                if (ba_is_inf(cur.a) || ba_is_inf(cur.b) || fq_is_zero(den)) kind = ba_special_den(cur.a, cur.b, den);
            }
            const fq dinv = ba_mul(inv, cur.pre);
            inv = ba_mul(inv, den);
            affine_t r;
            if (kind == 0 || kind == 3) {
                fq num;
                if (kind == 0) {
                    num = fq_sub(cur.b.y, cur.a.y);
                } else {
                    fq xx = ba_mul(cur.a.x, cur.a.x);
                    num = fq_add(fq_dbl(xx), xx);
                }
                const fq lam = ba_mul(num, dinv);
                r.x = fq_sub(fq_sub(ba_mul(lam, lam), cur.a.x), cur.b.x);
                r.y = fq_sub(ba_mul(lam, fq_sub(cur.a.x, r.x)), cur.a.y);
            } else if (kind == 1) {
                r = cur.b;
            } else if (kind == 2) {
                r = cur.a;
            } else {
                r.x = fq_zero();
                r.y = fq_zero();
            }
            char* c = reinterpret_cast<char*>(dst + co);
            fq_store(c, r.x);
            fq_store(c + 32, r.y);
        }
    }
    return m_out;
}

// Scratch per thread: buffer A (3L/4 records), buffer B (9L/16 records), prefix products (L/2 field elements).
__host__ __device__ inline size_t ba_scratch_bytes_per_thread(uint32_t L) {
    return (size_t)(3 * L / 4 + 1) * 64 + (size_t)(9 * L / 16 + 1) * 64 + (size_t)(L / 2 + 1) * 32;
}

// Persistent grid: CTA b processes chunk groups b, b + gridDim.x, ... ; a group is BA_THREADS consecutive chunks.
// Same contract as k_accumulate for (buckets, head, tail): chunk t = absolute entry positions [t L, (t+1) L).
__global__ void __launch_bounds__(BA_THREADS, BA_MIN_CTAS) k_accumulate_ba(const affine_t* __restrict__ bases, const fq* __restrict__ xb,
                                                              uint32_t n, const uint32_t* __restrict__ entries,
                                                              const uint32_t* __restrict__ ends, uint32_t g_lo, uint32_t g_hi,
                                                              uint32_t L, uint32_t min_pairs, xyzz_t* __restrict__ buckets,
                                                              xyzz_t* __restrict__ head, xyzz_t* __restrict__ tail,
                                                              uint8_t* __restrict__ scratch) {
    __shared__ uint32_t sm_bits[3 * BA_MASK_WORDS * BA_THREADS];
    const ba_bits bitsA = {sm_bits + threadIdx.x};
    const ba_bits bitsB = {sm_bits + BA_MASK_WORDS * BA_THREADS + threadIdx.x};
    const ba_bits pairm = {sm_bits + 2 * BA_MASK_WORDS * BA_THREADS + threadIdx.x};
    const uint32_t P0 = g_lo ? ends[g_lo - 1] : 0;
    const uint32_t P1 = ends[g_hi - 1];
    uint8_t* my = scratch + ((size_t)blockIdx.x * BA_THREADS + threadIdx.x) * ba_scratch_bytes_per_thread(L);
    affine_t* bufA = reinterpret_cast<affine_t*>(my);
    affine_t* bufB = bufA + (3 * L / 4 + 1);
    fq* prefix = reinterpret_cast<fq*>(bufB + (9 * L / 16 + 1));
    const uint32_t mask_words = (L + 31) / 32;
    for (uint64_t grp = blockIdx.x;; grp += gridDim.x) {
        const uint64_t t64 = (uint64_t)(P0 / L) + grp * BA_THREADS + threadIdx.x;
        if ((uint64_t)(P0 / L) * L + grp * BA_THREADS * L >= P1) break;   // uniform: the whole group is past the end
        const uint64_t clo64 = t64 * L;
        if (clo64 >= P1) continue;
        const uint32_t t = (uint32_t)t64;
        const uint32_t clo = (uint32_t)clo64;
        const uint32_t chi = (uint32_t)min((uint64_t)0xffffffffu, clo64 + L);
        const uint32_t lo = max(clo, P0), hi = min(chi, P1);
        if (lo >= hi) continue;
        // smallest g in [g_lo, g_hi) with ends[g] > lo
        uint32_t a = g_lo, b = g_hi - 1;
        while (a < b) {
            uint32_t mid = (a + b) >> 1;
            if (__ldg(ends + mid) > lo) b = mid; else a = mid + 1;
        }
        const uint32_t g0 = a;
        const uint32_t bstart0 = g0 ? __ldg(ends + g0 - 1) : 0;
        uint32_t m = hi - lo;
        // segment-start bits of the chunk's entries
        for (uint32_t w = 0; w < mask_words; w++) bitsA.set_word(w, 0);
        bitsA.set(0);
        for (uint32_t g = g0;; g++) {
            const uint32_t e = __ldg(ends + g);
            if (e >= hi) break;
            bitsA.set(e - lo);
        }
        ba_item_src src = {bases, xb, n, entries + lo, nullptr};
        ba_bits bcur = bitsA, bnext = bitsB;
        affine_t* dst = bufA;
        for (;;) {
            uint32_t nseg = 0;
            for (uint32_t w = 0; w < (m + 31) / 32; w++) {
                uint32_t v = bcur.word(w);
                if ((w + 1) * 32 > m) v &= (m & 31) ? ((1u << (m & 31)) - 1) : 0xffffffffu;
                nseg += __popc(v);
            }
            // at least (m - nseg) / 2 pairs; a round must remove a quarter of the items (bounds the scratch buffers)
            if (2 * nseg > m || (m - nseg) / 2 < min_pairs) break;
            m = ba_round(src, m, bcur, bnext, pairm, dst, prefix);
            src.entries = nullptr;
            src.buf = dst;
            dst = dst == bufA ? bufB : bufA;
            const ba_bits tmp = bcur;
            bcur = bnext;
            bnext = tmp;
        }
        // ---- survivors: XYZZ accumulation, bucket by bucket (k_accumulate's epilogue)
        uint32_t g = g0;
        bool first_seg = true;
        xyzz_t acc = xyzz_inf();
        for (uint32_t j = 0; j < m; j++) {
            if (j > 0 && bcur.get(j)) {
                xyzz_t* d = (first_seg && bstart0 < clo) ? head + t : buckets + g;
                xyzz_store(d, acc);
                first_seg = false;
                uint32_t prev = __ldg(ends + g);
                do { g++; } while (__ldg(ends + g) == prev);   // next non-empty bucket
                acc = xyzz_inf();
            }
            affine_t p = src.load(j);
            if (!ba_is_inf(p)) xyzz_madd(acc, p);
        }
        const uint32_t bend = __ldg(ends + g);
        xyzz_t* d;
        if (first_seg && bstart0 < clo) d = head + t;
        else if (bend > chi) d = tail + t;
        else d = buckets + g;
        xyzz_store(d, acc);
    }
}
