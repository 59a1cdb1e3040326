// libb200msm.so -- host orchestration + C ABI (include/b200msm.h).
//
// Replaces, for the one hot path `metal_variable_base_msm`, the reference's
//   L4 public API            metal_msm.rs:642-695
//   L3 host orchestration    metal_msm.rs:40-262 (MetalMSMPipeline::execute_pipeline, 4 stage structs)
//   L2 host GPU runtime      metal_msm/host/{metal_wrapper,shader_manager,gpu}.rs
// with a persistent context: one stream and one grow-only buffer pool per device, every stage
// device-resident, ONE host synchronisation per MSM (the reference drains the queue and copies
// whole buffers to host Vecs >= 9 times per MSM: metal_wrapper.rs:131-132, gpu.rs:15-26).
// There is no CPU fallback anywhere in this file: without a CUDA device every entry point fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <string>
#include <vector>

#include "../../include/b200msm.h"
#include "../../include/b200math.h"
#include "msm_kernels.cuh"
#include "msm_ba_kernels.cuh"
#include "msm_psort_kernels.cuh"
#include "msm_g2_kernels.cuh"
#include "testkit_kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            return fail(_e == cudaErrorMemoryAllocation ? B200MSM_ENOMEM : B200MSM_ECUDA,         \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                      \
        }                                                                                         \
    } while (0)
// No exception may cross the C ABI (a Rust / ctypes caller cannot unwind through it): every int-returning entry point is a
// function-try-block closed by this handler.
#define B200_CATCH                                                                                              \
    catch (const std::bad_alloc&) { return fail(B200MSM_ENOMEM, "out of host memory"); }                        \
    catch (const std::exception& e) { return fail(B200MSM_ECUDA, std::string("internal error: ") + e.what()); } \
    catch (...) { return fail(B200MSM_ECUDA, "internal error"); }
#ifndef B200MSM_BUILD_ID
#define B200MSM_BUILD_ID "unknown"
#endif
#define RET_TRY(expr)              \
    do {                           \
        int _r = (expr);           \
        if (_r != B200MSM_OK) return _r; \
    } while (0)

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return B200MSM_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8;  // head-room so sweeps do not reallocate every size
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            p = nullptr;
            return fail(B200MSM_ENOMEM, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
        }
        cap = want;
        return B200MSM_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// A few helper threads for host-side memcpy: arkworks callers hand over ordinary (pageable) Vec memory, which the CUDA
// driver would stage through one internal buffer with a single-threaded copy (~10 GB/s).  The library stages it itself:
// several threads copy a chunk into a pinned ring slot while the DMA engine drains the previous slots.
class CopyPool {
  public:
    explicit CopyPool(int workers) {
        for (int k = 0; k < workers; k++) th_.emplace_back([this, k] { run(k); });
        jobs_.resize(workers);
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (auto& t : th_) t.join();
    }
    // blocking parallel memcpy; concurrent callers (one enqueue thread per device) take turns
    void copy(void* dst, const void* src, size_t bytes) {
        std::lock_guard<std::mutex> turn(call_mu_);
        const size_t parts = th_.size() + 1;
        const size_t per = ((bytes + parts - 1) / parts + 4095) & ~(size_t)4095;
        if (bytes < (1u << 20) || th_.empty()) {
            std::memcpy(dst, src, bytes);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (size_t k = 0; k < th_.size(); k++) {
                const size_t off = std::min(bytes, (k + 1) * per), end = std::min(bytes, (k + 2) * per);
                jobs_[k] = {(uint8_t*)dst + off, (const uint8_t*)src + off, end - off};
            }
            pending_ = (int)th_.size();
            generation_++;
        }
        cv_work_.notify_all();
        std::memcpy(dst, src, std::min(bytes, per));
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [this] { return pending_ == 0; });
    }

  private:
    struct Job { uint8_t* dst; const uint8_t* src; size_t bytes; };
    void run(int id) {
        unsigned long long seen = 0;
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_work_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                j = jobs_[id];
            }
            if (j.bytes) std::memcpy(j.dst, j.src, j.bytes);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) cv_done_.notify_one();
            }
        }
    }
    std::vector<std::thread> th_;
    std::vector<Job> jobs_;
    std::mutex mu_, call_mu_;
    std::condition_variable cv_work_, cv_done_;
    unsigned long long generation_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

constexpr int MAX_SLICES = 8;
constexpr int STAGE_SLOTS = 4;
constexpr size_t STAGE_BYTES = 8u << 20;
enum { EV_START = 0, EV_H2D, EV_DECOMP, EV_SORT, EV_ACC, EV_RED, EV_COUNT };

struct DevState {
    int ordinal = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool owns_stream = true;
    cudaStream_t stream2 = nullptr;  // high-priority side stream: reduce chains (window groups, batches), slice uploads
    cudaStream_t stream3 = nullptr;  // copy stream of the batch pipeline (scalar uploads ahead of the arithmetic)
    cudaStream_t stream4 = nullptr;  // high-priority sort stream of the sliced host pipeline: K1 + K2 of slice k+1 under K3 of slice k
    // Feedback for the slice plan of the host-buffer call: device timestamps around the uploads and at the end of the last
    // sliced call.  When the transfer took most of the call (several ranks sharing the host's H2D bandwidth), the next call
    // uses equal slices (see the host entry point).
    cudaEvent_t ev_cp_begin = nullptr, ev_cp_end = nullptr, ev_call_end = nullptr;
    bool cp_valid = false;
    double cp_compute_est_ms = 0;   // what the arithmetic of that call takes with resident inputs (model below)
    bool copy_bound = false;
    cudaEvent_t ev_sorted[MAX_SLICES] = {};
    cudaEvent_t ev_acc[8] = {};
    cudaEvent_t ev_done = nullptr;
    cudaEvent_t ev_bases = nullptr;
    cudaEvent_t ev[EV_COUNT] = {};
    Buf digits, ranks, skeys, parts, bslots, chunkg, giant, ends, wtotal, entries, buckets, head, tail, wpart, out, longlist, xb, redbuf, ba_scratch;
    int ba_ctas_per_sm = 0;   // occupancy of k_accumulate_ba (queried once)
    int hw_sm_count = 148;    // the device's SM count (sm_count may be overridden by option "sm_count")
    int* occ_flag = nullptr;  // mapped host flag of the SM blocker (test kit)
    unsigned int* occ_started = nullptr;
    cudaStream_t occ_stream = nullptr;
    Buf raw, bases, infmask, scalars_raw, scalars, scalars_alt, partials;
    Buf g2_bases, g2_buckets, g2_head, g2_tail, g2_wpart, g2_out, g2_redbuf;   // G2 MSM (Fq2 points: twice the bytes of G1)
    // Work sets of slices 1.. of a sliced host-input MSM (slice 0 uses the buffers above)
    struct SliceWork {
        Buf digits, ranks, skeys, parts, bslots, chunkg, giant, ends, wtotal, entries, buckets, head, tail, longlist;
        Buf g2_buckets, g2_head, g2_tail;
    } extra[MAX_SLICES - 1];
    cudaEvent_t ev_slice[2 * MAX_SLICES] = {};   // [2k] scalars of slice k on the device, [2k+1] bases
    // pinned staging ring for uploads from pageable host memory (pool: the context's copy threads)
    CopyPool* pool = nullptr;
    uint8_t* stage[STAGE_SLOTS] = {};
    cudaEvent_t stage_ev[STAGE_SLOTS] = {};
    bool stage_used[STAGE_SLOTS] = {};
    unsigned stage_next = 0;
};

// The per-(sub-)MSM scratch one sort + accumulate + fix-up pass works on.
struct WorkView {
    void *digits, *ranks, *ends, *wtotal, *entries, *buckets, *head, *tail, *longlist, *skeys, *parts, *chunkg, *bslots, *giant;
};
WorkView view_main(DevState& d) {
    return {d.digits.p, d.ranks.p, d.ends.p, d.wtotal.p, d.entries.p, d.buckets.p, d.head.p, d.tail.p, d.longlist.p, d.skeys.p, d.parts.p, d.chunkg.p, d.bslots.p, d.giant.p};
}
WorkView view_slice(DevState& d, int k) {
    if (k == 0) return view_main(d);
    auto& e = d.extra[k - 1];
    return {e.digits.p, e.ranks.p, e.ends.p, e.wtotal.p, e.entries.p, e.buckets.p, e.head.p, e.tail.p, e.longlist.p, e.skeys.p, e.parts.p, e.chunkg.p, e.bslots.p, e.giant.p};
}

// Shape of one single-device MSM.
struct Plan {
    uint32_t n = 0;
    int c = 0, W = 0;
    uint32_t half = 0, nb = 0, G = 0;
    bool wide_digits = false;
    uint32_t L = 0, nchunks = 0;
    uint32_t bpw = 0, log2Bsz = 0;
    int ngroups = 1;
    bool glv = false;
    uint32_t n_eff = 0;  // pseudo-points: n, or 2n with the GLV split
    // precomputed-table mode: every digit window is served by table[w][i] = 2^(c*w) P_i, so all W windows of digits feed
    // ONE set of buckets (Wb = 1) and the Horner step disappears.  Wb = W otherwise.
    uint32_t tstride = 0;  // points per table window (0 = not in table mode)
    int Wb = 0;
    // cooperative bucket reduce: levels of the recursive weighted sum (k_reduce_level)
    int red_nl = 0;
    uint32_t red_lb[8] = {}, red_ctas[8] = {};
    size_t red_slots = 0;  // XYZZ slots needed for the level buffers
    bool coop_reduce = true;
    bool rowcol = false;   // K4 through row / column sums (k_rowcol_sums + k_reduce_rowcol) instead of the level recursion
    uint32_t rc_ra = 0, rc_cb = 0;
    bool ranked = true;   // ranked sort: ranks from the histogram pass, scatter without atomics
    bool fix_chunks = false;   // chunk-boundary fix-up with one thread per chunk (k_fixup_chunks) instead of one per bucket
    bool psort = false;   // partitioned sort (msm_psort_kernels.cuh): shared-memory radix partition, no per-digit global atomic
    psort_shape ps = {};
    uint32_t psort_tile_pts = 256;
    bool ba = false;      // K3 with batched affine additions (k_accumulate_ba): chunk-local tree reduction, one inversion per round
    uint32_t ba_min_pairs = 24;
};

// bits = 254 for plain scalars (< r < 2^254), 127 for the GLV half-scalars (|k| < 2^127)
int num_windows_for(int c, int bits = 254) {
    int W = (bits + c - 1) / c;
    if (bits - c * (W - 1) == c) W += 1;  // no head-room for the top carry
    return W;
}

// (GLV?, window size) per (points on this device, SM count).  Replaces the reference's hand table
// (metal_msm.rs:661-673: 8 / 13 / 15 / 16) and its unused cuZK cost model (utils/window_size_optimizer.rs:38-76, which
// takes the core count: :57-76).  Every breakpoint is MEASURED, on the full chip (148 SMs) and with half of the SMs taken
// away (74; tools/autotune_sweep.py with b200msm_testkit_occupy_sms -- the situation of a MIG slice or a green context):
// profiles/r02m_autotune_sweep_148_74sm.jsonl holds the total device time of every admissible (split, c) at n = 2^12..2^24
// for both, profiles/r02m_autotune_ncu_148sm.json the ncu counters of each choice (integer-pipe utilisation of
// k_accumulate, achieved HBM GB/s of k_decompose / k_scatter_ranked).  What the two sweeps show:
//   * the throughput trade-offs do not depend on the SM count (accumulate ~ W n / SMs against reduce ~ W 2^c / SMs): the
//     GLV split wins up to 2^22 points (half the windows for K4, half the doublings of K5), above it plain windows win
//     because ceil(254/c) wastes less than 2 ceil(127/c); c = 17 up to 2^23, c = 20 from 2^24 -- for 148 and for 74 SMs;
//   * the small-size breakpoints (c = 8 / 13 / 16 under GLV) are where the bucket reduce turns from latency-bound (time
//     independent of the SM count) to throughput-bound (time ~ buckets / SMs), so they move with the SM count: they are
//     looked up with n scaled by SMs / 148 (74 SMs: 2^16 points take c = 13, measured 2.36 ms against 2.56 for c = 16;
//     148 SMs: c = 16, 1.03 against 1.09).
// Under GLV only window sizes whose TOP digit keeps >= 6 bits are admissible: a 1-bit top window (c = 9, 14, 18, 21)
// means a handful of buckets holding n/2 points each.
int round_log2(double x) {
    int lg = 0;
    while (lg < 62 && std::ldexp(1.0, lg + 1) <= x) lg++;
    if (x - std::ldexp(1.0, lg) > std::ldexp(1.0, lg) / 2) lg++;   // nearest power of two
    return lg;
}
void auto_policy(size_t n, int sm_count, bool glv_allowed, bool* glv, int* c) {
    const int sms = sm_count > 0 ? sm_count : 148;
    const int lg = round_log2((double)n);
    const int lg_small = round_log2(std::max(1.0, (double)n * sms / 148.0));
    if (glv_allowed && lg <= 22) {
        *glv = true;
        *c = lg_small <= 12 ? 8 : lg_small <= 15 ? 13 : 16;
        return;
    }
    *glv = false;
    *c = lg_small <= 10 ? 8 : lg_small <= 14 ? 12 : lg_small == 15 ? 13 : lg <= 20 ? 16 : lg <= 23 ? 17 : 20;
}
// Window size of the precomputed table for n registered points (one bucket set, no Horner step: the reduce costs
// 2 * 2^(c-1) additions ONCE instead of per window).  MEASURED on B200 (profiles/r01e_table_sweep.jsonl, device time
// of a registered MSM, plain -> table): 2^12 0.64 -> 0.29 ms (c = 8), 2^16 1.11 -> 0.63 (16), 2^20 4.13 -> 3.44 (17),
// 2^22 13.6 -> 11.3 (20), 2^24 46.6 -> 44.4 (20).  Only windows whose TOP digit keeps >= 6 bits are candidates
// (8, 13, 15, 16, 17, 19, 20, 22, 24): a 1-2 bit top window means a handful of buckets with n/4 points each.
int table_window_bits(size_t n) {
    int lg = 0;
    while (lg < 63 && (1ull << (lg + 1)) <= n) lg++;
    if ((1ull << lg) < n && n - (1ull << lg) > (1ull << lg) / 2) lg++;
    return lg <= 13 ? 8 : lg <= 19 ? 16 : lg <= 21 ? 17 : 20;
}
// Batched-affine accumulation pays when every resident thread gets several 256-entry chunks (measured: DESIGN.md).
bool ba_auto(uint64_t max_entries, int sm_count) {
    (void)max_entries; (void)sm_count;
    return false;
}
int auto_window_bits(size_t n, int sm_count) {
    bool g;
    int c;
    auto_policy(n, sm_count, true, &g, &c);
    return c;
}

}  // namespace

struct b200msm_bases {
    struct Shard {
        int dev_index;
        size_t begin, len;
        void* d_xy = nullptr;   // len x 64 B; in table mode the [tW][len] table, whose window 0 is the bases themselves
        void* d_inf = nullptr;  // nullptr when the set has no infinity flags
        int tc = 0, tW = 0;     // table mode: window size and window count the table was built for (0 = plain bases)
    };
    std::vector<Shard> shards;
    size_t n = 0;
};

struct b200msm_ctx {
    std::vector<DevState> devs;
    std::mutex mu;
    int opt_window_bits = 0;
    int opt_chunk = 0;
    int opt_timing = 0;
    int opt_reduce_log2 = -1;
    int opt_groups = 0;
    int opt_glv = -1;
    int opt_coop_reduce = -1;
    int opt_rowcol = -1;
    int opt_slices = 0;
    int opt_ranked_sort = -1;
    int opt_fix_chunks = -1;
    int opt_precompute = 0;
    int opt_adaptive_slices = -1;  // host-buffer call: equal slices when the last call was transfer-bound (-1 / 1 on, 0 off)
    int opt_sort_overlap = -1;  // sliced host pipeline: sort slice k+1 on the sort stream under slice k's accumulation (-1 auto: by size)
    int opt_slice_ratio = 0;    // percent: length of slice k+1 / length of slice k; 0 = auto (by slice count)
    int opt_batch_affine = -1;  // -1 auto, 0 XYZZ chunks (k_accumulate), 1 batched affine (k_accumulate_ba)
    int opt_ba_chunk = 0;       // 0 auto; else entries per batched-affine thread (64..512)
    int opt_ba_min_pairs = 0;   // 0 auto; else the smallest round worth an inversion
    b200msm_timings last = {};
    CopyPool* pool = nullptr;     // host copy threads that stage pageable input (shared by all devices of the context)
    uint8_t* h_pinned = nullptr;  // result / partial staging
    size_t h_pinned_bytes = 0;
    Plan last_plan;
    int pending_timings_dev = -1;
};

namespace {

// force_c > 0: take (window size, GLV) from the caller (the slices of one MSM share the whole MSM's shape so
// that their bucket arrays can be merged) instead of the policy / options.
int make_plan(const b200msm_ctx* ctx, const DevState& d, size_t n, Plan* out, int force_c = 0, bool force_glv = false,
              size_t table_stride = 0) {
    Plan p;
    if (n == 0) return fail(B200MSM_EINVAL, "Empty input");
    if (n >= (1ull << 31)) return fail(B200MSM_EINVAL, "n must be < 2^31 per device");
    p.n = (uint32_t)n;
    // "glv": -1 auto (measured policy), 0 off, 1 forced on
    bool glv_auto = false;
    int c_auto = 16;
    auto_policy(n, d.sm_count, ctx->opt_glv != 0 && n < (1ull << 30), &glv_auto, &c_auto);
    p.glv = ctx->opt_glv < 0 ? glv_auto : (ctx->opt_glv != 0 && n < (1ull << 30));
    if (ctx->opt_glv > 0 && !glv_auto) {  // forced on outside the auto range: nearest admissible window
        c_auto = 16;
    }
    if (ctx->opt_glv == 0 && glv_auto) {  // forced off: plain-window table
        bool g2;
        auto_policy(n, d.sm_count, false, &g2, &c_auto);
    }
    if (force_c > 0) p.glv = force_glv;
    p.n_eff = p.glv ? 2 * p.n : p.n;
    p.c = force_c > 0 ? force_c : ctx->opt_window_bits ? ctx->opt_window_bits : c_auto;
    if (p.c < 4 || p.c > 24) return fail(B200MSM_EINVAL, "window_bits must be in [4, 24]");
    p.W = num_windows_for(p.c, p.glv ? 127 : 254);
    if ((uint64_t)p.W * p.n_eff >= (1ull << 32)) return fail(B200MSM_EINVAL, "num_windows * n must be < 2^32 per device");
    p.tstride = (uint32_t)table_stride;
    p.Wb = table_stride ? 1 : p.W;
    if (table_stride && (uint64_t)p.W * table_stride >= (1ull << 31)) return fail(B200MSM_EINVAL, "table too large for 31-bit entries");
    p.half = 1u << (p.c - 1);
    p.nb = p.half + 1;
    p.G = (uint32_t)p.Wb * p.nb;
    p.wide_digits = p.c > 16;
    p.ranked = ctx->opt_ranked_sort != 0;
    uint64_t max_entries = (uint64_t)p.W * p.n_eff;
    // Sort engine.  "ranked_sort": -1 auto, 0 cursor atomics, 1 ranked, 2 partitioned.  The partitioned sort needs enough
    // digits to fill the machine (measured, profiles/r02t_psort_stage_times.jsonl: K1+K2 at 2^16 points 0.076 vs 0.064 ms
    // ranked, 2^18 0.109 vs 0.119, 2^20 0.249 vs 0.326, 2^22 0.93 vs 1.23, 2^24 2.57 vs 4.31) and a partition grid that
    // fits its shared-memory counters.
    {
        // buckets per partition: ~16K digits each (what k_place sorts inside shared memory), at most PSORT_MAX_SLOTS buckets,
        // at most PSORT_MAX_NPW partitions per window and PSORT_MAX_NP in total
        auto shift_for = [&](uint64_t mags) {
            uint32_t sh = 0;
            while (sh < 12 && ((uint64_t)p.n_eff << sh) < 16384ull * mags) sh++;
            return sh;
        };
        auto parts_of = [](uint64_t mags, uint32_t sh) { return (uint32_t)((mags + (1ull << sh) - 1) >> sh); };
        // the top window only reaches magnitudes <= 2^(bits - c (W - 1)) (its digit is narrower; + 1 for the carry is covered
        // because the range is a power of two and the recoding wraps at half)
        const int rem_bits = (p.glv ? 127 : 254) - p.c * (p.W - 1);
        const uint64_t top_mags = table_stride ? p.half : (rem_bits <= 0 ? 1ull : std::min<uint64_t>(p.half, 1ull << rem_bits));
        uint32_t shift = shift_for(p.half), shift_top = shift_for(top_mags);
        auto total_parts = [&]() {
            return table_stride ? (uint64_t)parts_of(p.half, shift) : (uint64_t)(p.W - 1) * parts_of(p.half, shift) + parts_of(top_mags, shift_top);
        };
        while (shift < 12 && (parts_of(p.half, shift) > PSORT_MAX_NPW || total_parts() > PSORT_MAX_NP)) shift++;
        while (shift_top < 12 && parts_of(top_mags, shift_top) > PSORT_MAX_NPW) shift_top++;
        const bool fits = parts_of(p.half, shift) <= PSORT_MAX_NPW && parts_of(top_mags, shift_top) <= PSORT_MAX_NPW &&
                          total_parts() <= PSORT_MAX_NP;                                        // hard limits (shared-memory counters)
        const bool dense = ((uint64_t)p.n_eff << shift) >= 2048ull * p.half;                    // >= 2K digits per partition
        p.psort = fits && (ctx->opt_ranked_sort == 2 || (ctx->opt_ranked_sort < 0 && dense && max_entries >= (1ull << 22)));
        if (ctx->opt_ranked_sort == 2 && !fits) return fail(B200MSM_EINVAL, "ranked_sort = 2: the partitioned sort does not fit this window size");
        p.ps.shift = shift;
        p.ps.npw = parts_of(p.half, shift);
        p.ps.shift_top = table_stride ? shift : shift_top;
        p.ps.npw_top = table_stride ? p.ps.npw : parts_of(top_mags, shift_top);
        p.ps.shared_set = table_stride ? 1u : 0u;
        p.ps.top = table_stride ? 0xffffffffu : (uint32_t)(p.W - 1);
        p.ps.np = (uint32_t)total_parts();
        p.ps.heavy = (uint32_t)std::min<uint64_t>(0xffffffffull, std::max<uint64_t>(PSORT_HEAVY, 4 * (max_entries / std::max<uint64_t>(1, total_parts()))));
        const uint32_t per_point = (uint32_t)p.W * (p.glv ? 2u : 1u);
        uint32_t tile = 256;
        while (tile < 4096 && (uint64_t)tile * per_point < 8ull * p.ps.np) tile *= 2;
        while (tile > 256 && (uint64_t)p.n / tile < 2ull * (uint64_t)d.sm_count) tile /= 2;
        p.psort_tile_pts = tile;
    }
    uint32_t L = 64;
    if (ctx->opt_chunk > 0) {
        L = (uint32_t)ctx->opt_chunk;
    } else {
        uint64_t want = max_entries / ((uint64_t)d.sm_count * 1024);
        L = 8;
        while (L * 2 <= want && L < 64) L *= 2;
        // Crowded buckets (the single bucket set of a big table-mode MSM: hundreds of entries each) straddle many
        // 64-entry chunks and the per-bucket fix-up becomes a long serial chain; longer chunks remove it as long as
        // >= 10 waves of chunks remain (measured, 2^24 with the window table: L = 64 -> 256 takes 44.4 -> 42.0 ms).
        const uint64_t occupancy = max_entries / ((uint64_t)p.Wb * p.half);
        const uint64_t resident_threads = (uint64_t)d.sm_count * 512;
        while (L >= 64 && L < 256 && occupancy >= 2 * L && max_entries / (2 * L) >= 10 * resident_threads) L *= 2;
    }
    // Batched-affine accumulation: "batch_affine" -1 auto, 0 off, 1 on.  Its chunks are long (a lane needs >= 64
    // independent additions per inversion) and live in a global scratch slice per resident thread.
    p.ba = ctx->opt_batch_affine > 0 || (ctx->opt_batch_affine < 0 && ba_auto(max_entries, d.sm_count));
    if (p.ba) {
        L = ctx->opt_ba_chunk > 0 ? (uint32_t)ctx->opt_ba_chunk : 256;
        p.ba_min_pairs = ctx->opt_ba_min_pairs > 0 ? (uint32_t)ctx->opt_ba_min_pairs : 24;
    }
    p.L = L;
    p.nchunks = (uint32_t)((max_entries + L - 1) / L);
    // thread-per-chunk fix-up; "fix_chunks" -1 auto, 0, 1.  Measured (profiles/r02z_fixup_chunks*.jsonl, whole MSM, per bucket ->
    // per chunk): 2^12 0.613 -> 0.621 ms, 2^14 0.725 -> 0.728, 2^16 1.018 -> 0.998, 2^18 1.576 -> 1.526, 2^20 3.677 -> 3.658,
    // 2^22 12.15 -> 12.09, 2^24 41.90 -> 41.25 (its three launches only pay from 2^20 digits)
    // ... or when the average bucket spans >= 4 chunks of a small input (2^12 points at c = 8: 63 entries per bucket over 8-entry
    // chunks): the queued buckets then get a warp each and a log-depth tree (k_fixup_medium), 2^12 0.615 -> 0.557 ms
    // (profiles/r02f_fixup_warp.jsonl)
    p.fix_chunks = !p.ba && (ctx->opt_fix_chunks > 0 ||
                             (ctx->opt_fix_chunks < 0 && (max_entries >= (1ull << 20) || max_entries >= 4ull * L * (uint64_t)p.Wb * p.half)));
    // bucket-reduce shape: Bsz = 2^log2Bsz magnitudes per thread, bpw CTAs of 128 threads per window (<= 128,
    // the widest k_window_finish); short per-thread chains matter because every EC addition of a lone warp
    // costs ~7 us
    uint32_t lb = ctx->opt_reduce_log2 >= 0 ? (uint32_t)ctx->opt_reduce_log2 : 4;
    while ((((uint64_t)p.half + ((uint64_t)RED_THREADS << lb) - 1) / ((uint64_t)RED_THREADS << lb)) > 128) lb++;
    p.log2Bsz = lb;
    p.bpw = (uint32_t)(((uint64_t)p.half + ((uint64_t)RED_THREADS << lb) - 1) / ((uint64_t)RED_THREADS << lb));
    // Cooperative reduce levels.  Level 0 chains own 2^lb0 buckets: 16 when the reduce is latency-bound, more
    // when there are so many buckets that it is throughput-bound (the per-CTA combine is ~20 additions).
    // "coop_reduce": -1 auto (cooperative engine while the reduce is latency-bound: <= 2^20 buckets in total; the
    // thread-per-segment kernels in the throughput regime, measured 6.6 vs 7.6 ms at 2^24 / c = 20), 0 / 1 forced
    p.coop_reduce = ctx->opt_coop_reduce < 0 ? ((uint64_t)p.Wb * p.half <= (1ull << 20)) : ctx->opt_coop_reduce != 0;
    {
        const uint64_t total_buckets = (uint64_t)p.Wb * p.half;
        uint32_t lb0 = total_buckets <= (1ull << 19) ? 4 : total_buckets <= (1ull << 21) ? 5 : 6;
        while (lb0 > 0 && ((uint64_t)32 << (lb0 - 1)) >= p.half) lb0--;   // one CTA already covers the window
        if (ctx->opt_reduce_log2 >= 0) lb0 = (uint32_t)ctx->opt_reduce_log2;
        uint64_t cnt = p.half;
        uint32_t lbl = lb0;
        p.red_nl = 0;
        p.red_slots = 0;
        while (p.red_nl < 8) {
            uint32_t ctas = (uint32_t)((cnt + ((uint64_t)32 << lbl) - 1) / ((uint64_t)32 << lbl));
            p.red_lb[p.red_nl] = lbl;
            p.red_ctas[p.red_nl] = ctas;
            p.red_slots += 2 * (size_t)p.W * ctas;
            p.red_nl++;
            if (ctas == 1) break;
            cnt = ctas;
            lbl = 0;
            while (((uint64_t)32 << lbl) < cnt && lbl < 4) lbl++;
        }
    }
    // Row / column sums: "rowcol_reduce" 1 = on, -1 / 0 = off.  MEASURED NEUTRAL on B200 (profiles/r02_rowcol_reduce.jsonl,
    // r02_rowcol_ncu.csv; 2^20, c = 16: k_rowcol_sums 0.281 + k_reduce_rowcol 0.100 + 0.05 of extra Horner additions against
    // 0.327 + 0.085 for the two cooperative levels; whole MSM 2^14 0.752 -> 0.733 ms, 2^16 1.000 -> 0.994, 2^20 3.650 -> 3.672):
    // the 2 plain additions per bucket are throughput work (7.3 M products) that 3 resident warps per sub-partition run at
    // ~60 % of the pipe, which cancels the shorter dependency chain.  Kept as an option, not the default.
    // R = 2^ra rows x C = 2^cb columns, cb >= ra (a column sum has R terms, a row sum C).
    p.rc_ra = (uint32_t)(p.c - 1) / 2;
    p.rc_cb = (uint32_t)(p.c - 1) - p.rc_ra;
    p.rowcol = p.c >= 3 && p.rc_cb <= 10 && ctx->opt_rowcol > 0;
    if (p.rowcol) p.red_slots = std::max(p.red_slots, (size_t)p.W * ((size_t)(1u << p.rc_ra) + (1u << p.rc_cb)) + 3 * (size_t)p.W + 2);
    // Window groups (accumulate of group k+1 on the main stream overlapping the reduce chain of group k on the
    // side stream).  MEASURED NEGATIVE on B200 (profiles/r01_groups_experiment.jsonl: 2^20 4.57 -> 5.98 ms with 4
    // groups): the reduce chain is latency-bound per group, so splitting multiplies it.  Default: one group.
    int ng = ctx->opt_groups > 0 ? ctx->opt_groups : 1;
    p.ngroups = std::max(1, std::min({ng, p.Wb, 8}));
    *out = p;
    return B200MSM_OK;
}

// upper bound of the (partition, slice) work items of heavy partitions: every item but the last of a partition is a full slice
inline size_t psort_heavy_items_cap(uint64_t max_entries, uint32_t np) {
    return (size_t)(max_entries / PSORT_SLICE + std::min<uint64_t>(np, max_entries / PSORT_HEAVY + 1) + 2);
}

// queue of giant buckets + one ticket counter per queued bucket, padded to 128 B
inline size_t giant_list_bytes(uint32_t nchunks) { return (((size_t)nchunks / FIX_GIANT + 2) * 8 + 127) & ~(size_t)127; }

// Scratch of one sort + accumulate + fix-up pass (slice k of a sliced MSM, or the whole MSM for k = 0).
int ensure_work(DevState& d, const Plan& p, int k = 0, bool shared_buckets = false) {
    Buf *ranks = k > 0 ? &d.extra[k - 1].ranks : &d.ranks;
    if (p.ranked || p.psort) RET_TRY(ranks->ensure((size_t)p.W * p.n_eff * 4));   // psort: the staged entries
    if (p.psort) {
        Buf *skeys = k > 0 ? &d.extra[k - 1].skeys : &d.skeys, *parts = k > 0 ? &d.extra[k - 1].parts : &d.parts;
        RET_TRY(skeys->ensure((size_t)p.W * p.n_eff * 2));
        // part_count | part_base (+1) | part_cursor | heavy_total | heavy (partition, slice) items
        RET_TRY(parts->ensure(((size_t)3 * p.ps.np + 8 + (PSORT_MAX_W + 2) + 2 * psort_heavy_items_cap((uint64_t)p.W * p.n_eff, p.ps.np)) * 4));
        RET_TRY((k > 0 ? &d.extra[k - 1].bslots : &d.bslots)->ensure((size_t)2 * p.G * 4));   // per-bucket counts and cursors (heavy partitions)
    }
    Buf *digits = &d.digits, *ends = &d.ends, *wtotal = &d.wtotal, *entries = &d.entries, *buckets = &d.buckets,
        *head = &d.head, *tail = &d.tail, *longlist = &d.longlist;
    if (k > 0) {
        auto& e = d.extra[k - 1];
        digits = &e.digits; ends = &e.ends; wtotal = &e.wtotal; entries = &e.entries; buckets = &e.buckets;
        head = &e.head; tail = &e.tail; longlist = &e.longlist;
    }
    RET_TRY(digits->ensure((size_t)p.W * p.n_eff * (p.wide_digits ? 4 : 2)));
    RET_TRY(ends->ensure((size_t)p.G * 4));
    RET_TRY(wtotal->ensure(128 * 4));
    RET_TRY(longlist->ensure(((size_t)p.nchunks / FIX_LONG + 2) * 4 * 8));
    RET_TRY(entries->ensure((size_t)p.W * p.n_eff * 4));
    if (!shared_buckets) RET_TRY(buckets->ensure((size_t)p.G * sizeof(xyzz_t)));   // shared: the slice adds into slice 0's array
    RET_TRY(head->ensure((size_t)p.nchunks * sizeof(xyzz_t)));
    RET_TRY(tail->ensure((size_t)p.nchunks * sizeof(xyzz_t)));
    // giant buckets (> FIX_GIANT chunks): queue (padded to 128 B) | segment partial sums
    RET_TRY((k > 0 ? &d.extra[k - 1].giant : &d.giant)->ensure(giant_list_bytes(p.nchunks) + ((size_t)p.nchunks / FIX_SEG + (size_t)p.nchunks / FIX_GIANT + 4) * sizeof(xyzz_t)));
    if (p.fix_chunks)   // chunk_g[nchunks + 2] | medium-bucket queue of every window group
        RET_TRY((k > 0 ? &d.extra[k - 1].chunkg : &d.chunkg)->ensure((((size_t)p.nchunks + 2) + (size_t)p.ngroups * ((size_t)p.nchunks / 2 + 2)) * 4));
    return B200MSM_OK;
}
// Per-device buffers of the bucket reduce and the result (shape depends on (c, W) only).
int ensure_reduce(DevState& d, const Plan& p) {
    RET_TRY(d.wpart.ensure(((size_t)p.W * p.bpw * 2 + p.W + 1) * sizeof(xyzz_t)));
    RET_TRY(d.redbuf.ensure((p.red_slots + 2) * sizeof(xyzz_t)));
    RET_TRY(d.out.ensure(sizeof(jac_t)));
    return B200MSM_OK;
}
int ensure_workspace(DevState& d, const Plan& p) {
    RET_TRY(ensure_work(d, p, 0));
    if (p.glv) RET_TRY(d.xb.ensure((size_t)p.n * 32));
    return ensure_reduce(d, p);
}

inline unsigned cdiv(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }
void shard_ranges(size_t n, size_t parts, std::vector<std::pair<size_t, size_t>>* out);

// K1 + K2 on stream s: digits, histogram, scan, scatter.  Afterwards w.ends holds bucket end offsets.
int launch_sort(const WorkView& w, const Plan& p, const void* d_scalars, const void* d_inf, cudaStream_t s, cudaEvent_t after_decompose) {
    const uint4* sc = (const uint4*)d_scalars;
    const uint8_t* inf = (const uint8_t*)d_inf;
    if (p.psort) {
        uint32_t* part_count = (uint32_t*)w.parts;
        uint32_t* part_base = part_count + p.ps.np;
        uint32_t* part_cursor = part_base + p.ps.np + 1;
        uint32_t* heavy = part_cursor + p.ps.np;         // [0] work-item count, [1 + w] window w has a heavy partition
        uint2* heavy_items = (uint2*)(part_count + (((size_t)3 * p.ps.np + 1 + PSORT_MAX_W + 2 + 1) & ~(size_t)1));   // 8-byte aligned
        uint32_t* bcount = (uint32_t*)w.bslots;
        uint32_t* bcursor = bcount + p.G;
        const size_t heavy_cap = psort_heavy_items_cap((uint64_t)p.W * p.n_eff, p.ps.np);
        CU_TRY(cudaMemsetAsync(part_count, 0, (size_t)p.ps.np * 4, s));
        CU_TRY(cudaMemsetAsync(bcount, 0, (size_t)2 * p.G * 4, s));
        const unsigned g1 = cdiv(p.n, p.psort_tile_pts);
        const size_t sm1 = (size_t)p.ps.np * 4;
#define B200_DECOUNT(DT, GLVF)                                                                                                  \
    do {                                                                                                                        \
        if (sm1 > 48 * 1024) CU_TRY(cudaFuncSetAttribute(k_decompose_count<DT, GLVF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1)); \
        k_decompose_count<DT, GLVF><<<g1, 256, sm1, s>>>(sc, inf, p.n, p.c, p.W, p.psort_tile_pts, p.ps, (DT*)w.digits, part_count);          \
    } while (0)
        if (p.wide_digits) { if (p.glv) B200_DECOUNT(int32_t, true); else B200_DECOUNT(int32_t, false); }
        else               { if (p.glv) B200_DECOUNT(int16_t, true); else B200_DECOUNT(int16_t, false); }
#undef B200_DECOUNT
        if (after_decompose) CU_TRY(cudaEventRecord(after_decompose, s));
        k_pscan<<<1, 1024, 0, s>>>(part_count, p.ps.np, p.ps.npw, p.ps.top, p.ps.heavy, part_base, part_cursor, heavy, heavy_items);
        const uint32_t tiles = cdiv(p.n_eff, PSORT_TILE);
        const uint32_t npw_max = std::max(p.ps.npw, p.ps.npw_top);
        const size_t sm2 = (size_t)8 * ((npw_max + PART_THREADS - 1) & ~(uint32_t)(PART_THREADS - 1)) + (size_t)PSORT_TILE * 8;
        if (p.wide_digits) {
            if (sm2 > 48 * 1024) CU_TRY(cudaFuncSetAttribute(k_partition<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
            k_partition<int32_t><<<(unsigned)p.W * tiles, PART_THREADS, sm2, s>>>((const int32_t*)w.digits, p.n_eff, tiles, p.tstride, p.ps, part_cursor,
                                                                      (uint32_t*)w.ranks, (uint16_t*)w.skeys, part_base, heavy, p.tstride ? 0u : p.nb, bcount);
        } else {
            if (sm2 > 48 * 1024) CU_TRY(cudaFuncSetAttribute(k_partition<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
            k_partition<int16_t><<<(unsigned)p.W * tiles, PART_THREADS, sm2, s>>>((const int16_t*)w.digits, p.n_eff, tiles, p.tstride, p.ps, part_cursor,
                                                                      (uint32_t*)w.ranks, (uint16_t*)w.skeys, part_base, heavy, p.tstride ? 0u : p.nb, bcount);
        }
        const size_t sm3 = (size_t)(PSORT_MAX_SLOTS + PLACE_CAP) * 4;
        CU_TRY(cudaFuncSetAttribute(k_place, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
        k_place<<<p.ps.np, PLACE_THREADS, sm3, s>>>((const uint32_t*)w.ranks, (const uint16_t*)w.skeys, part_base, p.ps, p.half,
                                                   p.tstride ? 0u : p.nb, (uint32_t*)w.ends, (uint32_t*)w.entries);
        // heavy partitions (skewed scalars): a persistent grid walks k_pscan's (partition, slice) list; no list, no work
        (void)heavy_cap;
        k_place_heavy<<<2 * 148, PLACE_THREADS, (size_t)2 * PSORT_MAX_SLOTS * 4, s>>>((const uint32_t*)w.ranks, (const uint16_t*)w.skeys, part_base, p.ps,
                                                                                     p.half, p.tstride ? 0u : p.nb, (uint32_t*)w.ends,
                                                                                     (uint32_t*)w.entries, heavy, heavy_items, bcount, bcursor);
        CU_TRY(cudaGetLastError());
        return B200MSM_OK;
    }
    CU_TRY(cudaMemsetAsync(w.ends, 0, (size_t)p.G * 4, s));
    uint32_t* hist = (uint32_t*)w.ends;
    const unsigned g1 = cdiv(p.n, 256);
    const uint32_t wstride = p.tstride ? 0 : p.nb;
    uint32_t* rk = (uint32_t*)w.ranks;
#define B200_DECOMPOSE(DT, GLVF)                                                                                              \
    do {                                                                                                                      \
        if (p.ranked) k_decompose<DT, GLVF, true><<<g1, 256, 0, s>>>(sc, inf, p.n, p.c, p.W, wstride, (DT*)w.digits, hist, rk); \
        else k_decompose<DT, GLVF, false><<<g1, 256, 0, s>>>(sc, inf, p.n, p.c, p.W, wstride, (DT*)w.digits, hist, nullptr);   \
    } while (0)
    if (p.wide_digits) {
        if (p.glv) B200_DECOMPOSE(int32_t, true); else B200_DECOMPOSE(int32_t, false);
    } else {
        if (p.glv) B200_DECOMPOSE(int16_t, true); else B200_DECOMPOSE(int16_t, false);
    }
#undef B200_DECOMPOSE
    if (after_decompose) CU_TRY(cudaEventRecord(after_decompose, s));
    // flat two-level scan over all G counters: up to 64 segments of >= 4096 counters, whatever the window structure
    const uint32_t nseg = std::max(1u, std::min(64u, p.G / 4096));
    const uint32_t seg = (p.G + nseg - 1) / nseg;
    k_scan_windows<<<nseg, 1024, 0, s>>>(hist, seg, p.G, (uint32_t*)w.wtotal, p.ranked ? 1 : 0);
    k_add_window_base<<<cdiv(p.G, 256), 256, 0, s>>>(hist, seg, p.G, (const uint32_t*)w.wtotal);
    const unsigned g2 = cdiv((uint64_t)p.W * p.n_eff, 256);
    if (p.ranked) {
        if (p.wide_digits)
            k_scatter_ranked<int32_t><<<g2, 256, 0, s>>>((const int32_t*)w.digits, rk, p.n_eff, p.W, wstride, p.tstride, hist, (uint32_t*)w.entries);
        else
            k_scatter_ranked<int16_t><<<g2, 256, 0, s>>>((const int16_t*)w.digits, rk, p.n_eff, p.W, wstride, p.tstride, hist, (uint32_t*)w.entries);
    } else if (p.wide_digits) {
        k_scatter<int32_t><<<g2, 256, 0, s>>>((const int32_t*)w.digits, p.n_eff, p.W, wstride, p.tstride, hist, (uint32_t*)w.entries);
    } else {
        k_scatter<int16_t><<<g2, 256, 0, s>>>((const int16_t*)w.digits, p.n_eff, p.W, wstride, p.tstride, hist, (uint32_t*)w.entries);
    }
    CU_TRY(cudaGetLastError());
    return B200MSM_OK;
}

// K3 for windows [w_lo, w_hi): chunked accumulation on stream s (caller zeroed the long-bucket counters).
int launch_accumulate(DevState& d, const WorkView& w, const Plan& p, const void* d_bases, const fq* d_xb, int w_lo, int w_hi, cudaStream_t s,
                      bool into = false) {
    const uint32_t g_lo = (uint32_t)w_lo * p.nb, g_hi = (uint32_t)w_hi * p.nb;
    const uint64_t max_chunks = ((uint64_t)(p.tstride ? p.W : w_hi - w_lo) * p.n_eff + p.L - 1) / p.L + 2;
    if (p.ba) {
        if (!d.ba_ctas_per_sm) {
            int nb = 0;
            CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_accumulate_ba, BA_THREADS, 0));
            d.ba_ctas_per_sm = std::max(1, nb);
        }
        const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)d.sm_count * d.ba_ctas_per_sm, cdiv(max_chunks, BA_THREADS));
        RET_TRY(d.ba_scratch.ensure((size_t)grid * BA_THREADS * ba_scratch_bytes_per_thread(p.L)));
        k_accumulate_ba<<<grid, BA_THREADS, 0, s>>>((const affine_t*)d_bases, d_xb, p.tstride ? 0xffffffffu : p.n,
                                                    (const uint32_t*)w.entries, (const uint32_t*)w.ends, g_lo, g_hi, p.L, p.ba_min_pairs,
                                                    (xyzz_t*)w.buckets, (xyzz_t*)w.head, (xyzz_t*)w.tail, (uint8_t*)d.ba_scratch.p);
        CU_TRY(cudaGetLastError());
        return B200MSM_OK;
    }
    // table mode: entries index the [W][tstride] table directly (never the endomorphism branch)
    k_accumulate<<<cdiv(max_chunks, ACC_THREADS), ACC_THREADS, 0, s>>>((const affine_t*)d_bases, d_xb, p.tstride ? 0xffffffffu : p.n,
                                                                       (const uint32_t*)w.entries,
                                                                       (const uint32_t*)w.ends, g_lo, g_hi, p.L, (xyzz_t*)w.buckets,
                                                                       (xyzz_t*)w.head, (xyzz_t*)w.tail,
                                                                       p.fix_chunks ? (uint32_t*)w.chunkg : nullptr, into ? 1 : 0,
                                                                       w_hi == p.Wb ? 1 : 0);
    CU_TRY(cudaGetLastError());
    return B200MSM_OK;
}

// Chunk-boundary fix-up for windows [w_lo, w_hi) on stream r; k selects the long-bucket list slot.
int launch_fixup(const DevState& d, const WorkView& w, const Plan& p, int w_lo, int w_hi, int k, cudaStream_t r, bool into = false) {
    const uint32_t g_lo = (uint32_t)w_lo * p.nb, g_hi = (uint32_t)w_hi * p.nb;
    uint32_t* long_count = (uint32_t*)w.wtotal + 64;
    const uint32_t long_cap = p.nchunks / FIX_LONG + 2;
    uint32_t* giant_count = long_count + 24 + k;                                        // zeroed together with the long counters
    uint32_t* giant_list = (uint32_t*)w.giant;
    uint32_t* giant_done = giant_list + ((size_t)p.nchunks / FIX_GIANT + 2);            // per-bucket ticket counters
    CU_TRY(cudaMemsetAsync(giant_done, 0, ((size_t)p.nchunks / FIX_GIANT + 2) * 4, r));
    xyzz_t* gpart = (xyzz_t*)((uint8_t*)w.giant + giant_list_bytes(p.nchunks));
    if (p.fix_chunks) {
        const uint64_t max_chunks = ((uint64_t)(p.tstride ? p.W : w_hi - w_lo) * p.n_eff + p.L - 1) / p.L + 2;
        if (!into) k_fixup_empty<<<cdiv(g_hi - g_lo, 256), 256, 0, r>>>((const uint32_t*)w.ends, g_lo, g_hi, (xyzz_t*)w.buckets);
        uint32_t* medium_count = long_count + 8 + k;                                   // zeroed together with the long counters
        const size_t medium_cap = (size_t)p.nchunks / 2 + 2;
        uint32_t* medium_list = (uint32_t*)w.chunkg + ((size_t)p.nchunks + 2) + (size_t)k * medium_cap;
        k_fixup_chunks<<<cdiv(max_chunks, 128), 128, 0, r>>>((const uint32_t*)w.ends, g_lo, g_hi, p.L, (xyzz_t*)w.buckets,
                                                             (const xyzz_t*)w.head, (const xyzz_t*)w.tail, (const uint32_t*)w.chunkg,
                                                             long_count + k, (uint32_t*)w.longlist + (size_t)k * long_cap, medium_count,
                                                             medium_list, giant_count, giant_list);
        k_fixup_medium<<<d.sm_count * 8, 128, 0, r>>>((const uint32_t*)w.ends, p.L, (xyzz_t*)w.buckets, (const xyzz_t*)w.head,
                                                      (const xyzz_t*)w.tail, medium_count, medium_list);
    } else
    k_fixup<<<cdiv(g_hi - g_lo, 128), 128, 0, r>>>((const uint32_t*)w.ends, g_lo, g_hi, p.L, (xyzz_t*)w.buckets, (const xyzz_t*)w.head,
                                                   (const xyzz_t*)w.tail, long_count + k, (uint32_t*)w.longlist + (size_t)k * long_cap, into ? 1 : 0, giant_count,
                                                   giant_list);
    k_fixup_long<<<d.sm_count * 2, FIXL_THREADS, 0, r>>>((const uint32_t*)w.ends, p.L, (xyzz_t*)w.buckets, (const xyzz_t*)w.head,
                                                         (const xyzz_t*)w.tail, long_count + k,
                                                         (const uint32_t*)w.longlist + (size_t)k * long_cap);
    k_fixup_giant<<<d.sm_count * 2, FIXL_THREADS, 0, r>>>((const uint32_t*)w.ends, p.L, (xyzz_t*)w.buckets, (const xyzz_t*)w.head,
                                                          (const xyzz_t*)w.tail, giant_count, giant_list, gpart, giant_done);
    CU_TRY(cudaGetLastError());
    return B200MSM_OK;
}

// K4 + K5 for windows [w_lo, w_hi) of `buckets` on stream r: window sums, then the Horner segment.
int launch_reduce(DevState& d, const Plan& p, const void* buckets, int w_lo, int w_hi, bool first, cudaStream_t r, void* d_out,
                  int* nlaunch) {
    xyzz_t* wpartR = (xyzz_t*)d.wpart.p;
    xyzz_t* wpartT = wpartR + (size_t)p.W * p.bpw;
    xyzz_t* wsum = wpartT + (size_t)p.W * p.bpw;
    uint32_t* hstate = (uint32_t*)(wsum + p.W);
    const xyzz_t* wsum2 = nullptr;
    if (p.rowcol) {
        const uint32_t R = 1u << p.rc_ra, C = 1u << p.rc_cb;
        xyzz_t* rc = (xyzz_t*)d.redbuf.p;
        xyzz_t* scratch = rc + (size_t)p.W * (R + C);
        xyzz_t* wcols = scratch + 2 * (size_t)p.W;
        const uint32_t problems = (uint32_t)(w_hi - w_lo) * (R + C);
        k_rowcol_sums<<<cdiv(problems, RC_WARPS), RC_WARPS * 32, 0, r>>>((const xyzz_t*)buckets, p.nb, p.rc_ra, p.rc_cb, (uint32_t)w_lo, problems, rc);
        k_reduce_rowcol<<<(w_hi - w_lo) * 2, CL_THREADS, 0, r>>>(rc, p.rc_ra, p.rc_cb, (uint32_t)w_lo, scratch, wsum, wcols);
        wsum2 = wcols;
        *nlaunch += 2;
    } else if (p.coop_reduce) {
        // recursive weighted sum on the lane-parallel cooperative engine
        const xyzz_t* Ain = (const xyzz_t*)buckets;
        const xyzz_t* Xin = nullptr;
        uint32_t in_stride = p.nb, in_off = 1, cnt = p.half, log2u = 0, delta = 0;
        xyzz_t* buf = (xyzz_t*)d.redbuf.p;
        for (int l = 0; l < p.red_nl; l++) {
            const uint32_t ctas = p.red_ctas[l];
            xyzz_t* Aout = buf;
            xyzz_t* Xout = ctas == 1 ? wsum : buf + (size_t)p.W * ctas;
            k_reduce_level<<<(w_hi - w_lo) * ctas, CL_THREADS, 0, r>>>(Ain, Xin, in_stride, in_off, cnt, p.red_lb[l], log2u, delta,
                                                                      ctas, (uint32_t)w_lo, Aout, Xout);
            Ain = Aout;
            Xin = Xout;
            in_stride = ctas;
            in_off = 0;
            cnt = ctas;
            log2u += 5 + p.red_lb[l];
            delta = 1;
            buf += 2 * (size_t)p.W * ctas;
            *nlaunch += 1;
        }
    } else {
        k_bucket_reduce<<<(w_hi - w_lo) * p.bpw, RED_THREADS, 0, r>>>((const xyzz_t*)buckets, p.nb, p.log2Bsz, p.bpw, (uint32_t)w_lo,
                                                                      wpartR, wpartT);
        if (p.bpw <= 32)
            k_window_finish<32><<<w_hi - w_lo, 32, 0, r>>>(wpartR, wpartT, p.bpw, p.log2Bsz + 7, (uint32_t)w_lo, wsum);
        else
            k_window_finish<128><<<w_hi - w_lo, 128, 0, r>>>(wpartR, wpartT, p.bpw, p.log2Bsz + 7, (uint32_t)w_lo, wsum);
        *nlaunch += 2;
    }
    k_window_combine<<<1, CMB_THREADS, 0, r>>>(wsum, wsum2, w_lo, w_hi, p.c, hstate, first, w_lo == 0, (jac_t*)d_out);
    *nlaunch += 1;
    CU_TRY(cudaGetLastError());
    return B200MSM_OK;
}

// Enqueue the whole single-device pipeline on d.stream.  No host synchronisation.
int enqueue_msm(b200msm_ctx* ctx, DevState& d, const Plan& p, const void* d_bases, const void* d_inf,
                const void* d_scalars, void* d_out, unsigned long long* launches, cudaEvent_t bases_ready = nullptr,
                const void* d_xb_pre = nullptr) {
    cudaStream_t s = d.stream;
    const bool timing = ctx->opt_timing != 0;
    const WorkView w = view_main(d);
    RET_TRY(launch_sort(w, p, d_scalars, d_inf, s, timing ? d.ev[EV_DECOMP] : nullptr));
    if (bases_ready) CU_TRY(cudaStreamWaitEvent(s, bases_ready, 0));  // bases were uploaded on the side stream meanwhile
    const fq* d_xb = (const fq*)d_xb_pre;
    if (p.glv && !d_xb) {   // x coordinates of phi(P_i) = (beta * x_i, y_i)
        k_endo_x<<<cdiv(p.n, 256), 256, 0, s>>>((const affine_t*)d_bases, p.n, (fq*)d.xb.p);
        d_xb = (const fq*)d.xb.p;
    }
    // sort_ms ends here (k_endo_x included), so that accumulate_ms brackets exactly the k_accumulate launches
    if (timing) CU_TRY(cudaEventRecord(d.ev[EV_SORT], s));
    // Window groups, top group first.  Accumulation of group k+1 runs on the main stream while the fix-up,
    // bucket reduce and the Horner segment of group k run on the high-priority side stream.
    cudaStream_t s2 = d.stream2;
    const int NG = p.ngroups;
    const int gw = (p.Wb + NG - 1) / NG;
    CU_TRY(cudaMemsetAsync((uint32_t*)w.wtotal + 64, 0, 4 * 48, s));
    int nlaunch = p.glv ? 5 : 4;   // decompose, scan, add-base, scatter (+ endo)
    for (int k = 0; k < NG; k++) {
        const int w_hi = p.Wb - k * gw;
        const int w_lo = std::max(0, w_hi - gw);
        if (w_hi <= 0) break;
        RET_TRY(launch_accumulate(d, w, p, d_bases, d_xb, w_lo, w_hi, s));
        cudaStream_t r = NG > 1 ? s2 : s;
        if (NG > 1) {
            CU_TRY(cudaEventRecord(d.ev_acc[k], s));
            CU_TRY(cudaStreamWaitEvent(s2, d.ev_acc[k], 0));
        }
        // the last group that actually runs (the loop ends early when gw * k >= Wb)
        if (timing && (k == NG - 1 || p.Wb - (k + 1) * gw <= 0)) CU_TRY(cudaEventRecord(d.ev[EV_ACC], s));
        RET_TRY(launch_fixup(d, w, p, w_lo, w_hi, k, r));
        nlaunch += 1 + (p.fix_chunks ? 4 : 2);   // accumulate + fix-up kernels
        RET_TRY(launch_reduce(d, p, w.buckets, w_lo, w_hi, k == 0, r, d_out, &nlaunch));
    }
    if (NG > 1) {
        CU_TRY(cudaEventRecord(d.ev_done, s2));
        CU_TRY(cudaStreamWaitEvent(s, d.ev_done, 0));
    }
    if (timing) CU_TRY(cudaEventRecord(d.ev[EV_RED], s));
    CU_TRY(cudaGetLastError());
    if (launches) *launches += nlaunch;
    return B200MSM_OK;
}

int collect_timings(b200msm_ctx* ctx, DevState& d, const Plan& p, bool had_h2d) {
    b200msm_timings t = {};
    t.window_bits = p.c;
    t.num_windows = p.W;
    if (ctx->opt_timing) {
        // a stage whose event was never recorded (option combinations that skip it) reads as 0 instead of failing the MSM
        auto span = [&](int a, int b) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, d.ev[a], d.ev[b]) != cudaSuccess) { cudaGetLastError(); ms = 0; }
            return ms;
        };
        if (had_h2d) t.h2d_ms = span(EV_START, EV_H2D);
        t.decompose_ms = span(EV_H2D, EV_DECOMP);
        t.sort_ms = span(EV_DECOMP, EV_SORT);
        t.accumulate_ms = span(EV_SORT, EV_ACC);
        t.reduce_ms = span(EV_ACC, EV_RED);
        t.total_ms = span(EV_H2D, EV_RED);
        uint32_t total = 0;
        CU_TRY(cudaMemcpy(&total, (const uint32_t*)d.ends.p + (p.G - 1), 4, cudaMemcpyDeviceToHost));
        t.entries = total;
    }
    t.kernel_launches = ctx->last.kernel_launches;
    ctx->last = t;
    return B200MSM_OK;
}

int check_layout(size_t base_stride, size_t x_off, size_t y_off, size_t inf_off) {
    if (base_stride % 8 || x_off % 8 || y_off % 8) return fail(B200MSM_EINVAL, "base stride/offsets must be multiples of 8");
    if (x_off + 32 > base_stride || y_off + 32 > base_stride) return fail(B200MSM_EINVAL, "x/y offset outside the record");
    if (inf_off != B200MSM_NO_INF && inf_off >= base_stride) return fail(B200MSM_EINVAL, "infinity offset outside the record");
    return B200MSM_OK;
}

// Ordinary malloc / Vec memory (not pinned, not managed)?
bool host_is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// Host -> device copy on stream st.  Pinned sources go straight to the DMA engine; pageable sources are staged by the
// library: the copy threads fill an 8 MB pinned ring slot while the DMA engine drains the previous ones (the host thread
// blocks only for the memcpy part, the tail of the transfer stays asynchronous).
int h2d(DevState& d, void* dst, const void* src, size_t bytes, cudaStream_t st, bool pageable) {
    if (!pageable || !d.pool) {
        CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return B200MSM_OK;
    }
    for (size_t off = 0; off < bytes; off += STAGE_BYTES) {
        const size_t len = std::min(STAGE_BYTES, bytes - off);
        const unsigned slot = d.stage_next++ % STAGE_SLOTS;
        if (!d.stage[slot]) {
            if (cudaHostAlloc((void**)&d.stage[slot], STAGE_BYTES, cudaHostAllocDefault) != cudaSuccess) {
                d.stage[slot] = nullptr;
                return fail(B200MSM_ENOMEM, "pinned staging allocation failed");
            }
            CU_TRY(cudaEventCreateWithFlags(&d.stage_ev[slot], cudaEventDisableTiming));
        }
        if (d.stage_used[slot]) CU_TRY(cudaEventSynchronize(d.stage_ev[slot]));
        d.pool->copy(d.stage[slot], (const uint8_t*)src + off, len);
        CU_TRY(cudaMemcpyAsync((uint8_t*)dst + off, d.stage[slot], len, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaEventRecord(d.stage_ev[slot], st));
        d.stage_used[slot] = true;
    }
    return B200MSM_OK;
}

// Upload + repack `len` base records starting at host pointer `src` into (d_xy, d_inf) on device d.
int upload_bases(DevState& d, const uint8_t* src, size_t stride, size_t x_off, size_t y_off, size_t inf_off, size_t len,
                 void* d_xy, void* d_inf, unsigned long long* launches, cudaStream_t st = nullptr, bool pageable = false) {
    if (!st) st = d.stream;
    if (stride == 64 && x_off == 0 && y_off == 32 && inf_off == B200MSM_NO_INF) return h2d(d, d_xy, src, len * 64, st, pageable);
    RET_TRY(d.raw.ensure(len * stride));
    RET_TRY(h2d(d, d.raw.p, src, len * stride, st, pageable));
    k_repack_bases<<<cdiv(len * 8, 256), 256, 0, st>>>((const uint8_t*)d.raw.p, stride, x_off, y_off, inf_off, (uint32_t)len,
                                                            (uint64_t*)d_xy, (uint8_t*)d_inf);
    CU_TRY(cudaGetLastError());
    if (launches) *launches += 1;
    return B200MSM_OK;
}

// Upload `len` scalar records from host `src` into the 32-byte records at d_dst, on stream st.  The staging buffer for
// strided records (d.scalars_raw) must already hold len * stride bytes.
int upload_scalars_to(DevState& d, const uint8_t* src, size_t stride, size_t len, void* d_dst, cudaStream_t st,
                      unsigned long long* launches, bool pageable = false) {
    if (stride == 32) {
        RET_TRY(h2d(d, d_dst, src, len * 32, st, pageable));
    } else {
        RET_TRY(h2d(d, d.scalars_raw.p, src, len * stride, st, pageable));
        k_repack_scalars<<<cdiv(len * 4, 256), 256, 0, st>>>((const uint8_t*)d.scalars_raw.p, stride, (uint32_t)len, (uint64_t*)d_dst);
        CU_TRY(cudaGetLastError());
        if (launches) *launches += 1;
    }
    return B200MSM_OK;
}
int upload_scalars(DevState& d, const uint8_t* src, size_t stride, size_t len, void** d_scalars, unsigned long long* launches,
                   bool pageable = false) {
    if (stride % 8 || stride < 32) return fail(B200MSM_EINVAL, "scalar stride must be a multiple of 8 and >= 32");
    RET_TRY(d.scalars.ensure(len * 32));
    if (stride != 32) RET_TRY(d.scalars_raw.ensure(len * stride));
    RET_TRY(upload_scalars_to(d, src, stride, len, d.scalars.p, d.stream, launches, pageable));
    *d_scalars = d.scalars.p;
    return B200MSM_OK;
}

// [0, n) cut into at most S contiguous, non-empty slices whose lengths grow by `ratio` from one to the next.
void slice_ranges(size_t n, int S, double ratio, std::vector<std::pair<size_t, size_t>>* out) {
    out->clear();
    double wsum = 0, wk = 1;
    for (int k = 0; k < S; k++, wk *= ratio) wsum += wk;
    size_t begin = 0;
    wk = 1;
    double acc = 0;
    for (int k = 0; k < S; k++, wk *= ratio) {
        acc += wk;
        size_t end = k == S - 1 ? n : std::min(n, (size_t)std::llround((double)n * acc / wsum));
        if (end > begin) out->push_back({begin, end - begin});
        begin = end;
    }
}

// Host-input MSM in S slices of the point range (S >= 2).  The copy stream uploads scalars and bases slice by slice;
// the main stream sorts, accumulates and fixes up slice k as soon as its data has landed, each slice into its own
// bucket array, so all but the first slice's transfer hides behind the arithmetic of the slices before it.  The
// bucket arrays are then merged (S-1 additions per bucket) and reduced once.  Resident-input callers
// (b200msm_msm_device, registered bases) have nothing to hide and never come here.
// With `res` the bases are already on the device (registered handle, possibly with its window table) and only the
// scalars are uploaded slice by slice.
struct ResidentBases {
    const void* d_xy;
    const void* d_inf;
    int tc;
    size_t tstride;
};
int enqueue_sliced(b200msm_ctx* ctx, DevState& d, const Plan& whole, int S, const uint8_t* bases, size_t base_stride, size_t x_off,
                   size_t y_off, size_t inf_off, const uint8_t* scalars, size_t scalar_stride, void* d_out,
                   unsigned long long* launches, const ResidentBases* res = nullptr, int ratio_pct = 0) {
    if (scalar_stride % 8 || scalar_stride < 32) return fail(B200MSM_EINVAL, "scalar stride must be a multiple of 8 and >= 32");
    const size_t n = whole.n;
    // Slice lengths grow geometrically: slice k+1's transfer has to fit under slice k's arithmetic, and on B200 behind
    // PCIe gen5 the arithmetic of a point range takes ~1.7x its transfer (measured, 2^20 and 2^22), so a ratio of 1.6
    // keeps the copy stream ahead while the first (exposed) transfer stays short.
    std::vector<std::pair<size_t, size_t>> sl;
    // auto ratio: the more slices, the flatter the growth (measured with the slice counts below, profiles/r02_e2e_slices_inplace*.jsonl)
    const int ratio_auto = S <= 3 ? 160 : S <= 5 ? 140 : 125;
    slice_ranges(n, S, (ratio_pct > 0 ? ratio_pct : ctx->opt_slice_ratio > 0 ? ctx->opt_slice_ratio : ratio_auto) / 100.0, &sl);
    S = (int)sl.size();
    size_t max_len = 0;
    for (auto& r : sl) max_len = std::max(max_len, r.second);
    std::vector<Plan> plans(S);
    for (int k = 0; k < S; k++) {
        RET_TRY(make_plan(ctx, d, sl[k].second, &plans[k], whole.c, whole.glv, whole.tstride));
        RET_TRY(ensure_work(d, plans[k], k, k > 0 && !whole.ba));
    }
    RET_TRY(ensure_reduce(d, whole));
    RET_TRY(d.scalars.ensure(n * 32));
    if (!res) RET_TRY(d.bases.ensure(n * 64));
    if (whole.glv) RET_TRY(d.xb.ensure(n * 32));
    if (scalar_stride != 32) RET_TRY(d.scalars_raw.ensure(max_len * scalar_stride));
    const bool packed = base_stride == 64 && x_off == 0 && y_off == 32 && inf_off == B200MSM_NO_INF;
    if (!res && !packed) RET_TRY(d.raw.ensure(max_len * base_stride));
    const bool timing = ctx->opt_timing != 0;
    const bool pg_sc = host_is_pageable(scalars), pg_b = !res && host_is_pageable(bases);
    cudaStream_t s = d.stream, cs = d.stream2;
    if (timing) CU_TRY(cudaEventRecord(d.ev[EV_START], s));
    CU_TRY(cudaEventRecord(d.ev_acc[7], s));   // the copy stream starts after earlier main-stream work (buffer reuse)
    CU_TRY(cudaStreamWaitEvent(cs, d.ev_acc[7], 0));
    // K1 + K2 of slices 1.. run on the high-priority sort stream: they need only the slice's scalars (which land before its
    // bases) and their short, barrier-bound kernels would otherwise sit between two accumulations on the main stream.
    // Measured (profiles/r02_sort_overlap*.jsonl): 2^20 4.66 -> 4.50 ms (the two later slices' ~0.08 ms barrier chains leave the
    // critical path); from 2^22 the sort is HBM-bound, sits in the shadow of the transfer anyway and only disturbs the
    // accumulation it runs beside (2^24, 6 slices: 47.2 -> 48.8 ms), so auto = below 3 * 2^20 points.
    const bool sort_ahead = ctx->opt_sort_overlap < 0 ? n < (3u << 20) : ctx->opt_sort_overlap != 0;
    cudaStream_t ss = d.stream4;
    if (sort_ahead) CU_TRY(cudaStreamWaitEvent(ss, d.ev_acc[7], 0));
    const bool measure = !res && !pg_sc && !pg_b;   // pinned host input: the copy stream runs DMA back to back
    d.cp_valid = false;
    if (measure) CU_TRY(cudaEventRecord(d.ev_cp_begin, cs));
    int nlaunch = 0;
    merge_srcs ms = {};
    for (int k = 0; k < S; k++) {
        const Plan& p = plans[k];
        const size_t off = sl[k].first, len = sl[k].second;
        uint8_t* d_sc = (uint8_t*)d.scalars.p + off * 32;
        // a table handle is indexed w * tstride + i: offsetting the base pointer by the slice start keeps that true
        const uint8_t* d_xy = res ? (const uint8_t*)res->d_xy + off * 64 : (const uint8_t*)d.bases.p + off * 64;
        const uint8_t* d_inf = res && res->d_inf ? (const uint8_t*)res->d_inf + off : nullptr;
        fq* d_xb = whole.glv ? (fq*)d.xb.p + off : nullptr;
        RET_TRY(upload_scalars_to(d, scalars + off * scalar_stride, scalar_stride, len, d_sc, cs, launches, pg_sc));
        CU_TRY(cudaEventRecord(d.ev_slice[2 * k], cs));
        if (!res) {
            RET_TRY(upload_bases(d, bases + off * base_stride, base_stride, x_off, y_off, inf_off, len, (void*)d_xy, nullptr, launches, cs,
                                 pg_b));
            CU_TRY(cudaEventRecord(d.ev_slice[2 * k + 1], cs));
        }
        WorkView w = view_slice(d, k);
        // the slices add up IN PLACE in slice 0's bucket array (k_accumulate `into`); the batched-affine variant keeps the
        // separate arrays + merge pass
        const bool into = k > 0 && !whole.ba;
        if (into) w.buckets = d.buckets.p;
        if (sort_ahead && k > 0) {
            CU_TRY(cudaStreamWaitEvent(ss, d.ev_slice[2 * k], 0));
            RET_TRY(launch_sort(w, p, d_sc, d_inf, ss, nullptr));
            CU_TRY(cudaEventRecord(d.ev_sorted[k], ss));
            CU_TRY(cudaStreamWaitEvent(s, d.ev_sorted[k], 0));
        } else {
            CU_TRY(cudaStreamWaitEvent(s, d.ev_slice[2 * k], 0));
            if (timing && k == 0) CU_TRY(cudaEventRecord(d.ev[EV_H2D], s));
            RET_TRY(launch_sort(w, p, d_sc, d_inf, s, timing && k == 0 ? d.ev[EV_DECOMP] : nullptr));
            if (timing && k == 0) CU_TRY(cudaEventRecord(d.ev[EV_SORT], s));
        }
        if (!res) CU_TRY(cudaStreamWaitEvent(s, d.ev_slice[2 * k + 1], 0));
        if (whole.glv) k_endo_x<<<cdiv(len, 256), 256, 0, s>>>((const affine_t*)d_xy, (uint32_t)len, d_xb);
        CU_TRY(cudaMemsetAsync((uint32_t*)w.wtotal + 64, 0, 4 * 48, s));
        RET_TRY(launch_accumulate(d, w, p, d_xy, d_xb, 0, p.Wb, s, into));
        RET_TRY(launch_fixup(d, w, p, 0, p.Wb, 0, s, into));
        nlaunch += (whole.glv ? 5 : 4) + 1 + (p.fix_chunks ? (into ? 3 : 4) : 2);   // sort (+ endo), accumulate, fix-up kernels
        if (k > 0) ms.p[k - 1] = (const xyzz_t*)w.buckets;
    }
    if (measure) CU_TRY(cudaEventRecord(d.ev_cp_end, cs));
    if (whole.ba) {
        k_merge_buckets<<<cdiv(whole.G, 128), 128, 0, s>>>((xyzz_t*)d.buckets.p, ms, S - 1, whole.G);
        nlaunch += 1;
    }
    if (timing) CU_TRY(cudaEventRecord(d.ev[EV_ACC], s));
    RET_TRY(launch_reduce(d, whole, d.buckets.p, 0, whole.Wb, true, s, d_out, &nlaunch));
    if (timing) CU_TRY(cudaEventRecord(d.ev[EV_RED], s));
    if (measure) {
        CU_TRY(cudaEventRecord(d.ev_call_end, s));
        // resident-input time of this MSM on a B200: 0.155 ns per bucket entry (K3) + 0.25 ns per point (K1, K2, fix-up) + 0.9 ms
        // (K4, K5); 2^20: 3.8 ms, 2^24: 39 ms against 3.7 / 41.4 measured
        d.cp_compute_est_ms = ((double)whole.W * whole.n_eff * 1.55e-7 + (double)whole.n * 2.5e-7 + 0.9) * 148.0 / std::max(1, d.sm_count);
        d.cp_valid = true;
    }
    CU_TRY(cudaGetLastError());
    if (launches) *launches += nlaunch;
    return B200MSM_OK;
}

// After every shard's pipeline has been enqueued with its partial copied to h_pinned slot k:
// synchronise, then (if more than one shard) add the partials on the first shard's device.
int finish_and_combine(b200msm_ctx* ctx, const std::vector<int>& dev_indices, uint8_t* slots, uint64_t out[12],
                       unsigned long long* launches) {
    for (int di : dev_indices) {
        CU_TRY(cudaSetDevice(ctx->devs[di].ordinal));
        CU_TRY(cudaStreamSynchronize(ctx->devs[di].stream));
    }
    if (dev_indices.size() == 1) {
        std::memcpy(out, slots, 96);
        return B200MSM_OK;
    }
    DevState& d0 = ctx->devs[dev_indices[0]];
    CU_TRY(cudaSetDevice(d0.ordinal));
    size_t cnt = dev_indices.size();
    RET_TRY(d0.partials.ensure(cnt * 96 + 96));
    CU_TRY(cudaMemcpyAsync(d0.partials.p, slots, cnt * 96, cudaMemcpyHostToDevice, d0.stream));
    jac_t* dout = (jac_t*)((uint8_t*)d0.partials.p + cnt * 96);
    k_sum_partials<<<1, 32, 0, d0.stream>>>((const jac_t*)d0.partials.p, (int)cnt, dout);
    CU_TRY(cudaGetLastError());
    if (launches) *launches += 1;
    CU_TRY(cudaMemcpyAsync(slots, dout, 96, cudaMemcpyDeviceToHost, d0.stream));
    CU_TRY(cudaStreamSynchronize(d0.stream));
    std::memcpy(out, slots, 96);
    return B200MSM_OK;
}

void shard_ranges(size_t n, size_t parts, std::vector<std::pair<size_t, size_t>>* out) {
    out->clear();
    size_t per = (n + parts - 1) / parts;
    for (size_t k = 0; k < parts; k++) {
        size_t b = std::min(n, k * per), e = std::min(n, (k + 1) * per);
        if (e > b) out->push_back({b, e - b});
    }
}

}  // namespace

// =========================================================================================== C ABI
extern "C" {

const char* b200msm_last_error(const b200msm_ctx*) { return g_err.c_str(); }
const char* b200msm_build_id(void) { return B200MSM_BUILD_ID; }

int b200msm_create(b200msm_ctx** out, const int* devices, int n_devices) try {
    if (!out) return fail(B200MSM_EINVAL, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(B200MSM_ENODEV, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count == 0"));
    std::vector<int> ords;
    if (!devices || n_devices <= 0) ords.push_back(0);
    else ords.assign(devices, devices + n_devices);
    b200msm_ctx* ctx = new (std::nothrow) b200msm_ctx();
    if (!ctx) return fail(B200MSM_ENOMEM, "out of host memory");
    for (int o : ords) {
        if (o < 0 || o >= count) {
            delete ctx;
            return fail(B200MSM_ENODEV, "device ordinal out of range");
        }
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, o) != cudaSuccess || prop.major < 10) {
            delete ctx;
            return fail(B200MSM_ENODEV, "device is not sm_100 or newer (this library is built for sm_100a only)");
        }
        DevState d;
        d.ordinal = o;
        d.sm_count = prop.multiProcessorCount;
        d.hw_sm_count = prop.multiProcessorCount;
        ctx->devs.push_back(d);
    }
    for (auto& d : ctx->devs) {
        if (cudaSetDevice(d.ordinal) != cudaSuccess || cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess) {
            b200msm_destroy(ctx);
            return fail(B200MSM_ECUDA, "stream creation failed");
        }
        for (int k = 0; k < EV_COUNT; k++) cudaEventCreate(&d.ev[k]);
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (cudaStreamCreateWithPriority(&d.stream2, cudaStreamNonBlocking, prio_hi) != cudaSuccess) {
            b200msm_destroy(ctx);
            return fail(B200MSM_ECUDA, "side stream creation failed");
        }
        if (cudaStreamCreateWithFlags(&d.stream3, cudaStreamNonBlocking) != cudaSuccess) {
            b200msm_destroy(ctx);
            return fail(B200MSM_ECUDA, "copy stream creation failed");
        }
        if (cudaStreamCreateWithPriority(&d.stream4, cudaStreamNonBlocking, prio_hi) != cudaSuccess) {
            b200msm_destroy(ctx);
            return fail(B200MSM_ECUDA, "sort stream creation failed");
        }
        for (int k = 0; k < MAX_SLICES; k++) cudaEventCreateWithFlags(&d.ev_sorted[k], cudaEventDisableTiming);
        cudaEventCreate(&d.ev_cp_begin);
        cudaEventCreate(&d.ev_cp_end);
        cudaEventCreate(&d.ev_call_end);
        for (int k = 0; k < 8; k++) cudaEventCreateWithFlags(&d.ev_acc[k], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&d.ev_done, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&d.ev_bases, cudaEventDisableTiming);
        for (int k = 0; k < 2 * MAX_SLICES; k++) cudaEventCreateWithFlags(&d.ev_slice[k], cudaEventDisableTiming);
    }
    // staging copy threads: 6 saturate the upload (profiles/r01h_pageable_threads.jsonl: 2^24 from pageable memory 173 ms with
    // one thread, 62 ms with six, 59.5 ms from pinned memory); leave room for one process per GPU on the same host
    const int hw = (int)std::thread::hardware_concurrency();
    const int copy_threads = std::max(1, std::min(6, hw > 0 ? hw / std::max(1, count) : 4));
    ctx->pool = new (std::nothrow) CopyPool(copy_threads - 1);
    for (auto& d : ctx->devs) d.pool = ctx->pool;
    cudaSetDevice(ctx->devs[0].ordinal);
    ctx->h_pinned_bytes = 1 << 16;
    if (cudaMallocHost((void**)&ctx->h_pinned, ctx->h_pinned_bytes) != cudaSuccess) {
        b200msm_destroy(ctx);
        return fail(B200MSM_ENOMEM, "pinned staging allocation failed");
    }
    *out = ctx;
    return B200MSM_OK;
} B200_CATCH

void b200msm_destroy(b200msm_ctx* ctx) {
    if (!ctx) return;
    for (auto& d : ctx->devs) {
        cudaSetDevice(d.ordinal);
        if (d.occ_flag) {
            *d.occ_flag = 1;
            if (d.occ_stream) { cudaStreamSynchronize(d.occ_stream); cudaStreamDestroy(d.occ_stream); }
            cudaFreeHost(d.occ_flag);
            cudaFree(d.occ_started);
        }
        if (d.stream) cudaStreamSynchronize(d.stream);
        for (Buf* b : {&d.digits, &d.ranks, &d.skeys, &d.parts, &d.bslots, &d.chunkg, &d.giant, &d.ends, &d.wtotal, &d.entries, &d.buckets, &d.head, &d.tail, &d.wpart, &d.out, &d.longlist, &d.xb, &d.redbuf, &d.ba_scratch, &d.raw, &d.bases,
                       &d.infmask, &d.scalars_raw, &d.scalars, &d.scalars_alt, &d.partials, &d.g2_bases, &d.g2_buckets, &d.g2_head, &d.g2_tail,
                       &d.g2_wpart, &d.g2_out, &d.g2_redbuf})
            b->release();
        for (auto& e : d.extra)
            for (Buf* b : {&e.g2_buckets, &e.g2_head, &e.g2_tail, &e.digits, &e.ranks, &e.skeys, &e.parts, &e.bslots, &e.chunkg, &e.giant, &e.ends, &e.wtotal, &e.entries, &e.buckets, &e.head, &e.tail, &e.longlist}) b->release();
        for (int k = 0; k < 2 * MAX_SLICES; k++)
            if (d.ev_slice[k]) cudaEventDestroy(d.ev_slice[k]);
        for (int k = 0; k < EV_COUNT; k++)
            if (d.ev[k]) cudaEventDestroy(d.ev[k]);
        if (d.stream && d.owns_stream) cudaStreamDestroy(d.stream);
        if (d.stream2) { cudaStreamSynchronize(d.stream2); cudaStreamDestroy(d.stream2); }
        if (d.stream3) { cudaStreamSynchronize(d.stream3); cudaStreamDestroy(d.stream3); }
        if (d.stream4) { cudaStreamSynchronize(d.stream4); cudaStreamDestroy(d.stream4); }
        for (int k = 0; k < MAX_SLICES; k++)
            if (d.ev_sorted[k]) cudaEventDestroy(d.ev_sorted[k]);
        for (cudaEvent_t e : {d.ev_cp_begin, d.ev_cp_end, d.ev_call_end})
            if (e) cudaEventDestroy(e);
        for (int k = 0; k < STAGE_SLOTS; k++) {
            if (d.stage[k]) cudaFreeHost(d.stage[k]);
            if (d.stage_ev[k]) cudaEventDestroy(d.stage_ev[k]);
        }
        for (int k = 0; k < 8; k++) if (d.ev_acc[k]) cudaEventDestroy(d.ev_acc[k]);
        if (d.ev_done) cudaEventDestroy(d.ev_done);
        if (d.ev_bases) cudaEventDestroy(d.ev_bases);
    }
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    delete ctx->pool;
    delete ctx;
}

int b200msm_device_count(const b200msm_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }

int b200msm_set_option(b200msm_ctx* ctx, const char* key, long long value) try {
    if (!ctx || !key) return fail(B200MSM_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    std::string k(key);
    if (k == "window_bits") {
        if (value != 0 && (value < 4 || value > 24)) return fail(B200MSM_EINVAL, "window_bits must be 0 (auto) or in [4, 24]");
        ctx->opt_window_bits = (int)value;
    } else if (k == "chunk") {
        if (value < 0 || value > 4096) return fail(B200MSM_EINVAL, "chunk must be in [0, 4096]");
        ctx->opt_chunk = (int)value;
    } else if (k == "reduce_log2") {
        if (value < -1 || value > 16) return fail(B200MSM_EINVAL, "reduce_log2 must be in [-1, 16]");
        ctx->opt_reduce_log2 = (int)value;
    } else if (k == "coop_reduce") {
        if (value < -1 || value > 1) return fail(B200MSM_EINVAL, "coop_reduce must be -1 (auto), 0 or 1");
        ctx->opt_coop_reduce = (int)value;
    } else if (k == "glv") {
        if (value < -1 || value > 1) return fail(B200MSM_EINVAL, "glv must be -1 (auto), 0 or 1");
        ctx->opt_glv = (int)value;
    } else if (k == "groups") {
        if (value < 0 || value > 8) return fail(B200MSM_EINVAL, "groups must be in [0, 8]");
        ctx->opt_groups = (int)value;
    } else if (k == "precompute") {
        if (value != 0 && value != 1 && (value < 8 || value > 24)) return fail(B200MSM_EINVAL, "precompute must be 0, 1 (auto window) or a window size in [8, 24]");
        ctx->opt_precompute = (int)value;
    } else if (k == "copy_threads") {
        if (value < 1 || value > 32) return fail(B200MSM_EINVAL, "copy_threads must be in [1, 32]");
        for (auto& d : ctx->devs) {   // nothing may be in flight on the staging ring while the pool is swapped
            cudaSetDevice(d.ordinal);
            cudaStreamSynchronize(d.stream);
            cudaStreamSynchronize(d.stream2);
            cudaStreamSynchronize(d.stream3);
        }
        delete ctx->pool;
        ctx->pool = new (std::nothrow) CopyPool((int)value - 1);
        for (auto& d : ctx->devs) d.pool = ctx->pool;
    } else if (k == "rowcol_reduce") {
        if (value < -1 || value > 1) return fail(B200MSM_EINVAL, "rowcol_reduce must be -1 (auto), 0 or 1");
        ctx->opt_rowcol = (int)value;
    } else if (k == "fix_chunks") {
        if (value < -1 || value > 1) return fail(B200MSM_EINVAL, "fix_chunks must be -1 (auto), 0 or 1");
        ctx->opt_fix_chunks = (int)value;
    } else if (k == "ranked_sort") {
        if (value < -1 || value > 2) return fail(B200MSM_EINVAL, "ranked_sort must be -1 (auto), 0 (cursor atomics), 1 (ranked) or 2 (partitioned)");
        ctx->opt_ranked_sort = (int)value;
    } else if (k == "slice_ratio") {
        if (value != 0 && (value < 100 || value > 400)) return fail(B200MSM_EINVAL, "slice_ratio (percent) must be 0 (auto) or in [100, 400]");
        ctx->opt_slice_ratio = (int)value;
    } else if (k == "adaptive_slices") {
        if (value < -1 || value > 1) return fail(B200MSM_EINVAL, "adaptive_slices must be -1 (auto = on), 0 or 1");
        ctx->opt_adaptive_slices = (int)value;
        if (value == 0) for (auto& d : ctx->devs) d.copy_bound = false;
    } else if (k == "sort_overlap") {
        if (value < -1 || value > 1) return fail(B200MSM_EINVAL, "sort_overlap must be -1 (auto), 0 or 1");
        ctx->opt_sort_overlap = (int)value;
    } else if (k == "slices") {
        if (value < 0 || value > MAX_SLICES) return fail(B200MSM_EINVAL, "slices must be in [0, 8]");
        ctx->opt_slices = (int)value;
    } else if (k == "sm_count") {
        // SMs the policy and the persistent grids should assume (0 = the device's): a MIG slice / green context / shared GPU
        if (value < 0 || value > 1024) return fail(B200MSM_EINVAL, "sm_count must be in [0, 1024]");
        for (auto& d : ctx->devs) d.sm_count = value ? (int)value : d.hw_sm_count;
    } else if (k == "batch_affine") {
        if (value < -1 || value > 1) return fail(B200MSM_EINVAL, "batch_affine must be -1 (auto), 0 or 1");
        ctx->opt_batch_affine = (int)value;
    } else if (k == "ba_chunk") {
        if (value != 0 && (value < 32 || value > BA_MAXL || (value & 15))) return fail(B200MSM_EINVAL, "ba_chunk must be 0 (auto) or a multiple of 16 in [32, 512]");
        ctx->opt_ba_chunk = (int)value;
    } else if (k == "ba_min_pairs") {
        if (value < 0 || value > 256) return fail(B200MSM_EINVAL, "ba_min_pairs must be in [0, 256]");
        ctx->opt_ba_min_pairs = (int)value;
    } else if (k == "timing") {
        ctx->opt_timing = value != 0;
    } else {
        return fail(B200MSM_EINVAL, "unknown option: " + k);
    }
    return B200MSM_OK;
} B200_CATCH

int b200msm_last_timings(const b200msm_ctx* ctx, b200msm_timings* out) try {
    if (!ctx || !out) return fail(B200MSM_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(const_cast<b200msm_ctx*>(ctx)->mu);
    *out = ctx->last;
    return B200MSM_OK;
} B200_CATCH

int b200msm_last_sort_engine(const b200msm_ctx* ctx) try {
    if (!ctx) return fail(B200MSM_EINVAL, "null context");
    std::lock_guard<std::mutex> lk(const_cast<b200msm_ctx*>(ctx)->mu);
    return ctx->last_plan.psort ? 2 : ctx->last_plan.ranked ? 1 : 0;
} B200_CATCH

int b200msm_auto_window_bits(const b200msm_ctx* ctx, size_t n) try {
    if (!ctx || n == 0) return fail(B200MSM_EINVAL, "null context or n == 0");
    return auto_window_bits(n, ctx->devs[0].sm_count);
} B200_CATCH

void* b200msm_stream(b200msm_ctx* ctx, int dev_index) {
    if (!ctx || dev_index < 0 || dev_index >= (int)ctx->devs.size()) return nullptr;
    return (void*)ctx->devs[dev_index].stream;
}

int b200msm_set_stream(b200msm_ctx* ctx, int dev_index, void* stream) try {
    if (!ctx || dev_index < 0 || dev_index >= (int)ctx->devs.size()) return fail(B200MSM_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState& d = ctx->devs[dev_index];
    CU_TRY(cudaSetDevice(d.ordinal));
    CU_TRY(cudaStreamSynchronize(d.stream));
    if (d.owns_stream && d.stream) cudaStreamDestroy(d.stream);
    d.stream = (cudaStream_t)stream;
    d.owns_stream = false;
    return B200MSM_OK;
} B200_CATCH

int b200msm_sync(b200msm_ctx* ctx) try {
    if (!ctx) return fail(B200MSM_EINVAL, "null context");
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (auto& d : ctx->devs) {
        CU_TRY(cudaSetDevice(d.ordinal));
        CU_TRY(cudaStreamSynchronize(d.stream));
    }
    if (ctx->pending_timings_dev >= 0) {   // stage timings of the last asynchronous b200msm_msm_device
        DevState& d = ctx->devs[ctx->pending_timings_dev];
        ctx->pending_timings_dev = -1;
        CU_TRY(cudaSetDevice(d.ordinal));
        RET_TRY(collect_timings(ctx, d, ctx->last_plan, false));
    }
    return B200MSM_OK;
} B200_CATCH

int b200msm_msm_device(b200msm_ctx* ctx, int dev_index, const void* d_bases, const void* d_inf_mask, const void* d_scalars,
                       size_t n, void* d_out, int sync) try {
    if (!ctx || !d_bases || !d_scalars || !d_out) return fail(B200MSM_EINVAL, "null argument");
    if (dev_index < 0 || dev_index >= (int)ctx->devs.size()) return fail(B200MSM_EINVAL, "bad dev_index");
    if (n == 0) return fail(B200MSM_EINVAL, "Empty input");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState& d = ctx->devs[dev_index];
    CU_TRY(cudaSetDevice(d.ordinal));
    Plan p;
    RET_TRY(make_plan(ctx, d, n, &p));
    RET_TRY(ensure_workspace(d, p));
    ctx->last.kernel_launches = 0;
    if (ctx->opt_timing) CU_TRY(cudaEventRecord(d.ev[EV_H2D], d.stream));
    RET_TRY(enqueue_msm(ctx, d, p, d_bases, d_inf_mask, d_scalars, d_out, &ctx->last.kernel_launches));
    ctx->last_plan = p;
    if (sync) {
        CU_TRY(cudaStreamSynchronize(d.stream));
        RET_TRY(collect_timings(ctx, d, p, false));
    } else {
        ctx->last.window_bits = p.c;
        ctx->last.num_windows = p.W;
        ctx->pending_timings_dev = dev_index;   // b200msm_sync() reads the stage events once the stream has drained
    }
    return B200MSM_OK;
} B200_CATCH

int b200msm_sum_partials_device(b200msm_ctx* ctx, int dev_index, const void* d_partials, int count, void* d_out, int sync) try {
    if (!ctx || !d_partials || !d_out || count <= 0) return fail(B200MSM_EINVAL, "bad argument");
    if (dev_index < 0 || dev_index >= (int)ctx->devs.size()) return fail(B200MSM_EINVAL, "bad dev_index");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState& d = ctx->devs[dev_index];
    CU_TRY(cudaSetDevice(d.ordinal));
    k_sum_partials<<<1, 32, 0, d.stream>>>((const jac_t*)d_partials, count, (jac_t*)d_out);
    CU_TRY(cudaGetLastError());
    if (sync) CU_TRY(cudaStreamSynchronize(d.stream));
    return B200MSM_OK;
} B200_CATCH

int b200msm_bn254_g1_msm(b200msm_ctx* ctx, const void* bases, size_t base_stride, size_t x_off, size_t y_off, size_t inf_off,
                         const void* scalars, size_t scalar_stride, size_t n, uint64_t out_jacobian[12]) try {
    if (!ctx || !out_jacobian) return fail(B200MSM_EINVAL, "null argument");
    if (n == 0 || !bases || !scalars) return fail(B200MSM_EINVAL, "Empty input");
    RET_TRY(check_layout(base_stride, x_off, y_off, inf_off));
    std::lock_guard<std::mutex> lk(ctx->mu);
    std::vector<std::pair<size_t, size_t>> ranges;
    shard_ranges(n, ctx->devs.size(), &ranges);
    std::vector<int> used;
    ctx->last.kernel_launches = 0;
    Plan plan0;
    // One device = one enqueue job.  With several devices every job runs on its own host thread, so that the uploads of
    // all shards are in flight together: staging pageable memory blocks the enqueuing thread (the copy threads fill a pinned
    // ring slot while the DMA engine drains the previous ones), and a sequential loop would feed the GPUs one after another.
    const size_t ndev = ranges.size();
    std::vector<int> rcs(ndev, B200MSM_OK);
    std::vector<std::string> errs(ndev);
    std::vector<unsigned long long> nlaunch(ndev, 0);
    std::vector<Plan> plans(ndev);
    auto job = [&](size_t k) -> int {
        DevState& d = ctx->devs[k];
        CU_TRY(cudaSetDevice(d.ordinal));
        size_t begin = ranges[k].first, len = ranges[k].second;
        Plan p;
        RET_TRY(make_plan(ctx, d, len, &p));
        plans[k] = p;
        // "slices": 0 = auto, 1 = off.  Measured on B200 behind PCIe gen5 (profiles/r01e_e2e_slices.jsonl): 2 slices pay
        // from 2^17 points per device, 3 from ~2^20 (2^20: 6.6 -> 5.0 ms, 2^22: 20.8 -> 16.4, 2^24: 73.9 -> 59.5).  Since the
        // slices add up in place (k_accumulate `into`: no per-slice bucket array, no merge pass) a further slice only costs its
        // sort launches and fix-up, and the arithmetic of 2^22+ points is longer than their transfer, so more and flatter slices
        // shorten the exposed first transfer (profiles/r02_e2e_slices_inplace*.jsonl: 2^22 14.8 -> 14.4 ms with 4 slices,
        // 2^24 54.4 (3 slices, ratio 1.6) -> 47.3 (6 slices, ratio 1.25); 2^20: 3 and 4 slices equal).
        int S = ctx->opt_slices > 0 ? ctx->opt_slices : len >= (3u << 22) ? 6 : len >= (3u << 20) ? 4 : len >= (3u << 18) ? 3 : len >= (1u << 17) ? 2 : 1;
        // Feedback from the previous sliced call on this device (it was synchronous, so its events are complete): the time its
        // uploads took against what its arithmetic takes with resident inputs.  Alone on the host a B200 gets ~50 GB/s (2^20:
        // 2.1 ms of transfer under 3.7 ms of arithmetic); with eight ranks uploading at once it gets ~23 GB/s, the call is
        // TRANSFER-bound and what counts is the work left after the last byte has landed: equal slices keep the last one small
        // (profiles/r02_e2e_slices_8ranks_2e2{0,1}.jsonl: 8 ranks, 2^20 per GPU 6.92 -> 6.21 ms with 4 equal slices, 2^21
        // 13.0 -> 10.97 with 8).  Hysteresis: on when the transfer exceeds the arithmetic, off below 0.85 x (one GPU alone: 0.56).
        int ratio_pct = 0;
        if (ctx->opt_adaptive_slices != 0 && ctx->opt_slices == 0 && ctx->opt_slice_ratio == 0) {
            if (d.cp_valid) {
                float cp_ms = 0;
                if (cudaEventElapsedTime(&cp_ms, d.ev_cp_begin, d.ev_cp_end) == cudaSuccess && d.cp_compute_est_ms > 0) {
                    if (cp_ms > 1.0 * d.cp_compute_est_ms) d.copy_bound = true;
                    else if (cp_ms < 0.85 * d.cp_compute_est_ms) d.copy_bound = false;
                } else {
                    cudaGetLastError();
                }
                d.cp_valid = false;
            }
            if (d.copy_bound && len >= (3u << 18)) {
                S = len >= (1u << 21) ? 8 : 4;
                ratio_pct = 100;
            }
        }
        if (p.ngroups > 1) S = 1;
        S = (int)std::min<size_t>((size_t)S, len);
        if (S > 1) {
            RET_TRY(ensure_reduce(d, p));   // d.out must exist before its address is passed on
            RET_TRY(enqueue_sliced(ctx, d, p, S, (const uint8_t*)bases + begin * base_stride, base_stride, x_off, y_off, inf_off,
                                   (const uint8_t*)scalars + begin * scalar_stride, scalar_stride, d.out.p, &nlaunch[k], nullptr, ratio_pct));
        } else {
            RET_TRY(ensure_workspace(d, p));
            RET_TRY(d.bases.ensure(len * 64));
            if (ctx->opt_timing) CU_TRY(cudaEventRecord(d.ev[EV_START], d.stream));
            void* d_scalars = nullptr;
            // Scalars go first on the main stream; decomposition and the sort do not need the bases, which are
            // uploaded and repacked on the side stream meanwhile (infinity records become the (0,0) marker that
            // k_accumulate skips).  The main stream waits for them just before the accumulation.
            RET_TRY(upload_scalars(d, (const uint8_t*)scalars + begin * scalar_stride, scalar_stride, len, &d_scalars, &nlaunch[k],
                                   host_is_pageable(scalars)));
            CU_TRY(cudaEventRecord(d.ev_acc[7], d.stream));          // orders the side stream after earlier main-stream work
            CU_TRY(cudaStreamWaitEvent(d.stream2, d.ev_acc[7], 0));
            RET_TRY(upload_bases(d, (const uint8_t*)bases + begin * base_stride, base_stride, x_off, y_off, inf_off, len, d.bases.p,
                                 nullptr, &nlaunch[k], d.stream2, host_is_pageable(bases)));
            CU_TRY(cudaEventRecord(d.ev_bases, d.stream2));
            if (ctx->opt_timing) CU_TRY(cudaEventRecord(d.ev[EV_H2D], d.stream));
            RET_TRY(enqueue_msm(ctx, d, p, d.bases.p, nullptr, d_scalars, d.out.p, &nlaunch[k], d.ev_bases));
        }
        CU_TRY(cudaMemcpyAsync(ctx->h_pinned + k * 96, d.out.p, 96, cudaMemcpyDeviceToHost, d.stream));
        return B200MSM_OK;
    };
    if (ndev == 1) {
        RET_TRY(job(0));
    } else {
        std::vector<std::thread> workers;
        for (size_t k = 0; k < ndev; k++)
            workers.emplace_back([&, k] {
                try {
                    rcs[k] = job(k);
                } catch (const std::bad_alloc&) {
                    rcs[k] = fail(B200MSM_ENOMEM, "out of host memory");
                } catch (...) {
                    rcs[k] = fail(B200MSM_ECUDA, "internal error");
                }
                if (rcs[k] != B200MSM_OK) errs[k] = g_err;   // the message is thread-local: hand it to the caller
            });
        for (auto& t : workers) t.join();
        for (size_t k = 0; k < ndev; k++)
            if (rcs[k] != B200MSM_OK) {
                for (auto& dd : ctx->devs) {   // do not leave work of the other devices in flight behind an error return
                    cudaSetDevice(dd.ordinal);
                    cudaStreamSynchronize(dd.stream);
                }
                return fail(rcs[k], errs[k]);
            }
    }
    plan0 = plans[0];
    for (size_t k = 0; k < ndev; k++) {
        ctx->last.kernel_launches += nlaunch[k];
        used.push_back((int)k);
    }
    RET_TRY(finish_and_combine(ctx, used, ctx->h_pinned, out_jacobian, &ctx->last.kernel_launches));
    ctx->last_plan = plan0;
    CU_TRY(cudaSetDevice(ctx->devs[0].ordinal));
    return collect_timings(ctx, ctx->devs[0], plan0, true);
} B200_CATCH

// ------------------------------------------------------------------------------------------- G2
struct b200msm_g2_bases {
    int dev_index = 0;
    size_t n = 0;
    void* d_pts = nullptr;   // n x 128 B, or the [tW][n] window table whose window 0 is the bases
    int tc = 0, tW = 0;
};

namespace {

// K4 shape for G2: 64-thread CTAs, Bsz = 2^lb magnitudes per thread, at most 64 CTAs per window
void g2_reduce_shape(const Plan& p, uint32_t* lb, uint32_t* bpw) {
    uint32_t l = 3;
    while ((((uint64_t)p.half + ((uint64_t)G2_RED_THREADS << l) - 1) / ((uint64_t)G2_RED_THREADS << l)) > G2_RED_THREADS) l++;
    *lb = l;
    *bpw = (uint32_t)(((uint64_t)p.half + ((uint64_t)G2_RED_THREADS << l) - 1) / ((uint64_t)G2_RED_THREADS << l));
}

// K3 of one (sub-)MSM on stream s: chunked accumulation + boundary fix-up into (bk, hd, tl)
int g2_launch_accumulate(DevState& d, const WorkView& w, const Plan& pk, const g2_affine_t* pts, uint32_t n_glv, g2_xyzz_t* bk,
                         g2_xyzz_t* hd, g2_xyzz_t* tl, cudaStream_t s, bool into = false) {
    const uint64_t max_chunks = ((uint64_t)pk.W * pk.n_eff + pk.L - 1) / pk.L + 2;
    k_g2_accumulate<<<cdiv(max_chunks, G2_ACC_THREADS), G2_ACC_THREADS, 0, s>>>(pts, n_glv, (const uint32_t*)w.entries,
                                                                                (const uint32_t*)w.ends, pk.G, pk.L, bk, hd, tl, into ? 1 : 0);
    uint32_t* long_count = (uint32_t*)w.wtotal + 64;
    CU_TRY(cudaMemsetAsync(long_count, 0, 4 * 16, s));
    k_g2_fixup<<<cdiv(pk.G, 128), 128, 0, s>>>((const uint32_t*)w.ends, pk.G, pk.L, bk, hd, tl, long_count, (uint32_t*)w.longlist, into ? 1 : 0);
    k_g2_fixup_long<<<d.sm_count * 2, G2_FIXL_THREADS, 0, s>>>((const uint32_t*)w.ends, pk.L, bk, hd, tl, long_count,
                                                              (const uint32_t*)w.longlist);
    CU_TRY(cudaGetLastError());
    return B200MSM_OK;
}

// K4 + K5 over d.g2_buckets on stream s, result copied to the pinned staging and returned after a stream sync
int g2_reduce_and_read(b200msm_ctx* ctx, DevState& d, const Plan& p, uint64_t out_jacobian[24], int slot = 0, bool wait = true) {
    cudaStream_t s = d.stream;
    uint32_t lb, bpw;
    g2_reduce_shape(p, &lb, &bpw);
    g2_xyzz_t* wpartR = (g2_xyzz_t*)d.g2_wpart.p;
    g2_xyzz_t* wpartT = wpartR + (size_t)p.Wb * bpw;
    g2_xyzz_t* wsum = wpartT + (size_t)p.Wb * bpw;
    if (p.coop_reduce) {
        // recursive weighted sum on the lane-parallel cooperative engine (the G1 levels of launch_reduce over Fq2)
        CU_TRY(cudaFuncSetAttribute(k_g2_reduce_level, cudaFuncAttributeMaxDynamicSharedMemorySize, G2CL_SMEM_BYTES));   // per device
        const g2_xyzz_t* Ain = (const g2_xyzz_t*)d.g2_buckets.p;
        const g2_xyzz_t* Xin = nullptr;
        uint32_t in_stride = p.nb, in_off = 1, cnt = p.half, log2u = 0, delta = 0;
        g2_xyzz_t* buf = (g2_xyzz_t*)d.g2_redbuf.p;
        for (int l = 0; l < p.red_nl; l++) {
            const uint32_t ctas = p.red_ctas[l];
            g2_xyzz_t* Aout = buf;
            g2_xyzz_t* Xout = ctas == 1 ? wsum : buf + (size_t)p.W * ctas;
            k_g2_reduce_level<<<p.Wb * ctas, G2CL_THREADS, G2CL_SMEM_BYTES, s>>>(Ain, Xin, in_stride, in_off, cnt, p.red_lb[l], log2u, delta, ctas,
                                                                                  Aout, Xout);
            Ain = Aout;
            Xin = Xout;
            in_stride = ctas;
            in_off = 0;
            cnt = ctas;
            log2u += 5 + p.red_lb[l];
            delta = 1;
            buf += 2 * (size_t)p.W * ctas;
            ctx->last.kernel_launches += 1;
        }
    } else {
        k_g2_bucket_reduce<<<p.Wb * bpw, G2_RED_THREADS, 0, s>>>((const g2_xyzz_t*)d.g2_buckets.p, p.nb, lb, bpw, wpartR, wpartT);
        k_g2_window_finish<<<p.Wb, G2_RED_THREADS, 0, s>>>(wpartR, wpartT, bpw, lb + 6, wsum);
        ctx->last.kernel_launches += 2;
    }
    k_g2_combine<<<1, G2_CMB_THREADS, 0, s>>>(wsum, p.Wb, p.c, (g2_jac_t*)d.g2_out.p);
    CU_TRY(cudaGetLastError());
    ctx->last.kernel_launches += 1;
    CU_TRY(cudaMemcpyAsync(ctx->h_pinned + (size_t)slot * sizeof(g2_jac_t), d.g2_out.p, sizeof(g2_jac_t), cudaMemcpyDeviceToHost, s));
    ctx->last.window_bits = p.c;
    ctx->last.num_windows = p.W;
    if (!wait) return B200MSM_OK;   // the caller synchronises every shard, then combines
    CU_TRY(cudaStreamSynchronize(s));
    std::memcpy(out_jacobian, ctx->h_pinned + (size_t)slot * sizeof(g2_jac_t), sizeof(g2_jac_t));
    return B200MSM_OK;
}

int g2_check_layout(size_t base_stride, size_t x_off, size_t y_off, size_t inf_off) {
    if (base_stride % 8 || x_off % 8 || y_off % 8) return fail(B200MSM_EINVAL, "base stride/offsets must be multiples of 8");
    if (x_off + 64 > base_stride || y_off + 64 > base_stride) return fail(B200MSM_EINVAL, "x/y offset outside the record");
    if (inf_off != B200MSM_NO_INF && inf_off >= base_stride) return fail(B200MSM_EINVAL, "infinity offset outside the record");
    return B200MSM_OK;
}

}  // namespace

namespace {

// One device's share of a G2 MSM: upload + sort + accumulate (sliced like the G1 host call) + reduce, result copied to
// pinned slot `slot`; with wait = false nothing is synchronised (multi-device: the caller joins all shards).
int g2_msm_shard(b200msm_ctx* ctx, DevState& d, const void* bases, size_t base_stride, size_t x_off, size_t y_off, size_t inf_off,
                 const void* scalars, size_t scalar_stride, size_t n, uint64_t out_jacobian[24], int slot, bool wait) {
    CU_TRY(cudaSetDevice(d.ordinal));
    // same (scalar split, window) policy and options as G1: phi acts on G2 as well (k_g2_accumulate)
    Plan p;
    RET_TRY(make_plan(ctx, d, n, &p));
    // The point range is uploaded and accumulated in slices exactly like the G1 host call (enqueue_sliced): per-slice sort,
    // accumulation and fix-up into per-slice bucket arrays, one merge, one reduce.  S = 1 is the same code with one slice.
    int S = ctx->opt_slices > 0 ? ctx->opt_slices : n >= (3u << 18) ? 3 : n >= (1u << 17) ? 2 : 1;
    std::vector<std::pair<size_t, size_t>> sl;
    slice_ranges(n, S, (ctx->opt_slice_ratio > 0 ? ctx->opt_slice_ratio : 160) / 100.0, &sl);
    S = (int)sl.size();
    size_t max_len = 0;
    for (auto& r : sl) max_len = std::max(max_len, r.second);
    std::vector<Plan> plans(S);
    for (int k = 0; k < S; k++) {
        RET_TRY(make_plan(ctx, d, sl[k].second, &plans[k], p.c, p.glv));
        RET_TRY(ensure_work(d, plans[k], k, true));   // the G1 bucket array of the work set is not used by the G2 kernels
        Buf& hd = k ? d.extra[k - 1].g2_head : d.g2_head;
        Buf& tl = k ? d.extra[k - 1].g2_tail : d.g2_tail;
        if (k == 0) RET_TRY(d.g2_buckets.ensure((size_t)p.G * sizeof(g2_xyzz_t)));   // one bucket array: the slices add up in place
        RET_TRY(hd.ensure((size_t)plans[k].nchunks * sizeof(g2_xyzz_t)));
        RET_TRY(tl.ensure((size_t)plans[k].nchunks * sizeof(g2_xyzz_t)));
    }
    uint32_t lb, bpw;
    g2_reduce_shape(p, &lb, &bpw);
    RET_TRY(d.g2_bases.ensure(n * sizeof(g2_affine_t)));
    RET_TRY(d.g2_wpart.ensure(((size_t)p.Wb * bpw * 2 + p.Wb) * sizeof(g2_xyzz_t)));
    RET_TRY(d.g2_out.ensure(sizeof(g2_jac_t)));
    RET_TRY(d.g2_redbuf.ensure((p.red_slots + 2) * sizeof(g2_xyzz_t)));
    RET_TRY(d.raw.ensure(max_len * base_stride));
    RET_TRY(d.scalars.ensure(n * 32));
    if (scalar_stride != 32) RET_TRY(d.scalars_raw.ensure(max_len * scalar_stride));
    const bool pg_sc = host_is_pageable(scalars), pg_b = host_is_pageable(bases);
    cudaStream_t s = d.stream, cs = d.stream2;
    CU_TRY(cudaEventRecord(d.ev_acc[7], s));
    CU_TRY(cudaStreamWaitEvent(cs, d.ev_acc[7], 0));
    g2_merge_srcs ms = {};
    for (int k = 0; k < S; k++) {
        const Plan& pk = plans[k];
        const size_t off = sl[k].first, len = sl[k].second;
        uint8_t* d_sc = (uint8_t*)d.scalars.p + off * 32;
        g2_affine_t* d_pts = (g2_affine_t*)d.g2_bases.p + off;
        RET_TRY(upload_scalars_to(d, (const uint8_t*)scalars + off * scalar_stride, scalar_stride, len, d_sc, cs,
                                  &ctx->last.kernel_launches, pg_sc));
        CU_TRY(cudaEventRecord(d.ev_slice[2 * k], cs));
        RET_TRY(h2d(d, d.raw.p, (const uint8_t*)bases + off * base_stride, len * base_stride, cs, pg_b));
        k_g2_repack<<<cdiv(len * 16, 256), 256, 0, cs>>>((const uint8_t*)d.raw.p, base_stride, x_off, y_off, inf_off, (uint32_t)len,
                                                         (uint64_t*)d_pts);
        CU_TRY(cudaEventRecord(d.ev_slice[2 * k + 1], cs));
        const WorkView w = view_slice(d, k);
        g2_xyzz_t* bk = (g2_xyzz_t*)d.g2_buckets.p;   // every slice adds into the same bucket array (k_g2_accumulate `into`)
        g2_xyzz_t* hd = (g2_xyzz_t*)(k ? d.extra[k - 1].g2_head.p : d.g2_head.p);
        g2_xyzz_t* tl = (g2_xyzz_t*)(k ? d.extra[k - 1].g2_tail.p : d.g2_tail.p);
        CU_TRY(cudaStreamWaitEvent(s, d.ev_slice[2 * k], 0));
        RET_TRY(launch_sort(w, pk, d_sc, nullptr, s, nullptr));
        CU_TRY(cudaStreamWaitEvent(s, d.ev_slice[2 * k + 1], 0));
        RET_TRY(g2_launch_accumulate(d, w, pk, d_pts, pk.glv ? pk.n : 0xffffffffu, bk, hd, tl, s, k > 0));
        ctx->last.kernel_launches += 8;
    }
    (void)ms;
    return g2_reduce_and_read(ctx, d, p, out_jacobian, slot, wait);
}

}  // namespace

int b200msm_bn254_g2_msm(b200msm_ctx* ctx, const void* bases, size_t base_stride, size_t x_off, size_t y_off, size_t inf_off,
                         const void* scalars, size_t scalar_stride, size_t n, uint64_t out_jacobian[24]) try {
    if (!ctx || !out_jacobian) return fail(B200MSM_EINVAL, "null argument");
    if (n == 0 || !bases || !scalars) return fail(B200MSM_EINVAL, "Empty input");
    RET_TRY(g2_check_layout(base_stride, x_off, y_off, inf_off));
    if (scalar_stride % 8 || scalar_stride < 32) return fail(B200MSM_EINVAL, "scalar stride must be a multiple of 8 and >= 32");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->last.kernel_launches = 0;
    // Point-range sharding over the context's devices, exactly like the G1 call (small inputs stay on the first device:
    // a shard below ~2^12 points is all latency).
    std::vector<std::pair<size_t, size_t>> ranges;
    const size_t want = std::max<size_t>(1, std::min(ctx->devs.size(), n >> 12));
    shard_ranges(n, want, &ranges);
    if (ranges.size() == 1)
        return g2_msm_shard(ctx, ctx->devs[0], bases, base_stride, x_off, y_off, inf_off, scalars, scalar_stride, n, out_jacobian, 0, true);
    if (ranges.size() * sizeof(g2_jac_t) > ctx->h_pinned_bytes) return fail(B200MSM_EINVAL, "too many devices");
    for (size_t k = 0; k < ranges.size(); k++)
        RET_TRY(g2_msm_shard(ctx, ctx->devs[k], (const uint8_t*)bases + ranges[k].first * base_stride, base_stride, x_off, y_off, inf_off,
                             (const uint8_t*)scalars + ranges[k].first * scalar_stride, scalar_stride, ranges[k].second, out_jacobian,
                             (int)k, false));
    for (size_t k = 0; k < ranges.size(); k++) {
        CU_TRY(cudaSetDevice(ctx->devs[k].ordinal));
        CU_TRY(cudaStreamSynchronize(ctx->devs[k].stream));
    }
    DevState& d0 = ctx->devs[0];
    CU_TRY(cudaSetDevice(d0.ordinal));
    const size_t cnt = ranges.size();
    RET_TRY(d0.partials.ensure((cnt + 1) * sizeof(g2_jac_t)));
    CU_TRY(cudaMemcpyAsync(d0.partials.p, ctx->h_pinned, cnt * sizeof(g2_jac_t), cudaMemcpyHostToDevice, d0.stream));
    g2_jac_t* dout = (g2_jac_t*)d0.partials.p + cnt;
    k_g2_sum_partials<<<1, 32, 0, d0.stream>>>((const g2_jac_t*)d0.partials.p, (int)cnt, dout);
    CU_TRY(cudaGetLastError());
    ctx->last.kernel_launches += 1;
    CU_TRY(cudaMemcpyAsync(ctx->h_pinned, dout, sizeof(g2_jac_t), cudaMemcpyDeviceToHost, d0.stream));
    CU_TRY(cudaStreamSynchronize(d0.stream));
    std::memcpy(out_jacobian, ctx->h_pinned, sizeof(g2_jac_t));
    return B200MSM_OK;
} B200_CATCH

// Registered G2 base set (the B2 bases of a proving key are fixed): bases stay on the first device; precompute = 1 / 8..24
// also builds the window table 2^(c w) * P_i (W x 128 B per point), after which an MSM is one bucket set and no Horner chain.
int b200msm_g2_register_bases(b200msm_ctx* ctx, const void* bases, size_t base_stride, size_t x_off, size_t y_off, size_t inf_off,
                              size_t n, int precompute, b200msm_g2_bases** out) try {
    if (!ctx || !bases || !out) return fail(B200MSM_EINVAL, "null argument");
    if (n == 0) return fail(B200MSM_EINVAL, "Empty input");
    if (precompute != 0 && precompute != 1 && (precompute < 8 || precompute > 24))
        return fail(B200MSM_EINVAL, "precompute must be 0, 1 (auto window) or a window size in [8, 24]");
    RET_TRY(g2_check_layout(base_stride, x_off, y_off, inf_off));
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState& d = ctx->devs[0];
    CU_TRY(cudaSetDevice(d.ordinal));
    b200msm_g2_bases* h = new (std::nothrow) b200msm_g2_bases();
    if (!h) return fail(B200MSM_ENOMEM, "out of host memory");
    h->n = n;
    if (precompute) {
        h->tc = precompute >= 8 ? precompute : table_window_bits(n);
        h->tW = num_windows_for(h->tc, 254);
        if ((uint64_t)h->tW * n >= (1ull << 31)) { h->tc = 0; h->tW = 0; }
    }
    const size_t windows = h->tc ? (size_t)h->tW : 1;
    cudaError_t e = cudaMalloc(&h->d_pts, windows * n * sizeof(g2_affine_t));
    int rc = e == cudaSuccess ? d.raw.ensure(n * base_stride) : fail(B200MSM_ENOMEM, std::string("g2_register_bases: ") + cudaGetErrorString(e));
    if (rc == B200MSM_OK) rc = h2d(d, d.raw.p, bases, n * base_stride, d.stream, host_is_pageable(bases));
    if (rc == B200MSM_OK) {
        k_g2_repack<<<cdiv(n * 16, 256), 256, 0, d.stream>>>((const uint8_t*)d.raw.p, base_stride, x_off, y_off, inf_off, (uint32_t)n,
                                                             (uint64_t*)h->d_pts);
        if (h->tc) k_g2_build_table<<<cdiv(n, 64), 64, 0, d.stream>>>((uint32_t)n, n, h->tc, h->tW, (g2_affine_t*)h->d_pts);
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(d.stream) != cudaSuccess)
            rc = fail(B200MSM_ECUDA, "g2_register_bases: upload / table build failed");
    }
    if (rc != B200MSM_OK) {
        std::string keep = g_err;
        if (h->d_pts) cudaFree(h->d_pts);
        delete h;
        g_err = keep;
        return rc;
    }
    *out = h;
    return B200MSM_OK;
} B200_CATCH

int b200msm_g2_release_bases(b200msm_ctx* ctx, b200msm_g2_bases* h) try {
    if (!ctx || !h) return fail(B200MSM_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState& d = ctx->devs[h->dev_index];
    cudaSetDevice(d.ordinal);
    cudaStreamSynchronize(d.stream);
    if (h->d_pts) cudaFree(h->d_pts);
    delete h;
    return B200MSM_OK;
} B200_CATCH

int b200msm_g2_msm_registered(b200msm_ctx* ctx, const b200msm_g2_bases* h, const void* scalars, size_t scalar_stride, size_t n,
                              uint64_t out_jacobian[24]) try {
    if (!ctx || !h || !scalars || !out_jacobian) return fail(B200MSM_EINVAL, "null argument");
    if (n == 0) return fail(B200MSM_EINVAL, "Empty input");
    if (n > h->n) return fail(B200MSM_EINVAL, "more scalars than registered bases");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState& d = ctx->devs[h->dev_index];
    CU_TRY(cudaSetDevice(d.ordinal));
    Plan p;
    if (h->tc) RET_TRY(make_plan(ctx, d, n, &p, h->tc, false, h->n));
    else RET_TRY(make_plan(ctx, d, n, &p));
    RET_TRY(ensure_work(d, p, 0));
    uint32_t lb, bpw;
    g2_reduce_shape(p, &lb, &bpw);
    RET_TRY(d.g2_buckets.ensure((size_t)p.G * sizeof(g2_xyzz_t)));
    RET_TRY(d.g2_head.ensure((size_t)p.nchunks * sizeof(g2_xyzz_t)));
    RET_TRY(d.g2_tail.ensure((size_t)p.nchunks * sizeof(g2_xyzz_t)));
    RET_TRY(d.g2_wpart.ensure(((size_t)p.Wb * bpw * 2 + p.Wb) * sizeof(g2_xyzz_t)));
    RET_TRY(d.g2_out.ensure(sizeof(g2_jac_t)));
    RET_TRY(d.g2_redbuf.ensure((p.red_slots + 2) * sizeof(g2_xyzz_t)));
    ctx->last.kernel_launches = 0;
    void* d_scalars = nullptr;
    RET_TRY(upload_scalars(d, (const uint8_t*)scalars, scalar_stride, n, &d_scalars, &ctx->last.kernel_launches,
                           host_is_pageable(scalars)));
    const WorkView w = view_main(d);
    RET_TRY(launch_sort(w, p, d_scalars, nullptr, d.stream, nullptr));
    // table handle: entries index the [W][n] table directly; plain handle: pseudo-points >= n are phi(P)
    RET_TRY(g2_launch_accumulate(d, w, p, (const g2_affine_t*)h->d_pts, p.tstride || !p.glv ? 0xffffffffu : p.n,
                                 (g2_xyzz_t*)d.g2_buckets.p, (g2_xyzz_t*)d.g2_head.p, (g2_xyzz_t*)d.g2_tail.p, d.stream));
    ctx->last.kernel_launches += 7;
    return g2_reduce_and_read(ctx, d, p, out_jacobian);
} B200_CATCH

// precompute: -1 = take the context's "precompute" option, 0 = bases only, 1 = window table with the automatic window,
// 8..24 = window table with that window size
static int register_on(b200msm_ctx* ctx, const void* bases, size_t base_stride, size_t x_off, size_t y_off, size_t inf_off,
                       size_t n, const int* dev_indices, int n_dev, b200msm_bases** out, int precompute = -1) {
    if (!ctx || !bases || !out) return fail(B200MSM_EINVAL, "null argument");
    if (n == 0) return fail(B200MSM_EINVAL, "Empty input");
    RET_TRY(check_layout(base_stride, x_off, y_off, inf_off));
    std::vector<int> devs;
    if (!dev_indices || n_dev <= 0) for (int k = 0; k < (int)ctx->devs.size(); k++) devs.push_back(k);
    else devs.assign(dev_indices, dev_indices + n_dev);
    for (int di : devs)
        if (di < 0 || di >= (int)ctx->devs.size()) return fail(B200MSM_EINVAL, "bad dev_index");
    std::lock_guard<std::mutex> lk(ctx->mu);
    b200msm_bases* h = new (std::nothrow) b200msm_bases();
    if (!h) return fail(B200MSM_ENOMEM, "out of host memory");
    h->n = n;
    std::vector<std::pair<size_t, size_t>> ranges;
    shard_ranges(n, devs.size(), &ranges);
    const bool has_inf = inf_off != B200MSM_NO_INF;
    for (size_t k = 0; k < ranges.size(); k++) {
        DevState& d = ctx->devs[devs[k]];
        b200msm_bases::Shard sh;
        sh.dev_index = devs[k];
        sh.begin = ranges[k].first;
        sh.len = ranges[k].second;
        cudaError_t e = cudaSetDevice(d.ordinal);
        const int pre = precompute >= 0 ? precompute : ctx->opt_precompute;
        if (pre) {
            sh.tc = pre >= 8 ? pre : table_window_bits(sh.len);
            sh.tW = num_windows_for(sh.tc, 254);
            if ((uint64_t)sh.tW * sh.len >= (1ull << 31)) { sh.tc = 0; sh.tW = 0; }   // 31-bit entry indices: fall back to plain bases
        }
        const size_t windows = sh.tc ? (size_t)sh.tW : 1;
        if (e == cudaSuccess) e = cudaMalloc(&sh.d_xy, windows * sh.len * 64);
        if (e == cudaSuccess && has_inf) e = cudaMalloc(&sh.d_inf, sh.len);
        h->shards.push_back(sh);
        int rc = e == cudaSuccess ? upload_bases(d, (const uint8_t*)bases + sh.begin * base_stride, base_stride, x_off, y_off, inf_off,
                                                 sh.len, sh.d_xy, sh.d_inf, nullptr, nullptr, host_is_pageable(bases))
                                  : fail(B200MSM_ENOMEM, std::string("register_bases: ") + cudaGetErrorString(e));
        if (rc == B200MSM_OK && sh.tc) {
            k_build_table<<<cdiv(sh.len, 128), 128, 0, d.stream>>>((uint32_t)sh.len, sh.len, sh.tc, sh.tW, (affine_t*)sh.d_xy);
            if (cudaGetLastError() != cudaSuccess) rc = fail(B200MSM_ECUDA, "register_bases: table kernel launch failed");
        }
        if (rc == B200MSM_OK && cudaStreamSynchronize(d.stream) != cudaSuccess) rc = fail(B200MSM_ECUDA, "register_bases: sync failed");
        if (rc != B200MSM_OK) {
            std::string keep = g_err;
            for (auto& s2 : h->shards) {
                cudaSetDevice(ctx->devs[s2.dev_index].ordinal);
                if (s2.d_xy) cudaFree(s2.d_xy);
                if (s2.d_inf) cudaFree(s2.d_inf);
            }
            delete h;
            g_err = keep;
            return rc;
        }
    }
    *out = h;
    return B200MSM_OK;
}

int b200msm_register_bases(b200msm_ctx* ctx, const void* bases, size_t base_stride, size_t x_off, size_t y_off, size_t inf_off,
                           size_t n, b200msm_bases** out) try {
    return register_on(ctx, bases, base_stride, x_off, y_off, inf_off, n, nullptr, 0, out);
} B200_CATCH

int b200msm_register_bases_on(b200msm_ctx* ctx, const void* bases, size_t base_stride, size_t x_off, size_t y_off, size_t inf_off,
                              size_t n, const int* dev_indices, int n_dev, b200msm_bases** out) try {
    return register_on(ctx, bases, base_stride, x_off, y_off, inf_off, n, dev_indices, n_dev, out);
} B200_CATCH

int b200msm_register_bases_ex(b200msm_ctx* ctx, const void* bases, size_t base_stride, size_t x_off, size_t y_off, size_t inf_off,
                              size_t n, const int* dev_indices, int n_dev, int precompute, b200msm_bases** out) try {
    if (precompute != 0 && precompute != 1 && (precompute < 8 || precompute > 24))
        return fail(B200MSM_EINVAL, "precompute must be 0, 1 (auto window) or a window size in [8, 24]");
    return register_on(ctx, bases, base_stride, x_off, y_off, inf_off, n, dev_indices, n_dev, out, precompute);
} B200_CATCH

int b200msm_release_bases(b200msm_ctx* ctx, b200msm_bases* h) try {
    if (!ctx || !h) return fail(B200MSM_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (auto& s : h->shards) {
        cudaSetDevice(ctx->devs[s.dev_index].ordinal);
        cudaStreamSynchronize(ctx->devs[s.dev_index].stream);
        if (s.d_xy) cudaFree(s.d_xy);
        if (s.d_inf) cudaFree(s.d_inf);
    }
    delete h;
    return B200MSM_OK;
} B200_CATCH

size_t b200msm_bases_len(const b200msm_bases* h) { return h ? h->n : 0; }

int b200msm_msm_batch(b200msm_ctx* ctx, int count, const b200msm_bases* const* handles, const void* const* scalars,
                      const size_t* n, uint64_t (*out_jacobian)[12]) try {
    if (!ctx || count <= 0 || !handles || !scalars || !n || !out_jacobian) return fail(B200MSM_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    // slot layout in pinned staging: msm m, shard k -> (m * 16 + k) * 96
    if ((size_t)count * 16 * 96 > ctx->h_pinned_bytes) return fail(B200MSM_EINVAL, "batch too large (max 42 MSMs)");
    ctx->last.kernel_launches = 0;
    // Every (MSM, shard) pair is one item of its device's three-stage pipeline:
    //   copy stream    scalars of item j+1 go up while item j computes                      (two scalar buffers)
    //   main stream    sort + accumulate + fix-up of item j into work set j & 1              (two work sets)
    //   reduce stream  bucket reduce + Horner + result read-back of item j (latency-bound, high priority) under the
    //                  accumulation of item j+1; one set of reduce buffers, the chains are serial on that stream anyway
    struct Item { int m; int dev; size_t begin, len; const b200msm_bases::Shard* sh; Plan p; int slot; };
    std::vector<Item> items;
    std::vector<int> per_dev(ctx->devs.size(), 0);
    std::vector<std::vector<int>> used(count);
    for (int m = 0; m < count; m++) {
        const b200msm_bases* h = handles[m];
        if (!h || !scalars[m]) return fail(B200MSM_EINVAL, "null handle or scalars");
        if (n[m] == 0) return fail(B200MSM_EINVAL, "Empty input");
        if (n[m] > h->n) return fail(B200MSM_EINVAL, "more scalars than registered bases");
        if (h->shards.size() > 16) return fail(B200MSM_EINVAL, "too many shards");
        for (size_t k = 0; k < h->shards.size(); k++) {
            const auto& sh = h->shards[k];
            if (sh.begin >= n[m]) break;
            Item it;
            it.m = m;
            it.dev = sh.dev_index;
            it.begin = sh.begin;
            it.len = std::min(sh.len, n[m] - sh.begin);
            it.sh = &sh;
            DevState& d = ctx->devs[sh.dev_index];
            RET_TRY(make_plan(ctx, d, it.len, &it.p, sh.tc, false, sh.tc ? sh.len : 0));
            if (it.p.ngroups > 1) return fail(B200MSM_EINVAL, "window groups are not supported on the registered / batch path");
            it.slot = per_dev[sh.dev_index]++ & 1;
            items.push_back(it);
        }
    }
    // size every buffer before anything is in flight (growing one later would free memory a queued kernel still uses)
    for (const Item& it : items) {
        DevState& d = ctx->devs[it.dev];
        CU_TRY(cudaSetDevice(d.ordinal));
        RET_TRY(ensure_work(d, it.p, it.slot));
        RET_TRY(ensure_reduce(d, it.p));
        if (it.p.glv) RET_TRY(d.xb.ensure((size_t)it.p.n * 32));
        RET_TRY((it.slot ? d.scalars_alt : d.scalars).ensure(it.len * 32));
    }
    const bool timing = ctx->opt_timing != 0 && count == 1;
    std::vector<int> seen(ctx->devs.size() * 2, 0);   // work set (dev, slot) already used in this call
    for (const Item& it : items) {
        DevState& d = ctx->devs[it.dev];
        CU_TRY(cudaSetDevice(d.ordinal));
        // A lone MSM has no neighbour to hide its scalar upload behind: from 2^21 points per device the upload is cut in
        // two slices (1 : 4) and the first slice's arithmetic covers the second slice's transfer.  Measured
        // (profiles/r01e_registered_slices.jsonl, plain / table handle): 2^20 4.76 -> 4.70 / 4.09 -> 4.05 ms,
        // 2^22 16.0 -> 14.8 / 13.7 -> 12.3, 2^24 56.3 -> 52.0 / 54.1 -> 44.4.
        // Round 2 (slices add up in place): 3 slices growing 2.5x from 2^21 points, 4 growing 2x from 3 * 2^22
        // (profiles/r02_registered_slices_inplace.jsonl: 2^22 13.36 -> 13.24 ms, 2^24 44.7 -> 43.6).
        // With the later slices' sorts on the sort stream (sort_overlap) two slices (1 : 4) pay from 2^20 points
        // (interleaved A/B on one box, profiles/r02_ab_sort_overlap_2e20.jsonl: 2^20 4.32 -> 4.29 ms, with the window table
        // 3.85 -> 3.78; the host-buffer call 4.66 -> 4.49).
        const int S1 = ctx->opt_slices > 0 ? ctx->opt_slices : it.len >= (3u << 22) ? 4 : it.len >= (1u << 21) ? 3 : it.len >= (1u << 20) ? 2 : 1;
        const int ratio1 = ctx->opt_slice_ratio > 0 ? ctx->opt_slice_ratio : S1 >= 4 ? 200 : S1 == 3 ? 250 : 400;
        if (count == 1 && S1 > 1) {
            ResidentBases rb = {it.sh->d_xy, it.sh->d_inf, it.sh->tc, it.sh->len};
            RET_TRY(enqueue_sliced(ctx, d, it.p, S1, nullptr, 0, 0, 0, 0, (const uint8_t*)scalars[it.m] + it.begin * 32, 32, d.out.p,
                                   &ctx->last.kernel_launches, &rb, ratio1));
            CU_TRY(cudaMemcpyAsync(ctx->h_pinned + ((size_t)it.m * 16 + used[it.m].size()) * 96, d.out.p, 96, cudaMemcpyDeviceToHost, d.stream));
            used[it.m].push_back(it.dev);
            continue;
        }
        cudaStream_t s = d.stream, rs = d.stream2, cs = d.stream3;
        cudaEvent_t ev_sc = d.ev_slice[it.slot], ev_front = d.ev_slice[2 + it.slot], ev_red = d.ev_slice[4 + it.slot];
        const bool first_use = !seen[it.dev * 2 + it.slot];
        const bool first_on_dev = !seen[it.dev * 2] && !seen[it.dev * 2 + 1];
        seen[it.dev * 2 + it.slot] = 1;
        void* d_sc = (it.slot ? d.scalars_alt : d.scalars).p;
        if (timing && first_on_dev) CU_TRY(cudaEventRecord(d.ev[EV_START], s));
        if (first_on_dev) {   // the side streams start after whatever the main stream was doing before this call
            CU_TRY(cudaEventRecord(d.ev_acc[7], s));
            CU_TRY(cudaStreamWaitEvent(cs, d.ev_acc[7], 0));
            CU_TRY(cudaStreamWaitEvent(rs, d.ev_acc[7], 0));
        }
        if (!first_use) CU_TRY(cudaStreamWaitEvent(cs, ev_front, 0));   // scalar buffer: free once item j-2 was decomposed
        RET_TRY(h2d(d, d_sc, (const uint8_t*)scalars[it.m] + it.begin * 32, it.len * 32, cs, host_is_pageable(scalars[it.m])));
        CU_TRY(cudaEventRecord(ev_sc, cs));
        CU_TRY(cudaStreamWaitEvent(s, ev_sc, 0));
        if (!first_use) CU_TRY(cudaStreamWaitEvent(s, ev_red, 0));      // work set: free once item j-2 was reduced
        if (timing) CU_TRY(cudaEventRecord(d.ev[EV_H2D], s));
        const WorkView w = view_slice(d, it.slot);
        const Plan& p = it.p;
        RET_TRY(launch_sort(w, p, d_sc, it.sh->d_inf, s, timing ? d.ev[EV_DECOMP] : nullptr));
        if (timing) CU_TRY(cudaEventRecord(d.ev[EV_SORT], s));
        const fq* d_xb = nullptr;
        if (p.glv) {
            k_endo_x<<<cdiv(p.n, 256), 256, 0, s>>>((const affine_t*)it.sh->d_xy, p.n, (fq*)d.xb.p);
            d_xb = (const fq*)d.xb.p;
        }
        CU_TRY(cudaMemsetAsync((uint32_t*)w.wtotal + 64, 0, 4 * 48, s));
        RET_TRY(launch_accumulate(d, w, p, it.sh->d_xy, d_xb, 0, p.Wb, s));
        if (timing) CU_TRY(cudaEventRecord(d.ev[EV_ACC], s));
        RET_TRY(launch_fixup(d, w, p, 0, p.Wb, 0, s));
        CU_TRY(cudaEventRecord(ev_front, s));
        CU_TRY(cudaStreamWaitEvent(rs, ev_front, 0));
        int nlaunch = p.glv ? 8 : 7;
        RET_TRY(launch_reduce(d, p, w.buckets, 0, p.Wb, true, rs, d.out.p, &nlaunch));
        CU_TRY(cudaMemcpyAsync(ctx->h_pinned + ((size_t)it.m * 16 + used[it.m].size()) * 96, d.out.p, 96, cudaMemcpyDeviceToHost, rs));
        CU_TRY(cudaEventRecord(ev_red, rs));
        ctx->last.kernel_launches += nlaunch;
        used[it.m].push_back(it.dev);
    }
    // the main stream (the one finish_and_combine synchronises) rejoins the reduce stream
    for (size_t dv = 0; dv < ctx->devs.size(); dv++) {
        DevState& d = ctx->devs[dv];
        if (!seen[dv * 2] && !seen[dv * 2 + 1]) continue;
        CU_TRY(cudaSetDevice(d.ordinal));
        for (int slot = 0; slot < 2; slot++)
            if (seen[dv * 2 + slot]) CU_TRY(cudaStreamWaitEvent(d.stream, d.ev_slice[4 + slot], 0));
        if (timing) CU_TRY(cudaEventRecord(d.ev[EV_RED], d.stream));
    }
    for (int m = 0; m < count; m++)
        RET_TRY(finish_and_combine(ctx, used[m], ctx->h_pinned + (size_t)m * 16 * 96, out_jacobian[m], &ctx->last.kernel_launches));
    const Plan plan0 = items[0].p;
    ctx->last_plan = plan0;
    DevState& d0 = ctx->devs[items[0].dev];
    CU_TRY(cudaSetDevice(d0.ordinal));
    if (count == 1) return collect_timings(ctx, d0, plan0, true);
    ctx->last.window_bits = plan0.c;
    ctx->last.num_windows = plan0.W;
    return B200MSM_OK;
} B200_CATCH

int b200msm_msm_registered(b200msm_ctx* ctx, const b200msm_bases* h, const void* scalars, size_t scalar_stride, size_t n,
                           uint64_t out_jacobian[12]) try {
    if (scalar_stride != 32) return fail(B200MSM_EINVAL, "registered path requires scalar_stride == 32");
    const b200msm_bases* hs[1] = {h};
    const void* sc[1] = {scalars};
    size_t ns[1] = {n};
    return b200msm_msm_batch(ctx, 1, hs, sc, ns, (uint64_t(*)[12])out_jacobian);
} B200_CATCH

// ------------------------------------------------------------------------------------------- test kit
int b200msm_testkit_generate(b200msm_ctx* ctx, int dev_index, uint64_t seed, size_t n, void* d_bases, void* d_scalars,
                             uint8_t* h_t1, uint8_t* h_t2) try {
    if (!ctx || n == 0 || n >= (1ull << 31)) return fail(B200MSM_EINVAL, "bad argument");
    if (dev_index < 0 || dev_index >= (int)ctx->devs.size()) return fail(B200MSM_EINVAL, "bad dev_index");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState& d = ctx->devs[dev_index];
    CU_TRY(cudaSetDevice(d.ordinal));
    if (d_scalars) {
        k_tk_gen_scalars<<<cdiv(n, 256), 256, 0, d.stream>>>(seed ^ 0x5ca1a75ull, (uint32_t)n, (uint64_t*)d_scalars);
        CU_TRY(cudaGetLastError());
    }
    if (d_bases) {
        size_t n2 = (n + 4095) / 4096;
        std::vector<uint64_t> dl((4096 + n2) * 4);
        for (size_t k = 0; k < 4096; k++) tk_random_below_r(seed ^ 0x7ab1e001ull, k, &dl[k * 4]);
        for (size_t k = 0; k < n2; k++) tk_random_below_r(seed ^ 0x7ab1e002ull, k, &dl[(4096 + k) * 4]);
        if (h_t1) std::memcpy(h_t1, dl.data(), 4096 * 32);
        if (h_t2) std::memcpy(h_t2, dl.data() + 4096 * 4, n2 * 32);
        void *d_dl = nullptr, *d_tab = nullptr;
        CU_TRY(cudaMalloc(&d_dl, dl.size() * 8));
        cudaError_t e = cudaMalloc(&d_tab, (4096 + n2) * 64);
        if (e != cudaSuccess) {
            cudaFree(d_dl);
            return fail(B200MSM_ENOMEM, "testkit table allocation failed");
        }
        cudaMemcpyAsync(d_dl, dl.data(), dl.size() * 8, cudaMemcpyHostToDevice, d.stream);
        k_tk_gen_table<<<cdiv(4096 + n2, 64), 64, 0, d.stream>>>((const uint32_t*)d_dl, (uint32_t)(4096 + n2), (affine_t*)d_tab);
        k_tk_gen_bases<<<cdiv(n, 128), 128, 0, d.stream>>>((const affine_t*)d_tab, (const affine_t*)d_tab + 4096, (uint32_t)n,
                                                         (affine_t*)d_bases);
        e = cudaStreamSynchronize(d.stream);
        cudaFree(d_dl);
        cudaFree(d_tab);
        if (e != cudaSuccess) return fail(B200MSM_ECUDA, std::string("testkit_generate: ") + cudaGetErrorString(e));
    }
    CU_TRY(cudaStreamSynchronize(d.stream));
    return B200MSM_OK;
} B200_CATCH

// SM blocker: occupy `n_sms` SMs of the device with sleeping CTAs until b200msm_testkit_release_sms (or max_seconds).
int b200msm_testkit_occupy_sms(b200msm_ctx* ctx, int dev_index, int n_sms, double max_seconds) try {
    if (!ctx || dev_index < 0 || dev_index >= (int)ctx->devs.size()) return fail(B200MSM_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState& d = ctx->devs[dev_index];
    if (n_sms < 1 || n_sms >= d.hw_sm_count) return fail(B200MSM_EINVAL, "n_sms must be in [1, SM count)");
    if (max_seconds <= 0 || max_seconds > 120) return fail(B200MSM_EINVAL, "max_seconds must be in (0, 120]");
    if (d.occ_flag && *d.occ_flag == 0) return fail(B200MSM_EINVAL, "SMs are already occupied");
    CU_TRY(cudaSetDevice(d.ordinal));
    if (!d.occ_flag) {
        CU_TRY(cudaHostAlloc((void**)&d.occ_flag, 64, cudaHostAllocMapped));
        CU_TRY(cudaMalloc((void**)&d.occ_started, 4));
        CU_TRY(cudaStreamCreateWithFlags(&d.occ_stream, cudaStreamNonBlocking));
    }
    *d.occ_flag = 0;
    CU_TRY(cudaMemsetAsync(d.occ_started, 0, 4, d.occ_stream));
    int* dflag = nullptr;
    CU_TRY(cudaHostGetDevicePointer((void**)&dflag, d.occ_flag, 0));
    int khz = 1965000;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, d.ordinal);
    k_tk_occupy<<<n_sms, 1024, 0, d.occ_stream>>>(dflag, d.occ_started, (long long)(max_seconds * 1e3 * khz));
    CU_TRY(cudaGetLastError());
    // wait until every blocker CTA is resident (a CTA that cannot start yet would start later, on an SM the MSM has freed)
    for (int spin = 0; spin < 20000; spin++) {
        unsigned int started = 0;
        CU_TRY(cudaMemcpyAsync(&started, d.occ_started, 4, cudaMemcpyDeviceToHost, d.stream2));
        CU_TRY(cudaStreamSynchronize(d.stream2));
        if (started == (unsigned)n_sms) return B200MSM_OK;
        std::this_thread::sleep_for(std::chrono::microseconds(100));
    }
    *d.occ_flag = 1;
    return fail(B200MSM_ECUDA, "blocker CTAs did not all become resident");
} B200_CATCH

int b200msm_testkit_release_sms(b200msm_ctx* ctx, int dev_index) try {
    if (!ctx || dev_index < 0 || dev_index >= (int)ctx->devs.size()) return fail(B200MSM_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState& d = ctx->devs[dev_index];
    if (!d.occ_flag) return B200MSM_OK;
    *d.occ_flag = 1;
    CU_TRY(cudaSetDevice(d.ordinal));
    CU_TRY(cudaStreamSynchronize(d.occ_stream));
    return B200MSM_OK;
} B200_CATCH

// Plain IMAD.WIDE.U32 issue rate (no carries), 8 independent chains per thread: the measured
// integer-multiply roofline denominator on THIS device at ITS current clocks.
int b200msm_testkit_imad_peak(b200msm_ctx* ctx, int dev_index, double* macs_per_s) try {
    if (!ctx || !macs_per_s || dev_index < 0 || dev_index >= (int)ctx->devs.size()) return fail(B200MSM_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState& d = ctx->devs[dev_index];
    CU_TRY(cudaSetDevice(d.ordinal));
    const int blocks = d.sm_count * 4, threads = 256, iters = 8192;
    void* out = nullptr;
    CU_TRY(cudaMalloc(&out, (size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0, d.stream);
        k_tk_imad_wide<<<blocks, threads, 0, d.stream>>>((uint64_t*)out, iters, 3, 7);
        cudaEventRecord(e1, d.stream);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { cudaFree(out); return fail(B200MSM_ECUDA, cudaGetErrorString(e)); }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 1 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *macs_per_s = 32.0 * iters * blocks * threads / (best * 1e-3);
    return B200MSM_OK;
} B200_CATCH

int b200msm_testkit_op(b200msm_ctx* ctx, int op, const void* a, const void* b, void* out, size_t count) try {
    if (!ctx || !a || !out || count == 0) return fail(B200MSM_EINVAL, "bad argument");
    size_t sa, sb, so;
    switch (op) {
        case 0: case 1: case 2: sa = 32; sb = 32; so = 32; break;
        case 3: case 4: case 5: case 6: case 7: sa = 32; sb = 0; so = 32; break;
        case 8: sa = 64; sb = 64; so = 32; break;
        case 14: sa = 96; sb = 0; so = 96; break;
        case 15: sa = 64; sb = 8; so = 128; break;
        case 10: sa = 128; sb = 64; so = 128; break;
        case 11: sa = 128; sb = 128; so = 128; break;
        case 12: sa = 128; sb = 0; so = 128; break;
        case 13: sa = 128; sb = 0; so = 96; break;
        case 20: sa = 32; sb = 0; so = 32; break;
        case 30: sa = 64; sb = 64; so = 64; break;
        case 31: sa = 64; sb = 0; so = 64; break;
        case 32: sa = 256; sb = 128; so = 256; break;
        case 33: sa = 256; sb = 256; so = 256; break;
        case 34: sa = 256; sb = 0; so = 256; break;
        default: return fail(B200MSM_EINVAL, "unknown op");
    }
    if (sb && !b) return fail(B200MSM_EINVAL, "operand b required");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState& d = ctx->devs[0];
    CU_TRY(cudaSetDevice(d.ordinal));
    void *da = nullptr, *db = nullptr, *dout = nullptr;
    {   // all three or none: an allocation failure must not leak the earlier buffers
        cudaError_t ea = cudaMalloc(&da, sa * count);
        if (ea == cudaSuccess && sb) ea = cudaMalloc(&db, sb * count);
        if (ea == cudaSuccess) ea = cudaMalloc(&dout, so * count);
        if (ea != cudaSuccess) {
            if (da) cudaFree(da);
            if (db) cudaFree(db);
            return fail(B200MSM_ENOMEM, std::string("testkit_op: ") + cudaGetErrorString(ea));
        }
    }
    cudaMemcpyAsync(da, a, sa * count, cudaMemcpyHostToDevice, d.stream);
    if (sb) cudaMemcpyAsync(db, b, sb * count, cudaMemcpyHostToDevice, d.stream);
    if (op >= 30)
        k_g2_tk_op<<<cdiv(count, 64), 64, 0, d.stream>>>(op, (const uint8_t*)da, (const uint8_t*)db, (uint8_t*)dout, (uint32_t)count);
    else
        k_tk_op<<<cdiv(count, 128), 128, 0, d.stream>>>(op, (const uint8_t*)da, (const uint8_t*)db, (uint8_t*)dout, (uint32_t)count);
    cudaMemcpyAsync(out, dout, so * count, cudaMemcpyDeviceToHost, d.stream);
    cudaError_t e = cudaStreamSynchronize(d.stream);
    cudaFree(da);
    if (db) cudaFree(db);
    cudaFree(dout);
    if (e != cudaSuccess) return fail(B200MSM_ECUDA, std::string("testkit_op: ") + cudaGetErrorString(e));
    return B200MSM_OK;
} B200_CATCH

// Public math-library entry (include/b200math.h): the same element-wise dispatcher under its documented operation names.
int b200math_apply(b200msm_ctx* ctx, b200math_op op, const void* a, const void* b, void* out, size_t count) {
    return b200msm_testkit_op(ctx, (int)op, a, b, out, count);
}

int b200msm_testkit_sort(b200msm_ctx* ctx, const void* scalars, size_t n, int window_bits, uint32_t* ends, uint32_t* entries,
                         uint64_t* n_entries, int* num_windows, uint64_t* n_pseudo) try {
    if (!ctx || !scalars || !ends || !entries || !n_entries || !num_windows || !n_pseudo || n == 0)
        return fail(B200MSM_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState& d = ctx->devs[0];
    CU_TRY(cudaSetDevice(d.ordinal));
    int saved = ctx->opt_window_bits;
    ctx->opt_window_bits = window_bits;
    Plan p;
    int rc = make_plan(ctx, d, n, &p);
    ctx->opt_window_bits = saved;
    RET_TRY(rc);
    RET_TRY(ensure_workspace(d, p));
    void* d_scalars = nullptr;
    RET_TRY(upload_scalars(d, (const uint8_t*)scalars, 32, n, &d_scalars, nullptr));
    cudaStream_t s = d.stream;
    RET_TRY(launch_sort(view_main(d), p, d_scalars, nullptr, s, nullptr));
    CU_TRY(cudaMemcpyAsync(ends, d.ends.p, (size_t)p.G * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    uint32_t total = ends[p.G - 1];
    *n_entries = total;
    *num_windows = p.W;
    *n_pseudo = p.n_eff;
    if (total) CU_TRY(cudaMemcpy(entries, d.entries.p, (size_t)total * 4, cudaMemcpyDeviceToHost));
    return B200MSM_OK;
} B200_CATCH

// Run the full pipeline on host inputs and copy back the per-window sums G_w (XYZZ, 16 u64 each)
// that feed the Horner step: the stage-4 probe (reference: tests/cuzk/pbpr.rs:26-247).
int b200msm_testkit_window_sums(b200msm_ctx* ctx, const void* bases64, const void* scalars, size_t n, int window_bits,
                                uint64_t* out_wsum, int* num_windows) try {
    if (!ctx || !bases64 || !scalars || !out_wsum || !num_windows || n == 0) return fail(B200MSM_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState& d = ctx->devs[0];
    CU_TRY(cudaSetDevice(d.ordinal));
    int saved = ctx->opt_window_bits;
    ctx->opt_window_bits = window_bits;
    Plan p;
    int rc = make_plan(ctx, d, n, &p);
    ctx->opt_window_bits = saved;
    RET_TRY(rc);
    RET_TRY(ensure_workspace(d, p));
    RET_TRY(d.bases.ensure(n * 64));
    void* d_scalars = nullptr;
    RET_TRY(upload_scalars(d, (const uint8_t*)scalars, 32, n, &d_scalars, nullptr));
    CU_TRY(cudaMemcpyAsync(d.bases.p, bases64, n * 64, cudaMemcpyHostToDevice, d.stream));
    RET_TRY(enqueue_msm(ctx, d, p, d.bases.p, nullptr, d_scalars, d.out.p, nullptr));
    xyzz_t* wsum = (xyzz_t*)d.wpart.p + (size_t)p.W * p.bpw * 2;
    if (p.rowcol) {   // the window sum is kept as (row term, column term): add them for the probe
        const xyzz_t* wcols = (const xyzz_t*)d.redbuf.p + (size_t)p.W * ((size_t)(1u << p.rc_ra) + (1u << p.rc_cb)) + 2 * (size_t)p.W;
        k_add_into<<<cdiv(p.W, 32), 32, 0, d.stream>>>(wsum, wcols, p.W);
    }
    CU_TRY(cudaMemcpyAsync(out_wsum, wsum, (size_t)p.W * sizeof(xyzz_t), cudaMemcpyDeviceToHost, d.stream));
    CU_TRY(cudaStreamSynchronize(d.stream));
    *num_windows = p.W;
    return B200MSM_OK;
} B200_CATCH

// G2 stage-4 probe: run the G2 MSM on host inputs (bases: n x 128 B x.c0|x.c1|y.c0|y.c1, scalars: n x 32 B) with the given
// window size and return the per-window sums G_w = sum_m m * bucket[w][m] as XYZZ over Fq2 (32 u64 each): what
// k_g2_bucket_reduce + k_g2_window_finish (g2_block_weighted_sum) produce, before the Horner chain.
int b200msm_testkit_g2_window_sums(b200msm_ctx* ctx, const void* bases128, const void* scalars, size_t n, int window_bits,
                                   uint64_t* out_wsum, int* num_windows) try {
    if (!ctx || !bases128 || !scalars || !out_wsum || !num_windows || n == 0) return fail(B200MSM_EINVAL, "bad argument");
    int saved;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        saved = ctx->opt_window_bits;
        ctx->opt_window_bits = window_bits;
    }
    uint64_t res[24];
    int rc = b200msm_bn254_g2_msm(ctx, bases128, 128, 0, 64, B200MSM_NO_INF, scalars, 32, n, res);
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState& d = ctx->devs[0];
    Plan p;
    if (rc == B200MSM_OK) rc = make_plan(ctx, d, n, &p);
    ctx->opt_window_bits = saved;
    RET_TRY(rc);
    CU_TRY(cudaSetDevice(d.ordinal));
    uint32_t lb, bpw;
    g2_reduce_shape(p, &lb, &bpw);
    const g2_xyzz_t* wsum = (const g2_xyzz_t*)d.g2_wpart.p + (size_t)p.Wb * bpw * 2;
    CU_TRY(cudaMemcpy(out_wsum, wsum, (size_t)p.Wb * sizeof(g2_xyzz_t), cudaMemcpyDeviceToHost));
    *num_windows = p.Wb;
    return B200MSM_OK;
} B200_CATCH

// Host-only probes (no device, no context): the slice plan and the parallel staging copy.
int b200msm_testkit_slice_plan(size_t n, int slices, int ratio_pct, size_t* begins, size_t* lens, int* count) try {
    if (!begins || !lens || !count || n == 0 || slices < 1 || slices > MAX_SLICES || ratio_pct < 100 || ratio_pct > 400)
        return fail(B200MSM_EINVAL, "bad argument");
    std::vector<std::pair<size_t, size_t>> sl;
    slice_ranges(n, slices, ratio_pct / 100.0, &sl);
    *count = (int)sl.size();
    for (size_t k = 0; k < sl.size(); k++) {
        begins[k] = sl[k].first;
        lens[k] = sl[k].second;
    }
    return B200MSM_OK;
} B200_CATCH

int b200msm_testkit_parallel_copy(void* dst, const void* src, size_t bytes, int threads) try {
    if (!dst || !src || threads < 1 || threads > 32) return fail(B200MSM_EINVAL, "bad argument");
    CopyPool pool(threads - 1);
    pool.copy(dst, src, bytes);
    pool.copy(dst, src, bytes);   // a pool serves many copies: the second one must work too
    return B200MSM_OK;
} B200_CATCH

int b200msm_testkit_table(b200msm_ctx* ctx, const b200msm_bases* h, int window, size_t count, void* out_xy64, int* window_bits,
                          int* num_windows) try {
    if (!ctx || !h || !out_xy64 || !window_bits || !num_windows || h->shards.empty()) return fail(B200MSM_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    const auto& sh = h->shards[0];
    if (!sh.tc) return fail(B200MSM_EINVAL, "handle has no precomputed table");
    if (window < 0 || window >= sh.tW || count > sh.len) return fail(B200MSM_EINVAL, "window / count out of range");
    DevState& d = ctx->devs[sh.dev_index];
    CU_TRY(cudaSetDevice(d.ordinal));
    CU_TRY(cudaStreamSynchronize(d.stream));
    CU_TRY(cudaMemcpy(out_xy64, (const uint8_t*)sh.d_xy + (size_t)window * sh.len * 64, count * 64, cudaMemcpyDeviceToHost));
    *window_bits = sh.tc;
    *num_windows = sh.tW;
    return B200MSM_OK;
} B200_CATCH

// ------------------------------------------------------------------------------------------- instance files
// Decode `count` compressed G1 points (32 B each, host) into 64-byte Montgomery x||y records (host) on the GPU.
int b200msm_decompress_g1(b200msm_ctx* ctx, const void* compressed, size_t count, void* out_xy64, uint64_t* n_invalid) try {
    if (!ctx || !compressed || !out_xy64 || count == 0 || count >= (1ull << 31)) return fail(B200MSM_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState& d = ctx->devs[0];
    CU_TRY(cudaSetDevice(d.ordinal));
    RET_TRY(d.raw.ensure(count * 32 + 16));
    RET_TRY(d.bases.ensure(count * 64));
    unsigned long long* d_bad = (unsigned long long*)((uint8_t*)d.raw.p + ((count * 32 + 7) & ~(size_t)7));
    CU_TRY(cudaMemcpyAsync(d.raw.p, compressed, count * 32, cudaMemcpyHostToDevice, d.stream));
    CU_TRY(cudaMemsetAsync(d_bad, 0, 8, d.stream));
    k_decompress_g1<<<cdiv(count, 128), 128, 0, d.stream>>>((const uint8_t*)d.raw.p, (uint32_t)count, (affine_t*)d.bases.p, d_bad);
    CU_TRY(cudaGetLastError());
    unsigned long long bad = 0;
    CU_TRY(cudaMemcpyAsync(out_xy64, d.bases.p, count * 64, cudaMemcpyDeviceToHost, d.stream));
    CU_TRY(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, d.stream));
    CU_TRY(cudaStreamSynchronize(d.stream));
    if (n_invalid) *n_invalid = bad;
    return B200MSM_OK;
} B200_CATCH

// canonical 32-byte little-endian scalars (the `scalars` file) -> Fr Montgomery words (`&[Fr]` memory).
int b200msm_fr_to_montgomery(b200msm_ctx* ctx, const void* canonical, size_t count, void* out) try {
    if (!ctx || !canonical || !out || count == 0 || count >= (1ull << 31)) return fail(B200MSM_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevState& d = ctx->devs[0];
    CU_TRY(cudaSetDevice(d.ordinal));
    RET_TRY(d.scalars_raw.ensure(count * 32));
    RET_TRY(d.scalars.ensure(count * 32));
    CU_TRY(cudaMemcpyAsync(d.scalars_raw.p, canonical, count * 32, cudaMemcpyHostToDevice, d.stream));
    k_fr_to_mont<<<cdiv(count, 256), 256, 0, d.stream>>>((const uint4*)d.scalars_raw.p, (uint32_t)count, (uint4*)d.scalars.p);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(out, d.scalars.p, count * 32, cudaMemcpyDeviceToHost, d.stream));
    CU_TRY(cudaStreamSynchronize(d.stream));
    return B200MSM_OK;
} B200_CATCH

}  // extern "C"
