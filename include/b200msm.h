/* b200msm.h -- C ABI of libb200msm.so: BN254 G1 variable-base MSM on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for ONE hot path of zkmopro/gpu-acceleration (`mopro-msm`):
 *
 *     pub fn metal_variable_base_msm(bases: &[G1Affine], scalars: &[Fr])
 *         -> Result<G1Projective, Box<dyn Error>>
 *     /root/reference/mopro-msm/src/msm/metal_msm/metal_msm.rs:642-695
 *     (re-exported at /root/reference/mopro-msm/src/msm/metal_msm/mod.rs:7)
 *
 * A Rust shim `cuda_variable_base_msm` with the same signature binds these entry points
 * (gpu-acceleration_b200/rust/src/cuda_msm.rs; binding shown in INTEGRATION.md).  All
 * pointers are plain; no C++/torch types cross the boundary; nothing throws or aborts:
 * every function returns 0 on success or a negative B200MSM_E* code, with a message
 * retrievable through b200msm_last_error().
 *
 * Memory conventions (arkworks 0.4, ark-ff 0.4.1 MontBackend<_,4>):
 *   - a field element is 4 little-endian u64 holding a*R mod m, R = 2^256 (Montgomery form);
 *   - `bases`   : n records of `base_stride` bytes; Fq x at +x_off, Fq y at +y_off, and a
 *                 1-byte `infinity` flag at +inf_off (pass B200MSM_NO_INF if the record has none);
 *                 the shim passes size_of::<G1Affine>() and offset_of! values, so the
 *                 non-repr(C) Rust layout is handled by construction;
 *   - `scalars` : n records of `scalar_stride` bytes, Fr at offset 0;
 *   - result    : 12 u64 = Jacobian (X, Y, Z) in Montgomery form, i.e. the in-memory
 *                 content of `G1Projective { x, y, z }`; infinity is (R, R, 0).
 *                 Coordinates are fully reduced (< p).  Only the group element is specified,
 *                 not the representative: compare with `==` / after normalisation, as the
 *                 reference's tests do (tests/cuzk/e2e.rs:58-61).
 */
#ifndef B200MSM_H
#define B200MSM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MSM_OK 0
#define B200MSM_EINVAL (-1)   /* bad argument (null pointer, n == 0 -> "Empty input", bad option) */
#define B200MSM_ECUDA (-2)    /* CUDA runtime error (message has the cudaError string)            */
#define B200MSM_ENOMEM (-3)   /* device or pinned-host allocation failed                          */
#define B200MSM_ENODEV (-4)   /* no usable sm_100 device / bad device ordinal                     */
#define B200MSM_NO_INF ((size_t)-1)

typedef struct b200msm_ctx b200msm_ctx;     /* opaque: devices, streams, pooled buffers          */
typedef struct b200msm_bases b200msm_bases; /* opaque: device-resident (registered) base set     */

/* Stage timings of the most recent MSM on device 0 of the context, CUDA-event milliseconds. */
typedef struct b200msm_timings {
    float h2d_ms;        /* host->device copies + repack (0 for resident inputs) */
    float decompose_ms;  /* K1: Fr Montgomery->canonical, signed digits, bucket histogram */
    float sort_ms;       /* K2: scan + scatter (CSR build)                              */
    float accumulate_ms; /* K3: bucket accumulation (+ boundary fix-up)                 */
    float reduce_ms;     /* K4+K5: bucket running-sum reduce + window combine           */
    float total_ms;      /* first kernel -> result point ready on device                */
    int window_bits;     /* c actually used                                            */
    int num_windows;
    unsigned long long entries; /* non-zero digits accumulated                          */
    unsigned long long kernel_launches;
} b200msm_timings;

/* ---- context ------------------------------------------------------------------------------
 * Replaces MetalMSMPipeline::with_default_config()/ShaderManager::new, which the reference
 * rebuilds on EVERY call (metal_msm.rs:693, host/shader_manager.rs:100-135).  A context is
 * persistent, owns one stream + buffer pool per device and is internally serialised.
 * devices == NULL / n_devices == 0 selects device 0.                                        */
int b200msm_create(b200msm_ctx** out, const int* devices, int n_devices);
void b200msm_destroy(b200msm_ctx* ctx);
/* Thread-local message for the last failure on this thread (ctx may be NULL). */
const char* b200msm_last_error(const b200msm_ctx* ctx);
/* Hash of the sources this library was compiled from (gpu-acceleration_b200/build_id.py computes the same value from a
 * checkout): lets a host binding refuse a stale binary.                                      */
const char* b200msm_build_id(void);
int b200msm_device_count(const b200msm_ctx* ctx);

/* Options (replace the hard-coded size->(window_size, scale_factor) tables, metal_msm.rs:661-691):
 *   "window_bits"   0 = auto-tune per (n, SM count) [default]; 4..24 forces c
 *   "chunk"         0 = auto; else entries per accumulate thread
 *   "glv"           -1 = auto [default: on for n <= 2^22 per device], 1 = always split scalars with the BN254
 *                   endomorphism (127-bit half-scalars over 2n pseudo-points), 0 = plain 254-bit windows
 *   "coop_reduce"   -1 = auto [default], 1 = bucket reduce on the lane-parallel cooperative engine, 0 = thread-per-segment kernels
 *   "groups"        0 = auto; else number of window groups pipelined between accumulate and reduce (1..8)
 *   "reduce_log2"   -1 = auto; else log2 of the bucket magnitudes each bucket-reduce thread owns
 *   "precompute"    applies to LATER b200msm_register_bases calls: 0 = keep the bases only [default]; 1 = also build the
 *                   window table 2^(c*w) * P_i (w < ceil(254/c), c chosen per size; W x 64 B per point of HBM) so that MSMs
 *                   over the handle use one bucket set and no Horner step; 8..24 = the same with that window size
 *   "slices"        0 = auto [default], 1 = off, 2..8 = slices of the point range the host-buffer call uploads and
 *                   accumulates one after the other (transfer of slice k+1 under the arithmetic of slice k)
 *   "copy_threads"  host threads (caller included) that stage PAGEABLE input memory into the pinned upload ring
 *                   [default min(6, host cores / visible GPUs)]
 *   "ranked_sort"   sort engine of K1+K2: -1 = auto [default]: shared-memory radix partition (2) from 2^22 digits, ranked (1) below;
 *                   0 = cursor atomics in the scatter, 1 = ranks from the histogram pass + atomic-free scatter,
 *                   2 = partitioned sort forced (EINVAL when the window is too wide for its counters: c >= 23)
 *   "rowcol_reduce" 1 = bucket reduce (K4) through row / column sums of the bucket matrix + one cooperative level (measured neutral
 *                   on B200); -1 / 0 = the recursive cooperative levels [default]
 *   "fix_chunks"    chunk-boundary fix-up: -1 = auto [default]: one thread per chunk (full warps) from 2^20 digits, one per bucket below; 0 / 1 forced
 *   "slice_ratio"   percent, length of slice k+1 / slice k; 0 = auto [default]: 160 up to 3 slices, 140 for 4-5, 125 above (100 = equal)
 *   "adaptive_slices" host-buffer call: -1 / 1 = on [default]: when the uploads of the previous call on the device took longer than
 *                   its arithmetic does (several processes sharing the host's H2D bandwidth), the next one uses 4 (8 from 2^21
 *                   points) EQUAL slices, so that little work is left after the last byte has landed; off again below
 *                   0.85 x; 0 = off
 *   "sort_overlap"  sliced host call: K1 + K2 of slice k+1 on a high-priority side stream under the accumulation of slice k;
 *                   -1 = auto [default]: below 3 * 2^20 points (where the sort is a latency-bound kernel chain), 0 / 1 forced
 *   "batch_affine"  -1 = auto [default], 1 = bucket accumulation with batched affine additions (chunk-local tree rounds sharing one
 *                   safegcd inversion per lane and round), 0 = XYZZ chunks; "ba_chunk" 0 = auto / 32..512 entries per thread,
 *                   "ba_min_pairs" 0 = auto / smallest round worth an inversion (measured slower than XYZZ on B200: DESIGN.md)
 *   "sm_count"      0 = the device's SM count [default]; else the SMs the window policy and the persistent grids assume
 *   "timing"        1 = record per-stage CUDA-event timings (adds event records only)      */
int b200msm_set_option(b200msm_ctx* ctx, const char* key, long long value);
int b200msm_last_timings(const b200msm_ctx* ctx, b200msm_timings* out);
/* Sort engine (K1 + K2) the most recent MSM of the context used: 0 cursor atomics, 1 ranked, 2 partitioned (see "ranked_sort"). */
int b200msm_last_sort_engine(const b200msm_ctx* ctx);
/* The window size the auto-tuner picks for n points per device (cuZK cost model corrected by
 * measurement; replaces utils/window_size_optimizer.rs:57-76). */
int b200msm_auto_window_bits(const b200msm_ctx* ctx, size_t n);

/* ---- the drop-in call ---------------------------------------------------------------------
 * Host buffers in, one point out; blocking; borrows the inputs only for the duration of the
 * call (pinned host memory lets the transfers overlap the arithmetic; pageable memory works, slower).  Shards [0,n) by contiguous point range over the context's devices and combines the
 * per-device partial sums on device 0.  n == 0 -> B200MSM_EINVAL ("Empty input",
 * metal_msm.rs:647-649).  Length-mismatch truncation (metal_msm.rs:652-656) is the shim's job
 * (it passes min(len)).                                                                     */
int b200msm_bn254_g1_msm(b200msm_ctx* ctx,
                         const void* bases, size_t base_stride, size_t x_off, size_t y_off, size_t inf_off,
                         const void* scalars, size_t scalar_stride,
                         size_t n, uint64_t out_jacobian[12]);

/* ---- BN254 G2 (nothing to replace: the reference has no G2 path) -------------------------------
 * The B2 multi-scalar multiplication of a Groth16 prover: sum_i scalars[i] * bases[i] over G2Affine points.
 * bases: arkworks `G2Affine {x: Fq2, y: Fq2, infinity}` records; Fq2 = {c0, c1}, each 4 LE u64 in Montgomery form, so
 * x occupies 64 bytes at x_off and y 64 bytes at y_off (c1 at +32).  scalars: as for G1.  out: G2Projective memory,
 * Jacobian (X.c0, X.c1, Y.c0, Y.c1, Z.c0, Z.c1), 24 u64.  Sharded by point range over the context's devices (shards of
 * >= 2^12 points), partials added on the first device.
 * PRECONDITION: every base lies in the order-r subgroup G2 (what arkworks' checked deserialisation guarantees; proving
 * keys do).  With the scalar split on ("glv" -1 / 1) the engine uses phi(P) = (beta^2 x, y) = lambda P, which holds on
 * that subgroup only -- the twist has a large cofactor.  For unchecked on-curve points outside G2 set "glv" to 0.       */
int b200msm_bn254_g2_msm(b200msm_ctx* ctx,
                         const void* bases, size_t base_stride, size_t x_off, size_t y_off, size_t inf_off,
                         const void* scalars, size_t scalar_stride,
                         size_t n, uint64_t out_jacobian[24]);

/* Registered G2 base sets (B2 of a proving key): bases uploaded once to the context's first device; precompute = 0 keeps
 * the bases only, 1 also builds the window table 2^(c*w) * P_i with the automatic window size (W x 128 B per point),
 * 8..24 the same with that window size.  MSMs over a handle move only the scalars.                                  */
typedef struct b200msm_g2_bases b200msm_g2_bases;
int b200msm_g2_register_bases(b200msm_ctx* ctx,
                              const void* bases, size_t base_stride, size_t x_off, size_t y_off, size_t inf_off,
                              size_t n, int precompute, b200msm_g2_bases** out);
int b200msm_g2_release_bases(b200msm_ctx* ctx, b200msm_g2_bases* h);
int b200msm_g2_msm_registered(b200msm_ctx* ctx, const b200msm_g2_bases* h, const void* scalars, size_t scalar_stride,
                              size_t n, uint64_t out_jacobian[24]);

/* ---- registered (device-resident) bases: SURVEY §8(f) rank 1 / BASELINE config #5 --------- */
int b200msm_register_bases(b200msm_ctx* ctx,
                           const void* bases, size_t base_stride, size_t x_off, size_t y_off, size_t inf_off,
                           size_t n, b200msm_bases** out);
/* Same, sharded over a subset of the context's devices (indices into the context's device
 * list), so that a batch of MSMs can be spread over disjoint GPU groups.                    */
int b200msm_register_bases_on(b200msm_ctx* ctx,
                              const void* bases, size_t base_stride, size_t x_off, size_t y_off, size_t inf_off,
                              size_t n, const int* dev_indices, int n_dev, b200msm_bases** out);
/* The same with the table choice in the call instead of the "precompute" option (safe when several threads share a
 * context): precompute = 0 bases only, 1 window table with the automatic window size, 8..24 that window size.
 * dev_indices == NULL / n_dev == 0 = all devices of the context.                                                  */
int b200msm_register_bases_ex(b200msm_ctx* ctx,
                              const void* bases, size_t base_stride, size_t x_off, size_t y_off, size_t inf_off,
                              size_t n, const int* dev_indices, int n_dev, int precompute, b200msm_bases** out);
int b200msm_release_bases(b200msm_ctx* ctx, b200msm_bases* h);
size_t b200msm_bases_len(const b200msm_bases* h);
/* scalars: host memory, n <= registered length (uses the first n bases). */
int b200msm_msm_registered(b200msm_ctx* ctx, const b200msm_bases* h,
                           const void* scalars, size_t scalar_stride, size_t n,
                           uint64_t out_jacobian[12]);
/* count independent MSMs submitted together; each runs on its handle's devices, all overlap. */
int b200msm_msm_batch(b200msm_ctx* ctx, int count, const b200msm_bases* const* handles,
                      const void* const* scalars, const size_t* n, uint64_t (*out_jacobian)[12]);

/* ---- device-pointer entry points (one rank per GPU: torch.distributed / NCCL callers) -----
 * All pointers are DEVICE memory on context device `dev_index`.
 *   d_bases   : n x 64 bytes, x || y Montgomery LE, 16-byte aligned; the pair (0,0) encodes infinity
 *               (an optional d_inf_mask, n bytes, non-zero = infinity, is honoured as well; may be NULL)
 *   d_scalars : n x 32 bytes Fr Montgomery LE, 16-byte aligned
 *   d_out     : 96 bytes, Jacobian Montgomery (same content as out_jacobian above)
 * Asynchronous on the context's stream unless `sync` != 0.                                  */
int b200msm_msm_device(b200msm_ctx* ctx, int dev_index,
                       const void* d_bases, const void* d_inf_mask, const void* d_scalars,
                       size_t n, void* d_out, int sync);
/* d_out <- sum of `count` Jacobian points at d_partials (96 bytes each): the multi-GPU
 * combine after an all-gather of per-rank partial sums.                                     */
int b200msm_sum_partials_device(b200msm_ctx* ctx, int dev_index, const void* d_partials, int count,
                                void* d_out, int sync);
/* Waits for everything enqueued on the context's streams; with "timing" on it also collects the stage timings of the last
 * asynchronous b200msm_msm_device call (b200msm_last_timings then returns them). */
int b200msm_sync(b200msm_ctx* ctx);
/* Make the context launch on a caller-owned cudaStream_t (e.g. torch's current stream) for
 * dev_index, so MSM kernels, NCCL collectives and the caller's events share one stream order. */
int b200msm_set_stream(b200msm_ctx* ctx, int dev_index, void* stream);
/* The cudaStream_t (as void*) the context launches on for dev_index, so callers can order
 * their own work / events against it. */
void* b200msm_stream(b200msm_ctx* ctx, int dev_index);

/* ---- benchmark-instance files ---------------------------------------------
 * The reference stores MSM instances as two files, `points` and `scalars`, written with arkworks'
 * compressed serialisation (src/msm/utils/preprocess.rs:181-225; read back by FileInputIterator :101-131 and
 * timed by arkworks_pippenger.rs:45-75).  These two calls decode the payloads on the GPU:
 *   compressed point = 32 B: canonical LE x; byte 31 bit 7 = y is the larger of (y, p-y); bit 6 = infinity
 *   scalar           = 32 B canonical LE (`BigInt<4>`); the MSM entry points take Montgomery `Fr` words.
 * out_xy64: count x 64 B Montgomery x||y ((0,0) for infinity); *n_invalid = records that are not curve points. */
int b200msm_decompress_g1(b200msm_ctx* ctx, const void* compressed, size_t count, void* out_xy64, uint64_t* n_invalid);
int b200msm_fr_to_montgomery(b200msm_ctx* ctx, const void* canonical, size_t count, void* out);

/* ---- test kit (mirrors the reference's public test_utils, metal_msm.rs:698-731, and its
 * single-purpose test kernels) -- NOT part of the drop-in surface ---------------
 * Deterministic synthetic inputs generated on the device: base i = T1[i mod 4096] + T2[i / 4096]
 * (tables of seeded multiples of G), scalars uniform in [0, r) in Montgomery form.
 * d_bases: n x 64 B, d_scalars: n x 32 B (device).  The two discrete-log tables are returned
 * to the host (canonical LE, 32 B each) so a checker can predict sum s_i*P_i in O(n) field ops.*/
int b200msm_testkit_generate(b200msm_ctx* ctx, int dev_index, uint64_t seed, size_t n,
                             void* d_bases, void* d_scalars,
                             uint8_t* h_table1_dlogs /*4096*32 or NULL*/,
                             uint8_t* h_table2_dlogs /*ceil(n/4096)*32 or NULL*/);
/* Auto-tuner evidence for another SM count (a MIG slice, a green context, a GPU shared with another tenant): occupy `n_sms`
 * SMs of the device with sleeping CTAs until b200msm_testkit_release_sms (or max_seconds, whichever comes first), so that
 * MSMs issued meanwhile run on the remaining SMs.  Tell the engine with b200msm_set_option("sm_count", remaining).        */
int b200msm_testkit_occupy_sms(b200msm_ctx* ctx, int dev_index, int n_sms, double max_seconds);
int b200msm_testkit_release_sms(b200msm_ctx* ctx, int dev_index);
/* Measured plain IMAD.WIDE.U32 rate (32x32+64 multiply-adds per second) on this device at its
 * current clocks: the denominator of the integer-multiply roofline.                         */
int b200msm_testkit_imad_peak(b200msm_ctx* ctx, int dev_index, double* macs_per_s);
/* Element-wise field / curve operations on HOST arrays through the production device
 * functions (one thread per element), for the limb -> field -> curve test pyramid.
 *   op: 0 fq_mul  1 fq_add  2 fq_sub  3 fq_sqr  4 fq_neg  5 fq_inv (0 -> 0)  6 fq_dbl  7 fq_inv_by (safegcd inversion, 0 -> 0)   (a, b, out: count x 32 B)
 *       8 fq_mulsub = a*b - c*d with one reduction (a: count x 64 B [a | c], b: count x 64 B [b | d], out: count x 32 B)
 *       10 xyzz_madd (a: count x 128 B XYZZ, b: count x 64 B affine, out: count x 128 B)
 *       11 xyzz_add  (a, b, out: count x 128 B)
 *       12 xyzz_dbl  (a, out: count x 128 B)
 *       13 xyzz_to_jacobian (a: count x 128 B, out: count x 96 B)
 *       14 jac_dbl (dbl-2009-l on a finite point; a, out: count x 96 B)
 *       15 k*P for a 32-bit k (a: count x 64 B affine, b: count x 8 B holding k, out: count x 128 B XYZZ)
 *       20 fr_from_mont (a, out: count x 32 B)
 *       30 fq2_mul  31 fq2_sqr (a, b, out: count x 64 B)
 *       32 g2 madd (a: count x 256 B XYZZ over Fq2, b: count x 128 B affine)  33 g2 add  34 g2 dbl (256 B records)                                              */
int b200msm_testkit_op(b200msm_ctx* ctx, int op, const void* a, const void* b, void* out, size_t count);
/* Stage-level access: run K1+K2 only and copy the CSR (bucket end offsets and sorted entries)
 * back to the host.  ends: up to 64*(2^(c-1)+1) u32; entries: up to 2*n*(ceil(127/c)+1) or n*ceil(254/c) u32
 * (pseudo-point index | sign<<31; with the GLV split on, index i < n is P_i with k1_i and n + i is
 * phi(P_i) with k2_i).  Returns the entry count, the window count and the pseudo-point count.        */
int b200msm_testkit_sort(b200msm_ctx* ctx, const void* scalars, size_t n, int window_bits,
                         uint32_t* ends, uint32_t* entries, uint64_t* n_entries,
                         int* num_windows, uint64_t* n_pseudo);

/* Stage-4 probe: run the pipeline on host inputs (bases: n x 64 B x||y, scalars: n x 32 B) and
 * return the per-window sums G_w = sum_m m*bucket[w][m] as XYZZ (16 u64 each, up to 64 windows). */
int b200msm_testkit_window_sums(b200msm_ctx* ctx, const void* bases64, const void* scalars, size_t n, int window_bits,
                                uint64_t* out_wsum, int* num_windows);

/* The same probe for the G2 pipeline (bases: n x 128 B x.c0|x.c1|y.c0|y.c1): per-window sums as XYZZ over Fq2, 32 u64 each. */
int b200msm_testkit_g2_window_sums(b200msm_ctx* ctx, const void* bases128, const void* scalars, size_t n, int window_bits,
                                   uint64_t* out_wsum, int* num_windows);

/* Host-only probes (no device needed): the slice plan of the host-buffer call ([0, n) cut into <= `slices` contiguous
 * ranges growing by ratio_pct percent; begins/lens have room for 8 entries) and the parallel copy that stages
 * pageable memory (`threads` includes the caller). */
int b200msm_testkit_slice_plan(size_t n, int slices, int ratio_pct, size_t* begins, size_t* lens, int* count);
int b200msm_testkit_parallel_copy(void* dst, const void* src, size_t bytes, int threads);

/* Table probe: copy `count` records (64 B x||y, Montgomery) of window `window` of a handle registered with "precompute"
 * (first shard) to the host, and return the table's window size / window count.  B200MSM_EINVAL for a plain handle. */
int b200msm_testkit_table(b200msm_ctx* ctx, const b200msm_bases* h, int window, size_t count, void* out_xy64,
                          int* window_bits, int* num_windows);

#ifdef __cplusplus
}
#endif
#endif /* B200MSM_H */
