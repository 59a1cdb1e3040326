/* cpu_msm.c -- CPU restatement of arkworks' BN254 G1 variable-base MSM.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this file.  The product (gpu-acceleration_b200/) never does.
 *
 * What it restates: the CPU side of the reference's metric, "arkworks CPU MSM":
 *     <G as VariableBaseMSM>::msm(&bases, &scalars)
 *     call sites: /root/reference/mopro-msm/src/msm/metal_msm/metal_msm.rs:756,
 *                 /root/reference/mopro-msm/src/msm/metal_msm/tests/cuzk/e2e.rs:55,
 *                 /root/reference/mopro-msm/benches/e2e.rs:56-60,
 *                 /root/reference/mopro-msm/src/msm/arkworks_pippenger.rs:26-29
 * The arithmetic lives in the un-vendored crates ark-ec 0.4.1 / ark-ff 0.4.1 / ark-bn254 0.4.0
 * (Cargo.toml:25-35).  Their published algorithm (`msm_bigint_wnaf`) is restated from the crate's
 * documented behaviour -- it is a "port" baseline, not the reference binary (no Rust toolchain):
 *   - scalars: Montgomery -> canonical (`into_bigint`);
 *   - window c = 3 if n < 32 else ln_without_floats(n) + 2,  ln_without_floats(n) = log2(n)*69/100;
 *   - `make_digits`: signed radix-2^c digits, carry = (coef + 2^(c-1)) >> c;
 *   - per window (windows in parallel, as rayon's `cfg_into_iter!(0..digits_count)`): 2^(c-1) used
 *     buckets of Jacobian points, `bucket += base` = mixed addition (madd-2007-bl), `-=` for
 *     negative digits; running-sum reduction (add-2007-bl);
 *   - fold from the top window down with c doublings (dbl-2009-l) per window.
 * Field: 4 x 64-bit limbs, Montgomery R = 2^256 (the same in-memory form as arkworks' Fp).
 * The only in-tree witness of the window rule is the dead-code ZPrize variant
 * (src/msm/trapdoortech_zprize_msm/local_msm.rs:377-380, which uses +1 with unsigned buckets).
 *
 * Pinning: constants are checked against the reference's literals by oracle/bn254.py self_check();
 * this file is validated against bn254.py (an independent big-int implementation) in
 * tests/test_oracle_c.py.  "Parity unpinned" by stored reference MSM vectors: none exist.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t v[4]; } fe;
typedef struct { fe x, y, z; } jac;   /* infinity <=> z == 0 */
typedef struct { fe x, y; int inf; } aff;

static const uint64_t Pm[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const uint64_t Rm[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
#define P_N0 0x87d20782e4866389ull
#define R_N0 0xc2e1f593efffffffull
static const fe FQ_ONE = {{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}};

static inline int ge(const uint64_t* a, const uint64_t* m) {
    for (int i = 3; i >= 0; i--) { if (a[i] != m[i]) return a[i] > m[i]; }
    return 1;
}
static inline void sub_n(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a[i] - b[i] - (uint64_t)br; r[i] = (uint64_t)d; br = (d >> 64) & 1; }
}
static inline void mont_mul(fe* r, const fe* a, const fe* b, const uint64_t* m, uint64_t n0) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a->v[j] * b->v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t q = t[0] * n0;
        c = (u128)q * m[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)q * m[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || ge(t, m)) sub_n(r->v, t, m); else memcpy(r->v, t, 32);
}
static inline void fq_mul(fe* r, const fe* a, const fe* b) { mont_mul(r, a, b, Pm, P_N0); }
static inline void fq_sqr(fe* r, const fe* a) { mont_mul(r, a, a, Pm, P_N0); }
static inline void fq_add(fe* r, const fe* a, const fe* b) {
    u128 c = 0; uint64_t t[4];
    for (int i = 0; i < 4; i++) { c += (u128)a->v[i] + b->v[i]; t[i] = (uint64_t)c; c >>= 64; }
    if (c || ge(t, Pm)) sub_n(r->v, t, Pm); else memcpy(r->v, t, 32);
}
static inline void fq_sub(fe* r, const fe* a, const fe* b) {
    uint64_t t[4]; u128 br = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a->v[i] - b->v[i] - (uint64_t)br; t[i] = (uint64_t)d; br = (d >> 64) & 1; }
    if (br) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)t[i] + Pm[i]; t[i] = (uint64_t)c; c >>= 64; } }
    memcpy(r->v, t, 32);
}
static inline void fq_dbl(fe* r, const fe* a) { fq_add(r, a, a); }
static inline int fe_is_zero(const fe* a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) { return memcmp(a, b, 32) == 0; }

static void jac_set_inf(jac* r) { r->x = FQ_ONE; r->y = FQ_ONE; memset(&r->z, 0, 32); }

/* dbl-2009-l, a = 0 */
static void jac_dbl(jac* r, const jac* p) {
    if (fe_is_zero(&p->z)) { *r = *p; return; }
    fe A, B, C, D, E, F, t, t2;
    fq_sqr(&A, &p->x); fq_sqr(&B, &p->y); fq_sqr(&C, &B);
    fq_add(&t, &p->x, &B); fq_sqr(&t, &t); fq_sub(&t, &t, &A); fq_sub(&t, &t, &C); fq_dbl(&D, &t);
    fq_dbl(&E, &A); fq_add(&E, &E, &A);
    fq_sqr(&F, &E);
    fe Z3; fq_mul(&Z3, &p->y, &p->z); fq_dbl(&Z3, &Z3);
    fq_dbl(&t, &D); fq_sub(&r->x, &F, &t);
    fq_sub(&t, &D, &r->x); fq_mul(&t, &E, &t);
    fq_dbl(&t2, &C); fq_dbl(&t2, &t2); fq_dbl(&t2, &t2);
    fq_sub(&r->y, &t, &t2);
    r->z = Z3;
}

/* madd-2007-bl with arkworks' special cases */
static void jac_madd(jac* r, const aff* q) {
    if (q->inf) return;
    if (fe_is_zero(&r->z)) { r->x = q->x; r->y = q->y; r->z = FQ_ONE; return; }
    fe Z1Z1, U2, S2, H, HH, I, J, rr, V, t, t2;
    fq_sqr(&Z1Z1, &r->z);
    fq_mul(&U2, &q->x, &Z1Z1);
    fq_mul(&S2, &q->y, &r->z); fq_mul(&S2, &S2, &Z1Z1);
    if (fe_eq(&r->x, &U2)) {
        if (fe_eq(&r->y, &S2)) { jac t3 = *r; jac_dbl(r, &t3); } else jac_set_inf(r);
        return;
    }
    fq_sub(&H, &U2, &r->x);
    fq_sqr(&HH, &H);
    fq_dbl(&I, &HH); fq_dbl(&I, &I);
    fq_mul(&J, &H, &I);
    fq_sub(&rr, &S2, &r->y); fq_dbl(&rr, &rr);
    fq_mul(&V, &r->x, &I);
    fe X3, Y3, Z3;
    fq_sqr(&X3, &rr); fq_sub(&X3, &X3, &J); fq_dbl(&t, &V); fq_sub(&X3, &X3, &t);
    fq_sub(&t, &V, &X3); fq_mul(&t, &rr, &t);
    fq_mul(&t2, &r->y, &J); fq_dbl(&t2, &t2);
    fq_sub(&Y3, &t, &t2);
    fq_add(&Z3, &r->z, &H); fq_sqr(&Z3, &Z3); fq_sub(&Z3, &Z3, &Z1Z1); fq_sub(&Z3, &Z3, &HH);
    r->x = X3; r->y = Y3; r->z = Z3;
}

/* add-2007-bl, complete */
static void jac_add(jac* r, const jac* q) {
    if (fe_is_zero(&q->z)) return;
    if (fe_is_zero(&r->z)) { *r = *q; return; }
    fe Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t, t2;
    fq_sqr(&Z1Z1, &r->z); fq_sqr(&Z2Z2, &q->z);
    fq_mul(&U1, &r->x, &Z2Z2); fq_mul(&U2, &q->x, &Z1Z1);
    fq_mul(&S1, &r->y, &q->z); fq_mul(&S1, &S1, &Z2Z2);
    fq_mul(&S2, &q->y, &r->z); fq_mul(&S2, &S2, &Z1Z1);
    if (fe_eq(&U1, &U2)) {
        if (fe_eq(&S1, &S2)) { jac t3 = *r; jac_dbl(r, &t3); } else jac_set_inf(r);
        return;
    }
    fq_sub(&H, &U2, &U1);
    fq_dbl(&I, &H); fq_sqr(&I, &I);
    fq_mul(&J, &H, &I);
    fq_sub(&rr, &S2, &S1); fq_dbl(&rr, &rr);
    fq_mul(&V, &U1, &I);
    fe X3, Y3, Z3;
    fq_sqr(&X3, &rr); fq_sub(&X3, &X3, &J); fq_dbl(&t, &V); fq_sub(&X3, &X3, &t);
    fq_sub(&t, &V, &X3); fq_mul(&t, &rr, &t);
    fq_mul(&t2, &S1, &J); fq_dbl(&t2, &t2);
    fq_sub(&Y3, &t, &t2);
    fq_add(&Z3, &r->z, &q->z); fq_sqr(&Z3, &Z3); fq_sub(&Z3, &Z3, &Z1Z1); fq_sub(&Z3, &Z3, &Z2Z2); fq_mul(&Z3, &Z3, &H);
    r->x = X3; r->y = Y3; r->z = Z3;
}

int oracle_ark_window(size_t n) {
    if (n < 32) return 3;
    int lg = 0;
    while (((size_t)1 << lg) < n) lg++;   /* ark_std::log2 = ceil(log2 n) */
    return lg * 69 / 100 + 2;
}

typedef struct {
    const uint8_t* bases; size_t stride, x_off, y_off, inf_off;
    const int32_t* digits; size_t n; int c, windows;
    jac* window_sums;
    int next; pthread_mutex_t mu;
} job_t;

static void window_sum(job_t* J, int w) {
    size_t nb = (size_t)1 << (J->c - 1);
    jac* buckets = (jac*)malloc(nb * sizeof(jac));
    for (size_t k = 0; k < nb; k++) jac_set_inf(&buckets[k]);
    for (size_t i = 0; i < J->n; i++) {
        int32_t d = J->digits[i * J->windows + w];
        if (d == 0) continue;
        const uint8_t* rec = J->bases + i * J->stride;
        aff q;
        memcpy(&q.x, rec + J->x_off, 32); memcpy(&q.y, rec + J->y_off, 32);
        q.inf = J->inf_off != (size_t)-1 && rec[J->inf_off] != 0;
        if (d > 0) jac_madd(&buckets[d - 1], &q);
        else { fe ny; fe zero = {{0, 0, 0, 0}}; fq_sub(&ny, &zero, &q.y); q.y = ny; jac_madd(&buckets[-d - 1], &q); }
    }
    jac run, res;
    jac_set_inf(&run); jac_set_inf(&res);
    for (size_t k = nb; k-- > 0;) { jac_add(&run, &buckets[k]); jac_add(&res, &run); }
    J->window_sums[w] = res;
    free(buckets);
}

static void* worker(void* arg) {
    job_t* J = (job_t*)arg;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        int w = J->next++;
        pthread_mutex_unlock(&J->mu);
        if (w >= J->windows) break;
        window_sum(J, w);
    }
    return NULL;
}

/* scalars: n x scalar_stride bytes, Fr Montgomery LE.  window_bits = 0 -> arkworks' rule.
 * Returns the number of threads actually used (<= windows: arkworks parallelises over windows). */
int oracle_msm(const uint8_t* bases, size_t stride, size_t x_off, size_t y_off, size_t inf_off,
               const uint8_t* scalars, size_t scalar_stride, size_t n, int threads, int window_bits,
               uint64_t out_jac[12]) {
    jac total; jac_set_inf(&total);
    if (n == 0) { memcpy(out_jac, &total, 96); return 0; }
    int c = window_bits ? window_bits : oracle_ark_window(n);
    int windows = (254 + c - 1) / c;
    if (254 - c * (windows - 1) == c) windows++;
    int32_t* digits = (int32_t*)malloc(n * (size_t)windows * 4);
    const fe one = {{1, 0, 0, 0}};
    for (size_t i = 0; i < n; i++) {
        fe s, m; memcpy(&m, scalars + i * scalar_stride, 32);
        mont_mul(&s, &m, &one, Rm, R_N0);  /* into_bigint */
        uint64_t carry = 0;
        for (int k = 0; k < windows; k++) {   /* make_digits */
            int bit = k * c, word = bit >> 6, sh = bit & 63;
            uint64_t v = word < 4 ? s.v[word] >> sh : 0;
            if (sh && word + 1 < 4) v |= s.v[word + 1] << (64 - sh);
            int64_t coef = (int64_t)((v & (((uint64_t)1 << c) - 1)) + carry);
            carry = (uint64_t)((coef + ((int64_t)1 << (c - 1))) >> c);
            digits[i * windows + k] = (int32_t)(coef - ((int64_t)carry << c));
        }
    }
    job_t J = {bases, stride, x_off, y_off, inf_off, digits, n, c, windows, NULL, 0, PTHREAD_MUTEX_INITIALIZER};
    J.window_sums = (jac*)malloc(windows * sizeof(jac));
    if (threads < 1) threads = 1;
    if (threads > windows) threads = windows;
    pthread_t* th = (pthread_t*)malloc(threads * sizeof(pthread_t));
    for (int t = 1; t < threads; t++) pthread_create(&th[t], NULL, worker, &J);
    worker(&J);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
    for (int w = windows - 1; w >= 1; w--) {
        jac_add(&total, &J.window_sums[w]);
        for (int k = 0; k < c; k++) { jac t3 = total; jac_dbl(&total, &t3); }
    }
    jac_add(&total, &J.window_sums[0]);
    memcpy(out_jac, &total, 96);
    free(th); free(J.window_sums); free(digits);
    return threads;
}

/* sum_i s_i * (t1[i mod 4096] + t2[i div 4096]) mod r; scalars Montgomery words, tables canonical;
 * result canonical.  The O(n) "checksum of checksums" for device-generated inputs (tests, bench). */
void oracle_dlog_checksum(const uint64_t* scalars_mont, size_t n, const uint64_t* t1, const uint64_t* t2, uint64_t out[4]) {
    fe acc = {{0, 0, 0, 0}};
    for (size_t i = 0; i < n; i++) {
        fe s, d, a, b, prod;
        memcpy(&s, scalars_mont + 4 * i, 32);                 /* s*R */
        memcpy(&a, t1 + 4 * (i & 4095), 32); memcpy(&b, t2 + 4 * (i >> 12), 32);
        u128 c = 0; uint64_t t[4];
        for (int k = 0; k < 4; k++) { c += (u128)a.v[k] + b.v[k]; t[k] = (uint64_t)c; c >>= 64; }
        if (c || ge(t, Rm)) sub_n(d.v, t, Rm); else memcpy(d.v, t, 32);
        mont_mul(&prod, &s, &d, Rm, R_N0);                    /* (sR)(d)/R = s*d canonical */
        c = 0;
        for (int k = 0; k < 4; k++) { c += (u128)acc.v[k] + prod.v[k]; t[k] = (uint64_t)c; c >>= 64; }
        if (c || ge(t, Rm)) sub_n(acc.v, t, Rm); else memcpy(acc.v, t, 32);
    }
    memcpy(out, &acc, 32);
}

/* k*G for canonical k (4 words) -> Jacobian Montgomery words. */
void oracle_scalar_mul_gen(const uint64_t k[4], uint64_t out_jac[12]) {
    aff G; G.x = FQ_ONE; fq_dbl(&G.y, &FQ_ONE); G.inf = 0;
    jac acc; jac_set_inf(&acc);
    for (int bit = 255; bit >= 0; bit--) {
        jac t3 = acc; jac_dbl(&acc, &t3);
        if ((k[bit >> 6] >> (bit & 63)) & 1) jac_madd(&acc, &G);
    }
    memcpy(out_jac, &acc, 96);
}

/* out = a + b (Jacobian Montgomery words): the host-side check of partial-sum combination. */
void oracle_jac_add(const uint64_t a[12], const uint64_t b[12], uint64_t out[12]) {
    jac x, y; memcpy(&x, a, 96); memcpy(&y, b, 96);
    jac_add(&x, &y);
    memcpy(out, &x, 96);
}

/* ---- synthetic inputs on the CPU (the `--impl reference` arm of bench.py must not load the CUDA library) ------------
 * Restates the device test kit (gpu-acceleration_b200/csrc/testkit_kernels.cuh: tk_mix64, tk_random_below_r,
 * k_tk_gen_table, k_tk_gen_bases) byte for byte: scalars[i] = random_below_r(seed ^ 0x5ca1a75, i),
 * T1[k] = random_below_r(seed ^ 0x7ab1e001, k) * G (k < 4096), T2[k] = random_below_r(seed ^ 0x7ab1e002, k) * G,
 * base[i] = T1[i mod 4096] + T2[i div 4096] in affine Montgomery words.  tests/test_gpu_msm.py compares the two
 * generators, which is one more independent check of the device field and curve arithmetic.                      */
static uint64_t tk_mix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
static void tk_random_below_r(uint64_t seed, uint64_t index, uint64_t out[4]) {
    for (uint64_t attempt = 0; attempt < 64; attempt++) {
        for (int k = 0; k < 4; k++) out[k] = tk_mix64(seed ^ tk_mix64((index * 4 + k) + (attempt << 44)));
        out[3] &= 0x3fffffffffffffffull;
        if (!ge(out, Rm)) return;
    }
    out[0] = 1; out[1] = out[2] = out[3] = 0;
}
static void fq_inv(fe* r, const fe* a) {   /* a^(p-2) */
    uint64_t e[4]; memcpy(e, Pm, 32); e[0] -= 2;
    fe acc = FQ_ONE;
    for (int bit = 253; bit >= 0; bit--) {
        fq_sqr(&acc, &acc);
        if ((e[bit >> 6] >> (bit & 63)) & 1) fq_mul(&acc, &acc, a);
    }
    *r = acc;
}
static void jac_to_aff(aff* r, const jac* p) {
    fe zi, zi2, zi3;
    fq_inv(&zi, &p->z); fq_sqr(&zi2, &zi); fq_mul(&zi3, &zi2, &zi);
    fq_mul(&r->x, &p->x, &zi2); fq_mul(&r->y, &p->y, &zi3); r->inf = 0;
}
/* bases_xy64: n x 64 B (may be NULL), scalars: n x 32 B (may be NULL), t1: 4096 x 32 B, t2: ceil(n/4096) x 32 B (may be NULL) */
int oracle_testkit_generate(uint64_t seed, size_t n, uint8_t* bases_xy64, uint8_t* scalars, uint8_t* t1, uint8_t* t2) {
    if (scalars)
        for (size_t i = 0; i < n; i++) tk_random_below_r(seed ^ 0x5ca1a75ull, i, (uint64_t*)(scalars + 32 * i));
    size_t n2 = (n + 4095) / 4096;
    uint64_t* dl = (uint64_t*)malloc((4096 + n2) * 32);
    if (!dl) return -1;
    for (size_t k = 0; k < 4096; k++) tk_random_below_r(seed ^ 0x7ab1e001ull, k, dl + 4 * k);
    for (size_t k = 0; k < n2; k++) tk_random_below_r(seed ^ 0x7ab1e002ull, k, dl + 4 * (4096 + k));
    if (t1) memcpy(t1, dl, 4096 * 32);
    if (t2) memcpy(t2, dl + 4096 * 4, n2 * 32);
    if (bases_xy64) {
        aff* tab = (aff*)malloc((4096 + n2) * sizeof(aff));
        for (size_t k = 0; k < 4096 + n2; k++) {
            uint64_t out[12]; jac j;
            oracle_scalar_mul_gen(dl + 4 * k, out);
            memcpy(&j, out, 96);
            jac_to_aff(&tab[k], &j);
        }
        /* affine sums with one inversion per block of 1024 (Montgomery's trick); T1[a] != +-T2[b] for random dlogs */
        enum { BLK = 1024 };
        fe den[BLK], pre[BLK];
        for (size_t i0 = 0; i0 < n; i0 += BLK) {
            size_t m = n - i0 < BLK ? n - i0 : BLK;
            fe run = FQ_ONE;
            for (size_t j = 0; j < m; j++) {
                const aff *a = &tab[(i0 + j) & 4095], *b = &tab[4096 + ((i0 + j) >> 12)];
                fq_sub(&den[j], &b->x, &a->x);
                pre[j] = run;
                fq_mul(&run, &run, &den[j]);
            }
            fe inv; fq_inv(&inv, &run);
            for (size_t j = m; j-- > 0;) {
                const aff *a = &tab[(i0 + j) & 4095], *b = &tab[4096 + ((i0 + j) >> 12)];
                fe dinv, lam, x3, y3, t;
                fq_mul(&dinv, &inv, &pre[j]);
                fq_mul(&inv, &inv, &den[j]);
                fq_sub(&t, &b->y, &a->y); fq_mul(&lam, &t, &dinv);
                fq_sqr(&x3, &lam); fq_sub(&x3, &x3, &a->x); fq_sub(&x3, &x3, &b->x);
                fq_sub(&t, &a->x, &x3); fq_mul(&y3, &lam, &t); fq_sub(&y3, &y3, &a->y);
                memcpy(bases_xy64 + 64 * (i0 + j), &x3, 32); memcpy(bases_xy64 + 64 * (i0 + j) + 32, &y3, 32);
            }
        }
        free(tab);
    }
    free(dl);
    return 0;
}
