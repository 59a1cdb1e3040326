// cuda_msm.hpp -- C++ host-side mirror of the reference's operator interface for the hot path,
// above the C ABI (include/b200msm.h).  Written in C++ because the reference is compiled code
// (Rust) and no Rust toolchain exists in this image; the Rust shim a maintainer would add is in
// ../rust/src/cuda_msm.rs and binds the very same entry point.
//
// Mirrors (names, argument meaning, error behaviour):
//   mopro_msm::msm::metal_msm::metal_variable_base_msm(&[G1Affine], &[Fr]) -> Result<G1Projective, Box<dyn Error>>
//   /root/reference/mopro-msm/src/msm/metal_msm/metal_msm.rs:642-695
#pragma once
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <string>

#include "../../include/b200msm.h"

namespace mopro_msm::msm::cuda_msm {

// arkworks memory layouts restated as standard-layout structs (what `&[G1Affine]` / `&[Fr]` hold).
struct Fq { uint64_t limbs[4]; };            // a*R mod p, R = 2^256, little-endian limbs
struct Fr { uint64_t limbs[4]; };            // s*R mod r
struct G1Affine { Fq x, y; bool infinity; }; // Affine::identity() = {0, 0, true}
struct G1Projective { Fq x, y, z; };         // Jacobian; identity <=> z == 0

template <typename T>
struct Result {                               // Result<T, Box<dyn Error>>
    bool ok = false;
    T value{};
    std::string error;
    explicit operator bool() const { return ok; }
};

inline b200msm_ctx* default_context(std::string* err) {
    static std::mutex mu;
    static b200msm_ctx* ctx = nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!ctx && b200msm_create(&ctx, nullptr, 0) != B200MSM_OK) {
        if (err) *err = b200msm_last_error(nullptr);
        ctx = nullptr;
    }
    return ctx;
}

inline Result<G1Projective> cuda_variable_base_msm(const G1Affine* bases, size_t bases_len, const Fr* scalars, size_t scalars_len,
                                                   b200msm_ctx* ctx = nullptr) {
    Result<G1Projective> r;
    if (bases_len == 0 || scalars_len == 0) {  // metal_msm.rs:647-649
        r.error = "Empty input";
        return r;
    }
    const size_t n = bases_len < scalars_len ? bases_len : scalars_len;  // metal_msm.rs:652-656
    if (!ctx) ctx = default_context(&r.error);
    if (!ctx) return r;
    uint64_t out[12];
    const int rc = b200msm_bn254_g1_msm(ctx, bases, sizeof(G1Affine), offsetof(G1Affine, x), offsetof(G1Affine, y),
                                        offsetof(G1Affine, infinity), scalars, sizeof(Fr), n, out);
    if (rc != B200MSM_OK) {
        r.error = b200msm_last_error(ctx);
        return r;
    }
    for (int k = 0; k < 4; k++) {
        r.value.x.limbs[k] = out[k];
        r.value.y.limbs[k] = out[4 + k];
        r.value.z.limbs[k] = out[8 + k];
    }
    r.ok = true;
    return r;
}

// BN254 G2 (the B2 MSM of a Groth16 prover; the reference has no G2 path).  Same contract as the G1 call.
struct Fq2 { Fq c0, c1; };
struct G2Affine { Fq2 x, y; bool infinity; };
struct G2Projective { Fq2 x, y, z; };

inline Result<G2Projective> cuda_variable_base_msm_g2(const G2Affine* bases, size_t bases_len, const Fr* scalars, size_t scalars_len,
                                                      b200msm_ctx* ctx = nullptr) {
    Result<G2Projective> r;
    if (bases_len == 0 || scalars_len == 0) {
        r.error = "Empty input";
        return r;
    }
    const size_t n = bases_len < scalars_len ? bases_len : scalars_len;
    if (!ctx) ctx = default_context(&r.error);
    if (!ctx) return r;
    uint64_t out[24];
    if (b200msm_bn254_g2_msm(ctx, bases, sizeof(G2Affine), offsetof(G2Affine, x), offsetof(G2Affine, y), offsetof(G2Affine, infinity),
                             scalars, sizeof(Fr), n, out) != B200MSM_OK) {
        r.error = b200msm_last_error(ctx);
        return r;
    }
    Fq* dst[6] = {&r.value.x.c0, &r.value.x.c1, &r.value.y.c0, &r.value.y.c1, &r.value.z.c0, &r.value.z.c1};
    for (int f = 0; f < 6; f++)
        for (int k = 0; k < 4; k++) dst[f]->limbs[k] = out[4 * f + k];
    r.ok = true;
    return r;
}

// A base set kept on the GPU(s) across MSMs -- the proving-key pattern (BASELINE config #5): the
// bases are uploaded once, each later call moves only the scalars.  With precompute = true the one-time window table
// 2^(c*w) * P_i is built as well (W x 64 B of HBM per point; MSMs then need no Horner step).  Nothing like it exists in
// the reference, whose every call re-uploads and re-converts the points (metal_msm.rs:74-201).
class RegisteredBases {
  public:
    RegisteredBases() = default;
    RegisteredBases(const RegisteredBases&) = delete;
    RegisteredBases& operator=(const RegisteredBases&) = delete;
    ~RegisteredBases() { release(); }

    // Err("Empty input") for an empty slice, like the MSM itself.
    Result<size_t> register_bases(const G1Affine* bases, size_t len, bool precompute = false, b200msm_ctx* ctx = nullptr) {
        Result<size_t> r;
        release();
        if (len == 0) { r.error = "Empty input"; return r; }
        if (!ctx) ctx = default_context(&r.error);
        if (!ctx) return r;
        const int rc = b200msm_register_bases_ex(ctx, bases, sizeof(G1Affine), offsetof(G1Affine, x), offsetof(G1Affine, y),
                                                 offsetof(G1Affine, infinity), len, nullptr, 0, precompute ? 1 : 0, &handle_);
        if (rc != B200MSM_OK) { r.error = b200msm_last_error(ctx); handle_ = nullptr; return r; }
        ctx_ = ctx;
        r.ok = true;
        r.value = len;
        return r;
    }

    // sum_i scalars[i] * bases[i] over the first min(scalars_len, len()) registered points.
    Result<G1Projective> msm(const Fr* scalars, size_t scalars_len) const {
        Result<G1Projective> r;
        if (!handle_ || scalars_len == 0) { r.error = "Empty input"; return r; }
        const size_t n = scalars_len < len() ? scalars_len : len();
        uint64_t out[12];
        if (b200msm_msm_registered(ctx_, handle_, scalars, sizeof(Fr), n, out) != B200MSM_OK) {
            r.error = b200msm_last_error(ctx_);
            return r;
        }
        for (int k = 0; k < 4; k++) {
            r.value.x.limbs[k] = out[k];
            r.value.y.limbs[k] = out[4 + k];
            r.value.z.limbs[k] = out[8 + k];
        }
        r.ok = true;
        return r;
    }

    size_t len() const { return handle_ ? b200msm_bases_len(handle_) : 0; }
    void release() {
        if (handle_) b200msm_release_bases(ctx_, handle_);
        handle_ = nullptr;
    }

  private:
    b200msm_ctx* ctx_ = nullptr;
    b200msm_bases* handle_ = nullptr;
};

}  // namespace mopro_msm::msm::cuda_msm
