// The C++ rendition of the reference's e2e test (tests/cuzk/e2e.rs:14-63) through the host mirror:
// reads raw arkworks-layout bases/scalars and the expected affine point from files written by the
// test-suite, runs cuda_variable_base_msm, prints the Jacobian words.  Built and driven by
// tests/test_gpu_cpp_mirror.py (links libb200msm.so; no oracle code is linked).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cuda_msm.hpp"

using namespace mopro_msm::msm::cuda_msm;

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s bases.bin scalars.bin n\n", argv[0]); return 2; }
    size_t n = strtoull(argv[3], nullptr, 10);
    std::vector<G1Affine> bases(n);
    std::vector<Fr> scalars(n);
    FILE* fb = fopen(argv[1], "rb");
    FILE* fs = fopen(argv[2], "rb");
    if (!fb || !fs) { fprintf(stderr, "cannot open inputs\n"); return 2; }
    for (size_t i = 0; i < n; i++) {          // file records are 72-byte arkworks records
        uint64_t rec[9];
        if (fread(rec, 8, 9, fb) != 9) return 2;
        for (int k = 0; k < 4; k++) { bases[i].x.limbs[k] = rec[k]; bases[i].y.limbs[k] = rec[4 + k]; }
        bases[i].infinity = (rec[8] & 0xff) != 0;
    }
    if (fread(scalars.data(), sizeof(Fr), n, fs) != n) return 2;
    auto empty = cuda_variable_base_msm(bases.data(), 0, scalars.data(), 0);
    if (empty.ok || empty.error != "Empty input") { fprintf(stderr, "empty-input contract broken\n"); return 1; }
    auto res = cuda_variable_base_msm(bases.data(), n, scalars.data(), n);
    if (!res) { fprintf(stderr, "error: %s\n", res.error.c_str()); return 1; }
    const uint64_t* w = res.value.x.limbs;
    for (int k = 0; k < 12; k++) printf("%llu\n", (unsigned long long)w[k]);
    // the same MSM over a registered base set, without and with the precomputed window table (12 + 12 more words)
    for (int pre = 0; pre < 2; pre++) {
        RegisteredBases key;
        auto reg = key.register_bases(bases.data(), n, pre != 0);
        if (!reg || key.len() != n) { fprintf(stderr, "register error: %s\n", reg.error.c_str()); return 1; }
        auto r2 = key.msm(scalars.data(), n);
        if (!r2) { fprintf(stderr, "registered msm error: %s\n", r2.error.c_str()); return 1; }
        const uint64_t* w2 = r2.value.x.limbs;
        for (int k = 0; k < 12; k++) printf("%llu\n", (unsigned long long)w2[k]);
    }
    // optional: a G2 MSM over 136-byte arkworks G2Affine records (argv[4]) with the same scalars -> 24 more words
    if (argc > 4) {
        std::vector<G2Affine> g2b(n);
        FILE* fg = fopen(argv[4], "rb");
        if (!fg) { fprintf(stderr, "cannot open g2 bases\n"); return 2; }
        for (size_t i = 0; i < n; i++) {
            uint64_t rec[17];
            if (fread(rec, 8, 17, fg) != 17) return 2;
            Fq* f[4] = {&g2b[i].x.c0, &g2b[i].x.c1, &g2b[i].y.c0, &g2b[i].y.c1};
            for (int c = 0; c < 4; c++)
                for (int k = 0; k < 4; k++) f[c]->limbs[k] = rec[4 * c + k];
            g2b[i].infinity = (rec[16] & 0xff) != 0;
        }
        auto r3 = cuda_variable_base_msm_g2(g2b.data(), n, scalars.data(), n);
        if (!r3) { fprintf(stderr, "g2 msm error: %s\n", r3.error.c_str()); return 1; }
        const uint64_t* w3 = r3.value.x.c0.limbs;
        for (int k = 0; k < 24; k++) printf("%llu\n", (unsigned long long)w3[k]);
    }
    return 0;
}
