"""GPU parity at BASELINE.json's full single-GPU size (2^20) through size-independent properties:
the device-generated bases have known discrete logs (base_i = (a[i mod 4096] + b[i div 4096])*G), so
sum_i s_i*P_i = (sum_i s_i*(a+b) mod r)*G costs the checker O(n) scalar-field multiplications and
ONE scalar multiplication -- a checksum of checksums -- plus linearity (halves sum to the whole)
and window-size independence."""
import numpy as np
import pytest
import torch

import bn254 as o
import helpers as h

pytestmark = pytest.mark.gpu


def _ints(arr):  # (n,4) uint64 -> python ints
    a = arr.astype(object)
    return (a[:, 0] + (a[:, 1] << 64) + (a[:, 2] << 128) + (a[:, 3] << 192)).tolist()


def _generate(ctx, n, seed):
    d_bases = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    d_scalars = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    t1, t2 = ctx.testkit_generate(seed, n, d_bases, d_scalars, want_dlogs=True)
    return d_bases, d_scalars, _ints(t1), _ints(t2)


def _expected(d_scalars, n, t1, t2, lo=0, hi=None):
    hi = n if hi is None else hi
    sm = _ints(d_scalars.cpu().numpy().view(np.uint64).reshape(n, 4))
    acc = 0
    for i in range(lo, hi):
        acc += sm[i] * (t1[i & 4095] + t2[i >> 12])
    dlog = acc * o.RINV_R % o.R_ORDER  # scalars are Montgomery words: s = sm * R^-1
    return o.jac_to_affine(o.jac_scalar_mul(dlog, o.affine_to_jac(o.GEN)))


def _run(ctx, d_bases, d_scalars, n, off=0):
    d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.msm_device(d_bases.data_ptr() + off * 64, d_scalars.data_ptr() + off * 32, n, d_out)
    return o.jac_to_affine(o.decode_jacobian(d_out.cpu().numpy().view(np.uint64)))


def test_generated_inputs_are_what_they_claim(ctx):
    n = 10000
    d_bases, d_scalars, t1, t2 = _generate(ctx, n, 7)
    b = d_bases.cpu().numpy().view(np.uint64).reshape(n, 8)
    G = o.affine_to_jac(o.GEN)
    for i in (0, 1, 4095, 4096, 9999):
        pt = (o.from_mont(h.unwords(b[i, 0:4])), o.from_mont(h.unwords(b[i, 4:8])))
        assert pt == o.jac_to_affine(o.jac_scalar_mul((t1[i & 4095] + t2[i >> 12]) % o.R_ORDER, G))
    s = _ints(d_scalars.cpu().numpy().view(np.uint64).reshape(n, 4))
    assert all(v < o.R_ORDER for v in s) and len(set(s)) == n


def test_full_size_2_20_checksum_linearity_windows(ctx):
    n = 1 << 20
    d_bases, d_scalars, t1, t2 = _generate(ctx, n, 0xB200)
    want = _expected(d_scalars, n, t1, t2)
    for w in (0, 13, 15, 16, 18):
        ctx.set_option("window_bits", w)
        assert _run(ctx, d_bases, d_scalars, n) == want, w
    ctx.set_option("window_bits", 0)
    # linearity on device pointers: [0, n/2) + [n/2, n)
    half = n // 2
    a = _run(ctx, d_bases, d_scalars, half)
    b = _run(ctx, d_bases, d_scalars, half, off=half)
    assert a == _expected(d_scalars, n, t1, t2, 0, half)
    assert o.jac_to_affine(o.jac_add(o.affine_to_jac(a), o.affine_to_jac(b))) == want
    # multi-GPU combine kernel on the two partials
    parts = torch.zeros(2 * 96, dtype=torch.uint8, device="cuda")
    outp = torch.zeros(96, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.msm_device(d_bases, d_scalars, half, parts.data_ptr())
    ctx.msm_device(d_bases.data_ptr() + half * 64, d_scalars.data_ptr() + half * 32, half, parts.data_ptr() + 96)
    ctx.sum_partials_device(parts, 2, outp)
    assert o.jac_to_affine(o.decode_jacobian(outp.cpu().numpy().view(np.uint64))) == want


def test_not_power_of_two_and_skewed_device_inputs(ctx):
    n = (1 << 17) + 12345
    d_bases, d_scalars, t1, t2 = _generate(ctx, n, 99)
    assert _run(ctx, d_bases, d_scalars, n) == _expected(d_scalars, n, t1, t2)
    # skew: every scalar equal (one bucket per window holds all n points: the load-imbalance case)
    sc = d_scalars.view(torch.int64).reshape(n, 4)
    sc[:] = sc[0:1].clone()
    torch.cuda.synchronize()
    assert _run(ctx, d_bases, d_scalars, n) == _expected(d_scalars, n, t1, t2)


def test_witness_like_scalars_2_20(ctx):
    """Groth16-witness-like skew at full size: ~45 % zeros, ~45 % ones, 10 % full-width scalars.  All the ones land
    in ONE bucket (window 0, magnitude 1) holding ~2^19 points: the warp-aggregated histogram/cursor atomics and the
    per-CTA long-bucket fix-up carry it.  Result checked with the discrete-log checksum; also must not be slow."""
    import time
    n = 1 << 20
    d_bases, d_scalars, t1, t2 = _generate(ctx, n, 0x517)
    sc = d_scalars.view(torch.int64).reshape(n, 4)
    one_mont = torch.tensor(h.words(o.R_MOD_R), dtype=torch.uint64).view(torch.int64).to(sc.device)
    sel = torch.rand(n, device=sc.device)
    sc[sel < 0.45] = 0
    sc[(sel >= 0.45) & (sel < 0.90)] = one_mont
    torch.cuda.synchronize()
    want = _expected(d_scalars, n, t1, t2)
    for glv in (-1, 0):
        ctx.set_option("glv", glv)
        try:
            assert _run(ctx, d_bases, d_scalars, n) == want
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _run(ctx, d_bases, d_scalars, n)
            dt = (time.perf_counter() - t0) * 1e3
        finally:
            ctx.set_option("glv", -1)
        assert dt < 50.0, f"skewed 2^20 MSM took {dt:.1f} ms"


def test_host_entry_slices_2_20(ctx):
    """2^20 points through the HOST-buffer entry point (the reference-facing call): unsliced, the automatic slice
    count and the maximum must all give the checksum's group element."""
    n = 1 << 20
    d_bases, d_scalars, t1, t2 = _generate(ctx, n, 0x51CE)
    want = _expected(d_scalars, n, t1, t2)
    hb = d_bases.cpu().numpy().view(np.uint64).reshape(n, 8)
    hs = d_scalars.cpu().numpy().view(np.uint64).reshape(n, 4)
    try:
        for slices in (1, 0, 8):
            ctx.set_option("slices", slices)
            assert h.result_affine(ctx.msm(hb, hs)) == want, slices
        # slices 1.. sorted on the side stream under the accumulation before them (auto at this size), and on the main stream
        for overlap in (1, 0):
            ctx.set_option("sort_overlap", overlap)
            for slices in (3, 8):
                ctx.set_option("slices", slices)
                for _ in range(2):   # back to back: the second call reuses every slice's work set
                    assert h.result_affine(ctx.msm(hb, hs)) == want, (overlap, slices)
        ctx.set_option("sort_overlap", -1)
        # transfer-bound feedback: pretend the arithmetic is ~7x faster (sm_count = 1024 scales the model), so that the uploads
        # of one call exceed it and the NEXT call takes the equal-slice plan; results must not move
        ctx.set_option("slices", 0)
        ctx.set_option("sm_count", 1024)
        try:
            for _ in range(3):
                assert h.result_affine(ctx.msm(hb, hs)) == want, "adaptive slices"
        finally:
            ctx.set_option("sm_count", 0)
        ctx.set_option("adaptive_slices", 0)
        assert h.result_affine(ctx.msm(hb, hs)) == want
        ctx.set_option("adaptive_slices", -1)
        # 72-byte arkworks records (repack path) with the automatic slice count
        ctx.set_option("slices", 0)
        hb9 = np.zeros((n, 9), dtype=np.uint64)
        hb9[:, :8] = hb
        assert h.result_affine(ctx.msm(hb9, hs)) == want
        # witness-like skew through the slices: ~half the points share one bucket per slice (long-bucket queue per slice)
        sc = d_scalars.view(torch.int64).reshape(n, 4)
        one_mont = torch.tensor(h.words(o.R_MOD_R), dtype=torch.uint64).view(torch.int64).to(sc.device)
        sel = torch.rand(n, device=sc.device)
        sc[sel < 0.45] = 0
        sc[(sel >= 0.45) & (sel < 0.90)] = one_mont
        torch.cuda.synchronize()
        want2 = _expected(d_scalars, n, t1, t2)
        hs2 = d_scalars.cpu().numpy().view(np.uint64).reshape(n, 4)
        for slices in (0, 1, 5):
            ctx.set_option("slices", slices)
            assert h.result_affine(ctx.msm(hb, hs2)) == want2, ("skew", slices)
    finally:
        ctx.set_option("slices", 0)
        ctx.set_option("sort_overlap", -1)


def test_precomputed_table_2_20(ctx):
    """2^20 registered bases with the precomputed window table (auto window): checksum parity, twice over the same handle."""
    n = 1 << 20
    d_bases, d_scalars, t1, t2 = _generate(ctx, n, 0x7AB1E)
    want = _expected(d_scalars, n, t1, t2)
    hb = d_bases.cpu().numpy().view(np.uint64).reshape(n, 8)
    hs = d_scalars.cpu().numpy().view(np.uint64).reshape(n, 4)
    ctx.set_option("precompute", 1)
    try:
        handle = ctx.register_bases(hb)
    finally:
        ctx.set_option("precompute", 0)
    try:
        assert h.result_affine(ctx.msm_registered(handle, hs)) == want
        half = n // 2
        assert h.result_affine(ctx.msm_registered(handle, hs[:half])) == _expected(d_scalars, n, t1, t2, 0, half)
    finally:
        handle.release()


def test_size_2_16_plus_1(ctx):
    """n = 2^16 + 1 (SURVEY 8(d) edge set): one point past a power of two, device inputs, checksum-verified; the last point
    alone must account for the difference to the 2^16 prefix."""
    n = (1 << 16) + 1
    d_bases, d_scalars, t1, t2 = _generate(ctx, n, 0x10001)
    assert _run(ctx, d_bases, d_scalars, n) == _expected(d_scalars, n, t1, t2)
    assert _run(ctx, d_bases, d_scalars, 1, off=n - 1) == _expected(d_scalars, n, t1, t2, n - 1, n)


def test_device_generator_equals_cpu_generator(ctx):
    """The device test kit and its CPU restatement (oracle/cpu_msm.c: oracle_testkit_generate, used by the reference arm of
    bench.py so that it never loads the CUDA library) produce the same bytes: 2^13 + 5 affine additions + inversions on the
    device against independent 64-bit-limb CPU arithmetic."""
    import cpu_msm
    n = (1 << 13) + 5
    d_bases, d_scalars, t1, t2 = _generate(ctx, n, 0xB2000000)
    cb, cs, ct1, ct2 = cpu_msm.testkit_generate(0xB2000000, n)
    assert np.array_equal(d_bases.cpu().numpy().view(np.uint64).reshape(n, 8), cb)
    assert np.array_equal(d_scalars.cpu().numpy().view(np.uint64).reshape(n, 4), cs)
    assert _ints(ct1) == t1 and _ints(ct2) == t2


def _expected_c(d_scalars, n, t1w, t2w, lo=0, hi=None):
    """The checksum of checksums through the C oracle (python big ints are too slow at 2^24)."""
    import cpu_msm
    hi = n if hi is None else hi
    sc = d_scalars.cpu().numpy().view(np.uint64).reshape(n, 4)
    # dlog_checksum indexes the tables by absolute position: feed it whole 4096-aligned prefixes and subtract
    def upto(k):
        if k == 0:
            return 0
        return h.unwords(cpu_msm.dlog_checksum(sc[:k], t1w, t2w))
    dlog = (upto(hi) - upto(lo)) % o.R_ORDER
    return o.jac_to_affine(o.decode_jacobian(cpu_msm.scalar_mul_gen(np.array(h.words(dlog), dtype=np.uint64))))


def test_north_star_size_2_24_checksum_and_halves(ctx):
    """BASELINE configs[2] size on ONE GPU: 2^24 points (auto policy: plain windows, c = 20), checked by the discrete-log
    checksum; the two halves (the 2-GPU shards) sum to the whole; the batched-affine engine agrees."""
    n = 1 << 24
    d_bases = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    d_scalars = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    t1w, t2w = ctx.testkit_generate(0xB224, n, d_bases, d_scalars, want_dlogs=True)
    want = _expected_c(d_scalars, n, t1w, t2w)
    assert _run(ctx, d_bases, d_scalars, n) == want
    half = n // 2
    a = _run(ctx, d_bases, d_scalars, half)
    b = _run(ctx, d_bases, d_scalars, half, off=half)
    assert a == _expected_c(d_scalars, n, t1w, t2w, 0, half)
    assert o.jac_to_affine(o.jac_add(o.affine_to_jac(a), o.affine_to_jac(b))) == want
    ctx.set_option("batch_affine", 1)
    try:
        assert _run(ctx, d_bases, d_scalars, n) == want
    finally:
        ctx.set_option("batch_affine", -1)


def test_groth16_batch_4x2_22_registered(ctx):
    """BASELINE configs[4] on one GPU: four MSMs of 2^22 points over registered bases through b200msm_msm_batch, plain and
    with the precomputed window table; every result against the checksum."""
    n = 1 << 22
    sets, wants = [], []
    for k in range(4):
        d_bases = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
        d_scalars = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        t1w, t2w = ctx.testkit_generate(0xB250 + k, n, d_bases, d_scalars, want_dlogs=True)
        wants.append(_expected_c(d_scalars, n, t1w, t2w))
        sets.append((d_bases.cpu().numpy().view(np.uint64).reshape(n, 8), d_scalars.cpu().numpy().view(np.uint64).reshape(n, 4)))
        del d_bases, d_scalars
    for pre in (0, 1):
        handles = [ctx.register_bases(hb, precompute=pre) for hb, _ in sets]
        try:
            res = ctx.msm_batch(handles, [hs for _, hs in sets])
        finally:
            for hd in handles:
                hd.release()
        for k in range(4):
            assert h.result_affine(res[k]) == wants[k], (pre, k)


@pytest.mark.parametrize("glv,w,chunk", [(-1, 0, 0), (0, 13, 16), (1, 8, 64), (0, 17, 32)])
def test_fixup_per_chunk_equals_fixup_per_bucket(ctx, glv, w, chunk):
    """k_fixup_empty + k_fixup_chunks (one thread per chunk whose last bucket runs on; the default from 2^22 digits) against
    k_fixup (one thread per bucket) and the checksum: random scalars, then a skewed set (every scalar equal: buckets that
    span hundreds of chunks go through the long-bucket queue) -- at chunk lengths that make buckets span 1, 2 and many chunks."""
    n = (1 << 16) + 777
    d_bases, d_scalars, t1, t2 = _generate(ctx, n, 0xF1C + w)
    for skew in (False, True):
        if skew:
            sc = d_scalars.view(torch.int64).reshape(n, 4)
            sc[: n // 2] = sc[0:1].clone()
            torch.cuda.synchronize()
        want = _expected(d_scalars, n, t1, t2)
        ctx.set_option("glv", glv)
        ctx.set_option("window_bits", w)
        ctx.set_option("chunk", chunk)
        try:
            for fc in (1, 0):
                ctx.set_option("fix_chunks", fc)
                assert _run(ctx, d_bases, d_scalars, n) == want, (skew, fc)
        finally:
            for k, v in (("glv", -1), ("window_bits", 0), ("chunk", 0), ("fix_chunks", -1)):
                ctx.set_option(k, v)


def test_option_fuzz_large_and_skewed(ctx):
    """Random knob combinations at a size where the large-input engines and their skew paths are real (2^17 points: with
    every scalar equal one bucket per window holds 2^18 digits = a heavy partition of the sort, and spans 32 768 chunks = a
    giant bucket of the fix-up): every combination must give the checksum's group element."""
    import random
    rng = random.Random(7117)
    n = (1 << 17) + 99
    d_bases, d_scalars, t1, t2 = _generate(ctx, n, 0xF022)
    base = d_scalars.clone()
    one = torch.tensor(h.words(o.R_MOD_R), dtype=torch.uint64).view(torch.int64).to(d_scalars.device)
    knobs = ("window_bits", "glv", "chunk", "ranked_sort", "fix_chunks", "rowcol_reduce", "coop_reduce", "groups")
    try:
        for kind in ("uniform", "zeros_ones", "all_equal", "half_equal"):
            d_scalars.copy_(base)
            sc = d_scalars.view(torch.int64).reshape(n, 4)
            u = torch.rand(n, device=sc.device)
            if kind == "zeros_ones":
                sc[u < 0.45] = 0
                sc[(u >= 0.45) & (u < 0.9)] = one
            elif kind == "all_equal":
                sc[:] = sc[0:1].clone()
            elif kind == "half_equal":
                sc[u < 0.5] = sc[3:4].clone()
            torch.cuda.synchronize()
            want = _expected(d_scalars, n, t1, t2)
            for trial in range(7):
                glv = rng.choice((-1, 0, 1))
                admissible = (8, 13, 16, 20) if glv != 0 else (8, 13, 16, 17, 20)
                opts = {"window_bits": rng.choice((0,) + admissible), "glv": glv, "chunk": rng.choice((0, 0, 8, 32, 64)),
                        "ranked_sort": rng.choice((1, 2, 2)), "fix_chunks": rng.choice((0, 1, 1)), "rowcol_reduce": rng.choice((-1, 1)),
                        "coop_reduce": rng.choice((-1, 0, 1)), "groups": rng.choice((0, 0, 2))}
                for k in knobs:
                    ctx.set_option(k, opts[k])
                assert _run(ctx, d_bases, d_scalars, n) == want, (kind, trial, opts)
    finally:
        for k, v in (("window_bits", 0), ("glv", -1), ("chunk", 0), ("ranked_sort", -1), ("fix_chunks", -1), ("rowcol_reduce", -1),
                     ("coop_reduce", -1), ("groups", 0)):
            ctx.set_option(k, v)
