"""GPU parity, end to end, through the reference-facing call `cuda_variable_base_msm` and the C ABI.
Reads like the reference's own e2e test (tests/cuzk/e2e.rs:14-63, metal_msm.rs:739-760):
    assert_eq!(metal_variable_base_msm(&bases, &scalars).unwrap(), G::msm(&bases, &scalars).unwrap())
with the oracle standing in for arkworks.  Bar: equal group element (bit-exact after normalisation)."""
import os
import random

import numpy as np
import pytest

import b200msm
import bn254 as o
import helpers as h
from b200msm import cuda_variable_base_msm

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "msm_cases.npz")


def _expect(pts, sc):
    return o.jac_to_affine(o.msm_pippenger(pts, sc, 8 if len(pts) > 64 else 5))


@pytest.mark.parametrize("glv", [1, 0])
def test_golden_fixtures_all_windows(ctx, glv):
    z = np.load(GOLDEN)
    ctx.set_option("glv", glv)
    try:
        for name in z["names"]:
            bases, scalars, exp = z[f"{name}/bases"], z[f"{name}/scalars"], z[f"{name}/expected"]
            want = None if int(exp[8]) else (h.unwords(exp[0:4]), h.unwords(exp[4:8]))
            for w in (0, 4, 7, 8, 13, 15, 16, 17, 20):  # 0 = auto; 8/13/15/16 are the reference's table (metal_msm.rs:661-673)
                ctx.set_option("window_bits", w)
                res = cuda_variable_base_msm(bases, scalars, ctx)
                assert h.result_affine(res) == want, (name, w, glv)
                assert res.into_affine() == want
    finally:
        ctx.set_option("window_bits", 0)
        ctx.set_option("glv", -1)


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 1000, 4097])
def test_random_sizes(ctx, n):
    pts = o.random_points(n, 1000 + n)
    sc = o.random_scalars(n, 2000 + n)
    res = cuda_variable_base_msm(h.pack_bases(pts), h.pack_scalars(sc), ctx)
    assert h.result_affine(res) == _expect(pts, sc)


def test_reference_config_2_16(ctx):
    """BASELINE config #1 size (2^16, the reference's own test size). Bases are 2^16 distinct points
    built as sums of two small tables so the expected value costs O(n) scalar-field ops + one
    scalar multiplication in the oracle (the 'checksum of checksums' route)."""
    n = 1 << 16
    rng = random.Random(42)
    t1 = [rng.randrange(1, o.R_ORDER) for _ in range(256)]
    t2 = [rng.randrange(1, o.R_ORDER) for _ in range(256)]
    G = o.affine_to_jac(o.GEN)
    T1 = [o.jac_scalar_mul(k, G) for k in t1]
    T2 = [o.jac_scalar_mul(k, G) for k in t2]
    pts = _batch_affine([o.jac_add(T1[i & 255], T2[i >> 8]) for i in range(n)])
    sc = o.random_scalars(n, 4242)
    dlog = sum(s * (t1[i & 255] + t2[i >> 8]) for i, s in enumerate(sc)) % o.R_ORDER
    want = o.jac_to_affine(o.jac_scalar_mul(dlog, G))
    bases = h.pack_bases(pts)
    scal = h.pack_scalars(sc)
    for w in (0, 13, 16):  # 13 is the reference's choice at 2^16
        ctx.set_option("window_bits", w)
        assert h.result_affine(cuda_variable_base_msm(bases, scal, ctx)) == want
    ctx.set_option("window_bits", 0)
    # linearity: halves sum to the whole
    a = cuda_variable_base_msm(bases[: n // 2], scal[: n // 2], ctx)
    b = cuda_variable_base_msm(bases[n // 2:], scal[n // 2:], ctx)
    assert o.jac_to_affine(o.jac_add(o.decode_jacobian(a.words), o.decode_jacobian(b.words))) == want


def _batch_affine(jacs):
    """Montgomery batch inversion so 2^16 normalisations stay cheap in Python."""
    zs = [j[2] for j in jacs]
    pref = [1]
    for z in zs:
        pref.append(pref[-1] * z % o.P)
    inv = pow(pref[-1], -1, o.P)
    out = [None] * len(jacs)
    for i in range(len(jacs) - 1, -1, -1):
        zi = inv * pref[i] % o.P
        inv = inv * zs[i] % o.P
        zi2 = zi * zi % o.P
        out[i] = (jacs[i][0] * zi2 % o.P, jacs[i][1] * zi2 * zi % o.P)
    return out


def test_length_mismatch_truncates(ctx):
    # metal_msm.rs:652-656
    pts = o.random_points(10, 5)
    sc = o.random_scalars(7, 6)
    res = cuda_variable_base_msm(h.pack_bases(pts), h.pack_scalars(sc), ctx)
    assert h.result_affine(res) == _expect(pts[:7], sc)


def test_layouts_and_strides(ctx):
    pts = o.random_points(50, 9)
    sc = o.random_scalars(50, 10)
    want = _expect(pts, sc)
    # 64-byte records without infinity flag (fast path: no repack)
    assert h.result_affine(ctx.msm(h.pack_bases(pts, with_inf=False), h.pack_scalars(sc))) == want
    # odd layout: 96-byte records, y before x, infinity flag at +80; 40-byte scalar records
    raw = np.zeros((50, 12), dtype=np.uint64)
    b = h.pack_bases(pts)
    raw[:, 1:5] = b[:, 4:8]
    raw[:, 6:10] = b[:, 0:4]
    raw[:, 10] = b[:, 8]
    sraw = np.zeros((50, 5), dtype=np.uint64)
    sraw[:, 0:4] = h.pack_scalars(sc)
    sraw[:, 4] = 0xDEADBEEF
    res = ctx.msm_raw(raw.ctypes.data, 96, 48, 8, 80, sraw.ctypes.data, 40, 50)
    assert h.result_affine(res) == want


def test_registered_bases_and_batch(ctx):
    pts = o.random_points(200, 31)
    hb = ctx.register_bases(h.pack_bases(pts))
    try:
        assert len(hb) == 200
        scs = [o.random_scalars(200, 40 + k) for k in range(3)] + [o.random_scalars(120, 50)]
        for sc in scs:
            assert h.result_affine(ctx.msm_registered(hb, h.pack_scalars(sc))) == _expect(pts[: len(sc)], sc)
        outs = ctx.msm_batch([hb] * 4, [h.pack_scalars(sc) for sc in scs])  # BASELINE config #5 shape (A, B1, C, H)
        for sc, r in zip(scs, outs):
            assert h.result_affine(r) == _expect(pts[: len(sc)], sc)
    finally:
        hb.release()


def test_bad_arguments(ctx):
    with pytest.raises(b200msm.MsmError):
        ctx.set_option("window_bits", 99)
    with pytest.raises(b200msm.MsmError):
        ctx.set_option("nope", 1)
    with pytest.raises(b200msm.MsmError, match="Empty input"):
        ctx.msm(np.zeros((0, 9), dtype=np.uint64), np.zeros((0, 4), dtype=np.uint64), 0)
    a = np.zeros((4, 9), dtype=np.uint64)
    with pytest.raises(b200msm.MsmError):
        ctx.msm_raw(a.ctypes.data, 70, 0, 32, 64, a.ctypes.data, 32, 4)  # stride not a multiple of 8


def test_timings_and_launch_count(ctx):
    ctx.set_option("timing", 1)
    pts = o.random_points(64, 3)
    sc = o.random_scalars(64, 4)
    ctx.msm(h.pack_bases(pts), h.pack_scalars(sc))
    t = ctx.timings()
    ctx.set_option("timing", 0)
    assert t["kernel_launches"] >= 8 and t["total_ms"] > 0 and t["entries"] > 0
    assert t["num_windows"] == o.num_windows_for(t["window_bits"], 127)  # small n: GLV half-scalars by default


def test_reference_srs_points(ctx):
    """MSM over the externally produced G1 points shipped in the reference tree (example-app/ios/*_srs.bin ->
    tests/golden/ref_srs_g1.npz): bytes neither the oracle nor the kernels generated, consumed as 64-byte
    records without any conversion."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_srs_g1.npz"))
    bases = np.concatenate([z["plonk"], z["gemini"], z["hyperplonk"]])
    pts = [(o.from_mont(h.unwords(b[0:4])), o.from_mont(h.unwords(b[4:8]))) for b in bases]
    sc = o.random_scalars(len(pts), 808)
    res = cuda_variable_base_msm(bases, h.pack_scalars(sc), ctx)
    assert h.result_affine(res) == o.jac_to_affine(o.msm_naive(pts, sc))


@pytest.mark.parametrize("slices", [2, 3, 8])
def test_sliced_host_pipeline_small(ctx, slices):
    """The host-buffer entry point uploads and accumulates the point range in slices (copy of slice k+1 under the
    arithmetic of slice k), every slice adding into the same bucket array, then one reduce.  Any slice count must give the
    same group element, including n < slices, ragged tails, infinity records, strided layouts, both scalar
    splits and a wide-digit window."""
    ctx.set_option("slices", slices)
    try:
        for n, seed in ((1, 1), (5, 2), (1000, 3), (4099, 4)):
            pts = o.random_points(n, 700 + seed)
            sc = o.random_scalars(n, 800 + seed)
            if n >= 5:
                pts[1] = None          # infinity record in slice 0
                pts[n - 1] = None      # ... and in the last slice
                sc[2] = 0
                sc[n - 2] = o.R_ORDER - 1
            want = _expect(pts, sc)
            bases, scal = h.pack_bases(pts), h.pack_scalars(sc)
            for glv, w in ((-1, 0), (0, 0), (1, 13), (0, 17)):
                ctx.set_option("glv", glv)
                ctx.set_option("window_bits", w)
                assert h.result_affine(ctx.msm(bases, scal)) == want, (n, glv, w)
            ctx.set_option("glv", -1)
            ctx.set_option("window_bits", 0)
            # odd layout through the slice path: 96-byte records, y before x, flag at +80; 40-byte scalar records
            raw = np.zeros((n, 12), dtype=np.uint64)
            raw[:, 1:5] = bases[:, 4:8]
            raw[:, 6:10] = bases[:, 0:4]
            raw[:, 10] = bases[:, 8]
            sraw = np.zeros((n, 5), dtype=np.uint64)
            sraw[:, 0:4] = scal
            sraw[:, 4] = 0xDEADBEEF
            assert h.result_affine(ctx.msm_raw(raw.ctypes.data, 96, 48, 8, 80, sraw.ctypes.data, 40, n)) == want
            # a lone MSM over a registered handle (plain, then with the window table) uploads its scalars in slices too
            for pre in (0, 13):
                ctx.set_option("precompute", pre)
                hb = ctx.register_bases(bases)
                ctx.set_option("precompute", 0)
                try:
                    assert h.result_affine(ctx.msm_registered(hb, scal)) == want, (n, "registered", pre)
                    if n > 7:
                        assert h.result_affine(ctx.msm_registered(hb, scal[:n - 3])) == _expect(pts[:n - 3], sc[:n - 3])
                finally:
                    hb.release()
    finally:
        ctx.set_option("slices", 0)
        ctx.set_option("precompute", 0)
        ctx.set_option("glv", -1)
        ctx.set_option("window_bits", 0)


def test_sliced_first_call_on_fresh_context():
    """Regression: the sliced path must allocate the result buffer itself (a fresh context whose first call is sliced
    used to hand a null output pointer to the Horner kernel)."""
    c2 = b200msm.Context()
    try:
        c2.set_option("slices", 2)
        pts = o.random_points(300, 77)
        sc = o.random_scalars(300, 78)
        assert h.result_affine(c2.msm(h.pack_bases(pts), h.pack_scalars(sc))) == _expect(pts, sc)
    finally:
        c2.close()


@pytest.mark.parametrize("precompute", [1, 8, 13, 17])
def test_registered_bases_with_precomputed_table(ctx, precompute):
    """registering a base set with "precompute" builds table[w][i] = 2^(c*w) P_i once; every later MSM
    over the handle then feeds all digit windows into one bucket set (no Horner step).  Same group element as the
    oracle for full, shorter and batched scalar sets, with infinity records in the set."""
    pts = o.random_points(333, 91)
    pts[0] = None
    pts[200] = None
    ctx.set_option("precompute", precompute)
    try:
        hb = ctx.register_bases(h.pack_bases(pts))
    finally:
        ctx.set_option("precompute", 0)
    try:
        scs = [o.random_scalars(333, 60 + k) for k in range(2)] + [o.random_scalars(150, 70), [0] * 333,
                                                                   [o.R_ORDER - 1] * 333, [1] * 333]
        for sc in scs:
            assert h.result_affine(ctx.msm_registered(hb, h.pack_scalars(sc))) == _expect(pts[: len(sc)], sc)
        outs = ctx.msm_batch([hb] * len(scs), [h.pack_scalars(sc) for sc in scs])
        for sc, r in zip(scs, outs):
            assert h.result_affine(r) == _expect(pts[: len(sc)], sc)
    finally:
        hb.release()
    # a plain handle registered afterwards is unaffected
    hb2 = ctx.register_bases(h.pack_bases(pts))
    try:
        assert h.result_affine(ctx.msm_registered(hb2, h.pack_scalars(scs[0]))) == _expect(pts, scs[0])
    finally:
        hb2.release()


def test_option_fuzz(ctx):
    """Every tuning knob is result-neutral: random combinations of window size, scalar split, chunk length, reduce engine,
    reduce chain length and slice count over three fixed inputs (expected value computed once per input)."""
    rng = random.Random(20261017)
    inputs = []
    for n, seed in ((1, 1), (257, 2), (1500, 3)):
        pts = o.random_points(n, 3100 + seed)
        sc = o.random_scalars(n, 3200 + seed)
        if n > 10:
            pts[7] = None
            sc[9] = 0
            sc[10] = o.R_ORDER - 1
            pts[12] = pts[11]                     # repeated base: P + P inside a bucket when the digits agree
            sc[12] = sc[11]
            pts[14] = o.affine_neg(pts[13])       # P + (-P)
            sc[14] = sc[13]
        inputs.append((h.pack_bases(pts), h.pack_scalars(sc), _expect(pts, sc)))
    knobs = ("window_bits", "glv", "chunk", "coop_reduce", "reduce_log2", "slices", "ranked_sort", "fix_chunks", "rowcol_reduce", "groups", "batch_affine")
    try:
        for trial in range(60):
            glv = rng.choice((-1, 0, 1))
            admissible = (4, 5, 7, 8, 10, 11, 12, 13, 15, 16, 19, 20) if glv != 0 else tuple(range(4, 21))
            opts = {"window_bits": rng.choice((0,) + admissible), "glv": glv, "chunk": rng.choice((0, 0, 1, 3, 8, 64, 500)),
                    "coop_reduce": rng.choice((-1, 0, 1)), "reduce_log2": rng.choice((-1, -1, 0, 2, 5)),
                    "slices": rng.choice((0, 1, 2, 5)), "ranked_sort": rng.choice((-1, 0, 1, 2)),
                    "fix_chunks": rng.choice((-1, 0, 1)), "rowcol_reduce": rng.choice((-1, 0, 1)), "groups": rng.choice((0, 0, 2, 4)),
                    "batch_affine": rng.choice((-1, -1, -1, 1))}
            for k in knobs:
                ctx.set_option(k, opts[k])
            for bases, scal, want in inputs:
                assert h.result_affine(ctx.msm(bases, scal)) == want, (trial, opts, len(scal))
    finally:
        for k, v in (("window_bits", 0), ("glv", -1), ("chunk", 0), ("coop_reduce", -1), ("reduce_log2", -1), ("slices", 0),
                     ("ranked_sort", -1), ("fix_chunks", -1), ("rowcol_reduce", -1), ("groups", 0), ("batch_affine", -1)):
            ctx.set_option(k, v)


def test_slices_add_up_in_place_with_cancellation(ctx):
    """The slices of the host call accumulate into ONE bucket array (k_accumulate `into`): a bucket filled by slice 0 must be
    picked up, doubled or cancelled correctly by the later slices.  Second half of the points = the first half again (P + P
    across slices), third quarter negated (P + (-P) across slices), equal slice lengths so that the halves face each other."""
    n = 2048
    pts = o.random_points(n // 2, 5151)
    sc = o.random_scalars(n // 2, 5252)
    pts2 = list(pts)
    for i in range(n // 4):
        pts2[i] = o.affine_neg(pts2[i])          # slice 1 cancels the first quarter of slice 0 ...
    all_pts, all_sc = pts + pts2, sc + sc       # ... and doubles the second quarter
    want = _expect(all_pts, all_sc)
    bases, scal = h.pack_bases(all_pts), h.pack_scalars(all_sc)
    ctx.set_option("slice_ratio", 100)
    try:
        for slices in (2, 4):
            ctx.set_option("slices", slices)
            for glv, w, fc in ((-1, 0, -1), (0, 8, 1), (1, 13, 0), (0, 16, 1)):
                ctx.set_option("glv", glv)
                ctx.set_option("window_bits", w)
                ctx.set_option("fix_chunks", fc)
                assert h.result_affine(ctx.msm(bases, scal)) == want, (slices, glv, w, fc)
        # everything cancels: the result is the identity
        neg = [o.affine_neg(p) for p in pts]
        ctx.set_option("slices", 2)
        ctx.set_option("glv", -1)
        ctx.set_option("window_bits", 0)
        assert h.result_affine(ctx.msm(h.pack_bases(pts + neg), scal)) is None
    finally:
        for k, v in (("slices", 0), ("slice_ratio", 0), ("glv", -1), ("window_bits", 0), ("fix_chunks", -1)):
            ctx.set_option(k, v)


def test_register_bases_ex(ctx):
    """The table choice as a call argument (b200msm_register_bases_ex) instead of the context option."""
    pts = o.random_points(120, 17)
    sc = o.random_scalars(120, 18)
    want = _expect(pts, sc)
    for pre in (0, 1, 15):
        hb = ctx.register_bases(h.pack_bases(pts), precompute=pre)
        try:
            assert h.result_affine(ctx.msm_registered(hb, h.pack_scalars(sc))) == want
            if pre:
                assert ctx.testkit_table(hb, 0, 1)[1] == (pre if pre >= 8 else 8)   # 120 points: automatic window 8
            else:
                with pytest.raises(b200msm.MsmError):
                    ctx.testkit_table(hb, 0, 1)
        finally:
            hb.release()
    with pytest.raises(b200msm.MsmError):
        ctx.register_bases(h.pack_bases(pts), precompute=5)


def test_concurrent_callers_share_a_context(ctx):
    """A context is internally serialised (SURVEY 8(b) threading row): host threads calling into it at the same time --
    ctypes drops the GIL during the call -- each get their own correct result; a second context runs truly in parallel."""
    import threading
    jobs = []
    for k in range(6):
        n = 200 + 37 * k
        pts = o.random_points(n, 4000 + k)
        sc = o.random_scalars(n, 4100 + k)
        jobs.append((h.pack_bases(pts), h.pack_scalars(sc), _expect(pts, sc)))
    other = b200msm.Context()
    results = [None] * (2 * len(jobs))

    def work(i, c):
        bases, scal, _ = jobs[i % len(jobs)]
        for _ in range(5):
            results[i] = h.result_affine(c.msm(bases, scal))

    try:
        threads = [threading.Thread(target=work, args=(i, ctx if i < len(jobs) else other)) for i in range(2 * len(jobs))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    finally:
        other.close()
    for i, got in enumerate(results):
        assert got == jobs[i % len(jobs)][2], i


def test_survey_edge_set(ctx):
    """The correctness-only inputs SURVEY 8(d) lists: every base identical, P / -P pairs, all scalars below 2^32, all
    scalars zero, all one, all r - 1."""
    n = 5000
    P0 = o.random_points(1, 1234)[0]
    sc = o.random_scalars(n, 1235)
    same = h.pack_bases([P0] * n)
    want = o.jac_to_affine(o.jac_scalar_mul(sum(sc) % o.R_ORDER, o.affine_to_jac(P0)))
    assert h.result_affine(ctx.msm(same, h.pack_scalars(sc))) == want            # every base identical: P + P everywhere
    pairs = []
    for k in range(n // 2):
        pairs += [P0, o.affine_neg(P0)]
    sc2 = []
    for k in range(n // 2):
        sc2 += [sc[k], sc[k]]
    assert h.result_affine(ctx.msm(h.pack_bases(pairs), h.pack_scalars(sc2))) is None   # P and -P pairs cancel exactly
    pts = o.random_points(2000, 1236)
    small = [s & 0xFFFFFFFF for s in o.random_scalars(2000, 1237)]
    assert h.result_affine(ctx.msm(h.pack_bases(pts), h.pack_scalars(small))) == _expect(pts, small)
    total = o.jac_to_affine(o.msm_naive(pts[:200], [1] * 200))
    bases200 = h.pack_bases(pts[:200])
    assert h.result_affine(ctx.msm(bases200, h.pack_scalars([0] * 200))) is None
    assert h.result_affine(ctx.msm(bases200, h.pack_scalars([1] * 200))) == total
    assert h.result_affine(ctx.msm(bases200, h.pack_scalars([o.R_ORDER - 1] * 200))) == o.affine_neg(total)
