"""GPU parity, stage level: K1 (signed digits + histogram) and K2 (scan + scatter) against the
oracle's restatement of the reference's convert/transpose kernels (SURVEY Appendix A;
tests/cuzk/convert_point_coords_and_decompose_scalars.rs:104-245, tests/cuzk/transpose.rs:6-118).
Order inside a bucket is unspecified here (the reference's stable order is not needed for a
commutative sum), so buckets are compared as sets."""
import random

import numpy as np
import pytest

import bn254 as o
import helpers as h

pytestmark = pytest.mark.gpu


def _pseudo(pts, scalars, glv):
    """The (pseudo-point, signed scalar) list the engine sorts: plain, or the GLV-expanded 2n-term problem."""
    if glv:
        return o.glv_expand(pts if pts is not None else [o.GEN] * len(scalars), scalars)
    return (list(pts) if pts is not None else [o.GEN] * len(scalars)), [s % o.R_ORDER for s in scalars]


def _check(ctx, scalars, w, glv=1):
    # both sort variants: ranks from the histogram pass + atomic-free scatter (default), and cursor atomics in the scatter
    # ... and the shared-memory radix partition (2), the default above 2^19 digits
    for ranked in (2, 1, 0):
        _check_one(ctx, scalars, w, glv, ranked)


def _check_one(ctx, scalars, w, glv, ranked):
    ctx.set_option("glv", glv)
    ctx.set_option("ranked_sort", ranked)
    try:
        ends, entries, npseudo = ctx.testkit_sort(h.pack_scalars(scalars), w)
    finally:
        ctx.set_option("glv", -1)
        ctx.set_option("ranked_sort", -1)
    _, ks = _pseudo(None, scalars, glv)
    K = o.num_windows_for(w, 127 if glv else 254)
    half = 1 << (w - 1)
    assert npseudo == len(ks) and ends.shape == (K, half + 1)
    digs = [o.signed_digits_signed(k, w, K) for k in ks]
    flat_end = ends.reshape(-1)
    assert np.all(np.diff(flat_end.astype(np.int64)) >= 0)
    nonzero = sum(1 for d in digs for x in d if x != 0)
    assert len(entries) == nonzero == int(flat_end[-1])
    start = 0
    for k in range(K):
        want = {}
        for i in range(len(ks)):
            d = digs[i][k]
            if d:
                want.setdefault(abs(d), set()).add(i | ((1 << 31) if d < 0 else 0))
        for m in range(half + 1):
            end = int(ends[k, m])
            got = set(int(e) for e in entries[start:end])
            assert got == want.get(m, set()), (w, k, m)
            assert end - start == len(got)
            start = end


@pytest.mark.parametrize("glv", [0, 1])
@pytest.mark.parametrize("w", [4, 8, 13, 16, 17, 20])
def test_sort_random(ctx, w, glv):
    _check(ctx, o.random_scalars(300, 100 + w), w, glv)


def test_sort_skewed(ctx):
    r = o.R_ORDER
    rng = random.Random(1)
    sc = [0] * 40 + [1] * 70 + [r - 1] * 33 + [5] * 64 + [rng.randrange(1 << 32) for _ in range(50)] + [1 << 253, (1 << 15), (1 << 16) - 1]
    sc += [o.GLV_LAMBDA, r - o.GLV_LAMBDA, r // 2, r // 2 + 1, r - 2]
    # scalars whose GLV halves hit the int16 corner digit (+-2^15 at w = 16), both signs
    for k1, k2 in ((0x8000, -0x8000), (-(0x8000 << 16), 0x8000 + (0x8000 << 32)), (-0x8000, -(0x8000 + (0x8000 << 16)))):
        sc.append((k1 + k2 * o.GLV_LAMBDA) % r)
    for w in (8, 16):
        for glv in (0, 1):
            _check(ctx, sc, w, glv)


def test_sort_ragged_sizes(ctx):
    for n in (1, 2, 31, 33, 257):
        _check(ctx, o.random_scalars(n, n), 13)


def _csr_canonical(ends, entries):
    """Entries sorted inside every bucket (the order there is unspecified)."""
    flat = ends.reshape(-1).astype(np.int64)
    sizes = np.diff(np.concatenate([[0], flat]))
    bucket_of = np.repeat(np.arange(len(flat)), sizes)
    order = np.lexsort((entries, bucket_of))
    return entries[order]


@pytest.mark.parametrize("log_n,w,glv", [(17, 16, 1), (17, 13, 0), (18, 17, 0), (16, 8, 1), (17, 20, 0), (15, 11, 1)])
def test_partitioned_sort_equals_ranked_sort(ctx, log_n, w, glv):
    """The shared-memory radix partition (k_decompose_count / k_pscan / k_partition / k_place) against the ranked sort at
    sizes with many partitions per window and ragged tiles: identical bucket ends, identical bucket contents."""
    n = (1 << log_n) + 4099
    rng = np.random.default_rng(log_n * 100 + w)
    sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    sc[:, 3] &= np.uint64((1 << 60) - 1)            # Montgomery words below r
    sc[: n // 8] = sc[0]                            # a heavy bucket in every window
    sc[n // 8: n // 4, 1:] = 0
    out = {}
    for ranked in (1, 2):
        ctx.set_option("glv", glv)
        ctx.set_option("ranked_sort", ranked)
        try:
            out[ranked] = ctx.testkit_sort(sc, w)
        finally:
            ctx.set_option("glv", -1)
            ctx.set_option("ranked_sort", -1)
    (e1, x1, n1), (e2, x2, n2) = out[1], out[2]
    assert n1 == n2 and np.array_equal(e1, e2) and len(x1) == len(x2)
    assert np.array_equal(_csr_canonical(e1, x1), _csr_canonical(e2, x2))


@pytest.mark.parametrize("kind", ["all_equal", "half_equal_half_small", "zeros_and_ones"])
@pytest.mark.parametrize("w,glv", [(16, 1), (13, 0), (20, 0)])
def test_partitioned_sort_heavy_partitions(ctx, kind, w, glv):
    """Skewed scalars: a partition (often a single bucket) holds far more than PSORT_HEAVY digits and is cut into slices that
    take their positions from global per-bucket counters / cursors.  CSR identical to the ranked sort's."""
    import helpers as hh
    n = (1 << 18) + 321
    rng = np.random.default_rng(w * 7 + glv)
    sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    sc[:, 3] &= np.uint64((1 << 60) - 1)
    one = np.array(hh.words(o.R_MOD_R), dtype=np.uint64)            # Montgomery form of 1
    if kind == "all_equal":
        sc[:] = sc[0]
    elif kind == "half_equal_half_small":
        sc[: n // 2] = sc[1]
        small = [np.array(hh.words(k * o.R_MOD_R % o.R_ORDER), dtype=np.uint64) for k in range(2, 18)]
        idx = rng.integers(0, 16, size=n - n // 2)
        sc[n // 2:] = np.stack(small)[idx]
    else:
        u = rng.random(n)
        sc[u < 0.45] = 0
        sc[(u >= 0.45) & (u < 0.9)] = one
    out = {}
    for ranked in (1, 2):
        ctx.set_option("glv", glv)
        ctx.set_option("ranked_sort", ranked)
        try:
            out[ranked] = ctx.testkit_sort(sc, w)
        finally:
            ctx.set_option("glv", -1)
            ctx.set_option("ranked_sort", -1)
    (e1, x1, n1), (e2, x2, n2) = out[1], out[2]
    assert n1 == n2 and np.array_equal(e1, e2) and len(x1) == len(x2)
    assert np.array_equal(_csr_canonical(e1, x1), _csr_canonical(e2, x2))


def _check_window_sums(ctx, n, w, seed, glv):
    """Stage 3+4: per-window sums G_w = sum_m m * bucket[m] against the oracle's bucket/reduce
    restatement (smvp.metal:14-107 + pbpr.metal:33-148; tests/cuzk/smvp.rs:245-302, pbpr.rs:161-216)."""
    pts = o.random_points(n, seed)
    sc = o.random_scalars(n, seed + 1)
    ctx.set_option("glv", glv)
    try:
        got = h.unpack_xyzz(ctx.testkit_window_sums(h.pack_bases(pts, with_inf=False), h.pack_scalars(sc), w))
    finally:
        ctx.set_option("glv", -1)
    pts2, ks = _pseudo(pts, sc, glv)
    K = o.num_windows_for(w, 127 if glv else 254)
    half = 1 << (w - 1)
    assert len(got) == K
    digs = [o.signed_digits_signed(k, w, K) for k in ks]
    for k in range(K):
        buckets = o.stage_bucket_sums(pts2, [d[k] for d in digs], half)
        want = o.xyzz_to_affine(o.stage_bucket_reduce(buckets))
        assert o.xyzz_to_affine(got[k]) == want, (n, w, k)


@pytest.mark.parametrize("n,w", [(1, 6), (2, 6), (3, 4), (50, 5), (300, 8), (300, 11), (2000, 13), (3000, 16)])
@pytest.mark.parametrize("coop", [2, 1, 0])
@pytest.mark.parametrize("glv", [0, 1])
def test_window_sums(ctx, n, w, glv, coop):
    """All bucket-reduce implementations: row / column sums + one cooperative level (2), the recursive cooperative levels (1)
    and the thread-per-segment kernels (0)."""
    ctx.set_option("coop_reduce", 1 if coop else 0)
    ctx.set_option("rowcol_reduce", 1 if coop == 2 else 0)
    try:
        _check_window_sums(ctx, n, w, 900 + n + w, glv)
    finally:
        ctx.set_option("coop_reduce", -1)
        ctx.set_option("rowcol_reduce", -1)


@pytest.mark.parametrize("c", [8, 13, 20])
def test_precomputed_table_contents(ctx, c):
    """table[w][i] = 2^(c*w) * P_i, affine, infinity marker preserved: every window of the device table against the
    oracle's table_expand."""
    pts = o.random_points(9, 300 + c)
    pts[4] = None
    want = o.table_expand(pts, c)
    ctx.set_option("precompute", c)
    try:
        hb = ctx.register_bases(h.pack_bases(pts))
    finally:
        ctx.set_option("precompute", 0)
    try:
        for w in range(len(want)):
            rec, cc, nw = ctx.testkit_table(hb, w, len(pts))
            assert cc == c and nw == len(want) == o.num_windows_for(c)
            for i, pt in enumerate(want[w]):
                x, y = h.unwords(rec[i, 0:4]), h.unwords(rec[i, 4:8])
                if pt is None:
                    assert (x, y) == (0, 0)
                else:
                    assert (o.from_mont(x), o.from_mont(y)) == pt, (w, i)
    finally:
        hb.release()
