"""GPU parity of the batched-affine bucket accumulation (k_accumulate_ba: chunk-local tree reduction, affine additions
sharing one safegcd inversion per round) against the same oracles as the XYZZ path.  Replaces the same reference stage
(shader/cuzk/smvp.metal:14-107); bar: equal group element, every option combination."""
import os
import random

import numpy as np
import pytest
import torch

import bn254 as o
import cpu_msm
import helpers as h

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _aff(words):
    return o.jac_to_affine(o.decode_jacobian(words))


def _fixtures():
    for fn in ("msm_cases.npz", "msm_sympy.npz"):
        z = np.load(os.path.join(HERE, "golden", fn))
        for name in z["names"]:
            exp = z[f"{name}/expected"]
            want = None if int(exp[8]) else (h.unwords(exp[0:4]), h.unwords(exp[4:8]))
            yield str(name), z[f"{name}/bases"], z[f"{name}/scalars"], want


@pytest.fixture()
def ba(ctx):
    ctx.set_option("batch_affine", 1)
    yield ctx
    for k, v in (("batch_affine", -1), ("ba_chunk", 0), ("ba_min_pairs", 0), ("window_bits", 0), ("glv", -1), ("slices", 0)):
        ctx.set_option(k, v)


@pytest.mark.parametrize("chunk,min_pairs", [(32, 1), (64, 2), (256, 24), (512, 8), (48, 1)])
def test_golden_fixtures_batched_affine(ba, chunk, min_pairs):
    """Small windows crowd the buckets, so every chunk runs several tree rounds; min_pairs = 1 forces rounds down to a
    single pair.  The fixtures include infinity bases, repeated bases (P + P inside a round), P / -P pairs (cancellation
    inside a round) and all-equal scalars (one bucket per window)."""
    ba.set_option("ba_chunk", chunk)
    ba.set_option("ba_min_pairs", min_pairs)
    for name, bases, scalars, want in _fixtures():
        for glv, wb in ((0, 4), (0, 7), (1, 5), (-1, 0), (0, 13)):
            ba.set_option("glv", glv)
            ba.set_option("window_bits", wb)
            assert _aff(ba.msm(bases, scalars).words) == want, (name, chunk, min_pairs, glv, wb)


def test_special_cases_inside_rounds(ba):
    """Adjacent equal points (doubling in the batch), adjacent opposite points (cancellation -> the infinity marker feeds
    the next round), infinity records next to finite ones, and whole buckets that cancel."""
    rng = random.Random(5)
    pts = o.random_points(40, 77)
    P, Q, S = pts[0], pts[1], pts[2]
    bases = [P, P, P, P, Q, o.affine_neg(Q), Q, None, S, None, None, S, P, o.affine_neg(P), o.affine_neg(P), P] * 6 + pts
    for sc_kind in ("same", "two", "random"):
        if sc_kind == "same":
            sc = [12345] * len(bases)
        elif sc_kind == "two":
            sc = [3 if i % 2 else 3 + (1 << 20) for i in range(len(bases))]
        else:
            sc = [rng.randrange(o.R_ORDER) for _ in bases]
        want = o.jac_to_affine(o.msm_pippenger(bases, sc, 8))
        for chunk, mp in ((32, 1), (256, 1), (512, 24)):
            ba.set_option("ba_chunk", chunk)
            ba.set_option("ba_min_pairs", mp)
            for glv, wb in ((0, 6), (0, 11), (1, 8)):
                ba.set_option("glv", glv)
                ba.set_option("window_bits", wb)
                assert _aff(ba.msm(h.pack_bases(bases), h.pack_scalars(sc)).words) == want, (sc_kind, chunk, mp, glv, wb)


def test_registered_table_and_slices_with_batched_affine(ba):
    pts = o.random_points(3000, 801)
    pts[17] = None
    sc = o.random_scalars(3000, 802)
    want = o.jac_to_affine(o.msm_pippenger(pts, sc, 9))
    hb, hs = h.pack_bases(pts), h.pack_scalars(sc)
    ba.set_option("ba_chunk", 64)
    ba.set_option("ba_min_pairs", 2)
    ba.set_option("window_bits", 8)
    for slices in (1, 2, 3):
        ba.set_option("slices", slices)
        assert _aff(ba.msm(hb, hs).words) == want, slices
    ba.set_option("slices", 0)
    for pre in (0, 8, 13):
        key = ba.register_bases(hb, precompute=pre)
        try:
            assert _aff(ba.msm_registered(key, hs).words) == want, pre
        finally:
            key.release()


def _device_case(ctx, n, seed):
    d_bases = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    d_scalars = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    t1, t2 = ctx.testkit_generate(seed, n, d_bases, d_scalars, want_dlogs=True)
    dl = h.unwords(cpu_msm.dlog_checksum(d_scalars.cpu().numpy().view(np.uint64).reshape(n, 4), t1, t2))
    want = _aff(cpu_msm.scalar_mul_gen(np.array(h.words(dl), dtype=np.uint64)))
    return d_bases, d_scalars, want


def _run_device(ctx, d_bases, d_scalars, n):
    d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.msm_device(d_bases, d_scalars, n, d_out)
    return _aff(d_out.cpu().numpy().view(np.uint64))


@pytest.mark.parametrize("log_n", [16, 20])
def test_checksum_sizes_batched_affine(ba, log_n):
    """2^16 (the reference's own test size) and 2^20 (BASELINE configs[1]) through the discrete-log checksum, plain and
    GLV windows, chunk 256 and 512."""
    n = 1 << log_n
    d_bases, d_scalars, want = _device_case(ba, n, 0xBA00 + log_n)
    for chunk in (256, 512):
        ba.set_option("ba_chunk", chunk)
        for glv, wb in ((-1, 0), (0, 13), (0, 16)):
            ba.set_option("glv", glv)
            ba.set_option("window_bits", wb)
            assert _run_device(ba, d_bases, d_scalars, n) == want, (log_n, chunk, glv, wb)


def test_batched_affine_equals_xyzz_on_skewed_scalars(ba):
    """Witness-like scalars (most are 0 or 1, a few huge buckets): both accumulation engines must agree."""
    n = 1 << 15
    rng = np.random.default_rng(3)
    d_bases = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    d_scalars = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ba.testkit_generate(99, n, d_bases, d_scalars)
    sc = d_scalars.cpu().numpy().view(np.uint64).reshape(n, 4).copy()
    one = np.array(h.words(o.to_mont(1, o.R_ORDER)), dtype=np.uint64)
    kind = rng.integers(0, 100, n)
    sc[kind < 45] = 0
    sc[(kind >= 45) & (kind < 90)] = one
    d_scalars.copy_(torch.from_numpy(sc.view(np.uint8).reshape(-1)))
    got = _run_device(ba, d_bases, d_scalars, n)
    ba.set_option("batch_affine", 0)
    assert got == _run_device(ba, d_bases, d_scalars, n)
