"""Single-process multi-device paths of the C library: b200msm_create over several ordinals shards every MSM
by contiguous point range and adds the 96-byte partials on the first device.  Skipped on a one-GPU box."""
import numpy as np
import pytest
import torch

import b200msm
import bn254 as o
import helpers as h

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mctx():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    c = b200msm.Context(list(range(min(4, torch.cuda.device_count()))))
    yield c
    c.close()


def _expect(pts, sc):
    return o.jac_to_affine(o.msm_pippenger(pts, sc, 8))


def test_sharded_host_call(mctx):
    for n, seed in ((1, 1), (3, 2), (1001, 3), (5000, 4)):
        pts = o.random_points(n, 900 + seed)
        sc = o.random_scalars(n, 950 + seed)
        if n > 3:
            pts[n // 2] = None
        want = _expect(pts, sc)
        for slices in (0, 1, 3):
            mctx.set_option("slices", slices)
            assert h.result_affine(mctx.msm(h.pack_bases(pts), h.pack_scalars(sc))) == want, (n, slices)
    mctx.set_option("slices", 0)


@pytest.mark.parametrize("precompute", [0, 1, 13])
def test_sharded_registered_and_batch(mctx, precompute):
    pts = o.random_points(777, 41)
    pts[5] = None
    mctx.set_option("precompute", precompute)
    try:
        hb = mctx.register_bases(h.pack_bases(pts))
    finally:
        mctx.set_option("precompute", 0)
    try:
        scs = [o.random_scalars(777, 60), o.random_scalars(300, 61), o.random_scalars(1, 62)]
        for sc in scs:
            assert h.result_affine(mctx.msm_registered(hb, h.pack_scalars(sc))) == _expect(pts[: len(sc)], sc)
        outs = mctx.msm_batch([hb] * 3, [h.pack_scalars(sc) for sc in scs])
        for sc, r in zip(scs, outs):
            assert h.result_affine(r) == _expect(pts[: len(sc)], sc)
    finally:
        hb.release()


def test_sharded_2_18_matches_single_device(mctx):
    """1-GPU result == N-GPU result on device-generated inputs (cross-size consistency needs no oracle)."""
    n = 1 << 18
    one = b200msm.Context([0])
    try:
        d_bases = torch.empty(n * 64, dtype=torch.uint8, device="cuda:0")
        d_scalars = torch.empty(n * 32, dtype=torch.uint8, device="cuda:0")
        torch.cuda.synchronize()
        one.testkit_generate(0x5A4D, n, d_bases, d_scalars)
        hb = d_bases.cpu().numpy().view(np.uint64).reshape(n, 8)
        hs = d_scalars.cpu().numpy().view(np.uint64).reshape(n, 4)
        a = one.msm(hb, hs)
        b = mctx.msm(hb, hs)
        assert h.result_affine(a) == h.result_affine(b)
    finally:
        one.close()


def test_sharded_pageable_and_pinned_agree_2_20(mctx):
    """The per-device enqueue threads (pageable input staged for all devices at once) against one device, 2^20 points,
    from ordinary numpy memory and from pinned memory."""
    n = 1 << 20
    one = b200msm.Context([0])
    try:
        d_bases = torch.empty(n * 64, dtype=torch.uint8, device="cuda:0")
        d_scalars = torch.empty(n * 32, dtype=torch.uint8, device="cuda:0")
        torch.cuda.synchronize()
        one.testkit_generate(0x5A4E, n, d_bases, d_scalars)
        hb = np.zeros((n, 9), dtype=np.uint64)
        hb[:, :8] = d_bases.cpu().numpy().view(np.uint64).reshape(n, 8)
        hs = d_scalars.cpu().numpy().view(np.uint64).reshape(n, 4).copy()
        want = h.result_affine(one.msm(hb, hs))
        assert h.result_affine(mctx.msm(hb, hs)) == want                       # pageable
        pb, ps = torch.from_numpy(hb).pin_memory(), torch.from_numpy(hs).pin_memory()
        got = mctx.msm_raw(pb.data_ptr(), 72, 0, 32, 64, ps.data_ptr(), 32, n)  # pinned
        assert h.result_affine(got) == want
    finally:
        one.close()


def test_g2_sharded_matches_oracle_and_single_device(mctx):
    """b200msm_bn254_g2_msm shards by point range over the context's devices (>= 2^12 points per shard) and adds the
    192-byte partials on the first device."""
    import bn254_g2 as g2
    n = 1 << 13
    pts = g2.random_points(64, 9)
    rng = np.random.default_rng(4)
    idx = rng.integers(0, 64, n)
    idx[:64] = np.arange(64)
    base_rows = np.array([g2.encode_base(pt) for pt in pts], dtype=np.uint64)
    bases = np.ascontiguousarray(base_rows[idx])
    sc = o.random_scalars(n, 77)
    # oracle: group the scalars by base (the MSM is linear in the scalars of one base)
    per_base = [0] * 64
    for i, s in zip(idx, sc):
        per_base[int(i)] = (per_base[int(i)] + s) % o.R_ORDER
    want = g2.jac_to_affine(g2.msm_naive(pts, per_base))
    hs = h.pack_scalars(sc)
    got = mctx.msm_g2(bases, hs)
    assert g2.jac_to_affine(g2.decode_jacobian(got)) == want
    one = b200msm.Context([0])
    try:
        assert g2.jac_to_affine(g2.decode_jacobian(one.msm_g2(bases, hs))) == want
    finally:
        one.close()
