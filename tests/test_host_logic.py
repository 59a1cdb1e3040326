"""CPU suite: host-side logic of the C library that needs no device -- the slice plan of the host-buffer call and the
parallel copy that stages pageable memory into the pinned upload ring."""
import ctypes as C

import numpy as np
import pytest

import b200msm


def _plan(lib, n, slices, ratio):
    begins = (C.c_size_t * 8)()
    lens = (C.c_size_t * 8)()
    cnt = C.c_int()
    rc = lib.b200msm_testkit_slice_plan(n, slices, ratio, begins, lens, C.byref(cnt))
    assert rc == 0
    return [(begins[k], lens[k]) for k in range(cnt.value)]


def test_slice_plan_covers_the_range_and_grows():
    lib = b200msm.load_library()
    for n in (1, 2, 3, 7, 1000, 4099, 1 << 20, (1 << 24) + 5):
        for slices in range(1, 9):
            for ratio in (100, 130, 160, 400):
                pl = _plan(lib, n, slices, ratio)
                assert 1 <= len(pl) <= min(slices, n)
                pos = 0
                for b, ln in pl:               # contiguous, non-empty, in order
                    assert b == pos and ln >= 1
                    pos += ln
                assert pos == n
                if n >= (1 << 20) and len(pl) == slices:
                    for (_, a), (_, c) in zip(pl, pl[1:]):   # geometric growth within rounding
                        assert abs(c / a - ratio / 100) < 0.01
    assert _plan(lib, 1 << 20, 3, 160)[0][1] == round((1 << 20) / (1 + 1.6 + 2.56))
    # argument validation
    cnt = C.c_int()
    arr = (C.c_size_t * 8)()
    assert lib.b200msm_testkit_slice_plan(0, 3, 160, arr, arr, C.byref(cnt)) != 0
    assert lib.b200msm_testkit_slice_plan(10, 9, 160, arr, arr, C.byref(cnt)) != 0
    assert lib.b200msm_testkit_slice_plan(10, 3, 50, arr, arr, C.byref(cnt)) != 0


@pytest.mark.parametrize("threads", [1, 2, 4, 7])
def test_parallel_copy_is_a_memcpy(threads):
    lib = b200msm.load_library()
    rng = np.random.default_rng(threads)
    for nbytes in (0, 1, 4095, 4096, (1 << 20) - 1, 1 << 20, (1 << 20) + 1, 3 * (1 << 20) + 12345, 8 << 20):
        src = rng.integers(0, 256, size=nbytes + 64, dtype=np.uint8)
        dst = np.full(nbytes + 64, 0xAB, dtype=np.uint8)
        # unaligned source and destination on purpose
        assert lib.b200msm_testkit_parallel_copy(dst.ctypes.data + 3, src.ctypes.data + 5, nbytes, threads) == 0
        assert np.array_equal(dst[3:3 + nbytes], src[5:5 + nbytes])
        assert np.all(dst[:3] == 0xAB) and np.all(dst[3 + nbytes:] == 0xAB)   # nothing written outside the range
