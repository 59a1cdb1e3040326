"""CPU suite: the C oracle (oracle/cpu_msm.c: arkworks `msm_bigint_wnaf` restated) against the
independent Python big-int oracle and the committed golden fixtures."""
import os
import random
import time

import numpy as np

import bn254 as o
import cpu_msm
import helpers as h

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "msm_cases.npz")


def _aff(words):
    return o.jac_to_affine(o.decode_jacobian(words))


def test_window_rule_matches_python():
    for n in (1, 31, 32, 1000, 1 << 16, 1 << 20, (1 << 20) + 1, 1 << 24, 1 << 26):
        assert cpu_msm.lib().oracle_ark_window(n) == o.ark_window_size(n)


def test_golden_fixtures_c_oracle():
    z = np.load(GOLDEN)
    for name in z["names"]:
        exp = z[f"{name}/expected"]
        want = None if int(exp[8]) else (h.unwords(exp[0:4]), h.unwords(exp[4:8]))
        for threads, w in ((1, 0), (4, 0), (3, 7), (8, 13)):
            out, _ = cpu_msm.msm(z[f"{name}/bases"], z[f"{name}/scalars"], threads, w)
            assert _aff(out) == want, (name, threads, w)


def test_c_oracle_matches_python_oracle_random():
    pts = o.random_points(3000, 91)
    sc = o.random_scalars(3000, 92)
    want = o.jac_to_affine(o.msm_pippenger(pts, sc, 9))
    out, used = cpu_msm.msm(h.pack_bases(pts), h.pack_scalars(sc))
    assert _aff(out) == want and 1 <= used <= 64


def test_dlog_checksum_and_scalar_mul():
    rng = random.Random(4)
    n = 10000
    t1 = np.array([h.words(rng.randrange(o.R_ORDER)) for _ in range(4096)], dtype=np.uint64)
    t2 = np.array([h.words(rng.randrange(o.R_ORDER)) for _ in range(3)], dtype=np.uint64)
    sc = [rng.randrange(o.R_ORDER) for _ in range(n)]
    got = h.unwords(cpu_msm.dlog_checksum(h.pack_scalars(sc), t1, t2))
    want = sum(s * (h.unwords(t1[i & 4095]) + h.unwords(t2[i >> 12])) for i, s in enumerate(sc)) % o.R_ORDER
    assert got == want
    k = rng.randrange(o.R_ORDER)
    assert _aff(cpu_msm.scalar_mul_gen(np.array(h.words(k), dtype=np.uint64))) == o.jac_to_affine(o.jac_scalar_mul(k, o.affine_to_jac(o.GEN)))
    a = cpu_msm.scalar_mul_gen(np.array(h.words(5), dtype=np.uint64))
    b = cpu_msm.scalar_mul_gen(np.array(h.words(7), dtype=np.uint64))
    assert _aff(cpu_msm.jac_add(a, b)) == o.jac_to_affine(o.jac_scalar_mul(12, o.affine_to_jac(o.GEN)))


def test_reference_config_2_16_cpu():
    """BASELINE config #0: 2^16 random bases/scalars on the host CPU (the reference's own test size,
    tests/cuzk/e2e.rs:14-63).  Bases are multiples of G with known logs so the expected value is
    independent of the MSM code under test."""
    n = 1 << 16
    rng = random.Random(16)
    t1 = np.array([h.words(rng.randrange(1, o.R_ORDER)) for _ in range(4096)], dtype=np.uint64)
    t2 = np.array([h.words(rng.randrange(1, o.R_ORDER)) for _ in range(n >> 12)], dtype=np.uint64)
    tab1 = [cpu_msm.scalar_mul_gen(t) for t in t1[:64]]  # 64 x 16 distinct sums are enough: reuse rows
    tab2 = [cpu_msm.scalar_mul_gen(t) for t in t2]
    pts = []
    for i in range(64 * 16):
        pts.append(o.jac_to_affine(o.decode_jacobian(cpu_msm.jac_add(tab1[i & 63], tab2[i >> 6]))))
    bases = np.tile(h.pack_bases(pts), (n // len(pts), 1))
    scal = np.frombuffer(np.random.default_rng(5).bytes(n * 32), dtype=np.uint64).reshape(n, 4).copy()
    scal[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)  # < r
    t0 = time.time()
    out, used = cpu_msm.msm(bases, scal)
    dt = time.time() - t0
    # expected via logs: base i has log t1[(i%1024)&63] + t2[(i%1024)>>6]
    sm = [h.unwords(r) for r in scal]
    acc = 0
    for i, s in enumerate(sm):
        j = i % 1024
        acc += s * (h.unwords(t1[j & 63]) + h.unwords(t2[j >> 6]))
    dlog = acc * o.RINV_R % o.R_ORDER
    assert _aff(out) == o.jac_to_affine(o.jac_scalar_mul(dlog, o.affine_to_jac(o.GEN)))
    print(f"cpu oracle 2^16: {dt*1e3:.0f} ms on {used} threads")
