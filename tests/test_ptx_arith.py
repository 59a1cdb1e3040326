"""CPU suite: execute the exact PTX instruction stream of the production field arithmetic
(gpu-acceleration_b200/csrc/gen_fq_asm.py -> fq_asm.inc) in an interpreter and compare with the
oracle.  Mirrors the reference's limb/field/Montgomery test levels (tests/bigint, tests/field,
tests/mont_backend/mont_mul_cios.rs:19-76) without needing a GPU."""
import os
import random
import sys

import bn254 as o
import ptx_sim

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "gpu-acceleration_b200", "csrc"))
import gen_fq_asm as g  # noqa: E402

EDGE = [0, 1, 2, o.P - 1, o.P - 2, o.R_MOD_P, o.R2_MOD_P, (1 << 253), (1 << 32) - 1, (1 << 64), o.P >> 1]


def _cases(k=150):
    rng = random.Random(99)
    cs = [(a, b) for a in EDGE for b in EDGE]
    cs += [(rng.randrange(o.P), rng.randrange(o.P)) for _ in range(k)]
    return cs


def test_mont_mul_stream():
    body = g.mul_body()
    for a, b in _cases():
        assert ptx_sim.call(body, a, b) == o.mont_mul(a, b)


def test_mont_sqr_stream():
    """Dedicated squaring (108 wide MACs: off-diagonal once, doubled, + diagonal) == mul(a, a)."""
    body = g.sqr_body()
    rng = random.Random(7)
    for a in EDGE + [rng.randrange(o.P) for _ in range(300)]:
        assert ptx_sim.call(body, a) == o.mont_mul(a, a)


def test_mont_mulsub_stream():
    """Fused a*b - c*d with one reduction (200 wide MACs): equals mont_mul(a, b) - mont_mul(c, d) mod p, including
    c = 0 (nc = p), a*b = c*d (result 0) and the largest operands (the < 2p bound before the single subtraction)."""
    body = g.mulsub_body()
    rng = random.Random(17)
    quads = [(a, b, c, d) for a in EDGE[:6] for b in EDGE[:6] for c in EDGE[:6] for d in EDGE[:6]]
    quads += [(o.P - 1, o.P - 1, 1, 1), (o.P - 1, o.P - 1, 0, 0), (o.P - 1, o.P - 1, o.P - 1, 0), (5, 7, 7, 5), (0, 0, o.P - 1, o.P - 1)]
    quads += [tuple(rng.randrange(o.P) for _ in range(4)) for _ in range(400)]
    for a, b, c, d in quads:
        assert ptx_sim.call(body, a, b, c, d) == (o.mont_mul(a, b) - o.mont_mul(c, d)) % o.P, (a, b, c, d)


def test_mont_muladd_stream():
    """Fused a*b + c*d with one reduction (200 wide MACs): equals mont_mul(a, b) + mont_mul(c, d) mod p, including the largest
    operands (the < 2p bound before the single subtraction) and sums that land exactly on p."""
    body = g.muladd_body()
    rng = random.Random(23)
    quads = [(a, b, c, d) for a in EDGE[:6] for b in EDGE[:6] for c in EDGE[:6] for d in EDGE[:6]]
    quads += [(o.P - 1, o.P - 1, o.P - 1, o.P - 1), (o.P - 1, o.P - 1, 0, 0), (0, 0, o.P - 1, o.P - 1), (5, 7, o.P - 5, 7), (0, 0, 0, 0)]
    quads += [tuple(rng.randrange(o.P) for _ in range(4)) for _ in range(400)]
    for a, b, c, d in quads:
        assert ptx_sim.call(body, a, b, c, d) == (o.mont_mul(a, b) + o.mont_mul(c, d)) % o.P, (a, b, c, d)


def test_add_sub_streams():
    add, sub = g.add_body(), g.sub_body()
    for a, b in _cases():
        assert ptx_sim.call(add, a, b) == (a + b) % o.P  # overflow + a+b in [p, 2p) cases: tests/field/ff_reduce.rs:84-114
        assert ptx_sim.call(sub, a, b) == (a - b) % o.P  # underflow: tests/bigint/bigint_sub.rs


def test_generated_file_is_current():
    import io
    import contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        g.main()
    path = os.path.join(os.path.dirname(g.__file__), "fq_asm.inc")
    assert open(path).read() == buf.getvalue(), "fq_asm.inc is stale: re-run gen_fq_asm.py"


def test_constants_in_stream():
    assert g.N0 == o.N0_32_P and g.P == o.P
