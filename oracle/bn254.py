"""CPU oracle for the BN254 G1 variable-base MSM path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (gpu-acceleration_b200/) never does.

What it restates
----------------
The observable behaviour of the reference's hot path
    metal_variable_base_msm(&[G1Affine], &[Fr]) -> Result<G1Projective>
        /root/reference/mopro-msm/src/msm/metal_msm/metal_msm.rs:642-695
whose contract is "equals arkworks `G1Projective::msm(bases, scalars)`"
        /root/reference/mopro-msm/src/msm/metal_msm/tests/cuzk/e2e.rs:50-61
        /root/reference/mopro-msm/src/msm/metal_msm/metal_msm.rs:739-760
plus stage-level restatements of the cuZK pipeline (signed digits, CSC, bucket slots,
running-sum reduce, Horner) following SURVEY.md Appendix A, which cites
        shader/cuzk/convert_point_coords_and_decompose_scalars.metal:95-121  (signed digits)
        shader/cuzk/transpose.metal:8-65                                     (stable CSC)
        shader/cuzk/smvp.metal:14-107                                        (bucket slots / signs)
        shader/cuzk/pbpr.metal:33-148                                        (two-pass bucket reduce)
        metal_msm.rs:204-262                                                 (Horner)

Third-party arithmetic that is NOT vendored under /root/reference: arkworks
`ark-ec 0.4.1`, `ark-ff 0.4.1`, `ark-bn254 0.4.0` (Cargo.toml:25-35, Cargo.lock).  Their
published algorithm is restated here: Fp = MontBackend<_, 4> holding a*R mod p with
R = 2^256 in four little-endian u64 limbs; G1: y^2 = x^3 + 3, generator (1, 2);
G1Projective is Jacobian (X/Z^2, Y/Z^3), infinity <=> Z == 0.

Pinning ("parity pinned by constants, unpinned by stored MSM vectors")
---------------------------------------------------------------------
The reference holds NO golden MSM result vectors (every test draws fresh random inputs and
compares with arkworks in the same process; arkworks cannot be built here: no Rust).
What the reference DOES hold as known-answer literals is checked by `self_check()`:
    p            shader/constants.metal:30-47   (BN254_BASEFIELD_MODULUS)
    R mod p      shader/constants.metal:229-246 (BN254_ONE_XR)
    n0 (16-bit)  utils/mont_params.rs:122       (25481)
    R^-1 mod p   utils/mont_params.rs:116-121
    Barrett mu   utils/barrett_params.rs:25-28
    r            utils/mont_params.rs:9
The MSM value sum_i s_i*P_i is a unique group element, so any correct implementation
agrees with arkworks after normalisation; tests compare affine-normalised points.
Reference-side DATA that pins the memory layout: tests/golden/ref_srs_g1.npz, 80 BN254 G1 points extracted
from /root/reference/example-app/ios/{plonk,gemini,hyperplonk}_fibonacci_srs.bin (raw Montgomery x || y words;
tests/golden/make_ref_srs_fixture.py).  Every one decodes to a curve point under the conventions below and the
first of each file is the generator (tests/test_oracle.py::test_reference_srs_points_pin_memory_layout).
"""
from __future__ import annotations

import hashlib
import struct
from typing import Iterable, List, Optional, Sequence, Tuple

# ----------------------------------------------------------------------------- constants
P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # base field
R_ORDER = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # scalar field
MONT_R = 1 << 256
B_COEFF = 3
GEN = (1, 2)

# known-answer literals held by the reference (see module docstring for file:line)
_REF_R_MOD_P = 0x0E0A77C19A07DF2F666EA36F7879462C0A78EB28F5C70B3DD35D438DC58F0D9D
_REF_RINV = 20988524275117001072002809824448087578619730785600314334253784976379291040311
_REF_N0_16 = 25481
_REF_BARRETT_MU = 38284845454613504619394467267190322316455732053192006567598327834621704638693

R_MOD_P = MONT_R % P
R2_MOD_P = (MONT_R * MONT_R) % P
RINV_P = pow(MONT_R, -1, P)
N0_32_P = (-pow(P, -1, 1 << 32)) % (1 << 32)
N0_64_P = (-pow(P, -1, 1 << 64)) % (1 << 64)
R_MOD_R = MONT_R % R_ORDER
R2_MOD_R = (MONT_R * MONT_R) % R_ORDER
RINV_R = pow(MONT_R, -1, R_ORDER)
N0_32_R = (-pow(R_ORDER, -1, 1 << 32)) % (1 << 32)


def self_check() -> None:
    """Pin the oracle's constants against every known-answer literal in the reference."""
    assert R_MOD_P == _REF_R_MOD_P
    assert RINV_P == _REF_RINV
    assert (-pow(P, -1, 1 << 16)) % (1 << 16) == _REF_N0_16
    assert N0_32_P == 0xE4866389 and N0_64_P == 0x87D20782E4866389
    assert (1 << 508) // P == _REF_BARRETT_MU  # mu = floor(2^(2*254)/p), barrett_params.rs:3-7
    assert (2 * R_MOD_P) % P == to_mont(2)  # generator y = 2 in Montgomery form (constants.metal:247-264)
    assert is_on_curve(GEN)
    assert jac_is_inf(jac_scalar_mul(R_ORDER, affine_to_jac(GEN)))  # r*G = inf pins r
    assert N0_32_R == 0xEFFFFFFF


# ----------------------------------------------------------------------------- field helpers
def to_mont(a: int, mod: int = P) -> int:
    return (a * MONT_R) % mod


def from_mont(a: int, mod: int = P) -> int:
    return (a * (RINV_P if mod == P else RINV_R)) % mod


def mont_mul(a: int, b: int, mod: int = P) -> int:
    """montmul(aR, bR) = abR mod p   (mont_backend/mont.metal:105-181 semantics, R = 2^256)."""
    return (a * b * (RINV_P if mod == P else RINV_R)) % mod


def mont_mul_cios32(a: int, b: int, mod: int = P) -> int:
    """Word-level CIOS with 8x32-bit limbs; the shape the CUDA kernels use.  Output in [0, mod)."""
    n0 = N0_32_P if mod == P else N0_32_R
    mask = (1 << 32) - 1
    t = 0
    for i in range(8):
        t += a * ((b >> (32 * i)) & mask)
        m = ((t & mask) * n0) & mask
        t += m * mod
        assert t & mask == 0
        t >>= 32
    if t >= mod:
        t -= mod
    return t


def limbs_u64(a: int) -> Tuple[int, int, int, int]:
    return tuple((a >> (64 * i)) & ((1 << 64) - 1) for i in range(4))  # type: ignore


def to_bytes32(a: int) -> bytes:
    return a.to_bytes(32, "little")


def from_bytes32(b: bytes) -> int:
    return int.from_bytes(b, "little")


def fq_sqrt(a: int) -> Optional[int]:
    """p = 3 mod 4, so sqrt is one exponentiation."""
    y = pow(a, (P + 1) // 4, P)
    return y if (y * y) % P == a % P else None


# ----------------------------------------------------------------------------- curve (affine = (x, y) or None)
Affine = Optional[Tuple[int, int]]
Jac = Tuple[int, int, int]
JAC_INF: Jac = (1, 1, 0)


def is_on_curve(pt: Affine) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B_COEFF) % P == 0


def affine_neg(pt: Affine) -> Affine:
    return None if pt is None else (pt[0], (-pt[1]) % P)


def affine_to_jac(pt: Affine) -> Jac:
    return JAC_INF if pt is None else (pt[0], pt[1], 1)


def jac_is_inf(a: Jac) -> bool:
    return a[2] % P == 0


def jac_to_affine(a: Jac) -> Affine:
    if jac_is_inf(a):
        return None
    zi = pow(a[2], -1, P)
    zi2 = zi * zi % P
    return (a[0] * zi2 % P, a[1] * zi2 * zi % P)


def jac_dbl(a: Jac) -> Jac:
    """EFD dbl-2009-l (a = 0)  -- curve/jacobian.metal:11-44."""
    X1, Y1, Z1 = a
    if Z1 % P == 0:
        return JAC_INF
    A = X1 * X1 % P
    B = Y1 * Y1 % P
    C = B * B % P
    D = 2 * ((X1 + B) * (X1 + B) - A - C) % P
    E = 3 * A % P
    F = E * E % P
    X3 = (F - 2 * D) % P
    Y3 = (E * (D - X3) - 8 * C) % P
    Z3 = 2 * Y1 * Z1 % P
    return (X3, Y3, Z3)


def jac_add(a: Jac, b: Jac) -> Jac:
    """EFD add-2007-bl with COMPLETE handling of inf / P+P / P+(-P)
    (curve/jacobian.metal:46-100 handles P+P only by limb equality -- SURVEY §2.3 gap 2;
    arkworks handles it by cross-multiplied comparison, which is what this does)."""
    if jac_is_inf(a):
        return b
    if jac_is_inf(b):
        return a
    X1, Y1, Z1 = a
    X2, Y2, Z2 = b
    Z1Z1 = Z1 * Z1 % P
    Z2Z2 = Z2 * Z2 % P
    U1 = X1 * Z2Z2 % P
    U2 = X2 * Z1Z1 % P
    S1 = Y1 * Z2 * Z2Z2 % P
    S2 = Y2 * Z1 * Z1Z1 % P
    if U1 == U2:
        return jac_dbl(a) if S1 == S2 else JAC_INF
    H = (U2 - U1) % P
    I = 4 * H * H % P
    J = H * I % P
    r = 2 * (S2 - S1) % P
    V = U1 * I % P
    X3 = (r * r - J - 2 * V) % P
    Y3 = (r * (V - X3) - 2 * S1 * J) % P
    Z3 = ((Z1 + Z2) * (Z1 + Z2) - Z1Z1 - Z2Z2) * H % P
    return (X3, Y3, Z3)


def jac_neg(a: Jac) -> Jac:
    return (a[0], (-a[1]) % P, a[2])


def jac_add_affine(a: Jac, pt: Affine) -> Jac:
    return a if pt is None else jac_add(a, (pt[0], pt[1], 1))


def jac_scalar_mul(k: int, a: Jac) -> Jac:
    acc = JAC_INF
    for bit in bin(k)[2:] if k else "":
        acc = jac_dbl(acc)
        if bit == "1":
            acc = jac_add(acc, a)
    return acc


def jac_eq(a: Jac, b: Jac) -> bool:
    """arkworks `Projective == Projective`: cross-multiplied equality."""
    if jac_is_inf(a) or jac_is_inf(b):
        return jac_is_inf(a) and jac_is_inf(b)
    z1z1, z2z2 = a[2] * a[2] % P, b[2] * b[2] % P
    return (a[0] * z2z2 - b[0] * z1z1) % P == 0 and (a[1] * z2z2 * b[2] - b[1] * z1z1 * a[2]) % P == 0


# XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2) -- what the CUDA bucket kernels use.
XYZZ = Tuple[int, int, int, int]
XYZZ_INF: XYZZ = (0, 0, 0, 0)


def xyzz_to_affine(a: XYZZ) -> Affine:
    if a[2] % P == 0:
        return None
    return (a[0] * pow(a[2], -1, P) % P, a[1] * pow(a[3], -1, P) % P)


def xyzz_madd(a: XYZZ, pt: Affine) -> XYZZ:
    """EFD madd-2008-s, complete."""
    if pt is None:
        return a
    if a[2] % P == 0:
        return (pt[0], pt[1], 1, 1)
    X1, Y1, ZZ1, ZZZ1 = a
    U2 = pt[0] * ZZ1 % P
    S2 = pt[1] * ZZZ1 % P
    Pp = (U2 - X1) % P
    Rr = (S2 - Y1) % P
    if Pp == 0:
        if Rr == 0:
            return xyzz_dbl_affine(pt)
        return XYZZ_INF
    PP = Pp * Pp % P
    PPP = Pp * PP % P
    Q = X1 * PP % P
    X3 = (Rr * Rr - PPP - 2 * Q) % P
    Y3 = (Rr * (Q - X3) - Y1 * PPP) % P
    return (X3, Y3, ZZ1 * PP % P, ZZZ1 * PPP % P)


def xyzz_dbl_affine(pt: Affine) -> XYZZ:
    """EFD mdbl-2008-s-1 (a = 0)."""
    if pt is None:
        return XYZZ_INF
    X1, Y1 = pt
    U = 2 * Y1 % P
    V = U * U % P
    W = U * V % P
    S = X1 * V % P
    M = 3 * X1 * X1 % P
    X3 = (M * M - 2 * S) % P
    Y3 = (M * (S - X3) - W * Y1) % P
    return (X3, Y3, V, W)


def xyzz_dbl(a: XYZZ) -> XYZZ:
    """EFD dbl-2008-s-1 (a = 0)."""
    if a[2] % P == 0:
        return XYZZ_INF
    X1, Y1, ZZ1, ZZZ1 = a
    U = 2 * Y1 % P
    V = U * U % P
    W = U * V % P
    S = X1 * V % P
    M = 3 * X1 * X1 % P
    X3 = (M * M - 2 * S) % P
    Y3 = (M * (S - X3) - W * Y1) % P
    return (X3, Y3, V * ZZ1 % P, W * ZZZ1 % P)


def xyzz_add(a: XYZZ, b: XYZZ) -> XYZZ:
    """EFD add-2008-s, complete."""
    if a[2] % P == 0:
        return b
    if b[2] % P == 0:
        return a
    X1, Y1, ZZ1, ZZZ1 = a
    X2, Y2, ZZ2, ZZZ2 = b
    U1 = X1 * ZZ2 % P
    U2 = X2 * ZZ1 % P
    S1 = Y1 * ZZZ2 % P
    S2 = Y2 * ZZZ1 % P
    Pp = (U2 - U1) % P
    Rr = (S2 - S1) % P
    if Pp == 0:
        return xyzz_dbl(a) if Rr == 0 else XYZZ_INF
    PP = Pp * Pp % P
    PPP = Pp * PP % P
    Q = U1 * PP % P
    X3 = (Rr * Rr - PPP - 2 * Q) % P
    Y3 = (Rr * (Q - X3) - S1 * PPP) % P
    return (X3, Y3, ZZ1 * ZZ2 % P * PP % P, ZZZ1 * ZZZ2 % P * PPP % P)


def xyzz_to_jac(a: XYZZ) -> Jac:
    """(X*ZZ^2, Y*ZZ^3, ZZZ) is a Jacobian representative of the same point (lambda = ZZ)."""
    if a[2] % P == 0:
        return JAC_INF
    zz2 = a[2] * a[2] % P
    return (a[0] * zz2 % P, a[1] * zz2 * a[2] % P, a[3])


# ----------------------------------------------------------------------------- MSM
def msm_naive(bases: Sequence[Affine], scalars: Sequence[int]) -> Jac:
    """sum_i s_i * P_i by double-and-add.  The definition; for tiny n."""
    acc = JAC_INF
    for pt, s in zip(bases, scalars):
        acc = jac_add(acc, jac_scalar_mul(s % R_ORDER, affine_to_jac(pt)))
    return acc


def ark_window_size(n: int) -> int:
    """ark-ec 0.4.1 `msm_bigint_wnaf`: c = 3 if n < 32 else ln_without_floats(n) + 2,
    ln_without_floats(n) = log2(n) * 69 / 100 (integer; log2 = ceil-ish `ark_std::log2`)."""
    if n < 32:
        return 3
    log2 = (n - 1).bit_length()  # ark_std::log2(n) = ceil(log2 n)
    return log2 * 69 // 100 + 2


def signed_digits(s: int, w: int, num_windows: Optional[int] = None) -> List[int]:
    """Signed radix-2^w digits in [-2^(w-1), 2^(w-1)) -- convert kernel :95-121 (the reference
    stores digit + 2^(w-1)); the final carry is appended as an extra window if it is non-zero
    (the reference silently drops it, SURVEY §2.3 gap 4)."""
    L = 1 << w
    half = L >> 1
    K = num_windows if num_windows is not None else num_windows_for(w)
    out = []
    carry = 0
    for k in range(K):
        v = ((s >> (k * w)) & (L - 1)) + carry
        if v >= half:
            v -= L
            carry = 1
        else:
            carry = 0
        out.append(v)
    assert carry == 0, "window count too small for this scalar"
    return out


def num_windows_for(w: int, bits: int = 254) -> int:
    K = -(-bits // w)
    # top window must have head-room for the carry (scalars < r < 2^254)
    if bits - w * (K - 1) == w:
        K += 1
    return K


def msm_pippenger(bases: Sequence[Affine], scalars: Sequence[int], w: Optional[int] = None) -> Jac:
    """Signed-bucket Pippenger (the algorithm of ark-ec 0.4.1 `msm_bigint_wnaf`, also the
    cuZK pipeline's mathematical content: SURVEY Appendix A).  Buckets are XYZZ here only
    for speed; the result is a group element and representation-independent."""
    n = min(len(bases), len(scalars))
    if w is None:
        w = ark_window_size(n)
    K = num_windows_for(w)
    half = 1 << (w - 1)
    window_sums: List[XYZZ] = []
    digs = [signed_digits(scalars[i] % R_ORDER, w, K) for i in range(n)]
    for k in range(K):
        buckets: dict = {}
        for i in range(n):
            d = digs[i][k]
            if d == 0 or bases[i] is None:
                continue
            pt = bases[i] if d > 0 else affine_neg(bases[i])
            m = abs(d)
            buckets[m] = xyzz_madd(buckets.get(m, XYZZ_INF), pt)
        # running-sum reduce: sum_m m * B[m]
        running = XYZZ_INF
        total = XYZZ_INF
        if buckets:
            keys = sorted(buckets.keys(), reverse=True)
            prev = keys[0]
            for m in keys:
                # gap of (prev - m) steps where running is unchanged
                gap = prev - m
                if gap:
                    total = xyzz_add(total, _xyzz_small_mul(gap, running))
                running = xyzz_add(running, buckets[m])
                prev = m
            total = xyzz_add(total, _xyzz_small_mul(prev, running))
        window_sums.append(total)
    acc = JAC_INF
    for k in reversed(range(K)):
        for _ in range(w):
            acc = jac_dbl(acc)
        acc = jac_add(acc, xyzz_to_jac(window_sums[k]))
    return acc


def _xyzz_small_mul(k: int, a: XYZZ) -> XYZZ:
    acc = XYZZ_INF
    for bit in bin(k)[2:] if k else "":
        acc = xyzz_dbl(acc)
        if bit == "1":
            acc = xyzz_add(acc, a)
    return acc


# ----------------------------------------------------------------------------- stage-level restatements
def stage_rows(scalars: Sequence[int], w: int) -> List[List[int]]:
    """rows[k][i] = digit + half (what `chunks[k*n + i]` holds in the reference)."""
    K = num_windows_for(w)
    half = 1 << (w - 1)
    digs = [signed_digits(s, w, K) for s in scalars]
    return [[digs[i][k] + half for i in range(len(scalars))] for k in range(K)]


def stage_csc(row: Sequence[int], num_cols: int) -> Tuple[List[int], List[int]]:
    """Stable counting sort of one window (transpose.metal:8-65; tests/cuzk/transpose.rs:95-118)."""
    col_ptr = [0] * (num_cols + 1)
    for d in row:
        col_ptr[d + 1] += 1
    for d in range(num_cols):
        col_ptr[d + 1] += col_ptr[d]
    cur = list(col_ptr[:-1])
    val_idx = [0] * len(row)
    for i, d in enumerate(row):
        val_idx[cur[d]] = i
        cur[d] += 1
    return col_ptr, val_idx


def stage_bucket_sums(bases: Sequence[Affine], digits: Sequence[int], half: int) -> List[XYZZ]:
    """bucket[m] = sum_{digit=+m} P - sum_{digit=-m} P for m in 1..half (index m; slot 0 unused).
    Same content as smvp.metal:14-107 (which stores magnitude `half` in slot 0)."""
    out = [XYZZ_INF] * (half + 1)
    for pt, d in zip(bases, digits):
        if d == 0 or pt is None:
            continue
        out[abs(d)] = xyzz_madd(out[abs(d)], pt if d > 0 else affine_neg(pt))
    return out


def stage_bucket_reduce(buckets: Sequence[XYZZ]) -> XYZZ:
    """sum_m m * bucket[m]  (pbpr.metal:33-148 + the CPU sum in metal_msm.rs:214-247)."""
    running = XYZZ_INF
    total = XYZZ_INF
    for m in range(len(buckets) - 1, 0, -1):
        running = xyzz_add(running, buckets[m])
        total = xyzz_add(total, running)
    return total


def table_expand(bases: Sequence[Affine], w: int) -> List[List[Affine]]:
    """Precomputed-table mode (engine-internal): table[k][i] = 2^(w*k) * P_i in affine form for every
    digit window k.  Infinity stays infinity."""
    K = num_windows_for(w)
    rows = [list(bases)]
    for _ in range(1, K):
        prev = rows[-1]
        nxt = []
        for pt in prev:
            if pt is None:
                nxt.append(None)
                continue
            j = affine_to_jac(pt)
            for _ in range(w):
                j = jac_dbl(j)
            nxt.append(jac_to_affine(j))
        rows.append(nxt)
    return rows


def msm_table_mode(bases: Sequence[Affine], scalars: Sequence[int], w: int) -> Jac:
    """The MSM as the engine computes it over a precomputed table: every signed digit d of window k of scalar i sends
    table[k][i] (negated for d < 0) to bucket |d| of ONE bucket set; the result is sum_m m * bucket[m] -- no Horner step."""
    K = num_windows_for(w)
    half = 1 << (w - 1)
    table = table_expand(bases, w)
    buckets = [XYZZ_INF] * (half + 1)
    for i, s in enumerate(scalars):
        for k, d in enumerate(signed_digits(s % R_ORDER, w, K)):
            pt = table[k][i]
            if d == 0 or pt is None:
                continue
            buckets[abs(d)] = xyzz_madd(buckets[abs(d)], pt if d > 0 else affine_neg(pt))
    return xyzz_to_jac(stage_bucket_reduce(buckets))


# ----------------------------------------------------------------------------- deterministic inputs & serialisation
def _prng_words(seed: int, count: int, tag: bytes) -> Iterable[int]:
    ctr = 0
    while ctr < count:
        h = hashlib.sha256(tag + struct.pack("<QQ", seed, ctr)).digest()
        yield int.from_bytes(h, "little")
        ctr += 1


def random_scalars(n: int, seed: int) -> List[int]:
    """Uniform-ish in [0, r): 256-bit hash reduced mod r (bias 2^-2, irrelevant for tests)."""
    return [v % R_ORDER for v in _prng_words(seed, n, b"scalar")]


def random_points(n: int, seed: int) -> List[Affine]:
    """n distinct-looking G1 points: try-and-increment on x (cofactor 1, so every curve point is in G1)."""
    out: List[Affine] = []
    for v in _prng_words(seed, n, b"point"):
        x = v % P
        while True:
            y = fq_sqrt((x * x * x + B_COEFF) % P)
            if y is not None:
                break
            x = (x + 1) % P
        if (v >> 255) & 1:
            y = (-y) % P
        out.append((x, y))
    return out


def encode_bases(bases: Sequence[Affine], stride: int = 64, x_off: int = 0, y_off: int = 32,
                 inf_off: Optional[int] = None) -> bytes:
    """Raw arkworks memory: Montgomery-form x, y as LE u64[4]; optional infinity byte.
    Infinity is (0, 0, true) as `Affine::identity()` builds it."""
    buf = bytearray(stride * len(bases))
    for i, pt in enumerate(bases):
        o = i * stride
        if pt is None:
            if inf_off is not None:
                buf[o + inf_off] = 1
            continue
        buf[o + x_off:o + x_off + 32] = to_bytes32(to_mont(pt[0]))
        buf[o + y_off:o + y_off + 32] = to_bytes32(to_mont(pt[1]))
    return bytes(buf)


def encode_scalars(scalars: Sequence[int]) -> bytes:
    """`&[Fr]` memory: s*R mod r, LE."""
    return b"".join(to_bytes32(to_mont(s % R_ORDER, R_ORDER)) for s in scalars)


def decode_jacobian(words: Sequence[int]) -> Jac:
    """12 u64 words (X, Y, Z Montgomery LE) -> canonical Jacobian ints."""
    vals = []
    for c in range(3):
        v = 0
        for j in range(4):
            v |= int(words[4 * c + j]) << (64 * j)
        assert v < P, "coordinate not fully reduced"
        vals.append(from_mont(v))
    return (vals[0], vals[1], vals[2])


if __name__ == "__main__":
    self_check()
    pts = random_points(40, 1)
    sc = random_scalars(40, 2)
    a = jac_to_affine(msm_naive(pts, sc))
    for w in (3, 5, 8, 13, 16):
        assert jac_to_affine(msm_pippenger(pts, sc, w)) == a, w
    print("oracle self-check ok")


# ----------------------------------------------------------------------------- GLV endomorphism (engine-internal)
# BN254 G1 has the endomorphism phi(x, y) = (BETA*x, y) = LAMBDA*(x, y).  The CUDA engine splits every scalar
# s = k1 + k2*LAMBDA (mod r) with |k1|, |k2| < 2^127 so that the 254-bit MSM over n points becomes a 127-bit
# MSM over 2n points: same number of bucket additions, half the windows for the latency-bound reduce stage.
# arkworks does the same for single scalar multiplications (ark-ec `GLVConfig`, ark-bn254 g1.rs); the MSM
# result is unchanged as a group element.  This block restates the engine's EXACT integer procedure
# (shift-based rounding with 256 fractional bits) so that tests can compare digit-for-digit.
GLV_LAMBDA = 0xB3C4D79D41A917585BFC41088D8DAAA78B17EA66B99C90DD
GLV_BETA = 0x59E26BCEA0D48BACD4F263F1ACDB5C4F5763473177FFFFFE
# short lattice basis of {(a, b): a + b*LAMBDA = 0 mod r}
GLV_A1, GLV_B1 = 9931322734385697763, -147946756881789319000765030803803410728
GLV_A2, GLV_B2 = 147946756881789319010696353538189108491, 9931322734385697763
GLV_G1 = (GLV_B2 << 256) // R_ORDER          # round(s*b2/r)  ~ (s*G1 + 2^255) >> 256
GLV_G2 = ((-GLV_B1) << 256) // R_ORDER       # round(-s*b1/r) ~ (s*G2 + 2^255) >> 256


def glv_self_check() -> None:
    assert (GLV_LAMBDA * GLV_LAMBDA + GLV_LAMBDA + 1) % R_ORDER == 0
    assert pow(GLV_BETA, 3, P) == 1 and GLV_BETA != 1
    assert (GLV_A1 + GLV_B1 * GLV_LAMBDA) % R_ORDER == 0 and (GLV_A2 + GLV_B2 * GLV_LAMBDA) % R_ORDER == 0
    assert GLV_A1 * GLV_B2 - GLV_A2 * GLV_B1 == R_ORDER
    q = jac_to_affine(jac_scalar_mul(GLV_LAMBDA, affine_to_jac(GEN)))
    assert q == (GLV_BETA * GEN[0] % P, GEN[1])


def glv_decompose(s: int) -> Tuple[int, int]:
    """s in [0, r) -> (k1, k2) signed, k1 + k2*LAMBDA = s (mod r), |k1|, |k2| < 2^127.
    Arithmetic is done modulo 2^256 exactly as on the device (wrap-around, then sign from bit 255)."""
    M = (1 << 256) - 1
    c1 = (s * GLV_G1 + (1 << 255)) >> 256
    c2 = (s * GLV_G2 + (1 << 255)) >> 256
    k1 = (s - c1 * GLV_A1 - c2 * GLV_A2) & M
    k2 = (c1 * (-GLV_B1) - c2 * GLV_B2) & M
    if k1 >> 255:
        k1 -= 1 << 256
    if k2 >> 255:
        k2 -= 1 << 256
    return k1, k2


def glv_expand(bases: Sequence[Affine], scalars: Sequence[int]) -> Tuple[List[Affine], List[int]]:
    """The 2n-term problem the engine actually accumulates: pseudo-point i < n is P_i with k1_i, pseudo-point
    n + i is phi(P_i) with k2_i (signs kept on the scalars)."""
    n = min(len(bases), len(scalars))
    pts: List[Affine] = list(bases[:n]) + [None if b is None else (GLV_BETA * b[0] % P, b[1]) for b in bases[:n]]
    ks = [glv_decompose(s % R_ORDER) for s in scalars[:n]]
    return pts, [k[0] for k in ks] + [k[1] for k in ks]


def signed_digits_signed(k: int, w: int, num_windows: int) -> List[int]:
    """Signed-digit recoding of a SIGNED value, every digit in [-2^(w-1), 2^(w-1)) (the engine stores int16
    digits at w = 16): k >= 0 uses the reference's rule (v >= half wraps); k < 0 recodes |k| with the mirrored
    rule (v > half wraps) and flips every sign."""
    if k >= 0:
        return signed_digits(k, w, num_windows)
    L = 1 << w
    half = L >> 1
    out, carry, s = [], 0, -k
    for i in range(num_windows):
        v = ((s >> (i * w)) & (L - 1)) + carry
        if v > half:
            v -= L
            carry = 1
        else:
            carry = 0
        out.append(-v)
    assert carry == 0
    return out


# ----------------------------------------------------------------------------- arkworks compressed serialisation
# Restated from ark-serialize / ark-ec 0.4 (un-vendored): the format of the reference's instance files
# (src/msm/utils/preprocess.rs:181-225).  Unpinned by stored files: the reference ships none.
def ark_compress_g1(pt: Affine) -> bytes:
    """x canonical LE, 32 bytes; byte 31 bit 7 set iff y > p - y ("YIsNegative"); bit 6 = infinity (x = 0)."""
    if pt is None:
        b = bytearray(32)
        b[31] |= 1 << 6
        return bytes(b)
    b = bytearray(pt[0].to_bytes(32, "little"))
    if pt[1] > (P - pt[1]) % P:
        b[31] |= 1 << 7
    return bytes(b)


def ark_decompress_g1(b: bytes) -> Affine:
    flags = b[31] >> 6
    if flags & 1:
        return None
    x = int.from_bytes(b[:31] + bytes([b[31] & 0x3F]), "little")
    y = fq_sqrt((x * x * x + B_COEFF) % P)
    if y is None or x >= P:
        raise ValueError("not a curve point")
    if (y > (P - y) % P) != bool(flags & 2):
        y = (P - y) % P
    return (x, y)


def ark_serialize_instance(bases: Sequence[Affine], scalars: Sequence[int]) -> Tuple[bytes, bytes]:
    """(`points` record, `scalars` record): u64 LE length + 32-byte elements each."""
    pts = len(bases).to_bytes(8, "little") + b"".join(ark_compress_g1(p) for p in bases)
    sc = len(scalars).to_bytes(8, "little") + b"".join((s % R_ORDER).to_bytes(32, "little") for s in scalars)
    return pts, sc
