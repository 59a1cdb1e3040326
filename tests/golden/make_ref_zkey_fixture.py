#!/usr/bin/env python3
"""Derive tests/golden/ref_zkey_points.npz from the Groth16 proving key that ships INSIDE the reference tree:

    /root/reference/example-app/test-vectors/circom/multiplier2_final.zkey      (snarkjs zkey v1, 10 sections)

Section 2 (Groth16 header) holds alpha1, beta1 (G1), beta2, gamma2 (G2), delta1 (G1), delta2 (G2); sections 3, 5, 6, 8, 9
hold the IC, A, B1, C and H queries (G1); section 7 holds the B2 query (G2).  snarkjs stores coordinates as 32-byte
little-endian MONTGOMERY words (R = 2^256) -- the in-memory form of arkworks' `Fq` -- and a G2 point as
x.c0 | x.c1 | y.c0 | y.c1 (128 bytes), which is exactly the `G2Affine {x: Fq2 {c0, c1}, y: Fq2}` record the C ABI's G2 entry
points consume (x_off = 0, y_off = 64).  A point at infinity is stored as all-zero bytes.

These are bytes neither the oracle nor the kernels produced.  They pin the Fq2 layout / Montgomery assumptions of
oracle/bn254_g2.py and of the CUDA G2 path the way ref_srs_g1.npz pins G1: every record must decode to a point of the
curve (twist for G2) of order r under this repo's conventions.  Only the point records are extracted (19 G1 + 7 G2
points, 2 KiB); no source code is copied.
    python tests/golden/make_ref_zkey_fixture.py
"""
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import bn254 as o  # noqa: E402
import bn254_g2 as g2  # noqa: E402

SRC = "/root/reference/example-app/test-vectors/circom/multiplier2_final.zkey"


def sections(b):
    assert b[:4] == b"zkey"
    _, count = struct.unpack("<II", b[4:12])
    off, out = 12, {}
    for _ in range(count):
        t, sz = struct.unpack("<IQ", b[off:off + 12])
        out[t] = b[off + 12:off + 12 + sz]
        off += 12 + sz
    return out


def fq(b):
    return o.from_mont(int.from_bytes(b, "little"))


def main():
    s = sections(open(SRC, "rb").read())
    hdr = s[2]
    n8q = struct.unpack("<I", hdr[:4])[0]
    assert n8q == 32 and int.from_bytes(hdr[4:36], "little") == o.P
    assert int.from_bytes(hdr[40:72], "little") == o.R_ORDER
    pts = hdr[72 + 12:]
    assert len(pts) == 576
    g1_names, g1_blobs = ["alpha1", "beta1", "delta1"], [pts[0:64], pts[64:128], pts[384:448]]
    g2_names, g2_blobs = ["beta2", "gamma2", "delta2"], [pts[128:256], pts[256:384], pts[448:576]]
    for sec, nm in ((3, "ic"), (5, "a"), (6, "b1"), (8, "c"), (9, "h")):
        for k in range(len(s[sec]) // 64):
            g1_names.append(f"{nm}{k}")
            g1_blobs.append(s[sec][64 * k:64 * k + 64])
    for k in range(len(s[7]) // 128):
        g2_names.append(f"b2_{k}")
        g2_blobs.append(s[7][128 * k:128 * k + 128])
    n_inf = 0
    for nm, blob in zip(g1_names, g1_blobs):
        if blob == bytes(64):
            n_inf += 1
            continue
        pt = (fq(blob[:32]), fq(blob[32:]))
        assert o.is_on_curve(pt), nm
    for nm, blob in zip(g2_names, g2_blobs):
        if blob == bytes(128):
            n_inf += 1
            continue
        pt = ((fq(blob[0:32]), fq(blob[32:64])), (fq(blob[64:96]), fq(blob[96:128])))
        assert g2.is_on_curve(pt), nm
        assert g2.jac_is_inf(g2.jac_scalar_mul_raw(o.R_ORDER, g2.affine_to_jac(pt))), nm + ": not in the order-r subgroup"
    out = {"g1_names": np.array(g1_names), "g2_names": np.array(g2_names),
           "g1": np.frombuffer(b"".join(g1_blobs), dtype=np.uint64).reshape(-1, 8).copy(),
           "g2": np.frombuffer(b"".join(g2_blobs), dtype=np.uint64).reshape(-1, 16).copy()}
    np.savez_compressed(os.path.join(HERE, "ref_zkey_points.npz"), **out)
    print(len(g1_names), "G1 +", len(g2_names), "G2 points,", n_inf, "at infinity; all on curve, G2 in the r-torsion")


if __name__ == "__main__":
    main()
