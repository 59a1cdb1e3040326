#!/usr/bin/env python3
"""Derive tests/golden/ref_srs_g1.npz from data blobs that ship INSIDE the reference tree:

    /root/reference/example-app/ios/{plonk,gemini,hyperplonk}_fibonacci_srs.bin

They are structured-reference-string files of the demo app: a 4-byte header followed by consecutive BN254 G1
affine points stored as raw Montgomery-form words (x || y, 32 + 32 bytes, little-endian, R = 2^256) -- exactly
the in-memory representation of arkworks' `Fq`, i.e. the layout the C ABI of this repo consumes
(`x_off = 0, y_off = 32, stride = 64`).  The first point of every file is the generator (1, 2).
This is the only externally produced BN254 point data in the reference; it pins the layout/Montgomery
assumptions of the oracle and of the CUDA kernels against bytes that neither of them generated.

Only the 64-byte point records are extracted (80 points, 5 KiB); no source code is copied.
    python tests/golden/make_ref_srs_fixture.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import bn254 as o  # noqa: E402

SRC = "/root/reference/example-app/ios"
FILES = {"plonk": ("plonk_fibonacci_srs.bin", 16), "gemini": ("gemini_fibonacci_srs.bin", 32),
         "hyperplonk": ("hyperplonk_fibonacci_srs.bin", 32)}


def main():
    out = {}
    for key, (fn, count) in FILES.items():
        b = open(os.path.join(SRC, fn), "rb").read()
        pts = np.frombuffer(b[4:4 + 64 * count], dtype=np.uint64).reshape(count, 8).copy()
        for row in pts:
            x = o.from_mont(sum(int(row[j]) << (64 * j) for j in range(4)))
            y = o.from_mont(sum(int(row[4 + j]) << (64 * j) for j in range(4)))
            assert o.is_on_curve((x, y)), fn
        assert (o.from_mont(sum(int(pts[0][j]) << (64 * j) for j in range(4))),
                o.from_mont(sum(int(pts[0][4 + j]) << (64 * j) for j in range(4)))) == o.GEN
        out[key] = pts
        print(fn, count, "points, first = generator")
    np.savez_compressed(os.path.join(HERE, "ref_srs_g1.npz"), **out)


if __name__ == "__main__":
    main()
