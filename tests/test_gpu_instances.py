"""GPU parity for the reference's benchmark-instance path: `points`/`scalars` files in
arkworks' compressed serialisation (src/msm/utils/preprocess.rs:181-225) decoded on the GPU and fed to the MSM,
i.e. `benchmark_msm` (src/msm/arkworks_pippenger.rs:7-43) with the CUDA engine in place of `G::msm`.
The format restatement is unpinned by stored files (the reference ships none): round trip vs the oracle."""
import numpy as np
import pytest

import b200msm
import bn254 as o
import helpers as h

pytestmark = pytest.mark.gpu


def test_decompress_and_scalar_conversion(ctx):
    pts = o.random_points(300, 444) + [None, o.GEN, o.affine_neg(o.GEN)]
    comp = np.frombuffer(b"".join(o.ark_compress_g1(p) for p in pts), dtype=np.uint8).reshape(-1, 32)
    assert [o.ark_decompress_g1(bytes(c)) for c in comp] == pts  # oracle round trip
    out, bad = ctx.decompress_g1(comp)
    assert bad == 0
    want = h.pack_bases(pts)[:, :8]  # infinity -> (0, 0)
    assert np.array_equal(out, want)
    # an x with no square root, and an x >= p, are reported, not silently accepted
    x = 5
    while o.fq_sqrt((x ** 3 + 3) % o.P) is not None:
        x += 1
    bogus = np.frombuffer(x.to_bytes(32, "little") + (o.P + 1).to_bytes(32, "little"), dtype=np.uint8).reshape(2, 32)
    _, bad = ctx.decompress_g1(bogus)
    assert bad == 2
    sc = [0, 1, o.R_ORDER - 1] + o.random_scalars(200, 445)
    canon = np.array([h.words(s) for s in sc], dtype=np.uint64)
    assert np.array_equal(ctx.fr_to_montgomery(canon), h.pack_scalars(sc))


def test_instance_files_end_to_end(ctx, tmp_path):
    # two instances appended to the same pair of files, as gen_vectors writes them (preprocess.rs:181-225)
    want = []
    with open(tmp_path / "points", "wb") as fp, open(tmp_path / "scalars", "wb") as fs:
        for k, n in enumerate((257, 1024)):
            pts = o.random_points(n, 600 + k)
            pts[3] = None
            sc = o.random_scalars(n, 700 + k)
            a, b = o.ark_serialize_instance(pts, sc)
            fp.write(a)
            fs.write(b)
            want.append(o.jac_to_affine(o.msm_pippenger(pts, sc, 8)))
    got = [h.result_affine(b200msm.msm_from_instance(ctx, p, s)) for p, s in b200msm.read_instance_files(str(tmp_path))]
    assert got == want


def test_instance_writer_round_trip_and_replay(ctx, tmp_path):
    """write_instance_files (the gen_vectors side, preprocess.rs:181-225) produces what the oracle's restatement of the
    format produces, byte for byte; bench.py --replay (run_benchmark, arkworks_pippenger.rs:45-75) runs over the files."""
    import json
    import os
    import subprocess
    import sys
    insts, blobs = [], [b"", b""]
    for k, n in enumerate((33, 300)):
        pts = o.random_points(n, 800 + k)
        pts[2] = None
        sc = o.random_scalars(n, 900 + k)
        insts.append((h.pack_bases(pts), h.pack_scalars(sc)))
        a, b = o.ark_serialize_instance(pts, sc)
        blobs[0] += a
        blobs[1] += b
    b200msm.write_instance_files(str(tmp_path), insts)
    assert open(tmp_path / "points", "rb").read() == blobs[0]
    assert open(tmp_path / "scalars", "rb").read() == blobs[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--replay", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["num_instance"] == 2 and line["instance_size"] == 33 and line["verified_equal"]
