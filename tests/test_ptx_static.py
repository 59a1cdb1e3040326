"""CPU suite: compile the library to PTX (no GPU needed) and check every kernel for registers that
are read before any definition -- the signature of the cicc miscompile documented in
gpu-acceleration_b200/csrc/msm_kernels.cuh (block_weighted_sum takes its operands BY VALUE because of it)."""
import os
import subprocess

import ptx_undef_check

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_use_before_def_in_ptx(tmp_path):
    ptx = tmp_path / "b200msm.ptx"
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-ptx", "-o", str(ptx),
                    os.path.join(ROOT, "gpu-acceleration_b200", "csrc", "b200msm.cu")], check=True, capture_output=True)
    bad = ptx_undef_check.check(ptx.read_text())
    assert not bad, bad[:5]
