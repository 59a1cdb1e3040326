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


def test_checker_flags_a_never_written_register_and_follows_loops():
    """The checker itself: a register no path writes is reported (the miscompile's signature); a loop-carried register
    that is only written textually AFTER its use (cicc's block layout for rotated loops) is not."""
    ptx = """
.visible .entry k_demo(
    .param .u64 p0
)
{
    .reg .b32 %r<9>;
    .reg .pred %p<3>;
    mov.u32 %r1, 0;
    bra.uni $L__BB0_3;
$L__BB0_2:
    add.u32 %r3, %r2, %r1;
    setp.lt.u32 %p1, %r3, 10;
    @%p1 bra $L__BB0_3;
    bra.uni $L__BB0_4;
$L__BB0_3:
    add.u32 %r2, %r1, 1;
    bra.uni $L__BB0_2;
$L__BB0_4:
    st.shared.v4.u32 [%r3], {%r5, %r1, %r2, %r3};
    ret;
}
"""
    bad = ptx_undef_check.check(ptx)
    assert [(fn, r) for fn, r, _ in bad] == [("k_demo", "%r5")]
