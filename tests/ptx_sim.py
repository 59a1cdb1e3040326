"""A tiny interpreter for the PTX subset emitted by gpu-acceleration_b200/csrc/gen_fq_asm.py.

Lets the `-m "not gpu"` suite execute the exact instruction stream the CUDA kernels run
(carry-flag semantics included) against the big-int oracle, without a GPU.
"""
from __future__ import annotations

import re
from typing import Dict, List

M32 = 0xFFFFFFFF


def run(body: List[str], operands: Dict[str, int]) -> Dict[str, int]:
    """body: PTX lines; operands: {'%8': value, ...} for inputs. Returns all registers."""
    reg: Dict[str, int] = dict(operands)
    pred: Dict[str, bool] = {}
    cf = 0

    def val(tok: str) -> int:
        tok = tok.strip()
        if tok.startswith("0x"):
            return int(tok, 16)
        if re.fullmatch(r"\d+", tok):
            return int(tok)
        return reg[tok]

    for line in body:
        line = line.strip().rstrip(";")
        if not line or line.startswith("."):
            continue
        op, rest = line.split(None, 1)
        args = [a.strip() for a in rest.split(",")]
        d = args[0]
        if op == "mov.u32":
            reg[d] = val(args[1])
        elif op == "mul.lo.u32":
            reg[d] = (val(args[1]) * val(args[2])) & M32
        elif op in ("mad.lo.cc.u32", "madc.lo.cc.u32", "mad.lo.u32"):
            s = ((val(args[1]) * val(args[2])) & M32) + val(args[3]) + (cf if op.startswith("madc") else 0)
            reg[d] = s & M32
            if ".cc" in op:
                cf = s >> 32
        elif op in ("mad.hi.cc.u32", "madc.hi.cc.u32", "madc.hi.u32", "mad.hi.u32"):
            s = ((val(args[1]) * val(args[2])) >> 32) + val(args[3]) + (cf if op.startswith("madc") else 0)
            reg[d] = s & M32
            if ".cc" in op:
                cf = s >> 32
        elif op in ("add.cc.u32", "addc.cc.u32", "addc.u32", "add.u32"):
            s = val(args[1]) + val(args[2]) + (cf if op.startswith("addc") else 0)
            reg[d] = s & M32
            if ".cc" in op:
                cf = s >> 32
        elif op in ("sub.cc.u32", "subc.cc.u32", "subc.u32", "sub.u32"):
            s = val(args[1]) - val(args[2]) - (cf if op.startswith("subc") else 0)
            reg[d] = s & M32
            if ".cc" in op:
                cf = 1 if s < 0 else 0
        elif op == "shf.l.wrap.b32":
            lo, hi, sh = val(args[1]), val(args[2]), val(args[3]) & 31
            reg[d] = (((hi << 32) | lo) << sh >> 32) & M32
        elif op == "shl.b32":
            reg[d] = (val(args[1]) << val(args[2])) & M32
        elif op == "and.b32":
            reg[d] = val(args[1]) & val(args[2])
        elif op == "setp.eq.u32":
            pred[d] = val(args[1]) == val(args[2])
        elif op == "selp.u32":
            reg[d] = val(args[1]) if pred[args[3]] else val(args[2])
        else:
            raise ValueError(f"unsupported PTX op: {line}")
    return reg


def call(body: List[str], a: int, b: int | None = None, c: int | None = None, d: int | None = None) -> int:
    ops = {f"%{8 + k}": (a >> (32 * k)) & M32 for k in range(8)}
    for base, v in ((16, b), (24, c), (32, d)):
        if v is not None:
            ops.update({f"%{base + k}": (v >> (32 * k)) & M32 for k in range(8)})
    out = run(body, ops)
    return sum(out[f"%{k}"] << (32 * k) for k in range(8))
