"""Static check of the generated PTX for registers that are USED BEFORE ANY DEFINITION inside a
function (first textual occurrence is a source operand).  cicc (CUDA 12.9) was observed to drop the
reload of word 0 of a by-reference accumulator after an out-of-line call, which shows up exactly
like this (`st.shared.v4.u32 [..], {%r380, ...}` with %r380 never written).  PTX emitted by nvcc
initialises loop-carried registers before the loop, so a textual first-use is a reliable signal."""
from __future__ import annotations

import re
import sys

REG = re.compile(r"%(?:rd|rs|r|p|fd|f|rh)\d+")
FUNC = re.compile(r"^\s*(?:\.visible\s+|\.weak\s+)?(?:\.entry|\.func)\b")


def split_operands(line: str):
    """-> (dest_regs, src_regs) for one PTX instruction line (best effort)."""
    line = line.split("//")[0].strip().rstrip(";")
    if not line or line.startswith((".", "{", "}", "$", "@")) and not line.startswith("@"):
        if not line.startswith("@"):
            return [], []
    pred_src = []
    if line.startswith("@"):
        g, _, line = line.partition(" ")
        pred_src = REG.findall(g)
        line = line.strip()
    parts = line.split(None, 1)
    if len(parts) < 2:
        return [], pred_src
    op, rest = parts
    if op.startswith(("st.", "bra", "call", "ret", "bar", "red.", "exit", "membar", "trap")) and not op.startswith(("bar.red", "barrier.red")):
        return [], pred_src + REG.findall(rest)
    # destination = first operand (possibly a {..} vector or "a|b" pair)
    depth, idx = 0, len(rest)
    for i, ch in enumerate(rest):
        if ch in "{[":
            depth += 1
        elif ch in "}]":
            depth -= 1
        elif ch == "," and depth == 0:
            idx = i
            break
    dst, src = rest[:idx], rest[idx + 1:]
    if "[" in dst:  # e.g. atom/ld forms never have memory dest here; treat as source
        return [], pred_src + REG.findall(rest)
    return REG.findall(dst), pred_src + REG.findall(src)


def check(ptx_text: str):
    """Returns a list of (function, register, line) for first-use-before-def registers."""
    bad = []
    fn, defined, in_asm = None, set(), False
    for raw in ptx_text.splitlines():
        if FUNC.match(raw):
            m = re.search(r"(_Z\w+|\w+)\s*\(", raw) or re.search(r"(_Z\w+)", raw)
            fn = m.group(1) if m else raw.strip()
            defined = set()
            continue
        if fn is None:
            continue
        s = raw.strip()
        if s.startswith(".param") or s.startswith(".reg") or s.startswith(".local") or s.startswith(".shared"):
            continue
        if "// begin inline asm" in s:
            in_asm = True
            continue
        if "// end inline asm" in s:
            in_asm = False
            continue
        # a line may hold several ';'-separated instructions, possibly wrapped in { .reg ...; ... }
        for piece in s.replace("{ .reg", ".reg").split(";"):
            piece = piece.strip().lstrip("{").rstrip("}").strip()
            if not piece or piece.startswith(".reg"):
                continue
            dst, src = split_operands(piece)
            for r in src:
                if r not in defined:
                    bad.append((fn, r, s))
                    defined.add(r)
            defined.update(dst)
    return bad


if __name__ == "__main__":
    res = check(open(sys.argv[1]).read())
    for fn, r, line in res[:50]:
        print(fn, r, "::", line)
    print(len(res), "suspicious first-uses")
