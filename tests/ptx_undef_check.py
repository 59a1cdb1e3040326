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


LABEL = re.compile(r"^(\$?[\w$]+):$")


def _function_bodies(ptx_text: str):
    """-> [(name, [instruction strings])]: one entry per .entry/.func, inline-asm wrappers flattened."""
    out = []
    fn, body = None, []
    for raw in ptx_text.splitlines():
        if FUNC.match(raw):
            if fn is not None:
                out.append((fn, body))
            m = re.search(r"(_Z\w+|\w+)\s*\(", raw) or re.search(r"(_Z\w+)", raw)
            fn, body = (m.group(1) if m else raw.strip()), []
            continue
        if fn is None:
            continue
        s = raw.split("//")[0].strip()
        if not s or s.startswith((".param", ".reg", ".local", ".shared", ".maxntid", ".minnctapersm", ".pragma")):
            continue
        # a line may hold several ';'-separated instructions, possibly wrapped in { .reg ...; ... }
        for piece in s.replace("{ .reg", ".reg").split(";"):
            piece = piece.strip().lstrip("{").rstrip("}").strip()
            if not piece or piece.startswith(".reg"):
                continue
            body.append(piece)
    if fn is not None:
        out.append((fn, body))
    return out


def check(ptx_text: str):
    """Returns a list of (function, register, line) for registers that are read although NO path from the function
    entry writes them first (forward may-be-defined data flow over the basic blocks; cicc lays loop bodies out in any
    textual order, so textual first-use is not a signal)."""
    bad = []
    for fn, body in _function_bodies(ptx_text):
        # basic blocks
        blocks, cur, labels = [], [], {}
        for ins in body:
            m = LABEL.match(ins)
            if m:
                if cur:
                    blocks.append(cur)
                    cur = []
                labels[m.group(1)] = len(blocks)
                continue
            cur.append(ins)
            op = ins.split(None, 1)[1] if ins.startswith("@") and " " in ins else ins
            if op.startswith(("bra", "ret", "exit", "brx", "trap")):
                blocks.append(cur)
                cur = []
        if cur:
            blocks.append(cur)
        nb = len(blocks)
        succ = [[] for _ in range(nb)]
        for b, ins_list in enumerate(blocks):
            last = ins_list[-1] if ins_list else ""
            guarded = last.startswith("@")
            op = last.split(None, 1)[1] if guarded and " " in last else last
            targets = [labels[t] for t in re.findall(r"\$?[\w$]*L__BB[\w$]+|\$L[\w$]+", op) if t in labels] if op.startswith(("bra", "brx")) else []
            succ[b].extend(t for t in targets if t < nb)
            falls = not op.startswith(("bra", "ret", "exit", "brx", "trap")) or guarded
            if falls and b + 1 < nb:
                succ[b].append(b + 1)
        defs = []
        for ins_list in blocks:
            d = set()
            for ins in ins_list:
                d.update(split_operands(ins)[0])
            defs.append(d)
        inset = [set() for _ in range(nb)]
        reached = [False] * nb
        if nb:
            reached[0] = True
        work = [0] if nb else []
        while work:
            b = work.pop()
            out = inset[b] | defs[b]
            for t in succ[b]:
                if not reached[t] or not out <= inset[t]:
                    inset[t] |= out
                    reached[t] = True
                    work.append(t)
        for b, ins_list in enumerate(blocks):
            if not reached[b]:
                continue
            have = set(inset[b])
            for ins in ins_list:
                dst, src = split_operands(ins)
                for r in src:
                    if r not in have:
                        bad.append((fn, r, ins))
                        have.add(r)
                have.update(dst)
    return bad


if __name__ == "__main__":
    res = check(open(sys.argv[1]).read())
    for fn, r, line in res[:50]:
        print(fn, r, "::", line)
    print(len(res), "suspicious first-uses")
