#!/usr/bin/env python3
"""Emit fq_asm.inc: fully unrolled PTX bodies for 256-bit Montgomery arithmetic on BN254 Fq.

Why generated: the whole product must be ONE asm block so that the PTX carry flag (CC) is
never live across compiler-scheduled code, and so that ptxas can pair every
`mad.lo.cc / madc.hi.cc` on the same operands into a single IMAD.WIDE.U32(.X) (64-bit
multiply-add with carry-in/carry-out in a predicate).  For that pairing the 64-bit addend
must sit in an aligned register pair, so partial products are split over two accumulators:
X holds pairs that start at EVEN absolute limb positions, Y pairs that start at ODD ones.
Row i of the CIOS loop therefore runs two independent carry chains (products whose position
i+j is even go to X, odd to Y) and one single-limb "fold" (S[i] += D[i]) whose carry-out
feeds the other accumulator's chain, which starts one limb higher.  Nothing is physically
shifted: positions are absolute and ptxas's allocator retires the low limbs.

The modulus limbs and -p^-1 mod 2^32 are emitted as immediates, so the reduction half of
the multiply needs no registers for p.

Replaces (does not translate) the reference's 16x16-bit-limb `mont_mul_cios`
(/root/reference/mopro-msm/src/msm/metal_msm/shader/mont_backend/mont.metal:105-181).
"""
import sys

P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
N0 = (-pow(P, -1, 1 << 32)) % (1 << 32)
PL = [(P >> (32 * i)) & 0xFFFFFFFF for i in range(8)]


class Emit:
    def __init__(self):
        self.lines = []

    def __call__(self, s):
        self.lines.append(s)


def mul_body(square_hint=False):
    """Operands: %0-%7 = r (out), %8-%15 = a, %16-%23 = b."""
    e = Emit()
    e(".reg .u32 x<17>, y<17>, m, t<8>;")
    e(".reg .pred pb;")
    for k in range(17):
        e(f"mov.u32 x{k}, 0;")
        e(f"mov.u32 y{k}, 0;")
    a = [f"%{8 + j}" for j in range(8)]
    b = [f"%{16 + j}" for j in range(8)]

    def chain(acc, pos0, mults, carry_in):
        """mults: list of (opA, opB) placed at positions pos0, pos0+2, ...; ends with addc into the next limb."""
        first = not carry_in
        pos = pos0
        for (u, v) in mults:
            lo = "mad.lo.cc.u32" if first else "madc.lo.cc.u32"
            e(f"{lo} {acc}{pos}, {u}, {v}, {acc}{pos};")
            e(f"madc.hi.cc.u32 {acc}{pos + 1}, {u}, {v}, {acc}{pos + 1};")
            first = False
            pos += 2
        e(f"addc.u32 {acc}{pos}, {acc}{pos}, 0;")

    for i in range(8):
        S, D = ("x", "y") if i % 2 == 0 else ("y", "x")
        # D-chain: products a[j]*b[i] with j odd -> positions i+1, i+3, i+5, i+7
        if i > 0:
            e(f"add.cc.u32 {S}{i}, {S}{i}, {D}{i};")
        chain(D, i + 1, [(a[j], b[i]) for j in (1, 3, 5, 7)], carry_in=(i > 0))
        # S-chain: j even -> positions i, i+2, i+4, i+6
        chain(S, i, [(a[j], b[i]) for j in (0, 2, 4, 6)], carry_in=False)
        # reduction row
        e(f"mul.lo.u32 m, {S}{i}, 0x{N0:08x};")
        chain(S, i, [("m", f"0x{PL[j]:08x}") for j in (0, 2, 4, 6)], carry_in=False)
        chain(D, i + 1, [("m", f"0x{PL[j]:08x}") for j in (1, 3, 5, 7)], carry_in=False)
    # merge: result limb k = x[8+k] + y[8+k]
    for k in range(8):
        op = "add.cc.u32" if k == 0 else "addc.cc.u32"
        e(f"{op} x{8 + k}, x{8 + k}, y{8 + k};")
    # conditional subtract p (inputs < p => result < 2p)
    for k in range(8):
        op = "sub.cc.u32" if k == 0 else "subc.cc.u32"
        e(f"{op} t{k}, x{8 + k}, 0x{PL[k]:08x};")
    e("subc.u32 m, 0, 0;")  # m = borrow ? 0xffffffff : 0
    e("setp.eq.u32 pb, m, 0;")  # no borrow -> take t
    for k in range(8):
        e(f"selp.u32 %{k}, t{k}, x{8 + k}, pb;")
    return e.lines


def sqr_body():
    """r = a*a*R^-1 mod p with 28 + 8 + 72 = 108 wide MACs (mul_body: 136).
    Operands: %0-%7 = r, %8-%15 = a.
      1. off-diagonal products a_i*a_j (i < j) into the even/odd-aligned accumulators X / Y;
      2. t = 2*(X + Y)            (merge, then a 1-bit funnel shift over 16 limbs);
      3. t += sum_i a_i^2 * 2^(64 i)   (all pairs even-aligned: ONE carry chain of 8 wide MACs);
      4. X = t, Y = 0 and the eight reduction rows of mul_body (m = S[i]*n0', += m*p); because X now holds
         live limbs above each chain's end, X-chains propagate their carry-out to limb 16."""
    e = Emit()
    e(".reg .u32 x<18>, y<18>, m, t<8>;")
    e(".reg .pred pb;")
    for k in range(18):
        e(f"mov.u32 x{k}, 0;")
        e(f"mov.u32 y{k}, 0;")
    a = [f"%{8 + j}" for j in range(8)]

    def chain(acc, pos0, mults, carry_in, propagate_to=None):
        first = not carry_in
        pos = pos0
        for (u, v) in mults:
            lo = "mad.lo.cc.u32" if first else "madc.lo.cc.u32"
            e(f"{lo} {acc}{pos}, {u}, {v}, {acc}{pos};")
            e(f"madc.hi.cc.u32 {acc}{pos + 1}, {u}, {v}, {acc}{pos + 1};")
            first = False
            pos += 2
        if propagate_to is None or pos >= propagate_to:
            e(f"addc.u32 {acc}{pos}, {acc}{pos}, 0;")
        else:
            while pos < propagate_to:
                e(f"addc.cc.u32 {acc}{pos}, {acc}{pos}, 0;")
                pos += 1
            e(f"addc.u32 {acc}{pos}, {acc}{pos}, 0;")

    # 1. off-diagonal triangle: row i multiplies a_i by a_j, j > i; position i+j even -> X, odd -> Y
    for i in range(7):
        ev = [j for j in range(i + 1, 8) if (i + j) % 2 == 0]
        od = [j for j in range(i + 1, 8) if (i + j) % 2 == 1]
        if ev:
            chain("x", i + ev[0], [(a[i], a[j]) for j in ev], carry_in=False)
        if od:
            chain("y", i + od[0], [(a[i], a[j]) for j in od], carry_in=False)
    # 2. merge (positions 1..15) and double
    for k in range(1, 16):
        op = "add.cc.u32" if k == 1 else "addc.cc.u32"
        e(f"{op} x{k}, x{k}, y{k};")
    e("addc.u32 x16, x16, y16;")
    for k in range(16, 0, -1):
        e(f"shf.l.wrap.b32 x{k}, x{k - 1}, x{k}, 1;")
    e("shl.b32 x0, x0, 1;")
    for k in range(18):
        e(f"mov.u32 y{k}, 0;")
    # 3. diagonal squares at even positions 0, 2, ..., 14
    chain("x", 0, [(a[i], a[i]) for i in range(8)], carry_in=False)
    # 4. reduction rows
    for i in range(8):
        S, D = ("x", "y") if i % 2 == 0 else ("y", "x")
        if i > 0:
            e(f"add.cc.u32 {S}{i}, {S}{i}, {D}{i};")
            # carry of the fold goes to position i+1 of D (as in mul_body, where it feeds D's a*b chain)
            e(f"addc.cc.u32 {D}{i + 1}, {D}{i + 1}, 0;")
            k = i + 2
            if D == "x":
                while k < 16:
                    e(f"addc.cc.u32 {D}{k}, {D}{k}, 0;")
                    k += 1
            e(f"addc.u32 {D}{k}, {D}{k}, 0;")
        e(f"mul.lo.u32 m, {S}{i}, 0x{N0:08x};")
        chain(S, i, [("m", f"0x{PL[j]:08x}") for j in (0, 2, 4, 6)], carry_in=False, propagate_to=16 if S == "x" else None)
        chain(D, i + 1, [("m", f"0x{PL[j]:08x}") for j in (1, 3, 5, 7)], carry_in=False, propagate_to=16 if D == "x" else None)
    for k in range(8):
        op = "add.cc.u32" if k == 0 else "addc.cc.u32"
        e(f"{op} x{8 + k}, x{8 + k}, y{8 + k};")
    for k in range(8):
        op = "sub.cc.u32" if k == 0 else "subc.cc.u32"
        e(f"{op} t{k}, x{8 + k}, 0x{PL[k]:08x};")
    e("subc.u32 m, 0, 0;")
    e("setp.eq.u32 pb, m, 0;")
    for k in range(8):
        e(f"selp.u32 %{k}, t{k}, x{8 + k}, pb;")
    return e.lines


def mulsub_body():
    """r = (a*b - c*d) * R^-1 mod p with ONE reduction: 64 + 64 + 72 = 200 wide MACs instead of 2 x 136 for two products
    and a subtraction (the Y3 = R (Q - X3) - Y1 PPP step of every XYZZ addition / doubling).
    Operands: %0-%7 = r, %8-%15 = a, %16-%23 = b, %24-%31 = c, %32-%39 = d.
    c is replaced by nc = p - c (in [1, p]), so every row accumulates a*b_i + nc*d_i >= 0; the result before the final
    conditional subtraction is (a b + nc d + M p) / R < (2 p^2 + R p) / R < 1.38 p, so one subtraction still suffices."""
    e = Emit()
    e(".reg .u32 x<17>, y<17>, m, t<8>, nc<8>;")
    e(".reg .pred pb;")
    for k in range(17):
        e(f"mov.u32 x{k}, 0;")
        e(f"mov.u32 y{k}, 0;")
    a = [f"%{8 + j}" for j in range(8)]
    b = [f"%{16 + j}" for j in range(8)]
    d = [f"%{32 + j}" for j in range(8)]
    for k in range(8):
        op = "sub.cc.u32" if k == 0 else ("subc.cc.u32" if k < 7 else "subc.u32")
        e(f"{op} nc{k}, 0x{PL[k]:08x}, %{24 + k};")
    nc = [f"nc{j}" for j in range(8)]

    def chain(acc, pos0, mults, carry_in):
        first = not carry_in
        pos = pos0
        for (u, v) in mults:
            lo = "mad.lo.cc.u32" if first else "madc.lo.cc.u32"
            e(f"{lo} {acc}{pos}, {u}, {v}, {acc}{pos};")
            e(f"madc.hi.cc.u32 {acc}{pos + 1}, {u}, {v}, {acc}{pos + 1};")
            first = False
            pos += 2
        e(f"addc.u32 {acc}{pos}, {acc}{pos}, 0;")

    for i in range(8):
        S, D = ("x", "y") if i % 2 == 0 else ("y", "x")
        if i > 0:
            e(f"add.cc.u32 {S}{i}, {S}{i}, {D}{i};")
        chain(D, i + 1, [(a[j], b[i]) for j in (1, 3, 5, 7)], carry_in=(i > 0))
        chain(D, i + 1, [(nc[j], d[i]) for j in (1, 3, 5, 7)], carry_in=False)
        chain(S, i, [(a[j], b[i]) for j in (0, 2, 4, 6)], carry_in=False)
        chain(S, i, [(nc[j], d[i]) for j in (0, 2, 4, 6)], carry_in=False)
        e(f"mul.lo.u32 m, {S}{i}, 0x{N0:08x};")
        chain(S, i, [("m", f"0x{PL[j]:08x}") for j in (0, 2, 4, 6)], carry_in=False)
        chain(D, i + 1, [("m", f"0x{PL[j]:08x}") for j in (1, 3, 5, 7)], carry_in=False)
    for k in range(8):
        op = "add.cc.u32" if k == 0 else "addc.cc.u32"
        e(f"{op} x{8 + k}, x{8 + k}, y{8 + k};")
    for k in range(8):
        op = "sub.cc.u32" if k == 0 else "subc.cc.u32"
        e(f"{op} t{k}, x{8 + k}, 0x{PL[k]:08x};")
    e("subc.u32 m, 0, 0;")
    e("setp.eq.u32 pb, m, 0;")
    for k in range(8):
        e(f"selp.u32 %{k}, t{k}, x{8 + k}, pb;")
    return e.lines


def muladd_body():
    """r = (a*b + c*d) * R^-1 mod p with ONE reduction: 64 + 64 + 72 = 200 wide MACs (the imaginary part a0 b1 + a1 b0 of an Fq2
    product; with mulsub_body for the real part an Fq2 product is 400 wide MACs and no Karatsuba additions).
    Operands: %0-%7 = r, %8-%15 = a, %16-%23 = b, %24-%31 = c, %32-%39 = d.  Before the final conditional subtraction the value
    is (a b + c d + M p) / R < (2 p^2 + R p) / R < 1.38 p, so one subtraction suffices."""
    e = Emit()
    e(".reg .u32 x<17>, y<17>, m, t<8>;")
    e(".reg .pred pb;")
    for k in range(17):
        e(f"mov.u32 x{k}, 0;")
        e(f"mov.u32 y{k}, 0;")
    a = [f"%{8 + j}" for j in range(8)]
    b = [f"%{16 + j}" for j in range(8)]
    c = [f"%{24 + j}" for j in range(8)]
    d = [f"%{32 + j}" for j in range(8)]

    def chain(acc, pos0, mults, carry_in):
        first = not carry_in
        pos = pos0
        for (u, v) in mults:
            lo = "mad.lo.cc.u32" if first else "madc.lo.cc.u32"
            e(f"{lo} {acc}{pos}, {u}, {v}, {acc}{pos};")
            e(f"madc.hi.cc.u32 {acc}{pos + 1}, {u}, {v}, {acc}{pos + 1};")
            first = False
            pos += 2
        e(f"addc.u32 {acc}{pos}, {acc}{pos}, 0;")

    for i in range(8):
        S, D = ("x", "y") if i % 2 == 0 else ("y", "x")
        if i > 0:
            e(f"add.cc.u32 {S}{i}, {S}{i}, {D}{i};")
        chain(D, i + 1, [(a[j], b[i]) for j in (1, 3, 5, 7)], carry_in=(i > 0))
        chain(D, i + 1, [(c[j], d[i]) for j in (1, 3, 5, 7)], carry_in=False)
        chain(S, i, [(a[j], b[i]) for j in (0, 2, 4, 6)], carry_in=False)
        chain(S, i, [(c[j], d[i]) for j in (0, 2, 4, 6)], carry_in=False)
        e(f"mul.lo.u32 m, {S}{i}, 0x{N0:08x};")
        chain(S, i, [("m", f"0x{PL[j]:08x}") for j in (0, 2, 4, 6)], carry_in=False)
        chain(D, i + 1, [("m", f"0x{PL[j]:08x}") for j in (1, 3, 5, 7)], carry_in=False)
    for k in range(8):
        op = "add.cc.u32" if k == 0 else "addc.cc.u32"
        e(f"{op} x{8 + k}, x{8 + k}, y{8 + k};")
    for k in range(8):
        op = "sub.cc.u32" if k == 0 else "subc.cc.u32"
        e(f"{op} t{k}, x{8 + k}, 0x{PL[k]:08x};")
    e("subc.u32 m, 0, 0;")
    e("setp.eq.u32 pb, m, 0;")
    for k in range(8):
        e(f"selp.u32 %{k}, t{k}, x{8 + k}, pb;")
    return e.lines


def add_body():
    """r = a + b mod p; %0-7 r, %8-15 a, %16-23 b."""
    e = Emit()
    e(".reg .u32 s<8>, t<8>, m;")
    e(".reg .pred pb;")
    for k in range(8):
        op = "add.cc.u32" if k == 0 else "addc.cc.u32"
        e(f"{op} s{k}, %{8 + k}, %{16 + k};")
    for k in range(8):
        op = "sub.cc.u32" if k == 0 else "subc.cc.u32"
        e(f"{op} t{k}, s{k}, 0x{PL[k]:08x};")
    e("subc.u32 m, 0, 0;")
    e("setp.eq.u32 pb, m, 0;")
    for k in range(8):
        e(f"selp.u32 %{k}, t{k}, s{k}, pb;")
    return e.lines


def sub_body():
    """r = a - b mod p."""
    e = Emit()
    e(".reg .u32 s<8>, m, q<8>;")
    for k in range(8):
        op = "sub.cc.u32" if k == 0 else "subc.cc.u32"
        e(f"{op} s{k}, %{8 + k}, %{16 + k};")
    e("subc.u32 m, 0, 0;")  # all-ones if borrow
    for k in range(8):
        e(f"and.b32 q{k}, m, 0x{PL[k]:08x};")
    for k in range(8):
        op = "add.cc.u32" if k == 0 else ("addc.cc.u32" if k < 7 else "addc.u32")
        e(f"{op} %{k}, s{k}, q{k};")
    return e.lines


def wrap(name, body, nin):
    outs = ", ".join(f'"=r"(r[{k}])' for k in range(8))
    names = "abcd"[:nin]
    ins = ", ".join(", ".join(f'"r"({v}[{k}])' for k in range(8)) for v in names)
    text = "\\n\\t".join(body)
    sig = ", ".join(f"const uint32_t (&{v})[8]" for v in names)
    return (f"__device__ __forceinline__ void {name}(uint32_t (&r)[8], {sig}) {{\n"
            f'    asm("{{\\n\\t{text}\\n\\t}}"\n        : {outs}\n        : {ins});\n}}\n')


def main():
    out = ["// GENERATED by gen_fq_asm.py -- do not edit.  BN254 Fq, 8x32-bit limbs, R = 2^256.\n",
           "#pragma once\n#include <cstdint>\n"]
    out.append(wrap("fq_mul_asm", mul_body(), 2))
    out.append(wrap("fq_sqr_asm", sqr_body(), 1))
    out.append(wrap("fq_mulsub_asm", mulsub_body(), 4))
    out.append(wrap("fq_muladd_asm", muladd_body(), 4))
    out.append(wrap("fq_add_asm", add_body(), 2))
    out.append(wrap("fq_sub_asm", sub_body(), 2))
    sys.stdout.write("\n".join(out))


if __name__ == "__main__":
    main()
