#!/usr/bin/env python
"""Instruction mix of the hot loop(s) of a kernel, from `cuobjdump -sass` (no GPU needed).

    python profiles/tools/sass_loop_mix.py file.cubin|lib.so kernel_name [--all]

Finds every backward branch of the kernel, takes the loop bodies (innermost first, the largest by default),
and counts SASS opcodes by class: FP64 (DFMA/DMUL/DADD/DSETP/MUFU.RCP64H...), shared-memory loads, global
loads/stores, local (spill) traffic, integer/address, moves, branches/barriers."""
import re
import subprocess
import sys
from collections import Counter


def sass(path, kernel):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", kernel, path], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def klass(op):
    op = op.split()[0] if not op.startswith("@") else op.split()[1]
    base = op.split(".")[0]
    if base in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"):
        return "fp64"
    if base == "MUFU":
        return "mufu"
    if base in ("LDS", "LDSM"):
        return "lds"
    if base in ("STS",):
        return "sts"
    if base in ("LDG", "LD", "LDGSTS"):
        return "ldg"
    if base in ("STG", "ST"):
        return "stg"
    if base in ("LDL", "STL"):
        return "local"
    if base in ("LDC", "ULDC", "LDCU"):
        return "ldc"
    if base in ("MOV", "UMOV", "IMAD.MOV", "PRMT", "SEL", "FSEL", "R2UR", "S2R", "CS2R", "S2UR"):
        return "mov"
    if base in ("SHFL",):
        return "shfl"
    if base in ("BRA", "BSSY", "BSYNC", "BAR", "EXIT", "WARPSYNC", "NOP", "BREAK", "CALL", "RET", "DEPBAR", "ERRBAR", "MEMBAR"):
        return "ctrl"
    if base in ("IMAD", "IADD3", "IADD", "LEA", "ISETP", "LOP3", "SHF", "UIADD3", "ULEA", "UIMAD", "ULOP3", "USHF", "UISETP",
                "IMNMX", "VIADD", "PLOP3", "UPLOP3", "IABS", "VIMNMX", "UVIADD", "UIADD"):
        if op.startswith("IMAD.MOV"):
            return "mov"
        return "int"
    return "other:" + base


def main():
    path, kernel = sys.argv[1], sys.argv[2]
    ins = sass(path, kernel)
    if not ins:
        sys.exit("no SASS for " + kernel)
    addr_index = {a: k for k, (a, _) in enumerate(ins)}
    loops = []
    for k, (a, text) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.[A-Z.]+)?\s+(?:[!A-Z0-9]+,\s*)?`?\(?0x([0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((addr_index[tgt], k))
    print(f"{kernel}: {len(ins)} SASS instructions, {len(loops)} backward branches")
    loops.sort(key=lambda lk: lk[1] - lk[0], reverse=True)
    sel = loops if "--all" in sys.argv else loops[:3]
    for lo, hi in sel:
        body = ins[lo:hi + 1]
        c = Counter(klass(t) for _, t in body)
        tot = sum(c.values())
        fp = c.get("fp64", 0)
        print(f"  loop 0x{ins[lo][0]:x}..0x{ins[hi][0]:x}: {tot} instr; " +
              ", ".join(f"{k}={v}" for k, v in sorted(c.items(), key=lambda kv: -kv[1])) +
              f"; model FP64-pipe share 2*fp64/(2*fp64+other) = {2 * fp / max(1, 2 * fp + tot - fp):.2f}")


if __name__ == "__main__":
    main()
