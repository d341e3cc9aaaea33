"""FP64 issue cost of a kernel's hot loop from its SASS (no GPU needed), with the operand rule measured on B200 by
profiles/tools/fp64_latency_map.cu: a warp-wide FP64 instruction occupies its scheduler's FP64 pipe for 2 cycles when it
reads at most two vector registers (DADD, DMUL, and DFMA with a constant-bank / uniform-register / immediate operand) and
for 3 cycles when it reads three distinct vector registers (the general DFMA).

    python profiles/tools/sass_fp64_cost.py lib.so|file.cubin kernel_name [lo hi]      (loop = largest backward branch span)"""
import re, subprocess, sys
path, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", "-fun", kern, path], capture_output=True, text=True).stdout
ins = []
for line in out.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
if len(sys.argv) >= 5:
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
else:
    best = None
    for a, s in ins:
        m = re.search(r"BRA\S*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", s)
        if m and int(m.group(1), 16) < a:
            span = a - int(m.group(1), 16)
            if best is None or span > best[0]:
                best = (span, int(m.group(1), 16), a)
    lo, hi = best[1], best[2]
n2 = n3 = other = 0
kinds = {}
for a, s in ins:
    if not (lo <= a <= hi):
        continue
    t = s.split()
    if t[0].startswith("@"):
        t = t[1:]
    op = t[0].split(".")[0]
    if op in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"):
        ops = " ".join(t[1:]).split(",")
        srcs = ops[1:] if op != "DSETP" else ops[2:]
        vec = set()
        for o in srcs:
            m = re.search(r"(?<![A-Z])R(\d+)", o.strip())
            if m and "UR" not in o:
                vec.add(m.group(1))
        if op == "DFMA" and len(vec) >= 3:
            n3 += 1
        else:
            n2 += 1
        kinds[op] = kinds.get(op, 0) + 1
    else:
        other += 1
cyc = 2 * n2 + 3 * n3
print(f"{kern}: loop 0x{lo:x}-0x{hi:x}: FP64 {n2 + n3} ({kinds}), of which {n3} DFMA with three vector-register operands; "
      f"other instructions {other}")
print(f"  FP64 pipe cycles per warp iteration: {cyc} (2 x {n2} + 3 x {n3}); at 2 cycles each it would be {2 * (n2 + n3)}: "
      f"pipe 'active' ceiling {2 * (n2 + n3) / cyc:.3f}")
