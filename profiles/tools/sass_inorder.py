"""In-order issue model of a SASS loop body (no GPU needed): how many cycles one warp alone needs per iteration when
FP64 results are usable 8 cycles after issue and the FP64 pipe takes a warp instruction every 2 cycles (measured:
profiles/tools/fp64_latency_map.cu), against the 2 cycles per FP64 instruction the pipe itself needs.  The ratio is the
number of warps per scheduler the loop needs to keep the pipe busy.

    python profiles/tools/sass_inorder.py file.sass 0x2c00 0x4960"""
import re, sys
lines = open(sys.argv[1]).read().splitlines()
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
ins = []
for l in lines:
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and lo <= int(m.group(1), 16) <= hi:
        ins.append(m.group(2).strip())
ready = {}          # register -> cycle its value is available
cycle = 0
pipe_free = 0
nfp = 0
stall_hist = {}
for s in ins:
    t = s.split()
    if t[0].startswith("@"):
        t = t[1:]
    op = t[0].split(".")[0]
    regs = re.findall(r"\bR(\d+)\b", " ".join(t[1:]))
    if not regs:
        cycle += 1
        continue
    fp = op in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
    wide = fp or ".64" in t[0] or ".128" in t[0]
    dst = int(regs[0])
    srcs = [int(r) for r in regs[1:]]
    if op in ("STG", "STS", "ST", "BRA", "ISETP", "DSETP", "LDGSTS"):
        srcs = [int(r) for r in regs]
        dst = None
    need = cycle + 1
    for r in srcs:
        for rr in ((r, r + 1) if fp else (r,)):
            need = max(need, ready.get(rr, 0))
    if fp:
        need = max(need, pipe_free)
    stall = need - (cycle + 1)
    stall_hist[op] = stall_hist.get(op, 0) + stall
    cycle = need
    if fp:
        pipe_free = cycle + 2
        nfp += 1
    lat = 8 if fp else (24 if op in ("LDS", "LDC", "LDCU") else (400 if op in ("LDG",) else 5))
    if dst is not None:
        for rr in ((dst, dst + 1) if wide else (dst,)):
            ready[rr] = cycle + lat
print(f"{len(ins)} instructions, {nfp} FP64; one warp in order: {cycle} cycles; FP64 pipe alone: {2 * nfp} cycles; "
      f"warps per scheduler to fill the pipe: {cycle / (2.0 * nfp):.2f}")
print("stall cycles by waiting opcode:", dict(sorted(stall_hist.items(), key=lambda kv: -kv[1])[:6]))
