"""Per-kernel launch count, mean duration and share of the time of the repository's own kernels in an
`ncu --metrics gpu__time_duration.sum --csv` launch list.
    python profiles/tools/launch_shares.py launches.csv [substring ...]   (kernels whose name contains a substring count as "ours";
                                                                          default: ocs_ spg_ dense_project basis_eval soa cgb_)"""
import csv
import sys

ours = sys.argv[2:] or ["ocs_", "spg_", "dense_project", "basis_eval", "soa_kernel", "cgb_", "broadcast_rows", "reduce_kernel"]
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.reader(lines))
hdr = rows[0]
kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
mu = hdr.index("Metric Unit")
agg = {}
order = []
for r in rows[1:]:
    if r[mn] != "gpu__time_duration.sum":
        continue
    v = float(r[mv].replace(",", ""))
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[mu], 1.0)
    name = r[kn]
    if name not in agg:
        agg[name] = [0, 0.0]
        order.append(name)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(t for k, (c, t) in agg.items() if any(o in k for o in ours))
print("kernel, launches, mean ns, total ms, share of this repository's kernels")
for k in order:
    c, t = agg[k]
    mine = any(o in k for o in ours)
    print(f"{k[:90]}, {c}, {t / c:.0f}, {t / 1e6:.3f}" + (f", {t / tot:.3f}" if mine else ""))
