"""Compile one sweep-kernel unit of a generated problem for sm_100a with nvcc (what NVRTC does on the GPU box, same
defines) and report registers, spills and the FP64 instruction mix of its SASS.
usage: python profiles/tools/gen_sass.py "OCS_REV_KERNEL(ocs_gen::Problem, gen, 2, 0, 0, 0)" [extra -D...]"""
import importlib, os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
ocs = importlib.import_module("optimal-control-solvers_b200")
unit = sys.argv[1]
extra = sys.argv[2:]
prob = getattr(ocs, os.environ.get("PROB", "harvest_chain_problem"))()
pre = "#define OCS_NS ocs_gen\n"
if prob.nStates >= 4:       # csrc/nvrtc_problem.cu, ocs_compile_unit
    pre += ("#define OCS_THREADS 64\n#define OCS_TILE 16\n#define OCS_REV_UNROLL 1\n#define OCS_RING 4\n#define OCS_FB_RING 4\n"
            "#define OCS_MINBLOCKS_REV 6\n#define OCS_MINBLOCKS_FWD 6\n#define OCS_MINBLOCKS_FB 4\n"
            "#define OCS_LB_FWD_DENSE 6\n#define OCS_LB_REV_GRID 4\n#define OCS_LB_FWD_GRID 8\n")
for d in extra:
    if d.startswith("-D"):
        k, _, v = d[2:].partition("=")
        pre += f"#undef {k}\n#define {k} {v}\n"
out = os.environ.get("OUT", "/tmp/sass/gen")
open(out + ".cu", "w").write(pre + '#include "rk4_kernels.cuh"\n' + prob.source + "\n" + unit + "\n")
r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-lineinfo", "-Xptxas", "-v",
                    "-I", os.path.join(ROOT, "optimal-control-solvers_b200", "csrc"), "-cubin", "-o", out + ".cubin", out + ".cu"],
                   capture_output=True, text=True)
print("\n".join(l for l in r.stderr.splitlines() if "registers" in l or "spill" in l or "error" in l))
if r.returncode:
    print(r.stderr); sys.exit(1)
sass = subprocess.run(["cuobjdump", "-sass", out + ".cubin"], capture_output=True, text=True).stdout
open(out + ".sass", "w").write(sass)
cnt = collections.Counter()
for l in sass.splitlines():
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        cnt[m.group(1).split(".")[0]] += 1
print({k: v for k, v in cnt.most_common(25)})
