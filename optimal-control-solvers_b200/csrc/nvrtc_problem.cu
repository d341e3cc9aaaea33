// nvrtc_problem.cu -- run-time compilation of a generated OCProblem.
//
// The make_from_symbolic CUDA backend (optimal-control-solvers_b200/codegen) prints the problem's
// F / dFdx_times_vec / dFdu_times_vec / ControlChar as a `struct Problem` of device
// functions (replacing matlabFunction, functions/make_from_symbolic.m:27-31).  Here
// that source is compiled with NVRTC for sm_100a TOGETHER with rk4_kernels.cuh, so the
// right-hand sides inline into the sweep kernels (no function-pointer calls), and the
// resulting cubin is loaded as a cudaLibrary_t whose kernels launch through the same
// cudaLaunchKernel path as the built-in ones.
//
// libnvrtc is opened lazily with dlopen: the library itself loads (and exports its
// symbols) on machines without the CUDA toolkit's NVRTC or without a driver.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ocs_b200.h"
#include "kernel_table.h"

namespace {

typedef struct _nvrtcProgram *nvrtcProgram;
typedef int nvrtcResult;

struct Nvrtc {
    void *lib = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *);
    nvrtcResult (*DestroyProgram)(nvrtcProgram *);
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *);
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *);
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char *);
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *);
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char *);
    const char *(*GetErrorString)(nvrtcResult);
};

bool load_nvrtc(Nvrtc &n, std::string *log)
{
    static Nvrtc cached;
    static std::mutex mu;                       // a multi-device handle compiles on one thread per device
    std::lock_guard<std::mutex> lock(mu);
    if (cached.lib) {
        n = cached;
        return true;
    }
    const char *names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", nullptr};
    void *lib = nullptr;
    for (int k = 0; names[k] && !lib; ++k) lib = dlopen(names[k], RTLD_NOW | RTLD_LOCAL);
    if (!lib) {
        *log = std::string("cannot open libnvrtc: ") + dlerror();
        return false;
    }
    n.lib = lib;
#define SYM(field, name)                                            \
    *(void **)(&n.field) = dlsym(lib, name);                        \
    if (!n.field) {                                                 \
        *log = std::string("libnvrtc lacks ") + name;               \
        return false;                                               \
    }
    SYM(CreateProgram, "nvrtcCreateProgram")
    SYM(DestroyProgram, "nvrtcDestroyProgram")
    SYM(CompileProgram, "nvrtcCompileProgram")
    SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    SYM(GetCUBIN, "nvrtcGetCUBIN")
    SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    SYM(GetProgramLog, "nvrtcGetProgramLog")
    SYM(GetErrorString, "nvrtcGetErrorString")
#undef SYM
    cached = n;
    return true;
}

std::string default_include_dir()
{
    Dl_info info;
    if (dladdr((const void *)&default_include_dir, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        const size_t k = p.find_last_of('/');
        p = (k == std::string::npos) ? std::string(".") : p.substr(0, k);
        return p + "/csrc";
    }
    return "csrc";
}

}  // namespace

// Compiles `source` (which defines `namespace ocs_gen { struct Problem {...}; }`) followed by
// `unit` -- one or more OCS_*_KERNEL macro invocations -- and returns the named kernels of the
// resulting library.  Sweep variants are compiled one at a time, on first use: a 6-state problem
// takes a second or two per variant instead of a minute for all of them.
int ocs_compile_unit(const char *source, const char *unit, int strict, const char *include_dir,
                     const char *const *names, int nnames, const void **kernels, void **library, std::string *log,
                     int n_states)
{
    Nvrtc rt;
    if (!load_nvrtc(rt, log)) return OCS_B200_ERR_COMPILE;

    std::string src;
    src += "#define OCS_NS ocs_gen\n";
    // the strict build keeps the reference's operation order everywhere it matters: the sweep bodies (rk4_kernels.cuh)
    // and the generated right-hand sides (codegen/fma_forms.py prints re-associated forms for the other build only)
    // `strict` carries two flags: bit 0 = the strict arithmetic mode, bit 1 = the grid's steps agree to rounding, so the
    // FMA-contracted build may take the step sizes as constant-bank operands (OCS_UNIFORM_H, rk4_kernels.cuh)
    const bool uniform_h = (strict & 2) != 0;
    strict &= 1;
    src += strict ? "#define OCS_STRICT 1\n" : "#define OCS_STRICT 0\n";
    src += (uniform_h && !strict) ? "#define OCS_UNIFORM_H 1\n" : "#define OCS_UNIFORM_H 0\n";
    // Sweep-kernel tuning by problem size (B200, profiles/r2_notes.md).  A problem with several states carries
    // 150+ live registers per trajectory, so only 8-12 warps are resident per SM: nothing hides a global-load
    // latency.  Its per-step loads therefore go through a cp.async ring 4 steps ahead, the reverse step loop is not
    // unrolled (fewer live values) and the kernels are capped at 168 registers (12 resident warps) except the two
    // gradient/costate sweeps over a per-trajectory control grid, which spill below ~220 (8 warps).  Blocks are 64
    // threads with 16-step record tiles: with so few rounds of blocks per kernel (10^5 trajectories are 2-3) the last
    // round is filled in finer units (alone-kernel times -3 %; same-box A/B r2j/r2p).
    if (n_states >= 4) {
        src += "#define OCS_THREADS 64\n#define OCS_TILE 16\n#define OCS_REV_UNROLL 1\n#define OCS_RING 4\n#define OCS_FB_RING 4\n"
               "#define OCS_MINBLOCKS_REV 6\n#define OCS_MINBLOCKS_FWD 6\n#define OCS_MINBLOCKS_FB 4\n"
               "#define OCS_LB_FWD_DENSE 6\n#define OCS_LB_REV_GRID 4\n#define OCS_LB_FWD_GRID 8\n";
    }
    // kernel tuning experiments: OCS_B200_NVRTC_DEFS="OCS_MINBLOCKS_REV=3;OCS_TPT_FWD=2" -> #define lines
    if (const char *defs = getenv("OCS_B200_NVRTC_DEFS")) {
        std::string d(defs);
        size_t pos = 0;
        while (pos < d.size()) {
            size_t end = d.find(';', pos);
            if (end == std::string::npos) end = d.size();
            std::string item = d.substr(pos, end - pos);
            const size_t eq = item.find('=');
            if (!item.empty()) {
                src += "#undef " + (eq == std::string::npos ? item : item.substr(0, eq)) + "\n";
                src += "#define " + (eq == std::string::npos ? item : item.substr(0, eq) + " " + item.substr(eq + 1)) + "\n";
            }
            pos = end + 1;
        }
    }
    src += "#include \"rk4_kernels.cuh\"\n";
    src += source;
    src += "\n";
    src += unit;
    src += "\n";

    nvrtcProgram prog = nullptr;
    nvrtcResult r = rt.CreateProgram(&prog, src.c_str(), "ocs_generated_problem.cu", 0, nullptr, nullptr);
    if (r != 0) {
        *log = std::string("nvrtcCreateProgram: ") + rt.GetErrorString(r);
        return OCS_B200_ERR_COMPILE;
    }
    const std::string inc = std::string("-I") + (include_dir ? std::string(include_dir) : default_include_dir());
    // -default-device: generic lambdas inside the sweep bodies carry no execution-space annotation
    std::vector<const char *> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device", inc.c_str(),
                                      strict ? "--fmad=false" : "--fmad=true"};
    r = rt.CompileProgram(prog, (int)opts.size(), opts.data());
    size_t logsz = 0;
    rt.GetProgramLogSize(prog, &logsz);
    if (logsz > 1) {
        std::vector<char> buf(logsz + 1, 0);
        rt.GetProgramLog(prog, buf.data());
        *log = buf.data();
    }
    if (r != 0) {
        *log += std::string("\nnvrtcCompileProgram: ") + rt.GetErrorString(r);
        rt.DestroyProgram(&prog);
        return OCS_B200_ERR_COMPILE;
    }
    size_t sz = 0;
    rt.GetCUBINSize(prog, &sz);
    std::vector<char> cubin(sz);
    r = rt.GetCUBIN(prog, cubin.data());
    rt.DestroyProgram(&prog);
    if (r != 0 || sz == 0) {
        *log += "\nnvrtcGetCUBIN returned nothing";
        return OCS_B200_ERR_COMPILE;
    }
    cudaLibrary_t lib = nullptr;
    cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) {
        *log += std::string("\ncudaLibraryLoadData: ") + cudaGetErrorString(e);
        return OCS_B200_ERR_CUDA;
    }
    for (int k = 0; k < nnames; ++k) {
        cudaKernel_t kern = nullptr;
        if (cudaLibraryGetKernel(&kern, lib, names[k]) != cudaSuccess) {
            cudaGetLastError();
            *log += std::string("\ncompiled library lacks kernel ") + names[k];
            cudaLibraryUnload(lib);
            return OCS_B200_ERR_COMPILE;
        }
        kernels[k] = (const void *)kern;
    }
    *library = (void *)lib;
    return 0;
}

// Base unit of a generated problem: the table-fill and info kernels; fills the problem constants
// of `table` (every sweep entry stays null until ocs_compile_unit resolves it).
int ocs_compile_problem(const char *source, int n, int nc, int np, int strict, const char *include_dir,
                        ocs_kernel_table *table, void **library, std::string *log)
{
    std::string unit;
    unit += "static_assert(ocs_gen::Problem::n == " + std::to_string(n) + ", \"state count mismatch\");\n";
    unit += "static_assert(ocs_gen::Problem::nc == " + std::to_string(nc) + ", \"control count mismatch\");\n";
    unit += "static_assert(ocs_gen::Problem::np == " + std::to_string(np) + ", \"parameter count mismatch\");\n";
    unit += "OCS_TT_KERNEL(ocs_gen::Problem, gen)\n";
    const char *names[2] = {"ocs_tt_gen", "ocs_info_gen"};
    const void *kern[2] = {nullptr, nullptr};
    int rc = ocs_compile_unit(source, unit.c_str(), strict, include_dir, names, 2, kern, library, log, n);
    if (rc != 0) return rc;
    memset(table, 0, sizeof *table);
    table->tt = kern[0];
    int *d_info = nullptr, inf[8] = {0};
    cudaError_t e = cudaMalloc(&d_info, sizeof inf);
    if (e == cudaSuccess) {
        void *args[] = {(void *)&d_info};
        e = cudaLaunchKernel(kern[1], dim3(1), dim3(1), args, 0, nullptr);
        if (e == cudaSuccess) e = cudaMemcpy(inf, d_info, sizeof inf, cudaMemcpyDeviceToHost);
        cudaFree(d_info);
    }
    if (e != cudaSuccess) {
        *log += std::string("\nocs_info_gen launch: ") + cudaGetErrorString(e);
        cudaLibraryUnload((cudaLibrary_t)*library);
        *library = nullptr;
        return OCS_B200_ERR_CUDA;
    }
    table->n = inf[0];
    table->nc = inf[1];
    table->np = inf[2];
    for (int hm = 0; hm < 3; ++hm) table->ntt[hm] = inf[3 + hm] & 0xffff;
    table->tpt_fwd_coef = (inf[5] >> 16) & 0xff;
    table->rev_ring_pb = (inf[5] >> 24) & 0xff;
    table->tpt_fwd_pb = (inf[4] >> 16) & 0xf;
    table->tpt_rev_pb = (inf[4] >> 20) & 0xf;
    table->tile = inf[6] & 0xffff;
    table->rev_ring = (inf[6] >> 16) & 0xff;
    table->fb_ring = (inf[6] >> 24) & 0xf;
    table->async_stage = (inf[6] >> 28) & 1;
    table->threads = inf[7] & 0xffff;
    table->tpt_rev = (inf[7] >> 16) & 0xff;
    table->tpt_fwd = (inf[7] >> 24) & 0xf;
    table->tpt_rev_coef = (inf[7] >> 28) & 0xf;
    table->kd = 0;
    return 0;
}

void ocs_unload_problem(void *library)
{
    if (library) cudaLibraryUnload((cudaLibrary_t)library);
}
