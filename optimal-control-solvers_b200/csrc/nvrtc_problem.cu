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
#include <cstring>
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

// Compiles `source` (which defines `namespace ocs_gen { struct Problem {...}; }` with static
// constexpr n, nc, np, ntt) and fills `table` with the kernels of the resulting library.
int ocs_compile_problem(const char *source, int n, int nc, int np, int kd, int strict, const char *include_dir,
                        ocs_kernel_table *table, void **library, std::string *log)
{
    Nvrtc rt;
    if (!load_nvrtc(rt, log)) return OCS_B200_ERR_COMPILE;

    std::string src;
    src += "#define OCS_NS ocs_gen\n";
    src += "#include \"rk4_kernels.cuh\"\n";
    src += source;
    src += "\nstatic_assert(ocs_gen::Problem::n == " + std::to_string(n) + ", \"state count mismatch\");\n";
    src += "static_assert(ocs_gen::Problem::nc == " + std::to_string(nc) + ", \"control count mismatch\");\n";
    src += "static_assert(ocs_gen::Problem::np == " + std::to_string(np) + ", \"parameter count mismatch\");\n";
    src += "OCS_DEFINE_KERNELS(ocs_gen::Problem, gen, " + std::to_string(kd) + ")\n";

    nvrtcProgram prog = nullptr;
    nvrtcResult r = rt.CreateProgram(&prog, src.c_str(), "ocs_generated_problem.cu", 0, nullptr, nullptr);
    if (r != 0) {
        *log = std::string("nvrtcCreateProgram: ") + rt.GetErrorString(r);
        return OCS_B200_ERR_COMPILE;
    }
    const std::string inc = std::string("-I") + (include_dir ? std::string(include_dir) : default_include_dir());
    std::vector<const char *> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", inc.c_str(),
                                      strict ? "--fmad=false" : "--fmad=true", "-default-device"};
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
    auto get = [&](const std::string &name) -> const void * {
        cudaKernel_t k = nullptr;
        if (cudaLibraryGetKernel(&k, lib, name.c_str()) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (const void *)k;
    };
    memset(table, 0, sizeof *table);
    table->n = n;
    table->nc = nc;
    table->np = np;
    table->kd = kd;
    table->threads = 128;
    const int combos_f[4][2] = {{0, 0}, {1, 1}, {1, 2}, {2, 0}};
    const int combos_r[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 0}};
    bool ok = true;
    for (int pb = 0; pb < 2; ++pb) {
        char name[96];
        for (auto &c : combos_f) {
            snprintf(name, sizeof name, "ocs_fwd_gen_u%d_b%d_p%d", c[0], c[1], pb);
            ok &= (table->fwd[c[0]][c[1]][pb] = get(name)) != nullptr;
            snprintf(name, sizeof name, "ocs_fwdg_gen_u%d_b%d_p%d", c[0], c[1], pb);
            ok &= (table->fwdg[c[0]][c[1]][pb] = get(name)) != nullptr;
            snprintf(name, sizeof name, "ocs_revg_gen_u%d_b%d_p%d", c[0], c[1], pb);
            ok &= (table->revg[c[0]][c[1]][pb] = get(name)) != nullptr;
        }
        for (auto &c : combos_r) {
            snprintf(name, sizeof name, "ocs_rev_gen_u%d_b%d_p%d", c[0], c[1], pb);
            ok &= (table->rev[c[0]][c[1]][pb] = get(name)) != nullptr;
        }
        snprintf(name, sizeof name, "ocs_fbrev_gen_p%d", pb);
        ok &= (table->fbrev[pb] = get(name)) != nullptr;
    }
    ok &= (table->tt = get("ocs_tt_gen")) != nullptr;
    const void *info = get("ocs_info_gen");
    if (!ok || !info) {
        *log += "\ncompiled library lacks an expected kernel";
        cudaLibraryUnload(lib);
        return OCS_B200_ERR_COMPILE;
    }
    // ask the compiled code for its hoisted-table widths and tile constants
    int *d_info = nullptr, inf[8] = {0};
    e = cudaMalloc(&d_info, sizeof inf);
    if (e == cudaSuccess) {
        void *args[] = {(void *)&d_info};
        e = cudaLaunchKernel(info, dim3(1), dim3(1), args, 0, nullptr);
        if (e == cudaSuccess) e = cudaMemcpy(inf, d_info, sizeof inf, cudaMemcpyDeviceToHost);
        cudaFree(d_info);
    }
    if (e != cudaSuccess) {
        *log += std::string("\nocs_gen_info launch: ") + cudaGetErrorString(e);
        cudaLibraryUnload(lib);
        return OCS_B200_ERR_CUDA;
    }
    for (int hm = 0; hm < 3; ++hm) table->ntt[hm] = inf[3 + hm];
    table->tile = inf[6];
    table->threads = inf[7];
    *library = (void *)lib;
    return 0;
}

void ocs_unload_problem(void *library)
{
    if (library) cudaLibraryUnload((cudaLibrary_t)library);
}
