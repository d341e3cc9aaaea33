// ocs_api.cu -- host side of the C ABI declared in include/ocs_b200.h.
//
// Owns the device-resident trajectory store (the xK cache of
// Integrator/RK4Integrator.m:8-10, node states only), packs the batch-invariant
// grid / control tables into per-grid-point records, picks the sweep kernel variant
// and launches it.  No arithmetic of the path runs on the host except forming the
// grid (diff / midpoints, RK4Integrator.m:17-24), h/2, h/3, h/6 and -- for a control
// shared by the whole batch -- the product u = v*B for the G grid points.
// There is no CPU fallback: without a CUDA device every compute call fails.
#include "ocs_internal.h"
#include "solution_interp.cuh"

#define OCS_NS ocs_host
#include "rk4_kernels.cuh"
using ocs_host::SweepArgs;
using ocs_host::EqArgs;

extern "C" void ocs_builtin_table_test_fast(ocs_kernel_table *t);
extern "C" void ocs_builtin_table_test_fastv(ocs_kernel_table *t);
extern "C" void ocs_builtin_table_test_strict(ocs_kernel_table *t);
// nvrtc_problem.cu
int ocs_compile_problem(const char *source, int n, int nc, int np, int strict, const char *include_dir,
                        ocs_kernel_table *table, void **library, std::string *log);
int ocs_compile_unit(const char *source, const char *unit, int strict, const char *include_dir,
                     const char *const *names, int nnames, const void **kernels, void **library, std::string *log,
                     int n_states);
void ocs_unload_problem(void *library);

namespace {

using ocs_detail::Segment;
using ocs_detail::DevBuf;
using ocs_detail::OCS_FOLD_MAX;
using namespace ocs_detail;

std::string g_create_error = "";
std::mutex g_mutex;

int ensure(H *h, DevBuf &b, size_t n)
{
    if (b.cap >= n && b.p) return 0;
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    cudaError_t e = cudaMalloc(&b.p, std::max<size_t>(n, 1) * sizeof(double));
    if (e != cudaSuccess) {
        b.p = nullptr;
        cudaGetLastError();
        return h->fail(OCS_B200_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", n * sizeof(double),
                       cudaGetErrorString(e));
    }
    b.cap = n;
    return 0;
}

int ensure_raw(H *h, double **p, size_t *cap, size_t n)
{
    DevBuf b;
    b.p = *p;
    b.cap = *cap;
    int rc = ensure(h, b, n);
    *p = b.p;
    *cap = b.cap;
    return rc;
}

// ---------------------------------------------------------------- caller buffers
// Host <-> device copies of caller-owned host arrays.  Pinned / registered memory goes straight to the copy
// engine; large pageable arrays (every MATLAB mxArray, every NumPy array) go through the handle's pinned ring and
// copy threads (host_stager.h).  Everything is ordered on `st`; the data is in place after finish_host().
int h2d(H *h, void *dst, const void *src, size_t bytes, cudaStream_t st = nullptr)
{
    if (bytes == 0) return 0;
    if (!st) st = h->stream;
    if (bytes >= HostStager::kMinBytes && HostStager::enabled() && !HostStager::is_pinned(src)) {
        CK(h->stager.h2d(h->device, dst, src, bytes, st));
        h->staged = true;
        return 0;
    }
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return 0;
}
int d2h(H *h, void *dst, const void *src, size_t bytes, cudaStream_t st = nullptr)
{
    if (bytes == 0) return 0;
    if (!st) st = h->stream;
    if (bytes >= HostStager::kMinBytes && HostStager::enabled() && !HostStager::is_pinned(dst)) {
        CK(h->stager.d2h(h->device, dst, src, bytes, st));
        h->staged = true;
        return 0;
    }
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    return 0;
}
// The end of every host-buffer call: streams drained, staged copies delivered, and the flag a kernel raises when it
// was compiled for another record layout than the host packed (it then returns without writing anything) turned
// into an error instead of stale results.
int finish_host(H *h)
{
    if (h->copy_stream) CK(cudaStreamSynchronize(h->copy_stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->staged) {
        h->staged = false;
        CK(h->stager.finish());
    }
    if (h->h_errflag && *h->h_errflag) {
        *h->h_errflag = 0;
        h->have_states = false;
        return h->fail(OCS_B200_ERR_STATE, "a sweep kernel was compiled for another record layout than the host packed "
                                           "(stale kernel variant or OCS_B200_NVRTC_DEFS mismatch); no results were written");
    }
    return 0;
}

// ---------------------------------------------------------------- layout kernels
// dst[perm(r)][ldd] (SoA, batch fastest)  <-  src[b][rows] (MATLAB column-major rows x nb)
// perm(r) = (r % nc) * (rows / nc) + r / nc   when nc > 1 regroups [c + nc*j] -> [c*G + j]
__global__ void to_soa_kernel(const double *__restrict__ src, double *__restrict__ dst, long rows, long nb, long ldd,
                              int nc)
{
    __shared__ double tile[32][33];
    const long b0 = (long)blockIdx.x * 32, r0 = (long)blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const long b = b0 + k, r = r0 + threadIdx.x;
        if (b < nb && r < rows) tile[k][threadIdx.x] = src[(size_t)b * rows + r];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const long r = r0 + k, b = b0 + threadIdx.x;
        if (b < nb && r < rows) {
            const long rp = nc > 1 ? (r % nc) * (rows / nc) + r / nc : r;
            dst[(size_t)rp * ldd + b] = tile[threadIdx.x][k];
        }
    }
}
// inverse: dst[b][rows] <- src[perm(r)][lds]
__global__ void from_soa_kernel(const double *__restrict__ src, double *__restrict__ dst, long rows, long nb, long lds,
                                int nc)
{
    __shared__ double tile[32][33];
    const long b0 = (long)blockIdx.x * 32, r0 = (long)blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const long r = r0 + k, b = b0 + threadIdx.x;
        if (b < nb && r < rows) {
            const long rp = nc > 1 ? (r % nc) * (rows / nc) + r / nc : r;
            tile[k][threadIdx.x] = src[(size_t)rp * lds + b];
        }
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const long b = b0 + k, r = r0 + threadIdx.x;
        if (b < nb && r < rows) dst[(size_t)b * rows + r] = tile[threadIdx.x][k];
    }
}

int to_soa(H *h, const double *d_src, double *d_dst, long rows, long nb, long ldd, int nc = 1)
{
    if (rows == 0 || nb == 0) return 0;
    dim3 grid((unsigned)((nb + 31) / 32), (unsigned)((rows + 31) / 32)), block(32, 8);
    to_soa_kernel<<<grid, block, 0, h->stream>>>(d_src, d_dst, rows, nb, ldd, nc);
    CK(cudaGetLastError());
    return 0;
}
int from_soa(H *h, const double *d_src, double *d_dst, long rows, long nb, long lds, int nc = 1,
             cudaStream_t stream = nullptr)
{
    if (rows == 0 || nb == 0) return 0;
    dim3 grid((unsigned)((nb + 31) / 32), (unsigned)((rows + 31) / 32)), block(32, 8);
    from_soa_kernel<<<grid, block, 0, stream ? stream : h->stream>>>(d_src, d_dst, rows, nb, lds, nc);
    CK(cudaGetLastError());
    return 0;
}

// host [rows x nb] (column-major) -> device SoA [rows][nb]
int upload_soa(H *h, const double *host, DevBuf &dst, long rows, long nb, int nc = 1)
{
    TRY(ensure(h, dst, (size_t)rows * nb));
    if (rows == 1) return h2d(h, dst.p, host, sizeof(double) * nb);
    TRY(ensure(h, h->d_stage, (size_t)rows * nb));
    TRY(h2d(h, h->d_stage.p, host, sizeof(double) * rows * nb));
    return to_soa(h, h->d_stage.p, dst.p, rows, nb, nb, nc);
}
// device SoA [rows][ld] (first nb columns from b0) -> host [rows x nb]
int download_soa(H *h, const double *d_src, long ld, long b0, double *host, long rows, long nb, int nc = 1)
{
    if (rows == 1) return d2h(h, host, d_src + b0, sizeof(double) * nb);
    TRY(ensure(h, h->d_stage, (size_t)rows * nb));
    TRY(from_soa(h, d_src + b0, h->d_stage.p, rows, nb, ld, nc));
    return d2h(h, host, h->d_stage.p, sizeof(double) * rows * nb);
}

// ---------------------------------------------------------------- batch reductions (warp shuffles)
struct RedPair {
    double v;
    long i;
};
template <int OP>
__device__ __forceinline__ RedPair red_combine(RedPair a, RedPair b)
{
    if (OP == OCS_B200_REDUCE_SUM) return RedPair{a.v + b.v, 0};
    bool take_b;
    if (OP == OCS_B200_REDUCE_MIN)
        take_b = (b.v < a.v) || (b.v == a.v && b.i < a.i);
    else
        take_b = (b.v > a.v) || (b.v == a.v && b.i < a.i);
    return take_b ? b : a;
}
template <int OP>
__device__ __forceinline__ RedPair warp_reduce(RedPair x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        RedPair y;
        y.v = __shfl_down_sync(0xffffffffu, x.v, o);
        y.i = __shfl_down_sync(0xffffffffu, x.i, o);
        x = red_combine<OP>(x, y);
    }
    return x;
}
// one block per partial; final pass with one block.  Deterministic (fixed tree).
template <int OP>
__global__ void reduce_kernel(const double *__restrict__ J, const long *__restrict__ idx_in, long nvals, double *out_v,
                              long *out_i)
{
    __shared__ RedPair ws[32];
    const double ident = OP == OCS_B200_REDUCE_SUM ? 0.0 : (OP == OCS_B200_REDUCE_MIN ? INFINITY : -INFINITY);
    RedPair acc{ident, 0x7fffffffffffffffL};
    for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < nvals; k += (long)gridDim.x * blockDim.x) {
        RedPair x{J[k], idx_in ? idx_in[k] : k};
        acc = red_combine<OP>(acc, x);
    }
    acc = warp_reduce<OP>(acc);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) ws[w] = acc;
    __syncthreads();
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        RedPair x = lane < nw ? ws[lane] : RedPair{ident, 0x7fffffffffffffffL};
        x = warp_reduce<OP>(x);
        if (lane == 0) {
            out_v[blockIdx.x] = x.v;
            out_i[blockIdx.x] = x.i;
        }
    }
}

// fb_sweep bookkeeping: freeze trajectories whose normalised control change is <= 1
__global__ void fb_update_kernel(const double *__restrict__ maxchg, int *frozen, int *sweeps, int sweep, long nb,
                                 int *n_open)
{
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    int open = 0;
    if (b < nb && !frozen[b]) {
        sweeps[b] = sweep;
        if (maxchg[b] <= 1.0)
            frozen[b] = 1;   // fb_sweep.m:110-111
        else
            open = 1;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, open);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_open, __popc(bal));
}

// Spectral projected gradient update, one trajectory per thread (rows of z are coalesced across threads).
//   z, g, J   : accepted point, its gradient and objective
//   zt, gt, Jt: trial point the sweeps just evaluated
//   Jhist     : the last SPG_M accepted objectives (non-monotone reference value)
// first != 0: the trial IS the starting point (accepted unconditionally).
#define SPG_M 5
__global__ void spg_update_kernel(long nz, long nb, double *z, double *g, double *zt, double *gt, double *J,
                                  const double *Jt, double *Jhist, double *alpha, const double *lb, const double *ub,
                                  int *frozen, int *iters, int iter, double tol, int first, int *n_open)
{
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    int open = 0;
    if (b < nb && !frozen[b]) {
        bool accept = first != 0;
        double a = alpha[b];
        if (!first) {
            double pred = 0.0;                                   // g.(z - zt) >= 0 for a projected step
            for (long k = 0; k < nz; ++k) pred += g[k * nb + b] * (z[k * nb + b] - zt[k * nb + b]);
            double Jref = -INFINITY;
            for (int j = 0; j < SPG_M; ++j) Jref = fmax(Jref, Jhist[j * nb + b]);
            accept = isfinite(Jt[b]) && Jt[b] <= Jref - 1e-4 * pred;
        }
        if (accept) {
            double sts = 0.0, sty = 0.0, pg = 0.0, gmax = 0.0;
            for (long k = 0; k < nz; ++k) {
                const size_t i = (size_t)k * nb + b;
                const double zn = zt[i], gn = gt[i];
                const double sk = zn - z[i], yk = gn - g[i];
                sts += sk * sk;
                sty += sk * yk;
                // (the accepted point stays where it is: the host swaps the roles of (z, g) and (zt, gt) after this kernel)
                const double proj = fmin(ub[k], fmax(lb[k], zn - gn));
                pg = fmax(pg, fabs(zn - proj));
                gmax = fmax(gmax, fabs(gn));
            }
            J[b] = Jt[b];
            Jhist[(size_t)(iter % SPG_M) * nb + b] = Jt[b];
            if (first) {
                for (int j = 0; j < SPG_M; ++j) Jhist[(size_t)j * nb + b] = Jt[b];
                a = 1.0 / fmax(gmax, 1e-8);
            } else {
                a = sty > 0.0 ? fmin(1e10, fmax(1e-10, sts / sty)) : fmin(1e10, 10.0 * a);
            }
            if (pg <= tol) frozen[b] = 1;
        } else {
            a *= 0.25;                                           // rejected: shorter step from the same point
            if (a < 1e-14) frozen[b] = 2;                        // stalled
        }
        alpha[b] = a;
        iters[b] = iter;
        // Roles after this kernel (the host swaps the pointers): (zt, gt) hold the accepted point, z receives the new
        // trial point and g the gradient the next sweep writes.  An accepted trial is therefore not copied at all; a
        // rejected one is replaced by the old accepted point.
        const bool fz = frozen[b] != 0;
        if (!fz) open = 1;
        for (long k = 0; k < nz; ++k) {
            const size_t i = (size_t)k * nb + b;
            double zc, gc;                                       // the accepted point
            if (accept) {
                zc = zt[i];
                gc = gt[i];
            } else {
                zc = z[i];
                gc = g[i];
                zt[i] = zc;
                gt[i] = gc;
            }
            z[i] = fz ? zc : fmin(ub[k], fmax(lb[k], zc - a * gc));   // frozen: the minimiser in both buffers
        }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, open);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_open, __popc(bal));
}

// The same update with the rows of a trajectory split over G thread groups (block = 32 trajectories x G groups,
// group kg owns rows kg, kg + G, ...): every row of zt, gt, z, g is read once and only the new trial point is written
// (5 passes over the (nc K) x Bt arrays; the accepted point changes buffers by a pointer swap on the host) -- the trial rows wait in
// registers while the Barzilai-Borwein sums are reduced across the groups through shared memory.  The Armijo
// term g.(z - zt) of the NEXT acceptance test is formed while zt is written and kept in pred[b].  Sums run in a
// different (fixed) order than in spg_update_kernel, so the two differ by rounding only.
template <int VPT, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) spg_update_split_kernel(
    long nz, long nb, double *__restrict__ z, double *__restrict__ g, double *__restrict__ zt, double *__restrict__ gt,
    double *J, const double *Jt, double *Jhist, double *alpha, double *pred, const double *__restrict__ lb,
    const double *__restrict__ ub, int *frozen, int *iters, int iter, double tol, int first, int *n_open)
{
    extern __shared__ double spg_sm[];                   // [5][G][32]
    const int lane = threadIdx.x, kg = threadIdx.y, G = blockDim.y;
    const long b = (long)blockIdx.x * 32 + lane;
    const bool live = b < nb && !frozen[b];
    bool accept = first != 0;
    double a = 0.0, Jtb = 0.0;
    if (live) {
        a = alpha[b];
        Jtb = Jt[b];
        if (!first) {
            double Jref = -INFINITY;
            for (int j = 0; j < SPG_M; ++j) Jref = fmax(Jref, Jhist[(size_t)j * nb + b]);
            accept = isfinite(Jtb) && Jtb <= Jref - 1e-4 * pred[b];
        }
    }
    double zn[VPT], gn[VPT];
    double sts = 0.0, sty = 0.0, pg = 0.0, gmax = 0.0;
    if (live) {
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const long k = kg + (long)G * j;
            zn[j] = gn[j] = 0.0;
            if (k < nz) {
                const size_t i = (size_t)k * nb + b;
                if (accept) {
                    zn[j] = zt[i];
                    gn[j] = gt[i];
                    const double sk = zn[j] - z[i], yk = gn[j] - g[i];
                    sts += sk * sk;
                    sty += sk * yk;
                    // (z and g are overwritten in the second phase: with no store in this loop the compiler is free to
                    //  issue all 4 x VPT loads of the thread up front)
                    const double proj = fmin(ub[k], fmax(lb[k], zn[j] - gn[j]));
                    pg = fmax(pg, fabs(zn[j] - proj));
                    gmax = fmax(gmax, fabs(gn[j]));
                } else {                     // rejected: a shorter step from the accepted point
                    zn[j] = z[i];
                    gn[j] = g[i];
                }
            }
        }
    }
    double *r0 = spg_sm, *r1 = r0 + G * 32, *r2 = r1 + G * 32, *r3 = r2 + G * 32, *r4 = r3 + G * 32;
    r0[kg * 32 + lane] = sts;
    r1[kg * 32 + lane] = sty;
    r2[kg * 32 + lane] = pg;
    r3[kg * 32 + lane] = gmax;
    __syncthreads();
    sts = sty = pg = gmax = 0.0;
    for (int q = 0; q < G; ++q) {
        sts += r0[q * 32 + lane];
        sty += r1[q * 32 + lane];
        pg = fmax(pg, r2[q * 32 + lane]);
        gmax = fmax(gmax, r3[q * 32 + lane]);
    }
    int fz = 0;
    if (live) {
        if (accept) {
            if (first) a = 1.0 / fmax(gmax, 1e-8);
            else a = sty > 0.0 ? fmin(1e10, fmax(1e-10, sts / sty)) : fmin(1e10, 10.0 * a);
            if (pg <= tol) fz = 1;
        } else {
            a *= 0.25;
            if (a < 1e-14) fz = 2;                               // stalled
        }
    }
    double pr = 0.0;
    if (live) {
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const long k = kg + (long)G * j;
            if (k < nz) {
                const size_t i = (size_t)k * nb + b;
                const double zn_new = fz ? zn[j] : fmin(ub[k], fmax(lb[k], zn[j] - a * gn[j]));   // frozen: the minimiser stays
                // Roles after this kernel (the host swaps the pointers): (zt, gt) hold the accepted point and z receives
                // the new trial point.  An accepted trial is not copied at all (5 passes over the coefficient arrays
                // instead of 7); a rejected one is replaced by the old accepted point.
                if (!accept) {
                    zt[i] = zn[j];
                    gt[i] = gn[j];
                }
                z[i] = zn_new;
                pr += gn[j] * (zn[j] - zn_new);
            }
        }
    }
    r4[kg * 32 + lane] = pr;
    __syncthreads();
    if (kg == 0) {
        int open = 0;
        if (live) {
            pr = 0.0;
            for (int q = 0; q < G; ++q) pr += r4[q * 32 + lane];
            pred[b] = pr;
            alpha[b] = a;
            iters[b] = iter;
            if (accept) {
                J[b] = Jtb;
                Jhist[(size_t)(iter % SPG_M) * nb + b] = Jtb;
                if (first)
                    for (int j = 0; j < SPG_M; ++j) Jhist[(size_t)j * nb + b] = Jtb;
            }
            if (fz) frozen[b] = fz;
            else open = 1;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, open);
        if (lane == 0 && bal) atomicAdd(n_open, __popc(bal));
    }
}

// The split update without the register-held rows (the default).  spg_update_split_kernel keeps a trajectory's trial
// rows in registers between its two phases (2 x 16 doubles per thread: 124 registers, two 224-thread blocks per SM at
// K = 101) and reaches 2.4 TB/s.  Here phase 1 only accumulates and phase 2 reads the accepted rows a second time: a
// block's rows (52 KB at K = 101) are still in the 126 MB L2, so DRAM sees the same 5 passes, no thread holds more than
// the UNR rows in flight, and twice the warps are resident: 4.3 TB/s (profiles/r2_ab_spg_update.txt; wider blocks of
// 64 / 128 adjacent trajectories and 1344-1792 resident threads at 36-40 registers measured 3-30 % slower than
// this shape, 4 / 8 rows in flight per thread 7 / 65 % slower than 2).  Same row groups and summation order as the
// split kernel: bit-identical results.
template <int TW, int UNR>
__global__ void __launch_bounds__(TW * 512, 2 / TW) spg_update_l2_kernel(
    long nz, long nb, double *z, const double *g, double *zt, double *gt,
    double *J, const double *Jt, double *Jhist, double *alpha, double *pred, const double *__restrict__ lb,
    const double *__restrict__ ub, int *frozen, int *iters, int iter, double tol, int first, int *n_open)
{
    extern __shared__ double spg_sm[];                   // [5][G][32 TW]
    constexpr int W = 32 * TW;
    const int lane = threadIdx.x, kg = threadIdx.y, G = blockDim.y;
    const long b = (long)blockIdx.x * W + lane;
    const bool live = b < nb && !frozen[b];
    bool accept = first != 0;
    double a = 0.0, Jtb = 0.0;
    if (live) {
        a = alpha[b];
        Jtb = Jt[b];
        if (!first) {
            double Jref = -INFINITY;
            for (int j = 0; j < SPG_M; ++j) Jref = fmax(Jref, Jhist[(size_t)j * nb + b]);
            accept = isfinite(Jtb) && Jtb <= Jref - 1e-4 * pred[b];
        }
    }
    double sts = 0.0, sty = 0.0, pg = 0.0, gmax = 0.0;
    if (live && accept) {
#pragma unroll UNR
        for (long k = kg; k < nz; k += G) {
            const size_t i = (size_t)k * nb + b;
            const double zn = zt[i], gn = gt[i];
            const double sk = zn - z[i], yk = gn - g[i];
            sts += sk * sk;
            sty += sk * yk;
            const double proj = fmin(ub[k], fmax(lb[k], zn - gn));
            pg = fmax(pg, fabs(zn - proj));
            gmax = fmax(gmax, fabs(gn));
        }
    }
    double *r0 = spg_sm, *r1 = r0 + G * W, *r2 = r1 + G * W, *r3 = r2 + G * W, *r4 = r3 + G * W;
    r0[kg * W + lane] = sts;
    r1[kg * W + lane] = sty;
    r2[kg * W + lane] = pg;
    r3[kg * W + lane] = gmax;
    __syncthreads();
    sts = sty = pg = gmax = 0.0;
    for (int q = 0; q < G; ++q) {
        sts += r0[q * W + lane];
        sty += r1[q * W + lane];
        pg = fmax(pg, r2[q * W + lane]);
        gmax = fmax(gmax, r3[q * W + lane]);
    }
    int fz = 0;
    if (live) {
        if (accept) {
            if (first) a = 1.0 / fmax(gmax, 1e-8);
            else a = sty > 0.0 ? fmin(1e10, fmax(1e-10, sts / sty)) : fmin(1e10, 10.0 * a);
            if (pg <= tol) fz = 1;
        } else {
            a *= 0.25;
            if (a < 1e-14) fz = 2;                               // stalled
        }
    }
    double pr = 0.0;
    if (live) {
        const double *zs = accept ? zt : z, *gs = accept ? gt : g;   // the accepted point
#pragma unroll UNR
        for (long k = kg; k < nz; k += G) {
            const size_t i = (size_t)k * nb + b;
            const double zn = zs[i], gn = gs[i];
            const double zn_new = fz ? zn : fmin(ub[k], fmax(lb[k], zn - a * gn));
            if (!accept) {
                zt[i] = zn;
                gt[i] = gn;
            }
            z[i] = zn_new;
            pr += gn * (zn - zn_new);
        }
    }
    r4[kg * W + lane] = pr;
    __syncthreads();
    if (kg == 0) {
        int open = 0;
        if (live) {
            pr = 0.0;
            for (int q = 0; q < G; ++q) pr += r4[q * W + lane];
            pred[b] = pr;
            alpha[b] = a;
            iters[b] = iter;
            if (accept) {
                J[b] = Jtb;
                Jhist[(size_t)(iter % SPG_M) * nb + b] = Jtb;
                if (first)
                    for (int j = 0; j < SPG_M; ++j) Jhist[(size_t)j * nb + b] = Jtb;
            }
            if (fz) frozen[b] = fz;
            else open = 1;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, open);
        if ((lane & 31) == 0 && bal) atomicAdd(n_open, __popc(bal));
    }
}

// ControlBounds on the grid as constraints of the batched solve (functions/single_shooting.m:106-111: A v <= b from
// control.compute_lincon(prob.ControlBounds); for a dense basis A = [kron(B', I); -kron(B', I)], b = [ub; -lb], i.e.
// lb_c <= u_c(t_j) <= ub_c at every grid point).  Augmented Lagrangian (Powell-Hestenes-Rockafellar): the evaluation
// adds  psi = ( max(0, mu+ + rho (u - ub))^2 - mu+^2 + max(0, mu- + rho (lb - u))^2 - mu-^2 ) / (2 rho)  per grid
// value to J and  max(0, mu+ + rho (u - ub)) - max(0, mu- + rho (lb - u))  to dJdu, on the two SoA grids the
// dense-basis gradient path materialises anyway (u from the forward sweep, dJdu before its projection onto B).
// One thread per trajectory: every access is coalesced across the warp.
__global__ void __launch_bounds__(128) cgb_penalty_kernel(const double *__restrict__ ug, double *__restrict__ dug,
                                                          const double *__restrict__ mu, const double *__restrict__ lohi,
                                                          double rho, int nc, int G, long nb, long ldb, double *J,
                                                          const int *frozen)
{
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb || (frozen && frozen[b])) return;
    const double half_inv = 0.5 / rho;
    double P = 0.0;
    for (int c = 0; c < nc; ++c) {
        const double lo = lohi[2 * c], hi = lohi[2 * c + 1];
        const double *up = ug + (size_t)c * G * ldb + b, *mp = mu + (size_t)c * G * ldb + b,
                     *mm = mu + (size_t)(nc + c) * G * ldb + b;
        double *dp = dug + (size_t)c * G * ldb + b;
        for (int j = 0; j < G; ++j) {
            const size_t o = (size_t)j * ldb;
            const double u = up[o], a = mp[o], d = mm[o];
            const double tp = fmax(0.0, a + rho * (u - hi)), tm = fmax(0.0, d + rho * (lo - u));
            P += half_inv * ((tp * tp - a * a) + (tm * tm - d * d));
            if (tp != 0.0 || tm != 0.0) dp[o] += tp - tm;
        }
    }
    J[b] += P;
}

// End of an outer iteration: multipliers mu <- max(0, mu + rho (u - ub)) (and the lower twin) at the accepted points,
// the largest bound violation per trajectory, and trajectories that still violate the bounds are reopened.
__global__ void __launch_bounds__(128) cgb_update_kernel(const double *__restrict__ ug, double *__restrict__ mu,
                                                         const double *__restrict__ lohi, double rho, int nc, int G, long nb,
                                                         long ldb, double *viol, int *frozen, double feas_tol, int *n_reopened,
                                                         unsigned long long *worst_bits)
{
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    int reopened = 0;
    if (b < nb) {
        double worst = 0.0;
        for (int c = 0; c < nc; ++c) {
            const double lo = lohi[2 * c], hi = lohi[2 * c + 1];
            const double *up = ug + (size_t)c * G * ldb + b;
            double *mp = mu + (size_t)c * G * ldb + b, *mm = mu + (size_t)(nc + c) * G * ldb + b;
            for (int j = 0; j < G; ++j) {
                const size_t o = (size_t)j * ldb;
                const double u = up[o];
                worst = fmax(worst, fmax(u - hi, lo - u));
                mp[o] = fmax(0.0, mp[o] + rho * (u - hi));
                mm[o] = fmax(0.0, mm[o] + rho * (lo - u));
            }
        }
        viol[b] = worst;
        if (isfinite(worst)) atomicMax(worst_bits, (unsigned long long)__double_as_longlong(worst));   // worst >= 0: the bit pattern is monotone
        if (worst > feas_tol && frozen[b] != 0) {      // (not finite J etc. never freeze as converged: left alone)
            frozen[b] = 0;
            reopened = 1;
        } else if (worst > feas_tol) {
            reopened = 1;
        }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, reopened);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_reopened, __popc(bal));
}

// dJdv = dJdu * B' for a dense basis (ChebyshevControl.m:34-37):
// dug [nc][G][ldb] (SoA, written by the reverse sweep), basrec [G][BW] = B(:,j), dJdv [(k*nc+c)][ldb].
// HBM-bound (every dJdu value is read once, K FMAs each), so what matters is bytes in flight and shared-memory
// reads per FMA.  A block is 32 trajectories (lanes) x PROJ_W warps; warp w sums the grid points of segment w of
// [0, G).  A thread owns one trajectory and NS controls (2 when the problem has at least two): the row B(:,j) it
// reads from the warp-private shared-memory tile (LDS.128: every lane reads the same entries, a broadcast) feeds
// NS*KC FMAs, and 8 grid points of dJdu are in flight per stream.  The PROJ_W partial sums of a trajectory are added
// in ascending segment order through shared memory (fixed order: deterministic).  History (cfg3, 100k x 2001 x K=17,
// 3.2 GB): 4 trajectories x 18 accumulators per thread, 254 registers, 196 blocks: 1.61 ms (latency-bound); one
// trajectory and control per thread: 1.79 ms (one LDS per FMA); this form: see profiles/r2_notes.md.
#define PROJ_JT 32
#define PROJ_W 8
template <int KC, int NS>
__global__ void __launch_bounds__(32 * PROJ_W, 2) dense_project_kernel(const double *__restrict__ dug, const double *__restrict__ basrec,
                                                                      int BW, int K, int nc, int G, long nb, long ldb, double *dJdv)
{
    extern __shared__ double2 proj_smem2[];
    double *sm = reinterpret_cast<double *>(proj_smem2);
    const int lane = threadIdx.x, w = threadIdx.y, c0 = blockIdx.y * NS;
    double *Bs = sm + (size_t)w * 2 * PROJ_JT * BW;             // 2 x [PROJ_JT][BW], private to the warp
    const long b = (long)blockIdx.x * 32 + lane;
    const long bb = b < nb ? b : nb - 1;
    const int seg = (G + PROJ_W - 1) / PROJ_W;
    const int ja = w * seg, jb = ja + seg < G ? ja + seg : G;
    const double *dp[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) dp[s] = dug + (size_t)(c0 + s < nc ? c0 + s : nc - 1) * G * ldb + bb;
    const unsigned ld8 = (unsigned)ldb * 8u;
    static_assert(KC % 2 == 0, "KC accumulators are updated in double2 pairs");
    constexpr int R = 8 / NS;                 // grid points per batch and stream: 8 loads per thread and batch
    const int tile = PROJ_JT * BW;            // doubles of one B tile; the warp owns two (the next one is fetched with
    double *Bt0 = Bs, *Bt1 = Bs + tile;       // cp.async while the current one is consumed)
    auto fetch_tile = [&](double *dst, int j0) {
        const int nj = jb - j0 < PROJ_JT ? jb - j0 : PROJ_JT;
        for (int k = 2 * lane; k < nj * BW; k += 64) ocs_host::cp_async_16(dst + k, basrec + (size_t)j0 * BW + k);
        ocs_host::cp_async_commit();
    };
    auto load_batch = [&](double (&d)[NS][R], int j) {      // dJdu of grid points j .. j+R-1 (clamped: they meet zero B rows)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int jc = j + r < G ? j + r : G - 1;
#pragma unroll
            for (int s = 0; s < NS; ++s) d[s][r] = *ocs_host::soa_row(dp[s], (unsigned)jc, ld8);
        }
    };
    for (int k0 = 0; k0 < K; k0 += KC) {
        double acc[NS][KC];
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) acc[s][kk] = 0.0;
        auto fma_batch = [&](const double (&d)[NS][R], const double *Bt, int jj) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double2 *Bj = reinterpret_cast<const double2 *>(Bt + (jj + r) * BW + k0);   // BW, k0 even: 16-byte aligned
#pragma unroll
                for (int kk = 0; kk < KC; kk += 2) {
                    if (k0 + kk < BW) {
                        const double2 bq = Bj[kk >> 1];
#pragma unroll
                        for (int s = 0; s < NS; ++s) {
                            acc[s][kk] = acc[s][kk] + d[s][r] * bq.x;
                            acc[s][kk + 1] = acc[s][kk + 1] + d[s][r] * bq.y;
                        }
                    }
                }
            }
        };
        __syncwarp();
        if (ja < jb) fetch_tile(Bt0, ja);
        double dA[NS][R], dB[NS][R];
        load_batch(dA, ja);
        for (int j0 = ja; j0 < jb; j0 += PROJ_JT) {
            const int nj = jb - j0 < PROJ_JT ? jb - j0 : PROJ_JT;
            double *Bc = ((j0 - ja) / PROJ_JT) & 1 ? Bt1 : Bt0, *Bn = Bc == Bt0 ? Bt1 : Bt0;
            ocs_host::cp_async_wait<0>();
            __syncwarp();                                 // the current tile is complete for every lane; the other one is free
            if (j0 + PROJ_JT < jb) fetch_tile(Bn, j0 + PROJ_JT);
            // rows past the end of the segment inside the last batch multiply by zero
            for (int k = nj * BW + lane; k < ((nj + 2 * R - 1) / (2 * R)) * (2 * R) * BW && k < tile; k += 32) Bc[k] = 0.0;
            __syncwarp();
            // two batches per iteration: the loads of one are in flight while the other is multiplied
#pragma unroll 1
            for (int jj = 0; jj < nj; jj += 2 * R) {
                load_batch(dB, j0 + jj + R);
                fma_batch(dA, Bc, jj);
                load_batch(dA, j0 + jj + 2 * R);
                if (jj + R < nj) fma_batch(dB, Bc, jj + R);
            }
        }
        // partial sums of the PROJ_W segments, added in ascending segment order
        __syncthreads();
        double *red = sm;                                        // [PROJ_W][NS*KC][32]
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) red[((size_t)w * NS * KC + s * KC + kk) * 32 + lane] = acc[s][kk];
        __syncthreads();
        for (int e = w; e < NS * KC; e += PROJ_W) {
            double sum = red[(size_t)e * 32 + lane];
            for (int q = 1; q < PROJ_W; ++q) sum = sum + red[((size_t)q * NS * KC + e) * 32 + lane];
            const int s = e / KC, kk = e % KC;
            if (b < nb && k0 + kk < K && c0 + s < nc) dJdv[(size_t)((k0 + kk) * nc + c0 + s) * ldb + b] = sum;
        }
        __syncthreads();
    }
}

template <int KC, int NS>
static cudaError_t launch_project(cudaStream_t st, const double *dug, const double *basrec, int BW, int K, int nc, int G, long nb,
                                  long ldb, double *dJdv)
{
    const size_t psm = std::max((size_t)PROJ_W * 2 * PROJ_JT * BW, (size_t)PROJ_W * NS * KC * 32) * sizeof(double);
    if (psm > 200 * 1024) return cudaErrorInvalidValue;
    if (psm > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(dense_project_kernel<KC, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm);
        if (e != cudaSuccess) return e;
    }
    const dim3 grid((unsigned)((nb + 31) / 32), (unsigned)((nc + NS - 1) / NS)), block(32, PROJ_W);
    dense_project_kernel<KC, NS><<<grid, block, psm, st>>>(dug, basrec, BW, K, nc, G, nb, ldb, dJdv);
    return cudaGetLastError();
}

// The same projection with B as an OPERAND instead of a shared-memory read.  The kernel above issues one LDS.128 per four
// FMAs, and a broadcast LDS.128 still delivers 512 bytes to the register file (4 cycles of the SM's shared-memory path):
// 9 of them per grid point and warp against 18 SM-cycles of FP64 issue -- it is bound by shared-memory bandwidth
// (short_scoreboard / wait stalls, FP64 pipe 36 % active, 0.45 of HBM).  Here up to 220 columns of B travel in the
// kernel's PARAMETER space (constant bank 0, 32 764 bytes per launch since CUDA 12.1): the compiler reads them with
// LDCU.64 into uniform registers -- once per warp, not once per lane -- and every FMA takes the form
// DFMA acc, d.reuse, UR, acc, which also occupies the FP64 pipe for 2 cycles instead of 3 (profiles/r2_notes.md section 1).
// One thread per trajectory and NS controls, KC accumulators each (128 registers), two batches of 8 / NS grid points in
// flight per stream; G / 220 launches, each adding its columns' contribution to dJdv in ascending grid order (the first
// one stores).  K <= KC only; wider bases keep the kernel above.  Measured (cfg3, same box, profiles/r2_ab_project.txt):
// 1.14 -> 1.08 ms -- the accumulators of 100k trajectories are three quarters of the GPU's registers, so every launch
// runs 1.3 rounds of blocks at 16 warps per SM and the loads' latency is what remains; without the double-buffered
// loads 1.41-1.46 ms, one control per thread at 6 blocks per SM 1.20 ms.
constexpr int PROJC_DOUBLES = 3960;
struct ProjChunk {
    double b[PROJC_DOUBLES];      // [nj][KC]: B(0..K-1, j0 + jj), zero-padded to KC
};
template <int KC, int NS>
__global__ void __launch_bounds__(128, 4) dense_project_const_kernel(const __grid_constant__ ProjChunk B, const double *__restrict__ dug,
                                                                    int K, int nc, int G, int j0, int nj, long nb, long ldb,
                                                                    double *dJdv, int first)
{
    const long b = (long)blockIdx.x * 128 + threadIdx.x;
    const int c0 = blockIdx.y * NS;
    const long bb = b < nb ? b : nb - 1;
    const double *dp[NS];
    int cs[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        cs[s] = c0 + s < nc ? c0 + s : nc - 1;
        dp[s] = dug + ((size_t)cs[s] * G + j0) * ldb + bb;
    }
    double acc[NS][KC];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int k = 0; k < KC; ++k) acc[s][k] = (first || k >= K) ? 0.0 : dJdv[(size_t)(k * nc + cs[s]) * ldb + bb];
    // two batches of R grid points: the loads of one are in flight while the other is multiplied
    constexpr int R = 8 / NS;
    auto load_batch = [&](double (&d)[NS][R], int jj) {      // clamped: entries past nj are loaded but never multiplied
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int jc = jj + r < nj ? jj + r : nj - 1;
#pragma unroll
            for (int s = 0; s < NS; ++s) d[s][r] = dp[s][(size_t)jc * ldb];
        }
    };
    auto fma_batch = [&](const double (&d)[NS][R], int jj) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (jj + r < nj) {                               // uniform
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const double bk = B.b[(jj + r) * KC + k];
#pragma unroll
                    for (int s = 0; s < NS; ++s) acc[s][k] = fma(d[s][r], bk, acc[s][k]);
                }
            }
        }
    };
    double dA[NS][R], dB[NS][R];
    load_batch(dA, 0);
#pragma unroll 1
    for (int jj = 0; jj < nj; jj += 2 * R) {
        load_batch(dB, jj + R);
        fma_batch(dA, jj);
        load_batch(dA, jj + 2 * R);
        fma_batch(dB, jj + R);
    }
    if (b < nb) {
#pragma unroll
        for (int s = 0; s < NS; ++s)
            if (c0 + s < nc)
#pragma unroll
                for (int k = 0; k < KC; ++k)
                    if (k < K) dJdv[(size_t)(k * nc + c0 + s) * ldb + b] = acc[s][k];
    }
}

// bdense: the host's B, K x G column-major (B(k, j) at bdense[k + K j])
template <int KC, int NS>
static cudaError_t launch_project_const(cudaStream_t st, const double *dug, const double *bdense, int K, int nc, int G, long nb, long ldb,
                                        double *dJdv, long *launches)
{
    static_assert((size_t)PROJC_DOUBLES * sizeof(double) + 128 <= 32764, "B chunk must fit the kernel parameter space");
    const int JT = PROJC_DOUBLES / KC;
    const dim3 grid((unsigned)((nb + 127) / 128), (unsigned)((nc + NS - 1) / NS));
    ProjChunk chunk;
    for (int j0 = 0; j0 < G; j0 += JT) {
        const int nj = G - j0 < JT ? G - j0 : JT;
        for (int jj = 0; jj < nj; ++jj)
            for (int k = 0; k < KC; ++k) chunk.b[jj * KC + k] = k < K ? bdense[k + (size_t)K * (j0 + jj)] : 0.0;
        dense_project_const_kernel<KC, NS><<<grid, 128, 0, st>>>(chunk, dug, K, nc, G, j0, nj, nb, ldb, dJdv, j0 == 0);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (launches) ++*launches;
    }
    return cudaSuccess;
}

// ug[c][j][b] = sum_k v[(k*nc+c)][b] B(k,j): control.compute_u for every trajectory (Control/*Control.m compute_u),
// from the basis records of build_records (sparse: the two weights by knot parity; dense: the column B(:,j)), in
// the summation order of ControlEval.  Used once per fb_sweep solve that starts from coefficients.
__global__ void __launch_bounds__(128) basis_eval_kernel(const double *__restrict__ v, const double *__restrict__ basrec,
                                                         int BW, int sparse, int K, int nc, int G, long nb, long ldb,
                                                         double *__restrict__ ug)
{
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    for (int j = 0; j < G; ++j) {
        const double *rec = basrec + (size_t)j * BW;
        for (int c = 0; c < nc; ++c) {
            double u;
            if (sparse) {
                const int k = ocs_host::rec_k(rec);
                const int ke = (k & 1) ? k + 1 : k, ko = (k & 1) ? k : k + 1;
                const double vE = ke < K ? v[(size_t)(ke * nc + c) * ldb + b] : 0.0;
                const double vO = ko < K ? v[(size_t)(ko * nc + c) * ldb + b] : 0.0;
                u = vE * rec[1] + vO * rec[2];
            } else {
                u = v[(size_t)c * ldb + b] * rec[0];
                for (int k = 1; k < K; ++k) u = u + v[(size_t)(k * nc + c) * ldb + b] * rec[k];
            }
            ug[((size_t)c * G + j) * ldb + b] = u;
        }
    }
}

// The dense-basis case of basis_eval_kernel at full speed (a Chebyshev solve that starts from coefficients evaluates
// u0 = v*B for every candidate: 2 nc K G FMAs and nc G stores per trajectory; the generic kernel above re-reads every
// coefficient from L1 for every grid value and took longer than a whole fb_sweep iteration at 10^5 candidates).  A block
// is 128 trajectories x one segment of the grid (blockIdx.y); a control's K <= BE_KC coefficients sit in registers, the
// rows B(:, j) of a 32-point tile in shared memory (every lane reads the same entries: a broadcast), and the sum runs in
// the order of ControlEval / basis_eval_kernel, so the results are bit-identical to theirs.
#define BE_KC 32
#define BE_JT 32
__global__ void __launch_bounds__(128) basis_eval_dense_kernel(const double *__restrict__ v, const double *__restrict__ basrec, int BW,
                                                               int K, int nc, int G, int seg, long nb, long ldb,
                                                               double *__restrict__ ug)
{
    extern __shared__ double be_sm[];                    // [BE_JT][BW]
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long bb = b < nb ? b : nb - 1;
    const int ja = blockIdx.y * seg, jb = ja + seg < G ? ja + seg : G;
    for (int c = 0; c < nc; ++c) {
        double vc[BE_KC];
#pragma unroll
        for (int k = 0; k < BE_KC; ++k) vc[k] = k < K ? v[(size_t)(k * nc + c) * ldb + bb] : 0.0;
        for (int j0 = ja; j0 < jb; j0 += BE_JT) {
            const int nj = jb - j0 < BE_JT ? jb - j0 : BE_JT;
            __syncthreads();
            for (int q = threadIdx.x; q < nj * BW; q += blockDim.x) be_sm[q] = basrec[(size_t)j0 * BW + q];
            __syncthreads();
            for (int jj = 0; jj < nj; ++jj) {
                const double *rec = be_sm + jj * BW;
                double u = vc[0] * rec[0];
#pragma unroll
                for (int k = 1; k < BE_KC; ++k)
                    if (k < K) u = u + vc[k] * rec[k];
                if (b < nb) ug[((size_t)c * G + j0 + jj) * ldb + b] = u;
            }
        }
    }
}

// x grid (MATLAB layout m x NP x nb) read back from the resident node store [n][NP][ld]; the running-objective row is
// not kept there: NaN, except its last entry = J
__global__ void store_to_grid_kernel(const double *__restrict__ xs, long ld, int n, int NP, long nb, const double *__restrict__ J,
                                     double *__restrict__ xgrid)
{
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const int m = n + 1;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = 0; i < NP; ++i) {
        for (int c = 0; c < n; ++c) xgrid[c + (size_t)m * (i + (size_t)NP * b)] = xs[((size_t)c * NP + i) * ld + b];
        xgrid[n + (size_t)m * (i + (size_t)NP * b)] = i == NP - 1 ? J[b] : qnan;
    }
}

__global__ void broadcast_rows_kernel(const double *__restrict__ col, double *dst, long rows, long nb)
{
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    for (long r = 0; r < rows; ++r) dst[(size_t)r * nb + b] = col[r];
}

// ---------------------------------------------------------------- records
int seg_umode(const H *h, int si)
{
    if (si == 1) return 0;
    switch (h->ctrl_mode) {
        case CTRL_V_BATCH: return 1;
        case CTRL_U_BATCH: return 2;
        default: return 0;
    }
}

int basis_mode(const H *h, int si) { return si == 0 ? h->basis_kind : BASIS_NONE; }

// u(:,j) = reshape(v,nc,[]) * B(:,j) for a shared v -- same operation order as ControlEval::eval
void shared_u_from_v(const H *h, int j, double *u)
{
    const int nc = h->kt.nc;
    if (h->basis_kind == BASIS_SPARSE) {
        const int k = h->bk[j];
        for (int c = 0; c < nc; ++c) {
            double s = h->ctrl_shared[(size_t)k * nc + c] * h->bw0[j];
            if (h->bnnz[j] > 1) s = s + h->ctrl_shared[(size_t)(k + 1) * nc + c] * h->bw1[j];
            u[c] = s;
        }
    } else {
        const int K = h->K;
        for (int c = 0; c < nc; ++c) {
            double s = h->ctrl_shared[c] * h->bdense[(size_t)K * j];
            for (int k = 1; k < K; ++k) s = s + h->ctrl_shared[(size_t)k * nc + c] * h->bdense[k + (size_t)K * j];
            u[c] = s;
        }
    }
}

// Packs the batch-invariant tables of one segment:
//   grid-point records [ t | u(nc) | hoisted table | pad ]   (rebuilt when the shared control or parameters change)
//   step records       [ h, h/2, h/3, h/6 ]                  (once)
//   basis records      sparse [ (int k, int two) | w0 | w1 | pad ], dense [ B(:,j) | pad ]   (when the basis changes)
int build_records(H *h, int si)
{
    Segment &s = h->seg[si];
    const int N = s.N, G = 2 * N + 1, nc = h->kt.nc;
    if (s.basis_dirty) {
        const int bm = basis_mode(h, si);
        const int BW = ocs_host::ocs_basis_width(bm, h->K);
        if (BW > 0) {
            std::vector<double> bas((size_t)G * BW, 0.0);
            for (int j = 0; j < G; ++j) {
                double *r = &bas[(size_t)j * BW];
                if (bm == BASIS_SPARSE) {
                    // weights by knot parity (wE: the even knot of {k, k+1}); flags bit0 = two non-zeros,
                    // bit1 = the single non-zero sits on the odd slot
                    const int k = h->bk[j];
                    const bool two = h->bnnz[j] > 1, odd = (k & 1) != 0;
                    // bit2 (on midpoint records): the reverse sweep's knot window moves while it handles this
                    // step's right node or midpoint, i.e. k differs from the grid point processed just before
                    int moves = 0;
                    if (j & 1) {
                        const int kprev = (j + 2 < G) ? h->bk[j + 2] : -0x40000000;   // handled before point j+1
                        moves = (h->bk[j + 1] != kprev) || (h->bk[j] != h->bk[j + 1]);
                    }
                    // bit3 / bit4 (on midpoint records): the CONTROL window of a per-trajectory-coefficient sweep moves
                    // while the reverse / forward sweep evaluates this step's controls (reverse: midpoint after the
                    // right node, left node after the midpoint; forward: midpoint after the left node, right node
                    // after the midpoint), so that the sweeps test one flag per step instead of comparing knots at
                    // every grid point
                    int cmoves_rev = 0, cmoves_fwd = 0;
                    if (j & 1) {
                        cmoves_rev = (h->bk[j] != h->bk[j + 1]) || (h->bk[j - 1] != h->bk[j]);
                        cmoves_fwd = (h->bk[j] != h->bk[j - 1]) || (h->bk[j + 1] != h->bk[j]);
                    }
                    int32_t kk[2] = {k, (two ? 1 : 0) | ((!two && odd) ? 2 : 0) | (moves ? 4 : 0) | (cmoves_rev ? 8 : 0) |
                                            (cmoves_fwd ? 16 : 0)};
                    memcpy(r, kk, sizeof kk);
                    r[1] = odd ? h->bw1[j] : h->bw0[j];
                    r[2] = odd ? h->bw0[j] : h->bw1[j];
                } else {
                    for (int k = 0; k < h->K; ++k) r[k] = h->bdense[k + (size_t)h->K * j];
                }
            }
            if (s.basrec_cap < bas.size()) {
                if (s.d_basrec) CK(cudaFree(s.d_basrec));
                s.d_basrec = nullptr;
                CK(cudaMalloc(&s.d_basrec, bas.size() * sizeof(double)));
                s.basrec_cap = bas.size();
            }
            CK(cudaMemcpyAsync(s.d_basrec, bas.data(), bas.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            CK(cudaStreamSynchronize(h->stream));   // bas is a local vector
        }
        s.BW = BW;
        s.basis_dirty = false;
    }
    if (!s.records_dirty) return 0;
    const int hm = ocs_host::ocs_hoist_mode(seg_umode(h, si), h->params_nbatch > 1);
    const int ntt = h->kt.ntt[hm];
    const int R = ocs_host::ocs_rec_width(nc, ntt);
    const size_t nrec = (size_t)G * R;
    if (s.h_rec_cap < nrec) {
        if (s.h_rec) CK(cudaFreeHost(s.h_rec));
        s.h_rec = nullptr;
        CK(cudaHostAlloc((void **)&s.h_rec, nrec * sizeof(double), cudaHostAllocDefault));
        s.h_rec_cap = nrec;
    }
    if (s.gprec_cap < nrec) {
        if (s.d_gprec) CK(cudaFree(s.d_gprec));
        s.d_gprec = nullptr;
        CK(cudaMalloc(&s.d_gprec, nrec * sizeof(double)));
        s.gprec_cap = nrec;
    }
    if (!s.rec_uploaded) CK(cudaEventCreateWithFlags(&s.rec_uploaded, cudaEventDisableTiming));
    CK(cudaEventSynchronize(s.rec_uploaded));   // the previous upload no longer reads the staging buffer
    double *rec = s.h_rec;
    memset(rec, 0, nrec * sizeof(double));
    std::vector<double> u(nc);
    for (int j = 0; j < G; ++j) {
        double *r = &rec[(size_t)j * R];
        r[0] = s.t[j];
        if (si == 1) {
            for (int c = 0; c < nc; ++c) r[1 + c] = h->uStar[c];   // RK4InfiniteIntegrator.m:15
        } else if (h->ctrl_mode == CTRL_V_SHARED) {
            shared_u_from_v(h, j, u.data());
            for (int c = 0; c < nc; ++c) r[1 + c] = u[c];
        } else if (h->ctrl_mode == CTRL_U_SHARED) {
            for (int c = 0; c < nc; ++c) r[1 + c] = h->ctrl_shared[c + (size_t)nc * j];
        }
    }
    CK(cudaMemcpyAsync(s.d_gprec, rec, nrec * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(s.rec_uploaded, h->stream));
    if (!s.steps_uploaded) {
        std::vector<double> st((size_t)N * 4);
        for (int i = 0; i < N; ++i) {
            st[4 * i] = s.h[i];
            st[4 * i + 1] = s.h[i] / 2;   // obj.h(i)/2   RK4Integrator.m:40
            st[4 * i + 2] = s.h[i] / 3;   // obj.h(i)/3   :77
            st[4 * i + 3] = s.h[i] / 6;   // obj.h(i)/6   :51
        }
        CK(cudaMalloc(&s.d_steprec, st.size() * sizeof(double)));
        CK(cudaMemcpyAsync(s.d_steprec, st.data(), st.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));   // st is a local vector
        s.steps_uploaded = true;
    }
    s.R = R;
    s.hm = hm;
    if (ntt > 0) {
        int hmv = hm, Gv = G;
        void *args[] = {(void *)&s.d_gprec, (void *)&Gv, (void *)&hmv, (void *)&h->d_params.p};
        CK(cudaLaunchKernel(h->kt.tt, dim3((G + 127) / 128), dim3(128), args, 0, h->stream));
        h->launches++;
    }
    s.fold_nf = h->kt.fold ? h->kt.nfold[hm] : 0;
    if (s.fold_nf > 0) {        // stage weights folded into the table entries they multiply, once per step
        if (!s.d_foldrec) CK(cudaMalloc(&s.d_foldrec, (size_t)N * 3 * (OCS_FOLD_MAX + 1) * sizeof(double)));
        if (s.fold_nf > OCS_FOLD_MAX) return h->fail(OCS_B200_ERR_ARG, "too many folded table entries");
        int hmv = hm, Nv = N;
        void *args[] = {(void *)&s.d_gprec, (void *)&s.d_steprec, (void *)&s.d_foldrec, (void *)&Nv, (void *)&hmv};
        CK(cudaLaunchKernel(h->kt.fold, dim3((N + 127) / 128), dim3(128), args, 0, h->stream));
        h->launches++;
    }
    s.records_dirty = false;
    return 0;
}

void mark_dirty(H *h)
{
    h->seg[0].records_dirty = h->seg[1].records_dirty = true;
    h->seg[0].basis_dirty = h->seg[1].basis_dirty = true;
}

int check_ready(H *h)
{
    if (h->sticky) return h->sticky;
    if (!h->have_problem) return h->fail(OCS_B200_ERR_STATE, "no problem bound (ocs_b200_set_problem_*)");
    if (h->params_nbatch == 0 && h->kt.np > 0) return h->fail(OCS_B200_ERR_STATE, "no parameters set (ocs_b200_set_params)");
    if (h->has_tail && (int)h->uStar.size() != h->kt.nc)
        return h->fail(OCS_B200_ERR_STATE, "tail horizon without uStar (ocs_b200_set_tail_control)");
    return 0;
}

// fills the batch-invariant part of a SweepArgs
void base_args(H *h, int si, SweepArgs &a)
{
    Segment &s = h->seg[si];
    memset(&a, 0, sizeof a);
    a.N = s.N;
    a.K = h->K;
    a.RB = s.R;
    a.BW = s.BW;
    a.gprec = s.d_gprec;
    a.steprec = s.d_steprec;
    a.foldrec = s.fold_nf > 0 ? s.d_foldrec : nullptr;
    {   // steps equal to rounding (linspace grids are): see ocs_ntexp in rk4_kernels.cuh
        double dmax = 0.0;
        for (double hi : s.h) dmax = std::max(dmax, std::fabs(hi - s.h[0]));
        a.texp_uniform = dmax <= 1e-12 * std::fabs(s.h[0]) ? 1 : 0;
        // the mean step (the horizon's length is kept exactly); only the OCS_UNIFORM_H kernels read it, and those are
        // only selected for grids whose steps agree to rounding (steps_uniform)
        const double hm = (s.tspan[s.N] - s.tspan[0]) / s.N;
        a.hc[0] = hm;
        a.hc[1] = hm / 2;
        a.hc[2] = hm / 3;
        a.hc[3] = hm / 6;
    }
    a.basrec = s.d_basrec;
    a.p = h->d_params.p;
    a.ldp = h->params_nbatch > 1 ? h->params_nbatch : 0;
    if (h->params_nbatch == 1)      // batch-uniform parameters also travel by value (constant-bank operands in the kernels)
        for (int k = 0; k < h->kt.np && k < OCS_NPC; ++k) a.pc[k] = h->params_shared[k];
    a.error_flag = h->d_errflag;
    a.frozen = h->fb_frozen;
}

// Upper bound of what any kernel variant lays out: the record tile (twice when the next tile is fetched asynchronously
// into a second buffer) followed by `acc_doubles` of per-thread rings.
size_t smem_bytes(const H *h, const SweepArgs &a, size_t acc_doubles)
{
    const int S = h->kt.tile;
    int nf = 0;     // fold records of the tile (same width for every hoisting mode that folds)
    for (int hm = 0; hm < 3; ++hm) nf = std::max(nf, h->kt.fold ? h->kt.nfold[hm] : 0);
    const size_t fold = a.foldrec ? (size_t)S * ((3 * nf + 1) & ~1) : 0;
    const size_t tile = (size_t)(2 * S + 1) * a.RB + 4 * S + (size_t)(2 * S + 1) * a.BW + fold;
    return ((h->kt.async_stage ? 2 : 1) * tile + acc_doubles) * sizeof(double);
}

// acc_doubles: shared-memory doubles per block beyond the record tiles (dense-basis gradient accumulators)
int launch(H *h, const void *kernel, SweepArgs &a, bool is_forward, int tpt = 1, size_t acc_doubles = 0)
{
    if (!kernel) return h->fail(OCS_B200_ERR_ARG, "this control/basis/parameter combination has no kernel variant");
    if (a.nb <= 0) return 0;
    const int threads = h->kt.threads;
    const unsigned blocks = (unsigned)((a.nb + (long)threads * tpt - 1) / ((long)threads * tpt));
    void *args[] = {(void *)&a};
    if (h->profiling) CK(cudaEventRecord(h->ev[0], h->stream));
    const size_t smem = smem_bytes(h, a, acc_doubles);
    if (smem > 48 * 1024) {
        if (smem > 200 * 1024) return h->fail(OCS_B200_ERR_ARG, "control basis too wide for the shared-memory tile");
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    CK(cudaLaunchKernel(kernel, dim3(blocks), dim3(threads), args, smem, h->stream));
    if (h->profiling) {
        CK(cudaEventRecord(h->ev[1], h->stream));
        CK(cudaEventSynchronize(h->ev[1]));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
        (is_forward ? h->ms_fwd : h->ms_rev) += ms;
    }
    h->launches++;
    return 0;
}

int pb_index(const H *h) { return h->params_nbatch > 1 ? 1 : 0; }

// trajectories per thread of the multi-trajectory sweep variants (OCS_FWD_T / OCS_REV_T in rk4_kernels.cuh)
int fwd_tpt(const H *h, int um)
{
    return std::max(h->params_nbatch > 1 ? h->kt.tpt_fwd_pb : (um == 1 ? h->kt.tpt_fwd_coef : h->kt.tpt_fwd), 1);
}
int rev_tpt(const H *h, int um)
{
    return std::max(h->params_nbatch > 1 ? h->kt.tpt_rev_pb : (um == 1 ? h->kt.tpt_rev_coef : h->kt.tpt_rev), 1);
}

// Every segment's steps agree to rounding (any linspace grid does: 1e-12 relative, the bound of the time-exponential
// recurrence): the FMA-contracted kernels may then read ONE step size per segment as a constant-bank operand.
bool steps_uniform(const H *h)
{
    for (int si = 0; si < (h->has_tail ? 2 : 1); ++si) {
        const Segment &s = h->seg[si];
        if (s.N < 1) continue;
        const double hm = (s.tspan[s.N] - s.tspan[0]) / s.N;
        for (double hi : s.h)
            if (!(std::fabs(hi - hm) <= 1e-12 * std::fabs(hm))) return false;
    }
    return true;
}

// flags of ocs_compile_unit / ocs_compile_problem: bit 0 strict arithmetic, bit 1 uniform steps
int nvrtc_flags(const H *h)
{
    const bool strict = h->math_mode == OCS_B200_MATH_STRICT;
    return (strict ? 1 : 0) | ((!strict && steps_uniform(h)) ? 2 : 0);
}

enum { KIND_FWD = 0, KIND_FWDG, KIND_REV, KIND_REVG, KIND_FB, KIND_FBC, KIND_FBCG };

// The kernel of one sweep variant.  Built-in problems have every variant compiled in; generated
// problems compile the variant with NVRTC the first time it is needed.
int resolve(H *h, int kind, int um, int bm, int pb, const void **out)
{
    const void **slot = kind == KIND_FWD ? &h->kt.fwd[um][bm][pb]
                      : kind == KIND_FWDG ? &h->kt.fwdg[um][bm][pb]
                      : kind == KIND_REV ? &h->kt.rev[um][bm][pb]
                      : kind == KIND_REVG ? &h->kt.revg[um][bm][pb]
                      : kind == KIND_FB ? &h->kt.fbrev[pb] : kind == KIND_FBC ? &h->kt.fbcont[pb] : &h->kt.fbcontg[pb];
    if (!*slot && h->is_generated) {
        static const char *macro[] = {"OCS_FWD_KERNEL", "OCS_FWDG_KERNEL", "OCS_REV_KERNEL", "OCS_REVG_KERNEL", "OCS_FB_KERNEL",
                                      "OCS_FBC_KERNEL", "OCS_FBCG_KERNEL"};
        static const char *prefix[] = {"ocs_fwd_", "ocs_fwdg_", "ocs_rev_", "ocs_revg_", "ocs_fbrev_", "ocs_fbcont_", "ocs_fbcontg_"};
        char unit[256], name[96];
        const int kd = bm == 2 ? h->K : 0;
        if (kind == KIND_FB || kind == KIND_FBC || kind == KIND_FBCG) {
            snprintf(unit, sizeof unit, "%s(ocs_gen::Problem, gen, %d)", macro[kind], pb);
            snprintf(name, sizeof name, "%sgen_u2_b0_p%d", prefix[kind], pb);
        } else {
            snprintf(unit, sizeof unit, "%s(ocs_gen::Problem, gen, %d, %d, %d, %d)", macro[kind], um, bm, pb, kd);
            snprintf(name, sizeof name, "%sgen_u%d_b%d_p%d", prefix[kind], um, bm, pb);
        }
        const char *names[1] = {name};
        const void *kern[1] = {nullptr};
        void *lib = nullptr;
        std::string log;
        int rc = ocs_compile_unit(h->gen_source.c_str(), unit, nvrtc_flags(h),
                                  h->gen_include.empty() ? nullptr : h->gen_include.c_str(), names, 1, kern, &lib, &log,
                                  h->gen_n);
        if (rc != 0) return h->fail(rc, "kernel %s of the generated problem did not compile:\n%s", name, log.c_str());
        h->gen_libraries.push_back(lib);
        *slot = kern[0];
    }
    if (!*slot) return h->fail(OCS_B200_ERR_ARG, "this control/basis/parameter combination has no kernel variant");
    *out = *slot;
    return 0;
}

int main_umode(const H *h) { return seg_umode(h, 0); }

// forward sweep(s) over trajectories [b0, b0+nb) of the resident batch; xgrid optional (fetch)
int run_forward(H *h, long b0, long nb, int only_segment, double *d_xgrid, bool write_outputs)
{
    const int n = h->kt.n;
    const long Bt = h->Bt;
    for (int si = 0; si < (h->has_tail ? 2 : 1); ++si) {
        if (only_segment >= 0 && si != only_segment) continue;
        TRY(build_records(h, si));
        Segment &s = h->seg[si];
        SweepArgs a;
        base_args(h, si, a);
        a.nb = nb;
        a.ldb = Bt;
        if (a.ldp) a.p = h->d_params.p + b0;
        if (a.frozen) a.frozen += b0;
        a.x0 = (si == 0 ? h->d_x0.p : h->d_xT.p) + b0;
        a.xs = s.d_xs + b0;
        a.xgrid = d_xgrid;
        if (write_outputs) {
            a.J = h->d_J.p + b0;
            a.J_accumulate = si == 1;
            a.status = h->d_status + b0;
            if (si == 0) a.xT = h->d_xT.p + b0;
        }
        int um = si == 0 ? main_umode(h) : 0, bm = 0;
        if (um == 1) {
            bm = h->basis_kind;
            a.v = (h->v_ext ? h->v_ext : h->d_v.p) + b0;
            if (bm == BASIS_DENSE) {     // keep the evaluated control grid for the reverse sweep
                TRY(ensure(h, h->d_ug, (size_t)h->kt.nc * (2 * s.N + 1) * Bt));
                a.ug_out = h->d_ug.p + b0;
            }
        } else if (um == 2) {
            a.ug = h->d_ug.p + b0;
        }
        const void *kern = nullptr;
        TRY(resolve(h, d_xgrid ? KIND_FWDG : KIND_FWD, um, bm, pb_index(h), &kern));
        TRY(launch(h, kern, a, true, (d_xgrid || bm == 2) ? 1 : fwd_tpt(h, um)));
    }
    return 0;
}

// reverse sweep(s): tail first (RK4InfiniteIntegrator.m:28), then the main horizon (:29)
int run_reverse(H *h, long b0, long nb, int only_segment, const double *d_lamT, long ldl, double *d_dJdv,
                double *d_lam0, double *d_lamgrid, double *d_dJdugrid)
{
    const long Bt = h->Bt;
    const int m = h->kt.n + 1;
    for (int si = (h->has_tail ? 1 : 0); si >= 0; --si) {
        if (only_segment >= 0 && si != only_segment) {
            // fetching the main grid still needs the tail's lam(:,1); it is cached in d_lamT_tail
            continue;
        }
        TRY(build_records(h, si));
        Segment &s = h->seg[si];
        SweepArgs a;
        base_args(h, si, a);
        a.nb = nb;
        a.ldb = Bt;
        if (a.ldp) a.p = h->d_params.p + b0;
        if (a.frozen) a.frozen += b0;
        a.xs = s.d_xs + b0;
        int um = 0, bm = 0;
        bool project = false;
        a.lamJ_one = 1;      // lamT = e_m (RK4Integrator.m:63-66) unless the caller supplies one
        if (si == 1) {
            TRY(ensure(h, h->d_lamT_tail, (size_t)m * Bt));
            a.lam0 = h->d_lamT_tail.p + b0;
            if (only_segment == 1) {
                a.lamgrid = d_lamgrid;
                a.dJdugrid = d_dJdugrid;
                if (h->lams_req && h->lams_segment == 1) {
                    a.lams = h->lams_req;
                    a.ldlams = h->lams_ld;
                }
            }
        } else {
            if (h->has_tail) {
                // lam2(:,1) of the tail sweep, which started from e_m: its last row is still exactly 1
                // (the last row of dFdx_times_vec is zero, OCProblem.m:17-19)
                a.lamT = h->d_lamT_tail.p + b0;
                a.ldl = Bt;
            } else if (d_lamT) {
                a.lamT = d_lamT + (ldl ? b0 : 0);
                a.ldl = ldl;
                a.lamJ_one = 0;      // arbitrary terminal costate: the unfolded adjoint chain
            }
            a.dJdv = d_dJdv;
            a.dJdv_store = h->bk_monotone ? 1 : 0;
            a.lam0 = d_lam0;
            a.lamgrid = d_lamgrid;
            a.dJdugrid = d_dJdugrid;
            if (h->lams_req && h->lams_segment == 0) {
                a.lams = h->lams_req;
                a.ldlams = h->lams_ld;
            }
            um = main_umode(h);
            const bool grids_req = d_lamgrid || d_dJdugrid || a.lams;
            if (d_dJdv && h->basis_kind == BASIS_NONE)
                return h->fail(OCS_B200_ERR_STATE, "dJdv requested but no control basis is attached");
            if (um == 2 && d_dJdv)
                return h->fail(OCS_B200_ERR_ARG, "dJdv needs coefficients v, not a per-trajectory u grid");
            if (h->basis_kind == BASIS_DENSE && !grids_req) {
                // dense basis: the sweep writes dJdu (SoA) and dense_project_kernel forms dJdu*B' afterwards;
                // a per-trajectory control is read from the grid the forward sweep evaluated (d_ug)
                if (um == 1) {
                    um = 2;
                    a.ug = h->d_ug.p + b0;
                }
                if (d_dJdv) {
                    TRY(ensure(h, h->d_dug, (size_t)h->kt.nc * (2 * s.N + 1) * Bt));
                    a.dug = h->d_dug.p + b0;
                    a.dJdv = nullptr;
                    project = true;
                }
            } else if (um == 1) {
                bm = h->basis_kind;
                a.v = (h->v_ext ? h->v_ext : h->d_v.p) + b0;
            } else if (um == 2) {
                a.ug = h->d_ug.p + b0;
            } else if (d_dJdv) {
                bm = h->basis_kind;
            }
        }
        const bool grids = a.lamgrid || a.dJdugrid || a.lams;
        if (grids && um != 1) bm = 0;   // the fetch variants write dJdu itself; only UMODE_COEF needs the basis
        if (grids && si == 0 && h->fb_resident && h->fb_costate == OCS_B200_COSTATE_CONTINUOUS) {
            // soln.lam of a continuous-costate fb_sweep solve is that costate (compute_x_lam.m:11-14), not the discrete
            // adjoint: the fb_sweep kernel's fetch variant re-integrates it and leaves the control grid alone
            if (a.dJdugrid) return h->fail(OCS_B200_ERR_ARG, "dJdu belongs to the discrete adjoint; the continuous costate has none");
            a.ug = h->d_ug.p + b0;
            a.bounds = h->d_bounds;
            a.relax = 1.0;
            a.fb_readonly = 1;
            a.lamT = nullptr;
            a.dJdv = nullptr;
            a.lam0 = nullptr;
            const void *kc = nullptr;
            TRY(resolve(h, KIND_FBCG, 2, 0, pb_index(h), &kc));
            TRY(launch(h, kc, a, false, 1, (size_t)h->kt.fb_ring * (h->kt.n + 2 * h->kt.nc) * h->kt.threads));
            continue;
        }
        const void *kern = nullptr;
        TRY(resolve(h, grids ? KIND_REVG : KIND_REV, um, bm, pb_index(h), &kern));
        // dense-basis kernels keep K coefficients and K accumulators per trajectory in registers: one per thread
        {
            const int tpt = (grids || bm == 2) ? 1 : rev_tpt(h, um);
            // the gradient kernels (not the fetch variants) keep a cp.async ring of node states in shared memory
            // (with a per-trajectory control grid a ring slot also carries the step's two controls)
            const int depth = pb_index(h) ? h->kt.rev_ring_pb : h->kt.rev_ring;
            const size_t ring = grids ? 0 : (size_t)depth * tpt * (h->kt.n + (um == 2 ? 2 * h->kt.nc : 0)) * h->kt.threads;
            TRY(launch(h, kern, a, false, tpt, ring));
        }
        if (project) {
            const int G = 2 * s.N + 1;
            if (h->cgb_active) {       // grid control bounds of the solve: dJdu += d(penalty)/du, J += penalty
                cgb_penalty_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, h->stream>>>(
                    a.ug, a.dug, h->d_cgb_mu.p + b0, h->d_cgb_b.p, h->cgb_rho, h->kt.nc, G, nb, Bt, h->d_J.p + b0,
                    h->fb_frozen ? h->fb_frozen + b0 : nullptr);
                CK(cudaGetLastError());
                h->launches++;
            }
            if (h->profiling) CK(cudaEventRecord(h->ev[2], h->stream));
            // K <= 18: one pass with 18 accumulators per stream; larger K in chunks of 16.  Two controls per thread when
            // the problem has them (the B row read from shared memory then feeds twice the FMAs)
            {
                const int ncv = h->kt.nc;
                cudaError_t pe;
                // K <= 18: B travels as a kernel parameter (dense_project_const_kernel); OCS_B200_PROJECT=lds keeps the
                // shared-memory kernel (A/B)
                static const bool proj_lds = getenv("OCS_B200_PROJECT") && !strcmp(getenv("OCS_B200_PROJECT"), "lds");
                if (h->K <= 18 && !proj_lds) {
                    long nl = 0;
                    pe = ncv >= 2 ? launch_project_const<18, 2>(h->stream, a.dug, h->bdense.data(), h->K, ncv, G, nb, Bt, d_dJdv, &nl)
                                  : launch_project_const<18, 1>(h->stream, a.dug, h->bdense.data(), h->K, ncv, G, nb, Bt, d_dJdv, &nl);
                    h->launches += nl - 1;      // (one launch is counted below)
                } else if (h->K <= 18) pe = ncv >= 2 ? launch_project<18, 2>(h->stream, a.dug, s.d_basrec, s.BW, h->K, ncv, G, nb, Bt, d_dJdv)
                                              : launch_project<18, 1>(h->stream, a.dug, s.d_basrec, s.BW, h->K, ncv, G, nb, Bt, d_dJdv);
                else pe = ncv >= 2 ? launch_project<16, 2>(h->stream, a.dug, s.d_basrec, s.BW, h->K, ncv, G, nb, Bt, d_dJdv)
                                   : launch_project<16, 1>(h->stream, a.dug, s.d_basrec, s.BW, h->K, ncv, G, nb, Bt, d_dJdv);
                if (pe == cudaErrorInvalidValue) return h->fail(OCS_B200_ERR_ARG, "control basis too wide for the projection kernel");
                CK(pe);
            }
            CK(cudaGetLastError());
            if (h->profiling) {
                CK(cudaEventRecord(h->ev[3], h->stream));
                CK(cudaEventSynchronize(h->ev[3]));
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]));
                h->ms_rev += ms;
                h->ms_proj += ms;
            }
            h->launches++;
        }
    }
    return 0;
}

int alloc_batch(H *h, long Bt)
{
    // the kernels address SoA rows with a 32-bit byte stride (soa_row in rk4_kernels.cuh)
    if (Bt >= (1L << 29)) return h->fail(OCS_B200_ERR_ARG, "batches of 2^29 trajectories or more are not supported");
    const int n = h->kt.n, m = n + 1;
    TRY(ensure(h, h->d_x0, (size_t)n * Bt));
    TRY(ensure(h, h->d_J, (size_t)Bt));
    TRY(ensure(h, h->d_xT, (size_t)n * Bt));
    if (h->status_cap < (size_t)Bt) {
        if (h->d_status) CK(cudaFree(h->d_status));
        h->d_status = nullptr;
        CK(cudaMalloc(&h->d_status, sizeof(int) * Bt));
        h->status_cap = Bt;
    }
    for (int si = 0; si < (h->has_tail ? 2 : 1); ++si) {
        Segment &s = h->seg[si];
        TRY(ensure_raw(h, &s.d_xs_base, &s.xs_cap, (size_t)n * (s.N + 1) * Bt + OCS_XS_SLACK * (size_t)Bt));
        s.d_xs = s.d_xs_base + OCS_XS_SLACK * (size_t)Bt;
    }
    (void)m;
    return 0;
}

int validate_control(H *h, long Bt)
{
    const int nc = h->kt.nc;
    switch (h->ctrl_mode) {
        case CTRL_NONE: return h->fail(OCS_B200_ERR_STATE, "no control set (ocs_b200_set_control_v / _u)");
        case CTRL_V_SHARED:
        case CTRL_V_BATCH:
            if (h->basis_kind == BASIS_NONE)
                return h->fail(OCS_B200_ERR_STATE, "control coefficients v need a basis (ocs_b200_set_control_basis)");
            if (h->ctrl_mode == CTRL_V_BATCH && h->ctrl_nbatch != Bt)
                return h->fail(OCS_B200_ERR_ARG, "v has %ld columns but the batch has %ld trajectories", h->ctrl_nbatch, Bt);
            if (h->ctrl_mode == CTRL_V_SHARED && (long)h->ctrl_shared.size() != (long)nc * h->K)
                return h->fail(OCS_B200_ERR_ARG, "v must have nc*K = %d entries", nc * h->K);
            if (h->basis_kind == BASIS_DENSE && h->kt.kd == 0 && h->K > OCS_KDENSE_MAX && h->ctrl_mode == CTRL_V_BATCH)
                return h->fail(OCS_B200_ERR_ARG, "dense basis with K = %d > %d needs a generated problem", h->K, OCS_KDENSE_MAX);
            return 0;
        case CTRL_U_BATCH:
            if (h->ctrl_nbatch != Bt)
                return h->fail(OCS_B200_ERR_ARG, "u has %ld pages but the batch has %ld trajectories", h->ctrl_nbatch, Bt);
            return 0;
        default: return 0;
    }
}

// Control and parameters fit a batch of Bt trajectories.  Checked BEFORE anything is (re)allocated for a new batch and
// again by every call that sweeps the resident one (a setter may have changed them in between: the kernels index
// d_v / d_ug / d_params with the batch stride Bt).
int validate_batch(H *h, long Bt)
{
    TRY(validate_control(h, Bt));
    if (h->params_nbatch > 1 && h->params_nbatch != Bt)
        return h->fail(OCS_B200_ERR_ARG, "params have %ld columns but the batch has %ld trajectories", h->params_nbatch, Bt);
    return 0;
}

int states_common(H *h, long Bt)
{
    h->Bt = Bt;
    h->fb_resident = false;
    TRY(run_forward(h, 0, Bt, -1, nullptr, true));
    h->have_states = true;
    return 0;
}

// The sparse-basis sweep stores every dJdv column exactly once when the knot window is monotone and reaches
// every knot; otherwise (dense bases, exotic sparse ones, frozen trajectories) the buffer starts from zero.
bool needs_zero_fill(const ocs_b200_handle *h)
{
    return !(h->basis_kind == BASIS_SPARSE && h->bk_monotone && h->bk_covered && !h->fb_frozen);
}

// ------------------------------------------------------------------------------------------------ two lanes
// A sweep kernel whose threads each carry a several-state trajectory keeps 8-12 warps resident per SM, and a batch of
// 10^5 trajectories is then only two or three rounds of blocks: the last, partly filled round leaves SMs idle for
// 12-16 % of the kernel (ncu: sm__cycles_active.avg against sm__cycles_elapsed.max, profiles/r2_notes.md).  Such a batch
// is swept as two halves on two streams.  The halves are independent (trajectories are), every kernel of one half
// follows the previous kernel of the SAME half only, and the block scheduler fills the SMs that one half's last round
// leaves free with the other half's next kernel.  Same-box measurement on cfg3 (100k candidates of the 6-state
// problem): 6.74 -> 5.9 ms per evaluation.

// resident warps per SM and rounds of blocks of a sweep kernel launched over the whole batch
void lane_geometry(H *h, const void *kernel, long Bt, int tpt, size_t smem, int *resident_warps, double *rounds)
{
    int nblk = 0;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, kernel, h->kt.threads, smem) != cudaSuccess || nblk < 1) {
        cudaGetLastError();
        nblk = 1;
    }
    if (h->sm_count <= 0) {
        cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device);
        if (h->sm_count <= 0) h->sm_count = 148;
    }
    const double blocks = (double)((Bt + (long)h->kt.threads * tpt - 1) / ((long)h->kt.threads * tpt));
    *resident_warps = nblk * h->kt.threads / 32;
    *rounds = blocks / ((double)h->sm_count * nblk);
}

// Whether a call whose widest kernel is `kernel` sweeps the batch as two lanes.  Automatic: at most 12 resident warps
// per SM (above that a partly filled round runs proportionally faster: the FP64 pipe is shared by fewer warps, and
// there is nothing to gain) and fewer than 8 rounds (the last round's share of the kernel).  Never under per-kernel
// profiling, which times each kernel alone over the whole batch.
bool use_lanes(H *h, const void *kernel, long Bt, int tpt, size_t smem)
{
    if (h->profiling || h->lanes_mode == 0) return false;
    const long gran = (long)h->kt.threads * 12;
    if (Bt < 2 * gran) return false;
    if (h->lanes_mode == 1) return true;
    int warps = 0;
    double rounds = 0;
    lane_geometry(h, kernel, Bt, tpt, smem, &warps, &rounds);
    return warps <= 12 && rounds < 8.0;
}

// the two halves: whole blocks of every kernel variant (1, 2, 3 or 4 trajectories per thread)
void lane_split(const H *h, long Bt, long b0[2], long nb[2])
{
    const long gran = (long)h->kt.threads * 12;
    const long half = ((Bt + 1) / 2 + gran - 1) / gran * gran;
    b0[0] = 0;
    nb[0] = std::min(half, Bt);
    b0[1] = nb[0];
    nb[1] = Bt - nb[0];
}

// everything queued on the handle's stream so far happens before the lanes start
int lanes_fork(H *h)
{
    for (int l = 0; l < 2; ++l) {
        if (!h->lane_stream[l]) CK(cudaStreamCreateWithFlags(&h->lane_stream[l], cudaStreamNonBlocking));
        if (!h->lane_done[l]) CK(cudaEventCreateWithFlags(&h->lane_done[l], cudaEventDisableTiming));
    }
    if (!h->lane_fork) CK(cudaEventCreateWithFlags(&h->lane_fork, cudaEventDisableTiming));
    CK(cudaEventRecord(h->lane_fork, h->stream));
    for (int l = 0; l < 2; ++l) CK(cudaStreamWaitEvent(h->lane_stream[l], h->lane_fork, 0));
    return 0;
}

// ... and everything queued on the lanes happens before what is queued on the handle's stream from now on
int lanes_join(H *h)
{
    for (int l = 0; l < 2; ++l) {
        CK(cudaEventRecord(h->lane_done[l], h->lane_stream[l]));
        CK(cudaStreamWaitEvent(h->stream, h->lane_done[l], 0));
    }
    return 0;
}

// runs body(b0, nb) for each half with the handle's stream pointing at the half's lane
template <class F>
int on_lanes(H *h, long Bt, F body)
{
    long b0[2], nb[2];
    lane_split(h, Bt, b0, nb);
    TRY(lanes_fork(h));
    cudaStream_t main_stream = h->stream;
    int rc = 0;
    for (int l = 0; l < 2 && rc == 0; ++l) {
        if (nb[l] <= 0) continue;
        h->stream = h->lane_stream[l];
        rc = body(b0[l], nb[l]);
    }
    h->stream = main_stream;
    if (rc != 0) {              // leave nothing running behind the caller's back
        for (int l = 0; l < 2; ++l) cudaStreamSynchronize(h->lane_stream[l]);
        return rc;
    }
    return lanes_join(h);
}

// The kernel and launch shape that decide whether a gradient evaluation uses lanes: the forward sweep of the main horizon.
bool evaluation_uses_lanes(H *h, long Bt)
{
    if (h->profiling || h->lanes_mode == 0 || h->has_tail) return false;
    if (build_records(h, 0) != 0) return false;
    int um = main_umode(h), bm = um == 1 ? h->basis_kind : 0;
    const void *kern = nullptr;
    if (resolve(h, KIND_FWD, um, bm, pb_index(h), &kern) != 0) return false;
    SweepArgs a;
    base_args(h, 0, a);
    const int tpt = bm == 2 ? 1 : fwd_tpt(h, um);
    return use_lanes(h, kern, Bt, tpt, smem_bytes(h, a, 0));
}

// Forward sweep and (optionally) adjoint sweep of the whole resident batch, as one pass or as two lanes.  Everything
// the halves share -- records and tables, buffers the sweeps would otherwise allocate on first use, the zero-fill of
// dJdv -- is prepared on the handle's stream before the lanes fork.
int sweeps_common(H *h, long Bt, bool forward, bool reverse, const double *d_lamT, long ldl, double *d_dJdv, double *d_lam0)
{
    if (forward) {
        h->Bt = Bt;
        h->fb_resident = false;
    }
    const bool lanes = evaluation_uses_lanes(h, Bt);
    if (lanes) {
        const int nc = h->kt.nc;
        const Segment &s = h->seg[0];
        const int um = main_umode(h);
        if ((um == 1 && h->basis_kind == BASIS_DENSE) || um == 2) TRY(ensure(h, h->d_ug, (size_t)nc * (2 * s.N + 1) * Bt));
        if (reverse && d_dJdv && h->basis_kind == BASIS_DENSE) TRY(ensure(h, h->d_dug, (size_t)nc * (2 * s.N + 1) * Bt));
    }
    if (reverse && d_dJdv && needs_zero_fill(h))
        CK(cudaMemsetAsync(d_dJdv, 0, sizeof(double) * h->kt.nc * h->K * Bt, h->stream));
    if (!lanes) {
        if (forward) TRY(run_forward(h, 0, Bt, -1, nullptr, true));
        if (forward) h->have_states = true;
        if (reverse) TRY(run_reverse(h, 0, Bt, -1, d_lamT, ldl, d_dJdv, d_lam0, nullptr, nullptr));
        return 0;
    }
    TRY(on_lanes(h, Bt, [&](long b0, long nb) -> int {
        if (forward) TRY(run_forward(h, b0, nb, -1, nullptr, true));
        if (reverse) TRY(run_reverse(h, b0, nb, -1, d_lamT, ldl, d_dJdv ? d_dJdv + b0 : nullptr, d_lam0 ? d_lam0 + b0 : nullptr,
                                     nullptr, nullptr));
        return 0;
    }));
    if (forward) h->have_states = true;
    return 0;
}

// Calls that use the resident sweep: it exists and still matches the control / parameters that are set now.
int check_resident(H *h, const char *who)
{
    if (!h->have_states) return h->fail(OCS_B200_ERR_STATE, "%s needs a preceding compute_states on this handle", who);
    return validate_batch(h, h->Bt);
}

}  // namespace

// =============================================================================== C ABI

extern "C" {

int ocs_b200_abi_version(void) { return OCS_B200_ABI_VERSION; }

const char *ocs_b200_last_error(const ocs_b200_handle *h)
{
    if (!h) return g_create_error.c_str();
    return h->err.c_str();
}

static int check_tspan(const double *ts, int N)
{
    for (int i = 0; i <= N; ++i)
        if (!std::isfinite(ts[i])) return 1;
    return 0;
}

int ocs_b200_create(int N, const double *tspan, int N2, const double *tspan_extra, int device, int math_mode,
                    ocs_b200_handle **out)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!out) {
        g_create_error = "out is NULL";
        return OCS_B200_ERR_ARG;
    }
    *out = nullptr;
    if (N < 1 || !tspan || (N2 > 0 && !tspan_extra) || N2 < 0 || check_tspan(tspan, N) ||
        (N2 > 0 && check_tspan(tspan_extra, N2))) {
        g_create_error = "tspan must hold at least two finite times (and tspan_extra when N2 > 0)";
        return OCS_B200_ERR_ARG;
    }
    if (math_mode != OCS_B200_MATH_FAST && math_mode != OCS_B200_MATH_STRICT) {
        g_create_error = "unknown math_mode";
        return OCS_B200_ERR_ARG;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        g_create_error = std::string("no CUDA device available (there is no CPU fallback): ") +
                         (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return OCS_B200_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        g_create_error = "device ordinal out of range";
        return OCS_B200_ERR_ARG;
    }
    ocs_b200_handle *h = new (std::nothrow) ocs_b200_handle();
    if (!h) {
        g_create_error = "out of host memory";
        return OCS_B200_ERR_NOMEM;
    }
    h->device = device;
    h->math_mode = math_mode;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    for (int k = 0; k < 4 && e == cudaSuccess; ++k) e = cudaEventCreate(&h->ev[k]);
    // the kernels' "wrong record layout" flag lives in mapped pinned host memory: raised by the device (only on an
    // error), read by the host at every synchronisation point without a copy
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&h->h_errflag, sizeof(int), cudaHostAllocMapped);
    if (e == cudaSuccess) {
        *h->h_errflag = 0;
        e = cudaHostGetDevicePointer((void **)&h->d_errflag, h->h_errflag, 0);
    }
    if (e != cudaSuccess) {
        g_create_error = std::string("CUDA initialisation failed: ") + cudaGetErrorString(e);
        delete h;
        return OCS_B200_ERR_CUDA;
    }
    h->stream = h->own_stream;
    if (const char *lm = getenv("OCS_B200_LANES")) h->lanes_mode = atoi(lm) > 0 ? 1 : (lm[0] == '0' ? 0 : -1);
    for (int si = 0; si < 2; ++si) {
        const int Ns = si == 0 ? N : N2;
        const double *ts = si == 0 ? tspan : tspan_extra;
        if (Ns == 0) continue;
        Segment &s = h->seg[si];
        s.N = Ns;
        s.tspan.assign(ts, ts + Ns + 1);
        s.h.resize(Ns);
        s.t.resize(2 * Ns + 1);
        for (int i = 0; i < Ns; ++i) s.h[i] = ts[i + 1] - ts[i];                      // diff(tspan)          :17
        for (int i = 0; i <= Ns; ++i) s.t[2 * i] = ts[i];                             // t(1:2:end) = tspan    :22
        for (int i = 0; i < Ns; ++i) s.t[2 * i + 1] = (ts[i] + ts[i + 1]) / 2;        // midpoints             :23
    }
    h->has_tail = N2 > 0;
    *out = h;
    return 0;
}

int ocs_b200_destroy(ocs_b200_handle *h)
{
    if (!h) return 0;
    if (h->multi) return ocs_multi_destroy(h);
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int si = 0; si < 2; ++si) {
        cudaFree(h->seg[si].d_gprec);
        cudaFree(h->seg[si].d_steprec);
        cudaFree(h->seg[si].d_foldrec);
        cudaFree(h->seg[si].d_basrec);
        if (h->seg[si].h_rec) cudaFreeHost(h->seg[si].h_rec);
        if (h->seg[si].rec_uploaded) cudaEventDestroy(h->seg[si].rec_uploaded);
        cudaFree(h->seg[si].d_xs_base);
    }
    DevBuf *bufs[] = {&h->d_params, &h->d_v,    &h->d_ug,    &h->d_x0,     &h->d_J,     &h->d_xT,   &h->d_lamT_tail,
                      &h->d_lamT_user, &h->d_dJdv, &h->d_lam0, &h->d_stage, &h->d_stage2, &h->d_maxchg,
                      &h->d_pipe[0], &h->d_pipe[1], &h->d_z, &h->d_g, &h->d_zt, &h->d_gt, &h->d_Jacc, &h->d_Jhist,
                      &h->d_alpha, &h->d_lbub, &h->d_dug, &h->d_pred, &h->d_cgb_mu, &h->d_cgb_b, &h->d_cgb_viol};
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (int k = 0; k < 2; ++k) {
        if (h->lane_stream[k]) cudaStreamDestroy(h->lane_stream[k]);
        if (h->lane_done[k]) cudaEventDestroy(h->lane_done[k]);
        for (int q = 0; q < 2; ++q)
            if (h->lane_sweep[k][q]) cudaEventDestroy(h->lane_sweep[k][q]);
    }
    if (h->lane_fork) cudaEventDestroy(h->lane_fork);
    cudaFree(h->d_lane_open);
    if (h->h_lane_open) cudaFreeHost(h->h_lane_open);
    for (int k = 0; k < 2; ++k) {
        if (h->chunk_done[k]) cudaEventDestroy(h->chunk_done[k]);
        if (h->stage_free[k]) cudaEventDestroy(h->stage_free[k]);
    }
    for (DevBuf *b : bufs) cudaFree(b->p);
    cudaFree(h->d_status);
    cudaFree(h->d_frozen);
    cudaFree(h->d_sweeps);
    cudaFree(h->d_bounds);
    cudaFree(h->d_red);
    if (h->h_errflag) cudaFreeHost(h->h_errflag);
    h->stager.shutdown();
    for (int k = 0; k < 4; ++k)
        if (h->ev[k]) cudaEventDestroy(h->ev[k]);
    for (int k = 0; k < 2; ++k)
        if (h->ev_loop[k]) cudaEventDestroy(h->ev_loop[k]);
    for (void *lib : h->gen_libraries) ocs_unload_problem(lib);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    cudaGetLastError();
    delete h;
    return 0;
}

int ocs_b200_get_grid(const ocs_b200_handle *h, int segment, double *t, double *hstep)
{
    if (h && h->multi) return ocs_multi_get_grid(h, segment, t, hstep);
    if (!h || segment < 0 || segment > 1 || h->seg[segment].N == 0) return OCS_B200_ERR_ARG;
    const Segment &s = h->seg[segment];
    if (t) memcpy(t, s.t.data(), sizeof(double) * s.t.size());
    if (hstep) memcpy(hstep, s.h.data(), sizeof(double) * s.h.size());
    return 0;
}

int ocs_b200_set_problem_builtin(ocs_b200_handle *h, int problem_id)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_set_problem_builtin(h, problem_id);
    if (problem_id != OCS_B200_PROBLEM_TEST) return h->fail(OCS_B200_ERR_ARG, "unknown built-in problem %d", problem_id);
    if (h->math_mode == OCS_B200_MATH_STRICT)
        ocs_builtin_table_test_strict(&h->kt);
    else if (steps_uniform(h))
        ocs_builtin_table_test_fast(&h->kt);       // step sizes as constant-bank operands
    else
        ocs_builtin_table_test_fastv(&h->kt);      // per-step sizes from the step records
    h->have_problem = true;
    h->is_generated = false;
    h->have_states = false;
    mark_dirty(h);
    return 0;
}

int ocs_b200_set_problem_source(ocs_b200_handle *h, const char *cuda_source, int n, int nc, int np,
                                const char *kernel_include_dir)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_set_problem_source(h, cuda_source, n, nc, np, kernel_include_dir);
    if (!cuda_source || n < 1 || nc < 1 || np < 0) return h->fail(OCS_B200_ERR_ARG, "bad generated-problem description");
    try {
        CK(cudaSetDevice(h->device));
        std::string log;
        ocs_kernel_table kt{};
        void *lib = nullptr;
        int rc = ocs_compile_problem(cuda_source, n, nc, np, nvrtc_flags(h),
                                     (kernel_include_dir && kernel_include_dir[0]) ? kernel_include_dir : nullptr, &kt, &lib, &log);
        // a failed compilation leaves the handle as it was (problem, lazily compiled variants and their source)
        if (rc != 0) return h->fail(rc, "generated problem did not compile:\n%s", log.c_str());
        h->gen_source = cuda_source;
        h->gen_include = kernel_include_dir ? kernel_include_dir : "";
        h->gen_n = n;
        h->gen_nc = nc;
        h->gen_np = np;
        for (void *old : h->gen_libraries) ocs_unload_problem(old);
        h->gen_libraries.clear();
        h->gen_libraries.push_back(lib);
        h->kt = kt;
        h->have_problem = true;
        h->is_generated = true;
        h->have_states = false;
        mark_dirty(h);
        return 0;
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

int ocs_b200_set_params(ocs_b200_handle *h, const double *params, int np, long nbatch)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_set_params(h, params, np, nbatch);
    if (!h->have_problem) return h->fail(OCS_B200_ERR_STATE, "bind a problem before setting parameters");
    if (np != h->kt.np || nbatch < 1 || (!params && np > 0))
        return h->fail(OCS_B200_ERR_ARG, "the problem takes %d parameters (got %d, nbatch %ld)", h->kt.np, np, nbatch);
    try {
        CK(cudaSetDevice(h->device));
        if (nbatch == 1) {
            h->params_shared.assign(params, params + np);
            TRY(ensure(h, h->d_params, np));
            TRY(h2d(h, h->d_params.p, params, sizeof(double) * np));
            CK(cudaStreamSynchronize(h->stream));
        } else {
            TRY(upload_soa(h, params, h->d_params, np, nbatch));
            CK(cudaStreamSynchronize(h->stream));
        }
        h->params_nbatch = nbatch;
        mark_dirty(h);
        return 0;
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

int ocs_b200_set_tail_control(ocs_b200_handle *h, const double *uStar, int nc)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_set_tail_control(h, uStar, nc);
    if (!h->have_problem) return h->fail(OCS_B200_ERR_STATE, "bind a problem before setting uStar");
    if (!uStar || nc != h->kt.nc) return h->fail(OCS_B200_ERR_ARG, "uStar must have nc = %d entries", h->kt.nc);
    h->uStar.assign(uStar, uStar + nc);
    h->seg[1].records_dirty = true;
    return 0;
}

int ocs_b200_set_control_basis(ocs_b200_handle *h, int K, const double *B)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_set_control_basis(h, K, B);
    try {
        if (K == 0) {
            h->basis_kind = BASIS_NONE;
            h->K = 0;
            if (h->ctrl_mode == CTRL_V_SHARED || h->ctrl_mode == CTRL_V_BATCH) {
                h->ctrl_mode = CTRL_NONE;     // coefficients without a basis mean nothing
                h->have_states = false;
            }
            mark_dirty(h);
            return 0;
        }
        if (K < 0 || !B) return h->fail(OCS_B200_ERR_ARG, "bad basis");
        const int G = 2 * h->seg[0].N + 1;
        std::vector<int32_t> bk(G), nnz(G);
        std::vector<double> w0(G), w1(G);
        bool sparse = true;
        for (int j = 0; j < G; ++j) {
            const double *col = B + (size_t)K * j;
            int first = -1, last = -1, cnt = 0;
            for (int k = 0; k < K; ++k) {
                if (!std::isfinite(col[k])) return h->fail(OCS_B200_ERR_ARG, "basis matrix holds a non-finite entry");
                if (col[k] != 0.0) {
                    if (first < 0) first = k;
                    last = k;
                    ++cnt;
                }
            }
            nnz[j] = cnt;
            if (cnt == 0) {   // an all-zero column: u_j = 0
                bk[j] = 0;
                w0[j] = 0.0;
                w1[j] = 0.0;
            } else if (cnt <= 2 && last - first <= 1) {
                bk[j] = first;
                w0[j] = col[first];
                w1[j] = last > first ? col[last] : 0.0;
            } else {
                sparse = false;
            }
        }
        h->K = K;
        h->bnnz = nnz;
        if (sparse) {
            h->basis_kind = BASIS_SPARSE;
            h->bk_monotone = true;      // the sliding knot window never returns to a knot it has left
            for (int j = 1; j < G; ++j)
                if (bk[j] < bk[j - 1]) h->bk_monotone = false;
            {
                std::vector<char> seen((size_t)K + 1, 0);
                for (int j = 0; j < G; ++j) seen[bk[j]] = seen[bk[j] + 1] = 1;
                h->bk_covered = h->bk_monotone;
                for (int k = 0; k < K; ++k)
                    if (!seen[k]) h->bk_covered = false;
            }
            h->bk = bk;
            h->bw0 = w0;
            h->bw1 = w1;
            h->bdense.clear();
        } else {
            h->basis_kind = BASIS_DENSE;
            h->bdense.assign(B, B + (size_t)K * G);
            h->bk.assign(G, 0);
            for (int j = 0; j < G; ++j) h->bnnz[j] = K;
        }
        // a generated problem's dense-basis kernels are specialised on K: forget them when K changes
        if (h->is_generated && h->basis_kind == BASIS_DENSE && h->kt.kd != K) {
            for (int um = 0; um < 3; ++um)
                for (int pb = 0; pb < 2; ++pb)
                    h->kt.fwd[um][2][pb] = h->kt.rev[um][2][pb] = h->kt.fwdg[um][2][pb] = h->kt.revg[um][2][pb] = nullptr;
            h->kt.kd = K;
        }
        h->ctrl_mode = CTRL_NONE;   // coefficients belong to the previous basis
        h->have_states = false;     // ... and the resident sweep to the previous control
        h->ufunc_kind = 0;          // ... and so does its uFunc
        h->ufunc_knots.clear();
        mark_dirty(h);
        return 0;
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

int ocs_b200_get_control_index(const ocs_b200_handle *h, int32_t *k_index, int32_t *nnz)
{
    if (h && h->multi) return ocs_multi_get_control_index(h, k_index, nnz);
    if (!h || h->basis_kind == BASIS_NONE) return OCS_B200_ERR_STATE;
    const int G = 2 * h->seg[0].N + 1;
    if (k_index) memcpy(k_index, h->bk.data(), sizeof(int32_t) * G);
    if (nnz) memcpy(nnz, h->bnnz.data(), sizeof(int32_t) * G);
    return 0;
}

int ocs_b200_set_control_v(ocs_b200_handle *h, const double *v, long nbatch)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_set_control_v(h, v, nbatch);
    if (!h->have_problem) return h->fail(OCS_B200_ERR_STATE, "bind a problem before setting the control");
    if (h->basis_kind == BASIS_NONE) return h->fail(OCS_B200_ERR_STATE, "attach a control basis before setting v");
    if (!v || nbatch < 1) return h->fail(OCS_B200_ERR_ARG, "bad v");
    try {
        CK(cudaSetDevice(h->device));
        const long rows = (long)h->kt.nc * h->K;
        h->v_ext = nullptr;
        if (nbatch == 1) {
            h->ctrl_shared.assign(v, v + rows);
            h->ctrl_mode = CTRL_V_SHARED;
        } else {
            TRY(upload_soa(h, v, h->d_v, rows, nbatch));
            CK(cudaStreamSynchronize(h->stream));
            h->ctrl_mode = CTRL_V_BATCH;
        }
        h->ctrl_nbatch = nbatch;
        h->seg[0].records_dirty = true;
        return 0;
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

int ocs_b200_set_control_v_dev(ocs_b200_handle *h, const double *v_soa, long nbatch)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi)
        return h->fail(OCS_B200_ERR_ARG, "ocs_b200_set_control_v_dev takes pointers / a stream of ONE device; on a multi-device handle use "
                                          "ocs_b200_evaluate_resident, ocs_b200_gather_results_dev and ocs_b200_get_results");
    if (h->basis_kind == BASIS_NONE) return h->fail(OCS_B200_ERR_STATE, "attach a control basis before setting v");
    if (!v_soa || nbatch < 2) return h->fail(OCS_B200_ERR_ARG, "device v needs one column per trajectory");
    h->v_ext = v_soa;
    h->ctrl_mode = CTRL_V_BATCH;
    h->ctrl_nbatch = nbatch;
    h->seg[0].records_dirty = true;
    return 0;
}

int ocs_b200_set_control_u(ocs_b200_handle *h, const double *u, long nbatch)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_set_control_u(h, u, nbatch);
    if (!h->have_problem) return h->fail(OCS_B200_ERR_STATE, "bind a problem before setting the control");
    if (!u || nbatch < 1) return h->fail(OCS_B200_ERR_ARG, "bad u");
    try {
        CK(cudaSetDevice(h->device));
        const int nc = h->kt.nc, G = 2 * h->seg[0].N + 1;
        h->v_ext = nullptr;
        if (nbatch == 1) {
            h->ctrl_shared.assign(u, u + (size_t)nc * G);
            h->ctrl_mode = CTRL_U_SHARED;
        } else {
            TRY(upload_soa(h, u, h->d_ug, (long)nc * G, nbatch, nc));
            CK(cudaStreamSynchronize(h->stream));
            h->ctrl_mode = CTRL_U_BATCH;
        }
        h->ctrl_nbatch = nbatch;
        h->seg[0].records_dirty = true;
        return 0;
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

int ocs_b200_get_control_u(ocs_b200_handle *h, long b0, long nb, double *u)
{
    if (!h || !u) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_get_control_u(h, b0, nb, u);
    if (!h->have_problem || h->ctrl_mode == CTRL_NONE) return h->fail(OCS_B200_ERR_STATE, "no control set");
    try {
        CK(cudaSetDevice(h->device));
        const int nc = h->kt.nc, G = 2 * h->seg[0].N + 1;
        if (h->ctrl_mode == CTRL_U_SHARED) {
            memcpy(u, h->ctrl_shared.data(), sizeof(double) * nc * G);
            return 0;
        }
        if (h->ctrl_mode == CTRL_V_SHARED) {
            for (int j = 0; j < G; ++j) shared_u_from_v(h, j, u + (size_t)nc * j);
            return 0;
        }
        if (h->ctrl_mode == CTRL_U_BATCH) {
            if (b0 < 0 || nb < 1 || b0 + nb > h->ctrl_nbatch) return h->fail(OCS_B200_ERR_ARG, "trajectory range out of bounds");
            TRY(download_soa(h, h->d_ug.p, h->ctrl_nbatch, b0, u, (long)nc * G, nb, nc));
            return finish_host(h);
        }
        // per-trajectory coefficients: evaluate v*B on the host for the requested columns
        if (b0 < 0 || nb < 1 || b0 + nb > h->ctrl_nbatch) return h->fail(OCS_B200_ERR_ARG, "trajectory range out of bounds");
        const long rows = (long)nc * h->K;
        std::vector<double> vcols((size_t)rows * nb);
        TRY(download_soa(h, h->v_ext ? h->v_ext : h->d_v.p, h->ctrl_nbatch, b0, vcols.data(), rows, nb));
        TRY(finish_host(h));
        std::vector<double> keep = h->ctrl_shared;
        for (long b = 0; b < nb; ++b) {
            h->ctrl_shared.assign(vcols.begin() + b * rows, vcols.begin() + (b + 1) * rows);
            for (int j = 0; j < G; ++j) shared_u_from_v(h, j, u + (size_t)nc * (j + (size_t)G * b));
        }
        h->ctrl_shared = keep;
        return 0;
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

int ocs_b200_set_stream(ocs_b200_handle *h, void *cuda_stream)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi)
        return h->fail(OCS_B200_ERR_ARG, "ocs_b200_set_stream takes pointers / a stream of ONE device; on a multi-device handle use "
                                          "ocs_b200_evaluate_resident, ocs_b200_gather_results_dev and ocs_b200_get_results");
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return 0;
}

int ocs_b200_synchronize(ocs_b200_handle *h)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_synchronize(h);
    CK(cudaSetDevice(h->device));
    return finish_host(h);
}

int ocs_b200_compute_states_dev(ocs_b200_handle *h, const double *x0_soa, long Bt, double *J, double *xT_soa,
                                int32_t *status)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi)
        return h->fail(OCS_B200_ERR_ARG, "ocs_b200_compute_states_dev takes pointers / a stream of ONE device; on a multi-device handle use "
                                          "ocs_b200_evaluate_resident, ocs_b200_gather_results_dev and ocs_b200_get_results");
    TRY(check_ready(h));
    if (!x0_soa || Bt < 1) return h->fail(OCS_B200_ERR_ARG, "bad x0 / batch size");
    try {
        CK(cudaSetDevice(h->device));
        const int n = h->kt.n;
        h->have_states = false;         // a failure below must not leave a half-replaced batch looking valid
        TRY(validate_batch(h, Bt));
        TRY(alloc_batch(h, Bt));
        CK(cudaMemcpyAsync(h->d_x0.p, x0_soa, sizeof(double) * n * Bt, cudaMemcpyDeviceToDevice, h->stream));
        TRY(states_common(h, Bt));
        if (J) CK(cudaMemcpyAsync(J, h->d_J.p, sizeof(double) * Bt, cudaMemcpyDeviceToDevice, h->stream));
        if (xT_soa) CK(cudaMemcpyAsync(xT_soa, h->d_xT.p, sizeof(double) * n * Bt, cudaMemcpyDeviceToDevice, h->stream));
        if (status) CK(cudaMemcpyAsync(status, h->d_status, sizeof(int) * Bt, cudaMemcpyDeviceToDevice, h->stream));
        return 0;
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

// nlpObjective (functions/single_shooting.m:137-150: compute_states followed by compute_adjoints) on caller-owned
// DEVICE buffers in ONE call, asynchronous on the handle's stream: both sweeps are queued at once, as two lanes when
// that pays (sweeps_common).
int ocs_b200_evaluate_dev(ocs_b200_handle *h, const double *x0_soa, long Bt, const double *lamT_soa, long lamT_nbatch,
                          double *J, double *xT_soa, int32_t *status, double *dJdv_soa, double *lam0_soa)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi)
        return h->fail(OCS_B200_ERR_ARG, "ocs_b200_evaluate_dev takes pointers / a stream of ONE device; on a multi-device handle use "
                                          "ocs_b200_evaluate_resident, ocs_b200_gather_results_dev and ocs_b200_get_results");
    TRY(check_ready(h));
    if (!x0_soa || Bt < 1) return h->fail(OCS_B200_ERR_ARG, "bad x0 / batch size");
    if (lamT_soa && lamT_nbatch != 1 && lamT_nbatch != Bt) return h->fail(OCS_B200_ERR_ARG, "lamT must have 1 or Bt columns");
    try {
        CK(cudaSetDevice(h->device));
        const int n = h->kt.n;
        h->have_states = false;
        TRY(validate_batch(h, Bt));
        if (dJdv_soa && h->basis_kind == BASIS_NONE) return h->fail(OCS_B200_ERR_STATE, "dJdv requested but no control basis is attached");
        TRY(alloc_batch(h, Bt));
        CK(cudaMemcpyAsync(h->d_x0.p, x0_soa, sizeof(double) * n * Bt, cudaMemcpyDeviceToDevice, h->stream));
        h->user_lamT = false;
        TRY(sweeps_common(h, Bt, true, dJdv_soa || lam0_soa, lamT_soa, lamT_soa && lamT_nbatch > 1 ? Bt : 0, dJdv_soa, lam0_soa));
        if (J) CK(cudaMemcpyAsync(J, h->d_J.p, sizeof(double) * Bt, cudaMemcpyDeviceToDevice, h->stream));
        if (xT_soa) CK(cudaMemcpyAsync(xT_soa, h->d_xT.p, sizeof(double) * n * Bt, cudaMemcpyDeviceToDevice, h->stream));
        if (status) CK(cudaMemcpyAsync(status, h->d_status, sizeof(int) * Bt, cudaMemcpyDeviceToDevice, h->stream));
        return 0;
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

int ocs_b200_compute_states(ocs_b200_handle *h, const double *x0, long Bt, double *J, double *xT, int32_t *status)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_compute_states(h, x0, Bt, J, xT, status);
    TRY(check_ready(h));
    if (!x0 || Bt < 1) return h->fail(OCS_B200_ERR_ARG, "bad x0 / batch size");
    try {
        CK(cudaSetDevice(h->device));
        const int n = h->kt.n;
        h->have_states = false;         // a failure below must not leave a half-replaced batch looking valid
        TRY(validate_batch(h, Bt));
        TRY(alloc_batch(h, Bt));
        TRY(upload_soa(h, x0, h->d_x0, n, Bt));
        TRY(states_common(h, Bt));
        if (J) TRY(d2h(h, J, h->d_J.p, sizeof(double) * Bt));
        if (xT) TRY(download_soa(h, h->d_xT.p, Bt, 0, xT, n, Bt));
        if (status) TRY(d2h(h, status, h->d_status, sizeof(int) * Bt));
        return finish_host(h);
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

// Trajectory chunks of a host-buffer evaluation whose gradient download (rows*Bt doubles over PCIe) overlaps the
// sweeps.  Short horizons (the download dominates): the first chunk is small so that the transfer starts almost at
// once, later chunks double (the sweeps run about three times faster than PCIe drains them) up to a quarter of the
// batch.  Long horizons (the sweeps dominate: a block takes N steps whatever the chunk size, so a chunk below one
// wave of resident blocks costs a whole wave's time -- cfg4, 20 000 steps, took 60 ms in 7 geometric chunks against
// 27 ms of sweeps): equal chunks of at least one wave each, or a single chunk.
static std::vector<std::pair<long, long>> chunk_schedule(const ocs_b200_handle *h, long Bt)
{
    std::vector<std::pair<long, long>> out;
    const long gran = 128 * 12;                      // whole blocks of every kernel variant (1, 2, 3 or 4 trajectories per thread)
    const long ntot = h->seg[0].N + (h->has_tail ? h->seg[1].N : 0);
    if (ntot > 2500) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        const long wave = (long)sms * 1280;          // trajectories resident at once in the widest sweep kernel
        long nch = std::min<long>(8, Bt / wave);
        if (nch < 2) {
            out.emplace_back(0, Bt);
            return out;
        }
        const long c = ((Bt + nch - 1) / nch + gran - 1) / gran * gran;
        for (long b0 = 0; b0 < Bt; b0 += c) out.emplace_back(b0, std::min(c, Bt - b0));
        return out;
    }
    long c = std::max<long>(gran, (Bt / 32 + gran - 1) / gran * gran);
    const long cmax = std::max<long>(gran, (Bt / 4 + gran - 1) / gran * gran);
    for (long b0 = 0; b0 < Bt;) {
        const long nb = std::min(c, Bt - b0);
        out.emplace_back(b0, nb);
        b0 += nb;
        c = std::min(2 * c, cmax);
    }
    return out;
}

// Sweeps trajectories chunk by chunk on the compute stream (forward too when `with_forward`) and transposes the
// chunk's gradient into one of two staging buffers, while a second stream downloads the previous chunk.
// OCS_B200_TRACE=1 prints the timeline (sweeps done / transposed / downloaded, per chunk) to stderr.
static int pipelined_gradient(ocs_b200_handle *h, bool with_forward, const double *d_lamT, long ldl, double *d_dJdv,
                              double *d_lam0, double *dJdv, double *lam0)
{
    const int m = h->kt.n + 1, nc = h->kt.nc;
    const long Bt = h->Bt, rows = (long)nc * h->K;
    if (!h->copy_stream) CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
        if (!h->chunk_done[k]) CK(cudaEventCreateWithFlags(&h->chunk_done[k], cudaEventDisableTiming));
        if (!h->stage_free[k]) CK(cudaEventCreateWithFlags(&h->stage_free[k], cudaEventDisableTiming));
    }
    const auto sched = chunk_schedule(h, Bt);
    long cmax = 0;
    for (auto &c : sched) cmax = std::max(cmax, c.second);
    for (int k = 0; k < 2; ++k) TRY(ensure(h, h->d_pipe[k], (size_t)(rows + m) * cmax));
    if (needs_zero_fill(h)) CK(cudaMemsetAsync(d_dJdv, 0, sizeof(double) * rows * Bt, h->stream));
    const bool trace = getenv("OCS_B200_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t st) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); } };
    mark(h->stream);
    // Software pipeline: the sweeps and the transposition of chunk c are launched first, THEN the download of chunk
    // c-1 is issued.  Issuing a download may block this thread (a pageable destination drains through the pinned
    // ring at host-memcpy speed); the device is busy with chunk c meanwhile.
    auto download = [&](int c) -> int {
        const long b0 = sched[c].first, nb = sched[c].second;
        double *stg = h->d_pipe[c & 1].p;
        CK(cudaStreamWaitEvent(h->copy_stream, h->chunk_done[c & 1], 0));
        TRY(d2h(h, dJdv + (size_t)rows * b0, stg, sizeof(double) * rows * nb, h->copy_stream));
        if (lam0) TRY(d2h(h, lam0 + (size_t)m * b0, stg + (size_t)rows * cmax, sizeof(double) * m * nb, h->copy_stream));
        CK(cudaEventRecord(h->stage_free[c & 1], h->copy_stream));
        mark(h->copy_stream);
        return 0;
    };
    const int nchunks = (int)sched.size();
    for (int c = 0; c < nchunks; ++c) {
        const long b0 = sched[c].first, nb = sched[c].second;
        if (with_forward) TRY(run_forward(h, b0, nb, -1, nullptr, true));
        TRY(run_reverse(h, b0, nb, -1, d_lamT, ldl, d_dJdv + b0, d_lam0 + b0, nullptr, nullptr));
        // transpose on the compute stream (it has slack: the sweeps run ~3x faster than PCIe drains them) into the
        // staging buffer whose previous download has finished, so that the copy stream only ever holds downloads
        // and the transfer never pauses
        mark(h->stream);
        double *stg = h->d_pipe[c & 1].p;
        if (c >= 2) CK(cudaStreamWaitEvent(h->stream, h->stage_free[c & 1], 0));
        TRY(from_soa(h, d_dJdv + b0, stg, rows, nb, Bt, 1, h->stream));
        if (lam0) TRY(from_soa(h, d_lam0 + b0, stg + (size_t)rows * cmax, m, nb, Bt, 1, h->stream));
        CK(cudaEventRecord(h->chunk_done[c & 1], h->stream));
        mark(h->stream);
        if (c >= 1) TRY(download(c - 1));
    }
    TRY(download(nchunks - 1));
    if (trace) {
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamSynchronize(h->stream);
        for (size_t i = 1; i < tev.size(); ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, tev[0], tev[i]);
            fprintf(stderr, "trace event %zu: %.3f ms\n", i, ms);
        }
        for (auto e : tev) cudaEventDestroy(e);
    }
    return 0;
}

static int adjoints_common(ocs_b200_handle *h, const double *d_lamT, long ldl, double *d_dJdv, double *d_lam0)
{
    const int nc = h->kt.nc;
    if (d_dJdv && needs_zero_fill(h)) CK(cudaMemsetAsync(d_dJdv, 0, sizeof(double) * nc * h->K * h->Bt, h->stream));
    return run_reverse(h, 0, h->Bt, -1, d_lamT, ldl, d_dJdv, d_lam0, nullptr, nullptr);
}

int ocs_b200_compute_adjoints_dev(ocs_b200_handle *h, const double *lamT_soa, long lamT_nbatch, double *dJdv_soa,
                                  double *lam0_soa)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi)
        return h->fail(OCS_B200_ERR_ARG, "ocs_b200_compute_adjoints_dev takes pointers / a stream of ONE device; on a multi-device handle use "
                                          "ocs_b200_evaluate_resident, ocs_b200_gather_results_dev and ocs_b200_get_results");
    TRY(check_ready(h));
    TRY(check_resident(h, "compute_adjoints"));
    if (lamT_soa && lamT_nbatch != 1 && lamT_nbatch != h->Bt) return h->fail(OCS_B200_ERR_ARG, "lamT must have 1 or Bt columns");
    try {
        CK(cudaSetDevice(h->device));
        h->user_lamT = false;
        return adjoints_common(h, lamT_soa, lamT_soa && lamT_nbatch > 1 ? h->Bt : 0, dJdv_soa, lam0_soa);
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

int ocs_b200_compute_adjoints(ocs_b200_handle *h, const double *lamT, long lamT_nbatch, double *dJdv, double *lam0)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_compute_adjoints(h, lamT, lamT_nbatch, dJdv, lam0);
    TRY(check_ready(h));
    TRY(check_resident(h, "compute_adjoints"));
    if (lamT && lamT_nbatch != 1 && lamT_nbatch != h->Bt) return h->fail(OCS_B200_ERR_ARG, "lamT must have 1 or Bt columns");
    try {
        CK(cudaSetDevice(h->device));
        const int n = h->kt.n, m = n + 1, nc = h->kt.nc;
        const long Bt = h->Bt;
        const double *d_lamT = nullptr;
        h->user_lamT = lamT != nullptr;
        h->user_lamT_ld = lamT && lamT_nbatch > 1 ? Bt : 0;
        if (lamT) {
            TRY(upload_soa(h, lamT, h->d_lamT_user, m, lamT_nbatch));
            d_lamT = h->d_lamT_user.p;
        }
        double *d_dJdv = nullptr, *d_lam0 = nullptr;
        if (dJdv) {
            TRY(ensure(h, h->d_dJdv, (size_t)nc * h->K * Bt));
            d_dJdv = h->d_dJdv.p;
        }
        TRY(ensure(h, h->d_lam0, (size_t)m * Bt));
        d_lam0 = h->d_lam0.p;
        const long rows = (long)nc * h->K;
        const long ldl = lamT && lamT_nbatch > 1 ? Bt : 0;
        if (!dJdv || Bt < 262144) {
            TRY(adjoints_common(h, d_lamT, ldl, d_dJdv, d_lam0));
            if (dJdv) TRY(download_soa(h, d_dJdv, Bt, 0, dJdv, rows, Bt));
            if (lam0) TRY(download_soa(h, d_lam0, Bt, 0, lam0, m, Bt));   // d_stage reuse is ordered by the stream
        } else {
            // Large batch: the gradient download (rows*Bt doubles over PCIe) dominates the call; it overlaps
            // the sweeps chunk by chunk.
            TRY(pipelined_gradient(h, false, d_lamT, ldl, d_dJdv, d_lam0, dJdv, lam0));
        }
        return finish_host(h);
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

int ocs_b200_evaluate(ocs_b200_handle *h, const double *x0, long Bt, const double *lamT, long lamT_nbatch, double *J,
                      double *xT, int32_t *status, double *dJdv, double *lam0)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_evaluate(h, x0, Bt, lamT, lamT_nbatch, J, xT, status, dJdv, lam0);
    TRY(check_ready(h));
    if (!x0 || Bt < 1) return h->fail(OCS_B200_ERR_ARG, "bad x0 / batch size");
    if (lamT && lamT_nbatch != 1 && lamT_nbatch != Bt) return h->fail(OCS_B200_ERR_ARG, "lamT must have 1 or Bt columns");
    // small batches (one chunk would not fill the device) or no gradient: the two calls in sequence.  With a tail
    // horizon a chunk runs main forward, tail forward, tail reverse, main reverse (RK4InfiniteIntegrator.m:20-30 for
    // the chunk's trajectories): trajectories are independent, so the chunking is the same.
    if (!dJdv && !lam0) return ocs_b200_compute_states(h, x0, Bt, J, xT, status);
    if (!dJdv || Bt < 262144) {
        // one pass: both sweeps are queued before anything is downloaded (and run as two lanes when that pays:
        // sweeps_common), then J, x(T), status and the gradient come down in that order
        try {
            CK(cudaSetDevice(h->device));
            const int n = h->kt.n, m = n + 1, nc = h->kt.nc;
            h->have_states = false;
            TRY(validate_batch(h, Bt));
            if (dJdv && h->basis_kind == BASIS_NONE) return h->fail(OCS_B200_ERR_STATE, "dJdv requested but no control basis is attached");
            TRY(alloc_batch(h, Bt));
            TRY(upload_soa(h, x0, h->d_x0, n, Bt));
            const double *d_lamT = nullptr;
            h->user_lamT = lamT != nullptr;
            h->user_lamT_ld = lamT && lamT_nbatch > 1 ? Bt : 0;
            if (lamT) {
                TRY(upload_soa(h, lamT, h->d_lamT_user, m, lamT_nbatch));
                d_lamT = h->d_lamT_user.p;
            }
            double *d_dJdv = nullptr;
            if (dJdv) {
                TRY(ensure(h, h->d_dJdv, (size_t)nc * h->K * Bt));
                d_dJdv = h->d_dJdv.p;
            }
            TRY(ensure(h, h->d_lam0, (size_t)m * Bt));
            TRY(sweeps_common(h, Bt, true, true, d_lamT, h->user_lamT_ld, d_dJdv, h->d_lam0.p));
            if (J) TRY(d2h(h, J, h->d_J.p, sizeof(double) * Bt));
            if (xT) TRY(download_soa(h, h->d_xT.p, Bt, 0, xT, n, Bt));
            if (status) TRY(d2h(h, status, h->d_status, sizeof(int) * Bt));
            if (dJdv) {
                TRY(finish_host(h));      // download_soa reuses one staging buffer
                TRY(download_soa(h, d_dJdv, Bt, 0, dJdv, (long)nc * h->K, Bt));
            }
            if (lam0) {
                TRY(finish_host(h));
                TRY(download_soa(h, h->d_lam0.p, Bt, 0, lam0, m, Bt));
            }
            return finish_host(h);
        } catch (const std::exception &ex) {
            return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
        }
    }
    try {
        CK(cudaSetDevice(h->device));
        const int n = h->kt.n, m = n + 1, nc = h->kt.nc;
        h->have_states = false;
        TRY(validate_batch(h, Bt));
        if (h->basis_kind == BASIS_NONE) return h->fail(OCS_B200_ERR_STATE, "dJdv requested but no control basis is attached");
        TRY(alloc_batch(h, Bt));
        TRY(upload_soa(h, x0, h->d_x0, n, Bt));
        h->Bt = Bt;
        h->fb_resident = false;
        const double *d_lamT = nullptr;
        h->user_lamT = lamT != nullptr;
        h->user_lamT_ld = lamT && lamT_nbatch > 1 ? Bt : 0;
        if (lamT) {
            TRY(upload_soa(h, lamT, h->d_lamT_user, m, lamT_nbatch));
            d_lamT = h->d_lamT_user.p;
        }
        TRY(ensure(h, h->d_dJdv, (size_t)nc * h->K * Bt));
        TRY(ensure(h, h->d_lam0, (size_t)m * Bt));
        TRY(pipelined_gradient(h, true, d_lamT, h->user_lamT_ld, h->d_dJdv.p, h->d_lam0.p, dJdv, lam0));
        h->have_states = true;
        if (J) TRY(d2h(h, J, h->d_J.p, sizeof(double) * Bt));
        if (xT) TRY(download_soa(h, h->d_xT.p, Bt, 0, xT, n, Bt));
        if (status) TRY(d2h(h, status, h->d_status, sizeof(int) * Bt));
        return finish_host(h);
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

// Evaluate on resident data, keep every result on the device: forward sweep (and, with want_gradient, the adjoint
// sweep with dJdv, lam0 into the handle's own buffers) of the batch whose initial states are x0 [n x Bt] (host), or
// of the resident initial states when x0 == NULL.  Asynchronous: returns once the work is queued.
int ocs_b200_evaluate_resident(ocs_b200_handle *h, const double *x0, long Bt, int want_gradient)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_evaluate_resident(h, x0, Bt, want_gradient);
    TRY(check_ready(h));
    try {
        CK(cudaSetDevice(h->device));
        const int n = h->kt.n, m = n + 1, nc = h->kt.nc;
        if (x0) {
            if (Bt < 1) return h->fail(OCS_B200_ERR_ARG, "bad batch size");
            h->have_states = false;
            TRY(validate_batch(h, Bt));
            TRY(alloc_batch(h, Bt));
            TRY(upload_soa(h, x0, h->d_x0, n, Bt));
        } else {
            if (h->Bt < 1 || h->d_x0.cap < (size_t)n * h->Bt) return h->fail(OCS_B200_ERR_STATE, "no resident initial states: pass x0 once");
            Bt = h->Bt;
            h->have_states = false;
            TRY(validate_batch(h, Bt));
        }
        if (want_gradient) {
            if (h->basis_kind == BASIS_NONE) return h->fail(OCS_B200_ERR_STATE, "a gradient needs an attached control basis");
            TRY(ensure(h, h->d_dJdv, (size_t)nc * h->K * Bt));
            TRY(ensure(h, h->d_lam0, (size_t)m * Bt));
            h->user_lamT = false;
            TRY(sweeps_common(h, Bt, true, true, nullptr, 0, h->d_dJdv.p, h->d_lam0.p));
        } else {
            TRY(states_common(h, Bt));
        }
        return 0;
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

// Results of the resident batch into host arrays (any may be NULL): J [Bt], xT [n x Bt], status [Bt], and after a
// gradient evaluation dJdv [(nc*K) x Bt], lam0 [m x Bt].
int ocs_b200_get_results(ocs_b200_handle *h, double *J, double *xT, int32_t *status, double *dJdv, double *lam0)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_get_results(h, J, xT, status, dJdv, lam0);
    TRY(check_ready(h));
    if (!h->have_states) return h->fail(OCS_B200_ERR_STATE, "no resident results");
    try {
        CK(cudaSetDevice(h->device));
        const int n = h->kt.n, m = n + 1, nc = h->kt.nc;
        const long Bt = h->Bt;
        if ((dJdv && h->d_dJdv.cap < (size_t)nc * h->K * Bt) || (lam0 && h->d_lam0.cap < (size_t)m * Bt))
            return h->fail(OCS_B200_ERR_STATE, "no resident gradient: evaluate with want_gradient first");
        if (J) TRY(d2h(h, J, h->d_J.p, sizeof(double) * Bt));
        if (xT) TRY(download_soa(h, h->d_xT.p, Bt, 0, xT, n, Bt));
        if (status) TRY(d2h(h, status, h->d_status, sizeof(int) * Bt));
        if (dJdv) {
            TRY(finish_host(h));      // download_soa reuses one staging buffer
            TRY(download_soa(h, h->d_dJdv.p, Bt, 0, dJdv, (long)nc * h->K, Bt));
        }
        if (lam0) {
            TRY(finish_host(h));
            TRY(download_soa(h, h->d_lam0.p, Bt, 0, lam0, m, Bt));
        }
        return finish_host(h);
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

// J [Bt] and x(T) [n][Bt] of the resident batch into DEVICE buffers on the root device (the handle's device, or
// devices[0] of a multi-device handle, which gathers the shards over NVLink).  Asynchronous on the root's stream.
int ocs_b200_gather_results_dev(ocs_b200_handle *h, double *J_root, double *xT_root)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_gather_results_dev(h, J_root, xT_root);
    if (!h->have_states) return h->fail(OCS_B200_ERR_STATE, "no resident results to gather");
    CK(cudaSetDevice(h->device));
    if (J_root) CK(cudaMemcpyAsync(J_root, h->d_J.p, sizeof(double) * h->Bt, cudaMemcpyDeviceToDevice, h->stream));
    if (xT_root) CK(cudaMemcpyAsync(xT_root, h->d_xT.p, sizeof(double) * h->kt.n * h->Bt, cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}

// Devices behind the handle and the shard offsets of the resident batch (offsets has ndev + 1 entries).
int ocs_b200_get_devices(const ocs_b200_handle *h, int *ndev, int *devices, long *offsets, int cap)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_device_info(h, ndev, devices, offsets, cap);
    if (ndev) *ndev = 1;
    if (devices && cap > 0) devices[0] = h->device;
    if (offsets && cap > 0) offsets[0] = 0;
    if (offsets && cap > 1) offsets[1] = h->have_states ? h->Bt : 0;
    return 0;
}

int ocs_b200_fetch_states(ocs_b200_handle *h, int segment, long b0, long nb, double *x)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_fetch_states(h, segment, b0, nb, x);
    TRY(check_ready(h));
    TRY(check_resident(h, "fetch_states"));
    if (segment < 0 || segment > 1 || (segment == 1 && !h->has_tail)) return h->fail(OCS_B200_ERR_ARG, "no such segment");
    if (b0 < 0 || nb < 1 || b0 + nb > h->Bt) return h->fail(OCS_B200_ERR_ARG, "trajectory range out of bounds");
    if (!x) return 0;
    try {
        CK(cudaSetDevice(h->device));
        const int m = h->kt.n + 1;
        const size_t cnt = (size_t)m * (h->seg[segment].N + 1) * nb;
        TRY(ensure(h, h->d_stage2, cnt));
        if (h->fb_resident) {
            // after fb_sweep the control grid already holds uNew while the resident states belong to the control the
            // last sweep started from (final_sweep(u), fb_sweep.m:82,117-122): they are read, not re-swept.  The
            // running-objective row is not part of the store: NaN except x(end,end) = J.
            const int NP = h->seg[segment].N + 1;
            store_to_grid_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, h->stream>>>(h->seg[segment].d_xs + b0, h->Bt, h->kt.n, NP, nb,
                                                                                      h->d_J.p + b0, h->d_stage2.p);
            CK(cudaGetLastError());
            h->launches++;
        } else {
            // re-run the (deterministic) forward sweep of the requested trajectories with the full grid
            // written out, J row included; the resident store is rewritten with identical values
            TRY(run_forward(h, b0, nb, segment, h->d_stage2.p, false));
        }
        TRY(d2h(h, x, h->d_stage2.p, sizeof(double) * cnt));
        return finish_host(h);
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

int ocs_b200_fetch_adjoints(ocs_b200_handle *h, int segment, long b0, long nb, double *lam, double *dJdu)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_fetch_adjoints(h, segment, b0, nb, lam, dJdu);
    TRY(check_ready(h));
    TRY(check_resident(h, "fetch_adjoints"));
    if (segment < 0 || segment > 1 || (segment == 1 && !h->has_tail)) return h->fail(OCS_B200_ERR_ARG, "no such segment");
    if (b0 < 0 || nb < 1 || b0 + nb > h->Bt) return h->fail(OCS_B200_ERR_ARG, "trajectory range out of bounds");
    try {
        CK(cudaSetDevice(h->device));
        const int m = h->kt.n + 1, nc = h->kt.nc;
        const int N = h->seg[segment].N, G = 2 * N + 1;
        const size_t nl = (size_t)m * (N + 1) * nb, nu = (size_t)nc * G * nb;
        TRY(ensure(h, h->d_stage2, nl + nu));
        double *d_lam = h->d_stage2.p, *d_du = h->d_stage2.p + nl;
        if (segment == 0 && h->has_tail) {
            // lamT of the main horizon = lam2(:,1) of the tail (RK4InfiniteIntegrator.m:28-29)
            TRY(run_reverse(h, b0, nb, 1, nullptr, 0, nullptr, nullptr, nullptr, nullptr));
        }
        // the terminal costate of the last compute_adjoints (host API) is reused
        const double *d_lamT = (segment == 0 && !h->has_tail && h->user_lamT) ? h->d_lamT_user.p : nullptr;
        TRY(run_reverse(h, b0, nb, segment, d_lamT, h->user_lamT_ld, nullptr, nullptr, lam ? d_lam : nullptr,
                        dJdu ? d_du : nullptr));
        if (lam) TRY(d2h(h, lam, d_lam, sizeof(double) * nl));
        if (dJdu) TRY(d2h(h, dJdu, d_du, sizeof(double) * nu));
        return finish_host(h);
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

int ocs_b200_fb_sweep(ocs_b200_handle *h, const double *x0, long Bt, const double *u0, long u0_nbatch,
                      const double *bounds, double uRelTol, double uAbsTol, int nSWEEPS, double relax, double *J,
                      int32_t *sweeps, int32_t *converged, double *max_change)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_fb_sweep(h, x0, Bt, u0, u0_nbatch, bounds, uRelTol, uAbsTol, nSWEEPS, relax, J, sweeps, converged, max_change);
    TRY(check_ready(h));
    if (!x0 || !bounds || Bt < 1 || nSWEEPS < 1 || (u0 && u0_nbatch != 1 && u0_nbatch != Bt))
        return h->fail(OCS_B200_ERR_ARG, "bad fb_sweep arguments");
    // u0 == NULL: start from the coefficients of ocs_b200_set_control_v, u0 = reshape(v,nc,[])*B
    std::vector<double> u0_from_v;
    bool u0_on_device = false;
    if (!u0) {
        if (h->basis_kind == BASIS_NONE || (h->ctrl_mode != CTRL_V_SHARED && h->ctrl_mode != CTRL_V_BATCH))
            return h->fail(OCS_B200_ERR_STATE, "fb_sweep without u0 needs a basis and coefficients v (ocs_b200_set_control_v)");
        TRY(validate_control(h, Bt));
        if (h->ctrl_mode == CTRL_V_SHARED) {
            const int G0 = 2 * h->seg[0].N + 1, nc0 = h->kt.nc;
            u0_from_v.resize((size_t)nc0 * G0);
            for (int j = 0; j < G0; ++j) shared_u_from_v(h, j, &u0_from_v[(size_t)nc0 * j]);
            u0 = u0_from_v.data();
            u0_nbatch = 1;
        } else {
            u0_on_device = true;
        }
    }
    if (h->has_tail) return h->fail(OCS_B200_ERR_ARG, "fb_sweep runs on a finite horizon");
    try {
        CK(cudaSetDevice(h->device));
        const int n = h->kt.n, nc = h->kt.nc, G = 2 * h->seg[0].N + 1;
        if (h->params_nbatch > 1 && h->params_nbatch != Bt) return h->fail(OCS_B200_ERR_ARG, "params batch mismatch");
        h->have_states = false;        // a failure below must not leave a half-replaced batch looking valid
        TRY(alloc_batch(h, Bt));
        h->Bt = Bt;
        h->fb_resident = false;
        TRY(upload_soa(h, x0, h->d_x0, n, Bt));
        // control on the grid, one copy per trajectory
        if (u0_on_device) {
            h->seg[0].basis_dirty = h->seg[0].basis_dirty || h->seg[0].BW == 0;
            TRY(build_records(h, 0));      // basis records of the attached control
            TRY(ensure(h, h->d_ug, (size_t)nc * G * Bt));
            const double *vp = h->v_ext ? h->v_ext : h->d_v.p;
            if (h->basis_kind == BASIS_DENSE && h->K <= BE_KC && getenv("OCS_B200_BASIS_EVAL_GENERIC") == nullptr) {
                // enough blocks for the device whatever the batch: the grid is cut into segments (blockIdx.y)
                const unsigned bx = (unsigned)((Bt + 127) / 128);
                int sms = 148;
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
                int ny = (int)std::max<long>(1, std::min<long>(32, (8L * sms + bx - 1) / bx));
                const int segl = ((G + ny - 1) / ny + BE_JT - 1) / BE_JT * BE_JT;
                ny = (G + segl - 1) / segl;
                basis_eval_dense_kernel<<<dim3(bx, (unsigned)ny), 128, sizeof(double) * BE_JT * h->seg[0].BW, h->stream>>>(
                    vp, h->seg[0].d_basrec, h->seg[0].BW, h->K, nc, G, segl, Bt, Bt, h->d_ug.p);
            } else {
                basis_eval_kernel<<<(unsigned)((Bt + 127) / 128), 128, 0, h->stream>>>(
                    vp, h->seg[0].d_basrec, h->seg[0].BW, h->basis_kind == BASIS_SPARSE, h->K, nc, G, Bt, Bt, h->d_ug.p);
            }
            CK(cudaGetLastError());
            h->launches++;
        } else if (u0_nbatch == Bt) {
            TRY(upload_soa(h, u0, h->d_ug, (long)nc * G, Bt, nc));
        } else {
            TRY(ensure(h, h->d_ug, (size_t)nc * G * Bt));
            TRY(ensure(h, h->d_stage, (size_t)nc * G));
            std::vector<double> perm((size_t)nc * G);
            for (int j = 0; j < G; ++j)
                for (int c = 0; c < nc; ++c) perm[(size_t)c * G + j] = u0[c + (size_t)nc * j];
            CK(cudaMemcpyAsync(h->d_stage.p, perm.data(), sizeof(double) * nc * G, cudaMemcpyHostToDevice, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            broadcast_rows_kernel<<<(unsigned)((Bt + 255) / 256), 256, 0, h->stream>>>(h->d_stage.p, h->d_ug.p, (long)nc * G, Bt);
            CK(cudaGetLastError());
        }
        h->v_ext = nullptr;
        h->ctrl_mode = CTRL_U_BATCH;
        h->ctrl_nbatch = Bt;
        h->seg[0].records_dirty = true;
        if (h->frozen_cap < (size_t)Bt) {
            cudaFree(h->d_frozen);
            cudaFree(h->d_sweeps);
            h->d_frozen = h->d_sweeps = nullptr;
            CK(cudaMalloc(&h->d_frozen, sizeof(int) * (Bt + 1)));
            CK(cudaMalloc(&h->d_sweeps, sizeof(int) * Bt));
            h->frozen_cap = Bt;
        }
        int *d_nopen = h->d_frozen + Bt;
        CK(cudaMemsetAsync(h->d_frozen, 0, sizeof(int) * (Bt + 1), h->stream));
        CK(cudaMemsetAsync(h->d_sweeps, 0, sizeof(int) * Bt, h->stream));
        TRY(ensure(h, h->d_maxchg, Bt));
        if (!h->d_bounds) CK(cudaMalloc(&h->d_bounds, sizeof(double) * 2 * 64));
        if (nc > 64) return h->fail(OCS_B200_ERR_ARG, "too many controls");
        {
            // ControlBounds is nc x 2 column-major; the kernels read [c][lo,hi]
            std::vector<double> bb(2 * nc);
            for (int c = 0; c < nc; ++c) {
                bb[2 * c] = bounds[c];
                bb[2 * c + 1] = bounds[c + nc];
            }
            CK(cudaMemcpyAsync(h->d_bounds, bb.data(), sizeof(double) * 2 * nc, cudaMemcpyHostToDevice, h->stream));
            CK(cudaStreamSynchronize(h->stream));
        }
        TRY(build_records(h, 0));
        const void *kf = nullptr, *kr = nullptr;
        TRY(resolve(h, KIND_FWD, 2, 0, pb_index(h), &kf));
        TRY(resolve(h, h->fb_costate == OCS_B200_COSTATE_CONTINUOUS ? KIND_FBC : KIND_FB, 2, 0, pb_index(h), &kr));
        const size_t fb_ring = (size_t)h->kt.fb_ring * (n + 2 * nc) * h->kt.threads;
        // one sweep (fb_sweep.m:79-112) of trajectories [b0, b0 + nb) on the handle's current stream:
        // x = forward(u); uNew = ControlChar(t, x, lam(u)) with the convergence measure; then the bookkeeping
        auto one_sweep = [&](long b0, long nb, int sweep, int *d_open) -> int {
            SweepArgs a;
            base_args(h, 0, a);
            a.nb = nb;
            a.ldb = Bt;
            if (a.ldp) a.p = h->d_params.p + b0;
            a.x0 = h->d_x0.p + b0;
            a.xs = h->seg[0].d_xs + b0;
            a.ug = h->d_ug.p + b0;
            a.J = h->d_J.p + b0;
            a.xT = h->d_xT.p + b0;
            a.status = h->d_status + b0;
            a.frozen = h->d_frozen + b0;
            TRY(launch(h, kf, a, true, fwd_tpt(h, 2)));
            SweepArgs r;
            base_args(h, 0, r);
            r.nb = nb;
            r.ldb = Bt;
            if (r.ldp) r.p = h->d_params.p + b0;
            r.xs = h->seg[0].d_xs + b0;
            r.ug = h->d_ug.p + b0;
            r.bounds = h->d_bounds;
            r.uRelTol = uRelTol;
            r.uAbsTol = uAbsTol;
            r.relax = relax;
            r.max_change = h->d_maxchg.p + b0;
            r.frozen = h->d_frozen + b0;
            TRY(launch(h, kr, r, false, 1, fb_ring));
            CK(cudaMemsetAsync(d_open, 0, sizeof(int), h->stream));
            fb_update_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, h->stream>>>(h->d_maxchg.p + b0, h->d_frozen + b0,
                                                                                 h->d_sweeps + b0, sweep, nb, d_open);
            CK(cudaGetLastError());
            return 0;
        };
        SweepArgs geom;
        base_args(h, 0, geom);
        for (int k = 0; k < 2; ++k)
            if (!h->ev_loop[k]) CK(cudaEventCreate(&h->ev_loop[k]));
        CK(cudaEventRecord(h->ev_loop[0], h->stream));
        if (!use_lanes(h, kr, Bt, 1, smem_bytes(h, geom, fb_ring))) {
            for (int sweep = 1; sweep <= nSWEEPS; ++sweep) {
                TRY(one_sweep(0, Bt, sweep, d_nopen));
                int nopen = 0;
                CK(cudaMemcpyAsync(&nopen, d_nopen, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
                CK(cudaStreamSynchronize(h->stream));
                if (nopen == 0) break;
            }
        } else {
            // Two lanes (see lanes_fork): each half runs its own chain of sweeps, so one half's forward sweep fills the
            // SMs the other half's costate sweep leaves idle.  The host looks at the open-trajectory counts ONE sweep
            // late (while the next sweep is already queued) so that the lanes never drain; a sweep queued after
            // everything has converged finds every trajectory frozen and changes nothing.
            if (!h->d_lane_open) {
                CK(cudaMalloc(&h->d_lane_open, sizeof(int) * 4));
                CK(cudaHostAlloc((void **)&h->h_lane_open, sizeof(int) * 4, cudaHostAllocDefault));
            }
            for (int l = 0; l < 2; ++l)
                for (int q = 0; q < 2; ++q)
                    if (!h->lane_sweep[l][q]) CK(cudaEventCreateWithFlags(&h->lane_sweep[l][q], cudaEventDisableTiming));
            long lb0[2], lnb[2];
            lane_split(h, Bt, lb0, lnb);
            TRY(lanes_fork(h));
            cudaStream_t main_stream = h->stream;
            int rc = 0;
            for (int sweep = 1; sweep <= nSWEEPS && rc == 0; ++sweep) {
                const int q = sweep & 1;
                for (int l = 0; l < 2 && rc == 0; ++l) {
                    h->stream = h->lane_stream[l];
                    rc = one_sweep(lb0[l], lnb[l], sweep, h->d_lane_open + 2 * l + q);
                    if (rc == 0 && cudaMemcpyAsync(h->h_lane_open + 2 * l + q, h->d_lane_open + 2 * l + q, sizeof(int),
                                                   cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) rc = OCS_B200_ERR_CUDA;
                    if (rc == 0 && cudaEventRecord(h->lane_sweep[l][q], h->stream) != cudaSuccess) rc = OCS_B200_ERR_CUDA;
                }
                h->stream = main_stream;
                if (rc != 0 || sweep == 1) continue;
                const int qp = (sweep - 1) & 1;      // the previous sweep's counts
                int open = 0;
                for (int l = 0; l < 2; ++l) {
                    if (cudaEventSynchronize(h->lane_sweep[l][qp]) != cudaSuccess) rc = OCS_B200_ERR_CUDA;
                    open += h->h_lane_open[2 * l + qp];
                }
                if (open == 0) break;
            }
            h->stream = main_stream;
            if (rc != 0) {
                for (int l = 0; l < 2; ++l) cudaStreamSynchronize(h->lane_stream[l]);
                return rc == OCS_B200_ERR_CUDA ? h->fail(rc, "CUDA error while queueing the fb_sweep lanes: %s",
                                                         cudaGetErrorString(cudaGetLastError())) : rc;
            }
            TRY(lanes_join(h));
        }
        CK(cudaEventRecord(h->ev_loop[1], h->stream));
        h->have_states = true;
        h->fb_resident = true;      // x, J: the sweep's input control (final_sweep(u), fb_sweep.m:82); d_ug: uNew = soln.u (:123)
        if (J) TRY(d2h(h, J, h->d_J.p, sizeof(double) * Bt));
        if (sweeps) TRY(d2h(h, sweeps, h->d_sweeps, sizeof(int) * Bt));
        if (converged) TRY(d2h(h, converged, h->d_frozen, sizeof(int) * Bt));
        if (max_change) TRY(d2h(h, max_change, h->d_maxchg.p, sizeof(double) * Bt));
        TRY(finish_host(h));
        CK(cudaEventSynchronize(h->ev_loop[1]));
        float ms_loop = 0;
        CK(cudaEventElapsedTime(&ms_loop, h->ev_loop[0], h->ev_loop[1]));
        h->ms_loop += ms_loop;
        return 0;
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

int ocs_b200_set_control_ufunc(ocs_b200_handle *h, int kind, int nknots, const double *knots)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_set_control_ufunc(h, kind, nknots, knots);
    if (kind == OCS_B200_UFUNC_NONE) {
        h->ufunc_kind = kind;
        h->ufunc_knots.clear();
        return 0;
    }
    if (kind < OCS_B200_UFUNC_PREVIOUS || kind > OCS_B200_UFUNC_CHEBYSHEV || !knots)
        return h->fail(OCS_B200_ERR_ARG, "bad uFunc description");
    if (h->basis_kind == BASIS_NONE) return h->fail(OCS_B200_ERR_STATE, "attach the control basis first (ocs_b200_set_control_basis)");
    if ((kind == OCS_B200_UFUNC_CHEBYSHEV && nknots != 2) || (kind != OCS_B200_UFUNC_CHEBYSHEV && nknots != h->K))
        return h->fail(OCS_B200_ERR_ARG, "uFunc needs K = %d knots (Chebyshev: the two end times)", h->K);
    for (int k = 0; k < nknots; ++k)
        if (!std::isfinite(knots[k]) || (k > 0 && !(knots[k] > knots[k - 1])))
            return h->fail(OCS_B200_ERR_ARG, "uFunc knots must be finite and strictly increasing");
    h->ufunc_kind = kind;
    h->ufunc_knots.assign(knots, knots + nknots);
    return 0;
}

int ocs_b200_eval_solution(ocs_b200_handle *h, int segment, long b0, long nb, int nq, const double *tq, double *x,
                           double *lam, double *u)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_eval_solution(h, segment, b0, nb, nq, tq, x, lam, u);
    TRY(check_ready(h));
    TRY(check_resident(h, "eval_solution"));
    if (segment < 0 || segment > 1 || (segment == 1 && !h->has_tail)) return h->fail(OCS_B200_ERR_ARG, "no such segment");
    if (b0 < 0 || nb < 1 || b0 + nb > h->Bt) return h->fail(OCS_B200_ERR_ARG, "trajectory range out of bounds");
    if (nq < 1 || !tq) return h->fail(OCS_B200_ERR_ARG, "no query times");
    for (int q = 0; q < nq; ++q)
        if (!std::isfinite(tq[q])) return h->fail(OCS_B200_ERR_ARG, "query times must be finite");
    try {
        CK(cudaSetDevice(h->device));
        using namespace ocs_interp;
        const int n = h->kt.n, nc = h->kt.nc;
        Segment &s = h->seg[segment];
        const int N = s.N, NP = N + 1, G = 2 * N + 1;
        const long Bt = h->Bt;
        const int maxrows = std::max(n, nc);
        // scratch: abscissae (G) | tq | query records | knots | SoA result | MATLAB-layout result | lam node store
        const size_t n_out = (size_t)maxrows * nq * nb;
        const size_t n_lams = lam ? (size_t)n * NP * nb : 0;
        const size_t n_knots = std::max<size_t>(h->ufunc_knots.size(), 2);
        const size_t total = (size_t)G + nq + (size_t)QREC * nq + n_knots + 2 * n_out + n_lams + (size_t)nc * h->K + 8;
        TRY(ensure(h, h->d_stage2, total));
        double *d_abs = h->d_stage2.p, *d_tq = d_abs + G, *d_qrec = d_tq + nq, *d_knots = d_qrec + (size_t)QREC * nq,
               *d_soa = d_knots + n_knots, *d_out = d_soa + n_out, *d_lams = d_out + n_out, *d_vsh = d_lams + n_lams;
        TRY(h2d(h, d_tq, tq, sizeof(double) * nq));
        const int qchunk = 32;
        const dim3 grid((unsigned)((nb + 127) / 128), (unsigned)((nq + qchunk - 1) / qchunk));
        auto prepare = [&](const std::vector<double> &xs) -> int {
            CK(cudaMemcpyAsync(d_abs, xs.data(), sizeof(double) * xs.size(), cudaMemcpyHostToDevice, h->stream));
            interp_prepare_kernel<<<(nq + 127) / 128, 128, 0, h->stream>>>(d_abs, (int)xs.size(), d_tq, nq, d_qrec);
            CK(cudaGetLastError());
            h->launches++;
            return 0;
        };
        auto deliver = [&](double *host, int rows) -> int {
            TRY(from_soa(h, d_soa, d_out, (long)rows * nq, nb, nb, rows));
            TRY(d2h(h, host, d_out, sizeof(double) * rows * nq * nb));
            CK(cudaStreamSynchronize(h->stream));   // the scratch is reused by the next output
            return 0;
        };
        if (x || lam) TRY(prepare(s.tspan));
        if (x) {      // soln.x: pchip through the resident node states   single_shooting.m:129
            pchip_eval_kernel<<<grid, 128, 0, h->stream>>>(s.d_xs + b0, Bt, NP, n, d_qrec, nq, qchunk, d_soa, nb, nb);
            CK(cudaGetLastError());
            h->launches++;
            TRY(deliver(x, n));
        }
        if (lam) {    // soln.lam: the costate sweep is re-run for the range with its node values kept   :130
            if (segment == 0 && h->has_tail) TRY(run_reverse(h, b0, nb, 1, nullptr, 0, nullptr, nullptr, nullptr, nullptr));
            const double *d_lamT = (segment == 0 && !h->has_tail && h->user_lamT) ? h->d_lamT_user.p : nullptr;
            h->lams_req = d_lams;
            h->lams_ld = nb;
            h->lams_segment = segment;
            int rc = run_reverse(h, b0, nb, segment, d_lamT, h->user_lamT_ld, nullptr, nullptr, nullptr, nullptr);
            h->lams_req = nullptr;
            if (rc != 0) return rc;
            pchip_eval_kernel<<<grid, 128, 0, h->stream>>>(d_lams, nb, NP, n, d_qrec, nq, qchunk, d_soa, nb, nb);
            CK(cudaGetLastError());
            h->launches++;
            TRY(deliver(lam, n));
        }
        if (u) {      // soln.u   single_shooting.m:128 / fb_sweep.m:123
            const int um = seg_umode(h, segment);
            if (um == 2) {
                // control on the grid (fb_sweep): pchip through the G grid values, as fb_sweep.m:123 does
                // through its interpPts
                TRY(prepare(s.t));
                pchip_eval_kernel<<<grid, 128, 0, h->stream>>>(h->d_ug.p + b0, Bt, G, nc, d_qrec, nq, qchunk, d_soa, nb, nb);
            } else if (um == 1 || (segment == 0 && h->ctrl_mode == CTRL_V_SHARED)) {
                if (h->ufunc_kind == OCS_B200_UFUNC_NONE)
                    return h->fail(OCS_B200_ERR_STATE, "soln.u needs the control's uFunc (ocs_b200_set_control_ufunc)");
                CK(cudaMemcpyAsync(d_knots, h->ufunc_knots.data(), sizeof(double) * h->ufunc_knots.size(),
                                   cudaMemcpyHostToDevice, h->stream));
                const double *vp;
                long ldv;
                int vb;
                if (h->ctrl_mode == CTRL_V_SHARED) {
                    CK(cudaMemcpyAsync(d_vsh, h->ctrl_shared.data(), sizeof(double) * nc * h->K, cudaMemcpyHostToDevice,
                                       h->stream));
                    vp = d_vsh;
                    ldv = 1;
                    vb = 0;
                } else {
                    vp = (h->v_ext ? h->v_ext : h->d_v.p) + b0;
                    ldv = Bt;
                    vb = 1;
                }
                ufunc_eval_kernel<<<grid, 128, 0, h->stream>>>(h->ufunc_kind, vp, ldv, vb, h->K, nc, d_knots, d_tq, nq, qchunk,
                                                               d_soa, nb, nb);
            } else {
                return h->fail(OCS_B200_ERR_STATE, "soln.u needs control coefficients v or a per-trajectory control grid");
            }
            CK(cudaGetLastError());
            h->launches++;
            TRY(deliver(u, nc));
        }
        return finish_host(h);
    } catch (const std::exception &ex) {
        h->lams_req = nullptr;
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

int ocs_b200_compute_equilibrium(ocs_b200_handle *h, const double *params, int np, long nparams, long nsets, const double *guess,
                                 long nguess, const double *lb, const double *ub, const double *r, long nr, int max_iter,
                                 double tol, double *y, double *resnorm, double *residual, int32_t *exitflag, int32_t *iters)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_compute_equilibrium(h, params, np, nparams, nsets, guess, nguess, lb, ub, r, nr, max_iter, tol, y, resnorm, residual, exitflag, iters);
    if (h->sticky) return h->sticky;
    if (!h->have_problem) return h->fail(OCS_B200_ERR_STATE, "no problem bound (ocs_b200_set_problem_*)");
    if (np != h->kt.np) return h->fail(OCS_B200_ERR_ARG, "the problem has %d parameters, got %d", h->kt.np, np);
    if ((np > 0 && !params) || !guess || !r || !y || nsets < 1 || (nguess != 1 && nguess != nsets) ||
        (nr != 1 && nr != nsets) || (nparams != 1 && nparams != nsets) || max_iter < 1 || !(tol > 0.0))
        return h->fail(OCS_B200_ERR_ARG, "bad compute_equilibrium arguments");
    try {
        CK(cudaSetDevice(h->device));
        const int NY = 2 * h->kt.n + h->kt.nc;
        if (!h->kt.eq && h->is_generated) {
            const char *names[1] = {"ocs_eq_gen"};
            const void *kern[1] = {nullptr};
            void *lib = nullptr;
            std::string log;
            int rc = ocs_compile_unit(h->gen_source.c_str(), "OCS_EQ_KERNEL(ocs_gen::Problem, gen)",
                                      nvrtc_flags(h),
                                      h->gen_include.empty() ? nullptr : h->gen_include.c_str(), names, 1, kern, &lib, &log,
                                      h->gen_n);
            if (rc != 0) return h->fail(rc, "equilibrium kernel of the generated problem did not compile:\n%s", log.c_str());
            h->gen_libraries.push_back(lib);
            h->kt.eq = kern[0];
        }
        if (!h->kt.eq) return h->fail(OCS_B200_ERR_ARG, "the bound problem has no equilibrium kernel");
        // one scratch allocation: params | guess | lb | ub | r | y | resnorm | residual | (int) exitflag, iters
        const size_t n_par = (size_t)std::max(np, 1) * nparams, n_gs = (size_t)NY * nguess, n_y = (size_t)NY * nsets;
        const size_t total = n_par + n_gs + 2 * NY + (size_t)nr + n_y + (size_t)nsets + n_y + (size_t)nsets + 2;
        TRY(ensure(h, h->d_stage2, total));
        double *d = h->d_stage2.p;
        double *d_par = d, *d_gs = d_par + n_par, *d_lb = d_gs + n_gs, *d_ub = d_lb + NY, *d_r = d_ub + NY, *d_y = d_r + nr,
               *d_rn = d_y + n_y, *d_res = d_rn + nsets;
        int *d_flag = reinterpret_cast<int *>(d_res + n_y), *d_it = d_flag + nsets;
        if (np > 0) TRY(h2d(h, d_par, params, sizeof(double) * np * nparams));
        TRY(h2d(h, d_gs, guess, sizeof(double) * n_gs));
        if (lb) TRY(h2d(h, d_lb, lb, sizeof(double) * NY));
        if (ub) TRY(h2d(h, d_ub, ub, sizeof(double) * NY));
        TRY(h2d(h, d_r, r, sizeof(double) * nr));
        EqArgs a;
        memset(&a, 0, sizeof a);
        a.nsets = nsets;
        a.np = np;
        a.params = d_par;
        a.nparams = nparams;
        a.guess = d_gs;
        a.nguess = nguess;
        a.lb = lb ? d_lb : nullptr;
        a.ub = ub ? d_ub : nullptr;
        a.r = d_r;
        a.nr = nr;
        a.max_iter = max_iter;
        a.tol = tol;
        a.y = d_y;
        a.resnorm = d_rn;
        a.residual = d_res;
        a.exitflag = d_flag;
        a.iters = d_it;
        void *args[] = {(void *)&a};
        CK(cudaLaunchKernel(h->kt.eq, dim3((unsigned)((nsets + 127) / 128)), dim3(128), args, 0, h->stream));
        h->launches++;
        TRY(d2h(h, y, d_y, sizeof(double) * n_y));
        if (resnorm) TRY(d2h(h, resnorm, d_rn, sizeof(double) * nsets));
        if (residual) TRY(d2h(h, residual, d_res, sizeof(double) * n_y));
        if (exitflag) TRY(d2h(h, exitflag, d_flag, sizeof(int) * nsets));
        if (iters) TRY(d2h(h, iters, d_it, sizeof(int) * nsets));
        return finish_host(h);
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

int ocs_b200_set_fb_costate(ocs_b200_handle *h, int mode)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_set_fb_costate(h, mode);
    if (mode != OCS_B200_COSTATE_DISCRETE && mode != OCS_B200_COSTATE_CONTINUOUS)
        return h->fail(OCS_B200_ERR_ARG, "costate mode must be OCS_B200_COSTATE_DISCRETE or _CONTINUOUS");
    h->fb_costate = mode;
    return 0;
}

int ocs_b200_single_shooting(ocs_b200_handle *h, const double *x0, long Bt, const double *v0, long v0_nbatch,
                             const double *Lb, const double *Ub, int nfree, const int32_t *free_states,
                             const double *free_lb, const double *free_ub, int max_iter, double tol, double *v_opt,
                             double *J, int32_t *iters, int32_t *converged)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_single_shooting(h, x0, Bt, v0, v0_nbatch, Lb, Ub, nfree, free_states, free_lb, free_ub, max_iter, tol, v_opt, J, iters, converged);
    TRY(check_ready(h));
    if (!x0 || !v0 || !Lb || !Ub || Bt < 1 || max_iter < 1 || (v0_nbatch != 1 && v0_nbatch != Bt) || nfree < 0 ||
        (nfree > 0 && (!free_states || !free_lb || !free_ub)))
        return h->fail(OCS_B200_ERR_ARG, "bad single_shooting arguments");
    if (h->basis_kind == BASIS_NONE) return h->fail(OCS_B200_ERR_STATE, "single_shooting needs an attached control basis");
    try {
        CK(cudaSetDevice(h->device));
        const int n = h->kt.n, m = n + 1, nc = h->kt.nc;
        const long nv = (long)nc * h->K, nz = nv + nfree;
        for (int f = 0; f < nfree; ++f)
            if (free_states[f] < 0 || free_states[f] >= n) return h->fail(OCS_B200_ERR_ARG, "free state index out of range");
        if (h->params_nbatch > 1 && h->params_nbatch != Bt) return h->fail(OCS_B200_ERR_ARG, "params batch mismatch");
        h->have_states = false;        // a failure below must not leave a half-replaced batch looking valid
        TRY(alloc_batch(h, Bt));
        h->Bt = Bt;
        h->fb_resident = false;
        TRY(upload_soa(h, x0, h->d_x0, n, Bt));
        TRY(ensure(h, h->d_z, (size_t)nz * Bt));
        TRY(ensure(h, h->d_g, (size_t)nz * Bt));
        TRY(ensure(h, h->d_zt, (size_t)nz * Bt));
        TRY(ensure(h, h->d_gt, (size_t)nz * Bt));
        TRY(ensure(h, h->d_Jacc, Bt));
        TRY(ensure(h, h->d_Jhist, (size_t)SPG_M * Bt));
        TRY(ensure(h, h->d_alpha, Bt));
        TRY(ensure(h, h->d_pred, Bt));
        TRY(ensure(h, h->d_lbub, 2 * (size_t)nz));
        TRY(ensure(h, h->d_lam0, (size_t)m * Bt));
        // starting point: zt = [v0; x0(free)] projected onto the box
        {
            std::vector<double> lbub(2 * (size_t)nz);
            for (long k = 0; k < nv; ++k) { lbub[k] = Lb[k]; lbub[nz + k] = Ub[k]; }
            for (int f = 0; f < nfree; ++f) { lbub[nv + f] = free_lb[f]; lbub[nz + nv + f] = free_ub[f]; }
            CK(cudaMemcpyAsync(h->d_lbub.p, lbub.data(), sizeof(double) * 2 * nz, cudaMemcpyHostToDevice, h->stream));
            if (v0_nbatch == Bt) {
                TRY(ensure(h, h->d_stage, (size_t)nv * Bt));
                TRY(h2d(h, h->d_stage.p, v0, sizeof(double) * nv * Bt));
                TRY(to_soa(h, h->d_stage.p, h->d_zt.p, nv, Bt, Bt));
            } else {
                std::vector<double> v0p(v0, v0 + nv);
                for (long k = 0; k < nv; ++k) v0p[k] = std::min(Ub[k], std::max(Lb[k], v0p[k]));
                TRY(ensure(h, h->d_stage, (size_t)nv));
                CK(cudaMemcpyAsync(h->d_stage.p, v0p.data(), sizeof(double) * nv, cudaMemcpyHostToDevice, h->stream));
                CK(cudaStreamSynchronize(h->stream));
                broadcast_rows_kernel<<<(unsigned)((Bt + 255) / 256), 256, 0, h->stream>>>(h->d_stage.p, h->d_zt.p, nv, Bt);
                CK(cudaGetLastError());
            }
            for (int f = 0; f < nfree; ++f)
                CK(cudaMemcpyAsync(h->d_zt.p + (size_t)(nv + f) * Bt, h->d_x0.p + (size_t)free_states[f] * Bt,
                                   sizeof(double) * Bt, cudaMemcpyDeviceToDevice, h->stream));
            CK(cudaStreamSynchronize(h->stream));   // lbub is a local vector
        }
        if (h->frozen_cap < (size_t)Bt) {
            cudaFree(h->d_frozen);
            cudaFree(h->d_sweeps);
            h->d_frozen = h->d_sweeps = nullptr;
            CK(cudaMalloc(&h->d_frozen, sizeof(int) * (Bt + 1)));
            CK(cudaMalloc(&h->d_sweeps, sizeof(int) * Bt));
            h->frozen_cap = Bt;
        }
        int *d_nopen = h->d_frozen + Bt;
        CK(cudaMemsetAsync(h->d_frozen, 0, sizeof(int) * (Bt + 1), h->stream));
        CK(cudaMemsetAsync(h->d_sweeps, 0, sizeof(int) * Bt, h->stream));
        // the sweeps read the trial coefficients in place
        h->v_ext = h->d_zt.p;
        h->ctrl_mode = CTRL_V_BATCH;
        h->ctrl_nbatch = Bt;
        h->seg[0].records_dirty = true;
        const bool spg_serial = nz > 512 || getenv("OCS_B200_SPG_SERIAL") != nullptr;   // (the env switch is for the A/B test)
        // A/B switches: OCS_B200_SPG_L2 = 0 selects the register-held split kernel, 2 the L2 re-read kernel over 64 adjacent
        // trajectories per block (default 1: 32); OCS_B200_SPG_UNROLL = 1 | 4 its rows in flight per thread (default 2)
        int spg_l2 = 1, spg_unr = 2;
        if (const char *e = getenv("OCS_B200_SPG_L2")) spg_l2 = atoi(e);
        if (const char *e = getenv("OCS_B200_SPG_UNROLL")) spg_unr = atoi(e);
        if (spg_l2 != 1 && spg_l2 != 2) spg_l2 = 0;
        // ControlBounds on the grid (ocs_b200_set_solve_control_bounds): an outer augmented-Lagrangian iteration around
        // the projected-gradient solve.  Only the dense-basis gradient path has the two grids the penalty works on;
        // a sparse basis bounds u(t) through the coefficient box already (compute_nlp_bounds).
        const bool cgb = h->cgb_on && h->basis_kind == BASIS_DENSE && !h->has_tail;
        const int G0 = 2 * h->seg[0].N + 1;
        if (cgb) {
            if ((int)h->cgb_bounds.size() != 2 * nc) return h->fail(OCS_B200_ERR_ARG, "solve control bounds do not match the problem's controls");
            TRY(ensure(h, h->d_cgb_mu, 2 * (size_t)nc * G0 * Bt));
            TRY(ensure(h, h->d_cgb_b, 2 * (size_t)nc + 1));      // [nc][lo, hi] and the batch's largest violation
            TRY(ensure(h, h->d_cgb_viol, (size_t)Bt));
            CK(cudaMemsetAsync(h->d_cgb_mu.p, 0, sizeof(double) * 2 * nc * G0 * Bt, h->stream));
            CK(cudaMemcpyAsync(h->d_cgb_b.p, h->cgb_bounds.data(), sizeof(double) * 2 * nc, cudaMemcpyHostToDevice, h->stream));
            h->cgb_rho = h->cgb_rho0;
        }
        struct ActiveGuard {       // the penalty is part of this call's evaluations only
            ocs_b200_handle *h;
            ~ActiveGuard() { h->cgb_active = false; h->fb_frozen = nullptr; }
        } guard{h};
        h->cgb_active = cgb;
        double cgb_worst_prev = INFINITY;
        int iter = 0;
        for (int outer = 0; outer < (cgb ? h->cgb_outer : 1) && iter < max_iter; ++outer) {
        for (int inner = 0; iter < max_iter; ++iter, ++inner) {
            for (int f = 0; f < nfree; ++f)      // x0(FreeInitStates) = v(end-nFREE_STATES+1:end)   single_shooting.m:145
                CK(cudaMemcpyAsync(h->d_x0.p + (size_t)free_states[f] * Bt, h->d_zt.p + (size_t)(nv + f) * Bt,
                                   sizeof(double) * Bt, cudaMemcpyDeviceToDevice, h->stream));
            // nlpObjective for the whole batch; converged trajectories are skipped by the kernels (frozen)
            {
                const int n_seg = h->has_tail ? 2 : 1;
                for (int si = 0; si < n_seg; ++si) {
                    TRY(build_records(h, si));
                }
            }
            h->fb_frozen = h->d_frozen;
            int rc = run_forward(h, 0, Bt, -1, nullptr, true);
            if (rc == 0) {
                // store-once gradient columns (sparse basis, monotone knot index reaching every knot) need no
                // zero-fill; frozen trajectories are skipped by the sweeps and their gt is never read again
                if (!(h->basis_kind == BASIS_SPARSE && h->bk_monotone && h->bk_covered))
                    CK(cudaMemsetAsync(h->d_gt.p, 0, sizeof(double) * nv * Bt, h->stream));
                rc = run_reverse(h, 0, Bt, -1, nullptr, 0, h->d_gt.p, h->d_lam0.p, nullptr, nullptr);
            }
            h->fb_frozen = nullptr;
            if (rc != 0) return rc;
            for (int f = 0; f < nfree; ++f)      // dJdv = [compute_dJdv(dJdu); lam(FreeInitStates,1)]   :148
                CK(cudaMemcpyAsync(h->d_gt.p + (size_t)(nv + f) * Bt, h->d_lam0.p + (size_t)free_states[f] * Bt,
                                   sizeof(double) * Bt, cudaMemcpyDeviceToDevice, h->stream));
            CK(cudaMemsetAsync(d_nopen, 0, sizeof(int), h->stream));
            if (h->profiling) CK(cudaEventRecord(h->ev[2], h->stream));
            {
                // rows of a trajectory over G thread groups of up to 16 rows each (nz <= 512); beyond that the
                // one-thread-per-trajectory kernel
                const int Gk = (int)((nz + 15) / 16);
                const dim3 blk(32, Gk), grd((unsigned)((Bt + 31) / 32));
                const size_t sm = sizeof(double) * 5 * Gk * 32;
                if (spg_serial) {
                    spg_update_kernel<<<(unsigned)((Bt + 127) / 128), 128, 0, h->stream>>>(
                        nz, Bt, h->d_z.p, h->d_g.p, h->d_zt.p, h->d_gt.p, h->d_Jacc.p, h->d_J.p, h->d_Jhist.p, h->d_alpha.p,
                        h->d_lbub.p, h->d_lbub.p + nz, h->d_frozen, h->d_sweeps, iter + 1, tol, inner == 0, d_nopen);
                } else if (spg_l2 && Gk * 32 * spg_l2 <= 512 * spg_l2) {
                    const dim3 blk2(32 * spg_l2, Gk), grd2((unsigned)((Bt + 32 * spg_l2 - 1) / (32 * spg_l2)));
                    auto kern = spg_l2 == 2 ? spg_update_l2_kernel<2, 2> : spg_unr == 1 ? spg_update_l2_kernel<1, 1>
                              : spg_unr == 4 ? spg_update_l2_kernel<1, 4> : spg_update_l2_kernel<1, 2>;
                    kern<<<grd2, blk2, sm * spg_l2, h->stream>>>(
                        nz, Bt, h->d_z.p, h->d_g.p, h->d_zt.p, h->d_gt.p, h->d_Jacc.p, h->d_J.p, h->d_Jhist.p, h->d_alpha.p,
                        h->d_pred.p, h->d_lbub.p, h->d_lbub.p + nz, h->d_frozen, h->d_sweeps, iter + 1, tol, inner == 0, d_nopen);
                } else if (Gk <= 16) {
                    spg_update_split_kernel<16, 512><<<grd, blk, sm, h->stream>>>(
                        nz, Bt, h->d_z.p, h->d_g.p, h->d_zt.p, h->d_gt.p, h->d_Jacc.p, h->d_J.p, h->d_Jhist.p, h->d_alpha.p,
                        h->d_pred.p, h->d_lbub.p, h->d_lbub.p + nz, h->d_frozen, h->d_sweeps, iter + 1, tol, inner == 0, d_nopen);
                } else {
                    spg_update_split_kernel<16, 1024><<<grd, blk, sm, h->stream>>>(
                        nz, Bt, h->d_z.p, h->d_g.p, h->d_zt.p, h->d_gt.p, h->d_Jacc.p, h->d_J.p, h->d_Jhist.p, h->d_alpha.p,
                        h->d_pred.p, h->d_lbub.p, h->d_lbub.p + nz, h->d_frozen, h->d_sweeps, iter + 1, tol, inner == 0, d_nopen);
                }
            }
            h->launches++;
            CK(cudaGetLastError());
            // the update left the accepted point in (zt, gt) and wrote the new trial point into z: swap the roles
            std::swap(h->d_z, h->d_zt);
            std::swap(h->d_g, h->d_gt);
            h->v_ext = h->d_zt.p;
            if (h->profiling) {
                CK(cudaEventRecord(h->ev[3], h->stream));
                CK(cudaEventSynchronize(h->ev[3]));
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]));
                h->ms_upd += ms;
            }
            int nopen = 0;
            CK(cudaMemcpyAsync(&nopen, d_nopen, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            if (nopen == 0) {
                ++iter;
                break;
            }
        }
        // leave the device state at the minimisers: one last evaluation of the accepted points (zt == z for
        // frozen trajectories; the others get their best accepted point) -- J without any penalty term
        // Without the outer (constraint) iteration the minimisers in z are final here: their download -- the transposition
        // and the 8 nz Bt bytes over PCIe, 17 ms of a 2^20 x 101 solve -- starts on the copy stream now and runs under the
        // last evaluation instead of after it.
        const bool early_v = !cgb && v_opt && nz > 1;
        if (early_v) {
            if (!h->copy_stream) CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
            if (!h->chunk_done[0]) CK(cudaEventCreateWithFlags(&h->chunk_done[0], cudaEventDisableTiming));
            TRY(ensure(h, h->d_stage, (size_t)nz * Bt));
            CK(cudaEventRecord(h->chunk_done[0], h->stream));
        }
        CK(cudaMemcpyAsync(h->d_zt.p, h->d_z.p, sizeof(double) * nz * Bt, cudaMemcpyDeviceToDevice, h->stream));
        for (int f = 0; f < nfree; ++f)
            CK(cudaMemcpyAsync(h->d_x0.p + (size_t)free_states[f] * Bt, h->d_zt.p + (size_t)(nv + f) * Bt,
                               sizeof(double) * Bt, cudaMemcpyDeviceToDevice, h->stream));
        TRY(run_forward(h, 0, Bt, -1, nullptr, true));
        if (early_v) {
            // issued after the sweep's launches: a pageable destination blocks this thread while it drains through the
            // pinned ring, and the device has its work queued meanwhile
            CK(cudaStreamWaitEvent(h->copy_stream, h->chunk_done[0], 0));
            TRY(from_soa(h, h->d_z.p, h->d_stage.p, nz, Bt, Bt, 1, h->copy_stream));
            TRY(d2h(h, v_opt, h->d_stage.p, sizeof(double) * nz * Bt, h->copy_stream));
            v_opt = nullptr;      // delivered by finish_host
        }
        if (!cgb) break;
        // multipliers, violations, and the trajectories that go round again (u of the accepted points is on the grid now)
        CK(cudaMemsetAsync(d_nopen, 0, sizeof(int), h->stream));
        unsigned long long *d_worst = reinterpret_cast<unsigned long long *>(h->d_cgb_b.p + 2 * nc);
        CK(cudaMemsetAsync(d_worst, 0, sizeof(unsigned long long), h->stream));
        cgb_update_kernel<<<(unsigned)((Bt + 127) / 128), 128, 0, h->stream>>>(h->d_ug.p, h->d_cgb_mu.p, h->d_cgb_b.p, h->cgb_rho, nc,
                                                                              G0, Bt, Bt, h->d_cgb_viol.p, h->d_frozen, h->cgb_tol,
                                                                              d_nopen, d_worst);
        CK(cudaGetLastError());
        h->launches++;
        int reopened = 0;
        double worst = 0.0;
        CK(cudaMemcpyAsync(&reopened, d_nopen, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(&worst, d_worst, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (reopened == 0) break;
        // the usual safeguard: the penalty parameter grows only when the multiplier update alone did not cut the
        // largest violation of the batch to a quarter (a large rho makes the inner problem ill-conditioned)
        if (worst > 0.25 * cgb_worst_prev) h->cgb_rho = std::min(h->cgb_rho * 4.0, 1e8);
        cgb_worst_prev = worst;
        }
        h->cgb_active = false;
        h->have_states = true;
        if (v_opt) TRY(download_soa(h, h->d_z.p, Bt, 0, v_opt, nz, Bt));
        if (J) TRY(d2h(h, J, h->d_J.p, sizeof(double) * Bt));
        if (iters) TRY(d2h(h, iters, h->d_sweeps, sizeof(int) * Bt));
        if (converged) TRY(d2h(h, converged, h->d_frozen, sizeof(int) * Bt));
        return finish_host(h);
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

int ocs_b200_reduce_J(ocs_b200_handle *h, int op, double *value, long *index)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_reduce_J(h, op, value, index);
    if (!h->have_states) return h->fail(OCS_B200_ERR_STATE, "reduce_J needs a preceding compute_states");
    if (!value || op < 0 || op > 2) return h->fail(OCS_B200_ERR_ARG, "bad reduction");
    try {
        CK(cudaSetDevice(h->device));
        const int nblk = 296, thr = 256;
        if (!h->d_red) CK(cudaMalloc(&h->d_red, (sizeof(double) + sizeof(long)) * (nblk + 1)));
        double *pv = h->d_red;
        long *pi = reinterpret_cast<long *>(h->d_red + nblk + 1);
#define RUN(OP)                                                                                            \
    reduce_kernel<OP><<<nblk, thr, 0, h->stream>>>(h->d_J.p, nullptr, h->Bt, pv, pi);                        \
    reduce_kernel<OP><<<1, thr, 0, h->stream>>>(pv, pi, nblk, pv + nblk, pi + nblk);
        if (op == OCS_B200_REDUCE_SUM) { RUN(OCS_B200_REDUCE_SUM) }
        else if (op == OCS_B200_REDUCE_MIN) { RUN(OCS_B200_REDUCE_MIN) }
        else { RUN(OCS_B200_REDUCE_MAX) }
#undef RUN
        CK(cudaGetLastError());
        long idx = 0;
        CK(cudaMemcpyAsync(value, pv + nblk, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(&idx, pi + nblk, sizeof(long), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (index) *index = op == OCS_B200_REDUCE_SUM ? -1 : idx;
        return 0;
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

int ocs_b200_set_profiling(ocs_b200_handle *h, int on)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_set_profiling(h, on);
    h->profiling = on != 0;
    return 0;
}

// ControlBounds at every grid point as linear constraints of ocs_b200_single_shooting (see cgb_penalty_kernel).
int ocs_b200_set_solve_control_bounds(ocs_b200_handle *h, const double *bounds, int nc, double rho0, double feas_tol, int max_outer)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_set_solve_control_bounds(h, bounds, nc, rho0, feas_tol, max_outer);
    if (!bounds) {
        h->cgb_on = false;
        return 0;
    }
    if (nc < 1 || nc > 64) return h->fail(OCS_B200_ERR_ARG, "bad control count");
    h->cgb_bounds.resize(2 * (size_t)nc);
    for (int c = 0; c < nc; ++c) {          // nc x 2 column-major -> [c][lo, hi]
        h->cgb_bounds[2 * c] = bounds[c];
        h->cgb_bounds[2 * c + 1] = bounds[c + nc];
        if (!(h->cgb_bounds[2 * c] <= h->cgb_bounds[2 * c + 1])) return h->fail(OCS_B200_ERR_ARG, "control bounds must satisfy lb <= ub");
    }
    h->cgb_rho0 = rho0 > 0 ? rho0 : 100.0;
    h->cgb_tol = feas_tol > 0 ? feas_tol : 1e-4;
    h->cgb_outer = max_outer > 0 ? max_outer : 12;
    h->cgb_on = true;
    return 0;
}

// Largest violation of the grid control bounds per trajectory after the last constrained solve: [Bt].
int ocs_b200_get_solve_violation(ocs_b200_handle *h, double *viol)
{
    if (!h || !viol) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_get_solve_violation(h, viol);
    if (!h->have_states || h->d_cgb_viol.cap < (size_t)h->Bt) return h->fail(OCS_B200_ERR_STATE, "no constrained solve on this handle");
    try {
        CK(cudaSetDevice(h->device));
        TRY(d2h(h, viol, h->d_cgb_viol.p, sizeof(double) * h->Bt));
        return finish_host(h);
    } catch (const std::exception &ex) {
        return h->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
    }
}

int ocs_b200_kernel_stats(ocs_b200_handle *h, int reset, double *ms_forward, double *ms_reverse, long *launches)
{
    if (!h) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_kernel_stats(h, reset, ms_forward, ms_reverse, launches);
    if (ms_forward) *ms_forward = h->ms_fwd;
    if (ms_reverse) *ms_reverse = h->ms_rev;
    if (launches) *launches = h->launches;
    if (reset) {
        h->ms_fwd = h->ms_rev = h->ms_proj = h->ms_upd = 0;
        h->launches = 0;
    }
    return 0;
}

int ocs_b200_kernel_stats_ex(ocs_b200_handle *h, int reset, double *ms, int nms, long *launches)
{
    if (!h || (nms > 0 && !ms)) return OCS_B200_ERR_ARG;
    if (h->multi) return ocs_multi_kernel_stats_ex(h, reset, ms, nms, launches);
    const double v[5] = {h->ms_fwd, h->ms_rev - h->ms_proj, h->ms_proj, h->ms_upd, h->ms_loop};
    for (int k = 0; k < nms; ++k) ms[k] = k < 5 ? v[k] : 0.0;
    if (launches) *launches = h->launches;
    if (reset) {
        h->ms_fwd = h->ms_rev = h->ms_proj = h->ms_upd = h->ms_loop = 0;
        h->launches = 0;
    }
    return 0;
}

int ocs_b200_get_dims(const ocs_b200_handle *h, int32_t dims[8])
{
    if (!h || !dims) return OCS_B200_ERR_ARG;
    if (h->multi) h = ocs_multi_first(h);
    dims[0] = h->have_problem ? h->kt.n : 0;
    dims[1] = h->have_problem ? h->kt.nc : 0;
    dims[2] = h->have_problem ? h->kt.np : 0;
    dims[3] = h->K;
    dims[4] = h->seg[0].N;
    dims[5] = h->seg[1].N;
    dims[6] = 0;
    dims[7] = 0;
    return 0;
}

long ocs_b200_get_batch(const ocs_b200_handle *h)
{
    if (!h) return -1;
    if (h->multi) return ocs_multi_batch(h);
    return h->have_states ? h->Bt : 0;
}

}  // extern "C"
