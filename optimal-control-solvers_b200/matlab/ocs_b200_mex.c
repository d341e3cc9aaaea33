/* ocs_b200_mex.c -- MEX gateway between the MATLAB classes in this directory and the C ABI of
 * include/ocs_b200.h (libocs_b200.so).
 *
 *     out = ocs_b200_mex(command, args...)
 *
 * One thin dispatcher: every command unpacks mxArrays into plain pointers, calls exactly one
 * ocs_b200_* entry point, and packs the outputs.  MATLAB owns every prhs (read-only) and every
 * plhs created here; the library never keeps a host pointer past a call.  A non-zero status is
 * turned into a MATLAB error (mexErrMsgIdAndTxt) AFTER temporaries are released; nothing
 * throws across the C ABI.  The handle travels as a uint64 scalar and is freed by
 * delete(obj) -> 'destroy' (CudaRK4Integrator.m).
 *
 * Lifecycle: the file locks itself (mexLock) while an integrator is alive and frees every remaining handle from a
 * mexAtExit hook; stale handle values are rejected.  Output sizes come from ocs_b200_get_dims / ocs_b200_get_batch
 * and input shapes are checked against them before the library is called.
 *
 * Build (MATLAB):  mex -I../../include ocs_b200_mex.c -L.. -locs_b200          (see build_mex.m)
 * Not compiled in this repository's image (no MATLAB / mex.h); tests/test_mex_gateway.py compiles it
 * against a stub mex.h and drives it end to end.
 */
#include <stdint.h>
#include <string.h>

#include "mex.h"
#include "ocs_b200.h"

#define FAIL_IF(cond, id, msg)                     \
    do {                                           \
        if (cond) mexErrMsgIdAndTxt(id, "%s", msg);\
    } while (0)

/* ---- handle registry and MEX lifecycle.
 * Every live integrator handle is recorded here.  While one exists the MEX file is locked (mexLock), so `clear mex` /
 * `clear all` cannot unload the library underneath MATLAB objects that still hold device memory; when MATLAB exits or
 * the file is unlocked and cleared, the mexAtExit hook frees whatever is left (the device-resident state grids are
 * tens of gigabytes).  A uint64 that is not in the registry -- a handle of a deleted object, a value from before a
 * `clear mex` -- is rejected instead of being dereferenced. */
#define OCS_MAX_HANDLES 256
static ocs_b200_handle *g_handles[OCS_MAX_HANDLES];
static int g_nhandles = 0;
static int g_atexit_set = 0;

static void ocs_at_exit(void)
{
    for (int k = 0; k < OCS_MAX_HANDLES; ++k)
        if (g_handles[k]) {
            ocs_b200_destroy(g_handles[k]);
            g_handles[k] = NULL;
        }
    if (g_nhandles > 0) mexUnlock();      /* the lock was held for the handles that are gone now */
    g_nhandles = 0;
}

static void registry_add(ocs_b200_handle *h)
{
    for (int k = 0; k < OCS_MAX_HANDLES; ++k)
        if (!g_handles[k]) {
            g_handles[k] = h;
            if (g_nhandles++ == 0) mexLock();
            return;
        }
    ocs_b200_destroy(h);
    mexErrMsgIdAndTxt("ocs_b200:handle", "too many live integrators (%d)", OCS_MAX_HANDLES);
}

static void registry_remove(ocs_b200_handle *h)
{
    for (int k = 0; k < OCS_MAX_HANDLES; ++k)
        if (g_handles[k] == h) {
            g_handles[k] = NULL;
            if (--g_nhandles == 0) mexUnlock();
            return;
        }
}

static ocs_b200_handle *get_handle(const mxArray *a)
{
    FAIL_IF(!mxIsUint64(a) || mxGetNumberOfElements(a) != 1, "ocs_b200:handle", "handle must be a uint64 scalar");
    uint64_t v = *(const uint64_t *)mxGetData(a);
    FAIL_IF(v == 0, "ocs_b200:handle", "integrator handle was already deleted");
    for (int k = 0; k < OCS_MAX_HANDLES; ++k)
        if (g_handles[k] && (uint64_t)(uintptr_t)g_handles[k] == v) return g_handles[k];
    mexErrMsgIdAndTxt("ocs_b200:handle", "not a live integrator handle (deleted, or from before `clear mex`)");
    return NULL;
}

/* Sizes come from the library (ocs_b200_get_dims / ocs_b200_get_batch), never from caller-supplied scalars: the C ABI
 * writes n-, N- and Bt-sized data, so an output sized from a stale MATLAB property would be a heap overflow. */
typedef struct { long n, nc, np, K, N, N2, Bt; } ocs_dims;
static ocs_dims dims_of(ocs_b200_handle *h)
{
    int32_t d[8];
    ocs_dims o;
    if (ocs_b200_get_dims(h, d) != 0) mexErrMsgIdAndTxt("ocs_b200:handle", "cannot read the integrator's dimensions");
    o.n = d[0]; o.nc = d[1]; o.np = d[2]; o.K = d[3]; o.N = d[4]; o.N2 = d[5];
    o.Bt = ocs_b200_get_batch(h);
    return o;
}
static void need_problem(const ocs_dims *d)
{
    FAIL_IF(d->n < 1, "ocs_b200:state", "no problem bound yet (set_problem_builtin / set_problem_source)");
}
static void need_batch(const ocs_dims *d)
{
    FAIL_IF(d->Bt < 1, "ocs_b200:state", "no resident batch: call compute_states first (RK4Integrator.m:8-10)");
}
static void need_rows(const mxArray *a, long rows, const char *what)
{
    if (!mxIsEmpty(a) && (long)mxGetM(a) != rows)
        mexErrMsgIdAndTxt("ocs_b200:size", "%s must have %ld rows, got %ld", what, rows, (long)mxGetM(a));
}
static void need_range(const ocs_dims *d, long b0, long nb)
{
    FAIL_IF(b0 < 0 || nb < 1 || b0 + nb > d->Bt, "ocs_b200:size", "trajectory range outside the resident batch");
}

static const double *dbl(const mxArray *a, const char *what)
{
    if (mxIsEmpty(a)) return NULL;
    if (!mxIsDouble(a) || mxIsComplex(a)) mexErrMsgIdAndTxt("ocs_b200:type", "%s must be a real double array", what);
    return mxGetPr(a);
}

static double scalar(const mxArray *a) { return mxGetScalar(a); }

/* pages along the last dimension: size(a, d) for d = ndims wanted (1 if absent) */
static long dim(const mxArray *a, int d)
{
    mwSize nd = mxGetNumberOfDimensions(a);
    const mwSize *s = mxGetDimensions(a);
    return d < (int)nd ? (long)s[d] : 1;
}

static void check(int rc, ocs_b200_handle *h)
{
    if (rc != 0) mexErrMsgIdAndTxt("ocs_b200:failed", "ocs_b200 error %d: %s", rc, ocs_b200_last_error(h));
}

static mxArray *int32_row(long n)
{
    mwSize d[2];
    d[0] = 1;
    d[1] = (mwSize)n;
    return mxCreateNumericArray(2, d, mxINT32_CLASS, mxREAL);
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    char cmd[64];
    FAIL_IF(nrhs < 1 || !mxIsChar(prhs[0]) || mxGetString(prhs[0], cmd, sizeof cmd) != 0, "ocs_b200:command",
            "first argument must be a command string");

    /* ---- h = mex('create', tspan, tspanExtra | [], device, math_mode)      RK4Integrator.m:16-25 */
    if (!strcmp(cmd, "create")) {
        FAIL_IF(nrhs < 5, "ocs_b200:nargin", "create(tspan, tspanExtra, device, math_mode)");
        const double *ts = dbl(prhs[1], "tspan"), *te = dbl(prhs[2], "tspanExtra");
        int N = (int)mxGetNumberOfElements(prhs[1]) - 1;
        int N2 = te ? (int)mxGetNumberOfElements(prhs[2]) - 1 : 0;
        ocs_b200_handle *h = NULL;
        int rc;
        FAIL_IF(N < 1 || (te && N2 < 1), "ocs_b200:size", "tspan / tspanExtra need at least two times");
        const size_t ndev = mxGetNumberOfElements(prhs[3]);
        FAIL_IF(ndev < 1 || ndev > 64 || !mxIsDouble(prhs[3]), "ocs_b200:create", "'Device' must be a CUDA ordinal or a vector of them");
        if (ndev == 1) {
            rc = ocs_b200_create(N, ts, N2, te, (int)scalar(prhs[3]), (int)scalar(prhs[4]), &h);
        } else {                      /* CudaRK4Integrator(tspan, 'Device', 0:7): one handle, the batch sharded over the GPUs */
            int devs[64];
            for (size_t k = 0; k < ndev; ++k) devs[k] = (int)mxGetPr(prhs[3])[k];
            rc = ocs_b200_create_multi(N, ts, N2, te, (int)ndev, devs, (int)scalar(prhs[4]), &h);
        }
        if (rc != 0) mexErrMsgIdAndTxt("ocs_b200:create", "ocs_b200 error %d: %s", rc, ocs_b200_last_error(NULL));
        if (!g_atexit_set) {
            mexAtExit(ocs_at_exit);
            g_atexit_set = 1;
        }
        registry_add(h);
        plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
        *(uint64_t *)mxGetData(plhs[0]) = (uint64_t)(uintptr_t)h;
        return;
    }
    FAIL_IF(nrhs < 2, "ocs_b200:nargin", "missing handle");
    ocs_b200_handle *h = get_handle(prhs[1]);

    if (!strcmp(cmd, "destroy")) {                                   /* delete(obj) */
        registry_remove(h);
        ocs_b200_destroy(h);
        return;
    }
    if (!strcmp(cmd, "get_dims")) {                                  /* [n nc np K N N2 Bt] = mex('get_dims', h) */
        const ocs_dims d = dims_of(h);
        plhs[0] = mxCreateDoubleMatrix(1, 7, mxREAL);
        double *o = mxGetPr(plhs[0]);
        o[0] = d.n; o[1] = d.nc; o[2] = d.np; o[3] = d.K; o[4] = d.N; o[5] = d.N2; o[6] = d.Bt;
        return;
    }
    if (!strcmp(cmd, "get_grid")) {                                  /* [t, h] = mex('get_grid', h, segment [, N]) */
        int seg = nrhs > 2 ? (int)scalar(prhs[2]) : 0;
        const ocs_dims dd = dims_of(h);
        FAIL_IF(seg < 0 || seg > 1 || (seg == 1 && dd.N2 == 0), "ocs_b200:size", "no such segment");
        const long N = seg == 0 ? dd.N : dd.N2;          /* the handle's own step count, not the caller's */
        plhs[0] = mxCreateDoubleMatrix(1, 2 * N + 1, mxREAL);
        mxArray *hh = mxCreateDoubleMatrix(1, N, mxREAL);
        int rc = ocs_b200_get_grid(h, seg, mxGetPr(plhs[0]), mxGetPr(hh));
        if (nlhs > 1) plhs[1] = hh; else mxDestroyArray(hh);
        check(rc, h);
        return;
    }
    if (!strcmp(cmd, "set_problem_builtin")) {
        FAIL_IF(nrhs < 3, "ocs_b200:nargin", "set_problem_builtin(h, id)");
        check(ocs_b200_set_problem_builtin(h, (int)scalar(prhs[2])), h);
        return;
    }
    if (!strcmp(cmd, "set_problem_source")) {                        /* (h, source, n, nc, np, include_dir) */
        FAIL_IF(nrhs < 6, "ocs_b200:nargin", "set_problem_source(h, source, n, nc, np [, include_dir])");
        char *src = mxArrayToString(prhs[2]);
        char *inc = (nrhs > 6 && !mxIsEmpty(prhs[6])) ? mxArrayToString(prhs[6]) : NULL;
        int rc = ocs_b200_set_problem_source(h, src, (int)scalar(prhs[3]), (int)scalar(prhs[4]), (int)scalar(prhs[5]), inc);
        mxFree(src);
        if (inc) mxFree(inc);
        check(rc, h);
        return;
    }
    if (!strcmp(cmd, "set_params")) {                                /* (h, params [np x nbatch]) */
        FAIL_IF(nrhs < 3, "ocs_b200:nargin", "set_params(h, params)");
        { const ocs_dims d = dims_of(h); need_problem(&d); need_rows(prhs[2], d.np, "params"); }
        check(ocs_b200_set_params(h, dbl(prhs[2], "params"), (int)mxGetM(prhs[2]), (long)mxGetN(prhs[2])), h);
        return;
    }
    if (!strcmp(cmd, "set_tail_control")) {
        FAIL_IF(nrhs < 3, "ocs_b200:nargin", "set_tail_control(h, uStar)");
        { const ocs_dims d = dims_of(h); need_problem(&d);
          FAIL_IF((long)mxGetNumberOfElements(prhs[2]) != d.nc, "ocs_b200:size", "uStar must have nControls entries"); }
        check(ocs_b200_set_tail_control(h, dbl(prhs[2], "uStar"), (int)mxGetNumberOfElements(prhs[2])), h);
        return;
    }
    if (!strcmp(cmd, "set_control_basis")) {                         /* (h, B [K x G]) ; [] detaches */
        FAIL_IF(nrhs < 3, "ocs_b200:nargin", "set_control_basis(h, B)");
        { const ocs_dims d = dims_of(h);
          FAIL_IF(!mxIsEmpty(prhs[2]) && (long)mxGetN(prhs[2]) != 2 * d.N + 1, "ocs_b200:size",
                  "B must have one column per point of integrator.t (2*nSTEPS+1)"); }
        check(ocs_b200_set_control_basis(h, (int)mxGetM(prhs[2]), dbl(prhs[2], "B")), h);
        return;
    }
    if (!strcmp(cmd, "set_control_v")) {                             /* (h, v [(nc*K) x nbatch]) */
        FAIL_IF(nrhs < 3, "ocs_b200:nargin", "set_control_v(h, v)");
        { const ocs_dims d = dims_of(h); need_problem(&d);
          FAIL_IF(d.K < 1, "ocs_b200:state", "attach a control basis before setting v");
          need_rows(prhs[2], d.nc * d.K, "v"); }
        check(ocs_b200_set_control_v(h, dbl(prhs[2], "v"), (long)mxGetN(prhs[2])), h);
        return;
    }
    if (!strcmp(cmd, "set_control_u")) {                             /* (h, u [nc x G x nbatch]) */
        FAIL_IF(nrhs < 3, "ocs_b200:nargin", "set_control_u(h, u)");
        { const ocs_dims d = dims_of(h); need_problem(&d); need_rows(prhs[2], d.nc, "u");
          FAIL_IF(dim(prhs[2], 1) != 2 * d.N + 1, "ocs_b200:size", "u must have one column per point of integrator.t"); }
        check(ocs_b200_set_control_u(h, dbl(prhs[2], "u"), dim(prhs[2], 2)), h);
        return;
    }
    /* ---- [J, xT, status] = mex('compute_states', h, x0 [n x Bt])            RK4Integrator.m:28-56 */
    if (!strcmp(cmd, "compute_states")) {
        FAIL_IF(nrhs < 3, "ocs_b200:nargin", "compute_states(h, x0)");
        const ocs_dims d = dims_of(h);
        need_problem(&d);
        need_rows(prhs[2], d.n, "x0");
        long n = d.n, Bt = (long)mxGetN(prhs[2]);
        FAIL_IF(Bt < 1, "ocs_b200:size", "x0 is empty");
        plhs[0] = mxCreateDoubleMatrix(1, Bt, mxREAL);
        mxArray *xT = mxCreateDoubleMatrix(n, Bt, mxREAL);
        mxArray *st = int32_row(Bt);
        int rc = ocs_b200_compute_states(h, dbl(prhs[2], "x0"), Bt, mxGetPr(plhs[0]), mxGetPr(xT), (int32_t *)mxGetData(st));
        if (nlhs > 1) plhs[1] = xT; else mxDestroyArray(xT);
        if (nlhs > 2) plhs[2] = st; else mxDestroyArray(st);
        check(rc, h);
        return;
    }
    /* ---- [lam0, dJdv] = mex('compute_adjoints', h, lamT | [], m, Bt, nv)    RK4Integrator.m:59-121 */
    if (!strcmp(cmd, "compute_adjoints")) {
        FAIL_IF(nrhs < 3, "ocs_b200:nargin", "compute_adjoints(h, lamT [, m, Bt, nv])");
        const double *lamT = dbl(prhs[2], "lamT");
        const ocs_dims d = dims_of(h);
        need_problem(&d);
        need_batch(&d);
        need_rows(prhs[2], d.n + 1, "lamT");
        /* the trailing scalars of older callers only say whether a gradient is wanted (nv > 0); sizes are the handle's */
        long m = d.n + 1, Bt = d.Bt, nv = (nrhs > 5 && scalar(prhs[5]) <= 0) ? 0 : d.nc * d.K;
        plhs[0] = mxCreateDoubleMatrix(m, Bt, mxREAL);
        mxArray *dJdv = (nlhs > 1 && nv > 0) ? mxCreateDoubleMatrix(nv, Bt, mxREAL) : NULL;
        int rc = ocs_b200_compute_adjoints(h, lamT, lamT ? (long)mxGetN(prhs[2]) : 0, dJdv ? mxGetPr(dJdv) : NULL,
                                           mxGetPr(plhs[0]));
        if (nlhs > 1) plhs[1] = dJdv ? dJdv : mxCreateDoubleMatrix(0, 0, mxREAL);
        check(rc, h);
        return;
    }
    /* ---- [J, xT, status, lam0, dJdv] = mex('evaluate', h, x0, lamT | [], nv)   nlpObjective, single_shooting.m:137-150:
     *      compute_states + compute_adjoints in one call (chunked, gradient download overlapped) */
    if (!strcmp(cmd, "evaluate")) {
        FAIL_IF(nrhs < 4, "ocs_b200:nargin", "evaluate(h, x0, lamT [, nv])");
        const double *lamT = dbl(prhs[3], "lamT");
        const ocs_dims d = dims_of(h);
        need_problem(&d);
        need_rows(prhs[2], d.n, "x0");
        need_rows(prhs[3], d.n + 1, "lamT");
        long n = d.n, Bt = (long)mxGetN(prhs[2]), nv = (nrhs > 4 && scalar(prhs[4]) <= 0) ? 0 : d.nc * d.K;
        FAIL_IF(Bt < 1, "ocs_b200:size", "x0 is empty");
        plhs[0] = mxCreateDoubleMatrix(1, Bt, mxREAL);
        mxArray *xT = mxCreateDoubleMatrix(n, Bt, mxREAL), *st = int32_row(Bt);
        mxArray *l0 = mxCreateDoubleMatrix(n + 1, Bt, mxREAL), *g = mxCreateDoubleMatrix(nv, Bt, mxREAL);
        int rc = ocs_b200_evaluate(h, dbl(prhs[2], "x0"), Bt, lamT, lamT ? (long)mxGetN(prhs[3]) : 0, mxGetPr(plhs[0]),
                                   mxGetPr(xT), (int32_t *)mxGetData(st), nv > 0 ? mxGetPr(g) : NULL, mxGetPr(l0));
        if (nlhs > 1) plhs[1] = xT; else mxDestroyArray(xT);
        if (nlhs > 2) plhs[2] = st; else mxDestroyArray(st);
        if (nlhs > 3) plhs[3] = l0; else mxDestroyArray(l0);
        if (nlhs > 4) plhs[4] = g; else mxDestroyArray(g);
        check(rc, h);
        return;
    }
    /* ---- x = mex('fetch_states', h, segment, b0, nb, m, N)                  x of RK4Integrator.m:54 */
    if (!strcmp(cmd, "fetch_states")) {
        FAIL_IF(nrhs < 5, "ocs_b200:nargin", "fetch_states(h, segment, b0, nb)");
        const ocs_dims dd = dims_of(h);
        need_problem(&dd);
        need_batch(&dd);
        const int seg = (int)scalar(prhs[2]);
        FAIL_IF(seg < 0 || seg > 1 || (seg == 1 && dd.N2 == 0), "ocs_b200:size", "no such segment");
        need_range(&dd, (long)scalar(prhs[3]), (long)scalar(prhs[4]));
        mwSize d[3];
        d[0] = (mwSize)(dd.n + 1);
        d[1] = (mwSize)((seg == 0 ? dd.N : dd.N2) + 1);
        d[2] = (mwSize)scalar(prhs[4]);
        plhs[0] = mxCreateNumericArray(3, d, mxDOUBLE_CLASS, mxREAL);
        check(ocs_b200_fetch_states(h, (int)scalar(prhs[2]), (long)scalar(prhs[3]), (long)scalar(prhs[4]), mxGetPr(plhs[0])), h);
        return;
    }
    /* ---- [lam, dJdu] = mex('fetch_adjoints', h, segment, b0, nb, m, N, nc) */
    if (!strcmp(cmd, "fetch_adjoints")) {
        FAIL_IF(nrhs < 5, "ocs_b200:nargin", "fetch_adjoints(h, segment, b0, nb)");
        const ocs_dims dd = dims_of(h);
        need_problem(&dd);
        need_batch(&dd);
        const int seg = (int)scalar(prhs[2]);
        FAIL_IF(seg < 0 || seg > 1 || (seg == 1 && dd.N2 == 0), "ocs_b200:size", "no such segment");
        need_range(&dd, (long)scalar(prhs[3]), (long)scalar(prhs[4]));
        mwSize d[3];
        long N = seg == 0 ? dd.N : dd.N2;
        d[0] = (mwSize)(dd.n + 1);
        d[1] = (mwSize)N + 1;
        d[2] = (mwSize)scalar(prhs[4]);
        plhs[0] = mxCreateNumericArray(3, d, mxDOUBLE_CLASS, mxREAL);
        mxArray *du = NULL;
        if (nlhs > 1) {
            d[0] = (mwSize)dd.nc;
            d[1] = (mwSize)(2 * N + 1);
            du = mxCreateNumericArray(3, d, mxDOUBLE_CLASS, mxREAL);
            plhs[1] = du;
        }
        check(ocs_b200_fetch_adjoints(h, (int)scalar(prhs[2]), (long)scalar(prhs[3]), (long)scalar(prhs[4]),
                                      mxGetPr(plhs[0]), du ? mxGetPr(du) : NULL), h);
        return;
    }
    /* ---- u = mex('get_control_u', h, b0, nb, nc, G) */
    if (!strcmp(cmd, "get_control_u")) {
        FAIL_IF(nrhs < 4, "ocs_b200:nargin", "get_control_u(h, b0, nb)");
        const ocs_dims dd = dims_of(h);
        need_problem(&dd);
        FAIL_IF(scalar(prhs[3]) < 1, "ocs_b200:size", "nb must be positive");
        mwSize d[3];
        d[0] = (mwSize)dd.nc;
        d[1] = (mwSize)(2 * dd.N + 1);
        d[2] = (mwSize)scalar(prhs[3]);
        plhs[0] = mxCreateNumericArray(3, d, mxDOUBLE_CLASS, mxREAL);
        check(ocs_b200_get_control_u(h, (long)scalar(prhs[2]), (long)scalar(prhs[3]), mxGetPr(plhs[0])), h);
        return;
    }
    /* ---- [J, sweeps, converged, maxChange] = mex('fb_sweep', h, x0, u0, bounds, uRelTol, uAbsTol, nSWEEPS, relax[, costate])
     *      functions/fb_sweep.m:79-115 on the RK4 grid; costate 0 = discrete adjoint (default), 1 = the v1 API's
     *      adjointRHS integrated backwards (functions/compute_x_lam.m:11-12) */
    if (!strcmp(cmd, "fb_sweep")) {
        FAIL_IF(nrhs < 9, "ocs_b200:nargin", "fb_sweep(h, x0, u0, bounds, uRelTol, uAbsTol, nSWEEPS, relax[, costate])");
        check(ocs_b200_set_fb_costate(h, nrhs > 9 ? (int)scalar(prhs[9]) : OCS_B200_COSTATE_DISCRETE), h);
        const ocs_dims d = dims_of(h);
        need_problem(&d);
        need_rows(prhs[2], d.n, "x0");
        need_rows(prhs[3], d.nc, "u0");
        FAIL_IF(!mxIsEmpty(prhs[3]) && dim(prhs[3], 1) != 2 * d.N + 1, "ocs_b200:size", "u0 must have one column per point of integrator.t");
        FAIL_IF((long)mxGetNumberOfElements(prhs[4]) != 2 * d.nc, "ocs_b200:size", "bounds must be nControls x 2");
        long Bt = (long)mxGetN(prhs[2]);
        FAIL_IF(Bt < 1, "ocs_b200:size", "x0 is empty");
        plhs[0] = mxCreateDoubleMatrix(1, Bt, mxREAL);
        mxArray *sw = int32_row(Bt), *cv = int32_row(Bt), *mc = mxCreateDoubleMatrix(1, Bt, mxREAL);
        int rc = ocs_b200_fb_sweep(h, dbl(prhs[2], "x0"), Bt, dbl(prhs[3], "u0"), dim(prhs[3], 2), dbl(prhs[4], "bounds"),
                                   scalar(prhs[5]), scalar(prhs[6]), (int)scalar(prhs[7]), scalar(prhs[8]),
                                   mxGetPr(plhs[0]), (int32_t *)mxGetData(sw), (int32_t *)mxGetData(cv), mxGetPr(mc));
        if (nlhs > 1) plhs[1] = sw; else mxDestroyArray(sw);
        if (nlhs > 2) plhs[2] = cv; else mxDestroyArray(cv);
        if (nlhs > 3) plhs[3] = mc; else mxDestroyArray(mc);
        check(rc, h);
        return;
    }
    /* ---- mex('set_control_ufunc', h, kind, knots): how control.compute_uFunc interpolates v
     *      (1 'previous' over intervalStarts, 2 'linear' over controlPts, 3 Chebyshev over [t(1) t(end)]) */
    if (!strcmp(cmd, "set_control_ufunc")) {
        FAIL_IF(nrhs < 4, "ocs_b200:nargin", "set_control_ufunc(h, kind, knots)");
        check(ocs_b200_set_control_ufunc(h, (int)scalar(prhs[2]), (int)mxGetNumberOfElements(prhs[3]), dbl(prhs[3], "knots")), h);
        return;
    }
    /* ---- [x, lam, u] = mex('eval_solution', h, segment, b0, nb, tq, n, nc): soln.x(tq), soln.lam(tq), soln.u(tq) of
     *      functions/single_shooting.m:128-130 for trajectories b0 .. b0+nb-1; only the requested outputs are computed */
    if (!strcmp(cmd, "eval_solution")) {
        FAIL_IF(nrhs < 6, "ocs_b200:nargin", "eval_solution(h, segment, b0, nb, tq)");
        long nb = (long)scalar(prhs[4]);
        int nq = (int)mxGetNumberOfElements(prhs[5]);
        const ocs_dims dd = dims_of(h);
        need_problem(&dd);
        need_batch(&dd);
        need_range(&dd, (long)scalar(prhs[3]), nb);
        FAIL_IF(nq < 1, "ocs_b200:size", "no query times");
        mwSize d[3];
        d[1] = (mwSize)nq;
        d[2] = (mwSize)nb;
        d[0] = (mwSize)dd.n;
        plhs[0] = mxCreateNumericArray(3, d, mxDOUBLE_CLASS, mxREAL);
        mxArray *lam = nlhs > 1 ? mxCreateNumericArray(3, d, mxDOUBLE_CLASS, mxREAL) : NULL;
        d[0] = (mwSize)dd.nc;
        mxArray *uu = nlhs > 2 ? mxCreateNumericArray(3, d, mxDOUBLE_CLASS, mxREAL) : NULL;
        int rc = ocs_b200_eval_solution(h, (int)scalar(prhs[2]), (long)scalar(prhs[3]), nb, nq, dbl(prhs[5], "tq"),
                                        mxGetPr(plhs[0]), lam ? mxGetPr(lam) : NULL, uu ? mxGetPr(uu) : NULL);
        if (lam) plhs[1] = lam;
        if (uu) plhs[2] = uu;
        check(rc, h);
        return;
    }
    /* ---- [y, resnorm, residual, exitflag] = mex('compute_equilibrium', h, params, guess, lb, ub, r, maxIter, tol)
     *      functions/compute_equilibrium.m:1-33, one steady state per column of params */
    if (!strcmp(cmd, "compute_equilibrium")) {
        FAIL_IF(nrhs < 9, "ocs_b200:nargin", "compute_equilibrium(h, params, guess, lb, ub, r, maxIter, tol)");
        const ocs_dims dd = dims_of(h);
        need_problem(&dd);
        need_rows(prhs[2], dd.np, "params");
        need_rows(prhs[3], 2 * dd.n + dd.nc, "guess");
        FAIL_IF((!mxIsEmpty(prhs[4]) && (long)mxGetNumberOfElements(prhs[4]) != 2 * dd.n + dd.nc) ||
                (!mxIsEmpty(prhs[5]) && (long)mxGetNumberOfElements(prhs[5]) != 2 * dd.n + dd.nc), "ocs_b200:size",
                "lb / ub must have 2*nStates + nControls entries");
        long nsets = (long)mxGetN(prhs[2]);
        long ny = 2 * dd.n + dd.nc;
        if (nsets < 1) nsets = 1;
        if ((long)mxGetN(prhs[3]) > nsets) nsets = (long)mxGetN(prhs[3]);
        if ((long)mxGetNumberOfElements(prhs[6]) > nsets) nsets = (long)mxGetNumberOfElements(prhs[6]);
        FAIL_IF(mxGetN(prhs[2]) > 1 && (long)mxGetN(prhs[2]) != nsets, "ocs_b200:size", "params must have one column per set");
        plhs[0] = mxCreateDoubleMatrix(ny, nsets, mxREAL);
        mxArray *rn = mxCreateDoubleMatrix(1, nsets, mxREAL), *rs = mxCreateDoubleMatrix(ny, nsets, mxREAL), *fl = int32_row(nsets);
        int rc = ocs_b200_compute_equilibrium(h, dbl(prhs[2], "params"), (int)mxGetM(prhs[2]),
                                              mxGetN(prhs[2]) > 1 ? (long)mxGetN(prhs[2]) : 1, nsets, dbl(prhs[3], "guess"), (long)mxGetN(prhs[3]),
                                              dbl(prhs[4], "lb"), dbl(prhs[5], "ub"), dbl(prhs[6], "r"),
                                              (long)mxGetNumberOfElements(prhs[6]), (int)scalar(prhs[7]), scalar(prhs[8]),
                                              mxGetPr(plhs[0]), mxGetPr(rn), mxGetPr(rs), (int32_t *)mxGetData(fl), NULL);
        if (nlhs > 1) plhs[1] = rn; else mxDestroyArray(rn);
        if (nlhs > 2) plhs[2] = rs; else mxDestroyArray(rs);
        if (nlhs > 3) plhs[3] = fl; else mxDestroyArray(fl);
        check(rc, h);
        return;
    }
    /* ---- [v, J, iters, converged] = mex('single_shooting', h, x0, v0, Lb, Ub, free (0-based int32 | []), freeLb, freeUb, maxIter, tol)
     *      the fmincon solve of functions/single_shooting.m:114 for every column of x0, on the device */
    if (!strcmp(cmd, "single_shooting")) {
        FAIL_IF(nrhs < 11, "ocs_b200:nargin", "single_shooting(h, x0, v0, Lb, Ub, free, freeLb, freeUb, maxIter, tol)");
        const ocs_dims d = dims_of(h);
        need_problem(&d);
        need_rows(prhs[2], d.n, "x0");
        long Bt = (long)mxGetN(prhs[2]);
        long nv = d.nc * d.K;
        int nfree = (int)mxGetNumberOfElements(prhs[6]);
        FAIL_IF(Bt < 1 || d.K < 1, "ocs_b200:size", "x0 is empty or no control basis is attached");
        need_rows(prhs[3], nv, "v0");
        FAIL_IF((long)mxGetNumberOfElements(prhs[4]) != nv || (long)mxGetNumberOfElements(prhs[5]) != nv, "ocs_b200:size",
                "Lb / Ub must have nControls*K entries (control.compute_nlp_bounds)");
        FAIL_IF((long)mxGetNumberOfElements(prhs[7]) != nfree || (long)mxGetNumberOfElements(prhs[8]) != nfree, "ocs_b200:size",
                "FreeStateBounds must have one row per free state");
        FAIL_IF(nfree > 0 && mxGetClassID(prhs[6]) != mxINT32_CLASS, "ocs_b200:type", "free state indices must be int32 (0-based)");
        plhs[0] = mxCreateDoubleMatrix(nv + nfree, Bt, mxREAL);
        mxArray *Jo = mxCreateDoubleMatrix(1, Bt, mxREAL), *it = int32_row(Bt), *cv = int32_row(Bt);
        int rc = ocs_b200_single_shooting(h, dbl(prhs[2], "x0"), Bt, dbl(prhs[3], "v0"), (long)mxGetN(prhs[3]),
                                          dbl(prhs[4], "Lb"), dbl(prhs[5], "Ub"), nfree,
                                          nfree ? (const int32_t *)mxGetData(prhs[6]) : NULL, dbl(prhs[7], "freeLb"),
                                          dbl(prhs[8], "freeUb"), (int)scalar(prhs[9]), scalar(prhs[10]),
                                          mxGetPr(plhs[0]), mxGetPr(Jo), (int32_t *)mxGetData(it), (int32_t *)mxGetData(cv));
        if (nlhs > 1) plhs[1] = Jo; else mxDestroyArray(Jo);
        if (nlhs > 2) plhs[2] = it; else mxDestroyArray(it);
        if (nlhs > 3) plhs[3] = cv; else mxDestroyArray(cv);
        check(rc, h);
        return;
    }
    /* ---- mex('set_solve_control_bounds', h, ControlBounds | [] [, rho0, feasTol, maxOuter])
     *      [A, b] = control.compute_lincon(prob.ControlBounds) of functions/single_shooting.m:106-111 for a dense
     *      (Chebyshev) basis: ControlBounds (nc x 2) at every grid point; [] switches the constraints off */
    if (!strcmp(cmd, "set_solve_control_bounds")) {
        FAIL_IF(nrhs < 3, "ocs_b200:nargin", "set_solve_control_bounds(h, ControlBounds [, rho0, feasTol, maxOuter])");
        const double *cb = dbl(prhs[2], "ControlBounds");
        FAIL_IF(cb && mxGetN(prhs[2]) != 2, "ocs_b200:size", "ControlBounds must be nc x 2");
        check(ocs_b200_set_solve_control_bounds(h, cb, cb ? (int)mxGetM(prhs[2]) : 0, nrhs > 3 ? scalar(prhs[3]) : 0.0,
                                                nrhs > 4 ? scalar(prhs[4]) : 0.0, nrhs > 5 ? (int)scalar(prhs[5]) : 0), h);
        return;
    }
    /* ---- viol = mex('get_solve_violation', h)   largest bound violation per trajectory of the last constrained solve */
    if (!strcmp(cmd, "get_solve_violation")) {
        const ocs_dims dd = dims_of(h);
        need_batch(&dd);
        plhs[0] = mxCreateDoubleMatrix(1, (mwSize)dd.Bt, mxREAL);
        check(ocs_b200_get_solve_violation(h, mxGetPr(plhs[0])), h);
        return;
    }
    /* ---- [value, index] = mex('reduce_J', h, op)   op: 0 sum, 1 min, 2 max (index 1-based) */
    if (!strcmp(cmd, "reduce_J")) {
        FAIL_IF(nrhs < 3, "ocs_b200:nargin", "reduce_J(h, op)");
        double val = 0;
        long idx = -1;
        check(ocs_b200_reduce_J(h, (int)scalar(prhs[2]), &val, &idx), h);
        plhs[0] = mxCreateDoubleScalar(val);
        if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(idx >= 0 ? (double)(idx + 1) : 0.0);
        return;
    }
    mexErrMsgIdAndTxt("ocs_b200:command", "unknown command '%s'", cmd);
}
