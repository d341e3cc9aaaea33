/* Drives mexFunction of ocs_b200_mex.c the way CudaRK4Integrator.m does -- TEST INFRASTRUCTURE.
 *   mex_driver <in.bin> <out.bin>
 * in.bin : int32 N, int32 nc, then tspan[N+1], params[3], x0, u[nc*(2N+1)]   (doubles)
 * out.bin: J, x[2*(N+1)], lam[2*(N+1)], dJdu[nc*(2N+1)]
 * Exit code 3 + message on a MATLAB-style error (mexErrMsgIdAndTxt). */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mex.h"

static jmp_buf g_jmp;
static char g_err[2048];

/* ---- stub mx API */
static size_t elsize(mxClassID c) { return c == mxDOUBLE_CLASS ? 8 : c == mxINT32_CLASS ? 4 : c == mxUINT64_CLASS ? 8 : 1; }
int mxIsUint64(const mxArray *a) { return a->cls == mxUINT64_CLASS; }
int mxIsDouble(const mxArray *a) { return a->cls == mxDOUBLE_CLASS; }
int mxIsChar(const mxArray *a) { return a->cls == mxCHAR_CLASS; }
int mxIsComplex(const mxArray *a) { (void)a; return 0; }
mxClassID mxGetClassID(const mxArray *a) { return a->cls; }
size_t mxGetNumberOfElements(const mxArray *a) { size_t n = 1; for (mwSize k = 0; k < a->ndim; ++k) n *= a->dims[k]; return n; }
int mxIsEmpty(const mxArray *a) { return mxGetNumberOfElements(a) == 0; }
mwSize mxGetNumberOfDimensions(const mxArray *a) { return a->ndim; }
const mwSize *mxGetDimensions(const mxArray *a) { return a->dims; }
size_t mxGetM(const mxArray *a) { return a->dims[0]; }
size_t mxGetN(const mxArray *a) { size_t n = 1; for (mwSize k = 1; k < a->ndim; ++k) n *= a->dims[k]; return n; }
void *mxGetData(const mxArray *a) { return a->data; }
double *mxGetPr(const mxArray *a) { return (double *)a->data; }
double mxGetScalar(const mxArray *a)
{
    if (a->cls == mxDOUBLE_CLASS) return ((double *)a->data)[0];
    if (a->cls == mxINT32_CLASS) return ((int32_t *)a->data)[0];
    return (double)((uint64_t *)a->data)[0];
}
int mxGetString(const mxArray *a, char *buf, mwSize buflen)
{
    size_t n = mxGetNumberOfElements(a);
    if (n + 1 > buflen) return 1;
    memcpy(buf, a->data, n);
    buf[n] = 0;
    return 0;
}
char *mxArrayToString(const mxArray *a)
{
    size_t n = mxGetNumberOfElements(a);
    char *s = malloc(n + 1);
    memcpy(s, a->data, n);
    s[n] = 0;
    return s;
}
void mxFree(void *p) { free(p); }
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c)
{
    (void)c;
    mxArray *a = calloc(1, sizeof *a);
    a->cls = cls;
    a->ndim = ndim < 2 ? 2 : ndim;
    a->dims[0] = a->dims[1] = a->dims[2] = a->dims[3] = 1;
    for (mwSize k = 0; k < ndim; ++k) a->dims[k] = dims[k];
    a->data = calloc(mxGetNumberOfElements(a) + 1, elsize(cls));
    return a;
}
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c)
{
    mwSize d[2] = {m, n};
    return mxCreateNumericArray(2, d, cls, c);
}
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) { return mxCreateNumericMatrix(m, n, mxDOUBLE_CLASS, c); }
mxArray *mxCreateDoubleScalar(double v)
{
    mxArray *a = mxCreateDoubleMatrix(1, 1, mxREAL);
    ((double *)a->data)[0] = v;
    return a;
}
mxArray *mxCreateString(const char *s)
{
    mwSize d[2] = {1, strlen(s)};
    mxArray *a = mxCreateNumericArray(2, d, mxCHAR_CLASS, mxREAL);
    memcpy(a->data, s, strlen(s));
    return a;
}
void mxDestroyArray(mxArray *a) { if (a) { free(a->data); free(a); } }
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    int k = snprintf(g_err, sizeof g_err, "[%s] ", id);
    vsnprintf(g_err + k, sizeof g_err - k, fmt, ap);
    va_end(ap);
    longjmp(g_jmp, 1);
}

/* ---- MEX lifecycle: lock count and the at-exit hook MATLAB would run on `clear mex` / exit */
static int g_lock = 0;
static void (*g_atexit)(void) = NULL;
void mexLock(void) { ++g_lock; }
void mexUnlock(void) { if (g_lock > 0) --g_lock; }
int mexIsLocked(void) { return g_lock > 0; }
int mexAtExit(void (*fn)(void)) { g_atexit = fn; return 0; }

/* ---- helpers */
static mxArray *dmat(mwSize m, mwSize n, const double *src)
{
    mxArray *a = mxCreateDoubleMatrix(m, n, mxREAL);
    if (src) memcpy(a->data, src, sizeof(double) * m * n);
    return a;
}
static mxArray *call(const char *cmd, int nlhs, mxArray **out, int nargs, ...)
{
    const mxArray *in[12];
    in[0] = mxCreateString(cmd);
    va_list ap;
    va_start(ap, nargs);
    for (int k = 0; k < nargs; ++k) in[k + 1] = va_arg(ap, const mxArray *);
    va_end(ap);
    mxArray *lhs[4] = {0, 0, 0, 0};
    mexFunction(nlhs, lhs, nargs + 1, in);
    for (int k = 0; k < nlhs && out; ++k) out[k] = lhs[k];
    return lhs[0];
}

/* expects the call to raise a MATLAB error whose text contains `needle`; returns 1 if it did */
static int g_expect_active = 0;
static jmp_buf g_expect_jmp;
#define EXPECT_ERROR(needle, stmt)                                                            \
    do {                                                                                      \
        jmp_buf saved;                                                                        \
        memcpy(saved, g_jmp, sizeof saved);                                                   \
        int raised = 0;                                                                       \
        if (setjmp(g_jmp) == 0) { stmt; } else raised = 1;                                    \
        memcpy(g_jmp, saved, sizeof saved);                                                   \
        if (!raised || !strstr(g_err, needle)) {                                              \
            fprintf(stderr, "expected an error containing '%s', got '%s'\n", needle, raised ? g_err : "(none)"); \
            return 4;                                                                         \
        }                                                                                     \
    } while (0)

/* mex_driver --lifecycle <in.bin>: locking, the at-exit hook with live handles, stale handles, shape checks */
static int lifecycle(int N, const double *tspan, const double *params)
{
    (void)g_expect_active; (void)g_expect_jmp;
    mxArray *empty = mxCreateDoubleMatrix(0, 0, mxREAL);
    if (mexIsLocked()) return 5;
    mxArray *h1 = call("create", 1, NULL, 4, dmat(1, N + 1, tspan), empty, mxCreateDoubleScalar(0), mxCreateDoubleScalar(0));
    mxArray *h2 = call("create", 1, NULL, 4, dmat(1, N + 1, tspan), empty, mxCreateDoubleScalar(0), mxCreateDoubleScalar(0));
    if (!mexIsLocked() || !g_atexit) return 6;              /* live integrators pin the MEX file */
    call("set_problem_builtin", 0, NULL, 2, h1, mxCreateDoubleScalar(0));
    call("set_params", 0, NULL, 2, h1, dmat(3, 1, params));
    /* sizes are the library's, and wrong shapes are refused before the C ABI sees them */
    mxArray *d = call("get_dims", 1, NULL, 1, h1);
    if (mxGetPr(d)[0] != 1 || mxGetPr(d)[1] != 1 || mxGetPr(d)[2] != 3 || mxGetPr(d)[4] != N || mxGetPr(d)[6] != 0) return 7;
    double two[2] = {1.0, 2.0};
    EXPECT_ERROR("x0 must have 1 rows", call("compute_states", 1, NULL, 2, h1, dmat(2, 1, two)));
    EXPECT_ERROR("params must have 3 rows", call("set_params", 0, NULL, 2, h1, dmat(2, 1, two)));
    EXPECT_ERROR("no resident batch", call("fetch_states", 1, NULL, 4, h1, mxCreateDoubleScalar(0), mxCreateDoubleScalar(0), mxCreateDoubleScalar(1)));
    EXPECT_ERROR("one column per point", call("set_control_u", 0, NULL, 2, h1, dmat(1, 5, NULL)));
    call("destroy", 0, NULL, 1, h1);
    if (!mexIsLocked()) return 8;                            /* h2 is still alive */
    EXPECT_ERROR("not a live integrator handle", call("get_dims", 1, NULL, 1, h1));
    /* `clear mex` / MATLAB exit with a live handle: the hook frees it, later use of the value is refused */
    g_atexit();
    EXPECT_ERROR("not a live integrator handle", call("get_dims", 1, NULL, 1, h2));
    mxArray *bogus = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
    *(uint64_t *)mxGetData(bogus) = 0xdeadbeefULL;
    EXPECT_ERROR("not a live integrator handle", call("get_dims", 1, NULL, 1, bogus));
    /* a fresh integrator after the hook works and locks again */
    mxArray *h3 = call("create", 1, NULL, 4, dmat(1, N + 1, tspan), empty, mxCreateDoubleScalar(0), mxCreateDoubleScalar(0));
    if (!mexIsLocked()) return 9;
    call("destroy", 0, NULL, 1, h3);
    if (mexIsLocked()) return 10;
    printf("lifecycle ok\n");
    return 0;
}

int main(int argc, char **argv)
{
    int life = argc >= 3 && !strcmp(argv[1], "--lifecycle");
    if (life) { argv[1] = argv[2]; }
    if (argc < 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t N, nc;
    if (fread(&N, 4, 1, f) != 1 || fread(&nc, 4, 1, f) != 1) return 2;
    const int G = 2 * N + 1;
    double *tspan = malloc(8 * (N + 1)), params[3], x0, *u = malloc(8 * nc * G);
    if (fread(tspan, 8, N + 1, f) != (size_t)(N + 1) || fread(params, 8, 3, f) != 3 || fread(&x0, 8, 1, f) != 1 ||
        fread(u, 8, nc * G, f) != (size_t)(nc * G)) return 2;
    fclose(f);
    if (setjmp(g_jmp)) {
        fprintf(stderr, "MATLAB error: %s\n", g_err);
        return 3;
    }
    if (life) return lifecycle(N, tspan, params);
    mxArray *empty = mxCreateDoubleMatrix(0, 0, mxREAL);
    mxArray *h = call("create", 1, NULL, 4, dmat(1, N + 1, tspan), empty, mxCreateDoubleScalar(0), mxCreateDoubleScalar(0));
    call("set_problem_builtin", 0, NULL, 2, h, mxCreateDoubleScalar(0));
    call("set_params", 0, NULL, 2, h, dmat(3, 1, params));
    call("set_control_u", 0, NULL, 2, h, dmat(nc, G, u));
    mxArray *o[4];
    call("compute_states", 3, o, 2, h, dmat(1, 1, &x0));
    double J = mxGetPr(o[0])[0];
    mxArray *x = call("fetch_states", 1, NULL, 6, h, mxCreateDoubleScalar(0), mxCreateDoubleScalar(0), mxCreateDoubleScalar(1),
                      mxCreateDoubleScalar(2), mxCreateDoubleScalar(N));
    call("compute_adjoints", 2, o, 5, h, empty, mxCreateDoubleScalar(2), mxCreateDoubleScalar(1), mxCreateDoubleScalar(0));
    mxArray *g[2];
    call("fetch_adjoints", 2, g, 7, h, mxCreateDoubleScalar(0), mxCreateDoubleScalar(0), mxCreateDoubleScalar(1),
         mxCreateDoubleScalar(2), mxCreateDoubleScalar(N), mxCreateDoubleScalar(nc));
    call("destroy", 0, NULL, 1, h);
    f = fopen(argv[2], "wb");
    fwrite(&J, 8, 1, f);
    fwrite(mxGetPr(x), 8, 2 * (N + 1), f);
    fwrite(mxGetPr(g[0]), 8, 2 * (N + 1), f);
    fwrite(mxGetPr(g[1]), 8, nc * G, f);
    fclose(f);
    printf("mex gateway ok J=%.17g\n", J);
    return 0;
}
