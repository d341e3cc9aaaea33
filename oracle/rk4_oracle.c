/* CPU oracle (C restatement) for the RK4 trajectory hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 * PARITY PINNING: pinned against golden vectors produced by executing the reference's
 * unmodified .m files under oracle/mlab (the .npz files in tests/golden, tests/test_oracle_golden.py);
 * see the header of oracle/rk4_oracle.py for what that does and does not cover (no
 * MATLAB/Octave in the image; MATLAB's built-ins are restated).
 *
 * Plain C, one trajectory at a time, the exact operation order of
 *   Integrator/RK4Integrator.m:28-56   (compute_states)
 *   Integrator/RK4Integrator.m:59-94   (compute_adjoints)
 *   Integrator/RK4Integrator.m:97-121  (compute_dJdu)
 *   Integrator/RK4InfiniteIntegrator.m:20-30
 *   Control/{PWConstant,PWLinear,Chebyshev}Control.m  compute_u (v*B) / compute_dJdv (dJdu*B')
 *   tests/TestOCProblem.m:22-38
 * Build with -ffp-contract=off so that no FMA is formed (MATLAB's element-wise
 * arithmetic rounds every operation).  OpenMP only parallelises ACROSS
 * trajectories; each trajectory runs the reference's serial loop.
 *
 * Array conventions are MATLAB's (column-major): xK is m x (N+1) x 4,
 * u is nc x G, B is K x G.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int n;        /* true states */
    int nc;       /* controls    */
    const double *p;
    void (*F)(const double *p, double t, const double *y, const double *u, double *out);
    void (*X)(const double *p, double t, const double *y, const double *u, const double *v, double *out);
    void (*U)(const double *p, double t, const double *y, const double *u, const double *v, double *out);
} oc_problem;

/* ---- tests/TestOCProblem.m, p = {c, m, r} ------------------------------------ */
static void test_F(const double *p, double t, const double *y, const double *u, double *o)
{   /* :22-26 */
    double x = y[0];
    o[0] = x * (p[1] - x) - u[0];
    o[1] = exp(-p[2] * t) * (x * x + p[0] * (u[0] * u[0]));
}
static void test_X(const double *p, double t, const double *y, const double *u, const double *v, double *o)
{   /* :29-33 */
    (void)u;
    double x = y[0];
    o[0] = (p[1] - 2 * x) * v[0] + 2 * exp(-p[2] * t) * x * v[1];
    o[1] = 0;
}
static void test_U(const double *p, double t, const double *y, const double *u, const double *v, double *o)
{   /* :36-38 */
    (void)y;
    o[0] = -v[0] + 2 * p[0] * exp(-p[2] * t) * u[0] * v[1];
}

#define MAXM 16
#define XK(c, i, k) xK[(c) + m * ((i) + (size_t)(N + 1) * (k))]
#define DK(c, i, k) dJdk[(c) + m * ((i) + (size_t)N * (k))]

/* RK4Integrator.m:28-56.  u is nc x G for this trajectory. */
static void compute_states(const oc_problem *pr, int N, const double *t, const double *h,
                           const double *x0, const double *u, double *xK)
{
    int n = pr->n, m = n + 1, nc = pr->nc;
    double F1[MAXM], F2[MAXM], F3[MAXM], F4[MAXM], y[MAXM];
    for (int c = 0; c < n; ++c) XK(c, 0, 0) = x0[c];
    XK(n, 0, 0) = 0;
    for (int i = 0; i < N; ++i) {
        double hi = h[i];
        for (int c = 0; c < m; ++c) y[c] = XK(c, i, 0);
        pr->F(pr->p, t[2 * i], y, u + nc * (2 * i), F1);
        for (int c = 0; c < m; ++c) y[c] = XK(c, i, 1) = XK(c, i, 0) + hi / 2 * F1[c];
        pr->F(pr->p, t[2 * i + 1], y, u + nc * (2 * i + 1), F2);
        for (int c = 0; c < m; ++c) y[c] = XK(c, i, 2) = XK(c, i, 0) + hi / 2 * F2[c];
        pr->F(pr->p, t[2 * i + 1], y, u + nc * (2 * i + 1), F3);
        for (int c = 0; c < m; ++c) y[c] = XK(c, i, 3) = XK(c, i, 0) + hi * F3[c];
        pr->F(pr->p, t[2 * i + 2], y, u + nc * (2 * i + 2), F4);
        for (int c = 0; c < m; ++c)
            XK(c, i + 1, 0) = XK(c, i, 0) + hi / 6 * (F1[c] + 2 * F2[c] + 2 * F3[c] + F4[c]);
    }
}

/* RK4Integrator.m:59-94.  lam is m x (N+1), dJdk is m x N x 4. */
static void compute_adjoints(const oc_problem *pr, int N, const double *t, const double *h,
                             const double *u, const double *xK, const double *lamT,
                             double *lam, double *dJdk)
{
    int n = pr->n, m = n + 1, nc = pr->nc;
    double g1[MAXM], g2[MAXM], g3[MAXM], g0[MAXM], y[MAXM], v[MAXM];
    for (int c = 0; c < m; ++c) lam[c + (size_t)m * N] = lamT ? lamT[c] : (c == n ? 1.0 : 0.0);
    for (int i = N - 1; i >= 0; --i) {
        double hi = h[i];
        const double *ln = lam + (size_t)m * (i + 1);
        for (int c = 0; c < m; ++c) { v[c] = DK(c, i, 3) = hi / 6 * ln[c]; y[c] = XK(c, i, 3); }
        pr->X(pr->p, t[2 * i + 2], y, u + nc * (2 * i + 2), v, g3);
        for (int c = 0; c < m; ++c) { v[c] = DK(c, i, 2) = hi / 3 * ln[c] + hi * g3[c]; y[c] = XK(c, i, 2); }
        pr->X(pr->p, t[2 * i + 1], y, u + nc * (2 * i + 1), v, g2);
        for (int c = 0; c < m; ++c) { v[c] = DK(c, i, 1) = hi / 3 * ln[c] + hi / 2 * g2[c]; y[c] = XK(c, i, 1); }
        pr->X(pr->p, t[2 * i + 1], y, u + nc * (2 * i + 1), v, g1);
        for (int c = 0; c < m; ++c) { v[c] = DK(c, i, 0) = hi / 6 * ln[c] + hi / 2 * g1[c]; y[c] = XK(c, i, 0); }
        pr->X(pr->p, t[2 * i], y, u + nc * (2 * i), v, g0);
        for (int c = 0; c < m; ++c) lam[c + (size_t)m * i] = ln[c] + g1[c] + g2[c] + g3[c] + g0[c];
    }
}

/* RK4Integrator.m:97-121.  dJdu is nc x G. */
static void compute_dJdu(const oc_problem *pr, int N, const double *t, const double *u,
                         const double *xK, const double *dJdk, double *dJdu)
{
    int n = pr->n, m = n + 1, nc = pr->nc, G = 2 * N + 1;
    double a[MAXM], b[MAXM], y[MAXM], v[MAXM];
#define LOADYV(i, k) for (int c = 0; c < m; ++c) { y[c] = XK(c, i, k); v[c] = DK(c, i, k); }
    LOADYV(0, 0); pr->U(pr->p, t[0], y, u, v, a);                               /* :101-102 */
    for (int c = 0; c < nc; ++c) dJdu[c] = a[c];
    for (int i = 0; i < N; ++i) {                                                 /* :105-109 */
        int j = 2 * i + 1;
        LOADYV(i, 1); pr->U(pr->p, t[j], y, u + nc * j, v, a);
        LOADYV(i, 2); pr->U(pr->p, t[j], y, u + nc * j, v, b);
        for (int c = 0; c < nc; ++c) dJdu[c + nc * j] = a[c] + b[c];
    }
    for (int i = 1; i < N; ++i) {                                                 /* :112-116 */
        int j = 2 * i;
        LOADYV(i, 0); pr->U(pr->p, t[j], y, u + nc * j, v, a);
        LOADYV(i - 1, 3); pr->U(pr->p, t[j], y, u + nc * j, v, b);
        for (int c = 0; c < nc; ++c) dJdu[c + nc * j] = a[c] + b[c];
    }
    LOADYV(N - 1, 3); pr->U(pr->p, t[G - 1], y, u + nc * (G - 1), v, a);        /* :119-120 */
    for (int c = 0; c < nc; ++c) dJdu[c + nc * (G - 1)] = a[c];
#undef LOADYV
}

/* u = reshape(v,nc,[])*B  (PWLinearControl.m:59-62);  B is K x G.  klo/khi bracket the
 * non-zero rows of column j (terms with B(k,j)==0 add an exact zero and are skipped). */
static void compute_u(int nc, int K, int G, const double *v, const double *B,
                      const int *klo, const int *khi, double *u)
{
    for (int j = 0; j < G; ++j)
        for (int c = 0; c < nc; ++c) {
            double s = 0;
            for (int k = klo[j]; k <= khi[j]; ++k) {
                double b = B[k + (size_t)K * j];
                if (b != 0) s += v[c + nc * k] * b;
            }
            u[c + nc * j] = s;
        }
}
/* dJdv = reshape(dJdu*B',[],1)  (PWLinearControl.m:53-56) */
static void compute_dJdv(int nc, int K, int G, const double *dJdu, const double *B,
                         const int *klo, const int *khi, double *dJdv)
{
    for (int k = 0; k < nc * K; ++k) dJdv[k] = 0;
    for (int j = 0; j < G; ++j)
        for (int k = klo[j]; k <= khi[j]; ++k) {
            double b = B[k + (size_t)K * j];
            if (b != 0)
                for (int c = 0; c < nc; ++c) dJdv[c + nc * k] += dJdu[c + nc * j] * b;
        }
}

static int make_problem(int problem_id, const double *params, oc_problem *pr)
{
    if (problem_id != 0) return -1;
    pr->n = 1; pr->nc = 1; pr->p = params;
    pr->F = test_F; pr->X = test_X; pr->U = test_U;
    return 0;
}

/* One batched objective+gradient evaluation = single_shooting.m:137-150 for every
 * trajectory b (and the RK4InfiniteIntegrator chaining when N2 > 0).
 *   x0      n x Bt
 *   v       (nc*K) x (v_batched ? Bt : 1)
 *   params  np x (p_batched ? Bt : 1)
 * outputs (any may be NULL):
 *   J [Bt], xT n x Bt, dJdv (nc*K) x Bt, lam0 m x Bt
 *   x_grid m x (N+1) x Bt, lam_grid m x (N+1) x Bt, dJdu_grid nc x G x Bt
 * mode: 0 = forward only, 1 = forward + adjoint.
 * returns 0 on success. */
int oracle_evaluate(int problem_id, const double *params, int np, int p_batched,
                    int N, const double *tspan,
                    int N2, const double *tspan_extra, const double *uStar,
                    int K, const double *B,
                    const double *x0, const double *v, int v_batched, long Bt, int mode,
                    double *J, double *xT, double *dJdv, double *lam0,
                    double *x_grid, double *lam_grid, double *dJdu_grid, int nthreads)
{
    oc_problem proto;
    if (make_problem(problem_id, params, &proto)) return -1;
    const int n = proto.n, m = n + 1, nc = proto.nc, G = 2 * N + 1, G2 = 2 * N2 + 1;
    /* RK4Integrator.m:16-25 */
    double *h = malloc(sizeof(double) * N), *t = malloc(sizeof(double) * G);
    for (int i = 0; i < N; ++i) h[i] = tspan[i + 1] - tspan[i];
    for (int i = 0; i <= N; ++i) t[2 * i] = tspan[i];
    for (int i = 0; i < N; ++i) t[2 * i + 1] = (tspan[i] + tspan[i + 1]) / 2;
    double *h2 = NULL, *t2 = NULL, *u2 = NULL;
    if (N2 > 0) {
        h2 = malloc(sizeof(double) * N2); t2 = malloc(sizeof(double) * G2);
        u2 = malloc(sizeof(double) * nc * G2);
        for (int i = 0; i < N2; ++i) h2[i] = tspan_extra[i + 1] - tspan_extra[i];
        for (int i = 0; i <= N2; ++i) t2[2 * i] = tspan_extra[i];
        for (int i = 0; i < N2; ++i) t2[2 * i + 1] = (tspan_extra[i] + tspan_extra[i + 1]) / 2;
        for (int j = 0; j < G2; ++j) for (int c = 0; c < nc; ++c) u2[c + nc * j] = uStar[c]; /* Inf:15 */
    }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    int *klo = malloc(sizeof(int) * G), *khi = malloc(sizeof(int) * G);
    for (int j = 0; j < G; ++j) {
        klo[j] = K; khi[j] = -1;
        for (int k = 0; k < K; ++k)
            if (B[k + (size_t)K * j] != 0) { if (k < klo[j]) klo[j] = k; khi[j] = k; }
    }
    int fail = 0;
#pragma omp parallel
    {
        oc_problem pr = proto;
        double *xK = malloc(sizeof(double) * m * (N + 1) * 4);
        double *dJdk = malloc(sizeof(double) * m * N * 4);
        double *lam = malloc(sizeof(double) * m * (N + 1));
        double *u = malloc(sizeof(double) * nc * G);
        double *dJdu = malloc(sizeof(double) * nc * G);
        double *xK2 = N2 > 0 ? malloc(sizeof(double) * m * (N2 + 1) * 4) : NULL;
        double *dJdk2 = N2 > 0 ? malloc(sizeof(double) * m * N2 * 4) : NULL;
        double *lam2 = N2 > 0 ? malloc(sizeof(double) * m * (N2 + 1)) : NULL;
        if (!xK || !dJdk || !lam || !u || !dJdu) {
#pragma omp atomic write
            fail = 1;
        } else {
        if (!v_batched) compute_u(nc, K, G, v, B, klo, khi, u);
#pragma omp for schedule(static)
        for (long b = 0; b < Bt; ++b) {
            if (p_batched) pr.p = params + (size_t)np * b;
            if (v_batched) compute_u(nc, K, G, v + (size_t)nc * K * b, B, klo, khi, u);
            compute_states(&pr, N, t, h, x0 + (size_t)n * b, u, xK);
            double Jb = xK[n + m * (size_t)N];                         /* :55 */
            const double *lamT = NULL;
            if (N2 > 0) {                                              /* Inf:21-23 */
                double xe[MAXM];
                for (int c = 0; c < n; ++c) xe[c] = xK[c + m * (size_t)N];
                compute_states(&pr, N2, t2, h2, xe, u2, xK2);
                Jb = Jb + xK2[n + m * (size_t)N2];
            }
            if (J) J[b] = Jb;
            if (xT) for (int c = 0; c < n; ++c) xT[c + (size_t)n * b] = xK[c + m * (size_t)N];
            if (x_grid)
                for (int i = 0; i <= N; ++i) for (int c = 0; c < m; ++c)
                    x_grid[c + m * (i + (size_t)(N + 1) * b)] = xK[c + m * (size_t)i];
            if (mode == 0) continue;
            if (N2 > 0) {                                              /* Inf:28-29 */
                compute_adjoints(&pr, N2, t2, h2, u2, xK2, NULL, lam2, dJdk2);
                lamT = lam2;
            }
            compute_adjoints(&pr, N, t, h, u, xK, lamT, lam, dJdk);
            compute_dJdu(&pr, N, t, u, xK, dJdk, dJdu);
            if (dJdv) compute_dJdv(nc, K, G, dJdu, B, klo, khi, dJdv + (size_t)nc * K * b);
            if (lam0) for (int c = 0; c < m; ++c) lam0[c + (size_t)m * b] = lam[c];
            if (lam_grid) memcpy(lam_grid + (size_t)m * (N + 1) * b, lam, sizeof(double) * m * (N + 1));
            if (dJdu_grid) memcpy(dJdu_grid + (size_t)nc * G * b, dJdu, sizeof(double) * nc * G);
        }
        }
        free(xK); free(dJdk); free(lam); free(u); free(dJdu); free(xK2); free(dJdk2); free(lam2);
    }
    free(h); free(t); free(h2); free(t2); free(u2); free(klo); free(khi);
    return fail;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
