// problem_test.cuh -- device form of the reference's bundled test problem
// (tests/TestOCProblem.m:13-38):   F = [ x (m - x) - u ;  exp(-r t) (x^2 + c u^2) ]
// params p = {c, m, r}.  Operation order is the MATLAB source's, left to right.
//
// Hoisted table (see HM_* in rk4_kernels.cuh).  Batch-uniform sub-expressions are
// evaluated once per grid point with the very operations (and order) of the in-thread
// form, so hoisting does not change a single bit of the result:
//   HM_TIME          tt[0] = exp(-r t)        (F :25)
//                    tt[1] = exp(r t)/(2c)    (ControlChar: u* = lam * (exp(r t) / (2c)))
//                    tt[2] = 2*exp(-r t)      (dFdx_times_vec :31, left-to-right prefix)
//                    tt[3] = 2*c*exp(-r t)    (dFdu_times_vec :37, left-to-right prefix of 2*c*exp(-r*t).*u)
//   HM_TIME_CONTROL  tt[0] = exp(-r t)
//                    tt[1] = c*u.^2           (F :25)
//                    tt[2] = 2*exp(-r t)      (dFdx_times_vec :31, left-to-right prefix)
//                    tt[3] = 2*c*exp(-r t).*u (dFdu_times_vec :37, left-to-right prefix)
//
// Weight folding (fast arithmetic only; see ocs_nfold in rk4_kernels.cuh).  The objective row enters every
// right-hand side through a factor that is batch-uniform once the RK4 stage weight is known: the forward
// quadrature adds (h/6, h/3, h/3, h/6) * E * (x^2 + c u^2), and in the adjoint sweep the objective row of each
// stage costate dJdk is the stage weight itself when lam(end) = 1 (RK4Integrator.m:63-66,73-85: the last row of
// dFdx_times_vec is zero, so lam(end,:) never changes).  fold() multiplies the table entries that meet that
// weight by it, once per step and stage position; Ff / Xf / Uf are F / dFdx_times_vec / dFdu_times_vec with
// the products folded in.  This re-associates E*(..)*w into (E*w)*(..): results move by rounding only (held
// to the 1e-10 parity bound like the FMA contraction of this mode); the strict build does not fold.
//   fw[0] = exp(-r t) * w      fw[1] = 2*exp(-r t) * w      fw[2] = tt[3] * w  (2c exp(-r t) [u] w)
#pragma once

namespace OCS_NS {

struct TestOCProblem {
    static constexpr int n = 1, nc = 1, np = 3;
    __host__ __device__ static constexpr int ntt(int hm) { return hm == 0 ? 0 : 4; }

    template <int HM>
    __device__ __forceinline__ static void table(double t, const double *u, const double *p, double *tt)
    {
        if constexpr (HM == 1) {
            const double E = exp(-p[2] * t);
            tt[0] = E;
            tt[1] = exp(p[2] * t) / (2.0 * p[0]);
            tt[2] = 2.0 * E;
            tt[3] = 2.0 * p[0] * E;
        } else if constexpr (HM == 2) {
            const double E = exp(-p[2] * t);
            tt[0] = E;
            tt[1] = p[0] * (u[0] * u[0]);
            tt[2] = 2.0 * E;
            tt[3] = 2.0 * p[0] * E * u[0];
        }
    }

    // Per-trajectory parameters: exp(-r t) advances along the grid by recurrence (ocs_ntexp in rk4_kernels.cuh) and
    // the right-hand sides run in their HM_TIME form on a per-trajectory table with the entries of table<1> (tt[1],
    // the ControlChar factor, is not needed by F / dFdx_times_vec / dFdu_times_vec).
    static constexpr int ntexp = 1;
    __device__ __forceinline__ static void texp_rates(const double *p, double *a) { a[0] = -p[2]; }
    __device__ __forceinline__ static void table_from_texp(const double *E, double t, const double *p, double *tt)
    {
        (void)t;
        tt[0] = E[0];
        tt[1] = 0.0;
        tt[2] = 2.0 * E[0];
        tt[3] = 2.0 * p[0] * E[0];
    }

    static constexpr int nfold = 3;
    template <int HM>
    __device__ __forceinline__ static void fold(const double *tt, double w, double *fw)
    {
        fw[0] = tt[0] * w;
        fw[1] = tt[2] * w;
        fw[2] = tt[3] * w;
    }

    // state rows of F into o[0..n-1]; the objective row, times the stage weight, is added to J
    template <int HM>
    __device__ __forceinline__ static void Ff(const double *tt, const double *fw, double t, const double *y,
                                              const double *u, const double *p, double *o, double &J)
    {
        (void)t;
        const double cu2 = HM == 2 ? tt[1] : p[0] * (u[0] * u[0]);
        const double x = y[0];
        o[0] = x * (p[1] - x) - u[0];
        J = fma(fw[0], x * x + cu2, J);
    }

    // dFdx_times_vec(t, y, u, [v; w]) with the stage weight w folded into fw
    template <int HM>
    __device__ __forceinline__ static void Xf(const double *tt, const double *fw, double t, const double *y,
                                              const double *u, const double *p, const double *v, double *o)
    {
        (void)tt; (void)t; (void)u;
        const double x = y[0];
        o[0] = fma(fma(-2.0, x, p[1]), v[0], fw[1] * x);
    }

    // dFdu_times_vec(t, y, u, [v; w])
    template <int HM>
    __device__ __forceinline__ static void Uf(const double *tt, const double *fw, double t, const double *y,
                                              const double *u, const double *p, const double *v, double *o)
    {
        (void)tt; (void)t; (void)y; (void)p;
        o[0] = HM == 2 ? fw[2] - v[0] : fma(fw[2], u[0], -v[0]);
    }

    // value = [x.*(m - x) - u; exp(-r*t).*(x.^2 + c*u.^2)]          TestOCProblem.m:22-26
    template <int HM>
    __device__ __forceinline__ static void F(const double *tt, double t, const double *y, const double *u,
                                             const double *p, double *o)
    {
        const double E = HM ? tt[0] : exp(-p[2] * t);
        const double cu2 = HM == 2 ? tt[1] : p[0] * (u[0] * u[0]);
        const double x = y[0];
        o[0] = x * (p[1] - x) - u[0];
        o[1] = E * (x * x + cu2);
    }

    // value = [(m - 2*x).*v(1) + 2*exp(-r*t).*x.*v(2); 0]           TestOCProblem.m:29-33
    // (the zero last row is implicit: o has n entries)
    template <int HM>
    __device__ __forceinline__ static void X(const double *tt, double t, const double *y, const double *u,
                                             const double *p, const double *v, double *o)
    {
        (void)u;
        const double E2 = HM ? tt[2] : 2.0 * exp(-p[2] * t);
        const double x = y[0];
        // m - 2*x: the product 2*x is exact, so the fused form rounds identically to MATLAB's two operations
        o[0] = fma(-2.0, x, p[1]) * v[0] + E2 * x * v[1];
    }

    // value = -v(1) + 2*c*exp(-r*t).*u.*v(2)                        TestOCProblem.m:36-38
    template <int HM>
    __device__ __forceinline__ static void U(const double *tt, double t, const double *y, const double *u,
                                             const double *p, const double *v, double *o)
    {
        (void)y;
        const double CEU = HM == 2 ? tt[3] : (HM == 1 ? tt[3] * u[0] : 2.0 * p[0] * exp(-p[2] * t) * u[0]);
        o[0] = -v[0] + CEU * v[1];
    }

    // u solving dFdu_times_vec(t, y, u, [lam; 1]) = 0, clamped to ControlBounds
    // (functions/make_from_symbolic.m:17-23,111):  u* = lam*exp(r*t)/(2*c), written as lam * (exp(r*t)/(2*c)) so
    // that the batch-uniform factor is one table entry (the reference's form is whatever `solve` prints)
    template <int HM>
    __device__ __forceinline__ static void ControlChar(const double *tt, double t, const double *x,
                                                       const double *lam, const double *p,
                                                       const double *bounds, double *o)
    {
        (void)x;
        const double Ei2c = HM == 1 ? tt[1] : exp(p[2] * t) / (2.0 * p[0]);
        const double uu = lam[0] * Ei2c;
        o[0] = fmin(bounds[1], fmax(bounds[0], uu));
    }
};

}  // namespace OCS_NS
