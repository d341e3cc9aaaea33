// problem_test.cuh -- device form of the reference's bundled test problem
// (tests/TestOCProblem.m:13-38):   F = [ x (m - x) - u ;  exp(-r t) (x^2 + c u^2) ]
// params p = {c, m, r}.  Operation order is the MATLAB source's, left to right.
//
// Time table (used when the parameters are shared by the batch): the sub-expressions
// that depend only on (t, r) are evaluated once per grid point,
//   tt[0] = exp(-r t)   (F :25, dFdx_times_vec :31, dFdu_times_vec :37)
//   tt[1] = exp(r t)    (ControlChar: u* = lam exp(r t) / (2c), the root of dFdu_times_vec = 0)
#pragma once

namespace OCS_NS {

struct TestOCProblem {
    static constexpr int n = 1, nc = 1, np = 3, ntt = 2;

    __device__ __forceinline__ static void time_table(double t, const double *p, double *tt)
    {
        tt[0] = exp(-p[2] * t);
        tt[1] = exp(p[2] * t);
    }

    // value = [x.*(m - x) - u; exp(-r*t).*(x.^2 + c*u.^2)]          TestOCProblem.m:22-26
    template <bool TT>
    __device__ __forceinline__ static void F(const double *tt, double t, const double *y, const double *u,
                                             const double *p, double *o)
    {
        const double E = TT ? tt[0] : exp(-p[2] * t);
        const double x = y[0];
        o[0] = x * (p[1] - x) - u[0];
        o[1] = E * (x * x + p[0] * (u[0] * u[0]));
    }

    // value = [(m - 2*x).*v(1) + 2*exp(-r*t).*x.*v(2); 0]           TestOCProblem.m:29-33
    // (the zero last row is implicit: o has n entries)
    template <bool TT>
    __device__ __forceinline__ static void X(const double *tt, double t, const double *y, const double *u,
                                             const double *p, const double *v, double *o)
    {
        (void)u;
        const double E = TT ? tt[0] : exp(-p[2] * t);
        const double x = y[0];
        o[0] = (p[1] - 2.0 * x) * v[0] + 2.0 * E * x * v[1];
    }

    // value = -v(1) + 2*c*exp(-r*t).*u.*v(2)                        TestOCProblem.m:36-38
    template <bool TT>
    __device__ __forceinline__ static void U(const double *tt, double t, const double *y, const double *u,
                                             const double *p, const double *v, double *o)
    {
        (void)y;
        const double E = TT ? tt[0] : exp(-p[2] * t);
        o[0] = -v[0] + 2.0 * p[0] * E * u[0] * v[1];
    }

    // u solving dFdu_times_vec(t, y, u, [lam; 1]) = 0, clamped to ControlBounds
    // (functions/make_from_symbolic.m:17-23,111):  u* = lam*exp(r*t)/(2*c)
    template <bool TT>
    __device__ __forceinline__ static void ControlChar(const double *tt, double t, const double *x,
                                                       const double *lam, const double *p,
                                                       const double *bounds, double *o)
    {
        (void)x;
        const double Ei = TT ? tt[1] : exp(p[2] * t);
        const double uu = lam[0] * Ei / (2.0 * p[0]);
        o[0] = fmin(bounds[1], fmax(bounds[0], uu));
    }
};

}  // namespace OCS_NS
