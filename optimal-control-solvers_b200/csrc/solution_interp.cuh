// solution_interp.cuh -- the `soln` interpolants of the reference evaluated on the device for a batch.
//
//   soln.x   = vectorInterpolant(tspan, xOpt(1:end-1,:), 'pchip')      functions/single_shooting.m:129
//   soln.lam = vectorInterpolant(tspan, lamOpt(1:end-1,:), 'pchip')    functions/single_shooting.m:130
//   soln.u   = control.compute_uFunc(vOpt)                             functions/single_shooting.m:128
//              PWLinearControl.m:74-77 ('linear'), PWConstantControl.m:58-61 ('previous')
//   soln.u   = vectorInterpolant(interpPts, ControlChar(...), 'pchip') functions/fb_sweep.m:123
//
// vectorInterpolant.m:1-12 builds one griddedInterpolant per row; 'pchip' is MATLAB's shape-preserving
// piecewise cubic Hermite interpolant (Fritsch & Carlson 1980; slopes as in MATLAB's pchip.m: weighted
// harmonic mean of the neighbouring secants at interior points where they agree in sign, zero otherwise,
// and the non-centred three-point formula with the shape-preserving clamps at the two ends).
//
// Everything that depends only on the query times (interval index, local coordinate, neighbouring step
// sizes) is batch-uniform: interp_prepare_kernel tabulates it once per query, and the evaluation kernels
// read the source arrays in their SoA layout [row][point][batch] so that a warp's 32 trajectories touch
// 256 contiguous bytes per load.  Output is SoA [(row*nq + q)][batch]; the caller transposes it to the
// MATLAB layout rows x nq x nb.
#pragma once

namespace ocs_interp {

// per-query record: [k (as double) | s = tq - x_k | h_{k-1} | h_k | h_{k+1}]   (h = 0 where no neighbour exists)
constexpr int QREC = 6;

__global__ void interp_prepare_kernel(const double *__restrict__ xs, int NP, const double *__restrict__ tq, int nq,
                                      double *__restrict__ qrec)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const double t = tq[q];
    // k = last index with xs[k] <= t, clamped to [0, NP-2] (queries outside the grid use the end polynomials)
    int lo = 0, hi = NP - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (xs[mid] <= t) lo = mid; else hi = mid;
    }
    const int k = lo;
    double *r = qrec + (size_t)q * QREC;
    r[0] = (double)k;
    r[1] = t - xs[k];
    r[2] = k > 0 ? xs[k] - xs[k - 1] : 0.0;
    r[3] = xs[k + 1] - xs[k];
    r[4] = k + 2 < NP ? xs[k + 2] - xs[k + 1] : 0.0;
    r[5] = 0.0;
}

__device__ __forceinline__ double sgn(double a) { return (a > 0.0) - (a < 0.0); }

// interior slope: secants dl (over hl) and dr (over hr) on either side of the point
__device__ __forceinline__ double pchip_interior(double hl, double hr, double dl, double dr)
{
    if (!(sgn(dl) * sgn(dr) > 0.0)) return 0.0;
    const double hs = hl + hr;
    const double w1 = (hl + hs) / (3.0 * hs), w2 = (hs + hr) / (3.0 * hs);
    const double al = fabs(dl), ar = fabs(dr);
    const double dmax = fmax(al, ar), dmin = fmin(al, ar);
    return dmin / (w1 * (dl / dmax) + w2 * (dr / dmax));
}

// end slope: h0, d0 belong to the end interval, h1, d1 to its neighbour
__device__ __forceinline__ double pchip_end(double h0, double h1, double d0, double d1)
{
    double d = ((2.0 * h0 + h1) * d0 - h0 * d1) / (h0 + h1);
    if (sgn(d) != sgn(d0)) d = 0.0;
    else if (sgn(d0) != sgn(d1) && fabs(d) > fabs(3.0 * d0)) d = 3.0 * d0;
    return d;
}

// src: SoA [rows][NP][ld]; out: SoA [(row*nq + q)][ldo].  Thread = one trajectory, blockIdx.y = a chunk of queries.
__global__ void __launch_bounds__(128) pchip_eval_kernel(const double *__restrict__ src, long ld, int NP, int rows,
                                                         const double *__restrict__ qrec, int nq, int qchunk,
                                                         double *__restrict__ out, long ldo, long nb)
{
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const int q0 = blockIdx.y * qchunk, q1 = min(nq, q0 + qchunk);
    for (int q = q0; q < q1; ++q) {
        const double *r = qrec + (size_t)q * QREC;
        const int k = (int)r[0];
        const double s = r[1], hl = r[2], h = r[3], hr = r[4];
        for (int c = 0; c < rows; ++c) {
            const double *y = src + ((size_t)c * NP + k) * ld + b;
            const double y0 = y[0], y1 = y[ld];
            const double del = (y1 - y0) / h;
            double d0, d1;
            if (NP == 2) {
                d0 = d1 = del;
            } else {
                const double dl = k > 0 ? (y0 - y[-ld]) / hl : 0.0;
                const double dr = k + 2 < NP ? (y[2 * ld] - y1) / hr : 0.0;
                d0 = k > 0 ? pchip_interior(hl, h, dl, del) : pchip_end(h, hr, del, dr);
                d1 = k + 2 < NP ? pchip_interior(h, hr, del, dr) : pchip_end(h, hl, del, dl);
            }
            const double cc = (3.0 * del - 2.0 * d0 - d1) / h, bb = (d0 - 2.0 * del + d1) / (h * h);
            out[((size_t)c * nq + q) * ldo + b] = y0 + s * (d0 + s * (cc + s * bb));
        }
    }
}

enum { UFUNC_NONE = 0, UFUNC_PREVIOUS = 1, UFUNC_LINEAR = 2, UFUNC_CHEBYSHEV = 3 };

// control.compute_uFunc(v)(tq) for every trajectory.  v: SoA [(k*nc + c)][ldv] (vstride = 0: one shared column).
// knots: intervalStarts (PREVIOUS, K), controlPts (LINEAR, K), [t(1), t(end)] (CHEBYSHEV).
__global__ void __launch_bounds__(128) ufunc_eval_kernel(int kind, const double *__restrict__ v, long ldv, int vbatched, int K,
                                                         int nc, const double *__restrict__ knots,
                                                         const double *__restrict__ tq, int nq, int qchunk,
                                                         double *__restrict__ out, long ldo, long nb)
{
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const long bv = vbatched ? b : 0;
    const int q0 = blockIdx.y * qchunk, q1 = min(nq, q0 + qchunk);
    for (int q = q0; q < q1; ++q) {
        const double t = tq[q];
        if (kind == UFUNC_CHEBYSHEV) {
            // the recurrence of ChebyshevControl.m:23-30 at the query time
            const double tau = 2.0 * (t - knots[0]) / (knots[1] - knots[0]) - 1.0;
            for (int c = 0; c < nc; ++c) {
                double Tm = 1.0, Tk = tau, acc = v[(size_t)c * ldv + bv];
                if (K > 1) acc = acc + v[(size_t)(nc + c) * ldv + bv] * tau;
                for (int k = 2; k < K; ++k) {
                    const double Tn = 2.0 * tau * Tk - Tm;
                    acc = acc + v[((size_t)k * nc + c) * ldv + bv] * Tn;
                    Tm = Tk;
                    Tk = Tn;
                }
                out[((size_t)c * nq + q) * ldo + b] = acc;
            }
            continue;
        }
        int lo = 0, hi = K;                 // k = last knot <= t, clamped into the table
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (knots[mid] <= t) lo = mid; else hi = mid;
        }
        if (kind == UFUNC_PREVIOUS || K == 1) {
            for (int c = 0; c < nc; ++c) out[((size_t)c * nq + q) * ldo + b] = v[((size_t)lo * nc + c) * ldv + bv];
        } else {
            const int k = lo < K - 1 ? lo : K - 2;
            const double tc = fmin(fmax(t, knots[0]), knots[K - 1]);
            const double w = (tc - knots[k]) / (knots[k + 1] - knots[k]);
            for (int c = 0; c < nc; ++c) {
                const double va = v[((size_t)k * nc + c) * ldv + bv], vb = v[((size_t)(k + 1) * nc + c) * ldv + bv];
                out[((size_t)c * nq + q) * ldo + b] = va + w * (vb - va);
            }
        }
    }
}

}  // namespace ocs_interp
