/* ocs_b200.h -- C ABI of the B200-native RK4 trajectory sweeps
 *
 * Drop-in boundary for ONE hot path of DrDanRyan/Optimal-Control-Solvers: the
 * fixed-step RK4 forward state sweep and the backward discrete-adjoint (costate)
 * sweep behind single_shooting / an RK4-grid fb_sweep, batched over independent
 * trajectories.  The reference is pure MATLAB and has no FFI; every entry point
 * below names the MATLAB method it replaces (paths relative to the reference
 * tree).  A MEX gateway (optimal-control-solvers_b200/matlab/ocs_b200_mex.c) and a Python ctypes
 * binding (optimal-control-solvers_b200/_lib.py) bind exactly these symbols.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from ocs_b200_last_error(handle) (or with handle == NULL for
 *     create-time failures).  Nothing throws or longjmps across this boundary.
 *   - all host arrays are FP64, MATLAB (column-major) layout, read-only inputs
 *     are never retained past the call.  Host arrays may be pageable (every mxArray is): large transfers
 *     then go through a pinned ring and copy threads inside the handle instead of the driver's slow
 *     pageable path; pinned / registered arrays are handed to the copy engine as they are.
 *   - n = number of true states, m = n+1 (augmented with the running objective,
 *     last row), nc = controls, N = RK4 steps, G = 2N+1 grid points (nodes and
 *     interval midpoints), K = control basis size, Bt = trajectories in a batch.
 *   - there is NO CPU fallback: every compute entry point needs a CUDA device
 *     and fails with OCS_B200_ERR_CUDA when there is none.
 */
#ifndef OCS_B200_H
#define OCS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define OCS_B200_API __declspec(dllexport)
#else
#define OCS_B200_API __attribute__((visibility("default")))
#endif

#define OCS_B200_ABI_VERSION 1

/* status codes */
enum {
    OCS_B200_OK = 0,
    OCS_B200_ERR_ARG = 1,      /* bad argument / size / call order (MATLAB would raise a size error) */
    OCS_B200_ERR_CUDA = 2,     /* CUDA runtime/driver error, sticky for the handle */
    OCS_B200_ERR_COMPILE = 3,  /* NVRTC compilation of a generated problem failed (log in last_error) */
    OCS_B200_ERR_STATE = 4,    /* compute_adjoints before compute_states, etc. */
    OCS_B200_ERR_NOMEM = 5
};

/* built-in problems (device functions compiled into the library) */
enum {
    OCS_B200_PROBLEM_TEST = 0  /* tests/TestOCProblem.m:13-38, params = {c, m, r} */
};

/* arithmetic mode */
enum {
    OCS_B200_MATH_FAST = 0,    /* FMA contraction on (nvcc -fmad=true) and, for the built-in problem, the RK4 stage weights
                                  folded into the batch-uniform tables (re-associated products): results within 1e-10
                                  of the reference's operation order, not bit-identical to it */
    OCS_B200_MATH_STRICT = 1   /* every operation rounded separately, like MATLAB's element-wise ops */
};

/* reductions over the batch (ocs_b200_reduce_J) */
enum { OCS_B200_REDUCE_SUM = 0, OCS_B200_REDUCE_MIN = 1, OCS_B200_REDUCE_MAX = 2 };

typedef struct ocs_b200_handle ocs_b200_handle;

/* ------------------------------------------------------------------ lifetime */

/* Replaces the constructors RK4Integrator(tspan) (Integrator/RK4Integrator.m:16-25) and
 * RK4InfiniteIntegrator(tspan, tspanExtra, uStar) (Integrator/RK4InfiniteIntegrator.m:12-17).
 *   tspan        [N+1]   node times of the main horizon
 *   tspan_extra  [N2+1]  node times of the tail horizon, or NULL / N2 == 0 for a finite horizon
 *   uStar        [nc]    constant tail control (read at problem-binding time; may be NULL when N2 == 0)
 *   device       CUDA device ordinal
 *   math_mode    OCS_B200_MATH_FAST or OCS_B200_MATH_STRICT
 * h = diff(tspan) and the augmented grid t (nodes + midpoints) are formed with the same
 * single IEEE operations as RK4Integrator.m:17-24 and can be read back with ocs_b200_get_grid. */
OCS_B200_API int ocs_b200_create(int N, const double *tspan, int N2, const double *tspan_extra,
                                 int device, int math_mode, ocs_b200_handle **out);

/* The same integrator spread over several GPUs of one node behind ONE handle: the reference's host is a single
 * MATLAB process that constructs one integrator object (functions/single_shooting.m:41-45), so multi-GPU lives inside
 * the handle, not in extra processes.  devices [ndev]: distinct CUDA ordinals; devices[0] is the root.  A batch of Bt
 * trajectories is cut into contiguous shards (the first Bt mod ndev take one trajectory more); every host-buffer
 * entry point below forwards shard k to device k on its own host thread and stream, reading and writing the
 * caller's arrays directly (column-major with the trajectory index last: a shard is one contiguous block), so there is
 * no collective on that path.  Entry points that take DEVICE pointers or a stream of one device
 * (ocs_b200_set_stream, *_dev) are refused on such a handle; use ocs_b200_evaluate_resident,
 * ocs_b200_gather_results_dev and ocs_b200_get_results instead. */
OCS_B200_API int ocs_b200_create_multi(int N, const double *tspan, int N2, const double *tspan_extra, int ndev,
                                       const int *devices, int math_mode, ocs_b200_handle **out);

/* Devices behind a handle and the shard offsets of its resident batch: devices [<= cap], offsets [<= cap] with
 * offsets[k] .. offsets[k+1] the trajectories of device k (ndev + 1 entries when cap allows).  Any pointer may be NULL. */
OCS_B200_API int ocs_b200_get_devices(const ocs_b200_handle *h, int *ndev, int *devices, long *offsets, int cap);

/* Sizes the handle was built / bound with: dims = { n, nc, np, K, N, N2, 0, 0 } (n, nc, np are 0 before a problem is
 * bound, K is 0 without a basis).  A binding (the MEX gateway) sizes its outputs from these, not from caller-supplied
 * scalars.  ocs_b200_get_batch: trajectories of the resident batch (0: none). */
OCS_B200_API int ocs_b200_get_dims(const ocs_b200_handle *h, int32_t dims[8]);
OCS_B200_API long ocs_b200_get_batch(const ocs_b200_handle *h);

/* delete(obj): frees the device-resident trajectory store (the xK cache of RK4Integrator.m:8-10). */
OCS_B200_API int ocs_b200_destroy(ocs_b200_handle *h);

/* Message of the last failure on this handle (h == NULL: last create-time failure). Never NULL. */
OCS_B200_API const char *ocs_b200_last_error(const ocs_b200_handle *h);

OCS_B200_API int ocs_b200_abi_version(void);

/* Property t / h of RK4Integrator (Integrator/RK4Integrator.m:4-5).  segment 0 = main, 1 = tail.
 * t [2N+1], h [N]; either may be NULL. */
OCS_B200_API int ocs_b200_get_grid(const ocs_b200_handle *h, int segment, double *t, double *hstep);

/* ------------------------------------------------------------------ problem (OCProblem/OCProblem.m:3-21) */

/* Bind a built-in OCProblem (F, dFdx_times_vec, dFdu_times_vec compiled into the library). */
OCS_B200_API int ocs_b200_set_problem_builtin(ocs_b200_handle *h, int problem_id);

/* Bind a generated OCProblem: CUDA source of the device functions emitted by the
 * make_from_symbolic CUDA backend (replaces matlabFunction in functions/make_from_symbolic.m:27-31),
 * compiled here with NVRTC for sm_100a together with the sweep kernels.
 *   kernel_include_dir: directory holding rk4_kernels.cuh (NULL = the library's own csrc directory). */
OCS_B200_API int ocs_b200_set_problem_source(ocs_b200_handle *h, const char *cuda_source,
                                             int n, int nc, int np, const char *kernel_include_dir);

/* Problem parameters (TestOCProblem.m:16-18 c,m,r; make_from_symbolic.m:3-4 params struct).
 *   params [np x nbatch] column-major; nbatch == 1: shared by the batch, else one column per trajectory. */
OCS_B200_API int ocs_b200_set_params(ocs_b200_handle *h, const double *params, int np, long nbatch);

/* uStar of RK4InfiniteIntegrator.m:15 (constant tail control, [nc]). */
OCS_B200_API int ocs_b200_set_tail_control(ocs_b200_handle *h, const double *uStar, int nc);

/* ------------------------------------------------------------------ control (Control/Control.m:4-14) */

/* Attach the basis matrix B of a Control object (PWConstantControl.m:30-38, PWLinearControl.m:31-50,
 * ChebyshevControl.m:21-31): B is K x G column-major, built by the host on integrator.t.
 * Columns with <= 2 adjacent non-zeros are stored as an (index, w0, w1) table -- the index is the
 * reference's control-grid index, taken from B itself, hence bit-exact; denser B is kept dense.
 * With a basis attached, compute_u (u = reshape(v,nc,[])*B) and compute_dJdv (dJdv = dJdu*B') are
 * fused into the sweeps.  K == 0 detaches. */
OCS_B200_API int ocs_b200_set_control_basis(ocs_b200_handle *h, int K, const double *B);

/* Reads back the compressed basis: k_index[G] (0-based first non-zero row) and nnz[G]; NULL to skip.
 * For a dense basis k_index is all 0 and nnz all K. */
OCS_B200_API int ocs_b200_get_control_index(const ocs_b200_handle *h, int32_t *k_index, int32_t *nnz);

/* Control coefficients v (argument of compute_u, PWLinearControl.m:59-62): [(nc*K) x nbatch],
 * element (k*nc + c) = control c, basis k (0-based).  nbatch == 1: shared; else one column per trajectory. */
OCS_B200_API int ocs_b200_set_control_v(ocs_b200_handle *h, const double *v, long nbatch);

/* Control values on the grid (argument u of compute_states, RK4Integrator.m:28): [nc x G x nbatch]. */
OCS_B200_API int ocs_b200_set_control_u(ocs_b200_handle *h, const double *u, long nbatch);

/* u = control.compute_u(v) for the current v: [nc x G x nb] for trajectories [b0, b0+nb)
 * (nb is clamped to 1 for a shared control). */
OCS_B200_API int ocs_b200_get_control_u(ocs_b200_handle *h, long b0, long nb, double *u);

/* ------------------------------------------------------------------ sweeps */

/* [x, J] = integrator.compute_states(prob, x0, u)  (RK4Integrator.m:28-56; with a tail:
 * RK4InfiniteIntegrator.m:20-24) for Bt trajectories at once.
 *   x0     [n x Bt]
 *   J      [Bt]       objective (J1 + J2 with a tail)                         (may be NULL)
 *   xT     [n x Bt]   state at the end of the main horizon                    (may be NULL)
 *   status [Bt]       0 = finite, 1 = J or x(T) not finite                    (may be NULL)
 * Node states stay on the device for compute_adjoints / fetch_states. */
OCS_B200_API int ocs_b200_compute_states(ocs_b200_handle *h, const double *x0, long Bt,
                                         double *J, double *xT, int32_t *status);

/* [lam, dJdu] = integrator.compute_adjoints(prob, u [, lamT])  (RK4Integrator.m:59-121,
 * RK4InfiniteIntegrator.m:27-30) fused with dJdv = control.compute_dJdv(dJdu) when a basis is attached.
 * Uses the trajectories of the last compute_states on this handle (RK4Integrator.m:8-10).
 *   lamT  [m x nbatch] terminal costate of the main horizon, nbatch in {1, Bt}; NULL = e_m
 *         (ignored when the handle has a tail: the tail sweep supplies it, RK4InfiniteIntegrator.m:28-29)
 *   dJdv  [(nc*K) x Bt]   (NULL to skip; requires a basis)
 *   lam0  [m x Bt]        lam(:,1), the free-initial-state gradient of single_shooting.m:147-148 (NULL to skip) */
OCS_B200_API int ocs_b200_compute_adjoints(ocs_b200_handle *h, const double *lamT, long lamT_nbatch,
                                           double *dJdv, double *lam0);

/* nlpObjective of functions/single_shooting.m:137-150 in ONE call on host buffers: compute_states followed by
 * compute_adjoints (same arguments, same results as the two calls in sequence).  For large batches the batch is
 * swept in trajectory chunks -- forward and reverse sweep of a chunk back to back -- while the gradient of the
 * previous chunk is transposed and downloaded on a second stream, so the PCIe transfer of dJdv ((nc*K) x Bt
 * doubles) overlaps all of the arithmetic.  Any output may be NULL. */
OCS_B200_API int ocs_b200_evaluate(ocs_b200_handle *h, const double *x0, long Bt, const double *lamT, long lamT_nbatch,
                                   double *J, double *xT, int32_t *status, double *dJdv, double *lam0);

/* Full grids for trajectories [b0, b0+nb), exactly the arrays the MATLAB methods return:
 *   x    [m x (N+1) x nb]   (RK4Integrator.m:54)     lam  [m x (N+1) x nb]  (RK4Integrator.m:68-89)
 *   dJdu [nc x G x nb]      (RK4Integrator.m:97-121)
 * segment 0 = main horizon, 1 = tail.  Any pointer may be NULL. */
OCS_B200_API int ocs_b200_fetch_states(ocs_b200_handle *h, int segment, long b0, long nb, double *x);
OCS_B200_API int ocs_b200_fetch_adjoints(ocs_b200_handle *h, int segment, long b0, long nb,
                                         double *lam, double *dJdu);

/* One fb_sweep iteration loop on the RK4 grid (functions/fb_sweep.m:79-115 composed with the RK4
 * sweeps): u lives on the G grid points per trajectory; per sweep  x = forward(u), lam = adjoint(u),
 * uNew = ControlChar(t, x, lam) clamped to bounds (make_from_symbolic.m:102-112), convergence
 * max_j |uNew-u| / (uRelTol |u| + uAbsTol) <= 1 (fb_sweep.m:107-110) per trajectory.
 *   x0 [n x Bt]; u0 [nc x G x nbatch] initial control (nbatch in {1, Bt}), or NULL: start from the coefficients of
 *   ocs_b200_set_control_v through the attached basis, u0 = reshape(v,nc,[])*B evaluated on the device (one
 *   control candidate per trajectory without uploading nc x G x Bt values);
 *   bounds [nc x 2] ControlBounds; relax = 1 is the reference's pure replacement (fb_sweep.m:96).
 *   J [Bt], sweeps [Bt] (sweeps used), converged [Bt] (0/1), max_change [Bt]: any may be NULL.
 * The converged control stays on the device (ocs_b200_get_control_u), states/adjoints via fetch_*. */
OCS_B200_API int ocs_b200_fb_sweep(ocs_b200_handle *h, const double *x0, long Bt,
                                   const double *u0, long u0_nbatch, const double *bounds,
                                   double uRelTol, double uAbsTol, int nSWEEPS, double relax,
                                   double *J, int32_t *sweeps, int32_t *converged, double *max_change);

/* Which costate the fb_sweep iteration uses (default DISCRETE).
 *   DISCRETE:   lam = the v2 Integrator's discrete adjoint of the RK4 scheme (Integrator/RK4Integrator.m:59-94).
 *   CONTINUOUS: the v1 struct API's costate equation  lam' = prob.adjointRHS(t, x, lam, u) = -gradient(H, x)
 *               (functions/make_from_symbolic.m:11-14, functions/compute_x_lam.m:11-12, compute_x_lam_J.m:18-19)
 *               integrated backwards from lam(T) = 0 by classical RK4 on the fixed grid (the reference integrates
 *               it with the absent odevr7); midpoint state = mean of the neighbouring nodes.  adjointRHS is
 *               evaluated as -dFdx_times_vec(t, y, u, [lam; 1]), the same expression. */
#define OCS_B200_COSTATE_DISCRETE 0
#define OCS_B200_COSTATE_CONTINUOUS 1
OCS_B200_API int ocs_b200_set_fb_costate(ocs_b200_handle *h, int mode);

/* Batched single shooting: the solve behind functions/single_shooting.m:114
 *     [vOpt, J] = fmincon(@nlpObjective, v0, [], [], [], [], Lb, Ub)
 * for every trajectory at once, entirely on the device.  fmincon (SQP, serial, host) is replaced by a
 * spectral projected-gradient method (Barzilai-Borwein step, non-monotone Armijo safeguard, box projection
 * onto [Lb, Ub] from control.compute_nlp_bounds): each iteration is one fused objective/gradient evaluation
 * (compute_states + compute_adjoints) of the whole batch plus one update kernel; trajectories converge and
 * freeze independently.  The minimiser, not the iterate path, is what is comparable with fmincon.
 *   x0 [n x Bt]; v0 [(nc*K) x nbatch] (nbatch in {1, Bt}); Lb, Ub [(nc*K)]
 *   free_states [nfree] 0-based rows of x0 that are optimised too (single_shooting.m:82-97,143-149) with
 *   bounds free_lb/free_ub [nfree]; nfree = 0 and NULLs for none
 *   tol: stop when the projected-gradient max-norm ||v - P(v - dJdv)|| <= tol;  max_iter evaluations at most
 *   v_opt [(nc*K + nfree) x Bt], J [Bt], iters [Bt], converged [Bt]: any may be NULL
 * Requires an attached basis.  States/adjoints of the minimisers stay on the device for fetch_*. */
OCS_B200_API int ocs_b200_single_shooting(ocs_b200_handle *h, const double *x0, long Bt,
                                          const double *v0, long v0_nbatch, const double *Lb, const double *Ub,
                                          int nfree, const int32_t *free_states, const double *free_lb,
                                          const double *free_ub, int max_iter, double tol,
                                          double *v_opt, double *J, int32_t *iters, int32_t *converged);

/* Linear constraints of the batched solve (functions/single_shooting.m:106-111: [A, b] = control.compute_lincon(
 * prob.ControlBounds), fmincon(..., A, b, ...)).  The reference's PW controls need none (their coefficient box bounds
 * u(t), compute_nlp_bounds) and its ChebyshevControl left compute_lincon empty (Control/ChebyshevControl.m:51-53) although
 * a box on Chebyshev coefficients does not bound u(t).  The finished constraint is ControlBounds at every grid point,
 *     lb_c <= (reshape(v, nc, []) * B)(c, j) <= ub_c      i.e.  A = [kron(B', I); -kron(B', I)],  b = [ub; -lb],
 * and that is what this switches on for the solves of a handle with a dense (Chebyshev) basis: an augmented-Lagrangian
 * outer iteration (Powell-Hestenes-Rockafellar multipliers per grid value and trajectory, penalty parameter rho0, raised fourfold
 * whenever a round does not cut the largest violation to a quarter, at most max_outer rounds) around the projected-gradient solve; the penalty and its gradient are applied on
 * the control and dJdu grids the dense-basis gradient path holds anyway, before dJdu is projected onto B.  Trajectories
 * are finished when their projected gradient meets tol AND their largest bound violation is <= feas_tol.
 *   bounds [nc x 2] column-major ControlBounds (nc must equal the problem's control count when the solve runs), or
 *   NULL to switch the constraints off
 *   rho0, feas_tol, max_outer: <= 0 selects 100, 1e-4, 12.  The inner problems of an augmented Lagrangian are
 *   ill-conditioned for a first-order method: expect 10^4 evaluations per solve where the unconstrained solve takes 10^1-10^2.
 * ocs_b200_get_solve_violation: the largest violation per trajectory [Bt] after such a solve.  The reported J is the
 * objective alone (no penalty term). */
OCS_B200_API int ocs_b200_set_solve_control_bounds(ocs_b200_handle *h, const double *bounds, int nc, double rho0,
                                                   double feas_tol, int max_outer);
OCS_B200_API int ocs_b200_get_solve_violation(ocs_b200_handle *h, double *viol);

/* The `soln` struct of the reference, evaluated on the device for trajectories [b0, b0+nb) at the query
 * times tq[nq]:
 *   x   [n  x nq x nb] = vectorInterpolant(tspan, xOpt(1:end-1,:), 'pchip')(tq)     functions/single_shooting.m:129
 *   lam [n  x nq x nb] = vectorInterpolant(tspan, lamOpt(1:end-1,:), 'pchip')(tq)   functions/single_shooting.m:130
 *   u   [nc x nq x nb] = control.compute_uFunc(vOpt)(tq)                            functions/single_shooting.m:128
 *                        (with a per-trajectory control grid, i.e. after fb_sweep: the pchip interpolant of the grid
 *                         values, functions/fb_sweep.m:123)
 * 'pchip' is MATLAB's shape-preserving cubic (functions/vectorInterpolant.m:1-12 -> griddedInterpolant).  x reads the
 * resident node states; lam re-runs the costate sweep of the range (same terminal costate as the last
 * compute_adjoints); u needs ocs_b200_set_control_ufunc.  Any output may be NULL.  segment as in fetch_*. */
OCS_B200_API int ocs_b200_eval_solution(ocs_b200_handle *h, int segment, long b0, long nb, int nq, const double *tq,
                                        double *x, double *lam, double *u);

/* How control.compute_uFunc interpolates the coefficients v of the attached basis:
 *   PREVIOUS  vectorInterpolant(intervalStarts, v, 'previous')  Control/PWConstantControl.m:58-61, knots = intervalStarts [K]
 *   LINEAR    vectorInterpolant(controlPts, v, 'linear')        Control/PWLinearControl.m:74-77,   knots = controlPts [K]
 *   CHEBYSHEV sum_k v_k T_k(tau(t)) with the recurrence of Control/ChebyshevControl.m:23-30,       knots = [t(1), t(end)]
 * Queries outside the knots take the nearest end value (PREVIOUS, LINEAR). */
#define OCS_B200_UFUNC_NONE 0
#define OCS_B200_UFUNC_PREVIOUS 1
#define OCS_B200_UFUNC_LINEAR 2
#define OCS_B200_UFUNC_CHEBYSHEV 3
OCS_B200_API int ocs_b200_set_control_ufunc(ocs_b200_handle *h, int kind, int nknots, const double *knots);

/* Batched compute_equilibrium (functions/compute_equilibrium.m:1-33): for every parameter set, the steady state
 * [xStar; lamStar; uStar] of the infinite-horizon problem, i.e. the solution within [lb, ub] of (:10-21)
 *     F(0,x,u)(1:n) = 0;   r*lam - dFdx_times_vec(0,x,u,[lam;1])(1:n) = 0;   dFdu_times_vec(0,x,u,[lam;1]) = 0.
 * The reference solves ONE system per call with lsqnonlin (:23-27); here every parameter set is solved by its own
 * thread (Levenberg-Marquardt, central-difference Jacobian, steps projected onto the bounds), using the bound
 * problem's device functions -- parameter sweeps of RK4InfiniteIntegrator need one uStar per set.
 *   params [np x nparams], nparams in {1, nsets}; guess [(2n+nc) x nguess] = [xGuess; lamGuess; uGuess], nguess in {1, nsets};
 *   lb, ub [(2n+nc)] or NULL (unbounded); r [nr] discount rate(s), nr in {1, nsets};
 *   max_iter (reference: MaxIter 4000), tol: stop when max_k |step_k| <= tol (1 + |y_k|) or resnorm <= tol^2.
 *   y [(2n+nc) x nsets]; resnorm [nsets] = sum(residual.^2); residual [(2n+nc) x nsets];
 *   exitflag [nsets] (1 converged, 0 iteration limit, -1 not finite); iters [nsets].  All but y may be NULL.
 * Needs a bound problem only (no grid data is used). */
OCS_B200_API int ocs_b200_compute_equilibrium(ocs_b200_handle *h, const double *params, int np, long nparams, long nsets,
                                              const double *guess, long nguess, const double *lb, const double *ub,
                                              const double *r, long nr, int max_iter, double tol, double *y,
                                              double *resnorm, double *residual, int32_t *exitflag, int32_t *iters);

/* Batch reduction of the last J (warp-shuffle tree): value and arg index (index may be NULL). */
OCS_B200_API int ocs_b200_reduce_J(ocs_b200_handle *h, int op, double *value, long *index);

/* ------------------------------------------------------------------ device-resident variants
 * Same sweeps with caller-owned DEVICE buffers in structure-of-arrays layout [row][Bt]
 * (batch fastest) -- for pipelines that keep the batch in HBM on ONE device.
 * stream: a cudaStream_t (NULL = the handle's own stream). */
OCS_B200_API int ocs_b200_set_stream(ocs_b200_handle *h, void *cuda_stream);
OCS_B200_API int ocs_b200_set_control_v_dev(ocs_b200_handle *h, const double *v_soa, long nbatch);
OCS_B200_API int ocs_b200_compute_states_dev(ocs_b200_handle *h, const double *x0_soa, long Bt,
                                             double *J, double *xT_soa, int32_t *status);
OCS_B200_API int ocs_b200_compute_adjoints_dev(ocs_b200_handle *h, const double *lamT_soa, long lamT_nbatch,
                                               double *dJdv_soa, double *lam0_soa);
/* nlpObjective of functions/single_shooting.m:137-150 (compute_states followed by compute_adjoints) as ONE asynchronous
 * call on device buffers.  Both sweeps are queued together, which lets a batch whose kernels keep few warps resident
 * (several-state problems) run as two half batches on two internal streams -- one half's adjoint sweep fills the SMs
 * that the other half's forward sweep leaves idle in its last round of blocks.  Results are complete in stream order
 * on the handle's stream, as for the two calls above.  OCS_B200_LANES=0 / 1 in the environment at create time
 * forces the one-pass / two-lane form; the default decides from the kernels' occupancy and the batch size. */
OCS_B200_API int ocs_b200_evaluate_dev(ocs_b200_handle *h, const double *x0_soa, long Bt, const double *lamT_soa,
                                       long lamT_nbatch, double *J, double *xT_soa, int32_t *status, double *dJdv_soa,
                                       double *lam0_soa);
OCS_B200_API int ocs_b200_synchronize(ocs_b200_handle *h);

/* Device-resident evaluation for solver loops and multi-GPU pipelines (works on single- and multi-device handles):
 *   ocs_b200_evaluate_resident  nlpObjective (functions/single_shooting.m:137-150) of the batch whose initial states are
 *                               x0 [n x Bt] (host; uploaded and kept), or of the resident initial states when x0 == NULL
 *                               (Bt ignored) -- e.g. after ocs_b200_set_control_v with a new candidate.  want_gradient
 *                               != 0 also runs the adjoint sweep.  All results stay on the device(s); asynchronous.
 *   ocs_b200_gather_results_dev J [Bt] and x(T) [n][Bt] (SoA) into DEVICE buffers on the root device -- the one
 *                               exchange of the multi-GPU path (north star: "NCCL ... only to gather the small
 *                               per-trajectory J and final-state results"): a gather to root with grouped NCCL
 *                               ncclSend/ncclRecv (ncclCommInitAll inside this process), or peer copies on the copy
 *                               engines when libnccl is not loadable or OCS_B200_GATHER=p2p.  Asynchronous; ordered on
 *                               the root device's stream.  Either pointer may be NULL.
 *   ocs_b200_get_results        the resident results into host arrays (any may be NULL). */
OCS_B200_API int ocs_b200_evaluate_resident(ocs_b200_handle *h, const double *x0, long Bt, int want_gradient);
OCS_B200_API int ocs_b200_gather_results_dev(ocs_b200_handle *h, double *J_root, double *xT_root_soa);
OCS_B200_API int ocs_b200_get_results(ocs_b200_handle *h, double *J, double *xT, int32_t *status, double *dJdv, double *lam0);

/* Device time (ms, CUDA events on the launch stream) and launch count of the sweep kernels since
 * the last call with reset != 0. */
OCS_B200_API int ocs_b200_kernel_stats(ocs_b200_handle *h, int reset, double *ms_forward, double *ms_reverse,
                                       long *launches);
/* The same split further: ms[0] forward sweeps, ms[1] reverse sweeps, ms[2] dense-basis projection dJdu*B',
 * ms[3] solver update (single_shooting); ms[4] device time of the ocs_b200_fb_sweep iteration loops from the first
 * kernel queued to the last one's end (measured whether or not profiling is on; with profiling off the sweeps of a
 * low-occupancy problem overlap as two lanes, so ms[4] is what a solve costs while ms[0..1] -- profiling on -- time each
 * kernel alone); entries beyond nms are not written.  Multi-device: the slowest device. */
OCS_B200_API int ocs_b200_kernel_stats_ex(ocs_b200_handle *h, int reset, double *ms, int nms, long *launches);
/* enable/disable the CUDA-event bracketing used by ocs_b200_kernel_stats (off by default) */
OCS_B200_API int ocs_b200_set_profiling(ocs_b200_handle *h, int on);

/* ------------------------------------------------------------------ measurement helpers (not on the path) */

/* FP64 FMA peak of the device (dependent DFMA chains, all SMs): TFLOP/s. */
OCS_B200_API int ocs_b200_measure_fp64_peak(int device, double *tflops);
/* The same with longer runs and the SM clock the device ran at meanwhile (SM cycles of the kernel / its duration):
 * nominal peak at that clock = SMs x 64 FMA/clk x 2 x sm_mhz. */
OCS_B200_API int ocs_b200_measure_fp64_peak_ex(int device, double *tflops, double *sm_mhz);
/* HBM streaming bandwidth with the sweeps' access pattern (FP64 coalesced write + read): GB/s each. */
OCS_B200_API int ocs_b200_measure_hbm(int device, size_t bytes, double *write_gbs, double *read_gbs);

#ifdef __cplusplus
}
#endif
#endif /* OCS_B200_H */
