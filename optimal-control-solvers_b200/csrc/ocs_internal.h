// ocs_internal.h -- what the translation units behind the C ABI share: the handle (the device-resident
// trajectory store of one integrator on ONE device, or the fan-out over several of them) and the error macros.
// Not part of the public interface (include/ocs_b200.h is).
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/ocs_b200.h"
#include "host_stager.h"
#include "kernel_table.h"

struct ocs_multi;      // ocs_multi.cu: one handle per device behind a single handle

namespace ocs_detail {

constexpr int OCS_FOLD_MAX = 15;     // most folded table entries per stage position a problem may declare

struct Segment {
    int N = 0;
    std::vector<double> tspan, t, h;
    double *d_gprec = nullptr, *d_steprec = nullptr, *d_basrec = nullptr;
    double *d_foldrec = nullptr;            // weight-folded table entries [N][even(3*nfold)] (rk4_kernels.cuh), or null
    int fold_nf = 0;                        // nfold the fold records were built for (0: none)
    size_t gprec_cap = 0, basrec_cap = 0;   // doubles
    double *h_rec = nullptr;  // pinned staging of the records (async upload)
    size_t h_rec_cap = 0;
    int BW = 0;               // doubles per basis record
    int hm = 0;               // hoisting mode the records were built for
    bool basis_dirty = true;
    cudaEvent_t rec_uploaded = nullptr;
    bool steps_uploaded = false;
    double *d_xs_base = nullptr;   // allocation: OCS_XS_SLACK*Bt doubles of slack, then node states [n][N+1][Bt]
    double *d_xs = nullptr;        // = d_xs_base + OCS_XS_SLACK*Bt
    size_t xs_cap = 0;             // doubles (whole allocation)
    int R = 0;
    bool records_dirty = true;
};

enum { CTRL_NONE = 0, CTRL_V_SHARED, CTRL_V_BATCH, CTRL_U_SHARED, CTRL_U_BATCH };
enum { BASIS_NONE = 0, BASIS_SPARSE = 1, BASIS_DENSE = 2 };

struct DevBuf {
    double *p = nullptr;
    size_t cap = 0;   // doubles
};

}  // namespace ocs_detail

struct ocs_b200_handle {
    typedef ocs_detail::Segment Segment;
    typedef ocs_detail::DevBuf DevBuf;
    ocs_multi *multi = nullptr;        // non-null: this handle fans every call out over one handle per device (ocs_multi.cu)
    ocs_detail::HostStager stager;     // pinned ring + copy threads for pageable caller buffers (host_stager.h)
    int device = 0;
    int math_mode = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::string err;
    int sticky = 0;

    Segment seg[2];
    bool has_tail = false;

    // problem
    bool have_problem = false;
    ocs_kernel_table kt{};
    std::vector<void *> gen_libraries;
    std::string gen_source, gen_include;
    int gen_n = 0, gen_nc = 0, gen_np = 0;
    bool is_generated = false;

    // params
    std::vector<double> params_shared;
    long params_nbatch = 0;
    DevBuf d_params;
    std::vector<double> uStar;

    // control
    int basis_kind = ocs_detail::BASIS_NONE;
    int K = 0;
    std::vector<int32_t> bk, bnnz;
    bool bk_monotone = false;          // knot index non-decreasing along the grid: dJdv columns are stored once
    bool bk_covered = false;           // ... and every knot is reached by the window, so dJdv needs no zero-fill
    std::vector<double> bw0, bw1, bdense;
    int ctrl_mode = ocs_detail::CTRL_NONE;
    std::vector<double> ctrl_shared;   // v (nc*K) or u (nc*G)
    long ctrl_nbatch = 0;
    DevBuf d_v, d_ug;
    const double *v_ext = nullptr;     // caller-owned device v (set_control_v_dev)
    const int *fb_frozen = nullptr;    // trajectories the sweeps must leave untouched (single_shooting / fb_sweep)
    double *lams_req = nullptr;        // set while ocs_b200_eval_solution wants the node costates of a segment (SoA)
    long lams_ld = 0;
    int lams_segment = 0;
    int ufunc_kind = 0;                // compute_uFunc of the attached control (ocs_b200_set_control_ufunc)
    std::vector<double> ufunc_knots;
    int fb_costate = 0;                // OCS_B200_COSTATE_* used by ocs_b200_fb_sweep
    bool user_lamT = false;            // last compute_adjoints carried an explicit lamT (kept in d_lamT_user)
    long user_lamT_ld = 0;

    // batch state
    long Bt = 0;
    bool have_states = false;
    DevBuf d_x0, d_J, d_xT, d_lamT_tail, d_lamT_user, d_dJdv, d_lam0, d_stage, d_stage2, d_maxchg;
    DevBuf d_dug;                     // dJdu grid [nc][G][Bt] of the dense-basis gradient path
    DevBuf d_z, d_g, d_zt, d_gt, d_Jacc, d_Jhist, d_alpha, d_lbub, d_pred;   // single_shooting (SPG) state
    DevBuf d_pipe[2];                 // double-buffered transpose staging of the chunked dJdv download
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t chunk_done[2] = {nullptr, nullptr};
    cudaEvent_t stage_free[2] = {nullptr, nullptr};   // the download out of staging buffer k has finished
    // Two lanes (ocs_api.cu, lanes_*): sweeps whose kernels keep few warps resident (a several-state problem: 8-12 per
    // SM) leave the SMs of their last, partly filled round idle; the batch is then swept as two halves on two streams
    // so that one half's next kernel fills what the other half's last round leaves free.
    cudaStream_t lane_stream[2] = {nullptr, nullptr};
    cudaEvent_t lane_fork = nullptr, lane_done[2] = {nullptr, nullptr};
    cudaEvent_t lane_sweep[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // fb_sweep: sweep s of lane l has been counted
    int *d_lane_open = nullptr;        // fb_sweep: open trajectories per lane and sweep parity [2][2]
    int *h_lane_open = nullptr;        // ... pinned host copy
    int lanes_mode = -1;               // -1 automatic, 0 never, 1 whenever the batch can be halved (OCS_B200_LANES)
    int sm_count = 0;
    // ControlBounds at every grid point as constraints of the batched solve (ocs_b200_set_solve_control_bounds):
    // augmented-Lagrangian multipliers mu [2][nc][G][Bt] (upper, lower), the bounds [nc][2], the largest violation
    // per trajectory, and the penalty parameter of the evaluation in flight (cgb_active: run_reverse adds the term)
    bool cgb_on = false, cgb_active = false;
    std::vector<double> cgb_bounds;
    double cgb_rho0 = 100.0, cgb_tol = 1e-4, cgb_rho = 0.0;
    int cgb_outer = 12;
    DevBuf d_cgb_mu, d_cgb_b, d_cgb_viol;
    int *d_status = nullptr;
    size_t status_cap = 0;
    int *d_frozen = nullptr, *d_sweeps = nullptr;
    size_t frozen_cap = 0;
    double *d_bounds = nullptr;
    double *d_red = nullptr;   // reduction scratch
    int *d_errflag = nullptr;  // set by a kernel whose compiled record layout disagrees with the host's: the device
    int *h_errflag = nullptr;  // alias of a mapped pinned host word, read (no copy) at every synchronisation point
    bool staged = false;       // copies are pending in the pinned ring (finish_host delivers them)
    bool fb_resident = false;  // the resident states/J belong to the control BEFORE the last fb_sweep update (fb_sweep.m:82:
                               // final_sweep(u)), while d_ug already holds uNew = soln.u: fetch_states must not re-sweep

    // profiling
    bool profiling = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    double ms_fwd = 0, ms_rev = 0;          // ms_rev includes ms_proj (the dense-basis projection after the sweep)
    double ms_proj = 0, ms_upd = 0;         // projection kernel; solver update kernel (single_shooting)
    cudaEvent_t ev_loop[2] = {nullptr, nullptr};
    double ms_loop = 0;                     // device time of the fb_sweep iteration loops (first kernel queued to the
                                            // last one's end on the handle's stream, lanes included); always measured
    long launches = 0;

    int fail(int code, const char *fmt, ...)
    {
        char buf[1024];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        if (code == OCS_B200_ERR_CUDA) sticky = code;
        return code;
    }
};

typedef ocs_b200_handle H;

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return h->fail(OCS_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                           __FILE__, __LINE__);                                                         \
    } while (0)
#define TRY(expr)                  \
    do {                           \
        int rc_ = (expr);          \
        if (rc_ != 0) return rc_;  \
    } while (0)


// ---------------------------------------------------------------- multi-device fan-out (ocs_multi.cu)
// A handle made by ocs_b200_create_multi owns one ordinary handle per device and a worker thread for each; every
// entry point of the C ABI forwards to these when h->multi is set.
const ocs_b200_handle *ocs_multi_first(const ocs_b200_handle *h);
long ocs_multi_batch(const ocs_b200_handle *h);
int ocs_multi_destroy(ocs_b200_handle *h);
int ocs_multi_get_grid(const ocs_b200_handle *h, int segment, double *t, double *hstep);
int ocs_multi_set_problem_builtin(ocs_b200_handle *h, int problem_id);
int ocs_multi_set_problem_source(ocs_b200_handle *h, const char *src, int n, int nc, int np, const char *inc);
int ocs_multi_set_params(ocs_b200_handle *h, const double *params, int np, long nbatch);
int ocs_multi_set_tail_control(ocs_b200_handle *h, const double *uStar, int nc);
int ocs_multi_set_control_basis(ocs_b200_handle *h, int K, const double *B);
int ocs_multi_get_control_index(const ocs_b200_handle *h, int32_t *k_index, int32_t *nnz);
int ocs_multi_set_control_v(ocs_b200_handle *h, const double *v, long nbatch);
int ocs_multi_set_control_u(ocs_b200_handle *h, const double *u, long nbatch);
int ocs_multi_get_control_u(ocs_b200_handle *h, long b0, long nb, double *u);
int ocs_multi_compute_states(ocs_b200_handle *h, const double *x0, long Bt, double *J, double *xT, int32_t *status);
int ocs_multi_compute_adjoints(ocs_b200_handle *h, const double *lamT, long lamT_nbatch, double *dJdv, double *lam0);
int ocs_multi_evaluate(ocs_b200_handle *h, const double *x0, long Bt, const double *lamT, long lamT_nbatch, double *J,
                       double *xT, int32_t *status, double *dJdv, double *lam0);
int ocs_multi_fetch_states(ocs_b200_handle *h, int segment, long b0, long nb, double *x);
int ocs_multi_fetch_adjoints(ocs_b200_handle *h, int segment, long b0, long nb, double *lam, double *dJdu);
int ocs_multi_fb_sweep(ocs_b200_handle *h, const double *x0, long Bt, const double *u0, long u0_nbatch, const double *bounds,
                       double uRelTol, double uAbsTol, int nSWEEPS, double relax, double *J, int32_t *sweeps,
                       int32_t *converged, double *max_change);
int ocs_multi_set_fb_costate(ocs_b200_handle *h, int mode);
int ocs_multi_set_solve_control_bounds(ocs_b200_handle *h, const double *bounds, int nc, double rho0, double feas_tol, int max_outer);
int ocs_multi_get_solve_violation(ocs_b200_handle *h, double *viol);
int ocs_multi_single_shooting(ocs_b200_handle *h, const double *x0, long Bt, const double *v0, long v0_nbatch, const double *Lb,
                              const double *Ub, int nfree, const int32_t *free_states, const double *free_lb,
                              const double *free_ub, int max_iter, double tol, double *v_opt, double *J, int32_t *iters,
                              int32_t *converged);
int ocs_multi_eval_solution(ocs_b200_handle *h, int segment, long b0, long nb, int nq, const double *tq, double *x, double *lam,
                            double *u);
int ocs_multi_set_control_ufunc(ocs_b200_handle *h, int kind, int nknots, const double *knots);
int ocs_multi_compute_equilibrium(ocs_b200_handle *h, const double *params, int np, long nparams, long nsets, const double *guess,
                                  long nguess, const double *lb, const double *ub, const double *r, long nr, int max_iter,
                                  double tol, double *y, double *resnorm, double *residual, int32_t *exitflag, int32_t *iters);
int ocs_multi_reduce_J(ocs_b200_handle *h, int op, double *value, long *index);
int ocs_multi_synchronize(ocs_b200_handle *h);
int ocs_multi_kernel_stats(ocs_b200_handle *h, int reset, double *ms_forward, double *ms_reverse, long *launches);
int ocs_multi_kernel_stats_ex(ocs_b200_handle *h, int reset, double *ms, int nms, long *launches);
int ocs_multi_set_profiling(ocs_b200_handle *h, int on);
int ocs_multi_evaluate_resident(ocs_b200_handle *h, const double *x0, long Bt, int want_gradient);
int ocs_multi_gather_results_dev(ocs_b200_handle *h, double *J_root, double *xT_root);
int ocs_multi_get_results(ocs_b200_handle *h, double *J, double *xT, int32_t *status, double *dJdv, double *lam0);
int ocs_multi_device_info(const ocs_b200_handle *h, int *ndev, int *devices, long *offsets, int cap);
