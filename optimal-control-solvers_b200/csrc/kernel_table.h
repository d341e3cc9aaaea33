// kernel_table.h -- the set of sweep kernels bound to one problem (host side).
// Entries are launchable through cudaLaunchKernel: either host symbols of kernels
// compiled into the library or cudaKernel_t handles of an NVRTC-compiled library.
#pragma once

struct ocs_kernel_table {
    int n, nc, np;
    int ntt[3];    // hoisted-table width per hoisting mode (HM_NONE, HM_TIME, HM_TIME_CONTROL)
    int kd;        // compile-time dense-basis size (0 = run-time K <= OCS_KDENSE_MAX)
    int threads;   // block size the kernels were built for
    int tile;      // RK4 steps per shared-memory tile
    int tpt_rev;   // trajectories per thread of the reverse (gradient) kernels
    int tpt_rev_coef;   // ... of the gradient kernels with per-trajectory coefficients (UMODE_COEF)
    int tpt_fwd;   // ... and of the forward kernels
    int tpt_fwd_coef;
    int tpt_fwd_pb, tpt_rev_pb;   // ... with per-trajectory parameters
    int fb_ring;   // ... and of the fb_sweep costate kernels (ring*(n + 2 nc)*threads doubles)
    int rev_ring;  // depth of the gradient kernels' cp.async node-state ring (shared memory: ring*tpt*(n [+ 2 nc])*threads doubles)
    int rev_ring_pb;   // ... of the variants with per-trajectory parameters
    int async_stage;   // multi-trajectory bodies keep two record tiles in shared memory (the next one is fetched with cp.async)
    const void *fwd[3][3][2];   // [UMODE][BMODE][params batched]
    const void *rev[3][3][2];
    const void *fwdg[3][3][2];  // fetch variants: also write the full grids (x with the J row; lam, dJdu)
    const void *revg[3][3][2];
    const void *fbrev[2];
    const void *fbcont[2];      // fb_sweep with the continuous costate equation (v1 adjointRHS) instead of the discrete adjoint
    const void *fbcontg[2];     // ... its fetch variant: writes the costate grid, leaves the control alone
    const void *tt;
    int nfold[3];               // weight-folded table entries per hoisting mode (0: the problem / build does not fold)
    const void *fold;           // fills the fold records [N][even(3*nfold)] from the grid-point and step records
    const void *eq;             // batched compute_equilibrium (one parameter set per thread)
};
