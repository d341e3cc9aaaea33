// kernel_table.h -- the set of sweep kernels bound to one problem (host side).
// Entries are launchable through cudaLaunchKernel: either host symbols of kernels
// compiled into the library or cudaKernel_t handles of an NVRTC-compiled library.
#pragma once

struct ocs_kernel_table {
    int n, nc, np, ntt;
    int kd;        // compile-time dense-basis size (0 = run-time K <= OCS_KDENSE_MAX)
    int threads;   // block size the kernels were built for
    const void *fwd[3][3][2];   // [UMODE][BMODE][params batched]
    const void *rev[3][3][2];
    const void *fbrev[2];
    const void *tt;
};
