// kernels_builtin.cu -- instantiates the sweep kernels for the built-in problems.
// Compiled twice: -DOCS_STRICT=0 (FMA contraction on, namespace ocs_fast, kernel
// tag test_fast) and -DOCS_STRICT=1 -fmad=false (namespace ocs_strict, tag
// test_strict), so one library carries both arithmetic modes.
#include <cuda_runtime.h>

#if OCS_STRICT
#define OCS_NS ocs_strict
#define OCS_TAG(x) x##_strict
#else
#define OCS_NS ocs_fast
#define OCS_TAG(x) x##_fast
#endif

#include "rk4_kernels.cuh"
#include "problem_test.cuh"
#include "kernel_table.h"

#if OCS_STRICT
OCS_DEFINE_KERNELS(ocs_strict::TestOCProblem, test_strict, 0)
#define T(name) (const void *)name##test_strict
#else
OCS_DEFINE_KERNELS(ocs_fast::TestOCProblem, test_fast, 0)
#define T(name) (const void *)name##test_fast
#endif

#define FWD(UM, BM, PB) (const void *)OCS_CAT5(ocs_fwd_, OCS_TAGNAME, _u##UM, _b##BM, _p##PB)

#if OCS_STRICT
#define OCS_TAGNAME test_strict
#else
#define OCS_TAGNAME test_fast
#endif
#define OCS_CAT5_(a, b, c, d, e) a##b##c##d##e
#define OCS_CAT5(a, b, c, d, e) OCS_CAT5_(a, b, c, d, e)
#define OCS_CAT3_(a, b, c) a##b##c
#define OCS_CAT3(a, b, c) OCS_CAT3_(a, b, c)
#define REV(UM, BM, PB) (const void *)OCS_CAT5(ocs_rev_, OCS_TAGNAME, _u##UM, _b##BM, _p##PB)
#define FBR(PB) (const void *)OCS_CAT3(ocs_fbrev_, OCS_TAGNAME, _p##PB)

static void fill(ocs_kernel_table *t)
{
    t->n = OCS_NS::TestOCProblem::n;
    t->nc = OCS_NS::TestOCProblem::nc;
    t->np = OCS_NS::TestOCProblem::np;
    t->ntt = OCS_NS::TestOCProblem::ntt;
    t->kd = 0;
    t->threads = OCS_THREADS;
    for (int pb = 0; pb < 2; ++pb)
        for (int um = 0; um < 3; ++um)
            for (int bm = 0; bm < 3; ++bm) t->fwd[um][bm][pb] = t->rev[um][bm][pb] = nullptr;
    t->fwd[0][0][0] = FWD(0, 0, 0); t->fwd[0][0][1] = FWD(0, 0, 1);
    t->fwd[1][1][0] = FWD(1, 1, 0); t->fwd[1][1][1] = FWD(1, 1, 1);
    t->fwd[1][2][0] = FWD(1, 2, 0); t->fwd[1][2][1] = FWD(1, 2, 1);
    t->fwd[2][0][0] = FWD(2, 0, 0); t->fwd[2][0][1] = FWD(2, 0, 1);
    t->rev[0][0][0] = REV(0, 0, 0); t->rev[0][0][1] = REV(0, 0, 1);
    t->rev[0][1][0] = REV(0, 1, 0); t->rev[0][1][1] = REV(0, 1, 1);
    t->rev[0][2][0] = REV(0, 2, 0); t->rev[0][2][1] = REV(0, 2, 1);
    t->rev[1][1][0] = REV(1, 1, 0); t->rev[1][1][1] = REV(1, 1, 1);
    t->rev[1][2][0] = REV(1, 2, 0); t->rev[1][2][1] = REV(1, 2, 1);
    t->rev[2][0][0] = REV(2, 0, 0); t->rev[2][0][1] = REV(2, 0, 1);
    t->fbrev[0] = FBR(0); t->fbrev[1] = FBR(1);
    t->tt = (const void *)OCS_CAT3(ocs_tt_, OCS_TAGNAME, );
}

#if OCS_STRICT
extern "C" void ocs_builtin_table_test_strict(ocs_kernel_table *t) { fill(t); }
#else
extern "C" void ocs_builtin_table_test_fast(ocs_kernel_table *t) { fill(t); }
#endif
