// kernels_builtin.cu -- instantiates the sweep kernels for the built-in problems.
// Compiled three times, so one library carries both arithmetic modes:
//   -DOCS_STRICT=1 -fmad=false                   namespace ocs_strict, kernel tag test_strict
//   -DOCS_STRICT=0 -DOCS_UNIFORM_H=1 -fmad=true  namespace ocs_fast,   tag test_fast   (grids whose steps agree to
//                                                rounding -- every linspace grid: step sizes as constant operands)
//   -DOCS_STRICT=0 -DOCS_UNIFORM_H=0 -fmad=true  namespace ocs_fastv,  tag test_fastv  (any other grid: per-step sizes)
#include <cuda_runtime.h>

#if OCS_STRICT
#define OCS_NS ocs_strict
#define OCS_TAGNAME test_strict
#elif defined(OCS_UNIFORM_H) && OCS_UNIFORM_H
#define OCS_NS ocs_fast
#define OCS_TAGNAME test_fast
#else
#define OCS_NS ocs_fastv
#define OCS_TAGNAME test_fastv
#endif

// Two trajectories per thread at 5 resident blocks/SM (92 registers) measured fastest for the test problem's
// reverse sweep on B200 (profiles/r1_notes.md); generated problems keep the defaults (one trajectory per
// thread, no register cap).
#ifndef OCS_MINBLOCKS_REV
#define OCS_MINBLOCKS_REV 5
#endif
#ifndef OCS_TPT_REV
#define OCS_TPT_REV 2
#endif
// fb_sweep costate kernel (HBM-bound, 40 B/step): occupancy is the lever.  Same-box A/B of (min blocks/SM,
// control prefetch depth): (1,1: 92 regs) 10.35 ms / (6,1) 9.71 / (7,2) 8.90 / (8,1: 64 regs) 8.38 / (8,2) 8.79 /
// (9,1) 8.56 / (10,1: spills) 8.79.
#ifndef OCS_MINBLOCKS_FB
#define OCS_MINBLOCKS_FB 8
#endif
#ifndef OCS_TPT_FWD
#define OCS_TPT_FWD 4
#endif
// Per-trajectory coefficients v (UMODE_COEF, what a batched single_shooting solve evaluates; bench.py --config cfg2v).
// Same-box A/B of (trajectories per thread, min blocks/SM): forward (3,1: 134 regs) 3.88 ms / (3,4) 3.32 / (2,6) 3.29 /
// (1,8: 64 regs) 2.91 / (1,12) 2.90; reverse (2,5: spills) 5.88 / (2,4: 126 regs) 5.70 / (1,8) 5.98 / (1,10) 5.97.
// (with the per-step window flag and folded weights, session 4: forward (1,8) 2.80 / (2,5: 94 regs) 2.69 ms)
#ifndef OCS_TPT_FWD_COEF
#define OCS_TPT_FWD_COEF 2
#endif
#ifndef OCS_MINBLOCKS_FWD_COEF
#define OCS_MINBLOCKS_FWD_COEF 5
#endif
#ifndef OCS_TPT_REV_COEF
#define OCS_TPT_REV_COEF 2
#endif
// (with the step sizes as constant operands, r3j: reverse (2,4: 128 regs) 4.125 ms / (2,5: 96 regs, 16 B spilled) 3.993;
//  forward (2,6: 76 regs) and the shared-control reverse at (2,6: 80 regs) unchanged within 0.4 %)
#ifndef OCS_MINBLOCKS_REV_COEF
#if !OCS_STRICT && defined(OCS_UNIFORM_H) && OCS_UNIFORM_H
#define OCS_MINBLOCKS_REV_COEF 5
#else
#define OCS_MINBLOCKS_REV_COEF 4      // per-step sizes need the registers: 40-52 B spilled at 5
#endif
#endif
// Per-trajectory parameters (a parameter sweep, bench.py --config cfg2p / cfg2pv): exp(-r t) runs in the thread and
// (c, m, r) live in registers, so the uniform-parameter settings above spilled (808 B in the grid-control forward
// kernel, 136-360 B in the reverse kernels).  Same-box A/B of (trajectories per thread, min blocks/SM) forward;reverse,
// cfg2p / cfg2pv in 1e10 trajectory-steps/s: inherited (4,-;2,5: spills) 6.59 / 5.50, (1,6;1,6) 7.99 / 6.69,
// (1,8;1,5) 8.00 / 7.06, (2,4;2,4) 7.72 / 6.57, (2,3;2,3) 7.70 / 6.47.  One trajectory per thread it is; (1,5;1,4) is
// the fastest setting without a single spilled byte in any _p1 variant (profiles/r2_notes.md).
#ifndef OCS_TPT_FWD_PB
#define OCS_TPT_FWD_PB 1
#endif
#ifndef OCS_MINBLOCKS_FWD_PB
#define OCS_MINBLOCKS_FWD_PB 5
#endif
#ifndef OCS_RING_PB
#define OCS_RING_PB 4       // cp.async node-state ring of the per-trajectory-parameter gradient kernels (r2u A/B: 5.32 -> 4.96 ms)
#endif
#ifndef OCS_TPT_REV_PB
#define OCS_TPT_REV_PB 1
#endif
#ifndef OCS_MINBLOCKS_REV_PB
#define OCS_MINBLOCKS_REV_PB 4
#endif
#ifndef OCS_MINBLOCKS_FB_PB
#define OCS_MINBLOCKS_FB_PB 3
#endif
#ifndef OCS_TILE
#define OCS_TILE 64         // RK4 steps per shared-memory tile: 32 / 64 / 128 measured 5.13 / 5.06 / 5.02 ms per cfg2 evaluation (fewer block barriers)
#endif
#include "rk4_kernels.cuh"
#include "problem_test.cuh"
#include "kernel_table.h"

#if OCS_STRICT
OCS_DEFINE_KERNELS(ocs_strict::TestOCProblem, test_strict, 0)
#elif defined(OCS_UNIFORM_H) && OCS_UNIFORM_H
OCS_DEFINE_KERNELS(ocs_fast::TestOCProblem, test_fast, 0)
#else
OCS_DEFINE_KERNELS(ocs_fastv::TestOCProblem, test_fastv, 0)
#endif

#define OCS_CAT5_(a, b, c, d, e) a##b##c##d##e
#define OCS_CAT5(a, b, c, d, e) OCS_CAT5_(a, b, c, d, e)
#define OCS_CAT2_(a, b) a##b
#define OCS_CAT2(a, b) OCS_CAT2_(a, b)
#define OCS_CAT3_(a, b, c) a##b##c
#define OCS_CAT3(a, b, c) OCS_CAT3_(a, b, c)
#define FWD(UM, BM, PB) (const void *)OCS_CAT5(ocs_fwd_, OCS_TAGNAME, _u##UM, _b##BM, _p##PB)
#define REV(UM, BM, PB) (const void *)OCS_CAT5(ocs_rev_, OCS_TAGNAME, _u##UM, _b##BM, _p##PB)
#define FWDG(UM, BM, PB) (const void *)OCS_CAT5(ocs_fwdg_, OCS_TAGNAME, _u##UM, _b##BM, _p##PB)
#define REVG(UM, BM, PB) (const void *)OCS_CAT5(ocs_revg_, OCS_TAGNAME, _u##UM, _b##BM, _p##PB)
#define FBR(PB) (const void *)OCS_CAT3(ocs_fbrev_, OCS_TAGNAME, _u2_b0_p##PB)
#define FBC(PB) (const void *)OCS_CAT3(ocs_fbcont_, OCS_TAGNAME, _u2_b0_p##PB)
#define FBCG(PB) (const void *)OCS_CAT3(ocs_fbcontg_, OCS_TAGNAME, _u2_b0_p##PB)

extern "C" void OCS_CAT2(ocs_builtin_table_, OCS_TAGNAME)(ocs_kernel_table *t)
{
    typedef OCS_NS::TestOCProblem P;
    t->n = P::n; t->nc = P::nc; t->np = P::np;
    for (int hm = 0; hm < 3; ++hm) t->ntt[hm] = P::ntt(hm);
    t->kd = 0;
    t->threads = OCS_THREADS;
    t->tile = OCS_TILE;
    t->tpt_rev = OCS_TPT_REV;
    t->tpt_rev_coef = OCS_TPT_REV_COEF;
    t->tpt_fwd = OCS_TPT_FWD;
    t->tpt_fwd_coef = OCS_TPT_FWD_COEF;
    t->tpt_fwd_pb = OCS_TPT_FWD_PB;
    t->tpt_rev_pb = OCS_TPT_REV_PB;
    t->rev_ring = OCS_RING;
    t->rev_ring_pb = OCS_RING_PB;
    t->async_stage = OCS_ASYNC_STAGE;
    t->fb_ring = OCS_FB_RING;
    for (int pb = 0; pb < 2; ++pb)
        for (int um = 0; um < 3; ++um)
            for (int bm = 0; bm < 3; ++bm)
                t->fwd[um][bm][pb] = t->rev[um][bm][pb] = t->fwdg[um][bm][pb] = t->revg[um][bm][pb] = nullptr;
    t->fwd[0][0][0] = FWD(0, 0, 0); t->fwd[0][0][1] = FWD(0, 0, 1);
    t->fwd[1][1][0] = FWD(1, 1, 0); t->fwd[1][1][1] = FWD(1, 1, 1);
    t->fwd[1][2][0] = FWD(1, 2, 0); t->fwd[1][2][1] = FWD(1, 2, 1);
    t->fwd[2][0][0] = FWD(2, 0, 0); t->fwd[2][0][1] = FWD(2, 0, 1);
    t->rev[0][0][0] = REV(0, 0, 0); t->rev[0][0][1] = REV(0, 0, 1);
    t->rev[0][1][0] = REV(0, 1, 0); t->rev[0][1][1] = REV(0, 1, 1);
    t->rev[1][1][0] = REV(1, 1, 0); t->rev[1][1][1] = REV(1, 1, 1);
    t->rev[2][0][0] = REV(2, 0, 0); t->rev[2][0][1] = REV(2, 0, 1);
    t->fwdg[0][0][0] = FWDG(0, 0, 0); t->fwdg[0][0][1] = FWDG(0, 0, 1);
    t->fwdg[1][1][0] = FWDG(1, 1, 0); t->fwdg[1][1][1] = FWDG(1, 1, 1);
    t->fwdg[1][2][0] = FWDG(1, 2, 0); t->fwdg[1][2][1] = FWDG(1, 2, 1);
    t->fwdg[2][0][0] = FWDG(2, 0, 0); t->fwdg[2][0][1] = FWDG(2, 0, 1);
    t->revg[0][0][0] = REVG(0, 0, 0); t->revg[0][0][1] = REVG(0, 0, 1);
    t->revg[1][1][0] = REVG(1, 1, 0); t->revg[1][1][1] = REVG(1, 1, 1);
    t->revg[1][2][0] = REVG(1, 2, 0); t->revg[1][2][1] = REVG(1, 2, 1);
    t->revg[2][0][0] = REVG(2, 0, 0); t->revg[2][0][1] = REVG(2, 0, 1);
    t->fbrev[0] = FBR(0); t->fbrev[1] = FBR(1);
    t->fbcont[0] = FBC(0); t->fbcont[1] = FBC(1);
    t->fbcontg[0] = FBCG(0); t->fbcontg[1] = FBCG(1);
    t->tt = (const void *)OCS_CAT2(ocs_tt_, OCS_TAGNAME);
    for (int hm = 0; hm < 3; ++hm) t->nfold[hm] = OCS_NS::ocs_nfold<P>(hm);
    t->fold = (const void *)OCS_CAT2(ocs_fold_, OCS_TAGNAME);
    t->eq = (const void *)OCS_CAT2(ocs_eq_, OCS_TAGNAME);
}
