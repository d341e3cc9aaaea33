// rk4_kernels.cuh -- sm_100a sweep kernels of the B200-native RK4 trajectory path.
//
// One trajectory per thread.  The small augmented state [x; J], the costate and
// the RK4 stage values live in registers; node states needed by the reverse sweep
// are stored in HBM in structure-of-arrays layout [component][step][batch] so that
// a warp's 32 trajectories write/read 256 contiguous bytes.  Everything that is the
// same for every trajectory of the batch (times, step sizes, hoisted uniform
// sub-expressions, shared control values, control basis entries) is packed into
// fixed-width per-grid-point records and staged through shared memory tile by tile;
// record widths and the tile size are compile-time constants so that every
// shared-memory operand is an immediate offset.
//
// Arithmetic follows, operation by operation,
//   Integrator/RK4Integrator.m:28-56   compute_states
//   Integrator/RK4Integrator.m:59-94   compute_adjoints
//   Integrator/RK4Integrator.m:97-121  compute_dJdu
//   Control/*Control.m                 compute_u (v*B) and compute_dJdv (dJdu*B')
// of the reference (paths relative to the reference tree).
//
// The file is compiled three ways: by nvcc with FMA contraction (namespace
// ocs_fast), by nvcc with -fmad=false (namespace ocs_strict) and by NVRTC at run
// time for generated problems (namespace ocs_gen).  It therefore only uses what
// NVRTC provides without host headers.
#pragma once

#ifndef OCS_NS
#error "define OCS_NS before including rk4_kernels.cuh"
#endif

#ifndef OCS_THREADS
#define OCS_THREADS 128
#endif
#ifndef OCS_TILE
#define OCS_TILE 32         // RK4 steps per shared-memory tile
#endif
#ifndef OCS_TILE_SINGLE
#define OCS_TILE_SINGLE (OCS_TILE < 32 ? OCS_TILE : 32)   // ... of the one-trajectory bodies (fetch and fb_sweep kernels); never
#endif                                                    // larger than OCS_TILE, which sizes the host's allocation
#define OCS_XS_SLACK 8      // nodes of slack in front of the node-state store (>= any OCS_PREFETCH)
#ifndef OCS_PREFETCH
#define OCS_PREFETCH 2      // node states prefetched ahead by the reverse sweep (slack in front of the store)
#endif
#ifndef OCS_REV_UNROLL
#define OCS_REV_UNROLL 2    // unroll factor of the reverse step loop
#endif
#ifndef OCS_TPT_FWD
#define OCS_TPT_FWD 1       // trajectories per thread of the forward kernels
#endif
#ifndef OCS_MINBLOCKS_FWD
#define OCS_MINBLOCKS_FWD 1
#endif
#ifndef OCS_TPT_REV
#define OCS_TPT_REV 1       // trajectories per thread of the reverse (gradient) kernels
#endif
#ifndef OCS_MINBLOCKS_REV
#define OCS_MINBLOCKS_REV 1 // __launch_bounds__ min blocks/SM of the reverse kernels (register cap)
#endif
#ifndef OCS_TPT_FWD_COEF
#define OCS_TPT_FWD_COEF OCS_TPT_FWD              // trajectories per thread of the forward kernels with per-trajectory v
#endif
#ifndef OCS_MINBLOCKS_FWD_COEF
#define OCS_MINBLOCKS_FWD_COEF OCS_MINBLOCKS_FWD
#endif
#ifndef OCS_TPT_REV_COEF
#define OCS_TPT_REV_COEF OCS_TPT_REV              // ... of the gradient kernels with per-trajectory coefficients v
#endif
#ifndef OCS_MINBLOCKS_REV_COEF
#define OCS_MINBLOCKS_REV_COEF OCS_MINBLOCKS_REV
#endif
#ifndef OCS_MINBLOCKS_FB
#define OCS_MINBLOCKS_FB 1  // __launch_bounds__ min blocks/SM of the fb_sweep costate kernels
#endif
// Per-trajectory parameters (a parameter sweep: nothing is batch-uniform, exp() and the like run in the thread and
// every trajectory carries its parameters in registers) want their own trajectories-per-thread / register cap:
// the values tuned for uniform parameters spilled (808 B in the grid-control forward kernel).  Defaults: inherit.
#ifndef OCS_TPT_FWD_PB
#define OCS_TPT_FWD_PB OCS_TPT_FWD
#endif
#ifndef OCS_MINBLOCKS_FWD_PB
#define OCS_MINBLOCKS_FWD_PB OCS_MINBLOCKS_FWD
#endif
#ifndef OCS_TPT_REV_PB
#define OCS_TPT_REV_PB OCS_TPT_REV
#endif
#ifndef OCS_MINBLOCKS_REV_PB
#define OCS_MINBLOCKS_REV_PB OCS_MINBLOCKS_REV
#endif
#ifndef OCS_MINBLOCKS_FB_PB
#define OCS_MINBLOCKS_FB_PB OCS_MINBLOCKS_FB
#endif
#ifndef OCS_UPREFETCH
#define OCS_UPREFETCH 1     // steps of per-trajectory grid control loaded ahead by reverse_body (UMODE_GRID)
#endif
#ifndef OCS_FB_RING
#define OCS_FB_RING 0       // fb_sweep costate kernels: node state and the two grid controls of a step are fetched this
                            // many steps ahead by cp.async into a per-thread shared-memory ring; 0 = register prefetch
#endif
#ifndef OCS_RING
#define OCS_RING 0          // reverse_multi_body: node states are fetched this many steps ahead by cp.async into a
                            // per-thread shared-memory ring (power of two, <= OCS_XS_SLACK); 0 = register prefetch.
                            // Measured on B200 (profiles/r1_notes.md): 0: 3.71 ms, 2: 3.90, 4: 3.75, 8: 3.76 -- the
                            // FP64-bound gradient kernel does not gain from it, the HBM-bound fb_sweep kernel does.
#endif
#ifndef OCS_RING_PB
#define OCS_RING_PB OCS_RING  // ... the same for the variants with per-trajectory parameters: their reverse step is long
                              // (nothing is tabulated) and their top stall was the node-state load (long_scoreboard
                              // 2.8 warps per issue); ring of 4: 5.32 -> 4.96 ms on cfg2p, while the uniform-parameter
                              // kernels lose 5 % with it (profiles/r2_notes.md)
#endif
// How the sweeps with per-trajectory coefficients move the sparse-basis control window, per sweep direction:
//   0  compare the knot index at every grid point, reload under a branch
//   1  ... reload by predicated loads (ld_global_if)
//   2  ... branch on a warp vote
//   3  test ONE host-computed flag per step (bits 3 / 4 of the midpoint's basis record: the window moves within this
//      step) and only then run the compares and reloads -- the common step carries no window instructions at all
// and whether the forward sweep with per-trajectory coefficients folds the stage weights.  Same-box A/B on cfg2v
// (forward / reverse ms): mode 0 without folding 2.91 / 5.59; 0 + fold 3.16 / 6.01; 1 + fold 3.74 / 5.24 (4.97 with
// 64-step tiles); 2: 2.89 / 6.02; 3 + fold: 2.80 / 4.29, and 2.69 with two trajectories per thread in the forward
// sweep (profiles/r1_notes.md).
#ifndef OCS_SLIDE_PRED_FWD
#define OCS_SLIDE_PRED_FWD 3
#endif
#ifndef OCS_SLIDE_PRED_REV
#define OCS_SLIDE_PRED_REV 3
#endif
#ifndef OCS_FOLD_FWD_COEF
#define OCS_FOLD_FWD_COEF 1
#endif
#ifndef OCS_PRED_CLOBBER
#define OCS_PRED_CLOBBER 1
#endif
#ifndef OCS_KDENSE_MAX
#define OCS_KDENSE_MAX 32   // run-time K limit of the dense-basis kernels (compile-time K has none)
#endif
#define OCS_NPC 16          // batch-uniform problem parameters travel inside SweepArgs (kernel-parameter constant bank) up
                            // to this count: an FP64 instruction takes them as a constant operand, so they cost neither
                            // registers nor loads (6 parameters of the generated 6-state problem were 12 registers)
#ifndef OCS_ASYNC_STAGE
#define OCS_ASYNC_STAGE 1   // multi-trajectory bodies: the record tile of the NEXT block of steps is fetched with cp.async
                            // into a second shared-memory buffer while the current one is consumed (one block barrier per
                            // tile instead of two, no exposed load latency at the tile switch -- with the 8 resident warps
                            // per SM of a generated 6-state problem the synchronous staging was 12 % of the forward sweep)
#endif

namespace OCS_NS {

// How the control reaches a trajectory.
enum { UMODE_UNIFORM = 0,   // u_j shared by the batch, read from the grid-point record
       UMODE_COEF = 1,      // per-trajectory coefficients v, evaluated through the basis
       UMODE_GRID = 2 };    // per-trajectory values on the G grid points
// How dJdu is folded into dJdv (and how UMODE_COEF evaluates u).
enum { BMODE_NONE = 0, BMODE_SPARSE = 1, BMODE_DENSE = 2 };
// Hoisting mode: which sub-expressions of the right-hand sides are batch-uniform and therefore
// evaluated once per grid point (into the record's table slots) instead of once per trajectory.
enum { HM_NONE = 0,         // per-trajectory parameters: nothing is uniform
       HM_TIME = 1,         // uniform parameters: functions of (t, p)
       HM_TIME_CONTROL = 2 };   // uniform parameters and control: functions of (t, p, u)

__host__ __device__ constexpr int ocs_even(int x) { return (x + 1) & ~1; }
// grid-point record: [ t | u(0..nc-1) | table(0..ntt-1) | pad ]
__host__ __device__ constexpr int ocs_rec_width(int nc, int ntt) { return ocs_even(1 + nc + ntt); }
// basis record: sparse [ (int k, int two) | w0 | w1 | pad ], dense [ B(0..K-1, j) | pad ]
__host__ __device__ constexpr int ocs_basis_width(int bmode, int K) { return bmode == 1 ? 4 : (bmode == 2 ? ocs_even(K) : 0); }
__host__ __device__ constexpr int ocs_hoist_mode(int umode, bool params_batched)
{
    return params_batched ? 0 : (umode == 0 ? 2 : 1);
}

// Weight folding (see problem_test.cuh): a problem may declare `nfold` table entries that meet the RK4 stage
// weight of the objective row; they are multiplied by that weight once per step and stage position into the
// fold records [N][ocs_fold_width] = [ L: fold(tt_L, h/6) | M: fold(tt_M, h/3) | R: fold(tt_R, h/6) | pad ] and
// the sweeps call the problem's Ff / Xf / Uf.  Only in the FMA-contracted build (it re-associates products),
// only when a table exists (uniform parameters), and in the adjoint sweep only while lam(end) == 1.
template <bool B>
struct ocs_bool { static constexpr bool value = B; };
template <class P, class Tag>
struct ocs_dep { typedef P type; };      // P, made dependent on Tag (calls in discarded if-constexpr branches of generic lambdas)
template <class P>
struct ocs_fold_traits {
    template <class Q> __host__ __device__ static constexpr int get(decltype(Q::nfold) *) { return Q::nfold; }
    template <class Q> __host__ __device__ static constexpr int get(...) { return 0; }
    static constexpr int value = get<P>(nullptr);
};
template <class P>
__host__ __device__ constexpr int ocs_nfold(int hm)
{
#if (defined(OCS_STRICT) && OCS_STRICT) || defined(OCS_NOFOLD)
    return 0;
#else
    return hm == 0 ? 0 : ocs_fold_traits<P>::value;
#endif
}
__host__ __device__ constexpr int ocs_fold_width(int nfold) { return ocs_even(3 * nfold); }

// Time exponentials by recurrence (per-trajectory parameters, FMA-contracted build only).  With one parameter set per
// trajectory nothing can be tabulated per grid point, and a problem like the bundled one evaluates exp(-r t) in every
// right-hand side: ~40 FP64 instructions each, 4 per forward and 8 per reverse step -- the parameter-sweep kernels
// were spending 3/4 of their FP64 issue slots inside exp().  A problem may therefore declare `ntexp` exponentials
// E_k(t) = exp(a_k t) (texp_rates) and how its HM_TIME table follows from them (table_from_texp).  The multi-trajectory
// sweeps then keep E_k per trajectory and advance it by one multiplication per grid point, E(t + h/2) = E(t) exp(a h/2),
// re-anchored with a real exp() at every tile of steps (<= 2*OCS_TILE products in a row: relative error below 3e-14, well
// inside the 1e-10 parity bound; the strict build and the single-trajectory fetch / fb_sweep kernels keep exp()).
// On a grid whose steps differ by more than rounding the factor is recomputed every step (one exp instead of four).
template <class P>
struct ocs_texp_traits {
    template <class Q> __host__ __device__ static constexpr int get(decltype(Q::ntexp) *) { return Q::ntexp; }
    template <class Q> __host__ __device__ static constexpr int get(...) { return 0; }
    static constexpr int value = get<P>(nullptr);
};
template <class P>
__host__ __device__ constexpr int ocs_ntexp(bool params_batched)
{
#if (defined(OCS_STRICT) && OCS_STRICT) || defined(OCS_NOTEXP)
    return 0;
#else
    return params_batched ? ocs_texp_traits<P>::value : 0;
#endif
}

// Argument block of one sweep over one RK4 segment (main horizon or tail).
// SoA device arrays have leading dimension ldb (the batch stride); the kernel
// handles trajectories [0, nb) of them, so a sub-range is a pointer offset.
struct SweepArgs {
    int N;                  // RK4 steps
    int K;                  // basis size
    int RB;                 // doubles per grid-point record (cross-checked against the kernel's constant)
    int BW;                 // doubles per basis record
    long nb;                // trajectories handled
    long ldb;               // batch stride of SoA arrays
    long ldp;               // batch stride of per-trajectory params
    const double *gprec;    // [2N+1][RB]
    const double *steprec;  // [N][4]     h, h/2, h/3, h/6
    const double *foldrec;  // [N][ocs_fold_width(nfold)] weight-folded table entries, or null (problem without fold)
    double hc[4];           // OCS_UNIFORM_H builds (grids whose steps agree to rounding, FMA-contracted build): h, h/2, h/3, h/6 of
                            // the segment's mean step, read as constant-bank operands.  A DFMA that reads three vector
                            // registers occupies the FP64 pipe for 3 cycles, one with a constant operand for 2
                            // (profiles/r2_notes.md): the step sizes are operands of a quarter of the sweeps' FMAs
    int texp_uniform;       // the segment's steps agree to rounding: the half-step factor of a time exponential is
                            // computed once per tile instead of once per step (ocs_ntexp)
    int lamJ_one;           // reverse sweep: lam(end) == 1 for every trajectory (lamT null, or chained from a tail
                            // sweep that started from e_m), so the folded adjoint chain applies
    const double *basrec;   // [2N+1][BW]
    const double *x0;       // [n][ldb]
    const double *p;        // [np] or [np][ldp]
    double pc[OCS_NPC];     // batch-uniform parameters by value (np <= OCS_NPC), read as constant-bank operands
    const double *v;        // UMODE_COEF: [(k*nc+c)][ldb]
    double *ug;             // UMODE_GRID: [nc][2N+1][ldb]   (fb_sweep updates it in place)
    double *ug_out;         // dense UMODE_COEF forward sweep: the evaluated control grid is kept here for the
                            // reverse sweep ([nc][2N+1][ldb]), or null
    double *dug;            // BMODE_NONE reverse sweep: dJdu in SoA layout [nc][2N+1][ldb], or null
    double *xs;             // node-state store [n][N+1][ldb]; OCS_XS_SLACK*ldb doubles of slack precede it so
                            // that the reverse sweep's prefetch never needs a bounds check
    // forward outputs
    double *J;              // [nb]
    int J_accumulate;       // J = J_in + J_segment (RK4InfiniteIntegrator.m:23)
    double *xT;             // [n][ldb]
    int *status;            // [nb]
    double *xgrid;          // MATLAB layout m x (N+1) x nb, or null
    // reverse inputs / outputs
    const double *lamT;     // [m][ldl] or null (= e_m, RK4Integrator.m:63-66)
    long ldl;               // batch stride of lamT; 0 = shared column
    double *dJdv;           // [(k*nc+c)][ldb], zero-initialised, or null
    int dJdv_store;         // 1: every knot is flushed exactly once (knot index monotone along the grid), so a
                            // finished accumulator is stored; 0: added to what the buffer holds
    double *lam0;           // [m][ldb] or null
    double *lamgrid;        // MATLAB layout m x (N+1) x nb, or null
    double *dJdugrid;       // MATLAB layout nc x (2N+1) x nb, or null
    double *lams;           // fetch variant: node costates (state rows) in SoA layout [n][N+1][ldlams], or null
    long ldlams;
    // fb_sweep (reverse sweep with control update)
    const double *bounds;   // [nc][2]
    double uRelTol, uAbsTol, relax;
    double *max_change;     // [nb]
    const int *frozen;      // [nb] trajectories already converged (skipped), or null
    int fb_readonly;        // fb_sweep kernels used as a costate fetch: the control grid is read, never updated
    int *error_flag;        // set to 1 when the host's record layout disagrees with the kernel's
};

__device__ __forceinline__ size_t soa(long row, long ld, long b) { return (size_t)row * (size_t)ld + (size_t)b; }
// Row `row` of an SoA array whose batch offset is already folded into `base`.  Row index and byte stride are
// 32-bit (the host refuses batches of 2^29 trajectories or more), so the address is ONE IMAD.WIDE.U32 instead of
// a 64x64-bit multiply-add chain -- the per-step grid accesses of the fb_sweep kernels were spending more issue
// slots on that chain than on arithmetic.
__device__ __forceinline__ double *soa_row(double *base, unsigned row, unsigned ld8)
{
    return reinterpret_cast<double *>(reinterpret_cast<char *>(base) + (size_t)row * ld8);
}
__device__ __forceinline__ const double *soa_row(const double *base, unsigned row, unsigned ld8)
{
    return reinterpret_cast<const double *>(reinterpret_cast<const char *>(base) + (size_t)row * ld8);
}

// The first slot of a sparse basis record packs two ints (knot index k, flags) into one double; they are
// taken apart with a bit cast of the loaded double, never through an int-typed pointer (no aliasing games).
// dst = *p if pred (a predicated load, never a branch)
__device__ __forceinline__ void ld_global_if(double &dst, const double *p, bool pred)
{
#if OCS_PRED_CLOBBER
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.f64 %0, [%1];\n\t}"
                 : "+d"(dst) : "l"(p), "r"((unsigned)pred) : "memory");
#else
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.f64 %0, [%1];\n\t}"
        : "+d"(dst) : "l"(p), "r"((unsigned)pred));
#endif
}
__device__ __forceinline__ int rec_k(const double *bas) { return (int)(__double_as_longlong(bas[0]) & 0xffffffffLL); }
__device__ __forceinline__ int rec_flags(const double *bas) { return (int)(__double_as_longlong(bas[0]) >> 32); }

// ------------------------------------------------------------------ control evaluation
// Evaluates u(:,j) for one trajectory while the sweep walks the grid monotonically
// (up in the forward sweep, down in the reverse sweep).
template <int nc, int UMODE, int BMODE, int KD>
struct ControlEval {
    static constexpr int KA = (KD > 0 ? KD : OCS_KDENSE_MAX);
    int kcur;
    const double *ugb;      // UMODE_GRID: a.ug + trajectory
    unsigned ld8;           // byte stride between grid rows
    // sparse basis: the window {k, k+1} of live coefficients, held by knot parity so that sliding the
    // window replaces one slot and never shuffles registers
    double vE[(UMODE == UMODE_COEF && BMODE == BMODE_SPARSE) ? nc : 1];
    double vO[(UMODE == UMODE_COEF && BMODE == BMODE_SPARSE) ? nc : 1];
    double vd[(UMODE == UMODE_COEF && BMODE == BMODE_DENSE) ? KA * nc : 1];

    __device__ __forceinline__ void init(const SweepArgs &a, long bb)
    {
        kcur = -0x40000000;
#pragma unroll
        for (int c = 0; c < ((UMODE == UMODE_COEF && BMODE == BMODE_SPARSE) ? nc : 1); ++c) vE[c] = vO[c] = 0.0;
        ugb = UMODE == UMODE_GRID ? a.ug + bb : nullptr;
        ld8 = (unsigned)a.ldb * 8u;
        if constexpr (UMODE == UMODE_COEF && BMODE == BMODE_DENSE) {
            const int K = KD > 0 ? KD : a.K;
#pragma unroll
            for (int k = 0; k < KA; ++k)
                if (k < K) {
#pragma unroll
                    for (int c = 0; c < nc; ++c) vd[k * nc + c] = a.v[soa(k * nc + c, a.ldb, bb)];
                }
        }
    }

    // Sparse basis, part 1: make {k, k+1} the live window of grid point j.  The branch is uniform across the
    // batch (every trajectory shares the grid) and taken ~K times per sweep.
    template <int PRED = 0>     // 0: branch, 1: predicated loads, 2: branch on a warp vote (provably uniform)
    __device__ __forceinline__ void slide(const SweepArgs &a, const double *bas, long bb)
    {
        const int k = rec_k(bas);
        if constexpr (PRED == 2) {
            if (__any_sync(0xffffffffu, k != kcur)) {
                const int ke = (k & 1) ? k + 1 : k, ko = (k & 1) ? k : k + 1;
#pragma unroll
                for (int c = 0; c < nc; ++c) {
                    vE[c] = (ke < a.K) ? a.v[soa((long)ke * nc + c, a.ldb, bb)] : 0.0;
                    vO[c] = (ko < a.K) ? a.v[soa((long)ko * nc + c, a.ldb, bb)] : 0.0;
                }
                kcur = k;
            }
            return;
        }
        if constexpr (PRED == 1) {
        // Predicated loads instead of a branch: the move is uniform across the batch, but the compiler cannot
        // know that and may wrap a branch in BSSY/BSYNC, splitting the arithmetic of the step into basic blocks
        // (measured: +8 % on the forward sweep).  A slot whose knot lies outside the basis keeps its previous
        // (finite; 0 at the start) value and meets weight 0 in value().
        const bool mv = k != kcur;
        const int ke = (k & 1) ? k + 1 : k, ko = (k & 1) ? k : k + 1;
        const unsigned ld8v = (unsigned)a.ldb * 8u;
#pragma unroll
        for (int c = 0; c < nc; ++c) {
            ld_global_if(vE[c], soa_row(a.v + bb, (unsigned)(ke * nc + c), ld8v), mv && ke < a.K);
            ld_global_if(vO[c], soa_row(a.v + bb, (unsigned)(ko * nc + c), ld8v), mv && ko < a.K);
        }
        kcur = k;
        return;
        }
        if (k != kcur) {
            const int ke = (k & 1) ? k + 1 : k, ko = (k & 1) ? k : k + 1;
#pragma unroll
            for (int c = 0; c < nc; ++c) {
                vE[c] = (ke < a.K) ? a.v[soa((long)ke * nc + c, a.ldb, bb)] : 0.0;
                vO[c] = (ko < a.K) ? a.v[soa((long)ko * nc + c, a.ldb, bb)] : 0.0;
            }
            kcur = k;
        }
    }
    // Sparse basis, part 2: u = v_k B(k,j) + v_{k+1} B(k+1,j) with the two weights by knot parity, branch-free
    // so that the chains of several trajectories share a basic block.  An absent entry has weight 0 and adds an
    // exact zero -- MATLAB's v*B multiplies by those zeros as well -- so a PWConstantControl value stays an exact
    // copy of its coefficient; with separately rounded products (strict mode) the sum is commutative, i.e. the
    // reference's summation order.
    __device__ __forceinline__ void value(const double *bas, double *u) const
    {
        const double wE = bas[1], wO = bas[2];
#pragma unroll
        for (int c = 0; c < nc; ++c) u[c] = vE[c] * wE + vO[c] * wO;
    }

    // u(:,j) = reshape(v,nc,[]) * B(:,j)   (PWLinearControl.m:59-62 and twins)
    //   rec: grid-point record, bas: basis record of grid point j
    __device__ __forceinline__ void eval(const SweepArgs &a, const double *rec, const double *bas, int j, long bb,
                                         double *u)
    {
        if constexpr (UMODE == UMODE_UNIFORM) {
#pragma unroll
            for (int c = 0; c < nc; ++c) u[c] = rec[1 + c];
        } else if constexpr (UMODE == UMODE_GRID) {
            const int G = 2 * a.N + 1;
#pragma unroll
            for (int c = 0; c < nc; ++c) u[c] = *soa_row(ugb, (unsigned)(c * G + j), ld8);
        } else if constexpr (BMODE == BMODE_SPARSE) {
            slide(a, bas, bb);
            value(bas, u);
        } else {
            const int K = KD > 0 ? KD : a.K;
#pragma unroll
            for (int c = 0; c < nc; ++c) u[c] = vd[c] * bas[0];
#pragma unroll
            for (int k = 1; k < KA; ++k)
                if (k < K) {
                    const double bk = bas[k];
#pragma unroll
                    for (int c = 0; c < nc; ++c) u[c] = u[c] + vd[k * nc + c] * bk;
                }
        }
    }
};

// dJdv = dJdu * B'  (PWLinearControl.m:53-56 and twins) accumulated while the reverse
// sweep walks the grid downwards; dJdu itself is never materialised.
template <int nc, int BMODE, int KD>
struct GradAcc {
    static constexpr int KA = (KD > 0 ? KD : OCS_KDENSE_MAX);
    int kcur;
    // sparse basis: accumulators of the two live knots {k, k+1}, held by knot parity (aE: even knot)
    double aE[BMODE == BMODE_SPARSE ? nc : 1];
    double aO[BMODE == BMODE_SPARSE ? nc : 1];
    // (a dense basis is not folded in the sweep: the reverse kernel writes dJdu and a separate projection
    //  kernel forms dJdu*B', see dense_project_kernel in ocs_api.cu)

    __device__ __forceinline__ void init()
    {
        kcur = -0x40000000;
        if constexpr (BMODE == BMODE_SPARSE) {
#pragma unroll
            for (int c = 0; c < nc; ++c) aE[c] = aO[c] = 0.0;
        }
    }

    // writes the finished knot k out and clears its accumulator (explicit branches on the knot parity: the
    // accumulators must stay in registers, so they are never selected through a run-time pointer)
    __device__ __forceinline__ void flush(const SweepArgs &a, int k, long b, bool active)
    {
        const bool write = k >= 0 && k < a.K && active;
        if (k & 1) {
            if (write) {
#pragma unroll
                for (int c = 0; c < nc; ++c) {
                    const size_t idx = soa((long)k * nc + c, a.ldb, b);
                    // one owner thread per column.  A plain store when the knot cannot come back: the load of the
                    // read-modify-write form stalled every warp of the block at the same step (long scoreboard)
                    a.dJdv[idx] = a.dJdv_store ? aO[c] : a.dJdv[idx] + aO[c];
                }
            }
#pragma unroll
            for (int c = 0; c < nc; ++c) aO[c] = 0.0;
        } else {
            if (write) {
#pragma unroll
                for (int c = 0; c < nc; ++c) {
                    const size_t idx = soa((long)k * nc + c, a.ldb, b);
                    a.dJdv[idx] = a.dJdv_store ? aE[c] : a.dJdv[idx] + aE[c];
                }
            }
#pragma unroll
            for (int c = 0; c < nc; ++c) aE[c] = 0.0;
        }
    }

    __device__ __forceinline__ void add(const SweepArgs &a, const double *bas, const double *d, long b, bool active)
    {
        if constexpr (BMODE == BMODE_SPARSE) {
            const int k = rec_k(bas);
            const double wE = bas[1], wO = bas[2];
            if (k != kcur) {                    // uniform across the block; ~K times per sweep
                flush(a, kcur + 1, b, active);  // the window slides down: knot kcur+1 is complete
                if (k != kcur - 1) flush(a, kcur, b, active);
                kcur = k;
            }
#pragma unroll
            for (int c = 0; c < nc; ++c) {
                aE[c] = aE[c] + d[c] * wE;
                aO[c] = aO[c] + d[c] * wO;     // a zero weight adds an exact zero
            }
        }
    }

    // sparse basis, window known not to move at this grid point
    __device__ __forceinline__ void add_fast(const double *bas, const double *d)
    {
        if constexpr (BMODE == BMODE_SPARSE) {
            const double wE = bas[1], wO = bas[2];
#pragma unroll
            for (int c = 0; c < nc; ++c) {
                aE[c] = aE[c] + d[c] * wE;
                aO[c] = aO[c] + d[c] * wO;
            }
        }
    }

    __device__ __forceinline__ void finish(const SweepArgs &a, long b, bool active)
    {
        if constexpr (BMODE == BMODE_SPARSE) {
            flush(a, kcur + 1, b, active);
            flush(a, kcur, b, active);
        }
    }
};

// Running maximum of ratios num/den (den >= 0) without a division per sample: the current maximiser is
// kept as a (num, den) pair and compared by cross-multiplication; one division at the end.  A zero
// denominator with a positive numerator is +inf and dominates; 0/0 samples are NaN in the reference's
// formula and are ignored here like fmax ignores NaN.
struct MaxRatio {
    double num = 0.0, den = 1.0;
    __device__ __forceinline__ void update(double n_, double d_)
    {
        if (n_ * den > num * d_) { num = n_; den = d_; }
    }
    __device__ __forceinline__ double value() const { return num / den; }
};

// ------------------------------------------------------------------ asynchronous per-thread prefetch (cp.async)
// The reverse sweep consumes one node state per trajectory and step.  Prefetching it into registers ties the
// prefetch distance to the register rotation of the unrolled loop (the rotation moves wait on the load);
// cp.async (LDGSTS) writes it into a shared-memory slot owned by the thread instead, any number of steps ahead,
// and the thread only waits (cp.async.wait_group) for the slot it is about to read.
__device__ __forceinline__ void cp_async_8(double *smem_dst, const double *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_16(double *smem_dst, const double *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPEND>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory"); }

// ------------------------------------------------------------------ tile staging
template <int RB, int TS = OCS_TILE>
struct TileLayout {
    static constexpr int S = TS;
    static constexpr int GP = (2 * S + 1) * RB;   // grid-point records
    static constexpr int ST = 4 * S;              // step records
};

__device__ __forceinline__ void copy_to_smem(double *dst, const double *src, int ndoubles)
{
    // both sides are 16-byte aligned and ndoubles is even
    const double2 *s2 = reinterpret_cast<const double2 *>(src);
    double2 *d2 = reinterpret_cast<double2 *>(dst);
    const int cnt = ndoubles >> 1;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) d2[k] = s2[k];
}

// Copies the grid-point records [2*i0, 2*(i0+ns)], step records [i0, i0+ns) and (BW > 0) basis
// records of one tile into shared memory.  All threads of the block take part.
template <int RB>
__device__ __forceinline__ void stage_tile(const SweepArgs &a, int i0, int ns, int BW, double *gp, double *st, double *bs,
                                           double *fo = nullptr, int FW = 0)
{
    __syncthreads();   // the previous tile has been consumed by every warp
    copy_to_smem(gp, a.gprec + (size_t)2 * i0 * RB, (2 * ns + 1) * RB);
    copy_to_smem(st, a.steprec + (size_t)4 * i0, 4 * ns);
    if (BW > 0) copy_to_smem(bs, a.basrec + (size_t)2 * i0 * BW, (2 * ns + 1) * BW);
    if (FW > 0) copy_to_smem(fo, a.foldrec + (size_t)FW * i0, FW * ns);
    __syncthreads();
}

// The same copies issued with cp.async (16 bytes each; both sides are 16-byte aligned, ndoubles is even): no barrier,
// no commit -- the caller commits the group and waits for it before the block barrier that publishes the tile.
__device__ __forceinline__ void copy_to_smem_async(double *dst, const double *src, int ndoubles)
{
    const int cnt = ndoubles >> 1;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) cp_async_16(dst + 2 * k, src + 2 * k);
}
template <int RB>
__device__ __forceinline__ void stage_tile_async(const SweepArgs &a, int i0, int ns, int BW, double *gp, double *st, double *bs,
                                                 double *fo, int FW)
{
    copy_to_smem_async(gp, a.gprec + (size_t)2 * i0 * RB, (2 * ns + 1) * RB);
    copy_to_smem_async(st, a.steprec + (size_t)4 * i0, 4 * ns);
    if (BW > 0) copy_to_smem_async(bs, a.basrec + (size_t)2 * i0 * BW, (2 * ns + 1) * BW);
    if (FW > 0) copy_to_smem_async(fo, a.foldrec + (size_t)FW * i0, FW * ns);
}

// ------------------------------------------------------------------ forward sweep
// RK4Integrator.m:28-56 for one trajectory per thread.
template <class P, int UMODE, int BMODE, bool PB, int KD, bool GRIDS>
__device__ __forceinline__ void forward_body(const SweepArgs &a, double *sm)
{
    constexpr int n = P::n, m = n + 1, nc = P::nc, np = P::np;
    constexpr int HM = ocs_hoist_mode(UMODE, PB);
    constexpr int RB = ocs_rec_width(nc, P::ntt(HM));
    constexpr bool NEEDB = (UMODE == UMODE_COEF);
    constexpr int BWc = NEEDB ? (BMODE == BMODE_SPARSE ? 4 : (KD > 0 ? ocs_even(KD) : -1)) : 0;
    typedef TileLayout<RB, OCS_TILE_SINGLE> TL;
    const int BW = BWc >= 0 ? BWc : a.BW;
    const int N = a.N;
    double *gp = sm, *st = sm + TL::GP, *bs = st + TL::ST;
    if (a.RB != RB || (NEEDB && a.BW != BW)) {
        if (threadIdx.x == 0 && blockIdx.x == 0 && a.error_flag) *a.error_flag = 1;
        return;
    }

    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    bool active = b < a.nb;
    const long bb = active ? b : a.nb - 1;   // idle lanes shadow the last trajectory and never store
    if (a.frozen && a.frozen[bb]) active = false;   // fb_sweep: converged trajectories are left as they are

    double p[np > 0 ? np : 1];
#pragma unroll
    for (int k = 0; k < np; ++k) p[k] = PB ? a.p[soa(k, a.ldp, bb)] : (np <= OCS_NPC ? a.pc[k] : a.p[k]);

    double y[m];
#pragma unroll
    for (int c = 0; c < n; ++c) y[c] = a.x0[soa(c, a.ldb, bb)];
    y[n] = 0.0;                                                   // xK(:,1,1) = [x0; 0]   :33

    ControlEval<nc, UMODE, BMODE, KD> ctl;
    ctl.init(a, bb);
    double uL[nc], uM[nc], uR[nc];

    double *xsp = a.xs + b;                        // running pointer into the node-state store
    constexpr bool gridout = GRIDS;   // fetch variant: also writes the full grid, J row included

    for (int i0 = 0; i0 < N; i0 += TL::S) {
        const int ns = (N - i0 < TL::S) ? (N - i0) : TL::S;
        stage_tile<RB>(a, i0, ns, NEEDB ? BW : 0, gp, st, bs);
        if (i0 == 0) {
            ctl.eval(a, gp, bs, 0, bb, uL);
            if constexpr (UMODE == UMODE_COEF && BMODE == BMODE_DENSE) {
                if (a.ug_out && active) {
#pragma unroll
                    for (int c = 0; c < nc; ++c) a.ug_out[soa((long)c * (2 * N + 1), a.ldb, b)] = uL[c];
                }
            }
        }
#pragma unroll 4
        for (int s = 0; s < ns; ++s) {
            const int i = i0 + s;
            const double *rL = gp + (2 * s) * RB, *rM = rL + RB, *rR = rM + RB;
            const double *bM = bs + (2 * s + 1) * BW, *bR = bM + BW;
#if defined(OCS_UNIFORM_H) && OCS_UNIFORM_H
            const double h = a.hc[0], h2 = a.hc[1], h6 = a.hc[3];
#else
            const double h = st[4 * s], h2 = st[4 * s + 1], h6 = st[4 * s + 3];
#endif
            if constexpr (UMODE == UMODE_UNIFORM) ctl.eval(a, rL, bs, 2 * i, bb, uL);   // a shared-memory read
            if (active) {
#pragma unroll
                for (int c = 0; c < n; ++c) xsp[(size_t)c * (N + 1) * a.ldb] = y[c];
                if (gridout) {
#pragma unroll
                    for (int c = 0; c < m; ++c) a.xgrid[c + (size_t)m * (i + (size_t)(N + 1) * b)] = y[c];
                }
            }
            xsp += a.ldb;
            ctl.eval(a, rM, bM, 2 * i + 1, bb, uM);
            ctl.eval(a, rR, bR, 2 * i + 2, bb, uR);
            if constexpr (UMODE == UMODE_COEF && BMODE == BMODE_DENSE) {
                if (a.ug_out && active) {   // the reverse sweep reads the grid instead of re-evaluating v*B
#pragma unroll
                    for (int c = 0; c < nc; ++c) {
                        a.ug_out[soa((long)c * (2 * N + 1) + 2 * i + 1, a.ldb, b)] = uM[c];
                        a.ug_out[soa((long)c * (2 * N + 1) + 2 * i + 2, a.ldb, b)] = uR[c];
                    }
                }
            }

            double F[m], acc[m], ys[n];
            P::template F<HM>(rL + 1 + nc, rL[0], y, uL, p, F);                  // F1   :39
#pragma unroll
            for (int c = 0; c < m; ++c) acc[c] = F[c];
#pragma unroll
            for (int c = 0; c < n; ++c) ys[c] = y[c] + h2 * F[c];                 // xK2  :40
            P::template F<HM>(rM + 1 + nc, rM[0], ys, uM, p, F);                 // F2   :42
#pragma unroll
            for (int c = 0; c < m; ++c) acc[c] = acc[c] + 2.0 * F[c];
#pragma unroll
            for (int c = 0; c < n; ++c) ys[c] = y[c] + h2 * F[c];                 // xK3  :43
            P::template F<HM>(rM + 1 + nc, rM[0], ys, uM, p, F);                 // F3   :45
#pragma unroll
            for (int c = 0; c < m; ++c) acc[c] = acc[c] + 2.0 * F[c];
#pragma unroll
            for (int c = 0; c < n; ++c) ys[c] = y[c] + h * F[c];                  // xK4  :46
            P::template F<HM>(rR + 1 + nc, rR[0], ys, uR, p, F);                 // F4   :48
#pragma unroll
            for (int c = 0; c < m; ++c) y[c] = y[c] + h6 * (acc[c] + F[c]);       // :50-51
            if constexpr (UMODE != UMODE_UNIFORM) {   // per-trajectory controls are carried, not re-evaluated
#pragma unroll
                for (int c = 0; c < nc; ++c) uL[c] = uR[c];
            }
        }
    }

    if (active) {
#pragma unroll
        for (int c = 0; c < n; ++c) xsp[(size_t)c * (N + 1) * a.ldb] = y[c];
        if (gridout) {
#pragma unroll
            for (int c = 0; c < m; ++c) a.xgrid[c + (size_t)m * (N + (size_t)(N + 1) * b)] = y[c];
        }
        if (a.xT) {
#pragma unroll
            for (int c = 0; c < n; ++c) a.xT[soa(c, a.ldb, b)] = y[c];
        }
        double Jv = y[n];                                           // J = x(end,end)   :55
        if (a.J) {
            if (a.J_accumulate) Jv = a.J[b] + Jv;                   // J = J1 + J2  (RK4InfiniteIntegrator.m:23)
            a.J[b] = Jv;
        }
        if (a.status) {
            bool ok = isfinite(Jv);
#pragma unroll
            for (int c = 0; c < n; ++c) ok = ok && isfinite(y[c]);
            a.status[b] = (a.J_accumulate ? a.status[b] : 0) | (ok ? 0 : 1);
        }
    }
}

// ------------------------------------------------------------------ forward sweep, T trajectories per thread
// RK4Integrator.m:28-56 with T independent trajectories per thread (see reverse_multi_body for the
// mapping): record reads, address arithmetic and loop control are issued once for T trajectories and
// the T dependency chains interleave.  Same arithmetic as forward_body; no grid output (fetch uses
// forward_body).
template <class P, int UMODE, int BMODE, bool PB, int KD, int T>
__device__ __forceinline__ void forward_multi_body(const SweepArgs &a, double *sm)
{
    constexpr int n = P::n, m = n + 1, nc = P::nc, np = P::np;
    constexpr int HM = ocs_hoist_mode(UMODE, PB);
    constexpr int RB = ocs_rec_width(nc, P::ntt(HM));
    constexpr bool NEEDB = (UMODE == UMODE_COEF);
    constexpr int BWc = NEEDB ? (BMODE == BMODE_SPARSE ? 4 : (KD > 0 ? ocs_even(KD) : -1)) : 0;
    typedef TileLayout<RB> TL;
    constexpr int NF = (UMODE == UMODE_COEF && !OCS_FOLD_FWD_COEF) ? 0 : ocs_nfold<P>(HM), FW = ocs_fold_width(NF);
    const int BW = BWc >= 0 ? BWc : a.BW;
    const int N = a.N;
    // one tile = [grid-point records | step records | basis records | fold records]; two of them when the next tile is
    // fetched asynchronously while this one is consumed
    const int TD = TL::GP + TL::ST + (2 * TL::S + 1) * (NEEDB ? BW : 0) + TL::S * FW;
    double *gp = sm, *st = sm + TL::GP, *bs = st + TL::ST;
    double *fo = bs + (size_t)(2 * TL::S + 1) * (NEEDB ? BW : 0);       // fold records of the tile
    if (a.RB != RB || (NEEDB && a.BW != BW) || (NF > 0 && !a.foldrec)) {
        if (threadIdx.x == 0 && blockIdx.x == 0 && a.error_flag) *a.error_flag = 1;
        return;
    }

    long b[T], bb[T];
    bool active[T];
    double p[T][np > 0 ? np : 1];
    double y[T][m];
    ControlEval<nc, UMODE, BMODE, KD> ctl[T];
    double uL[T][nc], uM[T][nc], uR[T][nc];
    double uMn[T][nc], uRn[T][nc];      // UMODE_GRID: control of the next step, loaded one step ahead
    double *xsp[T];
    // time exponentials by recurrence (see ocs_ntexp): the right-hand sides are then called in HM_TIME form on a
    // per-trajectory table
    constexpr int NE = ocs_ntexp<P>(PB);
    constexpr int HMF = NE > 0 ? (int)HM_TIME : HM;
    constexpr int NT1 = NE > 0 ? P::ntt(HM_TIME) : 1;
    double Ecur[T][NE > 0 ? NE : 1], Eq[T][NE > 0 ? NE : 1], Erate[T][NE > 0 ? NE : 1];
#pragma unroll
    for (int q = 0; q < T; ++q) {
        b[q] = ((long)blockIdx.x * T + q) * blockDim.x + threadIdx.x;
        active[q] = b[q] < a.nb;
        bb[q] = active[q] ? b[q] : a.nb - 1;
        if (a.frozen && a.frozen[bb[q]]) active[q] = false;
#pragma unroll
        for (int k = 0; k < np; ++k) p[q][k] = PB ? a.p[soa(k, a.ldp, bb[q])] : (np <= OCS_NPC ? a.pc[k] : a.p[k]);
#pragma unroll
        for (int c = 0; c < n; ++c) y[q][c] = a.x0[soa(c, a.ldb, bb[q])];
        y[q][n] = 0.0;                                                // xK(:,1,1) = [x0; 0]   :33
        ctl[q].init(a, bb[q]);
        xsp[q] = a.xs + b[q];
        if constexpr (NE > 0) ocs_dep<P, ocs_bool<(NE > 0)>>::type::texp_rates(p[q], Erate[q]);
    }

    if constexpr (OCS_ASYNC_STAGE) {
        stage_tile_async<RB>(a, 0, N < TL::S ? N : TL::S, NEEDB ? BW : 0, gp, st, bs, fo, FW);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
    }
    for (int i0 = 0; i0 < N; i0 += TL::S) {
        const int ns = (N - i0 < TL::S) ? (N - i0) : TL::S;
        if constexpr (OCS_ASYNC_STAGE) {
            // this tile is complete and visible (barrier at the end of the previous iteration); fetch the next one
            // into the other buffer while this one is consumed
            if (i0 + TL::S < N) {
                const int nn = (N - i0 - TL::S < TL::S) ? (N - i0 - TL::S) : TL::S;
                double *ng = gp == sm ? sm + TD : sm;
                stage_tile_async<RB>(a, i0 + TL::S, nn, NEEDB ? BW : 0, ng, ng + TL::GP, ng + TL::GP + TL::ST,
                                     ng + TL::GP + TL::ST + (size_t)(2 * TL::S + 1) * (NEEDB ? BW : 0), FW);
                cp_async_commit();
            }
        } else {
            stage_tile<RB>(a, i0, ns, NEEDB ? BW : 0, gp, st, bs, fo, FW);
        }
        if constexpr (NE > 0) {     // anchor: exp() at the tile's first node, and the half-step factor of its first step
#pragma unroll
            for (int q = 0; q < T; ++q) {
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    Ecur[q][e] = exp(Erate[q][e] * gp[0]);
                    Eq[q][e] = exp(Erate[q][e] * st[1]);
                }
            }
        }
        if (i0 == 0) {
#pragma unroll
            for (int q = 0; q < T; ++q) {
                ctl[q].eval(a, gp, bs, 0, bb[q], uL[q]);
                if constexpr (UMODE == UMODE_GRID) {
                    ctl[q].eval(a, gp, bs, 1, bb[q], uMn[q]);
                    ctl[q].eval(a, gp, bs, 2, bb[q], uRn[q]);
                }
                if constexpr (UMODE == UMODE_COEF && BMODE == BMODE_DENSE) {
                    if (a.ug_out && active[q]) {
#pragma unroll
                        for (int c = 0; c < nc; ++c) a.ug_out[soa((long)c * (2 * N + 1), a.ldb, b[q])] = uL[q][c];
                    }
                }
            }
        }
#pragma unroll 4
        for (int s = 0; s < ns; ++s) {
            const int i = i0 + s;
            const double *rL = gp + (2 * s) * RB, *rM = rL + RB, *rR = rM + RB;
            const double *bM = bs + (2 * s + 1) * BW, *bR = bM + BW;
#if defined(OCS_UNIFORM_H) && OCS_UNIFORM_H
            const double h = a.hc[0], h2 = a.hc[1], h6 = a.hc[3];
#else
            const double h = st[4 * s], h2 = st[4 * s + 1], h6 = st[4 * s + 3];
#endif
            constexpr bool SPLIT = (UMODE == UMODE_COEF && BMODE == BMODE_SPARSE);
            if constexpr (SPLIT) {
                // controls of all T trajectories first: the (uniform, rare) window moves happen here, outside the
                // arithmetic below, whose T dependency chains then share one basic block and interleave
                if constexpr (OCS_SLIDE_PRED_FWD == 3) {
                    // one host-computed flag per step (bit4 of the midpoint's basis record): does the window move here?
                    if (rec_flags(bM) & 16) {
#pragma unroll
                        for (int q = 0; q < T; ++q) ctl[q].template slide<0>(a, bM, bb[q]);
#pragma unroll
                        for (int q = 0; q < T; ++q) ctl[q].value(bM, uM[q]);
#pragma unroll
                        for (int q = 0; q < T; ++q) ctl[q].template slide<0>(a, bR, bb[q]);
#pragma unroll
                        for (int q = 0; q < T; ++q) ctl[q].value(bR, uR[q]);
                    } else {
#pragma unroll
                        for (int q = 0; q < T; ++q) ctl[q].value(bM, uM[q]);
#pragma unroll
                        for (int q = 0; q < T; ++q) ctl[q].value(bR, uR[q]);
                    }
                } else {
#pragma unroll
                for (int q = 0; q < T; ++q) ctl[q].template slide<OCS_SLIDE_PRED_FWD>(a, bM, bb[q]);
#pragma unroll
                for (int q = 0; q < T; ++q) ctl[q].value(bM, uM[q]);
#pragma unroll
                for (int q = 0; q < T; ++q) ctl[q].template slide<OCS_SLIDE_PRED_FWD>(a, bR, bb[q]);
#pragma unroll
                for (int q = 0; q < T; ++q) ctl[q].value(bR, uR[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < T; ++q) {
                if (active[q]) {
#pragma unroll
                    for (int c = 0; c < n; ++c) xsp[q][(size_t)c * (N + 1) * a.ldb] = y[q][c];
                }
                xsp[q] += a.ldb;
                if constexpr (UMODE == UMODE_UNIFORM) ctl[q].eval(a, rL, bs, 2 * i, bb[q], uL[q]);
                if constexpr (UMODE == UMODE_GRID) {
                    // per-trajectory grid control: use the values loaded during the previous step and issue the
                    // loads of the next step now (clamped at the last grid point), hiding the HBM latency
#pragma unroll
                    for (int c = 0; c < nc; ++c) { uM[q][c] = uMn[q][c]; uR[q][c] = uRn[q][c]; }
                    const int jm = 2 * i + 3 < 2 * N ? 2 * i + 3 : 2 * N, jr = 2 * i + 4 < 2 * N ? 2 * i + 4 : 2 * N;
                    ctl[q].eval(a, rM, bM, jm, bb[q], uMn[q]);
                    ctl[q].eval(a, rR, bR, jr, bb[q], uRn[q]);
                } else if constexpr (!SPLIT) {
                    ctl[q].eval(a, rM, bM, 2 * i + 1, bb[q], uM[q]);
                    ctl[q].eval(a, rR, bR, 2 * i + 2, bb[q], uR[q]);
                }
                if constexpr (UMODE == UMODE_COEF && BMODE == BMODE_DENSE) {
                    if (a.ug_out && active[q]) {   // the reverse sweep reads the grid instead of re-evaluating v*B
#pragma unroll
                        for (int c = 0; c < nc; ++c) {
                            *soa_row(a.ug_out + b[q], (unsigned)(c * (2 * N + 1) + 2 * i + 1), (unsigned)a.ldb * 8u) = uM[q][c];
                            *soa_row(a.ug_out + b[q], (unsigned)(c * (2 * N + 1) + 2 * i + 2), (unsigned)a.ldb * 8u) = uR[q][c];
                        }
                    }
                }

                double F[m], acc[m], ys[n];
                if constexpr (NF > 0) {
                    // weight-folded objective row: J += (w E)(..) per stage, state rows as below
                    const double *fL = fo + s * FW, *fM = fL + NF, *fR = fM + NF;
                    P::template Ff<HM>(rL + 1 + nc, fL, rL[0], y[q], uL[q], p[q], F, y[q][n]);   // F1   :39
#pragma unroll
                    for (int c = 0; c < n; ++c) acc[c] = F[c];
#pragma unroll
                    for (int c = 0; c < n; ++c) ys[c] = y[q][c] + h2 * F[c];                     // xK2  :40
                    P::template Ff<HM>(rM + 1 + nc, fM, rM[0], ys, uM[q], p[q], F, y[q][n]);     // F2   :42
#pragma unroll
                    for (int c = 0; c < n; ++c) acc[c] = acc[c] + 2.0 * F[c];
#pragma unroll
                    for (int c = 0; c < n; ++c) ys[c] = y[q][c] + h2 * F[c];                     // xK3  :43
                    P::template Ff<HM>(rM + 1 + nc, fM, rM[0], ys, uM[q], p[q], F, y[q][n]);     // F3   :45
#pragma unroll
                    for (int c = 0; c < n; ++c) acc[c] = acc[c] + 2.0 * F[c];
#pragma unroll
                    for (int c = 0; c < n; ++c) ys[c] = y[q][c] + h * F[c];                      // xK4  :46
                    P::template Ff<HM>(rR + 1 + nc, fR, rR[0], ys, uR[q], p[q], F, y[q][n]);     // F4   :48
#pragma unroll
                    for (int c = 0; c < n; ++c) y[q][c] = y[q][c] + h6 * (acc[c] + F[c]);        // :50-51
                } else {
                // tables of the three grid points: the staged record, or (per-trajectory parameters, time
                // exponentials by recurrence) this trajectory's own, one multiplication per grid point
                double tqL[NT1], tqM[NT1], tqR[NT1];
                if constexpr (NE > 0) {
                    typedef typename ocs_dep<P, ocs_bool<(NE > 0)>>::type PE;
                    double EM[NE], ER[NE];
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        if (!a.texp_uniform) Eq[q][e] = exp(Erate[q][e] * h2);
                        EM[e] = Ecur[q][e] * Eq[q][e];
                        ER[e] = EM[e] * Eq[q][e];
                    }
                    PE::table_from_texp(Ecur[q], rL[0], p[q], tqL);
                    PE::table_from_texp(EM, rM[0], p[q], tqM);
                    PE::table_from_texp(ER, rR[0], p[q], tqR);
#pragma unroll
                    for (int e = 0; e < NE; ++e) Ecur[q][e] = ER[e];
                }
                const double *ttL = NE > 0 ? tqL : rL + 1 + nc, *ttM = NE > 0 ? tqM : rM + 1 + nc, *ttR = NE > 0 ? tqR : rR + 1 + nc;
                P::template F<HMF>(ttL, rL[0], y[q], uL[q], p[q], F);                // F1   :39
#pragma unroll
                for (int c = 0; c < m; ++c) acc[c] = F[c];
#pragma unroll
                for (int c = 0; c < n; ++c) ys[c] = y[q][c] + h2 * F[c];             // xK2  :40
                P::template F<HMF>(ttM, rM[0], ys, uM[q], p[q], F);                  // F2   :42
#pragma unroll
                for (int c = 0; c < m; ++c) acc[c] = acc[c] + 2.0 * F[c];
#pragma unroll
                for (int c = 0; c < n; ++c) ys[c] = y[q][c] + h2 * F[c];             // xK3  :43
                P::template F<HMF>(ttM, rM[0], ys, uM[q], p[q], F);                  // F3   :45
#pragma unroll
                for (int c = 0; c < m; ++c) acc[c] = acc[c] + 2.0 * F[c];
#pragma unroll
                for (int c = 0; c < n; ++c) ys[c] = y[q][c] + h * F[c];              // xK4  :46
                P::template F<HMF>(ttR, rR[0], ys, uR[q], p[q], F);                  // F4   :48
#pragma unroll
                for (int c = 0; c < m; ++c) y[q][c] = y[q][c] + h6 * (acc[c] + F[c]);   // :50-51
                }
                if constexpr (UMODE != UMODE_UNIFORM) {
#pragma unroll
                    for (int c = 0; c < nc; ++c) uL[q][c] = uR[q][c];
                }
            }
        }
        if constexpr (OCS_ASYNC_STAGE) {
            if (i0 + TL::S < N) {
                cp_async_wait<0>();     // this thread's copies of the next tile have landed ...
                __syncthreads();        // ... and so have everybody else's; nobody reads the old tile any more
                gp = gp == sm ? sm + TD : sm;
                st = gp + TL::GP;
                bs = st + TL::ST;
                fo = bs + (size_t)(2 * TL::S + 1) * (NEEDB ? BW : 0);
            }
        }
    }

#pragma unroll
    for (int q = 0; q < T; ++q) {
        if (!active[q]) continue;
#pragma unroll
        for (int c = 0; c < n; ++c) xsp[q][(size_t)c * (N + 1) * a.ldb] = y[q][c];
        if (a.xT) {
#pragma unroll
            for (int c = 0; c < n; ++c) a.xT[soa(c, a.ldb, b[q])] = y[q][c];
        }
        double Jv = y[q][n];                                          // J = x(end,end)   :55
        if (a.J) {
            if (a.J_accumulate) Jv = a.J[b[q]] + Jv;                  // J = J1 + J2  (RK4InfiniteIntegrator.m:23)
            a.J[b[q]] = Jv;
        }
        if (a.status) {
            bool ok = isfinite(Jv);
#pragma unroll
            for (int c = 0; c < n; ++c) ok = ok && isfinite(y[q][c]);
            a.status[b[q]] = (a.J_accumulate ? a.status[b[q]] : 0) | (ok ? 0 : 1);
        }
    }
}

// ------------------------------------------------------------------ reverse sweep
// RK4Integrator.m:59-121 for one trajectory per thread.  Stage states 2..4 are
// recomputed from the stored node state with the same device functions as the forward
// sweep (16 B/step of HBM traffic instead of 64).  FB != 0 = fb_sweep variant: no dJdu; the
// control on the grid is replaced by ControlChar(t, x, lam) and the normalised change
// is reduced per trajectory (functions/fb_sweep.m:96,107-110).  FB == 1: lam is the discrete
// adjoint of the RK4 scheme (the v2 Integrator's).  FB == 2: lam solves the continuous costate
// equation lam' = adjointRHS(t, x, lam, u) = -gradient(H, x) (functions/make_from_symbolic.m:11-14,
// functions/compute_x_lam.m:11-12) backwards from lam(T) = 0 by classical RK4 on the same grid, with
// the midpoint state taken as the mean of the neighbouring nodes (the textbook forward-backward sweep);
// with H = f + lam*g, -gradient(H, x) is -dFdx_times_vec(t, y, u, [lam; 1]).
template <class P, int UMODE, int BMODE, bool PB, int KD, int FB, bool GRIDS>
__device__ __forceinline__ void reverse_body(const SweepArgs &a, double *sm)
{
    constexpr int n = P::n, m = n + 1, nc = P::nc, np = P::np;
    constexpr int HM = ocs_hoist_mode(UMODE, PB);
    constexpr int NT = P::ntt(HM);
    constexpr int RB = ocs_rec_width(nc, NT);
    constexpr bool NEEDB = (BMODE != BMODE_NONE);
    constexpr bool GACC = NEEDB && !GRIDS && FB == 0;   // fold dJdu into dJdv (the fetch variant writes dJdu itself)
    constexpr int BWc = NEEDB ? (BMODE == BMODE_SPARSE ? 4 : (KD > 0 ? ocs_even(KD) : -1)) : 0;
    typedef TileLayout<RB, OCS_TILE_SINGLE> TL;
    const int BW = BWc >= 0 ? BWc : a.BW;
    const int N = a.N, G = 2 * N + 1;
    double *gp = sm, *st = sm + TL::GP, *bs = st + TL::ST;
    if (a.RB != RB || (NEEDB && a.BW != BW)) {
        if (threadIdx.x == 0 && blockIdx.x == 0 && a.error_flag) *a.error_flag = 1;
        return;
    }

    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    bool active = b < a.nb;
    const long bb = active ? b : a.nb - 1;
    if (a.frozen && a.frozen[bb]) active = false;       // fb_sweep: converged trajectories keep their control

    double p[np > 0 ? np : 1];
#pragma unroll
    for (int k = 0; k < np; ++k) p[k] = PB ? a.p[soa(k, a.ldp, bb)] : (np <= OCS_NPC ? a.pc[k] : a.p[k]);

    double lam[m];
#pragma unroll
    for (int c = 0; c < m; ++c)
        lam[c] = a.lamT ? (a.ldl ? a.lamT[soa(c, a.ldl, bb)] : a.lamT[c]) : (c == n ? 1.0 : 0.0);   // :63-66
    const bool lamout = GRIDS && a.lamgrid != nullptr, duout = GRIDS && a.dJdugrid != nullptr;
    const bool lamsout = GRIDS && a.lams != nullptr;
    const bool want_du = GACC || duout;
    if (lamout && active) {
#pragma unroll
        for (int c = 0; c < m; ++c) a.lamgrid[c + (size_t)m * (N + (size_t)(N + 1) * b)] = lam[c];
    }
    if (lamsout && active) {
#pragma unroll
        for (int c = 0; c < n; ++c) a.lams[soa((long)c * (N + 1) + N, a.ldlams, b)] = lam[c];
    }

    ControlEval<nc, UMODE, BMODE, KD> ctl;
    ctl.init(a, bb);
    GradAcc<nc, GACC ? BMODE : BMODE_NONE, KD> gacc;
    gacc.init();

    double uL[nc], uM[nc], uR[nc], dcar[nc];
    constexpr int DU = OCS_UPREFETCH;
    double uMn[DU][nc], uLn[DU][nc];   // UMODE_GRID: controls of the next DU steps, loaded ahead (ring, [0] = next)
    double y1[n], yn1[n], yn2[n];      // node states: current, and the next two (prefetched)
    double xR[n];                      // FB: node state to the right (for midpoint averages)
    MaxRatio chg;                      // FB: running max of |uNew-u| / (uRelTol |u| + uAbsTol)
    bool unfinite = true;              // FB: every uNew so far was finite (a NaN compares false everywhere and would
                                       // otherwise leave the maximum at 0, i.e. "converged")
    const bool fbwrite = (FB != 0) && !a.fb_readonly;
    double *ugw = (FB != 0) ? a.ug + bb : nullptr;     // FB: this trajectory's column of the control grid (bb == b when active)
    double bnd[2 * nc];                // FB: ControlBounds, read once (ControlChar clamps with them twice per step)
#pragma unroll
    for (int c = 0; c < 2 * nc; ++c) bnd[c] = (FB != 0) ? a.bounds[c] : 0.0;
    const unsigned ld8 = (unsigned)a.ldb * 8u;
#pragma unroll
    for (int c = 0; c < nc; ++c) dcar[c] = 0.0;
    const double *xsp = a.xs + (size_t)(N - 1) * a.ldb + bb;   // node N-1
    // fb_sweep variant: cp.async ring [DP][n + 2 nc][threads] of (x_i, u_{2i+1}, u_{2i}) per step
    constexpr int DP = (FB != 0 && UMODE == UMODE_GRID) ? OCS_FB_RING : 0;
    static_assert(DP == 0 || ((DP & (DP - 1)) == 0 && DP <= OCS_XS_SLACK), "OCS_FB_RING must be a power of two <= OCS_XS_SLACK");
    constexpr int RW = n + 2 * nc;
    double *ring = bs + (size_t)(2 * TL::S + 1) * BW + threadIdx.x;
    auto ring_fill = [&](double *slot, int node) {     // node < 0: x from the slack (never used), u clamped
        const int jm = 2 * node + 1 > 0 ? 2 * node + 1 : 0, jl = node > 0 ? 2 * node : 0;
        // one IMAD.WIDE per address (row index relative to the start of the slack in front of the store)
        const double *xbase = a.xs + bb - (size_t)OCS_XS_SLACK * a.ldb;
#pragma unroll
        for (int c = 0; c < n; ++c)
            cp_async_8(slot + (size_t)c * OCS_THREADS, soa_row(xbase, (unsigned)(c * (N + 1) + node + OCS_XS_SLACK), ld8));
#pragma unroll
        for (int c = 0; c < nc; ++c) {
            cp_async_8(slot + (size_t)(n + c) * OCS_THREADS, soa_row(a.ug + bb, (unsigned)(c * G + jm), ld8));
            cp_async_8(slot + (size_t)(n + nc + c) * OCS_THREADS, soa_row(a.ug + bb, (unsigned)(c * G + jl), ld8));
        }
        cp_async_commit();
    };
    if constexpr (DP > 0) {
#pragma unroll
        for (int d = 0; d < DP; ++d) ring_fill(ring + (size_t)(d * RW) * OCS_THREADS, N - 1 - d);
    }
#pragma unroll
    for (int c = 0; c < n; ++c) {
        yn1[c] = xsp[(size_t)c * (N + 1) * a.ldb];
        yn2[c] = xsp[(size_t)c * (N + 1) * a.ldb - a.ldb];   // node N-2 (slack when N == 1)
    }
    if constexpr (FB != 0) {
#pragma unroll
        for (int c = 0; c < n; ++c) xR[c] = xsp[(size_t)c * (N + 1) * a.ldb + a.ldb];
    }
    xsp -= a.ldb;                                              // node N-2: the deepest prefetch so far

    const int i0_last = ((N - 1) / TL::S) * TL::S;
    for (int i0 = i0_last; i0 >= 0; i0 -= TL::S) {
        const int ns = (N - i0 < TL::S) ? (N - i0) : TL::S;
        stage_tile<RB>(a, i0, ns, NEEDB ? BW : 0, gp, st, bs);
        if (i0 == i0_last) {
            ctl.eval(a, gp + (2 * ns) * RB, bs + (2 * ns) * BW, 2 * N, bb, uR);
            if constexpr (UMODE == UMODE_GRID && DP == 0) {
#pragma unroll
                for (int d = 0; d < DU; ++d) {
                    const int jm = 2 * (N - 1 - d) + 1 > 0 ? 2 * (N - 1 - d) + 1 : 0;
                    const int jl = 2 * (N - 1 - d) > 0 ? 2 * (N - 1 - d) : 0;
                    ctl.eval(a, gp, bs, jm, bb, uMn[d]);
                    ctl.eval(a, gp, bs, jl, bb, uLn[d]);
                }
            }
        }
#pragma unroll 2
        for (int s = ns - 1; s >= 0; --s) {
            const int i = i0 + s;
            const double *rL = gp + (2 * s) * RB, *rM = rL + RB, *rR = rM + RB;
            const double *bL = bs + (2 * s) * BW, *bM = bL + BW, *bR = bM + BW;
#if defined(OCS_UNIFORM_H) && OCS_UNIFORM_H
            const double h = a.hc[0], h2 = a.hc[1], h3 = a.hc[2], h6 = a.hc[3];
#else
            const double h = st[4 * s], h2 = st[4 * s + 1], h3 = st[4 * s + 2], h6 = st[4 * s + 3];
#endif
            if constexpr (DP > 0) {
                // step i sits in slot (N-1-i) mod DP; all but the DP-1 youngest groups have landed.  fb_sweep only
                // overwrites grid points of steps >= i, the ring only reads steps <= i - 1: no hazard.
                double *slot = ring + (size_t)(((N - 1 - i) & (DP - 1)) * RW) * OCS_THREADS;
                cp_async_wait<DP - 1>();
#pragma unroll
                for (int c = 0; c < n; ++c) y1[c] = slot[(size_t)c * OCS_THREADS];
#pragma unroll
                for (int c = 0; c < nc; ++c) {
                    uM[c] = slot[(size_t)(n + c) * OCS_THREADS];
                    uL[c] = slot[(size_t)(n + nc + c) * OCS_THREADS];
                }
                ring_fill(slot, i - DP);
            } else {
#pragma unroll
            for (int c = 0; c < n; ++c) { y1[c] = yn1[c]; yn1[c] = yn2[c]; }
            xsp -= a.ldb;
            // prefetch two nodes ahead (node i-2): the HBM latency overlaps two steps of arithmetic.
            // Unconditional: for i < 2 it reads the slack in front of the store and is never used.
#pragma unroll
            for (int c = 0; c < n; ++c) yn2[c] = xsp[(size_t)c * (N + 1) * a.ldb];
            }
            if constexpr (UMODE == UMODE_UNIFORM) ctl.eval(a, rR, bR, 2 * i + 2, bb, uR);   // shared-memory reads
            if constexpr (DP > 0) {
                // controls came out of the ring
            } else if constexpr (UMODE == UMODE_GRID) {
                // per-trajectory grid control: loaded one step ahead (fb_sweep only overwrites grid points of the
                // current step after they were read, so the early loads never see an updated value)
#pragma unroll
                for (int c = 0; c < nc; ++c) { uM[c] = uMn[0][c]; uL[c] = uLn[0][c]; }
#pragma unroll
                for (int d = 0; d + 1 < DU; ++d) {
#pragma unroll
                    for (int c = 0; c < nc; ++c) { uMn[d][c] = uMn[d + 1][c]; uLn[d][c] = uLn[d + 1][c]; }
                }
                const int jm = 2 * (i - DU) + 1 > 0 ? 2 * (i - DU) + 1 : 0, jl = 2 * (i - DU) > 0 ? 2 * (i - DU) : 0;
                ctl.eval(a, rM, bM, jm, bb, uMn[DU - 1]);
                ctl.eval(a, rL, bL, jl, bb, uLn[DU - 1]);
            } else {
                ctl.eval(a, rM, bM, 2 * i + 1, bb, uM);
                ctl.eval(a, rL, bL, 2 * i, bb, uL);
            }
            const double *ttR = rR + 1 + nc, *ttM = rM + 1 + nc, *ttL = rL + 1 + nc;
            const double tR = rR[0], tM = rM[0], tL = rL[0];

            double lamR[n];
          if constexpr (FB == 2) {
            // k_s = adjointRHS = -g_s;  lam - (h/2) k = lam + (h/2) g  (sign flips are exact)
            double xm[n], vs[m], g1[n], g2[n], g3[n], g4[n];
            vs[n] = lam[n];
#pragma unroll
            for (int c = 0; c < n; ++c) { vs[c] = lam[c]; xm[c] = (y1[c] + xR[c]) / 2; }
            P::template X<HM>(ttR, tR, xR, uR, p, vs, g1);
#pragma unroll
            for (int c = 0; c < n; ++c) vs[c] = lam[c] + h2 * g1[c];
            P::template X<HM>(ttM, tM, xm, uM, p, vs, g2);
#pragma unroll
            for (int c = 0; c < n; ++c) vs[c] = lam[c] + h2 * g2[c];
            P::template X<HM>(ttM, tM, xm, uM, p, vs, g3);
#pragma unroll
            for (int c = 0; c < n; ++c) vs[c] = lam[c] + h * g3[c];
            P::template X<HM>(ttL, tL, y1, uL, p, vs, g4);
#pragma unroll
            for (int c = 0; c < n; ++c) {
                lamR[c] = lam[c];
                lam[c] = lam[c] + h6 * (g1[c] + 2.0 * g2[c] + 2.0 * g3[c] + g4[c]);
            }
            (void)h3;
          } else {
            // stage states of the forward step (RK4Integrator.m:39-46), recomputed
            double F[m], y2[n], y3[n], y4[n];
            P::template F<HM>(ttL, tL, y1, uL, p, F);
#pragma unroll
            for (int c = 0; c < n; ++c) y2[c] = y1[c] + h2 * F[c];
            P::template F<HM>(ttM, tM, y2, uM, p, F);
#pragma unroll
            for (int c = 0; c < n; ++c) y3[c] = y1[c] + h2 * F[c];
            P::template F<HM>(ttM, tM, y3, uM, p, F);
#pragma unroll
            for (int c = 0; c < n; ++c) y4[c] = y1[c] + h * F[c];

            double k4[m], k3[m], k2[m], k1[m], g3[n], g2[n], g1[n], g0[n];
#pragma unroll
            for (int c = 0; c < m; ++c) k4[c] = h6 * lam[c];                              // :73
            P::template X<HM>(ttR, tR, y4, uR, p, k4, g3);                               // :74-75
#pragma unroll
            for (int c = 0; c < n; ++c) k3[c] = h3 * lam[c] + h * g3[c];                  // :77
            k3[n] = h3 * lam[n];
            P::template X<HM>(ttM, tM, y3, uM, p, k3, g2);                               // :78-79
#pragma unroll
            for (int c = 0; c < n; ++c) k2[c] = h3 * lam[c] + h2 * g2[c];                 // :81
            k2[n] = h3 * lam[n];
            P::template X<HM>(ttM, tM, y2, uM, p, k2, g1);                               // :82-83
#pragma unroll
            for (int c = 0; c < n; ++c) k1[c] = h6 * lam[c] + h2 * g1[c];                 // :85
            k1[n] = h6 * lam[n];
            P::template X<HM>(ttL, tL, y1, uL, p, k1, g0);                               // :86-88

            if (want_du) {
                // dJdu at the right node: stage 1 of step i+1 (carried) + stage 4 of step i   :112-120
                double d[nc], e[nc];
                P::template U<HM>(ttR, tR, y4, uR, p, k4, d);
#pragma unroll
                for (int c = 0; c < nc; ++c) d[c] = dcar[c] + d[c];   // dcar == 0 at the right end
                // the host flags the steps in which the knot window moves; all others take the straight path
                const bool moves = GACC && BMODE == BMODE_SPARSE ? (rec_flags(bM) & 4) != 0 : true;
                if (moves) gacc.add(a, bR, d, b, active); else gacc.add_fast(bR, d);
                if (duout && active) {
#pragma unroll
                    for (int c = 0; c < nc; ++c) a.dJdugrid[c + (size_t)nc * ((2 * i + 2) + (size_t)G * b)] = d[c];
                }
                // midpoint: stages 2 and 3   :105-109
                P::template U<HM>(ttM, tM, y2, uM, p, k2, d);
                P::template U<HM>(ttM, tM, y3, uM, p, k3, e);
#pragma unroll
                for (int c = 0; c < nc; ++c) d[c] = d[c] + e[c];
                if (moves) gacc.add(a, bM, d, b, active); else gacc.add_fast(bM, d);
                if (duout && active) {
#pragma unroll
                    for (int c = 0; c < nc; ++c) a.dJdugrid[c + (size_t)nc * ((2 * i + 1) + (size_t)G * b)] = d[c];
                }
                // left node: stage 1 of this step, completed by the next iteration (or the left end :101-102)
                P::template U<HM>(ttL, tL, y1, uL, p, k1, dcar);
            }

#pragma unroll
            for (int c = 0; c < n; ++c) {
                lamR[c] = lam[c];
                lam[c] = lam[c] + g1[c] + g2[c] + g3[c] + g0[c];                           // :86
            }
          }
            if (lamout && active) {
#pragma unroll
                for (int c = 0; c < m; ++c) a.lamgrid[c + (size_t)m * (i + (size_t)(N + 1) * b)] = lam[c];
            }
            if (lamsout && active) {
#pragma unroll
                for (int c = 0; c < n; ++c) a.lams[soa((long)c * (N + 1) + i, a.ldlams, b)] = lam[c];
            }

            if constexpr (FB != 0) {
                // uNew at the right node and at the midpoint of step i (midpoint x, lam = mean of the
                // neighbouring nodes, the same formula the grid uses for t, RK4Integrator.m:23)
                double un[nc], xm[n], lm[n];
                P::template ControlChar<HM>(ttR, tR, xR, lamR, p, bnd, un);
#pragma unroll
                for (int c = 0; c < nc; ++c) {
                    chg.update(fabs(un[c] - uR[c]), a.uRelTol * fabs(uR[c]) + a.uAbsTol);          // fb_sweep.m:107-108
                    unfinite = unfinite && isfinite(un[c]);
                    const double uw = (a.relax == 1.0) ? un[c] : (1.0 - a.relax) * uR[c] + a.relax * un[c];
                    if (active && fbwrite) *soa_row(ugw, (unsigned)(c * G + 2 * i + 2), ld8) = uw;
                }
#pragma unroll
                for (int c = 0; c < n; ++c) { xm[c] = (y1[c] + xR[c]) / 2; lm[c] = (lam[c] + lamR[c]) / 2; }
                P::template ControlChar<HM>(ttM, tM, xm, lm, p, bnd, un);
#pragma unroll
                for (int c = 0; c < nc; ++c) {
                    chg.update(fabs(un[c] - uM[c]), a.uRelTol * fabs(uM[c]) + a.uAbsTol);
                    unfinite = unfinite && isfinite(un[c]);
                    const double uw = (a.relax == 1.0) ? un[c] : (1.0 - a.relax) * uM[c] + a.relax * un[c];
                    if (active && fbwrite) *soa_row(ugw, (unsigned)(c * G + 2 * i + 1), ld8) = uw;
                }
#pragma unroll
                for (int c = 0; c < n; ++c) xR[c] = y1[c];
            }
            if constexpr (UMODE != UMODE_UNIFORM) {   // per-trajectory controls are carried, not re-evaluated
#pragma unroll
                for (int c = 0; c < nc; ++c) uR[c] = uL[c];
            }
        }
    }

    if constexpr (FB == 0) {
        if (want_du) {
            gacc.add(a, bs, dcar, b, active);                                              // left end :101-102
            if (duout && active) {
#pragma unroll
                for (int c = 0; c < nc; ++c) a.dJdugrid[c + (size_t)nc * ((size_t)G * b)] = dcar[c];
            }
            gacc.finish(a, b, active);
        }
    } else {
        double un[nc];
        double lam0[n];
#pragma unroll
        for (int c = 0; c < n; ++c) lam0[c] = lam[c];
        P::template ControlChar<HM>(gp + 1 + nc, gp[0], xR, lam0, p, bnd, un);
#pragma unroll
        for (int c = 0; c < nc; ++c) {
            chg.update(fabs(un[c] - uR[c]), a.uRelTol * fabs(uR[c]) + a.uAbsTol);
            unfinite = unfinite && isfinite(un[c]);
            const double uw = (a.relax == 1.0) ? un[c] : (1.0 - a.relax) * uR[c] + a.relax * un[c];
            if (active && fbwrite) *soa_row(ugw, (unsigned)(c * G), ld8) = uw;
        }
        // a non-finite uNew never counts as converged (the reference's max() would propagate the NaN and fail `<= 1`)
        if (active && a.max_change) a.max_change[b] = unfinite ? chg.value() : __longlong_as_double(0x7ff0000000000000LL);
    }
    if (a.lam0 && active) {
#pragma unroll
        for (int c = 0; c < m; ++c) a.lam0[soa(c, a.ldb, b)] = lam[c];
    }
}

// ------------------------------------------------------------------ reverse sweep, T trajectories per thread
// The plain gradient sweep (no grids, no fb_sweep) with T independent trajectories per thread:
// thread tid of a block handles trajectories blockIdx*blockDim*T + q*blockDim + tid, q = 0..T-1, so
// every q-slice is coalesced.  Everything read from shared memory (records, weights, step sizes) and
// all address / loop / branch instructions are issued once for T trajectories, and the T dependency
// chains interleave in the FP64 pipe.  Same arithmetic, operation for operation, as reverse_body.
template <class P, int UMODE, int BMODE, bool PB, int KD, int T>
__device__ __forceinline__ void reverse_multi_body(const SweepArgs &a, double *sm)
{
    constexpr int n = P::n, m = n + 1, nc = P::nc, np = P::np;
    constexpr int HM = ocs_hoist_mode(UMODE, PB);
    constexpr int NT = P::ntt(HM);
    constexpr int RB = ocs_rec_width(nc, NT);
    constexpr bool NEEDB = (BMODE != BMODE_NONE);
    constexpr int BWc = NEEDB ? (BMODE == BMODE_SPARSE ? 4 : (KD > 0 ? ocs_even(KD) : -1)) : 0;
    typedef TileLayout<RB> TL;
    constexpr int NF = ocs_nfold<P>(HM), FW = ocs_fold_width(NF);
    const int BW = BWc >= 0 ? BWc : a.BW;
    const int N = a.N, G = 2 * N + 1;
    const bool dugout = a.dug != nullptr;
    // one tile = [grid-point records | step records | basis records | fold records]; two of them when the next tile is
    // fetched asynchronously while this one is consumed (forward_multi_body has the details); the node-state ring follows
    const int TD = TL::GP + TL::ST + (2 * TL::S + 1) * BW + TL::S * FW;
    double *gp = sm, *st = sm + TL::GP, *bs = st + TL::ST;
    if (a.RB != RB || (NEEDB && a.BW != BW) || (NF > 0 && !a.foldrec)) {
        if (threadIdx.x == 0 && blockIdx.x == 0 && a.error_flag) *a.error_flag = 1;
        return;
    }

    long b[T], bb[T];
    bool active[T];
    double p[T][np > 0 ? np : 1];
    double lam[T][m];
    ControlEval<nc, UMODE, BMODE, KD> ctl[T];
    GradAcc<nc, BMODE, KD> gacc[T];
    double uL[T][nc], uM[T][nc], uR[T][nc], dcar[T][nc], dR[T][nc], dM[T][nc];
    double uMn[T][nc], uLn[T][nc];      // UMODE_GRID: control of the next step, loaded one step ahead
    // time exponentials by recurrence (see ocs_ntexp), walking the grid downwards
    constexpr int NE = ocs_ntexp<P>(PB);
    constexpr int HMF = NE > 0 ? (int)HM_TIME : HM;
    constexpr int NT1 = NE > 0 ? P::ntt(HM_TIME) : 1;
    double Ecur[T][NE > 0 ? NE : 1], Eq[T][NE > 0 ? NE : 1], Erate[T][NE > 0 ? NE : 1];
    constexpr int D = OCS_PREFETCH;          // ring of node states in flight: yn[0] is the next one consumed
    constexpr int UNR = OCS_REV_UNROLL;
    constexpr int DP = PB ? OCS_RING_PB : OCS_RING;   // > 0: cp.async ring in shared memory [DP][T][n (+ 2 nc)][threads] instead of yn
    static_assert(DP == 0 || ((DP & (DP - 1)) == 0 && DP <= OCS_XS_SLACK), "OCS_RING / OCS_RING_PB must be a power of two <= OCS_XS_SLACK");
    // with a per-trajectory control grid the ring also carries the step's two controls (midpoint, left node): every
    // per-step global load of the sweep then completes through cp.async groups, DP steps ahead, instead of the
    // scoreboard -- with 8 resident warps per SM a quarter of the generated 6-state reverse sweep was ONE instruction
    // waiting for a scoreboard shared with the loads issued just before it (profiles/r2_notes.md)
    constexpr int RW = n + (UMODE == UMODE_GRID ? 2 * nc : 0);
    double y1[T][n], yn[D][T][n];            // (yn is unused, and optimised away, with the shared-memory ring)
    const double *xsp[T];
    double *ring = sm + (size_t)(OCS_ASYNC_STAGE ? 2 : 1) * TD + threadIdx.x;
    double *fo = bs + (size_t)(2 * TL::S + 1) * BW;    // fold records of the tile
    const bool folded = NF > 0 && a.lamJ_one != 0;     // uniform: the whole grid takes one of the two chains
    // ring slot of one node (step): states x_node, then u at grid points 2 node + 1 and 2 node (clamped at 0)
    auto ring_issue = [&](double *slot, int q, int node) {
        // one IMAD.WIDE per address: row index relative to the start of the slack in front of the store (node >= -SLACK)
        const double *xbase = a.xs + bb[q] - (size_t)OCS_XS_SLACK * a.ldb;
#pragma unroll
        for (int c = 0; c < n; ++c)
            cp_async_8(slot + (size_t)(q * RW + c) * OCS_THREADS,
                       soa_row(xbase, (unsigned)(c * (N + 1) + node + OCS_XS_SLACK), (unsigned)a.ldb * 8u));
        if constexpr (UMODE == UMODE_GRID) {
            const int jm = 2 * node + 1 > 0 ? 2 * node + 1 : 0, jl = node > 0 ? 2 * node : 0;
#pragma unroll
            for (int c = 0; c < nc; ++c) {
                cp_async_8(slot + (size_t)(q * RW + n + c) * OCS_THREADS, soa_row(a.ug + bb[q], (unsigned)(c * G + jm), (unsigned)a.ldb * 8u));
                cp_async_8(slot + (size_t)(q * RW + n + nc + c) * OCS_THREADS, soa_row(a.ug + bb[q], (unsigned)(c * G + jl), (unsigned)a.ldb * 8u));
            }
        }
    };
#pragma unroll
    for (int q = 0; q < T; ++q) {
        b[q] = ((long)blockIdx.x * T + q) * blockDim.x + threadIdx.x;
        active[q] = b[q] < a.nb;
        bb[q] = active[q] ? b[q] : a.nb - 1;
        if (a.frozen && a.frozen[bb[q]]) active[q] = false;
#pragma unroll
        for (int k = 0; k < np; ++k) p[q][k] = PB ? a.p[soa(k, a.ldp, bb[q])] : (np <= OCS_NPC ? a.pc[k] : a.p[k]);
#pragma unroll
        for (int c = 0; c < m; ++c)
            lam[q][c] = a.lamT ? (a.ldl ? a.lamT[soa(c, a.ldl, bb[q])] : a.lamT[c]) : (c == n ? 1.0 : 0.0);
        ctl[q].init(a, bb[q]);
        gacc[q].init();
        if constexpr (NE > 0) P::texp_rates(p[q], Erate[q]);
#pragma unroll
        for (int c = 0; c < nc; ++c) dcar[q][c] = 0.0;
        xsp[q] = a.xs + (size_t)(N - 1) * a.ldb + bb[q];     // node N-1; nodes below 0 fall into the slack
        if constexpr (DP == 0) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
#pragma unroll
                for (int c = 0; c < n; ++c) yn[d][q][c] = xsp[q][(size_t)c * (N + 1) * a.ldb - (size_t)d * a.ldb];
            }
            xsp[q] -= (size_t)(D - 1) * a.ldb;
        }
    }
    const int i0_last = ((N - 1) / TL::S) * TL::S;
    if constexpr (OCS_ASYNC_STAGE) {
        stage_tile_async<RB>(a, i0_last, N - i0_last, NEEDB ? BW : 0, gp, st, bs, fo, folded ? FW : 0);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
    }
    if constexpr (DP > 0) {
        // fill the ring: nodes N-1 ... N-DP, one commit group per node (slot = step count mod DP)
#pragma unroll
        for (int d = 0; d < DP; ++d) {
#pragma unroll
            for (int q = 0; q < T; ++q) ring_issue(ring + (size_t)(d * T * RW) * OCS_THREADS, q, N - 1 - d);
            cp_async_commit();
        }
    }

    for (int i0 = i0_last; i0 >= 0; i0 -= TL::S) {
        const int ns = (N - i0 < TL::S) ? (N - i0) : TL::S;
        if constexpr (OCS_ASYNC_STAGE) {
            // this tile is complete and visible; fetch the one below it (always a full tile) into the other buffer
            if (i0 > 0) {
                double *ng = gp == sm ? sm + TD : sm;
                stage_tile_async<RB>(a, i0 - TL::S, TL::S, NEEDB ? BW : 0, ng, ng + TL::GP, ng + TL::GP + TL::ST,
                                     ng + TL::GP + TL::ST + (size_t)(2 * TL::S + 1) * BW, folded ? FW : 0);
                cp_async_commit();
            }
        } else {
            stage_tile<RB>(a, i0, ns, NEEDB ? BW : 0, gp, st, bs, fo, folded ? FW : 0);
        }
        if (i0 == i0_last) {
#pragma unroll
            for (int q = 0; q < T; ++q) {
                ctl[q].eval(a, gp + (2 * ns) * RB, bs + (2 * ns) * BW, 2 * N, bb[q], uR[q]);
                if constexpr (UMODE == UMODE_GRID && DP == 0) {
                    ctl[q].eval(a, gp, bs, 2 * N - 1, bb[q], uMn[q]);
                    ctl[q].eval(a, gp, bs, 2 * N - 2, bb[q], uLn[q]);
                }
            }
        }
        if constexpr (NE > 0) {     // anchor: exp() at the tile's last node, and the (inverse) half-step factor of its last step
#pragma unroll
            for (int q = 0; q < T; ++q) {
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    Ecur[q][e] = exp(Erate[q][e] * gp[(2 * ns) * RB]);
                    Eq[q][e] = exp(-(Erate[q][e] * st[4 * (ns - 1) + 1]));
                }
            }
        }
        // one RK4 step (all T trajectories) as a lambda, so that the step loop below can be unrolled by hand:
        // `#pragma unroll 2` on this loop miscompiled the peeled remainder iteration for odd step counts
        // (nvcc 12.9, caught by tests/test_gpu_edges.py); pairs + one explicit odd step are equivalent and safe.
        auto rk4_reverse_step = [&](auto fold_tag, const int s) {
            constexpr bool FO = decltype(fold_tag)::value;
            typedef typename ocs_dep<P, decltype(fold_tag)>::type PF;
            const int i = i0 + s;
            const double *rL = gp + (2 * s) * RB, *rM = rL + RB, *rR = rM + RB;
            const double *bL = bs + (2 * s) * BW, *bM = bL + BW, *bR = bM + BW;
#if defined(OCS_UNIFORM_H) && OCS_UNIFORM_H
            const double h = a.hc[0], h2 = a.hc[1], h3 = a.hc[2], h6 = a.hc[3];
#else
            const double h = st[4 * s], h2 = st[4 * s + 1], h3 = st[4 * s + 2], h6 = st[4 * s + 3];
#endif
            const double *ttR0 = rR + 1 + nc, *ttM0 = rM + 1 + nc, *ttL0 = rL + 1 + nc;
            const double tR = rR[0], tM = rM[0], tL = rL[0];
            const bool moves = BMODE == BMODE_SPARSE ? (rec_flags(bM) & 4) != 0 : true;
            if constexpr (DP > 0) {
                // node i sits in slot (N-1-i) mod DP; every group but the DP-1 youngest has landed (a tile fetch
                // committed in between only makes this wait for one ring group more than it needs)
                double *slot = ring + (size_t)(((N - 1 - i) & (DP - 1)) * T * RW) * OCS_THREADS;
                cp_async_wait<DP - 1>();
#pragma unroll
                for (int q = 0; q < T; ++q) {
#pragma unroll
                    for (int c = 0; c < n; ++c) y1[q][c] = slot[(size_t)(q * RW + c) * OCS_THREADS];
                    if constexpr (UMODE == UMODE_GRID) {
#pragma unroll
                        for (int c = 0; c < nc; ++c) {
                            uM[q][c] = slot[(size_t)(q * RW + n + c) * OCS_THREADS];
                            uL[q][c] = slot[(size_t)(q * RW + n + nc + c) * OCS_THREADS];
                        }
                    }
                }
                // refill the slot with node i-DP (below node 0: the slack in front of the store, never used)
#pragma unroll
                for (int q = 0; q < T; ++q) ring_issue(slot, q, i - DP);
                cp_async_commit();
            }
            if constexpr (DP == 0) {
#pragma unroll
                for (int q = 0; q < T; ++q) {
                    // rotate the ring (register renaming under the unroll) and issue the load D nodes ahead;
                    // unconditional: below node 0 it reads the slack in front of the store and is never used
#pragma unroll
                    for (int c = 0; c < n; ++c) y1[q][c] = yn[0][q][c];
#pragma unroll
                    for (int d = 0; d + 1 < D; ++d) {
#pragma unroll
                        for (int c = 0; c < n; ++c) yn[d][q][c] = yn[d + 1][q][c];
                    }
                    xsp[q] -= a.ldb;
#pragma unroll
                    for (int c = 0; c < n; ++c) yn[D - 1][q][c] = xsp[q][(size_t)c * (N + 1) * a.ldb];
                }
            }
            constexpr bool SPLIT = (UMODE == UMODE_COEF && BMODE == BMODE_SPARSE);
            if constexpr (SPLIT) {      // controls of all T trajectories first (see forward_multi_body)
                if constexpr (OCS_SLIDE_PRED_REV == 3) {
                    if (rec_flags(bM) & 8) {        // host-computed: the control window moves within this step
#pragma unroll
                        for (int q = 0; q < T; ++q) ctl[q].template slide<0>(a, bM, bb[q]);
#pragma unroll
                        for (int q = 0; q < T; ++q) ctl[q].value(bM, uM[q]);
#pragma unroll
                        for (int q = 0; q < T; ++q) ctl[q].template slide<0>(a, bL, bb[q]);
#pragma unroll
                        for (int q = 0; q < T; ++q) ctl[q].value(bL, uL[q]);
                    } else {
#pragma unroll
                        for (int q = 0; q < T; ++q) ctl[q].value(bM, uM[q]);
#pragma unroll
                        for (int q = 0; q < T; ++q) ctl[q].value(bL, uL[q]);
                    }
                } else {
#pragma unroll
                for (int q = 0; q < T; ++q) ctl[q].template slide<OCS_SLIDE_PRED_REV>(a, bM, bb[q]);
#pragma unroll
                for (int q = 0; q < T; ++q) ctl[q].value(bM, uM[q]);
#pragma unroll
                for (int q = 0; q < T; ++q) ctl[q].template slide<OCS_SLIDE_PRED_REV>(a, bL, bb[q]);
#pragma unroll
                for (int q = 0; q < T; ++q) ctl[q].value(bL, uL[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < T; ++q) {
                if constexpr (UMODE == UMODE_UNIFORM) ctl[q].eval(a, rR, bR, 2 * i + 2, bb[q], uR[q]);
                if constexpr (UMODE == UMODE_GRID && DP > 0) {
                    // controls came out of the ring
                } else if constexpr (UMODE == UMODE_GRID) {     // loaded one step ahead, see reverse_body
#pragma unroll
                    for (int c = 0; c < nc; ++c) { uM[q][c] = uMn[q][c]; uL[q][c] = uLn[q][c]; }
                    const int jm = 2 * i - 1 > 0 ? 2 * i - 1 : 0, jl = 2 * i - 2 > 0 ? 2 * i - 2 : 0;
                    ctl[q].eval(a, rM, bM, jm, bb[q], uMn[q]);
                    ctl[q].eval(a, rL, bL, jl, bb[q], uLn[q]);
                } else if constexpr (!SPLIT) {
                    ctl[q].eval(a, rM, bM, 2 * i + 1, bb[q], uM[q]);
                    ctl[q].eval(a, rL, bL, 2 * i, bb[q], uL[q]);
                }

                // tables of the step's three grid points: the staged records, or this trajectory's own (time
                // exponentials by recurrence: one multiplication per grid point, downwards)
                double tqL[NT1], tqM[NT1], tqR[NT1];
                if constexpr (NE > 0) {
                    typedef typename ocs_dep<P, decltype(fold_tag)>::type PE;
                    double EM[NE], EL[NE];
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        if (!a.texp_uniform) Eq[q][e] = exp(-(Erate[q][e] * h2));
                        EM[e] = Ecur[q][e] * Eq[q][e];
                        EL[e] = EM[e] * Eq[q][e];
                    }
                    PE::table_from_texp(Ecur[q], tR, p[q], tqR);
                    PE::table_from_texp(EM, tM, p[q], tqM);
                    PE::table_from_texp(EL, tL, p[q], tqL);
#pragma unroll
                    for (int e = 0; e < NE; ++e) Ecur[q][e] = EL[e];
                }
                const double *ttR = NE > 0 ? tqR : ttR0, *ttM = NE > 0 ? tqM : ttM0, *ttL = NE > 0 ? tqL : ttL0;
                double F[m], y2[n], y3[n], y4[n];
                P::template F<HMF>(ttL, tL, y1[q], uL[q], p[q], F);
#pragma unroll
                for (int c = 0; c < n; ++c) y2[c] = y1[q][c] + h2 * F[c];
                P::template F<HMF>(ttM, tM, y2, uM[q], p[q], F);
#pragma unroll
                for (int c = 0; c < n; ++c) y3[c] = y1[q][c] + h2 * F[c];
                P::template F<HMF>(ttM, tM, y3, uM[q], p[q], F);
#pragma unroll
                for (int c = 0; c < n; ++c) y4[c] = y1[q][c] + h * F[c];

                // Adjoint chain :73-88 with each dFdu_times_vec (:101-120) evaluated as soon as its stage costate
                // dJdk is known, so that stage states and dJdk's die early (the sums are formed in the reference's
                // order; only the evaluation order of independent terms differs).
                constexpr bool GA = (BMODE != BMODE_NONE);
                const bool du = GA || dugout;
                double k4[m], k3[m], k2[m], k1[m], g3[n], g2[n], g1[n], g0[n];
                double dRq[nc], dM3[nc], dM2[nc];
                // Larger problems (FMA-contracted build only) add each g to a running sum as soon as it is known --
                // lam + g3 + g2 + g1 + g0 instead of the reference's lam + g1 + g2 + g3 + g0 (:86): the same terms in
                // another order (rounding only), and 2 n fewer live doubles across the step
#if defined(OCS_STRICT) && OCS_STRICT
                constexpr bool RUNSUM = false;
#else
                constexpr bool RUNSUM = n >= 3;
#endif
                double ls[n];
                const double *fL = fo + s * FW, *fM = fL + NF, *fR = fM + NF;      // (FO only)
                if constexpr (FO) {
                    // lam(end) == 1: the objective row of every dJdk is the stage weight, already folded into fL/fM/fR
#pragma unroll
                    for (int c = 0; c < n; ++c) k4[c] = h6 * lam[q][c];                          // :73
                    if (du) PF::template Uf<HM>(ttR, fR, tR, y4, uR[q], p[q], k4, dRq);          // :119-120 / :116
                    PF::template Xf<HM>(ttR, fR, tR, y4, uR[q], p[q], k4, g3);                   // :74-75
#pragma unroll
                    for (int c = 0; c < n; ++c) k3[c] = h3 * lam[q][c] + h * g3[c];              // :77
                    if constexpr (RUNSUM) {
#pragma unroll
                        for (int c = 0; c < n; ++c) ls[c] = lam[q][c] + g3[c];
                    }
                    if (du) PF::template Uf<HM>(ttM, fM, tM, y3, uM[q], p[q], k3, dM3);          // :108-109
                    PF::template Xf<HM>(ttM, fM, tM, y3, uM[q], p[q], k3, g2);                   // :78-79
#pragma unroll
                    for (int c = 0; c < n; ++c) k2[c] = h3 * lam[q][c] + h2 * g2[c];             // :81
                    if constexpr (RUNSUM) {
#pragma unroll
                        for (int c = 0; c < n; ++c) ls[c] = ls[c] + g2[c];
                    }
                    if (du) PF::template Uf<HM>(ttM, fM, tM, y2, uM[q], p[q], k2, dM2);          // :106-107
                    PF::template Xf<HM>(ttM, fM, tM, y2, uM[q], p[q], k2, g1);                   // :82-83
#pragma unroll
                    for (int c = 0; c < n; ++c) k1[c] = h6 * lam[q][c] + h2 * g1[c];             // :85
                    if constexpr (RUNSUM) {
#pragma unroll
                        for (int c = 0; c < n; ++c) ls[c] = ls[c] + g1[c];
                    }
                    PF::template Xf<HM>(ttL, fL, tL, y1[q], uL[q], p[q], k1, g0);                // :86-88
                } else {
#pragma unroll
                for (int c = 0; c < m; ++c) k4[c] = h6 * lam[q][c];                          // :73
                if (du) P::template U<HMF>(ttR, tR, y4, uR[q], p[q], k4, dRq);               // :119-120 / :116
                P::template X<HMF>(ttR, tR, y4, uR[q], p[q], k4, g3);                        // :74-75
#pragma unroll
                for (int c = 0; c < n; ++c) k3[c] = h3 * lam[q][c] + h * g3[c];              // :77
                k3[n] = h3 * lam[q][n];
                if constexpr (RUNSUM) {
#pragma unroll
                    for (int c = 0; c < n; ++c) ls[c] = lam[q][c] + g3[c];
                }
                if (du) P::template U<HMF>(ttM, tM, y3, uM[q], p[q], k3, dM3);               // :108-109
                P::template X<HMF>(ttM, tM, y3, uM[q], p[q], k3, g2);                        // :78-79
#pragma unroll
                for (int c = 0; c < n; ++c) k2[c] = h3 * lam[q][c] + h2 * g2[c];             // :81
                k2[n] = h3 * lam[q][n];
                if constexpr (RUNSUM) {
#pragma unroll
                    for (int c = 0; c < n; ++c) ls[c] = ls[c] + g2[c];
                }
                if (du) P::template U<HMF>(ttM, tM, y2, uM[q], p[q], k2, dM2);               // :106-107
                P::template X<HMF>(ttM, tM, y2, uM[q], p[q], k2, g1);                        // :82-83
#pragma unroll
                for (int c = 0; c < n; ++c) k1[c] = h6 * lam[q][c] + h2 * g1[c];             // :85
                k1[n] = h6 * lam[q][n];
                if constexpr (RUNSUM) {
#pragma unroll
                    for (int c = 0; c < n; ++c) ls[c] = ls[c] + g1[c];
                }
                P::template X<HMF>(ttL, tL, y1[q], uL[q], p[q], k1, g0);                     // :86-88
                }

                if constexpr (GA) {
#pragma unroll
                    for (int c = 0; c < nc; ++c) {
                        dR[q][c] = dcar[q][c] + dRq[c];        // stage 1 of step i+1 (carried) + stage 4 of step i
                        dM[q][c] = dM2[c] + dM3[c];            // stages 2 + 3
                    }
                    if constexpr (FO) PF::template Uf<HM>(ttL, fL, tL, y1[q], uL[q], p[q], k1, dcar[q]);
                    else P::template U<HMF>(ttL, tL, y1[q], uL[q], p[q], k1, dcar[q]);      // :101-102 / :112-116
                } else {
                    if (dugout) {            // dJdu itself, SoA [nc][G][batch], coalesced (dense bases project it later)
#pragma unroll
                        for (int c = 0; c < nc; ++c) {
                            if (active[q]) {
                                *soa_row(a.dug + b[q], (unsigned)(c * G + 2 * i + 2), (unsigned)a.ldb * 8u) = dcar[q][c] + dRq[c];
                                *soa_row(a.dug + b[q], (unsigned)(c * G + 2 * i + 1), (unsigned)a.ldb * 8u) = dM2[c] + dM3[c];
                            }
                        }
                        if constexpr (FO) PF::template Uf<HM>(ttL, fL, tL, y1[q], uL[q], p[q], k1, dcar[q]);
                        else P::template U<HMF>(ttL, tL, y1[q], uL[q], p[q], k1, dcar[q]);
                    }
                }
#pragma unroll
                for (int c = 0; c < n; ++c)
                    lam[q][c] = RUNSUM ? ls[c] + g0[c] : lam[q][c] + g1[c] + g2[c] + g3[c] + g0[c];   // :86
                if constexpr (UMODE != UMODE_UNIFORM) {
#pragma unroll
                    for (int c = 0; c < nc; ++c) uR[q][c] = uL[q][c];
                }
            }
            // fold dJdu of the right node and the midpoint into dJdv: ONE uniform branch for all T trajectories,
            // outside the arithmetic above so that the T dependency chains share a basic block and interleave
            if constexpr (BMODE != BMODE_NONE) {
                if (moves) {
#pragma unroll
                    for (int q = 0; q < T; ++q) {
                        gacc[q].add(a, bR, dR[q], b[q], active[q]);
                        gacc[q].add(a, bM, dM[q], b[q], active[q]);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < T; ++q) {
                        gacc[q].add_fast(bR, dR[q]);
                        gacc[q].add_fast(bM, dM[q]);
                    }
                }
            }
        };
        auto run_tile = [&](auto fold_tag) {
            int s = ns - 1;
            if constexpr (UNR >= 2) {
#pragma unroll 1
                for (; s >= 1; s -= 2) {
                    rk4_reverse_step(fold_tag, s);
                    rk4_reverse_step(fold_tag, s - 1);
                }
            }
#pragma unroll 1
            for (; s >= 0; --s) rk4_reverse_step(fold_tag, s);
        };
        if constexpr (NF > 0) {
            if (folded) run_tile(ocs_bool<true>{});
            else run_tile(ocs_bool<false>{});
        } else {
            run_tile(ocs_bool<false>{});
        }
        if constexpr (OCS_ASYNC_STAGE) {
            if (i0 > 0) {
                // the tile fetch is older than the last ns ring groups: waiting for all but DP-1 groups covers it
                // whenever the tile had at least DP steps (only the first, partial tile can be shorter)
                if (DP > 0 && ns >= DP) cp_async_wait<(DP > 0 ? DP - 1 : 0)>();
                else cp_async_wait<0>();
                __syncthreads();        // everybody's copies have landed; nobody reads the old tile any more
                gp = gp == sm ? sm + TD : sm;
                st = gp + TL::GP;
                bs = st + TL::ST;
                fo = bs + (size_t)(2 * TL::S + 1) * BW;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < T; ++q) {
        if constexpr (BMODE != BMODE_NONE) {
            gacc[q].add(a, bs, dcar[q], b[q], active[q]);                                    // left end :101-102
            gacc[q].finish(a, b[q], active[q]);
        } else {
            if (dugout && active[q]) {
#pragma unroll
                for (int c = 0; c < nc; ++c) a.dug[soa((long)c * G, a.ldb, b[q])] = dcar[q][c];
            }
        }
        if (a.lam0 && active[q]) {
#pragma unroll
            for (int c = 0; c < m; ++c) a.lam0[soa(c, a.ldb, b[q])] = lam[q][c];
        }
    }
}

// ------------------------------------------------------------------ small helper kernels

// Fills the table slots of the grid-point records: every sub-expression of the right-hand
// sides that is batch-uniform -- a function of t and shared parameters (HM_TIME) and, when
// the control is shared too, of u (HM_TIME_CONTROL) -- is evaluated once per grid point here,
// with the same device code the in-thread variant would run.
template <class P>
__device__ __forceinline__ void table_body(double *gprec, int G, int hm, const double *p)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= G) return;
    constexpr int nc = P::nc;
    double pr[P::np > 0 ? P::np : 1];
#pragma unroll
    for (int k = 0; k < P::np; ++k) pr[k] = p[k];
    if (hm == HM_TIME) {
        constexpr int NT = P::ntt(HM_TIME), RB = ocs_rec_width(nc, NT);
        double *r = gprec + (size_t)j * RB;
        double tt[NT > 0 ? NT : 1];
        P::template table<HM_TIME>(r[0], r + 1, pr, tt);
#pragma unroll
        for (int k = 0; k < NT; ++k) r[1 + nc + k] = tt[k];
    } else if (hm == HM_TIME_CONTROL) {
        constexpr int NT = P::ntt(HM_TIME_CONTROL), RB = ocs_rec_width(nc, NT);
        double *r = gprec + (size_t)j * RB;
        double tt[NT > 0 ? NT : 1], u[nc];
#pragma unroll
        for (int c = 0; c < nc; ++c) u[c] = r[1 + c];
        P::template table<HM_TIME_CONTROL>(r[0], u, pr, tt);
#pragma unroll
        for (int k = 0; k < NT; ++k) r[1 + nc + k] = tt[k];
    }
}

// Fills the fold records of one segment (after the table kernel): for every step the problem's fold entries at
// the left node (weight h/6), the midpoint (h/3) and the right node (h/6).
template <class P, int HM>
__device__ __forceinline__ void fold_rows(const double *gprec, const double *steprec, double *foldrec, int N)
{
    constexpr int NF = ocs_nfold<P>(HM);
    if constexpr (NF > 0) {
        constexpr int RB = ocs_rec_width(P::nc, P::ntt(HM)), FW = ocs_fold_width(NF);
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= N) return;
        const double h3 = steprec[4 * i + 2], h6 = steprec[4 * i + 3];
        const double *ttL = gprec + (size_t)(2 * i) * RB + 1 + P::nc;
        double *f = foldrec + (size_t)i * FW;
        P::template fold<HM>(ttL, h6, f);
        P::template fold<HM>(ttL + RB, h3, f + NF);
        P::template fold<HM>(ttL + 2 * RB, h6, f + 2 * NF);
        for (int k = 3 * NF; k < FW; ++k) f[k] = 0.0;
    }
}

// ------------------------------------------------------------------ batched compute_equilibrium
// functions/compute_equilibrium.m:1-33 for one parameter set per thread.  The reference hands the
// residual system (:10-21)
//     value(x)   = F(0, x, u)(1:n)
//     value(lam) = r*lam - dFdx_times_vec(0, x, u, [lam; 1])(1:n)
//     value(u)   = dFdu_times_vec(0, x, u, [lam; 1])
// to lsqnonlin with bounds (:23-27).  Here every thread runs a Levenberg-Marquardt iteration on its
// own system (central-difference Jacobian, Cholesky solve of the damped normal equations, active-set
// treatment of [lb, ub] plus projection of the step); the root, not the iterate path, is what compares
// with lsqnonlin.  exitflag 1 = converged to a (possibly local) minimiser of sum(residual.^2): resnorm
// tells whether it is a root.
// All host-layout arrays are MATLAB column-major with one column per parameter set.
struct EqArgs {
    long nsets;
    int np;
    const double *params;   // [np x nparams]
    long nparams;           // 1 or nsets
    const double *guess;    // [(2n+nc) x nguess]
    long nguess;            // 1 or nsets
    const double *lb, *ub;  // [(2n+nc)] or null
    const double *r;        // [nr]
    long nr;                // 1 or nsets
    int max_iter;
    double tol;             // stop when max |step_k| <= tol*(1+|y_k|) or sum(residual.^2) <= tol^2
    double *y;              // [(2n+nc) x nsets]  = [xStar; lamStar; uStar]
    double *resnorm;        // [nsets] sum(residual.^2)
    double *residual;       // [(2n+nc) x nsets] or null
    int *exitflag;          // [nsets] 1 converged, 0 iteration limit, -1 not finite
    int *iters;             // [nsets] or null
};

template <class P>
__device__ __forceinline__ void eq_residual(const double *y, const double *p, double r, double *f)
{
    constexpr int n = P::n, m = n + 1, nc = P::nc;
    double ya[m], v[m], Fo[m], gx[n], gu[nc], u[nc];
    const double dummy[2] = {0.0, 0.0};
#pragma unroll
    for (int c = 0; c < n; ++c) { ya[c] = y[c]; v[c] = y[n + c]; }
    ya[n] = 0.0;
    v[n] = 1.0;                                                        // [lam; 1]   :16,:19
#pragma unroll
    for (int c = 0; c < nc; ++c) u[c] = y[2 * n + c];
    P::template F<HM_NONE>(dummy, 0.0, ya, u, p, Fo);                  // :13
    P::template X<HM_NONE>(dummy, 0.0, ya, u, p, v, gx);               // :16
    P::template U<HM_NONE>(dummy, 0.0, ya, u, p, v, gu);               // :19
#pragma unroll
    for (int c = 0; c < n; ++c) { f[c] = Fo[c]; f[n + c] = r * y[n + c] - gx[c]; }   // :14,:17
#pragma unroll
    for (int c = 0; c < nc; ++c) f[2 * n + c] = gu[c];
}

template <class P>
__device__ __forceinline__ void equilibrium_body(const EqArgs &a)
{
    constexpr int n = P::n, nc = P::nc, np = P::np, NY = 2 * n + nc;
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.nsets) return;
    double p[np > 0 ? np : 1];
    for (int k = 0; k < np; ++k) p[k] = a.params[k + (size_t)np * (a.nparams > 1 ? b : 0)];
    const double r = a.r[a.nr > 1 ? b : 0];
    double y[NY], f[NY], yt[NY], ft[NY], fm[NY], g[NY], dy[NY], lo[NY], hi[NY];
    double Jm[NY][NY], A[NY][NY], Lc[NY][NY];
    for (int k = 0; k < NY; ++k) {
        lo[k] = a.lb ? a.lb[k] : -1.0e308 * 10.0;
        hi[k] = a.ub ? a.ub[k] : 1.0e308 * 10.0;
        y[k] = fmin(hi[k], fmax(lo[k], a.guess[k + (size_t)NY * (a.nguess > 1 ? b : 0)]));
    }
    eq_residual<P>(y, p, r, f);
    double Fv = 0.0;
    for (int k = 0; k < NY; ++k) Fv += f[k] * f[k];
    double mu = -1.0, nu = 2.0;
    int flag = 0, it = 0;
    bool need_jac = true;
    for (it = 0; it < a.max_iter; ++it) {
        if (!(Fv == Fv) || Fv > 1.0e300) { flag = -1; break; }
        if (Fv <= a.tol * a.tol) { flag = 1; break; }
        if (need_jac) {
            for (int k = 0; k < NY; ++k) yt[k] = y[k];
            for (int k = 0; k < NY; ++k) {
                const double d = 6.0e-6 * fmax(1.0, fabs(y[k]));     // ~eps^(1/3): central differences
                yt[k] = y[k] + d;
                eq_residual<P>(yt, p, r, ft);
                yt[k] = y[k] - d;
                eq_residual<P>(yt, p, r, fm);
                yt[k] = y[k];
                const double inv = 1.0 / (2.0 * d);
                for (int q = 0; q < NY; ++q) Jm[q][k] = (ft[q] - fm[q]) * inv;
            }
            double amax = 0.0;
            for (int i = 0; i < NY; ++i) {
                double s = 0.0;
                for (int q = 0; q < NY; ++q) s += Jm[q][i] * f[q];
                g[i] = s;
                for (int j = 0; j <= i; ++j) {
                    double t = 0.0;
                    for (int q = 0; q < NY; ++q) t += Jm[q][i] * Jm[q][j];
                    A[i][j] = A[j][i] = t;
                }
                amax = fmax(amax, A[i][i]);
            }
            if (mu < 0.0) mu = 1.0e-3 * fmax(amax, 1.0e-300);
            need_jac = false;
        }
        // active set: a variable on a bound whose gradient points outwards is held there for this step
        bool act[NY];
        for (int k = 0; k < NY; ++k) act[k] = (y[k] <= lo[k] && g[k] > 0.0) || (y[k] >= hi[k] && g[k] < 0.0);
        // (A + mu I) dy = -g on the free variables by Cholesky; a failed factorisation raises the damping
        bool spd = true;
        for (int i = 0; i < NY && spd; ++i) {
            for (int j = 0; j <= i; ++j) {
                double s = (i == j) ? (act[i] ? 1.0 : A[i][i]) + mu : ((act[i] || act[j]) ? 0.0 : A[i][j]);
                for (int q = 0; q < j; ++q) s -= Lc[i][q] * Lc[j][q];
                if (i == j) {
                    if (!(s > 0.0)) { spd = false; break; }
                    Lc[i][i] = sqrt(s);
                } else {
                    Lc[i][j] = s / Lc[j][j];
                }
            }
        }
        if (!spd) { mu *= nu; nu *= 2.0; continue; }
        for (int i = 0; i < NY; ++i) {
            double s = act[i] ? 0.0 : -g[i];
            for (int q = 0; q < i; ++q) s -= Lc[i][q] * dy[q];
            dy[i] = s / Lc[i][i];
        }
        for (int i = NY - 1; i >= 0; --i) {
            double s = dy[i];
            for (int q = i + 1; q < NY; ++q) s -= Lc[q][i] * dy[q];
            dy[i] = s / Lc[i][i];
        }
        bool tiny = true;
        for (int k = 0; k < NY; ++k) {
            yt[k] = fmin(hi[k], fmax(lo[k], y[k] + dy[k]));           // projected step
            dy[k] = yt[k] - y[k];
            if (fabs(dy[k]) > a.tol * (1.0 + fabs(y[k]))) tiny = false;
        }
        eq_residual<P>(yt, p, r, ft);
        double Fn = 0.0;
        for (int k = 0; k < NY; ++k) Fn += ft[k] * ft[k];
        if (Fn < Fv) {
            // gain ratio against the damped model: F - model = dy'(mu dy - g) for the unprojected step
            double pred = 0.0;
            for (int k = 0; k < NY; ++k) pred += dy[k] * (mu * dy[k] - (act[k] ? 0.0 : g[k]));
            const double rho = pred > 0.0 ? (Fv - Fn) / (2.0 * pred) : 1.0;   // F = |f|^2 = 2 * (|f|^2 / 2)
            const double c = 2.0 * rho - 1.0;
            mu *= fmax(1.0 / 3.0, 1.0 - c * c * c);
            nu = 2.0;
            for (int k = 0; k < NY; ++k) { y[k] = yt[k]; f[k] = ft[k]; }
            Fv = Fn;
            need_jac = true;
            if (tiny) { flag = 1; ++it; break; }
        } else {
            if (tiny) { flag = 1; break; }     // no further decrease within the step tolerance
            mu *= nu;
            nu *= 2.0;
        }
    }
    for (int k = 0; k < NY; ++k) {
        a.y[k + (size_t)NY * b] = y[k];
        if (a.residual) a.residual[k + (size_t)NY * b] = f[k];
    }
    a.resnorm[b] = Fv;
    a.exitflag[b] = flag;
    if (a.iters) a.iters[b] = it;
}

}  // namespace OCS_NS

// Kernel entry points.  Each macro emits one extern "C" __global__ wrapper; the host looks the
// kernels up by name (built-in problems: OCS_DEFINE_KERNELS instantiates every variant at build
// time; generated problems: nvrtc_problem.cu compiles single variants on demand).
#define OCS_KERNEL_(prefix, body_call, tag, UM, BM, PB, LB)                                            \
    extern "C" __global__ void __launch_bounds__(OCS_THREADS, LB)                                      \
        prefix##tag##_u##UM##_b##BM##_p##PB(const OCS_NS::SweepArgs a)                                 \
    {                                                                                                  \
        extern __shared__ double2 ocs_smem2[];                                                         \
        OCS_NS::body_call(a, reinterpret_cast<double *>(ocs_smem2));                                   \
    }
#define OCS_COMMA ,
// trajectories per thread / min blocks per SM of a sweep variant (the host mirrors these in kernel_tpt())
#define OCS_FWD_T(UM, BM, PB) ((BM) == 2 ? 1 : ((PB) != 0 ? OCS_TPT_FWD_PB : ((UM) == 1 ? OCS_TPT_FWD_COEF : OCS_TPT_FWD)))
// explicit register caps (blocks per SM) of the variants the tuned defaults derive from the macros above
#ifndef OCS_LB_FWD_DENSE
#define OCS_LB_FWD_DENSE 1      // forward sweep evaluating a dense (Chebyshev) basis per trajectory
#endif
#ifndef OCS_LB_FWD_GRID
#define OCS_LB_FWD_GRID 4       // forward sweep reading its control from the per-trajectory grid
#endif
#define OCS_FWD_LB(UM, BM, PB) ((BM) == 2 ? OCS_LB_FWD_DENSE : ((PB) != 0 ? OCS_MINBLOCKS_FWD_PB : ((UM) == 2 ? OCS_LB_FWD_GRID : ((UM) == 1 ? OCS_MINBLOCKS_FWD_COEF : OCS_MINBLOCKS_FWD))))
#define OCS_REV_T(UM, BM, PB) ((BM) == 2 ? 1 : ((PB) != 0 ? OCS_TPT_REV_PB : ((UM) == 1 ? OCS_TPT_REV_COEF : OCS_TPT_REV)))
// (the variants without a sparse basis carry the generic dJdu-grid path and need a few registers more: one block less)
#ifndef OCS_LB_REV_GRID
#define OCS_LB_REV_GRID (OCS_MINBLOCKS_REV > 2 ? OCS_MINBLOCKS_REV - 2 : OCS_MINBLOCKS_REV)   // gradient sweep over the per-trajectory control grid
#endif
#define OCS_REV_LB(UM, BM, PB) ((BM) == 2 ? 1 : ((PB) != 0 ? ((UM) == 2 && OCS_MINBLOCKS_REV_PB > 1 ? OCS_MINBLOCKS_REV_PB - 1 : OCS_MINBLOCKS_REV_PB) : ((UM) == 1 ? OCS_MINBLOCKS_REV_COEF : ((BM) == 0 && (UM) == 2 ? OCS_LB_REV_GRID : ((BM) == 0 && OCS_MINBLOCKS_REV > 2 ? OCS_MINBLOCKS_REV - 1 : OCS_MINBLOCKS_REV)))))
#define OCS_FWD_KERNEL(P, tag, UM, BM, PB, KD) \
    OCS_KERNEL_(ocs_fwd_, forward_multi_body<P OCS_COMMA UM OCS_COMMA BM OCS_COMMA(PB != 0) OCS_COMMA KD OCS_COMMA OCS_FWD_T(UM, BM, PB)>, tag, UM, BM, PB, OCS_FWD_LB(UM, BM, PB))
#define OCS_FWDG_KERNEL(P, tag, UM, BM, PB, KD) \
    OCS_KERNEL_(ocs_fwdg_, forward_body<P OCS_COMMA UM OCS_COMMA BM OCS_COMMA(PB != 0) OCS_COMMA KD OCS_COMMA true>, tag, UM, BM, PB, 1)
#define OCS_REV_KERNEL(P, tag, UM, BM, PB, KD) \
    OCS_KERNEL_(ocs_rev_, reverse_multi_body<P OCS_COMMA UM OCS_COMMA BM OCS_COMMA(PB != 0) OCS_COMMA KD OCS_COMMA OCS_REV_T(UM, BM, PB)>, tag, UM, BM, PB, OCS_REV_LB(UM, BM, PB))
#define OCS_REVG_KERNEL(P, tag, UM, BM, PB, KD) \
    OCS_KERNEL_(ocs_revg_, reverse_body<P OCS_COMMA UM OCS_COMMA BM OCS_COMMA(PB != 0) OCS_COMMA KD OCS_COMMA 0 OCS_COMMA true>, tag, UM, BM, PB, 1)
#define OCS_FB_KERNEL(P, tag, PB) \
    OCS_KERNEL_(ocs_fbrev_, reverse_body<P OCS_COMMA 2 OCS_COMMA 0 OCS_COMMA(PB != 0) OCS_COMMA 0 OCS_COMMA 1 OCS_COMMA false>, tag, 2, 0, PB, ((PB) != 0 ? OCS_MINBLOCKS_FB_PB : OCS_MINBLOCKS_FB))
#define OCS_FBC_KERNEL(P, tag, PB) \
    OCS_KERNEL_(ocs_fbcont_, reverse_body<P OCS_COMMA 2 OCS_COMMA 0 OCS_COMMA(PB != 0) OCS_COMMA 0 OCS_COMMA 2 OCS_COMMA false>, tag, 2, 0, PB, ((PB) != 0 ? OCS_MINBLOCKS_FB_PB : OCS_MINBLOCKS_FB))
#define OCS_FBCG_KERNEL(P, tag, PB) \
    OCS_KERNEL_(ocs_fbcontg_, reverse_body<P OCS_COMMA 2 OCS_COMMA 0 OCS_COMMA(PB != 0) OCS_COMMA 0 OCS_COMMA 2 OCS_COMMA true>, tag, 2, 0, PB, 1)
#define OCS_TT_KERNEL(P, tag)                                                                         \
    extern "C" __global__ void ocs_tt_##tag(double *gprec, int G, int hm, const double *p)             \
    {                                                                                                  \
        OCS_NS::table_body<P>(gprec, G, hm, p);                                                        \
    }                                                                                                  \
    extern "C" __global__ void ocs_fold_##tag(const double *gprec, const double *steprec, double *foldrec, \
                                              int N, int hm)                                           \
    {                                                                                                  \
        if (hm == 1) OCS_NS::fold_rows<P, 1>(gprec, steprec, foldrec, N);                              \
        else if (hm == 2) OCS_NS::fold_rows<P, 2>(gprec, steprec, foldrec, N);                         \
    }                                                                                                  \
    extern "C" __global__ void ocs_info_##tag(int *o)                                                  \
    {                                                                                                  \
        o[0] = P::n; o[1] = P::nc; o[2] = P::np;                                                       \
        o[3] = P::ntt(0); o[4] = P::ntt(1) | (OCS_TPT_FWD_PB << 16) | (OCS_TPT_REV_PB << 20);          \
        o[5] = P::ntt(2) | (OCS_TPT_FWD_COEF << 16) | (OCS_RING_PB << 24);                              \
        o[6] = OCS_TILE | (OCS_RING << 16) | (OCS_FB_RING << 24) | (OCS_ASYNC_STAGE << 28); o[7] = OCS_THREADS | (OCS_TPT_REV << 16) | (OCS_TPT_FWD << 24) | (OCS_TPT_REV_COEF << 28); \
    }

#define OCS_EQ_KERNEL(P, tag)                                                                         \
    extern "C" __global__ void __launch_bounds__(128) ocs_eq_##tag(const OCS_NS::EqArgs a)            \
    {                                                                                                  \
        OCS_NS::equilibrium_body<P>(a);                                                               \
    }

#define OCS_DEFINE_KERNELS_PB(P, tag, PB, KD)   \
    OCS_FWD_KERNEL(P, tag, 0, 0, PB, KD)        \
    OCS_FWD_KERNEL(P, tag, 1, 1, PB, KD)        \
    OCS_FWD_KERNEL(P, tag, 1, 2, PB, KD)        \
    OCS_FWD_KERNEL(P, tag, 2, 0, PB, KD)        \
    OCS_FWDG_KERNEL(P, tag, 0, 0, PB, KD)       \
    OCS_FWDG_KERNEL(P, tag, 1, 1, PB, KD)       \
    OCS_FWDG_KERNEL(P, tag, 1, 2, PB, KD)       \
    OCS_FWDG_KERNEL(P, tag, 2, 0, PB, KD)       \
    OCS_REV_KERNEL(P, tag, 0, 0, PB, KD)        \
    OCS_REV_KERNEL(P, tag, 0, 1, PB, KD)        \
    OCS_REV_KERNEL(P, tag, 1, 1, PB, KD)        \
    OCS_REV_KERNEL(P, tag, 2, 0, PB, KD)        \
    OCS_REVG_KERNEL(P, tag, 0, 0, PB, KD)       \
    OCS_REVG_KERNEL(P, tag, 1, 1, PB, KD)       \
    OCS_REVG_KERNEL(P, tag, 1, 2, PB, KD)       \
    OCS_REVG_KERNEL(P, tag, 2, 0, PB, KD)       \
    OCS_FB_KERNEL(P, tag, PB)                   \
    OCS_FBC_KERNEL(P, tag, PB)                  \
    OCS_FBCG_KERNEL(P, tag, PB)

#define OCS_DEFINE_KERNELS(P, tag, KD)   \
    OCS_DEFINE_KERNELS_PB(P, tag, 0, KD) \
    OCS_DEFINE_KERNELS_PB(P, tag, 1, KD) \
    OCS_TT_KERNEL(P, tag)                \
    OCS_EQ_KERNEL(P, tag)
