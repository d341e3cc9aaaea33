// ocs_multi.cu -- several B200s behind ONE handle of the C ABI.
//
// The reference's host is a single MATLAB process that constructs one integrator object
// (functions/single_shooting.m:41-45; the object owns the state cache, Integrator/RK4Integrator.m:8-10).
// ocs_b200_create_multi keeps that shape: one handle, one caller thread, and inside it one ordinary
// single-device handle per GPU with a worker thread each.  Trajectories are independent, so a batch is cut
// into contiguous shards (the first Bt mod ndev shards take one trajectory more) and every host-buffer entry
// point forwards shard k to device k with plain pointer offsets -- host arrays are MATLAB column-major with
// the trajectory index last, so a shard of any argument or result is one contiguous block and every device
// reads / writes the caller's arrays directly.  No collective is on that path.
//
// The only exchange is the gather of the small per-trajectory results J and x(T) to the ROOT device
// (devices[0]) for device-resident pipelines (ocs_b200_gather_results_dev): NCCL point-to-point over NVLink
// (ncclCommInitAll in this process, grouped ncclSend / ncclRecv -- a gather to root, not an all-gather), or
// peer copies on the copy engines (cudaMemcpyPeerAsync) when libnccl cannot be loaded or OCS_B200_GATHER=p2p.
#include "ocs_internal.h"

#include <dlfcn.h>

#include <condition_variable>
#include <functional>
#include <memory>
#include <thread>

#if __has_include(<nccl.h>)
#include <nccl.h>
#define OCS_HAVE_NCCL_H 1
#else
#define OCS_HAVE_NCCL_H 0
#endif

namespace {

// One thread per device: tasks posted to it run in order; wait() returns when the queue is empty.
class Worker {
public:
    explicit Worker(int device) : device_(device), th_([this] { run(); }) {}
    ~Worker()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        th_.join();
    }
    void post(std::function<void()> f)
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            q_.push_back(std::move(f));
            ++pending_;
        }
        cv_.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
    }

private:
    void run()
    {
        cudaSetDevice(device_);
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                f = std::move(q_.front());
                q_.erase(q_.begin());
            }
            f();
            {
                std::lock_guard<std::mutex> lk(mu_);
                --pending_;
            }
            done_.notify_all();
        }
    }
    int device_;
    bool stop_ = false;
    int pending_ = 0;
    std::vector<std::function<void()>> q_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::thread th_;
};

#if OCS_HAVE_NCCL_H
struct Nccl {
    void *lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load()
    {
        const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
        for (int k = 0; names[k] && !lib; ++k) lib = dlopen(names[k], RTLD_NOW | RTLD_LOCAL);
        if (!lib) return false;
#define SYM(f, n)                              \
    *(void **)(&f) = dlsym(lib, n);            \
    if (!f) return false;
        SYM(CommInitAll, "ncclCommInitAll")
        SYM(CommDestroy, "ncclCommDestroy")
        SYM(GroupStart, "ncclGroupStart")
        SYM(GroupEnd, "ncclGroupEnd")
        SYM(Send, "ncclSend")
        SYM(Recv, "ncclRecv")
        SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
        return true;
    }
};
#endif

}  // namespace

struct ocs_multi {
    std::vector<ocs_b200_handle *> dev;
    std::vector<int> device;
    std::vector<std::unique_ptr<Worker>> worker;
    long Bt = 0;                 // resident batch and its shard offsets (ndev + 1 entries)
    std::vector<long> off;
    long cB = 0;                 // shards of the last per-trajectory control / parameter upload
    std::vector<long> coff;
    std::vector<cudaEvent_t> ev;     // one per device: "my part of the gather has been issued"
    int gather_mode = -1;            // -1 undecided, 0 peer copies, 1 NCCL
#if OCS_HAVE_NCCL_H
    Nccl nccl;
    std::vector<ncclComm_t> comm;
#endif
    int ndev() const { return (int)dev.size(); }
};

namespace {

typedef ocs_multi M;

std::vector<long> split(long n, int parts)
{
    std::vector<long> off(parts + 1, 0);
    const long base = n / parts, rem = n % parts;
    for (int k = 0; k < parts; ++k) off[k + 1] = off[k] + base + (k < rem ? 1 : 0);
    return off;
}

// f(k, child) on every device's worker thread at once; the first non-zero status wins and its message becomes the
// multi handle's.  Shards that hold no trajectory (fewer trajectories than devices) are skipped when `skip_empty`.
template <class F>
int fan(ocs_b200_handle *h, F f, const std::vector<long> *off = nullptr)
{
    M *m = h->multi;
    const int nd = m->ndev();
    std::vector<int> rc(nd, 0);
    for (int k = 0; k < nd; ++k) {
        if (off && (*off)[k + 1] == (*off)[k]) continue;
        ocs_b200_handle *c = m->dev[k];
        m->worker[k]->post([&rc, k, c, &f] {
            try {
                rc[k] = f(k, c);
            } catch (const std::exception &ex) {
                rc[k] = c->fail(OCS_B200_ERR_NOMEM, "exception: %s", ex.what());
            }
        });
    }
    for (int k = 0; k < nd; ++k) m->worker[k]->wait();
    for (int k = 0; k < nd; ++k)
        if (rc[k] != 0) return h->fail(rc[k], "device %d: %s", m->device[k], ocs_b200_last_error(m->dev[k]));
    return 0;
}

template <class T>
T *at(T *p, size_t stride, long b)
{
    return p ? p + stride * (size_t)b : nullptr;
}

void dims(const ocs_b200_handle *h, int &n, int &nc, int &np, int &K, int &N, int &N2)
{
    int32_t d[8];
    ocs_b200_get_dims(h->multi->dev[0], d);
    n = d[0]; nc = d[1]; np = d[2]; K = d[3]; N = d[4]; N2 = d[5];
}

}  // namespace

// =============================================================================== lifetime

extern "C" int ocs_b200_create_multi(int N, const double *tspan, int N2, const double *tspan_extra, int ndev,
                                     const int *devices, int math_mode, ocs_b200_handle **out)
{
    if (!out) return OCS_B200_ERR_ARG;
    *out = nullptr;
    if (ndev < 1 || !devices) return OCS_B200_ERR_ARG;
    for (int a = 0; a < ndev; ++a)
        for (int b = a + 1; b < ndev; ++b)
            if (devices[a] == devices[b]) return OCS_B200_ERR_ARG;        // one shard per physical device
    ocs_b200_handle *h = new (std::nothrow) ocs_b200_handle();
    if (!h) return OCS_B200_ERR_NOMEM;
    M *m = new (std::nothrow) M();
    if (!m) {
        delete h;
        return OCS_B200_ERR_NOMEM;
    }
    h->multi = m;
    h->device = devices[0];
    h->math_mode = math_mode;
    for (int k = 0; k < ndev; ++k) {
        ocs_b200_handle *c = nullptr;
        const int rc = ocs_b200_create(N, tspan, N2, tspan_extra, devices[k], math_mode, &c);
        if (rc != 0) {          // the create-time message (ocs_b200_last_error(NULL)) is already the failing device's
            for (ocs_b200_handle *d : m->dev) ocs_b200_destroy(d);
            delete m;
            delete h;
            return rc;
        }
        m->dev.push_back(c);
        m->device.push_back(devices[k]);
    }
    for (int k = 0; k < ndev; ++k) m->worker.emplace_back(new Worker(devices[k]));
    m->ev.assign(ndev, nullptr);
    // peers map each other's memory so that the root gather can run on the copy engines / NVLink
    for (int a = 0; a < ndev; ++a) {
        cudaSetDevice(devices[a]);
        cudaEventCreateWithFlags(&m->ev[a], cudaEventDisableTiming);
        for (int b = 0; b < ndev; ++b) {
            int can = 0;
            if (a != b && cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) == cudaSuccess && can)
                if (cudaDeviceEnablePeerAccess(devices[b], 0) != cudaSuccess) cudaGetLastError();   // already enabled is fine
        }
    }
    cudaGetLastError();
    h->has_tail = N2 > 0;
    *out = h;
    return 0;
}

int ocs_multi_destroy(ocs_b200_handle *h)
{
    M *m = h->multi;
    fan(h, [](int, ocs_b200_handle *c) { return ocs_b200_synchronize(c); });
#if OCS_HAVE_NCCL_H
    for (ncclComm_t c : m->comm)
        if (c) m->nccl.CommDestroy(c);
#endif
    for (int k = 0; k < m->ndev(); ++k) {
        ocs_b200_handle *c = m->dev[k];
        cudaEvent_t e = m->ev[k];
        m->worker[k]->post([c, e] {
            if (e) cudaEventDestroy(e);
            ocs_b200_destroy(c);
        });
    }
    for (auto &w : m->worker) w->wait();
    m->worker.clear();
    delete m;
    h->multi = nullptr;
    delete h;
    return 0;
}

const ocs_b200_handle *ocs_multi_first(const ocs_b200_handle *h) { return h->multi->dev[0]; }
long ocs_multi_batch(const ocs_b200_handle *h) { return h->multi->Bt; }

int ocs_multi_device_info(const ocs_b200_handle *h, int *ndev, int *devices, long *offsets, int cap)
{
    const M *m = h->multi;
    if (ndev) *ndev = m->ndev();
    for (int k = 0; k < m->ndev() && k < cap; ++k) {
        if (devices) devices[k] = m->device[k];
        if (offsets) offsets[k] = (size_t)k < m->off.size() ? m->off[k] : 0;
    }
    if (offsets && m->ndev() < cap) offsets[m->ndev()] = m->Bt;
    return 0;
}

// =============================================================================== batch-invariant state: broadcast

int ocs_multi_get_grid(const ocs_b200_handle *h, int segment, double *t, double *hstep)
{
    return ocs_b200_get_grid(h->multi->dev[0], segment, t, hstep);
}
int ocs_multi_set_problem_builtin(ocs_b200_handle *h, int id)
{
    h->multi->Bt = 0;
    return fan(h, [=](int, ocs_b200_handle *c) { return ocs_b200_set_problem_builtin(c, id); });
}
int ocs_multi_set_problem_source(ocs_b200_handle *h, const char *src, int n, int nc, int np, const char *inc)
{
    h->multi->Bt = 0;
    return fan(h, [=](int, ocs_b200_handle *c) { return ocs_b200_set_problem_source(c, src, n, nc, np, inc); });
}
int ocs_multi_set_tail_control(ocs_b200_handle *h, const double *uStar, int nc)
{
    return fan(h, [=](int, ocs_b200_handle *c) { return ocs_b200_set_tail_control(c, uStar, nc); });
}
int ocs_multi_set_control_basis(ocs_b200_handle *h, int K, const double *B)
{
    h->multi->Bt = 0;
    return fan(h, [=](int, ocs_b200_handle *c) { return ocs_b200_set_control_basis(c, K, B); });
}
int ocs_multi_get_control_index(const ocs_b200_handle *h, int32_t *k_index, int32_t *nnz)
{
    return ocs_b200_get_control_index(h->multi->dev[0], k_index, nnz);
}
int ocs_multi_set_control_ufunc(ocs_b200_handle *h, int kind, int nknots, const double *knots)
{
    return fan(h, [=](int, ocs_b200_handle *c) { return ocs_b200_set_control_ufunc(c, kind, nknots, knots); });
}
int ocs_multi_set_fb_costate(ocs_b200_handle *h, int mode)
{
    return fan(h, [=](int, ocs_b200_handle *c) { return ocs_b200_set_fb_costate(c, mode); });
}
int ocs_multi_set_solve_control_bounds(ocs_b200_handle *h, const double *bounds, int nc, double rho0, double feas_tol, int max_outer)
{
    return fan(h, [=](int, ocs_b200_handle *c) { return ocs_b200_set_solve_control_bounds(c, bounds, nc, rho0, feas_tol, max_outer); });
}
int ocs_multi_get_solve_violation(ocs_b200_handle *h, double *viol)
{
    M *m = h->multi;
    if (m->Bt < 1 || m->off.empty()) return h->fail(OCS_B200_ERR_STATE, "no constrained solve on this handle");
    return fan(h, [&](int k, ocs_b200_handle *c) { return ocs_b200_get_solve_violation(c, viol + m->off[k]); }, &m->off);
}
int ocs_multi_set_profiling(ocs_b200_handle *h, int on)
{
    return fan(h, [=](int, ocs_b200_handle *c) { return ocs_b200_set_profiling(c, on); });
}
int ocs_multi_synchronize(ocs_b200_handle *h)
{
    return fan(h, [](int, ocs_b200_handle *c) { return ocs_b200_synchronize(c); });
}

// =============================================================================== per-trajectory inputs: sharded

// A batch of `nbatch` columns that arrives BEFORE the sweep it belongs to (parameters, coefficients, controls) is
// cut exactly like the batch of the sweep will be: split() depends on the column count only.
int ocs_multi_set_params(ocs_b200_handle *h, const double *params, int np, long nbatch)
{
    if (nbatch <= 1 || !params) return fan(h, [=](int, ocs_b200_handle *c) { return ocs_b200_set_params(c, params, np, nbatch); });
    const std::vector<long> off = split(nbatch, h->multi->ndev());
    return fan(h, [&](int k, ocs_b200_handle *c) { return ocs_b200_set_params(c, at(params, np, off[k]), np, off[k + 1] - off[k]); }, &off);
}
int ocs_multi_set_control_v(ocs_b200_handle *h, const double *v, long nbatch)
{
    if (nbatch <= 1 || !v) return fan(h, [=](int, ocs_b200_handle *c) { return ocs_b200_set_control_v(c, v, nbatch); });
    int n, nc, np, K, N, N2;
    dims(h, n, nc, np, K, N, N2);
    M *m = h->multi;
    m->coff = split(nbatch, m->ndev());
    m->cB = nbatch;
    const std::vector<long> &off = m->coff;
    return fan(h, [&](int k, ocs_b200_handle *c) { return ocs_b200_set_control_v(c, at(v, (size_t)nc * K, off[k]), off[k + 1] - off[k]); }, &off);
}
int ocs_multi_set_control_u(ocs_b200_handle *h, const double *u, long nbatch)
{
    if (nbatch <= 1 || !u) return fan(h, [=](int, ocs_b200_handle *c) { return ocs_b200_set_control_u(c, u, nbatch); });
    int n, nc, np, K, N, N2;
    dims(h, n, nc, np, K, N, N2);
    M *m = h->multi;
    m->coff = split(nbatch, m->ndev());
    m->cB = nbatch;
    const std::vector<long> &off = m->coff;
    const size_t st = (size_t)nc * (2 * N + 1);
    return fan(h, [&](int k, ocs_b200_handle *c) { return ocs_b200_set_control_u(c, at(u, st, off[k]), off[k + 1] - off[k]); }, &off);
}

// trajectories [b0, b0+nb) of a sharded batch -> (device, local range) pieces; f(k, child, local b0, count, global b0)
template <class F>
int ranged(ocs_b200_handle *h, const std::vector<long> &off, long b0, long nb, F f)
{
    M *m = h->multi;
    if (off.empty() || b0 < 0 || nb < 1 || b0 + nb > off.back()) return h->fail(OCS_B200_ERR_ARG, "trajectory range out of bounds");
    std::vector<long> piece(m->ndev() + 1, 0);        // fan() skips devices whose piece is empty
    std::vector<long> lo(m->ndev()), cnt(m->ndev());
    for (int k = 0; k < m->ndev(); ++k) {
        const long a = std::max(b0, off[k]), e = std::min(b0 + nb, off[k + 1]);
        lo[k] = a;
        cnt[k] = e > a ? e - a : 0;
        piece[k + 1] = piece[k] + cnt[k];
    }
    return fan(h, [&](int k, ocs_b200_handle *c) { return f(k, c, lo[k] - off[k], cnt[k], lo[k]); }, &piece);
}

int ocs_multi_get_control_u(ocs_b200_handle *h, long b0, long nb, double *u)
{
    M *m = h->multi;
    int n, nc, np, K, N, N2;
    dims(h, n, nc, np, K, N, N2);
    const int mode = m->dev[0]->ctrl_mode;
    if (mode == ocs_detail::CTRL_V_SHARED || mode == ocs_detail::CTRL_U_SHARED || mode == ocs_detail::CTRL_NONE)
        return ocs_b200_get_control_u(m->dev[0], b0, nb, u) == 0 ? 0 : h->fail(OCS_B200_ERR_STATE, "%s", ocs_b200_last_error(m->dev[0]));
    // per-trajectory control: after fb_sweep / single_shooting it is sharded like the batch, otherwise like its upload
    const bool solver = m->Bt > 0 && m->dev[0]->ctrl_nbatch == m->off[1] - m->off[0] && m->cB != m->Bt;
    const std::vector<long> &off = (solver || m->coff.empty()) ? m->off : m->coff;
    const size_t st = (size_t)nc * (2 * N + 1);
    return ranged(h, off, b0, nb, [&](int, ocs_b200_handle *c, long l0, long cnt, long g0) {
        return ocs_b200_get_control_u(c, l0, cnt, u + st * (size_t)(g0 - b0));
    });
}

// =============================================================================== sweeps

static void start_batch(M *m, long Bt)
{
    m->off = split(Bt, m->ndev());
    m->Bt = Bt;
}

int ocs_multi_compute_states(ocs_b200_handle *h, const double *x0, long Bt, double *J, double *xT, int32_t *status)
{
    if (!x0 || Bt < 1) return h->fail(OCS_B200_ERR_ARG, "bad x0 / batch size");
    M *m = h->multi;
    int n, nc, np, K, N, N2;
    dims(h, n, nc, np, K, N, N2);
    m->Bt = 0;
    const std::vector<long> off = split(Bt, m->ndev());
    int rc = fan(h, [&](int k, ocs_b200_handle *c) {
        return ocs_b200_compute_states(c, at(x0, n, off[k]), off[k + 1] - off[k], at(J, 1, off[k]), at(xT, n, off[k]), at(status, 1, off[k]));
    }, &off);
    if (rc == 0) start_batch(m, Bt);
    return rc;
}

int ocs_multi_compute_adjoints(ocs_b200_handle *h, const double *lamT, long lamT_nbatch, double *dJdv, double *lam0)
{
    M *m = h->multi;
    if (m->Bt < 1) return h->fail(OCS_B200_ERR_STATE, "compute_adjoints needs a preceding compute_states on this handle");
    if (lamT && lamT_nbatch != 1 && lamT_nbatch != m->Bt) return h->fail(OCS_B200_ERR_ARG, "lamT must have 1 or Bt columns");
    int n, nc, np, K, N, N2;
    dims(h, n, nc, np, K, N, N2);
    const int mm = n + 1;
    const std::vector<long> &off = m->off;
    return fan(h, [&](int k, ocs_b200_handle *c) {
        const bool per = lamT && lamT_nbatch > 1;
        return ocs_b200_compute_adjoints(c, per ? at(lamT, mm, off[k]) : lamT, per ? off[k + 1] - off[k] : lamT_nbatch,
                                         at(dJdv, (size_t)nc * K, off[k]), at(lam0, mm, off[k]));
    }, &off);
}

int ocs_multi_evaluate(ocs_b200_handle *h, const double *x0, long Bt, const double *lamT, long lamT_nbatch, double *J,
                       double *xT, int32_t *status, double *dJdv, double *lam0)
{
    if (!x0 || Bt < 1) return h->fail(OCS_B200_ERR_ARG, "bad x0 / batch size");
    if (lamT && lamT_nbatch != 1 && lamT_nbatch != Bt) return h->fail(OCS_B200_ERR_ARG, "lamT must have 1 or Bt columns");
    M *m = h->multi;
    int n, nc, np, K, N, N2;
    dims(h, n, nc, np, K, N, N2);
    const int mm = n + 1;
    m->Bt = 0;
    const std::vector<long> off = split(Bt, m->ndev());
    int rc = fan(h, [&](int k, ocs_b200_handle *c) {
        const bool per = lamT && lamT_nbatch > 1;
        return ocs_b200_evaluate(c, at(x0, n, off[k]), off[k + 1] - off[k], per ? at(lamT, mm, off[k]) : lamT,
                                 per ? off[k + 1] - off[k] : lamT_nbatch, at(J, 1, off[k]), at(xT, n, off[k]), at(status, 1, off[k]),
                                 at(dJdv, (size_t)nc * K, off[k]), at(lam0, mm, off[k]));
    }, &off);
    if (rc == 0) start_batch(m, Bt);
    return rc;
}

int ocs_multi_fetch_states(ocs_b200_handle *h, int segment, long b0, long nb, double *x)
{
    M *m = h->multi;
    if (m->Bt < 1) return h->fail(OCS_B200_ERR_STATE, "fetch_states needs a preceding compute_states");
    int n, nc, np, K, N, N2;
    dims(h, n, nc, np, K, N, N2);
    const size_t st = (size_t)(n + 1) * ((segment == 0 ? N : N2) + 1);
    return ranged(h, m->off, b0, nb, [&](int, ocs_b200_handle *c, long l0, long cnt, long g0) {
        return ocs_b200_fetch_states(c, segment, l0, cnt, at(x, st, g0 - b0));
    });
}

int ocs_multi_fetch_adjoints(ocs_b200_handle *h, int segment, long b0, long nb, double *lam, double *dJdu)
{
    M *m = h->multi;
    if (m->Bt < 1) return h->fail(OCS_B200_ERR_STATE, "fetch_adjoints needs a preceding compute_states");
    int n, nc, np, K, N, N2;
    dims(h, n, nc, np, K, N, N2);
    const int Ns = segment == 0 ? N : N2;
    const size_t sl = (size_t)(n + 1) * (Ns + 1), su = (size_t)nc * (2 * Ns + 1);
    return ranged(h, m->off, b0, nb, [&](int, ocs_b200_handle *c, long l0, long cnt, long g0) {
        return ocs_b200_fetch_adjoints(c, segment, l0, cnt, at(lam, sl, g0 - b0), at(dJdu, su, g0 - b0));
    });
}

int ocs_multi_eval_solution(ocs_b200_handle *h, int segment, long b0, long nb, int nq, const double *tq, double *x, double *lam,
                            double *u)
{
    M *m = h->multi;
    if (m->Bt < 1) return h->fail(OCS_B200_ERR_STATE, "eval_solution needs a preceding compute_states");
    int n, nc, np, K, N, N2;
    dims(h, n, nc, np, K, N, N2);
    return ranged(h, m->off, b0, nb, [&](int, ocs_b200_handle *c, long l0, long cnt, long g0) {
        return ocs_b200_eval_solution(c, segment, l0, cnt, nq, tq, at(x, (size_t)n * nq, g0 - b0), at(lam, (size_t)n * nq, g0 - b0),
                                      at(u, (size_t)nc * nq, g0 - b0));
    });
}

int ocs_multi_fb_sweep(ocs_b200_handle *h, const double *x0, long Bt, const double *u0, long u0_nbatch, const double *bounds,
                       double uRelTol, double uAbsTol, int nSWEEPS, double relax, double *J, int32_t *sweeps,
                       int32_t *converged, double *max_change)
{
    if (!x0 || Bt < 1 || (u0 && u0_nbatch != 1 && u0_nbatch != Bt)) return h->fail(OCS_B200_ERR_ARG, "bad fb_sweep arguments");
    M *m = h->multi;
    int n, nc, np, K, N, N2;
    dims(h, n, nc, np, K, N, N2);
    const size_t su = (size_t)nc * (2 * N + 1);
    m->Bt = 0;
    const std::vector<long> off = split(Bt, m->ndev());
    int rc = fan(h, [&](int k, ocs_b200_handle *c) {
        const bool per = u0 && u0_nbatch > 1;
        return ocs_b200_fb_sweep(c, at(x0, n, off[k]), off[k + 1] - off[k], per ? at(u0, su, off[k]) : u0,
                                 per ? off[k + 1] - off[k] : u0_nbatch, bounds, uRelTol, uAbsTol, nSWEEPS, relax, at(J, 1, off[k]),
                                 at(sweeps, 1, off[k]), at(converged, 1, off[k]), at(max_change, 1, off[k]));
    }, &off);
    if (rc == 0) start_batch(m, Bt);
    return rc;
}

int ocs_multi_single_shooting(ocs_b200_handle *h, const double *x0, long Bt, const double *v0, long v0_nbatch, const double *Lb,
                              const double *Ub, int nfree, const int32_t *free_states, const double *free_lb,
                              const double *free_ub, int max_iter, double tol, double *v_opt, double *J, int32_t *iters,
                              int32_t *converged)
{
    if (!x0 || !v0 || Bt < 1 || (v0_nbatch != 1 && v0_nbatch != Bt) || nfree < 0)
        return h->fail(OCS_B200_ERR_ARG, "bad single_shooting arguments");
    M *m = h->multi;
    int n, nc, np, K, N, N2;
    dims(h, n, nc, np, K, N, N2);
    const size_t nv = (size_t)nc * K;
    m->Bt = 0;
    const std::vector<long> off = split(Bt, m->ndev());
    int rc = fan(h, [&](int k, ocs_b200_handle *c) {
        const bool per = v0_nbatch > 1;
        return ocs_b200_single_shooting(c, at(x0, n, off[k]), off[k + 1] - off[k], per ? at(v0, nv, off[k]) : v0,
                                        per ? off[k + 1] - off[k] : v0_nbatch, Lb, Ub, nfree, free_states, free_lb, free_ub, max_iter,
                                        tol, at(v_opt, nv + nfree, off[k]), at(J, 1, off[k]), at(iters, 1, off[k]),
                                        at(converged, 1, off[k]));
    }, &off);
    if (rc == 0) start_batch(m, Bt);
    return rc;
}

int ocs_multi_compute_equilibrium(ocs_b200_handle *h, const double *params, int np, long nparams, long nsets, const double *guess,
                                  long nguess, const double *lb, const double *ub, const double *r, long nr, int max_iter,
                                  double tol, double *y, double *resnorm, double *residual, int32_t *exitflag, int32_t *iters)
{
    if (nsets < 1 || !y) return h->fail(OCS_B200_ERR_ARG, "bad compute_equilibrium arguments");
    M *m = h->multi;
    int n, nc, npp, K, N, N2;
    dims(h, n, nc, npp, K, N, N2);
    const size_t NY = 2 * (size_t)n + nc;
    const std::vector<long> off = split(nsets, m->ndev());
    return fan(h, [&](int k, ocs_b200_handle *c) {
        const long cnt = off[k + 1] - off[k];
        return ocs_b200_compute_equilibrium(c, nparams > 1 ? at(params, np, off[k]) : params, np, nparams > 1 ? cnt : nparams, cnt,
                                            nguess > 1 ? at(guess, NY, off[k]) : guess, nguess > 1 ? cnt : nguess, lb, ub,
                                            nr > 1 ? at(r, 1, off[k]) : r, nr > 1 ? cnt : nr, max_iter, tol, at(y, NY, off[k]),
                                            at(resnorm, 1, off[k]), at(residual, NY, off[k]), at(exitflag, 1, off[k]),
                                            at(iters, 1, off[k]));
    }, &off);
}

int ocs_multi_reduce_J(ocs_b200_handle *h, int op, double *value, long *index)
{
    M *m = h->multi;
    if (m->Bt < 1) return h->fail(OCS_B200_ERR_STATE, "reduce_J needs a preceding compute_states");
    if (!value || op < 0 || op > 2) return h->fail(OCS_B200_ERR_ARG, "bad reduction");
    const int nd = m->ndev();
    std::vector<double> v(nd, 0.0);
    std::vector<long> ix(nd, -1);
    TRY(fan(h, [&](int k, ocs_b200_handle *c) { return ocs_b200_reduce_J(c, op, &v[k], &ix[k]); }, &m->off));
    // shards in ascending order: the same tie-breaking (lowest index) and a fixed summation order
    double best = 0.0;
    long bi = -1;
    bool first = true;
    for (int k = 0; k < nd; ++k) {
        if (m->off[k + 1] == m->off[k]) continue;
        const long gi = ix[k] >= 0 ? ix[k] + m->off[k] : -1;
        if (first) { best = v[k]; bi = gi; first = false; continue; }
        if (op == OCS_B200_REDUCE_SUM) best += v[k];
        else if ((op == OCS_B200_REDUCE_MIN && v[k] < best) || (op == OCS_B200_REDUCE_MAX && v[k] > best)) { best = v[k]; bi = gi; }
    }
    *value = best;
    if (index) *index = op == OCS_B200_REDUCE_SUM ? -1 : bi;
    return 0;
}

int ocs_multi_kernel_stats_ex(ocs_b200_handle *h, int reset, double *ms, int nms, long *launches)
{
    M *m = h->multi;
    long total = 0;
    for (int q = 0; q < nms; ++q) ms[q] = 0.0;
    for (ocs_b200_handle *c : m->dev) {        // device times overlap: the slowest device is the step's time
        double v[8] = {0};
        long l = 0;
        ocs_b200_kernel_stats_ex(c, reset, v, nms < 8 ? nms : 8, &l);
        for (int q = 0; q < nms && q < 8; ++q) ms[q] = std::max(ms[q], v[q]);
        total += l;
    }
    if (launches) *launches = total;
    return 0;
}
int ocs_multi_kernel_stats(ocs_b200_handle *h, int reset, double *ms_forward, double *ms_reverse, long *launches)
{
    double v[4] = {0};
    ocs_multi_kernel_stats_ex(h, reset, v, 4, launches);
    if (ms_forward) *ms_forward = v[0];
    if (ms_reverse) *ms_reverse = v[1] + v[2];
    return 0;
}

// =============================================================================== device-resident evaluation and the root gather

int ocs_multi_evaluate_resident(ocs_b200_handle *h, const double *x0, long Bt, int want_gradient)
{
    M *m = h->multi;
    int n, nc, np, K, N, N2;
    dims(h, n, nc, np, K, N, N2);
    if (x0) {
        if (Bt < 1) return h->fail(OCS_B200_ERR_ARG, "bad batch size");
        m->Bt = 0;
        const std::vector<long> off = split(Bt, m->ndev());
        int rc = fan(h, [&](int k, ocs_b200_handle *c) { return ocs_b200_evaluate_resident(c, at(x0, n, off[k]), off[k + 1] - off[k], want_gradient); }, &off);
        if (rc == 0) start_batch(m, Bt);
        return rc;
    }
    if (m->Bt < 1) return h->fail(OCS_B200_ERR_STATE, "no resident initial states: pass x0 once");
    return fan(h, [&](int, ocs_b200_handle *c) { return ocs_b200_evaluate_resident(c, nullptr, 0, want_gradient); }, &m->off);
}

int ocs_multi_get_results(ocs_b200_handle *h, double *J, double *xT, int32_t *status, double *dJdv, double *lam0)
{
    M *m = h->multi;
    if (m->Bt < 1) return h->fail(OCS_B200_ERR_STATE, "no resident results");
    int n, nc, np, K, N, N2;
    dims(h, n, nc, np, K, N, N2);
    const std::vector<long> &off = m->off;
    return fan(h, [&](int k, ocs_b200_handle *c) {
        return ocs_b200_get_results(c, at(J, 1, off[k]), at(xT, n, off[k]), at(status, 1, off[k]), at(dJdv, (size_t)nc * K, off[k]),
                                    at(lam0, n + 1, off[k]));
    }, &off);
}

// J [Bt] and x(T) [n][Bt] (SoA, leading dimension Bt) of the resident batch, gathered into buffers on the root device
// (devices[0]).  Ordered after each device's sweeps on its own stream; the root's stream waits for every piece, so
// work queued on the root afterwards (ocs_b200_set_stream / the root handle's stream) sees the gathered arrays.
int ocs_multi_gather_results_dev(ocs_b200_handle *h, double *J_root, double *xT_root)
{
    M *m = h->multi;
    if (m->Bt < 1) return h->fail(OCS_B200_ERR_STATE, "no resident results to gather");
    int n, nc, np, K, N, N2;
    dims(h, n, nc, np, K, N, N2);
    const int nd = m->ndev();
    const long Bt = m->Bt;
    if (m->gather_mode < 0) {
        m->gather_mode = 0;
#if OCS_HAVE_NCCL_H
        const char *e = getenv("OCS_B200_GATHER");
        if (nd > 1 && !(e && !strcmp(e, "p2p")) && m->nccl.load()) {
            m->comm.assign(nd, nullptr);
            if (m->nccl.CommInitAll(m->comm.data(), nd, m->device.data()) == ncclSuccess) m->gather_mode = 1;
            else m->comm.clear();
        }
#endif
    }
    ocs_b200_handle *root = m->dev[0];
#if OCS_HAVE_NCCL_H
    if (m->gather_mode == 1) {
        // gather to root = one grouped set of point-to-point transfers: every peer sends its slices, the root posts
        // the matching receives (rows of x(T) are separate slices of the root's [n][Bt] array)
        ncclResult_t r = m->nccl.GroupStart();
        for (int k = 1; k < nd && r == ncclSuccess; ++k) {
            const long cnt = m->off[k + 1] - m->off[k];
            if (cnt == 0) continue;
            ocs_b200_handle *c = m->dev[k];
            if (J_root) {
                r = m->nccl.Send(c->d_J.p, cnt, ncclDouble, 0, m->comm[k], c->stream);
                if (r == ncclSuccess) r = m->nccl.Recv(J_root + m->off[k], cnt, ncclDouble, k, m->comm[0], root->stream);
            }
            for (int row = 0; xT_root && row < n && r == ncclSuccess; ++row) {
                r = m->nccl.Send(c->d_xT.p + (size_t)row * cnt, cnt, ncclDouble, 0, m->comm[k], c->stream);
                if (r == ncclSuccess) r = m->nccl.Recv(xT_root + (size_t)row * Bt + m->off[k], cnt, ncclDouble, k, m->comm[0], root->stream);
            }
        }
        const ncclResult_t r2 = m->nccl.GroupEnd();
        if (r != ncclSuccess || r2 != ncclSuccess)
            return h->fail(OCS_B200_ERR_CUDA, "NCCL gather failed: %s", m->nccl.GetErrorString(r != ncclSuccess ? r : r2));
    }
#endif
    // the root's own shard, and with peer copies every shard: copy engines, each on the source device's stream
    for (int k = 0; k < nd; ++k) {
        const long cnt = m->off[k + 1] - m->off[k];
        if (cnt == 0 || (k > 0 && m->gather_mode == 1)) continue;
        ocs_b200_handle *c = m->dev[k];
        CK(cudaSetDevice(m->device[k]));
        if (J_root) CK(cudaMemcpyPeerAsync(J_root + m->off[k], m->device[0], c->d_J.p, m->device[k], sizeof(double) * cnt, c->stream));
        for (int row = 0; xT_root && row < n; ++row)
            CK(cudaMemcpyPeerAsync(xT_root + (size_t)row * Bt + m->off[k], m->device[0], c->d_xT.p + (size_t)row * cnt, m->device[k],
                                   sizeof(double) * cnt, c->stream));
        if (k > 0) {
            CK(cudaEventRecord(m->ev[k], c->stream));
            CK(cudaSetDevice(m->device[0]));
            CK(cudaStreamWaitEvent(root->stream, m->ev[k], 0));
        }
    }
    CK(cudaSetDevice(m->device[0]));
    return 0;
}
