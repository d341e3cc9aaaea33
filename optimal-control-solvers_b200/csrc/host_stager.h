// host_stager.h -- host <-> device copies for PAGEABLE caller buffers.
//
// The C ABI takes plain host pointers.  MATLAB's mxArrays (and NumPy arrays) are pageable memory:
// a cudaMemcpyAsync on them is staged by the driver through one small pinned buffer on the calling
// thread and reaches ~5 GB/s, a tenth of what the link carries from pinned memory (round 1 measured
// 8.1 ms pinned vs 89 ms pageable for the same 394 MB).  The stager owns a ring of pinned slots and a
// few copy threads: a device->host transfer is cut into slot-sized DMAs on the caller's stream, and
// as each DMA lands (one CUDA event per slot) a copy thread moves the slot into the caller's buffer
// while the next DMAs are in flight -- the PCIe transfer and the host memcpy overlap, and the memcpy
// (page faults of a fresh mxArray included) is spread over several cores.  Host->device is the mirror
// image: slot <- caller memory on the calling thread, then an asynchronous DMA out of the slot.
//
// Pinned (cudaHostAlloc / cudaHostRegister) and managed pointers bypass the ring: one cudaMemcpyAsync.
//
//   OCS_B200_STAGER=0            always plain cudaMemcpyAsync (A/B switch)
//   OCS_B200_STAGER_THREADS=k    copy threads (default: a quarter of the host's hardware threads, 2..8)
//   OCS_B200_STAGER_SLOT_MB=m    slot size (default 4 MiB; the ring holds 3 slots per copy thread)
#pragma once

#include <cuda_runtime.h>

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

extern "C" void ocs_host_copy(void *dst, const void *src, size_t n);   // host_copy.cpp: non-temporal stores for large blocks

namespace ocs_detail {

class HostStager {
public:
    HostStager() = default;
    HostStager(const HostStager &) = delete;
    HostStager &operator=(const HostStager &) = delete;
    ~HostStager() { shutdown(); }

    // true when `p` can be handed to cudaMemcpyAsync as is (pinned, registered or managed memory)
    static bool is_pinned(const void *p)
    {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
            cudaGetLastError();          // pre-11 runtimes report plain malloc memory as an error
            return false;
        }
        return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged || at.type == cudaMemoryTypeDevice;
    }

    static bool enabled()
    {
        const char *e = getenv("OCS_B200_STAGER");
        return !(e && e[0] == '0');
    }

    // transfers below this size go through the driver's own (synchronous) pageable path
    static constexpr size_t kMinBytes = 1u << 20;

    // dst (host, pageable) <- src (device), ordered on `stream`.  Returns once every DMA has been issued;
    // the data is in dst after finish().
    cudaError_t d2h(int device, void *dst, const void *src, size_t bytes, cudaStream_t stream)
    {
        cudaError_t e = start(device);
        if (e != cudaSuccess) return e;
        char *d = static_cast<char *>(dst);
        const char *s = static_cast<const char *>(src);
        for (size_t off = 0; off < bytes; off += slot_bytes_) {
            const size_t n = bytes - off < slot_bytes_ ? bytes - off : slot_bytes_;
            const int k = acquire();
            if (k < 0) return error_;
            e = cudaMemcpyAsync(slots_[k].p, s + off, n, cudaMemcpyDeviceToHost, stream);
            if (e == cudaSuccess) e = cudaEventRecord(slots_[k].ev, stream);
            if (e != cudaSuccess) {
                release(k);
                return e;
            }
            push(Job{k, d + off, n});
        }
        return cudaSuccess;
    }

    // dst (device) <- src (host, pageable), ordered on `stream`.  Returns once src has been read completely.
    cudaError_t h2d(int device, void *dst, const void *src, size_t bytes, cudaStream_t stream)
    {
        cudaError_t e = start(device);
        if (e != cudaSuccess) return e;
        char *d = static_cast<char *>(dst);
        const char *s = static_cast<const char *>(src);
        for (size_t off = 0; off < bytes; off += slot_bytes_) {
            const size_t n = bytes - off < slot_bytes_ ? bytes - off : slot_bytes_;
            const int k = acquire();
            if (k < 0) return error_;
            ocs_host_copy(slots_[k].p, s + off, n);
            e = cudaMemcpyAsync(d + off, slots_[k].p, n, cudaMemcpyHostToDevice, stream);
            if (e == cudaSuccess) e = cudaEventRecord(slots_[k].ev, stream);
            if (e != cudaSuccess) {
                release(k);
                return e;
            }
            push(Job{k, nullptr, 0});      // a copy thread frees the slot when the DMA has read it
        }
        return cudaSuccess;
    }

    // every transfer handed to d2h()/h2d() so far has completed (host side included)
    cudaError_t finish()
    {
        if (!started_) return cudaSuccess;
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return pending_ == 0; });
        const cudaError_t e = error_;
        error_ = cudaSuccess;
        return e;
    }

    void shutdown()
    {
        if (!started_) return;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_done_.wait(lk, [&] { return pending_ == 0; });
            stop_ = true;
        }
        cv_job_.notify_all();
        for (auto &t : threads_) t.join();
        threads_.clear();
        cudaSetDevice(device_);
        for (auto &s : slots_) {
            if (s.ev) cudaEventDestroy(s.ev);
            if (s.p) cudaFreeHost(s.p);
        }
        slots_.clear();
        started_ = false;
        stop_ = false;
    }

    int threads() const { return (int)threads_.size(); }

private:
    struct Slot {
        char *p = nullptr;
        cudaEvent_t ev = nullptr;
        bool busy = false;
    };
    struct Job {
        int slot;
        char *dst;      // null: host->device, only the slot is released
        size_t n;
    };

    cudaError_t start(int device)
    {
        if (started_) return cudaSuccess;
        device_ = device;
        unsigned hw = std::thread::hardware_concurrency();
        int nt = (int)(hw / 4);
        if (nt < 2) nt = 2;
        if (nt > 8) nt = 8;
        if (const char *e = getenv("OCS_B200_STAGER_THREADS")) nt = atoi(e) > 0 ? atoi(e) : nt;
        size_t mb = 4;
        if (const char *e = getenv("OCS_B200_STAGER_SLOT_MB")) mb = atoi(e) > 0 ? (size_t)atoi(e) : mb;
        slot_bytes_ = mb << 20;
        slots_.resize((size_t)3 * nt);
        for (auto &s : slots_) {
            cudaError_t e = cudaHostAlloc((void **)&s.p, slot_bytes_, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming);
            if (e != cudaSuccess) {
                for (auto &q : slots_) {
                    if (q.ev) cudaEventDestroy(q.ev);
                    if (q.p) cudaFreeHost(q.p);
                }
                slots_.clear();
                return e;
            }
        }
        started_ = true;
        for (int k = 0; k < nt; ++k) threads_.emplace_back([this] { run(); });
        return cudaSuccess;
    }

    int acquire()
    {
        std::unique_lock<std::mutex> lk(mu_);
        int found = -1;
        cv_slot_.wait(lk, [&] {
            if (error_ != cudaSuccess) return true;
            for (size_t k = 0; k < slots_.size(); ++k)
                if (!slots_[k].busy) {
                    found = (int)k;
                    return true;
                }
            return false;
        });
        if (error_ != cudaSuccess) return -1;
        slots_[found].busy = true;
        return found;
    }

    void release(int k)
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            slots_[k].busy = false;
        }
        cv_slot_.notify_one();
    }

    void push(const Job &j)
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            jobs_.push_back(j);
            ++pending_;
        }
        cv_job_.notify_one();
    }

    void run()
    {
        cudaSetDevice(device_);
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_job_.wait(lk, [&] { return stop_ || !jobs_.empty(); });
                if (jobs_.empty()) return;       // stop_ and nothing left
                j = jobs_.front();
                jobs_.pop_front();
            }
            const cudaError_t e = cudaEventSynchronize(slots_[j.slot].ev);
            if (e == cudaSuccess && j.dst) ocs_host_copy(j.dst, slots_[j.slot].p, j.n);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (e != cudaSuccess && error_ == cudaSuccess) error_ = e;
                slots_[j.slot].busy = false;
                --pending_;
            }
            cv_slot_.notify_one();
            cv_done_.notify_all();
        }
    }

    int device_ = 0;
    bool started_ = false, stop_ = false;
    size_t slot_bytes_ = 4u << 20;
    std::vector<Slot> slots_;
    std::vector<std::thread> threads_;
    std::deque<Job> jobs_;
    long pending_ = 0;
    cudaError_t error_ = cudaSuccess;
    std::mutex mu_;
    std::condition_variable cv_job_, cv_slot_, cv_done_;
};

}  // namespace ocs_detail
