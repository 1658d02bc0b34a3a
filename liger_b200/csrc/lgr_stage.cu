// liger_b200 -- copies between CALLER host memory and the device.
//
// The reference's caller is R: the dgCMatrix slots and the factor matrices an Rcpp shim hands over live in R's heap,
// i.e. in PAGEABLE memory (SURVEY 8b "Ownership").  A cudaMemcpyAsync from pageable memory is staged by the driver through
// one bounce buffer on the calling thread and reaches a fraction of the PCIe rate.  Here pageable buffers go through a
// ring of page-locked bounce buffers that a few host threads fill (H2D) or drain (D2H) while the DMA engine works on the
// previous chunks; page-locked caller buffers (lgr_host_alloc, cudaHostRegister, torch pinned tensors, or the caller's
// promise LGR_FLAG_PINNED_IO) are copied from directly.
#include <string.h>

#include <algorithm>
#include <exception>
#include <mutex>
#include <thread>

#include "lgr_common.cuh"

namespace lgr {

bool host_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

namespace {
constexpr size_t CHUNK = (size_t)4 << 20;
constexpr int NTHREADS = 4;
constexpr int NBUF = 2 * NTHREADS;   // two bounce buffers per thread

struct BouncePool {
    std::mutex mu;   // one staged copy at a time per process
    void* buf[NBUF] = {};
    bool ready = false;
    void ensure() {
        if (ready) return;
        for (int b = 0; b < NBUF; ++b) LGR_CUDA(cudaHostAlloc(&buf[b], CHUNK, cudaHostAllocPortable));
        ready = true;
    }
};
BouncePool& pool() {
    static BouncePool p;   // lives until process exit (32 MB of page-locked memory)
    return p;
}

struct Events {
    cudaEvent_t ev[NBUF] = {};
    Events() {
        for (int b = 0; b < NBUF; ++b) LGR_CUDA(cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming));
    }
    ~Events() {
        for (int b = 0; b < NBUF; ++b)
            if (ev[b]) cudaEventDestroy(ev[b]);
    }
};

template <typename Body>
void run_threads(int nthreads, Body body) {
    std::exception_ptr err;
    std::mutex emu;
    std::vector<std::thread> th;
    auto wrapped = [&](int t) {
        try {
            body(t);
        } catch (...) {
            std::lock_guard<std::mutex> g(emu);
            if (!err) err = std::current_exception();
        }
    };
    for (int t = 1; t < nthreads; ++t) th.emplace_back(wrapped, t);
    wrapped(0);
    for (auto& x : th) x.join();
    if (err) std::rethrow_exception(err);
}
}  // namespace

// Host -> device.  Returns after the last byte has left the caller's buffer (pageable) or after enqueueing (pinned).
void copy_in(void* dst, const void* src, size_t bytes, cudaStream_t cs, bool pinned_promise) {
    if (!bytes) return;
    if (pinned_promise || bytes < (CHUNK >> 2) || host_is_pinned(src)) {
        LGR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, cs));
        return;
    }
    BouncePool& P = pool();
    std::lock_guard<std::mutex> lock(P.mu);
    P.ensure();
    Events E;
    int dev = 0;
    LGR_CUDA(cudaGetDevice(&dev));
    const size_t nchunks = (bytes + CHUNK - 1) / CHUNK;
    const int T = (int)std::min<size_t>(NTHREADS, nchunks);
    run_threads(T, [&](int t) {
        LGR_CUDA(cudaSetDevice(dev));
        bool used[2] = {false, false};
        int turn = 0;
        for (size_t c = (size_t)t; c < nchunks; c += (size_t)T, turn ^= 1) {
            const int b = 2 * t + turn;
            if (used[turn]) LGR_CUDA(cudaEventSynchronize(E.ev[b]));   // the DMA of this buffer's previous chunk is done
            const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
            memcpy(P.buf[b], static_cast<const char*>(src) + off, len);
            LGR_CUDA(cudaMemcpyAsync(static_cast<char*>(dst) + off, P.buf[b], len, cudaMemcpyHostToDevice, cs));
            LGR_CUDA(cudaEventRecord(E.ev[b], cs));
            used[turn] = true;
        }
        for (int u = 0; u < 2; ++u)
            if (used[u]) LGR_CUDA(cudaEventSynchronize(E.ev[2 * t + u]));
    });
}

// Device -> host (contiguous).  The source must be complete in stream order on `st`.  Returns when dst is filled
// (pageable) or after enqueueing (pinned: the caller synchronises the stream).
void copy_out(void* dst, const void* src, size_t bytes, cudaStream_t st, bool pinned_promise) {
    if (!bytes) return;
    if (pinned_promise || bytes < (CHUNK >> 2) || host_is_pinned(dst)) {
        LGR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
        return;
    }
    BouncePool& P = pool();
    std::lock_guard<std::mutex> lock(P.mu);
    P.ensure();
    Events E;
    int dev = 0;
    LGR_CUDA(cudaGetDevice(&dev));
    const size_t nchunks = (bytes + CHUNK - 1) / CHUNK;
    const int T = (int)std::min<size_t>(NTHREADS, nchunks);
    run_threads(T, [&](int t) {
        LGR_CUDA(cudaSetDevice(dev));
        auto issue = [&](size_t c, int b) {
            const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
            LGR_CUDA(cudaMemcpyAsync(P.buf[b], static_cast<const char*>(src) + off, len, cudaMemcpyDeviceToHost, st));
            LGR_CUDA(cudaEventRecord(E.ev[b], st));
        };
        // two chunks in flight per thread: while one is drained into the caller's memory the DMA fills the other
        size_t c = (size_t)t;
        if (c < nchunks) issue(c, 2 * t);
        if (c + T < nchunks) issue(c + T, 2 * t + 1);
        int turn = 0;
        for (; c < nchunks; c += (size_t)T, turn ^= 1) {
            const int b = 2 * t + turn;
            LGR_CUDA(cudaEventSynchronize(E.ev[b]));
            const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
            memcpy(static_cast<char*>(dst) + off, P.buf[b], len);
            if (c + 2 * (size_t)T < nchunks) issue(c + 2 * (size_t)T, b);
        }
    });
}

}  // namespace lgr
