// Asynchronous file I/O engine for the NVMe offload tier (ZeRO-Infinity swapping, DeepNVMe).
//
// Role parity: reference csrc/aio/common/* + csrc/aio/py_lib/* (N10: `aio_handle` with
// block_size / queue_depth / single_submit / overlap_events / intra_op_parallelism, sync and async
// pread/pwrite, wait()).  Implementation is independent: Linux native AIO through raw syscalls
// (io_setup/io_submit/io_getevents from <linux/aio_abi.h>; no libaio dependency) with O_DIRECT
// when buffer, offset and length are 512-byte aligned, and a buffered pread/pwrite fallback
// otherwise.  Each handle owns `intra_op_parallelism` worker threads; one request is split into
// per-thread contiguous spans, each span streamed as `queue_depth` in-flight `block_size` iocbs.
#include <errno.h>
#include <fcntl.h>
#include <linux/aio_abi.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#define DSB_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

inline int sys_io_setup(unsigned nr, aio_context_t* ctx) { return static_cast<int>(syscall(__NR_io_setup, nr, ctx)); }
inline int sys_io_destroy(aio_context_t ctx) { return static_cast<int>(syscall(__NR_io_destroy, ctx)); }
inline int sys_io_submit(aio_context_t ctx, long n, iocb** cbs)
{
    return static_cast<int>(syscall(__NR_io_submit, ctx, n, cbs));
}
inline int sys_io_getevents(aio_context_t ctx, long min_nr, long nr, io_event* ev)
{
    return static_cast<int>(syscall(__NR_io_getevents, ctx, min_nr, nr, ev, nullptr));
}

struct Span {
    int fd;
    bool write;
    char* buf;
    int64_t offset;  // file offset
    int64_t bytes;
    bool direct;
};

struct Request {
    std::atomic<int> remaining{0};
    std::atomic<int64_t> error{0};
    int fd = -1;
    int fd_direct = -1;
};

struct Task {
    Span span;
    Request* req;
};

class Handle {
public:
    Handle(int64_t block_size, int queue_depth, bool single_submit, bool overlap_events, int threads)
        : block_size_(block_size > 0 ? block_size : (1 << 20)),
          queue_depth_(queue_depth > 0 ? queue_depth : 8),
          single_submit_(single_submit),
          overlap_events_(overlap_events),
          nthreads_(threads > 0 ? threads : 1)
    {
        for (int i = 0; i < nthreads_; ++i) workers_.emplace_back([this] { this->worker(); });
    }

    ~Handle()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
        for (auto* r : inflight_) finish(r);
    }

    // Returns 0 on success (async: queued), negative errno otherwise.
    int64_t submit(const char* path, void* buffer, int64_t bytes, int64_t file_offset, bool write, bool async)
    {
        int flags = write ? (O_WRONLY | O_CREAT) : O_RDONLY;
        int fd = open(path, flags, 0644);
        if (fd < 0) return -errno;
        int fd_direct = -1;
        const bool aligned = ((reinterpret_cast<uintptr_t>(buffer) | static_cast<uint64_t>(file_offset) |
                               static_cast<uint64_t>(bytes)) & 511) == 0;
        if (aligned) fd_direct = open(path, flags | O_DIRECT, 0644);
        auto* req = new Request();
        req->fd = fd;
        req->fd_direct = fd_direct;
        // split across worker threads on block boundaries
        int64_t nblocks = (bytes + block_size_ - 1) / block_size_;
        int parts = static_cast<int>(nblocks < nthreads_ ? (nblocks > 0 ? nblocks : 1) : nthreads_);
        int64_t per = ((nblocks + parts - 1) / parts) * block_size_;
        std::vector<Task> tasks;
        for (int i = 0; i < parts; ++i) {
            int64_t lo = i * per, hi = lo + per < bytes ? lo + per : bytes;
            if (lo >= hi) break;
            Task t;
            t.span = Span{fd_direct >= 0 ? fd_direct : fd, write, static_cast<char*>(buffer) + lo, file_offset + lo,
                          hi - lo, fd_direct >= 0};
            t.req = req;
            tasks.push_back(t);
        }
        req->remaining.store(static_cast<int>(tasks.size()));
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (auto& t : tasks) queue_.push_back(t);
            inflight_.push_back(req);
        }
        cv_.notify_all();
        if (!async) return wait_one(req);
        return 0;
    }

    // Wait for every outstanding request; returns number completed or negative errno.
    int64_t wait_all()
    {
        std::vector<Request*> reqs;
        {
            std::lock_guard<std::mutex> lk(mu_);
            reqs.swap(inflight_);
        }
        int64_t err = 0;
        for (auto* r : reqs) {
            spin_until_done(r);
            int64_t e = finish(r);
            if (e < 0 && err == 0) err = e;
        }
        return err < 0 ? err : static_cast<int64_t>(reqs.size());
    }

    int64_t block_size() const { return block_size_; }
    int queue_depth() const { return queue_depth_; }
    bool single_submit() const { return single_submit_; }
    bool overlap_events() const { return overlap_events_; }
    int threads() const { return nthreads_; }

private:
    void spin_until_done(Request* r)
    {
        std::unique_lock<std::mutex> lk(done_mu_);
        done_cv_.wait(lk, [r] { return r->remaining.load() == 0; });
    }

    int64_t finish(Request* r)
    {
        int64_t e = r->error.load();
        if (r->fd >= 0) close(r->fd);
        if (r->fd_direct >= 0) close(r->fd_direct);
        delete r;
        return e;
    }

    int64_t wait_one(Request* req)
    {
        spin_until_done(req);
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (size_t i = 0; i < inflight_.size(); ++i)
                if (inflight_[i] == req) {
                    inflight_.erase(inflight_.begin() + i);
                    break;
                }
        }
        return finish(req);
    }

    void worker()
    {
        aio_context_t ctx = 0;
        bool have_ctx = sys_io_setup(static_cast<unsigned>(queue_depth_), &ctx) == 0;
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return stop_ || !queue_.empty(); });
                if (stop_ && queue_.empty()) break;
                t = queue_.front();
                queue_.pop_front();
            }
            int64_t err = 0;
            if (have_ctx && t.span.direct)
                err = run_kernel_aio(ctx, t.span);
            else
                err = run_buffered(t.span);
            if (err < 0) t.req->error.store(err);
            if (t.req->remaining.fetch_sub(1) == 1) {
                std::lock_guard<std::mutex> lk(done_mu_);
                done_cv_.notify_all();
            }
        }
        if (have_ctx) sys_io_destroy(ctx);
    }

    int64_t run_buffered(const Span& s)
    {
        int64_t done = 0;
        while (done < s.bytes) {
            int64_t chunk = s.bytes - done < block_size_ ? s.bytes - done : block_size_;
            ssize_t r = s.write ? pwrite(s.fd, s.buf + done, chunk, s.offset + done)
                                : pread(s.fd, s.buf + done, chunk, s.offset + done);
            if (r < 0) {
                if (errno == EINTR) continue;
                return -errno;
            }
            if (r == 0) break;  // short file
            done += r;
        }
        return done;
    }

    int64_t run_kernel_aio(aio_context_t ctx, const Span& s)
    {
        const int64_t nblocks = (s.bytes + block_size_ - 1) / block_size_;
        std::vector<iocb> cbs(static_cast<size_t>(queue_depth_));
        std::vector<iocb*> ptrs(static_cast<size_t>(queue_depth_));
        std::vector<io_event> events(static_cast<size_t>(queue_depth_));
        int64_t next = 0, completed = 0, inflight = 0;
        std::vector<int> free_slots;
        for (int i = queue_depth_ - 1; i >= 0; --i) free_slots.push_back(i);
        while (completed < nblocks) {
            int nsub = 0;
            while (next < nblocks && !free_slots.empty()) {
                const int slot = free_slots.back();
                free_slots.pop_back();
                iocb& cb = cbs[slot];
                memset(&cb, 0, sizeof(cb));
                const int64_t off = next * block_size_;
                const int64_t len = s.bytes - off < block_size_ ? s.bytes - off : block_size_;
                cb.aio_fildes = static_cast<uint32_t>(s.fd);
                cb.aio_lio_opcode = s.write ? IOCB_CMD_PWRITE : IOCB_CMD_PREAD;
                cb.aio_buf = reinterpret_cast<uint64_t>(s.buf + off);
                cb.aio_nbytes = static_cast<uint64_t>(len);
                cb.aio_offset = s.offset + off;
                cb.aio_data = static_cast<uint64_t>(slot);
                ptrs[nsub++] = &cb;
                ++next;
                if (single_submit_) {
                    int r = sys_io_submit(ctx, 1, &ptrs[nsub - 1]);
                    if (r < 0) return -errno;
                    ++inflight;
                    nsub = 0;
                }
            }
            if (nsub > 0) {
                int sent = 0;
                while (sent < nsub) {
                    int r = sys_io_submit(ctx, nsub - sent, ptrs.data() + sent);
                    if (r < 0) {
                        if (errno == EAGAIN || errno == EINTR) continue;
                        return -errno;
                    }
                    sent += r;
                }
                inflight += nsub;
            }
            const long min_nr = overlap_events_ ? 1 : inflight;
            int got = sys_io_getevents(ctx, min_nr, static_cast<long>(inflight), events.data());
            if (got < 0) {
                if (errno == EINTR) continue;
                return -errno;
            }
            for (int i = 0; i < got; ++i) {
                if (static_cast<int64_t>(events[i].res) < 0) return static_cast<int64_t>(events[i].res);
                free_slots.push_back(static_cast<int>(events[i].data));
            }
            completed += got;
            inflight -= got;
        }
        return s.bytes;
    }

    const int64_t block_size_;
    const int queue_depth_;
    const bool single_submit_;
    const bool overlap_events_;
    const int nthreads_;
    std::vector<std::thread> workers_;
    std::deque<Task> queue_;
    std::vector<Request*> inflight_;
    std::mutex mu_, done_mu_;
    std::condition_variable cv_, done_cv_;
    bool stop_ = false;
};

}  // namespace

DSB_EXPORT void* dsb_aio_create(int64_t block_size, int queue_depth, int single_submit, int overlap_events,
                                int threads)
{
    return new Handle(block_size, queue_depth, single_submit != 0, overlap_events != 0, threads);
}

DSB_EXPORT void dsb_aio_destroy(void* h) { delete static_cast<Handle*>(h); }

DSB_EXPORT int64_t dsb_aio_pread(void* h, void* buffer, int64_t bytes, const char* path, int64_t file_offset,
                                 int async)
{
    return static_cast<Handle*>(h)->submit(path, buffer, bytes, file_offset, false, async != 0);
}

DSB_EXPORT int64_t dsb_aio_pwrite(void* h, const void* buffer, int64_t bytes, const char* path, int64_t file_offset,
                                  int async)
{
    return static_cast<Handle*>(h)->submit(path, const_cast<void*>(buffer), bytes, file_offset, true, async != 0);
}

DSB_EXPORT int64_t dsb_aio_wait(void* h) { return static_cast<Handle*>(h)->wait_all(); }

// Page-aligned, mlock'ed host buffers (the reference's new_cpu_locked_tensor).
DSB_EXPORT void* dsb_aio_alloc_locked(int64_t bytes)
{
    void* p = nullptr;
    const int64_t rounded = (bytes + 4095) & ~int64_t(4095);
    if (posix_memalign(&p, 4096, static_cast<size_t>(rounded)) != 0) return nullptr;
    mlock(p, static_cast<size_t>(rounded));  // best effort
    return p;
}

DSB_EXPORT void dsb_aio_free_locked(void* p, int64_t bytes)
{
    if (!p) return;
    munlock(p, static_cast<size_t>((bytes + 4095) & ~int64_t(4095)));
    free(p);
}

DSB_EXPORT int64_t dsb_file_size(const char* path)
{
    struct stat st;
    if (stat(path, &st) != 0) return -errno;
    return static_cast<int64_t>(st.st_size);
}

// Multi-threaded memcpy (the reference's deepspeed_memcpy).
DSB_EXPORT void dsb_parallel_memcpy(void* dst, const void* src, int64_t bytes, int threads)
{
    if (threads <= 1 || bytes < (1 << 22)) {
        memcpy(dst, src, static_cast<size_t>(bytes));
        return;
    }
    std::vector<std::thread> ts;
    const int64_t per = ((bytes / threads) + 63) & ~int64_t(63);
    for (int i = 0; i < threads; ++i) {
        const int64_t lo = i * per, hi = (i == threads - 1) ? bytes : (lo + per < bytes ? lo + per : bytes);
        if (lo >= hi) break;
        ts.emplace_back([=] { memcpy(static_cast<char*>(dst) + lo, static_cast<const char*>(src) + lo, hi - lo); });
    }
    for (auto& t : ts) t.join();
}
