// Shared-memory collectives for the host tier (CPU inference all-reduce, gloo-free small messages).
//
// Role parity: reference csrc/cpu/comm/{shm.cpp,shm_interface.cpp} (N17: `inference_all_reduce`).
// Independent design: one POSIX shm segment per communicator holding, per rank, a data slot plus a
// sense-reversing barrier built from two 64-bit atomics.  all_reduce = (1) every rank copies its
// input into its slot, (2) barrier, (3) rank r reduces the r-th 1/world slice across all slots
// (fp32 accumulate; bf16/fp16 inputs widened on the fly) and writes it to the shared result area,
// (4) barrier, (5) every rank copies the result out.  Small messages (< 32 KiB) skip the slicing
// and have every rank reduce everything (one barrier less).
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <new>

#define DSB_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

enum { kF32 = 0, kF16 = 1, kBF16 = 2 };

struct Header {
    std::atomic<uint64_t> arrived;
    std::atomic<uint64_t> generation;
    uint64_t world;
    uint64_t slot_bytes;
    char pad[64 - 4 * sizeof(uint64_t)];
};

struct Comm {
    int rank, world;
    size_t slot_bytes, total_bytes;
    char* base;
    Header* hdr;
    char name[128];
    bool owner;
    char* slot(int r) const { return base + sizeof(Header) + static_cast<size_t>(r) * slot_bytes; }
    char* result() const { return slot(world); }
};

void barrier(Comm* c)
{
    const uint64_t gen = c->hdr->generation.load(std::memory_order_acquire);
    if (c->hdr->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == static_cast<uint64_t>(c->world)) {
        c->hdr->arrived.store(0, std::memory_order_relaxed);
        c->hdr->generation.store(gen + 1, std::memory_order_release);
    } else {
        int spins = 0;
        while (c->hdr->generation.load(std::memory_order_acquire) == gen) {
            if (++spins > 1024) {
                usleep(1);
                spins = 0;
            }
        }
    }
}

inline float widen(const void* p, int dt, size_t i)
{
    if (dt == kF32) return static_cast<const float*>(p)[i];
    const uint16_t h = static_cast<const uint16_t*>(p)[i];
    if (dt == kBF16) {
        uint32_t u = static_cast<uint32_t>(h) << 16;
        float f;
        memcpy(&f, &u, 4);
        return f;
    }
    return static_cast<float>(*reinterpret_cast<const _Float16*>(&h));
}

inline void narrow(void* p, int dt, size_t i, float v)
{
    if (dt == kF32) {
        static_cast<float*>(p)[i] = v;
    } else if (dt == kBF16) {
        uint32_t u;
        memcpy(&u, &v, 4);
        u += 0x7fffu + ((u >> 16) & 1u);
        static_cast<uint16_t*>(p)[i] = static_cast<uint16_t>(u >> 16);
    } else {
        _Float16 h = static_cast<_Float16>(v);
        memcpy(static_cast<uint16_t*>(p) + i, &h, 2);
    }
}

}  // namespace

DSB_EXPORT void* dsb_shm_create(const char* name, int rank, int world, int64_t max_bytes)
{
    Comm* c = new (std::nothrow) Comm();
    if (!c) return nullptr;
    c->rank = rank;
    c->world = world;
    c->slot_bytes = (static_cast<size_t>(max_bytes) + 63) & ~size_t(63);
    c->total_bytes = sizeof(Header) + c->slot_bytes * static_cast<size_t>(world + 1);
    snprintf(c->name, sizeof(c->name), "/dsb200_%s", name);
    c->owner = rank == 0;
    int fd = -1;
    if (rank == 0) {
        shm_unlink(c->name);
        fd = shm_open(c->name, O_CREAT | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, static_cast<off_t>(c->total_bytes)) != 0) {
            delete c;
            return nullptr;
        }
    } else {
        for (int tries = 0; tries < 20000 && fd < 0; ++tries) {
            fd = shm_open(c->name, O_RDWR, 0600);
            if (fd >= 0) {
                struct stat st;
                if (fstat(fd, &st) != 0 || static_cast<size_t>(st.st_size) < c->total_bytes) {
                    close(fd);
                    fd = -1;
                }
            }
            if (fd < 0) usleep(500);
        }
        if (fd < 0) {
            delete c;
            return nullptr;
        }
    }
    void* p = mmap(nullptr, c->total_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) {
        delete c;
        return nullptr;
    }
    c->base = static_cast<char*>(p);
    c->hdr = reinterpret_cast<Header*>(p);
    if (rank == 0) {
        c->hdr->arrived.store(0);
        c->hdr->generation.store(0);
        c->hdr->slot_bytes = c->slot_bytes;
        c->hdr->world = static_cast<uint64_t>(world);  // written last: peers poll on it
    } else {
        while (reinterpret_cast<volatile uint64_t&>(c->hdr->world) != static_cast<uint64_t>(world)) usleep(100);
    }
    barrier(c);
    return c;
}

DSB_EXPORT void dsb_shm_destroy(void* h)
{
    Comm* c = static_cast<Comm*>(h);
    if (!c) return;
    barrier(c);
    munmap(c->base, c->total_bytes);
    if (c->owner) shm_unlink(c->name);
    delete c;
}

DSB_EXPORT int dsb_shm_barrier(void* h)
{
    barrier(static_cast<Comm*>(h));
    return 0;
}

// In-place sum all-reduce of `n` elements of dtype `dt`.
DSB_EXPORT int dsb_shm_all_reduce(void* h, void* data, int64_t n, int dt)
{
    Comm* c = static_cast<Comm*>(h);
    const size_t esz = dt == kF32 ? 4 : 2;
    const size_t bytes = static_cast<size_t>(n) * esz;
    if (bytes > c->slot_bytes) return -EMSGSIZE;
    memcpy(c->slot(c->rank), data, bytes);
    barrier(c);
    if (bytes < 32768) {
        for (size_t i = 0; i < static_cast<size_t>(n); ++i) {
            float acc = 0.f;
            for (int r = 0; r < c->world; ++r) acc += widen(c->slot(r), dt, i);
            narrow(data, dt, i, acc);
        }
        barrier(c);  // nobody may overwrite its slot before all have read
        return 0;
    }
    const size_t per = (static_cast<size_t>(n) + c->world - 1) / c->world;
    const size_t lo = per * c->rank, hi = (lo + per < static_cast<size_t>(n)) ? lo + per : static_cast<size_t>(n);
    for (size_t i = lo; i < hi; ++i) {
        float acc = 0.f;
        for (int r = 0; r < c->world; ++r) acc += widen(c->slot(r), dt, i);
        narrow(c->result(), dt, i, acc);
    }
    barrier(c);
    memcpy(data, c->result(), bytes);
    barrier(c);
    return 0;
}

// all_gather: out[r*n:(r+1)*n] = rank r's data.
DSB_EXPORT int dsb_shm_all_gather(void* h, const void* data, void* out, int64_t bytes)
{
    Comm* c = static_cast<Comm*>(h);
    if (static_cast<size_t>(bytes) > c->slot_bytes) return -EMSGSIZE;
    memcpy(c->slot(c->rank), data, static_cast<size_t>(bytes));
    barrier(c);
    for (int r = 0; r < c->world; ++r) memcpy(static_cast<char*>(out) + r * bytes, c->slot(r), static_cast<size_t>(bytes));
    barrier(c);
    return 0;
}
