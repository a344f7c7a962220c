// Symmetric (peer-mapped) device memory for single-node NVSwitch boxes.
//
// Each rank allocates physical memory with the CUDA VMM API (cuMemCreate), exports it as a POSIX
// file descriptor, ships the descriptor to every other rank of the node over unix-domain sockets
// (SCM_RIGHTS) and maps all peers' allocations into its own address space, so device code can
// load/store any rank's buffer over NVLink.  When the driver supports NVLS a multicast object is
// created on rank 0, shared the same way, bound to every rank's allocation and mapped, which gives
// one "multicast" address whose `multimem.ld_reduce` / `multimem.st` are executed by the switch.
//
// The reference has no equivalent (all its collectives are torch.distributed calls,
// deepspeed/comm/torch.py:96); SURVEY.md 5.8 item 2 describes this layer.
//
// Driver entry points are resolved at run time through cudaGetDriverEntryPoint, so the library
// links only against libcudart and loads on machines without a GPU.
#include <cuda.h>
#include <cuda_runtime_api.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/syscall.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#include <thread>
#include <vector>

#define DSB_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

template <typename F>
bool load_sym(const char* name, F* out)
{
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &st) != cudaSuccess || fn == nullptr ||
        st != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return false;
    }
    *out = reinterpret_cast<F>(fn);
    return true;
}

struct Driver {
    bool ok = false;
    bool mc_ok = false;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
    CUresult (*MemRelease)(CUmemGenericAllocationHandle);
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
    CUresult (*MemAddressFree)(CUdeviceptr, size_t);
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
    CUresult (*MemUnmap)(CUdeviceptr, size_t);
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
    CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType,
                                           unsigned long long);
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
    CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
    CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice);
    CUresult (*DeviceGet)(CUdevice*, int);
    CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
    CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice);
    CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                                 unsigned long long);
    CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
    CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t);
};

Driver& drv()
{
    static Driver d;
    static bool init = false;
    if (init) return d;
    init = true;
    cudaFree(nullptr);  // make sure a primary context exists
    bool ok = load_sym("cuMemCreate", &d.MemCreate) && load_sym("cuMemRelease", &d.MemRelease) &&
              load_sym("cuMemAddressReserve", &d.MemAddressReserve) &&
              load_sym("cuMemAddressFree", &d.MemAddressFree) && load_sym("cuMemMap", &d.MemMap) &&
              load_sym("cuMemUnmap", &d.MemUnmap) && load_sym("cuMemSetAccess", &d.MemSetAccess) &&
              load_sym("cuMemExportToShareableHandle", &d.MemExportToShareableHandle) &&
              load_sym("cuMemImportFromShareableHandle", &d.MemImportFromShareableHandle) &&
              load_sym("cuMemGetAllocationGranularity", &d.MemGetAllocationGranularity) &&
              load_sym("cuDeviceGetAttribute", &d.DeviceGetAttribute) && load_sym("cuDeviceGet", &d.DeviceGet);
    d.ok = ok;
    d.mc_ok = ok && load_sym("cuMulticastCreate", &d.MulticastCreate) &&
              load_sym("cuMulticastAddDevice", &d.MulticastAddDevice) &&
              load_sym("cuMulticastBindMem", &d.MulticastBindMem) &&
              load_sym("cuMulticastGetGranularity", &d.MulticastGetGranularity) &&
              load_sym("cuMulticastUnbind", &d.MulticastUnbind);
    return d;
}

CUmemAllocationProp alloc_prop(int device)
{
    CUmemAllocationProp p;
    memset(&p, 0, sizeof(p));
    p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    p.location.id = device;
    p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return p;
}

int map_rw(Driver& d, int device, CUmemGenericAllocationHandle h, size_t size, CUdeviceptr* out)
{
    CUdeviceptr va = 0;
    CUresult r = d.MemAddressReserve(&va, size, 0, 0, 0);
    if (r != CUDA_SUCCESS) return -static_cast<int>(r);
    r = d.MemMap(va, size, 0, h, 0);
    if (r != CUDA_SUCCESS) {
        d.MemAddressFree(va, size);
        return -static_cast<int>(r);
    }
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    r = d.MemSetAccess(va, size, &acc, 1);
    if (r != CUDA_SUCCESS) {
        d.MemUnmap(va, size);
        d.MemAddressFree(va, size);
        return -static_cast<int>(r);
    }
    *out = va;
    return 0;
}

// ---- unix-domain descriptor passing ------------------------------------------------------------------
void make_addr(const char* prefix, int rank, sockaddr_un* addr, socklen_t* len)
{
    memset(addr, 0, sizeof(*addr));
    addr->sun_family = AF_UNIX;
    // abstract namespace: first byte NUL, no filesystem entry to clean up
    int n = snprintf(addr->sun_path + 1, sizeof(addr->sun_path) - 2, "%s-%d", prefix, rank);
    *len = static_cast<socklen_t>(offsetof(sockaddr_un, sun_path) + 1 + n);
}

int send_fd(int sock, int fd, int tag)
{
    msghdr msg;
    memset(&msg, 0, sizeof(msg));
    char ctrl[CMSG_SPACE(sizeof(int))];
    memset(ctrl, 0, sizeof(ctrl));
    iovec io;
    io.iov_base = &tag;
    io.iov_len = sizeof(tag);
    msg.msg_iov = &io;
    msg.msg_iovlen = 1;
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof(ctrl);
    cmsghdr* c = CMSG_FIRSTHDR(&msg);
    c->cmsg_level = SOL_SOCKET;
    c->cmsg_type = SCM_RIGHTS;
    c->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(c), &fd, sizeof(int));
    return sendmsg(sock, &msg, 0) < 0 ? -errno : 0;
}

int recv_fd(int sock, int* fd, int* tag)
{
    msghdr msg;
    memset(&msg, 0, sizeof(msg));
    char ctrl[CMSG_SPACE(sizeof(int))];
    iovec io;
    io.iov_base = tag;
    io.iov_len = sizeof(*tag);
    msg.msg_iov = &io;
    msg.msg_iovlen = 1;
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof(ctrl);
    ssize_t n = recvmsg(sock, &msg, 0);
    if (n <= 0) return n < 0 ? -errno : -EPIPE;
    cmsghdr* c = CMSG_FIRSTHDR(&msg);
    if (!c || c->cmsg_type != SCM_RIGHTS) return -EPROTO;
    memcpy(fd, CMSG_DATA(c), sizeof(int));
    return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------
// capability queries
// ------------------------------------------------------------------------------------------------------
// bit0: VMM + POSIX-fd handles, bit1: multicast (NVLS)
DSB_EXPORT int dsb_symm_caps(int device)
{
    Driver& d = drv();
    if (!d.ok) return 0;
    CUdevice dev;
    if (d.DeviceGet(&dev, device) != CUDA_SUCCESS) return 0;
    int vmm = 0, fd = 0, mc = 0;
    d.DeviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev);
    d.DeviceGetAttribute(&fd, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev);
    if (d.mc_ok) d.DeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
    return ((vmm && fd) ? 1 : 0) | (mc ? 2 : 0);
}

DSB_EXPORT int64_t dsb_symm_granularity(int device, int world, int want_multicast)
{
    Driver& d = drv();
    if (!d.ok) return -1;
    CUmemAllocationProp p = alloc_prop(device);
    size_t g = 0;
    if (d.MemGetAllocationGranularity(&g, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS) return -1;
    if (want_multicast && d.mc_ok) {
        CUmulticastObjectProp mp;
        memset(&mp, 0, sizeof(mp));
        mp.numDevices = static_cast<unsigned>(world);
        mp.size = g;
        mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        size_t mg = 0;
        if (d.MulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > g) g = mg;
    }
    return static_cast<int64_t>(g);
}

// ------------------------------------------------------------------------------------------------------
// allocation / import
// ------------------------------------------------------------------------------------------------------
// Allocate `bytes` (multiple of the granularity) on `device`, map it locally, export an fd.
DSB_EXPORT int dsb_symm_alloc(int device, int64_t bytes, uint64_t* ptr_out, uint64_t* handle_out, int* fd_out)
{
    Driver& d = drv();
    if (!d.ok) return -1000;
    CUmemAllocationProp p = alloc_prop(device);
    CUmemGenericAllocationHandle h;
    CUresult r = d.MemCreate(&h, static_cast<size_t>(bytes), &p, 0);
    if (r != CUDA_SUCCESS) return -static_cast<int>(r);
    CUdeviceptr va = 0;
    int rc = map_rw(d, device, h, static_cast<size_t>(bytes), &va);
    if (rc != 0) {
        d.MemRelease(h);
        return rc;
    }
    int fd = -1;
    r = d.MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (r != CUDA_SUCCESS) {
        d.MemUnmap(va, static_cast<size_t>(bytes));
        d.MemAddressFree(va, static_cast<size_t>(bytes));
        d.MemRelease(h);
        return -static_cast<int>(r);
    }
    *ptr_out = static_cast<uint64_t>(va);
    *handle_out = static_cast<uint64_t>(h);
    *fd_out = fd;
    return 0;
}

// Import a peer allocation from its fd and map it read/write for `device`.
DSB_EXPORT int dsb_symm_import(int device, int fd, int64_t bytes, uint64_t* ptr_out, uint64_t* handle_out)
{
    Driver& d = drv();
    if (!d.ok) return -1000;
    CUmemGenericAllocationHandle h;
    CUresult r = d.MemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<intptr_t>(fd)),
                                                CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    if (r != CUDA_SUCCESS) return -static_cast<int>(r);
    CUdeviceptr va = 0;
    int rc = map_rw(d, device, h, static_cast<size_t>(bytes), &va);
    if (rc != 0) {
        d.MemRelease(h);
        return rc;
    }
    *ptr_out = static_cast<uint64_t>(va);
    *handle_out = static_cast<uint64_t>(h);
    return 0;
}

DSB_EXPORT int dsb_symm_unmap(uint64_t ptr, uint64_t handle, int64_t bytes)
{
    Driver& d = drv();
    if (!d.ok) return -1000;
    if (ptr) {
        d.MemUnmap(static_cast<CUdeviceptr>(ptr), static_cast<size_t>(bytes));
        d.MemAddressFree(static_cast<CUdeviceptr>(ptr), static_cast<size_t>(bytes));
    }
    if (handle) d.MemRelease(static_cast<CUmemGenericAllocationHandle>(handle));
    return 0;
}

DSB_EXPORT void dsb_close_fd(int fd)
{
    if (fd >= 0) close(fd);
}

// ------------------------------------------------------------------------------------------------------
// multicast (NVLS)
// ------------------------------------------------------------------------------------------------------
DSB_EXPORT int dsb_mc_create(int world, int64_t bytes, uint64_t* handle_out, int* fd_out)
{
    Driver& d = drv();
    if (!d.mc_ok) return -1000;
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof(mp));
    mp.numDevices = static_cast<unsigned>(world);
    mp.size = static_cast<size_t>(bytes);
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUmemGenericAllocationHandle h;
    CUresult r = d.MulticastCreate(&h, &mp);
    if (r != CUDA_SUCCESS) return -static_cast<int>(r);
    int fd = -1;
    r = d.MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (r != CUDA_SUCCESS) {
        d.MemRelease(h);
        return -static_cast<int>(r);
    }
    *handle_out = static_cast<uint64_t>(h);
    *fd_out = fd;
    return 0;
}

DSB_EXPORT int dsb_mc_import(int fd, uint64_t* handle_out)
{
    Driver& d = drv();
    if (!d.mc_ok) return -1000;
    CUmemGenericAllocationHandle h;
    CUresult r = d.MemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<intptr_t>(fd)),
                                                CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    if (r != CUDA_SUCCESS) return -static_cast<int>(r);
    *handle_out = static_cast<uint64_t>(h);
    return 0;
}

DSB_EXPORT int dsb_mc_add_device(uint64_t mc_handle, int device)
{
    Driver& d = drv();
    if (!d.mc_ok) return -1000;
    CUdevice dev;
    if (d.DeviceGet(&dev, device) != CUDA_SUCCESS) return -1;
    CUresult r = d.MulticastAddDevice(static_cast<CUmemGenericAllocationHandle>(mc_handle), dev);
    return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r);
}

// Bind this rank's physical allocation at offset 0 and map the multicast object.
DSB_EXPORT int dsb_mc_bind_and_map(uint64_t mc_handle, uint64_t mem_handle, int device, int64_t bytes,
                                   uint64_t* mc_ptr_out)
{
    Driver& d = drv();
    if (!d.mc_ok) return -1000;
    CUresult r = d.MulticastBindMem(static_cast<CUmemGenericAllocationHandle>(mc_handle), 0,
                                    static_cast<CUmemGenericAllocationHandle>(mem_handle), 0,
                                    static_cast<size_t>(bytes), 0);
    if (r != CUDA_SUCCESS) return -static_cast<int>(r);
    CUdeviceptr va = 0;
    int rc = map_rw(d, device, static_cast<CUmemGenericAllocationHandle>(mc_handle), static_cast<size_t>(bytes), &va);
    if (rc != 0) return rc;
    *mc_ptr_out = static_cast<uint64_t>(va);
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// all-to-all exchange of one fd per rank over abstract unix sockets
// ------------------------------------------------------------------------------------------------------
// fds_out[r] receives rank r's descriptor (fds_out[rank] = dup of my_fd).  Returns 0 or -errno.
DSB_EXPORT int dsb_exchange_fds(const char* prefix, int rank, int world, int my_fd, int* fds_out, int timeout_s)
{
    for (int i = 0; i < world; ++i) fds_out[i] = -1;
    fds_out[rank] = dup(my_fd);
    if (world == 1) return 0;
    int lsock = socket(AF_UNIX, SOCK_STREAM, 0);
    if (lsock < 0) return -errno;
    sockaddr_un addr;
    socklen_t alen;
    make_addr(prefix, rank, &addr, &alen);
    if (bind(lsock, reinterpret_cast<sockaddr*>(&addr), alen) != 0 || listen(lsock, world) != 0) {
        int e = errno;
        close(lsock);
        return -e;
    }
    int send_err = 0;
    std::thread sender([&] {
        for (int k = 1; k < world; ++k) {
            const int peer = (rank + k) % world;
            sockaddr_un pa;
            socklen_t pl;
            make_addr(prefix, peer, &pa, &pl);
            int s = socket(AF_UNIX, SOCK_STREAM, 0);
            if (s < 0) {
                send_err = -errno;
                return;
            }
            const time_t deadline = time(nullptr) + timeout_s;
            int rc = -1;
            while (time(nullptr) < deadline) {
                rc = connect(s, reinterpret_cast<sockaddr*>(&pa), pl);
                if (rc == 0) break;
                usleep(2000);  // peer's listener not up yet
            }
            if (rc != 0) {
                send_err = -ETIMEDOUT;
                close(s);
                return;
            }
            int e = send_fd(s, my_fd, rank);
            close(s);
            if (e != 0) {
                send_err = e;
                return;
            }
        }
    });
    int recv_err = 0;
    for (int k = 1; k < world; ++k) {
        timeval tv;
        tv.tv_sec = timeout_s;
        tv.tv_usec = 0;
        setsockopt(lsock, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
        int c = accept(lsock, nullptr, nullptr);
        if (c < 0) {
            recv_err = -errno;
            break;
        }
        int fd = -1, from = -1;
        int e = recv_fd(c, &fd, &from);
        close(c);
        if (e != 0 || from < 0 || from >= world) {
            recv_err = e ? e : -EPROTO;
            break;
        }
        fds_out[from] = fd;
    }
    sender.join();
    close(lsock);
    if (send_err) return send_err;
    return recv_err;
}

// ------------------------------------------------------------------------------------------------------
// copy-engine all-gather: (world-1) asynchronous peer->local DMA copies on `stream`, zero SM footprint.
// shards[p] = peer p's shard address (peer-mapped VA); own shard copied only if it is not already in place.
// ------------------------------------------------------------------------------------------------------
DSB_EXPORT int dsb_symm_all_gather_ce(void* const* shards, void* full, int64_t shard_bytes, int rank, int world,
                                      cudaStream_t stream)
{
    char* out = static_cast<char*>(full);
    for (int k = 1; k <= world; ++k) {
        const int peer = (rank + k) % world;  // start with rank+1: spreads load over the links
        char* dst = out + static_cast<int64_t>(peer) * shard_bytes;
        if (dst == shards[peer]) continue;  // in-place (stage <= 2): own shard already there
        cudaError_t e = cudaMemcpyAsync(dst, shards[peer], static_cast<size_t>(shard_bytes), cudaMemcpyDeviceToDevice,
                                        stream);
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    return 0;
}

// ---- exact-size page-locked host arenas ------------------------------------------------------------------------------------
// The ZeRO-Offload tier keeps fp32 master weights, Adam moments and the reduced gradient shard in pinned host memory -- four
// arenas of (parameters / ranks) * 4 bytes each.  torch's CachingHostAllocator rounds every request up to the next power of
// two (a 70 GB arena becomes 128 GB), which at Llama-70B scale overshoots the node's memory; these entry points give the
// arena exactly the bytes it asked for.  The pages are first-touched by several threads (page faulting is what dominates a
// >10 GB allocation) and then registered with the driver.
// Interleave the pages of the calling thread's next allocations over every NUMA node it may use (the CPU optimizer streams
// these arenas with all cores of the rank: one node's memory controllers would cap it).  Best effort: raw syscall, no libnuma.
static void set_interleave(bool on)
{
#if defined(__linux__) && defined(__x86_64__)
    constexpr long kSetMempolicy = 238;
    constexpr int kDefault = 0, kInterleave = 3;
    if (!on) {
        syscall(kSetMempolicy, kDefault, nullptr, 0);
        return;
    }
    unsigned long mask[16] = {0};
    FILE* f = fopen("/sys/devices/system/node/online", "r");
    if (f == nullptr) return;
    char buf[256] = {0};
    const bool got = fgets(buf, sizeof(buf), f) != nullptr;
    fclose(f);
    if (!got) return;
    int nodes = 0;
    for (char* tok = strtok(buf, ","); tok != nullptr; tok = strtok(nullptr, ",")) {
        int lo = 0, hi = 0;
        if (sscanf(tok, "%d-%d", &lo, &hi) == 2) {
        } else if (sscanf(tok, "%d", &lo) == 1) {
            hi = lo;
        } else {
            continue;
        }
        for (int n = lo; n <= hi && n < 1024; ++n) {
            mask[n / 64] |= 1ul << (n % 64);
            ++nodes;
        }
    }
    if (nodes > 1) syscall(kSetMempolicy, kInterleave, mask, 1024 + 1);
#else
    (void)on;
#endif
}

DSB_EXPORT int dsb_pinned_alloc(void** out, int64_t bytes, int touch_threads)
{
    if (bytes <= 0 || out == nullptr) return -2;
    void* p = nullptr;
    const size_t align = 2u << 20;
    const size_t sz = (static_cast<size_t>(bytes) + align - 1) / align * align;
    set_interleave(true);  // inherited by the first-touch threads below
    struct Restore {
        ~Restore() { set_interleave(false); }
    } restore;
    if (posix_memalign(&p, align, sz) != 0) return -ENOMEM;
    const int nt = touch_threads > 0 ? touch_threads : 1;
    std::vector<std::thread> th;
    const size_t per = (sz / 4096 + nt - 1) / nt * 4096;
    for (int t = 0; t < nt; ++t) {
        th.emplace_back([=]() {
            char* b = static_cast<char*>(p);
            const size_t lo = static_cast<size_t>(t) * per, hi = lo + per < sz ? lo + per : sz;
            for (size_t i = lo; i < hi; i += 4096) b[i] = 0;
        });
    }
    for (auto& t : th) t.join();
    cudaError_t e = cudaHostRegister(p, sz, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        free(p);
        return static_cast<int>(e);
    }
    *out = p;
    return 0;
}

DSB_EXPORT int dsb_pinned_free(void* p)
{
    if (p == nullptr) return 0;
    cudaError_t e = cudaHostUnregister(p);
    free(p);
    return static_cast<int>(e);
}
