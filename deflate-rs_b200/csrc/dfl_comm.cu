// dfl_comm.cu -- the one exchange step of the sharded encode path: the compressed
// streams of all ranks are brought to one rank over NCCL (NVLink 5 / NVSwitch inside a node).
//
// The encode itself needs no collective: DEFLATE blocks only refer to plaintext every rank already holds.
// What remains is a gather-v of byte streams whose sizes are only known after the encode.  NCCL has no
// gather-v, so: one ncclAllGather of the 8-byte sizes, then one group of ncclSend / ncclRecv at the
// prefix offsets, everything on the caller's stream so that the transfer of one batch of streams runs while
// the next batch is being encoded.
//
// NCCL is loaded at run time (dlopen of libnccl.so.2 -- the copy PyTorch ships when the caller is a
// torch.distributed job, the system's otherwise): the library stays loadable, and every other entry point
// usable, on a box without NCCL; the calls below return DFL_E_NCCL there.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/deflate_b200.h"

namespace {

struct NcclApi {
    void* so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            api.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.so) break;
        }
        if (!api.so) { api.why = std::string("dlopen(libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"); return; }
#define DFL_SYM(field, name)                                                       \
        api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.so, name));    \
        if (!api.field) { api.why = std::string("missing NCCL symbol ") + name; return; }
        DFL_SYM(GetUniqueId, "ncclGetUniqueId")
        DFL_SYM(CommInitRank, "ncclCommInitRank")
        DFL_SYM(CommDestroy, "ncclCommDestroy")
        DFL_SYM(AllGather, "ncclAllGather")
        DFL_SYM(Send, "ncclSend")
        DFL_SYM(Recv, "ncclRecv")
        DFL_SYM(GroupStart, "ncclGroupStart")
        DFL_SYM(GroupEnd, "ncclGroupEnd")
        DFL_SYM(GetErrorString, "ncclGetErrorString")
#undef DFL_SYM
        api.ok = true;
    });
    return api;
}

thread_local std::string t_comm_err;

int nccl_fail(ncclResult_t r, const char* where) {
    NcclApi& a = nccl();
    t_comm_err = std::string(where) + ": " + (a.GetErrorString ? a.GetErrorString(r) : "NCCL error");
    return DFL_E_NCCL;
}
int cuda_fail(cudaError_t e, const char* where) {
    t_comm_err = std::string(where) + ": " + cudaGetErrorString(e);
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? DFL_E_NODEVICE : DFL_E_CUDA;
}
#define NK(expr)                                                  \
    do {                                                          \
        ncclResult_t r_ = (expr);                                 \
        if (r_ != ncclSuccess) return nccl_fail(r_, #expr);       \
    } while (0)
#define CK(expr)                                                  \
    do {                                                          \
        cudaError_t e_ = (expr);                                  \
        if (e_ != cudaSuccess) return cuda_fail(e_, #expr);       \
    } while (0)

}  // namespace

struct dfl_comm {
    ncclComm_t comm = nullptr;
    int world = 0, rank = 0, device = 0;
    unsigned long long* d_sizes = nullptr;   // world + 1 words: [0..world) gathered sizes, [world] this rank's size
    unsigned long long* h_sizes = nullptr;   // pinned, world words
    cudaEvent_t ev = nullptr;
};

static_assert(sizeof(ncclUniqueId) == DFL_COMM_ID_BYTES, "dfl_comm_unique_id hands out a whole ncclUniqueId");

extern "C" const char* dfl_comm_last_error(void) { return t_comm_err.c_str(); }

extern "C" int dfl_comm_unique_id(uint8_t* id) {
    if (!id) return DFL_E_ARG;
    NcclApi& a = nccl();
    if (!a.ok) { t_comm_err = a.why; return DFL_E_NCCL; }
    ncclUniqueId u;
    NK(a.GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return DFL_OK;
}

extern "C" int dfl_comm_init(dfl_comm** out, int world, int rank, const uint8_t* id) {
    if (!out || !id || world < 1 || rank < 0 || rank >= world) return DFL_E_ARG;
    *out = nullptr;
    NcclApi& a = nccl();
    if (!a.ok) { t_comm_err = a.why; return DFL_E_NCCL; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { (void)cudaGetLastError(); t_comm_err = "no CUDA device"; return DFL_E_NODEVICE; }
    dfl_comm* c = new dfl_comm();
    c->world = world; c->rank = rank;
    cudaError_t e = cudaGetDevice(&c->device);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaGetDevice"); }
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclResult_t r = a.CommInitRank(&c->comm, world, u, rank);
    if (r != ncclSuccess) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
    if ((e = cudaMalloc(reinterpret_cast<void**>(&c->d_sizes), (size_t)(world + 1) * 8)) != cudaSuccess ||
        (e = cudaMallocHost(reinterpret_cast<void**>(&c->h_sizes), (size_t)(world + 1) * 8)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c->ev, cudaEventDisableTiming)) != cudaSuccess) {
        dfl_comm_free(c);
        return cuda_fail(e, "dfl_comm_init: scratch");
    }
    *out = c;
    return DFL_OK;
}

extern "C" void dfl_comm_free(dfl_comm* c) {
    if (!c) return;
    if (c->ev) cudaEventDestroy(c->ev);
    if (c->h_sizes) cudaFreeHost(c->h_sizes);
    if (c->d_sizes) cudaFree(c->d_sizes);
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    delete c;
}

extern "C" int dfl_comm_world(const dfl_comm* c) { return c ? c->world : 0; }
extern "C" int dfl_comm_rank(const dfl_comm* c) { return c ? c->rank : -1; }

// Every rank contributes the first n bytes at d_src; on `root` they arrive back to back, in rank order, at d_dst.
// sizes (host, world entries, may be NULL) receives every rank's n on every rank.  The size exchange costs one
// short host wait (the counts of ncclSend/ncclRecv are host arguments); the payload moves asynchronously on
// `stream`: the call returns once it is queued.
extern "C" int dfl_gather_device(dfl_comm* c, const void* d_src, size_t n, void* d_dst, size_t dst_cap, size_t* sizes, int root,
                                 void* stream) {
    if (!c || root < 0 || root >= c->world || (n && !d_src)) return DFL_E_ARG;
    NcclApi& a = nccl();
    if (!a.ok) { t_comm_err = a.why; return DFL_E_NCCL; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int W = c->world;
    c->h_sizes[W] = (unsigned long long)n;
    CK(cudaMemcpyAsync(c->d_sizes + W, c->h_sizes + W, 8, cudaMemcpyHostToDevice, st));
    NK(a.AllGather(c->d_sizes + W, c->d_sizes, 1, ncclUint64, c->comm, st));
    CK(cudaMemcpyAsync(c->h_sizes, c->d_sizes, (size_t)W * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(c->ev, st));
    CK(cudaEventSynchronize(c->ev));
    size_t total = 0;
    for (int r = 0; r < W; r++) { if (sizes) sizes[r] = (size_t)c->h_sizes[r]; total += (size_t)c->h_sizes[r]; }
    if (c->rank == root) {
        if (!d_dst || total > dst_cap) { t_comm_err = "dfl_gather_device: destination too small"; return DFL_E_OVERFLOW; }
        uint8_t* dst = static_cast<uint8_t*>(d_dst);
        size_t off = 0, my_off = 0;
        NK(a.GroupStart());
        for (int r = 0; r < W; r++) {
            const size_t len = (size_t)c->h_sizes[r];
            if (r == root) my_off = off;
            else if (len) {
                ncclResult_t rr = a.Recv(dst + off, len, ncclUint8, r, c->comm, st);
                if (rr != ncclSuccess) { a.GroupEnd(); return nccl_fail(rr, "ncclRecv"); }
            }
            off += len;
        }
        NK(a.GroupEnd());
        if (n && dst + my_off != d_src) CK(cudaMemcpyAsync(dst + my_off, d_src, n, cudaMemcpyDeviceToDevice, st));
    } else if (n) {
        NK(a.GroupStart());
        ncclResult_t rr = a.Send(d_src, n, ncclUint8, root, c->comm, st);
        if (rr != ncclSuccess) { a.GroupEnd(); return nccl_fail(rr, "ncclSend"); }
        NK(a.GroupEnd());
    }
    return DFL_OK;
}
