// comm.cu -- collectives for the time-frame-sharded mode: NCCL (dlopen) or caller callbacks.  See comm.h.
#include "comm.h"

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace ainmf {

namespace {
// Minimal NCCL ABI (nccl.h 2.x): opaque comm, 128-byte unique id passed by value, enums below.
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
enum { kNcclInt32 = 2, kNcclFloat32 = 7, kNcclFloat64 = 8 };
enum { kNcclSum = 0, kNcclMax = 2 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

int load_nccl(char* err, size_t errlen) {
    if (g_nccl.lib) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) { snprintf(err, errlen, "cannot dlopen libnccl.so.2: %s", dlerror()); return -1; }
#define SYM(field, name)                                                                         \
    *(void**)(&g_nccl.field) = dlsym(lib, name);                                                  \
    if (!g_nccl.field) { snprintf(err, errlen, "libnccl lacks %s", name); dlclose(lib); return -1; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = lib;
    return 0;
}
}  // namespace

constexpr int kMaxPeers = 16;
struct PeerPtrs { char* base[kMaxPeers]; };

struct Comm {
    int rank = 0, nranks = 1;
    NcclComm nccl = nullptr;
    comm_allreduce_cb ar = nullptr;
    comm_sendrecv_cb sr = nullptr;
    void* user = nullptr;
    // peer mailboxes (see comm.h): box layout = [2 parities][nranks] int flags (256-byte block), then
    // [2 parities][nranks sources][cap] payload slots
    char* box = nullptr;              // this rank's mailbox (cudaMalloc, exported through CUDA IPC)
    PeerPtrs peers{};                 // every rank's mailbox as mapped here (own entry = box)
    size_t cap = 0;                   // payload bytes per slot
    unsigned* push_count = nullptr;   // [nranks] blocks that finished their share of a push
    unsigned epoch = 0;               // one per exchange, the same on every rank; parity = epoch & 1
    unsigned pushed = 0;              // blocks launched by all pushes so far (per destination)
};

#ifndef AINMF_EMU
namespace {
constexpr size_t kFlagBytes = 256;
constexpr int kPushChunk = 16384;     // bytes per block of the push kernel

__device__ __forceinline__ size_t slot_offset(int parity, int src, int nranks, size_t cap) {
    return kFlagBytes + ((size_t)parity * nranks + src) * cap;
}

// grid = (chunks, nranks): block (x, d) copies chunk x of this rank's contribution into its slot on rank d; the last
// block to finish for d publishes the epoch in d's flag word (release, system scope).
__global__ void __launch_bounds__(256)
peer_push_kernel(PeerPtrs peers, const char* __restrict__ src, size_t bytes, int rank, int nranks, size_t cap, unsigned epoch,
                 unsigned* __restrict__ push_count, unsigned target) {
    const int d = blockIdx.y, parity = (int)(epoch & 1u);
    char* dst = peers.base[d] + slot_offset(parity, rank, nranks, cap);
    const size_t b0 = (size_t)blockIdx.x * kPushChunk;
    const size_t b1 = b0 + kPushChunk < bytes ? b0 + kPushChunk : bytes;
    if ((bytes & 15) == 0) {
        for (size_t o = b0 + 16 * (size_t)threadIdx.x; o < b1; o += 16 * 256)
            *reinterpret_cast<float4*>(dst + o) = *reinterpret_cast<const float4*>(src + o);
    } else {                          // small payloads (a scalar): 4-byte words
        for (size_t o = b0 + 4 * (size_t)threadIdx.x; o < b1; o += 4 * 256)
            *reinterpret_cast<float*>(dst + o) = *reinterpret_cast<const float*>(src + o);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(&push_count[d], 1u) + 1u;      // running total over all pushes (wraps consistently)
        if (done == target) {
            __threadfence_system();
            unsigned* flag = reinterpret_cast<unsigned*>(peers.base[d]) + parity * nranks + rank;
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
        }
    }
}

// Waits until every rank's contribution of this epoch has landed in this rank's mailbox, then out = sum of the slots in
// rank order (the same order on every rank: replicated results stay bit-identical).  A peer that never arrives traps
// after ~10 s instead of hanging the device.
template <typename T4, typename T>
__global__ void __launch_bounds__(256)
peer_sum_kernel(const char* __restrict__ box, char* __restrict__ out, size_t bytes, int nranks, size_t cap, unsigned epoch) {
    const int parity = (int)(epoch & 1u);
    if (threadIdx.x < nranks) {
        const unsigned* flag = reinterpret_cast<const unsigned*>(box) + parity * nranks + threadIdx.x;
        unsigned v;
        long long t0 = clock64();
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v == epoch) break;
            if (clock64() - t0 > 20000000000LL) asm volatile("trap;");
        }
    }
    __syncthreads();
    const char* s0 = box + slot_offset(parity, 0, nranks, cap);
    if ((bytes % sizeof(T4)) == 0) {
        const size_t n = bytes / sizeof(T4);
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
            T4 acc = __ldcg(reinterpret_cast<const T4*>(s0) + i);
            for (int g = 1; g < nranks; ++g) {
                const T4 v = __ldcg(reinterpret_cast<const T4*>(s0 + (size_t)g * cap) + i);
                acc.x += v.x; acc.y += v.y;
                if constexpr (sizeof(T4) == 4 * sizeof(T)) { acc.z += v.z; acc.w += v.w; }
            }
            reinterpret_cast<T4*>(out)[i] = acc;
        }
    } else {
        const size_t n = bytes / sizeof(T);
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
            T acc = __ldcg(reinterpret_cast<const T*>(s0) + i);
            for (int g = 1; g < nranks; ++g) acc += __ldcg(reinterpret_cast<const T*>(s0 + (size_t)g * cap) + i);
            reinterpret_cast<T*>(out)[i] = acc;
        }
    }
}
}  // namespace
#endif

int comm_unique_id(uint8_t id_out[128], char* err, size_t errlen) {
    if (load_nccl(err, errlen)) return -1;
    NcclUniqueId id;
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc) { snprintf(err, errlen, "ncclGetUniqueId: %s", g_nccl.GetErrorString(rc)); return -1; }
    memcpy(id_out, id.internal, 128);
    return 0;
}

int comm_create_nccl(Comm** out, const uint8_t id_in[128], int rank, int nranks, char* err, size_t errlen) {
    if (load_nccl(err, errlen)) return -1;
    NcclUniqueId id;
    memcpy(id.internal, id_in, 128);
    Comm* c = new Comm();
    c->rank = rank; c->nranks = nranks;
    const int rc = g_nccl.CommInitRank(&c->nccl, nranks, id, rank);
    if (rc) { snprintf(err, errlen, "ncclCommInitRank: %s", g_nccl.GetErrorString(rc)); delete c; return -1; }
    *out = c;
    return 0;
}

int comm_create_callbacks(Comm** out, int rank, int nranks, comm_allreduce_cb ar, comm_sendrecv_cb sr, void* user) {
    Comm* c = new Comm();
    c->rank = rank; c->nranks = nranks; c->ar = ar; c->sr = sr; c->user = user;
    *out = c;
    return 0;
}

#ifndef AINMF_EMU
static void peer_teardown(Comm* c) {
    if (!c->box) return;
    cudaDeviceSynchronize();
    for (int r = 0; r < c->nranks; ++r)
        if (r != c->rank && c->peers.base[r]) cudaIpcCloseMemHandle(c->peers.base[r]);
    cudaFree(c->box);
    cudaFree(c->push_count);
    c->box = nullptr; c->push_count = nullptr; c->cap = 0; c->pushed = 0;
    for (int r = 0; r < kMaxPeers; ++r) c->peers.base[r] = nullptr;
}
#endif

int comm_peer_active(const Comm* c) { return (c && c->box) ? 1 : 0; }

int comm_peer_setup(Comm* c, size_t cap_bytes, cudaStream_t s, char* err, size_t errlen) {
#ifdef AINMF_EMU
    (void)c; (void)cap_bytes; (void)s; (void)err; (void)errlen;
    return 0;
#else
    if (!c || c->nranks == 1 || !c->nccl || c->nranks > kMaxPeers) return 0;      // callbacks transport: no mailboxes
    const char* off = getenv("AINMF_PEER_EXCHANGE");
    if (off && off[0] == '0') return 0;
    cap_bytes = (cap_bytes + 255) / 256 * 256;
    if (c->box && cap_bytes <= c->cap) return 0;
    cudaError_t e;
#define PCU(call) do { if ((e = (call)) != cudaSuccess) { snprintf(err, errlen, "peer mailbox: %s: %s", #call, cudaGetErrorString(e)); return -1; } } while (0)
    // (re)allocation is collective: nobody frees a mailbox a peer may still be writing to
    PCU(cudaStreamSynchronize(s));
    if (c->box) {
        int* d_bar = reinterpret_cast<int*>(c->push_count);
        const int rc = g_nccl.AllReduce(d_bar, d_bar, 1, kNcclInt32, kNcclMax, c->nccl, s);
        if (rc) { snprintf(err, errlen, "ncclAllReduce: %s", g_nccl.GetErrorString(rc)); return -1; }
        PCU(cudaStreamSynchronize(s));
        peer_teardown(c);
    }
    const size_t total = kFlagBytes + 2 * (size_t)c->nranks * cap_bytes;
    PCU(cudaMalloc(&c->box, total));
    PCU(cudaMemset(c->box, 0, kFlagBytes));
    PCU(cudaMalloc(&c->push_count, sizeof(unsigned) * kMaxPeers));
    PCU(cudaMemset(c->push_count, 0, sizeof(unsigned) * kMaxPeers));
    // every rank's IPC handle, gathered as an int32 sum of one-hot rows (the transport we already have)
    constexpr int HW = (int)(sizeof(cudaIpcMemHandle_t) / sizeof(int));
    cudaIpcMemHandle_t mine;
    PCU(cudaIpcGetMemHandle(&mine, c->box));
    int* d_h = nullptr;
    PCU(cudaMalloc(&d_h, sizeof(int) * HW * c->nranks));
    PCU(cudaMemset(d_h, 0, sizeof(int) * HW * c->nranks));
    PCU(cudaMemcpy(d_h + (size_t)HW * c->rank, &mine, sizeof mine, cudaMemcpyHostToDevice));
    PCU(cudaDeviceSynchronize());
    const int rc = g_nccl.AllReduce(d_h, d_h, (size_t)HW * c->nranks, kNcclInt32, kNcclSum, c->nccl, s);
    if (rc) { snprintf(err, errlen, "ncclAllReduce: %s", g_nccl.GetErrorString(rc)); return -1; }
    PCU(cudaStreamSynchronize(s));
    cudaIpcMemHandle_t all[kMaxPeers];
    PCU(cudaMemcpy(all, d_h, sizeof(cudaIpcMemHandle_t) * c->nranks, cudaMemcpyDeviceToHost));
    PCU(cudaFree(d_h));
    for (int r = 0; r < c->nranks; ++r) {
        if (r == c->rank) { c->peers.base[r] = c->box; continue; }
        void* p = nullptr;
        PCU(cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess));
        c->peers.base[r] = (char*)p;
    }
    c->cap = cap_bytes;
#undef PCU
    return 0;
#endif
}

void comm_destroy(Comm* c) {
    if (!c) return;
#ifndef AINMF_EMU
    peer_teardown(c);
#endif
    if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl);
    delete c;
}

int comm_rank(const Comm* c) { return c ? c->rank : 0; }
int comm_size(const Comm* c) { return c ? c->nranks : 1; }

int comm_allreduce(Comm* c, void* buf, size_t count, int dtype, int op, cudaStream_t s, char* err, size_t errlen) {
    if (!c || c->nranks == 1 || count == 0) return 0;
    if (c->ar) {
        const int rc = c->ar(c->user, buf, count, dtype, op, (void*)s);
        if (rc) snprintf(err, errlen, "all-reduce callback failed (%d)", rc);
        return rc;
    }
#ifndef AINMF_EMU
    {
        const size_t bytes = count * (dtype == COMM_F64 ? 8 : 4);
        if (c->box && op == COMM_SUM && (dtype == COMM_F32 || dtype == COMM_F64) && bytes <= c->cap) {
            const unsigned epoch = ++c->epoch;
            const unsigned chunks = (unsigned)((bytes + kPushChunk - 1) / kPushChunk);
            c->pushed += chunks;
            AINMF_LAUNCH(peer_push_kernel, dim3(chunks, c->nranks), dim3(256), 0, s, c->peers, (const char*)buf, bytes, c->rank,
                         c->nranks, c->cap, epoch, c->push_count, c->pushed);
            unsigned blocks = (unsigned)((bytes / 16 + 255) / 256);
            if (blocks < 1) blocks = 1;
            if (blocks > 296) blocks = 296;
            auto sum32 = peer_sum_kernel<float4, float>;
            auto sum64 = peer_sum_kernel<double2, double>;
            if (dtype == COMM_F32) AINMF_LAUNCH(sum32, dim3(blocks), dim3(256), 0, s, c->box, (char*)buf, bytes, c->nranks, c->cap, epoch);
            else AINMF_LAUNCH(sum64, dim3(blocks), dim3(256), 0, s, c->box, (char*)buf, bytes, c->nranks, c->cap, epoch);
            const cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { snprintf(err, errlen, "peer exchange: %s", cudaGetErrorString(e)); return -1; }
            return 0;
        }
    }
#endif
    const int dt = dtype == COMM_F32 ? kNcclFloat32 : (dtype == COMM_F64 ? kNcclFloat64 : kNcclInt32);
    const int rc = g_nccl.AllReduce(buf, buf, count, dt, op == COMM_SUM ? kNcclSum : kNcclMax, c->nccl, s);
    if (rc) { snprintf(err, errlen, "ncclAllReduce: %s", g_nccl.GetErrorString(rc)); return -1; }
    return 0;
}

int comm_sendrecv(Comm* c, const void* sendbuf, int send_peer, void* recvbuf, int recv_peer, size_t n_floats,
                  cudaStream_t s, char* err, size_t errlen) {
    if (!c || c->nranks == 1 || n_floats == 0 || (send_peer < 0 && recv_peer < 0)) return 0;
    if (c->sr) {
        const int rc = c->sr(c->user, sendbuf, send_peer, recvbuf, recv_peer, n_floats, (void*)s);
        if (rc) snprintf(err, errlen, "send/recv callback failed (%d)", rc);
        return rc;
    }
    int rc = g_nccl.GroupStart();
    if (!rc && send_peer >= 0) rc = g_nccl.Send(sendbuf, n_floats, kNcclFloat32, send_peer, c->nccl, s);
    if (!rc && recv_peer >= 0) rc = g_nccl.Recv(recvbuf, n_floats, kNcclFloat32, recv_peer, c->nccl, s);
    const int rc2 = g_nccl.GroupEnd();
    if (rc || rc2) { snprintf(err, errlen, "ncclSend/Recv: %s", g_nccl.GetErrorString(rc ? rc : rc2)); return -1; }
    return 0;
}

}  // namespace ainmf
