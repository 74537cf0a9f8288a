// comm.cu -- collectives for the time-frame-sharded mode: NCCL (dlopen) or caller callbacks.  See comm.h.
#include "comm.h"

#include <dlfcn.h>
#include <stdio.h>
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

struct Comm {
    int rank = 0, nranks = 1;
    NcclComm nccl = nullptr;
    comm_allreduce_cb ar = nullptr;
    comm_sendrecv_cb sr = nullptr;
    void* user = nullptr;
};

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

void comm_destroy(Comm* c) {
    if (!c) return;
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
