// comm.h -- the collective layer of the time-frame-sharded mode.
//
// Two back ends behind one interface:
//   * NCCL over NVLink / NVSwitch, loaded at run time with dlopen("libnccl.so.2") so that libainmf.so has no
//     link-time dependency (inside a torch process this resolves to the NCCL torch already loaded);
//   * caller-supplied callbacks (ainmf_comm_set_callbacks) -- used by the CPU test-suite to run the sharded
//     algorithm over torch.distributed/gloo, and available to hosts that bring their own transport.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "common.cuh"

namespace ainmf {

enum CommDtype { COMM_F32 = 0, COMM_F64 = 1, COMM_I32 = 2 };
enum CommOp { COMM_SUM = 0, COMM_MAX = 1 };

typedef int (*comm_allreduce_cb)(void* user, void* buf, size_t count, int dtype, int op, void* stream);
typedef int (*comm_sendrecv_cb)(void* user, const void* sendbuf, int send_peer, void* recvbuf, int recv_peer,
                                size_t n_floats, void* stream);

struct Comm;

int comm_unique_id(uint8_t id_out[128], char* err, size_t errlen);
int comm_create_nccl(Comm** out, const uint8_t id[128], int rank, int nranks, char* err, size_t errlen);
int comm_create_callbacks(Comm** out, int rank, int nranks, comm_allreduce_cb ar, comm_sendrecv_cb sr, void* user);
void comm_destroy(Comm* c);
int comm_rank(const Comm* c);
int comm_size(const Comm* c);
// in-place all-reduce of `count` elements on `s`
int comm_allreduce(Comm* c, void* buf, size_t count, int dtype, int op, cudaStream_t s, char* err, size_t errlen);
// Peer mailboxes (NCCL back end only): every rank maps a slot buffer of every other rank (CUDA IPC over NVLink) so that
// the per-iteration all-reduce becomes two small kernels of our own -- push this rank's contribution into its slot on
// every rank, then sum the slots in rank order -- instead of a library collective between the kernels.  Collective:
// every rank calls it with the same capacity.  After it, comm_allreduce routes float32/float64 sums of up to
// `cap_bytes` through the mailboxes; AINMF_PEER_EXCHANGE=0 keeps everything on ncclAllReduce.
int comm_peer_setup(Comm* c, size_t cap_bytes, cudaStream_t s, char* err, size_t errlen);
int comm_peer_active(const Comm* c);
// simultaneous send to `send_peer` and receive from `recv_peer` (either may be -1 = none), float payloads
int comm_sendrecv(Comm* c, const void* sendbuf, int send_peer, void* recvbuf, int recv_peer, size_t n_floats,
                  cudaStream_t s, char* err, size_t errlen);

}  // namespace ainmf
