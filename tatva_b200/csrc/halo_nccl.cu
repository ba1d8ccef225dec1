// Halo exchange over an ncclComm_t, entirely behind the C ABI: pack kernel -> ONE grouped ncclSend / ncclRecv with every
// neighbour -> unpack kernel, all enqueued on the caller's stream (tatva/mpi.py:372-409 forward fill, :479-516 reverse
// add; there one blocking mpi4jax.sendrecv per neighbour, serialised, :403-405 / :509-511).  No allocation and no
// synchronisation here: the staging buffers are the caller's.  NCCL is opened with dlopen at first use — the copy the
// process already holds (torch's bundled libnccl.so.2) if there is one, else the system library — so libtatva_b200.so
// itself has no link-time dependency on it and loads on machines without NCCL.
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace {

// The slice of nccl.h this file needs (stable since NCCL 2.7: point-to-point + groups).
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclSuccess = 0, kNcclFloat64 = 8 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*CommCount)(const ncclComm_t, int*) = nullptr;
  int (*CommUserRank)(const ncclComm_t, int*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names)
      if ((api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD))) break;  // the copy already in the process
    if (!api.handle)
      for (const char* n : names)
        if ((api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!api.handle) return;
#define TATVA_NCCL_SYM(field, name) \
  if (!(*reinterpret_cast<void**>(&api.field) = dlsym(api.handle, name))) return;
    TATVA_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    TATVA_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    TATVA_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    TATVA_NCCL_SYM(CommCount, "ncclCommCount")
    TATVA_NCCL_SYM(CommUserRank, "ncclCommUserRank")
    TATVA_NCCL_SYM(GroupStart, "ncclGroupStart")
    TATVA_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    TATVA_NCCL_SYM(Send, "ncclSend")
    TATVA_NCCL_SYM(Recv, "ncclRecv")
    TATVA_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef TATVA_NCCL_SYM
    api.ok = true;
  });
  return api;
}

}  // namespace

extern "C" {

int tatva_nccl_version(int* version) {
  if (!version) return TATVA_E_INVALID;
  NcclApi& a = nccl();
  if (!a.ok) return TATVA_E_UNSUPPORTED;
  return a.GetVersion(version) == kNcclSuccess ? TATVA_OK : TATVA_E_INVALID;
}

// For hosts that have no communicator of their own to hand over (an MPI or JAX host passes its ncclComm_t straight to
// tatva_halo_exchange): rank 0 draws the 128-byte id, the host ships it to every rank by its own means, every rank calls
// tatva_halo_comm_create (collective, blocking) on its device.
int tatva_halo_comm_unique_id(void* id128) {
  if (!id128) return TATVA_E_INVALID;
  NcclApi& a = nccl();
  if (!a.ok) return TATVA_E_UNSUPPORTED;
  return a.GetUniqueId(reinterpret_cast<ncclUniqueId*>(id128)) == kNcclSuccess ? TATVA_OK : TATVA_E_INVALID;
}

int tatva_halo_comm_create(void** nccl_comm, const void* id128, int n_ranks, int rank) {
  if (!nccl_comm || !id128 || n_ranks <= 0 || rank < 0 || rank >= n_ranks) return TATVA_E_INVALID;
  NcclApi& a = nccl();
  if (!a.ok) return TATVA_E_UNSUPPORTED;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  if (a.CommInitRank(&c, n_ranks, id, rank) != kNcclSuccess) return TATVA_E_INVALID;
  *nccl_comm = c;
  return TATVA_OK;
}

int tatva_halo_comm_destroy(void* nccl_comm) {
  if (!nccl_comm) return TATVA_OK;
  NcclApi& a = nccl();
  if (!a.ok) return TATVA_E_UNSUPPORTED;
  return a.CommDestroy(reinterpret_cast<ncclComm_t>(nccl_comm)) == kNcclSuccess ? TATVA_OK : TATVA_E_INVALID;
}

// One direction of an ExchangePlan on `nccl_comm` (an ncclComm_t):
//   d_send_buf[k] = d_src[d_send_idx[k]]                      k < sum(send_counts)        (tatva_halo_pack)
//   rank r receives send_counts[r] doubles, sends us recv_counts[r]  (grouped ncclSend / ncclRecv, rank order)
//   d_dst[d_recv_idx[k]] (+)= d_recv_buf[k]                   k < sum(recv_counts)        (tatva_halo_unpack_set / _add)
// send_counts / recv_counts: HOST arrays of n_ranks entries (the entry of the own rank must be 0: local copies are not
// an exchange).  d_src may equal d_dst (ghost refresh and reverse add both work in place on a local vector).
int tatva_halo_exchange(void* nccl_comm, const double* d_src, const int64_t* d_send_idx, const int64_t* send_counts,
                        double* d_send_buf, double* d_recv_buf, const int64_t* recv_counts, const int64_t* d_recv_idx,
                        double* d_dst, int add, tatva_stream_t stream) {
  if (!nccl_comm || !d_src || !d_dst || !send_counts || !recv_counts) return TATVA_E_INVALID;
  NcclApi& a = nccl();
  if (!a.ok) return TATVA_E_UNSUPPORTED;
  ncclComm_t comm = reinterpret_cast<ncclComm_t>(nccl_comm);
  cudaStream_t st = (cudaStream_t)stream;
  int n_ranks = 0, me = -1;
  if (a.CommCount(comm, &n_ranks) != kNcclSuccess || a.CommUserRank(comm, &me) != kNcclSuccess) return TATVA_E_INVALID;
  int64_t n_send = 0, n_recv = 0;
  for (int r = 0; r < n_ranks; ++r) {
    if (send_counts[r] < 0 || recv_counts[r] < 0 || (r == me && (send_counts[r] || recv_counts[r]))) return TATVA_E_INVALID;
    n_send += send_counts[r];
    n_recv += recv_counts[r];
  }
  if ((n_send && (!d_send_idx || !d_send_buf)) || (n_recv && (!d_recv_idx || !d_recv_buf))) return TATVA_E_INVALID;
  if (n_send) {
    const int rc = tatva_halo_pack(d_src, d_send_idx, n_send, d_send_buf, stream);
    if (rc != TATVA_OK) return rc;
  }
  if (n_send || n_recv) {
    if (a.GroupStart() != kNcclSuccess) return TATVA_E_INVALID;
    int64_t so = 0, ro = 0;
    int bad = 0;
    for (int r = 0; r < n_ranks; ++r) {
      if (send_counts[r]) bad |= a.Send(d_send_buf + so, (size_t)send_counts[r], kNcclFloat64, r, comm, st) != kNcclSuccess;
      if (recv_counts[r]) bad |= a.Recv(d_recv_buf + ro, (size_t)recv_counts[r], kNcclFloat64, r, comm, st) != kNcclSuccess;
      so += send_counts[r];
      ro += recv_counts[r];
    }
    if (a.GroupEnd() != kNcclSuccess || bad) return TATVA_E_INVALID;
  }
  if (n_recv) {
    const int rc = add ? tatva_halo_unpack_add(d_recv_buf, d_recv_idx, n_recv, d_dst, stream)
                       : tatva_halo_unpack_set(d_recv_buf, d_recv_idx, n_recv, d_dst, stream);
    if (rc != TATVA_OK) return rc;
  }
  return TATVA_OK;
}

}  // extern "C"
