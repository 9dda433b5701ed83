// Inter-GPU part of a halo exchange as ONE C call on the caller's communication stream:
//   fv3_halo_exchange_nccl <- the Isend / Irecv pairs HaloUpdater.start posts, one per neighbour
//                             (util/pace/util/halo_updater.py:217-303), here one grouped ncclSend / ncclRecv per peer GPU
//                             on the packed segments of fv3_halo_pack_segments / fv3_halo_unpack_segments.
// NCCL is not a link-time dependency of libfv3b200.so: its entry points are resolved at run time from the libnccl.so.2
// the process already has (the one torch.distributed loaded; never a second copy), so the library also loads on a box
// without NCCL.
#include <dlfcn.h>

#include "common.h"

namespace {

struct NcclUniqueId {
  char internal[128];
};
typedef void *ncclComm_t;
typedef int ncclResult_t;
constexpr int NCCL_FLOAT64 = 8;  // ncclDataType_t::ncclFloat64 (nccl.h)

struct NcclApi {
  ncclResult_t (*GetUniqueId)(NcclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, NcclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, void *) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, void *) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

const NcclApi &nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
#ifndef FV3_HOSTSIM
  // ONLY the copy the process already uses (torch.distributed's): loading a second libnccl.so.2 from the system path
  // would shadow the one a later `import torch` needs.  No NCCL in the process yet = not available (try again later).
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) {
    tried = false;
    return api;
  }
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
  api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
  api.Send = (decltype(api.Send))dlsym(h, "ncclSend");
  api.Recv = (decltype(api.Recv))dlsym(h, "ncclRecv");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send && api.Recv;
#endif
  return api;
}

int fail(const char *what, ncclResult_t r) {
  char msg[256];
  const NcclApi &a = nccl();
  snprintf(msg, sizeof msg, "%s: %s", what, a.GetErrorString ? a.GetErrorString(r) : "NCCL error");
  fv3::set_error(msg);
  return r ? r : -1;
}

}  // namespace

extern "C" {

int fv3_nccl_available(void) { return nccl().ok ? 1 : 0; }

// 128 bytes identifying a new communicator: rank 0 calls this and hands the bytes to the other ranks (any transport)
int fv3_nccl_unique_id(char *id128) {
  const NcclApi &a = nccl();
  if (!a.ok) {
    fv3::set_error("fv3_nccl_unique_id: libnccl.so.2 is not available in this process");
    return -1;
  }
  NcclUniqueId id;
  const ncclResult_t r = a.GetUniqueId(&id);
  if (r) return fail("ncclGetUniqueId", r);
  memcpy(id128, id.internal, 128);
  return 0;
}

// collective over the nranks processes (one per GPU, current CUDA device); *comm receives the ncclComm_t
int fv3_nccl_comm_create(void **comm, int nranks, const char *id128, int rank) {
  const NcclApi &a = nccl();
  if (!a.ok) {
    fv3::set_error("fv3_nccl_comm_create: libnccl.so.2 is not available in this process");
    return -1;
  }
  NcclUniqueId id;
  memcpy(id.internal, id128, 128);
  ncclComm_t c = nullptr;
  const ncclResult_t r = a.CommInitRank(&c, nranks, id, rank);
  if (r) return fail("ncclCommInitRank", r);
  *comm = c;
  return 0;
}

int fv3_nccl_comm_destroy(void *comm) {
  const NcclApi &a = nccl();
  if (!a.ok || !comm) return 0;
  const ncclResult_t r = a.CommDestroy((ncclComm_t)comm);
  return r ? fail("ncclCommDestroy", r) : 0;
}

// One grouped exchange: segment p of send_buf (send_cnt[p] doubles from send_off[p]) goes to process send_peer[p], segment
// p of recv_buf is filled by process recv_peer[p]; asynchronous on `stream` (capturable into a CUDA graph).
int fv3_halo_exchange_nccl(void *nccl_comm, const double *send_buf, const int64_t *send_off, const int64_t *send_cnt,
                           const int32_t *send_peer, int n_send, double *recv_buf, const int64_t *recv_off,
                           const int64_t *recv_cnt, const int32_t *recv_peer, int n_recv, void *stream) {
  const NcclApi &a = nccl();
  if (!a.ok || !nccl_comm) {
    fv3::set_error("fv3_halo_exchange_nccl: no NCCL communicator");
    return -1;
  }
  ncclResult_t r = a.GroupStart();
  if (r) return fail("ncclGroupStart", r);
  for (int p = 0; p < n_send && !r; ++p)
    r = a.Send(send_buf + send_off[p], (size_t)send_cnt[p], NCCL_FLOAT64, send_peer[p], (ncclComm_t)nccl_comm, stream);
  for (int p = 0; p < n_recv && !r; ++p)
    r = a.Recv(recv_buf + recv_off[p], (size_t)recv_cnt[p], NCCL_FLOAT64, recv_peer[p], (ncclComm_t)nccl_comm, stream);
  const ncclResult_t e = a.GroupEnd();
  if (r) return fail("ncclSend / ncclRecv", r);
  if (e) return fail("ncclGroupEnd", e);
  return 0;
}

}  // extern "C"
