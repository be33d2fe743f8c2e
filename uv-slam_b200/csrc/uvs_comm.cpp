// uvs_comm.cpp — NCCL for the factor-parallel mode (uvs_comm_init_nccl), bound at run time.
//
// The exchange step of the hot path: every rank owns the landmarks k with k % nranks == rank, builds a
// partial reduced camera system, and ONE ncclAllReduce per LM iteration sums [S | gS | g | column norms | per-window
// accumulators] over NVLink / NVSwitch.  The reference's only reduction of this kind is the four-thread sum of
// MarginalizationInfo (factor/marginalization_factor.cpp:232-261).
//
// libnccl.so.2 is dlopen()ed on first use: a process that already holds NCCL (torch.distributed) shares that copy, a
// plain C++ host gets the system library, and single-GPU users need no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "uvs_handle.h"

namespace uvs {
namespace {
struct Api {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
Api &api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib) break;
    }
    if (!a.lib) return;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.lib, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.lib, "ncclCommInitRank");
    a.AllReduce = (decltype(a.AllReduce))dlsym(a.lib, "ncclAllReduce");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.CommDestroy;
  });
  return a;
}
constexpr int kNoLibrary = 1000;
}  // namespace

int nccl_unique_id(unsigned char id[128]) {
  Api &a = api();
  if (!a.ok) return kNoLibrary;
  ncclUniqueId u;
  const ncclResult_t r = a.GetUniqueId(&u);
  if (r != ncclSuccess) return (int)r;
  static_assert(sizeof(u) == 128, "ncclUniqueId is 128 bytes");
  std::memcpy(id, &u, 128);
  return 0;
}

int nccl_init_rank(void **comm, const unsigned char id[128], int rank, int nranks) {
  Api &a = api();
  if (!a.ok) return kNoLibrary;
  ncclUniqueId u;
  std::memcpy(&u, id, 128);
  ncclComm_t c = nullptr;
  const ncclResult_t r = a.CommInitRank(&c, nranks, u, rank);
  if (r != ncclSuccess) return (int)r;
  *comm = (void *)c;
  return 0;
}

int nccl_all_reduce_sum(void *comm, double *buf, size_t count, cudaStream_t st) {
  Api &a = api();
  if (!a.ok) return kNoLibrary;
  return (int)a.AllReduce(buf, buf, count, ncclDouble, ncclSum, (ncclComm_t)comm, st);
}

void nccl_destroy(void *comm) {
  Api &a = api();
  if (a.ok && comm) a.CommDestroy((ncclComm_t)comm);
}

const char *nccl_error_string(int rc) {
  if (rc == kNoLibrary) return "libnccl.so.2 not found (dlopen)";
  Api &a = api();
  return a.GetErrorString ? a.GetErrorString((ncclResult_t)rc) : "NCCL error";
}

}  // namespace uvs
