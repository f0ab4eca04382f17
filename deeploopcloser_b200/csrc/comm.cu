// NCCL plumbing of the row-sharded matcher behind the C ABI (dlc_match_topk_sharded): one communicator per process,
// created from a unique id the host distributes (any out-of-band channel: the Python host uses torch.distributed).
// libnccl is resolved at run time with dlopen - the library the host process already loaded (PyTorch's bundled NCCL)
// is reused; libdlc.so itself has no link-time dependency on it, so single-GPU users never need NCCL.
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "util.h"

namespace dlc {

struct NcclId {
  char internal[128];
};
typedef void* NcclComm;
struct NcclApi {
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
  });
  return api.ok ? &api : nullptr;
}

}  // namespace dlc

struct dlc_comm {
  dlc::NcclComm comm = nullptr;
  int rank = 0, world = 1;
};

using namespace dlc;

namespace dlc {
// all-gather of `bytes` bytes per rank on `stream` (stream-ordered like a kernel launch; no host synchronisation)
int comm_all_gather(dlc_comm* c, const void* send, void* recv, size_t bytes, cudaStream_t stream) {
  NcclApi* api = nccl_api();
  if (!api || !c || !c->comm) return fail(DLC_EUNSUPPORTED, "dlc_comm: NCCL is not available");
  const int rc = api->AllGather(send, recv, bytes, /*ncclChar*/ 0, c->comm, stream);
  if (rc != 0) return fail(DLC_ECUDA, "dlc_comm: ncclAllGather failed: %s", api->GetErrorString(rc));
  return DLC_OK;
}
int comm_rank(const dlc_comm* c) { return c ? c->rank : 0; }
int comm_world(const dlc_comm* c) { return c ? c->world : 1; }
}  // namespace dlc

extern "C" int dlc_comm_unique_id(void* id_out_host) {
  DLC_CHECK_ARG(id_out_host);
  NcclApi* api = nccl_api();
  if (!api) return fail(DLC_EUNSUPPORTED, "dlc_comm_unique_id: libnccl.so.2 not found (dlopen)");
  const int rc = api->GetUniqueId(static_cast<NcclId*>(id_out_host));
  if (rc != 0) return fail(DLC_ECUDA, "dlc_comm_unique_id: %s", api->GetErrorString(rc));
  return DLC_OK;
}

extern "C" int dlc_comm_create(dlc_comm** c, const void* id_host, int rank, int world) {
  DLC_CHECK_ARG(c && id_host);
  DLC_CHECK_ARG(world >= 1 && rank >= 0 && rank < world);
  NcclApi* api = nccl_api();
  if (!api) return fail(DLC_EUNSUPPORTED, "dlc_comm_create: libnccl.so.2 not found (dlopen)");
  dlc_comm* out = new dlc_comm();
  out->rank = rank;
  out->world = world;
  NcclId id;
  memcpy(&id, id_host, sizeof(id));
  const int rc = api->CommInitRank(&out->comm, world, id, rank);
  if (rc != 0) {
    delete out;
    return fail(DLC_ECUDA, "dlc_comm_create: ncclCommInitRank failed: %s", api->GetErrorString(rc));
  }
  *c = out;
  return DLC_OK;
}

extern "C" int dlc_comm_destroy(dlc_comm* c) {
  if (!c) return DLC_OK;
  NcclApi* api = nccl_api();
  if (api && c->comm) api->CommDestroy(c->comm);
  delete c;
  return DLC_OK;
}
