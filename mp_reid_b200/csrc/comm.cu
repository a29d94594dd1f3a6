// C-ABI multi-GPU surface (mpreid_comm_{init,broadcast,allgather,destroy}): thin wrappers over NCCL for a
// consumer that is not PyTorch.  The path needs exactly three collectives: the gallery broadcast, the
// all-gather of per-query results / neighbour lists, and a max all-reduce of the row maxima of the sharded all-pairs pass.
// NCCL is resolved at run time with dlopen("libnccl.so.2"): inside a PyTorch process that is the copy torch already
// loaded (same SONAME), elsewhere the system library; the library itself has no link-time dependency on NCCL.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace mpreid {

struct NcclId { char internal[128]; };
typedef int (*fn_get_unique_id)(NcclId*);
typedef int (*fn_comm_init_rank)(void**, int, NcclId, int);
typedef int (*fn_comm_destroy)(void*);
typedef int (*fn_broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*fn_error_string)(int);

struct NcclApi {
  void* handle;
  fn_get_unique_id get_unique_id; fn_comm_init_rank comm_init_rank; fn_comm_destroy comm_destroy;
  fn_broadcast broadcast; fn_all_gather all_gather; fn_all_reduce all_reduce; fn_error_string error_string;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static int state = 0;   // 0 = not tried, 1 = ok, -1 = unavailable
  if (state == 0) {
    const char* names[] = {getenv("MPREID_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) if (n && !h) h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (h) {
      api.handle = h;
      api.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
      api.comm_init_rank = (fn_comm_init_rank)dlsym(h, "ncclCommInitRank");
      api.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
      api.broadcast = (fn_broadcast)dlsym(h, "ncclBroadcast");
      api.all_gather = (fn_all_gather)dlsym(h, "ncclAllGather");
      api.all_reduce = (fn_all_reduce)dlsym(h, "ncclAllReduce");
      api.error_string = (fn_error_string)dlsym(h, "ncclGetErrorString");
    }
    state = (h && api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.broadcast && api.all_gather && api.all_reduce) ? 1 : -1;
  }
  return state == 1 ? &api : nullptr;
}

}  // namespace mpreid

using namespace mpreid;

struct mpreid_comm {
  void* nccl;      // ncclComm_t
  int world, rank;
  int owned;       // created by mpreid_comm_init (destroyed with the handle) or adopted from the caller
};

#define MPREID_NCCL_CHECK(api, expr)                                                                       \
  do {                                                                                                     \
    int _r = (expr);                                                                                       \
    if (_r != 0) {                                                                                         \
      set_error("%s failed: %s", #expr, (api)->error_string ? (api)->error_string(_r) : "NCCL error");    \
      return MPREID_ERR_CUDA;                                                                              \
    }                                                                                                      \
  } while (0)

extern "C" int mpreid_comm_unique_id(void* id_out_128_bytes) {
  NcclApi* api = nccl_api();
  if (!api) { set_error("comm: libnccl.so.2 could not be loaded (set MPREID_NCCL_LIB)"); return MPREID_ERR_UNSUPPORTED; }
  MPREID_REQUIRE(id_out_128_bytes, "comm_unique_id: null pointer");
  MPREID_NCCL_CHECK(api, api->get_unique_id((NcclId*)id_out_128_bytes));
  return MPREID_OK;
}

extern "C" int mpreid_comm_init(mpreid_comm** comm, int world, int rank, const void* unique_id_128_bytes) {
  NcclApi* api = nccl_api();
  if (!api) { set_error("comm: libnccl.so.2 could not be loaded (set MPREID_NCCL_LIB)"); return MPREID_ERR_UNSUPPORTED; }
  MPREID_REQUIRE(comm && unique_id_128_bytes && world >= 1 && rank >= 0 && rank < world, "comm_init: bad arguments (rank %d of %d)", rank, world);
  NcclId id;
  memcpy(&id, unique_id_128_bytes, sizeof(id));
  void* nc = nullptr;
  MPREID_NCCL_CHECK(api, api->comm_init_rank(&nc, world, id, rank));   // binds to the calling thread's current CUDA device
  mpreid_comm* c = (mpreid_comm*)malloc(sizeof(mpreid_comm));
  if (!c) { api->comm_destroy(nc); set_error("comm_init: out of memory"); return MPREID_ERR_INVALID; }
  c->nccl = nc; c->world = world; c->rank = rank; c->owned = 1;
  *comm = c;
  return MPREID_OK;
}

extern "C" int mpreid_comm_from_nccl(mpreid_comm** comm, void* nccl_comm, int world, int rank) {
  NcclApi* api = nccl_api();
  if (!api) { set_error("comm: libnccl.so.2 could not be loaded (set MPREID_NCCL_LIB)"); return MPREID_ERR_UNSUPPORTED; }
  MPREID_REQUIRE(comm && nccl_comm && world >= 1 && rank >= 0 && rank < world, "comm_from_nccl: bad arguments");
  mpreid_comm* c = (mpreid_comm*)malloc(sizeof(mpreid_comm));
  if (!c) { set_error("comm_from_nccl: out of memory"); return MPREID_ERR_INVALID; }
  c->nccl = nccl_comm; c->world = world; c->rank = rank; c->owned = 0;
  *comm = c;
  return MPREID_OK;
}

extern "C" int mpreid_comm_size(const mpreid_comm* comm, int* world, int* rank) {
  MPREID_REQUIRE(comm, "comm_size: null communicator");
  if (world) *world = comm->world;
  if (rank) *rank = comm->rank;
  return MPREID_OK;
}

// gallery features / labels from `root` to every rank, in place (utils/metrics.py has no such step: the reference
// evaluates on rank 0 only, processor/processor.py:117-118)
extern "C" int mpreid_comm_broadcast(mpreid_comm* comm, void* buf, size_t bytes, int root, void* stream) {
  NcclApi* api = nccl_api();
  MPREID_REQUIRE(api && comm && buf && root >= 0 && root < comm->world, "comm_broadcast: bad arguments");
  MPREID_NCCL_CHECK(api, api->broadcast(buf, buf, bytes, /*ncclUint8*/ 1, root, comm->nccl, (cudaStream_t)stream));
  return MPREID_OK;
}

// recv [world, bytes_per_rank]: per-query (AP, first_hit, num_rel) buffers, neighbour lists, partial top-k keys, V0 rows
extern "C" int mpreid_comm_allgather(mpreid_comm* comm, const void* send, void* recv, size_t bytes_per_rank, void* stream) {
  NcclApi* api = nccl_api();
  MPREID_REQUIRE(api && comm && send && recv, "comm_allgather: bad arguments");
  MPREID_NCCL_CHECK(api, api->all_gather(send, recv, bytes_per_rank, /*ncclUint8*/ 1, comm->nccl, (cudaStream_t)stream));
  return MPREID_OK;
}

// in-place max over the ranks (row maxima of the row-sharded all-pairs pass, utils/reranking.py:46)
extern "C" int mpreid_comm_allreduce_max_f32(mpreid_comm* comm, float* buf, size_t count, void* stream) {
  NcclApi* api = nccl_api();
  MPREID_REQUIRE(api && comm && buf, "comm_allreduce_max_f32: bad arguments");
  MPREID_NCCL_CHECK(api, api->all_reduce(buf, buf, count, /*ncclFloat32*/ 7, /*ncclMax*/ 2, comm->nccl, (cudaStream_t)stream));
  return MPREID_OK;
}

extern "C" int mpreid_comm_destroy(mpreid_comm* comm) {
  if (!comm) return MPREID_OK;
  NcclApi* api = nccl_api();
  int rc = MPREID_OK;
  if (comm->owned && api && api->comm_destroy(comm->nccl) != 0) { set_error("ncclCommDestroy failed"); rc = MPREID_ERR_CUDA; }
  free(comm);
  return rc;
}
