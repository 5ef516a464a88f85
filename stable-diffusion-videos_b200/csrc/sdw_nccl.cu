// sdw_nccl.cu — the three exchanges of the frame-sharded walk as thin NCCL wrappers behind the C ABI:
// communicator from a broadcast ncclUniqueId, ONE broadcast of the packed weights, decoded uint8 frames sent to rank 0.
// The reference's only multi-device precedent is the Flax twin's pmap over the frame axis
// (flax_stable_diffusion_pipeline.py:546, 568-578, 594-597, 898-902, 935: replicate params, shard, unshard).
//
// NCCL is bound at run time (dlopen of the libnccl the process already carries — torch's bundled one — or
// $SDW_NCCL_LIB), so libsdwalk.so keeps linking libcudart only.  No collective exists inside the sampler: frames are
// independent, which is why there is no fused compute + collective kernel here.
#include "sdw_internal.h"
#include "../../include/sdwalk.h"

#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace sdw {
namespace {

struct NcclId { char internal[128]; };  // ncclUniqueId
using comm_t = void*;
constexpr int kUint8 = 1;               // ncclUint8

struct NcclApi {
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(comm_t*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(comm_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    if (const char* e = std::getenv("SDW_NCCL_LIB")) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy torch has already loaded
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    auto sym = [&](const char* n) { return dlsym(h, n); };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
    a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
    a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Broadcast && a.Send && a.Recv && a.GroupStart &&
           a.GroupEnd && a.GetErrorString;
  });
  return a;
}

int need_api() {
  if (api().ok) return 0;
  set_error("NCCL is not available: libnccl.so.2 could not be loaded (set SDW_NCCL_LIB to its path)");
  return 1;
}

#define SDW_NCCL_OK(expr)                                                                  \
  do {                                                                                      \
    const int _r = (expr);                                                                  \
    if (_r != 0) {                                                                          \
      ::sdw::set_error(std::string(#expr) + ": " + api().GetErrorString(_r));              \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

}  // namespace
}  // namespace sdw

using namespace sdw;

struct sdw_comm {
  comm_t comm = nullptr;
  int rank = 0, world = 1;
};

extern "C" {

int sdw_nccl_unique_id(void* id128) {
  SDW_REQUIRE(id128 != nullptr, "null id");
  if (int e = need_api()) return e;
  NcclId id;
  SDW_NCCL_OK(api().GetUniqueId(&id));
  std::memcpy(id128, &id, sizeof id);
  return 0;
}

int sdw_nccl_init(const void* id128, int rank, int world, sdw_comm** out) {
  SDW_REQUIRE(id128 != nullptr && out != nullptr, "null argument");
  SDW_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank outside [0, world)");
  if (int e = need_api()) return e;
  NcclId id;
  std::memcpy(&id, id128, sizeof id);
  sdw_comm* c = new sdw_comm;
  c->rank = rank;
  c->world = world;
  const int r = api().CommInitRank(&c->comm, world, id, rank);  // uses the calling thread's current CUDA device
  if (r != 0) {
    set_error(std::string("ncclCommInitRank: ") + api().GetErrorString(r));
    delete c;
    return 1;
  }
  *out = c;
  return 0;
}

void sdw_nccl_destroy(sdw_comm* c) {
  if (!c) return;
  if (c->comm && api().ok) api().CommDestroy(c->comm);
  delete c;
}

// in place on every rank: rank `root` holds the data, the others receive it (the flat fp16 weight buffer, once)
int sdw_nccl_broadcast_weights(sdw_comm* c, void* buf, uint64_t bytes, int root, void* stream) {
  SDW_REQUIRE(c != nullptr && buf != nullptr, "null argument");
  SDW_REQUIRE(root >= 0 && root < c->world, "root outside the communicator");
  if (bytes == 0) return 0;
  SDW_NCCL_OK(api().Broadcast(buf, buf, static_cast<size_t>(bytes), kUint8, root, c->comm, static_cast<cudaStream_t>(stream)));
  return 0;
}

// every rank contributes `bytes_per_rank` bytes (its padded frame block); only `root` receives: recv (root only) is
// [world][bytes_per_rank].  One grouped send / recv round — the bytes on the wire are the frames themselves
int sdw_nccl_gather_frames(sdw_comm* c, const void* send, void* recv, uint64_t bytes_per_rank, int root, void* stream) {
  SDW_REQUIRE(c != nullptr && send != nullptr, "null argument");
  SDW_REQUIRE(root >= 0 && root < c->world, "root outside the communicator");
  SDW_REQUIRE(c->rank != root || recv != nullptr, "the root needs a receive buffer");
  if (bytes_per_rank == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = static_cast<size_t>(bytes_per_rank);
  SDW_NCCL_OK(api().GroupStart());
  if (c->rank == root) {
    for (int r = 0; r < c->world; ++r) {
      uint8_t* dst = static_cast<uint8_t*>(recv) + static_cast<size_t>(r) * n;
      if (r == root) {
        if (cudaMemcpyAsync(dst, send, n, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
          api().GroupEnd();
          set_error("cudaMemcpyAsync of the root's own block failed");
          return 1;
        }
      } else {
        const int rr = api().Recv(dst, n, kUint8, r, c->comm, st);
        if (rr != 0) {
          api().GroupEnd();
          set_error(std::string("ncclRecv: ") + api().GetErrorString(rr));
          return 1;
        }
      }
    }
  } else {
    const int rs = api().Send(send, n, kUint8, root, c->comm, st);
    if (rs != 0) {
      api().GroupEnd();
      set_error(std::string("ncclSend: ") + api().GetErrorString(rs));
      return 1;
    }
  }
  SDW_NCCL_OK(api().GroupEnd());
  return 0;
}

}  // extern "C"
