// In-library exchange for one proof on several GPUs: one process per GPU, an NCCL communicator owned by the
// context, every collective enqueued on the context's own stream.  What crosses NVLink is tiny - one partial POINT per MSM
// and GPU (72 bytes with its infinity flag), at most 17 partial field sums per sharded sumcheck round - so the collective is a
// plain ncclAllGather of bytes followed by the (exact, order-independent) group / field additions on the host; nothing goes
// through torch, numpy or a caller callback.  The reference has no distributed code to mirror (single-process rayon).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy the process already loaded - torch bundles one - or the system
// library), so the library still loads where NCCL is absent; ja_comm_* then fail with JA_ERR_UNSUPPORTED.
// Rendezvous: rank 0 calls ja_comm_unique_id and hands the 128 bytes to the other ranks by any means (bench.py / parallel.py
// broadcast them through torch.distributed's store; a Rust host would use its own launcher), then every rank calls ja_comm_init.
#include <dlfcn.h>

#include "common.hpp"

namespace {

struct NcclUid { char internal[128]; };
typedef int (*GetUidFn)(NcclUid*);
typedef int (*InitRankFn)(void**, int, NcclUid, int);
typedef int (*AllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*DestroyFn)(void*);
typedef const char* (*ErrStrFn)(int);
struct Nccl {
  void* lib = nullptr;
  GetUidFn get_uid = nullptr; InitRankFn init_rank = nullptr; AllGatherFn all_gather = nullptr; DestroyFn destroy = nullptr;
  ErrStrFn err_str = nullptr;
  bool ok() const { return get_uid && init_rank && all_gather && destroy; }
};
Nccl& nccl() {
  static Nccl n = [] {
    Nccl r;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      r.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (r.lib) break;
    }
    if (!r.lib) return r;
    r.get_uid = (GetUidFn)dlsym(r.lib, "ncclGetUniqueId");
    r.init_rank = (InitRankFn)dlsym(r.lib, "ncclCommInitRank");
    r.all_gather = (AllGatherFn)dlsym(r.lib, "ncclAllGather");
    r.destroy = (DestroyFn)dlsym(r.lib, "ncclCommDestroy");
    r.err_str = (ErrStrFn)dlsym(r.lib, "ncclGetErrorString");
    return r;
  }();
  return n;
}
int32_t nccl_fail(const char* what, int rc) {
  Nccl& n = nccl();
  return fail(JA_ERR_CUDA, std::string(what) + ": " + (n.err_str ? n.err_str(rc) : "NCCL error") + " (" + std::to_string(rc) + ")");
}

}  // namespace

// all-gather of `bytes` from every rank of the context's communicator: host buffers in and out, rank-major; NCCL on the context stream
int32_t comm_allgather(ja_ctx* c, const void* send, size_t bytes, void* recv) {
  JA_REQUIRE(c && c->comm && send && recv && bytes, "comm_allgather: no communicator / null argument");
  Nccl& n = nccl();
  const size_t world = c->comm_world;
  char* d = nullptr;
  int32_t st = dev_alloc(c, bytes * (world + 1), (void**)&d);
  if (st) return st;
  cudaError_t e = cudaMemcpyAsync(d, send, bytes, cudaMemcpyHostToDevice, c->stream);
  int rc = 0;
  if (e == cudaSuccess) rc = n.all_gather(d, d + bytes, bytes, 1 /* ncclUint8 */, c->comm, c->stream);
  if (e == cudaSuccess && rc == 0) e = cudaMemcpyAsync(recv, d + bytes, bytes * world, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess && rc == 0) e = cudaStreamSynchronize(c->stream);
  dev_free(c, d);
  if (rc) return nccl_fail("ncclAllGather", rc);
  if (e != cudaSuccess) return fail(JA_ERR_CUDA, std::string("comm_allgather: ") + cudaGetErrorString(e));
  return JA_OK;
}

// partial affine points (xy: count x 8 limbs, inf: count flags) of every rank -> their sums, identical on every rank
int32_t comm_combine_points(ja_ctx* c, uint64_t* xy, int32_t* inf, size_t count) {
  const size_t world = c->comm_world;
  std::vector<uint64_t> mine(count * 9), all(count * 9 * world);
  for (size_t i = 0; i < count; i++) { memcpy(&mine[9 * i], xy + 8 * i, 64); mine[9 * i + 8] = (uint64_t)(inf ? inf[i] : 0); }
  int32_t st = comm_allgather(c, mine.data(), mine.size() * 8, all.data());
  if (st) return st;
  std::vector<uint64_t> pts(8 * world);
  std::vector<int32_t> flags(world);
  for (size_t i = 0; i < count; i++) {
    for (size_t r = 0; r < world; r++) { memcpy(&pts[8 * r], &all[(r * count + i) * 9], 64); flags[r] = (int32_t)all[(r * count + i) * 9 + 8]; }
    int32_t oi = 0;
    if ((st = ja_g1_sum_affine(pts.data(), flags.data(), world, xy + 8 * i, &oi))) return st;
    if (inf) inf[i] = oi;
  }
  return JA_OK;
}

extern "C" {

int32_t ja_comm_unique_id(uint8_t out[128]) {
  JA_REQUIRE(out, "ja_comm_unique_id: null argument");
  Nccl& n = nccl();
  if (!n.ok()) return fail(JA_ERR_UNSUPPORTED, "ja_comm_unique_id: libnccl.so.2 is not available in this process");
  NcclUid id;
  const int rc = n.get_uid(&id);
  if (rc) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(out, id.internal, 128);
  return JA_OK;
}

int32_t ja_comm_init(ja_ctx* c, uint32_t rank, uint32_t world, const uint8_t id[128]) {
  JA_REQUIRE(c && id && world >= 1 && rank < world, "ja_comm_init: bad argument");
  Nccl& n = nccl();
  if (!n.ok()) return fail(JA_ERR_UNSUPPORTED, "ja_comm_init: libnccl.so.2 is not available in this process");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  if (c->comm) { n.destroy(c->comm); c->comm = nullptr; }
  NcclUid uid;
  memcpy(uid.internal, id, 128);
  void* comm = nullptr;
  const int rc = n.init_rank(&comm, (int)world, uid, (int)rank);
  if (rc) return nccl_fail("ncclCommInitRank", rc);
  c->comm = comm; c->comm_rank = rank; c->comm_world = world;
  return JA_OK;
}

void ja_comm_free(ja_ctx* c) {
  if (!c || !c->comm) return;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  nccl().destroy(c->comm);
  c->comm = nullptr; c->comm_world = 1; c->comm_rank = 0;
}

int32_t ja_comm_allgather(ja_ctx* c, const void* send, size_t bytes, void* recv) {
  JA_REQUIRE(c, "ja_comm_allgather: null context");
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  JA_CUDA(cudaSetDevice(c->device));
  return comm_allgather(c, send, bytes, recv);
}

}  // extern "C"
