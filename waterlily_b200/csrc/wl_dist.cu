// wl_dist.cu — z-slab domain decomposition: one process per GPU, NCCL over NVLink/NVSwitch.
//
// The reference has no multi-GPU path (README.md:157 lists it as a goal); this is new work
// (SURVEY.md §8e).  Each rank owns Nz/R interior z-planes of every field plus one ghost plane per
// side (two extra for u/u⁰: QUICK reaches I-2δ…I+2δ).  z is the slowest index, so a plane is one
// contiguous run of floats and a halo exchange is 2 sends + 2 receives of contiguous memory,
// issued at exactly the sites where the reference calls BC!/perBC!.  Global reductions are
// ncclAllReduce on 1-2 Float64 partials that stay on the device; every rank then takes the same
// data-dependent decision from the same bits.
//
// NCCL is loaded with dlopen so that a single-GPU build has no link-time dependency, and so that
// inside a torch process the library torch already loaded (same soname) is reused.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <map>

#include "wl_internal.h"

namespace {
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
} N;

bool load_nccl(std::string& err) {
  if (N.h) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    N.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (N.h) break;
  }
  if (!N.h) { err = std::string("cannot load NCCL: ") + dlerror(); return false; }
#define SYM(field, name) \
  *(void**)(&N.field) = dlsym(N.h, name); \
  if (!N.field) { err = std::string("NCCL symbol missing: ") + name; N.h = nullptr; return false; }
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
  SYM(GetErrorString, "ncclGetErrorString") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
  SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather")
#undef SYM
  return true;
}

int nccl_check(Sim& s, ncclResult_t r, const char* what) {
  if (r == ncclSuccess) return 0;
  if (!s.sticky) s.sticky = WL_ERR_NCCL;
  return wl_fail(WL_ERR_NCCL, std::string("NCCL error in ") + what + ": " + (N.GetErrorString ? N.GetErrorString(r) : "?"));
}
#define WL_NCCL(s, call) do { int _rc = nccl_check((s), (call), #call); if (_rc) return _rc; } while (0)
}  // namespace

int dist_unique_id(char* out128) {
  std::string err;
  if (!load_nccl(err)) return wl_fail(WL_ERR_NCCL, err);
  ncclUniqueId id;
  ncclResult_t r = N.GetUniqueId(&id);
  if (r != ncclSuccess) return wl_fail(WL_ERR_NCCL, std::string("ncclGetUniqueId: ") + N.GetErrorString(r));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, 128);
  return 0;
}

int dist_init(Sim& s, int rank, int nranks, const char* id128) {
  s.dist.rank = rank;
  s.dist.nranks = nranks;
  if (nranks == 1) return 0;
  std::string err;
  if (!load_nccl(err)) return wl_fail(WL_ERR_NCCL, err);
  // a unique id creates exactly one communicator: handles made with the same id share it
  // (it lives until the process exits)
  static std::map<std::string, ncclComm_t> comms;
  const std::string key(id128, 128);
  auto it = comms.find(key);
  if (it == comms.end()) {
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    WL_NCCL(s, N.CommInitRank(&comm, nranks, id, rank));
    it = comms.emplace(key, comm).first;
  }
  s.dist.comm = (void*)it->second;
  return 0;
}

void dist_destroy(Sim& s) { s.dist.comm = nullptr; }

// Exchange `depth` z-planes per side of `ncomp` components of a field living on grid g.
// base points at cell (0,0,0) of component 0; plane k starts at base + k*s2.
// lower ghost planes [1-depth … 0] ← lower neighbour's top interior planes [n-1-depth … n-2]
// upper ghost planes [n-1 … n-2+depth] ← upper neighbour's bottom interior planes [1 … depth]
int dist_halo(Sim& s, const GridD& g, float* base, int ncomp, int depth) {
  if (s.dist.nranks == 1 || (!g.zlo_halo && !g.zhi_halo)) return 0;
  ncclComm_t comm = (ncclComm_t)s.dist.comm;
  const int R = s.dist.nranks, me = s.dist.rank;
  const int up = (me + 1) % R, dn = (me + R - 1) % R;
  const size_t cnt = (size_t)depth * g.s[2];
  const int n = g.n[2];
  WL_NCCL(s, N.GroupStart());
  for (int c = 0; c < ncomp; ++c) {
    float* b = base + c * g.comp;
    // order send_up, recv_dn, send_dn, recv_up: pairs match even when up == dn (R == 2, periodic)
    if (g.zhi_halo) WL_NCCL(s, N.Send(b + (int64_t)(n - 1 - depth) * g.s[2], cnt, ncclFloat, up, comm, s.stream));
    if (g.zlo_halo) WL_NCCL(s, N.Recv(b + (int64_t)(1 - depth) * g.s[2], cnt, ncclFloat, dn, comm, s.stream));
    if (g.zlo_halo) WL_NCCL(s, N.Send(b + (int64_t)1 * g.s[2], cnt, ncclFloat, dn, comm, s.stream));
    if (g.zhi_halo) WL_NCCL(s, N.Recv(b + (int64_t)(n - 1) * g.s[2], cnt, ncclFloat, up, comm, s.stream));
  }
  WL_NCCL(s, N.GroupEnd());
  ++s.collectives;
  return 0;
}

int dist_allreduce(Sim& s, double* dptr, int count, bool is_max) {
  if (s.dist.nranks == 1) return 0;
  WL_NCCL(s, N.AllReduce(dptr, dptr, (size_t)count, ncclDouble, is_max ? ncclMax : ncclSum, (ncclComm_t)s.dist.comm, s.stream));
  ++s.collectives;
  return 0;
}

// in-place all-gather of per-rank chunks of `count` floats laid out rank-major at base
int dist_allgather(Sim& s, float* base, size_t count) {
  if (s.dist.nranks == 1) return 0;
  WL_NCCL(s, N.AllGather(base + (size_t)s.dist.rank * count, base, count, ncclFloat, (ncclComm_t)s.dist.comm, s.stream));
  ++s.collectives;
  return 0;
}
