// wl_dist.cu — z-slab domain decomposition: one process per GPU of an NVLink/NVSwitch box.
//
// The reference has no multi-GPU path (README.md:157 lists it as a goal); this is new work
// (SURVEY.md §8e).  Each rank owns Nz/R interior z-planes of every field plus one ghost plane per
// side (two extra for u/u⁰: QUICK reaches I-2δ…I+2δ).  z is the slowest index, so a plane is one
// contiguous run of floats and a halo exchange moves contiguous memory, at exactly the sites where
// the reference calls BC!/perBC!.  Global reductions combine 1-2 Float64 partials that stay on the
// device; every rank then takes the same data-dependent decision from the same bits.
//
// Two data paths.  Default: peer memory (wl_p2p.cuh) — the ranks map each other's field arena over
// CUDA IPC and kernels store halo planes / reduction partials straight into the neighbours' memory
// (k_p2p_push here, p2p_allreduce2 in the producing kernels, push_segments inside the persistent
// solver kernel).  Fallback (WL_P2P=0 or a failed mapping): grouped ncclSend/ncclRecv and
// ncclAllReduce + a finalise kernel.  NCCL also bootstraps the peer path (IPC handle exchange).
//
// NCCL is loaded with dlopen so that a single-GPU build has no link-time dependency, and so that
// inside a torch process the library torch already loaded (same soname) is reused.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

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

struct CommRef { ncclComm_t comm; int refs; };
std::mutex g_comm_mu;                      // wl_create_dist may be called from several host threads
std::map<std::string, CommRef> g_comms;   // NCCL unique id → communicator
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
  // a unique id creates exactly one communicator (NCCL ids are single-use): handles made with the
  // same id share it, and it lives until wl_dist_finalize() or process exit
  std::lock_guard<std::mutex> lock(g_comm_mu);
  const std::string key(id128, 128);
  auto it = g_comms.find(key);
  if (it == g_comms.end()) {
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    WL_NCCL(s, N.CommInitRank(&comm, nranks, id, rank));
    it = g_comms.emplace(key, CommRef{comm, 0}).first;
  }
  ++it->second.refs;
  s.dist.comm = (void*)it->second.comm;
  return 0;
}

// ---- peer-memory path -----------------------------------------------------------------------
namespace {
// One push = a list of contiguous float4 runs copied from this rank's arena into the same field of a
// peer's arena, bracketed by a two-phase epoch handshake with every peer involved:
//   1. tell the peers "I have reached epoch e" (all my earlier kernels are complete, so my ghost planes
//      may be overwritten) and wait for the same from them;
//   2. store the runs into the peers' arenas (NVLink), fence;
//   3. the last block tells the peers "my data of epoch e is visible" and waits for theirs.
// When the kernel retires, the ghost planes of this rank hold the neighbours' epoch-e data.
struct PushSeg { int peer; int pad; int64_t src_off, dst_off, n4; };  // byte offsets into the arenas, float4 count
#define WL_PUSH_MAXSEG 24
struct PushPlan {
  int nseg, npeers;
  int peers[WL_MAXRANKS];
  PushSeg seg[WL_PUSH_MAXSEG];
};
#define WL_PUSH_CHUNK 2048  // float4 per block-chunk (32 KiB)

__global__ void __launch_bounds__(256) k_p2p_push(const P2PDev* __restrict__ ctxp, PushPlan plan) {
  __shared__ P2PDev c;
  __shared__ bool last;
  if (threadIdx.x == 0) c = *ctxp;
  __syncthreads();
  P2PFlags* mine = p2p_flags(c, c.rank);
  // the epoch counter lives on the device (the fused V-cycle kernel pushes without the host); every
  // block samples it before taking its ticket, the last block advances it
  const unsigned long long epoch = *((const volatile unsigned long long*)&mine->push_epoch) + 1;
  const int t = threadIdx.x;
  if (t < plan.npeers) {
    const int pr = plan.peers[t];
    if (blockIdx.x == 0) p2p_st_release(&p2p_flags(c, pr)->ready[c.rank], epoch);
    p2p_wait<false>(c, &mine->ready[pr], epoch);
  }
  __syncthreads();
  // chunks of all segments, grid-strided
  int64_t first = 0;
  for (int sgi = 0; sgi < plan.nseg; ++sgi) {
    const PushSeg sg = plan.seg[sgi];
    const int64_t nchunk = (sg.n4 + WL_PUSH_CHUNK - 1) / WL_PUSH_CHUNK;
    const float4* __restrict__ src = reinterpret_cast<const float4*>(c.base[c.rank] + sg.src_off);
    float4* __restrict__ dst = reinterpret_cast<float4*>(c.base[sg.peer] + sg.dst_off);
    // first chunk of this segment handled by this block: chunk ids are global (first + k)
    int64_t k = ((int64_t)blockIdx.x - first % gridDim.x + gridDim.x) % gridDim.x;
    for (; k < nchunk; k += gridDim.x) {
      const int64_t lo = k * WL_PUSH_CHUNK, hi = min(lo + (int64_t)WL_PUSH_CHUNK, sg.n4);
      for (int64_t i = lo + t; i < hi; i += blockDim.x) dst[i] = src[i];
    }
    first += nchunk;
  }
  __threadfence_system();
  __syncthreads();
  if (t == 0) {
    const unsigned int tk = atomicAdd(&mine->ticket, 1u);
    last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  if (t < plan.npeers) {
    const int pr = plan.peers[t];
    p2p_st_release(&p2p_flags(c, pr)->done[c.rank], epoch);
    p2p_wait<false>(c, &mine->done[pr], epoch);
  }
  if (t == 0) { mine->ticket = 0u; mine->push_epoch = epoch; }
}

int push(Sim& s, PushPlan& plan, int64_t total4) {
  int64_t nchunk = (total4 + WL_PUSH_CHUNK - 1) / WL_PUSH_CHUNK + plan.nseg;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(nchunk, 2 * s.nsm));
  k_p2p_push<<<grid, 256, 0, s.stream>>>(s.dist.dev, plan);
  ++s.launches;
  ++s.dist.pushes;
  return wl_cuda_check(s, cudaGetLastError(), "k_p2p_push");
}
void plan_add_peer(PushPlan& p, int r) {
  for (int i = 0; i < p.npeers; ++i) if (p.peers[i] == r) return;
  p.peers[p.npeers++] = r;
}
}  // namespace

// Map the arenas of all ranks into this process.  The 64-byte IPC handles travel through NCCL (the
// communicator exists anyway for bootstrap); any failure leaves the NCCL data path in place.
int dist_p2p_setup(Sim& s) {
  if (s.dist.nranks == 1) return 0;
  const char* env = getenv("WL_P2P");
  if (env && atoi(env) == 0) return 0;
  const int R = s.dist.nranks, me = s.dist.rank;
  if (R > WL_MAXRANKS) return 0;
  ncclComm_t comm = (ncclComm_t)s.dist.comm;
  struct Rec { cudaIpcMemHandle_t h; int ok; int pad[15]; };
  static_assert(sizeof(Rec) == 128, "IPC record");
  Rec mine{};
  mine.ok = cudaIpcGetMemHandle(&mine.h, s.arena) == cudaSuccess ? 1 : 0;
  if (!mine.ok) cudaGetLastError();
  Rec* d_recs = nullptr;
  WL_CUDA(s, cudaMalloc((void**)&d_recs, sizeof(Rec) * R));
  s.allocs.push_back(d_recs);
  WL_CUDA(s, cudaMemcpyAsync(d_recs + me, &mine, sizeof(Rec), cudaMemcpyHostToDevice, s.stream));
  WL_NCCL(s, N.AllGather(d_recs + me, d_recs, sizeof(Rec), ncclChar, comm, s.stream));
  std::vector<Rec> recs(R);
  WL_CUDA(s, cudaMemcpyAsync(recs.data(), d_recs, sizeof(Rec) * R, cudaMemcpyDeviceToHost, s.stream));
  WL_CUDA(s, cudaStreamSynchronize(s.stream));
  bool all_ok = true;
  for (int r = 0; r < R; ++r) all_ok &= recs[r].ok != 0;
  P2PDev h{};
  h.rank = me; h.nranks = R;
  const char* to = getenv("WL_P2P_TIMEOUT_MS");
  h.timeout_ns = (unsigned long long)(to ? atoll(to) : 20000) * 1000000ull;
  int mapped = 1;
  if (all_ok) {
    for (int r = 0; r < R; ++r) {
      if (r == me) { h.base[r] = s.arena; continue; }
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, recs[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mapped = 0; break; }
      h.base[r] = (char*)ptr;
    }
  } else {
    mapped = 0;
  }
  // everybody must agree: one failed mapping anywhere keeps all ranks on NCCL
  int* d_flag = (int*)d_recs;  // reuse (the records were consumed)
  WL_CUDA(s, cudaMemcpyAsync(d_flag, &mapped, sizeof(int), cudaMemcpyHostToDevice, s.stream));
  WL_NCCL(s, N.AllReduce(d_flag, d_flag, 1, ncclInt, ncclMin, comm, s.stream));
  int agreed = 0;
  WL_CUDA(s, cudaMemcpyAsync(&agreed, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  WL_CUDA(s, cudaStreamSynchronize(s.stream));
  if (!agreed) {
    for (int r = 0; r < R; ++r) if (r != me && h.base[r]) cudaIpcCloseMemHandle(h.base[r]);
    if (getenv("WL_P2P_VERBOSE")) fprintf(stderr, "[waterlily_b200] rank %d: peer mapping unavailable, using NCCL\n", me);
    return 0;
  }
  s.dist.host = h;
  WL_CUDA(s, cudaMalloc((void**)&s.dist.dev, sizeof(P2PDev)));
  s.allocs.push_back(s.dist.dev);
  WL_CUDA(s, cudaMemcpyAsync(s.dist.dev, &h, sizeof(P2PDev), cudaMemcpyHostToDevice, s.stream));
  WL_CUDA(s, cudaStreamSynchronize(s.stream));
  s.dist.p2p = true;
  return 0;
}

int dist_check_p2p_error(Sim& s) {
  if (!s.dist.p2p) return 0;
  int err = 0;
  cudaError_t e = cudaMemcpyAsync(&err, &((P2PFlags*)s.arena)->err, sizeof(int), cudaMemcpyDeviceToHost, s.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
  if (e != cudaSuccess) return wl_cuda_check(s, e, "p2p error flag");
  if (err) {
    if (!s.sticky) s.sticky = WL_ERR_NCCL;
    return wl_fail(WL_ERR_NCCL, "peer-memory wait timed out (a rank fell out of the collective call sequence)");
  }
  return 0;
}

// all ranks leave together: nobody frees an arena a peer may still be writing to.  The farewell
// handshake is collective, so it only runs for a handle that finished wl_create_dist on every rank
// (s.dist.ready) and is not in a failed state — a rank that failed during create must not make its
// peers spin until the timeout.
void dist_destroy(Sim& s) {
  if (s.dist.p2p && s.dist.ready && !s.sticky) {
    PushPlan plan{};
    for (int r = 0; r < s.dist.nranks; ++r) if (r != s.dist.rank) plan_add_peer(plan, r);
    push(s, plan, 0);
    cudaStreamSynchronize(s.stream);
  }
  if (s.dist.p2p)
    for (int r = 0; r < s.dist.nranks; ++r) if (r != s.dist.rank && s.dist.host.base[r]) cudaIpcCloseMemHandle(s.dist.host.base[r]);
  s.dist.p2p = false;
  if (s.dist.comm) {
    std::lock_guard<std::mutex> lock(g_comm_mu);
    for (auto& kv : g_comms) if ((void*)kv.second.comm == s.dist.comm) { --kv.second.refs; break; }
  }
  s.dist.comm = nullptr;
}


// destroys every cached communicator that no live handle uses (call once, after the last wl_destroy)
int dist_finalize() {
  std::lock_guard<std::mutex> lock(g_comm_mu);
  for (auto it = g_comms.begin(); it != g_comms.end();) {
    if (it->second.refs <= 0) {
      if (N.CommDestroy) N.CommDestroy(it->second.comm);
      it = g_comms.erase(it);
    } else {
      ++it;
    }
  }
  return 0;
}

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
  if (s.dist.p2p) {
    // planes start 3 floats (WL_OX) past a 16-byte boundary and cnt is a multiple of 4: the run shifted
    // down by 3 floats is float4-aligned; what it gains/loses at its ends is row padding
    PushPlan plan{};
    int64_t total4 = 0;
    auto add = [&](int peer, float* src, float* dst) {
      PushSeg& sg = plan.seg[plan.nseg++];
      sg.peer = peer;
      sg.src_off = ((char*)(src - WL_OX)) - s.arena;
      sg.dst_off = ((char*)(dst - WL_OX)) - s.arena;
      sg.n4 = (int64_t)cnt / 4;
      total4 += sg.n4;
      plan_add_peer(plan, peer);
    };
    if (2 * ncomp > WL_PUSH_MAXSEG) return wl_fail(WL_ERR_ARG, "internal: too many halo segments");
    for (int c = 0; c < ncomp; ++c) {
      float* b = base + c * g.comp;
      // my top interior planes → lower ghost planes of `up`; my bottom interior planes → upper ghosts of `dn`
      if (g.zhi_halo) add(up, b + (int64_t)(n - 1 - depth) * g.s[2], b + (int64_t)(1 - depth) * g.s[2]);
      if (g.zlo_halo) add(dn, b + (int64_t)1 * g.s[2], b + (int64_t)(n - 1) * g.s[2]);
    }
    ++s.collectives;
    return push(s, plan, total4);
  }
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
  if (s.dist.p2p) {  // push my chunk into every peer's copy (same float4 alignment shift as the halos)
    PushPlan plan{};
    float* mine = base + (size_t)s.dist.rank * count;
    for (int r = 0; r < s.dist.nranks; ++r) {
      if (r == s.dist.rank) continue;
      PushSeg& sg = plan.seg[plan.nseg++];
      sg.peer = r;
      sg.src_off = sg.dst_off = ((char*)(mine - WL_OX)) - s.arena;
      sg.n4 = (int64_t)count / 4;
      plan_add_peer(plan, r);
    }
    ++s.collectives;
    return push(s, plan, (int64_t)(count / 4) * (s.dist.nranks - 1));
  }
  WL_NCCL(s, N.AllGather(base + (size_t)s.dist.rank * count, base, count, ncclFloat, (ncclComm_t)s.dist.comm, s.stream));
  ++s.collectives;
  return 0;
}
