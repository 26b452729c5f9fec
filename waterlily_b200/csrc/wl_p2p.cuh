// wl_p2p.cuh — peer-memory communication between the z-slab ranks of one NVSwitch box.
//
// Every rank allocates ALL of its fields from one arena (a single cudaMalloc) whose first bytes are
// a P2PFlags block.  The arenas are exported with cudaIpcGetMemHandle and mapped by every other
// rank, and because every rank performs the same allocation sequence the arenas are symmetric: the
// address of "my field X, plane k" on rank r is  peer_base[r] + (my pointer − my base).  Kernels
// therefore store halo planes and reduction partials straight into the neighbour's memory over
// NVLink and signal with monotonically increasing epochs — no NCCL call, no extra launch for a
// reduction (the last block of the producing kernel exchanges the partials itself).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define WL_MAXRANKS 8

struct P2PMail {
  double v[2];
  unsigned long long epoch;
  unsigned long long pad;
};
// lives at offset 0 of every arena; written by peers (and by this rank for its own slots)
struct P2PFlags {
  unsigned long long ready[WL_MAXRANKS];  // ready[src]: src has reached push epoch e (its earlier kernels are done)
  unsigned long long done[WL_MAXRANKS];   // done[src]:  src's stores of push epoch e into this arena are visible
  P2PMail mbox[2][WL_MAXRANKS];           // all-reduce mailboxes, double-buffered by epoch parity
  unsigned long long ar_epoch;            // local: all-reduces performed so far (device-side counter)
  unsigned long long push_epoch;          // local: pushes performed so far (device-side: the fused V-cycle pushes on its own)
  unsigned int ticket;                    // local: block counter of the push kernel
  int err;                                // local: a wait timed out (the handle becomes sticky-failed)
  unsigned long long pad[3];
};
#define WL_P2P_FLAG_BYTES 1024  // sizeof(P2PFlags) rounded up; the first field starts behind it

struct P2PDev {
  int rank, nranks;
  unsigned long long timeout_ns;
  char* base[WL_MAXRANKS];  // arena of every rank as mapped into this process (base[rank] = own)
};

#ifdef __CUDACC__
static_assert(sizeof(P2PFlags) <= WL_P2P_FLAG_BYTES, "P2PFlags must fit its reserved block");

__device__ __forceinline__ unsigned long long p2p_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long p2p_ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void p2p_st_release(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ P2PFlags* p2p_flags(const P2PDev& c, int r) { return reinterpret_cast<P2PFlags*>(c.base[r]); }

// spin until *p >= want (EQ: == want); false on timeout (sets the local error flag)
template <bool EQ>
__device__ __forceinline__ bool p2p_wait(const P2PDev& c, const unsigned long long* p, unsigned long long want) {
  const unsigned long long t0 = p2p_now();
  for (;;) {
    const unsigned long long v = p2p_ld_acquire(p);
    if (EQ ? v == want : v >= want) return true;
    if (p2p_now() - t0 > c.timeout_ns) {
      p2p_flags(c, c.rank)->err = 1;
      return false;
    }
    __nanosleep(20);
  }
}

// All-reduce of two doubles across the ranks, called by ONE thread per rank (the thread that holds the
// grid totals).  Every rank adds the R contributions in rank order, so all ranks get identical bits.
// Mailboxes are double-buffered by epoch parity: a rank can start all-reduce e+1 (other parity) while a
// slow peer still reads e, but not e+2, which needs that peer's contribution to e+1.
__device__ __forceinline__ void p2p_allreduce2(const P2PDev& c, double& v0, double& v1, bool is_max) {
  P2PFlags* mine = p2p_flags(c, c.rank);
  const unsigned long long e = ++mine->ar_epoch;
  const int slot = (int)(e & 1ull), me = c.rank, R = c.nranks;
  for (int r = 0; r < R; ++r) {
    P2PMail* m = &p2p_flags(c, r)->mbox[slot][me];
    ((volatile double*)m->v)[0] = v0;
    ((volatile double*)m->v)[1] = v1;
  }
  __threadfence_system();
  for (int r = 0; r < R; ++r) p2p_st_release(&p2p_flags(c, r)->mbox[slot][me].epoch, e);
  double a0 = is_max ? -1e300 : 0.0, a1 = is_max ? -1e300 : 0.0;
  for (int r = 0; r < R; ++r) {
    const P2PMail* m = &mine->mbox[slot][r];
    if (!p2p_wait<true>(c, &m->epoch, e)) break;
    const double b0 = ((const volatile double*)m->v)[0], b1 = ((const volatile double*)m->v)[1];
    a0 = is_max ? fmax(a0, b0) : a0 + b0;
    a1 = is_max ? fmax(a1, b1) : a1 + b1;
  }
  v0 = a0;
  v1 = a1;
}
#endif
