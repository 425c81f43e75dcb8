// voxe_collective.cu -- the one exchange step of the path: the all-reduce (SUM) of the packed voxel-gradient volume over
// the data-parallel ranks, once per optimiser step (SURVEY.md row e; the reference has no multi-GPU code).
//
// Two implementations behind the C ABI:
//   * voxe_allreduce_grads_peer -- own kernel over peer-mapped memory (NVLink / NVSwitch).  Every rank's gradient volume
//     is mapped into every process (CUDA IPC / VMM handles; the Python binding takes them from torch's symmetric-memory
//     allocator, a C host exchanges them itself), optionally also through an NVLS multicast mapping.  One launch, in
//     place, two-shot: rank r owns the r-th part of every CTA's slice; it sums that part over all ranks (plain 16-byte
//     loads from the peers' memory, or ONE multimem.ld_reduce -- the switch adds) and stores the sum into every rank's
//     volume (16-byte stores to each peer, or ONE multimem.st -- the switch replicates).  CTA b of every rank
//     synchronises only with CTA b of the other ranks (one flag per (CTA, peer) in a peer-mapped signal pad, set and
//     consumed with system-scope compare-and-swap), so there is no grid-wide barrier and the transfer of one slice
//     overlaps the reduction of the others.
//   * voxe_allreduce_grads -- ncclAllReduce on a caller-supplied communicator (libnccl is opened at run time, so the
//     library has no link-time dependency on it), for hosts whose volumes are not peer-mapped.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdlib.h>
#include <stdint.h>

#include "voxe.h"
#include "voxe_launch.h"

namespace voxe {
namespace {

constexpr int kMaxPeers = VOXE_MAX_PEERS;
constexpr int kBlocks = VOXE_SIGNAL_WORDS / (2 * VOXE_MAX_PEERS);  // flag slots: [phase 0|1][CTA][source rank]
constexpr int kThreads = 512;
constexpr long long kSpinLimit = 4000000000LL;  // ~2 s of SM clocks: a lost peer ends the kernel instead of hanging the GPU

struct PeerParams {
  float4* buf[kMaxPeers];
  uint32_t* sig[kMaxPeers];
  float4* mc;
  int world, rank;
  int mc_share;  // of every 8 consecutive CTAs, how many take the multicast path (the others: plain peer loads / stores)
  long long n_vec;
};

__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* addr, uint32_t expect, uint32_t value) {
  uint32_t old;
  asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(expect), "r"(value) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* addr, uint32_t expect, uint32_t value) {
  uint32_t old;
  asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(expect), "r"(value) : "memory");
  return old;
}

// CTA-to-CTA barrier across ranks: thread k < world raises this CTA's flag in rank k's pad (0 -> 1; it waits there while
// the previous use of the slot has not been consumed) and then consumes rank k's flag in the local pad (1 -> 0), so every
// slot is back at 0 when the barrier completes and can be reused by the next call.  Returns false on a timeout.
__device__ __forceinline__ bool peer_barrier(const PeerParams& p, int phase, unsigned* s_fail) {
  __syncthreads();  // everything this CTA wrote before the barrier ...
  if (threadIdx.x < p.world) {
    __threadfence_system();  // ... is visible to the peers before the flag is
    const int peer = threadIdx.x;
    const long long t0 = clock64();
    uint32_t* put = p.sig[peer] + ((size_t)phase * kBlocks + blockIdx.x) * kMaxPeers + p.rank;
    while (cas_release_sys(put, 0u, 1u) != 0u)
      if (clock64() - t0 > kSpinLimit) { atomicOr(s_fail, 1u); break; }
    uint32_t* get = p.sig[p.rank] + ((size_t)phase * kBlocks + blockIdx.x) * kMaxPeers + peer;
    while (cas_acquire_sys(get, 1u, 0u) != 1u)
      if (clock64() - t0 > kSpinLimit) { atomicOr(s_fail, 1u); break; }
  }
  __syncthreads();
  return *s_fail == 0u;
}

__device__ __forceinline__ float4 ld_peer(const float4* p) {  // peer memory changes between calls: no non-coherent path
  float4 v;
  asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(float4* p, const float4& v) {
  asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 multimem_ld_reduce(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float4* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// The two paths are bound by different resources -- the multicast path by the switch's reduction rate (~330 GB/s per GPU
// whatever the rank or CTA count), the plain path by the links (it moves twice the bytes) -- so a launch may split its CTAs
// between them (PeerParams::mc_share).
__global__ void __launch_bounds__(kThreads, 1) allreduce_peer_kernel(const __grid_constant__ PeerParams p, unsigned* fail_flag) {
  const bool MULTICAST = p.mc != nullptr && (int)(blockIdx.x & 7) < p.mc_share;
  __shared__ unsigned s_fail;
  if (threadIdx.x == 0) s_fail = 0u;
  // this CTA's slice of the volume, and this rank's part of it (16-byte vectors)
  const long long per_cta = (p.n_vec + gridDim.x - 1) / gridDim.x;
  const long long c0 = blockIdx.x * per_cta, c1 = min(c0 + per_cta, p.n_vec);
  const long long per_rank = (max(c1 - c0, 0LL) + p.world - 1) / p.world;
  const long long r0 = min(c0 + p.rank * per_rank, c1), r1 = min(r0 + per_rank, c1);

  // phase 0: the peers' backward kernels have finished (their streams reached this launch)
  bool ok = peer_barrier(p, 0, &s_fail);
  if (ok) {
    if (MULTICAST) {
      constexpr int U = 8;  // independent 16-byte requests per thread in flight (a round trip through the switch is ~2 us)
      for (long long i = r0 + threadIdx.x; i < r1; i += (long long)kThreads * U) {
        float4 acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (i + (long long)u * kThreads < r1) acc[u] = multimem_ld_reduce(p.mc + i + (long long)u * kThreads);
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (i + (long long)u * kThreads < r1) multimem_st(p.mc + i + (long long)u * kThreads, acc[u]);
      }
    } else {
      constexpr int U = 2, G = 8;  // all loads of a pass (up to 8 peers x 2 vectors) are issued before the first add
      for (long long i = r0 + threadIdx.x; i < r1; i += (long long)kThreads * U) {
        float4 acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k0 = 0; k0 < p.world; k0 += G) {
          float4 v[G][U];
#pragma unroll
          for (int k = 0; k < G; ++k)
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (k0 + k < p.world && i + (long long)u * kThreads < r1) v[k][u] = ld_peer(p.buf[k0 + k] + i + (long long)u * kThreads);
#pragma unroll
          for (int k = 0; k < G; ++k)  // fixed rank order: every element is summed identically on every rank and run
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (k0 + k < p.world && i + (long long)u * kThreads < r1) {
                acc[u].x += v[k][u].x; acc[u].y += v[k][u].y; acc[u].z += v[k][u].z; acc[u].w += v[k][u].w;
              }
        }
        for (int k = 0; k < p.world; ++k) {
          const int dst = (p.rank + k) % p.world;  // start with the local copy, spread the peers over the links
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (i + (long long)u * kThreads < r1) st_peer(p.buf[dst] + i + (long long)u * kThreads, acc[u]);
        }
      }
    }
    // phase 1: every rank has written its part of this slice into every volume
    ok = peer_barrier(p, 1, &s_fail);
  }
  if (!ok && threadIdx.x == 0 && fail_flag != nullptr) atomicOr(fail_flag, 1u);
}

}  // namespace

cudaError_t launch_allreduce_peer(const VoxePeerDesc& d, int64_t n_floats, unsigned* fail_flag, cudaStream_t stream) {
  PeerParams p{};
  p.world = d.world_size;
  p.rank = d.rank;
  p.n_vec = n_floats / 4;
  for (int k = 0; k < d.world_size; ++k) {
    p.buf[k] = reinterpret_cast<float4*>(d.buffers[k]);
    p.sig[k] = d.signals[k];
  }
  p.mc = reinterpret_cast<float4*>(d.multicast);
  // enough CTAs to keep ~1.5 MB of 16-byte requests in flight per direction, few enough that CTA b of every rank is
  // resident at the same time whatever else runs (one CTA per SM at most, see __launch_bounds__)
  const long long want = (p.n_vec + (long long)kThreads * 16 - 1) / ((long long)kThreads * 16);
  static const int block_cap = [] {  // tuning runs: VOXE_ALLREDUCE_BLOCKS caps the CTA count (must agree on every rank)
    const char* v = getenv("VOXE_ALLREDUCE_BLOCKS");
    const int n = v ? atoi(v) : 0;
    return (n >= 1 && n <= kBlocks) ? n : kBlocks;
  }();
  const int blocks = (int)(want < 1 ? 1 : (want > block_cap ? block_cap : want));
  p.mc_share = p.mc == nullptr ? 0 : (d.multicast_share >= 1 && d.multicast_share <= 8 ? d.multicast_share : 8);
  allreduce_peer_kernel<<<blocks, kThreads, 0, stream>>>(p, fail_flag);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// NCCL, opened at run time (nccl.h is not needed: the five entry points used here have had this shape since NCCL 2.0)
// ---------------------------------------------------------------------------------------------------------
struct NcclUniqueId {
  char internal[128];
};
static_assert(sizeof(NcclUniqueId) == VOXE_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");

struct NcclApi {
  void* handle = nullptr;
  int (*get_unique_id)(NcclUniqueId*) = nullptr;
  int (*comm_init_rank)(void**, int, NcclUniqueId, int) = nullptr;
  int (*comm_destroy)(void*) = nullptr;
  int (*all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*get_error_string)(int) = nullptr;
  const char* error = nullptr;
};

const NcclApi& nccl_api() {
  static const NcclApi api = [] {
    NcclApi a;
    // a process that already carries NCCL (torch imports its bundled copy) resolves the soname to that copy
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) {
      a.error = "libnccl.so.2 not found (set LD_LIBRARY_PATH to the directory that holds it)";
      return a;
    }
    a.get_unique_id = reinterpret_cast<decltype(a.get_unique_id)>(dlsym(a.handle, "ncclGetUniqueId"));
    a.comm_init_rank = reinterpret_cast<decltype(a.comm_init_rank)>(dlsym(a.handle, "ncclCommInitRank"));
    a.comm_destroy = reinterpret_cast<decltype(a.comm_destroy)>(dlsym(a.handle, "ncclCommDestroy"));
    a.all_reduce = reinterpret_cast<decltype(a.all_reduce)>(dlsym(a.handle, "ncclAllReduce"));
    a.get_error_string = reinterpret_cast<decltype(a.get_error_string)>(dlsym(a.handle, "ncclGetErrorString"));
    if (!a.get_unique_id || !a.comm_init_rank || !a.comm_destroy || !a.all_reduce) a.error = "libnccl lacks an expected entry point";
    return a;
  }();
  return api;
}

const char* nccl_unavailable() { return nccl_api().error; }

const char* nccl_error_string(int rc) {
  const NcclApi& a = nccl_api();
  return a.get_error_string ? a.get_error_string(rc) : "unknown NCCL error";
}

int nccl_unique_id(void* out) { return nccl_api().get_unique_id(reinterpret_cast<NcclUniqueId*>(out)); }

int nccl_comm_create(void** comm, int world, int rank, const void* id) {
  NcclUniqueId uid;
  __builtin_memcpy(&uid, id, sizeof(uid));
  return nccl_api().comm_init_rank(comm, world, uid, rank);
}

int nccl_comm_destroy(void* comm) { return nccl_api().comm_destroy(comm); }

int nccl_allreduce_sum_f32(void* comm, float* buf, size_t n, cudaStream_t stream) {
  return nccl_api().all_reduce(buf, buf, n, /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm, stream);
}

}  // namespace voxe
