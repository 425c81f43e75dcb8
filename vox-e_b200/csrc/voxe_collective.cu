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

// ---------------------------------------------------------------------------------------------------------------------
// Sparse variant: exchange only the 2x2x2 bricks some rank's backward touched.
//
// A 65 536-ray batch through a 512^3 SH-2 grid (cfg 5) writes a few percent of a 15 GB gradient volume; reducing the
// whole volume costs ~40 ms at N = 8, several times the render itself.  The backward kernels leave a trail -- one byte
// per brick, `tag` where a sample's corner 0 fell (voxe_render_bwd) -- so the exchange can follow it:
//   kernel A (flags_union_kernel): an all-reduce of the FLAG arrays with "some rank has the tag" as the operator: rank r
//     owns the r-th part of every CTA's slice, reads that part of every rank's flags (16 flags per load), and stores the
//     union into every rank's array.  17.6 MB at 512^3.  Same per-CTA barriers as the dense kernel; because CTA b leaves
//     only when CTA b of every peer has stored, the kernel's completion on a rank implies every peer's stores have landed.
//   kernel B (allreduce_sparse_kernel): walks the bricks in tiles of 32 (one lane per brick), dealt round-robin over the
//     ranks and then over the CTAs and warps -- a batch's trail is a slab of the volume, so contiguous ownership would leave
//     most CTAs idle.  A lane tests its brick against the dilated union (a brick holds gradients if one of the 8 flags at
//     (bx-dx, by-dy, bz-dz) carries the tag: a sample's corners reach one brick further in +x, +y, +z), the warp ballots,
//     and all lanes together reduce the tagged bricks' vectors (multimem.ld_reduce + multimem.st, or peer loads / stores).
// Afterwards every rank's flags hold the union, so its voxe_consume_grad visits exactly the bricks that now hold sums.
struct SparseParams {
  PeerParams base;
  unsigned char* touched[kMaxPeers];
  unsigned tagword;       // the tag in all four bytes
  unsigned n_bricks;
  long long n_flag_vec;   // 16-flag vectors (the arrays are padded to whole vectors)
  int BY, BZ;
  int vec_per_brick;      // 8 * CV
};

__device__ __forceinline__ uint4 ld_peer_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.relaxed.sys.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer_u4(uint4* p, const uint4& v) {
  asm volatile("st.global.relaxed.sys.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// bytes of `mine` replaced by the tag wherever the same byte of `theirs` is the tag
__device__ __forceinline__ unsigned merge_tag(unsigned mine, unsigned theirs, unsigned tagword) {
  const unsigned m = __vcmpeq4(theirs, tagword);
  return (mine & ~m) | (tagword & m);
}

__global__ void __launch_bounds__(kThreads, 1) flags_union_kernel(const __grid_constant__ SparseParams p, unsigned* fail_flag) {
  __shared__ unsigned s_fail;
  if (threadIdx.x == 0) s_fail = 0u;
  const long long per_cta = (p.n_flag_vec + gridDim.x - 1) / gridDim.x;
  const long long c0 = blockIdx.x * per_cta, c1 = min(c0 + per_cta, p.n_flag_vec);
  const long long per_rank = (max(c1 - c0, 0LL) + p.base.world - 1) / p.base.world;
  const long long r0 = min(c0 + p.base.rank * per_rank, c1), r1 = min(r0 + per_rank, c1);
  bool ok = peer_barrier(p.base, 0, &s_fail);  // the peers' backward kernels have finished
  if (ok) {
    for (long long i = r0 + threadIdx.x; i < r1; i += kThreads) {
      uint4 u = ld_peer_u4(reinterpret_cast<const uint4*>(p.touched[p.base.rank]) + i);
      for (int k = 0; k < p.base.world; ++k) {
        if (k == p.base.rank) continue;
        const uint4 v = ld_peer_u4(reinterpret_cast<const uint4*>(p.touched[k]) + i);
        u.x = merge_tag(u.x, v.x, p.tagword);
        u.y = merge_tag(u.y, v.y, p.tagword);
        u.z = merge_tag(u.z, v.z, p.tagword);
        u.w = merge_tag(u.w, v.w, p.tagword);
      }
      // a byte that is not the tag carries this rank's stale value to the peers: any value but the tag means the same
      for (int k = 0; k < p.base.world; ++k) st_peer_u4(reinterpret_cast<uint4*>(p.touched[(p.base.rank + k) % p.base.world]) + i, u);
    }
    ok = peer_barrier(p.base, 1, &s_fail);
  }
  if (!ok && threadIdx.x == 0 && fail_flag != nullptr) atomicOr(fail_flag, 1u);
}

__device__ __forceinline__ unsigned char ld_flag(const unsigned char* p) {  // written by the peers during the previous kernel
  unsigned v;
  asm volatile("ld.global.relaxed.gpu.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return (unsigned char)v;
}

__global__ void __launch_bounds__(kThreads, 1) allreduce_sparse_kernel(const __grid_constant__ SparseParams p, unsigned* fail_flag) {
  __shared__ unsigned s_fail;
  if (threadIdx.x == 0) s_fail = 0u;
  __syncthreads();
  const bool MULTICAST = p.base.mc != nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = kThreads / 32;
  const unsigned char* flags = p.touched[p.base.rank];
  const unsigned char want = (unsigned char)(p.tagword & 0xffu);
  const unsigned n_tiles = (p.n_bricks + 31u) / 32u;
  const int vpb = p.vec_per_brick;
  // tile t belongs to rank t % world; a rank's tiles are dealt over its CTAs, then over the warps of a CTA
  for (unsigned long long q = (unsigned long long)blockIdx.x + (unsigned long long)gridDim.x * warp;; q += (unsigned long long)gridDim.x * n_warps) {
    const unsigned long long t = q * p.base.world + p.base.rank;
    if (t >= n_tiles) break;
    const unsigned brick = (unsigned)t * 32u + lane;
    bool any = false;
    if (brick < p.n_bricks) {
      const unsigned bz = brick % (unsigned)p.BZ, r = brick / (unsigned)p.BZ;
      const unsigned by = r % (unsigned)p.BY, bx = r / (unsigned)p.BY;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const unsigned dx = k >> 2, dy = (k >> 1) & 1, dz = k & 1;
        if (bx >= dx && by >= dy && bz >= dz) any |= ld_flag(flags + ((bx - dx) * (unsigned)p.BY + (by - dy)) * (unsigned)p.BZ + (bz - dz)) == want;
      }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, any);
    const int n = __popc(mask) * vpb;  // vectors of this tile's tagged bricks, shared out over the lanes
    const long long tile0 = (long long)t * 32 * vpb;
    constexpr int U = 4;
    for (int i0 = lane; i0 < n; i0 += 32 * U) {
      long long off[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int id = i0 + 32 * u;
        if (id < n) {
          const int j = id / vpb;
          off[u] = tile0 + (long long)__fns(mask, 0u, j + 1) * vpb + (id - j * vpb);
        } else {
          off[u] = -1;
        }
      }
      float4 acc[U];
      if (MULTICAST) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (off[u] >= 0) acc[u] = multimem_ld_reduce(p.base.mc + off[u]);
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (off[u] >= 0) multimem_st(p.base.mc + off[u], acc[u]);
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        constexpr int G = 4;  // peers whose loads are in flight together (G * U 16-byte requests per thread)
        for (int k0 = 0; k0 < p.base.world; k0 += G) {
          float4 v[G][U];
#pragma unroll
          for (int k = 0; k < G; ++k)
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (k0 + k < p.base.world && off[u] >= 0) v[k][u] = ld_peer(p.base.buf[k0 + k] + off[u]);
#pragma unroll
          for (int k = 0; k < G; ++k)  // fixed rank order: the same sum on every rank and run
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (k0 + k < p.base.world && off[u] >= 0) {
                acc[u].x += v[k][u].x; acc[u].y += v[k][u].y; acc[u].z += v[k][u].z; acc[u].w += v[k][u].w;
              }
        }
        for (int k = 0; k < p.base.world; ++k) {
          const int dst = (p.base.rank + k) % p.base.world;
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (off[u] >= 0) st_peer(p.base.buf[dst] + off[u], acc[u]);
        }
      }
    }
  }
  // every rank has written its tiles into every volume (the peers' backward kernels were awaited by kernel A)
  const bool ok = peer_barrier(p.base, 1, &s_fail);
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

cudaError_t launch_allreduce_peer_sparse(const VoxePeerDesc& d, unsigned char* const* touched_peers, int tag, const int dims[3],
                                         int channels, unsigned* fail_flag, cudaStream_t stream) {
  SparseParams p{};
  p.base.world = d.world_size;
  p.base.rank = d.rank;
  for (int k = 0; k < d.world_size; ++k) {
    p.base.buf[k] = reinterpret_cast<float4*>(d.buffers[k]);
    p.base.sig[k] = d.signals[k];
    p.touched[k] = touched_peers[k];
  }
  p.base.mc = reinterpret_cast<float4*>(d.multicast);
  p.base.mc_share = p.base.mc ? 8 : 0;
  const int64_t bricks = packed_bricks(dims);
  p.n_bricks = (unsigned)bricks;
  p.n_flag_vec = (bricks + 15) / 16;
  p.BY = (dims[1] + 3) / 2;
  p.BZ = (dims[2] + 3) / 2;
  p.vec_per_brick = 8 * (channels / 4);
  p.base.n_vec = bricks * p.vec_per_brick;
  p.tagword = (unsigned)(tag & 0xff) * 0x01010101u;
  const long long want_a = (p.n_flag_vec + kThreads * 4 - 1) / (kThreads * 4);
  const int blocks_a = (int)(want_a < 1 ? 1 : (want_a > kBlocks ? kBlocks : want_a));
  flags_union_kernel<<<blocks_a, kThreads, 0, stream>>>(p, fail_flag);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const long long tiles_per_rank = ((bricks + 31) / 32 + d.world_size - 1) / d.world_size;
  const long long want_b = (tiles_per_rank + (kThreads / 32) - 1) / (kThreads / 32);
  const int blocks_b = (int)(want_b < 1 ? 1 : (want_b > kBlocks ? kBlocks : want_b));
  allreduce_sparse_kernel<<<blocks_b, kThreads, 0, stream>>>(p, fail_flag);
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
