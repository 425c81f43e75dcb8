// bulkred_micro.cu -- how fast can an SM add a 112-byte vector (one SH-2 voxel: 28 floats) into scattered global
// addresses?  (a) 7 x red.global.add.v4.f32 per lane, as render_bwd_kernel does, against (b) staging the 112 bytes in
// shared memory and issuing ONE cp.reduce.async.bulk (TMA reduce) per lane.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bulkred_micro tools/bulkred_micro.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

constexpr int kVec = 7;  // float4 per voxel

__device__ __forceinline__ unsigned pcg(unsigned v) {
  const unsigned s = v * 747796405u + 2891336453u;
  const unsigned w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
  return (w >> 22u) ^ w;
}

__global__ void __launch_bounds__(128) red_kernel(float4* vol, unsigned n_vox, int ops, int local) {
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned h = pcg(tid);
  for (int k = 0; k < ops; ++k) {
    h = pcg(h + k);
    // `local`: neighbouring lanes hit neighbouring voxels (like neighbouring rays); else fully random
    const unsigned v = local ? ((pcg(tid / 8 + k * 7919u) + (tid & 7)) % n_vox) : (h % n_vox);
    float4* dst = vol + (size_t)v * kVec;
    const float x = 1e-6f * (float)(k + 1);
#pragma unroll
    for (int j = 0; j < kVec; ++j)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(x), "f"(x), "f"(x), "f"(x) : "memory");
  }
}

__global__ void __launch_bounds__(128) bulk_kernel(float4* vol, unsigned n_vox, int ops, int local) {
  __shared__ __align__(16) float4 stage[2][128][kVec];  // double-buffered 112-byte staging slot per thread
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned h = pcg(tid);
  for (int k = 0; k < ops; ++k) {
    h = pcg(h + k);
    const unsigned v = local ? ((pcg(tid / 8 + k * 7919u) + (tid & 7)) % n_vox) : (h % n_vox);
    float4* dst = vol + (size_t)v * kVec;
    const float x = 1e-6f * (float)(k + 1);
    float4* s = stage[k & 1][threadIdx.x];
    // the bulk op issued two iterations ago has finished reading this buffer
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
#pragma unroll
    for (int j = 0; j < kVec; ++j) s[j] = make_float4(x, x, x, x);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(s);
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(saddr), "n"(kVec * 16)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
  const unsigned n_vox = argc > 1 ? (unsigned)atoi(argv[1]) : 160u * 160u * 160u / 8u;  // default 57 MB: L2-resident
  const int ops = 64, blocks = 148 * 8, threads = 128;
  float4* vol;
  cudaMalloc(&vol, (size_t)n_vox * kVec * sizeof(float4));
  cudaMemset(vol, 0, (size_t)n_vox * kVec * sizeof(float4));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int local = 0; local < 2; ++local) {
    for (int which = 0; which < 2; ++which) {
      float best = 1e9f;
      for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        if (which == 0) red_kernel<<<blocks, threads>>>(vol, n_vox, ops, local);
        else bulk_kernel<<<blocks, threads>>>(vol, n_vox, ops, local);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
      }
      const double n = (double)blocks * threads * ops;
      printf("%-28s %-8s %8.3f ms  %7.2f G voxel-adds/s  %7.1f GB/s payload  (%s)\n", which ? "cp.reduce.async.bulk 112 B" : "7 x red.v4.f32",
             local ? "local" : "random", best, n / best * 1e-6, n * 112 / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
    }
  }
  // check: every add must have landed (sum of the volume)
  return 0;
}
