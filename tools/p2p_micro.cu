// p2p_micro.cu -- what SM-issued loads / stores over NVLink reach between two B200s, against the copy engine: the ceiling
// for voxe_allreduce_grads_peer's plain path.  One process, two devices, peer access enabled both ways.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/p2p_micro tools/p2p_micro.cu && gpurun --gpus 2 -- tools/p2p_micro
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ float4 ld_sys(const float4* p) {
  float4 v;
  asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(float4* p, const float4& v) {
  asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// mode 0: dst[i] = src[i];  mode 1: dst[i] = src[i] + src2[i] and dst2[i] = same (the all-reduce's traffic at N = 2)
template <int U, int MODE>
__global__ void __launch_bounds__(512) stream_kernel(const float4* __restrict__ src, const float4* __restrict__ src2, float4* __restrict__ dst,
                                                     float4* __restrict__ dst2, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * U) {
    float4 a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < n) {
        a[u] = ld_sys(src + i + u * stride);
        if (MODE == 1) b[u] = ld_sys(src2 + i + u * stride);
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < n) {
        if (MODE == 1) { a[u].x += b[u].x; a[u].y += b[u].y; a[u].z += b[u].z; a[u].w += b[u].w; }
        st_sys(dst + i + u * stride, a[u]);
        if (MODE == 1) st_sys(dst2 + i + u * stride, a[u]);
      }
  }
}

template <int U, int MODE>
float run(int dev, int grid, const float4* src, const float4* src2, float4* dst, float4* dst2, long long n, int other_dev = -1,
          const float4* osrc = nullptr, const float4* osrc2 = nullptr, float4* odst = nullptr, float4* odst2 = nullptr) {
  cudaEvent_t e0, e1;
  CK(cudaSetDevice(dev));
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaSetDevice(dev));
    CK(cudaDeviceSynchronize());
    if (other_dev >= 0) { CK(cudaSetDevice(other_dev)); CK(cudaDeviceSynchronize()); CK(cudaSetDevice(dev)); }
    CK(cudaEventRecord(e0));
    stream_kernel<U, MODE><<<grid, 512>>>(src, src2, dst, dst2, n);
    CK(cudaEventRecord(e1));
    if (other_dev >= 0) {
      CK(cudaSetDevice(other_dev));
      stream_kernel<U, MODE><<<grid, 512>>>(osrc, osrc2, odst, odst2, n);
      CK(cudaDeviceSynchronize());
      CK(cudaSetDevice(dev));
    }
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int n_dev = 0;
  CK(cudaGetDeviceCount(&n_dev));
  if (n_dev < 2) { printf("needs 2 GPUs\n"); return 0; }
  const long long bytes = 256ll << 20, n = bytes / 16;
  float4 *a0, *b0, *a1, *b1;
  CK(cudaSetDevice(0));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaMalloc(&a0, bytes));
  CK(cudaMalloc(&b0, bytes));
  CK(cudaMemset(a0, 0, bytes));
  CK(cudaMemset(b0, 0, bytes));
  CK(cudaSetDevice(1));
  CK(cudaDeviceEnablePeerAccess(0, 0));
  CK(cudaMalloc(&a1, bytes));
  CK(cudaMalloc(&b1, bytes));
  CK(cudaMemset(a1, 0, bytes));
  CK(cudaMemset(b1, 0, bytes));
  CK(cudaDeviceSynchronize());
  CK(cudaSetDevice(0));
  // copy engine
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    CK(cudaMemcpyPeerAsync(a0, 0, a1, 1, bytes, 0));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep == 2) printf("copy engine dev1 -> dev0, 256 MB: %.0f GB/s\n", bytes / ms * 1e-6);
  }
  const int grids[] = {32, 64, 128, 148, 296, 592};
  printf("%-44s", "GB/s of payload per direction; grid =");
  for (int g : grids) printf("%7d", g);
  printf("\n");
#define ROW(label, U, MODE, ...)                                                              \
  printf("%-44s", label);                                                                     \
  for (int g : grids) printf("%7.0f", (MODE == 1 ? bytes / 2 : bytes) / run<U, MODE>(__VA_ARGS__) * 1e-6); \
  printf("\n");
  // remote read -> local write
  ROW("read remote, write local   U=1", 1, 0, 0, g, a1, nullptr, a0, nullptr, n)
  ROW("read remote, write local   U=2", 2, 0, 0, g, a1, nullptr, a0, nullptr, n)
  ROW("read remote, write local   U=4", 4, 0, 0, g, a1, nullptr, a0, nullptr, n)
  ROW("read remote, write local   U=8", 8, 0, 0, g, a1, nullptr, a0, nullptr, n)
  ROW("read local, write remote   U=2", 2, 0, 0, g, a0, nullptr, a1, nullptr, n)
  ROW("read local, write remote   U=8", 8, 0, 0, g, a0, nullptr, a1, nullptr, n)
  ROW("read remote, write remote  U=4", 4, 0, 0, g, a1, nullptr, b1, nullptr, n)
  ROW("both GPUs: read remote, write local  U=4", 4, 0, 0, g, a1, nullptr, a0, nullptr, n, 1, a0, nullptr, a1, nullptr)
  // the two-shot all-reduce at N = 2: each GPU owns half; reads local + remote, writes local + remote (n/2 vectors each)
  ROW("both GPUs: all-reduce traffic  U=2", 2, 1, 0, g, a0, a1, b0, b1, n / 2, 1, a1 + n / 2, a0 + n / 2, b1 + n / 2, b0 + n / 2)
  ROW("both GPUs: all-reduce traffic  U=4", 4, 1, 0, g, a0, a1, b0, b1, n / 2, 1, a1 + n / 2, a0 + n / 2, b1 + n / 2, b0 + n / 2)
  ROW("both GPUs: all-reduce traffic  U=8", 8, 1, 0, g, a0, a1, b0, b1, n / 2, 1, a1 + n / 2, a0 + n / 2, b1 + n / 2, b0 + n / 2)
  ROW("one GPU:   all-reduce traffic  U=4", 4, 1, 0, g, a0, a1, b0, b1, n / 2)
  return 0;
}
