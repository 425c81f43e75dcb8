// Which multi-stream read-modify-write patterns reach HBM speed on B200?
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int NS, int PER>
__global__ void __launch_bounds__(256) rmw(float4* base, int64_t stride_vec, int64_t n) {
  int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  float4 v[NS][PER];
#pragma unroll
  for (int p = 0; p < PER; ++p) {
    const int64_t i = t + (int64_t)p * gridDim.x * blockDim.x;
    if (i < n) {
#pragma unroll
      for (int s = 0; s < NS; ++s) v[s][p] = base[s * stride_vec + i];
    }
  }
#pragma unroll
  for (int p = 0; p < PER; ++p) {
    const int64_t i = t + (int64_t)p * gridDim.x * blockDim.x;
    if (i < n) {
#pragma unroll
      for (int s = 0; s < NS; ++s) { float4 x = v[s][p]; x.x = x.x * 0.999f + 1.f; x.y += x.x; base[s * stride_vec + i] = x; }
    }
  }
}
template <int NS>
__global__ void __launch_bounds__(256) rd_wr(const float4* __restrict__ in, float4* __restrict__ out, int64_t stride_vec, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
  for (int s = 0; s < NS; ++s) { float4 x = in[s * stride_vec + i]; acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w; }
#pragma unroll
  for (int s = 0; s < NS; ++s) out[s * stride_vec + i] = acc;
}
int main() {
  const int64_t n = 4096000;  // float4 per stream = 65.5 MB
  float4* buf; const int64_t slack = 1 << 20;  // in float4
  cudaMalloc(&buf, (size_t)(10 * (n + slack)) * 16); cudaMemset(buf, 0, (size_t)(10 * (n + slack)) * 16);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto time = [&](auto launch, const char* name, double bytes) {
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(a); for (int i = 0; i < 20; ++i) launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); double us = 1e3 * ms / 20;
    printf("%-56s %8.1f us %7.0f GB/s (%s)\n", name, us, bytes / us / 1e3, cudaGetErrorString(cudaGetLastError()));
  };
  const unsigned blocks = (unsigned)((n + 255) / 256);
  const double sb = n * 16.0;
  time([&] { rmw<1, 1><<<blocks, 256>>>(buf, n, n); }, "RMW 1 stream", 2 * sb);
  time([&] { rmw<2, 1><<<blocks, 256>>>(buf, n, n); }, "RMW 2 streams, contiguous arrays", 4 * sb);
  time([&] { rmw<4, 1><<<blocks, 256>>>(buf, n, n); }, "RMW 4 streams, contiguous arrays", 8 * sb);
  time([&] { rmw<4, 1><<<blocks, 256>>>(buf, n + 4097, n); }, "RMW 4 streams, arrays skewed by 64 KB+16 B", 8 * sb);
  time([&] { rmw<4, 1><<<blocks, 256>>>(buf, n + 333333, n); }, "RMW 4 streams, arrays skewed by 5.3 MB", 8 * sb);
  time([&] { rmw<4, 2><<<blocks / 2 + 1, 256>>>(buf, n, n); }, "RMW 4 streams, 2 vec/thread", 8 * sb);
  time([&] { rmw<4, 4><<<blocks / 4 + 1, 256>>>(buf, n, n); }, "RMW 4 streams, 4 vec/thread", 8 * sb);
  time([&] { rd_wr<1><<<blocks, 256>>>(buf, buf + 5 * (n + slack), n, n); }, "read 1 -> write 1 (copy)", 2 * sb);
  time([&] { rd_wr<4><<<blocks, 256>>>(buf, buf + 5 * (n + slack), n, n); }, "read 4 -> write 4 (out of place)", 8 * sb);
  time([&] { cudaMemcpyAsync(buf + 5 * (n + slack), buf, 4 * n * 16, cudaMemcpyDeviceToDevice); }, "cudaMemcpy 262 MB", 8 * sb);
  return 0;
}
