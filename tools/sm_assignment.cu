// tools/sm_assignment.cu -- which SM does CTA i of a (grid, block, regs, smem) launch land on, and when?
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/sm_assignment tools/sm_assignment.cu && tools/sm_assignment 512 128
// Prints, per SM, the CTAs it received in order of arrival -- the question behind the 2x spread of executed instructions
// between SMs that ncu shows for a 4096-ray batch (profiles/r2_*): do CTA i and CTA i + 148 share an SM?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
__global__ void __launch_bounds__(128, 5) probe(unsigned* smid, long long* t0, int spin) {
  __shared__ float pad[880];
  unsigned s;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
  long long t = clock64();
  float x = threadIdx.x;
  for (int i = 0; i < spin; ++i) x = x * 1.0001f + 0.5f;
  pad[threadIdx.x] = x;
  if (threadIdx.x == 0) { smid[blockIdx.x] = s; t0[blockIdx.x] = t; if (x == 12345.f) printf("%f", pad[3]); }
}
int main(int argc, char** argv) {
  int grid = argc > 1 ? atoi(argv[1]) : 512, block = argc > 2 ? atoi(argv[2]) : 128;
  unsigned* d_s; long long* d_t;
  cudaMalloc(&d_s, grid * 4); cudaMalloc(&d_t, grid * 8);
  for (int rep = 0; rep < 2; ++rep) probe<<<grid, block>>>(d_s, d_t, 20000);
  cudaDeviceSynchronize();
  std::vector<unsigned> s(grid); std::vector<long long> t(grid);
  cudaMemcpy(s.data(), d_s, grid * 4, cudaMemcpyDeviceToHost); cudaMemcpy(t.data(), d_t, grid * 8, cudaMemcpyDeviceToHost);
  std::vector<std::vector<int>> per(256);
  for (int i = 0; i < grid; ++i) per[s[i]].push_back(i);
  int shown = 0;
  for (int k = 0; k < 256 && shown < 40; ++k) if (!per[k].empty()) { printf("sm %3d:", k); for (int c : per[k]) printf(" %d", c); printf("\n"); ++shown; }
  std::vector<int> counts; for (auto& v : per) if (!v.empty()) counts.push_back((int)v.size());
  std::sort(counts.begin(), counts.end());
  printf("SMs used %zu, CTAs per SM min %d max %d\n", counts.size(), counts.front(), counts.back());
  printf("first 40 CTAs -> SM:"); for (int i = 0; i < 40 && i < grid; ++i) printf(" %u", s[i]); printf("\n");
  return 0;
}
