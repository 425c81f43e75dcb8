// Micro-benchmark for the fused optimiser-step pass: which part of the access pattern costs the time?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/adam_micro.cu -o tools/adam_micro && tools/adam_micro
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

struct BD { int X, Y, Z, BY, BZ; };

__device__ __forceinline__ bool slot_to_voxel(const BD& d, int64_t slot, int64_t& v) {
  const int64_t brick = slot >> 3; const int within = (int)(slot & 7);
  const int bz = (int)(brick % d.BZ); const int64_t t = brick / d.BZ; const int by = (int)(t % d.BY); const int bx = (int)(t / d.BY);
  const int x = 2 * bx + (within >> 2), y = 2 * by + ((within >> 1) & 1), z = 2 * bz + (within & 1);
  v = ((int64_t)x * d.Y + y) * d.Z + z; return x < d.X && y < d.Y && z < d.Z;
}
__device__ __forceinline__ bool slot_to_voxel32(const BD& d, unsigned slot, unsigned& v) {
  const unsigned brick = slot >> 3; const unsigned within = slot & 7;
  const unsigned bz = brick % d.BZ; const unsigned t = brick / d.BZ; const unsigned by = t % d.BY; const unsigned bx = t / d.BY;
  const unsigned x = 2 * bx + (within >> 2), y = 2 * by + ((within >> 1) & 1), z = 2 * bz + (within & 1);
  v = (x * d.Y + y) * d.Z + z; return x < d.X && y < d.Y && z < d.Z;
}

template <int MODE>  // 0 full(int64) 1 no scattered write 2 int32 math 3 int32 + no scatter 4 only streams no math
__global__ void __launch_bounds__(256) k(float4* __restrict__ packed, float4* __restrict__ pg, float4* __restrict__ pm,
                                         float4* __restrict__ pv, float* __restrict__ dens, float* __restrict__ feat, int64_t n, BD d) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int64_t v = 0;
  if (MODE == 0 || MODE == 1 || MODE == 9 || MODE == 11 || MODE == 13) { if (!slot_to_voxel(d, t, v)) return; }
  else if (MODE == 2 || MODE == 3) { unsigned vv; if (!slot_to_voxel32(d, (unsigned)t, vv)) return; v = vv; }
  const float4 g4 = pg[t];
  if (MODE == 8 || MODE == 9) pg[t] = make_float4(g4.y, g4.x, g4.w, g4.z);
  else if (MODE == 11 || MODE == 12) {}
  else if (MODE == 15) pg[t] = make_float4(1e-30f, 1e-30f, 1e-30f, 1e-30f);
  else if (MODE == 16) pg[t] = make_float4(0.f, 0.f, 0.f, 1e-30f);
  else pg[t] = make_float4(0, 0, 0, 0);
  const float4 p4 = packed[t], m4 = pm[t], v4 = pv[t];
  float g[4] = {g4.x, g4.y, g4.z, g4.w}, p[4] = {p4.x, p4.y, p4.z, p4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    m[c] = m[c] + (g[c] - m[c]) * 0.1f; vv[c] = vv[c] * 0.999f + g[c] * g[c] * 0.001f;
    if (MODE == 6) p[c] = p[c] - 0.03f * m[c] * vv[c];
    else if (MODE == 7) p[c] = p[c] - 0.03f * __fdividef(m[c], __fdividef(__fsqrt_rn(vv[c]), 0.0316f) + 1e-8f);
    else p[c] = p[c] - 0.03f * (m[c] / (sqrtf(vv[c]) / 0.0316f + 1e-8f));
    if (MODE == 0 || MODE == 2 || MODE == 9 || MODE == 11 || MODE == 13) { float* dst = c < 3 ? feat + v * 3 + c : dens + v; *dst = p[c]; }
  }
  packed[t] = make_float4(p[0], p[1], p[2], p[3]); pm[t] = make_float4(m[0], m[1], m[2], m[3]); pv[t] = make_float4(vv[0], vv[1], vv[2], vv[3]);
}

// variant 5: CTA = 32 bricks along z (256 slots); results staged in smem and written back as full rows
__global__ void __launch_bounds__(256) k_rows(float4* __restrict__ packed, float4* __restrict__ pg, float4* __restrict__ pm,
                                              float4* __restrict__ pv, float* __restrict__ dens, float* __restrict__ feat, BD d, int BX) {
  __shared__ float sf[4][64 * 3 + 1];
  __shared__ float sd[4][64 + 1];
  // blockIdx.x -> (bx, by, bz chunk of 32 bricks)
  const int chunks = (d.BZ + 31) / 32;
  const int cz = blockIdx.x % chunks; const int t2 = blockIdx.x / chunks; const int by = t2 % d.BY; const int bx = t2 / d.BY;
  const int bz = cz * 32 + (threadIdx.x >> 3); const int within = threadIdx.x & 7;
  const int x = 2 * bx + (within >> 2), y = 2 * by + ((within >> 1) & 1), z = 2 * bz + (within & 1);
  const bool ok = bz < d.BZ && x < d.X && y < d.Y && z < d.Z;
  const int64_t t = (((int64_t)bx * d.BY + by) * d.BZ + bz) * 8 + within;
  const int row = within >> 1, zl = (threadIdx.x >> 3) * 2 + (within & 1);
  if (ok) {
    const float4 g4 = pg[t]; pg[t] = make_float4(0, 0, 0, 0);
    const float4 p4 = packed[t], m4 = pm[t], v4 = pv[t];
    float g[4] = {g4.x, g4.y, g4.z, g4.w}, p[4] = {p4.x, p4.y, p4.z, p4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      m[c] = m[c] + (g[c] - m[c]) * 0.1f; vv[c] = vv[c] * 0.999f + g[c] * g[c] * 0.001f;
      p[c] = p[c] - 0.03f * (m[c] / (sqrtf(vv[c]) / 0.0316f + 1e-8f));
    }
    packed[t] = make_float4(p[0], p[1], p[2], p[3]); pm[t] = make_float4(m[0], m[1], m[2], m[3]); pv[t] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    sf[row][zl * 3 + 0] = p[0]; sf[row][zl * 3 + 1] = p[1]; sf[row][zl * 3 + 2] = p[2]; sd[row][zl] = p[3];
  }
  __syncthreads();
  const int z0 = cz * 64;
  const int nz = min(64, d.Z - z0);
  for (int i = threadIdx.x; i < 4 * 192; i += 256) {
    const int r = i / 192, j = i % 192; const int xx = 2 * bx + (r >> 1), yy = 2 * by + (r & 1);
    if (xx < d.X && yy < d.Y && j < nz * 3) feat[(((int64_t)xx * d.Y + yy) * d.Z + z0) * 3 + j] = sf[r][j];
  }
  { const int r = threadIdx.x >> 6, j = threadIdx.x & 63; const int xx = 2 * bx + (r >> 1), yy = 2 * by + (r & 1);
    if (xx < d.X && yy < d.Y && j < nz) dens[((int64_t)xx * d.Y + yy) * d.Z + z0 + j] = sd[r][j]; }
}

int main() {
  const int X = 160, Y = 160, Z = 160; BD d{X, Y, Z, (Y + 1) / 2, (Z + 1) / 2}; const int BX = (X + 1) / 2;
  const int64_t n = (int64_t)BX * d.BY * d.BZ * 8;
  float4 *packed, *pg, *pm, *pv; float *dens, *feat;
  cudaMalloc(&packed, n * 16); cudaMalloc(&pg, n * 16); cudaMalloc(&pm, n * 16); cudaMalloc(&pv, n * 16);
  cudaMalloc(&dens, (size_t)X * Y * Z * 4); cudaMalloc(&feat, (size_t)X * Y * Z * 12);
  cudaMemset(packed, 0, n * 16); cudaMemset(pg, 0, n * 16); cudaMemset(pm, 0, n * 16); cudaMemset(pv, 0, n * 16);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto time = [&](auto launch, const char* name) {
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(a); for (int i = 0; i < 20; ++i) launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); printf("%-44s %8.1f us  (%s)\n", name, 1e3 * ms / 20, cudaGetErrorString(cudaGetLastError()));
  };
  const unsigned blocks = (unsigned)((n + 255) / 256);
  time([&] { k<0><<<blocks, 256>>>(packed, pg, pm, pv, dens, feat, n, d); }, "0 full, int64 index math");
  time([&] { k<1><<<blocks, 256>>>(packed, pg, pm, pv, dens, feat, n, d); }, "1 no scattered write-back, int64");
  time([&] { k<2><<<blocks, 256>>>(packed, pg, pm, pv, dens, feat, n, d); }, "2 full, int32 index math");
  time([&] { k<3><<<blocks, 256>>>(packed, pg, pm, pv, dens, feat, n, d); }, "3 no scattered write-back, int32");
  time([&] { k<4><<<blocks, 256>>>(packed, pg, pm, pv, dens, feat, n, d); }, "4 streams only (no index math, no scatter)");
  time([&] { k<6><<<blocks, 256>>>(packed, pg, pm, pv, dens, feat, n, d); }, "6 streams only, FMA-only math");
  time([&] { k<7><<<blocks, 256>>>(packed, pg, pm, pv, dens, feat, n, d); }, "7 streams only, approx div/sqrt");
  {  // non-trivial data: gradients ~ +-1, kept alive between launches
    float* h = (float*)malloc(n * 16); for (int64_t i = 0; i < n * 4; ++i) h[i] = ((i * 2654435761u) % 2001) / 1000.f - 1.f;
    cudaMemcpy(pg, h, n * 16, cudaMemcpyHostToDevice); free(h);
  }
  time([&] { k<8><<<blocks, 256>>>(packed, pg, pm, pv, dens, feat, n, d); }, "8 streams only, IEEE math, non-zero gradients");
  time([&] { k<9><<<blocks, 256>>>(packed, pg, pm, pv, dens, feat, n, d); }, "9 non-zero data, scatter write-back, pg swizzled");
  time([&] { k<11><<<blocks, 256>>>(packed, pg, pm, pv, dens, feat, n, d); }, "11 non-zero data, scatter write-back, pg untouched");
  time([&] { k<12><<<blocks, 256>>>(packed, pg, pm, pv, dens, feat, n, d); }, "12 non-zero data, no scatter, pg untouched");
  { float4* noise; cudaMalloc(&noise, n * 16); cudaMemcpy(noise, pg, n * 16, cudaMemcpyDeviceToDevice);
    time([&] { cudaMemcpyAsync(pg, noise, n * 16, cudaMemcpyDeviceToDevice); }, "   (refill of pg alone)");
    time([&] { cudaMemcpyAsync(pg, noise, n * 16, cudaMemcpyDeviceToDevice); k<13><<<blocks, 256>>>(packed, pg, pm, pv, dens, feat, n, d); },
         "13 refill + full kernel (scatter, pg zeroed)");
    time([&] { cudaMemcpyAsync(pg, noise, n * 16, cudaMemcpyDeviceToDevice); k<1><<<blocks, 256>>>(packed, pg, pm, pv, dens, feat, n, d); },
         "14 refill + no-scatter kernel (pg zeroed)"); }
  { float4* noise; cudaMalloc(&noise, n * 16); cudaMemcpy(noise, pm, n * 16, cudaMemcpyDeviceToDevice);
    time([&] { cudaMemcpyAsync(pg, noise, n * 16, cudaMemcpyDeviceToDevice); k<15><<<blocks, 256>>>(packed, pg, pm, pv, dens, feat, n, d); },
         "15 refill + kernel writing 1e-30 to pg (no scatter)");
    time([&] { cudaMemcpyAsync(pg, noise, n * 16, cudaMemcpyDeviceToDevice); k<16><<<blocks, 256>>>(packed, pg, pm, pv, dens, feat, n, d); },
         "16 refill + kernel writing (0,0,0,1e-30) to pg");
    time([&] { cudaMemsetAsync(pg, 0, n * 16); k<12><<<blocks, 256>>>(packed, pg, pm, pv, dens, feat, n, d); },
         "17 memset(0) of pg + kernel that only reads pg");
    time([&] { cudaMemsetAsync(pg, 0, n * 16); }, "   (memset(0) alone)");
    time([&] { cudaMemcpyAsync(pg, noise, n * 16, cudaMemcpyDeviceToDevice); k<12><<<blocks, 256>>>(packed, pg, pm, pv, dens, feat, n, d); cudaMemsetAsync(pg, 0, n * 16); },
         "18 refill + read-only kernel + memset(0)"); }
  const unsigned rb = (unsigned)(BX * d.BY * ((d.BZ + 31) / 32));
  time([&] { k_rows<<<rb, 256>>>(packed, pg, pm, pv, dens, feat, d, BX); }, "5 brick columns, smem-staged row write-back");
  return 0;
}
