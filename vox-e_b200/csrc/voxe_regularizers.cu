// voxe_regularizers.cu -- the per-step full-grid regularisers of Vox-E's edit loop as streaming kernels (sm_100a).
//
// SURVEY.md row f2.  The reference evaluates, once per optimiser step and over the WHOLE grid (sds_trainer.py:290-326):
//   * the density-correlation loss between the edited and the pretrained density grid (sds_trainer.py:494-524;
//     weight 200 by default, edit_pretrained_relu_field.py:171) or its L2 / L1 variants (:498-503), and
//   * total-variation losses on ReLU(_densities) and on _features (sds_trainer.py:318-326, :563-567; the same function
//     regularises the attention grids, attn_grid_trainer.py:659, grid_refine.py:709),
// each as ~10-25 elementwise / reduction launches over [X,Y,Z,C] tensors plus their autograd backward.  Here a loss is
// one read of the grid (+ one 3-to-5-value reduction) and its gradient is one more read fused with the accumulation
// into the dense gradient: pure HBM streams.
#include <algorithm>
#include <cmath>
#include <cstdint>

#include "voxe_launch.h"

namespace voxe {
namespace {

__device__ __forceinline__ float sgn(float d) { return (float)(d > 0.f) - (float)(d < 0.f); }  // torch.sign: 0 at 0

constexpr int kMaxCtas = 148 * 8;   // persistent launch: one wave of 8 resident 256-thread CTAs per SM
constexpr int kPrefetch = 6;       // iterations (of 256 elements per CTA) the TV kernel prefetches ahead into L2
constexpr int kStripPrefetch = 6;  // rows the strip kernel prefetches ahead into L2
constexpr int kTicket = 15;         // workspace slot of the last-CTA ticket
constexpr int kPartials = 16;       // workspace: [0, 16) results / statistics, then N partial sums per CTA

// Block-reduce N per-thread doubles and store them as this CTA's partial sums (no atomics: 25 600 same-address double
// atomics cost more than the whole streaming pass, and per-CTA slots make the loss bitwise reproducible).
template <int N>
__device__ __forceinline__ void block_store(double (&val)[N], double* __restrict__ ws) {
  __shared__ double part[N][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double v = val[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) part[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += part[threadIdx.x][w];
    ws[kPartials + threadIdx.x * kMaxCtas + blockIdx.x] = v;  // [N][kMaxCtas]: the fold below reads each row coalesced
    __threadfence();                                           // the partial before this CTA's ticket
  }
}

// True in exactly one CTA of the grid: the last one to get here, after every CTA's partial sums are visible.  The ticket
// (workspace slot 15, as an unsigned) wraps back to 0 on the last increment (atomicInc), so a workspace that was zeroed
// once at allocation stays ready for every later call; the reduction needs no second launch.
__device__ __forceinline__ bool last_cta_done(double* __restrict__ ws) {
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicInc(reinterpret_cast<unsigned*>(ws + kTicket), gridDim.x - 1);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) __threadfence();
  return last;
}

// Sum the per-CTA partials (the 256 threads of the last CTA): every thread returns with the N totals.
template <int N>
__device__ __forceinline__ void gather_partials(const double* __restrict__ ws, int n_ctas, double (&tot)[N]) {
  __shared__ double part[N][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // block_store's use of its shared buffer is over
  // all N * ceil(kMaxCtas / 256) loads are independent and issued together: one L2 round trip, not one per partial
  constexpr int kPer = (kMaxCtas + 255) / 256;
  double ld[N][kPer];
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int c = threadIdx.x + i * 256;
      ld[k][i] = c < n_ctas ? __ldcg(ws + kPartials + k * kMaxCtas + c) : 0.0;  // L2: written by other CTAs
    }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double v = 0.0;
#pragma unroll
    for (int i = 0; i < kPer; ++i) v += ld[k][i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) part[k][warp] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += part[k][w];
    tot[k] = v;
  }
}

__device__ __forceinline__ void tv_finalize(const double* __restrict__ ws, float* __restrict__ loss, double n0, double n1,
                                            double n2);

// ---- total variation -------------------------------------------------------------------------------------------
// grid [X,Y,Z,C], z fastest, channels innermost, taken as X*Y rows of Z*C contiguous floats.  Persistent launch: CTA b
// owns the flat element span [b * chunk, (b+1) * chunk); its threads walk the span 256 elements at a time, so every load
// instruction of a warp is one contiguous 128-byte request, and each thread carries its (x, y, offset-in-row) along
// incrementally (no divisions in the loop).  An element's six axis neighbours sit at +-C (z), +-Z*C (y), +-Y*Z*C (x):
// the z neighbours are the same lines in L1, the y / x neighbours are lines other CTAs stream at the same time (L2), so
// DRAM sees the grid once.
//   loss  = (mean|d_x h| + mean|d_y h| + mean|d_z h|) / 3,  h = RELU ? max(g, 0) : g      (sds_trainer.py:563-567)
//   dloss/dg[v] = [g[v] > 0] * sum_a 1/(3 n_a) * (sign(h[v] - h[v-1_a]) - sign(h[v+1_a] - h[v]))
template <bool RELU, bool DO_SUM, bool DO_GRAD>
__global__ void __launch_bounds__(256) tv_kernel(const float* __restrict__ g, float* __restrict__ grad, double* __restrict__ ws,
                                                 int X, int Y, int ZC, int C, int64_t sx, int64_t total, int64_t chunk,
                                                 float cx, float cy, float cz, const float* __restrict__ upstream, int accumulate,
                                                 float* __restrict__ loss, double n0, double n1, double n2) {
  const int64_t begin = (int64_t)blockIdx.x * chunk;
  const int64_t end = begin + chunk < total ? begin + chunk : total;
  int64_t idx = begin + threadIdx.x;
  int64_t row = idx / ZC;
  int e = (int)(idx - row * ZC);
  int x = (int)(row / Y), y = (int)(row - (int64_t)x * Y);
  float up = 1.f;
  if (DO_GRAD && upstream) up = __ldg(upstream);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (; idx < end; idx += 256) {
    // Branch-free: a missing neighbour (grid face) reads the element itself, so its difference -- and with it |d| and
    // sign(d) -- is exactly 0; all loads of an iteration are independent and issued up front.
    const float* r = g + idx;
    // A thread touches 4 fresh bytes of the grid (and of the gradient) per iteration -- far too little in flight to cover
    // DRAM latency even at full occupancy (measured: 1.8 TB/s).  So each warp pulls the lines it will need kPrefetch
    // iterations from now into L2; the loads below then see L2 latency only.
    if (idx + kPrefetch * 256 < end) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(r + kPrefetch * 256));
      if (DO_GRAD && accumulate) asm volatile("prefetch.global.L2 [%0];" ::"l"(grad + idx + kPrefetch * 256));
    }
    float old = 0.f;
    if (DO_GRAD && accumulate) old = grad[idx];
    const float raw = __ldg(r);
    float zp = __ldg(e + C < ZC ? r + C : r);
    float yp = __ldg(y < Y - 1 ? r + ZC : r);
    float xp = __ldg(x < X - 1 ? r + sx : r);
    float zm = raw, ym = raw, xm = raw;
    if (DO_GRAD) {
      zm = __ldg(e >= C ? r - C : r);
      ym = __ldg(y > 0 ? r - ZC : r);
      xm = __ldg(x > 0 ? r - sx : r);
    }
    float v = raw;
    if (RELU) {
      v = fmaxf(raw, 0.f);
      zp = fmaxf(zp, 0.f); yp = fmaxf(yp, 0.f); xp = fmaxf(xp, 0.f);
      zm = fmaxf(zm, 0.f); ym = fmaxf(ym, 0.f); xm = fmaxf(xm, 0.f);
    }
    const float dzp = zp - v, dyp = yp - v, dxp = xp - v;
    if (DO_SUM) {
      s2 += fabsf(dzp);
      s1 += fabsf(dyp);
      s0 += fabsf(dxp);
    }
    if (DO_GRAD) {
      const float gz = sgn(v - zm) - sgn(dzp), gy = sgn(v - ym) - sgn(dyp), gx = sgn(v - xm) - sgn(dxp);
      float t = (cx * gx + cy * gy + cz * gz) * up;
      if (RELU && !(raw > 0.f)) t = 0.f;
      grad[idx] = old + t;
    }
    e += 256;  // advance (x, y, e) with the flat index
    while (e >= ZC) {
      e -= ZC;
      if (++y == Y) {
        y = 0;
        ++x;
      }
    }
  }
  if (DO_SUM) {
    double v[3] = {(double)s0, (double)s1, (double)s2};
    block_store<3>(v, ws);
    if (last_cta_done(ws)) tv_finalize(ws, loss, n0, n1, n2);
  }
}

// Vector variant (rows of Z*C floats with Z*C % 4 == 0, 16-byte aligned buffers -- every grid the renderer packs):
// a thread owns FOUR consecutive floats of a row per iteration, so the index / predicate arithmetic is paid once per 16
// bytes and every load is a 16-byte one (the scalar kernel above spends ~110 instructions per float and is issue-bound at
// 2 TB/s).  The y / x neighbours of an aligned group are aligned groups.  The z neighbours sit +-C floats away:
//   ZMODE 1..3  C = ZMODE: they lie in the window [previous group | own group | next group]
//   ZMODE 4     C % 4 == 0: aligned groups at +-C
//   ZMODE 0     any other C: eight scalar loads
// A neighbour beyond a grid face is replaced by the element itself (difference exactly 0), as in the scalar kernel.
template <bool RELU, bool DO_SUM, bool DO_GRAD, int ZMODE>
__global__ void __launch_bounds__(256) tv_vec_kernel(const float* __restrict__ g, float* __restrict__ grad, double* __restrict__ ws,
                                                     int X, int Y, int GR, int C, int64_t sxg, int64_t total_g, int64_t chunk,
                                                     float cx, float cy, float cz, const float* __restrict__ upstream,
                                                     int accumulate, float* __restrict__ loss, double n0, double n1, double n2) {
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* grad4 = reinterpret_cast<float4*>(grad);
  const int64_t begin = (int64_t)blockIdx.x * chunk;
  const int64_t end = begin + chunk < total_g ? begin + chunk : total_g;
  int64_t idx = begin + threadIdx.x;  // group index
  int64_t row = idx / GR;
  int eg = (int)(idx - row * GR);
  int x = (int)(row / Y), y = (int)(row - (int64_t)x * Y);
  const int ZC = 4 * GR;
  float up = 1.f;
  if (DO_GRAD && upstream) up = __ldg(upstream);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (; idx < end; idx += 256) {
    const float4* r = g4 + idx;
    if (idx + kPrefetch * 256 < end) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(r + kPrefetch * 256));
      if (DO_GRAD && accumulate) asm volatile("prefetch.global.L2 [%0];" ::"l"(grad4 + idx + kPrefetch * 256));
    }
    float4 old4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (DO_GRAD && accumulate) old4 = grad4[idx];
    const float4 own4 = __ldg(r);
    const float4 yp4 = __ldg(y < Y - 1 ? r + GR : r), xp4 = __ldg(x < X - 1 ? r + sxg : r);
    float4 ym4 = own4, xm4 = own4;
    if (DO_GRAD) {
      ym4 = __ldg(y > 0 ? r - GR : r);
      xm4 = __ldg(x > 0 ? r - sxg : r);
    }
    float own[4] = {own4.x, own4.y, own4.z, own4.w};
    float zp[4], zm[4];
    const int e0 = 4 * eg;
    if (ZMODE >= 1 && ZMODE <= 3) {
      const float4 nx4 = __ldg(eg < GR - 1 ? r + 1 : r);
      float4 pv4 = own4;
      if (DO_GRAD) pv4 = __ldg(eg > 0 ? r - 1 : r);
      const float w[12] = {pv4.x, pv4.y, pv4.z, pv4.w, own4.x, own4.y, own4.z, own4.w, nx4.x, nx4.y, nx4.z, nx4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        zp[k] = (e0 + k + ZMODE < ZC) ? w[4 + k + ZMODE] : own[k];
        zm[k] = (e0 + k >= ZMODE) ? w[4 + k - ZMODE] : own[k];
      }
    } else if (ZMODE == 4) {
      const int cg = C >> 2;
      const float4 a4 = __ldg(eg + cg < GR ? r + cg : r);
      float4 b4 = own4;
      if (DO_GRAD) b4 = __ldg(eg >= cg ? r - cg : r);
      zp[0] = a4.x; zp[1] = a4.y; zp[2] = a4.z; zp[3] = a4.w;
      zm[0] = b4.x; zm[1] = b4.y; zm[2] = b4.z; zm[3] = b4.w;
    } else {
      const float* rs = g + 4 * idx;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        zp[k] = __ldg(e0 + k + C < ZC ? rs + k + C : rs + k);
        zm[k] = DO_GRAD ? __ldg(e0 + k >= C ? rs + k - C : rs + k) : own[k];
      }
    }
    const float yp[4] = {yp4.x, yp4.y, yp4.z, yp4.w}, ym[4] = {ym4.x, ym4.y, ym4.z, ym4.w};
    const float xp[4] = {xp4.x, xp4.y, xp4.z, xp4.w}, xm[4] = {xm4.x, xm4.y, xm4.z, xm4.w};
    float out[4] = {old4.x, old4.y, old4.z, old4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float raw = own[k];
      float v = raw, a = zp[k], b = yp[k], c = xp[k], d = zm[k], e = ym[k], f = xm[k];
      if (RELU) {
        v = fmaxf(v, 0.f);
        a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); c = fmaxf(c, 0.f);
        d = fmaxf(d, 0.f); e = fmaxf(e, 0.f); f = fmaxf(f, 0.f);
      }
      const float dzp = a - v, dyp = b - v, dxp = c - v;
      if (DO_SUM) {
        s2 += fabsf(dzp);
        s1 += fabsf(dyp);
        s0 += fabsf(dxp);
      }
      if (DO_GRAD) {
        const float gz = sgn(v - d) - sgn(dzp), gy = sgn(v - e) - sgn(dyp), gx = sgn(v - f) - sgn(dxp);
        float t = (cx * gx + cy * gy + cz * gz) * up;
        if (RELU && !(raw > 0.f)) t = 0.f;
        out[k] += t;
      }
    }
    if (DO_GRAD) grad4[idx] = make_float4(out[0], out[1], out[2], out[3]);
    eg += 256;  // advance (x, y, eg) with the group index
    while (eg >= GR) {
      eg -= GR;
      if (++y == Y) {
        y = 0;
        ++x;
      }
    }
  }
  if (DO_SUM) {
    double v[3] = {(double)s0, (double)s1, (double)s2};
    block_store<3>(v, ws);
    if (last_cta_done(ws)) tv_finalize(ws, loss, n0, n1, n2);
  }
}

// Gradient variant (with or without the loss): a thread owns one 16-byte group of a z-row and walks a STRIP of rows along
// y.  What the walk saves over tv_vec_kernel, which is issue-bound at ~72 SASS instructions per float:
//   * the row above is the previous iteration's own row and the row below becomes the next one: one new row group per
//     step instead of three, and max(., 0) is applied once per loaded value;
//   * sign(h[y+1] - h[y]) is the forward term of row y and the backward term of row y + 1: carried in registers; along z
//     the same holds inside the group (C <= 3: the backward sign of element k is the forward sign of element k - C);
//   * x, the z-group and every face predicate but y's are fixed for the strip: the index arithmetic leaves the loop.
// The loads of row y + 1 are issued before row y is evaluated (two rows of state in registers).  Strips are SEG rows long,
// chosen on the host so that the launch has ~150 k threads; a strip that starts inside the grid fetches the group above it
// once for its first backward sign.
template <bool RELU>
__device__ __forceinline__ float tv_act(float v) { return RELU ? fmaxf(v, 0.f) : v; }

template <bool RELU>
__device__ __forceinline__ void tv_unpack(const float4& a, float (&out)[4]) {
  out[0] = tv_act<RELU>(a.x); out[1] = tv_act<RELU>(a.y); out[2] = tv_act<RELU>(a.z); out[3] = tv_act<RELU>(a.w);
}

template <bool RELU, bool DO_SUM, int ZMODE>
__global__ void __launch_bounds__(256, 4) tv_strip_kernel(const float* __restrict__ g, float* __restrict__ grad, double* __restrict__ ws,
                                                          int X, int Y, int GR, int C, int64_t sxg, int SEG, int NSEG, int64_t n_strips,
                                                          float cx, float cy, float cz, const float* __restrict__ upstream,
                                                          int accumulate, float* __restrict__ loss, double n0, double n1, double n2) {
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* grad4 = reinterpret_cast<float4*>(grad);
  const float up = upstream ? __ldg(upstream) : 1.f;
  const float cxu = cx * up, cyu = cy * up, czu = cz * up;
  const int ZC = 4 * GR;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int64_t strip = (int64_t)blockIdx.x * 256 + threadIdx.x; strip < n_strips; strip += (int64_t)gridDim.x * 256) {
    const int eg = (int)(strip % GR), e0 = 4 * eg;
    const int64_t t = strip / GR;
    const int yseg = (int)(t % NSEG), x = (int)(t / NSEG);
    const int y0 = yseg * SEG, y1 = min(Y, y0 + SEG);
    const int64_t dxm = x > 0 ? -sxg : 0, dxp = x < X - 1 ? sxg : 0;  // a missing neighbour reads the element itself
    const float4* r = g4 + ((int64_t)x * Y + y0) * GR + eg;
    float4* gr = grad4 + ((int64_t)x * Y + y0) * GR + eg;
    // the strip's lines of the grid and of the gradient, kStripPrefetch rows ahead of the walk, into L2
#pragma unroll
    for (int k = 1; k <= kStripPrefetch; ++k)
      if (y0 + k < Y) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r + (int64_t)k * GR));
        if (accumulate) asm volatile("prefetch.global.L2 [%0];" ::"l"(gr + (int64_t)k * GR));
      }
    float v[4], sy[4] = {0.f, 0.f, 0.f, 0.f};
    tv_unpack<RELU>(__ldg(r), v);
    if (y0 > 0) {  // first backward sign of the strip: the row above belongs to another strip
      float above[4];
      tv_unpack<RELU>(__ldg(r - GR), above);
#pragma unroll
      for (int k = 0; k < 4; ++k) sy[k] = sgn(v[k] - above[k]);
    }
    for (int y = y0; y < y1; ++y) {
      // all loads of the row first: the group below, the x neighbours, the z neighbours, the old gradient
      const float4 yp4 = __ldg(y < Y - 1 ? r + GR : r);
      const float4 xm4 = __ldg(r + dxm), xp4 = __ldg(r + dxp);
      float4 za4, zb4;  // ZMODE 1..3: next / previous group; ZMODE 4: the groups at +-C
      float zs[8];      // ZMODE 0: scalars
      if (ZMODE >= 1 && ZMODE <= 3) {
        za4 = __ldg(eg < GR - 1 ? r + 1 : r);
        zb4 = __ldg(eg > 0 ? r - 1 : r);
      } else if (ZMODE == 4) {
        const int cg = C >> 2;
        za4 = __ldg(eg + cg < GR ? r + cg : r);
        zb4 = __ldg(eg >= cg ? r - cg : r);
      } else {
        const float* rs = reinterpret_cast<const float*>(r);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          zs[k] = __ldg(e0 + k + C < ZC ? rs + k + C : rs + k);
          zs[4 + k] = __ldg(e0 + k >= C ? rs + k - C : rs + k);
        }
      }
      const float4 old4 = accumulate ? *gr : make_float4(0.f, 0.f, 0.f, 0.f);
      if (y + 1 + kStripPrefetch < Y) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r + (int64_t)(1 + kStripPrefetch) * GR));
        if (accumulate) asm volatile("prefetch.global.L2 [%0];" ::"l"(gr + (int64_t)(1 + kStripPrefetch) * GR));
      }
      float yp[4], xm[4], xp[4], zp[4], zm[4];
      tv_unpack<RELU>(yp4, yp);
      tv_unpack<RELU>(xm4, xm);
      tv_unpack<RELU>(xp4, xp);
      if (ZMODE >= 1 && ZMODE <= 3) {
        float nx[4], pv[4];
        tv_unpack<RELU>(za4, nx);
        tv_unpack<RELU>(zb4, pv);
        const float w[12] = {pv[0], pv[1], pv[2], pv[3], v[0], v[1], v[2], v[3], nx[0], nx[1], nx[2], nx[3]};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          zp[k] = (e0 + k + ZMODE < ZC) ? w[4 + k + ZMODE] : v[k];
          zm[k] = (e0 + k >= ZMODE) ? w[4 + k - ZMODE] : v[k];
        }
      } else if (ZMODE == 4) {
        tv_unpack<RELU>(za4, zp);
        tv_unpack<RELU>(zb4, zm);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          zp[k] = tv_act<RELU>(zs[k]);
          zm[k] = tv_act<RELU>(zs[4 + k]);
        }
      }
      float out[4] = {old4.x, old4.y, old4.z, old4.w};
      float fz[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float dzp = zp[k] - v[k], dyp = yp[k] - v[k], dxpk = xp[k] - v[k];
        if (DO_SUM) {
          s2 += fabsf(dzp);
          s1 += fabsf(dyp);
          s0 += fabsf(dxpk);
        }
        fz[k] = sgn(dzp);
        // along z inside the group the backward sign of element k is the forward sign of element k - C
        const float bz = (ZMODE >= 1 && ZMODE <= 3 && k >= ZMODE) ? fz[k >= ZMODE ? k - ZMODE : 0] : sgn(v[k] - zm[k]);
        const float fy = sgn(dyp), by = sy[k];
        sy[k] = fy;  // ... and along y it is the forward sign of the row above: carried to the next iteration
        const float gx = sgn(v[k] - xm[k]) - sgn(dxpk);
        float tt = fmaf(cxu, gx, fmaf(cyu, by - fy, czu * (bz - fz[k])));
        if (RELU && !(v[k] > 0.f)) tt = 0.f;
        out[k] += tt;
      }
      *gr = make_float4(out[0], out[1], out[2], out[3]);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = yp[k];
      r += GR;
      gr += GR;
    }
  }
  if (DO_SUM) {
    double v[3] = {(double)s0, (double)s1, (double)s2};
    block_store<3>(v, ws);
    if (last_cta_done(ws)) tv_finalize(ws, loss, n0, n1, n2);
  }
}

// Last CTA of a TV launch: fold the partials into the loss.
__device__ __forceinline__ void tv_finalize(const double* __restrict__ ws, float* __restrict__ loss, double n0, double n1,
                                            double n2) {
  double tot[3];
  gather_partials<3>(ws, gridDim.x, tot);
  if (threadIdx.x != 0) return;
  // mean over an empty difference tensor (an axis of extent 1) is NaN in torch; 0/0 reproduces that
  const float m0 = (float)(tot[0] / n0), m1 = (float)(tot[1] / n1), m2 = (float)(tot[2] / n2);
  *loss = (m0 + m1 + m2) / 3.f;
}

// ---- density correlation / L2 / L1 between two grids --------------------------------------------------------------
// workspace doubles: [5] mean_a, [6] mean_b, [7] var_a, [8] var_b, [9] cov, [10] sqrt(var_a*var_b); per-CTA partials from 16 on
enum { kCorr = 0, kL2 = 1, kL1 = 2 };

// Last CTA of a statistics launch: fold the partials into the loss and the statistics the gradient pass reads.
__device__ __forceinline__ void pair_finalize(double* __restrict__ ws, float* __restrict__ loss, double n, int mode, float eps) {
  double tot[5];
  gather_partials<5>(ws, gridDim.x, tot);
  if (threadIdx.x != 0) return;
  if (mode != kCorr) {
    *loss = (float)(tot[0] / n);
    return;
  }
  const double ma = tot[0] / n, mb = tot[1] / n;
  const double va = fmax(tot[2] / n - ma * ma, 0.0), vb = fmax(tot[3] / n - mb * mb, 0.0);
  const double cov = tot[4] / n - ma * mb;
  const double d = sqrt(va * vb);
  ws[5] = ma; ws[6] = mb; ws[7] = va; ws[8] = vb; ws[9] = cov; ws[10] = d;
  // 1 - mean(covariance_grid / (denominator + eps))                                   (sds_trainer.py:520-524)
  *loss = 1.f - (float)(cov / (d + (double)eps));
}

template <int MODE>
__device__ __forceinline__ void pair_accumulate(float xf, float yf, double (&acc)[5]) {
  if (MODE == kCorr) {
    const double x = (double)xf, y = (double)yf;
    acc[0] += x;
    acc[1] += y;
    acc[2] = fma(x, x, acc[2]);
    acc[3] = fma(y, y, acc[3]);
    acc[4] = fma(x, y, acc[4]);
  } else if (MODE == kL2) {
    const double d = (double)(xf - yf);  // the fp32 difference the reference squares
    acc[0] = fma(d, d, acc[0]);
  } else {
    acc[0] += (double)fabsf(xf - yf);
  }
}

// n4 = number of whole float4 groups the vector loop covers (0 when the pointers are not 16-byte aligned)
template <int MODE>
__global__ void __launch_bounds__(256) pair_stats_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                         int64_t n4, double* __restrict__ sums, float* __restrict__ loss, float eps) {
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  for (int64_t i = t0; i < n4; i += stride) {
    const float4 x = __ldg(a4 + i), y = __ldg(b4 + i);
    pair_accumulate<MODE>(x.x, y.x, acc);
    pair_accumulate<MODE>(x.y, y.y, acc);
    pair_accumulate<MODE>(x.z, y.z, acc);
    pair_accumulate<MODE>(x.w, y.w, acc);
  }
  for (int64_t i = 4 * n4 + t0; i < n; i += stride) pair_accumulate<MODE>(__ldg(a + i), __ldg(b + i), acc);
  block_store<5>(acc, sums);
  if (last_cta_done(sums)) pair_finalize(sums, loss, (double)n, MODE, eps);
}

// correlation_grid = (a - mean_a)(b - mean_b) / (denominator + eps): the second return value of the reference function
__global__ void __launch_bounds__(256) corr_grid_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                        const double* __restrict__ ws, float* __restrict__ out, float eps) {
  const float ma = (float)ws[5], mb = (float)ws[6];
  const float inv = 1.f / ((float)ws[10] + eps);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (__ldg(a + i) - ma) * (__ldg(b + i) - mb) * inv;
}

// dloss/da.  Correlation mode, with D = d + eps, d = sqrt(var_a var_b):
//   dloss/da_i = -[(b_i - mean_b) / D - cov * var_b * (a_i - mean_a) / (D^2 d)] / n
template <int MODE>
__global__ void __launch_bounds__(256) pair_grad_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                        int64_t n4, const double* __restrict__ ws, const float* __restrict__ upstream,
                                                        float scale, float eps, float* __restrict__ grad, int accumulate) {
  float up = scale;
  if (upstream) up *= __ldg(upstream);
  float k1 = 0.f, k2 = 0.f, ma = 0.f, mb = 0.f;
  if (MODE == kCorr) {
    const double d = ws[10], D = d + (double)eps, cov = ws[9], vb = ws[8];
    ma = (float)ws[5];
    mb = (float)ws[6];
    k1 = (float)(-(double)up / (D * (double)n));
    k2 = (float)((double)up * cov * vb / (D * D * d * (double)n));  // d == 0 (a constant grid): inf / NaN, as autograd's
  } else if (MODE == kL2) {
    k1 = (float)(2.0 * (double)up / (double)n);
  } else {
    k1 = (float)((double)up / (double)n);
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto one = [&](float x, float y) -> float {
    if (MODE == kCorr) return k1 * (y - mb) + k2 * (x - ma);
    if (MODE == kL2) return k1 * (x - y);
    return k1 * sgn(x - y);
  };
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  float4* g4 = reinterpret_cast<float4*>(grad);
  for (int64_t i = t0; i < n4; i += stride) {
    const float4 x = __ldg(a4 + i), y = __ldg(b4 + i);
    float4 o = accumulate ? g4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    o.x += one(x.x, y.x);
    o.y += one(x.y, y.y);
    o.z += one(x.z, y.z);
    o.w += one(x.w, y.w);
    g4[i] = o;
  }
  for (int64_t i = 4 * n4 + t0; i < n; i += stride) {
    const float t = one(__ldg(a + i), __ldg(b + i));
    grad[i] = accumulate ? (grad[i] + t) : t;
  }
}

// whole float4 groups a vector loop may cover: all of them when every pointer is 16-byte aligned, none otherwise
int64_t vec_groups(int64_t n, const void* p0, const void* p1, const void* p2) {
  const uintptr_t bits = (uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2;
  return (bits & 15) ? 0 : n / 4;
}

int stream_blocks(int64_t n, int ctas_per_sm = 8) {
  const int64_t want = (n + 255) / 256;
  const int64_t cap = 148 * ctas_per_sm;  // <= kMaxCtas; threads stride over the rest
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

cudaError_t launch_tv(const float* grid, const int dims[3], int channels, bool relu, double* workspace, float* loss,
                      const float* upstream, float scale, float* grad, bool accumulate, cudaStream_t stream, int* launches) {
  const int X = dims[0], Y = dims[1], Z = dims[2], C = channels;
  const int ZC = Z * C;
  const double per = (double)C;
  const double n0 = (double)(X - 1) * Y * Z * per, n1 = (double)X * (Y - 1) * Z * per, n2 = (double)X * Y * (Z - 1) * per;
  const float cx = X > 1 ? (float)((double)scale / (3.0 * n0)) : 0.f;
  const float cy = Y > 1 ? (float)((double)scale / (3.0 * n1)) : 0.f;
  const float cz = Z > 1 ? (float)((double)scale / (3.0 * n2)) : 0.f;
  const int64_t total = (int64_t)X * Y * ZC;
  const bool vec = (ZC % 4 == 0) && vec_groups(4, grid, grad, nullptr) != 0;
  const int64_t units = vec ? total / 4 : total;  // float4 groups or floats
  const int n_ctas = stream_blocks(units, vec ? 4 : 8);  // the vector kernel holds 57 registers: 4 resident CTAs per SM
  const int64_t chunk = (((units + n_ctas - 1) / n_ctas + 255) / 256) * 256;  // whole 256-unit strides per CTA
  cudaError_t e = cudaSuccess;
  *launches = 0;
  const bool do_sum = loss != nullptr, do_grad = grad != nullptr;
  const int acc = accumulate ? 1 : 0;
  if (vec && do_grad) {  // strips along y (see tv_strip_kernel)
    const int zmode = C <= 3 ? C : (C % 4 == 0 ? 4 : 0);
    const int GR = ZC / 4;
    const int64_t columns = (int64_t)X * GR;
    const int64_t nseg_want = std::max<int64_t>(1, (int64_t)148 * 4 * 256 / columns);
    int SEG = (int)std::max<int64_t>(4, (Y + nseg_want - 1) / nseg_want);
    if (SEG > Y) SEG = Y;
    if (const char* v = getenv("VOXE_TV_SEG")) { SEG = std::min(Y, std::max(1, atoi(v))); }  // tuning runs: rows per strip
    const int NSEG = (Y + SEG - 1) / SEG;
    const int64_t n_strips = columns * NSEG;
    const int64_t want = (n_strips + 255) / 256;
    const int ctas = (int)(want < kMaxCtas ? want : kMaxCtas);
#define VOXE_TV_STRIP(R_, S_, Z_)                                                                                            \
  tv_strip_kernel<R_, S_, Z_><<<ctas, 256, 0, stream>>>(grid, grad, workspace, X, Y, GR, C, (int64_t)Y * GR, SEG, NSEG, n_strips, \
                                                        cx, cy, cz, upstream, acc, loss, n0, n1, n2)
#define VOXE_TV_STRIP_Z(R_, S_)             \
  switch (zmode) {                          \
    case 1: VOXE_TV_STRIP(R_, S_, 1); break; \
    case 2: VOXE_TV_STRIP(R_, S_, 2); break; \
    case 3: VOXE_TV_STRIP(R_, S_, 3); break; \
    case 4: VOXE_TV_STRIP(R_, S_, 4); break; \
    default: VOXE_TV_STRIP(R_, S_, 0);       \
  }
    if (relu) {
      if (do_sum) { VOXE_TV_STRIP_Z(true, true) } else { VOXE_TV_STRIP_Z(true, false) }
    } else {
      if (do_sum) { VOXE_TV_STRIP_Z(false, true) } else { VOXE_TV_STRIP_Z(false, false) }
    }
#undef VOXE_TV_STRIP_Z
#undef VOXE_TV_STRIP
  } else if (vec) {
    const int zmode = C <= 3 ? C : (C % 4 == 0 ? 4 : 0);
#define VOXE_TV_VEC(R_, S_, G_, Z_)                                                                                          \
  tv_vec_kernel<R_, S_, G_, Z_><<<n_ctas, 256, 0, stream>>>(grid, grad, workspace, X, Y, ZC / 4, C, (int64_t)Y * (ZC / 4), units, \
                                                            chunk, cx, cy, cz, upstream, acc, loss, n0, n1, n2)
#define VOXE_TV_VEC_Z(R_, S_, G_)              \
  switch (zmode) {                             \
    case 1: VOXE_TV_VEC(R_, S_, G_, 1); break; \
    case 2: VOXE_TV_VEC(R_, S_, G_, 2); break; \
    case 3: VOXE_TV_VEC(R_, S_, G_, 3); break; \
    case 4: VOXE_TV_VEC(R_, S_, G_, 4); break; \
    default: VOXE_TV_VEC(R_, S_, G_, 0);       \
  }
    if (relu) {
      if (do_sum && do_grad) { VOXE_TV_VEC_Z(true, true, true) } else if (do_sum) { VOXE_TV_VEC_Z(true, true, false) } else { VOXE_TV_VEC_Z(true, false, true) }
    } else {
      if (do_sum && do_grad) { VOXE_TV_VEC_Z(false, true, true) } else if (do_sum) { VOXE_TV_VEC_Z(false, true, false) } else { VOXE_TV_VEC_Z(false, false, true) }
    }
#undef VOXE_TV_VEC_Z
#undef VOXE_TV_VEC
  } else {
#define VOXE_TV_LAUNCH(R_, S_, G_)                                                                                       \
  tv_kernel<R_, S_, G_><<<n_ctas, 256, 0, stream>>>(grid, grad, workspace, X, Y, ZC, C, (int64_t)Y * ZC, total, chunk, cx, cy, \
                                                    cz, upstream, acc, loss, n0, n1, n2)
    if (relu) {
      if (do_sum && do_grad) VOXE_TV_LAUNCH(true, true, true);
      else if (do_sum) VOXE_TV_LAUNCH(true, true, false);
      else VOXE_TV_LAUNCH(true, false, true);
    } else {
      if (do_sum && do_grad) VOXE_TV_LAUNCH(false, true, true);
      else if (do_sum) VOXE_TV_LAUNCH(false, true, false);
      else VOXE_TV_LAUNCH(false, false, true);
    }
#undef VOXE_TV_LAUNCH
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  ++*launches;
  return e;
}

cudaError_t launch_pair_loss(const float* a, const float* b, int64_t n, int mode, float eps, double* workspace, float* loss,
                             float* corr_grid, cudaStream_t stream, int* launches) {
  *launches = 0;
  cudaError_t e = cudaSuccess;
  const int64_t n4 = vec_groups(n, a, b, nullptr);
  const int blocks = stream_blocks(n4 ? n4 : n);
  if (mode == kCorr) pair_stats_kernel<kCorr><<<blocks, 256, 0, stream>>>(a, b, n, n4, workspace, loss, eps);
  else if (mode == kL2) pair_stats_kernel<kL2><<<blocks, 256, 0, stream>>>(a, b, n, n4, workspace, loss, eps);
  else pair_stats_kernel<kL1><<<blocks, 256, 0, stream>>>(a, b, n, n4, workspace, loss, eps);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  ++*launches;
  if (corr_grid && mode == kCorr) {
    corr_grid_kernel<<<blocks, 256, 0, stream>>>(a, b, n, workspace, corr_grid, eps);
    e = cudaGetLastError();
    if (e == cudaSuccess) ++*launches;
  }
  return e;
}

cudaError_t launch_pair_grad(const float* a, const float* b, int64_t n, int mode, float eps, const double* workspace,
                             const float* upstream, float scale, float* grad, bool accumulate, cudaStream_t stream) {
  const int64_t n4 = vec_groups(n, a, b, grad);
  const int blocks = stream_blocks(n4 ? n4 : n);
  const int acc = accumulate ? 1 : 0;
  if (mode == kCorr) pair_grad_kernel<kCorr><<<blocks, 256, 0, stream>>>(a, b, n, n4, workspace, upstream, scale, eps, grad, acc);
  else if (mode == kL2) pair_grad_kernel<kL2><<<blocks, 256, 0, stream>>>(a, b, n, n4, workspace, upstream, scale, eps, grad, acc);
  else pair_grad_kernel<kL1><<<blocks, 256, 0, stream>>>(a, b, n, n4, workspace, upstream, scale, eps, grad, acc);
  return cudaGetLastError();
}

}  // namespace voxe
