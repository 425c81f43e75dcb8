// voxe_render.cu -- fused forward / backward ray-marching kernels (sm_100a).
//
// Work decomposition (both kernels)
//   A CTA owns `rpc` consecutive rays; thread (r, seg) owns L consecutive samples [seg*L, seg*L+L) of ray r, with
//   r the fastest thread index.  A warp is therefore 32 neighbouring rays at the same depth range (or 16 rays x 2
//   adjacent depth ranges, ...): neighbouring rays of a render hit the same voxel cells, so one warp-wide gather
//   touches a handful of 128-byte lines instead of 32, and the per-thread samples are consecutive along the ray so
//   consecutive cells share corners in L1.
//   Compositing is associative over depth segments: a segment is summarised by (T = prod(1-alpha), sum w*colour,
//   sum w*z, sum w) computed with a local transmittance starting at 1; segments are stitched with an exclusive
//   product scan of T over `seg` (warp shuffles on a shared-memory transpose), which is the only cross-thread
//   communication.  Nothing of size O(R*S) is ever written to memory.
//
// Backward (closed form of SURVEY.md 8.A): phase 1 re-gathers and keeps the L samples' (alpha, local T, delta, z,
//   sigmoid(rgb), post') in registers; the scan additionally yields, per segment, the suffix sum of w*q over all
//   later segments; phase 2 walks the L samples back to front, forms dL/dsigma and dL/draw and scatters them to
//   the 8 corners with 16-byte vector REDs (red.global.add.v4.f32 -> REDG.E.ADD.F32x4).
#include "voxe_device.cuh"
#include "voxe_launch.h"

namespace voxe {

namespace {

constexpr int kMaxThreads = 512;

template <int DEG, int NCOL>
struct Layout {
  static constexpr int K = (DEG + 1) * (DEG + 1);
  static constexpr int F = NCOL * K;            // feature channels
  static constexpr int CV = (F + 1 + 3) / 4;    // float4 vectors per voxel
  static constexpr int DCH = F / 4;             // vector / component holding the density channel
  static constexpr int DCO = F % 4;
};

// One sample: gather the 8 corners, interpolate, SH-contract.  Returns the (pre-activated, interpolated) raw
// density and writes raw colour logits.  `signs` gets one bit per corner: d pre(x)/dx < 0 (abs pre-activation).
template <int DEG, int NCOL>
__device__ __forceinline__ float gather_sample(const KParams& p, const Corners& c, const float (&Y)[(DEG + 1) * (DEG + 1)],
                                               float (&raw)[NCOL], unsigned& signs) {
  using LT = Layout<DEG, NCOL>;
  float fe[LT::CV * 4];
#pragma unroll
  for (int k = 0; k < LT::CV * 4; ++k) fe[k] = 0.f;
  float sig = 0.f;
  signs = 0u;
  float4 v[8][LT::CV];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4* src = p.grid + (size_t)c.idx[q] * LT::CV;
#pragma unroll
    for (int j = 0; j < LT::CV; ++j) v[q][j] = __ldg(src + j);
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float w = c.w[q];
#pragma unroll
    for (int j = 0; j < LT::CV; ++j) {
      fe[4 * j + 0] = fmaf(w, v[q][j].x, fe[4 * j + 0]);
      fe[4 * j + 1] = fmaf(w, v[q][j].y, fe[4 * j + 1]);
      fe[4 * j + 2] = fmaf(w, v[q][j].z, fe[4 * j + 2]);
      fe[4 * j + 3] = fmaf(w, v[q][j].w, fe[4 * j + 3]);
    }
    float dv = f4_get(v[q][LT::DCH], LT::DCO) * p.dscale;  // voxels.py:303-305: pre(density * scale) at the voxel
    if (p.preact == kPreAbs) {
      if (dv < 0.f) signs |= (1u << q);
      dv = fabsf(dv);
    }
    sig = fmaf(w, dv, sig);
  }
#pragma unroll
  for (int ch = 0; ch < NCOL; ++ch) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < LT::K; ++k) a = fmaf(Y[k], fe[ch * LT::K + k], a);
    raw[ch] = a;
  }
  return sig;
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
  return x;
}

// inclusive product scan across lanes
__device__ __forceinline__ float warp_scan_mul(float x, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x *= y;
  }
  return x;
}

// inclusive suffix sum across lanes (lane l gets sum over lanes >= l)
__device__ __forceinline__ float warp_scan_suffix_add(float x, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float y = __shfl_down_sync(0xffffffffu, x, d);
    if (lane + d < 32) x += y;
  }
  return x;
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
template <int DEG, int NCOL, int L>
__global__ void __launch_bounds__(kMaxThreads) render_fwd_kernel(const __grid_constant__ KParams p) {
  using LT = Layout<DEG, NCOL>;
  constexpr int NV = NCOL + 2;  // colour..., depth, acc
  extern __shared__ float smem[];
  const int rpc = p.rpc, nseg = p.nseg, stride = rpc + 1;
  float* sT = smem;                   // [nseg][stride]
  float* sV = smem + nseg * stride;   // [NV][nseg][stride]

  const int r_in = threadIdx.x % rpc, seg = threadIdx.x / rpc;
  const int ray = blockIdx.x * rpc + r_in;
  const bool active = (seg < nseg) && (ray < p.R);

  float Tl = 1.f, V[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) V[k] = 0.f;

  if (active) {
    RayCtx rc;
    load_ray(p, ray, rc);
    const int i0 = seg * L;
    float z[L + 1];
    segment_depths<L>(p, rc, ray, i0, z);
    const bool use_noise = (p.noise_std != 0.f);

    // cheap reject: is any sample of this segment inside the grid?
    unsigned in_mask = 0u;
    float px[L], py[L], pz[L];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      px[j] = __fadd_rn(rc.o[0], __fmul_rn(rc.d[0], z[j]));   // sample.py:67: o + d * z (mul, then add)
      py[j] = __fadd_rn(rc.o[1], __fmul_rn(rc.d[1], z[j]));
      pz[j] = __fadd_rn(rc.o[2], __fmul_rn(rc.d[2], z[j]));
      if (i0 + j < p.S && inside_aabb(p, px[j], py[j], pz[j])) in_mask |= (1u << j);
    }
    if (in_mask != 0u || use_noise) {
      float Y[LT::K];
      const float inv = 1.0f / rc.dnorm;
      sh_basis<DEG>(rc.d[0] * inv, rc.d[1] * inv, rc.d[2] * inv, (p.flags & kDiffuse) != 0, Y);
#pragma unroll
      for (int j = 0; j < L; ++j) {
        const int i = i0 + j;
        if (i >= p.S) break;
        const bool in = (in_mask >> j) & 1u;
        float sigma = 0.f, col[NCOL];
#pragma unroll
        for (int k = 0; k < NCOL; ++k) col[k] = 0.f;   // sigmoid(-1e10) == 0 outside the grid (process.py:80-84)
        if (in) {
          Corners c;
          make_corners(p, px[j], py[j], pz[j], c);
          float raw[NCOL], dpost;
          unsigned signs;
          const float sraw = gather_sample<DEG, NCOL>(p, c, Y, raw, signs);
          sigma = post_act(p.postact, sraw, dpost);
#pragma unroll
          for (int k = 0; k < NCOL; ++k) col[k] = sigmoidf(raw[k]);
        } else if (!use_noise) {
          continue;  // alpha == 0 exactly
        }
        if (use_noise) sigma = fmaf(__ldg(p.noise + (size_t)ray * p.S + i), p.noise_std, sigma);
        const float delta = ((i == p.S - 1) ? kInfinity : __fsub_rn(z[j + 1], z[j])) * rc.dnorm;  // accumulate.py:49-55
        const float alpha = 1.0f - expf(-(sigma * delta));
        const float w = alpha * Tl;
#pragma unroll
        for (int k = 0; k < NCOL; ++k) V[k] = fmaf(w, col[k], V[k]);
        V[NCOL] = fmaf(w, z[j], V[NCOL]);
        V[NCOL + 1] += w;
        Tl *= (1.0f - alpha);
      }
    }
  }
  if (seg < nseg) {
    sT[seg * stride + r_in] = Tl;
#pragma unroll
    for (int k = 0; k < NV; ++k) sV[(k * nseg + seg) * stride + r_in] = V[k];
  }
  __syncthreads();

  // stitch the segments: one warp per ray, lanes = segments
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = warp; r < rpc; r += nwarps) {
    const int ray2 = blockIdx.x * rpc + r;
    if (ray2 >= p.R) break;
    float carry = 1.f, tot[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) tot[k] = 0.f;
    for (int s0 = 0; s0 < nseg; s0 += 32) {
      const int s = s0 + lane;
      const bool ok = s < nseg;
      const float T = ok ? sT[s * stride + r] : 1.f;
      const float inc = warp_scan_mul(T, lane);
      float excl = __shfl_up_sync(0xffffffffu, inc, 1);
      if (lane == 0) excl = 1.f;
      excl *= carry;
      carry *= __shfl_sync(0xffffffffu, inc, 31);
      if (ok) {
#pragma unroll
        for (int k = 0; k < NV; ++k) tot[k] = fmaf(excl, sV[(k * nseg + s) * stride + r], tot[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) tot[k] = warp_sum(tot[k]);
    if (lane == 0) {
      const float acc = tot[NCOL + 1], depth = tot[NCOL];
      const bool white = (p.flags & kWhite) && !(p.flags & kAttn);  // accumulate.py:79-83 / :166
#pragma unroll
      for (int k = 0; k < NCOL; ++k) p.colour[(size_t)ray2 * NCOL + k] = white ? tot[k] + (1.0f - acc) : tot[k];
      p.depth[ray2] = depth;
      p.acc[ray2] = acc;
      if (p.disp != nullptr) {
        const float q = depth / acc;  // NaN when the ray saw nothing, like the reference (accumulate.py:85-88)
        const float m = (q != q) ? q : fmaxf(kZeroPlus, q);
        p.disp[ray2] = 1.0f / m;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
template <int DEG, int NCOL, int L>
__global__ void __launch_bounds__(kMaxThreads) render_bwd_kernel(const __grid_constant__ KParams p) {
  using LT = Layout<DEG, NCOL>;
  constexpr int NV = NCOL + 2;
  extern __shared__ float smem[];
  const int rpc = p.rpc, nseg = p.nseg, stride = rpc + 1;
  float* sT = smem;                          // [nseg][stride]   T of the segment, then T at segment start
  float* sV = smem + nseg * stride;          // [NV][nseg][stride]; plane 0 is reused for the suffix sums
  float* sG = sV + NV * nseg * stride;       // [2][rpc]  effective dL/ddepth, dL/dacc per ray

  const int r_in = threadIdx.x % rpc, seg = threadIdx.x / rpc;
  const int ray = blockIdx.x * rpc + r_in;
  const bool active = (seg < nseg) && (ray < p.R);
  const int i0 = seg * L;

  // per-sample state kept in registers between the two phases
  float s_alpha[L], s_T[L], s_delta[L], s_z[L], s_dpost[L], s_col[L][NCOL];
  unsigned s_signs = 0u;  // 8 bits per sample (L <= 4) or split over two words
  unsigned s_signs_hi = 0u;
  unsigned in_mask = 0u;
  RayCtx rc;
  float Tl = 1.f, V[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) V[k] = 0.f;
  float Y[LT::K];
  const bool use_noise = (p.noise_std != 0.f);

  if (active) {
    load_ray(p, ray, rc);
    float z[L + 1];
    segment_depths<L>(p, rc, ray, i0, z);
    float px[L], py[L], pz[L];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      px[j] = __fadd_rn(rc.o[0], __fmul_rn(rc.d[0], z[j]));
      py[j] = __fadd_rn(rc.o[1], __fmul_rn(rc.d[1], z[j]));
      pz[j] = __fadd_rn(rc.o[2], __fmul_rn(rc.d[2], z[j]));
      if (i0 + j < p.S && inside_aabb(p, px[j], py[j], pz[j])) in_mask |= (1u << j);
      s_z[j] = z[j];
      s_delta[j] = ((i0 + j >= p.S - 1) ? kInfinity : __fsub_rn(z[j + 1], z[j])) * rc.dnorm;
      s_alpha[j] = 0.f;
      s_T[j] = 1.f;
      s_dpost[j] = 0.f;
#pragma unroll
      for (int k = 0; k < NCOL; ++k) s_col[j][k] = 0.f;
    }
    if (in_mask != 0u || use_noise) {
      const float inv = 1.0f / rc.dnorm;
      sh_basis<DEG>(rc.d[0] * inv, rc.d[1] * inv, rc.d[2] * inv, (p.flags & kDiffuse) != 0, Y);
#pragma unroll
      for (int j = 0; j < L; ++j) {
        const int i = i0 + j;
        if (i >= p.S) break;
        const bool in = (in_mask >> j) & 1u;
        float sigma = 0.f;
        if (in) {
          Corners c;
          make_corners(p, px[j], py[j], pz[j], c);
          float raw[NCOL];
          unsigned signs;
          const float sraw = gather_sample<DEG, NCOL>(p, c, Y, raw, signs);
          sigma = post_act(p.postact, sraw, s_dpost[j]);
#pragma unroll
          for (int k = 0; k < NCOL; ++k) s_col[j][k] = sigmoidf(raw[k]);
          if (j < 4) s_signs |= signs << (8 * j); else s_signs_hi |= signs << (8 * (j - 4));
        } else if (!use_noise) {
          continue;
        }
        if (use_noise) sigma = fmaf(__ldg(p.noise + (size_t)ray * p.S + i), p.noise_std, sigma);
        const float alpha = 1.0f - expf(-(sigma * s_delta[j]));
        const float w = alpha * Tl;
        s_alpha[j] = alpha;
        s_T[j] = Tl;
#pragma unroll
        for (int k = 0; k < NCOL; ++k) V[k] = fmaf(w, s_col[j][k], V[k]);
        V[NCOL] = fmaf(w, s_z[j], V[NCOL]);
        V[NCOL + 1] += w;
        Tl *= (1.0f - alpha);
      }
    }
  }
  if (seg < nseg) {
    sT[seg * stride + r_in] = Tl;
#pragma unroll
    for (int k = 0; k < NV; ++k) sV[(k * nseg + seg) * stride + r_in] = V[k];
  }
  __syncthreads();

  // stitch: T at segment start, totals, effective output gradients, suffix sums of w*q over later segments
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = warp; r < rpc; r += nwarps) {
    const int ray2 = blockIdx.x * rpc + r;
    if (ray2 >= p.R) break;
    float carry = 1.f, totD = 0.f, totA = 0.f;
    for (int s0 = 0; s0 < nseg; s0 += 32) {
      const int s = s0 + lane;
      const bool ok = s < nseg;
      const float T = ok ? sT[s * stride + r] : 1.f;
      const float inc = warp_scan_mul(T, lane);
      float excl = __shfl_up_sync(0xffffffffu, inc, 1);
      if (lane == 0) excl = 1.f;
      excl *= carry;
      carry *= __shfl_sync(0xffffffffu, inc, 31);
      if (ok) {
        sT[s * stride + r] = excl;
        totD = fmaf(excl, sV[(NCOL * nseg + s) * stride + r], totD);
        totA = fmaf(excl, sV[((NCOL + 1) * nseg + s) * stride + r], totA);
      }
    }
    totD = warp_sum(totD);
    totA = warp_sum(totA);
    float gc[NCOL], gsum = 0.f;
#pragma unroll
    for (int k = 0; k < NCOL; ++k) {
      gc[k] = __ldg(p.g_colour + (size_t)ray2 * NCOL + k);
      gsum += gc[k];
    }
    float gd = p.g_depth ? __ldg(p.g_depth + ray2) : 0.f;
    float ga = p.g_acc ? __ldg(p.g_acc + ray2) : 0.f;
    if ((p.flags & kWhite) && !(p.flags & kAttn)) ga -= gsum;  // colour += 1 - acc
    if (p.g_disp) {
      const float gq = __ldg(p.g_disp + ray2);
      const float q = totD / totA;
      if (gq != 0.f && q > kZeroPlus) {  // disparity = acc / depth on this branch of the max()
        ga += gq / totD;
        gd -= gq * totA / (totD * totD);
      }
    }
    float carry_q = 0.f;
    for (int s0 = ((nseg - 1) / 32) * 32; s0 >= 0; s0 -= 32) {
      const int s = s0 + lane;
      const bool ok = s < nseg;
      float q = 0.f;
      if (ok) {
#pragma unroll
        for (int k = 0; k < NCOL; ++k) q = fmaf(gc[k], sV[(k * nseg + s) * stride + r], q);
        q = fmaf(gd, sV[(NCOL * nseg + s) * stride + r], q);
        q = fmaf(ga, sV[((NCOL + 1) * nseg + s) * stride + r], q);
        q *= sT[s * stride + r];
      }
      const float inc = warp_scan_suffix_add(q, lane);
      float excl = __shfl_down_sync(0xffffffffu, inc, 1);
      if (lane == 31) excl = 0.f;
      excl += carry_q;
      carry_q += __shfl_sync(0xffffffffu, inc, 0);
      if (ok) sV[s * stride + r] = excl;  // plane 0 <- suffix (own q was read above by this same lane)
    }
    if (lane == 0) {
      sG[r] = gd;
      sG[rpc + r] = ga;
    }
  }
  __syncthreads();

  if (!active || in_mask == 0u) return;

  // phase 2: back to front over this thread's samples
  const float Tstart = sT[seg * stride + r_in];
  float suffix = sV[seg * stride + r_in];
  const float gd = sG[r_in], ga = sG[rpc + r_in];
  float gc[NCOL];
#pragma unroll
  for (int k = 0; k < NCOL; ++k) gc[k] = __ldg(p.g_colour + (size_t)ray * NCOL + k);

#pragma unroll
  for (int j = L - 1; j >= 0; --j) {
    if (i0 + j >= p.S) continue;
    const float alpha = s_alpha[j];
    const float T = Tstart * s_T[j];
    const float w = alpha * T;
    float q = fmaf(gd, s_z[j], ga);
#pragma unroll
    for (int k = 0; k < NCOL; ++k) q = fmaf(gc[k], s_col[j][k], q);
    // dL/dsigma_i = delta_i * (T_{i+1} * q_i - sum_{j>i} w_j q_j)
    const float dsigma = s_delta[j] * (T * (1.0f - alpha) * q - suffix);
    suffix = fmaf(w, q, suffix);
    if (!((in_mask >> j) & 1u)) continue;  // masked samples pass no gradient (process.py:80-91)
    const float dsraw = dsigma * s_dpost[j] * p.dscale;
    float draw[NCOL];
    bool any = (dsraw != 0.f);
#pragma unroll
    for (int k = 0; k < NCOL; ++k) {
      const float sc = s_col[j][k];
      draw[k] = gc[k] * w * sc * (1.0f - sc);
      any |= (draw[k] != 0.f);
    }
    if (!any) continue;
    // gradient w.r.t. the interpolated channel vector
    float gfe[LT::CV * 4];
#pragma unroll
    for (int k = 0; k < LT::CV * 4; ++k) gfe[k] = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCOL; ++ch)
#pragma unroll
      for (int k = 0; k < LT::K; ++k) gfe[ch * LT::K + k] = draw[ch] * Y[k];
    const unsigned signs = (j < 4) ? (s_signs >> (8 * j)) : (s_signs_hi >> (8 * (j - 4)));
    Corners c;
    make_corners(p, __fadd_rn(rc.o[0], __fmul_rn(rc.d[0], s_z[j])), __fadd_rn(rc.o[1], __fmul_rn(rc.d[1], s_z[j])),
                 __fadd_rn(rc.o[2], __fmul_rn(rc.d[2], s_z[j])), c);
#pragma unroll
    for (int qn = 0; qn < 8; ++qn) {
      const float wq = c.w[qn];
      if (wq == 0.f) continue;
      float4* dst = p.grad + (size_t)c.idx[qn] * LT::CV;
      const float gdens = ((signs >> qn) & 1u) ? -dsraw : dsraw;
#pragma unroll
      for (int jv = 0; jv < LT::CV; ++jv) {
        float a = gfe[4 * jv + 0], b = gfe[4 * jv + 1], cc = gfe[4 * jv + 2], d = gfe[4 * jv + 3];
        if (jv == LT::DCH) {
          if (LT::DCO == 0) a = gdens; else if (LT::DCO == 1) b = gdens; else if (LT::DCO == 2) cc = gdens; else d = gdens;
        }
        red_add_v4(dst + jv, wq * a, wq * b, wq * cc, wq * d);
      }
    }
  }
}

template <int DEG, int NCOL, int L>
cudaError_t launch_pair(const KParams& p, bool backward, cudaStream_t stream) {
  const int threads = ((p.rpc * p.nseg + 31) / 32) * 32;
  const int blocks = (p.R + p.rpc - 1) / p.rpc;
  const int nv = NCOL + 2;
  size_t smem = sizeof(float) * ((size_t)(1 + nv) * p.nseg * (p.rpc + 1) + (backward ? 2 * p.rpc : 0));
  if (backward) {
    auto k = render_bwd_kernel<DEG, NCOL, L>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<blocks, threads, smem, stream>>>(p);
  } else {
    auto k = render_fwd_kernel<DEG, NCOL, L>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<blocks, threads, smem, stream>>>(p);
  }
  return cudaGetLastError();
}

template <int L>
cudaError_t dispatch_deg(const KParams& p, int deg, int ncol, bool backward, cudaStream_t stream) {
  if (ncol == 1) {
    if (deg == 0) return launch_pair<0, 1, L>(p, backward, stream);
    return cudaErrorInvalidValue;
  }
  switch (deg) {
    case 0: return launch_pair<0, 3, L>(p, backward, stream);
    case 1: return launch_pair<1, 3, L>(p, backward, stream);
    case 2: return launch_pair<2, 3, L>(p, backward, stream);
    case 3: return launch_pair<3, 3, L>(p, backward, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_render(const KParams& p, int deg, int ncol, int samples_per_thread, bool backward,
                          cudaStream_t stream) {
  if (samples_per_thread == 4) return dispatch_deg<4>(p, deg, ncol, backward, stream);
  if (samples_per_thread == 8) return dispatch_deg<8>(p, deg, ncol, backward, stream);
  return cudaErrorInvalidValue;
}

int max_threads_per_cta() { return kMaxThreads; }

}  // namespace voxe
