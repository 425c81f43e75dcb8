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
// Forward: streams over its samples (gather 8 corners -> interpolate -> SH -> alpha), then the stitch.  When a
//   `saved` workspace is given (a backward will follow) it also stores
//     * per in-grid sample the tone-mapped colour and the interpolated raw density -- one 16-byte vector, laid out
//       [slot][ray] so that the rays of a warp write one 128-byte line per depth segment -- and
//     * per (ray, segment) the transmittance at the segment start and the segment's local sums ((NCOL+3) floats).
//
// Backward (closed form of SURVEY.md 8.A): reads the segment summaries, turns them into "sum of w*q over everything
//   behind this segment" with one suffix scan, and then streams over its samples ONCE: reload the sample vector (one
//   LDG.128 instead of re-gathering 8 x CV corner vectors), recompute alpha / T / w, form
//   dL/dsigma_i = delta_i (T_{i+1} q_i - sum_{j>i} w_j q_j) and dL/draw, and scatter to the 8 corners with 16-byte
//   vector REDs (red.global.add.v4.f32 -> REDG.E.ADD.F32x4).  The workspace is 16 bytes per sample, against the ~140
//   bytes per sample autograd keeps for the reference.
#include "voxe_device.cuh"
#include <stdlib.h>

#include "voxe_launch.h"

#ifndef VOXE_UNROLL
#define VOXE_UNROLL 2
#endif

namespace voxe {

namespace {

constexpr int kSampleUnroll = VOXE_UNROLL;  // unroll factor of the per-thread sample loop

template <int DEG, int NCOL>
struct Layout {
  static constexpr int K = (DEG + 1) * (DEG + 1);
  static constexpr int F = NCOL * K;            // feature channels
  static constexpr int CV = (F + 1 + 3) / 4;    // float4 vectors per voxel
  static constexpr int DCH = F / 4;             // vector / component holding the density channel
  static constexpr int DCO = F % 4;
  static constexpr int NV = NCOL + 2;           // per-segment sums: colour..., depth, acc
  static constexpr int SV = NCOL + 3;           // saved floats per (ray, segment): T_start + NV sums
};

// One sample: gather the 8 corners, interpolate, SH-contract.  Returns the (pre-activated, interpolated) raw
// density and writes raw colour logits.  `signs` gets one bit per corner: d pre(x)/dx < 0 (abs pre-activation).
template <int DEG, int NCOL, class SP = SpecDynamic>
__device__ __forceinline__ float gather_sample(const KParams& p, const Corners& c, const float (&Y)[(DEG + 1) * (DEG + 1)],
                                               float (&raw)[NCOL], unsigned& signs) {
  using LT = Layout<DEG, NCOL>;
  float fe[LT::CV * 4];
#pragma unroll
  for (int k = 0; k < LT::CV * 4; ++k) fe[k] = 0.f;
  float sig = 0.f;
  signs = 0u;
  // Corners are gathered kGroup at a time: all loads of a group are issued before its first multiply-add.  Single-vector
  // voxels (SH-0) take all 8 corners at once (8 float4 in flight); multi-vector voxels (SH >= 1: CV = 4 / 7 / 13 vectors
  // per corner) take them in pairs -- 8 x CV vectors do not fit the register file next to the CV * 4 accumulators, and
  // letting the compiler find that out cost 64-88 bytes of local-memory spills per thread at 128 registers.
  constexpr int kGroup = LT::CV == 1 ? 8 : (LT::CV <= 4 ? 4 : 2);
  float4 v[8][LT::CV];  // only kGroup rows are live at a time
#pragma unroll
  for (int q0 = 0; q0 < 8; q0 += kGroup) {
#pragma unroll
    for (int q = q0; q < q0 + kGroup; ++q) {
      const float4* src = p.grid + (size_t)c.idx[q] * LT::CV;
#pragma unroll
      for (int j = 0; j < LT::CV; ++j) v[q][j] = __ldg(src + j);
    }
#pragma unroll
    for (int q = q0; q < q0 + kGroup; ++q) {
      const float w = c.w[q];
#pragma unroll
      for (int j = 0; j < LT::CV; ++j) {
        fe[4 * j + 0] = fmaf(w, v[q][j].x, fe[4 * j + 0]);
        fe[4 * j + 1] = fmaf(w, v[q][j].y, fe[4 * j + 1]);
        fe[4 * j + 2] = fmaf(w, v[q][j].z, fe[4 * j + 2]);
        fe[4 * j + 3] = fmaf(w, v[q][j].w, fe[4 * j + 3]);
      }
    }
  }
  // voxels.py:303-305: pre(density * scale) is applied at the voxels, then interpolated
  if (SP::pre_abs(p.preact)) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float dv = f4_get(v[q][LT::DCH], LT::DCO);
      if (dv * p.dscale < 0.f) signs |= (1u << q);
      sig = fmaf(c.w[q], fabsf(dv), sig);
    }
    sig *= fabsf(p.dscale);
  } else {
    sig = fe[LT::F] * p.dscale;  // identity: the blend of the density channel, scaled once
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

// Layout of the `saved` workspace: [nseg * L][R] sample vectors (float4), then [(NCOL+3)][nseg][R] segment summaries.
__device__ __forceinline__ float4* sample_slots(const KParams& p, int seg, int ray) {
  return reinterpret_cast<float4*>(p.saved) + (size_t)seg * p.L * p.R + ray;  // + j * R for the thread's j-th sample
}
__device__ __forceinline__ float* summaries(const KParams& p) { return p.saved + (size_t)4 * p.nseg * p.L * p.R; }

// Sign bits of the 8 corner densities (abs pre-activation only): d|x|/dx < 0.
template <int DEG, int NCOL>
__device__ __forceinline__ unsigned corner_signs(const KParams& p, const Corners& c) {
  using LT = Layout<DEG, NCOL>;
  unsigned signs = 0u;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float dv = f4_get(__ldg(p.grid + (size_t)c.idx[q] * LT::CV + LT::DCH), LT::DCO);
    if (dv * p.dscale < 0.f) signs |= (1u << q);
  }
  return signs;
}

// Multi-vector voxels (SH degree >= 1): a corner's gradient is CV consecutive 16-byte vectors.  Instead of CV vector
// REDs per lane, the thread stages the CV vectors in its shared-memory slot and issues ONE TMA reduction
// (cp.reduce.async.bulk ... .add.f32, CV * 16 bytes) -- measured on B200 (tools/bulkred_micro.cu): 44.8 G voxel-adds/s
// against 26.7 G/s for 7 x red.v4.f32 per lane (L2-resident volume), 15.7 against 8.1 G/s for a DRAM-sized one.
// Slots are an odd number of vectors apart so that a quarter-warp's 16-byte stores hit distinct banks.
template <int CV>
struct BulkStage {
  static constexpr int kSlot = CV | 1;
};

__host__ __device__ __forceinline__ int bwd_stage_offset_floats(int nv, int nseg, int rpc) {
  const int floats = (1 + nv) * nseg * (rpc + 1) + 2 * rpc;
  return (floats + 3) & ~3;  // 16-byte aligned
}

__device__ __forceinline__ void bulk_reduce_add(float4* gdst, const float4* ssrc, int bytes) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the staged values were written through the generic proxy
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// Samples [i0, i1) of depth segment `seg` of one ray: the ray's in-grid index range (sample_range) split evenly over
// the nseg threads of the ray, so every thread of a ray streams the same number of (almost always in-grid) samples
// whatever part of [near, far] the grid occupies; at most L = ceil(S / nseg) of them.  Forward and backward evaluate
// this identically.
__device__ __forceinline__ void thread_samples(const KParams& p, const RayCtx& rc, int seg, int& i0, int& i1) {
  int a, b;
  sample_range(p, rc, a, b);
  const int per = (b - a + p.nseg - 1) / p.nseg;
  i0 = a + seg * per;
  i1 = min(i0 + per, b);
}

// Which group of `rpc` consecutive rays a CTA renders.  The block scheduler hands CTA i and CTA i + (number of SMs) to the
// same SM, and on a scan-line batch (rows of W pixels) those two are 148 * rpc rays apart -- almost a whole number of rows at
// the benchmark's W = 400 -- i.e. at the same image column: an SM ends up with only centre-of-image groups (long in-grid
// ranges) or only edge groups (short or empty ones); ncu shows a 2x spread of executed instructions between SMs and 35 % of
// the launch spent waiting for the slowest one.  Rotating every round of `round` CTAs by a further `rot` groups keeps
// neighbouring CTAs neighbours (they share cache lines) but gives each SM a mix of columns.  rot == 0: identity.
__device__ __forceinline__ int ray_group(const KParams& p) {
  const int i = (int)blockIdx.x;
  if (p.group_rot == 0) return i;
  const int r = i / p.group_round, base = r * p.group_round;
  const int n = min(p.group_round, (int)gridDim.x - base);  // the last round may be partial
  return base + (i - base + r * p.group_rot) % n;
}

template <int REGCAP>
struct Bounds {
  // register budget of the variant = 65536 / (kThreads * kMinBlocks): 64 | 80 (85) | 96 (102) | 128
  static constexpr int kThreads = (REGCAP <= 64) ? 1024 : (REGCAP >= 128 ? 512 : 128);
  static constexpr int kMinBlocks = (REGCAP == 80) ? 6 : (REGCAP == 96 ? 5 : 1);
};

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
template <int DEG, int NCOL, int REGCAP, class SP>
__global__ void __launch_bounds__(Bounds<REGCAP>::kThreads, Bounds<REGCAP>::kMinBlocks) render_fwd_kernel(const __grid_constant__ KParams p) {
  using LT = Layout<DEG, NCOL>;
  constexpr int NV = LT::NV;
  extern __shared__ float smem[];
  const int rpc = p.rpc, nseg = p.nseg, stride = rpc + 1;
  float* sT = smem;                   // [nseg][stride]
  float* sV = smem + nseg * stride;   // [NV][nseg][stride]

  const int r_in = threadIdx.x % rpc, seg = threadIdx.x / rpc;
  const int group = ray_group(p);
  const int ray = group * rpc + r_in;
  const bool active = (seg < nseg) && (ray < p.R);
  pdl_wait();  // everything below may read what the previous kernel of the stream wrote (rays, the packed volume, ...)

  float Tl = 1.f, V[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) V[k] = 0.f;

  if (active) {
    RayCtx rc;
    load_ray(p, ray, rc);
    int i0, i1;
    thread_samples(p, rc, seg, i0, i1);
    JitterSource u_row;
    u_row.init(p, ray);
    float4* samples = p.saved ? sample_slots(p, seg, ray) : nullptr;
    DepthWalker zw;
    if (i0 < i1) zw.init<SP>(p, rc, u_row, i0);
    const bool use_noise = SP::noise(p.noise_std);
    float Y[LT::K];
    const float inv = 1.0f / rc.dnorm;
    sh_basis<DEG>(rc.d[0] * inv, rc.d[1] * inv, rc.d[2] * inv, (p.flags & kDiffuse) != 0, Y);
    // The body is straight-line code (no early exits): a batch keeps only ~13 warps on an SM, so the kernel is bound by
    // the dependent-instruction latency of one sample; unrolled by two, the loads and arithmetic of two consecutive
    // samples interleave.  Samples outside the box are rare here (thread_samples) and are masked, not skipped: their
    // corner addresses are clamped into the volume and their sigma / colour are forced to 0 (process.py:80-91).
    // Multi-vector voxels are not unrolled: two samples' accumulators and corner vectors do not fit 128 registers.
    constexpr int kFwdUnroll = LT::CV == 1 ? kSampleUnroll : 1;
#pragma unroll(kFwdUnroll)
    for (int i = i0; i < i1; ++i, zw.advance<SP>(p, rc, u_row, i - 1)) {
      const float zi = zw.cur;
      const float px = __fadd_rn(rc.o[0], __fmul_rn(rc.d[0], zi));   // sample.py:67: o + d * z (mul, then add)
      const float py = __fadd_rn(rc.o[1], __fmul_rn(rc.d[1], zi));
      const float pz = __fadd_rn(rc.o[2], __fmul_rn(rc.d[2], zi));
      const bool in = inside_aabb(p, px, py, pz);
      Corners c;
      make_corners(p, px, py, pz, c);
      float raw[NCOL], dpost, col[NCOL];
      unsigned signs;
      const float sraw = gather_sample<DEG, NCOL, SP>(p, c, Y, raw, signs);
      float sigma = in ? post_act(SP::postact(p.postact), sraw, dpost) : 0.f;
#pragma unroll
      for (int k = 0; k < NCOL; ++k) col[k] = in ? sigmoid_fast(raw[k]) : 0.f;   // sigmoid(-1e10) == 0 outside the grid
      if (samples != nullptr) {  // what the backward needs of this sample: colour(s) and the raw density (every slot
        // of the thread's range is written, so the backward's one-ahead prefetch never reads uninitialised memory)
        float4 sv = make_float4(col[0], 0.f, 0.f, in ? sraw : 0.f);
        if (NCOL > 1) sv.y = col[1];
        if (NCOL > 2) sv.z = col[2];
        samples[(size_t)(i - i0) * p.R] = sv;
      }
      if (use_noise) sigma = fmaf(__ldg(p.noise + (size_t)ray * p.S + i), p.noise_std, sigma);
      const float delta = ((i == p.S - 1) ? kInfinity : __fsub_rn(zw.next, zi)) * rc.dnorm;  // accumulate.py:49-55
      const float alpha = 1.0f - exp_fast(-(sigma * delta));
      const float w = alpha * Tl;
#pragma unroll
      for (int k = 0; k < NCOL; ++k) V[k] = fmaf(w, col[k], V[k]);
      V[NCOL] = fmaf(w, zi, V[NCOL]);
      V[NCOL + 1] += w;
      Tl *= (1.0f - alpha);
    }
  }
  pdl_launch_dependents();  // the sample loop is over: the next kernel's CTAs may take the slots this grid frees from here on
  if (seg < nseg) {
    sT[seg * stride + r_in] = Tl;
#pragma unroll
    for (int k = 0; k < NV; ++k) sV[(k * nseg + seg) * stride + r_in] = V[k];
  }
  __syncthreads();

  // stitch the segments: one warp per ray, lanes = segments
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = warp; r < rpc; r += nwarps) {
    const int ray2 = group * rpc + r;
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
        if (p.saved) sT[s * stride + r] = excl;  // T at segment start, written out below
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
  if (p.saved) {  // segment summaries for the backward: saved[k][seg][ray], coalesced over rays
    __syncthreads();
    if (active) {
      const size_t plane = (size_t)nseg * p.R;
      float* dst = summaries(p) + (size_t)seg * p.R + ray;
      dst[0] = sT[seg * stride + r_in];
#pragma unroll
      for (int k = 0; k < NV; ++k) dst[(k + 1) * plane] = V[k];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
template <int DEG, int NCOL, int REGCAP, class SP>
__global__ void __launch_bounds__(Bounds<REGCAP>::kThreads, Bounds<REGCAP>::kMinBlocks) render_bwd_kernel(const __grid_constant__ KParams p) {
  using LT = Layout<DEG, NCOL>;
  constexpr int NV = LT::NV;
  extern __shared__ float smem[];
  const int rpc = p.rpc, nseg = p.nseg, stride = rpc + 1;
  float* sT = smem;                          // [nseg][stride]   T at segment start (from the forward)
  float* sV = smem + nseg * stride;          // [NV][nseg][stride] local sums; plane 0 becomes q_s, plane 1 the suffix
  float* sG = sV + NV * nseg * stride;       // [2][rpc]  effective dL/ddepth, dL/dacc per ray
  // [2][threads][kStageSlot] float4: double-buffered staging of one voxel's gradient vector per thread (CV > 1 only)
  float4* sStage = reinterpret_cast<float4*>(smem + bwd_stage_offset_floats(NV, nseg, rpc));

  const int r_in = threadIdx.x % rpc, seg = threadIdx.x / rpc;
  const int group = ray_group(p);
  const int ray = group * rpc + r_in;
  const bool active = (seg < nseg) && (ray < p.R);
  pdl_wait();  // the forward's workspace, the upstream gradients

  // load the forward's segment summaries (coalesced over rays) and transpose through shared memory
  if (seg < nseg) {
    const size_t plane = (size_t)nseg * p.R;
    const float* src = summaries(p) + (size_t)seg * p.R + ray;
    sT[seg * stride + r_in] = active ? __ldg(src) : 1.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) sV[(k * nseg + seg) * stride + r_in] = active ? __ldg(src + (k + 1) * plane) : 0.f;
  }
  __syncthreads();

  // per ray: totals, effective output gradients, q_s = T_start * <g, sums_s>, and its exclusive suffix sum
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = warp; r < rpc; r += nwarps) {
    const int ray2 = group * rpc + r;
    if (ray2 >= p.R) break;
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
      if (gq != 0.f) {
        float totD = 0.f, totA = 0.f;
        for (int s = lane; s < nseg; s += 32) {
          const float ts = sT[s * stride + r];
          totD = fmaf(ts, sV[(NCOL * nseg + s) * stride + r], totD);
          totA = fmaf(ts, sV[((NCOL + 1) * nseg + s) * stride + r], totA);
        }
        totD = warp_sum(totD);
        totA = warp_sum(totA);
        if (totD / totA > kZeroPlus) {  // disparity = acc / depth on this branch of the max()
          ga += gq / totD;
          gd -= gq * totA / (totD * totD);
        }
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
      if (ok) {
        sV[s * stride + r] = q;                    // plane 0 <- sum of w*q inside segment s
        sV[(nseg + s) * stride + r] = excl;        // plane 1 <- sum of w*q behind segment s
      }
    }
    if (lane == 0) {
      sG[r] = gd;
      sG[rpc + r] = ga;
    }
  }
  __syncthreads();
  // Unlike the forward, the backward lets the next kernel's CTAs in from here: its threads return one by one as their
  // segments finish (there is no common point after the sample loop), and the hand-over kernel that follows is tiny.
  pdl_launch_dependents();
  if (!active) return;

  // stream over this thread's samples, front to back
  float Tcur = sT[seg * stride + r_in];
  const float q_seg = sV[seg * stride + r_in];
  const float q_behind = sV[(nseg + seg) * stride + r_in];
  if (Tcur == 0.f && q_seg == 0.f && q_behind == 0.f) return;  // fully occluded and nothing behind: all gradients are 0
  const float gd = sG[r_in], ga = sG[rpc + r_in];
  float gc[NCOL];
#pragma unroll
  for (int k = 0; k < NCOL; ++k) gc[k] = __ldg(p.g_colour + (size_t)ray * NCOL + k);

  RayCtx rc;
  load_ray(p, ray, rc);
  int i0, i1;
  thread_samples(p, rc, seg, i0, i1);
  if (i0 >= i1) return;
  JitterSource u_row;
  u_row.init(p, ray);
  const float4* samples = sample_slots(p, seg, ray);
  DepthWalker zw;
  zw.init<SP>(p, rc, u_row, i0);
  const bool use_noise = SP::noise(p.noise_std);
  float Y[LT::K];
  const float inv = 1.0f / rc.dnorm;
  sh_basis<DEG>(rc.d[0] * inv, rc.d[1] * inv, rc.d[2] * inv, (p.flags & kDiffuse) != 0, Y);
  float prefix = 0.f;  // sum of w*q over this segment's samples up to and including the current one
  int stage_buf = 0;
  unsigned n_in = 0, n_scatter = 0, n_cell_leader = 0, n_corner_leader = 0;  // p.stats only

  // The sample vectors are fetched TWO iterations ahead of their use: with ReLU every second iteration is a short one
  // (no scatter), and ncu showed the move out of a one-ahead prefetch as the kernel's top stall (long scoreboard).
  float4 sv_next = __ldg(samples), sv_next2 = sv_next;
  if (i0 + 1 < i1) sv_next2 = __ldg(samples + (size_t)p.R);
#pragma unroll(kSampleUnroll)
  for (int i = i0; i < i1; ++i, zw.advance<SP>(p, rc, u_row, i - 1)) {
    const float4 sv = sv_next;
    sv_next = sv_next2;
    if (i + 2 < i1) sv_next2 = __ldg(samples + (size_t)(i + 2 - i0) * p.R);
    const float zi = zw.cur;
    const float px = __fadd_rn(rc.o[0], __fmul_rn(rc.d[0], zi));
    const float py = __fadd_rn(rc.o[1], __fmul_rn(rc.d[1], zi));
    const float pz = __fadd_rn(rc.o[2], __fmul_rn(rc.d[2], zi));
    const bool in = inside_aabb(p, px, py, pz);
    if (!in && !use_noise) continue;
    float sigma = 0.f, dpost = 0.f, col[NCOL];
#pragma unroll
    for (int k = 0; k < NCOL; ++k) col[k] = 0.f;
    if (in) {
      col[0] = sv.x;  // colour(s) + raw density stored by the forward
      if (NCOL > 1) col[1] = sv.y;
      if (NCOL > 2) col[2] = sv.z;
      sigma = post_act(SP::postact(p.postact), sv.w, dpost);
    }
    if (use_noise) sigma = fmaf(__ldg(p.noise + (size_t)ray * p.S + i), p.noise_std, sigma);
    const bool last = (i == p.S - 1);
    const float delta = (last ? kInfinity : __fsub_rn(zw.next, zi)) * rc.dnorm;
    const float alpha = 1.0f - exp_fast(-(sigma * delta));
    const float w = alpha * Tcur;
    float q = fmaf(gd, zi, ga);
#pragma unroll
    for (int k = 0; k < NCOL; ++k) q = fmaf(gc[k], col[k], q);
    prefix = fmaf(w, q, prefix);
    const float Tnext = Tcur * (1.0f - alpha);
    // dL/dsigma_i = delta_i * (T_{i+1} q_i - sum_{j>i} w_j q_j); the sum behind the very last sample is exactly 0
    const float behind = last ? 0.f : (q_behind + (q_seg - prefix));
    const float dsigma = delta * (Tnext * q - behind);
    Tcur = Tnext;
    if (!in) continue;  // masked samples pass no gradient
    ++n_in;
    const float dsraw = dsigma * dpost * p.dscale;
    float draw[NCOL];
    bool any = (dsraw != 0.f);
#pragma unroll
    for (int k = 0; k < NCOL; ++k) {
      draw[k] = gc[k] * w * col[k] * (1.0f - col[k]);
      any |= (draw[k] != 0.f);
    }
    if (!any) continue;
    ++n_scatter;
    Corners c;
    make_corners(p, px, py, pz, c);
    if (p.touched != nullptr) p.touched[c.idx[0] >> 3] = (unsigned char)p.touch_tag;  // brick of corner 0 (8 slots per brick)
    if (p.stats != nullptr) {
      // measurement only: how many of the lanes that scatter in this warp instruction hit a cell / a voxel no lower lane
      // hits -- the number of REDs a perfect intra-warp merge (match.any + shuffle reduction) would still issue
      const unsigned active = __activemask();
      const int lane_id = threadIdx.x & 31;
      n_cell_leader += (__ffs(__match_any_sync(active, c.idx[0])) - 1 == lane_id);
#pragma unroll
      for (int qn = 0; qn < 8; ++qn) n_corner_leader += (__ffs(__match_any_sync(active, c.idx[qn])) - 1 == lane_id);
    }
    const unsigned signs = SP::pre_abs(p.preact) ? corner_signs<DEG, NCOL>(p, c) : 0u;
    float gfe[LT::CV * 4];
#pragma unroll
    for (int k = 0; k < LT::CV * 4; ++k) gfe[k] = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCOL; ++ch)
#pragma unroll
      for (int k = 0; k < LT::K; ++k) gfe[ch * LT::K + k] = draw[ch] * Y[k];
#pragma unroll
    for (int qn = 0; qn < 8; ++qn) {
      const float wq = c.w[qn];  // corners beyond the grid land in the zero apron; unpack drops what they receive
      float4* dst = p.grad + (size_t)c.idx[qn] * LT::CV;
      const float gdens = ((signs >> qn) & 1u) ? -dsraw : dsraw;
      float4* slot = nullptr;
      if constexpr (LT::CV > 1) {
        slot = sStage + ((size_t)(stage_buf * blockDim.x) + threadIdx.x) * BulkStage<LT::CV>::kSlot;
        stage_buf ^= 1;
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the reduction issued two corners ago has read this slot
      }
#pragma unroll
      for (int jv = 0; jv < LT::CV; ++jv) {
        float a = gfe[4 * jv + 0], b = gfe[4 * jv + 1], cc = gfe[4 * jv + 2], d = gfe[4 * jv + 3];
        if (jv == LT::DCH) {
          if (LT::DCO == 0) a = gdens; else if (LT::DCO == 1) b = gdens; else if (LT::DCO == 2) cc = gdens; else d = gdens;
        }
        if constexpr (LT::CV > 1) slot[jv] = make_float4(wq * a, wq * b, wq * cc, wq * d);
        else red_add_v4(dst + jv, wq * a, wq * b, wq * cc, wq * d);
      }
      if constexpr (LT::CV > 1) bulk_reduce_add(dst, slot, LT::CV * 16);
    }
  }
  if constexpr (LT::CV > 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // staging slots die with the CTA
  if (p.stats != nullptr) {  // measurement runs only (bench.py bills the bytes the kernel really moves)
    atomicAdd(p.stats, (unsigned long long)n_in);
    atomicAdd(p.stats + 1, (unsigned long long)n_scatter);
    atomicAdd(p.stats + 2, (unsigned long long)n_cell_leader);
    atomicAdd(p.stats + 3, (unsigned long long)n_corner_leader);
  }
}

// ---------------------------------------------------------------------------------------------------------
// whole-camera inference
// ---------------------------------------------------------------------------------------------------------
// Forward-only render of a pinhole camera: one thread per pixel.  The ray is generated in the kernel from the pose and
// the intrinsics (cast_rays, misc.py:12-50: pixel centres, dir = ((x+.5-W/2)/f, -(y+.5-H/2)/f, -1), d = R dir, not
// normalised), so no ray tensors exist; the thread walks the ray's in-grid sample range front to back with a running
// transmittance and may stop once T < min_transmittance (the remaining samples can change a pixel by less than that;
// 0 disables it).  With a whole camera in one launch there are enough rays to fill the machine without splitting a ray
// over threads, and the 32 lanes of a warp are 32 neighbouring pixels at the same depth index -- the most coherent gather
// this volume layout can get.  Shares every device function with the training kernels; never used when a backward follows.
template <int DEG, int NCOL>
__global__ void __launch_bounds__(128, 5) render_camera_kernel(const __grid_constant__ KParams p, const __grid_constant__ CameraParams cam) {
  using LT = Layout<DEG, NCOL>;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.R) return;
  const long long pixel = cam.first_pixel + t;
  const int row = (int)(pixel / max(cam.W, 1)), col = (int)(pixel - (long long)row * cam.W);
  RayCtx rc;
  if (p.rays_o != nullptr) {  // caller-supplied rays (bit-identical geometry to the training kernels), see voxe_render_infer
    load_ray(p, (int)t, rc);
  } else {
    const float dx = __fdiv_rn(__fsub_rn((float)col + 0.5f, (float)cam.W * 0.5f), cam.focal);
    const float dy = -__fdiv_rn(__fsub_rn((float)row + 0.5f, (float)cam.H * 0.5f), cam.focal);
    const float dz = -1.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      rc.o[a] = cam.trans[a];
      rc.d[a] = fmaf(cam.rot[3 * a + 2], dz, fmaf(cam.rot[3 * a + 1], dy, cam.rot[3 * a + 0] * dx));
    }
    finish_ray(p, rc);
  }
  int a, b;
  sample_range(p, rc, a, b);
  JitterSource u_row;
  u_row.init(p, (int)t);
  float Y[LT::K];
  const float inv = 1.0f / rc.dnorm;
  sh_basis<DEG>(rc.d[0] * inv, rc.d[1] * inv, rc.d[2] * inv, (p.flags & kDiffuse) != 0, Y);
  float T = 1.f, V[LT::NV];
#pragma unroll
  for (int k = 0; k < LT::NV; ++k) V[k] = 0.f;
  DepthWalker zw;
  if (a < b) zw.init(p, rc, u_row, a);
#pragma unroll(kSampleUnroll)
  for (int i = a; i < b; ++i, zw.advance(p, rc, u_row, i - 1)) {
    const float zi = zw.cur;
    const float px = __fadd_rn(rc.o[0], __fmul_rn(rc.d[0], zi));
    const float py = __fadd_rn(rc.o[1], __fmul_rn(rc.d[1], zi));
    const float pz = __fadd_rn(rc.o[2], __fmul_rn(rc.d[2], zi));
    const bool in = inside_aabb(p, px, py, pz);
    Corners c;
    make_corners(p, px, py, pz, c);
    float raw[NCOL], dpost, col_k[NCOL];
    unsigned signs;
    const float sraw = gather_sample<DEG, NCOL>(p, c, Y, raw, signs);
    const float sigma = in ? post_act(p.postact, sraw, dpost) : 0.f;
#pragma unroll
    for (int k = 0; k < NCOL; ++k) col_k[k] = in ? sigmoid_fast(raw[k]) : 0.f;
    const float delta = ((i == p.S - 1) ? kInfinity : __fsub_rn(zw.next, zi)) * rc.dnorm;
    const float alpha = 1.0f - exp_fast(-(sigma * delta));
    const float w = alpha * T;
#pragma unroll
    for (int k = 0; k < NCOL; ++k) V[k] = fmaf(w, col_k[k], V[k]);
    V[NCOL] = fmaf(w, zi, V[NCOL]);
    V[NCOL + 1] += w;
    T *= (1.0f - alpha);
    if (T < cam.min_transmittance) break;
  }
  const float acc = V[NCOL + 1], depth = V[NCOL];
  const bool white = (p.flags & kWhite) && !(p.flags & kAttn);
#pragma unroll
  for (int k = 0; k < NCOL; ++k) p.colour[(size_t)t * NCOL + k] = white ? V[k] + (1.0f - acc) : V[k];
  p.depth[t] = depth;
  p.acc[t] = acc;
  if (p.disp != nullptr) {
    const float q = depth / acc;
    const float m = (q != q) ? q : fmaxf(kZeroPlus, q);
    p.disp[t] = 1.0f / m;
  }
}

template <int DEG, int NCOL>
cudaError_t launch_camera_t(const KParams& p, const CameraParams& cam, cudaStream_t stream) {
  const int threads = 128;
  const long long blocks = ((long long)p.R + threads - 1) / threads;
  render_camera_kernel<DEG, NCOL><<<(unsigned)blocks, threads, 0, stream>>>(p, cam);
  return cudaGetLastError();
}

template <int DEG, int NCOL, int REGCAP, class SP = SpecDynamic>
cudaError_t launch_pair(const KParams& p, bool backward, cudaStream_t stream) {
  using LT = Layout<DEG, NCOL>;
  const int threads = ((p.rpc * p.nseg + 31) / 32) * 32;
  if (threads > Bounds<REGCAP>::kThreads) return cudaErrorInvalidConfiguration;
  const int blocks = (p.R + p.rpc - 1) / p.rpc;
  size_t smem = sizeof(float) * ((size_t)(1 + LT::NV) * p.nseg * (p.rpc + 1) + (backward ? 2 * p.rpc : 0));
  if (backward && LT::CV > 1)
    smem = sizeof(float) * bwd_stage_offset_floats(LT::NV, p.nseg, p.rpc) + sizeof(float4) * 2 * (size_t)threads * BulkStage<LT::CV>::kSlot;
  auto k = backward ? render_bwd_kernel<DEG, NCOL, REGCAP, SP> : render_fwd_kernel<DEG, NCOL, REGCAP, SP>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  return launch_chained(k, dim3(blocks), dim3(threads), smem, stream, p);
}

template <int REGCAP>
cudaError_t dispatch_deg(const KParams& p, int deg, int ncol, bool backward, cudaStream_t stream) {
  if (ncol == 1) {
    if (deg == 0) return launch_pair<0, 1, REGCAP>(p, backward, stream);
    return cudaErrorInvalidValue;
  }
  switch (deg) {
    case 0: return launch_pair<0, 3, REGCAP>(p, backward, stream);
    case 1: return launch_pair<1, 3, REGCAP>(p, backward, stream);
    case 2: return launch_pair<2, 3, REGCAP>(p, backward, stream);
    case 3: return launch_pair<3, 3, REGCAP>(p, backward, stream);
  }
  return cudaErrorInvalidValue;
}

// Specialised variants (see Spec in voxe_device.cuh) exist for what the training loops actually run: RGB grids, the
// default register budgets of pick_shape (80 / 96 at SH-0, 128 above), ReLU or Softplus post-activation, with or without
// stratified jitter.  Everything else takes the generic kernels.
template <int DEG, int REGCAP>
cudaError_t dispatch_spec(const KParams& p, bool backward, cudaStream_t stream) {
  const bool perturb = (p.flags & kPerturb) != 0;
  if (p.postact == kPostRelu)
    return perturb ? launch_pair<DEG, 3, REGCAP, Spec<1, kPostRelu, true>>(p, backward, stream)
                   : launch_pair<DEG, 3, REGCAP, Spec<0, kPostRelu, true>>(p, backward, stream);
  return perturb ? launch_pair<DEG, 3, REGCAP, Spec<1, kPostSoftplus, true>>(p, backward, stream)
                 : launch_pair<DEG, 3, REGCAP, Spec<0, kPostSoftplus, true>>(p, backward, stream);
}

bool specialised_variant_exists(const KParams& p, int deg, int ncol, int regcap) {
  if (ncol != 3 || (p.postact != kPostRelu && p.postact != kPostSoftplus)) return false;
  const bool disparity = (p.flags & kDisparity) && !(p.flags & kAabb);
  if (disparity || p.noise_std != 0.f || p.preact != kPreIdentity || p.jitter != nullptr) return false;
  return deg == 0 ? (regcap == 80 || regcap == 96) : regcap == 128;
}

}  // namespace

cudaError_t launch_render(const KParams& p, int deg, int ncol, int regcap, bool backward, bool specialise, cudaStream_t stream,
                          bool* took_specialised) {
  *took_specialised = specialise && specialised_variant_exists(p, deg, ncol, regcap);
  if (*took_specialised) {
    switch (deg) {
      case 0: return regcap == 80 ? dispatch_spec<0, 80>(p, backward, stream) : dispatch_spec<0, 96>(p, backward, stream);
      case 1: return dispatch_spec<1, 128>(p, backward, stream);
      case 2: return dispatch_spec<2, 128>(p, backward, stream);
      case 3: return dispatch_spec<3, 128>(p, backward, stream);
    }
  }
  if (regcap <= 64) return dispatch_deg<64>(p, deg, ncol, backward, stream);
  if (regcap == 80) return dispatch_deg<80>(p, deg, ncol, backward, stream);
  if (regcap == 96) return dispatch_deg<96>(p, deg, ncol, backward, stream);
  return dispatch_deg<128>(p, deg, ncol, backward, stream);
}

cudaError_t launch_camera(const KParams& p, const CameraParams& cam, int deg, int ncol, cudaStream_t stream) {
  if (ncol == 1) return deg == 0 ? launch_camera_t<0, 1>(p, cam, stream) : cudaErrorInvalidValue;
  switch (deg) {
    case 0: return launch_camera_t<0, 3>(p, cam, stream);
    case 1: return launch_camera_t<1, 3>(p, cam, stream);
    case 2: return launch_camera_t<2, 3>(p, cam, stream);
    case 3: return launch_camera_t<3, 3>(p, cam, stream);
  }
  return cudaErrorInvalidValue;
}

int max_threads_per_cta(int regcap) {
  return regcap <= 64 ? Bounds<64>::kThreads : (regcap == 80 ? Bounds<80>::kThreads : (regcap == 96 ? Bounds<96>::kThreads : Bounds<128>::kThreads));
}

int saved_floats_per_segment(int ncol, int samples_per_segment) { return ncol + 3 + 4 * samples_per_segment; }

}  // namespace voxe

namespace voxe {
namespace {
__global__ void __launch_bounds__(256) jitter_fill_kernel(KParams p, float* out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)p.R * p.S) return;
  JitterSource u;
  u.init(p, (int)(t / p.S));
  out[t] = u.at(p, (int)(t % p.S));
}
}  // namespace

cudaError_t launch_jitter_fill(int R, int S, unsigned long long seed, unsigned long long offset, const long long* seed_dev,
                               const long long* offset_dev, unsigned long long intragraph, float* out, cudaStream_t stream) {
  KParams p{};
  p.R = R;
  p.S = S;
  p.rng_seed = seed;
  p.rng_offset = offset;
  if (seed_dev != nullptr && offset_dev != nullptr) {
    p.rng_seed_dev = seed_dev;
    p.rng_offset_dev = offset_dev;
    p.rng_intragraph = intragraph;
  }
  const long long n = (long long)R * S;
  jitter_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, out);
  return cudaGetLastError();
}
}  // namespace voxe
