// voxe_device.cuh -- device-side building blocks of the fused ray-marcher (sm_100a).
//
// Semantics follow SURVEY.md 8.A, i.e. the reference's
//   sample.py:15-68,71-184 (depths, slab test)      voxels.py:225-234,263-342 (normalise, inside test, trilinear)
//   process.py:46-91 (SH colour, mask)              accumulate.py:49-88 (compositing)
// Geometry (ray interval, depths, sample positions, inside test) is evaluated op for op in fp32 with explicit
// round-to-nearest intrinsics (no FMA contraction): with AABB-bound sampling the first/last sample lies exactly
// on a grid face and the reference's own fp32 rounding decides whether it is inside.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace voxe {

constexpr float kZeroPlus = 1e-10f;   // thre3d_atom/utils/constants.py:8
constexpr float kInfinity = 1e10f;    // thre3d_atom/utils/constants.py:9

enum : int { kPerturb = 1, kAabb = 2, kDisparity = 4, kWhite = 8, kDiffuse = 16, kAttn = 32 };
enum : int { kPreIdentity = 0, kPreAbs = 1 };
enum : int { kPostIdentity = 0, kPostRelu = 1, kPostSoftplus = 2 };

// What a kernel variant knows about the call at compile time.  The generic variant (SpecDynamic) reads every switch from
// KParams per sample; the specialised variants fix the stratified-jitter switch and the density post-activation and
// assume the plain case for the rest (linear sampling, no density noise, identity pre-activation, in-kernel jitter
// draws) -- the uniform branches, their convergence barriers and the dead alternatives leave the sample loop
// (forward: 760 -> ~470 SASS instructions per two samples).  Same arithmetic, same results.
template <int PERTURB /* -1: read p.flags */, int POST /* -1: read p.postact */, bool PLAIN>
struct Spec {
  static constexpr bool kPlain = PLAIN;
  __device__ static __forceinline__ bool perturb(int flags) {
    if constexpr (PERTURB >= 0) return PERTURB != 0; else return (flags & 1) != 0;  // kPerturb
  }
  __device__ static __forceinline__ int postact(int kind) {
    if constexpr (POST >= 0) return POST; else return kind;
  }
  __device__ static __forceinline__ bool disparity(bool d) {
    if constexpr (PLAIN) return false; else return d;
  }
  __device__ static __forceinline__ bool noise(float noise_std) {
    if constexpr (PLAIN) return false; else return noise_std != 0.f;
  }
  __device__ static __forceinline__ bool pre_abs(int preact) {
    if constexpr (PLAIN) return false; else return preact == 1;  // kPreAbs
  }
  __device__ static __forceinline__ bool jitter_buffer(const float* row) {
    if constexpr (PLAIN) return false; else return row != nullptr;
  }
};
using SpecDynamic = Spec<-1, -1, false>;

// Kernel parameter block (passed by value, lives in constant bank 0).
struct KParams {
  const float4* grid;   // packed [X*Y*Z][CV] float4
  float4* grad;         // packed gradient (backward only)
  const float* rays_o;
  const float* rays_d;
  const float* jitter;  // [R,S] or null
  const float* noise;   // [R,S] or null
  float* colour;        // fwd outputs
  float* depth;
  float* acc;
  float* disp;
  const float* g_colour;  // bwd inputs
  const float* g_depth;
  const float* g_acc;
  const float* g_disp;
  int R, S;
  int X, Y, Z;
  int sbx, sby;   // brick strides in voxel slots; with x' = x+1 etc. (one-voxel zero apron) voxel (x,y,z) lives at
                  // (x'>>1)*sbx + (y'>>1)*sby + (z'>>1)*8 + (x'&1)*4 + (y'&1)*2 + (z'&1)
  float lo[3], hi[3];
  float ua[3], ub[3];   // apron-shifted voxel coordinate u = p*ua + ub  (= ((p*nscale + nbias + 1) * N - 1) / 2 + 1, folded on the host)
  float near, far, dscale, noise_std, lin_step;
  int flags, preact, postact;
  int rpc, nseg, L;  // rays per CTA, sample segments per ray, samples per segment (threads = rpc * nseg -> x32)
  int group_round, group_rot;  // CTA -> ray-group mapping (ray_group in voxe_render.cu); group_rot == 0: identity
  unsigned long long rng_seed, rng_offset;  // in-kernel jitter (kPerturb with jitter == nullptr)
  const long long* rng_seed_dev;            // non-null: the generator state lives in device memory (CUDA-graph replays):
  const long long* rng_offset_dev;          //   seed = *rng_seed_dev, offset = *rng_offset_dev + rng_intragraph
  unsigned long long rng_intragraph;
  float* saved;      // workspace written by the forward / read by the backward (or null): [nseg*L][R] float4 sample
                     // vectors, then [NCOL+3][nseg][R] segment summaries (16-byte aligned)
  unsigned char* touched;  // backward, optional: one byte per 2x2x2 brick of the gradient volume; a sample that scatters
  int touch_tag;           // stores touch_tag at the brick of its corner 0 (its 8 corners lie in that brick and its +1
                           // neighbours), so that voxe_consume_grad visits only what this call wrote
  unsigned long long* stats;  // backward, optional device counters: [0] += in-grid samples, [1] += samples that scattered
};

// ---------------------------------------------------------------------------------------------------------
// depths
// ---------------------------------------------------------------------------------------------------------
// torch.linspace(0, 1, S): symmetric evaluation from both ends (exact 0 and 1 at the end points).
__device__ __forceinline__ float lin_t(int i, int S, float step) {
  return (i < (S >> 1)) ? __fmul_rn(step, (float)i) : __fsub_rn(1.0f, __fmul_rn(step, (float)(S - 1 - i)));
}

// Un-jittered depth of sample i (sample.py:46-54).  inv_near/inv_far are only read for disparity sampling.
__device__ __forceinline__ float depth_plain(const KParams& p, float near, float far, float inv_near, float inv_far,
                                             bool disparity, int i) {
  const float t = lin_t(i, p.S, p.lin_step);
  const float omt = __fsub_rn(1.0f, t);
  if (disparity) return __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(inv_near, omt), __fmul_rn(inv_far, t)));
  return __fadd_rn(__fmul_rn(near, omt), __fmul_rn(far, t));
}

// Per-ray (near, far): camera bounds, or the slab test of sample.py:71-184 (misses keep the camera bounds,
// hits ignore them; both ends clipped at 0).
__device__ __forceinline__ void ray_interval(const KParams& p, const float (&o)[3], const float (&d)[3], float& near,
                                             float& far) {
  near = p.near;
  far = p.far;
  if (!(p.flags & kAabb)) return;
  float lo = 0.f, hi = 0.f;
  bool hit = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float den = __fadd_rn(d[a], kZeroPlus);
    const float t0 = __fdiv_rn(__fsub_rn(p.lo[a], o[a]), den);
    const float t1 = __fdiv_rn(__fsub_rn(p.hi[a], o[a]), den);
    const float alo = (t0 > t1) ? t1 : t0;
    const float ahi = (t0 > t1) ? t0 : t1;
    if (a == 0) {
      lo = alo;
      hi = ahi;
    } else {
      if (lo > ahi || alo > hi) hit = false;
      lo = (alo > lo) ? alo : lo;
      hi = (ahi < hi) ? ahi : hi;
    }
  }
  if (hit) {
    near = lo;
    far = hi;
  }
  near = (near < 0.f) ? 0.f : near;  // torch.clip(min=0); NaN stays NaN
  far = (far < 0.f) ? 0.f : far;
}

// ---------------------------------------------------------------------------------------------------------
// SH basis with the reference's constants and signs (spherical_harmonics.py:33-50, 86-116)
// ---------------------------------------------------------------------------------------------------------
template <int DEG>
__device__ __forceinline__ void sh_basis(float x, float y, float z, bool diffuse, float (&Y)[(DEG + 1) * (DEG + 1)]) {
  Y[0] = 0.28209479177387814f;
  if constexpr (DEG > 0) {
    Y[1] = -0.4886025119029199f * y;
    Y[2] = 0.4886025119029199f * z;
    Y[3] = -0.4886025119029199f * x;
  }
  if constexpr (DEG > 1) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    Y[4] = 1.0925484305920792f * xy;
    Y[5] = -1.0925484305920792f * yz;
    Y[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
    Y[7] = -1.0925484305920792f * xz;
    Y[8] = 0.5462742152960396f * (xx - yy);
    if constexpr (DEG > 2) {
      Y[9] = -0.5900435899266435f * y * (3.f * xx - yy);
      Y[10] = 2.890611442640554f * xy * z;
      Y[11] = -0.4570457994644658f * y * (4.f * zz - xx - yy);
      Y[12] = 0.3731763325901154f * z * (2.f * zz - 3.f * xx - 3.f * yy);
      Y[13] = -0.4570457994644658f * x * (4.f * zz - xx - yy);
      Y[14] = 1.445305721320277f * z * (xx - yy);
      Y[15] = -0.5900435899266435f * x * (xx - 3.f * yy);
    }
  }
  if (diffuse) {  // render_diffuse: only the degree-0 coefficient (process.py:59-63)
#pragma unroll
    for (int k = 1; k < (DEG + 1) * (DEG + 1); ++k) Y[k] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------------
// trilinear corner set (grid_sample: bilinear, zeros padding, align_corners=False)
// ---------------------------------------------------------------------------------------------------------
struct Corners {
  int idx[8];   // voxel slot in the bricked, zero-aproned volume
  float w[8];
};

// Strict inside test on the world-space point (voxels.py:263-285).
__device__ __forceinline__ bool inside_aabb(const KParams& p, float px, float py, float pz) {
  return (px > p.lo[0]) & (px < p.hi[0]) & (py > p.lo[1]) & (py < p.hi[1]) & (pz > p.lo[2]) & (pz < p.hi[2]);
}

// One axis of the trilinear footprint.  The packed volume carries a one-voxel apron of zeros on every face, and the
// voxel coordinate is shifted by +1 accordingly: u = (p - lo)/voxel + 0.5, evaluated with one FMA (coefficients folded
// on the host in double; unlike the inside test this is a continuous function of p, so the different rounding
// relative to the reference's normalise -> unnormalise chain (voxels.py:225-234 + grid_sampler) only moves results by
// ~1e-7 * N).  For a point strictly inside the box floor(u) lies in [0, N], so both corners i and i+1 are always
// addressable: a corner beyond the grid reads an apron zero -- exactly grid_sample's zeros padding -- and no range
// checks or weight masking are needed.  (The clamp only guards the address against rounding at extreme coordinates.)
__device__ __forceinline__ void axis_setup(float pc, float ua, float ub, int N, int& i0, float& w0, float& w1) {
  const float u = fmaf(pc, ua, ub);
  const float fl = floorf(u);
  w1 = u - fl;
  w0 = 1.0f - w1;
  i0 = min(max((int)fl, 0), N);
}

__device__ __forceinline__ void make_corners(const KParams& p, float px, float py, float pz, Corners& c) {
  int x0, y0, z0;
  float wx0, wx1, wy0, wy1, wz0, wz1;
  axis_setup(px, p.ua[0], p.ub[0], p.X, x0, wx0, wx1);
  axis_setup(py, p.ua[1], p.ub[1], p.Y, y0, wy0, wy1);
  axis_setup(pz, p.ua[2], p.ub[2], p.Z, z0, wz0, wz1);
  // 2x2x2-brick addressing (one brick of SH-0 voxels = one 128-byte line): per-axis partial offsets, then 8 sums.
  // Index i+1 stays in the brick of i when i is even and moves to the next brick (low slot) when i is odd.
  const int xo = x0 & 1, yo = y0 & 1, zo = z0 & 1;
  const int xa0 = (x0 >> 1) * p.sbx + (xo << 2), xa1 = xa0 + (xo ? p.sbx - 4 : 4);
  const int ya0 = (y0 >> 1) * p.sby + (yo << 1), ya1 = ya0 + (yo ? p.sby - 2 : 2);
  const int za0 = ((z0 >> 1) << 3) + zo, za1 = za0 + (zo ? 7 : 1);
  const int r00 = xa0 + ya0, r01 = xa0 + ya1, r10 = xa1 + ya0, r11 = xa1 + ya1;
  const float w00 = wx0 * wy0, w01 = wx0 * wy1, w10 = wx1 * wy0, w11 = wx1 * wy1;
  c.idx[0] = r00 + za0; c.w[0] = w00 * wz0;
  c.idx[1] = r00 + za1; c.w[1] = w00 * wz1;
  c.idx[2] = r01 + za0; c.w[2] = w01 * wz0;
  c.idx[3] = r01 + za1; c.w[3] = w01 * wz1;
  c.idx[4] = r10 + za0; c.w[4] = w10 * wz0;
  c.idx[5] = r10 + za1; c.w[5] = w10 * wz1;
  c.idx[6] = r11 + za0; c.w[6] = w11 * wz0;
  c.idx[7] = r11 + za1; c.w[7] = w11 * wz1;
}

__device__ __forceinline__ float f4_get(const float4& v, int k) {
  return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}
__device__ __forceinline__ void f4_set(float4& v, int k, float x) {
  if (k == 0) v.x = x; else if (k == 1) v.y = x; else if (k == 2) v.z = x; else v.w = x;
}

// Fast transcendental forms (ex2.approx / rcp.approx based, ~2 ulp): three orders of magnitude below the 1e-4 pixel
// tolerance, and they keep the per-sample instruction count (the kernels are issue/latency bound, not DRAM bound).
// ex2.approx.ftz directly: __expf() wraps the same instruction in a range fix-up for denormal results (compare, two
// conditional multiplies: 3 extra instructions per call, 4 calls per sample); flushing exp(x) to 0 below x = -87.3
// changes a sigmoid or an alpha by less than 1e-38.
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float exp_fast(float x) { return ex2_ftz(x * 1.4426950408889634f); }
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + ex2_ftz(x * -1.4426950408889634f)); }

// post-activation and its derivative w.r.t. the interpolated (pre-activated) density
__device__ __forceinline__ float post_act(int kind, float x, float& dydx) {
  if (kind == kPostRelu) {
    dydx = (x > 0.f) ? 1.f : 0.f;
    return fmaxf(x, 0.f);
  }
  if (kind == kPostSoftplus) {  // torch.nn.Softplus(beta=1, threshold=20)
    if (x > 20.f) {
      dydx = 1.f;
      return x;
    }
    const float e = __expf(x);
    dydx = __fdividef(e, 1.f + e);
    return log1pf(e);
  }
  dydx = 1.f;
  return x;
}

// 16-byte vector reduction into global memory (REDG.E.ADD.F32x4 on sm_100a).
__device__ __forceinline__ void red_add_v4(float4* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// have its CTAs made resident while the previous kernel of the stream is still draining; pdl_wait() blocks until that
// kernel has completed and its writes are visible (a no-op without the attribute), pdl_launch_dependents() tells the
// scheduler that the NEXT kernel's CTAs may be placed as soon as every CTA of this grid has got here or exited.  The
// render kernels of a training loop are a chain of short launches (fwd -> bwd -> hand-over -> fwd ...), each with ~2-3 us
// of launch ramp on an otherwise idle GPU; this overlaps the ramp of one with the tail of the previous.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Everything a thread needs to know about its ray.
struct RayCtx {
  float o[3], d[3];
  float near, far, inv_near, inv_far, dnorm;
  bool disparity;
};

// Everything derived from a ray's origin and direction (rc.o, rc.d already set).
__device__ __forceinline__ void finish_ray(const KParams& p, RayCtx& rc) {
  ray_interval(p, rc.o, rc.d, rc.near, rc.far);
  rc.disparity = (p.flags & kDisparity) && !(p.flags & kAabb);  // renderers.py:66-78
  rc.inv_near = rc.inv_far = 0.f;
  if (rc.disparity) {
    rc.inv_near = __fdiv_rn(1.0f, __fadd_rn(rc.near, kZeroPlus));
    rc.inv_far = __fdiv_rn(1.0f, rc.far);
  }
  rc.dnorm = sqrtf(rc.d[0] * rc.d[0] + rc.d[1] * rc.d[1] + rc.d[2] * rc.d[2]);
}

__device__ __forceinline__ void load_ray(const KParams& p, int ray, RayCtx& rc) {
  const float* po = p.rays_o + 3 * (size_t)ray;
  const float* pd = p.rays_d + 3 * (size_t)ray;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    rc.o[a] = __ldg(po + a);
    rc.d[a] = __ldg(pd + a);
  }
  finish_ray(p, rc);
}

// Pinhole camera of the whole-camera inference kernel (cast_rays, rendering/volumetric/utils/misc.py:12-50).
struct CameraParams {
  int H, W;
  float focal;
  float rot[9];    // camera-to-world rotation, row-major
  float trans[3];  // camera position
  long long first_pixel;  // flat pixel index (row * W + col) of ray 0 of this launch
  float min_transmittance;  // early termination: stop a ray once T < this (0 = never)
};

// Conservative index range [a, b) of the samples of one ray that can lie inside the grid AABB.  Samples outside the
// box contribute exactly nothing (sigma = 0 -> alpha = 0, process.py:80-91), so the kernels only distribute [a, b)
// over a ray's threads; the exact per-sample inside test (inside_aabb) still decides, this range only has to be a
// superset.  Plain depths are z_i = near + (far-near) * i/(S-1); stratified jitter keeps z'_i between the mid-points to
// its neighbours, i.e. within half a step of z_i.  The range is widened by one further sample on each side against
// fp32 rounding of the slab arithmetic (a sample step is ~1e4 ulp of z).  Disparity sampling (non-linear in i), density
// noise (every sample contributes) and degenerate intervals use the full range.
__device__ __forceinline__ void sample_range(const KParams& p, const RayCtx& rc, int& a, int& b) {
  a = 0;
  b = p.S;
  if (rc.disparity || p.noise_std != 0.f) return;
  const float span = rc.far - rc.near;
  if (!(span > 0.f)) return;
  float zin = -kInfinity, zout = kInfinity;
  bool empty = false;
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    const float o = rc.o[ax], d = rc.d[ax];
    if (d == 0.f) {
      empty |= !(o > p.lo[ax] && o < p.hi[ax]);
    } else {
      const float inv = 1.0f / d;
      const float t0 = (p.lo[ax] - o) * inv, t1 = (p.hi[ax] - o) * inv;
      zin = fmaxf(zin, fminf(t0, t1));
      zout = fminf(zout, fmaxf(t0, t1));
    }
  }
  if (empty || !(zin <= zout)) {  // also catches NaN
    b = 0;
    return;
  }
  const float per_z = (float)(p.S - 1) / span;  // samples per unit depth
  const float fs = (float)p.S;
  const float xa = fminf(fmaxf((zin - rc.near) * per_z, -2.f), fs + 2.f);
  const float xb = fminf(fmaxf((zout - rc.near) * per_z, -2.f), fs + 2.f);
  a = max((int)floorf(xa) - 1, 0);
  b = min((int)ceilf(xb) + 2, p.S);
  if (b < a) b = a;
}

// PCG output hash (O'Neill's PCG-RXS-M-XS-32 permutation; the best quality-per-instruction 32-bit hash in Jarzynski &
// Olano, "Hash Functions for GPU Rendering", JCGT 2020): ~7 integer instructions, no state.
__device__ __forceinline__ unsigned pcg_hash(unsigned v) {
  const unsigned state = v * 747796405u + 2891336453u;
  const unsigned word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
  return (word >> 22u) ^ word;
}

// Stratified-jitter draws u[ray][i] in [0, 1): either the caller's [R,S] buffer (torch.rand, sample.py:63) or generated
// here as a counter-based hash of (seed, offset, ray, sample) -- 24 mantissa bits per draw like torch's uniform.  The
// kernels are bound by per-thread instruction latency, so the generator is a short hash rather than a Philox block
// (measured: Philox4x32-7 added 22 % to the instruction count of the forward kernel).  Forward and backward regenerate
// identical values; a draw depends on (seed, offset, ray, sample) only.
struct JitterSource {
  const float* row;
  unsigned base;

  __device__ __forceinline__ void init(const KParams& p, int ray_index) {
    row = p.jitter ? p.jitter + (size_t)ray_index * p.S : nullptr;
    unsigned long long seed = p.rng_seed, offset = p.rng_offset;
    if (p.rng_seed_dev != nullptr) {
      seed = (unsigned long long)__ldg(p.rng_seed_dev);
      offset = (unsigned long long)__ldg(p.rng_offset_dev) + p.rng_intragraph;
    }
    unsigned k = pcg_hash((unsigned)seed ^ pcg_hash((unsigned)(seed >> 32)));
    k = pcg_hash(k ^ (unsigned)offset);
    k = pcg_hash(k ^ (unsigned)(offset >> 32));
    base = pcg_hash(k ^ (unsigned)ray_index) + (unsigned)ray_index * 0x9E3779B9u;  // per-ray stream start
  }
  template <class SP = SpecDynamic>
  __device__ __forceinline__ float at(const KParams& p, int i) const {
    if (SP::jitter_buffer(row)) return __ldg(row + i);
    // (ray, sample) hashed jointly: the sample index enters through a multiply and an XOR, so two rays whose stream
    // starts happen to lie within S of each other do not replay each other's draws with a shift
    return (float)(pcg_hash(base ^ ((unsigned)i * 0x85EBCA6Bu)) >> 8) * (1.0f / 16777216.0f);
  }
};

// Rolling evaluation of the (optionally jittered) sample depths along one ray: `cur` is the depth of sample i,
// `next` the depth of sample i+1 (it closes the interval delta_i).  Stratified jitter follows sample.py:57-64:
// z'_i = lower_i + (upper_i - lower_i) * u_i with lower/upper the mid-points to the neighbouring plain depths.
struct DepthWalker {
  float a, b, c, d;  // plain depths of samples i-1, i, i+1, i+2 (indices clamped into [0, S-1])
  float cur, next;
  float u_pre;       // jitter of sample i+2, fetched one iteration before it is needed

  template <class SP = SpecDynamic>
  __device__ __forceinline__ float plain(const KParams& p, const RayCtx& rc, int k) const {
    return depth_plain(p, rc.near, rc.far, rc.inv_near, rc.inv_far, SP::disparity(rc.disparity), min(max(k, 0), p.S - 1));
  }
  __device__ __forceinline__ static float jittered(const KParams& p, float lo, float mid, float hi, int i, float u) {
    const float lower = (i <= 0) ? mid : __fmul_rn(0.5f, __fadd_rn(mid, lo));
    const float upper = (i >= p.S - 1) ? mid : __fmul_rn(0.5f, __fadd_rn(hi, mid));
    return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), u));
  }
  template <class SP = SpecDynamic>
  __device__ __forceinline__ void init(const KParams& p, const RayCtx& rc, JitterSource& u, int i) {
    b = plain<SP>(p, rc, i);
    c = plain<SP>(p, rc, i + 1);
    if (SP::perturb(p.flags)) {
      const float u0 = u.at<SP>(p, min(i, p.S - 1));
      const float u1 = u.at<SP>(p, min(i + 1, p.S - 1));
      u_pre = u.at<SP>(p, min(i + 2, p.S - 1));
      a = plain<SP>(p, rc, i - 1);
      d = plain<SP>(p, rc, i + 2);
      cur = jittered(p, a, b, c, i, u0);
      next = jittered(p, b, c, d, i + 1, u1);
    } else {
      cur = b;
      next = c;
    }
  }
  // step from sample i to sample i+1
  template <class SP = SpecDynamic>
  __device__ __forceinline__ void advance(const KParams& p, const RayCtx& rc, JitterSource& u, int i) {
    cur = next;
    if (SP::perturb(p.flags)) {
      a = b;
      b = c;
      c = d;
      d = plain<SP>(p, rc, i + 3);
      next = jittered(p, b, c, d, i + 2, u_pre);
      u_pre = u.at<SP>(p, min(i + 3, p.S - 1));
    } else {
      next = plain<SP>(p, rc, i + 2);
    }
  }
};

}  // namespace voxe
