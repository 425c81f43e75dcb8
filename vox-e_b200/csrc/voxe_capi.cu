// voxe_capi.cu -- the C ABI of libvoxe_sm100a.so (see include/voxe.h for the contract and the reference
// interfaces each entry point replaces).  Plain pointers and sizes in, error codes out; no torch types.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "voxe.h"
#include "voxe_device.cuh"
#include "voxe_launch.h"

namespace {

thread_local char g_error[512] = "";
std::atomic<int64_t> g_launches{0}, g_specialised{0};
std::atomic<int> g_tune_l{0}, g_tune_rpc{0}, g_tune_reg{0};

// The flag-specialised kernel variants (same arithmetic, per-call switches resolved at compile time: -14 % forward time on
// the benchmark batch, parity suites green under them, profiles/r2a_specialised_ab.txt) are taken wherever they exist;
// VOXE_SPECIALISED_KERNELS=0 forces the generic kernels (A/B runs, tests/test_specialised_kernels.py).
bool specialised_kernels() {
  static const bool on = [] {
    const char* v = std::getenv("VOXE_SPECIALISED_KERNELS");
    return !(v != nullptr && v[0] == '0');
  }();
  return on;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(VOXE_ERR_CUDA_BASE + (int)e, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

int check_grid(const VoxeGridDesc* g) {
  if (!g) return fail(VOXE_ERR_INVALID_ARGUMENT, "grid descriptor is NULL");
  for (int a = 0; a < 3; ++a)
    if (g->dims[a] < 1) return fail(VOXE_ERR_INVALID_ARGUMENT, "grid dims must be >= 1 (got %d on axis %d)", g->dims[a], a);
  if (g->n_features < 1) return fail(VOXE_ERR_INVALID_ARGUMENT, "n_features must be >= 1");
  if (g->channels != voxe_packed_channels(g->n_features))
    return fail(VOXE_ERR_INVALID_ARGUMENT, "channels must be roundup4(n_features+1) = %d (got %d)",
                voxe_packed_channels(g->n_features), g->channels);
  const int64_t nvec = voxe::packed_voxel_slots(g->dims) * (g->channels / 4);
  if (nvec >= (int64_t)1 << 31) return fail(VOXE_ERR_UNSUPPORTED, "packed grid has %lld 16-byte vectors; limit is 2^31", (long long)nvec);
  if (g->preact != VOXE_PREACT_IDENTITY && g->preact != VOXE_PREACT_ABS)
    return fail(VOXE_ERR_UNSUPPORTED, "density pre-activation %d is not in the fused set {identity, abs}", g->preact);
  if (g->postact < VOXE_POSTACT_IDENTITY || g->postact > VOXE_POSTACT_SOFTPLUS)
    return fail(VOXE_ERR_UNSUPPORTED, "density post-activation %d is not in the fused set {identity, relu, softplus}", g->postact);
  return VOXE_OK;
}

int check_render(const VoxeGridDesc* g, const VoxeRenderDesc* r, const float* jitter, const float* noise) {
  if (!r) return fail(VOXE_ERR_INVALID_ARGUMENT, "render descriptor is NULL");
  if (r->num_samples < 2) return fail(VOXE_ERR_INVALID_ARGUMENT, "num_samples must be >= 2 (got %d)", r->num_samples);
  if (r->sh_degree < 0 || r->sh_degree > 3)
    return fail(VOXE_ERR_UNSUPPORTED, "only SH degrees 0..3 are supported (got %d)", r->sh_degree);
  const bool attn = (r->flags & VOXE_FLAG_ATTN) != 0;
  if (r->n_colour != (attn ? 1 : 3)) return fail(VOXE_ERR_INVALID_ARGUMENT, "n_colour must be %d", attn ? 1 : 3);
  if (attn && r->sh_degree != 0) return fail(VOXE_ERR_UNSUPPORTED, "the attention render uses a single SH-0 channel");
  const int k = (r->sh_degree + 1) * (r->sh_degree + 1);
  if (g->n_features != r->n_colour * k)
    return fail(VOXE_ERR_INVALID_ARGUMENT, "n_features %d does not match n_colour*(deg+1)^2 = %d", g->n_features, r->n_colour * k);
  (void)jitter;  // NULL with VOXE_FLAG_PERTURB: in-kernel draws from (rng_seed, rng_offset)
  if (r->noise_std != 0.f && !noise) return fail(VOXE_ERR_INVALID_ARGUMENT, "noise_std != 0 needs the noise buffer [R,S]");
  return VOXE_OK;
}

// Launch shape: depth segments per ray nseg = ceil(S / L) (each thread streams 1/nseg of the ray's in-grid samples, see
// thread_samples in voxe_render.cu), rays per CTA, register budget of the kernel variant.
// Measured on B200 at S=256 (profiles/r1_sweep*.txt): small CTAs (four warps: 8 rays x 16 segments) balance best over
// the 148 SMs; the 128-register variant has no spills and, with fewer and longer threads, leaves room for a second
// batch's kernels to run beside it (a frame keeps 2-3 batches in flight): 30 us per 4096-ray batch fwd+bwd against
// 38 us for 4 rays x 32 segments at 64 registers, although the kernels timed alone are equal.
int sm_count() {
  static const int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1) v = 148;
    return v;
  }();
  return n;
}

// `backward` only affects rays per CTA (the workspace layout depends on L and nseg alone, so the two kernels of a render
// may group rays differently).
int pick_shape(int S, int sh_degree, int64_t R, bool backward, int& L, int& nseg, int& rpc, int& regcap) {
  // Register budget (does not affect the workspace layout).  With several batches in flight, or one large launch, the
  // number of resident CTAs is limited by registers, and the SH-0 kernels fit 96 / 80 registers without spilling:
  // 96 costs nothing on a lone 4096-ray launch and gains 5 % with three batches in flight; 80 is best for launches that
  // fill the machine several times over (one 160 000-ray launch: 955 us against 1058 us at 128).  Higher SH degrees keep
  // the 128-register variant (their corner gathers hold CV vectors per corner).
  regcap = g_tune_reg.load();
  if (regcap != 64 && regcap != 80 && regcap != 96 && regcap != 128) regcap = (sh_degree == 0) ? (R >= 16384 ? 80 : 96) : 128;
  const int max_threads = voxe::max_threads_per_cta(regcap);
  // Multi-vector voxels (SH degree >= 1: 64 / 112 / 208 bytes per corner): the gather wants WIDE ray bundles -- 16
  // neighbouring rays at the same depth read neighbouring voxels, so a warp-wide load touches a few long runs instead of
  // eight scattered ones -- and therefore few, long depth segments (~8-11 per ray).  Measured on the 15 GB grid of
  // BASELINE.json's configuration 5 (512^3 SH-2, S = 512, 65 536 rays; profiles/r2_shape_sweeps.txt): forward 7.57 ms with
  // 4 rays x 32 segments per CTA, 2.44 ms with 16 rays x 11 segments (8.4 TB/s of corner bytes: the L2 now serves the
  // overlap between neighbouring rays); the backward (one TMA reduction per corner) is indifferent and keeps 8 rays.
  const bool wide = sh_degree >= 1;
  L = g_tune_l.load();
  if (L < 1 || L > 256) {
    if (wide && S > 128) L = (((S + 10) / 11) + 15) / 16 * 16;  // ~11 segments, a multiple of 16 samples each
    else L = (S <= 32) ? 4 : (S < 128 ? 8 : 16);
  }
  while ((S + L - 1) / L > max_threads) L *= 2;  // very long rays: more samples per thread
  nseg = (S + L - 1) / L;
  rpc = g_tune_rpc.load();
  if (rpc <= 0 || rpc > 32) {
    if (wide && !backward && (R + 15) / 16 >= 2 * (int64_t)sm_count()) {
      rpc = 16;
      while (rpc > 8 && rpc * nseg > 256) rpc >>= 1;
    } else {
      rpc = 32;
      while (rpc > 1 && rpc * nseg > 128) rpc >>= 1;
    }
  }
  while (rpc > 1 && rpc * nseg > max_threads) --rpc;
  return VOXE_OK;
}

void fill_grid(const VoxeGridDesc* g, voxe::KParams& p) {
  p.X = g->dims[0];
  p.Y = g->dims[1];
  p.Z = g->dims[2];
  p.sby = ((g->dims[2] + 3) / 2) * 8;  // bricks over the aproned extent (N + 2 voxels per axis)
  p.sbx = ((g->dims[1] + 3) / 2) * p.sby;
  for (int a = 0; a < 3; ++a) {
    p.lo[a] = g->aabb_lo[a];
    p.hi[a] = g->aabb_hi[a];
    // u = ((p*scale + bias + 1) * N - 1) / 2 folded into one multiply-add
    p.ua[a] = (float)(0.5 * (double)g->norm_scale[a] * g->dims[a]);
    p.ub[a] = (float)(0.5 * (((double)g->norm_bias[a] + 1.0) * g->dims[a] - 1.0) + 1.0);  // +1: zero apron
  }
  p.dscale = g->density_scale;
  p.preact = g->preact;
  p.postact = g->postact;
}

int fill_params(const VoxeGridDesc* g, const VoxeRenderDesc* r, int64_t R, bool backward, voxe::KParams& p, int& regcap) {
  std::memset(&p, 0, sizeof(p));
  fill_grid(g, p);
  p.R = (int)R;
  p.S = r->num_samples;
  p.near = r->near;
  p.far = r->far;
  p.noise_std = r->noise_std;
  p.rng_seed = r->rng_seed;
  p.rng_offset = r->rng_offset;
  if (r->rng_seed_dev != nullptr && r->rng_offset_dev != nullptr) {
    p.rng_seed_dev = reinterpret_cast<const long long*>(r->rng_seed_dev);
    p.rng_offset_dev = reinterpret_cast<const long long*>(r->rng_offset_dev);
    p.rng_intragraph = r->rng_offset_intragraph;
  }
  p.lin_step = 1.0f / (float)(r->num_samples - 1);
  p.flags = r->flags;
  if (int rc = pick_shape(p.S, r->sh_degree, R, backward, p.L, p.nseg, p.rpc, regcap)) return rc;
  // CTA -> ray-group rotation (see ray_group in voxe_render.cu), off by default: on the benchmark frames the work per
  // group varies by +-8 % only and rotating the rounds changed nothing (profiles/r2_shape_sweeps.txt); kept as a tuning
  // knob (VOXE_GROUP_ROTATION = groups of rotation per round of one CTA per SM) for batches with emptier image borders.
  static const int rot_override = [] {
    const char* v = std::getenv("VOXE_GROUP_ROTATION");
    return v ? std::atoi(v) : 0;
  }();
  const int groups = (int)((R + p.rpc - 1) / p.rpc);
  p.group_round = sm_count();
  p.group_rot = (rot_override > 0 && groups >= 2 * sm_count()) ? rot_override : 0;
  return VOXE_OK;
}

}  // namespace

extern "C" {

int voxe_abi_version(void) { return VOXE_ABI_VERSION; }

const char* voxe_last_error(void) { return g_error; }

int voxe_packed_channels(int n_features) { return ((n_features + 1 + 3) / 4) * 4; }

int64_t voxe_packed_floats(const VoxeGridDesc* grid) {
  if (!grid || grid->dims[0] < 1 || grid->dims[1] < 1 || grid->dims[2] < 1 || grid->channels < 4) return 0;
  return voxe::packed_voxel_slots(grid->dims) * grid->channels;
}

int64_t voxe_launch_count(void) { return g_launches.load(); }

int64_t voxe_specialised_launch_count(void) { return g_specialised.load(); }

int voxe_set_tuning(int samples_per_thread, int rays_per_cta, int register_cap) {
  if (samples_per_thread < 0 || samples_per_thread > 256)
    return fail(VOXE_ERR_INVALID_ARGUMENT, "samples_per_thread must be in 0..256");
  if (rays_per_cta < 0 || rays_per_cta > 32) return fail(VOXE_ERR_INVALID_ARGUMENT, "rays_per_cta must be in 0..32");
  if (register_cap != 0 && register_cap != 64 && register_cap != 80 && register_cap != 96 && register_cap != 128)
    return fail(VOXE_ERR_INVALID_ARGUMENT, "register_cap must be 0, 64, 80, 96 or 128");
  g_tune_l.store(samples_per_thread);
  g_tune_rpc.store(rays_per_cta);
  g_tune_reg.store(register_cap);
  return VOXE_OK;
}

int64_t voxe_saved_floats(const VoxeRenderDesc* render, int64_t num_rays) {
  if (!render || render->num_samples < 2 || num_rays < 0) return 0;
  int L, nseg, rpc, regcap;
  if (pick_shape(render->num_samples, render->sh_degree, num_rays, false, L, nseg, rpc, regcap)) return 0;
  return (int64_t)voxe::saved_floats_per_segment(render->n_colour, L) * nseg * num_rays;
}

int voxe_pack_grid(const VoxeGridDesc* grid, const float* densities, const float* features, float* packed,
                   voxe_stream_t stream) {
  if (int rc = check_grid(grid)) return rc;
  if (!densities || !features || !packed) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_pack_grid: NULL buffer");
  cudaError_t e = voxe::launch_pack_grid(densities, features, packed, grid->dims, grid->n_features, grid->channels,
                                         (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "voxe_pack_grid launch");
  g_launches.fetch_add(1);
  return VOXE_OK;
}

int voxe_unpack_grad(const VoxeGridDesc* grid, const float* packed_grad, float* d_densities, float* d_features,
                     int accumulate, voxe_stream_t stream) {
  if (int rc = check_grid(grid)) return rc;
  if (!packed_grad) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_unpack_grad: NULL packed_grad");
  if (!d_densities && !d_features) return VOXE_OK;
  cudaError_t e = voxe::launch_unpack_grad(packed_grad, d_densities, d_features, grid->dims, grid->n_features,
                                           grid->channels, accumulate != 0, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "voxe_unpack_grad launch");
  g_launches.fetch_add(1);
  return VOXE_OK;
}

int voxe_jitter_fill(const VoxeRenderDesc* render, float* out, int64_t num_rays, voxe_stream_t stream) {
  if (!render || render->num_samples < 1 || !out || num_rays < 0 || num_rays > 0x7fffffff)
    return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_jitter_fill: bad arguments");
  if (num_rays == 0) return VOXE_OK;
  cudaError_t e = voxe::launch_jitter_fill((int)num_rays, render->num_samples, render->rng_seed, render->rng_offset,
                                           reinterpret_cast<const long long*>(render->rng_seed_dev),
                                           reinterpret_cast<const long long*>(render->rng_offset_dev), render->rng_offset_intragraph,
                                           out, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "voxe_jitter_fill launch");
  g_launches.fetch_add(1);
  return VOXE_OK;
}

int64_t voxe_touched_bytes(const VoxeGridDesc* grid) {
  if (!grid || grid->dims[0] < 1 || grid->dims[1] < 1 || grid->dims[2] < 1) return 0;
  return voxe::packed_bricks(grid->dims);
}

int voxe_consume_grad(const VoxeGridDesc* grid, float* packed_grad, float* d_densities, float* d_features,
                      const uint8_t* touched, int32_t touch_tag, voxe_stream_t stream) {
  if (int rc = check_grid(grid)) return rc;
  if (!packed_grad) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_consume_grad: NULL packed_grad");
  if (touched && (touch_tag < 1 || touch_tag > 255)) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_consume_grad: touch_tag must be in 1..255");
  cudaError_t e = voxe::launch_consume_grad(packed_grad, d_densities, d_features, grid->dims, grid->n_features,
                                            grid->channels, touched, touch_tag, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "voxe_consume_grad launch");
  g_launches.fetch_add(1);
  return VOXE_OK;
}

int voxe_adam_step(const VoxeGridDesc* grid, const VoxeAdamDesc* adam, float* densities, float* features, float* packed,
                   float* packed_grad, const float* dense_d_densities, const float* dense_d_features, float* packed_m,
                   float* packed_v, voxe_stream_t stream) {
  if (int rc = check_grid(grid)) return rc;
  if (!adam) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_adam_step: NULL descriptor");
  if ((!densities && !features) || !packed || !packed_m || !packed_v)
    return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_adam_step: NULL parameters / packed volume / moment buffer");
  if (adam->step < 1) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_adam_step: step must be >= 1 (got %d)", adam->step);
  if (!(adam->beta1 >= 0.0 && adam->beta1 < 1.0 && adam->beta2 >= 0.0 && adam->beta2 < 1.0 && adam->eps >= 0.0))
    return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_adam_step: betas must be in [0,1) and eps >= 0");
  cudaError_t e = voxe::launch_adam_step(packed, packed_grad, packed_m, packed_v, densities, features, dense_d_densities,
                                         dense_d_features, grid->dims, grid->n_features, grid->channels, adam->lr,
                                         adam->beta1, adam->beta2, adam->eps, adam->step, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "voxe_adam_step launch");
  g_launches.fetch_add(1);
  return VOXE_OK;
}

int voxe_resample_grid(const float* grid_in, const int32_t in_dims[3], int32_t channels, float* grid_out, const int32_t out_dims[3],
                       voxe_stream_t stream) {
  if (!grid_in || !grid_out || !in_dims || !out_dims) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_resample_grid: NULL buffer / dims");
  for (int a = 0; a < 3; ++a)
    if (in_dims[a] < 1 || out_dims[a] < 1) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_resample_grid: dims must be >= 1");
  if (channels < 1) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_resample_grid: channels must be >= 1");
  if ((int64_t)out_dims[0] * out_dims[1] * out_dims[2] * channels >= ((int64_t)1 << 31) * 256)
    return fail(VOXE_ERR_UNSUPPORTED, "voxe_resample_grid: output too large for one launch");
  const int di[3] = {in_dims[0], in_dims[1], in_dims[2]}, dout[3] = {out_dims[0], out_dims[1], out_dims[2]};
  cudaError_t e = voxe::launch_resample_grid(grid_in, di, channels, grid_out, dout, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "voxe_resample_grid launch");
  g_launches.fetch_add(1);
  return VOXE_OK;
}

static int query_points(const VoxeGridDesc* grid, const float* packed, const float* points, float* out, const float* g_out,
                        float* packed_grad, int64_t n, voxe_stream_t stream, const char* what) {
  if (int rc = check_grid(grid)) return rc;
  if (n < 0) return fail(VOXE_ERR_INVALID_ARGUMENT, "%s: num_points must be >= 0", what);
  if (n == 0) return VOXE_OK;
  if (!packed || !points || (!out && !g_out) || (g_out && !packed_grad)) return fail(VOXE_ERR_INVALID_ARGUMENT, "%s: NULL buffer", what);
  voxe::KParams p;
  std::memset(&p, 0, sizeof(p));
  fill_grid(grid, p);
  p.grid = reinterpret_cast<const float4*>(packed);
  p.grad = reinterpret_cast<float4*>(packed_grad);
  cudaError_t e = voxe::launch_query_points(p, points, out, g_out, (long long)n, grid->channels / 4, grid->n_features, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, what);
  g_launches.fetch_add(1);
  return VOXE_OK;
}

int voxe_query_points(const VoxeGridDesc* grid, const float* packed, const float* points, float* out, int64_t num_points,
                      voxe_stream_t stream) {
  return query_points(grid, packed, points, out, nullptr, nullptr, num_points, stream, "voxe_query_points");
}

int voxe_query_points_bwd(const VoxeGridDesc* grid, const float* packed, const float* points, const float* g_out,
                          float* packed_grad, int64_t num_points, voxe_stream_t stream) {
  if (!g_out) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_query_points_bwd: NULL g_out");
  return query_points(grid, packed, points, nullptr, g_out, packed_grad, num_points, stream, "voxe_query_points_bwd");
}

int voxe_tv_regularizer(const float* grid, const int32_t dims[3], int32_t channels, int32_t relu, double* workspace,
                        float* loss, const float* upstream, float scale, float* grad, int32_t accumulate, voxe_stream_t stream) {
  if (!grid || !dims) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_tv_regularizer: NULL grid / dims");
  if (dims[0] < 1 || dims[1] < 1 || dims[2] < 1 || channels < 1)
    return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_tv_regularizer: dims and channels must be >= 1");
  if ((int64_t)dims[2] * channels > 0x7fffffff || (int64_t)dims[0] * dims[1] > 0x7fffffff)
    return fail(VOXE_ERR_UNSUPPORTED, "voxe_tv_regularizer: grid rows out of range");
  if (!loss && !grad) return VOXE_OK;
  if (loss && !workspace) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_tv_regularizer: the loss needs the workspace");
  const int d[3] = {dims[0], dims[1], dims[2]};
  int launches = 0;
  cudaError_t e = voxe::launch_tv(grid, d, channels, relu != 0, workspace, loss, upstream, scale, grad, accumulate != 0,
                                  (cudaStream_t)stream, &launches);
  g_launches.fetch_add(launches);
  if (e != cudaSuccess) return cuda_fail(e, "voxe_tv_regularizer launch");
  return VOXE_OK;
}

int voxe_pair_loss(const float* a, const float* b, int64_t n, int32_t mode, double* workspace, float* loss,
                   float* correlation_grid, voxe_stream_t stream) {
  if (!a || !b || !workspace || !loss) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_pair_loss: NULL buffer");
  if (n < 1) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_pair_loss: n must be >= 1");
  if (mode < VOXE_PAIR_CORRELATION || mode > VOXE_PAIR_L1) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_pair_loss: unknown mode %d", mode);
  int launches = 0;
  cudaError_t e = voxe::launch_pair_loss(a, b, n, mode, 1e-7f, workspace, loss, correlation_grid, (cudaStream_t)stream, &launches);
  g_launches.fetch_add(launches);
  if (e != cudaSuccess) return cuda_fail(e, "voxe_pair_loss launch");
  return VOXE_OK;
}

int voxe_pair_loss_grad(const float* a, const float* b, int64_t n, int32_t mode, const double* workspace, const float* upstream,
                        float scale, float* grad, int32_t accumulate, voxe_stream_t stream) {
  if (!a || !b || !grad) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_pair_loss_grad: NULL buffer");
  if (n < 1) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_pair_loss_grad: n must be >= 1");
  if (mode < VOXE_PAIR_CORRELATION || mode > VOXE_PAIR_L1)
    return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_pair_loss_grad: unknown mode %d", mode);
  if (mode == VOXE_PAIR_CORRELATION && !workspace)
    return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_pair_loss_grad: correlation mode reads the workspace voxe_pair_loss filled");
  cudaError_t e = voxe::launch_pair_grad(a, b, n, mode, 1e-7f, workspace, upstream, scale, grad, accumulate != 0, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "voxe_pair_loss_grad launch");
  g_launches.fetch_add(1);
  return VOXE_OK;
}

static int check_peers(const VoxePeerDesc* peers, int64_t n_floats, const char* what) {
  if (!peers) return fail(VOXE_ERR_INVALID_ARGUMENT, "%s: NULL descriptor", what);
  if (peers->world_size < 1 || peers->world_size > VOXE_MAX_PEERS || peers->rank < 0 || peers->rank >= peers->world_size)
    return fail(VOXE_ERR_INVALID_ARGUMENT, "%s: world_size must be in 1..%d and rank inside it", what, VOXE_MAX_PEERS);
  if (n_floats < 0 || (n_floats & 3)) return fail(VOXE_ERR_INVALID_ARGUMENT, "%s: n_floats must be a multiple of 4", what);
  for (int k = 0; k < peers->world_size; ++k) {
    if (!peers->buffers[k] || !peers->signals[k]) return fail(VOXE_ERR_INVALID_ARGUMENT, "%s: NULL buffer / signal pad of rank %d", what, k);
    if (reinterpret_cast<uintptr_t>(peers->buffers[k]) & 15) return fail(VOXE_ERR_INVALID_ARGUMENT, "%s: buffers must be 16-byte aligned", what);
  }
  if (reinterpret_cast<uintptr_t>(peers->multicast) & 15) return fail(VOXE_ERR_INVALID_ARGUMENT, "%s: multicast mapping must be 16-byte aligned", what);
  return VOXE_OK;
}

int voxe_allreduce_grads_peer(const VoxePeerDesc* peers, int64_t n_floats, uint32_t* fail_flag, voxe_stream_t stream) {
  if (int rc = check_peers(peers, n_floats, "voxe_allreduce_grads_peer")) return rc;
  if (n_floats == 0 || peers->world_size == 1) return VOXE_OK;
  cudaError_t e = voxe::launch_allreduce_peer(*peers, n_floats, fail_flag, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "voxe_allreduce_grads_peer launch");
  g_launches.fetch_add(1);
  return VOXE_OK;
}

int64_t voxe_peer_touched_bytes(const VoxeGridDesc* grid) { return (voxe_touched_bytes(grid) + 15) / 16 * 16; }

int voxe_allreduce_grads_peer_sparse(const VoxePeerDesc* peers, const VoxeGridDesc* grid, uint8_t* const* touched_peers,
                                     int32_t touch_tag, uint32_t* fail_flag, voxe_stream_t stream) {
  const char* what = "voxe_allreduce_grads_peer_sparse";
  if (int rc = check_peers(peers, 0, what)) return rc;
  if (int rc = check_grid(grid)) return rc;
  if (touch_tag < 1 || touch_tag > 255) return fail(VOXE_ERR_INVALID_ARGUMENT, "%s: touch_tag must be in 1..255", what);
  if (!touched_peers) return fail(VOXE_ERR_INVALID_ARGUMENT, "%s: NULL touched_peers", what);
  for (int k = 0; k < peers->world_size; ++k)
    if (!touched_peers[k] || (reinterpret_cast<uintptr_t>(touched_peers[k]) & 15))
      return fail(VOXE_ERR_INVALID_ARGUMENT, "%s: flag array of rank %d is NULL or not 16-byte aligned", what, k);
  if (peers->world_size == 1) return VOXE_OK;
  cudaError_t e = voxe::launch_allreduce_peer_sparse(*peers, touched_peers, touch_tag, grid->dims, grid->channels, fail_flag, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "voxe_allreduce_grads_peer_sparse launch");
  g_launches.fetch_add(2);
  return VOXE_OK;
}

static int nccl_fail(int rc, const char* what) {
  return fail(VOXE_ERR_NCCL_BASE + rc, "%s: NCCL error %d (%s)", what, rc, voxe::nccl_error_string(rc));
}

int voxe_nccl_unique_id(void* id_out) {
  if (!id_out) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_nccl_unique_id: NULL output");
  if (const char* why = voxe::nccl_unavailable()) return fail(VOXE_ERR_UNSUPPORTED, "NCCL is not available: %s", why);
  if (int rc = voxe::nccl_unique_id(id_out)) return nccl_fail(rc, "ncclGetUniqueId");
  return VOXE_OK;
}

int voxe_nccl_comm_create(void** nccl_comm_out, int32_t world_size, int32_t rank, const void* id) {
  if (!nccl_comm_out || !id || world_size < 1 || rank < 0 || rank >= world_size)
    return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_nccl_comm_create: bad arguments");
  if (const char* why = voxe::nccl_unavailable()) return fail(VOXE_ERR_UNSUPPORTED, "NCCL is not available: %s", why);
  if (int rc = voxe::nccl_comm_create(nccl_comm_out, world_size, rank, id)) return nccl_fail(rc, "ncclCommInitRank");
  return VOXE_OK;
}

int voxe_nccl_comm_destroy(void* nccl_comm) {
  if (!nccl_comm) return VOXE_OK;
  if (const char* why = voxe::nccl_unavailable()) return fail(VOXE_ERR_UNSUPPORTED, "NCCL is not available: %s", why);
  if (int rc = voxe::nccl_comm_destroy(nccl_comm)) return nccl_fail(rc, "ncclCommDestroy");
  return VOXE_OK;
}

int voxe_allreduce_grads(void* nccl_comm, float* buf, int64_t n_floats, voxe_stream_t stream) {
  if (!nccl_comm || !buf || n_floats < 0) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_allreduce_grads: NULL communicator / buffer");
  if (const char* why = voxe::nccl_unavailable()) return fail(VOXE_ERR_UNSUPPORTED, "NCCL is not available: %s", why);
  if (n_floats == 0) return VOXE_OK;
  if (int rc = voxe::nccl_allreduce_sum_f32(nccl_comm, buf, (size_t)n_floats, (cudaStream_t)stream)) return nccl_fail(rc, "ncclAllReduce");
  g_launches.fetch_add(1);
  return VOXE_OK;
}

int voxe_sample_rays(const VoxeSamplerDesc* sampler, const float* poses, const float* src_rays_o, const float* src_rays_d,
                     const float* pixels, const int64_t* indices_in, int64_t sample_size, int64_t* indices_out, float* rays_o,
                     float* rays_d, float* pixels_out, voxe_stream_t stream) {
  if (!sampler) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_sample_rays: NULL descriptor");
  if (sampler->num_pixels < 1) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_sample_rays: num_pixels must be >= 1");
  if (sample_size < 0 || sample_size > 0x7fffffff) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_sample_rays: sample_size out of range");
  if (!indices_in && sample_size > sampler->num_pixels)
    return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_sample_rays: cannot draw %lld distinct indices out of %lld",
                (long long)sample_size, (long long)sampler->num_pixels);
  if ((rays_o == nullptr) != (rays_d == nullptr)) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_sample_rays: rays_o and rays_d go together");
  if (rays_o) {
    if (poses) {
      const int64_t per = (int64_t)sampler->height * sampler->width;
      if (sampler->height < 1 || sampler->width < 1 || !(sampler->focal > 0.f) || sampler->num_pixels % per != 0)
        return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_sample_rays: camera mode needs H, W >= 1, focal > 0 and num_pixels = B*H*W");
    } else if (!src_rays_o || !src_rays_d) {
      return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_sample_rays: give poses (camera mode) or source rays (gather mode)");
    }
  }
  if (pixels_out && (!pixels || sampler->pixel_channels < 1))
    return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_sample_rays: pixels_out needs pixels and pixel_channels >= 1");
  if (sample_size == 0) return VOXE_OK;
  cudaError_t e = voxe::launch_sample_rays(sampler->num_pixels, sample_size, sampler->height, sampler->width, sampler->pixel_channels,
                                           sampler->focal, sampler->rng_seed, sampler->rng_offset, poses, src_rays_o, src_rays_d,
                                           pixels, reinterpret_cast<const long long*>(indices_in),
                                           reinterpret_cast<long long*>(indices_out), rays_o, rays_d, pixels_out, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "voxe_sample_rays launch");
  g_launches.fetch_add(1);
  return VOXE_OK;
}

int voxe_render_fwd(const VoxeGridDesc* grid, const VoxeRenderDesc* render, const float* packed,
                    const float* rays_o, const float* rays_d, const float* jitter, const float* noise,
                    float* colour, float* depth, float* acc, float* disparity, float* saved, int64_t num_rays,
                    voxe_stream_t stream) {
  if (int rc = check_grid(grid)) return rc;
  if (int rc = check_render(grid, render, jitter, noise)) return rc;
  if (num_rays < 0 || num_rays > 0x7fffffff) return fail(VOXE_ERR_INVALID_ARGUMENT, "num_rays out of range");
  if (num_rays == 0) return VOXE_OK;
  if (!packed || !rays_o || !rays_d || !colour || !depth || !acc)
    return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_render_fwd: NULL buffer");
  voxe::KParams p;
  int regcap = 0;
  if (int rc = fill_params(grid, render, num_rays, false, p, regcap)) return rc;
  p.grid = reinterpret_cast<const float4*>(packed);
  p.rays_o = rays_o;
  p.rays_d = rays_d;
  p.jitter = (render->flags & VOXE_FLAG_PERTURB) ? jitter : nullptr;
  p.noise = (render->noise_std != 0.f) ? noise : nullptr;
  p.saved = saved;
  p.colour = colour;
  p.depth = depth;
  p.acc = acc;
  p.disp = disparity;
  bool took_specialised = false;
  cudaError_t e = voxe::launch_render(p, render->sh_degree, render->n_colour, regcap, false, specialised_kernels(), (cudaStream_t)stream,
                                      &took_specialised);
  if (took_specialised) g_specialised.fetch_add(1);
  if (e != cudaSuccess) return cuda_fail(e, "voxe_render_fwd launch");
  g_launches.fetch_add(1);
  return VOXE_OK;
}

int voxe_render_camera(const VoxeGridDesc* grid, const VoxeRenderDesc* render, const VoxeCameraDesc* camera,
                       const float* packed, int64_t first_pixel, int64_t num_pixels, float* colour, float* depth,
                       float* acc, float* disparity, float min_transmittance, voxe_stream_t stream) {
  if (int rc = check_grid(grid)) return rc;
  if (int rc = check_render(grid, render, nullptr, nullptr)) return rc;
  if (!camera || camera->height < 1 || camera->width < 1 || !(camera->focal > 0.f))
    return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_render_camera: camera needs height, width >= 1 and focal > 0");
  if (render->noise_std != 0.f) return fail(VOXE_ERR_UNSUPPORTED, "voxe_render_camera does not take density noise; use voxe_render_fwd");
  const int64_t total = (int64_t)camera->height * camera->width;
  if (first_pixel < 0 || num_pixels < 0 || first_pixel + num_pixels > total || num_pixels > 0x7fffffff)
    return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_render_camera: pixel range outside the %d x %d image", camera->height, camera->width);
  if (num_pixels == 0) return VOXE_OK;
  if (!packed || !colour || !depth || !acc) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_render_camera: NULL buffer");
  if (!(min_transmittance >= 0.f)) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_render_camera: min_transmittance must be >= 0");
  voxe::KParams p;
  int regcap = 0;
  if (int rc = fill_params(grid, render, num_pixels, false, p, regcap)) return rc;
  p.grid = reinterpret_cast<const float4*>(packed);
  p.colour = colour;
  p.depth = depth;
  p.acc = acc;
  p.disp = disparity;
  voxe::CameraParams cam;
  cam.H = camera->height;
  cam.W = camera->width;
  cam.focal = camera->focal;
  for (int k = 0; k < 9; ++k) cam.rot[k] = camera->rotation[k];
  for (int k = 0; k < 3; ++k) cam.trans[k] = camera->translation[k];
  cam.first_pixel = first_pixel;
  cam.min_transmittance = min_transmittance;
  cudaError_t e = voxe::launch_camera(p, cam, render->sh_degree, render->n_colour, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "voxe_render_camera launch");
  g_launches.fetch_add(1);
  return VOXE_OK;
}

int voxe_render_infer(const VoxeGridDesc* grid, const VoxeRenderDesc* render, const float* packed, const float* rays_o,
                      const float* rays_d, const float* jitter, float* colour, float* depth, float* acc, float* disparity,
                      int64_t num_rays, float min_transmittance, voxe_stream_t stream) {
  if (int rc = check_grid(grid)) return rc;
  if (int rc = check_render(grid, render, jitter, nullptr)) return rc;
  if (render->noise_std != 0.f) return fail(VOXE_ERR_UNSUPPORTED, "voxe_render_infer does not take density noise; use voxe_render_fwd");
  if (num_rays < 0 || num_rays > 0x7fffffff) return fail(VOXE_ERR_INVALID_ARGUMENT, "num_rays out of range");
  if (num_rays == 0) return VOXE_OK;
  if (!packed || !rays_o || !rays_d || !colour || !depth || !acc) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_render_infer: NULL buffer");
  if (!(min_transmittance >= 0.f)) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_render_infer: min_transmittance must be >= 0");
  voxe::KParams p;
  int regcap = 0;
  if (int rc = fill_params(grid, render, num_rays, false, p, regcap)) return rc;
  p.grid = reinterpret_cast<const float4*>(packed);
  p.rays_o = rays_o;
  p.rays_d = rays_d;
  p.jitter = (render->flags & VOXE_FLAG_PERTURB) ? jitter : nullptr;
  p.colour = colour;
  p.depth = depth;
  p.acc = acc;
  p.disp = disparity;
  voxe::CameraParams cam{};
  cam.W = 1;
  cam.min_transmittance = min_transmittance;
  cudaError_t e = voxe::launch_camera(p, cam, render->sh_degree, render->n_colour, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "voxe_render_infer launch");
  g_launches.fetch_add(1);
  return VOXE_OK;
}

int voxe_render_bwd(const VoxeGridDesc* grid, const VoxeRenderDesc* render, const float* packed,
                    const float* rays_o, const float* rays_d, const float* jitter, const float* noise,
                    const float* saved, const float* g_colour, const float* g_depth, const float* g_acc,
                    const float* g_disp, float* packed_grad, uint8_t* touched, int32_t touch_tag, int64_t num_rays,
                    voxe_stream_t stream) {
  if (int rc = check_grid(grid)) return rc;
  if (int rc = check_render(grid, render, jitter, noise)) return rc;
  if (touched && (touch_tag < 1 || touch_tag > 255)) return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_render_bwd: touch_tag must be in 1..255");
  if (num_rays < 0 || num_rays > 0x7fffffff) return fail(VOXE_ERR_INVALID_ARGUMENT, "num_rays out of range");
  if (num_rays == 0) return VOXE_OK;
  if (!packed || !rays_o || !rays_d || !g_colour || !packed_grad || !saved)
    return fail(VOXE_ERR_INVALID_ARGUMENT, "voxe_render_bwd: NULL buffer (saved is the workspace filled by voxe_render_fwd)");
  voxe::KParams p;
  int regcap = 0;
  if (int rc = fill_params(grid, render, num_rays, true, p, regcap)) return rc;
  p.saved = const_cast<float*>(saved);
  p.grid = reinterpret_cast<const float4*>(packed);
  p.grad = reinterpret_cast<float4*>(packed_grad);
  p.rays_o = rays_o;
  p.rays_d = rays_d;
  p.jitter = (render->flags & VOXE_FLAG_PERTURB) ? jitter : nullptr;
  p.noise = (render->noise_std != 0.f) ? noise : nullptr;
  p.g_colour = g_colour;
  p.g_depth = g_depth;
  p.g_acc = g_acc;
  p.g_disp = g_disp;
  p.touched = touched;
  p.touch_tag = touch_tag;
  p.stats = reinterpret_cast<unsigned long long*>(render->stats);
  bool took_specialised = false;
  cudaError_t e = voxe::launch_render(p, render->sh_degree, render->n_colour, regcap, true, specialised_kernels(), (cudaStream_t)stream,
                                      &took_specialised);
  if (took_specialised) g_specialised.fetch_add(1);
  if (e != cudaSuccess) return cuda_fail(e, "voxe_render_bwd launch");
  g_launches.fetch_add(1);
  return VOXE_OK;
}

}  // extern "C"
