// voxe_torch.cpp -- host-side bridge between torch (tensors, autograd, streams, RNG) and the C ABI of
// libvoxe_sm100a.so.  Built as the Python extension module `voxe_b200._voxe_torch`.
//
// It replaces, for one render call, everything the reference does between `render_sh_voxel_grid(...)`
// (thre3d_atom/thre3d_reprs/renderers.py:50-105) and the ATen kernels: the sampler -> processor -> accumulator
// chain of render_interface.py:140-171 and its autograd graph become ONE autograd node (RenderNode below) created by a
// voxe_render_fwd launch and whose backward is a voxe_render_bwd launch.  The node lives in C++ so that a 4096-ray
// training batch costs a few microseconds of host time per direction instead of ~100 (Python autograd.Function + ctypes
// marshalling).
//
// Gradient hand-over (backward): the kernel scatter-adds into the grid's persistent packed gradient volume
// (always all-zero between calls), then one of
//   kSink    deferred gradients: leave them there; whoever owns the volume materialises / consumes them later
//            (VoxelGrid.materialize_render_gradients, FusedVoxelAdam);
//   kDirect  plain `loss.backward()`: add the touched voxels straight into `param.grad` (voxe_consume_grad) -- what
//            AccumulateGrad would do with a dense gradient, minus three full-grid passes per call.  Only taken when the
//            engine is accumulating into leaves (no `inputs=` / torch.autograd.grad) and the parameter carries no hooks;
//   dense    otherwise: return dense gradients shaped like the parameters, exactly as autograd expects.
#include <torch/extension.h>

#include <ATen/cuda/CUDAGeneratorImpl.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/csrc/autograd/functions/utils.h>
#include <torch/csrc/autograd/graph_task.h>

#include <cstring>
#include <string>

#include "voxe.h"

namespace {

using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

enum : int64_t { kDense = 0, kDirect = 1, kSink = 2 };

void check(int rc, const char* what) {
  if (rc == VOXE_OK) return;
  const std::string msg = std::string(what) + " failed with code " + std::to_string(rc) + ": " + voxe_last_error();
  TORCH_CHECK_NOT_IMPLEMENTED(rc != VOXE_ERR_UNSUPPORTED, msg);
  TORCH_CHECK(false, msg);
}

const float* cptr(const Tensor& t) { return t.defined() ? t.data_ptr<float>() : nullptr; }
float* mptr(const Tensor& t) { return t.defined() ? t.data_ptr<float>() : nullptr; }

// Checked, gradient-free, contiguous view of an input -- without creating tensor objects when it already is all that (the
// common case: every dispatcher call costs about a microsecond of the ~40 a forward call takes on the host).
Tensor prep(const Tensor& t, const c10::Device& dev, const char* name) {
  TORCH_CHECK(t.device() == dev, "all render inputs must live on one device (", name, " is on ", t.device(), ", expected ", dev, ")");
  TORCH_CHECK_TYPE(t.scalar_type() == at::kFloat, "the render path computes in fp32 (", name, " is ", t.scalar_type(), ")");
  if (!t.requires_grad() && t.is_contiguous()) return t;
  return t.detach().contiguous();
}

voxe_stream_t stream_of(const c10::Device& dev) {
  return reinterpret_cast<voxe_stream_t>(c10::cuda::getCurrentCUDAStream(dev.index()).stream());
}

template <class T>
T unpack_desc(const std::string& bytes) {
  T d;
  TORCH_INTERNAL_ASSERT(bytes.size() == sizeof(T));
  std::memcpy(&d, bytes.data(), sizeof(T));
  return d;
}

struct Outputs {
  Tensor colour, depth, acc, disp;
};

Outputs run_forward(const VoxeGridDesc& gd, const VoxeRenderDesc& rd, const Tensor& packed, const Tensor& rays_o,
                    const Tensor& rays_d, const Tensor& jitter, const Tensor& noise, const Tensor& saved) {
  const int64_t R = rays_o.size(0);
  const auto opts = rays_o.options();
  Outputs o{at::empty({R, (int64_t)rd.n_colour}, opts), at::empty({R, 1}, opts), at::empty({R, 1}, opts), at::empty({R, 1}, opts)};
  check(voxe_render_fwd(&gd, &rd, cptr(packed), cptr(rays_o), cptr(rays_d), cptr(jitter), cptr(noise), mptr(o.colour),
                        mptr(o.depth), mptr(o.acc), mptr(o.disp), mptr(saved), R, stream_of(rays_o.device())),
        "voxe_render_fwd");
  return o;
}

// True when the running backward pass accumulates into every leaf it reaches (`tensor.backward()` without `inputs=`).
bool engine_accumulates_into_leaves() {
  const auto* exec_info = torch::autograd::get_current_graph_task_exec_info();
  return exec_info == nullptr || exec_info->empty();
}

// May this node add into `param.grad` itself instead of handing a dense gradient to AccumulateGrad?
bool direct_ok(const Tensor& param) {
  if (!param.defined() || !param.is_leaf() || !param.requires_grad()) return false;
  if (!torch::autograd::impl::hooks(param).empty()) return false;
  if (torch::autograd::impl::post_acc_grad_hooks(param)) return false;
  if (auto acc = torch::autograd::impl::try_get_grad_accumulator(param)) {
    if (!acc->pre_hooks().empty() || !acc->post_hooks().empty() || !acc->tensor_pre_hooks().empty()) return false;
  }
  const Tensor& g = param.grad();
  if (!g.defined()) return true;
  return g.is_contiguous() && g.scalar_type() == at::kFloat && g.device() == param.device() && g.sizes() == param.sizes() &&
         g.layout() == at::kStrided && !g.requires_grad();
}

// The autograd node of one render call.  A hand-written torch::autograd::Node rather than a torch::autograd::Function:
// the per-call cost of the Function machinery (AutogradContext, SavedVariable wrapping of eight tensors, string-keyed
// saved_data, output wrapping and validation -- ~15 us per forward/backward pair on the bench hosts) is a third of the host
// time of a 4096-ray batch, and none of it is needed here: the only inputs autograd differentiates are the two parameter
// tensors, everything else the backward reads is private to this node.
struct RenderNode : public torch::autograd::Node {
  Tensor densities, features;  // the parameters this render read (next edges 0 and 1)
  uint32_t dens_version = 0, feat_version = 0;
  Tensor packed, rays_o, rays_d, jitter, noise, work, grad_volume, dirty_flag, touched, touch_tag;
  VoxeGridDesc gd{};
  VoxeRenderDesc rd{};  // carries this call's (rng_seed, rng_offset)
  int64_t mode = kDense;
  int64_t work_floats = 0;
  bool released = false;

  std::string name() const override { return "VoxeRenderBackward"; }

  void release_variables() override {
    std::lock_guard<std::mutex> lock(mutex_);
    released = true;
    packed.reset(); rays_o.reset(); rays_d.reset(); jitter.reset(); noise.reset(); work.reset();
    grad_volume.reset(); dirty_flag.reset(); touched.reset(); touch_tag.reset();
  }

  variable_list apply(variable_list&& grads) override {
    std::lock_guard<std::mutex> lock(mutex_);
    variable_list out(2);  // dL/d densities, dL/d features (undefined = no gradient through that edge)
    TORCH_CHECK(!released, "Trying to backward through the render a second time (its workspace has already been freed). "
                           "Specify retain_graph=True when calling backward the first time.");
    const bool need_d = task_should_compute_output(0), need_f = task_should_compute_output(1);
    if (!need_d && !need_f) return out;
    bool any = false;
    for (const auto& g : grads) any |= g.defined();
    if (!any) return out;
    // what SavedVariable would have checked for the two parameters: an in-place write (an optimiser step, say) between
    // this render's forward and its backward would make the backward differentiate another function
    TORCH_CHECK((!densities.defined() || densities._version() == dens_version) && (!features.defined() || features._version() == feat_version),
                "one of the variables needed for gradient computation has been modified by an inplace operation: the voxel grid's "
                "densities / features changed between a render's forward and its backward");
    const auto dev = packed.device();
    const c10::cuda::CUDAGuard guard(dev);
    const int64_t R = rays_o.size(0);
    // the workspace layout follows the launch shape, which voxe_set_tuning can change between the two calls
    TORCH_CHECK(voxe_saved_floats(&rd, R) == work_floats,
                "voxe_set_tuning changed the launch shape between a render's forward and its backward (workspace of ", work_floats,
                " floats, the backward now expects ", voxe_saved_floats(&rd, R), "); retune only between whole forward/backward pairs");

    Tensor g[4];
    for (int k = 0; k < 4; ++k)
      if (grads[k].defined())
        g[k] = (grads[k].scalar_type() == at::kFloat && grads[k].is_contiguous()) ? grads[k] : grads[k].to(at::kFloat).contiguous();
    if (!g[0].defined()) g[0] = at::zeros({R, (int64_t)rd.n_colour}, rays_o.options());
    Tensor volume = grad_volume;
    if (!volume.defined()) volume = at::zeros_like(packed);  // no persistent volume attached: a fresh one

    const auto stream = stream_of(dev);
    // Sparse hand-over: the kernel tags the bricks it scatters into and voxe_consume_grad visits only those.  A fresh tag
    // per backward; stale tags of earlier calls only cost the consume pass a few reads (it skips all-zero vectors), so the
    // flags are never cleared.
    // A sink that keeps a trail (deferred gradients on a large grid: the step's exchange and its hand-over follow the flags,
    // PackedGradAccumulator.sparse_sink) tags every backward of an optimiser step alike; the accumulator advances the tag.
    const bool sparse = mode != kSink && touched.defined() && grad_volume.defined();
    const bool sink_trail = mode == kSink && touched.defined() && touch_tag.defined() && grad_volume.defined();
    int32_t tag = 0;
    if (sparse) {
      int64_t* last = touch_tag.data_ptr<int64_t>();
      *last = (*last % 255) + 1;
      tag = (int32_t)*last;
    } else if (sink_trail) {
      tag = (int32_t)touch_tag.data_ptr<int64_t>()[0];
      TORCH_CHECK(tag >= 1 && tag <= 255, "the gradient sink's brick-flag tag must be in 1..255 (got ", tag, ")");
    }
    uint8_t* touched_ptr = (sparse || sink_trail) ? touched.data_ptr<uint8_t>() : nullptr;
    check(voxe_render_bwd(&gd, &rd, cptr(packed), cptr(rays_o), cptr(rays_d), cptr(jitter), cptr(noise), cptr(work),
                          cptr(g[0]), cptr(g[1]), cptr(g[2]), cptr(g[3]), mptr(volume), touched_ptr, tag, R, stream),
          "voxe_render_bwd");
    if (mode == kSink) {
      if (dirty_flag.defined()) dirty_flag.data_ptr<int64_t>()[0] = 1;  // CPU flag owned by the accumulator
      return out;
    }
    const bool direct = mode == kDirect && engine_accumulates_into_leaves() && (!need_d || direct_ok(densities)) &&
                        (!need_f || direct_ok(features));
    Tensor d_dens, d_feat;
    if (direct) {
      // AccumulateGrad's job, done sparsely: create a zero gradient on first use, then add what this call touched.
      if (need_d && !densities.grad().defined()) densities.mutable_grad() = at::zeros(densities.sizes(), densities.options());
      if (need_f && !features.grad().defined()) features.mutable_grad() = at::zeros(features.sizes(), features.options());
      if (need_d) d_dens = densities.grad();
      if (need_f) d_feat = features.grad();
    } else {
      if (need_d) d_dens = at::zeros(densities.sizes(), densities.options());
      if (need_f) d_feat = at::zeros(features.sizes(), features.options());
    }
    check(voxe_consume_grad(&gd, mptr(volume), mptr(d_dens), mptr(d_feat), touched_ptr, tag, stream), "voxe_consume_grad");
    if (!direct) {
      out[0] = d_dens;
      out[1] = d_feat;
    }
    return out;
  }
};

// colour [R,C], depth [R,1], acc [R,1], disparity [R,1] = render(...).  `gdesc` / `rdesc` are the raw bytes of a
// VoxeGridDesc / VoxeRenderDesc (built once per distinct description on the Python side).  `jitter` / `noise` may be
// None.  Stratified jitter is then generated inside the kernels from a (seed, offset) pair taken from `generator` (or
// the device's default generator, which is advanced, so torch.manual_seed reproduces a run); with `strict_rng` the
// reference's own draws are made instead -- at::rand [R,S] for the jitter and at::randn [R,S] for the density noise
// (drawn and discarded when noise_std == 0), the same generator, shapes and order as sample.py:63 and
// accumulate.py:59-62.  Density noise with noise_std != 0 is always an at::randn draw.
std::vector<Tensor> render(const Tensor& densities, const Tensor& features, const Tensor& packed, const Tensor& rays_o_in,
                           const Tensor& rays_d_in, const c10::optional<Tensor>& jitter_in, const c10::optional<Tensor>& noise_in,
                           const c10::optional<Tensor>& grad_volume, const c10::optional<Tensor>& dirty_flag,
                           const c10::optional<Tensor>& touched, const c10::optional<Tensor>& touch_tag,
                           const std::string& gdesc, const std::string& rdesc, int64_t mode, bool strict_rng,
                           const c10::optional<at::Generator>& generator) {
  const auto dev = packed.device();
  TORCH_CHECK(dev.is_cuda(), "the fused Vox-E render path runs on CUDA only (tensors are on '", dev,
              "'); there is deliberately no CPU fallback in this package");
  TORCH_CHECK(gdesc.size() == sizeof(VoxeGridDesc) && rdesc.size() == sizeof(VoxeRenderDesc), "descriptor size mismatch (ABI)");
  auto rd = unpack_desc<VoxeRenderDesc>(rdesc);
  TORCH_CHECK(rays_o_in.dim() == 2 && rays_d_in.dim() == 2 && rays_o_in.size(1) == 3 && rays_o_in.sizes() == rays_d_in.sizes(),
              "Please note that the RENDER interface only works with FLAT RAYS!");
  const c10::cuda::CUDAGuard guard(dev);
  Tensor rays_o, rays_d, jitter, noise;
  {
    at::NoGradGuard no_grad;
    rays_o = prep(rays_o_in, dev, "ray origins");
    rays_d = prep(rays_d_in, dev, "ray directions");
    const int64_t R = rays_o.size(0), S = rd.num_samples;
    if (rd.flags & VOXE_FLAG_PERTURB) {
      if (jitter_in.has_value() && jitter_in->defined()) {
        jitter = prep(*jitter_in, dev, "jitter");
      } else if (strict_rng) {
        jitter = at::rand({R, S}, generator, rays_o.options());
      } else {  // in-kernel draws: take (seed, offset) from the generator and advance it
        auto* gen = at::get_generator_or_default<at::CUDAGeneratorImpl>(generator, at::cuda::detail::getDefaultCUDAGenerator(dev.index()));
        std::lock_guard<std::mutex> lock(gen->mutex_);
        const at::PhiloxCudaState st = gen->philox_cuda_state(4);
        if (st.captured_) {
          // under CUDA-graph capture the generator's state lives in device memory and is advanced by the graph on every
          // replay (the generator is registered with the graph by torch): the kernels read it there, so replays of a
          // captured render draw fresh jitter, forward and backward of one call the same
          rd.rng_seed_dev = st.seed_.ptr;
          rd.rng_offset_dev = st.offset_.ptr;
          rd.rng_offset_intragraph = st.offset_intragraph_;
        } else {
          rd.rng_seed = st.seed_.val;
          rd.rng_offset = st.offset_.val;
        }
      }
      TORCH_CHECK(!jitter.defined() || (jitter.dim() == 2 && jitter.size(0) == R && jitter.size(1) == S), "jitter must be [R, S]");
    }
    if (rd.noise_std != 0.f) {
      noise = noise_in.has_value() && noise_in->defined() ? prep(*noise_in, dev, "noise")
                                                          : at::randn({R, S}, generator, rays_o.options());
      TORCH_CHECK(noise.dim() == 2 && noise.size(0) == R && noise.size(1) == S, "noise must be [R, S]");
    } else if (strict_rng) {
      at::randn({R, S}, generator, rays_o.options());  // drawn and discarded, as upstream
    }
  }
  const bool differentiable = at::GradMode::is_enabled() && ((densities.defined() && densities.requires_grad()) ||
                                                             (features.defined() && features.requires_grad()));
  const auto gd = unpack_desc<VoxeGridDesc>(gdesc);
  if (!differentiable) {
    Outputs o = run_forward(gd, rd, packed, rays_o, rays_d, jitter, noise, Tensor());
    return {o.colour, o.depth, o.acc, o.disp};
  }
  const int64_t R = rays_o.size(0);
  auto node = std::shared_ptr<RenderNode>(new RenderNode(), torch::autograd::deleteNode);
  node->set_next_edges(torch::autograd::collect_next_edges(densities, features));
  node->densities = densities;
  node->features = features;
  node->dens_version = densities.defined() ? densities._version() : 0;
  node->feat_version = features.defined() ? features._version() : 0;
  node->work_floats = voxe_saved_floats(&rd, R);
  Outputs o;
  {
    at::NoGradGuard no_grad;
    node->work = at::empty({node->work_floats}, rays_o.options());
    o = run_forward(gd, rd, packed, rays_o, rays_d, jitter, noise, node->work);
  }
  node->packed = packed;
  node->rays_o = rays_o;
  node->rays_d = rays_d;
  node->jitter = jitter;
  node->noise = noise;
  node->grad_volume = grad_volume.value_or(Tensor());
  node->dirty_flag = dirty_flag.value_or(Tensor());
  node->touched = touched.value_or(Tensor());      // uint8 [bricks], device: trail of the backward's scatter
  node->touch_tag = touch_tag.value_or(Tensor());  // int64 [1], host: last tag handed out for `touched`
  node->gd = gd;
  node->rd = rd;
  node->mode = mode;
  torch::autograd::set_history({o.colour, o.depth, o.acc, o.disp}, node);
  return {o.colour, o.depth, o.acc, o.disp};
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "torch <-> libvoxe_sm100a.so bridge (autograd node of the fused ray-marcher)";
  m.attr("MODE_DENSE") = (int64_t)kDense;
  m.attr("MODE_DIRECT") = (int64_t)kDirect;
  m.attr("MODE_SINK") = (int64_t)kSink;
  m.def("render", &render, py::arg("densities"), py::arg("features"), py::arg("packed"), py::arg("rays_o"), py::arg("rays_d"),
        py::arg("jitter"), py::arg("noise"), py::arg("grad_volume"), py::arg("dirty_flag"), py::arg("touched"), py::arg("touch_tag"),
        py::arg("gdesc"), py::arg("rdesc"),
        py::arg("mode"), py::arg("strict_rng"), py::arg("generator"));
  m.def("abi_version", []() { return voxe_abi_version(); });
}
