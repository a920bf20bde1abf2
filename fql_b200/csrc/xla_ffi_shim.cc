// xla_ffi_shim.cc -- jax.ffi (XLA typed FFI) handlers over the C ABI of include/fql_b200.h.
//
// NOT built by default and NOT testable in this image: jaxlib (which ships xla/ffi/api/ffi.h) is not installed
// (SURVEY F1/F2).  Build where jax is present:
//   g++ -O2 -shared -fPIC -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -Iinclude \
//       fql_b200/csrc/xla_ffi_shim.cc -Lfql_b200 -lfql_b200 -o fql_b200/libfql_b200_xla.so
// Python side: see INTEGRATION.md ("jax.ffi binding").
#if __has_include("xla/ffi/api/ffi.h")
#include "xla/ffi/api/ffi.h"

#include <cuda_runtime.h>

#include "../../include/fql_b200.h"

namespace ffi = xla::ffi;

static FqlContext* ctx_for_thread() {
  static thread_local FqlContext* ctx = nullptr;
  if (!ctx) fql_context_create(&ctx);
  return ctx;
}

// One FQLAgent.update (agents/fql.py:122-133).  Operands: the four arenas + count (aliased to the results with
// input_output_aliases), the batch and the five noise tensors, a workspace; attributes: the static config.
static ffi::Error UpdateStepImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> mu, ffi::Buffer<ffi::F32> nu,
                                 ffi::Buffer<ffi::F32> grads, ffi::Buffer<ffi::S32> count, ffi::Buffer<ffi::F32> obs,
                                 ffi::Buffer<ffi::F32> actions, ffi::Buffer<ffi::F32> next_obs, ffi::Buffer<ffi::F32> rewards,
                                 ffi::Buffer<ffi::F32> masks, ffi::Buffer<ffi::F32> z_next, ffi::Buffer<ffi::F32> x0, ffi::Buffer<ffi::F32> t,
                                 ffi::Buffer<ffi::F32> z, ffi::Buffer<ffi::F32> z_metric, ffi::Buffer<ffi::U8> workspace,
                                 ffi::Buffer<ffi::U8> shadow, int32_t batch, int32_t obs_dim, int32_t action_dim, int32_t hidden,
                                 int32_t num_hidden, bool critic_layer_norm, bool actor_layer_norm, bool q_agg_min, bool normalize_q_loss,
                                 int32_t flow_steps, int32_t num_seeds, int32_t precision, float lr, float discount, float tau, float alpha,
                                 ffi::ResultBuffer<ffi::F32> info) {
  FqlDims d{};
  d.batch = d.global_batch = batch; d.obs_dim = obs_dim; d.action_dim = action_dim; d.hidden = hidden; d.num_hidden = num_hidden;
  d.critic_layer_norm = critic_layer_norm; d.actor_layer_norm = actor_layer_norm; d.q_agg_min = q_agg_min;
  d.normalize_q_loss = normalize_q_loss; d.flow_steps = flow_steps; d.num_seeds = num_seeds; d.precision = precision;
  FqlHparams hp{};
  hp.lr = lr; hp.beta1 = 0.9f; hp.beta2 = 0.999f; hp.eps = 1e-8f; hp.discount = discount; hp.tau = tau; hp.alpha = alpha;
  hp.one_minus_beta1 = (float)(1.0 - 0.9); hp.one_minus_beta2 = (float)(1.0 - 0.999); hp.one_minus_tau = (float)(1.0 - (double)tau);
  FqlBatch b{obs.typed_data(), actions.typed_data(), next_obs.typed_data(), rewards.typed_data(), masks.typed_data(),
             z_next.typed_data(), x0.typed_data(), t.typed_data(), z.typed_data(), z_metric.typed_data()};
  FqlState st{params.typed_data(), mu.typed_data(), nu.typed_data(), grads.typed_data(), count.typed_data(),
              shadow.element_count() ? (void*)shadow.typed_data() : nullptr};
  if (fql_update_step(ctx_for_thread(), &d, &hp, &b, &st, info->typed_data(), workspace.typed_data(), workspace.element_count(), stream))
    return ffi::Error(ffi::ErrorCode::kInternal, fql_last_error());
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    fql_update_step_ffi, UpdateStepImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::S32>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::U8>>().Arg<ffi::Buffer<ffi::U8>>()
        .Attr<int32_t>("batch").Attr<int32_t>("obs_dim").Attr<int32_t>("action_dim").Attr<int32_t>("hidden").Attr<int32_t>("num_hidden")
        .Attr<bool>("critic_layer_norm").Attr<bool>("actor_layer_norm").Attr<bool>("q_agg_min").Attr<bool>("normalize_q_loss")
        .Attr<int32_t>("flow_steps").Attr<int32_t>("num_seeds").Attr<int32_t>("precision")
        .Attr<float>("lr").Attr<float>("discount").Attr<float>("tau").Attr<float>("alpha")
        .Ret<ffi::Buffer<ffi::F32>>());
#else
// xla/ffi/api/ffi.h not found: nothing to build (see the header comment).
#endif
