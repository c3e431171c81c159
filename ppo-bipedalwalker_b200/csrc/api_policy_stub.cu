// TEMPORARY bring-up stub: policy exports return WB_ERR_UNSUPPORTED until mlp.cu/api_policy.cu land.
#include "common.h"
using namespace wb;
#define STUB(name, ...) int32_t name(__VA_ARGS__) { return fail(WB_ERR_UNSUPPORTED, #name ": policy kernels not built yet"); }
extern "C" {
STUB(wb_policy_create, int32_t, int32_t, const int32_t*, const int32_t*, int32_t, const int32_t*, const int32_t*, int32_t, const wb_hyperparams*, wb_policy**)
STUB(wb_policy_destroy, wb_policy*)
STUB(wb_policy_set_stream, wb_policy*, void*)
STUB(wb_policy_sync, wb_policy*)
STUB(wb_policy_set_hyperparams, wb_policy*, const wb_hyperparams*)
STUB(wb_policy_num_params, const wb_policy*, int32_t, int32_t*)
STUB(wb_policy_set_weights, wb_policy*, int32_t, const float*)
STUB(wb_policy_get_weights, wb_policy*, int32_t, float*)
STUB(wb_policy_get_grads, wb_policy*, int32_t, float*)
STUB(wb_policy_get_adam, wb_policy*, int32_t, float*, float*, int32_t*)
STUB(wb_policy_set_adam, wb_policy*, int32_t, const float*, const float*, const int32_t*)
STUB(wb_policy_forward, wb_policy*, int32_t, const float*, float*, float*)
STUB(wb_policy_forward_dev, wb_policy*, int32_t, const float*, float*, float*)
STUB(wb_policy_sample, wb_policy*, int32_t, const float*, const float*, float*, float*, float*)
STUB(wb_policy_sample_dev, wb_policy*, int32_t, const float*, const float*, float*, float*, float*)
STUB(wb_policy_sample_philox_dev, wb_policy*, int32_t, const float*, uint64_t, uint64_t, float*, float*, float*)
STUB(wb_ppo_grad, wb_policy*, int32_t, const float*, const float*, const float*, const float*, const float*, float*, int32_t*)
STUB(wb_ppo_grad_dev, wb_policy*, int32_t, const float*, const float*, const float*, const float*, const float*)
STUB(wb_adam_step, wb_policy*)
STUB(wb_policy_grad_buffer, wb_policy*, void**, int32_t*)
STUB(wb_policy_launch_count, const wb_policy*, int64_t*)
STUB(wb_returns_advantages, wb_policy*, int32_t, const float*, const float*, float*, float*)
}
