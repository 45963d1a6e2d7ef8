// Backward pass (placeholder until the hand-written gradients land).
#include "common.cuh"
using namespace chromo;
extern "C" int chromo_backward(const chromo_config_t* cfg, const float* params, const chromo_batch_t* in,
                               const float* dlogits, float* grads, float* workspace, int64_t workspace_floats,
                               int32_t flags, void* stream) {
    (void)cfg; (void)params; (void)in; (void)dlogits; (void)grads; (void)workspace; (void)workspace_floats;
    (void)flags; (void)stream;
    set_error("chromo_backward: not implemented yet");
    return CHROMO_EINVAL;
}
