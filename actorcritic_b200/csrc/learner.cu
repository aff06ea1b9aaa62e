// learner.cu - placeholder until the learner engine lands (keeps the C ABI complete for the first GPU bring-up)
#include "common.cuh"
extern "C" {
size_t acx_learner_arena_bytes(const acx_learner_config_t*) { return 0; }
acx_learner_t* acx_learner_create(const acx_learner_config_t*, void*, size_t) { acx::set_error("learner not built yet"); return nullptr; }
void acx_learner_destroy(acx_learner_t*) {}
size_t acx_learner_num_params(const acx_learner_t*) { return 0; }
int acx_learner_set_params(acx_learner_t*, const float*, void*) { return 1; }
int acx_learner_get_params(acx_learner_t*, float*, void*) { return 1; }
float* acx_learner_buffer(acx_learner_t*, const char*, size_t*) { return nullptr; }
uint8_t* acx_learner_obs_buffer(acx_learner_t*, size_t*) { return nullptr; }
int acx_learner_phase1(acx_learner_t*, const int32_t*, const float*, void*) { return 1; }
int acx_learner_phase2(acx_learner_t*, void*) { return 1; }
int64_t acx_learner_global_step(const acx_learner_t*) { return 0; }
void acx_learner_set_global_step(acx_learner_t*, int64_t) {}
int acx_learner_act(acx_learner_t*, const uint8_t*, int, const float*, int, int32_t*, float*, float*, void*) { return 1; }
}
