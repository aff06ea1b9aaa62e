// returns.cu - K-RET: n-step discounted returns and advantages (objectives.py:123-130, 178-214).
// The reference builds [E,T,T] discount matrices in a py_func and multiplies; the value is exactly the
// reverse recursion R_T = V(s_T), R_t = r_t + gamma * (1 - term_t) * R_{t+1}, evaluated here
// by one thread per environment in fp32.  Algorithmic bytes per update: 17*N + 4*E.
#include "common.cuh"

namespace acx {

__global__ void returns_kernel(const float* __restrict__ rewards, const uint8_t* __restrict__ terminals,
                               const float* __restrict__ values, const float* __restrict__ bootstrap, float gamma,
                               int num_envs, int num_steps, float* __restrict__ targets, float* __restrict__ adv) {
  pdl_enter();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= num_envs) return;
  float run = bootstrap[e];
  const size_t base = (size_t)e * num_steps;
  for (int t = num_steps - 1; t >= 0; --t) {
    if (terminals[base + t]) run = 0.0f;
    run = __fadd_rn(rewards[base + t], __fmul_rn(gamma, run));
    if (targets) targets[base + t] = run;
    if (adv) adv[base + t] = run - values[base + t];
  }
}

int returns_launch(const float* rewards, const uint8_t* terminals, const float* values, const float* bootstrap,
                   float gamma, int num_envs, int num_steps, float* targets, float* adv, cudaStream_t st) {
  ACX_PDL_LAUNCH(returns_kernel, ceil_div(num_envs, 64), 64, 0, st, rewards, terminals, values, bootstrap, gamma, num_envs, num_steps, targets, adv);
  return 0;
}

}  // namespace acx

extern "C" int acx_returns_adv(const float* d_rewards, const uint8_t* d_terminals, const float* d_values,
                               const float* d_bootstrap_values, float gamma, int num_envs, int num_steps,
                               float* d_targets, float* d_advantages, void* stream) {
  ACX_CHECK(num_envs >= 0 && num_steps >= 0, "negative shape");
  if (num_envs == 0 || num_steps == 0) return 0;
  ACX_CHECK(d_rewards && d_terminals && d_bootstrap_values, "null input");
  ACX_CHECK(d_advantages == nullptr || d_values != nullptr, "advantages need values");
  return acx::returns_launch(d_rewards, d_terminals, d_values, d_bootstrap_values, gamma, num_envs, num_steps, d_targets,
                             d_advantages, reinterpret_cast<cudaStream_t>(stream));
}
