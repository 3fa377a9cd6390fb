// Rollout-side kernels behind the C ABI (include/myo_b200.h, myo_gae / myo_running_moments_* / myo_vecnorm_reward):
// SB3 VecNormalize's running moments and reward scaling, and sb3-contrib's RecurrentRolloutBuffer GAE scan, on the
// device so the PPO rollout never leaves HBM. All three are streaming kernels: one pass
// over [n] or [T][n] arrays, coalesced along the world index, fp64 accumulation where SB3 keeps fp64 state.
//
// Reference semantics (third-party, restated in oracle/rollout_oracle.py):
//   stable_baselines3/common/running_mean_std.py  RunningMeanStd.update / update_from_moments
//   stable_baselines3/common/vec_env/vec_normalize.py  VecNormalize.step_wait / normalize_reward
//   sb3_contrib/common/recurrent/buffers.py -> stable_baselines3 RolloutBuffer.compute_returns_and_advantage
// reached from /root/reference/src/main_baoding.py:75 and /root/reference/src/train/trainer.py:67-71.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/myo_b200.h"

namespace myo { void set_error(const std::string& msg); }

namespace {

constexpr int kMomThreads = 256;
constexpr int kMomMaxCtas = 148;   // one per SM: the merge walks the partials serially in fp64 (~0.15 us each), so their count is what the update costs

#define RCK(call)                                                                 \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) {                                                      \
      myo::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));        \
      return MYO_E_CUDA;                                                          \
    }                                                                             \
  } while (0)

// Chan et al. pairwise merge of (count, mean, M2)
__device__ __forceinline__ void chan_merge(double& na, double& ma, double& m2a, double nb, double mb, double m2b) {
  if (nb == 0.0) return;
  if (na == 0.0) { na = nb; ma = mb; m2a = m2b; return; }
  const double n = na + nb, delta = mb - ma;
  ma += delta * nb / n;
  m2a += m2b + delta * delta * na * nb / n;
  na = n;
}

// Stage 1: per-CTA Welford moments of the columns of x[n][d] (optionally of y = ret * gamma + x for the reward path,
// which also writes y back). Threads are laid out (row lane, column): dc = columns per pass (power of two <= 256),
// rl = kMomThreads / dc row lanes; a thread walks its rows with stride rl * gridDim.x, so a warp reads dc contiguous
// floats of a row (coalesced). Row lanes are merged in lane order through shared memory: one partial per (CTA, column).
template <bool RETURNS>
__global__ void __launch_bounds__(kMomThreads) moments_partial_kernel(const float* __restrict__ x, double* __restrict__ ret, double gamma,
                                                                      int n, int d, int dc, double* __restrict__ part) {
  __shared__ double s_n[kMomThreads], s_m[kMomThreads], s_q[kMomThreads];
  const int rl = kMomThreads / dc, col0 = threadIdx.x % dc, lane = threadIdx.x / dc;
  for (int cbase = 0; cbase < d; cbase += dc) {
    const int col = cbase + col0;
    double cnt = 0.0, mean = 0.0, m2 = 0.0;
    if (col < d)
      for (int r = blockIdx.x * rl + lane; r < n; r += rl * gridDim.x) {
        double v = (double)x[(size_t)r * d + col];
        if (RETURNS) { v = ret[r] * gamma + v; ret[r] = v; }
        cnt += 1.0;
        const double dl = v - mean;
        mean += dl / cnt;
        m2 += dl * (v - mean);
      }
    s_n[threadIdx.x] = cnt; s_m[threadIdx.x] = mean; s_q[threadIdx.x] = m2;
    __syncthreads();
    if (lane == 0 && col < d) {
      for (int k = 1; k < rl; k++) chan_merge(cnt, mean, m2, s_n[k * dc + col0], s_m[k * dc + col0], s_q[k * dc + col0]);
      double* p = part + ((size_t)blockIdx.x * d + col) * 3;
      p[0] = cnt; p[1] = mean; p[2] = m2;
    }
    __syncthreads();
  }
}

// Stage 2: one thread per column merges the CTA partials in CTA order (deterministic), then folds the batch into the
// running state exactly as RunningMeanStd.update_from_moments does (batch_var = M2 / n, population variance).
// state = mean[d], var[d], count. Optionally refreshes the fp32 copies the policy's fused normalize_obs reads.
__global__ void moments_merge_kernel(const double* __restrict__ part, int nparts, int d, double* __restrict__ state,
                                     float* __restrict__ mean_f, float* __restrict__ var_f) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= d) return;
  double cnt = 0.0, mean = 0.0, m2 = 0.0;
  for (int k = 0; k < nparts; k++) {
    const double* p = part + ((size_t)k * d + col) * 3;
    chan_merge(cnt, mean, m2, p[0], p[1], p[2]);
  }
  const double count = state[2 * d];
  if (cnt > 0.0) {
    const double bvar = m2 / cnt, old_mean = state[col], old_var = state[d + col];
    const double delta = mean - old_mean, tot = count + cnt;
    const double new_mean = old_mean + delta * cnt / tot;
    const double M2 = old_var * count + bvar * cnt + delta * delta * count * cnt / tot;
    state[col] = new_mean;
    state[d + col] = M2 / tot;
    if (mean_f) mean_f[col] = (float)new_mean;
    if (var_f) var_f[col] = (float)(M2 / tot);
  }
}
__global__ void moments_count_kernel(double* state, int d, double add) { state[2 * d] += add; }

__global__ void moments_export_kernel(const double* __restrict__ state, int d, float* __restrict__ mean_f, float* __restrict__ var_f) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= d) return;
  mean_f[col] = (float)state[col];
  var_f[col] = (float)state[d + col];
}

// VecNormalize.normalize_reward + `returns[dones] = 0`
__global__ void reward_norm_kernel(const float* __restrict__ r, const uint8_t* __restrict__ done, double* __restrict__ ret,
                                   const double* __restrict__ state, double eps, double clip, int norm, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = r[i];
  if (norm) {
    const double s = sqrt(state[1] + eps);
    v = (float)fmin(fmax((double)v / s, -clip), clip);
  }
  out[i] = v;
  if (done[i]) ret[i] = 0.0;
}

// GAE(lambda) reverse scan, one thread per world, arrays [T][n] (SB3 buffer layout: step-major), coalesced over n.
__global__ void gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ starts,
                           const float* __restrict__ last_values, const uint8_t* __restrict__ last_dones, int T, int n, float gamma,
                           float lam, float* __restrict__ adv, float* __restrict__ returns) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  float next_value = last_values[w], next_nonterminal = 1.f - (float)(last_dones[w] != 0), gae = 0.f;
  for (int t = T - 1; t >= 0; t--) {
    const size_t i = (size_t)t * n + w;
    const float v = values[i];
    const float delta = rewards[i] + gamma * next_value * next_nonterminal - v;
    gae = delta + gamma * lam * next_nonterminal * gae;
    adv[i] = gae;
    returns[i] = gae + v;
    next_value = v;
    next_nonterminal = 1.f - (float)(starts[i] != 0);
  }
}

int moments_geometry(int n, int d, int* dc, int* ctas) {
  int c = 1;
  while (c < d && c < kMomThreads) c <<= 1;
  const int rl = kMomThreads / c;
  int g = (n + rl - 1) / rl;
  // enough rows per thread to amortise the merge, never more CTAs than the bound
  g = (g + 15) / 16;
  if (g < 1) g = 1;
  if (g > kMomMaxCtas) g = kMomMaxCtas;
  *dc = c; *ctas = g;
  return 0;
}

__global__ void normalize_obs_kernel(const float* __restrict__ obs, const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                     float clip, float* __restrict__ out, int64_t total, int d) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % d);
    const float v = (obs[i] - mean[c]) / sqrtf(var[c] + eps);
    out[i] = fminf(fmaxf(v, -clip), clip);
  }
}

}  // namespace

extern "C" {

int myo_running_moments_scratch(int n, int d) {
  if (n <= 0 || d <= 0) return 0;
  int dc, ctas;
  moments_geometry(n, d, &dc, &ctas);
  return ctas * d * 3;
}

int myo_running_moments_update(double* state_dev, const float* x_dev, int n, int d, double* scratch_dev, float* mean_f_dev,
                               float* var_f_dev, void* stream) {
  if (!state_dev || !x_dev || !scratch_dev || n <= 0 || d <= 0) { myo::set_error("bad argument to myo_running_moments_update"); return MYO_E_ARG; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dc, ctas;
  moments_geometry(n, d, &dc, &ctas);
  moments_partial_kernel<false><<<ctas, kMomThreads, 0, st>>>(x_dev, nullptr, 0.0, n, d, dc, scratch_dev);
  moments_merge_kernel<<<(d + 127) / 128, 128, 0, st>>>(scratch_dev, ctas, d, state_dev, mean_f_dev, var_f_dev);
  moments_count_kernel<<<1, 1, 0, st>>>(state_dev, d, (double)n);
  RCK(cudaGetLastError());
  return MYO_OK;
}

int myo_running_moments_export(const double* state_dev, int d, float* mean_f_dev, float* var_f_dev, void* stream) {
  if (!state_dev || !mean_f_dev || !var_f_dev || d <= 0) { myo::set_error("bad argument to myo_running_moments_export"); return MYO_E_ARG; }
  moments_export_kernel<<<(d + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(state_dev, d, mean_f_dev, var_f_dev);
  RCK(cudaGetLastError());
  return MYO_OK;
}

int myo_vecnorm_reward(double* ret_state_dev, double* returns_dev, const float* reward_dev, const uint8_t* done_dev, float* out_reward_dev,
                       int n, double gamma, double epsilon, double clip_reward, int training, int norm_reward, double* scratch_dev,
                       void* stream) {
  if (!ret_state_dev || !returns_dev || !reward_dev || !done_dev || !out_reward_dev || !scratch_dev || n <= 0) {
    myo::set_error("bad argument to myo_vecnorm_reward");
    return MYO_E_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (training) {
    int dc, ctas;
    moments_geometry(n, 1, &dc, &ctas);
    moments_partial_kernel<true><<<ctas, kMomThreads, 0, st>>>(reward_dev, returns_dev, gamma, n, 1, dc, scratch_dev);
    moments_merge_kernel<<<1, 128, 0, st>>>(scratch_dev, ctas, 1, ret_state_dev, nullptr, nullptr);
    moments_count_kernel<<<1, 1, 0, st>>>(ret_state_dev, 1, (double)n);
  }
  reward_norm_kernel<<<(n + 255) / 256, 256, 0, st>>>(reward_dev, done_dev, returns_dev, ret_state_dev, epsilon, clip_reward, norm_reward,
                                                     out_reward_dev, n);
  RCK(cudaGetLastError());
  return MYO_OK;
}

int myo_gae(const float* rewards_dev, const float* values_dev, const uint8_t* episode_starts_dev, const float* last_values_dev,
            const uint8_t* last_dones_dev, int n_steps, int n, float gamma, float gae_lambda, float* advantages_dev, float* returns_dev,
            void* stream) {
  if (!rewards_dev || !values_dev || !episode_starts_dev || !last_values_dev || !last_dones_dev || !advantages_dev || !returns_dev ||
      n_steps <= 0 || n <= 0) {
    myo::set_error("bad argument to myo_gae");
    return MYO_E_ARG;
  }
  gae_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(rewards_dev, values_dev, episode_starts_dev, last_values_dev,
                                                                              last_dones_dev, n_steps, n, gamma, gae_lambda, advantages_dev,
                                                                              returns_dev);
  RCK(cudaGetLastError());
  return MYO_OK;
}

int myo_normalize_obs(const float* obs_dev, const float* mean_f_dev, const float* var_f_dev, float epsilon, float clip_obs, float* out_dev,
                      int n, int d, void* stream) {
  if (!obs_dev || !mean_f_dev || !var_f_dev || !out_dev || n <= 0 || d <= 0) { myo::set_error("bad argument to myo_normalize_obs"); return MYO_E_ARG; }
  const int64_t total = (int64_t)n * d;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  normalize_obs_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(obs_dev, mean_f_dev, var_f_dev, epsilon, clip_obs, out_dev, total, d);
  RCK(cudaGetLastError());
  return MYO_OK;
}

}  // extern "C"
