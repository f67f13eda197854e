// Fused collector and evaluation rollouts.
//
// Reference: the inner loop of scripts/rl_multiplicative.py:185-273
//     next_state, reward, env_done, risk = env.step(action)
//     agent.store_transistion(state, action, reward, next_state, learn_done)
//     state = next_state            (env.reset() once the episode is done)
// and tools/eval_episodes.py:233-262 (n_eval episodes under one constant action
// each, until done or max_eval_steps).
//
// collect_step_kernel: one thread per environment does the env step (the shared
// core of menv_core.cuh), appends the transition to ITS lane of the replay
// memory with the reference's episode bookkeeping (replay.cu restates it; a lane
// is a reference buffer of lane_len transitions with its own 8-word header,
// local index j of lane e at slot j * n_envs + e so that one step's appends are
// contiguous rows), and
// carries the observation forward (reset state once the episode is done).  No
// host round trip between step, store and sample: the step counter that feeds
// the Philox draws lives on the device, so one collect-step + sample pair can be
// captured in a CUDA graph and replayed.
//
// rollout_kernel: one thread per evaluation episode, wealth in registers, the
// whole episode in one launch.
#include "menv_core.cuh"

namespace b200 {

enum { H_MEM_IDX = 0, H_EPISODES = 1, H_E0 = 2, H_ELAST = 3, H_RUN_START = 4 };
enum { CTR_STEP = 0, CTR_SAMPLE = 1 };

int replay_sample_lanes(const b200_replay_desc* d, int64_t lanes, int64_t lane_len, const int64_t* idx,
                        int64_t n_batches, int32_t batch, int64_t filled, int32_t multi_steps,
                        const float* gamma_pow_host, int32_t additive, uint64_t seed, uint64_t draw_index,
                        const int64_t* draw_counter, int64_t* out_idx, float* out_state, float* out_action,
                        float* out_reward, float* out_next_state, uint8_t* out_done, int64_t* out_eff,
                        void* stream, int64_t* bump, int* bumped);

__device__ __forceinline__ void reset_state(const b200_env_desc& d, double* __restrict__ s, int S) {
  s[0] = d.initial_value / d.max_value;
  s[1] = 0.0 / d.max_value;
  s[2] = 1.0 / d.max_value;
  s[3] = 1.0 / d.max_value;
  for (int i = 4; i < S; ++i) s[i] = 0.0;
}

__global__ void __launch_bounds__(128)
collect_reset_kernel(const __grid_constant__ b200_collect_desc c, double* __restrict__ wealth,
                     int32_t* __restrict__ time, double* __restrict__ cur_state, int64_t* __restrict__ counter) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e == 0) { counter[CTR_STEP] = 0; counter[CTR_SAMPLE] = 0; }
  if (e >= c.n_envs) return;
  int S, A, R;
  env_dims_dev(c.env, S, A, R);
  wealth[e] = c.env.initial_value;
  time[e] = 1;
  reset_state(c.env, cur_state + e * S, S);
  for (int i = 0; i < 8; ++i) c.replay.header[e * 8 + i] = 0;
}

template <int NG>
__global__ void __launch_bounds__(128)
collect_step_kernel(const __grid_constant__ b200_collect_desc c, double* __restrict__ wealth,
                    int32_t* __restrict__ time, double* __restrict__ cur_state, const double* __restrict__ action,
                    const double* __restrict__ r_in, const int64_t* __restrict__ counter,
                    double* __restrict__ reward_out, uint8_t* __restrict__ done_out, double* __restrict__ risk_out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= c.n_envs) return;
  const b200_env_desc& d = c.env;
  int S, A, R;
  env_dims_dev(d, S, A, R);
  using Dm = EnvDims<NG>;
  const int n = NG > 0 ? NG : d.n_gambles;
  double r[Dm::G];
  if (r_in != nullptr) {
    copy_n<Dm::G>(r, r_in + e * n, n);
  } else {
    env_draw_returns(d, e, (uint64_t)counter[CTR_STEP], n, r);
  }
  double a[Dm::A];
  copy_n<Dm::A>(a, action + e * A, A);
  double* st = cur_state + e * S;
  EnvStep<NG> o;
  const int t = time[e];
  env_step_core<NG>(d, a, r, wealth[e], t, o);

  // ---- append (state, action, reward, next_state, learn_done) to lane e
  const b200_replay_desc& m = c.replay;
  int64_t* h = m.header + e * 8;
  const int64_t pos = h[H_MEM_IDX];
  const int64_t local = pos % c.lane_len;
  const int64_t slot = local * c.n_envs + e;   // slot-major: the lanes' writes of one step coalesce
  copy_n<Dm::S>(m.state_memory + slot * S, st, S);
  copy_n<Dm::S>(m.next_state_memory + slot * S, o.ns, S);
  copy_n<Dm::A>(m.action_memory + slot * A, a, A);
  m.reward_memory[slot] = (float)(o.reward > c.reward_floor ? o.reward : c.reward_floor);
  m.terminal_memory[slot] = o.learn_done ? 1 : 0;
  m.episode_start[slot] = (int32_t)h[H_RUN_START];
  if (o.learn_done) {
    if (h[H_EPISODES] == 0) h[H_E0] = pos;
    h[H_ELAST] = pos;
    h[H_RUN_START] = pos + 1;
    h[H_EPISODES] += 1;
  }
  h[H_MEM_IDX] = pos + 1;

  // ---- outputs of the step, then carry the observation (auto-reset when done)
  if (reward_out != nullptr) reward_out[e] = o.reward;
  if (done_out != nullptr) { done_out[e * 2] = o.done ? 1 : 0; done_out[e * 2 + 1] = o.learn_done ? 1 : 0; }
  if (risk_out != nullptr) copy_n<Dm::R>(risk_out + e * R, o.rk, R);
  if (o.done) {
    wealth[e] = d.initial_value;
    time[e] = 1;
    reset_state(d, st, S);
  } else {
    wealth[e] = o.w;
    time[e] = t + 1;
    copy_n<Dm::S>(st, o.ns, S);
  }
}

__global__ void counter_bump_kernel(int64_t* __restrict__ counter, int which) { counter[which] += 1; }

// One evaluation episode per thread: constant action, until done or max_steps.
template <int NG>
__global__ void __launch_bounds__(128)
rollout_kernel(const __grid_constant__ b200_env_desc d, int64_t E, const double* __restrict__ action,
               const double* __restrict__ r_in, int64_t r_stride_t, uint64_t draw_base, int32_t max_steps,
               double* __restrict__ reward_out, int32_t* __restrict__ steps_out, double* __restrict__ risk_out,
               double* __restrict__ state_out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int S, A, R;
  env_dims_dev(d, S, A, R);
  using Dm = EnvDims<NG>;
  const int n = NG > 0 ? NG : d.n_gambles;
  double a[Dm::A];
  copy_n<Dm::A>(a, action + e * A, A);
  double w = d.initial_value;
  EnvStep<NG> o;
  o.reward = 0.0;
  o.done = false;
  int step = 0;
  while (step < max_steps) {
    double r[Dm::G];
    if (r_in != nullptr) {
      copy_n<Dm::G>(r, r_in + (int64_t)step * r_stride_t + e * n, n);
    } else {
      env_draw_returns(d, e, draw_base + (uint64_t)step, n, r);
    }
    env_step_core<NG>(d, a, r, w, step + 1, o);
    w = o.w;
    ++step;
    if (o.done) break;
  }
  reward_out[e] = o.reward;
  steps_out[e] = step;
  if (step > 0) {
    copy_n<Dm::R>(risk_out + e * R, o.rk, R);
    if (state_out != nullptr) copy_n<Dm::S>(state_out + e * S, o.ns, S);
  }
}

static int check_collect(const b200_collect_desc* c, int* S, int* A, int* R) {
  B200_REQUIRE(c != nullptr, "collect: desc is NULL");
  int32_t s, a, r;
  if (int rc = b200_menv_dims(&c->env, &s, &a, &r)) return rc;
  B200_REQUIRE(c->n_envs >= 1 && c->lane_len >= 1, "collect: n_envs and lane_len must be positive");
  B200_REQUIRE(c->n_envs * c->lane_len <= c->replay.mem_size, "collect: n_envs x lane_len exceeds the replay mem_size");
  B200_REQUIRE(c->lane_len < (1ll << 31), "collect: lane_len must be < 2^31");
  B200_REQUIRE(c->replay.state_dim == s && c->replay.action_dim == a,
               "collect: the replay buffer's state/action dims (%d,%d) differ from the env's (%d,%d)",
               c->replay.state_dim, c->replay.action_dim, s, a);
  B200_REQUIRE(c->replay.state_memory && c->replay.action_memory && c->replay.reward_memory &&
                   c->replay.next_state_memory && c->replay.terminal_memory && c->replay.episode_start &&
                   c->replay.header,
               "collect: a replay buffer pointer is NULL");
  *S = s; *A = a; *R = r;
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_collect_reset(const b200_collect_desc* c, double* wealth, int32_t* time, double* cur_state,
                                  int64_t* counter, void* stream) {
  int S, A, R;
  if (int rc = check_collect(c, &S, &A, &R)) return rc;
  B200_REQUIRE(wealth && time && cur_state && counter, "collect_reset: NULL buffer");
  cudaStream_t st = (cudaStream_t)stream;
  B200_CUDA(cudaMemsetAsync(c->replay.terminal_memory, 0, (size_t)c->replay.mem_size, st));
  B200_CUDA(cudaMemsetAsync(c->replay.episode_start, 0, (size_t)c->replay.mem_size * sizeof(int32_t), st));
  const unsigned blocks = (unsigned)((c->n_envs + 127) / 128);
  collect_reset_kernel<<<blocks, 128, 0, st>>>(*c, wealth, time, cur_state, counter);
  return check_cuda(cudaGetLastError(), "collect_reset launch");
}

extern "C" int b200_collect_step(const b200_collect_desc* c, double* wealth, int32_t* time, double* cur_state,
                                 const double* action, const double* returns_in, int64_t* counter, double* reward,
                                 uint8_t* done, double* risk, void* stream) {
  int S, A, R;
  if (int rc = check_collect(c, &S, &A, &R)) return rc;
  B200_REQUIRE(wealth && time && cur_state && action && counter, "collect_step: NULL buffer");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((c->n_envs + 127) / 128);
  B200_ENV_DISPATCH(env_ng(c->env), (collect_step_kernel<NG><<<blocks, 128, 0, st>>>(
                                        *c, wealth, time, cur_state, action, returns_in, counter, reward, done, risk)));
  counter_bump_kernel<<<1, 1, 0, st>>>(counter, CTR_STEP);
  return check_cuda(cudaGetLastError(), "collect_step launch");
}

extern "C" int b200_collect_sample(const b200_collect_desc* c, const int64_t* idx, int64_t n_batches, int32_t batch,
                                   int32_t multi_steps, const float* gamma_pow_host, int32_t additive, uint64_t seed,
                                   int64_t* counter, int64_t* out_idx, float* out_state, float* out_action,
                                   float* out_reward, float* out_next_state, uint8_t* out_done, int64_t* out_eff,
                                   void* stream) {
  int S, A, R;
  if (int rc = check_collect(c, &S, &A, &R)) return rc;
  B200_REQUIRE(idx != nullptr || counter != nullptr, "collect_sample: counter is NULL while indices are drawn");
  int rc = replay_sample_lanes(&c->replay, c->n_envs, c->lane_len, idx, n_batches, batch, -1, multi_steps,
                               gamma_pow_host, additive, seed, 0, idx ? nullptr : counter + CTR_SAMPLE, out_idx,
                               out_state, out_action, out_reward, out_next_state, out_done, out_eff, stream, nullptr,
                               nullptr);
  if (rc) return rc;
  if (idx == nullptr && n_batches * batch > 0) {
    counter_bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, CTR_SAMPLE);
    return check_cuda(cudaGetLastError(), "collect_sample counter");
  }
  return 0;
}

extern "C" int b200_menv_rollout(const b200_env_desc* desc, int64_t n_episodes, const double* action,
                                 const double* returns_in, uint64_t draw_base, int32_t max_steps, double* reward,
                                 int32_t* steps, double* risk, double* last_state, void* stream) {
  int32_t S, A, R;
  if (int rc = b200_menv_dims(desc, &S, &A, &R)) return rc;
  B200_REQUIRE(n_episodes >= 0 && max_steps >= 0, "menv_rollout: negative size");
  if (n_episodes == 0) return 0;
  B200_REQUIRE(action && reward && steps && risk, "menv_rollout: NULL buffer");
  const int n = desc->family == B200_ENV_DICE_SH ? 1 : desc->n_gambles;
  const unsigned blocks = (unsigned)((n_episodes + 127) / 128);
  B200_ENV_DISPATCH(env_ng(*desc), (rollout_kernel<NG><<<blocks, 128, 0, (cudaStream_t)stream>>>(
                                       *desc, n_episodes, action, returns_in, n_episodes * n, draw_base, max_steps,
                                       reward, steps, risk, last_state)));
  return check_cuda(cudaGetLastError(), "menv_rollout launch");
}
