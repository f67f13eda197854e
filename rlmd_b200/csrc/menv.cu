// K4: batched step / reset of the multiplicative gym environments.
//
// Reference: envs/coin_flip_envs.py:150-233,290-379,436-538 (Coin_Inv{A,B,C}),
// envs/dice_roll_envs.py (same, trinary sampling :174-176, stop-loss map :311),
// envs/gbm_envs.py:147-212,304-323,449-474, envs/dice_roll_sh_envs.py:160-235,
// 290-365,420-502,557-645, and the done flags of tools/env_resources.py:26-137.
//
// One thread per environment, all arithmetic in fp64 in the reference's order
// (-fmad=false: no contraction), so the exact-equality termination tests
// (wealth == floor, return == MIN_RETURN, |lev| == eps1*eta) fire exactly where
// the reference's do; only exp/log in the reward differ in the last ulp.
// Returns are injected (parity runs) or drawn with Philox4x32-10:
//   counter = (env lo, env hi, draw_index lo, TAG_ENV + block), key = seed ^ draw_index hi.
#include <algorithm>

#include "menv_core.cuh"

namespace b200 {

__global__ void __launch_bounds__(128)
menv_reset_kernel(const __grid_constant__ b200_env_desc d, int64_t E, double* __restrict__ wealth,
                  int32_t* __restrict__ time, double* __restrict__ state, const uint8_t* __restrict__ mask,
                  int S) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  if (mask != nullptr && !mask[e]) return;
  wealth[e] = d.initial_value;
  time[e] = 1;
  if (state != nullptr) {
    double* s = state + e * S;
    s[0] = d.initial_value / d.max_value;
    s[1] = 0.0 / d.max_value;
    s[2] = 1.0 / d.max_value;
    s[3] = 1.0 / d.max_value;
    for (int i = 4; i < S; ++i) s[i] = 0.0;
  }
}

// The [E, S] / [E, A] / [E, R] rows of a block's 128 environments are one
// contiguous span each: they move between global and shared memory with
// unit-stride accesses, a thread touches its own row in shared memory only.
constexpr int MENV_THREADS = 128;

__device__ __forceinline__ void span_to_smem(double* __restrict__ sm, const double* __restrict__ g, int64_t words) {
  for (int64_t i = threadIdx.x; i < words; i += MENV_THREADS) sm[i] = __ldcs(g + i);
}
__device__ __forceinline__ void smem_to_span(double* __restrict__ g, const double* __restrict__ sm, int64_t words) {
  for (int64_t i = threadIdx.x; i < words; i += MENV_THREADS) __stcs(g + i, sm[i]);
}

template <int NG>
__global__ void __launch_bounds__(MENV_THREADS)
menv_step_kernel(const __grid_constant__ b200_env_desc d, int64_t E, double* __restrict__ wealth,
                 int32_t* __restrict__ time, const double* __restrict__ action, const double* __restrict__ r_in,
                 uint64_t draw_index, double* __restrict__ next_state, double* __restrict__ reward_out,
                 uint8_t* __restrict__ done_out, double* __restrict__ risk, int S, int A, int R) {
  using Dm = EnvDims<NG>;
  extern __shared__ double stage[];   // MENV_THREADS * max(S, A, R) doubles
  const int64_t e0 = (int64_t)blockIdx.x * MENV_THREADS;
  const int64_t e = e0 + threadIdx.x;
  const int64_t live = min((int64_t)MENV_THREADS, E - e0);   // environments of this block
  const bool on = e < E;
  const int n = NG > 0 ? NG : d.n_gambles;

  // ---- actions (and injected returns) through shared memory
  double a[Dm::A];
  span_to_smem(stage, action + e0 * A, live * A);
  __syncthreads();
  if (on) copy_n<Dm::A>(a, stage + threadIdx.x * A, A);
  __syncthreads();
  double r[Dm::G];
  if (r_in != nullptr) {
    span_to_smem(stage, r_in + e0 * n, live * n);
    __syncthreads();
    if (on) copy_n<Dm::G>(r, stage + threadIdx.x * n, n);
    __syncthreads();
  } else if (on) {
    env_draw_returns(d, e, draw_index, n, r);
  }

  EnvStep<NG> o;
  if (on) {
    const int t = time[e];
    env_step_core<NG>(d, a, r, wealth[e], t, o);
    done_out[e * 2 + 0] = o.done ? 1 : 0;
    done_out[e * 2 + 1] = o.learn_done ? 1 : 0;
    reward_out[e] = o.reward;
    wealth[e] = o.w;
    time[e] = t + 1;
    copy_n<Dm::S>(stage + threadIdx.x * S, o.ns, S);
  }
  __syncthreads();
  smem_to_span(next_state + e0 * S, stage, live * S);
  __syncthreads();
  if (on) copy_n<Dm::R>(stage + threadIdx.x * R, o.rk, R);
  __syncthreads();
  smem_to_span(risk + e0 * R, stage, live * R);
}

static int env_dims(const b200_env_desc* d, int* S, int* A, int* R) {
  B200_REQUIRE(d != nullptr, "menv: desc is NULL");
  B200_REQUIRE(d->family >= B200_ENV_COIN && d->family <= B200_ENV_DICE_SH, "menv: unknown family %d", d->family);
  B200_REQUIRE(d->investor >= B200_INV_A && d->investor <= B200_INV_INSURED, "menv: unknown investor %d", d->investor);
  if (d->family == B200_ENV_DICE_SH) {
    *S = 6; *R = 7;
    *A = d->investor == B200_INV_INSURED ? 1 : 2 + d->investor;
    return 0;
  }
  B200_REQUIRE(d->investor != B200_INV_INSURED, "menv: the INSURED investor exists for dice_sh only");
  B200_REQUIRE(d->n_gambles >= 1 && d->n_gambles <= ENV_MAX_GAMBLES, "menv: n_gambles must be in 1..%d", ENV_MAX_GAMBLES);
  const int n = d->n_gambles;
  *S = 4 + n;
  *A = d->investor + n;
  *R = (n == 1 ? 3 : 4) + d->investor + n;
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_menv_dims(const b200_env_desc* desc, int32_t* state_dim, int32_t* action_dim, int32_t* risk_dim) {
  int S, A, R;
  int rc = env_dims(desc, &S, &A, &R);
  if (rc) return rc;
  if (state_dim) *state_dim = S;
  if (action_dim) *action_dim = A;
  if (risk_dim) *risk_dim = R;
  return 0;
}

extern "C" int b200_menv_reset(const b200_env_desc* desc, int64_t n_envs, double* wealth, int32_t* time, double* state,
                               const uint8_t* mask, void* stream) {
  int S, A, R;
  int rc = env_dims(desc, &S, &A, &R);
  if (rc) return rc;
  B200_REQUIRE(n_envs >= 0, "menv_reset: negative n_envs");
  if (n_envs == 0) return 0;
  B200_REQUIRE(wealth != nullptr && time != nullptr, "menv_reset: wealth/time is NULL");
  const unsigned blocks = (unsigned)((n_envs + 127) / 128);
  menv_reset_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(*desc, n_envs, wealth, time, state, mask, S);
  return check_cuda(cudaGetLastError(), "menv_reset launch");
}

extern "C" int b200_menv_step(const b200_env_desc* desc, int64_t n_envs, double* wealth, int32_t* time,
                              const double* action, const double* returns_in, uint64_t draw_index, double* next_state,
                              double* reward, uint8_t* done, double* risk, void* stream) {
  int S, A, R;
  int rc = env_dims(desc, &S, &A, &R);
  if (rc) return rc;
  B200_REQUIRE(n_envs >= 0, "menv_step: negative n_envs");
  if (n_envs == 0) return 0;
  B200_REQUIRE(wealth && time && action && next_state && reward && done && risk, "menv_step: NULL buffer");
  const unsigned blocks = (unsigned)((n_envs + MENV_THREADS - 1) / MENV_THREADS);
  const size_t smem = (size_t)MENV_THREADS * std::max(S, std::max(A, R)) * sizeof(double);
  B200_ENV_DISPATCH(env_ng(*desc), (menv_step_kernel<NG><<<blocks, MENV_THREADS, smem, (cudaStream_t)stream>>>(
                                       *desc, n_envs, wealth, time, action, returns_in, draw_index, next_state, reward,
                                       done, risk, S, A, R)));
  return check_cuda(cudaGetLastError(), "menv_step launch");
}
