// Batched step / reset of the market environments.
//
// Reference: envs/market_envs.py - Market_Inv{A,B,C}_D1 :133-202, :283-358,
// :440-528 and Market_Inv{A,B,C}_Dx :611-682, :765-841, :924-1013 (resets :204-222,
// :684-703) with the done flags of tools/env_resources.py:140-200 (market_dones).
//
// The wealth update is the multiplicative envs' (menv_core.cuh) fed by historical
// prices instead of sampled returns: with `assets` the (flattened, obs_days x
// n_assets) prices handed to reset() - the reference never advances them, so
// every return is relative to the episode's first observation (:150,:628) -
//     hist = next_assets / assets - 1,  r = hist[:n_assets],
//     step_return = clip(sum(lev * r), MIN_RETURN, MAX_RETURN),  lev = a * 3.
// One thread per environment, fp64 in the reference's order.  np.sum / np.mean
// over a contiguous fp64 vector are NumPy's pairwise sums (8 running partial
// sums once n >= 8); np_sum() below reproduces that order for n <= 128 so the
// exact-equality termination tests fire where the reference's do.
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ double clampm(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

__global__ void __launch_bounds__(128)
market_reset_kernel(const __grid_constant__ b200_market_desc d, int64_t E, double* __restrict__ wealth,
                    int32_t* __restrict__ time, const double* __restrict__ assets_in, double* __restrict__ assets,
                    double* __restrict__ state, const uint8_t* __restrict__ mask) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  if (mask != nullptr && !mask[e]) return;
  const int W = d.obs_days * d.n_assets;
  wealth[e] = d.initial_value;
  time[e] = 1;
  for (int k = 0; k < W; ++k) assets[e * W + k] = assets_in[e * W + k];
  if (state != nullptr) {
    double* s = state + e * (4 + W);
    s[0] = d.initial_value / d.max_value;
    s[1] = 0.0 / d.max_value;
    s[2] = 1.0 / d.max_value;
    s[3] = 1.0 / d.max_value;
    // D1 shows zeros (:219), Dx the observed prices (:700)
    for (int k = 0; k < W; ++k) s[4 + k] = (d.obs_days == 1 ? 0.0 : assets_in[e * W + k]) / d.max_value;
  }
}

__global__ void __launch_bounds__(128)
market_step_kernel(const __grid_constant__ b200_market_desc d, int64_t E, double* __restrict__ wealth,
                   int32_t* __restrict__ time, const double* __restrict__ assets,
                   const double* __restrict__ action, const double* __restrict__ next_assets,
                   double* __restrict__ next_state, double* __restrict__ reward_out, uint8_t* __restrict__ done_out,
                   double* __restrict__ risk, int A, int R) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int n = d.n_assets, W = d.obs_days * n, S = 4 + W;
  const int off = d.investor;  // A 0, B 1, C 2
  const double* a = action + e * A;
  const double* p0 = assets + e * W;
  const double* p1 = next_assets + e * W;
  const bool has_stop = off >= 1, has_ret = off >= 2;
  const double stop = has_stop ? (a[0] + d.max_abs_action) / 2 : 0.0;
  const double retention = has_ret ? (a[1] + d.max_abs_action) / 2 : 0.0;

  const double total = np_sum(n, [&](int i) { return (a[off + i] * d.lev_factor) * (p1[i] / p0[i] - 1); });
  const double step_return = clampm(total, d.min_return, d.max_return);

  const double w0 = wealth[e];
  double wmin, w, active = 1.0;
  if (!has_stop) {
    wmin = d.min_value;
    w = clampm(w0 * (1 + step_return), d.min_value, d.max_value);
  } else {
    const double floor_b = fmax(d.initial_value * stop, d.min_value);
    if (!has_ret || w0 <= d.initial_value) wmin = floor_b;
    else wmin = d.initial_value + (w0 - d.initial_value) * retention;
    active = fmax(w0 - wmin, 0.0);
    w = clampm(wmin + active * (1 + step_return), wmin, d.max_value);
  }
  const int t = time[e];
  const double growth = w / d.initial_value;
  const double rew = exp(log(growth) / (double)t);

  double* ns = next_state + e * S;
  const double s0 = w / d.max_value, s1 = step_return / d.max_value, s2 = growth / d.max_value,
               s3 = rew / d.max_value;
  ns[0] = s0; ns[1] = s1; ns[2] = s2; ns[3] = s3;
  bool done_state = (s0 >= 1.0) || (s1 >= 1.0) || (s2 >= 1.0) || (s3 >= 1.0);
  for (int k = 0; k < W; ++k) {
    const double q = (p1[k] / p0[k] - 1) / d.max_value;
    ns[4 + k] = q;
    done_state = done_state || q >= 1.0;
  }
  const double lev_cap = d.max_abs_action * d.lev_factor;
  bool all_max = true, all_min = true;
  for (int i = 0; i < n; ++i) {
    const double al = fabs(a[off + i] * d.lev_factor);
    all_max = all_max && (al == lev_cap);
    all_min = all_min && (al < d.min_weight);
  }
  const bool done_time = t == d.time_length;
  bool done = done_time || (w == wmin) || (rew < d.min_reward) || (step_return == d.min_return) || all_max ||
              all_min || done_state;
  if (has_stop) done = done || (active == 0.0);
  done_out[e * 2 + 0] = done ? 1 : 0;
  done_out[e * 2 + 1] = (done && !(done_time || done_state)) ? 1 : 0;
  reward_out[e] = rew;

  double* rk = risk + e * R;
  rk[0] = rew; rk[1] = w; rk[2] = step_return;
  rk[3] = np_sum(n, [&](int i) { return a[off + i] * d.lev_factor; }) / (double)n;
  int c = 4;
  if (has_stop) rk[c++] = stop;
  if (has_ret) rk[c++] = retention;
  if (n > 1)
    for (int i = 0; i < n; ++i) rk[c++] = a[off + i] * d.lev_factor;

  wealth[e] = w;
  time[e] = t + 1;
}

static int market_dims(const b200_market_desc* d, int* S, int* A, int* R) {
  B200_REQUIRE(d != nullptr, "market: desc is NULL");
  B200_REQUIRE(d->investor >= B200_INV_A && d->investor <= B200_INV_C, "market: unknown investor %d", d->investor);
  B200_REQUIRE(d->n_assets >= 1 && d->n_assets <= B200_MARKET_MAX_ASSETS, "market: n_assets must be in 1..%d",
               B200_MARKET_MAX_ASSETS);
  B200_REQUIRE(d->obs_days >= 1 && d->obs_days <= 4096, "market: obs_days must be in 1..4096");
  const int n = d->n_assets;
  *S = 4 + d->obs_days * n;
  *A = d->investor + n;
  *R = (n == 1 ? 3 : 4) + d->investor + n;
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_market_dims(const b200_market_desc* desc, int32_t* state_dim, int32_t* action_dim,
                                int32_t* risk_dim) {
  int S, A, R;
  if (int rc = market_dims(desc, &S, &A, &R)) return rc;
  if (state_dim) *state_dim = S;
  if (action_dim) *action_dim = A;
  if (risk_dim) *risk_dim = R;
  return 0;
}

extern "C" int b200_market_reset(const b200_market_desc* desc, int64_t n_envs, double* wealth, int32_t* time,
                                 const double* assets_in, double* assets, double* state, const uint8_t* mask,
                                 void* stream) {
  int S, A, R;
  if (int rc = market_dims(desc, &S, &A, &R)) return rc;
  B200_REQUIRE(n_envs >= 0, "market_reset: negative n_envs");
  if (n_envs == 0) return 0;
  B200_REQUIRE(wealth && time && assets_in && assets, "market_reset: NULL buffer");
  const unsigned blocks = (unsigned)((n_envs + 127) / 128);
  market_reset_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(*desc, n_envs, wealth, time, assets_in, assets, state,
                                                                mask);
  return check_cuda(cudaGetLastError(), "market_reset launch");
}

extern "C" int b200_market_step(const b200_market_desc* desc, int64_t n_envs, double* wealth, int32_t* time,
                                const double* assets, const double* action, const double* next_assets,
                                double* next_state, double* reward, uint8_t* done, double* risk, void* stream) {
  int S, A, R;
  if (int rc = market_dims(desc, &S, &A, &R)) return rc;
  B200_REQUIRE(n_envs >= 0, "market_step: negative n_envs");
  if (n_envs == 0) return 0;
  B200_REQUIRE(wealth && time && assets && action && next_assets && next_state && reward && done && risk,
               "market_step: NULL buffer");
  const unsigned blocks = (unsigned)((n_envs + 127) / 128);
  market_step_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(*desc, n_envs, wealth, time, assets, action,
                                                               next_assets, next_state, reward, done, risk, A, R);
  return check_cuda(cudaGetLastError(), "market_step launch");
}
