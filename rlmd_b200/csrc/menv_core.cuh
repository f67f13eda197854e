// Device core of the multiplicative env step, shared by the step kernel
// (menv.cu), the fused collector and the evaluation rollouts (collect.cu).
//
// Reference: envs/coin_flip_envs.py:150-233,290-379,436-538, envs/dice_roll_envs.py
// :174-176,:311, envs/gbm_envs.py:147-212,304-323,449-474,
// envs/dice_roll_sh_envs.py:160-235,290-365,420-502,557-645 and the done flags
// of tools/env_resources.py:26-137.  All arithmetic is fp64 in the reference's
// order (the library is compiled with -fmad=false).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int ENV_MAX_GAMBLES = B200_ENV_MAX_GAMBLES;
constexpr int ENV_MAX_STATE = 4 + ENV_MAX_GAMBLES;
constexpr int ENV_MAX_RISK = 6 + ENV_MAX_GAMBLES;

__device__ __forceinline__ double clampd(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

// Returns of one step of env `e`: counter = (env lo, env hi, draw_index lo,
// TAG_ENV + block), key = seed ^ draw_index hi.
template <int GS>
__device__ __forceinline__ void env_draw_returns(const b200_env_desc& d, int64_t e, uint64_t draw_index, int n,
                                                 double (&r)[GS]) {
  const bool gbm = d.family == B200_ENV_GBM;
  const uint32_t k0 = (uint32_t)d.seed ^ (uint32_t)(draw_index >> 32), k1 = (uint32_t)(d.seed >> 32);
  uint32_t thr0, thr1;
  {
    const double c0 = d.probs[0] * 4294967296.0, c1 = (d.probs[0] + d.probs[1]) * 4294967296.0;
    thr0 = c0 >= 4294967295.0 ? 0xffffffffu : (uint32_t)floor(c0);
    thr1 = c1 >= 4294967295.0 ? 0xffffffffu : (uint32_t)floor(c1);
  }
#pragma unroll
  for (int j = 0; j * 4 < GS; ++j) {
    if (j * 4 >= n) break;
    const Philox4 p = philox4x32_10((uint32_t)e, (uint32_t)((uint64_t)e >> 32), (uint32_t)draw_index,
                                    PHILOX_TAG_ENV + (uint32_t)j, k0, k1);
    const uint32_t u[4] = {p.x, p.y, p.z, p.w};
    if (gbm) {
      float z[4];
      box_muller(u[0], u[1], z[0], z[1]);
      box_muller(u[2], u[3], z[2], z[3]);
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (j * 4 + b < GS && j * 4 + b < n) r[j * 4 + b] = d.log_mean + d.vol * (double)z[b];
    } else {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (j * 4 + b < GS && j * 4 + b < n) {
          int code = (u[b] >= thr0);
          if (d.family != B200_ENV_COIN) code += (u[b] >= thr1);
          r[j * 4 + b] = code == 0 ? d.returns[0] : code == 1 ? d.returns[1] : d.returns[2];
        }
      }
    }
  }
}

// NG > 0: the number of gambles is the compile-time constant NG (every loop
// unrolls, the per-gamble arrays live in registers); NG == 0: read from the desc.
// The dice_sh family has one gamble but a 6-word state and a 7-word risk vector.
template <int NG> struct EnvDims {
  static constexpr int G = NG > 0 ? NG : ENV_MAX_GAMBLES;          // gamble slots
  static constexpr int S = NG == 1 ? 6 : 4 + G;                    // state slots
  static constexpr int A = NG == 1 ? 4 : 2 + G;                    // action slots
  static constexpr int R = NG == 1 ? 7 : 6 + G;                    // risk slots
};

template <int NG = 0>
struct EnvStep {
  double w;        // wealth after the step
  double reward;
  bool done, learn_done;
  double ns[EnvDims<NG>::S];   // next state (normalised), S entries
  double rk[EnvDims<NG>::R];   // risk vector, R entries
};

// dst[i] = src[i] for i < count, fully unrolled over MAXC slots (register arrays stay registers)
template <int MAXC, typename D, typename S>
__device__ __forceinline__ void copy_n(D* __restrict__ dst, const S* __restrict__ src, int count) {
#pragma unroll
  for (int i = 0; i < MAXC; ++i)
    if (i < count) dst[i] = (D)src[i];
}

// One step from wealth w0 at env time t with action a[A] and returns r[n].
template <int NG>
__device__ __forceinline__ void env_step_core(const b200_env_desc& d, const double* __restrict__ a,
                                              const double (&r)[EnvDims<NG>::G], double w0, int t, EnvStep<NG>& o) {
  constexpr int GS = EnvDims<NG>::G;
  const bool sh = NG <= 1 && d.family == B200_ENV_DICE_SH;
  const bool gbm = d.family == B200_ENV_GBM;
  const int n = NG > 0 ? NG : (sh ? 1 : d.n_gambles);
  // ---- actions -> leverages, stop-loss, retention
  const int off = sh ? (d.investor == B200_INV_INSURED ? 0 : d.investor) : d.investor;  // A 0, B 1, C 2
  // a[off + i] with compile-time indices (an action held in registers stays there)
  auto act = [&](int i) { return off == 0 ? a[i] : off == 1 ? a[i + 1] : a[i + 2]; };
  double lev[GS];
  double lev_sh = 0.0, r_sh = 0.0;
  double stop = 0.0, retention = 0.0;
  const bool has_stop = d.investor == B200_INV_B || d.investor == B200_INV_C;
  const bool has_ret = d.investor == B200_INV_C;
  if (has_stop) stop = d.stop_abs ? fabs(a[0]) : (a[0] + d.max_abs_action) / 2;
  if (has_ret) retention = (a[1] + d.max_abs_action) / 2;

  double step_return, factor;
  if (sh) {
    r_sh = (r[0] == d.returns[2]) ? d.sh_returns[2] : (r[0] == d.returns[0]) ? d.sh_returns[0] : d.sh_returns[1];
    if (d.investor == B200_INV_INSURED) {
      lev[0] = a[0] * d.i_lev_factor;
      lev_sh = 1 - lev[0];
    } else {
      lev[0] = act(0) * d.lev_factor;
      lev_sh = (act(1) + d.max_abs_action) / 2 * d.sh_lev_factor;
    }
    step_return = clampd(lev[0] * r[0] + lev_sh * r_sh, d.min_return, d.max_return);
    factor = 1 + step_return;
  } else {
#pragma unroll
    for (int i = 0; i < GS; ++i)
      if (i < n) lev[i] = act(i) * d.lev_factor;
    const double total = np_sum<NG>(n, [&](int i) { return lev[i] * r[i]; });   // np.sum(lev * r)
    if (gbm) {
      step_return = fmax(total, d.min_return);
      factor = fmin(exp(step_return), 1 + d.max_return);
    } else {
      step_return = clampd(total, d.min_return, d.max_return);
      factor = 1 + step_return;
    }
  }

  // ---- wealth update
  double wmin, w, active = 1.0;
  if (!has_stop) {
    wmin = d.min_value;
    w = clampd(w0 * factor, d.min_value, d.max_value);
  } else {
    const double floor_b = fmax(d.initial_value * stop, d.min_value);
    if (!has_ret || w0 <= d.initial_value) wmin = floor_b;
    else wmin = d.initial_value + (w0 - d.initial_value) * retention;
    active = fmax(w0 - wmin, 0.0);
    w = clampd(wmin + active * factor, wmin, d.max_value);
  }
  const double growth = w / d.initial_value;
  const double rew = exp(log(growth) / (double)t);

  // ---- next state (normalised), done flags
  double* ns = o.ns;
  const double s0 = w / d.max_value, s1 = step_return / d.max_value, s2 = growth / d.max_value,
               s3 = rew / d.max_value;
  ns[0] = s0; ns[1] = s1; ns[2] = s2; ns[3] = s3;
  bool done_state = (s0 >= 1.0) || (s1 >= 1.0) || (s2 >= 1.0) || (s3 >= 1.0);
  if (sh) {
    const double q0 = r[0] / d.max_value, q1 = r_sh / d.max_value;
    ns[4] = q0; ns[5] = q1;
    done_state = done_state || q0 >= 1.0 || q1 >= 1.0;
  } else {
#pragma unroll
    for (int i = 0; i < GS; ++i) {
      if (i < n) {
        const double q = r[i] / d.max_value;
        ns[4 + i] = q;
        done_state = done_state || q >= 1.0;
      }
    }
  }
  const double lev_cap = d.max_abs_action * d.lev_factor;
  bool any_max = false, all_max = true, all_min = true;
#pragma unroll
  for (int i = 0; i < GS; ++i) {
    if (i < n) {
      const double al = fabs(lev[i]);
      any_max = any_max || (al == lev_cap);
      all_max = all_max && (al == lev_cap);
      all_min = all_min && (al < d.min_weight);
    }
  }
  bool done = (w == wmin) || (rew < d.min_reward) || (step_return == d.min_return) || (gbm ? all_max : any_max) ||
              all_min || done_state;
  if (has_stop) done = done || (active == 0.0);
  o.done = done;
  o.learn_done = done && !done_state;
  o.reward = rew;

  // ---- risk vector
  double* rk = o.rk;
  rk[0] = rew; rk[1] = w; rk[2] = step_return;
  if (sh) {
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    rk[3] = lev[0];
    rk[4] = has_stop ? stop : qnan;
    rk[5] = has_ret ? retention : qnan;
    rk[6] = lev_sh;
  } else {
    rk[3] = np_sum<NG>(n, [&](int i) { return lev[i]; }) / (double)n;           // np.mean(lev)
    // stop / retention / per-gamble leverages follow without gaps; written with
    // compile-time indices (a running index would put rk[] in local memory)
    const int base = 4 + (has_stop ? 1 : 0) + (has_ret ? 1 : 0);
    if (has_stop) rk[4] = stop;
    if (has_ret) rk[5] = retention;
    if (n > 1) {
#pragma unroll
      for (int i = 0; i < GS; ++i) {
        if (i < n) {
          const double v = lev[i];
          if (base == 4) rk[4 + i] = v;
          else if (base == 5) rk[5 + i] = v;
          else rk[6 + i] = v;
        }
      }
    }
  }

  o.w = w;
}

// Kernel instantiation for a desc: its number of gambles when a specialisation
// exists (1, 2, 3; dice_sh is the one-gamble layout), else 0 = generic.
__host__ __device__ __forceinline__ int env_ng(const b200_env_desc& d) {
  if (d.family == B200_ENV_DICE_SH) return 1;
  return d.n_gambles <= 3 ? d.n_gambles : 0;
}
#define B200_ENV_DISPATCH(ng, CALL)                 \
  switch (ng) {                                     \
    case 1: { constexpr int NG = 1; CALL; } break;  \
    case 2: { constexpr int NG = 2; CALL; } break;  \
    case 3: { constexpr int NG = 3; CALL; } break;  \
    default: { constexpr int NG = 0; CALL; } break; \
  }

// State / risk vector lengths (env_dims on the host computes the same).
__device__ __forceinline__ void env_dims_dev(const b200_env_desc& d, int& S, int& A, int& R) {
  if (d.family == B200_ENV_DICE_SH) {
    S = 6; R = 7;
    A = d.investor == B200_INV_INSURED ? 1 : 2 + d.investor;
    return;
  }
  const int n = d.n_gambles;
  S = 4 + n;
  A = d.investor + n;
  R = (n == 1 ? 3 : 4) + d.investor + n;
}

}  // namespace b200

