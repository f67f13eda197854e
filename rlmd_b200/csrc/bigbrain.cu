// K3: state-dependent leverage ("big brain") sweeps over a stop-loss x retention grid.
//
// Reference: lev/lev_exp.py coin_optimal_lev :240-267, coin_big_brain_lev :270-452,
// dice_optimal_lev :704-738, dice_big_brain_lev :741-932.  Per grid point
// (retention phi, stop-loss lambda), V_min = lambda * V0 and
//     lev  = eta * (1 - V_min / V)                                   (phi == 0)
//     lev  = eta * (1 - (V <= V0 ? V_min : V0 + phi (V - V0)) / V)   (phi != 0)
//     V   <- V * (1 + lev * g_t)
// with per-step statistics of the leverage BEFORE the update and of the wealth
// AFTER it (data[R,S,26,H-1]).
//
// Dtypes follow the reference's promotion rules (pinned by tests/golden/bigbrain_*):
//   coin: returns, wealth and leverage fp32; the float64 LEV_FACTOR is rounded to fp32;
//   dice: the outcomes are cast to float64 (:791), so returns and the wealth chain
//         are float64; phi == 0: leverage float64 with the float64 LEV_FACTOR;
//         phi != 0: leverage fp32, from the fp32-ROUNDED wealth (:731), promoted in
//         the update;
//   step 0 uses a scalar leverage computed on 0-dim operands (float64; host side).
// Compiled with -fmad=false and IEEE division, so the chains are the reference's
// operation for operation.
//
// One thread per investor and PT grid points (wealth and leverage in registers);
// the wealth / leverage after every step are streamed to a [point, step, 2, N]
// fp32 dump (coalesced over investors) for b200_rowstats - order statistics
// commute with the monotone fp64 -> fp32 rounding, so medians stay exact.
#include "common.cuh"

namespace b200 {

constexpr int BB_PT = 4;  // grid points per thread

struct BBPoints {
  float vmin[B200_BB_MAX_POINTS];   // fl32(stop * V0)
  float roll[B200_BB_MAX_POINTS];   // retention ratio (fp32 grid value)
  double lev0[B200_BB_MAX_POINTS];  // leverage of step 0 (float64, 0-dim arithmetic)
};

template <typename T>
struct BBTraits;
template <>
struct BBTraits<float> {
  static __device__ __forceinline__ float opt(float v, float v0, float vmin, float roll, float eta32, double) {
    if (roll == 0.0f) return eta32 * (1.0f - vmin / v);
    const float floor_ = v <= v0 ? vmin : v0 + roll * (v - v0);
    return eta32 * (1.0f - floor_ / v);
  }
};
template <>
struct BBTraits<double> {
  static __device__ __forceinline__ double opt(double v, float v0, float vmin, float roll, float eta32, double eta64) {
    if (roll == 0.0f) return eta64 * (1.0 - (double)vmin / v);
    const float vf = (float)v;  // T.tensor(value_t, dtype=T.float) (:731)
    const float floor_ = vf <= v0 ? vmin : v0 + roll * (vf - v0);
    return (double)(eta32 * (1.0f - floor_ / vf));
  }
};

template <typename T>
__global__ void __launch_bounds__(128)
bigbrain_chunk_kernel(const __grid_constant__ b200_bigbrain_desc d, const __grid_constant__ BBPoints pts,
                      const uint8_t* __restrict__ outcomes, int32_t s_begin, int32_t s_end, T* __restrict__ state,
                      float* __restrict__ dump) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n_investors) return;
  const int p0 = blockIdx.y * BB_PT;
  const int N_P = d.n_points;
  const int64_t N = d.n_investors;
  const float v0 = d.value_0, eta32 = d.lev_factor32;
  const double eta64 = d.lev_factor64;
  T ret[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) ret[k] = (T)d.returns[k];  // coin: fl32(return); dice: the Python double
  const uint8_t* __restrict__ row = outcomes + i * d.ld_outcomes;

  T V[BB_PT], L[BB_PT];
  float vmin[BB_PT], roll[BB_PT];
#pragma unroll
  for (int q = 0; q < BB_PT; ++q) {
    const int p = min(p0 + q, N_P - 1);
    vmin[q] = pts.vmin[p];
    roll[q] = pts.roll[p];
  }
  auto pick = [&](unsigned code) -> T { return code == 0 ? ret[0] : (code == 1 ? ret[1] : ret[2]); };
  int s = s_begin;
  if (s == 0) {
    const T g = pick(row[0]);
#pragma unroll
    for (int q = 0; q < BB_PT; ++q) {
      const int p = min(p0 + q, N_P - 1);
      V[q] = (T)v0 * ((T)1 + (T)pts.lev0[p] * g);                // :330 / :810
      L[q] = BBTraits<T>::opt(V[q], v0, vmin[q], roll[q], eta32, eta64);
    }
    s = 1;
  } else {
#pragma unroll
    for (int q = 0; q < BB_PT; ++q) {
      const int p = min(p0 + q, N_P - 1);
      V[q] = state[(int64_t)p * N + i];
      L[q] = state[((int64_t)N_P + p) * N + i];
    }
  }
  const int s_lo = max(s_begin, 1);
  const int64_t tc = s_end - s_lo;
  auto step = [&](unsigned code, int at) {
    const T g = pick(code);
#pragma unroll
    for (int q = 0; q < BB_PT; ++q) {
      const int p = p0 + q;
      float* out = dump + (((int64_t)min(p, N_P - 1) * tc + (at - s_lo)) * 2) * N + i;
      if (p < N_P && dump != nullptr) __stcs(out, (float)L[q]);        // leverage before the step (:337)
      V[q] = V[q] * ((T)1 + L[q] * g);                                 // :374 / :855
      L[q] = BBTraits<T>::opt(V[q], v0, vmin[q], roll[q], eta32, eta64);
      if (p < N_P && dump != nullptr) __stcs(out + N, (float)V[q]);    // wealth after it (:381)
    }
  };
  while (s < s_end) {
    if (((uintptr_t)(row + s) & 15) == 0 && s + 16 <= s_end) {         // 16 outcomes per load
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(row + s));
      const unsigned words[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int b = 0; b < 16; ++b) step((words[b >> 2] >> (8 * (b & 3))) & 0xffu, s + b);
      s += 16;
    } else {
      step(__ldg(row + s), s);
      ++s;
    }
  }
#pragma unroll
  for (int q = 0; q < BB_PT; ++q) {
    const int p = p0 + q;
    if (p < N_P) {
      state[(int64_t)p * N + i] = V[q];
      state[((int64_t)N_P + p) * N + i] = L[q];
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_bigbrain_chunk(const b200_bigbrain_desc* d, const uint8_t* outcomes, const float* stop_vmin_host,
                                   const float* roll_host, const double* lev0_host, int32_t s_begin, int32_t s_end,
                                   void* state, float* dump, void* stream) {
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_REQUIRE(d != nullptr, "bigbrain: desc is NULL");
  B200_REQUIRE(d->n_investors >= 0 && d->horizon >= 1, "bigbrain: bad sizes");
  B200_REQUIRE(d->kind == B200_BB_COIN || d->kind == B200_BB_DICE, "bigbrain: kind must be B200_BB_COIN / _DICE");
  B200_REQUIRE(d->n_points >= 1, "bigbrain: n_points < 1");
  if (d->n_points > B200_BB_MAX_POINTS)
    return set_error(B200_ELIMIT, "bigbrain: n_points %d exceeds %d per call", d->n_points, B200_BB_MAX_POINTS);
  B200_REQUIRE(d->ld_outcomes >= d->horizon, "bigbrain: ld_outcomes < horizon");
  B200_REQUIRE(0 <= s_begin && s_begin < s_end && s_end <= d->horizon, "bigbrain: need 0 <= s_begin < s_end <= horizon");
  B200_REQUIRE(stop_vmin_host && roll_host && lev0_host, "bigbrain: a host table pointer is NULL");
  if (d->n_investors == 0) return 0;
  B200_REQUIRE(outcomes != nullptr && state != nullptr, "bigbrain: outcomes / state is NULL");
  BBPoints pts;
  for (int p = 0; p < B200_BB_MAX_POINTS; ++p) {
    const int q = p < d->n_points ? p : d->n_points - 1;
    pts.vmin[p] = stop_vmin_host[q];
    pts.roll[p] = roll_host[q];
    pts.lev0[p] = lev0_host[q];
  }
  dim3 grid((unsigned)((d->n_investors + 127) / 128), (unsigned)((d->n_points + BB_PT - 1) / BB_PT));
  cudaStream_t st = (cudaStream_t)stream;
  if (d->kind == B200_BB_COIN)
    bigbrain_chunk_kernel<float><<<grid, 128, 0, st>>>(*d, pts, outcomes, s_begin, s_end, (float*)state, dump);
  else
    bigbrain_chunk_kernel<double><<<grid, 128, 0, st>>>(*d, pts, outcomes, s_begin, s_end, (double*)state, dump);
  B200_CUDA(cudaGetLastError());
  return 0;
}
