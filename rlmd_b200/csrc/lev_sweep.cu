// K1: leverage sweep, final-time wealth for the whole grid in one launch.
//
// Reference: lev/lev_exp.py - the per-leverage loop of *_fixed_final_lev
// (:83-87, :539-545, :963-967, :1158-1168) and the sequential chain of
// *_smart_lev (:167-175, :629-640, :1048-1055, :1258-1273).
//
// CHAIN mode (discrete): one thread per investor, G wealth registers, the exact
//   fp32 product ((V0*m_0)*m_1)*... in time order.  Outcome bytes [N,ld] are
//   staged through shared memory in [128 investors x 128 steps] tiles by TMA
//   (cp.async.bulk.tensor.2d, 128-byte swizzle => the per-thread row reads are
//   16-byte LDS without bank conflicts) behind a ring of mbarriers.  Rows whose
//   stride or base is not 16-byte aligned take a cooperative plain-load path
//   into the same swizzled layout.
// LOG mode (discrete): wealth depends on the outcomes only through their
//   counts, so one warp sweeps one investor row with coalesced 16-byte loads and
//   counts codes with dp4a; HBM-bound, G-independent.
// LOG mode (GBM): running sum of x with its running extremes (to reproduce the
//   reference dtype's overflow/underflow saturation), thread per investor over
//   TMA-staged fp32 tiles, or Philox + Box-Muller draws in registers.
#include <cuda.h>

#include <cmath>
#include <cstring>

#include "common.cuh"

namespace b200 {

constexpr int TILE_ROWS = 128;   // investors per block
constexpr int TILE_BYTES = 128;  // bytes of one investor's row per tile (swizzle span)
constexpr int STAGES = 4;
constexpr int TILE_SMEM = TILE_ROWS * TILE_BYTES;  // 16 KB

struct FactorTable {
  float m[B200_MAX_OUTCOMES][32];  // [k][g] for one grid tile of <= 32 points
};
struct LogFactorTable {
  double lm[B200_MAX_OUTCOMES][B200_MAX_GRID];  // log m[k][g]
};
struct LevGrid {
  float lev[B200_MAX_GRID];
};

// ------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// 16-byte chunk `c` of tile row `r` under the 128-byte swizzle.
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)(r * TILE_BYTES + ((c ^ (r & 7)) << 4)); }

// ------------------------------------------------------- one chain step
// Three bit-identical ways to apply w[g] *= m[code][g] for the whole grid tile
// (each is one IEEE fp32 multiply per path-step; they differ in how the factor
// is selected, i.e. in which pipes they load - see DESIGN.md "chain variants"):
//   V_FSEL : per g, (K-1) FSEL from constant-bank factors + 1 FMUL
//   V_PRED2: K divergent arms of packed FMUL2 (mul.rn.f32x2) by constant-bank factors
//   V_LDS  : factor rows m[code][*] fetched from a shared-memory table with
//            LDS.128 (broadcast across lanes that saw the same code) + FMUL2
enum { V_FSEL = 0, V_PRED2 = 1, V_LDS = 2 };

template <int GT>
struct GridTile {
  static constexpr int PAD = (GT + 3) & ~3;                    // floats per table row, multiple of 4
  static constexpr int STRIDE = (PAD % 16 == 0) ? PAD + 4 : PAD;  // rows of different codes on disjoint banks
};

template <int GT, int K, int V>
struct ChainState {
  float w[GT];
  __device__ __forceinline__ void init(float v0) {
#pragma unroll
    for (int g = 0; g < GT; ++g) w[g] = v0;
  }
};

template <int GT, int K, int V>
__device__ __forceinline__ void chain_step(float (&w)[GT], const FactorTable& f, const float* __restrict__ tab,
                                           uint32_t code) {
  if (V == V_LDS) {
    constexpr int S = GridTile<GT>::STRIDE;
    const float4* __restrict__ row = reinterpret_cast<const float4*>(tab + code * S);
#pragma unroll
    for (int c = 0; c < GridTile<GT>::PAD / 4; ++c) {
      const float4 m = row[c];
      const int g = 4 * c;
      if (g + 1 < GT) {
        const float2 r = __fmul2_rn(make_float2(w[g], w[g + 1]), make_float2(m.x, m.y));
        w[g] = r.x; w[g + 1] = r.y;
      } else if (g < GT) {
        w[g] = __fmul_rn(w[g], m.x);
      }
      if (g + 3 < GT) {
        const float2 r = __fmul2_rn(make_float2(w[g + 2], w[g + 3]), make_float2(m.z, m.w));
        w[g + 2] = r.x; w[g + 3] = r.y;
      } else if (g + 2 < GT) {
        w[g + 2] = __fmul_rn(w[g + 2], m.z);
      }
    }
  } else if (V == V_PRED2) {
    // divergent branches: each arm is straight-line packed multiplies by
    // constant-bank factors; lanes that saw another code sit the arm out
    static_assert(GT % 2 == 0, "packed variant needs an even grid tile");
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (code == (uint32_t)k) {
#pragma unroll
        for (int g = 0; g < GT; g += 2) {
          const float2 r = __fmul2_rn(make_float2(w[g], w[g + 1]), make_float2(f.m[k][g], f.m[k][g + 1]));
          w[g] = r.x; w[g + 1] = r.y;
        }
      }
    }
  } else {
    const bool is1 = code == 1, is2 = code == 2, is3 = code == 3;
#pragma unroll
    for (int g = 0; g < GT; ++g) {
      float m = f.m[0][g];
      m = (K == 2 ? code != 0 : is1) ? f.m[1][g] : m;
      if (K >= 3) m = is2 ? f.m[2][g] : m;
      if (K >= 4) m = is3 ? f.m[3][g] : m;
      w[g] = __fmul_rn(w[g], m);
    }
  }
}

template <int GT, int K, int V>
__device__ __forceinline__ void chain_word(float (&w)[GT], const FactorTable& f, const float* __restrict__ tab,
                                           uint32_t word) {
#pragma unroll
  for (int b = 0; b < 4; ++b) chain_step<GT, K, V>(w, f, tab, (word >> (8 * b)) & 0xffu);
}

template <int GT, int K>
__device__ __forceinline__ void fill_table(float* tab, const FactorTable& f) {
  constexpr int S = GridTile<GT>::STRIDE;
  for (int i = threadIdx.x; i < K * S; i += blockDim.x) {
    const int k = i / S, g = i - k * S;
    tab[i] = g < GT ? f.m[k][g] : 1.0f;
  }
}

// --------------------------------------------- CHAIN, discrete, streamed
// USE_TMA: tiles arrive by cp.async.bulk.tensor; otherwise all threads copy.
template <int GT, int K, int V, bool USE_TMA>
__global__ void __launch_bounds__(TILE_ROWS)
chain_discrete_stream_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ outcomes,
                             int64_t ld, const __grid_constant__ FactorTable f, int32_t H, int64_t N,
                             int32_t G, float V0, float* __restrict__ data_T, int64_t ldT) {
  extern __shared__ __align__(1024) uint8_t tiles[];
  __shared__ __align__(8) uint64_t full[STAGES];
  __shared__ __align__(16) float tab[V == V_LDS ? K * GridTile<GT>::STRIDE : 4];

  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * TILE_ROWS;
  const int ntiles = (H + TILE_BYTES - 1) / TILE_BYTES;

  if (V == V_LDS) fill_table<GT, K>(tab, f);
  if (USE_TMA) {
    if (tid == 0) {
      for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
      fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
      const int pre = ntiles < STAGES ? ntiles : STAGES;
      for (int s = 0; s < pre; ++s) {
        mbar_expect_tx(&full[s], TILE_SMEM);
        tma_load_2d(tiles + s * TILE_SMEM, &tmap, &full[s], s * TILE_BYTES, (int)row0);
      }
    }
  } else {
    __syncthreads();
  }

  float w[GT];
#pragma unroll
  for (int g = 0; g < GT; ++g) w[g] = V0;

  for (int kt = 0; kt < ntiles; ++kt) {
    const int s = USE_TMA ? kt % STAGES : 0;
    uint8_t* tile = tiles + s * TILE_SMEM;
    if (USE_TMA) {
      mbar_wait(&full[s], (uint32_t)((kt / STAGES) & 1));
    } else {
      // cooperative copy: consecutive threads read consecutive bytes of a row
      __syncthreads();
      const int t0 = kt * TILE_BYTES;
      for (int idx = tid; idx < TILE_ROWS * TILE_BYTES; idx += TILE_ROWS) {
        const int r = idx >> 7, b = idx & 127;
        const int64_t row = row0 + r;
        uint8_t v = 0;
        if (row < N && t0 + b < H) v = outcomes[row * ld + t0 + b];
        tile[swz(r, b >> 4) + (b & 15)] = v;
      }
      __syncthreads();
    }
    const int steps = min(TILE_BYTES, H - kt * TILE_BYTES);
    if (steps == TILE_BYTES) {
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        const uint4 q = *reinterpret_cast<const uint4*>(tile + swz(tid, c));
        chain_word<GT, K, V>(w, f, tab, q.x);
        chain_word<GT, K, V>(w, f, tab, q.y);
        chain_word<GT, K, V>(w, f, tab, q.z);
        chain_word<GT, K, V>(w, f, tab, q.w);
      }
    } else {
#pragma unroll 1
      for (int t = 0; t < steps; ++t) {
        const uint32_t code = tile[swz(tid, t >> 4) + (t & 15)];
        chain_step<GT, K, V>(w, f, tab, code);
      }
    }
    if (USE_TMA) {
      __syncthreads();  // every thread is done with stage s
      if (tid == 0 && kt + STAGES < ntiles) {
        mbar_expect_tx(&full[s], TILE_SMEM);
        tma_load_2d(tile, &tmap, &full[s], (kt + STAGES) * TILE_BYTES, (int)row0);
      }
    }
  }

  const int64_t row = row0 + tid;
  if (row < N) {
#pragma unroll
    for (int g = 0; g < GT; ++g)
      if (g < G) data_T[(int64_t)g * ldT + row] = w[g];
  }
}

// ------------------------------------------------- Philox outcome draws
// Four consecutive time steps share one Philox block: counter =
// (investor lo, investor hi, t/4, TAG), key = seed.  Code = #{k: u >= thr[k]}.
struct Thresholds {
  uint32_t t[B200_MAX_OUTCOMES];
};
template <int K>
__device__ __forceinline__ uint32_t draw_code(uint32_t u, const Thresholds& th) {
  uint32_t c = (u >= th.t[0]);
  if (K >= 3) c += (u >= th.t[1]);
  if (K >= 4) c += (u >= th.t[2]);
  return c;
}

template <int GT, int K, int V>
__global__ void __launch_bounds__(128)
chain_discrete_philox_kernel(const __grid_constant__ FactorTable f, const __grid_constant__ Thresholds th,
                             uint64_t seed, int64_t investor_offset, int32_t H, int64_t N, int32_t G, float V0,
                             float* __restrict__ data_T, int64_t ldT) {
  __shared__ __align__(16) float tab[V == V_LDS ? K * GridTile<GT>::STRIDE : 4];
  if (V == V_LDS) {
    fill_table<GT, K>(tab, f);
    __syncthreads();
  }
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= N) return;
  const uint64_t id = (uint64_t)(row + investor_offset);
  const uint32_t c0 = (uint32_t)id, c1 = (uint32_t)(id >> 32);
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  float w[GT];
#pragma unroll
  for (int g = 0; g < GT; ++g) w[g] = V0;
  const int nblk = H >> 2;
#pragma unroll 2
  for (int j = 0; j < nblk; ++j) {
    const Philox4 r = philox4x32_10(c0, c1, (uint32_t)j, PHILOX_TAG_LEV, k0, k1);
    chain_step<GT, K, V>(w, f, tab, draw_code<K>(r.x, th));
    chain_step<GT, K, V>(w, f, tab, draw_code<K>(r.y, th));
    chain_step<GT, K, V>(w, f, tab, draw_code<K>(r.z, th));
    chain_step<GT, K, V>(w, f, tab, draw_code<K>(r.w, th));
  }
  if (H & 3) {
    const Philox4 r = philox4x32_10(c0, c1, (uint32_t)nblk, PHILOX_TAG_LEV, k0, k1);
    const uint32_t u[4] = {r.x, r.y, r.z, r.w};
    for (int t = 0; t < (H & 3); ++t) chain_step<GT, K, V>(w, f, tab, draw_code<K>(u[t], th));
  }
#pragma unroll
  for (int g = 0; g < GT; ++g)
    if (g < G) data_T[(int64_t)g * ldT + row] = w[g];
}

template <int K>
__global__ void __launch_bounds__(128)
draw_discrete_kernel(const __grid_constant__ Thresholds th, uint64_t seed, int64_t investor_offset, int32_t H,
                     int64_t N, int64_t ld, uint8_t* __restrict__ out) {
  // one thread per (investor, block of 4 steps)
  const int nblk = (H + 3) >> 2;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * nblk) return;
  const int64_t row = idx / nblk;
  const int j = (int)(idx - row * nblk);
  const uint64_t id = (uint64_t)(row + investor_offset);
  const Philox4 r = philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)j, PHILOX_TAG_LEV,
                                  (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t u[4] = {r.x, r.y, r.z, r.w};
  for (int b = 0; b < 4; ++b) {
    const int t = j * 4 + b;
    if (t < H) out[row * ld + t] = (uint8_t)draw_code<K>(u[b], th);
  }
}

// ------------------------------------------------ LOG, discrete: counting
// One warp per investor row; lanes stride over 16-byte chunks (coalesced 512 B
// per warp instruction).  Codes 0..3 are counted with dp4a on bit-planes:
//   S1 = sum(code & 1), S2 = sum(code >> 1)   =>   for K <= 3 (codes 0,1,2):
//   n1 = S1, n2 = S2, n0 = H - n1 - n2; K = 4 adds n3 via (code == 3).
template <int K>
__device__ __forceinline__ void count_word(uint32_t wd, uint32_t& s1, uint32_t& s2, uint32_t& s3) {
  s1 = __dp4a(wd & 0x01010101u, 0x01010101u, s1);
  if (K >= 3) s2 = __dp4a((wd >> 1) & 0x01010101u, 0x01010101u, s2);
  if (K >= 4) s3 = __dp4a(wd & (wd >> 1) & 0x01010101u, 0x01010101u, s3);
}

constexpr int COUNT_WARPS = 8;

template <int K>
__global__ void __launch_bounds__(COUNT_WARPS * 32)
log_discrete_stream_kernel(const uint8_t* __restrict__ outcomes, int64_t ld, int32_t H, int64_t N, int32_t G,
                           const __grid_constant__ LogFactorTable lf, double logV0, float* __restrict__ data_T,
                           double* __restrict__ log_w, int32_t* __restrict__ counts, int64_t ldT) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * COUNT_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * COUNT_WARPS;
  for (int64_t row = warp_global; row < N; row += nwarps) {
    const uint8_t* __restrict__ p = outcomes + row * ld;
    uint32_t s1 = 0, s2 = 0, s3 = 0;
    // head bytes up to 16-byte alignment, body in uint4, tail bytes
    const uintptr_t addr = (uintptr_t)p;
    int head = (int)((16 - (addr & 15)) & 15);
    if (head > H) head = H;
    const int body = (H - head) >> 4;
    const int tail0 = head + (body << 4);
    for (int t = lane; t < head; t += 32) {
      const uint32_t c = p[t];
      s1 += c & 1; if (K >= 3) s2 += (c >> 1) & 1; if (K >= 4) s3 += (c == 3);
    }
    const uint4* __restrict__ q = reinterpret_cast<const uint4*>(p + head);
    int i = lane;
    for (; i + 96 < body; i += 128) {  // 4 independent 16-byte loads in flight per lane
      const uint4 a = __ldcs(q + i), b = __ldcs(q + i + 32), c = __ldcs(q + i + 64), d = __ldcs(q + i + 96);
      count_word<K>(a.x, s1, s2, s3); count_word<K>(a.y, s1, s2, s3);
      count_word<K>(a.z, s1, s2, s3); count_word<K>(a.w, s1, s2, s3);
      count_word<K>(b.x, s1, s2, s3); count_word<K>(b.y, s1, s2, s3);
      count_word<K>(b.z, s1, s2, s3); count_word<K>(b.w, s1, s2, s3);
      count_word<K>(c.x, s1, s2, s3); count_word<K>(c.y, s1, s2, s3);
      count_word<K>(c.z, s1, s2, s3); count_word<K>(c.w, s1, s2, s3);
      count_word<K>(d.x, s1, s2, s3); count_word<K>(d.y, s1, s2, s3);
      count_word<K>(d.z, s1, s2, s3); count_word<K>(d.w, s1, s2, s3);
    }
    for (; i < body; i += 32) {
      const uint4 a = __ldcs(q + i);
      count_word<K>(a.x, s1, s2, s3); count_word<K>(a.y, s1, s2, s3);
      count_word<K>(a.z, s1, s2, s3); count_word<K>(a.w, s1, s2, s3);
    }
    for (int t = tail0 + lane; t < H; t += 32) {
      const uint32_t c = p[t];
      s1 += c & 1; if (K >= 3) s2 += (c >> 1) & 1; if (K >= 4) s3 += (c == 3);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      if (K >= 3) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      if (K >= 4) s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    }
    int n[4];
    if (K == 2) { n[1] = (int)s1; n[0] = H - n[1]; n[2] = n[3] = 0; }
    else if (K == 3) { n[1] = (int)s1; n[2] = (int)s2; n[0] = H - n[1] - n[2]; n[3] = 0; }
    else { n[3] = (int)s3; n[1] = (int)(s1 - s3); n[2] = (int)(s2 - s3); n[0] = H - n[1] - n[2] - n[3]; }
    if (counts != nullptr && lane < K) counts[row * K + lane] = n[lane];
    for (int g = lane; g < G; g += 32) {
      double lw = logV0;
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (n[k] > 0) lw += (double)n[k] * lf.lm[k][g];
      if (log_w != nullptr) log_w[(int64_t)g * ldT + row] = lw;
      if (data_T != nullptr) data_T[(int64_t)g * ldT + row] = (float)exp(lw);
    }
  }
}

// ----------------------------------------------------------- LOG, GBM
// Saturation of the reference's fp32 chain, decided from the running extremes
// of log-wealth: once the running wealth exceeded FLT_MAX it is inf for good,
// once it fell below half the smallest denormal it is 0 for good.
__device__ __forceinline__ void gbm_finish(double S, double Smax, double Smin, int32_t G,
                                           const LevGrid& lv, double logV0, int64_t row, int64_t ldT,
                                           float* __restrict__ data_T, double* __restrict__ log_w) {
  const double LOG_FLT_MAX = 88.72283905206835;    // ln(3.4028234664e38)
  const double LOG_FLT_ZERO = -103.97207708399179; // ln(2^-150)
  for (int g = 0; g < G; ++g) {
    const double l = (double)lv.lev[g];
    const double lw = logV0 + l * S;
    const double hi = logV0 + (l >= 0 ? l * Smax : l * Smin);
    const double lo = logV0 + (l >= 0 ? l * Smin : l * Smax);
    if (log_w != nullptr) log_w[(int64_t)g * ldT + row] = lw;
    if (data_T != nullptr) {
      float v;
      if (hi > LOG_FLT_MAX) v = __int_as_float(0x7f800000);
      else if (lo < LOG_FLT_ZERO) v = 0.0f;
      else v = (float)exp(lw);
      data_T[(int64_t)g * ldT + row] = v;
    }
  }
}

// 32 steps in fp32 relative to the fp64 base, then fold into the base.
struct GbmAcc {
  double S = 0.0, Smax = 0.0, Smin = 0.0;  // running sum and its extremes (t >= 0; start value 0)
  float p = 0.f, pmax = -3.0e38f, pmin = 3.0e38f;
  __device__ __forceinline__ void step(float x) {
    p += x;
    pmax = fmaxf(pmax, p);
    pmin = fminf(pmin, p);
  }
  __device__ __forceinline__ void fold() {
    Smax = fmax(Smax, S + (double)pmax);
    Smin = fmin(Smin, S + (double)pmin);
    S += (double)p;
    p = 0.f; pmax = -3.0e38f; pmin = 3.0e38f;
  }
};

template <bool USE_TMA>
__global__ void __launch_bounds__(TILE_ROWS)
log_gbm_stream_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ x, int64_t ld,
                      const __grid_constant__ LevGrid lv, int32_t H, int64_t N, int32_t G, double logV0,
                      float* __restrict__ data_T, double* __restrict__ log_w, int64_t ldT) {
  extern __shared__ __align__(1024) uint8_t tiles[];
  __shared__ __align__(8) uint64_t full[STAGES];
  constexpr int TILE_STEPS = TILE_BYTES / 4;  // 32 floats per row per tile

  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * TILE_ROWS;
  const int ntiles = (H + TILE_STEPS - 1) / TILE_STEPS;

  if (USE_TMA) {
    if (tid == 0) {
      for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
      fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
      const int pre = ntiles < STAGES ? ntiles : STAGES;
      for (int s = 0; s < pre; ++s) {
        mbar_expect_tx(&full[s], TILE_SMEM);
        tma_load_2d(tiles + s * TILE_SMEM, &tmap, &full[s], s * TILE_STEPS, (int)row0);
      }
    }
  }

  GbmAcc acc;
  for (int kt = 0; kt < ntiles; ++kt) {
    const int s = USE_TMA ? kt % STAGES : 0;
    uint8_t* tile = tiles + s * TILE_SMEM;
    if (USE_TMA) {
      mbar_wait(&full[s], (uint32_t)((kt / STAGES) & 1));
    } else {
      __syncthreads();
      const int t0 = kt * TILE_STEPS;
      for (int idx = tid; idx < TILE_ROWS * TILE_STEPS; idx += TILE_ROWS) {
        const int r = idx >> 5, e = idx & 31;
        const int64_t row = row0 + r;
        float v = 0.f;
        if (row < N && t0 + e < H) v = x[row * ld + t0 + e];
        *reinterpret_cast<float*>(tile + swz(r, e >> 2) + (e & 3) * 4) = v;
      }
      __syncthreads();
    }
    const int steps = min(TILE_STEPS, H - kt * TILE_STEPS);
    if (steps == TILE_STEPS) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 q = *reinterpret_cast<const float4*>(tile + swz(tid, c));
        acc.step(q.x); acc.step(q.y); acc.step(q.z); acc.step(q.w);
      }
    } else {
      for (int t = 0; t < steps; ++t)
        acc.step(*reinterpret_cast<const float*>(tile + swz(tid, t >> 2) + (t & 3) * 4));
    }
    acc.fold();
    if (USE_TMA) {
      __syncthreads();
      if (tid == 0 && kt + STAGES < ntiles) {
        mbar_expect_tx(&full[s], TILE_SMEM);
        tma_load_2d(tile, &tmap, &full[s], (kt + STAGES) * TILE_STEPS, (int)row0);
      }
    }
  }
  const int64_t row = row0 + tid;
  if (row < N) gbm_finish(acc.S, acc.Smax, acc.Smin, G, lv, logV0, row, ldT, data_T, log_w);
}

// x_t = log_mean + sigma * z_t; four steps per Philox block, two Box-Muller pairs.
__device__ __forceinline__ void gbm_draw4(uint32_t c0, uint32_t c1, uint32_t j, uint32_t k0, uint32_t k1,
                                          float log_mean, float sigma, float (&x)[4]) {
  const Philox4 r = philox4x32_10(c0, c1, j, PHILOX_TAG_LEV, k0, k1);
  float z0, z1, z2, z3;
  box_muller(r.x, r.y, z0, z1);
  box_muller(r.z, r.w, z2, z3);
  x[0] = fmaf(sigma, z0, log_mean);
  x[1] = fmaf(sigma, z1, log_mean);
  x[2] = fmaf(sigma, z2, log_mean);
  x[3] = fmaf(sigma, z3, log_mean);
}

__global__ void __launch_bounds__(128)
log_gbm_philox_kernel(const __grid_constant__ LevGrid lv, uint64_t seed, int64_t investor_offset, float log_mean,
                      float sigma, int32_t H, int64_t N, int32_t G, double logV0, float* __restrict__ data_T,
                      double* __restrict__ log_w, int64_t ldT) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= N) return;
  const uint64_t id = (uint64_t)(row + investor_offset);
  const uint32_t c0 = (uint32_t)id, c1 = (uint32_t)(id >> 32);
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  GbmAcc acc;
  const int nblk = H >> 2;
  int j = 0;
  for (; j + 8 <= nblk; j += 8) {  // fold every 32 steps
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float x[4];
      gbm_draw4(c0, c1, (uint32_t)(j + u), k0, k1, log_mean, sigma, x);
      acc.step(x[0]); acc.step(x[1]); acc.step(x[2]); acc.step(x[3]);
    }
    acc.fold();
  }
  for (; j < nblk; ++j) {
    float x[4];
    gbm_draw4(c0, c1, (uint32_t)j, k0, k1, log_mean, sigma, x);
    acc.step(x[0]); acc.step(x[1]); acc.step(x[2]); acc.step(x[3]);
  }
  if (H & 3) {
    float x[4];
    gbm_draw4(c0, c1, (uint32_t)nblk, k0, k1, log_mean, sigma, x);
    for (int t = 0; t < (H & 3); ++t) acc.step(x[t]);
  }
  acc.fold();
  gbm_finish(acc.S, acc.Smax, acc.Smin, G, lv, logV0, row, ldT, data_T, log_w);
}

__global__ void __launch_bounds__(128)
draw_gbm_kernel(uint64_t seed, int64_t investor_offset, float log_mean, float sigma, int32_t H, int64_t N,
                int64_t ld, float* __restrict__ out) {
  const int nblk = (H + 3) >> 2;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * nblk) return;
  const int64_t row = idx / nblk;
  const int j = (int)(idx - row * nblk);
  const uint64_t id = (uint64_t)(row + investor_offset);
  float x[4];
  gbm_draw4((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)j, (uint32_t)seed, (uint32_t)(seed >> 32), log_mean,
            sigma, x);
  for (int b = 0; b < 4; ++b) {
    const int t = j * 4 + b;
    if (t < H) out[row * ld + t] = x[b];
  }
}

// ------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      cudaGetLastError();
  }
  return fn;
}

// [N rows, row_bytes] byte view with 128-byte x 128-row boxes, 128B swizzle.
static int make_row_tile_map(CUtensorMap* map, const void* base, int64_t n_rows, int64_t row_elems,
                             int64_t ld_elems, int elem_bytes) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return set_error(B200_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2] = {(cuuint64_t)row_elems, (cuuint64_t)n_rows};
  const cuuint64_t strides[1] = {(cuuint64_t)(ld_elems * elem_bytes)};
  const cuuint32_t box[2] = {(cuuint32_t)(TILE_BYTES / elem_bytes), (cuuint32_t)TILE_ROWS};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(B200_ECUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return 0;
}

// 1 = V_FSEL, 2 = V_PRED2, 3 = V_LDS; chosen from the B200 measurements in profiles/
static int default_chain_variant() { return 1; }

static bool tma_ok(const void* base, int64_t ld_bytes, int64_t n_rows) {
  return ((uintptr_t)base % 16 == 0) && (ld_bytes % 16 == 0) && n_rows < (int64_t)1 << 31;
}

template <int GT, int K, int V>
static int launch_chain_discrete(const b200_lev_desc& d, const uint8_t* outcomes, const FactorTable& f, int g_cnt,
                                 float* data_T, cudaStream_t st) {
  const int64_t N = d.n_investors;
  if (d.source == B200_SRC_PHILOX) {
    Thresholds th;
    for (int k = 0; k < B200_MAX_OUTCOMES; ++k) th.t[k] = d.thresholds[k];
    const unsigned blocks = (unsigned)((N + 127) / 128);
    chain_discrete_philox_kernel<GT, K, V><<<blocks, 128, 0, st>>>(f, th, d.seed, d.investor_offset, d.horizon, N,
                                                                   g_cnt, d.value_0, data_T, N);
    return check_cuda(cudaGetLastError(), "chain_discrete_philox launch");
  }
  const unsigned blocks = (unsigned)((N + TILE_ROWS - 1) / TILE_ROWS);
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (tma_ok(outcomes, d.ld_outcomes, N)) {
    int rc = make_row_tile_map(&map, outcomes, N, d.horizon, d.ld_outcomes, 1);
    if (rc) return rc;
    auto kern = chain_discrete_stream_kernel<GT, K, V, true>;
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * TILE_SMEM));
    kern<<<blocks, TILE_ROWS, STAGES * TILE_SMEM, st>>>(map, outcomes, d.ld_outcomes, f, d.horizon, N, g_cnt,
                                                        d.value_0, data_T, N);
  } else {
    chain_discrete_stream_kernel<GT, K, V_FSEL, false><<<blocks, TILE_ROWS, TILE_SMEM, st>>>(
        map, outcomes, d.ld_outcomes, f, d.horizon, N, g_cnt, d.value_0, data_T, N);
  }
  return check_cuda(cudaGetLastError(), "chain_discrete_stream launch");
}

template <int K, int V>
static int dispatch_chain_gt(const b200_lev_desc& d, const uint8_t* outcomes, const FactorTable& f, int g_cnt,
                             float* data_T, cudaStream_t st) {
  if (g_cnt <= 4) return launch_chain_discrete<4, K, V>(d, outcomes, f, g_cnt, data_T, st);
  if (g_cnt <= 10) return launch_chain_discrete<10, K, V>(d, outcomes, f, g_cnt, data_T, st);
  if (g_cnt <= 20) return launch_chain_discrete<20, K, V>(d, outcomes, f, g_cnt, data_T, st);
  return launch_chain_discrete<32, K, V>(d, outcomes, f, g_cnt, data_T, st);
}

template <int K>
static int dispatch_chain_variant(const b200_lev_desc& d, const uint8_t* outcomes, const FactorTable& f, int g_cnt,
                                  float* data_T, cudaStream_t st) {
  int v = d.variant;
  if (v == 0) v = default_chain_variant();
  switch (v) {
    case 2: return dispatch_chain_gt<K, V_PRED2>(d, outcomes, f, g_cnt, data_T, st);
    case 3: return dispatch_chain_gt<K, V_LDS>(d, outcomes, f, g_cnt, data_T, st);
    default: return dispatch_chain_gt<K, V_FSEL>(d, outcomes, f, g_cnt, data_T, st);
  }
}

static int run_chain_discrete(const b200_lev_desc& d, const uint8_t* outcomes, const float* factors_host,
                              float* data_T, cudaStream_t st) {
  // grid tiles of <= 32 points; each tile re-streams (or re-draws) the outcomes
  for (int g0 = 0; g0 < d.n_grid; g0 += 32) {
    const int g_cnt = d.n_grid - g0 < 32 ? d.n_grid - g0 : 32;
    FactorTable f;
    for (int k = 0; k < B200_MAX_OUTCOMES; ++k)
      for (int g = 0; g < 32; ++g)
        f.m[k][g] = (g < g_cnt && k < d.n_outcomes) ? factors_host[(int64_t)(g0 + g) * d.n_outcomes + k] : 1.0f;
    float* out = data_T + (int64_t)g0 * d.n_investors;
    int rc;
    switch (d.n_outcomes) {
      case 2: rc = dispatch_chain_variant<2>(d, outcomes, f, g_cnt, out, st); break;
      case 3: rc = dispatch_chain_variant<3>(d, outcomes, f, g_cnt, out, st); break;
      default: rc = dispatch_chain_variant<4>(d, outcomes, f, g_cnt, out, st); break;
    }
    if (rc) return rc;
  }
  return 0;
}

static int run_log_discrete(const b200_lev_desc& d, const uint8_t* outcomes, const float* factors_host,
                            float* data_T, double* log_w, int32_t* counts, cudaStream_t st) {
  LogFactorTable lf;
  for (int k = 0; k < B200_MAX_OUTCOMES; ++k)
    for (int g = 0; g < B200_MAX_GRID; ++g) {
      double v = 0.0;
      if (g < d.n_grid && k < d.n_outcomes) {
        const float m = factors_host[(int64_t)g * d.n_outcomes + k];
        if (m < 0.0f) return set_error(B200_EINVAL, "LOG mode needs factors >= 0 (m[%d][%d] = %g)", g, k, (double)m);
        v = log((double)m);
      }
      lf.lm[k][g] = v;
    }
  const double logV0 = log((double)d.value_0);
  const int64_t N = d.n_investors;
  int64_t blocks = (N + COUNT_WARPS - 1) / COUNT_WARPS;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  switch (d.n_outcomes) {
    case 2:
      log_discrete_stream_kernel<2><<<(unsigned)blocks, COUNT_WARPS * 32, 0, st>>>(
          outcomes, d.ld_outcomes, d.horizon, N, d.n_grid, lf, logV0, data_T, log_w, counts, N);
      break;
    case 3:
      log_discrete_stream_kernel<3><<<(unsigned)blocks, COUNT_WARPS * 32, 0, st>>>(
          outcomes, d.ld_outcomes, d.horizon, N, d.n_grid, lf, logV0, data_T, log_w, counts, N);
      break;
    default:
      log_discrete_stream_kernel<4><<<(unsigned)blocks, COUNT_WARPS * 32, 0, st>>>(
          outcomes, d.ld_outcomes, d.horizon, N, d.n_grid, lf, logV0, data_T, log_w, counts, N);
      break;
  }
  return check_cuda(cudaGetLastError(), "log_discrete_stream launch");
}

static int run_log_gbm(const b200_lev_desc& d, const float* x, const float* lev_host, float* data_T, double* log_w,
                       cudaStream_t st) {
  LevGrid lv;
  for (int g = 0; g < B200_MAX_GRID; ++g) lv.lev[g] = g < d.n_grid ? lev_host[g] : 0.f;
  const double logV0 = log((double)d.value_0);
  const int64_t N = d.n_investors;
  if (d.source == B200_SRC_PHILOX) {
    const unsigned blocks = (unsigned)((N + 127) / 128);
    log_gbm_philox_kernel<<<blocks, 128, 0, st>>>(lv, d.seed, d.investor_offset, d.log_mean, d.sigma, d.horizon, N,
                                                  d.n_grid, logV0, data_T, log_w, N);
    return check_cuda(cudaGetLastError(), "log_gbm_philox launch");
  }
  const unsigned blocks = (unsigned)((N + TILE_ROWS - 1) / TILE_ROWS);
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (tma_ok(x, d.ld_outcomes * 4, N)) {
    int rc = make_row_tile_map(&map, x, N, d.horizon, d.ld_outcomes, 4);
    if (rc) return rc;
    auto kern = log_gbm_stream_kernel<true>;
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * TILE_SMEM));
    kern<<<blocks, TILE_ROWS, STAGES * TILE_SMEM, st>>>(map, x, d.ld_outcomes, lv, d.horizon, N, d.n_grid, logV0,
                                                        data_T, log_w, N);
  } else {
    log_gbm_stream_kernel<false><<<blocks, TILE_ROWS, TILE_SMEM, st>>>(map, x, d.ld_outcomes, lv, d.horizon, N,
                                                                       d.n_grid, logV0, data_T, log_w, N);
  }
  return check_cuda(cudaGetLastError(), "log_gbm_stream launch");
}

static int validate(const b200_lev_desc* d) {
  B200_REQUIRE(d != nullptr, "lev: desc is NULL");
  B200_REQUIRE(d->n_investors >= 0 && d->n_investors < ((int64_t)1 << 40), "lev: n_investors out of range");
  B200_REQUIRE(d->horizon >= 1, "lev: horizon must be >= 1");
  B200_REQUIRE(d->n_grid >= 1 && d->n_grid <= B200_MAX_GRID, "lev: n_grid must be in 1..%d", B200_MAX_GRID);
  B200_REQUIRE(d->kind == B200_LEV_DISCRETE || d->kind == B200_LEV_GBM, "lev: unknown kind %d", d->kind);
  B200_REQUIRE(d->source == B200_SRC_STREAM || d->source == B200_SRC_PHILOX, "lev: unknown source %d", d->source);
  B200_REQUIRE(d->variant >= 0 && d->variant <= 3, "lev: variant must be 0 (auto) .. 3");
  if (d->kind == B200_LEV_DISCRETE) {
    B200_REQUIRE(d->n_outcomes >= 2 && d->n_outcomes <= B200_MAX_OUTCOMES, "lev: n_outcomes must be in 2..%d",
                 B200_MAX_OUTCOMES);
    if (d->source == B200_SRC_PHILOX)
      for (int k = 1; k < d->n_outcomes - 1; ++k)
        B200_REQUIRE(d->thresholds[k] >= d->thresholds[k - 1], "lev: thresholds must ascend");
  }
  if (d->source == B200_SRC_STREAM) B200_REQUIRE(d->ld_outcomes >= d->horizon, "lev: ld_outcomes < horizon");
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_lev_sweep(const b200_lev_desc* desc, const void* outcomes, const float* factors, float* data_T,
                              double* log_w, int32_t* counts, void* stream) {
  int rc = validate(desc);
  if (rc) return rc;
  const b200_lev_desc& d = *desc;
  cudaStream_t st = (cudaStream_t)stream;
  B200_REQUIRE(factors != nullptr, "lev_sweep: factors is NULL");
  B200_REQUIRE(d.source == B200_SRC_PHILOX || outcomes != nullptr || d.n_investors == 0,
               "lev_sweep: outcomes is NULL for a streamed sweep");
  if (d.n_investors == 0) return 0;

  // the tiny factor table travels as a kernel parameter
  const float* host_f = factors;

  if (d.kind == B200_LEV_DISCRETE) {
    if (d.mode == B200_MODE_CHAIN) {
      B200_REQUIRE(data_T != nullptr, "lev_sweep: CHAIN mode needs data_T");
      return run_chain_discrete(d, (const uint8_t*)outcomes, host_f, data_T, st);
    }
    if (d.mode == B200_MODE_LOG) {
      B200_REQUIRE(d.source == B200_SRC_STREAM, "lev_sweep: discrete LOG mode takes streamed outcomes");
      B200_REQUIRE(data_T || log_w || counts, "lev_sweep: LOG mode needs at least one output");
      return run_log_discrete(d, (const uint8_t*)outcomes, host_f, data_T, log_w, counts, st);
    }
    return set_error(B200_EINVAL, "lev_sweep: unknown mode %d", d.mode);
  }
  B200_REQUIRE(d.mode == B200_MODE_LOG, "lev_sweep: GBM runs in LOG mode only (expf chains are not reproducible)");
  B200_REQUIRE(data_T || log_w, "lev_sweep: LOG mode needs at least one output");
  return run_log_gbm(d, (const float*)outcomes, host_f, data_T, log_w, st);
}

extern "C" int b200_lev_draw(const b200_lev_desc* desc, void* out, void* stream) {
  int rc = validate(desc);
  if (rc) return rc;
  const b200_lev_desc& d = *desc;
  cudaStream_t st = (cudaStream_t)stream;
  B200_REQUIRE(out != nullptr || d.n_investors == 0, "lev_draw: out is NULL");
  B200_REQUIRE(d.ld_outcomes >= d.horizon, "lev_draw: ld_outcomes < horizon");
  if (d.n_investors == 0) return 0;
  const int64_t work = d.n_investors * (int64_t)((d.horizon + 3) >> 2);
  const unsigned blocks = (unsigned)((work + 127) / 128);
  if (d.kind == B200_LEV_GBM) {
    draw_gbm_kernel<<<blocks, 128, 0, st>>>(d.seed, d.investor_offset, d.log_mean, d.sigma, d.horizon,
                                            d.n_investors, d.ld_outcomes, (float*)out);
  } else {
    Thresholds th;
    for (int k = 0; k < B200_MAX_OUTCOMES; ++k) th.t[k] = d.thresholds[k];
    switch (d.n_outcomes) {
      case 2: draw_discrete_kernel<2><<<blocks, 128, 0, st>>>(th, d.seed, d.investor_offset, d.horizon,
                                                              d.n_investors, d.ld_outcomes, (uint8_t*)out); break;
      case 3: draw_discrete_kernel<3><<<blocks, 128, 0, st>>>(th, d.seed, d.investor_offset, d.horizon,
                                                              d.n_investors, d.ld_outcomes, (uint8_t*)out); break;
      default: draw_discrete_kernel<4><<<blocks, 128, 0, st>>>(th, d.seed, d.investor_offset, d.horizon,
                                                               d.n_investors, d.ld_outcomes, (uint8_t*)out); break;
    }
  }
  return check_cuda(cudaGetLastError(), "lev_draw launch");
}
