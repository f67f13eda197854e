// Ingest: the outcome arrays exactly as the reference's scripts hold them ->
// the engine's formats, counted on the way (b200_lev_ingest).
//
//   coin  lev/coin_flip.py:158-161   Bernoulli(p).sample((N,H))        fp32 {0,1}
//         lev/lev_exp.py:85          where(outcomes == 1, up, down)     code = (x == 1)
//   dice  lev/dice_roll.py:145-148   Categorical(probs).sample((N,H))   int64 {0,1,2}
//         lev/lev_exp.py:537-543     cast to fp32; == 0 up, == 1 down, == 2 mid;
//                                    any other value would be used AS the factor
//                                    (:541 `where(outcomes == 0, ., outcomes)`) -
//                                    counted in the tally's "bad" word, not supported
//
// One pass over the source at its own width (8 bytes per outcome for int64): every
// thread owns runs of 16 consecutive outcomes - 16-byte loads, one 16-byte store of
// the uint8 codes - and a row's counts are reduced in the warp / block that owns the
// row.  Sinks (any combination): uint8 codes for the CHAIN kernels, counts [N,K],
// and the tally of count tuples that feeds the final-time statistics (tally.cu).
// Bound by the source read: HBM for a device-resident array, PCIe when the caller
// streams chunks of a host array through a staging buffer (engine.py).
#include "tally.cuh"

namespace b200 {

template <typename T> struct SrcTraits;
template <> struct SrcTraits<uint8_t> { static constexpr int VEC = 16; };   // elements per 16-byte load
template <> struct SrcTraits<int32_t> { static constexpr int VEC = 4; };
template <> struct SrcTraits<int64_t> { static constexpr int VEC = 2; };
template <> struct SrcTraits<float> { static constexpr int VEC = 4; };
template <> struct SrcTraits<double> { static constexpr int VEC = 2; };

// code of one source value; `bad` counts values the discrete sweep cannot represent
template <typename T, int K>
__device__ __forceinline__ uint32_t code_of(T v, uint32_t& bad) {
  if (K == 2) return v == (T)1 ? 1u : 0u;   // the coin: everything that is not 1 is a down move
  uint32_t c = 0;
  bool ok = false;
#pragma unroll
  for (int k = 0; k < K; ++k)
    if (v == (T)k) { c = (uint32_t)k; ok = true; }
  bad += ok ? 0u : 1u;
  return c;
}

template <int K>
__device__ __forceinline__ void tally_code(uint32_t c, uint32_t& s1, uint32_t& s2, uint32_t& s3) {
  s1 += (c == 1u);
  if (K >= 3) s2 += (c == 2u);
  if (K >= 4) s3 += (c == 3u);
}

// RT threads per row (32: a warp per row, 8 rows per block; 256: a block per row)
template <typename T, int K, int RT>
__global__ void __launch_bounds__(256)
ingest_kernel(const T* __restrict__ src, int64_t ld_src, int32_t H, int64_t N, uint8_t* __restrict__ codes,
              int64_t ld_codes, int32_t* __restrict__ counts, const TallyDev tally, long long* __restrict__ bad_word) {
  constexpr int VEC = SrcTraits<T>::VEC;
  constexpr int LOADS = 16 / VEC;                       // 16-byte loads per run of 16 outcomes
  __shared__ uint32_t red[4][8];
  const int rows_per_block = 256 / RT;
  const int t = threadIdx.x % RT;
  const int64_t row = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / RT;
  uint32_t s1 = 0, s2 = 0, s3 = 0, bad = 0;
  if (row < N) {
    const T* __restrict__ p = src + row * ld_src;
    uint8_t* __restrict__ q = codes ? codes + row * ld_codes : nullptr;
    const bool vec_src = (((uintptr_t)p) & 15) == 0;
    const bool vec_dst = q == nullptr || (((uintptr_t)q) & 15) == 0;
    const int runs = H >> 4;
    if (vec_src && vec_dst) {
      for (int r = t; r < runs; r += RT) {
        const uint4* __restrict__ v4 = reinterpret_cast<const uint4*>(p + (int64_t)r * 16);
        uint4 raw[LOADS];
#pragma unroll
        for (int u = 0; u < LOADS; ++u) raw[u] = __ldcs(v4 + u);
        const T* e = reinterpret_cast<const T*>(raw);
        uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t c = code_of<T, K>(e[i], bad);
          tally_code<K>(c, s1, s2, s3);
          w[i >> 2] |= c << (8 * (i & 3));
        }
        if (q) *reinterpret_cast<uint4*>(q + (int64_t)r * 16) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    } else {
      for (int i = t; i < runs * 16; i += RT) {
        const uint32_t c = code_of<T, K>(p[i], bad);
        tally_code<K>(c, s1, s2, s3);
        if (q) q[i] = (uint8_t)c;
      }
    }
    for (int i = runs * 16 + t; i < H; i += RT) {
      const uint32_t c = code_of<T, K>(p[i], bad);
      tally_code<K>(c, s1, s2, s3);
      if (q) q[i] = (uint8_t)c;
    }
  }
  // reduce over the row's threads
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    if (K >= 3) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    if (K >= 4) s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    if (K >= 3) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if (RT == 256) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { red[0][wid] = s1; red[1][wid] = s2; red[2][wid] = s3; red[3][wid] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
      s1 = s2 = s3 = bad = 0;
      for (int w = 0; w < 8; ++w) { s1 += red[0][w]; s2 += red[1][w]; s3 += red[2][w]; bad += red[3][w]; }
    }
  }
  if (t == 0 && row < N) {
    int n[4];
    n[1] = (int)s1; n[2] = K >= 3 ? (int)s2 : 0; n[3] = K >= 4 ? (int)s3 : 0;
    n[0] = H - n[1] - n[2] - n[3];
    if (counts != nullptr)
      for (int k = 0; k < K; ++k) counts[row * K + k] = n[k];
    if (tally.keys != nullptr) tally_insert(tally, tally_key(n[1], n[2], n[3]), 1u);
    if (bad != 0 && bad_word != nullptr) atomicAdd(reinterpret_cast<unsigned long long*>(bad_word), (unsigned long long)bad);
  }
}

template <typename T, int K>
static int launch_ingest(const void* src, int64_t ld_src, int32_t H, int64_t N, uint8_t* codes, int64_t ld_codes,
                         int32_t* counts, const TallyDev& t, long long* bad, cudaStream_t st) {
  // short rows: a warp per row; long rows: a block per row (enough loads in flight either way)
  if ((int64_t)H * (int64_t)sizeof(T) <= 16384) {
    const int64_t blocks = (N + 7) / 8;
    B200_REQUIRE(blocks <= 0x7fffffff, "lev_ingest: too many rows for one launch");
    ingest_kernel<T, K, 32><<<(unsigned)blocks, 256, 0, st>>>((const T*)src, ld_src, H, N, codes, ld_codes, counts, t, bad);
  } else {
    B200_REQUIRE(N <= 0x7fffffff, "lev_ingest: too many rows for one launch");
    ingest_kernel<T, K, 256><<<(unsigned)N, 256, 0, st>>>((const T*)src, ld_src, H, N, codes, ld_codes, counts, t, bad);
  }
  return check_cuda(cudaGetLastError(), "lev_ingest launch");
}

template <typename T>
static int ingest_k(int K, const void* src, int64_t ld_src, int32_t H, int64_t N, uint8_t* codes, int64_t ld_codes,
                    int32_t* counts, const TallyDev& t, long long* bad, cudaStream_t st) {
  switch (K) {
    case 2: return launch_ingest<T, 2>(src, ld_src, H, N, codes, ld_codes, counts, t, bad, st);
    case 3: return launch_ingest<T, 3>(src, ld_src, H, N, codes, ld_codes, counts, t, bad, st);
    default: return launch_ingest<T, 4>(src, ld_src, H, N, codes, ld_codes, counts, t, bad, st);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_lev_ingest(const void* src, int32_t src_type, int64_t n_investors, int32_t horizon, int64_t ld_src,
                               int32_t n_outcomes, uint8_t* codes, int64_t ld_codes, int32_t* counts,
                               const b200_tally_plan* plan, void* workspace, void* stream) {
  B200_REQUIRE(n_investors >= 0 && horizon >= 1, "lev_ingest: need n_investors >= 0 and horizon >= 1");
  B200_REQUIRE(n_outcomes >= 2 && n_outcomes <= B200_MAX_OUTCOMES, "lev_ingest: n_outcomes must be in 2..%d",
               B200_MAX_OUTCOMES);
  B200_REQUIRE(ld_src >= horizon, "lev_ingest: ld_src < horizon");
  B200_REQUIRE(codes == nullptr || ld_codes >= horizon, "lev_ingest: ld_codes < horizon");
  B200_REQUIRE(codes != nullptr || counts != nullptr || plan != nullptr, "lev_ingest: no sink");
  TallyDev t{nullptr, nullptr, nullptr, 0};
  long long* bad = nullptr;
  if (plan != nullptr) {
    if (int rc = tally_device_view(plan, workspace, horizon, &t)) return rc;
    bad = t.header + TH_BAD;
  }
  if (n_investors == 0) return 0;
  B200_REQUIRE(src != nullptr, "lev_ingest: src is NULL");
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  cudaStream_t st = (cudaStream_t)stream;
  switch (src_type) {
    case B200_DT_U8: return ingest_k<uint8_t>(n_outcomes, src, ld_src, horizon, n_investors, codes, ld_codes, counts, t, bad, st);
    case B200_DT_I32: return ingest_k<int32_t>(n_outcomes, src, ld_src, horizon, n_investors, codes, ld_codes, counts, t, bad, st);
    case B200_DT_I64: return ingest_k<int64_t>(n_outcomes, src, ld_src, horizon, n_investors, codes, ld_codes, counts, t, bad, st);
    case B200_DT_F32: return ingest_k<float>(n_outcomes, src, ld_src, horizon, n_investors, codes, ld_codes, counts, t, bad, st);
    case B200_DT_F64: return ingest_k<double>(n_outcomes, src, ld_src, horizon, n_investors, codes, ld_codes, counts, t, bad, st);
    default: return set_error(B200_EINVAL, "lev_ingest: unknown src_type %d", src_type);
  }
}
