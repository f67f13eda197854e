// Row statistics: the reference's summary-statistic block (lev/lev_exp.py:89-104,
// :177-192 and 11 more inlined copies) for many rows at once, without sorting.
//
//   sort(descending); top = s[:K]; adj = s[K:];
//   mean, population std, MAD about the mean, lower median  of all / top / adj.
//
// Exact order statistics come from a 3-level radix select (11+11+10 bits) on the
// order-preserving key of the fp32 bit pattern; four target ranks are resolved
// together (App. B of SURVEY.md):
//   j=0 med_all  = (n-1)/2        j=1 thr     = n-K      (smallest of the top)
//   j=2 med_top  = n-K+(K-1)/2    j=3 med_adj = (n-K-1)/2
// Moments are accumulated in fp64, two-pass (mean first), and the top / adj sums
// are accumulated directly over {w > thr} and {w < thr} plus the tie share of
// thr itself (never as all-minus-top: the top-K hold most of the mass).
//
// Five passes over the data, each HBM-bound (4 B per element per pass):
//   pass 0: sum, level-1 histogram          pass 3: sums/counts above, below thr
//   pass 1: |w-mean|, (w-mean)^2, level 2   pass 4: deviations of top / adj
//   pass 2: level 3
// Between passes a one-block-per-row "resolve" kernel turns histograms into
// prefixes/ranks.  Every cross-block quantity lives in the caller's workspace as
// 8-byte words so that a multi-GPU caller can all-reduce it between passes.
#include "common.cuh"

namespace b200 {

constexpr int L1_BITS = 11, L2_BITS = 11, L3_BITS = 10;
constexpr int L1_BINS = 1 << L1_BITS, L2_BINS = 1 << L2_BITS, L3_BINS = 1 << L3_BITS;
constexpr int NT = 4;  // targets

// Per-row workspace, 8-byte words.  Doubles first, then integers.
struct RowWS {
  // ---- doubles (exchange region D) -------------------------------------
  double sum_all;      // pass 0
  double absdev_all;   // pass 1
  double sqdev_all;    // pass 1
  double sum_gt;       // pass 3: sum of w with key > key(thr)
  double sum_lt;       // pass 3
  double absdev_gt;    // pass 4 (about mean_top)
  double sqdev_gt;
  double absdev_lt;    // pass 4 (about mean_adj)
  double sqdev_lt;
  double pad_d[7];
  // ---- integers (exchange region I) ------------------------------------
  long long cnt_gt;    // pass 3
  long long cnt_lt;    // pass 3
  long long pad_i[6];
  long long hist1[L1_BINS];          // pass 0
  long long hist2[NT][L2_BINS];      // pass 1
  long long hist3[NT][L3_BINS];      // pass 2
  // ---- resolved by the row-resolve kernels (identical on every rank) ---
  long long rank[NT];       // remaining rank inside the current prefix
  unsigned long long prefix[NT];  // key prefix found so far (left aligned)
  double mean_all, mean_top, mean_adj;
  double value[NT];         // selected values (as double)
  long long ties_top, ties_adj;
  long long nonfinite[3];   // all / top / adj group holds +-inf or NaN (torch.std_mean -> nan)
  long long has_nan[3];     // all / top / adj group holds a NaN (torch.median -> nan)
  long long pad_r[2];
};
static_assert(sizeof(RowWS) % 8 == 0, "8-byte words");

constexpr int64_t D_WORDS = 16;  // doubles at the head
constexpr int64_t OFF_CNT = D_WORDS;
constexpr int64_t OFF_H1 = OFF_CNT + 8;
constexpr int64_t OFF_H2 = OFF_H1 + L1_BINS;
constexpr int64_t OFF_H3 = OFF_H2 + (int64_t)NT * L2_BINS;
constexpr int64_t OFF_RES = OFF_H3 + (int64_t)NT * L3_BINS;
constexpr int64_t ROW_WORDS = sizeof(RowWS) / 8;

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;  // elements per thread per block-slice iteration

__device__ __forceinline__ void atomic_add_f64(double* p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add_i64(long long* p, long long v) {
  atomicAdd((unsigned long long*)p, (unsigned long long)v);
}

// grid = (slices, rows).  Each block walks its slice of one row.
template <int PASS>
__global__ void __launch_bounds__(RS_THREADS)
rowstats_pass_kernel(const float* __restrict__ values, int64_t n, int64_t ld, RowWS* __restrict__ ws) {
  extern __shared__ unsigned int smem_hist[];
  __shared__ double red_d[32];
  __shared__ long long red_i[32];

  const int64_t row = blockIdx.y;
  const float* __restrict__ v = values + row * ld;
  RowWS* w = ws + row;

  constexpr int HBINS = PASS == 0 ? L1_BINS : PASS == 1 ? NT * L2_BINS : PASS == 2 ? NT * L3_BINS : 0;
  for (int i = threadIdx.x; i < HBINS; i += RS_THREADS) smem_hist[i] = 0;

  // Targets that fell into the same bin so far need the same histogram: only the
  // first of them (its representative) counts, the others copy it at the flush.
  // cmp[j] is the bin prefix a key must show to be counted for target j; a
  // duplicate gets a prefix no key can have.
  __shared__ int rep_s[NT];
  uint32_t cmp[NT];
  double mean_a = 0, mean_t = 0, mean_j = 0;
  uint32_t thr_key = 0;
  if (PASS == 1 || PASS == 2) {
    constexpr int SH = PASS == 1 ? 32 - L1_BITS : L3_BITS;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      cmp[j] = (uint32_t)w->prefix[j] >> SH;
      int rep = j;
#pragma unroll
      for (int jj = NT - 1; jj >= 0; --jj)
        if (jj < j && ((uint32_t)w->prefix[jj] >> SH) == cmp[j]) rep = jj;
      if (threadIdx.x == 0) rep_s[j] = rep;
      if (rep != j) cmp[j] = 0xffffffffu;
    }
  }
  if (PASS == 1) mean_a = w->mean_all;
  if (PASS >= 3) thr_key = (uint32_t)w->prefix[1];
  if (PASS == 4) { mean_t = w->mean_top; mean_j = w->mean_adj; }
  __syncthreads();

  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  long long c0 = 0, c1 = 0;

  auto process = [&](const float x) {
    const uint32_t k = float_key(x);
    if (PASS == 0) {
      a0 += (double)x;
      atomicAdd(&smem_hist[k >> (32 - L1_BITS)], 1u);
    } else if (PASS == 1) {
      const double d = (double)x - mean_a;
      a0 += fabs(d);
      a1 += d * d;
      const uint32_t top = k >> (32 - L1_BITS);
      const uint32_t mid = (k >> L3_BITS) & (L2_BINS - 1);
#pragma unroll
      for (int j = 0; j < NT; ++j)
        if (top == cmp[j]) atomicAdd(&smem_hist[j * L2_BINS + mid], 1u);
    } else if (PASS == 2) {
      const uint32_t hi = k >> L3_BITS;
      const uint32_t lo = k & (L3_BINS - 1);
#pragma unroll
      for (int j = 0; j < NT; ++j)
        if (hi == cmp[j]) atomicAdd(&smem_hist[j * L3_BINS + lo], 1u);
    } else if (PASS == 3) {
      if (k > thr_key) { a0 += (double)x; ++c0; }
      else if (k < thr_key) { a1 += (double)x; ++c1; }
    } else {
      if (k > thr_key) { const double d = (double)x - mean_t; a0 += fabs(d); a1 += d * d; }
      else if (k < thr_key) { const double d = (double)x - mean_j; a2 += fabs(d); a3 += d * d; }
    }
  };

  // RS_ITEMS elements per thread per tile, loaded as independent 16-byte vectors
  // before any of them is consumed (memory-level parallelism: the passes are
  // HBM streams); rows whose base is not 16-byte aligned take scalar loads.
  constexpr int64_t TILE = (int64_t)RS_THREADS * RS_ITEMS;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(v) & 15) == 0);
  for (int64_t base = (int64_t)blockIdx.x * TILE; base < n; base += (int64_t)gridDim.x * TILE) {
    if (vec_ok && base + TILE <= n) {
      const float4* __restrict__ v4 = reinterpret_cast<const float4*>(v + base);
      float4 t[RS_ITEMS / 4];
#pragma unroll
      for (int u = 0; u < RS_ITEMS / 4; ++u) t[u] = __ldcs(v4 + u * RS_THREADS + threadIdx.x);
#pragma unroll
      for (int u = 0; u < RS_ITEMS / 4; ++u) { process(t[u].x); process(t[u].y); process(t[u].z); process(t[u].w); }
    } else {
      float t[RS_ITEMS];
#pragma unroll
      for (int u = 0; u < RS_ITEMS; ++u) {
        const int64_t i = base + (int64_t)u * RS_THREADS + threadIdx.x;
        t[u] = i < n ? __ldcs(v + i) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < RS_ITEMS; ++u)
        if (base + (int64_t)u * RS_THREADS + threadIdx.x < n) process(t[u]);
    }
  }

  // block-level sums -> one atomic per block per quantity
  if (PASS == 0) {
    double s = block_sum(a0, red_d);
    if (threadIdx.x == 0) atomic_add_f64(&w->sum_all, s);
  } else if (PASS == 1) {
    double s0 = block_sum(a0, red_d), s1 = block_sum(a1, red_d);
    if (threadIdx.x == 0) { atomic_add_f64(&w->absdev_all, s0); atomic_add_f64(&w->sqdev_all, s1); }
  } else if (PASS == 3) {
    double s0 = block_sum(a0, red_d), s1 = block_sum(a1, red_d);
    long long n0 = block_sum(c0, red_i), n1 = block_sum(c1, red_i);
    if (threadIdx.x == 0) {
      atomic_add_f64(&w->sum_gt, s0); atomic_add_f64(&w->sum_lt, s1);
      atomic_add_i64(&w->cnt_gt, n0); atomic_add_i64(&w->cnt_lt, n1);
    }
  } else if (PASS == 4) {
    double s0 = block_sum(a0, red_d), s1 = block_sum(a1, red_d);
    double s2 = block_sum(a2, red_d), s3 = block_sum(a3, red_d);
    if (threadIdx.x == 0) {
      atomic_add_f64(&w->absdev_gt, s0); atomic_add_f64(&w->sqdev_gt, s1);
      atomic_add_f64(&w->absdev_lt, s2); atomic_add_f64(&w->sqdev_lt, s3);
    }
  }

  if (HBINS > 0) {
    __syncthreads();
    long long* gh = PASS == 0 ? w->hist1 : PASS == 1 ? &w->hist2[0][0] : &w->hist3[0][0];
    constexpr int BINS = PASS == 0 ? L1_BINS : PASS == 1 ? L2_BINS : L3_BINS;
    for (int i = threadIdx.x; i < HBINS; i += RS_THREADS) {
      const int j = i / BINS;
      const unsigned int c = PASS == 0 ? smem_hist[i] : smem_hist[rep_s[j] * BINS + (i - j * BINS)];
      if (c) atomic_add_i64(gh + i, (long long)c);
    }
  }
}

// One block per row: walk a histogram to find the bin that holds `rank`.
// blockDim.x == 256: each thread sums a contiguous run of bins, a block-wide
// inclusive scan of the 256 partials (warp shuffles + 8 warp totals) names the
// thread whose run holds the rank, and that thread walks its <= 8 bins.
__device__ void find_bin(const long long* __restrict__ hist, int bins, long long rank, int* bin_out,
                         long long* rem_out, long long* scratch /* >= 8 */) {
  const int per = bins / 256;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  long long local = 0;
  for (int i = 0; i < per; ++i) local += hist[threadIdx.x * per + i];
  long long incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  __syncthreads();  // scratch may still be read by a previous call
  if (lane == 31) scratch[wid] = incl;
  __syncthreads();
  long long before_warp = 0, total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const long long t = scratch[w];
    if (w < wid) before_warp += t;
    total += t;
  }
  incl += before_warp;
  const long long excl = incl - local;
  // the run that holds the rank; a rank beyond the total (cannot happen for
  // consistent counts) is given to the last thread
  const bool owner = (rank >= excl && rank < incl) || (threadIdx.x == 255 && rank >= total);
  if (owner) {
    long long r = rank - excl;
    int b = threadIdx.x * per;
    for (int i = 0; i < per; ++i) {
      const long long c = hist[threadIdx.x * per + i];
      if (r < c || i == per - 1) { b = threadIdx.x * per + i; break; }
      r -= c;
    }
    *bin_out = b;
    *rem_out = r;
  }
  __syncthreads();
}

// STEP 0: after pass 0   STEP 1: after pass 1   STEP 2: after pass 2
// STEP 3: after pass 3   STEP 4: after pass 4 (writes stats)
template <int STEP>
__global__ void __launch_bounds__(256)
rowstats_resolve_kernel(RowWS* __restrict__ ws, int64_t n_total, int64_t top, double* __restrict__ stats) {
  __shared__ long long scratch[256];
  __shared__ int bin_s;
  __shared__ long long rem_s;
  RowWS* w = ws + blockIdx.x;
  const long long n = n_total, K = top;

  if (STEP == 0) {
    const long long ranks[NT] = {(n - 1) / 2, n - K, n - K + (K - 1) / 2, (n - K - 1) / 2};
    for (int j = 0; j < NT; ++j) {
      find_bin(w->hist1, L1_BINS, ranks[j], &bin_s, &rem_s, scratch);
      if (threadIdx.x == 0) {
        w->prefix[j] = (unsigned long long)((uint32_t)bin_s << (32 - L1_BITS));
        w->rank[j] = rem_s;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      w->mean_all = w->sum_all / (double)n;
      // key bins that can only hold non-finite values: 3 = -inf, 2044 = +inf, 2047 = NaN
      const long long ninf = w->hist1[3], pinf = w->hist1[2044], nan = w->hist1[2047];
      const long long hi = pinf + nan;  // sort to the top
      w->nonfinite[0] = (hi + ninf) > 0;
      w->nonfinite[1] = hi > 0 || ninf > n - K;
      w->nonfinite[2] = hi > K || ninf > 0;
      w->has_nan[0] = nan > 0;
      w->has_nan[1] = nan > 0;
      w->has_nan[2] = nan > K;
    }
  } else if (STEP == 1) {
    for (int j = 0; j < NT; ++j) {
      find_bin(w->hist2[j], L2_BINS, w->rank[j], &bin_s, &rem_s, scratch);
      if (threadIdx.x == 0) {
        w->prefix[j] |= (unsigned long long)((uint32_t)bin_s << L3_BITS);
        w->rank[j] = rem_s;
      }
      __syncthreads();
    }
  } else if (STEP == 2) {
    for (int j = 0; j < NT; ++j) {
      find_bin(w->hist3[j], L3_BINS, w->rank[j], &bin_s, &rem_s, scratch);
      if (threadIdx.x == 0) {
        w->prefix[j] |= (unsigned long long)(uint32_t)bin_s;
        w->rank[j] = rem_s;
        w->value[j] = (double)key_float((uint32_t)w->prefix[j]);
      }
      __syncthreads();
    }
  } else if (STEP == 3) {
    if (threadIdx.x == 0) {
      const double thr = w->value[1];
      const long long n_eq = n - w->cnt_gt - w->cnt_lt;
      const long long tt = K - w->cnt_gt;  // ties that belong to the top group
      w->ties_top = tt;
      w->ties_adj = n_eq - tt;
      // a tie share of zero must not contribute 0 * inf
      w->mean_top = (w->sum_gt + (tt > 0 ? (double)tt * thr : 0.0)) / (double)K;
      w->mean_adj = (w->sum_lt + (n_eq - tt > 0 ? (double)(n_eq - tt) * thr : 0.0)) / (double)(n - K);
    }
  } else {
    if (threadIdx.x == 0) {
      const double thr = w->value[1];
      const double dt = thr - w->mean_top, da = thr - w->mean_adj;
      const double tt = (double)w->ties_top, ta = (double)w->ties_adj;
      // a tie share of zero must not contribute inf*0
      const double abs_top = w->absdev_gt + (tt > 0 ? tt * fabs(dt) : 0.0);
      const double sq_top = w->sqdev_gt + (tt > 0 ? tt * dt * dt : 0.0);
      const double abs_adj = w->absdev_lt + (ta > 0 ? ta * fabs(da) : 0.0);
      const double sq_adj = w->sqdev_lt + (ta > 0 ? ta * da * da : 0.0);
      double* s = stats + (int64_t)blockIdx.x * 12;
      s[0] = w->mean_all; s[1] = w->mean_top; s[2] = w->mean_adj;
      s[3] = w->absdev_all / (double)n; s[4] = abs_top / (double)K; s[5] = abs_adj / (double)(n - K);
      s[6] = sqrt(w->sqdev_all / (double)n); s[7] = sqrt(sq_top / (double)K);
      s[8] = sqrt(sq_adj / (double)(n - K));
      s[9] = w->value[0]; s[10] = w->value[2]; s[11] = w->value[3];
      // torch semantics: Welford's std_mean turns any non-finite member into
      // nan mean/std (hence nan MAD); median propagates NaN
      const double qnan = __longlong_as_double(0x7ff8000000000000LL);
      for (int j = 0; j < 3; ++j) {
        if (w->nonfinite[j]) { s[0 + j] = qnan; s[3 + j] = qnan; s[6 + j] = qnan; }
        if (w->has_nan[j]) s[9 + j] = qnan;
      }
    }
  }
}

static int launch_pass(int pass, const float* values, int64_t rows, int64_t n, int64_t ld, RowWS* ws,
                       cudaStream_t st) {
  if (n <= 0 || rows <= 0) return 0;
  // enough blocks to fill the machine several times over, bounded per row
  const int sms = sm_count();
  int64_t want = ((int64_t)sms * 8 + rows - 1) / rows;
  int64_t max_slices = (n + (int64_t)RS_THREADS * RS_ITEMS - 1) / ((int64_t)RS_THREADS * RS_ITEMS);
  int64_t slices = want < 1 ? 1 : want;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  if (rows > 65535) return set_error(B200_ELIMIT, "rowstats: rows=%lld > 65535 per call", (long long)rows);
  dim3 grid((unsigned)slices, (unsigned)rows);
  switch (pass) {
    case 0: rowstats_pass_kernel<0><<<grid, RS_THREADS, L1_BINS * 4, st>>>(values, n, ld, ws); break;
    case 1: rowstats_pass_kernel<1><<<grid, RS_THREADS, NT * L2_BINS * 4, st>>>(values, n, ld, ws); break;
    case 2: rowstats_pass_kernel<2><<<grid, RS_THREADS, NT * L3_BINS * 4, st>>>(values, n, ld, ws); break;
    case 3: rowstats_pass_kernel<3><<<grid, RS_THREADS, 0, st>>>(values, n, ld, ws); break;
    case 4: rowstats_pass_kernel<4><<<grid, RS_THREADS, 0, st>>>(values, n, ld, ws); break;
  }
  return check_cuda(cudaGetLastError(), "rowstats pass launch");
}

static int launch_resolve(int step, int64_t rows, int64_t n_total, int64_t top, RowWS* ws, double* stats,
                          cudaStream_t st) {
  switch (step) {
    case 0: rowstats_resolve_kernel<0><<<(unsigned)rows, 256, 0, st>>>(ws, n_total, top, stats); break;
    case 1: rowstats_resolve_kernel<1><<<(unsigned)rows, 256, 0, st>>>(ws, n_total, top, stats); break;
    case 2: rowstats_resolve_kernel<2><<<(unsigned)rows, 256, 0, st>>>(ws, n_total, top, stats); break;
    case 3: rowstats_resolve_kernel<3><<<(unsigned)rows, 256, 0, st>>>(ws, n_total, top, stats); break;
    case 4: rowstats_resolve_kernel<4><<<(unsigned)rows, 256, 0, st>>>(ws, n_total, top, stats); break;
  }
  return check_cuda(cudaGetLastError(), "rowstats resolve launch");
}

}  // namespace b200

using namespace b200;

extern "C" int64_t b200_rowstats_workspace_bytes(int64_t rows) {
  return rows < 0 ? 0 : rows * (int64_t)sizeof(RowWS);
}

// Phases for a multi-GPU caller:
//   phase 0: clear + pass 0                      -> exchange D[0..1) and hist1
//   phase 1: resolve 0 + pass 1                  -> exchange D[1..3) and hist2
//   phase 2: resolve 1 + pass 2                  -> exchange hist3
//   phase 3: resolve 2 + pass 3                  -> exchange D[3..5) and cnt
//   phase 4: resolve 3 + pass 4                  -> exchange D[5..9)
//   phase 5: resolve 4 (writes stats)
// The workspace is laid out row-major, so an exchange region is strided by
// ROW_WORDS; callers reduce the whole typed slab view instead (see
// b200_rowstats_exchange): the slabs below are contiguous PER ROW only, hence
// the multi-GPU path reduces the full workspace viewed as int64 for integer
// regions and as float64 for the 16 leading doubles of every row.
extern "C" int b200_rowstats(const float* values, int64_t rows, int64_t n, int64_t ld, int64_t n_total,
                             int64_t top, void* workspace, double* stats, int32_t phase, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_REQUIRE(rows >= 0 && n >= 0, "rowstats: negative size");
  if (rows == 0) return 0;
  B200_REQUIRE(values != nullptr || n == 0, "rowstats: values is NULL");
  B200_REQUIRE(workspace != nullptr && stats != nullptr, "rowstats: workspace/stats is NULL");
  B200_REQUIRE(n_total >= 2 && n <= n_total, "rowstats: need n_total >= 2 and n <= n_total (n=%lld n_total=%lld)",
               (long long)n, (long long)n_total);
  B200_REQUIRE(top >= 1 && top < n_total, "rowstats: need 1 <= top < n_total (top=%lld)", (long long)top);
  B200_REQUIRE(ld >= n, "rowstats: ld < n");
  B200_REQUIRE(phase >= -1 && phase <= 5, "rowstats: phase out of range");
  RowWS* ws = (RowWS*)workspace;
  const int first = phase < 0 ? 0 : phase, last = phase < 0 ? 5 : phase;
  for (int p = first; p <= last; ++p) {
    int rc = 0;
    if (p == 0) {
      B200_CUDA(cudaMemsetAsync(ws, 0, (size_t)rows * sizeof(RowWS), st));
    } else {
      rc = launch_resolve(p - 1, rows, n_total, top, ws, stats, st);
      if (rc) return rc;
    }
    if (p <= 4) {
      rc = launch_pass(p, values, rows, n, ld, ws, st);
      if (rc) return rc;
    }
  }
  return 0;
}

// Exchange description: per row, which 8-byte words a rank must sum with its
// peers after `phase`.
extern "C" int b200_rowstats_exchange(int32_t phase, int64_t out[5]) {
  B200_REQUIRE(out != nullptr, "rowstats_exchange: NULL output");
  int64_t io = 0, ic = 0, d_o = 0, dc = 0;
  switch (phase) {
    case 0: io = OFF_H1; ic = L1_BINS; d_o = 0; dc = 1; break;
    case 1: io = OFF_H2; ic = (int64_t)NT * L2_BINS; d_o = 1; dc = 2; break;
    case 2: io = OFF_H3; ic = (int64_t)NT * L3_BINS; break;
    case 3: io = OFF_CNT; ic = 2; d_o = 3; dc = 2; break;
    case 4: d_o = 5; dc = 4; break;
    default: break;
  }
  out[0] = io; out[1] = ic; out[2] = d_o; out[3] = dc; out[4] = ROW_WORDS;
  return 0;
}
