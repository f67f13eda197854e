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
// Four passes over the data, each an HBM stream (4 B per element per pass):
//   pass 0: sum, level-1 histogram
//   pass 1: level-2 histograms
//   pass 2: level-3 histograms + sums/counts of the elements whose 22-bit prefix is
//           above / below thr's (the elements that share thr's prefix all sit in
//           thr's level-3 histogram, and a level-3 bin IS one fp32 value, so the
//           resolve step finishes the top / adj split exactly from the bin counts)
//   pass 3: |w-mean|, (w-mean)^2 of all, and the deviations of top / adj about
//           their own means (pass 3 has the fewest instructions per element left:
//           the all-group moments ride there, not on the histogram passes)
// Between passes a one-block-per-row "resolve" kernel turns histograms into
// prefixes/ranks.  Every cross-block quantity lives in the caller's workspace as
// 8-byte words so that a multi-GPU caller can all-reduce it between passes.
#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace b200 {

constexpr int L1_BITS = 11, L2_BITS = 11, L3_BITS = 10;
constexpr int L1_BINS = 1 << L1_BITS, L2_BINS = 1 << L2_BITS, L3_BINS = 1 << L3_BITS;
constexpr int NT = 4;  // targets

// Per-row workspace, 8-byte words.  Doubles first, then integers.
struct RowWS {
  // ---- doubles (exchange region D) -------------------------------------
  double sum_all;      // pass 0
  double sum_gt;       // pass 2: sum of w whose 22-bit key prefix is above thr's
  double sum_lt;       // pass 2: ... below thr's
  double absdev_all;   // pass 3 (about mean_all)
  double sqdev_all;    // pass 3
  double absdev_gt;    // pass 3 (about mean_top), over key > key(thr)
  double sqdev_gt;
  double absdev_lt;    // pass 3 (about mean_adj), over key < key(thr)
  double sqdev_lt;
  double pad_d[7];
  // ---- integers (exchange region I) ------------------------------------
  long long hist1[L1_BINS];          // pass 0
  long long hist2[NT][L2_BINS];      // pass 1
  long long cnt_gt;                  // pass 2 (the count that goes with sum_gt)
  long long pad_i[7];
  long long hist3[NT][L3_BINS];      // pass 2
  // ---- resolved by the row-resolve kernels (identical on every rank) ---
  long long rank[NT];       // remaining rank inside the current prefix
  unsigned long long prefix[NT];  // key prefix found so far (left aligned)
  double mean_all, mean_top, mean_adj;
  double value[NT];         // selected values (as double)
  long long ties_top, ties_adj;
  long long nonfinite[3];   // all / top / adj group holds +-inf or NaN (torch.std_mean -> nan)
  long long has_nan[3];     // all / top / adj group holds a NaN (torch.median -> nan)
  long long key_mode;       // KEY_FULL / KEY_NO_NAN / KEY_POSITIVE for passes 1..3 (from hist1)
  long long pad_r[2];
  double stats_out[12];     // multi-GPU: the row's statistics as its owner rank resolved them (step 3)
};
static_assert(sizeof(RowWS) % 16 == 0, "8-byte words, rows 16-byte aligned (the peer gather loads word pairs)");
static_assert(offsetof(RowWS, hist1) == 16 * 8 && offsetof(RowWS, cnt_gt) == (16 + L1_BINS + NT * L2_BINS) * 8 &&
                  offsetof(RowWS, hist3) == offsetof(RowWS, cnt_gt) + 64,
              "exchange offsets follow the struct");

constexpr int64_t D_WORDS = 16;  // doubles at the head
constexpr int64_t OFF_H1 = D_WORDS;
constexpr int64_t OFF_H2 = OFF_H1 + L1_BINS;
constexpr int64_t OFF_CNT = OFF_H2 + (int64_t)NT * L2_BINS;
constexpr int64_t OFF_H3 = OFF_CNT + 8;
constexpr int64_t OFF_RES = OFF_H3 + (int64_t)NT * L3_BINS;
constexpr int64_t ROW_WORDS = sizeof(RowWS) / 8;

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;  // elements per thread per block-slice iteration

// Key modes.  Pass 0 must use the full map (float_key of common.cuh: sign-dependent
// flip, every NaN to the top: four instructions).  Its level-1 histogram tells
// whether the row holds a NaN or anything with the sign bit set; the later passes
// of a NaN-free row skip the NaN test (two instructions), and those of an
// all-positive row - every wealth row - order by the raw bit pattern: there
// key = bits | 0x80000000, so the kernel compares raw bits against constants
// with the top bit removed and spends nothing on the key.
enum { KEY_FULL = 0, KEY_NO_NAN = 1, KEY_POSITIVE = 2 };
template <int MODE>
__device__ __forceinline__ uint32_t float_key_fast(float f) {
  const uint32_t b = __float_as_uint(f);
  if (MODE == KEY_POSITIVE) return b;
  const uint32_t k = b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
  return (MODE == KEY_FULL && f != f) ? 0xffffffffu : k;
}
// a key-space constant (shifted right by SH) in the units float_key_fast<MODE> returns
template <int MODE>
__device__ __forceinline__ uint32_t key_const(uint32_t v, int sh) {
  return MODE == KEY_POSITIVE ? v ^ (0x80000000u >> sh) : v;
}

__device__ __forceinline__ void atomic_add_f64(double* p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add_i64(long long* p, long long v) {
  atomicAdd((unsigned long long*)p, (unsigned long long)v);
}

// One block walks its slice of one row.  Per-element work is kept to a dozen
// instructions so that every pass stays an HBM stream: the ALU pipe issues 2 warp
// instructions per clock per SM, i.e. ~11 ALU instructions per element at the
// measured HBM rate.
template <int PASS, int MODE>
__device__ __forceinline__ void rowstats_pass_body(const float* __restrict__ v, int64_t n, RowWS* __restrict__ w,
                                                   unsigned int* smem_hist) {
  __shared__ double red_d[32];
  __shared__ long long red_i[32];

  constexpr int HBINS = PASS == 0 ? L1_BINS : PASS == 1 ? NT * L2_BINS : PASS == 2 ? NT * L3_BINS : 0;
  for (int i = threadIdx.x; i < HBINS; i += RS_THREADS) smem_hist[i] = 0;
  // pass 1: level-1 bin -> byte offset of the level-2 histogram that counts it (-1: none)
  int* lut = reinterpret_cast<int*>(smem_hist + HBINS);
  if (PASS == 1)
    for (int i = threadIdx.x; i < L1_BINS; i += RS_THREADS) lut[i] = -1;

  // Targets that fell into the same bin so far need the same histogram: only the
  // first of them (its representative) counts, the others copy it at the flush.
  // cmp[j] is the bin prefix a key must show to be counted for target j; a
  // duplicate gets a prefix no key can have, so a key matches at most one target.
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
      cmp[j] = rep != j ? 0xffffffffu : key_const<MODE>(cmp[j], SH);
    }
  }
  if (PASS == 1) {
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int j = 0; j < NT; ++j)
        if (cmp[j] != 0xffffffffu) lut[cmp[j]] = j * L2_BINS * 4;
    }
  }
  if (PASS == 2) thr_key = key_const<MODE>((uint32_t)w->prefix[1] >> L3_BITS, L3_BITS);   // thr's 22-bit prefix
  if (PASS == 3) {
    thr_key = key_const<MODE>((uint32_t)w->prefix[1], 0);
    mean_a = w->mean_all;
    mean_t = w->mean_top;
    mean_j = w->mean_adj;
  }
  __syncthreads();

  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;
  unsigned int c0 = 0;  // per-thread count: a thread sees far fewer than 2^32 elements

  auto process = [&](const float x) {
    const uint32_t k = float_key_fast<MODE>(x);
    if (PASS == 0) {
      a0 += (double)x;
      atomicAdd(&smem_hist[k >> (32 - L1_BITS)], 1u);
    } else if (PASS == 1) {
      const int off = lut[k >> (32 - L1_BITS)];
      if (off >= 0)
        atomicAdd(reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(smem_hist) + off +
                                                  ((k >> (L3_BITS - 2)) & ((L2_BINS - 1) << 2))), 1u);
    } else if (PASS == 2) {
      // Sums on either side of thr's 22-bit prefix (the count below it follows
      // from the total at the resolve step).  Nearly every element lies below
      // thr (the top group is small) and outside the targets' 22-bit prefixes:
      // its whole cost is the key, three compares and one predicated fp64 add;
      // everything else sits behind one rarely taken branch.  (med_top's prefix
      // is never below thr's, so "not below thr" covers targets 1 and 2.)
      const uint32_t hi = k >> L3_BITS;
      a1 += (double)(hi < thr_key ? x : 0.0f);   // one fp32 select, not two on the fp64 sum
      if ((hi >= thr_key) | (hi == cmp[0]) | (hi == cmp[3])) {
        if (hi > thr_key) { a0 += (double)x; ++c0; }
        if ((hi == cmp[0]) | (hi == cmp[1]) | (hi == cmp[2]) | (hi == cmp[3])) {
          const uint32_t j = hi == cmp[0] ? 0u : hi == cmp[1] ? 1u : hi == cmp[2] ? 2u : 3u;
          atomicAdd(&smem_hist[j * L3_BINS + (k & (L3_BINS - 1))], 1u);
        }
      }
    } else {
      const double xd = (double)x;
      const double da = xd - mean_a;
      a4 += fabs(da);
      a5 = __fma_rn(da, da, a5);
      if (k < thr_key) { const double d = xd - mean_j; a2 += fabs(d); a3 = __fma_rn(d, d, a3); }
      else if (k > thr_key) { const double d = xd - mean_t; a0 += fabs(d); a1 = __fma_rn(d, d, a1); }
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
      if (PASS == 1) {
        // eight table lookups are issued before the first (dependent) histogram branch
        const float* e = reinterpret_cast<const float*>(t);
#pragma unroll
        for (int h = 0; h < RS_ITEMS; h += 8) {
          int off[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) off[u] = lut[float_key_fast<MODE>(e[h + u]) >> (32 - L1_BITS)];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (off[u] >= 0) {
              const uint32_t k = float_key_fast<MODE>(e[h + u]);
              atomicAdd(reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(smem_hist) + off[u] +
                                                        ((k >> (L3_BITS - 2)) & ((L2_BINS - 1) << 2))), 1u);
            }
          }
        }
      } else {
#pragma unroll
        for (int u = 0; u < RS_ITEMS / 4; ++u) { process(t[u].x); process(t[u].y); process(t[u].z); process(t[u].w); }
      }
    } else {
      for (int u = 0; u < RS_ITEMS; ++u) {
        const int64_t i = base + (int64_t)u * RS_THREADS + threadIdx.x;
        if (i < n) process(__ldcs(v + i));
      }
    }
  }

  // block-level sums -> one atomic per block per quantity
  if (PASS == 0) {
    double s = block_sum(a0, red_d);
    if (threadIdx.x == 0) atomic_add_f64(&w->sum_all, s);
  } else if (PASS == 2) {
    double s0 = block_sum(a0, red_d), s1 = block_sum(a1, red_d);
    long long n0 = block_sum((long long)c0, red_i);
    if (threadIdx.x == 0) {
      if (n0) { atomic_add_f64(&w->sum_gt, s0); atomic_add_i64(&w->cnt_gt, n0); }
      atomic_add_f64(&w->sum_lt, s1);
    }
  } else if (PASS == 3) {
    double s0 = block_sum(a0, red_d), s1 = block_sum(a1, red_d);
    double s2 = block_sum(a2, red_d), s3 = block_sum(a3, red_d);
    double s4 = block_sum(a4, red_d), s5 = block_sum(a5, red_d);
    if (threadIdx.x == 0) {
      atomic_add_f64(&w->absdev_gt, s0); atomic_add_f64(&w->sqdev_gt, s1);
      atomic_add_f64(&w->absdev_lt, s2); atomic_add_f64(&w->sqdev_lt, s3);
      atomic_add_f64(&w->absdev_all, s4); atomic_add_f64(&w->sqdev_all, s5);
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

// grid = (slices, rows)
template <int PASS>
__global__ void __launch_bounds__(RS_THREADS)
rowstats_pass_kernel(const float* __restrict__ values, int64_t n, int64_t ld, RowWS* __restrict__ ws) {
  extern __shared__ unsigned int smem_hist[];
  const int64_t row = blockIdx.y;
  const float* __restrict__ v = values + row * ld;
  RowWS* w = ws + row;
  const long long mode = PASS == 0 ? (long long)KEY_FULL : w->key_mode;   // block-uniform
  if (mode == KEY_POSITIVE) rowstats_pass_body<PASS, KEY_POSITIVE>(v, n, w, smem_hist);
  else if (mode == KEY_NO_NAN) rowstats_pass_body<PASS, KEY_NO_NAN>(v, n, w, smem_hist);
  else rowstats_pass_body<PASS, KEY_FULL>(v, n, w, smem_hist);
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

// A step's histogram(s) from the workspace (L2) into shared memory: every thread's
// 16-byte loads are issued before the first store (a rolled loop would pay the L2
// latency once per iteration: 8 dependent round trips for the 64 KB of step 1).
template <int COUNT>
__device__ __forceinline__ void copy_hist(long long* __restrict__ dst, const long long* __restrict__ src) {
  static_assert(COUNT % 512 == 0, "256 threads x word pairs");
  constexpr int PER = COUNT / 512;
  const longlong2* __restrict__ s2 = reinterpret_cast<const longlong2*>(src);
  longlong2* __restrict__ d2 = reinterpret_cast<longlong2*>(dst);
  longlong2 v[PER];
#pragma unroll
  for (int u = 0; u < PER; ++u) v[u] = s2[u * 256 + threadIdx.x];
#pragma unroll
  for (int u = 0; u < PER; ++u) d2[u * 256 + threadIdx.x] = v[u];
  __syncthreads();
}

// STEP 0: after pass 0   STEP 1: after pass 1   STEP 2: after pass 2 (order
// statistics, top / adj split and means)   STEP 3: after pass 3 (writes the 12 statistics)
// One block (256 threads) per row.  `hs`: the step's histogram(s), summed over the
// world, in shared memory (STEP 0: hist1, STEP 1: hist2[NT], STEP 2: hist3[NT]);
// `sd`: the row's 16 leading doubles, summed over the world; `cnt_gt_sum`: cnt_gt
// likewise (STEP 2).  The resolved prefixes / ranks / means go to `w`.
template <int STEP>
__device__ void resolve_row(const long long* __restrict__ hs, const double* __restrict__ sd, long long cnt_gt_sum,
                            RowWS* __restrict__ w, long long n, long long K, double* __restrict__ stats_row) {
  __shared__ long long scratch[256];
  __shared__ double red_d[32];
  __shared__ long long red_i[32];
  __shared__ int bin_s;
  __shared__ long long rem_s;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  if (STEP == 0) {
    const long long ranks[NT] = {(n - 1) / 2, n - K, n - K + (K - 1) / 2, (n - K - 1) / 2};
    for (int j = 0; j < NT; ++j) {
      find_bin(hs, L1_BINS, ranks[j], &bin_s, &rem_s, scratch);
      if (threadIdx.x == 0) {
        w->prefix[j] = (unsigned long long)((uint32_t)bin_s << (32 - L1_BITS));
        w->rank[j] = rem_s;
      }
      __syncthreads();
    }
    // anything with the sign bit set (level-1 bins 0..1023)?  4 bins per thread
    int neg = 0;
    for (int i = 0; i < 4; ++i) neg |= hs[threadIdx.x * 4 + i] != 0;
    neg = __syncthreads_or(neg);
    if (threadIdx.x == 0) {
      w->key_mode = hs[2047] > 0 ? KEY_FULL : neg ? KEY_NO_NAN : KEY_POSITIVE;
      w->mean_all = sd[0] / (double)n;
      // key bins that can only hold non-finite values: 3 = -inf, 2044 = +inf, 2047 = NaN
      const long long ninf = hs[3], pinf = hs[2044], nan = hs[2047];
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
      find_bin(hs + j * L2_BINS, L2_BINS, w->rank[j], &bin_s, &rem_s, scratch);
      if (threadIdx.x == 0) {
        w->prefix[j] |= (unsigned long long)((uint32_t)bin_s << L3_BITS);
        w->rank[j] = rem_s;
      }
      __syncthreads();
    }
  } else if (STEP == 2) {
    for (int j = 0; j < NT; ++j) {
      find_bin(hs + j * L3_BINS, L3_BINS, w->rank[j], &bin_s, &rem_s, scratch);
      if (threadIdx.x == 0) {
        w->prefix[j] |= (unsigned long long)(uint32_t)bin_s;
        w->rank[j] = rem_s;
        w->value[j] = (double)key_float((uint32_t)w->prefix[j]);
      }
      __syncthreads();
    }
    // The elements that share thr's 22-bit prefix: bin `lo` of thr's level-3
    // histogram holds c copies of the one fp32 value with key (prefix22 | lo).
    const uint32_t thr_key = (uint32_t)w->prefix[1];
    const uint32_t thr_lo = thr_key & (L3_BINS - 1), base = thr_key & ~(uint32_t)(L3_BINS - 1);
    double s_gt = 0, s_lt = 0;
    long long c_gt = 0, c_lt = 0, c_all = 0;
    for (uint32_t lo = threadIdx.x; lo < (uint32_t)L3_BINS; lo += blockDim.x) {
      const long long c = hs[L3_BINS + lo];
      c_all += c;
      if (c == 0 || lo == thr_lo) continue;   // an empty bin must not contribute 0 * inf
      const double v = (double)c * (double)key_float(base | lo);
      if (lo > thr_lo) { s_gt += v; c_gt += c; } else { s_lt += v; c_lt += c; }
    }
    s_gt = block_sum(s_gt, red_d); s_lt = block_sum(s_lt, red_d);
    c_gt = block_sum(c_gt, red_i); c_lt = block_sum(c_lt, red_i); c_all = block_sum(c_all, red_i);
    if (threadIdx.x == 0) {
      const double thr = w->value[1];
      const long long cnt_gt = cnt_gt_sum;
      const double coarse_gt = sd[1], coarse_lt = sd[2];
      // below thr's prefix = everything that is neither above it nor inside it
      const long long n_gt = cnt_gt + c_gt, n_lt = (n - cnt_gt - c_all) + c_lt;
      const double sum_gt = c_gt ? coarse_gt + s_gt : coarse_gt, sum_lt = c_lt ? coarse_lt + s_lt : coarse_lt;
      const long long n_eq = n - n_gt - n_lt;
      const long long tt = K - n_gt;  // ties that belong to the top group
      w->ties_top = tt;
      w->ties_adj = n_eq - tt;
      // a tie share of zero must not contribute 0 * inf
      w->mean_top = (sum_gt + (tt > 0 ? (double)tt * thr : 0.0)) / (double)K;
      w->mean_adj = (sum_lt + (n_eq - tt > 0 ? (double)(n_eq - tt) * thr : 0.0)) / (double)(n - K);
    }
  } else {
    if (threadIdx.x == 0) {
      const double absdev_all = sd[3], sqdev_all = sd[4];
      const double absdev_gt = sd[5], sqdev_gt = sd[6], absdev_lt = sd[7], sqdev_lt = sd[8];
      const double thr = w->value[1];
      const double dt = thr - w->mean_top, da = thr - w->mean_adj;
      const double tt = (double)w->ties_top, ta = (double)w->ties_adj;
      // a tie share of zero must not contribute inf*0
      const double abs_top = absdev_gt + (tt > 0 ? tt * fabs(dt) : 0.0);
      const double sq_top = sqdev_gt + (tt > 0 ? tt * dt * dt : 0.0);
      const double abs_adj = absdev_lt + (ta > 0 ? ta * fabs(da) : 0.0);
      const double sq_adj = sqdev_lt + (ta > 0 ? ta * da * da : 0.0);
      double* s = stats_row;
      s[0] = w->mean_all; s[1] = w->mean_top; s[2] = w->mean_adj;
      s[3] = absdev_all / (double)n; s[4] = abs_top / (double)K; s[5] = abs_adj / (double)(n - K);
      s[6] = sqrt(sqdev_all / (double)n); s[7] = sqrt(sq_top / (double)K);
      s[8] = sqrt(sq_adj / (double)(n - K));
      s[9] = w->value[0]; s[10] = w->value[2]; s[11] = w->value[3];
      // torch semantics: Welford's std_mean turns any non-finite member into
      // nan mean/std (hence nan MAD); median propagates NaN
      // ... except that a ONE-element group's mean is the element itself (Welford's
      // first update is exact: mean = +-inf, then std = MAD = nan)
      const long long size[3] = {n, K, n - K};
      const double single[3] = {qnan, w->value[1], w->value[3]};   // top = {thr}, adj = {its median}
      for (int j = 0; j < 3; ++j) {
        if (w->nonfinite[j]) { s[0 + j] = size[j] == 1 ? single[j] : qnan; s[3 + j] = qnan; s[6 + j] = qnan; }
        if (w->has_nan[j]) s[9 + j] = qnan;
      }
    }
  }
}

// One GPU (or partial sums all-reduced in place by the caller): grid = rows.
template <int STEP>
__global__ void __launch_bounds__(256)
rowstats_resolve_kernel(RowWS* __restrict__ ws, int64_t n_total, int64_t top, double* __restrict__ stats) {
  extern __shared__ long long hs[];
  const int64_t row = blockIdx.x;
  RowWS* w = ws + row;
  const long long* si = reinterpret_cast<const long long*>(w);
  if (STEP == 0) copy_hist<L1_BINS>(hs, si + OFF_H1);
  if (STEP == 1) copy_hist<NT * L2_BINS>(hs, si + OFF_H2);
  if (STEP == 2) copy_hist<NT * L3_BINS>(hs, si + OFF_H3);
  resolve_row<STEP>(hs, reinterpret_cast<const double*>(w), si[OFF_CNT], w, n_total, top, stats + row * 12);
}

// ----------------------------------------------------------------------------
// Multi-GPU without a collective library: every rank's workspace is mapped into
// every process (CUDA IPC / symmetric memory).  Rows are OWNED round-robin
// (row % world).  After each pass a rank PUSHES the partial sums of every row to
// the staging area of the row's owner (posted stores: NVLink at line rate, where
// loads out of a peer's memory were bound by their round trips - 0.25 TB/s
// measured) and its last block raises a flag at every peer; the owner then sums
// its rows' partials out of its own memory into shared memory, resolves them there
// and stores the few resolved words (prefixes, ranks, means, finally the 12
// statistics) into every rank's workspace.  So a histogram crosses NVLink once (to
// its owner) instead of once per rank, the resolve work is split over the ranks,
// and every rank continues from the same resolved words: bit-identical statistics
// everywhere.  Doubles are added in rank order.
struct PeerSet {
  RowWS* ws[B200_MAX_PEERS];            // [rank] is this rank's own workspace
  uint32_t* flags[B200_MAX_PEERS];      // flag blocks (B200_PEER_FLAG_WORDS uint32 each)
  long long* stage[B200_MAX_PEERS];     // staging of every rank: [source rank][owned row][STAGE_WORDS]
  int64_t stage_rows;                   // owned rows the staging holds per source rank
  int32_t world, rank;
  uint32_t epoch;
};
// A staged row (8-byte words): the 16 leading doubles of the source's RowWS (word 15 carries cnt_gt), then the
// step's histogram(s) as 32-bit counts.
constexpr int64_t STAGE_HEAD = 16, STAGE_CNT = 15, STAGE_WORDS = STAGE_HEAD + (int64_t)NT * L2_BINS / 2;
// Flag block of a rank: for every peer r, words [r][0..3] "r's partial sums of pass p
// are in memory" and [r][4..7] "the rows r owns are resolved after pass p and stored
// here"; then the error word, then one block ticket per pass for the push kernels and one for the owner kernels.
constexpr int FLAG_STRIDE = 8, FLAG_RESOLVED = 4;
constexpr int FLAG_ERROR_WORD = B200_PEER_FLAG_ERROR_WORD, FLAG_TICKET = FLAG_ERROR_WORD + 1;
static_assert(FLAG_ERROR_WORD == B200_MAX_PEERS * FLAG_STRIDE && FLAG_TICKET + 8 <= B200_PEER_FLAG_WORDS, "flag block layout");

// Warp 0 of a block (all 32 lanes call): raises flag `word` of this rank at every peer.
__device__ __forceinline__ void signal_peers(const PeerSet& P, int word) {
  const int r = threadIdx.x;
  if (r < P.world && r != P.rank) {
    __threadfence_system();
    uint32_t* f = P.flags[r] + P.rank * FLAG_STRIDE + word;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(P.epoch) : "memory");
  }
}

// Block-wide wait for flag `word` of every peer, one lane per peer; false after
// ~60 s (a peer died) or when an earlier wait already gave up: the error word is
// set and the call's statistics become NaN instead of the device hanging.
__device__ bool wait_peers(const PeerSet& P, int word) {
  __shared__ int ok_s;
  if (threadIdx.x < 32) {
    int ok = 1;
    const int r = threadIdx.x;
    volatile uint32_t* err = P.flags[P.rank] + FLAG_ERROR_WORD;
    if (r < P.world && r != P.rank) {
      const uint32_t* f = P.flags[P.rank] + r * FLAG_STRIDE + word;
      unsigned long long t0;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if ((int32_t)(v - P.epoch) >= 0) break;
        if (*err) { ok = 0; break; }
        __nanosleep(100);
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 60000000000ull) { ok = 0; break; }
      }
    }
    ok = __all_sync(0xffffffffu, ok);
    if (threadIdx.x == 0) {
      if (!ok) *err = 1u;
      ok_s = ok;
    }
  }
  __syncthreads();
  return ok_s != 0;
}

// Where rank r's partial sums of owned row `row` are on THIS (the owner's) GPU: the own
// workspace row, or the staged copy the source pushed.  `head`: the doubles; + hist: the histogram words.
struct RowSource {
  const long long* head;
  const long long* hist;
};
template <int STEP>
__device__ __forceinline__ RowSource row_source(const PeerSet& P, int r, int64_t row) {
  constexpr int64_t IOFF = STEP == 0 ? OFF_H1 : STEP == 1 ? OFF_H2 : OFF_H3;
  RowSource s;
  if (r == P.rank) {
    s.head = reinterpret_cast<const long long*>(P.ws[P.rank] + row);
    s.hist = s.head + IOFF;
  } else {
    s.head = P.stage[P.rank] + ((int64_t)r * P.stage_rows + row / P.world) * STAGE_WORDS;
    s.hist = s.head + STAGE_HEAD;
  }
  return s;
}

// After pass STEP, before the owners resolve: every row's partial sums to the row's
// owner.  Blocks stride over the rows; the last block to finish raises this rank's
// flag at every peer (one system-scope fence per block).
template <int STEP>
__global__ void __launch_bounds__(256)
rowstats_push_kernel(const __grid_constant__ PeerSet P, int64_t rows) {
  constexpr int64_t IOFF = STEP == 0 ? OFF_H1 : STEP == 1 ? OFF_H2 : OFF_H3;
  constexpr int ICNT = STEP == 0 ? L1_BINS : STEP == 1 ? NT * L2_BINS : STEP == 2 ? NT * L3_BINS : 0;
  constexpr int DOFF = STEP == 0 ? 0 : STEP == 2 ? 1 : 3;
  constexpr int DCNT = STEP == 0 ? 1 : STEP == 1 ? 0 : STEP == 2 ? 2 : 6;
  constexpr int PER4 = ICNT / 1024;    // groups of four counts per thread and row
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const int owner = (int)(row % P.world);
    if (owner == P.rank) continue;
    const long long* src = reinterpret_cast<const long long*>(P.ws[P.rank] + row);
    long long* dst = P.stage[owner] + ((int64_t)P.rank * P.stage_rows + row / P.world) * STAGE_WORDS;
    if (PER4 > 0) {   // the counts travel as 32-bit words (a rank holds fewer than 2^32 elements): half the NVLink bytes
      const longlong2* s2 = reinterpret_cast<const longlong2*>(src + IOFF);
      uint4* d4 = reinterpret_cast<uint4*>(dst + STAGE_HEAD);
      constexpr int CH = PER4 > 4 ? 4 : (PER4 > 0 ? PER4 : 1);
#pragma unroll 1
      for (int u0 = 0; u0 < PER4; u0 += CH) {
        longlong2 v[CH][2];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          v[c][0] = __ldcg(s2 + 2 * ((u0 + c) * 256 + threadIdx.x));
          v[c][1] = __ldcg(s2 + 2 * ((u0 + c) * 256 + threadIdx.x) + 1);
        }
#pragma unroll
        for (int c = 0; c < CH; ++c)
          d4[(u0 + c) * 256 + threadIdx.x] = make_uint4((uint32_t)v[c][0].x, (uint32_t)v[c][0].y, (uint32_t)v[c][1].x,
                                                        (uint32_t)v[c][1].y);
      }
    }
    if ((int)threadIdx.x < DCNT) dst[DOFF + threadIdx.x] = src[DOFF + threadIdx.x];
    if (STEP == 2 && threadIdx.x == 32) dst[STAGE_CNT] = src[OFF_CNT];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    int last = 0;
    if (threadIdx.x == 0) {
      unsigned int* ticket = P.flags[P.rank] + FLAG_TICKET + STEP;
      __threadfence_system();
      last = atomicAdd(ticket, 1u) == gridDim.x - 1;
      if (last) *ticket = 0u;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) signal_peers(P, STEP);
  }
}

// COUNT 8-byte words of the step's histogram(s) of owned row `row`, summed over the
// world (own workspace + the staged copies, all in this GPU's memory), into shared memory.
template <int STEP, int COUNT>
__device__ __forceinline__ void gather_hist(long long* __restrict__ hs, const PeerSet& P, int64_t row) {
  static_assert(COUNT % 1024 == 0, "256 threads x four counts");
  constexpr int PER4 = COUNT / 1024;
  const uint4* staged[B200_MAX_PEERS];    // the peers' counts: 32-bit words ([rank]: unused)
#pragma unroll
  for (int r = 0; r < B200_MAX_PEERS; ++r)
    staged[r] = reinterpret_cast<const uint4*>(row_source<STEP>(P, r < P.world && r != P.rank ? r : (P.rank + 1) % P.world, row).hist);
  const longlong2* own = reinterpret_cast<const longlong2*>(row_source<STEP>(P, P.rank, row).hist);
#pragma unroll 1
  for (int u = 0; u < PER4; ++u) {
    const int i = u * 256 + threadIdx.x;
    uint4 v[B200_MAX_PEERS];
    const longlong2 o0 = __ldcg(own + 2 * i), o1 = __ldcg(own + 2 * i + 1);
#pragma unroll
    for (int r = 0; r < B200_MAX_PEERS; ++r) v[r] = __ldcg(staged[r] + i);   // absent ranks: a cached re-read, not added
    long long a = o0.x, b = o0.y, c = o1.x, d = o1.y;
#pragma unroll
    for (int r = 0; r < B200_MAX_PEERS; ++r)
      if (r < P.world && r != P.rank) { a += v[r].x; b += v[r].y; c += v[r].z; d += v[r].w; }
    reinterpret_cast<longlong2*>(hs)[2 * i] = make_longlong2(a, b);
    reinterpret_cast<longlong2*>(hs)[2 * i + 1] = make_longlong2(c, d);
  }
}

// After the push of step STEP: grid = the rows this rank owns (at least one block:
// a rank without rows still tells its peers it is done), 256 threads.
template <int STEP>
__global__ void __launch_bounds__(256)
rowstats_owner_kernel(const __grid_constant__ PeerSet P, int64_t rows, int64_t n_total, int64_t top) {
  extern __shared__ long long hs[];
  __shared__ double sd_s[16];
  __shared__ long long cnt_s;
  constexpr int DOFF = STEP == 0 ? 0 : STEP == 2 ? 1 : 3;
  constexpr int DCNT = STEP == 0 ? 1 : STEP == 1 ? 0 : STEP == 2 ? 2 : 6;
  const bool ok = wait_peers(P, STEP);
  const int64_t row = P.rank + (int64_t)blockIdx.x * P.world;
  if (ok && row < rows) {
    if (STEP == 0) gather_hist<0, L1_BINS>(hs, P, row);
    if (STEP == 1) gather_hist<1, NT * L2_BINS>(hs, P, row);
    if (STEP == 2) gather_hist<2, NT * L3_BINS>(hs, P, row);
    if ((int)threadIdx.x < DCNT) {
      double acc = 0.0;
      for (int r = 0; r < P.world; ++r) {
        const double v = __longlong_as_double(__ldcg(row_source<STEP>(P, r, row).head + DOFF + threadIdx.x));
        acc = r == 0 ? v : acc + v;
      }
      sd_s[DOFF + threadIdx.x] = acc;
    }
    if (STEP == 2 && threadIdx.x == 32) {
      long long c = 0;
      for (int r = 0; r < P.world; ++r)
        c += __ldcg(row_source<STEP>(P, r, row).head + (r == P.rank ? OFF_CNT : STAGE_CNT));
      cnt_s = c;
    }
    __syncthreads();
    RowWS* w = P.ws[P.rank] + row;
    resolve_row<STEP>(hs, sd_s, cnt_s, w, n_total, top, w->stats_out);
    __syncthreads();
    // the resolved words of this row into every peer's workspace
    if (threadIdx.x < ROW_WORDS - OFF_RES) {
      const long long v = reinterpret_cast<const long long*>(w)[OFF_RES + threadIdx.x];
      for (int r = 0; r < P.world; ++r)
        if (r != P.rank) reinterpret_cast<long long*>(P.ws[r] + row)[OFF_RES + threadIdx.x] = v;
    }
  }
  // the last block to get here tells the peers that this rank's rows are resolved
  __syncthreads();
  if (threadIdx.x < 32) {
    int last = 0;
    if (threadIdx.x == 0) {
      unsigned int* ticket = P.flags[P.rank] + FLAG_TICKET + 4 + STEP;
      __threadfence_system();
      last = atomicAdd(ticket, 1u) == gridDim.x - 1;
      if (last) *ticket = 0u;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last && ok) signal_peers(P, FLAG_RESOLVED + STEP);
  }
}

// One warp, before the next pass reads the resolved words: every owner has stored them here.
__global__ void rowstats_wait_kernel(const __grid_constant__ PeerSet P, int word) { wait_peers(P, word); }

// The statistics of all rows (resolved by their owners after pass 3) out of the workspace.
__global__ void __launch_bounds__(256)
rowstats_collect_kernel(const __grid_constant__ PeerSet P, int64_t rows, double* __restrict__ stats) {
  const bool ok = wait_peers(P, FLAG_RESOLVED + 3) && P.flags[P.rank][FLAG_ERROR_WORD] == 0u;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < rows * 12)
    stats[i] = ok ? P.ws[P.rank][i / 12].stats_out[i % 12] : __longlong_as_double(0x7ff8000000000000LL);
}

static int launch_pass(int pass, const float* values, int64_t rows, int64_t n, int64_t ld, RowWS* ws,
                       cudaStream_t st) {
  if (n <= 0 || rows <= 0) return 0;
  // Several waves of blocks (the last, partly filled wave then costs little), but
  // at least four tiles of work per block so that zeroing and flushing its
  // shared-memory histograms stays a small share of its time.  (Measured: a single
  // wave of 16-tile blocks is 30 % slower on 20 rows x 1e6 and 8 % slower on
  // 320 x 1e6 - the stream needs the many blocks to keep enough loads in flight.)
  const int sms = sm_count();
  const int64_t tile = (int64_t)RS_THREADS * RS_ITEMS;
  int64_t slices = ((int64_t)sms * 32 + rows - 1) / rows;
  const int64_t max_slices = std::max<int64_t>(1, n / (4 * tile));
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  if (rows > 65535) return set_error(B200_ELIMIT, "rowstats: rows=%lld > 65535 per call", (long long)rows);
  dim3 grid((unsigned)slices, (unsigned)rows);
  switch (pass) {
    case 0: rowstats_pass_kernel<0><<<grid, RS_THREADS, L1_BINS * 4, st>>>(values, n, ld, ws); break;
    case 1: rowstats_pass_kernel<1><<<grid, RS_THREADS, NT * L2_BINS * 4 + L1_BINS * 4, st>>>(values, n, ld, ws); break;
    case 2: rowstats_pass_kernel<2><<<grid, RS_THREADS, NT * L3_BINS * 4, st>>>(values, n, ld, ws); break;
    case 3: rowstats_pass_kernel<3><<<grid, RS_THREADS, 0, st>>>(values, n, ld, ws); break;
  }
  return check_cuda(cudaGetLastError(), "rowstats pass launch");
}

static int launch_resolve(int step, int64_t rows, int64_t n_total, int64_t top, RowWS* ws, double* stats,
                          cudaStream_t st) {
  static bool attr_set[64] = {false};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {   // step 1 walks four 2048-bin histograms of 8-byte counts: 64 KB
    B200_CUDA(cudaFuncSetAttribute(rowstats_resolve_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   NT * L2_BINS * 8));
    B200_CUDA(cudaFuncSetAttribute(rowstats_owner_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   NT * L2_BINS * 8));
    attr_set[dev] = true;
  }
  const unsigned g = (unsigned)rows;
  switch (step) {
    case 0: rowstats_resolve_kernel<0><<<g, 256, L1_BINS * 8, st>>>(ws, n_total, top, stats); break;
    case 1: rowstats_resolve_kernel<1><<<g, 256, NT * L2_BINS * 8, st>>>(ws, n_total, top, stats); break;
    case 2: rowstats_resolve_kernel<2><<<g, 256, NT * L3_BINS * 8, st>>>(ws, n_total, top, stats); break;
    case 3: rowstats_resolve_kernel<3><<<g, 256, 0, st>>>(ws, n_total, top, stats); break;
  }
  return check_cuda(cudaGetLastError(), "rowstats resolve launch");
}

static int launch_push(int step, int64_t rows, const PeerSet& P, cudaStream_t st) {
  const unsigned g = (unsigned)std::max<int64_t>(1, std::min<int64_t>(rows, (int64_t)sm_count() * 4));
  switch (step) {
    case 0: rowstats_push_kernel<0><<<g, 256, 0, st>>>(P, rows); break;
    case 1: rowstats_push_kernel<1><<<g, 256, 0, st>>>(P, rows); break;
    case 2: rowstats_push_kernel<2><<<g, 256, 0, st>>>(P, rows); break;
    case 3: rowstats_push_kernel<3><<<g, 256, 0, st>>>(P, rows); break;
  }
  return check_cuda(cudaGetLastError(), "rowstats push launch");
}

static int launch_owner(int step, int64_t rows, int64_t n_total, int64_t top, const PeerSet& P, cudaStream_t st) {
  static bool attr_set[64] = {false};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    B200_CUDA(cudaFuncSetAttribute(rowstats_owner_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   NT * L2_BINS * 8));
    attr_set[dev] = true;
  }
  const int64_t owned = rows > P.rank ? (rows - P.rank + P.world - 1) / P.world : 0;
  const unsigned g = (unsigned)std::max<int64_t>(owned, 1);
  switch (step) {
    case 0: rowstats_owner_kernel<0><<<g, 256, L1_BINS * 8, st>>>(P, rows, n_total, top); break;
    case 1: rowstats_owner_kernel<1><<<g, 256, NT * L2_BINS * 8, st>>>(P, rows, n_total, top); break;
    case 2: rowstats_owner_kernel<2><<<g, 256, NT * L3_BINS * 8, st>>>(P, rows, n_total, top); break;
    case 3: rowstats_owner_kernel<3><<<g, 256, 0, st>>>(P, rows, n_total, top); break;
  }
  return check_cuda(cudaGetLastError(), "rowstats owner launch");
}

}  // namespace b200

using namespace b200;

extern "C" int64_t b200_rowstats_stage_bytes(int64_t rows, int32_t world) {
  if (rows < 0 || world < 1) return 0;
  const int64_t owned = (rows + world - 1) / world;
  return (int64_t)world * owned * STAGE_WORDS * 8;
}

extern "C" int64_t b200_rowstats_workspace_bytes(int64_t rows) {
  return rows < 0 ? 0 : rows * (int64_t)sizeof(RowWS);
}

// Phases for a multi-GPU caller:
//   phase 0: clear + pass 0                      -> exchange D[0..1) and hist1
//   phase 1: resolve 0 + pass 1                  -> exchange hist2
//   phase 2: resolve 1 + pass 2                  -> exchange D[1..3), cnt and hist3
//   phase 3: resolve 2 + pass 3                  -> exchange D[3..9)
//   phase 4: resolve 3 (writes stats)
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
  B200_REQUIRE(phase >= -1 && phase <= 4, "rowstats: phase out of range");
  RowWS* ws = (RowWS*)workspace;
  const int first = phase < 0 ? 0 : phase, last = phase < 0 ? 4 : phase;
  for (int p = first; p <= last; ++p) {
    int rc = 0;
    if (p == 0) {
      B200_CUDA(cudaMemsetAsync(ws, 0, (size_t)rows * sizeof(RowWS), st));
    } else {
      rc = launch_resolve(p - 1, rows, n_total, top, ws, stats, st);
      if (rc) return rc;
    }
    if (p <= 3) {
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
    case 1: io = OFF_H2; ic = (int64_t)NT * L2_BINS; break;
    case 2: io = OFF_CNT; ic = 8 + (int64_t)NT * L3_BINS; d_o = 1; dc = 2; break;  // cnt words, then hist3
    case 3: d_o = 3; dc = 6; break;
    default: break;
  }
  out[0] = io; out[1] = ic; out[2] = d_o; out[3] = dc; out[4] = ROW_WORDS;
  return 0;
}

// The whole statistic block over investor shards of `world` GPUs of one node, the
// cross-GPU sums taken by the resolve kernels out of the peers' workspaces.
extern "C" int b200_rowstats_p2p(const float* values, int64_t rows, int64_t n, int64_t ld, int64_t n_total, int64_t top,
                                 const b200_peer_set* peers, double* stats, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_REQUIRE(peers != nullptr, "rowstats_p2p: peers is NULL");
  B200_REQUIRE(peers->world >= 1 && peers->world <= B200_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world,
               "rowstats_p2p: need 1 <= world <= %d and 0 <= rank < world", B200_MAX_PEERS);
  B200_REQUIRE(rows >= 0 && n >= 0, "rowstats_p2p: negative size");
  if (rows == 0) return 0;
  B200_REQUIRE(rows <= 65535, "rowstats_p2p: rows > 65535 per call");
  B200_REQUIRE(values != nullptr || n == 0, "rowstats_p2p: values is NULL");
  B200_REQUIRE(stats != nullptr, "rowstats_p2p: stats is NULL");
  B200_REQUIRE(n_total >= 2 && n <= n_total, "rowstats_p2p: need n_total >= 2 and n <= n_total");
  B200_REQUIRE(top >= 1 && top < n_total, "rowstats_p2p: need 1 <= top < n_total (top=%lld)", (long long)top);
  B200_REQUIRE(ld >= n, "rowstats_p2p: ld < n");
  B200_REQUIRE(n < (1ll << 32), "rowstats_p2p: a rank's partial counts travel as 32-bit words (n < 2^32 per rank)");
  PeerSet P;
  memset(&P, 0, sizeof(P));
  P.world = peers->world; P.rank = peers->rank; P.epoch = peers->epoch;
  for (int r = 0; r < peers->world; ++r) {
    B200_REQUIRE(peers->workspace[r] != nullptr && (peers->world == 1 || peers->flags[r] != nullptr),
                 "rowstats_p2p: workspace / flags of rank %d is NULL", r);
    P.ws[r] = (RowWS*)peers->workspace[r];
    P.flags[r] = peers->flags[r];
    P.stage[r] = (long long*)peers->stage[r];
    B200_REQUIRE(peers->world == 1 || P.stage[r] != nullptr, "rowstats_p2p: staging of rank %d is NULL", r);
  }
  P.stage_rows = peers->stage_rows;
  B200_REQUIRE(P.world == 1 || P.stage_rows * P.world >= rows,
               "rowstats_p2p: the staging holds %lld owned rows per rank, %lld rows need %lld", (long long)P.stage_rows,
               (long long)rows, (long long)((rows + P.world - 1) / P.world));
  RowWS* ws = P.ws[P.rank];
#ifdef B200_P2P_DEBUG   // per-kernel times of one call (tools/bench_series_n.py --debug): -DB200_P2P_DEBUG builds only
  cudaEvent_t ev[24];
  int nev = 0;
  for (auto& e : ev) cudaEventCreate(&e);
#define P2P_STAMP() cudaEventRecord(ev[nev++], st)
#else
#define P2P_STAMP()
#endif
  P2P_STAMP();
  B200_CUDA(cudaMemsetAsync(ws, 0, (size_t)rows * sizeof(RowWS), st));
  for (int p = 0; p < 4; ++p) {
    P2P_STAMP();
    if (int rc = launch_pass(p, values, rows, n, ld, ws, st)) return rc;
    P2P_STAMP();
    if (P.world == 1) {
      if (int rc = launch_resolve(p, rows, n_total, top, ws, stats, st)) return rc;
      continue;
    }
    if (int rc = launch_push(p, rows, P, st)) return rc;
    P2P_STAMP();
    if (int rc = launch_owner(p, rows, n_total, top, P, st)) return rc;
    P2P_STAMP();
    if (p < 3) {
      rowstats_wait_kernel<<<1, 32, 0, st>>>(P, FLAG_RESOLVED + p);
    } else {
      rowstats_collect_kernel<<<(unsigned)((rows * 12 + 255) / 256), 256, 0, st>>>(P, rows, stats);
    }
    B200_CUDA(cudaGetLastError());
  }
  P2P_STAMP();
#ifdef B200_P2P_DEBUG
  cudaStreamSynchronize(st);
  if (getenv("B200_P2P_PRINT")) {
    fprintf(stderr, "[p2p rank %d rows %lld]", P.rank, (long long)rows);
    for (int i = 1; i < nev; ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      fprintf(stderr, " %.1f", ms * 1e3f);
    }
    fprintf(stderr, " us\n");
  }
  for (auto& e : ev) cudaEventDestroy(e);
#endif
  return 0;
}
