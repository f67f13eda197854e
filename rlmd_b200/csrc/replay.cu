// K5 / K6: replay buffer - batched append with episode bookkeeping, unique
// index draws and the (multi-step) gather.
//
// Reference: tools/replay_torch.py (ReplayBufferTorch) - store_exp :167-197 with
// _episode_history :117-165; sample_exp :360-412 with _construct_history
// :199-247 and the n-step return :273-310, :336-345.  tools/replay.py is the
// same algorithm on NumPy fp64.
//
// The reference keeps, per finished episode, Python slices of the memories and
// searches them per sample.  Here the history of a sampled slot follows in
// O(1) from one int32 per slot (first slot of its episode) and a five-word
// header kept in device memory (restated and tested against
// the list bookkeeping in oracle/replay_oracle.py):
//   finished episode j >= 1 : start = episode_start[i], len = i - start + 1 (+1
//                             unless slot i is terminal: the history carries
//                             one FUTURE step)
//   first episode / nothing finished yet : start = 0, len = i + 1
//   episode still running   : the reference falls back to the prefix of
//                             episode 0: start = 0, len = min(i - e_last + 1, e_0 + 1)
//   eff = min(len, n);  R = sum / prod_{t < eff-1} fl32(gamma^t) * r[first + t],
//   first = start + len - eff;  s0 = next_state[first], a0 = action[first].
// Everything a sample needs is on the device, so store -> sample sequences run
// without a host round trip (and can be captured in a CUDA graph).
#include <algorithm>

#include <cstdlib>

#include "common.cuh"

namespace b200 {

enum { H_MEM_IDX = 0, H_EPISODES = 1, H_E0 = 2, H_ELAST = 3, H_RUN_START = 4 };

constexpr int STORE_HOST_MAX_WORDS = 480;  // doubles travelling as a kernel parameter

struct HostRow {
  double v[STORE_HOST_MAX_WORDS];
};

// ------------------------------------------------------------------ store
template <typename T>
__device__ __forceinline__ float to_reward(T r, double floor_) {
  const double x = (double)r;
  return (float)(x > floor_ ? x : floor_);  // max(reward, r_abs_zero) in double, then fp32 (:189)
}

// rows: flattened copy / conversion of state, action, next_state
template <typename T>
__global__ void __launch_bounds__(256)
replay_store_rows_kernel(const b200_replay_desc d, const T* __restrict__ state, const T* __restrict__ action,
                         const T* __restrict__ next_state, int64_t count, int64_t position) {
  const int S = d.state_dim, A = d.action_dim, W = 2 * S + A;
  const int64_t pos = position >= 0 ? position : d.header[H_MEM_IDX];
  const int64_t total = count * W;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = e / W;
    const int c = (int)(e - k * W);
    const int64_t slot = (pos + k) % d.mem_size;
    if (c < S) d.state_memory[slot * S + c] = (float)state[k * S + c];
    else if (c < S + A) d.action_memory[slot * A + (c - S)] = (float)action[k * A + (c - S)];
    else d.next_state_memory[slot * S + (c - S - A)] = (float)next_state[k * S + (c - S - A)];
  }
}

// scalars: reward, terminal flag and the first slot of the episode of every new
// slot = max(run_start before this call, 1 + last terminal slot before it in
// this call).  One 1024-wide tile per block: an inclusive max-scan inside the
// tile, and a backward search over the earlier `done` bytes for the tile prefix.
template <typename T>
__global__ void __launch_bounds__(1024)
replay_store_index_kernel(const b200_replay_desc d, const T* __restrict__ reward,
                          const uint8_t* __restrict__ done, int64_t count, int64_t position,
                          double reward_floor) {
  __shared__ long long warp_max[32];
  __shared__ long long tile_prefix;
  const int64_t pos = position >= 0 ? position : d.header[H_MEM_IDX];
  const int64_t tile0 = (int64_t)blockIdx.x * 1024;
  const int64_t k = tile0 + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  // boundary carried by the transitions before this tile
  if (threadIdx.x == 0) tile_prefix = -1;
  __syncthreads();
  for (int64_t hi = tile0; hi > 0; hi -= 1024) {
    const int64_t j = hi - 1 - threadIdx.x;
    long long cand = (j >= 0 && done[j]) ? (long long)(pos + j + 1) : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand = max(cand, __shfl_xor_sync(0xffffffffu, cand, o));
    if (lane == 0 && cand >= 0) atomicMax(&tile_prefix, cand);
    __syncthreads();
    if (tile_prefix >= 0) break;
    __syncthreads();
  }
  __syncthreads();
  long long carry = tile_prefix >= 0 ? tile_prefix : (long long)d.header[H_RUN_START];

  // exclusive max-scan of (pos + k + 1 where done) inside the tile
  const bool mine = k < count;
  const bool dn = mine && done[k] != 0;
  long long v = dn ? (long long)(pos + k + 1) : -1;
  long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long up = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc = max(inc, up);
  }
  if (lane == 31) warp_max[wid] = inc;
  __syncthreads();
  long long before = carry;
  for (int w = 0; w < wid; ++w) before = max(before, warp_max[w]);
  long long excl = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) excl = -1;
  const long long start = max(before, excl);

  if (mine) {
    const int64_t slot = (pos + k) % d.mem_size;
    d.reward_memory[slot] = to_reward(reward[k], reward_floor);
    d.terminal_memory[slot] = dn ? 1 : 0;
    d.episode_start[slot] = (int32_t)start;
  }
}

// header: finished-episode count, first / last terminal slot, write position.
// A few blocks scan the flags (16 per load, four loads in flight per thread, a word's four flags
// counted and located with bit operations); their partial results meet in the three spare header
// words (count + block ticket, first and last as running maxima: all zero between calls) and the
// last block to arrive updates the header.  (One block reading byte by byte took 166 us for a
// million flags - 2.5x the row-store kernel it follows; this takes ~10.)
enum { H_SCRATCH_CNT = 5, H_SCRATCH_FIRST = 6, H_SCRATCH_LAST = 7 };
constexpr int COMMIT_TICKET_SHIFT = 48;
constexpr long long COMMIT_BIG = 1ll << 62;

__global__ void __launch_bounds__(1024)
replay_commit_kernel(const b200_replay_desc d, const uint8_t* __restrict__ done, int64_t count, int64_t position) {
  __shared__ long long s_cnt[32], s_first[32], s_last[32];
  long long cnt = 0, first = COMMIT_BIG, last = -1;
  const uintptr_t addr = (uintptr_t)done;
  int64_t head = (int64_t)((16 - (addr & 15)) & 15);      // head / tail bytes around the aligned body: block 0
  if (head > count) head = count;
  const int64_t body = (count - head) >> 4;
  auto flag = [&](int64_t k, uint32_t v) {
    if (v) { ++cnt; first = min(first, (long long)k); last = max(last, (long long)k); }
  };
  if (blockIdx.x == 0) {
    for (int64_t k = threadIdx.x; k < head; k += blockDim.x) flag(k, done[k]);
    for (int64_t k = head + body * 16 + threadIdx.x; k < count; k += blockDim.x) flag(k, done[k]);
  }
  const uint4* __restrict__ q = reinterpret_cast<const uint4*>(done + head);
  constexpr int UN = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < body; i0 += UN * stride) {
    uint4 v[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t i = i0 + u * stride;
      v[u] = i < body ? q[i] : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if ((v[u].x | v[u].y | v[u].z | v[u].w) == 0u) continue;
      const int64_t i = i0 + u * stride;
      const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t t = w[c] | (w[c] >> 4);          // bit 0 of every byte = "the byte is not zero"
        t |= t >> 2;
        t |= t >> 1;
        t &= 0x01010101u;
        if (t) {
          const long long base = (long long)(head + i * 16 + c * 4);
          cnt += __popc(t);
          first = min(first, base + ((__ffs(t) - 1) >> 3));
          last = max(last, base + ((31 - __clz(t)) >> 3));
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_cnt[wid] = cnt; s_first[wid] = first; s_last[wid] = last; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      cnt += s_cnt[w];
      first = min(first, s_first[w]);
      last = max(last, s_last[w]);
    }
    int64_t* h = d.header;
    unsigned long long* sc = reinterpret_cast<unsigned long long*>(h);
    if (cnt > 0) {
      atomicMax(sc + H_SCRATCH_FIRST, (unsigned long long)(COMMIT_BIG - first));    // the smallest index wins
      atomicMax(sc + H_SCRATCH_LAST, (unsigned long long)(last + 1));
    }
    __threadfence();
    const unsigned long long mine = (unsigned long long)cnt + (1ull << COMMIT_TICKET_SHIFT);
    const unsigned long long old = atomicAdd(sc + H_SCRATCH_CNT, mine);
    if ((old >> COMMIT_TICKET_SHIFT) == (unsigned long long)gridDim.x - 1) {        // every block has arrived
      __threadfence();
      const long long total = (long long)((old + mine) & ((1ull << COMMIT_TICKET_SHIFT) - 1));
      const long long f = COMMIT_BIG - (long long)atomicMax(sc + H_SCRATCH_FIRST, 0ull);
      const long long l = (long long)atomicMax(sc + H_SCRATCH_LAST, 0ull) - 1;
      const int64_t pos = position >= 0 ? position : h[H_MEM_IDX];
      if (total > 0) {
        if (h[H_EPISODES] == 0) h[H_E0] = pos + f;
        h[H_ELAST] = pos + l;
        h[H_RUN_START] = pos + l + 1;
        h[H_EPISODES] += total;
      }
      h[H_MEM_IDX] = pos + count;
      h[H_SCRATCH_CNT] = 0; h[H_SCRATCH_FIRST] = 0; h[H_SCRATCH_LAST] = 0;
    }
  }
}

// one transition handed over as a kernel parameter (no H2D copy): the
// store_exp(state, action, reward, next_state, done) call of the training loop
__global__ void __launch_bounds__(128)
replay_store_host_kernel(const b200_replay_desc d, const __grid_constant__ HostRow row, int done, int64_t position,
                         double reward_floor) {
  const int S = d.state_dim, A = d.action_dim;
  int64_t* h = d.header;
  const int64_t pos = position >= 0 ? position : h[H_MEM_IDX];
  __syncthreads();  // everybody has read the header before thread 0 advances it
  const int64_t slot = pos % d.mem_size;
  for (int c = threadIdx.x; c < 2 * S + A; c += blockDim.x) {
    const float v = (float)row.v[c];
    if (c < S) d.state_memory[slot * S + c] = v;
    else if (c < S + A) d.action_memory[slot * A + (c - S)] = v;
    else d.next_state_memory[slot * S + (c - S - A)] = v;
  }
  if (threadIdx.x == 0) {
    d.reward_memory[slot] = to_reward(row.v[2 * S + A], reward_floor);
    d.terminal_memory[slot] = done ? 1 : 0;
    d.episode_start[slot] = (int32_t)h[H_RUN_START];
    if (done) {
      if (h[H_EPISODES] == 0) h[H_E0] = pos;
      h[H_ELAST] = pos;
      h[H_RUN_START] = pos + 1;
      h[H_EPISODES] += 1;
    }
    h[H_MEM_IDX] = pos + 1;
  }
}

// ------------------------------------------------------------------- draw
// B distinct uniform indices in [0, filled) per batch, one block per batch.
// Round r: every pending position t draws
//   v = mulhi64(philox(t, r, batch, draw_index lo; seed ^ draw_index hi).xy, filled)
// and claims v in a shared-memory hash set; the claim with the smallest
// (round, t) owns the value, everybody else redraws in round r+1.  The result
// depends only on (seed, draw_index, batch, filled, B) - oracle/replay_oracle.py
// restates it word for word.
constexpr unsigned long long SET_EMPTY = ~0ull;

__device__ __forceinline__ void draw_block(const b200_replay_desc& d, int64_t filled_arg, int32_t B, uint64_t seed,
                                           uint64_t draw_index, const int64_t* __restrict__ draw_counter,
                                           int64_t lanes, int64_t lane_len, int32_t table_size,
                                           int64_t* __restrict__ out_idx, unsigned long long* table) {
  // lanes > 1 (collector): every lane holds the same number of transitions and the
  // memory is slot-major (slot = local * lanes + lane), so the filled part is the
  // prefix [0, lanes * lane_filled) and a drawn value is a slot
  const int64_t lane_filled = filled_arg >= 0 ? filled_arg : min(d.header[H_MEM_IDX], lane_len);
  const int64_t filled = lanes * lane_filled;
  if (draw_counter != nullptr) draw_index += (uint64_t)draw_counter[0];
  const uint32_t mask = (uint32_t)table_size - 1u;
  const uint32_t k0 = (uint32_t)seed ^ (uint32_t)(draw_index >> 32), k1 = (uint32_t)(seed >> 32);
  for (int s = threadIdx.x; s < table_size; s += blockDim.x) table[s] = SET_EMPTY;
  __syncthreads();
  // positions owned by this thread: t = threadIdx.x + j * blockDim.x (B <= 8 * 1024)
  uint32_t pending_bits = 0;
  uint32_t val[8];
  for (int j = 0; j < 8; ++j)
    if ((int)threadIdx.x + j * (int)blockDim.x < B) pending_bits |= 1u << j;
  if (filled <= 0 || filled < B) {  // cannot draw B distinct values: flag every position
    for (int j = 0; j < 8; ++j)
      if (pending_bits >> j & 1) out_idx[(int64_t)blockIdx.x * B + threadIdx.x + j * blockDim.x] = -1;
    return;
  }
  for (uint32_t round = 0;; ++round) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (!(pending_bits >> j & 1)) continue;
      const uint32_t t = threadIdx.x + j * blockDim.x;
      const Philox4 p = philox4x32_10(t, round, blockIdx.x, (uint32_t)draw_index ^ PHILOX_TAG_REPLAY, k0, k1);
      const unsigned long long u = ((unsigned long long)p.x << 32) | p.y;
      const uint32_t v = (uint32_t)__umul64hi(u, (unsigned long long)filled);
      val[j] = v;
      const unsigned long long mine = ((unsigned long long)v << 32) | ((unsigned long long)round << 16) | t;
      uint32_t s = (v * 0x9E3779B1u) & mask;
      for (;;) {
        unsigned long long cur = table[s];
        if (cur == SET_EMPTY) {
          cur = atomicCAS(&table[s], SET_EMPTY, mine);
          if (cur == SET_EMPTY) break;
        }
        if ((uint32_t)(cur >> 32) == v) {
          atomicMin(&table[s], mine);
          break;
        }
        s = (s + 1) & mask;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (!(pending_bits >> j & 1)) continue;
      const uint32_t t = threadIdx.x + j * blockDim.x;
      const uint32_t v = val[j];
      uint32_t s = (v * 0x9E3779B1u) & mask;
      while ((uint32_t)(table[s] >> 32) != v) s = (s + 1) & mask;
      if ((uint32_t)(table[s] & 0xffffffffu) == ((round << 16) | t)) {
        out_idx[(int64_t)blockIdx.x * B + t] = (int64_t)v;
        pending_bits &= ~(1u << j);
      }
    }
    if (!__syncthreads_or(pending_bits != 0)) break;
  }
}

__global__ void __launch_bounds__(1024)
replay_draw_kernel(const b200_replay_desc d, int64_t filled_arg, int32_t B, uint64_t seed, uint64_t draw_index,
                   const int64_t* __restrict__ draw_counter, int64_t lanes, int64_t lane_len, int32_t table_size,
                   int64_t* __restrict__ out_idx) {
  extern __shared__ unsigned long long table[];
  draw_block(d, filled_arg, B, seed, draw_index, draw_counter, lanes, lane_len, table_size, out_idx, table);
}

// ----------------------------------------------------------------- gather
struct GammaPow {
  float v[B200_REPLAY_MAX_STEPS];
};

template <int LANES>
__global__ void __launch_bounds__(256)
replay_gather_kernel(const b200_replay_desc d, const int64_t* __restrict__ idx, int64_t n_samples,
                     int64_t filled_arg, int64_t lanes, int64_t lane_len, int32_t n_steps, int32_t additive,
                     const __grid_constant__ GammaPow gp,
                     float* __restrict__ out_state, float* __restrict__ out_action, float* __restrict__ out_reward,
                     float* __restrict__ out_next_state, uint8_t* __restrict__ out_done,
                     int64_t* __restrict__ out_eff) {
  const int S = d.state_dim, A = d.action_dim;
  const int lane = threadIdx.x % LANES;
  const unsigned gmask = LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x & 31) / LANES * LANES));
  const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / LANES;
  // A lane is a reference buffer of lane_len transitions with its own header,
  // stored slot-major: local index j of lane e sits at slot j * lanes + e.  The
  // plain replay buffer is the one-lane case (lane_len = mem_size).
  for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES; q < n_samples; q += groups) {
    const int64_t slot = idx[q];
    const int64_t lane_id = (slot >= 0 && lanes > 1) ? slot % lanes : 0;
    const int64_t i = (slot >= 0 && lanes > 1) ? slot / lanes : slot;
    auto at = [&](int64_t j) { return j * lanes + lane_id; };
    const int64_t* __restrict__ hdr = d.header + lane_id * 8;
    const int64_t filled = filled_arg >= 0 ? filled_arg : min(hdr[H_MEM_IDX], lane_len);
    const int64_t episodes = hdr[H_EPISODES], e0 = hdr[H_E0], elast = hdr[H_ELAST];
    const bool ok = slot >= 0 && i < filled;
    int64_t first = i;
    int eff = ok ? 1 : 0;
    float R = 0.0f;
    const float* src_state = d.state_memory;
    uint8_t term = 0;
    if (ok) {
      term = d.terminal_memory[at(i)];
      if (n_steps <= 1) {
        R = d.reward_memory[at(i)];
      } else {
        int64_t start, len;
        if (episodes == 0 || i <= e0) { start = 0; len = i + 1; }
        else if (i <= elast) { start = d.episode_start[at(i)]; len = i - start + 1 + (term ? 0 : 1); }
        else { start = 0; len = min(i - elast + 1, e0 + 1); }
        eff = (int)min(len, (int64_t)n_steps);
        first = start + len - eff;
        src_state = d.next_state_memory;  // the reference's histories hold next_state (:132)
        float acc = additive ? 0.0f : 1.0f;
        for (int t0 = 0; t0 < n_steps - 1; t0 += LANES) {
          const int t = t0 + lane;
          float term_t = 0.0f;
          if (t < eff - 1) term_t = gp.v[t] * d.reward_memory[at(first + t)];
          const int lim = min(LANES, n_steps - 1 - t0);
          for (int j = 0; j < lim; ++j) {
            const float x = __shfl_sync(gmask, term_t, j, LANES);
            if (t0 + j < eff - 1) acc = additive ? acc + x : acc * x;
          }
        }
        R = acc;
      }
    }
    for (int c = lane; c < S; c += LANES) {
      out_state[q * S + c] = ok ? src_state[at(first) * S + c] : 0.0f;
      out_next_state[q * S + c] = ok ? d.next_state_memory[at(i) * S + c] : 0.0f;
    }
    for (int c = lane; c < A; c += LANES) out_action[q * A + c] = ok ? d.action_memory[at(first) * A + c] : 0.0f;
    if (lane == 0) {
      out_reward[q] = R;
      out_done[q] = term;
      out_eff[q] = eff;
    }
  }
}

// Narrow rows (state_dim < 32, the multiplicative envs' 5..12 floats): one THREAD
// per sample.  Every load of a sample that does not depend on another one is
// issued before the first use, so a resident warp keeps 32 samples x ~12 loads in
// flight (the lane-group kernel above keeps 4): the gather is a chain of three
// dependent memory latencies (slot -> bookkeeping -> rows), and throughput is
// the number of chains in flight.  Same arithmetic, term by term.
__device__ __forceinline__ void gather_sample(const b200_replay_desc& d, const int64_t* __restrict__ idx, int64_t q,
                                              int64_t filled_arg, int64_t lanes, int64_t lane_len, int32_t n_steps,
                                              int32_t additive, const GammaPow& gp, float* __restrict__ out_state,
                                              float* __restrict__ out_action, float* __restrict__ out_reward,
                                              float* __restrict__ out_next_state, uint8_t* __restrict__ out_done,
                                              int64_t* __restrict__ out_eff) {
  const int S = d.state_dim, A = d.action_dim;
  {
    const int64_t slot = idx[q];
    const int64_t lane_id = (slot >= 0 && lanes > 1) ? slot % lanes : 0;
    const int64_t i = (slot >= 0 && lanes > 1) ? slot / lanes : slot;
    auto at = [&](int64_t j) { return j * lanes + lane_id; };
    const int64_t* __restrict__ hdr = d.header + lane_id * 8;
    const int64_t filled = filled_arg >= 0 ? filled_arg : min(hdr[H_MEM_IDX], lane_len);
    const int64_t episodes = hdr[H_EPISODES], e0 = hdr[H_E0], elast = hdr[H_ELAST];
    const bool ok = slot >= 0 && i < filled;
    int64_t first = ok ? i : 0;
    int eff = ok ? 1 : 0;
    float R = 0.0f;
    uint8_t term = 0;
    const float* src_state = d.state_memory;
    if (ok) {
      // independent of each other: issued together
      term = d.terminal_memory[at(i)];
      const int32_t es = d.episode_start[at(i)];
      const float ri = d.reward_memory[at(i)];
      if (n_steps <= 1) {
        R = ri;
      } else {
        int64_t start, len;
        if (episodes == 0 || i <= e0) { start = 0; len = i + 1; }
        else if (i <= elast) { start = es; len = i - start + 1 + (term ? 0 : 1); }
        else { start = 0; len = min(i - elast + 1, e0 + 1); }
        eff = (int)min(len, (int64_t)n_steps);
        first = start + len - eff;
        src_state = d.next_state_memory;
        float acc = additive ? 0.0f : 1.0f;
        for (int t0 = 0; t0 < eff - 1; t0 += 8) {
          float rw[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) rw[u] = (t0 + u < eff - 1) ? d.reward_memory[at(first + t0 + u)] : 0.0f;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (t0 + u < eff - 1) {
              const float x = gp.v[t0 + u] * rw[u];
              acc = additive ? acc + x : acc * x;
            }
          }
        }
        R = acc;
      }
    }
    const float* __restrict__ rs = src_state + at(first) * S;
    const float* __restrict__ rn = d.next_state_memory + (ok ? at(i) : 0) * S;
    const float* __restrict__ ra = d.action_memory + at(first) * A;
    for (int c0 = 0; c0 < S; c0 += 8) {
      float vs[8], vn[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool in = ok && c0 + u < S;
        vs[u] = in ? rs[c0 + u] : 0.0f;
        vn[u] = in ? rn[c0 + u] : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (c0 + u < S) { out_state[q * S + c0 + u] = vs[u]; out_next_state[q * S + c0 + u] = vn[u]; }
    }
    for (int c = 0; c < A; ++c) out_action[q * A + c] = ok ? ra[c] : 0.0f;
    out_reward[q] = R;
    out_done[q] = term;
    out_eff[q] = eff;
  }
}

__global__ void __launch_bounds__(256)
replay_gather_thread_kernel(const b200_replay_desc d, const int64_t* __restrict__ idx, int64_t n_samples,
                            int64_t filled_arg, int64_t lanes, int64_t lane_len, int32_t n_steps, int32_t additive,
                            const __grid_constant__ GammaPow gp, float* __restrict__ out_state,
                            float* __restrict__ out_action, float* __restrict__ out_reward,
                            float* __restrict__ out_next_state, uint8_t* __restrict__ out_done,
                            int64_t* __restrict__ out_eff) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_samples;
       q += (int64_t)gridDim.x * blockDim.x)
    gather_sample(d, idx, q, filled_arg, lanes, lane_len, n_steps, additive, gp, out_state, out_action, out_reward,
                  out_next_state, out_done, out_eff);
}

// Many samples per launch (>= 4096, narrow rows).  The per-thread gather above reads a 20-byte
// row as five scattered 4-byte loads and writes it as five strided stores: every 128-byte line
// is touched several times by different warp instructions.  Here a warp owns 32 samples: the owner lane of a sample resolves
// the bookkeeping (slot -> terminal flag, episode start -> first, eff), then QUADS of
// lanes fetch each row - and, for one-lane buffers, the run of rewards - as aligned
// 16-byte vectors of the window that contains it (every line is touched once per warp
// instruction), into shared memory; the n-step return is folded from there in the same
// order, and all outputs leave as contiguous runs of the warp's 32 samples.  Measured at
// 16384 x 256 samples per launch from a 1e6 buffer (n = 1 / 5 / 10): 2.06 / 1.93 / 1.88e10
// samples/s against 2.01 / 1.81 / 1.69e10 for the per-thread gather.
constexpr int BULK_WIN = 16;        // floats of a row window: 3 + state_dim <= 16
constexpr int BULK_RWIN = 12;       // floats of the reward window: 3 + (n_steps - 1) <= 12
constexpr int BULK_WARP_FLOATS = 32 * (2 * BULK_WIN + BULK_RWIN) + 32;   // + the samples' three window offsets

// The aligned 16-byte vector `v` of the window that starts at element `w0` (a multiple of 4) of `base`
// (16-byte aligned, `total` elements); elements beyond the array read as 0.
__device__ __forceinline__ float4 window_vec(const float* __restrict__ base, int64_t w0, int v, int64_t total) {
  const int64_t e = w0 + 4 * v;
  if (e + 4 <= total) return __ldg(reinterpret_cast<const float4*>(base + e));
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (e < total) r.x = base[e];
  if (e + 1 < total) r.y = base[e + 1];
  if (e + 2 < total) r.z = base[e + 2];
  return r;
}

// One group of 32 consecutive samples (q = g * 32 + lane), one warp; `sw`: the warp's BULK_WARP_FLOATS of shared memory.
__device__ __forceinline__ void bulk_gather_group(const b200_replay_desc& d, const int64_t* __restrict__ idx,
                                                  int64_t g, int64_t n_samples, int64_t filled_arg, int64_t lanes,
                                                  int64_t lane_len, int32_t n_steps, int32_t additive, const GammaPow& gp,
                                                  float* __restrict__ sw, float* __restrict__ out_state,
                                                  float* __restrict__ out_action, float* __restrict__ out_reward,
                                                  float* __restrict__ out_next_state, uint8_t* __restrict__ out_done,
                                                  int64_t* __restrict__ out_eff) {
  const int S = d.state_dim, A = d.action_dim;
  const int lane = threadIdx.x & 31;
  float* nw = sw + 32 * BULK_WIN;                               // [32][BULK_WIN] next_state windows (sw: state windows)
  float* rw = nw + 32 * BULK_WIN;                               // [32][BULK_RWIN] reward windows
  uint32_t* offs = reinterpret_cast<uint32_t*>(rw + 32 * BULK_RWIN);   // [32] o_state | o_next << 8 | o_reward << 16 | ok << 24
  const int64_t row_total = d.mem_size * (int64_t)S;
  const bool reward_windows = n_steps > 1 && lanes == 1;
  {
    const int64_t q = g * 32 + lane;
    // ---- the owner lane resolves its sample (same arithmetic as gather_sample)
    int64_t e_state = 0, e_next = 0, e_rew = 0;     // element offsets of the two rows / the first reward
    const float* src_state = d.state_memory;
    bool ok = false;
    int eff = 0;
    uint8_t term = 0;
    float R = 0.0f;
    int64_t first = 0, lane_id = 0;
    if (q < n_samples) {
      const int64_t slot = idx[q];
      lane_id = (slot >= 0 && lanes > 1) ? slot % lanes : 0;
      const int64_t i = (slot >= 0 && lanes > 1) ? slot / lanes : slot;
      const int64_t* __restrict__ hdr = d.header + lane_id * 8;
      const int64_t filled = filled_arg >= 0 ? filled_arg : min(hdr[H_MEM_IDX], lane_len);
      const int64_t episodes = hdr[H_EPISODES], e0 = hdr[H_E0], elast = hdr[H_ELAST];
      ok = slot >= 0 && i < filled;
      first = ok ? i : 0;
      eff = ok ? 1 : 0;
      if (ok) {
        const int64_t si = i * lanes + lane_id;
        term = d.terminal_memory[si];
        if (n_steps <= 1) {
          R = d.reward_memory[si];
        } else {
          const int32_t es = d.episode_start[si];
          int64_t start, len;
          if (episodes == 0 || i <= e0) { start = 0; len = i + 1; }
          else if (i <= elast) { start = es; len = i - start + 1 + (term ? 0 : 1); }
          else { start = 0; len = min(i - elast + 1, e0 + 1); }
          eff = (int)min(len, (int64_t)n_steps);
          first = start + len - eff;
          src_state = d.next_state_memory;
        }
        e_state = (first * lanes + lane_id) * S;
        e_next = si * S;
        e_rew = first;                               // lanes == 1 when the window is used
      }
    }
    const bool nstep_src = src_state != d.state_memory;
    offs[lane] = (uint32_t)(e_state & 3) | (uint32_t)(e_next & 3) << 8 | (uint32_t)(e_rew & 3) << 16 | (ok ? 1u << 24 : 0u);
    // ---- quads fetch the windows: pass k serves samples 8k .. 8k+7, lane & 3 is the vector of the window
    const int vec = lane & 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int sj = 8 * k + (lane >> 2);
      const int64_t ws = __shfl_sync(0xffffffffu, e_state & ~(int64_t)3, sj);
      const int64_t wn = __shfl_sync(0xffffffffu, e_next & ~(int64_t)3, sj);
      const int64_t wr = __shfl_sync(0xffffffffu, e_rew & ~(int64_t)3, sj);
      const int okj = __shfl_sync(0xffffffffu, (int)ok, sj);
      const int nsj = __shfl_sync(0xffffffffu, (int)nstep_src, sj);
      const int effj = __shfl_sync(0xffffffffu, eff, sj);
      const int os = (int)(__shfl_sync(0xffffffffu, (int)(e_state & 3), sj));
      const int on = (int)(__shfl_sync(0xffffffffu, (int)(e_next & 3), sj));
      const int orw = (int)(__shfl_sync(0xffffffffu, (int)(e_rew & 3), sj));
      if (okj) {
        if (4 * vec < os + S)
          reinterpret_cast<float4*>(sw + sj * BULK_WIN)[vec] =
              window_vec(nsj ? d.next_state_memory : d.state_memory, ws, vec, row_total);
        if (4 * vec < on + S)
          reinterpret_cast<float4*>(nw + sj * BULK_WIN)[vec] = window_vec(d.next_state_memory, wn, vec, row_total);
        if (reward_windows && vec < 3 && 4 * vec < orw + effj - 1)
          reinterpret_cast<float4*>(rw + sj * BULK_RWIN)[vec] = window_vec(d.reward_memory, wr, vec, d.mem_size);
      }
    }
    __syncwarp();
    // ---- the n-step return, folded in the reference's order
    if (ok && n_steps > 1) {
      float acc = additive ? 0.0f : 1.0f;
      if (reward_windows) {
        const float* r = rw + lane * BULK_RWIN + (int)(e_rew & 3);
        for (int t = 0; t < eff - 1; ++t) {
          const float x = gp.v[t] * r[t];
          acc = additive ? acc + x : acc * x;
        }
      } else {
        for (int t = 0; t < eff - 1; ++t) {
          const float x = gp.v[t] * d.reward_memory[(first + t) * lanes + lane_id];
          acc = additive ? acc + x : acc * x;
        }
      }
      R = acc;
    }
    if (q < n_samples) {
      for (int c = 0; c < A; ++c) out_action[q * A + c] = ok ? d.action_memory[(first * lanes + lane_id) * A + c] : 0.0f;
      out_reward[q] = R;
      out_done[q] = term;
      out_eff[q] = eff;
    }
    // ---- the rows of the warp's 32 samples are one contiguous run of each output
    const int64_t run0 = g * 32 * S;
    const int64_t run = min((int64_t)32, n_samples - g * 32) * S;
    for (int e = lane; e < run; e += 32) {
      const int sj = e / S, c = e - sj * S;
      const uint32_t o = offs[sj];
      const bool okj = o >> 24 & 1u;
      out_state[run0 + e] = okj ? sw[sj * BULK_WIN + (o & 3u) + c] : 0.0f;
      out_next_state[run0 + e] = okj ? nw[sj * BULK_WIN + (o >> 8 & 3u) + c] : 0.0f;
    }
  }
}

__global__ void __launch_bounds__(256)
replay_gather_bulk_kernel(const b200_replay_desc d, const int64_t* __restrict__ idx, int64_t n_samples,
                          int64_t filled_arg, int64_t lanes, int64_t lane_len, int32_t n_steps, int32_t additive,
                          const __grid_constant__ GammaPow gp, float* __restrict__ out_state,
                          float* __restrict__ out_action, float* __restrict__ out_reward,
                          float* __restrict__ out_next_state, uint8_t* __restrict__ out_done,
                          int64_t* __restrict__ out_eff) {
  extern __shared__ float bulk_smem[];
  const int warp = threadIdx.x >> 5;
  float* sw = bulk_smem + (size_t)warp * BULK_WARP_FLOATS;
  const int64_t n_groups = (n_samples + 31) / 32;
  for (int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; g < n_groups; g += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    bulk_gather_group(d, idx, g, n_samples, filled_arg, lanes, lane_len, n_steps, additive, gp, sw, out_state, out_action,
                      out_reward, out_next_state, out_done, out_eff);
    __syncwarp();    // the windows are rewritten by the next group
  }
}

// Narrow rows, indices drawn on the device: ONE launch per call - the block that
// drew a mini-batch's indices gathers its samples (a 256-sample call is bound by
// launch latency: two dependent launches cost twice as much as one).
__global__ void __launch_bounds__(1024)
replay_draw_gather_kernel(const b200_replay_desc d, int64_t filled_arg, int32_t B, uint64_t seed, uint64_t draw_index,
                          const int64_t* __restrict__ draw_counter, int64_t lanes, int64_t lane_len,
                          int32_t table_size, int64_t* __restrict__ out_idx, int32_t n_steps, int32_t additive,
                          const __grid_constant__ GammaPow gp, float* __restrict__ out_state,
                          float* __restrict__ out_action, float* __restrict__ out_reward,
                          float* __restrict__ out_next_state, uint8_t* __restrict__ out_done,
                          int64_t* __restrict__ out_eff, int64_t* __restrict__ bump) {
  extern __shared__ unsigned long long table[];
  draw_block(d, filled_arg, B, seed, draw_index, draw_counter, lanes, lane_len, table_size, out_idx, table);
  __syncthreads();   // the block's own writes of out_idx
  for (int t = threadIdx.x; t < B; t += blockDim.x)
    gather_sample(d, out_idx, (int64_t)blockIdx.x * B + t, filled_arg, lanes, lane_len, n_steps, additive, gp,
                  out_state, out_action, out_reward, out_next_state, out_done, out_eff);
  // bump = {draw counter, ticket}: the last block to get here advances the counter (every block has read
  // it by then) - the counted call is this ONE launch, nothing to bump afterwards
  if (bump != nullptr && threadIdx.x == 0) {
    const unsigned long long t = atomicAdd(reinterpret_cast<unsigned long long*>(bump + 1), 1ull);
    if (t == (unsigned long long)gridDim.x - 1) { bump[1] = 0; bump[0] += 1; }
  }
}

static int check_desc(const b200_replay_desc* d) {
  B200_REQUIRE(d != nullptr, "replay: desc is NULL");
  B200_REQUIRE(d->mem_size > 0 && d->mem_size < (1ll << 31), "replay: mem_size %lld outside 1..2^31-1",
               (long long)d->mem_size);
  B200_REQUIRE(d->state_dim > 0 && d->action_dim > 0, "replay: state_dim / action_dim must be positive");
  B200_REQUIRE(d->state_memory && d->action_memory && d->reward_memory && d->next_state_memory &&
                   d->terminal_memory && d->episode_start && d->header,
               "replay: a buffer pointer in the desc is NULL");
  return 0;
}

static int require_device() {
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  return 0;
}

template <typename T>
static int store_impl(const b200_replay_desc* d, const void* state, const void* action, const void* reward,
                      const void* next_state, const uint8_t* done, int64_t count, int64_t position,
                      double reward_floor, cudaStream_t st) {
  const int W = 2 * d->state_dim + d->action_dim;
  const int64_t total = count * W;
  const int grid_rows = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 8);
  replay_store_rows_kernel<T><<<grid_rows, 256, 0, st>>>(*d, (const T*)state, (const T*)action,
                                                         (const T*)next_state, count, position);
  const int tiles = (int)((count + 1023) / 1024);
  replay_store_index_kernel<T><<<tiles, 1024, 0, st>>>(*d, (const T*)reward, done, count, position, reward_floor);
  const int cblocks = (int)std::max<int64_t>(1, std::min<int64_t>(64, (count + 65535) / 65536));
  replay_commit_kernel<<<cblocks, 1024, 0, st>>>(*d, done, count, position);
  B200_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_replay_reset(const b200_replay_desc* d, void* stream) {
  if (int rc = require_device()) return rc;
  if (int rc = check_desc(d)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  B200_CUDA(cudaMemsetAsync(d->header, 0, 8 * sizeof(int64_t), st));
  B200_CUDA(cudaMemsetAsync(d->terminal_memory, 0, (size_t)d->mem_size, st));
  B200_CUDA(cudaMemsetAsync(d->episode_start, 0, (size_t)d->mem_size * sizeof(int32_t), st));
  return 0;
}

extern "C" int b200_replay_store(const b200_replay_desc* d, const void* state, const void* action,
                                 const void* reward, const void* next_state, const uint8_t* done, int64_t count,
                                 int32_t is_f64, int64_t position, double reward_floor, void* stream) {
  if (int rc = require_device()) return rc;
  if (int rc = check_desc(d)) return rc;
  B200_REQUIRE(count >= 0, "replay_store: count %lld < 0", (long long)count);
  if (count == 0) return 0;
  B200_REQUIRE(state && action && reward && next_state && done, "replay_store: an input pointer is NULL");
  B200_REQUIRE(count <= d->mem_size, "replay_store: count %lld exceeds mem_size %lld", (long long)count,
               (long long)d->mem_size);
  cudaStream_t st = (cudaStream_t)stream;
  return is_f64 ? store_impl<double>(d, state, action, reward, next_state, done, count, position, reward_floor, st)
                : store_impl<float>(d, state, action, reward, next_state, done, count, position, reward_floor, st);
}

extern "C" int b200_replay_store_host(const b200_replay_desc* d, const double* state_host, const double* action_host,
                                      double reward, const double* next_state_host, int32_t done,
                                      int64_t position, double reward_floor, void* stream) {
  if (int rc = require_device()) return rc;
  if (int rc = check_desc(d)) return rc;
  B200_REQUIRE(state_host && action_host && next_state_host, "replay_store_host: an input pointer is NULL");
  const int S = d->state_dim, A = d->action_dim;
  if (2 * S + A + 1 > STORE_HOST_MAX_WORDS)
    return set_error(B200_ELIMIT, "replay_store_host: 2*state_dim + action_dim + 1 = %d exceeds %d words; use "
                     "b200_replay_store with device buffers", 2 * S + A + 1, STORE_HOST_MAX_WORDS);
  HostRow row;
  for (int c = 0; c < S; ++c) row.v[c] = state_host[c];
  for (int c = 0; c < A; ++c) row.v[S + c] = action_host[c];
  for (int c = 0; c < S; ++c) row.v[S + A + c] = next_state_host[c];
  row.v[2 * S + A] = reward;
  replay_store_host_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(*d, row, done, position, reward_floor);
  B200_CUDA(cudaGetLastError());
  return 0;
}

namespace b200 {
int replay_sample_lanes(const b200_replay_desc* d, int64_t lanes, int64_t lane_len, const int64_t* idx,
                        int64_t n_batches, int32_t batch, int64_t filled, int32_t multi_steps,
                        const float* gamma_pow_host, int32_t additive, uint64_t seed, uint64_t draw_index,
                        const int64_t* draw_counter, int64_t* out_idx, float* out_state, float* out_action,
                        float* out_reward, float* out_next_state, uint8_t* out_done, int64_t* out_eff,
                        void* stream, int64_t* bump, int* bumped) {
  if (bumped) *bumped = 0;
  if (int rc = require_device()) return rc;
  if (int rc = check_desc(d)) return rc;
  B200_REQUIRE(lanes >= 1 && lane_len >= 1 && lanes * lane_len <= d->mem_size, "replay_sample: lanes x lane_len exceeds mem_size");
  B200_REQUIRE(n_batches >= 0 && batch >= 0, "replay_sample: negative batch count / size");
  const int64_t n_samples = n_batches * batch;
  if (n_samples == 0) return 0;
  B200_REQUIRE(multi_steps >= 1, "replay_sample: multi_steps %d < 1", multi_steps);
  if (multi_steps > B200_REPLAY_MAX_STEPS)
    return set_error(B200_ELIMIT, "replay_sample: multi_steps %d exceeds %d", multi_steps, B200_REPLAY_MAX_STEPS);
  B200_REQUIRE(multi_steps == 1 || gamma_pow_host != nullptr, "replay_sample: gamma_pow_host is NULL");
  B200_REQUIRE(out_state && out_action && out_reward && out_next_state && out_done && out_eff,
               "replay_sample: an output pointer is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  GammaPow gp;
  for (int t = 0; t < B200_REPLAY_MAX_STEPS; ++t) gp.v[t] = (multi_steps > 1 && t < multi_steps) ? gamma_pow_host[t] : 0.0f;
  // >= 4096 samples of narrow rows per launch: the warp-cooperative gather (B200_REPLAY_NO_BULK: A/B timing)
  const auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  static const bool no_bulk = getenv("B200_REPLAY_NO_BULK") != nullptr;
  const bool bulk_ok = !no_bulk && n_samples >= 4096 && d->state_dim + 3 <= BULK_WIN &&
                       (multi_steps <= 1 || lanes > 1 || multi_steps - 1 + 3 <= BULK_RWIN) &&
                       aligned16(d->state_memory) && aligned16(d->next_state_memory) && aligned16(d->reward_memory);
  if (idx == nullptr) {
    B200_REQUIRE(out_idx != nullptr, "replay_sample: out_idx is NULL while indices are drawn on the device");
    if (batch > 8192) return set_error(B200_ELIMIT, "replay_sample: on-device draw supports batch <= 8192 (got %d)", batch);
    B200_REQUIRE(n_batches <= 0x7fffffff, "replay_sample: too many batches");
    int table = 64;
    while (table < 2 * batch) table <<= 1;
    const int threads = std::min(1024, (batch + 31) / 32 * 32);
    const size_t smem = (size_t)table * sizeof(unsigned long long);
    static bool attr_set[64] = {false};
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    if (smem > 48 * 1024 && dev < 64 && !attr_set[dev]) {
      B200_CUDA(cudaFuncSetAttribute(replay_draw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
      B200_CUDA(cudaFuncSetAttribute(replay_draw_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     128 * 1024));
      attr_set[dev] = true;
    }
    // narrow rows: ONE launch, the block that drew a mini-batch gathers it.  A 256-sample call is bound by launch
    // latency (two dependent launches cost twice as much); with many mini-batches, blocks in their draw phase
    // (Philox, shared-memory atomics) and blocks in their gather phase (memory) share an SM: 1.61 / 1.45 / 1.35e10
    // samples/s at 16384 x 256 (n = 1 / 5 / 10) against 1.28 / 1.23 / 1.22e10 for a draw launch followed by a
    // gather launch (and 1.31 / 1.26 / 1.25e10 with the warp-cooperative gather inside, whose shared memory
    // halves the resident blocks).
    if (d->state_dim < 32) {
      replay_draw_gather_kernel<<<(unsigned)n_batches, threads, smem, st>>>(
          *d, filled, batch, seed, draw_index, draw_counter, lanes, lane_len, table, out_idx, multi_steps, additive,
          gp, out_state, out_action, out_reward, out_next_state, out_done, out_eff, bump);
      B200_CUDA(cudaGetLastError());
      if (bumped && bump) *bumped = 1;
      return 0;
    }
    replay_draw_kernel<<<(unsigned)n_batches, threads, smem, st>>>(*d, filled, batch, seed, draw_index, draw_counter,
                                                                    lanes, lane_len, table, out_idx);
    idx = out_idx;
  }
  if (bulk_ok) {
    const int64_t want_blocks = (n_samples + 255) / 256;
    const int grid = (int)std::min<int64_t>(want_blocks, (int64_t)sm_count() * 4);
    replay_gather_bulk_kernel<<<grid, 256, 8 * BULK_WARP_FLOATS * sizeof(float), st>>>(
        *d, idx, n_samples, filled, lanes, lane_len, multi_steps, additive, gp, out_state, out_action, out_reward,
        out_next_state, out_done, out_eff);
  } else if (d->state_dim >= 32) {
    const int64_t want_blocks = (n_samples * 32 + 255) / 256;
    const int grid = (int)std::min<int64_t>(want_blocks, (int64_t)sm_count() * 8);
    replay_gather_kernel<32><<<grid, 256, 0, st>>>(*d, idx, n_samples, filled, lanes, lane_len, multi_steps, additive,
                                                    gp, out_state, out_action, out_reward, out_next_state, out_done,
                                                    out_eff);
  } else {
    const int64_t want_blocks = (n_samples + 255) / 256;
    const int grid = (int)std::min<int64_t>(want_blocks, (int64_t)sm_count() * 8);
    replay_gather_thread_kernel<<<grid, 256, 0, st>>>(*d, idx, n_samples, filled, lanes, lane_len, multi_steps,
                                                      additive, gp, out_state, out_action, out_reward, out_next_state,
                                                      out_done, out_eff);
  }
  B200_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace b200

extern "C" int b200_replay_sample(const b200_replay_desc* d, const int64_t* idx, int64_t n_batches, int32_t batch,
                                  int64_t filled, int32_t multi_steps, const float* gamma_pow_host,
                                  int32_t additive, uint64_t seed, uint64_t draw_index, int64_t* out_idx,
                                  float* out_state, float* out_action, float* out_reward, float* out_next_state,
                                  uint8_t* out_done, int64_t* out_eff, void* stream) {
  if (d == nullptr) return set_error(B200_EINVAL, "replay: desc is NULL");
  return replay_sample_lanes(d, 1, d->mem_size, idx, n_batches, batch, filled, multi_steps, gamma_pow_host, additive,
                             seed, draw_index, nullptr, out_idx, out_state, out_action, out_reward, out_next_state,
                             out_done, out_eff, stream, nullptr, nullptr);
}


namespace b200 {
__global__ void replay_counter_bump_kernel(int64_t* __restrict__ counter) { counter[0] += 1; }
}  // namespace b200

extern "C" int b200_replay_sample_counted(const b200_replay_desc* d, int64_t n_batches, int32_t batch,
                                          int32_t multi_steps, const float* gamma_pow_host, int32_t additive,
                                          uint64_t seed, int64_t* counter, int64_t* out_idx, float* out_state,
                                          float* out_action, float* out_reward, float* out_next_state,
                                          uint8_t* out_done, int64_t* out_eff, void* stream) {
  if (d == nullptr) return set_error(B200_EINVAL, "replay: desc is NULL");
  B200_REQUIRE(counter != nullptr, "replay_sample_counted: counter is NULL");
  int bumped = 0;
  int rc = replay_sample_lanes(d, 1, d->mem_size, nullptr, n_batches, batch, -1, multi_steps, gamma_pow_host, additive,
                               seed, 0, counter, out_idx, out_state, out_action, out_reward, out_next_state, out_done,
                               out_eff, stream, counter, &bumped);
  if (rc) return rc;
  if (n_batches * batch > 0 && !bumped) {
    replay_counter_bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter);
    return check_cuda(cudaGetLastError(), "replay_sample_counted counter");
  }
  return 0;
}
