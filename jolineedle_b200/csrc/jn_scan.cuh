// K3 -- segmented scans: per-episode returns and supervised trajectory expansion.
#pragma once

#include "jn_device.cuh"

namespace jnk {

// ------------------------------------------------------------------------------------------
// returns (reinforce.py:186-202)
// ------------------------------------------------------------------------------------------
// One thread per episode, reading the step-major [T, n] buffers the env wrote (coalesced
// across the warp) and producing the reference's episode-major tensors.  The scan runs from
// the last step backwards with a float64 accumulator that is rounded to float32 once per
// element -- the arithmetic of torch's CPU cumsum on the flipped, masked rewards.  The order
// of additions is the reference's, so the result is bit-identical, not merely close.
__global__ void returns_kernel(const float* __restrict__ rewards_tn, const uint8_t* __restrict__ terminated_tn, int T,
                               int n, float* __restrict__ rewards_out, uint8_t* __restrict__ masks,
                               uint8_t* __restrict__ logit_masks, float* __restrict__ returns) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double acc = 0.0;
  // blocks of 8 steps from the last one backwards: the loads of a block are independent of the running sum and
  // go out together; only the additions are serial
  constexpr int kBlock = 8;
  for (int hi = T; hi > 0; hi -= kBlock) {
    const int lo = hi > kBlock ? hi - kBlock : 0;
    float r[kBlock];
    uint8_t term[kBlock + 1];  // term[k] = terminated after step lo + k - 1 (slot 0: the step before the block)
#pragma unroll
    for (int k = 0; k < kBlock; ++k)
      if (lo + k < hi) r[k] = rewards_tn[(long long)(lo + k) * n + e];
#pragma unroll
    for (int k = 0; k <= kBlock; ++k)
      if (lo + k <= hi) term[k] = (lo + k == 0) ? 0 : terminated_tn[(long long)(lo + k - 1) * n + e];
#pragma unroll
    for (int k = kBlock - 1; k >= 0; --k) {
      const int t = lo + k;
      if (t >= hi) continue;
      // logit_masks[:, t] = masks[:, t]: column 0 is True, column t = !terminated after step t-1
      const uint8_t alive = term[k] ? 0 : 1;
      const float product = __fmul_rn(r[k], alive ? 1.0f : 0.0f);  // fp32 product first (keeps -0.0 / NaN behaviour)
      acc += (double)product;
      returns[(long long)e * T + t] = (float)acc;
      rewards_out[(long long)e * T + t] = r[k];
      logit_masks[(long long)e * T + t] = alive;
      masks[(long long)e * (T + 1) + t + 1] = term[k + 1] ? 0 : 1;
    }
  }
  masks[(long long)e * (T + 1)] = 1;
}

__global__ void returns_rows_kernel(const float* __restrict__ rewards, long long r_stride,
                                    const uint8_t* __restrict__ logit_masks, long long m_stride, int T, int n,
                                    float* __restrict__ returns) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double acc = 0.0;
  for (int t = T - 1; t >= 0; --t) {
    const float term = __fmul_rn(rewards[e * r_stride + t], logit_masks[e * m_stride + t] ? 1.0f : 0.0f);
    acc += (double)term;
    returns[(long long)e * T + t] = (float)acc;
  }
}

// ------------------------------------------------------------------------------------------
// tile lookup: which detection patches were already gathered as trajectory glimpses?
// ------------------------------------------------------------------------------------------
// The detection patches of an image are its box patches plus one empty patch (simple_env.py:397-419)
// and the trajectory visits exactly those box patches, so most detection tiles already sit in the
// [n, T, C, P, P] trajectory buffer.  For query d (episode e = q_src[d], patch q_pos[d]) this finds a
// recorded slot t of the same episode at the same patch and redirects the query to image
// `slab_base + e*T + t`, patch (0, 0) -- the trajectory buffer registered as a slab of n*T one-patch
// images.  When the images live in pinned host memory this keeps those tiles off PCIe.
__global__ void tile_lookup_kernel(const int64_t* __restrict__ traj_pos, const int32_t* __restrict__ traj_src, int T,
                                   const int64_t* __restrict__ q_pos, const int32_t* __restrict__ q_src, int n_queries,
                                   int slab_base, int64_t* __restrict__ out_pos, int32_t* __restrict__ out_src) {
  for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < n_queries; d += gridDim.x * blockDim.x) {
    const int e = q_src[d];
    long long y = q_pos[2 * (long long)d], x = q_pos[2 * (long long)d + 1];
    int src = e;
    if (e >= 0) {
      for (int t = 0; t < T; ++t) {
        const long long k = (long long)e * T + t;
        if (traj_src[k] == e && traj_pos[2 * k] == y && traj_pos[2 * k + 1] == x) {
          src = slab_base + (int)k;
          y = 0; x = 0;
          break;
        }
      }
    }
    out_pos[2 * (long long)d] = y; out_pos[2 * (long long)d + 1] = x;
    out_src[d] = src;
  }
}

// First occurrences and repeats among the recorded slots of each episode (a walk with detours
// revisits patches: ~10 % of the slots at BASELINE cfg 2).  Slot k = e*T + t whose patch was already
// recorded at an earlier slot t' of the same episode gets first_src = -2 (skip) and repeat_src = e*T + t'
// (the earliest such slot, itself a first occurrence); every other slot keeps its source in first_src
// (-1 = padded slot, zero-filled) and gets repeat_src = -2.  Two gathers then fill the buffer: first
// occurrences out of the images, repeats out of the buffer itself -- with host-resident images a
// revisited tile crosses PCIe once.
__global__ void tile_dedupe_kernel(const int64_t* __restrict__ traj_pos, const int32_t* __restrict__ traj_src,
                                   int n_slots, int T, int32_t* __restrict__ first_src,
                                   int32_t* __restrict__ repeat_src) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_slots; k += gridDim.x * blockDim.x) {
    const int e = traj_src[k];
    int first = e, repeat = -2;
    if (e >= 0) {
      const long long y = traj_pos[2 * (long long)k], x = traj_pos[2 * (long long)k + 1];
      const int k0 = k - k % T;
      for (int j = k0; j < k; ++j)
        if (traj_src[j] == e && traj_pos[2 * (long long)j] == y && traj_pos[2 * (long long)j + 1] == x) {
          first = -2; repeat = j;
          break;
        }
    }
    first_src[k] = first;
    repeat_src[k] = repeat;
  }
}

// ------------------------------------------------------------------------------------------
// trajectory expansion (simple_env.py:481-664)
// ------------------------------------------------------------------------------------------
// One warp per episode.  The host planner supplies the key points in visiting order as
// segments (walk to `to`, best action points at `tgt`) and the pre-drawn replacement moves.
//
//   phase 1  lanes = segments: length L_k = Chebyshev distance from the previous end point,
//            number of replacement draws the segment consumes (one if the group's opening
//            "best action" is STOP, one if the walk passes over `tgt`), warp scans of both.
//   phase 2  lanes = records: record t >= 1 belongs to the segment with pre_k < t <= pre_k+L_k;
//            its position is closed form, start + sign(d) * min(j, |d|) per axis.
//   phase 3  tail truncation: with ep_len > T only records [ep_len - T, ep_len) are kept.
//
// Segment tables live in shared memory, kMaxSeg segments per warp at a time: longer episodes (a box that
// covers hundreds of patches of a large grid) are expanded block by block -- a first pass over all segments
// yields the episode length (only the LAST T records are kept, so it has to be known first), then every block
// of kMaxSeg segments fills the table and writes the kept records it owns.  A record's `next_actions`
// overwrite may come from a later block than its position; blocks run in order and a slot is always handled
// by the same lane, so the later write wins as it does in the reference.
constexpr int kMaxSeg = 128;
constexpr int kTrajWarps = 4;

struct SegTable {
  int16_t ay[kMaxSeg], ax[kMaxSeg];  // start of the walk
  int16_t by[kMaxSeg], bx[kMaxSeg];  // end of the walk (to_visit)
  int16_t ty[kMaxSeg], tx[kMaxSeg];  // true target
  int32_t pre[kMaxSeg];              // records before this segment's first step, minus the start record
  int32_t len[kMaxSeg];
  int32_t draw0[kMaxSeg];            // first draw index of the segment
  int16_t hit[kMaxSeg];              // step j (1-based) at which the walk stands on tgt, 0 = never
  uint8_t first[kMaxSeg];            // opens a keypoint group
  uint8_t open_draw[kMaxSeg];        // the group-opening best action is STOP (consumes draw0)
};

__device__ __forceinline__ int sgn(int v) { return (v > 0) - (v < 0); }

// Phase 1 for segments [k0, k1) of an episode (a warp; lanes = segments, 32 at a time): Chebyshev lengths and
// draw counts with running warp scans.  `table` null: totals only.  run_len / run_draw carry the prefix in and out.
__device__ __forceinline__ void scan_segments(SegTable* table, const int32_t* __restrict__ seg_to,
                                              const int32_t* __restrict__ seg_tgt,
                                              const uint8_t* __restrict__ seg_flags, int s0, int k0, int k1, int sy,
                                              int sx, int lane, int& run_len, int& run_draw) {
  for (int base = k0; base < k1; base += 32) {
    const int k = base + lane;
    int L = 0, dcount = 0, hit = 0, ay = 0, ax = 0, by = 0, bx = 0, ty = 0, tx = 0, first = 0, open_draw = 0;
    if (k < k1) {
      by = seg_to[2 * (s0 + k)]; bx = seg_to[2 * (s0 + k) + 1];
      ty = seg_tgt[2 * (s0 + k)]; tx = seg_tgt[2 * (s0 + k) + 1];
      if (k == 0) { ay = sy; ax = sx; } else { ay = seg_to[2 * (s0 + k - 1)]; ax = seg_to[2 * (s0 + k - 1) + 1]; }
      first = seg_flags[s0 + k] & 1;
      const int dy = by - ay, dx = bx - ax;
      L = imax(iabs(dy), iabs(dx));
      open_draw = (first && ay == ty && ax == tx) ? 1 : 0;
      // the walk stands on tgt at step j iff both axes agree; positions are monotone per axis
      for (int j = 1; j <= L; ++j) {
        const int py = ay + sgn(dy) * imin(j, iabs(dy)), px = ax + sgn(dx) * imin(j, iabs(dx));
        if (py == ty && px == tx) { hit = j; break; }
      }
      dcount = open_draw + (hit ? 1 : 0);
    }
    // inclusive warp scans
    int incl_len = L, incl_draw = dcount;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int a = __shfl_up_sync(0xffffffffu, incl_len, off), b = __shfl_up_sync(0xffffffffu, incl_draw, off);
      if (lane >= off) { incl_len += a; incl_draw += b; }
    }
    if (table && k < k1) {
      SegTable& s = *table;
      const int i = k - k0;
      s.ay[i] = (int16_t)ay; s.ax[i] = (int16_t)ax; s.by[i] = (int16_t)by; s.bx[i] = (int16_t)bx;
      s.ty[i] = (int16_t)ty; s.tx[i] = (int16_t)tx;
      s.len[i] = L; s.pre[i] = run_len + incl_len - L;
      s.draw0[i] = run_draw + incl_draw - dcount;
      s.hit[i] = (int16_t)hit; s.first[i] = (uint8_t)first; s.open_draw[i] = (uint8_t)open_draw;
    }
    run_len += __shfl_sync(0xffffffffu, incl_len, 31);
    run_draw += __shfl_sync(0xffffffffu, incl_draw, 31);
  }
}

__global__ void __launch_bounds__(kTrajWarps * 32)
traj_expand_kernel(const int32_t* __restrict__ start_yx, const int32_t* __restrict__ seg_begin,
                   const int32_t* __restrict__ seg_to, const int32_t* __restrict__ seg_tgt,
                   const uint8_t* __restrict__ seg_flags, const int32_t* __restrict__ draw_begin,
                   const uint8_t* __restrict__ draws, const uint32_t* __restrict__ area, int words_per_item,
                   const int32_t* __restrict__ cols_arr, int n, int T, int64_t* __restrict__ positions,
                   int64_t* __restrict__ cur_act, int64_t* __restrict__ next_act, int64_t* __restrict__ labels,
                   float* __restrict__ masks, int32_t* __restrict__ gather_src, int32_t* __restrict__ ep_len_out,
                   int32_t* __restrict__ status) {
  __shared__ SegTable tables[kTrajWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  SegTable& s = tables[warp];
  for (int e = blockIdx.x * kTrajWarps + warp; e < n; e += gridDim.x * kTrajWarps) {
    const int s0 = seg_begin[e], ns = seg_begin[e + 1] - s0;
    const int d0 = draw_begin[e], nd = draw_begin[e + 1] - d0;
    const int sy = start_yx[2 * e], sx = start_yx[2 * e + 1];
    const int cols = cols_arr[e];
    const uint32_t* bm = area + (long long)e * words_per_item;
    const long long o = (long long)e * T;

    // ---- phase 1: episode length and draw count (episodes of up to kMaxSeg segments fill the table right away)
    const bool one_block = ns <= kMaxSeg;
    int total_len = 0, total_draw = 0;
    scan_segments(one_block ? &s : nullptr, seg_to, seg_tgt, seg_flags, s0, 0, ns, sy, sx, lane, total_len, total_draw);
    __syncwarp();
    const int ep_len = 1 + total_len;
    if (total_draw > nd) {  // planner supplied too few replacement moves: the episode is left masked out
      if (lane == 0 && status) atomicOr(status, 8);
      for (int t = lane; t < T; t += 32) {
        positions[2 * (o + t)] = 0; positions[2 * (o + t) + 1] = 0;
        cur_act[o + t] = 0; next_act[o + t] = 0; labels[o + t] = 0; masks[o + t] = 0.f; gather_src[o + t] = -1;
      }
      if (lane == 0) ep_len_out[e] = 0;
      __syncwarp();
      continue;
    }

    // ---- phases 2+3: one lane per kept record, one block of segments at a time
    const int drop = ep_len > T ? ep_len - T : 0;  // keep the LAST T records (simple_env.py:580-584)
    int run_len = 0, run_draw = 0;
    for (int k0 = 0; k0 == 0 || k0 < ns; k0 += kMaxSeg) {
      const int k1 = imin(ns, k0 + kMaxSeg), nb = k1 - k0;
      if (!one_block) {
        __syncwarp();  // the previous block's table has been read by every lane
        scan_segments(&s, seg_to, seg_tgt, seg_flags, s0, k0, k1, sy, sx, lane, run_len, run_draw);
        __syncwarp();
      }
      for (int slot = lane; slot < T; slot += 32) {
        const int t = slot + drop;  // record index in the untruncated episode
        if (t >= ep_len) {
          if (k0 == 0) {
            positions[2 * (o + slot)] = 0; positions[2 * (o + slot) + 1] = 0;
            cur_act[o + slot] = 0; next_act[o + slot] = 0; labels[o + slot] = 0; masks[o + slot] = 0.f;
            gather_src[o + slot] = -1;
          }
          continue;
        }
        // owning segment: records of segment k are t = pre_k + j, j = 1..len_k (the start record t = 0 has none)
        // overwrite: the last group-opening segment with pre_k == t rewrites next_actions[t]
        int own = -1, over = -1;
        for (int i = 0; i < nb; ++i) {
          const int pre = s.pre[i];
          if (t > pre && t <= pre + s.len[i]) own = i;
          if (s.first[i] && pre == t) over = i;
        }
        const bool start_record = (t == 0 && k0 == 0);
        if (own >= 0 || start_record) {
          int py = sy, px = sx, act = 0 /*LEFT placeholder, simple_env.py:527-534*/, best = 0;
          if (own >= 0) {
            const int j = t - s.pre[own];
            const int dy = s.by[own] - s.ay[own], dx = s.bx[own] - s.ax[own];
            const int qy = s.ay[own] + sgn(dy) * imin(j - 1, iabs(dy)), qx = s.ax[own] + sgn(dx) * imin(j - 1, iabs(dx));
            py = s.ay[own] + sgn(dy) * imin(j, iabs(dy)); px = s.ax[own] + sgn(dx) * imin(j, iabs(dx));
            act = direction_code(s.by[own] - qy, s.bx[own] - qx);
            best = direction_code(s.ty[own] - py, s.tx[own] - px);
            if (best == kStop) best = draws[d0 + s.draw0[own] + s.open_draw[own]];  // j == hit[own]
          }
          const int bit = py * cols + px;
          positions[2 * (o + slot)] = py; positions[2 * (o + slot) + 1] = px;
          cur_act[o + slot] = act; next_act[o + slot] = best;
          labels[o + slot] = (bm[bit >> 5] >> (bit & 31)) & 1u;
          masks[o + slot] = 1.0f;
          gather_src[o + slot] = e;
        }
        if (over >= 0) {
          int best = direction_code(s.ty[over] - s.ay[over], s.tx[over] - s.ax[over]);
          if (best == kStop) best = draws[d0 + s.draw0[over]];
          next_act[o + slot] = best;
        }
      }
    }
    if (lane == 0) ep_len_out[e] = ep_len;
    __syncwarp();
  }
}

}  // namespace jnk
