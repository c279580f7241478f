// K0 (patch x bbox tables) and K2 (warp-per-episode reset / step / props).
//
// Bitmap convention: patch (y, x) of an episode with `cols` columns is bit (y*cols + x) & 31
// of word (y*cols + x) >> 5.  A warp owns one episode; lane l holds words l, l+32, ... in
// registers while it works, so a 32x32 grid (1024 patches, BASELINE cfg 4) is exactly one
// word per lane and every count is a popc + warp reduction.
#pragma once

#include "jn_device.cuh"

namespace jnk {

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(kFullMask, v); }

// ------------------------------------------------------------------------------------------
// K0: overlap bitmaps
// ------------------------------------------------------------------------------------------
// rule 0 (any pixel): general_env.py:360-379 -- box rasterised with inclusive x2/y2 after
//   clamping to the image, max-pooled by P: patch set iff it meets [x1c, x2c) x [y1c, y2c).
// rule 1 (area 5%): simple_env.py:270-321 -- patch set iff it lies in the box's patch range
//   and 20 * overlap_area > P^2 (the float test `area / P**2 > 0.05` in exact integers), or
//   it holds the box centre; always restricted to the grid.
//
// One warp per episode.  Boxes are taken one at a time (their patch ranges are warp-uniform scalars); for the
// few bitmap words a box can touch, lane = bit evaluates its patch and a ballot assembles the word, which
// lane (word % 32) ORs into its accumulator -- so a 5x6 grid keeps 30 lanes busy instead of one, and a
// 32x32 grid only visits the one or two words under each box.
__global__ void patch_bitmaps_kernel(const int64_t* __restrict__ bboxes, const int32_t* __restrict__ n_boxes, int n,
                                     int max_boxes, int P, int grid_rows, int grid_cols,
                                     const int32_t* __restrict__ rows_arr, const int32_t* __restrict__ cols_arr,
                                     int rule, uint32_t* __restrict__ out, int words_per_item) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int e = blockIdx.x * warps_per_block + (threadIdx.x >> 5); e < n; e += gridDim.x * warps_per_block) {
    const int rows = rows_arr ? rows_arr[e] : grid_rows;
    const int cols = cols_arr ? cols_arr[e] : grid_cols;
    const int nb = n_boxes ? n_boxes[e] : max_boxes;
    const long long H = (long long)rows * P, W = (long long)cols * P;
    const int n_bits = rows * cols;
    for (int wb = 0; wb < words_per_item; wb += 32) {  // blocks of 32 words: lane l accumulates word wb + l
      uint32_t acc = 0;
      for (int k = 0; k < nb; ++k) {
        const int64_t* b = bboxes + ((long long)e * max_boxes + k) * 4;
        const long long x1 = b[0], y1 = b[1], x2 = b[2], y2 = b[3];
        long long px_lo, px_hi, py_lo, py_hi;  // inclusive candidate patch range
        long long cpx = -1, cpy = -1;          // centre patch (rule 1)
        if (rule == 0) {
          const long long x1c = lmin(lmax(x1, 0), W), x2c = lmin(lmax(x2 + 1, 0), W);
          const long long y1c = lmin(lmax(y1, 0), H), y2c = lmin(lmax(y2 + 1, 0), H);
          if (x1c >= x2c || y1c >= y2c) continue;
          px_lo = x1c / P; px_hi = (x2c - 1) / P; py_lo = y1c / P; py_hi = (y2c - 1) / P;
        } else {
          px_lo = floordiv(x1, P); px_hi = floordiv(x2, P); py_lo = floordiv(y1, P); py_hi = floordiv(y2, P);
          cpx = floordiv(floordiv(x1 + x2, 2), P); cpy = floordiv(floordiv(y1 + y2, 2), P);
        }
        // bits the box can set: its patch range clipped to the grid, plus the centre patch when that is in the grid
        const long long cx_lo = lmax(px_lo, 0), cx_hi = lmin(px_hi, cols - 1);
        const long long cy_lo = lmax(py_lo, 0), cy_hi = lmin(py_hi, rows - 1);
        long long idx_lo = n_bits, idx_hi = -1;
        if (cx_lo <= cx_hi && cy_lo <= cy_hi) { idx_lo = cy_lo * cols + cx_lo; idx_hi = cy_hi * cols + cx_hi; }
        if (rule != 0 && cpx >= 0 && cpx < cols && cpy >= 0 && cpy < rows) {
          idx_lo = lmin(idx_lo, cpy * cols + cpx); idx_hi = lmax(idx_hi, cpy * cols + cpx);
        }
        if (idx_hi < idx_lo) continue;
        const int w_lo = imax((int)(idx_lo >> 5), wb), w_hi = imin((int)(idx_hi >> 5), imin(wb + 31, words_per_item - 1));
        for (int w = w_lo; w <= w_hi; ++w) {
          const int idx = w * 32 + lane;
          bool hit = false;
          if (idx < n_bits) {
            const int y = idx / cols, x = idx - y * cols;
            if (rule == 0) {
              hit = (x >= px_lo && x <= px_hi && y >= py_lo && y <= py_hi);
            } else {
              hit = (x == cpx && y == cpy);
              if (!hit && x >= px_lo && x <= px_hi && y >= py_lo && y <= py_hi) {
                const long long oh = lmin((long long)(y + 1) * P, y2) - lmax((long long)y * P, y1);
                const long long ow = lmin((long long)(x + 1) * P, x2) - lmax((long long)x * P, x1);
                hit = 20 * (oh * ow) > (long long)P * P;
              }
            }
          }
          const uint32_t word = __ballot_sync(kFullMask, hit);
          if (lane == w - wb) acc |= word;
        }
      }
      if (wb + lane < words_per_item) out[(long long)e * words_per_item + wb + lane] = acc;
    }
  }
}

__global__ void bitmap_unpack_kernel(const uint32_t* __restrict__ words, int n, int bits_per_item, int words_per_item,
                                     uint8_t* __restrict__ out) {
  const long long total = (long long)n * bits_per_item;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i / bits_per_item), bit = (int)(i - (long long)e * bits_per_item);
    out[i] = (words[(long long)e * words_per_item + (bit >> 5)] >> (bit & 31)) & 1u;
  }
}

// parse_bboxes (general_env.py:381-504) in closed form: box k contributes to every patch of
// [x1//P .. x2//P] x [y1//P .. y2//P] its intersection with that patch, in local inclusive
// coordinates.  One thread per (episode, patch, box).  Degenerate boxes (x2 < x1 or y2 < y1)
// follow the recursion's behaviour: only the top-left patch is written, un-clamped below.
__global__ void split_boxes_kernel(const int64_t* __restrict__ bboxes, int n, int max_boxes, int P, int rows, int cols,
                                   int64_t* __restrict__ local, uint8_t* __restrict__ present,
                                   int32_t* __restrict__ status) {
  const long long total = (long long)n * rows * cols * max_boxes;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % max_boxes);
    long long t = i / max_boxes;
    const int x = (int)(t % cols); t /= cols;
    const int y = (int)(t % rows);
    const int e = (int)(t / rows);
    const int64_t* b = bboxes + ((long long)e * max_boxes + k) * 4;
    // `.int()` truncation of the reference is a no-op for int64 inputs in range
    const long long x1 = b[0], y1 = b[1], x2 = b[2], y2 = b[3];
    const long long px1 = floordiv(x1, P), py1 = floordiv(y1, P);
    const long long px2 = lmax(floordiv(x2, P), px1), py2 = lmax(floordiv(y2, P), py1);
    if (px1 < 0 || py1 < 0 || px2 >= cols || py2 >= rows) {
      if (status && x == 0 && y == 0) atomicOr(status, 4);
    }
    int64_t* o = local + i * 4;
    const bool hit = (x >= px1 && x <= px2 && y >= py1 && y <= py2);
    if (hit) {
      const long long ox = (long long)x * P, oy = (long long)y * P;
      o[0] = lmax(x1, ox) - ox;
      o[1] = lmax(y1, oy) - oy;
      o[2] = lmin(x2 - ox, (long long)P - 1);
      o[3] = lmin(y2 - oy, (long long)P - 1);
    } else {
      o[0] = o[1] = o[2] = o[3] = 0;
    }
    present[i] = hit ? 1 : 0;
  }
}

// local_bboxes (simple_env.py:231-268): one thread per (item, box).
__global__ void local_boxes_kernel(const int64_t* __restrict__ bboxes, const int32_t* __restrict__ n_boxes,
                                   int max_boxes, int P, const int64_t* __restrict__ positions,
                                   const int32_t* __restrict__ src_index, int n_items, float* __restrict__ out) {
  const long long total = (long long)n_items * max_boxes;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int item = (int)(i / max_boxes), k = (int)(i - (long long)item * max_boxes);
    float* o = out + i * 6;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f, v5 = 0.f;
    const int e = src_index ? src_index[item] : item;
    if (e >= 0 && k < (n_boxes ? n_boxes[e] : max_boxes)) {
      const int64_t* b = bboxes + ((long long)e * max_boxes + k) * 4;
      const long long px1 = positions[2 * (long long)item + 1] * P, py1 = positions[2 * (long long)item] * P;
      const long long px2 = px1 + P, py2 = py1 + P;
      const long long x1 = lmax(px1, b[0]), y1 = lmax(py1, b[1]);
      const long long x2 = lmin(px2, b[2]), y2 = lmin(py2, b[3]);
      if (x1 < x2 && y1 < y2) {  // the px1 <= x1 and x2 <= px2 halves hold by construction
        v1 = (float)(x1 - px1); v2 = (float)(y1 - py1); v3 = (float)(x2 - px1); v4 = (float)(y2 - py1); v5 = 1.f;
      }
    }
    o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3; o[4] = v4; o[5] = v5;
  }
}

// ------------------------------------------------------------------------------------------
// K2: reset / step / props -- one warp per episode
// ------------------------------------------------------------------------------------------
__global__ void env_reset_kernel(const int64_t* __restrict__ positions, uint32_t* __restrict__ visited,
                                 int64_t* __restrict__ steps, uint8_t* __restrict__ has_stopped, int n, int rows,
                                 int cols, int words, int32_t* __restrict__ status) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int e = blockIdx.x * warps_per_block + (threadIdx.x >> 5); e < n; e += gridDim.x * warps_per_block) {
    const long long y = positions[2 * (long long)e], x = positions[2 * (long long)e + 1];
    const bool ok = (y >= 0 && y < rows && x >= 0 && x < cols);
    const int bit = ok ? (int)(y * cols + x) : -1;
    for (int w = lane; w < words; w += 32)
      visited[(long long)e * words + w] = (bit >= 0 && (bit >> 5) == w) ? (1u << (bit & 31)) : 0u;
    if (lane == 0) {
      steps[e] = 0;
      has_stopped[e] = 0;
      if (!ok && status) atomicOr(status, 1);
    }
  }
}

// A group of G lanes (G = the power of two >= the number of bitmap words, at most 32) owns an episode:
// lane k of the group owns words k, k+G, ...; a warp steps 32/G episodes at once (32x32 grid: 32 words, one
// warp per episode; 8x8 grid: 2 words, 16 episodes per warp).  Counts are reduced inside the group with
// xor-shuffles.  Every lane of a group computes the scalar tail (same values, uniform instructions) and lane 0
// of the group stores, so that no instruction runs with a single active lane.
template <int G>
__global__ void env_step_group_kernel(const int64_t* __restrict__ pos_in, const int64_t* __restrict__ actions,
                                      int64_t* __restrict__ pos_out, uint32_t* __restrict__ visited,
                                      const uint32_t* __restrict__ bbox, int64_t* __restrict__ steps,
                                      uint8_t* __restrict__ has_stopped, float* __restrict__ rewards,
                                      uint8_t* __restrict__ terminated, uint8_t* __restrict__ truncated, int n,
                                      int rows, int cols, int words, int max_ep_len, float cost, int stop_enabled,
                                      int32_t* __restrict__ status) {
  constexpr int kPerWarp = 32 / G;
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31, sub = lane & (G - 1), group = lane / G;
  const int stride = gridDim.x * warps_per_block * kPerWarp;
  for (int base = (blockIdx.x * warps_per_block + (threadIdx.x >> 5)) * kPerWarp; base < n; base += stride) {
    const int e = base + group;
    const bool live = e < n;  // lanes past the end still take part in the shuffles
    // --- move + clamp, sticky stop (general_env.py:209-233)
    const long long a_raw = live ? actions[e] : (long long)kStop;
    const bool bad_action = a_raw < 0 || a_raw > kStop;
    if (bad_action && sub == 0 && status) atomicOr(status, 2);
    const long long a = bad_action ? (long long)kStop : a_raw;  // invalid code: no move (flagged); the reference raises
    long long y = (live ? pos_in[2 * (long long)e] : 0) + kActionDy[a];
    long long x = (live ? pos_in[2 * (long long)e + 1] : 0) + kActionDx[a];
    y = lmin(lmax(y, 0), rows - 1);
    x = lmin(lmax(x, 0), cols - 1);
    const bool stopped = live && ((has_stopped[e] != 0) || a_raw == kStop);
    const int bit = (int)(y * cols + x);
    // --- bitmaps
    int found = 0, every = 0, missing_after = 0, fresh = 0;
    if (live) {
      for (int w = sub; w < words; w += G) {
        const uint32_t v = visited[(long long)e * words + w];
        const uint32_t b = bbox[(long long)e * words + w];
        found += __popc(v & b);  // counts use the map BEFORE marking (general_env.py:347)
        every += __popc(b);
        uint32_t v_new = v;
        if ((bit >> 5) == w) {
          const uint32_t m = 1u << (bit & 31);
          fresh = ((b & m) != 0 && (v & m) == 0) ? 1 : 0;
          v_new = v | m;
          visited[(long long)e * words + w] = v_new;
        }
        missing_after += __popc(b & ~v_new);
      }
    }
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
      found += __shfl_xor_sync(kFullMask, found, off);
      every += __shfl_xor_sync(kFullMask, every, off);
      missing_after += __shfl_xor_sync(kFullMask, missing_after, off);
      fresh |= __shfl_xor_sync(kFullMask, fresh, off);
    }
    // --- reward = fl32(fl32(fresh + cost) + stop_eval), general_env.py:334-358
    float r = __fadd_rn(fresh ? 1.0f : 0.0f, cost);
    if (stop_enabled) {
      const int stop_eval = stopped ? (found == every ? found : found - every) : 0;
      r = __fadd_rn(r, (float)stop_eval);
    }
    const long long s = (live ? steps[e] : 0) + 1;
    if (live && sub == 0) {
      rewards[e] = r;
      steps[e] = s;
      has_stopped[e] = stopped ? 1 : 0;
      truncated[e] = s >= max_ep_len ? 1 : 0;
      terminated[e] = stop_enabled ? (stopped ? 1 : 0) : (missing_after == 0 ? 1 : 0);
      pos_out[2 * (long long)e] = y;
      pos_out[2 * (long long)e + 1] = x;
    }
  }
}

// Single-word grids (rows*cols <= 32, e.g. the 5x6 LARD grid of cfg 2/3): the whole bitmap of an
// episode is one register, so a *lane* owns an episode and a warp steps 32 of them with every lane
// busy with 16-byte position loads and stores.  Same arithmetic, same order of operations as the group
// kernel above (whose G = 1 instance serves one-word grids with unaligned position buffers).
__global__ void env_step_lane_kernel(const int64_t* __restrict__ pos_in, const int64_t* __restrict__ actions,
                                     int64_t* __restrict__ pos_out, uint32_t* __restrict__ visited,
                                     const uint32_t* __restrict__ bbox, int64_t* __restrict__ steps,
                                     uint8_t* __restrict__ has_stopped, float* __restrict__ rewards,
                                     uint8_t* __restrict__ terminated, uint8_t* __restrict__ truncated, int n, int rows,
                                     int cols, int max_ep_len, float cost, int stop_enabled,
                                     int32_t* __restrict__ status) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    long long a = actions[e];
    const bool valid = (a >= 0 && a <= kStop);
    if (!valid) {
      if (status) atomicOr(status, 2);
      a = kStop;
    }
    const longlong2 p = reinterpret_cast<const longlong2*>(pos_in)[e];  // (y, x), 16-byte aligned rows
    const long long y = lmin(lmax(p.x + kActionDy[a], 0), rows - 1);
    const long long x = lmin(lmax(p.y + kActionDx[a], 0), cols - 1);
    const bool stopped = (has_stopped[e] != 0) || (valid && a == kStop);
    const uint32_t m = 1u << (int)(y * cols + x);
    const uint32_t v = visited[e], b = bbox[e];
    const int found = __popc(v & b), every = __popc(b);  // from the map BEFORE marking
    const bool fresh = (b & m) != 0 && (v & m) == 0;
    const uint32_t v_new = v | m;
    visited[e] = v_new;
    float r = __fadd_rn(fresh ? 1.0f : 0.0f, cost);
    if (stop_enabled) r = __fadd_rn(r, (float)(stopped ? (found == every ? found : found - every) : 0));
    rewards[e] = r;
    const long long s = steps[e] + 1;
    steps[e] = s;
    has_stopped[e] = stopped ? 1 : 0;
    truncated[e] = s >= max_ep_len ? 1 : 0;
    terminated[e] = stop_enabled ? (stopped ? 1 : 0) : ((b & ~v_new) == 0 ? 1 : 0);
    reinterpret_cast<longlong2*>(pos_out)[e] = make_longlong2(y, x);
  }
}

// The `rewards` property of the reference (general_env.py:321-358) evaluated on the CURRENT state, outside
// of a step: fresh = bbox[pos] & ~visited[pos] with whatever `visited` holds now.  One thread per episode.
__global__ void env_rewards_kernel(const int64_t* __restrict__ positions, const uint32_t* __restrict__ visited,
                                   const uint32_t* __restrict__ bbox, const uint8_t* __restrict__ has_stopped, int n,
                                   int rows, int cols, int words, float cost, int stop_enabled,
                                   float* __restrict__ rewards) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const long long y = positions[2 * (long long)e], x = positions[2 * (long long)e + 1];
    int found = 0, every = 0;
    bool fresh = false;
    const int bit = (y >= 0 && y < rows && x >= 0 && x < cols) ? (int)(y * cols + x) : -1;
    for (int w = 0; w < words; ++w) {
      const uint32_t v = visited[(long long)e * words + w], b = bbox[(long long)e * words + w];
      found += __popc(v & b);
      every += __popc(b);
      if (bit >= 0 && (bit >> 5) == w) fresh = ((b >> (bit & 31)) & 1u) && !((v >> (bit & 31)) & 1u);
    }
    float r = __fadd_rn(fresh ? 1.0f : 0.0f, cost);
    if (stop_enabled) r = __fadd_rn(r, (float)(has_stopped[e] ? (found == every ? found : found - every) : 0));
    rewards[e] = r;
  }
}

__global__ void env_props_kernel(const uint32_t* __restrict__ visited, const uint32_t* __restrict__ bbox,
                                 const uint8_t* __restrict__ has_stopped, int n, int words, int stop_enabled,
                                 float* __restrict__ prop_patches, uint8_t* __restrict__ terminated) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int e = blockIdx.x * warps_per_block + (threadIdx.x >> 5); e < n; e += gridDim.x * warps_per_block) {
    int found = 0, every = 0;
    for (int w = lane; w < words; w += 32) {
      const uint32_t v = visited[(long long)e * words + w], b = bbox[(long long)e * words + w];
      found += __popc(v & b);
      every += __popc(b);
    }
    found = warp_sum(found);
    every = warp_sum(every);
    if (lane == 0) {
      // int64 / int64 true division in torch -> float32 operands (general_env.py:308-315)
      if (prop_patches) prop_patches[e] = __fdiv_rn((float)found, (float)(every == 0 ? 1 : every));
      if (terminated) terminated[e] = stop_enabled ? has_stopped[e] : (found == every ? 1 : 0);
    }
  }
}

// ------------------------------------------------------------------------------------------
// first-visit table of the zero-copy env (host-resident images + crop history)
// ------------------------------------------------------------------------------------------
// One thread per episode: has this episode seen its current patch before, and in which history slot?
__global__ void visit_sources_kernel(const int64_t* __restrict__ positions, int32_t* __restrict__ first_slot, int n,
                                     int rows, int cols, int slots, int t, int32_t* __restrict__ host_src,
                                     int32_t* __restrict__ history_src, int32_t* __restrict__ status) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const long long y = positions[2 * (long long)i], x = positions[2 * (long long)i + 1];
    if (y < 0 || x < 0 || y >= rows || x >= cols) {
      if (status) atomicOr(status, 1);  // out-of-grid position: the gather skips it too
      host_src[i] = -2; history_src[i] = -2;
      continue;
    }
    int32_t* cell = first_slot + (long long)i * rows * cols + y * cols + x;
    const int seen = *cell;
    if (seen < 0) {
      *cell = t;
      host_src[i] = i; history_src[i] = -2;
    } else {
      host_src[i] = -2; history_src[i] = i * slots + seen;
    }
  }
}

}  // namespace jnk
