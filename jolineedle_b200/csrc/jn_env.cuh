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
//
// Box coordinates are int64 pixels, or float64 (kFloat: rule 1 only) for boxes that are not whole pixels -- the
// dataset's minimum-size resize (dataset.py:258-270) scales them -- with the reference's python float arithmetic
// restated in IEEE doubles: floor(v / P) patch ranges, `oh * ow / P**2 > 0.05`, centre floor((a + b) / 2).
template <bool kFloat>
__global__ void patch_bitmaps_kernel(const void* __restrict__ bboxes_, const int32_t* __restrict__ n_boxes, int n,
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
        const long long box_at = ((long long)e * max_boxes + k) * 4;
        long long x1 = 0, y1 = 0, x2 = 0, y2 = 0;
        double fx1 = 0, fy1 = 0, fx2 = 0, fy2 = 0;
        if (kFloat) {
          const double* b = static_cast<const double*>(bboxes_) + box_at;
          fx1 = b[0]; fy1 = b[1]; fx2 = b[2]; fy2 = b[3];
        } else {
          const int64_t* b = static_cast<const int64_t*>(bboxes_) + box_at;
          x1 = b[0]; y1 = b[1]; x2 = b[2]; y2 = b[3];
        }
        long long px_lo, px_hi, py_lo, py_hi;  // inclusive candidate patch range
        long long cpx = -1, cpy = -1;          // centre patch (rule 1)
        if (kFloat) {
          const double dp = (double)P;
          px_lo = (long long)floor(__ddiv_rn(fx1, dp)); px_hi = (long long)floor(__ddiv_rn(fx2, dp));
          py_lo = (long long)floor(__ddiv_rn(fy1, dp)); py_hi = (long long)floor(__ddiv_rn(fy2, dp));
          cpx = (long long)floor(__ddiv_rn(floor(__ddiv_rn(__dadd_rn(fx1, fx2), 2.0)), dp));
          cpy = (long long)floor(__ddiv_rn(floor(__ddiv_rn(__dadd_rn(fy1, fy2), 2.0)), dp));
        } else if (rule == 0) {
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
                if (kFloat) {
                  const double oh = __dsub_rn(fmin((double)((long long)(y + 1) * P), fy2), fmax((double)((long long)y * P), fy1));
                  const double ow = __dsub_rn(fmin((double)((long long)(x + 1) * P), fx2), fmax((double)((long long)x * P), fx1));
                  hit = __ddiv_rn(__dmul_rn(oh, ow), (double)((long long)P * P)) > 0.05;
                } else {
                  const long long oh = lmin((long long)(y + 1) * P, y2) - lmax((long long)y * P, y1);
                  const long long ow = lmin((long long)(x + 1) * P, x2) - lmax((long long)x * P, x1);
                  hit = 20 * (oh * ow) > (long long)P * P;
                }
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

// local_bboxes (simple_env.py:231-268): one thread per (item, box).  kFloat: float64 boxes, differences taken in
// double and rounded to float32 once, as python floats going into a FloatTensor are.
template <bool kFloat>
__global__ void local_boxes_kernel(const void* __restrict__ bboxes_, const int32_t* __restrict__ n_boxes,
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
      const long long box_at = ((long long)e * max_boxes + k) * 4;
      const long long px1 = positions[2 * (long long)item + 1] * P, py1 = positions[2 * (long long)item] * P;
      const long long px2 = px1 + P, py2 = py1 + P;
      if (kFloat) {
        const double* b = static_cast<const double*>(bboxes_) + box_at;
        const double x1 = fmax((double)px1, b[0]), y1 = fmax((double)py1, b[1]);
        const double x2 = fmin((double)px2, b[2]), y2 = fmin((double)py2, b[3]);
        if (x1 < x2 && y1 < y2) {
          v1 = (float)__dsub_rn(x1, (double)px1); v2 = (float)__dsub_rn(y1, (double)py1);
          v3 = (float)__dsub_rn(x2, (double)px1); v4 = (float)__dsub_rn(y2, (double)py1); v5 = 1.f;
        }
      } else {
        const int64_t* b = static_cast<const int64_t*>(bboxes_) + box_at;
        const long long x1 = lmax(px1, b[0]), y1 = lmax(py1, b[1]);
        const long long x2 = lmin(px2, b[2]), y2 = lmin(py2, b[3]);
        if (x1 < x2 && y1 < y2) {  // the px1 <= x1 and x2 <= px2 halves hold by construction
          v1 = (float)(x1 - px1); v2 = (float)(y1 - py1); v3 = (float)(x2 - px1); v4 = (float)(y2 - py1); v5 = 1.f;
        }
      }
    }
    o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3; o[4] = v4; o[5] = v5;
  }
}

// ------------------------------------------------------------------------------------------
// K2: reset / step / props
// ------------------------------------------------------------------------------------------
// Per-step argument block of the env kernels (borrowed device pointers; see jn_env_step / jn_env_step_gather).
struct StepArgs {
  const int64_t* pos_in;   // [n, 2] positions before the move
  const int64_t* actions;  // [n]
  int64_t* pos_out;        // [n, 2] positions after the move (may alias pos_in)
  uint32_t* visited;       // [n, words]
  const uint32_t* bbox;    // [n, words]
  int64_t* steps;          // [n]
  uint8_t* has_stopped;    // [n]
  float* rewards;          // [n]
  uint8_t* terminated;     // [n]
  uint8_t* truncated;      // [n]
  // first-visit table of the zero-copy env (all null when unused): see visit_sources_kernel
  int32_t* first_slot;     // [n, rows*cols]
  int32_t* host_src;       // [n]
  int32_t* history_src;    // [n]
  unsigned long long* host_tiles;  // running count of first visits (= tiles read from the host), or null
  int32_t* status;
  int n, rows, cols, words, max_ep_len, stop_enabled, slots, t;
  float cost;
};

// Clears the episode state and marks the start patch (init_env_variables + the tail of reset,
// general_env.py:117-142,164).  One thread per bitmap word; with a first-visit table, one thread per patch of
// the table as well (entry = slot 0 for the start patch, -1 elsewhere) and the sources of slot 0.
__global__ void env_reset_kernel(const StepArgs a) {
  pdl_launch_dependents();
  const long long cells = a.first_slot ? (long long)a.rows * a.cols : 0;
  const long long per = a.words > cells ? a.words : cells;
  const long long total = (long long)a.n * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i / per), k = (int)(i - (long long)e * per);
    const long long y = a.pos_out[2 * (long long)e], x = a.pos_out[2 * (long long)e + 1];
    const bool ok = (y >= 0 && y < a.rows && x >= 0 && x < a.cols);
    const int bit = ok ? (int)(y * a.cols + x) : -1;
    if (k < a.words) a.visited[(long long)e * a.words + k] = (bit >= 0 && (bit >> 5) == k) ? (1u << (bit & 31)) : 0u;
    if (k < cells) a.first_slot[(long long)e * cells + k] = (k == bit) ? 0 : -1;
    if (k == 0) {
      a.steps[e] = 0;
      a.has_stopped[e] = 0;
      if (!ok && a.status) atomicOr(a.status, 1);
      if (a.first_slot) {
        a.host_src[e] = ok ? e : -2; a.history_src[e] = -2;
        if (ok && a.host_tiles) atomicAdd(a.host_tiles, 1ull);
      }
    }
  }
}

// One env step (general_env.py:172-207).  A warp owns 32 consecutive episodes and every instruction of the
// kernel runs with all of its lanes on useful work:
//
//   phase A  lane = episode: action, position (coalesced loads), move + clamp, sticky stop flag, bit index;
//   phase B  lanes = bitmap words: G lanes (G = the power of two >= the number of words, at most 32) read the
//            visited / bbox words of one episode, 32/G episodes at a time; popc counts are packed two to a
//            register and reduced inside the group (one REDUX per register when the group is the warp),
//            then handed to the lane that owns the episode.  Loads only -- the visited bit is set in phase C
//            with a RED.OR, so that the unrolled loop keeps every load of the warp in flight at once;
//   phase C  lane = episode: reward = fl32(fl32(fresh + cost) + stop_eval), flags, counters; coalesced stores.
//
// One-word grids (5x6 LARD grid, G = 1) degenerate to a lane per episode; the 32x32 aerial grid (G = 32) to 32
// warp-wide iterations; wider grids (kWide) keep a second block of 32 words in registers and loop over any
// further ones.  Counts are 16-bit fields: grids of up to 65535 patches.
template <int G, bool kWide = false>  // kWide: more than 32 bitmap words (G == 32)
__global__ void __launch_bounds__(64) env_step_kernel(const __grid_constant__ StepArgs a) {
  pdl_launch_dependents();  // the gather that follows derives its positions from pos_in + actions itself
  constexpr int kPer = 32 / G;
  const int lane = threadIdx.x & 31, sub = lane & (G - 1), grp = lane / G;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const bool vec = ((reinterpret_cast<uintptr_t>(a.pos_in) | reinterpret_cast<uintptr_t>(a.pos_out)) & 15) == 0;
  for (int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; base < a.n; base += n_warps * 32) {
    // ---- phase A
    const int e = base + lane;
    const bool live = e < a.n;
    long long act = live ? a.actions[e] : (long long)kStop;
    const bool bad = act < 0 || act > kStop;  // invalid code: flagged, no move (the reference raises)
    if (bad) {
      if (a.status) atomicOr(a.status, 2);
      act = kStop;
    }
    long long y = 0, x = 0;
    if (live) {
      if (vec) {
        const longlong2 p = reinterpret_cast<const longlong2*>(a.pos_in)[e];
        y = p.x; x = p.y;
      } else {
        y = a.pos_in[2 * (long long)e]; x = a.pos_in[2 * (long long)e + 1];
      }
    }
    y = lmin(lmax(y + kActionDy[act], 0), a.rows - 1);  // general_env.py:214-231
    x = lmin(lmax(x + kActionDx[act], 0), a.cols - 1);
    const bool stopped = live && (a.has_stopped[e] != 0 || (!bad && act == kStop));  // sticky, :233
    const int bit = (int)(y * a.cols + x);
    // ---- phase B
    // all loads first (independent: 2 * G in flight per lane), then the arithmetic
    uint32_t vv[G], bb[G];
#pragma unroll
    for (int it = 0; it < G; ++it) {
      const int ej = base + it * kPer + grp;
      const bool ok = ej < a.n && sub < a.words;
      const long long idx = (long long)ej * a.words + sub;
      vv[it] = ok ? a.visited[idx] : 0u;
      bb[it] = ok ? a.bbox[idx] : 0u;
    }
    // grids of 1025 ... 2048 patches: the second block of 32 words, loaded up front as well (a block of two warps
    // has registers to spare; loading them inside the loop below exposed one memory round trip per episode)
    constexpr int kTail = kWide ? 32 : 1;
    uint32_t tv[kTail], tb[kTail];
    if (kWide) {
#pragma unroll
      for (int it = 0; it < kTail; ++it) {
        const bool ok = base + it < a.n && sub + 32 < a.words;
        const long long idx = (long long)(base + it) * a.words + sub + 32;
        tv[it] = ok ? a.visited[idx] : 0u;
        tb[it] = ok ? a.bbox[idx] : 0u;
      }
    }
    uint32_t mine_fe = 0, mine_mf = 0;  // found | every << 16;  fresh | missing_after << 1
#pragma unroll
    for (int it = 0; it < G; ++it) {
      const int j = it * kPer + grp;  // episode of this group, relative to base
      const int bit_j = __shfl_sync(kFullMask, bit, j);
      const uint32_t v = vv[it], b = bb[it];
      const uint32_t m = (bit_j >> 5) == sub ? (1u << (bit_j & 31)) : 0u;
      uint32_t fe = __popc(v & b) | (__popc(b) << 16);  // counts use the map BEFORE marking (general_env.py:347)
      uint32_t mf = ((b & m & ~v) ? 1u : 0u) | (__popc(b & ~(v | m)) << 1);
      if (kWide) {
        const uint32_t v2 = tv[kWide ? it : 0], b2 = tb[kWide ? it : 0];
        const uint32_t m2 = (bit_j >> 5) == sub + 32 ? (1u << (bit_j & 31)) : 0u;
        fe += __popc(v2 & b2) | (__popc(b2) << 16);
        mf += ((b2 & m2 & ~v2) ? 1u : 0u) | (__popc(b2 & ~(v2 | m2)) << 1);
      }
      if (kWide && a.words > 64 && base + j < a.n) {  // grids of more than 2048 patches: the remaining words
        const long long row = (long long)(base + j) * a.words;
#pragma unroll 1
        for (int w = sub + 64; w < a.words; w += 32) {
          const uint32_t v2 = a.visited[row + w], b2 = a.bbox[row + w];
          const uint32_t m2 = (bit_j >> 5) == w ? (1u << (bit_j & 31)) : 0u;
          fe += __popc(v2 & b2) | (__popc(b2) << 16);
          mf += ((b2 & m2 & ~v2) ? 1u : 0u) | (__popc(b2 & ~(v2 | m2)) << 1);
        }
      }
      if (G == 32) {
        fe = __reduce_add_sync(kFullMask, fe);
        mf = __reduce_add_sync(kFullMask, mf);
        if (lane == it) { mine_fe = fe; mine_mf = mf; }
      } else {
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) {
          fe += __shfl_xor_sync(kFullMask, fe, off);
          mf += __shfl_xor_sync(kFullMask, mf, off);
        }
        // lane L owns episode L: handled in iteration L / kPer by group L % kPer
        const uint32_t got_fe = __shfl_sync(kFullMask, fe, (lane % kPer) * G);
        const uint32_t got_mf = __shfl_sync(kFullMask, mf, (lane % kPer) * G);
        if (lane / kPer == it) { mine_fe = got_fe; mine_mf = got_mf; }
      }
    }
    // ---- phase C
    bool first_visit = false;
    if (live) {
      const int found = (int)(mine_fe & 0xFFFFu), every = (int)(mine_fe >> 16);
      const bool fresh = (mine_mf & 1u) != 0, done = (mine_mf >> 1) == 0;
      float r = __fadd_rn(fresh ? 1.0f : 0.0f, a.cost);  // general_env.py:334-358
      if (a.stop_enabled) r = __fadd_rn(r, (float)(stopped ? (found == every ? found : found - every) : 0));
      const long long s = a.steps[e] + 1;
      atomicOr(a.visited + (long long)e * a.words + (bit >> 5), 1u << (bit & 31));
      a.rewards[e] = r;
      a.steps[e] = s;
      a.has_stopped[e] = stopped ? 1 : 0;
      a.truncated[e] = s >= a.max_ep_len ? 1 : 0;
      a.terminated[e] = a.stop_enabled ? (stopped ? 1 : 0) : (done ? 1 : 0);
      if (vec) {
        reinterpret_cast<longlong2*>(a.pos_out)[e] = make_longlong2(y, x);
      } else {
        a.pos_out[2 * (long long)e] = y; a.pos_out[2 * (long long)e + 1] = x;
      }
      if (a.first_slot) {  // zero-copy env: first visit -> read the tile from the host, else from the history
        int32_t* cell = a.first_slot + (long long)e * a.rows * a.cols + bit;
        const int seen = *cell;
        first_visit = seen < 0;
        if (first_visit) *cell = a.t;
        a.host_src[e] = first_visit ? e : -2;
        a.history_src[e] = first_visit ? -2 : e * a.slots + seen;
      }
    }
    if (a.host_tiles) {
      const unsigned firsts = __ballot_sync(kFullMask, first_visit);
      if (lane == 0 && firsts) atomicAdd(a.host_tiles, (unsigned long long)__popc(firsts));
    }
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
