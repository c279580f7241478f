// K1 -- the glimpse gather.  Four kernels behind one argument block:
//
//   copy   kernel: same dtype, plain [C,P,P] layout.  Pure DMA: one warp per CTA drives a
//                  ring of shared-memory stages, TMA loads (tensor-map tiles or per-row bulk
//                  copies) in, one TMA bulk store per stage out.  No thread touches a pixel.
//                  Chunks are dealt round-robin to a persistent grid (q = blockIdx.x + j * gridDim.x).
//   xform  kernel: uint8 -> float32 normalisation, the Focus space-to-depth layout, translated
//                  and / or virtually padded sources.  Warp 0 is the TMA producer: it draws
//                  tickets (small batches of chunks) from a global counter three batches ahead,
//                  decodes them as a software pipeline (no load is consumed in the iteration
//                  that issued it) and hands every stage a descriptor through shared memory; the
//                  other warps read the staged chunk, convert, and write coalesced vector stores.
//                  With `actions` it also applies an env step's move to the positions it reads
//                  (jn_env_step_gather), which makes it independent of the step kernel.
//   rows   kernel: plain loads, one warp per tile row: what the TMA unit cannot address --
//                  lists of images combined with a translation or padding, uint8 -> uint8 copies
//                  with unaligned offsets.
//   ldg    kernel: element-wise last resort (patch sizes that are not multiples of 4, outputs
//                  that are not 16-byte aligned); with the rows kernel it is the "ldg" engine
//                  and the in-GPU cross-check of the TMA engines in the tests.
//
// Work decomposition: a *chunk* is `rows` consecutive rows of one channel of one tile
// (rows * P * elem bytes, contiguous in the plain output).
#pragma once

#include "jn_device.cuh"

namespace jnk {

struct ImageRec {        // one per image when the set has several slabs (<= 32 bytes: jn_images_table_bytes)
  const uint8_t* base;   // first byte of this image's channel 0
  int32_t height, width; // pixels
  int32_t slab;          // slab this image lives in
  int32_t plane0;        // index of this image's channel-0 plane inside its slab
};

struct GatherArgs {
  const uint8_t* base;        // single-slab sets: first byte of image 0
  const ImageRec* images;     // multi-slab sets: per-image records (device memory), else null
  const int64_t* positions;   // [n_items, 2] (y, x) patch coordinates; null = patch (0, 0) for every item
  const int32_t* src_index;   // [n_items] image per item (negative = zero fill) or null = identity
  const int32_t* shifts;      // [n_images, 2] (ty, tx) integer translation of each image (zero fill) or null
  uint8_t* out;
  long long out_item_stride;  // bytes
  long long image_stride;     // single slab: bytes between images
  int32_t* status;            // device flag word or null
  int n_items, n_images;
  int channels, height, width;  // single slab geometry
  int patch, elem;              // patch size (pixels), source element size (bytes)
  int rows;                     // rows per chunk
  int chunks_per_plane;         // patch / rows
  int total_chunks;             // n_items * channels * chunks_per_plane
  int box_w, kbox;              // tensor engine: patch = box_w * kbox
  int pitch, stage_bytes;       // xform kernel: bytes per staged row (patch * elem, + 16 when translating), per stage
  uint32_t wpr_magic;           // xform kernel: ceil(2^32 / (patch / 4)), division by multiply-high
  int* work_counter;            // xform kernel: {next ticket, CTAs done}; zero before and after a launch
  // xform kernel: the ticket schedule (see claim_schedule in jn_api.cu).  Ticket k of segment j
  // (sched_ticket[j] <= k < sched_ticket[j+1]) is the batch of up to sched_size[j] chunks starting at
  // sched_chunk[j] + (k - sched_ticket[j]) * sched_size[j], cut at sched_chunk[j+1].
  int sched_n;
  int sched_size[6], sched_ticket[7], sched_chunk[7];
  int padded;                   // image sizes are rounded up to the patch grid, pixels outside the image are zeros
  int skip_negative;            // negative src_index: 1 = leave the output tile untouched, 0 = zero-fill it
  // fused env step (jn_env_step_gather): `positions` are the positions BEFORE the move and the kernel applies
  // actions[item] itself (move + clamp to grid_rows x grid_cols, general_env.py:209-231), so that it does not
  // depend on the step kernel it was launched behind
  const int64_t* actions;
  int grid_rows, grid_cols;
  int stream_stores;            // 1 = the converting kernel writes with st.global.cs (uint8 sources; JN_STREAM_STORES)
  int wait_prior;               // 1 = griddepcontrol.wait before the first index load (src_index comes from the step kernel)
};

struct Chunk {
  const uint8_t* src;  // first byte of the chunk's first row in the source image (null = zero fill)
  long long src_row_bytes;
  int item, channel, row0;  // tile-local first row
  int plane, px, py;  // plane = index of (image, channel) inside the slab: the tensor map's outer coordinate
  int sy, sx;  // translation of the source image: tile pixel (r, c) <- image pixel (py*P + r - sy, px*P + c - sx)
};

// Extent of the patch grid in pixels: the image size, rounded up to whole patches for padded sets.
__device__ __forceinline__ int grid_extent(const GatherArgs& a, int size) {
  return a.padded ? (size + a.patch - 1) / a.patch * a.patch : size;
}

// Patch coordinates of item `item`: positions[item], moved by actions[item] when the gather runs fused with
// an env step (same arithmetic as env_step_kernel: invalid action codes do not move).
__device__ __forceinline__ void item_position(const GatherArgs& a, long long item, long long& y, long long& x) {
  if (!a.positions) { y = 0; x = 0; return; }  // sets of one-patch images
  y = a.positions[2 * item];
  x = a.positions[2 * item + 1];
  if (a.actions) {
    long long act = a.actions[item];
    if (act < 0 || act > kStop) act = kStop;
    y = lmin(lmax(y + kActionDy[act], 0), a.grid_rows - 1);
    x = lmin(lmax(x + kActionDx[act], 0), a.grid_cols - 1);
  }
}

// Decode chunk q.  Out-of-grid positions are reported once and treated as "skip" (src = null,
// skip = true); negative src_index means zero fill (src = null, skip = false).
__device__ __forceinline__ bool decode_chunk(const GatherArgs& a, int q, Chunk& c) {
  const int cpi = a.channels * a.chunks_per_plane;
  c.item = q / cpi;
  const int rem = q - c.item * cpi;
  c.channel = rem / a.chunks_per_plane;
  c.row0 = (rem - c.channel * a.chunks_per_plane) * a.rows;
  c.src = nullptr;
  c.plane = 0;
  const int img = a.src_index ? a.src_index[c.item] : c.item;
  if (img < 0) return a.skip_negative != 0 || img < -1;  // -1: zero fill (or skip on request); <= -2: always skip
  long long y, x;
  item_position(a, c.item, y, x);
  const uint8_t* base;
  int h, w;
  if (a.images) {
    const ImageRec r = a.images[img < a.n_images ? img : 0];  // (a bad index is reported below, never dereferenced)
    base = r.base; h = r.height; w = r.width;
    c.plane = r.plane0 + c.channel;
  } else {
    base = a.base + (long long)img * a.image_stride; h = a.height; w = a.width;
    c.plane = img * a.channels + c.channel;
  }
  if (img >= a.n_images || y < 0 || x < 0 || (y + 1) * a.patch > grid_extent(a, h) ||
      (x + 1) * a.patch > grid_extent(a, w)) {
    if (a.status && c.channel == 0 && c.row0 == 0) atomicOr(a.status, 1);
    return true;  // skip
  }
  c.px = (int)x; c.py = (int)y;
  c.sy = a.shifts ? a.shifts[2 * img] : 0;
  c.sx = a.shifts ? a.shifts[2 * img + 1] : 0;
  c.src_row_bytes = (long long)w * a.elem;
  c.src = base + (((long long)c.channel * h + y * a.patch + c.row0) * w + x * a.patch) * a.elem;
  return false;
}

// ------------------------------------------------------------------------------------------
// copy kernel (DMA only)
// ------------------------------------------------------------------------------------------
constexpr int kZeroBytes = 4096;

template <int kStages, int kAhead, bool kTensor>
__global__ void __launch_bounds__(32, 1)
gather_copy_kernel(const __grid_constant__ GatherArgs a, const __grid_constant__ CUtensorMap map0) {
  static_assert(kAhead >= 1 && kAhead < kStages, "lookahead must leave room for stores in flight");
  extern __shared__ __align__(1024) uint8_t smem[];
  const int lane = threadIdx.x;
  const uint32_t chunk_bytes = (uint32_t)a.rows * a.patch * a.elem;
  const uint32_t row_bytes = (uint32_t)a.patch * a.elem;
  uint8_t* zero = smem + (size_t)kStages * chunk_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(zero + kZeroBytes);

  if (lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
    if (kTensor) prefetch_tensormap(&map0);
  }
  for (int i = lane * 16; i < kZeroBytes; i += 32 * 16) *reinterpret_cast<uint4*>(zero + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();  // zeros (generic proxy) -> visible to the bulk stores (async proxy)
  __syncwarp();
  if (a.wait_prior) pdl_wait();  // everything above overlapped the tail of the kernel before us
  pdl_launch_dependents();       // a gather launched behind this one (history reuse) may run next to it

  const int grid = gridDim.x;
  const int mine = a.total_chunks > (int)blockIdx.x ? (a.total_chunks - (int)blockIdx.x + grid - 1) / grid : 0;

  for (int j = 0; j < mine + kAhead; ++j) {
    if (j < mine) {  // ---- load chunk j into stage j % kStages
      const int st = j % kStages;
      uint8_t* stage = smem + (size_t)st * chunk_bytes;
      if (j >= kStages && lane == 0) bulk_wait_read<kStages - kAhead - 1>();  // store of chunk j-kStages has read its stage
      __syncwarp();
      Chunk c;
      const bool skip = decode_chunk(a, (int)blockIdx.x + j * grid, c);
      if (c.src != nullptr && !skip) {
        if (lane == 0) mbar_arrive_expect_tx(&full[st], chunk_bytes);
        __syncwarp();
        if (kTensor) {
          if (lane == 0) {
            const CUtensorMap* m = &map0;
            if (a.shifts)  // 3-D map [W, H, planes]: x offsets in 16-byte steps, any y; out-of-image pixels arrive as zeros
              tensor_g2s_3d(stage, m, c.px * a.patch - c.sx, c.py * a.patch + c.row0 - c.sy, c.plane, &full[st]);
            else
              tensor_g2s_4d(stage, m, 0, c.px * a.kbox, c.py * a.patch + c.row0, c.plane, &full[st]);
          }
        } else {
          for (int r = lane; r < a.rows; r += 32)
            bulk_g2s(stage + (size_t)r * row_bytes, c.src + (long long)r * c.src_row_bytes, row_bytes, &full[st]);
        }
      } else if (lane == 0) {
        mbar_arrive(&full[st]);  // nothing to load: complete the phase by hand
      }
    }
    const int k = j - kAhead;
    if (k >= 0 && lane == 0) {  // ---- store chunk k
      const int st = k % kStages;
      Chunk c;
      const bool skip = decode_chunk(a, (int)blockIdx.x + k * grid, c);
      mbar_wait(&full[st], (uint32_t)(k / kStages) & 1u);
      if (!skip) {
        uint8_t* dst = a.out + (long long)c.item * a.out_item_stride +
                       ((long long)c.channel * a.patch + c.row0) * row_bytes;
        if (c.src != nullptr) {
          bulk_s2g(dst, smem + (size_t)st * chunk_bytes, chunk_bytes);
        } else {
          for (uint32_t off = 0; off < chunk_bytes; off += kZeroBytes)
            bulk_s2g(dst + off, zero, min((uint32_t)kZeroBytes, chunk_bytes - off));
        }
      }
      bulk_commit();  // one group per chunk, even when empty: keeps the wait_group arithmetic exact
    }
  }
  if (lane == 0) bulk_wait<0>();
}

// ------------------------------------------------------------------------------------------
// xform kernel (uint8 -> float32, Focus layout, unaligned translation)
// ------------------------------------------------------------------------------------------
// Warp 0 is the TMA producer; the consumer warps read the staged chunk from shared memory,
// convert and write coalesced vector stores.
//
//  * The producer decodes a ticket's chunks one per lane and hands each stage a 16-byte
//    descriptor through shared memory, so no consumer touches the index arrays.  The descriptor
//    is written before the producer's arrive on the stage's `full` barrier (release) and read
//    after the consumers' wait (acquire).
//  * kShift: integer translation with ANY x offset.  The TMA unit only takes 16-byte aligned
//    inner coordinates, so the producer loads the aligned superset of each row (P*elem + 16
//    bytes, 3-D map of 8-byte elements, out-of-image bytes arrive as zeros) and the consumers
//    read it back at the residual byte offset (funnel shift for uint8, word select for float32).
enum XformMode { kNormPlain = 0, kF32Focus = 1, kNormFocus = 2, kF32Plain = 3 };

struct __align__(16) StageDesc {
  unsigned long long dst;  // first byte of the chunk's output item
  int chrow;               // channel << 16 | first tile row
  int flags;               // kDescSkip | kDescZero | kDescStop | residual byte offset << 8 | frame limits (below)
};
constexpr int kDescSkip = 1, kDescZero = 2, kDescStop = 4;
// Translated gathers out of padded sets: the translated image is clipped to its own H x W frame before the
// padding (dataset order: transform, then padded_collate_fn), so tile pixels whose FRAME coordinate lies in the
// padding are zeros even if the shifted source pixel exists.  Per chunk: rows / columns of the tile that are
// inside the frame (rows <= 256 -> 9 bits at 12, columns <= 2047 -> 11 bits at 21).
__device__ __forceinline__ int desc_pack_limits(int row_lim, int col_lim) { return (row_lim << 12) | (col_lim << 21); }
__device__ __forceinline__ int desc_offset(int flags) { return ((unsigned)flags >> 8) & 15; }
__device__ __forceinline__ int desc_row_limit(int flags) { return ((unsigned)flags >> 12) & 511; }
__device__ __forceinline__ int desc_col_limit(int flags) { return (unsigned)flags >> 21; }

// 4 staged source pixels of group g of a staged row (uint8: one word; float32: one float4)
template <bool kShift>
__device__ __forceinline__ uint32_t staged_u8x4(const uint8_t* row, int g, int off) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(row) + g;
  if (!kShift) return w[0];
  w += off >> 2;
  return __funnelshift_r(w[0], w[1], (off & 3) * 8);
}
template <bool kShift>
__device__ __forceinline__ float4 staged_f32x4(const uint8_t* row, int g, int off) {
  const float4* v = reinterpret_cast<const float4*>(row) + g;
  const float4 A = v[0];
  if (!kShift) return A;
  const float4 B = v[1];
  switch (off >> 2) {  // uniform over the CTA
    case 0: return A;
    case 1: return make_float4(A.y, A.z, A.w, B.x);
    case 2: return make_float4(A.z, A.w, B.x, B.y);
    default: return make_float4(A.w, B.x, B.y, B.z);
  }
}

// Zero the pixels of group (row i / wpr, columns 4g .. 4g+3) that lie outside the frame limits.
__device__ __forceinline__ uint32_t clip_u8x4(uint32_t u, int i, int g, int rows, int wpr, uint32_t magic, int row_lim,
                                             int col_lim) {
  const int r = (int)__umulhi((uint32_t)i, magic), gg = i - r * wpr;
  (void)g; (void)rows;
  const int valid = col_lim - 4 * gg;  // pixels of this group inside the frame
  if (r >= row_lim || valid <= 0) return 0u;
  return valid >= 4 ? u : (u & (0xFFFFFFFFu >> (8 * (4 - valid))));
}
__device__ __forceinline__ float4 clip_f32x4(float4 v, int i, int g, int rows, int wpr, uint32_t magic, int row_lim,
                                             int col_lim) {
  const int r = (int)__umulhi((uint32_t)i, magic), gg = i - r * wpr;
  (void)g; (void)rows;
  const int valid = (r >= row_lim) ? 0 : col_lim - 4 * gg;
  return make_float4(valid > 0 ? v.x : 0.f, valid > 1 ? v.y : 0.f, valid > 2 ? v.z : 0.f, valid > 3 ? v.w : 0.f);
}

template <int kMode, bool kShift>
__device__ __forceinline__ void xform_chunk(const GatherArgs& a, const StageDesc& d, const uint8_t* stage, int tid,
                                            int nthreads) {
  // Every iteration is independent of the previous one (row / group of item i come from one multiply-high,
  // not from a carried counter), so the unrolled body keeps several shared-memory loads in flight.
  const int P = a.patch, wpr = P >> 2;  // wpr = 4-pixel groups per row
  const int n = a.rows * wpr, pitch = a.pitch;
  const uint32_t magic = a.wpr_magic;  // ceil(2^32 / wpr): i / wpr == umulhi(i, magic) for every i < n
  const int ch = d.chrow >> 16, row0 = d.chrow & 0xFFFF, off = desc_offset(d.flags);
  const bool zero = (d.flags & kDescZero) != 0;
  // frame clipping (translated + padded sets only): uniform over the chunk
  const int row_lim = kShift ? desc_row_limit(d.flags) : a.rows, col_lim = kShift ? desc_col_limit(d.flags) : P;
  const bool clip = kShift && (row_lim < a.rows || col_lim < P);
  const bool cs = a.stream_stores != 0;
  float* out_item = reinterpret_cast<float*>(d.dst);
  if (kMode == kNormPlain || kMode == kF32Plain) {
    // rows * P source pixels in, rows * P floats out, contiguous in the output
    float* dst = out_item + ((long long)ch * P + row0) * P;
    if (zero) {
      for (int i = tid; i < n; i += nthreads) st_f4(dst + 4 * i, 0.f, 0.f, 0.f, 0.f);
      return;
    }
#pragma unroll 4
    for (int i = tid; i < n; i += nthreads) {
      const uint8_t* row = stage;  // not translated: the stage is contiguous, group i sits at word / float4 i
      int g = i;
      if (kShift) {
        const int r = (int)__umulhi((uint32_t)i, magic);
        g = i - r * wpr;
        row = stage + r * pitch;
      }
      if (kMode == kNormPlain) {
        uint32_t u = staged_u8x4<kShift>(row, g, off);
        if (clip) u = clip_u8x4(u, i, g, a.rows, wpr, magic, row_lim, col_lim);
        if (cs) st_f4_cs(dst + 4 * i, byte_to_unit<0>(u), byte_to_unit<1>(u), byte_to_unit<2>(u), byte_to_unit<3>(u));
        else st_f4(dst + 4 * i, byte_to_unit<0>(u), byte_to_unit<1>(u), byte_to_unit<2>(u), byte_to_unit<3>(u));
      } else {
        float4 v = staged_f32x4<kShift>(row, g, off);
        if (clip) v = clip_f32x4(v, i, g, a.rows, wpr, magic, row_lim, col_lim);
        if (cs) st_f4_cs(dst + 4 * i, v.x, v.y, v.z, v.w);
        else st_f4(dst + 4 * i, v.x, v.y, v.z, v.w);
      }
    }
  } else {
    // Focus: out[(dy + 2*dx) * C + ch][i][j] = tile[ch][2i + dy][2j + dx]; chunks start on even rows
    const int half = P >> 1, plane = half * half;
    const int dy_planes = a.channels * plane, dx_planes = 2 * dy_planes;
    float* base = out_item + (long long)ch * plane + (row0 >> 1) * half;
    if (zero) {
      for (int i = tid; i < n; i += nthreads) {
        const int r = (int)__umulhi((uint32_t)i, magic), g = i - r * wpr;
        float* even = base + ((r & 1) * dy_planes + (r >> 1) * half + 2 * g);
        st_f2(even, 0.f, 0.f);
        st_f2(even + dx_planes, 0.f, 0.f);
      }
      return;
    }
#pragma unroll 4
    for (int i = tid; i < n; i += nthreads) {
      const int r = (int)__umulhi((uint32_t)i, magic), g = i - r * wpr;
      const uint8_t* row = kShift ? stage + r * pitch : stage;
      const int gi = kShift ? g : i;
      float e0, o0, e1, o1;
      if (kMode == kF32Focus) {
        float4 v = staged_f32x4<kShift>(row, gi, off);
        if (clip) v = clip_f32x4(v, i, g, a.rows, wpr, magic, row_lim, col_lim);
        e0 = v.x; o0 = v.y; e1 = v.z; o1 = v.w;
      } else {
        uint32_t u = staged_u8x4<kShift>(row, gi, off);
        if (clip) u = clip_u8x4(u, i, g, a.rows, wpr, magic, row_lim, col_lim);
        e0 = byte_to_unit<0>(u); o0 = byte_to_unit<1>(u); e1 = byte_to_unit<2>(u); o1 = byte_to_unit<3>(u);
      }
      float* even = base + ((r & 1) * dy_planes + (r >> 1) * half + 2 * g);
      if (cs) { st_f2_cs(even, e0, e1); st_f2_cs(even + dx_planes, o0, o1); }
      else { st_f2(even, e0, e1); st_f2(even + dx_planes, o0, o1); }
    }
  }
}

__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, int src_lane) {
  const uint32_t lo = __shfl_sync(0xFFFFFFFFu, (uint32_t)v, src_lane);
  const uint32_t hi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), src_lane);
  return ((unsigned long long)hi << 32) | lo;
}

// What the producer needs to know about one chunk, decoded by one lane.
struct ChunkPlan {
  unsigned long long dst, src;  // output item base; first source byte of the chunk (bulk engine)
  int src_row_bytes;
  int chrow, flags;             // as in StageDesc
  int cx, cy, plane;            // tensor-map coordinates
};

// The decode of a chunk needs three dependent memory round trips (claim -> src_index -> position / action /
// shift / image record).  The producer takes them one batch apart (see gather_xform_kernel), so the decode is
// split where the dependencies are: each function only ISSUES loads whose addresses it can already compute and
// hands the raw registers on; nothing looks at a loaded value before the next stage, one batch later.

// Stage B: image index of chunk q's item.
__device__ __forceinline__ int load_chunk_image(const GatherArgs& a, int q, bool valid) {
  if (!valid) return -1;
  const int item = q / (a.channels * a.chunks_per_plane);
  return a.src_index ? a.src_index[item] : item;
}

// Stage C: everything else the plan of a chunk reads from memory.
struct ChunkLoads {
  int q, img;           // chunk (negative: lane idle) and image index
  long long y, x, act;  // patch position and action code as stored
  int sy, sx;           // translation of the image
  ImageRec rec;         // multi-slab sets: the image's record
};

__device__ __forceinline__ ChunkLoads load_chunk_inputs(const GatherArgs& a, int q, bool valid, int img) {
  ChunkLoads c;
  c.q = valid ? q : -1; c.img = img;
  c.y = c.x = 0; c.act = kStop; c.sy = c.sx = 0;
  c.rec.base = nullptr; c.rec.height = c.rec.width = c.rec.slab = c.rec.plane0 = 0;
  if (!valid || img < 0) return c;
  const long long item = q / (a.channels * a.chunks_per_plane);
  if (a.positions) { c.y = a.positions[2 * item]; c.x = a.positions[2 * item + 1]; }
  if (a.actions) c.act = a.actions[item];
  if (img < a.n_images) {  // (a bad index is reported by plan_chunk, never dereferenced)
    if (a.shifts) { c.sy = a.shifts[2 * img]; c.sx = a.shifts[2 * img + 1]; }
    if (a.images) c.rec = a.images[img];
  }
  return c;
}

// Stage D: pure arithmetic on the loaded values -> what the TMA issue needs.
template <bool kShift>
__device__ __forceinline__ ChunkPlan plan_chunk(const GatherArgs& a, const ChunkLoads& c) {
  ChunkPlan p;
  p.dst = p.src = 0; p.src_row_bytes = 0; p.chrow = 0; p.flags = kDescSkip; p.cx = p.cy = p.plane = 0;
  if (c.q < 0) return p;
  const int q = c.q, img = c.img;
  const int cpi = a.channels * a.chunks_per_plane;
  const int item = q / cpi;
  const int rem = q - item * cpi;
  const int channel = rem / a.chunks_per_plane;
  const int row0 = (rem - channel * a.chunks_per_plane) * a.rows;
  p.dst = reinterpret_cast<unsigned long long>(a.out + (long long)item * a.out_item_stride);
  p.chrow = (channel << 16) | row0;
  if (img < 0) {  // -1: zero fill (or skip when the caller asked for that); <= -2: always skip
    p.flags = (a.skip_negative || img < -1) ? kDescSkip : kDescZero;
    return p;
  }
  long long y = c.y, x = c.x;
  if (a.actions) {  // fused env step: same move + clamp as env_step_kernel (see item_position)
    long long act = c.act;
    if (act < 0 || act > kStop) act = kStop;
    y = lmin(lmax(y + kActionDy[act], 0), a.grid_rows - 1);
    x = lmin(lmax(x + kActionDx[act], 0), a.grid_cols - 1);
  }
  const uint8_t* base;
  int h, w;
  if (a.images) {
    base = c.rec.base; h = c.rec.height; w = c.rec.width;
    p.plane = c.rec.plane0 + channel;
  } else {
    base = a.base + (long long)img * a.image_stride; h = a.height; w = a.width;
    p.plane = img * a.channels + channel;
  }
  if (img >= a.n_images || y < 0 || x < 0 || (y + 1) * a.patch > grid_extent(a, h) ||
      (x + 1) * a.patch > grid_extent(a, w)) {
    if (a.status && channel == 0 && row0 == 0) atomicOr(a.status, 1);
    return p;  // out-of-grid position: reported once, tile skipped
  }
  p.flags = 0;
  const int px = (int)x, py = (int)y;
  const int sy = c.sy, sx = c.sx;
  p.src_row_bytes = w * a.elem;
  p.src = reinterpret_cast<unsigned long long>(base + (((long long)channel * h + y * a.patch + row0) * w + x * a.patch) * a.elem);
  if (kShift) {
    const int xb = (px * a.patch - sx) * a.elem;  // byte offset of the tile's first pixel in its image row
    const int xa = xb & ~15;                       // 16-byte floor (also for negative offsets)
    p.flags |= (xb - xa) << 8;
    p.cx = xa >> 3;                                // in 8-byte map elements
    p.cy = py * a.patch + row0 - sy;
    // rows / columns of this chunk inside the image's own frame (everything, unless the set is padded)
    const int row_lim = a.padded ? imax(0, imin(a.rows, h - (py * a.patch + row0))) : a.rows;
    const int col_lim = a.padded ? imax(0, imin(a.patch, w - px * a.patch)) : a.patch;
    p.flags |= desc_pack_limits(row_lim, col_lim);
  } else {
    p.cx = px * a.kbox;
    p.cy = py * a.patch + row0;
  }
  return p;
}

template <int kMode, int kConsumerWarps, bool kShift>
__global__ void __launch_bounds__((kConsumerWarps + 1) * 32)
gather_xform_kernel(const __grid_constant__ GatherArgs a, const __grid_constant__ CUtensorMap map0, const int stages,
                    const int tensor) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tx_bytes = (uint32_t)a.rows * a.pitch;
  const uint32_t row_bytes = (uint32_t)a.patch * a.elem;
  StageDesc* desc = reinterpret_cast<StageDesc*>(smem + (size_t)stages * a.stage_bytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(desc + stages);
  uint64_t* empty = full + stages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kConsumerWarps);
    }
    fence_mbar_init();
    if (tensor) prefetch_tensormap(&map0);
  }
  __syncthreads();
  if (a.wait_prior) pdl_wait();  // barrier init / tensor-map prefetch overlapped the tail of the kernel before us
  // A gather launched behind this one with the programmatic attribute may start now: in the zero-copy env that is
  // the copy of revisited patches out of the crop history (HBM -> HBM), which then runs NEXT TO this PCIe-bound
  // kernel instead of after it.  Everything it reads (history_src from the step kernel, older history slots) is
  // complete by now: this kernel has just waited for the step kernel.  A normally launched successor still waits
  // for this grid to finish.
  pdl_launch_dependents();

  int st = 0;
  uint32_t phase = 0;  // parity of the ring round this warp is in

  if (warp == 0) {  // ---- TMA producer
    // Work is handed out as TICKETS from a global counter (a statically dealt grid drifts apart: SMs that see
    // a faster memory path finish early and leave a tail).  Ticket k stands for a small batch of chunks fixed
    // by a schedule the host computed for this launch (claim_schedule in jn_api.cu: 4 chunks, then 2, then
    // single chunks at the end).  The sizes depend on the ticket number only, not on how much work seemed to be
    // left when it was drawn -- tickets are drawn three batches ahead (below), when that estimate would be stale.
    //
    // Decoding a batch takes three dependent round trips to memory (ticket, src_index, positions & co): ~3 us,
    // as long as 3-4 chunks take to convert.  They run as a software pipeline, one batch apart, so that the
    // warp never looks at a value in the iteration that asked for it:
    //     A  atomicAdd for the ticket of batch b+3      (result read in the next iteration)
    //     B  src_index loads of batch b+2               (   "    )
    //     C  position / action / shift / record loads of batch b+1
    //     D  arithmetic + TMA issue of batch b
    // The first ticket of every CTA is its own index (no round trip before the first load); the counter hands
    // out the tickets from gridDim.x on.  Claims still in flight when the tickets run out are harmless.
    const int total = a.total_chunks, grid = gridDim.x;
    auto ticket_range = [&](int k, int& base, int& size) {
      base = total; size = 0;
#pragma unroll
      for (int j = 0; j < 6; ++j)
        if (j < a.sched_n && k >= a.sched_ticket[j] && k < a.sched_ticket[j + 1]) {
          base = a.sched_chunk[j] + (k - a.sched_ticket[j]) * a.sched_size[j];
          size = min(a.sched_size[j], a.sched_chunk[j + 1] - base);
        }
    };
    auto claim = [&]() { return lane == 0 ? atomicAdd(a.work_counter, 1) : 0; };
    // A: ticket of batch 1
    int a_ret = claim();
    // B: batch 0
    int b_base, b_size;
    ticket_range((int)blockIdx.x, b_base, b_size);
    int b_img = load_chunk_image(a, b_base + lane, lane < b_size);
    // C: batch 0 (waits for B once); B: batch 1 (waits for A once); A: batch 2
    int c_base = b_base, c_size = b_size;
    ChunkLoads c = load_chunk_inputs(a, c_base + lane, lane < c_size, b_img);
    ticket_range(grid + __shfl_sync(0xFFFFFFFFu, a_ret, 0), b_base, b_size);
    b_img = load_chunk_image(a, b_base + lane, lane < b_size);
    a_ret = claim();
    bool ring_used = false;  // true once every stage has been filled once
    while (c_base < total) {
      const ChunkPlan cur = plan_chunk<kShift>(a, c);  // D: the loads of C were issued one batch ago
      const int count = c_size;
      // advance the pipeline before issuing, so that its loads fly while the ring is being fed
      c_base = b_base; c_size = b_size;
      c = load_chunk_inputs(a, c_base + lane, lane < c_size, b_img);
      ticket_range(grid + __shfl_sync(0xFFFFFFFFu, a_ret, 0), b_base, b_size);
      b_img = load_chunk_image(a, b_base + lane, lane < b_size);
      a_ret = claim();
      for (int k = 0; k < count; ++k) {
        const int flags = __shfl_sync(0xFFFFFFFFu, cur.flags, k);
        const int chrow = __shfl_sync(0xFFFFFFFFu, cur.chrow, k);
        const int cx = __shfl_sync(0xFFFFFFFFu, cur.cx, k), cy = __shfl_sync(0xFFFFFFFFu, cur.cy, k);
        const int plane = __shfl_sync(0xFFFFFFFFu, cur.plane, k);
        const int src_row_bytes = __shfl_sync(0xFFFFFFFFu, cur.src_row_bytes, k);
        const unsigned long long dst = shfl_u64(cur.dst, k);
        const unsigned long long src = shfl_u64(cur.src, k);
        uint8_t* stage = smem + (size_t)st * a.stage_bytes;
        if (ring_used) mbar_wait(&empty[st], phase ^ 1u);
        if (lane == 0) {
          StageDesc d;
          d.dst = dst; d.chrow = chrow; d.flags = flags;
          desc[st] = d;
        }
        if ((flags & (kDescSkip | kDescZero)) == 0) {
          if (lane == 0) mbar_arrive_expect_tx(&full[st], tx_bytes);
          __syncwarp();
          if (tensor) {
            if (lane == 0) {
              if (kShift) tensor_g2s_3d(stage, &map0, cx, cy, plane, &full[st]);
              else tensor_g2s_4d(stage, &map0, 0, cx, cy, plane, &full[st]);
            }
          } else {
            const uint8_t* from = reinterpret_cast<const uint8_t*>(src);
            for (int r = lane; r < a.rows; r += 32)
              bulk_g2s(stage + (size_t)r * row_bytes, from + (long long)r * src_row_bytes, row_bytes, &full[st]);
          }
        } else if (lane == 0) {
          mbar_arrive(&full[st]);  // nothing to load: complete the phase by hand
        }
        if (++st == stages) { st = 0; phase ^= 1u; ring_used = true; }
      }
    }
    // the claim that is still in flight has to land before this CTA reports itself done: whoever resets the
    // counters for the next launch must be the last one to touch them
    asm volatile("" ::"r"(a_ret), "r"(b_img) : "memory");
    __threadfence();
    // tell the consumers to leave, then clean the counters up for the next launch: the last CTA whose claims
    // have all come back empty knows that nobody will touch them again
    if (ring_used) mbar_wait(&empty[st], phase ^ 1u);
    if (lane == 0) {
      StageDesc d;
      d.dst = 0; d.chrow = 0; d.flags = kDescStop;
      desc[st] = d;
      mbar_arrive(&full[st]);
      if (atomicAdd(a.work_counter + 1, 1) == (int)gridDim.x - 1) {
        a.work_counter[1] = 0;
        __threadfence();
        a.work_counter[0] = 0;
      }
    }
  } else {  // ---- consumers
    const int tid = threadIdx.x - 32;
    for (;;) {
      mbar_wait(&full[st], phase);
      const StageDesc d = desc[st];
      if (d.flags & kDescStop) break;
      if ((d.flags & kDescSkip) == 0)
        xform_chunk<kMode, kShift>(a, d, smem + (size_t)st * a.stage_bytes, tid, kConsumerWarps * 32);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
      if (++st == stages) { st = 0; phase ^= 1u; }
    }
  }
}

// ------------------------------------------------------------------------------------------
// rows kernel (plain loads, one warp per tile row)
// ------------------------------------------------------------------------------------------
// For what TMA cannot address: source rows that do not start on a 16-byte boundary (arbitrary integer
// translation: the TMA unit traps on inner coordinates that are not 16-byte multiples) and lists of
// images combined with a translation.  A warp owns one row of one channel of one tile; each lane
// loads four consecutive source pixels with scalar loads (any alignment; the warp still covers whole
// cache lines) and writes one aligned 16-byte store (two 8-byte stores in the Focus layout).  Pixels
// outside the translated image are zeros.  Needs P % 4 == 0 and a 16-byte aligned output.
__global__ void __launch_bounds__(256)
gather_rows_kernel(const GatherArgs a, const int out_f32, const int normalize, const int focus) {
  if (a.wait_prior) pdl_wait();
  const int P = a.patch, lane = threadIdx.x & 31;
  const long long units = (long long)a.n_items * a.channels * P;  // (item, channel, row)
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long u = warp0; u < units; u += n_warps) {
    const int item = (int)(u / ((long long)a.channels * P));
    const int rem = (int)(u - (long long)item * a.channels * P);
    const int ch = rem / P, r = rem - ch * P;
    const int img = a.src_index ? a.src_index[item] : item;
    const uint8_t* base = nullptr;
    int h = 0, w = 0;
    long long y = 0, x = 0;
    bool zero_row = img < 0;
    if (zero_row && (a.skip_negative || img < -1)) continue;
    if (!zero_row) {
      item_position(a, item, y, x);
      if (a.images) {
        const ImageRec rec = a.images[img < a.n_images ? img : 0];
        base = rec.base; h = rec.height; w = rec.width;
      } else {
        base = a.base + (long long)img * a.image_stride; h = a.height; w = a.width;
      }
      if (img >= a.n_images || y < 0 || x < 0 || (y + 1) * P > grid_extent(a, h) || (x + 1) * P > grid_extent(a, w)) {
        if (a.status && lane == 0 && ch == 0 && r == 0) atomicOr(a.status, 1);
        continue;  // out-of-grid position: tile skipped
      }
    }
    const long long sy = zero_row ? -1 : y * P + r - (a.shifts ? a.shifts[2 * img] : 0);
    const long long sx0 = zero_row ? 0 : x * P - (a.shifts ? a.shifts[2 * img + 1] : 0);
    // (padded sets: the translated image is clipped to its own frame before the padding)
    const bool row_ok = !zero_row && sy >= 0 && sy < h && y * P + r < h;
    const uint8_t* row = row_ok ? base + (((long long)ch * h + sy) * w) * a.elem : nullptr;
    uint8_t* dst_item = a.out + (long long)item * a.out_item_stride;
    for (int g = lane; g < P / 4; g += 32) {
      float f[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t packed = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long sx = sx0 + 4 * g + k;
        if (row_ok && sx >= 0 && sx < w && x * P + 4 * g + k < w) {
          if (a.elem == 4) {
            f[k] = reinterpret_cast<const float*>(row)[sx];
          } else {
            const uint32_t b = row[sx];
            packed |= b << (8 * k);
            f[k] = normalize ? u8_to_unit((float)b) : (float)b;
          }
        }
      }
      if (!out_f32) {  // uint8 -> uint8 (plain layout only; Focus for uint8 output stays on the element kernel)
        reinterpret_cast<uint32_t*>(dst_item + ((long long)ch * P + r) * P)[g] = packed;
      } else if (!focus) {
        st_f4(reinterpret_cast<float*>(dst_item) + ((long long)ch * P + r) * P + 4 * g, f[0], f[1], f[2], f[3]);
      } else {
        const int half = P / 2;
        const long long plane = (long long)half * half;
        float* even = reinterpret_cast<float*>(dst_item) + (long long)((r & 1) * a.channels + ch) * plane +
                      (long long)(r >> 1) * half + 2 * g;
        st_f2(even, f[0], f[2]);
        st_f2(even + 2ll * a.channels * plane, f[1], f[3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// ldg kernel (element-wise fallback)
// ------------------------------------------------------------------------------------------
// One thread per output element of the plain layout; handles every dtype / flag combination.
__global__ void gather_ldg_kernel(const GatherArgs a, const int out_f32, const int normalize, const int focus) {
  if (a.wait_prior) pdl_wait();
  const int P = a.patch;
  const long long per_item = (long long)a.channels * P * P;
  const long long total = per_item * a.n_items;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int item = (int)(idx / per_item);
    long long rem = idx - (long long)item * per_item;
    const int ch = (int)(rem / ((long long)P * P));
    rem -= (long long)ch * P * P;
    const int r = (int)(rem / P), col = (int)(rem - (long long)r * P);
    const int img = a.src_index ? a.src_index[item] : item;
    float fv = 0.f;
    uint8_t bv = 0;
    if (img < 0 && (a.skip_negative || img < -1)) continue;
    if (img >= 0) {
      long long y, x;
      item_position(a, item, y, x);
      const uint8_t* base;
      int h, w;
      if (a.images) {
        const ImageRec rec = a.images[img < a.n_images ? img : 0];
        base = rec.base; h = rec.height; w = rec.width;
      } else {
        base = a.base + (long long)img * a.image_stride; h = a.height; w = a.width;
      }
      if (img >= a.n_images || y < 0 || x < 0 || (y + 1) * P > grid_extent(a, h) || (x + 1) * P > grid_extent(a, w)) {
        if (a.status && rem == 0 && ch == 0) atomicOr(a.status, 1);
        continue;
      }
      const long long sy = y * P + r - (a.shifts ? a.shifts[2 * img] : 0);
      const long long sx = x * P + col - (a.shifts ? a.shifts[2 * img + 1] : 0);
      if (sy >= 0 && sy < h && sx >= 0 && sx < w && y * P + r < h && x * P + col < w) {  // outside the (translated) image or its frame: zero fill
        const long long e = ((long long)ch * h + sy) * w + sx;
        if (a.elem == 4) {
          fv = reinterpret_cast<const float*>(base)[e];
        } else {
          bv = base[e];
          fv = normalize ? u8_to_unit((float)bv) : (float)bv;
        }
      } else if (normalize) {
        fv = 0.f;
      }
    }
    long long o;
    if (focus) {
      const int half = P / 2;
      const int pl = ((r & 1) + 2 * (col & 1)) * a.channels + ch;
      o = ((long long)pl * half + (r >> 1)) * half + (col >> 1);
    } else {
      o = rem + (long long)ch * P * P;
    }
    uint8_t* dst = a.out + (long long)item * a.out_item_stride;
    if (out_f32) reinterpret_cast<float*>(dst)[o] = fv;
    else dst[o] = bv;
  }
}

}  // namespace jnk
