// K1 -- the glimpse gather.  Three engines behind one argument block:
//
//   copy   kernel: same dtype, plain [C,P,P] layout.  Pure DMA: one warp per CTA drives a
//                  ring of shared-memory stages, TMA loads (tensor-map tiles or per-row bulk
//                  copies) in, one TMA bulk store per stage out.  No thread touches a pixel.
//   xform  kernel: uint8 -> float32 normalisation and / or the Focus space-to-depth layout.
//                  Warp 0 is the TMA producer, the other warps read the staged tile from
//                  shared memory, convert, and write coalesced vector stores.
//   rows   kernel: plain loads, one warp per tile row: what the TMA unit cannot address --
//                  source rows that do not start on a 16-byte boundary (arbitrary integer
//                  translation), lists of images combined with a translation.
//   ldg    kernel: element-wise last resort (patch sizes that are not multiples of 4, outputs
//                  that are not 16-byte aligned); with the rows kernel it is the "ldg" engine
//                  and the in-GPU cross-check of the TMA engines in the tests.
//
// Work decomposition: a *chunk* is `rows` consecutive rows of one channel of one tile
// (rows * P * elem bytes, contiguous in the plain output).  Chunks are dealt round-robin to
// a persistent grid (q = blockIdx.x + j * gridDim.x), so neighbouring SMs stream
// neighbouring rows of the same tile.
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
  const int64_t* positions;   // [n_items, 2] (y, x) patch coordinates
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
  int skip_negative;            // negative src_index: 1 = leave the output tile untouched, 0 = zero-fill it
};

struct Chunk {
  const uint8_t* src;  // first byte of the chunk's first row in the source image (null = zero fill)
  long long src_row_bytes;
  int item, channel, row0;  // tile-local first row
  int plane, px, py;  // plane = index of (image, channel) inside the slab: the tensor map's outer coordinate
  int sy, sx;  // translation of the source image: tile pixel (r, c) <- image pixel (py*P + r - sy, px*P + c - sx)
};

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
  if (img < 0) return a.skip_negative != 0;  // zero fill, or skip when the caller asked for that
  const long long y = a.positions[2 * (long long)c.item], x = a.positions[2 * (long long)c.item + 1];
  const uint8_t* base;
  int h, w;
  if (a.images) {
    const ImageRec r = a.images[img];
    base = r.base; h = r.height; w = r.width;
    c.plane = r.plane0 + c.channel;
  } else {
    base = a.base + (long long)img * a.image_stride; h = a.height; w = a.width;
    c.plane = img * a.channels + c.channel;
  }
  if (img >= a.n_images || y < 0 || x < 0 || (y + 1) * a.patch > h || (x + 1) * a.patch > w) {
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
// xform kernel (uint8 -> float32 / Focus)
// ------------------------------------------------------------------------------------------
enum XformMode { kNormPlain = 0, kF32Focus = 1, kNormFocus = 2 };

template <int kMode>
__device__ __forceinline__ void xform_chunk(const GatherArgs& a, const Chunk& c, const uint8_t* stage, int tid,
                                            int nthreads) {
  const int P = a.patch;
  float* out_item = reinterpret_cast<float*>(a.out + (long long)c.item * a.out_item_stride);
  const bool zero = (c.src == nullptr);
  if (kMode == kNormPlain) {
    // chunk = rows*P bytes in, rows*P floats out, both contiguous
    float* dst = out_item + ((long long)c.channel * P + c.row0) * P;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(stage);
    const int n = a.rows * P / 4;
#pragma unroll 4
    for (int i = tid; i < n; i += nthreads) {
      const uint32_t u = zero ? 0u : w[i];
      st_f4(dst + 4 * (long long)i, u8_to_unit((float)(u & 0xFF)), u8_to_unit((float)((u >> 8) & 0xFF)),
            u8_to_unit((float)((u >> 16) & 0xFF)), u8_to_unit((float)(u >> 24)));
    }
  } else {
    // Focus: out[(dy + 2*dx) * C + ch][i][j] = tile[ch][2i + dy][2j + dx]
    const int half = P / 2;
    const int q4 = P / 4;  // 4-pixel groups per row
    const int n = a.rows * q4;
    const long long plane = (long long)half * half;
#pragma unroll 2
    for (int i = tid; i < n; i += nthreads) {
      const int r = i / q4, g = i - r * q4;
      const int y = c.row0 + r;
      float e0, o0, e1, o1;
      if (kMode == kF32Focus) {
        const float4 v = zero ? make_float4(0.f, 0.f, 0.f, 0.f)
                              : reinterpret_cast<const float4*>(stage)[i];
        e0 = v.x; o0 = v.y; e1 = v.z; o1 = v.w;
      } else {
        const uint32_t u = zero ? 0u : reinterpret_cast<const uint32_t*>(stage)[i];
        e0 = u8_to_unit((float)(u & 0xFF)); o0 = u8_to_unit((float)((u >> 8) & 0xFF));
        e1 = u8_to_unit((float)((u >> 16) & 0xFF)); o1 = u8_to_unit((float)(u >> 24));
      }
      const int dy = y & 1;
      float* even = out_item + ((long long)(dy * a.channels + c.channel)) * plane + (long long)(y >> 1) * half + 2 * g;
      float* odd = even + 2ll * a.channels * plane;
      st_f2(even, e0, e1);
      st_f2(odd, o0, o1);
    }
  }
}

template <int kMode, int kStages, int kConsumerWarps, bool kTensor>
__global__ void __launch_bounds__((kConsumerWarps + 1) * 32)
gather_xform_kernel(const __grid_constant__ GatherArgs a, const __grid_constant__ CUtensorMap map0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t chunk_bytes = (uint32_t)a.rows * a.patch * a.elem;
  const uint32_t row_bytes = (uint32_t)a.patch * a.elem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * chunk_bytes);
  uint64_t* empty = full + kStages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kConsumerWarps);
    }
    fence_mbar_init();
    if (kTensor) prefetch_tensormap(&map0);
  }
  __syncthreads();

  const int grid = gridDim.x;
  const int mine = a.total_chunks > (int)blockIdx.x ? (a.total_chunks - (int)blockIdx.x + grid - 1) / grid : 0;

  if (warp == 0) {  // ---- TMA producer
    for (int j = 0; j < mine; ++j) {
      const int st = j % kStages;
      uint8_t* stage = smem + (size_t)st * chunk_bytes;
      if (j >= kStages) mbar_wait(&empty[st], (uint32_t)(j / kStages - 1) & 1u);
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
        mbar_arrive(&full[st]);
      }
    }
  } else {  // ---- consumers
    const int tid = threadIdx.x - 32;
    for (int j = 0; j < mine; ++j) {
      const int st = j % kStages;
      Chunk c;
      const bool skip = decode_chunk(a, (int)blockIdx.x + j * grid, c);
      mbar_wait(&full[st], (uint32_t)(j / kStages) & 1u);
      if (!skip) xform_chunk<kMode>(a, c, smem + (size_t)st * chunk_bytes, tid, kConsumerWarps * 32);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
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
    if (zero_row && a.skip_negative) continue;
    if (!zero_row) {
      y = a.positions[2 * (long long)item]; x = a.positions[2 * (long long)item + 1];
      if (a.images) {
        const ImageRec rec = a.images[img];
        base = rec.base; h = rec.height; w = rec.width;
      } else {
        base = a.base + (long long)img * a.image_stride; h = a.height; w = a.width;
      }
      if (img >= a.n_images || y < 0 || x < 0 || (y + 1) * P > h || (x + 1) * P > w) {
        if (a.status && lane == 0 && ch == 0 && r == 0) atomicOr(a.status, 1);
        continue;  // out-of-grid position: tile skipped
      }
    }
    const long long sy = zero_row ? -1 : y * P + r - (a.shifts ? a.shifts[2 * img] : 0);
    const long long sx0 = zero_row ? 0 : x * P - (a.shifts ? a.shifts[2 * img + 1] : 0);
    const bool row_ok = !zero_row && sy >= 0 && sy < h;
    const uint8_t* row = row_ok ? base + (((long long)ch * h + sy) * w) * a.elem : nullptr;
    uint8_t* dst_item = a.out + (long long)item * a.out_item_stride;
    for (int g = lane; g < P / 4; g += 32) {
      float f[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t packed = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long sx = sx0 + 4 * g + k;
        if (row_ok && sx >= 0 && sx < w) {
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
    if (img < 0 && a.skip_negative) continue;
    if (img >= 0) {
      const long long y = a.positions[2 * (long long)item], x = a.positions[2 * (long long)item + 1];
      const uint8_t* base;
      int h, w;
      if (a.images) {
        const ImageRec rec = a.images[img];
        base = rec.base; h = rec.height; w = rec.width;
      } else {
        base = a.base + (long long)img * a.image_stride; h = a.height; w = a.width;
      }
      if (img >= a.n_images || y < 0 || x < 0 || (y + 1) * P > h || (x + 1) * P > w) {
        if (a.status && rem == 0 && ch == 0) atomicOr(a.status, 1);
        continue;
      }
      const long long sy = y * P + r - (a.shifts ? a.shifts[2 * img] : 0);
      const long long sx = x * P + col - (a.shifts ? a.shifts[2 * img + 1] : 0);
      if (sy >= 0 && sy < h && sx >= 0 && sx < w) {  // outside the (translated) image: zero fill
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
