// C ABI of jolineedle_b200 (see include/jolineedle_b200.h): argument validation, TMA tensor-map
// encoding, launch configuration.  No torch types cross this boundary.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/jolineedle_b200.h"
#include "jn_env.cuh"
#include "jn_gather.cuh"
#include "jn_pyramid.cuh"
#include "jn_scan.cuh"

namespace {

thread_local char g_error[512] = "";
// Kernel launches issued by this library since it was loaded (bench.py reports the launches of its timed region).
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

#define JN_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t err__ = (expr);                                                                    \
    if (err__ != cudaSuccess)                                                                      \
      return fail(JN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
  } while (0)

#define JN_LAUNCHED()                                      \
  do {                                                     \
    g_launches.fetch_add(1, std::memory_order_relaxed);    \
    JN_CUDA(cudaGetLastError());                           \
  } while (0)

#define JN_REQUIRE(cond, ...)                              \
  do {                                                     \
    if (!(cond)) return fail(JN_ERR_INVALID, __VA_ARGS__); \
  } while (0)

struct DeviceInfo {
  int device = -1, sm_count = 0, cc_major = 0, cc_minor = 0, max_smem_optin = 0;
};

int current_device_info(DeviceInfo& info) {
  static std::mutex mu;
  static std::vector<DeviceInfo> cache;
  int dev = 0;
  JN_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  for (const auto& d : cache)
    if (d.device == dev) { info = d; return JN_OK; }
  DeviceInfo d;
  d.device = dev;
  JN_CUDA(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
  JN_CUDA(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  JN_CUDA(cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  JN_CUDA(cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  cache.push_back(d);
  info = d;
  return JN_OK;
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// Work counters of the xform kernel: {next unclaimed chunk, CTAs done} pairs that are zero between
// launches (the kernel cleans up after itself).  Launches rotate through a pool so that gathers running
// concurrently on different streams do not share a pair.  Allocated once per device.
int* next_work_counter(int device) {
  constexpr int kSlots = 1024;
  struct Pool { int device; int* base; unsigned seq; };
  static std::mutex mu;
  static std::vector<Pool> pools;
  std::lock_guard<std::mutex> lock(mu);
  for (auto& p : pools)
    if (p.device == device) return p.base + 2 * (p.seq++ % kSlots);
  int* base = nullptr;
  if (cudaMalloc(&base, kSlots * 2 * sizeof(int)) != cudaSuccess ||
      cudaMemset(base, 0, kSlots * 2 * sizeof(int)) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  pools.push_back({device, base, 1u});
  return base;
}

// One launch through cudaLaunchKernelEx; `pdl` adds the programmatic-stream-serialization attribute: the kernel
// may start while the previous kernel of the stream is still running (after that one's
// griddepcontrol.launch_dependents), and orders itself with griddepcontrol.wait where it needs to.
template <typename... Params, typename... Args>
cudaError_t launch_kernel(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                          Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  static const bool pdl_enabled = [] { const char* e = getenv("JN_PDL"); return !(e && e[0] == '0'); }();  // A/B knob
  if (pdl && pdl_enabled) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
}

inline int grid_for(long long work_items, int per_block, int cap) {
  long long g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

}  // namespace

// ------------------------------------------------------------------------------------------
struct jn_images {
  int n_slabs = 0, n_images = 0, channels = 0, dtype = 0, elem = 0, patch = 0;
  // single slab
  const uint8_t* base = nullptr;
  int height = 0, width = 0;
  long long image_stride = 0;
  // several slabs
  jnk::ImageRec* d_recs = nullptr;
  bool owns_recs = true;
  bool padded = false;  // sizes are rounded up to the patch grid; pixels outside the image read as zeros
  bool host_mapped = false;  // slab 0 is page-locked HOST memory: the gathers read it over PCIe
  // engines
  bool bulk_ok = false, tensor_ok = false;
  int box_w = 0, kbox = 0;
  int device = 0;
  // tensor maps by chunk geometry, encoded on first use (see cached_map)
  struct CachedMap { int kind, rows, pitch; CUtensorMap map; };
  mutable std::mutex mu;
  mutable std::vector<CachedMap> maps;
};

namespace {

// Largest divisor of `patch` that is <= 256 elements and a multiple of 16 bytes.
int pick_box_width(int patch, int elem) {
  for (int w = patch < 256 ? patch : 256; w >= 1; --w)
    if (patch % w == 0 && (w * elem) % 16 == 0) return w;
  return 0;
}

// Rows per chunk: largest divisor of patch (even when `need_even`) with rows*patch*elem <= target.
int pick_rows(int patch, int elem, int target_bytes, bool need_even) {
  int best = 0;
  for (int r = 1; r <= patch && r <= 256; ++r) {
    if (patch % r) continue;
    if (need_even && (r & 1)) continue;
    if ((long long)r * patch * elem <= target_bytes) best = r;
  }
  return best;
}

// L2 promotion of the tensor maps: how far the L2 widens a TMA read beyond the bytes asked for.  Tile rows are
// short (128 ... 448 bytes of uint8) and start anywhere on a 64-byte grid, so 256-byte promotion fetched 29-33 % more
// from DRAM than the tiles hold (ncu: 0.79 GB read for 0.62 GB of cfg-3 tiles); 64 bytes reads what is asked for
// (profiles/r02/sweep_l2_promotion.jsonl: uint8 P = 448 0.945 -> 0.991 of the HBM peak, float32 unchanged).
// JN_TMA_L2_PROMOTION = none | 64 | 128 | 256 overrides the choice (A/B knob).
CUtensorMapL2promotion l2_promotion(CUtensorMapL2promotion chosen) {
  static const int forced = [] {
    const char* e = std::getenv("JN_TMA_L2_PROMOTION");
    if (!e || !*e) return -1;
    if (!std::strcmp(e, "none")) return (int)CU_TENSOR_MAP_L2_PROMOTION_NONE;
    if (!std::strcmp(e, "64")) return (int)CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
    if (!std::strcmp(e, "128")) return (int)CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    if (!std::strcmp(e, "256")) return (int)CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    return -1;
  }();
  return forced < 0 ? chosen : (CUtensorMapL2promotion)forced;
}

int encode_slab_map(CUtensorMap* map, const void* ptr, int elem, int n_planes, int height, int width, int box_w,
                    int kbox, int rows, bool three_d = false) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(JN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  if (three_d) {
    // 3-D view [W] x [H] x [planes], box = [patch] x [rows] x 1: element-granular x / y offsets (used by
    // translated gathers; pixels outside the image are zero-filled by the TMA unit)
    cuuint64_t dims3[3] = {(cuuint64_t)width, (cuuint64_t)height, (cuuint64_t)n_planes};
    cuuint64_t strides3[2] = {(cuuint64_t)width * elem, (cuuint64_t)width * elem * height};
    cuuint32_t box3[3] = {(cuuint32_t)(box_w * kbox), (cuuint32_t)rows, 1};
    cuuint32_t estr3[3] = {1, 1, 1};
    const CUtensorMapDataType dt3 = elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    CUresult r3 = fn(map, dt3, 3, const_cast<void*>(ptr), dims3, strides3, box3, estr3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, l2_promotion(CU_TENSOR_MAP_L2_PROMOTION_L2_64B), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r3 != CUDA_SUCCESS) return fail(JN_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int)r3);
    return JN_OK;
  }
  // 4-D view (innermost first): [box_w] x [width / box_w] x [height] x [planes]
  cuuint64_t dims[4] = {(cuuint64_t)box_w, (cuuint64_t)(width / box_w), (cuuint64_t)height, (cuuint64_t)n_planes};
  cuuint64_t strides[3] = {(cuuint64_t)box_w * elem, (cuuint64_t)width * elem, (cuuint64_t)width * elem * height};
  cuuint32_t box[4] = {(cuuint32_t)box_w, (cuuint32_t)kbox, (cuuint32_t)rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapDataType dt = elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUresult r = fn(map, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, l2_promotion(CU_TENSOR_MAP_L2_PROMOTION_L2_64B), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(JN_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (dims %llu x %llu x %llu x %llu, box %u x %u x %u)",
                (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                (unsigned long long)dims[3], box[0], box[1], box[2]);
  return JN_OK;
}

// Translated gathers with arbitrary x offsets: 3-D view [W*elem/8] x [H] x [planes] of 8-byte elements,
// box = [pitch/8] x [rows] x 1 with pitch = patch*elem + 16: the kernel asks for the 16-byte aligned
// superset of every tile row; bytes outside the image arrive as zeros (rows are multiples of 8 bytes,
// so the zero fill is exact at pixel granularity).
int encode_shift_map(CUtensorMap* map, const void* ptr, int elem, int n_planes, int height, int width, int pitch,
                     int rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(JN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t row_bytes = (cuuint64_t)width * elem;
  cuuint64_t dims[3] = {row_bytes / 8, (cuuint64_t)height, (cuuint64_t)n_planes};
  cuuint64_t strides[2] = {row_bytes, row_bytes * height};
  cuuint32_t box[3] = {(cuuint32_t)(pitch / 8), (cuuint32_t)rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  l2_promotion(CU_TENSOR_MAP_L2_PROMOTION_L2_64B), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(JN_ERR_CUDA, "cuTensorMapEncodeTiled (translated) failed with CUresult %d", (int)r);
  return JN_OK;
}

}  // namespace

namespace {

constexpr int kXformWarps = 8;

// Pipeline shape of the gather kernels: stages of the shared-memory ring, load lookahead (copy
// kernel), bytes per chunk and resident CTAs per SM.  The defaults come from sweeps on B200
// (tools/tune_gather.py, profiles/); JN_GATHER_TUNE="copy_stages,copy_ahead,copy_chunk_bytes,
// copy_ctas_per_sm,xform_stages,xform_chunk_bytes,xform_ctas_per_sm,xform_batch" overrides them (0 keeps a
// default).
struct GatherTune {
  int copy_stages = 6, copy_ahead = 3, copy_chunk = 32768, copy_ctas = 0;
  int xform_stages = 4, xform_chunk = 0, xform_ctas = 0;
  int xform_batch = 4;  // chunks per ticket for the bulk of a launch (power of two <= 32), see claim_schedule
};

GatherTune gather_tune(int patch, int elem, bool plain_copy, bool focus, int engine) {
  GatherTune t;
  (void)engine; (void)plain_copy;
  // Buckets by patch size (measured at P = 128 / 256 / 448 / 1024, both TMA engines; the entry is the
  // setting whose *worse* engine was best -- profiles/r01/tune_sweep_summary.txt).  Shallow rings
  // with several CTAs per SM beat one deep ring: 2-3 stages, lookahead 1, 2-4 CTAs/SM.
  const int bucket = patch <= 128 ? 0 : patch <= 256 ? 1 : patch < 1024 ? 2 : 3;
  static const int copy[4][4] = {{2, 1, 32768, 3}, {3, 1, 32768, 2}, {3, 1, 32768, 2}, {2, 1, 16384, 3}};
  // xform kernel (dynamic chunk claiming, profiles/r01/tune_xform_v2_summary.txt): {stages, chunk bytes, CTAs/SM};
  // the sweep is flat within 2-3 % around these
  static const int norm_plain[4][3] = {{3, 4096, 3}, {3, 8192, 2}, {3, 16384, 1}, {4, 8192, 1}};
  static const int norm_focus[4][3] = {{3, 4096, 3}, {3, 8192, 2}, {3, 16384, 1}, {4, 8192, 1}};
  static const int f32_focus[4][3] = {{4, 4096, 4}, {3, 16384, 2}, {6, 16384, 1}, {3, 32768, 1}};
  t.copy_stages = copy[bucket][0]; t.copy_ahead = copy[bucket][1]; t.copy_chunk = copy[bucket][2];
  t.copy_ctas = copy[bucket][3];
  const int(*x)[3] = elem == 4 ? f32_focus : (focus ? norm_focus : norm_plain);
  t.xform_stages = x[bucket][0]; t.xform_chunk = x[bucket][1]; t.xform_ctas = x[bucket][2];
  // chunks per ticket: 4 with tensor tiles (one instruction per chunk), 8 with per-row bulk copies (lists of
  // images: the producer issues a copy per row, a few more chunks per decode amortise that) -- r02 sweep
  t.xform_batch = engine == JN_ENGINE_BULK ? 8 : 4;
  if (const char* env = getenv("JN_GATHER_TUNE")) {
    int v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    sscanf(env, "%d,%d,%d,%d,%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]);
    if (v[0] > 0) t.copy_stages = v[0];
    if (v[1] > 0) t.copy_ahead = v[1];
    if (v[2] > 0) t.copy_chunk = v[2];
    if (v[3] > 0) t.copy_ctas = v[3];
    if (v[4] > 0) t.xform_stages = v[4];
    if (v[5] > 0) t.xform_chunk = v[5];
    if (v[6] > 0) t.xform_ctas = v[6];
    if (v[7] > 0) t.xform_batch = v[7];
  }
  int b = 1;
  while (b * 2 <= t.xform_batch && b < 32) b *= 2;  // power of two
  t.xform_batch = b;
  (void)plain_copy;
  return t;
}

// Resident CTAs per SM of a persistent gather kernel at a given shared-memory size: asked of the runtime once
// per (kernel, size, device) -- cudaFuncSetAttribute + the occupancy query cost ~3 us per launch otherwise,
// which matters for the 256-tile launches of the batched env.
template <typename Kernel>
int persistent_ctas_per_sm(Kernel kernel, int threads, size_t smem, int device, int* per_sm) {
  struct Entry { const void* fn; size_t smem; int device, per_sm; };
  struct Limit { const void* fn; int device; size_t smem; };  // largest dynamic size the function was opted into
  static std::mutex mu;
  static std::vector<Entry> cache;
  static std::vector<Limit> limits;
  const void* fn = reinterpret_cast<const void*>(kernel);
  std::lock_guard<std::mutex> lock(mu);
  for (const auto& e : cache)
    if (e.fn == fn && e.smem == smem && e.device == device) { *per_sm = e.per_sm; return JN_OK; }
  // the opt-in limit is one value per function: only ever raise it, or a later launch at an earlier, larger
  // size would be refused
  Limit* lim = nullptr;
  for (auto& l : limits)
    if (l.fn == fn && l.device == device) lim = &l;
  if (!lim) { limits.push_back({fn, device, 0}); lim = &limits.back(); }
  if (smem > lim->smem) {
    JN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lim->smem = smem;
  }
  int n = 0;
  JN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem));
  cache.push_back({fn, smem, device, n});
  *per_sm = n;
  return JN_OK;
}

// Ticket schedule of the converting gather (gather_xform_kernel): which chunks a ticket stands for.  The
// producer claims tickets three batches ahead (that is what hides the decode latency), so a CTA that falls
// behind still owns three batches when the tickets run out.  The bulk of the launch goes out in batches of `head`
// chunks (a power of two <= 32, picked per pipeline shape: GatherTune::xform_batch), followed, per CTA, by two
// batches each of head/2 ... 4 chunks, three of 2 and six single chunks -- so that what a slow CTA still holds is
// about what every other CTA draws from the fine-grained end.  A launch too short for all of it drops the large
// sizes first.  Segments are stored in launch order.
void claim_schedule(jnk::GatherArgs& a, int grid, int head) {
  int sizes[5], per_cta[5], n_tail = 0;
  for (int sz = 1; sz < head && n_tail < 5; sz *= 2) {
    sizes[n_tail] = sz;
    per_cta[n_tail] = sz == 1 ? 6 : sz == 2 ? 3 : 2;
    ++n_tail;
  }
  int rem = a.total_chunks, counts[5] = {0, 0, 0, 0, 0};
  for (int j = 0; j < n_tail && rem > 0; ++j) {
    long long want = (long long)per_cta[j] * grid;  // batches
    const bool all = want * sizes[j] <= rem;
    if (!all) want = rem / sizes[j];
    counts[j] = (int)want;
    rem -= (int)want * sizes[j];
    if (!all) break;  // out of work: what is left (< sizes[j] chunks) opens the launch
  }
  int n = 0, ticket = 0, chunk = 0;
  auto push = [&](int size, int batches, int chunks) {
    if (batches <= 0) return;
    a.sched_size[n] = size; a.sched_ticket[n] = ticket; a.sched_chunk[n] = chunk;
    ticket += batches; chunk += chunks; ++n;
  };
  push(head, (rem + head - 1) / head, rem);  // head, the last batch possibly partial
  for (int j = n_tail - 1; j >= 0; --j) push(sizes[j], counts[j], counts[j] * sizes[j]);
  a.sched_n = n;
  for (int j = n; j < 7; ++j) { a.sched_ticket[j] = ticket; a.sched_chunk[j] = chunk; }
  for (int j = n; j < 6; ++j) a.sched_size[j] = 1;
}

template <typename Kernel, typename... Extra>
int launch_persistent(Kernel kernel, const jnk::GatherArgs& args, const CUtensorMap& map, int threads, size_t smem,
                      const DeviceInfo& dev, cudaStream_t stream, int ctas_cap, bool pdl, int batch, Extra... extra) {
  int per_sm = 0;
  if (int rc = persistent_ctas_per_sm(kernel, threads, smem, dev.device, &per_sm)) return rc;
  if (per_sm < 1) return fail(JN_ERR_CUDA, "gather kernel does not fit on an SM (%zu bytes of shared memory)", smem);
  if (ctas_cap > 0 && per_sm > ctas_cap) per_sm = ctas_cap;
  long long grid = (long long)dev.sm_count * per_sm;
  if (grid > args.total_chunks) grid = args.total_chunks;
  jnk::GatherArgs a = args;
  if (a.work_counter) {
    claim_schedule(a, (int)grid, batch);
    if (a.sched_ticket[a.sched_n] < grid) grid = a.sched_ticket[a.sched_n];  // never more CTAs than tickets
  }
  JN_CUDA(launch_kernel(kernel, dim3((unsigned)grid), dim3((unsigned)threads), smem, stream, pdl, a, map, extra...));
  return JN_OK;
}

}  // namespace

extern "C" {

int jn_abi_version(void) { return JN_ABI_VERSION; }
const char* jn_last_error(void) { return g_error; }

int jn_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  DeviceInfo d;
  if (int rc = current_device_info(d)) return rc;
  if (sm_count) *sm_count = d.sm_count;
  if (cc_major) *cc_major = d.cc_major;
  if (cc_minor) *cc_minor = d.cc_minor;
  return JN_OK;
}

// Host-only self test of the two pieces of arithmetic shared with the device code: the
// direction table and the uint8 -> [0,1] normalisation.  `unit_out` receives the 256 values.
int jn_selftest_host(float* unit_out /*HOST [256]*/, int* direction_out /*HOST [9], index (sy+1)*3+(sx+1)*/) {
  for (int i = 0; i < 256; ++i) {
    const float a = jnk::u8_to_unit((float)i), b = jnk::unit_from_scaled((float)i / 256.0f);
    if (a != b) return fail(JN_ERR_INVALID, "normalisation formulas disagree for %d: %.9g vs %.9g", i, a, b);
    if (unit_out) unit_out[i] = b;
  }
  if (direction_out)
    for (int sy = -1; sy <= 1; ++sy)
      for (int sx = -1; sx <= 1; ++sx) direction_out[(sy + 1) * 3 + (sx + 1)] = jnk::direction_code(sy * 3, sx * 5);
  return JN_OK;
}

// ------------------------------------------------------------------------------------------
// image sets
// ------------------------------------------------------------------------------------------
static int images_create(jn_images** out, int n_slabs, const void* const* slab_ptrs, const int32_t* counts,
                         const int32_t* heights, const int32_t* widths, int channels, int dtype, int patch_size,
                         void* table_host, void* table_dev, void* stream, bool padded) {
  JN_REQUIRE(out != nullptr, "jn_images_create: out is NULL");
  *out = nullptr;
  JN_REQUIRE(n_slabs >= 1 && slab_ptrs && counts && heights && widths, "jn_images_create: empty image set");
  JN_REQUIRE(dtype == JN_U8 || dtype == JN_F32, "jn_images_create: dtype must be JN_U8 or JN_F32");
  JN_REQUIRE(channels >= 1 && patch_size >= 1, "jn_images_create: channels and patch_size must be positive");
  const int elem = dtype == JN_F32 ? 4 : 1;
  long long total = 0;
  bool aligned = (patch_size * elem) % 16 == 0;
  for (int k = 0; k < n_slabs; ++k) {
    JN_REQUIRE(slab_ptrs[k] != nullptr && counts[k] >= 1, "jn_images_create: slab %d is empty", k);
    // same precondition as the reference envs (general_env.py:50-51, simple_env.py:68-69)
    JN_REQUIRE(heights[k] > 0 && widths[k] > 0 &&
                   (padded || (heights[k] % patch_size == 0 && widths[k] % patch_size == 0)),
               "image size %dx%d is not a multiple of patch_size %d", heights[k], widths[k], patch_size);
    total += counts[k];
    aligned = aligned && (reinterpret_cast<uintptr_t>(slab_ptrs[k]) % 16 == 0) && ((long long)widths[k] * elem) % 16 == 0;
  }
  JN_REQUIRE(total < (1ll << 30), "jn_images_create: too many images");

  jn_images* s = new jn_images();
  s->n_slabs = n_slabs; s->n_images = (int)total; s->channels = channels; s->dtype = dtype; s->elem = elem;
  s->patch = patch_size;
  s->padded = padded;
  cudaGetDevice(&s->device);
  {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, slab_ptrs[0]) == cudaSuccess) s->host_mapped = attr.type == cudaMemoryTypeHost;
    else cudaGetLastError();
  }
  s->bulk_ok = aligned;
  s->box_w = aligned ? pick_box_width(patch_size, elem) : 0;
  s->kbox = s->box_w ? patch_size / s->box_w : 0;
  s->tensor_ok = aligned && s->box_w > 0 && encode_tiled_fn() != nullptr;

  auto bail = [&](int rc) { jn_images_destroy(s); return rc; };
  if (n_slabs == 1) {
    s->base = static_cast<const uint8_t*>(slab_ptrs[0]);
    s->height = heights[0]; s->width = widths[0];
    s->image_stride = (long long)channels * heights[0] * widths[0] * elem;
  } else {
    std::vector<jnk::ImageRec> recs;
    recs.reserve((size_t)total);
    for (int k = 0; k < n_slabs; ++k)
      for (int i = 0; i < counts[k]; ++i) {
        jnk::ImageRec r;
        r.base = static_cast<const uint8_t*>(slab_ptrs[k]) + (long long)i * channels * heights[k] * widths[k] * elem;
        r.height = heights[k]; r.width = widths[k]; r.slab = k; r.plane0 = i * channels;
        recs.push_back(r);
      }
    if (table_host && table_dev) {
      // Caller-owned scratch: the table is written into `table_host` here and the CALLER copies it
      // to `table_dev` on its stream before the first gather (with its own pinned-memory
      // bookkeeping).  No cudaMalloc / cudaFree / memcpy in the library: nothing synchronises.
      memcpy(table_host, recs.data(), recs.size() * sizeof(jnk::ImageRec));
      s->d_recs = static_cast<jnk::ImageRec*>(table_dev);
      s->owns_recs = false;
    } else {
      if (cudaMalloc(&s->d_recs, recs.size() * sizeof(jnk::ImageRec)) != cudaSuccess)
        return bail(fail(JN_ERR_CUDA, "cudaMalloc(image records) failed: %s", cudaGetErrorString(cudaGetLastError())));
      // pageable source: the runtime stages the bytes before returning, so `recs` may die right after
      if (cudaMemcpyAsync(s->d_recs, recs.data(), recs.size() * sizeof(jnk::ImageRec), cudaMemcpyHostToDevice,
                          (cudaStream_t)stream) != cudaSuccess)
        return bail(fail(JN_ERR_CUDA, "upload of image records failed: %s", cudaGetErrorString(cudaGetLastError())));
    }
  }
  *out = s;
  return JN_OK;
}

int jn_images_create(jn_images** out, int n_slabs, const void* const* slab_ptrs, const int32_t* counts,
                     const int32_t* heights, const int32_t* widths, int channels, int dtype, int patch_size,
                     void* table_host, void* table_dev, void* stream) {
  return images_create(out, n_slabs, slab_ptrs, counts, heights, widths, channels, dtype, patch_size, table_host,
                       table_dev, stream, false);
}

int jn_images_create_padded(jn_images** out, int n_slabs, const void* const* slab_ptrs, const int32_t* counts,
                            const int32_t* heights, const int32_t* widths, int channels, int dtype, int patch_size,
                            void* table_host, void* table_dev, void* stream) {
  return images_create(out, n_slabs, slab_ptrs, counts, heights, widths, channels, dtype, patch_size, table_host,
                       table_dev, stream, true);
}

void jn_images_destroy(jn_images* s) {
  if (!s) return;
  if (s->d_recs && s->owns_recs) cudaFree(s->d_recs);
  delete s;
}

int jn_images_tma_ok(const jn_images* s, int engine) {
  if (!s) return 0;
  if (engine == JN_ENGINE_TENSOR) return s->tensor_ok && s->n_slabs == 1 ? 1 : 0;
  if (engine == JN_ENGINE_BULK) return s->bulk_ok ? 1 : 0;
  return 1;
}

// ------------------------------------------------------------------------------------------
// K1 gather
// ------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

// What a gather is asked to do: the arguments of jn_gather plus the fused-step extras of jn_env_step_gather.
struct GatherRequest {
  const int64_t* positions = nullptr;
  const int32_t* src_index = nullptr;
  const int32_t* shifts = nullptr;
  int n_items = 0;
  void* out = nullptr;
  int64_t out_item_stride_bytes = 0;
  uint32_t flags = 0;
  int engine = JN_ENGINE_AUTO;
  int32_t* status = nullptr;
  const int64_t* actions = nullptr;  // fused step: positions are pre-move, the kernel applies the action itself
  int grid_rows = 0, grid_cols = 0;
  bool pdl = false;         // launch with the programmatic-serialization attribute (behind the step kernel)
  bool wait_prior = false;  // ... and wait for that kernel before the first index load
};

// Tensor map of an image set for one chunk geometry; encoded once and kept with the set
// (cuTensorMapEncodeTiled costs a few microseconds, as much as the launch of a small gather).
int cached_map(const jn_images* set, int kind /*0 tiled 4-D, 1 tiled 3-D, 2 superset rows*/, int rows, int pitch,
               CUtensorMap* out) {
  std::lock_guard<std::mutex> lock(set->mu);
  for (const auto& m : set->maps)
    if (m.kind == kind && m.rows == rows && m.pitch == pitch) { *out = m.map; return JN_OK; }
  jn_images::CachedMap m;
  m.kind = kind; m.rows = rows; m.pitch = pitch;
  memset(&m.map, 0, sizeof(m.map));
  const int planes = set->n_images * set->channels;
  if (int rc = kind == 2 ? encode_shift_map(&m.map, set->base, set->elem, planes, set->height, set->width, pitch, rows)
                         : encode_slab_map(&m.map, set->base, set->elem, planes, set->height, set->width, set->box_w,
                                           set->kbox, rows, kind == 1))
    return rc;
  set->maps.push_back(m);
  *out = m.map;
  return JN_OK;
}

int gather_launch(const jn_images* set, const GatherRequest& rq, cudaStream_t stream) {
  const int64_t* positions = rq.positions;
  const int32_t *src_index = rq.src_index, *shifts = rq.shifts;
  const int n_items = rq.n_items;
  void* out = rq.out;
  const int64_t out_item_stride_bytes = rq.out_item_stride_bytes;
  const uint32_t flags = rq.flags;
  int engine = rq.engine;
  int32_t* status = rq.status;
  JN_REQUIRE(set != nullptr, "jn_gather: image set is NULL");
  JN_REQUIRE(n_items >= 0, "jn_gather: negative item count");
  if (n_items == 0) return JN_OK;
  JN_REQUIRE(out != nullptr, "jn_gather: NULL out");
  JN_REQUIRE(positions != nullptr || rq.actions == nullptr, "jn_gather: actions without positions");
  const bool normalize = (flags & JN_GATHER_NORMALIZE) != 0, focus = (flags & JN_GATHER_FOCUS) != 0;
  JN_REQUIRE(!(normalize && set->dtype != JN_U8), "JN_GATHER_NORMALIZE needs a uint8 image set");
  JN_REQUIRE(!(focus && (set->patch % 2)), "JN_GATHER_FOCUS needs an even patch size");
  JN_REQUIRE(src_index != nullptr || n_items <= set->n_images,
             "jn_gather: %d items but only %d images and no src_index", n_items, set->n_images);
  const int P = set->patch, C = set->channels;
  const int out_elem = (normalize || set->dtype == JN_F32) ? 4 : 1;
  const long long tile_out_bytes = (long long)C * P * P * out_elem;
  JN_REQUIRE(out_item_stride_bytes >= tile_out_bytes, "jn_gather: out_item_stride_bytes %lld < tile size %lld",
             (long long)out_item_stride_bytes, tile_out_bytes);
  DeviceInfo dev;
  if (int rc = current_device_info(dev)) return rc;

  jnk::GatherArgs a;
  memset(&a, 0, sizeof(a));
  a.base = set->base; a.images = set->d_recs;
  a.positions = positions; a.src_index = src_index; a.shifts = shifts;
  a.out = static_cast<uint8_t*>(out); a.out_item_stride = out_item_stride_bytes; a.image_stride = set->image_stride;
  a.status = status; a.n_items = n_items; a.n_images = set->n_images;
  a.channels = C; a.height = set->height; a.width = set->width; a.patch = P; a.elem = set->elem;
  a.box_w = set->box_w; a.kbox = set->kbox;
  a.skip_negative = (flags & JN_GATHER_SKIP_NEGATIVE) ? 1 : 0;
  a.padded = set->padded ? 1 : 0;
  a.actions = rq.actions; a.grid_rows = rq.grid_rows; a.grid_cols = rq.grid_cols;
  a.wait_prior = rq.wait_prior ? 1 : 0;
  // Streaming stores (st.global.cs) for uint8 -> float32 launches whose crops are of the order of the L2 (<= 256 MB):
  // +1.7 % on cfg 4 and +2 % on 256 tiles of 256^2; launches of GBs gain nothing or lose (cfg 2's trajectory
  // gather -1.6 %), float32 pass-through -0.5 % (profiles/r02/sweep_l2_hint.jsonl, stream_stores_ab.jsonl).
  // JN_STREAM_STORES = 0 / 1 forces.
  static const int stream_stores = [] { const char* e = std::getenv("JN_STREAM_STORES"); return e && *e ? std::atoi(e) : -1; }();
  a.stream_stores = stream_stores >= 0 ? stream_stores
                    : (((rq.flags & JN_GATHER_NORMALIZE) && (long long)n_items * tile_out_bytes <= (256ll << 20)) ? 1 : 0);

  const bool plain_copy = !normalize && !focus;
  const bool out_aligned = reinterpret_cast<uintptr_t>(out) % 16 == 0 && out_item_stride_bytes % 16 == 0;
  // u8 -> u8 Focus has no TMA kernel (nobody asks for it); it runs on the LDG engine.
  // (the xform kernel divides by patch / 4 with a 32-bit multiply-high: patches of at least 8 pixels)
  const bool tma_mode = plain_copy || (P >= 8 && (normalize || (focus && set->dtype == JN_F32)));
  // Translated sources (zero fill outside the image), one slab:
  //  * plain same-dtype copies whose x shifts the caller vouches for (JN_GATHER_SHIFT_ALIGNED: every
  //    tx * elem % 16 == 0) stay pure DMA: copy kernel, element-typed 3-D map, patch <= 256;
  //  * everything else that converts or re-lays out (and float32 plain copies) goes through the xform
  //    kernel's superset loads: any x offset (the TMA unit itself traps on inner coordinates that are not
  //    16-byte multiples), rows of up to 2032 bytes;
  //  * the rest (lists of images, uint8 -> uint8, wider rows) is served by the plain-load engine.
  // Padded sets (sizes that are not multiples of the patch) need the same out-of-image zero fill: they ride the
  // superset loads as well, translated or not.
  const bool shift_copy_ok = plain_copy && set->tensor_ok && set->n_slabs == 1 && set->kbox == 1 &&
                             (flags & JN_GATHER_SHIFT_ALIGNED) != 0 && !set->padded;
  const int shift_pitch = P * set->elem + 16;
  const bool shift_xform_ok = set->tensor_ok && set->n_slabs == 1 && shift_pitch / 8 <= 256 && P >= 8 &&
                              (normalize || set->dtype == JN_F32);
  // Plain float32 copies of resident single-slab sets with tiles of 256 pixels or more: the converting kernel's
  // float32 pass-through mode (tickets, tensor tiles) instead of the pure-DMA copy kernel with its static round
  // robin -- 1.03 / 1.04 / 1.06 of the measured copy peak at P = 256 / 448 / 1024 against 1.00 / 1.00 / 0.99
  // (profiles/r02/f32_plain_routes.jsonl); smaller tiles stay on the copy kernel (0.92 vs 0.82 at P = 128).
  // JN_F32_PLAIN=copy restores the old route (A/B).  Only under `auto`: an explicit engine means the copy kernel.
  static const bool f32_plain_copy_only = [] { const char* e = getenv("JN_F32_PLAIN"); return e && e[0] == 'c'; }();
  const bool f32_via_xform = engine == JN_ENGINE_AUTO && !f32_plain_copy_only && plain_copy && set->dtype == JN_F32 &&
                             P >= 256 && out_aligned && set->tensor_ok && set->n_slabs == 1 && !set->host_mapped &&
                             !shifts && !set->padded;
  if (f32_via_xform) engine = JN_ENGINE_TENSOR;
  if (engine == JN_ENGINE_AUTO) {
    if (!tma_mode || !out_aligned || !set->bulk_ok) engine = JN_ENGINE_LDG;
    else if (shifts || set->padded) engine = (shift_copy_ok || shift_xform_ok) ? JN_ENGINE_TENSOR : JN_ENGINE_LDG;
    else {
      // pure-DMA copy kernel, tiles wider than one TMA box (P > 256: two or more boxes per row): per-row bulk
      // copies measured 1-7 % faster than tensor tiles (profiles/r01/micro_quick_v2.jsonl).  The converting
      // kernel, whose producer feeds the ring in small batches, is better off with ONE tensor-tile instruction
      // per chunk than with a bulk copy per row at every patch size (profiles/r02/sweep_batch_engine.jsonl:
      // +2-12 % at P = 128 / 256 / 448, a tie at 1024)
      // Over PCIe (page-locked host images read in place) whole-row bulk copies keep the wider tiles at the
      // link rate where tensor tiles of uint8 pixels lose a tenth (profiles/r02/zero_copy_pcie.jsonl: 50.1 vs
      // 44.9 GB/s at P = 448).
      const bool prefer_bulk = set->kbox > 1 && (plain_copy || set->host_mapped);
      engine = (set->tensor_ok && set->n_slabs == 1 && !prefer_bulk) ? JN_ENGINE_TENSOR : JN_ENGINE_BULK;
    }
  }
  if ((shifts || set->padded) && engine == JN_ENGINE_TENSOR)
    JN_REQUIRE(shift_copy_ok || shift_xform_ok,
               "translated gathers on the tensor engine need one slab and tile rows of at most 2032 bytes "
               "(uint8 -> uint8 copies: patch_size <= 256 and JN_GATHER_SHIFT_ALIGNED)");
  if ((shifts || set->padded) && engine == JN_ENGINE_BULK)
    return fail(JN_ERR_INVALID, "the bulk engine cannot translate or pad (no zero fill outside the image); use auto");
  if (engine == JN_ENGINE_TENSOR)
    JN_REQUIRE(tma_mode && out_aligned && set->tensor_ok && set->n_slabs == 1,
               "tensor-map engine unavailable for this image set / flags (needs one slab, 16-byte aligned rows)");
  if (engine == JN_ENGINE_BULK)
    JN_REQUIRE(tma_mode && out_aligned && set->bulk_ok, "bulk engine needs 16-byte aligned bases, rows and patches");
  // translated gathers on the xform kernel (superset loads) -- also plain float32 copies that are not vouched for
  const bool shift_xform = (shifts != nullptr || set->padded) && engine == JN_ENGINE_TENSOR && !shift_copy_ok;

  if (engine == JN_ENGINE_LDG) {
    a.rows = 1; a.chunks_per_plane = P; a.total_chunks = 0;
    const bool rows_ok = P % 4 == 0 && out_aligned && !(focus && out_elem == 1);
    if (rows_ok) {  // one warp per tile row, 16-byte stores
      const long long units = (long long)n_items * C * P;
      JN_CUDA(launch_kernel(jnk::gather_rows_kernel, dim3(grid_for(units, 8, dev.sm_count * 8)), dim3(256), 0, stream,
                            rq.pdl, a, out_elem == 4 ? 1 : 0, normalize ? 1 : 0, focus ? 1 : 0));
    } else {  // any patch size / alignment: one thread per element
      const long long total = (long long)n_items * C * P * P;
      JN_CUDA(launch_kernel(jnk::gather_ldg_kernel, dim3(grid_for(total, 256 * 8, dev.sm_count * 16)), dim3(256), 0,
                            stream, rq.pdl, a, out_elem == 4 ? 1 : 0, normalize ? 1 : 0, focus ? 1 : 0));
    }
    return JN_OK;
  }

  // TMA engines: chunk geometry
  const bool use_copy = plain_copy && !shift_xform && !f32_via_xform;
  const GatherTune tune = gather_tune(P, set->elem, use_copy, focus, engine);
  const int target = use_copy ? tune.copy_chunk : tune.xform_chunk;
  const int rows = pick_rows(P, set->elem, target, focus);
  JN_REQUIRE(rows > 0, "patch row of %d bytes does not fit a shared-memory stage", P * set->elem);
  a.rows = rows; a.chunks_per_plane = P / rows;
  const long long total_chunks = (long long)n_items * C * a.chunks_per_plane;
  JN_REQUIRE(total_chunks < (1ll << 31) - (1ll << 22), "jn_gather: too many chunks (%lld)", total_chunks);
  a.total_chunks = (int)total_chunks;
  const size_t chunk_bytes = (size_t)rows * P * set->elem;
  a.pitch = shift_xform ? shift_pitch : P * set->elem;
  a.stage_bytes = (int)(((size_t)rows * a.pitch + 127) / 128 * 128);
  a.wpr_magic = (uint32_t)(((1ull << 32) + (P / 4) - 1) / (P / 4));
  if (!use_copy) {
    a.work_counter = next_work_counter(dev.device);
    if (!a.work_counter) return fail(JN_ERR_CUDA, "cudaMalloc(work counters) failed");
  }

  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (engine == JN_ENGINE_TENSOR) {
    if (int rc = cached_map(set, shift_xform ? 2 : (shifts != nullptr ? 1 : 0), rows, shift_xform ? a.pitch : 0, &map))
      return rc;
  }

  if (use_copy) {
    const size_t smem = tune.copy_stages * chunk_bytes + jnk::kZeroBytes + tune.copy_stages * sizeof(uint64_t);
    const bool tensor = engine == JN_ENGINE_TENSOR;
#define JN_COPY(S, D)                                                                                              \
  if (tune.copy_stages == S && tune.copy_ahead == D)                                                               \
    return tensor ? launch_persistent(jnk::gather_copy_kernel<S, D, true>, a, map, 32, smem, dev, stream,          \
                                      tune.copy_ctas, rq.pdl, 1)                                                    \
                  : launch_persistent(jnk::gather_copy_kernel<S, D, false>, a, map, 32, smem, dev, stream,         \
                                      tune.copy_ctas, rq.pdl, 1);
    JN_COPY(6, 3) JN_COPY(2, 1) JN_COPY(3, 1) JN_COPY(3, 2) JN_COPY(4, 1) JN_COPY(4, 2) JN_COPY(4, 3) JN_COPY(6, 2)
    JN_COPY(6, 4) JN_COPY(6, 5) JN_COPY(8, 4) JN_COPY(8, 6) JN_COPY(12, 6) JN_COPY(12, 9)
#undef JN_COPY
    return fail(JN_ERR_INVALID, "JN_GATHER_TUNE: no copy kernel with %d stages / lookahead %d", tune.copy_stages,
                tune.copy_ahead);
  }
  const int threads = (kXformWarps + 1) * 32;
  const int tensor = engine == JN_ENGINE_TENSOR ? 1 : 0;
  const int stages = tune.xform_stages;
  JN_REQUIRE(stages >= 2 && stages <= 16, "JN_GATHER_TUNE: xform stages must be 2..16, got %d", stages);
  const size_t smem = (size_t)stages * (a.stage_bytes + sizeof(jnk::StageDesc) + 2 * sizeof(uint64_t));
  JN_REQUIRE(smem <= (size_t)dev.max_smem_optin, "xform pipeline of %d x %d bytes does not fit shared memory", stages,
             a.stage_bytes);
#define JN_XFORM(mode)                                                                                             \
  return shift_xform ? launch_persistent(jnk::gather_xform_kernel<mode, kXformWarps, true>, a, map, threads, smem, \
                                         dev, stream, tune.xform_ctas, rq.pdl, tune.xform_batch, stages, tensor)   \
                     : launch_persistent(jnk::gather_xform_kernel<mode, kXformWarps, false>, a, map, threads, smem, \
                                         dev, stream, tune.xform_ctas, rq.pdl, tune.xform_batch, stages, tensor);
  if (normalize && !focus) { JN_XFORM(jnk::kNormPlain) }
  else if (normalize && focus) { JN_XFORM(jnk::kNormFocus) }
  else if (focus) { JN_XFORM(jnk::kF32Focus) }
  else { JN_XFORM(jnk::kF32Plain) }
#undef JN_XFORM
}

}  // namespace

extern "C" {

int jn_gather(const jn_images* set, const int64_t* positions, const int32_t* src_index, const int32_t* shifts,
              int n_items, void* out, int64_t out_item_stride_bytes, uint32_t flags, int engine, int32_t* status,
              void* stream) {
  GatherRequest rq;
  rq.positions = positions; rq.src_index = src_index; rq.shifts = shifts; rq.n_items = n_items; rq.out = out;
  rq.out_item_stride_bytes = out_item_stride_bytes; rq.flags = flags; rq.engine = engine; rq.status = status;
  return gather_launch(set, rq, (cudaStream_t)stream);
}

long long jn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// Host-only view of the converting gather's ticket schedule (tests: the tickets must tile [0, total) exactly).
int jn_claim_schedule_host(int total_chunks, int grid, int batch, int32_t* sizes /*[6]*/, int32_t* tickets /*[7]*/,
                           int32_t* chunks /*[7]*/) {
  JN_REQUIRE(total_chunks >= 1 && grid >= 1 && sizes && tickets && chunks, "jn_claim_schedule_host: bad arguments");
  JN_REQUIRE(batch >= 1 && batch <= 32 && (batch & (batch - 1)) == 0, "jn_claim_schedule_host: batch must be 1, 2, 4 ... 32");
  jnk::GatherArgs a;
  memset(&a, 0, sizeof(a));
  a.total_chunks = total_chunks;
  claim_schedule(a, grid, batch);
  for (int j = 0; j < 6; ++j) sizes[j] = a.sched_size[j];
  for (int j = 0; j < 7; ++j) { tickets[j] = a.sched_ticket[j]; chunks[j] = a.sched_chunk[j]; }
  return a.sched_n;
}

#ifndef JN_SOURCE_HASH
#define JN_SOURCE_HASH "unknown"
#endif
const char* jn_source_hash(void) { return JN_SOURCE_HASH; }

// ------------------------------------------------------------------------------------------
// K0 tables
// ------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

int patch_bitmaps(const void* bboxes, bool f64, const int32_t* n_boxes, int n, int max_boxes, int patch_size,
                  int grid_rows, int grid_cols, const int32_t* rows, const int32_t* cols, int rule, uint32_t* out,
                  int words_per_item, void* stream) {
  JN_REQUIRE(n >= 0 && max_boxes >= 0 && patch_size >= 1, "jn_patch_bitmaps: bad sizes");
  if (n == 0) return JN_OK;
  JN_REQUIRE(out != nullptr && (bboxes != nullptr || max_boxes == 0), "jn_patch_bitmaps: NULL pointer");
  JN_REQUIRE(rule == JN_RULE_ANY_PIXEL || rule == JN_RULE_AREA5, "jn_patch_bitmaps: unknown rule %d", rule);
  JN_REQUIRE(!f64 || rule == JN_RULE_AREA5, "jn_patch_bitmaps_f64 serves JN_RULE_AREA5 only");
  JN_REQUIRE((rows && cols) || (grid_rows >= 1 && grid_cols >= 1), "jn_patch_bitmaps: grid size missing");
  JN_REQUIRE((rows && cols) ? words_per_item >= 1 : words_per_item >= jn_bitmap_words(grid_rows, grid_cols),
             "jn_patch_bitmaps: words_per_item too small");
  DeviceInfo dev;
  if (int rc = current_device_info(dev)) return rc;
  const int wpb = 4;
  const dim3 grid(grid_for(n, wpb, dev.sm_count * 8)), block(wpb * 32);
  if (f64)
    jnk::patch_bitmaps_kernel<true><<<grid, block, 0, (cudaStream_t)stream>>>(
        bboxes, n_boxes, n, max_boxes, patch_size, grid_rows, grid_cols, rows, cols, rule, out, words_per_item);
  else
    jnk::patch_bitmaps_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(
        bboxes, n_boxes, n, max_boxes, patch_size, grid_rows, grid_cols, rows, cols, rule, out, words_per_item);
  JN_LAUNCHED();
  return JN_OK;
}

int local_boxes(const void* bboxes, bool f64, const int32_t* n_boxes, int max_boxes, int patch_size,
                const int64_t* positions, const int32_t* src_index, int n_items, float* out, void* stream) {
  JN_REQUIRE(n_items >= 0 && max_boxes >= 0 && patch_size >= 1, "jn_local_boxes: bad sizes");
  const long long total = (long long)n_items * max_boxes;
  if (total == 0) return JN_OK;
  JN_REQUIRE(bboxes && positions && out, "jn_local_boxes: NULL pointer");
  DeviceInfo dev;
  if (int rc = current_device_info(dev)) return rc;
  const dim3 grid(grid_for(total, 256, dev.sm_count * 8)), block(256);
  if (f64)
    jnk::local_boxes_kernel<true><<<grid, block, 0, (cudaStream_t)stream>>>(bboxes, n_boxes, max_boxes, patch_size,
                                                                          positions, src_index, n_items, out);
  else
    jnk::local_boxes_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(bboxes, n_boxes, max_boxes, patch_size,
                                                                           positions, src_index, n_items, out);
  JN_LAUNCHED();
  return JN_OK;
}

}  // namespace

extern "C" {

int jn_patch_bitmaps(const int64_t* bboxes, const int32_t* n_boxes, int n, int max_boxes, int patch_size,
                     int grid_rows, int grid_cols, const int32_t* rows, const int32_t* cols, int rule, uint32_t* out,
                     int words_per_item, void* stream) {
  return patch_bitmaps(bboxes, false, n_boxes, n, max_boxes, patch_size, grid_rows, grid_cols, rows, cols, rule, out,
                       words_per_item, stream);
}

int jn_patch_bitmaps_f64(const double* bboxes, const int32_t* n_boxes, int n, int max_boxes, int patch_size,
                         int grid_rows, int grid_cols, const int32_t* rows, const int32_t* cols, int rule,
                         uint32_t* out, int words_per_item, void* stream) {
  return patch_bitmaps(bboxes, true, n_boxes, n, max_boxes, patch_size, grid_rows, grid_cols, rows, cols, rule, out,
                       words_per_item, stream);
}

int jn_bitmap_unpack(const uint32_t* words, int n, int rows, int cols, uint8_t* out, void* stream) {
  JN_REQUIRE(n >= 0 && rows >= 1 && cols >= 1, "jn_bitmap_unpack: bad sizes");
  if (n == 0) return JN_OK;
  JN_REQUIRE(words && out, "jn_bitmap_unpack: NULL pointer");
  DeviceInfo dev;
  if (int rc = current_device_info(dev)) return rc;
  const long long total = (long long)n * rows * cols;
  jnk::bitmap_unpack_kernel<<<grid_for(total, 256, dev.sm_count * 8), 256, 0, (cudaStream_t)stream>>>(
      words, n, rows * cols, jn_bitmap_words(rows, cols), out);
  JN_LAUNCHED();
  return JN_OK;
}

int jn_split_boxes(const int64_t* bboxes, int n, int max_boxes, int patch_size, int rows, int cols, int64_t* local,
                   uint8_t* present, int32_t* status, void* stream) {
  JN_REQUIRE(n >= 0 && max_boxes >= 0 && patch_size >= 1 && rows >= 1 && cols >= 1, "jn_split_boxes: bad sizes");
  const long long total = (long long)n * rows * cols * max_boxes;
  if (total == 0) return JN_OK;
  JN_REQUIRE(bboxes && local && present, "jn_split_boxes: NULL pointer");
  DeviceInfo dev;
  if (int rc = current_device_info(dev)) return rc;
  jnk::split_boxes_kernel<<<grid_for(total, 256, dev.sm_count * 8), 256, 0, (cudaStream_t)stream>>>(
      bboxes, n, max_boxes, patch_size, rows, cols, local, present, status);
  JN_LAUNCHED();
  return JN_OK;
}

int jn_local_boxes(const int64_t* bboxes, const int32_t* n_boxes, int max_boxes, int patch_size,
                   const int64_t* positions, const int32_t* src_index, int n_items, float* out, void* stream) {
  return local_boxes(bboxes, false, n_boxes, max_boxes, patch_size, positions, src_index, n_items, out, stream);
}

int jn_local_boxes_f64(const double* bboxes, const int32_t* n_boxes, int max_boxes, int patch_size,
                       const int64_t* positions, const int32_t* src_index, int n_items, float* out, void* stream) {
  return local_boxes(bboxes, true, n_boxes, max_boxes, patch_size, positions, src_index, n_items, out, stream);
}

// ------------------------------------------------------------------------------------------
// K2 env
// ------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

int check_step_args(const jn_env_step_args* p, bool stepping, const char* who) {
  JN_REQUIRE(p != nullptr, "%s: args is NULL", who);
  JN_REQUIRE(p->n >= 0 && p->rows >= 1 && p->cols >= 1 && p->max_ep_len >= 1, "%s: bad sizes", who);
  JN_REQUIRE((long long)p->rows * p->cols <= 65535, "%s: grids of more than 65535 patches are not supported", who);
  if (p->n == 0) return JN_OK;
  JN_REQUIRE(p->pos_out && p->visited && p->steps && p->has_stopped, "%s: NULL state pointer", who);
  if (stepping)
    JN_REQUIRE(p->pos_in && p->actions && p->bbox && p->rewards && p->terminated && p->truncated,
               "%s: NULL pointer", who);
  if (p->first_slot) {
    JN_REQUIRE(p->host_src && p->history_src, "%s: first_slot needs host_src and history_src", who);
    JN_REQUIRE(p->slots >= 1 && p->t >= 0 && p->t < p->slots, "%s: history slot %d outside [0, %d)", who, p->t, p->slots);
    JN_REQUIRE((long long)p->n * p->slots < (1ll << 31), "%s: history of %d x %d slots is too large", who, p->n, p->slots);
  }
  return JN_OK;
}

jnk::StepArgs to_step_args(const jn_env_step_args* p) {
  jnk::StepArgs a;
  a.pos_in = p->pos_in; a.actions = p->actions; a.pos_out = p->pos_out; a.visited = p->visited; a.bbox = p->bbox;
  a.steps = p->steps; a.has_stopped = p->has_stopped; a.rewards = p->rewards; a.terminated = p->terminated;
  a.truncated = p->truncated; a.first_slot = p->first_slot; a.host_src = p->host_src; a.history_src = p->history_src;
  a.host_tiles = p->host_tiles; a.status = p->status;
  a.n = p->n; a.rows = p->rows; a.cols = p->cols; a.words = jn_bitmap_words(p->rows, p->cols);
  a.max_ep_len = p->max_ep_len; a.stop_enabled = p->stop_enabled; a.slots = p->slots; a.t = p->t; a.cost = p->cost;
  return a;
}

int launch_reset(const jnk::StepArgs& a, cudaStream_t stream) {
  DeviceInfo dev;
  if (int rc = current_device_info(dev)) return rc;
  const long long cells = a.first_slot ? (long long)a.rows * a.cols : 0;
  const long long total = (long long)a.n * (a.words > cells ? a.words : cells);
  JN_CUDA(launch_kernel(jnk::env_reset_kernel, dim3(grid_for(total, 128, dev.sm_count * 16)), dim3(128), 0, stream,
                        false, a));
  return JN_OK;
}

int launch_step(const jnk::StepArgs& a, cudaStream_t stream, bool pdl = false) {
  DeviceInfo dev;
  if (int rc = current_device_info(dev)) return rc;
  // G lanes per episode in the bitmap phase: the power of two >= the number of bitmap words, at most a warp
  int g = 1;
  while (g < a.words && g < 32) g <<= 1;
  const dim3 grid(grid_for(a.n, 64, dev.sm_count * 16)), block(64);  // a warp steps 32 episodes
#define JN_STEP(G) \
  case G: JN_CUDA(launch_kernel(jnk::env_step_kernel<G>, grid, block, 0, stream, pdl, a)); break;
  if (a.words > 32) {  // its own instantiation: the second block of words costs 120 registers
    JN_CUDA(launch_kernel(jnk::env_step_kernel<32, true>, grid, block, 0, stream, pdl, a));
    return JN_OK;
  }
  switch (g) { JN_STEP(1) JN_STEP(2) JN_STEP(4) JN_STEP(8) JN_STEP(16) JN_STEP(32) }
#undef JN_STEP
  return JN_OK;
}

// The gather that follows a reset / step inside the same call.
int gather_after(const jn_images* set, const jn_images* history_set, const jn_env_step_args* p, bool stepping,
                 cudaStream_t stream) {
  if (!set || !p->out) return JN_OK;  // state only
  const bool zero_copy = p->first_slot != nullptr;
  GatherRequest rq;
  rq.shifts = p->shifts; rq.n_items = p->n; rq.out = p->out; rq.out_item_stride_bytes = p->out_item_stride_bytes;
  rq.flags = p->flags; rq.engine = p->engine; rq.status = p->status;
  rq.grid_rows = p->rows; rq.grid_cols = p->cols;
  rq.pdl = true;
  if (zero_copy) {
    // sources come from the kernel before us: positions after the move, first visits only
    rq.positions = p->pos_out; rq.src_index = p->host_src; rq.wait_prior = true;
  } else if (stepping) {
    // independent of the step kernel: the gather moves the old positions itself and runs next to it
    rq.positions = p->pos_in; rq.actions = p->actions;
  } else {
    rq.positions = p->pos_out;  // reset: the start positions (the reset kernel does not write them)
  }
  if (int rc = gather_launch(set, rq, stream)) return rc;
  if (zero_copy && history_set) {
    // revisited patches: copied from the history slot that first held them (one-patch images, patch (0, 0))
    GatherRequest rh;
    rh.src_index = p->history_src; rh.n_items = p->n; rh.out = p->out;
    rh.out_item_stride_bytes = p->out_item_stride_bytes; rh.flags = 0; rh.engine = p->engine; rh.status = p->status;
    // behind the host gather with the programmatic attribute and WITHOUT a wait: that gather triggers its
    // dependents right after it has itself waited for the step kernel, so this HBM -> HBM copy overlaps the
    // PCIe-bound reads; the two write disjoint tiles of the slot
    rh.pdl = true;
    if (int rc = gather_launch(history_set, rh, stream)) return rc;
  }
  return JN_OK;
}

}  // namespace

extern "C" {

int jn_env_reset(const int64_t* positions, uint32_t* visited, int64_t* steps, uint8_t* has_stopped, int n, int rows,
                 int cols, int32_t* status, void* stream) {
  jn_env_step_args p;
  memset(&p, 0, sizeof(p));
  p.pos_out = const_cast<int64_t*>(positions); p.visited = visited; p.steps = steps; p.has_stopped = has_stopped;
  p.n = n; p.rows = rows; p.cols = cols; p.max_ep_len = 1; p.status = status;
  if (int rc = check_step_args(&p, false, "jn_env_reset")) return rc;
  if (n == 0) return JN_OK;
  return launch_reset(to_step_args(&p), (cudaStream_t)stream);
}

int jn_env_step(const int64_t* pos_in, const int64_t* actions, int64_t* pos_out, uint32_t* visited,
                const uint32_t* bbox, int64_t* steps, uint8_t* has_stopped, float* rewards, uint8_t* terminated,
                uint8_t* truncated, int n, int rows, int cols, int max_ep_len, float cost, int stop_enabled,
                int32_t* status, void* stream) {
  jn_env_step_args p;
  memset(&p, 0, sizeof(p));
  p.pos_in = pos_in; p.actions = actions; p.pos_out = pos_out; p.visited = visited; p.bbox = bbox; p.steps = steps;
  p.has_stopped = has_stopped; p.rewards = rewards; p.terminated = terminated; p.truncated = truncated;
  p.n = n; p.rows = rows; p.cols = cols; p.max_ep_len = max_ep_len; p.cost = cost; p.stop_enabled = stop_enabled;
  p.status = status;
  if (int rc = check_step_args(&p, true, "jn_env_step")) return rc;
  if (n == 0) return JN_OK;
  return launch_step(to_step_args(&p), (cudaStream_t)stream);
}

int jn_env_reset_gather(const jn_images* set, const jn_images* history_set, const jn_env_step_args* args,
                        void* stream) {
  if (int rc = check_step_args(args, false, "jn_env_reset_gather")) return rc;
  if (args->n == 0) return JN_OK;
  if (int rc = launch_reset(to_step_args(args), (cudaStream_t)stream)) return rc;
  return gather_after(set, history_set, args, false, (cudaStream_t)stream);
}

int jn_env_step_gather(const jn_images* set, const jn_images* history_set, const jn_env_step_args* args,
                       void* stream) {
  if (int rc = check_step_args(args, true, "jn_env_step_gather")) return rc;
  if (args->n == 0) return JN_OK;
  JN_REQUIRE(args->pos_in != args->pos_out || !set || !args->out,
             "jn_env_step_gather: pos_in and pos_out must not alias (the gather reads pos_in while the step runs)");
  if (set && args->out && !args->first_slot) {
    // The gather does not need the step kernel (it moves the old positions itself) and is the long one of the
    // two: it goes FIRST, and the step kernel is launched behind it with the programmatic attribute -- the gather
    // releases its dependents in its prologue -- so the state update runs next to the gather instead of in
    // front of it.  The next step's launches are ordinary ones and wait for both.
    GatherRequest rq;
    rq.positions = args->pos_in; rq.actions = args->actions; rq.shifts = args->shifts; rq.n_items = args->n;
    rq.out = args->out; rq.out_item_stride_bytes = args->out_item_stride_bytes; rq.flags = args->flags;
    rq.engine = args->engine; rq.status = args->status; rq.grid_rows = args->rows; rq.grid_cols = args->cols;
    if (int rc = gather_launch(set, rq, (cudaStream_t)stream)) return rc;
    return launch_step(to_step_args(args), (cudaStream_t)stream, true);
  }
  if (int rc = launch_step(to_step_args(args), (cudaStream_t)stream)) return rc;
  return gather_after(set, history_set, args, true, (cudaStream_t)stream);
}

int jn_env_rewards(const int64_t* positions, const uint32_t* visited, const uint32_t* bbox, const uint8_t* has_stopped,
                   int n, int rows, int cols, float cost, int stop_enabled, float* rewards, void* stream) {
  JN_REQUIRE(n >= 0 && rows >= 1 && cols >= 1, "jn_env_rewards: bad sizes");
  if (n == 0) return JN_OK;
  JN_REQUIRE(positions && visited && bbox && has_stopped && rewards, "jn_env_rewards: NULL pointer");
  DeviceInfo dev;
  if (int rc = current_device_info(dev)) return rc;
  jnk::env_rewards_kernel<<<grid_for(n, 128, dev.sm_count * 16), 128, 0, (cudaStream_t)stream>>>(
      positions, visited, bbox, has_stopped, n, rows, cols, jn_bitmap_words(rows, cols), cost, stop_enabled, rewards);
  JN_LAUNCHED();
  return JN_OK;
}

int jn_env_props(const uint32_t* visited, const uint32_t* bbox, const uint8_t* has_stopped, int n, int rows, int cols,
                 int stop_enabled, float* prop_patches, uint8_t* terminated, void* stream) {
  JN_REQUIRE(n >= 0 && rows >= 1 && cols >= 1, "jn_env_props: bad sizes");
  if (n == 0) return JN_OK;
  JN_REQUIRE(visited && bbox && (has_stopped || !stop_enabled || !terminated), "jn_env_props: NULL pointer");
  DeviceInfo dev;
  if (int rc = current_device_info(dev)) return rc;
  const int wpb = 4;
  jnk::env_props_kernel<<<grid_for(n, wpb, dev.sm_count * 16), wpb * 32, 0, (cudaStream_t)stream>>>(
      visited, bbox, has_stopped, n, jn_bitmap_words(rows, cols), stop_enabled, prop_patches, terminated);
  JN_LAUNCHED();
  return JN_OK;
}

// ------------------------------------------------------------------------------------------
// glimpse pyramid
// ------------------------------------------------------------------------------------------
int jn_resize_aa_reflect(const void* src, int64_t src_image_stride_bytes, float* tmp, void* dst,
                         int64_t dst_image_stride_bytes, int dtype, int n_images, int channels, int height, int width,
                         int pad, const int32_t* first_x, const int32_t* count_x, const float* weights_x, int k_x,
                         const int32_t* first_y, const int32_t* count_y, const float* weights_y, int k_y,
                         void* stream) {
  JN_REQUIRE(dtype == JN_U8 || dtype == JN_F32, "jn_resize_aa_reflect: dtype must be JN_U8 or JN_F32");
  JN_REQUIRE(n_images >= 0 && channels >= 1 && height >= 2 && width >= 2, "jn_resize_aa_reflect: bad sizes");
  JN_REQUIRE(pad >= 0 && pad < height && pad < width, "jn_resize_aa_reflect: reflect padding needs pad < image size");
  if (n_images == 0) return JN_OK;
  JN_REQUIRE(src && tmp && dst && first_x && count_x && weights_x && first_y && count_y && weights_y && k_x >= 1 &&
                 k_y >= 1,
             "jn_resize_aa_reflect: NULL pointer");
  const int elem = dtype == JN_F32 ? 4 : 1;
  JN_REQUIRE(src_image_stride_bytes % elem == 0 && dst_image_stride_bytes % elem == 0, "jn_resize_aa_reflect: odd stride");
  DeviceInfo dev;
  if (int rc = current_device_info(dev)) return rc;
  const long long total = (long long)n_images * channels * height * width;
  const dim3 grid(grid_for(total, 256 * 4, dev.sm_count * 16)), block(256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == JN_F32)
    jnk::resize_aa_rows_kernel<float><<<grid, block, 0, st>>>(static_cast<const float*>(src), src_image_stride_bytes / 4,
                                                             tmp, n_images, channels, height, width, pad, first_x,
                                                             count_x, weights_x, k_x);
  else
    jnk::resize_aa_rows_kernel<uint8_t><<<grid, block, 0, st>>>(static_cast<const uint8_t*>(src), src_image_stride_bytes,
                                                               tmp, n_images, channels, height, width, pad, first_x,
                                                               count_x, weights_x, k_x);
  JN_LAUNCHED();
  if (dtype == JN_F32)
    jnk::resize_aa_cols_kernel<float><<<grid, block, 0, st>>>(tmp, static_cast<float*>(dst), dst_image_stride_bytes / 4,
                                                             n_images, channels, height, width, pad, first_y, count_y,
                                                             weights_y, k_y);
  else
    jnk::resize_aa_cols_kernel<uint8_t><<<grid, block, 0, st>>>(tmp, static_cast<uint8_t*>(dst), dst_image_stride_bytes,
                                                               n_images, channels, height, width, pad, first_y, count_y,
                                                               weights_y, k_y);
  JN_LAUNCHED();
  return JN_OK;
}

// ------------------------------------------------------------------------------------------
// K3 scans
// ------------------------------------------------------------------------------------------
int jn_returns(const float* rewards_tn, const uint8_t* terminated_tn, int T, int n, float* rewards_out,
               uint8_t* masks, uint8_t* logit_masks, float* returns, void* stream) {
  JN_REQUIRE(T >= 0 && n >= 0, "jn_returns: bad sizes");
  if (n == 0) return JN_OK;
  JN_REQUIRE(masks && (T == 0 || (rewards_tn && terminated_tn && rewards_out && logit_masks && returns)),
             "jn_returns: NULL pointer");
  jnk::returns_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rewards_tn, terminated_tn, T, n, rewards_out,
                                                                         masks, logit_masks, returns);
  JN_LAUNCHED();
  return JN_OK;
}

int jn_returns_rows(const float* rewards, int64_t rewards_row_stride, const uint8_t* logit_masks,
                    int64_t masks_row_stride, int T, int n, float* returns, void* stream) {
  JN_REQUIRE(T >= 0 && n >= 0, "jn_returns_rows: bad sizes");
  if (n == 0 || T == 0) return JN_OK;
  JN_REQUIRE(rewards && logit_masks && returns, "jn_returns_rows: NULL pointer");
  jnk::returns_rows_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rewards, rewards_row_stride, logit_masks,
                                                                              masks_row_stride, T, n, returns);
  JN_LAUNCHED();
  return JN_OK;
}

int jn_tile_lookup(const int64_t* traj_positions, const int32_t* traj_src, int T, const int64_t* query_positions,
                   const int32_t* query_src, int n_queries, int slab_base, int64_t* out_positions, int32_t* out_src,
                   void* stream) {
  JN_REQUIRE(T >= 1 && n_queries >= 0 && slab_base >= 0, "jn_tile_lookup: bad sizes");
  if (n_queries == 0) return JN_OK;
  JN_REQUIRE(traj_positions && traj_src && query_positions && query_src && out_positions && out_src,
             "jn_tile_lookup: NULL pointer");
  DeviceInfo dev;
  if (int rc = current_device_info(dev)) return rc;
  jnk::tile_lookup_kernel<<<grid_for(n_queries, 128, dev.sm_count * 8), 128, 0, (cudaStream_t)stream>>>(
      traj_positions, traj_src, T, query_positions, query_src, n_queries, slab_base, out_positions, out_src);
  JN_LAUNCHED();
  return JN_OK;
}

int jn_visit_sources(const int64_t* positions, int32_t* first_slot, int n, int rows, int cols, int slots, int t,
                     int32_t* host_src, int32_t* history_src, int32_t* status, void* stream) {
  JN_REQUIRE(n >= 0 && rows >= 1 && cols >= 1 && slots >= 1 && t >= 0 && t < slots, "jn_visit_sources: bad sizes");
  JN_REQUIRE((long long)n * slots < (1ll << 31), "jn_visit_sources: history of %d x %d slots is too large", n, slots);
  if (n == 0) return JN_OK;
  JN_REQUIRE(positions && first_slot && host_src && history_src, "jn_visit_sources: NULL pointer");
  DeviceInfo dev;
  if (int rc = current_device_info(dev)) return rc;
  jnk::visit_sources_kernel<<<grid_for(n, 128, dev.sm_count * 8), 128, 0, (cudaStream_t)stream>>>(
      positions, first_slot, n, rows, cols, slots, t, host_src, history_src, status);
  JN_LAUNCHED();
  return JN_OK;
}

int jn_tile_dedupe(const int64_t* traj_positions, const int32_t* traj_src, int n_slots, int T, int32_t* first_src,
                   int32_t* repeat_src, void* stream) {
  JN_REQUIRE(T >= 1 && n_slots >= 0 && n_slots % T == 0, "jn_tile_dedupe: n_slots must be a multiple of T");
  if (n_slots == 0) return JN_OK;
  JN_REQUIRE(traj_positions && traj_src && first_src && repeat_src, "jn_tile_dedupe: NULL pointer");
  DeviceInfo dev;
  if (int rc = current_device_info(dev)) return rc;
  jnk::tile_dedupe_kernel<<<grid_for(n_slots, 128, dev.sm_count * 8), 128, 0, (cudaStream_t)stream>>>(
      traj_positions, traj_src, n_slots, T, first_src, repeat_src);
  JN_LAUNCHED();
  return JN_OK;
}

int jn_traj_expand(const int32_t* start_yx, const int32_t* seg_begin, const int32_t* seg_to_yx,
                   const int32_t* seg_tgt_yx, const uint8_t* seg_flags, const int32_t* draw_begin,
                   const uint8_t* draws, const uint32_t* area_bitmaps, int words_per_item, const int32_t* cols, int n,
                   int T, int64_t* positions, int64_t* current_actions, int64_t* next_actions, int64_t* labels,
                   float* masks, int32_t* gather_src, int32_t* ep_len, int32_t* status, void* stream) {
  JN_REQUIRE(n >= 0 && T >= 1 && words_per_item >= 1, "jn_traj_expand: bad sizes");
  if (n == 0) return JN_OK;
  JN_REQUIRE(start_yx && seg_begin && draw_begin && area_bitmaps && cols && positions && current_actions &&
                 next_actions && labels && masks && gather_src && ep_len,
             "jn_traj_expand: NULL pointer");
  DeviceInfo dev;
  if (int rc = current_device_info(dev)) return rc;
  jnk::traj_expand_kernel<<<grid_for(n, jnk::kTrajWarps, dev.sm_count * 8), jnk::kTrajWarps * 32, 0,
                           (cudaStream_t)stream>>>(start_yx, seg_begin, seg_to_yx, seg_tgt_yx, seg_flags, draw_begin,
                                                   draws, area_bitmaps, words_per_item, cols, n, T, positions,
                                                   current_actions, next_actions, labels, masks, gather_src, ep_len,
                                                   status);
  JN_LAUNCHED();
  return JN_OK;
}

}  // extern "C"
