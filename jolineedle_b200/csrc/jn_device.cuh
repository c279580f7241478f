// Device-side helpers shared by the kernels: PTX wrappers for mbarrier / TMA (sm_100a),
// the action table and small integer utilities.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace jnk {

// (dy, dx) per action code -- contract with jolineedle_b200/env/common.py (reference:
// src/env/common.py:17-27).  Packed as two 4-bit-per-entry tables would be cute; two small
// constant arrays are clearer and live in the constant bank.
__device__ __constant__ const int8_t kActionDy[9] = {0, 0, -1, 1, -1, -1, 1, 1, 0};
__device__ __constant__ const int8_t kActionDx[9] = {-1, 1, 0, 0, -1, 1, -1, 1, 0};
constexpr int kStop = 8;

// Greedy 8-neighbour direction of the gradient (dy, dx); same decision table as the
// reference's move_towards (simple_env.py:84-125).  index = (sign(dy)+1)*3 + sign(dx)+1.
__host__ __device__ __forceinline__ int direction_code(int dy, int dx) {
  const int sy = (dy > 0) - (dy < 0), sx = (dx > 0) - (dx < 0);
  // nibble i (from the low end) = code for index i:
  //   0 LEFT_UP(4) 1 UP(2) 2 RIGHT_UP(5) | 3 LEFT(0) 4 STOP(8) 5 RIGHT(1) | 6 LEFT_DOWN(6) 7 DOWN(3) 8 RIGHT_DOWN(7)
  const unsigned long long lut = 0x736180524ull;
  return (int)((lut >> (4 * ((sy + 1) * 3 + (sx + 1)))) & 0xF);
}

__device__ __forceinline__ int iabs(int v) { return v < 0 ? -v : v; }
__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }
__device__ __forceinline__ long long lmin(long long a, long long b) { return a < b ? a : b; }
__device__ __forceinline__ long long lmax(long long a, long long b) { return a > b ? a : b; }
// floor division / modulo for possibly negative numerators (python semantics, b > 0)
__device__ __forceinline__ long long floordiv(long long a, long long b) {
  long long q = a / b;
  return (a % b != 0 && (a < 0)) ? q - 1 : q;
}

// uint8 -> float32 value / 255, correctly rounded (== torch's `x.float() / 255`).
// q0 = x * fl(1/255); one Newton correction with exact residual via FMA.  Verified
// exhaustively for all 256 inputs on the host (tests/test_oracle_cpu.py) and on the device
// (tests/test_gather_gpu.py).
__host__ __device__ __forceinline__ float u8_to_unit(float x) {
  const float r = 1.0f / 255.0f;  // constant-folded, correctly rounded fl32(1/255)
  const float q0 = x * r;
#ifdef __CUDA_ARCH__
  const float rem = __fmaf_rn(-255.0f, q0, x);
  return __fmaf_rn(rem, r, q0);
#else
  const float rem = fmaf(-255.0f, q0, x);
  return fmaf(rem, r, q0);
#endif
}

// The same value from x/256 (exact in float32): x/255 = x/256 + (x/256)/255 is ONE rounding in
// fma(xs, fl(1/255), xs) -- the error of fl(1/255) is scaled by 1/255 and stays far inside half an ulp
// (all 256 inputs are compared with u8_to_unit in jn_selftest_host and with torch on the device).
__host__ __device__ __forceinline__ float unit_from_scaled(float xs) {
  const float r = 1.0f / 255.0f;
#ifdef __CUDA_ARCH__
  return __fmaf_rn(xs, r, xs);
#else
  return fmaf(xs, r, xs);
#endif
}
#ifdef __CUDACC__
// Byte k of a packed word -> value / 255 in three full-rate instructions: PRMT builds the float
// 32768 + b/256 (one ulp is 2^-8 there), FADD strips the 32768 exactly, FFMA as above.
template <int k>
__device__ __forceinline__ float byte_to_unit(uint32_t u) {
  return unit_from_scaled(__uint_as_float(__byte_perm(u, 0x47000000u, 0x7440 + k)) - 32768.0f);
}
#endif

// ---- shared-memory address / mbarrier ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA: 1-D bulk copies -------------------------------------------------------------------
// global -> shared, completion signalled on an mbarrier (complete_tx).  16-byte aligned
// addresses and sizes.  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
template <int kPending>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(kPending) : "memory");
}

// ---- TMA: tensor-map tiles -------------------------------------------------------------------
// 4-D tiled load: coordinates are (innermost .. outermost) element offsets.  SASS: UTMALDG.
__device__ __forceinline__ void tensor_g2s_4d(void* dst_smem, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                              uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, "
      "%5}], [%6];" ::"r"(smem_u32(dst_smem)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tensor_g2s_3d(void* dst_smem, const CUtensorMap* map, int c0, int c1, int c2,
                                              uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], "
      "[%5];" ::"r"(smem_u32(dst_smem)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---- programmatic dependent launch (sm_90+) ---------------------------------------------------
// A kernel that executes pdl_launch_dependents() lets the next kernel of the stream -- when that one was
// launched with the programmatic-serialization attribute -- start before this one has finished; pdl_wait() in
// the dependent kernel blocks until the prerequisite grid has completed and its writes are visible (a no-op
// for a normally launched kernel).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// streaming stores (the crops are consumed by the next kernel, never re-read by us)
__device__ __forceinline__ void st_f4(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_f2(float* p, float a, float b) {
  asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
// streaming variants (evict-first in L2): the crops of a uint8 -> float32 gather are written once and are four
// times the bytes read; used for launches whose crops are of the order of the L2 (jn_api.cu: stream_stores)
__device__ __forceinline__ void st_f4_cs(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_f2_cs(float* p, float a, float b) {
  asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

}  // namespace jnk
