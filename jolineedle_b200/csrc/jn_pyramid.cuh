// Glimpse pyramid: one level = reflect-pad by `pad` pixels + antialiased bilinear resize back to H x W
// (general_env.py:84-115: torchvision TF.pad(..., "reflect") + TF.resize(..., antialias=True) on the CPU).
//
// ATen's CPU kernel is separable (rows, then columns) and computes every output pixel as the chain
//     t = s[0] * w[0];  t = fma(s[j], w[j], t)   j = 1 .. count-1
// over the float32 weights of that output index (its AVX2 / AVX-512 builds contract the multiply-add).  The two
// kernels below run the same chains with __fmul_rn / __fmaf_rn in the same order, reading the weight tables the
// host computed operation by operation like ATen (jolineedle_b200/pyramid.py:aa_weights), so a level equals the
// reference's bit for bit.  The reflect padding is never materialised: padded index i maps to source index
// reflect(i - pad).  Construction-time work (once per env), one thread per output pixel.
//
// uint8 images: torchvision casts them to float32, resizes, and casts back through torch.round (half to even) --
// so the rows pass reads bytes as floats and the columns pass ends in a round-to-nearest-even conversion.
#pragma once

#include "jn_device.cuh"

namespace jnk {

__device__ __forceinline__ int reflect_index(int i, int n) {  // torch 'reflect': no repeated edge
  i = i < 0 ? -i : i;
  return i >= n ? 2 * (n - 1) - i : i;
}

__device__ __forceinline__ float pixel_as_float(float v) { return v; }
__device__ __forceinline__ float pixel_as_float(uint8_t v) { return (float)v; }
__device__ __forceinline__ void store_pixel(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_pixel(uint8_t* p, float v) {  // torch.round, then the cast
  *p = (uint8_t)imin(imax(__float2int_rn(v), 0), 255);
}

// tmp[b, c, y, xo] from src[b, c, y, :]  (src images `src_image_stride` ELEMENTS apart, [C, H, W] inside)
template <typename T>
__global__ void resize_aa_rows_kernel(const T* __restrict__ src, long long src_image_stride, float* __restrict__ tmp,
                                      int B, int C, int H, int W, int pad, const int32_t* __restrict__ first,
                                      const int32_t* __restrict__ count, const float* __restrict__ wts, int k) {
  const long long total = (long long)B * C * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xo = (int)(i % W);
    long long r = i / W;
    const int y = (int)(r % H); r /= H;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    const T* row = src + b * src_image_stride + ((long long)c * H + y) * W;
    const float* w = wts + (long long)xo * k;
    const int x0 = first[xo] - pad, n = count[xo];
    float t = __fmul_rn(pixel_as_float(row[reflect_index(x0, W)]), w[0]);
    for (int j = 1; j < n; ++j) t = __fmaf_rn(pixel_as_float(row[reflect_index(x0 + j, W)]), w[j], t);
    tmp[i] = t;
  }
}

// dst[b, c, yo, x] from tmp[b, c, :, x]  (tmp contiguous [B, C, H, W]; dst images `dst_image_stride` ELEMENTS apart)
template <typename T>
__global__ void resize_aa_cols_kernel(const float* __restrict__ tmp, T* __restrict__ dst, long long dst_image_stride,
                                      int B, int C, int H, int W, int pad, const int32_t* __restrict__ first,
                                      const int32_t* __restrict__ count, const float* __restrict__ wts, int k) {
  const long long total = (long long)B * C * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    long long r = i / W;
    const int yo = (int)(r % H); r /= H;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    const float* plane = tmp + ((long long)b * C + c) * H * W + x;
    const float* w = wts + (long long)yo * k;
    const int y0 = first[yo] - pad, n = count[yo];
    float t = __fmul_rn(plane[(long long)reflect_index(y0, H) * W], w[0]);
    for (int j = 1; j < n; ++j) t = __fmaf_rn(plane[(long long)reflect_index(y0 + j, H) * W], w[j], t);
    store_pixel(dst + b * dst_image_stride + ((long long)c * H + yo) * W + x, t);
  }
}

}  // namespace jnk
