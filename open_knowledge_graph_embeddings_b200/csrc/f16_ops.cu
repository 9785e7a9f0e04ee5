// FP16 operands of the tensor-core contractions (gemm_tc.cu).
//
// The reference scores in fp32 (torch.mm, openkge/model.py:206-215, 270-272). The tensor cores take 16-bit or TF32
// inputs; we feed them IEEE fp16: the same 10-bit mantissa as TF32 at twice the rate and half the bytes, rounded to
// nearest (tcgen05's own fp32 -> tf32 conversion truncates). What fp16 lacks is exponent range, so every operand
// carries ONE power-of-two scale: x16 = fp16(x * scale), scale = 2^(8 - e) with |x|max in [2^(e-1), 2^e), i.e. the
// largest element lands in [128, 256). Elements down to 2^-22 of the largest keep all 11 significant bits, smaller ones
// round with an absolute error below 2^-33 of the largest -- invisible in a norm-wise bound. The scale is exact to
// undo (the epilogues multiply the fp32 accumulator by the inverse scales), and conversions saturate instead of
// producing infinities.
//
// Split precision (evaluation): lo = fp16(x * scale - hi) holds the next 11 bits. The three-term product
// q_hi e_hi + q_hi e_lo + q_lo e_hi (one contraction with three passes over K) reproduces the fp32 product to ~2^-21
// relative, norm-wise, so the filtered ranks are those of an fp32 scorer up to genuine near-ties.
#include "okge_common.cuh"

#include <cuda_fp16.h>
#include <math.h>

namespace okge {

namespace {

constexpr int kParts = OKGE_F16_ABSMAX_PARTS;
constexpr int kThreads = 256;

__device__ __forceinline__ float block_max(float v) {
  __shared__ float red[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < kThreads / 32; ++i) r = fmaxf(r, red[i]);
  __syncthreads();
  return r;
}

// partials[b] = max |x| over the rows block b strides through (0 for blocks without rows); NaNs are ignored
__global__ void __launch_bounds__(kThreads)
absmax_kernel(const float* __restrict__ x, int64_t ld, int64_t rows, int cols, float* __restrict__ partials) {
  pdl_wait_and_trigger();
  float m = 0.f;
  const bool vec = (cols % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0);
  if (vec) {
    const int c4 = cols / 4;
    const int64_t total = rows * c4;
    const bool flat = ld == cols;                 // contiguous rows: no division per element
    for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
      int64_t off = 4 * i;
      if (!flat) {
        const int64_t r = i / c4;
        off = r * ld + 4 * (i - r * c4);
      }
      const float4 v = ldg_nc_f4(reinterpret_cast<const float4*>(x + off));
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
  } else {
    const int64_t total = rows * cols;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
      const int64_t r = i / cols;
      m = fmaxf(m, fabsf(__ldg(x + r * ld + (i - r * cols))));
    }
  }
  m = block_max(m);
  if (threadIdx.x == 0) partials[blockIdx.x] = m;
}

__device__ __forceinline__ float scale_for(float amax) {
  if (!(amax > 0.f) || isinf(amax)) return 1.f;
  int e;
  frexpf(amax, &e);                     // amax = m 2^e, m in [0.5, 1)
  e = max(-118, min(128, e));            // keeps scale and 1 / scale normal fp32 numbers
  return exp2f(static_cast<float>(8 - e));
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float2 unpack2(uint32_t h) {
  return __half22float2(*reinterpret_cast<const __half2*>(&h));
}

// hi[r, c] = fp16(x[r, c] * scale), lo[r, c] = fp16(x * scale - hi) (optional). scale: from the absmax partials
// (dynamic) or `fixed_scale`; block 0 publishes 1 / scale.
__global__ void __launch_bounds__(kThreads)
quantize_kernel(const float* __restrict__ x, int64_t ld, int64_t rows, int cols, const float* __restrict__ partials,
                float fixed_scale, __half* __restrict__ hi, __half* __restrict__ lo, int64_t ld16,
                float* __restrict__ inv_scale_out) {
  pdl_wait_and_trigger();
  float scale = fixed_scale;
  if (partials != nullptr) scale = scale_for(block_max(threadIdx.x < kParts ? __ldg(partials + threadIdx.x) : 0.f));
  if (blockIdx.x == 0 && threadIdx.x == 0 && inv_scale_out != nullptr) inv_scale_out[0] = 1.0f / scale;
  const bool vec = (cols % 8 == 0) && (ld % 4 == 0) && (ld16 % 8 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(hi) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(lo) & 15u) == 0);
  if (vec) {
    const int c8 = cols / 8;
    const int64_t total = rows * c8;
    const bool flat = ld == cols && ld16 == cols;      // contiguous rows: no division per element
    for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
      int64_t off = 8 * i, off16 = 8 * i;
      if (!flat) {
        const int64_t r = i / c8;
        const int64_t c = 8 * (i - r * c8);
        off = r * ld + c;
        off16 = r * ld16 + c;
      }
      const float4* src = reinterpret_cast<const float4*>(x + off);
      const float4 a = ldg_nc_f4(src), b = ldg_nc_f4(src + 1);
      const float v[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, b.x * scale, b.y * scale, b.z * scale, b.w * scale};
      uint4 h;
      h.x = pack2(v[0], v[1]); h.y = pack2(v[2], v[3]); h.z = pack2(v[4], v[5]); h.w = pack2(v[6], v[7]);
      *reinterpret_cast<uint4*>(hi + off16) = h;
      if (lo != nullptr) {
        const float2 h0 = unpack2(h.x), h1 = unpack2(h.y), h2 = unpack2(h.z), h3 = unpack2(h.w);
        uint4 l;
        l.x = pack2(v[0] - h0.x, v[1] - h0.y); l.y = pack2(v[2] - h1.x, v[3] - h1.y);
        l.z = pack2(v[4] - h2.x, v[5] - h2.y); l.w = pack2(v[6] - h3.x, v[7] - h3.y);
        *reinterpret_cast<uint4*>(lo + off16) = l;
      }
    }
  } else {
    const int64_t total = rows * cols;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
      const int64_t r = i / cols;
      const int64_t c = i - r * cols;
      const float v = __ldg(x + r * ld + c) * scale;
      const uint32_t h = pack2(v, 0.f);
      hi[r * ld16 + c] = __ushort_as_half(static_cast<unsigned short>(h & 0xFFFFu));
      if (lo != nullptr) {
        const uint32_t l = pack2(v - unpack2(h).x, 0.f);
        lo[r * ld16 + c] = __ushort_as_half(static_cast<unsigned short>(l & 0xFFFFu));
      }
    }
  }
}

int stream_grid(int64_t n_items) {
  int64_t blocks = ceil_div64(n_items, kThreads);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace

}  // namespace okge

using namespace okge;

// dst = keep ? src : 0 for the inverted-dropout mask dropout_kernel (embed_ops.cu) draws for (p, seed, offset [, step]) over
// the flattened [rows, cols] matrix; the 1 / (1 - p) of the kept elements goes into the operand's inverse scale.
__global__ void __launch_bounds__(kThreads)
mask_dropout_kernel(const __half* __restrict__ src, int64_t ld_src, int64_t rows, int cols4, float p, uint64_t seed,
                    uint64_t offset, const unsigned long long* __restrict__ step_dev, const float* __restrict__ src_inv,
                    __half* __restrict__ dst, int64_t ld_dst, float* __restrict__ dst_inv) {
  pdl_wait_and_trigger();
  if (step_dev != nullptr) offset += static_cast<uint64_t>(*step_dev) << 44;
  if (blockIdx.x == 0 && threadIdx.x == 0 && dst_inv != nullptr)
    dst_inv[0] = (src_inv != nullptr ? __ldg(src_inv) : 1.0f) / (1.0f - p);
  const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const int64_t total = rows * cols4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * kThreads) {
    const int64_t r = i / cols4;
    const int c = static_cast<int>(i - r * cols4) * 4;
    const uint64_t ctr = static_cast<uint64_t>(i) + offset;         // 4 consecutive elements of the flattened matrix
    const uint4 rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), 0u, 0u), key);
    uint2 v = *reinterpret_cast<const uint2*>(src + r * ld_src + c);   // 4 halves
    if (!((rnd.x >> 8) * (1.0f / 16777216.0f) >= p)) v.x &= 0xffff0000u;
    if (!((rnd.y >> 8) * (1.0f / 16777216.0f) >= p)) v.x &= 0x0000ffffu;
    if (!((rnd.z >> 8) * (1.0f / 16777216.0f) >= p)) v.y &= 0xffff0000u;
    if (!((rnd.w >> 8) * (1.0f / 16777216.0f) >= p)) v.y &= 0x0000ffffu;
    *reinterpret_cast<uint2*>(dst + r * ld_dst + c) = v;
  }
}

extern "C" int okge_f16_absmax(const float* x, int64_t ld, int64_t rows, int64_t cols, float* partials,
                               okge_stream_t stream) {
  OKGE_REQUIRE(x != nullptr && partials != nullptr, "null pointer");
  OKGE_REQUIRE(rows >= 0 && cols > 0 && ld >= cols && cols < (1 << 30), "bad shape");
  OKGE_LAUNCH((absmax_kernel), kParts, kThreads, 0, static_cast<cudaStream_t>(stream), x, ld, rows, static_cast<int>(cols), partials);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_f16_quantize(const float* x, int64_t ld, int64_t rows, int64_t cols, const float* partials,
                                 float fixed_scale, okge_half_t* hi, okge_half_t* lo, int64_t ld16, float* inv_scale,
                                 okge_stream_t stream) {
  OKGE_REQUIRE(x != nullptr && hi != nullptr, "null pointer");
  OKGE_REQUIRE(rows >= 0 && cols > 0 && ld >= cols && ld16 >= cols && cols < (1 << 30), "bad shape");
  OKGE_REQUIRE(partials != nullptr || fixed_scale > 0.f, "a fixed scale must be positive");
  if (rows == 0 && inv_scale == nullptr) return OKGE_OK;
  OKGE_LAUNCH((quantize_kernel), stream_grid(rows * ((cols + 7) / 8)), kThreads, 0, static_cast<cudaStream_t>(stream), x, ld, rows, static_cast<int>(cols), partials, fixed_scale, reinterpret_cast<__half*>(hi),
      reinterpret_cast<__half*>(lo), ld16, inv_scale);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_f16_mask_dropout(const okge_half_t* src, int64_t ld_src, int64_t rows, int64_t cols, float p, uint64_t seed,
                                     uint64_t offset, const uint64_t* step_dev, const float* src_inv_scale, okge_half_t* dst,
                                     int64_t ld_dst, float* dst_inv_scale, okge_stream_t stream) {
  OKGE_REQUIRE(src != nullptr && dst != nullptr, "null pointer");
  OKGE_REQUIRE(rows >= 0 && cols > 0 && cols % 4 == 0 && ld_src >= cols && ld_dst >= cols && ld_src % 4 == 0 && ld_dst % 4 == 0,
               "cols and row pitches must be multiples of 4");
  OKGE_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 7u) == 0, "operands must be 8-byte aligned");
  OKGE_REQUIRE(p > 0.f && p < 1.f, "dropout probability must be in (0, 1)");
  if (rows == 0 && dst_inv_scale == nullptr) return OKGE_OK;
  OKGE_LAUNCH((mask_dropout_kernel), stream_grid(rows * (cols / 4)), kThreads, 0, static_cast<cudaStream_t>(stream),
              reinterpret_cast<const __half*>(src), ld_src, rows, static_cast<int>(cols / 4), p, seed, offset,
              reinterpret_cast<const unsigned long long*>(step_dev), src_inv_scale, reinterpret_cast<__half*>(dst), ld_dst,
              dst_inv_scale);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}
