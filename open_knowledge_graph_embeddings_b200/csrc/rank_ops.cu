// Filtered ranking as integer counting (openkge/dataset.py:423-453) for sm_100a.
//
//   rank_count            one CTA per ranked answer, streaming its prefix row of a MATERIALISED
//                         score matrix once with 16-byte loads (HBM-bound: B*N*4 bytes);
//   rank_true_score /     the sparse corrections that complete the fused path, where the dense
//   rank_filter_correct   count-greater runs inside the scoring GEMM epilogue (gemm_tf32.cu).
//
// Counting identity used by both paths. With masked[n] = (n in F) ? -1e8 : s[n] the reference counts
//   greater = #{n : t < masked[n]} = #{n : t < s[n]} - #{f in F : t < s[f]} + |F| * [t < -1e8]
//   equal   = #{n : t == masked[n]} = (same with ==)
// which is exact for any inputs because F is a set (unique column indices).
#include "okge_common.cuh"

namespace okge {
namespace {

constexpr float kMaskFill = -1e8f;  // openkge/dataset.py:440

__global__ void __launch_bounds__(256)
rank_count_kernel(const float* __restrict__ scores, int64_t lds, int64_t N,
                  const int32_t* __restrict__ ans_row, const int32_t* __restrict__ alt_ptr,
                  const int32_t* __restrict__ alt_idx, const int32_t* __restrict__ filt_ptr,
                  const int32_t* __restrict__ filt_idx, float* __restrict__ true_score,
                  int32_t* __restrict__ greater, int32_t* __restrict__ equal) {
  pdl_wait_and_trigger();
  __shared__ float s_true;
  __shared__ int s_cnt[2];
  const int j = blockIdx.x;
  const int row = __ldg(ans_row + j);
  const float* srow = scores + static_cast<int64_t>(row) * lds;

  // true score: max over the alternative mentions of this answer, from the UNMASKED row (:436-438)
  if (threadIdx.x < 32) {
    float t = -INFINITY;
    for (int a = __ldg(alt_ptr + j) + threadIdx.x; a < __ldg(alt_ptr + j + 1); a += 32)
      t = fmaxf(t, srow[__ldg(alt_idx + a)]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
    if (threadIdx.x == 0) { s_true = t; s_cnt[0] = 0; s_cnt[1] = 0; }
  }
  __syncthreads();
  const float t = s_true;

  int g = 0, e = 0;
  // dense pass over all candidates
  const bool vec = ((reinterpret_cast<uintptr_t>(srow) & 15u) == 0);
  if (vec) {
    const int64_t n4 = N / 4;
    const float4* s4 = reinterpret_cast<const float4*>(srow);
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = ldg_nc_f4(s4 + i);
      g += (t < v.x) + (t < v.y) + (t < v.z) + (t < v.w);
      e += (t == v.x) + (t == v.y) + (t == v.z) + (t == v.w);
    }
    for (int64_t i = n4 * 4 + threadIdx.x; i < N; i += blockDim.x) {
      const float v = srow[i];
      g += (t < v); e += (t == v);
    }
  } else {
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
      const float v = srow[i];
      g += (t < v); e += (t == v);
    }
  }
  // filter correction
  const int f_lo = __ldg(filt_ptr + row), f_hi = __ldg(filt_ptr + row + 1);
  for (int f = f_lo + threadIdx.x; f < f_hi; f += blockDim.x) {
    const float v = srow[__ldg(filt_idx + f)];
    g -= (t < v); e -= (t == v);
  }
  if (threadIdx.x == 0) {
    const int nf = f_hi - f_lo;
    g += nf * (t < kMaskFill);
    e += nf * (t == kMaskFill);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    g += __shfl_xor_sync(0xffffffffu, g, o);
    e += __shfl_xor_sync(0xffffffffu, e, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s_cnt[0], g); atomicAdd(&s_cnt[1], e); }
  __syncthreads();
  if (threadIdx.x == 0) {
    true_score[j] = t;
    greater[j] = s_cnt[0];
    equal[j] = s_cnt[1];
  }
}

__global__ void __launch_bounds__(256)
rank_true_score_kernel(const float* __restrict__ sel, int64_t lds, const int32_t* __restrict__ ans_row,
                       const int32_t* __restrict__ alt_ptr, const int32_t* __restrict__ alt_pos,
                       int64_t Q, float* __restrict__ true_score) {
  pdl_wait_and_trigger();
  for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < Q;
       j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float* srow = sel + static_cast<int64_t>(__ldg(ans_row + j)) * lds;
    float t = true_score[j];
    for (int a = __ldg(alt_ptr + j); a < __ldg(alt_ptr + j + 1); ++a) {
      const int pos = __ldg(alt_pos + a);
      if (pos >= 0) t = fmaxf(t, srow[pos]);
    }
    true_score[j] = t;
  }
}

// one warp per ranked answer
__global__ void __launch_bounds__(256)
rank_filter_correct_kernel(const float* __restrict__ sel, int64_t lds,
                           const int32_t* __restrict__ ans_row, int64_t Q,
                           const int32_t* __restrict__ filt_ptr, const int32_t* __restrict__ filt_pos,
                           const float* __restrict__ thresh, int add_mask_terms,
                           int32_t* __restrict__ greater, int32_t* __restrict__ equal) {
  pdl_wait_and_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t j = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; j < Q; j += warps) {
    const int row = __ldg(ans_row + j);
    const float* srow = sel + static_cast<int64_t>(row) * lds;
    const float t = __ldg(thresh + j);
    const int f_lo = __ldg(filt_ptr + row), f_hi = __ldg(filt_ptr + row + 1);
    int g = 0, e = 0;
    for (int f = f_lo + lane; f < f_hi; f += 32) {
      const int pos = __ldg(filt_pos + f);
      if (pos >= 0) {
        const float v = srow[pos];
        g -= (t < v); e -= (t == v);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      g += __shfl_xor_sync(0xffffffffu, g, o);
      e += __shfl_xor_sync(0xffffffffu, e, o);
    }
    if (lane == 0) {
      if (add_mask_terms) {
        const int nf = f_hi - f_lo;
        g += nf * (t < kMaskFill);
        e += nf * (t == kMaskFill);
      }
      if (g) atomicAdd(greater + j, g);
      if (e) atomicAdd(equal + j, e);
    }
  }
}

}  // namespace
}  // namespace okge

using namespace okge;

extern "C" int okge_rank_count(const float* scores, int64_t lds, int64_t B, int64_t N,
                               const int32_t* ans_row, const int32_t* alt_ptr, const int32_t* alt_idx,
                               int64_t Q, const int32_t* filt_ptr, const int32_t* filt_idx,
                               float* true_score, int32_t* greater, int32_t* equal,
                               okge_stream_t stream) {
  if (Q == 0) return OKGE_OK;
  OKGE_REQUIRE(scores && ans_row && alt_ptr && alt_idx && filt_ptr && true_score && greater && equal,
               "null pointer");
  OKGE_REQUIRE(B > 0 && N > 0 && lds >= N, "bad score matrix shape");
  OKGE_REQUIRE(Q < 2147483647LL, "too many ranked answers");
  OKGE_LAUNCH((rank_count_kernel), static_cast<unsigned>(Q), 256, 0, static_cast<cudaStream_t>(stream), scores, lds, N, ans_row, alt_ptr, alt_idx, filt_ptr, filt_idx, true_score, greater, equal);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_rank_true_score(const float* sel_scores, int64_t lds, const int32_t* ans_row,
                                    const int32_t* alt_ptr, const int32_t* alt_pos, int64_t Q,
                                    float* true_score, okge_stream_t stream) {
  if (Q == 0) return OKGE_OK;
  OKGE_REQUIRE(sel_scores && ans_row && alt_ptr && alt_pos && true_score, "null pointer");
  int64_t blocks = ceil_div64(Q, 256);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  OKGE_LAUNCH((rank_true_score_kernel), static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream), sel_scores, lds, ans_row, alt_ptr, alt_pos, Q, true_score);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_rank_filter_correct(const float* sel_scores, int64_t lds, const int32_t* ans_row,
                                        int64_t Q, const int32_t* filt_ptr, const int32_t* filt_pos,
                                        const float* thresh, int32_t add_mask_terms, int32_t* greater,
                                        int32_t* equal, okge_stream_t stream) {
  if (Q == 0) return OKGE_OK;
  OKGE_REQUIRE(sel_scores && ans_row && filt_ptr && filt_pos && thresh && greater && equal, "null pointer");
  int64_t blocks = ceil_div64(Q, 8);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  OKGE_LAUNCH((rank_filter_correct_kernel), static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream), sel_scores, lds, ans_row, Q, filt_ptr, filt_pos, thresh, add_mask_terms, greater, equal);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}
