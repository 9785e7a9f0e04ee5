// Optimizer updates of the OpenKGE hot path for sm_100a (HBM-bound, 16-byte vector accesses).
//
// torch.optim.Adagrad / Adam exactly as the reference drives them through OptimRegime
// (utils/optim.py:28-29, 139-160, 194-201): the regime re-instantiates the configured optimizer on
// the param_groups of a bootstrap Adam(lr=0), so Adagrad runs with eps = 1e-8 and the yaml's
// weight_decay, DENSE over every row of every table. The dense kernels stream p, g, G once
// (5 x 4 bytes per element: read p, g, G; write p, G); the row-wise kernels touch listed rows only.
#include "okge_common.cuh"

namespace okge {
namespace {

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float lr, float b1,
                                          float b2, float eps, float wd, float bc1, float sqrt_bc2) {
  g = __fadd_rn(g, __fmul_rn(wd, p));
  m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), 1.f - b1));                  // lerp_(g, 1 - b1)
  v = __fadd_rn(__fmul_rn(v, b2), __fmul_rn(__fmul_rn(g, g), 1.f - b2));   // mul_(b2).addcmul_(g, g, 1 - b2)
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), sqrt_bc2), eps);
  const float step_size = lr / bc1;
  p = __fadd_rn(p, __fmul_rn(-step_size, __fdiv_rn(m, denom)));
}

__global__ void __launch_bounds__(256)
adagrad_dense_kernel(float* __restrict__ param, const float* __restrict__ grad,
                     float* __restrict__ state, int64_t n, float clr, float eps, float wd, int vec) {
  pdl_wait_and_trigger();
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t nth = static_cast<int64_t>(gridDim.x) * blockDim.x;
  if (vec) {
    const int64_t n4 = n / 4;
    float4* p4 = reinterpret_cast<float4*>(param);
    const float4* g4 = reinterpret_cast<const float4*>(grad);
    float4* s4 = reinterpret_cast<float4*>(state);
    for (int64_t i = tid; i < n4; i += nth) {
      float4 p = p4[i], G = s4[i];
      const float4 g = ldg_nc_f4(g4 + i);
      adagrad_elem(p.x, g.x, G.x, clr, eps, wd);
      adagrad_elem(p.y, g.y, G.y, clr, eps, wd);
      adagrad_elem(p.z, g.z, G.z, clr, eps, wd);
      adagrad_elem(p.w, g.w, G.w, clr, eps, wd);
      p4[i] = p;
      s4[i] = G;
    }
    for (int64_t i = n4 * 4 + tid; i < n; i += nth) adagrad_elem(param[i], grad[i], state[i], clr, eps, wd);
  } else {
    for (int64_t i = tid; i < n; i += nth) adagrad_elem(param[i], grad[i], state[i], clr, eps, wd);
  }
}

__global__ void __launch_bounds__(256)
adam_dense_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m_,
                  float* __restrict__ v_, int64_t n, float lr, float b1, float b2, float eps, float wd,
                  float bc1, float sqrt_bc2, int vec) {
  pdl_wait_and_trigger();
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t nth = static_cast<int64_t>(gridDim.x) * blockDim.x;
  if (vec) {
    const int64_t n4 = n / 4;
    float4* p4 = reinterpret_cast<float4*>(param);
    const float4* g4 = reinterpret_cast<const float4*>(grad);
    float4* m4 = reinterpret_cast<float4*>(m_);
    float4* v4 = reinterpret_cast<float4*>(v_);
    for (int64_t i = tid; i < n4; i += nth) {
      float4 p = p4[i], m = m4[i], v = v4[i];
      const float4 g = ldg_nc_f4(g4 + i);
      adam_elem(p.x, g.x, m.x, v.x, lr, b1, b2, eps, wd, bc1, sqrt_bc2);
      adam_elem(p.y, g.y, m.y, v.y, lr, b1, b2, eps, wd, bc1, sqrt_bc2);
      adam_elem(p.z, g.z, m.z, v.z, lr, b1, b2, eps, wd, bc1, sqrt_bc2);
      adam_elem(p.w, g.w, m.w, v.w, lr, b1, b2, eps, wd, bc1, sqrt_bc2);
      p4[i] = p; m4[i] = m; v4[i] = v;
    }
    for (int64_t i = n4 * 4 + tid; i < n; i += nth)
      adam_elem(param[i], grad[i], m_[i], v_[i], lr, b1, b2, eps, wd, bc1, sqrt_bc2);
  } else {
    for (int64_t i = tid; i < n; i += nth)
      adam_elem(param[i], grad[i], m_[i], v_[i], lr, b1, b2, eps, wd, bc1, sqrt_bc2);
  }
}

// one warp per listed row
__global__ void __launch_bounds__(256)
adagrad_rows_kernel(float* __restrict__ param, float* __restrict__ state, int64_t ld,
                    const float* __restrict__ grad_rows, int64_t ld_grad,
                    const int32_t* __restrict__ row_ids, const int32_t* __restrict__ slot_map, int64_t n_rows, int D,
                    float clr, float eps, float wd) {
  pdl_wait_and_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t i = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; i < n_rows; i += warps) {
    const int64_t r = __ldg(row_ids + i);
    // with a slot map the list may repeat ids: only the position that owns the id's slot carries its (summed) gradient
    if (r < 0 || (slot_map != nullptr && __ldg(slot_map + r) != static_cast<int32_t>(i))) continue;
    float* p = param + r * ld;
    float* G = state + r * ld;
    const float* g = grad_rows + i * ld_grad;
    for (int c = lane; c < D; c += 32) adagrad_elem(p[c], g[c], G[c], clr, eps, wd);
  }
}

__global__ void __launch_bounds__(256)
adam_rows_kernel(float* __restrict__ param, float* __restrict__ m_, float* __restrict__ v_, int64_t ld,
                 const float* __restrict__ grad_rows, int64_t ld_grad,
                 const int32_t* __restrict__ row_ids, const int32_t* __restrict__ slot_map, int64_t n_rows, int D,
                 float lr, float b1, float b2, float eps, float wd, float bc1, float sqrt_bc2) {
  pdl_wait_and_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t i = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; i < n_rows; i += warps) {
    const int64_t r = __ldg(row_ids + i);
    if (r < 0 || (slot_map != nullptr && __ldg(slot_map + r) != static_cast<int32_t>(i))) continue;
    const float* g = grad_rows + i * ld_grad;
    for (int c = lane; c < D; c += 32)
      adam_elem(param[r * ld + c], g[c], m_[r * ld + c], v_[r * ld + c], lr, b1, b2, eps, wd, bc1, sqrt_bc2);
  }
}

// ---- sparse extra-gradient rows for the fused dE + Adagrad epilogue --------------------------------------------------
// The B looked-up rows of a batch receive, besides their 1-vs-all gradient row dE[id], the gradient of the lookup
// itself. Duplicated ids share ONE slot (the largest batch position with that id), so the epilogue adds exactly one
// extra row per table row. slot_map is a persistent [table rows] int32 buffer of -1 that is restored after the step.
__global__ void row_slots_build_kernel(const int32_t* __restrict__ ids, int64_t n, int32_t skip_id,
                                       int32_t* __restrict__ slot_map) {
  pdl_wait_and_trigger();
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int32_t id = __ldg(ids + i);
  if (id != skip_id) atomicMax(slot_map + id, static_cast<int32_t>(i));
}

__global__ void row_slots_clear_kernel(const int32_t* __restrict__ ids, int64_t n, int32_t skip_id,
                                       int32_t* __restrict__ slot_map) {
  pdl_wait_and_trigger();
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int32_t id = __ldg(ids + i);
  if (id != skip_id) slot_map[id] = -1;
}

// one warp per batch row: extra[slot_map[ids[i]], :] += grad[i, :]
__global__ void __launch_bounds__(256)
row_slots_accumulate_kernel(const float* __restrict__ grad, int64_t ld_grad, const int32_t* __restrict__ ids, int64_t n,
                            int D, int32_t skip_id, const int32_t* __restrict__ slot_map, float* __restrict__ extra,
                            int64_t ld_extra) {
  pdl_wait_and_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t i = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; i < n; i += warps) {
    const int32_t id = __ldg(ids + i);
    if (id == skip_id) continue;
    const int32_t slot = __ldg(slot_map + id);
    if (slot < 0) continue;
    const float* g = grad + i * ld_grad;
    float* out = extra + static_cast<int64_t>(slot) * ld_extra;
    for (int c = lane; c < D; c += 32) atomicAdd(out + c, g[c]);
  }
}

// Adagrad on a few leading rows whose gradient is only the optional slot row (the PAD / UNK rows of an entity table:
// no 1-vs-all gradient): g = slot_map[r] >= 0 ? extra[slot_map[r], :] : 0. One warp per row.
__global__ void __launch_bounds__(256)
adagrad_slot_rows_kernel(float* __restrict__ param, float* __restrict__ state, int64_t ld, int n_rows, int D,
                         const int32_t* __restrict__ slot_map, const float* __restrict__ extra, int64_t ld_extra,
                         float clr, float eps, float wd) {
  pdl_wait_and_trigger();
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rows) return;
  const int32_t slot = slot_map != nullptr ? __ldg(slot_map + r) : -1;
  const float* g = (slot >= 0 && extra != nullptr) ? extra + static_cast<int64_t>(slot) * ld_extra : nullptr;
  for (int c = lane; c < D; c += 32) adagrad_elem(param[r * ld + c], g != nullptr ? g[c] : 0.f, state[r * ld + c], clr, eps, wd);
}

// Adagrad over a WHOLE table whose gradient lives in a compact slot table (token tables of a batch-shared step: a few
// ten thousand of 200 k rows receive a gradient, but weight decay makes the dense reference step touch every row):
// g = slot_map[r] >= 0 ? slot_grad[slot_map[r], :] : 0. One warp per row, float4: param and state once each way
// (16 B/element) plus the touched gradient rows - no dense gradient to zero-fill, write and read back.
__global__ void __launch_bounds__(256)
adagrad_slot_table_kernel(float* __restrict__ param, float* __restrict__ state, int64_t n_rows, int D4,
                          const int32_t* __restrict__ slot_map, const float* __restrict__ slot_grad, float clr, float eps,
                          float wd) {
  pdl_wait_and_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; r < n_rows; r += warps) {
    const int32_t slot = __ldg(slot_map + r);
    float4* p4 = reinterpret_cast<float4*>(param) + r * D4;
    float4* s4 = reinterpret_cast<float4*>(state) + r * D4;
    const float4* g4 = slot >= 0 ? reinterpret_cast<const float4*>(slot_grad) + static_cast<int64_t>(slot) * D4 : nullptr;
    for (int c = lane; c < D4; c += 32) {
      float4 p = p4[c], st = s4[c];
      const float4 g = g4 != nullptr ? __ldg(g4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      adagrad_elem(p.x, g.x, st.x, clr, eps, wd);
      adagrad_elem(p.y, g.y, st.y, clr, eps, wd);
      adagrad_elem(p.z, g.z, st.z, clr, eps, wd);
      adagrad_elem(p.w, g.w, st.w, clr, eps, wd);
      p4[c] = p;
      s4[c] = st;
    }
  }
}

int dense_grid(int64_t n) {
  int64_t blocks = ceil_div64(ceil_div64(n, 4), 256);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

bool all_aligned16(const void* a, const void* b, const void* c, const void* d) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
           reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(d)) & 15u) == 0;
}

}  // namespace
}  // namespace okge

using namespace okge;

extern "C" int okge_adagrad_dense(float* param, const float* grad, float* state_sum, int64_t n,
                                  float clr, float eps, float weight_decay, okge_stream_t stream) {
  if (n == 0) return OKGE_OK;
  OKGE_REQUIRE(param && grad && state_sum, "null pointer");
  const int vec = all_aligned16(param, grad, state_sum, nullptr) ? 1 : 0;
  OKGE_LAUNCH((adagrad_dense_kernel), dense_grid(n), 256, 0, static_cast<cudaStream_t>(stream), param, grad, state_sum, n, clr, eps, weight_decay, vec);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_adagrad_rows(float* param, float* state_sum, int64_t ld, const float* grad_rows,
                                 int64_t ld_grad, const int32_t* row_ids, const int32_t* slot_map, int64_t n_rows,
                                 int64_t D, float clr, float eps, float weight_decay, okge_stream_t stream) {
  if (n_rows == 0) return OKGE_OK;
  OKGE_REQUIRE(param && state_sum && grad_rows && row_ids, "null pointer");
  OKGE_REQUIRE(D > 0 && ld >= D && ld_grad >= D, "bad row shape");
  int64_t blocks = ceil_div64(n_rows, 8);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  OKGE_LAUNCH((adagrad_rows_kernel), static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream), param, state_sum, ld, grad_rows, ld_grad, row_ids, slot_map, n_rows, static_cast<int>(D), clr, eps,
      weight_decay);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_adam_dense(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                               int64_t n, float lr, float beta1, float beta2, float eps,
                               float weight_decay, float bias_correction1, float bias_correction2,
                               okge_stream_t stream) {
  if (n == 0) return OKGE_OK;
  OKGE_REQUIRE(param && grad && exp_avg && exp_avg_sq, "null pointer");
  OKGE_REQUIRE(bias_correction1 > 0.f && bias_correction2 > 0.f, "bias corrections must be > 0");
  const int vec = all_aligned16(param, grad, exp_avg, exp_avg_sq) ? 1 : 0;
  OKGE_LAUNCH((adam_dense_kernel), dense_grid(n), 256, 0, static_cast<cudaStream_t>(stream), param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, bias_correction1,
      sqrtf(bias_correction2), vec);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_adam_rows(float* param, float* exp_avg, float* exp_avg_sq, int64_t ld,
                              const float* grad_rows, int64_t ld_grad, const int32_t* row_ids,
                              const int32_t* slot_map, int64_t n_rows, int64_t D, float lr, float beta1, float beta2, float eps,
                              float weight_decay, float bias_correction1, float bias_correction2,
                              okge_stream_t stream) {
  if (n_rows == 0) return OKGE_OK;
  OKGE_REQUIRE(param && exp_avg && exp_avg_sq && grad_rows && row_ids, "null pointer");
  OKGE_REQUIRE(D > 0 && ld >= D && ld_grad >= D, "bad row shape");
  OKGE_REQUIRE(bias_correction1 > 0.f && bias_correction2 > 0.f, "bias corrections must be > 0");
  int64_t blocks = ceil_div64(n_rows, 8);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  OKGE_LAUNCH((adam_rows_kernel), static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream), param, exp_avg, exp_avg_sq, ld, grad_rows, ld_grad, row_ids, slot_map, n_rows, static_cast<int>(D), lr,
      beta1, beta2, eps, weight_decay, bias_correction1, sqrtf(bias_correction2));
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_row_slots_build(const int32_t* ids, int64_t n, int32_t skip_id, int32_t* slot_map,
                                    okge_stream_t stream) {
  if (n == 0) return OKGE_OK;
  OKGE_REQUIRE(ids && slot_map, "null pointer");
  OKGE_LAUNCH((row_slots_build_kernel), static_cast<unsigned>(ceil_div64(n, 256)), 256, 0, static_cast<cudaStream_t>(stream), ids, n, skip_id, slot_map);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_row_slots_accumulate(const float* grad, int64_t ld_grad, const int32_t* ids, int64_t n, int64_t D,
                                         int32_t skip_id, const int32_t* slot_map, float* extra, int64_t ld_extra,
                                         okge_stream_t stream) {
  if (n == 0) return OKGE_OK;
  OKGE_REQUIRE(grad && ids && slot_map && extra, "null pointer");
  OKGE_REQUIRE(D > 0 && ld_grad >= D && ld_extra >= D, "bad row shape");
  int64_t blocks = ceil_div64(n, 8);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  OKGE_LAUNCH((row_slots_accumulate_kernel), static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream), grad, ld_grad, ids, n, static_cast<int>(D), skip_id, slot_map, extra, ld_extra);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_row_slots_clear(const int32_t* ids, int64_t n, int32_t skip_id, int32_t* slot_map,
                                    okge_stream_t stream) {
  if (n == 0) return OKGE_OK;
  OKGE_REQUIRE(ids && slot_map, "null pointer");
  OKGE_LAUNCH((row_slots_clear_kernel), static_cast<unsigned>(ceil_div64(n, 256)), 256, 0, static_cast<cudaStream_t>(stream), ids, n, skip_id, slot_map);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_adagrad_slot_rows(float* param, float* state_sum, int64_t ld, int64_t n_rows, int64_t D,
                                      const int32_t* slot_map, const float* extra, int64_t ld_extra, float clr, float eps,
                                      float weight_decay, okge_stream_t stream) {
  if (n_rows == 0) return OKGE_OK;
  OKGE_REQUIRE(param && state_sum, "null pointer");
  OKGE_REQUIRE(D > 0 && ld >= D && n_rows < (1 << 20), "bad row shape");
  OKGE_LAUNCH((adagrad_slot_rows_kernel), static_cast<unsigned>(ceil_div64(n_rows, 8)), 256, 0, static_cast<cudaStream_t>(stream), param, state_sum, ld, static_cast<int>(n_rows), static_cast<int>(D), slot_map, extra, ld_extra, clr, eps, weight_decay);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_adagrad_slot_table(float* param, float* state_sum, int64_t n_rows, int64_t D, const int32_t* slot_map,
                                       const float* slot_grad, float clr, float eps, float weight_decay,
                                       okge_stream_t stream) {
  if (n_rows == 0) return OKGE_OK;
  OKGE_REQUIRE(param && state_sum && slot_map && slot_grad, "null pointer");
  OKGE_REQUIRE(D > 0 && D % 4 == 0, "D must be a multiple of 4");
  OKGE_REQUIRE(all_aligned16(param, state_sum, slot_grad, nullptr), "param / state / slot_grad must be 16-byte aligned");
  int64_t blocks = ceil_div64(n_rows, 8);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  OKGE_LAUNCH((adagrad_slot_table_kernel), static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream), param, state_sum, n_rows, static_cast<int>(D / 4), slot_map, slot_grad, clr, eps, weight_decay);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}
