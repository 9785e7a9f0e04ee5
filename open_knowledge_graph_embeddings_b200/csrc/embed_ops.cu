// HBM-bound embedding kernels of the OpenKGE hot path for sm_100a:
//   gather_rows / scatter_add_rows   Lookup embedder              openkge/model.py:455-480
//   gather_pool fwd / bwd            UnigramPooling embedder      openkge/model.py:762-774
//   dropout                          F.dropout on [n, D] operands openkge/model.py:461-470, 783-786
//   fold_query fwd / bwd             ComplEx / DistMult prefix    openkge/model.py:206-215, 270-272
//
// Layout: one warp per output row, each lane owns 16-byte (float4) column slices, so every global
// access is a fully coalesced 512-byte warp transaction; token ids of a row are loaded once by the
// first L lanes and broadcast with shuffles; grids are sized in multiples of the SM count.
#include "okge_common.cuh"

#include <math.h>

namespace okge {

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kMaxL = 32;  // token slots per row handled by one warp-load of ids

__device__ __forceinline__ int64_t global_warp_id() {
  return (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
}
__device__ __forceinline__ int64_t global_warp_count() {
  return (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
}

int rows_grid(int64_t n_rows) {
  int64_t blocks = ceil_div64(n_rows, kWarpsPerBlock);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;  // 16 resident 256-thread CTAs max / SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ------------------------------------------------------------------------------------------
// Lookup gather / scatter-add
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(kThreads)
gather_rows_kernel(const float* __restrict__ table, int64_t ld_table, const int32_t* __restrict__ ids,
                   int64_t n, int D4, float* __restrict__ out, int64_t ld_out) {
  pdl_wait_and_trigger();
  const int lane = threadIdx.x & 31;
  for (int64_t i = global_warp_id(); i < n; i += global_warp_count()) {
    const int64_t r = __ldg(ids + i);
    const float4* src = reinterpret_cast<const float4*>(table + r * ld_table);
    float4* dst = reinterpret_cast<float4*>(out + i * ld_out);
    for (int c = lane; c < D4; c += 32) dst[c] = __ldg(src + c);
  }
}

__global__ void __launch_bounds__(kThreads)
scatter_add_rows_kernel(const float* __restrict__ grad, int64_t ld_grad,
                        const int32_t* __restrict__ ids, int64_t n, int D4, int32_t skip_id,
                        float* __restrict__ grad_table, int64_t ld_table) {
  pdl_wait_and_trigger();
  const int lane = threadIdx.x & 31;
  for (int64_t i = global_warp_id(); i < n; i += global_warp_count()) {
    const int32_t r = __ldg(ids + i);
    if (r == skip_id) continue;
    const float4* src = reinterpret_cast<const float4*>(grad + i * ld_grad);
    float* dst = grad_table + static_cast<int64_t>(r) * ld_table;
    for (int c = lane; c < D4; c += 32) red_add_f4(dst + 4 * c, __ldg(src + c));
  }
}

// ------------------------------------------------------------------------------------------
// token gather + pooling
// ------------------------------------------------------------------------------------------

template <int MODE>
__global__ void __launch_bounds__(kThreads)
gather_pool_fwd_kernel(const float* __restrict__ tok_table, int64_t ld_table,
                       const int32_t* __restrict__ id_rows, int L, const int32_t* __restrict__ ids,
                       int64_t id_start, int64_t n, int D4, float* __restrict__ out, int64_t ld_out) {
  pdl_wait_and_trigger();
  const int lane = threadIdx.x & 31;
  for (int64_t i = global_warp_id(); i < n; i += global_warp_count()) {
    const int64_t row = (ids != nullptr) ? static_cast<int64_t>(__ldg(ids + i)) : id_start + i;
    const int32_t my_tok = (lane < L) ? __ldg(id_rows + row * L + lane) : 0;
    float len = 1.f;
    if (MODE == OKGE_POOL_MEAN) {
      const unsigned nz = __ballot_sync(0xffffffffu, lane < L && my_tok > 0);
      len = static_cast<float>(__popc(nz)) + 1e-12f;   // (input > 0).sum() + 1e-12, model.py:771-772
    }
    float4* dst = reinterpret_cast<float4*>(out + i * ld_out);
    for (int c0 = 0; c0 < D4; c0 += 32) {      // warp-uniform trip count (shuffles inside)
      const int c = c0 + lane;
      const bool active = c < D4;
      float4 acc;
      if (MODE == OKGE_POOL_MAX) acc = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      else acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 5
      for (int l = 0; l < L; ++l) {
        const int32_t tok = __shfl_sync(0xffffffffu, my_tok, l);
        if (active) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(tok_table + static_cast<int64_t>(tok) * ld_table) + c);
          if (MODE == OKGE_POOL_MAX) {
            acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y);
            acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w);
          } else {
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
        }
      }
      if (MODE == OKGE_POOL_MEAN) {
        // the reference divides sum / (len + 1e-12): keep a true division for identical rounding
        acc.x = acc.x / len; acc.y = acc.y / len; acc.z = acc.z / len; acc.w = acc.w / len;
      }
      if (active) dst[c] = acc;
    }
  }
}

// Scatter-add backward. Tokens 2 (BOS) and 3 (EOS) appear in (nearly) every row
// (openkge/dataset.py insert_start/insert_end), so they are accumulated per block in shared
// memory and flushed with one vector atomic per block and column slice instead of one per row.
template <int MODE>
__global__ void __launch_bounds__(kThreads)
gather_pool_bwd_kernel(const float* __restrict__ grad_out, int64_t ld_grad,
                       const float* __restrict__ tok_table, int64_t ld_table,
                       const int32_t* __restrict__ id_rows, int L, const int32_t* __restrict__ ids,
                       int64_t id_start, int64_t n, int D4, float* __restrict__ grad_tok,
                       int hot_lo, int hot_n, const int32_t* __restrict__ slot_map) {
  pdl_wait_and_trigger();
  extern __shared__ float4 hot_acc[];  // [hot_n][D4]
  const int lane = threadIdx.x & 31;
  for (int t = threadIdx.x; t < hot_n * D4; t += blockDim.x) hot_acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  for (int64_t i = global_warp_id(); i < n; i += global_warp_count()) {
    const int64_t row = (ids != nullptr) ? static_cast<int64_t>(__ldg(ids + i)) : id_start + i;
    const int32_t my_tok = (lane < L) ? __ldg(id_rows + row * L + lane) : 0;
    float len = 1.f;
    if (MODE == OKGE_POOL_MEAN) {
      const unsigned nz = __ballot_sync(0xffffffffu, lane < L && my_tok > 0);
      len = static_cast<float>(__popc(nz)) + 1e-12f;
    }
    const float4* g4 = reinterpret_cast<const float4*>(grad_out + i * ld_grad);
    for (int c0 = 0; c0 < D4; c0 += 32) {
      const int c = c0 + lane;
      const bool active = c < D4;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active) g = __ldg(g4 + c);
      if (MODE == OKGE_POOL_MEAN) { g.x = g.x / len; g.y = g.y / len; g.z = g.z / len; g.w = g.w / len; }

      // for MAX: gradient goes to the first slot that attains the maximum (torch.max(dim) semantics)
      int arg_x = 0, arg_y = 0, arg_z = 0, arg_w = 0;
      if (MODE == OKGE_POOL_MAX) {
        float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (int l = 0; l < L; ++l) {
          const int32_t tok = __shfl_sync(0xffffffffu, my_tok, l);
          if (active) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(tok_table + static_cast<int64_t>(tok) * ld_table) + c);
            if (v.x > best.x) { best.x = v.x; arg_x = l; }
            if (v.y > best.y) { best.y = v.y; arg_y = l; }
            if (v.z > best.z) { best.z = v.z; arg_z = l; }
            if (v.w > best.w) { best.w = v.w; arg_w = l; }
          }
        }
      }

      for (int l = 0; l < L; ++l) {
        const int32_t tok = __shfl_sync(0xffffffffu, my_tok, l);
        if (!active || tok == 0) continue;  // padding_idx = 0 never receives gradient
        float4 gl = g;
        if (MODE == OKGE_POOL_MAX) {
          gl.x = (arg_x == l) ? g.x : 0.f; gl.y = (arg_y == l) ? g.y : 0.f;
          gl.z = (arg_z == l) ? g.z : 0.f; gl.w = (arg_w == l) ? g.w : 0.f;
        }
        const int h = tok - hot_lo;
        if (h >= 0 && h < hot_n) {
          float* a = reinterpret_cast<float*>(&hot_acc[h * D4 + c]);
          atomicAdd(a + 0, gl.x); atomicAdd(a + 1, gl.y); atomicAdd(a + 2, gl.z); atomicAdd(a + 3, gl.w);
        } else {
          // destination row: the token's row of the dense gradient table, or its slot in a compact gradient table
          const int64_t dst = slot_map != nullptr ? static_cast<int64_t>(__ldg(slot_map + tok)) : static_cast<int64_t>(tok);
          red_add_f4(grad_tok + dst * ld_table + 4 * c, gl);
        }
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < hot_n * D4; t += blockDim.x) {
    const int h = t / D4, c = t % D4;
    const float4 v = hot_acc[t];
    if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
      const int64_t dst = slot_map != nullptr ? static_cast<int64_t>(__ldg(slot_map + hot_lo + h)) : static_cast<int64_t>(hot_lo + h);
      red_add_f4(grad_tok + dst * ld_table + 4 * c, v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// dropout (Philox4x32-10, one 128-bit block per 4 consecutive elements)
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
dropout_kernel(const float* __restrict__ x, int64_t n, float p, float scale, uint64_t seed,
               uint64_t offset, const unsigned long long* __restrict__ step_dev, float* __restrict__ out) {
  pdl_wait_and_trigger();
  const int64_t n4 = (n + 3) / 4;
  if (step_dev != nullptr) offset += static_cast<uint64_t>(*step_dev) << 44;   // per-step stream of a replayed launch
  const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint64_t c = static_cast<uint64_t>(i) + offset;
    const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(c), static_cast<uint32_t>(c >> 32), 0u, 0u), key);
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
    const int64_t base = i * 4;
    if (base + 3 < n && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
      float4 o;
      // u in [0,1): keep iff u >= p
      o.x = ((rr[0] >> 8) * (1.0f / 16777216.0f) >= p) ? v.x * scale : 0.f;
      o.y = ((rr[1] >> 8) * (1.0f / 16777216.0f) >= p) ? v.y * scale : 0.f;
      o.z = ((rr[2] >> 8) * (1.0f / 16777216.0f) >= p) ? v.z * scale : 0.f;
      o.w = ((rr[3] >> 8) * (1.0f / 16777216.0f) >= p) ? v.w * scale : 0.f;
      reinterpret_cast<float4*>(out)[i] = o;
    } else {
      for (int k = 0; k < 4 && base + k < n; ++k)
        out[base + k] = ((rr[k] >> 8) * (1.0f / 16777216.0f) >= p) ? x[base + k] * scale : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------
// query folding
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
fold_query_kernel(int kind, const int32_t* __restrict__ kinds, const float* __restrict__ a,
                  const float* __restrict__ b, int64_t Bq, int D, float* __restrict__ q) {
  pdl_wait_and_trigger();
  const int H = D / 2;
  const int64_t total = (kind == OKGE_FOLD_DISTMULT) ? Bq * D : Bq * H;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (kind == OKGE_FOLD_DISTMULT) {
      q[i] = a[i] * b[i];
    } else {
      const int64_t r = i / H;
      const int c = static_cast<int>(i % H);
      const float a1 = a[r * D + c], a2 = a[r * D + H + c];
      const float b1 = b[r * D + c], b2 = b[r * D + H + c];
      const int row_kind = kinds != nullptr ? __ldg(kinds + r) : kind;     // per-row kind: po and sp rows in one launch
      if (row_kind == OKGE_FOLD_COMPLEX_SP) {
        q[r * D + c] = a1 * b1 - a2 * b2;
        q[r * D + H + c] = a2 * b1 + a1 * b2;
      } else {
        q[r * D + c] = a1 * b1 + a2 * b2;
        q[r * D + H + c] = a2 * b1 - a1 * b2;
      }
    }
  }
}

__global__ void __launch_bounds__(256)
fold_query_bwd_kernel(int kind, const int32_t* __restrict__ kinds, const float* __restrict__ a,
                      const float* __restrict__ b, const float* __restrict__ gq, int64_t Bq, int D,
                      float* __restrict__ ga, float* __restrict__ gb) {
  pdl_wait_and_trigger();
  const int H = D / 2;
  const int64_t total = (kind == OKGE_FOLD_DISTMULT) ? Bq * D : Bq * H;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (kind == OKGE_FOLD_DISTMULT) {
      const float g = gq[i];
      ga[i] = g * b[i];
      gb[i] = g * a[i];
    } else {
      const int64_t r = i / H;
      const int c = static_cast<int>(i % H);
      const int64_t i1 = r * D + c, i2 = r * D + H + c;
      const float a1 = a[i1], a2 = a[i2], b1 = b[i1], b2 = b[i2];
      const float g1 = gq[i1], g2 = gq[i2];
      const int row_kind = kinds != nullptr ? __ldg(kinds + r) : kind;
      if (row_kind == OKGE_FOLD_COMPLEX_SP) {
        // q1 = a1 b1 - a2 b2 ; q2 = a2 b1 + a1 b2
        ga[i1] = g1 * b1 + g2 * b2;
        ga[i2] = -g1 * b2 + g2 * b1;
        gb[i1] = g1 * a1 + g2 * a2;
        gb[i2] = -g1 * a2 + g2 * a1;
      } else {
        // q1 = a1 b1 + a2 b2 ; q2 = a2 b1 - a1 b2
        ga[i1] = g1 * b1 - g2 * b2;
        ga[i2] = g1 * b2 + g2 * b1;
        gb[i1] = g1 * a1 + g2 * a2;
        gb[i2] = g1 * a2 - g2 * a1;
      }
    }
  }
}

int elementwise_grid(int64_t n, int threads) {
  int64_t blocks = ceil_div64(n, threads);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace okge

using namespace okge;

extern "C" int okge_gather_rows(const float* table, int64_t ld_table, const int32_t* ids, int64_t n,
                                int64_t D, float* out, int64_t ld_out, okge_stream_t stream) {
  if (n == 0) return OKGE_OK;
  OKGE_REQUIRE(table && ids && out, "null pointer");
  OKGE_REQUIRE(D > 0 && D % 4 == 0 && ld_table % 4 == 0 && ld_out % 4 == 0, "D and leading dimensions must be multiples of 4");
  OKGE_REQUIRE(aligned16(table) && aligned16(out), "table/out must be 16-byte aligned");
  OKGE_LAUNCH((gather_rows_kernel), rows_grid(n), kThreads, 0, static_cast<cudaStream_t>(stream), table, ld_table, ids, n, static_cast<int>(D / 4), out, ld_out);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_scatter_add_rows(const float* grad, int64_t ld_grad, const int32_t* ids, int64_t n,
                                     int64_t D, int32_t skip_id, float* grad_table, int64_t ld_table,
                                     okge_stream_t stream) {
  if (n == 0) return OKGE_OK;
  OKGE_REQUIRE(grad && ids && grad_table, "null pointer");
  OKGE_REQUIRE(D > 0 && D % 4 == 0 && ld_table % 4 == 0 && ld_grad % 4 == 0, "D and leading dimensions must be multiples of 4");
  OKGE_REQUIRE(aligned16(grad) && aligned16(grad_table), "grad/grad_table must be 16-byte aligned");
  OKGE_LAUNCH((scatter_add_rows_kernel), rows_grid(n), kThreads, 0, static_cast<cudaStream_t>(stream), grad, ld_grad, ids, n, static_cast<int>(D / 4), skip_id, grad_table, ld_table);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_gather_pool_fwd(const float* tok_table, int64_t ld_table, const int32_t* id_rows,
                                    int32_t L, const int32_t* ids, int64_t id_start, int64_t n,
                                    int64_t D, int32_t mode, float* out, int64_t ld_out,
                                    okge_stream_t stream) {
  if (n == 0) return OKGE_OK;
  OKGE_REQUIRE(tok_table && id_rows && out, "null pointer");
  OKGE_REQUIRE(L >= 1 && L <= kMaxL, "L must be in [1, 32]");
  OKGE_REQUIRE(D > 0 && D % 4 == 0 && ld_table % 4 == 0 && ld_out % 4 == 0, "D and leading dimensions must be multiples of 4");
  OKGE_REQUIRE(aligned16(tok_table) && aligned16(out), "tok_table/out must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int D4 = static_cast<int>(D / 4);
  const int grid = rows_grid(n);
  switch (mode) {
    case OKGE_POOL_SUM:
      OKGE_LAUNCH((gather_pool_fwd_kernel<OKGE_POOL_SUM>), grid, kThreads, 0, s, tok_table, ld_table, id_rows, L, ids, id_start, n, D4, out, ld_out);
      break;
    case OKGE_POOL_MEAN:
      OKGE_LAUNCH((gather_pool_fwd_kernel<OKGE_POOL_MEAN>), grid, kThreads, 0, s, tok_table, ld_table, id_rows, L, ids, id_start, n, D4, out, ld_out);
      break;
    case OKGE_POOL_MAX:
      OKGE_LAUNCH((gather_pool_fwd_kernel<OKGE_POOL_MAX>), grid, kThreads, 0, s, tok_table, ld_table, id_rows, L, ids, id_start, n, D4, out, ld_out);
      break;
    default:
      OKGE_REQUIRE(false, "unknown pooling mode");
  }
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

static int gather_pool_bwd_impl(const float* grad_out, int64_t ld_grad, const float* tok_table,
                                int64_t ld_table, const int32_t* id_rows, int32_t L,
                                const int32_t* ids, int64_t id_start, int64_t n, int64_t D,
                                int32_t mode, float* grad_tok_table, const int32_t* slot_map, okge_stream_t stream);

extern "C" int okge_gather_pool_bwd(const float* grad_out, int64_t ld_grad, const float* tok_table,
                                    int64_t ld_table, const int32_t* id_rows, int32_t L,
                                    const int32_t* ids, int64_t id_start, int64_t n, int64_t D,
                                    int32_t mode, float* grad_tok_table, okge_stream_t stream) {
  return gather_pool_bwd_impl(grad_out, ld_grad, tok_table, ld_table, id_rows, L, ids, id_start, n, D, mode, grad_tok_table,
                              nullptr, stream);
}

extern "C" int okge_gather_pool_bwd_slots(const float* grad_out, int64_t ld_grad, const float* tok_table,
                                          int64_t ld_table, const int32_t* id_rows, int32_t L,
                                          const int32_t* ids, int64_t id_start, int64_t n, int64_t D,
                                          int32_t mode, const int32_t* slot_map, float* slot_grad, okge_stream_t stream) {
  OKGE_REQUIRE(slot_map != nullptr, "null slot map");
  return gather_pool_bwd_impl(grad_out, ld_grad, tok_table, ld_table, id_rows, L, ids, id_start, n, D, mode, slot_grad,
                              slot_map, stream);
}

static int gather_pool_bwd_impl(const float* grad_out, int64_t ld_grad, const float* tok_table,
                                int64_t ld_table, const int32_t* id_rows, int32_t L,
                                const int32_t* ids, int64_t id_start, int64_t n, int64_t D,
                                int32_t mode, float* grad_tok_table, const int32_t* slot_map, okge_stream_t stream) {
  if (n == 0) return OKGE_OK;
  OKGE_REQUIRE(grad_out && id_rows && grad_tok_table, "null pointer");
  OKGE_REQUIRE(mode != OKGE_POOL_MAX || tok_table != nullptr, "max pooling backward needs tok_table");
  OKGE_REQUIRE(L >= 1 && L <= kMaxL, "L must be in [1, 32]");
  OKGE_REQUIRE(D > 0 && D % 4 == 0 && ld_table % 4 == 0 && ld_grad % 4 == 0, "D and leading dimensions must be multiples of 4");
  OKGE_REQUIRE(aligned16(grad_out) && aligned16(grad_tok_table), "grad_out/grad_tok_table must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int D4 = static_cast<int>(D / 4);
  // hot tokens: UNK=1, BOS=2, EOS=3 (openkge/index_mapper.py:14) accumulated per block in smem
  const int hot_lo = 1, hot_n = 3;
  const size_t smem = static_cast<size_t>(hot_n) * D4 * sizeof(float4);
  OKGE_REQUIRE(smem <= 48 * 1024, "D too large for the hot-token accumulator (max 1024)");
  // Every block flushes the privatised hot rows once, so large calls take fat blocks (16 rows per warp). A warp walks its
  // rows one after the other (L dependent rounds of reductions each): at batch sizes (512 ... 8,192 rows) that serial chain
  // IS the kernel's duration -- 36 us for 512 rows x 10 tokens x 64 columns on 4 blocks, 0.13 ms for 8,192 x 10 x 512 on
  // 64 blocks -- so small calls spread to at least ~4 blocks per SM first.
  int64_t rows_per_warp = n / (static_cast<int64_t>(sm_count()) * kWarpsPerBlock * 4);
  rows_per_warp = rows_per_warp < 1 ? 1 : (rows_per_warp > 16 ? 16 : rows_per_warp);
  int64_t blocks = ceil_div64(n, kWarpsPerBlock * rows_per_warp);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const int grid = static_cast<int>(blocks);
  switch (mode) {
    case OKGE_POOL_SUM:
      OKGE_LAUNCH((gather_pool_bwd_kernel<OKGE_POOL_SUM>), grid, kThreads, smem, s, grad_out, ld_grad, tok_table, ld_table, id_rows, L, ids, id_start, n, D4, grad_tok_table, hot_lo, hot_n, slot_map);
      break;
    case OKGE_POOL_MEAN:
      OKGE_LAUNCH((gather_pool_bwd_kernel<OKGE_POOL_MEAN>), grid, kThreads, smem, s, grad_out, ld_grad, tok_table, ld_table, id_rows, L, ids, id_start, n, D4, grad_tok_table, hot_lo, hot_n, slot_map);
      break;
    case OKGE_POOL_MAX:
      OKGE_LAUNCH((gather_pool_bwd_kernel<OKGE_POOL_MAX>), grid, kThreads, smem, s, grad_out, ld_grad, tok_table, ld_table, id_rows, L, ids, id_start, n, D4, grad_tok_table, hot_lo, hot_n, slot_map);
      break;
    default:
      OKGE_REQUIRE(false, "unknown pooling mode");
  }
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_dropout(const float* x, int64_t n, float p, uint64_t seed, uint64_t offset,
                            float* out, okge_stream_t stream) {
  if (n == 0) return OKGE_OK;
  OKGE_REQUIRE(x && out, "null pointer");
  OKGE_REQUIRE(p >= 0.f && p < 1.f, "dropout probability must be in [0, 1)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p == 0.f) {
    if (x != out) OKGE_CUDA_TRY(cudaMemcpyAsync(out, x, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return OKGE_OK;
  }
  OKGE_LAUNCH((dropout_kernel), elementwise_grid((n + 3) / 4, 256), 256, 0, s, x, n, p, 1.f / (1.f - p), seed, offset, nullptr, out);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_dropout_step(const float* x, int64_t n, float p, uint64_t seed, uint64_t offset, const uint64_t* step_dev,
                                 float* out, okge_stream_t stream) {
  if (n == 0) return OKGE_OK;
  OKGE_REQUIRE(x && out && step_dev, "null pointer");
  OKGE_REQUIRE(p > 0.f && p < 1.f, "dropout probability must be in (0, 1)");
  OKGE_LAUNCH((dropout_kernel), elementwise_grid((n + 3) / 4, 256), 256, 0, static_cast<cudaStream_t>(stream), x, n, p, 1.f / (1.f - p), seed, offset, reinterpret_cast<const unsigned long long*>(step_dev), out);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_fold_query_rows(const int32_t* kinds, const float* a, const float* b, int64_t Bq, int64_t D, float* q,
                                    okge_stream_t stream) {
  if (Bq == 0) return OKGE_OK;
  OKGE_REQUIRE(kinds && a && b && q, "null pointer");
  OKGE_REQUIRE(D > 0 && D % 2 == 0, "per-row kinds are the two ComplEx folds: D must be even");
  OKGE_LAUNCH((fold_query_kernel), elementwise_grid(Bq * D, 256), 256, 0, static_cast<cudaStream_t>(stream), OKGE_FOLD_COMPLEX_SP, kinds, a, b, Bq, static_cast<int>(D), q);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

// Row kinds and batch-norm segment bounds of a training batch whose po rows come first, from the device copy of the
// number of po rows -- the bookkeeping a graph-replayed step needs per batch, in one launch (see okge_b200.h).
__global__ void __launch_bounds__(256)
batch_layout_kernel(const int32_t* __restrict__ n_po_dev, const int32_t* __restrict__ count_dev, int rows, int n_cols,
                    int kind_po, int kind_sp, int32_t* __restrict__ kinds, int32_t* __restrict__ segments, int token_model,
                    unsigned long long* __restrict__ step_counter) {
  pdl_wait_and_trigger();
  const int b_po = __ldg(n_po_dev);
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (kinds != nullptr && i < rows) kinds[i] = i < b_po ? kind_po : kind_sp;
  if (i == 0) {
    if (segments != nullptr) {
      if (token_model) {
        // candidates (the real ones of a padded list), po block, sp block of the entity rows; po, sp block of the relation rows
        const int first = count_dev != nullptr ? __ldg(count_dev) : n_cols;
        const int seg[10] = {0, first, n_cols, n_cols + b_po, n_cols + b_po, n_cols + rows, 0, b_po, b_po, rows};
        for (int k = 0; k < 10; ++k) segments[k] = seg[k];
      } else {
        segments[0] = 0; segments[1] = b_po; segments[2] = b_po; segments[3] = rows;
      }
    }
    if (step_counter != nullptr) *step_counter += 1ull;
  }
}

extern "C" int okge_batch_layout(const int32_t* n_po_dev, const int32_t* count_dev, int64_t rows, int64_t n_cols, int32_t kind_po,
                                 int32_t kind_sp, int32_t* kinds, int32_t* segments, int32_t token_model,
                                 uint64_t* step_counter, okge_stream_t stream) {
  OKGE_REQUIRE(n_po_dev != nullptr, "null pointer");
  OKGE_REQUIRE(rows > 0 && rows < (int64_t(1) << 30) && n_cols >= 0 && n_cols < (int64_t(1) << 30), "bad shape");
  OKGE_LAUNCH((batch_layout_kernel), static_cast<unsigned>(kinds != nullptr ? ceil_div64(rows, 256) : 1), 256, 0,
              static_cast<cudaStream_t>(stream), n_po_dev, count_dev, static_cast<int>(rows), static_cast<int>(n_cols), kind_po,
              kind_sp, kinds, segments, token_model, reinterpret_cast<unsigned long long*>(step_counter));
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_fold_query_rows_bwd(const int32_t* kinds, const float* a, const float* b, const float* grad_q, int64_t Bq,
                                        int64_t D, float* grad_a, float* grad_b, okge_stream_t stream) {
  if (Bq == 0) return OKGE_OK;
  OKGE_REQUIRE(kinds && a && b && grad_q && grad_a && grad_b, "null pointer");
  OKGE_REQUIRE(D > 0 && D % 2 == 0, "per-row kinds are the two ComplEx folds: D must be even");
  OKGE_LAUNCH((fold_query_bwd_kernel), elementwise_grid(Bq * D, 256), 256, 0, static_cast<cudaStream_t>(stream), OKGE_FOLD_COMPLEX_SP, kinds, a, b, grad_q, Bq, static_cast<int>(D), grad_a, grad_b);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_fold_query(int32_t kind, const float* a, const float* b, int64_t Bq, int64_t D,
                               float* q, okge_stream_t stream) {
  if (Bq == 0) return OKGE_OK;
  OKGE_REQUIRE(a && b && q, "null pointer");
  OKGE_REQUIRE(kind >= OKGE_FOLD_COMPLEX_SP && kind <= OKGE_FOLD_DISTMULT, "unknown fold kind");
  OKGE_REQUIRE(kind == OKGE_FOLD_DISTMULT || D % 2 == 0, "ComplEx needs an even embedding width");
  OKGE_LAUNCH((fold_query_kernel), elementwise_grid(Bq * D, 256), 256, 0, static_cast<cudaStream_t>(stream), kind, nullptr, a, b, Bq, static_cast<int>(D), q);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_fold_query_bwd(int32_t kind, const float* a, const float* b, const float* grad_q,
                                   int64_t Bq, int64_t D, float* grad_a, float* grad_b,
                                   okge_stream_t stream) {
  if (Bq == 0) return OKGE_OK;
  OKGE_REQUIRE(a && b && grad_q && grad_a && grad_b, "null pointer");
  OKGE_REQUIRE(kind >= OKGE_FOLD_COMPLEX_SP && kind <= OKGE_FOLD_DISTMULT, "unknown fold kind");
  OKGE_REQUIRE(kind == OKGE_FOLD_DISTMULT || D % 2 == 0, "ComplEx needs an even embedding width");
  OKGE_LAUNCH((fold_query_bwd_kernel), elementwise_grid(Bq * D, 256), 256, 0, static_cast<cudaStream_t>(stream), kind, nullptr, a, b, grad_q, Bq, static_cast<int>(D), grad_a, grad_b);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}
