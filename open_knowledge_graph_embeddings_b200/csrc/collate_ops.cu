// Batch-shared collate of a training batch on the device.
//
// Reference: the collate function of OneToNMentionRelationDataset with use_batch_shared_entities (openkge/dataset.py:
// 813-868 builds the candidate list of the batch: the entities that occur as answers, topped up to
// min_size_batch_labels with negatives from numpy.random.choice; :885-932 writes the po rows first, then the sp rows, and
// the [B, n_candidates] label matrix). The reference does this with Python dicts per batch on the host; here a batch is
// B prefix-row indices that are already in HBM and the collate is six small launches with FIXED shapes (capacities
// instead of data-dependent sizes), so it can be captured in the same CUDA graph as the training step it feeds:
//
//   rows        1 CTA     stable partition po | sp, prefix ids, CSR row pointer of the labels (block scans)
//   mark        grid      every positive label: entity -> bit in a bitmap over all entities
//   tile sums   grid      popcount of every 1024-word tile of the bitmap
//   rank        grid      exclusive popcount prefix per word (= column of the word's first entity in the candidate list)
//                         + the candidate list of the positives, ascending by entity id
//   index+draw  grid      column of every label (prefix + popcount below its bit: ascending within a row because the
//                         index stores every answer list ascending); n_draw uniform entity draws (Philox), first arrival
//                         per entity by atomicMin
//   negatives   1 CTA     the first min_size - n_unique draws that are neither positives nor repeats, appended to the list;
//                         scalars (count, 1 / (B * count), overflow and label counters, draw counter)
//
// Same distribution as the host collate (dataset.collate_shared, the reference-exact one: same numpy stream and order),
// not the same stream: the positives are ordered by entity id instead of first occurrence -- the order of the columns
// does not enter the loss -- and the negatives come from Philox. Deterministic: no result depends on the order in which
// atomics land.
#include "okge_common.cuh"

#include <limits.h>

namespace okge {

namespace {

constexpr int kCta = 1024;          // single-CTA kernels
constexpr int kTileWords = 1024;    // bitmap words per tile (256 threads x uint4)

// exclusive block scan (kCta or fewer threads, whole warps); returns the exclusive prefix, `total` = block sum
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T& total, T* warp_sums /* [32] shared */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
  T incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const T up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  __syncthreads();                                   // warp_sums may still be read from the previous call
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    T w = lane < n_warps ? warp_sums[lane] : T(0);
    T wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const T up = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += up;
    }
    warp_sums[lane] = wi - w;                        // exclusive prefix of the warp sums
    if (lane == 31) warp_sums[32] = wi;              // block total
  }
  __syncthreads();
  total = warp_sums[32];
  return warp_sums[warp] + incl - v;
}

// scalars[]: see OKGE_COLLATE_* in okge_b200.h
__global__ void __launch_bounds__(kCta)
collate_rows_kernel(const int64_t* __restrict__ rows, int B, const int64_t* __restrict__ lab_ptr,
                    const int32_t* __restrict__ prefix, const int32_t* __restrict__ slot, int cap_nnz,
                    int64_t* __restrict__ row_start, int32_t* __restrict__ ent, int32_t* __restrict__ rel,
                    int32_t* __restrict__ is_po, int32_t* __restrict__ ptr, int64_t* __restrict__ scalars) {
  pdl_wait_and_trigger();
  __shared__ int64_t sums[33];
  int64_t total;
  int64_t mine = 0;
  for (int i = threadIdx.x; i < B; i += kCta) mine += (slot[rows[i]] == 0);
  block_exclusive_scan<int64_t>(mine, total, sums);
  const int b_po = static_cast<int>(total);
  // stable partition: po rows keep their order in [0, b_po), sp rows in [b_po, B)
  int run_po = 0;
  for (int base = 0; base < B; base += kCta) {
    const int i = base + threadIdx.x;
    int64_t r = 0;
    int flag = 0;
    if (i < B) {
      r = rows[i];
      flag = slot[r] == 0;
    }
    const int64_t excl = block_exclusive_scan<int64_t>(flag, total, sums);
    if (i < B) {
      const int po_before = run_po + static_cast<int>(excl);
      const int pos = flag ? po_before : b_po + (i - po_before);
      const int64_t s = lab_ptr[r], len = lab_ptr[r + 1] - s;
      const int32_t a = prefix[2 * r], b = prefix[2 * r + 1];
      row_start[pos] = s;
      ptr[pos + 1] = static_cast<int32_t>(min(len, static_cast<int64_t>(INT_MAX)));
      ent[pos] = flag ? b : a;                       // po prefix = (rel, obj), sp prefix = (subj, rel)
      rel[pos] = flag ? a : b;
      is_po[pos] = flag;
    }
    run_po += static_cast<int>(total);
  }
  __syncthreads();
  // CSR row pointer: inclusive sum of the row lengths in the new order, clamped to the label capacity (a batch with more
  // positives than the buffer holds keeps the labels that fit)
  int64_t carry = 0;
  for (int base = 0; base < B; base += kCta) {
    const int p = base + threadIdx.x;
    const int64_t len = p < B ? ptr[p + 1] : 0;
    const int64_t excl = block_exclusive_scan<int64_t>(len, total, sums);
    if (p < B) ptr[p + 1] = static_cast<int32_t>(min(carry + excl + len, static_cast<int64_t>(cap_nnz)));
    carry += total;
  }
  if (threadIdx.x == 0) {
    ptr[0] = 0;
    scalars[OKGE_COLLATE_B_PO] = b_po;
    scalars[OKGE_COLLATE_NNZ] = min(carry, static_cast<int64_t>(cap_nnz));
    scalars[OKGE_COLLATE_LABELS_ALL] = carry;
  }
}

__global__ void __launch_bounds__(256)
collate_mark_kernel(const int32_t* __restrict__ ptr, int B, const int64_t* __restrict__ row_start,
                    const int32_t* __restrict__ lab_idx, const int64_t* __restrict__ scalars, int cap_nnz,
                    int32_t* __restrict__ e_flat, int32_t* __restrict__ idx, uint32_t* __restrict__ bitmap) {
  pdl_wait_and_trigger();
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= cap_nnz) return;
  if (j >= static_cast<int>(scalars[OKGE_COLLATE_NNZ])) {
    idx[j] = -1;
    return;
  }
  int lo = 0, hi = B;                                // the row with ptr[row] <= j < ptr[row + 1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (ptr[mid] <= j) lo = mid; else hi = mid;
  }
  const int32_t e = lab_idx[row_start[lo] + (j - ptr[lo])];
  e_flat[j] = e;
  atomicOr(bitmap + (e >> 5), 1u << (e & 31));
}

__global__ void __launch_bounds__(256)
collate_tile_sums_kernel(const uint32_t* __restrict__ bitmap, int n_words, int32_t* __restrict__ tile_sum) {
  pdl_wait_and_trigger();
  __shared__ int32_t sums[33];
  const int w = blockIdx.x * kTileWords + 4 * threadIdx.x;
  int32_t c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) c += (w + k < n_words) ? __popc(bitmap[w + k]) : 0;
  int32_t total;
  block_exclusive_scan<int32_t>(c, total, sums);
  if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256)
collate_rank_kernel(const uint32_t* __restrict__ bitmap, int n_words, const int32_t* __restrict__ tile_sum,
                    int32_t id_offset, int cap_cols, int32_t* __restrict__ word_prefix, int32_t* __restrict__ cand,
                    int64_t* __restrict__ scalars) {
  pdl_wait_and_trigger();
  __shared__ int32_t sums[33];
  int32_t before = 0, total;
  for (int t = threadIdx.x; t < static_cast<int>(blockIdx.x); t += 256) before += tile_sum[t];
  block_exclusive_scan<int32_t>(before, total, sums);
  const int32_t tile_base = total;
  const int w = blockIdx.x * kTileWords + 4 * threadIdx.x;
  uint32_t bits[4];
  int32_t c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    bits[k] = (w + k < n_words) ? bitmap[w + k] : 0u;
    c += __popc(bits[k]);
  }
  int32_t col = tile_base + block_exclusive_scan<int32_t>(c, total, sums);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (w + k < n_words) word_prefix[w + k] = col;
    uint32_t m = bits[k];
    while (m) {
      const int bit = __ffs(m) - 1;
      m &= m - 1;
      if (col < cap_cols) cand[col] = (w + k) * 32 + bit + id_offset;
      ++col;
    }
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) scalars[OKGE_COLLATE_N_UNIQUE] = tile_base + total;
}

__device__ __forceinline__ int32_t draw_entity(uint64_t seed, uint64_t call, uint32_t i, int64_t n_entities) {
  const uint4 r = philox4x32_10(make_uint4(i, 0x636f6c6cu, static_cast<uint32_t>(call), static_cast<uint32_t>(call >> 32)),
                                make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  const uint64_t u = (static_cast<uint64_t>(r.x) << 32) | r.y;
  return static_cast<int32_t>(__umul64hi(u, static_cast<uint64_t>(n_entities)));     // uniform in [0, n), bias < 2^-40
}

__global__ void __launch_bounds__(256)
collate_index_kernel(const int32_t* __restrict__ e_flat, const uint32_t* __restrict__ bitmap,
                     const int32_t* __restrict__ word_prefix, const int64_t* __restrict__ scalars, int cap_nnz,
                     int n_draw, int64_t n_entities, uint64_t seed, int32_t* __restrict__ idx,
                     int32_t* __restrict__ draws, int32_t* __restrict__ first_draw) {
  pdl_wait_and_trigger();
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j < cap_nnz) {
    if (j < static_cast<int>(scalars[OKGE_COLLATE_NNZ])) {
      const int32_t e = e_flat[j];
      idx[j] = word_prefix[e >> 5] + __popc(bitmap[e >> 5] & ((1u << (e & 31)) - 1u));
    }
    return;
  }
  const int i = j - cap_nnz;
  if (i >= n_draw) return;
  const int32_t e = draw_entity(seed, static_cast<uint64_t>(scalars[OKGE_COLLATE_CALLS]), static_cast<uint32_t>(i), n_entities);
  draws[i] = e;
  if (!((bitmap[e >> 5] >> (e & 31)) & 1u)) atomicMin(first_draw + e, i);
}

__global__ void __launch_bounds__(kCta)
collate_negatives_kernel(const int32_t* __restrict__ draws, int n_draw, const uint32_t* __restrict__ bitmap,
                         int32_t* __restrict__ first_draw, int B, int min_size, int cap_cols, int64_t n_entities,
                         int32_t id_offset, int32_t* __restrict__ cand, int64_t* __restrict__ scalars,
                         int32_t* __restrict__ count_out, float* __restrict__ inv_norm) {
  pdl_wait_and_trigger();
  __shared__ int32_t sums[33];
  const int n_u_all = static_cast<int>(scalars[OKGE_COLLATE_N_UNIQUE]);
  const int n_u = min(n_u_all, cap_cols);
  const int need = max(min(min_size, cap_cols) - n_u, 0);
  int taken = 0;
  for (int base = 0; base < n_draw; base += kCta) {
    const int i = base + threadIdx.x;
    int32_t e = 0;
    int ok = 0;
    if (i < n_draw) {
      e = draws[i];
      ok = !((bitmap[e >> 5] >> (e & 31)) & 1u) && first_draw[e] == i;
    }
    int32_t total;
    const int k = taken + block_exclusive_scan<int32_t>(ok, total, sums);
    if (ok && k < need) cand[n_u + k] = e + id_offset;
    taken += total;
    if (taken >= need) break;                        // block-uniform: the list is full (usually after the first 1,024 draws)
  }
  taken = min(taken, need);
  const int count = n_u + taken;
  __syncthreads();                                   // every first_draw read above precedes the resets below
  for (int i = threadIdx.x; i < n_draw; i += kCta) first_draw[draws[i]] = INT_MAX;
  for (int c = count + threadIdx.x; c < cap_cols; c += kCta) cand[c] = id_offset;      // padding: a valid entity id
  if (threadIdx.x == 0) {
    const int64_t want = min(static_cast<int64_t>(min(min_size, cap_cols)), n_entities);
    scalars[OKGE_COLLATE_COUNT] = count;
    scalars[OKGE_COLLATE_OVERFLOW] += (scalars[OKGE_COLLATE_LABELS_ALL] > scalars[OKGE_COLLATE_NNZ]) || (n_u_all > cap_cols) ||
                                      (count < want);
    scalars[OKGE_COLLATE_NNZ_TOTAL] += scalars[OKGE_COLLATE_NNZ];
    scalars[OKGE_COLLATE_CALLS] += 1;
    if (count_out != nullptr) count_out[0] = count;
    if (inv_norm != nullptr) inv_norm[0] = 1.0f / (static_cast<float>(B) * static_cast<float>(max(count, 1)));
  }
}

}  // namespace

}  // namespace okge

using namespace okge;

extern "C" int okge_collate_shared(const int64_t* rows, int64_t n_rows, const int64_t* lab_ptr, const int32_t* lab_idx,
                                   const int32_t* prefix, const int32_t* slot, int64_t n_entities, int32_t id_offset,
                                   int64_t min_size, int64_t cap_nnz, int64_t cap_cols, int64_t n_draw, uint64_t seed,
                                   uint32_t* bitmap, int32_t* word_prefix, int32_t* tile_sum, int32_t* first_draw,
                                   int32_t* e_flat, int64_t* row_start, int32_t* ent, int32_t* rel, int32_t* is_po,
                                   int32_t* ptr, int32_t* idx, int32_t* cand, int64_t* scalars, int32_t* count_out,
                                   float* inv_norm, okge_stream_t stream) {
  OKGE_REQUIRE(rows && lab_ptr && lab_idx && prefix && slot && bitmap && word_prefix && tile_sum && first_draw && e_flat &&
                   row_start && ent && rel && is_po && ptr && idx && cand && scalars,
               "null pointer");
  OKGE_REQUIRE(n_rows > 0 && n_rows < (1 << 24), "between 1 and 2^24 - 1 prefix rows");
  OKGE_REQUIRE(n_entities > 0 && n_entities < (int64_t(1) << 31) - 64, "entity ids must fit int32");
  OKGE_REQUIRE(cap_nnz > 0 && cap_cols > 0 && n_draw >= 0 && min_size >= 0 && cap_nnz + n_draw < (int64_t(1) << 30) &&
                   cap_cols < (int64_t(1) << 30) && min_size < (int64_t(1) << 30),
               "bad capacities");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int B = static_cast<int>(n_rows), C = static_cast<int>(cap_nnz), S = static_cast<int>(n_draw);
  const int n_words = static_cast<int>((n_entities + 31) / 32);
  const int n_tiles = (n_words + kTileWords - 1) / kTileWords;
  OKGE_CUDA_TRY(cudaMemsetAsync(bitmap, 0, sizeof(uint32_t) * static_cast<size_t>(n_words), s));
  OKGE_LAUNCH((collate_rows_kernel), 1, kCta, 0, s, rows, B, lab_ptr, prefix, slot, C, row_start, ent, rel, is_po, ptr, scalars);
  OKGE_LAUNCH((collate_mark_kernel), (C + 255) / 256, 256, 0, s, ptr, B, row_start, lab_idx, scalars, C, e_flat, idx, bitmap);
  OKGE_LAUNCH((collate_tile_sums_kernel), n_tiles, 256, 0, s, bitmap, n_words, tile_sum);
  OKGE_LAUNCH((collate_rank_kernel), n_tiles, 256, 0, s, bitmap, n_words, tile_sum, id_offset, static_cast<int>(cap_cols), word_prefix,
                                              cand, scalars);
  OKGE_LAUNCH((collate_index_kernel), (C + S + 255) / 256, 256, 0, s, e_flat, bitmap, word_prefix, scalars, C, S, n_entities, seed, idx,
                                                           e_flat + C, first_draw);
  OKGE_LAUNCH((collate_negatives_kernel), 1, kCta, 0, s, e_flat + C, S, bitmap, first_draw, B, static_cast<int>(min_size),
                                              static_cast<int>(cap_cols), n_entities, id_offset, cand, scalars, count_out,
                                              inv_norm);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}
