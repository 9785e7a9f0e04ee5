// Batch normalisation over the rows of an [n, D] operand, for the token-based embedders of the OpenKGE hot path
// (torch.nn.BatchNorm1d applied to the pooled vectors, openkge/model.py:597-612, 777-780; Lookup embedders :463-465).
//
// Why its own kernels: the reference normalises each encode call separately (all candidates, po relations, po objects,
// sp subjects, sp relations - openkge/trainer.py:69-87), so the statistics of one batch are taken over row SEGMENTS whose
// bounds (the po / sp split) change from batch to batch. Here the bounds are device data (`seg`: one [begin, end) pair
// per segment; segments may leave gaps, e.g. the padding of a fixed-capacity candidate list), so the launches of a
// step have fixed shapes (CUDA-graph replay) and the two query blocks of a table are normalised by one launch sequence;
// running statistics are updated segment by segment in the reference's call order.
//
// Layout: HBM-bound streaming. A block covers a 128-column tile (lane -> one float4) x a chunk of a segment's rows,
// warp w of the block walks rows w, w+8, ...: every access is a coalesced 512-byte warp transaction. Column sums are
// accumulated in fp64 (mean / biased variance to fp32 round-off without a Welford chain), merged through shared memory
// and a [segments, chunks, D, 2] workspace; a one-thread-per-column kernel finishes the statistics.
#include "okge_common.cuh"

#include <math.h>

namespace okge {

namespace {

constexpr int kBnWarps = 8;
constexpr int kBnThreads = kBnWarps * 32;
constexpr int kBnTileCols = 128;
constexpr int kBnMaxSegments = 8;

struct BnGrid {
  int col_tiles;
  int chunks;
};

// Inverted dropout fused behind the normalisation (F.dropout of the embedders, openkge/model.py:783-786): y <- mask * y /
// (1 - p) in the forward, dy <- mask * dy / (1 - p) on the way in in the backward. Same Philox stream as okge_dropout on
// the flattened [n, D] output (element (r, c) has index r * D + c), so the mask never exists in memory. p == 0: off.
struct BnDropout {
  float p, scale;
  uint64_t seed, offset;
  const unsigned long long* step_dev;   // nullable: per-step stream position of a replayed (CUDA graph) launch
};

__device__ __forceinline__ uint64_t drop_offset(const BnDropout& d) {
  return d.step_dev != nullptr ? d.offset + (static_cast<uint64_t>(*d.step_dev) << 44) : d.offset;
}

BnGrid bn_grid(int64_t n_rows, int D) {
  BnGrid g;
  g.col_tiles = (D + kBnTileCols - 1) / kBnTileCols;
  // >= 16 rows per warp: the finalize kernels add up `chunks` partials per column on D / 32 blocks only, and with 4 rows
  // per warp a narrow operand (15,053 x 64: 592 chunks, 2 finalize blocks) spent 26 - 36 us there
  int64_t chunks = ceil_div64(n_rows, 16 * kBnWarps);
  const int64_t cap = (static_cast<int64_t>(sm_count()) * 4 + g.col_tiles - 1) / g.col_tiles;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  g.chunks = static_cast<int>(chunks);
  return g;
}

// rows [lo, hi) of segment z handled by chunk blockIdx.y
__device__ __forceinline__ void chunk_rows(const int32_t* seg, int64_t n_rows, int64_t& lo, int64_t& hi, int64_t& n_seg_rows) {
  int64_t s0 = 0, s1 = n_rows;
  if (seg != nullptr) {
    s0 = seg[2 * blockIdx.z];
    s1 = seg[2 * blockIdx.z + 1];
  }
  n_seg_rows = s1 - s0;
  const int64_t per = (n_seg_rows + gridDim.y - 1) / gridDim.y;
  lo = s0 + per * blockIdx.y;
  hi = lo + per < s1 ? lo + per : s1;
}

// Partial column sums over a chunk: (sum a, sum a*b') where, for the forward statistics, a = x and the second sum is x^2;
// for the backward, a = dy and the second sum is dy * xhat.
template <bool BWD>
__global__ void __launch_bounds__(kBnThreads)
bn_partial_kernel(const float* __restrict__ a, int64_t ld_a, const float* __restrict__ x, int64_t ld_x,
                  const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                  const int32_t* __restrict__ seg, int64_t n_rows, int D, double* __restrict__ partial, BnDropout drop) {
  pdl_wait_and_trigger();
  __shared__ double red[kBnWarps][32][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * kBnTileCols + lane * 4;
  int64_t lo, hi, n_seg_rows;
  chunk_rows(seg, n_rows, lo, hi, n_seg_rows);
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (col < D) {
    float4 mu = make_float4(0, 0, 0, 0), is = make_float4(0, 0, 0, 0);
    if (BWD) {
      mu = *reinterpret_cast<const float4*>(save_mean + static_cast<int64_t>(blockIdx.z) * D + col);
      is = *reinterpret_cast<const float4*>(save_invstd + static_cast<int64_t>(blockIdx.z) * D + col);
    }
    const uint64_t doff = (BWD && drop.p > 0.f) ? drop_offset(drop) : 0;
    for (int64_t r = lo + warp; r < hi; r += kBnWarps) {
      float4 v = __ldg(reinterpret_cast<const float4*>(a + r * ld_a + col));
      if (BWD) {
        if (drop.p > 0.f) v = dropout4(v, static_cast<uint64_t>(r * D + col) >> 2, drop.p, drop.scale, drop.seed, doff);
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + r * ld_x + col));
        s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
        q[0] += static_cast<double>(v.x) * ((xv.x - mu.x) * is.x);
        q[1] += static_cast<double>(v.y) * ((xv.y - mu.y) * is.y);
        q[2] += static_cast<double>(v.z) * ((xv.z - mu.z) * is.z);
        q[3] += static_cast<double>(v.w) * ((xv.w - mu.w) * is.w);
      } else {
        s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
        q[0] += static_cast<double>(v.x) * v.x; q[1] += static_cast<double>(v.y) * v.y;
        q[2] += static_cast<double>(v.z) * v.z; q[3] += static_cast<double>(v.w) * v.w;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[warp][lane][j] = s[j];
    red[warp][lane][4 + j] = q[j];
  }
  __syncthreads();
  // thread t < 256 owns (lane = t / 8, slot = t % 8) and adds the 8 warps in a fixed order (deterministic)
  const int l = threadIdx.x >> 3, slot = threadIdx.x & 7;
  double acc = 0;
#pragma unroll
  for (int w = 0; w < kBnWarps; ++w) acc += red[w][l][slot];
  const int c = blockIdx.x * kBnTileCols + l * 4 + (slot & 3);
  if (c < D) {
    const int64_t base = ((static_cast<int64_t>(blockIdx.z) * gridDim.y + blockIdx.y) * D + c) * 2;
    partial[base + (slot >> 2)] = acc;
  }
}

// Finalize kernels: a block owns 32 columns; 16 thread rows split the chunk partials of a column between them and are
// merged through shared memory in a fixed order. Thread row 0 then holds the column's (sum, second sum).
constexpr int kFinCols = 32;
constexpr int kFinLanes = 16;

__device__ __forceinline__ void reduce_partials(const double* __restrict__ partial, int z, int chunks, int D, int c,
                                                double (*red)[kFinCols][2], double& s, double& q) {
  const int kl = threadIdx.y;
  s = 0;
  q = 0;
  if (c < D) {
    // (independent loads in flight: the loop is bound by their latency, not by the additions)
    double s1 = 0, q1 = 0, s2 = 0, q2 = 0, s3 = 0, q3 = 0;
    const int64_t stride = static_cast<int64_t>(kFinLanes) * D * 2;
    const double* pp = partial + ((static_cast<int64_t>(z) * chunks + kl) * D + c) * 2;
    int k = kl;
    for (; k + 3 * kFinLanes < chunks; k += 4 * kFinLanes, pp += 4 * stride) {
      const double a0 = pp[0], b0 = pp[1], a1 = pp[stride], b1 = pp[stride + 1];
      const double a2 = pp[2 * stride], b2 = pp[2 * stride + 1], a3 = pp[3 * stride], b3 = pp[3 * stride + 1];
      s += a0; q += b0; s1 += a1; q1 += b1; s2 += a2; q2 += b2; s3 += a3; q3 += b3;
    }
    for (; k < chunks; k += kFinLanes, pp += stride) {
      s += pp[0];
      q += pp[1];
    }
    s += (s1 + s2) + s3;
    q += (q1 + q2) + q3;
  }
  __syncthreads();                               // the previous segment's readers are done with `red`
  red[kl][threadIdx.x][0] = s;
  red[kl][threadIdx.x][1] = q;
  __syncthreads();
  if (kl == 0) {
    s = 0;
    q = 0;
#pragma unroll
    for (int w = 0; w < kFinLanes; ++w) {
      s += red[w][threadIdx.x][0];
      q += red[w][threadIdx.x][1];
    }
  }
}

// Segments in order (the order of the reference's encode calls), so the running statistics see the same sequence of
// momentum updates.
__global__ void __launch_bounds__(kFinCols * kFinLanes)
bn_stats_finalize_kernel(const double* __restrict__ partial, const int32_t* __restrict__ seg, int n_seg,
                         int64_t n_rows, int chunks, int D, float momentum, float eps,
                         float* __restrict__ running_mean, float* __restrict__ running_var,
                         int64_t* __restrict__ num_batches_tracked,
                         float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  pdl_wait_and_trigger();
  __shared__ double red[kFinLanes][kFinCols][2];
  const int c = blockIdx.x * kFinCols + threadIdx.x;
  const bool owner = threadIdx.y == 0 && c < D;
  int nonempty = 0;
  for (int z = 0; z < n_seg; ++z) {
    const int64_t n = seg ? static_cast<int64_t>(seg[2 * z + 1]) - seg[2 * z] : n_rows;    // uniform across the block
    if (n <= 0) {
      if (owner) {
        save_mean[static_cast<int64_t>(z) * D + c] = 0.f;
        save_invstd[static_cast<int64_t>(z) * D + c] = 0.f;
      }
      continue;
    }
    ++nonempty;
    double s, q;
    reduce_partials(partial, z, chunks, D, c, red, s, q);
    if (!owner) continue;
    const double mean = s / static_cast<double>(n);
    double var = q / static_cast<double>(n) - mean * mean;          // biased, like the normalisation itself uses
    if (var < 0) var = 0;
    const float varf = static_cast<float>(var);
    save_mean[static_cast<int64_t>(z) * D + c] = static_cast<float>(mean);
    save_invstd[static_cast<int64_t>(z) * D + c] = 1.0f / sqrtf(varf + eps);
    if (running_mean != nullptr) {
      const float unbiased = static_cast<float>(var * (static_cast<double>(n) / static_cast<double>(n > 1 ? n - 1 : 1)));
      running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
      running_var[c] = (1.0f - momentum) * running_var[c] + momentum * unbiased;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0 && num_batches_tracked != nullptr)
    *num_batches_tracked += nonempty;
}

// y = (x - mean) * invstd * gamma + beta. EVAL: mean / invstd come from the running statistics.
template <bool EVAL>
__global__ void __launch_bounds__(kBnThreads)
bn_apply_kernel(const float* __restrict__ x, int64_t ld_x, const int32_t* __restrict__ seg, int64_t n_rows, int D,
                const float* __restrict__ mean, const float* __restrict__ invstd_or_var, float eps,
                const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y, int64_t ld_y,
                BnDropout drop) {
  pdl_wait_and_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * kBnTileCols + lane * 4;
  if (col >= D) return;
  int64_t lo, hi, n_seg_rows;
  chunk_rows(seg, n_rows, lo, hi, n_seg_rows);
  const uint64_t doff = drop.p > 0.f ? drop_offset(drop) : 0;
  const int64_t stat = EVAL ? col : static_cast<int64_t>(blockIdx.z) * D + col;
  const float4 mu = *reinterpret_cast<const float4*>(mean + stat);
  float4 is = *reinterpret_cast<const float4*>(invstd_or_var + stat);
  if (EVAL) {
    is.x = 1.0f / sqrtf(is.x + eps); is.y = 1.0f / sqrtf(is.y + eps);
    is.z = 1.0f / sqrtf(is.z + eps); is.w = 1.0f / sqrtf(is.w + eps);
  }
  const float4 g = gamma ? *reinterpret_cast<const float4*>(gamma + col) : make_float4(1, 1, 1, 1);
  const float4 b = beta ? *reinterpret_cast<const float4*>(beta + col) : make_float4(0, 0, 0, 0);
  for (int64_t r = lo + warp; r < hi; r += kBnWarps) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ld_x + col));
    float4 o;
    o.x = (v.x - mu.x) * is.x * g.x + b.x;
    o.y = (v.y - mu.y) * is.y * g.y + b.y;
    o.z = (v.z - mu.z) * is.z * g.z + b.z;
    o.w = (v.w - mu.w) * is.w * g.w + b.w;
    if (drop.p > 0.f) o = dropout4(o, static_cast<uint64_t>(r * D + col) >> 2, drop.p, drop.scale, drop.seed, doff);
    *reinterpret_cast<float4*>(y + r * ld_y + col) = o;
  }
}

// Backward statistics: per segment c1 = sum(dy) / n, c2 = sum(dy * xhat) / n (kept in `coef`), and the parameter
// gradients dbeta = sum over segments of sum(dy), dgamma = ... of sum(dy * xhat).
__global__ void __launch_bounds__(kFinCols * kFinLanes)
bn_bwd_finalize_kernel(const double* __restrict__ partial, const int32_t* __restrict__ seg, int n_seg,
                       int64_t n_rows, int chunks, int D, float* __restrict__ coef,
                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_wait_and_trigger();
  __shared__ double red[kFinLanes][kFinCols][2];
  const int c = blockIdx.x * kFinCols + threadIdx.x;
  const bool owner = threadIdx.y == 0 && c < D;
  double dg = 0, db = 0;
  for (int z = 0; z < n_seg; ++z) {
    const int64_t n = seg ? static_cast<int64_t>(seg[2 * z + 1]) - seg[2 * z] : n_rows;
    double s = 0, q = 0;
    if (n > 0) reduce_partials(partial, z, chunks, D, c, red, s, q);
    if (!owner) continue;
    db += s;
    dg += q;
    coef[(static_cast<int64_t>(z) * 2) * D + c] = n > 0 ? static_cast<float>(s / static_cast<double>(n)) : 0.f;
    coef[(static_cast<int64_t>(z) * 2 + 1) * D + c] = n > 0 ? static_cast<float>(q / static_cast<double>(n)) : 0.f;
  }
  if (owner && dgamma) dgamma[c] = static_cast<float>(dg);
  if (owner && dbeta) dbeta[c] = static_cast<float>(db);
}

// dx = gamma * invstd * (dy - c1 - xhat * c2)
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_apply_kernel(const float* __restrict__ dy, int64_t ld_dy, const float* __restrict__ x, int64_t ld_x,
                    const int32_t* __restrict__ seg, int64_t n_rows, int D, const float* __restrict__ save_mean,
                    const float* __restrict__ save_invstd, const float* __restrict__ coef,
                    const float* __restrict__ gamma, float* __restrict__ dx, int64_t ld_dx, BnDropout drop) {
  pdl_wait_and_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * kBnTileCols + lane * 4;
  if (col >= D) return;
  int64_t lo, hi, n_seg_rows;
  chunk_rows(seg, n_rows, lo, hi, n_seg_rows);
  const int64_t z = blockIdx.z;
  const float4 mu = *reinterpret_cast<const float4*>(save_mean + z * D + col);
  const float4 is = *reinterpret_cast<const float4*>(save_invstd + z * D + col);
  const float4 c1 = *reinterpret_cast<const float4*>(coef + (z * 2) * D + col);
  const float4 c2 = *reinterpret_cast<const float4*>(coef + (z * 2 + 1) * D + col);
  const float4 g = gamma ? *reinterpret_cast<const float4*>(gamma + col) : make_float4(1, 1, 1, 1);
  const uint64_t doff = drop.p > 0.f ? drop_offset(drop) : 0;
  for (int64_t r = lo + warp; r < hi; r += kBnWarps) {
    float4 d = __ldg(reinterpret_cast<const float4*>(dy + r * ld_dy + col));
    if (drop.p > 0.f) d = dropout4(d, static_cast<uint64_t>(r * D + col) >> 2, drop.p, drop.scale, drop.seed, doff);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ld_x + col));
    float4 o;
    o.x = g.x * is.x * (d.x - c1.x - (v.x - mu.x) * is.x * c2.x);
    o.y = g.y * is.y * (d.y - c1.y - (v.y - mu.y) * is.y * c2.y);
    o.z = g.z * is.z * (d.z - c1.z - (v.z - mu.z) * is.z * c2.z);
    o.w = g.w * is.w * (d.w - c1.w - (v.w - mu.w) * is.w * c2.w);
    *reinterpret_cast<float4*>(dx + r * ld_dx + col) = o;
  }
}

// Column sums of one segment out of the chunk partials: sums[c] = first sum, sums[D + c] = second sum (fp64). The split
// entry points (okge_bn_col_sums / okge_bn_normalize*) expose the three phases separately so that a caller whose rows are
// partitioned over several GPUs can all-reduce the sums between them (synchronised batch norm).
__global__ void __launch_bounds__(kFinCols * kFinLanes)
bn_sums_kernel(const double* __restrict__ partial, int chunks, int D, double* __restrict__ sums) {
  pdl_wait_and_trigger();
  __shared__ double red[kFinLanes][kFinCols][2];
  const int c = blockIdx.x * kFinCols + threadIdx.x;
  double s, q;
  reduce_partials(partial, 0, chunks, D, c, red, s, q);
  if (threadIdx.y == 0 && c < D) {
    sums[c] = s;
    sums[D + c] = q;
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

BnDropout make_drop(float p, uint64_t seed, uint64_t offset, const uint64_t* step_dev) {
  BnDropout d;
  d.p = p;
  d.scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  d.seed = seed;
  d.offset = offset;
  d.step_dev = reinterpret_cast<const unsigned long long*>(step_dev);
  return d;
}

int check_common(const void* x, int64_t ld_x, int n_seg, int64_t n_rows, int D) {
  OKGE_REQUIRE(x != nullptr, "null pointer");
  OKGE_REQUIRE(D > 0 && D % 4 == 0 && ld_x % 4 == 0, "D and leading dimensions must be multiples of 4");
  OKGE_REQUIRE(aligned16(x), "operands must be 16-byte aligned");
  OKGE_REQUIRE(n_seg >= 1 && n_seg <= kBnMaxSegments, "1..8 row segments");
  OKGE_REQUIRE(n_rows >= 0, "negative row count");
  return OKGE_OK;
}

}  // namespace

}  // namespace okge

using namespace okge;

extern "C" int64_t okge_bn_workspace_bytes(int64_t n_rows, int D, int n_seg) {
  if (n_rows < 0 || D <= 0 || n_seg < 1) return 0;
  const BnGrid g = bn_grid(n_rows, D);
  // partial sums (fp64 pairs) + the backward's per-segment coefficients (fp32 pairs)
  return static_cast<int64_t>(n_seg) * g.chunks * D * 2 * static_cast<int64_t>(sizeof(double)) +
         static_cast<int64_t>(n_seg) * 2 * D * static_cast<int64_t>(sizeof(float));
}

extern "C" int okge_bn_train_fwd(const float* x, int64_t ld_x, const int32_t* seg, int n_seg, int64_t n_rows, int D,
                                 const float* gamma, const float* beta, float* running_mean, float* running_var,
                                 int64_t* num_batches_tracked, float momentum, float eps, float* y, int64_t ld_y,
                                 float* save_mean, float* save_invstd, float drop_p, uint64_t drop_seed,
                                 uint64_t drop_offset_, const uint64_t* drop_step_dev, void* workspace, void* stream) {
  if (int rc = check_common(x, ld_x, n_seg, n_rows, D)) return rc;
  OKGE_REQUIRE(y && save_mean && save_invstd && workspace, "null pointer");
  OKGE_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || ld_y == D), "fused dropout: p in [0, 1), contiguous output");
  const BnDropout drop = make_drop(drop_p, drop_seed, drop_offset_, drop_step_dev), no_drop = make_drop(0.f, 0, 0, nullptr);
  OKGE_REQUIRE(ld_y % 4 == 0 && aligned16(y) && aligned16(save_mean) && aligned16(save_invstd), "outputs must be 16-byte aligned");
  OKGE_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "running_mean and running_var go together");
  OKGE_REQUIRE(seg != nullptr || n_seg == 1, "several segments need their bounds");
  if (int rc = okge_device_check()) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const BnGrid g = bn_grid(n_rows, D);
  double* partial = static_cast<double*>(workspace);
  const dim3 grid(g.col_tiles, g.chunks, n_seg);
  OKGE_LAUNCH((bn_partial_kernel<false>), grid, kBnThreads, 0, s, x, ld_x, nullptr, 0, nullptr, nullptr, seg, n_rows, D, partial, no_drop);
  OKGE_LAUNCH((bn_stats_finalize_kernel), (D + kFinCols - 1) / kFinCols, dim3(kFinCols, kFinLanes), 0, s, partial, seg, n_seg, n_rows, g.chunks, D, momentum, eps,
                                                           running_mean, running_var, num_batches_tracked, save_mean, save_invstd);
  if (n_rows > 0)
    OKGE_LAUNCH((bn_apply_kernel<false>), grid, kBnThreads, 0, s, x, ld_x, seg, n_rows, D, save_mean, save_invstd, eps, gamma, beta, y, ld_y, drop);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_bn_train_bwd(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x, const int32_t* seg, int n_seg,
                                 int64_t n_rows, int D, const float* gamma, const float* save_mean, const float* save_invstd,
                                 float* dx, int64_t ld_dx, float* dgamma, float* dbeta, float drop_p, uint64_t drop_seed,
                                 uint64_t drop_offset_, const uint64_t* drop_step_dev, void* workspace, void* stream) {
  if (int rc = check_common(x, ld_x, n_seg, n_rows, D)) return rc;
  OKGE_REQUIRE(dy && save_mean && save_invstd && workspace, "null pointer");
  OKGE_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "fused dropout: p in [0, 1)");
  const BnDropout drop = make_drop(drop_p, drop_seed, drop_offset_, drop_step_dev);
  OKGE_REQUIRE(ld_dy % 4 == 0 && aligned16(dy) && (dx == nullptr || (ld_dx % 4 == 0 && aligned16(dx))), "gradients must be 16-byte aligned");
  OKGE_REQUIRE(seg != nullptr || n_seg == 1, "several segments need their bounds");
  if (int rc = okge_device_check()) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const BnGrid g = bn_grid(n_rows, D);
  double* partial = static_cast<double*>(workspace);
  float* coef = reinterpret_cast<float*>(partial + static_cast<int64_t>(n_seg) * g.chunks * D * 2);
  const dim3 grid(g.col_tiles, g.chunks, n_seg);
  OKGE_LAUNCH((bn_partial_kernel<true>), grid, kBnThreads, 0, s, dy, ld_dy, x, ld_x, save_mean, save_invstd, seg, n_rows, D, partial, drop);
  OKGE_LAUNCH((bn_bwd_finalize_kernel), (D + kFinCols - 1) / kFinCols, dim3(kFinCols, kFinLanes), 0, s, partial, seg, n_seg, n_rows, g.chunks, D, coef, dgamma, dbeta);
  if (dx != nullptr && n_rows > 0)
    OKGE_LAUNCH((bn_bwd_apply_kernel), grid, kBnThreads, 0, s, dy, ld_dy, x, ld_x, seg, n_rows, D, save_mean, save_invstd, coef, gamma, dx, ld_dx, drop);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_bn_eval_fwd(const float* x, int64_t ld_x, int64_t n_rows, int D, const float* gamma, const float* beta,
                                const float* running_mean, const float* running_var, float eps, float* y, int64_t ld_y,
                                void* stream) {
  if (n_rows == 0) return OKGE_OK;
  if (int rc = check_common(x, ld_x, 1, n_rows, D)) return rc;
  OKGE_REQUIRE(y && running_mean && running_var, "null pointer");
  OKGE_REQUIRE(ld_y % 4 == 0 && aligned16(y) && aligned16(running_mean) && aligned16(running_var), "operands must be 16-byte aligned");
  if (int rc = okge_device_check()) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const BnGrid g = bn_grid(n_rows, D);
  const dim3 grid(g.col_tiles, g.chunks, 1);
  OKGE_LAUNCH((bn_apply_kernel<true>), grid, kBnThreads, 0, s, x, ld_x, nullptr, n_rows, D, running_mean, running_var, eps, gamma, beta, y, ld_y,
                                                    make_drop(0.f, 0, 0, nullptr));
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

// ---- split phases (synchronised batch norm over a row-partitioned operand) --------------------------------------------

extern "C" int okge_bn_col_sums(const float* a, int64_t ld_a, const float* x, int64_t ld_x, const float* mean,
                                const float* invstd, int64_t n_rows, int D, double* sums, void* workspace, void* stream) {
  if (int rc = check_common(a, ld_a, 1, n_rows, D)) return rc;
  OKGE_REQUIRE(sums && workspace, "null pointer");
  OKGE_REQUIRE(x == nullptr || (mean && invstd && ld_x % 4 == 0 && aligned16(x) && aligned16(mean) && aligned16(invstd)),
               "the backward sums need x, mean and invstd (16-byte aligned)");
  if (int rc = okge_device_check()) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const BnGrid g = bn_grid(n_rows, D);
  double* partial = static_cast<double*>(workspace);
  const dim3 grid(g.col_tiles, g.chunks, 1);
  if (x == nullptr)
    OKGE_LAUNCH((bn_partial_kernel<false>), grid, kBnThreads, 0, s, a, ld_a, nullptr, 0, nullptr, nullptr, nullptr, n_rows, D, partial,
                                                         make_drop(0.f, 0, 0, nullptr));
  else
    OKGE_LAUNCH((bn_partial_kernel<true>), grid, kBnThreads, 0, s, a, ld_a, x, ld_x, mean, invstd, nullptr, n_rows, D, partial,
                                                        make_drop(0.f, 0, 0, nullptr));
  OKGE_LAUNCH((bn_sums_kernel), (D + kFinCols - 1) / kFinCols, dim3(kFinCols, kFinLanes), 0, s, partial, g.chunks, D, sums);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_bn_normalize(const float* x, int64_t ld_x, int64_t n_rows, int D, const float* mean, const float* invstd,
                                 const float* gamma, const float* beta, float* y, int64_t ld_y, void* stream) {
  if (n_rows == 0) return OKGE_OK;
  if (int rc = check_common(x, ld_x, 1, n_rows, D)) return rc;
  OKGE_REQUIRE(y && mean && invstd, "null pointer");
  OKGE_REQUIRE(ld_y % 4 == 0 && aligned16(y) && aligned16(mean) && aligned16(invstd), "operands must be 16-byte aligned");
  if (int rc = okge_device_check()) return rc;
  const BnGrid g = bn_grid(n_rows, D);
  OKGE_LAUNCH((bn_apply_kernel<false>), dim3(g.col_tiles, g.chunks, 1), kBnThreads, 0, static_cast<cudaStream_t>(stream), x, ld_x, nullptr, n_rows, D, mean, invstd, 0.f, gamma, beta, y, ld_y, make_drop(0.f, 0, 0, nullptr));
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_bn_normalize_bwd(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x, int64_t n_rows, int D,
                                     const float* mean, const float* invstd, const float* coef, const float* gamma,
                                     float* dx, int64_t ld_dx, void* stream) {
  if (n_rows == 0) return OKGE_OK;
  if (int rc = check_common(x, ld_x, 1, n_rows, D)) return rc;
  OKGE_REQUIRE(dy && dx && mean && invstd && coef, "null pointer");
  OKGE_REQUIRE(ld_dy % 4 == 0 && ld_dx % 4 == 0 && aligned16(dy) && aligned16(dx) && aligned16(mean) && aligned16(invstd) &&
                   aligned16(coef), "operands must be 16-byte aligned");
  if (int rc = okge_device_check()) return rc;
  const BnGrid g = bn_grid(n_rows, D);
  OKGE_LAUNCH((bn_bwd_apply_kernel), dim3(g.col_tiles, g.chunks, 1), kBnThreads, 0, static_cast<cudaStream_t>(stream), dy, ld_dy, x, ld_x, nullptr, n_rows, D, mean, invstd, coef, gamma, dx, ld_dx, make_drop(0.f, 0, 0, nullptr));
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}
