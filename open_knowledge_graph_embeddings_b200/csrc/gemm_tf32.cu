// 1-vs-all scoring GEMM for sm_100a: C[M,N] = A[M,K] * B[N,K]^T with TF32 inputs / FP32 accumulate.
//
//   * operands: fp32 row-major (K contiguous) straight from the embedding tables, moved by TMA
//     (cp.async.bulk.tensor, 128B swizzle) into a 4-stage shared-memory ring;
//   * math: tcgen05.mma.kind::tf32, 128x256x8 per instruction, issued by one elected thread,
//     accumulators in TMEM (2 x 256 columns, double buffered so the epilogue of tile i overlaps
//     the MMAs of tile i+1);
//   * epilogue: 8 warps read TMEM with tcgen05.ld (one query row per thread, 32 columns per load)
//     and apply one of the fused epilogues below, so the B x N score matrix is never written
//     unless the caller asks for it (MODE_STORE).
//
// Persistent kernel: grid = min(#work items, #SMs); work item = (m_tile, n_tile, k_split), m fastest
// so that CTAs running at the same time share the (large) entity tile through L2.
//
// Reference semantics implemented by the epilogues (paths relative to the reference root):
//   MODE_STORE   q e^T                                   openkge/model.py:206-215, 270-272
//   MODE_BCE     BCEWithLogitsLoss(sum) and its gradient openkge/trainer.py:93-106
//   MODE_LSE     log_softmax(dim=1) row statistics       openkge/trainer.py:99-100
//   MODE_SMGRAD  gradient of KLDivLoss(log_softmax)      openkge/trainer.py:99-106
//   MODE_RANK    count-greater / count-equal             openkge/dataset.py:441-444
#include "okge_common.cuh"

#include <limits.h>
#include <math.h>
#include <string.h>

#include <type_traits>

#ifndef OKGE_ADAGRAD_CHUNK_COLS
#define OKGE_ADAGRAD_CHUNK_COLS 32
#endif

namespace okge {

namespace {

constexpr int kBM = 128;       // tile rows    = UMMA M = TMEM lanes
constexpr int kBN = 256;       // tile columns = UMMA N = TMEM columns per accumulator
constexpr int kBK = 32;        // fp32 elements per stage along K (K-major: one 128-byte SW128 row)
constexpr int kUmmaK = 8;      // K per tcgen05.mma for tf32 (32 bytes)
constexpr int kTmemCols = 512;                   // 2 accumulators x 256 columns
constexpr int kEpiStageBytes = 32 * 128;         // one 32-row x 32-column fp32 chunk per epilogue warp (SW128)

enum Mode : int { MODE_STORE = 0, MODE_BCE = 1, MODE_LSE = 2, MODE_SMGRAD = 3, MODE_RANK = 4, MODE_ADAGRAD = 5 };

// Operand source forms (a_mode / b_mode). 0/1 are K-major in shared memory (SWIZZLE_128B), 2/3 are MN-major
// (SWIZZLE_128B with 32-byte atoms, the only MN-major form tcgen05 accepts for 4-byte operands).
enum OperandMode : int { OP_ROW_MAJOR = OKGE_ROW_MAJOR, OP_K_PANELS = OKGE_K_PANELS, OP_COL_MAJOR = OKGE_COL_MAJOR,
                         OP_MN_PANELS = OKGE_MN_PANELS };

constexpr int kAdagradChunkCols = OKGE_ADAGRAD_CHUNK_COLS;

// Per-epilogue kernel shape. The loss epilogues do ~20 instructions per score, so they get 16 epilogue warps
// (4 per scheduler); to keep all 4 pipeline stages (the mainloop is TMA-latency bound with 3) their TMA-store
// staging is 2 KiB per warp: a 32 x 32 chunk leaves as two 32-row x 16-column halves (SWIZZLE_64B).
template <int MODE>
struct Cfg {
  static constexpr bool kStaged = MODE == MODE_STORE || MODE == MODE_BCE || MODE == MODE_SMGRAD || MODE == MODE_ADAGRAD;
  static constexpr bool kHalfChunks = MODE == MODE_BCE || MODE == MODE_SMGRAD;
  static constexpr int kEpiWarps = (MODE == MODE_STORE || MODE == MODE_ADAGRAD) ? 8 : 16;
  // MODE_ADAGRAD streams the parameter and its accumulator through shared memory (2 x (4 + 4) KiB per warp, loads one
  // chunk ahead) and is HBM-bound, so it gives up two pipeline stages for that staging.
  // Its operands are MN-major (dS^T panels, Q column-major), whose K extent per stage is free: 16-row stages keep
  // four loads in flight in the same 96 KiB.
  static constexpr int kStageK = MODE == MODE_ADAGRAD ? 16 : kBK;
  static constexpr int kABytes = kBM * kStageK * 4;
  static constexpr int kBBytes = kBN * kStageK * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = 4;
  static constexpr int kGroups = kEpiWarps / 4;            // column groups of the 256-column accumulator
  static constexpr int kColsPerGroup = kBN / kGroups;
  static constexpr int kThreads = 32 * (2 + kEpiWarps);
  static constexpr int kEpiWarpBytes = MODE == MODE_ADAGRAD ? 4 * kEpiStageBytes : (kHalfChunks ? kEpiStageBytes / 2 : kEpiStageBytes);
  static constexpr int kEpiBytes = kStaged ? kEpiWarps * kEpiWarpBytes : 0;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align slack*/ + 512 /*barriers*/;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KiB of shared memory a CTA can opt into");
};
constexpr int kLseGroups = Cfg<MODE_LSE>::kGroups;

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, splits;
  int n_fastest;           // work-item order, see decode_work
  int k_chunks, k_chunks_per_split;
  int a_mode, b_mode;      // OperandMode
  // shared-memory matrix descriptors of the two operands: lo = start address field | desc_lo, per UMMA K step += kadv
  uint32_t a_desc_lo, a_desc_hi, a_kadv;
  uint32_t b_desc_lo, b_desc_hi, b_kadv;
  uint32_t idesc;
  // MODE_STORE
  float* C;
  long long ldc;
  long long split_stride;  // elements between split partials (0 when splits == 1)
  float alpha;
  const float* alpha_dev;
  float acc_scale;  // truncation-bias correction applied to the raw accumulator (scores)
  // labels (BCE, SMGRAD): every label is y_base inside the tiles; positives are fixed up by sparse_label_fix_kernel
  float y_base;
  const int* n_limit_dev;  // BCE: columns >= *n_limit_dev are padding (no loss term, zero gradient); nullable
  int loss_rows;           // BCE + RANK4: rows >= loss_rows only carry ranking thresholds (no loss term)
  double* loss_sum;
  float* dS;   // K-panel layout of the [M, N] gradient:    [ceil(N/32)][M][32]   (TMA store through tmap_c)
  float* dST;  // K-panel layout of its transpose [N, M]:  [ceil(M/32)][N][32]   (optional, direct stores)
  // LSE
  float* part_max;  // [n_tiles * kLseGroups, M]
  float* part_sum;  // [n_tiles * kLseGroups, M]
  const float* row_lse;
  const float* row_weight;
  // RANK
  const float* thresh;
  int* greater;
  int* equal;
  // ADAGRAD (param / state go through tmap_c / tmap_d)
  float clr, eps, weight_decay;
  const int* extra_map;    // [M] slot of an additional gradient row per output row, -1 = none (nullable)
  const float* extra;      // [slots, N] row-major
  long long ld_extra;
};

// Shared-memory matrix descriptors (PTX "tcgen05 shared memory descriptor", version 1 = Blackwell).
//   K-major, SWIZZLE_128B (layout type 2): rows of 128 bytes, 8-row groups 1024 B apart (SBO); LBO unused (1).
//     One UMMA K step (8 tf32 = 32 bytes) advances the start address by 32 B  -> +2 in the (addr >> 4) field.
//   MN-major, SWIZZLE_128B with 32-byte atoms (layout type 1): each K index is a 128-byte row holding 32 consecutive
//     MN elements; 4-row groups are 512 B apart (SBO), 32-element MN blocks are one 32 x stage_k box apart (LBO).
//     One UMMA K step (8 rows = 1024 B) -> +64.
constexpr uint32_t kDescHiKMajor = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t kDescLoKMajor = 1u << 16;
constexpr uint32_t kKadvKMajor = 2;
constexpr uint32_t kDescHiMnMajor = (512u >> 4) | (1u << 14) | (1u << 29);
constexpr uint32_t desc_lo_mn_major(int stage_k) { return (static_cast<uint32_t>(32 * stage_k * 4) >> 4) << 16; }
constexpr uint32_t kKadvMnMajor = 1024u >> 4;

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lo_or, uint32_t hi) {
  return (static_cast<uint64_t>(hi) << 32) | static_cast<uint64_t>(lo_or | ((smem_addr >> 4) & 0x3FFFu));
}

// Instruction descriptor: D = F32, A = B = TF32, N = 256, M = 128; bit 15 / 16 = A / B is MN-major.
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) |
                                (static_cast<uint32_t>(kBN >> 3) << 17) |
                                (static_cast<uint32_t>(kBM >> 4) << 24);

struct WorkItem {
  int m, n, split;
};

// Work items that run at the same time should share their LARGE operand tile through L2, so the tile index of the
// dimension with fewer tiles runs fastest: the forward pass (4 query tiles x thousands of entity tiles) walks m
// fastest, dE = dS^T Q (thousands of entity tiles x 2 column tiles) walks n fastest.
__device__ __forceinline__ WorkItem decode_work(int w, const GemmParams& p) {
  WorkItem it;
  if (p.n_fastest) {
    it.n = w % p.n_tiles;
    const int rest = w / p.n_tiles;
    it.m = rest % p.m_tiles;
    it.split = rest / p.m_tiles;
  } else {
    it.m = w % p.m_tiles;
    const int rest = w / p.m_tiles;
    it.n = rest % p.n_tiles;
    it.split = rest / p.n_tiles;
  }
  return it;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;

// log(1 + x) for x in [0, 1]: x * P5(x), max abs error 6.0e-6, mean error 9e-8 (least-squares fit on Chebyshev nodes;
// the loss tolerance is 1e-3 relative and a softplus term is ~0.69). Keeps the softplus of the BCE epilogue at two MUFU
// ops per score (ex2, rcp); the rest runs on the FMA pipe.
__device__ __forceinline__ float log1p_unit(float x) {
  float p = -0.02397775463759899f;
  p = fmaf(p, x, 0.10149542987346649f);
  p = fmaf(p, x, -0.21028946340084076f);
  p = fmaf(p, x, 0.3252934515476227f);
  p = fmaf(p, x, -0.49937233328819275f);
  p = fmaf(p, x, 0.9999918341636658f);
  return p * x;
}

// One load of a pipeline stage: A (128 rows) and B (256 rows) of K chunk `kc`, in whichever form each operand has.
template <int kStageK>
__device__ __forceinline__ void load_operand(int mode, uint32_t dst, const CUtensorMap* tm, uint32_t bar, int row0,
                                             int rows, int kc) {
  if (mode == OP_ROW_MAJOR) {
    tma_load_2d(dst, tm, bar, kc * kStageK, row0);
  } else if (mode == OP_K_PANELS) {
    tma_load_3d(dst, tm, bar, 0, row0, kc);
  } else if (mode == OP_MN_PANELS) {
    tma_load_3d(dst, tm, bar, 0, kc * kStageK, row0 >> 5);
  } else {  // OP_COL_MAJOR: one 32 x kStageK box per 32-row MN block (boxes past the matrix edge are zero-filled)
    for (int i = 0; i < rows / 32; ++i)
      tma_load_2d(dst + static_cast<uint32_t>(i * 32 * kStageK * 4), tm, bar, row0 + 32 * i, kc * kStageK);
  }
}

// LIMIT (MODE_BCE only): the number of label-carrying columns is read from p.n_limit_dev; a separate instantiation so
// that the regular loss kernel, which sits at its register cap, is compiled without it.
// RANK4 (MODE_BCE only, evaluation): the loss pass also counts, for up to 4 ranked answers per query row, the scores
// above / equal to the answer's threshold (p.thresh [M, 4], +inf = unused slot; p.greater / p.equal [M, 4]) -- the filtered
// ranking of openkge/dataset.py:441-444 without a second contraction over the candidates.
template <int MODE, bool LIMIT = false, bool RANK4 = false>
__global__ void __launch_bounds__(Cfg<MODE>::kThreads, 1)
okge_gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a,
                      const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_c,
                      const __grid_constant__ CUtensorMap tmap_d, const GemmParams p) {
  using C = Cfg<MODE>;
  constexpr int kStages = C::kStages;
  constexpr int kStageK = C::kStageK;
  constexpr int kABytes = C::kABytes;
  constexpr int kStageBytes = C::kStageBytes;
  constexpr int kNumEpiWarps = C::kEpiWarps;
  constexpr int kColsPerGroup = C::kColsPerGroup;
  extern __shared__ uint8_t smem_raw[];
  // swizzled tiles need 1024-byte alignment.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + kStages * kStageBytes;   // 4 KB of TMA-store staging per epilogue warp
  const uint32_t bar_base = epi_base + C::kEpiBytes;
  // barrier layout (8 bytes each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_work = p.m_tiles * p.n_tiles * p.splits;
  // Warp roles: epilogue warps 0 .. kNumEpiWarps-1 (TMEM lane quarter = warp % 4), then the TMA producer and the MMA
  // issuer as the LAST two warps: the warp scheduler favours higher warp ids, and these two single-thread roles must
  // never wait behind busy epilogue warps for an issue slot (with them as warps 0 / 1 the loss epilogue starved them).
  constexpr int kProducerWarp = kNumEpiWarps, kMmaWarp = kNumEpiWarps + 1;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (C::kStaged) tma_prefetch_desc(&tmap_c);
    if (MODE == MODE_ADAGRAD) {
      tma_prefetch_desc(&tmap_d);
      for (int i = 0; i < 4 * kNumEpiWarps; ++i) mbar_init(bar_base + 8u * (2 * kStages + 5 + i), 1);   // <= 4 per warp
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), kNumEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc<kTmemCols>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == kProducerWarp) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const WorkItem it = decode_work(w, p);
        const int kc_begin = it.split * p.k_chunks_per_split;
        const int kc_end = min(kc_begin + p.k_chunks_per_split, p.k_chunks);
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), kStageBytes);
          const uint32_t sa = smem_base + stage * kStageBytes;
          load_operand<kStageK>(p.a_mode, sa, &tmap_a, full_bar(stage), it.m * kBM, kBM, kc);
          load_operand<kStageK>(p.b_mode, sa + kABytes, &tmap_b, full_bar(stage), it.n * kBN, kBN, kc);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const WorkItem it = decode_work(w, p);
        const int kc_begin = it.split * p.k_chunks_per_split;
        const int kc_end = min(kc_begin + p.k_chunks_per_split, p.k_chunks);
        mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kBN);
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint64_t adesc = make_desc(sa, p.a_desc_lo, p.a_desc_hi);
          const uint64_t bdesc = make_desc(sa + kABytes, p.b_desc_lo, p.b_desc_hi);
#pragma unroll
          for (int k = 0; k < kStageK / kUmmaK; ++k) {
            umma_tf32(tmem_d, adesc + static_cast<uint64_t>(p.a_kadv * k),
                      bdesc + static_cast<uint64_t>(p.b_kadv * k), p.idesc,
                      (kc > kc_begin || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tmem_full_bar(acc));  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if constexpr (MODE == MODE_ADAGRAD) {
    // ===================== epilogue: Adagrad step fused onto the gradient tile =====================
    // g = alpha * acc (+ extra row); g' = g + wd p; G += g'^2; p -= clr g' / (sqrt(G) + eps). Each warp walks its
    // half-chunks (32 rows x 16 columns) in order. p and G of the next kAhead half-chunks are in flight (TMA, 64-byte
    // swizzle) while one is updated in shared memory and TMA-stored back: the tables move through HBM exactly once
    // each way and the read latency is covered by 12 KiB in flight per warp.
    constexpr int kHalf = kAdagradChunkCols;         // columns per chunk (16: 64-byte rows / SWIZZLE_64B, 32: SWIZZLE_128B)
    constexpr int kBufBytes = 2 * 32 * kHalf * 4;    // p | G
    constexpr int kBufs = C::kEpiWarpBytes / kBufBytes, kAhead = kBufs - 1;
    constexpr int kChunks = kColsPerGroup / kHalf;
    const int ew = warp;
    const int quarter = warp & 3;
    const int group = ew >> 2;
    const uint32_t wbuf = epi_base + static_cast<uint32_t>(ew * C::kEpiWarpBytes);
    const uint32_t ldbar = bar_base + 8u * (2 * kStages + 5 + 4 * ew);                // one mbarrier per buffer (<= 4)
    float alpha_eff = p.alpha;
    if (p.alpha_dev != nullptr) alpha_eff *= __ldg(p.alpha_dev);
    const float clr = p.clr, eps = p.eps, wd = p.weight_decay;
    auto n_valid = [&](const WorkItem& it) {
      const int rem = p.N - (it.n * kBN + group * kColsPerGroup);
      return rem <= 0 ? 0 : min(kChunks, (rem + kHalf - 1) / kHalf);
    };
    int pw = blockIdx.x, pc = 0;          // prefetch cursor: (work item, half-chunk) of the next load
    int n_issued = 0, n_done = 0;
    auto skip_invalid = [&]() {
      while (pw < total_work) {
        if (pc < n_valid(decode_work(pw, p))) return;
        pw += gridDim.x;
        pc = 0;
      }
    };
    auto issue_next = [&]() {
      const WorkItem it = decode_work(pw, p);
      const int buf = n_issued % kBufs;
      const uint32_t pb = wbuf + static_cast<uint32_t>(buf * kBufBytes);
      const uint32_t bar = ldbar + 8u * buf;
      if (lane == 0) {
        // this buffer was last read by the store of half-chunk n_issued - kBufs; only the kBufs - kAhead stores
        // committed after that one may still be draining
        tma_store_wait_read_le<kBufs - kAhead>();
        mbar_arrive_expect_tx(bar, kBufBytes);
        const int c0 = it.n * kBN + group * kColsPerGroup + pc * kHalf, r0 = it.m * kBM + quarter * 32;
        tma_load_3d(pb, &tmap_c, bar, c0, r0, 0);
        tma_load_3d(pb + kBufBytes / 2, &tmap_d, bar, c0, r0, 0);
      }
      ++n_issued;
      ++pc;
      skip_invalid();
    };
    skip_invalid();
    for (int i = 0; i < kAhead && pw < total_work; ++i) issue_next();
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const WorkItem it = decode_work(w, p);
      const int row0 = it.m * kBM + quarter * 32;
      const int row = row0 + lane;
      int slot = -1;
      if (p.extra_map != nullptr && row < p.M) slot = __ldg(p.extra_map + row);
      const int nv = n_valid(it);
      mbar_wait(tmem_full_bar(acc), acc_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < nv; ++chunk) {
        const int col0 = it.n * kBN + group * kColsPerGroup + chunk * kHalf;
        uint32_t v[kHalf];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(acc * kBN + group * kColsPerGroup + chunk * kHalf);
        tmem_ld_chunk(taddr, v);
        tmem_ld_wait();
        float g[kHalf];
#pragma unroll
        for (int t = 0; t < kHalf; ++t) g[t] = alpha_eff * __uint_as_float(v[t]);
        if (slot >= 0) {   // rare (the batch's own entities): add the lookup gradient row of this table row
          const float* ex = p.extra + static_cast<long long>(slot) * p.ld_extra + col0;
#pragma unroll
          for (int t = 0; t < kHalf; ++t)
            if (col0 + t < p.N) g[t] += __ldg(ex + t);
        }
        const int buf = n_done % kBufs;
        const uint32_t pb = wbuf + static_cast<uint32_t>(buf * kBufBytes);
        mbar_wait(ldbar + 8u * buf, static_cast<uint32_t>((n_done / kBufs) & 1));
#pragma unroll
        for (int c = 0; c < kHalf / 4; ++c) {
          // row = lane; 16-byte chunk c of the row sits at the swizzled position of the TMA layout
          const uint32_t off = kHalf == 16
              ? static_cast<uint32_t>(lane) * 64u + (static_cast<uint32_t>(c ^ ((lane >> 1) & 3)) << 4)
              : static_cast<uint32_t>(lane) * 128u + (static_cast<uint32_t>(c ^ (lane & 7)) << 4);
          float pv[4], sv[4];
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pv[0]), "=f"(pv[1]), "=f"(pv[2]), "=f"(pv[3]) : "r"(pb + off));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sv[0]), "=f"(sv[1]), "=f"(sv[2]), "=f"(sv[3])
                       : "r"(pb + kBufBytes / 2 + off));
#pragma unroll
          for (int t = 0; t < 4; ++t) adagrad_elem_fast(pv[t], g[4 * c + t], sv[t], clr, eps, wd);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(pb + off), "f"(pv[0]), "f"(pv[1]), "f"(pv[2]), "f"(pv[3]) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(pb + kBufBytes / 2 + off), "f"(sv[0]), "f"(sv[1]),
                       "f"(sv[2]), "f"(sv[3]) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmap_c, pb, col0, row0, 0);
          tma_store_3d(&tmap_d, pb + kBufBytes / 2, col0, row0, 0);
          tma_store_commit();
        }
        ++n_done;
        if (pw < total_work) issue_next();   // refill: kAhead loads stay in flight
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (lane == 0) tma_store_wait_all();
    __syncwarp();
  } else {
    // ===================== epilogue =====================
    const int ew = warp;
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int group = ew >> 2;          // which block of kColsPerGroup columns
    const uint32_t stage_buf = epi_base + static_cast<uint32_t>(ew * C::kEpiWarpBytes);
    int acc = 0;
    uint32_t acc_phase = 0;
    double loss_acc = 0.0;
    float alpha_eff = 1.0f;
    if (MODE == MODE_STORE) {
      alpha_eff = p.alpha;
      if (p.alpha_dev != nullptr) alpha_eff *= __ldg(p.alpha_dev);
    }

    // registers -> swizzled smem chunk (row = lane, 128 bytes) -> one TMA tensor store per warp and chunk; the TMA
    // unit writes full rows coalesced and clips the box at the matrix edges
    auto stage_and_store = [&](const uint32_t (&v)[32], int c0, int c1, int c2) {
      if (lane == 0) tma_store_wait_read();        // the previous chunk's store has drained this buffer
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t addr = stage_buf + static_cast<uint32_t>(lane) * 128u + (static_cast<uint32_t>(c ^ (lane & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[4 * c + 0]), "r"(v[4 * c + 1]),
                     "r"(v[4 * c + 2]), "r"(v[4 * c + 3])
                     : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&tmap_c, stage_buf, c0, c1, c2);
        tma_store_commit();
      }
    };

    // the same through a 2 KB buffer: two 32-row x 16-column halves, 64-byte rows, SWIZZLE_64B (16-byte chunk index
    // XOR bits 7..8 of the address = (lane >> 1) & 3); used for the dS panels, c0 = first column inside the panel
    const uint64_t stream_policy = l2_policy_evict_first();
    auto stage_and_store_halves = [&](const uint32_t (&v)[32], int c1, int c2) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t addr = stage_buf + static_cast<uint32_t>(lane) * 64u + (static_cast<uint32_t>(c ^ ((lane >> 1) & 3)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[16 * h + 4 * c + 0]),
                       "r"(v[16 * h + 4 * c + 1]), "r"(v[16 * h + 4 * c + 2]), "r"(v[16 * h + 4 * c + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d_hint(&tmap_c, stage_buf, 16 * h, c1, c2, stream_policy);   // written once, read by the next kernels
          tma_store_commit();
        }
      }
    };

#ifdef OKGE_DS_DIRECT_STORE
    // The same 32 x 32 chunk written WITHOUT the bulk-store engine: in the K-panel layout the chunk is one contiguous
    // 4 KB block of dS ([panel][row][32]), so after the transposing trip through the 2 KB staging buffer every warp
    // instruction stores 8 rows x 64 contiguous bytes. No async-proxy fence, no store group to wait for before the buffer
    // is reused (a __syncwarp orders the shared-memory reads), and the TMA unit is left to the operand loads.
    auto store_halves_direct = [&](const uint32_t (&v)[32], int row0, int panel) {
      float* blk = p.dS + (static_cast<long long>(panel) * p.M + row0) * 32;
      const int rows_valid = p.M - row0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t addr = stage_buf + static_cast<uint32_t>(lane) * 64u + (static_cast<uint32_t>(c ^ ((lane >> 1) & 3)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[16 * h + 4 * c + 0]),
                       "r"(v[16 * h + 4 * c + 1]), "r"(v[16 * h + 4 * c + 2]), "r"(v[16 * h + 4 * c + 3])
                       : "memory");
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = 8 * j + (lane >> 2), c = lane & 3;
          const uint4 val = lds_v4(stage_buf + static_cast<uint32_t>(r) * 64u + (static_cast<uint32_t>(c ^ ((r >> 1) & 3)) << 4));
          if (r < rows_valid) stg_v4_hint(blk + r * 32 + 16 * h + 4 * c, val, stream_policy);
        }
      }
    };
#endif

    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const WorkItem it = decode_work(w, p);
      const int row = it.m * kBM + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const int cbeg = it.n * kBN + group * kColsPerGroup;   // first column of this warp group

      // ---- per-row prologue ----
      float row_lse = 0.f, row_w = 0.f, thr = 0.f;
      if (MODE == MODE_SMGRAD) {
        if (row_ok) { row_lse = __ldg(p.row_lse + row) * kLog2e; row_w = __ldg(p.row_weight + row); }
      }
      if (MODE == MODE_RANK) {
        if (row_ok) thr = __ldg(p.thresh + row);
      }
      float4 thr4 = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);   // RANK4: +inf never counts
      int cg0 = 0, cg1 = 0, cg2 = 0, cg3 = 0, ce0 = 0, ce1 = 0, ce2 = 0, ce3 = 0;
      if (RANK4) {
        if (row_ok) thr4 = __ldg(reinterpret_cast<const float4*>(p.thresh) + row);
      }
      float run_max = -INFINITY, run_sum = 0.f;   // LSE
      float tile_loss = 0.f;                      // BCE: fp32 inside a tile (<= 128 columns per thread), fp64 across tiles
      int cnt_g = 0, cnt_e = 0;                   // RANK

      mbar_wait(tmem_full_bar(acc), acc_phase);
      tcgen05_fence_after();

#pragma unroll 1
      for (int chunk = 0; chunk < kColsPerGroup / 32; ++chunk) {
        const int col0 = cbeg + chunk * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(acc * kBN + group * kColsPerGroup + chunk * 32);
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        // valid columns in this chunk (warp-uniform); <= 0 when the whole chunk is padding behind the device-side limit
        // Columns that carry labels (warp-uniform): all N, or fewer when the caller padded the candidate list to a fixed
        // capacity and keeps the real count in device memory (CUDA-graph replay of batch-shared candidate lists); <= 0
        // when the whole chunk is padding.
        int ncols = min(32, p.N - col0);
        if (LIMIT) ncols = min(ncols, __ldg(p.n_limit_dev) - col0);

        // `full` is a compile-time tag: the edge chunk (ncols < 32) runs a masked copy of the same body
        auto body = [&](auto full) {
          constexpr bool kFull = decltype(full)::value;
          if (MODE == MODE_STORE) {
#pragma unroll
            for (int t = 0; t < 32; ++t) v[t] = __float_as_uint(alpha_eff * __uint_as_float(v[t]));
            stage_and_store(v, col0, it.m * kBM + quarter * 32, it.split);
          } else if (MODE == MODE_BCE || MODE == MODE_SMGRAD) {
            // Dense part only: every label is y_base here; the (very sparse) positives are corrected afterwards by
            // sparse_label_fix_kernel, which keeps all CSR look-ups off this epilogue.
            float lsum = 0.f;
            if (MODE == MODE_BCE) {
              // s = c * raw (c: truncation-bias correction). softplus(s) - s y = c max(raw, 0) + log1p(exp(-|s|)) - y s,
              // sigmoid(s) = r or e r with e = exp(-|s|), r = 1 / (1 + e). ~18 instructions per score (20 with label
              // smoothing), two of them MUFU; `smooth` is warp-uniform and selects a second copy of the loop.
              const float c = p.acc_scale, k_exp = -p.acc_scale * kLog2e;
              auto scores = [&](auto smooth_tag) {
                constexpr bool kSmooth = decltype(smooth_tag)::value;
                const float y0 = p.y_base, cy0 = -p.acc_scale * p.y_base;
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                  const float raw = __uint_as_float(v[t]);
                  const float e = ex2_approx(fabsf(raw) * k_exp);   // exp(-|s|) in (0, 1]
                  float sig = 0.f;
                  if (!RANK4) {                                     // the gradient is not needed in evaluation
                    const float r = rcp_approx(1.f + e);
                    sig = (raw >= 0.f) ? r : e * r;
                  }
                  float term = fmaf(c, fmaxf(raw, 0.f), log1p_unit(e));
                  if (kSmooth) term = fmaf(cy0, raw, term);
                  if (kFull || t < ncols) lsum += term;
                  if (RANK4) {
                    const float s = raw * p.acc_scale;               // the score exactly as MODE_RANK forms it
                    if (kFull || t < ncols) {
                      cg0 += (thr4.x < s) ? 1 : 0; ce0 += (thr4.x == s) ? 1 : 0;
                      cg1 += (thr4.y < s) ? 1 : 0; ce1 += (thr4.y == s) ? 1 : 0;
                      cg2 += (thr4.z < s) ? 1 : 0; ce2 += (thr4.z == s) ? 1 : 0;
                      cg3 += (thr4.w < s) ? 1 : 0; ce3 += (thr4.w == s) ? 1 : 0;
                    }
                  }
                  if (!RANK4) {
                    const float g = kSmooth ? sig - y0 : sig;
                    // round to nearest TF32 (ties away, like cvt.rna; |g| < 1 so no inf / nan case): dS only feeds the
                    // gradient GEMMs, whose tensor cores would otherwise truncate it
                    v[t] = (kFull || t < ncols) ? ((__float_as_uint(g) + 0x1000u) & 0xFFFFE000u) : 0u;
                  }
                }
              };
              if (p.y_base == 0.f) scores(std::false_type{}); else scores(std::true_type{});
            } else {
#pragma unroll
              for (int t = 0; t < 32; ++t) {
                const float s = __uint_as_float(v[t]) * p.acc_scale;
                const float g = row_w * ex2_approx(fmaf(s, kLog2e, -row_lse)) - p.y_base;
                v[t] = (kFull || t < ncols) ? __float_as_uint(round_tf32(g)) : 0u;
              }
            }
            if (MODE == MODE_BCE && row_ok && (!RANK4 || row < p.loss_rows)) tile_loss += lsum;
            if (!RANK4 && p.dST != nullptr && row < ((p.M + 31) & ~31)) {
              // panel = block of 32 query rows (exactly this warp's lanes); rows >= M are the zero padding
              float* dcol = p.dST + (static_cast<long long>(row >> 5) * p.N + col0) * 32 + lane;
#pragma unroll
              for (int t = 0; t < 32; ++t)
                if (kFull || col0 + t < p.N) dcol[t * 32] = row_ok ? __uint_as_float(v[t]) : 0.f;
            }
            // dS panel = 32 columns x all rows: this chunk is rows [row0, row0 + 32) of panel col0 / 32
#ifdef OKGE_DS_DIRECT_STORE
            if (!RANK4 && p.dS != nullptr) store_halves_direct(v, it.m * kBM + quarter * 32, col0 >> 5);
#else
            if (!RANK4 && p.dS != nullptr) stage_and_store_halves(v, it.m * kBM + quarter * 32, col0 >> 5);
#endif
          } else if (MODE == MODE_LSE) {
            float cmax = -INFINITY;
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              v[t] = __float_as_uint(__uint_as_float(v[t]) * (p.acc_scale * kLog2e));   // base-2 domain
              if (kFull || t < ncols) cmax = fmaxf(cmax, __uint_as_float(v[t]));
            }
            const float new_max = fmaxf(run_max, cmax);
            float csum = 0.f;
#pragma unroll
            for (int t = 0; t < 32; ++t)
              if (kFull || t < ncols) csum += ex2_approx(__uint_as_float(v[t]) - new_max);
            run_sum = run_sum * ex2_approx(run_max - new_max) + csum;   // 2^(-inf) = 0 on the first chunk
            run_max = new_max;
          } else if (MODE == MODE_RANK) {
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              const float s = __uint_as_float(v[t]) * p.acc_scale;
              if (kFull || t < ncols) {
                cnt_g += (thr < s) ? 1 : 0;
                cnt_e += (thr == s) ? 1 : 0;
              }
            }
          }
        };
        if (ncols == 32) body(std::true_type{}); else body(std::false_type{});
      }

      // TMEM reads of this accumulator are done: hand it back to the MMA warp.
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;

      if (MODE == MODE_BCE) loss_acc += static_cast<double>(tile_loss);
      if (RANK4) {
        if (row_ok) {
          int* g4 = p.greater + 4 * static_cast<long long>(row);
          int* e4 = p.equal + 4 * static_cast<long long>(row);
          if (cg0) atomicAdd(g4 + 0, cg0);
          if (cg1) atomicAdd(g4 + 1, cg1);
          if (cg2) atomicAdd(g4 + 2, cg2);
          if (cg3) atomicAdd(g4 + 3, cg3);
          if (ce0) atomicAdd(e4 + 0, ce0);
          if (ce1) atomicAdd(e4 + 1, ce1);
          if (ce2) atomicAdd(e4 + 2, ce2);
          if (ce3) atomicAdd(e4 + 3, ce3);
        }
      }
      if (MODE == MODE_LSE) {
        if (row_ok) {   // natural-log domain again: max_e = max_2 / log2(e); the sum of exponentials is base-free
          const long long pidx = static_cast<long long>(it.n * C::kGroups + group) * p.M + row;
          p.part_max[pidx] = run_max * (1.0f / kLog2e);
          p.part_sum[pidx] = run_sum;
        }
      }
      if (MODE == MODE_RANK) {
        if (row_ok) {
          if (cnt_g) atomicAdd(p.greater + row, cnt_g);
          if (cnt_e) atomicAdd(p.equal + row, cnt_e);
        }
      }
    }

    if (C::kStaged) {
      if (lane == 0) tma_store_wait_all();
      __syncwarp();
    }
    if (MODE == MODE_BCE) {
      // one fp64 atomic per warp
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
      if (lane == 0 && loss_acc != 0.0) atomicAdd(p.loss_sum, loss_acc);
    }
  }

  // ---- teardown ----
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// The positives of the sparse label matrix, applied after the dense tile pass (one warp per CSR entry):
//   s = <q[b, :], e[n, :]> with the operand truncation and scale of the tensor-core pass,
//   MODE_BCE     loss -= s * (y_pos - y_base);  dS[b, n] = dST[n, b] = sigmoid(s) - y_pos
//   MODE_SMGRAD  dS[b, n] = dST[n, b] = row_weight[b] * exp(s - row_lse[b]) - y_pos
//   MODE_LSE     pos_score[p] = s
template <int MODE>
__global__ void sparse_label_fix_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ e,
                                        long long lde, int B, int N, int D, const int* __restrict__ pos_ptr,
                                        const int* __restrict__ pos_idx, float acc_scale, float y_pos,
                                        double* loss_sum, float y_delta, float* dS, float* dST, float* pos_score,
                                        const float* __restrict__ row_lse, const float* __restrict__ row_weight) {
  const int lane = threadIdx.x & 31;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int nnz = __ldg(pos_ptr + B);
  double loss_local = 0.0;
  for (int pidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; pidx < nnz; pidx += warps_total) {
    // row b with pos_ptr[b] <= pidx < pos_ptr[b + 1]
    int lo = 0, hi = B;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(pos_ptr + mid) <= pidx) lo = mid; else hi = mid;
    }
    const int b = lo;
    const int n = __ldg(pos_idx + pidx);
    if (n < 0 || n >= N) continue;
    const float* qr = q + static_cast<long long>(b) * ldq;
    const float* er = e + static_cast<long long>(n) * lde;
    float dot = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float qa = __uint_as_float(__float_as_uint(__ldg(qr + d)) & 0xFFFFE000u);   // the hardware truncates to TF32
      const float eb = __uint_as_float(__float_as_uint(__ldg(er + d)) & 0xFFFFE000u);
      dot = fmaf(qa, eb, dot);
    }
    dot = warp_sum(dot);
    const float s = dot * acc_scale;
    if (lane == 0) {
      float g = 0.f;
      if (MODE == MODE_BCE) {
        loss_local -= static_cast<double>(s) * static_cast<double>(y_delta);
        g = 1.f / (1.f + __expf(-s)) - y_pos;
      } else if (MODE == MODE_SMGRAD) {
        g = __ldg(row_weight + b) * __expf(s - __ldg(row_lse + b)) - y_pos;
      } else {
        pos_score[pidx] = s;
      }
      if (MODE != MODE_LSE) {
        g = round_tf32(g);
        if (dS != nullptr) dS[(static_cast<long long>(n >> 5) * B + b) * 32 + (n & 31)] = g;
        if (dST != nullptr) dST[(static_cast<long long>(b >> 5) * N + n) * 32 + (b & 31)] = g;
      }
    }
  }
  if (MODE == MODE_BCE && lane == 0 && loss_local != 0.0) atomicAdd(loss_sum, loss_local);
}

// C[m, n] = alpha * sum_s part[s, m, n]
__global__ void splitk_reduce_kernel(const float* __restrict__ part, long long split_stride,
                                     int splits, long long M, long long N, float alpha,
                                     const float* __restrict__ alpha_dev, float* __restrict__ C,
                                     long long ldc) {
  const long long total = M * N;
  float a = alpha;
  if (alpha_dev != nullptr) a *= __ldg(alpha_dev);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[k * split_stride + i];
    C[(i / N) * ldc + (i % N)] = a * s;
  }
}

// Two-stage merge of per-(column group, row) softmax partials into row_lse.
// stage 1: grid (ceil(M/128), kLseChunks): each thread merges a strided slice of P partials of its row
// stage 2: grid ceil(M/128): merges kLseChunks partials and writes log-sum-exp
constexpr int kLseChunks = 64;

__global__ void lse_merge_stage1(const float* __restrict__ pmax, const float* __restrict__ psum,
                                 int P, int M, float* __restrict__ omax, float* __restrict__ osum) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  float m = -INFINITY, s = 0.f;
  for (int pi = blockIdx.y; pi < P; pi += gridDim.y) {
    const float pm = pmax[static_cast<long long>(pi) * M + row];
    const float ps = psum[static_cast<long long>(pi) * M + row];
    const float nm = fmaxf(m, pm);
    if (nm > -INFINITY) s = s * __expf(m - nm) + ps * __expf(pm - nm);
    m = nm;
  }
  omax[static_cast<long long>(blockIdx.y) * M + row] = m;
  osum[static_cast<long long>(blockIdx.y) * M + row] = s;
}

__global__ void lse_merge_stage2(const float* __restrict__ pmax, const float* __restrict__ psum,
                                 int P, int M, float* __restrict__ row_lse) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  float m = -INFINITY;
  for (int pi = 0; pi < P; ++pi) m = fmaxf(m, pmax[static_cast<long long>(pi) * M + row]);
  float s = 0.f;
  for (int pi = 0; pi < P; ++pi) {
    const float pm = pmax[static_cast<long long>(pi) * M + row];
    if (pm > -INFINITY) s += psum[static_cast<long long>(pi) * M + row] * expf(pm - m);
  }
  row_lse[row] = m + logf(s);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
  }
  return fn;
}

// 2-D fp32 row-major [rows, k] tensor with `ld` elements between rows; box = [box_rows, 32 floats].
int make_tmap(CUtensorMap* out, const float* base, int64_t rows, int64_t k, int64_t ld,
              int box_rows) {   // K-major forms always use kBK = 32 (one 128-byte swizzle row)
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled not available from the driver");
    return OKGE_ERR_UNSUPPORTED;
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * sizeof(float)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof(msg),
             "cuTensorMapEncodeTiled failed (CUresult %d) rows=%lld k=%lld ld=%lld", (int)r,
             (long long)rows, (long long)k, (long long)ld);
    set_last_error(__FILE__, __LINE__, msg);
    return OKGE_ERR_CUDA;
  }
  return OKGE_OK;
}

// K-panel operand: memory [panels][rows][32 floats]; 3-D map {32, rows, panels}, box = [1][box_rows][32].
int make_tmap_panel(CUtensorMap* out, const float* base, int64_t rows, int64_t k, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled not available from the driver");
    return OKGE_ERR_UNSUPPORTED;
  }
  const int64_t panels = ceil_div64(k, kBK);
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(kBK), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(panels)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(kBK) * sizeof(float),
                           static_cast<cuuint64_t>(rows) * kBK * sizeof(float)};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled (panel) failed (CUresult %d) rows=%lld k=%lld", (int)r,
             (long long)rows, (long long)k);
    set_last_error(__FILE__, __LINE__, msg);
    return OKGE_ERR_CUDA;
  }
  return OKGE_OK;
}

// MN-major operand stored "column-major": logical [rows, K] lives in memory as [K][ld] with the rows contiguous
// (a row-major matrix read as its own transpose). 2-D map {rows, K}; box = 32 rows x 32 K, 128B swizzle with 32-byte
// atoms; the kernel issues one box per 32-row block so that edges are clipped / zero-filled by the TMA unit.
int make_tmap_colmajor(CUtensorMap* out, const float* base, int64_t rows, int64_t k, int64_t ld, int stage_k) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled not available from the driver");
    return OKGE_ERR_UNSUPPORTED;
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(k)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * sizeof(float)};
  cuuint32_t box[2] = {32, static_cast<cuuint32_t>(stage_k)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled (col-major) failed (CUresult %d) rows=%lld k=%lld ld=%lld", (int)r,
             (long long)rows, (long long)k, (long long)ld);
    set_last_error(__FILE__, __LINE__, msg);
    return OKGE_ERR_CUDA;
  }
  return OKGE_OK;
}

// MN-major operand as a 3-D map {32 rows, K, ceil(rows/32) row blocks}; box = [box_rows/32][32 K][32]: one TMA per
// stage. MN-panels: memory [ceil(rows/32)][K][32 floats] (= the K-panel storage of the transposed matrix), strides
// (32, 32 K) floats. A col-major operand whose row count is a multiple of 32 is the same map with strides (ld, 32).
int make_tmap_mnpanel(CUtensorMap* out, const float* base, int64_t rows, int64_t k, int box_rows,
                      int64_t k_stride_floats, int64_t panel_stride_floats, int stage_k) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled not available from the driver");
    return OKGE_ERR_UNSUPPORTED;
  }
  const int64_t panels = ceil_div64(rows, 32);
  cuuint64_t dims[3] = {32, static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(panels)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(k_stride_floats) * sizeof(float),
                           static_cast<cuuint64_t>(panel_stride_floats) * sizeof(float)};
  cuuint32_t box[3] = {32, static_cast<cuuint32_t>(stage_k), static_cast<cuuint32_t>(box_rows / 32)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled (MN panel) failed (CUresult %d) rows=%lld k=%lld", (int)r,
             (long long)rows, (long long)k);
    set_last_error(__FILE__, __LINE__, msg);
    return OKGE_ERR_CUDA;
  }
  return OKGE_OK;
}

int make_operand_tmap(CUtensorMap* out, int mode, const float* base, int64_t rows, int64_t k, int64_t ld, int box_rows,
                      int stage_k) {
  switch (mode) {
    case OP_ROW_MAJOR: return make_tmap(out, base, rows, k, ld, box_rows);
    case OP_K_PANELS: return make_tmap_panel(out, base, rows, k, box_rows);
    case OP_COL_MAJOR: return make_tmap_colmajor(out, base, rows, k, ld, stage_k);
    case OP_MN_PANELS: return make_tmap_mnpanel(out, base, rows, k, box_rows, 32, k * 32, stage_k);
  }
  set_last_error(__FILE__, __LINE__, "unknown operand layout");
  return OKGE_ERR_INVALID;
}

// Output map of MODE_STORE: [splits][M][N] fp32 with row pitch ldc, box = 32 rows x 32 columns, SW128.
int make_tmap_out(CUtensorMap* out, float* base, int64_t M, int64_t N, int64_t ldc, int64_t splits,
                  int64_t split_stride, int box_cols = 32) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled not available from the driver");
    return OKGE_ERR_UNSUPPORTED;
  }
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(M), static_cast<cuuint64_t>(splits)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ldc) * sizeof(float),
                           static_cast<cuuint64_t>(splits > 1 ? split_stride : M * ldc) * sizeof(float)};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled (output) failed (CUresult %d) M=%lld N=%lld ldc=%lld", (int)r,
             (long long)M, (long long)N, (long long)ldc);
    set_last_error(__FILE__, __LINE__, msg);
    return OKGE_ERR_CUDA;
  }
  return OKGE_OK;
}

template <int MODE, bool LIMIT = false, bool RANK4 = false>
int launch_mode(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& td,
                const GemmParams& p, int grid, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    OKGE_CUDA_TRY(cudaFuncSetAttribute(okge_gemm_tf32_kernel<MODE, LIMIT, RANK4>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<MODE>::kSmemBytes));
    attr_set = true;
  }
  okge_gemm_tf32_kernel<MODE, LIMIT, RANK4><<<grid, Cfg<MODE>::kThreads, Cfg<MODE>::kSmemBytes, stream>>>(ta, tb, tc, td, p);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

bool is_mn_major(int mode) { return mode == OP_COL_MAJOR || mode == OP_MN_PANELS; }

int launch_gemm(int mode, const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M,
                int64_t N, int64_t K, GemmParams p, cudaStream_t stream) {
  OKGE_REQUIRE(A != nullptr && B != nullptr, "null operand");
  OKGE_REQUIRE(M > 0 && N > 0 && K > 0, "empty GEMM (M, N, K must be > 0)");
  OKGE_REQUIRE(M < INT_MAX && N < INT_MAX && K < INT_MAX, "dimension exceeds int32");
  OKGE_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15u) == 0 && (reinterpret_cast<uintptr_t>(B) & 15u) == 0,
               "operand base pointers must be 16-byte aligned (TMA)");
  OKGE_REQUIRE(p.a_mode != OP_ROW_MAJOR || (lda % 4 == 0 && lda >= K), "lda must be a multiple of 4 and >= K (TMA)");
  OKGE_REQUIRE(p.b_mode != OP_ROW_MAJOR || (ldb % 4 == 0 && ldb >= K), "ldb must be a multiple of 4 and >= K (TMA)");
  OKGE_REQUIRE(p.a_mode != OP_COL_MAJOR || (lda % 4 == 0 && lda >= M), "col-major lda must be a multiple of 4 and >= M (TMA)");
  OKGE_REQUIRE(p.b_mode != OP_COL_MAJOR || (ldb % 4 == 0 && ldb >= N), "col-major ldb must be a multiple of 4 and >= N (TMA)");
  int st = okge_device_check();
  if (st != OKGE_OK) return st;

  CUtensorMap ta, tb;
  // A col-major operand made of whole 32-row blocks is loaded with one strided 3-D TMA per stage instead of one box
  // per block (if the driver rejects that map the per-box form is used).
  const int stage_k = mode == MODE_ADAGRAD ? Cfg<MODE_ADAGRAD>::kStageK : kBK;
  OKGE_REQUIRE(stage_k == kBK || (is_mn_major(p.a_mode) && is_mn_major(p.b_mode)),
               "the fused Adagrad contraction takes MN-major operands (OKGE_COL_MAJOR / OKGE_MN_PANELS)");
  st = OKGE_ERR_INVALID;
  if (p.a_mode == OP_COL_MAJOR && M % 32 == 0 && (st = make_tmap_mnpanel(&ta, A, M, K, kBM, lda, 32, stage_k)) == OKGE_OK)
    p.a_mode = OP_MN_PANELS;
  if (st != OKGE_OK) st = make_operand_tmap(&ta, p.a_mode, A, M, K, lda, kBM, stage_k);
  if (st != OKGE_OK) return st;
  st = OKGE_ERR_INVALID;
  if (p.b_mode == OP_COL_MAJOR && N % 32 == 0 && (st = make_tmap_mnpanel(&tb, B, N, K, kBN, ldb, 32, stage_k)) == OKGE_OK)
    p.b_mode = OP_MN_PANELS;
  if (st != OKGE_OK) st = make_operand_tmap(&tb, p.b_mode, B, N, K, ldb, kBN, stage_k);
  if (st != OKGE_OK) return st;
  const bool a_mn = is_mn_major(p.a_mode), b_mn = is_mn_major(p.b_mode);
  const uint32_t kDescLoMnMajor = desc_lo_mn_major(stage_k);
  p.a_desc_lo = a_mn ? kDescLoMnMajor : kDescLoKMajor;
  p.a_desc_hi = a_mn ? kDescHiMnMajor : kDescHiKMajor;
  p.a_kadv = a_mn ? kKadvMnMajor : kKadvKMajor;
  p.b_desc_lo = b_mn ? kDescLoMnMajor : kDescLoKMajor;
  p.b_desc_hi = b_mn ? kDescHiMnMajor : kDescHiKMajor;
  p.b_kadv = b_mn ? kKadvMnMajor : kKadvKMajor;
  p.idesc = kInstrDesc | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.K = static_cast<int>(K);
  p.m_tiles = static_cast<int>(ceil_div64(M, kBM));
  p.n_tiles = static_cast<int>(ceil_div64(N, kBN));
  p.n_fastest = p.n_tiles < p.m_tiles;
  p.k_chunks = static_cast<int>(ceil_div64(K, stage_k));
  if (p.splits < 1) p.splits = 1;
  p.k_chunks_per_split = static_cast<int>(ceil_div64(p.k_chunks, p.splits));
  p.splits = static_cast<int>(ceil_div64(p.k_chunks, p.k_chunks_per_split));
  const long long total = static_cast<long long>(p.m_tiles) * p.n_tiles * p.splits;
  OKGE_REQUIRE(total < INT_MAX, "too many tiles");
  const int grid = static_cast<int>(total < sm_count() ? total : sm_count());

  CUtensorMap tc, td;
  memset(&tc, 0, sizeof(tc));
  memset(&td, 0, sizeof(td));
  if (mode == MODE_ADAGRAD) {
    // p.C = parameter rows, p.dS = Adagrad accumulator rows, both [M, N] with row pitch p.ldc
    OKGE_REQUIRE(((reinterpret_cast<uintptr_t>(p.C) | reinterpret_cast<uintptr_t>(p.dS)) & 15u) == 0 && p.ldc % 4 == 0,
                 "param / state must be 16-byte aligned with a row pitch that is a multiple of 4 (TMA)");
    st = make_tmap_out(&tc, p.C, M, N, p.ldc, 1, 0, kAdagradChunkCols);
    if (st != OKGE_OK) return st;
    st = make_tmap_out(&td, p.dS, M, N, p.ldc, 1, 0, kAdagradChunkCols);
    if (st != OKGE_OK) return st;
  } else if (mode == MODE_STORE) {
    OKGE_REQUIRE((reinterpret_cast<uintptr_t>(p.C) & 15u) == 0 && p.ldc % 4 == 0,
                 "output must be 16-byte aligned with a leading dimension that is a multiple of 4 (TMA store)");
    st = make_tmap_out(&tc, p.C, M, N, p.ldc, p.splits, p.split_stride);
    if (st != OKGE_OK) return st;
  } else if ((mode == MODE_BCE || mode == MODE_SMGRAD) && p.dS != nullptr) {
    // dS as K-panels [ceil(N/32)][M][32]: a "[panels][M][32]" output whose 32-column rows are whole 128-byte lines
    st = make_tmap_out(&tc, p.dS, M, 32, 32, ceil_div64(N, 32), M * 32, 16);
    if (st != OKGE_OK) return st;
  }
  switch (mode) {
    case MODE_STORE: return launch_mode<MODE_STORE>(ta, tb, tc, td, p, grid, stream);
    case MODE_BCE:
      if (p.thresh != nullptr) return launch_mode<MODE_BCE, false, true>(ta, tb, tc, td, p, grid, stream);
      return p.n_limit_dev != nullptr ? launch_mode<MODE_BCE, true>(ta, tb, tc, td, p, grid, stream)
                                      : launch_mode<MODE_BCE, false>(ta, tb, tc, td, p, grid, stream);
    case MODE_LSE: return launch_mode<MODE_LSE>(ta, tb, tc, td, p, grid, stream);
    case MODE_SMGRAD: return launch_mode<MODE_SMGRAD>(ta, tb, tc, td, p, grid, stream);
    case MODE_RANK: return launch_mode<MODE_RANK>(ta, tb, tc, td, p, grid, stream);
    case MODE_ADAGRAD: return launch_mode<MODE_ADAGRAD>(ta, tb, tc, td, p, grid, stream);
  }
  set_last_error(__FILE__, __LINE__, "unknown epilogue mode");
  return OKGE_ERR_INVALID;
}

}  // namespace

}  // namespace okge

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------

using namespace okge;

extern "C" int okge_gemm_tf32_nt(const float* A, int64_t lda, int32_t a_layout, const float* B, int64_t ldb,
                                 int32_t b_layout, int64_t M, int64_t N, int64_t K, float alpha,
                                 const float* alpha_dev, float* C, int64_t ldc, int32_t splits, float* split_ws,
                                 okge_stream_t stream) {
  OKGE_REQUIRE(C != nullptr, "null output");
  OKGE_REQUIRE(ldc >= N, "ldc smaller than N");
  OKGE_REQUIRE(a_layout >= OKGE_ROW_MAJOR && a_layout <= OKGE_MN_PANELS && b_layout >= OKGE_ROW_MAJOR &&
                   b_layout <= OKGE_MN_PANELS, "unknown operand layout");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GemmParams p = {};
  p.a_mode = a_layout;
  p.b_mode = b_layout;
  p.splits = splits;
  if (splits > 1) {
    OKGE_REQUIRE(split_ws != nullptr, "split-K needs a workspace of splits*M*N floats");
    p.C = split_ws;
    p.ldc = N;
    p.split_stride = M * N;
    p.alpha = 1.0f;
    p.alpha_dev = nullptr;
    int st = launch_gemm(MODE_STORE, A, lda, B, ldb, M, N, K, p, s);
    if (st != OKGE_OK) return st;
    // recompute the effective split count exactly as launch_gemm did
    const int64_t k_chunks = ceil_div64(K, kBK);
    const int64_t per = ceil_div64(k_chunks, splits);
    const int eff_splits = static_cast<int>(ceil_div64(k_chunks, per));
    const long long total = M * N;
    int blocks = static_cast<int>(ceil_div64(total, 256));
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    splitk_reduce_kernel<<<blocks, 256, 0, s>>>(split_ws, M * N, eff_splits, M, N, alpha, alpha_dev,
                                                C, ldc);
    OKGE_CUDA_TRY(cudaGetLastError());
    return OKGE_OK;
  }
  p.C = C;
  p.ldc = ldc;
  p.split_stride = 0;
  p.alpha = alpha;
  p.alpha_dev = alpha_dev;
  p.acc_scale = 1.0f;
  return launch_gemm(MODE_STORE, A, lda, B, ldb, M, N, K, p, s);
}

extern "C" int okge_score_store(const float* q, int64_t ldq, const float* e, int64_t lde, int64_t B,
                                int64_t N, int64_t D, float* scores, int64_t lds,
                                okge_stream_t stream) {
  // q is TF32-rounded by okge_fold_query; e is a raw table operand truncated by the tensor core
  return okge_gemm_tf32_nt(q, ldq, OKGE_ROW_MAJOR, e, lde, OKGE_ROW_MAJOR, B, N, D, kTf32RawOperandScale, nullptr,
                           scores, lds, 1, nullptr, stream);
}

namespace {

// one warp per CSR entry; the entry count lives on the device, so the grid is fixed and the warps stride
template <int MODE>
int launch_label_fix(const float* q, int64_t ldq, const float* e, int64_t lde, int64_t B, int64_t N, int64_t D,
                     const int32_t* pos_ptr, const int32_t* pos_idx, float y_pos, double* loss_sum, float y_delta,
                     float* dS, float* dST, float* pos_score, const float* row_lse, const float* row_weight,
                     cudaStream_t s) {
  if (pos_idx == nullptr) return OKGE_OK;   // no positives at all
  sparse_label_fix_kernel<MODE><<<sm_count() * 4, 256, 0, s>>>(
      q, ldq, e, lde, static_cast<int>(B), static_cast<int>(N), static_cast<int>(D), pos_ptr, pos_idx,
      kTf32RawOperandScale, y_pos, loss_sum, y_delta, dS, dST, pos_score, row_lse, row_weight);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

}  // namespace

extern "C" int okge_score_bce(const float* q, int64_t ldq, const float* e, int64_t lde, int64_t B,
                              int64_t N, int64_t D, const int32_t* pos_ptr, const int32_t* pos_idx,
                              float y_base, float y_pos, const int32_t* n_cols_dev, double* loss_sum, float* dS,
                              float* dST, okge_stream_t stream) {
  OKGE_REQUIRE(pos_ptr != nullptr && loss_sum != nullptr, "null label pointer / loss output");
  OKGE_REQUIRE(((reinterpret_cast<uintptr_t>(dS) | reinterpret_cast<uintptr_t>(dST)) & 127u) == 0,
               "dS / dST panels must be 128-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  OKGE_CUDA_TRY(cudaMemsetAsync(loss_sum, 0, sizeof(double), s));
  GemmParams p = {};
  p.splits = 1;
  p.y_base = y_base;
  p.n_limit_dev = n_cols_dev;
  p.loss_sum = loss_sum;
  p.dS = dS;
  p.dST = dST;
  p.acc_scale = kTf32RawOperandScale;
  int st = launch_gemm(MODE_BCE, q, ldq, e, lde, B, N, D, p, s);
  if (st != OKGE_OK) return st;
  return launch_label_fix<MODE_BCE>(q, ldq, e, lde, B, N, D, pos_ptr, pos_idx, y_pos, loss_sum, y_pos - y_base, dS, dST,
                                    nullptr, nullptr, nullptr, s);
}

extern "C" int okge_score_bce_rank(const float* q, int64_t ldq, const float* e, int64_t lde, int64_t B, int64_t B_extra,
                                   int64_t N, int64_t D, const int32_t* pos_ptr, const int32_t* pos_idx, float y_base,
                                   float y_pos, const float* thresh4, int32_t* greater4, int32_t* equal4, double* loss_sum,
                                   okge_stream_t stream) {
  OKGE_REQUIRE(pos_ptr != nullptr && loss_sum != nullptr, "null label pointer / loss output");
  OKGE_REQUIRE(B_extra >= 0, "negative number of extra rows");
  OKGE_REQUIRE(thresh4 != nullptr && greater4 != nullptr && equal4 != nullptr, "null ranking pointer");
  OKGE_REQUIRE((reinterpret_cast<uintptr_t>(thresh4) & 15u) == 0, "thresh4 must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  OKGE_CUDA_TRY(cudaMemsetAsync(loss_sum, 0, sizeof(double), s));
  GemmParams p = {};
  p.splits = 1;
  p.y_base = y_base;
  p.loss_sum = loss_sum;
  p.thresh = thresh4;
  p.greater = greater4;
  p.equal = equal4;
  p.loss_rows = static_cast<int>(B);
  p.acc_scale = kTf32RawOperandScale;
  int st = launch_gemm(MODE_BCE, q, ldq, e, lde, B + B_extra, N, D, p, s);
  if (st != OKGE_OK) return st;
  return launch_label_fix<MODE_BCE>(q, ldq, e, lde, B, N, D, pos_ptr, pos_idx, y_pos, loss_sum, y_pos - y_base, nullptr, nullptr,
                                    nullptr, nullptr, nullptr, s);
}

extern "C" int64_t okge_score_lse_ws_floats(int64_t B, int64_t N) {
  const int64_t P = ceil_div64(N, kBN) * kLseGroups;
  return 2 * P * B + 2 * static_cast<int64_t>(kLseChunks) * B;
}

extern "C" int okge_score_lse(const float* q, int64_t ldq, const float* e, int64_t lde, int64_t B,
                              int64_t N, int64_t D, const int32_t* pos_ptr, const int32_t* pos_idx,
                              float* row_lse, float* pos_score, float* part_ws,
                              okge_stream_t stream) {
  OKGE_REQUIRE(pos_ptr != nullptr && row_lse != nullptr && part_ws != nullptr, "null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t P = ceil_div64(N, kBN) * kLseGroups;
  GemmParams p = {};
  p.splits = 1;
  p.part_max = part_ws;
  p.part_sum = part_ws + P * B;
  p.acc_scale = kTf32RawOperandScale;
  int st = launch_gemm(MODE_LSE, q, ldq, e, lde, B, N, D, p, s);
  if (st != OKGE_OK) return st;
  float* s1max = part_ws + 2 * P * B;
  float* s1sum = s1max + static_cast<int64_t>(kLseChunks) * B;
  const int chunks = static_cast<int>(P < kLseChunks ? P : kLseChunks);
  dim3 g1(static_cast<unsigned>(ceil_div64(B, 128)), static_cast<unsigned>(chunks));
  lse_merge_stage1<<<g1, 128, 0, s>>>(p.part_max, p.part_sum, static_cast<int>(P),
                                      static_cast<int>(B), s1max, s1sum);
  OKGE_CUDA_TRY(cudaGetLastError());
  lse_merge_stage2<<<static_cast<unsigned>(ceil_div64(B, 128)), 128, 0, s>>>(
      s1max, s1sum, chunks, static_cast<int>(B), row_lse);
  OKGE_CUDA_TRY(cudaGetLastError());
  if (pos_score == nullptr) return OKGE_OK;
  return launch_label_fix<MODE_LSE>(q, ldq, e, lde, B, N, D, pos_ptr, pos_idx, 1.f, nullptr, 1.f, nullptr, nullptr,
                                    pos_score, nullptr, nullptr, s);
}

extern "C" int okge_score_softmax_grad(const float* q, int64_t ldq, const float* e, int64_t lde,
                                       int64_t B, int64_t N, int64_t D, const int32_t* pos_ptr,
                                       const int32_t* pos_idx, const float* row_lse,
                                       const float* row_weight, float* dS, float* dST,
                                       okge_stream_t stream) {
  OKGE_REQUIRE(pos_ptr != nullptr && row_lse != nullptr && row_weight != nullptr, "null pointer");
  OKGE_REQUIRE(((reinterpret_cast<uintptr_t>(dS) | reinterpret_cast<uintptr_t>(dST)) & 127u) == 0,
               "dS / dST panels must be 128-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GemmParams p = {};
  p.splits = 1;
  p.y_base = 0.f;
  p.row_lse = row_lse;
  p.row_weight = row_weight;
  p.dS = dS;
  p.dST = dST;
  p.acc_scale = kTf32RawOperandScale;
  int st = launch_gemm(MODE_SMGRAD, q, ldq, e, lde, B, N, D, p, s);
  if (st != OKGE_OK) return st;
  return launch_label_fix<MODE_SMGRAD>(q, ldq, e, lde, B, N, D, pos_ptr, pos_idx, 1.f, nullptr, 1.f, dS, dST, nullptr,
                                       row_lse, row_weight, s);
}

extern "C" int okge_score_rank(const float* q, int64_t ldq, const float* e, int64_t lde, int64_t Q,
                               int64_t N, int64_t D, const float* thresh, int32_t* greater,
                               int32_t* equal, okge_stream_t stream) {
  OKGE_REQUIRE(thresh != nullptr && greater != nullptr && equal != nullptr, "null pointer");
  GemmParams p = {};
  p.splits = 1;
  p.thresh = thresh;
  p.greater = greater;
  p.equal = equal;
  p.acc_scale = kTf32RawOperandScale;
  return launch_gemm(MODE_RANK, q, ldq, e, lde, Q, N, D, p, static_cast<cudaStream_t>(stream));
}

extern "C" int okge_gemm_adagrad(const float* A, int64_t lda, int32_t a_layout, const float* B, int64_t ldb,
                                 int32_t b_layout, int64_t M, int64_t N, int64_t K, float alpha, const float* alpha_dev,
                                 const int32_t* extra_map, const float* extra, int64_t ld_extra, float* param,
                                 float* state_sum, int64_t ld, float clr, float eps, float weight_decay,
                                 okge_stream_t stream) {
  OKGE_REQUIRE(param != nullptr && state_sum != nullptr, "null parameter / accumulator");
  OKGE_REQUIRE(ld >= N, "row pitch smaller than N");
  OKGE_REQUIRE(a_layout >= OKGE_ROW_MAJOR && a_layout <= OKGE_MN_PANELS && b_layout >= OKGE_ROW_MAJOR &&
                   b_layout <= OKGE_MN_PANELS, "unknown operand layout");
  OKGE_REQUIRE(extra_map == nullptr || extra != nullptr, "extra_map without extra rows");
  GemmParams p = {};
  p.a_mode = a_layout;
  p.b_mode = b_layout;
  p.splits = 1;
  p.C = param;
  p.dS = state_sum;
  p.ldc = ld;
  p.alpha = alpha;
  p.alpha_dev = alpha_dev;
  p.acc_scale = 1.0f;
  p.clr = clr;
  p.eps = eps;
  p.weight_decay = weight_decay;
  p.extra_map = extra_map;
  p.extra = extra;
  p.ld_extra = ld_extra;
  return launch_gemm(MODE_ADAGRAD, A, lda, B, ldb, M, N, K, p, static_cast<cudaStream_t>(stream));
}
