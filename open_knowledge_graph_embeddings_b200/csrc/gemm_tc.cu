// 1-vs-all scoring contractions for sm_100a: C[M,N] = A[M,K] * B[N,K]^T on the 5th-generation tensor cores.
//
//   * operands: FP16 (10-bit mantissa like TF32, rounded to nearest, power-of-two scaled by the producer, see
//     f16_ops.cu) for every 1-vs-all mode; TF32 straight from fp32 memory for the generic okge_gemm_tf32_nt. Moved by
//     TMA (cp.async.bulk.tensor, 128B swizzle) into a 4-stage shared-memory ring;
//   * math: tcgen05.mma.kind::f16 128x256x16 (kind::tf32: 128x256x8), issued by one elected thread, FP32 accumulators
//     in TMEM (2 x 256 columns, double buffered so the epilogue of tile i overlaps the MMAs of tile i+1);
//   * epilogue: 8 or 16 warps read TMEM with tcgen05.ld (one output row per thread, 32 columns per load) and apply one
//     of the fused epilogues below, so the B x N score matrix is never written unless the caller asks for it.
//
// Persistent kernel: grid = min(#work items, #SMs); work item = (m_tile, n_tile, k_split); the dimension with FEWER
// tiles runs fastest so that CTAs running at the same time share the large operand tile through L2.
//
// Reference semantics implemented by the epilogues (paths relative to the reference root):
//   MODE_STORE   q e^T                                   openkge/model.py:206-215, 270-272
//   MODE_BCE     BCEWithLogitsLoss(sum) and its gradient openkge/trainer.py:93-106
//   MODE_LSE     log_softmax(dim=1) row statistics       openkge/trainer.py:99-100
//   MODE_SMGRAD  gradient of KLDivLoss(log_softmax)      openkge/trainer.py:99-106
//   MODE_RANK    count-greater / count-equal             openkge/dataset.py:441-444
//   MODE_ADAGRAD dE = dS^T Q + torch.optim.Adagrad.step  utils/optim.py:194-201
//
// Why not ONE tile-resident kernel for the whole training step (S -> dS -> dE + Adagrad, dQ by reduction)? TMEM has
// 512 fp32 columns per SM: the dE tile of 128 entities x D = 512 fills all of them, the score tile of the same
// entities against 256 queries needs 256 more, and a dQ partial [B, D] = 1 MB per CTA has no on-chip home at all
// (flushing it per entity tile would be 8 GB of L2 reductions per step). See DESIGN.md section 3.
#include "okge_common.cuh"

#include <cuda_fp16.h>
#include <limits.h>
#include <math.h>
#include <string.h>

#include <type_traits>

namespace okge {

namespace {

constexpr int kBM = 128;       // tile rows    = UMMA M = TMEM lanes
constexpr int kBN = 256;       // tile columns = UMMA N = TMEM columns per accumulator
constexpr int kTmemCols = 512;                   // 2 accumulators x 256 columns
constexpr int kEpiStageBytes = 32 * 128;         // one 32-row x 32-column fp32 chunk per epilogue warp (SW128)

enum Mode : int { MODE_STORE = 0, MODE_BCE = 1, MODE_LSE = 2, MODE_SMGRAD = 3, MODE_RANK = 4, MODE_ADAGRAD = 5,
                  MODE_ADAGRAD_DEEP = 6 };
// MODE_ADAGRAD_DEEP: the same fused update for contractions with a long K (many query rows per table row: the sharded
// step at 4 / 8 GPUs), where the kernel is tensor-bound and the operand ring, not the p / G staging, needs the bytes.
__host__ __device__ constexpr bool is_adagrad(int mode) { return mode == MODE_ADAGRAD || mode == MODE_ADAGRAD_DEEP; }

// Operand source forms (a_mode / b_mode). 0/1 are K-major in shared memory, 2/3 MN-major; see okge_b200.h.
enum OperandMode : int { OP_ROW_MAJOR = OKGE_ROW_MAJOR, OP_K_PANELS = OKGE_K_PANELS, OP_COL_MAJOR = OKGE_COL_MAJOR,
                         OP_MN_PANELS = OKGE_MN_PANELS };

// Element-type constants: a 128-byte swizzle row holds kRow elements (the K depth of a K-major stage and the width
// of a panel); one tcgen05.mma covers 32 bytes of K.
template <bool F16>
struct Elem {
  static constexpr int kBytes = F16 ? 2 : 4;
  static constexpr int kRow = 128 / kBytes;      // 64 (fp16) / 32 (tf32)
  static constexpr int kUmmaK = 32 / kBytes;     // 16 / 8
};

// Per-epilogue kernel shape. The loss epilogues do ~20 instructions per score, so they get 16 epilogue warps
// (4 per scheduler) and a 2 KiB TMA-store staging buffer per warp (a 32 x 32 fp16 chunk of dS, SWIZZLE_64B).
template <bool F16, int MODE>
struct Cfg {
  using E = Elem<F16>;
  static constexpr bool kStaged = MODE == MODE_STORE || MODE == MODE_BCE || MODE == MODE_SMGRAD || is_adagrad(MODE);
  static constexpr bool kHalfOut = MODE == MODE_BCE || MODE == MODE_SMGRAD;      // fp16 dS chunks
  static constexpr int kEpiWarps = (MODE == MODE_STORE || is_adagrad(MODE)) ? 8 : 16;
  // MODE_ADAGRAD streams the parameter and its accumulator through shared memory (2 x (4 + 4) KiB per warp, loads one
  // chunk ahead) and is HBM-bound, so it gives up half of every pipeline stage for that staging. Its operands are
  // MN-major (dS^T panels, Q column-major), whose K extent per stage is free.
  // HBM-bound shape (K <= ~1,024): 4 stages of 24 KiB and 32-column p / G chunks (2 x 8 KiB per warp). Tensor-bound shape
  // (MODE_ADAGRAD_DEEP): the 96 KiB ring covers only ~1,000 tensor-pipe cycles, less than an L2 round trip under load
  // (ncu at K = 4,096: tensor pipe 40 %, epilogue warps 58 % of their time waiting for an accumulator), so the p / G
  // chunks shrink to 16 columns (2 x 4 KiB per warp) and the ring grows to 144 KiB -- as 3 stages of 48 KiB: the issuer's
  // per-stage cost (barrier wait, commit) is then spread over four MMAs instead of two (6 x 24 KiB: 0.620 ms at
  // K = 4,096, 3 x 48 KiB: 0.599).
  static constexpr int kStageK = MODE == MODE_ADAGRAD ? E::kRow / 2 : E::kRow;
  static constexpr int kUpdCols = MODE == MODE_ADAGRAD_DEEP ? 16 : 32;          // columns of a staged p / G chunk
  static constexpr int kABytes = kBM * kStageK * E::kBytes;
  static constexpr int kBBytes = kBN * kStageK * E::kBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = MODE == MODE_ADAGRAD_DEEP ? 3 : 4;
  static constexpr int kGroups = kEpiWarps / 4;            // column groups of the 256-column accumulator
  static constexpr int kColsPerGroup = kBN / kGroups;
  static constexpr int kThreads = 32 * (2 + kEpiWarps);
  static constexpr int kEpiWarpBytes = is_adagrad(MODE) ? 2 * (2 * 32 * kUpdCols * 4)
                                                          : (kHalfOut ? kEpiStageBytes / 2 : kEpiStageBytes);
  static constexpr int kEpiBytes = kStaged ? kEpiWarps * kEpiWarpBytes : 0;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align slack*/ + 512 /*barriers*/;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KiB of shared memory a CTA can opt into");
};
constexpr int kLseGroups = Cfg<true, MODE_LSE>::kGroups;

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, splits;
  int n_fastest;           // work-item order, see decode_work
  int k_chunks, k_chunks_per_split;
  int terms;               // 1, or 3 = split-precision product A_hi B_hi + A_hi B_lo + A_lo B_hi (row-major operands)
  int a_mode, b_mode;      // OperandMode
  // shared-memory matrix descriptors of the two operands: lo = start address field | desc_lo, per UMMA K step += kadv
  uint32_t a_desc_lo, a_desc_hi, a_kadv;
  uint32_t b_desc_lo, b_desc_hi, b_kadv;
  uint32_t idesc;
  // scale of the raw accumulator: host factor x up to three device scalars (operand inverse scales, gradient scale)
  float alpha;
  const float* scale_dev[3];
  // MODE_STORE
  float* C;
  long long ldc;
  long long split_stride;  // elements between split partials (0 when splits == 1)
  // labels (BCE, SMGRAD): every label is y_base inside the tiles; positives are fixed up by sparse_label_fix_kernel
  float y_base;
  const int* n_limit_dev;  // BCE: columns >= *n_limit_dev are padding (no loss term, zero gradient); nullable
  int loss_rows;           // BCE + RANK: rows >= loss_rows only carry ranking thresholds (no loss term)
  int rank_slots;          // BCE + RANK: thresholds per row the kernel looks at (1, 2 or 4)
  double* loss_sum;
  __half* dS;              // K-panel layout of the [M, N] gradient: [ceil(N/64)][M][64] fp16 (TMA store through tmap_c)
  float ds_scale;          // dS is stored as fp16(ds_scale * gradient); a power of two, ds_log2 = log2(ds_scale)
  float ds_log2;
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
  float* param;
  float* state;
  float clr, eps, weight_decay;
  const int* extra_map;    // [M] slot of an additional gradient row per output row, -1 = none (nullable)
  const float* extra;      // [slots, N] row-major
  long long ld_extra;
  // ADAGRAD with LIMIT = true: the contraction is the gradient of a DROPPED-OUT operand (inverted dropout of the candidate
  // rows, openkge/model.py:461-470 applied by _get_all); the mask dropout_kernel would draw for (p, seed, offset [, step])
  // over the flattened [M, N] matrix is applied to the gradient tile before the update
  float drop_p, drop_scale;
  unsigned long long drop_seed, drop_offset;
  const unsigned long long* drop_step_dev;
  __half* shadow;          // fp16 copy of the updated parameter rows (the next step's scoring operand); nullable
  long long ld_shadow;
  const float* shadow_inv;  // device scalar: 1 / scale of the fp16 copy
};

// Shared-memory matrix descriptors (PTX "tcgen05 shared memory descriptor", version 1 = Blackwell).
//   K-major, SWIZZLE_128B (layout type 2): rows of 128 bytes, 8-row groups 1024 B apart (SBO); LBO unused (1).
//     One UMMA K step (32 bytes) advances the start address by 32 B  -> +2 in the (addr >> 4) field.
//   MN-major fp16, SWIZZLE_128B (layout type 2): each K index is a 128-byte row holding 64 consecutive MN elements;
//     8-row groups are 1024 B apart (SBO), 64-element MN blocks one 64 x stage_k box apart (LBO).
//     One UMMA K step (16 rows = 2048 B) -> +128.
//   MN-major tf32, SWIZZLE_128B with 32-byte atoms (layout type 1, the only MN-major form for 4-byte operands): a K
//     index is a 128-byte row of 32 MN elements; 4-row groups 512 B apart (SBO), 32-element MN blocks one 32 x stage_k
//     box apart (LBO). One UMMA K step (8 rows = 1024 B) -> +64.
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t kDescLoKMajor = 1u << 16;
constexpr uint32_t kKadvKMajor = 2;
constexpr uint32_t kDescHiMnMajorTf32 = (512u >> 4) | (1u << 14) | (1u << 29);
constexpr uint32_t desc_lo_mn_major(int stage_k) { return (static_cast<uint32_t>(128 * stage_k) >> 4) << 16; }
constexpr uint32_t kKadvMnMajorTf32 = 1024u >> 4;
constexpr uint32_t kKadvMnMajorF16 = 2048u >> 4;

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lo_or, uint32_t hi) {
  return (static_cast<uint64_t>(hi) << 32) | static_cast<uint64_t>(lo_or | ((smem_addr >> 4) & 0x3FFFu));
}

// Instruction descriptor: D = F32 (bit 4), A / B format at bits 7 / 10 (0 = F16, 2 = TF32), bit 15 / 16 = A / B is
// MN-major, N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t instr_desc(bool f16) {
  return (1u << 4) | ((f16 ? 0u : 2u) << 7) | ((f16 ? 0u : 2u) << 10) | (static_cast<uint32_t>(kBN >> 3) << 17) |
         (static_cast<uint32_t>(kBM >> 4) << 24);
}

struct WorkItem {
  int m, n, split;
};

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
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;

// two fp32 -> packed fp16x2 (lo = first), round to nearest even, saturating to the largest finite value
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ float param_scale(const GemmParams& p) {
  float s = p.alpha;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (p.scale_dev[i] != nullptr) s *= __ldg(p.scale_dev[i]);
  return s;
}

// One load of a pipeline stage operand: `rows` rows (128 for A, 256 for B) of K chunk `kc`, in whichever form the
// operand has. `plane` selects the hi / lo plane of a row-major fp16 operand.
template <bool F16, int kStageK>
__device__ __forceinline__ void load_operand(int mode, uint32_t dst, const CUtensorMap* tm, uint32_t bar, int row0,
                                             int rows, int kc, int plane) {
  constexpr int kRow = Elem<F16>::kRow;
  if (mode == OP_ROW_MAJOR) {
    if constexpr (F16) tma_load_3d(dst, tm, bar, kc * kStageK, row0, plane);
    else tma_load_2d(dst, tm, bar, kc * kStageK, row0);
  } else if (mode == OP_K_PANELS) {
    tma_load_3d(dst, tm, bar, 0, row0, kc);
  } else if (mode == OP_MN_PANELS) {
    tma_load_3d(dst, tm, bar, 0, kc * kStageK, row0 / kRow);
  } else {  // OP_COL_MAJOR: one kRow x kStageK box per MN block (boxes past the matrix edge are zero-filled)
    for (int i = 0; i < rows / kRow; ++i)
      tma_load_2d(dst + static_cast<uint32_t>(i * 128 * kStageK), tm, bar, row0 + kRow * i, kc * kStageK);
  }
}

template <bool F16>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if constexpr (F16) umma_f16(tmem_d, adesc, bdesc, idesc, acc);
  else umma_tf32(tmem_d, adesc, bdesc, idesc, acc);
}

// LIMIT (MODE_BCE only): the number of label-carrying columns is read from p.n_limit_dev; a separate instantiation so
// that the regular loss kernel, which sits at its register cap, is compiled without it.
// RANK = 1, 2 or 4 (MODE_BCE only, evaluation): the loss pass also counts, for up to RANK ranked answers per query row, the
// scores above / equal to the answer's threshold (p.thresh [M, 4], +inf = unused slot; p.greater / p.equal [M, 4]; slots
// >= RANK are ignored) -- the filtered ranking of openkge/dataset.py:441-444 without a second contraction over the
// candidates. Every slot costs four instructions per score, so the host picks the smallest RANK that fits the batch.
template <bool F16, int MODE, bool LIMIT = false, int RANK = 0>
__global__ void __launch_bounds__(Cfg<F16, MODE>::kThreads, 1)
okge_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c,
                    const __grid_constant__ CUtensorMap tmap_d, const GemmParams p) {
  static_assert(F16 || MODE == MODE_STORE, "the fused epilogues take fp16 operands");
  using C = Cfg<F16, MODE>;
  constexpr bool RANK4 = RANK > 0;             // evaluation variant of the loss epilogue (no gradient)
  constexpr int kStages = C::kStages;
  constexpr int kStageK = C::kStageK;
  constexpr int kABytes = C::kABytes;
  constexpr int kStageBytes = C::kStageBytes;
  constexpr int kNumEpiWarps = C::kEpiWarps;
  constexpr int kColsPerGroup = C::kColsPerGroup;
  constexpr int kUmmaK = Elem<F16>::kUmmaK;
  extern __shared__ uint8_t smem_raw[];
  // swizzled tiles need 1024-byte alignment.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + kStages * kStageBytes;
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
  const int all_chunks = p.k_chunks * p.terms;     // K chunks of one output tile, over all terms
  // Warp roles: epilogue warps 0 .. kNumEpiWarps-1 (TMEM lane quarter = warp % 4), then the TMA producer and the MMA
  // issuer as the LAST two warps: the warp scheduler favours higher warp ids, and these two single-thread roles must
  // never wait behind busy epilogue warps for an issue slot (with them as warps 0 / 1 the loss epilogue starved them).
  constexpr int kProducerWarp = kNumEpiWarps, kMmaWarp = kNumEpiWarps + 1;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (C::kStaged) tma_prefetch_desc(&tmap_c);
    if (is_adagrad(MODE)) {
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
  // everything above touched only this CTA's shared / tensor memory and the kernel parameters: it may overlap the tail of
  // the preceding kernel. From here on global memory is read and written.
  pdl_wait_and_trigger();

  if (warp == kProducerWarp) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const WorkItem it = decode_work(w, p);
        const int kc_begin = it.split * p.k_chunks_per_split;
        const int kc_end = min(kc_begin + p.k_chunks_per_split, all_chunks);
        for (int kk = kc_begin; kk < kc_end; ++kk) {
          // split-precision products walk K three times: (A_hi, B_hi), (A_hi, B_lo), (A_lo, B_hi)
          const int term = p.terms == 1 ? 0 : kk / p.k_chunks;
          const int kc = p.terms == 1 ? kk : kk - term * p.k_chunks;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), kStageBytes);
          const uint32_t sa = smem_base + stage * kStageBytes;
          load_operand<F16, kStageK>(p.a_mode, sa, &tmap_a, full_bar(stage), it.m * kBM, kBM, kc, term == 2 ? 1 : 0);
          load_operand<F16, kStageK>(p.b_mode, sa + kABytes, &tmap_b, full_bar(stage), it.n * kBN, kBN, kc, term == 1 ? 1 : 0);
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
        const int kc_end = min(kc_begin + p.k_chunks_per_split, all_chunks);
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
            umma<F16>(tmem_d, adesc + static_cast<uint64_t>(p.a_kadv * k),
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
  } else if constexpr (is_adagrad(MODE)) {
    // ===================== epilogue: Adagrad step fused onto the gradient tile =====================
    // g = alpha * acc (+ extra row); g' = g + wd p; G += g'^2; p -= clr g' / (sqrt(G) + eps). Each warp walks its
    // chunks (32 rows x 32 columns) in order. p and G of the next kAhead chunks are in flight (TMA, 128-byte swizzle)
    // while one is updated in shared memory and TMA-stored back: the tables move through HBM exactly once each way.
    // The fp16 copy of the new parameter values (the scoring operand of the next step) leaves straight from registers:
    // every lane owns 64 contiguous bytes of its row.
    constexpr int kCols = C::kUpdCols;               // columns per chunk: 128-byte rows / SWIZZLE_128B, or 64 / SWIZZLE_64B
    constexpr int kBufBytes = 2 * 32 * kCols * 4;    // p | G
    constexpr int kBufs = C::kEpiWarpBytes / kBufBytes, kAhead = kBufs - 1;
    constexpr int kChunks = kColsPerGroup / kCols;
    const int ew = warp;
    const int quarter = warp & 3;
    const int group = ew >> 2;
    const uint32_t wbuf = epi_base + static_cast<uint32_t>(ew * C::kEpiWarpBytes);
    const uint32_t ldbar = bar_base + 8u * (2 * kStages + 5 + 4 * ew);                // one mbarrier per buffer (<= 4)
    const float alpha_eff = param_scale(p);
    const float clr = p.clr, eps = p.eps, wd = p.weight_decay;
    const float hs = p.shadow != nullptr ? 1.0f / __ldg(p.shadow_inv) : 1.0f;     // power of two: exact
    unsigned long long drop_offset = 0;
    if constexpr (LIMIT) drop_offset = p.drop_offset + (p.drop_step_dev != nullptr ? (*p.drop_step_dev << 44) : 0ull);
    auto n_valid = [&](const WorkItem& it) {
      const int rem = p.N - (it.n * kBN + group * kColsPerGroup);
      return rem <= 0 ? 0 : min(kChunks, (rem + kCols - 1) / kCols);
    };
    int pw = blockIdx.x, pc = 0;          // prefetch cursor: (work item, chunk) of the next load
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
        // this buffer was last read by the store of chunk n_issued - kBufs; only the kBufs - kAhead stores committed
        // after that one may still be draining
        tma_store_wait_read_le<kBufs - kAhead>();
        mbar_arrive_expect_tx(bar, kBufBytes);
        const int c0 = it.n * kBN + group * kColsPerGroup + pc * kCols, r0 = it.m * kBM + quarter * 32;
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
        const int col0 = it.n * kBN + group * kColsPerGroup + chunk * kCols;
        uint32_t v[kCols];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(acc * kBN + group * kColsPerGroup + chunk * kCols);
        tmem_ld_chunk(taddr, v);
        tmem_ld_wait();
        float g[kCols];
#pragma unroll
        for (int t = 0; t < kCols; ++t) g[t] = alpha_eff * __uint_as_float(v[t]);
        if constexpr (LIMIT) {   // gradient through the dropout of the candidate rows: d raw = mask / (1 - p) * d dropped
          const uint64_t first4 = (static_cast<uint64_t>(row) * static_cast<uint64_t>(p.N) + static_cast<uint64_t>(col0)) >> 2;
#pragma unroll
          for (int c = 0; c < kCols / 4; ++c) {
            const float4 m = dropout4(make_float4(g[4 * c], g[4 * c + 1], g[4 * c + 2], g[4 * c + 3]), first4 + c, p.drop_p,
                                      p.drop_scale, p.drop_seed, drop_offset);
            g[4 * c] = m.x; g[4 * c + 1] = m.y; g[4 * c + 2] = m.z; g[4 * c + 3] = m.w;
          }
        }
        if (slot >= 0) {   // rare (the batch's own entities): add the lookup gradient row of this table row
          const float* ex = p.extra + static_cast<long long>(slot) * p.ld_extra + col0;
#pragma unroll
          for (int t = 0; t < kCols; ++t)
            if (col0 + t < p.N) g[t] += __ldg(ex + t);
        }
        const int buf = n_done % kBufs;
        const uint32_t pb = wbuf + static_cast<uint32_t>(buf * kBufBytes);
        mbar_wait(ldbar + 8u * buf, static_cast<uint32_t>((n_done / kBufs) & 1));
#pragma unroll
        for (int c = 0; c < kCols / 4; ++c) {
          // row = lane; 16-byte chunk c of the row sits at the swizzled position of the TMA layout
          const uint32_t off = kCols == 32
              ? static_cast<uint32_t>(lane) * 128u + (static_cast<uint32_t>(c ^ (lane & 7)) << 4)
              : static_cast<uint32_t>(lane) * 64u + (static_cast<uint32_t>(c ^ ((lane >> 1) & 3)) << 4);
          float pv[4], sv[4];
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pv[0]), "=f"(pv[1]), "=f"(pv[2]), "=f"(pv[3]) : "r"(pb + off));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sv[0]), "=f"(sv[1]), "=f"(sv[2]), "=f"(sv[3])
                       : "r"(pb + kBufBytes / 2 + off));
#pragma unroll
          for (int t = 0; t < 4; ++t) adagrad_elem_fast(pv[t], g[4 * c + t], sv[t], clr, eps, wd);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(pb + off), "f"(pv[0]), "f"(pv[1]), "f"(pv[2]), "f"(pv[3]) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(pb + kBufBytes / 2 + off), "f"(sv[0]), "f"(sv[1]),
                       "f"(sv[2]), "f"(sv[3]) : "memory");
          // the gradient registers are dead now: reuse two of them for the packed fp16 copy of the new values
          v[2 * c] = pack_half2(pv[0] * hs, pv[1] * hs);
          v[2 * c + 1] = pack_half2(pv[2] * hs, pv[3] * hs);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmap_c, pb, col0, row0, 0);
          tma_store_3d(&tmap_d, pb + kBufBytes / 2, col0, row0, 0);
          tma_store_commit();
        }
        if (p.shadow != nullptr && row < p.M) {
          __half* dst = p.shadow + static_cast<long long>(row) * p.ld_shadow + col0;
          const int ncols = p.N - col0;                     // multiple of 8 (checked on the host)
#pragma unroll
          for (int j = 0; j < kCols / 8; ++j)
            if (8 * j < ncols)
              *reinterpret_cast<uint4*>(dst + 8 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
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
    const float acc_scale = param_scale(p);     // scores = acc_scale * accumulator (MODE_STORE: alpha)

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

    // a 32 x 32 chunk of fp16 gradients (16 packed words per row) through a 2 KB buffer: 64-byte rows, SWIZZLE_64B
    // (16-byte chunk index XOR bits 7..8 of the address = (lane >> 1) & 3); c0 = first column inside the 64-wide panel
    const uint64_t stream_policy = l2_policy_evict_first();
    auto stage_and_store_half = [&](const uint32_t (&h)[16], int c0, int c1, int c2) {
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t addr = stage_buf + static_cast<uint32_t>(lane) * 64u + (static_cast<uint32_t>(c ^ ((lane >> 1) & 3)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(h[4 * c + 0]), "r"(h[4 * c + 1]),
                     "r"(h[4 * c + 2]), "r"(h[4 * c + 3])
                     : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d_hint(&tmap_c, stage_buf, c0, c1, c2, stream_policy);   // written once, read by the next kernels
        tma_store_commit();
      }
    };

    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const WorkItem it = decode_work(w, p);
      const int row = it.m * kBM + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const int cbeg = it.n * kBN + group * kColsPerGroup;   // first column of this warp group

      // ---- per-row prologue ----
      float row_lse = 0.f, row_w = 0.f, thr = 0.f;
      if (MODE == MODE_SMGRAD) {
        if (row_ok) { row_lse = __ldg(p.row_lse + row) * kLog2e; row_w = __ldg(p.row_weight + row) * p.ds_scale; }
      }
      if (MODE == MODE_RANK) {
        if (row_ok) thr = __ldg(p.thresh + row);
      }
      constexpr int kSlots = RANK > 0 ? RANK : 1;
      float thrs[kSlots];
      int cgs[kSlots], ces[kSlots];
#pragma unroll
      for (int k = 0; k < kSlots; ++k) { thrs[k] = INFINITY; cgs[k] = 0; ces[k] = 0; }       // +inf never counts
      if (RANK4) {
        if (row_ok) {
#pragma unroll
          for (int k = 0; k < kSlots; ++k) thrs[k] = __ldg(p.thresh + 4 * static_cast<long long>(row) + k);
        }
      }
      float run_max = -INFINITY, run_sum = 0.f;   // LSE
      float tile_loss = 0.f;                      // BCE: fp32 inside a tile (<= 128 columns per thread), fp64 across tiles
      int cnt_g = 0, cnt_e = 0;                   // RANK

      mbar_wait(tmem_full_bar(acc), acc_phase);
      tcgen05_fence_after();

#pragma unroll 1
      for (int chunk = 0; chunk < kColsPerGroup / 32; ++chunk) {
        const int col0 = cbeg + chunk * 32;
        // warp-uniform. The gradient panels are 64 columns wide: the chunk behind the last column of an odd 32-column
        // block is still written (as zeros), because the dQ contraction reads whole panels.
        if (col0 >= (C::kHalfOut ? ((p.N + 63) & ~63) : p.N)) break;
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(acc * kBN + group * kColsPerGroup + chunk * 32);
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        // Columns that carry labels (warp-uniform): all N, or fewer when the caller padded the candidate list to a fixed
        // capacity and keeps the real count in device memory (CUDA-graph replay of batch-shared candidate lists); <= 0
        // when the whole chunk is padding.
        int ncols = min(32, p.N - col0);
        if (LIMIT) ncols = min(ncols, __ldg(p.n_limit_dev) - col0);

        // `full` is a compile-time tag: the edge chunk (ncols < 32) runs a masked copy of the same body
        auto body = [&](auto full) {
          constexpr bool kFull = decltype(full)::value;
          if constexpr (MODE == MODE_STORE) {
#pragma unroll
            for (int t = 0; t < 32; ++t) v[t] = __float_as_uint(acc_scale * __uint_as_float(v[t]));
            stage_and_store(v, col0, it.m * kBM + quarter * 32, it.split);
          } else if constexpr (MODE == MODE_BCE || MODE == MODE_SMGRAD) {
            // Dense part only: every label is y_base here; the (very sparse) positives are corrected afterwards by
            // sparse_label_fix_kernel, which keeps all CSR look-ups off this epilogue.
            float lsum = 0.f;
            uint32_t h[16];
            if constexpr (MODE == MODE_BCE) {
              // s = c * raw (c: inverse operand scales). With e = exp(-|s|) and r = 1 / (1 + e):
              //   softplus(s) = max(s, 0) - ln r,    sigmoid(s) = r (s >= 0) or e r.
              // Two MUFU ops per score (ex2, rcp) plus one lg2 per 16 scores: -ln r = ln(1 + e) is summed as the log of a
              // running product of (1 + e) in (1, 2]. The two sums of the loss (max part, log part; with label
              // smoothing also the plain sum of s) are kept apart and combined once per chunk. The gradient scale gs
              // (a power of two) rides in the exponent argument of ex2: eg = gs e.
              const float gs = RANK4 ? 1.0f : p.ds_scale;
              const float k_exp = -acc_scale * kLog2e, gs_log2 = RANK4 ? 0.0f : p.ds_log2, inv_gs = 1.0f / gs;
              float acc_max = 0.f, acc_lg = 0.f, acc_raw = 0.f, prod = 1.f;
              auto scores = [&](auto smooth_tag) {
                constexpr bool kSmooth = decltype(smooth_tag)::value;
                const float gy0 = p.y_base * p.ds_scale;
#pragma unroll
                for (int t = 0; t < 32; t += 2) {
                  float gp[2];
#pragma unroll
                  for (int u = 0; u < 2; ++u) {
                    const float raw = __uint_as_float(v[t + u]);
                    const float eg = ex2_approx(fmaf(fabsf(raw), k_exp, gs_log2));   // gs exp(-|s|) in (0, gs]
                    const float one_e = fmaf(eg, inv_gs, 1.f);                       // 1 + e in (1, 2]
                    const float r = rcp_approx(one_e);                               // 1 / (1 + e) in [1/2, 1)
                    if (kFull || t + u < ncols) {
                      acc_max += fmaxf(raw, 0.f);
                      prod *= one_e;                                   // sum of logs = log of the product (<= 2^16)
                      if (kSmooth) acc_raw += raw;
                    }
                    if (RANK4) {
                      const float s = raw * acc_scale;                 // the score exactly as MODE_RANK forms it
                      if (kFull || t + u < ncols) {
#pragma unroll
                        for (int k = 0; k < kSlots; ++k) {
                          cgs[k] += (thrs[k] < s) ? 1 : 0;
                          ces[k] += (thrs[k] == s) ? 1 : 0;
                        }
                      }
                    } else {                                           // the gradient is not needed in evaluation
                      float g = ((raw >= 0.f) ? gs : eg) * r;          // gs sigmoid(s)
                      if (kSmooth) g -= gy0;
                      gp[u] = (kFull || t + u < ncols) ? g : 0.f;
                    }
                  }
                  if (!RANK4) h[t >> 1] = pack_half2(gp[0], gp[1]);
                  if ((t & 15) == 14) {                                // one lg2 per 16 scores
                    acc_lg -= lg2_approx(prod);
                    prod = 1.f;
                  }
                }
              };
              if (p.y_base == 0.f) scores(std::false_type{}); else scores(std::true_type{});
              lsum = fmaf(acc_scale, acc_max, -0.6931471805599453f * acc_lg) - acc_scale * p.y_base * acc_raw;
            } else {
              const float k2 = acc_scale * kLog2e, gy0 = p.y_base * p.ds_scale;
#pragma unroll
              for (int t = 0; t < 32; t += 2) {
                const float g0 = row_w * ex2_approx(fmaf(__uint_as_float(v[t]), k2, -row_lse)) - gy0;
                const float g1 = row_w * ex2_approx(fmaf(__uint_as_float(v[t + 1]), k2, -row_lse)) - gy0;
                h[t >> 1] = pack_half2((kFull || t < ncols) ? g0 : 0.f, (kFull || t + 1 < ncols) ? g1 : 0.f);
              }
            }
            if (MODE == MODE_BCE && row_ok && (!RANK4 || row < p.loss_rows)) tile_loss += lsum;
            // dS panel = 64 columns x all rows: this chunk is rows [row0, row0 + 32), columns col0 % 64 .. + 32 of panel col0 / 64
            if (!RANK4 && p.dS != nullptr) stage_and_store_half(h, col0 & 63, it.m * kBM + quarter * 32, col0 >> 6);
          } else if constexpr (MODE == MODE_LSE) {
            float cmax = -INFINITY;
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              v[t] = __float_as_uint(__uint_as_float(v[t]) * (acc_scale * kLog2e));   // base-2 domain
              if (kFull || t < ncols) cmax = fmaxf(cmax, __uint_as_float(v[t]));
            }
            const float new_max = fmaxf(run_max, cmax);
            float csum = 0.f;
#pragma unroll
            for (int t = 0; t < 32; ++t)
              if (kFull || t < ncols) csum += ex2_approx(__uint_as_float(v[t]) - new_max);
            run_sum = run_sum * ex2_approx(run_max - new_max) + csum;   // 2^(-inf) = 0 on the first chunk
            run_max = new_max;
          } else if constexpr (MODE == MODE_RANK) {
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              const float s = __uint_as_float(v[t]) * acc_scale;
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
#pragma unroll
          for (int k = 0; k < kSlots; ++k) {
            if (cgs[k]) atomicAdd(g4 + k, cgs[k]);
            if (ces[k]) atomicAdd(e4 + k, ces[k]);
          }
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
//   s = <q[b, :], e[n, :]> from the SAME fp16 operands (and inverse scales) as the tensor-core pass,
//   MODE_BCE     loss -= s * (y_pos - y_base);  dS[b, n] = sigmoid(s) - y_pos
//   MODE_SMGRAD  dS[b, n] = row_weight[b] * exp(s - row_lse[b]) - y_pos
//   MODE_LSE     pos_score[p] = s
// q_lo / e_lo (nullable): split-precision planes, s = q_hi e_hi + q_hi e_lo + q_lo e_hi like the 3-term contraction.
template <int MODE>
__global__ void sparse_label_fix_kernel(const __half* __restrict__ q, const __half* __restrict__ q_lo, long long ldq,
                                        const __half* __restrict__ e, const __half* __restrict__ e_lo, long long lde,
                                        int B, int N, int D, const int* __restrict__ pos_ptr,
                                        const int* __restrict__ pos_idx, const float* __restrict__ q_inv,
                                        const float* __restrict__ e_inv, float y_pos, double* loss_sum, float y_delta,
                                        __half* dS, float ds_scale, float* pos_score, const float* __restrict__ row_lse,
                                        const float* __restrict__ row_weight) {
  pdl_wait_and_trigger();
  const int lane = threadIdx.x & 31;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int nnz = __ldg(pos_ptr + B);
  float acc_scale = 1.f;
  if (q_inv != nullptr) acc_scale *= __ldg(q_inv);
  if (e_inv != nullptr) acc_scale *= __ldg(e_inv);
  const bool split = q_lo != nullptr && e_lo != nullptr;
  double loss_local = 0.0;
  for (int pidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; pidx < nnz; pidx += warps_total) {
    // row b with pos_ptr[b] <= pidx < pos_ptr[b + 1]: the last b whose pointer is <= pidx. 32-ary search, one probe per
    // lane and round (the probes that satisfy the test form a prefix): 3 dependent loads for B = 4,096 instead of 12
    int lo = 0, hi = B;
    while (hi - lo > 1) {
      const int step = (hi - lo + 31) >> 5;
      const int probe = lo + (lane + 1) * step;
      const bool le = probe < hi && __ldg(pos_ptr + probe) <= pidx;
      const int cnt = __popc(__ballot_sync(0xffffffffu, le));
      const int nlo = lo + cnt * step;
      hi = min(hi, nlo + step);
      lo = nlo;
    }
    const int b = lo;
    const int n = __ldg(pos_idx + pidx);
    if (n < 0 || n >= N) continue;
    const __half* qr = q + static_cast<long long>(b) * ldq;
    const __half* er = e + static_cast<long long>(n) * lde;
    float dot = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float qa = __half2float(qr[d]), eb = __half2float(er[d]);
      dot = fmaf(qa, eb, dot);
      if (split) {
        dot = fmaf(qa, __half2float(e_lo[static_cast<long long>(n) * lde + d]), dot);
        dot = fmaf(__half2float(q_lo[static_cast<long long>(b) * ldq + d]), eb, dot);
      }
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
      if (MODE != MODE_LSE && dS != nullptr)
        dS[(static_cast<long long>(n >> 6) * B + b) * 64 + (n & 63)] = __float2half_rn(g * ds_scale);
    }
  }
  if (MODE == MODE_BCE && lane == 0 && loss_local != 0.0) atomicAdd(loss_sum, loss_local);
}

// C[m, n] = alpha * sum_s part[s, m, n]
__global__ void splitk_reduce_kernel(const float* __restrict__ part, long long split_stride,
                                     int splits, long long M, long long N, float alpha,
                                     const float* __restrict__ s0, const float* __restrict__ s1,
                                     const float* __restrict__ s2, float* __restrict__ C, long long ldc) {
  pdl_wait_and_trigger();
  const long long total = M * N;
  float a = alpha;
  if (s0 != nullptr) a *= __ldg(s0);
  if (s1 != nullptr) a *= __ldg(s1);
  if (s2 != nullptr) a *= __ldg(s2);
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
  pdl_wait_and_trigger();
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
  pdl_wait_and_trigger();
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

// One tensor map: `rank` dimensions (innermost first) of fp32 or fp16 elements; strides[i] = byte pitch of dimension
// i + 1. What-strings name the operand in the error message.
int encode_tmap(CUtensorMap* out, bool f16, int rank, const void* base, const cuuint64_t* dims,
                const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapSwizzle swizzle,
                CUtensorMapL2promotion promo, const char* what) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled not available from the driver");
    return OKGE_ERR_UNSUPPORTED;
  }
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  static_cast<cuuint32_t>(rank), const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[200];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled (%s) failed (CUresult %d) dims=%llu,%llu,%llu stride0=%llu", what,
             (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1],
             (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)strides[0]);
    set_last_error(__FILE__, __LINE__, msg);
    return OKGE_ERR_CUDA;
  }
  return OKGE_OK;
}

struct OperandDesc {
  const void* base;
  const void* lo;      // second plane of a split-precision fp16 operand (row-major only), or nullptr
  int64_t ld;          // elements
  int mode;
};

// Tensor map of one GEMM operand (logical [rows, K]) in any of the four layouts. es = element size, row = elements per
// 128-byte swizzle row.
int make_operand_tmap(CUtensorMap* out, bool f16, OperandDesc* op, int64_t rows, int64_t k, int box_rows, int stage_k) {
  const int es = f16 ? 2 : 4;
  const int row = 128 / es;
  const CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  const CUtensorMapSwizzle mn_swizzle = f16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  switch (op->mode) {
    case OP_ROW_MAJOR: {
      if (!f16) {   // 2-D {K, rows}, box = [box_rows][32 floats]
        cuuint64_t dims[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows)};
        cuuint64_t strides[1] = {static_cast<cuuint64_t>(op->ld) * es};
        cuuint32_t box[2] = {static_cast<cuuint32_t>(row), static_cast<cuuint32_t>(box_rows)};
        return encode_tmap(out, f16, 2, op->base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, promo, "row-major");
      }
      // 3-D {K, rows, planes}: the lo plane of a split-precision operand is plane 1
      int64_t plane_stride = rows * op->ld * es;
      int planes = 1;
      if (op->lo != nullptr) {
        plane_stride = static_cast<const char*>(op->lo) - static_cast<const char*>(op->base);
        planes = 2;
        if (plane_stride <= 0 || plane_stride % 16 != 0) {
          set_last_error(__FILE__, __LINE__, "the lo plane must lie above the hi plane at a multiple of 16 bytes");
          return OKGE_ERR_INVALID;
        }
      }
      cuuint64_t dims[3] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(planes)};
      cuuint64_t strides[2] = {static_cast<cuuint64_t>(op->ld) * es, static_cast<cuuint64_t>(plane_stride)};
      cuuint32_t box[3] = {static_cast<cuuint32_t>(row), static_cast<cuuint32_t>(box_rows), 1};
      return encode_tmap(out, f16, 3, op->base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, promo, "row-major planes");
    }
    case OP_K_PANELS: {   // memory [panels][rows][row]; box = [1][box_rows][row]
      const int64_t panels = ceil_div64(k, row);
      cuuint64_t dims[3] = {static_cast<cuuint64_t>(row), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(panels)};
      cuuint64_t strides[2] = {128, static_cast<cuuint64_t>(rows) * 128};
      cuuint32_t box[3] = {static_cast<cuuint32_t>(row), static_cast<cuuint32_t>(box_rows), 1};
      return encode_tmap(out, f16, 3, op->base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, promo, "K panels");
    }
    case OP_COL_MAJOR: {
      // logical [rows, K] lives in memory as [K][ld] with the rows contiguous (a row-major matrix read as its own
      // transpose). Whole row blocks: one strided 3-D box per stage {row, K, rows / row}; else one 2-D box per block
      // (the TMA unit zero-fills past the edges).
      if (rows % row == 0) {
        cuuint64_t dims[3] = {static_cast<cuuint64_t>(row), static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows / row)};
        cuuint64_t strides[2] = {static_cast<cuuint64_t>(op->ld) * es, 128};
        cuuint32_t box[3] = {static_cast<cuuint32_t>(row), static_cast<cuuint32_t>(stage_k), static_cast<cuuint32_t>(box_rows / row)};
        if (encode_tmap(out, f16, 3, op->base, dims, strides, box, mn_swizzle, promo, "col-major blocks") == OKGE_OK) {
          op->mode = OP_MN_PANELS;
          return OKGE_OK;
        }
      }
      cuuint64_t dims[2] = {static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(k)};
      cuuint64_t strides[1] = {static_cast<cuuint64_t>(op->ld) * es};
      cuuint32_t box[2] = {static_cast<cuuint32_t>(row), static_cast<cuuint32_t>(stage_k)};
      return encode_tmap(out, f16, 2, op->base, dims, strides, box, mn_swizzle, promo, "col-major");
    }
    case OP_MN_PANELS: {   // memory [ceil(rows/row)][K][row] = the K-panel storage of the transposed matrix
      const int64_t panels = ceil_div64(rows, row);
      cuuint64_t dims[3] = {static_cast<cuuint64_t>(row), static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(panels)};
      cuuint64_t strides[2] = {128, static_cast<cuuint64_t>(k) * 128};
      cuuint32_t box[3] = {static_cast<cuuint32_t>(row), static_cast<cuuint32_t>(stage_k), static_cast<cuuint32_t>(box_rows / row)};
      return encode_tmap(out, f16, 3, op->base, dims, strides, box, mn_swizzle, promo, "MN panels");
    }
  }
  set_last_error(__FILE__, __LINE__, "unknown operand layout");
  return OKGE_ERR_INVALID;
}

// Output map of MODE_STORE / MODE_ADAGRAD: [splits][M][N] fp32 with row pitch ldc, box = 32 rows x 32 columns (SW128) or
// 32 rows x 16 columns (64-byte rows, SW64: MODE_ADAGRAD_DEEP).
int make_tmap_out(CUtensorMap* out, float* base, int64_t M, int64_t N, int64_t ldc, int64_t splits, int64_t split_stride,
                  int box_cols = 32) {
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(M), static_cast<cuuint64_t>(splits)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ldc) * sizeof(float),
                           static_cast<cuuint64_t>(splits > 1 ? split_stride : M * ldc) * sizeof(float)};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), 32, 1};
  return encode_tmap(out, false, 3, base, dims, strides, box,
                     box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     "output");
}

// dS as fp16 K-panels [ceil(N/64)][M][64]: box = 32 rows x 32 columns (64-byte rows, SWIZZLE_64B)
int make_tmap_ds(CUtensorMap* out, __half* base, int64_t M, int64_t N) {
  cuuint64_t dims[3] = {64, static_cast<cuuint64_t>(M), static_cast<cuuint64_t>(ceil_div64(N, 64))};
  cuuint64_t strides[2] = {128, static_cast<cuuint64_t>(M) * 128};
  cuuint32_t box[3] = {32, 32, 1};
  return encode_tmap(out, true, 3, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     "dS panels");
}

template <bool F16, int MODE, bool LIMIT = false, int RANK = 0>
int launch_mode(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& td,
                const GemmParams& p, int grid, cudaStream_t stream) {
  using C = Cfg<F16, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    OKGE_CUDA_TRY(cudaFuncSetAttribute(okge_gemm_tc_kernel<F16, MODE, LIMIT, RANK>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  OKGE_LAUNCH((okge_gemm_tc_kernel<F16, MODE, LIMIT, RANK>), grid, C::kThreads, C::kSmemBytes, stream, ta, tb, tc, td, p);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

bool is_mn_major(int mode) { return mode == OP_COL_MAJOR || mode == OP_MN_PANELS; }

// Contractions at least this long take the deep-ring instantiation of the fused update. Measured on B200 (D = 512, fp16
// operands, regular / deep): K = 512 x 10^6 rows 1.85 / 2.14 ms, K = 1,024 x 500 k 1.05 / 1.13, K = 2,048 x 250 k
// 0.78 / 0.72, K = 4,096 x 125 k 0.65 / 0.60.
constexpr int kAdagradDeepMinK = 2048;

int launch_gemm(bool f16, int mode, OperandDesc A, OperandDesc B, int64_t M, int64_t N, int64_t K, GemmParams p,
                cudaStream_t stream) {
  const int es = f16 ? 2 : 4;
  const int align = 16 / es;           // elements per 16 bytes
  OKGE_REQUIRE(A.base != nullptr && B.base != nullptr, "null operand");
  OKGE_REQUIRE(M > 0 && N > 0 && K > 0, "empty GEMM (M, N, K must be > 0)");
  OKGE_REQUIRE(M < INT_MAX && N < INT_MAX && K < INT_MAX, "dimension exceeds int32");
  OKGE_REQUIRE((reinterpret_cast<uintptr_t>(A.base) & 15u) == 0 && (reinterpret_cast<uintptr_t>(B.base) & 15u) == 0,
               "operand base pointers must be 16-byte aligned (TMA)");
  OKGE_REQUIRE(A.mode != OP_ROW_MAJOR || (A.ld % align == 0 && A.ld >= K), "lda must be a multiple of 16 bytes and >= K (TMA)");
  OKGE_REQUIRE(B.mode != OP_ROW_MAJOR || (B.ld % align == 0 && B.ld >= K), "ldb must be a multiple of 16 bytes and >= K (TMA)");
  OKGE_REQUIRE(A.mode != OP_COL_MAJOR || (A.ld % align == 0 && A.ld >= M), "col-major lda must be a multiple of 16 bytes and >= M (TMA)");
  OKGE_REQUIRE(B.mode != OP_COL_MAJOR || (B.ld % align == 0 && B.ld >= N), "col-major ldb must be a multiple of 16 bytes and >= N (TMA)");
  OKGE_REQUIRE((A.lo == nullptr) == (B.lo == nullptr), "split-precision products need the lo plane of both operands");
  OKGE_REQUIRE(A.lo == nullptr || (f16 && A.mode == OP_ROW_MAJOR && B.mode == OP_ROW_MAJOR && p.splits <= 1),
               "split-precision products take row-major fp16 operands and no split-K");
  int st = okge_device_check();
  if (st != OKGE_OK) return st;

  const int row = 128 / es;
  if (mode == MODE_ADAGRAD && K >= kAdagradDeepMinK) mode = MODE_ADAGRAD_DEEP;
  const int stage_k = mode == MODE_ADAGRAD ? row / 2 : row;
  OKGE_REQUIRE(!is_adagrad(mode) || (is_mn_major(A.mode) && is_mn_major(B.mode)),
               "the fused Adagrad contraction takes MN-major operands (OKGE_COL_MAJOR / OKGE_MN_PANELS)");
  CUtensorMap ta, tb;
  st = make_operand_tmap(&ta, f16, &A, M, K, kBM, stage_k);
  if (st != OKGE_OK) return st;
  st = make_operand_tmap(&tb, f16, &B, N, K, kBN, stage_k);
  if (st != OKGE_OK) return st;
  p.a_mode = A.mode;
  p.b_mode = B.mode;
  p.terms = A.lo != nullptr ? 3 : 1;
  const bool a_mn = is_mn_major(A.mode), b_mn = is_mn_major(B.mode);
  const uint32_t lo_mn = desc_lo_mn_major(stage_k);
  const uint32_t hi_mn = f16 ? kDescHiSw128 : kDescHiMnMajorTf32;
  const uint32_t kadv_mn = f16 ? kKadvMnMajorF16 : kKadvMnMajorTf32;
  p.a_desc_lo = a_mn ? lo_mn : kDescLoKMajor;
  p.a_desc_hi = a_mn ? hi_mn : kDescHiSw128;
  p.a_kadv = a_mn ? kadv_mn : kKadvKMajor;
  p.b_desc_lo = b_mn ? lo_mn : kDescLoKMajor;
  p.b_desc_hi = b_mn ? hi_mn : kDescHiSw128;
  p.b_kadv = b_mn ? kadv_mn : kKadvKMajor;
  p.idesc = instr_desc(f16) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.K = static_cast<int>(K);
  p.m_tiles = static_cast<int>(ceil_div64(M, kBM));
  p.n_tiles = static_cast<int>(ceil_div64(N, kBN));
  p.n_fastest = p.n_tiles < p.m_tiles;
  p.k_chunks = static_cast<int>(ceil_div64(K, stage_k));
  if (p.splits < 1) p.splits = 1;
  const int all_chunks = p.k_chunks * p.terms;
  p.k_chunks_per_split = static_cast<int>(ceil_div64(all_chunks, p.splits));
  p.splits = static_cast<int>(ceil_div64(all_chunks, p.k_chunks_per_split));
  const long long total = static_cast<long long>(p.m_tiles) * p.n_tiles * p.splits;
  OKGE_REQUIRE(total < INT_MAX, "too many tiles");
  const int grid = static_cast<int>(total < sm_count() ? total : sm_count());

  CUtensorMap tc, td;
  memset(&tc, 0, sizeof(tc));
  memset(&td, 0, sizeof(td));
  if (is_adagrad(mode)) {
    const int box_cols = mode == MODE_ADAGRAD_DEEP ? Cfg<true, MODE_ADAGRAD_DEEP>::kUpdCols : Cfg<true, MODE_ADAGRAD>::kUpdCols;
    OKGE_REQUIRE(((reinterpret_cast<uintptr_t>(p.param) | reinterpret_cast<uintptr_t>(p.state)) & 15u) == 0 && p.ldc % 4 == 0,
                 "param / state must be 16-byte aligned with a row pitch that is a multiple of 4 (TMA)");
    OKGE_REQUIRE(p.shadow == nullptr || ((reinterpret_cast<uintptr_t>(p.shadow) & 15u) == 0 && p.ld_shadow % 8 == 0 && N % 8 == 0),
                 "the fp16 parameter copy must be 16-byte aligned with N and its row pitch multiples of 8");
    st = make_tmap_out(&tc, p.param, M, N, p.ldc, 1, 0, box_cols);
    if (st != OKGE_OK) return st;
    st = make_tmap_out(&td, p.state, M, N, p.ldc, 1, 0, box_cols);
    if (st != OKGE_OK) return st;
  } else if (mode == MODE_STORE) {
    OKGE_REQUIRE((reinterpret_cast<uintptr_t>(p.C) & 15u) == 0 && p.ldc % 4 == 0,
                 "output must be 16-byte aligned with a leading dimension that is a multiple of 4 (TMA store)");
    st = make_tmap_out(&tc, p.C, M, N, p.ldc, p.splits, p.split_stride);
    if (st != OKGE_OK) return st;
  } else if ((mode == MODE_BCE || mode == MODE_SMGRAD) && p.dS != nullptr) {
    st = make_tmap_ds(&tc, p.dS, M, N);
    if (st != OKGE_OK) return st;
  }
  if (!f16) {
    OKGE_REQUIRE(mode == MODE_STORE, "fp32 operands only feed the plain contraction");
    return launch_mode<false, MODE_STORE>(ta, tb, tc, td, p, grid, stream);
  }
  switch (mode) {
    case MODE_STORE: return launch_mode<true, MODE_STORE>(ta, tb, tc, td, p, grid, stream);
    case MODE_BCE:
      if (p.thresh != nullptr) {
        if (p.rank_slots == 1) return launch_mode<true, MODE_BCE, false, 1>(ta, tb, tc, td, p, grid, stream);
        if (p.rank_slots == 2) return launch_mode<true, MODE_BCE, false, 2>(ta, tb, tc, td, p, grid, stream);
        return launch_mode<true, MODE_BCE, false, 4>(ta, tb, tc, td, p, grid, stream);
      }
      return p.n_limit_dev != nullptr ? launch_mode<true, MODE_BCE, true>(ta, tb, tc, td, p, grid, stream)
                                      : launch_mode<true, MODE_BCE, false>(ta, tb, tc, td, p, grid, stream);
    case MODE_LSE: return launch_mode<true, MODE_LSE>(ta, tb, tc, td, p, grid, stream);
    case MODE_SMGRAD: return launch_mode<true, MODE_SMGRAD>(ta, tb, tc, td, p, grid, stream);
    case MODE_RANK: return launch_mode<true, MODE_RANK>(ta, tb, tc, td, p, grid, stream);
    case MODE_ADAGRAD:
      return p.drop_p > 0.f ? launch_mode<true, MODE_ADAGRAD, true>(ta, tb, tc, td, p, grid, stream)
                            : launch_mode<true, MODE_ADAGRAD>(ta, tb, tc, td, p, grid, stream);
    case MODE_ADAGRAD_DEEP:
      return p.drop_p > 0.f ? launch_mode<true, MODE_ADAGRAD_DEEP, true>(ta, tb, tc, td, p, grid, stream)
                            : launch_mode<true, MODE_ADAGRAD_DEEP>(ta, tb, tc, td, p, grid, stream);
  }
  set_last_error(__FILE__, __LINE__, "unknown epilogue mode");
  return OKGE_ERR_INVALID;
}

// C = alpha * s0 * s1 * s2 * A B^T with optional split-K (fp32 or fp16 operands)
int gemm_store(bool f16, OperandDesc A, OperandDesc B, int64_t M, int64_t N, int64_t K, float alpha, const float* s0,
               const float* s1, const float* s2, float* C, int64_t ldc, int32_t splits, float* split_ws, cudaStream_t s) {
  OKGE_REQUIRE(C != nullptr, "null output");
  OKGE_REQUIRE(ldc >= N, "ldc smaller than N");
  OKGE_REQUIRE(A.mode >= OKGE_ROW_MAJOR && A.mode <= OKGE_MN_PANELS && B.mode >= OKGE_ROW_MAJOR && B.mode <= OKGE_MN_PANELS,
               "unknown operand layout");
  GemmParams p = {};
  p.splits = splits;
  if (splits > 1) {
    OKGE_REQUIRE(split_ws != nullptr, "split-K needs a workspace of splits*M*N floats");
    p.C = split_ws;
    p.ldc = N;
    p.split_stride = M * N;
    p.alpha = 1.0f;
    int st = launch_gemm(f16, MODE_STORE, A, B, M, N, K, p, s);
    if (st != OKGE_OK) return st;
    // recompute the effective split count exactly as launch_gemm did
    const int64_t k_chunks = ceil_div64(K, f16 ? 64 : 32);
    const int64_t per = ceil_div64(k_chunks, splits);
    const int eff_splits = static_cast<int>(ceil_div64(k_chunks, per));
    const long long total = M * N;
    int blocks = static_cast<int>(ceil_div64(total, 256));
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    OKGE_LAUNCH((splitk_reduce_kernel), blocks, 256, 0, s, split_ws, M * N, eff_splits, M, N, alpha, s0, s1, s2, C, ldc);
    OKGE_CUDA_TRY(cudaGetLastError());
    return OKGE_OK;
  }
  p.C = C;
  p.ldc = ldc;
  p.split_stride = 0;
  p.alpha = alpha;
  p.scale_dev[0] = s0;
  p.scale_dev[1] = s1;
  p.scale_dev[2] = s2;
  return launch_gemm(f16, MODE_STORE, A, B, M, N, K, p, s);
}

OperandDesc f16_rows(const okge_half_t* x, const okge_half_t* lo, int64_t ld) {
  return OperandDesc{x, lo, ld, OP_ROW_MAJOR};
}

// one warp per CSR entry; the entry count lives on the device, so the grid is fixed and the warps stride
template <int MODE>
int launch_label_fix(const okge_half_t* q, const okge_half_t* q_lo, int64_t ldq, const okge_half_t* e,
                     const okge_half_t* e_lo, int64_t lde, int64_t B, int64_t N, int64_t D, const int32_t* pos_ptr,
                     const int32_t* pos_idx, const float* q_inv, const float* e_inv, float y_pos, double* loss_sum,
                     float y_delta, okge_half_t* dS, float ds_scale, float* pos_score, const float* row_lse,
                     const float* row_weight, cudaStream_t s) {
  if (pos_idx == nullptr) return OKGE_OK;   // no positives at all
  OKGE_LAUNCH((sparse_label_fix_kernel<MODE>), sm_count() * 4, 256, 0, s, reinterpret_cast<const __half*>(q), reinterpret_cast<const __half*>(q_lo), ldq, reinterpret_cast<const __half*>(e),
      reinterpret_cast<const __half*>(e_lo), lde, static_cast<int>(B), static_cast<int>(N), static_cast<int>(D), pos_ptr,
      pos_idx, q_inv, e_inv, y_pos, loss_sum, y_delta, reinterpret_cast<__half*>(dS), ds_scale, pos_score, row_lse,
      row_weight);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
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
  return gemm_store(false, OperandDesc{A, nullptr, lda, a_layout}, OperandDesc{B, nullptr, ldb, b_layout}, M, N, K, alpha,
                    alpha_dev, nullptr, nullptr, C, ldc, splits, split_ws, static_cast<cudaStream_t>(stream));
}

extern "C" int okge_gemm_f16_nt(const okge_half_t* A, int64_t lda, int32_t a_layout, const okge_half_t* B, int64_t ldb,
                                int32_t b_layout, int64_t M, int64_t N, int64_t K, float alpha, const float* scale0,
                                const float* scale1, const float* scale2, float* C, int64_t ldc, int32_t splits,
                                float* split_ws, okge_stream_t stream) {
  return gemm_store(true, OperandDesc{A, nullptr, lda, a_layout}, OperandDesc{B, nullptr, ldb, b_layout}, M, N, K, alpha,
                    scale0, scale1, scale2, C, ldc, splits, split_ws, static_cast<cudaStream_t>(stream));
}

extern "C" int okge_score_store(const okge_half_t* q, const okge_half_t* q_lo, int64_t ldq, const okge_half_t* e,
                                const okge_half_t* e_lo, int64_t lde, int64_t B, int64_t N, int64_t D,
                                const float* q_inv, const float* e_inv, float* scores, int64_t lds,
                                okge_stream_t stream) {
  return gemm_store(true, f16_rows(q, q_lo, ldq), f16_rows(e, e_lo, lde), B, N, D, 1.0f, q_inv, e_inv, nullptr, scores, lds,
                    1, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int okge_score_bce(const okge_half_t* q, int64_t ldq, const okge_half_t* e, int64_t lde, int64_t B,
                              int64_t N, int64_t D, const float* q_inv, const float* e_inv, const int32_t* pos_ptr,
                              const int32_t* pos_idx, float y_base, float y_pos, const int32_t* n_cols_dev,
                              double* loss_sum, okge_half_t* dS, float ds_scale, okge_stream_t stream) {
  OKGE_REQUIRE(pos_ptr != nullptr && loss_sum != nullptr, "null label pointer / loss output");
  OKGE_REQUIRE((reinterpret_cast<uintptr_t>(dS) & 127u) == 0, "dS panels must be 128-byte aligned");
  int ds_exp = 0;
  OKGE_REQUIRE(dS == nullptr || (ds_scale > 0.f && frexpf(ds_scale, &ds_exp) == 0.5f), "ds_scale must be a power of two");
  if (dS == nullptr) ds_scale = 1.0f;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  OKGE_CUDA_TRY(cudaMemsetAsync(loss_sum, 0, sizeof(double), s));
  GemmParams p = {};
  p.splits = 1;
  p.alpha = 1.0f;
  p.scale_dev[0] = q_inv;
  p.scale_dev[1] = e_inv;
  p.y_base = y_base;
  p.n_limit_dev = n_cols_dev;
  p.loss_sum = loss_sum;
  p.dS = reinterpret_cast<__half*>(dS);
  p.ds_scale = ds_scale;
  p.ds_log2 = log2f(ds_scale);
  int st = launch_gemm(true, MODE_BCE, f16_rows(q, nullptr, ldq), f16_rows(e, nullptr, lde), B, N, D, p, s);
  if (st != OKGE_OK) return st;
  return launch_label_fix<MODE_BCE>(q, nullptr, ldq, e, nullptr, lde, B, N, D, pos_ptr, pos_idx, q_inv, e_inv, y_pos, loss_sum,
                                    y_pos - y_base, dS, ds_scale, nullptr, nullptr, nullptr, s);
}

extern "C" int okge_score_bce_rank(const okge_half_t* q, const okge_half_t* q_lo, int64_t ldq, const okge_half_t* e,
                                   const okge_half_t* e_lo, int64_t lde, int64_t B, int64_t B_extra, int64_t N, int64_t D,
                                   const float* q_inv, const float* e_inv, const int32_t* pos_ptr, const int32_t* pos_idx,
                                   float y_base, float y_pos, const float* thresh4, int32_t* greater4, int32_t* equal4,
                                   int32_t n_slots, double* loss_sum, okge_stream_t stream) {
  OKGE_REQUIRE(n_slots == 1 || n_slots == 2 || n_slots == 4, "n_slots must be 1, 2 or 4");
  OKGE_REQUIRE(pos_ptr != nullptr && loss_sum != nullptr, "null label pointer / loss output");
  OKGE_REQUIRE(B_extra >= 0, "negative number of extra rows");
  OKGE_REQUIRE(thresh4 != nullptr && greater4 != nullptr && equal4 != nullptr, "null ranking pointer");
  OKGE_REQUIRE((reinterpret_cast<uintptr_t>(thresh4) & 15u) == 0, "thresh4 must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  OKGE_CUDA_TRY(cudaMemsetAsync(loss_sum, 0, sizeof(double), s));
  GemmParams p = {};
  p.splits = 1;
  p.alpha = 1.0f;
  p.scale_dev[0] = q_inv;
  p.scale_dev[1] = e_inv;
  p.y_base = y_base;
  p.loss_sum = loss_sum;
  p.thresh = thresh4;
  p.greater = greater4;
  p.equal = equal4;
  p.loss_rows = static_cast<int>(B);
  p.rank_slots = n_slots;
  p.ds_scale = 1.0f;
  int st = launch_gemm(true, MODE_BCE, f16_rows(q, q_lo, ldq), f16_rows(e, e_lo, lde), B + B_extra, N, D, p, s);
  if (st != OKGE_OK) return st;
  return launch_label_fix<MODE_BCE>(q, q_lo, ldq, e, e_lo, lde, B, N, D, pos_ptr, pos_idx, q_inv, e_inv, y_pos, loss_sum,
                                    y_pos - y_base, nullptr, 1.f, nullptr, nullptr, nullptr, s);
}

extern "C" int64_t okge_score_lse_ws_floats(int64_t B, int64_t N) {
  const int64_t P = ceil_div64(N, kBN) * kLseGroups;
  return 2 * P * B + 2 * static_cast<int64_t>(kLseChunks) * B;
}

extern "C" int okge_score_lse(const okge_half_t* q, int64_t ldq, const okge_half_t* e, int64_t lde, int64_t B,
                              int64_t N, int64_t D, const float* q_inv, const float* e_inv, const int32_t* pos_ptr,
                              const int32_t* pos_idx, float* row_lse, float* pos_score, float* part_ws,
                              okge_stream_t stream) {
  OKGE_REQUIRE(pos_ptr != nullptr && row_lse != nullptr && part_ws != nullptr, "null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t P = ceil_div64(N, kBN) * kLseGroups;
  GemmParams p = {};
  p.splits = 1;
  p.alpha = 1.0f;
  p.scale_dev[0] = q_inv;
  p.scale_dev[1] = e_inv;
  p.part_max = part_ws;
  p.part_sum = part_ws + P * B;
  int st = launch_gemm(true, MODE_LSE, f16_rows(q, nullptr, ldq), f16_rows(e, nullptr, lde), B, N, D, p, s);
  if (st != OKGE_OK) return st;
  float* s1max = part_ws + 2 * P * B;
  float* s1sum = s1max + static_cast<int64_t>(kLseChunks) * B;
  const int chunks = static_cast<int>(P < kLseChunks ? P : kLseChunks);
  dim3 g1(static_cast<unsigned>(ceil_div64(B, 128)), static_cast<unsigned>(chunks));
  OKGE_LAUNCH((lse_merge_stage1), g1, 128, 0, s, p.part_max, p.part_sum, static_cast<int>(P),
                                      static_cast<int>(B), s1max, s1sum);
  OKGE_CUDA_TRY(cudaGetLastError());
  OKGE_LAUNCH((lse_merge_stage2), static_cast<unsigned>(ceil_div64(B, 128)), 128, 0, s, s1max, s1sum, chunks, static_cast<int>(B), row_lse);
  OKGE_CUDA_TRY(cudaGetLastError());
  if (pos_score == nullptr) return OKGE_OK;
  return launch_label_fix<MODE_LSE>(q, nullptr, ldq, e, nullptr, lde, B, N, D, pos_ptr, pos_idx, q_inv, e_inv, 1.f, nullptr,
                                    1.f, nullptr, 1.f, pos_score, nullptr, nullptr, s);
}

extern "C" int okge_score_softmax_grad(const okge_half_t* q, int64_t ldq, const okge_half_t* e, int64_t lde,
                                       int64_t B, int64_t N, int64_t D, const float* q_inv, const float* e_inv,
                                       const int32_t* pos_ptr, const int32_t* pos_idx, const float* row_lse,
                                       const float* row_weight, okge_half_t* dS, float ds_scale,
                                       okge_stream_t stream) {
  OKGE_REQUIRE(pos_ptr != nullptr && row_lse != nullptr && row_weight != nullptr && dS != nullptr, "null pointer");
  OKGE_REQUIRE((reinterpret_cast<uintptr_t>(dS) & 127u) == 0, "dS panels must be 128-byte aligned");
  OKGE_REQUIRE(ds_scale > 0.f, "ds_scale must be positive");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GemmParams p = {};
  p.splits = 1;
  p.alpha = 1.0f;
  p.scale_dev[0] = q_inv;
  p.scale_dev[1] = e_inv;
  p.y_base = 0.f;
  p.row_lse = row_lse;
  p.row_weight = row_weight;
  p.dS = reinterpret_cast<__half*>(dS);
  p.ds_scale = ds_scale;
  int st = launch_gemm(true, MODE_SMGRAD, f16_rows(q, nullptr, ldq), f16_rows(e, nullptr, lde), B, N, D, p, s);
  if (st != OKGE_OK) return st;
  return launch_label_fix<MODE_SMGRAD>(q, nullptr, ldq, e, nullptr, lde, B, N, D, pos_ptr, pos_idx, q_inv, e_inv, 1.f,
                                       nullptr, 1.f, dS, ds_scale, nullptr, row_lse, row_weight, s);
}

extern "C" int okge_score_rank(const okge_half_t* q, const okge_half_t* q_lo, int64_t ldq, const okge_half_t* e,
                               const okge_half_t* e_lo, int64_t lde, int64_t Q, int64_t N, int64_t D, const float* q_inv,
                               const float* e_inv, const float* thresh, int32_t* greater, int32_t* equal,
                               okge_stream_t stream) {
  OKGE_REQUIRE(thresh != nullptr && greater != nullptr && equal != nullptr, "null pointer");
  GemmParams p = {};
  p.splits = 1;
  p.alpha = 1.0f;
  p.scale_dev[0] = q_inv;
  p.scale_dev[1] = e_inv;
  p.thresh = thresh;
  p.greater = greater;
  p.equal = equal;
  return launch_gemm(true, MODE_RANK, f16_rows(q, q_lo, ldq), f16_rows(e, e_lo, lde), Q, N, D, p,
                     static_cast<cudaStream_t>(stream));
}

extern "C" int okge_gemm_adagrad_dropout(const okge_half_t* A, int64_t lda, int32_t a_layout, const okge_half_t* B, int64_t ldb,
                                         int32_t b_layout, int64_t M, int64_t N, int64_t K, float alpha, const float* scale0,
                                         const float* scale1, const float* scale2, const int32_t* extra_map,
                                         const float* extra, int64_t ld_extra, float* param, float* state_sum, int64_t ld,
                                         okge_half_t* shadow, int64_t ld_shadow, const float* shadow_inv_scale, float clr,
                                         float eps, float weight_decay, float drop_p, uint64_t drop_seed, uint64_t drop_offset,
                                         const uint64_t* drop_step_dev, okge_stream_t stream);

extern "C" int okge_gemm_adagrad(const okge_half_t* A, int64_t lda, int32_t a_layout, const okge_half_t* B, int64_t ldb,
                                 int32_t b_layout, int64_t M, int64_t N, int64_t K, float alpha, const float* scale0,
                                 const float* scale1, const float* scale2, const int32_t* extra_map, const float* extra,
                                 int64_t ld_extra, float* param, float* state_sum, int64_t ld, okge_half_t* shadow,
                                 int64_t ld_shadow, const float* shadow_inv_scale, float clr, float eps, float weight_decay,
                                 okge_stream_t stream) {
  return okge_gemm_adagrad_dropout(A, lda, a_layout, B, ldb, b_layout, M, N, K, alpha, scale0, scale1, scale2, extra_map, extra,
                                   ld_extra, param, state_sum, ld, shadow, ld_shadow, shadow_inv_scale, clr, eps, weight_decay,
                                   0.f, 0, 0, nullptr, stream);
}

extern "C" int okge_gemm_adagrad_dropout(const okge_half_t* A, int64_t lda, int32_t a_layout, const okge_half_t* B, int64_t ldb,
                                         int32_t b_layout, int64_t M, int64_t N, int64_t K, float alpha, const float* scale0,
                                         const float* scale1, const float* scale2, const int32_t* extra_map,
                                         const float* extra, int64_t ld_extra, float* param, float* state_sum, int64_t ld,
                                         okge_half_t* shadow, int64_t ld_shadow, const float* shadow_inv_scale, float clr,
                                         float eps, float weight_decay, float drop_p, uint64_t drop_seed, uint64_t drop_offset,
                                         const uint64_t* drop_step_dev, okge_stream_t stream) {
  OKGE_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "dropout probability must be in [0, 1)");
  OKGE_REQUIRE(drop_p == 0.f || N % 4 == 0, "the dropout mask is drawn per 4 consecutive elements: N must be a multiple of 4");
  OKGE_REQUIRE(param != nullptr && state_sum != nullptr, "null parameter / accumulator");
  OKGE_REQUIRE(ld >= N, "row pitch smaller than N");
  OKGE_REQUIRE(a_layout >= OKGE_ROW_MAJOR && a_layout <= OKGE_MN_PANELS && b_layout >= OKGE_ROW_MAJOR &&
                   b_layout <= OKGE_MN_PANELS, "unknown operand layout");
  OKGE_REQUIRE(extra_map == nullptr || extra != nullptr, "extra_map without extra rows");
  OKGE_REQUIRE(shadow == nullptr || (ld_shadow >= N && shadow_inv_scale != nullptr), "the fp16 copy needs a row pitch >= N and its device scale");
  GemmParams p = {};
  p.splits = 1;
  p.param = param;
  p.state = state_sum;
  p.ldc = ld;
  p.alpha = alpha;
  p.scale_dev[0] = scale0;
  p.scale_dev[1] = scale1;
  p.scale_dev[2] = scale2;
  p.clr = clr;
  p.eps = eps;
  p.weight_decay = weight_decay;
  p.extra_map = extra_map;
  p.extra = extra;
  p.ld_extra = ld_extra;
  p.shadow = reinterpret_cast<__half*>(shadow);
  p.ld_shadow = ld_shadow;
  p.shadow_inv = shadow_inv_scale;
  p.drop_p = drop_p;
  p.drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  p.drop_seed = drop_seed;
  p.drop_offset = drop_offset;
  p.drop_step_dev = reinterpret_cast<const unsigned long long*>(drop_step_dev);
  return launch_gemm(true, MODE_ADAGRAD, OperandDesc{A, nullptr, lda, a_layout}, OperandDesc{B, nullptr, ldb, b_layout}, M, N,
                     K, p, static_cast<cudaStream_t>(stream));
}
