// Shared device/host helpers for the sm_100a kernels of the OpenKGE hot path.
// Everything here is hand-written PTX wrappers (mbarrier, TMA, tcgen05/TMEM) plus
// small host utilities; no third-party headers.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/okge_b200.h"

namespace okge {

// ---------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------

#define OKGE_CUDA_TRY(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      okge::set_last_error(__FILE__, __LINE__, cudaGetErrorString(_e));                      \
      return OKGE_ERR_CUDA;                                                                  \
    }                                                                                        \
  } while (0)

#define OKGE_REQUIRE(cond, msg)                                                              \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      okge::set_last_error(__FILE__, __LINE__, msg);                                         \
      return OKGE_ERR_INVALID;                                                               \
    }                                                                                        \
  } while (0)

void set_last_error(const char* file, int line, const char* msg);
int sm_count();
bool pdl_enabled();   // lib.cu: programmatic dependent launch, on unless OKGE_PDL=0

// Every kernel of this library is launched through OKGE_LAUNCH: an ordinary stream launch that additionally allows
// PROGRAMMATIC DEPENDENT LAUNCH. The kernel may be scheduled while its predecessor in the stream is still draining; its
// first statement (pdl_wait_and_trigger) blocks until that predecessor has completed and flushed its memory, so the
// data dependencies are exactly those of a plain launch. What overlaps is the launch latency, the block scheduling and
// -- in the tensor-core kernel -- the prologue (barrier init, TMEM allocation, descriptor prefetch). The steps of the
// small configurations are chains of 20-30 dependent launches of a few microseconds each (inside one CUDA graph, where
// the edges become programmatic ones), which is where this pays.
template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr = {};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // the caller checks cudaGetLastError()
}
#define OKGE_UNPAREN(...) __VA_ARGS__
#define OKGE_LAUNCH(kernel, ...) okge::launch_kernel(OKGE_UNPAREN kernel, __VA_ARGS__)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------

// First statement of every kernel (see OKGE_LAUNCH): wait for the predecessor grid to complete and become visible, then
// let the successor be scheduled (it blocks in its own wait until this grid has completed). Must run in every thread
// before any early return: a block that exits without waiting would let the successor overtake the predecessor.
__device__ __forceinline__ void pdl_wait_and_trigger() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier -------------------------------------------------------------------------------

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// Spin on an mbarrier phase. A watchdog converts a protocol bug into a trap (an error the host
// sees) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 20000000000LL) {  // ~10 s at 2 GHz
      printf("okge: mbarrier watchdog fired (block %d thread %d bar %u parity %u)\n",
             (int)blockIdx.x, (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// ---- TMA ------------------------------------------------------------------------------------

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

// 2-D tiled load global -> shared, completion signalled on an mbarrier (complete_tx::bytes).
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// 3-D tiled load (K-panel operands: dims = {32 floats, rows, panels}).
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar,
                                            int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 3-D tiled store shared -> global (bulk async-group completion); out-of-bounds parts of the box are clipped.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t smem_src, int32_t c0, int32_t c1,
                                             int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// L2 eviction-priority policies for streaming data (written once / read once by this kernel)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap* m, uint32_t smem_src, int32_t c0, int32_t c1, int32_t c2,
                                                  uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
               : "memory");
}
// 16-byte shared-memory load / global store with an L2 eviction policy (epilogues that write a tile straight from the
// staging buffer with coalesced vector stores instead of a bulk tensor store)
__device__ __forceinline__ uint4 lds_v4(uint32_t smem_addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr));
  return v;
}
__device__ __forceinline__ void stg_v4_hint(void* gptr, uint4 v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(gptr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest 0 groups have finished READING shared memory (the buffer may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// at most kPending of the newest store groups may still be reading shared memory
template <int kPending>
__device__ __forceinline__ void tma_store_wait_read_le() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- tcgen05 / TMEM -------------------------------------------------------------------------

__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_result_addr),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 inputs, FP32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// The same with FP16 inputs (kind::f16: 128 x 256 x 16 per instruction at twice the tf32 rate).
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31},"
      " [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t (&v)[16]) { tmem_ld_32x16(taddr, v); }
__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld_32x32(taddr, v); }

__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- TF32 operand rounding -------------------------------------------------------------------
// tcgen05.mma.kind::tf32 TRUNCATES the low 13 mantissa bits of its fp32 operands. Operands produced
// by our own kernels are therefore rounded to nearest TF32 when they are written (then the
// truncation is exact); an operand that comes straight from a parameter table is truncated by the
// hardware, which shrinks it by a factor (1 - d) on average, d = E[2^-11 / m] = 2^-11 / (2 ln 2) for a
// log-uniform mantissa m in [1, 2). Epilogues multiply the accumulator by 1 / (1 - d) for each such
// operand, which centres the error (measured on B200: -7.0e-4 relative bias with both operands raw).
constexpr float kTf32TruncBias = 0.00035221f;                 // 2^-11 / (2 ln 2)
constexpr float kTf32RawOperandScale = 1.0f / (1.0f - kTf32TruncBias);

__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// ---- Adagrad element update (shared by the dense / row-wise kernels and the fused dE epilogue) ----------------------
// torch.optim.Adagrad as the reference runs it (utils/optim.py:139-160, 194-201): g' = g + wd p; G += g'^2;
// p -= clr g' / (sqrt(G) + eps).
__device__ __forceinline__ void adagrad_elem(float& p, float g, float& G, float clr, float eps, float wd) {
  // separate roundings (no FMA contraction) to follow torch's addcmul_/sqrt/add_/addcdiv_ sequence
  g = __fadd_rn(g, __fmul_rn(wd, p));
  G = __fadd_rn(G, __fmul_rn(g, g));
  const float std = __fadd_rn(__fsqrt_rn(G), eps);
  p = __fadd_rn(p, __fmul_rn(-clr, __fdiv_rn(g, std)));
}

// The same update for the fused dE epilogue, where a few epilogue warps do all the arithmetic of a tile: branch-free
// MUFU square root and reciprocal (sqrt.approx / rcp.approx: <= 2^-23 / 1 ulp relative error) instead of the IEEE
// sequences with their slow-path calls. The update term clr g'/(sqrt(G)+eps) is off by at most ~3 of ITS ulps.
__device__ __forceinline__ void adagrad_elem_fast(float& p, float g, float& G, float clr, float eps, float wd) {
  g = __fadd_rn(g, __fmul_rn(wd, p));
  G = __fadd_rn(G, __fmul_rn(g, g));
  float sq, rc;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(G));
  asm("rcp.approx.f32 %0, %1;" : "=f"(rc) : "f"(__fadd_rn(sq, eps)));
  p = __fadd_rn(p, __fmul_rn(-clr, __fmul_rn(g, rc)));
}

// ---- vector memory ops ----------------------------------------------------------------------

__device__ __forceinline__ float4 ldg_nc_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void red_add_f4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// Inverted dropout of 4 consecutive elements whose first element has index 4 * idx4 in the flattened operand: the same
// Philox block, keep rule (u >= p) and scale as dropout_kernel (embed_ops.cu), so a fused epilogue draws the mask the
// separate kernel would draw for (seed, offset).
__device__ __forceinline__ float4 dropout4(float4 v, uint64_t idx4, float p, float scale, uint64_t seed, uint64_t offset) {
  const uint64_t c = idx4 + offset;
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(c), static_cast<uint32_t>(c >> 32), 0u, 0u),
                                make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  float4 o;
  o.x = ((r.x >> 8) * (1.0f / 16777216.0f) >= p) ? v.x * scale : 0.f;
  o.y = ((r.y >> 8) * (1.0f / 16777216.0f) >= p) ? v.y * scale : 0.f;
  o.z = ((r.z >> 8) * (1.0f / 16777216.0f) >= p) ? v.z * scale : 0.f;
  o.w = ((r.w >> 8) * (1.0f / 16777216.0f) >= p) ? v.w * scale : 0.f;
  return o;
}

}  // namespace okge
