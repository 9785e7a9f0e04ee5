// Point-wise part of the LSTM token encoder (LSTMRelationEmbedder, openkge/model.py:912-998: a single-layer
// torch.nn.LSTM over the <= 10 tokens of a mention, output taken at the last real token).
//
// The matrix products of the recurrence run on the tensor-core kernel of gemm_tf32.cu (okge_gemm_tf32_nt):
//   Gx = X W_ih^T for all time steps at once ([L*n, D] x [D, 4D]), Gh_t = h_{t-1} W_hh^T per step,
//   backward: dh_{t-1} = dG_t W_hh, dW_hh = dG^T H_prev, dW_ih = dG^T X, dX = dG W_ih (one contraction each over all steps).
// What is left per time step is element-wise and HBM-bound: the gate non-linearities, the cell / hidden update, the
// selection of the output row at t == last_state[row], and their derivatives. One thread owns 4 consecutive hidden
// units of a row (float4): it reads the four gate slices (i, f, g, o: 4 coalesced 16-byte loads D floats apart), so
// every warp access is a contiguous 512-byte transaction.
#include "okge_common.cuh"

#include <math.h>

namespace okge {

namespace {

constexpr int kLstmThreads = 256;

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

#define OKGE_F4_MAP(out, expr)                    \
  do {                                            \
    { const int k = 0; (out).x = (expr); }        \
    { const int k = 1; (out).y = (expr); }        \
    { const int k = 2; (out).z = (expr); }        \
    { const int k = 3; (out).w = (expr); }        \
  } while (0)

__device__ __forceinline__ float comp(const float4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }

__global__ void __launch_bounds__(kLstmThreads)
lstm_cell_fwd_kernel(const float* __restrict__ gx, int64_t ld_gx, const float* __restrict__ gh, int64_t ld_gh,
                     const float* __restrict__ b_ih, const float* __restrict__ b_hh, const float* __restrict__ c_prev,
                     int64_t n, int D, int t, const int32_t* __restrict__ last_state, float* __restrict__ act,
                     float* __restrict__ c_out, float* __restrict__ h_out, float* __restrict__ out) {
  pdl_wait_and_trigger();
  const int D4 = D >> 2;
  const int64_t total = n * D4;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = e / D4;
    const int d = static_cast<int>(e - row * D4) * 4;
    float4 pre[4];
#pragma unroll
    for (int gate = 0; gate < 4; ++gate) {
      const int col = gate * D + d;
      float4 v = ld4(gx + row * ld_gx + col);
      const float4 bi = ld4(b_ih + col), bh = ld4(b_hh + col);
      v.x += bi.x + bh.x; v.y += bi.y + bh.y; v.z += bi.z + bh.z; v.w += bi.w + bh.w;
      if (gh != nullptr) {
        const float4 r = ld4(gh + row * ld_gh + col);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      pre[gate] = v;
    }
    float4 gi, gf, gg, go, cp = make_float4(0.f, 0.f, 0.f, 0.f), cn, hn;
    if (c_prev != nullptr) cp = ld4(c_prev + row * D + d);
    OKGE_F4_MAP(gi, sigmoidf_acc(comp(pre[0], k)));
    OKGE_F4_MAP(gf, sigmoidf_acc(comp(pre[1], k)));
    OKGE_F4_MAP(gg, tanhf(comp(pre[2], k)));
    OKGE_F4_MAP(go, sigmoidf_acc(comp(pre[3], k)));
    OKGE_F4_MAP(cn, comp(gf, k) * comp(cp, k) + comp(gi, k) * comp(gg, k));
    OKGE_F4_MAP(hn, comp(go, k) * tanhf(comp(cn, k)));
    if (act != nullptr) {
      float* a = act + row * (4 * static_cast<int64_t>(D)) + d;
      st4(a, gi); st4(a + D, gf); st4(a + 2 * D, gg); st4(a + 3 * D, go);
    }
    st4(c_out + row * D + d, cn);
    st4(h_out + row * D + d, hn);
    if (__ldg(last_state + row) == t) st4(out + row * D + d, hn);
  }
}

__global__ void __launch_bounds__(kLstmThreads)
lstm_cell_bwd_kernel(const float* __restrict__ act, const float* __restrict__ c_prev, const float* __restrict__ c,
                     const float* __restrict__ grad_out, const int32_t* __restrict__ last_state, int t,
                     const float* __restrict__ dh_in, float* __restrict__ dc, int64_t n, int D,
                     float* __restrict__ dgates) {
  pdl_wait_and_trigger();
  const int D4 = D >> 2;
  const int64_t total = n * D4;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = e / D4;
    const int d = static_cast<int>(e - row * D4) * 4;
    const float* a = act + row * (4 * static_cast<int64_t>(D)) + d;
    const float4 gi = ld4(a), gf = ld4(a + D), gg = ld4(a + 2 * D), go = ld4(a + 3 * D);
    float4 dh = make_float4(0.f, 0.f, 0.f, 0.f), cp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dh_in != nullptr) dh = ld4(dh_in + row * D + d);
    if (__ldg(last_state + row) == t) {         // the encoder's output is h_t of this row
      const float4 g = ld4(grad_out + row * D + d);
      dh.x += g.x; dh.y += g.y; dh.z += g.z; dh.w += g.w;
    }
    if (c_prev != nullptr) cp = ld4(c_prev + row * D + d);
    const float4 cn = ld4(c + row * D + d);
    const float4 dcn = *reinterpret_cast<const float4*>(dc + row * D + d);
    float4 tc, dct, di, df, dg, d_o, dcp;
    OKGE_F4_MAP(tc, tanhf(comp(cn, k)));
    OKGE_F4_MAP(dct, comp(dcn, k) + comp(dh, k) * comp(go, k) * (1.f - comp(tc, k) * comp(tc, k)));
    OKGE_F4_MAP(di, comp(dct, k) * comp(gg, k) * comp(gi, k) * (1.f - comp(gi, k)));
    OKGE_F4_MAP(df, comp(dct, k) * comp(cp, k) * comp(gf, k) * (1.f - comp(gf, k)));
    OKGE_F4_MAP(dg, comp(dct, k) * comp(gi, k) * (1.f - comp(gg, k) * comp(gg, k)));
    OKGE_F4_MAP(d_o, comp(dh, k) * comp(tc, k) * comp(go, k) * (1.f - comp(go, k)));
    OKGE_F4_MAP(dcp, comp(dct, k) * comp(gf, k));
    float* g = dgates + row * (4 * static_cast<int64_t>(D)) + d;
    st4(g, di); st4(g + D, df); st4(g + 2 * D, dg); st4(g + 3 * D, d_o);
    st4(dc + row * D + d, dcp);
  }
}

int lstm_grid(int64_t n, int64_t D) {
  int64_t blocks = ceil_div64(n * (D / 4), kLstmThreads);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

}  // namespace okge

using namespace okge;

extern "C" int okge_lstm_cell_fwd(const float* gx, int64_t ld_gx, const float* gh, int64_t ld_gh, const float* b_ih,
                                  const float* b_hh, const float* c_prev, int64_t n, int64_t D, int32_t t,
                                  const int32_t* last_state, float* act, float* c, float* h, float* out, void* stream) {
  if (n == 0) return OKGE_OK;
  OKGE_REQUIRE(gx && b_ih && b_hh && last_state && c && h && out, "null pointer");
  OKGE_REQUIRE(D > 0 && D % 4 == 0 && ld_gx % 4 == 0 && (gh == nullptr || ld_gh % 4 == 0), "D and leading dimensions must be multiples of 4");
  OKGE_REQUIRE(aligned16(gx) && aligned16(gh) && aligned16(b_ih) && aligned16(b_hh) && aligned16(c_prev) && aligned16(act) &&
                   aligned16(c) && aligned16(h) && aligned16(out), "operands must be 16-byte aligned");
  OKGE_LAUNCH((lstm_cell_fwd_kernel), lstm_grid(n, D), kLstmThreads, 0, static_cast<cudaStream_t>(stream), gx, ld_gx, gh, ld_gh, b_ih, b_hh, c_prev, n, static_cast<int>(D), t, last_state, act, c, h, out);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}

extern "C" int okge_lstm_cell_bwd(const float* act, const float* c_prev, const float* c, const float* grad_out,
                                  const int32_t* last_state, int32_t t, const float* dh, float* dc, int64_t n, int64_t D,
                                  float* dgates, void* stream) {
  if (n == 0) return OKGE_OK;
  OKGE_REQUIRE(act && c && grad_out && last_state && dc && dgates, "null pointer");
  OKGE_REQUIRE(D > 0 && D % 4 == 0, "D must be a multiple of 4");
  OKGE_REQUIRE(aligned16(act) && aligned16(c_prev) && aligned16(c) && aligned16(grad_out) && aligned16(dh) && aligned16(dc) &&
                   aligned16(dgates), "operands must be 16-byte aligned");
  OKGE_LAUNCH((lstm_cell_bwd_kernel), lstm_grid(n, D), kLstmThreads, 0, static_cast<cudaStream_t>(stream), act, c_prev, c, grad_out, last_state, t, dh, dc, n, static_cast<int>(D), dgates);
  OKGE_CUDA_TRY(cudaGetLastError());
  return OKGE_OK;
}
