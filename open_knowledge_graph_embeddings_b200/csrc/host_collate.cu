// Host-side collate of 1-vs-all training batches (see okge_b200.h): plain C++ loops, no GPU work. Lives in the same
// library so that the loader thread of the Python host calls it through the same binding; the call releases the
// interpreter lock (ctypes), which is the point: the per-batch numpy / Python of the collate was what bounded the input
// rate of the small, graph-replayed steps (0.09 ms of interpreter time per 512-row batch against a 0.13 ms GPU step).
#include "okge_common.cuh"

#include <string.h>

namespace okge {
namespace {
inline int64_t pad4(int64_t n) { return (n + 3) / 4 * 4; }
}  // namespace
}  // namespace okge

using namespace okge;

extern "C" int okge_host_collate_plan(const int64_t* rows, int64_t k, int64_t B, const int64_t* lab_ptr,
                                      int64_t n_prefix_rows, int64_t* counts, int64_t* starts) {
  OKGE_REQUIRE(rows && lab_ptr && counts && starts, "null pointer");
  OKGE_REQUIRE(k > 0 && B > 0 && B < (int64_t(1) << 24), "bad batch shape");
  const int64_t o_idx = 2 * pad4(B) + pad4(B + 2);
  starts[0] = 0;
  for (int64_t b = 0; b < k; ++b) {
    int64_t c = 0;
    for (int64_t i = 0; i < B; ++i) {
      const int64_t r = rows[b * B + i];
      OKGE_REQUIRE(r >= 0 && r < n_prefix_rows, "prefix row index out of range");
      c += lab_ptr[r + 1] - lab_ptr[r];
    }
    OKGE_REQUIRE(c < (int64_t(1) << 31) - 8, "more than 2^31 labels in one batch");
    counts[b] = c;
    starts[b + 1] = starts[b] + o_idx + pad4(c);
  }
  return OKGE_OK;
}

extern "C" int okge_host_collate_fill(const int64_t* rows, int64_t k, int64_t B, const uint8_t* row_is_sp,
                                      const int32_t* row_ent, const int32_t* row_rel, const int64_t* lab_ptr,
                                      const int32_t* lab_idx, const int64_t* starts, int32_t* packed, int32_t* n_po) {
  OKGE_REQUIRE(rows && row_is_sp && row_ent && row_rel && lab_ptr && lab_idx && starts && packed && n_po, "null pointer");
  OKGE_REQUIRE(k > 0 && B > 0, "bad batch shape");
  const int64_t o_rel = pad4(B), o_ptr = 2 * pad4(B), o_idx = o_ptr + pad4(B + 2);
  for (int64_t b = 0; b < k; ++b) {
    const int64_t* rb = rows + b * B;
    int32_t* out = packed + starts[b];
    int64_t po = 0;
    for (int64_t i = 0; i < B; ++i) po += row_is_sp[rb[i]] ? 0 : 1;
    // zero the padding of the fixed sections (the label padding is zeroed below)
    for (int64_t i = B; i < o_rel; ++i) out[i] = out[o_rel + i] = 0;
    for (int64_t i = o_ptr + B + 2; i < o_idx; ++i) out[i] = 0;
    // stable partition: the po rows in batch order, then the sp rows in batch order. Two cursors; the label columns of
    // a row go behind those of the rows placed before it, so the po rows are laid out first.
    int64_t nnz = 0;
    int64_t pos = 0;
    for (int pass = 0; pass < 2; ++pass) {
      for (int64_t i = 0; i < B; ++i) {
        const int64_t r = rb[i];
        if ((row_is_sp[r] != 0) != (pass == 1)) continue;
        out[pos] = row_ent[r];
        out[o_rel + pos] = row_rel[r];
        out[o_ptr + pos] = static_cast<int32_t>(nnz);
        const int64_t s = lab_ptr[r], len = lab_ptr[r + 1] - s;
        memcpy(out + o_idx + nnz, lab_idx + s, static_cast<size_t>(len) * sizeof(int32_t));
        nnz += len;
        ++pos;
      }
    }
    out[o_ptr + B] = static_cast<int32_t>(nnz);
    out[o_ptr + B + 1] = static_cast<int32_t>(po);
    for (int64_t i = nnz; i < pad4(nnz); ++i) out[o_idx + i] = 0;
    if (starts[b] + o_idx + pad4(nnz) != starts[b + 1]) {
      set_last_error(__FILE__, __LINE__, "starts do not match the label counts (okge_host_collate_plan of the same rows?)");
      return OKGE_ERR_INVALID;
    }
    n_po[b] = static_cast<int32_t>(po);
  }
  return OKGE_OK;
}
