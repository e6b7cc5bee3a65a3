// One-call composite of a VectorQuantize forward / one ResidualVQ stage.
//
// The Python glue used to issue ~10 ctypes calls + a dozen tiny torch ops per forward; at BASELINE config 2 the
// GPU work is ~0.5 ms and the host needed 1.6 ms to enqueue it.  vqb_vq_forward enqueues the whole chain
// (input staging -> tensor-core search with fused gather tail -> exact re-score -> EMA statistics -> EMA apply
// -> loss) from C++ in one call; the caller only provides outputs and one workspace.
#include "vqb_common.cuh"

using namespace vqb;

namespace {
inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

struct FwdWs {
  size_t x_eff, a_planes, flagged, counters, stats_ws, total;
};

FwdWs carve_fwd(int64_t N, int D, int K, int dtype, int metric, int update) {
  FwdWs w;
  size_t off = 0;
  const size_t esz = dtype == VQB_DTYPE_BF16 ? 2 : 4;
  const bool cosine = metric == VQB_METRIC_COSINE;
  w.x_eff = off;
  if (cosine) off = up256(off + static_cast<size_t>(N) * D * esz);
  w.a_planes = off;
  if (dtype == VQB_DTYPE_F32) off = up256(off + static_cast<size_t>(2) * N * D * 2);
  w.flagged = off;
  off = up256(off + static_cast<size_t>(N) * sizeof(vqb_flag_entry));
  w.counters = off;  // [0..7] flag_count (int32) ; [8..15] loss_sum (double)
  off = up256(off + 16);
  w.stats_ws = off;
  if (update) off = up256(off + vqb_ema_stats_workspace(N, K));
  w.total = off;
  return w;
}
}  // namespace

extern "C" size_t vqb_vq_forward_workspace(int64_t N, int D, int K, int dtype, int metric, int update) {
  if (N <= 0 || D <= 0 || K <= 0) return 0;
  return carve_fwd(N, D, K, dtype, metric, update).total;
}

extern "C" int vqb_vq_forward(const vqb_vq_forward_args* a, void* stream) {
  if (!a || !a->x || !a->embed || !a->planes || !a->bext || !a->cnorm2 || !a->cmax || !a->idx32 || !a->workspace)
    return VQB_E_INVALID;
  if (a->N <= 0 || a->D <= 0 || a->K <= 0) return VQB_E_INVALID;
  if (a->dtype != VQB_DTYPE_F32 && a->dtype != VQB_DTYPE_BF16) return VQB_E_INVALID;
  if (a->update && (!a->stats)) return VQB_E_INVALID;
  if (a->update == 2 && (!a->cluster_size || !a->embed_avg || !a->bias || !a->scratch)) return VQB_E_INVALID;
  const FwdWs w = carve_fwd(a->N, a->D, a->K, a->dtype, a->metric, a->update);
  if (w.total > a->workspace_bytes) return VQB_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(a->workspace) & 255) return VQB_E_ALIGN;
  uint8_t* ws = static_cast<uint8_t*>(a->workspace);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool cosine = a->metric == VQB_METRIC_COSINE;
  const bool l2 = cosine && !a->already_normalised;

  // ---- input staging (vqp:692, :1159 -> :376)
  const void* x_eff = a->x;
  const void* a_planes = a->x;
  int n_a = 1;
  int rc = VQB_OK;
  if (a->dtype == VQB_DTYPE_BF16) {
    if (l2) {
      rc = vqb_input_prepare(a->x, a->dtype, a->N, a->D, 1, ws + w.x_eff, nullptr, 0, stream);
      if (rc) return rc;
      x_eff = ws + w.x_eff;
      a_planes = x_eff;
    }
  } else {
    rc = vqb_input_prepare(a->x, a->dtype, a->N, a->D, l2 ? 1 : 0, l2 ? ws + w.x_eff : nullptr, ws + w.a_planes, 2, stream);
    if (rc) return rc;
    if (l2) x_eff = ws + w.x_eff;
    a_planes = ws + w.a_planes;
    n_a = 2;
  }
  int32_t* flag_count = reinterpret_cast<int32_t*>(ws + w.counters);
  double* loss_sum = reinterpret_cast<double*>(ws + w.counters + 8);
  cudaError_t e = cudaMemsetAsync(ws + w.counters, 0, 16, s);
  if (e != cudaSuccess) return static_cast<int>(e);

  // ---- search with the fused gather / loss / residual tail (vqp:743-747, :766, :1178, :1327; rvq:524-525)
  vqb_fused_outputs f;
  f.x_eff = x_eff; f.embed = a->embed; f.q_out = a->q_out; f.idx64_out = a->idx64_out; f.idx_stride = a->idx_stride;
  f.loss_sum = a->loss_out ? loss_sum : nullptr;
  f.x_raw = (x_eff != a->x) ? a->x : nullptr;
  f.resid_out = a->resid_out; f.qsum = a->qsum; f.dtype = a->dtype;
  const bool fused_stats = a->update && a->stats_mode == 0;
  f.stats_cnt = nullptr; f.stats_sum = nullptr;
  if (fused_stats) {  // statistics ride on the store warps: zero the packed buffer, accumulate with vector REDs
    if (!a->stats_accumulate) {
      e = cudaMemsetAsync(a->stats, 0, sizeof(float) * static_cast<size_t>(vqb_stats_floats(a->K, a->D)), s);
      if (e != cudaSuccess) return static_cast<int>(e);
    }
    f.stats_cnt = a->stats;
    f.stats_sum = a->stats + vqb_stats_offset(a->K);
  }
  const bool want_tail = a->q_out || a->idx64_out || a->loss_out || a->resid_out || a->qsum || fused_stats;
  vqb_flag_entry* flagged = reinterpret_cast<vqb_flag_entry*>(ws + w.flagged);
  if (a->ev_search_begin) cudaEventRecord(static_cast<cudaEvent_t>(a->ev_search_begin), s);
  rc = vqb_assign_ex(a_planes, n_a, a->N, a->D, a->planes, a->bext, a->cmax, a->K, a->margin_rel, 0, a->idx32, flagged,
                     flag_count, nullptr, want_tail ? &f : nullptr, a->metric, a->cnorm2, stream);
  if (rc) return rc;
  if (a->ev_search_end) cudaEventRecord(static_cast<cudaEvent_t>(a->ev_search_end), s);
  rc = vqb_fix_flagged(x_eff, a->dtype, a->N, a->D, a->embed, a->cnorm2, a->K, a->metric, flagged, flag_count, a->idx32,
                       want_tail ? &f : nullptr, stream);
  if (rc) return rc;
  if (a->loss_out) {
    rc = vqb_loss_finalize(loss_sum, a->N * a->D, a->dtype, a->loss_weight, a->loss_out, stream);
    if (rc) return rc;
  }
  // ---- EMA (vqp:586-617, :576-584)
  if (a->update && !fused_stats) {
    rc = vqb_ema_stats(x_eff, a->dtype, a->N, a->D, a->idx32, a->K, a->stats, ws + w.stats_ws,
                       vqb_ema_stats_workspace(a->N, a->K), stream);
    if (rc) return rc;
  }
  if (a->update) {
    if (a->update == 2) {
      rc = vqb_ema_apply(a->cluster_size, a->embed_avg, a->embed, a->stats, a->K, a->D, a->decay, a->eps, a->metric, 1,
                         a->do_normalise, a->planes, a->bext, a->bias, a->cnorm2, a->cmax, a->scratch, stream);
      if (rc) return rc;
    }
  }
  return VQB_OK;
}
