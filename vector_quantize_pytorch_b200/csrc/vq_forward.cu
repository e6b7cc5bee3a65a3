// One-call composite of a VectorQuantize forward / one ResidualVQ stage.
//
// The Python glue used to issue ~10 ctypes calls + a dozen tiny torch ops per forward; at BASELINE config 2 the
// GPU work is ~0.5 ms and the host needed 1.6 ms to enqueue it.  vqb_vq_forward enqueues the whole chain
// (input staging -> tensor-core search with fused gather tail -> exact re-score -> EMA statistics -> EMA apply
// -> loss) from C++ in one call; the caller only provides outputs and one workspace.
#include "vqb_common.cuh"
#include <stdlib.h>
#include <string.h>

using namespace vqb;

extern "C" int vqb_debug_active(void);  // vq_assign.cu: diagnostics (profile buffer / debug mode) are armed

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

static int vq_forward_enqueue(const vqb_vq_forward_args* a, void* stream);

// ---------------------------------------------------------------------------------------------
// CUDA-graph cache.  The chain is ~15 small launches around one big kernel; replaying it as a graph removes the
// launch gaps.  A call is identified by every pointer / size / flag of its argument struct; the first occurrence
// of a key is enqueued directly (also warms lazy module loading), the second is captured, later ones replay.
// It pays off when pointer sets repeat: the Python glue keeps its small outputs / scratch in persistent buffers and
// torch's allocator cycles through a handful of large blocks for the per-call outputs.  VQB_GRAPH=0 disables.
// Bypassed while profiling events are requested, or when the stream is already being captured
// (then the launches simply become part of the caller's graph).
// ---------------------------------------------------------------------------------------------
namespace {
struct GraphEntry {
  uint64_t key[40];
  cudaGraphExec_t exec;
  unsigned long long last_use;
};
constexpr int kMaxGraphs = 128;
int g_captures_since_replay = 0;  // safety valve: pointer sets that never repeat make capturing pure overhead
bool g_graph_disabled = false;
GraphEntry g_graphs[kMaxGraphs];
int g_num_graphs = 0;
unsigned long long g_tick = 0;

int graph_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("VQB_GRAPH");
    mode = (e && e[0] == '0') ? 0 : 1;
  }
  return mode;
}

void make_key(const vqb_vq_forward_args* a, void* stream, uint64_t* k) {
  int i = 0;
  auto P = [&](const void* p) { k[i++] = reinterpret_cast<uint64_t>(p); };
  auto I = [&](long long v) { k[i++] = static_cast<uint64_t>(v); };
  auto F = [&](double v) { uint64_t u; memcpy(&u, &v, 8); k[i++] = u; };
  P(a->x); I(a->dtype); I(a->metric); I(a->N); I(a->D); I(a->K); I(a->already_normalised);
  P(a->cluster_size); P(a->embed_avg); P(a->embed); P(a->planes); P(a->bext); P(a->bias); P(a->cnorm2); P(a->cmax);
  P(a->scratch); P(a->q_out); P(a->idx64_out); I(a->idx_stride); P(a->loss_out); F(a->loss_weight); P(a->resid_out);
  P(a->qsum); P(a->idx32); I(a->update); I(a->stats_mode); I(a->stats_accumulate); I(a->do_normalise); F(a->decay);
  F(a->eps); P(a->stats); F(a->margin_rel); P(a->workspace); I(static_cast<long long>(a->workspace_bytes)); P(stream);
  while (i < 40) k[i++] = 0;
}
}  // namespace

extern "C" int vqb_vq_forward(const vqb_vq_forward_args* a, void* stream) {
  if (!a) return VQB_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (!graph_mode() || g_graph_disabled || a->ev_search_begin || a->ev_search_end || vqb_debug_active() ||
      cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone)
    return vq_forward_enqueue(a, stream);
  uint64_t key[40];
  make_key(a, stream, key);
  ++g_tick;
  int slot = -1;
  for (int i = 0; i < g_num_graphs; ++i)
    if (memcmp(g_graphs[i].key, key, sizeof(key)) == 0) { slot = i; break; }
  if (slot >= 0 && g_graphs[slot].exec) {  // replay
    g_graphs[slot].last_use = g_tick;
    g_captures_since_replay = 0;
    const cudaError_t e = cudaGraphLaunch(g_graphs[slot].exec, s);
    return static_cast<int>(e);
  }
  if (slot < 0) {  // first sighting: remember the key, run directly
    if (g_num_graphs < kMaxGraphs) slot = g_num_graphs++;
    else {
      slot = 0;
      for (int i = 1; i < kMaxGraphs; ++i)
        if (g_graphs[i].last_use < g_graphs[slot].last_use) slot = i;
      if (g_graphs[slot].exec) cudaGraphExecDestroy(g_graphs[slot].exec);
    }
    memcpy(g_graphs[slot].key, key, sizeof(key));
    g_graphs[slot].exec = nullptr;
    g_graphs[slot].last_use = g_tick;
    return vq_forward_enqueue(a, stream);
  }
  // second sighting: capture, instantiate, launch
  g_graphs[slot].last_use = g_tick;
  if (++g_captures_since_replay > 48) g_graph_disabled = true;  // instantiations are not paying off: stop
  if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return vq_forward_enqueue(a, stream);
  }
  const int rc = vq_forward_enqueue(a, stream);
  cudaGraph_t graph = nullptr;
  const cudaError_t ee = cudaStreamEndCapture(s, &graph);
  if (rc != VQB_OK || ee != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    if (rc != VQB_OK) return rc;
    return vq_forward_enqueue(a, stream);  // capture failed: nothing ran, enqueue directly
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess || !exec) {
    cudaGetLastError();
    return vq_forward_enqueue(a, stream);
  }
  g_graphs[slot].exec = exec;
  return static_cast<int>(cudaGraphLaunch(exec, s));
}

static int vq_forward_enqueue(const vqb_vq_forward_args* a, void* stream) {
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
  // ResidualVQ stages need x again (residual, running sum): that generic tail is HBM-heavy and throttles the search
  // kernel when run by its four store warps (measured +0.3 ms per stage at config 3), so it runs as the stand-alone
  // gather kernel after the re-score instead.  The VectorQuantize tail (row copy + loss from the scores) stays fused.
  const bool split_tail = (a->resid_out || a->qsum) && !fused_stats;
  const bool want_tail = !split_tail && (a->q_out || a->idx64_out || a->loss_out || fused_stats);
  vqb_flag_entry* flagged = reinterpret_cast<vqb_flag_entry*>(ws + w.flagged);
  if (a->ev_search_begin) cudaEventRecord(static_cast<cudaEvent_t>(a->ev_search_begin), s);
  rc = vqb_assign_ex(a_planes, n_a, a->N, a->D, a->planes, a->bext, a->cmax, a->K, a->margin_rel, 0, a->idx32, flagged,
                     flag_count, nullptr, want_tail ? &f : nullptr, a->metric, a->cnorm2, stream);
  if (rc) return rc;
  if (a->ev_search_end) cudaEventRecord(static_cast<cudaEvent_t>(a->ev_search_end), s);
  rc = vqb_fix_flagged(x_eff, a->dtype, a->N, a->D, a->embed, a->cnorm2, a->K, a->metric, flagged, flag_count, a->idx32,
                       want_tail ? &f : nullptr, stream);
  if (rc) return rc;
  if (split_tail) {
    rc = vqb_gather(x_eff, a->dtype, a->N, a->D, a->embed, a->idx32, a->q_out, a->idx64_out, a->idx_stride,
                    a->loss_out ? loss_sum : nullptr, f.x_raw, a->resid_out, a->qsum, stream);
    if (rc) return rc;
  }
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
