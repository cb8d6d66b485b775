// yy_nn.cuh -- policy/value network inference on the leaf batch (interface used by yy_tree.cu).
// Implementation and memory layouts: yy_nn.cu.
#pragma once
#include "yy_common.cuh"

namespace yy {

struct NNState {
  int rows, cols, A, W, channels, blocks;
  int max_boards;            // capacity of the head-feature scratch (>= n_games)
  const void* weights;       // packed image (device), owned by the caller
  int64_t weight_bytes;
  void* scratch;             // head features etc. (inside the engine workspace)
  int64_t scratch_bytes;
  int device, num_sms;
  bool attrs_set;
  // optional CUDA-event timing of the persistent kernel launches (bench.py's live roofline figure)
  bool profiling;
  cudaEvent_t* ev;           // 2 * ev_cap events
  int ev_cap, ev_used;
  long long tower_launches; double tower_ms; long long tower_boards;
  long long* dbg;            // device buffer for per-layer clock stamps (developer tool), usually nullptr
};

int64_t nn_weight_bytes(int rows, int cols, int channels, int blocks);
size_t nn_workspace_bytes(const yy_engine_config& cfg);
int nn_init(NNState& nn, const yy_engine_config& cfg, void* scratch);
void nn_destroy(NNState& nn);
int nn_load_weights(NNState& nn, const void* weights_dev, int64_t bytes);
int nn_set_profiling(NNState& nn, int enable);
// drains recorded event pairs (synchronises) and returns totals since profiling was enabled
int nn_get_profile(NNState& nn, long long* launches, double* total_ms, long long* boards);
struct EngineDev;
// Persistent kernel (yy_fused.cu): `iterations` x (network forward on boards [0,count) -> heads -> optional tree step)
// in ONE launch.  dev == nullptr: plain forward into policy/value/logits.  dev != nullptr: whole search over the
// engine's leaf batch (slot = game), use_nn = false runs the deterministic-prior evaluator (tree steps only).
int nn_fused_run(NNState& nn, const EngineDev* dev, uint32_t rule_flags, const uint64_t* black, const uint64_t* white,
                 int64_t count, float* policy, float* value, float* logits, int iterations, bool use_nn, cudaStream_t stream);

}  // namespace yy
