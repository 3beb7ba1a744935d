// Host-side helpers shared by the translation units of libtrafficbots_b200.so.
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/trafficbots_b200.h"
#include "tb_device.cuh"
#include "tb_tc.cuh"

namespace tb {

constexpr int ROW_TILE = 16;  // feature rows per CTA in the row-tile kernels

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern std::atomic<long long> g_launches;  // diagnostics only (tb_launch_count)
inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
inline int launch_status() { return cudaGetLastError() == cudaSuccess ? TB_OK : TB_ERR_LAUNCH; }

int check_dims_host(const TbDims* d);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: every launcher keeps one bit per device ordinal
// (thread-safe; setting the attribute twice is harmless) instead of a process-wide "already set" flag.
inline uint64_t device_bit() {
  int dev = 0;
  cudaGetDevice(&dev);
  return 1ull << (dev & 63);
}
inline bool smem_attr_done(const std::atomic<uint64_t>& m) { return (m.load(std::memory_order_acquire) & device_bit()) != 0; }
inline void smem_attr_mark(std::atomic<uint64_t>& m) { m.fetch_or(device_bit(), std::memory_order_release); }
template <class K>
inline bool set_max_smem(K* kernel, int bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
}


struct StateView {
  float* agent_state;  // [B,A,4]
  uint8_t* valid;      // [2,B,A]
  uint8_t* killed;     // [B,A]
  float* vel;          // [B,A,2]
  float* acc;          // [B,A]
  float* yaw_rate;     // [B,A]
  uint8_t* goal_valid; // [B,A]
  uint8_t* sticky;     // [3,B,A]
  float* hidden;       // [3,B*A,128]
  // private scratch
  float* x0;           // [B,A,128]  map/traffic-light aware agent feature of the current step
  float* kv_int;       // [3,B,A,256] interaction K|V of the current step
  float* goal_in;      // [B,A,128]  add_goal.mlp_in(goal_feature) before mask/ReLU (loop invariant)
  float* latent_in;    // [B,A,128]  add_latent.mlp_in(latent_sample) before mask/ReLU (loop invariant)
  // Scratch of the persistent kernel in "agent-minor" layout [.., 32 column quads, A] float4 (the 32 lanes of a warp = 32
  // consecutive agents read 512 contiguous bytes), one copy per CTA of a cluster:
  float4* hidden_t;    // [n_cluster][3][B][32][A]  working copy of the GRU hidden state (converted from / to `hidden`)
  float4* x0_t;        // [n_cluster][B][32][A]     input of the interaction block
  float4* goal_in_t;   // [B][32][A]   = goal_in
  float4* latent_in_t; // [B][32][A]   = latent_in
  float4* goal_c_t;    // [B][32][A]   = W_out0[:, 128:256] relu(goal_in): step-invariant half of add_goal.mlp_out layer 0
  float4* latent_c_t;  // [B][32][A]   = the same for add_latent
  float4* dest_nodes;  // [B][20][A]   destination polyline nodes (x, y, unit direction), invalid nodes at 1e30
  // 64 < n_agent <= 128: the two CTAs of a scene-mode (64 agents each) exchange their valid masks and the readiness of their
  // interaction key blocks through 16 ints per scene-mode: [0..3] valid mask (lo, hi) of rank 0, 1; [4,5] step counters;
  // [6,7] key-block counters.  Zeroed by tb_rollout_init.
  int32_t* xch;
};

StateView state_view(const TbDims& d, void* base);

// tensor-core polyline encoder: per-CTA scratch of the initial node features (tb_tc_kernels.cu)
constexpr int MAP_TC_MAX_CTA = 148;
size_t map_tc_scratch_bytes(int n_cta);

// generic tensor-core cross-attention layer on compacted key blocks (tb_tc_xlayer.cu)
int launch_xlayer_tc(int block, int layer, const float* src, const uint8_t* src_valid, int n_batch, int n_src,
                     const unsigned char* blocks, const int32_t* n_key, int n_key_max, int kv_share, const float* packed, float* dst,
                     cudaStream_t st);

int launch_kv_project_tc(int block, int layer, const float* tgt, long n_row, const float* packed, float* kv, cudaStream_t st);
int launch_gru_seq_tc(int which, int mode, const float* x, const uint8_t* valid, int n_batch, int n_frame, int n_agent, int t_stride,
                      const float* packed, int gru_base_offset, void* workspace, float* out, uint8_t* out_valid, cudaStream_t st);
size_t dest_lists_bytes(int n_scene, int n_pl);  // admissible-polyline lists per (scene, agent class) + counts
int launch_dest_pairs_tc(const float* U, const float* V, int n_scene, int n_agent, int n_pl, const float* packed, float* logits,
                         const uint8_t* map_valid, const uint8_t* map_type, const uint8_t* agent_type, const uint8_t* agent_valid,
                         int32_t* lists_ws, cudaStream_t st);

// tensor-core polyline encoder: 4 threads per node row, compacted tiles (tb_tc_polyline.cu)
// plan_ws: map_plan_bytes(n_scene * n_pl) bytes = [live_pl | row_start (+1) | plan] int32 (compacted-tile plan, k_map_plan)
size_t map_plan_bytes(long n_pl_total);
int launch_map_polyline_tc2(const TbDims& d, const TbSceneIn& in, const float* packed, float* x0_scratch, int n_cta,
                            float* pl_feature, uint8_t* pl_valid, int32_t* plan_ws, cudaStream_t st);

// tensor-core decode step (tb_tc_rollout.cu)
int launch_pack_kv_tc(const float* kv, const uint8_t* key_valid, int n_set, int n_set_valid, int T, unsigned char* blocks,
                      int32_t* n_key, cudaStream_t st);
bool front_tc_supported(const TbDims& d, const TbRolloutIn& in);
int launch_step_front_tc(const TbDims& d, const TbRolloutIn& in, const float* packed, const StateView& sv, int t, cudaStream_t st);
extern long long* g_debug_trace;  // development aid (tb_debug_set_trace)
// persistent tensor-core rollout (tb_tc_persist.cu): all decode steps t_first..t_last in one launch, n_agent <= 128
bool rollout_tc_supported(const TbDims& d, const TbRolloutIn& in);
int rollout_tc_cluster_size(const TbDims& d);  // CTAs per scene-mode (1, 2 or 4; always 2 for 64 < n_agent <= 128: agent halves)
int launch_rollout_tc(const TbDims& d, const TbRolloutIn& in, const float* packed, const StateView& sv, const TbRolloutOut& out,
                      int t_first, int t_last, cudaStream_t st);
inline bool persist_enabled() {  // read per call: tests toggle it
  const char* e = getenv("TB_DISABLE_PERSIST");
  return !(e && e[0] == '1');
}
inline bool tc_enabled() {
  static const bool on = !(getenv("TB_DISABLE_TC") && getenv("TB_DISABLE_TC")[0] == '1');
  return on;
}

// tensor-core Linear kernels of the training path (tb_train_tc.cu): launches with many rows and 128-wide operands
bool train_tc_enabled();  // TB_TRAIN_NO_TC=1 switches them off (A/B runs)
int launch_train_linear_tc_fwd(const float* x, long M, int K, const float* w, long ldw, int N, const float* bias, int relu,
                               const uint8_t* keep_lin, const float* res, const uint8_t* keep_out, float* y, const uint32_t* drop_seed,
                               uint32_t drop_site, uint32_t drop_thresh, float drop_scale, long drop_offset, cudaStream_t st);
int launch_train_linear_tc_dx(const float* dy, long M, int K, int N, const float* w, long ldw, const float* ym, const uint8_t* rm1,
                              const uint8_t* rm2, float* dx, const uint32_t* drop_seed, uint32_t drop_site, uint32_t drop_thresh,
                              float drop_scale, long drop_offset, cudaStream_t st);
int launch_train_linear_tc_dw(const float* dy, const float* x, long M, int K, int N, const float* ym, const uint8_t* rm1,
                              const uint8_t* rm2, float* dw, long lddw, float* db, const uint32_t* drop_seed, uint32_t drop_site,
                              uint32_t drop_thresh, float drop_scale, long drop_offset, cudaStream_t st);

// the packed parameter buffer = [fp32 blob | pad to 1 KB | tensor-core blocks]
inline size_t tc_blob_offset_bytes() { return ((size_t)TB_PACKED_FLOATS * sizeof(float) + 1023) & ~(size_t)1023; }
inline const unsigned char* tc_blob(const float* packed) { return reinterpret_cast<const unsigned char*>(packed) + tc_blob_offset_bytes(); }

}  // namespace tb
