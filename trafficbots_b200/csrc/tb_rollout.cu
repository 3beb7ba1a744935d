// Closed-loop rollout (WaymoMotion.rollout / WaymoMotion.forward, reference src/pl_modules/waymo_motion.py:108-354).
//
// One decode step is two kernels over (agent row tile, scene-mode):
//   k_step_front : state embedding -> 3 agent->map layers -> 3 agent->traffic-light layers (all row-local: the keys
//                  are the pre-projected map / traffic-light K|V caches) -> interaction K|V of the tile's rows
//   k_step_back  : 3 agent<->agent layers (keys = all agents of the scene-mode, written by k_step_front) -> GRU x3
//                  -> add_goal -> add_latent -> action head -> dynamics -> override/spawn -> rule checks -> kill ->
//                  goal_valid -> reward -> outputs straight into the final [B,A,T,.] layout
// The only cross-row dependency inside a step (agent<->agent attention) is the kernel boundary; nothing returns to
// the host between steps, and the whole loop is CUDA-graph capturable (no allocation, no sync).
#include "tb_host.h"

namespace tb {

static inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

struct StateLayout {
  size_t off[TB_STATE_N_FIELD];
  size_t x0, kv_int, goal_in, latent_in, dest_nodes, hidden_t, x0_t, goal_in_t, latent_in_t, goal_c_t, latent_c_t, xch, total;
};

static StateLayout state_layout(const TbDims& d) {
  const size_t BA = (size_t)d.n_scene * d.n_mode * d.n_agent;
  StateLayout L;
  size_t o = 0;
  auto put = [&](size_t bytes) {
    const size_t at = o;
    o = align_up(o + bytes);
    return at;
  };
  L.off[TB_STATE_AGENT_STATE] = put(BA * 4 * sizeof(float));
  L.off[TB_STATE_VALID] = put(2 * BA);
  L.off[TB_STATE_KILLED] = put(BA);
  L.off[TB_STATE_VEL] = put(BA * 2 * sizeof(float));
  L.off[TB_STATE_ACC] = put(BA * sizeof(float));
  L.off[TB_STATE_YAW_RATE] = put(BA * sizeof(float));
  L.off[TB_STATE_GOAL_VALID] = put(BA);
  L.off[TB_STATE_STICKY] = put(3 * BA);
  L.off[TB_STATE_HIDDEN] = put(3 * BA * D * sizeof(float));
  L.x0 = put(BA * D * sizeof(float));
  // (with 64 < n_agent <= 128 the persistent kernel keeps its exchanged interaction key blocks here: [3][B][2] x 64 KB)
  const size_t a_kv = d.n_agent > 64 && d.n_agent < 128 ? 128 : (size_t)d.n_agent;
  L.kv_int = put(3 * (size_t)d.n_scene * d.n_mode * a_kv * 256 * sizeof(float));
  L.goal_in = put(BA * D * sizeof(float));
  L.latent_in = put(BA * D * sizeof(float));
  L.dest_nodes = put(BA * TB_PL_NODE * 4 * sizeof(float));
  const size_t n_cta = d.n_agent <= 128 ? (size_t)rollout_tc_cluster_size(d) : 0;  // persistent-kernel scratch
  L.hidden_t = put(n_cta * 3 * BA * D * sizeof(float));
  L.x0_t = put(n_cta * BA * D * sizeof(float));
  L.goal_in_t = put(BA * D * sizeof(float));
  L.latent_in_t = put(BA * D * sizeof(float));
  L.goal_c_t = put(BA * D * sizeof(float));
  L.latent_c_t = put(BA * D * sizeof(float));
  L.xch = put((size_t)d.n_scene * d.n_mode * 64);
  L.total = o;
  return L;
}

StateView state_view(const TbDims& d, void* base) {
  const StateLayout L = state_layout(d);
  char* p = reinterpret_cast<char*>(base);
  StateView v;
  v.agent_state = reinterpret_cast<float*>(p + L.off[TB_STATE_AGENT_STATE]);
  v.valid = reinterpret_cast<uint8_t*>(p + L.off[TB_STATE_VALID]);
  v.killed = reinterpret_cast<uint8_t*>(p + L.off[TB_STATE_KILLED]);
  v.vel = reinterpret_cast<float*>(p + L.off[TB_STATE_VEL]);
  v.acc = reinterpret_cast<float*>(p + L.off[TB_STATE_ACC]);
  v.yaw_rate = reinterpret_cast<float*>(p + L.off[TB_STATE_YAW_RATE]);
  v.goal_valid = reinterpret_cast<uint8_t*>(p + L.off[TB_STATE_GOAL_VALID]);
  v.sticky = reinterpret_cast<uint8_t*>(p + L.off[TB_STATE_STICKY]);
  v.hidden = reinterpret_cast<float*>(p + L.off[TB_STATE_HIDDEN]);
  v.x0 = reinterpret_cast<float*>(p + L.x0);
  v.kv_int = reinterpret_cast<float*>(p + L.kv_int);
  v.goal_in = reinterpret_cast<float*>(p + L.goal_in);
  v.latent_in = reinterpret_cast<float*>(p + L.latent_in);
  v.dest_nodes = reinterpret_cast<float4*>(p + L.dest_nodes);
  v.hidden_t = reinterpret_cast<float4*>(p + L.hidden_t);
  v.x0_t = reinterpret_cast<float4*>(p + L.x0_t);
  v.goal_in_t = reinterpret_cast<float4*>(p + L.goal_in_t);
  v.latent_in_t = reinterpret_cast<float4*>(p + L.latent_in_t);
  v.goal_c_t = reinterpret_cast<float4*>(p + L.goal_c_t);
  v.latent_c_t = reinterpret_cast<float4*>(p + L.latent_c_t);
  v.xch = reinterpret_cast<int32_t*>(p + L.xch);
  return v;
}

template <int R>
__device__ __forceinline__ void store_tile(const float* xs, float* __restrict__ dst, int nrow) {
  for (int i = threadIdx.x; i < R * (D / 4); i += NT) {
    const int r = i / (D / 4), c4 = i % (D / 4);
    if (r < nrow) reinterpret_cast<float4*>(dst + (size_t)r * D)[c4] = reinterpret_cast<const float4*>(xs + r * D)[c4];
  }
}
template <int R>
__device__ __forceinline__ void load_tile(float* xs, const float* __restrict__ src, int nrow) {
  for (int i = threadIdx.x; i < R * (D / 4); i += NT) {
    const int r = i / (D / 4), c4 = i % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nrow) v = reinterpret_cast<const float4*>(src + (size_t)r * D)[c4];  // plain load: written by the previous kernel
    reinterpret_cast<float4*>(xs + r * D)[c4] = v;
  }
}

// ------------------------------------------------------------------------------------------------------------
// rollout set-up
// ------------------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(NT) k_rollout_init(TbDims dm, TbRolloutIn in, const float* __restrict__ packed, StateView sv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<R>& sm = *reinterpret_cast<TileSmem<R>*>(smem_raw);
  constexpr int RPT = R / 4;
  const int A = dm.n_agent, K = dm.n_mode, B = dm.n_scene * K;
  const int b = blockIdx.y, a0 = blockIdx.x * R, s = b / K;
  const int nrow = min(R, A - a0);
  const int tid = threadIdx.x;
  const size_t BA = (size_t)B * A;

  // ---- Dynamics.init with GT frame 0 (dynamics.py:29-48), checker / goal flags ------------------------------------
  if (tid < nrow) {
    const int a = a0 + tid;
    const size_t ba = (size_t)b * A + a;
    const size_t g0 = ((size_t)s * dm.n_step_gt + 0) * A + a;
    sv.agent_state[ba * 4 + 0] = in.gt_pos[g0 * 2];
    sv.agent_state[ba * 4 + 1] = in.gt_pos[g0 * 2 + 1];
    sv.agent_state[ba * 4 + 2] = in.gt_yaw[g0];
    sv.agent_state[ba * 4 + 3] = in.gt_spd[g0];
    sv.valid[BA + ba] = in.gt_valid[g0];  // step t reads valid[t & 1]: the first step is t = 1
    sv.valid[ba] = 0;
    sv.killed[ba] = 0;
    sv.vel[ba * 2] = in.gt_vel[g0 * 2];
    sv.vel[ba * 2 + 1] = in.gt_vel[g0 * 2 + 1];
    sv.acc[ba] = in.gt_acc[g0];
    sv.yaw_rate[ba] = in.gt_yaw_rate[g0];
    sv.goal_valid[ba] = in.goal_valid[ba];
    sv.sticky[ba] = 0;
    sv.sticky[BA + ba] = 0;
    sv.sticky[2 * BA + ba] = 0;
    // destination polyline nodes (traffic_rule_checker.py:82-98) as (x, y, unit dir): loop-invariant operand of the
    // dest_reached check; invalid nodes are placed at 1e30 with a zero direction, so they never test true
    long dst = in.dest[ba];
    dst = dst < 0 ? 0 : (dst >= dm.n_pl ? dm.n_pl - 1 : dst);
    const size_t dp = (size_t)s * dm.n_pl + dst;
    for (int n = 0; n < TB_PL_NODE; ++n) {
      const size_t nd = dp * TB_PL_NODE + n;
      float4 o = make_float4(1e30f, 1e30f, 0.f, 0.f);
      if (in.map_valid[nd]) {
        const float ux = in.map_dir[nd * 2], uy = in.map_dir[nd * 2 + 1];
        const float nrm = sqrtf(ux * ux + uy * uy);
        o = make_float4(in.map_pos[nd * 2], in.map_pos[nd * 2 + 1], ux / nrm, uy / nrm);  // zero-length dir -> NaN (compares false)
      }
      sv.dest_nodes[((size_t)b * TB_PL_NODE + n) * A + a] = o;
    }
  }
  // GRU hidden starts at zero (agent_temporal.py:131, traffic_bots.py:159)
  for (int L = 0; L < 3; ++L)
    for (int i = tid; i < nrow * (D / 4); i += NT)
      reinterpret_cast<float4*>(sv.hidden + ((size_t)L * BA + (size_t)b * A + a0) * D)[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  // ---- add_goal.mlp_in(map_feature[dest]) : 3 x [Linear, LayerNorm], ReLU between (mlp.py:36-64) --------------------
  for (int i = tid; i < R * (D / 4); i += NT) {
    const int r = i / (D / 4), c4 = i % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nrow) {
      long dst = in.dest[(size_t)b * A + a0 + r];
      dst = dst < 0 ? 0 : (dst >= dm.n_pl ? dm.n_pl - 1 : dst);
      v = __ldg(reinterpret_cast<const float4*>(in.map_feature + ((size_t)s * dm.n_pl + dst) * D) + c4);
    }
    reinterpret_cast<float4*>(sm.x + r * D)[c4] = v;
  }
  __syncthreads();
  const int gw[3] = {tbw::model_add_goal_mlp_in_fc_layers_0_weight, tbw::model_add_goal_mlp_in_fc_layers_4_weight,
                     tbw::model_add_goal_mlp_in_fc_layers_8_weight};
  const int gb[3] = {tbw::model_add_goal_mlp_in_fc_layers_0_bias, tbw::model_add_goal_mlp_in_fc_layers_4_bias,
                     tbw::model_add_goal_mlp_in_fc_layers_8_bias};
  const int nw[3] = {tbw::model_add_goal_mlp_in_fc_layers_1_weight, tbw::model_add_goal_mlp_in_fc_layers_5_weight,
                     tbw::model_add_goal_mlp_in_fc_layers_9_weight};
  const int nb[3] = {tbw::model_add_goal_mlp_in_fc_layers_1_bias, tbw::model_add_goal_mlp_in_fc_layers_5_bias,
                     tbw::model_add_goal_mlp_in_fc_layers_9_bias};
#pragma unroll 1
  for (int j = 0; j < 3; ++j) {
    gemm128<RPT>(packed + gw[j], D, 0, D / 4, sm.x, D, [&](int r, int c, float v) { sm.t[r * D + c] = v + __ldg(packed + gb[j] + c); });
    __syncthreads();
    layernorm_rows(sm.t, D, sm.x, D, R, packed + nw[j], packed + nb[j]);
    __syncthreads();
    if (j < 2) {
      for (int i = tid; i < R * D; i += NT) sm.x[i] = fmaxf(sm.x[i], 0.f);
      __syncthreads();
    }
  }
  store_tile<R>(sm.x, sv.goal_in + ((size_t)b * A + a0) * D, nrow);
  for (int i = tid; i < R * (D / 4); i += NT) {  // agent-minor copy for the persistent kernel
    const int r = i % R, c4 = i / R;
    if (r < nrow) sv.goal_in_t[((size_t)b * (D / 4) + c4) * A + a0 + r] = reinterpret_cast<const float4*>(sm.x + r * D)[c4];
  }
  __syncthreads();
  // step-invariant half of add_goal.mlp_out layer 0: W[:, 128:256] relu(goal_in)  (cat[x, z] @ W^T = x-half + z-half,
  // add_latent_goal.py:64-77; the persistent kernel adds it in the epilogue where goal_valid(t) holds)
  for (int i = tid; i < R * D; i += NT) sm.t[i] = fmaxf(sm.x[i], 0.f);
  __syncthreads();
  gemm128<RPT>(packed + tbw::model_add_goal_mlp_out_fc_layers_0_weight + (size_t)(D / 4) * D * 4, D, 0, D / 4, sm.t, D,
               [&](int r, int c, float v) { sm.q[r * D + c] = v; });
  __syncthreads();
  for (int i = tid; i < R * (D / 4); i += NT) {
    const int r = i % R, c4 = i / R;
    if (r < nrow) sv.goal_c_t[((size_t)b * (D / 4) + c4) * A + a0 + r] = reinterpret_cast<const float4*>(sm.q + r * D)[c4];
  }
  __syncthreads();

  // ---- add_latent.mlp_in(latent_sample): Linear(16,128)-ReLU-Linear(128,128) -----------------------------------------
  for (int i = tid; i < R * TB_LATENT; i += NT) {
    const int r = i / TB_LATENT, c = i % TB_LATENT;
    sm.t[r * D + c] = r < nrow ? in.latent_sample[((size_t)b * A + a0 + r) * TB_LATENT + c] : 0.f;
  }
  __syncthreads();
  gemm128<RPT>(packed + tbw::model_add_latent_mlp_in_fc_layers_0_weight, D, 0, TB_LATENT / 4, sm.t, D, [&](int r, int c, float v) {
    sm.q[r * D + c] = fmaxf(v + __ldg(packed + tbw::model_add_latent_mlp_in_fc_layers_0_bias + c), 0.f);
  });
  __syncthreads();
  gemm128<RPT>(packed + tbw::model_add_latent_mlp_in_fc_layers_3_weight, D, 0, D / 4, sm.q, D, [&](int r, int c, float v) {
    sm.x[r * D + c] = v + __ldg(packed + tbw::model_add_latent_mlp_in_fc_layers_3_bias + c);
  });
  __syncthreads();
  store_tile<R>(sm.x, sv.latent_in + ((size_t)b * A + a0) * D, nrow);
  for (int i = tid; i < R * (D / 4); i += NT) {
    const int r = i % R, c4 = i / R;
    if (r < nrow) sv.latent_in_t[((size_t)b * (D / 4) + c4) * A + a0 + r] = reinterpret_cast<const float4*>(sm.x + r * D)[c4];
  }
  __syncthreads();
  for (int i = tid; i < R * D; i += NT) sm.t[i] = fmaxf(sm.x[i], 0.f);
  __syncthreads();
  gemm128<RPT>(packed + tbw::model_add_latent_mlp_out_fc_layers_0_weight + (size_t)(D / 4) * D * 4, D, 0, D / 4, sm.t, D,
               [&](int r, int c, float v) { sm.q[r * D + c] = v; });
  __syncthreads();
  for (int i = tid; i < R * (D / 4); i += NT) {
    const int r = i % R, c4 = i / R;
    if (r < nrow) sv.latent_c_t[((size_t)b * (D / 4) + c4) * A + a0 + r] = reinterpret_cast<const float4*>(sm.q + r * D)[c4];
  }
}

// ------------------------------------------------------------------------------------------------------------
// step, front half
// ------------------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(NT) k_step_front(TbDims dm, TbRolloutIn in, const float* __restrict__ packed, StateView sv, int t) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<R>& sm = *reinterpret_cast<TileSmem<R>*>(smem_raw);
  __shared__ float3 pose[R];
  const int A = dm.n_agent, K = dm.n_mode, S = dm.n_scene, B = S * K, P = dm.n_pl, TL = dm.n_tl, Th = dm.n_step_hist;
  const int b = blockIdx.y, a0 = blockIdx.x * R, s = b / K;
  const int nrow = min(R, A - a0);
  const int tid = threadIdx.x;
  const size_t BA = (size_t)B * A;
  const uint8_t* valid_cur = sv.valid + (size_t)(t & 1) * BA;

  // ---- get_agent_attr_and_pe (sc_input.py:142-165): [vel2, spd, yaw_rate, acc, size3, type3]; vel/acc/yaw_rate are the
  //      Dynamics.vel/.acc/.yaw_rate attributes that only overrides refresh --------------------------------
  if (tid < R) {
    const int r = tid;
    float* at = sm.t + r * D;
    for (int i = 0; i < 12; ++i) at[i] = 0.f;
    bool valid = false;
    float3 p = make_float3(0.f, 0.f, 0.f);
    if (r < nrow) {
      const int a = a0 + r;
      const size_t ba = (size_t)b * A + a, sa = (size_t)s * A + a;
      valid = valid_cur[ba] != 0;
      const float4 st = *reinterpret_cast<const float4*>(sv.agent_state + ba * 4);
      p = make_float3(st.x, st.y, st.z);
      at[0] = sv.vel[ba * 2];
      at[1] = sv.vel[ba * 2 + 1];
      at[2] = st.w;
      at[3] = sv.yaw_rate[ba];
      at[4] = sv.acc[ba];
      for (int i = 0; i < 3; ++i) {
        at[5 + i] = in.agent_size[sa * 3 + i];
        at[8 + i] = in.agent_type[sa * 3 + i] ? 1.f : 0.f;
      }
    }
    pose[r] = p;
    sm.row_valid[r] = valid;
  }
  __syncthreads();
  // agent_encoder (waymo_motion.py:155)
  {
    for (int i = tid; i < R * 48; i += NT) {
      const int r = i / 48, j = i % 48;
      pose_pe_elem(j, pose[r].x, pose[r].y, pose[r].z, packed + tbw::pre_processing_input_pose_pe_agent_pe_xy_freqs,
                   packed + tbw::pre_processing_input_pose_pe_agent_pe_yaw_freqs, sm.x + r * D + 32);
    }
    gemm_small(packed + tbw::model_agent_encoder_mlp_fc_layers_0_weight, 32, 3, sm.t, D, R, [&](int r, int n, float v) {
      sm.q[r * D + n] = fmaxf(v + __ldg(packed + tbw::model_agent_encoder_mlp_fc_layers_0_bias + n), 0.f);
    });
    __syncthreads();
    gemm_small(packed + tbw::model_agent_encoder_mlp_fc_layers_3_weight, 32, 8, sm.q, D, R, [&](int r, int n, float v) {
      sm.x[r * D + n] = v + __ldg(packed + tbw::model_agent_encoder_mlp_fc_layers_3_bias + n);
    });
    __syncthreads();
    for (int i = tid; i < R * D; i += NT)
      if (!sm.row_valid[i / D]) sm.x[i] = 0.f;
    __syncthreads();
  }
  // ---- transformer_as2pl (traffic_bots.py:205-211) ---------------------------------------------------------------------
#pragma unroll 1
  for (int L = 0; L < 3; ++L)
    xlayer_tile<R>(sm, packed + tbw::model_transformer_as2pl_layers_0_norm1_weight + L * tfl::STRIDE,
                   in.kv_map + ((size_t)L * S + s) * P * 256, in.map_feature_valid + (size_t)s * P, P, -1);
  // ---- transformer_as2tl (:213-219); traffic-light frame = min(t-1, Th-1) (waymo_motion.py:287) ---------------------------
  const int tl_t = min(t - 1, Th - 1);
#pragma unroll 1
  for (int L = 0; L < 3; ++L)
    xlayer_tile<R>(sm, packed + tbw::model_transformer_as2tl_layers_0_norm1_weight + L * tfl::STRIDE,
                   in.kv_tl + (((size_t)L * S + s) * Th + tl_t) * TL * 256, in.tl_valid + ((size_t)s * Th + tl_t) * TL, TL, -1);
  // ---- hand-over to the back half: x0 and the interaction K|V of these rows (agent_interaction.py:52: tgt = block input)
  store_tile<R>(sm.x, sv.x0 + ((size_t)b * A + a0) * D, nrow);
#pragma unroll 1
  for (int L = 0; L < 3; ++L)
    kv_project_tile<R>(sm, packed + tbw::model_agent_interaction_transformer_layers_0_norm1_weight + L * tfl::STRIDE,
                       sv.kv_int + (((size_t)L * B + b) * A + a0) * 256, nrow);
}

// ------------------------------------------------------------------------------------------------------------
// step, back half
// ------------------------------------------------------------------------------------------------------------
template <int R>
struct BackSmem {
  TileSmem<R> tile;
  float h[R * D];
  float mean[R * 2];
  int n_valid;
};

__device__ __forceinline__ float smooth_l1(float d) {
  const float a = fabsf(d);
  return a < 1.0f ? 0.5f * d * d : a - 0.5f;
}

template <int R>
__global__ void __launch_bounds__(NT) k_step_back(TbDims dm, TbRolloutIn in, const float* __restrict__ packed, StateView sv,
                                                  TbRolloutOut out, int t) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BackSmem<R>& bs = *reinterpret_cast<BackSmem<R>*>(smem_raw);
  TileSmem<R>& sm = bs.tile;
  constexpr int RPT = R / 4;
  const int A = dm.n_agent, K = dm.n_mode, S = dm.n_scene, B = S * K, T = dm.n_step, Tg = dm.n_step_gt;
  const int b = blockIdx.y, a0 = blockIdx.x * R, s = b / K;
  const int nrow = min(R, A - a0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t BA = (size_t)B * A;
  const uint8_t* valid_cur = sv.valid + (size_t)(t & 1) * BA;
  uint8_t* valid_next = sv.valid + (size_t)((t + 1) & 1) * BA;

  load_tile<R>(sm.x, sv.x0 + ((size_t)b * A + a0) * D, nrow);
  if (tid < R) sm.row_valid[tid] = (tid < nrow) ? valid_cur[(size_t)b * A + a0 + tid] : (uint8_t)0;
  if (warp == 0) {
    int cnt = 0;
    for (int a = lane; a < A; a += 32) cnt += valid_cur[(size_t)b * A + a] ? 1 : 0;
    cnt = (int)warp_sum((float)cnt);
    if (lane == 0) bs.n_valid = cnt;
  }
  __syncthreads();

  // ---- agent_interaction (agent_interaction.py:51-93): bypassed when exactly one agent of the scene is valid ---------------
  if (bs.n_valid != 1) {
#pragma unroll 1
    for (int L = 0; L < 3; ++L)
      xlayer_tile<R>(sm, packed + tbw::model_agent_interaction_transformer_layers_0_norm1_weight + L * tfl::STRIDE,
                     sv.kv_int + ((size_t)L * B + b) * A * 256, valid_cur + (size_t)b * A, A, a0);
  }

  // ---- agent_temporal: 3-layer GRU, one time step (agent_temporal.py:147-153) ------------------------------------------------
#pragma unroll 1
  for (int L = 0; L < 3; ++L) {
    float* hid = sv.hidden + ((size_t)L * BA + (size_t)b * A + a0) * D;
    load_tile<R>(bs.h, hid, nrow);
    __syncthreads();
    const float* gw = packed + gru::BASE + L * gru::STRIDE;
    const int cg = tid & 63, rg = tid >> 6;
    float rr[RPT][2], zz[RPT][2], ai[RPT][2], ah[RPT][2];
#pragma unroll
    for (int i = 0; i < RPT; ++i) rr[i][0] = rr[i][1] = zz[i][0] = zz[i][1] = ai[i][0] = ai[i][1] = ah[i][0] = ah[i][1] = 0.f;
    gemm_acc<RPT>(gw + gru::W_IH, 3 * D, 0, D / 4, sm.x, D, rr);
    gemm_acc<RPT>(gw + gru::W_HH, 3 * D, 0, D / 4, bs.h, D, rr);
    gemm_acc<RPT>(gw + gru::W_IH, 3 * D, D, D / 4, sm.x, D, zz);
    gemm_acc<RPT>(gw + gru::W_HH, 3 * D, D, D / 4, bs.h, D, zz);
    gemm_acc<RPT>(gw + gru::W_IH, 3 * D, 2 * D, D / 4, sm.x, D, ai);
    gemm_acc<RPT>(gw + gru::W_HH, 3 * D, 2 * D, D / 4, bs.h, D, ah);
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int r = rg * RPT + i;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = cg + 64 * j;
        const float rgate = sigmoidf_(rr[i][j] + __ldg(gw + gru::B_IH + c) + __ldg(gw + gru::B_HH + c));
        const float zgate = sigmoidf_(zz[i][j] + __ldg(gw + gru::B_IH + D + c) + __ldg(gw + gru::B_HH + D + c));
        const float n = tanhf(ai[i][j] + __ldg(gw + gru::B_IH + 2 * D + c) + rgate * (ah[i][j] + __ldg(gw + gru::B_HH + 2 * D + c)));
        sm.q[r * D + c] = (1.0f - zgate) * n + zgate * bs.h[r * D + c];
      }
    }
    __syncthreads();
    for (int i = tid; i < R * D; i += NT) {
      const int r = i / D;
      const float v = sm.q[i];
      sm.x[i] = v;  // the next GRU layer sees the unmasked output
      if (r < nrow) hid[i] = sm.row_valid[r] ? v : 0.f;  // h[:, ~valid] = 0
    }
    __syncthreads();
  }
  for (int i = tid; i < R * D; i += NT)
    if (!sm.row_valid[i / D]) sm.x[i] = 0.f;
  // (barrier below, after the goal operand is staged)

  // ---- add_goal (add_latent_goal.py:57-77, mode cat, res_add): z = relu(mask(mlp_in(goal), goal_valid)) ----------------------
  __shared__ uint8_t goal_valid_s[R];
  if (tid < R) goal_valid_s[tid] = (tid < nrow) ? sv.goal_valid[(size_t)b * A + a0 + tid] : (uint8_t)0;
  __syncthreads();
  for (int i = tid; i < R * D; i += NT) {
    const int r = i / D;
    sm.t[i] = (r < nrow && goal_valid_s[r]) ? fmaxf(sv.goal_in[((size_t)b * A + a0) * D + i], 0.f) : 0.f;
  }
  __syncthreads();
  gemm128_cat<RPT>(packed + tbw::model_add_goal_mlp_out_fc_layers_0_weight, D, 0, sm.x, D, D / 4, sm.t, D, D / 4,
                   [&](int r, int c, float v) { sm.q[r * D + c] = fmaxf(v + __ldg(packed + tbw::model_add_goal_mlp_out_fc_layers_0_bias + c), 0.f); });
  __syncthreads();
  gemm128<RPT>(packed + tbw::model_add_goal_mlp_out_fc_layers_3_weight, D, 0, D / 4, sm.q, D, [&](int r, int c, float v) {
    const float hg = fmaxf(v + __ldg(packed + tbw::model_add_goal_mlp_out_fc_layers_3_bias + c), 0.f);
    const float y = (goal_valid_s[r] ? hg : 0.f) + sm.x[r * D + c];
    sm.x[r * D + c] = sm.row_valid[r] ? y : 0.f;
  });
  __syncthreads();
  // ---- add_latent: z_valid = agent_valid ---------------------------------------------------------------------------------------
  for (int i = tid; i < R * D; i += NT) {
    const int r = i / D;
    sm.t[i] = (r < nrow && sm.row_valid[r]) ? fmaxf(sv.latent_in[((size_t)b * A + a0) * D + i], 0.f) : 0.f;
  }
  __syncthreads();
  gemm128_cat<RPT>(packed + tbw::model_add_latent_mlp_out_fc_layers_0_weight, D, 0, sm.x, D, D / 4, sm.t, D, D / 4,
                   [&](int r, int c, float v) { sm.q[r * D + c] = fmaxf(v + __ldg(packed + tbw::model_add_latent_mlp_out_fc_layers_0_bias + c), 0.f); });
  __syncthreads();
  gemm128<RPT>(packed + tbw::model_add_latent_mlp_out_fc_layers_3_weight, D, 0, D / 4, sm.q, D, [&](int r, int c, float v) {
    const float hz = fmaxf(v + __ldg(packed + tbw::model_add_latent_mlp_out_fc_layers_3_bias + c), 0.f);
    sm.x[r * D + c] = sm.row_valid[r] ? hz + sm.x[r * D + c] : 0.f;
  });
  __syncthreads();
  if (out.trace_policy_feature) {
    for (int i = tid; i < nrow * D; i += NT) {
      const int r = i / D, c = i % D;
      out.trace_policy_feature[(((size_t)b * A + a0 + r) * T + (t - 1)) * D + c] = sm.x[i];
    }
  }

  // ---- action head (action_head.py:70-87): per-type MLP 128->128->2, masked by type & valid, summed ------------------------
  __shared__ uint8_t type_s[R * 3];
  if (tid < R * 3) {
    const int r = tid / 3;
    type_s[tid] = (r < nrow) ? in.agent_type[((size_t)s * A + a0 + r) * 3 + tid % 3] : (uint8_t)0;
  }
  if (tid < R * 2) bs.mean[tid] = 0.f;
  const int hw1[3] = {tbw::action_head_mlp_mean_0_fc_layers_0_weight, tbw::action_head_mlp_mean_1_fc_layers_0_weight,
                      tbw::action_head_mlp_mean_2_fc_layers_0_weight};
  const int hb1[3] = {tbw::action_head_mlp_mean_0_fc_layers_0_bias, tbw::action_head_mlp_mean_1_fc_layers_0_bias,
                      tbw::action_head_mlp_mean_2_fc_layers_0_bias};
  const int hw2[3] = {tbw::action_head_mlp_mean_0_fc_layers_2_weight, tbw::action_head_mlp_mean_1_fc_layers_2_weight,
                      tbw::action_head_mlp_mean_2_fc_layers_2_weight};
  const int hb2[3] = {tbw::action_head_mlp_mean_0_fc_layers_2_bias, tbw::action_head_mlp_mean_1_fc_layers_2_bias,
                      tbw::action_head_mlp_mean_2_fc_layers_2_bias};
#pragma unroll 1
  for (int c3 = 0; c3 < 3; ++c3) {
    gemm128<RPT>(packed + hw1[c3], D, 0, D / 4, sm.x, D,
                 [&](int r, int c, float v) { sm.q[r * D + c] = fmaxf(v + __ldg(packed + hb1[c3] + c), 0.f); });
    __syncthreads();
    for (int item = warp; item < R * 2; item += NWARP) {
      const int r = item >> 1, d = item & 1;
      const float4 w = __ldg(reinterpret_cast<const float4*>(packed + hw2[c3]) + lane * 2 + d);  // Wt4[32][2][4]
      const float4 x = *reinterpret_cast<const float4*>(sm.q + r * D + lane * 4);
      const float v = warp_sum(x.x * w.x + x.y * w.y + x.z * w.z + x.w * w.w) + __ldg(packed + hb2[c3] + d);
      if (lane == 0 && type_s[r * 3 + c3] && sm.row_valid[r]) bs.mean[item] += v;
    }
    __syncthreads();
  }

  // ---- per-agent tail: dynamics, override, rule checks, kill, goal_valid, reward, outputs --------------------------------------
  if (tid < nrow) {
    const int r = tid, a = a0 + r;
    const size_t ba = (size_t)b * A + a, sa = (size_t)s * A + a;
    const bool valid = sm.row_valid[r] != 0;
    const bool ty0 = type_s[r * 3 + 0], ty1 = type_s[r * 3 + 1], ty2 = type_s[r * 3 + 2];
    const float mean0 = bs.mean[r * 2], mean1 = bs.mean[r * 2 + 1];
    // DiagGaussian log-prob of the deterministic sample (= mean), dynamics.py:77-80
    float logp = 0.f;
    if (valid) {
      for (int d = 0; d < 2; ++d) {
        float ls = 0.f;
        if (ty0) ls += __ldg(packed + tbw::action_head_log_std_0 + d);
        if (ty1) ls += __ldg(packed + tbw::action_head_log_std_1 + d);
        if (ty2) ls += __ldg(packed + tbw::action_head_log_std_2 + d);
        logp += -logf(expf(ls)) - 0.91893853320467267f;
      }
    }
    // MultiPathPP.process_action / update (dynamics.py:187-228); type order of the parameter tuples: veh, ped, cyc
    const float max_acc = (ty0 ? 5.0f : 0.f) + (ty1 ? 7.0f : 0.f) + (ty2 ? 6.0f : 0.f);
    const float max_yr = (ty0 ? 1.5f : 0.f) + (ty1 ? 7.0f : 0.f) + (ty2 ? 3.0f : 0.f);
    const float a_acc = valid ? tanhf(mean0) * max_acc : 0.f;
    const float a_yr = valid ? tanhf(mean1) * max_yr : 0.f;
    const float4 st = *reinterpret_cast<const float4*>(sv.agent_state + ba * 4);
    const float v_t = st.w + 0.05f * a_acc, th_t = st.z + 0.05f * a_yr;
    const bool has_type = ty0 || ty1 || ty2;
    float4 pred = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid && has_type) {
      pred.x = st.x + 0.1f * (v_t * cosf(th_t));
      pred.y = st.y + 0.1f * (v_t * sinf(th_t));
      pred.z = st.z + 0.1f * a_yr;
      pred.w = st.w + 0.1f * a_acc;
    }
    const size_t o = ba * T + (t - 1);
    *reinterpret_cast<float4*>(out.preds + o * 4) = pred;
    out.valid[o] = valid;
    out.action_log_probs[o] = logp;
    out.latent_log_probs[o] = in.latent_logp[ba];
    if (out.trace_action_mean) {
      out.trace_action_mean[o * 2] = mean0;
      out.trace_action_mean[o * 2 + 1] = mean1;
    }
    // Dynamics.override_states (dynamics.py:121-149)
    const bool has_gt = t < Tg;
    const size_t g = ((size_t)s * Tg + (has_gt ? t : 0)) * A + a;
    const bool ovr = has_gt && in.tf_mask[g] != 0;
    const bool gt_valid = has_gt && in.gt_valid[g] != 0;
    bool killed = sv.killed[ba] != 0;
    const bool m = ovr && !killed;
    bool nvalid = valid || m;
    float4 ns = pred;
    float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_gt) gs = make_float4(in.gt_pos[g * 2], in.gt_pos[g * 2 + 1], in.gt_yaw[g], in.gt_spd[g]);
    if (m) {
      ns = gs;
      sv.vel[ba * 2] = in.gt_vel[g * 2];
      sv.vel[ba * 2 + 1] = in.gt_vel[g * 2 + 1];
      sv.acc[ba] = in.gt_acc[g];
      sv.yaw_rate[ba] = in.gt_yaw_rate[g];
    }
    out.override_masks[o] = ovr;
    // TrafficRuleChecker.check, always-on subset (traffic_rule_checker.py:101-119,338-410,423-424,474-496)
    const float* mb = in.map_boundary + (size_t)s * 4;
    const bool out_t = nvalid && (ns.x > mb[1] || ns.x < mb[0] || ns.y > mb[3] || ns.y < mb[2]);
    bool outside = sv.sticky[ba] != 0, goal_r = sv.sticky[BA + ba] != 0, dest_r = sv.sticky[2 * BA + ba] != 0;
    outside |= out_t;
    bool goal_t = false;
    if (in.goal_gt) {
      const float* gg = in.goal_gt + sa * 4;
      const float dx = ns.x - gg[0], dy = ns.y - gg[1];
      const bool pos_ok = sqrtf(dx * dx + dy * dy) < in.agent_size[sa * 3] * 8.0f;
      // cast_rad (transform_utils.py:10-12): (a + pi) % (2 pi) - pi with Python's sign-of-divisor modulo
      const float PI_F = 3.14159265358979323846f, TWO_PI_F = 6.28318530717958647692f;
      float w = fmodf(ns.z - gg[2] + PI_F, TWO_PI_F);
      if (w < 0.f) w += TWO_PI_F;
      const bool rot_ok = fabsf(w - PI_F) < 0.26179938779914943654f;
      goal_t = pos_ok && rot_ok && nvalid && !goal_r;
    }
    goal_r |= goal_t;
    long dst = in.dest[ba];
    dst = dst < 0 ? 0 : (dst >= dm.n_pl ? dm.n_pl - 1 : dst);
    const size_t dp = (size_t)s * dm.n_pl + dst;
    const uint8_t* dtype = in.map_type + dp * TB_PL_TYPE;
    const bool lane_t = dtype[0] || dtype[1] || dtype[2] || dtype[3], edge_t = dtype[4] != 0;
    const float thresh = 50.0f * (1.0f - (edge_t ? 1.0f : 0.f) * 0.8f);
    bool pos_reached = false, rot_reached = false;
    const float hx = cosf(ns.z), hy = sinf(ns.z);
    for (int n = 0; n < TB_PL_NODE; ++n) {
      const size_t nd = dp * TB_PL_NODE + n;
      if (!in.map_valid[nd]) continue;
      const float dx = ns.x - in.map_pos[nd * 2], dy = ns.y - in.map_pos[nd * 2 + 1];
      pos_reached |= sqrtf(dx * dx + dy * dy) < thresh;
      const float ux = in.map_dir[nd * 2], uy = in.map_dir[nd * 2 + 1];
      const float nrm = sqrtf(ux * ux + uy * uy);
      rot_reached |= (hx * (ux / nrm) + hy * (uy / nrm)) > 0.86602540378443864676f;  // NaN (zero-length dir) compares false
    }
    const bool dest_t = !dest_r && nvalid && ((lane_t && pos_reached && rot_reached) || (edge_t && pos_reached));
    dest_r |= dest_t;
    const size_t vs = BA * T;
    out.violations[0 * vs + o] = outside;
    out.violations[1 * vs + o] = out_t;
    out.violations[2 * vs + o] = goal_r;
    out.violations[3 * vs + o] = goal_t;
    out.violations[4 * vs + o] = dest_r;
    out.violations[5 * vs + o] = dest_t;
    // Dynamics.kill (dynamics.py:151-167): outside_map_this_step & ~gt_valid
    const bool kill = out_t && !gt_valid;
    killed |= kill;
    nvalid = nvalid && !kill;
    // disable_goal_reached (goal_manager.py:155-161)
    const bool gv = goal_valid_s[r] && nvalid && !dest_r;
    // DifferentiableReward.get, imitation part (rewards.py:117-131)
    float reward = 0.f;
    bool rv = valid;
    if (has_gt) {
      rv = valid && gt_valid;
      if (rv) {
        const float e_pos = smooth_l1(gs.x - pred.x) + smooth_l1(gs.y - pred.y);
        const float e_rot = 0.5f * (1.0f - cosf(gs.z - pred.z));
        const float e_spd = smooth_l1(gs.w - pred.w);
        reward = 0.0f - (0.1f * e_pos + 10.0f * e_rot + 0.1f * e_spd);
      }
    }
    out.diffbar_rewards[o] = reward;
    out.diffbar_rewards_valid[o] = rv;
    // state for the next step
    *reinterpret_cast<float4*>(sv.agent_state + ba * 4) = ns;
    valid_next[ba] = nvalid;
    sv.killed[ba] = killed;
    sv.goal_valid[ba] = gv;
    sv.sticky[ba] = outside;
    sv.sticky[BA + ba] = goal_r;
    sv.sticky[2 * BA + ba] = dest_r;
  }
}

std::atomic<long long> g_launches{0};

}  // namespace tb

// ==============================================================================================================
// host side
// ==============================================================================================================
using namespace tb;

extern "C" int64_t tb_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" size_t tb_rollout_state_bytes(const TbDims* d) {
  if (check_dims_host(d) != TB_OK) return 0;
  return state_layout(*d).total;
}
extern "C" size_t tb_rollout_state_offset(const TbDims* d, int32_t field) {
  if (check_dims_host(d) != TB_OK || field < 0 || field >= TB_STATE_N_FIELD) return (size_t)-1;
  return state_layout(*d).off[field];
}

static int check_rollout_in(const TbRolloutIn* in) {
  if (!in) return TB_ERR_NULL;
  const void* req[] = {in->map_feature, in->map_feature_valid, in->kv_map,     in->kv_tl,    in->tl_valid,   in->gt_valid,
                       in->gt_pos,      in->gt_yaw,            in->gt_spd,     in->gt_vel,   in->gt_acc,     in->gt_yaw_rate,
                       in->tf_mask,     in->agent_type,        in->agent_size, in->map_boundary, in->map_valid, in->map_type,
                       in->map_pos,     in->map_dir,           in->latent_sample, in->latent_logp, in->dest, in->goal_valid};
  for (const void* p : req)
    if (!p) return TB_ERR_NULL;
  if (!aligned16(in->map_feature) || !aligned16(in->kv_map) || !aligned16(in->kv_tl)) return TB_ERR_ALIGN;
  return TB_OK;
}

template <int R>
static int set_rollout_attrs() {
  static std::atomic<uint64_t> done{0};
  if (smem_attr_done(done)) return TB_OK;
  if (!set_max_smem(k_rollout_init<R>, (int)sizeof(TileSmem<R>))) return TB_ERR_LAUNCH;
  if (!set_max_smem(k_step_front<R>, (int)sizeof(TileSmem<R>))) return TB_ERR_LAUNCH;
  if (!set_max_smem(k_step_back<R>, (int)sizeof(BackSmem<R>))) return TB_ERR_LAUNCH;
  smem_attr_mark(done);
  return TB_OK;
}

extern "C" int32_t tb_rollout_init(const TbDims* dims, const TbRolloutIn* in, const float* packed, void* state, void* stream) {
  int rc = check_dims_host(dims);
  if (rc != TB_OK) return rc;
  rc = check_rollout_in(in);
  if (rc != TB_OK) return rc;
  if (!packed || !state) return TB_ERR_NULL;
  if (!aligned16(packed) || (reinterpret_cast<uintptr_t>(state) & 255u)) return TB_ERR_ALIGN;
  constexpr int R = ROW_TILE;
  if (set_rollout_attrs<R>() != TB_OK) return TB_ERR_LAUNCH;
  const TbDims d = *dims;
  dim3 grid((d.n_agent + R - 1) / R, d.n_scene * d.n_mode);
  const StateView sv0 = state_view(d, state);
  if (cudaMemsetAsync(sv0.xch, 0, (size_t)d.n_scene * d.n_mode * 64, (cudaStream_t)stream) != cudaSuccess) return TB_ERR_LAUNCH;
  k_rollout_init<R><<<grid, NT, sizeof(TileSmem<R>), (cudaStream_t)stream>>>(d, *in, packed, sv0);
  count_launch();
  return launch_status();
}

// which: 1 = front half only, 2 = back half only, 3 = both
static int rollout_steps_impl(const TbDims* dims, const TbRolloutIn* in, const float* packed, void* state,
                              const TbRolloutOut* out, int32_t t_first, int32_t t_last, void* stream, int which) {
  int rc = check_dims_host(dims);
  if (rc != TB_OK) return rc;
  rc = check_rollout_in(in);
  if (rc != TB_OK) return rc;
  if (!packed || !state) return TB_ERR_NULL;
  if (which & 2) {
    if (!out) return TB_ERR_NULL;
    if (!out->preds || !out->valid || !out->override_masks || !out->diffbar_rewards || !out->diffbar_rewards_valid ||
        !out->action_log_probs || !out->latent_log_probs || !out->violations)
      return TB_ERR_NULL;
    if (!aligned16(out->preds)) return TB_ERR_ALIGN;
  }
  if (!aligned16(packed) || (reinterpret_cast<uintptr_t>(state) & 255u)) return TB_ERR_ALIGN;
  if (t_first < 1 || t_last > dims->n_step || t_first > t_last) return TB_ERR_BAD_SHAPE;
  constexpr int R = ROW_TILE;
  if (set_rollout_attrs<R>() != TB_OK) return TB_ERR_LAUNCH;
  const TbDims d = *dims;
  const StateView sv = state_view(d, state);
  dim3 grid((d.n_agent + R - 1) / R, d.n_scene * d.n_mode);
  cudaStream_t st = (cudaStream_t)stream;
  if (which == 3 && tc_enabled() && persist_enabled() && rollout_tc_supported(d, *in))
    return launch_rollout_tc(d, *in, packed, sv, *out, t_first, t_last, st);
  for (int t = t_first; t <= t_last; ++t) {
    if (which & 1) {
      if (tc_enabled() && front_tc_supported(d, *in)) {
        const int rc2 = launch_step_front_tc(d, *in, packed, sv, t, st);
        if (rc2 != TB_OK) return rc2;
      } else {
        k_step_front<R><<<grid, NT, sizeof(TileSmem<R>), st>>>(d, *in, packed, sv, t);
        count_launch();
      }
    }
    if (which & 2) {
      k_step_back<R><<<grid, NT, sizeof(BackSmem<R>), st>>>(d, *in, packed, sv, *out, t);
      count_launch();
    }
  }
  return launch_status();
}

extern "C" int32_t tb_rollout_steps(const TbDims* dims, const TbRolloutIn* in, const float* packed, void* state,
                                    const TbRolloutOut* out, int32_t t_first, int32_t t_last, void* stream) {
  return rollout_steps_impl(dims, in, packed, state, out, t_first, t_last, stream, 3);
}
extern "C" int32_t tb_step_front(const TbDims* dims, const TbRolloutIn* in, const float* packed, void* state, int32_t t,
                                 void* stream) {
  return rollout_steps_impl(dims, in, packed, state, nullptr, t, t, stream, 1);
}
extern "C" int32_t tb_step_back(const TbDims* dims, const TbRolloutIn* in, const float* packed, void* state,
                                const TbRolloutOut* out, int32_t t, void* stream) {
  return rollout_steps_impl(dims, in, packed, state, out, t, t, stream, 2);
}

extern "C" int32_t tb_rollout(const TbDims* dims, const TbRolloutIn* in, const float* packed, void* state,
                              const TbRolloutOut* out, void* stream) {
  int rc = tb_rollout_init(dims, in, packed, state, stream);
  if (rc != TB_OK) return rc;
  return tb_rollout_steps(dims, in, packed, state, out, 1, dims->n_step, stream);
}
