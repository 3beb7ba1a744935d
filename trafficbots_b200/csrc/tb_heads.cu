// Pre-rollout heads, fp32 row-tile kernels on the building blocks of tb_device.cuh:
//   tb_gru_sequence  -- `MultiAgentGRULoop.forward` 3-D branch (reference src/models/modules/agent_temporal.py:133-146) fused
//                       with `TemporalAggregate` (:31-36,43-44): prior / posterior latent encoder and destination predictor
//   tb_mlp_head      -- `MLP` 128 -> 128 -> N (mlp.py:20-85), e.g. the latent mean (latent_encoder.py:195-199)
//   tb_dest_logits   -- `DestPredictor.forward`, mode mlp (goal_manager.py:228-246,294-333): pairwise (agent, polyline) MLP,
//                       type masks and the Categorical normalisation (distributions.py:161-165)
// The cross-attention layers of the latent encoder reuse tb_xlayer / tb_kv_project (tb_encode.cu).
#include "tb_host.h"

namespace tb {

constexpr int HR = 16;  // rows per CTA

struct GruSmem {
  float x[HR * D];     // input / output of the current GRU layer
  float h[HR * D];     // hidden state of the current layer (loaded from hs)
  float q[HR * D];     // new hidden state
  float hs[3][HR * D]; // hidden states of the 3 layers, carried over the time steps
  float agg[HR * D];   // temporal aggregate
  float xin[HR * D];   // the step's input (residual of mode 1)
  uint8_t valid[HR];
  uint8_t any[HR];
};

// x [B, T_all, A, 128], valid [B, T_all, A]; frames t = 0, t_stride, 2 t_stride, ... (n_t of them).
// mode 0: out = max over the valid frames of the GRU output (0 if none)              -> TemporalAggregate max_valid
// mode 1: out = (GRU output + input) at the last valid frame (0 if none)             -> DestPredictor :298-300, last_valid
template <int R>
__global__ void __launch_bounds__(NT) k_gru_seq(const float* __restrict__ x, const uint8_t* __restrict__ valid, int T_all, int A,
                                                int t_stride, int n_t, const float* __restrict__ gw0, int mode,
                                                float* __restrict__ out, uint8_t* __restrict__ out_valid) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GruSmem& sm = *reinterpret_cast<GruSmem*>(smem_raw);
  constexpr int RPT = R / 4;
  const int b = blockIdx.y, a0 = blockIdx.x * R, tid = threadIdx.x;
  const int nrow = min(R, A - a0);
  for (int i = tid; i < 3 * R * D; i += NT) (&sm.hs[0][0])[i] = 0.f;  // h starts at zero (agent_temporal.py:131)
  for (int i = tid; i < R * D; i += NT) sm.agg[i] = mode == 0 ? -1e3f : 0.f;
  if (tid < R) sm.any[tid] = 0;
  __syncthreads();
#pragma unroll 1
  for (int it = 0; it < n_t; ++it) {
    const int t = it * t_stride;
    for (int i = tid; i < R * (D / 4); i += NT) {
      const int r = i / (D / 4), c4 = i % (D / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nrow) v = __ldg(reinterpret_cast<const float4*>(x + (((size_t)b * T_all + t) * A + a0 + r) * D) + c4);
      reinterpret_cast<float4*>(sm.x + r * D)[c4] = v;
      reinterpret_cast<float4*>(sm.xin + r * D)[c4] = v;
    }
    if (tid < R) sm.valid[tid] = tid < nrow ? valid[((size_t)b * T_all + t) * A + a0 + tid] : (uint8_t)0;
    __syncthreads();
#pragma unroll 1
    for (int L = 0; L < 3; ++L) {
      const float* gw = gw0 + L * gru::STRIDE;
      const int cg = tid & 63, rg = tid >> 6;
      float rr[RPT][2], zz[RPT][2], ai[RPT][2], ah[RPT][2];
#pragma unroll
      for (int i = 0; i < RPT; ++i) rr[i][0] = rr[i][1] = zz[i][0] = zz[i][1] = ai[i][0] = ai[i][1] = ah[i][0] = ah[i][1] = 0.f;
      const float* hl = sm.hs[L];
      gemm_acc<RPT>(gw + gru::W_IH, 3 * D, 0, D / 4, sm.x, D, rr);
      gemm_acc<RPT>(gw + gru::W_HH, 3 * D, 0, D / 4, hl, D, rr);
      gemm_acc<RPT>(gw + gru::W_IH, 3 * D, D, D / 4, sm.x, D, zz);
      gemm_acc<RPT>(gw + gru::W_HH, 3 * D, D, D / 4, hl, D, zz);
      gemm_acc<RPT>(gw + gru::W_IH, 3 * D, 2 * D, D / 4, sm.x, D, ai);
      gemm_acc<RPT>(gw + gru::W_HH, 3 * D, 2 * D, D / 4, hl, D, ah);
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = rg * RPT + i;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = cg + 64 * j;
          const float rgate = sigmoidf_(rr[i][j] + __ldg(gw + gru::B_IH + c) + __ldg(gw + gru::B_HH + c));
          const float zgate = sigmoidf_(zz[i][j] + __ldg(gw + gru::B_IH + D + c) + __ldg(gw + gru::B_HH + D + c));
          const float n = tanhf(ai[i][j] + __ldg(gw + gru::B_IH + 2 * D + c) + rgate * (ah[i][j] + __ldg(gw + gru::B_HH + 2 * D + c)));
          sm.q[r * D + c] = (1.0f - zgate) * n + zgate * hl[r * D + c];
        }
      }
      __syncthreads();
      for (int i = tid; i < R * D; i += NT) {
        const float v = sm.q[i];
        sm.x[i] = v;                                      // the next layer sees the unmasked output
        sm.hs[L][i] = sm.valid[i / D] ? v : 0.f;          // h[:, ~valid] = 0 after the step
      }
      __syncthreads();
    }
    // output of the step (zero where invalid) -> temporal aggregate
    for (int i = tid; i < R * D; i += NT) {
      const int r = i / D;
      const bool v = sm.valid[r] != 0;
      if (mode == 0) {
        sm.agg[i] = fmaxf(sm.agg[i], v ? sm.x[i] : -1e3f);
      } else if (v) {
        sm.agg[i] = sm.x[i] + sm.xin[i];
      }
    }
    if (tid < R && sm.valid[tid]) sm.any[tid] = 1;
    __syncthreads();
  }
  for (int i = tid; i < R * (D / 4); i += NT) {
    const int r = i / (D / 4), c4 = i % (D / 4);
    if (r < nrow) {
      float4 v = reinterpret_cast<const float4*>(sm.agg + r * D)[c4];
      if (!sm.any[r]) v = make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(out + ((size_t)b * A + a0 + r) * D)[c4] = v;
    }
  }
  if (tid < nrow) out_valid[(size_t)b * A + a0 + tid] = sm.any[tid];
}

// y[row, 0:n_out] = valid[row] ? W2 relu(W1 x[row] + b1) + b2 : 0        (W1 [128,128], W2 [n_out,128], n_out <= 128)
template <int R>
__global__ void __launch_bounds__(NT) k_mlp_head(const float* __restrict__ x, const uint8_t* __restrict__ valid, long n_row,
                                                 const float* __restrict__ w1, const float* __restrict__ b1,
                                                 const float* __restrict__ w2, const float* __restrict__ b2, int n_out,
                                                 float* __restrict__ y) {
  __shared__ __align__(16) float xs[R * D];
  __shared__ __align__(16) float hs[R * D];
  constexpr int RPT = R / 4;
  const long row0 = (long)blockIdx.x * R;
  const int nrow = (int)min((long)R, n_row - row0), tid = threadIdx.x;
  for (int i = tid; i < R * (D / 4); i += NT) {
    const int r = i / (D / 4), c4 = i % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nrow) v = __ldg(reinterpret_cast<const float4*>(x + (row0 + r) * D) + c4);
    reinterpret_cast<float4*>(xs + r * D)[c4] = v;
  }
  __syncthreads();
  gemm128<RPT>(w1, D, 0, D / 4, xs, D, [&](int r, int c, float v) { hs[r * D + c] = fmaxf(v + __ldg(b1 + c), 0.f); });
  __syncthreads();
  gemm_small(w2, n_out, D / 4, hs, D, nrow, [&](int r, int n, float v) {
    y[(row0 + r) * n_out + n] = valid[row0 + r] ? v + __ldg(b2 + n) : 0.f;
  });
}

// y[row] = W[:, k0 : k0 + 128] x[row] (+ bias)          W packed Wt4[K/4][128][4]
template <int R>
__global__ void __launch_bounds__(NT) k_row_linear(const float* __restrict__ x, long n_row, const float* __restrict__ w, int k0,
                                                   const float* __restrict__ bias, float* __restrict__ y) {
  __shared__ __align__(16) float xs[R * D];
  constexpr int RPT = R / 4;
  const long row0 = (long)blockIdx.x * R;
  const int nrow = (int)min((long)R, n_row - row0), tid = threadIdx.x;
  for (int i = tid; i < R * (D / 4); i += NT) {
    const int r = i / (D / 4), c4 = i % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nrow) v = __ldg(reinterpret_cast<const float4*>(x + (row0 + r) * D) + c4);
    reinterpret_cast<float4*>(xs + r * D)[c4] = v;
  }
  __syncthreads();
  gemm128<RPT>(w + (size_t)(k0 / 4) * D * 4, D, 0, D / 4, xs, D, [&](int r, int c, float v) {
    if (r < nrow) y[(row0 + r) * D + c] = v + (bias ? __ldg(bias + c) : 0.f);
  });
}

// Pairwise destination MLP.  CTA = (64 polylines p0.., agent a, scene s):
//   h1 = relu(LN1(U[s,p] + V[s,a]))   (U already holds W0[:, :128] map_feature + b0, V = W0[:, 128:] tgt)
//   h2 = relu(LN2(W3 h1 + b3));  logit = w6 . h2 + b6
constexpr int PR = 64;
struct PairSmem {
  float x[PR * D];
  float t[PR * D];
  float v[D];
};
__global__ void __launch_bounds__(NT) k_dest_pairs(const float* __restrict__ U, const float* __restrict__ V, int P, int A,
                                                   const float* __restrict__ packed, float* __restrict__ logits) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PairSmem& sm = *reinterpret_cast<PairSmem*>(smem_raw);
  constexpr int RPT = PR / 4;
  const int p0 = blockIdx.x * PR, a = blockIdx.y, s = blockIdx.z, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nrow = min(PR, P - p0);
  if (tid < D) sm.v[tid] = V[((size_t)s * A + a) * D + tid];
  __syncthreads();
  for (int i = tid; i < PR * (D / 4); i += NT) {
    const int r = i / (D / 4), c4 = i % (D / 4);
    float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nrow) u = __ldg(reinterpret_cast<const float4*>(U + ((size_t)s * P + p0 + r) * D) + c4);
    const float4 w = reinterpret_cast<const float4*>(sm.v)[c4];
    reinterpret_cast<float4*>(sm.t + r * D)[c4] = make_float4(u.x + w.x, u.y + w.y, u.z + w.z, u.w + w.w);
  }
  __syncthreads();
  layernorm_rows(sm.t, D, sm.x, D, PR, packed + tbw::model_goal_manager_goal_predictor_mlp_fc_layers_1_weight,
                 packed + tbw::model_goal_manager_goal_predictor_mlp_fc_layers_1_bias);
  __syncthreads();
  for (int i = tid; i < PR * D; i += NT) sm.x[i] = fmaxf(sm.x[i], 0.f);
  __syncthreads();
  gemm128<RPT>(packed + tbw::model_goal_manager_goal_predictor_mlp_fc_layers_3_weight, D, 0, D / 4, sm.x, D, [&](int r, int c, float v) {
    sm.t[r * D + c] = v + __ldg(packed + tbw::model_goal_manager_goal_predictor_mlp_fc_layers_3_bias + c);
  });
  __syncthreads();
  layernorm_rows(sm.t, D, sm.x, D, PR, packed + tbw::model_goal_manager_goal_predictor_mlp_fc_layers_4_weight,
                 packed + tbw::model_goal_manager_goal_predictor_mlp_fc_layers_4_bias);
  __syncthreads();
  // last Linear(128, 1): W6 packed Wt4[32][1][4] = the 128 weights in order
  const float4 w6 = __ldg(reinterpret_cast<const float4*>(packed + tbw::model_goal_manager_goal_predictor_mlp_fc_layers_6_weight) + lane);
  const float b6 = __ldg(packed + tbw::model_goal_manager_goal_predictor_mlp_fc_layers_6_bias);
  for (int r = warp; r < nrow; r += NWARP) {
    const float4 h = *reinterpret_cast<const float4*>(sm.x + r * D + lane * 4);
    const float d = warp_sum(fmaxf(h.x, 0.f) * w6.x + fmaxf(h.y, 0.f) * w6.y + fmaxf(h.z, 0.f) * w6.z + fmaxf(h.w, 0.f) * w6.w);
    if (lane == 0) logits[((size_t)s * A + a) * P + p0 + r] = d + b6;
  }
}

// masks + Categorical(logits=) normalisation, one warp per (scene, agent) row of P logits (goal_manager.py:328-333)
// (`logits` and `logp` may be the same buffer: the raw logits are normalised in place)
__global__ void __launch_bounds__(NT) k_dest_finish(float* logits, const uint8_t* __restrict__ map_valid,
                                                    const uint8_t* __restrict__ map_type, const uint8_t* __restrict__ agent_type,
                                                    const uint8_t* __restrict__ dist_valid, int S, int A, int P,
                                                    float* logp, float* __restrict__ probs) {
  const int row = blockIdx.x * NWARP + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= S * A) return;
  const int s = row / A;
  const uint8_t* at = agent_type + (size_t)row * 3;
  const bool veh = at[0] != 0, ped = at[1] != 0, cyc = at[2] != 0, dv = dist_valid[row] != 0;
  float* lg = logits + (size_t)row * P;
  auto masked = [&](int p) -> float {
    const uint8_t* mt = map_type + ((size_t)s * P + p) * TB_PL_TYPE;
    const bool t0 = mt[0], t1 = mt[1], t2 = mt[2], t3 = mt[3], t4 = mt[4];
    const bool type_mask = !(map_valid[(size_t)s * P + p] && (t0 || t1 || t2 || t3 || t4));
    const bool attn_mask = (veh && t3) || (ped && (t0 || t1 || t2 || t3)) || (cyc && (t0 || t1 || t2));
    float v = (type_mask || attn_mask) ? -INFINITY : lg[p];
    if (!dv) v = 0.f;
    return v;
  };
  float mx = -INFINITY;
  for (int p = lane; p < P; p += 32) mx = fmaxf(mx, masked(p));
  mx = warp_max(mx);
  const bool all_masked = mx == -INFINITY;  // rows without any admissible polyline become uniform
  if (all_masked) mx = 0.f;
  float sum = 0.f;
  for (int p = lane; p < P; p += 32) sum += expf((all_masked ? 0.f : masked(p)) - mx);
  sum = warp_sum(sum);
  const float lse = mx + logf(sum);
  for (int p = lane; p < P; p += 32) {
    const float v = (all_masked ? 0.f : masked(p)) - lse;
    logp[(size_t)row * P + p] = v;
    probs[(size_t)row * P + p] = expf(v);
  }
}

}  // namespace tb

using namespace tb;

static int gru_base(int which) {
  switch (which) {
    case TB_GRU_POLICY: return tbw::model_agent_temporal_rnn_weight_ih_l0;
    case TB_GRU_LATENT_PRIOR: return tbw::model_latent_encoder_agent_temporal_prior_rnn_weight_ih_l0;
    case TB_GRU_LATENT_POST: return tbw::model_latent_encoder_agent_temporal_post_rnn_weight_ih_l0;
    case TB_GRU_DEST: return tbw::model_goal_manager_goal_predictor_gru_as_rnn_weight_ih_l0;
  }
  return -1;
}

extern "C" size_t tb_gru_workspace_bytes(int32_t n_batch, int32_t n_agent) {
  if (n_batch < 1 || n_agent < 1) return 0;
  const size_t n_cta = ((size_t)n_batch * n_agent + 127) / 128;
  return n_cta * 3 * 128 * 128 * sizeof(float);  // hidden state of the 3 layers, 128 rows per CTA
}

extern "C" int32_t tb_gru_sequence(int32_t which, int32_t mode, const float* x, const uint8_t* valid, int32_t n_batch,
                                   int32_t n_frame, int32_t n_agent, int32_t t_stride, const float* packed, void* workspace,
                                   float* out, uint8_t* out_valid, void* stream) {
  if (!x || !valid || !packed || !out || !out_valid) return TB_ERR_NULL;
  if (gru_base(which) < 0 || (mode != 0 && mode != 1) || n_batch < 1 || n_batch > 65535 || n_frame < 1 || n_agent < 1 || t_stride < 1)
    return TB_ERR_BAD_SHAPE;
  if (!aligned16(x) || !aligned16(packed) || !aligned16(out) || (workspace && !aligned16(workspace))) return TB_ERR_ALIGN;
  if (workspace && tc_enabled())  // tensor-core kernel (tb_tc_xlayer.cu); without a workspace: the fp32 row-tile kernel
    return launch_gru_seq_tc(which, mode, x, valid, n_batch, n_frame, n_agent, t_stride, packed, gru_base(which), workspace, out,
                             out_valid, (cudaStream_t)stream);
  static std::atomic<uint64_t> attr_set{0};
  if (!smem_attr_done(attr_set)) {
    if (!set_max_smem(k_gru_seq<HR>, (int)sizeof(GruSmem))) return TB_ERR_LAUNCH;
    smem_attr_mark(attr_set);
  }
  const int n_t = (n_frame + t_stride - 1) / t_stride;
  dim3 grid((n_agent + HR - 1) / HR, n_batch);
  k_gru_seq<HR><<<grid, NT, sizeof(GruSmem), (cudaStream_t)stream>>>(x, valid, n_frame, n_agent, t_stride, n_t, packed + gru_base(which),
                                                                    mode, out, out_valid);
  count_launch();
  return launch_status();
}

extern "C" int32_t tb_mlp_head(int32_t which, const float* x, const uint8_t* valid, int64_t n_row, const float* packed, float* y,
                               void* stream) {
  if (!x || !valid || !packed || !y) return TB_ERR_NULL;
  if (n_row < 1) return TB_ERR_BAD_SHAPE;
  if (!aligned16(x) || !aligned16(packed)) return TB_ERR_ALIGN;
  int w1, b1, w2, b2, n_out;
  switch (which) {
    case TB_MLP_LATENT_PRIOR_MEAN:
      w1 = tbw::model_latent_encoder_latent_prior_dist_mlp_mean_fc_layers_0_weight, b1 = tbw::model_latent_encoder_latent_prior_dist_mlp_mean_fc_layers_0_bias;
      w2 = tbw::model_latent_encoder_latent_prior_dist_mlp_mean_fc_layers_2_weight, b2 = tbw::model_latent_encoder_latent_prior_dist_mlp_mean_fc_layers_2_bias;
      n_out = TB_LATENT;
      break;
    case TB_MLP_LATENT_POST_MEAN:
      w1 = tbw::model_latent_encoder_latent_post_dist_mlp_mean_fc_layers_0_weight, b1 = tbw::model_latent_encoder_latent_post_dist_mlp_mean_fc_layers_0_bias;
      w2 = tbw::model_latent_encoder_latent_post_dist_mlp_mean_fc_layers_2_weight, b2 = tbw::model_latent_encoder_latent_post_dist_mlp_mean_fc_layers_2_bias;
      n_out = TB_LATENT;
      break;
    default: return TB_ERR_BAD_SHAPE;
  }
  k_mlp_head<HR><<<(unsigned)((n_row + HR - 1) / HR), NT, 0, (cudaStream_t)stream>>>(x, valid, n_row, packed + w1, packed + b1, packed + w2,
                                                                                    packed + b2, n_out, y);
  count_launch();
  return launch_status();
}

extern "C" size_t tb_dest_workspace_bytes(int32_t n_scene, int32_t n_agent, int32_t n_pl) {
  if (n_scene < 1 || n_agent < 1 || n_pl < 1) return 0;
  return ((size_t)n_scene * n_pl * D + (size_t)n_scene * n_agent * D) * sizeof(float) + 1024 + dest_lists_bytes(n_scene, n_pl);
}

extern "C" int32_t tb_dest_logits(int32_t n_scene, int32_t n_agent, int32_t n_pl, const float* map_feature,
                                  const uint8_t* map_feature_valid, const uint8_t* map_type, const float* tgt,
                                  const uint8_t* tgt_valid, const uint8_t* agent_type, const float* packed, void* workspace,
                                  float* logp, float* probs, void* stream) {
  if (!map_feature || !map_feature_valid || !map_type || !tgt || !tgt_valid || !agent_type || !packed || !workspace || !logp || !probs)
    return TB_ERR_NULL;
  if (n_scene < 1 || n_scene > 65535 || n_agent < 1 || n_agent > 65535 || n_pl < 1) return TB_ERR_BAD_SHAPE;
  if (!aligned16(map_feature) || !aligned16(tgt) || !aligned16(packed) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return TB_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  float* U = reinterpret_cast<float*>(workspace);
  float* V = U + (((size_t)n_scene * n_pl * D + 63) & ~(size_t)63);
  const float* w0 = packed + tbw::model_goal_manager_goal_predictor_mlp_fc_layers_0_weight;  // [128, 256]: cat[map_feature, tgt]
  const long rows_u = (long)n_scene * n_pl, rows_v = (long)n_scene * n_agent;
  k_row_linear<HR><<<(unsigned)((rows_u + HR - 1) / HR), NT, 0, st>>>(map_feature, rows_u, w0, 0,
                                                                     packed + tbw::model_goal_manager_goal_predictor_mlp_fc_layers_0_bias, U);
  count_launch();
  k_row_linear<HR><<<(unsigned)((rows_v + HR - 1) / HR), NT, 0, st>>>(tgt, rows_v, w0, D, nullptr, V);
  count_launch();
  if (tc_enabled()) {  // pairwise MLP on the tensor pipe (tb_tc_xlayer.cu); raw logits staged in `logp`
    int32_t* lists = reinterpret_cast<int32_t*>(
        (reinterpret_cast<uintptr_t>(V + (((size_t)n_scene * n_agent * D + 63) & ~(size_t)63)) + 255) & ~(uintptr_t)255);
    const int rc = launch_dest_pairs_tc(U, V, n_scene, n_agent, n_pl, packed, logp, map_feature_valid, map_type, agent_type, tgt_valid,
                                        lists, st);
    if (rc != TB_OK) return rc;
  } else {
    static std::atomic<uint64_t> attr_set{0};
    if (!smem_attr_done(attr_set)) {
      if (!set_max_smem(k_dest_pairs, (int)sizeof(PairSmem))) return TB_ERR_LAUNCH;
      smem_attr_mark(attr_set);
    }
    dim3 grid((n_pl + PR - 1) / PR, n_agent, n_scene);
    k_dest_pairs<<<grid, NT, sizeof(PairSmem), st>>>(U, V, n_pl, n_agent, packed, logp);
    count_launch();
  }
  const int rows = n_scene * n_agent;
  k_dest_finish<<<(rows + NWARP - 1) / NWARP, NT, 0, st>>>(logp, map_feature_valid, map_type, agent_type, tgt_valid, n_scene, n_agent, n_pl,
                                                         logp, probs);
  count_launch();
  return launch_status();
}
