// The step right after the rollout: Waymo post-processing of the K joint futures and the packing of the
// per-scene tensors the Waymo motion-metrics op consumes.  HBM-bound gather / scatter work plus tiny per-agent sequential
// selections; no tensor cores.
//
//   k_post_process  one CTA per (scene, agent): WaymoPostProcessing.forward (data_modules/waymo_post_processing.py:33-81) --
//                   score normalisation, mode selection when n_pred > k_pred (mtr_nms :126-171 or top-k :173-193), mpa_nms
//                   (:83-124, the reference's triple Python loop), temperature softmax, and the [S,Tf,A,k,.] output layout.
//   k_womd_pack     one CTA per scene: WOMDMetrics.update (models/metrics/womd.py:60-145): agents to predict first, then the
//                   other fully observed agents (ordered compaction), down-sampled predictions, GT tracks, object types.
#include "tb_host.h"

namespace tb {
namespace post {

constexpr int MAXP = 32;  // modes per agent
constexpr int NT = 128;

struct PostSmem {
  float score[MAXP];      // normalised scores of the n_pred input modes
  float dist[MAXP][MAXP]; // pairwise distance (ADE over the future steps, or final displacement)
  int sel[MAXP];          // selected input mode of every output slot
  float out_score[MAXP];
};

// element (s, a, k, t) of the trajectories: base + s * str_s + a * str_a + k * str_k + t * 4
struct TrajView {
  const float* base;
  long long str_s, str_a, str_k;
};

__device__ __forceinline__ const float4* traj_ptr(const TrajView& v, int s, int a, int k) {
  return reinterpret_cast<const float4*>(v.base + s * v.str_s + a * v.str_a + k * v.str_k);
}

// distance matrix between the modes listed in `idx` (n of them) into sm.dist: warps take (i, j) pairs, lanes the steps
__device__ void pair_distances(PostSmem& sm, const TrajView& tv, int s, int a, const int* idx, int n, int n_step, bool use_ade) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int p = warp; p < n * n; p += NT / 32) {
    const int i = p / n, j = p - i * n;
    if (j < i) continue;
    float acc = 0.f;
    if (i != j) {
      const float4* ti = traj_ptr(tv, s, a, idx ? idx[i] : i);
      const float4* tj = traj_ptr(tv, s, a, idx ? idx[j] : j);
      if (use_ade) {
        for (int t = lane; t < n_step; t += 32) {
          const float4 u = ti[t], w = tj[t];
          const float dx = __fsub_rn(u.x, w.x), dy = __fsub_rn(u.y, w.y);
          acc += __fsqrt_rn(__fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        acc = acc / (float)n_step;
      } else {
        const float4 u = ti[n_step - 1], w = tj[n_step - 1];
        const float dx = __fsub_rn(u.x, w.x), dy = __fsub_rn(u.y, w.y);
        acc = __fsqrt_rn(__fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
      }
    }
    if (lane == 0) sm.dist[i][j] = sm.dist[j][i] = acc;
  }
}

__device__ __forceinline__ float type_thresh(const uint8_t* ty, const float* th, int n) {
  float t = 0.f;
  for (int i = 0; i < n; ++i) t = __fadd_rn(t, ty[i] ? th[i] : 0.f);
  return t;
}

__global__ void __launch_bounds__(NT) k_post_process(int n_agent, int n_pred, int n_step, TrajView tv, const float* __restrict__ scores,
                                                     const uint8_t* __restrict__ valid, const uint8_t* __restrict__ agent_type,
                                                     TbPostCfg cfg, float* __restrict__ w_trajs, float* __restrict__ w_yaw,
                                                     float* __restrict__ w_spd, float* __restrict__ w_scores, int32_t* __restrict__ mode_idx) {
  __shared__ PostSmem sm;
  const int a = blockIdx.x, s = blockIdx.y;
  const int K = cfg.k_pred < n_pred ? cfg.k_pred : n_pred;  // output modes
  const uint8_t* ty = agent_type + ((size_t)s * n_agent + a) * 3;
  const float* sc_in = scores + ((size_t)s * n_agent + a) * n_pred;
  if (threadIdx.x == 0) {
    float sum = 0.f;
    for (int k = 0; k < n_pred; ++k) sum = __fadd_rn(sum, sc_in[k]);
    for (int k = 0; k < n_pred; ++k) sm.score[k] = sc_in[k] / sum;
  }
  const bool select = n_pred > cfg.k_pred;
  const bool use_mtr = select && cfg.n_mtr > 0;
  if (use_mtr) pair_distances(sm, tv, s, a, nullptr, n_pred, n_step, cfg.use_ade != 0);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (!select) {
      for (int k = 0; k < K; ++k) {
        sm.sel[k] = k;
        sm.out_score[k] = sm.score[k];
      }
    } else {
      float work[MAXP];
      for (int k = 0; k < n_pred; ++k) work[k] = sm.score[k];
      const float th = use_mtr ? type_thresh(ty, cfg.mtr_nms_thresh, cfg.n_mtr) : 0.f;
      for (int k = 0; k < K; ++k) {
        int best = 0;
        for (int j = 1; j < n_pred; ++j)
          if (work[j] > work[best]) best = j;  // first maximum, like torch.max
        if (use_mtr)
          for (int j = 0; j < n_pred; ++j)  // suppress everything close to the pick: x 0.01 (waymo_post_processing.py:157-160)
            work[j] = __fmul_rn(work[j], __fadd_rn(sm.dist[best][j] < th ? 0.f : 0.99f, 0.01f));
        work[best] = -1.f;
        sm.sel[k] = best;
      }
      float sum = 0.f;
      for (int k = 0; k < K; ++k) sum = __fadd_rn(sum, sm.score[sm.sel[k]]);
      for (int k = 0; k < K; ++k) sm.out_score[k] = sm.score[sm.sel[k]] / sum;
    }
  }
  __syncthreads();
  if (cfg.n_mpa > 0) {
    pair_distances(sm, tv, s, a, sm.sel, K, n_step, cfg.use_ade != 0);
    __syncthreads();
    if (threadIdx.x == 0) {
      float* sc = sm.out_score;
      if (valid[(size_t)s * n_agent + a]) {
        const float th = type_thresh(ty, cfg.mpa_nms_thresh, cfg.n_mpa);
        int order[MAXP];
        for (int k = 0; k < K; ++k) order[k] = k;
        for (int i = 1; i < K; ++i) {  // descending insertion sort of the scores as they are BEFORE the in-place edits
          const int o = order[i];
          int j = i - 1;
          while (j >= 0 && sc[order[j]] < sc[o]) {
            order[j + 1] = order[j];
            --j;
          }
          order[j + 1] = o;
        }
        for (int q = 0; q < K; ++q) {
          const int k = order[q];
          bool hit = false;
          for (int j = 0; j < K; ++j) hit |= (sm.dist[k][j] < th) && (sc[j] > sc[k]);
          if (hit) sc[k] = 1e-3f;
        }
      }
      float sum = 0.f;
      for (int k = 0; k < K; ++k) sum = __fadd_rn(sum, sc[k]);
      for (int k = 0; k < K; ++k) sc[k] = sc[k] / sum;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float* sc = sm.out_score;
    if (cfg.score_temperature > 0.f) {  // softmax(log(scores) / temperature)
      float z[MAXP], m = -3.0e38f, sum = 0.f;
      for (int k = 0; k < K; ++k) {
        z[k] = logf(sc[k]) / cfg.score_temperature;
        m = fmaxf(m, z[k]);
      }
      for (int k = 0; k < K; ++k) {
        z[k] = expf(z[k] - m);
        sum += z[k];
      }
      for (int k = 0; k < K; ++k) sc[k] = z[k] / sum;
    }
    for (int k = 0; k < K; ++k) {
      w_scores[((size_t)s * n_agent + a) * K + k] = sc[k];
      if (mode_idx) mode_idx[((size_t)s * n_agent + a) * K + k] = sm.sel[k];
    }
  }
  // [S, Tf, A, k, .] outputs (trajs.movedim(3, 1), :68-79)
  for (int e = threadIdx.x; e < n_step * K; e += NT) {
    const int t = e / K, k = e - t * K;
    const float4 v = traj_ptr(tv, s, a, sm.sel[k])[t];
    const size_t o = (((size_t)s * n_step + t) * n_agent + a) * K + k;
    reinterpret_cast<float2*>(w_trajs)[o] = make_float2(v.x, v.y);
    if (w_yaw) w_yaw[o] = v.z;
    if (w_spd) w_spd[o] = v.w;
  }
}

// ------------------------------------------------------------------------------------------------------------------------
constexpr int WMAXA = 256;

__global__ void __launch_bounds__(256) k_womd_pack(TbWomdIn in, TbWomdOut out) {
  __shared__ int order[WMAXA];
  __shared__ int n_first, n_all;
  const int s = blockIdx.x;
  const int A = in.n_agent, K = in.n_pred, Tf = in.n_step_future, Tg = in.n_step_gt_frames;
  const int n_gt = in.step_gt + 1;                                 // frames kept in the GT tensors
  const int n_ds = (in.step_gt - in.step_current - 4 + 4) / 5;     // len(range(4, track_future_samples, 5))
  if (threadIdx.x == 0) {
    int n = 0;
    for (int a = 0; a < A; ++a)
      if (in.agent_role[((size_t)s * A + a) * 3 + 2]) order[n++] = a;
    n_first = n;
    for (int a = 0; a < A; ++a) {
      if (in.agent_role[((size_t)s * A + a) * 3 + 2]) continue;
      bool all = true;
      for (int t = 0; t <= in.step_current; ++t) all &= in.agent_valid[((size_t)s * Tg + t) * A + a] != 0;
      if (all) order[n++] = a;
    }
    n_all = n;
    if (n_first > in.m_joint && out.overflow) atomicAdd(out.overflow, 1);
  }
  __syncthreads();
  const size_t off = (size_t)s * (size_t)out.scene_stride_bytes;  // this scene's record
  float* p_traj = reinterpret_cast<float*>(reinterpret_cast<char*>(out.prediction_trajectory) + off);
  float* p_score = reinterpret_cast<float*>(reinterpret_cast<char*>(out.prediction_score) + off);
  float* g_traj = reinterpret_cast<float*>(reinterpret_cast<char*>(out.ground_truth_trajectory) + off);
  uint8_t* g_valid = out.ground_truth_is_valid + off;
  uint8_t* p_mask = out.prediction_ground_truth_indices_mask + off;
  float* o_type = reinterpret_cast<float*>(reinterpret_cast<char*>(out.object_type) + off);
  const int n_p = n_first < in.m_joint ? n_first : in.m_joint;
  // predictions: [m_joint, K, 1, n_ds, 2] = waymo_trajs[s, 4::5][:n_ds] of the agents to predict, zeros elsewhere
  for (int e = threadIdx.x; e < in.m_joint * K * n_ds; e += blockDim.x) {
    const int slot = e / (K * n_ds), r = e - slot * (K * n_ds), k = r / n_ds, i = r - k * n_ds;
    float2 v = make_float2(0.f, 0.f);
    if (slot < n_p) {
      const int t = 4 + 5 * i;
      v = reinterpret_cast<const float2*>(in.waymo_trajs)[(((size_t)s * Tf + t) * A + order[slot]) * K + k];
    }
    reinterpret_cast<float2*>(p_traj)[e] = v;
  }
  for (int e = threadIdx.x; e < in.m_joint * K; e += blockDim.x) {
    const int slot = e / K, k = e - slot * K;
    float v = 0.f;
    if (slot < n_p) v = in.waymo_scores ? in.waymo_scores[((size_t)s * A + order[slot]) * K + k] : 1.0f / (float)K;
    p_score[e] = v;
  }
  for (int e = threadIdx.x; e < in.m_joint; e += blockDim.x) p_mask[e] = e < n_p;
  // ground truth: [A, n_gt, 7] = (x, y, length, width, yaw, vx, vy), [A, n_gt] valid, [A] type in {1, 2, 3}
  for (int e = threadIdx.x; e < A * n_gt; e += blockDim.x) {
    const int slot = e / n_gt, t = e - slot * n_gt;
    float v[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    uint8_t ok = 0;
    if (slot < n_all) {
      const int a = order[slot];
      const size_t g = ((size_t)s * Tg + t) * A + a;
      const float2 p = reinterpret_cast<const float2*>(in.agent_pos)[g], vel = reinterpret_cast<const float2*>(in.agent_vel)[g];
      const float* sz = in.agent_size + ((size_t)s * A + a) * 3;
      v[0] = p.x; v[1] = p.y; v[2] = sz[0]; v[3] = sz[1]; v[4] = in.agent_yaw[g]; v[5] = vel.x; v[6] = vel.y;
      ok = in.agent_valid[g];
    }
#pragma unroll
    for (int c = 0; c < 7; ++c) g_traj[(size_t)e * 7 + c] = v[c];
    g_valid[e] = ok;
  }
  for (int slot = threadIdx.x; slot < A; slot += blockDim.x) {
    float ty = 0.f;
    if (slot < n_all) {
      const uint8_t* t3 = in.agent_type + ((size_t)s * A + order[slot]) * 3;
      ty = t3[0] ? 1.f : (t3[1] ? 2.f : (t3[2] ? 3.f : 1.f));  // argmax of the one-hot + 1 (first maximum)
    }
    o_type[slot] = ty;
  }
}

}  // namespace post
}  // namespace tb

using namespace tb;

extern "C" int32_t tb_post_process(int32_t n_scene, int32_t n_agent, int32_t n_pred, int32_t n_step, const float* trajs,
                                   int64_t stride_scene, int64_t stride_agent, int64_t stride_mode, const float* scores,
                                   const uint8_t* valid, const uint8_t* agent_type, const TbPostCfg* cfg, float* waymo_trajs,
                                   float* waymo_yaw, float* waymo_spd, float* waymo_scores, int32_t* mode_idx, void* stream) {
  if (!trajs || !scores || !valid || !agent_type || !cfg || !waymo_trajs || !waymo_scores) return TB_ERR_NULL;
  if (n_scene < 1 || n_agent < 1 || n_pred < 1 || n_step < 1 || n_scene > 65535) return TB_ERR_BAD_SHAPE;
  if (n_pred > post::MAXP || cfg->k_pred < 1 || cfg->n_mtr < 0 || cfg->n_mtr > 3 || cfg->n_mpa < 0 || cfg->n_mpa > 3) return TB_ERR_BAD_SHAPE;
  if (!aligned16(trajs) || (stride_scene & 3) || (stride_agent & 3) || (stride_mode & 3)) return TB_ERR_ALIGN;
  post::TrajView tv{trajs, stride_scene, stride_agent, stride_mode};
  post::k_post_process<<<dim3(n_agent, n_scene), post::NT, 0, (cudaStream_t)stream>>>(n_agent, n_pred, n_step, tv, scores, valid, agent_type,
                                                                                   *cfg, waymo_trajs, waymo_yaw, waymo_spd,
                                                                                   waymo_scores, mode_idx);
  count_launch();
  return launch_status();
}

extern "C" size_t tb_womd_record_bytes(int32_t n_agent, int32_t n_pred, int32_t step_gt, int32_t step_current, int32_t m_joint,
                                       int64_t* offsets6) {
  // one scene's record: [prediction_trajectory | prediction_score | ground_truth_trajectory | object_type (all f32) |
  //                      ground_truth_is_valid | prediction_ground_truth_indices_mask (u8)], padded to 16 bytes
  const int64_t n_ds = (step_gt - step_current) / 5, n_gt = step_gt + 1;
  int64_t o = 0, off[6];
  off[0] = o; o += (int64_t)m_joint * n_pred * n_ds * 2 * 4;
  off[1] = o; o += (int64_t)m_joint * n_pred * 4;
  off[2] = o; o += (int64_t)n_agent * n_gt * 7 * 4;
  off[5] = o; o += (int64_t)n_agent * 4;
  off[3] = o; o += (int64_t)n_agent * n_gt;
  off[4] = o; o += (int64_t)m_joint;
  if (offsets6)
    for (int i = 0; i < 6; ++i) offsets6[i] = off[i];
  return (size_t)((o + 15) & ~(int64_t)15);
}

extern "C" int32_t tb_womd_pack(int32_t n_scene, const TbWomdIn* in, const TbWomdOut* out, void* stream) {
  if (!in || !out) return TB_ERR_NULL;
  if (!in->agent_role || !in->agent_valid || !in->agent_pos || !in->agent_size || !in->agent_yaw || !in->agent_vel || !in->agent_type ||
      !in->waymo_trajs || !out->prediction_trajectory || !out->prediction_score || !out->ground_truth_trajectory ||
      !out->ground_truth_is_valid || !out->prediction_ground_truth_indices_mask || !out->object_type)
    return TB_ERR_NULL;
  if (n_scene < 1 || in->n_agent < 1 || in->n_agent > post::WMAXA || in->n_pred < 1 || in->m_joint < 1) return TB_ERR_BAD_SHAPE;
  if (in->step_gt + 1 > in->n_step_gt_frames || in->step_gt - in->step_current > in->n_step_future || in->step_current < 0) return TB_ERR_BAD_SHAPE;
  if ((in->step_gt - in->step_current) % 5 != 0) return TB_ERR_UNSUPPORTED;
  post::k_womd_pack<<<n_scene, 256, 0, (cudaStream_t)stream>>>(*in, *out);
  count_launch();
  return launch_status();
}
