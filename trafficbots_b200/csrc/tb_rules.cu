// Optional traffic-rule checks and the collision term of the differentiable reward, evaluated after
// the rollout in parallel over all (scene-mode, step) pairs -- see `tb_rule_checks` in include/trafficbots_b200.h.
//
// Reference semantics: utils/traffic_rule_checker.py:122-335,420-472,518-604 and utils/rewards.py:49-115.  The boolean
// outputs are thresholded fp32 geometry, so every comparison is evaluated with the reference's operation order and with
// explicitly rounded multiplies / adds (no FMA contraction) where torch's CPU kernels round separately.
//
// Kernels (all HBM / L2 bound integer-and-geometry work, no tensor cores):
//   k_rule_compact  one CTA per scene: road-edge segments (polyline types 4, 5, 7) and lane-centre nodes (types 0-2) of the
//                   valid map nodes compacted into dense lists (ordered block scan), read by all steps / modes of the scene.
//   k_rule_step     one CTA per (step, scene-mode): rebuilds the post-override state, then collided (A x A separating-axis
//                   tests), run_road_edge (edges x vehicles with a conservative distance cull before the exact ccw tests),
//                   run_red_light, the raw passive predicate, and the 5-circle collision term of the reward.
//   k_rule_sticky   one thread per (scene-mode, agent): running ORs over the steps and the passive counter (> 20 steps).
#include "tb_host.h"

namespace tb {
namespace rl {

constexpr int MAXA = 128;
constexpr int NT = 256;

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
// torch.norm over a last dim of 2 on the CPU build used for the fixtures: sqrt(fma(y, y, x * x))
__device__ __forceinline__ float norm2(float x, float y) { return __fsqrt_rn(__fmaf_rn(y, y, mul(x, x))); }
// ccw(A, B, C) (traffic_rule_checker.py:609-610)
__device__ __forceinline__ bool ccw(float ax, float ay, float bx, float by, float cx, float cy) {
  return mul(sub(cy, ay), sub(bx, ax)) > mul(sub(by, ay), sub(cx, ax));
}

// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_rule_compact(int n_pl, const uint8_t* __restrict__ map_valid, const uint8_t* __restrict__ map_type,
                                                     const float* __restrict__ map_pos, const float* __restrict__ map_dir,
                                                     float4* __restrict__ edges, float2* __restrict__ lanes, int32_t* __restrict__ counts) {
  const int s = blockIdx.x;
  const int n_node = n_pl * TB_PL_NODE;
  __shared__ int warp_e[NT / 32], warp_l[NT / 32];
  __shared__ int base_e, base_l;
  if (threadIdx.x == 0) base_e = base_l = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int n0 = 0; n0 < n_node; n0 += NT) {
    const int n = n0 + threadIdx.x;
    bool is_e = false, is_l = false;
    float2 p = make_float2(0.f, 0.f), d = p;
    if (n < n_node) {
      const size_t node = (size_t)s * n_node + n;
      const size_t pl = (size_t)s * n_pl + n / TB_PL_NODE;
      if (map_valid[node]) {
        const uint8_t* ty = map_type + pl * TB_PL_TYPE;
        is_e = ty[4] | ty[5] | ty[7];
        is_l = ty[0] | ty[1] | ty[2];
        if (is_e | is_l) {
          p = reinterpret_cast<const float2*>(map_pos)[node];
          d = reinterpret_cast<const float2*>(map_dir)[node];
        }
      }
    }
    const unsigned be = __ballot_sync(0xffffffffu, is_e), bl = __ballot_sync(0xffffffffu, is_l);
    if (lane == 0) {
      warp_e[w] = __popc(be);
      warp_l[w] = __popc(bl);
    }
    __syncthreads();
    int off_e = base_e, off_l = base_l;
    for (int k = 0; k < w; ++k) {
      off_e += warp_e[k];
      off_l += warp_l[k];
    }
    const unsigned below = (1u << lane) - 1u;
    if (is_e) edges[(size_t)s * n_node + off_e + __popc(be & below)] = make_float4(p.x, p.y, add(p.x, d.x), add(p.y, d.y));
    if (is_l) lanes[(size_t)s * n_node + off_l + __popc(bl & below)] = p;
    __syncthreads();
    if (threadIdx.x == 0) {
      int te = 0, tl = 0;
      for (int k = 0; k < NT / 32; ++k) {
        te += warp_e[k];
        tl += warp_l[k];
      }
      base_e += te;
      base_l += tl;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    counts[2 * s] = base_e;
    counts[2 * s + 1] = base_l;
  }
}

// ------------------------------------------------------------------------------------------------------------------------
struct StepSmem {
  // post-override state seen by TrafficRuleChecker.check
  float x[MAXA], y[MAXA], c[MAXA], s[MAXA], spd[MAXA];
  float bx[MAXA][4], by[MAXA][4];  // box corners (1.1 x size): rear-right, front-right, front-left, rear-left
  float cull[MAXA];                // half diagonal of the box + margin
  float rl_len[MAXA], rl_wid[MAXA];
  uint8_t valid[MAXA], veh[MAXA], ped[MAXA];
  // pre-override prediction (the state DifferentiableReward.get sees)
  float px[MAXA], py[MAXA], pc[MAXA], ps[MAXA], pd[MAXA], pr[MAXA];
  uint8_t pvalid[MAXA];
  uint32_t sep[MAXA][MAXA / 32];   // bit j of row i: an edge of box i separates box j
  int f_edge[MAXA], f_red[MAXA], f_near[MAXA], f_tl_ahead[MAXA], f_ag_ahead[MAXA];
  unsigned col_bits[MAXA];         // collision term, max-reduced (non-negative floats order like their bit patterns)
  float col_sum[MAXA];
  int n_pvalid;
};

__global__ void __launch_bounds__(NT) k_rule_step(TbDims dm, TbRuleIn in, TbRuleOut out, const float4* __restrict__ edges,
                                                  const float2* __restrict__ lanes, const int32_t* __restrict__ counts) {
  __shared__ StepSmem sm;
  const int t = blockIdx.x + 1;  // decode step 1..T, output slot t-1
  const int b = blockIdx.y;
  const int A = dm.n_agent, T = dm.n_step, Tg = dm.n_step_gt, TL = dm.n_tl;
  const int scene = b / dm.n_mode;
  const int tid = threadIdx.x;
  const int en = in.enable_mask;
  const bool want_reward = in.w_collision > 0.f && out.diffbar_rewards != nullptr;

  for (int a = tid; a < A; a += NT) {
    const size_t row = ((size_t)b * A + a) * T;
    bool killed = false;  // running OR of Dynamics.kill before this step (dynamics.py:161-167)
    for (int s = 1; s < t; ++s) {
      const bool gv = s < Tg ? in.gt_valid[((size_t)scene * Tg + s) * A + a] != 0 : false;
      killed |= (in.outside_map_this_step[row + s - 1] != 0) && !gv;
    }
    const bool ovr = t < Tg && in.override_masks[row + t - 1] != 0;
    const bool m = ovr && !killed;
    const bool vpre = in.valid[row + t - 1] != 0;
    const float4 pred = reinterpret_cast<const float4*>(in.preds)[row + t - 1];
    float4 st = pred;
    if (m) {
      const size_t g = ((size_t)scene * Tg + t) * A + a;
      const float2 gp = reinterpret_cast<const float2*>(in.gt_pos)[g];
      st = make_float4(gp.x, gp.y, in.gt_yaw[g], in.gt_spd[g]);
    }
    const float* size = in.agent_size + ((size_t)scene * A + a) * 3;
    const uint8_t* ty = in.agent_type + ((size_t)scene * A + a) * 3;
    float sn, cs;
    sincosf(st.z, &sn, &cs);
    sm.x[a] = st.x; sm.y[a] = st.y; sm.c[a] = cs; sm.s[a] = sn; sm.spd[a] = st.w;
    sm.valid[a] = vpre || m;
    sm.veh[a] = ty[0];
    sm.ped[a] = ty[1];
    const float L = mul(size[0], in.collision_size_scale), W = mul(size[1], in.collision_size_scale);
    const float fx = mul(mul(0.5f, L), cs), fy = mul(mul(0.5f, L), sn);
    const float rx = mul(mul(0.5f, W), sn), ry = mul(mul(0.5f, W), -cs);
    sm.bx[a][0] = add(st.x, add(-fx, rx)); sm.by[a][0] = add(st.y, add(-fy, ry));
    sm.bx[a][1] = add(st.x, add(fx, rx));  sm.by[a][1] = add(st.y, add(fy, ry));
    sm.bx[a][2] = add(st.x, sub(fx, rx));  sm.by[a][2] = add(st.y, sub(fy, ry));
    sm.bx[a][3] = add(st.x, sub(-fx, rx)); sm.by[a][3] = add(st.y, sub(-fy, ry));
    sm.cull[a] = 0.5f * sqrtf(L * L + W * W) + 0.05f;
    sm.rl_len[a] = mul(mul(size[0], 0.5f), 0.6f);
    sm.rl_wid[a] = mul(mul(size[1], 0.5f), 1.8f);
    // reward: pre-override prediction with the old valid (waymo_motion.py:325-331)
    float psn, pcs;
    sincosf(pred.z, &psn, &pcs);
    sm.px[a] = pred.x; sm.py[a] = pred.y; sm.pc[a] = pcs; sm.ps[a] = psn;
    const float wmin = fminf(size[0], size[1]), lmax = fmaxf(size[0], size[1]);
    sm.pd[a] = sub(lmax, wmin) / 4.0f;
    sm.pr[a] = add(wmin / 2.0f, 1.1920928955078125e-07f);
    sm.pvalid[a] = vpre;
    sm.f_edge[a] = sm.f_red[a] = sm.f_near[a] = sm.f_tl_ahead[a] = sm.f_ag_ahead[a] = 0;
    sm.col_bits[a] = 0u;
    sm.col_sum[a] = 0.f;
#pragma unroll
    for (int w = 0; w < MAXA / 32; ++w) sm.sep[a][w] = 0u;
  }
  if (tid == 0) sm.n_pvalid = 0;
  __syncthreads();
  if (want_reward && !in.reduce_collision_with_max) {
    int n = 0;
    for (int a = tid; a < A; a += NT) n += sm.pvalid[a];
    if (n) atomicAdd(&sm.n_pvalid, n);
  }

  // ---- collided: separating-axis test, lines of box i against the corners of box j (traffic_rule_checker.py:131-160) ----
  if (en & 1) {
    for (int idx = tid; idx < A * A; idx += NT) {
      const int i = idx / A, j = idx - i * A;
      if (i == j || !sm.valid[i] || !sm.valid[j] || (sm.ped[i] && sm.ped[j])) continue;  // never collide: handled below
      {  // boxes whose circumscribed circles are apart are separated by one of their edge lines (SAT): skipped here and below
        const float dx = sm.x[i] - sm.x[j], dy = sm.y[i] - sm.y[j], reach = sm.cull[i] + sm.cull[j];
        if (dx * dx + dy * dy > reach * reach) continue;
      }
      bool sepd = false;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float x0 = sm.bx[i][k], y0 = sm.by[i][k], x1 = sm.bx[i][(k + 1) & 3], y1 = sm.by[i][(k + 1) & 3];
        const float la = sub(y1, y0), lb = sub(x0, x1), lc = sub(mul(x1, y0), mul(y1, x0));
        bool all_out = true;
#pragma unroll
        for (int q = 0; q < 4; ++q) all_out &= add(add(mul(la, sm.bx[j][q]), mul(lb, sm.by[j][q])), lc) > 0.f;
        sepd |= all_out;
      }
      if (sepd) atomicOr(&sm.sep[i][j >> 5], 1u << (j & 31));
    }
  }

  // ---- run_road_edge: box edges x road-edge segments, vehicles only (:163-196) ----
  if (en & 2) {
    const int nE = counts[2 * scene];
    const float4* E = edges + (size_t)scene * dm.n_pl * TB_PL_NODE;
    for (int e = tid; e < nE; e += NT) {
      const float4 sg = E[e];
      const float mx = 0.5f * (sg.x + sg.z), my = 0.5f * (sg.y + sg.w);
      const float hl = 0.5f * sqrtf((sg.z - sg.x) * (sg.z - sg.x) + (sg.w - sg.y) * (sg.w - sg.y));
      for (int a = 0; a < A; ++a) {
        if (!(sm.valid[a] && sm.veh[a])) continue;
        const float dx = sm.x[a] - mx, dy = sm.y[a] - my, reach = sm.cull[a] + hl;
        if (dx * dx + dy * dy > reach * reach) continue;  // conservative: the box and the segment cannot touch
        bool hit = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float ax = sm.bx[a][k], ay = sm.by[a][k], bx = sm.bx[a][(k + 1) & 3], by = sm.by[a][(k + 1) & 3];
          hit |= (ccw(ax, ay, sg.x, sg.y, sg.z, sg.w) != ccw(bx, by, sg.x, sg.y, sg.z, sg.w)) &&
                 (ccw(ax, ay, bx, by, sg.x, sg.y) != ccw(ax, ay, bx, by, sg.z, sg.w));
        }
        if (hit) sm.f_edge[a] = 1;
      }
    }
  }

  // ---- traffic lights: run_red_light (:199-258) and "red light ahead" of the passive check (:300-313) ----
  const int tl_t = t < in.n_tl_frame ? t : in.n_tl_frame - 1;
  if (en & (4 | 8)) {
    for (int idx = tid; idx < A * TL; idx += NT) {
      const int a = idx / TL, k = idx - a * TL;
      if (!(sm.valid[a] && sm.veh[a])) continue;
      const size_t tl = ((size_t)scene * in.n_tl_frame + tl_t) * TL + k;
      if (!in.tl_valid[tl]) continue;
      const uint8_t* stt = in.tl_state + tl * TB_TL_STATE;
      const float2 tp = reinterpret_cast<const float2*>(in.tl_pos)[tl];
      const float cs = sm.c[a], sn = sm.s[a];
      if ((en & 4) && stt[1]) {
        const float d0x = sub(tp.x, sm.x[a]), d0y = sub(tp.y, sm.y[a]);
        const bool in0 = fabsf(add(mul(d0x, cs), mul(d0y, sn))) < sm.rl_len[a] && fabsf(add(mul(d0x, sn), mul(d0y, -cs))) < sm.rl_wid[a];
        if (in0) {
          const float adv = mul(0.1f, sm.spd[a]);
          const float d1x = sub(tp.x, add(sm.x[a], mul(adv, cs))), d1y = sub(tp.y, add(sm.y[a], mul(adv, sn)));
          const bool in1 = fabsf(add(mul(d1x, cs), mul(d1y, sn))) < sm.rl_len[a] && fabsf(add(mul(d1x, sn), mul(d1y, -cs))) < sm.rl_wid[a];
          if (!in1) sm.f_red[a] = 1;
        }
      }
      if ((en & 8) && (stt[0] | stt[1] | stt[2] | stt[4])) {
        const float vx = sub(tp.x, sm.x[a]), vy = sub(tp.y, sm.y[a]);
        const float n = norm2(vx, vy);
        if (n < 10.f && add(mul(cs, vx), mul(sn, vy)) / n > 0.95f) sm.f_tl_ahead[a] = 1;
      }
    }
  }

  // ---- passive (:261-335): near a lane centre, another agent ahead ----
  if (en & 8) {
    const int nL = counts[2 * scene + 1];
    const float2* Ln = lanes + (size_t)scene * dm.n_pl * TB_PL_NODE;
    for (int e = tid; e < nL; e += NT) {
      const float2 p = Ln[e];
      for (int a = 0; a < A; ++a) {
        if (!(sm.valid[a] && sm.veh[a]) || sm.f_near[a]) continue;
        const float dx = sub(sm.x[a], p.x), dy = sub(sm.y[a], p.y);
        if (fabsf(dx) >= 2.f || fabsf(dy) >= 2.f) continue;
        if (norm2(dx, dy) < 2.f) sm.f_near[a] = 1;
      }
    }
    for (int idx = tid; idx < A * A; idx += NT) {
      const int i = idx / A, j = idx - i * A;
      if (i == j || !sm.valid[i] || !sm.valid[j] || !sm.veh[i]) continue;
      const float vx = sub(sm.x[j], sm.x[i]), vy = sub(sm.y[j], sm.y[i]);
      const float n = norm2(vx, vy);
      if (n < 10.f && add(mul(sm.c[i], vx), mul(sm.s[i], vy)) / n > 0.95f) sm.f_ag_ahead[i] = 1;
    }
  }

  // ---- collision term of the reward: 5 circles per agent (rewards.py:49-114) ----
  if (want_reward) {
    const float eps = 1.1920928955078125e-07f;
    for (int idx = tid; idx < A * A; idx += NT) {
      const int i = idx / A, j = idx - i * A;
      if (i == j || !sm.pvalid[i] || !sm.pvalid[j]) continue;
      {  // every circle centre lies within 2 d of its agent's centre: if the agents are further apart than that plus the radii,
         // all 25 distances exceed r_i + r_j and the relaxed overlap clamps to exactly 0
        const float dx = sm.px[i] - sm.px[j], dy = sm.py[i] - sm.py[j];
        const float reach = 2.f * (sm.pd[i] + sm.pd[j]) + sm.pr[i] + sm.pr[j] + 1e-2f;
        if (dx * dx + dy * dy > reach * reach) continue;
      }
      float dmin = 3.0e38f;
#pragma unroll
      for (int p = 0; p < 5; ++p) {
        const float kp = (float)(p - 2);
        const float cix = add(sm.px[i], mul(mul(kp, sm.pc[i]), sm.pd[i])), ciy = add(sm.py[i], mul(mul(kp, sm.ps[i]), sm.pd[i]));
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const float kq = (float)(q - 2);
          const float cjx = add(sm.px[j], mul(mul(kq, sm.pc[j]), sm.pd[j])), cjy = add(sm.py[j], mul(mul(kq, sm.ps[j]), sm.pd[j]));
          dmin = fminf(dmin, add(norm2(sub(cix, cjx), sub(ciy, cjy)), eps));
        }
      }
      const float col = fmaxf(sub(1.f, dmin / add(sm.pr[j], sm.pr[i])), 0.f);
      if (col > 0.f) {
        if (in.reduce_collision_with_max) atomicMax(&sm.col_bits[i], __float_as_uint(col));
        else atomicAdd(&sm.col_sum[i], fminf(col, 1.f));
      }
    }
  }
  __syncthreads();

  // ---- per-agent results of this step ----
  const size_t plane = (size_t)dm.n_scene * dm.n_mode * A * T;
  for (int a = tid; a < A; a += NT) {
    const size_t o = ((size_t)b * A + a) * T + (t - 1);
    bool collided = false;
    if ((en & 1) && sm.valid[a]) {
      for (int j = 0; j < A; ++j) {
        if (j == a || !sm.valid[j] || (sm.ped[a] && sm.ped[j])) continue;
        const float dx = sm.x[a] - sm.x[j], dy = sm.y[a] - sm.y[j], reach = sm.cull[a] + sm.cull[j];
        if (dx * dx + dy * dy > reach * reach) continue;  // far apart: separated (same test as in the pair loop)
        const bool sepd = ((sm.sep[a][j >> 5] >> (j & 31)) & 1u) || ((sm.sep[j][a >> 5] >> (a & 31)) & 1u);
        collided |= !sepd;
      }
    }
    const bool vv = sm.valid[a] && sm.veh[a];
    out.violations[1 * plane + o] = collided;
    out.violations[3 * plane + o] = (en & 2) ? (sm.f_edge[a] != 0) : 0;
    out.violations[5 * plane + o] = (en & 4) ? (sm.f_red[a] != 0) : 0;
    // raw passive predicate of this step; k_rule_sticky turns it into the > 20-step counter
    out.violations[7 * plane + o] = (en & 8) ? (vv && sm.f_near[a] && sm.spd[a] < 5.f && !sm.f_tl_ahead[a] && !sm.f_ag_ahead[a]) : 0;
    if (want_reward) {
      float col = in.reduce_collision_with_max ? __uint_as_float(sm.col_bits[a]) : sm.col_sum[a] / (float)sm.n_pvalid;
      if (!sm.pvalid[a]) col = 0.f;
      const float r0 = sub(0.f, mul(in.w_collision, col));
      // reference: (0 - w * col) - il_loss, masked by reward_valid; tb_rollout wrote (0 - il_loss) masked
      out.diffbar_rewards[o] = out.diffbar_rewards_valid[o] ? add(r0, out.diffbar_rewards[o]) : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_rule_sticky(long n_row, int T, int enable_mask, uint8_t* __restrict__ v) {
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_row) return;
  const size_t plane = (size_t)n_row * T;
  uint8_t* row = v + (size_t)r * T;
  for (int c = 0; c < 3; ++c) {  // collided, run_road_edge, run_red_light: sticky = running OR (traffic_rule_checker.py:426-448)
    bool st = false;
    for (int t = 0; t < T; ++t) {
      st |= row[(2 * c + 1) * plane + t] != 0;
      row[(2 * c) * plane + t] = st;
    }
  }
  // passive: counter = (counter + raw) * raw; this_step = counter > 20; sticky OR (:331-334,:466-470)
  bool st = false;
  float counter = 0.f;
  for (int t = 0; t < T; ++t) {
    const float raw = row[7 * plane + t] ? 1.f : 0.f;
    counter = (counter + raw) * raw;
    const bool now = (enable_mask & 8) && counter > 20.f;
    st |= now;
    row[7 * plane + t] = now;
    row[6 * plane + t] = st;
  }
}

}  // namespace rl
}  // namespace tb

using namespace tb;

static size_t rule_align(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t tb_rule_workspace_bytes(const TbDims* d) {
  if (check_dims_host(d) != TB_OK) return 0;
  const size_t n_node = (size_t)d->n_scene * d->n_pl * TB_PL_NODE;
  return rule_align(n_node * sizeof(float4)) + rule_align(n_node * sizeof(float2)) + rule_align((size_t)d->n_scene * 2 * sizeof(int32_t));
}

extern "C" int32_t tb_rule_checks(const TbDims* dims, const TbRuleIn* in, const TbRuleOut* out, void* workspace, void* stream) {
  const int rc = check_dims_host(dims);
  if (rc != TB_OK) return rc;
  if (!in || !out || !workspace || !out->violations) return TB_ERR_NULL;
  if (!in->preds || !in->valid || !in->override_masks || !in->outside_map_this_step || !in->gt_valid || !in->gt_pos || !in->gt_yaw ||
      !in->gt_spd || !in->agent_type || !in->agent_size)
    return TB_ERR_NULL;
  const int en = in->enable_mask;
  if (en & ~15) return TB_ERR_BAD_SHAPE;
  if ((en & 8) && !(en & 4)) return TB_ERR_UNSUPPORTED;  // reference: NameError (tl_step undefined)
  if ((en & (2 | 8)) && (!in->map_valid || !in->map_type || !in->map_pos || !in->map_dir)) return TB_ERR_NULL;
  if ((en & (4 | 8)) && (!in->tl_valid || !in->tl_pos || !in->tl_state || in->n_tl_frame < 1)) return TB_ERR_NULL;
  if (in->w_collision > 0.f && (!out->diffbar_rewards || !out->diffbar_rewards_valid)) return TB_ERR_NULL;
  if (dims->n_agent > rl::MAXA) return TB_ERR_BAD_SHAPE;
  if (!aligned16(in->preds) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return TB_ERR_ALIGN;
  const TbDims d = *dims;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = reinterpret_cast<char*>(workspace);
  const size_t n_node = (size_t)d.n_scene * d.n_pl * TB_PL_NODE;
  float4* edges = reinterpret_cast<float4*>(ws);
  float2* lanes = reinterpret_cast<float2*>(ws + rule_align(n_node * sizeof(float4)));
  int32_t* counts = reinterpret_cast<int32_t*>(ws + rule_align(n_node * sizeof(float4)) + rule_align(n_node * sizeof(float2)));
  if (en & (2 | 8)) {
    rl::k_rule_compact<<<d.n_scene, rl::NT, 0, st>>>(d.n_pl, in->map_valid, in->map_type, in->map_pos, in->map_dir, edges, lanes, counts);
    count_launch();
  }
  const int B = d.n_scene * d.n_mode;
  rl::k_rule_step<<<dim3(d.n_step, B), rl::NT, 0, st>>>(d, *in, *out, edges, lanes, counts);
  count_launch();
  const long n_row = (long)B * d.n_agent;
  rl::k_rule_sticky<<<(unsigned)((n_row + 127) / 128), 128, 0, st>>>(n_row, d.n_step, en, out->violations);
  count_launch();
  return launch_status();
}
