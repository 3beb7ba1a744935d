// Persistent tensor-core rollout (tcgen05): the WHOLE decode step of WaymoMotion.rollout (reference
// src/pl_modules/waymo_motion.py:108-354) for ONE scene-mode per CTA, looping over the decode steps t_first..t_last
// without returning to the host: state embedding -> 3 agent->map layers -> 3 agent->traffic-light layers -> 3
// agent<->agent layers -> 3 GRU layers -> add_goal -> add_latent -> action head -> dynamics / override / rule checks /
// kill / reward -> outputs in the final [B,A,T,.] layout.  Supports n_agent <= 64 (larger scenes use the two-kernel path).
//
// Roles (320 threads):
//   warps 0..7  row workers.  TMEM lane l = 32 * (warp % 4) + lane; agent a = l % 64; `upper` = l / 64; `half` = warp / 4.
//               Every Linear is evaluated for 128 rows = the 64 agents TWICE (rows a and a + 64 hold the same values), which
//               costs nothing (M = 128 is the tensor-core tile height) and lets the attention run "head-stacked":
//   warp  8     issuer: one lane issues every tcgen05.mma and commits completion to mbarriers.
//   warp  9     loader: one lane streams the 64 KB operand blocks (weights, compacted K|V key blocks) into a 2-slot ring
//               with bulk-async copies, 32 KB halves with their own full / free barriers.
//
// Head-stacked attention (n_agent <= 64): for head pair hp = {2hp, 2hp+1} the A operand row l holds the query of agent a
// for head 2hp + upper in that head's 32 dims of the 64-dim K-block and zeros in the other head's dims, so ONE N = 64 MMA
// chain against the K-block [64 keys x 64 dims] yields S[l, key] = q_{2hp+upper}(a) . k_{2hp+upper}(key): all 128 lanes
// carry useful logits.  PV uses the same trick with N = 64 dims: D[l, 0:32] = P V_{2hp}, D[l, 32:64] = P V_{2hp+1}; lane l
// keeps columns 32 * upper .. +32.  Worker group `half` = hp owns pass hp, so the two groups alternate on the tensor pipe
// (one does its softmax while the other's QK^T / PV run).
//
// TMEM (512 columns):  [0,128) ACC0 (S buffers of pass 0 / 1 at [0,64) / [64,128) during attention)   [128,256) ACC1
//                      [256,384) ACC2 | second A operand (GRU hidden, goal / latent feature) | O accumulators of the 2 passes
//                      [384,512) A operand (bf16x2 packed: hi [384,448), lo [448,512))
// Numerics: every contraction is bf16x3 (hi*hi + lo*hi + hi*lo, fp32 accumulate), see tb_tc.cuh.
#include "tb_host.h"

namespace tb {
namespace pr {

constexpr int WORKERS = 256;
constexpr int THREADS = WORKERS + 64;
constexpr int MAXA = 64;
constexpr uint32_t BLK = 65536, HALF = 32768;
constexpr uint32_t T_ACC0 = 0, T_ACC1 = 128, T_ACC2 = 256, T_A2 = 256, T_O = 256, T_A = 384;
// parameter-vector sets ("phases"): 0-8 attention layers, 9-11 GRU layers, 12 add_goal, 13 add_latent, 14 head

struct Args {
  TbDims dm;
  TbRolloutIn in;
  const float* packed;
  const unsigned char* tcw;
  StateView sv;
  TbRolloutOut out;
  int t_first, t_last;
  long long* trace;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid / tanh through MUFU.EX2 + MUFU.RCP: ~2^-21 relative error, far inside the bf16x3 error of their arguments
__device__ __forceinline__ float fast_sigmoid(float x) {
  return __fdividef(1.0f, 1.0f + ex2_approx(-1.4426950408889634f * x));
}
__device__ __forceinline__ float fast_tanh(float x) { return fmaf(2.0f, fast_sigmoid(2.0f * x), -1.0f); }
using tc::elect_one;
using tc::uniform;

// Issuer / loader primitives are inlined into the step program: measured 3 % faster than out-of-line copies (A/B on one box,
// -DTB_NOINLINE_ISSUER), although the latter shrink the kernel by ~5 k instructions.
#ifdef TB_NOINLINE_ISSUER
#define TB_ROLE_FN __noinline__
#else
#define TB_ROLE_FN __forceinline__
#endif

// ---- thread-block cluster helpers (split of the agent->map attention over the CTAs of a cluster) -----------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_relaxed() {  // rendezvous only: the caller publishes no data
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t cta_smem_addr, uint32_t rank) {  // same offset in CTA `rank`'s shared memory
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32x4(uint32_t cluster_addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_f32x2(uint32_t cluster_addr, float2 v) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_cluster_f32(uint32_t cluster_addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_cluster_f32x4(uint32_t cluster_addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr) : "memory");
  return v;
}
__device__ __forceinline__ float2 ld_cluster_f32x2(uint32_t cluster_addr) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(cluster_addr) : "memory");
  return v;
}
// asynchronous remote stores: the data lands in CTA `rank`'s shared memory and completes `bytes` on THAT CTA's mbarrier, so
// the receiver waits on a local barrier for exactly its data instead of the whole cluster fencing (barrier.cluster release /
// acquire = MEMBAR.ALL.GPU on every thread)
__device__ __forceinline__ void st_async_f32x4(uint32_t cluster_addr, float4 v, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(cluster_addr),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(cluster_bar)
               : "memory");
}
__device__ __forceinline__ void st_async_f32x2(uint32_t cluster_addr, float2 v, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(cluster_addr), "f"(v.x),
               "f"(v.y), "r"(cluster_bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {  // acquire at cluster scope (remote st.async data)
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(tc::smem_u32(bar)), "r"(parity), "r"(1000000u)
        : "memory");
    if (ok) return;
    tc::watchdog_check(t0);
  }
}
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float smooth_l1(float d) {
  const float a = fabsf(d);
  return a < 1.0f ? 0.5f * d * d : a - 0.5f;
}

// offset (floats, in the packed blob) of parameter vector i of phase p, or -1
__device__ __forceinline__ int phase_vec(int p, int i) {
  if (p < 9) {
    const int kind = p / 3, L = p % 3;
    const int base = (kind == 0   ? tbw::model_transformer_as2pl_layers_0_norm1_weight
                      : kind == 1 ? tbw::model_transformer_as2tl_layers_0_norm1_weight
                                  : tbw::model_agent_interaction_transformer_layers_0_norm1_weight) +
                     L * tfl::STRIDE;
    switch (i) {
      case 0: return base + tfl::NORM1_W;
      case 1: return base + tfl::NORM1_B;
      case 2: return base + tfl::IN_B;
      case 3: return base + tfl::OUT_B;
      case 4: return base + tfl::NORM2_W;
      case 5: return base + tfl::NORM2_B;
      case 6: return base + tfl::L1_B;
      case 7: return base + tfl::L2_B;
      case 8: return kind == 2 ? base + tfl::NORMT_W : -1;
      case 9: return kind == 2 ? base + tfl::NORMT_B : -1;
      case 10: return kind == 2 ? base + tfl::IN_B + 128 : -1;
      case 11: return kind == 2 ? base + tfl::IN_B + 256 : -1;
    }
    return -1;
  }
  if (p < 12) {
    const int base = gru::BASE + (p - 9) * gru::STRIDE;
    if (i < 3) return base + gru::B_IH + 128 * i;
    if (i < 6) return base + gru::B_HH + 128 * (i - 3);
    return -1;
  }
  if (p == 12) return i == 0 ? tbw::model_add_goal_mlp_out_fc_layers_0_bias : i == 1 ? tbw::model_add_goal_mlp_out_fc_layers_3_bias : -1;
  if (p == 13) return i == 0 ? tbw::model_add_latent_mlp_out_fc_layers_0_bias : i == 1 ? tbw::model_add_latent_mlp_out_fc_layers_3_bias : -1;
  switch (i) {  // head
    case 0: return tbw::action_head_mlp_mean_0_fc_layers_0_bias;
    case 1: return tbw::action_head_mlp_mean_1_fc_layers_0_bias;
    case 2: return tbw::action_head_mlp_mean_2_fc_layers_0_bias;
    case 3: return tbw::action_head_mlp_mean_0_fc_layers_2_weight;  // Wt4[32][2][4] = 256 floats
    case 4: return tbw::action_head_mlp_mean_0_fc_layers_2_weight + 128;
    case 5: return tbw::action_head_mlp_mean_1_fc_layers_2_weight;
    case 6: return tbw::action_head_mlp_mean_1_fc_layers_2_weight + 128;
    case 7: return tbw::action_head_mlp_mean_2_fc_layers_2_weight;
    case 8: return tbw::action_head_mlp_mean_2_fc_layers_2_weight + 128;
  }
  return -1;
}

// ---- the operand-block sequence of one decode step, shared by the loader and the issuer ---------------------------------
struct StepCfg {
  int nblk_map, nblk_tl;  // nblk_map: key blocks of the scene; this CTA takes blocks rank, rank + n_cta, ...
  int rank, n_cta;
  const unsigned char* kv_map;  // + (L * S) * nT_map * BLK per layer
  const unsigned char* kv_tl;
  size_t kv_map_layer_stride, kv_tl_layer_stride;
  // 64 < n_agent <= 128 ("halves"): the scene-mode is shared by the 2 CTAs of a cluster, 64 agents each; the interaction key blocks
  // of both are exchanged through global memory (`kvx`: [layer][scene-mode][cluster rank] x 64 KB) and streamed like map keys
  bool halves;
  int crank, t;
  const unsigned char* kvx;  // + ((L * B + b) * 2) * BLK: block of rank 0, then of rank 1
  size_t kvx_layer_stride;
  int* xflag;                // this scene-mode's key-block counters [2] (StateView::xch + 6)
};

template <class R>
__device__ __forceinline__ void enumerate_step(const Args& a, const StepCfg& c, R& r) {
  auto W = [&](int first, int idx) { return a.tcw + (size_t)(first + idx) * BLK; };
  bool bypass = false;
  int nblk_int = 0;
#pragma unroll 1
  for (int Lx = 0; Lx < 9; ++Lx) {
    const int kind = Lx / 3, L = Lx % 3;
    if (Lx == 6) {
      const int nv = r.wait_cfg();
      bypass = nv == 1;
      nblk_int = nv > 0 ? 1 : 0;
    }
    if (kind == 2 && bypass) break;
    const int w0 = (kind == 0   ? tbb::model_transformer_as2pl_layers_0_attn_in_proj_weight
                    : kind == 1 ? tbb::model_transformer_as2tl_layers_0_attn_in_proj_weight
                                : tbb::model_agent_interaction_transformer_layers_0_attn_in_proj_weight) +
                   6 * L;
    const int nblk = kind == 0 ? c.nblk_map : kind == 1 ? c.nblk_tl : nblk_int;
    // blocks of this CTA: the agent->map attention is split over the cluster (online-softmax partials merged by the workers)
    const int nblk_my = kind == 0 ? (c.nblk_map > c.rank ? (c.nblk_map - c.rank + c.n_cta - 1) / c.n_cta : 0) : nblk;
    if (nblk > 0) {
      if (kind == 2) {
        r.gemm_begin();
        r.chain(W(w0, 1), T_ACC0, T_A, false);
        r.chain(W(w0, 2), T_ACC1, T_A, false);
        r.gemm_end();
        if (c.halves) r.kvx(c, L);
        else r.kvi();
      }
      r.gemm_begin();
      r.chain(W(w0, 0), kind == 2 ? R::kInteractionQ : T_ACC0, T_A, false);
      r.gemm_end();
      if (kind == 2 && c.halves)
        r.att(false, 2, c.kvx + L * c.kvx_layer_stride, (size_t)BLK);
      else
        r.att(kind == 2, nblk_my, kind == 0 ? c.kv_map + L * c.kv_map_layer_stride + (size_t)c.rank * BLK : c.kv_tl + L * c.kv_tl_layer_stride,
              kind == 0 ? (size_t)c.n_cta * BLK : (size_t)BLK);
      const bool split = kind == 0 && c.n_cta > 1;
      if (split) r.csync_before_wo();
      r.gemm_begin();
      r.chain(W(w0, 3), T_ACC0, T_A, false);
      r.gemm_end();
      if (split) r.csync_after_wo();
    }
    r.gemm_begin();
    r.chain(W(w0, 4), T_ACC0, T_A, false);
    r.gemm_end();
    r.gemm_begin();
    r.chain(W(w0, 5), T_ACC0, T_A, false);
    r.gemm_end();
  }
#pragma unroll 1
  for (int L = 0; L < 3; ++L) {  // GRU: gate blocks r, z, n of weight_ih / weight_hh
    const int wi = tbb::model_agent_temporal_rnn_weight_ih_l0 + 6 * L, wh = tbb::model_agent_temporal_rnn_weight_hh_l0 + 6 * L;
    r.gemm_begin();
    r.chain(W(wi, 0), T_ACC0, T_A, false);
    r.chain(W(wh, 0), T_ACC0, T_A2, true);
    r.chain(W(wh, 2), T_ACC1, T_A2, false);
    r.gemm_end();
    r.gemm_begin();
    r.chain(W(wi, 1), T_ACC0, T_A, false);
    r.chain(W(wh, 1), T_ACC0, T_A2, true);
    r.chain(W(wi, 2), T_ACC1, T_A, false);
    r.gemm_end();
  }
  {
    const int w[2][2] = {{tbb::model_add_goal_mlp_out_fc_layers_0_weight, tbb::model_add_goal_mlp_out_fc_layers_3_weight},
                         {tbb::model_add_latent_mlp_out_fc_layers_0_weight, tbb::model_add_latent_mlp_out_fc_layers_3_weight}};
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {
      r.gemm_begin();
      r.chain(W(w[j][0], 0), T_ACC0, T_A, false);
      if (!R::kAddHalfHoisted) r.chain(W(w[j][0], 1), T_ACC0, T_A2, true);
      r.gemm_end();
      r.gemm_begin();
      r.chain(W(w[j][1], 0), T_ACC0, T_A, false);
      r.gemm_end();
    }
  }
  r.gemm_begin();
  r.chain(W(tbb::action_head_mlp_mean_0_fc_layers_0_weight, 0), T_ACC0, T_A, false);
  r.chain(W(tbb::action_head_mlp_mean_1_fc_layers_0_weight, 0), T_ACC1, T_A, false);
  r.chain(W(tbb::action_head_mlp_mean_2_fc_layers_0_weight, 0), T_ACC2, T_A, false);
  r.gemm_end();
}

template <class SM>
struct LoaderT {  // run by a whole (converged) warp; one elected lane issues the copies
  static constexpr uint32_t kInteractionQ = SM::kInteractionQ;
  static constexpr bool kAddHalfHoisted = SM::kAddHalfHoisted;
  SM& sm;
  uint32_t g = 0, n_cfg = 0;
  __device__ LoaderT(SM& s) : sm(s) {}
  __device__ __forceinline__ void wait_free(uint32_t slot, uint32_t half) {
    const uint32_t use = g >> 1;
    if (use > 0) tc::mbar_wait(&sm.free_[slot][half], (use - 1) & 1);
  }
  __device__ TB_ROLE_FN void load(const unsigned char* ptr) {
    const uint32_t slot = g & 1;
#pragma unroll
    for (uint32_t h = 0; h < 2; ++h) {
      wait_free(slot, h);
      if (elect_one()) {
        tc::mbar_expect_tx(&sm.full[slot][h], HALF);
        tc::bulk_g2s(sm.ring[slot] + h * HALF, ptr + h * HALF, HALF, &sm.full[slot][h]);
      }
      __syncwarp();
    }
    ++g;
  }
  __device__ __forceinline__ int wait_cfg() {
    tc::mbar_wait(&sm.cfg, n_cfg & 1);
    ++n_cfg;
    return uniform(*reinterpret_cast<volatile int*>(&sm.n_valid));
  }
  __device__ __forceinline__ void gemm_begin() {}
  __device__ __forceinline__ void gemm_end() {}
  __device__ __forceinline__ void chain(const unsigned char* w, uint32_t, uint32_t, bool) { load(w); }
  __device__ __forceinline__ void kvi() {
    const uint32_t slot = g & 1;
    wait_free(slot, 0);
    wait_free(slot, 1);
    if (elect_one()) {
      *reinterpret_cast<volatile int*>(&sm.kvi_slot) = (int)slot;
      __threadfence_block();
      mbar_arrive(&sm.grant);
    }
    __syncwarp();
    ++g;
  }
  // halves mode: this CTA's key block of layer L is complete in global memory once all 16 worker warps arrived at `wfill`;
  // publish it (release), wait for the peer's (acquire), then the two blocks are loaded like any other key block
  uint32_t n_kvx = 0;
  __device__ __forceinline__ void kvx(const StepCfg& c, int L) {
    tc::mbar_wait(&sm.wfill, n_kvx & 1);
    ++n_kvx;
    if (elect_one()) {
      const int want = (c.t - 1) * 3 + L + 1;
      asm volatile("fence.proxy.async;" ::: "memory");
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(c.xflag + c.crank), "r"(want) : "memory");
      const long long t0 = clock64();
      int got;
      do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(got) : "l"(c.xflag + (c.crank ^ 1)) : "memory");
        if (got < want) {
          __nanosleep(64);
          tc::watchdog_check(t0);
        }
      } while (got < want);
      asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncwarp();
  }
  __device__ __forceinline__ void att(bool kvi_keys, int nblk, const unsigned char* blocks, size_t stride) {
    if (kvi_keys) return;
    for (int j = 0; j < nblk; ++j) load(blocks + (size_t)j * stride);
  }
  // the two cluster barriers of a split layer's merge: the loader joins them after the Wo block is on its way
  __device__ __forceinline__ void csync_before_wo() {}
  __device__ __forceinline__ void csync_after_wo() {
#pragma unroll
    for (int i = 0; i < SM::kMergeBarriers; ++i) cluster_sync_relaxed();
  }
};

template <class SM>
struct IssuerT {
  static constexpr uint32_t kInteractionQ = SM::kInteractionQ;
  static constexpr bool kAddHalfHoisted = SM::kAddHalfHoisted;
  SM& sm;
  uint32_t tm0;
  uint32_t g = 0, nf[2] = {0, 0}, n_ready = 0, n_cfg = 0, n_p[2] = {0, 0}, n_kvi = 0, kvi_slot = 0;
  __device__ IssuerT(SM& s, uint32_t t) : sm(s), tm0(t) {}
  __device__ __forceinline__ int wait_cfg() {
    tc::mbar_wait(&sm.cfg, n_cfg & 1);
    ++n_cfg;
    return uniform(*reinterpret_cast<volatile int*>(&sm.n_valid));
  }
  __device__ __forceinline__ void wait_ready() {
    tc::mbar_wait(&sm.ready, n_ready & 1);
    ++n_ready;
    tc::tc_fence_after();
  }
  bool need_ready = false;
  // the wait for the workers' operand is deferred into the first chain(): weight-slot wait and descriptor set-up come first
  __device__ __forceinline__ void gemm_begin() { need_ready = true; }
  __device__ __forceinline__ void gemm_end() {
    if (elect_one()) tc::mma_commit(&sm.mma);
    __syncwarp();
  }
  __device__ TB_ROLE_FN void chain(const unsigned char*, uint32_t dcol, uint32_t acol, bool accum) {
    const uint32_t slot = g & 1;
    tc::mbar_wait(&sm.full[slot][0], nf[slot] & 1);
    tc::mbar_wait(&sm.full[slot][1], nf[slot] & 1);
    ++nf[slot];
    tc::tc_fence_after();
    const uint32_t wh = tc::smem_u32(sm.ring[slot]);
    const uint64_t dh = tc::make_desc_sw128(wh), dl = tc::make_desc_sw128(wh + 2 * tc::KB_BYTES_128);
    const uint32_t idesc = tc::make_idesc_bf16(128, 128);
    if (need_ready) {
      wait_ready();
      need_ready = false;
    }
    if (elect_one()) {
#pragma unroll
      for (int term = 0; term < 3; ++term) {
        const uint32_t ta = tm0 + acol + (term == 1 ? 64 : 0);
        const uint64_t db = term == 2 ? dl : dh;
#pragma unroll
        for (int k = 0; k < 128; k += 16)
          tc::mma_bf16_ts(tm0 + dcol, ta + k / 2, db + (uint64_t)(((k >> 6) * tc::KB_BYTES_128 + (k & 63) * 2) >> 4), idesc,
                          (accum || term > 0 || k > 0) ? 1u : 0u);
      }
      tc::mma_commit(&sm.free_[slot][0]);
      tc::mma_commit(&sm.free_[slot][1]);
    }
    __syncwarp();
    ++g;
  }
  __device__ __forceinline__ void kvi() {
    kvi_slot = g & 1;
    ++g;
  }
  __device__ __forceinline__ void kvx(const StepCfg&, int) {}  // the exchanged key blocks arrive through the ring like map keys
  __device__ __forceinline__ void csync_before_wo() {  // the issuer joins the merge's four cluster barriers right away
#pragma unroll
    for (int i = 0; i < SM::kMergeBarriers; ++i) cluster_sync_relaxed();
  }
  __device__ __forceinline__ void csync_after_wo() {}
  // QK^T of pass hp against the K half of the block in `slot`:  S_hp[128 x 64 keys] = A_hp[128 x 64 dims] K_hp^T
  __device__ TB_ROLE_FN void issue_qk(uint32_t slot, int hp) {
    const uint32_t kb = tc::smem_u32(sm.ring[slot]) + hp * 8192;
    const uint64_t dh = tc::make_desc_sw128(kb), dl = tc::make_desc_sw128(kb + 16384);
    const uint32_t idesc = tc::make_idesc_bf16(128, 64);
#pragma unroll
    for (int term = 0; term < 3; ++term) {
      const uint32_t ta = tm0 + T_A + (term == 1 ? 64 : 0) + 32 * hp;
      const uint64_t db = term == 2 ? dl : dh;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        tc::mma_bf16_ts(tm0 + T_ACC0 + 64 * hp, ta + 8 * ks, db + (uint64_t)(2 * ks), idesc, (term > 0 || ks > 0) ? 1u : 0u);
    }
    tc::mma_commit(&sm.s[hp]);
  }
  // PV of pass hp:  O_hp[128 x 64 dims] (+)= P_hp[128 x 64 keys] V_hp   (P = bf16 hi | lo packed over the S columns)
  __device__ TB_ROLE_FN void issue_pv(uint32_t slot, int hp, bool accum) {
    const uint32_t vb = tc::smem_u32(sm.ring[slot]) + HALF + hp * 8192;
    const uint64_t dh = tc::make_desc_sw128(vb), dl = tc::make_desc_sw128(vb + 16384);
    const uint32_t idesc = tc::make_idesc_bf16(128, 64);
    const uint32_t sp = tm0 + T_ACC0 + 64 * hp;
#pragma unroll
    for (int term = 0; term < 3; ++term) {
      const uint32_t ta = sp + (term == 1 ? 32 : 0);
      const uint64_t db = term == 2 ? dl : dh;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        tc::mma_bf16_ts(tm0 + T_O + 64 * hp, ta + 8 * ks, db + (uint64_t)(2 * ks), idesc, (accum || term > 0 || ks > 0) ? 1u : 0u);
    }
  }
  __device__ __forceinline__ void att(bool kvi_keys, int nblk, const unsigned char*, size_t) {
    if (nblk == 0) {  // (a cluster rank without key blocks of its own)
      wait_ready();
      return;
    }
    const uint32_t g0 = kvi_keys ? kvi_slot : g;  // slot parity of block jb = (g0 + jb) & 1
    if (kvi_keys) {
      tc::mbar_wait(&sm.wfill, n_kvi & 1);
      ++n_kvi;
    } else {
      tc::mbar_wait(&sm.full[g0 & 1][0], nf[g0 & 1] & 1);
    }
    wait_ready();  // stacked queries written (the key block is normally there long before)
    if (elect_one()) {
      issue_qk(g0 & 1, 0);
      issue_qk(g0 & 1, 1);
      tc::mma_commit(&sm.free_[g0 & 1][0]);
    }
    __syncwarp();
    for (int jb = 0; jb < nblk; ++jb) {
      const uint32_t slot = (g0 + jb) & 1, slotn = slot ^ 1;
#pragma unroll
      for (int hp = 0; hp < 2; ++hp) {
        // operand blocks first (prefetched: these waits normally fall through), then the probabilities of this pass
        if (hp == 0 && !kvi_keys) {
          tc::mbar_wait(&sm.full[slot][1], nf[slot] & 1);
          ++nf[slot];
        }
        const bool more = jb + 1 < nblk;
        if (more && hp == 0) tc::mbar_wait(&sm.full[slotn][0], nf[slotn] & 1);
        tc::mbar_wait(&sm.p[hp], n_p[hp] & 1);
        ++n_p[hp];
        tc::tc_fence_after();
        if (elect_one()) {
          issue_pv(slot, hp, jb > 0);
          if (hp == 1) tc::mma_commit(&sm.free_[slot][1]);
          if (more) {
            issue_qk(slotn, hp);
            if (hp == 1) tc::mma_commit(&sm.free_[slotn][0]);
          } else {
            tc::mma_commit(&sm.o[hp]);
          }
        }
        __syncwarp();
      }
    }
    if (!kvi_keys) g += nblk;
  }
};

// =============================================================================================================================
// 16-worker-warp variant: FOUR threads per TMEM lane.  Thread (lane l, part 0..3) owns columns [32 part, 32 part + 32) of
// lane l in every operand write (both lanes of an agent write complete A operands), and 16 columns [32 part + 16 upper, +16)
// of the agent in the epilogues that end in shared / global memory.  In the attention the pass of a lane is shared by the
// two threads part = 2 hp, 2 hp + 1: each takes 32 of the block's 64 keys, and they agree on the running maximum through
// shared memory (one 64-thread named barrier per key block).  The first version of this kernel (8 worker
// warps, two threads per lane: 25.5 vs 21.7 ms per rollout, profiles/r1j_*) was removed in round 2.
// =============================================================================================================================
constexpr int WORKERS16 = 512;
constexpr int THREADS16 = WORKERS16 + 64;

struct Smem16 {
  unsigned char ring[2][BLK];
  float xs[128 * MAXA];  // residual stream, [col][agent]
  float xo[128 * MAXA];  // exchange buffer
  float lp[2][12][128];  // parameter vectors of the current / next phase
  float emb_w1[384], emb_b1[32], emb_b2[32], f_xy[24], f_yaw[48];
  float2 red[2][4][128];  // LayerNorm partials {sum, M2} [buffer][part][lane]
  float mxs[4][128];      // softmax exchange between the two threads of a (lane, pass): block max / final sum
  float mean_part[8][MAXA][2];
  float4 pose[MAXA];
  float2 vel[MAXA];
  float acc[MAXA], yaw_rate[MAXA];
  uint8_t valid[MAXA], killed[MAXA], goal_valid[MAXA], sticky[3][MAXA], type[MAXA][4];
  float tailc[9][MAXA];
  float map_boundary[4];
  uint8_t tflag[MAXA];
  uint64_t full[2][2], free_[2][2], grant, wfill, ready, mma, cfg, s[2], p[2], o[2];
  uint64_t rs_bar, ag_bar;  // merge of the cluster partials: bytes of the reduce-scatter / all-gather stage landed here
  static constexpr int kMergeBarriers = 2;
  static constexpr uint32_t kInteractionQ = T_ACC2;  // Q of the interaction layers lands beside K | V (their epilogue overlaps it)
  static constexpr bool kAddHalfHoisted = true;      // ... is precomputed per rollout (k_rollout_init) and added in the epilogue
  uint32_t tmem_base;
  int n_valid, kvi_slot;
  unsigned peer_vm[2];  // halves mode: valid mask (lo, hi) of the peer CTA's 64 agents at this step
};
static_assert(sizeof(Smem16) + 1024 <= 232448, "shared memory budget (227 KB per CTA)");

__device__ __forceinline__ void worker_sync16() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__global__ void __launch_bounds__(THREADS16, 1) k_rollout_tc16(Args a) {
  extern __shared__ unsigned char smem_raw[];
  Smem16& sm = *reinterpret_cast<Smem16*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
  const TbDims& dm = a.dm;
  const TbRolloutIn& in = a.in;
  const float* __restrict__ packed = a.packed;
  const int A = dm.n_agent, K = dm.n_mode, S = dm.n_scene, B = S * K, Th = dm.n_step_hist, T = dm.n_step, Tg = dm.n_step_gt;
  const int nT_map = (dm.n_pl + 63) / 64, nT_tl = (dm.n_tl + 63) / 64;
  // cluster: either the split of the agent->map attention over 1 / 2 / 4 CTAs (n_agent <= 64), or -- "halves", 64 < n_agent <= 128
  // -- two CTAs that each own 64 agents of the scene-mode (rows 0..63 of every per-agent array belong to cluster rank 0)
  const int crank = (int)cluster_ctarank(), n_cl = (int)cluster_nctarank();
  const bool halves = A > MAXA;
  const int rank = halves ? 0 : crank, n_cta = halves ? 1 : n_cl;  // rank / size of the attention split
  const int a0 = halves ? crank * MAXA : 0;                        // first agent of this CTA
  const int A_loc = min(MAXA, A - a0);
  const int b = blockIdx.x / n_cl, s = b / K;
  const int tid = threadIdx.x, warp = uniform(tid >> 5), lane = tid & 31;
  const size_t BA = (size_t)B * A;
  const int nkey_map = uniform(in.n_key_map[s]);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j) {
        tc::mbar_init(&sm.full[i][j], 1);
        tc::mbar_init(&sm.free_[i][j], 1);
      }
    tc::mbar_init(&sm.grant, 1);
    tc::mbar_init(&sm.wfill, 16);
    tc::mbar_init(&sm.ready, 16);
    tc::mbar_init(&sm.rs_bar, 1);
    tc::mbar_init(&sm.ag_bar, 1);
    tc::mbar_init(&sm.mma, 1);
    tc::mbar_init(&sm.cfg, 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&sm.s[i], 1);
      tc::mbar_init(&sm.p[i], 8);
      tc::mbar_init(&sm.o[i], 1);
    }
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&sm.tmem_base, 512);
  for (int i = tid; i < 384; i += THREADS16) sm.emb_w1[i] = __ldg(packed + tbw::model_agent_encoder_mlp_fc_layers_0_weight + i);
  if (tid < 32) {
    sm.emb_b1[tid] = __ldg(packed + tbw::model_agent_encoder_mlp_fc_layers_0_bias + tid);
    sm.emb_b2[tid] = __ldg(packed + tbw::model_agent_encoder_mlp_fc_layers_3_bias + tid);
  }
  if (tid < 24) sm.f_xy[tid] = __ldg(packed + tbw::pre_processing_input_pose_pe_agent_pe_xy_freqs + tid);
  if (tid < 48) sm.f_yaw[tid] = __ldg(packed + tbw::pre_processing_input_pose_pe_agent_pe_yaw_freqs + tid);
  if (tid < MAXA) {
    const int ag = tid;
    const bool live = a0 + ag < A;
    const size_t ba = (size_t)b * A + (live ? a0 + ag : 0), sa = (size_t)s * A + (live ? a0 + ag : 0);
    sm.pose[ag] = live ? *reinterpret_cast<const float4*>(a.sv.agent_state + ba * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    sm.vel[ag] = live ? make_float2(a.sv.vel[ba * 2], a.sv.vel[ba * 2 + 1]) : make_float2(0.f, 0.f);
    sm.acc[ag] = live ? a.sv.acc[ba] : 0.f;
    sm.yaw_rate[ag] = live ? a.sv.yaw_rate[ba] : 0.f;
    sm.valid[ag] = live ? a.sv.valid[(size_t)(a.t_first & 1) * BA + ba] : (uint8_t)0;
    sm.killed[ag] = live ? a.sv.killed[ba] : (uint8_t)0;
    sm.goal_valid[ag] = live ? a.sv.goal_valid[ba] : (uint8_t)0;
    for (int i = 0; i < 3; ++i) {
      sm.sticky[i][ag] = live ? a.sv.sticky[(size_t)i * BA + ba] : (uint8_t)0;
      sm.type[ag][i] = live ? in.agent_type[sa * 3 + i] : (uint8_t)0;
    }
  }
  float4* const hid_t = a.sv.hidden_t + ((size_t)rank * 3 * B + b) * 32 * A;
  float4* const x0_t = a.sv.x0_t + ((size_t)rank * B + b) * 32 * A;
  const float4* const goal_c_t = a.sv.goal_c_t + (size_t)b * 32 * A;
  const float4* const latent_c_t = a.sv.latent_c_t + (size_t)b * 32 * A;
  for (int L = 0; L < 3; ++L) {
    const float4* src = reinterpret_cast<const float4*>(a.sv.hidden + ((size_t)L * BA + (size_t)b * A) * D);
    float4* dst = hid_t + (size_t)L * B * 32 * A;
    for (int i = tid; i < A_loc * 32; i += THREADS16) {
      const int ag_ = a0 + i % A_loc, c4 = i / A_loc;
      dst[c4 * A + ag_] = src[ag_ * 32 + c4];
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  cluster_sync_all();
  const uint32_t tm0 = (uint32_t)uniform((int)sm.tmem_base);

  StepCfg cfg;
  cfg.nblk_map = (nkey_map + 63) / 64;
  cfg.rank = rank;
  cfg.n_cta = n_cta;
  cfg.kv_map = in.kv_map_tc + (size_t)s * nT_map * BLK;
  cfg.kv_map_layer_stride = (size_t)S * nT_map * BLK;
  cfg.kv_tl_layer_stride = (size_t)S * Th * nT_tl * BLK;
  cfg.halves = halves;
  cfg.crank = crank;
  cfg.t = 0;
  cfg.kvx = reinterpret_cast<const unsigned char*>(a.sv.kv_int) + (size_t)b * 2 * BLK;
  cfg.kvx_layer_stride = (size_t)B * 2 * BLK;
  cfg.xflag = a.sv.xch + (size_t)b * 16 + 6;

  if (warp == 17) {
    LoaderT<Smem16> ld(sm);
    for (int t = a.t_first; t <= a.t_last; ++t) {
      const int tl_t = min(t - 1, Th - 1);
      cfg.nblk_tl = (uniform(in.n_key_tl[(size_t)s * Th + tl_t]) + 63) / 64;
      cfg.kv_tl = in.kv_tl_tc + ((size_t)s * Th + tl_t) * nT_tl * BLK;
      cfg.t = t;
      enumerate_step(a, cfg, ld);
    }
  } else if (warp == 16) {
    IssuerT<Smem16> is(sm, tm0);
    for (int t = a.t_first; t <= a.t_last; ++t) {
      const int tl_t = min(t - 1, Th - 1);
      cfg.nblk_tl = (uniform(in.n_key_tl[(size_t)s * Th + tl_t]) + 63) / 64;
      cfg.kv_tl = nullptr;
      cfg.t = t;
      enumerate_step(a, cfg, is);
    }
  } else {
    // ========================================================================================================== workers
    const int quad = warp & 3, part = warp >> 2, half = part >> 1, sub = part & 1;
    const int l = quad * 32 + lane;  // TMEM lane
    const int ag = l & 63, upper = l >> 6;
    const int cq = 32 * part;             // operand columns of this thread
    const int ce = cq + 16 * upper;       // epilogue columns of this thread (16)
    const bool live = a0 + ag < A;
    const int agg = live ? a0 + ag : 0;  // agent index inside the scene (global arrays, [..][A] scratch layouts)
    const bool writer = upper == 0 && live && rank == 0;
    const size_t ba = (size_t)b * A + agg, sa = (size_t)s * A + agg;
    const uint32_t tm = tm0 + ((uint32_t)(quad * 32) << 16);
    const int pair_id = 2 + quad * 2 + half;  // named barrier of the two warps (quad, 2 half) and (quad, 2 half + 1)
    uint32_t n_mma = 0, n_s = 0, n_o = 0, n_grant = 0, n_ln = 0, n_lp = 0, n_merge = 0;
    int n_mark = 0;
    auto mark = [&]() {
      if (a.trace && blockIdx.x == 0 && tid == 0 && n_mark < 1000) a.trace[n_mark++] = clock64();
    };
#ifdef TB_TRACE_DETAIL
    int n_dmark = 0, t_cur = 0;
    auto dmark = [&](int id) {
      if (a.trace && blockIdx.x == 0 && (tid == 0 || tid == 256) && t_cur == a.t_first + 3 && n_dmark < 700) {
        long long* dst = a.trace + 1024 + (tid == 256 ? 1400 : 0) + 2 * n_dmark++;
        dst[0] = id;
        dst[1] = clock64();
      }
    };
#else
    auto dmark = [](int) {};
#endif
    auto signal_ready = [&]() {
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.ready);
    };
    auto wait_gemm = [&]() {
      tc::mbar_wait(&sm.mma, n_mma & 1);
      tc::tc_fence_after();
      ++n_mma;
    };
    auto xs_at = [&](int c) -> float& { return sm.xs[c * MAXA + ag]; };
    auto load_x = [&](float (&v)[32]) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = xs_at(cq + i);
    };
    auto write_A = [&](uint32_t col, const float (&v)[32]) {  // columns cq .. cq+31 of a K = 128 operand
      float ph[16], pl[16];
      tc::split32_packed(v, ph, pl);
      tc::tmem_st16(tm + col + cq / 2, ph);
      tc::tmem_st16(tm + col + 64 + cq / 2, pl);
    };
    auto load_acc = [&](uint32_t col, float (&v)[32]) {
      tc::tmem_ld32(tm + col + cq, v);
      tc::tmem_ld_wait();
    };
    auto load_acc16 = [&](uint32_t col, float (&v)[16]) {  // this thread's 16 epilogue columns
      tc::tmem_ld16(tm + col + ce, v);
      tc::tmem_ld_wait();
    };
    auto ln32 = [&](float (&v)[32], const float* g, const float* bt) {
      // packed pairs: lanes (0, 1) and (2, 3) of the former four-way accumulators, same order of operations per lane
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        tc::add2(s4[0], s4[1], s4[0], s4[1], v[i], v[i + 1]);
        tc::add2(s4[2], s4[3], s4[2], s4[3], v[i + 2], v[i + 3]);
      }
      const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      const float mloc = sum * (1.0f / 32);
      float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float d0, d1, d2, d3;
        tc::sub2(d0, d1, v[i], v[i + 1], mloc, mloc);
        tc::sub2(d2, d3, v[i + 2], v[i + 3], mloc, mloc);
        tc::fma2(q[0], q[1], d0, d1, d0, d1, q[0], q[1]);
        tc::fma2(q[2], q[3], d2, d3, d2, d3, q[2], q[3]);
      }
      const int buf = n_ln & 1;
      ++n_ln;
      sm.red[buf][part][l] = make_float2(sum, (q[0] + q[1]) + (q[2] + q[3]));
      worker_sync16();
      const float2 p0 = sm.red[buf][0][l], p1 = sm.red[buf][1][l], p2 = sm.red[buf][2][l], p3 = sm.red[buf][3][l];
      const float mean = ((p0.x + p1.x) + (p2.x + p3.x)) * (1.0f / 128);
      const float d0 = p0.x * (1.0f / 32) - mean, d1 = p1.x * (1.0f / 32) - mean, d2 = p2.x * (1.0f / 32) - mean, d3 = p3.x * (1.0f / 32) - mean;
      const float m2 = ((p0.y + p1.y) + (p2.y + p3.y)) + 32.0f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
      const float rstd = 1.0f / sqrtf(m2 * (1.0f / 128) + LN_EPS);
      const float nmr = -mean * rstd;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float t0, t1;
        tc::fma2(t0, t1, v[i], v[i + 1], rstd, rstd, nmr, nmr);
        tc::fma2(v[i], v[i + 1], t0, t1, g[cq + i], g[cq + i + 1], bt[cq + i], bt[cq + i + 1]);
      }
    };
    float4 pf0 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch_params = [&](int p) {
      const int o0 = warp < 12 ? phase_vec(p, warp) : -1;
      if (o0 >= 0) pf0 = __ldg(reinterpret_cast<const float4*>(packed + o0) + lane);
    };
    auto commit_params = [&]() {
      if (warp < 12) reinterpret_cast<float4*>(sm.lp[(n_lp + 1) & 1][warp])[lane] = pf0;
    };
    fetch_params(0);
    if (warp < 12) reinterpret_cast<float4*>(sm.lp[0][warp])[lane] = pf0;
    worker_sync16();

    const float sc = 0.17677669529663687f * 1.4426950408889634f;

    if (part == 0 && upper == 0 && live) {  // per-rollout constants of the tail
      const int b2o[3] = {tbw::action_head_mlp_mean_0_fc_layers_2_bias, tbw::action_head_mlp_mean_1_fc_layers_2_bias,
                          tbw::action_head_mlp_mean_2_fc_layers_2_bias};
      const int lso[3] = {tbw::action_head_log_std_0, tbw::action_head_log_std_1, tbw::action_head_log_std_2};
      float b2x = 0.f, b2y = 0.f, ls[2] = {0.f, 0.f};
      for (int c3 = 0; c3 < 3; ++c3)
        if (sm.type[ag][c3]) {
          b2x += __ldg(packed + b2o[c3]);
          b2y += __ldg(packed + b2o[c3] + 1);
          ls[0] += __ldg(packed + lso[c3]);
          ls[1] += __ldg(packed + lso[c3] + 1);
        }
      float logp = 0.f;
      for (int d = 0; d < 2; ++d) logp += -logf(expf(ls[d])) - 0.91893853320467267f;
      sm.tailc[0][ag] = b2x;
      sm.tailc[1][ag] = b2y;
      sm.tailc[2][ag] = logp;
      sm.tailc[3][ag] = in.latent_logp[ba];
      for (int i = 0; i < 3; ++i) sm.tailc[4 + i][ag] = in.goal_gt ? in.goal_gt[sa * 4 + i] : 0.f;
      sm.tailc[7][ag] = in.agent_size[sa * 3] * 8.0f;
      long dst = in.dest[ba];
      dst = dst < 0 ? 0 : (dst >= dm.n_pl ? dm.n_pl - 1 : dst);
      const uint8_t* dtype = in.map_type + ((size_t)s * dm.n_pl + dst) * TB_PL_TYPE;
      const bool lane_t = dtype[0] || dtype[1] || dtype[2] || dtype[3], edge_t = dtype[4] != 0;
      sm.tailc[8][ag] = 50.0f * (1.0f - (edge_t ? 1.0f : 0.f) * 0.8f);
      sm.tflag[ag] = (uint8_t)((lane_t ? 1 : 0) | (edge_t ? 2 : 0));
      if (ag == 0)
        for (int i = 0; i < 4; ++i) sm.map_boundary[i] = in.map_boundary[(size_t)s * 4 + i];
    }

#pragma unroll 1
    for (int t = a.t_first; t <= a.t_last; ++t) {
      const int tl_t = min(t - 1, Th - 1);
      const int nkey_tl = uniform(in.n_key_tl[(size_t)s * Th + tl_t]);
#ifdef TB_TRACE_DETAIL
      t_cur = t;
#endif
      mark();
      const bool valid = sm.valid[ag] != 0;
      const unsigned vm_lo = __ballot_sync(0xffffffffu, sm.valid[lane] != 0);
      const unsigned vm_hi = __ballot_sync(0xffffffffu, sm.valid[lane + 32] != 0);
      const unsigned long long vmask = ((unsigned long long)vm_hi << 32) | vm_lo;
      int n_valid = __popc(vm_lo) + __popc(vm_hi);  // valid agents of the SCENE (decides the interaction bypass)
      unsigned long long vmask_peer = 0ull;
      if (halves) {
        // exchange the valid masks of the two agent halves through global memory: [parity][rank] (lo, hi) + a step counter per rank
        if (tid == 0) {
          int* x = a.sv.xch + (size_t)b * 16;
          volatile unsigned* vm = reinterpret_cast<volatile unsigned*>(x) + 4 * (t & 1);
          vm[2 * crank] = vm_lo;
          vm[2 * crank + 1] = vm_hi;
          __threadfence();
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(x + 8 + crank), "r"(t) : "memory");
          const long long t0 = clock64();
          int got;
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(got) : "l"(x + 8 + (crank ^ 1)) : "memory");
            if (got < t) {
              __nanosleep(64);
              tc::watchdog_check(t0);
            }
          } while (got < t);
          const unsigned plo = vm[2 * (crank ^ 1)], phi = vm[2 * (crank ^ 1) + 1];
          sm.peer_vm[0] = plo;
          sm.peer_vm[1] = phi;
          *reinterpret_cast<volatile int*>(&sm.n_valid) = n_valid + __popc(plo) + __popc(phi);
          __threadfence_block();
          mbar_arrive(&sm.cfg);
        }
        worker_sync16();
        vmask_peer = ((unsigned long long)sm.peer_vm[1] << 32) | sm.peer_vm[0];
        n_valid += __popc(sm.peer_vm[0]) + __popc(sm.peer_vm[1]);
      } else if (tid == 0) {
        *reinterpret_cast<volatile int*>(&sm.n_valid) = n_valid;
        __threadfence_block();
        mbar_arrive(&sm.cfg);
      }
      // ---- state embedding: 8 threads per agent; each computes 4 hidden units, then (after one barrier) 4 MLP outputs and 12 of
      // the 96 PE values = ONE function of one coordinate for the 12 frequencies of its group (same arithmetic order as before) --
      {
        const int fidx = 2 * part + upper;
        const float4 st = sm.pose[ag];
        {
          float at[12];
          at[0] = sm.vel[ag].x, at[1] = sm.vel[ag].y, at[2] = st.w, at[3] = sm.yaw_rate[ag], at[4] = sm.acc[ag];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            at[5 + i] = in.agent_size[sa * 3 + i];
            at[8 + i] = sm.type[ag][i] ? 1.f : 0.f;
          }
          at[11] = 0.f;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int o = 4 * fidx + u;
            float acc = sm.emb_b1[o];
#pragma unroll
            for (int k4 = 0; k4 < 3; ++k4) {
              const float4 w = *reinterpret_cast<const float4*>(&sm.emb_w1[(k4 * 32 + o) * 4]);
              acc = fmaf(at[4 * k4 + 3], w.w, fmaf(at[4 * k4 + 2], w.z, fmaf(at[4 * k4 + 1], w.y, fmaf(at[4 * k4], w.x, acc))));
            }
            sm.xo[o * MAXA + ag] = fmaxf(acc, 0.f);
          }
        }
        // PE while the hidden units of the other threads land: [cos(x f) | sin(x f) | cos(y f) | sin(y f) | cos(yaw g) 24 | sin(yaw g) 24]
        {
          const bool is_sin = fidx == 1 || fidx == 3 || fidx >= 6;
          const float arg = fidx < 2 ? st.x : fidx < 4 ? st.y : st.z;
          const float* ft = fidx < 4 ? &sm.f_xy[is_sin ? 1 : 0] : &sm.f_yaw[(is_sin ? 1 : 0) + ((fidx & 1) ? 24 : 0)];
#pragma unroll 4
          for (int i = 0; i < 12; ++i) {
            const float ph = arg * ft[2 * i];
            const float v = is_sin ? sinf(ph) : cosf(ph);
            sm.xs[(32 + 12 * fidx + i) * MAXA + ag] = valid ? v : 0.f;
          }
        }
        worker_sync16();
        {
          float h[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) h[k] = sm.xo[k * MAXA + ag];
          const float* w2 = packed + tbw::model_agent_encoder_mlp_fc_layers_3_weight;  // Wt4[8][32][4], through L1
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int o = 4 * fidx + u;
            float acc = sm.emb_b2[o];
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
              const float4 w = __ldg(reinterpret_cast<const float4*>(w2) + k4 * 32 + o);
              acc = fmaf(h[4 * k4 + 3], w.w, fmaf(h[4 * k4 + 2], w.z, fmaf(h[4 * k4 + 1], w.y, fmaf(h[4 * k4], w.x, acc))));
            }
            sm.xs[o * MAXA + ag] = valid ? acc : 0.f;
          }
        }
      }
      mark();

      // ---- 9 pre-LN cross-attention layers ------------------------------------------------------------------------------------
      const bool bypass = n_valid == 1;
#pragma unroll 1
      for (int Lx = 0; Lx < 9; ++Lx) {
        const int kind = Lx / 3;
        if (kind == 2 && bypass) break;
        const int nkey = kind == 0 ? nkey_map : kind == 1 ? nkey_tl : (n_valid > 0 ? (halves ? 2 * MAXA : MAXA) : 0);
        const int nblk = (nkey + 63) / 64;
        worker_sync16();
        dmark(100 + Lx * 10);
        const float (*lp)[128] = sm.lp[n_lp & 1];
        {
          int pn = Lx + 1;
          if (pn == 6 && bypass) pn = 9;
          fetch_params(pn);
        }
        float v[32];
        if (nblk > 0) {
          if (kind == 2) {
            if (Lx == 6) {
              load_x(v);
              if (upper == 0 && live) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x0_t[(cq / 4 + i) * A + agg] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
              }
            } else if (live) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 q = x0_t[(cq / 4 + i) * A + agg];
                v[4 * i] = q.x, v[4 * i + 1] = q.y, v[4 * i + 2] = q.z, v[4 * i + 3] = q.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            ln32(v, lp[8], lp[9]);
            write_A(T_A, v);
            signal_ready();  // -> Wk (ACC0), Wv (ACC1)
            dmark(500 + Lx);
            // LayerNorm 1 of the query side while the K | V MMAs run; its operand replaces the key-side one as soon as they are
            // done, and the Q projection (-> ACC2) then runs under the K | V epilogue below
            load_x(v);
            ln32(v, lp[0], lp[1]);
            wait_gemm();
            dmark(510 + Lx);
            write_A(T_A, v);
            signal_ready();  // -> Wq (ACC2)
            if (!halves) {
              tc::mbar_wait(&sm.grant, n_grant & 1);
              ++n_grant;
            }
            {
              // the 64-key block of this CTA's agents: a ring slot granted by the loader, or (halves) global memory, from where both
              // CTAs of the scene-mode stream it after the exchange
              unsigned char* blk = halves ? const_cast<unsigned char*>(cfg.kvx) + (Lx - 6) * cfg.kvx_layer_stride + (size_t)crank * BLK
                                          : sm.ring[*reinterpret_cast<volatile int*>(&sm.kvi_slot)];
              float kk[16];
              load_acc16(T_ACC0, kk);  // K[ag, ce .. ce+15]: key row ag
#pragma unroll
              for (int i = 0; i < 16; i += 2) tc::add2(kk[i], kk[i + 1], kk[i], kk[i + 1], lp[10][ce + i], lp[10][ce + i + 1]);
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                uint4 hi, lo;
                tc::split8(kk + 8 * c, hi, lo);
                const uint32_t off = (uint32_t)((ce >> 6) * 8192) + tc::sw128_off(ag, ((ce & 63) >> 3) + c);
                *reinterpret_cast<uint4*>(blk + off) = hi;
                *reinterpret_cast<uint4*>(blk + 16384 + off) = lo;
              }
              load_acc16(T_ACC1, kk);  // V[ag, ce .. ce+15] -> V^T rows d = ce + i, key column ag
              unsigned char* vt = blk + HALF + (ag & 7) * 2;
              const uint32_t kc = (uint32_t)(ag >> 3);
#pragma unroll
              for (int i = 0; i < 16; i += 2) {
                float v0, v1;
                tc::add2(v0, v1, kk[i], kk[i + 1], lp[11][ce + i], lp[11][ce + i + 1]);
                uint32_t hh, ll;
                tc::split_pair(v0, v1, hh, ll);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int d = ce + i + e;
                  const uint32_t off = (uint32_t)((d >> 3) * 1024 + (d & 7) * 128) + ((kc ^ (uint32_t)(d & 7)) << 4);
                  *reinterpret_cast<unsigned short*>(vt + off) = (unsigned short)(e ? hh >> 16 : hh & 0xffffu);
                  *reinterpret_cast<unsigned short*>(vt + 16384 + off) = (unsigned short)(e ? ll >> 16 : ll & 0xffffu);
                }
              }
              if (halves) {
                __threadfence();
                asm volatile("fence.proxy.async;" ::: "memory");
              } else {
                tc::fence_proxy_async();
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.wfill);
            dmark(540 + Lx);
          } else {
            load_x(v);
            ln32(v, lp[0], lp[1]);
            write_A(T_A, v);
            signal_ready();  // -> Wq
          }
          dmark(101 + Lx * 10);
          const bool split_layer = kind == 0 && n_cta > 1;
          // split layer: from here to the merge this CTA touches neither the LayerNorm exchange area nor the exchange buffer,
          // so it already arrives at the "exchange buffers free" barrier of the merge (the wait sits right before the pushes)
          if (split_layer) cluster_arrive_relaxed();
          if (!split_layer) commit_params();
          wait_gemm();
          dmark(102 + Lx * 10);
          if ((part & 1) == upper) {  // this thread's 32 Q columns = head `part` = the head lane l needs in pass part / 2
            const int hp = part >> 1;
            float q[32];
            tc::tmem_ld32(tm + (kind == 2 ? T_ACC2 : T_ACC0) + cq, q);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              tc::add2(q[i], q[i + 1], q[i], q[i + 1], lp[2][cq + i], lp[2][cq + i + 1]);
              tc::mul2(q[i], q[i + 1], q[i], q[i + 1], sc, sc);
            }
            float ph[16], pl[16], zz[16];
            tc::split32_packed(q, ph, pl);
#pragma unroll
            for (int i = 0; i < 16; ++i) zz[i] = 0.f;
            tc::tmem_st16(tm + T_A + 32 * hp + 16 * upper, ph);
            tc::tmem_st16(tm + T_A + 64 + 32 * hp + 16 * upper, pl);
            tc::tmem_st16(tm + T_A + 32 * hp + 16 * (1 - upper), zz);
            tc::tmem_st16(tm + T_A + 64 + 32 * hp + 16 * (1 - upper), zz);
          }
          signal_ready();  // -> QK^T(0, .)
          dmark(103 + Lx * 10);
          // ---- online softmax of pass `half`; this thread: keys 32 sub .. 32 sub + 31 of every block ---------------------------
          float m_ref = -INFINITY, l_sum = 0.f;
          const uint32_t sbase = tm + T_ACC0 + 64 * half;
          const uint32_t obase = tm + T_O + 64 * half + 32 * upper + 16 * sub;
          const bool split = kind == 0 && n_cta > 1;
          const int nblk_my = kind == 0 ? (nblk > rank ? (nblk - rank + n_cta - 1) / n_cta : 0) : nblk;
#pragma unroll 1
          for (int jb = 0; jb < nblk_my; ++jb) {
            const int key0 = (kind == 0 ? rank + jb * n_cta : jb) * 64 + 32 * sub;
            tc::mbar_wait(&sm.s[half], n_s & 1);
            ++n_s;
            tc::tc_fence_after();
            float sv_[32];
            tc::tmem_ld32(sbase + 32 * sub, sv_);
            tc::tmem_ld_wait();
            if (kind == 2) {
              // keys = the agents of the scene (halves: block jb holds the 64 agents of cluster rank jb), without the query itself
              const bool own = !halves || jb == crank;
              const unsigned en = (unsigned)(((own ? vmask & ~(1ull << ag) : vmask_peer)) >> (32 * sub));
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (!((en >> j) & 1u)) sv_[j] = -INFINITY;
            } else if (key0 + 32 > nkey) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (key0 + j >= nkey) sv_[j] = -INFINITY;
            }
            float mx4[4] = {sv_[0], sv_[1], sv_[2], sv_[3]};
#pragma unroll
            for (int j = 4; j < 32; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], sv_[j]);
            float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            sm.mxs[part][l] = mx;
            pair_sync(pair_id);  // both threads of the pass have loaded their logits and published their maxima
            mx = fmaxf(mx, sm.mxs[part ^ 1][l]);
            float alpha = 1.f;
            bool resc = false;
            if (mx > m_ref + 8.0f) {
              alpha = (m_ref == -INFINITY) ? 0.f : exp2f(m_ref - mx);
              m_ref = mx;
              l_sum *= alpha;
              resc = jb > 0;
            }
            const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
            float ps4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              tc::add2(sv_[j], sv_[j + 1], sv_[j], sv_[j + 1], neg_m, neg_m);
              tc::add2(sv_[j + 2], sv_[j + 3], sv_[j + 2], sv_[j + 3], neg_m, neg_m);
#pragma unroll
              for (int e = 0; e < 4; ++e) sv_[j + e] = ex2_approx(sv_[j + e]);
              tc::add2(ps4[0], ps4[1], ps4[0], ps4[1], sv_[j], sv_[j + 1]);
              tc::add2(ps4[2], ps4[3], ps4[2], ps4[3], sv_[j + 2], sv_[j + 3]);
            }
            l_sum += (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
            {
              float ph[16], pl[16];
              tc::split32_packed(sv_, ph, pl);
              tc::tmem_st16(sbase + 16 * sub, ph);
              tc::tmem_st16(sbase + 32 + 16 * sub, pl);
            }
            if (__any_sync(0xffffffffu, resc)) {
              float o[16];
              tc::tmem_ld16(obase, o);
              tc::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] *= alpha;
              tc::tmem_st16(obase, o);
            }
            tc::tmem_st_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.p[half]);
          }
          dmark(104 + Lx * 10);
          {
            // total softmax denominator of the (lane, pass): the two threads' partial sums
            float* lxs = &sm.mean_part[0][0][0];  // [4][128] (the action-head partials are idle here; separate from mxs and from the merge's areas)
            lxs[part * 128 + l] = l_sum;
            pair_sync(pair_id);
            const float l_tot = l_sum + lxs[(part ^ 1) * 128 + l];
            float o[16];
            if (nblk_my > 0) {
              tc::mbar_wait(&sm.o[half], n_o & 1);
              ++n_o;
              tc::tc_fence_after();
              tc::tmem_ld16(obase, o);
              tc::tmem_ld_wait();
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = 0.f;
            }
            float* xo_mine = &sm.xo[(64 * half + 32 * upper + 16 * sub) * MAXA + ag];  // this thread's 16 outputs, stride MAXA
            if (split) {
              const int nq = 4 / n_cta;  // float4 per thread and range (16 outputs = 4 float4)
              float4* xp = reinterpret_cast<float4*>(sm.xo);                          // slots [src rank][nq][512] float4
              float2* mls = &sm.red[0][0][0];  // [src rank][256 (lane, pass)] (m, l): the whole LayerNorm exchange area (idle here)
              const int lp_id = half * 128 + l;
              const uint32_t xp_addr = tc::smem_u32(xp), mls_addr = tc::smem_u32(mls);
              const uint32_t rs_addr = tc::smem_u32(&sm.rs_bar), ag_addr = tc::smem_u32(&sm.ag_bar);
              dmark(600);
              cluster_wait();  // (arrived after LayerNorm 1) every CTA is done with its previous use of the exchange buffers
              dmark(601);
              if (tid == 0) {  // what this CTA receives: one range + the (m, l) pairs from every peer; one merged range from every peer
                tc::mbar_expect_tx(&sm.rs_bar, (uint32_t)((n_cta - 1) * (nq * WORKERS16 * 16 + 256 * 8)));
                tc::mbar_expect_tx(&sm.ag_bar, (uint32_t)((n_cta - 1) * nq * WORKERS16 * 16));
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int dst = q / nq, qq = q % nq;
                const float4 val = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
                const uint32_t off = (uint32_t)(((rank * nq + qq) * WORKERS16 + tid) * 16);
                if (dst == rank) xp[(rank * nq + qq) * WORKERS16 + tid] = val;
                else st_async_f32x4(mapa(xp_addr, (uint32_t)dst) + off, val, mapa(rs_addr, (uint32_t)dst));
              }
              if (sub == 0) {
                for (int dst = 0; dst < n_cta; ++dst)
                  if (dst != rank)
                    st_async_f32x2(mapa(mls_addr, (uint32_t)dst) + (uint32_t)((rank * 256 + lp_id) * 8), make_float2(m_ref, l_tot),
                                   mapa(rs_addr, (uint32_t)dst));
              }
              dmark(602);
              mbar_wait_cluster(&sm.rs_bar, n_merge & 1);  // the peers' partials of my range have landed
              dmark(603);
              float m_all = -INFINITY;
              float2 mlr[4];
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                mlr[r] = r == rank ? make_float2(m_ref, l_tot) : r < n_cta ? mls[r * 256 + lp_id] : make_float2(-INFINITY, 0.f);
                m_all = fmaxf(m_all, mlr[r].x);
              }
              float l_all = 0.f, w[4];
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                w[r] = mlr[r].x == -INFINITY ? 0.f : exp2f(mlr[r].x - m_all);
                l_all = fmaf(w[r], mlr[r].y, l_all);
              }
              const float inv = l_all > 0.f ? 1.0f / l_all : 0.f;
              float4 mg[2];
#pragma unroll
              for (int qq = 0; qq < 2; ++qq) {
                mg[qq] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (qq < nq) {
#pragma unroll
                  for (int r = 0; r < 4; ++r) {
                    if (r < n_cta) {
                      const float4 pv = xp[(r * nq + qq) * WORKERS16 + tid];
                      tc::fma2(mg[qq].x, mg[qq].y, w[r], w[r], pv.x, pv.y, mg[qq].x, mg[qq].y);
                      tc::fma2(mg[qq].z, mg[qq].w, w[r], w[r], pv.z, pv.w, mg[qq].z, mg[qq].w);
                    }
                  }
                  mg[qq].x *= inv, mg[qq].y *= inv, mg[qq].z *= inv, mg[qq].w *= inv;
                }
              }
              dmark(604);
              cluster_sync_relaxed();  // every CTA has consumed its slots: the exchange buffer becomes the merged output
              dmark(605);
#pragma unroll
              for (int qq = 0; qq < 2; ++qq) {
                if (qq < nq) {
                  const uint32_t off = (uint32_t)((((rank * nq + qq) * WORKERS16) + tid) * 16);
                  for (int dst = 0; dst < n_cta; ++dst) {
                    if (dst == rank) xp[(rank * nq + qq) * WORKERS16 + tid] = mg[qq];
                    else st_async_f32x4(mapa(xp_addr, (uint32_t)dst) + off, mg[qq], mapa(ag_addr, (uint32_t)dst));
                  }
                }
              }
              dmark(606);
              mbar_wait_cluster(&sm.ag_bar, n_merge & 1);  // the merged attention output is complete in this CTA
              dmark(607);
              ++n_merge;
              commit_params();
            } else {
              const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                float t0, t1;
                tc::mul2(t0, t1, o[j], o[j + 1], inv, inv);
                xo_mine[j * MAXA] = t0;
                xo_mine[(j + 1) * MAXA] = t1;
              }
            }
          }
          worker_sync16();
          if (split) {  // merged outputs as [float4 q][worker thread]; my columns cq..cq+31 = head `part` of agent ag
            const float4* xp = reinterpret_cast<const float4*>(sm.xo);
            const int l_src = ag + 64 * (part & 1);  // lane that held head `part`: lower lanes hold even heads
#pragma unroll
            for (int sb = 0; sb < 2; ++sb) {
              const int src_tid = (((2 * (part >> 1) + sb) * 4 + (l_src >> 5)) << 5) + (l_src & 31);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 t4 = xp[q * WORKERS16 + src_tid];
                v[16 * sb + 4 * q] = t4.x, v[16 * sb + 4 * q + 1] = t4.y, v[16 * sb + 4 * q + 2] = t4.z, v[16 * sb + 4 * q + 3] = t4.w;
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = sm.xo[(cq + i) * MAXA + ag];
          }
          write_A(T_A, v);
          signal_ready();  // -> Wo
          dmark(105 + Lx * 10);
          wait_gemm();
          dmark(106 + Lx * 10);
          {
            float o16[16];
            load_acc16(T_ACC0, o16);
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              float t0, t1;
              tc::add2(t0, t1, o16[i], o16[i + 1], lp[3][ce + i], lp[3][ce + i + 1]);
              tc::add2(xs_at(ce + i), xs_at(ce + i + 1), xs_at(ce + i), xs_at(ce + i + 1), t0, t1);
            }
          }
          worker_sync16();
        } else {
          commit_params();
        }
        load_x(v);
        ln32(v, lp[4], lp[5]);
        write_A(T_A, v);
        signal_ready();  // -> W1
        dmark(107 + Lx * 10);
        wait_gemm();
        dmark(108 + Lx * 10);
        load_acc(T_ACC0, v);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          tc::add2(v[i], v[i + 1], v[i], v[i + 1], lp[6][cq + i], lp[6][cq + i + 1]);
          v[i] = fmaxf(v[i], 0.f), v[i + 1] = fmaxf(v[i + 1], 0.f);
        }
        write_A(T_A, v);
        signal_ready();  // -> W2
        dmark(109 + Lx * 10);
        wait_gemm();
        dmark(190 + Lx);
        {
          float y[16];
          load_acc16(T_ACC0, y);
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float t0, t1;
            tc::add2(t0, t1, xs_at(ce + i), xs_at(ce + i + 1), y[i], y[i + 1]);
            tc::add2(t0, t1, t0, t1, lp[7][ce + i], lp[7][ce + i + 1]);
            xs_at(ce + i) = valid ? t0 : 0.f, xs_at(ce + i + 1) = valid ? t1 : 0.f;
          }
        }
        ++n_lp;
      }
      mark();

      // ---- agent_temporal: 3-layer GRU -------------------------------------------------------------------------------------------
#pragma unroll 1
      for (int L = 0; L < 3; ++L) {
        // h_{t-1} of this layer (written one step ago) is requested before the barrier: its L2 latency hides behind the barrier and
        // the x operand
        float4* hid = hid_t + (size_t)L * B * 32 * A + agg;
        float4 hq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) hq[i] = live ? hid[(cq / 4 + i) * A] : make_float4(0.f, 0.f, 0.f, 0.f);
        worker_sync16();
        const float (*lp)[128] = sm.lp[n_lp & 1];
        dmark(200 + L * 10);
        fetch_params(10 + L);
        {
          float x[32];
          load_x(x);
          write_A(T_A, x);
#pragma unroll
          for (int i = 0; i < 8; ++i) x[4 * i] = hq[i].x, x[4 * i + 1] = hq[i].y, x[4 * i + 2] = hq[i].z, x[4 * i + 3] = hq[i].w;
          write_A(T_A2, x);
        }
        signal_ready();
        dmark(201 + L * 10);
        commit_params();
        wait_gemm();
        dmark(202 + L * 10);
        {
          float r[16], rh[16];
          tc::tmem_ld16(tm + T_ACC0 + ce, r);
          tc::tmem_ld16(tm + T_ACC1 + ce, rh);
          tc::tmem_ld_wait();
          // both accumulators are in registers: the second MMA batch (same operands, z and n gates) may overwrite them now and
          // runs under the gate math below
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.ready);
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float a0, a1, h0, h1;
            tc::add2(a0, a1, r[i], r[i + 1], lp[0][ce + i], lp[0][ce + i + 1]);
            tc::add2(a0, a1, a0, a1, lp[3][ce + i], lp[3][ce + i + 1]);
            tc::add2(h0, h1, rh[i], rh[i + 1], lp[5][ce + i], lp[5][ce + i + 1]);
            tc::mul2(h0, h1, fast_sigmoid(a0), fast_sigmoid(a1), h0, h1);
            sm.xo[(ce + i) * MAXA + ag] = h0;
            sm.xo[(ce + i + 1) * MAXA + ag] = h1;
          }
        }
        dmark(203 + L * 10);
        float4 hp4v[4];  // this thread's 16 columns of h_{t-1}, in flight during the second MMA batch
#pragma unroll
        for (int i = 0; i < 4; ++i) hp4v[i] = live ? hid[(ce / 4 + i) * A] : make_float4(0.f, 0.f, 0.f, 0.f);
        wait_gemm();
        dmark(204 + L * 10);
        {
          float z[16], n[16];
          tc::tmem_ld16(tm + T_ACC0 + ce, z);
          tc::tmem_ld16(tm + T_ACC1 + ce, n);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float hp_[4] = {hp4v[i].x, hp4v[i].y, hp4v[i].z, hp4v[i].w};
            float hn_[4];
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
              const int c = ce + 4 * i + e;
              float z0, z1, n0, n1;
              tc::add2(z0, z1, z[4 * i + e], z[4 * i + e + 1], lp[1][c], lp[1][c + 1]);
              tc::add2(z0, z1, z0, z1, lp[4][c], lp[4][c + 1]);
              tc::add2(n0, n1, n[4 * i + e], n[4 * i + e + 1], lp[2][c], lp[2][c + 1]);
              tc::add2(n0, n1, n0, n1, sm.xo[c * MAXA + ag], sm.xo[(c + 1) * MAXA + ag]);
              const float zg0 = fast_sigmoid(z0), zg1 = fast_sigmoid(z1), ng0 = fast_tanh(n0), ng1 = fast_tanh(n1);
              hn_[e] = (1.0f - zg0) * ng0 + zg0 * hp_[e];
              hn_[e + 1] = (1.0f - zg1) * ng1 + zg1 * hp_[e + 1];
            }
            if (live) hid[(ce / 4 + i) * A] = valid ? make_float4(hn_[0], hn_[1], hn_[2], hn_[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int e = 0; e < 4; ++e) xs_at(ce + 4 * i + e) = (L == 2 && !valid) ? 0.f : hn_[e];
          }
        }
        ++n_lp;
      }
      mark();

      // ---- add_goal, add_latent ----------------------------------------------------------------------------------------------------
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {
        // z half of mlp_out layer 0 (W[:, 128:256] relu(z), step-invariant, from k_rollout_init): requested before the barrier, added
        // in the epilogue where z is valid
        const bool zv = j == 0 ? (sm.goal_valid[ag] != 0) : valid;
        const float4* zc = (j == 0 ? goal_c_t : latent_c_t) + (cq / 4) * A + agg;
        float4 zq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) zq[i] = (live && zv) ? __ldg(zc + i * A) : make_float4(0.f, 0.f, 0.f, 0.f);
        worker_sync16();
        const float (*lp)[128] = sm.lp[n_lp & 1];
        fetch_params(13 + j);
        {
          float z[32];
          load_x(z);
          write_A(T_A, z);
        }
        signal_ready();
        commit_params();
        wait_gemm();
        {
          float h1[32];
          load_acc(T_ACC0, h1);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            tc::add2(h1[4 * i], h1[4 * i + 1], h1[4 * i], h1[4 * i + 1], zq[i].x, zq[i].y);
            tc::add2(h1[4 * i + 2], h1[4 * i + 3], h1[4 * i + 2], h1[4 * i + 3], zq[i].z, zq[i].w);
            tc::add2(h1[4 * i], h1[4 * i + 1], h1[4 * i], h1[4 * i + 1], lp[0][cq + 4 * i], lp[0][cq + 4 * i + 1]);
            tc::add2(h1[4 * i + 2], h1[4 * i + 3], h1[4 * i + 2], h1[4 * i + 3], lp[0][cq + 4 * i + 2], lp[0][cq + 4 * i + 3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) h1[4 * i + e] = fmaxf(h1[4 * i + e], 0.f);
          }
          write_A(T_A, h1);
        }
        signal_ready();
        wait_gemm();
        {
          float h2[16];
          load_acc16(T_ACC0, h2);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float hz = fmaxf(h2[i] + lp[1][ce + i], 0.f);
            xs_at(ce + i) = valid ? (zv ? hz : 0.f) + xs_at(ce + i) : 0.f;
          }
        }
        ++n_lp;
      }
      // ---- action head ----------------------------------------------------------------------------------------------------------------
      const bool tail_thread = upper == 0 && live;  // the four column-part threads of an agent share the tail (below)
      const bool out_w = rank == 0;
      const bool has_gt = t < Tg;
      const size_t gidx = ((size_t)s * Tg + (has_gt ? t : 0)) * A + agg;
      float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
      float2 g_vel = make_float2(0.f, 0.f);
      float g_acc = 0.f, g_yr = 0.f;
      bool ovr = false, gt_valid = false;
      {
        worker_sync16();
        const float (*lp)[128] = sm.lp[n_lp & 1];
        fetch_params(0);
        {
          float x[32];
          load_x(x);
          write_A(T_A, x);
          if (a.out.trace_policy_feature && writer) {
            float* dst = a.out.trace_policy_feature + ((ba * T) + (t - 1)) * D + cq;
#pragma unroll
            for (int i = 0; i < 8; ++i) reinterpret_cast<float4*>(dst)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
          }
        }
        signal_ready();
        if (tail_thread && has_gt) {
          ovr = in.tf_mask[gidx] != 0;
          gt_valid = in.gt_valid[gidx] != 0;
          gs = make_float4(in.gt_pos[gidx * 2], in.gt_pos[gidx * 2 + 1], in.gt_yaw[gidx], in.gt_spd[gidx]);
          g_vel = make_float2(in.gt_vel[gidx * 2], in.gt_vel[gidx * 2 + 1]);
          g_acc = in.gt_acc[gidx];
          g_yr = in.gt_yaw_rate[gidx];
        }
        commit_params();
        wait_gemm();
        {
          float m0 = 0.f, m1 = 0.f;
#pragma unroll 1
          for (int c3 = 0; c3 < 3; ++c3) {
            float hdn[16];
            tc::tmem_ld16(tm + 128 * c3 + ce, hdn);
            tc::tmem_ld_wait();
            const bool on = sm.type[ag][c3] && valid;
            const float* w2 = &lp[3 + 2 * c3][0];
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float hv = fmaxf(hdn[i] + lp[c3][ce + i], 0.f);
              const int k = ce + i;
              tc::fma2(s0, s1, hv, hv, w2[((k >> 2) * 2 + 0) * 4 + (k & 3)], w2[((k >> 2) * 2 + 1) * 4 + (k & 3)], s0, s1);
            }
            if (on) {
              m0 += s0;
              m1 += s1;
            }
          }
          sm.mean_part[2 * part + upper][ag][0] = m0;
          sm.mean_part[2 * part + upper][ag][1] = m1;
        }
        ++n_lp;
        worker_sync16();
      }
      mark();

      dmark(400);
      // ---- per-agent tail: the 4 threads (part 0..3, lower lane) of an agent all integrate the dynamics (same inputs, same
      // arithmetic), then split the rest: part 0 outputs / override / map boundary / kill, part 1 goal check + reward, parts 2
      // and 3 ten destination nodes each; part 0 combines the flags after one barrier --------------------------------------
      float4 ns = make_float4(0.f, 0.f, 0.f, 0.f);
      bool nvalid = false, killed = false, outside = false, out_t = false;
      const size_t o = ba * T + (t - 1);
      uint8_t* tflags = reinterpret_cast<uint8_t*>(&sm.mxs[0][0]);  // [4 parts][MAXA] result flags (the softmax exchange area is idle here)
      if (tail_thread) {
        float mean0 = 0.f, mean1 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          mean0 += sm.mean_part[k][ag][0];
          mean1 += sm.mean_part[k][ag][1];
        }
        if (valid) {
          mean0 += sm.tailc[0][ag];
          mean1 += sm.tailc[1][ag];
        }
        const bool ty0 = sm.type[ag][0], ty1 = sm.type[ag][1], ty2 = sm.type[ag][2];
        const bool k_has_type = ty0 || ty1 || ty2;
        const float k_max_acc = (ty0 ? 5.0f : 0.f) + (ty1 ? 7.0f : 0.f) + (ty2 ? 6.0f : 0.f);
        const float k_max_yr = (ty0 ? 1.5f : 0.f) + (ty1 ? 7.0f : 0.f) + (ty2 ? 3.0f : 0.f);
        const float a_acc = valid ? tanhf(mean0) * k_max_acc : 0.f;
        const float a_yr = valid ? tanhf(mean1) * k_max_yr : 0.f;
        const float4 st = sm.pose[ag];
        const float v_t = st.w + 0.05f * a_acc, th_t = st.z + 0.05f * a_yr;
        float4 pred = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && k_has_type) {
          pred.x = st.x + 0.1f * (v_t * cosf(th_t));
          pred.y = st.y + 0.1f * (v_t * sinf(th_t));
          pred.z = st.z + 0.1f * a_yr;
          pred.w = st.w + 0.1f * a_acc;
        }
        killed = sm.killed[ag] != 0;
        const bool m = ovr && !killed;
        nvalid = valid || m;
        ns = m ? gs : pred;
        if (part == 0) {
          if (out_w) {
            *reinterpret_cast<float4*>(a.out.preds + o * 4) = pred;
            a.out.valid[o] = valid;
            a.out.action_log_probs[o] = valid ? sm.tailc[2][ag] : 0.f;
            a.out.latent_log_probs[o] = sm.tailc[3][ag];
            if (a.out.trace_action_mean) {
              a.out.trace_action_mean[o * 2] = mean0;
              a.out.trace_action_mean[o * 2 + 1] = mean1;
            }
            a.out.override_masks[o] = ovr;
          }
          out_t = nvalid && (ns.x > sm.map_boundary[1] || ns.x < sm.map_boundary[0] || ns.y > sm.map_boundary[3] || ns.y < sm.map_boundary[2]);
          outside = (sm.sticky[0][ag] != 0) || out_t;
        } else if (part == 1) {
          bool goal_t = false;
          if (in.goal_gt) {
            const float dx = ns.x - sm.tailc[4][ag], dy = ns.y - sm.tailc[5][ag];
            const bool pos_ok = sqrtf(dx * dx + dy * dy) < sm.tailc[7][ag];
            const float PI_F = 3.14159265358979323846f, TWO_PI_F = 6.28318530717958647692f;
            float w = fmodf(ns.z - sm.tailc[6][ag] + PI_F, TWO_PI_F);
            if (w < 0.f) w += TWO_PI_F;
            const bool rot_ok = fabsf(w - PI_F) < 0.26179938779914943654f;
            goal_t = pos_ok && rot_ok && nvalid && !(sm.sticky[1][ag] != 0);
          }
          tflags[1 * MAXA + ag] = goal_t;
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
          if (out_w) {
            a.out.diffbar_rewards[o] = reward;
            a.out.diffbar_rewards_valid[o] = rv;
          }
        } else {
          bool pos_reached = false, rot_reached = false;
          const float k_dest_thresh = sm.tailc[8][ag];
          const float hx = cosf(ns.z), hy = sinf(ns.z);
          const float4* dn = a.sv.dest_nodes + ((size_t)b * TB_PL_NODE + 10 * (part - 2)) * A + agg;
          float4 nd[10];
#pragma unroll
          for (int n = 0; n < 10; ++n) nd[n] = __ldg(dn + n * A);
#pragma unroll
          for (int n = 0; n < 10; ++n) {
            const float dx = ns.x - nd[n].x, dy = ns.y - nd[n].y;
            pos_reached |= sqrtf(dx * dx + dy * dy) < k_dest_thresh;
            rot_reached |= (hx * nd[n].z + hy * nd[n].w) > 0.86602540378443864676f;
          }
          tflags[part * MAXA + ag] = (uint8_t)((pos_reached ? 1 : 0) | (rot_reached ? 2 : 0));
        }
      }
      worker_sync16();
      if (tail_thread && part == 0) {
        const bool m = ovr && !killed;
        if (m) {
          sm.vel[ag] = g_vel;
          sm.acc[ag] = g_acc;
          sm.yaw_rate[ag] = g_yr;
        }
        bool goal_r = sm.sticky[1][ag] != 0, dest_r = sm.sticky[2][ag] != 0;
        const bool goal_t = tflags[1 * MAXA + ag] != 0;
        goal_r |= goal_t;
        const uint8_t f23 = tflags[2 * MAXA + ag] | tflags[3 * MAXA + ag];
        const bool pos_reached = (f23 & 1) != 0, rot_reached = (f23 & 2) != 0;
        const bool k_lane_t = (sm.tflag[ag] & 1) != 0, k_edge_t = (sm.tflag[ag] & 2) != 0;
        const bool dest_t = !dest_r && nvalid && ((k_lane_t && pos_reached && rot_reached) || (k_edge_t && pos_reached));
        dest_r |= dest_t;
        if (out_w) {
          const size_t vs = BA * T;
          a.out.violations[0 * vs + o] = outside;
          a.out.violations[1 * vs + o] = out_t;
          a.out.violations[2 * vs + o] = goal_r;
          a.out.violations[3 * vs + o] = goal_t;
          a.out.violations[4 * vs + o] = dest_r;
          a.out.violations[5 * vs + o] = dest_t;
        }
        const bool kill = out_t && !gt_valid;
        killed |= kill;
        nvalid = nvalid && !kill;
        const bool gv = sm.goal_valid[ag] && nvalid && !dest_r;
        sm.pose[ag] = ns;
        sm.valid[ag] = nvalid;
        sm.killed[ag] = killed;
        sm.goal_valid[ag] = gv;
        sm.sticky[0][ag] = outside;
        sm.sticky[1][ag] = goal_r;
        sm.sticky[2][ag] = dest_r;
      }
      dmark(401);
      worker_sync16();
      dmark(402);
    }
    mark();
    worker_sync16();
    if (rank == 0) {
      for (int L = 0; L < 3; ++L) {
        float4* dst = reinterpret_cast<float4*>(a.sv.hidden + ((size_t)L * BA + (size_t)b * A) * D);
        const float4* src = hid_t + (size_t)L * B * 32 * A;
        for (int i = tid; i < A_loc * 32; i += WORKERS16) {
          const int ag_ = a0 + i / 32, c4 = i % 32;
          dst[ag_ * 32 + c4] = __ldcg(src + c4 * A + ag_);
        }
      }
    }
    if (part == 0 && writer) {
      *reinterpret_cast<float4*>(a.sv.agent_state + ba * 4) = sm.pose[ag];
      a.sv.vel[ba * 2] = sm.vel[ag].x;
      a.sv.vel[ba * 2 + 1] = sm.vel[ag].y;
      a.sv.acc[ba] = sm.acc[ag];
      a.sv.yaw_rate[ba] = sm.yaw_rate[ag];
      a.sv.valid[(size_t)((a.t_last + 1) & 1) * BA + ba] = sm.valid[ag];
      a.sv.killed[ba] = sm.killed[ag];
      a.sv.goal_valid[ba] = sm.goal_valid[ag];
      for (int i = 0; i < 3; ++i) a.sv.sticky[(size_t)i * BA + ba] = sm.sticky[i][ag];
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(sm.tmem_base, 512);
  cluster_sync_all();
}

}  // namespace pr
}  // namespace tb

using namespace tb;

bool tb::rollout_tc_supported(const TbDims& d, const TbRolloutIn& in) {
  if (!in.kv_map_tc || !in.kv_tl_tc || !in.n_key_map || !in.n_key_tl) return false;
  return d.n_agent <= 2 * pr::MAXA;
}

int tb::rollout_tc_cluster_size(const TbDims& d) {
  // One CTA per scene-mode leaves most of the 148 SMs idle for small batches; the agent->map attention (the largest part of
  // a step) is then split over a cluster of 2 or 4 CTAs per scene-mode.
  if (d.n_agent > pr::MAXA) return 2;  // two CTAs per scene-mode, 64 agents each (the agent->map attention is not split)
  const char* env = getenv("TB_CLUSTER");  // development / test override
  const int forced = env ? atoi(env) : 0;
  if (forced == 1 || forced == 2 || forced == 4) return forced;
  if (d.n_cta_per_mode == 1 || d.n_cta_per_mode == 2 || d.n_cta_per_mode == 4) return d.n_cta_per_mode;
  const int B = d.n_scene * d.n_mode;
  return B * 4 <= 148 ? 4 : (B * 2 <= 148 ? 2 : 1);
}

int tb::launch_rollout_tc(const TbDims& d, const TbRolloutIn& in, const float* packed, const StateView& sv, const TbRolloutOut& out,
                          int t_first, int t_last, cudaStream_t st) {
  pr::Args a{d, in, packed, tc_blob(packed), sv, out, t_first, t_last, g_debug_trace};
  const int n_cta = rollout_tc_cluster_size(d);
  static std::atomic<uint64_t> attr16_set{0};
  const int smem16 = (int)sizeof(pr::Smem16) + 1024;
  if (!smem_attr_done(attr16_set)) {
    if (!set_max_smem(pr::k_rollout_tc16, smem16)) return TB_ERR_LAUNCH;
    smem_attr_mark(attr16_set);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(d.n_scene * d.n_mode * n_cta);
  cfg.blockDim = dim3(pr::THREADS16);
  cfg.dynamicSmemBytes = smem16;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = n_cta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, pr::k_rollout_tc16, a) != cudaSuccess) return TB_ERR_LAUNCH;
  count_launch();
  return launch_status();
}
