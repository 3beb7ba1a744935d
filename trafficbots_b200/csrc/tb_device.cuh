// Device-side building blocks shared by every kernel of the TrafficBots hot path (sm_100a, fp32).
//
// Work decomposition: a CTA of 256 threads owns a tile of R = 4*RPT feature rows (agents, polyline nodes or
// polylines) that live in shared memory for the whole kernel; every Linear of the reference becomes a
// "row-tile x packed-weight" GEMM whose weights stream from L2 through the read-only path as float4 (packed
// layout: tb_weights_gen.h), LayerNorm / masks / activations are fused as epilogues, and attention is a
// flash-style loop over 64-key tiles of the pre-projected K|V cache (nothing of size [rows, keys] is ever
// written to global memory; the reference materialises [B,4,A,P] logits and probabilities, attention.py:115-130).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "tb_weights_gen.h"

namespace tb {

constexpr int D = 128;         // hidden_dim
constexpr int NT = 256;        // threads per CTA
constexpr int NWARP = NT / 32;
constexpr int NHEAD = 4;
constexpr int DH = 32;
constexpr int TK = 64;         // keys per attention tile
constexpr int KPAD = D + 4;    // K tile row stride (floats): conflict-free float4 reads with one key per lane
constexpr int SPAD = TK + 4;   // probability tile row stride
constexpr float LN_EPS = 1e-5f;

// ---- offsets inside one packed TransformerCrossAttention layer (relative to its norm1.weight) --------------
namespace tfl {
constexpr int BASE = tbw::model_transformer_as2pl_layers_0_norm1_weight;
constexpr int NORM1_W = 0;
constexpr int NORM1_B = tbw::model_transformer_as2pl_layers_0_norm1_bias - BASE;
constexpr int NORMT_W = tbw::model_transformer_as2pl_layers_0_norm_tgt_weight - BASE;
constexpr int NORMT_B = tbw::model_transformer_as2pl_layers_0_norm_tgt_bias - BASE;
constexpr int IN_W = tbw::model_transformer_as2pl_layers_0_attn_in_proj_weight - BASE;   // Wt4[32][384][4]
constexpr int OUT_W = tbw::model_transformer_as2pl_layers_0_attn_out_proj_weight - BASE;
constexpr int IN_B = tbw::model_transformer_as2pl_layers_0_attn_in_proj_bias - BASE;
constexpr int OUT_B = tbw::model_transformer_as2pl_layers_0_attn_out_proj_bias - BASE;
constexpr int L1_W = tbw::model_transformer_as2pl_layers_0_linear1_weight - BASE;
constexpr int L1_B = tbw::model_transformer_as2pl_layers_0_linear1_bias - BASE;
constexpr int L2_W = tbw::model_transformer_as2pl_layers_0_linear2_weight - BASE;
constexpr int L2_B = tbw::model_transformer_as2pl_layers_0_linear2_bias - BASE;
constexpr int NORM2_W = tbw::model_transformer_as2pl_layers_0_norm2_weight - BASE;
constexpr int NORM2_B = tbw::model_transformer_as2pl_layers_0_norm2_bias - BASE;
constexpr int STRIDE = tbw::model_transformer_as2pl_layers_1_norm1_weight - BASE;
static_assert(tbw::model_agent_interaction_transformer_layers_1_norm1_weight -
                      tbw::model_agent_interaction_transformer_layers_0_norm1_weight == STRIDE, "layer stride");
static_assert(tbw::model_map_encoder_transformer_densetnt_layers_0_norm2_bias -
                      tbw::model_map_encoder_transformer_densetnt_layers_0_norm1_weight == NORM2_B, "layer layout");
}  // namespace tfl

// ---- offsets inside one packed GRU layer (relative to weight_ih_l0) ----------------------------------------
namespace gru {
constexpr int BASE = tbw::model_agent_temporal_rnn_weight_ih_l0;
constexpr int W_IH = 0;                                                        // Wt4[32][384][4], gate order r,z,n
constexpr int W_HH = tbw::model_agent_temporal_rnn_weight_hh_l0 - BASE;
constexpr int B_IH = tbw::model_agent_temporal_rnn_bias_ih_l0 - BASE;
constexpr int B_HH = tbw::model_agent_temporal_rnn_bias_hh_l0 - BASE;
constexpr int STRIDE = tbw::model_agent_temporal_rnn_weight_ih_l1 - BASE;
}  // namespace gru

__host__ __device__ inline int block_base(int block) {
  switch (block) {
    case 0: return tbw::model_map_encoder_transformer_densetnt_layers_0_norm1_weight;
    case 1: return tbw::model_map_encoder_transformer_self_attn_layers_0_norm1_weight;
    case 2: return tbw::model_transformer_as2pl_layers_0_norm1_weight;
    case 3: return tbw::model_transformer_as2tl_layers_0_norm1_weight;
    case 4: return tbw::model_agent_interaction_transformer_layers_0_norm1_weight;
    case 5: return tbw::model_latent_encoder_agent_interaction_prior_transformer_layers_0_norm1_weight;
    case 6: return tbw::model_latent_encoder_agent_interaction_post_transformer_layers_0_norm1_weight;
  }
  return -1;
}
__host__ __device__ inline int block_layers(int block) { return block == 1 ? 1 : 3; }

// ---- small helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- row-tile GEMM ------------------------------------------------------------------------------------------
// acc[i][c] += sum_k xs[(rg*RPT+i)*ldx + k] * W[n0 + cg + 64*c][k]     (k < 4*K4)
// thread (cg = tid&63, rg = tid>>6) owns rows rg*RPT..+RPT-1 and columns cg, cg+64 of a 128-column slab.
// `w` points at a packed Wt4[K4][ldn][4] tensor.  All lanes of a warp share rg -> the xs reads are broadcasts.
template <int RPT>
__device__ __forceinline__ void gemm_acc(const float* __restrict__ w, int ldn, int n0, int K4, const float* xs,
                                         int ldx, float (&acc)[RPT][2]) {
  const int cg = threadIdx.x & 63, rg = threadIdx.x >> 6;
  const float4* __restrict__ w4 = reinterpret_cast<const float4*>(w) + n0 + cg;
  const float* xr = xs + rg * RPT * ldx;
#pragma unroll 4
  for (int k4 = 0; k4 < K4; ++k4) {
    const float4 a = __ldg(w4 + (size_t)k4 * ldn);
    const float4 b = __ldg(w4 + (size_t)k4 * ldn + 64);
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const float4 x = *reinterpret_cast<const float4*>(xr + i * ldx + 4 * k4);
      acc[i][0] = fmaf(x.w, a.w, fmaf(x.z, a.z, fmaf(x.y, a.y, fmaf(x.x, a.x, acc[i][0]))));
      acc[i][1] = fmaf(x.w, b.w, fmaf(x.z, b.z, fmaf(x.y, b.y, fmaf(x.x, b.x, acc[i][1]))));
    }
  }
}

// Y = epi(row, col, X W^T) over a 128-column slab; epi is called once per owned element.
template <int RPT, class Epi>
__device__ __forceinline__ void gemm128(const float* __restrict__ w, int ldn, int n0, int K4, const float* xs,
                                        int ldx, Epi epi) {
  float acc[RPT][2];
#pragma unroll
  for (int i = 0; i < RPT; ++i) acc[i][0] = acc[i][1] = 0.f;
  gemm_acc<RPT>(w, ldn, n0, K4, xs, ldx, acc);
  const int cg = threadIdx.x & 63, rg = threadIdx.x >> 6;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    epi(rg * RPT + i, cg, acc[i][0]);
    epi(rg * RPT + i, cg + 64, acc[i][1]);
  }
}

// same with the K dimension split over two shared-memory sources (torch.cat([x1, x2], -1) @ W^T)
template <int RPT, class Epi>
__device__ __forceinline__ void gemm128_cat(const float* __restrict__ w, int ldn, int n0, const float* xs1, int ldx1,
                                            int K4a, const float* xs2, int ldx2, int K4b, Epi epi) {
  float acc[RPT][2];
#pragma unroll
  for (int i = 0; i < RPT; ++i) acc[i][0] = acc[i][1] = 0.f;
  gemm_acc<RPT>(w, ldn, n0, K4a, xs1, ldx1, acc);
  gemm_acc<RPT>(w + (size_t)K4a * ldn * 4, ldn, n0, K4b, xs2, ldx2, acc);
  const int cg = threadIdx.x & 63, rg = threadIdx.x >> 6;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    epi(rg * RPT + i, cg, acc[i][0]);
    epi(rg * RPT + i, cg + 64, acc[i][1]);
  }
}

// narrow outputs (N < 128): one (row, col) per loop trip, columns fastest -> coalesced weight reads
template <class Epi>
__device__ __forceinline__ void gemm_small(const float* __restrict__ w, int N, int K4, const float* xs, int ldx,
                                           int nrow, Epi epi) {
  const float4* __restrict__ w4 = reinterpret_cast<const float4*>(w);
  for (int idx = threadIdx.x; idx < nrow * N; idx += NT) {
    const int n = idx % N, r = idx / N;
    float acc = 0.f;
    for (int k4 = 0; k4 < K4; ++k4) {
      const float4 a = __ldg(w4 + k4 * N + n);
      const float4 x = *reinterpret_cast<const float4*>(xs + r * ldx + 4 * k4);
      acc = fmaf(x.w, a.w, fmaf(x.z, a.z, fmaf(x.y, a.y, fmaf(x.x, a.x, acc))));
    }
    epi(r, n, acc);
  }
}

// ---- LayerNorm over the rows of a shared-memory tile (eps 1e-5, affine; two-pass like ATen) ------------------
__device__ __forceinline__ void layernorm_rows(const float* xs, int ldx, float* ys, int ldy, int nrow,
                                               const float* __restrict__ g, const float* __restrict__ b) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + lane);
  const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + lane);
  for (int r = warp; r < nrow; r += NWARP) {
    const float4 v = *reinterpret_cast<const float4*>(xs + r * ldx + lane * 4);
    const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.0f / D);
    const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
    const float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / D);
    const float rstd = 1.0f / sqrtf(var + LN_EPS);
    float4 o;
    o.x = dx * rstd * gg.x + bb.x;
    o.y = dy * rstd * gg.y + bb.y;
    o.z = dz * rstd * gg.z + bb.z;
    o.w = dw * rstd * gg.w + bb.w;
    *reinterpret_cast<float4*>(ys + r * ldy + lane * 4) = o;
  }
}

// ---- sinusoidal pose embedding (utils/pose_pe.py:57-62, utils/pos_emb.py:23-26,53-56), 96 values ------------
// out[0:12]=cos(x f_i) out[12:24]=sin(x f_i) out[24:36]=cos(y f_i) out[36:48]=sin(y f_i)
// out[48:72]=cos(k yaw) out[72:96]=sin(k yaw); f = the registered `freqs` buffers (each value stored twice).
// Arguments reach hundreds of radians: full-range sinf/cosf, never the __sinf fast path.
__device__ __forceinline__ void pose_pe_elem(int j, float x, float y, float yaw, const float* __restrict__ f_xy,
                                             const float* __restrict__ f_yaw, float* out) {
  // j in [0,48): one (cos, sin) pair
  if (j < 12) {
    const float e = x * __ldg(f_xy + 2 * j);
    out[j] = cosf(e);
    out[12 + j] = sinf(x * __ldg(f_xy + 2 * j + 1));
  } else if (j < 24) {
    const int i = j - 12;
    out[24 + i] = cosf(y * __ldg(f_xy + 2 * i));
    out[36 + i] = sinf(y * __ldg(f_xy + 2 * i + 1));
  } else {
    const int i = j - 24;
    out[48 + i] = cosf(yaw * __ldg(f_yaw + 2 * i));
    out[72 + i] = sinf(yaw * __ldg(f_yaw + 2 * i + 1));
  }
}

// ---- shared-memory working set of a row tile -------------------------------------------------------------------
template <int R>
struct TileSmem {
  float x[R * D];          // residual stream of the tile
  float t[R * D];          // LayerNorm output / attention output
  float q[R * D];          // queries / FFN hidden
  float ks[TK * KPAD];     // key tile
  float vs[TK * D];        // value tile
  float s[R * NHEAD * SPAD];  // logits -> probabilities of the current key tile
  float m[R * NHEAD];      // running max
  float l[R * NHEAD];      // running sum
  float alpha[R * NHEAD];  // rescale factor of the current tile
  uint8_t kvalid[TK];
  uint8_t row_valid[R];
  uint8_t dead[R];
};

// Flash-style masked multi-head attention of the R query rows in sm.q against n_key pre-projected keys.
//   kv        [n_key,256] global: K = cols 0..127, V = cols 128..255 (already LN_tgt + projected + biased)
//   key_valid [n_key] global (attention.py:91-94 key padding mask)
//   self_base >= 0: query row r may not attend key (self_base + r)   (eye attn_mask, agent_interaction.py:57-59)
// Result: sm.t[r][:] = concat_h softmax(q_h K_h^T / sqrt(32)) V_h, sm.dead[r] = 1 (and a zero row) when the row
// has no enabled key (attention.py:101-107: the reference un-masks such rows and discards the result later).
template <int R>
__device__ void attention_tile(TileSmem<R>& sm, const float* __restrict__ kv, const uint8_t* __restrict__ key_valid,
                               int n_key, int self_base) {
  static_assert(R % 2 == 0, "R");
  constexpr int RH = R / 2;  // PV: rows per thread
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float scale = 0.17677669529663687f;  // 1/sqrt(32)

  for (int i = tid; i < R * NHEAD; i += NT) {
    sm.m[i] = -INFINITY;
    sm.l[i] = 0.f;
  }
  float oacc[RH];
#pragma unroll
  for (int i = 0; i < RH; ++i) oacc[i] = 0.f;
  const int pc = tid & 127, prh = tid >> 7;  // PV mapping: column pc, rows prh*RH..
  const int ph = pc >> 5;                    // head of that column

  for (int k0 = 0; k0 < n_key; k0 += TK) {
    __syncthreads();  // previous tile fully consumed (also orders the m/l init)
    // ---- stage K|V tile -----------------------------------------------------------------------------------
    for (int i = tid; i < TK * 64; i += NT) {
      const int key = i >> 6, c4 = i & 63;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + key < n_key) v = __ldg(reinterpret_cast<const float4*>(kv + (size_t)(k0 + key) * 256) + c4);
      if (c4 < 32)
        *reinterpret_cast<float4*>(sm.ks + key * KPAD + c4 * 4) = v;
      else
        *reinterpret_cast<float4*>(sm.vs + key * D + (c4 - 32) * 4) = v;
    }
    if (tid < TK) sm.kvalid[tid] = (k0 + tid < n_key) ? key_valid[k0 + tid] : (uint8_t)0;
    __syncthreads();
    // ---- logits: thread (key j, head h) against all R rows ---------------------------------------------------
    {
      const int j = tid & 63, h = tid >> 6;
      float kr[DH];
#pragma unroll
      for (int d4 = 0; d4 < DH / 4; ++d4) {
        const float4 v = *reinterpret_cast<const float4*>(sm.ks + j * KPAD + h * DH + d4 * 4);
        kr[d4 * 4 + 0] = v.x;
        kr[d4 * 4 + 1] = v.y;
        kr[d4 * 4 + 2] = v.z;
        kr[d4 * 4 + 3] = v.w;
      }
      const bool kval = sm.kvalid[j] != 0;
      const int kidx = k0 + j;
#pragma unroll 2
      for (int r = 0; r < R; ++r) {
        const float* qr = sm.q + r * D + h * DH;
        float acc = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < DH / 4; ++d4) {
          const float4 qv = *reinterpret_cast<const float4*>(qr + d4 * 4);
          acc = fmaf(qv.w, kr[d4 * 4 + 3],
                     fmaf(qv.z, kr[d4 * 4 + 2], fmaf(qv.y, kr[d4 * 4 + 1], fmaf(qv.x, kr[d4 * 4 + 0], acc))));
        }
        const bool on = kval && !(self_base >= 0 && kidx == self_base + r);
        sm.s[(r * NHEAD + h) * SPAD + j] = on ? acc * scale : -INFINITY;
      }
    }
    __syncthreads();
    // ---- online softmax: one warp per (row, head) ---------------------------------------------------------------
    for (int rh = warp; rh < R * NHEAD; rh += NWARP) {
      float* sp = sm.s + rh * SPAD;
      const float s0 = sp[lane], s1 = sp[lane + 32];
      const float m_old = sm.m[rh];
      const float m_new = fmaxf(m_old, warp_max(fmaxf(s0, s1)));
      float p0 = 0.f, p1 = 0.f, a = 1.f;
      if (m_new != -INFINITY) {
        p0 = (s0 == -INFINITY) ? 0.f : expf(s0 - m_new);
        p1 = (s1 == -INFINITY) ? 0.f : expf(s1 - m_new);
        a = (m_old == -INFINITY) ? 0.f : expf(m_old - m_new);
      }
      sp[lane] = p0;
      sp[lane + 32] = p1;
      const float sum = warp_sum(p0 + p1);
      if (lane == 0) {
        sm.m[rh] = m_new;
        sm.l[rh] = sm.l[rh] * a + sum;
        sm.alpha[rh] = a;
      }
    }
    __syncthreads();
    // ---- O = O*alpha + P V : thread (column pc, row half prh) ----------------------------------------------------
    {
#pragma unroll
      for (int i = 0; i < RH; ++i) oacc[i] *= sm.alpha[(prh * RH + i) * NHEAD + ph];
      for (int j4 = 0; j4 < TK / 4; ++j4) {
        const float v0 = sm.vs[(j4 * 4 + 0) * D + pc], v1 = sm.vs[(j4 * 4 + 1) * D + pc];
        const float v2 = sm.vs[(j4 * 4 + 2) * D + pc], v3 = sm.vs[(j4 * 4 + 3) * D + pc];
#pragma unroll
        for (int i = 0; i < RH; ++i) {
          const float4 p = *reinterpret_cast<const float4*>(sm.s + ((prh * RH + i) * NHEAD + ph) * SPAD + j4 * 4);
          oacc[i] = fmaf(p.w, v3, fmaf(p.z, v2, fmaf(p.y, v1, fmaf(p.x, v0, oacc[i]))));
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RH; ++i) {
    const int r = prh * RH + i;
    const float l = sm.l[r * NHEAD + ph];
    sm.t[r * D + pc] = (l > 0.f) ? oacc[i] / l : 0.f;
  }
  if (tid < R) sm.dead[tid] = (sm.l[tid * NHEAD] > 0.f) ? 0 : 1;  // the key mask is the same for all heads
  __syncthreads();
}

// One pre-LN cross-attention layer on the tile (transformer.py:186-237).  In/out: sm.x.  `lw` = packed layer.
//   s2 = LN1(x); q = s2 Wq + bq; a = MHA(q, K, V); x += (dead ? 0 : a Wo + bo);
//   x += W2 relu(W1 LN2(x) + b1) + b2; x[~row_valid] = 0
template <int R>
__device__ void xlayer_tile(TileSmem<R>& sm, const float* __restrict__ lw, const float* __restrict__ kv,
                            const uint8_t* __restrict__ key_valid, int n_key, int self_base) {
  constexpr int RPT = R / 4;
  layernorm_rows(sm.x, D, sm.t, D, R, lw + tfl::NORM1_W, lw + tfl::NORM1_B);
  __syncthreads();
  gemm128<RPT>(lw + tfl::IN_W, 3 * D, 0, D / 4, sm.t, D,
               [&](int r, int c, float v) { sm.q[r * D + c] = v + __ldg(lw + tfl::IN_B + c); });
  __syncthreads();
  attention_tile<R>(sm, kv, key_valid, n_key, self_base);  // -> sm.t, sm.dead   (ends with a barrier)
  gemm128<RPT>(lw + tfl::OUT_W, D, 0, D / 4, sm.t, D, [&](int r, int c, float v) {
    if (!sm.dead[r]) sm.x[r * D + c] += v + __ldg(lw + tfl::OUT_B + c);
  });
  __syncthreads();
  layernorm_rows(sm.x, D, sm.t, D, R, lw + tfl::NORM2_W, lw + tfl::NORM2_B);
  __syncthreads();
  gemm128<RPT>(lw + tfl::L1_W, D, 0, D / 4, sm.t, D,
               [&](int r, int c, float v) { sm.q[r * D + c] = fmaxf(v + __ldg(lw + tfl::L1_B + c), 0.f); });
  __syncthreads();
  gemm128<RPT>(lw + tfl::L2_W, D, 0, D / 4, sm.q, D, [&](int r, int c, float v) {
    const float y = sm.x[r * D + c] + v + __ldg(lw + tfl::L2_B + c);
    sm.x[r * D + c] = sm.row_valid[r] ? y : 0.f;
  });
  __syncthreads();
}

// K|V projection of the tile rows in sm.x for one layer: kv_out[row][0:256] = LN_tgt(x) Wkv + bkv.
// Uses sm.t as scratch.  `kv_out` points at the first row of the tile; rows >= nrow are not written.
template <int R>
__device__ void kv_project_tile(TileSmem<R>& sm, const float* __restrict__ lw, float* __restrict__ kv_out, int nrow) {
  constexpr int RPT = R / 4;
  layernorm_rows(sm.x, D, sm.t, D, R, lw + tfl::NORMT_W, lw + tfl::NORMT_B);
  __syncthreads();
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    gemm128<RPT>(lw + tfl::IN_W, 3 * D, D + half * D, D / 4, sm.t, D, [&](int r, int c, float v) {
      if (r < nrow) kv_out[(size_t)r * 256 + half * D + c] = v + __ldg(lw + tfl::IN_B + D + half * D + c);
    });
  }
  __syncthreads();
}

}  // namespace tb
