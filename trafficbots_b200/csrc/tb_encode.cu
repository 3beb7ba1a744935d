// Scene encoding kernels: parameter re-layout, polyline encoder, input-PE encoders, K|V projection and the
// stand-alone cross-attention layer.  See include/trafficbots_b200.h for the reference methods each replaces.
#define TB_WEIGHT_TABLE_IMPL
#include "tb_host.h"
#include <stdlib.h>

namespace tb {

// ------------------------------------------------------------------------------------------------------------
// parameter re-layout: W[N,K] -> Wt4[ceil(K/4)][N][4]; 1-D tensors are copied
// ------------------------------------------------------------------------------------------------------------
__global__ void k_pack_weight(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (cols == 0) {
    const int n4 = (rows + 3) / 4 * 4;
    if (i < n4) dst[i] = i < rows ? src[i] : 0.f;
    return;
  }
  const int k4n = (cols + 3) / 4;
  if (i >= k4n * rows * 4) return;
  const int j = i & 3, n = (i >> 2) % rows, k4 = (i >> 2) / rows;
  const int k = k4 * 4 + j;
  dst[i] = k < cols ? src[(size_t)n * cols + k] : 0.f;
}

// tensor-core copy of W[N,K] (N, K multiples of 128): 128x128 blocks, each [hi kb0 | hi kb1 | lo kb0 | lo kb1] x 16 KB,
// rows = output features (the UMMA B operand is N x K, K-major), 128-byte swizzle (tb_tc.cuh)
__global__ void k_pack_weight_tc(const float* __restrict__ src, unsigned char* __restrict__ dst, int rows, int cols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk (8 bf16) of either the hi or the lo plane
  const int chunks_per_plane = rows * cols / 8;
  if (i >= 2 * chunks_per_plane) return;
  const int plane = i / chunks_per_plane, j = i % chunks_per_plane;
  const int n = j / (cols / 8), k8 = j % (cols / 8);  // row n, elements k8*8 .. +7
  const int nb = n >> 7, r = n & 127, kblk128 = (k8 * 8) >> 7, kin = (k8 * 8) & 127;
  const int kb = kin >> 6, c16 = (kin & 63) >> 3;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = src[(size_t)n * cols + k8 * 8 + e];
  uint4 hi, lo;
  tc::split8(v, hi, lo);
  unsigned char* blk = dst + (size_t)(nb * (cols >> 7) + kblk128) * tc::BLOCK_BYTES;
  *reinterpret_cast<uint4*>(blk + plane * 2 * tc::KB_BYTES_128 + kb * tc::KB_BYTES_128 + tc::sw128_off(r, c16)) = plane ? lo : hi;
}

// ------------------------------------------------------------------------------------------------------------
// InputPeEncoder on a row tile (input_pe_encoder.py:41-61, pe_mode "cat"): x[r] = valid ? [MLP(attr) | PE] : 0.
// Expects attr (zero padded to 4*K4) in sm.t rows, pose in pose[r] = (x, y, yaw), validity in sm.row_valid.
// ------------------------------------------------------------------------------------------------------------
template <int R>
__device__ void input_pe_encode_tile(TileSmem<R>& sm, const float3* pose, const float* __restrict__ packed, int w1, int b1,
                                     int w2, int b2, int K4, int f_xy, int f_yaw) {
  for (int i = threadIdx.x; i < R * 48; i += NT) {
    const int r = i / 48, j = i % 48;
    pose_pe_elem(j, pose[r].x, pose[r].y, pose[r].z, packed + f_xy, packed + f_yaw, sm.x + r * D + 32);
  }
  gemm_small(packed + w1, 32, K4, sm.t, D, R,
             [&](int r, int n, float v) { sm.q[r * D + n] = fmaxf(v + __ldg(packed + b1 + n), 0.f); });
  __syncthreads();
  gemm_small(packed + w2, 32, 8, sm.q, D, R, [&](int r, int n, float v) { sm.x[r * D + n] = v + __ldg(packed + b2 + n); });
  __syncthreads();
  for (int i = threadIdx.x; i < R * D; i += NT)
    if (!sm.row_valid[i / D]) sm.x[i] = 0.f;
  __syncthreads();
}

// agent history encoder (sc_input.py:109-122 + traffic_bots.py:149): rows = [S,Th,A]
template <int R>
__global__ void __launch_bounds__(NT) k_encode_agent_hist(TbDims dm, TbSceneIn in, const float* __restrict__ packed,
                                                          float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<R>& sm = *reinterpret_cast<TileSmem<R>*>(smem_raw);
  __shared__ float3 pose[R];
  const int A = dm.n_agent;
  const long n_row = (long)dm.n_scene * dm.n_step_hist * A;
  const long row0 = (long)blockIdx.x * R;
  if (threadIdx.x < R) {
    const int r = threadIdx.x;
    const long row = row0 + r;
    float* at = sm.t + r * D;
    bool valid = false;
    float3 p = make_float3(0.f, 0.f, 0.f);
    for (int i = 0; i < 12; ++i) at[i] = 0.f;
    if (row < n_row) {
      const int a = (int)(row % A);
      const int s = (int)(row / ((long)dm.n_step_hist * A));
      valid = in.agent_valid[row] != 0;
      p = make_float3(in.agent_pos[row * 2], in.agent_pos[row * 2 + 1], in.agent_yaw[row]);
      at[0] = in.agent_vel[row * 2];
      at[1] = in.agent_vel[row * 2 + 1];
      at[2] = in.agent_spd[row];
      at[3] = in.agent_yaw_rate[row];
      at[4] = in.agent_acc[row];
      const long sa = (long)s * A + a;
      for (int i = 0; i < 3; ++i) {
        at[5 + i] = in.agent_size[sa * 3 + i];
        at[8 + i] = in.agent_type[sa * 3 + i] ? 1.f : 0.f;
      }
    }
    pose[r] = p;
    sm.row_valid[r] = valid;
  }
  __syncthreads();
  input_pe_encode_tile<R>(sm, pose, packed, tbw::model_agent_encoder_mlp_fc_layers_0_weight,
                          tbw::model_agent_encoder_mlp_fc_layers_0_bias, tbw::model_agent_encoder_mlp_fc_layers_3_weight,
                          tbw::model_agent_encoder_mlp_fc_layers_3_bias, 3, tbw::pre_processing_input_pose_pe_agent_pe_xy_freqs,
                          tbw::pre_processing_input_pose_pe_agent_pe_yaw_freqs);
  for (int i = threadIdx.x; i < R * (D / 4); i += NT) {
    const int r = i / (D / 4), c4 = i % (D / 4);
    if (row0 + r < n_row)
      reinterpret_cast<float4*>(out + (row0 + r) * D)[c4] = reinterpret_cast<const float4*>(sm.x + r * D)[c4];
  }
}

// traffic-light encoder (sc_input.py:136-139 + traffic_bots.py:150): rows = [S,Th,TL]
template <int R>
__global__ void __launch_bounds__(NT) k_encode_tl(TbDims dm, TbSceneIn in, const float* __restrict__ packed,
                                                  float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<R>& sm = *reinterpret_cast<TileSmem<R>*>(smem_raw);
  __shared__ float3 pose[R];
  const long n_row = (long)dm.n_scene * dm.n_step_hist * dm.n_tl;
  const long row0 = (long)blockIdx.x * R;
  if (threadIdx.x < R) {
    const int r = threadIdx.x;
    const long row = row0 + r;
    float* at = sm.t + r * D;
    bool valid = false;
    float3 p = make_float3(0.f, 0.f, 0.f);
    for (int i = 0; i < 8; ++i) at[i] = 0.f;
    if (row < n_row) {
      valid = in.tl_valid[row] != 0;
      p = make_float3(in.tl_pos[row * 2], in.tl_pos[row * 2 + 1], atan2f(in.tl_dir[row * 2 + 1], in.tl_dir[row * 2]));
      for (int i = 0; i < TB_TL_STATE; ++i) at[i] = in.tl_state[row * TB_TL_STATE + i] ? 1.f : 0.f;
    }
    pose[r] = p;
    sm.row_valid[r] = valid;
  }
  __syncthreads();
  input_pe_encode_tile<R>(sm, pose, packed, tbw::model_tl_encoder_mlp_fc_layers_0_weight,
                          tbw::model_tl_encoder_mlp_fc_layers_0_bias, tbw::model_tl_encoder_mlp_fc_layers_3_weight,
                          tbw::model_tl_encoder_mlp_fc_layers_3_bias, 2, tbw::pre_processing_input_pose_pe_tl_pe_xy_freqs,
                          tbw::pre_processing_input_pose_pe_tl_pe_yaw_freqs);
  for (int i = threadIdx.x; i < R * (D / 4); i += NT) {
    const int r = i / (D / 4), c4 = i % (D / 4);
    if (row0 + r < n_row)
      reinterpret_cast<float4*>(out + (row0 + r) * D)[c4] = reinterpret_cast<const float4*>(sm.x + r * D)[c4];
  }
}

// ------------------------------------------------------------------------------------------------------------
// Polyline encoder (map_encoder.py:72-106): NP polylines (NP*20 node rows) per CTA.
//   node feature = InputPeEncoder([type(11) | onehot(node)(20)], PE(pos, atan2(dir)))      (sc_input.py:124-134)
//   3 pre-LN layers over the 20 nodes of a polyline, tgt = the INITIAL node features (densetnt_vectornet)
//   masked max-pool over the valid nodes, zero for polylines without a valid node
// ------------------------------------------------------------------------------------------------------------
constexpr int MAP_NP = 2;
constexpr int MAP_R = MAP_NP * TB_PL_NODE;  // 40 rows
struct MapSmem {
  float x[MAP_R * D];
  float x0[MAP_R * D];
  float t[MAP_R * D];
  float q[MAP_R * D];
  float k[MAP_R * D];
  float v[MAP_R * D];
  float3 pose[MAP_R];
  uint8_t row_valid[MAP_R];
  uint8_t pl_valid[MAP_NP];
};

__global__ void __launch_bounds__(NT) k_map_polyline(TbDims dm, TbSceneIn in, const float* __restrict__ packed,
                                                     float* __restrict__ pl_feature, uint8_t* __restrict__ pl_valid_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MapSmem& sm = *reinterpret_cast<MapSmem*>(smem_raw);
  constexpr int R = MAP_R, RPT = MAP_R / 4, N = TB_PL_NODE;
  const int tid = threadIdx.x;
  const long n_pl_total = (long)dm.n_scene * dm.n_pl;
  const long pl0 = (long)blockIdx.x * MAP_NP;

  // ---- node attributes + pose --------------------------------------------------------------------------------
  if (tid < R) {
    const int r = tid, p = r / N, n = r % N;
    const long pl = pl0 + p;
    float* at = sm.t + r * D;
    for (int i = 0; i < 32; ++i) at[i] = 0.f;
    bool valid = false;
    float3 ps = make_float3(0.f, 0.f, 0.f);
    if (pl < n_pl_total) {
      const long node = pl * N + n;
      valid = in.map_valid[node] != 0;
      for (int i = 0; i < TB_PL_TYPE; ++i) at[i] = in.map_type[pl * TB_PL_TYPE + i] ? 1.f : 0.f;
      at[TB_PL_TYPE + n] = 1.f;  // pl_node_ohe = eye(20) (sc_input.py:60)
      ps = make_float3(in.map_pos[node * 2], in.map_pos[node * 2 + 1], atan2f(in.map_dir[node * 2 + 1], in.map_dir[node * 2]));
    }
    sm.pose[r] = ps;
    sm.row_valid[r] = valid;
  }
  __syncthreads();
  if (tid < MAP_NP) {
    bool any = false;
    for (int n = 0; n < N; ++n) any |= sm.row_valid[tid * N + n] != 0;
    sm.pl_valid[tid] = any;
  }
  for (int i = tid; i < R * 48; i += NT) {
    const int r = i / 48, j = i % 48;
    pose_pe_elem(j, sm.pose[r].x, sm.pose[r].y, sm.pose[r].z, packed + tbw::pre_processing_input_pose_pe_map_pe_xy_freqs,
                 packed + tbw::pre_processing_input_pose_pe_map_pe_yaw_freqs, sm.x + r * D + 32);
  }
  gemm_small(packed + tbw::model_map_encoder_input_pe_encoder_mlp_fc_layers_0_weight, 32, 8, sm.t, D, R, [&](int r, int n, float v) {
    sm.q[r * D + n] = fmaxf(v + __ldg(packed + tbw::model_map_encoder_input_pe_encoder_mlp_fc_layers_0_bias + n), 0.f);
  });
  __syncthreads();
  gemm_small(packed + tbw::model_map_encoder_input_pe_encoder_mlp_fc_layers_3_weight, 32, 8, sm.q, D, R, [&](int r, int n, float v) {
    sm.x[r * D + n] = v + __ldg(packed + tbw::model_map_encoder_input_pe_encoder_mlp_fc_layers_3_bias + n);
  });
  __syncthreads();
  for (int i = tid; i < R * D; i += NT) {
    const float v = sm.row_valid[i / D] ? sm.x[i] : 0.f;
    sm.x[i] = v;
    sm.x0[i] = v;
  }
  __syncthreads();

  // ---- 3 layers ------------------------------------------------------------------------------------------------
#pragma unroll 1
  for (int L = 0; L < 3; ++L) {
    const float* lw = packed + tbw::model_map_encoder_transformer_densetnt_layers_0_norm1_weight + L * tfl::STRIDE;
    // K|V from the initial features
    layernorm_rows(sm.x0, D, sm.t, D, R, lw + tfl::NORMT_W, lw + tfl::NORMT_B);
    __syncthreads();
    gemm128<RPT>(lw + tfl::IN_W, 3 * D, D, D / 4, sm.t, D,
                 [&](int r, int c, float v) { sm.k[r * D + c] = v + __ldg(lw + tfl::IN_B + D + c); });
    gemm128<RPT>(lw + tfl::IN_W, 3 * D, 2 * D, D / 4, sm.t, D,
                 [&](int r, int c, float v) { sm.v[r * D + c] = v + __ldg(lw + tfl::IN_B + 2 * D + c); });
    __syncthreads();
    layernorm_rows(sm.x, D, sm.t, D, R, lw + tfl::NORM1_W, lw + tfl::NORM1_B);
    __syncthreads();
    gemm128<RPT>(lw + tfl::IN_W, 3 * D, 0, D / 4, sm.t, D,
                 [&](int r, int c, float v) { sm.q[r * D + c] = v + __ldg(lw + tfl::IN_B + c); });
    __syncthreads();
    // attention inside each polyline: one thread per (polyline, head, query node)
    if (tid < MAP_NP * NHEAD * N) {
      const int i = tid % N, h = (tid / N) % NHEAD, p = tid / (N * NHEAD);
      const int row = p * N + i;
      float qv[DH];
#pragma unroll
      for (int d = 0; d < DH; ++d) qv[d] = sm.q[row * D + h * DH + d];
      float lg[N];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float* kr = sm.k + (p * N + j) * D + h * DH;
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) acc = fmaf(qv[d], kr[d], acc);
        lg[j] = sm.row_valid[p * N + j] ? acc * 0.17677669529663687f : -INFINITY;
        mx = fmaxf(mx, lg[j]);
      }
      float o[DH];
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = 0.f;
      float sum = 0.f;
      if (mx != -INFINITY) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const float pj = (lg[j] == -INFINITY) ? 0.f : expf(lg[j] - mx);
          sum += pj;
          const float* vr = sm.v + (p * N + j) * D + h * DH;
#pragma unroll
          for (int d = 0; d < DH; ++d) o[d] = fmaf(pj, vr[d], o[d]);
        }
      }
      const float inv = sum > 0.f ? 1.0f / sum : 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) sm.t[row * D + h * DH + d] = o[d] * inv;
    }
    __syncthreads();
    gemm128<RPT>(lw + tfl::OUT_W, D, 0, D / 4, sm.t, D, [&](int r, int c, float v) {
      if (sm.pl_valid[r / N]) sm.x[r * D + c] += v + __ldg(lw + tfl::OUT_B + c);  // dead rows: polyline without valid node
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
  // ---- masked max-pool (map_encoder.py:95-97,105-106) ---------------------------------------------------------
  for (int i = tid; i < MAP_NP * D; i += NT) {
    const int p = i / D, c = i % D;
    if (pl0 + p >= n_pl_total) continue;
    float mx = -INFINITY;
    for (int n = 0; n < N; ++n)
      if (sm.row_valid[p * N + n]) mx = fmaxf(mx, sm.x[(p * N + n) * D + c]);
    pl_feature[(pl0 + p) * D + c] = sm.pl_valid[p] ? mx : 0.f;
  }
  if (tid < MAP_NP && pl0 + tid < n_pl_total) pl_valid_out[pl0 + tid] = sm.pl_valid[tid];
}

// ------------------------------------------------------------------------------------------------------------
// stand-alone K|V projection and cross-attention layer (building blocks of the C ABI)
// ------------------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(NT) k_kv_project(const float* __restrict__ tgt, long n_row, const float* __restrict__ lw,
                                                   float* __restrict__ kv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<R>& sm = *reinterpret_cast<TileSmem<R>*>(smem_raw);
  const long row0 = (long)blockIdx.x * R;
  for (int i = threadIdx.x; i < R * (D / 4); i += NT) {
    const int r = i / (D / 4), c4 = i % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < n_row) v = __ldg(reinterpret_cast<const float4*>(tgt + (row0 + r) * D) + c4);
    reinterpret_cast<float4*>(sm.x + r * D)[c4] = v;
  }
  __syncthreads();
  const long left = n_row - row0;
  kv_project_tile<R>(sm, lw, kv + row0 * 256, left < R ? (int)left : R);
}

template <int R>
__global__ void __launch_bounds__(NT) k_xlayer(const float* __restrict__ src, const uint8_t* __restrict__ src_valid, int n_src,
                                               const float* __restrict__ kv, const uint8_t* __restrict__ key_valid, int n_key,
                                               int kv_share, int mask_self, const float* __restrict__ lw, float* __restrict__ dst) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<R>& sm = *reinterpret_cast<TileSmem<R>*>(smem_raw);
  const int b = blockIdx.y, r0 = blockIdx.x * R;
  const int kb = b / kv_share;
  for (int i = threadIdx.x; i < R * (D / 4); i += NT) {
    const int r = i / (D / 4), c4 = i % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < n_src) v = __ldg(reinterpret_cast<const float4*>(src + ((size_t)b * n_src + r0 + r) * D) + c4);
    reinterpret_cast<float4*>(sm.x + r * D)[c4] = v;
  }
  if (threadIdx.x < R) sm.row_valid[threadIdx.x] = (r0 + threadIdx.x < n_src) ? src_valid[(size_t)b * n_src + r0 + threadIdx.x] : (uint8_t)0;
  __syncthreads();
  xlayer_tile<R>(sm, lw, kv + (size_t)kb * n_key * 256, key_valid + (size_t)kb * n_key, n_key, mask_self ? r0 : -1);
  for (int i = threadIdx.x; i < R * (D / 4); i += NT) {
    const int r = i / (D / 4), c4 = i % (D / 4);
    if (r0 + r < n_src)
      reinterpret_cast<float4*>(dst + ((size_t)b * n_src + r0 + r) * D)[c4] = reinterpret_cast<const float4*>(sm.x + r * D)[c4];
  }
}

}  // namespace tb

// ==============================================================================================================
// host side
// ==============================================================================================================
using namespace tb;

extern "C" int32_t tb_weight_count(void) { return TB_N_WEIGHTS; }
extern "C" const char* tb_weight_name(int32_t i) { return (i >= 0 && i < TB_N_WEIGHTS) ? TB_WEIGHT_TABLE[i].name : nullptr; }
extern "C" int32_t tb_weight_rows(int32_t i) { return (i >= 0 && i < TB_N_WEIGHTS) ? TB_WEIGHT_TABLE[i].rows : -1; }
extern "C" int32_t tb_weight_cols(int32_t i) { return (i >= 0 && i < TB_N_WEIGHTS) ? TB_WEIGHT_TABLE[i].cols : -1; }
extern "C" int32_t tb_tc_block_count(void) { return TB_N_TC_BLOCKS; }
extern "C" int32_t tb_tc_first_block(int32_t i) {
  for (int j = 0; j < TB_N_TC_WEIGHTS; ++j)
    if (TB_TC_TABLE[j].weight == i) return TB_TC_TABLE[j].first_block;
  return -1;
}
extern "C" size_t tb_packed_weight_bytes(void) { return tc_blob_offset_bytes() + (size_t)TB_N_TC_BLOCKS * tc::BLOCK_BYTES; }

extern "C" int32_t tb_pack_weights(const float* const* params, float* packed, void* stream_) {
  if (!params || !packed) return TB_ERR_NULL;
  if (!aligned16(packed)) return TB_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)stream_;
  for (int i = 0; i < TB_N_WEIGHTS; ++i) {
    const TbWeightDesc& d = TB_WEIGHT_TABLE[i];
    if (!params[i]) return TB_ERR_NULL;
    const int n = d.cols == 0 ? (d.rows + 3) / 4 * 4 : (d.cols + 3) / 4 * d.rows * 4;
    k_pack_weight<<<(n + 255) / 256, 256, 0, st>>>(params[i], packed + d.offset, d.rows, d.cols);
    count_launch();
  }
  // tensor-core copies: bf16 hi / lo blocks in the swizzled UMMA layout
  unsigned char* tcb = reinterpret_cast<unsigned char*>(packed) + tc_blob_offset_bytes();
  for (int j = 0; j < TB_N_TC_WEIGHTS; ++j) {
    const TbWeightDesc& d = TB_WEIGHT_TABLE[TB_TC_TABLE[j].weight];
    const int chunks = d.rows * d.cols / 8 * 2;  // 16-byte chunks, hi and lo
    k_pack_weight_tc<<<(chunks + 255) / 256, 256, 0, st>>>(params[TB_TC_TABLE[j].weight],
                                                             tcb + (size_t)TB_TC_TABLE[j].first_block * tc::BLOCK_BYTES, d.rows, d.cols);
    count_launch();
  }
  return launch_status();
}

static int check_dims(const TbDims* d) {
  if (!d) return TB_ERR_NULL;
  if (d->n_scene < 1 || d->n_mode < 1 || d->n_agent < 1 || d->n_pl < 1 || d->n_tl < 1 || d->n_step_hist < 1 ||
      d->n_step_gt < 1 || d->n_step < 1)
    return TB_ERR_BAD_SHAPE;
  if ((long)d->n_scene * d->n_mode > 65535) return TB_ERR_BAD_SHAPE;  // gridDim.y
  if (d->n_cta_per_mode != 0 && d->n_cta_per_mode != 1 && d->n_cta_per_mode != 2 && d->n_cta_per_mode != 4) return TB_ERR_BAD_SHAPE;
  return TB_OK;
}
int tb::check_dims_host(const TbDims* d) { return check_dims(d); }

extern "C" size_t tb_encode_workspace_bytes(const TbDims* d) {
  if (check_dims(d) != TB_OK) return 0;
  const size_t rows = (size_t)d->n_scene * d->n_pl;
  // pooled polyline features + self-attention K|V + per-CTA scratch of the tensor-core polyline encoder + the tensor-core
  // key blocks and key counts of the map self-attention layer
  const size_t nT = (d->n_pl + 63) / 64;
  return rows * D * sizeof(float) + rows * 256 * sizeof(float) + map_tc_scratch_bytes(MAP_TC_MAX_CTA) + 1024 +
         (size_t)d->n_scene * nT * tc::BLOCK_BYTES + (((size_t)d->n_scene * sizeof(int32_t) + 255) & ~(size_t)255) + 256 +
         map_plan_bytes((long)rows);
}

template <int R>
static int launch_kv_project(const float* tgt, long n_row, const float* lw, float* kv, cudaStream_t st) {
  static std::atomic<uint64_t> attr_set{0};
  if (!smem_attr_done(attr_set)) {
    if (!set_max_smem(k_kv_project<R>, (int)sizeof(TileSmem<R>))) return TB_ERR_LAUNCH;
    smem_attr_mark(attr_set);
  }
  k_kv_project<R><<<(unsigned)((n_row + R - 1) / R), NT, sizeof(TileSmem<R>), st>>>(tgt, n_row, lw, kv);
  count_launch();
  return launch_status();
}

extern "C" int32_t tb_kv_project(int32_t block, int32_t layer, const float* tgt, int64_t n_row, const float* packed,
                                 float* kv, void* stream) {
  if (!tgt || !packed || !kv) return TB_ERR_NULL;
  if (block_base(block) < 0 || layer < 0 || layer >= block_layers(block) || n_row < 1) return TB_ERR_BAD_SHAPE;
  if (!aligned16(tgt) || !aligned16(packed) || !aligned16(kv)) return TB_ERR_ALIGN;
  if (tc_enabled() && n_row >= 128)  // tensor-core kernel for everything but tiny inputs
    return launch_kv_project_tc(block, layer, tgt, n_row, packed, kv, (cudaStream_t)stream);
  return launch_kv_project<ROW_TILE>(tgt, n_row, packed + block_base(block) + layer * tfl::STRIDE, kv, (cudaStream_t)stream);
}

extern "C" int32_t tb_xlayer(int32_t block, int32_t layer, const float* src, const uint8_t* src_valid, int32_t n_batch,
                             int32_t n_src, const float* kv, const uint8_t* key_valid, int32_t n_key, int32_t kv_share,
                             int32_t mask_self, const float* packed, float* dst, void* stream) {
  if (!src || !src_valid || !kv || !key_valid || !packed || !dst) return TB_ERR_NULL;
  if (block_base(block) < 0 || layer < 0 || layer >= block_layers(block) || n_batch < 1 || n_batch > 65535 || n_src < 1 ||
      n_key < 1 || kv_share < 1 || n_batch % kv_share != 0)
    return TB_ERR_BAD_SHAPE;
  if (!aligned16(src) || !aligned16(kv) || !aligned16(packed) || !aligned16(dst)) return TB_ERR_ALIGN;
  constexpr int R = ROW_TILE;
  static std::atomic<uint64_t> attr_set{0};
  if (!smem_attr_done(attr_set)) {
    if (!set_max_smem(k_xlayer<R>, (int)sizeof(TileSmem<R>))) return TB_ERR_LAUNCH;
    smem_attr_mark(attr_set);
  }
  dim3 grid((n_src + R - 1) / R, n_batch);
  k_xlayer<R><<<grid, NT, sizeof(TileSmem<R>), (cudaStream_t)stream>>>(src, src_valid, n_src, kv, key_valid, n_key, kv_share,
                                                                         mask_self, packed + block_base(block) + layer * tfl::STRIDE,
                                                                         dst);
  count_launch();
  return launch_status();
}

// the compacted-tile plan of the polyline encoder lives behind the key blocks / key counts of the map self-attention
static int32_t* map_plan_ws(const TbDims& d, void* workspace) {
  const size_t rows = (size_t)d.n_scene * d.n_pl, nT = (d.n_pl + 63) / 64;
  uintptr_t p = reinterpret_cast<uintptr_t>(workspace) + rows * D * sizeof(float) + rows * 256 * sizeof(float) +
                map_tc_scratch_bytes(MAP_TC_MAX_CTA);
  p = (p + 1023) & ~(uintptr_t)1023;
  p += (size_t)d.n_scene * nT * tc::BLOCK_BYTES + (((size_t)d.n_scene * sizeof(int32_t) + 255) & ~(size_t)255);
  return reinterpret_cast<int32_t*>((p + 255) & ~(uintptr_t)255);
}

extern "C" int32_t tb_encode_scene(const TbDims* dims, const TbSceneIn* in, const float* packed, const TbSceneOut* out,
                                   void* workspace, void* stream) {
  int rc = check_dims(dims);
  if (rc != TB_OK) return rc;
  if (!in || !packed || !out || !workspace) return TB_ERR_NULL;
  // in->map_valid == NULL: agents / traffic lights only (the map part of the scene was encoded by an earlier call, e.g. the
  // posterior pass over the full episode re-uses the map features and K|V caches of the history pass, sc_latent.py:143-148)
  const bool do_map = in->map_valid != nullptr;
  const void* req[] = {in->agent_valid, in->agent_pos, in->agent_yaw, in->agent_vel, in->agent_spd, in->agent_yaw_rate, in->agent_acc,
                       in->agent_size,  in->agent_type, in->tl_valid, in->tl_state, in->tl_pos,  in->tl_dir, out->agent_feature,
                       out->tl_feature, out->kv_tl};
  for (const void* p : req)
    if (!p) return TB_ERR_NULL;
  const void* req_map[] = {in->map_type, in->map_pos, in->map_dir, out->map_feature, out->map_feature_valid, out->kv_map};
  if (do_map)
    for (const void* p : req_map)
      if (!p) return TB_ERR_NULL;
  if (!aligned16(packed) || !aligned16(workspace) || !aligned16(out->agent_feature) || !aligned16(out->tl_feature) || !aligned16(out->kv_tl))
    return TB_ERR_ALIGN;
  if (do_map && (!aligned16(out->map_feature) || !aligned16(out->kv_map))) return TB_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const TbDims d = *dims;
  constexpr int R = ROW_TILE;
  const long n_pl = (long)d.n_scene * d.n_pl;
  float* pl_feature = reinterpret_cast<float*>(workspace);
  float* kv_self = pl_feature + n_pl * D;

  static std::atomic<uint64_t> attr_set{0};
  if (!smem_attr_done(attr_set)) {
    if (!set_max_smem(k_map_polyline, (int)sizeof(MapSmem))) return TB_ERR_LAUNCH;
    if (!set_max_smem(k_encode_agent_hist<R>, (int)sizeof(TileSmem<R>))) return TB_ERR_LAUNCH;
    if (!set_max_smem(k_encode_tl<R>, (int)sizeof(TileSmem<R>))) return TB_ERR_LAUNCH;
    smem_attr_mark(attr_set);
  }
  if (do_map) {
  // 1. polyline encoder: tcgen05 kernel; TB_DISABLE_TC=1 selects the fp32 CUDA-core kernel (verification aid)
  if (tc_enabled()) {
    rc = launch_map_polyline_tc2(d, *in, packed, kv_self + n_pl * 256, MAP_TC_MAX_CTA, pl_feature, out->map_feature_valid,
                                 map_plan_ws(d, workspace), st);
    if (rc != TB_OK) return rc;
  } else {
    k_map_polyline<<<(unsigned)((n_pl + MAP_NP - 1) / MAP_NP), NT, sizeof(MapSmem), st>>>(d, *in, packed, pl_feature,
                                                                                        out->map_feature_valid);
    count_launch();
  }
  // 2. global self-attention over the polylines of a scene (map_encoder.py:108-114)
  rc = tb_kv_project(TB_BLOCK_MAP_SELF_ATTN, 0, pl_feature, n_pl, packed, kv_self, stream);
  if (rc != TB_OK) return rc;
  if (tc_enabled()) {  // tensor-core layer on compacted key blocks (tb_tc_xlayer.cu)
    unsigned char* ws_end = reinterpret_cast<unsigned char*>(kv_self + n_pl * 256) + map_tc_scratch_bytes(MAP_TC_MAX_CTA);
    unsigned char* self_blocks = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ws_end) + 1023) & ~(uintptr_t)1023);
    const size_t nT = (d.n_pl + 63) / 64;
    int32_t* n_key_self = reinterpret_cast<int32_t*>(self_blocks + (size_t)d.n_scene * nT * tc::BLOCK_BYTES);
    rc = launch_pack_kv_tc(kv_self, out->map_feature_valid, d.n_scene, d.n_scene, d.n_pl, self_blocks, n_key_self, st);
    if (rc != TB_OK) return rc;
    rc = launch_xlayer_tc(TB_BLOCK_MAP_SELF_ATTN, 0, pl_feature, out->map_feature_valid, d.n_scene, d.n_pl, self_blocks, n_key_self,
                          d.n_pl, 1, packed, out->map_feature, st);
  } else {
    rc = tb_xlayer(TB_BLOCK_MAP_SELF_ATTN, 0, pl_feature, out->map_feature_valid, d.n_scene, d.n_pl, kv_self,
                   out->map_feature_valid, d.n_pl, 1, 0, packed, out->map_feature, stream);
  }
  if (rc != TB_OK) return rc;
  // 3. loop-invariant K|V of the policy's agent->map layers
  for (int L = 0; L < 3; ++L) {
    rc = tb_kv_project(TB_BLOCK_AS2PL, L, out->map_feature, n_pl, packed, out->kv_map + (size_t)L * n_pl * 256, stream);
    if (rc != TB_OK) return rc;
  }
  if (out->kv_map_tc && out->n_key_map) {  // tensor-core operand blocks of the same K|V (valid keys only)
    rc = launch_pack_kv_tc(out->kv_map, out->map_feature_valid, 3 * d.n_scene, d.n_scene, d.n_pl, out->kv_map_tc, out->n_key_map, st);
    if (rc != TB_OK) return rc;
  }
  }  // do_map
  // 4. agent / traffic-light history encoders
  const long n_ag = (long)d.n_scene * d.n_step_hist * d.n_agent;
  k_encode_agent_hist<R><<<(unsigned)((n_ag + R - 1) / R), NT, sizeof(TileSmem<R>), st>>>(d, *in, packed, out->agent_feature);
  count_launch();
  const long n_tl = (long)d.n_scene * d.n_step_hist * d.n_tl;
  k_encode_tl<R><<<(unsigned)((n_tl + R - 1) / R), NT, sizeof(TileSmem<R>), st>>>(d, *in, packed, out->tl_feature);
  count_launch();
  // 5. K|V of the agent->traffic-light layers for every history frame
  for (int L = 0; L < 3; ++L) {
    rc = tb_kv_project(TB_BLOCK_AS2TL, L, out->tl_feature, n_tl, packed, out->kv_tl + (size_t)L * n_tl * 256, stream);
    if (rc != TB_OK) return rc;
  }
  if (out->kv_tl_tc && out->n_key_tl) {
    rc = launch_pack_kv_tc(out->kv_tl, in->tl_valid, 3 * d.n_scene * d.n_step_hist, d.n_scene * d.n_step_hist, d.n_tl, out->kv_tl_tc,
                           out->n_key_tl, st);
    if (rc != TB_OK) return rc;
  }
  return launch_status();
}
