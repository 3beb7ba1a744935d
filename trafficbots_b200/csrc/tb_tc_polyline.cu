// Polyline encoder on the tensor pipe, second version (map_encoder.py:72-106): 512 threads, FOUR threads per node row.
// The first version (tb_tc_kernels.cu, one thread per row, 4 warps) was bound by dependent-issue latency with one warp per
// scheduler (issue active 17 %, tensor pipe 8 %: profiles/r1d_k_map_polyline_tc.txt).  Here thread (row, q4) owns the 32
// columns [32 q4, 32 q4 + 32) of its row -- which is exactly attention head q4 -- so 16 warps share the LayerNorm / epilogue
// work and each thread runs the 20-key attention of ONE head:
//   tile = 6 polylines x 20 nodes (120 rows), persistent over tiles; residual stream in registers (32 per thread);
//   TMEM: Q [0,128) (also Wo / W1 / W2 accumulator) | K [128,256) | V [256,384) | A operand [384,512) (bf16x2 hi | lo);
//   every Linear = bf16x3 tcgen05 GEMM with the A operand in tensor memory, weights through a 2 x 64 KB bulk-copy ring;
//   K of all 4 heads is staged in shared memory for the logits, then V in the same buffer for the weighted sum.
//
// Compacted tiles (round 2): invalid nodes are dead weight in the reference -- as keys they are masked, as queries their rows
// are zeroed (transformer.py:236-237) and excluded from the max-pool (map_encoder.py:95-97) -- and in the data most polylines
// have fewer than 20 valid nodes.  `k_map_plan` lists the non-empty polylines with the running count of their valid nodes;
// a tile is the run of polylines whose first valid-node row falls into [108 t, 108 t + 108) of that compacted row space
// (at most 108 + 19 = 127 rows of the 128-row MMA tile), so only valid nodes occupy rows, every key of a row's polyline is
// valid (no masking), and the per-head key loop is cut at the longest polyline of the warp.  Results are bit-identical to the
// dense 6 x 20 layout: rows are independent in every GEMM, and a masked key contributes an exact zero.
#include "tb_host.h"

namespace tb {
namespace pl2 {

constexpr int THREADS = 512;
constexpr int N = TB_PL_NODE, ROWS = 128;  // MMA tile height
constexpr int FILL = 108;                  // compacted rows per tile before the last polyline's overhang (108 + 19 < 128)
constexpr int PLAN_THREADS = 1024;
constexpr int KVS = 132;                              // staging row stride (floats)
constexpr uint32_t TQ = 0, TK = 128, TV = 256, TA = 384;

struct Smem {
  unsigned char w[2][tc::BLOCK_BYTES];
  float kv[ROWS * KVS];
  float2 red[2][4][128];  // LayerNorm partials {sum, M2} [buffer][column quarter][row]
  float lp[2][12][128];   // per-layer bias / LayerNorm vectors (double-buffered by layer parity)
  uint64_t bar_w[2], bar_mma;
  uint32_t tmem_base;
  int32_t seg_pl[128];      // polyline index of every segment (= non-empty polyline) of the tile
  uint8_t seg_start[128];   // first row of the segment
  uint8_t seg_cnt[128];     // valid nodes (= rows) of the segment
  uint8_t row_seg[128];     // segment of every row
  uint8_t row_node[128];    // node index (0..19) of every row inside its polyline
  int32_t n_seg, n_rows, j0;
};

// ---- plan: non-empty polylines and the prefix sum of their valid-node counts ------------------------------------------------
// k_map_count (one thread per polyline): valid-node count; empty polylines get their (all-zero, invalid) pooled feature here.
// k_map_plan (one CTA): live_pl[j] = index of the j-th non-empty polyline, row_start[j] = valid nodes before it
// (row_start[n_live] = total), plan = {n_live, total_rows}.
__global__ void __launch_bounds__(256) k_map_count(long n_pl_total, const uint8_t* __restrict__ map_valid, uint8_t* __restrict__ counts,
                                                   float* __restrict__ pl_feature, uint8_t* __restrict__ pl_valid_out) {
  const long pl = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pl >= n_pl_total) return;
  const uint32_t* v = reinterpret_cast<const uint32_t*>(map_valid + pl * N);  // 20 bytes, 4-byte aligned
  int c = 0;
#pragma unroll
  for (int i = 0; i < N / 4; ++i) {
    const uint32_t w = v[i];
    c += ((w & 0xffu) != 0) + ((w & 0xff00u) != 0) + ((w & 0xff0000u) != 0) + ((w & 0xff000000u) != 0);
  }
  counts[pl] = (uint8_t)c;
  if (c == 0) {
    pl_valid_out[pl] = 0;
    float4* f = reinterpret_cast<float4*>(pl_feature + pl * 128);
#pragma unroll 8
    for (int i = 0; i < 32; ++i) f[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void __launch_bounds__(PLAN_THREADS) k_map_plan(long n_pl_total, const uint8_t* __restrict__ counts, int32_t* __restrict__ live_pl,
                                                           int32_t* __restrict__ row_start, int32_t* __restrict__ plan) {
  __shared__ int w_live[PLAN_THREADS / 32], w_rows[PLAN_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long chunk = (n_pl_total + PLAN_THREADS - 1) / PLAN_THREADS;
  const long p0 = (long)tid * chunk, p1 = p0 + chunk < n_pl_total ? p0 + chunk : n_pl_total;
  int my_live = 0, my_rows = 0;
  for (long pl = p0; pl < p1; ++pl) {
    const int c = counts[pl];
    my_live += c > 0;
    my_rows += c;
  }
  // block-wide exclusive scan of (live, rows)
  int sl = my_live, sr = my_rows;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, sl, o), b = __shfl_up_sync(0xffffffffu, sr, o);
    if (lane >= o) sl += a, sr += b;
  }
  if (lane == 31) w_live[warp] = sl, w_rows[warp] = sr;
  __syncthreads();
  int base_l = 0, base_r = 0;
  for (int k = 0; k < warp; ++k) base_l += w_live[k], base_r += w_rows[k];
  int j = base_l + sl - my_live, row = base_r + sr - my_rows;
  for (long pl = p0; pl < p1; ++pl) {
    const int c = counts[pl];
    if (c > 0) {
      live_pl[j] = (int32_t)pl;
      row_start[j] = row;
      ++j;
      row += c;
    }
  }
  if (tid == PLAN_THREADS - 1) {
    plan[0] = j;
    plan[1] = row;
    row_start[j] = row;
  }
}

__device__ __forceinline__ int stage_block(uint32_t s) {  // stage s of the repeating 18-stage weight schedule
  const int L = (s % 18) / 6, j = s % 6;
  const int blk[6] = {1, 2, 0, 3, 4, 5};  // order inside a layer: Wk, Wv, Wq, Wo, W1, W2 (blocks: q,k,v,out,linear1,linear2)
  return tbb::model_map_encoder_transformer_densetnt_layers_0_attn_in_proj_weight + L * 6 + blk[j];
}

__global__ void __launch_bounds__(THREADS, 1) k_map_polyline_tc2(TbDims dm, TbSceneIn in, const float* __restrict__ packed,
                                                                 const unsigned char* __restrict__ tcw, float* __restrict__ x0_scratch,
                                                                 float* __restrict__ pl_feature, uint8_t* __restrict__ pl_valid_out,
                                                                 const int32_t* __restrict__ live_pl, const int32_t* __restrict__ row_start,
                                                                 const int32_t* __restrict__ plan) {
  extern __shared__ unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tc::uniform(tid >> 5), lane = tid & 31;
  const int quad = warp & 3, q4 = warp >> 2;
  const int r = quad * 32 + lane, c0 = 32 * q4;
  const int n_live = plan[0];
  const int n_tiles = (plan[1] + FILL - 1) / FILL;

  if (tid == 0) {
    tc::mbar_init(&sm.bar_w[0], 1);
    tc::mbar_init(&sm.bar_w[1], 1);
    tc::mbar_init(&sm.bar_mma, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&sm.tmem_base, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm0 = (uint32_t)tc::uniform((int)sm.tmem_base);
  const uint32_t tm = tm0 + ((uint32_t)(quad * 32) << 16);
  const uint32_t idesc = tc::make_idesc_bf16(128, 128);

  uint32_t w_loaded = 0, w_consumed = 0, n_mma = 0, n_ln = 0;
  auto prefetch = [&]() {  // thread 0: keep two weight stages in flight
    while (w_loaded < w_consumed + 2) {
      const uint32_t buf = w_loaded & 1;
      tc::mbar_expect_tx(&sm.bar_w[buf], tc::BLOCK_BYTES);
      tc::bulk_g2s(sm.w[buf], tcw + (size_t)stage_block(w_loaded) * tc::BLOCK_BYTES, tc::BLOCK_BYTES, &sm.bar_w[buf]);
      ++w_loaded;
    }
  };
  // A operand complete -> MMAs of `n_stage` consecutive weight stages into TMEM columns dst0, dst0 + 128, .. -> wait
  auto run_gemm = [&](int n_stage, uint32_t dst0) {
    tc::tmem_st_wait();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) {  // converged warp, one elected lane issues (operands stay warp-uniform)
      tc::tc_fence_after();
      for (int j = 0; j < n_stage; ++j) {
        const uint32_t s = w_consumed + j, buf = s & 1;
        tc::mbar_wait(&sm.bar_w[buf], (s >> 1) & 1);
        tc::tc_fence_after();
        const uint32_t wh = tc::smem_u32(sm.w[buf]);
        const uint64_t dh = tc::make_desc_sw128(wh), dl = tc::make_desc_sw128(wh + 2 * tc::KB_BYTES_128);
        if (tc::elect_one()) {
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint32_t ta = tm0 + TA + (term == 1 ? 64 : 0);
            const uint64_t db = term == 2 ? dl : dh;
#pragma unroll
            for (int k = 0; k < 128; k += 16)
              tc::mma_bf16_ts(tm0 + dst0 + 128 * j, ta + k / 2, db + (uint64_t)(((k >> 6) * tc::KB_BYTES_128 + (k & 63) * 2) >> 4), idesc,
                              (term > 0 || k > 0) ? 1u : 0u);
          }
        }
        __syncwarp();
      }
      if (tc::elect_one()) tc::mma_commit(&sm.bar_mma);
      __syncwarp();
    }
    tc::mbar_wait(&sm.bar_mma, n_mma & 1);
    tc::tc_fence_after();
    ++n_mma;
    w_consumed += n_stage;
    if (tid == 0) prefetch();
  };
  // LayerNorm over the 128 columns of a row held by 4 threads (32 columns each): one exchange of {sum, M2}
  auto ln32 = [&](float (&v)[32], const float* g, const float* bt) {  // g, bt: shared-memory vectors, read after the barrier
    float s4[4] = {0.f, 0.f, 0.f, 0.f};  // packed pairs (FADD2 / FFMA2): lanes (0, 1) and (2, 3) of the four-way accumulators
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
      float e0, e1, e2, e3;
      tc::sub2(e0, e1, v[i], v[i + 1], mloc, mloc);
      tc::sub2(e2, e3, v[i + 2], v[i + 3], mloc, mloc);
      tc::fma2(q[0], q[1], e0, e1, e0, e1, q[0], q[1]);
      tc::fma2(q[2], q[3], e2, e3, e2, e3, q[2], q[3]);
    }
    const int buf = n_ln & 1;
    ++n_ln;
    sm.red[buf][q4][r] = make_float2(sum, (q[0] + q[1]) + (q[2] + q[3]));
    __syncthreads();
    const float2 p0 = sm.red[buf][0][r], p1 = sm.red[buf][1][r], p2 = sm.red[buf][2][r], p3 = sm.red[buf][3][r];
    const float mean = ((p0.x + p1.x) + (p2.x + p3.x)) * (1.0f / 128);
    const float d0 = p0.x * (1.0f / 32) - mean, d1 = p1.x * (1.0f / 32) - mean, d2 = p2.x * (1.0f / 32) - mean, d3 = p3.x * (1.0f / 32) - mean;
    const float m2 = ((p0.y + p1.y) + (p2.y + p3.y)) + 32.0f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));  // Chan
    const float rstd = 1.0f / sqrtf(m2 * (1.0f / 128) + LN_EPS);
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float t0, t1;
      tc::sub2(t0, t1, v[i], v[i + 1], mean, mean);
      tc::mul2(t0, t1, t0, t1, rstd, rstd);
      tc::fma2(v[i], v[i + 1], t0, t1, g[c0 + i], g[c0 + i + 1], bt[c0 + i], bt[c0 + i + 1]);
    }
  };
  auto write_A = [&](const float (&v)[32]) {
    float ph[16], pl[16];
    tc::split32_packed(v, ph, pl);
    tc::tmem_st16(tm + TA + c0 / 2, ph);
    tc::tmem_st16(tm + TA + 64 + c0 / 2, pl);
  };
  auto load_acc = [&](uint32_t col, float (&v)[32]) {
    tc::tmem_ld32(tm + col + c0, v);
    tc::tmem_ld_wait();
  };
  if (tid == 0) prefetch();

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---- tile plan: the live polylines whose first row lies in [FILL tile, FILL tile + FILL) --------------------------------
    if (tid == 0) {
      const int lo = tile * FILL;
      int a = 0, b = n_live;  // first j with row_start[j] >= lo
      while (a < b) {
        const int m = (a + b) >> 1;
        if (row_start[m] < lo) a = m + 1; else b = m;
      }
      sm.j0 = a;
      sm.n_seg = 0;  // a tail tile may hold no polyline start at all (its rows belong to the previous tile's last polyline)
      sm.n_rows = 0;
    }
    __syncthreads();
    if (tid < 128) {
      const int j = sm.j0 + tid, hi = tile * FILL + FILL;
      const bool own = j < n_live && row_start[j] < hi;
      const bool last = own && !(j + 1 < n_live && row_start[j + 1] < hi);
      if (own) {
        const int base = row_start[sm.j0], st = row_start[j] - base, cnt = row_start[j + 1] - row_start[j];
        const int pl = live_pl[j];
        sm.seg_pl[tid] = pl;
        sm.seg_start[tid] = (uint8_t)st;
        sm.seg_cnt[tid] = (uint8_t)cnt;
        const uint8_t* v = in.map_valid + (long)pl * N;
        int k = 0;
        for (int n = 0; n < N; ++n)
          if (v[n]) {
            sm.row_seg[st + k] = (uint8_t)tid;
            sm.row_node[st + k] = (uint8_t)n;
            ++k;
          }
        if (last) {
          sm.n_seg = tid + 1;
          sm.n_rows = st + cnt;
        }
      }
    }
    __syncthreads();
    const int n_rows = sm.n_rows;
    const bool live = r < n_rows;
    const int p = live ? sm.row_seg[r] : 0, n = live ? sm.row_node[r] : 0;
    const long pl = live ? sm.seg_pl[p] : 0;
    const int ks = live ? sm.seg_start[p] : 0, kc = live ? sm.seg_cnt[p] : 0;  // key rows of this row's polyline
    float x[32];  // residual stream: columns c0 .. c0+31 of row r
    // ---- node features: InputPeEncoder([type | onehot(node)], PE(pos, atan2(dir)))  (sc_input.py:124-134) ---------
    bool valid = false;
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = 0.f;
    if (live) {
      const long node = pl * N + n;
      valid = true;  // compacted tiles hold valid nodes only
      {
        if (q4 == 0) {
          const float* w1 = packed + tbw::model_map_encoder_input_pe_encoder_mlp_fc_layers_0_weight;  // Wt4[8][32][4]
          float h[32];
#pragma unroll
          for (int o = 0; o < 32; ++o) h[o] = __ldg(packed + tbw::model_map_encoder_input_pe_encoder_mlp_fc_layers_0_bias + o);
          for (int k = 0; k < TB_PL_TYPE + N; ++k) {
            const bool on = k < TB_PL_TYPE ? (in.map_type[pl * TB_PL_TYPE + k] != 0) : (k - TB_PL_TYPE == n);
            if (!on) continue;
#pragma unroll
            for (int o = 0; o < 32; ++o) h[o] += __ldg(w1 + ((k >> 2) * 32 + o) * 4 + (k & 3));
          }
#pragma unroll
          for (int o = 0; o < 32; ++o) h[o] = fmaxf(h[o], 0.f);
          const float* w2 = packed + tbw::model_map_encoder_input_pe_encoder_mlp_fc_layers_3_weight;  // Wt4[8][32][4]
#pragma unroll
          for (int o = 0; o < 32; ++o) {
            float acc = __ldg(packed + tbw::model_map_encoder_input_pe_encoder_mlp_fc_layers_3_bias + o);
#pragma unroll
            for (int k = 0; k < 32; ++k) acc = fmaf(h[k], __ldg(w2 + ((k >> 2) * 32 + o) * 4 + (k & 3)), acc);
            x[o] = acc;
          }
        } else {
          const float px = in.map_pos[node * 2], py = in.map_pos[node * 2 + 1];
          const float yaw = atan2f(in.map_dir[node * 2 + 1], in.map_dir[node * 2]);
          const float* fxy = packed + tbw::pre_processing_input_pose_pe_map_pe_xy_freqs;
          const float* fyaw = packed + tbw::pre_processing_input_pose_pe_map_pe_yaw_freqs;
#pragma unroll 4
          for (int i = 0; i < 32; ++i) {
            const int j = 32 * (q4 - 1) + i;  // PE element: [cos(x f) 12 | sin(x f) 12 | cos(y f) 12 | sin(y f) 12 | cos(yaw g) 24 | sin(yaw g) 24]
            float v;
            if (j < 12) v = cosf(px * __ldg(fxy + 2 * j));
            else if (j < 24) v = sinf(px * __ldg(fxy + 2 * (j - 12) + 1));
            else if (j < 36) v = cosf(py * __ldg(fxy + 2 * (j - 24)));
            else if (j < 48) v = sinf(py * __ldg(fxy + 2 * (j - 36) + 1));
            else if (j < 72) v = cosf(yaw * __ldg(fyaw + 2 * (j - 48)));
            else v = sinf(yaw * __ldg(fyaw + 2 * (j - 72) + 1));
            x[i] = v;
          }
        }
      }
    }
    // initial node features = the attention target of all 3 layers: per-CTA scratch, row-minor [32 column quads][128 rows]
    float4* x0col = reinterpret_cast<float4*>(x0_scratch) + (size_t)blockIdx.x * 32 * 128 + (size_t)(8 * q4) * 128 + r;
#pragma unroll
    for (int i = 0; i < 8; ++i) x0col[i * 128] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
    const bool pvalid = live;

#pragma unroll 1
    for (int L = 0; L < 3; ++L) {
      const float* lw = packed + tbw::model_map_encoder_transformer_densetnt_layers_0_norm1_weight + L * tfl::STRIDE;
      // the layer's bias / LayerNorm vectors -> shared memory (visible after the barrier inside the first LayerNorm)
      float (*lp)[128] = sm.lp[L & 1];
      {
        const int off[12] = {tfl::NORMT_W, tfl::NORMT_B, tfl::NORM1_W, tfl::NORM1_B, tfl::IN_B, tfl::IN_B + 128, tfl::IN_B + 256,
                             tfl::OUT_B, tfl::NORM2_W, tfl::NORM2_B, tfl::L1_B, tfl::L2_B};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int i = tid + k * THREADS;
          lp[i >> 7][i & 127] = __ldg(lw + off[i >> 7] + (i & 127));
        }
      }
      enum { P_NT_W, P_NT_B, P_N1_W, P_N1_B, P_BQ, P_BK, P_BV, P_BO, P_N2_W, P_N2_B, P_B1, P_B2 };
      float t[32];
      // ---- K | V = LN_tgt(x0) Wkv  (tgt = the INITIAL node features in every layer, map_encoder.py:78-84) -----------
      if (L == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) t[i] = x[i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 q = x0col[i * 128];
          t[4 * i] = q.x, t[4 * i + 1] = q.y, t[4 * i + 2] = q.z, t[4 * i + 3] = q.w;
        }
      }
      ln32(t, lp[P_NT_W], lp[P_NT_B]);
      write_A(t);
      run_gemm(2, TK);  // -> K, V
      // ---- Q = LN1(x) Wq ------------------------------------------------------------------------------------------------
#pragma unroll
      for (int i = 0; i < 32; ++i) t[i] = x[i];
      ln32(t, lp[P_N1_W], lp[P_N1_B]);
      write_A(t);
      run_gemm(1, TQ);
      // ---- attention inside each polyline: thread (row, q4) = head q4 of the row ------------------------------------------
      load_acc(TK, t);
      if (r < ROWS) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          reinterpret_cast<float4*>(sm.kv + r * KVS + c0)[i] =
              make_float4(t[4 * i] + lp[P_BK][c0 + 4 * i], t[4 * i + 1] + lp[P_BK][c0 + 4 * i + 1], t[4 * i + 2] + lp[P_BK][c0 + 4 * i + 2],
                          t[4 * i + 3] + lp[P_BK][c0 + 4 * i + 3]);
      }
      __syncthreads();
      float pj[N];
      float inv = 0.f;
      load_acc(TQ, t);
      const int kmax = __reduce_max_sync(0xffffffffu, kc);  // longest polyline of the warp: later keys are skipped warp-wide
      if (live) {
#pragma unroll
        for (int i = 0; i < 32; ++i) t[i] += lp[P_BQ][c0 + i];
        float mx = -INFINITY;
        // keys in chunks of 4: one warp-uniform test per chunk (the longest polyline of the warp), the 4 keys of a chunk fully
        // unrolled so that their shared-memory loads overlap
#pragma unroll
        for (int j0 = 0; j0 < N; j0 += 4) {
          if (j0 >= kmax) {
#pragma unroll
            for (int e = 0; e < 4; ++e) pj[j0 + e] = -INFINITY;
            continue;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = j0 + e;
            const float* kr = sm.kv + (ks + (j < kc ? j : 0)) * KVS + c0;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
              const float4 k4 = reinterpret_cast<const float4*>(kr)[i], k5 = reinterpret_cast<const float4*>(kr)[i + 1];
              tc::fma2(a0, a1, t[4 * i], t[4 * i + 4], k4.x, k5.x, a0, a1);  // two independent chains, one packed FMA per step
              tc::fma2(a0, a1, t[4 * i + 1], t[4 * i + 5], k4.y, k5.y, a0, a1);
              tc::fma2(a0, a1, t[4 * i + 2], t[4 * i + 6], k4.z, k5.z, a0, a1);
              tc::fma2(a0, a1, t[4 * i + 3], t[4 * i + 7], k4.w, k5.w, a0, a1);
            }
            pj[j] = j < kc ? (a0 + a1) * 0.17677669529663687f : -INFINITY;
            mx = fmaxf(mx, pj[j]);
          }
        }
        if (mx != -INFINITY) {
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < N; ++j) {
            pj[j] = (pj[j] == -INFINITY) ? 0.f : expf(pj[j] - mx);
            sum += pj[j];
          }
          inv = 1.0f / sum;
        } else {
#pragma unroll
          for (int j = 0; j < N; ++j) pj[j] = 0.f;
        }
      }
      __syncthreads();  // every thread is done with K
      load_acc(TV, t);
      if (r < ROWS) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          reinterpret_cast<float4*>(sm.kv + r * KVS + c0)[i] =
              make_float4(t[4 * i] + lp[P_BV][c0 + 4 * i], t[4 * i + 1] + lp[P_BV][c0 + 4 * i + 1], t[4 * i + 2] + lp[P_BV][c0 + 4 * i + 2],
                          t[4 * i + 3] + lp[P_BV][c0 + 4 * i + 3]);
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 32; ++i) t[i] = 0.f;
      if (live) {
#pragma unroll
        for (int j0 = 0; j0 < N; j0 += 4) {
          if (j0 >= kmax) continue;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = j0 + e;
            const float* vr = sm.kv + (ks + (j < kc ? j : 0)) * KVS + c0;  // pj[j] = 0 for j >= kc
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 v4 = reinterpret_cast<const float4*>(vr)[i];
              tc::fma2(t[4 * i], t[4 * i + 1], pj[j], pj[j], v4.x, v4.y, t[4 * i], t[4 * i + 1]);
              tc::fma2(t[4 * i + 2], t[4 * i + 3], pj[j], pj[j], v4.z, v4.w, t[4 * i + 2], t[4 * i + 3]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) t[i] *= inv;
      }
      write_A(t);
      // ---- out-proj, residual (dead rows = polylines without a valid node get no attention update) ----------------------
      run_gemm(1, TQ);
      load_acc(TQ, t);
      if (pvalid) {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] += t[i] + lp[P_BO][c0 + i];
      }
      // ---- FFN ------------------------------------------------------------------------------------------------------------
#pragma unroll
      for (int i = 0; i < 32; ++i) t[i] = x[i];
      ln32(t, lp[P_N2_W], lp[P_N2_B]);
      write_A(t);
      run_gemm(1, TQ);
      load_acc(TQ, t);
#pragma unroll
      for (int i = 0; i < 32; ++i) t[i] = fmaxf(t[i] + lp[P_B1][c0 + i], 0.f);
      write_A(t);
      run_gemm(1, TQ);
      load_acc(TQ, t);
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = valid ? x[i] + t[i] + lp[P_B2][c0 + i] : 0.f;
    }
    // ---- masked max-pool over the valid nodes of each polyline (map_encoder.py:95-97,105-106) ------------------------------
    {
      float* stage = sm.kv;  // [row][col] with the K|V staging stride: 16-byte row writes and column-contiguous reads are both
                             // conflict-free (a [col][row] layout made every read of the pooling loop a 32-way bank conflict)
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i)
        reinterpret_cast<float4*>(stage + r * KVS + c0)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
      __syncthreads();
      const int n_seg = sm.n_seg;
      for (int idx = tid; idx < n_seg * 128; idx += THREADS) {
        const int pp = idx >> 7, c = idx & 127;
        const int st = sm.seg_start[pp], cnt = sm.seg_cnt[pp];
        float mx = -INFINITY;
        for (int j = 0; j < cnt; ++j) mx = fmaxf(mx, stage[(st + j) * KVS + c]);
        pl_feature[(long)sm.seg_pl[pp] * 128 + c] = mx;
      }
      if (tid < n_seg) pl_valid_out[sm.seg_pl[tid]] = 1;
      __syncthreads();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(sm.tmem_base, 512);
}

}  // namespace pl2
}  // namespace tb

size_t tb::map_plan_bytes(long n_pl_total) {  // [live_pl | row_start (+1) | plan (4)] int32 + counts u8
  return ((size_t)(2 * n_pl_total + 1 + 4) * sizeof(int32_t) + (size_t)n_pl_total + 255) & ~(size_t)255;
}

int tb::launch_map_polyline_tc2(const TbDims& d, const TbSceneIn& in, const float* packed, float* x0_scratch, int n_cta,
                                float* pl_feature, uint8_t* pl_valid, int32_t* plan_ws, cudaStream_t st) {
  static std::atomic<uint64_t> attr_set{0};
  const int smem = (int)sizeof(pl2::Smem) + 1024;
  if (!smem_attr_done(attr_set)) {
    if (!set_max_smem(pl2::k_map_polyline_tc2, smem)) return TB_ERR_LAUNCH;
    smem_attr_mark(attr_set);
  }
  const long n_pl = (long)d.n_scene * d.n_pl;
  uint8_t* counts = reinterpret_cast<uint8_t*>(plan_ws + 2 * n_pl + 1 + 4);
  pl2::k_map_count<<<(unsigned)((n_pl + 255) / 256), 256, 0, st>>>(n_pl, in.map_valid, counts, pl_feature, pl_valid);
  count_launch();
  pl2::k_map_plan<<<1, pl2::PLAN_THREADS, 0, st>>>(n_pl, counts, plan_ws, plan_ws + n_pl, plan_ws + 2 * n_pl + 1);
  count_launch();
  // persistent CTAs; the tile count is a device value (plan[1]): surplus CTAs exit at once
  const long max_tiles = (n_pl * pl2::N + pl2::FILL - 1) / pl2::FILL;
  const int grid = (int)(max_tiles < n_cta ? max_tiles : n_cta);
  pl2::k_map_polyline_tc2<<<grid, pl2::THREADS, smem, st>>>(d, in, packed, tc_blob(packed), x0_scratch, pl_feature, pl_valid, plan_ws,
                                                          plan_ws + n_pl, plan_ws + 2 * n_pl + 1);
  count_launch();
  return launch_status();
}
