// Tensor-core (tcgen05) kernels.  First: the self-test of the bf16x3 GEMM machinery used by the ABI's tb_tc_selftest.
#include "tb_host.h"

namespace tb {

struct SelftestSmem {
  unsigned char a_hi[2 * tc::KB_BYTES_128];  // 128 x 128 bf16, two K-blocks
  unsigned char a_lo[2 * tc::KB_BYTES_128];
  unsigned char w[tc::BLOCK_BYTES];          // [hi kb0 | hi kb1 | lo kb0 | lo kb1]
  uint64_t bar_w, bar_mma;
  uint32_t tmem_base;
};

// d[128,128] = a[128,128] @ W_block^T with W_block = one packed 128x128 tensor-core weight block
// mode 0: A operand from shared memory (SS); mode 1: A operand from tensor memory (TS)
__global__ void __launch_bounds__(128) k_tc_selftest(const float* __restrict__ a, const unsigned char* __restrict__ wblock,
                                                     float* __restrict__ d, int mode) {
  extern __shared__ unsigned char smem_raw[];
  SelftestSmem& sm = *reinterpret_cast<SelftestSmem*>(
      smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));  // SWIZZLE_128B tiles need 1 KB alignment
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    tc::mbar_init(&sm.bar_w, 1);
    tc::mbar_init(&sm.bar_mma, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&sm.tmem_base, 256);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  if (tid == 0) {
    tc::mbar_expect_tx(&sm.bar_w, tc::BLOCK_BYTES);
    tc::bulk_g2s(sm.w, wblock, tc::BLOCK_BYTES, &sm.bar_w);
  }
  // thread = row: split the fp32 row into bf16 hi / lo operand tiles
  for (int k0 = 0; k0 < 128; k0 += 32) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = a[tid * 128 + k0 + i];
    tc::store_row32_split(sm.a_hi, sm.a_lo, tid, k0, v);
    // TS mode: the same operand as packed bf16 pairs in TMEM, hi at columns [128,192), lo at [192,256)
    float ph[16], pl[16];
    tc::split32_packed(v, ph, pl);
    tc::tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 128 + k0 / 2, ph);
    tc::tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 192 + k0 / 2, pl);
  }
  tc::tmem_st_wait();
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc::mbar_wait(&sm.bar_w, 0);
    tc::tc_fence_after();
    const uint32_t idesc = tc::make_idesc_bf16(128, 128);
    const uint32_t ah = tc::smem_u32(sm.a_hi), al = tc::smem_u32(sm.a_lo), wh = tc::smem_u32(sm.w), wl = wh + 2 * tc::KB_BYTES_128;
    if (mode == 0) {
      tc::mma_tile(tmem, ah, tc::KB_BYTES_128, wh, tc::KB_BYTES_128, 128, idesc, false);
      tc::mma_tile(tmem, al, tc::KB_BYTES_128, wh, tc::KB_BYTES_128, 128, idesc, true);
      tc::mma_tile(tmem, ah, tc::KB_BYTES_128, wl, tc::KB_BYTES_128, 128, idesc, true);
    } else {
      for (int term = 0; term < 3; ++term) {
        const uint32_t ta = tmem + (term == 1 ? 192 : 128), wb = term == 2 ? wl : wh;
        for (int k = 0; k < 128; k += 16)
          tc::mma_bf16_ts(tmem, ta + k / 2, tc::make_desc_sw128(wb + (k >> 6) * tc::KB_BYTES_128 + (k & 63) * 2), idesc,
                          (term > 0 || k > 0) ? 1u : 0u);
      }
    }
    tc::mma_commit(&sm.bar_mma);
  }
  tc::mbar_wait(&sm.bar_mma, 0);
  tc::tc_fence_after();
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) d[tid * 128 + c0 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}


// ------------------------------------------------------------------------------------------------------------
// Polyline encoder on the tensor pipe (map_encoder.py:72-106).  One CTA = 128 threads, thread = one node row
// (= TMEM lane); a tile = 6 polylines x 20 nodes (120 rows, 8 idle).  Persistent over tiles.
//   TMEM columns:  X [0,128) residual stream | Q [128,256) | K [256,384) | V [384,512)   (fp32, lane = row)
//   every Linear = bf16x3 tcgen05 GEMM, A operand written by the row threads (LayerNorm / ReLU fused into the write),
//   weights streamed through a 2 x 64 KB ring by bulk-async copies; the 20x20 per-polyline attention runs on the
//   CUDA cores from a per-head fp32 staging buffer.
// ------------------------------------------------------------------------------------------------------------
constexpr int MT_NP = 6;                 // polylines per tile
constexpr int MT_ROWS = MT_NP * TB_PL_NODE;  // 120
constexpr int KVS = 68;                  // staging row stride (floats): conflict-free 16-byte row writes

struct MapTcSmem {
  unsigned char a_hi[2 * tc::KB_BYTES_128];
  unsigned char a_lo[2 * tc::KB_BYTES_128];
  unsigned char w[2][tc::BLOCK_BYTES];
  float kvs[MT_ROWS * KVS];
  uint64_t bar_w[2];
  uint64_t bar_mma;
  uint32_t tmem_base;
  uint8_t row_valid[128];
  uint8_t pl_valid[8];
};

struct WeightPipe {  // uniform across the CTA; only thread 0 touches the barriers / issues copies
  uint32_t loaded;    // weight stages whose copy has been issued
  uint32_t consumed;  // weight stages whose MMAs have completed
  uint32_t mma_count; // commits waited so far
};

__device__ __forceinline__ int map_stage_block(uint32_t s) {  // stage s of the repeating 18-stage schedule
  const int L = (s % 18) / 6, j = s % 6;
  // order inside a layer: Wk, Wv, Wq, Wo, W1, W2   (blocks of a layer: q,k,v,out,linear1,linear2)
  const int blk[6] = {1, 2, 0, 3, 4, 5};
  return tbb::model_map_encoder_transformer_densetnt_layers_0_attn_in_proj_weight + L * 6 + blk[j];
}

__device__ __forceinline__ void ld_row128(uint32_t taddr, float (&v)[128]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) tc::tmem_ld32(taddr + 32 * i, *reinterpret_cast<float(*)[32]>(&v[32 * i]));
  tc::tmem_ld_wait();
}
__device__ __forceinline__ void st_row128(uint32_t taddr, const float (&v)[128]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) tc::tmem_st32(taddr + 32 * i, *reinterpret_cast<const float(*)[32]>(&v[32 * i]));
  tc::tmem_st_wait();
}
// LayerNorm of a register row, written as the bf16 hi / lo A operand (row r of the 128-row tiles)
__device__ __forceinline__ void ln_row_to_A(const float (&v)[128], const float* __restrict__ g, const float* __restrict__ b,
                                            unsigned char* a_hi, unsigned char* a_lo, int r) {
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 128; ++i) s4[i & 3] += v[i];
  const float mean = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.0f / 128);
  float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 128; ++i) {
    const float d = v[i] - mean;
    q4[i & 3] = fmaf(d, d, q4[i & 3]);
  }
  const float q = (q4[0] + q4[1]) + (q4[2] + q4[3]);
  const float rstd = 1.0f / sqrtf(q * (1.0f / 128) + LN_EPS);
#pragma unroll
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = (v[c0 + i] - mean) * rstd * __ldg(g + c0 + i) + __ldg(b + c0 + i);
    tc::store_row32_split(a_hi, a_lo, r, c0, o);
  }
}

__global__ void __launch_bounds__(128, 1) k_map_polyline_tc(TbDims dm, TbSceneIn in, const float* __restrict__ packed,
                                                            const unsigned char* __restrict__ tcw, float* __restrict__ x0_scratch,
                                                            float* __restrict__ pl_feature, uint8_t* __restrict__ pl_valid_out,
                                                            int n_tiles) {
  extern __shared__ unsigned char smem_raw[];
  MapTcSmem& sm = *reinterpret_cast<MapTcSmem*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
  constexpr int N = TB_PL_NODE;
  const int tid = threadIdx.x, warp = tc::uniform(tid >> 5);
  const long n_pl_total = (long)dm.n_scene * dm.n_pl;

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
  const uint32_t tm = sm.tmem_base + ((uint32_t)(warp * 32) << 16);  // this warp's lane quadrant
  const uint32_t TX = tm, TQ = tm + 128, TK = tm + 256, TV = tm + 384;
  const uint32_t tm0 = (uint32_t)tc::uniform((int)sm.tmem_base);
  const uint32_t idesc = tc::make_idesc_bf16(128, 128);
  const uint32_t ah = tc::smem_u32(sm.a_hi), al = tc::smem_u32(sm.a_lo);

  WeightPipe wp{0, 0, 0};
  auto prefetch = [&]() {  // thread 0: keep two weight stages in flight
    while (wp.loaded < wp.consumed + 2) {
      const uint32_t s = wp.loaded, buf = s & 1;
      tc::mbar_expect_tx(&sm.bar_w[buf], tc::BLOCK_BYTES);
      tc::bulk_g2s(sm.w[buf], tcw + (size_t)map_stage_block(s) * tc::BLOCK_BYTES, tc::BLOCK_BYTES, &sm.bar_w[buf]);
      ++wp.loaded;
    }
  };
  // issue the three bf16x3 MMAs of weight stage `s` into TMEM columns `dst`: warp 0 runs this converged (operands stay
  // warp-uniform, see tc::elect_one), one elected lane issues
  auto issue = [&](uint32_t s, uint32_t dst_col) {
    const uint32_t buf = s & 1;
    tc::mbar_wait(&sm.bar_w[buf], (s >> 1) & 1);
    tc::tc_fence_after();
    const uint32_t wh = tc::smem_u32(sm.w[buf]), wl = wh + 2 * tc::KB_BYTES_128;
    if (tc::elect_one()) {
      tc::mma_tile(tm0 + dst_col, ah, tc::KB_BYTES_128, wh, tc::KB_BYTES_128, 128, idesc, false);
      tc::mma_tile(tm0 + dst_col, al, tc::KB_BYTES_128, wh, tc::KB_BYTES_128, 128, idesc, true);
      tc::mma_tile(tm0 + dst_col, ah, tc::KB_BYTES_128, wl, tc::KB_BYTES_128, 128, idesc, true);
    }
    __syncwarp();
  };
  // A operand complete -> MMAs of `n_stage` consecutive weight stages -> wait for completion (all threads)
  auto run_gemm = [&](int n_stage, uint32_t dst_col0) {
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc::tc_fence_after();
      for (int j = 0; j < n_stage; ++j) issue(wp.consumed + j, dst_col0 + 128 * j);
      if (tc::elect_one()) tc::mma_commit(&sm.bar_mma);
      __syncwarp();
    }
    tc::mbar_wait(&sm.bar_mma, wp.mma_count & 1);
    tc::tc_fence_after();
    wp.consumed += n_stage;
    ++wp.mma_count;
    if (tid == 0) prefetch();
  };
  if (tid == 0) prefetch();

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long pl0 = (long)tile * MT_NP;
    const int r = tid, p = r / N, n = r % N;
    const long pl = pl0 + p;
    const bool live = r < MT_ROWS && pl < n_pl_total;
    float v[128];
    // ---- node features: InputPeEncoder([type | onehot(node)], PE(pos, atan2(dir)))  (sc_input.py:124-134) ---------
    bool valid = false;
    {
#pragma unroll
      for (int i = 0; i < 128; ++i) v[i] = 0.f;
      if (live) {
        const long node = pl * N + n;
        valid = in.map_valid[node] != 0;
        const float px = in.map_pos[node * 2], py = in.map_pos[node * 2 + 1];
        const float yaw = atan2f(in.map_dir[node * 2 + 1], in.map_dir[node * 2]);
        if (valid) {
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
            v[o] = acc;
          }
          const float* fxy = packed + tbw::pre_processing_input_pose_pe_map_pe_xy_freqs;
          const float* fyaw = packed + tbw::pre_processing_input_pose_pe_map_pe_yaw_freqs;
#pragma unroll
          for (int i = 0; i < 12; ++i) {
            v[32 + i] = cosf(px * __ldg(fxy + 2 * i));
            v[44 + i] = sinf(px * __ldg(fxy + 2 * i + 1));
            v[56 + i] = cosf(py * __ldg(fxy + 2 * i));
            v[68 + i] = sinf(py * __ldg(fxy + 2 * i + 1));
          }
#pragma unroll
          for (int i = 0; i < 24; ++i) {
            v[80 + i] = cosf(yaw * __ldg(fyaw + 2 * i));
            v[104 + i] = sinf(yaw * __ldg(fyaw + 2 * i + 1));
          }
        }
      }
      sm.row_valid[r] = valid;
    }
    st_row128(TX, v);
    // per-CTA scratch in row-minor layout [32 column quads][128 rows] float4: a warp touches 512 contiguous bytes
    float4* x0col = reinterpret_cast<float4*>(x0_scratch) + (size_t)blockIdx.x * 32 * 128 + r;
#pragma unroll
    for (int i = 0; i < 32; ++i) x0col[i * 128] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    __syncthreads();
    if (tid < 8) {
      bool any = false;
      if (tid < MT_NP)
        for (int j = 0; j < N; ++j) any |= sm.row_valid[tid * N + j] != 0;
      sm.pl_valid[tid] = any;
    }
    __syncthreads();
    const bool pvalid = sm.pl_valid[p < 8 ? p : 7] != 0;

#pragma unroll 1
    for (int L = 0; L < 3; ++L) {
      const float* lw = packed + tbw::model_map_encoder_transformer_densetnt_layers_0_norm1_weight + L * tfl::STRIDE;
      // ---- K | V = LN_tgt(x0) Wkv  (tgt = the INITIAL node features in every layer, map_encoder.py:78-84) -----------
      if (L > 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float4 t = x0col[i * 128];
          v[4 * i] = t.x, v[4 * i + 1] = t.y, v[4 * i + 2] = t.z, v[4 * i + 3] = t.w;
        }
      }
      ln_row_to_A(v, lw + tfl::NORMT_W, lw + tfl::NORMT_B, sm.a_hi, sm.a_lo, r);
      run_gemm(2, 256);  // -> K, V
      // ---- Q = LN1(x) Wq -----------------------------------------------------------------------------------------------
      ld_row128(TX, v);
      ln_row_to_A(v, lw + tfl::NORM1_W, lw + tfl::NORM1_B, sm.a_hi, sm.a_lo, r);
      run_gemm(1, 128);  // -> Q
      // ---- attention inside each polyline, one head at a time ---------------------------------------------------------------
#pragma unroll 1
      for (int h = 0; h < NHEAD; ++h) {
        {
          float kk[32], vv[32];
          tc::tmem_ld32(TK + h * 32, kk);
          tc::tmem_ld32(TV + h * 32, vv);
          tc::tmem_ld_wait();
          float* row = sm.kvs + (r < MT_ROWS ? r : 0) * KVS;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (r >= MT_ROWS) break;
            float4 a, b;
            a.x = kk[4 * i] + __ldg(lw + tfl::IN_B + D + h * 32 + 4 * i);
            a.y = kk[4 * i + 1] + __ldg(lw + tfl::IN_B + D + h * 32 + 4 * i + 1);
            a.z = kk[4 * i + 2] + __ldg(lw + tfl::IN_B + D + h * 32 + 4 * i + 2);
            a.w = kk[4 * i + 3] + __ldg(lw + tfl::IN_B + D + h * 32 + 4 * i + 3);
            b.x = vv[4 * i] + __ldg(lw + tfl::IN_B + 2 * D + h * 32 + 4 * i);
            b.y = vv[4 * i + 1] + __ldg(lw + tfl::IN_B + 2 * D + h * 32 + 4 * i + 1);
            b.z = vv[4 * i + 2] + __ldg(lw + tfl::IN_B + 2 * D + h * 32 + 4 * i + 2);
            b.w = vv[4 * i + 3] + __ldg(lw + tfl::IN_B + 2 * D + h * 32 + 4 * i + 3);
            reinterpret_cast<float4*>(row)[i] = a;
            reinterpret_cast<float4*>(row + 32)[i] = b;
          }
        }
        __syncthreads();
        float o[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0.f;
        float q[32];
        tc::tmem_ld32(TQ + h * 32, q);  // .sync.aligned: the whole warp executes it, including the 8 idle rows
        tc::tmem_ld_wait();
        if (r < MT_ROWS) {
#pragma unroll
          for (int i = 0; i < 32; ++i) q[i] += __ldg(lw + tfl::IN_B + h * 32 + i);
          float lg[N];
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < N; ++j) {
            const float* kr = sm.kvs + (p * N + j) * KVS;
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 k4 = reinterpret_cast<const float4*>(kr)[i];
              acc = fmaf(q[4 * i + 3], k4.w, fmaf(q[4 * i + 2], k4.z, fmaf(q[4 * i + 1], k4.y, fmaf(q[4 * i], k4.x, acc))));
            }
            lg[j] = sm.row_valid[p * N + j] ? acc * 0.17677669529663687f : -INFINITY;
            mx = fmaxf(mx, lg[j]);
          }
          if (mx != -INFINITY) {
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < N; ++j) {
              const float pj = (lg[j] == -INFINITY) ? 0.f : expf(lg[j] - mx);
              sum += pj;
              const float* vr = sm.kvs + (p * N + j) * KVS + 32;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 v4 = reinterpret_cast<const float4*>(vr)[i];
                o[4 * i] = fmaf(pj, v4.x, o[4 * i]);
                o[4 * i + 1] = fmaf(pj, v4.y, o[4 * i + 1]);
                o[4 * i + 2] = fmaf(pj, v4.z, o[4 * i + 2]);
                o[4 * i + 3] = fmaf(pj, v4.w, o[4 * i + 3]);
              }
            }
            const float inv = 1.0f / sum;
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] *= inv;
          }
        }
        tc::store_row32_split(sm.a_hi, sm.a_lo, r, h * 32, o);
        __syncthreads();  // staging buffer is rewritten by the next head
      }
      // ---- out-proj, residual (dead rows = polylines without a valid node get no attention update) ----------------------
      run_gemm(1, 128);
      ld_row128(TX, v);
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float a[32];
        tc::tmem_ld32(TQ + c0, a);
        tc::tmem_ld_wait();
        if (pvalid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[c0 + i] += a[i] + __ldg(lw + tfl::OUT_B + c0 + i);
        }
      }
      st_row128(TX, v);
      // ---- FFN ------------------------------------------------------------------------------------------------------------
      ln_row_to_A(v, lw + tfl::NORM2_W, lw + tfl::NORM2_B, sm.a_hi, sm.a_lo, r);
      run_gemm(1, 128);
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float a[32];
        tc::tmem_ld32(TQ + c0, a);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = fmaxf(a[i] + __ldg(lw + tfl::L1_B + c0 + i), 0.f);
        tc::store_row32_split(sm.a_hi, sm.a_lo, r, c0, a);
      }
      run_gemm(1, 128);
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float a[32];
        tc::tmem_ld32(TQ + c0, a);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[c0 + i] = valid ? v[c0 + i] + a[i] + __ldg(lw + tfl::L2_B + c0 + i) : 0.f;
      }
      st_row128(TX, v);
    }
    // ---- masked max-pool over the valid nodes of each polyline (map_encoder.py:95-97,105-106) ------------------------------
    // stage the rows as fp32 in the (now idle) A-operand region: 128 x 128 floats = 64 KB = a_hi + a_lo
    {
      float* stage = reinterpret_cast<float*>(sm.a_hi);
      __syncthreads();  // all MMAs reading the A tiles have completed (run_gemm waited); make the reuse explicit
#pragma unroll
      for (int i = 0; i < 128; ++i) stage[i * 128 + r] = v[i];  // transposed: [col][row] -> conflict-free both ways
      __syncthreads();
      for (int pp = 0; pp < MT_NP; ++pp) {
        if (pl0 + pp >= n_pl_total) break;
        float mx = -INFINITY;
        for (int j = 0; j < N; ++j)
          if (sm.row_valid[pp * N + j]) mx = fmaxf(mx, stage[tid * 128 + pp * N + j]);
        pl_feature[(pl0 + pp) * 128 + tid] = sm.pl_valid[pp] ? mx : 0.f;
      }
      if (tid < MT_NP && pl0 + tid < n_pl_total) pl_valid_out[pl0 + tid] = sm.pl_valid[tid];
      __syncthreads();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(sm.tmem_base, 512);
}

}  // namespace tb

using namespace tb;

extern "C" int32_t tb_tc_selftest(const float* a, int32_t block, const float* packed, float* d, int32_t mode, void* stream) {
  if (!a || !packed || !d) return TB_ERR_NULL;
  if (block < 0 || block >= TB_N_TC_BLOCKS) return TB_ERR_BAD_SHAPE;
  if (!aligned16(a) || !aligned16(packed) || !aligned16(d)) return TB_ERR_ALIGN;
  static std::atomic<uint64_t> attr_set{0};
  if (!smem_attr_done(attr_set)) {
    if (!set_max_smem(k_tc_selftest, (int)sizeof(SelftestSmem) + 1024)) return TB_ERR_LAUNCH;
    smem_attr_mark(attr_set);
  }
  k_tc_selftest<<<1, 128, sizeof(SelftestSmem) + 1024, (cudaStream_t)stream>>>(a, tc_blob(packed) + (size_t)block * tc::BLOCK_BYTES, d, mode);
  count_launch();
  return launch_status();
}

// scratch: one fp32 copy of the initial node features per resident CTA
size_t tb::map_tc_scratch_bytes(int n_cta) { return (size_t)n_cta * 128 * 128 * sizeof(float); }

int tb::launch_map_polyline_tc(const TbDims& d, const TbSceneIn& in, const float* packed, float* x0_scratch, int n_cta,
                               float* pl_feature, uint8_t* pl_valid, cudaStream_t st) {
  static std::atomic<uint64_t> attr_set{0};
  const int smem = (int)sizeof(MapTcSmem) + 1024;
  if (!smem_attr_done(attr_set)) {
    if (!set_max_smem(k_map_polyline_tc, smem)) return TB_ERR_LAUNCH;
    smem_attr_mark(attr_set);
  }
  const long n_pl = (long)d.n_scene * d.n_pl;
  const int n_tiles = (int)((n_pl + MT_NP - 1) / MT_NP);
  const int grid = n_tiles < n_cta ? n_tiles : n_cta;
  k_map_polyline_tc<<<grid, 128, smem, st>>>(d, in, packed, tc_blob(packed), x0_scratch, pl_feature, pl_valid, n_tiles);
  count_launch();
  return launch_status();
}
