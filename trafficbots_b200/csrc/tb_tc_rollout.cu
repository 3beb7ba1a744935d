// Tensor-core decode step, front half (tcgen05): state embedding -> 3 agent->map layers -> 3 agent->traffic-light
// layers -> interaction K|V, for ONE scene-mode per CTA (rows = agents, up to 128; thread pair per row).
//
//   TMEM (512 columns, lane = agent row):
//     [  0,128)  S0   logits / probabilities buffer 0  (4 heads x 32 keys)   | GEMM accumulator outside attention
//     [128,256)  S1   logits / probabilities buffer 1                         | second GEMM accumulator (K|V pairs)
//     [256,384)  O    attention output accumulator (4 heads x 32 dims)
//     [384,512)  A    bf16x2-packed A operand of the next MMA: hi [384,448), lo [448,512)   (Q during attention)
//   shared memory: 2 x 64 KB ring (weight blocks and K|V key tiles arrive by bulk-async copies in one static order),
//     the fp32 residual stream x [col][row] (64 KB), LayerNorm exchange, barriers.
//
// Every Linear and both attention contractions are bf16x3 tcgen05 MMAs with the A operand in tensor memory:
//   QK^T:  S_h[128 x 32 keys] = Q_h[128 x 32] K_h^T     (N = 32, K = 32: 2 k-steps x 3 terms per head)
//   PV  :  O_h[128 x 32]     += P_h[128 x 32 keys] V_h   (P overwrites S in place as packed bf16 hi | lo)
// The softmax is an online softmax with lazy rescaling: O and l are rescaled only when a row's running maximum grows by
// more than 2^8, so the tensor-memory read-modify-write of O is rare; the final O / l is exact either way.
// QK^T of sub-tile u+2 and PV of sub-tile u are issued together, so the tensor pipe works while the CUDA cores do the
// exponentials of sub-tile u+1.
#include "tb_host.h"

namespace tb {

constexpr int KVT_KEYS = 64;  // keys per 64 KB block
constexpr int SUB_KEYS = 32;  // keys per softmax sub-tile

// ------------------------------------------------------------------------------------------------------------
// K|V cache -> tensor-core blocks (valid keys only, compacted)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_kv_tc(const float* __restrict__ kv, const uint8_t* __restrict__ key_valid, int T,
                                                    int nT, unsigned char* __restrict__ blocks, int32_t* __restrict__ n_key,
                                                    int n_set_valid /* key_valid sets; kv has gridDim.x sets */) {
  extern __shared__ int idx_s[];  // [nT * 64] compacted key indices (-1 = padding)
  __shared__ int count_s;
  const int set = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const float* kvs = kv + (size_t)set * T * 256;
  const uint8_t* kval = key_valid + (size_t)(set % n_set_valid) * T;
  unsigned char* out = blocks + (size_t)set * nT * tc::BLOCK_BYTES;
  for (int i = tid; i < nT * KVT_KEYS; i += 256) idx_s[i] = -1;
  __syncthreads();
  if (tid < 32) {
    int count = 0;
    for (int base = 0; base < T; base += 32) {
      const bool v = (base + lane < T) && kval[base + lane] != 0;
      const unsigned m = __ballot_sync(0xffffffffu, v);
      if (v) idx_s[count + __popc(m & ((1u << lane) - 1))] = base + lane;
      count += __popc(m);
    }
    if (lane == 0) {
      count_s = count;
      if (n_key && set < n_set_valid) n_key[set] = count;
    }
  }
  __syncthreads();
  for (int item = tid; item < nT * 2048; item += 256) {
    const int t = item >> 11, w = item & 2047;
    unsigned char* blk = out + (size_t)t * tc::BLOCK_BYTES;
    float v[8];
    if (w < 1024) {  // K: key slot i, 8 consecutive dims
      const int i = w >> 4, c = w & 15;
      const int key = idx_s[t * KVT_KEYS + i];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = key >= 0 ? kvs[(size_t)key * 256 + c * 8 + e] : 0.f;
      uint4 hi, lo;
      tc::split8(v, hi, lo);
      const uint32_t off = (c >> 3) * 8192 + tc::sw128_off(i, c & 7);
      *reinterpret_cast<uint4*>(blk + off) = hi;
      *reinterpret_cast<uint4*>(blk + 16384 + off) = lo;
    } else {  // V^T: dim d, 8 consecutive key slots
      const int d = (w - 1024) >> 3, kc = (w - 1024) & 7;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int key = idx_s[t * KVT_KEYS + kc * 8 + e];
        v[e] = key >= 0 ? kvs[(size_t)key * 256 + 128 + d] : 0.f;
      }
      uint4 hi, lo;
      tc::split8(v, hi, lo);
      const uint32_t off = tc::sw128_off(d, kc);
      *reinterpret_cast<uint4*>(blk + 32768 + off) = hi;
      *reinterpret_cast<uint4*>(blk + 49152 + off) = lo;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// front half of the decode step
// ------------------------------------------------------------------------------------------------------------
constexpr int MAX_STAGE = 192;

struct FrontTcSmem {
  unsigned char ring[2][tc::BLOCK_BYTES];
  float xs[128 * 128];  // residual stream, [col][row]
  float red[2][128];    // LayerNorm partial sums of the two column halves
  const unsigned char* sched[MAX_STAGE];
  uint64_t bar_ring[2];
  uint64_t bar_mma;
  uint64_t bar_s[2];
  uint64_t bar_pv;
  uint32_t tmem_base;
  int n_stage;
  uint8_t row_valid[128];
};

constexpr uint32_t T_S0 = 0, T_S1 = 128, T_O = 256, T_A = 384;

struct FrontArgs {
  TbDims dm;
  TbRolloutIn in;
  const float* packed;
  const unsigned char* tcw;
  StateView sv;
  int t;
};

__global__ void __launch_bounds__(256, 1) k_step_front_tc(FrontArgs a) {
  extern __shared__ unsigned char smem_raw[];
  FrontTcSmem& sm = *reinterpret_cast<FrontTcSmem*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
  const TbDims& dm = a.dm;
  const TbRolloutIn& in = a.in;
  const float* __restrict__ packed = a.packed;
  const int A = dm.n_agent, K = dm.n_mode, S = dm.n_scene, B = S * K, Th = dm.n_step_hist;
  const int nT_map = (dm.n_pl + KVT_KEYS - 1) / KVT_KEYS, nT_tl = (dm.n_tl + KVT_KEYS - 1) / KVT_KEYS;
  const int b = blockIdx.x, s = b / K, t = a.t;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2;  // TMEM lane quadrant / column half (heads 2*half, 2*half+1)
  const int r = quad * 32 + lane;               // agent row
  const int c0 = half * 64;                     // first owned column
  const size_t BA = (size_t)B * A;
  const bool live = r < A;
  const size_t ba = (size_t)b * A + (live ? r : 0), sa = (size_t)s * A + (live ? r : 0);
  const int tl_t = min(t - 1, Th - 1);
  const int nkey_map = in.n_key_map[s], nkey_tl = in.n_key_tl[(size_t)s * Th + tl_t];

  // ---- set-up: barriers, TMEM, static schedule of the 64 KB blocks this step consumes --------------------------------
  if (tid == 0) {
    tc::mbar_init(&sm.bar_ring[0], 1);
    tc::mbar_init(&sm.bar_ring[1], 1);
    tc::mbar_init(&sm.bar_mma, 1);
    tc::mbar_init(&sm.bar_s[0], 1);
    tc::mbar_init(&sm.bar_s[1], 1);
    tc::mbar_init(&sm.bar_pv, 1);
    tc::fence_mbar_init();
    int n = 0;
    auto wblk = [&](int first, int idx) { return a.tcw + (size_t)(first + idx) * tc::BLOCK_BYTES; };
    for (int L = 0; L < 3; ++L) {
      const int w0 = tbb::model_transformer_as2pl_layers_0_attn_in_proj_weight + 6 * L;
      sm.sched[n++] = wblk(w0, 0);  // Wq
      for (int j = 0; j < (nkey_map + KVT_KEYS - 1) / KVT_KEYS; ++j)
        sm.sched[n++] = in.kv_map_tc + (((size_t)L * S + s) * nT_map + j) * tc::BLOCK_BYTES;
      sm.sched[n++] = wblk(w0, 3);  // Wo
      sm.sched[n++] = wblk(w0, 4);  // W1
      sm.sched[n++] = wblk(w0, 5);  // W2
    }
    for (int L = 0; L < 3; ++L) {
      const int w0 = tbb::model_transformer_as2tl_layers_0_attn_in_proj_weight + 6 * L;
      sm.sched[n++] = wblk(w0, 0);
      for (int j = 0; j < (nkey_tl + KVT_KEYS - 1) / KVT_KEYS; ++j)
        sm.sched[n++] = in.kv_tl_tc + ((((size_t)L * S + s) * Th + tl_t) * nT_tl + j) * tc::BLOCK_BYTES;
      sm.sched[n++] = wblk(w0, 3);
      sm.sched[n++] = wblk(w0, 4);
      sm.sched[n++] = wblk(w0, 5);
    }
    for (int L = 0; L < 3; ++L) {
      const int w0 = tbb::model_agent_interaction_transformer_layers_0_attn_in_proj_weight + 6 * L;
      sm.sched[n++] = wblk(w0, 1);  // Wk
      sm.sched[n++] = wblk(w0, 2);  // Wv
    }
    sm.n_stage = n;
  }
  if (warp == 0) tc::tmem_alloc(&sm.tmem_base, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm0 = sm.tmem_base;                              // MMA addresses (lane 0)
  const uint32_t tm = tm0 + ((uint32_t)(quad * 32) << 16);        // this warp's lanes

  // ring state (uniform in all threads; only thread 0 touches barriers / issues copies)
  uint32_t loaded = 0, consumed = 0;
  auto prefetch = [&]() {  // thread 0
    while (loaded < consumed + 2 && (int)loaded < sm.n_stage) {
      const uint32_t buf = loaded & 1;
      tc::mbar_expect_tx(&sm.bar_ring[buf], tc::BLOCK_BYTES);
      tc::bulk_g2s(sm.ring[buf], sm.sched[loaded], tc::BLOCK_BYTES, &sm.bar_ring[buf]);
      ++loaded;
    }
  };
  auto ring_wait = [&](uint32_t g) {  // thread 0: block g has landed
    tc::mbar_wait(&sm.bar_ring[g & 1], (g >> 1) & 1);
    tc::tc_fence_after();
  };
  uint32_t n_mma = 0, n_s[2] = {0, 0}, n_pv = 0;  // completed phases of bar_mma / bar_s / bar_pv (uniform)
  if (tid == 0) prefetch();

  const uint32_t idesc128 = tc::make_idesc_bf16(128, 128), idesc32 = tc::make_idesc_bf16(128, 32);

  // D[128 x 128] at TMEM column `dcol` = A(tmem, K = 128) W_block^T for ring block g      (thread 0)
  auto issue_gemm = [&](uint32_t g, uint32_t dcol) {
    ring_wait(g);
    const uint32_t wh = tc::smem_u32(sm.ring[g & 1]), wl = wh + 2 * tc::KB_BYTES_128;
#pragma unroll 1
    for (int term = 0; term < 3; ++term) {
      const uint32_t ta = tm0 + T_A + (term == 1 ? 64 : 0), wb = term == 2 ? wl : wh;
#pragma unroll
      for (int k = 0; k < 128; k += 16)
        tc::mma_bf16_ts(tm0 + dcol, ta + k / 2, tc::make_desc_sw128(wb + (k >> 6) * tc::KB_BYTES_128 + (k & 63) * 2), idesc128,
                        (term > 0 || k > 0) ? 1u : 0u);
    }
  };
  // A operand written -> `n` GEMMs with consecutive ring blocks into S0 (, S1) -> wait (all threads)
  auto run_gemm = [&](int n) {
    tc::tmem_st_wait();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      for (int j = 0; j < n; ++j) issue_gemm(consumed + j, j == 0 ? T_S0 : T_S1);
      tc::mma_commit(&sm.bar_mma);
    }
    tc::mbar_wait(&sm.bar_mma, n_mma & 1);
    tc::tc_fence_after();
    ++n_mma;
    consumed += n;
    if (tid == 0) prefetch();
  };

  // ---- per-thread helpers on the owned 64 columns ------------------------------------------------------------------------
  auto xs_at = [&](int c) -> float& { return sm.xs[c * 128 + r]; };
  // write 64 fp32 values (columns c0..c0+63 of this row) as the packed bf16 hi | lo A operand
  auto write_A = [&](const float (&v)[64]) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float ph[16], pl[16];
      tc::split32_packed(*reinterpret_cast<const float(*)[32]>(&v[32 * j]), ph, pl);
      tc::tmem_st16(tm + T_A + (c0 + 32 * j) / 2, ph);
      tc::tmem_st16(tm + T_A + 64 + (c0 + 32 * j) / 2, pl);
    }
  };
  // LayerNorm over the full row (both halves) of values held as v[64] per thread; result in place
  auto layernorm64 = [&](float (&v)[64], const float* __restrict__ g, const float* __restrict__ bt) {
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) sum += v[i];
    sm.red[half][r] = sum;
    __syncthreads();
    const float mean = (sm.red[0][r] + sm.red[1][r]) * (1.0f / 128);
    __syncthreads();
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      const float d = v[i] - mean;
      q = fmaf(d, d, q);
    }
    sm.red[half][r] = q;
    __syncthreads();
    const float rstd = 1.0f / sqrtf((sm.red[0][r] + sm.red[1][r]) * (1.0f / 128) + LN_EPS);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = (v[i] - mean) * rstd * __ldg(g + c0 + i) + __ldg(bt + c0 + i);
  };
  auto load_x = [&](float (&v)[64]) {
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = xs_at(c0 + i);
  };
  auto load_acc = [&](uint32_t col, float (&v)[64]) {  // this thread's 64 columns of a 128-column accumulator
    tc::tmem_ld32(tm + col + c0, *reinterpret_cast<float(*)[32]>(&v[0]));
    tc::tmem_ld32(tm + col + c0 + 32, *reinterpret_cast<float(*)[32]>(&v[32]));
    tc::tmem_ld_wait();
  };

  // ---- state embedding: get_agent_attr_and_pe + agent_encoder (sc_input.py:142-165, input_pe_encoder.py:41-61) -----------
  bool valid = false;
  {
    const uint8_t* valid_cur = a.sv.valid + (size_t)(t & 1) * BA;
    float v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = 0.f;
    if (live) valid = valid_cur[ba] != 0;
    if (valid) {
      const float4 st = *reinterpret_cast<const float4*>(a.sv.agent_state + ba * 4);
      const float* fxy = packed + tbw::pre_processing_input_pose_pe_agent_pe_xy_freqs;
      const float* fyaw = packed + tbw::pre_processing_input_pose_pe_agent_pe_yaw_freqs;
      if (half == 0) {
        float at[12];
        at[0] = a.sv.vel[ba * 2], at[1] = a.sv.vel[ba * 2 + 1], at[2] = st.w, at[3] = a.sv.yaw_rate[ba], at[4] = a.sv.acc[ba];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          at[5 + i] = in.agent_size[sa * 3 + i];
          at[8 + i] = in.agent_type[sa * 3 + i] ? 1.f : 0.f;
        }
        at[11] = 0.f;
        const float* w1 = packed + tbw::model_agent_encoder_mlp_fc_layers_0_weight;  // Wt4[3][32][4]
        const float* w2 = packed + tbw::model_agent_encoder_mlp_fc_layers_3_weight;  // Wt4[8][32][4]
        float h[32];
#pragma unroll
        for (int o = 0; o < 32; ++o) {
          float acc = __ldg(packed + tbw::model_agent_encoder_mlp_fc_layers_0_bias + o);
#pragma unroll
          for (int k = 0; k < 12; ++k) acc = fmaf(at[k], __ldg(w1 + ((k >> 2) * 32 + o) * 4 + (k & 3)), acc);
          h[o] = fmaxf(acc, 0.f);
        }
#pragma unroll
        for (int o = 0; o < 32; ++o) {
          float acc = __ldg(packed + tbw::model_agent_encoder_mlp_fc_layers_3_bias + o);
#pragma unroll
          for (int k = 0; k < 32; ++k) acc = fmaf(h[k], __ldg(w2 + ((k >> 2) * 32 + o) * 4 + (k & 3)), acc);
          v[o] = acc;
        }
        // PE columns 32..63: cos(x f_i) i<12, sin(x f_i) i<12, cos(y f_i) i<8
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          v[32 + i] = cosf(st.x * __ldg(fxy + 2 * i));
          v[44 + i] = sinf(st.x * __ldg(fxy + 2 * i + 1));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[56 + i] = cosf(st.y * __ldg(fxy + 2 * i));
      } else {
        // columns 64..127: cos(y f_i) i=8..11, sin(y f_i) i<12, cos(k yaw) k<24, sin(k yaw) k<24
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = cosf(st.y * __ldg(fxy + 2 * (8 + i)));
#pragma unroll
        for (int i = 0; i < 12; ++i) v[4 + i] = sinf(st.y * __ldg(fxy + 2 * i + 1));
#pragma unroll
        for (int i = 0; i < 24; ++i) {
          v[16 + i] = cosf(st.z * __ldg(fyaw + 2 * i));
          v[40 + i] = sinf(st.z * __ldg(fyaw + 2 * i + 1));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 64; ++i) xs_at(c0 + i) = v[i];
    if (half == 0) sm.row_valid[r] = valid;
  }
  __syncthreads();

  // ---- one pre-LN cross-attention layer against `nkey` compacted keys streamed as 64 KB blocks ---------------------------
  auto xlayer = [&](const float* __restrict__ lw, int nkey) {
    float v[64];
    // Q = LN1(x) Wq + bq   -> packed into the A region (stays there for every QK^T of this layer)
    load_x(v);
    layernorm64(v, lw + tfl::NORM1_W, lw + tfl::NORM1_B);
    write_A(v);
    run_gemm(1);
    load_acc(T_S0, v);
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] += __ldg(lw + tfl::IN_B + c0 + i);
    write_A(v);
    tc::tmem_st_wait();
    tc::tc_fence_before();
    __syncthreads();

    // ---- attention ------------------------------------------------------------------------------------------------------
    const int n_sub = (nkey + SUB_KEYS - 1) / SUB_KEYS;
    const float sc = 0.17677669529663687f * 1.4426950408889634f;  // 1/sqrt(32) * log2(e): softmax in base 2
    float m_ref[2] = {-INFINITY, -INFINITY}, l_sum[2] = {0.f, 0.f};
    const uint32_t blk0 = consumed;  // ring index of the first key block
    // QK^T of sub-tile u into S[u & 1]                                                       (thread 0)
    auto issue_qk = [&](int u) {
      const uint32_t g = blk0 + (u >> 1);
      if ((u & 1) == 0) ring_wait(g);
      const uint32_t kb = tc::smem_u32(sm.ring[g & 1]);
      const uint32_t sd = tm0 + ((u & 1) ? T_S1 : T_S0);
#pragma unroll 1
      for (int h = 0; h < NHEAD; ++h) {
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t ta = tm0 + T_A + (term == 1 ? 64 : 0) + 16 * h;
          const uint32_t kk = kb + (term == 2 ? 16384 : 0) + (h >> 1) * 8192 + (u & 1) * 4096 + (h & 1) * 64;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            tc::mma_bf16_ts(sd + 32 * h, ta + 8 * ks, tc::make_desc_sw128(kk + 32 * ks), idesc32, (term > 0 || ks > 0) ? 1u : 0u);
        }
      }
      tc::mma_commit(&sm.bar_s[u & 1]);
    };
    // O_h += P_h V_h for sub-tile u (P = packed hi | lo in S[u & 1])                           (thread 0)
    auto issue_pv = [&](int u) {
      const uint32_t g = blk0 + (u >> 1);
      const uint32_t vb = tc::smem_u32(sm.ring[g & 1]) + 32768;
      const uint32_t sp = tm0 + ((u & 1) ? T_S1 : T_S0);
#pragma unroll 1
      for (int h = 0; h < NHEAD; ++h) {
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t ta = sp + 32 * h + (term == 1 ? 16 : 0);
          const uint32_t vv = vb + (term == 2 ? 16384 : 0) + h * 4096 + (u & 1) * 64;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            tc::mma_bf16_ts(tm0 + T_O + 32 * h, ta + 8 * ks, tc::make_desc_sw128(vv + 32 * ks), idesc32,
                            (u > 0 || term > 0 || ks > 0) ? 1u : 0u);
        }
      }
      tc::mma_commit(&sm.bar_pv);
    };
    if (tid == 0 && n_sub > 0) {
      tc::tc_fence_after();
      issue_qk(0);
      if (n_sub > 1) issue_qk(1);
    }
#pragma unroll 1
    for (int u = 0; u < n_sub; ++u) {
      const int bsel = u & 1;
      tc::mbar_wait(&sm.bar_s[bsel], n_s[bsel] & 1);
      tc::tc_fence_after();
      ++n_s[bsel];
      const uint32_t sbase = tm + (bsel ? T_S1 : T_S0);
      const int key0 = u * SUB_KEYS;
      bool need_rescale = false;
      float alpha[2] = {1.f, 1.f};
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int h = 2 * half + hh;
        float sv_[32];
        tc::tmem_ld32(sbase + 32 * h, sv_);
        tc::tmem_ld_wait();
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          sv_[j] = (key0 + j < nkey) ? sv_[j] * sc : -INFINITY;
          mx = fmaxf(mx, sv_[j]);
        }
        // lazy rescaling: keep the reference maximum unless the new maximum exceeds it by more than 8 (factor 256)
        if (mx > m_ref[hh] + 8.0f) {
          alpha[hh] = (m_ref[hh] == -INFINITY) ? 0.f : exp2f(m_ref[hh] - mx);
          m_ref[hh] = mx;
          l_sum[hh] *= alpha[hh];
          need_rescale = need_rescale || (u > 0);
        }
        float psum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          sv_[j] = exp2f(sv_[j] - m_ref[hh]);  // masked keys: exp2(-inf) = 0
          psum += sv_[j];
        }
        l_sum[hh] += psum;
        float ph[16], pl[16];
        tc::split32_packed(sv_, ph, pl);
        tc::tmem_st16(sbase + 32 * h, ph);
        tc::tmem_st16(sbase + 32 * h + 16, pl);
      }
      // the previous PV must have finished before O is touched / before its P buffer is reused by QK^T(u+1)... (see below)
      if (u > 0) {
        tc::mbar_wait(&sm.bar_pv, n_pv & 1);
        tc::tc_fence_after();
        ++n_pv;
        if (((u - 1) & 1) == 1) {  // PV(u-1) was the second half of its key block: the ring slot is free again
          ++consumed;
          if (tid == 0) prefetch();
        }
      }
      if (__any_sync(0xffffffffu, need_rescale)) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int h = 2 * half + hh;
          float o[32];
          tc::tmem_ld32(tm + T_O + 32 * h, o);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] *= alpha[hh];
          tc::tmem_st32(tm + T_O + 32 * h, o);
        }
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc::tc_fence_after();
        issue_pv(u);
        if (u + 2 < n_sub) issue_qk(u + 2);  // executes after PV(u) on the tensor pipe: S[u & 1] is free by then
      }
    }
    if (n_sub > 0) {
      tc::mbar_wait(&sm.bar_pv, n_pv & 1);
      tc::tc_fence_after();
      ++n_pv;
      // blocks not yet released: the last one (and, if n_sub is even, it is exactly the last block)
      consumed = blk0 + (n_sub + 1) / 2;
      if (tid == 0) prefetch();
    }
    // ---- O / l -> A operand; out-proj; residual (rows without any key: zero attention output, attention.py:144-146) ---------
    {
      float o[64];
      if (n_sub > 0) {
        load_acc(T_O, o);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const float inv = l_sum[hh] > 0.f ? 1.0f / l_sum[hh] : 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) o[32 * hh + j] *= inv;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) o[i] = 0.f;
      }
      write_A(o);
    }
    run_gemm(1);
    load_acc(T_S0, v);
    {
      float x[64];
      load_x(x);
      if (nkey > 0) {
#pragma unroll
        for (int i = 0; i < 64; ++i) x[i] += v[i] + __ldg(lw + tfl::OUT_B + c0 + i);
      }
#pragma unroll
      for (int i = 0; i < 64; ++i) xs_at(c0 + i) = x[i];
      // FFN
      layernorm64(x, lw + tfl::NORM2_W, lw + tfl::NORM2_B);
      write_A(x);
    }
    run_gemm(1);
    load_acc(T_S0, v);
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = fmaxf(v[i] + __ldg(lw + tfl::L1_B + c0 + i), 0.f);
    write_A(v);
    run_gemm(1);
    load_acc(T_S0, v);
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      const float y = xs_at(c0 + i) + v[i] + __ldg(lw + tfl::L2_B + c0 + i);
      xs_at(c0 + i) = valid ? y : 0.f;
    }
  };

#pragma unroll 1
  for (int L = 0; L < 3; ++L) xlayer(packed + tbw::model_transformer_as2pl_layers_0_norm1_weight + L * tfl::STRIDE, nkey_map);
#pragma unroll 1
  for (int L = 0; L < 3; ++L) xlayer(packed + tbw::model_transformer_as2tl_layers_0_norm1_weight + L * tfl::STRIDE, nkey_tl);

  // ---- hand-over: x0 and the interaction K|V (LN_tgt(x0) Wkv + b) of every row ------------------------------------------------
  {
    float x[64];
    load_x(x);
    if (live) {
      float* dst = a.sv.x0 + ba * D + c0;
#pragma unroll
      for (int i = 0; i < 16; ++i) reinterpret_cast<float4*>(dst)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
    }
#pragma unroll 1
    for (int L = 0; L < 3; ++L) {
      const float* lw = packed + tbw::model_agent_interaction_transformer_layers_0_norm1_weight + L * tfl::STRIDE;
      float v[64];
      load_x(v);
      layernorm64(v, lw + tfl::NORMT_W, lw + tfl::NORMT_B);
      write_A(v);
      run_gemm(2);  // K -> S0, V -> S1
#pragma unroll 1
      for (int kvsel = 0; kvsel < 2; ++kvsel) {
        load_acc(kvsel ? T_S1 : T_S0, v);
        if (live) {
          float* dst = a.sv.kv_int + (((size_t)L * B + b) * A + r) * 256 + kvsel * D + c0;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            reinterpret_cast<float4*>(dst)[i] =
                make_float4(v[4 * i] + __ldg(lw + tfl::IN_B + D + kvsel * D + c0 + 4 * i),
                            v[4 * i + 1] + __ldg(lw + tfl::IN_B + D + kvsel * D + c0 + 4 * i + 1),
                            v[4 * i + 2] + __ldg(lw + tfl::IN_B + D + kvsel * D + c0 + 4 * i + 2),
                            v[4 * i + 3] + __ldg(lw + tfl::IN_B + D + kvsel * D + c0 + 4 * i + 3));
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(sm.tmem_base, 512);
}

}  // namespace tb

using namespace tb;

extern "C" size_t tb_kv_tc_bytes(const TbDims* d, int32_t which) {
  if (check_dims_host(d) != TB_OK) return 0;
  const size_t nT_map = (d->n_pl + KVT_KEYS - 1) / KVT_KEYS, nT_tl = (d->n_tl + KVT_KEYS - 1) / KVT_KEYS;
  if (which == 0) return (size_t)3 * d->n_scene * nT_map * tc::BLOCK_BYTES;
  if (which == 1) return (size_t)3 * d->n_scene * d->n_step_hist * nT_tl * tc::BLOCK_BYTES;
  return 0;
}

int tb::launch_pack_kv_tc(const float* kv, const uint8_t* key_valid, int n_set, int n_set_valid, int T, unsigned char* blocks,
                          int32_t* n_key, cudaStream_t st) {
  const int nT = (T + KVT_KEYS - 1) / KVT_KEYS;
  k_pack_kv_tc<<<n_set, 256, nT * KVT_KEYS * sizeof(int), st>>>(kv, key_valid, T, nT, blocks, n_key, n_set_valid);
  count_launch();
  return launch_status();
}

bool tb::front_tc_supported(const TbDims& d, const TbRolloutIn& in) {
  if (!in.kv_map_tc || !in.kv_tl_tc || !in.n_key_map || !in.n_key_tl) return false;
  if (d.n_agent > 128) return false;
  const int nT_map = (d.n_pl + KVT_KEYS - 1) / KVT_KEYS, nT_tl = (d.n_tl + KVT_KEYS - 1) / KVT_KEYS;
  return 3 * (4 + nT_map) + 3 * (4 + nT_tl) + 6 <= MAX_STAGE;
}

int tb::launch_step_front_tc(const TbDims& d, const TbRolloutIn& in, const float* packed, const StateView& sv, int t,
                             cudaStream_t st) {
  static bool attr_set = false;
  const int smem = (int)sizeof(FrontTcSmem) + 1024;
  if (!attr_set) {
    cudaFuncSetAttribute(k_step_front_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_set = true;
  }
  FrontArgs a{d, in, packed, tc_blob(packed), sv, t};
  k_step_front_tc<<<d.n_scene * d.n_mode, 256, smem, st>>>(a);
  count_launch();
  return launch_status();
}
