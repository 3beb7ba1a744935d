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
//   warps 0..7 : row workers (thread pair per agent row: LayerNorm, softmax, epilogues; they never issue MMAs)
//   warp  8    : issuer -- one lane streams the 64 KB blocks into the ring and issues every tcgen05.mma; it is told that
//                an operand is ready through `bar_ready` (8 arrivals = one per worker warp) and reports completion with
//                tcgen05.commit on bar_mma / bar_s / bar_pv; ring slots are recycled on bar_free (commit after the last
//                MMA that reads the slot).
// ------------------------------------------------------------------------------------------------------------
constexpr int MAX_STAGE = 192;
constexpr int FRONT_THREADS = 288;

struct FrontTcSmem {
  unsigned char ring[2][tc::BLOCK_BYTES];
  float xs[128 * 128];  // residual stream, [col][row]
  float red[2][128];    // LayerNorm partial sums of the two column halves
  const unsigned char* sched[MAX_STAGE];
  uint64_t bar_ring[2];  // block landed in slot
  uint64_t bar_free[2];  // MMAs reading the slot have completed
  uint64_t bar_ready;    // workers -> issuer: operand written (8 arrivals)
  uint64_t bar_mma;      // issuer -> workers: GEMM batch done
  uint64_t bar_s[2];     // QK^T into S buffer done
  uint64_t bar_pv;       // PV done
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
  long long* trace;  // development aid: clock64 marks of CTA 0 / worker thread 0 (NULL = off)
};

// MUFU.EX2 (max relative error 2^-22), without exp2f's denormal-range scaling
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(FRONT_THREADS, 1) k_step_front_tc(FrontArgs a) {
  extern __shared__ unsigned char smem_raw[];
  FrontTcSmem& sm = *reinterpret_cast<FrontTcSmem*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
  const TbDims& dm = a.dm;
  const TbRolloutIn& in = a.in;
  const float* __restrict__ packed = a.packed;
  const int A = dm.n_agent, K = dm.n_mode, S = dm.n_scene, B = S * K, Th = dm.n_step_hist;
  const int nT_map = (dm.n_pl + KVT_KEYS - 1) / KVT_KEYS, nT_tl = (dm.n_tl + KVT_KEYS - 1) / KVT_KEYS;
  const int b = blockIdx.x, s = b / K, t = a.t;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tl_t = min(t - 1, Th - 1);
  const int nkey_map = in.n_key_map[s], nkey_tl = in.n_key_tl[(size_t)s * Th + tl_t];

  // ---- set-up: barriers, TMEM, static schedule of the 64 KB blocks this step consumes --------------------------------
  if (tid == 0) {
    tc::mbar_init(&sm.bar_ring[0], 1);
    tc::mbar_init(&sm.bar_ring[1], 1);
    tc::mbar_init(&sm.bar_free[0], 1);
    tc::mbar_init(&sm.bar_free[1], 1);
    tc::mbar_init(&sm.bar_ready, 8);
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
  const uint32_t tm0 = sm.tmem_base;
  const uint32_t idesc128 = tc::make_idesc_bf16(128, 128), idesc32 = tc::make_idesc_bf16(128, 32);

  if (warp == 8) {
    // =====================================================================================================================
    // issuer
    // =====================================================================================================================
    if (lane == 0) {
      uint32_t loaded = 0, n_ready = 0;
      const int n_stage = sm.n_stage;
      auto ensure_loaded = [&](uint32_t upto) {  // copies of blocks [loaded, upto] issued (slot reuse waits for bar_free)
        while (loaded <= upto && (int)loaded < n_stage) {
          const uint32_t slot = loaded & 1;
          if (loaded >= 2) tc::mbar_wait(&sm.bar_free[slot], ((loaded >> 1) - 1) & 1);
          tc::mbar_expect_tx(&sm.bar_ring[slot], tc::BLOCK_BYTES);
          tc::bulk_g2s(sm.ring[slot], sm.sched[loaded], tc::BLOCK_BYTES, &sm.bar_ring[slot]);
          ++loaded;
        }
      };
      // Loads are issued right after the last reader of the slot's previous block has been ISSUED (ensure_loaded blocks on
      // that reader's completion only), so a block is always in flight two uses ahead and never waits on the workers.
      auto ring_wait = [&](uint32_t g) {
        ensure_loaded(g);
        tc::mbar_wait(&sm.bar_ring[g & 1], (g >> 1) & 1);
        tc::tc_fence_after();
      };
      auto wait_ready = [&]() {
        tc::mbar_wait(&sm.bar_ready, n_ready & 1);
        ++n_ready;
        tc::tc_fence_after();
      };
      auto issue_gemm = [&](uint32_t g, uint32_t dcol) {  // D[128x128] @ dcol = A(tmem) W_g^T, then free the slot
        ring_wait(g);
        const uint32_t wh = tc::smem_u32(sm.ring[g & 1]);
        const uint64_t dh = tc::make_desc_sw128(wh), dl = tc::make_desc_sw128(wh + 2 * tc::KB_BYTES_128);
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t ta = tm0 + T_A + (term == 1 ? 64 : 0);
          const uint64_t db = term == 2 ? dl : dh;
#pragma unroll
          for (int k = 0; k < 128; k += 16)
            tc::mma_bf16_ts(tm0 + dcol, ta + k / 2, db + (uint64_t)(((k >> 6) * tc::KB_BYTES_128 + (k & 63) * 2) >> 4), idesc128,
                            (term > 0 || k > 0) ? 1u : 0u);
        }
        tc::mma_commit(&sm.bar_free[g & 1]);
      };
      uint32_t g = 0;  // next ring block
      ensure_loaded(1);
      auto attention = [&](int nkey) {
        const int n_sub = (nkey + SUB_KEYS - 1) / SUB_KEYS;
        const uint32_t blk0 = g;
        auto issue_qk = [&](int u) {
          const uint32_t gb = blk0 + (u >> 1);
          if ((u & 1) == 0) ring_wait(gb);
          const uint32_t kb = tc::smem_u32(sm.ring[gb & 1]) + (u & 1) * 4096;
          const uint64_t dh = tc::make_desc_sw128(kb), dl = tc::make_desc_sw128(kb + 16384);
          const uint32_t sd = tm0 + ((u & 1) ? T_S1 : T_S0);
#pragma unroll
          for (int h = 0; h < NHEAD; ++h) {
#pragma unroll
            for (int term = 0; term < 3; ++term) {
              const uint32_t ta = tm0 + T_A + (term == 1 ? 64 : 0) + 16 * h;
              const uint64_t db = (term == 2 ? dl : dh) + (uint64_t)(((h >> 1) * 8192 + (h & 1) * 64) >> 4);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                tc::mma_bf16_ts(sd + 32 * h, ta + 8 * ks, db + (uint64_t)(2 * ks), idesc32, (term > 0 || ks > 0) ? 1u : 0u);
            }
          }
          tc::mma_commit(&sm.bar_s[u & 1]);
        };
        auto issue_pv = [&](int u) {
          const uint32_t gb = blk0 + (u >> 1);
          const uint32_t vb = tc::smem_u32(sm.ring[gb & 1]) + 32768 + (u & 1) * 64;
          const uint64_t dh = tc::make_desc_sw128(vb), dl = tc::make_desc_sw128(vb + 16384);
          const uint32_t sp = tm0 + ((u & 1) ? T_S1 : T_S0);
#pragma unroll
          for (int h = 0; h < NHEAD; ++h) {
#pragma unroll
            for (int term = 0; term < 3; ++term) {
              const uint32_t ta = sp + 32 * h + (term == 1 ? 16 : 0);
              const uint64_t db = (term == 2 ? dl : dh) + (uint64_t)((h * 4096) >> 4);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                tc::mma_bf16_ts(tm0 + T_O + 32 * h, ta + 8 * ks, db + (uint64_t)(2 * ks), idesc32,
                                (u > 0 || term > 0 || ks > 0) ? 1u : 0u);
            }
          }
          tc::mma_commit(&sm.bar_pv);
          if ((u & 1) == 1 || u == n_sub - 1) tc::mma_commit(&sm.bar_free[gb & 1]);  // last reader of this key block
        };
        if (n_sub > 0) {
          wait_ready();  // Q packed in the A region
          issue_qk(0);
          if (n_sub > 1) issue_qk(1);
        }
        for (int u = 0; u < n_sub; ++u) {
          wait_ready();  // P(u) written, O rescaled
          issue_pv(u);
          if (u + 2 < n_sub) issue_qk(u + 2);
          if ((u & 1) == 1 || u == n_sub - 1) ensure_loaded(blk0 + (u >> 1) + 2);  // refill the slot this block leaves
        }
        g = blk0 + (n_sub + 1) / 2;
      };
      auto gemm_batch = [&](int n) {
        wait_ready();
        for (int j = 0; j < n; ++j) issue_gemm(g + j, j == 0 ? T_S0 : T_S1);
        tc::mma_commit(&sm.bar_mma);
        g += n;
        ensure_loaded(g + 1);  // blocks g-n .. g-1 were just read: refill their slots (waits for these MMAs only)
      };
      for (int L = 0; L < 6; ++L) {
        gemm_batch(1);                         // Wq
        attention(L < 3 ? nkey_map : nkey_tl);
        gemm_batch(1);                         // Wo
        gemm_batch(1);                         // W1
        gemm_batch(1);                         // W2
      }
      for (int L = 0; L < 3; ++L) gemm_batch(2);  // interaction Wk, Wv
    }
  } else {
    // =====================================================================================================================
    // row workers
    // =====================================================================================================================
    const int quad = warp & 3, half = warp >> 2;  // TMEM lane quadrant / column half (heads 2*half, 2*half+1)
    const int r = quad * 32 + lane;               // agent row
    const int c0 = half * 64;                     // first owned column
    const size_t BA = (size_t)B * A;
    const bool live = r < A;
    const size_t ba = (size_t)b * A + (live ? r : 0), sa = (size_t)s * A + (live ? r : 0);
    const uint32_t tm = tm0 + ((uint32_t)(quad * 32) << 16);
    uint32_t n_mma = 0, n_s[2] = {0, 0}, n_pv = 0;
    int n_mark = 0;
    auto mark = [&]() {
      if (a.trace && b == 0 && tid == 0 && n_mark < 500) a.trace[n_mark++] = clock64();
    };
    mark();

    // operand written (tcgen05.st) -> tell the issuer
    auto signal_ready = [&]() {
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.bar_ready);
    };
    auto wait_gemm = [&]() {
      tc::mbar_wait(&sm.bar_mma, n_mma & 1);
      tc::tc_fence_after();
      ++n_mma;
    };
    auto xs_at = [&](int c) -> float& { return sm.xs[c * 128 + r]; };
    auto write_A = [&](const float (&v)[64]) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float ph[16], pl[16];
        tc::split32_packed(*reinterpret_cast<const float(*)[32]>(&v[32 * j]), ph, pl);
        tc::tmem_st16(tm + T_A + (c0 + 32 * j) / 2, ph);
        tc::tmem_st16(tm + T_A + 64 + (c0 + 32 * j) / 2, pl);
      }
    };
    auto layernorm64 = [&](float (&v)[64], const float* __restrict__ g, const float* __restrict__ bt) {
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 64; ++i) sum += v[i];
      sm.red[half][r] = sum;
      worker_sync();
      const float mean = (sm.red[0][r] + sm.red[1][r]) * (1.0f / 128);
      worker_sync();
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float d = v[i] - mean;
        q = fmaf(d, d, q);
      }
      sm.red[half][r] = q;
      worker_sync();
      const float rstd = 1.0f / sqrtf((sm.red[0][r] + sm.red[1][r]) * (1.0f / 128) + LN_EPS);
      worker_sync();
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = (v[i] - mean) * rstd * __ldg(g + c0 + i) + __ldg(bt + c0 + i);
    };
    auto load_x = [&](float (&v)[64]) {
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = xs_at(c0 + i);
    };
    auto load_acc = [&](uint32_t col, float (&v)[64]) {
      tc::tmem_ld32(tm + col + c0, *reinterpret_cast<float(*)[32]>(&v[0]));
      tc::tmem_ld32(tm + col + c0 + 32, *reinterpret_cast<float(*)[32]>(&v[32]));
      tc::tmem_ld_wait();
    };

    // ---- state embedding: get_agent_attr_and_pe + agent_encoder (sc_input.py:142-165, input_pe_encoder.py:41-61) ---------
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
#pragma unroll
          for (int i = 0; i < 12; ++i) {
            v[32 + i] = cosf(st.x * __ldg(fxy + 2 * i));
            v[44 + i] = sinf(st.x * __ldg(fxy + 2 * i + 1));
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) v[56 + i] = cosf(st.y * __ldg(fxy + 2 * i));
        } else {
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
    worker_sync();
    mark();

    // ---- one pre-LN cross-attention layer against `nkey` compacted keys ---------------------------------------------------
    auto xlayer = [&](const float* __restrict__ lw, int nkey) {
      float v[64];
      load_x(v);
      layernorm64(v, lw + tfl::NORM1_W, lw + tfl::NORM1_B);
      write_A(v);
      signal_ready();  // -> Wq
      mark();
      wait_gemm();
      mark();
      load_acc(T_S0, v);
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] += __ldg(lw + tfl::IN_B + c0 + i);
      write_A(v);
      const int n_sub = (nkey + SUB_KEYS - 1) / SUB_KEYS;
      if (n_sub > 0) signal_ready();  // Q ready -> QK^T(0), QK^T(1)
      mark();

      const float sc = 0.17677669529663687f * 1.4426950408889634f;  // 1/sqrt(32) * log2(e): softmax in base 2
      float m_ref[2] = {-INFINITY, -INFINITY}, l_sum[2] = {0.f, 0.f};
#pragma unroll 1
      for (int u = 0; u < n_sub; ++u) {
        const int bsel = u & 1;
        tc::mbar_wait(&sm.bar_s[bsel], n_s[bsel] & 1);
        tc::tc_fence_after();
        ++n_s[bsel];
        if (u < 8) mark();
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
          if (key0 + SUB_KEYS > nkey) {  // only the last sub-tile of a layer has padding keys
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (key0 + j >= nkey) sv_[j] = -INFINITY;
          }
          float mx = sv_[0];
#pragma unroll
          for (int j = 1; j < 32; ++j) mx = fmaxf(mx, sv_[j]);
          mx *= sc;  // sc > 0: scale after the max
          // lazy rescaling: keep the reference maximum unless the new maximum exceeds it by more than 8 (factor 256)
          if (mx > m_ref[hh] + 8.0f) {
            alpha[hh] = (m_ref[hh] == -INFINITY) ? 0.f : exp2f(m_ref[hh] - mx);
            m_ref[hh] = mx;
            l_sum[hh] *= alpha[hh];
            need_rescale = need_rescale || (u > 0);
          }
          float psum = 0.f;
          const float neg_m = -m_ref[hh];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            sv_[j] = ex2_approx(fmaf(sv_[j], sc, neg_m));  // masked keys: ex2(-inf) = 0
            psum += sv_[j];
          }
          l_sum[hh] += psum;
          float ph[16], pl[16];
          tc::split32_packed(sv_, ph, pl);
          tc::tmem_st16(sbase + 32 * h, ph);
          tc::tmem_st16(sbase + 32 * h + 16, pl);
        }
        if (u < 8) mark();
        if (u > 0) {  // PV(u-1) must be complete before O is rescaled
          tc::mbar_wait(&sm.bar_pv, n_pv & 1);
          tc::tc_fence_after();
          ++n_pv;
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
        if (u < 8) mark();
        signal_ready();  // -> PV(u), QK^T(u+2)
        if (u < 8) mark();
      }
      mark();
      {
        float o[64];
        if (n_sub > 0) {
          tc::mbar_wait(&sm.bar_pv, n_pv & 1);
          tc::tc_fence_after();
          ++n_pv;
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
      signal_ready();  // -> Wo
      mark();
      wait_gemm();
      mark();
      load_acc(T_S0, v);
      {
        float x[64];
        load_x(x);
        if (nkey > 0) {  // rows without any key: zero attention output (attention.py:144-146)
#pragma unroll
          for (int i = 0; i < 64; ++i) x[i] += v[i] + __ldg(lw + tfl::OUT_B + c0 + i);
        }
#pragma unroll
        for (int i = 0; i < 64; ++i) xs_at(c0 + i) = x[i];
        layernorm64(x, lw + tfl::NORM2_W, lw + tfl::NORM2_B);
        write_A(x);
      }
      signal_ready();  // -> W1
      mark();
      wait_gemm();
      mark();
      load_acc(T_S0, v);
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = fmaxf(v[i] + __ldg(lw + tfl::L1_B + c0 + i), 0.f);
      write_A(v);
      signal_ready();  // -> W2
      mark();
      wait_gemm();
      mark();
      load_acc(T_S0, v);
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float y = xs_at(c0 + i) + v[i] + __ldg(lw + tfl::L2_B + c0 + i);
        xs_at(c0 + i) = valid ? y : 0.f;
      }
      mark();
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
        signal_ready();  // -> Wk (S0), Wv (S1)
        wait_gemm();
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

long long* tb::g_debug_trace = nullptr;
extern "C" void tb_debug_set_trace(void* dev_ptr) { tb::g_debug_trace = reinterpret_cast<long long*>(dev_ptr); }

bool tb::front_tc_supported(const TbDims& d, const TbRolloutIn& in) {
  if (!in.kv_map_tc || !in.kv_tl_tc || !in.n_key_map || !in.n_key_tl) return false;
  if (d.n_agent > 128) return false;
  const int nT_map = (d.n_pl + KVT_KEYS - 1) / KVT_KEYS, nT_tl = (d.n_tl + KVT_KEYS - 1) / KVT_KEYS;
  return 3 * (4 + nT_map) + 3 * (4 + nT_tl) + 6 <= MAX_STAGE;
}

int tb::launch_step_front_tc(const TbDims& d, const TbRolloutIn& in, const float* packed, const StateView& sv, int t,
                             cudaStream_t st) {
  static std::atomic<uint64_t> attr_set{0};
  const int smem = (int)sizeof(FrontTcSmem) + 1024;
  if (!smem_attr_done(attr_set)) {
    if (!set_max_smem(k_step_front_tc, smem)) return TB_ERR_LAUNCH;
    smem_attr_mark(attr_set);
  }
  FrontArgs a{d, in, packed, tc_blob(packed), sv, t, g_debug_trace};
  k_step_front_tc<<<d.n_scene * d.n_mode, FRONT_THREADS, smem, st>>>(a);
  count_launch();
  return launch_status();
}
