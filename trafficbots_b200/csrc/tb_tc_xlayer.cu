// One pre-LN cross-attention layer (`TransformerCrossAttention.forward`, reference src/models/modules/transformer.py:186-237)
// on the tensor pipe for ANY number of query rows: a CTA owns a tile of up to 128 rows of one batch element and runs
//   LN1 -> Wq -> flash attention against the compacted tensor-core key blocks of the element -> Wo -> residual -> LN2 -> FFN.
// Used for the map encoder's global self-attention over the polylines of a scene (map_encoder.py:108-114) and for the
// latent encoder's agent->map / agent->traffic-light layers (latent_encoder.py:108-122).  Same machinery as the decode
// kernels: A operands in tensor memory (bf16x3), weights and key blocks streamed through a 2 x 64 KB ring by bulk-async
// copies, online softmax with lazy rescaling; the issuer warp runs converged and one elected lane issues (tc::elect_one).
//   TMEM: [0,128) S0 / GEMM accumulator | [128,256) S1 | [256,384) O | [384,512) A operand (hi | lo)
#include "tb_host.h"

namespace tb {
namespace xl {

constexpr int KVT_KEYS = 64, SUB_KEYS = 32;
constexpr int MAX_STAGE = 160;
constexpr int THREADS = 288;
constexpr uint32_t T_S0 = 0, T_S1 = 128, T_O = 256, T_A = 384;

struct Smem {
  unsigned char ring[2][tc::BLOCK_BYTES];
  float xs[128 * 128];  // residual stream, [col][row]
  float red[2][128];
  const unsigned char* sched[MAX_STAGE];
  uint64_t bar_ring[2], bar_free[2], bar_ready, bar_mma, bar_s[2], bar_pv;
  uint32_t tmem_base;
  int n_stage;
};

struct Args {
  const float* src;            // [n_batch, n_src, 128]
  const uint8_t* src_valid;    // [n_batch, n_src]
  float* dst;
  int n_src;
  const unsigned char* blocks; // [n_batch / kv_share][nT] x 64 KB
  const int32_t* n_key;        // [n_batch / kv_share]
  int nT, kv_share;
  const float* lw;             // fp32 parameters of the layer (tfl:: offsets)
  const unsigned char* tcw;    // tensor-core weight blocks of the layer: q, k, v, out, linear1, linear2
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) k_xlayer_tc(Args a) {
  extern __shared__ unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
  const int b = blockIdx.y, r0 = blockIdx.x * 128;
  const int kb = b / a.kv_share;
  const int tid = threadIdx.x, warp = tc::uniform(tid >> 5), lane = tid & 31;
  const int nkey = tc::uniform(a.n_key[kb]);
  const int nblk = (nkey + KVT_KEYS - 1) / KVT_KEYS;

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
    sm.sched[n++] = a.tcw;  // Wq
    for (int j = 0; j < nblk; ++j) sm.sched[n++] = a.blocks + ((size_t)kb * a.nT + j) * tc::BLOCK_BYTES;
    sm.sched[n++] = a.tcw + 3 * (size_t)tc::BLOCK_BYTES;  // Wo
    sm.sched[n++] = a.tcw + 4 * (size_t)tc::BLOCK_BYTES;  // W1
    sm.sched[n++] = a.tcw + 5 * (size_t)tc::BLOCK_BYTES;  // W2
    sm.n_stage = n;
  }
  if (warp == 0) tc::tmem_alloc(&sm.tmem_base, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm0 = (uint32_t)tc::uniform((int)sm.tmem_base);
  const uint32_t idesc128 = tc::make_idesc_bf16(128, 128), idesc32 = tc::make_idesc_bf16(128, 32);

  if (warp == 8) {
    // ================================================================================================ issuer (converged warp)
    uint32_t loaded = 0, n_ready = 0;
    const int n_stage = tc::uniform(sm.n_stage);
    auto ensure_loaded = [&](uint32_t upto) {
      while (loaded <= upto && (int)loaded < n_stage) {
        const uint32_t slot = loaded & 1;
        if (loaded >= 2) tc::mbar_wait(&sm.bar_free[slot], ((loaded >> 1) - 1) & 1);
        if (tc::elect_one()) {
          tc::mbar_expect_tx(&sm.bar_ring[slot], tc::BLOCK_BYTES);
          tc::bulk_g2s(sm.ring[slot], sm.sched[loaded], tc::BLOCK_BYTES, &sm.bar_ring[slot]);
        }
        __syncwarp();
        ++loaded;
      }
    };
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
    uint32_t g = 0;
    ensure_loaded(1);
    auto gemm = [&]() {  // ACC0 = A(tmem) W_g^T
      wait_ready();
      ring_wait(g);
      const uint32_t wh = tc::smem_u32(sm.ring[g & 1]);
      const uint64_t dh = tc::make_desc_sw128(wh), dl = tc::make_desc_sw128(wh + 2 * tc::KB_BYTES_128);
      if (tc::elect_one()) {
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t ta = tm0 + T_A + (term == 1 ? 64 : 0);
          const uint64_t db = term == 2 ? dl : dh;
#pragma unroll
          for (int k = 0; k < 128; k += 16)
            tc::mma_bf16_ts(tm0 + T_S0, ta + k / 2, db + (uint64_t)(((k >> 6) * tc::KB_BYTES_128 + (k & 63) * 2) >> 4), idesc128,
                            (term > 0 || k > 0) ? 1u : 0u);
        }
        tc::mma_commit(&sm.bar_free[g & 1]);
        tc::mma_commit(&sm.bar_mma);
      }
      __syncwarp();
      ++g;
      ensure_loaded(g + 1);
    };
    auto attention = [&]() {
      const int n_sub = (nkey + SUB_KEYS - 1) / SUB_KEYS;
      const uint32_t blk0 = g;
      auto issue_qk = [&](int u) {
        const uint32_t gb = blk0 + (u >> 1);
        if ((u & 1) == 0) ring_wait(gb);
        const uint32_t kbase = tc::smem_u32(sm.ring[gb & 1]) + (u & 1) * 4096;
        const uint64_t dh = tc::make_desc_sw128(kbase), dl = tc::make_desc_sw128(kbase + 16384);
        const uint32_t sd = tm0 + ((u & 1) ? T_S1 : T_S0);
        if (tc::elect_one()) {
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
        }
        __syncwarp();
      };
      auto issue_pv = [&](int u) {
        const uint32_t gb = blk0 + (u >> 1);
        const uint32_t vb = tc::smem_u32(sm.ring[gb & 1]) + 32768 + (u & 1) * 64;
        const uint64_t dh = tc::make_desc_sw128(vb), dl = tc::make_desc_sw128(vb + 16384);
        const uint32_t sp = tm0 + ((u & 1) ? T_S1 : T_S0);
        const bool last_of_block = (u & 1) == 1 || u == n_sub - 1;
        if (tc::elect_one()) {
#pragma unroll
          for (int h = 0; h < NHEAD; ++h) {
#pragma unroll
            for (int term = 0; term < 3; ++term) {
              const uint32_t ta = sp + 32 * h + (term == 1 ? 16 : 0);
              const uint64_t db = (term == 2 ? dl : dh) + (uint64_t)((h * 4096) >> 4);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                tc::mma_bf16_ts(tm0 + T_O + 32 * h, ta + 8 * ks, db + (uint64_t)(2 * ks), idesc32, (u > 0 || term > 0 || ks > 0) ? 1u : 0u);
            }
          }
          tc::mma_commit(&sm.bar_pv);
          if (last_of_block) tc::mma_commit(&sm.bar_free[gb & 1]);
        }
        __syncwarp();
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
        if ((u & 1) == 1 || u == n_sub - 1) ensure_loaded(blk0 + (u >> 1) + 2);
      }
      g = blk0 + (n_sub + 1) / 2;
    };
    gemm();       // Wq
    attention();
    gemm();       // Wo
    gemm();       // W1
    gemm();       // W2
  } else {
    // ================================================================================================ row workers
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane;
    const int c0 = half * 64;
    const bool live = r0 + r < a.n_src;
    const size_t row = (size_t)b * a.n_src + r0 + (live ? r : 0);
    const uint32_t tm = tm0 + ((uint32_t)(quad * 32) << 16);
    uint32_t n_mma = 0, n_s[2] = {0, 0}, n_pv = 0;
    const float* __restrict__ lw = a.lw;

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
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 64; ++i) s4[i & 3] += v[i];
      sm.red[half][r] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      worker_sync();
      const float mean = (sm.red[0][r] + sm.red[1][r]) * (1.0f / 128);
      worker_sync();
      float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float d = v[i] - mean;
        q4[i & 3] = fmaf(d, d, q4[i & 3]);
      }
      sm.red[half][r] = (q4[0] + q4[1]) + (q4[2] + q4[3]);
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

    const bool valid = live && a.src_valid[row] != 0;
    {
      float v[64];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) q = __ldg(reinterpret_cast<const float4*>(a.src + row * D + c0) + i);
        v[4 * i] = q.x, v[4 * i + 1] = q.y, v[4 * i + 2] = q.z, v[4 * i + 3] = q.w;
      }
#pragma unroll
      for (int i = 0; i < 64; ++i) xs_at(c0 + i) = v[i];
      layernorm64(v, lw + tfl::NORM1_W, lw + tfl::NORM1_B);
      write_A(v);
      signal_ready();  // -> Wq
      wait_gemm();
      load_acc(T_S0, v);
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] += __ldg(lw + tfl::IN_B + c0 + i);
      write_A(v);
    }
    const int n_sub = (nkey + SUB_KEYS - 1) / SUB_KEYS;
    if (n_sub > 0) signal_ready();  // Q ready
    const float sc = 0.17677669529663687f * 1.4426950408889634f;
    float m_ref[2] = {-INFINITY, -INFINITY}, l_sum[2] = {0.f, 0.f};
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
        if (key0 + SUB_KEYS > nkey) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (key0 + j >= nkey) sv_[j] = -INFINITY;
        }
        float mx = sv_[0];
#pragma unroll
        for (int j = 1; j < 32; ++j) mx = fmaxf(mx, sv_[j]);
        mx *= sc;
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
          sv_[j] = ex2_approx(fmaf(sv_[j], sc, neg_m));
          psum += sv_[j];
        }
        l_sum[hh] += psum;
        float ph[16], pl[16];
        tc::split32_packed(sv_, ph, pl);
        tc::tmem_st16(sbase + 32 * h, ph);
        tc::tmem_st16(sbase + 32 * h + 16, pl);
      }
      if (u > 0) {
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
      signal_ready();
    }
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
    wait_gemm();
    {
      float v[64], x[64];
      load_acc(T_S0, v);
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
    wait_gemm();
    {
      float v[64];
      load_acc(T_S0, v);
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = fmaxf(v[i] + __ldg(lw + tfl::L1_B + c0 + i), 0.f);
      write_A(v);
    }
    signal_ready();  // -> W2
    wait_gemm();
    {
      float v[64];
      load_acc(T_S0, v);
      if (live) {
        float* out = a.dst + row * D + c0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float4 q;
          q.x = valid ? xs_at(c0 + 4 * i) + v[4 * i] + __ldg(lw + tfl::L2_B + c0 + 4 * i) : 0.f;
          q.y = valid ? xs_at(c0 + 4 * i + 1) + v[4 * i + 1] + __ldg(lw + tfl::L2_B + c0 + 4 * i + 1) : 0.f;
          q.z = valid ? xs_at(c0 + 4 * i + 2) + v[4 * i + 2] + __ldg(lw + tfl::L2_B + c0 + 4 * i + 2) : 0.f;
          q.w = valid ? xs_at(c0 + 4 * i + 3) + v[4 * i + 3] + __ldg(lw + tfl::L2_B + c0 + 4 * i + 3) : 0.f;
          reinterpret_cast<float4*>(out)[i] = q;
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(sm.tmem_base, 512);
}

}  // namespace xl
}  // namespace tb

using namespace tb;

static int tc_layer_first_block(int block, int layer) {
  switch (block) {
    case TB_BLOCK_MAP_SELF_ATTN: return tbb::model_map_encoder_transformer_self_attn_layers_0_attn_in_proj_weight;
    case TB_BLOCK_AS2PL: return tbb::model_transformer_as2pl_layers_0_attn_in_proj_weight + 6 * layer;
    case TB_BLOCK_AS2TL: return tbb::model_transformer_as2tl_layers_0_attn_in_proj_weight + 6 * layer;
    case TB_BLOCK_INTERACTION: return tbb::model_agent_interaction_transformer_layers_0_attn_in_proj_weight + 6 * layer;
  }
  return -1;
}

int tb::launch_xlayer_tc(int block, int layer, const float* src, const uint8_t* src_valid, int n_batch, int n_src,
                         const unsigned char* blocks, const int32_t* n_key, int n_key_max, int kv_share, const float* packed, float* dst,
                         cudaStream_t st) {
  const int first = tc_layer_first_block(block, layer);
  if (first < 0) return TB_ERR_UNSUPPORTED;
  const int nT = (n_key_max + xl::KVT_KEYS - 1) / xl::KVT_KEYS;
  if (nT + 4 > xl::MAX_STAGE) return TB_ERR_BAD_SHAPE;
  static std::atomic<uint64_t> attr_set{0};
  const int smem = (int)sizeof(xl::Smem) + 1024;
  if (!smem_attr_done(attr_set)) {
    if (!set_max_smem(xl::k_xlayer_tc, smem)) return TB_ERR_LAUNCH;
    smem_attr_mark(attr_set);
  }
  xl::Args a{src, src_valid, dst, n_src, blocks, n_key, nT, kv_share, packed + block_base(block) + layer * tfl::STRIDE,
             tc_blob(packed) + (size_t)first * tc::BLOCK_BYTES};
  dim3 grid((n_src + 127) / 128, n_batch);
  xl::k_xlayer_tc<<<grid, xl::THREADS, smem, st>>>(a);
  count_launch();
  return launch_status();
}

extern "C" int32_t tb_xlayer_tc(int32_t block, int32_t layer, const float* src, const uint8_t* src_valid, int32_t n_batch,
                                int32_t n_src, const uint8_t* key_blocks, const int32_t* n_key, int32_t n_key_max,
                                int32_t kv_share, const float* packed, float* dst, void* stream) {
  if (!src || !src_valid || !key_blocks || !n_key || !packed || !dst) return TB_ERR_NULL;
  if (block_base(block) < 0 || layer < 0 || layer >= block_layers(block) || n_batch < 1 || n_batch > 65535 || n_src < 1 ||
      n_key_max < 1 || kv_share < 1 || n_batch % kv_share != 0)
    return TB_ERR_BAD_SHAPE;
  if (!aligned16(src) || !aligned16(packed) || !aligned16(dst) || (reinterpret_cast<uintptr_t>(key_blocks) & 127u)) return TB_ERR_ALIGN;
  return launch_xlayer_tc(block, layer, src, src_valid, n_batch, n_src, key_blocks, n_key, n_key_max, kv_share, packed, dst,
                          (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------------
// Pairwise destination MLP on the tensor pipe (DestPredictor, mode mlp: reference src/models/goal_manager.py:301-307).
// A tile = 128 consecutive polylines of one (scene, agent) pair; persistent CTAs, the 128x128 weight block of the second
// Linear stays resident in shared memory:   h1 = relu(LN1(U[p] + V[a]))  -> [128 x 128] bf16x3 GEMM -> relu(LN2(. + b3)) . w6 + b6
// ------------------------------------------------------------------------------------------------------------
namespace tb {
namespace dp {

constexpr int THREADS = 288;
struct Smem {
  unsigned char w[tc::BLOCK_BYTES];
  float red[2][128];
  float dot[2][128];
  uint64_t bar_w, bar_ready, bar_mma;
  uint32_t tmem_base;
};
constexpr uint32_t T_ACC = 0, T_A = 128;

__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// Admissible destinations: the reference masks every (agent, polyline) pair whose polyline is invalid / not a lane or road edge
// (types 0-4) or does not fit the agent type (goal_manager.py:233-244,328-330) to -inf AFTER evaluating the pairwise MLP on all
// of them; ~80 % of the pairs are masked.  k_dest_lists compacts, per scene and agent class (0 vehicle, 1 pedestrian, 2 cyclist,
// 3 no type bit), the indices of the admissible polylines; the pair kernel only evaluates those (k_dest_finish never reads a
// masked logit), and skips agents without a valid history altogether (their row becomes uniform, :331).
__global__ void __launch_bounds__(256) k_dest_lists(int P, const uint8_t* __restrict__ map_valid, const uint8_t* __restrict__ map_type,
                                                    int32_t* __restrict__ lists, int32_t* __restrict__ counts) {
  const int s = blockIdx.x, c = blockIdx.y;
  __shared__ int wsum[8];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t* out = lists + ((size_t)s * 4 + c) * P;
  for (int p0 = 0; p0 < P; p0 += 256) {
    const int p = p0 + threadIdx.x;
    bool ok = false;
    if (p < P && map_valid[(size_t)s * P + p]) {
      const uint8_t* mt = map_type + ((size_t)s * P + p) * TB_PL_TYPE;
      const bool t012 = mt[0] | mt[1] | mt[2], t3 = mt[3], t4 = mt[4];
      // type_mask keeps types 0-4; vehicles exclude type 3, pedestrians 0-3, cyclists 0-2 (the masks combine as written in the
      // reference even if a polyline had several type bits set)
      const bool any = t012 || t3 || t4;
      ok = c == 0 ? (any && !t3) : c == 1 ? (t4 && !t012 && !t3) : c == 2 ? ((t3 || t4) && !t012) : any;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) wsum[w] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int k = 0; k < w; ++k) off += wsum[k];
    if (ok) out[off + __popc(bal & ((1u << lane) - 1u))] = p;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int k = 0; k < 8; ++k) t += wsum[k];
      base += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[s * 4 + c] = base;
}

__global__ void __launch_bounds__(THREADS, 2) k_dest_pairs_tc(const float* __restrict__ U, const float* __restrict__ V, int P, int A,
                                                              int n_sa, const float* __restrict__ packed,
                                                              const unsigned char* __restrict__ w3_block, float* __restrict__ logits,
                                                              const int32_t* __restrict__ lists, const int32_t* __restrict__ counts,
                                                              const uint8_t* __restrict__ agent_type, const uint8_t* __restrict__ agent_valid) {
  extern __shared__ unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tc::uniform(tid >> 5), lane = tid & 31;
  const int tiles_per_sa = (P + 127) / 128;
  const int n_tile = n_sa * tiles_per_sa;
  if (tid == 0) {
    tc::mbar_init(&sm.bar_w, 1);
    tc::mbar_init(&sm.bar_ready, 8);
    tc::mbar_init(&sm.bar_mma, 1);
    tc::fence_mbar_init();
    tc::mbar_expect_tx(&sm.bar_w, tc::BLOCK_BYTES);
    tc::bulk_g2s(sm.w, w3_block, tc::BLOCK_BYTES, &sm.bar_w);
  }
  if (warp == 0) tc::tmem_alloc(&sm.tmem_base, 256);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm0 = (uint32_t)tc::uniform((int)sm.tmem_base);
  // tile = (scene-agent sa, j-th block of 128 admissible polylines of its class); both roles skip the same dead tiles
  auto agent_class = [&](int sa) {
    const uint8_t* at = agent_type + (size_t)sa * 3;
    return at[0] ? 0 : at[1] ? 1 : at[2] ? 2 : 3;
  };
  auto tile_live = [&](int tile) {
    // block-major tile order (tile = j * n_sa + sa): with the agent-major order a CTA striding by gridDim.x = 8 x 37 would
    // always get the same block index j, i.e. either only live or only dead tiles
    const int sa = tile % n_sa, j = tile / n_sa;
    if (!agent_valid[sa]) return false;
    return j * 128 < counts[(sa / A) * 4 + agent_class(sa)];
  };

  if (warp == 8) {
    tc::mbar_wait(&sm.bar_w, 0);
    const uint32_t wh = tc::smem_u32(sm.w);
    const uint64_t dh = tc::make_desc_sw128(wh), dl = tc::make_desc_sw128(wh + 2 * tc::KB_BYTES_128);
    const uint32_t idesc = tc::make_idesc_bf16(128, 128);
    uint32_t n_ready = 0;
    for (int tile = blockIdx.x; tile < n_tile; tile += gridDim.x) {
      if (!tile_live(tile)) continue;
      tc::mbar_wait(&sm.bar_ready, n_ready & 1);
      ++n_ready;
      tc::tc_fence_after();
      if (tc::elect_one()) {
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t ta = tm0 + T_A + (term == 1 ? 64 : 0);
          const uint64_t db = term == 2 ? dl : dh;
#pragma unroll
          for (int k = 0; k < 128; k += 16)
            tc::mma_bf16_ts(tm0 + T_ACC, ta + k / 2, db + (uint64_t)(((k >> 6) * tc::KB_BYTES_128 + (k & 63) * 2) >> 4), idesc,
                            (term > 0 || k > 0) ? 1u : 0u);
        }
        tc::mma_commit(&sm.bar_mma);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane, c0 = half * 64;
    const uint32_t tm = tm0 + ((uint32_t)(quad * 32) << 16);
    uint32_t n_mma = 0;
    const float* ln1w = packed + tbw::model_goal_manager_goal_predictor_mlp_fc_layers_1_weight;
    const float* ln1b = packed + tbw::model_goal_manager_goal_predictor_mlp_fc_layers_1_bias;
    const float* b3 = packed + tbw::model_goal_manager_goal_predictor_mlp_fc_layers_3_bias;
    const float* ln2w = packed + tbw::model_goal_manager_goal_predictor_mlp_fc_layers_4_weight;
    const float* ln2b = packed + tbw::model_goal_manager_goal_predictor_mlp_fc_layers_4_bias;
    const float* w6 = packed + tbw::model_goal_manager_goal_predictor_mlp_fc_layers_6_weight;
    const float b6 = __ldg(packed + tbw::model_goal_manager_goal_predictor_mlp_fc_layers_6_bias);
    // LayerNorm statistics of a row held by the thread pair (r, half 0 / 1): exchange of {sum, M2} halves (Chan)
    auto ln_stats = [&](const float (&v)[64], float& mean, float& rstd) {
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 64; ++i) s4[i & 3] += v[i];
      const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      const float mloc = sum * (1.0f / 64);
      float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float d = v[i] - mloc;
        q4[i & 3] = fmaf(d, d, q4[i & 3]);
      }
      const float m2 = (q4[0] + q4[1]) + (q4[2] + q4[3]);
      sm.red[half][r] = sum;
      sm.dot[half][r] = m2;
      worker_sync();
      const float os = sm.red[half ^ 1][r], om = sm.dot[half ^ 1][r];
      worker_sync();
      mean = (sum + os) * (1.0f / 128);
      const float dm = (os - sum) * (1.0f / 64);
      rstd = 1.0f / sqrtf((m2 + om + dm * dm * 32.0f) * (1.0f / 128) + LN_EPS);
    };
    for (int tile = blockIdx.x; tile < n_tile; tile += gridDim.x) {
      if (!tile_live(tile)) continue;
      const int sa = tile % n_sa, j0 = (tile / n_sa) * 128;
      const int s = sa / A, cls = agent_class(sa);
      const bool live = j0 + r < counts[s * 4 + cls];
      const int p = live ? lists[((size_t)s * 4 + cls) * P + j0 + r] : 0;
      float v[64];
      {
        const float4* u4 = reinterpret_cast<const float4*>(U + ((size_t)s * P + (live ? p : 0)) * D + c0);
        const float4* v4 = reinterpret_cast<const float4*>(V + (size_t)sa * D + c0);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 a4 = __ldg(u4 + i), b4 = __ldg(v4 + i);
          v[4 * i] = a4.x + b4.x, v[4 * i + 1] = a4.y + b4.y, v[4 * i + 2] = a4.z + b4.z, v[4 * i + 3] = a4.w + b4.w;
        }
      }
      float mean, rstd;
      ln_stats(v, mean, rstd);
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = fmaxf((v[i] - mean) * rstd * __ldg(ln1w + c0 + i) + __ldg(ln1b + c0 + i), 0.f);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float ph[16], pl[16];
        tc::split32_packed(*reinterpret_cast<const float(*)[32]>(&v[32 * j]), ph, pl);
        tc::tmem_st16(tm + T_A + (c0 + 32 * j) / 2, ph);
        tc::tmem_st16(tm + T_A + 64 + (c0 + 32 * j) / 2, pl);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.bar_ready);
      tc::mbar_wait(&sm.bar_mma, n_mma & 1);
      ++n_mma;
      tc::tc_fence_after();
      tc::tmem_ld32(tm + T_ACC + c0, *reinterpret_cast<float(*)[32]>(&v[0]));
      tc::tmem_ld32(tm + T_ACC + c0 + 32, *reinterpret_cast<float(*)[32]>(&v[32]));
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] += __ldg(b3 + c0 + i);
      ln_stats(v, mean, rstd);
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 64; ++i)
        d = fmaf(fmaxf((v[i] - mean) * rstd * __ldg(ln2w + c0 + i) + __ldg(ln2b + c0 + i), 0.f), __ldg(w6 + c0 + i), d);
      sm.dot[half][r] = d;
      worker_sync();
      if (half == 0 && live) logits[(size_t)sa * P + p] = sm.dot[0][r] + sm.dot[1][r] + b6;
      worker_sync();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(sm.tmem_base, 256);
}

}  // namespace dp
}  // namespace tb

size_t tb::dest_lists_bytes(int n_scene, int n_pl) { return (((size_t)n_scene * 4 * n_pl + (size_t)n_scene * 4) * sizeof(int32_t) + 255) & ~(size_t)255; }

int tb::launch_dest_pairs_tc(const float* U, const float* V, int n_scene, int n_agent, int n_pl, const float* packed, float* logits,
                             const uint8_t* map_valid, const uint8_t* map_type, const uint8_t* agent_type, const uint8_t* agent_valid,
                             int32_t* lists_ws, cudaStream_t st) {
  static std::atomic<uint64_t> attr_set{0};
  const int smem = (int)sizeof(dp::Smem) + 1024;
  if (!smem_attr_done(attr_set)) {
    if (!set_max_smem(dp::k_dest_pairs_tc, smem)) return TB_ERR_LAUNCH;
    smem_attr_mark(attr_set);
  }
  const int n_sa = n_scene * n_agent;
  const long n_tile = (long)n_sa * ((n_pl + 127) / 128);
  const int grid = (int)(n_tile < 2 * 148 ? n_tile : 2 * 148);  // 2 CTAs per SM (66 KB shared memory, 256 TMEM columns each)
  int32_t* counts = lists_ws + (size_t)n_scene * 4 * n_pl;
  dp::k_dest_lists<<<dim3(n_scene, 4), 256, 0, st>>>(n_pl, map_valid, map_type, lists_ws, counts);
  count_launch();
  dp::k_dest_pairs_tc<<<grid, dp::THREADS, smem, st>>>(
      U, V, n_pl, n_agent, n_sa, packed,
      tc_blob(packed) + (size_t)tbb::model_goal_manager_goal_predictor_mlp_fc_layers_3_weight * tc::BLOCK_BYTES, logits, lists_ws, counts,
      agent_type, agent_valid);
  count_launch();
  return launch_status();
}

// ------------------------------------------------------------------------------------------------------------
// K|V projection on the tensor pipe: kv[row] = LN_tgt(tgt[row]) W_kv^T + b_kv  (transformer.py:192 + attention.py:86).
// Persistent CTAs of 512 threads (4 threads per row, 128-row tiles); the two 128x128 weight blocks (Wk, Wv) stay resident in
// shared memory; A operand in tensor memory, two accumulators.
// ------------------------------------------------------------------------------------------------------------
namespace tb {
namespace kvp {

constexpr int THREADS = 512;
constexpr uint32_t T_K = 0, T_V = 128, T_A = 256;
struct Smem {
  unsigned char w[2][tc::BLOCK_BYTES];
  float2 red[2][4][128];
  float lp[4][128];  // norm_tgt weight, bias, K bias, V bias
  uint64_t bar_w, bar_mma;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(THREADS, 1) k_kv_project_tc(const float* __restrict__ tgt, long n_row, const float* __restrict__ lw,
                                                              const unsigned char* __restrict__ wkv_blocks, float* __restrict__ kv) {
  extern __shared__ unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tc::uniform(tid >> 5), lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2, r = quad * 32 + lane, cq = 32 * part;
  const long n_tile = (n_row + 127) / 128;
  if (tid == 0) {
    tc::mbar_init(&sm.bar_w, 1);
    tc::mbar_init(&sm.bar_mma, 1);
    tc::fence_mbar_init();
    tc::mbar_expect_tx(&sm.bar_w, 2 * tc::BLOCK_BYTES);
    tc::bulk_g2s(sm.w[0], wkv_blocks, tc::BLOCK_BYTES, &sm.bar_w);                    // Wk = block 1 of in_proj
    tc::bulk_g2s(sm.w[1], wkv_blocks + tc::BLOCK_BYTES, tc::BLOCK_BYTES, &sm.bar_w);  // Wv = block 2
  }
  if (warp == 0) tc::tmem_alloc(&sm.tmem_base, 512);
  {
    const int off[4] = {tfl::NORMT_W, tfl::NORMT_B, tfl::IN_B + 128, tfl::IN_B + 256};
    sm.lp[tid >> 7][tid & 127] = __ldg(lw + off[tid >> 7] + (tid & 127));
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm0 = (uint32_t)tc::uniform((int)sm.tmem_base);
  const uint32_t tm = tm0 + ((uint32_t)(quad * 32) << 16);
  const uint32_t idesc = tc::make_idesc_bf16(128, 128);
  uint32_t n_mma = 0, n_ln = 0;
  bool w_ready = false;
  for (long tile = blockIdx.x; tile < n_tile; tile += gridDim.x) {
    const long row = tile * 128 + r;
    const bool live = row < n_row;
    float v[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) q = __ldg(reinterpret_cast<const float4*>(tgt + row * D + cq) + i);
      v[4 * i] = q.x, v[4 * i + 1] = q.y, v[4 * i + 2] = q.z, v[4 * i + 3] = q.w;
    }
    {  // LayerNorm over the row held by 4 threads
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; ++i) s4[i & 3] += v[i];
      const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      const float mloc = sum * (1.0f / 32);
      float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float d = v[i] - mloc;
        q[i & 3] = fmaf(d, d, q[i & 3]);
      }
      const int buf = n_ln & 1;
      ++n_ln;
      sm.red[buf][part][r] = make_float2(sum, (q[0] + q[1]) + (q[2] + q[3]));
      __syncthreads();
      const float2 p0 = sm.red[buf][0][r], p1 = sm.red[buf][1][r], p2 = sm.red[buf][2][r], p3 = sm.red[buf][3][r];
      const float mean = ((p0.x + p1.x) + (p2.x + p3.x)) * (1.0f / 128);
      const float d0 = p0.x * (1.0f / 32) - mean, d1 = p1.x * (1.0f / 32) - mean, d2 = p2.x * (1.0f / 32) - mean, d3 = p3.x * (1.0f / 32) - mean;
      const float m2 = ((p0.y + p1.y) + (p2.y + p3.y)) + 32.0f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
      const float rstd = 1.0f / sqrtf(m2 * (1.0f / 128) + LN_EPS);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = (v[i] - mean) * rstd * sm.lp[0][cq + i] + sm.lp[1][cq + i];
    }
    {
      float ph[16], pl[16];
      tc::split32_packed(v, ph, pl);
      tc::tmem_st16(tm + T_A + cq / 2, ph);
      tc::tmem_st16(tm + T_A + 64 + cq / 2, pl);
    }
    tc::tmem_st_wait();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc::tc_fence_after();
      if (!w_ready) tc::mbar_wait(&sm.bar_w, 0);
      tc::tc_fence_after();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint32_t wh = tc::smem_u32(sm.w[j]);
        const uint64_t dh = tc::make_desc_sw128(wh), dl = tc::make_desc_sw128(wh + 2 * tc::KB_BYTES_128);
        if (tc::elect_one()) {
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint32_t ta = tm0 + T_A + (term == 1 ? 64 : 0);
            const uint64_t db = term == 2 ? dl : dh;
#pragma unroll
            for (int k = 0; k < 128; k += 16)
              tc::mma_bf16_ts(tm0 + (j ? T_V : T_K), ta + k / 2, db + (uint64_t)(((k >> 6) * tc::KB_BYTES_128 + (k & 63) * 2) >> 4), idesc,
                              (term > 0 || k > 0) ? 1u : 0u);
          }
        }
        __syncwarp();
      }
      if (tc::elect_one()) tc::mma_commit(&sm.bar_mma);
      __syncwarp();
    }
    w_ready = true;
    tc::mbar_wait(&sm.bar_mma, n_mma & 1);
    ++n_mma;
    tc::tc_fence_after();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      tc::tmem_ld32(tm + (j ? T_V : T_K) + cq, v);
      tc::tmem_ld_wait();
      if (live) {
        float* dst = kv + row * 256 + 128 * j + cq;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i] + sm.lp[2 + j][cq + 4 * i], v[4 * i + 1] + sm.lp[2 + j][cq + 4 * i + 1],
                                                          v[4 * i + 2] + sm.lp[2 + j][cq + 4 * i + 2], v[4 * i + 3] + sm.lp[2 + j][cq + 4 * i + 3]);
      }
    }
    tc::tc_fence_before();  // the accumulators and the A operand are rewritten by the next tile
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(sm.tmem_base, 512);
}

}  // namespace kvp
}  // namespace tb

static int tc_in_proj_first_block(int block, int layer) {
  switch (block) {
    case TB_BLOCK_MAP_DENSETNT: return tbb::model_map_encoder_transformer_densetnt_layers_0_attn_in_proj_weight + 6 * layer;
    case TB_BLOCK_LATENT_PRIOR_INT: return tbb::model_latent_encoder_agent_interaction_prior_transformer_layers_0_attn_in_proj_weight + 6 * layer;
    case TB_BLOCK_LATENT_POST_INT: return tbb::model_latent_encoder_agent_interaction_post_transformer_layers_0_attn_in_proj_weight + 6 * layer;
    default: return tc_layer_first_block(block, layer);
  }
}

int tb::launch_kv_project_tc(int block, int layer, const float* tgt, long n_row, const float* packed, float* kv, cudaStream_t st) {
  const int first = tc_in_proj_first_block(block, layer);
  if (first < 0) return TB_ERR_UNSUPPORTED;
  static std::atomic<uint64_t> attr_set{0};
  const int smem = (int)sizeof(kvp::Smem) + 1024;
  if (!smem_attr_done(attr_set)) {
    if (!set_max_smem(kvp::k_kv_project_tc, smem)) return TB_ERR_LAUNCH;
    smem_attr_mark(attr_set);
  }
  const long n_tile = (n_row + 127) / 128;
  const int grid = (int)(n_tile < 148 ? n_tile : 148);
  kvp::k_kv_project_tc<<<grid, kvp::THREADS, smem, st>>>(tgt, n_row, packed + block_base(block) + layer * tfl::STRIDE,
                                                        tc_blob(packed) + (size_t)(first + 1) * tc::BLOCK_BYTES, kv);
  count_launch();
  return launch_status();
}

// ------------------------------------------------------------------------------------------------------------
// 3-layer GRU over the frames of a sequence on the tensor pipe (MultiAgentGRULoop 3-D branch + temporal aggregation, see
// tb_gru_sequence in the header).  A CTA owns 128 rows (flattened batch x agent) for the whole sequence; 512 threads, four
// per row; x and h of a layer are two TMEM A operands; gates r and W_hn h in one MMA batch, z and W_in x in the next
// (the same schedule as the decode kernel); the hidden state of the 3 layers lives in a global scratch in row-minor layout.
// ------------------------------------------------------------------------------------------------------------
namespace tb {
namespace gq {

constexpr int THREADS = 512;
constexpr uint32_t T_ACC0 = 0, T_ACC1 = 128, T_A2 = 256, T_A = 384;
struct Smem {
  unsigned char w[2][tc::BLOCK_BYTES];
  float agg[128 * 128];  // temporal aggregate, [col][row]
  float lp[2][6][128];   // b_ih (r, z, n), b_hh (r, z, n) of the current layer
  uint64_t bar_w[2], bar_free[2], bar_mma;
  uint32_t tmem_base;
};

__device__ __forceinline__ float fast_sigmoid(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  return __fdividef(1.0f, 1.0f + e);
}

__global__ void __launch_bounds__(THREADS, 1) k_gru_seq_tc(const float* __restrict__ x, const uint8_t* __restrict__ valid, int T_all,
                                                           int A, long n_rows, int t_stride, int n_t, const float* __restrict__ gw0,
                                                           const unsigned char* __restrict__ tcw, int blk_ih0, int blk_hh0, int mode,
                                                           float4* __restrict__ hscratch, float* __restrict__ out,
                                                           uint8_t* __restrict__ out_valid) {
  extern __shared__ unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tc::uniform(tid >> 5), lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2, r = quad * 32 + lane, cq = 32 * part;
  const long row = (long)blockIdx.x * 128 + r;
  const bool live = row < n_rows;
  const long bb = live ? row / A : 0;
  const int aa = live ? (int)(row % A) : 0;
  if (tid == 0) {
    tc::mbar_init(&sm.bar_w[0], 1);
    tc::mbar_init(&sm.bar_w[1], 1);
    tc::mbar_init(&sm.bar_free[0], 1);
    tc::mbar_init(&sm.bar_free[1], 1);
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
  float4* hs = hscratch + (size_t)blockIdx.x * 3 * 32 * 128 + r;  // layer L, column quad c4: hs[(L * 32 + c4) * 128]

  // weight stage s of the repeating 18-stage schedule (6 per layer): Wih_r, Whh_r, Whh_n | Wih_z, Whh_z, Wih_n
  auto stage_block = [&](uint32_t s) -> int {
    const int L = (s % 18) / 6, j = s % 6;
    const int ih[6] = {1, 0, 0, 1, 0, 1}, gate[6] = {0, 0, 2, 1, 1, 2};
    return (ih[j] ? blk_ih0 : blk_hh0) + 6 * L + gate[j];
  };
  // Weight ring (2 x 64 KB) driven by warp 0: stage j lives in slot j & 1.  A batch = 3 chains (stages s, s+1, s+2); the
  // third reuses the slot of the first, so every chain's completion is tracked on its slot's bar_free and the refill is
  // issued as soon as the slot is free.  All threads wait on bar_mma (committed after the last chain of a batch).
  uint32_t s_next = 0, n_w[2] = {0, 0}, n_free[2] = {0, 0}, n_mma = 0;
  auto load_stage = [&](uint32_t stage) {  // warp 0, converged
    const uint32_t buf = stage & 1;
    if (tc::elect_one()) {
      tc::mbar_expect_tx(&sm.bar_w[buf], tc::BLOCK_BYTES);
      tc::bulk_g2s(sm.w[buf], tcw + (size_t)stage_block(stage) * tc::BLOCK_BYTES, tc::BLOCK_BYTES, &sm.bar_w[buf]);
    }
    __syncwarp();
  };
  auto issue_chain = [&](uint32_t stage, uint32_t dst, uint32_t asel, bool accum, uint64_t* done_bar) {  // warp 0, converged
    const uint32_t buf = stage & 1;
    tc::mbar_wait(&sm.bar_w[buf], n_w[buf] & 1);
    ++n_w[buf];
    tc::tc_fence_after();
    const uint32_t wh = tc::smem_u32(sm.w[buf]);
    const uint64_t dh = tc::make_desc_sw128(wh), dl = tc::make_desc_sw128(wh + 2 * tc::KB_BYTES_128);
    if (tc::elect_one()) {
#pragma unroll
      for (int term = 0; term < 3; ++term) {
        const uint32_t ta = tm0 + asel + (term == 1 ? 64 : 0);
        const uint64_t db = term == 2 ? dl : dh;
#pragma unroll
        for (int k = 0; k < 128; k += 16)
          tc::mma_bf16_ts(tm0 + dst, ta + k / 2, db + (uint64_t)(((k >> 6) * tc::KB_BYTES_128 + (k & 63) * 2) >> 4), idesc,
                          (accum || term > 0 || k > 0) ? 1u : 0u);
      }
      tc::mma_commit(done_bar);
    }
    __syncwarp();
  };
  auto wait_free = [&](uint32_t buf) {
    tc::mbar_wait(&sm.bar_free[buf], n_free[buf] & 1);
    ++n_free[buf];
  };
  // batch of 3 chains: (dst column, A operand column) x 3; the second accumulates onto the first
  auto run_batch = [&](uint32_t d0, uint32_t a0, uint32_t d1, uint32_t a1, uint32_t d2, uint32_t a2) {
    tc::tmem_st_wait();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc::tc_fence_after();
      const uint32_t s0 = s_next, ba = s0 & 1, bb2 = ba ^ 1;
      issue_chain(s0, d0, a0, false, &sm.bar_free[ba]);
      issue_chain(s0 + 1, d1, a1, true, &sm.bar_free[bb2]);
      wait_free(ba);
      load_stage(s0 + 2);
      issue_chain(s0 + 2, d2, a2, false, &sm.bar_mma);
      wait_free(bb2);
      load_stage(s0 + 3);
    }
    tc::mbar_wait(&sm.bar_mma, n_mma & 1);
    ++n_mma;
    tc::tc_fence_after();
    if (warp == 0) load_stage(s_next + 4);  // slot of the batch's first / third chain is free again
    s_next += 3;
  };
  auto write_A = [&](uint32_t col, const float (&v)[32]) {
    float ph[16], pl[16];
    tc::split32_packed(v, ph, pl);
    tc::tmem_st16(tm + col + cq / 2, ph);
    tc::tmem_st16(tm + col + 64 + cq / 2, pl);
  };
  if (warp == 0) {
    load_stage(0);
    load_stage(1);
  }
  for (int i = tid; i < 128 * 128; i += THREADS) sm.agg[i] = mode == 0 ? -1e3f : 0.f;
  bool any = false;
  float xv[32];
#pragma unroll 1
  for (int it = 0; it < n_t; ++it) {
    const int t = it * t_stride;
    const bool v_t = live && valid[(bb * T_all + t) * A + aa] != 0;
    const float* xrow = x + ((bb * T_all + t) * A + aa) * (long)D + cq;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) q = __ldg(reinterpret_cast<const float4*>(xrow) + i);
      xv[4 * i] = q.x, xv[4 * i + 1] = q.y, xv[4 * i + 2] = q.z, xv[4 * i + 3] = q.w;
    }
#pragma unroll 1
    for (int L = 0; L < 3; ++L) {
      const float* gw = gw0 + L * gru::STRIDE;
      float (*lp)[128] = sm.lp[L & 1];
      if (tid < 384) {
        lp[tid >> 7][tid & 127] = __ldg(gw + gru::B_IH + tid);        // b_ih r, z, n
        lp[3 + (tid >> 7)][tid & 127] = __ldg(gw + gru::B_HH + tid);  // b_hh r, z, n
      }
      write_A(T_A, xv);
      {
        float hv[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
          if (it > 0) q = hs[(L * 32 + cq / 4 + i) * 128];
          hv[4 * i] = q.x, hv[4 * i + 1] = q.y, hv[4 * i + 2] = q.z, hv[4 * i + 3] = q.w;
        }
        write_A(T_A2, hv);
      }
      run_batch(T_ACC0, T_A, T_ACC0, T_A2, T_ACC1, T_A2);  // r (ACC0), W_hn h (ACC1)
      float rh[32];
      {
        float rr[32];
        tc::tmem_ld32(tm + T_ACC0 + cq, rr);
        tc::tmem_ld32(tm + T_ACC1 + cq, rh);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) rh[i] = fast_sigmoid(rr[i] + lp[0][cq + i] + lp[3][cq + i]) * (rh[i] + lp[5][cq + i]);
      }
      run_batch(T_ACC0, T_A, T_ACC0, T_A2, T_ACC1, T_A);  // z (ACC0), W_in x (ACC1)
      {
        float zz[32], nn[32];
        tc::tmem_ld32(tm + T_ACC0 + cq, zz);
        tc::tmem_ld32(tm + T_ACC1 + cq, nn);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          float4 hq = make_float4(0.f, 0.f, 0.f, 0.f);
          if (it > 0) hq = hs[(L * 32 + cq / 4 + i4) * 128];
          const float hp_[4] = {hq.x, hq.y, hq.z, hq.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = 4 * i4 + e;
            const float zg = fast_sigmoid(zz[i] + lp[1][cq + i] + lp[4][cq + i]);
            const float ng = 2.0f * fast_sigmoid(2.0f * (nn[i] + lp[2][cq + i] + rh[i])) - 1.0f;
            xv[i] = (1.0f - zg) * ng + zg * hp_[e];  // the next layer sees the unmasked output
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          hs[(L * 32 + cq / 4 + i) * 128] = v_t ? make_float4(xv[4 * i], xv[4 * i + 1], xv[4 * i + 2], xv[4 * i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // output of the frame (zero where invalid) -> temporal aggregate
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) sm.agg[(cq + i) * 128 + r] = fmaxf(sm.agg[(cq + i) * 128 + r], v_t ? xv[i] : -1e3f);
    } else if (v_t) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(xrow) + i);
        sm.agg[(cq + 4 * i) * 128 + r] = xv[4 * i] + q.x;
        sm.agg[(cq + 4 * i + 1) * 128 + r] = xv[4 * i + 1] + q.y;
        sm.agg[(cq + 4 * i + 2) * 128 + r] = xv[4 * i + 2] + q.z;
        sm.agg[(cq + 4 * i + 3) * 128 + r] = xv[4 * i + 3] + q.w;
      }
    }
    any = any || v_t;
  }
  if (live) {
    float* dst = out + row * D + cq;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 q = make_float4(sm.agg[(cq + 4 * i) * 128 + r], sm.agg[(cq + 4 * i + 1) * 128 + r], sm.agg[(cq + 4 * i + 2) * 128 + r],
                             sm.agg[(cq + 4 * i + 3) * 128 + r]);
      if (!any) q = make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(dst)[i] = q;
    }
    if (part == 0) out_valid[row] = any;
  }
  if (warp == 0) {  // the two prefetched stages beyond the last batch must land before the CTA exits
    tc::mbar_wait(&sm.bar_w[0], n_w[0] & 1);
    tc::mbar_wait(&sm.bar_w[1], n_w[1] & 1);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(sm.tmem_base, 512);
}

}  // namespace gq
}  // namespace tb

int tb::launch_gru_seq_tc(int which, int mode, const float* x, const uint8_t* valid, int n_batch, int n_frame, int n_agent, int t_stride,
                          const float* packed, int gru_base_offset, void* workspace, float* out, uint8_t* out_valid, cudaStream_t st) {
  int blk_ih0, blk_hh0;
  switch (which) {
    case TB_GRU_POLICY: blk_ih0 = tbb::model_agent_temporal_rnn_weight_ih_l0, blk_hh0 = tbb::model_agent_temporal_rnn_weight_hh_l0; break;
    case TB_GRU_LATENT_PRIOR:
      blk_ih0 = tbb::model_latent_encoder_agent_temporal_prior_rnn_weight_ih_l0, blk_hh0 = tbb::model_latent_encoder_agent_temporal_prior_rnn_weight_hh_l0;
      break;
    case TB_GRU_LATENT_POST:
      blk_ih0 = tbb::model_latent_encoder_agent_temporal_post_rnn_weight_ih_l0, blk_hh0 = tbb::model_latent_encoder_agent_temporal_post_rnn_weight_hh_l0;
      break;
    case TB_GRU_DEST:
      blk_ih0 = tbb::model_goal_manager_goal_predictor_gru_as_rnn_weight_ih_l0, blk_hh0 = tbb::model_goal_manager_goal_predictor_gru_as_rnn_weight_hh_l0;
      break;
    default: return TB_ERR_BAD_SHAPE;
  }
  static std::atomic<uint64_t> attr_set{0};
  const int smem = (int)sizeof(gq::Smem) + 1024;
  if (!smem_attr_done(attr_set)) {
    if (!set_max_smem(gq::k_gru_seq_tc, smem)) return TB_ERR_LAUNCH;
    smem_attr_mark(attr_set);
  }
  const long n_rows = (long)n_batch * n_agent;
  const int n_t = (n_frame + t_stride - 1) / t_stride;
  const int grid = (int)((n_rows + 127) / 128);
  gq::k_gru_seq_tc<<<grid, gq::THREADS, smem, st>>>(x, valid, n_frame, n_agent, n_rows, t_stride, n_t, packed + gru_base_offset, tc_blob(packed),
                                                   blk_ih0, blk_hh0, mode, reinterpret_cast<float4*>(workspace), out, out_valid);
  count_launch();
  return launch_status();
}
