// Tensor-core (tcgen05) Linear kernels of the training path for launches with many rows (encoders, destination predictor, the
// step-stacked backward: 92 k .. 1 M rows) and 128-wide operands.  Same numerics as the inference kernels: every fp32 contraction
// is three bf16 MMAs with fp32 accumulation in tensor memory (X_hi W_hi + X_lo W_hi + X_hi W_lo, ~2^-17 relative per product).
// Both operands are fp32 in global memory (activations and the CURRENT parameters) and are split into bf16 hi / lo operand tiles
// on the fly (128-byte-swizzled K-major shared-memory tiles, tb_tc.cuh); nothing is pre-packed, the optimizer may change the
// weights between any two launches.
//   k_tr_lin_tc<false> : Y  = epilogue(X W^T)          CTA = 128 output columns, loops over 128-row tiles (W tile split once)
//   k_tr_lin_tc<true>  : dX = dY' W                    (dY' = dY * relu' * row masks * dropout mask, applied while loading)
//   k_tr_lin_tc_dw     : dW += dY'^T X, db += colsum   CTA = a range of 128-row chunks accumulated in tensor memory
#include "tb_host.h"

namespace tb {
namespace {

struct TcDrop {  // see Drop in tb_train.cu (kept in sync: same hash)
  const uint32_t* seed;
  uint32_t site, thresh;
  float scale;
  long offset;
};
__device__ __forceinline__ uint32_t tc_mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float tc_drop_factor(const TcDrop& d, uint32_t key, long idx) {
  return tc_mix32((uint32_t)(idx + d.offset) ^ key) >= d.thresh ? d.scale : 0.f;
}

struct TcArgs {
  const float* bias;       // forward epilogue
  int relu;
  const uint8_t* keep_lin;
  const float* res;
  const uint8_t* keep_out;
  const float* ym;         // backward: ReLU mask source (the forward output)
  const uint8_t* rm1;      // backward: row masks
  const uint8_t* rm2;
  TcDrop drop;
  // geometry of one pass (operands wider than 128 are handled as several passes over 128-wide blocks)
  long a_ld, a_off;          // A operand: row stride and column offset (floats); the ReLU mask source and the dropout index share them
  long w_row, w_col;         // offset into W [N, K]
  long o_ld, o_off;          // output (and residual): row stride and column offset
  int acc_in;                // add the partial result already stored in the output (contraction split over passes)
  int final_pass;            // forward: apply the epilogue (last contraction pass)
};

struct LinSmem {
  unsigned char a_hi[2 * tc::KB_BYTES_128], a_lo[2 * tc::KB_BYTES_128];
  unsigned char b_hi[2 * tc::KB_BYTES_128], b_lo[2 * tc::KB_BYTES_128];
  uint64_t bar_mma;
  uint32_t tmem_base;
};

__device__ __forceinline__ LinSmem& lin_smem(unsigned char* raw) {
  return *reinterpret_cast<LinSmem*>(raw + ((1024u - (tc::smem_u32(raw) & 1023u)) & 1023u));  // SWIZZLE_128B: 1 KB alignment
}

__device__ __forceinline__ void issue_bf16x3(const LinSmem& sm, uint32_t tmem, bool accumulate) {
  const uint32_t idesc = tc::make_idesc_bf16(128, 128);
  const uint32_t ah = tc::smem_u32(sm.a_hi), al = tc::smem_u32(sm.a_lo), bh = tc::smem_u32(sm.b_hi), bl = tc::smem_u32(sm.b_lo);
  tc::mma_tile(tmem, ah, tc::KB_BYTES_128, bh, tc::KB_BYTES_128, 128, idesc, accumulate);
  tc::mma_tile(tmem, al, tc::KB_BYTES_128, bh, tc::KB_BYTES_128, 128, idesc, true);
  tc::mma_tile(tmem, ah, tc::KB_BYTES_128, bl, tc::KB_BYTES_128, 128, idesc, true);
}

// whole warp converged, one elected lane issues (see tb_tc.cuh: "warp-uniform issue")
__device__ __forceinline__ void issue_and_commit(LinSmem& sm, uint32_t tmem, bool accumulate) {
  tc::tc_fence_after();
  if (tc::elect_one()) {
    issue_bf16x3(sm, tmem, accumulate);
    tc::mma_commit(&sm.bar_mma);
  }
  __syncwarp();
}

// 128 x 128 fp32 tile (row stride 128 floats, rows [m0, m0 + 128) of `src`, rows >= M are zero) -> bf16 hi / lo operand tiles.
// 256 threads: thread t owns the 8-float chunk t % 16 of the rows t / 16 + 16 i: every warp instruction reads 2 x 512 contiguous
// bytes and all 16 loads of a thread are in flight together (one memory round trip per tile).
template <bool MASKED>
__device__ __forceinline__ void load_split_tile(unsigned char* hi, unsigned char* lo, const float* __restrict__ src, long m0, long M,
                                                const TcArgs& p, uint32_t dkey) {
  const int t = threadIdx.x, c16 = t & 15, r0 = t >> 4;
  float4 v[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long m = m0 + r0 + 16 * i;
    if (m < M) {
      const float4* q = reinterpret_cast<const float4*>(src + m * p.a_ld + p.a_off + c16 * 8);
      v[i][0] = q[0];
      v[i][1] = q[1];
    } else {
      v[i][0] = v[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + 16 * i;
    const long m = m0 + r;
    float f[8] = {v[i][0].x, v[i][0].y, v[i][0].z, v[i][0].w, v[i][1].x, v[i][1].y, v[i][1].z, v[i][1].w};
    if (MASKED && m < M) {
      bool on = true;
      if (p.rm1 && !p.rm1[m]) on = false;
      if (p.rm2 && !p.rm2[m]) on = false;
      if (p.ym) {
        const float4* q = reinterpret_cast<const float4*>(p.ym + m * p.a_ld + p.a_off + c16 * 8);
        const float4 y0 = q[0], y1 = q[1];
        const float yy[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (!(yy[e] > 0.f)) f[e] = 0.f;
      }
      if (p.drop.seed) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] *= tc_drop_factor(p.drop, dkey, m * p.a_ld + p.a_off + c16 * 8 + e);
      }
      if (!on) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = 0.f;
      }
    }
    uint4 h, l;
    tc::split8(f, h, l);
    const uint32_t off = (c16 >> 3) * tc::KB_BYTES_128 + tc::sw128_off(r, c16 & 7);
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
}

// grid (row-tile CTAs, N / 128); 256 threads
template <bool BWD>
__global__ void __launch_bounds__(256) k_tr_lin_tc(const float* __restrict__ a, const float* __restrict__ w, long ldw,
                                                   float* __restrict__ out, long M, TcArgs p) {
  extern __shared__ unsigned char smem_raw[];
  LinSmem& sm = lin_smem(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j0 = blockIdx.y * 128;
  if (tid == 0) {
    tc::mbar_init(&sm.bar_mma, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&sm.tmem_base, 128);
  // B tile, split once per CTA.  forward: row r = output column j0 + r of W [N, K]; dX: row r = input feature r, K index = n
  if (BWD) {
    if (tid < 128) {
      for (int k0 = 0; k0 < 128; k0 += 32) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = w[(p.w_row + k0 + i) * ldw + p.w_col + tid];
        tc::store_row32_split(sm.b_hi, sm.b_lo, tid, k0, v);
      }
    }
  } else {
    const int c16 = tid & 15, r0 = tid >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = r0 + 16 * i;
      const float4* q = reinterpret_cast<const float4*>(w + (p.w_row + j0 + r) * ldw + p.w_col + c16 * 8);
      const float4 x0 = q[0], x1 = q[1];
      const float f[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      uint4 h, l;
      tc::split8(f, h, l);
      const uint32_t off = (c16 >> 3) * tc::KB_BYTES_128 + tc::sw128_off(r, c16 & 7);
      *reinterpret_cast<uint4*>(sm.b_hi + off) = h;
      *reinterpret_cast<uint4*>(sm.b_lo + off) = l;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t dkey = p.drop.seed ? tc_mix32(p.drop.site * 0x9E3779B9u ^ p.drop.seed[0]) : 0u;
  const long n_tiles = (M + 127) / 128;
  uint32_t phase = 0;
  // accumulator read-back: warp w reads lanes 32 (w % 4) .. + 31 (a hardware restriction), columns 64 (w / 4) .. + 63
  const int erow = 32 * (warp & 3) + lane, ecol = 64 * (warp >> 2);
  for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long m0 = tile * 128;
    load_split_tile<BWD>(sm.a_hi, sm.a_lo, a, m0, M, p, dkey);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) issue_and_commit(sm, tmem, false);
    tc::mbar_wait(&sm.bar_mma, phase);
    phase ^= 1;
    tc::tc_fence_after();
    const long m = m0 + erow;
    const bool row_ok = m < M;
    const float kl = (!BWD && row_ok && p.keep_lin && !p.keep_lin[m]) ? 0.f : 1.f;
    const float ko = (!BWD && row_ok && p.keep_out && !p.keep_out[m]) ? 0.f : 1.f;
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      float v[32];
      tc::tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + ecol + c0, v);
      tc::tmem_ld_wait();
      if (!row_ok) continue;
      const long base = m * p.o_ld + p.o_off + j0 + ecol + c0;
      float4* dst = reinterpret_cast<float4*>(out + base);
      if (p.acc_in) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = dst[i];
          v[4 * i] += t.x, v[4 * i + 1] += t.y, v[4 * i + 2] += t.z, v[4 * i + 3] += t.w;
        }
      }
      if (!BWD && p.final_pass) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float t = v[i];
          if (p.bias) t += p.bias[j0 + ecol + c0 + i];
          if (p.relu) t = fmaxf(t, 0.f);
          t *= kl;
          if (p.drop.seed) t *= tc_drop_factor(p.drop, dkey, base + i);
          if (p.res) t += p.res[base + i];
          v[i] = t * ko;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
    tc::tc_fence_before();
    __syncthreads();  // the next tile overwrites the A operand and the accumulator
    tc::tc_fence_after();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

// dW [128, 128] += dY'^T X over the CTA's row chunks; db += column sums of dY'.  A tile row = output feature n (warps 0-3: thread n
// gathers column n of dY', coalesced across the warp), B tile row = input feature k (warps 4-7); the K dimension of the MMA is
// the row index m within the chunk.
// p.a_ld / p.a_off: row stride / column offset of dY (and its masks); p.o_ld / p.o_off: of X
__global__ void __launch_bounds__(256) k_tr_lin_tc_dw(const float* __restrict__ dy, const float* __restrict__ x, long M,
                                                      float* __restrict__ dw, long lddw, float* __restrict__ db, long chunks_per_cta,
                                                      TcArgs p) {
  extern __shared__ unsigned char smem_raw[];
  LinSmem& sm = lin_smem(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tc::mbar_init(&sm.bar_mma, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&sm.tmem_base, 128);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t dkey = p.drop.seed ? tc_mix32(p.drop.site * 0x9E3779B9u ^ p.drop.seed[0]) : 0u;
  const long n_chunks = (M + 127) / 128;
  const long c_lo = (long)blockIdx.x * chunks_per_cta, c_hi = min(n_chunks, c_lo + chunks_per_cta);
  const bool is_a = tid < 128;  // this thread fills a row of the A tile (dY' column) or of the B tile (X column)
  const int col = tid & 127;
  const float* __restrict__ src = is_a ? dy + p.a_off : x + p.o_off;
  const long ld = is_a ? p.a_ld : p.o_ld;
  unsigned char* t_hi = is_a ? sm.a_hi : sm.b_hi;
  unsigned char* t_lo = is_a ? sm.a_lo : sm.b_lo;
  const bool masked = p.rm1 || p.rm2 || p.ym || p.drop.seed;
  float colsum = 0.f;
  uint32_t phase = 0;
  for (long c = c_lo; c < c_hi; ++c) {
    const long m0 = c * 128;
    const bool full = m0 + 128 <= M;
    for (int i0 = 0; i0 < 128; i0 += 32) {
      float v[32];
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = src[(m0 + i0 + i) * ld + col];
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (m0 + i0 + i < M) ? src[(m0 + i0 + i) * ld + col] : 0.f;
      }
      if (is_a) {
        if (masked) {  // every mask source is loaded as an unconditional batch (full chunks), then applied arithmetically
          if (p.ym) {
            float yv[32];
            const float* ys = p.ym + p.a_off + col;
            if (full) {
#pragma unroll
              for (int i = 0; i < 32; ++i) yv[i] = ys[(m0 + i0 + i) * p.a_ld];
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) yv[i] = (m0 + i0 + i < M) ? ys[(m0 + i0 + i) * p.a_ld] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (!(yv[i] > 0.f)) v[i] = 0.f;
          }
          if (p.rm1 || p.rm2) {
            uint8_t k1[32], k2[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const long m = full ? m0 + i0 + i : min(m0 + i0 + i, M - 1);
              k1[i] = p.rm1 ? p.rm1[m] : 1;
              k2[i] = p.rm2 ? p.rm2[m] : 1;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (!k1[i] || !k2[i]) v[i] = 0.f;
          }
          if (p.drop.seed) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= tc_drop_factor(p.drop, dkey, (m0 + i0 + i) * p.a_ld + p.a_off + col);
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) colsum += v[i];
      }
      tc::store_row32_split(t_hi, t_lo, col, i0, v);
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) issue_and_commit(sm, tmem, c > c_lo);
    tc::mbar_wait(&sm.bar_mma, phase);  // the operand tiles are rewritten by the next chunk
    phase ^= 1;
    tc::tc_fence_after();
  }
  if (c_lo < c_hi) {
    const int erow = 32 * (warp & 3) + lane, ecol = 64 * (warp >> 2);
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      float v[32];
      tc::tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + ecol + c0, v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) atomicAdd(&dw[(long)erow * lddw + ecol + c0 + i], v[i]);
    }
    if (db && is_a) atomicAdd(&db[col], colsum);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

constexpr int LIN_SMEM = (int)sizeof(LinSmem) + 1024;

template <class Kern>
bool ensure_smem(Kern* k, std::atomic<uint64_t>& flag) {
  if (!smem_attr_done(flag)) {
    if (!set_max_smem(k, LIN_SMEM)) return false;
    smem_attr_mark(flag);
  }
  return true;
}

}  // namespace

// tb_train.cu calls these for launches that qualify (many rows, 128-wide operands); they return TB_OK or a launch error
bool train_tc_enabled() {
  const char* e = getenv("TB_TRAIN_NO_TC");
  return !(e && e[0] == '1');
}

int launch_train_linear_tc_fwd(const float* x, long M, int K, const float* w, long ldw, int N, const float* bias, int relu,
                               const uint8_t* keep_lin, const float* res, const uint8_t* keep_out, float* y, const uint32_t* drop_seed,
                               uint32_t drop_site, uint32_t drop_thresh, float drop_scale, long drop_offset, cudaStream_t st) {
  static std::atomic<uint64_t> flag{0};
  if (!ensure_smem(k_tr_lin_tc<false>, flag)) return TB_ERR_LAUNCH;
  const long n_tiles = (M + 127) / 128;
  const int n_col = N / 128, n_kb = K / 128;
  long gx = 148 / n_col;  // one CTA per SM (128 KB of operand tiles)
  if (gx < 1) gx = 1;
  if (gx > n_tiles) gx = n_tiles;
  for (int kb = 0; kb < n_kb; ++kb) {  // contraction blocks: partial sums pass through the output buffer
    TcArgs p{bias, relu, keep_lin, res, keep_out, nullptr, nullptr, nullptr, {drop_seed, drop_site, drop_thresh, drop_scale, drop_offset},
             K, kb * 128L, 0, kb * 128L, N, 0, kb > 0, kb == n_kb - 1};
    k_tr_lin_tc<false><<<dim3((unsigned)gx, n_col), 256, LIN_SMEM, st>>>(x, w, ldw, y, M, p);
    count_launch();
  }
  return launch_status();
}

int launch_train_linear_tc_dx(const float* dy, long M, int K, int N, const float* w, long ldw, const float* ym, const uint8_t* rm1,
                              const uint8_t* rm2, float* dx, const uint32_t* drop_seed, uint32_t drop_site, uint32_t drop_thresh,
                              float drop_scale, long drop_offset, cudaStream_t st) {
  static std::atomic<uint64_t> flag{0};
  if (!ensure_smem(k_tr_lin_tc<true>, flag)) return TB_ERR_LAUNCH;
  const long n_tiles = (M + 127) / 128;
  const long gx = n_tiles < 148 ? n_tiles : 148;
  for (int kb = 0; kb < K / 128; ++kb)      // block of input features (output columns of dX)
    for (int nb = 0; nb < N / 128; ++nb) {  // contraction block
      TcArgs p{nullptr, 0, nullptr, nullptr, nullptr, ym, rm1, rm2, {drop_seed, drop_site, drop_thresh, drop_scale, drop_offset},
               N, nb * 128L, nb * 128L, kb * 128L, K, kb * 128L, nb > 0, 0};
      k_tr_lin_tc<true><<<dim3((unsigned)gx, 1), 256, LIN_SMEM, st>>>(dy, w, ldw, dx, M, p);
      count_launch();
    }
  return launch_status();
}

int launch_train_linear_tc_dw(const float* dy, const float* x, long M, int K, int N, const float* ym, const uint8_t* rm1,
                              const uint8_t* rm2, float* dw, long lddw, float* db, const uint32_t* drop_seed, uint32_t drop_site,
                              uint32_t drop_thresh, float drop_scale, long drop_offset, cudaStream_t st) {
  static std::atomic<uint64_t> flag{0};
  if (!ensure_smem(k_tr_lin_tc_dw, flag)) return TB_ERR_LAUNCH;
  const long n_chunks = (M + 127) / 128;
  const int n_blk = (N / 128) * (K / 128);
  long n_cta = 148 / n_blk;  // the blocks of one Linear are independent launches that may overlap
  if (n_cta < 1) n_cta = 1;
  const long per = (n_chunks + n_cta - 1) / n_cta;
  const long gx = (n_chunks + per - 1) / per;
  for (int nb = 0; nb < N / 128; ++nb)
    for (int kb = 0; kb < K / 128; ++kb) {
      TcArgs p{nullptr, 0, nullptr, nullptr, nullptr, ym, rm1, rm2, {drop_seed, drop_site, drop_thresh, drop_scale, drop_offset},
               N, nb * 128L, 0, 0, K, kb * 128L, 0, 0};
      k_tr_lin_tc_dw<<<(unsigned)gx, 256, LIN_SMEM, st>>>(dy, x, M, dw + nb * 128L * lddw + kb * 128L, lddw,
                                                          (db && kb == 0) ? db + nb * 128L : nullptr, per, p);
      count_launch();
    }
  return launch_status();
}

}  // namespace tb
