// Training primitives (BASELINE.json configs[3]: training_step forward + backward), C ABI `tb_tr_*`.
//
// The reference trains through torch.autograd (src/pl_modules/waymo_motion.py:356-418 + Lightning backward).  Here every
// differentiable operation of that step is one of the primitives below: a forward kernel and a hand-derived backward kernel,
// fp32 throughout (>= the reference's AMP precision).  The host side (trafficbots_b200/train/tape.py, graph.py) only records
// which primitive produced which buffer and replays the backward kernels in reverse order.  All buffers are dense row-major
// fp32 [rows, cols]; masks are uint8.  Every entry point is asynchronous on the passed stream and allocates nothing.
//
// Contractions: one SIMT fp32 GEMM template (64x64x16 tiles, 4x4 register blocking) in its three operand layouts
// (Y = X W^T, dX = dY W, dW += dY^T X with the row dimension split over CTAs).  The shapes of a training step are
// [1k..1M rows] x [<= 384] x [<= 256]: the step is bound by launch latency and activation traffic, not by these GEMMs.
#include "tb_host.h"

namespace tb {
namespace {

constexpr int TR_D = 128;
constexpr int TR_H = 4;
constexpr int TR_DH = 32;

// =====================================================================================================================
// GEMM:  C[i, j] (op)= sum_k A(i, k) * B(k, j),   A(i, k) = a[i * sai + k * sak] (* relu mask ym[i * sai + k * sak] > 0)
// =====================================================================================================================
constexpr int GM = 64, GN = 64, GK = 16;
constexpr long SMALL_M = 8192;  // launches with at most this many rows use 32-row tiles
constexpr long TC_MIN_M = 8192; // launches with at least this many rows and 128-wide operands run on the tensor cores (tb_train_tc.cu)
enum { EPI_BIAS = 0, EPI_STORE = 1, EPI_ATOMIC = 2 };

// Per-tile options.  A operand: `ym` = ReLU mask source (same indexing as A: element kept where ym > 0), `rm1` / `rm2` = row masks
// over the M ("sample") dimension of A (uint8, element dropped where 0).  EPI_BIAS epilogue: v = acc + bias; ReLU; v *= keep_lin[i];
// v += res[i, j]; v *= keep_out[i].  `rowsum` (EPI_ATOMIC tiles with tile_j == 0): rowsum[i] += sum_k A(i, k) (bias gradient).
// Dropout (nn.Dropout in training mode: mlp.py:53-63, transformer.py:116-134, attention.py:131-132, nn.GRU inter-layer): the keep
// mask is a counter-based hash of (per-step seed on the device, site id of the call, element index) -- reproducible in the
// backward kernel without storing it.  seed == NULL switches it off.
struct Drop {
  const uint32_t* seed;
  uint32_t site;
  uint32_t thresh;  // element kept iff hash >= thresh (= p * 2^32)
  float scale;      // 1 / (1 - p)
  long offset;      // added to the element index: a launch that covers rows [t M, (t + 1) M) of a step-stacked buffer passes t M N,
                    // so that ONE backward launch over the whole stacked buffer regenerates the masks of all steps
};
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t drop_key(const Drop& d) { return d.seed ? mix32(d.site * 0x9E3779B9u ^ d.seed[0]) : 0u; }
__device__ __forceinline__ float drop_factor(const Drop& d, uint32_t key, long idx) {
  return mix32((uint32_t)(idx + d.offset) ^ key) >= d.thresh ? d.scale : 0.f;
}
inline Drop make_drop(const uint32_t* seed, uint32_t site, float p, long offset) {
  Drop d{nullptr, 0u, 0u, 1.f, 0};
  if (seed && p > 0.f) {
    d.seed = seed;
    d.site = site;
    d.offset = offset;
    d.thresh = (uint32_t)((double)p * 4294967296.0);
    d.scale = 1.f / (1.f - p);
  }
  return d;
}

struct GemmOpt {
  const float* ym;
  const uint8_t* rm1;
  const uint8_t* rm2;
  const float* bias;
  int relu;
  const uint8_t* keep_lin;
  const float* res;
  const uint8_t* keep_out;
  float* rowsum;
  Drop drop;  // forward: applied after ReLU / keep_lin, before the residual; backward: applied to the A operand (dY)
};

// The next k-slab is fetched into registers while the current one is multiplied out of shared memory (global-load latency
// hidden behind the FMAs); 16-byte loads where the layout allows.
// TM = rows of the output tile (64: 4 x 4 outputs per thread; 32: 2 x 4 -- twice the CTAs for the 1 k-row launches of a decode
// step, which are bound by the dependent k-slab iterations of a CTA, not by FLOPs)
template <int TM, bool A_KCONTIG, bool B_JCONTIG, int EPI>
__device__ __forceinline__ void gemm_tile(float (&As)[GK][GM + 4], float (&Bs)[GK][GN + 4], const float* __restrict__ a, long sai,
                                          long sak, const float* __restrict__ b, long sbk, long sbj, float* __restrict__ c, long ldc,
                                          long ni, int nj, long nk, long k_chunk, long tile_i, int tile_j, long tile_z,
                                          const GemmOpt& op) {
  const int tid = threadIdx.x;
  constexpr int RX = TM / 16;  // output rows per thread
  const long i0 = tile_i * TM;
  const int j0 = tile_j * GN;
  const long k_lo = tile_z * k_chunk;
  const long k_hi = min(nk, k_lo + k_chunk);
  const int ti = tid / 16, tj = tid % 16;  // 16 x 16 threads, 4 x 4 outputs each
  float acc[RX][4] = {};
  float rs = 0.f;
  const bool do_rowsum = (EPI == EPI_ATOMIC) && op.rowsum != nullptr && tile_j == 0;
  const float* __restrict__ ym = op.ym;
  // thread -> element mapping of the two tile loads (4 consecutive elements along the contiguous direction); with TM = 32 the A
  // tile has 512 elements: threads 0..127 load it
  const bool a_thread = tid < TM * 4;
  const int a_r = A_KCONTIG ? tid / 4 : (tid % (TM / 4)) * 4;   // row (i) offset
  const int a_k = A_KCONTIG ? (tid % 4) * 4 : tid / (TM / 4);   // k offset
  const int b_j = B_JCONTIG ? (tid % 16) * 4 : tid / 4;
  const int b_k = B_JCONTIG ? tid / 16 : (tid % 4) * 4;
  const bool a_vec = A_KCONTIG ? (sak == 1 && (sai & 3) == 0) : (sai == 1 && (sak & 3) == 0);
  const bool b_vec = B_JCONTIG ? (sbj == 1 && (sbk & 3) == 0) : (sbk == 1 && (sbj & 3) == 0);
  const bool a_al = a_vec && ((reinterpret_cast<uintptr_t>(a) & 15) == 0) && (!ym || (reinterpret_cast<uintptr_t>(ym) & 15) == 0);
  const bool b_al = b_vec && ((reinterpret_cast<uintptr_t>(b) & 15) == 0);
  float ra[4], rb[4];
  const bool a_drop = EPI != EPI_BIAS && op.drop.seed != nullptr;
  const uint32_t dkey = drop_key(op.drop);

  auto fetch = [&](long k0) {
    // ---- A ----
    if (!a_thread) {
    } else if (A_KCONTIG) {
      const long i = i0 + a_r, k = k0 + a_k;
      bool row_ok = i < ni;
      if (row_ok && op.rm1 && !op.rm1[i]) row_ok = false;
      if (row_ok && op.rm2 && !op.rm2[i]) row_ok = false;
      if (row_ok && a_al && k + 3 < k_hi) {
        const float4 v = *reinterpret_cast<const float4*>(a + i * sai + k);
        ra[0] = v.x, ra[1] = v.y, ra[2] = v.z, ra[3] = v.w;
        if (ym) {
          const float4 m = *reinterpret_cast<const float4*>(ym + i * sai + k);
          if (!(m.x > 0.f)) ra[0] = 0.f;
          if (!(m.y > 0.f)) ra[1] = 0.f;
          if (!(m.z > 0.f)) ra[2] = 0.f;
          if (!(m.w > 0.f)) ra[3] = 0.f;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float v = 0.f;
          if (row_ok && k + u < k_hi) {
            v = a[i * sai + (k + u) * sak];
            if (ym && !(ym[i * sai + (k + u) * sak] > 0.f)) v = 0.f;
          }
          ra[u] = v;
        }
      }
    } else {
      const long i = i0 + a_r, k = k0 + a_k;
      bool k_ok = k < k_hi;
      if (k_ok && op.rm1 && !op.rm1[k]) k_ok = false;
      if (k_ok && op.rm2 && !op.rm2[k]) k_ok = false;
      if (k_ok && a_al && i + 3 < ni) {
        const float4 v = *reinterpret_cast<const float4*>(a + i + k * sak);
        ra[0] = v.x, ra[1] = v.y, ra[2] = v.z, ra[3] = v.w;
        if (ym) {
          const float4 m = *reinterpret_cast<const float4*>(ym + i + k * sak);
          if (!(m.x > 0.f)) ra[0] = 0.f;
          if (!(m.y > 0.f)) ra[1] = 0.f;
          if (!(m.z > 0.f)) ra[2] = 0.f;
          if (!(m.w > 0.f)) ra[3] = 0.f;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float v = 0.f;
          if (k_ok && i + u < ni) {
            v = a[(i + u) * sai + k * sak];
            if (ym && !(ym[(i + u) * sai + k * sak] > 0.f)) v = 0.f;
          }
          ra[u] = v;
        }
      }
    }
    if (a_drop && a_thread) {  // the 4 elements are consecutive in dY: offsets off .. off + 3
      const long off = A_KCONTIG ? (i0 + a_r) * sai + (k0 + a_k) : (i0 + a_r) + (k0 + a_k) * sak;
#pragma unroll
      for (int u = 0; u < 4; ++u) ra[u] *= drop_factor(op.drop, dkey, off + (A_KCONTIG ? u * sak : u * sai));
    }
    // ---- B ----
    if (B_JCONTIG) {
      const long k = k0 + b_k;
      const int jj = j0 + b_j;
      if (k < k_hi && b_al && jj + 3 < nj) {
        const float4 v = *reinterpret_cast<const float4*>(b + k * sbk + jj);
        rb[0] = v.x, rb[1] = v.y, rb[2] = v.z, rb[3] = v.w;
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) rb[u] = (k < k_hi && jj + u < nj) ? b[k * sbk + (long)(jj + u) * sbj] : 0.f;
      }
    } else {
      const long k = k0 + b_k;
      const int jj = j0 + b_j;
      if (jj < nj && b_al && k + 3 < k_hi) {
        const float4 v = *reinterpret_cast<const float4*>(b + k + (long)jj * sbj);
        rb[0] = v.x, rb[1] = v.y, rb[2] = v.z, rb[3] = v.w;
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) rb[u] = (jj < nj && k + u < k_hi) ? b[(k + u) * sbk + (long)jj * sbj] : 0.f;
      }
    }
  };
  auto stash = [&]() {
    if (!a_thread) {
    } else if (A_KCONTIG) {
#pragma unroll
      for (int u = 0; u < 4; ++u) As[a_k + u][a_r] = ra[u];
    } else {
      *reinterpret_cast<float4*>(&As[a_k][a_r]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
    }
    if (B_JCONTIG) {
      *reinterpret_cast<float4*>(&Bs[b_k][b_j]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) Bs[b_k + u][b_j] = rb[u];
    }
  };

  if (k_lo < k_hi) fetch(k_lo);
  for (long k0 = k_lo; k0 < k_hi; k0 += GK) {
    stash();
    __syncthreads();
    if (k0 + GK < k_hi) fetch(k0 + GK);
    if (do_rowsum && tid < TM) {
#pragma unroll
      for (int kk = 0; kk < GK; ++kk) rs += As[kk][tid];
    }
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float ar[RX];
      if (RX == 4) {
        const float4 av = *reinterpret_cast<const float4*>(&As[kk][ti * 4]);
        ar[0] = av.x, ar[1] = av.y, ar[RX - 2] = av.z, ar[RX - 1] = av.w;
      } else {
        const float2 av = *reinterpret_cast<const float2*>(&As[kk][ti * 2]);
        ar[0] = av.x, ar[1] = av.y;
      }
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tj * 4]);
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int x = 0; x < RX; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(ar[x], br[y], acc[x][y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < RX; ++x) {
    const long i = i0 + ti * RX + x;
    if (i >= ni) continue;
    float kl = 1.f, ko = 1.f;
    if (EPI == EPI_BIAS) {
      if (op.keep_lin && !op.keep_lin[i]) kl = 0.f;
      if (op.keep_out && !op.keep_out[i]) ko = 0.f;
    }
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int j = j0 + tj * 4 + y;
      if (j >= nj) continue;
      float v = acc[x][y];
      if (EPI == EPI_BIAS) {
        if (op.bias) v += op.bias[j];
        if (op.relu) v = fmaxf(v, 0.f);
        v *= kl;
        if (op.drop.seed) v *= drop_factor(op.drop, dkey, i * ldc + j);
        if (op.res) v += op.res[i * ldc + j];
        v *= ko;
        c[i * ldc + j] = v;
      } else if (EPI == EPI_STORE) {
        c[i * ldc + j] = v;
      } else {
        atomicAdd(&c[i * ldc + j], v);
      }
    }
  }
  if (do_rowsum && tid < TM && i0 + tid < ni) atomicAdd(&op.rowsum[i0 + tid], rs);
}

template <int TM, bool A_KCONTIG, bool B_JCONTIG, int EPI>
__global__ void __launch_bounds__(256) k_tr_gemm(const float* __restrict__ a, long sai, long sak, const float* __restrict__ b, long sbk,
                                                 long sbj, float* __restrict__ c, long ldc, long ni, int nj, long nk, long k_chunk,
                                                 GemmOpt op) {
  __shared__ __align__(16) float As[GK][GM + 4];
  __shared__ __align__(16) float Bs[GK][GN + 4];
  gemm_tile<TM, A_KCONTIG, B_JCONTIG, EPI>(As, Bs, a, sai, sak, b, sbk, sbj, c, ldc, ni, nj, nk, k_chunk, blockIdx.x, blockIdx.y,
                                           blockIdx.z, op);
}

// backward of a Linear in ONE launch: CTAs [0, n_dx) compute tiles of dX = dY' W, the others tiles of dW += dY'^T X (rows split
// over `nz` chunks, partial sums added atomically) and, in the first column tile, db += colsum(dY');
// dY' = dY * relu'(Y) * rm1[row] * rm2[row]
template <int TM>
__global__ void __launch_bounds__(256) k_tr_linear_bwd(const float* __restrict__ dy, const float* __restrict__ x,
                                                       const float* __restrict__ w, long ldw, long M, int K, int N,
                                                       float* __restrict__ dx, float* __restrict__ dw, long lddw, int n_dx,
                                                       int dx_tiles_j, int dw_tiles_i, int dw_tiles_j, long chunk, GemmOpt op) {
  __shared__ __align__(16) float As[GK][GM + 4];
  __shared__ __align__(16) float Bs[GK][GN + 4];
  int t = blockIdx.x;
  if (t < n_dx) {
    gemm_tile<TM, true, true, EPI_STORE>(As, Bs, dy, N, 1, w, ldw, 1, dx, K, M, K, N, N, t / dx_tiles_j, t % dx_tiles_j, 0, op);
  } else {
    t -= n_dx;
    const int per_z = dw_tiles_i * dw_tiles_j;
    const int z = t / per_z, r = t % per_z;
    gemm_tile<TM, false, true, EPI_ATOMIC>(As, Bs, dy, 1, N, x, K, 1, dw, lddw, N, K, M, chunk, r / dw_tiles_j, r % dw_tiles_j, z, op);
  }
}

// =====================================================================================================================
// LayerNorm over 128 columns (+ReLU): one warp per row
// =====================================================================================================================
__global__ void __launch_bounds__(256) k_tr_ln_fwd(const float* __restrict__ x, const float* __restrict__ w,
                                                   const float* __restrict__ b, int relu, long M, float* __restrict__ y,
                                                   float* __restrict__ stats, Drop drop) {
  const long row = (long)blockIdx.x * 8 + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= M) return;
  const float4 v = reinterpret_cast<const float4*>(x + row * TR_D)[lane];
  float s = v.x + v.y + v.z + v.w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.f / TR_D);
  const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
  float q = dx * dx + dy * dy + dz * dz + dw * dw;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.f / TR_D) + 1e-5f);
  const float4 wv = reinterpret_cast<const float4*>(w)[lane], bv = reinterpret_cast<const float4*>(b)[lane];
  float4 o4 = make_float4(dx * rstd * wv.x + bv.x, dy * rstd * wv.y + bv.y, dz * rstd * wv.z + bv.z, dw * rstd * wv.w + bv.w);
  if (relu) o4 = make_float4(fmaxf(o4.x, 0.f), fmaxf(o4.y, 0.f), fmaxf(o4.z, 0.f), fmaxf(o4.w, 0.f));
  if (drop.seed) {
    const uint32_t key = drop_key(drop);
    const long e0 = row * TR_D + lane * 4;
    o4.x *= drop_factor(drop, key, e0), o4.y *= drop_factor(drop, key, e0 + 1), o4.z *= drop_factor(drop, key, e0 + 2),
        o4.w *= drop_factor(drop, key, e0 + 3);
  }
  reinterpret_cast<float4*>(y + row * TR_D)[lane] = o4;
  if (lane == 0) {
    stats[row * 2] = mean;
    stats[row * 2 + 1] = rstd;
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w;   dw += sum dy * xhat,  db += sum dy
__global__ void __launch_bounds__(256) k_tr_ln_bwd(const float* __restrict__ dy, const float* __restrict__ x,
                                                   const float* __restrict__ w, const float* __restrict__ stats,
                                                   const float* __restrict__ y, int relu, long M, long rows_per_block,
                                                   float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, Drop drop) {
  __shared__ float red[2][8][TR_D];
  const uint32_t dkey = drop_key(drop);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const long r_lo = (long)blockIdx.x * rows_per_block, r_hi = min(M, r_lo + rows_per_block);
  const float4 wv = reinterpret_cast<const float4*>(w)[lane];
  float aw[4] = {}, ab[4] = {};
  for (long row = r_lo + warp; row < r_hi; row += 8) {
    float4 g = reinterpret_cast<const float4*>(dy + row * TR_D)[lane];
    if (drop.seed) {
      const long e0 = row * TR_D + lane * 4;
      g.x *= drop_factor(drop, dkey, e0), g.y *= drop_factor(drop, dkey, e0 + 1), g.z *= drop_factor(drop, dkey, e0 + 2),
          g.w *= drop_factor(drop, dkey, e0 + 3);
    }
    if (relu) {
      const float4 yv = reinterpret_cast<const float4*>(y + row * TR_D)[lane];
      if (!(yv.x > 0.f)) g.x = 0.f;
      if (!(yv.y > 0.f)) g.y = 0.f;
      if (!(yv.z > 0.f)) g.z = 0.f;
      if (!(yv.w > 0.f)) g.w = 0.f;
    }
    const float4 xv = reinterpret_cast<const float4*>(x + row * TR_D)[lane];
    const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
    const float xh[4] = {(xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd};
    const float gy[4] = {g.x, g.y, g.z, g.w};
    const float wr[4] = {wv.x, wv.y, wv.z, wv.w};
    float gw[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      gw[u] = gy[u] * wr[u];
      s1 += gw[u];
      s2 += gw[u] * xh[u];
      aw[u] += gy[u] * xh[u];
      ab[u] += gy[u];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 *= (1.f / TR_D);
    s2 *= (1.f / TR_D);
    reinterpret_cast<float4*>(dx + row * TR_D)[lane] =
        make_float4(rstd * (gw[0] - s1 - xh[0] * s2), rstd * (gw[1] - s1 - xh[1] * s2), rstd * (gw[2] - s1 - xh[2] * s2),
                    rstd * (gw[3] - s1 - xh[3] * s2));
  }
  if (dw == nullptr) return;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    red[0][warp][lane * 4 + u] = aw[u];
    red[1][warp][lane * 4 + u] = ab[u];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * TR_D; c += 256) {
    const int which = c / TR_D, col = c % TR_D;
    float s = 0.f;
#pragma unroll
    for (int wi = 0; wi < 8; ++wi) s += red[which][wi][col];
    atomicAdd(which == 0 ? &dw[col] : &db[col], s);
  }
}

// =====================================================================================================================
// multi-head attention core: 4 heads x 32, masks, dead rows (models/modules/attention.py:89-141)
// =====================================================================================================================
constexpr int AT_QT = 8;  // queries per CTA (forward)

// grid (ceil(S / AT_QT), H, B), 128 threads; dynamic smem: AT_QT * T logits
__global__ void __launch_bounds__(128, 4) k_tr_attn_fwd(const float* __restrict__ q, const float* __restrict__ kv,
                                                     const uint8_t* __restrict__ key_valid, int eye, int S, int T,
                                                     float* __restrict__ o, float* __restrict__ p, uint8_t* __restrict__ alive,
                                                     Drop drop, long b_off) {
  extern __shared__ __align__(16) float sm[];
  float* lg = sm;                     // [AT_QT][T]
  __shared__ float qs[AT_QT][TR_DH];
  __shared__ float rmax[AT_QT], rinv[AT_QT];
  const int b = blockIdx.z, h = blockIdx.y, s0 = blockIdx.x * AT_QT;
  const int tid = threadIdx.x;
  const int nq = min(AT_QT, S - s0);
  for (int e = tid; e < AT_QT * TR_DH; e += 128) {
    const int qi = e / TR_DH, d = e % TR_DH;
    qs[qi][d] = qi < nq ? q[((long)b * S + s0 + qi) * TR_D + h * TR_DH + d] : 0.f;
  }
  __syncthreads();
  const float scale = 0.17677669529663688110f;  // 1 / sqrt(32), applied after the -inf fill (attention.py:128-130)
  const float NEG = -INFINITY;
  const uint32_t dkey = drop_key(drop);
  for (int j = tid; j < T; j += 128) {
    const float4* kp = reinterpret_cast<const float4*>(kv + ((long)b * T + j) * 2 * TR_D + h * TR_DH);
    float kr[TR_DH];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float4 t4 = kp[u];
      kr[u * 4] = t4.x, kr[u * 4 + 1] = t4.y, kr[u * 4 + 2] = t4.z, kr[u * 4 + 3] = t4.w;
    }
    const bool kvld = key_valid[(long)b * T + j] != 0;
#pragma unroll 2
    for (int qi = 0; qi < AT_QT; ++qi) {
      float dot = 0.f;
#pragma unroll
      for (int d = 0; d < TR_DH; ++d) dot = fmaf(qs[qi][d], kr[d], dot);
      const bool ok = kvld && !(eye && j == s0 + qi);
      lg[qi * T + j] = ok ? dot * scale : NEG;
    }
  }
  __syncthreads();
  const int warp = tid / 32, lane = tid % 32;
  for (int qi = warp; qi < AT_QT; qi += 4) {
    float m = NEG;
    for (int j = lane; j < T; j += 32) m = fmaxf(m, lg[qi * T + j]);
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, of));
    float s = 0.f;
    if (m > NEG) {
      for (int j = lane; j < T; j += 32) {
        const float e = expf(lg[qi * T + j] - m);
        lg[qi * T + j] = e;
        s += e;
      }
    } else {
      for (int j = lane; j < T; j += 32) lg[qi * T + j] = 0.f;
    }
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) s += __shfl_xor_sync(0xffffffffu, s, of);
    if (lane == 0) {
      rmax[qi] = m;
      rinv[qi] = m > NEG ? 1.f / s : 0.f;
      if (h == 0 && qi < nq) alive[(long)b * S + s0 + qi] = m > NEG ? 1 : 0;
    }
  }
  __syncthreads();
  for (int e = tid; e < nq * T; e += 128) {
    const int qi = e / T, j = e % T;
    const float pv = lg[qi * T + j] * rinv[qi];
    const long pidx = (((long)b * TR_H + h) * S + s0 + qi) * T + j;
    p[pidx] = pv;  // the un-dropped probabilities are what the backward needs; O uses the dropped ones (attention.py:131-136)
    lg[qi * T + j] = drop.seed ? pv * drop_factor(drop, dkey, pidx + b_off) : pv;
  }
  __syncthreads();
  // O = P V : thread (d, key slice g) accumulates all AT_QT queries over the keys {4 (g + 4 i) .. + 3}; the four slices are then
  // summed through shared memory (the logits buffer is free after the barrier)
  const int d = tid % 32, g = tid / 32;
  float acc[AT_QT] = {};
  const float* vbase = kv + (long)b * T * 2 * TR_D + TR_D + h * TR_DH + d;
  if ((T & 3) == 0) {
#pragma unroll 2
    for (int j4 = g * 4; j4 < T; j4 += 16) {
      const float v0 = vbase[(long)(j4 + 0) * 2 * TR_D], v1 = vbase[(long)(j4 + 1) * 2 * TR_D];
      const float v2 = vbase[(long)(j4 + 2) * 2 * TR_D], v3 = vbase[(long)(j4 + 3) * 2 * TR_D];
#pragma unroll
      for (int qi = 0; qi < AT_QT; ++qi) {
        const float4 pq = *reinterpret_cast<const float4*>(&lg[qi * T + j4]);
        acc[qi] = fmaf(pq.x, v0, fmaf(pq.y, v1, fmaf(pq.z, v2, fmaf(pq.w, v3, acc[qi]))));
      }
    }
  } else {
    for (int j = g; j < T; j += 4) {
      const float vv = vbase[(long)j * 2 * TR_D];
#pragma unroll
      for (int qi = 0; qi < AT_QT; ++qi) acc[qi] = fmaf(lg[qi * T + j], vv, acc[qi]);
    }
  }
  __syncthreads();
  float* red = lg;  // [4][AT_QT][32]
#pragma unroll
  for (int qi = 0; qi < AT_QT; ++qi) red[(g * AT_QT + qi) * 32 + d] = acc[qi];
  __syncthreads();
  for (int e = tid; e < nq * 32; e += 128) {
    const int qi = e / 32, dd = e % 32;
    o[((long)b * S + s0 + qi) * TR_D + h * TR_DH + dd] =
        red[(0 * AT_QT + qi) * 32 + dd] + red[(1 * AT_QT + qi) * 32 + dd] + red[(2 * AT_QT + qi) * 32 + dd] + red[(3 * AT_QT + qi) * 32 + dd];
  }
}

// backward: CTA = (key chunk of 64, head, batch); loops over query tiles of 16.  dK / dV of its keys are complete (plain
// stores), dQ partials are added atomically (dq zero-initialised by the caller).  Shared-memory rows are padded to 16-byte
// multiples and every inner product is register-blocked over 16-byte shared loads (the first version issued one 4-byte shared
// load per FMA and was bound by the LSU pipe).
constexpr int AB_KC = 64, AB_QT = 16;
constexpr int AB_LD = TR_DH + 4;   // 36: row stride of the [*, 32] tiles
constexpr int AB_LDS = AB_KC + 4;  // 68: row stride of the [16, 64] tiles
__global__ void __launch_bounds__(256) k_tr_attn_bwd(const float* __restrict__ dout, const float* __restrict__ q,
                                                     const float* __restrict__ kv, const float* __restrict__ p,
                                                     const float* __restrict__ o, int S, int T, float* __restrict__ dq,
                                                     float* __restrict__ dkv, Drop drop, int b_base, int kv_batch) {
  __shared__ __align__(16) float ks[AB_KC][AB_LD], vs[AB_KC][AB_LD];
  __shared__ __align__(16) float qs[AB_QT][AB_LD], gs[AB_QT][AB_LD];
  __shared__ __align__(16) float ps[AB_QT][AB_LDS], ds[AB_QT][AB_LDS];
  __shared__ __align__(16) float pf[AB_QT][AB_LDS];  // P * dropout factor (= P without dropout): what multiplied V in the forward
  const uint32_t dkey = drop_key(drop);
  __shared__ float delta[AB_QT];
  // kv_batch > 0: the K|V rows of batch element b are those of b % kv_batch (step-stacked queries [t][scene] against the per-scene
  // map keys); dK / dV of a key then receive contributions from several CTAs and are added atomically (dkv zero-initialised)
  const int b = blockIdx.z + b_base, h = blockIdx.y, j0 = blockIdx.x * AB_KC;
  const int bk = kv_batch > 0 ? b % kv_batch : b;
  const int tid = threadIdx.x;
  const int nk = min(AB_KC, T - j0);
  for (int e = tid; e < AB_KC * TR_DH; e += 256) {
    const int j = e / TR_DH, d = e % TR_DH;
    const bool ok = j < nk;
    const long base = ((long)bk * T + j0 + j) * 2 * TR_D + h * TR_DH + d;
    ks[j][d] = ok ? kv[base] : 0.f;
    vs[j][d] = ok ? kv[base + TR_D] : 0.f;
  }
  const float scale = 0.17677669529663688110f;
  // dK / dV accumulators: thread owns feature od and the 8 consecutive keys oj * 8 .. + 7
  float adk[8] = {}, adv[8] = {};
  const int od = tid % 32, oj = tid / 32;
  // dS: thread owns query sq and the keys sj + 16 u;  dQ: thread owns query sq and the features 2 sj, 2 sj + 1
  const int sq = tid / 16, sj = tid % 16;
  for (int s0 = 0; s0 < S; s0 += AB_QT) {
    const int nq = min(AB_QT, S - s0);
    __syncthreads();
    for (int e = tid; e < AB_QT * TR_DH; e += 256) {
      const int qi = e / TR_DH, d = e % TR_DH;
      const long idx = ((long)b * S + s0 + qi) * TR_D + h * TR_DH + d;
      qs[qi][d] = qi < nq ? q[idx] : 0.f;
      gs[qi][d] = qi < nq ? dout[idx] : 0.f;
    }
    for (int e = tid; e < AB_QT * AB_KC; e += 256) {
      const int qi = e / AB_KC, j = e % AB_KC;
      const long pidx = (((long)b * TR_H + h) * S + s0 + qi) * T + j0 + j;
      const float pv = (qi < nq && j < nk) ? p[pidx] : 0.f;
      ps[qi][j] = pv;
      pf[qi][j] = drop.seed ? pv * drop_factor(drop, dkey, pidx) : pv;
    }
    if (tid < AB_QT * 2) {  // delta = sum_d do * o, two half rows per query
      const int qi = tid / 2, half = tid % 2;
      float s = 0.f;
      if (qi < nq) {
        const long idx = ((long)b * S + s0 + qi) * TR_D + h * TR_DH + half * 16;
        for (int d = 0; d < 16; ++d) s = fmaf(dout[idx + d], o[idx + d], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (half == 0) delta[qi] = s;
    }
    __syncthreads();
    {  // dS = P * (dO V^T - delta) * scale
      float dp[4] = {};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 g4 = *reinterpret_cast<const float4*>(&gs[sq][c * 4]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 v4 = *reinterpret_cast<const float4*>(&vs[sj + 16 * u][c * 4]);
          dp[u] = fmaf(g4.x, v4.x, fmaf(g4.y, v4.y, fmaf(g4.z, v4.z, fmaf(g4.w, v4.w, dp[u]))));
        }
      }
      const float dl = delta[sq];
#pragma unroll
      for (int u = 0; u < 4; ++u) ds[sq][sj + 16 * u] = (pf[sq][sj + 16 * u] * dp[u] - ps[sq][sj + 16 * u] * dl) * scale;
    }
    __syncthreads();
    if (sq < nq) {  // dQ partial of (sq, 2 sj .. 2 sj + 1)
      float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
      for (int jj = 0; jj < AB_KC; jj += 4) {
        const float4 d4 = *reinterpret_cast<const float4*>(&ds[sq][jj]);
        const float2 k0 = *reinterpret_cast<const float2*>(&ks[jj + 0][2 * sj]);
        const float2 k1 = *reinterpret_cast<const float2*>(&ks[jj + 1][2 * sj]);
        const float2 k2 = *reinterpret_cast<const float2*>(&ks[jj + 2][2 * sj]);
        const float2 k3 = *reinterpret_cast<const float2*>(&ks[jj + 3][2 * sj]);
        a0 = fmaf(d4.x, k0.x, fmaf(d4.y, k1.x, fmaf(d4.z, k2.x, fmaf(d4.w, k3.x, a0))));
        a1 = fmaf(d4.x, k0.y, fmaf(d4.y, k1.y, fmaf(d4.z, k2.y, fmaf(d4.w, k3.y, a1))));
      }
      float* dst = dq + ((long)b * S + s0 + sq) * TR_D + h * TR_DH + 2 * sj;
      atomicAdd(dst, a0);
      atomicAdd(dst + 1, a1);
    }
    // dK, dV accumulation
#pragma unroll 4
    for (int qi = 0; qi < AB_QT; ++qi) {
      const float qv = qs[qi][od], gv = gs[qi][od];
      const float4 da = *reinterpret_cast<const float4*>(&ds[qi][oj * 8]), db = *reinterpret_cast<const float4*>(&ds[qi][oj * 8 + 4]);
      const float4 pa = *reinterpret_cast<const float4*>(&pf[qi][oj * 8]), pb = *reinterpret_cast<const float4*>(&pf[qi][oj * 8 + 4]);
      adk[0] = fmaf(da.x, qv, adk[0]), adk[1] = fmaf(da.y, qv, adk[1]), adk[2] = fmaf(da.z, qv, adk[2]), adk[3] = fmaf(da.w, qv, adk[3]);
      adk[4] = fmaf(db.x, qv, adk[4]), adk[5] = fmaf(db.y, qv, adk[5]), adk[6] = fmaf(db.z, qv, adk[6]), adk[7] = fmaf(db.w, qv, adk[7]);
      adv[0] = fmaf(pa.x, gv, adv[0]), adv[1] = fmaf(pa.y, gv, adv[1]), adv[2] = fmaf(pa.z, gv, adv[2]), adv[3] = fmaf(pa.w, gv, adv[3]);
      adv[4] = fmaf(pb.x, gv, adv[4]), adv[5] = fmaf(pb.y, gv, adv[5]), adv[6] = fmaf(pb.z, gv, adv[6]), adv[7] = fmaf(pb.w, gv, adv[7]);
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int j = oj * 8 + u;
    if (j < nk) {
      const long base = ((long)bk * T + j0 + j) * 2 * TR_D + h * TR_DH + od;
      if (kv_batch > 0) {
        atomicAdd(&dkv[base], adk[u]);
        atomicAdd(&dkv[base + TR_D], adv[u]);
      } else {
        dkv[base] = adk[u];
        dkv[base + TR_D] = adv[u];
      }
    }
  }
}

// ---- at most 32 keys (the 20 nodes of a polyline, map_encoder.py:72-88): one warp per (batch element, head), lane = key ----
// K rows live in registers, V in shared memory; per query the row of Q is broadcast lane by lane (shuffles), the softmax is two
// warp reductions, O = P V is accumulated with lane = feature.  Same arithmetic as the general kernel.
__global__ void __launch_bounds__(128) k_tr_attn_small_fwd(const float* __restrict__ q, const float* __restrict__ kv,
                                                           const uint8_t* __restrict__ key_valid, int eye, int S, int T,
                                                           float* __restrict__ o, float* __restrict__ p, uint8_t* __restrict__ alive,
                                                           Drop drop) {
  __shared__ float vs[TR_H][32][TR_DH + 1];
  const uint32_t dkey = drop_key(drop);
  const int b = blockIdx.x, h = threadIdx.x / 32, lane = threadIdx.x % 32;
  const bool has_key = lane < T;
  float kr[TR_DH];
  if (has_key) {
    const float4* kp = reinterpret_cast<const float4*>(kv + ((long)b * T + lane) * 2 * TR_D + h * TR_DH);
    const float4* vp = reinterpret_cast<const float4*>(kv + ((long)b * T + lane) * 2 * TR_D + TR_D + h * TR_DH);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float4 t4 = kp[u], v4 = vp[u];
      kr[u * 4] = t4.x, kr[u * 4 + 1] = t4.y, kr[u * 4 + 2] = t4.z, kr[u * 4 + 3] = t4.w;
      vs[h][lane][u * 4] = v4.x, vs[h][lane][u * 4 + 1] = v4.y, vs[h][lane][u * 4 + 2] = v4.z, vs[h][lane][u * 4 + 3] = v4.w;
    }
  } else {
#pragma unroll
    for (int d = 0; d < TR_DH; ++d) kr[d] = 0.f;
  }
  const bool kvld = has_key && key_valid[(long)b * T + lane] != 0;
  __syncwarp();
  const float scale = 0.17677669529663688110f;
  for (int s = 0; s < S; ++s) {
    const long row = (long)b * S + s;
    const float qv = q[row * TR_D + h * TR_DH + lane];
    float dot = 0.f;
#pragma unroll
    for (int d = 0; d < TR_DH; ++d) dot = fmaf(__shfl_sync(0xffffffffu, qv, d), kr[d], dot);
    const bool ok = kvld && !(eye && lane == s);
    const float lg = ok ? dot * scale : -INFINITY;
    float m = lg;
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, of));
    float e = (ok && m > -INFINITY) ? expf(lg - m) : 0.f;
    float sum = e;
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, of);
    const float pv = m > -INFINITY ? e * (1.f / sum) : 0.f;
    const long pidx = (((long)b * TR_H + h) * S + s) * T + lane;
    if (has_key) p[pidx] = pv;
    if (h == 0 && lane == 0) alive[row] = m > -INFINITY ? 1 : 0;
    const float pd = (drop.seed && has_key) ? pv * drop_factor(drop, dkey, pidx) : pv;
    float acc = 0.f;
    for (int j = 0; j < T; ++j) acc = fmaf(__shfl_sync(0xffffffffu, pd, j), vs[h][j][lane], acc);
    o[row * TR_D + h * TR_DH + lane] = acc;
  }
}

// backward of the same: lane = key holds its V row and the dK / dV accumulators in registers, K in shared memory; dQ of a query
// is complete within the warp (plain stores, no atomics)
__global__ void __launch_bounds__(128) k_tr_attn_small_bwd(const float* __restrict__ dout, const float* __restrict__ q,
                                                           const float* __restrict__ kv, const float* __restrict__ p,
                                                           const float* __restrict__ o, int S, int T, float* __restrict__ dq,
                                                           float* __restrict__ dkv, Drop drop) {
  __shared__ float ks[TR_H][32][TR_DH + 1];
  const uint32_t dkey = drop_key(drop);
  const int b = blockIdx.x, h = threadIdx.x / 32, lane = threadIdx.x % 32;
  const bool has_key = lane < T;
  float vr[TR_DH], dk[TR_DH], dv[TR_DH];
#pragma unroll
  for (int d = 0; d < TR_DH; ++d) vr[d] = 0.f, dk[d] = 0.f, dv[d] = 0.f;
  if (has_key) {
    const float4* kp = reinterpret_cast<const float4*>(kv + ((long)b * T + lane) * 2 * TR_D + h * TR_DH);
    const float4* vp = reinterpret_cast<const float4*>(kv + ((long)b * T + lane) * 2 * TR_D + TR_D + h * TR_DH);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float4 t4 = kp[u], v4 = vp[u];
      ks[h][lane][u * 4] = t4.x, ks[h][lane][u * 4 + 1] = t4.y, ks[h][lane][u * 4 + 2] = t4.z, ks[h][lane][u * 4 + 3] = t4.w;
      vr[u * 4] = v4.x, vr[u * 4 + 1] = v4.y, vr[u * 4 + 2] = v4.z, vr[u * 4 + 3] = v4.w;
    }
  }
  __syncwarp();
  const float scale = 0.17677669529663688110f;
  for (int s = 0; s < S; ++s) {
    const long idx = ((long)b * S + s) * TR_D + h * TR_DH + lane;
    const float gv = dout[idx], ov = o[idx], qv = q[idx];
    float delta = gv * ov;
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) delta += __shfl_xor_sync(0xffffffffu, delta, of);
    float dp = 0.f;
#pragma unroll
    for (int d = 0; d < TR_DH; ++d) dp = fmaf(__shfl_sync(0xffffffffu, gv, d), vr[d], dp);
    const long pidx = (((long)b * TR_H + h) * S + s) * T + lane;
    const float pj = has_key ? p[pidx] : 0.f;
    const float pjf = (drop.seed && has_key) ? pj * drop_factor(drop, dkey, pidx) : pj;
    const float dsj = (pjf * dp - pj * delta) * scale;
#pragma unroll
    for (int d = 0; d < TR_DH; ++d) {
      dv[d] = fmaf(pjf, __shfl_sync(0xffffffffu, gv, d), dv[d]);
      dk[d] = fmaf(dsj, __shfl_sync(0xffffffffu, qv, d), dk[d]);
    }
    float acc = 0.f;
    for (int j = 0; j < T; ++j) acc = fmaf(__shfl_sync(0xffffffffu, dsj, j), ks[h][j][lane], acc);
    dq[idx] = acc;
  }
  if (has_key) {
    float4* okp = reinterpret_cast<float4*>(dkv + ((long)b * T + lane) * 2 * TR_D + h * TR_DH);
    float4* ovp = reinterpret_cast<float4*>(dkv + ((long)b * T + lane) * 2 * TR_D + TR_D + h * TR_DH);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      okp[u] = make_float4(dk[u * 4], dk[u * 4 + 1], dk[u * 4 + 2], dk[u * 4 + 3]);
      ovp[u] = make_float4(dv[u * 4], dv[u * 4 + 1], dv[u * 4 + 2], dv[u * 4 + 3]);
    }
  }
}

// =====================================================================================================================
// elementwise glue
// =====================================================================================================================
// y = (a * keep_a[row] + b) * keep[row]
__global__ void k_tr_add_mask(const float* __restrict__ a, const uint8_t* __restrict__ keep_a, const float* __restrict__ b,
                              const uint8_t* __restrict__ keep, long M, int N, float* __restrict__ y) {
  const long total = M * N;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long m = e / N;
    float v = a[e];
    if (keep_a && !keep_a[m]) v = 0.f;
    if (b) v += b[e];
    if (keep && !keep[m]) v = 0.f;
    y[e] = v;
  }
}

// y = x * dropout factor(site, element index): nn.GRU's inter-layer dropout; the same call is its own backward
__global__ void k_tr_dropout(const float* __restrict__ x, long n, float* __restrict__ y, Drop drop) {
  const uint32_t key = drop_key(drop);
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x)
    y[e] = x[e] * drop_factor(drop, key, e);
}

__global__ void k_tr_axpy(float* __restrict__ dst, long ld_dst, const float* __restrict__ src, long ld_src, long M, int N) {
  const long total = M * N;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long m = e / N;
    const int n = (int)(e % N);
    dst[m * ld_dst + n] += src[m * ld_src + n];
  }
}

__global__ void k_tr_select(const uint8_t* __restrict__ mask, const float* __restrict__ a, const float* __restrict__ b, long M,
                            int N, float* __restrict__ y) {
  const long total = M * N;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x)
    y[e] = mask[e / N] ? a[e] : b[e];
}

__global__ void k_tr_select_bwd(const uint8_t* __restrict__ mask, const float* __restrict__ dy, long M, int N,
                                float* __restrict__ da, float* __restrict__ db) {
  const long total = M * N;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const bool m = mask[e / N] != 0;
    const float g = dy[e];
    da[e] = m ? g : 0.f;
    db[e] = m ? 0.f : g;
  }
}

__global__ void k_tr_cat2(const float* __restrict__ a, int ka, const float* __restrict__ b, int kb, long M, float* __restrict__ y) {
  const int n = ka + kb;
  const long total = M * n;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long m = e / n;
    const int c = (int)(e % n);
    y[e] = c < ka ? a[m * ka + c] : b[m * kb + (c - ka)];
  }
}

__global__ void k_tr_cat2_bwd(const float* __restrict__ dy, int ka, int kb, long M, float* __restrict__ da, float* __restrict__ db) {
  const int n = ka + kb;
  const long total = M * n;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long m = e / n;
    const int c = (int)(e % n);
    if (c < ka) {
      if (da) da[m * ka + c] = dy[e];
    } else if (db) {
      db[m * kb + (c - ka)] = dy[e];
    }
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// torch nn.GRU cell, gate order r, z, n
__global__ void k_tr_gru_fwd(const float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ h, long M,
                             float* __restrict__ hn) {
  const long total = M * TR_D;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long m = e / TR_D;
    const int c = (int)(e % TR_D);
    const float* a = gi + m * 3 * TR_D;
    const float* b = gh + m * 3 * TR_D;
    const float r = sigmoidf_(a[c] + b[c]);
    const float z = sigmoidf_(a[TR_D + c] + b[TR_D + c]);
    const float n = tanhf(a[2 * TR_D + c] + r * b[2 * TR_D + c]);
    hn[e] = (1.f - z) * n + z * h[e];
  }
}

__global__ void k_tr_gru_bwd(const float* __restrict__ dhn, const float* __restrict__ gi, const float* __restrict__ gh,
                             const float* __restrict__ h, long M, float* __restrict__ dgi, float* __restrict__ dgh,
                             float* __restrict__ dh) {
  const long total = M * TR_D;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long m = e / TR_D;
    const int c = (int)(e % TR_D);
    const float* a = gi + m * 3 * TR_D;
    const float* b = gh + m * 3 * TR_D;
    const float hn_ = b[2 * TR_D + c];
    const float r = sigmoidf_(a[c] + b[c]);
    const float z = sigmoidf_(a[TR_D + c] + b[TR_D + c]);
    const float n = tanhf(a[2 * TR_D + c] + r * hn_);
    const float g = dhn[e];
    const float dn = g * (1.f - z);
    const float dz = g * (h[e] - n);
    dh[e] = g * z;
    const float da = dn * (1.f - n * n);
    const float dr = da * hn_;
    const float dzz = dz * z * (1.f - z);
    const float drr = dr * r * (1.f - r);
    float* oa = dgi + m * 3 * TR_D;
    float* ob = dgh + m * 3 * TR_D;
    oa[c] = drr;
    ob[c] = drr;
    oa[TR_D + c] = dzz;
    ob[TR_D + c] = dzz;
    oa[2 * TR_D + c] = da;
    ob[2 * TR_D + c] = da * r;
  }
}

// x [O,R,I,D], valid [O,R,I] -> y [O,I,D], idx [O,I,D]
__global__ void k_tr_masked_max(const float* __restrict__ x, const uint8_t* __restrict__ valid, long O, int R, long I, int D,
                                float fill, float* __restrict__ y, int32_t* __restrict__ idx) {
  const long total = O * I * D;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int d = (int)(e % D);
    const long i = (e / D) % I;
    const long o = e / ((long)D * I);
    float best = 0.f;
    int bi = -1;
    bool best_valid = false, any = false;
    for (int r = 0; r < R; ++r) {
      const bool v = valid[(o * R + r) * I + i] != 0;
      const float val = v ? x[((o * R + r) * I + i) * D + d] : fill;
      any |= v;
      if (bi < 0 || val > best) {
        best = val;
        bi = r;
        best_valid = v;
      }
    }
    y[e] = any ? best : 0.f;
    idx[e] = (any && best_valid) ? bi : -1;
  }
}

__global__ void k_tr_masked_max_bwd(const float* __restrict__ dy, const int32_t* __restrict__ idx, long O, int R, long I, int D,
                                    float* __restrict__ dx) {
  const long total = O * R * I * D;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int d = (int)(e % D);
    const long i = (e / D) % I;
    const int r = (int)((e / ((long)D * I)) % R);
    const long o = e / ((long)D * I * R);
    const long oe = (o * I + i) * D + d;
    dx[e] = idx[oe] == r ? dy[oe] : 0.f;
  }
}

__global__ void k_tr_gather(const float* __restrict__ x, const int64_t* __restrict__ idx, long M, int D, float* __restrict__ y) {
  const long total = M * D;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x)
    y[e] = x[idx[e / D] * D + e % D];
}

__global__ void k_tr_scatter_add(const float* __restrict__ dy, const int64_t* __restrict__ idx, long M, int D,
                                 float* __restrict__ dx) {
  const long total = M * D;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x)
    atomicAdd(&dx[idx[e / D] * D + e % D], dy[e]);
}

// y[s,a,p,:] = u[s,p,:] + v[s,a,:]
__global__ void k_tr_pair_add(const float* __restrict__ u, const float* __restrict__ v, int S, int P, int A, float* __restrict__ y) {
  const long total = (long)S * A * P * TR_D;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int d = (int)(e % TR_D);
    const long row = e / TR_D;
    const int pp = (int)(row % P);
    const long sa = row / P;
    const long s = sa / A;
    y[e] = u[(s * P + pp) * TR_D + d] + v[sa * TR_D + d];
  }
}

// du[s,p,d] = sum_a dy[s,a,p,d]
__global__ void k_tr_pair_du(const float* __restrict__ dy, int S, int P, int A, float* __restrict__ du) {
  const long total = (long)S * P * TR_D;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int d = (int)(e % TR_D);
    const long sp = e / TR_D;
    const int pp = (int)(sp % P);
    const long s = sp / P;
    float acc = 0.f;
    for (int a = 0; a < A; ++a) acc += dy[(((s * A + a) * P) + pp) * TR_D + d];
    du[e] = acc;
  }
}

// dv[s,a,d] = sum_p dy[s,a,p,d]; one CTA of 128 threads per (s,a), 4-way split of p by blockIdx.y with atomics
__global__ void __launch_bounds__(128) k_tr_pair_dv(const float* __restrict__ dy, int P, float* __restrict__ dv) {
  const long sa = blockIdx.x;
  const int d = threadIdx.x;
  const int chunk = (P + gridDim.y - 1) / gridDim.y;
  const int p_lo = blockIdx.y * chunk, p_hi = min(P, p_lo + chunk);
  float acc = 0.f;
  for (int pp = p_lo; pp < p_hi; ++pp) acc += dy[(sa * P + pp) * TR_D + d];
  atomicAdd(&dv[sa * TR_D + d], acc);
}

// destination NLL (goal_manager.py:328-333 masks + Categorical(logits) + metrics/training.py:138-147); one CTA per (s,a)
__global__ void __launch_bounds__(256) k_tr_dest_nll(const float* __restrict__ logits, const uint8_t* __restrict__ pair_ok,
                                                     const uint8_t* __restrict__ row_valid, const int64_t* __restrict__ gt,
                                                     const uint8_t* __restrict__ loss_rows, const float* __restrict__ scale,
                                                     int P, float* __restrict__ nll_sum, float* __restrict__ dlogits) {
  __shared__ float red[256];
  __shared__ int any_ok;
  const long row = blockIdx.x;
  const int tid = threadIdx.x;
  const float* lg = logits + row * P;
  const uint8_t* ok = pair_ok + row * P;
  const bool rv = row_valid[row] != 0;
  if (tid == 0) any_ok = 0;
  __syncthreads();
  int loc = 0;
  for (int j = tid; j < P; j += 256) loc |= ok[j];
  if (loc) any_ok = 1;
  __syncthreads();
  const bool uniform = !rv || !any_ok;  // all logits reset to 0 (:330-331): constant, no gradient
  float m = -INFINITY;
  for (int j = tid; j < P; j += 256) {
    const float v = uniform ? 0.f : (ok[j] ? lg[j] : -INFINITY);
    m = fmaxf(m, v);
  }
  red[tid] = m;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) red[tid] = fmaxf(red[tid], red[tid + s]);
    __syncthreads();
  }
  m = red[0];
  __syncthreads();
  float se = 0.f;
  for (int j = tid; j < P; j += 256) {
    const float v = uniform ? 0.f : (ok[j] ? lg[j] : -INFINITY);
    se += expf(v - m);
  }
  red[tid] = se;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) red[tid] += red[tid + s];
    __syncthreads();
  }
  const float lse = m + logf(red[0]);
  const bool counted = loss_rows[row] != 0;
  const long g = gt[row];
  const float sc = scale[0];
  for (int j = tid; j < P; j += 256) {
    float grad = 0.f;
    if (counted && !uniform && ok[j]) grad = (expf(lg[j] - lse) - (j == g ? 1.f : 0.f)) * sc;
    dlogits[row * P + j] = grad;
  }
  if (tid == 0 && counted) {
    const float vg = uniform ? 0.f : (ok[g] ? lg[g] : -INFINITY);
    atomicAdd(nll_sum, lse - vg);
  }
}

__global__ void k_tr_rsample(const float* __restrict__ mean, const float* __restrict__ log_std, const float* __restrict__ eps,
                             long M, int E, float* __restrict__ z) {
  const long total = M * E;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x)
    z[e] = mean[e] + eps[e] * expf(log_std[e % E]);
}

__global__ void k_tr_rsample_bwd(const float* __restrict__ dz, const float* __restrict__ eps, const float* __restrict__ log_std,
                                 long M, int E, float* __restrict__ dlog_std) {
  // one thread per latent dim e: dlog_std[e] += sum_m dz * eps * std
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const float sd = expf(log_std[e]);
  float acc = 0.f;
  for (long m = 0; m < M; ++m) acc += dz[m * E + e] * eps[m * E + e] * sd;
  atomicAdd(&dlog_std[e], acc);  // several sub-batches of a step may run concurrently
}

// KL(N(mq, e^lq) || N(mp, e^lp)) per row = sum_e [lp - lq + (e^2lq + (mq - mp)^2) / (2 e^2lp) - 1/2], clamped at free_nats
__global__ void k_tr_kl(const float* __restrict__ mq, const float* __restrict__ lq, const float* __restrict__ mp,
                        const float* __restrict__ lp, const uint8_t* __restrict__ valid, float free_nats,
                        const float* __restrict__ scale, long M, int E, float* __restrict__ kl_sum, float* __restrict__ dmq,
                        float* __restrict__ dmp, float* __restrict__ dlq, float* __restrict__ dlp) {
  const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float kl = 0.f;
  for (int e = 0; e < E; ++e) {
    const float vq = expf(2.f * lq[e]), vp = expf(2.f * lp[e]);
    const float dm = mq[m * E + e] - mp[m * E + e];
    kl += lp[e] - lq[e] + (vq + dm * dm) / (2.f * vp) - 0.5f;
  }
  const bool on = valid[m] != 0;
  const bool active = on && !(free_nats > 0.f && kl < free_nats);  // torch.max(kl, free_nats): gradient only where kl wins
  const float sc = scale[0];
  if (on) atomicAdd(kl_sum, (free_nats > 0.f && kl < free_nats) ? free_nats : kl);
  for (int e = 0; e < E; ++e) {
    float gq = 0.f, gp = 0.f;
    if (active) {
      const float vq = expf(2.f * lq[e]), vp = expf(2.f * lp[e]);
      const float dm = mq[m * E + e] - mp[m * E + e];
      gq = dm / vp * sc;
      gp = -gq;
      atomicAdd(&dlq[e], (-1.f + vq / vp) * sc);
      atomicAdd(&dlp[e], (1.f - (vq + dm * dm) / vp) * sc);
    }
    dmq[m * E + e] = gq;
    dmp[m * E + e] = gp;
  }
}

// [cos(x f_even), sin(x f_odd), cos(y f_even), sin(y f_odd), cos(yaw g_even), sin(yaw g_odd)] (utils/pose_pe.py:57-62)
__global__ void k_tr_pose_pe(const float* __restrict__ xy, const float* __restrict__ yaw, const float* __restrict__ f_xy,
                             int n_xy, const float* __restrict__ f_yaw, int n_yaw, long M, float* __restrict__ pe) {
  const int W = 2 * n_xy + n_yaw;
  const long total = M * W;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long m = e / W;
    int c = (int)(e % W);
    float v;
    const float* f;
    int n;
    if (c < n_xy) {
      v = xy[m * 2], f = f_xy, n = n_xy;
    } else if (c < 2 * n_xy) {
      v = xy[m * 2 + 1], f = f_xy, n = n_xy, c -= n_xy;
    } else {
      v = yaw[m], f = f_yaw, n = n_yaw, c -= 2 * n_xy;
    }
    const int half = n / 2;
    pe[e] = c < half ? cosf(v * f[2 * c]) : sinf(v * f[2 * (c - half) + 1]);
  }
}

__global__ void k_tr_dir_to_yaw(const float* __restrict__ d, long M, float* __restrict__ yaw) {
  for (long m = (long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long)gridDim.x * blockDim.x)
    yaw[m] = atan2f(d[m * 2 + 1], d[m * 2]);
}

__constant__ float c_max_acc[3] = {5.0f, 7.0f, 6.0f};       // veh, ped, cyc (traffic_bots.yaml:142-155; dynamics.py:23-27)
__constant__ float c_max_yaw_rate[3] = {1.5f, 7.0f, 3.0f};

// Dynamics.update with the deterministic action (utils/dynamics.py:74-119) + MultiPathPP (:187-228); bwd when dpred != NULL
__global__ void k_tr_dynamics(const float* __restrict__ state, const float* __restrict__ mean, const uint8_t* __restrict__ a_type,
                              const uint8_t* __restrict__ valid, long M, float* __restrict__ pred, const float* __restrict__ dpred,
                              float* __restrict__ dstate, float* __restrict__ dmean) {
  const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float dt = 0.1f;
  float mx_a = 0.f, mx_y = 0.f;
  bool has_type = false;
  for (int c = 0; c < 3; ++c)
    if (a_type[m * 3 + c]) {
      mx_a += c_max_acc[c];
      mx_y += c_max_yaw_rate[c];
      has_type = true;
    }
  const float keep = valid[m] ? 1.f : 0.f;
  const float4 s = *reinterpret_cast<const float4*>(state + m * 4);
  const float th0 = tanhf(mean[m * 2]), th1 = tanhf(mean[m * 2 + 1]);
  const float a = th0 * mx_a * keep, w = th1 * mx_y * keep;
  const float vt = s.w + 0.5f * dt * a, tt = s.z + 0.5f * dt * w;
  const float ct = cosf(tt), st = sinf(tt);
  const float msk = has_type ? keep : 0.f;
  if (dpred == nullptr) {
    *reinterpret_cast<float4*>(pred + m * 4) =
        make_float4((s.x + dt * vt * ct) * msk, (s.y + dt * vt * st) * msk, (s.z + dt * w) * msk, (s.w + dt * a) * msk);
    return;
  }
  const float4 g4 = *reinterpret_cast<const float4*>(dpred + m * 4);
  const float gx = g4.x * msk, gy = g4.y * msk, gth = g4.z * msk, gv = g4.w * msk;
  const float dvt = gx * dt * ct + gy * dt * st;
  const float dtt = -gx * dt * vt * st + gy * dt * vt * ct;
  *reinterpret_cast<float4*>(dstate + m * 4) = make_float4(gx, gy, gth + dtt, gv + dvt);
  const float da = gv * dt + dvt * 0.5f * dt, dw = gth * dt + dtt * 0.5f * dt;
  dmean[m * 2] = da * mx_a * keep * (1.f - th0 * th0);
  dmean[m * 2 + 1] = dw * mx_y * keep * (1.f - th1 * th1);
}

__device__ __forceinline__ float smooth_l1_(float d) {
  const float a = fabsf(d);
  return a < 1.f ? 0.5f * d * d : a - 0.5f;
}
__device__ __forceinline__ float smooth_l1_grad_(float d) { return fabsf(d) < 1.f ? d : (d > 0.f ? 1.f : -1.f); }

// DifferentiableReward.get, IL part (utils/rewards.py:117-131); bwd when dr != NULL
__global__ void k_tr_reward(const float* __restrict__ pred, const float* __restrict__ gt, const uint8_t* __restrict__ rv, long M,
                            float* __restrict__ r, const float* __restrict__ dr, float* __restrict__ dpred) {
  const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const bool on = rv[m] != 0;
  const float4 p = *reinterpret_cast<const float4*>(pred + m * 4);
  const float4 g = *reinterpret_cast<const float4*>(gt + m * 4);
  if (dr == nullptr) {
    float v = 0.f;
    if (on) {
      const float e_pos = smooth_l1_(g.x - p.x) + smooth_l1_(g.y - p.y);
      const float e_rot = 0.5f * (1.f - cosf(g.z - p.z));
      const float e_spd = smooth_l1_(g.w - p.w);
      v = 0.f - (0.1f * e_pos + 10.f * e_rot + 0.1f * e_spd);
    }
    r[m] = v;
    return;
  }
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (on) {
    const float d = dr[m];
    o = make_float4(d * 0.1f * smooth_l1_grad_(g.x - p.x), d * 0.1f * smooth_l1_grad_(g.y - p.y), d * 5.f * sinf(g.z - p.z),
                    d * 0.1f * smooth_l1_grad_(g.w - p.w));
  }
  *reinterpret_cast<float4*>(dpred + m * 4) = o;
}

// rule check on the post-override state (always-on subset that feeds back), kill, disable_goal_reached
// (pl_modules/waymo_motion.py:311-320; utils/traffic_rule_checker.py:101-119,364-410; utils/dynamics.py:151-167;
//  models/goal_manager.py:155-161)
__global__ void k_tr_sim_flags(const float* __restrict__ state, const uint8_t* __restrict__ valid, const uint8_t* __restrict__ gt_valid,
                               const float* __restrict__ boundary, const float* __restrict__ dest_pos,
                               const float* __restrict__ dest_dir, const uint8_t* __restrict__ dest_valid,
                               const uint8_t* __restrict__ dest_lane, const uint8_t* __restrict__ dest_edge,
                               const uint8_t* __restrict__ killed, const uint8_t* __restrict__ dest_reached,
                               const uint8_t* __restrict__ goal_valid, int B, int A, uint8_t* __restrict__ o_valid,
                               uint8_t* __restrict__ o_killed, uint8_t* __restrict__ o_dest, uint8_t* __restrict__ o_goal) {
  const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= (long)B * A) return;
  const int bq = (int)(m / A);
  const float4 s = *reinterpret_cast<const float4*>(state + m * 4);
  bool v = valid[m] != 0;
  const float* mb = boundary + (long)bq * 4;
  const bool out_t = v && (s.x > mb[1] || s.x < mb[0] || s.y > mb[3] || s.y < mb[2]);
  const bool edge_t = dest_edge[m] != 0, lane_t = dest_lane[m] != 0;
  const float thresh = 50.0f * (1.0f - (edge_t ? 1.0f : 0.f) * 0.8f);
  bool pos_reached = false, rot_reached = false;
  const float hx = cosf(s.z), hy = sinf(s.z);
  for (int n = 0; n < TB_PL_NODE; ++n) {
    const long nd = m * TB_PL_NODE + n;
    if (!dest_valid[nd]) continue;
    const float dx = __fsub_rn(s.x, dest_pos[nd * 2]), dy = __fsub_rn(s.y, dest_pos[nd * 2 + 1]);
    pos_reached |= sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))) < thresh;
    const float ux = dest_dir[nd * 2], uy = dest_dir[nd * 2 + 1];
    const float nrm = sqrtf(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)));
    rot_reached |= __fadd_rn(__fmul_rn(hx, ux / nrm), __fmul_rn(hy, uy / nrm)) > 0.86602540378443864676f;
  }
  bool dr_ = dest_reached[m] != 0;
  const bool dest_t = !dr_ && v && ((lane_t && pos_reached && rot_reached) || (edge_t && pos_reached));
  dr_ |= dest_t;
  const bool kill = out_t && !(gt_valid && gt_valid[m]);
  v = v && !kill;
  o_valid[m] = v;
  o_killed[m] = (killed[m] != 0) || kill;
  o_dest[m] = dr_;
  o_goal[m] = goal_valid[m] && v && !dr_;
}

__global__ void k_tr_masked_sum(const float* __restrict__ x, const uint8_t* __restrict__ mask, long n, float* __restrict__ out) {
  __shared__ float red[256];
  float s = 0.f;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x)
    if (mask[e]) s += x[e];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out, red[0]);
}

__global__ void k_tr_mask_scale(const uint8_t* __restrict__ mask, const float* __restrict__ scale, long n, float* __restrict__ out) {
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x)
    out[e] = mask[e] ? scale[0] : 0.f;
}

__global__ void k_tr_scale(float* __restrict__ x, long n, float alpha) {
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) x[e] *= alpha;
}

__global__ void k_tr_sq_norm(const float* __restrict__ g, long n, float* __restrict__ out) {
  __shared__ float red[256];
  float s = 0.f;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) s = fmaf(g[e], g[e], s);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out, red[0]);
}

// torch.optim.Adam (no weight decay, no amsgrad) on the flat parameter buffer; the gradient is first multiplied by
// clip_coef = min(1, max_norm / (||g|| + 1e-6)) (torch.nn.utils.clip_grad_norm_), computed on the device from sq_norm
__global__ void k_tr_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n,
                          const float* __restrict__ lr_by_group, const int32_t* __restrict__ group_end, int n_group, float beta1,
                          float beta2, float eps, float bc1, float bc2_sqrt, const float* __restrict__ sq_norm, float max_norm) {
  float coef = 1.f;
  if (sq_norm && max_norm > 0.f) coef = fminf(1.f, max_norm / (sqrtf(sq_norm[0]) + 1e-6f));
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    int gi = 0;
    while (gi + 1 < n_group && e >= group_end[gi]) ++gi;
    const float gr = g[e] * coef;
    const float mm = beta1 * m[e] + (1.f - beta1) * gr;
    const float vv = beta2 * v[e] + (1.f - beta2) * gr * gr;
    m[e] = mm;
    v[e] = vv;
    p[e] -= lr_by_group[gi] / bc1 * mm / (sqrtf(vv) / bc2_sqrt + eps);
  }
}

inline long tc_min_m() {  // TB_TRAIN_TC_MIN_M overrides the row threshold of the tensor-core Linears (A/B runs)
  static const long v = [] {
    const char* e = getenv("TB_TRAIN_TC_MIN_M");
    return e ? atol(e) : TC_MIN_M;
  }();
  return v;
}

inline int grid_for(long total, int block = 256) {
  long g = (total + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > 148L * 16 ? 148L * 16 : g));
}

}  // namespace
}  // namespace tb

using namespace tb;

#define TR_CHECK(cond, code) \
  do {                       \
    if (!(cond)) return code; \
  } while (0)

extern "C" {

// y = (dropout(relu(x W^T + bias) * keep_lin[row]) + res) * keep_out[row]; bias / keep_lin / res / keep_out / drop_seed may be NULL
int32_t tb_tr_linear_fwd(const float* x, int64_t M, int32_t K, const float* w, int64_t ldw, int32_t N, const float* bias, int32_t relu,
                         const uint8_t* keep_lin, const float* res, const uint8_t* keep_out, float* y, const uint32_t* drop_seed,
                         uint32_t drop_site, float drop_p, int64_t drop_offset, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(x && w && y, TB_ERR_NULL);
  TR_CHECK(M > 0 && K > 0 && N > 0 && ldw >= K, TB_ERR_BAD_SHAPE);
  GemmOpt op{nullptr, nullptr, nullptr, bias, relu, keep_lin, res, keep_out, nullptr, make_drop(drop_seed, drop_site, drop_p, drop_offset)};
  if (M >= tc_min_m() && K % 128 == 0 && K <= 256 && N % 128 == 0 && (ldw & 3) == 0 && aligned16(x) && aligned16(w) && aligned16(y) &&
      (!res || aligned16(res)) && train_tc_enabled())
    return launch_train_linear_tc_fwd(x, M, K, w, ldw, N, bias, relu, keep_lin, res, keep_out, y, op.drop.seed, op.drop.site,
                                      op.drop.thresh, op.drop.scale, op.drop.offset, st);
  if (M <= SMALL_M) {
    dim3 grid((unsigned)((M + 31) / 32), (N + GN - 1) / GN, 1);
    k_tr_gemm<32, true, false, EPI_BIAS><<<grid, 256, 0, st>>>(x, K, 1, w, 1, ldw, y, N, M, N, K, K, op);
  } else {
    dim3 grid((unsigned)((M + GM - 1) / GM), (N + GN - 1) / GN, 1);
    k_tr_gemm<GM, true, false, EPI_BIAS><<<grid, 256, 0, st>>>(x, K, 1, w, 1, ldw, y, N, M, N, K, K, op);
  }
  count_launch();
  return launch_status();
}

// dx = dY' W; dw += dY'^T x; db += colsum(dY') with dY' = dy * relu'(y) * dropout mask * rm1[row] * rm2[row] (row masks may be
// NULL: the keep_lin / keep_out of the forward; drop_* as in the forward).  dx / dw / db may be NULL (skipped); db needs dw.
int32_t tb_tr_linear_bwd(const float* dy, const float* x, const float* w, int64_t ldw, const float* y, int32_t relu, const uint8_t* rm1,
                         const uint8_t* rm2, int64_t M, int32_t K, int32_t N, float* dx, float* dw, int64_t lddw, float* db,
                         const uint32_t* drop_seed, uint32_t drop_site, float drop_p, int64_t drop_offset, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(dy && x && w, TB_ERR_NULL);
  TR_CHECK(M > 0 && K > 0 && N > 0 && (!relu || y), TB_ERR_BAD_SHAPE);
  TR_CHECK(dw || !db, TB_ERR_BAD_SHAPE);  // the bias gradient rides on the weight-gradient tiles
  if (!dx && !dw) return TB_OK;
  if (M >= tc_min_m() && K % 128 == 0 && K <= 256 && N % 128 == 0 && N <= 384 && (ldw & 3) == 0 && aligned16(dy) && aligned16(x) &&
      aligned16(w) && (!dx || aligned16(dx)) && (!relu || aligned16(y)) && train_tc_enabled()) {
    const Drop d = make_drop(drop_seed, drop_site, drop_p, drop_offset);
    int rc = TB_OK;
    if (dx)
      rc = launch_train_linear_tc_dx(dy, M, K, N, w, ldw, relu ? y : nullptr, rm1, rm2, dx, d.seed, d.site, d.thresh, d.scale, d.offset,
                                     st);
    if (rc == TB_OK && dw)
      rc = launch_train_linear_tc_dw(dy, x, M, K, N, relu ? y : nullptr, rm1, rm2, dw, lddw, db, d.seed, d.site, d.thresh, d.scale,
                                     d.offset, st);
    return rc;
  }
  const int TMh = M <= SMALL_M ? 32 : GM;
  const int dx_tiles_j = (K + GN - 1) / GN;
  const long n_dx = dx ? ((M + TMh - 1) / TMh) * dx_tiles_j : 0;
  const int dw_tiles_i = (N + TMh - 1) / TMh, dw_tiles_j = (K + GN - 1) / GN;
  long nz = 0, chunk = 0;
  if (dw) {  // enough row chunks to fill the GPU twice, at least 128 rows each
    const long tiles = (long)dw_tiles_i * dw_tiles_j;
    nz = (296 + tiles - 1) / tiles;
    const long nz_max = (M + 127) / 128;
    if (nz > nz_max) nz = nz_max;
    if (nz < 1) nz = 1;
    chunk = ((M + nz - 1) / nz + GK - 1) / GK * GK;
    nz = (M + chunk - 1) / chunk;
  }
  const long total = n_dx + nz * dw_tiles_i * dw_tiles_j;
  TR_CHECK(total > 0 && total < 2147483647L, TB_ERR_BAD_SHAPE);
  GemmOpt op{relu ? y : nullptr, rm1, rm2, nullptr, 0, nullptr, nullptr, nullptr, db, make_drop(drop_seed, drop_site, drop_p, drop_offset)};
  if (TMh == 32)
    k_tr_linear_bwd<32><<<(unsigned)total, 256, 0, st>>>(dy, x, w, ldw, M, K, N, dx, dw, lddw, (int)n_dx, dx_tiles_j, dw_tiles_i,
                                                         dw_tiles_j, chunk, op);
  else
    k_tr_linear_bwd<GM><<<(unsigned)total, 256, 0, st>>>(dy, x, w, ldw, M, K, N, dx, dw, lddw, (int)n_dx, dx_tiles_j, dw_tiles_i,
                                                         dw_tiles_j, chunk, op);
  count_launch();
  return launch_status();
}

int32_t tb_tr_layernorm_fwd(const float* x, const float* w, const float* b, int32_t relu, int64_t M, int32_t D, float* y, float* stats,
                        const uint32_t* drop_seed, uint32_t drop_site,
                        float drop_p, int64_t drop_offset, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(x && w && b && y && stats, TB_ERR_NULL);
  TR_CHECK(M > 0 && D == TR_D, TB_ERR_BAD_SHAPE);
  TR_CHECK(aligned16(x) && aligned16(y) && aligned16(w) && aligned16(b), TB_ERR_ALIGN);
  k_tr_ln_fwd<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(x, w, b, relu, M, y, stats, make_drop(drop_seed, drop_site, drop_p, drop_offset));
  count_launch();
  return launch_status();
}

int32_t tb_tr_layernorm_bwd(const float* dy, const float* x, const float* w, const float* stats, const float* y, int32_t relu, int64_t M,
                        int32_t D, float* dx, float* dw, float* db, const uint32_t* drop_seed, uint32_t drop_site,
                        float drop_p, int64_t drop_offset, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(dy && x && w && stats && dx, TB_ERR_NULL);
  TR_CHECK(M > 0 && D == TR_D && (!relu || y) && ((dw == nullptr) == (db == nullptr)), TB_ERR_BAD_SHAPE);
  TR_CHECK(aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(w), TB_ERR_ALIGN);
  const long rows = 64;
  k_tr_ln_bwd<<<(unsigned)((M + rows - 1) / rows), 256, 0, st>>>(dy, x, w, stats, y, relu, M, rows, dx, dw, db,
                                                                 make_drop(drop_seed, drop_site, drop_p, drop_offset));
  count_launch();
  return launch_status();
}

// alive[b, s] = 0 for rows without any admissible key (their o and p are 0: attention.py:101-107,144-146), else 1
int32_t tb_tr_attention_fwd(const float* q, const float* kv, const uint8_t* key_valid, int32_t eye, int32_t B, int32_t S, int32_t T,
                        float* o, float* p, uint8_t* alive, const uint32_t* drop_seed, uint32_t drop_site,
                        float drop_p, int64_t drop_offset, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(q && kv && key_valid && o && p && alive, TB_ERR_NULL);
  TR_CHECK(B > 0 && S > 0 && T > 0 && T <= 6144 && (!eye || S == T) && B <= 65535 * 1024, TB_ERR_BAD_SHAPE);
  TR_CHECK(aligned16(kv), TB_ERR_ALIGN);
  if (T <= 32) {  // one warp per (batch element, head)
    k_tr_attn_small_fwd<<<B, 128, 0, st>>>(q, kv, key_valid, eye, S, T, o, p, alive, make_drop(drop_seed, drop_site, drop_p, drop_offset));
    count_launch();
    return launch_status();
  }
  const int smem = (AT_QT * T > 4 * AT_QT * 32 ? AT_QT * T : 4 * AT_QT * 32) * (int)sizeof(float);
  static std::atomic<uint64_t> attr_set{0};
  if (smem > 48 * 1024 && !smem_attr_done(attr_set)) {
    if (!set_max_smem(k_tr_attn_fwd, 200 * 1024)) return TB_ERR_LAUNCH;
    smem_attr_mark(attr_set);
  }
  for (int b0 = 0; b0 < B; b0 += 65535) {  // gridDim.z limit
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    dim3 grid((S + AT_QT - 1) / AT_QT, TR_H, nb);
    k_tr_attn_fwd<<<grid, 128, smem, st>>>(q + (size_t)b0 * S * TR_D, kv + (size_t)b0 * T * 2 * TR_D, key_valid + (size_t)b0 * T, eye, S,
                                           T, o + (size_t)b0 * S * TR_D, p + (size_t)b0 * TR_H * S * T, alive + (size_t)b0 * S,
                                           make_drop(drop_seed, drop_site, drop_p, drop_offset), (long)b0 * TR_H * S * T);
    count_launch();
  }
  return launch_status();
}

// dq must be zero-initialised by the caller (partials are added atomically); dkv is overwritten -- or, with kv_batch > 0 (K|V of
// batch element b = those of b % kv_batch; kv / dkv hold kv_batch elements), accumulated atomically into a zero-initialised buffer
int32_t tb_tr_attention_bwd(const float* dout, const float* q, const float* kv, const float* p, const float* o, int32_t B, int32_t S,
                            int32_t T, int32_t kv_batch, float* dq, float* dkv, const uint32_t* drop_seed, uint32_t drop_site,
                            float drop_p, int64_t drop_offset, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(dout && q && kv && p && o && dq && dkv, TB_ERR_NULL);
  TR_CHECK(B > 0 && S > 0 && T > 0 && kv_batch >= 0 && (kv_batch == 0 || B % kv_batch == 0), TB_ERR_BAD_SHAPE);
  if (T <= 32) {
    TR_CHECK(kv_batch == 0, TB_ERR_UNSUPPORTED);
    TR_CHECK(aligned16(kv) && aligned16(dkv), TB_ERR_ALIGN);
    k_tr_attn_small_bwd<<<B, 128, 0, st>>>(dout, q, kv, p, o, S, T, dq, dkv, make_drop(drop_seed, drop_site, drop_p, drop_offset));
    count_launch();
    return launch_status();
  }
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    dim3 grid((T + AB_KC - 1) / AB_KC, TR_H, nb);
    k_tr_attn_bwd<<<grid, 256, 0, st>>>(dout, q, kv, p, o, S, T, dq, dkv, make_drop(drop_seed, drop_site, drop_p, drop_offset), b0,
                                        kv_batch);
    count_launch();
  }
  return launch_status();
}

// y = (a * keep_a[row] + b) * keep[row]; keep_a / b / keep may be NULL
int32_t tb_tr_add_mask(const float* a, const uint8_t* keep_a, const float* b, const uint8_t* keep, int64_t M, int32_t N, float* y,
                       void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(a && y, TB_ERR_NULL);
  TR_CHECK(M > 0 && N > 0, TB_ERR_BAD_SHAPE);
  k_tr_add_mask<<<grid_for(M * N), 256, 0, st>>>(a, keep_a, b, keep, M, N, y);
  count_launch();
  return launch_status();
}

// y = x * dropout factor (inter-layer dropout of nn.GRU, agent_temporal.py:116); applied to dy it is its own backward
int32_t tb_tr_dropout(const float* x, int64_t n, float* y, const uint32_t* drop_seed, uint32_t drop_site, float drop_p, int64_t drop_offset, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(x && y && drop_seed, TB_ERR_NULL);
  k_tr_dropout<<<grid_for(n), 256, 0, st>>>(x, n, y, make_drop(drop_seed, drop_site, drop_p, drop_offset));
  count_launch();
  return launch_status();
}

int32_t tb_tr_axpy(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t M, int32_t N, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(dst && src, TB_ERR_NULL);
  TR_CHECK(M > 0 && N > 0, TB_ERR_BAD_SHAPE);
  k_tr_axpy<<<grid_for(M * N), 256, 0, st>>>(dst, ld_dst, src, ld_src, M, N);
  count_launch();
  return launch_status();
}

int32_t tb_tr_select_rows(const uint8_t* mask, const float* a, const float* b, int64_t M, int32_t N, float* y, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(mask && a && b && y, TB_ERR_NULL);
  k_tr_select<<<grid_for(M * N), 256, 0, st>>>(mask, a, b, M, N, y);
  count_launch();
  return launch_status();
}

int32_t tb_tr_select_rows_bwd(const uint8_t* mask, const float* dy, int64_t M, int32_t N, float* da, float* db, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(mask && dy && da && db, TB_ERR_NULL);
  k_tr_select_bwd<<<grid_for(M * N), 256, 0, st>>>(mask, dy, M, N, da, db);
  count_launch();
  return launch_status();
}

int32_t tb_tr_cat2(const float* a, int32_t ka, const float* b, int32_t kb, int64_t M, float* y, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(a && b && y, TB_ERR_NULL);
  k_tr_cat2<<<grid_for(M * (ka + kb)), 256, 0, st>>>(a, ka, b, kb, M, y);
  count_launch();
  return launch_status();
}

int32_t tb_tr_cat2_bwd(const float* dy, int32_t ka, int32_t kb, int64_t M, float* da, float* db, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(dy, TB_ERR_NULL);
  k_tr_cat2_bwd<<<grid_for(M * (ka + kb)), 256, 0, st>>>(dy, ka, kb, M, da, db);
  count_launch();
  return launch_status();
}

int32_t tb_tr_gru_gates_fwd(const float* gi, const float* gh, const float* h, int64_t M, float* hn, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(gi && gh && h && hn, TB_ERR_NULL);
  k_tr_gru_fwd<<<grid_for(M * TR_D), 256, 0, st>>>(gi, gh, h, M, hn);
  count_launch();
  return launch_status();
}

int32_t tb_tr_gru_gates_bwd(const float* dhn, const float* gi, const float* gh, const float* h, int64_t M, float* dgi, float* dgh, float* dh,
                        void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(dhn && gi && gh && h && dgi && dgh && dh, TB_ERR_NULL);
  k_tr_gru_bwd<<<grid_for(M * TR_D), 256, 0, st>>>(dhn, gi, gh, h, M, dgi, dgh, dh);
  count_launch();
  return launch_status();
}

int32_t tb_tr_masked_max_fwd(const float* x, const uint8_t* valid, int64_t O, int32_t R, int64_t I, int32_t D, float fill, float* y,
                         int32_t* idx, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(x && valid && y && idx, TB_ERR_NULL);
  k_tr_masked_max<<<grid_for(O * I * D), 256, 0, st>>>(x, valid, O, R, I, D, fill, y, idx);
  count_launch();
  return launch_status();
}

int32_t tb_tr_masked_max_bwd(const float* dy, const int32_t* idx, int64_t O, int32_t R, int64_t I, int32_t D, float* dx, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(dy && idx && dx, TB_ERR_NULL);
  k_tr_masked_max_bwd<<<grid_for(O * R * I * D), 256, 0, st>>>(dy, idx, O, R, I, D, dx);
  count_launch();
  return launch_status();
}

int32_t tb_tr_gather_rows(const float* x, const int64_t* idx, int64_t M, int32_t D, float* y, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(x && idx && y, TB_ERR_NULL);
  k_tr_gather<<<grid_for(M * D), 256, 0, st>>>(x, idx, M, D, y);
  count_launch();
  return launch_status();
}

// dx must be zero-initialised by the caller
int32_t tb_tr_scatter_add_rows(const float* dy, const int64_t* idx, int64_t M, int32_t D, float* dx, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(dy && idx && dx, TB_ERR_NULL);
  k_tr_scatter_add<<<grid_for(M * D), 256, 0, st>>>(dy, idx, M, D, dx);
  count_launch();
  return launch_status();
}

int32_t tb_tr_pair_add(const float* u, const float* v, int32_t S, int32_t P, int32_t A, float* y, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(u && v && y, TB_ERR_NULL);
  k_tr_pair_add<<<grid_for((long)S * A * P * TR_D), 256, 0, st>>>(u, v, S, P, A, y);
  count_launch();
  return launch_status();
}

// dv must be zero-initialised by the caller
int32_t tb_tr_pair_add_bwd(const float* dy, int32_t S, int32_t P, int32_t A, float* du, float* dv, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(dy && du && dv, TB_ERR_NULL);
  k_tr_pair_du<<<grid_for((long)S * P * TR_D), 256, 0, st>>>(dy, S, P, A, du);
  count_launch();
  k_tr_pair_dv<<<dim3(S * A, 4), 128, 0, st>>>(dy, P, dv);
  count_launch();
  return launch_status();
}

// nll_sum [1] must be zero-initialised by the caller
int32_t tb_tr_dest_nll(const float* logits, const uint8_t* pair_ok, const uint8_t* row_valid, const int64_t* gt, const uint8_t* loss_rows,
                   const float* scale, int64_t n_row, int32_t P, float* nll_sum, float* dlogits, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(logits && pair_ok && row_valid && gt && loss_rows && scale && nll_sum && dlogits, TB_ERR_NULL);
  k_tr_dest_nll<<<(unsigned)n_row, 256, 0, st>>>(logits, pair_ok, row_valid, gt, loss_rows, scale, P, nll_sum, dlogits);
  count_launch();
  return launch_status();
}

int32_t tb_tr_rsample(const float* mean, const float* log_std, const float* eps, int64_t M, int32_t E, float* z, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(mean && log_std && eps && z, TB_ERR_NULL);
  k_tr_rsample<<<grid_for(M * E), 256, 0, st>>>(mean, log_std, eps, M, E, z);
  count_launch();
  return launch_status();
}

int32_t tb_tr_rsample_bwd(const float* dz, const float* eps, const float* log_std, int64_t M, int32_t E, float* dlog_std, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(dz && eps && log_std && dlog_std, TB_ERR_NULL);
  k_tr_rsample_bwd<<<1, 64, 0, st>>>(dz, eps, log_std, M, E, dlog_std);
  count_launch();
  return launch_status();
}

// kl_sum [1] zero-initialised by the caller; dlq / dlp are accumulated into
int32_t tb_tr_kl(const float* mq, const float* lq, const float* mp, const float* lp, const uint8_t* valid, float free_nats, const float* scale,
             int64_t M, int32_t E, float* kl_sum, float* dmq, float* dmp, float* dlq, float* dlp, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(mq && lq && mp && lp && valid && scale && kl_sum && dmq && dmp && dlq && dlp, TB_ERR_NULL);
  k_tr_kl<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(mq, lq, mp, lp, valid, free_nats, scale, M, E, kl_sum, dmq, dmp, dlq, dlp);
  count_launch();
  return launch_status();
}

int32_t tb_tr_pose_pe(const float* xy, const float* yaw, const float* f_xy, int32_t n_xy, const float* f_yaw, int32_t n_yaw, int64_t M,
                  float* pe, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(xy && yaw && f_xy && f_yaw && pe, TB_ERR_NULL);
  k_tr_pose_pe<<<grid_for(M * (2 * n_xy + n_yaw)), 256, 0, st>>>(xy, yaw, f_xy, n_xy, f_yaw, n_yaw, M, pe);
  count_launch();
  return launch_status();
}

int32_t tb_tr_dir_to_yaw(const float* d, int64_t M, float* yaw, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(d && yaw, TB_ERR_NULL);
  k_tr_dir_to_yaw<<<grid_for(M), 256, 0, st>>>(d, M, yaw);
  count_launch();
  return launch_status();
}

int32_t tb_tr_dynamics(const float* state, const float* mean, const uint8_t* a_type, const uint8_t* valid, int64_t M, float* pred,
                   const float* dpred, float* dstate, float* dmean, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(state && mean && a_type && valid, TB_ERR_NULL);
  TR_CHECK((dpred == nullptr) ? (pred != nullptr) : (dstate && dmean), TB_ERR_NULL);
  k_tr_dynamics<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(state, mean, a_type, valid, M, pred, dpred, dstate, dmean);
  count_launch();
  return launch_status();
}

int32_t tb_tr_reward(const float* pred, const float* gt, const uint8_t* rv, int64_t M, float* r, const float* dr, float* dpred,
                 void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(pred && gt && rv, TB_ERR_NULL);
  TR_CHECK((dr == nullptr) ? (r != nullptr) : (dpred != nullptr), TB_ERR_NULL);
  k_tr_reward<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(pred, gt, rv, M, r, dr, dpred);
  count_launch();
  return launch_status();
}

int32_t tb_tr_sim_flags(const float* state, const uint8_t* valid, const uint8_t* gt_valid, const float* boundary, const float* dest_pos,
                    const float* dest_dir, const uint8_t* dest_valid, const uint8_t* dest_lane, const uint8_t* dest_edge,
                    const uint8_t* killed, const uint8_t* dest_reached, const uint8_t* goal_valid, int32_t B, int32_t A, uint8_t* o_valid,
                    uint8_t* o_killed, uint8_t* o_dest, uint8_t* o_goal, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(state && valid && boundary && dest_pos && dest_dir && dest_valid && dest_lane && dest_edge && killed && dest_reached &&
               goal_valid && o_valid && o_killed && o_dest && o_goal,
           TB_ERR_NULL);
  k_tr_sim_flags<<<(unsigned)(((long)B * A + 127) / 128), 128, 0, st>>>(state, valid, gt_valid, boundary, dest_pos, dest_dir, dest_valid,
                                                                        dest_lane, dest_edge, killed, dest_reached, goal_valid, B, A,
                                                                        o_valid, o_killed, o_dest, o_goal);
  count_launch();
  return launch_status();
}

// out [1] zero-initialised by the caller
int32_t tb_tr_masked_sum(const float* x, const uint8_t* mask, int64_t n, float* out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(x && mask && out, TB_ERR_NULL);
  k_tr_masked_sum<<<grid_for(n), 256, 0, st>>>(x, mask, n, out);
  count_launch();
  return launch_status();
}

int32_t tb_tr_mask_scale(const uint8_t* mask, const float* scale, int64_t n, float* out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(mask && scale && out, TB_ERR_NULL);
  k_tr_mask_scale<<<grid_for(n), 256, 0, st>>>(mask, scale, n, out);
  count_launch();
  return launch_status();
}

int32_t tb_tr_scale(float* x, int64_t n, float alpha, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(x, TB_ERR_NULL);
  k_tr_scale<<<grid_for(n), 256, 0, st>>>(x, n, alpha);
  count_launch();
  return launch_status();
}

// out [1] zero-initialised by the caller
int32_t tb_tr_sq_norm(const float* g, int64_t n, float* out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(g && out, TB_ERR_NULL);
  k_tr_sq_norm<<<grid_for(n), 256, 0, st>>>(g, n, out);
  count_launch();
  return launch_status();
}

int32_t tb_tr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_by_group, const int32_t* group_end,
                    int32_t n_group, float beta1, float beta2, float eps, int32_t step, const float* sq_norm, float max_norm,
                    void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TR_CHECK(p && g && m && v && lr_by_group && group_end, TB_ERR_NULL);
  TR_CHECK(n > 0 && n_group > 0 && step > 0, TB_ERR_BAD_SHAPE);
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  k_tr_adam<<<grid_for(n), 256, 0, st>>>(p, g, m, v, n, lr_by_group, group_end, n_group, beta1, beta2, eps, bc1, bc2_sqrt, sq_norm,
                                         max_norm);
  count_launch();
  return launch_status();
}

}  // extern "C"
