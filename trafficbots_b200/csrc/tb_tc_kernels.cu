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
