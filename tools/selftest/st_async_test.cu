#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__global__ void __cluster_dims__(2,1,1) k(float* out) {
  __shared__ float4 buf[128];
  __shared__ float2 b2[128];
  __shared__ uint64_t bar;
  uint32_t rank; asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\nbarrier.cluster.wait.aligned;" ::: "memory");
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(128 * 24) : "memory");
  uint32_t dst = mapa(smem_u32(&buf[threadIdx.x]), rank ^ 1), dst2 = mapa(smem_u32(&b2[threadIdx.x]), rank ^ 1), rb = mapa(smem_u32(&bar), rank ^ 1);
  float v = threadIdx.x + 1000.f * rank;
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst), "f"(v), "f"(v + 1), "f"(v + 2), "f"(v + 3), "r"(rb) : "memory");
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(dst2), "f"(v), "f"(-v), "r"(rb) : "memory");
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  out[(rank * 128 + threadIdx.x) * 2] = buf[threadIdx.x].x + buf[threadIdx.x].w;
  out[(rank * 128 + threadIdx.x) * 2 + 1] = b2[threadIdx.x].y;
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\nbarrier.cluster.wait.aligned;" ::: "memory");
}
int main() { float* d; cudaMalloc(&d, 4096); k<<<2, 128>>>(d); float h[512]; cudaMemcpy(h, d, 2048, cudaMemcpyDeviceToHost); printf("%s | %f %f %f %f\n", cudaGetErrorString(cudaGetLastError()), h[0], h[1], h[256], h[257]); return 0; }
