// Microbenchmark: cost of one tcgen05.mma.cta_group::1.kind::f16 (M=128, K=16) as a function of N, A from
// shared memory (SS) or tensor memory (TS).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0; d |= (uint64_t)((saddr & 0x3FFFFu) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
__host__ __device__ constexpr uint32_t idesc(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
template <int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) k(int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot; __shared__ __align__(8) uint64_t bar;
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  if (threadIdx.x < 32) {
    uint32_t pred; asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    long long t0 = clock64();
    if (pred) {
      const uint64_t ad = make_desc(sb), bd = make_desc(sb + 16384);
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t d = tm + (uint32_t)((j % NACC) * 256 / NACC * (NACC > 1 ? 1 : 0));
          if (TS) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(tm + 256 + 8 * (j & 3)), "l"(bd + 2 * (j & 3)), "r"(idesc(128, N)), "r"(1u) : "memory");
          else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(ad + 2 * (j & 3)), "l"(bd + 2 * (j & 3)), "r"(idesc(128, N)), "r"(1u) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
    }
    __syncwarp();
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (threadIdx.x < 32) { asm volatile("tcgen05.fence::after_thread_sync;"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u)); }
}
template <int N, bool TS, int NACC> void run(const char* name) {
  long long* d; cudaMalloc(&d, 8); const int reps = 2000;
  cudaFuncSetAttribute(k<N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  k<N, TS, NACC><<<148, 128, 64 * 1024>>>(10, d); cudaDeviceSynchronize();
  k<N, TS, NACC><<<148, 128, 64 * 1024>>>(reps, d); cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-28s N=%3d  %6.1f cycles/MMA  (%s)\n", name, N, (double)h / (reps * 8.0), cudaGetErrorString(e)); cudaFree(d);
}
int main() {
  run<32, false, 1>("SS one accumulator"); run<64, false, 1>("SS one accumulator"); run<128, false, 1>("SS one accumulator"); run<256, false, 1>("SS one accumulator");
  run<32, false, 2>("SS two accumulators"); run<64, false, 2>("SS two accumulators");
  run<32, true, 1>("TS one accumulator"); run<64, true, 1>("TS one accumulator"); run<128, true, 1>("TS one accumulator");
  run<32, true, 2>("TS two accumulators");
  return 0;
}
