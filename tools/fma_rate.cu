// Microbenchmark: issue rate of the FP32 FMA forms the depthwise kernel could use, per SM sub-partition, with the kernel's own
// register pattern (8 voxels x 3 taps per filter row, the tap weight shared by 8 consecutive FMAs):
//   ffma2_reuse   fma.rn.f32x2, three 64-bit register operands, weight operand repeated (what dwconv_s1_roll_kernel issues)
//   ffma2_norept  fma.rn.f32x2, all three operands change every instruction
//   ffma_reuse    scalar FFMA, three register operands, weight repeated
//   ffma_const    scalar FFMA, weight from the kernel-parameter constant bank (warp-uniform operand)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu ; ./fma_rate
// Output: cycles per warp instruction per sub-partition and FMA / clk / SM at 4, 2 and 1 warps per sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
struct W { float w[32]; };

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(const float* __restrict__ in, float* out, int reps, long long* cyc, const W cw) {
  const int tid = threadIdx.x;
  long long t0 = 0, t1 = 0;
  if (MODE == 0 || MODE == 1) {
    u64 x[10], w[3], y[8], acc[8];
    for (int j = 0; j < 10; ++j) x[j] = reinterpret_cast<const u64*>(in)[tid * 10 + j];
    for (int j = 0; j < 3; ++j) w[j] = reinterpret_cast<const u64*>(in)[8192 + tid * 3 + j];
    for (int j = 0; j < 8; ++j) { acc[j] = 0ull; y[j] = reinterpret_cast<const u64*>(in)[16384 + tid * 8 + j]; }
    __syncthreads();
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (MODE == 0) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j]) : "l"(x[j + kw]), "l"(w[kw]));
            else asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j]) : "l"(x[j + kw]), "l"(y[(j + kw) & 7]));
          }
    }
    t1 = clock64();
    u64 s = 0;
    for (int j = 0; j < 8; ++j) s ^= acc[j];
    if (s == 0x1234567ull) out[tid] = 1.f;
  } else {
    float x[20], w[6], acc[16];
    for (int j = 0; j < 20; ++j) x[j] = in[tid * 20 + j];
    for (int j = 0; j < 6; ++j) w[j] = in[16384 + tid * 6 + j];
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    __syncthreads();
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (MODE == 2) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[2 * j + e]) : "f"(x[2 * (j + kw) + e]), "f"(w[2 * kw + e]));
              else acc[2 * j + e] = fmaf(x[2 * (j + kw) + e], cw.w[(u * 3 + kw) * 2 + e], acc[2 * j + e]);
            }
    }
    t1 = clock64();
    float s = 0.f;
    for (int j = 0; j < 16; ++j) s += acc[j];
    if (s == 12345.678f) out[tid] = s;
  }
  if (tid == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE> void run(const char* name, const float* in, float* out, long long* cyc, int fma_per_inst, int inst_per_rep) {
  W cw;
  for (int i = 0; i < 32; ++i) cw.w[i] = 1.0f + 0.001f * i;
  const int reps = 4000;
  for (int threads = 512; threads >= 128; threads /= 2) {
    k<MODE><<<148, threads>>>(in, out, 10, cyc, cw);
    cudaDeviceSynchronize();
    k<MODE><<<148, threads>>>(in, out, reps, cyc, cw);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const int wps = threads / 128;                                    // warps per sub-partition
    const double inst = (double)reps * inst_per_rep * wps;            // warp instructions per sub-partition
    printf("%-14s %d warps/SMSP: %5.2f clk per warp instruction per SMSP, %6.1f FMA/clk/SM (%s)\n", name, wps, (double)h / inst,
           inst * 4 * 32 * fma_per_inst / (double)h, cudaGetErrorString(e));
  }
}
int main() {
  float* in; float* out; long long* cyc;
  cudaMalloc(&in, 1 << 20); cudaMemset(in, 0, 1 << 20); cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  run<0>("ffma2_reuse", in, out, cyc, 2, 96);
  run<1>("ffma2_norept", in, out, cyc, 2, 96);
  run<2>("ffma_reuse", in, out, cyc, 1, 192);
  run<3>("ffma_const", in, out, cyc, 1, 192);
  return 0;
}
