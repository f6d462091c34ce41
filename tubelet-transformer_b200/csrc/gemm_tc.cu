// tcgen05 "bf16x3" GEMM for sm_100a: the pointwise (1x1x1) convolutions of the CSN backbone, the
// 2048->256 projections and the decode-pool linear layers.
//
//   C[M,N] = act( scale[n] * sum_k A[m,k] W[n,k] + shift[n] + res[m % res_mod, n] )
//
// Operands are split-bf16 (common.cuh): A = A_hi + A_mid, W = W_hi + W_mid, and the kernel issues
// three tensor-core passes per k-block  A_hi*W_hi + A_hi*W_mid + A_mid*W_hi  into one fp32 TMEM
// accumulator (dropped terms are O(2^-16) relative).  Structure (one CTA per SM, persistent over
// output tiles, 192 threads):
//   warp 0      TMA producer: one cp.async.bulk.tensor.3d per operand and k-block brings the hi and
//               mid planes of a 128 x 64 (A) / BN x 64 (W) tile into 128B-swizzled shared memory.
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.kind::f16 (M=128, N=BN, K=16),
//               12 per k-block; tcgen05.commit releases the smem stage / publishes the accumulator.
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 32 columns per warp), folded-BN scale/shift or
//               bias, residual add, ReLU, fp32 or split-bf16 store.  The accumulator is double
//               buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps tile i+1's MMAs.
#include <cuda.h>
#include <stdio.h>
#include <string.h>

#include "kernels.h"

namespace tc {

constexpr int BM = 128, BK = 64;            // BK bf16 = 128 bytes = one swizzle-128B row
constexpr int NUM_THREADS = 192;
constexpr int A_PLANE_BYTES = BM * BK * 2;  // 16 KB

template <int BN> struct Cfg {
  static constexpr int W_PLANE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_PLANE_BYTES + 2 * W_PLANE_BYTES;
  static constexpr int STAGES = (BN == 64) ? 4 : 3;
  static constexpr int TMEM_COLS = 2 * BN;                      // power of two >= 32
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 4 * BN * 4 /*scale,shift x2*/;
};

struct Params {
  const float* scale; const float* shift;
  const void* res; int res_fmt; int ldr; int res_mod;
  float* Cf; int ldcf;          // fp32 output (nullable)
  void* Cs; int ldcs;           // split output (nullable)
  int M, N, K, act;
};

// ---- PTX wrappers -------------------------------------------------------------------------
TB_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

TB_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
TB_DEVINL void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
TB_DEVINL void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
TB_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
TB_DEVINL void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
TB_DEVINL void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
TB_DEVINL void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
TB_DEVINL void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32
TB_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
TB_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128B swizzle (cute::UMMA::SmemDescriptor):
// start address >> 4 in [0,14), LBO (unused for swizzled K-major, = 1) in [16,30), SBO = 1024 B
// (8 rows x 128 B) >> 4 in [32,46), version 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64).
TB_DEVINL uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b format
// BF16 (1) at [7,10)/[10,13), K-major A and B (0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, Params p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // swizzle-128B needs 1024 B alignment
  const uint32_t bar_base = smem_base + C::STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  // per-accumulator-stage copies of this tile's scale / shift columns: [2][2][BN] floats
  float* s_ss = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BM - 1) / BM, n_tiles = p.N / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int kblocks = p.K / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
          tma_load_3d(sa, &tmA, full_bar(stage), kb * BK, m_blk * BM, 0);
          tma_load_3d(sa + 2 * A_PLANE_BYTES, &tmW, full_bar(stage), kb * BK, n_blk * BN, 0);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1);          // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint64_t a_hi = make_smem_desc(sa), a_mid = make_smem_desc(sa + A_PLANE_BYTES);
          const uint64_t w_hi = make_smem_desc(sa + 2 * A_PLANE_BYTES);
          const uint64_t w_mid = make_smem_desc(sa + 2 * A_PLANE_BYTES + C::W_PLANE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {            // +32 bytes per K=16 step: +2 in the >>4 address field
            umma_bf16(tmem_d, a_mid + 2 * k, w_hi + 2 * k, idesc, (kb | k) != 0);
          }
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d, a_hi + 2 * k, w_mid + 2 * k, idesc, 1);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d, a_hi + 2 * k, w_hi + 2 * k, idesc, 1);
          umma_commit(empty_bar(stage));                 // smem stage reusable once these MMAs retire
          if (kb == kblocks - 1) umma_commit(tfull_bar(as));
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ================= epilogue (warps 2..5) =================
    const int lg = warp & 3;                              // TMEM lane group this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      float* s_scale = s_ss + as * 2 * BN;
      float* s_shift = s_scale + BN;
      {
        const int t = threadIdx.x - 64;                   // 0..127 over the four epilogue warps
        if (t < BN) {
          s_scale[t] = p.scale ? __ldg(p.scale + n_blk * BN + t) : 1.f;
          s_shift[t] = p.shift ? __ldg(p.shift + n_blk * BN + t) : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");    // epilogue warps only
      }
      mbar_wait(tfull_bar(as), aphase);
      tcgen05_fence_after();
      const long long row = (long long)m_blk * BM + lg * 32 + lane;
      const bool row_ok = row < p.M;
      const long long rrow = p.res_mod > 0 ? row % p.res_mod : row;
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN + cc * 32), r);
        const int n0 = n_blk * BN + cc * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 sc = *reinterpret_cast<const float4*>(s_scale + cc * 32 + 4 * j);
          const float4 sh = *reinterpret_cast<const float4*>(s_shift + cc * 32 + 4 * j);
          v[4 * j] = fmaf(__uint_as_float(r[4 * j]), sc.x, sh.x);
          v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), sc.y, sh.y);
          v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), sc.z, sh.z);
          v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), sc.w, sh.w);
        }
        if (row_ok) {
          if (p.res) {
            if (p.res_fmt == FMT_F32) {
              const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res) + rrow * p.ldr + n0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 t = __ldg(rp + j);
                v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
              }
            } else {
              const __nv_bfloat16* hp = split_hi(p.res, rrow, p.ldr) + n0;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 t = load_split4(hp + 4 * j, hp + p.ldr + 4 * j);
                v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
              }
            }
          }
          if (p.act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.Cf) {
            float4* op = reinterpret_cast<float4*>(p.Cf + row * p.ldcf + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (p.Cs) {
            __nv_bfloat16* hp = split_hi(p.Cs, row, p.ldcs) + n0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              store_split4(hp + 4 * j, hp + p.ldcs + 4 * j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static char g_err[256] = "";
static EncodeTiledFn g_encode = nullptr;
static int g_num_sms = 0;

static cudaError_t init_once() {
  if (g_encode) return cudaSuccess;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled entry point not available");
    return e != cudaSuccess ? e : cudaErrorNotSupported;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaFuncSetAttribute(gemm_bf16x3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_bf16x3_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return cudaSuccess;
}

// 3-D map over a split tensor: dims {K, rows, plane}, box {64, box_rows, 2}
static bool encode_split_map(CUtensorMap* map, const void* base, uint64_t k, uint64_t rows, uint64_t row_stride_bytes,
                             uint64_t plane_stride_bytes, uint32_t box_rows) {
  cuuint64_t dims[3] = {k, rows, 2};
  cuuint64_t strides[2] = {row_stride_bytes, plane_stride_bytes};
  cuuint32_t box[3] = {(cuuint32_t)BK, box_rows, 2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled failed (%d): k=%llu rows=%llu rs=%llu ps=%llu", (int)r,
             (unsigned long long)k, (unsigned long long)rows, (unsigned long long)row_stride_bytes,
             (unsigned long long)plane_stride_bytes);
    return false;
  }
  return true;
}

}  // namespace tc

const char* gemm_tc_last_error() { return tc::g_err; }

cudaError_t launch_gemm_tc(const GemmArgs& a, cudaStream_t st) {
  using namespace tc;
  cudaError_t e = init_once();
  if (e != cudaSuccess) return e;
  if (a.M <= 0 || a.N % 64 != 0 || a.K % 64 != 0 || a.lda % 8 != 0 || a.K > a.lda || a.a_fmt != FMT_SPLIT ||
      a.A2 != nullptr || a.Wp == nullptr || a.act == ACT_SIGMOID || (reinterpret_cast<uintptr_t>(a.A) & 15) ||
      (reinterpret_cast<uintptr_t>(a.Wp) & 15)) {
    snprintf(g_err, sizeof g_err, "gemm_tc: unsupported problem M=%d N=%d K=%d lda=%d a_fmt=%d", a.M, a.N, a.K,
             a.lda, a.a_fmt);
    return cudaErrorInvalidValue;
  }
  const int bn = (a.N % 128 == 0) ? 128 : 64;
  CUtensorMap tmA, tmW;
  if (!encode_split_map(&tmA, a.A, a.K, a.M, (uint64_t)a.lda * 4, (uint64_t)a.lda * 2, BM)) return cudaErrorInvalidValue;
  if (!encode_split_map(&tmW, a.Wp, a.K, a.N, (uint64_t)a.K * 2, (uint64_t)a.N * a.K * 2, bn)) return cudaErrorInvalidValue;
  Params p;
  p.scale = a.scale; p.shift = a.shift;
  p.res = a.res; p.res_fmt = a.res_fmt; p.ldr = a.ldr; p.res_mod = a.res_mod;
  if (a.c_fmt == FMT_F32) {
    p.Cf = reinterpret_cast<float*>(a.C); p.ldcf = a.ldc; p.Cs = a.C2; p.ldcs = a.ldc2;
  } else {
    p.Cs = a.C; p.ldcs = a.ldc; p.Cf = reinterpret_cast<float*>(a.C2); p.ldcf = a.ldc2;
  }
  p.M = a.M; p.N = a.N; p.K = a.K; p.act = a.act;
  const int tiles = ceil_div(a.M, BM) * (a.N / bn);
  const int grid = tiles < g_num_sms ? tiles : g_num_sms;
  if (bn == 128)
    gemm_bf16x3_kernel<128><<<grid, NUM_THREADS, Cfg<128>::SMEM_BYTES, st>>>(tmA, tmW, p);
  else
    gemm_bf16x3_kernel<64><<<grid, NUM_THREADS, Cfg<64>::SMEM_BYTES, st>>>(tmA, tmW, p);
  return cudaGetLastError();
}

// fp32 [N,K] -> bf16 [2][N][K] (hi plane, mid plane)
__global__ void pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  __nv_bfloat16 hi, mid;
  split_bf16(w[i], hi, mid);
  out[i] = hi;
  out[n + i] = mid;
}

cudaError_t launch_pack_weight(const float* w, void* out, int N, int K, cudaStream_t st) {
  long long n = (long long)N * K;
  pack_weight_kernel<<<ceil_div(n, 256), 256, 0, st>>>(w, reinterpret_cast<__nv_bfloat16*>(out), n);
  return cudaGetLastError();
}
