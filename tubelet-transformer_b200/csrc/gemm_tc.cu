// tcgen05 "bf16x3" GEMM for sm_100a: the pointwise (1x1x1) convolutions of the CSN backbone, the
// 2048->256 projections, the decode-pool and transformer linear layers.
//
//   C[M,N] = act( scale[n] * sum_k A[m,k] W[n,k] + shift[n] + res[m, n] )
//
// Operands are split-bf16 (common.cuh): A = A_hi + A_mid, W = W_hi + W_mid, and the kernel issues
// three tensor-core passes per k-block  A_mid*W_hi + A_hi*W_mid + A_hi*W_hi  into one fp32 TMEM
// accumulator (dropped terms are O(2^-16) relative).  All global traffic is TMA:
//
//   warp 0      operand producer: one cp.async.bulk.tensor.3d per operand and k-block brings the hi and
//               mid planes of a 128 x 64 (A) / BN x 64 (W) tile into 128B-swizzled shared memory.
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.kind::f16 (M=128, N=BN, K=16), 12 per
//               k-block; tcgen05.commit frees the operand stage / publishes the accumulator (TMEM,
//               double buffered: 2 x BN columns, so the epilogue of tile i overlaps the MMAs of i+1).
//   warp 2      panel producer: for every 128 x 64 output panel it claims a 32 KB shared-memory panel
//               buffer and, when there is a residual, TMA-loads the residual panel into it.
//   warps 3..6  epilogue: tcgen05.ld (thread = accumulator row), folded-BN scale/shift or bias, residual
//               from the panel, ReLU, written back IN PLACE into the panel (fp32 or split-bf16, in the
//               128B-swizzled layout the tensor map expects), then one thread issues the TMA store.
//
// A panel is 128 rows x 64 columns = 32 KB in either format, as two 16 KB sub-tiles of 128-byte rows:
//   fp32  : sub-tile s = columns [32s, 32s+32) of the panel  (tensor map: {32, rows, N/32} view of the matrix)
//   split : sub-tile 0 = hi plane, sub-tile 1 = mid plane   (tensor map: {N, rows, 2 planes})
// so thread r owns exactly row r of both sub-tiles in both formats and the in-place update is race free.
// The TMA residual path needs res format == output format; other cases (row-periodic residuals, mixed
// formats) use per-thread global loads (small token-sized GEMMs only).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kernels.h"

namespace tc {

constexpr int BM = 128, BK = 64;            // BK bf16 = 128 bytes = one swizzle-128B row
constexpr int NUM_THREADS = 224;            // 7 warps
constexpr int EPI_WARP0 = 3;                // warps 3..6 (warp % 4 = 3,0,1,2: all four TMEM lane groups)
constexpr int FUSED2_EPI_GROUPS = 2;        // gemm_fused2_kernel: epilogue warp quartets
constexpr int FUSED2_THREADS = 96 + 128 * FUSED2_EPI_GROUPS;
constexpr int A_PLANE_BYTES = BM * BK * 2;  // 16 KB
constexpr int PANEL_BYTES = 32768, SUB_BYTES = 16384;

enum ResMode { RES_NONE = 0, RES_TMA = 1, RES_DIRECT = 2 };

// EPI_GROUPS: independent epilogue warp quartets (each reaches all four TMEM lane groups) that take the tile's 64-column panels in
// turn.  With short K the epilogue -- one dependent chain tcgen05.ld -> scale/shift -> residual (shared memory) -> split -> st.shared
// per thread -- is the kernel: four warps (one per scheduler) cannot hide its latencies, eight can.
template <int BN_, int STAGES_, int PANELS_, int EPI_GROUPS_ = 1> struct Cfg {
  static constexpr int BN = BN_, STAGES = STAGES_, PANELS = PANELS_, EPI_GROUPS = EPI_GROUPS_;
  static constexpr int THREADS = 96 + 128 * EPI_GROUPS;
  static constexpr int W_PLANE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_PLANE_BYTES + 2 * W_PLANE_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;                      // power of two >= 32
  static constexpr int PANELS_PER_TILE = BN / 64;
  static constexpr int BAR_BYTES = 256;
  // scale / shift copies: per accumulator stage with one epilogue group (BN = 256: one copy, the panel barrier orders reuse, to
  // fit 227 KB); with several groups each group keeps its own single copy (it restages at the top of every tile it works on)
  static constexpr int SS_STAGES = EPI_GROUPS > 1 ? EPI_GROUPS : (BN > 128 ? 1 : 2);
  static constexpr int SS_BYTES = SS_STAGES * 2 * BN * 4;
  static constexpr int TEMPTY_COUNT = 4 * (EPI_GROUPS < PANELS_PER_TILE ? EPI_GROUPS : PANELS_PER_TILE);
  static_assert(EPI_GROUPS == 1 || PANELS >= EPI_GROUPS, "one panel buffer per epilogue group at least");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + PANELS * PANEL_BYTES + BAR_BYTES + SS_BYTES;
  static_assert(SMEM_BYTES <= 232448, "over the 227 KB shared-memory limit");
  static_assert(2 * STAGES + 4 + 2 * PANELS + 1 <= BAR_BYTES / 8, "barrier area too small");
};

struct Params {
  const float* scale; const float* shift;
  const void* res; int res_fmt; int ldr; int res_mod;   // RES_DIRECT only
  int res_mode;
  int out_fmt;
  int M, N, K, act;
  int kb1;                                              // k-blocks taken from the first A operand (the rest from the second)
  int group_rows;                                       // > 0: grouped GEMM (GemmArgs::group_rows), a multiple of BM
  int ksplit, part_rows;                                // split-K: work item = (tile, part); part s -> rows + s*part_rows of C
  int a2_wo;                                            // > 0: A2 is a strided 5-D map; output rows per (b,t) frame = a2_wo * a2_ho
  int a2_ho, a2_st, a2_ss;
  int tail_panels;                                      // gemm2 kernel, one tile per pair, no residual: the finished operand stages serve as panel buffers
};

// ---- PTX wrappers -------------------------------------------------------------------------
TB_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one lane of the (converged) warp; keeps the surrounding control flow warp-uniform so that descriptors and
// addresses stay in uniform registers instead of being moved there (R2UR) before every tcgen05.mma
TB_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}

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
TB_DEVINL void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
// second A operand: plain rows, or the strided voxel gather of a striding bottleneck's shortcut (128 consecutive output voxels
// = a box {Wo, bh, bbt} of the output grid = the same box with element strides in the input tensor)
struct Params;
TB_DEVINL void load_a2(uint32_t dst, const CUtensorMap* map, uint32_t bar, int k0, int row0, int a2_wo, int a2_ho, int a2_st, int a2_ss) {
  if (a2_wo <= 0) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(k0), "r"(row0), "r"(0) : "memory");
  } else {
    const int per_bt = a2_wo * a2_ho;
    const int bt = row0 / per_bt, ho = (row0 - bt * per_bt) / a2_wo;
    tma_load_5d(dst, map, bar, k0, 0, ho * a2_ss, bt * a2_st, 0);
  }
}
TB_DEVINL void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
TB_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> TB_DEVINL void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
TB_DEVINL void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
TB_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
TB_DEVINL uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
TB_DEVINL void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
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

// 16-byte chunk j of row r in a 128B-swizzled sub-tile
TB_DEVINL uint32_t swz(uint32_t sub_base, int r, int j) { return sub_base + (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4); }

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                   const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmC,
                   const __grid_constant__ CUtensorMap tmR, Params p) {
  constexpr int BN = C::BN;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);            // swizzle-128B tiles need 1024 B alignment
  const uint32_t panel_base = smem_base + C::STAGES * C::STAGE_BYTES;
  const uint32_t bar_base = panel_base + C::PANELS * PANEL_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 2 + s); };
  auto pfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 4 + s); };
  auto pfree_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 4 + C::PANELS + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4 + 2 * C::PANELS);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_base));
  // per-accumulator-stage copies of this tile's scale / shift columns: [2][2][BN] floats
  float* s_ss = reinterpret_cast<float*>(smem_raw + (bar_base + C::BAR_BYTES - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BM - 1) / BM, n_tiles = p.N / BN;
  const int ksplit = p.ksplit > 1 ? p.ksplit : 1;
  const int num_tiles = m_tiles * n_tiles * ksplit;         // work items: (output tile, K part); "tile" below = item index
  const int kblocks = p.K / BK / ksplit;                    // k-blocks per item
  const int kb_total = p.K / BK;

  if (warp == 0 && lane == 0) {
    if (smem_base & 1023u) __trap();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    if (p.kb1 < kb_total) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    if (p.res_mode == RES_TMA) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), C::TEMPTY_COUNT);
    }
    for (int s = 0; s < C::PANELS; ++s) {
      mbar_init(pfull_bar(s), 1);
      mbar_init(pfree_bar(s), 1);
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
  pdl_trigger();
  pdl_wait();                                               // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ================= operand producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int part = tile % ksplit, ot = tile / ksplit;
        const int m_blk = ot / n_tiles, n_blk = ot % n_tiles;
        for (int kb = part * kblocks; kb < (part + 1) * kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
          if (kb < p.kb1) tma_load_3d(sa, &tmA, full_bar(stage), kb * BK, m_blk * BM, 0);
          else load_a2(sa, &tmA2, full_bar(stage), (kb - p.kb1) * BK, m_blk * BM, p.a2_wo, p.a2_ho, p.a2_st, p.a2_ss);
          const int wrow = p.group_rows > 0 ? (m_blk * BM / p.group_rows) * p.N : 0;   // grouped: this row block's W rows
          tma_load_3d(sa + 2 * A_PLANE_BYTES, &tmW, full_bar(stage), kb * BK, wrow + n_blk * BN, 0);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp runs the loop, one elected lane issues) =================
    constexpr uint32_t idesc = make_idesc(BM, BN);
    const uint64_t desc0 = make_smem_desc(smem_base);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(tempty_bar(as), aphase ^ 1);            // epilogue has drained this accumulator
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
#pragma unroll 1
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tcgen05_fence_after();
        if (elect_one()) {
          // descriptor address field is in 16-byte units
          const uint64_t a_hi = desc0 + (uint64_t)(stage * (C::STAGE_BYTES >> 4)), a_mid = a_hi + (A_PLANE_BYTES >> 4);
          const uint64_t w_hi = a_hi + ((2 * A_PLANE_BYTES) >> 4), w_mid = w_hi + (C::W_PLANE_BYTES >> 4);
          const uint32_t first = kb == 0 ? 0u : 1u;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)              // +32 bytes per K=16 step: +2 in the >>4 address field
            umma_bf16(tmem_d, a_mid + 2 * k, w_hi + 2 * k, idesc, k == 0 ? first : 1u);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d, a_hi + 2 * k, w_mid + 2 * k, idesc, 1);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d, a_hi + 2 * k, w_hi + 2 * k, idesc, 1);
          umma_commit(empty_bar(stage));                 // smem stage reusable once these MMAs retire
          if (kb == kblocks - 1) umma_commit(tfull_bar(as));
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ================= panel producer =================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int part = tile % ksplit, ot = tile / ksplit;
        const int m_blk = ot / n_tiles, n_blk = ot % n_tiles;
        for (int j = 0; j < C::PANELS_PER_TILE; ++j) {
          mbar_wait(pfree_bar(slot), phase ^ 1);         // the store that last read this buffer has drained
          if (p.res_mode == RES_TMA && part == 0) {
            const int col = n_blk * BN + j * 64;
            mbar_expect_tx(pfull_bar(slot), PANEL_BYTES);
            if (p.out_fmt == FMT_F32) tma_load_3d(panel_base + slot * PANEL_BYTES, &tmR, pfull_bar(slot), 0, m_blk * BM, col >> 5);
            else tma_load_3d(panel_base + slot * PANEL_BYTES, &tmR, pfull_bar(slot), col, m_blk * BM, 0);
          } else {
            mbar_arrive(pfull_bar(slot));
          }
          if (++slot == C::PANELS) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ================= epilogue (warps 3..6 [, 7..10]): group eg takes the panels with (running panel index) % groups == eg ====
    constexpr int G = C::EPI_GROUPS, PPT = C::PANELS_PER_TILE;
    const int lg = warp & 3;                              // TMEM lane group this warp may access
    const int eg = (warp - EPI_WARP0) >> 2;               // epilogue group
    const int et = (threadIdx.x - EPI_WARP0 * 32) & 127;  // 0..127 inside the group
    const int r = lg * 32 + lane;                         // accumulator row = panel row of this thread
    const int gbar = 1 + eg;                              // named barrier of this group
    int it = 0, prev_slot = -1;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      if (G > 1 && PPT < G && (it * PPT) % G != eg) continue;   // one panel per tile: the groups alternate over tiles
      const int part = tile % ksplit, ot = tile / ksplit;
      const int m_blk = ot / n_tiles, n_blk = ot % n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      float* s_scale = s_ss + (G > 1 ? eg : (C::SS_STAGES == 2 ? as : 0)) * 2 * BN;
      float* s_shift = s_scale + BN;
      const int grp = p.group_rows > 0 ? m_blk * BM / p.group_rows : 0;
      const int gcol = grp * p.N;                           // grouped: column / parameter offset of this row block's group
      const int res_mode = part == 0 ? p.res_mode : RES_NONE;   // split-K: shift and residual go into part 0 only
      for (int cidx = et; cidx < BN; cidx += 128) {
        s_scale[cidx] = p.scale ? __ldg(p.scale + gcol + n_blk * BN + cidx) : 1.f;
        s_shift[cidx] = (p.shift && part == 0) ? __ldg(p.shift + gcol + n_blk * BN + cidx) : 0.f;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(gbar) : "memory");      // this epilogue group only
      mbar_wait(tfull_bar(as), aphase);
      tcgen05_fence_after();
      const long long row = (long long)m_blk * BM + r;
      const bool row_ok = row < p.M;
      const long long rrow = p.res_mod > 0 ? row % p.res_mod : row;
#pragma unroll 1
      for (int j = 0; j < PPT; ++j) {
        const int gp = it * PPT + j;                      // running panel index: the producer fills the buffers in this order
        if (G > 1 && gp % G != eg) continue;
        const int slot = gp % C::PANELS;
        const uint32_t pphase = (uint32_t)((gp / C::PANELS) & 1);
        mbar_wait(pfull_bar(slot), pphase);
        const uint32_t pb = panel_base + slot * PANEL_BYTES;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {                     // 32-column halves of the panel
          uint32_t acc[32];
          tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN + j * 64 + h * 32), acc);
          const int cl = j * 64 + h * 32;                 // column inside the tile
          float v[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 sc = *reinterpret_cast<const float4*>(s_scale + cl + 4 * q);
            const float4 sh = *reinterpret_cast<const float4*>(s_shift + cl + 4 * q);
            v[4 * q] = fmaf(__uint_as_float(acc[4 * q]), sc.x, sh.x);
            v[4 * q + 1] = fmaf(__uint_as_float(acc[4 * q + 1]), sc.y, sh.y);
            v[4 * q + 2] = fmaf(__uint_as_float(acc[4 * q + 2]), sc.z, sh.z);
            v[4 * q + 3] = fmaf(__uint_as_float(acc[4 * q + 3]), sc.w, sh.w);
          }
          if (res_mode == RES_TMA) {
            if (p.out_fmt == FMT_F32) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const uint4 t = lds128(swz(pb + h * SUB_BYTES, r, q));
                v[4 * q] += __uint_as_float(t.x); v[4 * q + 1] += __uint_as_float(t.y);
                v[4 * q + 2] += __uint_as_float(t.z); v[4 * q + 3] += __uint_as_float(t.w);
              }
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) {               // 8 columns per 16-byte chunk of each plane
                const uint4 a = lds128(swz(pb, r, 4 * h + q));
                const uint4 b = lds128(swz(pb + SUB_BYTES, r, 4 * h + q));
                v[8 * q] += bf16_lo_to_f32(a.x) + bf16_lo_to_f32(b.x); v[8 * q + 1] += bf16_hi_to_f32(a.x) + bf16_hi_to_f32(b.x);
                v[8 * q + 2] += bf16_lo_to_f32(a.y) + bf16_lo_to_f32(b.y); v[8 * q + 3] += bf16_hi_to_f32(a.y) + bf16_hi_to_f32(b.y);
                v[8 * q + 4] += bf16_lo_to_f32(a.z) + bf16_lo_to_f32(b.z); v[8 * q + 5] += bf16_hi_to_f32(a.z) + bf16_hi_to_f32(b.z);
                v[8 * q + 6] += bf16_lo_to_f32(a.w) + bf16_lo_to_f32(b.w); v[8 * q + 7] += bf16_hi_to_f32(a.w) + bf16_hi_to_f32(b.w);
              }
            }
          } else if (res_mode == RES_DIRECT && row_ok) {
            const int n0 = n_blk * BN + cl;
            if (p.res_fmt == FMT_F32) {
              const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res) + rrow * p.ldr + n0);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 t = __ldg(rp + q);
                v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
              }
            } else {
              const __nv_bfloat16* hp = split_hi(p.res, rrow, p.ldr) + n0;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 t = load_split4(hp + 4 * q, hp + p.ldr + 4 * q);
                v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
              }
            }
          }
          if (p.act == ACT_RELU) {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q], 0.f);
          }
          if (p.out_fmt == FMT_F32) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              sts128(swz(pb + h * SUB_BYTES, r, q), make_uint4(__float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]),
                                                               __float_as_uint(v[4 * q + 2]), __float_as_uint(v[4 * q + 3])));
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t hi[4], mid[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat16 h0, m0, h1, m1;
                split_bf16(v[8 * q + 2 * e], h0, m0);
                split_bf16(v[8 * q + 2 * e + 1], h1, m1);
                hi[e] = pack_bf16x2(h0, h1);
                mid[e] = pack_bf16x2(m0, m1);
              }
              sts128(swz(pb, r, 4 * h + q), make_uint4(hi[0], hi[1], hi[2], hi[3]));
              sts128(swz(pb + SUB_BYTES, r, 4 * h + q), make_uint4(mid[0], mid[1], mid[2], mid[3]));
            }
          }
        }
        if (j + G >= PPT) {                               // this group's last panel of the tile: its part of the accumulator is read
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(as));
        }
        fence_proxy_async();                              // generic-proxy writes -> visible to the TMA store
        asm volatile("bar.sync %0, 128;" ::"r"(gbar) : "memory");
        if (et == 0) {
          const int col = gcol + n_blk * BN + j * 64;
          const int orow = m_blk * BM - grp * p.group_rows + part * p.part_rows;
          if (p.out_fmt == FMT_F32) tma_store_3d(&tmC, pb, 0, orow, col >> 5);
          else tma_store_3d(&tmC, pb, col, orow, 0);
          bulk_commit();
          // keep at most PANELS-2 stores in flight, then recycle the buffer(s) whose store has drained
          if constexpr (C::PANELS == 1 || G > 1) {          // several groups: each drains its own store (the other group computes meanwhile)
            bulk_wait_read<0>();
            mbar_arrive(pfree_bar(slot));
          } else {
            bulk_wait_read<C::PANELS - 2>();
            if constexpr (C::PANELS == 2) {
              mbar_arrive(pfree_bar(slot));
            } else {
              if (prev_slot >= 0) mbar_arrive(pfree_bar(prev_slot));
              prev_slot = slot;
            }
          }
        }
      }
    }
    if (et == 0) bulk_wait_all();                         // shared memory must outlive the last stores
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}

// =============================================================================================================
// Fused pair of pointwise convolutions for the 256-channel stage (layer1 of the CSN backbone, ir_CSN_152.py:70-90):
//
//   x'[m, 0:256] = relu(scale * ([A | Ab] W4^T) + shift + res)        conv4 (+ shortcut) of bottleneck i    -> split, HBM
//   t1'[m, 0:N2] = relu(scale2 * (x' W1'^T) + shift2)                 conv1 of bottleneck i+1 (K2 = 256)    -> fp32,  HBM
//
// The second GEMM's A operand never leaves the SM: a finished 128 x 64 output panel of x' sits in shared memory in
// exactly the layout a K-major SWIZZLE_128B UMMA operand needs (hi plane | mid plane, 128-byte rows), so while the
// TMA store drains it the MMA warp multiplies it with k-block j of W1' into a second accumulator.  This removes the
// separate conv1 launch and its read of x' (1.07 GB per bottleneck at 8 clips) -- the layer is HBM bound.
// One CTA owns whole 128-row blocks (both 128-column halves of x'), W1' streams through the operand ring
// (N2 x 256 x 4 B per row block, from L2).
//   extra barriers: pready[slot]  epilogue -> MMA warp: panel is final and visible to the async proxy
//                   pcons[slot]   tcgen05.commit: the second GEMM has consumed the panel (it may be recycled once
//                                 the TMA store has drained too)
//                   d2full / d2empty  second accumulator handshake
// =============================================================================================================
struct Params2 {
  const float* scale2; const float* shift2;
  int N2;
};

// KBW_MAX caps the k-blocks of W1' per ring stage: with N2 = 64 a full stage holds the W1' rows of all four panels of a 256-column
// block, which pins one of the two ring slots for the whole block and forces the serial `delay` order whenever conv4 has more than one
// k-block (the first bottleneck of layer1, K = 64 + 64); half-filled stages (KBW = 2) are released per sub-tile instead.
template <int N1, int N2, int KBW_MAX = 4>   // N1 = channels of x' (256: layer1, 512: layer2), N2 = planes of the next bottleneck's conv1
__global__ void __launch_bounds__(FUSED2_THREADS, 1)
gemm_fused2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                   const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmC,
                   const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmW2,
                   const __grid_constant__ CUtensorMap tmC2, Params p, Params2 q) {
  constexpr int BN = 128, STAGES = 2, PANELS = 3, STAGE_BYTES = 65536;
  constexpr int NSUB = N1 / BN, P1 = N1 / 64, P2 = N2 / 64;  // x' = NSUB sub-tiles of 128 columns = P1 panels; t1' = P2 panels
  constexpr int W2_CHUNK = N2 * 256;                        // bytes of one k-block of W1' (hi + mid planes)
  constexpr int KBW = STAGE_BYTES / W2_CHUNK < KBW_MAX ? STAGE_BYTES / W2_CHUNK : KBW_MAX;   // k-blocks of the second GEMM per ring stage
  constexpr int D2_COL = 2 * BN;                            // TMEM: [0,256) two conv4 accumulators, [256, 256+N2) the second GEMM
  static_assert(KBW == 1 || KBW == 2 || KBW == 4, "W1' chunking");
  static_assert(D2_COL + N2 <= 512, "tensor memory: two conv4 accumulators + the second GEMM's accumulator");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t panel_base = smem_base + STAGES * STAGE_BYTES;
  const uint32_t bar_base = panel_base + PANELS * PANEL_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (4 + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (6 + s); };
  auto pfull_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto pfree_bar = [&](int s) { return bar_base + 8u * (11 + s); };
  auto pready_bar = [&](int s) { return bar_base + 8u * (14 + s); };
  auto pcons_bar = [&](int s) { return bar_base + 8u * (17 + s); };
  const uint32_t d2full_bar = bar_base + 8u * 20, d2empty_bar = bar_base + 8u * 21;
  const uint32_t tmem_slot = bar_base + 8u * 22;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BM - 1) / BM;
  const int kblocks = p.K / BK;
  // The k-th conv4 accumulator after the first two of a row block, g = sub + 2: sub-tile g of this block, or sub-tile g - NSUB
  // of the CTA's next block.  It is issued right after the panels of sub-tile `sub` have been consumed by the second GEMM --
  // unless a W1' chunk spans both sub-tiles' panels and a conv4 stage needs more than the one free ring slot (KBW = 4 and
  // more than one k-block): then all of them follow the block's last panel.
  const bool delay = KBW == 4 && kblocks > 1;

  if (warp == 0 && lane == 0) {
    if (smem_base & 1023u) __trap();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    if (p.kb1 < kblocks) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC2) : "memory");
    if (p.res_mode == RES_TMA) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4 * FUSED2_EPI_GROUPS); }
    for (int s = 0; s < PANELS; ++s) {
      mbar_init(pfull_bar(s), 1); mbar_init(pfree_bar(s), 1); mbar_init(pready_bar(s), 1); mbar_init(pcons_bar(s), 1);
    }
    mbar_init(d2full_bar, 1);
    mbar_init(d2empty_bar, 4 * (P2 < FUSED2_EPI_GROUPS ? P2 : FUSED2_EPI_GROUPS));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_trigger();
  pdl_wait();

  // Ring order = the MMA warp's consumption order.  S(sub, t) = the k-blocks [A | W4 rows of sub] of row block t, W2(c, t) =
  // chunk c (KBW k-blocks) of W1'.  After the first block's S(0), S(1), per block and sub-tile s:
  //   producer   [W2(c) when panel 2s starts a chunk]  S(s + 2)            MMA warp   G2(2s) G2(2s+1)  G1(s + 2)
  // so the next accumulator is ready before the epilogue has finished the current panels.
  if (warp == 0) {
    // ================= operand producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto load_s = [&](int sub, int m_blk) {
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          if (kb < p.kb1) tma_load_3d(sa, &tmA, full_bar(stage), kb * BK, m_blk * BM, 0);
          else load_a2(sa, &tmA2, full_bar(stage), (kb - p.kb1) * BK, m_blk * BM, p.a2_wo, p.a2_ho, p.a2_st, p.a2_ss);
          tma_load_3d(sa + 2 * A_PLANE_BYTES, &tmW, full_bar(stage), kb * BK, sub * BN, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      };
      auto load_w2 = [&](int c) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        mbar_expect_tx(full_bar(stage), KBW * W2_CHUNK);
        for (int e = 0; e < KBW; ++e) tma_load_3d(sa + e * W2_CHUNK, &tmW2, full_bar(stage), (c * KBW + e) * BK, 0, 0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      };
      if ((int)blockIdx.x < m_tiles) { load_s(0, blockIdx.x); load_s(1, blockIdx.x); }
      for (int m_blk = blockIdx.x; m_blk < m_tiles; m_blk += gridDim.x) {
        const int next = m_blk + gridDim.x;
        const bool has_next = next < m_tiles;
        auto load_g = [&](int g) {
          if (g < NSUB) load_s(g, m_blk);
          else if (has_next) load_s(g - NSUB, next);
        };
        for (int sub = 0; sub < NSUB; ++sub) {
          if ((2 * sub) % KBW == 0) load_w2(2 * sub / KBW);
          if (KBW == 1) load_w2(2 * sub + 1);                // one k-block of W1' per stage (N2 = 256): a chunk per panel
          if (!delay) load_g(sub + 2);
        }
        if (delay)
          for (int sub = 0; sub < NSUB; ++sub) load_g(sub + 2);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc1 = make_idesc(BM, BN), idesc2 = make_idesc(BM, N2);
    const uint64_t desc0 = make_smem_desc(smem_base);
    const uint64_t pdesc0 = make_smem_desc(panel_base);
    int stage = 0, pslot = 0;
    uint32_t phase = 0, ready_ph = 0;                        // ready_ph bit s: parity of the next pready[s] completion
    int acc_uses[2] = {0, 0};                                // how often each conv4 accumulator buffer has been issued
    // conv4 sub-tile `sub` into accumulator buffer sub & 1
    auto g1 = [&](int sub) {
      const int buf = sub & 1;
      mbar_wait(tempty_bar(buf), (uint32_t)((acc_uses[buf] & 1) ^ 1));
      ++acc_uses[buf];
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
#pragma unroll 1
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t a_hi = desc0 + (uint64_t)(stage * (STAGE_BYTES >> 4)), a_mid = a_hi + (A_PLANE_BYTES >> 4);
          const uint64_t w_hi = a_hi + ((2 * A_PLANE_BYTES) >> 4), w_mid = w_hi + ((BN * BK * 2) >> 4);
          const uint32_t first = kb == 0 ? 0u : 1u;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d, a_mid + 2 * k, w_hi + 2 * k, idesc1, k == 0 ? first : 1u);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d, a_hi + 2 * k, w_mid + 2 * k, idesc1, 1);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d, a_hi + 2 * k, w_hi + 2 * k, idesc1, 1);
          umma_commit(empty_bar(stage));
          if (kb == kblocks - 1) umma_commit(tfull_bar(buf));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    };
    int w2_stage = 0;                                       // ring slot of the W1' chunk in use (conv4 stages are consumed in between)
    // k-block j of the second GEMM = output panel j of x', read from its panel buffer
    auto g2 = [&](int j) {
      if (j % KBW == 0) {
        mbar_wait(full_bar(stage), phase);
        w2_stage = stage;
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      mbar_wait(pready_bar(pslot), (ready_ph >> pslot) & 1u);
      ready_ph ^= 1u << pslot;
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t tmem_d2 = tmem_base + (uint32_t)D2_COL;
        const uint64_t a_hi = pdesc0 + (uint64_t)(pslot * (PANEL_BYTES >> 4)), a_mid = a_hi + (SUB_BYTES >> 4);
        const uint64_t w_hi = desc0 + (uint64_t)(w2_stage * (STAGE_BYTES >> 4)) + (uint64_t)((j % KBW) * (W2_CHUNK >> 4));
        const uint64_t w_mid = w_hi + ((N2 * BK * 2) >> 4);
        const uint32_t first = j == 0 ? 0u : 1u;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d2, a_mid + 2 * k, w_hi + 2 * k, idesc2, k == 0 ? first : 1u);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d2, a_hi + 2 * k, w_mid + 2 * k, idesc2, 1);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d2, a_hi + 2 * k, w_hi + 2 * k, idesc2, 1);
        umma_commit(pcons_bar(pslot));
        if (j % KBW == KBW - 1) umma_commit(empty_bar(w2_stage));
        if (j == P1 - 1) umma_commit(d2full_bar);
      }
      __syncwarp();
      if (++pslot == PANELS) pslot = 0;
    };
    int tl = 0;
    if ((int)blockIdx.x < m_tiles) { g1(0); g1(1); }
    for (int m_blk = blockIdx.x; m_blk < m_tiles; m_blk += gridDim.x, ++tl) {
      const bool has_next = m_blk + (int)gridDim.x < m_tiles;
      mbar_wait(d2empty_bar, (uint32_t)((tl & 1) ^ 1));      // the epilogue has drained the previous block's second accumulator
      tcgen05_fence_after();
      for (int sub = 0; sub < NSUB; ++sub) {
        g2(2 * sub); g2(2 * sub + 1);
        const int g = sub + 2;
        if (!delay && (g < NSUB || has_next)) g1(g % NSUB);
      }
      if (delay)
        for (int sub = 0; sub < NSUB; ++sub)
          if (sub + 2 < NSUB || has_next) g1((sub + 2) % NSUB);
      for (int jp = 0; jp < P2; ++jp)                          // the t1' panels use ring slots too (not operands)
        if (++pslot == PANELS) pslot = 0;
    }
  } else if (warp == 2) {
    // ================= panel producer =================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (int m_blk = blockIdx.x; m_blk < m_tiles; m_blk += gridDim.x) {
        for (int j = 0; j < P1 + P2; ++j) {
          mbar_wait(pfree_bar(slot), phase ^ 1);
          if (j < P1 && p.res_mode == RES_TMA) {
            mbar_expect_tx(pfull_bar(slot), PANEL_BYTES);
            tma_load_3d(panel_base + slot * PANEL_BYTES, &tmR, pfull_bar(slot), j * 64, m_blk * BM, 0);
          } else {
            mbar_arrive(pfull_bar(slot));
          }
          if (++slot == PANELS) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ================= epilogue: two warp quartets (warps 3..6 and 7..10); group eg takes panel jj == eg of every 128-column
    // sub-tile of x' and panel jp % 2 == eg of t1'.  One quartet (one warp per scheduler) cannot hide the latencies of its
    // dependent chain tcgen05.ld -> scale/shift -> residual (shared memory) -> split -> st.shared; with two, one group computes
    // while the other drains its store.  Panels travel through the ring in the running order gp (slot = gp % PANELS), which both
    // groups count; the group that owns a panel publishes it (pready), stores it, and recycles its buffer as soon as the store has
    // read it and -- for an x' panel -- the second GEMM has consumed it (a lazy recycle would make the two groups take turns).
    constexpr int G = FUSED2_EPI_GROUPS;
    const int lg = warp & 3;
    const int eg = (warp - EPI_WARP0) >> 2;
    const int et = (threadIdx.x - EPI_WARP0 * 32) & 127;
    const int r = lg * 32 + lane;
    const int gbar = 1 + eg;
    int tl = 0;
    int acc_seen[2] = {0, 0};
    int xuse[PANELS];                                        // x' panels each ring slot has carried so far (parity of its pcons barrier)
#pragma unroll
    for (int i = 0; i < PANELS; ++i) xuse[i] = 0;
    for (int m_blk = blockIdx.x; m_blk < m_tiles; m_blk += gridDim.x, ++tl) {
      const int gp0 = tl * (P1 + P2);
      for (int sub = 0; sub < NSUB; ++sub) {
        const int as = sub & 1;
        mbar_wait(tfull_bar(as), (uint32_t)(acc_seen[as] & 1));
        ++acc_seen[as];
        tcgen05_fence_after();
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int gp = gp0 + 2 * sub + jj, slot = gp % PANELS;
          uint32_t cons_parity = 0;
#pragma unroll
          for (int i = 0; i < PANELS; ++i)
            if (i == slot) { cons_parity = (uint32_t)(xuse[i] & 1); ++xuse[i]; }
          if (jj % G != eg) continue;
          mbar_wait(pfull_bar(slot), (uint32_t)((gp / PANELS) & 1));
          const uint32_t pb = panel_base + slot * PANEL_BYTES;
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            uint32_t acc[32];
            tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN + jj * 64 + h * 32), acc);
            const int c0 = sub * BN + jj * 64 + h * 32;      // column of x'
            float v[32];
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              const float4 sc = p.scale ? __ldg(reinterpret_cast<const float4*>(p.scale + c0) + c4) : make_float4(1.f, 1.f, 1.f, 1.f);
              const float4 sh = p.shift ? __ldg(reinterpret_cast<const float4*>(p.shift + c0) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
              v[4 * c4] = fmaf(__uint_as_float(acc[4 * c4]), sc.x, sh.x);
              v[4 * c4 + 1] = fmaf(__uint_as_float(acc[4 * c4 + 1]), sc.y, sh.y);
              v[4 * c4 + 2] = fmaf(__uint_as_float(acc[4 * c4 + 2]), sc.z, sh.z);
              v[4 * c4 + 3] = fmaf(__uint_as_float(acc[4 * c4 + 3]), sc.w, sh.w);
            }
            if (p.res_mode == RES_TMA) {
#pragma unroll
              for (int c8 = 0; c8 < 4; ++c8) {
                const uint4 a = lds128(swz(pb, r, 4 * h + c8));
                const uint4 b = lds128(swz(pb + SUB_BYTES, r, 4 * h + c8));
                v[8 * c8] += bf16_lo_to_f32(a.x) + bf16_lo_to_f32(b.x); v[8 * c8 + 1] += bf16_hi_to_f32(a.x) + bf16_hi_to_f32(b.x);
                v[8 * c8 + 2] += bf16_lo_to_f32(a.y) + bf16_lo_to_f32(b.y); v[8 * c8 + 3] += bf16_hi_to_f32(a.y) + bf16_hi_to_f32(b.y);
                v[8 * c8 + 4] += bf16_lo_to_f32(a.z) + bf16_lo_to_f32(b.z); v[8 * c8 + 5] += bf16_hi_to_f32(a.z) + bf16_hi_to_f32(b.z);
                v[8 * c8 + 6] += bf16_lo_to_f32(a.w) + bf16_lo_to_f32(b.w); v[8 * c8 + 7] += bf16_hi_to_f32(a.w) + bf16_hi_to_f32(b.w);
              }
            }
            if (p.act == ACT_RELU) {
#pragma unroll
              for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c], 0.f);
            }
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              uint4 hi, mid;
              split_bf16x2(v[8 * c8], v[8 * c8 + 1], hi.x, mid.x);
              split_bf16x2(v[8 * c8 + 2], v[8 * c8 + 3], hi.y, mid.y);
              split_bf16x2(v[8 * c8 + 4], v[8 * c8 + 5], hi.z, mid.z);
              split_bf16x2(v[8 * c8 + 6], v[8 * c8 + 7], hi.w, mid.w);
              sts128(swz(pb, r, 4 * h + c8), hi);
              sts128(swz(pb + SUB_BYTES, r, 4 * h + c8), mid);
            }
          }
          // this group's only panel of the sub-tile: its columns of the accumulator have been read
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(as));
          fence_proxy_async();                               // visible to the TMA store AND to the second GEMM's MMAs
          asm volatile("bar.sync %0, 128;" ::"r"(gbar) : "memory");
          if (et == 0) {
            mbar_arrive(pready_bar(slot));
            tma_store_3d(&tmC, pb, sub * BN + jj * 64, m_blk * BM, 0);
            bulk_commit();
            bulk_wait_read<0>();                             // the store has read the panel ...
            mbar_wait(pcons_bar(slot), cons_parity);         // ... and so has the second GEMM
            mbar_arrive(pfree_bar(slot));
          }
        }
      }
      // ---- second accumulator: t1' = relu(scale2 * D2 + shift2), fp32 panels ----
      bool d2_waited = false;
#pragma unroll 1
      for (int jp = 0; jp < P2; ++jp) {
        const int gp = gp0 + P1 + jp, slot = gp % PANELS;
        if (jp % G != eg) continue;
        if (!d2_waited) {
          mbar_wait(d2full_bar, (uint32_t)(tl & 1));
          tcgen05_fence_after();
          d2_waited = true;
        }
        mbar_wait(pfull_bar(slot), (uint32_t)((gp / PANELS) & 1));
        const uint32_t pb = panel_base + slot * PANEL_BYTES;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          uint32_t acc[32];
          tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(D2_COL + jp * 64 + h * 32), acc);
          const int c0 = jp * 64 + h * 32;
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(q.scale2 + c0) + c4);
            const float4 sh = __ldg(reinterpret_cast<const float4*>(q.shift2 + c0) + c4);
            uint4 o;
            o.x = __float_as_uint(fmaxf(fmaf(__uint_as_float(acc[4 * c4]), sc.x, sh.x), 0.f));
            o.y = __float_as_uint(fmaxf(fmaf(__uint_as_float(acc[4 * c4 + 1]), sc.y, sh.y), 0.f));
            o.z = __float_as_uint(fmaxf(fmaf(__uint_as_float(acc[4 * c4 + 2]), sc.z, sh.z), 0.f));
            o.w = __float_as_uint(fmaxf(fmaf(__uint_as_float(acc[4 * c4 + 3]), sc.w, sh.w), 0.f));
            sts128(swz(pb + h * SUB_BYTES, r, c4), o);
          }
        }
        if (jp + G >= P2) {                                  // this group's last panel of t1'
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d2empty_bar);
        }
        fence_proxy_async();
        asm volatile("bar.sync %0, 128;" ::"r"(gbar) : "memory");
        if (et == 0) {
          tma_store_3d(&tmC2, pb, 0, m_blk * BM, (jp * 64) >> 5);
          bulk_commit();
          bulk_wait_read<0>();
          mbar_arrive(pfree_bar(slot));
        }
      }
    }
    if (et == 0) bulk_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// =============================================================================================================
// CTA-pair variant for the deep-K shapes (K >= 512, N % 256 == 0): tcgen05.mma.cta_group::2, one 256 x 256 output tile
// per pair.  The single-CTA kernel moves 64 KB from L2 into shared memory per k-block and 128 x 128 tile (operands
// are 4 bytes per element for three bf16 passes), i.e. ~85 B/clk/SM at the tensor rate -- twice what L2 delivers per SM, and
// these GEMMs ran at 50-60 % of the MMA rate.  In a pair each CTA loads its own 128 rows of A and its own 128 of the 256
// W rows (64 KB per k-block as before) but the MMA covers 256 x 256: 4x the FLOPs per byte fetched.
//   per CTA: 3 stages x (A 32 KB + W half 32 KB), one 32 KB panel buffer, 2 x 256 TMEM columns (double-buffered accumulator)
//   barriers: full[s]   local TMA completion;  xfull[s] (leader) the peer's relay warp reports ITS stage s full
//             empty[s]  tcgen05.commit multicast to both CTAs;  tfull[a] commit multicast;  tempty[a] (leader) 8 epilogue warps
// =============================================================================================================
TB_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
TB_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
TB_DEVINL uint32_t mapa_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
TB_DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// release at cluster scope: the arriving thread's (and, through the preceding CTA barrier, its group's) shared-memory writes are
// visible to whoever acquires the barrier in the other CTA (compiles to a MEMBAR: only where data is published, not for counters)
TB_DEVINL void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
TB_DEVINL void umma2_commit_mc(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
TB_DEVINL void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

// STAGES x 64 KB operand ring + PANELS x 32 KB epilogue panels: <3, 1> for the deep-K shapes (the next tile's main loop hides the
// one-panel epilogue), <2, 3> for short K with a residual (conv4 of the 1024-channel stage: K = 256, so the epilogue IS the kernel --
// panels pipelined load | compute | store as in gemm_bf16x3_kernel)
// BN2 = 128 (256 x 128 pair tiles, each CTA loads its 128 rows of A and 64 of the tile's 128 W rows: 48 KB per k-block instead of the
// 64 KB of a single-CTA 128 x 128 tile) keeps the tile count of the single-CTA kernel -- for the short-K, wide-N shapes that are bound
// by L2 -> shared-memory throughput and whose 256-column pair tiles quantise badly (conv4 of the 1024-channel stage: 512 tiles on 74 pairs).
template <int STAGES, int PANELS, int G, int BN2 = 256>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(96 + 128 * G, 1)
gemm2_bf16x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                    const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmC,
                    const __grid_constant__ CUtensorMap tmR, Params p) {
  constexpr int BNH = BN2 / 2, STAGE_BYTES = 2 * A_PLANE_BYTES + 2 * BNH * BK * 2, PPT = BN2 / 64;
  static_assert(BN2 == 256 || BN2 == 128, "pair tile width");
  static_assert(STAGE_BYTES % 1024 == 0, "operand stages must keep the 1024-byte alignment of the swizzled tiles");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t panel_base = smem_base + STAGES * STAGE_BYTES;
  const uint32_t bar_base = panel_base + PANELS * PANEL_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto xfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (3 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (3 * STAGES + 2 + s); };
  auto pfull_bar = [&](int s) { return bar_base + 8u * (3 * STAGES + 4 + s); };
  auto pfree_bar = [&](int s) { return bar_base + 8u * (3 * STAGES + 4 + PANELS + s); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * STAGES + 4 + 2 * PANELS);
  static_assert(8 * (3 * STAGES + 4 + 2 * PANELS + 1) <= 256, "barrier block");
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_base));
  float* s_ss = reinterpret_cast<float*>(smem_raw + (bar_base + 256 - smem_base));   // scale[256] | shift[256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int m_tiles = (p.M + 2 * BM - 1) / (2 * BM), n_tiles = p.N / BN2;
  const int num_tiles = m_tiles * n_tiles;
  const int kblocks = p.K / BK;

  if (warp == 0 && lane == 0) {
    if (smem_base & 1023u) __trap();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    if (p.kb1 < kblocks) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    if (p.res_mode == RES_TMA) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); mbar_init(xfull_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 8 * G); }
    for (int s = 0; s < PANELS; ++s) { mbar_init(pfull_bar(s), 1); mbar_init(pfree_bar(s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= operand producer: this CTA's 128 rows of A and 128 of the tile's 256 W rows =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        const int row0 = m_blk * 2 * BM + (int)rank * BM;
        const int wrow = (p.group_rows > 0 ? (m_blk * 2 * BM / p.group_rows) * p.N : 0) + n_blk * BN2 + (int)rank * BNH;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          if (kb < p.kb1) tma_load_3d(sa, &tmA, full_bar(stage), kb * BK, row0, 0);
          else load_a2(sa, &tmA2, full_bar(stage), (kb - p.kb1) * BK, row0, p.a2_wo, p.a2_ho, p.a2_st, p.a2_ss);
          tma_load_3d(sa + 2 * A_PLANE_BYTES, &tmW, full_bar(stage), kb * BK, wrow, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank != 0) {
      // ================= relay (peer CTA): report each locally completed stage to the leader =================
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        uint32_t remote[STAGES];
        for (int s = 0; s < STAGES; ++s) remote[s] = mapa_rank(xfull_bar(s), 0);
        for (int tile = pair; tile < num_tiles; tile += npairs)
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(full_bar(stage), phase);
            mbar_arrive_cluster(remote[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
      }
    } else {
      // ================= MMA issuer (leader CTA) =================
      constexpr uint32_t idesc = make_idesc(2 * BM, BN2);
      const uint64_t desc0 = make_smem_desc(smem_base);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs, ++it) {
        const int as = it & 1;
        mbar_wait(tempty_bar(as), ((it >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN2);
#pragma unroll 1
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          mbar_wait(xfull_bar(stage), phase);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint64_t a_hi = desc0 + (uint64_t)(stage * (STAGE_BYTES >> 4)), a_mid = a_hi + (A_PLANE_BYTES >> 4);
            const uint64_t w_hi = a_hi + ((2 * A_PLANE_BYTES) >> 4), w_mid = w_hi + ((BNH * BK * 2) >> 4);
            const uint32_t first = kb == 0 ? 0u : 1u;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma2_bf16(tmem_d, a_mid + 2 * k, w_hi + 2 * k, idesc, k == 0 ? first : 1u);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma2_bf16(tmem_d, a_hi + 2 * k, w_mid + 2 * k, idesc, 1);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma2_bf16(tmem_d, a_hi + 2 * k, w_hi + 2 * k, idesc, 1);
            umma2_commit_mc(empty_bar(stage));
            if (kb == kblocks - 1) umma2_commit_mc(tfull_bar(as));
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ================= panel producer =================
    if (lane == 0 && !p.tail_panels) {
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        const int row0 = m_blk * 2 * BM + (int)rank * BM;
        for (int j = 0; j < PPT; ++j) {
          mbar_wait(pfree_bar(slot), phase ^ 1);
          if (p.res_mode == RES_TMA) {
            const int col = n_blk * BN2 + j * 64;
            mbar_expect_tx(pfull_bar(slot), PANEL_BYTES);
            if (p.out_fmt == FMT_F32) tma_load_3d(panel_base + slot * PANEL_BYTES, &tmR, pfull_bar(slot), 0, row0, col >> 5);
            else tma_load_3d(panel_base + slot * PANEL_BYTES, &tmR, pfull_bar(slot), col, row0, 0);
          } else {
            mbar_arrive(pfull_bar(slot));
          }
          if (++slot == PANELS) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ================= epilogue (warps 3..6): this CTA's 128 rows x 256 columns =================
    const int lg = warp & 3;
    const int et = (threadIdx.x - EPI_WARP0 * 32) & 127;
    const int r = lg * 32 + lane;
    const uint32_t tempty_remote0 = mapa_rank(tempty_bar(0), 0), tempty_remote1 = mapa_rank(tempty_bar(1), 0);
    float* s_scale = s_ss;
    float* s_shift = s_ss + BN2;
    const int eg = (warp - EPI_WARP0) >> 2;               // epilogue group: takes the panels j with j % G == eg (PPT = 4)
    const int gbar = 1 + eg;
    int it = 0, prev_slot = -1;
    for (int tile = pair; tile < num_tiles; tile += npairs, ++it) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int as = it & 1;
      const int grp = p.group_rows > 0 ? m_blk * 2 * BM / p.group_rows : 0;
      const int gcol = grp * p.N;
      // one scale / shift copy for all groups: nobody restages it while a group still reads the previous tile's values
      if (G > 1) asm volatile("bar.sync 3, %0;" ::"r"(128 * G) : "memory");
      for (int c = threadIdx.x - EPI_WARP0 * 32; c < BN2; c += 128 * G) {
        s_scale[c] = p.scale ? __ldg(p.scale + gcol + n_blk * BN2 + c) : 1.f;
        s_shift[c] = p.shift ? __ldg(p.shift + gcol + n_blk * BN2 + c) : 0.f;
      }
      if (G > 1) asm volatile("bar.sync 3, %0;" ::"r"(128 * G) : "memory");
      else asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(tfull_bar(as), (it >> 1) & 1);
      tcgen05_fence_after();
      const int row0 = m_blk * 2 * BM + (int)rank * BM;
      const long long row = (long long)row0 + r;
      const bool row_ok = row < p.M;
      const long long rrow = p.res_mod > 0 ? row % p.res_mod : row;
#pragma unroll 1
      for (int j = 0; j < PPT; ++j) {
        if (G > 1 && j % G != eg) continue;
        const int gp = it * PPT + j;                      // running panel index (the producer's order)
        const int slot = gp % PANELS;
        const uint32_t pphase = (uint32_t)((gp / PANELS) & 1);
        // tail mode (one tile per pair, no residual): after tfull every MMA has read its operands and no load is outstanding, so
        // panel j > 0 is built in the first half of operand stage j - 1: four buffers, the stores of all panels in flight at once
        const bool tail = p.tail_panels != 0;
        if (!tail) mbar_wait(pfull_bar(slot), pphase);
        const uint32_t pb = tail ? (j == 0 ? panel_base : smem_base + (uint32_t)(j - 1) * STAGE_BYTES) : panel_base + slot * PANEL_BYTES;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          uint32_t acc[32];
          tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN2 + j * 64 + h * 32), acc);
          const int cl = j * 64 + h * 32;
          float v[32];
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 sc = *reinterpret_cast<const float4*>(s_scale + cl + 4 * q4);
            const float4 sh = *reinterpret_cast<const float4*>(s_shift + cl + 4 * q4);
            v[4 * q4] = fmaf(__uint_as_float(acc[4 * q4]), sc.x, sh.x);
            v[4 * q4 + 1] = fmaf(__uint_as_float(acc[4 * q4 + 1]), sc.y, sh.y);
            v[4 * q4 + 2] = fmaf(__uint_as_float(acc[4 * q4 + 2]), sc.z, sh.z);
            v[4 * q4 + 3] = fmaf(__uint_as_float(acc[4 * q4 + 3]), sc.w, sh.w);
          }
          if (p.res_mode == RES_TMA) {
            if (p.out_fmt == FMT_F32) {
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4) {
                const uint4 t = lds128(swz(pb + h * SUB_BYTES, r, q4));
                v[4 * q4] += __uint_as_float(t.x); v[4 * q4 + 1] += __uint_as_float(t.y);
                v[4 * q4 + 2] += __uint_as_float(t.z); v[4 * q4 + 3] += __uint_as_float(t.w);
              }
            } else {
#pragma unroll
              for (int q8 = 0; q8 < 4; ++q8) {
                const uint4 a = lds128(swz(pb, r, 4 * h + q8));
                const uint4 b = lds128(swz(pb + SUB_BYTES, r, 4 * h + q8));
                v[8 * q8] += bf16_lo_to_f32(a.x) + bf16_lo_to_f32(b.x); v[8 * q8 + 1] += bf16_hi_to_f32(a.x) + bf16_hi_to_f32(b.x);
                v[8 * q8 + 2] += bf16_lo_to_f32(a.y) + bf16_lo_to_f32(b.y); v[8 * q8 + 3] += bf16_hi_to_f32(a.y) + bf16_hi_to_f32(b.y);
                v[8 * q8 + 4] += bf16_lo_to_f32(a.z) + bf16_lo_to_f32(b.z); v[8 * q8 + 5] += bf16_hi_to_f32(a.z) + bf16_hi_to_f32(b.z);
                v[8 * q8 + 6] += bf16_lo_to_f32(a.w) + bf16_lo_to_f32(b.w); v[8 * q8 + 7] += bf16_hi_to_f32(a.w) + bf16_hi_to_f32(b.w);
              }
            }
          } else if (p.res_mode == RES_DIRECT && row_ok) {
            const int n0 = n_blk * BN2 + cl;
            if (p.res_fmt == FMT_F32) {
              const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res) + rrow * p.ldr + n0);
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4) {
                const float4 t = __ldg(rp + q4);
                v[4 * q4] += t.x; v[4 * q4 + 1] += t.y; v[4 * q4 + 2] += t.z; v[4 * q4 + 3] += t.w;
              }
            } else {
              const __nv_bfloat16* hp = split_hi(p.res, rrow, p.ldr) + n0;
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4) {
                const float4 t = load_split4(hp + 4 * q4, hp + p.ldr + 4 * q4);
                v[4 * q4] += t.x; v[4 * q4 + 1] += t.y; v[4 * q4 + 2] += t.z; v[4 * q4 + 3] += t.w;
              }
            }
          }
          if (p.act == ACT_RELU) {
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c], 0.f);
          }
          if (p.out_fmt == FMT_F32) {
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4)
              sts128(swz(pb + h * SUB_BYTES, r, q4), make_uint4(__float_as_uint(v[4 * q4]), __float_as_uint(v[4 * q4 + 1]),
                                                                __float_as_uint(v[4 * q4 + 2]), __float_as_uint(v[4 * q4 + 3])));
          } else {
#pragma unroll
            for (int q8 = 0; q8 < 4; ++q8) {
              uint4 hi, mid;
              split_bf16x2(v[8 * q8], v[8 * q8 + 1], hi.x, mid.x);
              split_bf16x2(v[8 * q8 + 2], v[8 * q8 + 3], hi.y, mid.y);
              split_bf16x2(v[8 * q8 + 4], v[8 * q8 + 5], hi.z, mid.z);
              split_bf16x2(v[8 * q8 + 6], v[8 * q8 + 7], hi.w, mid.w);
              sts128(swz(pb, r, 4 * h + q8), hi);
              sts128(swz(pb + SUB_BYTES, r, 4 * h + q8), mid);
            }
          }
        }
        if (j + G >= PPT) {                                  // this group's last panel: its part of the accumulator is read
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(as ? tempty_remote1 : tempty_remote0);
        }
        fence_proxy_async();
        asm volatile("bar.sync %0, 128;" ::"r"(gbar) : "memory");
        if (et == 0) {
          const int col = gcol + n_blk * BN2 + j * 64;
          const int orow = row0 - grp * p.group_rows;
          if (p.out_fmt == FMT_F32) tma_store_3d(&tmC, pb, 0, orow, col >> 5);
          else tma_store_3d(&tmC, pb, col, orow, 0);
          bulk_commit();
          // keep at most PANELS-2 stores in flight, then recycle the buffer(s) whose store has drained
          if (tail) {
            // nothing to recycle: every panel has its own buffer (bulk_wait_all below keeps shared memory alive)
          } else if constexpr (PANELS == 1 || G > 1) {
            bulk_wait_read<0>();
            mbar_arrive(pfree_bar(slot));
          } else {
            bulk_wait_read<PANELS - 2>();
            if constexpr (PANELS == 2) {
              mbar_arrive(pfree_bar(slot));
            } else {
              if (prev_slot >= 0) mbar_arrive(pfree_bar(prev_slot));
              prev_slot = slot;
            }
          }
        }
      }
    }
    if (et == 0) bulk_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                                        // the leader's MMAs read this CTA's shared memory until its last commit
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// =============================================================================================================
// CTA-pair form of the fused conv4 -> conv1 kernel for the 1024-channel stage (layer3 of ir-CSN-152: 36 bottlenecks, N1 = 1024,
// N2 = 256, K = 256).  gemm_fused2_kernel<1024, 256> is correct but one CTA per 128-row block pulls 4.1 MB through L2 -> shared
// memory (A re-read for each of its 8 sub-tiles, all of W4 and W1'), which is what bounds it (70 us per block at 8 clips against
// 30 + 37 us for the separate pair / single-CTA launches).  Here a cluster of two CTAs owns 256 rows and issues
// tcgen05.mma.cta_group::2 (M = 256): each CTA loads its own 128 rows of A and HALF of every weight tile, and the conv4 sub-tiles
// are 256 columns wide (A is read 4 times instead of 8): 2.6 MB per CTA.
//   tensor memory (per CTA): [0, 256) the conv4 accumulator of the current sub-tile (single: the second GEMM over its four
//   panels keeps the tensor pipe busy while the epilogue drains it), [256, 512) the second GEMM's accumulator.
//   barriers as in gemm_fused2_kernel; `tempty`, `pready`, `d2empty` live in the leader and count arrivals from both CTAs
//   (remote mbarrier.arrive), `xfull[s]` is the peer's relay of its TMA completions, every tcgen05.commit is multicast to both.
// =============================================================================================================
template <int N1, int N2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FUSED2_THREADS, 1)
gemm_fused2p_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                    const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmC,
                    const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmW2,
                    const __grid_constant__ CUtensorMap tmC2, Params p, Params2 q) {
  constexpr int BNS = 256, BNH = 128, STAGES = 2, PANELS = 3, STAGE_BYTES = 65536;
  constexpr int NSUB = N1 / BNS, PPS = BNS / 64, P1 = N1 / 64, P2 = N2 / 64;
  constexpr int W2_CHUNK = (N2 / 2) * 256;                  // bytes of one k-block of this CTA's half of W1' (hi + mid planes)
  constexpr int KBW = STAGE_BYTES / W2_CHUNK;               // k-blocks of the second GEMM per ring stage
  constexpr int D2_COL = BNS;
  constexpr int G = FUSED2_EPI_GROUPS;
  static_assert(N1 % BNS == 0 && (N2 == 256 || N2 == 128) && PPS % KBW == 0 && D2_COL + N2 <= 512, "built for the 1024 -> 256 and 512 -> 128 shapes");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t panel_base = smem_base + STAGES * STAGE_BYTES;
  const uint32_t bar_base = panel_base + PANELS * PANEL_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 + s); };
  auto xfull_bar = [&](int s) { return bar_base + 8u * (4 + s); };
  const uint32_t tfull_bar = bar_base + 8u * 6, tempty_bar = bar_base + 8u * 7;
  auto pfull_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto pfree_bar = [&](int s) { return bar_base + 8u * (11 + s); };
  auto pready_bar = [&](int s) { return bar_base + 8u * (14 + s); };
  auto pcons_bar = [&](int s) { return bar_base + 8u * (17 + s); };
  const uint32_t d2full_bar = bar_base + 8u * 20, d2empty_bar = bar_base + 8u * 21;
  const uint32_t tmem_slot = bar_base + 8u * 22;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int m_pairs = (p.M + 2 * BM - 1) / (2 * BM);
  const int kblocks = p.K / BK;

  if (warp == 0 && lane == 0) {
    if (smem_base & 1023u) __trap();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    if (p.kb1 < kblocks) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC2) : "memory");
    if (p.res_mode == RES_TMA) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); mbar_init(xfull_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 2 * 4 * G);
    for (int s = 0; s < PANELS; ++s) {
      mbar_init(pfull_bar(s), 1); mbar_init(pfree_bar(s), 1); mbar_init(pready_bar(s), 2); mbar_init(pcons_bar(s), 1);
    }
    mbar_init(d2full_bar, 1);
    mbar_init(d2empty_bar, 2 * 4 * (P2 < G ? P2 : G));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // Ring order = the MMA warp's consumption order, per row-block pair and sub-tile: the K stages [A rows | W4 half] of the
  // sub-tile, then the PPS / KBW chunks of W1' that its four panels multiply
  if (warp == 0) {
    // ================= operand producer (both CTAs): own 128 rows of A, own half of the weight rows =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pb = pair; pb < m_pairs; pb += npairs) {
        const int row0 = pb * 2 * BM + (int)rank * BM;
        for (int sub = 0; sub < NSUB; ++sub) {
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t sa = smem_base + stage * STAGE_BYTES;
            mbar_expect_tx(full_bar(stage), STAGE_BYTES);
            if (kb < p.kb1) tma_load_3d(sa, &tmA, full_bar(stage), kb * BK, row0, 0);
            else load_a2(sa, &tmA2, full_bar(stage), (kb - p.kb1) * BK, row0, p.a2_wo, p.a2_ho, p.a2_st, p.a2_ss);
            tma_load_3d(sa + 2 * A_PLANE_BYTES, &tmW, full_bar(stage), kb * BK, sub * BNS + (int)rank * BNH, 0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          for (int c = 0; c < PPS / KBW; ++c) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t sa = smem_base + stage * STAGE_BYTES;
            mbar_expect_tx(full_bar(stage), KBW * W2_CHUNK);
            for (int e = 0; e < KBW; ++e)
              tma_load_3d(sa + e * W2_CHUNK, &tmW2, full_bar(stage), ((sub * (PPS / KBW) + c) * KBW + e) * BK, (int)rank * (N2 / 2), 0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank != 0) {
      // ================= relay (peer CTA): report each locally completed stage to the leader =================
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        uint32_t remote[STAGES];
        for (int s = 0; s < STAGES; ++s) remote[s] = mapa_rank(xfull_bar(s), 0);
        for (int pb = pair; pb < m_pairs; pb += npairs)
          for (int i = 0; i < NSUB * (kblocks + PPS / KBW); ++i) {
            mbar_wait(full_bar(stage), phase);
            mbar_arrive_cluster(remote[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
      }
    } else {
      // ================= MMA issuer (leader CTA) =================
      constexpr uint32_t idesc1 = make_idesc(2 * BM, BNS), idesc2 = make_idesc(2 * BM, N2);
      const uint64_t desc0 = make_smem_desc(smem_base);
      const uint64_t pdesc0 = make_smem_desc(panel_base);
      int stage = 0, pslot = 0, tl = 0, nsub = 0;
      uint32_t phase = 0, ready_ph = 0;                      // ready_ph bit s: parity of the next pready[s] completion
      for (int pb = pair; pb < m_pairs; pb += npairs, ++tl) {
        for (int sub = 0; sub < NSUB; ++sub, ++nsub) {
          // ---- conv4 sub-tile -> D1 (both epilogues have drained the previous one) ----
          mbar_wait(tempty_bar, (uint32_t)((nsub & 1) ^ 1));
          tcgen05_fence_after();
#pragma unroll 1
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(full_bar(stage), phase);
            mbar_wait(xfull_bar(stage), phase);
            tcgen05_fence_after();
            if (elect_one()) {
              const uint64_t a_hi = desc0 + (uint64_t)(stage * (STAGE_BYTES >> 4)), a_mid = a_hi + (A_PLANE_BYTES >> 4);
              const uint64_t w_hi = a_hi + ((2 * A_PLANE_BYTES) >> 4), w_mid = w_hi + ((BNH * BK * 2) >> 4);
              const uint32_t first = kb == 0 ? 0u : 1u;
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma2_bf16(tmem_base, a_mid + 2 * k, w_hi + 2 * k, idesc1, k == 0 ? first : 1u);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma2_bf16(tmem_base, a_hi + 2 * k, w_mid + 2 * k, idesc1, 1);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma2_bf16(tmem_base, a_hi + 2 * k, w_hi + 2 * k, idesc1, 1);
              umma2_commit_mc(empty_bar(stage));
              if (kb == kblocks - 1) umma2_commit_mc(tfull_bar);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          // ---- second GEMM over the sub-tile's four panels (k-block j = output panel j of x', read from both CTAs' buffers) ----
          int w2_stage = 0;
          for (int jj = 0; jj < PPS; ++jj) {
            const int j = sub * PPS + jj;
            if (j == 0) {
              mbar_wait(d2empty_bar, (uint32_t)((tl & 1) ^ 1));   // both epilogues have drained the previous pair-block's t1'
              tcgen05_fence_after();
            }
            if (jj % KBW == 0) {
              mbar_wait(full_bar(stage), phase);
              mbar_wait(xfull_bar(stage), phase);
              w2_stage = stage;
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            mbar_wait(pready_bar(pslot), (ready_ph >> pslot) & 1u);
            ready_ph ^= 1u << pslot;
            tcgen05_fence_after();
            if (elect_one()) {
              const uint32_t tmem_d2 = tmem_base + (uint32_t)D2_COL;
              const uint64_t a_hi = pdesc0 + (uint64_t)(pslot * (PANEL_BYTES >> 4)), a_mid = a_hi + (SUB_BYTES >> 4);
              const uint64_t w_hi = desc0 + (uint64_t)(w2_stage * (STAGE_BYTES >> 4)) + (uint64_t)((jj % KBW) * (W2_CHUNK >> 4));
              const uint64_t w_mid = w_hi + (((N2 / 2) * BK * 2) >> 4);
              const uint32_t first = j == 0 ? 0u : 1u;
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma2_bf16(tmem_d2, a_mid + 2 * k, w_hi + 2 * k, idesc2, k == 0 ? first : 1u);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma2_bf16(tmem_d2, a_hi + 2 * k, w_mid + 2 * k, idesc2, 1);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma2_bf16(tmem_d2, a_hi + 2 * k, w_hi + 2 * k, idesc2, 1);
              umma2_commit_mc(pcons_bar(pslot));
              if (jj % KBW == KBW - 1) umma2_commit_mc(empty_bar(w2_stage));
              if (j == P1 - 1) umma2_commit_mc(d2full_bar);
            }
            __syncwarp();
            if (++pslot == PANELS) pslot = 0;
          }
        }
        for (int jp = 0; jp < P2; ++jp)                        // the t1' panels use ring slots too (not operands)
          if (++pslot == PANELS) pslot = 0;
      }
    }
  } else if (warp == 2) {
    // ================= panel producer (both CTAs) =================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (int pb = pair; pb < m_pairs; pb += npairs) {
        const int row0 = pb * 2 * BM + (int)rank * BM;
        for (int j = 0; j < P1 + P2; ++j) {
          mbar_wait(pfree_bar(slot), phase ^ 1);
          if (j < P1 && p.res_mode == RES_TMA) {
            mbar_expect_tx(pfull_bar(slot), PANEL_BYTES);
            tma_load_3d(panel_base + slot * PANEL_BYTES, &tmR, pfull_bar(slot), j * 64, row0, 0);
          } else {
            mbar_arrive(pfull_bar(slot));
          }
          if (++slot == PANELS) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ================= epilogue (both CTAs): two warp quartets; group eg takes the panels jj with jj % 2 == eg =================
    const int lg = warp & 3;
    const int eg = (warp - EPI_WARP0) >> 2;
    const int et = (threadIdx.x - EPI_WARP0 * 32) & 127;
    const int r = lg * 32 + lane;
    const int gbar = 1 + eg;
    const uint32_t tempty_remote = mapa_rank(tempty_bar, 0), d2empty_remote = mapa_rank(d2empty_bar, 0);
    uint32_t pready_remote[PANELS];
#pragma unroll
    for (int i = 0; i < PANELS; ++i) pready_remote[i] = mapa_rank(pready_bar(i), 0);
    int tl = 0, nsub = 0;
    int xuse[PANELS];                                        // x' panels each ring slot has carried so far (parity of its pcons barrier)
#pragma unroll
    for (int i = 0; i < PANELS; ++i) xuse[i] = 0;
    for (int pb = pair; pb < m_pairs; pb += npairs, ++tl) {
      const int row0 = pb * 2 * BM + (int)rank * BM;
      const int gp0 = tl * (P1 + P2);
      for (int sub = 0; sub < NSUB; ++sub, ++nsub) {
        mbar_wait(tfull_bar, (uint32_t)(nsub & 1));
        tcgen05_fence_after();
#pragma unroll
        for (int jj = 0; jj < PPS; ++jj) {
          const int gp = gp0 + PPS * sub + jj, slot = gp % PANELS;
          uint32_t cons_parity = 0, pr_remote = 0;
#pragma unroll
          for (int i = 0; i < PANELS; ++i)
            if (i == slot) { cons_parity = (uint32_t)(xuse[i] & 1); ++xuse[i]; pr_remote = pready_remote[i]; }
          if (jj % G != eg) continue;
          mbar_wait(pfull_bar(slot), (uint32_t)((gp / PANELS) & 1));
          const uint32_t pb_addr = panel_base + slot * PANEL_BYTES;
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            uint32_t acc[32];
            tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(jj * 64 + h * 32), acc);
            const int c0 = sub * BNS + jj * 64 + h * 32;     // column of x'
            float v[32];
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              const float4 sc = p.scale ? __ldg(reinterpret_cast<const float4*>(p.scale + c0) + c4) : make_float4(1.f, 1.f, 1.f, 1.f);
              const float4 sh = p.shift ? __ldg(reinterpret_cast<const float4*>(p.shift + c0) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
              v[4 * c4] = fmaf(__uint_as_float(acc[4 * c4]), sc.x, sh.x);
              v[4 * c4 + 1] = fmaf(__uint_as_float(acc[4 * c4 + 1]), sc.y, sh.y);
              v[4 * c4 + 2] = fmaf(__uint_as_float(acc[4 * c4 + 2]), sc.z, sh.z);
              v[4 * c4 + 3] = fmaf(__uint_as_float(acc[4 * c4 + 3]), sc.w, sh.w);
            }
            if (p.res_mode == RES_TMA) {
#pragma unroll
              for (int c8 = 0; c8 < 4; ++c8) {
                const uint4 a = lds128(swz(pb_addr, r, 4 * h + c8));
                const uint4 b = lds128(swz(pb_addr + SUB_BYTES, r, 4 * h + c8));
                v[8 * c8] += bf16_lo_to_f32(a.x) + bf16_lo_to_f32(b.x); v[8 * c8 + 1] += bf16_hi_to_f32(a.x) + bf16_hi_to_f32(b.x);
                v[8 * c8 + 2] += bf16_lo_to_f32(a.y) + bf16_lo_to_f32(b.y); v[8 * c8 + 3] += bf16_hi_to_f32(a.y) + bf16_hi_to_f32(b.y);
                v[8 * c8 + 4] += bf16_lo_to_f32(a.z) + bf16_lo_to_f32(b.z); v[8 * c8 + 5] += bf16_hi_to_f32(a.z) + bf16_hi_to_f32(b.z);
                v[8 * c8 + 6] += bf16_lo_to_f32(a.w) + bf16_lo_to_f32(b.w); v[8 * c8 + 7] += bf16_hi_to_f32(a.w) + bf16_hi_to_f32(b.w);
              }
            }
            if (p.act == ACT_RELU) {
#pragma unroll
              for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c], 0.f);
            }
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              uint4 hi, mid;
              split_bf16x2(v[8 * c8], v[8 * c8 + 1], hi.x, mid.x);
              split_bf16x2(v[8 * c8 + 2], v[8 * c8 + 3], hi.y, mid.y);
              split_bf16x2(v[8 * c8 + 4], v[8 * c8 + 5], hi.z, mid.z);
              split_bf16x2(v[8 * c8 + 6], v[8 * c8 + 7], hi.w, mid.w);
              sts128(swz(pb_addr, r, 4 * h + c8), hi);
              sts128(swz(pb_addr + SUB_BYTES, r, 4 * h + c8), mid);
            }
          }
          if (jj + G >= PPS) {                                 // this group's last panel of the sub-tile: its columns of D1 are read
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_remote);
          }
          fence_proxy_async();                               // visible to the TMA store AND to the second GEMM's MMAs
          asm volatile("bar.sync %0, 128;" ::"r"(gbar) : "memory");
          if (et == 0) {
            mbar_arrive_cluster_release(pr_remote);          // leader: this CTA's half of the panel is final (default-scope arrive: 70 vs 74 us)
            tma_store_3d(&tmC, pb_addr, sub * BNS + jj * 64, row0, 0);
            bulk_commit();
            bulk_wait_read<0>();                             // the store has read the panel ...
            mbar_wait(pcons_bar(slot), cons_parity);         // ... and so has the second GEMM
            mbar_arrive(pfree_bar(slot));
          }
        }
      }
      // ---- second accumulator: t1' = relu(scale2 * D2 + shift2), fp32 panels ----
      bool d2_waited = false;
#pragma unroll 1
      for (int jp = 0; jp < P2; ++jp) {
        const int gp = gp0 + P1 + jp, slot = gp % PANELS;
        if (jp % G != eg) continue;
        if (!d2_waited) {
          mbar_wait(d2full_bar, (uint32_t)(tl & 1));
          tcgen05_fence_after();
          d2_waited = true;
        }
        mbar_wait(pfull_bar(slot), (uint32_t)((gp / PANELS) & 1));
        const uint32_t pb_addr = panel_base + slot * PANEL_BYTES;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          uint32_t acc[32];
          tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(D2_COL + jp * 64 + h * 32), acc);
          const int c0 = jp * 64 + h * 32;
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(q.scale2 + c0) + c4);
            const float4 sh = __ldg(reinterpret_cast<const float4*>(q.shift2 + c0) + c4);
            uint4 o;
            o.x = __float_as_uint(fmaxf(fmaf(__uint_as_float(acc[4 * c4]), sc.x, sh.x), 0.f));
            o.y = __float_as_uint(fmaxf(fmaf(__uint_as_float(acc[4 * c4 + 1]), sc.y, sh.y), 0.f));
            o.z = __float_as_uint(fmaxf(fmaf(__uint_as_float(acc[4 * c4 + 2]), sc.z, sh.z), 0.f));
            o.w = __float_as_uint(fmaxf(fmaf(__uint_as_float(acc[4 * c4 + 3]), sc.w, sh.w), 0.f));
            sts128(swz(pb_addr + h * SUB_BYTES, r, c4), o);
          }
        }
        if (jp + G >= P2) {                                  // this group's last panel of t1'
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(d2empty_remote);
        }
        fence_proxy_async();
        asm volatile("bar.sync %0, 128;" ::"r"(gbar) : "memory");
        if (et == 0) {
          tma_store_3d(&tmC2, pb_addr, 0, row0, (jp * 64) >> 5);
          bulk_commit();
          bulk_wait_read<0>();
          mbar_arrive(pfree_bar(slot));
        }
      }
    }
    if (et == 0) bulk_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                                        // the leader's MMAs read this CTA's shared memory until its last commit
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

using CfgWide = Cfg<128, 2, 3, 2>;   // memory-bound shapes: panels pipelined (load residual | compute | store), two epilogue groups
using CfgDeep = Cfg<128, 3, 1>;   // K >= 512: deeper operand ring, one panel buffer
using CfgN64 = Cfg<64, 3, 2, 2>;     // N == 64 (or N % 128 != 0): one panel per tile, the two epilogue groups alternate over tiles
using CfgBig = Cfg<256, 2, 1>;    // K >= 512, N % 256 == 0 and enough row blocks: 128 x 256 tiles (A tile reused over 256 columns,
                                  // 25 % less L2 -> shared-memory traffic per FLOP; these shapes are L2-bandwidth / tensor bound)

constexpr int PAIR_SMEM_BYTES = 3 * 65536 + PANEL_BYTES + 256 + 2 * 256 * 4;   // gemm2_bf16x3_kernel<3, 1> and <2, 3> (2 x 64 KB + 3 panels: same size)
static_assert(2 * 65536 + 3 * PANEL_BYTES == 3 * 65536 + PANEL_BYTES, "pair kernel configurations share one shared-memory size");
static_assert(PAIR_SMEM_BYTES <= 232448, "over the 227 KB shared-memory limit");
// gemm2_bf16x3_kernel<2, 3, 2, 128>: 2 x 48 KB operand stages + 3 panels
constexpr int PAIR128_STAGE_BYTES = 2 * A_PLANE_BYTES + 2 * 64 * BK * 2;
constexpr int PAIR128_SMEM_BYTES = 2 * PAIR128_STAGE_BYTES + 3 * PANEL_BYTES + 256 + 2 * 256 * 4;
static_assert(PAIR128_SMEM_BYTES <= 232448, "over the 227 KB shared-memory limit");

static char g_err[256] = "";
static EncodeTiledFn g_encode = nullptr;
static DeviceOnce g_dev_once;
static cudaError_t init_once() {
  return g_dev_once.run([]() -> cudaError_t {
    if (!g_encode) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
      if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) return e != cudaSuccess ? e : cudaErrorNotSupported;
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_bf16x3_kernel<CfgWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgWide::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_bf16x3_kernel<CfgDeep>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgDeep::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_bf16x3_kernel<CfgN64>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgN64::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_bf16x3_kernel<CfgBig>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgBig::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm2_bf16x3_kernel<3, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm2_bf16x3_kernel<2, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm2_bf16x3_kernel<3, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm2_bf16x3_kernel<2, 3, 2, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR128_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_fused2_kernel<256, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgWide::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_fused2_kernel<256, 64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgWide::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_fused2_kernel<256, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgWide::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_fused2_kernel<512, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgWide::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_fused2_kernel<1024, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgWide::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_fused2_kernel<512, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgWide::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_fused2p_kernel<1024, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgWide::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_fused2p_kernel<512, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgWide::SMEM_BYTES);
    return e;
  });
}

static bool encode3(CUtensorMap* map, CUtensorMapDataType dt, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                    uint64_t s1_bytes, uint64_t s2_bytes, uint32_t b0, uint32_t b1, uint32_t b2) {
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1_bytes, s2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(map, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled failed (%d): dims %llu,%llu,%llu strides %llu,%llu box %u,%u,%u", (int)r,
             (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)s1_bytes,
             (unsigned long long)s2_bytes, b0, b1, b2);
    return false;
  }
  return true;
}

// split tensor [rows, ld] restricted to `cols` columns: dims {cols, rows, plane}, box {64, box_rows, 2}
static bool encode_split_map(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows) {
  return encode3(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, cols, rows, 2, ld * 4, ld * 2, 64, box_rows, 2);
}
// fp32 matrix [rows, ld] restricted to `cols` columns, as {32, rows, cols/32}: box {32, 128, 2} = one 64-column panel
static bool encode_f32_panel_map(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t ld) {
  return encode3(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, 32, rows, cols / 32, ld * 4, 128, 32, BM, 2);
}

// 128 x 256 tiles for the deep-K shapes when they still fill most of the machine
static bool use_big_tiles(const GemmArgs& a, int KT) {
  // measured: not faster than 128 x 128 (33-37 vs 31-33 us on M=16384 N=256 K=1024) -- the deep shapes are bound by L2 -> SM
  // bandwidth per CTA, which only operand sharing across a CTA pair fixes (gemm2_bf16x3_kernel); kept behind TUBER_BIG_TILES=1
  static const bool on = [] { const char* e = getenv("TUBER_BIG_TILES"); return e && e[0] == '1'; }();
  if (!on || KT < 512 || a.N % 256 != 0 || a.ksplit > 1) return false;
  const long long tiles = (long long)ceil_div(a.M, BM) * (a.N / 256);
  return tiles >= 100;
}

// CTA pairs (cta_group::2, 256 x 256 tiles) for the deep-K shapes
static bool use_pair_tiles(const GemmArgs& a, int KT) {
  const char* npe = getenv("TUBER_NO_PAIR_GEMM");           // read per call: the tests switch it between two plans of one process
  const bool off = npe && npe[0] == '1';
  static const int min_k = [] { const char* e = getenv("TUBER_PAIR_MINK"); return e ? atoi(e) : 512; }();
  if (off || KT < min_k || a.N % 256 != 0 || a.ksplit > 1) return false;
  if (a.group_rows > 0 && a.group_rows % (2 * BM) != 0) return false;
  const long long tiles = (long long)ceil_div(a.M, 2 * BM) * (a.N / 256);
  return tiles >= 48;                                       // at least ~2/3 of the 74 pairs busy
}

// CTA pairs with 256 x 128 tiles for the short-K, wide-N shapes with several rounds of tiles (conv4 of the 1024-channel stage).
// Measured (profiles/r2_pair128_experiment.json): 17 % less operand traffic through L2 -> shared memory and NOT faster than the
// single-CTA 128 x 128 kernel (43.4 us at 1912 MHz against 41-43 us at 1815-1830 MHz per launch; 49.5 us with three stages and two
// panels) -- the shape is not bound by the L2 throughput cap alone.  Correct (tests/test_ops_gpu.py::test_gemm_tc_pair128), kept off.
static bool use_pair128_tiles(const GemmArgs& a, int KT) {
  const char* e = getenv("TUBER_PAIR128");                  // read per call (tests / experiments): 1 = on, 0 = off
  const bool on = e ? e[0] == '1' : false;
  if (!on || KT > 256 || KT < 128 || a.N % 128 != 0 || a.N < 512 || a.ksplit > 1 || a.group_rows > 0) return false;
  const long long tiles = (long long)ceil_div(a.M, 2 * BM) * (a.N / 128);
  return tiles >= 4 * (device_num_sms() / 2);               // at least four rounds over the pairs
}

// A2 as a strided gather: dims {K, Wi, Hi, BTi, 2 planes}, element strides {1, ss, ss, st, 1}, box = 128 output voxels
static bool encode_a2(CUtensorMap* map, const GemmArgs& a, Params& p) {
  p.a2_wo = 0;
  if (!a.Ab) return true;
  if (a.ab_Wo <= 0) return encode_split_map(map, a.Ab, a.Kb, a.M, a.ldb, BM);
  if (!gemm_tc_strided_ab_ok(a.M, a.ab_Wo, a.ab_Ho, a.ab_Wi, a.ab_Hi, a.ab_BTi, a.ab_st_t, a.ab_st_s)) return false;
  const int per_bt = a.ab_Wo * a.ab_Ho;
  const int bh = per_bt >= BM ? BM / a.ab_Wo : a.ab_Ho, bbt = per_bt >= BM ? 1 : BM / per_bt;
  cuuint64_t dims[5] = {(cuuint64_t)a.Kb, (cuuint64_t)a.ab_Wi, (cuuint64_t)a.ab_Hi, (cuuint64_t)a.ab_BTi, 2};
  cuuint64_t strides[4] = {(cuuint64_t)a.ldb * 4, (cuuint64_t)a.ab_Wi * a.ldb * 4, (cuuint64_t)a.ab_Hi * a.ab_Wi * a.ldb * 4, (cuuint64_t)a.ldb * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)(a.ab_Wo * a.ab_st_s), (cuuint32_t)(bh * a.ab_st_s), (cuuint32_t)(bbt * a.ab_st_t), 2};
  cuuint32_t es[5] = {1, (cuuint32_t)a.ab_st_s, (cuuint32_t)a.ab_st_s, (cuuint32_t)a.ab_st_t, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(a.Ab), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled (strided A2) failed (%d): Wo=%d Ho=%d Wi=%d Hi=%d BTi=%d st=%d/%d", (int)r, a.ab_Wo, a.ab_Ho,
             a.ab_Wi, a.ab_Hi, a.ab_BTi, a.ab_st_t, a.ab_st_s);
    return false;
  }
  p.a2_wo = a.ab_Wo; p.a2_ho = a.ab_Ho; p.a2_st = a.ab_st_t; p.a2_ss = a.ab_st_s;
  return true;
}

template <class C>
static cudaError_t launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmW, const CUtensorMap& tmC, const CUtensorMap& tmR,
                              const Params& p, cudaStream_t st) {
  const int tiles = ceil_div(p.M, BM) * (p.N / C::BN) * (p.ksplit > 1 ? p.ksplit : 1);
  const int grid = tiles < device_num_sms() ? tiles : device_num_sms();
  return launch_pdl(gemm_bf16x3_kernel<C>, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, tmA, tmA2, tmW, tmC, tmR, p);
}

}  // namespace tc

const char* gemm_tc_last_error() { return tc::g_err; }

cudaError_t launch_gemm_tc(const GemmArgs& a, cudaStream_t st) {
  using namespace tc;
  cudaError_t e = init_once();
  if (e != cudaSuccess) return e;
  auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const int KT = a.K + (a.Ab ? a.Kb : 0);
  if (a.M <= 0 || a.N % 64 != 0 || a.K % 64 != 0 || a.lda % 8 != 0 || a.K > a.lda || a.a_fmt != FMT_SPLIT || a.Wp == nullptr ||
      a.act == ACT_SIGMOID || a.C2 != nullptr || !aligned16(a.A) || !aligned16(a.Wp) || !aligned16(a.C) || a.ldc % 8 != 0 ||
      a.N > a.ldc || (a.Ab && (a.Kb % 64 != 0 || a.Kb <= 0 || a.ldb % 8 != 0 || a.Kb > a.ldb || !aligned16(a.Ab)))) {
    snprintf(g_err, sizeof g_err, "gemm_tc: unsupported problem M=%d N=%d K=%d+%d lda=%d ldc=%d a_fmt=%d", a.M, a.N, a.K, a.Ab ? a.Kb : 0,
             a.lda, a.ldc, a.a_fmt);
    return cudaErrorInvalidValue;
  }
  // 128-wide tiles unless that leaves most of the machine idle (token-sized GEMMs): then 64-wide tiles double the CTA count
  int bn = (a.N % 128 == 0) ? 128 : 64;
  if (bn == 128 && (long long)ceil_div(a.M, BM) * (a.N / 128) * 2 * (a.ksplit > 1 ? a.ksplit : 1) <= device_num_sms()) bn = 64;
  if (use_big_tiles(a, KT)) bn = 256;
  const bool pair_tiles = use_pair_tiles(a, KT);
  if (pair_tiles) bn = 128;                                 // W box: each CTA of the pair loads 128 of the tile's 256 rows
  const bool pair128 = !pair_tiles && use_pair128_tiles(a, KT);
  if (pair128) bn = 64;                                     // W box: 64 of the tile's 128 rows
  Params p{};
  p.scale = a.scale; p.shift = a.shift;
  p.out_fmt = a.c_fmt;
  p.M = a.M; p.N = a.N; p.K = KT; p.act = a.act;
  p.kb1 = a.K / BK;
  p.res_mode = RES_NONE;
  const int groups = a.group_rows > 0 ? a.M / a.group_rows : 1;
  if (a.group_rows > 0 && (a.group_rows % BM != 0 || a.M % a.group_rows != 0 || a.res || a.Ab || (long long)groups * a.N > a.ldc)) {
    snprintf(g_err, sizeof g_err, "gemm_tc: bad grouped problem M=%d group_rows=%d N=%d ldc=%d", a.M, a.group_rows, a.N, a.ldc);
    return cudaErrorInvalidValue;
  }
  p.group_rows = a.group_rows > 0 ? a.group_rows : 0;
  int c_rows = a.group_rows > 0 ? (a.group_out_rows > 0 ? a.group_out_rows : a.group_rows) : a.M;
  const int c_cols = groups * a.N;
  p.ksplit = 1;
  if (a.ksplit > 1) {
    if (a.group_rows > 0 || a.Ab || a.act != ACT_NONE || a.c_fmt != FMT_F32 || (a.K / BK) % a.ksplit != 0 || a.part_rows < a.M || a.part_rows % BM != 0) {
      snprintf(g_err, sizeof g_err, "gemm_tc: bad split-K problem K=%d ksplit=%d part_rows=%d M=%d", a.K, a.ksplit, a.part_rows, a.M);
      return cudaErrorInvalidValue;
    }
    p.ksplit = a.ksplit; p.part_rows = a.part_rows;
    c_rows = (a.ksplit - 1) * a.part_rows + a.M;
  }
  if (a.res) {
    const bool tma_ok = a.res_mod <= 0 && a.res_fmt == a.c_fmt && aligned16(a.res) && a.ldr % 8 == 0 && a.N <= a.ldr;
    p.res_mode = tma_ok ? RES_TMA : RES_DIRECT;
    p.res = a.res; p.res_fmt = a.res_fmt; p.ldr = a.ldr; p.res_mod = a.res_mod;
  }
  CUtensorMap tmA, tmA2, tmW, tmC, tmR;
  if (!encode_split_map(&tmA, a.A, a.K, a.M, a.lda, BM)) return cudaErrorInvalidValue;
  tmA2 = tmA;
  if (!encode_a2(&tmA2, a, p)) return cudaErrorInvalidValue;
  if (!encode3(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a.Wp, KT, (uint64_t)groups * a.N, 2, (uint64_t)KT * 2, (uint64_t)groups * a.N * KT * 2, 64, bn, 2))
    return cudaErrorInvalidValue;
  bool ok = a.c_fmt == FMT_F32 ? encode_f32_panel_map(&tmC, a.C, c_cols, c_rows, a.ldc) : encode_split_map(&tmC, a.C, c_cols, c_rows, a.ldc, BM);
  if (!ok) return cudaErrorInvalidValue;
  tmR = tmC;
  if (p.res_mode == RES_TMA) {
    ok = a.c_fmt == FMT_F32 ? encode_f32_panel_map(&tmR, a.res, a.N, a.M, a.ldr) : encode_split_map(&tmR, a.res, a.N, a.M, a.ldr, BM);
    if (!ok) return cudaErrorInvalidValue;
  }
  if (pair128) {
    const int tiles = ceil_div(p.M, 2 * BM) * (p.N / 128);
    const int pairs = tiles < device_num_sms() / 2 ? tiles : device_num_sms() / 2;
    return launch_pdl(gemm2_bf16x3_kernel<2, 3, 2, 128>, dim3(2 * pairs), dim3(96 + 128 * 2), PAIR128_SMEM_BYTES, st, tmA, tmA2, tmW, tmC, tmR, p);
  }
  if (pair_tiles) {
    const int tiles = ceil_div(p.M, 2 * BM) * (p.N / 256);
    const int pairs = tiles < device_num_sms() / 2 ? tiles : device_num_sms() / 2;
    // short K: the epilogue dominates, pipeline its panels.  Also when every pair gets at most one tile (conv1 of the 1024-channel
    // stage at 8 clips: 64 tiles for 74 pairs): nothing follows the tile, so its epilogue is fully exposed
    static const int force23 = [] { const char* e = getenv("TUBER_PAIR_CFG23"); return e ? atoi(e) : -1; }();
    const bool single_round = tiles <= device_num_sms() / 2;
    // K = 512 with a residual (conv4 of the 2048-channel stage): 8 k-blocks do not hide a one-buffer, one-group epilogue of four
    // residual panels either
    static const bool res512 = [] { const char* e = getenv("TUBER_PAIR_RES512"); return !(e && e[0] == '0'); }();
    if (force23 == 1 || (force23 != 0 && KT < 512) || (force23 == 2 && single_round) ||
        (force23 != 0 && res512 && KT <= 512 && p.res_mode != RES_NONE))
      return launch_pdl(gemm2_bf16x3_kernel<2, 3, 2>, dim3(2 * pairs), dim3(96 + 128 * 2), PAIR_SMEM_BYTES, st, tmA, tmA2, tmW, tmC, tmR, p);
    // one tile per pair and no residual (conv1 of the 1024-channel stage at 8 clips): nothing follows the tile, its four panels go
    // through the freed operand stages with two epilogue groups instead of one buffer and one group
    static const bool no_tail = [] { const char* e = getenv("TUBER_PAIR_NO_TAIL"); return e && e[0] == '1'; }();
    if (single_round && p.res_mode == RES_NONE && !no_tail) {
      p.tail_panels = 1;
      return launch_pdl(gemm2_bf16x3_kernel<3, 1, 2>, dim3(2 * pairs), dim3(96 + 128 * 2), PAIR_SMEM_BYTES, st, tmA, tmA2, tmW, tmC, tmR, p);
    }
    return launch_pdl(gemm2_bf16x3_kernel<3, 1, 1>, dim3(2 * pairs), dim3(NUM_THREADS), PAIR_SMEM_BYTES, st, tmA, tmA2, tmW, tmC, tmR, p);
  }
  if (bn == 256) return launch_cfg<CfgBig>(tmA, tmA2, tmW, tmC, tmR, p, st);
  if (bn == 64) return launch_cfg<CfgN64>(tmA, tmA2, tmW, tmC, tmR, p, st);
  if (KT >= 512) return launch_cfg<CfgDeep>(tmA, tmA2, tmW, tmC, tmR, p, st);
  return launch_cfg<CfgWide>(tmA, tmA2, tmW, tmC, tmR, p, st);
}

bool gemm_tc_strided_ab_ok(int M, int Wo, int Ho, int Wi, int Hi, int BTi, int st_t, int st_s) {
  if (Wo <= 0 || Ho <= 0 || st_t < 1 || st_s < 1 || tc::BM % Wo != 0) return false;
  const int per_bt = Wo * Ho;
  if (per_bt >= tc::BM ? per_bt % tc::BM != 0 : tc::BM % per_bt != 0) return false;
  if (M % per_bt != 0) return false;
  const int BTo = M / per_bt;
  // output frame bt must sit at input frame bt * st_t for every clip, and the strided box must stay inside the TMA limits
  if (BTi != BTo * st_t || (Wo - 1) * st_s >= Wi || (Ho - 1) * st_s >= Hi) return false;
  const int bh = per_bt >= tc::BM ? tc::BM / Wo : Ho, bbt = per_bt >= tc::BM ? 1 : tc::BM / per_bt;
  return Wo * st_s <= 256 && bh * st_s <= 256 && bbt * st_t <= 256;
}

// which template configuration launch_gemm_tc picks (the per-launch profile reports them as separate kernels)
const char* gemm_tc_config_name(const GemmArgs& a) {
  const int KT = a.K + (a.Ab ? a.Kb : 0);
  int bn = (a.N % 128 == 0) ? 128 : 64;
  if (bn == 128 && (long long)ceil_div(a.M, tc::BM) * (a.N / 128) * 2 * (a.ksplit > 1 ? a.ksplit : 1) <= device_num_sms()) bn = 64;
  if (tc::use_pair_tiles(a, KT)) return "gemm2_bf16x3_pair";
  if (tc::use_pair128_tiles(a, KT)) return "gemm2_bf16x3_pair128";
  if (tc::use_big_tiles(a, KT)) return "gemm_bf16x3_big";
  if (bn == 64) return "gemm_bf16x3_n64";
  return KT >= 512 ? "gemm_bf16x3_deep" : "gemm_bf16x3_wide";
}

// conv4 (a: split output, TMA or no residual) fused with the next bottleneck's conv1 (W2p packed [2][N2][N]; fp32 result
// C2 [M, ldc2]); (N, N2) in {(256, 64), (256, 128), (512, 128), (512, 256), (1024, 256)} = the 256-, 512- and 1024-channel stages
cudaError_t launch_gemm_tc_fused2(const GemmArgs& a, const void* W2p, const float* scale2, const float* shift2, float* C2, int N2,
                                  int ldc2, cudaStream_t st) {
  using namespace tc;
  cudaError_t e = init_once();
  if (e != cudaSuccess) return e;
  auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const int KT = a.K + (a.Ab ? a.Kb : 0);
  const bool res_ok = !a.res || (a.res_mod <= 0 && a.res_fmt == FMT_SPLIT && aligned16(a.res) && a.ldr % 8 == 0 && a.N <= a.ldr);
  if (a.M <= 0 || !((a.N == 256 && (N2 == 64 || N2 == 128)) || (a.N == 512 && (N2 == 128 || N2 == 256)) || (a.N == 1024 && N2 == 256)) || a.K % 64 != 0 || a.lda % 8 != 0 || a.K > a.lda || a.a_fmt != FMT_SPLIT || a.Wp == nullptr ||
      a.c_fmt != FMT_SPLIT || a.act == ACT_SIGMOID || !aligned16(a.A) || !aligned16(a.Wp) || !aligned16(a.C) || a.ldc % 8 != 0 ||
      a.N > a.ldc || (a.Ab && (a.Kb % 64 != 0 || a.Kb <= 0 || a.ldb % 8 != 0 || a.Kb > a.ldb || !aligned16(a.Ab))) || !res_ok ||
      W2p == nullptr || !aligned16(W2p) || C2 == nullptr || !aligned16(C2) || ldc2 % 8 != 0 || N2 > ldc2) {
    snprintf(g_err, sizeof g_err, "gemm_tc_fused2: unsupported problem M=%d N=%d K=%d N2=%d", a.M, a.N, KT, N2);
    return cudaErrorInvalidValue;
  }
  Params p{};
  p.scale = a.scale; p.shift = a.shift;
  p.out_fmt = FMT_SPLIT;
  p.M = a.M; p.N = a.N; p.K = KT; p.act = a.act;
  p.kb1 = a.K / BK;
  p.res_mode = a.res ? RES_TMA : RES_NONE;
  Params2 q{scale2, shift2, N2};
  CUtensorMap tmA, tmA2, tmW, tmC, tmR, tmW2, tmC2;
  if (!encode_split_map(&tmA, a.A, a.K, a.M, a.lda, BM)) return cudaErrorInvalidValue;
  tmA2 = tmA;
  if (!encode_a2(&tmA2, a, p)) return cudaErrorInvalidValue;
  if (!encode3(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a.Wp, KT, a.N, 2, (uint64_t)KT * 2, (uint64_t)a.N * KT * 2, 64, 128, 2))
    return cudaErrorInvalidValue;
  if (!encode_split_map(&tmC, a.C, a.N, a.M, a.ldc, BM)) return cudaErrorInvalidValue;
  tmR = tmC;
  if (a.res && !encode_split_map(&tmR, a.res, a.N, a.M, a.ldr, BM)) return cudaErrorInvalidValue;
  static const bool single1024 = [] { const char* e = getenv("TUBER_FUSE2_SINGLE"); return e && e[0] == '1'; }();
  // 512 -> 128 (the 512-channel stage): per 128-row block the single-CTA kernel pulls 1.34 MB through L2 -> shared memory (t2 re-read for each
  // of its four sub-tiles, all of W4 and W1') for 0.64 MB of HBM traffic, the pair form 0.96 MB.  Measured (profiles/r2_fused2p_layer2_experiment.json):
  // the first bottleneck of the stage (K = 128 + 256 with the shortcut: six k-blocks per sub-tile) 240 -> 219 us; the K = 128 ones 145 -> 170 us (two
  // k-blocks do not hide the single conv4 accumulator's epilogue -> MMA hand-over), so the pair form takes K >= 256 only.
  // TUBER_FUSE2P_L2 (read per call): 1 = every 512 -> 128 launch, 0 = none.
  const char* p2e = getenv("TUBER_FUSE2P_L2");
  const bool pair512 = a.N == 512 && N2 == 128 && (p2e && (p2e[0] == '0' || p2e[0] == '1') ? p2e[0] == '1' : KT >= 256);
  const bool pair_form = (a.N == 1024 && !single1024) || pair512;   // gemm_fused2p_kernel: each CTA of the pair loads half of the weight rows
  if (!encode3(&tmW2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, W2p, a.N, N2, 2, (uint64_t)a.N * 2, (uint64_t)N2 * a.N * 2, 64, pair_form ? N2 / 2 : N2, 2))
    return cudaErrorInvalidValue;
  if (!encode_f32_panel_map(&tmC2, C2, N2, a.M, ldc2)) return cudaErrorInvalidValue;
  const int m_tiles = ceil_div(a.M, BM);
  const int grid = m_tiles < device_num_sms() ? m_tiles : device_num_sms();
  if (pair_form) {
    const int m_pairs = ceil_div(a.M, 2 * BM), max_pairs = device_num_sms() / 2;
    const int pairs = m_pairs < max_pairs ? m_pairs : max_pairs;
    if (pair512)
      return launch_pdl(gemm_fused2p_kernel<512, 128>, dim3(2 * pairs), dim3(FUSED2_THREADS), CfgWide::SMEM_BYTES, st, tmA, tmA2, tmW, tmC, tmR, tmW2, tmC2, p, q);
    return launch_pdl(gemm_fused2p_kernel<1024, 256>, dim3(2 * pairs), dim3(FUSED2_THREADS), CfgWide::SMEM_BYTES, st, tmA, tmA2, tmW, tmC, tmR, tmW2, tmC2, p, q);
  }
  if (a.N == 1024) return launch_pdl(gemm_fused2_kernel<1024, 256>, dim3(grid), dim3(FUSED2_THREADS), CfgWide::SMEM_BYTES, st, tmA, tmA2, tmW, tmC, tmR, tmW2, tmC2, p, q);
  if (a.N == 512 && N2 == 256)                              // last bottleneck of the 512-channel stage -> first conv1 of the 1024-channel stage
    return launch_pdl(gemm_fused2_kernel<512, 256>, dim3(grid), dim3(FUSED2_THREADS), CfgWide::SMEM_BYTES, st, tmA, tmA2, tmW, tmC, tmR, tmW2, tmC2, p, q);
  if (a.N == 512) return launch_pdl(gemm_fused2_kernel<512, 128>, dim3(grid), dim3(FUSED2_THREADS), CfgWide::SMEM_BYTES, st, tmA, tmA2, tmW, tmC, tmR, tmW2, tmC2, p, q);
  static const bool kbw2 = [] { const char* e = getenv("TUBER_FUSE2_KBW4"); return !(e && e[0] == '1'); }();
  static const bool kbw2_all = [] { const char* e = getenv("TUBER_FUSE2_KBW2_ALL"); return e && e[0] == '1'; }();
  if (N2 == 64 && (KT > BK || kbw2_all) && kbw2)                          // more than one conv4 k-block: half-filled W1' stages instead of the serial order
    return launch_pdl(gemm_fused2_kernel<256, 64, 2>, dim3(grid), dim3(FUSED2_THREADS), CfgWide::SMEM_BYTES, st, tmA, tmA2, tmW, tmC, tmR, tmW2, tmC2, p, q);
  if (N2 == 64) return launch_pdl(gemm_fused2_kernel<256, 64>, dim3(grid), dim3(FUSED2_THREADS), CfgWide::SMEM_BYTES, st, tmA, tmA2, tmW, tmC, tmR, tmW2, tmC2, p, q);
  return launch_pdl(gemm_fused2_kernel<256, 128>, dim3(grid), dim3(FUSED2_THREADS), CfgWide::SMEM_BYTES, st, tmA, tmA2, tmW, tmC, tmR, tmW2, tmC2, p, q);
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
