// Stem of the CSN backbone on the tensor cores (sm_100a): Conv3d(3->64,(3,7,7),s(1,2,2),p(1,3,3)) + BN + ReLU +
// MaxPool3d((1,3,3),s(1,2,2),p(0,1,1))  (reference models/backbones/ir_CSN_152.py:109-122,176-179) as an
// implicit GEMM with tcgen05 and the max pool fused into the epilogue.
//
//   conv[pos, oc] = relu(scale[oc] * sum_k patch[pos, k] * w[oc, k] + shift[oc])
//
// GEMM view per tile: M = 128 consecutive output columns of one conv row (b, t, oh), N = 32 output channels (a CTA
// owns one half of the 64 channels so that its packed filter bank, 72 KB, stays resident in shared memory),
// K = 576 = 9 k-blocks x 8 groups x 8: k-block = input plane (c, kt), group = filter row kh (group 7 is all zero),
// and a group holds one zero tap followed by the 7 taps kw = 0..6 of that filter row.  The patch row of a group is simply 8
// consecutive pixels in[c, t+kt-1, 2*oh+kh-3, 2*ow-4 .. 2*ow+3], so the A operand is built by threads from a ring of
// input rows kept in shared memory:
//
//   ring      two planar arrays (bf16 hi, bf16 mid) of [9 (c,kt) planes][8 row slots][132 pixel pairs]; a new conv row
//             needs only two new input rows per plane (slot = input row & 7); zero padding is applied at staging.
//   builders  8 warps: stage the ring (global fp32 -> split), then write the 128 x 64 hi and mid A tiles of each
//             k-block straight into TENSOR MEMORY (tcgen05.st, thread = row; 4 x LDS.32 per plane and group, no
//             shuffling): the im2col expansion never touches shared memory.  A stage = the 3 k-blocks of one input
//             channel (192 TMEM columns), two stages.
//   MMA       1 elected lane: three bf16 passes (mid*hi + hi*mid + hi*hi) of tcgen05.mma M=128, N=32, K=16 with A
//             from tensor memory and the filter bank from shared memory, into two fp32 TMEM accumulators (K-step
//             parity), double buffered per tile.
//   epilogue  4 warps: tcgen05.ld, BN scale/shift + ReLU into a 3-row ring of conv rows in shared memory; every
//             second row the 3x3/s2 max pool of the last three rows is written in split-bf16 to [B*T, H2, W2, 64].
//             A unit of 16 conv rows recomputes the one conv row above it that its first pooled row needs.
//             (When W1 > 128 the conv rows go to HBM through TMA instead and a separate kernel pools.)
#include <cuda.h>
#include <stdio.h>

#include "kernels.h"

namespace stemtc {

constexpr int KB = 9;                       // k-blocks of 64 = input planes (c, kt)
constexpr int KTOT = KB * 64;
constexpr int NCH = 32;                     // output channels per CTA
constexpr int A_STAGES = 2;                 // A stages in TENSOR MEMORY: 3 k-blocks (one input channel) = 192 columns each
constexpr int KB_PER_STAGE = 3;
constexpr int STAGE_COLS = KB_PER_STAGE * 64;
constexpr int W_KB_BYTES = 2 * NCH * 128;   // 8 KB per k-block (hi 4 KB + mid 4 KB)
constexpr int W_BYTES = KB * W_KB_BYTES;    // 72 KB
constexpr int RING_PX = 264, RING_PAIRS = RING_PX / 2;
constexpr int RING_PLANE_BYTES = 9 * 8 * RING_PAIRS * 4;      // 38016: one of the two planar arrays (hi / mid)
constexpr int ROWBUF_BYTES = 128 * 128;     // one conv row: 128 positions x 32 fp32, 128B-swizzled
constexpr int OFF_W = 0;
constexpr int OFF_ROWS = OFF_W + W_BYTES;                     // 3 conv rows
constexpr int OFF_RING_HI = OFF_ROWS + 3 * ROWBUF_BYTES;
constexpr int OFF_RING_MID = OFF_RING_HI + RING_PLANE_BYTES;
constexpr int OFF_BAR = OFF_RING_MID + ((RING_PLANE_BYTES + 127) / 128) * 128;
constexpr int OFF_SS = OFF_BAR + 128;       // scale[32], shift[32] of this CTA's channels
constexpr int SMEM_BYTES = OFF_SS + 256;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(OFF_RING_HI % 128 == 0 && OFF_RING_MID % 4 == 0, "alignment");
constexpr int NUM_THREADS = 416;            // warp 0 MMA, warps 1-4 epilogue, warps 5-12 builders
constexpr int NBUILD = 256;
constexpr int NACC = 2;                     // independent TMEM accumulators per tile (K-step parity)
constexpr int A_COL0 = 2 * NACC * NCH;      // TMEM columns: [0, 128) accumulators (2 tile buffers), then the A stages
constexpr int TMEM_COLS = 512;              // 128 + 2 x 192
constexpr int ROWS_PER_UNIT = 32;

struct Params {
  const uint8_t* xu8; const float* lut;     // stem_tc2_kernel only: decoded frames (B,T,H,W,3) uint8 + value table [3][256] instead of x
  const float* x;                           // (B,3,T,H,W) fp32
  const float* scale; const float* shift;   // [64]
  void* pooled;                             // split [B*T*H2*W2, 64] (fused pool) or null
  int B, T, H, W, H1, W1, H2, W2;
  int fuse_pool;
  int ct_n;                                 // stem_tc2_kernel: column tiles per conv row (1 when W1 <= 128, see there)
  int vec4;                                 // stem_tc2_kernel: fp32 rows staged with 16-byte loads (W % 4 == 0, one column tile, x 16-byte aligned)
};

TB_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one lane of the (converged) warp; keeps the surrounding control flow warp-uniform so that descriptors and
// addresses stay in uniform registers instead of being moved there (R2UR) before every tcgen05.mma
TB_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
TB_DEVINL void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
TB_DEVINL void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
TB_DEVINL void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
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
TB_DEVINL void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
TB_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
TB_DEVINL void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
TB_DEVINL void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
TB_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
TB_DEVINL void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
TB_DEVINL void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
TB_DEVINL void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
TB_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A = 128 lanes (rows) x 8 columns, two bf16 per 32-bit column
TB_DEVINL void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// thread (lane L of the warp) -> TMEM lane (quadrant base + L), 16 consecutive columns
TB_DEVINL void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
TB_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
TB_DEVINL void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
TB_DEVINL uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
TB_DEVINL void sts64(uint32_t addr, uint2 v) { asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory"); }
TB_DEVINL void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// K-major, 128B-swizzle shared-memory matrix descriptor (same encoding as gemm_tc.cu)
TB_DEVINL uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
TB_DEVINL uint32_t swz(uint32_t base, int r, int j) { return base + (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4); }

TB_DEVINL uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
TB_DEVINL uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
stem_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut, Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar = sb + OFF_BAR;
  auto full_bar = [&](int s) { return bar + 8u * s; };                       // A stage written by the builders
  auto empty_bar = [&](int s) { return bar + 8u * (A_STAGES + s); };         // A stage consumed by the MMAs
  auto tfull_bar = [&](int s) { return bar + 8u * (2 * A_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar + 8u * (2 * A_STAGES + 2 + s); };
  const uint32_t w_bar = bar + 8u * (2 * A_STAGES + 4);
  const uint32_t tmem_slot = bar + 8u * (2 * A_STAGES + 5);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * (2 * A_STAGES + 5));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = blockIdx.x & 1;                                 // which 32 output channels
  const int owb_n = (p.W1 + 127) / 128;
  const int rb_n = (p.H1 + ROWS_PER_UNIT - 1) / ROWS_PER_UNIT;
  const int units = p.B * p.T * owb_n * rb_n;
  const int cta = blockIdx.x >> 1, ncta = gridDim.x >> 1;
  // conv rows of unit row-block rb: [first, r1); with the fused pool the row above the block is recomputed
  auto unit_rows = [&](int rb, int& first, int& r0, int& r1) {
    r0 = rb * ROWS_PER_UNIT;
    r1 = min(p.H1, r0 + ROWS_PER_UNIT);
    first = (p.fuse_pool && r0 > 0) ? r0 - 1 : r0;
  };

  if (threadIdx.x == 0) {
    if (sb & 1023u) __trap();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    if (!p.fuse_pool) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(full_bar(s), NBUILD / 32);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= filter bank load + MMA issuer (whole warp, one elected lane issues) =================
    if (lane == 0) {
      mbar_expect_tx(w_bar, W_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_3d(sb + OFF_W + kb * W_KB_BYTES, &tmW, w_bar, kb * 64, half * NCH, 0);
    }
    __syncwarp();
    mbar_wait(w_bar, 0);
    constexpr uint32_t idesc = make_idesc(128, NCH);
    const uint64_t w_desc0 = make_smem_desc(sb + OFF_W);
    int stage = 0, it = 0;
    uint32_t phase = 0;
    for (int u = cta; u < units; u += ncta) {
      int first, r0, r1;
      unit_rows(u % rb_n, first, r0, r1);
      for (int oh = first; oh < r1; ++oh, ++it) {
        const int as = it & 1;
        mbar_wait(tempty_bar(as), ((it >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * NACC * NCH);
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {                           // one A stage = the three planes of input channel c
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int q = 0; q < KB_PER_STAGE; ++q) {
              const int kb = c * KB_PER_STAGE + q;
              const uint32_t a_hi = tmem_base + (uint32_t)(A_COL0 + stage * STAGE_COLS + q * 64), a_mid = a_hi + 32;   // 8 columns per K=16 step
              const uint64_t w_hi = w_desc0 + (uint64_t)(kb * (W_KB_BYTES >> 4)), w_mid = w_hi + ((NCH * 128) >> 4);
              const uint32_t first_k = kb == 0 ? 0u : 1u;
#pragma unroll
              for (int k = 0; k < 4; ++k) {                    // even / odd K-steps accumulate separately
                const uint32_t d = tmem_d + (uint32_t)((k & 1) * NCH);
                umma_bf16_ts(d, a_mid + 8 * k, w_hi + 2 * k, idesc, k < 2 ? first_k : 1u);
                umma_bf16_ts(d, a_hi + 8 * k, w_mid + 2 * k, idesc, 1u);
                umma_bf16_ts(d, a_hi + 8 * k, w_hi + 2 * k, idesc, 1u);
              }
            }
            umma_commit(empty_bar(stage));
            if (c == 2) umma_commit(tfull_bar(as));
          }
          __syncwarp();
          if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp <= 4) {
    // ================= epilogue =================
    const int lg = warp & 3;
    const int m = lg * 32 + lane;                                  // conv column inside the tile
    const int et = threadIdx.x - 32;
    float* s_sc = reinterpret_cast<float*>(smem + OFF_SS);
    float* s_sh = s_sc + NCH;
    if (et < NCH) {
      s_sc[et] = __ldg(p.scale + half * NCH + et);
      s_sh[et] = __ldg(p.shift + half * NCH + et);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int pw = et >> 1, hq = et & 1;                           // pooled column, which 16 of the 32 channels
    int it = 0;
    for (int u = cta; u < units; u += ncta) {
      const int rb = u % rb_n, owb = (u / rb_n) % owb_n, bt = u / (rb_n * owb_n);
      int first, r0, r1;
      unit_rows(rb, first, r0, r1);
      for (int oh = first; oh < r1; ++oh, ++it) {
        const int as = it & 1;
        mbar_wait(tfull_bar(as), (it >> 1) & 1);
        tcgen05_fence_after();
        float acc[32];
        {
          const uint32_t t0 = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * NACC * NCH);
          uint32_t part[32];
          tmem_ld32(t0, part);
#pragma unroll
          for (int q = 0; q < 32; ++q) acc[q] = __uint_as_float(part[q]);
          tmem_ld32(t0 + NCH, part);
#pragma unroll
          for (int q = 0; q < 32; ++q) acc[q] += __uint_as_float(part[q]);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(as));
        const uint32_t rowb = sb + OFF_ROWS + (uint32_t)((p.fuse_pool ? (oh % 3) : 0) * ROWBUF_BYTES);
        if (!p.fuse_pool) {
          if (et == 0) bulk_wait_read0();                          // previous store has finished reading the panel
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 sc = *reinterpret_cast<const float4*>(s_sc + 4 * q), sh = *reinterpret_cast<const float4*>(s_sh + 4 * q);
          uint4 o;
          o.x = __float_as_uint(fmaxf(fmaf(acc[4 * q], sc.x, sh.x), 0.f));
          o.y = __float_as_uint(fmaxf(fmaf(acc[4 * q + 1], sc.y, sh.y), 0.f));
          o.z = __float_as_uint(fmaxf(fmaf(acc[4 * q + 2], sc.z, sh.z), 0.f));
          o.w = __float_as_uint(fmaxf(fmaf(acc[4 * q + 3], sc.w, sh.w), 0.f));
          sts128(swz(rowb, m, q), o);
        }
        if (!p.fuse_pool) {
          fence_proxy_async();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            tma_store_3d(&tmOut, rowb, half * NCH, owb * 128, bt * p.H1 + oh);
            bulk_commit();
          }
          continue;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");             // the conv row is complete in shared memory
        // pooled row ph (centre conv row 2*ph) is emitted by the unit that owns its centre, once row 2*ph+1 (or the
        // last conv row) is in: max over conv rows {2ph-1, 2ph, 2ph+1} x columns {2pw-1, 2pw, 2pw+1}, missing = skipped
        const bool last_row = oh == p.H1 - 1;
        if ((oh & 1) || last_row) {
          const int ph = (oh & 1) ? (oh - 1) >> 1 : oh >> 1;
          if (2 * ph >= r0 && pw < p.W2) {
            float4 mx[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) mx[e] = make_float4(0.f, 0.f, 0.f, 0.f);   // post-ReLU values are >= 0
#pragma unroll
            for (int dr = -1; dr <= 1; ++dr) {
              const int cr = 2 * ph + dr;
              if (cr < 0 || cr >= p.H1 || cr > oh) continue;
              const uint32_t rsrc = sb + OFF_ROWS + (uint32_t)((cr % 3) * ROWBUF_BYTES);
#pragma unroll
              for (int dc = -1; dc <= 1; ++dc) {
                const int cc = 2 * pw + dc;
                if (cc < 0 || cc >= p.W1) continue;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const uint4 v = lds128(swz(rsrc, cc, hq * 4 + e));
                  mx[e].x = fmaxf(mx[e].x, __uint_as_float(v.x)); mx[e].y = fmaxf(mx[e].y, __uint_as_float(v.y));
                  mx[e].z = fmaxf(mx[e].z, __uint_as_float(v.z)); mx[e].w = fmaxf(mx[e].w, __uint_as_float(v.w));
                }
              }
            }
            const long long vox = ((long long)bt * p.H2 + ph) * p.W2 + pw;
            __nv_bfloat16* hi = split_hi(p.pooled, vox, 64) + half * NCH + hq * 16;
#pragma unroll
            for (int e = 0; e < 4; ++e) store_split4(hi + 4 * e, hi + 64 + 4 * e, mx[e]);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");           // pool reads done before a row slot is rewritten
        }
      }
    }
    if (!p.fuse_pool && et == 0) bulk_wait_all();
  } else {
    // ================= ring staging + A-tile builders =================
    const int bt_ = threadIdx.x - 160;                             // 0..255
    const int m = (warp & 3) * 32 + lane;                          // tile row = TMEM lane this thread may write
    const int gh = (warp - 5) >> 2;                                // which 4 of the 8 groups (filter rows) of a k-block
    const uint32_t ring_hi = sb + OFF_RING_HI, ring_mid = sb + OFF_RING_MID;
    for (int i = bt_; i < RING_PLANE_BYTES / 4; i += NBUILD) {      // never feed stale NaN bits to the MMA
      sts32(ring_hi + 4u * i, 0u);
      sts32(ring_mid + 4u * i, 0u);
    }
    const uint32_t a_dst = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(A_COL0 + gh * 16);
    int stage = 0;
    uint32_t phase = 0;
    for (int u = cta; u < units; u += ncta) {
      const int rb = u % rb_n, owb = (u / rb_n) % owb_n, bt = u / (rb_n * owb_n);
      const int b = bt / p.T, t = bt % p.T;
      int first, r0, r1;
      unit_rows(rb, first, r0, r1);
      const int iw0 = 2 * owb * 128 - 4;                            // tap 0 is the zero tap (see stem_pack_weight_kernel): pairs start at even columns
      // ring pixel (plane, ih, j) <- x[b, c, t+kt-1, ih, iw0+j], 0 outside the clip (the conv's zero padding).
      // Staging map: thread = (pixel pair spr in [0,128), row parity srr); the 4 pairs 128..131 of each row are a
      // second, mostly idle step.  Row pointers are warp-uniform, column validity is a per-thread constant.
      const int spr = bt_ & 127, srr = bt_ >> 7;
      const int xpr = 128 + (bt_ & 3), xrow = bt_ >> 2;             // extra step: row xrow of the 18 (or 63) staged rows
      auto col_ok = [&](int j) { return iw0 + j >= 0 && iw0 + j < p.W; };
      const bool ok0 = col_ok(2 * spr), ok1 = col_ok(2 * spr + 1), xok0 = col_ok(2 * xpr), xok1 = col_ok(2 * xpr + 1);
      const float* xb = p.x + (long long)b * 3 * p.T * p.H * p.W + iw0;
      const long long plane_stride = (long long)p.H * p.W;
      auto row_ptr = [&](int plane, int ih) -> const float* {
        const int c = plane / 3, f = t + plane % 3 - 1;
        if (f < 0 || f >= p.T || ih < 0 || ih >= p.H) return nullptr;
        return xb + ((long long)c * p.T + f) * plane_stride + (long long)ih * p.W;
      };
      auto store_pair = [&](int plane, int ih, int pr, float v0, float v1) {
        __nv_bfloat16 h0, m0, h1, m1;
        split_bf16(v0, h0, m0);
        split_bf16(v1, h1, m1);
        const uint32_t off = (uint32_t)(((plane * 8 + (ih & 7)) * RING_PAIRS + pr) * 4);
        sts32(ring_hi + off, pack_bf16x2(h0, h1));
        sts32(ring_mid + off, pack_bf16x2(m0, m1));
      };
      asm volatile("bar.sync 2, 256;" ::: "memory");               // previous unit's readers are done
      for (int plane = 0; plane < 9; ++plane) {                     // rows 2*first-3 .. 2*first+3 of every plane
        float v0[4], v1[4];
#pragma unroll
        for (int g2 = 0; g2 < 4; ++g2) {
          const int rr = 2 * g2 + srr;
          const float* rp = rr < 7 ? row_ptr(plane, 2 * first - 3 + rr) : nullptr;
          v0[g2] = (rp && ok0) ? __ldg(rp + 2 * spr) : 0.f;
          v1[g2] = (rp && ok1) ? __ldg(rp + 2 * spr + 1) : 0.f;
        }
#pragma unroll
        for (int g2 = 0; g2 < 4; ++g2)
          if (2 * g2 + srr < 7) store_pair(plane, 2 * first - 3 + 2 * g2 + srr, spr, v0[g2], v1[g2]);
      }
      if (xrow < 63) {
        const int plane = xrow / 7, ih = 2 * first - 3 + xrow % 7;
        const float* rp = row_ptr(plane, ih);
        store_pair(plane, ih, xpr, (rp && xok0) ? __ldg(rp + 2 * xpr) : 0.f, (rp && xok1) ? __ldg(rp + 2 * xpr + 1) : 0.f);
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      for (int oh = first; oh < r1; ++oh) {
        // prefetch the two input rows the next conv row adds (ih = 2*oh+4, 2*oh+5) into registers
        float pf0[9], pf1[9], px0 = 0.f, px1 = 0.f;
        const bool more = oh + 1 < r1;
        if (more) {
#pragma unroll
          for (int plane = 0; plane < 9; ++plane) {
            const float* rp = row_ptr(plane, 2 * oh + 4 + srr);
            pf0[plane] = (rp && ok0) ? __ldg(rp + 2 * spr) : 0.f;
            pf1[plane] = (rp && ok1) ? __ldg(rp + 2 * spr + 1) : 0.f;
          }
          if (xrow < 18) {
            const float* rp = row_ptr(xrow >> 1, 2 * oh + 4 + (xrow & 1));
            px0 = (rp && xok0) ? __ldg(rp + 2 * xpr) : 0.f;
            px1 = (rp && xok1) ? __ldg(rp + 2 * xpr + 1) : 0.f;
          }
        }
        uint32_t slot_off[4];                                      // ring row of filter row kh = 4*gh + jj (kh = 7: zero weights)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) slot_off[jj] = (uint32_t)((((2 * oh - 3 + gh * 4 + jj) & 7) * RING_PAIRS + m) * 4);
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          tcgen05_fence_after();
#pragma unroll
          for (int q = 0; q < KB_PER_STAGE; ++q) {
            const uint32_t pl = (uint32_t)((c * KB_PER_STAGE + q) * 8 * RING_PAIRS * 4);
            uint32_t hi[16], mid[16];                              // group jj -> columns 4*jj .. 4*jj+3 (two bf16 per column)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                hi[4 * jj + e] = lds32(ring_hi + pl + slot_off[jj] + 4u * e);
                mid[4 * jj + e] = lds32(ring_mid + pl + slot_off[jj] + 4u * e);
              }
            }
            tmem_st16(a_dst + (uint32_t)(stage * STAGE_COLS + q * 64), hi);
            tmem_st16(a_dst + (uint32_t)(stage * STAGE_COLS + q * 64 + 32), mid);
          }
          tmem_st_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(full_bar(stage));
          if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");             // every builder has finished reading this row's window
        if (more) {
#pragma unroll
          for (int plane = 0; plane < 9; ++plane) store_pair(plane, 2 * oh + 4 + srr, spr, pf0[plane], pf1[plane]);
          if (xrow < 18) store_pair(xrow >> 1, 2 * oh + 4 + (xrow & 1), xpr, px0, px1);
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// =============================================================================================================
// Fused-pool variant on CTA PAIRS (tcgen05 cta_group::2): the kernel above lets both CTAs of an SM pair build the
// same A tile (each multiplies it with its own 32 of the 64 filters), and it is bound by exactly that build (shared
// memory reads + tcgen05.st; ncu: tensor pipe 27 % busy, LSU/L1 68 %).  Here the two CTAs of a cluster form one
// M = 256 x N = 64 MMA: each CTA builds the A rows of ITS OWN conv row (128 positions) in its own tensor memory and
// keeps its own half of the filter bank (32 channels, 72 KB) in shared memory; the leader CTA's elected lane issues
// tcgen05.mma.cta_group::2, which reads A from both CTAs' tensor memory and B from both CTAs' shared memory and leaves
// D[128 positions x 64 channels] in each CTA's tensor memory.  Every A element is now built once per 64 channels.
//
//   unit      (frame bt, block of 32 conv rows); the pair works on two units at a time (rank r takes unit 2i + r).
//             A unit always walks 33 rows (its first pooled row needs the conv row above the block); rows outside the
//             image are dummies: built from zero padding, ignored by the epilogue.
//   barriers  full[s]   builders of BOTH CTAs -> leader's MMA warp (remote mbarrier.arrive for rank 1)
//             empty[s]  tcgen05.commit (multicast to both CTAs) -> each CTA's builders
//             tfull[a]  tcgen05.commit (multicast)              -> each CTA's epilogue
//             tempty[a] epilogue warps of BOTH CTAs -> leader's MMA warp
//   epilogue  tcgen05.ld 64 channels, BN + ReLU into ONE conv-row buffer in shared memory, then the 3x3/s2 max pool as
//             a running maximum held in registers (thread = pooled column x 32 channels): horizontal 3-max of the
//             row from shared memory, vertical combination across rows {2ph-1, 2ph, 2ph+1} in registers.
// =============================================================================================================
namespace p2 {
constexpr int NCH2 = 64;                    // channels per pair (N of the MMA); NCH = 32 per CTA (B operand half)
constexpr int ROW2_BYTES = 128 * NCH2 * 4;  // conv row buffer: 128 positions x 64 fp32
constexpr int OFF_W2 = 0;
constexpr int OFF_ROW2 = OFF_W2 + W_BYTES;
constexpr int OFF_RING_HI2 = OFF_ROW2 + ROW2_BYTES;
constexpr int OFF_RING_MID2 = OFF_RING_HI2 + RING_PLANE_BYTES;
constexpr int OFF_BAR2 = OFF_RING_MID2 + ((RING_PLANE_BYTES + 127) / 128) * 128;
constexpr int OFF_SS2 = OFF_BAR2 + 128;     // scale[64], shift[64]
constexpr int SMEM2_BYTES = OFF_SS2 + 512;
static_assert(SMEM2_BYTES <= 232448, "shared memory budget");
static_assert(OFF_RING_HI2 % 128 == 0, "alignment");
constexpr int A_COL0_2 = 2 * NCH2;          // TMEM columns: [0, 128) two accumulator buffers, then the two A stages
constexpr int UNIT_ROWS = ROWS_PER_UNIT + 1;
}  // namespace p2

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
TB_DEVINL void mbar_wait_cluster(uint32_t bar, uint32_t parity) {          // waits on a barrier that remote CTAs arrive on
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
TB_DEVINL void umma2_commit_mc(uint32_t bar) {                             // arrives on `bar` of both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
TB_DEVINL void umma2_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

// U8: the staging threads read decoded uint8 frames through the value table instead of the fp32 clip (a separate instantiation:
// the fp32 kernel keeps its register allocation)
template <bool U8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
stem_tc2_kernel(const __grid_constant__ CUtensorMap tmW, Params p) {
  using namespace p2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar = sb + OFF_BAR2;
  auto full_bar = [&](int s) { return bar + 8u * s; };                       // leader: A stage written by both CTAs' builders
  auto empty_bar = [&](int s) { return bar + 8u * (A_STAGES + s); };         // each CTA: A stage consumed by the MMAs
  auto tfull_bar = [&](int s) { return bar + 8u * (2 * A_STAGES + s); };     // each CTA: accumulator complete
  auto tempty_bar = [&](int s) { return bar + 8u * (2 * A_STAGES + 2 + s); };  // leader: accumulator drained by both epilogues
  const uint32_t w_bar = bar + 8u * (2 * A_STAGES + 4);
  const uint32_t tmem_slot = bar + 8u * (2 * A_STAGES + 5);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR2 + 8 * (2 * A_STAGES + 5));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                         // = which 32 output channels this CTA's filter half holds
  const int rb_n = (p.H1 + ROWS_PER_UNIT - 1) / ROWS_PER_UNIT;
  // Conv rows wider than the 128 TMEM lanes (W1 = 171 / 228 at the 256 x 341 / 256 x 455 clips the reference's evaluation transform
  // produces) are cut into column tiles: tile ct computes the 128 conv columns from c0 = 126 ct - 1 and emits the 63 pooled columns
  // 63 ct .. 63 ct + 62, whose 3-wide windows (conv columns 2 pw - 1 .. 2 pw + 1) lie inside it -- a one-column halo on the left
  // instead of a second pass over the conv rows.  A single tile (W1 <= 128) keeps c0 = 0 and 64 pooled columns.
  const int ct_n = p.ct_n > 1 ? p.ct_n : 1;
  const int units = p.B * p.T * rb_n * ct_n;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int steps = (units + 1) / 2;                               // pair steps; step i = units 2i (rank 0) and 2i+1 (rank 1)

  if (threadIdx.x == 0) {
    if (sb & 1023u) __trap();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(full_bar(s), 2 * NBUILD / 32);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8);
    }
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                                              // both CTAs' barriers exist before anyone arrives remotely
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= filter half load (both CTAs) + MMA issuer (leader CTA only) =================
    if (lane == 0) {
      mbar_expect_tx(w_bar, W_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_3d(sb + OFF_W2 + kb * W_KB_BYTES, &tmW, w_bar, kb * 64, (int)rank * NCH, 0);
    }
    __syncwarp();
    if (rank == 0) {
      mbar_wait(w_bar, 0);                                         // (rank 1's half: its builders wait for it before their first arrive)
      constexpr uint32_t idesc = make_idesc(256, NCH2);
      const uint64_t w_desc0 = make_smem_desc(sb + OFF_W2);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int i = pair; i < steps; i += npairs) {
        for (int r = 0; r < UNIT_ROWS; ++r, ++it) {
          const int as = it & 1;
          mbar_wait_cluster(tempty_bar(as), ((it >> 1) & 1) ^ 1);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(as * NCH2);
#pragma unroll 1
          for (int c = 0; c < 3; ++c) {
            mbar_wait_cluster(full_bar(stage), phase);
            tcgen05_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int q = 0; q < KB_PER_STAGE; ++q) {
                const int kb = c * KB_PER_STAGE + q;
                const uint32_t a_hi = tmem_base + (uint32_t)(A_COL0_2 + stage * STAGE_COLS + q * 64), a_mid = a_hi + 32;
                const uint64_t w_hi = w_desc0 + (uint64_t)(kb * (W_KB_BYTES >> 4)), w_mid = w_hi + ((NCH * 128) >> 4);
                const uint32_t first_k = kb == 0 ? 0u : 1u;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma2_bf16_ts(tmem_d, a_mid + 8 * k, w_hi + 2 * k, idesc, k == 0 ? first_k : 1u);
                  umma2_bf16_ts(tmem_d, a_hi + 8 * k, w_mid + 2 * k, idesc, 1u);
                  umma2_bf16_ts(tmem_d, a_hi + 8 * k, w_hi + 2 * k, idesc, 1u);
                }
              }
              umma2_commit_mc(empty_bar(stage));
              if (c == 2) umma2_commit_mc(tfull_bar(as));
            }
            __syncwarp();
            if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp <= 4) {
    // ================= epilogue: BN + ReLU + 3x3/s2 max pool =================
    const int lg = warp & 3;
    const int m = lg * 32 + lane;                                  // conv column = TMEM lane
    const int et = threadIdx.x - 32;
    float* s_sc = reinterpret_cast<float*>(smem + OFF_SS2);
    float* s_sh = s_sc + NCH2;
    if (et < NCH2) {
      s_sc[et] = __ldg(p.scale + et);
      s_sh[et] = __ldg(p.shift + et);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int pw = et >> 1, hq = et & 1;                           // pooled column; hq: the even / odd 4-channel chunks (32 channels)
    const uint32_t rowb = sb + OFF_ROW2;
    // 16-byte chunk `chunk` (4 channels) of position `pos`: rows are 256 B apart (bank-aligned), so the low 3 chunk bits are XORed
    // with pos & 7 -- conflict-free both for the row write (8 consecutive positions, same chunk, per 128-bit phase) and for the
    // pool read (4 consecutive pooled columns = positions 2 apart, x the thread pair hq = 0/1 that owns even / odd chunks)
    auto rswz = [&](int pos, int chunk) { return rowb + (uint32_t)pos * 256u + (uint32_t)((((chunk ^ pos) & 7) | (chunk & 8)) << 4); };
    const uint32_t tempty_remote0 = mapa_rank(tempty_bar(0), 0), tempty_remote1 = mapa_rank(tempty_bar(1), 0);
    int it = 0;
    for (int i = pair; i < steps; i += npairs) {
      const int u = min(2 * i + (int)rank, units - 1);             // odd unit count: the last step's rank 1 repeats the last unit
      const int ct = u % ct_n, ur = u / ct_n;
      const int rb = ur % rb_n, bt = ur / rb_n;
      const int r0 = rb * ROWS_PER_UNIT;
      const int c0 = ct_n > 1 ? 126 * ct - 1 : 0;                  // first conv column of this tile (TMEM lane 0)
      const int pwg = (ct_n > 1 ? 63 * ct : 0) + pw;               // this thread's pooled column in the image
      const bool pw_ok = pw < (ct_n > 1 ? 63 : 64) && pwg < p.W2;
      float4 run[8];                                               // running maximum over the rows seen so far (post-ReLU: >= 0)
#pragma unroll
      for (int e = 0; e < 8; ++e) run[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < UNIT_ROWS; ++r, ++it) {
        const int oh = r0 - 1 + r;
        const int as = it & 1;
        mbar_wait(tfull_bar(as), (it >> 1) & 1);
        tcgen05_fence_after();
        const bool valid = oh >= 0 && oh < p.H1;                   // uniform over the CTA
        if (valid) {
          const uint32_t t0 = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * NCH2);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t part[32];
            tmem_ld32(t0 + 32 * hh, part);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 sc = *reinterpret_cast<const float4*>(s_sc + 32 * hh + 4 * q), sh = *reinterpret_cast<const float4*>(s_sh + 32 * hh + 4 * q);
              uint4 o;
              o.x = __float_as_uint(fmaxf(fmaf(__uint_as_float(part[4 * q]), sc.x, sh.x), 0.f));
              o.y = __float_as_uint(fmaxf(fmaf(__uint_as_float(part[4 * q + 1]), sc.y, sh.y), 0.f));
              o.z = __float_as_uint(fmaxf(fmaf(__uint_as_float(part[4 * q + 2]), sc.z, sh.z), 0.f));
              o.w = __float_as_uint(fmaxf(fmaf(__uint_as_float(part[4 * q + 3]), sc.w, sh.w), 0.f));
              sts128(rswz(m, 8 * hh + q), o);
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(as ? tempty_remote1 : tempty_remote0);
        if (!valid) continue;                                      // dummy row (above / below the image): contributes nothing
        asm volatile("bar.sync 1, 128;" ::: "memory");             // the conv row is complete in shared memory
        // horizontal 3-max at stride 2 of this conv row: columns 2pw-1, 2pw, 2pw+1
        float4 hm[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) hm[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pw_ok) {
#pragma unroll
          for (int dc = -1; dc <= 1; ++dc) {
            const int cc = 2 * pwg + dc;
            if (cc < 0 || cc >= p.W1) continue;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const uint4 v = lds128(rswz(cc - c0, 2 * e + hq));
              hm[e].x = fmaxf(hm[e].x, __uint_as_float(v.x)); hm[e].y = fmaxf(hm[e].y, __uint_as_float(v.y));
              hm[e].z = fmaxf(hm[e].z, __uint_as_float(v.z)); hm[e].w = fmaxf(hm[e].w, __uint_as_float(v.w));
            }
          }
        }
        // vertical: pooled row ph = max(rows 2ph-1, 2ph, 2ph+1); emitted at row 2ph+1, or at row 2ph when it is the last one
        const bool odd = oh & 1;
        const bool emit = r > 0 && (odd || oh == p.H1 - 1);
        if (!odd || r > 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            run[e].x = fmaxf(run[e].x, hm[e].x); run[e].y = fmaxf(run[e].y, hm[e].y);
            run[e].z = fmaxf(run[e].z, hm[e].z); run[e].w = fmaxf(run[e].w, hm[e].w);
          }
        }
        if (emit && pw_ok) {
          const int ph = oh >> 1;
          const long long vox = ((long long)bt * p.H2 + ph) * p.W2 + pwg;
          __nv_bfloat16* hi = split_hi(p.pooled, vox, 64) + hq * 4;
#pragma unroll
          for (int e = 0; e < 8; ++e) store_split4(hi + 8 * e, hi + 64 + 8 * e, run[e]);
        }
        if (odd) {                                                 // this row is row 2(ph+1)-1 of the next pooled row
#pragma unroll
          for (int e = 0; e < 8; ++e) run[e] = hm[e];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");             // pool reads done before the row buffer is rewritten
      }
    }
  } else {
    // ================= ring staging + A-tile builders (as in stem_tc_kernel; one unit per CTA) =================
    const int bt_ = threadIdx.x - 160;                             // 0..255
    const int m = (warp & 3) * 32 + lane;
    const int gh = (warp - 5) >> 2;
    const uint32_t ring_hi = sb + OFF_RING_HI2, ring_mid = sb + OFF_RING_MID2;
    for (int i = bt_; i < RING_PLANE_BYTES / 4; i += NBUILD) {
      sts32(ring_hi + 4u * i, 0u);
      sts32(ring_mid + 4u * i, 0u);
    }
    const uint32_t a_dst = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(A_COL0_2 + gh * 16);
    const uint32_t full_remote0 = mapa_rank(full_bar(0), 0), full_remote1 = mapa_rank(full_bar(1), 0);
    mbar_wait(w_bar, 0);                                           // this CTA's filter half has landed before its first "full" arrive
    int stage = 0;
    uint32_t phase = 0;
    for (int i = pair; i < steps; i += npairs) {
      const int u = min(2 * i + (int)rank, units - 1);
      const int ct = u % ct_n, ur = u / ct_n;
      const int rb = ur % rb_n, bt = ur / rb_n;
      const int b = bt / p.T, t = bt % p.T;
      const int first = rb * ROWS_PER_UNIT - 1, r1 = first + UNIT_ROWS;
      const int iw0 = 2 * (ct_n > 1 ? 126 * ct - 1 : 0) - 4;        // input column of tap 0 (the zero tap) of the tile's first conv column: even
      const int spr = bt_ & 127, srr = bt_ >> 7;
      const int xpr = 128 + (bt_ & 3), xrow = bt_ >> 2;
      auto col_ok = [&](int j) { return iw0 + j >= 0 && iw0 + j < p.W; };
      const bool ok0 = col_ok(2 * spr), ok1 = col_ok(2 * spr + 1), xok0 = col_ok(2 * xpr), xok1 = col_ok(2 * xpr + 1);
      const float* xb = p.x + (long long)b * 3 * p.T * p.H * p.W + iw0;
      const long long plane_stride = (long long)p.H * p.W;
      // input value at (plane, row ih, column iw0 + j): the fp32 clip, or -- uint8 path -- the decoded frame's byte through the value
      // table of the reference's ToTensor + Normalize (video_transforms.py:294-296,308-314; the table holds exactly the floats
      // normalize_u8_kernel would have written, so both paths feed the tensor core identical bits).  Zero outside the volume.
      auto in_val = [&](int plane, int ih, int j, bool ok) -> float {
        const int c = plane / 3, f = t + plane % 3 - 1;
        if (!ok || f < 0 || f >= p.T || ih < 0 || ih >= p.H) return 0.f;
        if constexpr (U8) {
          const uint8_t v = __ldg(p.xu8 + ((((long long)b * p.T + f) * p.H + ih) * p.W + (iw0 + j)) * 3 + c);
          return __ldg(p.lut + c * 256 + v);
        } else {
          return __ldg(xb + ((long long)c * p.T + f) * plane_stride + (long long)ih * p.W + j);
        }
      };
      auto store_pair = [&](int plane, int ih, int pr, float v0, float v1) {
        __nv_bfloat16 h0, m0, h1, m1;
        split_bf16(v0, h0, m0);
        split_bf16(v1, h1, m1);
        const uint32_t off = (uint32_t)(((plane * 8 + (ih & 7)) * RING_PAIRS + pr) * 4);
        sts32(ring_hi + off, pack_bf16x2(h0, h1));
        sts32(ring_mid + off, pack_bf16x2(m0, m1));
      };
      // Vector form of the staging (fp32 clip, W % 4 == 0, one column tile): a thread moves QUADS -- four pixels = two ring pairs =
      // one aligned 16-byte global load, two packed hi / mid splits, one 8-byte store into each ring -- instead of single pixels:
      // quad q = columns 4q - 4 .. 4q - 1 of a row, 64 quads + 2 border quads per row; thread = (quad bt_ & 63, row selector bt_ >> 6).
      const bool vec = !U8 && p.vec4 != 0;
      const int vq = bt_ & 63, vsel = bt_ >> 6;
      const bool vq_ok = vq >= 1 && 4 * vq <= p.W;                  // columns 4 vq - 4 .. 4 vq - 1 inside the row
      const int xq = 64 + (bt_ & 1);
      const bool xq_ok = 4 * xq <= p.W;
      auto in_quad = [&](int plane, int ih, int q, bool ok) -> float4 {
        const int c = plane / 3, f = t + plane % 3 - 1;
        if (!ok || f < 0 || f >= p.T || ih < 0 || ih >= p.H) return make_float4(0.f, 0.f, 0.f, 0.f);
        return __ldg(reinterpret_cast<const float4*>(xb + ((long long)c * p.T + f) * plane_stride + (long long)ih * p.W + 4 * q));
      };
      auto store_quad = [&](int plane, int ih, int q, const float4& v) {
        uint2 h, md;
        split_bf16x2(v.x, v.y, h.x, md.x);
        split_bf16x2(v.z, v.w, h.y, md.y);
        const uint32_t off = (uint32_t)(((plane * 8 + (ih & 7)) * RING_PAIRS + 2 * q) * 4);
        sts64(ring_hi + off, h);
        sts64(ring_mid + off, md);
      };
      asm volatile("bar.sync 2, 256;" ::: "memory");               // previous unit's readers are done
      if (vec) {                                                     // rows 2*first-3 .. 2*first+3 of every plane: 63 (plane, row) combinations
        for (int k = 0; k < 16; ++k) {
          const int combo = 4 * k + vsel;
          if (combo < 63) {
            const int plane = combo / 7, ih = 2 * first - 3 + combo % 7;
            store_quad(plane, ih, vq, in_quad(plane, ih, vq, vq_ok));
          }
        }
        if (bt_ < 126) {
          const int combo = bt_ >> 1, plane = combo / 7, ih = 2 * first - 3 + combo % 7;
          store_quad(plane, ih, xq, in_quad(plane, ih, xq, xq_ok));
        }
      } else {
      for (int plane = 0; plane < 9; ++plane) {                     // rows 2*first-3 .. 2*first+3 of every plane
        float v0[4], v1[4];
#pragma unroll
        for (int g2 = 0; g2 < 4; ++g2) {
          const int rr = 2 * g2 + srr;
          v0[g2] = in_val(plane, 2 * first - 3 + rr, 2 * spr, rr < 7 && ok0);
          v1[g2] = in_val(plane, 2 * first - 3 + rr, 2 * spr + 1, rr < 7 && ok1);
        }
#pragma unroll
        for (int g2 = 0; g2 < 4; ++g2)
          if (2 * g2 + srr < 7) store_pair(plane, 2 * first - 3 + 2 * g2 + srr, spr, v0[g2], v1[g2]);
      }
      if (xrow < 63) {
        const int plane = xrow / 7, ih = 2 * first - 3 + xrow % 7;
        store_pair(plane, ih, xpr, in_val(plane, ih, 2 * xpr, xok0), in_val(plane, ih, 2 * xpr + 1, xok1));
      }
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      for (int oh = first; oh < r1; ++oh) {
        float pf0[9], pf1[9], px0 = 0.f, px1 = 0.f;
        float4 pv[5], pxv = make_float4(0.f, 0.f, 0.f, 0.f);       // vector form: the 18 (plane, row) combinations of rows 2 oh + 4, 2 oh + 5
        const bool more = oh + 1 < r1;
        if (more && vec) {
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const int combo = 4 * k + vsel;
            pv[k] = in_quad(combo >> 1, 2 * oh + 4 + (combo & 1), vq, vq_ok && combo < 18);
          }
          if (bt_ < 36) pxv = in_quad(bt_ >> 2, 2 * oh + 4 + ((bt_ >> 1) & 1), xq, xq_ok);
        } else if (more) {
#pragma unroll
          for (int plane = 0; plane < 9; ++plane) {
            pf0[plane] = in_val(plane, 2 * oh + 4 + srr, 2 * spr, ok0);
            pf1[plane] = in_val(plane, 2 * oh + 4 + srr, 2 * spr + 1, ok1);
          }
          if (xrow < 18) {
            px0 = in_val(xrow >> 1, 2 * oh + 4 + (xrow & 1), 2 * xpr, xok0);
            px1 = in_val(xrow >> 1, 2 * oh + 4 + (xrow & 1), 2 * xpr + 1, xok1);
          }
        }
        uint32_t slot_off[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) slot_off[jj] = (uint32_t)((((2 * oh - 3 + gh * 4 + jj) & 7) * RING_PAIRS + m) * 4);
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          tcgen05_fence_after();
#pragma unroll
          for (int q = 0; q < KB_PER_STAGE; ++q) {
            const uint32_t pl = (uint32_t)((c * KB_PER_STAGE + q) * 8 * RING_PAIRS * 4);
            uint32_t hi[16], mid[16];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                hi[4 * jj + e] = lds32(ring_hi + pl + slot_off[jj] + 4u * e);
                mid[4 * jj + e] = lds32(ring_mid + pl + slot_off[jj] + 4u * e);
              }
            }
            tmem_st16(a_dst + (uint32_t)(stage * STAGE_COLS + q * 64), hi);
            tmem_st16(a_dst + (uint32_t)(stage * STAGE_COLS + q * 64 + 32), mid);
          }
          tmem_st_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(stage ? full_remote1 : full_remote0);
          if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (more && vec) {
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const int combo = 4 * k + vsel;
            if (combo < 18) store_quad(combo >> 1, 2 * oh + 4 + (combo & 1), vq, pv[k]);
          }
          if (bt_ < 36) store_quad(bt_ >> 2, 2 * oh + 4 + ((bt_ >> 1) & 1), xq, pxv);
        } else if (more) {
#pragma unroll
          for (int plane = 0; plane < 9; ++plane) store_pair(plane, 2 * oh + 4 + srr, spr, pf0[plane], pf1[plane]);
          if (xrow < 18) store_pair(xrow >> 1, 2 * oh + 4 + (xrow & 1), xpr, px0, px1);
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                                              // the leader's MMAs read this CTA's memories until its last commit
  if (warp == 0) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// =============================================================================================================
// stem_tc3_kernel: the pair kernel above with TWO conv rows per accumulator batch, sharing their A K-steps.
//
// A K-step (16 elements = 8 TMEM columns) of plane (c, kt) holds two filter-row groups = 8 pixels of each of TWO consecutive input
// rows (y, y + 1), y = 2 oh - 3 + 2 k for K-step k of conv row oh.  Conv row oh + 1 reads input rows two further down, so ITS K-step
// k is conv row oh's K-step k + 1: the rows (oh, oh + 1) together need 5 distinct K-steps per plane, not 8.  The builders therefore
// write ONE plane per stage -- 5 K-steps, [hi 40 columns | mid 40 columns] -- and the MMA warp multiplies K-steps r .. r + 3 of it with
// the plane's four weight K-steps into the accumulator of row r (two independent accumulators, issued alternately).  The stem is
// bound by the A build (shared-memory reads + tcgen05.st; tensor pipe 42 % busy in stem_tc2_kernel): this cuts the build work per
// conv row to 5/8 with the same MMAs.  Tensor memory: 2 accumulator sets x 2 rows x 64 columns + 3 stages x 80 columns = 496.
// The input ring keeps 12 row slots per plane (a batch reads rows 2 oh - 3 .. 2 oh + 6 and stages 2 oh + 6 .. 2 oh + 9 for the next).
// MEASURED (profiles/r2_stem3_experiment.json): correct, and NOT faster than stem_tc2_kernel (0.99 - 1.02 ms against 0.95 - 0.99 ms
// at 8 clips 32 x 256 x 256): knocking out one role at a time shows the A build is only 0.17 ms of the kernel; the skeleton of
// cluster-wide stage handshakes alone (every role's work removed) takes 0.37 ms.  Kept behind TUBER_STEM3=1 with its tests.
// =============================================================================================================
namespace p3 {
constexpr int RB = 2;                        // conv rows per batch
constexpr int KSTEPS = RB + 3;               // distinct A K-steps of a plane per batch
constexpr int HALF_COLS = KSTEPS * 8;        // 40 TMEM columns per precision part
constexpr int STAGE3_COLS = 2 * HALF_COLS;   // 80: [hi | mid] of one plane
constexpr int A3_STAGES = 3;
constexpr int ACC3_COLS = RB * 64;           // one accumulator set
constexpr int A_COL0_3 = 2 * ACC3_COLS;      // 256
static_assert(A_COL0_3 + A3_STAGES * STAGE3_COLS <= TMEM_COLS, "tensor memory budget");
constexpr int RING3_SLOTS = 12;
constexpr int RING3_PLANE_BYTES = 9 * RING3_SLOTS * RING_PAIRS * 4;   // 57024 per precision part
constexpr int OFF_W3 = 0;
constexpr int OFF_ROW3 = OFF_W3 + W_BYTES;
constexpr int OFF_RING_HI3 = OFF_ROW3 + p2::ROW2_BYTES;
constexpr int OFF_RING_MID3 = OFF_RING_HI3 + RING3_PLANE_BYTES;
constexpr int OFF_BAR3 = OFF_RING_MID3 + ((RING3_PLANE_BYTES + 127) / 128) * 128;
constexpr int OFF_SS3 = OFF_BAR3 + 128;
constexpr int SMEM3_BYTES = OFF_SS3 + 512;
static_assert(SMEM3_BYTES <= 232448, "shared memory budget");
static_assert(OFF_RING_HI3 % 128 == 0, "alignment");
constexpr int NBATCH = (p2::UNIT_ROWS + RB - 1) / RB;   // 17 batches walk 34 rows: the last one is a dummy
}  // namespace p3

TB_DEVINL void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
TB_DEVINL void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
stem_tc3_kernel(const __grid_constant__ CUtensorMap tmW, Params p) {
  using namespace p2;
  using namespace p3;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar = sb + OFF_BAR3;
  auto full_bar = [&](int s) { return bar + 8u * s; };                       // leader: A stage written by both CTAs' builders
  auto empty_bar = [&](int s) { return bar + 8u * (A3_STAGES + s); };        // each CTA: A stage consumed by the MMAs
  auto tfull_bar = [&](int s) { return bar + 8u * (2 * A3_STAGES + s); };    // each CTA: accumulator set complete
  auto tempty_bar = [&](int s) { return bar + 8u * (2 * A3_STAGES + 2 + s); };  // leader: accumulator set drained by both epilogues
  const uint32_t w_bar = bar + 8u * (2 * A3_STAGES + 4);
  const uint32_t tmem_slot = bar + 8u * (2 * A3_STAGES + 5);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR3 + 8 * (2 * A3_STAGES + 5));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int rb_n = (p.H1 + ROWS_PER_UNIT - 1) / ROWS_PER_UNIT;
  const int ct_n = p.ct_n > 1 ? p.ct_n : 1;                        // column tiles of a conv row, as in stem_tc2_kernel
  const int units = p.B * p.T * rb_n * ct_n;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int steps = (units + 1) / 2;

  if (threadIdx.x == 0) {
    if (sb & 1023u) __trap();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    for (int s = 0; s < A3_STAGES; ++s) {
      mbar_init(full_bar(s), 2 * NBUILD / 32);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8);
    }
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= filter half load (both CTAs) + MMA issuer (leader CTA only) =================
    if (lane == 0) {
      mbar_expect_tx(w_bar, W_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_3d(sb + OFF_W3 + kb * W_KB_BYTES, &tmW, w_bar, kb * 64, (int)rank * NCH, 0);
    }
    __syncwarp();
    if (rank == 0) {
      mbar_wait(w_bar, 0);
      constexpr uint32_t idesc = make_idesc(256, NCH2);
      const uint64_t w_desc0 = make_smem_desc(sb + OFF_W3);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int i = pair; i < steps; i += npairs) {
        for (int bch = 0; bch < NBATCH; ++bch, ++it) {
          const int as = it & 1;
          mbar_wait_cluster(tempty_bar(as), ((it >> 1) & 1) ^ 1);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(as * ACC3_COLS);
#pragma unroll 1
          for (int plane = 0; plane < KB; ++plane) {
            mbar_wait_cluster(full_bar(stage), phase);
            tcgen05_fence_after();
            if (elect_one()) {
              const uint32_t a0 = tmem_base + (uint32_t)(A_COL0_3 + stage * STAGE3_COLS);
              const uint64_t w_hi = w_desc0 + (uint64_t)(plane * (W_KB_BYTES >> 4)), w_mid = w_hi + ((NCH * 128) >> 4);
              const uint32_t first_k = plane == 0 ? 0u : 1u;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int r = 0; r < RB; ++r) {                     // row r reads K-steps r .. r + 3 of the plane
                  const uint32_t a_hi = a0 + (uint32_t)(8 * (k + r)), a_mid = a_hi + HALF_COLS, d = tmem_d + (uint32_t)(r * NCH2);
                  umma2_bf16_ts(d, a_mid, w_hi + 2 * k, idesc, k == 0 ? first_k : 1u);
                  umma2_bf16_ts(d, a_hi, w_mid + 2 * k, idesc, 1u);
                  umma2_bf16_ts(d, a_hi, w_hi + 2 * k, idesc, 1u);
                }
              }
              umma2_commit_mc(empty_bar(stage));
              if (plane == KB - 1) umma2_commit_mc(tfull_bar(as));
            }
            __syncwarp();
            if (++stage == A3_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp <= 4) {
    // ================= epilogue: BN + ReLU + 3x3/s2 max pool, the batch's two rows one after the other =================
    const int lg = warp & 3;
    const int m = lg * 32 + lane;
    const int et = threadIdx.x - 32;
    float* s_sc = reinterpret_cast<float*>(smem + OFF_SS3);
    float* s_sh = s_sc + NCH2;
    if (et < NCH2) {
      s_sc[et] = __ldg(p.scale + et);
      s_sh[et] = __ldg(p.shift + et);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int pw = et >> 1, hq = et & 1;
    const uint32_t rowb = sb + OFF_ROW3;
    auto rswz = [&](int pos, int chunk) { return rowb + (uint32_t)pos * 256u + (uint32_t)((((chunk ^ pos) & 7) | (chunk & 8)) << 4); };
    const uint32_t tempty_remote0 = mapa_rank(tempty_bar(0), 0), tempty_remote1 = mapa_rank(tempty_bar(1), 0);
    int it = 0;
    for (int i = pair; i < steps; i += npairs) {
      const int u = min(2 * i + (int)rank, units - 1);
      const int ct = u % ct_n, ur = u / ct_n;
      const int rb = ur % rb_n, bt = ur / rb_n;
      const int r0 = rb * ROWS_PER_UNIT;
      const int c0 = ct_n > 1 ? 126 * ct - 1 : 0;
      const int pwg = (ct_n > 1 ? 63 * ct : 0) + pw;
      const bool pw_ok = pw < (ct_n > 1 ? 63 : 64) && pwg < p.W2;
      float4 run[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) run[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int bch = 0; bch < NBATCH; ++bch, ++it) {
        const int as = it & 1;
        mbar_wait(tfull_bar(as), (it >> 1) & 1);
        tcgen05_fence_after();
#pragma unroll 1
        for (int rr = 0; rr < RB; ++rr) {
          const int r = RB * bch + rr;
          const int oh = r0 - 1 + r;
          const bool valid = r < UNIT_ROWS && oh >= 0 && oh < p.H1;  // uniform over the CTA
          if (valid) {
            const uint32_t t0 = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * ACC3_COLS + rr * NCH2);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t part[32];
              tmem_ld32(t0 + 32 * hh, part);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 sc = *reinterpret_cast<const float4*>(s_sc + 32 * hh + 4 * q), sh = *reinterpret_cast<const float4*>(s_sh + 32 * hh + 4 * q);
                uint4 o;
                o.x = __float_as_uint(fmaxf(fmaf(__uint_as_float(part[4 * q]), sc.x, sh.x), 0.f));
                o.y = __float_as_uint(fmaxf(fmaf(__uint_as_float(part[4 * q + 1]), sc.y, sh.y), 0.f));
                o.z = __float_as_uint(fmaxf(fmaf(__uint_as_float(part[4 * q + 2]), sc.z, sh.z), 0.f));
                o.w = __float_as_uint(fmaxf(fmaf(__uint_as_float(part[4 * q + 3]), sc.w, sh.w), 0.f));
                sts128(rswz(m, 8 * hh + q), o);
              }
            }
          }
          if (rr == RB - 1) {                                      // both rows of the set have left tensor memory
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(as ? tempty_remote1 : tempty_remote0);
          }
          if (!valid) continue;
          asm volatile("bar.sync 1, 128;" ::: "memory");
          float4 hm[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) hm[e] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (pw_ok) {
#pragma unroll
            for (int dc = -1; dc <= 1; ++dc) {
              const int cc = 2 * pwg + dc;
              if (cc < 0 || cc >= p.W1) continue;
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const uint4 v = lds128(rswz(cc - c0, 2 * e + hq));
                hm[e].x = fmaxf(hm[e].x, __uint_as_float(v.x)); hm[e].y = fmaxf(hm[e].y, __uint_as_float(v.y));
                hm[e].z = fmaxf(hm[e].z, __uint_as_float(v.z)); hm[e].w = fmaxf(hm[e].w, __uint_as_float(v.w));
              }
            }
          }
          const bool odd = oh & 1;
          const bool emit = r > 0 && (odd || oh == p.H1 - 1);
          if (!odd || r > 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              run[e].x = fmaxf(run[e].x, hm[e].x); run[e].y = fmaxf(run[e].y, hm[e].y);
              run[e].z = fmaxf(run[e].z, hm[e].z); run[e].w = fmaxf(run[e].w, hm[e].w);
            }
          }
          if (emit && pw_ok) {
            const int ph = oh >> 1;
            const long long vox = ((long long)bt * p.H2 + ph) * p.W2 + pwg;
            __nv_bfloat16* hi = split_hi(p.pooled, vox, 64) + hq * 4;
#pragma unroll
            for (int e = 0; e < 8; ++e) store_split4(hi + 8 * e, hi + 64 + 8 * e, run[e]);
          }
          if (odd) {
#pragma unroll
            for (int e = 0; e < 8; ++e) run[e] = hm[e];
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      }
    }
  } else {
    // ================= ring staging + A builders: warps 5-8 write the hi part of a stage, warps 9-12 the mid part =================
    const int bt_ = threadIdx.x - 160;                             // 0..255
    const int m = (warp & 3) * 32 + lane;
    const int gh = (warp - 5) >> 2;                                // 0: hi, 1: mid
    const uint32_t ring_hi = sb + OFF_RING_HI3, ring_mid = sb + OFF_RING_MID3;
    for (int i = bt_; i < RING3_PLANE_BYTES / 4; i += NBUILD) {
      sts32(ring_hi + 4u * i, 0u);
      sts32(ring_mid + 4u * i, 0u);
    }
    const uint32_t ring_mine = gh ? ring_mid : ring_hi;
    const uint32_t a_dst = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(A_COL0_3 + gh * HALF_COLS);
    uint32_t full_remote[A3_STAGES];
#pragma unroll
    for (int s = 0; s < A3_STAGES; ++s) full_remote[s] = mapa_rank(full_bar(s), 0);
    mbar_wait(w_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = pair; i < steps; i += npairs) {
      const int u = min(2 * i + (int)rank, units - 1);
      const int ct = u % ct_n, ur = u / ct_n;
      const int rb = ur % rb_n, bt = ur / rb_n;
      const int b = bt / p.T, t = bt % p.T;
      const int first = rb * ROWS_PER_UNIT - 1;
      const int iw0 = 2 * (ct_n > 1 ? 126 * ct - 1 : 0) - 4;
      const int spr = bt_ & 127, srr = bt_ >> 7;
      const int xpr = 128 + (bt_ & 3), xrow = bt_ >> 2;
      auto col_ok = [&](int j) { return iw0 + j >= 0 && iw0 + j < p.W; };
      const bool ok0 = col_ok(2 * spr), ok1 = col_ok(2 * spr + 1), xok0 = col_ok(2 * xpr), xok1 = col_ok(2 * xpr + 1);
      const float* xb = p.x + (long long)b * 3 * p.T * p.H * p.W + iw0;
      const long long plane_stride = (long long)p.H * p.W;
      auto in_val = [&](int plane, int ih, int j, bool ok) -> float {
        const int c = plane / 3, f = t + plane % 3 - 1;
        if (!ok || f < 0 || f >= p.T || ih < 0 || ih >= p.H) return 0.f;
        return __ldg(xb + ((long long)c * p.T + f) * plane_stride + (long long)ih * p.W + j);
      };
      auto slot_of = [&](int ih) { return (ih + 2 * RING3_SLOTS) % RING3_SLOTS; };   // ih >= -5
      auto store_pair = [&](int plane, int ih, int pr, float v0, float v1) {
        __nv_bfloat16 h0, m0, h1, m1;
        split_bf16(v0, h0, m0);
        split_bf16(v1, h1, m1);
        const uint32_t off = (uint32_t)(((plane * RING3_SLOTS + slot_of(ih)) * RING_PAIRS + pr) * 4);
        sts32(ring_hi + off, pack_bf16x2(h0, h1));
        sts32(ring_mid + off, pack_bf16x2(m0, m1));
      };
      asm volatile("bar.sync 2, 256;" ::: "memory");               // previous unit's readers are done
      const int y0 = 2 * first - 3;                                // rows y0 .. y0 + 8 feed the first batch (y0 + 9 only meets zero weights)
      for (int plane = 0; plane < 9; ++plane) {
        float v0[5], v1[5];
#pragma unroll
        for (int g2 = 0; g2 < 5; ++g2) {
          const int rr = 2 * g2 + srr;
          v0[g2] = in_val(plane, y0 + rr, 2 * spr, rr < 9 && ok0);
          v1[g2] = in_val(plane, y0 + rr, 2 * spr + 1, rr < 9 && ok1);
        }
#pragma unroll
        for (int g2 = 0; g2 < 5; ++g2)
          if (2 * g2 + srr < 9) store_pair(plane, y0 + 2 * g2 + srr, spr, v0[g2], v1[g2]);
      }
      for (int idx = xrow; idx < 81; idx += 64) {
        const int plane = idx / 9, ih = y0 + idx % 9;
        store_pair(plane, ih, xpr, in_val(plane, ih, 2 * xpr, xok0), in_val(plane, ih, 2 * xpr + 1, xok1));
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      for (int bch = 0; bch < NBATCH; ++bch) {
        const int oh = first + RB * bch;                            // the batch's first conv row
        // rows 2 oh + 6 .. 2 oh + 9 for the next batch, fetched while this one is built
        float pf0[18], pf1[18], px0 = 0.f, px1 = 0.f;
        const bool more = bch + 1 < NBATCH;
        if (more) {
#pragma unroll
          for (int plane = 0; plane < 9; ++plane) {
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              pf0[2 * plane + h2] = in_val(plane, 2 * oh + 6 + 2 * h2 + srr, 2 * spr, ok0);
              pf1[2 * plane + h2] = in_val(plane, 2 * oh + 6 + 2 * h2 + srr, 2 * spr + 1, ok1);
            }
          }
          if (xrow < 36) {
            px0 = in_val(xrow >> 2, 2 * oh + 6 + (xrow & 3), 2 * xpr, xok0);
            px1 = in_val(xrow >> 2, 2 * oh + 6 + (xrow & 3), 2 * xpr + 1, xok1);
          }
        }
        uint32_t slot_off[2 * KSTEPS];                              // group g = input row 2 oh - 3 + g
#pragma unroll
        for (int g = 0; g < 2 * KSTEPS; ++g) slot_off[g] = (uint32_t)((slot_of(2 * oh - 3 + g) * RING_PAIRS + m) * 4);
#pragma unroll 1
        for (int plane = 0; plane < 9; ++plane) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          tcgen05_fence_after();
          const uint32_t pl = ring_mine + (uint32_t)(plane * RING3_SLOTS * RING_PAIRS * 4);
          uint32_t v[8 * KSTEPS];
#pragma unroll
          for (int g = 0; g < 2 * KSTEPS; ++g) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[4 * g + e] = lds32(pl + slot_off[g] + 4u * e);
          }
          tmem_st32(a_dst + (uint32_t)(stage * STAGE3_COLS), v);
          tmem_st8(a_dst + (uint32_t)(stage * STAGE3_COLS + 32), v + 32);
          tmem_st_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(full_remote[stage]);
          if (++stage == A3_STAGES) { stage = 0; phase ^= 1; }
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (more) {
#pragma unroll
          for (int plane = 0; plane < 9; ++plane) {
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) store_pair(plane, 2 * oh + 6 + 2 * h2 + srr, spr, pf0[2 * plane + h2], pf1[2 * plane + h2]);
          }
          if (xrow < 36) store_pair(xrow >> 2, 2 * oh + 6 + (xrow & 3), xpr, px0, px1);
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// reference filter (64,3,3,7,7) = [oc][441] fp32 -> packed split [2][64][576] bf16 in the kernel's K order
__global__ void stem_pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * KTOT) return;
  const int oc = i / KTOT, k = i % KTOT, plane = k / 64, kh = (k % 64) / 8, kw = k % 8;
  // tap slot 0 of a group is the zero tap, slots 1..7 are kw = 0..6: a group's 8 pixels then start at the EVEN input column
  // 2 ow - 4, so the ring's pixel pairs are aligned pairs of the input row (8- / 16-byte global loads in the staging threads)
  const float v = (kh < 7 && kw >= 1) ? w[oc * 441 + (plane * 7 + kh) * 7 + kw - 1] : 0.f;
  __nv_bfloat16 hi, mid;
  split_bf16(v, hi, mid);
  out[i] = hi;
  out[64 * KTOT + i] = mid;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
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
    if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, p2::SMEM2_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, p2::SMEM2_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p3::SMEM3_BYTES);
    return e;
  });
}

}  // namespace stemtc

cudaError_t launch_stem_pack_weight(const float* w_oc441, void* out, cudaStream_t st) {
  stemtc::stem_pack_weight_kernel<<<(64 * stemtc::KTOT + 255) / 256, 256, 0, st>>>(w_oc441, reinterpret_cast<__nv_bfloat16*>(out));
  return cudaGetLastError();
}

// conv rows of any width run on the pair kernel with the pool fused (column tiles, see stem_tc2_kernel); TUBER_STEM_SINGLE=1 in the
// environment keeps rows wider than 128 outputs on the single-CTA kernel + separate pool (the cross-check of the tests)
bool stem_pool_is_fused(int W1) {
  static const bool single = [] { const char* e = getenv("TUBER_STEM_SINGLE"); return e && e[0] == '1'; }();
  return W1 <= 128 || !single;
}

// y: fp32 conv rows [B,T,H1,W1,64] (only written when the pool is not fused; may be null otherwise);
// pooled: split [B,T,H2,W2,64] (only written when stem_pool_is_fused(W1))
cudaError_t launch_stem_conv(const float* x, const void* wpk, const float* scale, const float* shift, float* y, void* pooled, int B,
                             int T, int H, int W, int H1, int W1, cudaStream_t st, const uint8_t* frames_u8, const float* lut) {
  using namespace stemtc;
  if (frames_u8 && !(stem_pool_is_fused(W1) && pooled != nullptr && lut != nullptr)) return cudaErrorInvalidValue;   // uint8 input: pair kernel only
  cudaError_t e = init_once();
  if (e != cudaSuccess) return e;
  CUtensorMap tmW, tmOut;
  {
    cuuint64_t dims[3] = {KTOT, 64, 2};
    cuuint64_t strides[2] = {KTOT * 2, 64 * KTOT * 2};
    cuuint32_t box[3] = {64, NCH, 2}, es[3] = {1, 1, 1};
    if (g_encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wpk), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  const bool fuse = stem_pool_is_fused(W1) && pooled != nullptr;
  if (fuse) {                                                      // CTA pairs (cta_group::2), pool fused
    const int W2 = (W1 - 1) / 2 + 1;
    const int ct_n = W1 <= 128 ? 1 : (W2 + 62) / 63;
    const char* nv = getenv("TUBER_STEM_NO_VEC");                    // read per call: scalar staging (the tests' cross-check)
    const int vec4 = (!frames_u8 && ct_n == 1 && W % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && !(nv && nv[0] == '1')) ? 1 : 0;
    Params p{frames_u8, lut, x, scale, shift, pooled, B, T, H, W, H1, W1, (H1 - 1) / 2 + 1, W2, 1, ct_n, vec4};
    const int units = B * T * ((H1 + ROWS_PER_UNIT - 1) / ROWS_PER_UNIT) * ct_n;
    int pairs = device_num_sms() / 2;
    if (pairs > (units + 1) / 2) pairs = (units + 1) / 2;
    if (pairs < 1) pairs = 1;
    const char* s3 = getenv("TUBER_STEM3");                          // read per call: the tests compare both forms in one process
    const bool stem3 = s3 && s3[0] == '1';
    if (frames_u8) stem_tc2_kernel<true><<<2 * pairs, NUM_THREADS, p2::SMEM2_BYTES, st>>>(tmW, p);
    else if (stem3) stem_tc3_kernel<<<2 * pairs, NUM_THREADS, p3::SMEM3_BYTES, st>>>(tmW, p);
    else stem_tc2_kernel<false><<<2 * pairs, NUM_THREADS, p2::SMEM2_BYTES, st>>>(tmW, p);
    return cudaGetLastError();
  }
  tmOut = tmW;
  if (!fuse) {
    if (y == nullptr) return cudaErrorInvalidValue;
    cuuint64_t dims[3] = {64, (cuuint64_t)W1, (cuuint64_t)B * T * H1};
    cuuint64_t strides[2] = {64 * 4, (cuuint64_t)W1 * 64 * 4};
    cuuint32_t box[3] = {NCH, 128, 1}, es[3] = {1, 1, 1};
    if (g_encode(&tmOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  Params p{nullptr, nullptr, x, scale, shift, fuse ? pooled : nullptr, B, T, H, W, H1, W1, (H1 - 1) / 2 + 1, (W1 - 1) / 2 + 1, fuse ? 1 : 0, 1, 0};
  const int units = B * T * ((W1 + 127) / 128) * ((H1 + ROWS_PER_UNIT - 1) / ROWS_PER_UNIT);
  int pairs = device_num_sms() / 2;
  if (pairs > units) pairs = units;
  if (pairs < 1) pairs = 1;
  stem_tc_kernel<<<2 * pairs, NUM_THREADS, SMEM_BYTES, st>>>(tmW, tmOut, p);
  return cudaGetLastError();
}
