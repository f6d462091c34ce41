// Attention core on tcgen05 (sm_100a): O = softmax(scale * Q K^T + key padding mask) V for head_dim 32 wherever a (sequence, head)
// fills at least half of a 128-query tile and has at least one 128-key chunk: the encoder self-attention (L = S = H'W', 256 at
// 256 x 256 input, transformer.py:158-164), the class branch's spatial attention (L = S = H'W', batch T'B, transformer_layers.py:79-84),
// the class cross-attention (L = DEC_LAYERS x Q queries over S = T'H'W' tokens, tuber_ava.py:137-139), JHMDB's decoder self-attention
// (L = S = 320) and the long-term context layer (tuber_forward_ltc: 1024 tokens over a 64-clip window = 16 384 keys).  Only the
// 15-query decoder attentions and the per-pixel temporal attention (T' <= 8 keys) stay on the warp-level kernels of attn_mma.cu.
// A key padding mask (utils/misc.py:385-399, True = padding) travels as one flag byte per key next to the chunk in shared memory.
//
// One CTA = one (sequence n, head h, tile of 128 queries); it walks the keys in chunks of 128, flash-attention style, with the
// library's usual precision (common.cuh): every fp32 operand is split into bf16 hi + mid and a product is evaluated as
// a_hi*b_hi + a_hi*b_mid + a_mid*b_hi with fp32 accumulation in tensor memory.
//
//   warps 9-12 loaders: Q tile once, then per chunk K (row = key: hi[32] | mid[32] bf16 = one 128-byte swizzled row, K-major)
//              and V TRANSPOSED (row = head dim, 128 keys per row as two 64-key swizzle atoms; hi and mid planes) into a
//              2-stage shared-memory ring, converted from the fp32 projections on the fly.
//   warp 0     MMA issuer.  S = Q K^T: 6 tcgen05.mma (M=128, N=128, K=16; 2 dim halves x 3 passes) into TMEM columns [0,128).
//              O_c = P V: 24 tcgen05.mma (M=128, N=32, K=16; 8 key steps x 3 passes) with P read FROM TENSOR MEMORY (A operand)
//              and V^T from shared memory, into TMEM columns [192,224) -- a fresh accumulator per chunk.
//   warps 1-8  softmax, two threads per query row (its TMEM lane), each owning 64 of the chunk's keys and 16 output dims (the halves
//              swap their maxima through shared memory): pass 1 reads the scores for the row maximum, pass 2 reads
//              them again, exponentiates, splits into bf16 hi | mid and writes P back to tensor memory (hi of key half g over the
//              score columns [64g, 64g+32) that thread has already consumed, mid to columns [128,192)); then adds its half of O_c to the running output in
//              registers with the usual rescaling  o = o * exp(m_old - m_new) + O_c  -- no TMEM accumulator to rescale.
// Two CTAs share an SM (81 KB of shared memory and 256 TMEM columns each), so one CTA's softmax overlaps the other's MMAs.
//
// PREP variant (used by the plan, which has workspace to give): every (key sequence, head) is read by all the query tiles of all
// the clips that share it, so attn_tc_prep_kernel converts K / V ONCE into the exact 32 KB shared-memory image of each chunk
// (swizzled K rows + transposed V planes) in global memory, and the attention CTAs fetch a chunk with one cp.async.bulk
// (completion on the stage's mbarrier) -- the per-CTA conversion (a quarter of all issued instructions) disappears.
#include <cuda.h>

#include "kernels.h"

namespace attn_tc {

constexpr int D = 32, QT = 128, KC = 128;
constexpr int Q_BYTES = QT * 128;                 // 128 rows x (hi 64 B | mid 64 B)
constexpr int K_BYTES = KC * 128;
constexpr int VT_PLANE = 2 * 32 * 128;            // [2 key atoms][32 dims][64 keys] bf16 = 8 KB
constexpr int STAGE_BYTES = K_BYTES + 2 * VT_PLANE;   // 32 KB
constexpr int OFF_Q = 0, OFF_STAGE = Q_BYTES, OFF_BAR = OFF_STAGE + 2 * STAGE_BYTES;
constexpr int OFF_XCH = OFF_BAR + 128;              // [2][128] floats
constexpr int OFF_MASK = OFF_XCH + 1024;            // [2 stages][128 keys] bytes: 1 = the key is padding (key_padding_mask)
constexpr int SMEM_BYTES = OFF_MASK + 256;
constexpr int THREADS = 416;                      // warp 0 MMA, warps 1-8 softmax (two per query row), warps 9-12 loaders
constexpr int THREADS_PREP = 320;                 // PREP: warp 9 = one thread issuing bulk copies; softmax group 0 stages Q
constexpr int TMEM_COLS = 256;
constexpr int COL_S = 0, COL_PMID = 128, COL_O = 192;

TB_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
TB_DEVINL long long seq_row0(const SeqMap& m, int n) {
  return (long long)(n / m.inner) * m.outer + (long long)(n % m.inner) * m.inner_stride;
}
TB_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
TB_DEVINL void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
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
TB_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
TB_DEVINL void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
TB_DEVINL void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
TB_DEVINL void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
TB_DEVINL void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: A = 128 lanes (rows) x 8 columns per K = 16 step, two bf16 per 32-bit column
TB_DEVINL void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
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
TB_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
TB_DEVINL void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
TB_DEVINL void sts16(uint32_t addr, uint16_t v) { asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(v) : "memory"); }
TB_DEVINL float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32 pairs (one issue slot per two operations: the softmax warps are issue bound)
typedef unsigned long long u64;
TB_DEVINL u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
TB_DEVINL void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
TB_DEVINL u64 add2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
TB_DEVINL u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// K-major, 128B-swizzle shared-memory matrix descriptor (8-row groups 1024 B apart; same encoding as gemm_tc.cu / stem_tc.cu)
TB_DEVINL uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {      // bf16 x bf16 -> fp32, both operands K-major
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// 16-byte chunk j of row r of a 128-byte-row swizzled tile
TB_DEVINL uint32_t swz(uint32_t base, int r, int j) { return base + (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4); }

// 32 fp32 values of a row (8 x float4) times `mul` -> one swizzled operand row: chunks 0-3 = hi, 4-7 = mid
TB_DEVINL void stage_row(uint32_t tile, int r, const float* __restrict__ src, float mul, bool valid) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 hi = make_uint4(0u, 0u, 0u, 0u), mid = hi;
    if (valid) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * j), b = __ldg(reinterpret_cast<const float4*>(src) + 2 * j + 1);
      split_bf16x2(a.x * mul, a.y * mul, hi.x, mid.x);
      split_bf16x2(a.z * mul, a.w * mul, hi.y, mid.y);
      split_bf16x2(b.x * mul, b.y * mul, hi.z, mid.z);
      split_bf16x2(b.z * mul, b.w * mul, hi.w, mid.w);
    }
    sts128(swz(tile, r, j), hi);
    sts128(swz(tile, r, 4 + j), mid);
  }
}

// thread t (0..127) = key t of chunk c: K row (hi | mid) and the key's column of the transposed V planes -> the stage image at
// shared address st_base (K_BYTES of K, then V^T hi plane, then V^T mid plane)
TB_DEVINL void stage_chunk(uint32_t st_base, int t, const AttnArgs& p, int h, long long k0, int c) {
  const int key = c * KC + t;
  const bool valid = key < p.S;
  const long long krow = k0 + (long long)key * p.km.step;
  stage_row(st_base, t, p.k + krow * p.ldk + h * D, 1.f, valid);
  // V^T: value (key t, dim d) -> row d of key atom t / 64, 2-byte column t % 64
  const uint32_t vt = st_base + K_BYTES + (uint32_t)((t >> 6) * 4096), kc = (uint32_t)(t & 63);
  const float* vp = p.v + krow * p.ldv + h * D;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) a = __ldg(reinterpret_cast<const float4*>(vp) + j);
    const float vals[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = 4 * j + e;
      __nv_bfloat16 hi, mid;
      split_bf16(vals[e], hi, mid);
      const uint32_t off = (uint32_t)d * 128u + ((((kc >> 3) ^ (uint32_t)(d & 7)) << 4) | ((kc & 7u) << 1));
      sts16(vt + off, __bfloat16_as_ushort(hi));
      sts16(vt + VT_PLANE + off, __bfloat16_as_ushort(mid));
    }
  }
}

// K / V of one (key sequence, head, chunk) -> the chunk's shared-memory image, stored in global memory (see PREP above)
__global__ void __launch_bounds__(128)
attn_tc_prep_kernel(AttnArgs p, uint8_t* __restrict__ img) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  const int c = blockIdx.x, h = blockIdx.y, nkv = blockIdx.z, nchunks = gridDim.x;
  pdl_trigger();
  pdl_wait();
  stage_chunk(sb, threadIdx.x, p, h, seq_row0(p.km, nkv), c);
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(img + ((size_t)(nkv * p.H + h) * nchunks + c) * STAGE_BYTES);
  const uint4* src = reinterpret_cast<const uint4*>(smem);
  for (int i = threadIdx.x; i < STAGE_BYTES / 16; i += 128) dst[i] = src[i];
}

template <bool PREP>
__global__ void __launch_bounds__(PREP ? THREADS_PREP : THREADS, 2)
attn_tc_kernel(AttnArgs p, const uint8_t* __restrict__ img) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar = sb + OFF_BAR;
  auto kv_full = [&](int s) { return bar + 8u * s; };          // loaders (128 arrivals) -> MMA
  auto kv_empty = [&](int s) { return bar + 8u * (2 + s); };   // tcgen05.commit after P V -> loaders
  const uint32_t s_full = bar + 32, p_full = bar + 40, o_full = bar + 48, q_full = bar + 56, tmem_slot = bar + 64;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.z, h = blockIdx.y, l0 = blockIdx.x * QT;
  const int nchunks = (p.S + KC - 1) / KC;

  if (threadIdx.x == 0) {
    if (sb & 1023u) __trap();
    mbar_init(kv_full(0), PREP ? 1 : 128); mbar_init(kv_full(1), PREP ? 1 : 128);
    mbar_init(q_full, 128);
    mbar_init(kv_empty(0), 1); mbar_init(kv_empty(1), 1);
    mbar_init(s_full, 1); mbar_init(p_full, 256); mbar_init(o_full, 1);
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
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc_s = make_idesc(128, KC), idesc_o = make_idesc(128, D);
    const uint64_t q_desc = make_smem_desc(sb + OFF_Q);
    mbar_wait(q_full, 0);
    for (int c = 0; c < nchunks; ++c) {
      const int stage = c & 1;
      const uint32_t st_base = sb + OFF_STAGE + stage * STAGE_BYTES;
      mbar_wait(kv_full(stage), (uint32_t)((c >> 1) & 1));
      tcgen05_fence_after();
      if (elect_one()) {
        const uint64_t k_desc = make_smem_desc(st_base);
        // 32 bytes (16 bf16) per K step: row bytes [0,64) = hi dims, [64,128) = mid dims; descriptor address unit = 16 B
#pragma unroll
        for (int dh = 0; dh < 2; ++dh) {
          umma_ss(tmem_base + COL_S, q_desc + 4 + 2 * dh, k_desc + 2 * dh, idesc_s, dh);        // Q_mid K_hi
          umma_ss(tmem_base + COL_S, q_desc + 2 * dh, k_desc + 4 + 2 * dh, idesc_s, 1u);        // Q_hi  K_mid
          umma_ss(tmem_base + COL_S, q_desc + 2 * dh, k_desc + 2 * dh, idesc_s, 1u);            // Q_hi  K_hi
        }
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, (uint32_t)(c & 1));
      tcgen05_fence_after();
      if (elect_one()) {
        const uint64_t vh_desc = make_smem_desc(st_base + K_BYTES), vm_desc = make_smem_desc(st_base + K_BYTES + VT_PLANE);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {                         // 16 keys per step; 64-key atoms are 4 KB apart
          const uint64_t off = (uint64_t)((ks >> 2) * (4096 >> 4) + 2 * (ks & 3));
          const uint32_t p_hi = tmem_base + COL_S + 64 * (ks >> 2) + 8 * (ks & 3), p_mid = tmem_base + COL_PMID + 8 * ks;   // hi: per key half, see softmax
          umma_ts(tmem_base + COL_O, p_mid, vh_desc + off, idesc_o, ks == 0 ? 0u : 1u);
          umma_ts(tmem_base + COL_O, p_hi, vm_desc + off, idesc_o, 1u);
          umma_ts(tmem_base + COL_O, p_hi, vh_desc + off, idesc_o, 1u);
        }
        umma_commit(o_full);
        umma_commit(kv_empty(stage));
      }
      __syncwarp();
    }
  } else if (warp <= 8) {
    // ================= softmax + output accumulation: two threads per query row, each owns 64 of the chunk's 128 keys and 16 of
    // the 32 output dims (group g = 0: warps 1-4, g = 1: warps 5-8; a warp reaches the TMEM lanes 32 (warp % 4) ..) =================
    const int g = (warp - 1) >> 2, lg = warp & 3, row = lg * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(lg * 32) << 16);
    const uint32_t col_s = COL_S + 64 * g, col_pm = COL_PMID + 32 * g;
    float* xch = reinterpret_cast<float*>(smem + OFF_XCH);        // [2][128]: the two halves of a row exchange their maxima / sums
    float o[D / 2];
#pragma unroll
    for (int i = 0; i < D / 2; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f;
    if (PREP && g == 0) {                                        // no loader warps: group 0 stages the query rows
      const int qrow = l0 + row;
      const long long grow = seq_row0(p.qm, n) + (long long)qrow * p.qm.step;
      stage_row(sb + OFF_Q, row, p.q + grow * p.ldq + h * D, p.scale * 1.4426950408889634f, qrow < p.L);
      fence_proxy_async();
      mbar_arrive(q_full);
    }
    const bool has_mask = !PREP && p.kpm != nullptr;
    for (int c = 0; c < nchunks; ++c) {
      const int nvalid = min(KC, p.S - c * KC) - 64 * g;           // valid keys among this thread's 64 (may be <= 0 in the last chunk)
      mbar_wait(s_full, (uint32_t)(c & 1));
      tcgen05_fence_after();
      // padding flags of this thread's 64 keys (written by the loaders before the chunk's kv_full arrive)
      const uint8_t* mflag = smem + OFF_MASK + (c & 1) * 128 + 64 * g;
      bool plain = nvalid >= 64;
      if (has_mask) {
        const uint4* mf = reinterpret_cast<const uint4*>(mflag);
        const uint4 f0 = mf[0], f1 = mf[1], f2 = mf[2], f3 = mf[3];
        plain = plain && ((f0.x | f0.y | f0.z | f0.w | f1.x | f1.y | f1.z | f1.w | f2.x | f2.y | f2.z | f2.w | f3.x | f3.y | f3.z | f3.w) == 0u);
      }
      float cm = -INFINITY;
      if (plain) {                                               // full half chunk without padding: no key predicates
#pragma unroll 1
        for (int b = 0; b < 2; ++b) {
          uint32_t v[32];
          tmem_ld32(t_lane + col_s + 32 * b, v);
          float c0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])), c1 = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
#pragma unroll
          for (int j = 4; j < 32; j += 4) {
            c0 = fmaxf(c0, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
            c1 = fmaxf(c1, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
          }
          cm = fmaxf(cm, fmaxf(c0, c1));
        }
      } else {
#pragma unroll 1
        for (int b = 0; b < 2; ++b) {
          uint32_t v[32];
          tmem_ld32(t_lane + col_s + 32 * b, v);
          uint32_t fw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          if (has_mask) {
            const uint4 a = *reinterpret_cast<const uint4*>(mflag + 32 * b), bb = *reinterpret_cast<const uint4*>(mflag + 32 * b + 16);
            fw[0] = a.x; fw[1] = a.y; fw[2] = a.z; fw[3] = a.w; fw[4] = bb.x; fw[5] = bb.y; fw[6] = bb.z; fw[7] = bb.w;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const bool use = (32 * b + j < nvalid) && ((fw[j >> 2] >> (8 * (j & 3))) & 0xffu) == 0u;
            cm = fmaxf(cm, use ? __uint_as_float(v[j]) : -INFINITY);
          }
        }
      }
      xch[g * 128 + row] = cm;
      asm volatile("bar.sync 1, 256;" ::: "memory");             // (the slot is rewritten only after every thread has passed p_full of this chunk)
      cm = fmaxf(cm, xch[(g ^ 1) * 128 + row]);
      const float m_new = fmaxf(m, cm);                          // -inf only while every key seen so far was padding
      const float m_sub = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = (m_new == -INFINITY) ? 1.f : ex2(m - m_new);   // first chunk with a key: 2^-inf = 0
      float psum = 0.f;
#pragma unroll 1
      for (int b = 0; b < 2; ++b) {
        uint32_t v[32];
        tmem_ld32(t_lane + col_s + 32 * b, v);
        uint32_t hi[16], mid[16];
        if (plain) {
          const u64 negm = pack2(-m_new, -m_new), neg1 = pack2(-1.f, -1.f);
          u64 ps2 = pack2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float x0, x1;
            unpack2(add2(pack2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), negm), x0, x1);
            const u64 pp = pack2(ex2(x0), ex2(x1));
            ps2 = add2(ps2, pp);
            float p0, p1;
            unpack2(pp, p0, p1);
            const __nv_bfloat162 hb = __floats2bfloat162_rn(p0, p1);
            hi[j] = *reinterpret_cast<const uint32_t*>(&hb);
            float d0, d1;                                        // p - float(hi), exactly
            unpack2(fma2(pack2(__uint_as_float(hi[j] << 16), __uint_as_float(hi[j] & 0xffff0000u)), neg1, pp), d0, d1);
            const __nv_bfloat162 mb = __floats2bfloat162_rn(d0, d1);
            mid[j] = *reinterpret_cast<const uint32_t*>(&mb);
          }
          float s0, s1;
          unpack2(ps2, s0, s1);
          psum += s0 + s1;
        } else {
          uint32_t fw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          if (has_mask) {
            const uint4 a = *reinterpret_cast<const uint4*>(mflag + 32 * b), bb = *reinterpret_cast<const uint4*>(mflag + 32 * b + 16);
            fw[0] = a.x; fw[1] = a.y; fw[2] = a.z; fw[3] = a.w; fw[4] = bb.x; fw[5] = bb.y; fw[6] = bb.z; fw[7] = bb.w;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const bool u0 = (32 * b + 2 * j < nvalid) && ((fw[j >> 1] >> (16 * (j & 1))) & 0xffu) == 0u;
            const bool u1 = (32 * b + 2 * j + 1 < nvalid) && ((fw[j >> 1] >> (16 * (j & 1) + 8)) & 0xffu) == 0u;
            const float p0 = u0 ? ex2(__uint_as_float(v[2 * j]) - m_sub) : 0.f;
            const float p1 = u1 ? ex2(__uint_as_float(v[2 * j + 1]) - m_sub) : 0.f;
            psum += p0 + p1;
            split_bf16x2(p0, p1, hi[j], mid[j]);
          }
        }
        tmem_st16(t_lane + col_s + 16 * b, hi);                  // over score columns this thread has already consumed
        tmem_st16(t_lane + col_pm + 16 * b, mid);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      mbar_arrive(p_full);
      l = l * corr + psum;
      m = m_new;
      mbar_wait(o_full, (uint32_t)(c & 1));
      tcgen05_fence_after();
      {
        uint32_t oc[16];
        tmem_ld16(t_lane + COL_O + 16 * g, oc);
#pragma unroll
        for (int i = 0; i < D / 2; ++i) o[i] = fmaf(o[i], corr, __uint_as_float(oc[i]));
      }
      tcgen05_fence_before();
    }
    // the row's normaliser = the sum of both halves
    asm volatile("bar.sync 1, 256;" ::: "memory");
    xch[g * 128 + row] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l += xch[(g ^ 1) * 128 + row];
    if (l0 + row < p.L) {
      const float inv = 1.f / l;
      const long long orow = seq_row0(p.om, n) + (long long)(l0 + row) * p.om.step;
      const int col = h * D + 16 * g;
      if (p.o_f32) {
        float4* dst = reinterpret_cast<float4*>(p.o_f32 + orow * p.ldo + col);
#pragma unroll
        for (int i = 0; i < D / 8; ++i) dst[i] = make_float4(o[4 * i] * inv, o[4 * i + 1] * inv, o[4 * i + 2] * inv, o[4 * i + 3] * inv);
      }
      if (p.o_split) {
        __nv_bfloat16* hp = split_hi(p.o_split, orow, p.ldo) + col;
#pragma unroll
        for (int i = 0; i < D / 8; ++i)
          store_split4(hp + 4 * i, hp + p.ldo + 4 * i, make_float4(o[4 * i] * inv, o[4 * i + 1] * inv, o[4 * i + 2] * inv, o[4 * i + 3] * inv));
      }
    }
  } else {
    // ================= loaders: thread = query row (once), then = key of the chunk =================
    const int t = threadIdx.x - 288;                             // 0..127 (PREP: 0..31, only thread 0 works)
    if (!PREP) {
      const int qrow = l0 + t;
      const long long grow = seq_row0(p.qm, n) + (long long)qrow * p.qm.step;
      // scores in the log2 domain: softmax(x) = 2^(x log2 e - max)
      stage_row(sb + OFF_Q, t, p.q + grow * p.ldq + h * D, p.scale * 1.4426950408889634f, qrow < p.L);
      fence_proxy_async();
      mbar_arrive(q_full);
    }
    if (PREP) {
      // one thread fetches the pre-converted 32 KB image of each chunk with a bulk copy that completes on the stage's barrier
      if (t == 0) {
        const int nkv = p.tc_shared_kv ? 0 : n;
        const uint8_t* src = img + (size_t)(nkv * p.H + h) * nchunks * STAGE_BYTES;
        for (int c = 0; c < nchunks; ++c) {
          const int stage = c & 1;
          const uint32_t st_base = sb + OFF_STAGE + stage * STAGE_BYTES;
          if (c >= 2) mbar_wait(kv_empty(stage), (uint32_t)(((c >> 1) - 1) & 1));
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(kv_full(stage)), "r"((uint32_t)STAGE_BYTES) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(st_base), "l"(src + (size_t)c * STAGE_BYTES), "r"((uint32_t)STAGE_BYTES), "r"(kv_full(stage)) : "memory");
        }
      }
    } else {
      const long long k0 = seq_row0(p.km, n);
      for (int c = 0; c < nchunks; ++c) {
        const int stage = c & 1;
        const uint32_t st_base = sb + OFF_STAGE + stage * STAGE_BYTES;
        if (c >= 2) mbar_wait(kv_empty(stage), (uint32_t)(((c >> 1) - 1) & 1));
        stage_chunk(st_base, t, p, h, k0, c);
        if (p.kpm) {                                             // padding flag of key t of this chunk (read by the softmax warps)
          const int key = c * KC + t;
          smem[OFF_MASK + stage * 128 + t] = key < p.S ? p.kpm[(long long)(n / p.kpm_div) * p.S + key] : (uint8_t)0;
        }
        fence_proxy_async();                                     // generic-proxy writes -> visible to the tensor core's reads
        mbar_arrive(kv_full(stage));
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

}  // namespace attn_tc

bool attention_tc_supported(const AttnArgs& a) {
  return a.D == attn_tc::D && a.L >= attn_tc::QT / 2 && a.S >= attn_tc::KC && a.NB > 0 && a.NB <= 65535 && a.H <= 65535 &&
         a.ldq % 4 == 0 && a.ldk % 4 == 0 && a.ldv % 4 == 0 && a.ldo % 4 == 0;
}

// pre-converted K / V images pay when several query tiles read the same keys: the long-term context layer (8 tiles per clip, or
// every clip's tiles on a shared window) and JHMDB's class cross-attention (15 tiles); not with a key padding mask (the image has none)
bool attention_tc_wants_prep(const AttnArgs& a) {
  if (!attention_tc_supported(a) || a.kpm != nullptr) return false;
  const bool shared = a.km.outer == 0 && a.km.inner_stride == 0;
  const long long tiles_per_kv = (long long)ceil_div(a.L, attn_tc::QT) * (shared ? a.NB : 1);
  return a.S >= 512 && tiles_per_kv >= 4;
}

size_t attention_tc_scratch_bytes(const AttnArgs& a) {
  const bool shared = a.km.outer == 0 && a.km.inner_stride == 0;
  return (size_t)(shared ? 1 : a.NB) * a.H * ceil_div(a.S, attn_tc::KC) * attn_tc::STAGE_BYTES;
}

// a.tc_scratch (attention_tc_scratch_bytes, 16-byte aligned): K / V are converted once by attn_tc_prep_kernel (second launch of this
// call); without scratch every CTA converts its own chunks
cudaError_t launch_attention_tc(const AttnArgs& a0, cudaStream_t st) {
  using namespace attn_tc;
  if (!attention_tc_supported(a0)) return cudaErrorNotSupported;
  static DeviceOnce once;
  {
    cudaError_t e = once.run([]() {
      cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      return e;
    });
    if (e != cudaSuccess) return e;
  }
  AttnArgs a = a0;
  a.tc_shared_kv = (a.km.outer == 0 && a.km.inner_stride == 0) ? 1 : 0;
  dim3 grid(ceil_div(a.L, QT), a.H, a.NB);
  if (a.tc_scratch == nullptr || a.kpm != nullptr)
    return launch_pdl(attn_tc_kernel<false>, grid, dim3(THREADS), SMEM_BYTES, st, a, (const uint8_t*)nullptr);
  dim3 pgrid(ceil_div(a.S, KC), a.H, a.tc_shared_kv ? 1 : a.NB);
  cudaError_t e = launch_pdl(attn_tc_prep_kernel, pgrid, dim3(128), STAGE_BYTES, st, a, (uint8_t*)a.tc_scratch);
  if (e != cudaSuccess) return e;
  return launch_pdl(attn_tc_kernel<true>, grid, dim3(THREADS_PREP), SMEM_BYTES, st, a, (const uint8_t*)a.tc_scratch);
}
