// Attention cores on the warp-level tensor-core path (mma.sync m16n8k16, bf16 operands, fp32 accumulation) for
// head_dim 32 -- every nn.MultiheadAttention site of the TubeR forward with more than a handful of keys:
//   DETR encoder self-attention      transformer.py:158-164      L = S = T'H'W' (256)
//   DETR decoder self / cross        transformer.py:225-241      L = Q (15), S = Q / T'H'W'
//   class-branch spatial attention   transformer_layers.py:79-84 L = S = H'W' (256), one sequence per frame
//   class cross-attention            tuber_ava.py:137-139        L = DEC_LAYERS * Q (90), S = T'H'W' (1024)
// and a dedicated kernel for the class branch's per-pixel temporal attention (transformer_layers.py:86-91, L = S = T' <= 8).
//
// The tiles are far too small for tcgen05 (a 128-row UMMA tile per (sequence, head) would be > 85 % padding and the
// TMEM / barrier set-up costs more than the whole product), so a warp owns 16 queries and keeps Q, the scores and
// the output in registers, flash-attention style: per chunk of 64 keys  S = Q K^T  ->  online softmax  ->  O += P V.
// Precision follows the rest of the library (common.cuh): every fp32 operand x is split into bf16 hi + mid and a
// product a*b is evaluated as a_hi*b_hi + a_hi*b_mid + a_mid*b_hi with fp32 accumulation (~2^-16 operand error).
#include "kernels.h"

namespace attn_mma {

constexpr int D = 32;            // head dim
constexpr int KC = 64;           // keys per chunk
constexpr int QW = 16;           // queries per warp
// warps per CTA = NW (template): 4 (64 queries) or 8 (128 queries: half the K/V staging per query for L > 64)
constexpr int LDS_ROW = 40;      // bf16 per shared-memory row (32 + 8 pad: ldmatrix rows 80 B apart are conflict free)

TB_DEVINL long long seq_row0(const SeqMap& m, int n) {
  return (long long)(n / m.inner) * m.outer + (long long)(n % m.inner) * m.inner_stride;
}
TB_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
TB_DEVINL void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
TB_DEVINL void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
TB_DEVINL void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (x, y) -> packed bf16x2 of the hi parts and of the mid parts
TB_DEVINL void split2(float x, float y, uint32_t& hi, uint32_t& mid) {
  __nv_bfloat16 hx, mx, hy, my;
  split_bf16(x, hx, mx);
  split_bf16(y, hy, my);
  hi = pack_bf16x2(hx, hy);
  mid = pack_bf16x2(mx, my);
}

struct __align__(16) Smem {
  __nv_bfloat16 k_hi[KC][LDS_ROW], k_mid[KC][LDS_ROW], v_hi[KC][LDS_ROW], v_mid[KC][LDS_ROW];
  float msk[KC];                 // 0 or -inf per key of the chunk
};

// one chunk of up to 64 keys (staged in `sm` as bf16 hi | mid) for the 16 queries of a warp: scores, online softmax, O += P V
TB_DEVINL void process_chunk(const Smem& sm, int valid, const uint32_t (&q_hi)[2][4], const uint32_t (&q_mid)[2][4], float (&o)[4][4],
                              float (&mx)[2], float (&ls)[2], int lane) {
  const int t = lane & 3;
  const int mi = lane >> 3, mr = lane & 7;                 // ldmatrix lane addressing: matrix mi = lane / 8, row r = lane % 8
  const uint32_t k_hi_b = smem_u32(&sm.k_hi[mr][mi * 8]), k_mid_b = smem_u32(&sm.k_mid[mr][mi * 8]);
  const uint32_t v_hi_b = smem_u32(&sm.v_hi[(mi & 1) * 8 + mr][(mi >> 1) * 8]), v_mid_b = smem_u32(&sm.v_mid[(mi & 1) * 8 + mr][(mi >> 1) * 8]);
    const int ntiles = (valid + 7) >> 3;                 // 8-key score tiles that hold at least one real key

    // ---- scores: s[j] = Q K^T for keys 8j .. 8j+7 of the chunk ----
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
      if (j < ntiles) {
        uint32_t bh[4], bm[4];
        ldsm_x4(k_hi_b + (uint32_t)(j * 8 * LDS_ROW * 2), bh);
        ldsm_x4(k_mid_b + (uint32_t)(j * 8 * LDS_ROW * 2), bm);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          mma_bf16(s[j], q_mid[ks], bh[2 * ks], bh[2 * ks + 1]);
          mma_bf16(s[j], q_hi[ks], bm[2 * ks], bm[2 * ks + 1]);
          mma_bf16(s[j], q_hi[ks], bh[2 * ks], bh[2 * ks + 1]);
        }
      }
    }
    // ---- mask + online softmax (rows g and g+8; a row's 64 scores live in the 4 lanes of a quad) ----
    float cm[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 mk = *reinterpret_cast<const float2*>(&sm.msk[j * 8 + 2 * t]);
      s[j][0] += mk.x; s[j][1] += mk.y; s[j][2] += mk.x; s[j][3] += mk.y;
      cm[0] = fmaxf(cm[0], fmaxf(s[j][0], s[j][1]));
      cm[1] = fmaxf(cm[1], fmaxf(s[j][2], s[j][3]));
    }
    float corr[2], mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      cm[r] = fmaxf(cm[r], __shfl_xor_sync(0xffffffffu, cm[r], 1));
      cm[r] = fmaxf(cm[r], __shfl_xor_sync(0xffffffffu, cm[r], 2));
      mnew[r] = fmaxf(mx[r], cm[r]);
      corr[r] = (mnew[r] == -INFINITY) ? 1.f : expf(mx[r] - mnew[r]);     // mx = -inf -> 0
      mx[r] = mnew[r];
      ls[r] *= corr[r];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1]; }
    const float sub0 = (mnew[0] == -INFINITY) ? 0.f : mnew[0], sub1 = (mnew[1] == -INFINITY) ? 0.f : mnew[1];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = expf(s[j][0] - sub0); s[j][1] = expf(s[j][1] - sub0);     // exp(-inf) = 0 for masked keys
      s[j][2] = expf(s[j][2] - sub1); s[j][3] = expf(s[j][3] - sub1);
      ls[0] += s[j][0] + s[j][1];
      ls[1] += s[j][2] + s[j][3];
    }
    // ---- O += P V: the score tiles 2kk, 2kk+1 are exactly the A fragment of k step kk ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (2 * kk < ntiles) {
        uint32_t p_hi[4], p_mid[4];
        split2(s[2 * kk][0], s[2 * kk][1], p_hi[0], p_mid[0]);
        split2(s[2 * kk][2], s[2 * kk][3], p_hi[1], p_mid[1]);
        split2(s[2 * kk + 1][0], s[2 * kk + 1][1], p_hi[2], p_mid[2]);
        split2(s[2 * kk + 1][2], s[2 * kk + 1][3], p_hi[3], p_mid[3]);
#pragma unroll
        for (int dp = 0; dp < 2; ++dp) {                 // output dims 16dp .. 16dp+15 (two 8-wide tiles)
          uint32_t vh[4], vm[4];
          const uint32_t off = (uint32_t)((kk * 16 * LDS_ROW + dp * 16) * 2);
          ldsm_x4_t(v_hi_b + off, vh);
          ldsm_x4_t(v_mid_b + off, vm);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            mma_bf16(o[2 * dp + q], p_mid, vh[2 * q], vh[2 * q + 1]);
            mma_bf16(o[2 * dp + q], p_hi, vm[2 * q], vm[2 * q + 1]);
            mma_bf16(o[2 * dp + q], p_hi, vh[2 * q], vh[2 * q + 1]);
          }
        }
      }
    }
  }

template <int NW>
__global__ void __launch_bounds__(NW * 32)
attn_mma_kernel(AttnArgs p) {
  constexpr int QT = QW * NW;      // queries per CTA
  constexpr int TPK = NW / 2;      // staging threads per key (each converts 32 / TPK dims of K and of V)
  constexpr int F4 = 8 / TPK;      // float4 per thread and tensor
  __shared__ Smem sm;
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.z, h = blockIdx.y, l0 = blockIdx.x * QT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const long long q0 = seq_row0(p.qm, n), k0 = seq_row0(p.km, n);
  const uint8_t* mrow = p.kpm ? p.kpm + (long long)(n / p.kpm_div) * p.S : nullptr;
  const int lw = l0 + warp * QW;                         // first query of this warp
  const bool warp_active = lw < p.L;

  // ---- Q fragments (A operand, rows g and g+8, k = 2t,2t+1 (+8) of each 16-wide k step), pre-scaled ----
  uint32_t q_hi[2][4], q_mid[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int row = lw + g + (e & 1) * 8, col = ks * 16 + (e >> 1) * 8 + 2 * t;
      float2 v = make_float2(0.f, 0.f);
      if (row < p.L) v = __ldg(reinterpret_cast<const float2*>(p.q + (q0 + (long long)row * p.qm.step) * p.ldq + h * D + col));
      split2(v.x * p.scale, v.y * p.scale, q_hi[ks][e], q_mid[ks][e]);
    }

  // chunk staging: thread -> (key = tid / TPK, dims (tid % TPK) * 4 * F4 ..): F4 float4 of K and of V
  const int skey = tid / TPK, sd0 = (tid % TPK) * 4 * F4;
  float4 pk[F4], pv[F4];
  auto fetch = [&](int c) {
    const int s = c * KC + skey;
    if (s < p.S) {
      const long long krow = k0 + (long long)s * p.km.step;
      const float4* kp = reinterpret_cast<const float4*>(p.k + krow * p.ldk + h * D + sd0);
      const float4* vp = reinterpret_cast<const float4*>(p.v + krow * p.ldv + h * D + sd0);
#pragma unroll
      for (int i = 0; i < F4; ++i) { pk[i] = __ldg(kp + i); pv[i] = __ldg(vp + i); }
    } else {
#pragma unroll
      for (int i = 0; i < F4; ++i) { pk[i] = make_float4(0.f, 0.f, 0.f, 0.f); pv[i] = pk[i]; }
    }
  };
  auto stage = [&](int c) {
#pragma unroll
    for (int i = 0; i < F4; ++i) {
      store_split4(&sm.k_hi[skey][sd0 + 4 * i], &sm.k_mid[skey][sd0 + 4 * i], pk[i]);
      store_split4(&sm.v_hi[skey][sd0 + 4 * i], &sm.v_mid[skey][sd0 + 4 * i], pv[i]);
    }
    if (tid < KC) {
      const int s = c * KC + tid;
      sm.msk[tid] = (s >= p.S || (mrow && mrow[s])) ? -INFINITY : 0.f;
    }
  };

  float o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[i][e] = 0.f;
  float mx[2] = {-INFINITY, -INFINITY}, ls[2] = {0.f, 0.f};

  const int nchunks = (p.S + KC - 1) / KC;
  fetch(0);
  for (int c = 0; c < nchunks; ++c) {
    __syncthreads();                                     // the previous chunk's readers are done
    stage(c);
    __syncthreads();
    if (c + 1 < nchunks) fetch(c + 1);                   // in flight while this chunk is computed
    if (!warp_active) continue;
    process_chunk(sm, min(KC, p.S - c * KC), q_hi, q_mid, o, mx, ls, lane);
  }
  if (!warp_active) return;
  // ---- normalise and store (fp32 and / or split) ----
  const long long o0 = seq_row0(p.om, n);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = ls[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const int row = lw + g + r * 8;
    if (row >= p.L) continue;
    const float inv = 1.f / l;                           // every key masked: 0 / 0 = NaN, as the reference's softmax
    const long long orow = o0 + (long long)row * p.om.step;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float x = o[i][2 * r] * inv, y = o[i][2 * r + 1] * inv;
      const int col = h * D + i * 8 + 2 * t;
      if (p.o_f32) *reinterpret_cast<float2*>(p.o_f32 + orow * p.ldo + col) = make_float2(x, y);
      if (p.o_split) {
        uint32_t hi, mid;
        split2(x, y, hi, mid);
        __nv_bfloat16* hp = split_hi(p.o_split, orow, p.ldo);
        *reinterpret_cast<uint32_t*>(hp + col) = hi;
        *reinterpret_cast<uint32_t*>(hp + p.ldo + col) = mid;
      }
    }
  }
}

// L <= 16 (the DETR decoder's cross-attention: 15 tubelet queries over the T'H'W' memory tokens, transformer.py:232-241): one query
// tile per (sequence, head), so instead of one warp walking all key chunks in turn while three warps only help staging, every
// warp takes its own chunks (warp w: chunks w, w+4, ...) with a private staging area and no CTA-wide barriers; the partial
// (max, sum, O) triples are merged through shared memory at the end.  S = 256: one chunk latency instead of four.
constexpr int SPLIT_NW = 4;
__global__ void __launch_bounds__(SPLIT_NW * 32)
attn_mma_split_kernel(AttnArgs p) {
  extern __shared__ __align__(16) uint8_t split_smem[];
  Smem* sms = reinterpret_cast<Smem*>(split_smem);
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.z, h = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const long long q0 = seq_row0(p.qm, n), k0 = seq_row0(p.km, n);
  const uint8_t* mrow = p.kpm ? p.kpm + (long long)(n / p.kpm_div) * p.S : nullptr;
  Smem& sm = sms[warp];

  uint32_t q_hi[2][4], q_mid[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int row = g + (e & 1) * 8, col = ks * 16 + (e >> 1) * 8 + 2 * t;
      float2 v = make_float2(0.f, 0.f);
      if (row < p.L) v = __ldg(reinterpret_cast<const float2*>(p.q + (q0 + (long long)row * p.qm.step) * p.ldq + h * D + col));
      split2(v.x * p.scale, v.y * p.scale, q_hi[ks][e], q_mid[ks][e]);
    }
  float o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[i][e] = 0.f;
  float mx[2] = {-INFINITY, -INFINITY}, ls[2] = {0.f, 0.f};

  const int nchunks = (p.S + KC - 1) / KC;
  for (int c = warp; c < nchunks; c += SPLIT_NW) {
    __syncwarp();
    // this warp stages its chunk: lane -> keys lane and lane + 32, all 32 dims of K and V
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int key = lane + 32 * half, sidx = c * KC + key;
      if (sidx < p.S) {
        const long long krow = k0 + (long long)sidx * p.km.step;
        const float4* kp = reinterpret_cast<const float4*>(p.k + krow * p.ldk + h * D);
        const float4* vp = reinterpret_cast<const float4*>(p.v + krow * p.ldv + h * D);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          store_split4(&sm.k_hi[key][4 * i], &sm.k_mid[key][4 * i], __ldg(kp + i));
          store_split4(&sm.v_hi[key][4 * i], &sm.v_mid[key][4 * i], __ldg(vp + i));
        }
      } else {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          store_split4(&sm.k_hi[key][4 * i], &sm.k_mid[key][4 * i], z);
          store_split4(&sm.v_hi[key][4 * i], &sm.v_mid[key][4 * i], z);
        }
      }
      sm.msk[key] = (sidx >= p.S || (mrow && mrow[sidx])) ? -INFINITY : 0.f;
    }
    __syncwarp();
    process_chunk(sm, min(KC, p.S - c * KC), q_hi, q_mid, o, mx, ls, lane);
  }
  // ---- merge the warps' partial results (rows g and g + 8 of each warp's quad lanes) ----
  __syncthreads();                                           // staging areas are free: reuse them as float scratch
  float* scratch = reinterpret_cast<float*>(split_smem);      // [warp][row 16][m, l, o[32]] = 34 floats per row
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = ls[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    float* row = scratch + ((warp * 16) + g + 8 * r) * 36;
    if (t == 0) { row[0] = mx[r]; row[1] = l; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { row[2 + i * 8 + 2 * t] = o[i][2 * r]; row[2 + i * 8 + 2 * t + 1] = o[i][2 * r + 1]; }
  }
  __syncthreads();
  // thread -> (query row = tid / 8, 4 output dims = (tid % 8) * 4)
  const int row = tid >> 3, d0 = (tid & 7) * 4;
  if (row >= p.L) return;
  float m = -INFINITY;
#pragma unroll
  for (int w = 0; w < SPLIT_NW; ++w) m = fmaxf(m, scratch[(w * 16 + row) * 36]);
  float l = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int w = 0; w < SPLIT_NW; ++w) {
    const float* pr = scratch + (w * 16 + row) * 36;
    const float f = (pr[0] == -INFINITY) ? 0.f : expf(pr[0] - m);   // a warp without chunks / fully masked chunks contributes nothing
    l += pr[1] * f;
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] += pr[2 + d0 + e] * f;
  }
  const float inv = 1.f / l;                                 // every key masked: NaN, as the reference's softmax
  const long long orow = seq_row0(p.om, n) + (long long)row * p.om.step;
  const int col = h * D + d0;
  const float4 res = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
  if (p.o_f32) *reinterpret_cast<float4*>(p.o_f32 + orow * p.ldo + col) = res;
  if (p.o_split) {
    __nv_bfloat16* hp = split_hi(p.o_split, orow, p.ldo);
    store_split4(hp + col, hp + p.ldo + col, res);
  }
}

// Tiny sequences (L, S <= 8; the per-pixel temporal attention, transformer_layers.py:86-91): one warp per (sequence,
// pair of heads); half-warp = head, lane = (query slot = 0..?)...  Kept simple: 4 lanes share a (head, query) dot product.
// A warp handles one sequence and all heads in turn: lane = (head-local dim group), see below.
//   lane l: dims 8*(l&3) .. +7 of head (l >> 2) + 8 * pass  -> with H = 8 heads one pass covers all 256 channels.
template <int MAXT>
__global__ void __launch_bounds__(128)
attn_tiny_kernel(AttnArgs p) {
  pdl_trigger();
  pdl_wait();
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int hgroups = p.H / 8;                           // 8 heads per warp pass
  if (w >= (long long)p.NB * hgroups) return;
  const int n = (int)(w / hgroups), h = (int)(w % hgroups) * 8 + (lane >> 2), d0 = (lane & 3) * 8;
  const long long q0 = seq_row0(p.qm, n), k0 = seq_row0(p.km, n), o0 = seq_row0(p.om, n);
  const uint8_t* mrow = p.kpm ? p.kpm + (long long)(n / p.kpm_div) * p.S : nullptr;
  float k[MAXT][8], v[MAXT][8];
#pragma unroll
  for (int j = 0; j < MAXT; ++j) {
    if (j < p.S) {
      const long long krow = k0 + (long long)j * p.km.step;
      const float4* kp = reinterpret_cast<const float4*>(p.k + krow * p.ldk + h * D + d0);
      const float4* vp = reinterpret_cast<const float4*>(p.v + krow * p.ldv + h * D + d0);
      const float4 a = __ldg(kp), b = __ldg(kp + 1), c = __ldg(vp), d = __ldg(vp + 1);
      k[j][0] = a.x; k[j][1] = a.y; k[j][2] = a.z; k[j][3] = a.w; k[j][4] = b.x; k[j][5] = b.y; k[j][6] = b.z; k[j][7] = b.w;
      v[j][0] = c.x; v[j][1] = c.y; v[j][2] = c.z; v[j][3] = c.w; v[j][4] = d.x; v[j][5] = d.y; v[j][6] = d.z; v[j][7] = d.w;
    }
  }
#pragma unroll
  for (int i = 0; i < MAXT; ++i) {
    if (i < p.L) {
      const float4* qp = reinterpret_cast<const float4*>(p.q + (q0 + (long long)i * p.qm.step) * p.ldq + h * D + d0);
      const float4 a = __ldg(qp), b = __ldg(qp + 1);
      const float q[8] = {a.x * p.scale, a.y * p.scale, a.z * p.scale, a.w * p.scale, b.x * p.scale, b.y * p.scale, b.z * p.scale, b.w * p.scale};
      float sc[MAXT], m = -INFINITY;
#pragma unroll
      for (int j = 0; j < MAXT; ++j) {
        sc[j] = -INFINITY;
        if (j < p.S) {
          float dot = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) dot = fmaf(q[e], k[j][e], dot);
          dot += __shfl_xor_sync(0xffffffffu, dot, 1);
          dot += __shfl_xor_sync(0xffffffffu, dot, 2);
          sc[j] = (mrow && mrow[j]) ? -INFINITY : dot;
          m = fmaxf(m, sc[j]);
        }
      }
      float l = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < MAXT; ++j) {
        if (j < p.S) {
          const float pr = expf(sc[j] - m);              // all keys masked: exp(nan) -> NaN, as the reference
          l += pr;
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(pr, v[j][e], acc[e]);
        }
      }
      const float inv = 1.f / l;
      const long long orow = o0 + (long long)i * p.om.step;
      const int col = h * D + d0;
      if (p.o_f32) {
        float4* op = reinterpret_cast<float4*>(p.o_f32 + orow * p.ldo + col);
        op[0] = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
        op[1] = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
      }
      if (p.o_split) {
        __nv_bfloat16* hp = split_hi(p.o_split, orow, p.ldo);
        store_split4(hp + col, hp + p.ldo + col, make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv));
        store_split4(hp + col + 4, hp + p.ldo + col + 4, make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv));
      }
    }
  }
}

}  // namespace attn_mma

// head_dim 32 only; returns cudaErrorNotSupported for anything the two kernels do not cover (the caller falls back
// to the CUDA-core kernels of kernels_simt.cu)
cudaError_t launch_attention_mma(const AttnArgs& a, cudaStream_t st) {
  using namespace attn_mma;
  if (a.D != D || a.NB <= 0 || a.L <= 0 || a.S <= 0) return cudaErrorNotSupported;
  if (a.ldq % 4 || a.ldk % 4 || a.ldv % 4 || a.ldo % 4) return cudaErrorNotSupported;
  if (a.L <= 8 && a.S <= 8 && a.H % 8 == 0) {
    const long long warps = (long long)a.NB * (a.H / 8);
    return launch_pdl(attn_tiny_kernel<8>, dim3(ceil_div(warps * 32, 128)), dim3(128), 0, st, a);
  }
  if (a.NB > 65535 || a.H > 65535) return cudaErrorNotSupported;
  if (a.L <= QW && a.S > KC) {                               // one query tile, several key chunks: split the keys over the warps
    static DeviceOnce once;
    const size_t smem = SPLIT_NW * sizeof(Smem);
    cudaError_t e = once.run([smem]() { return cudaFuncSetAttribute(attn_mma_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
    if (e != cudaSuccess) return e;
    return launch_pdl(attn_mma_split_kernel, dim3(1, a.H, a.NB), dim3(SPLIT_NW * 32), smem, st, a);
  }
  if (a.L > 64) {
    dim3 grid(ceil_div(a.L, 8 * QW), a.H, a.NB);
    return launch_pdl(attn_mma_kernel<8>, grid, dim3(8 * 32), 0, st, a);
  } else {
    dim3 grid(ceil_div(a.L, 4 * QW), a.H, a.NB);
    return launch_pdl(attn_mma_kernel<4>, grid, dim3(4 * 32), 0, st, a);
  }
}
