// The DETR decoder stack (reference models/transformer/transformer.py:49-64 forward, :98-127 TransformerDecoder,
// :218-249 TransformerDecoderLayer.forward_post) for the tubelet-query sizes of the AVA configurations as ONE persistent
// cooperative kernel.
//
// With Q = 15 tubelet queries a decoder layer works on B*Q = 120 rows of 256 values: every one of its 11 launches
// (projections, two attentions, three LayerNorms, the feed-forward pair) was a few microseconds of dependent latency
// with almost no work -- 60 launches and 0.55 ms per forward for ~1 GFLOP.  Here the whole stack runs as a sequence of
// phases separated by grid barriers; the state never leaves fp32 and the arithmetic is plain fp32 FMA on the CUDA
// cores (the work is far too small for tensor-core tiles to pay), so the results are at least as close to the
// reference's fp32 ops as the split-bf16 GEMM sequence it replaces.
//
//   per layer i (layer 0 starts from tgt = 0: its self-attention block and cross-attention query were folded at
//   tuber_plan_finalize into dec0_c1 / dec0_qc, plan.cu):
//     A   qkv = tgt Win^T + bin + query_pos terms            columns distributed over the CTAs, rows staged in shared memory
//     --- grid barrier ---
//     BC  one CTA per row (b, q) (RB rows of a clip per CTA; RB = 1 measured best: 24 vs 30 us with 3): self-attention over the clip's Q rows (warp = head) -> out_proj + residual -> norm1
//         -> cross-attention query -> cross-attention over the clip's memory tokens (warp = head, K / V of all layers
//         were projected by one tcgen05 GEMM before the kernel) -> out_proj + residual -> norm2
//     --- grid barrier ---
//     D   feed-forward: CTA c owns 16 hidden units: h = relu(tgt W1_c^T + b1_c), partial_c = h W2[:, c]^T   (all rows)
//     --- grid barrier ---
//     E   one CTA per row: sum of the partials in a fixed order (deterministic) + b2 + residual -> norm3 -> tgt;
//         the decoder's shared final norm of it -> hs[b, i, q] (split-bf16: operand of the head GEMMs)
//     --- grid barrier ---
//
// Weights stay fp32: row-major [N, K] where a CTA reads whole rows of a slice (A, D), K-major copies [K, N] made at
// finalize where a thread owns output columns of a matrix-vector product (BC).
#include "kernels.h"

namespace dmk {

constexpr int D = 256, NH = 8, HD = 32, THREADS = 256;
constexpr int XS = 260;          // shared-memory row stride of staged rows (floats): conflict-free LDS.128 for 8 consecutive rows
constexpr int RCH = 120;         // rows staged per chunk
constexpr int RG = RCH / 8;      // row groups of the feed-forward register tile (a thread owns rows rg + RG * i, i < 8)
constexpr int FU = 16;           // hidden units per feed-forward slice
constexpr int ACOLS = 6;         // qkv columns per phase-A work item (768 / 6 = 128 items)
constexpr int OFF_X = 0;                         // [RCH][XS]
constexpr int OFF_W = OFF_X + RCH * XS;          // [FU][XS]  (phase A uses ACOLS rows of it)
constexpr int OFF_H = OFF_W + FU * XS;           // [RCH][FU]
// row-local phases alias the X region: vectors [att | t1 | qc | part(4 x) | red] for RB rows, then the score rows [RB][8][ntok_pad];
// the two 64 KB weight buffers of the streamed matrix-vector products sit at the top of the allocation
constexpr int RB = 1;             // rows of a clip per CTA in the row-local phase BC
constexpr int V_ATT = 0, V_T1 = RB * 256, V_QC = 2 * RB * 256, V_PART = 3 * RB * 256, V_RED = 7 * RB * 256, V_SC = V_RED + 64;
constexpr int WCH_ROWS = 64, WCH_FLOATS = WCH_ROWS * D, WCH_BYTES = WCH_FLOATS * 4;   // one chunk of a K-major weight: 64 k-rows
constexpr int SMEM_FLOATS = 57600;               // 225 KB
constexpr int OFF_BAR = SMEM_FLOATS - 8;         // two mbarriers (8-byte aligned)
constexpr int OFF_WB = OFF_BAR - 2 * WCH_FLOATS; // 16-byte aligned
static_assert(OFF_H + RCH * FU <= OFF_BAR, "staged operands overlap the barriers");   // (they may overlap the weight buffers: other phases)
static_assert((OFF_WB * 4) % 16 == 0 && (OFF_BAR * 4) % 8 == 0, "alignment");
static_assert(SMEM_FLOATS * 4 <= 232448, "over the 227 KB shared-memory limit");
static_assert(4 * RCH * FU <= RCH * XS, "feed-forward k-split partials live in the X region");

TB_DEVINL float ldcg(const float* p) { return __ldcg(p); }
TB_DEVINL float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
TB_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
TB_DEVINL void cp_async16(float* dst, const float* src) {    // L2 -> shared memory, no register staging, bypasses L1
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
TB_DEVINL void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
TB_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n"
      ::"r"(bar), "r"(parity) : "memory");
}

// grid barrier: monotonic arrival counter (zeroed by the launcher), every CTA waits for nblocks * (barrier index) arrivals
TB_DEVINL void grid_sync(unsigned* ctr, unsigned& target, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblocks;
    __threadfence();                                         // this CTA's writes (ordered before by the barrier above) -> gpu scope
    atomicAdd(ctr, 1u);
    unsigned seen, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
      if (++spins > (1u << 28)) __trap();                    // never reached when all CTAs are resident (cooperative launch)
    } while (seen < target);
    __threadfence();
  }
  __syncthreads();
}

// sums over the CTA of N values per thread; red: [N][8] floats
template <int N>
TB_DEVINL void block_sum(float (&v)[N], float* red) {
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();                                           // previous readers of red are done
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) red[i * 8 + (threadIdx.x >> 5)] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) s += red[i * 8 + w];
    v[i] = s;
  }
}
// LayerNorm over the 256 values held one per thread, N independent rows at once (in place)
template <int N>
TB_DEVINL void block_layernorm(float (&y)[N], const float* g, const float* b, float eps, float* red) {
  float t[N];
#pragma unroll
  for (int i = 0; i < N; ++i) t[i] = y[i];
  block_sum(t, red);
  float dlt[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { dlt[i] = y[i] - t[i] * (1.f / D); t[i] = dlt[i] * dlt[i]; }
  block_sum(t, red);
  const float gg = __ldg(g + threadIdx.x), bb = __ldg(b + threadIdx.x);
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = dlt[i] * (1.f / sqrtf(t[i] * (1.f / D) + eps)) * gg + bb;
}

// ---- y = x Wt for a K-major fp32 weight Wt [256][256], streamed through two 64 KB shared-memory buffers with bulk async
// copies (4 chunks of 64 k-rows).  A per-thread global-load loop left this at ~4 us per product (latency bound at 8 warps);
// the copy engine keeps the whole chunk in flight.  mv_prefetch may be called as soon as the previous product has finished
// (its first two chunks then load behind whatever the CTA does in between).
struct MatVec {
  float* wb; uint32_t bar0; uint32_t phase0, phase1;
  TB_DEVINL void issue(const float* Wt, int chunk) const {    // one thread
    const uint32_t bar = bar0 + 8u * (chunk & 1);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)WCH_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(wb + (chunk & 1) * WCH_FLOATS)), "l"(Wt + (size_t)chunk * WCH_FLOATS), "r"((uint32_t)WCH_BYTES), "r"(bar) : "memory");
  }
  TB_DEVINL void prefetch(const float* Wt) const {
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy use of the buffers is ordered before the copies
      issue(Wt, 0);
      issue(Wt, 1);
    }
  }
  // x: [RB][256] floats in shared memory -> y[rr] = (x[rr] Wt)[threadIdx.x]; `part`: [4][RB][256] scratch.
  // thread = (4 columns, a quarter of each chunk); every weight value read from shared memory serves the RB rows
  TB_DEVINL void run(const float* x, const float* Wt, float* part, float (&y)[RB]) {
    const int n4 = threadIdx.x & 63, kq = threadIdx.x >> 6;
    float4 acc[RB];
#pragma unroll
    for (int rr = 0; rr < RB; ++rr) acc[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int c = 0; c < D / WCH_ROWS; ++c) {
      const int b = c & 1;
      mbar_wait(bar0 + 8u * b, b ? phase1 : phase0);
      if (b) phase1 ^= 1u; else phase0 ^= 1u;
      const float4* wp = reinterpret_cast<const float4*>(wb + b * WCH_FLOATS) + (kq * 16) * 64 + n4;
      const float* xp = x + c * WCH_ROWS + kq * 16;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float4 w = wp[k * 64];
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) {
          const float xv = xp[rr * D + k];
          acc[rr].x = fmaf(xv, w.x, acc[rr].x); acc[rr].y = fmaf(xv, w.y, acc[rr].y);
          acc[rr].z = fmaf(xv, w.z, acc[rr].z); acc[rr].w = fmaf(xv, w.w, acc[rr].w);
        }
      }
      __syncthreads();                                       // every reader is done with buffer b
      if (c + 2 < D / WCH_ROWS && threadIdx.x == 0) issue(Wt, c + 2);
    }
#pragma unroll
    for (int rr = 0; rr < RB; ++rr)                            // (previous readers of part passed the barriers above)
      *reinterpret_cast<float4*>(part + (kq * RB + rr) * 256 + n4 * 4) = acc[rr];
    __syncthreads();
    const int n = threadIdx.x;
#pragma unroll
    for (int rr = 0; rr < RB; ++rr)
      y[rr] = (part[rr * 256 + n] + part[(RB + rr) * 256 + n]) + (part[(2 * RB + rr) * 256 + n] + part[(3 * RB + rr) * 256 + n]);
  }
};

// rows [r0, r0 + nr) of the fp32 state -> shared memory (asynchronous; cp_async_wait_all + barrier before use)
TB_DEVINL void stage_rows(float* X, const float* tgt, int r0, int nr) {
  for (int i = threadIdx.x; i < nr * (D / 4); i += THREADS) {
    const int r = i >> 6, k4 = i & 63;
    cp_async16(X + r * XS + k4 * 4, tgt + (size_t)(r0 + r) * D + k4 * 4);
  }
}

}  // namespace dmk

__global__ void __launch_bounds__(dmk::THREADS, 1) decoder_mega_kernel(const __grid_constant__ DecMegaArgs a) {
  using namespace dmk;
  extern __shared__ __align__(128) float dsm[];
  float* X = dsm + OFF_X;
  float* Ws = dsm + OFF_W;
  float* Hs = dsm + OFF_H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, Mq = a.B * a.Q, Q = a.Q, Ntok = a.Ntok;
  const int nslices = a.dim_ff / FU;
  unsigned target = 0;
  const float scale = 1.f / sqrtf((float)HD);
  MatVec mv;
  mv.wb = dsm + OFF_WB;
  mv.bar0 = smem_u32(dsm + OFF_BAR);
  mv.phase0 = mv.phase1 = 0u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mv.bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mv.bar0 + 8u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int tr = 0;
  auto stamp = [&]() {                                       // per-phase profile (tuber_set_kernel_profiling)
    if (a.trace && blockIdx.x == 0 && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.trace[tr] = t;
    }
    ++tr;
  };
  stamp();
  int tr2 = 64;
  auto sub = [&]() {                                         // finer stamps of CTA 0's own work inside a phase (debug)
    if (a.trace && blockIdx.x == 0 && tid == 0 && tr2 < 256) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.trace[tr2] = t;
    }
    ++tr2;
  };

  for (int li = 0; li < a.Ld; ++li) {
    const DecMegaLayer& L = a.L[li];
    const bool folded = li == 0 && a.dec0_c1 != nullptr;

    // ================= A: qkv = tgt Win^T + bin + (query_pos terms) =================
    if (!folded) {
      for (int item = blockIdx.x; item < 3 * D / ACOLS; item += G) {
        const int c0 = item * ACOLS;
        __syncthreads();
        for (int i = tid; i < ACOLS * (D / 4); i += THREADS) {
          const int c = i >> 6, k4 = i & 63;
          cp_async16(Ws + c * XS + k4 * 4, L.sa_in_w + (size_t)(c0 + c) * D + k4 * 4);
        }
        for (int r0 = 0; r0 < Mq; r0 += RCH) {
          const int nr = min(RCH, Mq - r0);
          __syncthreads();
          stage_rows(X, a.tgt, r0, nr);
          cp_async_wait_all();
          __syncthreads();
          const int r = tid & 127, half = tid >> 7;
          if (r < nr) {
            float acc[3] = {0.f, 0.f, 0.f};
            const float* xr = X + r * XS;
            const float* wr = Ws + (half * 3) * XS;
#pragma unroll 4
            for (int k4 = 0; k4 < D / 4; ++k4) {
              const float4 x = *reinterpret_cast<const float4*>(xr + k4 * 4);
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const float4 w = *reinterpret_cast<const float4*>(wr + c * XS + k4 * 4);
                acc[c] = fmaf(x.x, w.x, acc[c]); acc[c] = fmaf(x.y, w.y, acc[c]);
                acc[c] = fmaf(x.z, w.z, acc[c]); acc[c] = fmaf(x.w, w.w, acc[c]);
              }
            }
            const int row = r0 + r, q = row % Q;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const int col = c0 + half * 3 + c;
              __stcg(a.qkv + (size_t)row * 3 * D + col, acc[c] + __ldg(L.sa_in_b + col) + __ldg(L.pq_sa + (size_t)q * 3 * D + col));
            }
          }
        }
      }
      grid_sync(a.barrier, target, G);
    }
    stamp();

    // ================= BC: one CTA per group of RB rows of a clip (weights and the clip's K / V are read once per group) ====
    {
      const int gpc = (Q + RB - 1) / RB;                       // groups per clip
      for (int grp = blockIdx.x; grp < a.B * gpc; grp += G) {
        const int b = grp / gpc, q0 = (grp % gpc) * RB;
        float* v_att = X + V_ATT;                              // [RB][256]
        float* v_t1 = X + V_T1;
        float* v_qc = X + V_QC;
        float* v_part = X + V_PART;                            // [4][RB][256]
        float* v_red = X + V_RED;
        float* v_sc = X + V_SC + warp * a.ntok_pad;            // row rr: + rr * NH * ntok_pad
        const int sc_row = NH * a.ntok_pad;
        __syncthreads();
        sub();
        mv.prefetch(folded ? L.ca_out_t : L.sa_out_t);
        float y[RB];
        if (!folded) {
          // ---- self-attention over the clip's Q rows: warp = head, lane = key (scores) / dim (output); the clip's K / V rows
          // are loaded once for the RB query rows of the group ----
          const int h = warp;
          float kk[HD];
          if (lane < Q) {
            const float* kp = a.qkv + (size_t)(b * Q + lane) * 3 * D + D + h * HD;
#pragma unroll
            for (int i = 0; i < HD / 4; ++i) {
              const float4 t = ldcg4(kp + 4 * i);
              kk[4 * i] = t.x; kk[4 * i + 1] = t.y; kk[4 * i + 2] = t.z; kk[4 * i + 3] = t.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < HD; ++i) kk[i] = 0.f;
          }
          const float* vcol = a.qkv + (size_t)(b * Q) * 3 * D + 2 * D + h * HD + lane;
          float vv[32];                                        // V rows of the clip, this lane's dim (Q <= 32)
#pragma unroll
          for (int j = 0; j < 32; ++j) vv[j] = j < Q ? ldcg(vcol + (size_t)j * 3 * D) : 0.f;
#pragma unroll
          for (int rr = 0; rr < RB; ++rr) {
            const int q = min(q0 + rr, Q - 1);                 // (a group's tail rows repeat the last query; never stored)
            const float qv = ldcg(a.qkv + (size_t)(b * Q + q) * 3 * D + h * HD + lane);
            float s = 0.f;
#pragma unroll
            for (int dd = 0; dd < HD; ++dd) s = fmaf(__shfl_sync(0xffffffffu, qv, dd), kk[dd], s);
            s = lane < Q ? s * scale : -INFINITY;
            const float mx = warp_max(s);
            float pr = lane < Q ? expf(s - mx) : 0.f;
            pr /= warp_sum(pr);
            float o = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) o = fmaf(__shfl_sync(0xffffffffu, pr, j), vv[j], o);
            v_att[rr * D + h * HD + lane] = o;
          }
          __syncthreads();
          sub();
          // ---- out_proj + residual -> norm1 ----
          mv.run(v_att, L.sa_out_t, v_part, y);
          mv.prefetch(L.ca_q_t);
#pragma unroll
          for (int rr = 0; rr < RB; ++rr)
            y[rr] += __ldg(L.sa_out_b + tid) + ldcg(a.tgt + (size_t)(b * Q + min(q0 + rr, Q - 1)) * D + tid);
          block_layernorm(y, L.n1_g, L.n1_b, a.eps, v_red);
#pragma unroll
          for (int rr = 0; rr < RB; ++rr) v_t1[rr * D + tid] = y[rr];
          __syncthreads();
          sub();
          // ---- cross-attention query ----
          mv.run(v_t1, L.ca_q_t, v_part, y);
          mv.prefetch(L.ca_out_t);
#pragma unroll
          for (int rr = 0; rr < RB; ++rr)
            v_qc[rr * D + tid] = y[rr] + __ldg(L.ca_q_b + tid) + __ldg(L.pq_ca + (size_t)min(q0 + rr, Q - 1) * D + tid);
        } else {
#pragma unroll
          for (int rr = 0; rr < RB; ++rr) {
            v_t1[rr * D + tid] = __ldg(a.dec0_c1 + tid);
            v_qc[rr * D + tid] = __ldg(a.dec0_qc + (size_t)min(q0 + rr, Q - 1) * D + tid);
          }
        }
        __syncthreads();
        sub();
        {
          // ---- cross-attention over the clip's memory tokens: warp = head; a load instruction covers 4 keys x 128 bytes
          // (lane = key-in-group ks, 16-byte chunk c8 of the head's 32 dims), every K / V value serves the RB query rows ----
          const int h = warp, ks = lane >> 3, c8 = lane & 7;
          float4 q4[RB];
#pragma unroll
          for (int rr = 0; rr < RB; ++rr) q4[rr] = *reinterpret_cast<const float4*>(v_qc + rr * D + h * HD + 4 * c8);
          const float* kbase = a.memkv + (size_t)b * Ntok * a.kv_ld + (size_t)li * 2 * D + h * HD + 4 * c8;
          const uint8_t* km = a.kpm ? a.kpm + (size_t)b * Ntok : nullptr;
          for (int j0 = 0; j0 < Ntok; j0 += 64) {              // 16 loads in flight per lane
            float4 kv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int j = j0 + 4 * i + ks;
              kv[i] = j < Ntok ? __ldg(reinterpret_cast<const float4*>(kbase + (size_t)j * a.kv_ld)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int rr = 0; rr < RB; ++rr) {
              // partial dot products of 16 keys, then a transposing butterfly over the 8 lanes of the key group: every step
              // halves the values a lane carries (14 shuffles instead of 48), lane c8 ends with the sums of keys i = 2 c8, 2 c8 + 1.
              // Branch-free: selects and two predicated stores (a divergent branch per key made this loop issue bound)
              float s[16];
#pragma unroll
              for (int i = 0; i < 16; ++i)
                s[i] = fmaf(q4[rr].x, kv[i].x, fmaf(q4[rr].y, kv[i].y, fmaf(q4[rr].z, kv[i].z, q4[rr].w * kv[i].w)));
              const bool b2 = (c8 & 4) != 0, b1 = (c8 & 2) != 0, b0 = (c8 & 1) != 0;
              float r8[8], r4[4], r2[2];
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                const float mine = b2 ? s[8 + t] : s[t], theirs = b2 ? s[t] : s[8 + t];
                r8[t] = mine + __shfl_xor_sync(0xffffffffu, theirs, 4);
              }
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float mine = b1 ? r8[4 + t] : r8[t], theirs = b1 ? r8[t] : r8[4 + t];
                r4[t] = mine + __shfl_xor_sync(0xffffffffu, theirs, 2);
              }
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                const float mine = b0 ? r4[2 + t] : r4[t], theirs = b0 ? r4[t] : r4[2 + t];
                r2[t] = mine + __shfl_xor_sync(0xffffffffu, theirs, 1);
              }
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                const int j = j0 + 4 * (2 * c8 + t) + ks;
                if (j < Ntok) v_sc[rr * sc_row + j] = r2[t] * scale;
              }
            }
          }
          sub();
          __syncwarp();
          if (km) {                                            // key padding mask (1 = ignore the key)
            for (int j = lane; j < Ntok; j += 32)
              if (km[j]) {
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) v_sc[rr * sc_row + j] = -INFINITY;
              }
            __syncwarp();
          }
          float inv[RB];
#pragma unroll
          for (int rr = 0; rr < RB; ++rr) {
            float m = -INFINITY;
            for (int j = lane; j < Ntok; j += 32) m = fmaxf(m, v_sc[rr * sc_row + j]);
            m = warp_max(m);
            float sum = 0.f;
            for (int j = lane; j < Ntok; j += 32) {
              const float pj = expf(v_sc[rr * sc_row + j] - m);
              v_sc[rr * sc_row + j] = pj;
              sum += pj;
            }
            inv[rr] = 1.f / warp_sum(sum);
          }
          __syncwarp();
          const float* vbase = kbase + D;
          float4 o4[RB];
#pragma unroll
          for (int rr = 0; rr < RB; ++rr) o4[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int j0 = 0; j0 < Ntok; j0 += 64) {
            float4 vv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int j = j0 + 4 * i + ks;
              vv[i] = j < Ntok ? __ldg(reinterpret_cast<const float4*>(vbase + (size_t)j * a.kv_ld)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int j = j0 + 4 * i + ks;
#pragma unroll
              for (int rr = 0; rr < RB; ++rr) {
                const float pj = j < Ntok ? v_sc[rr * sc_row + j] : 0.f;
                o4[rr].x = fmaf(pj, vv[i].x, o4[rr].x); o4[rr].y = fmaf(pj, vv[i].y, o4[rr].y);
                o4[rr].z = fmaf(pj, vv[i].z, o4[rr].z); o4[rr].w = fmaf(pj, vv[i].w, o4[rr].w);
              }
            }
          }
#pragma unroll
          for (int rr = 0; rr < RB; ++rr) {                    // the four key groups of the warp hold partial sums of the same dims
            o4[rr].x += __shfl_xor_sync(0xffffffffu, o4[rr].x, 8);  o4[rr].y += __shfl_xor_sync(0xffffffffu, o4[rr].y, 8);
            o4[rr].z += __shfl_xor_sync(0xffffffffu, o4[rr].z, 8);  o4[rr].w += __shfl_xor_sync(0xffffffffu, o4[rr].w, 8);
            o4[rr].x += __shfl_xor_sync(0xffffffffu, o4[rr].x, 16); o4[rr].y += __shfl_xor_sync(0xffffffffu, o4[rr].y, 16);
            o4[rr].z += __shfl_xor_sync(0xffffffffu, o4[rr].z, 16); o4[rr].w += __shfl_xor_sync(0xffffffffu, o4[rr].w, 16);
            if (ks == 0)
              *reinterpret_cast<float4*>(v_att + rr * D + h * HD + 4 * c8) =
                  make_float4(o4[rr].x * inv[rr], o4[rr].y * inv[rr], o4[rr].z * inv[rr], o4[rr].w * inv[rr]);
          }
        }
        __syncthreads();
        sub();
        // ---- out_proj + residual -> norm2 -> state ----
        mv.run(v_att, L.ca_out_t, v_part, y);
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) y[rr] += __ldg(L.ca_out_b + tid) + v_t1[rr * D + tid];
        block_layernorm(y, L.n2_g, L.n2_b, a.eps, v_red);
#pragma unroll
        for (int rr = 0; rr < RB; ++rr)
          if (q0 + rr < Q) __stcg(a.tgt + (size_t)(b * Q + q0 + rr) * D + tid, y[rr]);
        sub();
      }
    }
    grid_sync(a.barrier, target, G);
    stamp();

    // ================= D: feed-forward partial sums, 16 hidden units per work item =================
    for (int c = blockIdx.x; c < nslices; c += G) {
      __syncthreads();
      for (int i = tid; i < FU * (D / 4); i += THREADS) {
        const int u = i >> 6, k4 = i & 63;
        cp_async16(Ws + u * XS + k4 * 4, L.lin1_w + (size_t)(c * FU + u) * D + k4 * 4);
      }
      float w2[FU];                                            // this thread's output column: W2[tid][c*16 .. c*16+16)
#pragma unroll
      for (int i = 0; i < FU / 4; ++i) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(L.lin2_w + (size_t)tid * a.dim_ff + c * FU) + i);
        w2[4 * i] = t.x; w2[4 * i + 1] = t.y; w2[4 * i + 2] = t.z; w2[4 * i + 3] = t.w;
      }
      for (int r0 = 0; r0 < Mq; r0 += RCH) {
        const int nr = min(RCH, Mq - r0);
        __syncthreads();
        sub();
        stage_rows(X, a.tgt, r0, nr);
        cp_async_wait_all();
        __syncthreads();
        sub();
        // ---- h = relu(x W1_c^T + b1_c): register tile 8 rows x 4 units x a quarter of K per thread (240 threads),
        // 12 LDS.128 per 128 FMAs; rows rg + 15 i keep the 8 row loads of a quarter-warp on different banks ----
        float acc[8][4];
        const int kq = tid / (4 * RG), ug = (tid / RG) & 3, rg = tid % RG;
        if (tid < 16 * RG) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[i][u] = 0.f;
          const float* xr = X + rg * XS + kq * 64;
          const float* wr = Ws + (ug * 4) * XS + kq * 64;
#pragma unroll 2
          for (int k4 = 0; k4 < 16; ++k4) {
            float4 xv[8], wv[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) xv[i] = *reinterpret_cast<const float4*>(xr + (size_t)(RG * i) * XS + k4 * 4);
#pragma unroll
            for (int u = 0; u < 4; ++u) wv[u] = *reinterpret_cast<const float4*>(wr + u * XS + k4 * 4);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                acc[i][u] = fmaf(xv[i].x, wv[u].x, acc[i][u]); acc[i][u] = fmaf(xv[i].y, wv[u].y, acc[i][u]);
                acc[i][u] = fmaf(xv[i].z, wv[u].z, acc[i][u]); acc[i][u] = fmaf(xv[i].w, wv[u].w, acc[i][u]);
              }
          }
        }
        __syncthreads();                                       // every reader of X is done: its space takes the k-split partials
        if (tid < 16 * RG) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(X + ((size_t)kq * RCH + rg + RG * i) * FU + ug * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        __syncthreads();
        for (int i = tid; i < nr * FU; i += THREADS) {
          const int u = i & (FU - 1);
          const float hsum = (X[i] + X[RCH * FU + i]) + (X[2 * RCH * FU + i] + X[3 * RCH * FU + i]);
          Hs[i] = fmaxf(hsum + __ldg(L.lin1_b + c * FU + u), 0.f);
        }
        __syncthreads();
        sub();
        // ---- partial_c = h W2[:, c]^T: thread = output column, 4 rows in flight ----
        float* pout = a.part + ((size_t)c * Mq + r0) * D + tid;
        int r = 0;
        for (; r + 4 <= nr; r += 4) {
          float y[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < FU / 4; ++i) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              const float4 hv = *reinterpret_cast<const float4*>(Hs + (r + rr) * FU + 4 * i);
              y[rr] = fmaf(hv.x, w2[4 * i], y[rr]); y[rr] = fmaf(hv.y, w2[4 * i + 1], y[rr]);
              y[rr] = fmaf(hv.z, w2[4 * i + 2], y[rr]); y[rr] = fmaf(hv.w, w2[4 * i + 3], y[rr]);
            }
          }
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) __stcg(pout + (size_t)(r + rr) * D, y[rr]);
        }
        for (; r < nr; ++r) {
          float y = 0.f;
#pragma unroll
          for (int i = 0; i < FU / 4; ++i) {
            const float4 hv = *reinterpret_cast<const float4*>(Hs + r * FU + 4 * i);
            y = fmaf(hv.x, w2[4 * i], y); y = fmaf(hv.y, w2[4 * i + 1], y); y = fmaf(hv.z, w2[4 * i + 2], y); y = fmaf(hv.w, w2[4 * i + 3], y);
          }
          __stcg(pout + (size_t)r * D, y);
        }
        sub();
      }
    }
    grid_sync(a.barrier, target, G);
    stamp();

    // ================= E: sum of the partials + bias + residual -> norm3 -> state; final norm -> hs =================
    for (int row = blockIdx.x; row < Mq; row += G) {
      const int b = row / Q, q = row % Q;
      float* v_part = X + V_PART;
      float* v_red = X + V_RED;
      __syncthreads();
      {
        // thread = (4 columns, a quarter of the slices): independent 16-byte loads, fixed summation order (deterministic)
        const int n4 = tid & 63, cq = tid >> 6;
        const int per = (nslices + 3) / 4, cbeg = cq * per, cend = min(nslices, cbeg + per);
        const size_t pstride = (size_t)Mq * D;
        const float* pp = a.part + (size_t)row * D + n4 * 4;
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
        int c = cbeg;
        for (; c + 8 <= cend; c += 8) {
          float4 t[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = ldcg4(pp + (size_t)(c + i) * pstride);
#pragma unroll
          for (int i = 0; i < 8; ++i) { s4.x += t[i].x; s4.y += t[i].y; s4.z += t[i].z; s4.w += t[i].w; }
        }
        for (; c < cend; ++c) {
          const float4 t = ldcg4(pp + (size_t)c * pstride);
          s4.x += t.x; s4.y += t.y; s4.z += t.z; s4.w += t.w;
        }
        *reinterpret_cast<float4*>(v_part + cq * 256 + n4 * 4) = s4;
      }
      __syncthreads();
      float y = ((v_part[tid] + v_part[256 + tid]) + (v_part[512 + tid] + v_part[768 + tid])) + __ldg(L.lin2_b + tid) +
                ldcg(a.tgt + (size_t)row * D + tid);
      float y1[1] = {y};
      block_layernorm(y1, L.n3_g, L.n3_b, a.eps, v_red);
      __stcg(a.tgt + (size_t)row * D + tid, y1[0]);
      block_layernorm(y1, a.nf_g, a.nf_b, a.eps, v_red);
      __nv_bfloat16 hi, mid;
      split_bf16(y1[0], hi, mid);
      __nv_bfloat16* hp = split_hi(a.hs, ((long long)b * a.Ld + li) * Q + q, D);
      hp[tid] = hi;
      hp[D + tid] = mid;
    }
    if (li + 1 < a.Ld) grid_sync(a.barrier, target, G);
    stamp();
  }
}

bool decoder_mega_supported(int d_model, int nhead, int dim_ff, int Q, int Ntok) {
  using namespace dmk;
  return d_model == D && nhead == NH && dim_ff % FU == 0 && Q >= 1 && Q <= 32 && Ntok >= 1 &&
         V_SC + RB * NH * ((Ntok + 3) & ~3) <= OFF_WB;
}

cudaError_t launch_decoder_mega(DecMegaArgs a, cudaStream_t st) {
  using namespace dmk;
  static DeviceOnce once;
  const size_t smem = (size_t)SMEM_FLOATS * sizeof(float);
  cudaError_t e = once.run([&]() -> cudaError_t {
    return cudaFuncSetAttribute(decoder_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  });
  if (e != cudaSuccess) return e;
  if (!decoder_mega_supported(D, NH, a.dim_ff, a.Q, a.Ntok) || a.Ld < 1 || a.Ld > DEC_MEGA_MAX_LAYERS) return cudaErrorInvalidValue;
  a.ntok_pad = (a.Ntok + 3) & ~3;
  e = cudaMemsetAsync(a.barrier, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(device_num_sms());
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;               // all CTAs resident: the grid barriers cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, decoder_mega_kernel, a);
}
