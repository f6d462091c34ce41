// CUDA-core kernels of the TubeR forward path (sm_100a): stem max pool,
// depthwise 3x3x3 stencil, strided gathers, temporal pooling, fp32 GEMM for the token-sized
// linear layers, LayerNorm, attention cores, padding-mask resize and the 3-D sine position code.
// The tensor-core GEMM lives in gemm_tc.cu.  All activations are channels-last (NDHWC).
#include <cuda.h>
#include <float.h>
#include <math.h>

#include "kernels.h"

// MaxPool3d((1,3,3),s(1,2,2),p(0,1,1))  (ir_CSN_152.py:122,179): fp32 -> split
__global__ void maxpool_hw_kernel(const float* __restrict__ in, void* __restrict__ out, int BT, int H1, int W1,
                                  int H2, int W2, int C) {
  const int c4n = C >> 2;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)BT * H2 * W2 * c4n;
  if (idx >= total) return;
  int c4 = (int)(idx % c4n);
  long long vox = idx / c4n;
  int w2 = (int)(vox % W2), h2 = (int)((vox / W2) % H2);
  long long bt = vox / ((long long)W2 * H2);
  float4 m = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
#pragma unroll
  for (int dh = -1; dh <= 1; ++dh) {
    int h = 2 * h2 + dh;
    if (h < 0 || h >= H1) continue;
#pragma unroll
    for (int dw = -1; dw <= 1; ++dw) {
      int w = 2 * w2 + dw;
      if (w < 0 || w >= W1) continue;
      float4 v = __ldg(reinterpret_cast<const float4*>(in + ((bt * H1 + h) * W1 + w) * C) + c4);
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  }
  __nv_bfloat16* hi = split_hi(out, vox, C) + c4 * 4;
  store_split4(hi, hi + C, m);
}

cudaError_t launch_maxpool_hw(const float* in, void* out_split, int BT, int H1, int W1, int H2, int W2, int C,
                              cudaStream_t st) {
  long long total = (long long)BT * H2 * W2 * (C / 4);
  maxpool_hw_kernel<<<ceil_div(total, 256), 256, 0, st>>>(in, out_split, BT, H1, W1, H2, W2, C);
  return cudaGetLastError();
}

// =============================================================================================
// Depthwise 3x3x3 conv (groups = C), stride (st,ss,ss), pad 1, + BN + ReLU  (ir_CSN_152.py:48-56,77-79)
// HBM-bound stencil: a thread owns 4 channels x 4 consecutive output columns, so the 9 input
// rows it touches are read as float4 along the contiguous channel axis and each loaded value
// feeds up to 3 outputs from registers.  fp32 in, split-bf16 out (operand of the next GEMM).
// =============================================================================================
template <int SS>
__global__ void __launch_bounds__(256)
dwconv_kernel(const float* __restrict__ in, const float* __restrict__ wpk, const float* __restrict__ scale,
              const float* __restrict__ shift, void* __restrict__ out, int B, int Ti, int Hi, int Wi, int C,
              int st_t, int To, int Ho, int Wo) {
  constexpr int WS = 4, NIN = (WS - 1) * SS + 3;
  const int c4n = C >> 2;
  const int wstrips = (Wo + WS - 1) / WS;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * To * Ho * wstrips * c4n;
  if (idx >= total) return;
  int c4 = (int)(idx % c4n);
  long long r = idx / c4n;
  int wsi = (int)(r % wstrips); r /= wstrips;
  int ho = (int)(r % Ho); r /= Ho;
  int to = (int)(r % To);
  int b = (int)(r / To);
  const int wo0 = wsi * WS;
  const int iw0 = wo0 * SS - 1;

  float4 acc[WS];
#pragma unroll
  for (int j = 0; j < WS; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);

#pragma unroll
  for (int kt = 0; kt < 3; ++kt) {
    int it = to * st_t - 1 + kt;
    if (it < 0 || it >= Ti) continue;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      int ih = ho * SS - 1 + kh;
      if (ih < 0 || ih >= Hi) continue;
      const float* irow = in + (((long long)b * Ti + it) * Hi + ih) * (long long)Wi * C;
      float4 x[NIN];
#pragma unroll
      for (int q = 0; q < NIN; ++q) {
        int iw = iw0 + q;
        x[q] = (iw >= 0 && iw < Wi) ? __ldg(reinterpret_cast<const float4*>(irow + (long long)iw * C) + c4)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        float4 w = __ldg(reinterpret_cast<const float4*>(wpk + ((kt * 3 + kh) * 3 + kw) * C) + c4);
#pragma unroll
        for (int j = 0; j < WS; ++j) {
          float4 v = x[j * SS + kw];
          acc[j].x = fmaf(v.x, w.x, acc[j].x);
          acc[j].y = fmaf(v.y, w.y, acc[j].y);
          acc[j].z = fmaf(v.z, w.z, acc[j].z);
          acc[j].w = fmaf(v.w, w.w, acc[j].w);
        }
      }
    }
  }
  float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + c4);
  float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + c4);
  long long orow0 = (((long long)b * To + to) * Ho + ho) * Wo + wo0;
#pragma unroll
  for (int j = 0; j < WS; ++j) {
    if (wo0 + j >= Wo) break;
    float4 o;
    o.x = fmaxf(fmaf(acc[j].x, sc.x, sh.x), 0.f);
    o.y = fmaxf(fmaf(acc[j].y, sc.y, sh.y), 0.f);
    o.z = fmaxf(fmaf(acc[j].z, sc.z, sh.z), 0.f);
    o.w = fmaxf(fmaf(acc[j].w, sc.w, sh.w), 0.f);
    __nv_bfloat16* hi = split_hi(out, orow0 + j, C) + c4 * 4;
    store_split4(hi, hi + C, o);
  }
}

// Stride-1 variant with a shared-memory halo tile (the 24 stride-1 blocks of CSN-152 / 13 of CSN-50 are the
// ones that matter): a tile = TT x TH x TW outputs for CC = 32 channels, computed from a (TT+2) x (TH+2) x (TW+2)
// input block that ONE 5-D TMA load brings in (out-of-volume voxels are zero-filled by the TMA unit = the conv's
// padding), so every input element leaves L2 ~2.3x instead of 9x and no thread computes a load address.
// CTAs are persistent (one per SM) and double buffered: the halo block of tile i+1 streams in while tile i is
// computed.  A thread owns one output row of TW = 8 voxels x 4 channels.
namespace dwt {
constexpr int TT = 4, TH = 8, TW = 8, CC = 32;
constexpr int IT = TT + 2, IH = TH + 2, IW = TW + 2;
constexpr int IN_F4 = IT * IH * IW * (CC / 4);            // float4 slots of the input tile
constexpr int W_F4 = 27 * (CC / 4);                       // this chunk's 27 x CC weights
constexpr int STAGE_F4 = IN_F4 + W_F4 + 8;                // keep stages 128-byte aligned
constexpr int STAGE_TX = (IN_F4 + W_F4) * 16;
constexpr int NSTAGE = 2;                                 // double buffered, one CTA per SM (two single-stage CTAs measured slower)
constexpr int SMEM_BYTES = NSTAGE * STAGE_F4 * 16 + 16;   // + mbarriers
constexpr int THREADS = TT * TH * (CC / 4);               // 256
}  // namespace dwt

__global__ void __launch_bounds__(dwt::THREADS, 1)
dwconv_s1_tiled_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmW,
                       const float* __restrict__ scale, const float* __restrict__ shift, void* __restrict__ out, int B, int T,
                       int H, int W, int C, int ntiles) {
  using namespace dwt;
  extern __shared__ __align__(128) float4 dw_smem[];
  const int tid = threadIdx.x;
  const int wt = (W + TW - 1) / TW, ht = (H + TH - 1) / TH, tt = (T + TT - 1) / TT, nch = C / CC;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(dw_smem);
  const uint32_t bar0 = sbase + NSTAGE * STAGE_F4 * 16;

  auto decode = [&](int tile, int& b, int& t0, int& h0, int& w0, int& cb) {
    cb = (tile % nch) * CC; tile /= nch;                    // channel chunks of one spatial tile are neighbours (L2 reuse)
    w0 = (tile % wt) * TW; tile /= wt;
    h0 = (tile % ht) * TH; tile /= ht;
    t0 = (tile % tt) * TT;
    b = tile / tt;
  };
  auto prefetch = [&](int tile, int buf) {                  // one thread: two bulk tensor loads onto the stage's mbarrier
    int b, t0, h0, w0, cb;
    decode(tile, b, t0, h0, w0, cb);
    const uint32_t dst = sbase + buf * STAGE_F4 * 16, bar = bar0 + 8 * buf;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)STAGE_TX) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(&tmIn), "r"(bar), "r"(cb), "r"(w0 - 1), "r"(h0 - 1), "r"(t0 - 1), "r"(b) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst + IN_F4 * 16), "l"(&tmW), "r"(bar), "r"(cb), "r"(0) : "memory");
  };

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int c4 = tid % (CC / 4);
  const int lh = (tid / (CC / 4)) % TH, lt = tid / ((CC / 4) * TH);
  int tile = blockIdx.x, buf = 0;
  uint32_t phase[2] = {0, 0};
  if (tid == 0 && tile < ntiles) prefetch(tile, 0);
  for (; tile < ntiles; tile += gridDim.x, buf = (buf + 1) % NSTAGE) {
    const int next = tile + gridDim.x;
    if (NSTAGE == 2 && tid == 0 && next < ntiles) prefetch(next, buf ^ 1);   // that stage was released by the barrier ending the last iteration
    {
      const uint32_t bar = bar0 + 8 * buf, par = phase[buf];
      asm volatile(
          "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n"
          ::"r"(bar), "r"(par) : "memory");
      phase[buf] ^= 1;
    }
    int b, t0, h0, w0, cb;
    decode(tile, b, t0, h0, w0, cb);
    const float4* tsm = dw_smem + buf * STAGE_F4;
    const float4* wsm = tsm + IN_F4;
    const int oh = h0 + lh, ot = t0 + lt;
    if (oh < H && ot < T) {
      float4 acc[TW];
#pragma unroll
      for (int j = 0; j < TW; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const float4* row = tsm + (((lt + kt) * IH + lh + kh) * IW) * (CC / 4) + c4;
          float4 x[IW];
#pragma unroll
          for (int q = 0; q < IW; ++q) x[q] = row[q * (CC / 4)];
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float4 wv = wsm[((kt * 3 + kh) * 3 + kw) * (CC / 4) + c4];
#pragma unroll
            for (int j = 0; j < TW; ++j) {
              acc[j].x = fmaf(x[j + kw].x, wv.x, acc[j].x);
              acc[j].y = fmaf(x[j + kw].y, wv.y, acc[j].y);
              acc[j].z = fmaf(x[j + kw].z, wv.z, acc[j].z);
              acc[j].w = fmaf(x[j + kw].w, wv.w, acc[j].w);
            }
          }
        }
      }
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + cb) + c4);
      const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + cb) + c4);
      const long long orow0 = (((long long)b * T + ot) * H + oh) * W + w0;
#pragma unroll
      for (int j = 0; j < TW; ++j) {
        if (w0 + j >= W) break;
        float4 o;
        o.x = fmaxf(fmaf(acc[j].x, sc.x, sh.x), 0.f);
        o.y = fmaxf(fmaf(acc[j].y, sc.y, sh.y), 0.f);
        o.z = fmaxf(fmaf(acc[j].z, sc.z, sh.z), 0.f);
        o.w = fmaxf(fmaf(acc[j].w, sc.w, sh.w), 0.f);
        __nv_bfloat16* hi = split_hi(out, orow0 + j, C) + cb + c4 * 4;
        store_split4(hi, hi + C, o);
      }
    }
    __syncthreads();                                        // every reader is done: the stage may be refilled
    if (NSTAGE == 1 && tid == 0 && next < ntiles) prefetch(next, 0);
  }
}

// Stride-1 variant #2, "rolling t": a work item = (clip, block of TC output frames, 16x16 spatial tile, 32 channels).
// The CTA walks along t: every step ONE input frame of the halo tile (18 x 18 x 32 ch, 41 KB, one 5-D TMA load, zero
// filled outside the volume) arrives in a 4-slot ring, and a thread (one output row of 8 voxels x 4 channels) adds that
// frame's 3 x 3 taps to THREE accumulator sets held in registers -- output frames it-1 (kt = 2), it (kt = 1), it+1 (kt = 0) --
// then emits frame it-1.  Every input value is read from shared memory 3 times instead of 9 (per voxel x 4 channels: 3.75
// LDS.128 for the input + 3.4 for the broadcast weights, against 108 FMAs), which moves the kernel from shared-memory bound
// to HBM / FMA bound; items, steps and ring slots run as one flat software pipeline (prefetch distance 3) across items.
namespace dwr {
constexpr int TH = 16, TW = 16, CC = 32, IH = TH + 2, IW = TW + 2, NSLOT = 4, TCMAX = 16;
constexpr int PLANE_F4 = IH * IW * (CC / 4);              // float4 slots of one input frame of the halo tile
constexpr int PLANE_BYTES = PLANE_F4 * 16;                // 41472
constexpr int W_F4 = 27 * (CC / 4);
constexpr int OFF_W = NSLOT * PLANE_BYTES;                // two weight buffers (item parity)
constexpr int OFF_OUT = OFF_W + 2 * W_F4 * 16;            // one output frame of the tile in split format: [h][w][hi 64 B | mid 64 B]
constexpr int OUT_BYTES = TH * TW * CC * 4;               // 32 KB
constexpr int OFF_BAR = OFF_OUT + OUT_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + NSLOT * 8;
static_assert(OFF_OUT % 128 == 0 && SMEM_BYTES <= 232448, "shared-memory layout");
constexpr int THREADS = TH * (TW / 8) * (CC / 4);         // 256 (CPT = 4); the CPT = 2 variant runs 512
}  // namespace dwr

// CPT = channels per thread.  4: 8 warps, a thread owns 8 voxels x 4 channels (two packed pairs per voxel).  2: 16 warps, 8 voxels x
// one packed pair -- the same shared-memory wavefronts and FFMA2 count per CTA, but four warps per scheduler instead of two to hide
// the LDS -> FFMA2 latencies (the kernel is issue / latency bound, not FLOP or HBM bound: profiles/r1_ncu_dwroll_v9.txt).
template <int CPT>
__global__ void __launch_bounds__(dwr::TH * (dwr::TW / 8) * (dwr::CC / CPT), 1)
dwconv_s1_roll_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut, const float* __restrict__ wpk,
                      const float* __restrict__ scale, const float* __restrict__ shift, int B, int T, int H, int W, int C, int TC, int nitems) {
  using namespace dwr;
  constexpr int NP = CPT / 2;                               // packed fp32 pairs per voxel and thread
  constexpr int CG = CC / CPT;                              // threads along the channel slice
  extern __shared__ __align__(128) float4 dwr_smem[];
  const int tid = threadIdx.x;
  const int wt = (W + TW - 1) / TW, ht = (H + TH - 1) / TH, tcn = (T + TC - 1) / TC, nch = C / CC;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(dwr_smem);
  const uint32_t bar0 = sbase + OFF_BAR;
  const int S = TC + 2;                                     // steps (input frames) per item

  auto decode = [&](int item, int& b, int& t0, int& h0, int& w0, int& cb) {
    cb = (item % nch) * CC; item /= nch;
    w0 = (item % wt) * TW; item /= wt;
    h0 = (item % ht) * TH; item /= ht;
    t0 = (item % tcn) * TC;
    b = item / tcn;
  };
  // flat sequence q = (k-th item of this CTA, step s): frame t0 - 1 + s of that item's halo tile -> ring slot q % NSLOT
  const int my_items = blockIdx.x < nitems ? (nitems - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const int total = my_items * S;
  auto issue = [&](int q) {                                 // one thread
    const int k = q / S, s = q - k * S;
    int b, t0, h0, w0, cb;
    decode(blockIdx.x + k * gridDim.x, b, t0, h0, w0, cb);
    const uint32_t dst = sbase + (q % NSLOT) * PLANE_BYTES, bar = bar0 + 8 * (q % NSLOT);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)PLANE_BYTES) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(&tmIn), "r"(bar), "r"(cb), "r"(w0 - 1), "r"(h0 - 1), "r"(t0 - 1 + s), "r"(b) : "memory");
  };
  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_trigger();
  pdl_wait();
  if (tid == 0)
    for (int q = 0; q < NSLOT - 1 && q < total; ++q) issue(q);

  const int cg = tid % CG, wh = (tid / CG) & 1, lh = tid / (2 * CG);
  // accumulators as packed fp32 pairs (FFMA2: one issue slot per two FMAs -- the kernel is issue bound, not FLOP bound)
  typedef unsigned long long u64;
  u64 accA[8][NP], accB[8][NP], accC[8][NP];
  auto zero = [](u64 (&a)[8][NP]) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < NP; ++e) a[j][e] = 0ull;
  };
  // acc += the 3 taps (kw) of filter row (kt, kh) applied to the input row x; a thread's CPT channels are NP consecutive u64
  auto add_row = [&](u64 (&acc)[8][NP], const u64 (&x)[10][NP], const u64* wsm, int kt, int kh) {
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      u64 wv[NP];
#pragma unroll
      for (int e = 0; e < NP; ++e) wv[e] = wsm[(((kt * 3 + kh) * 3 + kw) * CG + cg) * NP + e];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < NP; ++e) asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j][e]) : "l"(x[j + kw][e]), "l"(wv[e]));
    }
  };

  int q = 0;
  for (int k = 0; k < my_items; ++k) {
    int b, t0, h0, w0, cb;
    decode(blockIdx.x + k * gridDim.x, b, t0, h0, w0, cb);
    float4* wsm4 = reinterpret_cast<float4*>(reinterpret_cast<char*>(dwr_smem) + OFF_W) + (k & 1) * W_F4;
    if (tid < W_F4) wsm4[tid] = __ldg(reinterpret_cast<const float4*>(wpk + (tid >> 3) * C + cb) + (tid & 7));
    const u64* wsm = reinterpret_cast<const u64*>(wsm4);
    float sc[CPT], sh[CPT];
#pragma unroll
    for (int e = 0; e < CPT; ++e) { sc[e] = __ldg(scale + cb + cg * CPT + e); sh[e] = __ldg(shift + cb + cg * CPT + e); }
    __syncthreads();
    const int oh = h0 + lh, ow0 = w0 + wh * 8;
    const int tend = min(t0 + TC, T);                       // output frames of this item: [t0, tend)
    zero(accA); zero(accB); zero(accC);
    // roles rotate every step: frame it updates P (output it-1, kt=2), Cc (output it, kt=1), N (output it+1, kt=0)
    auto step = [&](int s, u64 (&P)[8][NP], u64 (&Cc)[8][NP], u64 (&N)[8][NP]) {
      if (tid == 0 && q + NSLOT - 1 < total) issue(q + NSLOT - 1);   // its slot was released by the barrier ending step q-1
      {
        const uint32_t bar = bar0 + 8 * (q % NSLOT), par = (uint32_t)((q / NSLOT) & 1);
        asm volatile(
            "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n"
            ::"r"(bar), "r"(par) : "memory");
      }
      const int it = t0 - 1 + s;
      const u64* tsm = reinterpret_cast<const u64*>(dwr_smem) + (size_t)(q % NSLOT) * PLANE_F4 * 2;
      if (it >= 0 && it < T) {                               // (frames outside the clip are all zero: nothing to add)
        const bool do_p = it - 1 >= t0, do_c = it >= t0 && it < tend, do_n = it + 1 < tend;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          u64 x[10][NP];
#pragma unroll
          for (int j = 0; j < 10; ++j) {
            const u64* xp = tsm + (((lh + kh) * IW + wh * 8 + j) * CG + cg) * NP;
            if constexpr (NP == 2) {
              const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(xp);
              x[j][0] = v.x; x[j][1] = v.y;
            } else {
              x[j][0] = xp[0];
            }
          }
          if (do_p) add_row(P, x, wsm, 2, kh);
          if (do_c) add_row(Cc, x, wsm, 1, kh);
          if (do_n) add_row(N, x, wsm, 0, kh);
        }
      }
      // emit output frame it-1 through a shared-memory staging tile and ONE asynchronous TMA store (per-thread 4 / 8-byte global
      // stores kept the warps in the LSU: the store phase did not overlap the next step's FMAs).  The tensor map clips the tile
      // at the volume's edges.
      const int ot = it - 1;
      const bool emit = ot >= t0 && ot < tend;               // uniform over the CTA
      if (emit && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous store has read the staging tile
      __syncthreads();                                       // every reader is done with this input slot; the staging tile is free
      if (emit) {
        char* stg = reinterpret_cast<char*>(dwr_smem) + OFF_OUT + (lh * TW + wh * 8) * (CC * 4) + cg * CPT * 2;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float o[CPT];
#pragma unroll
          for (int e = 0; e < NP; ++e) {
            o[2 * e] = fmaxf(fmaf(__uint_as_float((uint32_t)P[j][e]), sc[2 * e], sh[2 * e]), 0.f);
            o[2 * e + 1] = fmaxf(fmaf(__uint_as_float((uint32_t)(P[j][e] >> 32)), sc[2 * e + 1], sh[2 * e + 1]), 0.f);
          }
          char* vp = stg + j * (CC * 4);
          if constexpr (CPT == 4) {
            uint2 h2, m2;
            split_bf16x2(o[0], o[1], h2.x, m2.x);
            split_bf16x2(o[2], o[3], h2.y, m2.y);
            *reinterpret_cast<uint2*>(vp) = h2;
            *reinterpret_cast<uint2*>(vp + CC * 2) = m2;
          } else {
            uint32_t h2, m2;
            split_bf16x2(o[0], o[1], h2, m2);
            *reinterpret_cast<uint32_t*>(vp) = h2;
            *reinterpret_cast<uint32_t*>(vp + CC * 2) = m2;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      zero(P);                                               // becomes the N set of the next step
      if (emit) {
        __syncthreads();                                     // staging tile complete
        if (tid == 0) {
          asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                       ::"l"(&tmOut), "r"(sbase + OFF_OUT), "r"(cb), "r"(0), "r"(w0), "r"(h0), "r"(b * T + ot) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      ++q;
    };
    int s = 0;
    for (; s + 2 < S; s += 3) {
      step(s, accA, accB, accC);
      step(s + 1, accB, accC, accA);
      step(s + 2, accC, accA, accB);
    }
    if (s < S) { step(s, accA, accB, accC); ++s; }
    if (s < S) { step(s, accB, accC, accA); ++s; }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory must outlive the last store
}

// Stride-(2,2,2) variant of the rolling-t kernel (first block of layer2 / layer3, ir_CSN_152.py:48-51 with stride 2): a work
// item = (clip, TC output frames, 8 x 16 output tile, 32 channels); input frame it = 2*ot - 1 + kt feeds output frame ot, so an
// odd input frame 2j-1 finishes output j-1 (kt = 2, emitted) and starts output j (kt = 0), an even frame 2j adds kt = 1 to
// output j: two accumulator sets alternate.  A thread owns 4 output columns x 4 channels and reads the 3 x 9 input window of a
// frame from the halo tile (17 x 33 x 32 ch, one 5-D TMA load per frame, 3-slot ring).
namespace dws {
constexpr int TH = 8, TW = 16, CC = 32, IH = 2 * TH + 1, IW = 2 * TW + 1, NSLOT = 3, TCMAX = 64;
constexpr int PLANE_F4 = IH * IW * (CC / 4);
constexpr int PLANE_BYTES = PLANE_F4 * 16;                // 71808
constexpr int W_F4 = 27 * (CC / 4);
constexpr int OFF_W = NSLOT * PLANE_BYTES;
constexpr int OFF_BAR = OFF_W + 2 * W_F4 * 16;
constexpr int SMEM_BYTES = OFF_BAR + NSLOT * 8;
constexpr int THREADS = TH * (TW / 4) * (CC / 4);         // 256
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
}  // namespace dws

__global__ void __launch_bounds__(dws::THREADS, 1)
dwconv_s2_roll_kernel(const __grid_constant__ CUtensorMap tmIn, const float* __restrict__ wpk, const float* __restrict__ scale,
                      const float* __restrict__ shift, void* __restrict__ out, int B, int Ti, int To, int Ho, int Wo, int C, int TC,
                      int nitems) {
  using namespace dws;
  extern __shared__ __align__(128) float4 dws_smem[];
  const int tid = threadIdx.x;
  const int wt = (Wo + TW - 1) / TW, ht = (Ho + TH - 1) / TH, tcn = (To + TC - 1) / TC, nch = C / CC;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(dws_smem);
  const uint32_t bar0 = sbase + OFF_BAR;
  const int S = 2 * TC + 1;                                 // input frames per item: 2*t0 - 1 .. 2*(t0 + TC - 1) + 1

  auto decode = [&](int item, int& b, int& t0, int& h0, int& w0, int& cb) {
    cb = (item % nch) * CC; item /= nch;
    w0 = (item % wt) * TW; item /= wt;
    h0 = (item % ht) * TH; item /= ht;
    t0 = (item % tcn) * TC;
    b = item / tcn;
  };
  const int my_items = (int)blockIdx.x < nitems ? (nitems - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total = my_items * S;
  auto issue = [&](int q) {
    const int k = q / S, s = q - k * S;
    int b, t0, h0, w0, cb;
    decode(blockIdx.x + k * gridDim.x, b, t0, h0, w0, cb);
    const uint32_t dst = sbase + (q % NSLOT) * PLANE_BYTES, bar = bar0 + 8 * (q % NSLOT);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)PLANE_BYTES) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(&tmIn), "r"(bar), "r"(cb), "r"(2 * w0 - 1), "r"(2 * h0 - 1), "r"(2 * t0 - 1 + s), "r"(b) : "memory");
  };
  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0)
    for (int q = 0; q < NSLOT - 1 && q < total; ++q) issue(q);

  const int c4 = tid & 7, wq = (tid >> 3) & 3, lh = tid >> 5;
  typedef unsigned long long u64;
  u64 accX[4][2], accY[4][2];
  auto zero = [](u64 (&a)[4][2]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j][0] = a[j][1] = 0ull;
  };
  int q = 0;
  for (int k = 0; k < my_items; ++k) {
    int b, t0, h0, w0, cb;
    decode(blockIdx.x + k * gridDim.x, b, t0, h0, w0, cb);
    float4* wsm4 = reinterpret_cast<float4*>(reinterpret_cast<char*>(dws_smem) + OFF_W) + (k & 1) * W_F4;
    if (tid < W_F4) wsm4[tid] = __ldg(reinterpret_cast<const float4*>(wpk + (tid >> 3) * C + cb) + (tid & 7));
    const ulonglong2* wsm = reinterpret_cast<const ulonglong2*>(wsm4);
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + cb) + c4);
    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + cb) + c4);
    __syncthreads();
    const int oh = h0 + lh, ow0 = w0 + wq * 4;
    const int tend = min(t0 + TC, To);
    zero(accX); zero(accY);
    // one input frame: acc_fin (output frame jf, taps kt_fin) is finished and emitted, acc_add (taps kt_add) is added to
    auto frame = [&](int it, u64 (*fin)[2], int jf, u64 (*add)[2], int kt_add, bool do_add) {
      if (tid == 0 && q + NSLOT - 1 < total) issue(q + NSLOT - 1);
      {
        const uint32_t bar = bar0 + 8 * (q % NSLOT), par = (uint32_t)((q / NSLOT) & 1);
        asm volatile(
            "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n"
            ::"r"(bar), "r"(par) : "memory");
      }
      const ulonglong2* tsm = reinterpret_cast<const ulonglong2*>(dws_smem) + (q % NSLOT) * PLANE_F4;
      const bool do_fin = fin != nullptr && jf >= t0 && jf < tend;
      if (it >= 0 && it < Ti && (do_fin || do_add)) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          ulonglong2 x[9];
#pragma unroll
          for (int j = 0; j < 9; ++j) x[j] = tsm[((2 * lh + kh) * IW + wq * 8 + j) * (CC / 4) + c4];
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            if (do_fin) {
              const ulonglong2 wv = wsm[((2 * 3 + kh) * 3 + kw) * (CC / 4) + c4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(fin[j][0]) : "l"(x[2 * j + kw].x), "l"(wv.x));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(fin[j][1]) : "l"(x[2 * j + kw].y), "l"(wv.y));
              }
            }
            if (do_add) {
              const ulonglong2 wv = wsm[((kt_add * 3 + kh) * 3 + kw) * (CC / 4) + c4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(add[j][0]) : "l"(x[2 * j + kw].x), "l"(wv.x));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(add[j][1]) : "l"(x[2 * j + kw].y), "l"(wv.y));
              }
            }
          }
        }
      }
      if (do_fin) {
        if (oh < Ho) {
          const long long orow0 = (((long long)b * To + jf) * Ho + oh) * Wo + ow0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (ow0 + j < Wo) {
              float4 o;
              o.x = fmaxf(fmaf(__uint_as_float((uint32_t)fin[j][0]), sc.x, sh.x), 0.f);
              o.y = fmaxf(fmaf(__uint_as_float((uint32_t)(fin[j][0] >> 32)), sc.y, sh.y), 0.f);
              o.z = fmaxf(fmaf(__uint_as_float((uint32_t)fin[j][1]), sc.z, sh.z), 0.f);
              o.w = fmaxf(fmaf(__uint_as_float((uint32_t)(fin[j][1] >> 32)), sc.w, sh.w), 0.f);
              __nv_bfloat16* hi = split_hi(out, orow0 + j, C) + cb + c4 * 4;
              store_split4(hi, hi + C, o);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) fin[j][0] = fin[j][1] = 0ull;
      }
      __syncthreads();
      ++q;
    };
    // output frame j lives in accX for even (j - t0), accY for odd
    for (int jj = 0; jj <= TC; jj += 2) {
      const int j = t0 + jj;
      frame(2 * j - 1, accY, j - 1, accX, 0, j < tend);                   // odd input frame: finish j-1 (kt = 2), start j (kt = 0)
      if (jj < TC) frame(2 * j, nullptr, 0, accX, 1, j < tend);           // even input frame: kt = 1 of j
      if (jj + 1 <= TC) {
        frame(2 * j + 1, accX, j, accY, 0, j + 1 < tend);
        if (jj + 1 < TC) frame(2 * j + 2, nullptr, 0, accY, 1, j + 1 < tend);
      }
    }
  }
}

namespace dwt {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
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
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dwconv_s1_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dwconv_s1_roll_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, dwr::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dwconv_s1_roll_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dwr::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dwconv_s2_roll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dws::SMEM_BYTES);
    return e;
  });
}
}  // namespace dwt

cudaError_t launch_dwconv(const float* in, const float* wpk, const float* scale, const float* shift,
                          void* out_split, int B, int Ti, int Hi, int Wi, int C, int st_t, int st_s, int To,
                          int Ho, int Wo, cudaStream_t st) {
  if (st_t == 1 && st_s == 1 && C % dwt::CC == 0) {
    using namespace dwt;
    cudaError_t e = init_once();
    if (e != cudaSuccess) return e;
    CUtensorMap tmIn, tmW;
    {
      cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)Ti, (cuuint64_t)B};
      cuuint64_t strides[4] = {(cuuint64_t)C * 4, (cuuint64_t)Wi * C * 4, (cuuint64_t)Hi * Wi * C * 4, (cuuint64_t)Ti * Hi * Wi * C * 4};
      cuuint32_t box[5] = {CC, IW, IH, IT, 1}, es[5] = {1, 1, 1, 1, 1};
      if (g_encode(&tmIn, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    }
    {
      cuuint64_t dims[2] = {(cuuint64_t)C, 27};
      cuuint64_t strides[1] = {(cuuint64_t)C * 4};
      cuuint32_t box[2] = {CC, 27}, es[2] = {1, 1};
      if (g_encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(wpk), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    }
    static const bool use_tiled = [] { const char* e = getenv("TUBER_DW_TILED"); return e && e[0] == '1'; }();
    if (!use_tiled) {
      CUtensorMap tmR;
      cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)Ti, (cuuint64_t)B};
      cuuint64_t strides[4] = {(cuuint64_t)C * 4, (cuuint64_t)Wi * C * 4, (cuuint64_t)Hi * Wi * C * 4, (cuuint64_t)Ti * Hi * Wi * C * 4};
      cuuint32_t box[5] = {dwr::CC, dwr::IW, dwr::IH, 1, 1}, es[5] = {1, 1, 1, 1, 1};
      if (g_encode(&tmR, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
      // frames per item: as many as possible (fewer halo frames) while the grid still fills the machine about twice
      const long long cols = (long long)B * ceil_div(Hi, dwr::TH) * ceil_div(Wi, dwr::TW) * (C / dwr::CC);
      // frames per item: minimise (rounds over the SMs) x (steps per item = TC + 2 halo frames); ties go to the longer item
      int TC = 1;
      long long best = -1;
      for (int k = 1; k <= Ti; ++k) {
        const int tc = ceil_div(Ti, k);
        if (tc > dwr::TCMAX) continue;
        const long long cost = (long long)ceil_div(cols * ceil_div(Ti, tc), device_num_sms()) * (tc + 2);
        if (best < 0 || cost < best) { best = cost; TC = tc; }
      }
      const long long items = cols * ceil_div(Ti, TC);
      const int grid = (int)(items < device_num_sms() ? items : device_num_sms());
      CUtensorMap tmO;                                       // split output [B*T, H, W, 2 planes, C] bf16, box = one 16 x 16 x 32-channel frame tile
      {
        cuuint64_t od[5] = {(cuuint64_t)C, 2, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)B * Ti};
        cuuint64_t os[4] = {(cuuint64_t)C * 2, (cuuint64_t)C * 4, (cuuint64_t)Wi * C * 4, (cuuint64_t)Hi * Wi * C * 4};
        cuuint32_t ob[5] = {dwr::CC, 2, dwr::TW, dwr::TH, 1}, oe[5] = {1, 1, 1, 1, 1};
        if (g_encode(&tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, out_split, od, os, ob, oe, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
          return cudaErrorInvalidValue;
      }
      // 16 warps (a packed channel pair per thread) unless TUBER_DW_WARPS8=1 (the 8-warp variant: the tests' cross-check)
      const char* w8e = getenv("TUBER_DW_WARPS8");            // read per call: the tests switch it between two plans of one process
      const bool warps8 = w8e && w8e[0] == '1';
      if (warps8)
        return launch_pdl(dwconv_s1_roll_kernel<4>, dim3(grid), dim3(dwr::THREADS), dwr::SMEM_BYTES, st, tmR, tmO, wpk, scale, shift, B, Ti, Hi, Wi, C, TC, (int)items);
      return launch_pdl(dwconv_s1_roll_kernel<2>, dim3(grid), dim3(2 * dwr::THREADS), dwr::SMEM_BYTES, st, tmR, tmO, wpk, scale, shift, B, Ti, Hi, Wi, C, TC, (int)items);
    }
    const long long tiles = (long long)B * ceil_div(Ti, TT) * ceil_div(Hi, TH) * ceil_div(Wi, TW) * (C / CC);
    const long long slots = (long long)device_num_sms() * (NSTAGE == 1 ? 2 : 1);
    const int grid = (int)(tiles < slots ? tiles : slots);
    dwconv_s1_tiled_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(tmIn, tmW, scale, shift, out_split, B, Ti, Hi, Wi, C, (int)tiles);
    return cudaGetLastError();
  }
  if (st_t == 2 && st_s == 2 && C % dws::CC == 0) {
    using namespace dws;
    cudaError_t e = dwt::init_once();
    if (e != cudaSuccess) return e;
    CUtensorMap tmS;
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)Ti, (cuuint64_t)B};
    cuuint64_t strides[4] = {(cuuint64_t)C * 4, (cuuint64_t)Wi * C * 4, (cuuint64_t)Hi * Wi * C * 4, (cuuint64_t)Ti * Hi * Wi * C * 4};
    cuuint32_t box[5] = {CC, IW, IH, 1, 1}, es[5] = {1, 1, 1, 1, 1};
    if (dwt::g_encode(&tmS, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
    const long long cols = (long long)B * ceil_div(Ho, TH) * ceil_div(Wo, TW) * (C / CC);
    int TC = 1;
    long long best = -1;
    for (int k = 1; k <= To; ++k) {                          // minimise rounds x (2 TC + 1) input frames per item
      const int tc = ceil_div(To, k);
      if (tc > TCMAX) continue;
      const long long cost = (long long)ceil_div(cols * ceil_div(To, tc), device_num_sms()) * (2 * tc + 1);
      if (best < 0 || cost < best) { best = cost; TC = tc; }
    }
    const long long items = cols * ceil_div(To, TC);
    const int grid = (int)(items < device_num_sms() ? items : device_num_sms());
    dwconv_s2_roll_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(tmS, wpk, scale, shift, out_split, B, Ti, To, Ho, Wo, C, TC, (int)items);
    return cudaGetLastError();
  }
  long long total = (long long)B * To * Ho * ((Wo + 3) / 4) * (C / 4);
  int grid = ceil_div(total, 256);
  if (st_s == 1)
    dwconv_kernel<1><<<grid, 256, 0, st>>>(in, wpk, scale, shift, out_split, B, Ti, Hi, Wi, C, st_t, To, Ho, Wo);
  else if (st_s == 2)
    dwconv_kernel<2><<<grid, 256, 0, st>>>(in, wpk, scale, shift, out_split, B, Ti, Hi, Wi, C, st_t, To, Ho, Wo);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// =============================================================================================
// Row gathers (raw 16-byte chunks; format-agnostic because a split row has fp32's byte count)
// =============================================================================================
__global__ void gather_rows_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int chunks, int B,
                                   int Ti, int Hi, int Wi, int st_t, int st_s, int To, int Ho, int Wo) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * To * Ho * Wo * chunks;
  if (idx >= total) return;
  int ch = (int)(idx % chunks);
  long long r = idx / chunks;
  int wo = (int)(r % Wo); r /= Wo;
  int ho = (int)(r % Ho); r /= Ho;
  int to = (int)(r % To);
  long long b = r / To;
  long long src = ((b * Ti + (long long)to * st_t) * Hi + (long long)ho * st_s) * Wi + (long long)wo * st_s;
  out[idx] = __ldg(in + src * chunks + ch);
}

cudaError_t launch_gather_rows(const void* in, void* out, int row_bytes, int B, int Ti, int Hi, int Wi, int st_t,
                               int st_s, int To, int Ho, int Wo, cudaStream_t st) {
  int chunks = row_bytes / 16;
  long long total = (long long)B * To * Ho * Wo * chunks;
  gather_rows_kernel<<<ceil_div(total, 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(in),
                                                          reinterpret_cast<uint4*>(out), chunks, B, Ti, Hi, Wi,
                                                          st_t, st_s, To, Ho, Wo);
  return cudaGetLastError();
}

cudaError_t launch_slice_frames(const void* in, void* out, int row_bytes, int B, int Tin, int HW, int t0, int Tn,
                                cudaStream_t st) {
  // rows of frame range [t0, t0+Tn) are contiguous per clip: B strided 2-D copies
  return cudaMemcpy2DAsync(out, (size_t)Tn * HW * row_bytes,
                           reinterpret_cast<const char*>(in) + (size_t)t0 * HW * row_bytes,
                           (size_t)Tin * HW * row_bytes, (size_t)Tn * HW * row_bytes, B,
                           cudaMemcpyDeviceToDevice, st);
}

// Temporal avg / max pool with window = stride = k  (backbone_builder.py:43-46,73)
__global__ void tpool_kernel(const void* __restrict__ in, void* __restrict__ out, int B, int Tin, int HW, int C,
                             int k, int Tout, int is_max) {
  const int c4n = C >> 2;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * Tout * HW * c4n;
  if (idx >= total) return;
  int c4 = (int)(idx % c4n);
  long long r = idx / c4n;
  int p = (int)(r % HW); r /= HW;
  int to = (int)(r % Tout);
  long long b = r / Tout;
  float4 a = is_max ? make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int i = 0; i < k; ++i) {
    long long row = (b * Tin + (long long)to * k + i) * HW + p;
    const __nv_bfloat16* hi = split_hi(in, row, C) + c4 * 4;
    float4 v = load_split4(hi, hi + C);
    if (is_max) {
      a.x = fmaxf(a.x, v.x); a.y = fmaxf(a.y, v.y); a.z = fmaxf(a.z, v.z); a.w = fmaxf(a.w, v.w);
    } else {
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
  }
  if (!is_max) {
    float kk = (float)k;
    a.x /= kk; a.y /= kk; a.z /= kk; a.w /= kk;
  }
  __nv_bfloat16* ho = split_hi(out, (b * Tout + to) * HW + p, C) + c4 * 4;
  store_split4(ho, ho + C, a);
}

cudaError_t launch_tpool(const void* in_split, void* out_split, int B, int Tin, int HW, int C, int k, int Tout,
                         int is_max, cudaStream_t st) {
  long long total = (long long)B * Tout * HW * (C / 4);
  tpool_kernel<<<ceil_div(total, 256), 256, 0, st>>>(in_split, out_split, B, Tin, HW, C, k, Tout, is_max);
  return cudaGetLastError();
}

// =============================================================================================
// Decode pool (backbone_builder.py:75-78; transformer_layers.py:306-366): the cross-attention of the learned pooled
// query over the Tf frame tokens of one pixel, with both projections moved out of the token loop.  The query is input
// independent, so   q_h . (Wk_h x_t + bk_h) = u_h . x_t + const   with u_h = scale * Wk_h^T q_h (folded at finalize; the
// constant cancels in the softmax over t), and   sum_t p_t (Wv_h x_t + bv_h) = Wv_h (sum_t p_t x_t) + bv_h.
// This kernel computes, per pixel and head, the Tf scores, their softmax and the mixed token y_h = sum_t p[h][t] x_t
// (d = 2048); the value projection then runs on B*HW*8 mixed rows instead of B*Tf*HW tokens x (K and V) -- 4.7x fewer
// FLOPs than projecting every token to K and V.  One CTA per pixel, thread = 8 consecutive channels.
//   xt: split [B*Tf*HW, 2048]; U: fp32 [8][2048]; y: split [8 * Mp, 2048], row = h * Mp + pixel (Mp >= B*HW: group stride)
// =============================================================================================
template <int MAXT>
__global__ void __launch_bounds__(256)
pool_mix_kernel(const void* __restrict__ xt, const float* __restrict__ U, void* __restrict__ y, int B, int Tf, int HW, long long Mp) {
  constexpr int D = 2048, NH = 8;
  __shared__ float red[8][NH * MAXT];
  __shared__ float prob[NH * MAXT];
  pdl_trigger();
  pdl_wait();
  const int pix = blockIdx.x, b = pix / HW, hw = pix % HW;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, c0 = tid * 8;
  float x[MAXT][8];
#pragma unroll
  for (int t = 0; t < MAXT; ++t) {
    if (t < Tf) {
      const __nv_bfloat16* hp = split_hi(xt, ((long long)b * Tf + t) * HW + hw, D) + c0;
      const float4 a = load_split4(hp, hp + D), c = load_split4(hp + 4, hp + D + 4);
      x[t][0] = a.x; x[t][1] = a.y; x[t][2] = a.z; x[t][3] = a.w; x[t][4] = c.x; x[t][5] = c.y; x[t][6] = c.z; x[t][7] = c.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) x[t][e] = 0.f;
    }
  }
  float part[NH][MAXT];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const float4 u0 = __ldg(reinterpret_cast<const float4*>(U + h * D + c0)), u1 = __ldg(reinterpret_cast<const float4*>(U + h * D + c0) + 1);
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      float d = x[t][0] * u0.x;
      d = fmaf(x[t][1], u0.y, d); d = fmaf(x[t][2], u0.z, d); d = fmaf(x[t][3], u0.w, d);
      d = fmaf(x[t][4], u1.x, d); d = fmaf(x[t][5], u1.y, d); d = fmaf(x[t][6], u1.z, d); d = fmaf(x[t][7], u1.w, d);
      part[h][t] = warp_sum(d);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int h = 0; h < NH; ++h)
#pragma unroll
      for (int t = 0; t < MAXT; ++t) red[warp][h * MAXT + t] = part[h][t];
  }
  __syncthreads();
  if (tid < NH) {                                            // softmax over the Tf frames of head tid
    float sc[MAXT], m = -INFINITY;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += red[w][tid * MAXT + t];
      sc[t] = t < Tf ? v : -INFINITY;
      m = fmaxf(m, sc[t]);
    }
    float l = 0.f;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) { sc[t] = expf(sc[t] - m); l += sc[t]; }
    const float inv = 1.f / l;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) prob[tid * MAXT + t] = sc[t] * inv;
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const float pr = prob[h * MAXT + t];                   // 0 for t >= Tf
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaf(pr, x[t][e], o[e]);
    }
    __nv_bfloat16* hp = split_hi(y, (long long)h * Mp + pix, D) + c0;
    store_split4(hp, hp + D, make_float4(o[0], o[1], o[2], o[3]));
    store_split4(hp + 4, hp + D + 4, make_float4(o[4], o[5], o[6], o[7]));
  }
}

cudaError_t launch_pool_mix(const void* xt_split, const float* U, void* y_split, int B, int Tf, int HW, long long group_stride,
                            cudaStream_t st) {
  if (Tf < 1 || Tf > 8) return cudaErrorInvalidValue;
  if (Tf <= 4) return launch_pdl(pool_mix_kernel<4>, dim3(B * HW), dim3(256), 0, st, xt_split, U, y_split, B, Tf, HW, group_stride);
  return launch_pdl(pool_mix_kernel<8>, dim3(B * HW), dim3(256), 0, st, xt_split, U, y_split, B, Tf, HW, group_stride);
}

// AdaptiveAvgPool3d(1) over all positions (tuber_ava.py:48,124): split [B,N,C] -> fp32 [B,C].  A CTA = (clip, 64 channels):
// 16 lanes x 4 channels (8-byte loads of each plane) x 16 row groups, every row group summing its rows i = g, g+16, ... in order,
// then the 16 partial sums are added in order -- 16 independent load chains per channel instead of one thread walking all N rows
// (53 us for the JHMDB head's 512 rows x 2048 channels per clip).
__global__ void __launch_bounds__(256)
global_avgpool_kernel(const void* __restrict__ in, float* __restrict__ out, int N, int C) {
  __shared__ float4 part[16][16];
  const int cq = threadIdx.x & 15, g = threadIdx.x >> 4;
  const int c = blockIdx.x * 64 + cq * 4, b = blockIdx.y;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
#pragma unroll 4
    for (int i = g; i < N; i += 16) {
      const __nv_bfloat16* hi = split_hi(in, (long long)b * N + i, C) + c;
      const float4 v = load_split4(hi, hi + C);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  part[g][cq] = acc;
  __syncthreads();
  if (g == 0 && c < C) {
    float4 t = part[0][cq];
#pragma unroll
    for (int j = 1; j < 16; ++j) { const float4 u = part[j][cq]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
    const float inv = 1.f / (float)N;
    *reinterpret_cast<float4*>(out + (long long)b * C + c) = make_float4(t.x * inv, t.y * inv, t.z * inv, t.w * inv);
  }
}

cudaError_t launch_global_avgpool(const void* in_split, float* out, int B, int N, int C, cudaStream_t st) {
  if (C % 4 != 0) return cudaErrorInvalidValue;
  dim3 grid(ceil_div(C, 64), B);
  global_avgpool_kernel<<<grid, 256, 0, st>>>(in_split, out, N, C);
  return cudaGetLastError();
}

// =============================================================================================
// fp32 CUDA-core GEMM: the tiny-N heads (tuber_ava.py:64-73, criterion.py:485-497) and the debugging
// cross-check of the tensor-core kernel.  64x64 tile, BK=16, 256 threads, 4x4 micro-tile, register
// prefetch of the next k-slab.  A / res / C may be fp32 or split-bf16.
// =============================================================================================
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return v;
}

__global__ void __launch_bounds__(256)
sgemm_kernel(GemmArgs p) {
  constexpr int BM = 64, BN = 64, BK = 16, LDS = BM + 4;
  __shared__ __align__(16) float As[BK][LDS];
  __shared__ __align__(16) float Ws[BK][LDS];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lr = tid >> 2, lk = (tid & 3) * 4;   // loader: row within tile, k offset
  const int ty = tid >> 4, tx = tid & 15;

  const int am = m0 + lr, wn = n0 + lr;
  const bool a_ok = am < p.M, w_ok = wn < p.N;
  const int KT = p.K + p.Kb;                     // [A | Ab] along K
  const float* a_ptr = nullptr;
  const __nv_bfloat16* a_hi = nullptr;
  const float* b_ptr = nullptr;
  const __nv_bfloat16* b_hi = nullptr;
  if (p.a_fmt == FMT_F32) {
    a_ptr = reinterpret_cast<const float*>(p.A) + (long long)am * p.lda + lk;
    if (p.Ab) b_ptr = reinterpret_cast<const float*>(p.Ab) + (long long)am * p.ldb + lk;
  } else {
    a_hi = split_hi(p.A, am, p.lda) + lk;
    if (p.Ab) b_hi = split_hi(p.Ab, am, p.ldb) + lk;
  }
  const float* w_ptr = p.Wf + (long long)wn * KT + lk;

  auto load_a = [&](int k0) -> float4 {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a_ok) {
      if (k0 < p.K) v = a_ptr ? __ldg(reinterpret_cast<const float4*>(a_ptr + k0)) : load_split4(a_hi + k0, a_hi + p.lda + k0);
      else v = b_ptr ? __ldg(reinterpret_cast<const float4*>(b_ptr + k0 - p.K)) : load_split4(b_hi + k0 - p.K, b_hi + p.ldb + k0 - p.K);
    }
    return v;
  };
  auto load_w = [&](int k0) -> float4 {
    return w_ok ? __ldg(reinterpret_cast<const float4*>(w_ptr + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 ra = load_a(0), rw = load_w(0);
  for (int k0 = 0; k0 < KT; k0 += BK) {
    As[lk + 0][lr] = ra.x; As[lk + 1][lr] = ra.y; As[lk + 2][lr] = ra.z; As[lk + 3][lr] = ra.w;
    Ws[lk + 0][lr] = rw.x; Ws[lk + 1][lr] = rw.y; Ws[lk + 2][lr] = rw.z; Ws[lk + 3][lr] = rw.w;
    __syncthreads();
    if (k0 + BK < KT) {
      ra = load_a(k0 + BK);
      rw = load_w(k0 + BK);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    const long long rr = p.res_mod > 0 ? m % p.res_mod : m;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.scale) v *= __ldg(p.scale + n);
      if (p.shift) v += __ldg(p.shift + n);
      if (p.res) {
        if (p.res_fmt == FMT_F32) {
          v += __ldg(reinterpret_cast<const float*>(p.res) + rr * p.ldr + n);
        } else {
          const __nv_bfloat16* h = split_hi(p.res, rr, p.ldr);
          v += __bfloat162float(h[n]) + __bfloat162float(h[p.ldr + n]);
        }
      }
      v = apply_act(v, p.act);
      void* cf = p.c_fmt == FMT_F32 ? p.C : p.C2;
      void* cs = p.c_fmt == FMT_F32 ? p.C2 : p.C;
      const int ldf = p.c_fmt == FMT_F32 ? p.ldc : p.ldc2, lds = p.c_fmt == FMT_F32 ? p.ldc2 : p.ldc;
      if (cf) reinterpret_cast<float*>(cf)[(long long)m * ldf + n] = v;
      if (cs) {
        __nv_bfloat16 hi, mid;
        split_bf16(v, hi, mid);
        __nv_bfloat16* h = split_hi(cs, m, lds);
        h[n] = hi;
        h[lds + n] = mid;
      }
    }
  }
}

template <int LANES> __global__ void head_gemm_kernel(GemmArgs p);

cudaError_t launch_sgemm(const GemmArgs& a, cudaStream_t st) {
  if (a.K % 16 != 0 || a.Kb % 16 != 0 || a.M <= 0 || a.N <= 0 || a.Wf == nullptr) return cudaErrorInvalidValue;
  if (a.N <= 96 && !a.Ab && !a.res && !a.C2 && a.c_fmt == FMT_F32 && a.K <= 2048 && a.lda % 4 == 0) {
    if ((long long)a.M * a.N <= 1024 && a.K >= 1024)      // few outputs, long K: a warp per output
      return launch_pdl(head_gemm_kernel<32>, dim3(ceil_div((long long)a.M * a.N * 32, 256)), dim3(256), 0, st, a);
    const long long total = (long long)a.M * a.N * 8;
    return launch_pdl(head_gemm_kernel<8>, dim3(ceil_div(total, 256)), dim3(256), 0, st, a);
  }
  dim3 grid(ceil_div(a.N, 64), ceil_div(a.M, 64));
  sgemm_kernel<<<grid, 256, 0, st>>>(a);
  return cudaGetLastError();
}

// =============================================================================================
// LayerNorm (eps 1e-5) over C = 32*4*NV elements, one warp per row, optional fused residual add,
// fp32 and/or split-bf16 outputs, optional output-row remap (stacks decoder layers, concat).
// =============================================================================================
template <int NV, bool CHAIN>   // float4 chunks per lane: C = 128 * NV; CHAIN: second norm on the first result (LnArgs::gamma2)
__global__ void __launch_bounds__(256)
layernorm_kernel(LnArgs p) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= p.rows) return;
  const long long r = warp;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int c = i * 128 + lane * 4;
    if (p.x_fmt == FMT_F32) {
      v[i] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.x) + r * p.ldx + c));
      for (int s = 1; s < p.x_parts; ++s) {                // split-K partial sums, added in a fixed order
        const float4 u = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.x) + (r + s * p.x_part_stride) * p.ldx + c));
        v[i].x += u.x; v[i].y += u.y; v[i].z += u.z; v[i].w += u.w;
      }
    } else {
      const __nv_bfloat16* h = split_hi(p.x, r, p.ldx) + c;
      v[i] = load_split4(h, h + p.ldx);
    }
    if (p.res) {
      float4 u;
      if (p.res_fmt == FMT_F32) {
        u = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res) + r * p.ldr + c));
      } else {
        const __nv_bfloat16* h = split_hi(p.res, r, p.ldr) + c;
        u = load_split4(h, h + p.ldr);
      }
      v[i].x += u.x; v[i].y += u.y; v[i].z += u.z; v[i].w += u.w;
    }
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) / (float)p.C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += a * a + b * b + c * c + d * d;
  }
  const float rstd = 1.f / sqrtf(warp_sum(q) / (float)p.C + p.eps);
  long long orow = p.rpg > 0 ? (r / p.rpg) * p.group_stride + (r % p.rpg) + p.row_off : r + p.row_off;
  constexpr bool chain = CHAIN;
  float s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int c = i * 128 + lane * 4;
    float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + c));
    float4 b = __ldg(reinterpret_cast<const float4*>(p.beta + c));
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + orow * p.ldo + c) = o;
    if (p.out_split) {
      __nv_bfloat16* hi = split_hi(p.out_split, chain ? r : orow, p.lds) + p.split_col_off + c;
      store_split4(hi, hi + p.lds, o);
    }
    if constexpr (CHAIN) {                                  // the second norm sees the first result as stored (hi + mid)
      uint32_t h0, m0, h1, m1;
      split_bf16x2(o.x, o.y, h0, m0);
      split_bf16x2(o.z, o.w, h1, m1);
      v[i].x = bf16_lo_to_f32(h0) + bf16_lo_to_f32(m0); v[i].y = bf16_hi_to_f32(h0) + bf16_hi_to_f32(m0);
      v[i].z = bf16_lo_to_f32(h1) + bf16_lo_to_f32(m1); v[i].w = bf16_hi_to_f32(h1) + bf16_hi_to_f32(m1);
      s2 += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  if constexpr (!CHAIN) return;
  const float mean2 = warp_sum(s2) / (float)p.C;
  float q2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float a = v[i].x - mean2, b = v[i].y - mean2, c = v[i].z - mean2, d = v[i].w - mean2;
    q2 += a * a + b * b + c * c + d * d;
  }
  const float rstd2 = 1.f / sqrtf(warp_sum(q2) / (float)p.C + p.eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int c = i * 128 + lane * 4;
    float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma2 + c));
    float4 b = __ldg(reinterpret_cast<const float4*>(p.beta2 + c));
    float4 o;
    o.x = (v[i].x - mean2) * rstd2 * g.x + b.x;
    o.y = (v[i].y - mean2) * rstd2 * g.y + b.y;
    o.z = (v[i].z - mean2) * rstd2 * g.z + b.z;
    o.w = (v[i].w - mean2) * rstd2 * g.w + b.w;
    __nv_bfloat16* hi = split_hi(p.out2_split, orow, p.lds2) + c;
    store_split4(hi, hi + p.lds2, o);
  }
}

// =============================================================================================
// Tiny-N linear heads (class_embed_b N = 3 / 2, bbox_embed.layers.2 N = 4 + sigmoid, class_fc N = 80 / 22;
// tuber_ava.py:64-73,121-125,141-142): one thread per output element, K <= 2048, A fp32 or split.  The tiled SIMT GEMM
// above needs 64 x 64 tiles to be efficient and took ~20 us for these 0.1 - 30 MFLOP products.
// =============================================================================================
template <int LANES>   // lanes per output element: 8, or 32 for the products with few outputs and a long K (JHMDB clip head: 8 x 2, K = 2048)
__global__ void __launch_bounds__(256)
head_gemm_kernel(GemmArgs p) {
  pdl_trigger();
  pdl_wait();
  // LANES lanes per output element: lane s takes the float4 chunks s, s+LANES, ... of the K axis, so the lanes read contiguous
  // runs of the weight row and of the activation row per step; xor-shuffle reduction at the end
  constexpr int SH = LANES == 32 ? 5 : 3;
  const long long idx = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> SH;
  const int s = threadIdx.x & (LANES - 1);
  const bool live = idx < (long long)p.M * p.N;
  const int n = live ? (int)(idx % p.N) : 0;
  const long long m = live ? idx / p.N : 0;
  const float4* w = reinterpret_cast<const float4*>(p.Wf + (long long)n * p.K);
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  if (p.a_fmt == FMT_F32) {
    const float4* a = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.A) + m * p.lda);
#pragma unroll 4
    for (int k = s; k < p.K / 4; k += LANES) {
      const float4 x = __ldg(a + k), y = __ldg(w + k);
      acc0 = fmaf(x.x, y.x, acc0); acc1 = fmaf(x.y, y.y, acc1); acc2 = fmaf(x.z, y.z, acc2); acc3 = fmaf(x.w, y.w, acc3);
    }
  } else {
    const __nv_bfloat16* h = split_hi(p.A, m, p.lda);
#pragma unroll 4
    for (int k = s; k < p.K / 4; k += LANES) {
      const float4 x = load_split4(h + 4 * k, h + p.lda + 4 * k), y = __ldg(w + k);
      acc0 = fmaf(x.x, y.x, acc0); acc1 = fmaf(x.y, y.y, acc1); acc2 = fmaf(x.z, y.z, acc2); acc3 = fmaf(x.w, y.w, acc3);
    }
  }
  float v = (acc0 + acc1) + (acc2 + acc3);
#pragma unroll
  for (int o = 1; o < LANES; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (!live || s != 0) return;
  if (p.scale) v *= __ldg(p.scale + n);
  if (p.shift) v += __ldg(p.shift + n);
  if (p.act == ACT_RELU) v = fmaxf(v, 0.f);
  else if (p.act == ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
  reinterpret_cast<float*>(p.C)[m * p.ldc + n] = v;
}

cudaError_t launch_layernorm(const LnArgs& a, cudaStream_t st) {
  int grid = ceil_div((long long)a.rows * 32, 256);
  if (a.C == 256)
    return a.gamma2 ? launch_pdl(layernorm_kernel<2, true>, dim3(grid), dim3(256), 0, st, a)
                    : launch_pdl(layernorm_kernel<2, false>, dim3(grid), dim3(256), 0, st, a);
  else if (a.C == 2048)
    // (the chained variant holds 168 registers: 128-thread CTAs keep 12 warps per SM resident instead of 8)
    return a.gamma2 ? launch_pdl(layernorm_kernel<16, true>, dim3(ceil_div((long long)a.rows * 32, 128)), dim3(128), 0, st, a)
                    : launch_pdl(layernorm_kernel<16, false>, dim3(grid), dim3(256), 0, st, a);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// =============================================================================================
// Attention cores.  Q/K/V are fp32 slices of projection buffers; a sequence is addressed through
// a SeqMap so the same kernels serve per-clip, per-frame and per-pixel (strided) sequences.
// =============================================================================================
__device__ __forceinline__ long long seq_row0(const SeqMap& m, int n) {
  return (long long)(n / m.inner) * m.outer + (long long)(n % m.inner) * m.inner_stride;
}

// (b) any head_dim = 32*DPL, one warp per (sequence, head, query), keys streamed from global:
//     for the tiny-S sites (S = T' = 4 temporal attention, 15-query decoder self-attention,
//     the decode pool's 1x4 cross-attention with head_dim 256).
template <int DPL>
__global__ void __launch_bounds__(128)
attn_warp_kernel(AttnArgs p) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)p.NB * p.H * p.L;
  if (w >= total) return;
  const int l = (int)(w % p.L);
  const int h = (int)((w / p.L) % p.H);
  const int n = (int)(w / ((long long)p.L * p.H));
  const int D = 32 * DPL;
  const float* qp = p.q + (seq_row0(p.qm, n) + (long long)l * p.qm.step) * p.ldq + h * D;
  const long long k0 = seq_row0(p.km, n);
  const uint8_t* mrow = p.kpm ? p.kpm + (long long)(n / p.kpm_div) * p.S : nullptr;
  float q[DPL], acc[DPL];
#pragma unroll
  for (int i = 0; i < DPL; ++i) {
    q[i] = __ldg(qp + lane + 32 * i) * p.scale;
    acc[i] = 0.f;
  }
  float m = -INFINITY, lsum = 0.f;
  for (int j = 0; j < p.S; ++j) {
    if (mrow && mrow[j]) continue;
    const long long krow = k0 + (long long)j * p.km.step;
    const float* kp = p.k + krow * p.ldk + h * D;
    const float* vp = p.v + krow * p.ldv + h * D;
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < DPL; ++i) dot = fmaf(q[i], __ldg(kp + lane + 32 * i), dot);
    dot = warp_sum(dot);
    const float mnew = fmaxf(m, dot);
    const float corr = expf(m - mnew);
    const float pr = expf(dot - mnew);
    lsum = lsum * corr + pr;
    m = mnew;
#pragma unroll
    for (int i = 0; i < DPL; ++i) acc[i] = fmaf(pr, __ldg(vp + lane + 32 * i), acc[i] * corr);
  }
  const float inv = 1.f / lsum;
  const long long orow = seq_row0(p.om, n) + (long long)l * p.om.step;
#pragma unroll
  for (int i = 0; i < DPL; ++i) {
    float o = acc[i] * inv;
    int c = h * D + lane + 32 * i;
    if (p.o_f32) p.o_f32[orow * p.ldo + c] = o;
    if (p.o_split) {
      __nv_bfloat16 hi, mid;
      split_bf16(o, hi, mid);
      __nv_bfloat16* hp = split_hi(p.o_split, orow, p.ldo);
      hp[c] = hi;
      hp[p.ldo + c] = mid;
    }
  }
}

// (c) head_dim 32, L and S beyond a handful: a CTA = one (sequence, head, 32-query tile), 4 warps x 8 queries.
//     Keys / values stream through shared memory in chunks of 32 (cp.async, double buffered).  Scores: lane =
//     key, the key's 32 dims live in registers and the 8 queries are broadcast from shared memory; online
//     softmax with warp reductions; P.V: lane = output dim, the chunk's V column lives in registers and the
//     probabilities are broadcast from shared memory.  ~3.5 instructions per (query, key) pair.
namespace attile {
constexpr int QT = 32, KC = 32, D = 32, QW = 8;           // queries per CTA, keys per chunk, head dim, queries per warp
}
__global__ void __launch_bounds__(128)
attn_tile_kernel(AttnArgs p) {
  using namespace attile;
  __shared__ __align__(16) float Qs[QT][D];
  __shared__ __align__(16) float Ks[2][KC][D + 4];         // +4: lane-strided row reads stay conflict free with 16 B alignment
  __shared__ __align__(16) float Vs[2][KC][D];
  __shared__ __align__(16) float Ps[4][QW][KC];
  __shared__ uint8_t Ms[2][KC];
  const int n = blockIdx.z, h = blockIdx.y, l0 = blockIdx.x * QT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long q0 = seq_row0(p.qm, n), k0 = seq_row0(p.km, n);
  const uint8_t* mrow = p.kpm ? p.kpm + (long long)(n / p.kpm_div) * p.S : nullptr;

  for (int i = tid; i < QT * (D / 4); i += 128) {
    const int q = i / (D / 4), d4 = i % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (l0 + q < p.L) {
      v = __ldg(reinterpret_cast<const float4*>(p.q + (q0 + (long long)(l0 + q) * p.qm.step) * p.ldq + h * D) + d4);
      v.x *= p.scale; v.y *= p.scale; v.z *= p.scale; v.w *= p.scale;
    }
    *reinterpret_cast<float4*>(&Qs[q][d4 * 4]) = v;
  }
  auto load_chunk = [&](int c, int buf) {
    const int s0 = c * KC;
    for (int i = tid; i < KC * (D / 4); i += 128) {
      const int j = i / (D / 4), d4 = i % (D / 4);
      const bool ok = s0 + j < p.S;
      const long long krow = k0 + (long long)(ok ? s0 + j : 0) * p.km.step;
      const int nb = ok ? 16 : 0;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(&Ks[buf][j][d4 * 4])),
                   "l"(p.k + krow * p.ldk + h * D + d4 * 4), "r"(nb) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(&Vs[buf][j][d4 * 4])),
                   "l"(p.v + krow * p.ldv + h * D + d4 * 4), "r"(nb) : "memory");
    }
    if (tid < KC) Ms[buf][tid] = (s0 + tid >= p.S) || (mrow && mrow[s0 + tid]);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int nchunks = (p.S + KC - 1) / KC;
  load_chunk(0, 0);
  float acc[QW], mx[QW], ls[QW];
#pragma unroll
  for (int q = 0; q < QW; ++q) { acc[q] = 0.f; mx[q] = -INFINITY; ls[q] = 0.f; }
  const bool warp_active = l0 + warp * QW < p.L;

  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) {
      load_chunk(c + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (warp_active) {
      float kr[D];
#pragma unroll
      for (int d4 = 0; d4 < D / 4; ++d4) {
        const float4 t = *reinterpret_cast<const float4*>(&Ks[buf][lane][d4 * 4]);
        kr[4 * d4] = t.x; kr[4 * d4 + 1] = t.y; kr[4 * d4 + 2] = t.z; kr[4 * d4 + 3] = t.w;
      }
      const bool masked = Ms[buf][lane];
#pragma unroll
      for (int q = 0; q < QW; ++q) {
        const float* qv = &Qs[warp * QW + q][0];
        float sdot = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < D / 4; ++d4) {
          const float4 t = *reinterpret_cast<const float4*>(qv + d4 * 4);
          sdot = fmaf(t.x, kr[4 * d4], sdot); sdot = fmaf(t.y, kr[4 * d4 + 1], sdot);
          sdot = fmaf(t.z, kr[4 * d4 + 2], sdot); sdot = fmaf(t.w, kr[4 * d4 + 3], sdot);
        }
        if (masked) sdot = -INFINITY;
        const float cm = warp_max(sdot);
        const float mnew = fmaxf(mx[q], cm);
        const float pr = (mnew == -INFINITY) ? 0.f : expf(sdot - mnew);      // whole row masked so far: nothing to add
        const float corr = (mnew == -INFINITY) ? 1.f : expf(mx[q] - mnew);    // mx = -inf -> 0
        ls[q] = ls[q] * corr + warp_sum(pr);
        acc[q] *= corr;
        mx[q] = mnew;
        Ps[warp][q][lane] = pr;
      }
      __syncwarp();
      float vr[KC];
#pragma unroll
      for (int j = 0; j < KC; ++j) vr[j] = Vs[buf][j][lane];
#pragma unroll
      for (int q = 0; q < QW; ++q) {
        float a = acc[q];
#pragma unroll
        for (int j4 = 0; j4 < KC / 4; ++j4) {
          const float4 t = *reinterpret_cast<const float4*>(&Ps[warp][q][j4 * 4]);
          a = fmaf(t.x, vr[4 * j4], a); a = fmaf(t.y, vr[4 * j4 + 1], a);
          a = fmaf(t.z, vr[4 * j4 + 2], a); a = fmaf(t.w, vr[4 * j4 + 3], a);
        }
        acc[q] = a;
      }
    }
    __syncthreads();
  }
  if (!warp_active) return;
  const long long o0 = seq_row0(p.om, n);
#pragma unroll
  for (int q = 0; q < QW; ++q) {
    const int l = l0 + warp * QW + q;
    if (l >= p.L) break;
    const float o = acc[q] / ls[q];
    const long long orow = o0 + (long long)l * p.om.step;
    const int cidx = h * D + lane;
    if (p.o_f32) p.o_f32[orow * p.ldo + cidx] = o;
    if (p.o_split) {
      __nv_bfloat16 hi, mid;
      split_bf16(o, hi, mid);
      __nv_bfloat16* hp = split_hi(p.o_split, orow, p.ldo);
      hp[cidx] = hi;
      hp[p.ldo + cidx] = mid;
    }
  }
}

// which kernel launch_attention picks for a shape (unit tests assert the dispatch through tuber_op_attention_kernel)
const char* attention_kernel_name(const AttnArgs& a) {
  const char* fs = getenv("TUBER_ATTN_SIMT");
  const char* nt = getenv("TUBER_ATTN_NO_TC");
  const bool force_simt = fs && fs[0] == '1', no_tc = nt && nt[0] == '1';
  if (!force_simt && !no_tc && attention_tc_supported(a)) return "attn_tc_kernel";
  if (!force_simt && a.D == 32 && a.ldq % 4 == 0 && a.ldk % 4 == 0 && a.ldv % 4 == 0 && a.ldo % 4 == 0) {
    if (a.L <= 8 && a.S <= 8 && a.H % 8 == 0) return "attn_tiny_kernel";
    if (a.NB <= 65535 && a.H <= 65535) return (a.L <= 16 && a.S > 64) ? "attn_mma_split_kernel" : "attn_mma_kernel";
  }
  return "attn_simt_kernel";
}

cudaError_t launch_attention(const AttnArgs& a, cudaStream_t st) {
  static const bool force_simt = [] { const char* e = getenv("TUBER_ATTN_SIMT"); return e && e[0] == '1'; }();
  static const bool no_tc = [] { const char* e = getenv("TUBER_ATTN_NO_TC"); return e && e[0] == '1'; }();
  if (!force_simt && !no_tc && attention_tc_supported(a)) return launch_attention_tc(a, st);
  if (!force_simt && a.D == 32) {
    cudaError_t e = launch_attention_mma(a, st);
    if (e != cudaErrorNotSupported) return e;
  }
  return launch_attention_simt(a, st);
}

cudaError_t launch_attention_simt(const AttnArgs& a, cudaStream_t st) {
  if (a.D % 32 != 0 || a.NB <= 0 || a.L <= 0 || a.S <= 0) return cudaErrorInvalidValue;
  const bool small = a.S <= 16 || a.D != 32;
  if (!small) {
    dim3 grid(ceil_div(a.L, attile::QT), a.H, a.NB);
    attn_tile_kernel<<<grid, 128, 0, st>>>(a);
  } else {
    long long warps = (long long)a.NB * a.H * a.L;
    int grid = ceil_div(warps * 32, 128);
    if (a.D == 32)
      attn_warp_kernel<1><<<grid, 128, 0, st>>>(a);
    else if (a.D == 256)
      attn_warp_kernel<8><<<grid, 128, 0, st>>>(a);
    else
      return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// =============================================================================================
// Post-processing fused with the packing of the detection rows the evaluation loop writes
// (reference models/criterion.py:413-482 PostProcess / PostProcessAVA; utils/video_action_recognition.py:311-346,411-415:
// one text line per (clip, query) = boxes xyxy scaled | class scores | foreground probability):
//   AVA   : p = softmax(logits_b)[1]; scores = sigmoid(logits) * (p > 0.8 ? p : 0)
//   other : scores = softmax(logits); p = softmax(logits_b[clip])[1]   (logits_b is per clip, 2 values)
//   boxes : (cx,cy,w,h) in [0,1] -> (x1,y1,x2,y2) * (W,H,W,H) with (H,W) = sizes[clip]
// One warp per (clip, query); out row = [4 | C | 1] floats.
// =============================================================================================
__global__ void __launch_bounds__(128)
postprocess_kernel(const float* __restrict__ logits, long long l_sb, const float* __restrict__ boxes, long long b_sb,
                   const float* __restrict__ logits_b, long long lb_sb, const float* __restrict__ sizes, float* __restrict__ out,
                   int B, int Q, int C, int ava) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B * Q) return;
  const int b = w / Q, q = w % Q;
  const float* lg = logits + b * l_sb + (long long)q * C;
  const float* bx = boxes + b * b_sb + (long long)q * 4;
  float* o = out + (long long)w * (C + 5);
  float p;
  if (ava) {
    const float* lb = logits_b + b * lb_sb + (long long)q * 3;
    const float a0 = lb[0], a1 = lb[1], a2 = lb[2], m = fmaxf(a0, fmaxf(a1, a2));
    const float e0 = expf(a0 - m), e1 = expf(a1 - m), e2 = expf(a2 - m);
    p = e1 / (e0 + e1 + e2);
    const float gate = p > 0.8f ? p : 0.f;
    for (int c = lane; c < C; c += 32) o[4 + c] = gate / (1.f + expf(-lg[c]));
  } else {
    const float* lb = logits_b + b * lb_sb;
    const float m2 = fmaxf(lb[0], lb[1]), e0 = expf(lb[0] - m2), e1 = expf(lb[1] - m2);
    p = e1 / (e0 + e1);
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, lg[c]);
    m = warp_max(m);
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) sum += expf(lg[c] - m);
    sum = warp_sum(sum);
    for (int c = lane; c < C; c += 32) o[4 + c] = expf(lg[c] - m) / sum;
  }
  if (lane == 0) {
    const float H = sizes[2 * b], W = sizes[2 * b + 1];
    const float cx = bx[0], cy = bx[1], bw = bx[2], bh = bx[3];
    o[0] = (cx - 0.5f * bw) * W; o[1] = (cy - 0.5f * bh) * H; o[2] = (cx + 0.5f * bw) * W; o[3] = (cy + 0.5f * bh) * H;
    o[4 + C] = p;
  }
}

cudaError_t launch_postprocess(const float* logits, long long l_sb, const float* boxes, long long b_sb, const float* logits_b,
                               long long lb_sb, const float* sizes, float* out, int B, int Q, int C, int ava, cudaStream_t st) {
  if (B <= 0 || Q <= 0 || C <= 0) return cudaErrorInvalidValue;
  postprocess_kernel<<<ceil_div((long long)B * Q * 32, 128), 128, 0, st>>>(logits, l_sb, boxes, b_sb, logits_b, lb_sb, sizes, out, B, Q, C, ava);
  return cudaGetLastError();
}

// =============================================================================================
// uint8 frames -> normalised fp32 clip: the reference's ToTensor + Normalize + stack + permute
// (datasets/video_transforms.py:294-296,308-314; datasets/ava_frame.py:71-72) on the device, so that the host ships 3 bytes per
// pixel instead of 12.  in [B][T*H*W pixels][3] (decoded RGB frames, HWC), out [B][3][T*H*W] = NestedTensor.tensors.
// lut[c][u] = ((float)u / 255 - mean[c]) / std[c], evaluated on the host in fp32 exactly as torch evaluates it: bit-identical
// to the reference transform.  HBM bound: 3 B read + 12 B written per pixel.
// =============================================================================================
__global__ void __launch_bounds__(256)
normalize_u8_kernel(const uint8_t* __restrict__ in, const float* __restrict__ lut, float* __restrict__ out, long long P, int B, int vec) {
  __shared__ float s_lut[768];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = __ldg(lut + i);
  __syncthreads();
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
  if (vec) {                                                         // P % 4 == 0, aligned bases: 4 pixels = 12 bytes = three words per thread step
    const long long G = P >> 2, total = G * B;
    for (long long g = tid; g < total; g += nthr) {
      const long long b = g / G, p4 = g - b * G;
      const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (b * P + 4 * p4) * 3);
      const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);   // r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
      float* o = out + b * 3 * P + 4 * p4;
      const float4 r = make_float4(s_lut[w0 & 255u], s_lut[w0 >> 24], s_lut[(w1 >> 16) & 255u], s_lut[(w2 >> 8) & 255u]);
      const float4 gg = make_float4(s_lut[256 + ((w0 >> 8) & 255u)], s_lut[256 + (w1 & 255u)], s_lut[256 + (w1 >> 24)],
                                    s_lut[256 + ((w2 >> 16) & 255u)]);
      const float4 bb = make_float4(s_lut[512 + ((w0 >> 16) & 255u)], s_lut[512 + ((w1 >> 8) & 255u)], s_lut[512 + (w2 & 255u)],
                                    s_lut[512 + (w2 >> 24)]);
      *reinterpret_cast<float4*>(o) = r;
      *reinterpret_cast<float4*>(o + P) = gg;
      *reinterpret_cast<float4*>(o + 2 * P) = bb;
    }
  } else {
    const long long total = P * B;
    for (long long i = tid; i < total; i += nthr) {
      const long long b = i / P, px = i - b * P;
      const uint8_t* src = in + i * 3;
      float* o = out + b * 3 * P + px;
      o[0] = s_lut[src[0]];
      o[P] = s_lut[256 + src[1]];
      o[2 * P] = s_lut[512 + src[2]];
    }
  }
}

cudaError_t launch_normalize_u8(const uint8_t* frames, const float* lut, float* out, int B, long long pixels_per_clip, cudaStream_t st) {
  if (B <= 0 || pixels_per_clip <= 0) return cudaErrorInvalidValue;
  const int vec = (pixels_per_clip & 3) == 0 && (reinterpret_cast<uintptr_t>(frames) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const long long work = vec ? (pixels_per_clip >> 2) * B : pixels_per_clip * B;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long want = ceil_div(work, 256);
  const int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);   // 8 resident CTAs per SM, grid-stride
  normalize_u8_kernel<<<grid, 256, 0, st>>>(frames, lut, out, pixels_per_clip, B, vec);
  return cudaGetLastError();
}

// =============================================================================================
// Padding mask at feature resolution and the 3-D sine position code
// =============================================================================================
// nearest-neighbour resize (B,H,W) -> (B,T,Hf,Wf), repeated over T  (backbone_builder.py:85-86)
__global__ void mask_resize_kernel(const uint8_t* __restrict__ mask, uint8_t* __restrict__ fmask, int B, int H,
                                   int W, int T, int Hf, int Wf) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * T * Hf * Wf;
  if (idx >= total) return;
  int x = (int)(idx % Wf), y = (int)((idx / Wf) % Hf);
  long long b = idx / ((long long)Wf * Hf * T);
  const float sh = (float)H / (float)Hf, sw = (float)W / (float)Wf;
  int sy = min((int)floorf((float)y * sh), H - 1);
  int sx = min((int)floorf((float)x * sw), W - 1);
  fmask[idx] = mask ? mask[(b * H + sy) * W + sx] : 0;
}

cudaError_t launch_mask_resize(const uint8_t* mask, uint8_t* fmask, int B, int H, int W, int T, int Hf, int Wf,
                               cudaStream_t st) {
  long long total = (long long)B * T * Hf * Wf;
  mask_resize_kernel<<<ceil_div(total, 256), 256, 0, st>>>(mask, fmask, B, H, W, T, Hf, Wf);
  return cudaGetLastError();
}

// PositionEmbeddingSine_3D(normalize=True)  (position_encoding.py:32-72), token-major output
// pos[b, (t,y,x), c], c in [0,nt) from t, [nt,nt+ns) from y, [nt+ns,nt+2ns) from x.
__global__ void posenc_kernel(const uint8_t* __restrict__ fmask, const float* __restrict__ dim_t,
                              const float* __restrict__ dim_s, float* __restrict__ pos, int B, int T, int H, int W,
                              int nt, int ns) {
  const int d = nt + 2 * ns;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * T * H * W * d;
  if (idx >= total) return;
  int c = (int)(idx % d);
  long long tok = idx / d;
  int x = (int)(tok % W), y = (int)((tok / W) % H), t = (int)((tok / ((long long)W * H)) % T);
  long long b = tok / ((long long)W * H * T);
  const uint8_t* mb = fmask + b * T * H * W;
  float e = 0.f, last = 0.f, div;
  int ci;
  if (c < nt) {
    for (int i = 0; i < T; ++i) {
      float nm = mb[((long long)i * H + y) * W + x] ? 0.f : 1.f;
      last += nm;
      if (i <= t) e += nm;
    }
    ci = c; div = dim_t[ci];
  } else if (c < nt + ns) {
    for (int i = 0; i < H; ++i) {
      float nm = mb[((long long)t * H + i) * W + x] ? 0.f : 1.f;
      last += nm;
      if (i <= y) e += nm;
    }
    ci = c - nt; div = dim_s[ci];
  } else {
    for (int i = 0; i < W; ++i) {
      float nm = mb[((long long)t * H + y) * W + i] ? 0.f : 1.f;
      last += nm;
      if (i <= x) e += nm;
    }
    ci = c - nt - ns; div = dim_s[ci];
  }
  const float two_pi = 6.283185307179586f;
  float ph = (e / (last + 1e-6f) * two_pi) / div;
  pos[idx] = (ci & 1) ? cosf(ph) : sinf(ph);
}

cudaError_t launch_posenc(const uint8_t* fmask, const float* dim_t, const float* dim_s, float* pos, int B, int T,
                          int H, int W, int nt, int ns, cudaStream_t st) {
  long long total = (long long)B * T * H * W * (nt + 2 * ns);
  posenc_kernel<<<ceil_div(total, 256), 256, 0, st>>>(fmask, dim_t, dim_s, pos, B, T, H, W, nt, ns);
  return cudaGetLastError();
}

// =============================================================================================
// fp32 <-> split-bf16
// =============================================================================================
__global__ void to_split_kernel(const float* __restrict__ in, int ldi, void* __restrict__ out, int ldo,
                                long long rows, int cols) {
  const int c4n = cols >> 2;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * c4n) return;
  int c4 = (int)(idx % c4n);
  long long r = idx / c4n;
  float4 v = __ldg(reinterpret_cast<const float4*>(in + r * ldi) + c4);
  __nv_bfloat16* hi = split_hi(out, r, ldo) + c4 * 4;
  store_split4(hi, hi + ldo, v);
}
__global__ void from_split_kernel(const void* __restrict__ in, int ldi, float* __restrict__ out, int ldo,
                                  long long rows, int cols) {
  const int c4n = cols >> 2;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * c4n) return;
  int c4 = (int)(idx % c4n);
  long long r = idx / c4n;
  const __nv_bfloat16* hi = split_hi(in, r, ldi) + c4 * 4;
  reinterpret_cast<float4*>(out + r * ldo)[c4] = load_split4(hi, hi + ldi);
}

cudaError_t launch_to_split(const float* in, int ldi, void* out, int ldo, long long rows, int cols, cudaStream_t st) {
  to_split_kernel<<<ceil_div(rows * (cols / 4), 256), 256, 0, st>>>(in, ldi, out, ldo, rows, cols);
  return cudaGetLastError();
}
cudaError_t launch_from_split(const void* in, int ldi, float* out, int ldo, long long rows, int cols,
                              cudaStream_t st) {
  from_split_kernel<<<ceil_div(rows * (cols / 4), 256), 256, 0, st>>>(in, ldi, out, ldo, rows, cols);
  return cudaGetLastError();
}
