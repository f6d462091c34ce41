// Host-side launch interface of the tuber_b200 kernels (internal; the public C-ABI is
// include/tuber_b200.h).  Every launcher is asynchronous on the given stream and returns the
// cudaError_t of the launch.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

// ---- stem (ir_CSN_152.py:109-122,176-179), stem_tc.cu --------------------------------------
// filter (64,3,3,7,7) fp32 = [oc][441] -> packed split bf16 [2][64][576] in the kernel's K order
cudaError_t launch_stem_pack_weight(const float* w_oc441, void* out, cudaStream_t st);
// x NCDHW fp32 (B,3,T,H,W); wpk from launch_stem_pack_weight.  When stem_pool_is_fused(W1) the kernel writes the
// max-pooled result (split [B,T,H2,W2,64]) itself and y is not touched; otherwise it writes the conv rows
// y = relu(conv*scale+shift) (fp32 NDHWC [B,T,H1,W1,64]) and launch_maxpool_hw must follow.
bool stem_pool_is_fused(int W1);
// frames_u8 / lut (optional, fused-pool pair kernel only): decoded uint8 frames (B,T,H,W,3) + the value table [3][256] of the reference's
// ToTensor + Normalize instead of x -- the stem's staging threads normalise while they fill the input ring
cudaError_t launch_stem_conv(const float* x, const void* wpk, const float* scale, const float* shift, float* y, void* pooled,
                             int B, int T, int H, int W, int H1, int W1, cudaStream_t st, const uint8_t* frames_u8 = nullptr,
                             const float* lut = nullptr);
// (1,3,3)/s(1,2,2)/p(0,1,1) max pool, fp32 [BT,H1,W1,C] -> split [BT,H2,W2,C]
cudaError_t launch_maxpool_hw(const float* in, void* out_split, int BT, int H1, int W1, int H2, int W2,
                              int C, cudaStream_t st);

// ---- depthwise 3x3x3 (ir_CSN_152.py:48-51) + BN + ReLU -----------------------------------
// in fp32 [B,Ti,Hi,Wi,C]; wpk [27][C] (tap = (kt*3+kh)*3+kw); out split [B,To,Ho,Wo,C]
cudaError_t launch_dwconv(const float* in, const float* wpk, const float* scale, const float* shift,
                          void* out_split, int B, int Ti, int Hi, int Wi, int C, int st_t, int st_s,
                          int To, int Ho, int Wo, cudaStream_t st);

// strided voxel gather (rows of row_bytes, multiple of 16): out[b,to,ho,wo] = in[b,to*st_t,ho*st_s,wo*st_s]
cudaError_t launch_gather_rows(const void* in, void* out, int row_bytes, int B, int Ti, int Hi, int Wi,
                               int st_t, int st_s, int To, int Ho, int Wo, cudaStream_t st);
// generic frame slice: out[b, 0:Tn, hw] = in[b, t0:t0+Tn, hw]  (rows of row_bytes)
cudaError_t launch_slice_frames(const void* in, void* out, int row_bytes, int B, int Tin, int HW, int t0,
                                int Tn, cudaStream_t st);

// temporal pooling of split features [B,Tin,HW,C] -> split [B,Tout,HW,C], window k, stride k
cudaError_t launch_tpool(const void* in_split, void* out_split, int B, int Tin, int HW, int C, int k,
                         int Tout, int is_max, cudaStream_t st);
// decode pool: per pixel and head, scores u_h . x_t over the Tf frames, softmax, mixed token sum_t p x_t (see kernels_simt.cu)
//   xt split [B*Tf*HW, 2048], U fp32 [8][2048], y split [8*group_stride, 2048] (row = h*group_stride + pixel); Tf <= 8
cudaError_t launch_pool_mix(const void* xt_split, const float* U, void* y_split, int B, int Tf, int HW, long long group_stride,
                            cudaStream_t st);
// mean over N rows of a split tensor [B,N,C] -> fp32 [B,C]
cudaError_t launch_global_avgpool(const void* in_split, float* out, int B, int N, int C, cudaStream_t st);

// ---- GEMM: C = act( scale[n] * ([A | Ab] W^T)[m,n] + shift[n] + res[m % res_mod, n] ) -----------------------
// One argument block for both implementations:
//   launch_gemm_tc   tcgen05 bf16x3 tensor-core kernel (gemm_tc.cu): A must be FMT_SPLIT, needs Wp,
//                    N % 64 == 0, K % 64 == 0, Kb % 64 == 0.
//   launch_sgemm     fp32 CUDA-core kernel (kernels_simt.cu): any A format, needs Wf, K % 16 == 0; used for
//                    the tiny-N heads (N = 2, 3, 4, 80) and as the debugging cross-check of the TC kernel.
struct GemmArgs {
  const void* A; int a_fmt; int lda;         // [M, lda] fp32 or split
  const void* Ab; int ldb; int Kb;           // optional second operand concatenated along K (same format as A): W is [N, K+Kb]
  // Ab as a strided voxel gather (the shortcut of a striding bottleneck, ir_CSN_152.py:155-161): row m = output voxel (bt, ho, wo)
  // reads row (bt*st_t, ho*st_s, wo*st_s) of the split tensor Ab [BTi, Hi, Wi, ldb]; Wo = 0: plain rows.  Only the tcgen05
  // kernels (5-D TMA with element strides); see gemm_tc_strided_ab_ok
  int ab_Wo, ab_Ho, ab_Wi, ab_Hi, ab_BTi, ab_st_t, ab_st_s;
  const float* Wf;                           // fp32 [N, K+Kb] row-major
  const void* Wp;                            // packed split weights: bf16 [2][N][K+Kb]
  const float* scale; const float* shift;    // [N] each, nullable (1 / 0)
  const void* res; int res_fmt; int ldr; int res_mod;   // nullable residual, added before act
  void* C; int c_fmt; int ldc;
  void* C2; int ldc2;                        // optional second copy of C in the OTHER format
  int M, N, K;
  int act;                                   // ACT_NONE / ACT_RELU / ACT_SIGMOID (sigmoid: SIMT only)
  // grouped form (tcgen05 kernel only; 0 = plain): rows [g*group_rows, (g+1)*group_rows) of A are multiplied with rows
  // [g*N, (g+1)*N) of W (scale / shift likewise) and written to C[row - g*group_rows, g*N + n]; M = groups * group_rows
  int group_rows;
  // split-K (tcgen05 kernel only; 0/1 = off): the K range is cut into ksplit parts computed by different CTAs; part s is
  // written (fp32, no activation; shift and residual only in part 0) to rows [s*part_rows, s*part_rows + M) of C, and the
  // consumer (launch_layernorm with x_parts) adds the parts -- deterministic, unlike atomics
  int ksplit; int part_rows;
  int group_out_rows;                        // rows of C actually written per group (<= group_rows; 0 = group_rows): padding rows are clipped
};
cudaError_t launch_sgemm(const GemmArgs& a, cudaStream_t st);

// ---- LayerNorm over the last dim (C in {256, 2048}) of x (+ res) ---------------------------
struct LnArgs {
  const void* x; int x_fmt; int ldx;         // fp32 or split
  int x_parts; long long x_part_stride;      // > 1: x = sum of x_parts fp32 slabs, x_part_stride rows apart (split-K partial sums)
  const void* res; int res_fmt; int ldr;     // optional: normalise x + res (fp32 or split; ldr 0 = broadcast row)
  const float* gamma; const float* beta; float eps;
  int rows, C;
  float* out_f32; int ldo;                   // nullable; out row = (r / rpg) * group_stride + r % rpg + row_off
  int rpg; long long group_stride; long long row_off;
  void* out_split; int lds; int split_col_off;   // nullable; row r, columns split_col_off + [0,C)
  // chained second norm (the decoder's shared final norm on every layer output, transformer.py:116-126; the decode pool's
  // norm3 -> pool_decoder.norm): when gamma2 is set, the first result goes to out_split at row r (no remap), and
  // LayerNorm(first result as stored, gamma2, beta2) goes to out2_split at the remapped row
  const float* gamma2; const float* beta2;
  void* out2_split; int lds2;
};
cudaError_t launch_layernorm(const LnArgs& a, cudaStream_t st);

// ---- attention core: O = softmax(scale * Q K^T + mask) V, per (sequence n, head h) ---------
struct SeqMap {        // first row of sequence n = (n / inner) * outer + (n % inner) * inner_stride; token i at + i * step
  int inner; long long outer, inner_stride, step;
};
struct AttnArgs {
  const float* q; int ldq; SeqMap qm;
  const float* k; const float* v; int ldk, ldv; SeqMap km;
  float* o_f32; void* o_split; int ldo; SeqMap om;
  const uint8_t* kpm; int kpm_div;           // key padding mask [NB / kpm_div, S] (1 = ignore key) or null
  int NB, H, L, S, D;                        // D = head dim (32, or any multiple of 32 for the warp kernel)
  float scale;
  void* tc_scratch; int tc_shared_kv;        // attn_tc.cu only: optional workspace (attention_tc_scratch_bytes) / set by its launcher
};
// dispatcher: head_dim 32 -> the mma.sync kernels of attn_mma.cu (TUBER_ATTN_SIMT=1 in the environment forces the
// CUDA-core kernels, the cross-check of the tests); other head dims -> the CUDA-core warp kernel
cudaError_t launch_attention(const AttnArgs& a, cudaStream_t st);
cudaError_t launch_attention_simt(const AttnArgs& a, cudaStream_t st);
const char* attention_kernel_name(const AttnArgs& a);     // the kernel launch_attention dispatches this shape to
// attn_mma.cu; cudaErrorNotSupported when the shape is outside what it covers
cudaError_t launch_attention_mma(const AttnArgs& a, cudaStream_t st);
// attn_tc.cu: tcgen05 flash attention (head dim 32, L >= 64 queries and S >= 128 keys per sequence, optional key padding mask);
// TUBER_ATTN_NO_TC=1 in the environment keeps those shapes on the mma.sync kernel (the tests' cross-check)
bool attention_tc_supported(const AttnArgs& a);
bool attention_tc_wants_prep(const AttnArgs& a);
size_t attention_tc_scratch_bytes(const AttnArgs& a);
cudaError_t launch_attention_tc(const AttnArgs& a, cudaStream_t st);

// ---- the DETR decoder stack as one persistent cooperative kernel (decoder_mega.cu; transformer.py:98-127,218-249) ----------
// fp32 weights: *_w row-major [N, K] (the GEMM's Lin::wf), *_t K-major copies [K, N]; biases [N]; LayerNorm gamma / beta [256]
constexpr int DEC_MEGA_MAX_LAYERS = 8;
struct DecMegaLayer {
  const float* sa_in_w; const float* sa_in_b; const float* pq_sa;     // [3d, d], [3d], [Q, 3d] query_pos terms of q and k
  const float* sa_out_t; const float* sa_out_b; const float* n1_g; const float* n1_b;
  const float* ca_q_t; const float* ca_q_b; const float* pq_ca;       // [d, d] K-major, [d], [Q, d]
  const float* ca_out_t; const float* ca_out_b; const float* n2_g; const float* n2_b;
  const float* lin1_w; const float* lin1_b; const float* lin2_w; const float* lin2_b;   // [ff, d], [ff], [d, ff], [d]
  const float* n3_g; const float* n3_b;
};
struct DecMegaArgs {
  DecMegaLayer L[DEC_MEGA_MAX_LAYERS];
  int Ld, B, Q, Ntok, dim_ff, kv_ld, ntok_pad;
  float eps;
  const float* memkv;                        // [B*Ntok, kv_ld]: per layer [K | V] of the cross attention (kv_ld = Ld * 2d)
  const uint8_t* kpm;                        // key padding mask [B, Ntok] or null
  const float* dec0_c1; const float* dec0_qc;   // layer 0 folded at finalize ([d], [Q, d]); both set
  const float* nf_g; const float* nf_b;      // the decoder's shared final norm
  float* tgt;                                // [B*Q, d] fp32 state (scratch)
  float* qkv;                                // [B*Q, 3d] scratch
  float* part;                               // [dim_ff / 16][B*Q][d] feed-forward partial sums (scratch)
  void* hs;                                  // split [B, Ld, Q, d]: every layer's output after the final norm
  unsigned* barrier;                         // 4 bytes of device memory for the grid barrier
  unsigned long long* trace;                 // optional [1 + 4 * Ld]: %globaltimer of CTA 0 at the start and after every phase
};
bool decoder_mega_supported(int d_model, int nhead, int dim_ff, int Q, int Ntok);
cudaError_t launch_decoder_mega(DecMegaArgs a, cudaStream_t st);

// ---- post-processing + detection rows (criterion.py:413-482; video_action_recognition.py:411-415) -------------
// logits [B, Q, C] (clip stride l_sb floats), boxes [B, Q, 4] (b_sb), logits_b [B, Q, 3] (AVA) or [B, 2] (lb_sb), sizes [B, 2] = (H, W);
// out [B*Q, 4 + C + 1] = xyxy scaled | scores | foreground probability
cudaError_t launch_postprocess(const float* logits, long long l_sb, const float* boxes, long long b_sb, const float* logits_b,
                               long long lb_sb, const float* sizes, float* out, int B, int Q, int C, int ava, cudaStream_t st);

// ---- uint8 frames -> normalised fp32 clip (video_transforms.py:294-296,308-314; ava_frame.py:71-72) ------------
// frames [B][pixels_per_clip = T*H*W][3] uint8 (RGB, HWC), lut [3][256] fp32 (see tuber_set_input_norm), out [B][3][pixels_per_clip] fp32
cudaError_t launch_normalize_u8(const uint8_t* frames, const float* lut, float* out, int B, long long pixels_per_clip, cudaStream_t st);

// ---- masks and position code (backbone_builder.py:85-86, position_encoding.py:32-72) ------
cudaError_t launch_mask_resize(const uint8_t* mask, uint8_t* fmask, int B, int H, int W, int T, int Hf,
                               int Wf, cudaStream_t st);
cudaError_t launch_posenc(const uint8_t* fmask, const float* dim_t, const float* dim_s, float* pos, int B,
                          int T, int H, int W, int nt, int ns, cudaStream_t st);

// ---- format conversion (tests, plumbing) --------------------------------------------------
cudaError_t launch_to_split(const float* in, int ldi, void* out, int ldo, long long rows, int cols,
                            cudaStream_t st);
cudaError_t launch_from_split(const void* in, int ldi, float* out, int ldo, long long rows, int cols,
                              cudaStream_t st);

// ---- tcgen05 bf16x3 GEMM (gemm_tc.cu) -------------------------------------------------------
cudaError_t launch_gemm_tc(const GemmArgs& a, cudaStream_t st);
const char* gemm_tc_last_error();
// name of the template configuration launch_gemm_tc will use: gemm_bf16x3_wide (<128,2,3>, memory-bound shapes), _deep (<128,3,1>,
// K >= 512), _n64 (<64,3,2>, N = 64 or token-sized)
const char* gemm_tc_config_name(const GemmArgs& a);
// can the tcgen05 kernels read Ab through the strided tensor map for this geometry (128-row tiles must be boxes of the output grid)?
bool gemm_tc_strided_ab_ok(int M, int Wo, int Ho, int Wi, int Hi, int BTi, int st_t, int st_s);
// conv4 (+ shortcut / residual) of one bottleneck fused with conv1 of the next (256-channel stage): a describes the first
// GEMM (N = 256, split output), the second is C2[M, N2] = relu(scale2 * (C W2^T) + shift2) in fp32, N2 in {64, 128}
cudaError_t launch_gemm_tc_fused2(const GemmArgs& a, const void* W2p, const float* scale2, const float* shift2, float* C2, int N2,
                                  int ldc2, cudaStream_t st);
// fp32 [N,K] -> packed split bf16 [2][N][K]
cudaError_t launch_pack_weight(const float* w, void* out, int N, int K, cudaStream_t st);
