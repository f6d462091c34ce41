/* tuber_b200 -- C ABI of the B200-native TubeR forward path (libtuber_b200.so).
 *
 * The reference (amazon-science/tubelet-transformer) is pure Python and has no FFI; its
 * "plugin API" for this path is
 *     models/tuber_ava.py:160   build_model(cfg) -> (model, criterion, postprocessors)
 *     models/tuber_ava.py:97    DETR.forward(samples: NestedTensor) -> dict
 * called from utils/video_action_recognition.py:303 (`outputs = model(samples)`).
 * The functions below are what a binding for that path needs, and each one names the reference
 * interface it replaces.  Conventions:
 *   - plain C, no torch types; every function returns 0 on success or a negative TuberStatus and
 *     never throws; tuber_last_error() gives the message of the calling thread's last failure.
 *   - "dev" pointers are device pointers owned by the caller; "host" pointers are host memory
 *     (pinned for the *_host entry points if the copies are to overlap).
 *   - all device work is stream-ordered on the cudaStream_t passed in (as void*); no hidden
 *     synchronisation except where stated.
 *   - a plan is bound to the CUDA device current at tuber_plan_create(); one plan per
 *     (device, config); calls on one plan must not overlap, different plans are independent.
 */
#ifndef TUBER_B200_H_
#define TUBER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TUBER_ABI_VERSION 1

typedef enum TuberStatus {
  TUBER_OK = 0,
  TUBER_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
  TUBER_ERR_MISSING = -2,     /* a required weight tensor was never set   */
  TUBER_ERR_SHAPE = -3,       /* weight or input shape mismatch           */
  TUBER_ERR_CUDA = -4,        /* a CUDA call failed                        */
  TUBER_ERR_STATE = -5        /* call order violated (e.g. forward before finalize) */
} TuberStatus;

/* Temporal down-sampling of the backbone feature map (models/backbone_builder.py:41-53,70-80). */
typedef enum TuberPool { TUBER_POOL_AVG = 0, TUBER_POOL_MAX = 1, TUBER_POOL_DECODE = 2, TUBER_POOL_CENTER = 3,
                         TUBER_POOL_NONE = 4 /* SINGLE_FRAME: False */ } TuberPool;

/* Model hyper-parameters: the cfg.CONFIG.* keys build_model reads (tuber_ava.py:160-184,
 * backbone_builder.py:109-113, transformer.py:303-314). */
typedef struct TuberConfig {
  int32_t abi_version;        /* = TUBER_ABI_VERSION */
  int32_t blocks[4];          /* bottlenecks per stage: {3,8,36,3} CSN-152, {3,4,6,3} CSN-50 */
  int32_t last_stride;        /* MODEL.LAST_STRIDE: spatial stride 2 in layer4 when non-zero */
  int32_t pool;               /* TuberPool */
  int32_t pool_kernel;        /* MODEL.TEMP_LEN / MODEL.DS_RATE (avg / max window)            */
  int32_t d_model, nhead, enc_layers, dec_layers, dim_ff;
  int32_t num_queries;        /* rows of query_embed (QUERY_NUM, or QUERY_NUM*TEMP_LEN off-AVA) */
  int32_t num_classes;        /* columns of class_fc                                          */
  int32_t ava_mode;           /* 1: class_embed_b = Linear(d,3) on hs; 0: Linear(2048,2) on pooled xt */
} TuberConfig;

typedef struct TuberPlan TuberPlan;

/* ---- life cycle (replaces build_model(cfg), tuber_ava.py:160) --------------------------- */
int tuber_plan_create(const TuberConfig* cfg, TuberPlan** out_plan);
void tuber_plan_destroy(TuberPlan* plan);

/* Hand one state_dict entry to the plan (replaces nn.Module.load_state_dict /
 * utils/model_utils.py:66-95 for this path).  `name` is the reference's parameter or buffer name
 * ("backbone.body.layer2.0.conv1.weight", "transformer.encoder.layers.0.self_attn.in_proj_weight",
 * ...); `data` is host fp32, row-major in the reference's shape; it is
 * copied before the call returns.  Unknown names are ignored (returns TUBER_OK) so a whole
 * checkpoint can be streamed through. */
int tuber_plan_set_weight(TuberPlan* plan, const char* name, const float* host_data, const int64_t* shape,
                          int32_t ndim);

/* Fold BatchNorm (eps 1e-3) into per-channel scale/shift, split weights to bf16 hi/mid planes,
 * pack everything into the device arena, pre-compute the input-independent part of the decode
 * pool.  Synchronises the device.  Fails with TUBER_ERR_MISSING naming the first absent tensor. */
int tuber_plan_finalize(TuberPlan* plan);

/* ---- forward (replaces DETR.forward, tuber_ava.py:97-148) ------------------------------- */
/* clips_dev : fp32 (B,3,T,H,W) NCDHW, ImageNet-normalised (NestedTensor.tensors)
 * mask_dev  : uint8 (B,H,W), 1 = padding (NestedTensor.mask); NULL = no padding anywhere
 * outputs, for ALL decoder layers l (the reference returns l = L-1 plus 'aux_outputs' for l < L-1):
 *   logits_dev   fp32 (B, L, Q, num_classes)
 *   boxes_dev    fp32 (B, L, Q, 4)   cx,cy,w,h in (0,1)
 *   logits_b_dev fp32 (B, L, Q, 3) in ava_mode, else (B, 2) (identical for every layer)
 * Asynchronous on `stream`.  Grows the plan's workspace on first use of a larger shape (that
 * call synchronises). */
int tuber_forward(TuberPlan* plan, const float* clips_dev, const uint8_t* mask_dev, int32_t B, int32_t T, int32_t H,
                  int32_t W, float* logits_dev, float* boxes_dev, float* logits_b_dev, void* stream);

/* Same, from/to host memory: H2D copy of clips (+mask), forward, D2H copy of the three outputs,
 * then waits for the stream.  This is the call timed as `e2e` by bench.py. */
int tuber_forward_host(TuberPlan* plan, const float* clips_host, const uint8_t* mask_host, int32_t B, int32_t T,
                       int32_t H, int32_t W, float* logits_host, float* boxes_host, float* logits_b_host,
                       void* stream);

/* Pipelined form of tuber_forward_host for a stream of batches: submit() enqueues the H2D copies of `slot`
 * (0 or 1) on a private copy stream and the forward + D2H copies on a private compute stream and returns at once;
 * wait() blocks until that slot's outputs are in host memory.  Submitting slot 1 while slot 0 computes overlaps
 * the next batch's input copy with the current batch's kernels (host buffers must be pinned for the overlap and
 * must stay valid until wait()).  A slot must be waited on before it is submitted again. */
int tuber_forward_host_submit(TuberPlan* plan, int32_t slot, const float* clips_host, const uint8_t* mask_host, int32_t B,
                              int32_t T, int32_t H, int32_t W, float* logits_host, float* boxes_host, float* logits_b_host);
int tuber_forward_host_wait(TuberPlan* plan, int32_t slot);

/* ---- uint8 frames in (replaces the tail of the reference's input pipeline) ----------------------
 * The reference decodes JPEG frames to uint8 RGB and then runs, on the host, ToTensor (uint8 HWC -> float32 CHW, / 255) and
 * Normalize ((x - mean) / std) per frame, stacks and permutes to (3,T,H,W) (datasets/video_transforms.py:294-296,308-314;
 * datasets/ava_frame.py:71-72,158-162) -- 12 bytes per pixel then cross the bus.  These entry points take the decoded frames
 * themselves, uint8 (B,T,H,W,3) RGB, and apply that transform on the device (3 bytes per pixel cross the bus); the values fed
 * to the stem are bit-identical to the reference transform's (a 3 x 256 value table evaluated in fp32 the way torch does).
 * tuber_set_input_norm sets mean / std (3 floats each; default = the reference's ImageNet constants, ava_frame.py:159-162);
 * it synchronises the device.  tuber_input_lut is the host-only table builder (no device needed), lut_out = float[3*256].
 * tuber_forward_host_u8_submit shares its slots and tuber_forward_host_wait with tuber_forward_host_submit. */
int tuber_input_lut(const float* mean, const float* std, float* lut_out);
int tuber_set_input_norm(TuberPlan* plan, const float* mean, const float* std);
int tuber_forward_u8(TuberPlan* plan, const uint8_t* frames_dev, const uint8_t* mask_dev, int32_t B, int32_t T, int32_t H,
                     int32_t W, float* logits_dev, float* boxes_dev, float* logits_b_dev, void* stream);
int tuber_forward_host_u8(TuberPlan* plan, const uint8_t* frames_host, const uint8_t* mask_host, int32_t B, int32_t T,
                          int32_t H, int32_t W, float* logits_host, float* boxes_host, float* logits_b_host, void* stream);
int tuber_forward_host_u8_submit(TuberPlan* plan, int32_t slot, const uint8_t* frames_host, const uint8_t* mask_host, int32_t B,
                                 int32_t T, int32_t H, int32_t W, float* logits_host, float* boxes_host, float* logits_b_host);

/* ---- long-term context bank (BASELINE.json configs[3]) ------------------------------------------
 * NOT IN THE REFERENCE: its README (README.md:16-18,86) announces the long-term context variant but the code was never
 * released, so this layer is defined here, after the paper's description, and has no reference parity (its oracle is
 * oracle/tuber_oracle.py::forward(bank=...)).  Definition:
 *   bank entry of a clip   e = mean over the T' feature frames of class_proj(xt)             -> (H'W' tokens, d) fp32
 *   context layer          mem_c <- LayerNorm(mem_c + MHA_8heads(q = mem_c, k = v = bank))   (weights ltc_attn.*, ltc_norm.*)
 * applied to the class-branch tokens between the class encoder (tuber_ava.py:133-135) and the class cross-attention (:137-139).
 * bank_dev: fp32 (bank_clips, bank_tokens, d), bank_clips = 1 (one window shared by the batch) or B (one window per clip),
 * typically the entries of the 64 clips around the current one (16 384 tokens); NULL = no context layer (fill pass).
 * bank_new_dev: fp32 (B, H'W', d), receives this batch's entries; NULL = not wanted.  Everything else as tuber_forward.
 * The layer exists when the state_dict streamed through tuber_plan_set_weight held ltc_attn.in_proj_weight / in_proj_bias /
 * out_proj.weight / out_proj.bias and ltc_norm.weight / bias (tuber_has_ltc). */
int tuber_has_ltc(TuberPlan* plan);
int tuber_forward_ltc(TuberPlan* plan, const float* clips_dev, const uint8_t* mask_dev, int32_t B, int32_t T, int32_t H, int32_t W,
                      const float* bank_dev, int32_t bank_clips, int32_t bank_tokens, float* bank_new_dev, float* logits_dev,
                      float* boxes_dev, float* logits_b_dev, void* stream);

/* Output geometry for an input shape: feature-map size after the backbone (T',H',W'), tokens seen
 * by the DETR encoder, number of kernel launches one forward issues. */
typedef struct TuberShapeInfo {
  int32_t Tf, Hf, Wf, Tp, enc_tokens, cls_tokens, launches;
  int64_t workspace_bytes;
} TuberShapeInfo;
int tuber_query_shapes(TuberPlan* plan, int32_t B, int32_t T, int32_t H, int32_t W, TuberShapeInfo* out);

/* ---- execution options -------------------------------------------------------------------- */
/* Replay the launch sequence of a forward as one CUDA graph: the first call for a given
 * (shape, buffer pointers) runs eagerly and records, later calls with the same key are a single
 * cudaGraphLaunch.  Ignored while profiling or debug-keep is on. */
int tuber_set_graph(TuberPlan* plan, int32_t enabled);
/* Debugging cross-check: route every GEMM through the fp32 CUDA-core kernel instead of the tcgen05
 * one (also selectable with the environment variable TUBER_FORCE_SIMT=1 at plan creation). */
int tuber_set_force_simt(TuberPlan* plan, int32_t enabled);
/* Keep a private copy of every named intermediate of the next forwards (tuber_debug_fetch can then
 * also return "stem", "layer1".."layer4", which live in recycled buffers). */
int tuber_set_debug_keep(TuberPlan* plan, int32_t enabled);
/* Kernel launches (+ memsets / copies) issued by the last eager forward. */
int tuber_last_launches(TuberPlan* plan);
/* Recorded CUDA graphs the plan currently holds (one per (shape, buffer addresses); tuber_set_graph). */
int tuber_graph_count(TuberPlan* plan);

/* ---- instrumentation ---------------------------------------------------------------------- */
/* When enabled, tuber_forward brackets each stage with CUDA events (adds event records only). */
#define TUBER_NUM_STAGES 10
int tuber_set_profiling(TuberPlan* plan, int32_t enabled);
/* ms of the last profiled forward per stage: stem, layer1..4, pool, proj, encoder, decoder,
 * class-branch+heads.  Synchronises on the last event. */
int tuber_get_stage_ms(TuberPlan* plan, float* ms_out /* [TUBER_NUM_STAGES] */);
/* ALGORITHMIC bytes / flops of the last forward per stage (same stages, same accounting as tuber_get_kernel_profile: each operand
 * read once, each result written once, flops = 2 x MACs; SURVEY 8d's per-clip figures x the clips of the call) -- with
 * tuber_get_stage_ms the per-stage roofline fractions bench.py reports (BASELINE north_star: "each stage reported as achieved
 * fraction of its HBM-or-tensor-core roofline").  No reference counterpart (the reference has no instrumentation). */
int tuber_get_stage_work(TuberPlan* plan, double* bytes_out /* [TUBER_NUM_STAGES] */, double* flops_out /* [TUBER_NUM_STAGES] */);
const char* tuber_stage_name(int32_t i);
/* Per-kernel profile: when enabled, every launch of a forward is bracketed by its own pair of CUDA
 * events on the launching stream (disables graph replay).  tuber_get_kernel_profile aggregates the last
 * forward by kernel: launches, summed device ms, and the summed ALGORITHMIC bytes / flops of those
 * launches (each operand read once, each result written once; flops = 2 x MACs) -- the numerators of
 * the roofline bench.py reports.  Pass out = NULL to query the number of rows. */
typedef struct TuberKernelStat {
  char name[32];
  int32_t launches;
  float ms;
  double bytes, flops;
} TuberKernelStat;
int tuber_set_kernel_profiling(TuberPlan* plan, int32_t enabled);
int tuber_get_kernel_profile(TuberPlan* plan, TuberKernelStat* out, int32_t capacity, int32_t* n_out);
/* Copy an intermediate of the last forward to fp32 device memory, for tests:
 * "xt" (B,T',H',W',2048) channels-last, "xs" (B,tokens,2048), "src" / "memory" (B,tokens,d),
 * "hs" (B,L,Q,d), "mem_c" (B,T'H'W',d), "mem_ltc" (same, after the context layer), "pos" (B or 1,tokens,d); with debug-keep also "stem",
 * "layer1".."layer4" (B,T,H,W,C) channels-last.  Returns element count through *n_out; dst_dev may be NULL to query. */
int tuber_debug_fetch(TuberPlan* plan, const char* what, float* dst_dev, int64_t* n_out, void* stream);

/* Post-processing of one decoder layer's outputs fused with the packing of the detection rows the reference's evaluation loop
 * writes (models/criterion.py:413-482 PostProcess / PostProcessAVA; utils/video_action_recognition.py:311-346,411-415):
 * out_dev [B*Q, 4 + C + 1] = boxes (x1,y1,x2,y2) scaled to sizes_dev [B,2] = (H,W) | class scores | foreground probability.
 * logits_dev / boxes_dev / logits_b_dev are the (B,L,Q,.) tensors of tuber_forward ((B,2) actor-ness logits outside AVA);
 * `layer` selects the decoder layer (L-1 = the model's prediction). */
int tuber_postprocess(TuberPlan* plan, const float* logits_dev, const float* boxes_dev, const float* logits_b_dev,
                      const float* sizes_dev, int32_t B, int32_t layer, float* out_dev, void* stream);

/* ---- single operators (unit tests / microbenchmarks; same kernels the plan launches) ------- */
int tuber_op_to_split(const float* in_dev, void* out_dev, int64_t rows, int32_t cols, void* stream);
int tuber_op_from_split(const void* in_dev, float* out_dev, int64_t rows, int32_t cols, void* stream);
/* tcgen05 bf16x3 GEMM: C = act(scale*(A W^T) + shift + res); A split [M,K], W fp32 host-packed by
 * tuber_op_pack_weight (split planes [2][N][K] bf16), C/res fp32 (fmt 0) or split (fmt 1). */
int tuber_op_pack_weight(const float* w_dev /* [N,K] fp32 */, void* out_dev, int32_t N, int32_t K, void* stream);
int tuber_op_gemm_tc(const void* a_split_dev, const void* w_packed_dev, const float* scale_dev,
                     const float* shift_dev, const void* res_dev, int32_t res_fmt, int32_t res_mod, void* c_dev,
                     int32_t c_fmt, int32_t M, int32_t N, int32_t K, int32_t relu, void* stream);
/* Two chained pointwise convolutions in one kernel (reference ir_CSN_152.py:84-90 of block i then :73-75 of block i+1):
 * C[M,N1] = relu(scale*([A | Ab] W^T) + shift + res) (split),  C2[M,N2] = relu(scale2*(C W2^T) + shift2) (fp32).
 * A split [M,K], Ab split [M,Kb] or NULL, W packed [2][N1][K+Kb], res split [M,N1] or NULL, W2 packed [2][N2][N1];
 * (N1, N2) in {(256, 64), (256, 128), (512, 128), (512, 256), (1024, 256)}. */
int tuber_op_gemm_tc_fused2(const void* a_split_dev, const void* ab_split_dev, const void* w_packed_dev, const float* scale_dev,
                            const float* shift_dev, const void* res_split_dev, void* c_split_dev, int32_t M, int32_t K, int32_t Kb,
                            const void* w2_packed_dev, const float* scale2_dev, const float* shift2_dev, float* c2_dev, int32_t N1,
                            int32_t N2, void* stream);
/* fp32 CUDA-core GEMM: C = act(A W^T + bias + res), act: 0 none, 1 relu, 2 sigmoid */
int tuber_op_sgemm(const float* a_dev, const float* w_dev, const float* bias_dev, const float* res_dev, float* c_dev,
                   int32_t M, int32_t N, int32_t K, int32_t act, void* stream);
/* depthwise 3x3x3 + scale/shift + ReLU: in fp32 [B,Ti,Hi,Wi,C], w [27][C], out split */
int tuber_op_dwconv(const float* in_dev, const float* w27c_dev, const float* scale_dev, const float* shift_dev,
                    void* out_split_dev, int32_t B, int32_t Ti, int32_t Hi, int32_t Wi, int32_t C, int32_t stride_t,
                    int32_t stride_s, void* stream);
/* stem conv + scale/shift + ReLU (fp32 NDHWC out) and the (1,3,3) max pool (split out); w = the reference's
 * conv1.weight (64,3,3,7,7) fp32 on the device.  Synchronises the stream. */
int tuber_op_stem(const float* x_ncdhw_dev, const float* w_oc441_dev, const float* scale_dev, const float* shift_dev,
                  float* conv_out_dev, void* pooled_split_dev, int32_t B, int32_t T, int32_t H, int32_t W, void* stream);
/* LayerNorm(x + res) over C in {256, 2048} */
int tuber_op_layernorm(const float* x_dev, const float* res_dev, const float* gamma_dev, const float* beta_dev,
                       float* out_dev, int64_t rows, int32_t C, void* stream);
/* softmax(scale QK^T + mask)V for contiguous (NB, L|S, H*D) fp32 tensors */
int tuber_op_attention(const float* q_dev, const float* k_dev, const float* v_dev, const uint8_t* kpm_dev,
                       float* out_dev, int32_t NB, int32_t H, int32_t L, int32_t S, int32_t D, float scale,
                       void* stream);
/* Name of the kernel tuber_op_attention / the plan dispatch a contiguous (NB, L|S, H*D) problem to: "attn_tc_kernel" (tcgen05: head
 * dim 32, L >= 64, S >= 128, with or without a key padding mask), "attn_mma_kernel" / "attn_mma_split_kernel" / "attn_tiny_kernel"
 * (mma.sync / registers: the 15-query decoder attentions, T' <= 8 temporal attention), "attn_simt_kernel" (other head dims). */
const char* tuber_op_attention_kernel(int32_t NB, int32_t H, int32_t L, int32_t S, int32_t D, int32_t masked);
/* uint8 frames (B, pixels_per_clip, 3) -> fp32 (B, 3, pixels_per_clip) = ((u / 255) - mean[c]) / std[c].  Synchronises the stream. */
int tuber_op_normalize_u8(const uint8_t* frames_dev, const float* mean, const float* std, float* out_dev, int32_t B,
                          int64_t pixels_per_clip, void* stream);
/* 3-D sine position code from a feature-resolution mask (B,T,H,W) -> (B, T*H*W, d_model) */
int tuber_op_posenc(const uint8_t* fmask_dev, float* pos_dev, int32_t B, int32_t T, int32_t H, int32_t W,
                    int32_t d_model, void* stream);

/* ---- frame loading (widened row f4 of the scope contract) --------------------------------------------------------------------
 * Replaces, per frame, the reference loader's  Image.open(path)  +  .resize((w, h))  (datasets/ava_frame.py:146-150; Pillow's
 * default filter for RGB images: BICUBIC) -- bit-identical to Pillow on libjpeg(-turbo): Huffman decoding on host threads,
 * dequantisation + "islow" inverse DCT, "fancy" chroma upsampling, YCbCr -> RGB and Pillow's two-pass 8-bit resample on the GPU.
 * Baseline sequential YCbCr JPEGs (4:4:4 / 4:2:2 / 4:2:0, restart intervals); anything else is refused (TUBER_ERR_INVALID). */
typedef struct TuberFrameDecoder TuberFrameDecoder;
int tuber_frames_create(TuberFrameDecoder** out_decoder, int32_t max_host_threads /* <= 0: all */);
void tuber_frames_destroy(TuberFrameDecoder* decoder);
/* n JPEG byte ranges in host memory -> RGB uint8 [n, out_h, out_w, 3] on the device (the input layout of tuber_forward_u8).
 * Entropy decoding runs before the call returns; the copies and kernels are ordered on `stream` (use ONE stream per decoder: the
 * device buffers are reused from call to call; the pinned staging is double-buffered, so the host work of call i+1 overlaps the
 * device work of call i).  One call at a time per decoder. */
int tuber_frames_decode(TuberFrameDecoder* decoder, const uint8_t* const* jpeg_ptrs_host, const int64_t* jpeg_sizes, int32_t n,
                        int32_t out_h, int32_t out_w, uint8_t* frames_dev, void* stream);
const char* tuber_frames_last_error(void);
/* host only, no device needed: the entropy decoder by itself.  info_out[10] = {W, H, luma h, luma v, blocks_x / blocks_y of Y, Cb, Cr};
 * coef_out (may be NULL to query the sizes) receives the quantised coefficients, int16, natural order, [Y | Cb | Cr] blocks of 64. */
int tuber_op_jpeg_coefficients(const uint8_t* jpeg_host, int64_t size, int16_t* coef_out, int64_t capacity, int32_t* info_out);

const char* tuber_last_error(void);
int tuber_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TUBER_B200_H_ */
