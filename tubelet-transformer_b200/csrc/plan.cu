// C-ABI of the B200-native TubeR forward path (include/tuber_b200.h): weight ingest and packing,
// workspace management and the launch sequence of one forward.
//
// The launch sequence follows DETR.forward (reference models/tuber_ava.py:97-148):
//   backbone body      models/backbones/ir_CSN_152.py:172-186 (stem :176-179, bottlenecks :70-90)
//   temporal pooling   models/backbone_builder.py:70-80 (decode pool: transformer_layers.py:380-448)
//   mask + position    models/backbone_builder.py:85-89, models/transformer/position_encoding.py:32-72
//   encoder / decoder  models/transformer/transformer.py:49-64,153-168,218-249
//   class branch       models/transformer/transformer_layers.py:71-97, models/tuber_ava.py:129-141
//   heads              models/tuber_ava.py:121-125,141-142, models/criterion.py:485-497
// Activations are channels-last; "S" = split-bf16 (operand of the tcgen05 GEMM), "F" = fp32 (common.cuh).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/tuber_b200.h"
#include "kernels.h"

bool tuber_pdl_enabled() {
  // measured on B200 inside the captured graph: no gain without an early trigger (1194 vs 1193 clips/s), 1.5 % slower with
  // griddepcontrol.launch_dependents at kernel entry -> off unless TUBER_PDL=1
  static const bool on = [] { const char* e = getenv("TUBER_PDL"); return e && e[0] == '1'; }();
  return on;
}

namespace {

thread_local char g_last_error[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof g_last_error, fmt, ap);
  va_end(ap);
  return code;
}

#define CK(expr)                                                                                  \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return fail(TUBER_ERR_CUDA, "%s:%d %s -> %s %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e), \
                  gemm_tc_last_error());                                                          \
  } while (0)
#define TRY(expr)              \
  do {                         \
    int _s = (expr);           \
    if (_s != TUBER_OK) return _s; \
  } while (0)

constexpr float BN_EPS = 1e-3f;   // ir_CSN_152.py:15
constexpr float LN_EPS = 1e-5f;   // torch.nn.LayerNorm default
constexpr int POOL_DIM = 2048;    // backbone_builder.py:49-52
constexpr int CLS_FF = 2048;      // tuber_ava.py:60
constexpr int CLS_HEADS = 8;      // tuber_ava.py:60,62

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

struct Lin {          // y = scale * (x W^T) + shift
  float* wf = nullptr; void* wp = nullptr; float* scale = nullptr; float* shift = nullptr;
  int N = 0, K = 0;
};
struct Dw { float* w = nullptr; float* scale = nullptr; float* shift = nullptr; int C = 0; };
struct LnP { float* g = nullptr; float* b = nullptr; int C = 0; };

struct Block {
  Lin conv1, conv4;
  Lin c4ds;                 // first block of a stage: [s4*W4 | sd*Wd] with shift b4 + bd (conv4 and the shortcut as one GEMM)
  Dw dw;
  bool has_ds = false;
  int st_t = 1, st_s = 1, cin = 0, planes = 0, cout = 0;
};
struct EncLayer { Lin in, out, lin1, lin2; LnP n1, n2; };
struct DecLayer {
  Lin sa_in, sa_out, ca_q, ca_out, lin1, lin2;
  LnP n1, n2, n3;
  float* pq_sa = nullptr;   // [Q, 3d]: query_embed x [Wq;Wk;0]^T   (q = k = tgt + query_pos, transformer.py:225)
  float* pq_ca = nullptr;   // [Q, d] : query_embed x Wq^T           (query = tgt + query_pos, :232)
  float* sa_out_t = nullptr; float* ca_q_t = nullptr; float* ca_out_t = nullptr;   // K-major fp32 copies [d, d] (decoder_mega.cu)
};

constexpr int SIDE_CTAS_DEFAULT = 0;   // SM cap of the class-branch side stream (0 = none); TUBER_SIDE_CTAS overrides
struct KernelRec { const char* name; double bytes, flops; cudaEvent_t e0, e1; char tag[64]; };
struct Tap { const void* ptr; int fmt; long long rows; int cols; void* keep; };

struct Ws {   // bump allocator over one device buffer; in dry mode only counts
  char* base = nullptr; size_t cap = 0, off = 0, peak = 0; bool dry = false, overflow = false;
  void* alloc(size_t bytes) {
    size_t o = (off + 255) & ~(size_t)255;
    off = o + bytes;
    if (off > peak) peak = off;
    if (dry) return reinterpret_cast<void*>((uintptr_t)0x1000 + o);
    if (off > cap) { overflow = true; return base; }      // never hand out memory past the buffer
    return base + o;
  }
};

}  // namespace

struct TuberPlan {
  TuberConfig cfg;
  int device = 0;
  bool finalized = false;
  std::unordered_map<std::string, HostTensor> host;
  std::vector<void*> owned;          // device allocations of packed weights

  // packed model
  void* stem_w = nullptr; float* stem_scale = nullptr; float* stem_shift = nullptr;
  std::vector<Block> blocks[4];
  // decode pool (input-independent parts folded at finalize)
  float* pool_tgt1 = nullptr;        // [2048] LN1(query_pool + self_attn(query_pool))
  float* pool_q = nullptr;           // [2048] Wq tgt1 + bq of the cross attention
  Lin pool_kv, pool_v, pool_out, pool_lin1, pool_lin2;
  float* pool_u = nullptr;           // [8][2048]: scale * Wk_h^T q_h, the key projection folded into the pooled query
  LnP pool_n2, pool_n3, pool_nf;
  Lin input_proj, class_proj;
  std::vector<EncLayer> enc;
  std::vector<DecLayer> dec;
  LnP dec_norm;
  float* dec0_c1 = nullptr;          // [d]: decoder layer 0 after its (input independent) self-attention block: LN1(Wo bv + bo)
  float* dec0_qc = nullptr;          // [Q, d]: its cross-attention query Wq (dec0_c1 + query_embed) + bq
  Lin pos_proj;                      // N = Le*3d + Ld*2d: per encoder layer [Wq;Wk;0], per decoder layer [Wk_cross;0]
  Lin mem_kv;                        // N = Ld*2d: per decoder layer [Wk_cross;Wv_cross]
  float* dim_t = nullptr; float* dim_s = nullptr;
  // class branch
  Lin ct_in, ct_out, cs_in, cs_out, c_lin1, c_lin2, x_q, x_kv, x_out;
  LnP c_n1t, c_n1s, c_n2;
  Lin head_b, bbox0, bbox1, bbox2, class_fc;
  // long-term context layer (defined in this repo, the reference never released it): packed when the
  // state_dict holds ltc_attn.* / ltc_norm.*
  bool has_ltc = false;
  Lin ltc_q, ltc_kv, ltc_out;
  LnP ltc_n;
  // arguments of the tuber_forward_ltc call in progress (read by run_forward and by the graph key)
  const float* ltc_bank = nullptr; int ltc_bank_clips = 0, ltc_bank_tokens = 0;
  float* ltc_new = nullptr;

  // run-time state
  char* ws = nullptr; size_t ws_cap = 0;
  // forward_host staging: two slots so that the host copies of step i+1 overlap the kernels of step i
  char* stage_in[2] = {nullptr, nullptr}; size_t stage_in_cap[2] = {0, 0};
  char* stage_out[2] = {nullptr, nullptr}; size_t stage_out_cap[2] = {0, 0};
  cudaStream_t copy_stream = nullptr, run_stream = nullptr, out_stream = nullptr;
  cudaEvent_t h2d_done[2] = {nullptr, nullptr}, slot_done[2] = {nullptr, nullptr}, fwd_done[2] = {nullptr, nullptr};
  bool slot_busy[2] = {false, false};
  // uint8 input path (tuber_forward_u8*): value table of the reference's ToTensor + Normalize and the fp32 clip it expands into
  float* in_lut = nullptr;           // [3][256] on the device
  const uint8_t* in_frames = nullptr; // argument of the uint8 call in progress: the stem reads the frames itself (read by run_forward and the graph key)
  float* u8_clip = nullptr; size_t u8_clip_cap = 0;
  bool no_dec_mega = false, no_fuse2_s23 = false;
  unsigned long long* dec_trace = nullptr; int dec_trace_n = 0;   // per-phase timestamps of the decoder kernel (kernel profiling only)
  bool force_simt = false, no_fuse2 = false, fuse2_deep = false, pool_unfolded = false, no_strided_tma = false;
  bool profiling = false, debug_keep = false, use_graph = false;
  cudaEvent_t ev[TUBER_NUM_STAGES + 1] = {};
  bool ev_valid = false;
  double stage_bytes[TUBER_NUM_STAGES] = {}, stage_flops[TUBER_NUM_STAGES] = {};   // algorithmic work of the last forward, per stage
  std::map<std::string, Tap> taps;
  int launches = 0;
  bool kprof = false;
  cudaStream_t cap_stream = nullptr;
  // side branches of a forward (Ctx::side_begin / side_end / side_join): the position-code chain (needs only the mask) and the
  // class-branch encoder (needs only class_proj's output) run on their own streams beside the backbone / the DETR encoder + decoder,
  // whose token-sized launches leave most SMs idle; recorded into the CUDA graph as parallel branches
  static constexpr int NUM_SIDE = 2;
  cudaStream_t side[NUM_SIDE] = {}; cudaEvent_t ev_fork[NUM_SIDE] = {}, ev_join[NUM_SIDE] = {};
  bool side_valid = false, no_overlap = false;
  int side_ctas = 0;                 // > 0: CTAs (SMs) a persistent kernel of the class-branch side stream may occupy
  std::vector<KernelRec> kp; int kp_used = 0;
  struct GraphEntry { std::vector<uintptr_t> key; cudaGraphExec_t exec; };
  std::vector<GraphEntry> graphs;
};

namespace {

// ------------------------------------------------------------------------------------------
// weight ingest helpers
// ------------------------------------------------------------------------------------------
struct Packer {
  TuberPlan* p;
  std::string missing;
  int status = TUBER_OK;

  const HostTensor* get(const std::string& name, std::initializer_list<int64_t> shape) {
    auto it = p->host.find(name);
    if (it == p->host.end()) {
      if (status == TUBER_OK) { status = fail(TUBER_ERR_MISSING, "weight '%s' was never set", name.c_str()); }
      return nullptr;
    }
    if (shape.size()) {
      int64_t want = 1, have = 1;
      for (auto s : shape) want *= s;
      for (auto s : it->second.shape) have *= s;
      if (want != have) {
        if (status == TUBER_OK) status = fail(TUBER_ERR_SHAPE, "weight '%s' has %lld elements, expected %lld", name.c_str(), (long long)have, (long long)want);
        return nullptr;
      }
      // the expected shape is the logical (rows, cols) view: the tensor's own dimensions must flatten onto it in order
      // ((64,3,3,7,7) -> {64,441}, (256,2048,1,1,1) -> {256,2048}); a transposed or re-laid-out tensor of the same size does not
      auto dim = it->second.shape.begin();
      const auto end = it->second.shape.end();
      bool ok = true;
      for (auto s : shape) {
        int64_t acc = 1;
        while (acc < s && dim != end) acc *= *dim++;
        if (acc != s) { ok = false; break; }
      }
      for (; ok && dim != end; ++dim) ok = (*dim == 1);
      if (!ok) {
        std::string have_s, want_s;
        for (auto s : it->second.shape) have_s += (have_s.empty() ? "" : ",") + std::to_string(s);
        for (auto s : shape) want_s += (want_s.empty() ? "" : ",") + std::to_string(s);
        if (status == TUBER_OK) status = fail(TUBER_ERR_SHAPE, "weight '%s' has shape (%s), which does not flatten to (%s)", name.c_str(), have_s.c_str(), want_s.c_str());
        return nullptr;
      }
    }
    return &it->second;
  }

  template <typename T>
  T* upload(const std::vector<T>& v) {
    void* d = nullptr;
    if (cudaMalloc(&d, v.size() * sizeof(T) + 16) != cudaSuccess) {
      if (status == TUBER_OK) status = fail(TUBER_ERR_CUDA, "cudaMalloc of %zu bytes failed", v.size() * sizeof(T));
      return nullptr;
    }
    p->owned.push_back(d);
    if (cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {
      if (status == TUBER_OK) status = fail(TUBER_ERR_CUDA, "host -> device copy of %zu bytes failed", v.size() * sizeof(T));
      return nullptr;
    }
    return reinterpret_cast<T*>(d);
  }

  // fp32 [N,K] host rows -> Lin (both the fp32 copy and the packed split planes)
  Lin make_lin(const std::vector<float>& w, int N, int K, const std::vector<float>* scale, const std::vector<float>* shift) {
    Lin l;
    l.N = N; l.K = K;
    l.wf = upload(w);
    void* wp = nullptr;
    if (cudaMalloc(&wp, (size_t)N * K * 4 + 16) != cudaSuccess) {
      if (status == TUBER_OK) status = fail(TUBER_ERR_CUDA, "cudaMalloc failed");
      return l;
    }
    p->owned.push_back(wp);
    l.wp = wp;
    if (l.wf) launch_pack_weight(l.wf, wp, N, K, 0);
    if (scale) l.scale = upload(*scale);
    if (shift) l.shift = upload(*shift);
    return l;
  }

  // eval-mode BatchNorm3d folded to y = x*scale + shift
  bool bn(const std::string& prefix, int C, std::vector<float>& scale, std::vector<float>& shift) {
    const HostTensor* g = get(prefix + ".weight", {C});
    const HostTensor* b = get(prefix + ".bias", {C});
    const HostTensor* m = get(prefix + ".running_mean", {C});
    const HostTensor* v = get(prefix + ".running_var", {C});
    if (!g || !b || !m || !v) return false;
    scale.resize(C); shift.resize(C);
    for (int i = 0; i < C; ++i) {
      double s = (double)g->data[i] / sqrt((double)v->data[i] + (double)BN_EPS);
      scale[i] = (float)s;
      shift[i] = (float)((double)b->data[i] - (double)m->data[i] * s);
    }
    return true;
  }

  Lin conv_bn(const std::string& conv, const std::string& bnp, int N, int K) {
    const HostTensor* w = get(conv + ".weight", {N, K});
    std::vector<float> sc, sh;
    if (!w || !bn(bnp, N, sc, sh)) return Lin();
    return make_lin(w->data, N, K, &sc, &sh);
  }

  // rows [r0, r0+n) of a [R,K] weight (+ the same slice of its bias)
  Lin linear_rows(const std::string& wname, const std::string& bname, int R, int K, int r0, int n) {
    const HostTensor* w = get(wname, {R, K});
    const HostTensor* b = bname.empty() ? nullptr : get(bname, {R});
    if (!w || (!bname.empty() && !b)) return Lin();
    std::vector<float> ws(w->data.begin() + (size_t)r0 * K, w->data.begin() + (size_t)(r0 + n) * K);
    if (b) {
      std::vector<float> bs(b->data.begin() + r0, b->data.begin() + r0 + n);
      return make_lin(ws, n, K, nullptr, &bs);
    }
    return make_lin(ws, n, K, nullptr, nullptr);
  }
  Lin linear(const std::string& prefix, int N, int K) { return linear_rows(prefix + ".weight", prefix + ".bias", N, K, 0, N); }

  LnP ln(const std::string& prefix, int C) {
    LnP l;
    const HostTensor* g = get(prefix + ".weight", {C});
    const HostTensor* b = get(prefix + ".bias", {C});
    if (!g || !b) return l;
    l.g = upload(g->data); l.b = upload(b->data); l.C = C;
    return l;
  }
};

void host_layernorm(std::vector<double>& x, const std::vector<float>& g, const std::vector<float>& b) {
  const size_t n = x.size();
  double mean = 0, var = 0;
  for (double v : x) mean += v;
  mean /= (double)n;
  for (double v : x) var += (v - mean) * (v - mean);
  var /= (double)n;
  const double rstd = 1.0 / sqrt(var + (double)LN_EPS);
  for (size_t i = 0; i < n; ++i) x[i] = (x[i] - mean) * rstd * (double)g[i] + (double)b[i];
}

// y[n] = sum_k W[r0+n, k] x[k] (+ b[r0+n])
std::vector<double> host_matvec(const HostTensor& w, const HostTensor* b, int r0, int N, int K, const std::vector<double>& x) {
  std::vector<double> y(N);
  for (int n = 0; n < N; ++n) {
    const float* row = w.data.data() + (size_t)(r0 + n) * K;
    double acc = 0;
    for (int k = 0; k < K; ++k) acc += (double)row[k] * x[k];
    y[n] = acc + (b ? (double)b->data[r0 + n] : 0.0);
  }
  return y;
}

// out[q, n] = sum_k E[q,k] W[r0+n,k]   (fp64 accumulate, fp32 result)
void host_embed_proj(const HostTensor& e, int Q, int K, const HostTensor& w, int r0, int N, float* out, int ldo) {
  for (int q = 0; q < Q; ++q) {
    const float* er = e.data.data() + (size_t)q * K;
    for (int n = 0; n < N; ++n) {
      const float* wr = w.data.data() + (size_t)(r0 + n) * K;
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)er[k] * (double)wr[k];
      out[(size_t)q * ldo + n] = (float)acc;
    }
  }
}

int do_finalize(TuberPlan* p) {
  const TuberConfig& c = p->cfg;
  Packer pk{p};
  const int d = c.d_model, ff = c.dim_ff, Q = c.num_queries;
  const std::string bb = "backbone.body";

  // ---- stem (ir_CSN_152.py:109-120): filter (64,3,3,7,7) -> packed split bf16 [2][64][576] ----
  {
    const HostTensor* w = pk.get(bb + ".conv1.weight", {64, 441});
    std::vector<float> sc, sh;
    if (w && pk.bn(bb + ".bn1", 64, sc, sh)) {
      float* wf = pk.upload(w->data);
      void* wp = nullptr;
      if (cudaMalloc(&wp, 2 * 64 * 576 * 2) != cudaSuccess) return fail(TUBER_ERR_CUDA, "cudaMalloc failed");
      p->owned.push_back(wp);
      if (wf) launch_stem_pack_weight(wf, wp, 0);
      p->stem_w = wp;
      p->stem_scale = pk.upload(sc);
      p->stem_shift = pk.upload(sh);
    }
  }
  // ---- bottlenecks (ir_CSN_152.py:36-68,124-170) ----
  const int planes_of[4] = {64, 128, 256, 512};
  const int tstr[4] = {1, 2, 2, 2};
  const int sstr[4] = {1, 2, 2, c.last_stride ? 2 : 1};
  int in_planes = 64;
  for (int li = 0; li < 4 && pk.status == TUBER_OK; ++li) {
    const int planes = planes_of[li], outp = planes * 4;
    p->blocks[li].clear();
    for (int bi = 0; bi < c.blocks[li] && pk.status == TUBER_OK; ++bi) {
      Block b;
      char pre[96];
      snprintf(pre, sizeof pre, "%s.layer%d.%d", bb.c_str(), li + 1, bi);
      const std::string P = pre;
      b.cin = bi == 0 ? in_planes : outp;
      b.planes = planes; b.cout = outp;
      b.has_ds = bi == 0;
      b.st_t = bi == 0 ? tstr[li] : 1;
      b.st_s = bi == 0 ? sstr[li] : 1;
      b.conv1 = pk.conv_bn(P + ".conv1", P + ".bn1", planes, b.cin);
      {
        const HostTensor* w = pk.get(P + ".conv3.weight", {planes, 27});
        std::vector<float> sc, sh;
        if (w && pk.bn(P + ".bn3", planes, sc, sh)) {
          std::vector<float> t((size_t)27 * planes);
          for (int ch = 0; ch < planes; ++ch)
            for (int k = 0; k < 27; ++k) t[(size_t)k * planes + ch] = w->data[(size_t)ch * 27 + k];
          b.dw.w = pk.upload(t); b.dw.scale = pk.upload(sc); b.dw.shift = pk.upload(sh); b.dw.C = planes;
        }
      }
      b.conv4 = pk.conv_bn(P + ".conv4", P + ".bn4", outp, planes);
      if (b.has_ds) {
        // relu(bn4(W4 t2) + bn_d(Wd x)) = relu([t2 | x] [s4*W4 | sd*Wd]^T + (b4 + bd)): one GEMM, no shortcut tensor in HBM
        const HostTensor* w4 = pk.get(P + ".conv4.weight", {outp, planes});
        const HostTensor* wd = pk.get(P + ".down_sample.0.weight", {outp, b.cin});
        std::vector<float> s4, b4, sd, bd;
        if (w4 && wd && pk.bn(P + ".bn4", outp, s4, b4) && pk.bn(P + ".down_sample.1", outp, sd, bd)) {
          const int kt = planes + b.cin;
          std::vector<float> wcat((size_t)outp * kt), bias(outp);
          for (int n = 0; n < outp; ++n) {
            for (int k = 0; k < planes; ++k) wcat[(size_t)n * kt + k] = s4[n] * w4->data[(size_t)n * planes + k];
            for (int k = 0; k < b.cin; ++k) wcat[(size_t)n * kt + planes + k] = sd[n] * wd->data[(size_t)n * b.cin + k];
            bias[n] = b4[n] + bd[n];
          }
          b.c4ds = pk.make_lin(wcat, outp, kt, nullptr, &bias);
        }
      }
      p->blocks[li].push_back(b);
    }
    in_planes = outp;
  }
  if (pk.status != TUBER_OK) return pk.status;
  if (in_planes != POOL_DIM) return fail(TUBER_ERR_INVALID, "backbone width %d != 2048", in_planes);

  // ---- decode pool (backbone_builder.py:49-53; transformer_layers.py:407-448) ----
  if (c.pool == TUBER_POOL_DECODE) {
    const int D = POOL_DIM;
    const std::string L = "backbone.pool_decoder.layers.0";
    const HostTensor* qp = pk.get("backbone.query_pool.weight", {D});
    const HostTensor* sa_w = pk.get(L + ".self_attn.in_proj_weight", {3 * D, D});
    const HostTensor* sa_b = pk.get(L + ".self_attn.in_proj_bias", {3 * D});
    const HostTensor* so_w = pk.get(L + ".self_attn.out_proj.weight", {D, D});
    const HostTensor* so_b = pk.get(L + ".self_attn.out_proj.bias", {D});
    const HostTensor* n1g = pk.get(L + ".norm1.weight", {D});
    const HostTensor* n1b = pk.get(L + ".norm1.bias", {D});
    const HostTensor* ca_w = pk.get(L + ".multihead_attn.in_proj_weight", {3 * D, D});
    const HostTensor* ca_b = pk.get(L + ".multihead_attn.in_proj_bias", {3 * D});
    if (pk.status != TUBER_OK) return pk.status;
    // the pooled query is one token per pixel, so its self-attention softmax is identically 1 and the
    // whole first sub-layer is input independent: tgt1 = LN1(q + Wo (Wv q + bv) + bo)
    std::vector<double> q0(qp->data.begin(), qp->data.end());
    std::vector<double> v = host_matvec(*sa_w, sa_b, 2 * D, D, D, q0);
    std::vector<double> o = host_matvec(*so_w, so_b, 0, D, D, v);
    for (int i = 0; i < D; ++i) o[i] += q0[i];
    host_layernorm(o, n1g->data, n1b->data);
    std::vector<double> qc = host_matvec(*ca_w, ca_b, 0, D, D, o);
    std::vector<float> tgt1(o.begin(), o.end()), qcf(qc.begin(), qc.end());
    p->pool_tgt1 = pk.upload(tgt1);
    p->pool_q = pk.upload(qcf);
    p->pool_kv = pk.linear_rows(L + ".multihead_attn.in_proj_weight", L + ".multihead_attn.in_proj_bias", 3 * D, D, D, 2 * D);
    p->pool_v = pk.linear_rows(L + ".multihead_attn.in_proj_weight", L + ".multihead_attn.in_proj_bias", 3 * D, D, 2 * D, D);
    {
      // scores of head h: (q_h * scale) . (Wk_h x + bk_h) = u_h . x + const, u_h = scale * Wk_h^T q_h; the constant is the same for
      // every frame of a pixel and cancels in the softmax over frames (transformer_layers.py:338-352)
      const int NH = 8, HD = D / NH;
      const double scale = 1.0 / sqrt((double)HD);
      std::vector<float> u((size_t)NH * D);
      for (int h = 0; h < NH; ++h)
        for (int k = 0; k < D; ++k) {
          double acc = 0;
          for (int d2 = 0; d2 < HD; ++d2) acc += (double)ca_w->data[(size_t)(D + h * HD + d2) * D + k] * qc[h * HD + d2];
          u[(size_t)h * D + k] = (float)(acc * scale);
        }
      p->pool_u = pk.upload(u);
    }
    p->pool_out = pk.linear(L + ".multihead_attn.out_proj", D, D);
    p->pool_lin1 = pk.linear(L + ".linear1", 2048, D);
    p->pool_lin2 = pk.linear(L + ".linear2", D, 2048);
    p->pool_n2 = pk.ln(L + ".norm2", D);
    p->pool_n3 = pk.ln(L + ".norm3", D);
    p->pool_nf = pk.ln("backbone.pool_decoder.norm", D);
  }

  // ---- projections (tuber_ava.py:57-58) ----
  p->input_proj = pk.linear("input_proj", d, c.dim_ff);   // backbone.num_channels := DIM_FEEDFORWARD (backbone_builder.py:111)
  p->class_proj = pk.linear("class_proj", d, c.dim_ff);
  if (c.dim_ff != POOL_DIM) return fail(TUBER_ERR_INVALID, "DIM_FEEDFORWARD must equal the backbone width 2048");

  // ---- DETR encoder / decoder (transformer.py:131-149,193-211) ----
  p->enc.clear();
  p->dec.clear();
  const int NP = c.enc_layers * 3 * d + c.dec_layers * 2 * d;
  std::vector<float> posw((size_t)NP * d, 0.f);
  for (int i = 0; i < c.enc_layers && pk.status == TUBER_OK; ++i) {
    char pre[96];
    snprintf(pre, sizeof pre, "transformer.encoder.layers.%d", i);
    const std::string P = pre;
    EncLayer e;
    e.in = pk.linear_rows(P + ".self_attn.in_proj_weight", P + ".self_attn.in_proj_bias", 3 * d, d, 0, 3 * d);
    e.out = pk.linear(P + ".self_attn.out_proj", d, d);
    e.lin1 = pk.linear(P + ".linear1", ff, d);
    e.lin2 = pk.linear(P + ".linear2", d, ff);
    e.n1 = pk.ln(P + ".norm1", d);
    e.n2 = pk.ln(P + ".norm2", d);
    const HostTensor* w = pk.get(P + ".self_attn.in_proj_weight", {3 * d, d});
    if (w) memcpy(posw.data() + (size_t)i * 3 * d * d, w->data.data(), (size_t)2 * d * d * sizeof(float));   // Wq, Wk rows
    p->enc.push_back(e);
  }
  const HostTensor* qe = pk.get("query_embed.weight", {Q, d});
  std::vector<float> memw((size_t)c.dec_layers * 2 * d * d), memb((size_t)c.dec_layers * 2 * d);
  for (int i = 0; i < c.dec_layers && pk.status == TUBER_OK; ++i) {
    char pre[96];
    snprintf(pre, sizeof pre, "transformer.decoder.layers.%d", i);
    const std::string P = pre;
    DecLayer l;
    l.sa_in = pk.linear_rows(P + ".self_attn.in_proj_weight", P + ".self_attn.in_proj_bias", 3 * d, d, 0, 3 * d);
    l.sa_out = pk.linear(P + ".self_attn.out_proj", d, d);
    l.ca_q = pk.linear_rows(P + ".multihead_attn.in_proj_weight", P + ".multihead_attn.in_proj_bias", 3 * d, d, 0, d);
    l.ca_out = pk.linear(P + ".multihead_attn.out_proj", d, d);
    l.lin1 = pk.linear(P + ".linear1", ff, d);
    l.lin2 = pk.linear(P + ".linear2", d, ff);
    l.n1 = pk.ln(P + ".norm1", d);
    l.n2 = pk.ln(P + ".norm2", d);
    l.n3 = pk.ln(P + ".norm3", d);
    const HostTensor* sw = pk.get(P + ".self_attn.in_proj_weight", {3 * d, d});
    const HostTensor* cw = pk.get(P + ".multihead_attn.in_proj_weight", {3 * d, d});
    const HostTensor* cb = pk.get(P + ".multihead_attn.in_proj_bias", {3 * d});
    if (sw && cw && cb && qe) {
      std::vector<float> pq((size_t)Q * 3 * d, 0.f), pc((size_t)Q * d);
      host_embed_proj(*qe, Q, d, *sw, 0, 2 * d, pq.data(), 3 * d);
      host_embed_proj(*qe, Q, d, *cw, 0, d, pc.data(), d);
      l.pq_sa = pk.upload(pq);
      l.pq_ca = pk.upload(pc);
      const HostTensor* so = pk.get(P + ".self_attn.out_proj.weight", {d, d});
      const HostTensor* co = pk.get(P + ".multihead_attn.out_proj.weight", {d, d});
      if (so && co) {
        auto transposed = [&](const float* w) {              // rows [0, d) of w [*, d] -> K-major [d][d]
          std::vector<float> t((size_t)d * d);
          for (int n = 0; n < d; ++n)
            for (int k = 0; k < d; ++k) t[(size_t)k * d + n] = w[(size_t)n * d + k];
          return t;
        };
        l.sa_out_t = pk.upload(transposed(so->data.data()));
        l.ca_q_t = pk.upload(transposed(cw->data.data()));
        l.ca_out_t = pk.upload(transposed(co->data.data()));
      }
      memcpy(memw.data() + (size_t)i * 2 * d * d, cw->data.data() + (size_t)d * d, (size_t)2 * d * d * sizeof(float));
      memcpy(memb.data() + (size_t)i * 2 * d, cb->data.data() + d, (size_t)2 * d * sizeof(float));
      // position term of the cross-attention keys: k = (memory + pos) Wk^T
      memcpy(posw.data() + ((size_t)c.enc_layers * 3 * d + (size_t)i * 2 * d) * d, cw->data.data() + (size_t)d * d,
             (size_t)d * d * sizeof(float));
    }
    p->dec.push_back(l);
  }
  if (pk.status != TUBER_OK) return pk.status;
  {
    // decoder layer 0 sees tgt = 0: q = k = query_pos terms, v_j = bv for every j, so softmax(..) V = bv, the self-attention
    // block returns Wo bv + bo for every query, and after norm1 the state is one constant row c1 (transformer.py:60,225-231)
    const std::string P = "transformer.decoder.layers.0";
    const HostTensor* sb = pk.get(P + ".self_attn.in_proj_bias", {3 * d});
    const HostTensor* ow = pk.get(P + ".self_attn.out_proj.weight", {d, d});
    const HostTensor* ob = pk.get(P + ".self_attn.out_proj.bias", {d});
    const HostTensor* g1 = pk.get(P + ".norm1.weight", {d});
    const HostTensor* b1 = pk.get(P + ".norm1.bias", {d});
    const HostTensor* cw = pk.get(P + ".multihead_attn.in_proj_weight", {3 * d, d});
    const HostTensor* cb = pk.get(P + ".multihead_attn.in_proj_bias", {3 * d});
    if (pk.status != TUBER_OK) return pk.status;
    std::vector<double> bv(sb->data.begin() + 2 * d, sb->data.begin() + 3 * d);
    std::vector<double> c1 = host_matvec(*ow, ob, 0, d, d, bv);
    host_layernorm(c1, g1->data, b1->data);
    std::vector<float> c1f(c1.begin(), c1.end()), qcf((size_t)Q * d);
    for (int q = 0; q < Q; ++q) {
      std::vector<double> x(d);
      for (int k = 0; k < d; ++k) x[k] = c1[k] + (double)qe->data[(size_t)q * d + k];
      std::vector<double> y = host_matvec(*cw, cb, 0, d, d, x);
      for (int n = 0; n < d; ++n) qcf[(size_t)q * d + n] = (float)y[n];
    }
    p->dec0_c1 = pk.upload(c1f);
    p->dec0_qc = pk.upload(qcf);
  }
  p->dec_norm = pk.ln("transformer.decoder.norm", d);
  p->pos_proj = pk.make_lin(posw, NP, d, nullptr, nullptr);
  p->mem_kv = pk.make_lin(memw, c.dec_layers * 2 * d, d, nullptr, &memb);
  {
    // position_encoding.py:22-23,52-57: temperature ** (2 * (i // 2) / n), n = d/4 (t) and 3d/8 (y, x)
    const int nt = d / 8 * 2, ns = d / 8 * 3;
    std::vector<float> dt(nt), ds(ns);
    for (int i = 0; i < nt; ++i) dt[i] = powf(10000.f, (float)(2 * (i / 2)) / (float)nt);
    for (int i = 0; i < ns; ++i) ds[i] = powf(10000.f, (float)(2 * (i / 2)) / (float)ns);
    p->dim_t = pk.upload(dt);
    p->dim_s = pk.upload(ds);
  }

  // ---- class branch (transformer_layers.py:46-64; tuber_ava.py:60-62) ----
  {
    const std::string P = "encoder.layers.0";
    p->ct_in = pk.linear_rows(P + ".self_attn_t.in_proj_weight", P + ".self_attn_t.in_proj_bias", 3 * d, d, 0, 3 * d);
    p->ct_out = pk.linear(P + ".self_attn_t.out_proj", d, d);
    p->cs_in = pk.linear_rows(P + ".self_attn_s.in_proj_weight", P + ".self_attn_s.in_proj_bias", 3 * d, d, 0, 3 * d);
    p->cs_out = pk.linear(P + ".self_attn_s.out_proj", d, d);
    p->c_lin1 = pk.linear(P + ".linear1", CLS_FF, 2 * d);
    p->c_lin2 = pk.linear(P + ".linear2", d, CLS_FF);
    p->c_n1t = pk.ln(P + ".norm1_t", d);
    p->c_n1s = pk.ln(P + ".norm1_s", d);
    p->c_n2 = pk.ln(P + ".norm2", d);
    p->x_q = pk.linear_rows("cross_attn.in_proj_weight", "cross_attn.in_proj_bias", 3 * d, d, 0, d);
    p->x_kv = pk.linear_rows("cross_attn.in_proj_weight", "cross_attn.in_proj_bias", 3 * d, d, d, 2 * d);
    p->x_out = pk.linear("cross_attn.out_proj", d, d);
  }
  // ---- long-term context layer (optional) ----
  if (p->host.count("ltc_attn.in_proj_weight")) {
    p->ltc_q = pk.linear_rows("ltc_attn.in_proj_weight", "ltc_attn.in_proj_bias", 3 * d, d, 0, d);
    p->ltc_kv = pk.linear_rows("ltc_attn.in_proj_weight", "ltc_attn.in_proj_bias", 3 * d, d, d, 2 * d);
    p->ltc_out = pk.linear("ltc_attn.out_proj", d, d);
    p->ltc_n = pk.ln("ltc_norm", d);
    p->has_ltc = true;
  }
  // ---- heads (tuber_ava.py:64-73; criterion.py:485-492) ----
  p->head_b = c.ava_mode ? pk.linear("class_embed_b", 3, d) : pk.linear("class_embed_b", 2, POOL_DIM);
  p->bbox0 = pk.linear("bbox_embed.layers.0", d, d);
  p->bbox1 = pk.linear("bbox_embed.layers.1", d, d);
  p->bbox2 = pk.linear("bbox_embed.layers.2", 4, d);
  p->class_fc = pk.linear("class_fc", c.num_classes, d);
  if (pk.status != TUBER_OK) return pk.status;
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  p->host.clear();
  p->finalized = true;
  return TUBER_OK;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
struct Ctx {
  TuberPlan* p;
  Ws ws;
  cudaStream_t st;
  bool dry;
  int launches = 0;
  int status = TUBER_OK;
  char tag[64] = "";     // optional shape note for the next launch (per-kernel profile dump)

  bool ok() const { return status == TUBER_OK; }
  // ---- side branches: work that does not depend on the main chain runs on p->side[k] between side_begin(k) and side_end(k);
  // side_join(k) makes the main stream wait for it.  With `overlap` off (the default; always in the profiling modes) the
  // three calls do nothing and the branch simply runs in program order on the main stream.
  bool overlap = false;
  cudaStream_t main_st = nullptr;
  int side_open = -1;
  void side_begin(int k) {
    if (!overlap || dry || !ok()) return;
    main_st = st;
    cudaError_t e = cudaEventRecord(p->ev_fork[k], st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(p->side[k], p->ev_fork[k], 0);
    if (e != cudaSuccess) { status = fail(TUBER_ERR_CUDA, "side branch %d: %s", k, cudaGetErrorString(e)); return; }
    st = p->side[k];
    side_open = k;
    if (k == 1) sm_cap_ref() = p->side_ctas;                 // launchers called from this thread size their grids for that many SMs
  }
  void side_end(int k) {
    if (!overlap || dry || side_open != k) return;
    cudaError_t e = cudaEventRecord(p->ev_join[k], st);
    st = main_st;
    side_open = -1;
    sm_cap_ref() = 0;
    if (e != cudaSuccess && ok()) status = fail(TUBER_ERR_CUDA, "side branch %d: %s", k, cudaGetErrorString(e));
  }
  void side_join(int k) {
    if (!overlap || dry) return;                            // (also after an error: a capture in progress needs every branch joined)
    cudaError_t e = cudaStreamWaitEvent(st, p->ev_join[k], 0);
    if (e != cudaSuccess && ok()) status = fail(TUBER_ERR_CUDA, "side branch %d join: %s", k, cudaGetErrorString(e));
  }
  // every device operation of a forward goes through here: counted, optionally bracketed by CUDA events
  // (per-kernel profiling), skipped in the dry (sizing) pass
  // algorithmic bytes / flops of the launches since each stage mark (also counted in the sizing pass)
  int cur_stage = -1;
  double st_bytes[TUBER_NUM_STAGES] = {}, st_flops[TUBER_NUM_STAGES] = {};
  template <class F>
  void launch(const char* name, double bytes, double flops, F&& f) {
    if (!ok()) return;
    ++launches;
    if (cur_stage >= 0) { st_bytes[cur_stage] += bytes; st_flops[cur_stage] += flops; }
    if (dry) return;
    if (ws.overflow) { status = fail(TUBER_ERR_STATE, "workspace overflow before %s (sizing pass disagrees with the forward)", name); return; }
    const int idx = p->kprof ? kp_begin(name, bytes, flops) : -1;
    cudaError_t e = f();
    if (idx >= 0) cudaEventRecord(p->kp[idx].e1, st);
    if (e != cudaSuccess) status = fail(TUBER_ERR_CUDA, "%s: %s %s", name, cudaGetErrorString(e), gemm_tc_last_error());
  }
  int kp_begin(const char* name, double bytes, double flops) {
    if (p->kp_used == (int)p->kp.size()) {
      KernelRec r{};
      if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return -1;
      p->kp.push_back(r);
    }
    KernelRec& r = p->kp[p->kp_used];
    r.name = name; r.bytes = bytes; r.flops = flops;
    snprintf(r.tag, sizeof r.tag, "%s", tag);
    tag[0] = 0;
    cudaEventRecord(r.e0, st);
    return p->kp_used++;
  }
  float* f32(long long rows, int cols) { return reinterpret_cast<float*>(ws.alloc((size_t)rows * cols * 4)); }
  void* split(long long rows, int cols) { return ws.alloc((size_t)rows * cols * 4); }

  // C = act(scale * A W^T + shift + res)
  void gemm(const void* A, int a_fmt, int lda, long long M, const Lin& w, const void* res, int res_fmt, int ldr, int res_mod,
            void* C, int c_fmt, int ldc, int act, const void* Ab = nullptr, int ldb = 0, int Kb = 0, int ksplit = 0, int part_rows = 0,
            const int* ab_geo = nullptr) {
    if (!ok()) return;
    GemmArgs a{};
    a.ksplit = ksplit; a.part_rows = part_rows;
    if (ab_geo) { a.ab_Wo = ab_geo[0]; a.ab_Ho = ab_geo[1]; a.ab_Wi = ab_geo[2]; a.ab_Hi = ab_geo[3]; a.ab_BTi = ab_geo[4]; a.ab_st_t = ab_geo[5]; a.ab_st_s = ab_geo[6]; }
    a.A = A; a.a_fmt = a_fmt; a.lda = lda;
    a.Ab = Ab; a.ldb = ldb; a.Kb = Ab ? Kb : 0;
    a.Wf = w.wf; a.Wp = w.wp; a.scale = w.scale; a.shift = w.shift;
    a.res = res; a.res_fmt = res_fmt; a.ldr = ldr; a.res_mod = res_mod;
    a.C = C; a.c_fmt = c_fmt; a.ldc = ldc;
    a.M = (int)M; a.N = w.N; a.K = w.K - a.Kb; a.act = act;
    const bool tc_ok = a_fmt == FMT_SPLIT && w.N % 64 == 0 && w.K % 64 == 0 && act != ACT_SIGMOID && lda % 8 == 0 && ldc % 8 == 0;
    const bool tc = tc_ok && !p->force_simt;
    if (!tc && ksplit > 1) { status = fail(TUBER_ERR_STATE, "split-K GEMM needs the tensor-core path"); return; }
    if (p->kprof) snprintf(tag, sizeof tag, "M=%lld N=%d K=%d res=%d/%d out=%d act=%d ksplit=%d", M, w.N, w.K, res ? 1 + res_fmt : 0, res_mod, c_fmt, act, ksplit);
    const double mn = (double)M * w.N;
    const double bytes = 4.0 * ((double)M * w.K + (double)w.N * w.K + mn + (res ? (res_mod > 0 ? (double)res_mod * w.N : mn) : 0.0));
    launch(tc ? gemm_tc_config_name(a) : "sgemm_fp32", bytes, 2.0 * mn * w.K,
           [&] { return tc ? launch_gemm_tc(a, st) : launch_sgemm(a, st); });
  }

  // out = LayerNorm(x + res); x fp32 or split; fp32 and / or split outputs
  void layernorm(const void* x, int x_fmt, int ldx, const void* res, int res_fmt, int ldr, const LnP& ln, long long rows,
                 float* out_f32, int ldo, void* out_split, int lds, int split_col_off = 0, int rpg = 0,
                 long long group_stride = 0, long long row_off = 0, const LnP* ln2 = nullptr, void* out2_split = nullptr, int lds2 = 0,
                 int x_parts = 0, long long x_part_stride = 0) {
    if (!ok()) return;
    LnArgs a{};
    a.x_parts = x_parts; a.x_part_stride = x_part_stride;
    if (ln2) { a.gamma2 = ln2->g; a.beta2 = ln2->b; a.out2_split = out2_split; a.lds2 = lds2; }
    a.x = x; a.x_fmt = x_fmt; a.ldx = ldx; a.res = res; a.res_fmt = res_fmt; a.ldr = ldr;
    a.gamma = ln.g; a.beta = ln.b; a.eps = LN_EPS; a.rows = (int)rows; a.C = ln.C;
    a.out_f32 = out_f32; a.ldo = ldo; a.rpg = rpg; a.group_stride = group_stride; a.row_off = row_off;
    a.out_split = out_split; a.lds = lds; a.split_col_off = split_col_off;
    const double n = (double)rows * ln.C;
    launch("layernorm", 4.0 * n * (1 + (res ? 1 : 0) + (out_f32 ? 1 : 0) + (out_split ? 1 : 0) + (ln2 ? 1 : 0)), 8.0 * n * (ln2 ? 2 : 1),
           [&] { return launch_layernorm(a, st); });
  }

  void attention(const float* q, int ldq, SeqMap qm, const float* k, const float* v, int ldkv, SeqMap km, void* o_split, int ldo,
                 SeqMap om, const uint8_t* kpm, int NB, int H, int L, int S, int D) {
    if (!ok()) return;
    AttnArgs a{};
    a.q = q; a.ldq = ldq; a.qm = qm; a.k = k; a.v = v; a.ldk = ldkv; a.ldv = ldkv; a.km = km;
    a.o_f32 = nullptr; a.o_split = o_split; a.ldo = ldo; a.om = om;
    a.kpm = kpm; a.kpm_div = 1; a.NB = NB; a.H = H; a.L = L; a.S = S; a.D = D;
    a.scale = 1.f / sqrtf((float)D);
    // long sequences run on the tcgen05 kernel, which wants K / V pre-converted once into per-chunk tile images (attn_tc.cu)
    static const bool no_tc = [] { const char* e = getenv("TUBER_ATTN_NO_TC"); return e && e[0] == '1'; }();
    static const bool no_prep = [] { const char* e = getenv("TUBER_ATTN_NO_PREP"); return e && e[0] == '1'; }();
    static const bool attn_simt = [] { const char* e = getenv("TUBER_ATTN_SIMT"); return e && e[0] == '1'; }();
    const bool on_tc = !no_tc && !attn_simt && attention_tc_supported(a);
    if (on_tc && !no_prep && attention_tc_wants_prep(a)) {
      a.tc_scratch = ws.alloc(attention_tc_scratch_bytes(a));
      ++launches;                                              // the conversion launch (part of launch_attention_tc)
    }
    if (p->kprof) snprintf(tag, sizeof tag, "NB=%d H=%d L=%d S=%d D=%d", NB, H, L, S, D);
    const double e = (double)H * D;
    launch(on_tc ? "attention_tc" : "attention", 4.0 * e * ((double)NB * L * 2 + (double)NB * S * 2), 4.0 * (double)NB * H * L * S * D,
           [&] { return launch_attention(a, st); });
  }

  void tap(const char* name, const void* ptr, int fmt, long long rows, int cols) {
    if (dry || !ok()) return;
    Tap t{ptr, fmt, rows, cols, nullptr};
    auto it = p->taps.find(name);
    if (it != p->taps.end() && it->second.keep) { cudaFree(it->second.keep); }
    if (p->debug_keep) {
      void* keep = nullptr;
      size_t bytes = (size_t)rows * cols * 4;
      if (cudaMalloc(&keep, bytes) == cudaSuccess) {
        cudaMemcpyAsync(keep, ptr, bytes, cudaMemcpyDeviceToDevice, st);
        t.keep = keep;
        t.ptr = keep;
      }
    }
    p->taps[name] = t;
  }

  void stage_mark(int i) {
    cur_stage = i < TUBER_NUM_STAGES ? i : -1;
    if (dry || !p->profiling || !ok()) return;
    cudaEventRecord(p->ev[i], st);
  }
};

// split-K factor for a token-sized GEMM with a long K (FFN second layer: N = 256, K = 2048): enough parts to put ~128 CTAs
// to work, at least 4 k-blocks each
inline int choose_ksplit(long long M, int N, int K) {
  const long long tiles = ((M + 127) / 128) * (N / 64);
  int s = 1;
  while (s < 8 && tiles * s * 2 <= 160 && (K / 64) % (s * 2) == 0 && K / 64 / (s * 2) >= 4) s *= 2;
  return s;
}

inline SeqMap seqmap(int inner, long long outer, long long inner_stride, long long step) {
  SeqMap m; m.inner = inner; m.outer = outer; m.inner_stride = inner_stride; m.step = step;
  return m;
}

struct Geometry {
  int H1, W1, H2, W2;
  int Ti[4], Hi[4], Wi[4];       // stage INPUT dims
  int Tf, Hf, Wf, Tp;
};

int compute_geometry(const TuberConfig& c, int T, int H, int W, Geometry& g) {
  if (T < 1 || H < 7 || W < 7) return fail(TUBER_ERR_SHAPE, "clip %dx%dx%d too small", T, H, W);
  g.H1 = (H + 6 - 7) / 2 + 1; g.W1 = (W + 6 - 7) / 2 + 1;
  g.H2 = (g.H1 + 2 - 3) / 2 + 1; g.W2 = (g.W1 + 2 - 3) / 2 + 1;
  int t = T, h = g.H2, w = g.W2;
  const int tstr[4] = {1, 2, 2, 2};
  const int sstr[4] = {1, 2, 2, c.last_stride ? 2 : 1};
  for (int li = 0; li < 4; ++li) {
    g.Ti[li] = t; g.Hi[li] = h; g.Wi[li] = w;
    t = (t - 1) / tstr[li] + 1; h = (h - 1) / sstr[li] + 1; w = (w - 1) / sstr[li] + 1;
  }
  g.Tf = t; g.Hf = h; g.Wf = w;
  switch (c.pool) {
    case TUBER_POOL_AVG: case TUBER_POOL_MAX:
      if (c.pool_kernel < 1 || g.Tf < c.pool_kernel)
        return fail(TUBER_ERR_SHAPE, "temporal pool window %d exceeds the %d feature frames (input T=%d)", c.pool_kernel, g.Tf, T);
      g.Tp = g.Tf / c.pool_kernel;
      break;
    case TUBER_POOL_DECODE: case TUBER_POOL_CENTER: g.Tp = 1; break;
    default: g.Tp = g.Tf;
  }
  return TUBER_OK;
}

int run_forward(Ctx& cx, const float* clips, const uint8_t* mask, int B, int T, int H, int W, float* logits, float* boxes,
                float* logits_b) {
  TuberPlan* p = cx.p;
  const TuberConfig& c = p->cfg;
  const bool dry = cx.dry;
  cudaStream_t& st = cx.st;                                 // the CURRENT stream: a side branch switches it (Ctx::side_begin)
  const bool ov = cx.overlap;
  Geometry g;
  TRY(compute_geometry(c, T, H, W, g));
  const int d = c.d_model, nh = c.nhead, hd = d / nh, Q = c.num_queries, Le = c.enc_layers, Ld = c.dec_layers;
  const int Tf = g.Tf, HW = g.Hf * g.Wf, Tp = g.Tp, CB = POOL_DIM;
  const int Ntok = Tp * HW;
  const long long Mc = (long long)B * Tf * HW;             // class-branch tokens
  const long long Mtok = (long long)B * Ntok;              // DETR encoder tokens

  // ---- padding mask at feature resolution, 3-D sine position code and its projections: needs only the mask, so with overlap on
  // it runs as side branch 0 beside the stem (joined in front of the encoder) ----
  const int Bp = mask ? B : 1;                              // without padding every clip has the same code
  const int NP = Le * 3 * d + Ld * 2 * d;
  uint8_t* fmask = nullptr; float* pos = nullptr; float* posp = nullptr;
  auto position_codes = [&] {
    fmask = (uint8_t*)cx.ws.alloc((size_t)Bp * Ntok);
    pos = cx.f32((long long)Bp * Ntok, d);
    void* pos_s = cx.split((long long)Bp * Ntok, d);
    posp = cx.f32((long long)Bp * Ntok, NP);
    const double n = (double)Bp * Ntok * d;
    cx.launch("mask_resize", (double)Bp * Ntok, 0.0, [&] { return launch_mask_resize(mask, fmask, Bp, H, W, Tp, g.Hf, g.Wf, st); });
    cx.launch("posenc", 4.0 * n, 8.0 * n,
              [&] { return launch_posenc(fmask, p->dim_t, p->dim_s, pos, Bp, Tp, g.Hf, g.Wf, d / 8 * 2, d / 8 * 3, st); });
    cx.launch("to_split", 8.0 * n, 0.0, [&] { return launch_to_split(pos, d, pos_s, d, (long long)Bp * Ntok, d, st); });
    cx.gemm(pos_s, FMT_SPLIT, d, (long long)Bp * Ntok, p->pos_proj, nullptr, 0, 0, 0, posp, FMT_F32, NP, ACT_NONE);
  };
  if (ov) {
    cx.side_begin(0);
    position_codes();
    cx.side_end(0);
  }

  // ---- backbone buffers (ping-pong block outputs, conv1 / depthwise / shortcut scratch) ----
  size_t max_out = (size_t)B * T * g.H1 * g.W1 * 64, max_t1 = 0, max_t2 = 0, max_xg = 0;
  {
    size_t x0 = (size_t)B * T * g.H2 * g.W2 * 64;
    if (x0 > max_out) max_out = x0;
    int t = T, h = g.H2, w = g.W2;
    for (int li = 0; li < 4; ++li)
      for (const Block& b : p->blocks[li]) {
        const int to = (t - 1) / b.st_t + 1, ho = (h - 1) / b.st_s + 1, wo = (w - 1) / b.st_s + 1;
        const size_t vin = (size_t)B * t * h * w, vout = (size_t)B * to * ho * wo;
        max_t1 = std::max(max_t1, vin * b.planes);
        max_t2 = std::max(max_t2, vout * b.planes);
        max_out = std::max(max_out, vout * b.cout);
        if (b.has_ds) {
          if (b.st_t != 1 || b.st_s != 1) max_xg = std::max(max_xg, vout * b.cin);
        }
        t = to; h = ho; w = wo;
      }
  }
  char* bufA = (char*)cx.ws.alloc(max_out * 4);
  char* bufB = (char*)cx.ws.alloc(max_out * 4);
  float* t1 = (float*)cx.ws.alloc(max_t1 * 4);
  void* t2 = cx.ws.alloc(max_t2 * 4);
  void* xg = cx.ws.alloc(max_xg * 4 + 16);

  // ---- stem: conv + BN + ReLU (F, NDHWC) then the (1,3,3) max pool (S) ----
  cx.stage_mark(0);
  {
    const double vin = (double)B * 3 * T * H * W, v1 = (double)B * T * g.H1 * g.W1 * 64, v2 = (double)B * T * g.H2 * g.W2 * 64;
    const bool fused = stem_pool_is_fused(g.W1);
    cx.launch("stem_conv", 4.0 * (vin + (fused ? v2 : v1)), 2.0 * 441 * v1, [&] {
      return launch_stem_conv(p->in_frames ? nullptr : clips, p->stem_w, p->stem_scale, p->stem_shift, (float*)bufA, bufB, B, T, H, W, g.H1, g.W1, st,
                              p->in_frames, p->in_frames ? p->in_lut : nullptr);
    });
    if (!fused)
      cx.launch("maxpool", 4.0 * (v1 + v2), 9.0 * v2,
                [&] { return launch_maxpool_hw((const float*)bufA, bufB, B * T, g.H1, g.W1, g.H2, g.W2, 64, st); });
  }
  char* cur = bufB;
  char* nxt = bufA;
  int t = T, h = g.H2, w = g.W2;
  cx.tap("stem", cur, FMT_SPLIT, (long long)B * t * h * w, 64);

  // ---- bottlenecks: 1x1x1 -> depthwise 3x3x3 (stride) -> 1x1x1 (+ shortcut) -> ReLU ----
  bool t1_ready = false;                                     // conv1 of the coming block was already produced by the previous block's fused conv4
  for (int li = 0; li < 4; ++li) {
    cx.stage_mark(1 + li);
    for (size_t bi = 0; bi < p->blocks[li].size(); ++bi) {
      const Block& b = p->blocks[li][bi];
      const Block* nb = bi + 1 < p->blocks[li].size() ? &p->blocks[li][bi + 1] : (li + 1 < 4 ? &p->blocks[li + 1][0] : nullptr);
      const int to = (t - 1) / b.st_t + 1, ho = (h - 1) / b.st_s + 1, wo = (w - 1) / b.st_s + 1;
      const long long vin = (long long)B * t * h * w, vout = (long long)B * to * ho * wo;
      if (!t1_ready) cx.gemm(cur, FMT_SPLIT, b.cin, vin, b.conv1, nullptr, 0, 0, 0, t1, FMT_F32, b.planes, ACT_RELU);
      t1_ready = false;
      cx.launch("dwconv3x3x3", 4.0 * ((double)vin + vout) * b.planes, 54.0 * vout * b.planes,
                [&] { return launch_dwconv(t1, b.dw.w, b.dw.scale, b.dw.shift, t2, B, t, h, w, b.planes, b.st_t, b.st_s, to, ho, wo, st); });
      // 256-channel stage: conv4 of this block and conv1 of the next one run as one kernel (the next block's input tile is
      // multiplied while it still sits in shared memory), see gemm_fused2_kernel
      const bool fuse = !p->force_simt && !p->no_fuse2 && nb && nb->cin == b.cout && b.planes % 64 == 0 && b.cin % 64 == 0 &&
                        ((b.cout == 256 && (nb->planes == 64 || nb->planes == 128)) || (b.cout == 512 && nb->planes == 128) ||
                         (b.cout == 512 && nb->planes == 256 && !p->no_fuse2_s23) ||
                         (b.cout == 1024 && nb->planes == 256 && p->fuse2_deep));
      const void* xa = nullptr;                              // the shortcut's input rows (strided voxel gather when the block strides)
      const int geo[7] = {wo, ho, w, h, B * t, b.st_t, b.st_s};
      const int* ab_geo = nullptr;                           // set: the GEMM reads the strided rows itself (5-D TMA with element strides)
      if (b.has_ds) {
        xa = cur;
        if (b.st_t != 1 || b.st_s != 1) {
          if (!p->force_simt && !p->no_strided_tma && b.cin % 64 == 0 &&
              gemm_tc_strided_ab_ok((int)vout, wo, ho, w, h, B * t, b.st_t, b.st_s)) {
            ab_geo = geo;
          } else {
            cx.launch("gather_rows", 8.0 * vout * b.cin, 0.0,
                      [&] { return launch_gather_rows(cur, xg, b.cin * 4, B, t, h, w, b.st_t, b.st_s, to, ho, wo, st); });
            xa = xg;
          }
        }
      }
      if (fuse) {
        GemmArgs a{};
        const Lin& w4 = b.has_ds ? b.c4ds : b.conv4;
        a.A = t2; a.a_fmt = FMT_SPLIT; a.lda = b.planes;
        a.Ab = b.has_ds ? xa : nullptr; a.ldb = b.cin; a.Kb = b.has_ds ? b.cin : 0;
        if (ab_geo) { a.ab_Wo = geo[0]; a.ab_Ho = geo[1]; a.ab_Wi = geo[2]; a.ab_Hi = geo[3]; a.ab_BTi = geo[4]; a.ab_st_t = geo[5]; a.ab_st_s = geo[6]; }
        a.Wf = w4.wf; a.Wp = w4.wp; a.scale = w4.scale; a.shift = w4.shift;
        a.res = b.has_ds ? nullptr : cur; a.res_fmt = FMT_SPLIT; a.ldr = b.cin; a.res_mod = 0;
        a.C = nxt; a.c_fmt = FMT_SPLIT; a.ldc = b.cout;
        a.M = (int)vout; a.N = b.cout; a.K = b.planes; a.act = ACT_RELU;
        const Lin& c1 = nb->conv1;
        if (p->kprof) snprintf(cx.tag, sizeof cx.tag, "M=%lld N=%d K=%d -> N2=%d", vout, b.cout, w4.K, c1.N);
        const double mn = (double)vout * b.cout;
        cx.launch("gemm_fused2_tcgen05", 4.0 * ((double)vout * w4.K + (double)w4.N * w4.K + mn * (b.has_ds ? 1 : 2) + (double)vout * c1.N + (double)c1.N * c1.K),
                  2.0 * mn * w4.K + 2.0 * (double)vout * c1.N * c1.K,
                  [&] { return launch_gemm_tc_fused2(a, c1.wp, c1.scale, c1.shift, t1, c1.N, c1.N, st); });
        t1_ready = true;
      } else if (b.has_ds) {
        cx.gemm(t2, FMT_SPLIT, b.planes, vout, b.c4ds, nullptr, 0, 0, 0, nxt, FMT_SPLIT, b.cout, ACT_RELU, xa, b.cin, b.cin, 0, 0, ab_geo);
      } else {
        cx.gemm(t2, FMT_SPLIT, b.planes, vout, b.conv4, cur, FMT_SPLIT, b.cin, 0, nxt, FMT_SPLIT, b.cout, ACT_RELU);
      }
      std::swap(cur, nxt);
      t = to; h = ho; w = wo;
    }
    static const char* names[4] = {"layer1", "layer2", "layer3", "layer4"};
    cx.tap(names[li], cur, FMT_SPLIT, (long long)B * t * h * w, p->blocks[li].back().cout);
  }
  if (!cx.ok()) return cx.status;
  const void* xt = cur;                                    // S [B*Tf*HW, 2048]
  cx.tap("xt", xt, FMT_SPLIT, Mc, CB);

  // ---- temporal pooling (backbone_builder.py:70-80) ----
  cx.stage_mark(5);
  const void* xs = xt;
  if (c.pool == TUBER_POOL_AVG || c.pool == TUBER_POOL_MAX) {
    void* o = cx.split(Mtok, CB);
    cx.launch("tpool", 4.0 * ((double)Mc + Mtok) * CB, (double)Mc * CB,
              [&] { return launch_tpool(xt, o, B, Tf, HW, CB, c.pool_kernel, Tp, c.pool == TUBER_POOL_MAX, st); });
    xs = o;
  } else if (c.pool == TUBER_POOL_CENTER) {
    void* o = cx.split(Mtok, CB);
    cx.launch("slice_frames", 8.0 * Mtok * CB, 0.0, [&] { return launch_slice_frames(xt, o, CB * 4, B, Tf, HW, Tf / 2, 1, st); });
    xs = o;
  } else if (c.pool == TUBER_POOL_DECODE) {
    // per pixel: one learned query attends over the Tf frame tokens (d = 2048, 8 heads x 256)
    const long long Mp = (long long)B * HW;
    void* att = cx.split(Mp, CB);
    if (Tf <= 8 && !p->force_simt && !p->pool_unfolded) {
      // folded form: scores from u_h . x_t, frames mixed per head BEFORE the value projection, which then runs as a grouped
      // GEMM over the 8 heads (rows = (head, pixel), W rows = that head's slice of Wv): no K / V tensor of all tokens
      const long long Mg = (Mp + 127) / 128 * 128;            // group stride: a multiple of the GEMM's row block (padding rows are never stored)
      void* mix = cx.split(8 * Mg, CB);
      cx.launch("pool_mix", 4.0 * ((double)Mc + 8.0 * Mp) * CB, 2.0 * (double)Mc * CB * 16,
                [&] { return launch_pool_mix(xt, p->pool_u, mix, B, Tf, HW, Mg, st); });
      GemmArgs a{};
      a.A = mix; a.a_fmt = FMT_SPLIT; a.lda = CB;
      a.Wf = p->pool_v.wf; a.Wp = p->pool_v.wp; a.scale = nullptr; a.shift = p->pool_v.shift;
      a.C = att; a.c_fmt = FMT_SPLIT; a.ldc = CB;
      a.M = (int)(8 * Mg); a.N = CB / 8; a.K = CB; a.act = ACT_NONE; a.group_rows = (int)Mg; a.group_out_rows = (int)Mp;
      if (p->kprof) snprintf(cx.tag, sizeof cx.tag, "grouped 8 x (M=%lld N=%d K=%d)", Mp, CB / 8, CB);
      cx.launch("gemm_bf16x3_deep", 4.0 * (8.0 * Mp * CB + (double)CB * CB + (double)Mp * CB), 2.0 * 8.0 * Mp * (CB / 8) * CB,
                [&] { return launch_gemm_tc(a, st); });
    } else {
      float* kv = cx.f32(Mc, 2 * CB);
      cx.gemm(xt, FMT_SPLIT, CB, Mc, p->pool_kv, nullptr, 0, 0, 0, kv, FMT_F32, 2 * CB, ACT_NONE);
      cx.attention(p->pool_q, CB, seqmap(1, 0, 0, 0), kv, kv + CB, 2 * CB, seqmap(HW, (long long)Tf * HW, 1, HW), att, CB,
                   seqmap(1, 1, 0, 1), nullptr, (int)Mp, 8, 1, Tf, CB / 8);
    }
    void* o1 = cx.split(Mp, CB);
    cx.gemm(att, FMT_SPLIT, CB, Mp, p->pool_out, nullptr, 0, 0, 0, o1, FMT_SPLIT, CB, ACT_NONE);
    void* t2s = cx.split(Mp, CB);
    cx.layernorm(o1, FMT_SPLIT, CB, p->pool_tgt1, FMT_F32, 0, p->pool_n2, Mp, nullptr, 0, t2s, CB);
    void* hdn = cx.split(Mp, 2048);
    cx.gemm(t2s, FMT_SPLIT, CB, Mp, p->pool_lin1, nullptr, 0, 0, 0, hdn, FMT_SPLIT, 2048, ACT_RELU);
    cx.gemm(hdn, FMT_SPLIT, 2048, Mp, p->pool_lin2, t2s, FMT_SPLIT, CB, 0, o1, FMT_SPLIT, CB, ACT_NONE);
    void* t3 = cx.split(Mp, CB);
    void* o = cx.split(Mp, CB);
    cx.layernorm(o1, FMT_SPLIT, CB, nullptr, 0, 0, p->pool_n3, Mp, nullptr, 0, t3, CB, 0, 0, 0, 0, &p->pool_nf, o, CB);   // norm3, then the decoder's final norm
    xs = o;
  }
  cx.tap("xs", xs, FMT_SPLIT, Mtok, CB);

  // ---- padding mask at feature resolution, 3-D sine position code and its projections ----
  cx.stage_mark(6);
  if (ov) cx.side_join(0);
  else position_codes();
  const uint8_t* kpm = mask ? fmask : nullptr;
  const int pos_mod = mask ? 0 : Ntok;
  cx.tap("pos", pos, FMT_F32, (long long)Bp * Ntok, d);

  // input_proj / class_proj (tuber_ava.py:119,129); token activations live in split format from here on
  void* src_s = cx.split(Mtok, d);
  cx.gemm(xs, FMT_SPLIT, CB, Mtok, p->input_proj, nullptr, 0, 0, 0, src_s, FMT_SPLIT, d, ACT_NONE);
  void* srcc_s = cx.split(Mc, d);
  cx.gemm(xt, FMT_SPLIT, CB, Mc, p->class_proj, nullptr, 0, 0, 0, srcc_s, FMT_SPLIT, d, ACT_NONE);
  cx.tap("src", src_s, FMT_SPLIT, Mtok, d);
  // long-term context: this batch's bank entries = class_proj features averaged over the Tf feature frames, [B, HW, d] fp32
  if (p->ltc_new) {
    void* e_s = cx.split((long long)B * HW, d);
    cx.launch("tpool", 4.0 * ((double)Mc + (double)B * HW) * d, (double)Mc * d,
              [&] { return launch_tpool(srcc_s, e_s, B, Tf, HW, d, Tf, 1, 0, st); });
    cx.launch("from_split", 8.0 * B * HW * d, 0.0, [&] { return launch_from_split(e_s, d, p->ltc_new, d, (long long)B * HW, d, st); });
  }

  // The class branch up to the keys / values of its cross attention depends only on class_proj's output: with overlap on it is side
  // branch 1, running beside the DETR encoder and decoder (token-sized launches that leave most SMs idle) and joined in front
  // of the class cross-attention; otherwise it runs here in program order.
  void* memc_s = nullptr;
  float* kvx = nullptr;
  auto class_memory = [&] {
    // class-branch encoder layer (transformer_layers.py:71-97), evaluated once per clip: the reference runs
    // DEC_LAYERS identical replicas of it (tuber_ava.py:133-135)
    memc_s = cx.split(Mc, d);
    {
      float* qkv = cx.f32(Mc, 3 * d);
      void* att = cx.split(Mc, d);
      void* o = cx.split(Mc, d);
      void* cat = cx.split(Mc, 2 * d);
      void* hdn = cx.split(Mc, CLS_FF);
      const int ch = d / CLS_HEADS;
      // "_t": attention inside a frame over its HW positions
      cx.gemm(srcc_s, FMT_SPLIT, d, Mc, p->ct_in, nullptr, 0, 0, 0, qkv, FMT_F32, 3 * d, ACT_NONE);
      cx.attention(qkv, 3 * d, seqmap(1, HW, 0, 1), qkv + d, qkv + 2 * d, 3 * d, seqmap(1, HW, 0, 1), att, d, seqmap(1, HW, 0, 1),
                   nullptr, B * Tf, CLS_HEADS, HW, HW, ch);
      cx.gemm(att, FMT_SPLIT, d, Mc, p->ct_out, srcc_s, FMT_SPLIT, d, 0, o, FMT_SPLIT, d, ACT_NONE);
      cx.layernorm(o, FMT_SPLIT, d, nullptr, 0, 0, p->c_n1t, Mc, nullptr, 0, cat, 2 * d, 0);
      // "_s": attention inside a pixel over its Tf frames
      cx.gemm(srcc_s, FMT_SPLIT, d, Mc, p->cs_in, nullptr, 0, 0, 0, qkv, FMT_F32, 3 * d, ACT_NONE);
      const SeqMap pix = seqmap(HW, (long long)Tf * HW, 1, HW);
      cx.attention(qkv, 3 * d, pix, qkv + d, qkv + 2 * d, 3 * d, pix, att, d, pix, nullptr, B * HW, CLS_HEADS, Tf, Tf, ch);
      cx.gemm(att, FMT_SPLIT, d, Mc, p->cs_out, srcc_s, FMT_SPLIT, d, 0, o, FMT_SPLIT, d, ACT_NONE);
      cx.layernorm(o, FMT_SPLIT, d, nullptr, 0, 0, p->c_n1s, Mc, nullptr, 0, cat, 2 * d, d);
      cx.gemm(cat, FMT_SPLIT, 2 * d, Mc, p->c_lin1, nullptr, 0, 0, 0, hdn, FMT_SPLIT, CLS_FF, ACT_RELU);
      cx.gemm(hdn, FMT_SPLIT, CLS_FF, Mc, p->c_lin2, srcc_s, FMT_SPLIT, d, 0, o, FMT_SPLIT, d, ACT_NONE);
      cx.layernorm(o, FMT_SPLIT, d, nullptr, 0, 0, p->c_n2, Mc, nullptr, 0, memc_s, d);
    }
    cx.tap("mem_c", memc_s, FMT_SPLIT, Mc, d);
    // long-term context layer: every class-branch token of a clip attends over the bank window (bank_clips = 1: one window shared
    // by the batch; = B: one per clip), post-norm like every other layer of the model:  mem_c <- LN(mem_c + MHA(mem_c, bank, bank))
    if (p->ltc_bank && p->ltc_bank_tokens > 0) {
      const int Nb = p->ltc_bank_tokens, Bb = p->ltc_bank_clips, Nc = Tf * HW;
      const long long Mb = (long long)Bb * Nb;
      void* bank_s = cx.split(Mb, d);
      float* kvb = cx.f32(Mb, 2 * d);
      float* ql = cx.f32(Mc, d);
      void* att = cx.split(Mc, d);
      void* o = cx.split(Mc, d);
      void* memc2 = cx.split(Mc, d);
      cx.launch("to_split", 8.0 * Mb * d, 0.0, [&] { return launch_to_split(p->ltc_bank, d, bank_s, d, Mb, d, st); });
      cx.gemm(bank_s, FMT_SPLIT, d, Mb, p->ltc_kv, nullptr, 0, 0, 0, kvb, FMT_F32, 2 * d, ACT_NONE);
      cx.gemm(memc_s, FMT_SPLIT, d, Mc, p->ltc_q, nullptr, 0, 0, 0, ql, FMT_F32, d, ACT_NONE);
      cx.attention(ql, d, seqmap(1, Nc, 0, 1), kvb, kvb + d, 2 * d, seqmap(1, Bb == 1 ? 0 : Nb, 0, 1), att, d, seqmap(1, Nc, 0, 1), nullptr,
                   B, CLS_HEADS, Nc, Nb, d / CLS_HEADS);
      cx.gemm(att, FMT_SPLIT, d, Mc, p->ltc_out, memc_s, FMT_SPLIT, d, 0, o, FMT_SPLIT, d, ACT_NONE);
      cx.layernorm(o, FMT_SPLIT, d, nullptr, 0, 0, p->ltc_n, Mc, nullptr, 0, memc2, d);
      memc_s = memc2;
      cx.tap("mem_ltc", memc_s, FMT_SPLIT, Mc, d);
    }
    kvx = cx.f32(Mc, 2 * d);
    cx.gemm(memc_s, FMT_SPLIT, d, Mc, p->x_kv, nullptr, 0, 0, 0, kvx, FMT_F32, 2 * d, ACT_NONE);
  };
  if (ov) {
    cx.side_begin(1);
    class_memory();
    cx.side_end(1);
  }

  // ---- DETR encoder (transformer.py:153-168): post-norm, q = k = src + pos, v = src ----
  cx.stage_mark(7);
  {
    float* qkv = cx.f32(Mtok, 3 * d);
    void* att = cx.split(Mtok, d);
    void* o = cx.split(Mtok, d);
    void* hdn = cx.split(Mtok, c.dim_ff);
    const int ks2 = p->force_simt ? 1 : choose_ksplit(Mtok, d, c.dim_ff);
    const long long pr2 = (Mtok + 127) / 128 * 128;          // rows between partial-sum slabs (whole row blocks: a tile's padding rows stay inside its slab)
    float* o2 = ks2 > 1 ? cx.f32((long long)ks2 * pr2, d) : nullptr;
    for (int i = 0; i < Le; ++i) {
      const EncLayer& e = p->enc[i];
      cx.gemm(src_s, FMT_SPLIT, d, Mtok, e.in, posp + (size_t)i * 3 * d, FMT_F32, NP, pos_mod, qkv, FMT_F32, 3 * d, ACT_NONE);
      cx.attention(qkv, 3 * d, seqmap(1, Ntok, 0, 1), qkv + d, qkv + 2 * d, 3 * d, seqmap(1, Ntok, 0, 1), att, d,
                   seqmap(1, Ntok, 0, 1), kpm, B, nh, Ntok, Ntok, hd);
      cx.gemm(att, FMT_SPLIT, d, Mtok, e.out, src_s, FMT_SPLIT, d, 0, o, FMT_SPLIT, d, ACT_NONE);
      cx.layernorm(o, FMT_SPLIT, d, nullptr, 0, 0, e.n1, Mtok, nullptr, 0, src_s, d);
      cx.gemm(src_s, FMT_SPLIT, d, Mtok, e.lin1, nullptr, 0, 0, 0, hdn, FMT_SPLIT, c.dim_ff, ACT_RELU);
      if (ks2 > 1) {     // split-K partial sums (fp32 slabs of o2), added by the LayerNorm kernel
        cx.gemm(hdn, FMT_SPLIT, c.dim_ff, Mtok, e.lin2, src_s, FMT_SPLIT, d, 0, o2, FMT_F32, d, ACT_NONE, nullptr, 0, 0, ks2, (int)pr2);
        cx.layernorm(o2, FMT_F32, d, nullptr, 0, 0, e.n2, Mtok, nullptr, 0, src_s, d, 0, 0, 0, 0, nullptr, nullptr, 0, ks2, pr2);
      } else {
        cx.gemm(hdn, FMT_SPLIT, c.dim_ff, Mtok, e.lin2, src_s, FMT_SPLIT, d, 0, o, FMT_SPLIT, d, ACT_NONE);
        cx.layernorm(o, FMT_SPLIT, d, nullptr, 0, 0, e.n2, Mtok, nullptr, 0, src_s, d);
      }
    }
  }
  cx.tap("memory", src_s, FMT_SPLIT, Mtok, d);

  // ---- DETR decoder (transformer.py:218-249), all layers' outputs kept (return_intermediate) ----
  cx.stage_mark(8);
  const long long Mq = (long long)B * Q, Mh = (long long)B * Ld * Q;
  void* hs_s = cx.split(Mh, d);                             // [B, Ld, Q, d]
  {
    float* memkv = cx.f32(Mtok, Ld * 2 * d);                // per layer [K | V] of the cross attention
    cx.gemm(src_s, FMT_SPLIT, d, Mtok, p->mem_kv, posp + (size_t)Le * 3 * d, FMT_F32, NP, pos_mod, memkv, FMT_F32, Ld * 2 * d,
            ACT_NONE);
    const bool mega = !p->force_simt && !p->no_dec_mega && p->dec0_c1 && p->dec0_qc && Ld <= DEC_MEGA_MAX_LAYERS &&
                      decoder_mega_supported(d, nh, c.dim_ff, Q, Ntok) && p->dec[0].sa_out_t != nullptr;
    if (mega) {
      // the whole stack as one persistent cooperative kernel (decoder_mega.cu): fp32 state, grid barriers between the phases
      DecMegaArgs m{};
      for (int i = 0; i < Ld; ++i) {
        const DecLayer& l = p->dec[i];
        DecMegaLayer& o = m.L[i];
        o.sa_in_w = l.sa_in.wf; o.sa_in_b = l.sa_in.shift; o.pq_sa = l.pq_sa;
        o.sa_out_t = l.sa_out_t; o.sa_out_b = l.sa_out.shift; o.n1_g = l.n1.g; o.n1_b = l.n1.b;
        o.ca_q_t = l.ca_q_t; o.ca_q_b = l.ca_q.shift; o.pq_ca = l.pq_ca;
        o.ca_out_t = l.ca_out_t; o.ca_out_b = l.ca_out.shift; o.n2_g = l.n2.g; o.n2_b = l.n2.b;
        o.lin1_w = l.lin1.wf; o.lin1_b = l.lin1.shift; o.lin2_w = l.lin2.wf; o.lin2_b = l.lin2.shift;
        o.n3_g = l.n3.g; o.n3_b = l.n3.b;
      }
      m.Ld = Ld; m.B = B; m.Q = Q; m.Ntok = Ntok; m.dim_ff = c.dim_ff; m.kv_ld = Ld * 2 * d; m.eps = LN_EPS;
      m.memkv = memkv; m.kpm = kpm; m.dec0_c1 = p->dec0_c1; m.dec0_qc = p->dec0_qc; m.nf_g = p->dec_norm.g; m.nf_b = p->dec_norm.b;
      m.tgt = cx.f32(Mq, d); m.qkv = cx.f32(Mq, 3 * d); m.part = cx.f32((long long)(c.dim_ff / 16) * Mq, d);
      m.hs = hs_s; m.barrier = reinterpret_cast<unsigned*>(cx.ws.alloc(256));
      m.trace = reinterpret_cast<unsigned long long*>(cx.ws.alloc(8 * 256));
      if (!dry) { p->dec_trace = p->kprof ? m.trace : nullptr; p->dec_trace_n = 1 + 4 * Ld; }
      if (!p->kprof) m.trace = nullptr;
      const double wbytes = 4.0 * Ld * (6.0 * d * d + 2.0 * d * c.dim_ff);
      if (p->kprof) snprintf(cx.tag, sizeof cx.tag, "B=%d Q=%d Ntok=%d layers=%d", B, Q, Ntok, Ld);
      cx.launch("decoder_mega", wbytes + 4.0 * Mtok * Ld * 2 * d + 4.0 * Mh * d,
                2.0 * Ld * ((double)Mq * (6.0 * d * d + 2.0 * d * c.dim_ff) + 4.0 * Mq * (Q + Ntok) * d),
                [&] { return launch_decoder_mega(m, st); });
    }
    void* tgt_s = cx.split(Mq, d);
    // tgt = zeros_like(query_embed), transformer.py:60 (all-zero bits are zero in split format too); with layer 0 folded at
    // finalize the zero state is never read
    if (p->dec0_c1 == nullptr)
      cx.launch("memset", 4.0 * Mq * d, 0.0, [&] { return cudaMemsetAsync(tgt_s, 0, (size_t)Mq * d * 4, st); });
    float* qkv = cx.f32(Mq, 3 * d);
    float* qc = cx.f32(Mq, d);
    void* att = cx.split(Mq, d);
    void* o = cx.split(Mq, d);
    void* hdn = cx.split(Mq, c.dim_ff);
    const int ksd = p->force_simt ? 1 : choose_ksplit(Mq, d, c.dim_ff);
    const long long prd = (Mq + 127) / 128 * 128;
    float* od = ksd > 1 ? cx.f32((long long)ksd * prd, d) : nullptr;
    for (int i = 0; i < Ld && !mega; ++i) {
      const DecLayer& l = p->dec[i];
      const bool folded0 = i == 0 && p->dec0_c1 != nullptr;
      if (!folded0) {
        cx.gemm(tgt_s, FMT_SPLIT, d, Mq, l.sa_in, l.pq_sa, FMT_F32, 3 * d, Q, qkv, FMT_F32, 3 * d, ACT_NONE);
        cx.attention(qkv, 3 * d, seqmap(1, Q, 0, 1), qkv + d, qkv + 2 * d, 3 * d, seqmap(1, Q, 0, 1), att, d, seqmap(1, Q, 0, 1),
                     nullptr, B, nh, Q, Q, hd);
        cx.gemm(att, FMT_SPLIT, d, Mq, l.sa_out, tgt_s, FMT_SPLIT, d, 0, o, FMT_SPLIT, d, ACT_NONE);
        cx.layernorm(o, FMT_SPLIT, d, nullptr, 0, 0, l.n1, Mq, nullptr, 0, tgt_s, d);
        cx.gemm(tgt_s, FMT_SPLIT, d, Mq, l.ca_q, l.pq_ca, FMT_F32, d, Q, qc, FMT_F32, d, ACT_NONE);
      }
      // layer 0 starts from tgt = 0 (transformer.py:60), so its self-attention block and cross-attention query are input
      // independent and were folded at finalize: tgt after norm1 = dec0_c1 (one row for every query), query = dec0_qc [Q, d]
      cx.attention(folded0 ? p->dec0_qc : qc, d, folded0 ? seqmap(1, 0, 0, 1) : seqmap(1, Q, 0, 1), memkv + (size_t)i * 2 * d,
                   memkv + (size_t)i * 2 * d + d, Ld * 2 * d, seqmap(1, Ntok, 0, 1), att, d, seqmap(1, Q, 0, 1), kpm, B, nh, Q, Ntok, hd);
      if (folded0) cx.gemm(att, FMT_SPLIT, d, Mq, l.ca_out, p->dec0_c1, FMT_F32, d, 1, o, FMT_SPLIT, d, ACT_NONE);
      else cx.gemm(att, FMT_SPLIT, d, Mq, l.ca_out, tgt_s, FMT_SPLIT, d, 0, o, FMT_SPLIT, d, ACT_NONE);
      cx.layernorm(o, FMT_SPLIT, d, nullptr, 0, 0, l.n2, Mq, nullptr, 0, tgt_s, d);
      cx.gemm(tgt_s, FMT_SPLIT, d, Mq, l.lin1, nullptr, 0, 0, 0, hdn, FMT_SPLIT, c.dim_ff, ACT_RELU);
      // norm3, then the shared final norm on every layer's output (transformer.py:116-126), row (b, q) -> (b, i, q)
      if (ksd > 1) {
        cx.gemm(hdn, FMT_SPLIT, c.dim_ff, Mq, l.lin2, tgt_s, FMT_SPLIT, d, 0, od, FMT_F32, d, ACT_NONE, nullptr, 0, 0, ksd, (int)prd);
        cx.layernorm(od, FMT_F32, d, nullptr, 0, 0, l.n3, Mq, nullptr, 0, tgt_s, d, 0, Q, (long long)Ld * Q, (long long)i * Q, &p->dec_norm, hs_s, d,
                     ksd, prd);
      } else {
        cx.gemm(hdn, FMT_SPLIT, c.dim_ff, Mq, l.lin2, tgt_s, FMT_SPLIT, d, 0, o, FMT_SPLIT, d, ACT_NONE);
        cx.layernorm(o, FMT_SPLIT, d, nullptr, 0, 0, l.n3, Mq, nullptr, 0, tgt_s, d, 0, Q, (long long)Ld * Q, (long long)i * Q, &p->dec_norm, hs_s, d);
      }
    }
  }
  cx.tap("hs", hs_s, FMT_SPLIT, Mh, d);

  // ---- class branch + heads ----
  cx.stage_mark(9);
  // actor-ness head (tuber_ava.py:121-125)
  if (c.ava_mode) {
    cx.gemm(hs_s, FMT_SPLIT, d, Mh, p->head_b, nullptr, 0, 0, 0, logits_b, FMT_F32, 3, ACT_NONE);
  } else {
    float* gap = cx.f32(B, CB);
    cx.launch("global_avgpool", 4.0 * Mc * CB, (double)Mc * CB, [&] { return launch_global_avgpool(xt, gap, B, Tf * HW, CB, st); });
    cx.gemm(gap, FMT_F32, CB, B, p->head_b, nullptr, 0, 0, 0, logits_b, FMT_F32, 2, ACT_NONE);
  }
  // box head (criterion.py:494-497) + sigmoid (tuber_ava.py:142)
  {
    void* h1 = cx.split(Mh, d);
    void* h2 = cx.split(Mh, d);
    cx.gemm(hs_s, FMT_SPLIT, d, Mh, p->bbox0, nullptr, 0, 0, 0, h1, FMT_SPLIT, d, ACT_RELU);
    cx.gemm(h1, FMT_SPLIT, d, Mh, p->bbox1, nullptr, 0, 0, 0, h2, FMT_SPLIT, d, ACT_RELU);
    cx.gemm(h2, FMT_SPLIT, d, Mh, p->bbox2, nullptr, 0, 0, 0, boxes, FMT_F32, 4, ACT_SIGMOID);
  }
  if (!ov) class_memory();
  // class cross-attention (tuber_ava.py:137-139) + class_fc (:141; Dropout(0.5) is the identity in eval)
  {
    const int Nc = Tf * HW, LQ = Ld * Q;
    float* qx = cx.f32(Mh, d);
    void* att = cx.split(Mh, d);
    void* oc = cx.split(Mh, d);
    cx.gemm(hs_s, FMT_SPLIT, d, Mh, p->x_q, nullptr, 0, 0, 0, qx, FMT_F32, d, ACT_NONE);
    if (ov) cx.side_join(1);
    cx.attention(qx, d, seqmap(1, LQ, 0, 1), kvx, kvx + d, 2 * d, seqmap(1, Nc, 0, 1), att, d, seqmap(1, LQ, 0, 1), nullptr, B,
                 CLS_HEADS, LQ, Nc, d / CLS_HEADS);
    cx.gemm(att, FMT_SPLIT, d, Mh, p->x_out, nullptr, 0, 0, 0, oc, FMT_SPLIT, d, ACT_NONE);
    cx.gemm(oc, FMT_SPLIT, d, Mh, p->class_fc, nullptr, 0, 0, 0, logits, FMT_F32, c.num_classes, ACT_NONE);
  }
  cx.stage_mark(TUBER_NUM_STAGES);
  return cx.status;
}

int ensure_workspace(TuberPlan* p, int B, int T, int H, int W, bool has_mask, size_t* need_out, int* launches_out) {
  Ctx cx{p};
  cx.dry = true;
  cx.ws.dry = true;
  cx.st = 0;
  cx.overlap = !p->no_overlap && !p->profiling && !p->debug_keep && !p->kprof;   // same program order as the forward (same sizes either way)
  // the sizing pass must see the same mask / no-mask choice as the real one (per-clip position codes)
  const uint8_t* mask_tag = has_mask ? reinterpret_cast<const uint8_t*>((uintptr_t)0x1000) : nullptr;
  TRY(run_forward(cx, nullptr, mask_tag, B, T, H, W, nullptr, nullptr, nullptr));
  size_t need = cx.ws.peak + 4096;
  if (need_out) *need_out = need;
  if (launches_out) *launches_out = cx.launches;
  return TUBER_OK;
}

int check_forward_args(TuberPlan* p, int B, int T, int H, int W) {
  if (!p) return fail(TUBER_ERR_INVALID, "null plan");
  if (!p->finalized) return fail(TUBER_ERR_STATE, "tuber_forward before tuber_plan_finalize");
  if (B < 1) return fail(TUBER_ERR_INVALID, "batch %d", B);
  (void)T; (void)H; (void)W;
  return TUBER_OK;
}

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int tuber_abi_version(void) { return TUBER_ABI_VERSION; }
const char* tuber_last_error(void) { return g_last_error; }

int tuber_plan_create(const TuberConfig* cfg, TuberPlan** out_plan) {
  if (!cfg || !out_plan) return fail(TUBER_ERR_INVALID, "null argument");
  if (cfg->abi_version != TUBER_ABI_VERSION) return fail(TUBER_ERR_INVALID, "ABI version %d != %d", cfg->abi_version, TUBER_ABI_VERSION);
  if (cfg->d_model != 256 || cfg->nhead != 8)
    return fail(TUBER_ERR_INVALID, "d_model %d / nhead %d unsupported (kernels are built for 256 / 8)", cfg->d_model, cfg->nhead);
  if (cfg->dim_ff % 64 != 0 || cfg->enc_layers < 1 || cfg->dec_layers < 1 || cfg->num_queries < 1 || cfg->num_classes < 1)
    return fail(TUBER_ERR_INVALID, "bad transformer dimensions");
  if (cfg->pool < TUBER_POOL_AVG || cfg->pool > TUBER_POOL_NONE) return fail(TUBER_ERR_INVALID, "bad pool mode %d", cfg->pool);
  for (int i = 0; i < 4; ++i)
    if (cfg->blocks[i] < 1) return fail(TUBER_ERR_INVALID, "stage %d has no blocks", i);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(TUBER_ERR_CUDA, "no CUDA device: the tuber_b200 path has no CPU fallback");
  cudaDeviceProp prop;
  int dev = 0;
  cudaGetDevice(&dev);
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(TUBER_ERR_CUDA, "device '%s' is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
  TuberPlan* p = new TuberPlan();
  p->cfg = *cfg;
  p->device = dev;
  const char* fs = getenv("TUBER_FORCE_SIMT");
  p->force_simt = fs && fs[0] == '1';
  const char* nf = getenv("TUBER_NO_FUSE2");
  p->no_fuse2 = nf && nf[0] == '1';
  // 1024-channel stage: the fused conv4 -> conv1 kernels exist (gemm_fused2p_kernel on CTA pairs; gemm_fused2_kernel<1024, 256> with
  // TUBER_FUSE2_SINGLE=1) and are correct, but at 8 clips a CTA owns ONE 128-row block of the stage and its chain load -> conv4 ->
  // epilogue -> second GEMM has nothing to overlap with: 74 / 70 us per bottleneck against 30 + 37 us for the two separate launches
  // (profiles/r2_fused2_layer3_experiment.json).  Off unless TUBER_FUSE2_L3=1.
  const char* nfd = getenv("TUBER_FUSE2_L3");
  p->fuse2_deep = nfd && nfd[0] == '1';
  // last conv4 of the 512-channel stage + first conv1 of the 1024-channel stage (gemm_fused2_kernel<512, 256>): measured 193 us against
  // 109 + 80 us for the separate launches (the fused kernel is tensor bound on one CTA per row block) -> only with TUBER_FUSE2_S23=1
  const char* n23 = getenv("TUBER_FUSE2_S23");
  p->no_fuse2_s23 = !(n23 && n23[0] == '1');
  const char* ndm = getenv("TUBER_NO_DEC_MEGA");             // decoder as one launch per operation (cross-check of decoder_mega.cu)
  p->no_dec_mega = ndm && ndm[0] == '1';
  const char* pu = getenv("TUBER_POOL_UNFOLDED");
  p->pool_unfolded = pu && pu[0] == '1';
  const char* ns = getenv("TUBER_NO_STRIDED_TMA");
  p->no_strided_tma = ns && ns[0] == '1';
  // Side branches on their own streams: correct (tests/test_parity_gpu.py), but measured without gain at 8 clips -- 8.96 ms per step
  // against 8.88 ms in program order (profiles/r2_overlap_experiment.json): the encoder's launches are short but 64..148 CTAs wide
  // with one CTA per SM (shared memory), so a side kernel finds no idle SM, only a turn in the same queue.  On with TUBER_OVERLAP=1.
  const char* no = getenv("TUBER_OVERLAP");
  p->no_overlap = !(no && no[0] == '1');
  const char* sc = getenv("TUBER_SIDE_CTAS");
  p->side_ctas = sc ? (atoi(sc) & ~1) : SIDE_CTAS_DEFAULT;   // even: CTA-pair kernels
  *out_plan = p;
  return TUBER_OK;
}

void tuber_plan_destroy(TuberPlan* p) {
  if (!p) return;
  for (void* d : p->owned) cudaFree(d);
  for (auto& kv : p->taps) if (kv.second.keep) cudaFree(kv.second.keep);
  for (auto& g : p->graphs) cudaGraphExecDestroy(g.exec);
  if (p->cap_stream) cudaStreamDestroy(p->cap_stream);
  for (int k = 0; k < TuberPlan::NUM_SIDE; ++k) {
    if (p->side[k]) cudaStreamDestroy(p->side[k]);
    if (p->ev_fork[k]) cudaEventDestroy(p->ev_fork[k]);
    if (p->ev_join[k]) cudaEventDestroy(p->ev_join[k]);
  }
  if (p->ws) cudaFree(p->ws);
  if (p->in_lut) cudaFree(p->in_lut);
  if (p->u8_clip) cudaFree(p->u8_clip);
  for (int i = 0; i < 2; ++i) {
    if (p->stage_in[i]) cudaFree(p->stage_in[i]);
    if (p->stage_out[i]) cudaFree(p->stage_out[i]);
    if (p->h2d_done[i]) cudaEventDestroy(p->h2d_done[i]);
    if (p->slot_done[i]) cudaEventDestroy(p->slot_done[i]);
    if (p->fwd_done[i]) cudaEventDestroy(p->fwd_done[i]);
  }
  if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
  if (p->run_stream) cudaStreamDestroy(p->run_stream);
  if (p->out_stream) cudaStreamDestroy(p->out_stream);
  if (p->ev_valid) for (auto& e : p->ev) cudaEventDestroy(e);
  for (auto& r : p->kp) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  delete p;
}

int tuber_plan_set_weight(TuberPlan* p, const char* name, const float* host_data, const int64_t* shape, int32_t ndim) {
  if (!p || !name || !host_data || ndim < 0 || (ndim > 0 && !shape)) return fail(TUBER_ERR_INVALID, "null argument");
  if (p->finalized) return fail(TUBER_ERR_STATE, "plan already finalized");
  HostTensor t;
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] < 0 || (shape[i] > 0 && n > (int64_t(1) << 40) / shape[i])) return fail(TUBER_ERR_SHAPE, "weight '%s': bad dimension %lld", name, (long long)shape[i]);
    t.shape.push_back(shape[i]);
    n *= shape[i];
  }
  t.data.assign(host_data, host_data + n);
  p->host[name] = std::move(t);
  return TUBER_OK;
}

int tuber_plan_finalize(TuberPlan* p) {
  if (!p) return fail(TUBER_ERR_INVALID, "null plan");
  if (p->finalized) return fail(TUBER_ERR_STATE, "plan already finalized");
  CK(cudaSetDevice(p->device));
  return do_finalize(p);
}

int tuber_query_shapes(TuberPlan* p, int32_t B, int32_t T, int32_t H, int32_t W, TuberShapeInfo* out) {
  if (!out) return fail(TUBER_ERR_INVALID, "null argument");
  TRY(check_forward_args(p, B, T, H, W));
  Geometry g;
  TRY(compute_geometry(p->cfg, T, H, W, g));
  size_t need = 0;
  int launches = 0;
  TRY(ensure_workspace(p, B, T, H, W, true, &need, &launches));
  out->Tf = g.Tf; out->Hf = g.Hf; out->Wf = g.Wf; out->Tp = g.Tp;
  out->enc_tokens = g.Tp * g.Hf * g.Wf; out->cls_tokens = g.Tf * g.Hf * g.Wf;
  out->launches = launches; out->workspace_bytes = (int64_t)need;
  return TUBER_OK;
}

int tuber_forward(TuberPlan* p, const float* clips_dev, const uint8_t* mask_dev, int32_t B, int32_t T, int32_t H, int32_t W,
                  float* logits_dev, float* boxes_dev, float* logits_b_dev, void* stream) {
  TRY(check_forward_args(p, B, T, H, W));
  if (!clips_dev || !logits_dev || !boxes_dev || !logits_b_dev) return fail(TUBER_ERR_INVALID, "null device pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  size_t need = 0;
  TRY(ensure_workspace(p, B, T, H, W, mask_dev != nullptr, &need, nullptr));
  if (need > p->ws_cap) {
    CK(cudaDeviceSynchronize());
    if (p->ws) cudaFree(p->ws);
    p->ws = nullptr; p->ws_cap = 0;
    p->dec_trace = nullptr;                                  // (pointed into the old workspace)
    for (auto& g : p->graphs) cudaGraphExecDestroy(g.exec);     // captured pointers are stale now
    p->graphs.clear();
    void* w = nullptr;
    cudaError_t e = cudaMalloc(&w, need);
    if (e != cudaSuccess) return fail(TUBER_ERR_CUDA, "workspace of %zu bytes: %s", need, cudaGetErrorString(e));
    p->ws = (char*)w; p->ws_cap = need;
  }
  if (p->profiling && !p->ev_valid) {
    for (auto& e : p->ev) CK(cudaEventCreate(&e));
    p->ev_valid = true;
  }
  // side branches (position codes, class-branch encoder) on their own streams, except in the modes that time or copy per launch
  const bool overlap = !p->no_overlap && !p->profiling && !p->debug_keep && !p->kprof;
  if (overlap && !p->side_valid) {
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));          // lo = least urgent (numerically largest), hi = most urgent
    for (int k = 0; k < TuberPlan::NUM_SIDE; ++k) {
      CK(cudaStreamCreateWithPriority(&p->side[k], cudaStreamNonBlocking, lo));
      CK(cudaEventCreateWithFlags(&p->ev_fork[k], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&p->ev_join[k], cudaEventDisableTiming));
    }
    p->side_valid = true;
  }
  Ctx cx{p};
  cx.dry = false;
  cx.ws.base = p->ws; cx.ws.cap = p->ws_cap;
  cx.st = st;
  cx.overlap = overlap;
  p->kp_used = 0;
  const bool graph = p->use_graph && !p->profiling && !p->debug_keep && !p->kprof;
  if (!graph) {
    int s = run_forward(cx, clips_dev, mask_dev, B, T, H, W, logits_dev, boxes_dev, logits_b_dev);
    p->launches = cx.launches;
    for (int i = 0; i < TUBER_NUM_STAGES; ++i) { p->stage_bytes[i] = cx.st_bytes[i]; p->stage_flops[i] = cx.st_flops[i]; }
    return s;
  }
  std::vector<uintptr_t> key = {(uintptr_t)clips_dev, (uintptr_t)mask_dev, (uintptr_t)B, (uintptr_t)T, (uintptr_t)H, (uintptr_t)W,
                                (uintptr_t)logits_dev, (uintptr_t)boxes_dev, (uintptr_t)logits_b_dev, (uintptr_t)p->ltc_bank,
                                (uintptr_t)p->ltc_bank_clips, (uintptr_t)p->ltc_bank_tokens, (uintptr_t)p->ltc_new, (uintptr_t)p->in_frames};
  for (auto& g : p->graphs)
    if (g.key == key) { CK(cudaGraphLaunch(g.exec, st)); return TUBER_OK; }
  // first call for this (shape, buffers): run eagerly (this call's result; also performs the one-time
  // function-attribute set-up), then record the same launch sequence for the following calls
  {
    int s0 = run_forward(cx, clips_dev, mask_dev, B, T, H, W, logits_dev, boxes_dev, logits_b_dev);
    p->launches = cx.launches;
    for (int i = 0; i < TUBER_NUM_STAGES; ++i) { p->stage_bytes[i] = cx.st_bytes[i]; p->stage_flops[i] = cx.st_flops[i]; }
    if (s0 != TUBER_OK) return s0;
  }
  Ctx cap{p};
  cap.dry = false;
  cap.ws.base = p->ws; cap.ws.cap = p->ws_cap;
  cap.st = st;
  cap.overlap = overlap;
  // record on a private stream (the caller's may be the legacy default stream, which cannot capture);
  // the instantiated graph is launched on the caller's stream.  Most urgent priority: the main chain's kernel nodes go in front of
  // the side branches' (least urgent streams) whenever both have CTAs to place.
  if (!p->cap_stream) {
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&p->cap_stream, cudaStreamNonBlocking, hi));
  }
  cap.st = p->cap_stream;
  cudaGraph_t graph_obj = nullptr;
  CK(cudaStreamBeginCapture(p->cap_stream, cudaStreamCaptureModeThreadLocal));
  int s = run_forward(cap, clips_dev, mask_dev, B, T, H, W, logits_dev, boxes_dev, logits_b_dev);
  cudaError_t e = cudaStreamEndCapture(p->cap_stream, &graph_obj);
  if (s != TUBER_OK) { if (graph_obj) cudaGraphDestroy(graph_obj); return s; }
  if (e != cudaSuccess) return fail(TUBER_ERR_CUDA, "graph capture: %s", cudaGetErrorString(e));
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph_obj, 0);
  cudaGraphDestroy(graph_obj);
  if (e != cudaSuccess) return fail(TUBER_ERR_CUDA, "graph instantiate: %s", cudaGetErrorString(e));
  if (p->graphs.size() >= 8) { cudaGraphExecDestroy(p->graphs.front().exec); p->graphs.erase(p->graphs.begin()); }
  p->graphs.push_back({key, exec});
  return TUBER_OK;
}

namespace {

// The value table of the reference's input transform: to_tensor (uint8 -> float32, / 255) followed by Normalize
// ((x - mean) / std), datasets/video_transforms.py:294-296,308-314, evaluated per channel and byte value in fp32 the way torch does
// (volatile: no contraction, no double-precision intermediates).
void build_input_lut(const float mean[3], const float stdv[3], float* lut) {
  for (int c = 0; c < 3; ++c)
    for (int u = 0; u < 256; ++u) {
      volatile float v = (float)u;
      v = v / 255.0f;
      v = v - mean[c];
      v = v / stdv[c];
      lut[c * 256 + u] = v;
    }
}

int upload_input_lut(TuberPlan* p, const float mean[3], const float stdv[3]) {
  float lut[768];
  build_input_lut(mean, stdv, lut);
  CK(cudaSetDevice(p->device));
  CK(cudaDeviceSynchronize());                                      // forwards in flight may still read the old table
  if (!p->in_lut) {
    void* q = nullptr;
    CK(cudaMalloc(&q, sizeof(lut)));
    p->in_lut = (float*)q;
  }
  CK(cudaMemcpy(p->in_lut, lut, sizeof(lut), cudaMemcpyHostToDevice));
  return TUBER_OK;
}

const float kImageNetMean[3] = {0.485f, 0.456f, 0.406f}, kImageNetStd[3] = {0.229f, 0.224f, 0.225f};   // ava_frame.py:159-162

// uint8 frames straight into the stem (stem_tc2_kernel<true>) instead of through normalize_u8_kernel; read per call (the tests switch it)
bool stem_u8_enabled() {
  const char* e = getenv("TUBER_STEM_U8");
  return e && e[0] == '1';
}

int ensure_input_lut(TuberPlan* p) { return p->in_lut ? TUBER_OK : upload_input_lut(p, kImageNetMean, kImageNetStd); }

// H2D (+mask) -> forward -> D2H of one batch through staging slot `slot`.  `in_st` carries the input copies,
// `st` the kernels and the output copies; the caller decides whether they are the same stream.
// `frames_host` (uint8 [B,T,H,W,3]) replaces `clips_host` on the uint8 path: 3 bytes per pixel cross the bus and
// normalize_u8_kernel expands them into the slot's fp32 clip buffer on `st` before the forward.
int host_step(TuberPlan* p, int slot, const float* clips_host, const uint8_t* frames_host, const uint8_t* mask_host, int B, int T, int H,
              int W, float* logits_host, float* boxes_host, float* logits_b_host, cudaStream_t in_st, cudaStream_t st,
              cudaStream_t out_st = nullptr) {
  const TuberConfig& c = p->cfg;
  const size_t clip_bytes = (size_t)B * 3 * T * H * W * 4, mask_bytes = mask_host ? (size_t)B * H * W : 0;
  const size_t mask_room = ((mask_bytes + 255) & ~(size_t)255) + 256;
  const size_t u8_bytes = frames_host ? (size_t)B * 3 * T * H * W : 0;
  const size_t in_need = clip_bytes + mask_room + u8_bytes;
  if (frames_host) TRY(ensure_input_lut(p));
  const size_t n_logits = (size_t)B * c.dec_layers * c.num_queries * c.num_classes, n_boxes = (size_t)B * c.dec_layers * c.num_queries * 4;
  const size_t n_lb = c.ava_mode ? (size_t)B * c.dec_layers * c.num_queries * 3 : (size_t)B * 2;
  const size_t out_need = (n_logits + n_boxes + n_lb) * 4 + 1024;
  if (in_need > p->stage_in_cap[slot]) {
    CK(cudaDeviceSynchronize());
    if (p->stage_in[slot]) cudaFree(p->stage_in[slot]);
    p->stage_in[slot] = nullptr; p->stage_in_cap[slot] = 0;
    void* q = nullptr;
    CK(cudaMalloc(&q, in_need));
    p->stage_in[slot] = (char*)q; p->stage_in_cap[slot] = in_need;
  }
  if (out_need > p->stage_out_cap[slot]) {
    CK(cudaDeviceSynchronize());
    if (p->stage_out[slot]) cudaFree(p->stage_out[slot]);
    p->stage_out[slot] = nullptr; p->stage_out_cap[slot] = 0;
    void* q = nullptr;
    CK(cudaMalloc(&q, out_need));
    p->stage_out[slot] = (char*)q; p->stage_out_cap[slot] = out_need;
  }
  float* d_clips = (float*)p->stage_in[slot];
  uint8_t* d_mask = mask_host ? (uint8_t*)(p->stage_in[slot] + clip_bytes) : nullptr;
  float* d_logits = (float*)p->stage_out[slot];
  float* d_boxes = d_logits + ((n_logits + 63) & ~(size_t)63);
  float* d_lb = d_boxes + ((n_boxes + 63) & ~(size_t)63);
  uint8_t* d_frames = reinterpret_cast<uint8_t*>(p->stage_in[slot] + clip_bytes + mask_room);
  if (frames_host) CK(cudaMemcpyAsync(d_frames, frames_host, u8_bytes, cudaMemcpyHostToDevice, in_st));
  else CK(cudaMemcpyAsync(d_clips, clips_host, clip_bytes, cudaMemcpyHostToDevice, in_st));
  if (mask_host) CK(cudaMemcpyAsync(d_mask, mask_host, mask_bytes, cudaMemcpyHostToDevice, in_st));
  if (in_st != st) {
    CK(cudaEventRecord(p->h2d_done[slot], in_st));
    CK(cudaStreamWaitEvent(st, p->h2d_done[slot], 0));
  }
  // uint8 path: normalize_u8_kernel expands the frames into the slot's fp32 clip (40 us, one extra launch).  With TUBER_STEM_U8=1 the
  // pair stem reads the frames itself and normalises while it stages its input rows (stem_tc2_kernel<true>: bit-identical, no fp32
  // clip, no extra launch) -- measured SLOWER: the staging threads are the stem's bottleneck and a byte load + table lookup per
  // staged pixel costs them 0.15 ms per 8-clip step (e2e 912 against 929 clips/s), so it is not the default
  const bool stem_reads_u8 = frames_host && stem_u8_enabled() && stem_pool_is_fused((W + 6 - 7) / 2 + 1);
  if (frames_host && !stem_reads_u8) CK(launch_normalize_u8(d_frames, p->in_lut, d_clips, B, (long long)T * H * W, st));
  p->in_frames = stem_reads_u8 ? d_frames : nullptr;
  const int fwd_status = tuber_forward(p, d_clips, d_mask, B, T, H, W, d_logits, d_boxes, d_lb, st);
  p->in_frames = nullptr;
  TRY(fwd_status);
  // pipelined form: the result copies run on their own stream so that the next slot's kernels are not held up by them
  cudaStream_t ost = st;
  if (out_st && out_st != st) {
    CK(cudaEventRecord(p->fwd_done[slot], st));
    CK(cudaStreamWaitEvent(out_st, p->fwd_done[slot], 0));
    ost = out_st;
  }
  CK(cudaMemcpyAsync(logits_host, d_logits, n_logits * 4, cudaMemcpyDeviceToHost, ost));
  CK(cudaMemcpyAsync(boxes_host, d_boxes, n_boxes * 4, cudaMemcpyDeviceToHost, ost));
  CK(cudaMemcpyAsync(logits_b_host, d_lb, n_lb * 4, cudaMemcpyDeviceToHost, ost));
  return TUBER_OK;
}

int ensure_host_pipeline(TuberPlan* p) {
  if (p->copy_stream) return TUBER_OK;
  CK(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&p->run_stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&p->out_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CK(cudaEventCreateWithFlags(&p->h2d_done[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&p->slot_done[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&p->fwd_done[i], cudaEventDisableTiming));
  }
  return TUBER_OK;
}

}  // namespace

int tuber_forward_host(TuberPlan* p, const float* clips_host, const uint8_t* mask_host, int32_t B, int32_t T, int32_t H,
                       int32_t W, float* logits_host, float* boxes_host, float* logits_b_host, void* stream) {
  TRY(check_forward_args(p, B, T, H, W));
  if (!clips_host || !logits_host || !boxes_host || !logits_b_host) return fail(TUBER_ERR_INVALID, "null host pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int i = 0; i < 2; ++i)
    if (p->slot_busy[i]) return fail(TUBER_ERR_STATE, "tuber_forward_host while an asynchronous submission is in flight");
  TRY(host_step(p, 0, clips_host, nullptr, mask_host, B, T, H, W, logits_host, boxes_host, logits_b_host, st, st));
  CK(cudaStreamSynchronize(st));
  return TUBER_OK;
}

int tuber_forward_host_submit(TuberPlan* p, int32_t slot, const float* clips_host, const uint8_t* mask_host, int32_t B, int32_t T,
                              int32_t H, int32_t W, float* logits_host, float* boxes_host, float* logits_b_host) {
  TRY(check_forward_args(p, B, T, H, W));
  if (slot < 0 || slot > 1) return fail(TUBER_ERR_INVALID, "slot must be 0 or 1");
  if (!clips_host || !logits_host || !boxes_host || !logits_b_host) return fail(TUBER_ERR_INVALID, "null host pointer");
  if (p->slot_busy[slot]) return fail(TUBER_ERR_STATE, "slot %d still in flight: call tuber_forward_host_wait first", slot);
  TRY(ensure_host_pipeline(p));
  TRY(host_step(p, slot, clips_host, nullptr, mask_host, B, T, H, W, logits_host, boxes_host, logits_b_host, p->copy_stream, p->run_stream,
                p->out_stream));
  CK(cudaEventRecord(p->slot_done[slot], p->out_stream));
  p->slot_busy[slot] = true;
  return TUBER_OK;
}

int tuber_forward_host_wait(TuberPlan* p, int32_t slot) {
  if (!p) return fail(TUBER_ERR_INVALID, "null plan");
  if (slot < 0 || slot > 1) return fail(TUBER_ERR_INVALID, "slot must be 0 or 1");
  if (!p->slot_busy[slot]) return fail(TUBER_ERR_STATE, "slot %d has no submission in flight", slot);
  CK(cudaEventSynchronize(p->slot_done[slot]));
  p->slot_busy[slot] = false;
  return TUBER_OK;
}

// ---- uint8 input path -------------------------------------------------------------------------
int tuber_input_lut(const float* mean, const float* stdv, float* lut_out) {
  if (!mean || !stdv || !lut_out) return fail(TUBER_ERR_INVALID, "null argument");
  for (int c = 0; c < 3; ++c)
    if (!(stdv[c] != 0.f)) return fail(TUBER_ERR_INVALID, "std[%d] is zero or NaN", c);
  build_input_lut(mean, stdv, lut_out);
  return TUBER_OK;
}

int tuber_set_input_norm(TuberPlan* p, const float* mean, const float* stdv) {
  if (!p || !mean || !stdv) return fail(TUBER_ERR_INVALID, "null argument");
  for (int c = 0; c < 3; ++c)
    if (!(stdv[c] != 0.f)) return fail(TUBER_ERR_INVALID, "std[%d] is zero or NaN", c);
  for (int i = 0; i < 2; ++i)
    if (p->slot_busy[i]) return fail(TUBER_ERR_STATE, "tuber_set_input_norm while an asynchronous submission is in flight");
  return upload_input_lut(p, mean, stdv);
}

int tuber_forward_u8(TuberPlan* p, const uint8_t* frames_dev, const uint8_t* mask_dev, int32_t B, int32_t T, int32_t H, int32_t W,
                     float* logits_dev, float* boxes_dev, float* logits_b_dev, void* stream) {
  TRY(check_forward_args(p, B, T, H, W));
  if (!frames_dev || !logits_dev || !boxes_dev || !logits_b_dev) return fail(TUBER_ERR_INVALID, "null device pointer");
  TRY(ensure_input_lut(p));
  if (stem_u8_enabled() && stem_pool_is_fused((W + 6 - 7) / 2 + 1)) {
    // TUBER_STEM_U8=1: the pair stem normalises while it stages its input rows: no fp32 clip is written, no extra launch (see host_step)
    p->in_frames = frames_dev;
    const int s = tuber_forward(p, reinterpret_cast<const float*>(frames_dev), mask_dev, B, T, H, W, logits_dev, boxes_dev, logits_b_dev, stream);
    p->in_frames = nullptr;
    return s;
  }
  const size_t need = (size_t)B * 3 * T * H * W * 4;
  if (need > p->u8_clip_cap) {
    CK(cudaDeviceSynchronize());
    if (p->u8_clip) cudaFree(p->u8_clip);
    p->u8_clip = nullptr; p->u8_clip_cap = 0;
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, need);
    if (e != cudaSuccess) return fail(TUBER_ERR_CUDA, "clip buffer of %zu bytes: %s", need, cudaGetErrorString(e));
    p->u8_clip = (float*)q; p->u8_clip_cap = need;
  }
  CK(launch_normalize_u8(frames_dev, p->in_lut, p->u8_clip, B, (long long)T * H * W, reinterpret_cast<cudaStream_t>(stream)));
  return tuber_forward(p, p->u8_clip, mask_dev, B, T, H, W, logits_dev, boxes_dev, logits_b_dev, stream);
}

// ---- long-term context bank ------------------------------------------------------------------
int tuber_has_ltc(TuberPlan* p) { return (p && p->has_ltc) ? 1 : 0; }

int tuber_forward_ltc(TuberPlan* p, const float* clips_dev, const uint8_t* mask_dev, int32_t B, int32_t T, int32_t H, int32_t W,
                      const float* bank_dev, int32_t bank_clips, int32_t bank_tokens, float* bank_new_dev, float* logits_dev,
                      float* boxes_dev, float* logits_b_dev, void* stream) {
  TRY(check_forward_args(p, B, T, H, W));
  if (bank_dev) {
    if (!p->has_ltc) return fail(TUBER_ERR_MISSING, "the plan holds no long-term context layer (ltc_attn.* / ltc_norm.* were never set)");
    if (bank_tokens < 1 || (bank_clips != 1 && bank_clips != B))
      return fail(TUBER_ERR_SHAPE, "bank of %d clips x %d tokens: clips must be 1 (shared window) or the batch size %d", bank_clips, bank_tokens, B);
  }
  p->ltc_bank = bank_dev; p->ltc_bank_clips = bank_dev ? bank_clips : 0; p->ltc_bank_tokens = bank_dev ? bank_tokens : 0;
  p->ltc_new = bank_new_dev;
  const int s = tuber_forward(p, clips_dev, mask_dev, B, T, H, W, logits_dev, boxes_dev, logits_b_dev, stream);
  p->ltc_bank = nullptr; p->ltc_bank_clips = 0; p->ltc_bank_tokens = 0; p->ltc_new = nullptr;
  return s;
}

int tuber_forward_host_u8(TuberPlan* p, const uint8_t* frames_host, const uint8_t* mask_host, int32_t B, int32_t T, int32_t H,
                          int32_t W, float* logits_host, float* boxes_host, float* logits_b_host, void* stream) {
  TRY(check_forward_args(p, B, T, H, W));
  if (!frames_host || !logits_host || !boxes_host || !logits_b_host) return fail(TUBER_ERR_INVALID, "null host pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int i = 0; i < 2; ++i)
    if (p->slot_busy[i]) return fail(TUBER_ERR_STATE, "tuber_forward_host_u8 while an asynchronous submission is in flight");
  TRY(host_step(p, 0, nullptr, frames_host, mask_host, B, T, H, W, logits_host, boxes_host, logits_b_host, st, st));
  CK(cudaStreamSynchronize(st));
  return TUBER_OK;
}

int tuber_forward_host_u8_submit(TuberPlan* p, int32_t slot, const uint8_t* frames_host, const uint8_t* mask_host, int32_t B, int32_t T,
                                 int32_t H, int32_t W, float* logits_host, float* boxes_host, float* logits_b_host) {
  TRY(check_forward_args(p, B, T, H, W));
  if (slot < 0 || slot > 1) return fail(TUBER_ERR_INVALID, "slot must be 0 or 1");
  if (!frames_host || !logits_host || !boxes_host || !logits_b_host) return fail(TUBER_ERR_INVALID, "null host pointer");
  if (p->slot_busy[slot]) return fail(TUBER_ERR_STATE, "slot %d still in flight: call tuber_forward_host_wait first", slot);
  TRY(ensure_host_pipeline(p));
  TRY(host_step(p, slot, nullptr, frames_host, mask_host, B, T, H, W, logits_host, boxes_host, logits_b_host, p->copy_stream,
                p->run_stream, p->out_stream));
  CK(cudaEventRecord(p->slot_done[slot], p->out_stream));
  p->slot_busy[slot] = true;
  return TUBER_OK;
}

int tuber_set_profiling(TuberPlan* p, int32_t enabled) {
  if (!p) return fail(TUBER_ERR_INVALID, "null plan");
  p->profiling = enabled != 0;
  return TUBER_OK;
}
int tuber_set_debug_keep(TuberPlan* p, int32_t enabled) {
  if (!p) return fail(TUBER_ERR_INVALID, "null plan");
  p->debug_keep = enabled != 0;
  return TUBER_OK;
}
int tuber_set_graph(TuberPlan* p, int32_t enabled) {
  if (!p) return fail(TUBER_ERR_INVALID, "null plan");
  p->use_graph = enabled != 0;
  return TUBER_OK;
}
int tuber_set_force_simt(TuberPlan* p, int32_t enabled) {
  if (!p) return fail(TUBER_ERR_INVALID, "null plan");
  if (p->force_simt != (enabled != 0)) {                               // recorded graphs hold the other kernel choice
    for (auto& g : p->graphs) cudaGraphExecDestroy(g.exec);
    p->graphs.clear();
  }
  p->force_simt = enabled != 0;
  return TUBER_OK;
}
int tuber_last_launches(TuberPlan* p) { return p ? p->launches : 0; }
int tuber_graph_count(TuberPlan* p) { return p ? (int)p->graphs.size() : 0; }

int tuber_set_kernel_profiling(TuberPlan* p, int32_t enabled) {
  if (!p) return fail(TUBER_ERR_INVALID, "null plan");
  p->kprof = enabled != 0;
  return TUBER_OK;
}

int tuber_get_kernel_profile(TuberPlan* p, TuberKernelStat* out, int32_t capacity, int32_t* n_out) {
  if (!p || !n_out) return fail(TUBER_ERR_INVALID, "null argument");
  std::vector<TuberKernelStat> agg;
  const char* dump = getenv("TUBER_KPROF_DUMP");
  for (int i = 0; i < p->kp_used; ++i) {
    const KernelRec& r = p->kp[i];
    CK(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, r.e0, r.e1));
    if (dump && dump[0] == '1')
      fprintf(stderr, "kprof %3d %-22s %8.1f us %8.1f GB/s %8.1f TFLOP/s  %s\n", i, r.name, ms * 1e3, r.bytes / (ms * 1e-3) / 1e9,
              r.flops / (ms * 1e-3) / 1e12, r.tag);
    TuberKernelStat* s = nullptr;
    for (auto& a : agg) if (strcmp(a.name, r.name) == 0) s = &a;
    if (!s) {
      TuberKernelStat z{};
      snprintf(z.name, sizeof z.name, "%s", r.name);
      agg.push_back(z);
      s = &agg.back();
    }
    s->launches += 1; s->ms += ms; s->bytes += r.bytes; s->flops += r.flops;
  }
  if (dump && dump[0] == '1' && p->dec_trace && p->dec_trace_n > 1) {
    std::vector<unsigned long long> t(p->dec_trace_n);
    if (cudaMemcpy(t.data(), p->dec_trace, t.size() * 8, cudaMemcpyDeviceToHost) == cudaSuccess) {
      fprintf(stderr, "decoder_mega phases (us, A | BC | D | E per layer, barriers included):");
      for (int i = 1; i < p->dec_trace_n; ++i) fprintf(stderr, "%s%.1f", (i - 1) % 4 == 0 ? "  | " : " ", (double)(t[i] - t[i - 1]) * 1e-3);
      fprintf(stderr, "\n");
    }
    std::vector<unsigned long long> t2(192);
    if (cudaMemcpy(t2.data(), p->dec_trace + 64, t2.size() * 8, cudaMemcpyDeviceToHost) == cudaSuccess) {
      fprintf(stderr, "decoder_mega CTA 0 sub-stamps (us):");
      for (int i = 1; i < 60; ++i) fprintf(stderr, " %.1f", (double)((long long)(t2[i] - t2[i - 1])) * 1e-3);
      fprintf(stderr, "\n");
    }
  }
  *n_out = (int32_t)agg.size();
  if (out) for (int i = 0; i < (int)agg.size() && i < capacity; ++i) out[i] = agg[i];
  return TUBER_OK;
}

static const char* kStageNames[TUBER_NUM_STAGES] = {"stem", "layer1", "layer2", "layer3", "layer4", "pool", "proj", "encoder", "decoder", "class+heads"};
const char* tuber_stage_name(int32_t i) { return (i >= 0 && i < TUBER_NUM_STAGES) ? kStageNames[i] : ""; }

int tuber_get_stage_ms(TuberPlan* p, float* ms_out) {
  if (!p || !ms_out) return fail(TUBER_ERR_INVALID, "null argument");
  if (!p->ev_valid) return fail(TUBER_ERR_STATE, "no profiled forward yet");
  CK(cudaEventSynchronize(p->ev[TUBER_NUM_STAGES]));
  for (int i = 0; i < TUBER_NUM_STAGES; ++i) CK(cudaEventElapsedTime(ms_out + i, p->ev[i], p->ev[i + 1]));
  return TUBER_OK;
}

int tuber_get_stage_work(TuberPlan* p, double* bytes_out, double* flops_out) {
  if (!p || !bytes_out || !flops_out) return fail(TUBER_ERR_INVALID, "null argument");
  for (int i = 0; i < TUBER_NUM_STAGES; ++i) { bytes_out[i] = p->stage_bytes[i]; flops_out[i] = p->stage_flops[i]; }
  return TUBER_OK;
}

int tuber_debug_fetch(TuberPlan* p, const char* what, float* dst_dev, int64_t* n_out, void* stream) {
  if (!p || !what) return fail(TUBER_ERR_INVALID, "null argument");
  auto it = p->taps.find(what);
  if (it == p->taps.end()) return fail(TUBER_ERR_INVALID, "no intermediate named '%s'", what);
  const Tap& t = it->second;
  if (n_out) *n_out = t.rows * t.cols;
  if (!dst_dev) return TUBER_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (t.fmt == FMT_F32) CK(cudaMemcpyAsync(dst_dev, t.ptr, (size_t)t.rows * t.cols * 4, cudaMemcpyDeviceToDevice, st));
  else CK(launch_from_split(t.ptr, t.cols, dst_dev, t.cols, t.rows, t.cols, st));
  return TUBER_OK;
}

int tuber_postprocess(TuberPlan* p, const float* logits_dev, const float* boxes_dev, const float* logits_b_dev, const float* sizes_dev,
                      int32_t B, int32_t layer, float* out_dev, void* stream) {
  if (!p || !logits_dev || !boxes_dev || !logits_b_dev || !sizes_dev || !out_dev) return fail(TUBER_ERR_INVALID, "null argument");
  const TuberConfig& c = p->cfg;
  if (B < 1 || layer < 0 || layer >= c.dec_layers) return fail(TUBER_ERR_INVALID, "bad batch %d / layer %d", B, layer);
  const long long L = c.dec_layers, Q = c.num_queries, C = c.num_classes;
  const float* lg = logits_dev + (long long)layer * Q * C;
  const float* bx = boxes_dev + (long long)layer * Q * 4;
  const float* lb = c.ava_mode ? logits_b_dev + (long long)layer * Q * 3 : logits_b_dev;
  CK(launch_postprocess(lg, L * Q * C, bx, L * Q * 4, lb, c.ava_mode ? L * Q * 3 : 2, sizes_dev, out_dev, B, (int)Q, (int)C, c.ava_mode,
                        (cudaStream_t)stream));
  return TUBER_OK;
}

// ---- single operators ------------------------------------------------------------------------
int tuber_op_to_split(const float* in, void* out, int64_t rows, int32_t cols, void* stream) {
  CK(launch_to_split(in, cols, out, cols, rows, cols, (cudaStream_t)stream));
  return TUBER_OK;
}
int tuber_op_from_split(const void* in, float* out, int64_t rows, int32_t cols, void* stream) {
  CK(launch_from_split(in, cols, out, cols, rows, cols, (cudaStream_t)stream));
  return TUBER_OK;
}
int tuber_op_pack_weight(const float* w, void* out, int32_t N, int32_t K, void* stream) {
  CK(launch_pack_weight(w, out, N, K, (cudaStream_t)stream));
  return TUBER_OK;
}
int tuber_op_gemm_tc(const void* a_split, const void* w_packed, const float* scale, const float* shift, const void* res,
                     int32_t res_fmt, int32_t res_mod, void* c, int32_t c_fmt, int32_t M, int32_t N, int32_t K, int32_t relu,
                     void* stream) {
  GemmArgs a{};
  a.A = a_split; a.a_fmt = FMT_SPLIT; a.lda = K; a.Wp = w_packed; a.scale = scale; a.shift = shift;
  a.res = res; a.res_fmt = res_fmt; a.ldr = N; a.res_mod = res_mod; a.C = c; a.c_fmt = c_fmt; a.ldc = N;
  a.M = M; a.N = N; a.K = K; a.act = relu ? ACT_RELU : ACT_NONE;
  CK(launch_gemm_tc(a, (cudaStream_t)stream));
  return TUBER_OK;
}
int tuber_op_gemm_tc_fused2(const void* a_split, const void* ab_split, const void* w_packed, const float* scale, const float* shift,
                            const void* res_split, void* c_split, int32_t M, int32_t K, int32_t Kb, const void* w2_packed,
                            const float* scale2, const float* shift2, float* c2, int32_t N1, int32_t N2, void* stream) {
  GemmArgs a{};
  a.A = a_split; a.a_fmt = FMT_SPLIT; a.lda = K; a.Ab = ab_split; a.ldb = Kb; a.Kb = ab_split ? Kb : 0;
  a.Wp = w_packed; a.scale = scale; a.shift = shift;
  a.res = res_split; a.res_fmt = FMT_SPLIT; a.ldr = N1; a.res_mod = 0; a.C = c_split; a.c_fmt = FMT_SPLIT; a.ldc = N1;
  a.M = M; a.N = N1; a.K = K; a.act = ACT_RELU;
  CK(launch_gemm_tc_fused2(a, w2_packed, scale2, shift2, c2, N2, N2, (cudaStream_t)stream));
  return TUBER_OK;
}
int tuber_op_sgemm(const float* A, const float* W, const float* bias, const float* res, float* C, int32_t M, int32_t N, int32_t K,
                   int32_t act, void* stream) {
  GemmArgs a{};
  a.A = A; a.a_fmt = FMT_F32; a.lda = K; a.Wf = W; a.shift = bias; a.res = res; a.res_fmt = FMT_F32; a.ldr = N;
  a.C = C; a.c_fmt = FMT_F32; a.ldc = N; a.M = M; a.N = N; a.K = K; a.act = act;
  CK(launch_sgemm(a, (cudaStream_t)stream));
  return TUBER_OK;
}
int tuber_op_dwconv(const float* in, const float* w27c, const float* scale, const float* shift, void* out_split, int32_t B, int32_t Ti,
                    int32_t Hi, int32_t Wi, int32_t C, int32_t stride_t, int32_t stride_s, void* stream) {
  const int To = (Ti - 1) / stride_t + 1, Ho = (Hi - 1) / stride_s + 1, Wo = (Wi - 1) / stride_s + 1;
  CK(launch_dwconv(in, w27c, scale, shift, out_split, B, Ti, Hi, Wi, C, stride_t, stride_s, To, Ho, Wo, (cudaStream_t)stream));
  return TUBER_OK;
}
int tuber_op_stem(const float* x, const float* w_oc441, const float* scale, const float* shift, float* conv_out, void* pooled_split,
                  int32_t B, int32_t T, int32_t H, int32_t W, void* stream) {
  const int H1 = (H - 1) / 2 + 1, W1 = (W - 1) / 2 + 1, H2 = (H1 - 1) / 2 + 1, W2 = (W1 - 1) / 2 + 1;
  cudaStream_t st = (cudaStream_t)stream;
  void* wp = nullptr;
  CK(cudaMalloc(&wp, 2 * 64 * 576 * 2));
  cudaError_t e = launch_stem_pack_weight(w_oc441, wp, st);
  // the unit test wants both the conv rows and the pooled tensor: run the unfused path for the rows, and, when the
  // fused path exists for this width, overwrite the pooled tensor with it (so both epilogues are exercised)
  if (e == cudaSuccess) e = launch_stem_conv(x, wp, scale, shift, conv_out, nullptr, B, T, H, W, H1, W1, st);
  if (e == cudaSuccess && pooled_split) e = launch_maxpool_hw(conv_out, pooled_split, B * T, H1, W1, H2, W2, 64, st);
  if (e == cudaSuccess && pooled_split && stem_pool_is_fused(W1))
    e = launch_stem_conv(x, wp, scale, shift, nullptr, pooled_split, B, T, H, W, H1, W1, st);
  cudaStreamSynchronize(st);
  cudaFree(wp);
  CK(e);
  return TUBER_OK;
}
int tuber_op_layernorm(const float* x, const float* res, const float* gamma, const float* beta, float* out, int64_t rows, int32_t C,
                       void* stream) {
  LnArgs a{};
  a.x = x; a.x_fmt = FMT_F32; a.ldx = C; a.res = res; a.res_fmt = FMT_F32; a.ldr = C; a.gamma = gamma; a.beta = beta; a.eps = LN_EPS;
  a.rows = (int)rows; a.C = C; a.out_f32 = out; a.ldo = C;
  CK(launch_layernorm(a, (cudaStream_t)stream));
  return TUBER_OK;
}
int tuber_op_attention(const float* q, const float* k, const float* v, const uint8_t* kpm, float* out, int32_t NB, int32_t H, int32_t L,
                       int32_t S, int32_t D, float scale, void* stream) {
  AttnArgs a{};
  const int E = H * D;
  a.q = q; a.ldq = E; a.qm = seqmap(1, L, 0, 1);
  a.k = k; a.v = v; a.ldk = E; a.ldv = E; a.km = seqmap(1, S, 0, 1);
  a.o_f32 = out; a.ldo = E; a.om = seqmap(1, L, 0, 1);
  a.kpm = kpm; a.kpm_div = 1; a.NB = NB; a.H = H; a.L = L; a.S = S; a.D = D; a.scale = scale;
  // long sequences (tcgen05 kernel): with TUBER_OP_ATTN_PREP=1 give it the conversion workspace the plan would (synchronises)
  const char* prep = getenv("TUBER_OP_ATTN_PREP");
  if (prep && prep[0] == '1' && attention_tc_supported(a)) {
    void* scratch = nullptr;
    CK(cudaMalloc(&scratch, attention_tc_scratch_bytes(a)));
    a.tc_scratch = scratch;
    cudaError_t e = launch_attention(a, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(scratch);
    if (e != cudaSuccess) return fail(TUBER_ERR_CUDA, "attention: %s", cudaGetErrorString(e));
    return TUBER_OK;
  }
  CK(launch_attention(a, (cudaStream_t)stream));
  return TUBER_OK;
}
const char* tuber_op_attention_kernel(int32_t NB, int32_t H, int32_t L, int32_t S, int32_t D, int32_t masked) {
  AttnArgs a{};
  const int E = H * D;
  a.ldq = a.ldk = a.ldv = a.ldo = E;
  a.qm = seqmap(1, L, 0, 1); a.km = seqmap(1, S, 0, 1); a.om = seqmap(1, L, 0, 1);
  a.kpm = masked ? reinterpret_cast<const uint8_t*>(uintptr_t(16)) : nullptr;       // only tested for null-ness
  a.kpm_div = 1; a.NB = NB; a.H = H; a.L = L; a.S = S; a.D = D;
  return attention_kernel_name(a);
}
int tuber_op_normalize_u8(const uint8_t* frames, const float* mean, const float* stdv, float* out, int32_t B, int64_t pixels_per_clip,
                          void* stream) {
  if (!frames || !mean || !stdv || !out || B <= 0 || pixels_per_clip <= 0) return fail(TUBER_ERR_INVALID, "bad argument");
  float lut[768];
  TRY(tuber_input_lut(mean, stdv, lut));
  void* d = nullptr;
  CK(cudaMalloc(&d, sizeof(lut)));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemcpyAsync(d, lut, sizeof(lut), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = launch_normalize_u8(frames, (const float*)d, out, B, pixels_per_clip, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) return fail(TUBER_ERR_CUDA, "normalize_u8: %s", cudaGetErrorString(e));
  return TUBER_OK;
}

int tuber_op_posenc(const uint8_t* fmask, float* pos, int32_t B, int32_t T, int32_t H, int32_t W, int32_t d_model, void* stream) {
  const int nt = d_model / 8 * 2, ns = d_model / 8 * 3;
  std::vector<float> dt(nt), ds(ns);
  for (int i = 0; i < nt; ++i) dt[i] = powf(10000.f, (float)(2 * (i / 2)) / (float)nt);
  for (int i = 0; i < ns; ++i) ds[i] = powf(10000.f, (float)(2 * (i / 2)) / (float)ns);
  float *ddt = nullptr, *dds = nullptr;
  CK(cudaMalloc((void**)&ddt, nt * 4));
  CK(cudaMalloc((void**)&dds, ns * 4));
  CK(cudaMemcpy(ddt, dt.data(), nt * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dds, ds.data(), ns * 4, cudaMemcpyHostToDevice));
  cudaError_t e = launch_posenc(fmask, ddt, dds, pos, B, T, H, W, nt, ns, (cudaStream_t)stream);
  cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(ddt);
  cudaFree(dds);
  CK(e);
  return TUBER_OK;
}

}  // extern "C"
