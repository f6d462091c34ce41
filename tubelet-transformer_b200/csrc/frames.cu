// Frame loading for the TubeR forward path: what the reference's loader does per frame (datasets/ava_frame.py:146-150)
//
//     tmp = Image.open(video_frame_list[frame_idx])                          # baseline JPEG decode (Pillow -> libjpeg-turbo defaults)
//     tmp = tmp.resize((target['orig_size'][1], target['orig_size'][0]))     # Pillow's default filter for RGB images: BICUBIC
//
// with everything after the entropy decoder on the GPU, bit-identical to Pillow (tests/test_frames_gpu.py against
// oracle/frame_oracle.py and tests/golden/frames.npz):
//
//   host, one thread per frame      marker parsing + Huffman decoding (T.81 F.2.2) -> quantised coefficients, int16, natural
//                                   order, straight into pinned memory                                  [HuffmanDecoder below]
//   idct_islow_kernel               dequantisation + libjpeg's "islow" integer inverse DCT (jidctint.c), one thread per 8x8 block
//   upsample_color_kernel           libjpeg's "fancy" (triangle) chroma upsampling h2v1 / h2v2 (jdsample.c; the rows above the first and
//                                   below the last are their copies, jdmainct.c) + fixed-point YCbCr -> RGB (jdcolor.c)
//   resample_kernel x 2             Pillow's ImagingResample for 8-bit pixels (Resample.c): horizontal then vertical pass, each
//                                   rounded and clipped to 8 bits; the 22-bit fixed-point filter weights are computed on the host
//                                   in double precision with Pillow's own operation order (precompute_coeffs + normalize_coeffs_8bpc)
//
// Output: RGB uint8 [n, out_h, out_w, 3] on the device = the input of tuber_forward_u8 (ToTensor + Normalize run there).
// Supported: baseline sequential (SOF0), 8 bits, three components YCbCr with luma sampling 1x1 / 2x1 / 2x2 and 1x1 chroma, one
// interleaved scan, restart intervals.  Anything else is refused with TUBER_ERR_INVALID (no fallback).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/tuber_b200.h"
#include "kernels.h"

namespace frames {

// ---------------------------------------------------------------------------------------------------------------
// host: JPEG markers + Huffman decoding
// ---------------------------------------------------------------------------------------------------------------
static const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
  bool present = false;
  uint8_t vals[256];
  int32_t maxcode[18];       // largest code of each length (-1: none), maxcode[17] = sentinel
  int32_t valoff[17];        // vals index of the first code of a length minus that code
  uint16_t look[512];        // 9-bit lookahead: (length << 8) | symbol, 0 = longer code
  int16_t fast_ac[512];      // AC tables: code + magnitude bits fit in 9 bits -> (value << 8) | (run << 4) | total bits, 0 = take the long way
  void build(const uint8_t* counts, const uint8_t* symbols, int nsym) {
    memcpy(vals, symbols, nsym);
    memset(look, 0, sizeof look);
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
      valoff[len] = k - code;
      for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code)
        if (len <= 9) {
          const int first = code << (9 - len);
          for (int j = 0; j < (1 << (9 - len)); ++j) look[first + j] = (uint16_t)((len << 8) | symbols[k]);
        }
      maxcode[len] = counts[len - 1] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
    for (int i = 0; i < 512; ++i) {
      fast_ac[i] = 0;
      const uint16_t e = look[i];
      if (!e) continue;
      const int len = e >> 8, rs = e & 0xFF, run = rs >> 4, mag = rs & 15;
      if (mag == 0 || len + mag > 9) continue;
      int v = (i >> (9 - len - mag)) & ((1 << mag) - 1);          // the magnitude bits that follow the code
      if (v < (1 << (mag - 1))) v += -(1 << mag) + 1;              // EXTEND
      if (v >= -128 && v <= 127) fast_ac[i] = (int16_t)((v * 256) + (run * 16) + (len + mag));
    }
    present = true;
  }
};

struct Component { int id, h, v, tq, td, ta; };

struct Header {
  int width = 0, height = 0, ncomp = 0, ri = 0;
  Component comp[3];
  uint16_t qt[4][64];        // natural order
  bool qt_present[4] = {false, false, false, false};
  HuffTable dc[4], ac[4];
  const uint8_t* scan = nullptr;
  const uint8_t* end = nullptr;
  int hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
};

static inline int be16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

// -> nullptr on success, else a message
static const char* parse_header(const uint8_t* data, int64_t size, Header& h) {
  if (size < 4 || data[0] != 0xFF || data[1] != 0xD8) return "not a JPEG (no SOI)";
  int64_t pos = 2;
  bool sof = false;
  while (pos + 4 <= size) {
    if (data[pos] != 0xFF) return "marker expected";
    const int m = data[pos + 1];
    pos += 2;
    if (m == 0xFF) { pos -= 1; continue; }
    if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
    if (m == 0xD9) return "EOI before a scan";
    const int len = be16(data + pos);
    if (len < 2 || pos + len > size) return "truncated segment";
    const uint8_t* seg = data + pos + 2;
    const int n = len - 2;
    if (m == 0xDB) {
      int i = 0;
      while (i < n) {
        const int pq = seg[i] >> 4, tq = seg[i] & 15;
        ++i;
        if (tq > 3 || i + (pq ? 128 : 64) > n) return "bad DQT";
        for (int k = 0; k < 64; ++k) h.qt[tq][kZigzag[k]] = pq ? (uint16_t)be16(seg + i + 2 * k) : seg[i + k];
        i += pq ? 128 : 64;
        h.qt_present[tq] = true;
      }
    } else if (m == 0xC0) {
      if (n < 6 || seg[0] != 8) return "only 8-bit samples are supported";
      h.height = be16(seg + 1); h.width = be16(seg + 3); h.ncomp = seg[5];
      if (h.ncomp != 3 || n < 6 + 9) return "only three-component (YCbCr) JPEGs are supported";
      if (h.width < 1 || h.height < 1) return "empty image";
      for (int c = 0; c < 3; ++c) {
        h.comp[c].id = seg[6 + 3 * c]; h.comp[c].h = seg[7 + 3 * c] >> 4; h.comp[c].v = seg[7 + 3 * c] & 15; h.comp[c].tq = seg[8 + 3 * c] & 3;
      }
      sof = true;
    } else if (m == 0xC1 || m == 0xC2 || m == 0xC3 || (m >= 0xC5 && m <= 0xC7) || (m >= 0xC9 && m <= 0xCB) || (m >= 0xCD && m <= 0xCF)) {
      return "only baseline sequential (SOF0) JPEGs are supported";
    } else if (m == 0xC4) {
      int i = 0;
      while (i < n) {
        if (i + 17 > n) return "bad DHT";
        const int tc = seg[i] >> 4, th = seg[i] & 15;
        int nsym = 0;
        for (int k = 0; k < 16; ++k) nsym += seg[i + 1 + k];
        if (th > 3 || tc > 1 || nsym > 256 || i + 17 + nsym > n) return "bad DHT";
        (tc ? h.ac[th] : h.dc[th]).build(seg + i + 1, seg + i + 17, nsym);
        i += 17 + nsym;
      }
    } else if (m == 0xDD) {
      if (n < 2) return "bad DRI";
      h.ri = be16(seg);
    } else if (m == 0xDA) {
      if (!sof) return "SOS before SOF";
      if (n < 1 || seg[0] != 3 || n < 1 + 6 + 3) return "only one interleaved three-component scan is supported";
      for (int c = 0; c < 3; ++c) {
        if (seg[1 + 2 * c] != h.comp[c].id) return "scan component order differs from the frame header";
        h.comp[c].td = seg[2 + 2 * c] >> 4; h.comp[c].ta = seg[2 + 2 * c] & 15;
        if (h.comp[c].td > 3 || h.comp[c].ta > 3 || !h.dc[h.comp[c].td].present || !h.ac[h.comp[c].ta].present) return "missing Huffman table";
        if (!h.qt_present[h.comp[c].tq]) return "missing quantisation table";
      }
      h.scan = data + pos + len;
      h.end = data + size;
      for (int c = 0; c < 3; ++c) { h.hmax = h.comp[c].h > h.hmax ? h.comp[c].h : h.hmax; h.vmax = h.comp[c].v > h.vmax ? h.comp[c].v : h.vmax; }
      if (h.comp[1].h != 1 || h.comp[1].v != 1 || h.comp[2].h != 1 || h.comp[2].v != 1 || h.comp[0].h != h.hmax || h.comp[0].v != h.vmax ||
          !((h.hmax == 1 && h.vmax == 1) || (h.hmax == 2 && h.vmax == 1) || (h.hmax == 2 && h.vmax == 2)))
        return "unsupported chroma subsampling (supported: 4:4:4, 4:2:2, 4:2:0)";
      h.mcux = (h.width + 8 * h.hmax - 1) / (8 * h.hmax);
      h.mcuy = (h.height + 8 * h.vmax - 1) / (8 * h.vmax);
      return nullptr;
    }
    pos += len;
  }
  return "no scan found";
}

// MSB-first bit reader over the entropy-coded segment: removes 0xFF00 stuffing, stops at markers (then feeds zeros)
struct BitReader {
  const uint8_t* p; const uint8_t* end;
  uint64_t acc = 0; int n = 0; bool at_marker = false;
  int fake = 0;                                                 // zero bits fed after the data ended (marker or end of file): the tail of acc
  inline bool overrun() const { return n < fake; }              // bits beyond the real data were consumed: truncated / corrupt scan
  inline void fill() {
    if (n > 32) return;                                         // a symbol (<= 16 bits) + its value bits (<= 15) always fit
    if (!at_marker && p + 8 <= end) {
      // fast path: the next bytes hold no 0xFF (no stuffing, no marker): append as many whole bytes as fit
      uint64_t v;
      memcpy(&v, p, 8);
      v = __builtin_bswap64(v);
      const uint64_t inv = ~v;
      if (((inv - 0x0101010101010101ull) & ~inv & 0x8080808080808080ull) == 0) {      // no byte of v is 0xFF
        const int k = (64 - n) >> 3;
        acc |= (k == 8 ? v : (v >> (64 - 8 * k)) << (64 - 8 * k - n));
        p += k; n += 8 * k;
        return;
      }
    }
    while (n <= 56) {
      uint32_t b = 0;
      if (at_marker || p >= end) fake += 8;
      if (!at_marker && p < end) {
        b = *p;
        if (b == 0xFF) {
          if (p + 1 < end && p[1] == 0x00) { p += 2; }
          else { at_marker = true; b = 0; fake += 8; }
        } else {
          ++p;
        }
      }
      acc |= (uint64_t)b << (56 - n);
      n += 8;
    }
  }
  inline uint32_t peek(int k) { return (uint32_t)(acc >> (64 - k)); }
  inline void skip(int k) { acc <<= k; n -= k; }
  inline uint32_t get(int k) { const uint32_t v = peek(k); skip(k); return v; }
  // after an interval: drop the partial byte, consume the RSTn marker
  bool restart() {
    acc = 0; n = 0; fake = 0;
    while (p + 1 < end && p[0] == 0xFF && p[1] == 0xFF) ++p;
    if (p + 1 < end && p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7) { p += 2; at_marker = false; return true; }
    return false;
  }
};

static inline int decode_symbol(BitReader& br, const HuffTable& t) {
  br.fill();
  const uint16_t e = t.look[br.peek(9)];
  if (e) { br.skip(e >> 8); return e & 0xFF; }
  int code = (int)br.peek(10), len = 10;
  while (code > t.maxcode[len]) { ++len; if (len > 16) return -1; code = (int)br.peek(len); }
  br.skip(len);
  return t.vals[(code + t.valoff[len]) & 0xFF];
}
static inline int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }

// coefficients of one frame -> out (int16, natural order; every block is zeroed before it is filled): [Y blocks | Cb blocks | Cr blocks], each
// component as [blocks_y][blocks_x][64] over whole MCUs
static const char* decode_scan(const Header& h, int16_t* out) {
  int bx[3], by[3];
  int64_t off[3], o = 0;
  for (int c = 0; c < 3; ++c) { bx[c] = h.mcux * h.comp[c].h; by[c] = h.mcuy * h.comp[c].v; off[c] = o; o += (int64_t)bx[c] * by[c] * 64; }
  BitReader br{h.scan, h.end};
  int pred[3] = {0, 0, 0};
  const int total = h.mcux * h.mcuy;
  int until_restart = h.ri ? h.ri : total;
  for (int mcu = 0; mcu < total; ++mcu) {
    if (until_restart == 0) {
      if (!br.restart()) return "restart marker missing";
      pred[0] = pred[1] = pred[2] = 0;
      until_restart = h.ri;
    }
    --until_restart;
    const int my = mcu / h.mcux, mx = mcu - my * h.mcux;
    for (int c = 0; c < 3; ++c) {
      const HuffTable& dct = h.dc[h.comp[c].td];
      const HuffTable& act = h.ac[h.comp[c].ta];
      for (int v = 0; v < h.comp[c].v; ++v)
        for (int u = 0; u < h.comp[c].h; ++u) {
          int16_t* blk = out + off[c] + ((int64_t)(my * h.comp[c].v + v) * bx[c] + (mx * h.comp[c].h + u)) * 64;
          memset(blk, 0, 128);                                   // (zeroed right before it is filled: the block stays in L1)
          int t = decode_symbol(br, dct);
          if (t < 0 || t > 15) return "bad DC code";
          if (t) pred[c] += extend((int)br.get(t), t);           // (>= 17 bits are left after a symbol: no refill needed)
          blk[0] = (int16_t)pred[c];
          for (int k = 1; k < 64;) {
            br.fill();
            const int16_t fa = act.fast_ac[br.peek(9)];
            if (fa) {                                             // short code + small value: one table look-up
              k += (fa >> 4) & 15;
              if (k > 63) return "coefficient index out of range";
              br.skip(fa & 15);
              blk[kZigzag[k++]] = (int16_t)(fa >> 8);
              continue;
            }
            const int rs = decode_symbol(br, act);
            if (rs < 0) return "bad AC code";
            const int r = rs >> 4, s = rs & 15;
            if (s == 0) {
              if (r == 15) { k += 16; continue; }
              break;
            }
            k += r;
            if (k > 63) return "coefficient index out of range";
            blk[kZigzag[k]] = (int16_t)extend((int)br.get(s), s);
            ++k;
          }
        }
    }
    if (br.overrun()) return "truncated scan (the data ends before the last MCU)";
  }
  return nullptr;
}

// ---------------------------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------------------------
struct FrameMeta {
  int W, H;                      // image size
  int hmax, vmax;                // luma sampling factors (chroma is 1x1)
  int bx[3], by[3];              // blocks per component (whole MCUs)
  int dw[3], dh[3];              // real (downsampled) samples per component
  long long coef_off[3];         // int16 elements from the coefficient base
  long long plane_off[3];        // bytes from the plane base; plane c is [by*8][bx*8] uint8
  long long rgb_off;             // bytes from the decoded-image base ([H][W][3])
  long long tmp_off;             // bytes from the horizontal-pass base ([H][out_w][3])
  int hk_off, hk_ksize, hb_off;  // horizontal filter: weights at kk + hk_off ([out_w][ksize]), bounds at bounds + hb_off ([out_w][2]); ksize 0 = no pass
  int vk_off, vk_ksize, vb_off;
  unsigned short qt[3][64];
};

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// one 1-D pass of jpeg_idct_islow on d[0..7] (13-bit fixed-point constants), un-descaled outputs in o[0..7]
__device__ __forceinline__ void idct8(const int (&d)[8], int (&o)[8]) {
  int z2 = d[2], z3 = d[6];
  int z1 = (z2 + z3) * 4433;
  int tmp2 = z1 + z3 * (-15137);
  int tmp3 = z1 + z2 * 6270;
  z2 = d[0]; z3 = d[4];
  int tmp0 = (z2 + z3) << 13;
  int tmp1 = (z2 - z3) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = d[7]; tmp1 = d[5]; tmp2 = d[3]; tmp3 = d[1];
  z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * 9633;
  tmp0 *= 2446; tmp1 *= 16819; tmp2 *= 25172; tmp3 *= 12299;
  z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
  z3 += z5; z4 += z5;
  tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
  o[0] = tmp10 + tmp3; o[7] = tmp10 - tmp3;
  o[1] = tmp11 + tmp2; o[6] = tmp11 - tmp2;
  o[2] = tmp12 + tmp1; o[5] = tmp12 - tmp1;
  o[3] = tmp13 + tmp0; o[4] = tmp13 - tmp0;
}
// libjpeg's range-limit table applied to x + 128, indexed modulo 1024 like `& RANGE_MASK`
__device__ __forceinline__ unsigned range_limit(int x) {
  const int i = x & 1023;
  return (unsigned)(i < 128 ? i + 128 : (i < 512 ? 255 : (i < 896 ? 0 : i - 896)));
}

// dequantisation + islow inverse DCT: thread = one 8x8 block; grid.y = frame
__global__ void __launch_bounds__(128) idct_islow_kernel(const FrameMeta* __restrict__ metas, const short* __restrict__ coefs,
                                                         unsigned char* __restrict__ planes) {
  const FrameMeta& m = metas[blockIdx.y];
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  int c = 0;
  for (; c < 3; ++c) {
    const int nb = m.bx[c] * m.by[c];
    if (b < nb) break;
    b -= nb;
  }
  if (c == 3) return;
  const short* src = coefs + m.coef_off[c] + (long long)b * 64;
  int ws[64];
  // pass 1: columns, scaled up by 2^PASS1_BITS
#pragma unroll
  for (int col = 0; col < 8; ++col) {
    int d[8], o[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) d[r] = (int)src[r * 8 + col] * (int)m.qt[c][r * 8 + col];
    idct8(d, o);
#pragma unroll
    for (int r = 0; r < 8; ++r) ws[r * 8 + col] = descale(o[r], 13 - 2);
  }
  // pass 2: rows -> samples
  const int byi = b / m.bx[c], bxi = b - byi * m.bx[c];
  unsigned char* dst = planes + m.plane_off[c] + ((long long)byi * 8) * (m.bx[c] * 8) + bxi * 8;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int d[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = ws[r * 8 + k];
    idct8(d, o);
    unsigned lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) lo |= range_limit(descale(o[k], 13 + 2 + 3)) << (8 * k);
#pragma unroll
    for (int k = 0; k < 4; ++k) hi |= range_limit(descale(o[4 + k], 13 + 2 + 3)) << (8 * k);
    *reinterpret_cast<uint2*>(dst + (long long)r * (m.bx[c] * 8)) = make_uint2(lo, hi);
  }
}

// one chroma sample of the full-resolution image at (x, y): libjpeg's fancy upsampling of plane p (stride ps, dw x dh real samples)
__device__ __forceinline__ int chroma_at(const unsigned char* p, int ps, int dw, int dh, int hmax, int vmax, int x, int y) {
  if (hmax == 1) return p[(long long)y * ps + x];
  const int cx = x >> 1;
  if (vmax == 1) {                                                   // h2v1_fancy_upsample
    const unsigned char* row = p + (long long)y * ps;
    if (dw <= 2) return row[cx];                                     // (plain replication when the row is too short for the filter)
    const int v = row[cx];
    if (x & 1) return cx == dw - 1 ? v : (v * 3 + row[cx + 1] + 2) >> 2;
    return cx == 0 ? v : (v * 3 + row[cx - 1] + 1) >> 2;
  }
  const int cy = y >> 1;                                             // h2v2_fancy_upsample
  if (dw <= 2) return p[(long long)cy * ps + cx];
  int cyn = (y & 1) ? cy + 1 : cy - 1;                               // the further row: below for odd output rows, above for even ones
  cyn = cyn < 0 ? 0 : (cyn > dh - 1 ? dh - 1 : cyn);                 // (edge rows are duplicated, jdmainct.c)
  const unsigned char* r0 = p + (long long)cy * ps;
  const unsigned char* r1 = p + (long long)cyn * ps;
  const int cs = r0[cx] * 3 + r1[cx];
  if (x & 1) {
    if (cx == dw - 1) return (cs * 4 + 7) >> 4;
    return (cs * 3 + (r0[cx + 1] * 3 + r1[cx + 1]) + 7) >> 4;
  }
  if (cx == 0) return (cs * 4 + 8) >> 4;
  return (cs * 3 + (r0[cx - 1] * 3 + r1[cx - 1]) + 8) >> 4;
}

__device__ __forceinline__ unsigned clamp255(int v) { return (unsigned)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// fancy upsampling + YCbCr -> RGB (jdcolor.c: 16-bit fixed point): thread = one pixel; grid.y = frame
__global__ void __launch_bounds__(256) upsample_color_kernel(const FrameMeta* __restrict__ metas, const unsigned char* __restrict__ planes,
                                                             unsigned char* __restrict__ rgb) {
  const FrameMeta& m = metas[blockIdx.y];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)m.W * m.H) return;
  const int y = (int)(i / m.W), x = (int)(i - (long long)y * m.W);
  const int Y = planes[m.plane_off[0] + (long long)y * (m.bx[0] * 8) + x];
  const int cb = chroma_at(planes + m.plane_off[1], m.bx[1] * 8, m.dw[1], m.dh[1], m.hmax, m.vmax, x, y) - 128;
  const int cr = chroma_at(planes + m.plane_off[2], m.bx[2] * 8, m.dw[2], m.dh[2], m.hmax, m.vmax, x, y) - 128;
  const int r = Y + ((91881 * cr + 32768) >> 16);                    // FIX(1.40200)
  const int g = Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16);      // FIX(0.34414), FIX(0.71414)
  const int b = Y + ((116130 * cb + 32768) >> 16);                   // FIX(1.77200)
  unsigned char* o = rgb + m.rgb_off + i * 3;
  o[0] = (unsigned char)clamp255(r); o[1] = (unsigned char)clamp255(g); o[2] = (unsigned char)clamp255(b);
}

// one pass of Pillow's ImagingResample (8 bits per channel).  HORIZONTAL: src [H][W][3] -> dst [H][out_w][3]; else src [H][out_w][3]
// (or the decoded image when there was no horizontal pass) -> dst [out_h][out_w][3].  thread = one output pixel; grid.y = frame
template <bool HORIZONTAL>
__global__ void __launch_bounds__(256) resample_kernel(const FrameMeta* __restrict__ metas, const unsigned char* __restrict__ src_base,
                                                       unsigned char* __restrict__ dst_base, const int* __restrict__ kk,
                                                       const int* __restrict__ bounds, int out_h, int out_w, long long dst_frame_bytes,
                                                       int src_is_rgb) {
  const FrameMeta& m = metas[blockIdx.y];
  const int rows = HORIZONTAL ? m.H : out_h;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * out_w) return;
  const int y = (int)(i / out_w), x = (int)(i - (long long)y * out_w);
  const int ksize = HORIZONTAL ? m.hk_ksize : m.vk_ksize;
  unsigned char* dst = dst_base + (HORIZONTAL ? m.tmp_off : (long long)blockIdx.y * dst_frame_bytes) + i * 3;
  // the source of this pass: the decoded image, or the horizontal pass's result
  const int sw = HORIZONTAL ? m.W : (src_is_rgb ? m.W : out_w);
  const unsigned char* src = src_base + ((HORIZONTAL || src_is_rgb) ? m.rgb_off : m.tmp_off);
  if (ksize == 0) {                                                  // this axis keeps its size: copy
    const unsigned char* s = src + ((long long)y * sw + x) * 3;
    dst[0] = s[0]; dst[1] = s[1]; dst[2] = s[2];
    return;
  }
  const int o = HORIZONTAL ? x : y;
  const int* k = kk + (HORIZONTAL ? m.hk_off : m.vk_off) + (long long)o * ksize;
  const int* bd = bounds + (HORIZONTAL ? m.hb_off : m.vb_off) + 2 * o;
  const int first = bd[0], taps = bd[1];
  int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;                      // 1 << (PRECISION_BITS - 1)
  if (HORIZONTAL) {
    const unsigned char* s = src + ((long long)y * sw + first) * 3;
    for (int t = 0; t < taps; ++t) { const int w = k[t]; s0 += s[3 * t] * w; s1 += s[3 * t + 1] * w; s2 += s[3 * t + 2] * w; }
  } else {
    const unsigned char* s = src + ((long long)first * sw + x) * 3;
    for (int t = 0; t < taps; ++t) { const int w = k[t]; const unsigned char* q = s + (long long)t * sw * 3; s0 += q[0] * w; s1 += q[1] * w; s2 += q[2] * w; }
  }
  dst[0] = (unsigned char)clamp255(s0 >> 22); dst[1] = (unsigned char)clamp255(s1 >> 22); dst[2] = (unsigned char)clamp255(s2 >> 22);
}

// ---------------------------------------------------------------------------------------------------------------
// Pillow's filter weights (Resample.c: precompute_coeffs with the bicubic filter, normalize_coeffs_8bpc), double precision,
// the same operation order
// ---------------------------------------------------------------------------------------------------------------
static double bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}
struct Coeffs { int ksize; std::vector<int> kk, bounds; };
static Coeffs precompute(int in_size, int out_size) {
  Coeffs c;
  double scale = (double)in_size / out_size, filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  c.ksize = (int)ceil(support) * 2 + 1;
  c.kk.assign((size_t)out_size * c.ksize, 0);
  c.bounds.assign((size_t)out_size * 2, 0);
  std::vector<double> k(c.ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0 + (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) { const double w = bicubic((x + xmin - center + 0.5) * ss); k[x] = w; ww += w; }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      c.kk[(size_t)xx * c.ksize + x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << 22)) : (int)(0.5 + k[x] * (1 << 22));
    }
    c.bounds[2 * xx] = xmin; c.bounds[2 * xx + 1] = xmax;
  }
  return c;
}

}  // namespace frames

// ---------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------
struct TuberFrameDecoder {
  int threads = 1;
  // pinned host staging + device buffers, grown on demand
  // (the pinned staging exists twice: call i+1 decodes into one set while the copies of call i still read the other)
  void* h_coef_s[2] = {nullptr, nullptr}; size_t h_coef_cap_s[2] = {0, 0};
  void* h_meta_s[2] = {nullptr, nullptr}; size_t h_meta_cap_s[2] = {0, 0};
  void* h_tab_s[2] = {nullptr, nullptr}; size_t h_tab_cap_s[2] = {0, 0};
  int slot = 0;
  void* d_coef = nullptr; size_t d_coef_cap = 0;
  void* d_meta = nullptr; size_t d_meta_cap = 0;
  void* d_tab = nullptr; size_t d_tab_cap = 0;
  void* d_planes = nullptr; size_t d_planes_cap = 0;
  void* d_rgb = nullptr; size_t d_rgb_cap = 0;
  void* d_tmp = nullptr; size_t d_tmp_cap = 0;
  cudaEvent_t done_s[2] = {nullptr, nullptr}; // the copies out of a slot's pinned buffers have finished
  bool pending_s[2] = {false, false};
  std::map<std::pair<int, int>, frames::Coeffs> cache;
  char err[256] = "";
};

namespace {
thread_local char g_frames_error[256] = "";
int ffail(int code, const char* msg) { snprintf(g_frames_error, sizeof g_frames_error, "%s", msg); return code; }
bool grow(void** p, size_t* cap, size_t need, bool pinned) {
  if (need <= *cap) return true;
  if (*p) { if (pinned) cudaFreeHost(*p); else cudaFree(*p); }
  *p = nullptr; *cap = 0;
  need += need / 4 + 4096;
  if ((pinned ? cudaMallocHost(p, need) : cudaMalloc(p, need)) != cudaSuccess) return false;
  *cap = need;
  return true;
}
}  // namespace

extern "C" {

const char* tuber_frames_last_error(void) { return g_frames_error; }

int tuber_frames_create(TuberFrameDecoder** out, int32_t max_threads) {
  if (!out) return ffail(TUBER_ERR_INVALID, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return ffail(TUBER_ERR_CUDA, "no CUDA device: the frame decoder has no CPU fallback");
  TuberFrameDecoder* d = new TuberFrameDecoder();
  int hw = (int)std::thread::hardware_concurrency();
  if (hw < 1) hw = 1;
  d->threads = max_threads > 0 ? (max_threads < hw ? max_threads : hw) : hw;
  for (int i = 0; i < 2; ++i)
    if (cudaEventCreateWithFlags(&d->done_s[i], cudaEventDisableTiming) != cudaSuccess) { delete d; return ffail(TUBER_ERR_CUDA, "cudaEventCreate failed"); }
  *out = d;
  return TUBER_OK;
}

void tuber_frames_destroy(TuberFrameDecoder* d) {
  if (!d) return;
  for (int i = 0; i < 2; ++i) {
    if (d->pending_s[i]) cudaEventSynchronize(d->done_s[i]);
    if (d->h_coef_s[i]) cudaFreeHost(d->h_coef_s[i]);
    if (d->h_meta_s[i]) cudaFreeHost(d->h_meta_s[i]);
    if (d->h_tab_s[i]) cudaFreeHost(d->h_tab_s[i]);
    if (d->done_s[i]) cudaEventDestroy(d->done_s[i]);
  }
  for (void* p : {d->d_coef, d->d_meta, d->d_tab, d->d_planes, d->d_rgb, d->d_tmp})
    if (p) cudaFree(p);
  delete d;
}

// host only: the entropy decoder by itself (unit tests; needs no device)
int tuber_op_jpeg_coefficients(const uint8_t* jpeg, int64_t size, int16_t* coef_out, int64_t capacity, int32_t* info_out) {
  using namespace frames;
  if (!jpeg || !info_out) return ffail(TUBER_ERR_INVALID, "null argument");
  Header h;
  const char* msg = parse_header(jpeg, size, h);
  if (msg) return ffail(TUBER_ERR_INVALID, msg);
  int64_t total = 0;
  info_out[0] = h.width; info_out[1] = h.height; info_out[2] = h.hmax; info_out[3] = h.vmax;
  for (int c = 0; c < 3; ++c) {
    info_out[4 + 2 * c] = h.mcux * h.comp[c].h; info_out[5 + 2 * c] = h.mcuy * h.comp[c].v;
    total += (int64_t)info_out[4 + 2 * c] * info_out[5 + 2 * c] * 64;
  }
  if (!coef_out) return TUBER_OK;
  if (capacity < total) return ffail(TUBER_ERR_SHAPE, "coefficient buffer too small");
  msg = decode_scan(h, coef_out);
  if (msg) return ffail(TUBER_ERR_INVALID, msg);
  return TUBER_OK;
}

int tuber_frames_decode(TuberFrameDecoder* d, const uint8_t* const* jpeg_ptrs, const int64_t* jpeg_sizes, int32_t n, int32_t out_h,
                        int32_t out_w, uint8_t* frames_dev, void* stream) {
  using namespace frames;
  if (!d || !jpeg_ptrs || !jpeg_sizes || !frames_dev) return ffail(TUBER_ERR_INVALID, "null argument");
  if (n < 1 || out_h < 1 || out_w < 1) return ffail(TUBER_ERR_SHAPE, "bad frame count or output size");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // ---- headers (serial: cheap) and the layout of every buffer ----
  std::vector<Header> hdr(n);
  std::vector<FrameMeta> metas(n);
  std::vector<int> tab;                                             // filter weights and bounds of this call's (in, out) size pairs
  std::map<std::pair<int, int>, std::pair<int, int>> tab_pos;      // (in, out) -> (weights offset, bounds offset)
  long long coef_total = 0, plane_total = 0, rgb_total = 0, tmp_total = 0;
  int max_blocks = 0;
  long long max_pixels = 0, max_hpix = 0;
  auto filter_for = [&](int in_size, int out_size, int& k_off, int& ksize, int& b_off) {
    if (in_size == out_size) { k_off = 0; ksize = 0; b_off = 0; return; }
    const auto key = std::make_pair(in_size, out_size);
    auto it = d->cache.find(key);
    if (it == d->cache.end()) it = d->cache.emplace(key, precompute(in_size, out_size)).first;
    auto pos = tab_pos.find(key);
    if (pos == tab_pos.end()) {
      const int ko = (int)tab.size();
      tab.insert(tab.end(), it->second.kk.begin(), it->second.kk.end());
      const int bo = (int)tab.size();
      tab.insert(tab.end(), it->second.bounds.begin(), it->second.bounds.end());
      pos = tab_pos.emplace(key, std::make_pair(ko, bo)).first;
    }
    k_off = pos->second.first; b_off = pos->second.second; ksize = it->second.ksize;
  };
  // headers (markers, Huffman table construction) on the host threads too: ~14 us per frame adds up over a 256-frame call
  std::vector<const char*> errs(n, nullptr);
  const int T = d->threads < n ? d->threads : n;
  auto run_threads = [&](auto&& fn) {
    if (T <= 1) { fn(0); return; }
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) pool.emplace_back(fn, t);
    fn(0);
    for (auto& th : pool) th.join();
  };
  run_threads([&](int t) {
    for (int i = t; i < n; i += T) errs[i] = parse_header(jpeg_ptrs[i], jpeg_sizes[i], hdr[i]);
  });
  for (int i = 0; i < n; ++i) {
    if (errs[i]) {
      snprintf(g_frames_error, sizeof g_frames_error, "frame %d: %s", i, errs[i]);
      return TUBER_ERR_INVALID;
    }
    const Header& h = hdr[i];
    FrameMeta& m = metas[i];
    m.W = h.width; m.H = h.height; m.hmax = h.hmax; m.vmax = h.vmax;
    int blocks = 0;
    for (int c = 0; c < 3; ++c) {
      m.bx[c] = h.mcux * h.comp[c].h; m.by[c] = h.mcuy * h.comp[c].v;
      m.dw[c] = (h.width * h.comp[c].h + h.hmax - 1) / h.hmax; m.dh[c] = (h.height * h.comp[c].v + h.vmax - 1) / h.vmax;
      m.coef_off[c] = coef_total; coef_total += (long long)m.bx[c] * m.by[c] * 64;
      m.plane_off[c] = plane_total; plane_total += (long long)m.bx[c] * m.by[c] * 64;
      blocks += m.bx[c] * m.by[c];
      for (int k = 0; k < 64; ++k) m.qt[c][k] = h.qt[h.comp[c].tq][k];
    }
    plane_total = (plane_total + 15) & ~15LL;
    m.rgb_off = rgb_total; rgb_total += ((long long)m.W * m.H * 3 + 15) & ~15LL;
    m.tmp_off = tmp_total; tmp_total += ((long long)m.H * out_w * 3 + 15) & ~15LL;
    filter_for(m.W, out_w, m.hk_off, m.hk_ksize, m.hb_off);
    filter_for(m.H, out_h, m.vk_off, m.vk_ksize, m.vb_off);
    if (blocks > max_blocks) max_blocks = blocks;
    if ((long long)m.W * m.H > max_pixels) max_pixels = (long long)m.W * m.H;
    if ((long long)m.H * out_w > max_hpix) max_hpix = (long long)m.H * out_w;
  }
  if (tab.empty()) tab.push_back(0);
  // ---- buffers (the pinned ones may still be read by the previous call's copies) ----
  const int sl = d->slot;
  d->slot ^= 1;
  if (d->pending_s[sl]) {
    if (cudaEventSynchronize(d->done_s[sl]) != cudaSuccess) return ffail(TUBER_ERR_CUDA, "a previous decode failed");
    d->pending_s[sl] = false;
  }
  void*& h_coef = d->h_coef_s[sl];
  void*& h_meta = d->h_meta_s[sl];
  void*& h_tab = d->h_tab_s[sl];
  if (!grow(&h_coef, &d->h_coef_cap_s[sl], (size_t)coef_total * 2, true) || !grow(&h_meta, &d->h_meta_cap_s[sl], metas.size() * sizeof(FrameMeta), true) ||
      !grow(&h_tab, &d->h_tab_cap_s[sl], tab.size() * 4, true) || !grow(&d->d_coef, &d->d_coef_cap, (size_t)coef_total * 2, false) ||
      !grow(&d->d_meta, &d->d_meta_cap, metas.size() * sizeof(FrameMeta), false) || !grow(&d->d_tab, &d->d_tab_cap, tab.size() * 4, false) ||
      !grow(&d->d_planes, &d->d_planes_cap, (size_t)plane_total, false) || !grow(&d->d_rgb, &d->d_rgb_cap, (size_t)rgb_total, false) ||
      !grow(&d->d_tmp, &d->d_tmp_cap, (size_t)tmp_total, false))
    return ffail(TUBER_ERR_CUDA, "frame decoder: out of memory");
  // ---- entropy decoding: host threads, frame i on thread i % T, straight into pinned memory ----
  int16_t* hc = reinterpret_cast<int16_t*>(h_coef);
  run_threads([&](int t) {
    for (int i = t; i < n; i += T) {
      int16_t* dst = hc + metas[i].coef_off[0];
      long long cnt = 0;
      for (int c = 0; c < 3; ++c) cnt += (long long)metas[i].bx[c] * metas[i].by[c] * 64;
      (void)cnt;
      errs[i] = decode_scan(hdr[i], dst);                           // (zeroes every block it fills)
    }
  });
  for (int i = 0; i < n; ++i)
    if (errs[i]) {
      snprintf(g_frames_error, sizeof g_frames_error, "frame %d: %s", i, errs[i]);
      return TUBER_ERR_INVALID;
    }
  memcpy(h_meta, metas.data(), metas.size() * sizeof(FrameMeta));
  memcpy(h_tab, tab.data(), tab.size() * 4);
  // ---- device: copies, inverse DCT, upsampling + colour, resize ----
#define FCK(expr)                                                                                         \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess) {                                                                              \
      snprintf(g_frames_error, sizeof g_frames_error, "%s -> %s", #expr, cudaGetErrorString(_e));       \
      return TUBER_ERR_CUDA;                                                                              \
    }                                                                                                     \
  } while (0)
  FCK(cudaMemcpyAsync(d->d_coef, h_coef, (size_t)coef_total * 2, cudaMemcpyHostToDevice, st));
  FCK(cudaMemcpyAsync(d->d_meta, h_meta, metas.size() * sizeof(FrameMeta), cudaMemcpyHostToDevice, st));
  FCK(cudaMemcpyAsync(d->d_tab, h_tab, tab.size() * 4, cudaMemcpyHostToDevice, st));
  FCK(cudaEventRecord(d->done_s[sl], st));
  d->pending_s[sl] = true;
  const FrameMeta* dm = reinterpret_cast<const FrameMeta*>(d->d_meta);
  const int* dt = reinterpret_cast<const int*>(d->d_tab);
  unsigned char* planes = reinterpret_cast<unsigned char*>(d->d_planes);
  unsigned char* rgb = reinterpret_cast<unsigned char*>(d->d_rgb);
  unsigned char* tmp = reinterpret_cast<unsigned char*>(d->d_tmp);
  idct_islow_kernel<<<dim3((max_blocks + 127) / 128, n), 128, 0, st>>>(dm, reinterpret_cast<const short*>(d->d_coef), planes);
  upsample_color_kernel<<<dim3((unsigned)((max_pixels + 255) / 256), n), 256, 0, st>>>(dm, planes, rgb);
  bool any_h = false;
  for (int i = 0; i < n; ++i) any_h = any_h || metas[i].hk_ksize != 0;
  // frames of one call share the pass structure only when their widths agree with out_w alike; keep it simple: run the horizontal
  // pass for all frames whenever one of them needs it (a frame that does not is copied by the pass)
  if (any_h)
    resample_kernel<true><<<dim3((unsigned)((max_hpix + 255) / 256), n), 256, 0, st>>>(dm, rgb, tmp, dt, dt, out_h, out_w, 0, 0);
  resample_kernel<false><<<dim3((unsigned)(((long long)out_h * out_w + 255) / 256), n), 256, 0, st>>>(
      dm, any_h ? tmp : rgb, frames_dev, dt, dt, out_h, out_w, (long long)out_h * out_w * 3, any_h ? 0 : 1);
  FCK(cudaGetLastError());
#undef FCK
  return TUBER_OK;
}

}  // extern "C"
