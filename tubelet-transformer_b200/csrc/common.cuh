// Shared definitions for the tuber_b200 CUDA library (sm_100a only).
//
// Activation storage formats
// --------------------------
//   FMT_F32   : plain fp32, row-major [rows, ld].
//   FMT_SPLIT : "split-bf16".  Every fp32 value v is stored as two bf16 numbers
//               hi = bf16_rn(v), mid = bf16_rn(v - hi)  (16 mantissa bits in total).
//               Row r of a [rows, ld] tensor occupies 2*ld bf16 = 4*ld bytes -- the same
//               bytes as fp32 -- laid out as  hi[0..ld) | mid[0..ld).
//               This is the operand format of the tcgen05 bf16x3 GEMM (gemm_tc.cu): TMA moves
//               the hi and mid planes of a tile straight into the swizzled shared-memory
//               layout the tensor core reads, and three MMAs (hi*hi + hi*mid + mid*hi) with
//               fp32 accumulation in TMEM give ~2^-16 relative operand error, which the 1e-3
//               parity bar needs (single-pass TF32, 2^-11, measurably fails it).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

enum TuberFmt { FMT_F32 = 0, FMT_SPLIT = 1 };
enum TuberAct { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2 };

#define TB_DEVINL __device__ __forceinline__

TB_DEVINL uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// v -> (hi, mid) with hi + mid == v to 16 mantissa bits
TB_DEVINL void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& mid) {
  hi = __float2bfloat16_rn(v);
  mid = __float2bfloat16_rn(v - __bfloat162float(hi));
}

TB_DEVINL float bf16_lo_to_f32(uint32_t w) { return __uint_as_float(w << 16); }
TB_DEVINL float bf16_hi_to_f32(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// (x, y) -> packed bf16x2 of the hi parts and of the mid parts (x in the low half): two packed conversions,
// same round-to-nearest results as split_bf16 on each value
TB_DEVINL void split_bf16x2(float x, float y, uint32_t& hi, uint32_t& mid) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  __nv_bfloat162 m = __floats2bfloat162_rn(x - __uint_as_float(hi << 16), y - __uint_as_float(hi & 0xffff0000u));
  mid = *reinterpret_cast<uint32_t*>(&m);
}

// 4 consecutive values -> one 8-byte store into each plane
TB_DEVINL void store_split4(__nv_bfloat16* hi_ptr, __nv_bfloat16* mid_ptr, float4 v) {
  uint2 h, m;
  split_bf16x2(v.x, v.y, h.x, m.x);
  split_bf16x2(v.z, v.w, h.y, m.y);
  *reinterpret_cast<uint2*>(hi_ptr) = h;
  *reinterpret_cast<uint2*>(mid_ptr) = m;
}

TB_DEVINL float4 load_split4(const __nv_bfloat16* hi_ptr, const __nv_bfloat16* mid_ptr) {
  uint2 h = __ldg(reinterpret_cast<const uint2*>(hi_ptr));
  uint2 m = __ldg(reinterpret_cast<const uint2*>(mid_ptr));
  float4 r;
  r.x = bf16_lo_to_f32(h.x) + bf16_lo_to_f32(m.x);
  r.y = bf16_hi_to_f32(h.x) + bf16_hi_to_f32(m.x);
  r.z = bf16_lo_to_f32(h.y) + bf16_lo_to_f32(m.y);
  r.w = bf16_hi_to_f32(h.y) + bf16_hi_to_f32(m.y);
  return r;
}

// Row pointers of a split tensor (ld in elements)
TB_DEVINL const __nv_bfloat16* split_hi(const void* base, long long row, int ld) {
  return reinterpret_cast<const __nv_bfloat16*>(base) + row * 2 * ld;
}
TB_DEVINL __nv_bfloat16* split_hi(void* base, long long row, int ld) {
  return reinterpret_cast<__nv_bfloat16*>(base) + row * 2 * ld;
}

TB_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
TB_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- per-device one-time set-up ----------------------------------------------------------------------------
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count belong to a DEVICE, not to the process: a second plan on
// another GPU of the same process needs its own opt-in (include/tuber_b200.h: "one plan per (device, config), different plans are
// independent").  DeviceOnce::run executes `setup` once per device, under a mutex (concurrent first calls from two host threads).
struct DeviceOnce {
  static constexpr int MAX_DEVICES = 64;
  std::mutex mu;
  bool done[MAX_DEVICES] = {};
  template <class F>
  cudaError_t run(F&& setup) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= MAX_DEVICES) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(mu);
    if (done[dev]) return cudaSuccess;
    e = setup();
    if (e == cudaSuccess) done[dev] = true;
    return e;
  }
};
// Optional cap on the SMs a persistent launch may take (0 = none), per host thread: set around the launches of a side branch
// (plan.cu, Ctx::side_begin) so that a kernel of the branch cannot fill the machine in front of the main chain's next launch.
// Every launcher sizes its grid from device_num_sms(), so the kernels see a consistently smaller machine.
inline int& sm_cap_ref() {
  static thread_local int cap = 0;
  return cap;
}
// SM count of the CURRENT device (cached per device), limited by the calling thread's cap
static inline int device_num_sms_uncapped();
static inline int device_num_sms() {
  const int n = device_num_sms_uncapped(), cap = sm_cap_ref();
  return cap > 0 && cap < n ? cap : n;
}
static inline int device_num_sms_uncapped() {
  static int cache[DeviceOnce::MAX_DEVICES] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= DeviceOnce::MAX_DEVICES) return 1;
  int n = cache[dev];
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n;                                  // benign race: every writer stores the same value
  }
  return n > 0 ? n : 1;
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------
// A forward is a chain of ~190 dependent launches, most of them a few microseconds long.  Kernels launched through
// launch_pdl may be scheduled while their predecessor is still running: they set up (barriers, tensor memory, descriptor
// prefetch) and then block in pdl_wait() until the predecessor grid has completed and its writes are visible.  Every
// kernel launched this way calls pdl_wait() before its first global access; pdl_trigger() at its top lets ITS successor
// start early.  (Captured into CUDA graphs as programmatic dependency edges.)
TB_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
TB_DEVINL void pdl_trigger() {
#ifdef TUBER_PDL_EARLY_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

bool tuber_pdl_enabled();   // plan.cu: off by default (measured: no gain inside CUDA graphs), TUBER_PDL=1 turns it on

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tuber_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
