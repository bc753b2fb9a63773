// Shared device/host helpers for libb200robust (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>
#include "../../include/b200r.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb200robust is written for sm_100a (B200) only"
#endif

// ---------------------------------------------------------------------------------------------
// error plumbing (C-ABI: return codes + thread-local message)
// ---------------------------------------------------------------------------------------------
void b200r_set_error(const char* fmt, ...);

#define B200R_CHECK_ARG(cond, ...)                                                    \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      b200r_set_error(__VA_ARGS__);                                                   \
      return B200R_EINVAL;                                                            \
    }                                                                                 \
  } while (0)

#define B200R_CUDA(call)                                                              \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      b200r_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return B200R_ECUDA;                                                             \
    }                                                                                 \
  } while (0)

#define B200R_LAUNCH_CHECK() B200R_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(b200r_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int b200r_num_sms();  // cached SM count of the current device

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: no state, any element addressable.
//   key     = 64-bit seed
//   counter = (c0, c1, c2, c3)
// ---------------------------------------------------------------------------------------------
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

// R rounds: 10 is the Random123 default; 7 is the smallest count the authors report as Crush-resistant
// (Salmon et al., SC'11, table 2) and is what the per-byte noise kernels use (2 IMAD.WIDE + 2 LOP3 per round).
template <int R>
__host__ __device__ __forceinline__ uint4 philox4x32(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(PHILOX_M0, c.x), lo0 = PHILOX_M0 * c.x;
    uint32_t hi1 = __umulhi(PHILOX_M1, c.z), lo1 = PHILOX_M1 * c.z;
#else
    uint64_t p0 = (uint64_t)PHILOX_M0 * c.x, p1 = (uint64_t)PHILOX_M1 * c.z;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += PHILOX_W0;
    k1 += PHILOX_W1;
  }
  return c;
}
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) { return philox4x32<10>(c, k0, k1); }

// uniform in [0,1) from the top 24 bits (exactly representable, never 1.0)
__device__ __forceinline__ float u32_to_unit(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }
// uniform in (0,1] from 16 bits: (k+1)/65536
__device__ __forceinline__ float u16_to_unit_open0(uint32_t r16) { return (float)(r16 + 1u) * (1.0f / 65536.0f); }

// Box-Muller on two 16-bit uniforms -> two N(0,1) (MUFU lg2/sqrt/sin/cos; 2 MUFU per normal)
// u1 = (k+1)/65536 in (0,1], u2 = k/65536 in [0,1): built with bit tricks (mantissa injection) so the
// only XU-pipe work is the four MUFUs (lg2, sqrt, sin, cos).  `scale` multiplies both normals.
__device__ __forceinline__ void box_muller16(uint32_t r, float& z0, float& z1, float scale = 1.0f) {
  // [1,2) floats with the 16 random bits in the top of the mantissa
  const float a = __uint_as_float(0x3F800000u | ((r & 0xFFFFu) << 7));        // 1 + k1/65536
  const float b = __uint_as_float(0x3F800000u | ((r >> 16) << 7));            // 1 + k2/65536
  const float u1 = 2.0f - a;                                                  // 1 - k1/65536 in (0,1]
  float lg;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u1));
  float rad;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(lg * -1.3862943611198906f));  // sqrt(-2 ln u1)
  rad *= scale;
  const float ang = fmaf(b, 6.283185307179586f, -6.283185307179586f);         // 2*pi*u2
  float s, c;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(ang));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(ang));
  z0 = rad * c;
  z1 = rad * s;
}

// ---------------------------------------------------------------------------------------------
// streaming 128-bit global access (inputs are read once, outputs written once)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 ld_stream_f4(const void* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f4(void* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// byte k (0..3) of a 32-bit word as float, exact
__device__ __forceinline__ float byte_to_float(uint32_t word, int k) {
  return (float)((word >> (8 * k)) & 0xFFu);
}
// v in [0,1] -> uint8 trunc(v*255)  (np.uint8() truncation, imagenet_c/__init__.py:35).
// fused multiply-add with round-toward-zero onto 2^23 leaves trunc(v*255) in the low mantissa bits.
__device__ __forceinline__ uint32_t unit_to_u8(float v01) {
  return __float_as_uint(__fmaf_rz(v01, 255.0f, 8388608.0f)) & 0xFFu;
}
__device__ __forceinline__ float clamp01(float v) { return __saturatef(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// split-plane helpers: v ~= hi + lo, both IEEE fp16 (round-to-nearest-even): 11 + 11 significant bits.  (Round 1 used bf16
// pairs, 8 + 8 bits: 1.4e-3 max logit error on ResNet-50 at realistic logit magnitude -- over the 1e-3 bar; fp16 pairs cost the
// same bytes and the same three MMAs per product and leave ~2^-22 per stored value.)  Range: |v| < 65504; hi rounds to 0 below
// 3e-8, so gradients are carried loss-scaled (nets.GRAD_SCALE).
__device__ __forceinline__ uint16_t f32_to_plane_bits(float v) { return __half_as_ushort(__float2half_rn(v)); }
__device__ __forceinline__ float plane_bits_to_f32(uint16_t b) { return __half2float(__ushort_as_half(b)); }
// the two fp16 values packed in one 32-bit word of a plane (element 2j in the low half, 2j + 1 in the high half)
__device__ __forceinline__ float plane_lo16_f32(uint32_t w) { return __half2float(__ushort_as_half((uint16_t)(w & 0xFFFFu))); }
__device__ __forceinline__ float plane_hi16_f32(uint32_t w) { return __half2float(__ushort_as_half((uint16_t)(w >> 16))); }
// two values at once, packed in plane words (v0 in the low half): the packed conversions (F2FP.F16.F32.PACK_AB on the ALU pipe,
// HADD2.F32 back) instead of four scalar F2F on the quarter-rate XU pipe -- same round-to-nearest results bit for bit
__device__ __forceinline__ void split_pair2(float v0, float v1, uint32_t& hw, uint32_t& lw) {
  const __half2 h = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
  hw = *reinterpret_cast<const uint32_t*>(&h);
  lw = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void split_pair(float v, uint16_t& hi, uint16_t& lo) {
  const __half h = __float2half_rn(v);
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(__float2half_rn(v - __half2float(h)));
}
