// TEST INFRASTRUCTURE ONLY -- host versions of the inline-PTX helpers of csrc/common.cuh, pasted in after it by
// tests/test_kernel_emulation_cpu.py (same semantics; the cache hints are meaningless on the host, the approximate
// transcendental instructions of box_muller16 become libm calls, so device-RNG noise is NOT bit-comparable here).
#pragma once
inline uint4 ld_stream_u4(const void* p) { return *static_cast<const uint4*>(p); }
inline void st_stream_u4(void* p, uint4 v) { *static_cast<uint4*>(p) = v; }
inline float4 ld_stream_f4(const void* p) { return *static_cast<const float4*>(p); }
inline void st_stream_f4(void* p, float4 v) { *static_cast<float4*>(p) = v; }
inline void box_muller16(uint32_t r, float& z0, float& z1, float scale = 1.0f) {
  const float u1 = 1.0f - (float)(r & 0xFFFFu) * (1.0f / 65536.0f), u2 = (float)(r >> 16) * (1.0f / 65536.0f);
  const float rad = sqrtf(-2.0f * logf(u1)) * scale, ang = 6.283185307179586f * u2;
  z0 = rad * cosf(ang);
  z1 = rad * sinf(ang);
}
