// Internal interface between api.cu and the corruption translation units.
#pragma once
#include "common.cuh"

struct CorruptArgs {
  int id, severity;          // severity in 1..5
  const uint8_t* in;
  uint8_t* out;
  int n, h, w;
  uint64_t seed, image_offset;
  const float* ext;          // nullable
  void* ws;
  size_t ws_bytes;
  cudaStream_t stream;
};

// per-family entry points (return B200R_* codes)
int corrupt_pixel_family(const CorruptArgs& a);    // 0,1,2,8,9,10,11,15,18
int corrupt_stencil_family(const CorruptArgs& a);  // 3,4,5,6,7,12,16,17
int corrupt_codec_family(const CorruptArgs& a);    // 13,14

size_t corrupt_pixel_ws(int id, int sev, int n, int h, int w);
size_t corrupt_stencil_ws(int id, int sev, int n, int h, int w);
size_t corrupt_codec_ws(int id, int sev, int n, int h, int w);

size_t corrupt_ext_count(int id, int sev, int n, int h, int w);

// RNG stream tags (second Philox counter word, high half) so that different uses never overlap
enum RngStream : uint32_t {
  RNG_GAUSS = 1, RNG_SHOT = 2, RNG_IMPULSE = 3, RNG_SPECKLE = 4, RNG_FROST = 5, RNG_FOG = 6,
  RNG_GLASS = 7, RNG_MOTION = 8, RNG_SNOW = 9, RNG_ELASTIC = 10, RNG_SPATTER = 11, RNG_PGD = 12
};

__host__ __device__ __forceinline__ uint4 rng_counter(uint32_t idx, uint32_t stream_tag, uint32_t sub,
                                                      uint64_t image) {
  return make_uint4(idx, (stream_tag << 16) | (sub & 0xFFFFu), (uint32_t)image, (uint32_t)(image >> 32));
}
