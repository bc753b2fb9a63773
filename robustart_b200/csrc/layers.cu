// Memory-bound layers around the tensor-core contractions, all on "split-bf16" activations
// (planes[0:count] = hi = bf16(v), planes[count:2*count] = lo = bf16(v - hi)), NHWC.
//   split / merge                       : format conversion
//   stem im2col                         : 7x7/s2/p3 patches of the input image (resnet_official.py:221-224)
//   maxpool 3x3 s2 p1                   : resnet_official.py:227
//   global average pool                 : resnet_official.py:238
#include "common.cuh"
#include <cuda_fp16.h>

namespace {
constexpr int kThreads = 256;

// ---- fp16 single-plane helpers ------------------------------------------------------------------
__device__ __forceinline__ float2 h2_to_f2(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }
__device__ __forceinline__ uint32_t f2_to_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ uint32_t pack_u16x2(uint16_t a, uint16_t b) { return (uint32_t)a | ((uint32_t)b << 16); }

__global__ void __launch_bounds__(kThreads) split_kernel(const float4* __restrict__ in, uint2* __restrict__ hi,
                                                          uint2* __restrict__ lo, size_t count4, float scale) {
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < count4; i += (size_t)gridDim.x * kThreads) {
    float4 v = ld_stream_f4(in + i);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    uint32_t h[2], l[2];
    split_pair2(v.x, v.y, h[0], l[0]);
    split_pair2(v.z, v.w, h[1], l[1]);
    hi[i] = make_uint2(h[0], h[1]);
    lo[i] = make_uint2(l[0], l[1]);
  }
}

__global__ void __launch_bounds__(kThreads) merge_kernel(const uint2* __restrict__ hi, const uint2* __restrict__ lo,
                                                          float4* __restrict__ out, size_t count4) {
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < count4; i += (size_t)gridDim.x * kThreads) {
    uint2 h = hi[i], l = lo[i];
    float4 v;
    v.x = plane_bits_to_f32(h.x & 0xFFFF) + plane_bits_to_f32(l.x & 0xFFFF);
    v.y = plane_bits_to_f32(h.x >> 16) + plane_bits_to_f32(l.x >> 16);
    v.z = plane_bits_to_f32(h.y & 0xFFFF) + plane_bits_to_f32(l.y & 0xFFFF);
    v.w = plane_bits_to_f32(h.y >> 16) + plane_bits_to_f32(l.y >> 16);
    out[i] = v;
  }
}

// ---- stem im2col ------------------------------------------------------------------------------
// one thread per (output pixel, 8-column chunk): kpad = 192 columns = 24 chunks; column = ky*24 + kx*3 + c with
// kx in [0,7) -- every ky run is padded to 8 taps (columns ky*24+21..23 and 168..191 are zero) so that the fused
// stem's operand producer (gemm_sm100.cu) reads aligned 8-word chunks
constexpr int kStemK = 192;
struct Norm3 { float mean[3], std[3]; };

template <bool U8>
__global__ void __launch_bounds__(kThreads) stem_im2col_kernel(const void* __restrict__ img,
                                                                uint4* __restrict__ hi, uint4* __restrict__ lo,
                                                                int n, int h, int w, int ho, int wo, Norm3 nm) {
  const size_t total = (size_t)n * ho * wo * (kStemK / 8);
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int chunk = (int)(t % (kStemK / 8));
    const size_t pix = t / (kStemK / 8);
    const int ox = (int)(pix % wo);
    const int oy = (int)((pix / wo) % ho);
    const int im = (int)(pix / ((size_t)wo * ho));
    uint16_t hh[8], ll[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = chunk * 8 + j;
      float v = 0.f;
      const int ky = col / 24, rem = col - ky * 24;
      if (col < 168 && rem < 21) {
        const int kx = rem / 3, c = rem - kx * 3;
        const int iy = oy * 2 - 3 + ky, ix = ox * 2 - 3 + kx;
        if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
          float x;
          if (U8) x = __fdiv_rn((float)static_cast<const uint8_t*>(img)[(((size_t)im * h + iy) * w + ix) * 3 + c], 255.0f);
          else x = static_cast<const float*>(img)[(((size_t)im * 3 + c) * h + iy) * w + ix];
          v = (x - nm.mean[c]) / nm.std[c];
        }
      }
      split_pair(v, hh[j], ll[j]);
    }
    hi[t] = make_uint4(pack_u16x2(hh[0], hh[1]), pack_u16x2(hh[2], hh[3]), pack_u16x2(hh[4], hh[5]), pack_u16x2(hh[6], hh[7]));
    lo[t] = make_uint4(pack_u16x2(ll[0], ll[1]), pack_u16x2(ll[2], ll[3]), pack_u16x2(ll[4], ll[5]), pack_u16x2(ll[6], ll[7]));
  }
}

// ---- maxpool 3x3 s2 p1, 8 channels per thread ---------------------------------------------------
// CODES: also writes, per pooled element, the window position ky*3+kx of its (first) maximum -- 0xF where the maximum is not positive,
// i.e. with the backward of the ReLU that produced x folded in -- for b200r_maxpool3x3s2_bwd_codes_hi (the attack path's saved forward)
template <bool CODES>
__global__ void __launch_bounds__(kThreads) maxpool_kernel(const uint4* __restrict__ xhi, const uint4* __restrict__ xlo,
                                                            uint4* __restrict__ yhi, uint4* __restrict__ ylo, uint2* __restrict__ codes,
                                                            int n, int h, int w, int c8, int ho, int wo) {
  const size_t total = (size_t)n * ho * wo * c8;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int cc = (int)(t % c8);
    const size_t pix = t / c8;
    const int ox = (int)(pix % wo), oy = (int)((pix / wo) % ho), im = (int)(pix / ((size_t)wo * ho));
    float best[8];
    uint32_t bh[8], bl[8], bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bh[j] = 0xFF80u; bl[j] = 0; bi[j] = 0xFu; }
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - 1 + ky;
      if (iy < 0 || iy >= h) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - 1 + kx;
        if (ix < 0 || ix >= w) continue;
        const size_t idx = (((size_t)im * h + iy) * w + ix) * c8 + cc;
        uint4 a = __ldg(xhi + idx), b = __ldg(xlo + idx);
        uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t hb = (aw[j >> 1] >> (16 * (j & 1))) & 0xFFFF, lb = (bw[j >> 1] >> (16 * (j & 1))) & 0xFFFF;
          float v = plane_bits_to_f32((uint16_t)hb) + plane_bits_to_f32((uint16_t)lb);
          if (v > best[j]) { best[j] = v; bh[j] = hb; bl[j] = lb; if (CODES) bi[j] = (uint32_t)(ky * 3 + kx); }
        }
      }
    }
    yhi[t] = make_uint4(bh[0] | (bh[1] << 16), bh[2] | (bh[3] << 16), bh[4] | (bh[5] << 16), bh[6] | (bh[7] << 16));
    ylo[t] = make_uint4(bl[0] | (bl[1] << 16), bl[2] | (bl[3] << 16), bl[4] | (bl[5] << 16), bl[6] | (bl[7] << 16));
    if (CODES) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (!(best[j] > 0.f)) bi[j] = 0xFu;
      codes[t] = make_uint2(bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24), bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24));
    }
  }
}

// ---- fp16 twins ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) f32_to_f16_kernel(const float4* __restrict__ in, uint2* __restrict__ out, size_t count4,
                                                               float scale) {
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < count4; i += (size_t)gridDim.x * kThreads) {
    const float4 v = ld_stream_f4(in + i);
    // saturate instead of overflowing to inf (a scaled loss gradient must stay finite)
    auto c = [scale](float x) { return fminf(fmaxf(x * scale, -65504.f), 65504.f); };
    out[i] = make_uint2(f2_to_h2(c(v.x), c(v.y)), f2_to_h2(c(v.z), c(v.w)));
  }
}
__global__ void __launch_bounds__(kThreads) f16_to_f32_kernel(const uint2* __restrict__ in, float4* __restrict__ out, size_t count4,
                                                               float scale) {
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < count4; i += (size_t)gridDim.x * kThreads) {
    const uint2 h = in[i];
    const float2 a = h2_to_f2(h.x), b = h2_to_f2(h.y);
    out[i] = make_float4(a.x * scale, a.y * scale, b.x * scale, b.y * scale);
  }
}

// maxpool 3x3 s2 p1 on one fp16 plane: packed half2 maxima (HMNMX2), 8 channels per thread
__global__ void __launch_bounds__(kThreads) maxpool_f16_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int n, int h, int w,
                                                                int c8, int ho, int wo) {
  const size_t total = (size_t)n * ho * wo * c8;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int cc = (int)(t % c8);
    const size_t pix = t / c8;
    const int ox = (int)(pix % wo), oy = (int)((pix / wo) % ho), im = (int)(pix / ((size_t)wo * ho));
    const __half2 ninf = __half2half2(__ushort_as_half((unsigned short)0xFC00));
    __half2 best[4] = {ninf, ninf, ninf, ninf};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - 1 + ky;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - 1 + kx;
        if (ix < 0 || ix >= w) continue;
        const uint4 a = __ldg(x + (((size_t)im * h + iy) * w + ix) * c8 + cc);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) best[j] = __hmax2(best[j], *reinterpret_cast<const __half2*>(&aw[j]));
      }
    }
    y[t] = make_uint4(*reinterpret_cast<uint32_t*>(&best[0]), *reinterpret_cast<uint32_t*>(&best[1]),
                      *reinterpret_cast<uint32_t*>(&best[2]), *reinterpret_cast<uint32_t*>(&best[3]));
  }
}

// global average pool on one fp16 plane: one warp per (image, 8-channel chunk), fp32 sums
// Coalesced global average pool: block of 1024 threads = (image, up to 256 channel chunks); threads tile (pixel slice, chunk) so that a
// warp reads 512 contiguous bytes per plane; fixed summation order (slice partials through shared memory).  The
// warp-per-(image, chunk) kernels above read 16 bytes per 32-byte sector at a stride of C*2 bytes.
template <bool F16>
__global__ void __launch_bounds__(1024) avgpool_flat_kernel(const uint4* __restrict__ xhi, const uint4* __restrict__ xlo,
                                                            uint4* __restrict__ yhi, uint4* __restrict__ ylo, int hw, int c8) {
  __shared__ float part[8][1024];              // 1024 threads: one block per image must keep enough loads in flight by itself
  const int im = blockIdx.x, cb = blockIdx.y * 256, cw = min(256, c8 - cb);
  const int slices = 1024 / cw, tid = threadIdx.x;
  const bool active = tid < slices * cw;
  const int cc = tid % cw, sl = tid / cw;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (active) {
    const size_t base = (size_t)im * hw * c8 + cb + cc;
#pragma unroll 4
    for (int p = sl; p < hw; p += slices) {
      const size_t idx = base + (size_t)p * c8;
      const uint4 a = __ldg(xhi + idx);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
      if (F16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 v = h2_to_f2(aw[j]); s[2 * j] += v.x; s[2 * j + 1] += v.y; }
      } else {
        const uint4 b = __ldg(xlo + idx);
        const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[2 * j] += plane_lo16_f32(aw[j]) + plane_lo16_f32(bw[j]);
          s[2 * j + 1] += plane_hi16_f32(aw[j]) + plane_hi16_f32(bw[j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) part[j][tid] = s[j];
  __syncthreads();
  if (tid < cw) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = 0.f;
      for (int k = 0; k < slices; ++k) a += part[j][tid + k * cw];
      v[j] = a / (float)hw;
    }
    const size_t o = (size_t)im * c8 + cb + tid;
    if (F16) {
      yhi[o] = make_uint4(f2_to_h2(v[0], v[1]), f2_to_h2(v[2], v[3]), f2_to_h2(v[4], v[5]), f2_to_h2(v[6], v[7]));
    } else {
      uint16_t hh[8], ll[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split_pair(v[j], hh[j], ll[j]);
      yhi[o] = make_uint4(pack_u16x2(hh[0], hh[1]), pack_u16x2(hh[2], hh[3]), pack_u16x2(hh[4], hh[5]), pack_u16x2(hh[6], hh[7]));
      ylo[o] = make_uint4(pack_u16x2(ll[0], ll[1]), pack_u16x2(ll[2], ll[3]), pack_u16x2(ll[4], ll[5]), pack_u16x2(ll[6], ll[7]));
    }
  }
}

inline unsigned grid_for(size_t items) {
  size_t b = (items + kThreads - 1) / kThreads;
  size_t cap = (size_t)b200r_num_sms() * 16;
  return (unsigned)(b < cap ? (b ? b : 1) : cap);
}
}  // namespace

extern "C" {

int b200r_split_f32_scaled(const float* in, uint16_t* planes, size_t count, float scale, b200r_stream_t stream) {
  B200R_CHECK_ARG(in && planes, "null pointer");
  B200R_CHECK_ARG(count % 4 == 0, "count must be a multiple of 4");
  if (!count) return B200R_OK;
  split_kernel<<<grid_for(count / 4), kThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(in), reinterpret_cast<uint2*>(planes), reinterpret_cast<uint2*>(planes + count), count / 4, scale);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_split_f32(const float* in, uint16_t* planes, size_t count, b200r_stream_t stream) {
  return b200r_split_f32_scaled(in, planes, count, 1.0f, stream);
}

int b200r_merge_f32(const uint16_t* planes, float* out, size_t count, b200r_stream_t stream) {
  B200R_CHECK_ARG(out && planes, "null pointer");
  B200R_CHECK_ARG(count % 4 == 0, "count must be a multiple of 4");
  if (!count) return B200R_OK;
  merge_kernel<<<grid_for(count / 4), kThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint2*>(planes), reinterpret_cast<const uint2*>(planes + count), reinterpret_cast<float4*>(out), count / 4);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_f32_to_f16(const float* in, uint16_t* out, size_t count, float scale, b200r_stream_t stream) {
  B200R_CHECK_ARG(in && out, "null pointer");
  B200R_CHECK_ARG(count % 4 == 0, "count must be a multiple of 4");
  if (!count) return B200R_OK;
  f32_to_f16_kernel<<<grid_for(count / 4), kThreads, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(in),
                                                                              reinterpret_cast<uint2*>(out), count / 4, scale);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_f16_to_f32(const uint16_t* in, float* out, size_t count, float scale, b200r_stream_t stream) {
  B200R_CHECK_ARG(in && out, "null pointer");
  B200R_CHECK_ARG(count % 4 == 0, "count must be a multiple of 4");
  if (!count) return B200R_OK;
  f16_to_f32_kernel<<<grid_for(count / 4), kThreads, 0, as_stream(stream)>>>(reinterpret_cast<const uint2*>(in),
                                                                              reinterpret_cast<float4*>(out), count / 4, scale);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_maxpool3x3s2_nhwc_f16(const uint16_t* x, uint16_t* y, int n, int h, int w, int c, b200r_stream_t stream) {
  B200R_CHECK_ARG(x && y, "null pointer");
  B200R_CHECK_ARG(n > 0 && h > 0 && w > 0 && c % 8 == 0, "c must be a multiple of 8");
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  const size_t cout = (size_t)n * ho * wo * c;
  maxpool_f16_kernel<<<grid_for(cout / 8), kThreads, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y),
                                                                              n, h, w, c / 8, ho, wo);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_global_avgpool_nhwc_f16(const uint16_t* x, uint16_t* y, int n, int hw, int c, b200r_stream_t stream) {
  B200R_CHECK_ARG(x && y, "null pointer");
  B200R_CHECK_ARG(n > 0 && hw > 0 && c % 8 == 0, "c must be a multiple of 8");
  avgpool_flat_kernel<true><<<dim3(n, (c / 8 + 255) / 256), 1024, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(x), nullptr, reinterpret_cast<uint4*>(y), nullptr, hw, c / 8);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

static int stem_common(const void* img, uint16_t* planes, int n, int h, int w, const float* mean, const float* std,
                       bool u8, cudaStream_t s) {
  B200R_CHECK_ARG(img && planes && mean && std, "null pointer");
  B200R_CHECK_ARG(n > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, "bad shape");
  const int ho = h / 2, wo = w / 2;
  const size_t rows = (size_t)n * ho * wo;
  Norm3 nm;
  for (int i = 0; i < 3; ++i) { nm.mean[i] = mean[i]; nm.std[i] = std[i]; }
  uint4* hi = reinterpret_cast<uint4*>(planes);
  uint4* lo = reinterpret_cast<uint4*>(planes + rows * kStemK);
  if (u8) stem_im2col_kernel<true><<<grid_for(rows * (kStemK / 8)), kThreads, 0, s>>>(img, hi, lo, n, h, w, ho, wo, nm);
  else stem_im2col_kernel<false><<<grid_for(rows * (kStemK / 8)), kThreads, 0, s>>>(img, hi, lo, n, h, w, ho, wo, nm);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_stem_im2col_u8(const uint8_t* img, uint16_t* planes, int n, int h, int w, const float* mean_host,
                         const float* std_host, b200r_stream_t stream) {
  return stem_common(img, planes, n, h, w, mean_host, std_host, true, as_stream(stream));
}
int b200r_stem_im2col_f32(const float* img, uint16_t* planes, int n, int h, int w, const float* mean_host,
                          const float* std_host, b200r_stream_t stream) {
  return stem_common(img, planes, n, h, w, mean_host, std_host, false, as_stream(stream));
}

int b200r_maxpool3x3s2_nhwc(const uint16_t* x, uint16_t* y, int n, int h, int w, int c, b200r_stream_t stream) {
  B200R_CHECK_ARG(x && y, "null pointer");
  B200R_CHECK_ARG(n > 0 && h > 0 && w > 0 && c % 8 == 0, "c must be a multiple of 8");
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  const size_t cin = (size_t)n * h * w * c, cout = (size_t)n * ho * wo * c;
  maxpool_kernel<false><<<grid_for(cout / 8), kThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(x + cin), reinterpret_cast<uint4*>(y),
      reinterpret_cast<uint4*>(y + cout), nullptr, n, h, w, c / 8, ho, wo);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_maxpool3x3s2_nhwc_codes(const uint16_t* x, uint16_t* y, void* codes, int n, int h, int w, int c, b200r_stream_t stream) {
  B200R_CHECK_ARG(x && y && codes, "null pointer");
  B200R_CHECK_ARG(n > 0 && h > 0 && w > 0 && c % 8 == 0, "c must be a multiple of 8");
  B200R_CHECK_ARG((reinterpret_cast<uintptr_t>(codes) & 7) == 0, "codes must be 8-byte aligned");
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  const size_t cin = (size_t)n * h * w * c, cout = (size_t)n * ho * wo * c;
  maxpool_kernel<true><<<grid_for(cout / 8), kThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(x + cin), reinterpret_cast<uint4*>(y),
      reinterpret_cast<uint4*>(y + cout), reinterpret_cast<uint2*>(codes), n, h, w, c / 8, ho, wo);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_global_avgpool_nhwc(const uint16_t* x, uint16_t* y, int n, int hw, int c, b200r_stream_t stream) {
  B200R_CHECK_ARG(x && y, "null pointer");
  B200R_CHECK_ARG(n > 0 && hw > 0 && c % 8 == 0, "c must be a multiple of 8");
  const size_t cin = (size_t)n * hw * c, cout = (size_t)n * c;
  avgpool_flat_kernel<false><<<dim3(n, (c / 8 + 255) / 256), 1024, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(x + cin), reinterpret_cast<uint4*>(y),
      reinterpret_cast<uint4*>(y + cout), hw, c / 8);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

}  // extern "C"
