// Layers of the mobile families (MobileNetV2, EfficientNet-B0) that are NOT dense contractions, on split-bf16
// planes NHWC.  All HBM-bound, CUDA cores, fp32 math.
//   depthwise conv k x k (3 or 5), stride 1/2, pad k/2 + folded BN + ReLU6 / swish
//        mobilenet_v2.py:31-47,65 (ConvBNReLU groups=hidden); efficientnet.py:322-336
//   squeeze-excite channel scaling  out = x * w[n, c]      efficientnet.py:352-355
//   small-K im2col of the input image for the 3x3/s2 stems (mobilenet_v2.py:130, efficientnet.py:429-433)
#include "common.cuh"
#include <stdlib.h>

namespace {
constexpr int kThreads = 256;

__device__ __forceinline__ void unpack8(uint4 h, uint4 l, float* v) {
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = plane_lo16_f32(hw[j]) + plane_lo16_f32(lw[j]);
    v[2 * j + 1] = plane_hi16_f32(hw[j]) + plane_hi16_f32(lw[j]);
  }
}
__device__ __forceinline__ void pack8(const float* v, uint4& h, uint4& l) {
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_pair2(v[2 * j], v[2 * j + 1], hw[j], lw[j]);
  h = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  l = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}
// 256-bit global store (sm_100): one full 32-byte sector per thread
__device__ __forceinline__ void st_v8(uint4* p, const uint4& a, const uint4& b) {
#ifdef __CUDA_ARCH__
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z),
               "r"(b.w)
               : "memory");
#else            // host build of the kernel sources (tests/emu)
  p[0] = a;
  p[1] = b;
#endif
}
__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case B200R_ACT_RELU: return fmaxf(v, 0.f);
    case B200R_ACT_RELU6: return fminf(fmaxf(v, 0.f), 6.f);
    case B200R_ACT_SWISH: return v / (1.f + expf(-v));
    case B200R_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// thread = (output pixel, 8 channels); weights [k*k][c] float32 (tap major: coalesced across channels)
__global__ void __launch_bounds__(kThreads) dwconv_kernel(const uint4* __restrict__ xh, const uint4* __restrict__ xl,
                                                           const float* __restrict__ wgt, const float* __restrict__ scale,
                                                           const float* __restrict__ bias, uint4* __restrict__ yh,
                                                           uint4* __restrict__ yl, int n, int h, int w, int c8, int k, int stride,
                                                           int pad, int ho, int wo, int act) {
  const size_t total = (size_t)n * ho * wo * c8;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int cc = (int)(t % c8);
    const size_t pix = t / c8;
    const int ox = (int)(pix % wo), oy = (int)((pix / wo) % ho), im = (int)(pix / ((size_t)wo * ho));
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int ky = 0; ky < k; ++ky) {
      const int iy = oy * stride - pad + ky;
      if (iy < 0 || iy >= h) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ix = ox * stride - pad + kx;
        if (ix < 0 || ix >= w) continue;
        const size_t idx = (((size_t)im * h + iy) * w + ix) * c8 + cc;
        float v[8];
        unpack8(__ldg(xh + idx), __ldg(xl + idx), v);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wgt + ((size_t)(ky * k + kx) * c8 + cc) * 8));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wgt + ((size_t)(ky * k + kx) * c8 + cc) * 8) + 1);
        acc[0] = fmaf(v[0], w0.x, acc[0]); acc[1] = fmaf(v[1], w0.y, acc[1]); acc[2] = fmaf(v[2], w0.z, acc[2]); acc[3] = fmaf(v[3], w0.w, acc[3]);
        acc[4] = fmaf(v[4], w1.x, acc[4]); acc[5] = fmaf(v[5], w1.y, acc[5]); acc[6] = fmaf(v[6], w1.z, acc[6]); acc[7] = fmaf(v[7], w1.w, acc[7]);
      }
    }
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = act_apply(fmaf(acc[j], scale[cc * 8 + j], bias[cc * 8 + j]), act);
    uint4 hh, ll;
    pack8(o, hh, ll);
    yh[t] = hh; yl[t] = ll;
  }
}

// Strip version for the shapes the mobile families use (k = 3 / 5, stride 1 / 2): thread = (8 channels, 4 consecutive
// output columns of one output row).  The generic kernel re-reads (and re-unpacks) every input pixel k*k times through L1:
// 576 B of L1 traffic and ~330 instructions per 64-byte output, 3.3x its HBM roofline.  Here a pixel is loaded and
// unpacked once per kernel row and applied to every (output, tap) pair it belongs to; the row's K weight vectors sit in
// registers.  L1 traffic per output drops 4x (k = 3, s = 1), the unpack work likewise.
template <int K, int S>
__global__ void __launch_bounds__(kThreads) dwconv_strip_kernel(const uint4* __restrict__ xh, const uint4* __restrict__ xl,
                                                                 const float* __restrict__ wgt, const float* __restrict__ scale,
                                                                 const float* __restrict__ bias, uint4* __restrict__ yh,
                                                                 uint4* __restrict__ yl, int n, int h, int w, int c8, int ho, int wo,
                                                                 int act) {
  constexpr int OW = 4, PAD = K / 2, IW = (OW - 1) * S + K;
  const int strips = (wo + OW - 1) / OW;
  const size_t total = (size_t)n * ho * strips * c8;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int cc = (int)(t % c8);
    size_t r = t / c8;
    const int sx = (int)(r % strips);
    r /= strips;
    const int oy = (int)(r % ho), im = (int)(r / ho);
    const int ox0 = sx * OW, ix0 = ox0 * S - PAD;
    float acc[OW][8];
#pragma unroll
    for (int o = 0; o < OW; ++o)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[o][j] = 0.f;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const int iy = oy * S - PAD + ky;
      if (iy < 0 || iy >= h) continue;
      float wk[K][8];
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const float4* wp = reinterpret_cast<const float4*>(wgt + ((size_t)(ky * K + kx) * c8 + cc) * 8);
        const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
        wk[kx][0] = w0.x; wk[kx][1] = w0.y; wk[kx][2] = w0.z; wk[kx][3] = w0.w;
        wk[kx][4] = w1.x; wk[kx][5] = w1.y; wk[kx][6] = w1.z; wk[kx][7] = w1.w;
      }
      const size_t rowbase = ((size_t)im * h + iy) * w * c8 + cc;
      // all loads of the row first (clamped address, zero outside the image): IW x 2 independent 16-byte loads in flight
      uint4 rh[IW], rl[IW];
#pragma unroll
      for (int j = 0; j < IW; ++j) {
        const int ix = ix0 + j;
        const bool in = ix >= 0 && ix < w;
        const size_t at = rowbase + (size_t)(in ? ix : 0) * c8;
        rh[j] = __ldg(xh + at);
        rl[j] = __ldg(xl + at);
        if (!in) { rh[j] = make_uint4(0, 0, 0, 0); rl[j] = make_uint4(0, 0, 0, 0); }
      }
#pragma unroll
      for (int j = 0; j < IW; ++j) {
        float v[8];
        unpack8(rh[j], rl[j], v);
#pragma unroll
        for (int o = 0; o < OW; ++o) {
          const int kx = j - o * S;                  // compile-time after unrolling
          if (kx >= 0 && kx < K) {
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[o][q] = fmaf(v[q], wk[kx][q], acc[o][q]);
          }
        }
      }
    }
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + cc * 8)), s1 = __ldg(reinterpret_cast<const float4*>(scale + cc * 8) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + cc * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bias + cc * 8) + 1);
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, bi[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const size_t obase = (((size_t)im * ho + oy) * wo + ox0) * c8 + cc;
#pragma unroll
    for (int o = 0; o < OW; ++o) {
      if (ox0 + o >= wo) break;
      float ov[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float y = fmaf(acc[o][q], sc[q], bi[q]);
        ov[q] = act == B200R_ACT_RELU6 ? fminf(fmaxf(y, 0.f), 6.f)
              : act == B200R_ACT_SWISH ? __fdividef(y, 1.f + __expf(-y))
              : act_apply(y, act);
      }
      uint4 hh, ll;
      pack8(ov, hh, ll);
      yh[obase + (size_t)o * c8] = hh;
      yl[obase + (size_t)o * c8] = ll;
    }
  }
}

// Tiled version (round 2).  ncu on the strip kernel (MobileNetV2, batch 256): 104 registers -> 16 resident warps, every pixel loaded
// AND unpacked (hi/lo fp16 -> fp32: ~45 instructions per 8 channels) 4.5 times through L1, 600 instructions per 64-byte output, issue
// 60 %, DRAM 31-37 %: 2.2x the layers' HBM time.  Here a CTA owns a tile of TH x TW output pixels x CG channel groups (8 channels
// each): the input patch is loaded from global memory ONCE, unpacked ONCE and parked in shared memory as fp32
// [half][group][row][column] float4 planes (plane size = 16 bytes mod 128, so the 4 groups x 2 strips of a quarter-warp hit eight
// different 16-byte bank groups).  A thread then computes a strip of OW outputs along x: per kernel row the K weight vectors (two
// broadcast LDS.128 each) and the row's (OW-1)*S+K pixels are read once and every pixel feeds all the (output, tap) pairs it belongs to.
// Threads map to (group fastest, strip, row): four neighbours store 64 contiguous bytes per plane.
template <int K, int S, int OW>
__global__ void __launch_bounds__(kThreads) dwconv_tile_kernel(const uint4* __restrict__ xh, const uint4* __restrict__ xl,
                                                                const float* __restrict__ wgt, const float* __restrict__ scale,
                                                                const float* __restrict__ bias, uint4* __restrict__ yh,
                                                                uint4* __restrict__ yl, int h, int w, int c8, int ho, int wo, int act,
                                                                int TH, int TW, int CG, int tiles_x, int tiles_y, int tiles_c,
                                                                int plane4 /* float4 per (half, group) plane, padded */) {
  extern __shared__ __align__(16) float4 dw_smem[];
  constexpr int PAD = K / 2, IW = (OW - 1) * S + K;
  const int THI = (TH - 1) * S + K, TWI = (TW - 1) * S + K;
  float4* s_w = dw_smem;                                   // [K*K][CG][2]
  float4* s_x = dw_smem + K * K * CG * 2;                  // [2][CG][THI][TWI] (+ padding)
  const int half4 = CG * plane4;
  int t = blockIdx.x;
  const int tc = t % tiles_c; t /= tiles_c;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int im = t / tiles_y;
  const int cg0 = tc * CG, oy0 = ty * TH, ox0 = tx * TW, iy0 = oy0 * S - PAD, ix0 = ox0 * S - PAD;
  for (int i = threadIdx.x; i < K * K * CG * 2; i += kThreads) {
    const int hf = i & 1, g = (i >> 1) % CG, tap = (i >> 1) / CG;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cg0 + g < c8) v = __ldg(reinterpret_cast<const float4*>(wgt + ((size_t)tap * c8 + cg0 + g) * 8) + hf);
    s_w[i] = v;
  }
  // input patch: one pass.  Work item = (group, patch column, chunk of four rows) -- group fastest: 64 contiguous bytes per plane and
  // pixel -- with the item's eight loads in flight before the first conversion; three divisions per item, not per pixel.
  const int cols = TWI * CG, chunks = (THI + 3) >> 2;
  for (int e = threadIdx.x; e < cols * chunks; e += kThreads) {
    const int chunk = e / cols, rem = e - chunk * cols;
    const int px = rem / CG, g = rem - px * CG;
    const int ix = ix0 + px, py = chunk * 4;
    const bool col_ok = ix >= 0 && ix < w && cg0 + g < c8;
    const size_t col = (size_t)im * h * w * c8 + (size_t)(col_ok ? ix : 0) * c8 + (col_ok ? cg0 + g : 0);
    const size_t row_pitch = (size_t)w * c8;
    float4* d = s_x + g * plane4 + px;
    uint4 rh[4], rl[4];
    bool ok[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int iy = iy0 + py + r;
      ok[r] = col_ok && py + r < THI && iy >= 0 && iy < h;
      const size_t at = col + (size_t)(ok[r] ? iy : 0) * row_pitch;
      rh[r] = __ldg(xh + at);
      rl[r] = __ldg(xl + at);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (py + r >= THI) break;
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (ok[r]) unpack8(rh[r], rl[r], v);
      d[(py + r) * TWI] = make_float4(v[0], v[1], v[2], v[3]);
      d[half4 + (py + r) * TWI] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
  __syncthreads();
  const int SW = TW / OW;                                  // the host picks TW as a multiple of OW
  for (int it = threadIdx.x; it < TH * SW * CG; it += kThreads) {
    const int g = it % CG, sx = (it / CG) % SW, y = it / (CG * SW);
    const int oy = oy0 + y, oxs = ox0 + sx * OW, cc = cg0 + g;
    if (oy >= ho || oxs >= wo || cc >= c8) continue;
    float a[OW][8];
#pragma unroll
    for (int o = 0; o < OW; ++o)
#pragma unroll
      for (int q = 0; q < 8; ++q) a[o][q] = 0.f;
    const float4* row0 = s_x + g * plane4 + (y * S) * TWI + sx * OW * S;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      float4 w0[K], w1[K];
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        w0[kx] = s_w[((ky * K + kx) * CG + g) * 2];
        w1[kx] = s_w[((ky * K + kx) * CG + g) * 2 + 1];
      }
      const float4* rp = row0 + ky * TWI;
#pragma unroll
      for (int j = 0; j < IW; ++j) {
        const float4 v0 = rp[j], v1 = rp[half4 + j];
#pragma unroll
        for (int o = 0; o < OW; ++o) {
          const int kx = j - o * S;                          // compile-time after unrolling
          if (kx >= 0 && kx < K) {
            a[o][0] = fmaf(v0.x, w0[kx].x, a[o][0]); a[o][1] = fmaf(v0.y, w0[kx].y, a[o][1]);
            a[o][2] = fmaf(v0.z, w0[kx].z, a[o][2]); a[o][3] = fmaf(v0.w, w0[kx].w, a[o][3]);
            a[o][4] = fmaf(v1.x, w1[kx].x, a[o][4]); a[o][5] = fmaf(v1.y, w1[kx].y, a[o][5]);
            a[o][6] = fmaf(v1.z, w1[kx].z, a[o][6]); a[o][7] = fmaf(v1.w, w1[kx].w, a[o][7]);
          }
        }
      }
    }
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + cc * 8)), s1 = __ldg(reinterpret_cast<const float4*>(scale + cc * 8) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + cc * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bias + cc * 8) + 1);
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, bi[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const size_t ob = (((size_t)im * ho + oy) * wo + oxs) * c8 + cc;
#pragma unroll
    for (int o = 0; o < OW; ++o) {
      if (oxs + o >= wo) break;
      float ov[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float r = fmaf(a[o][q], sc[q], bi[q]);
        ov[q] = act == B200R_ACT_RELU6 ? fminf(fmaxf(r, 0.f), 6.f)
              : act == B200R_ACT_SWISH ? __fdividef(r, 1.f + __expf(-r))
              : act_apply(r, act);
      }
      uint4 hh, ll;
      pack8(ov, hh, ll);
      yh[ob + (size_t)o * c8] = hh;
      yl[ob + (size_t)o * c8] = ll;
    }
  }
}

// out[n, p, c] = x[n, p, c] * s[n, c]   (s: split planes [n, c_stride], first c used)
__global__ void __launch_bounds__(kThreads) channel_scale_kernel(const uint4* __restrict__ xh, const uint4* __restrict__ xl,
                                                                  const uint16_t* __restrict__ sh, const uint16_t* __restrict__ sl,
                                                                  uint4* __restrict__ yh, uint4* __restrict__ yl, int n, int hw, int c8,
                                                                  int s_stride) {
  const size_t total = (size_t)n * hw * c8;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int cc = (int)(t % c8);
    const int im = (int)(t / ((size_t)hw * c8));
    float v[8], o[8];
    unpack8(__ldg(xh + t), __ldg(xl + t), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const size_t si = (size_t)im * s_stride + cc * 8 + j;
      o[j] = v[j] * (plane_bits_to_f32(sh[si]) + plane_bits_to_f32(sl[si]));
    }
    uint4 hh, ll;
    pack8(o, hh, ll);
    yh[t] = hh; yl[t] = ll;
  }
}

// generic image im2col for tiny Cin=3 stems: planes [n*ho*wo, kpad], column = (ky*k + kx)*3 + c, zero padded
struct Norm3 { float mean[3], std[3]; };
template <bool U8>
__global__ void __launch_bounds__(kThreads) image_im2col_kernel(const void* __restrict__ img, uint4* __restrict__ hi, uint4* __restrict__ lo,
                                                                 int n, int h, int w, int k, int stride, int pad, int ho, int wo, int kpad,
                                                                 Norm3 nm) {
  const int k8 = kpad / 8;
  const size_t total = (size_t)n * ho * wo * k8;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int chunk = (int)(t % k8);
    const size_t pix = t / k8;
    const int ox = (int)(pix % wo), oy = (int)((pix / wo) % ho), im = (int)(pix / ((size_t)wo * ho));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = chunk * 8 + j;
      float val = 0.f;
      if (col < k * k * 3) {
        const int tap = col / 3, c = col - tap * 3, ky = tap / k, kx = tap - ky * k;
        const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
        if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
          float x;
          if (U8) x = __fdiv_rn((float)static_cast<const uint8_t*>(img)[(((size_t)im * h + iy) * w + ix) * 3 + c], 255.0f);
          else x = static_cast<const float*>(img)[(((size_t)im * 3 + c) * h + iy) * w + ix];
          val = (x - nm.mean[c]) / nm.std[c];
        }
      }
      v[j] = val;
    }
    uint4 hh, ll;
    pack8(v, hh, ll);
    hi[t] = hh; lo[t] = ll;
  }
}

// Direct 3x3 / stride 2 / pad 1 convolution from the image (3 -> cout, cout <= 64) + folded BN + activation: the stems of
// MobileNetV2 (mobilenet_v2.py:130) and EfficientNet-B0 (efficientnet.py:429-433).  As im2col (K = 27 padded to 32) + GEMM
// the layer wrote and re-read a 205 MB patch matrix per 128 images (0.25 + 0.15 ms, 10 % of a MobileNetV2 forward); the
// arithmetic is 1 728 FLOP per output pixel, so CUDA cores at fp32 do it in the time it takes to write the output.
// thread = output pixel (all output channels); weights [27][cout] in shared memory, ToTensor + Normalize on the fly.
template <bool U8>
__global__ void __launch_bounds__(kThreads) image_stem3x3s2_kernel(const void* __restrict__ img, const float* __restrict__ wgt,
                                                                    const float* __restrict__ scale, const float* __restrict__ bias,
                                                                    uint4* __restrict__ yh, uint4* __restrict__ yl, int n, int h, int w,
                                                                    int ho, int wo, int cout, int act, Norm3 nm) {
  __shared__ __align__(16) float sw[27 * 64];
  __shared__ float lut[U8 ? 768 : 1];                                // ToTensor + Normalize of every byte value, per channel
  __shared__ float ssc[64], sbi[64];
  if (threadIdx.x < cout) { ssc[threadIdx.x] = scale[threadIdx.x]; sbi[threadIdx.x] = bias[threadIdx.x]; }
  for (int i = threadIdx.x; i < 27 * cout; i += kThreads) {          // wgt is [cout][27] (ky, kx, c) -> sw[tap][cout]
    const int t = i / cout, oc = i - t * cout;
    sw[i] = wgt[oc * 27 + t];
  }
  if (U8)
    for (int i = threadIdx.x; i < 768; i += kThreads) lut[i] = (__fdiv_rn((float)(i & 255), 255.0f) - nm.mean[i >> 8]) / nm.std[i >> 8];
  __syncthreads();
  // thread = one output pixel, ALL output channels (round 2; was (pixel, 8 channels): four threads repeated the pixel's 27 byte
  // loads and 27 table lookups, and ncu showed the L1 / shared-memory path 98 % busy): the 27 normalised inputs are gathered once
  // into registers, then each group of 8 output channels is 27 x 2 warp-uniform (broadcast) LDS.128 of weights + 216 FFMA
  const int c8 = cout / 8;
  const size_t total = (size_t)n * ho * wo;
  for (size_t pix = (size_t)blockIdx.x * kThreads + threadIdx.x; pix < total; pix += (size_t)gridDim.x * kThreads) {
    const int ox = (int)(pix % wo), oy = (int)((pix / wo) % ho), im = (int)(pix / ((size_t)wo * ho));
    float x[27];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int iy = oy * 2 - 1 + ky, ix = ox * 2 - 1 + kx;
        const bool in = iy >= 0 && iy < h && ix >= 0 && ix < w;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float v = 0.f;
          if (in) {
            if (U8) v = lut[c * 256 + static_cast<const uint8_t*>(img)[(((size_t)im * h + iy) * w + ix) * 3 + c]];
            else v = (static_cast<const float*>(img)[(((size_t)im * 3 + c) * h + iy) * w + ix] - nm.mean[c]) / nm.std[c];
          }
          x[(ky * 3 + kx) * 3 + c] = v;
        }
      }
    for (int cc = 0; cc < c8; ++cc) {
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        const float4 w0 = *reinterpret_cast<const float4*>(sw + k * cout + cc * 8), w1 = *reinterpret_cast<const float4*>(sw + k * cout + cc * 8 + 4);
        acc[0] = fmaf(x[k], w0.x, acc[0]); acc[1] = fmaf(x[k], w0.y, acc[1]); acc[2] = fmaf(x[k], w0.z, acc[2]); acc[3] = fmaf(x[k], w0.w, acc[3]);
        acc[4] = fmaf(x[k], w1.x, acc[4]); acc[5] = fmaf(x[k], w1.y, acc[5]); acc[6] = fmaf(x[k], w1.z, acc[6]); acc[7] = fmaf(x[k], w1.w, acc[7]);
      }
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float y = fmaf(acc[j], ssc[cc * 8 + j], sbi[cc * 8 + j]);
        o[j] = act == B200R_ACT_RELU6 ? fminf(fmaxf(y, 0.f), 6.f) : act == B200R_ACT_SWISH ? __fdividef(y, 1.f + __expf(-y)) : act_apply(y, act);
      }
      uint4 hh, ll;
      pack8(o, hh, ll);
      yh[pix * c8 + cc] = hh; yl[pix * c8 + cc] = ll;
    }
  }
}

// 1x1 convolution with a SMALL input width (cin <= 32) + bias + activation (+ residual) on CUDA cores, fp32: the expansion /
// projection layers at 112 x 112 and 56 x 56 of the mobile families (mobilenet_v2.py:52-60; efficientnet.py:312-321).  On the tensor-core
// GEMM these layers have one k-block of mostly zero padding per tile and nothing to hide the epilogue behind: 3.4 us per 128-row tile
// whatever the width, 1.5-2.4 TB/s (ncu: tensor pipe 15 %, DRAM 19-29 %).  Here a thread owns one pixel: its cin inputs are unpacked
// once into registers, then each group of 8 output channels is cin x 2 warp-uniform (broadcast) LDS.128 of weights and 8 cin FFMA --
// at most 32 FFMA per output value, less time than writing it.  Exact fp32 arithmetic in the reference's k order.
template <int CIN8>
__global__ void __launch_bounds__(kThreads) pointwise_smallk_kernel(const uint4* __restrict__ xh, const uint4* __restrict__ xl,
                                                                     const float* __restrict__ wgt, const float* __restrict__ bias,
                                                                     const uint4* __restrict__ rh, const uint4* __restrict__ rl,
                                                                     uint4* __restrict__ yh, uint4* __restrict__ yl, size_t m, int cout8,
                                                                     int act) {
  constexpr int CIN = CIN8 * 8;
  extern __shared__ __align__(16) float4 pw_smem[];       // [cout8][CIN][2]: the 8 weights of output group cc against input k
  for (int i = threadIdx.x; i < cout8 * CIN * 2; i += kThreads) {
    const int hf = i & 1, k = (i >> 1) % CIN, cc = (i >> 1) / CIN;
    const float* wp = wgt + (size_t)(cc * 8 + hf * 4) * CIN + k;
    pw_smem[i] = make_float4(wp[0], wp[CIN], wp[2 * CIN], wp[3 * CIN]);
  }
  __syncthreads();
  for (size_t pix = (size_t)blockIdx.x * kThreads + threadIdx.x; pix < m; pix += (size_t)gridDim.x * kThreads) {
    float x[CIN];
#pragma unroll
    for (int g = 0; g < CIN8; ++g) unpack8(__ldg(xh + pix * CIN8 + g), __ldg(xl + pix * CIN8 + g), x + 8 * g);
    // two output groups (16 channels = one full 32-byte sector per plane) per step: a 16-byte store per thread would leave every
    // sector half written until the next group's store reaches it
    for (int cc = 0; cc < cout8; cc += 2) {
      const bool two = cc + 1 < cout8;
      float a[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) a[q] = 0.f;
      if (bias) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u && !two) break;
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + (cc + u) * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bias + (cc + u) * 8) + 1);
          a[8 * u + 0] = b0.x; a[8 * u + 1] = b0.y; a[8 * u + 2] = b0.z; a[8 * u + 3] = b0.w;
          a[8 * u + 4] = b1.x; a[8 * u + 5] = b1.y; a[8 * u + 6] = b1.z; a[8 * u + 7] = b1.w;
        }
      }
      const float4* wp = pw_smem + (size_t)cc * CIN * 2;
      const float4* wq = two ? wp + CIN * 2 : wp;
#pragma unroll
      for (int k = 0; k < CIN; ++k) {
        const float4 w0 = wp[2 * k], w1 = wp[2 * k + 1], w2 = wq[2 * k], w3 = wq[2 * k + 1];
        a[0] = fmaf(x[k], w0.x, a[0]); a[1] = fmaf(x[k], w0.y, a[1]); a[2] = fmaf(x[k], w0.z, a[2]); a[3] = fmaf(x[k], w0.w, a[3]);
        a[4] = fmaf(x[k], w1.x, a[4]); a[5] = fmaf(x[k], w1.y, a[5]); a[6] = fmaf(x[k], w1.z, a[6]); a[7] = fmaf(x[k], w1.w, a[7]);
        a[8] = fmaf(x[k], w2.x, a[8]); a[9] = fmaf(x[k], w2.y, a[9]); a[10] = fmaf(x[k], w2.z, a[10]); a[11] = fmaf(x[k], w2.w, a[11]);
        a[12] = fmaf(x[k], w3.x, a[12]); a[13] = fmaf(x[k], w3.y, a[13]); a[14] = fmaf(x[k], w3.z, a[14]); a[15] = fmaf(x[k], w3.w, a[15]);
      }
      const size_t o = pix * cout8 + cc;
      if (rh) {
        float r[16];
        unpack8(__ldg(rh + o), __ldg(rl + o), r);
        if (two) unpack8(__ldg(rh + o + 1), __ldg(rl + o + 1), r + 8);
#pragma unroll
        for (int q = 0; q < 16; ++q) a[q] += (q < 8 || two) ? r[q] : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 16; ++q)
        a[q] = act == B200R_ACT_RELU6 ? fminf(fmaxf(a[q], 0.f), 6.f)
             : act == B200R_ACT_SWISH ? __fdividef(a[q], 1.f + __expf(-a[q]))
             : act_apply(a[q], act);
      uint4 h0, l0, h1, l1;
      pack8(a, h0, l0);
      pack8(a + 8, h1, l1);
      if (two && (cout8 & 1) == 0) {        // 32-byte aligned pair: one 256-bit store per plane
        st_v8(yh + o, h0, h1);
        st_v8(yl + o, l0, l1);
      } else {
        yh[o] = h0; yl[o] = l0;
        if (two) { yh[o + 1] = h1; yl[o + 1] = l1; }
      }
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

int b200r_dwconv_nhwc(const uint16_t* x, const float* wgt, const float* scale, const float* bias, uint16_t* y, int n, int h, int w,
                      int c, int k, int stride, int pad, int act, b200r_stream_t stream) {
  B200R_CHECK_ARG(x && wgt && scale && bias && y, "null pointer");
  B200R_CHECK_ARG(n > 0 && h > 0 && w > 0 && c % 8 == 0 && k >= 1 && stride >= 1, "bad shape (c must be a multiple of 8)");
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  const size_t cin = (size_t)n * h * w * c, cout = (size_t)n * ho * wo * c;
  const uint4 *xh = reinterpret_cast<const uint4*>(x), *xl = reinterpret_cast<const uint4*>(x + cin);
  uint4 *yh = reinterpret_cast<uint4*>(y), *yl = reinterpret_cast<uint4*>(y + cout);
  static int use_strip = -1;      // B200R_DW_STRIP=0: generic kernel for every shape (A/B measurements)
  if (use_strip < 0) { const char* e = getenv("B200R_DW_STRIP"); use_strip = (e && e[0] == '0') ? 0 : 1; }
  const char* tile_env = getenv("B200R_DW_TILE");          // B200R_DW_TILE=0: the round-1 strip kernel
  // measured (batch 256): the tile kernel wins for 5 x 5 (25 taps of reuse: 0.31 -> 0.23 ms at 240 ch @28), the strip kernel for 3 x 3
  // (0.31 vs 0.40 ms at 144 ch @56: the two-phase tile does not hide its fill at 9 taps); B200R_DW_TILE=1 forces the tile kernel
  const bool want_tile = tile_env ? tile_env[0] == '1' : k == 5;
  if (want_tile && pad == k / 2 && (k == 3 || k == 5) && (stride == 1 || stride == 2)) {
    // strips of OW outputs (4 at stride 1, 2 at stride 2: the strip pitch stays 64 bytes, conflict-free); tile = TW x TH outputs x CG
    // channel groups with ~256 strips per CTA
    const int c8 = c / 8, OW = stride == 1 ? 4 : 2;
    int TW = wo < 16 ? ((wo + OW - 1) / OW) * OW : 16;
    const int SW = TW / OW;
    int TH = 64 / SW;                                        // with CG = 4: 256 strips
    if (TH > ho) TH = ho;
    int CG = 256 / (TH * SW);
    if (CG < 4) CG = 4;
    if (CG > c8) CG = c8;
    const int THI = (TH - 1) * stride + k, TWI = (TW - 1) * stride + k;
    int plane4 = THI * TWI;                                  // float4 units; pad to 1 mod 8: the next group starts 16 bytes mod 128 later
    plane4 += (9 - plane4 % 8) % 8;
    const int tiles_x = (wo + TW - 1) / TW, tiles_y = (ho + TH - 1) / TH, tiles_c = (c8 + CG - 1) / CG;
    const size_t smem = ((size_t)k * k * CG * 2 + (size_t)CG * 2 * plane4) * sizeof(float4);
    const long long blocks = (long long)n * tiles_y * tiles_x * tiles_c;
    if (smem <= 96 * 1024 && blocks < (1LL << 31)) {
#define B200R_DWT(K, S, O)                                                                                                           \
      do {                                                                                                                           \
        static size_t conf = 0;                                                                                                      \
        if (conf < smem) { B200R_CUDA(cudaFuncSetAttribute(dwconv_tile_kernel<K, S, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); conf = smem; } \
        dwconv_tile_kernel<K, S, O><<<(unsigned)blocks, kThreads, smem, as_stream(stream)>>>(xh, xl, wgt, scale, bias, yh, yl, h, w, c8, ho, wo, act, \
                                                                                            TH, TW, CG, tiles_x, tiles_y, tiles_c, plane4); \
      } while (0)
      if (k == 3 && stride == 1) B200R_DWT(3, 1, 4);
      else if (k == 3) B200R_DWT(3, 2, 2);
      else if (stride == 1) B200R_DWT(5, 1, 4);
      else B200R_DWT(5, 2, 2);
#undef B200R_DWT
      B200R_LAUNCH_CHECK();
      return B200R_OK;
    }
  }
  if (use_strip && pad == k / 2 && (k == 3 || k == 5) && (stride == 1 || stride == 2)) {
    const size_t threads = (size_t)n * ho * ((wo + 3) / 4) * (c / 8);
#define B200R_DW(K, S) dwconv_strip_kernel<K, S><<<grid_for(threads), kThreads, 0, as_stream(stream)>>>(xh, xl, wgt, scale, bias, yh, yl, n, h, w, c / 8, ho, wo, act)
    if (k == 3 && stride == 1) B200R_DW(3, 1);
    else if (k == 3) B200R_DW(3, 2);
    else if (stride == 1) B200R_DW(5, 1);
    else B200R_DW(5, 2);
#undef B200R_DW
  } else {
    dwconv_kernel<<<grid_for(cout / 8), kThreads, 0, as_stream(stream)>>>(xh, xl, wgt, scale, bias, yh, yl, n, h, w, c / 8, k, stride, pad, ho, wo, act);
  }
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_pointwise_smallk_nhwc(const uint16_t* x, const float* wgt, const float* bias, const uint16_t* res, uint16_t* y, size_t m, int cin,
                                int cout, int act, b200r_stream_t stream) {
  B200R_CHECK_ARG(x && wgt && y, "null pointer");
  B200R_CHECK_ARG(m > 0 && cin % 8 == 0 && cin >= 8 && cin <= 32 && cout % 8 == 0 && cout > 0, "pointwise_smallk: cin in {8, 16, 24, 32}, cout a multiple of 8");
  B200R_CHECK_ARG(act >= B200R_ACT_NONE && act <= B200R_ACT_SIGMOID, "bad activation");
  const size_t smem = (size_t)cout * cin * sizeof(float);
  B200R_CHECK_ARG(smem <= 96 * 1024, "pointwise_smallk: cout * cin too large for the shared-memory weight tile");
  const size_t xin = m * cin, yout = m * cout;
  const uint4 *xh = reinterpret_cast<const uint4*>(x), *xl = reinterpret_cast<const uint4*>(x + xin);
  const uint4 *rh = reinterpret_cast<const uint4*>(res), *rl = res ? reinterpret_cast<const uint4*>(res + yout) : nullptr;
  uint4 *yh = reinterpret_cast<uint4*>(y), *yl = reinterpret_cast<uint4*>(y + yout);
#define B200R_PW(C8)                                                                                                                  \
  do {                                                                                                                                \
    static size_t conf = 0;                                                                                                           \
    if (conf < smem) { B200R_CUDA(cudaFuncSetAttribute(pointwise_smallk_kernel<C8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); conf = smem; } \
    pointwise_smallk_kernel<C8><<<grid_for(m), kThreads, smem, as_stream(stream)>>>(xh, xl, wgt, bias, rh, rl, yh, yl, m, cout / 8, act); \
  } while (0)
  switch (cin / 8) {
    case 1: B200R_PW(1); break;
    case 2: B200R_PW(2); break;
    case 3: B200R_PW(3); break;
    default: B200R_PW(4); break;
  }
#undef B200R_PW
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_channel_scale(const uint16_t* x, const uint16_t* s, uint16_t* y, int n, int hw, int c, int s_stride, b200r_stream_t stream) {
  B200R_CHECK_ARG(x && s && y, "null pointer");
  B200R_CHECK_ARG(n > 0 && hw > 0 && c % 8 == 0 && s_stride >= c, "bad shape");
  const size_t cnt = (size_t)n * hw * c, scnt = (size_t)n * s_stride;
  channel_scale_kernel<<<grid_for(cnt / 8), kThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(x + cnt), s, s + scnt, reinterpret_cast<uint4*>(y),
      reinterpret_cast<uint4*>(y + cnt), n, hw, c / 8, s_stride);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

static int im2col_common(const void* img, uint16_t* planes, int n, int h, int w, int k, int stride, int pad, int kpad, const float* mean,
                         const float* stdv, bool u8, cudaStream_t s) {
  B200R_CHECK_ARG(img && planes && mean && stdv, "null pointer");
  B200R_CHECK_ARG(n > 0 && k >= 1 && stride >= 1 && kpad % 8 == 0 && kpad >= k * k * 3, "bad im2col geometry");
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  const size_t rows = (size_t)n * ho * wo;
  Norm3 nm;
  for (int i = 0; i < 3; ++i) { nm.mean[i] = mean[i]; nm.std[i] = stdv[i]; }
  uint4* hi = reinterpret_cast<uint4*>(planes);
  uint4* lo = reinterpret_cast<uint4*>(planes + rows * kpad);
  if (u8) image_im2col_kernel<true><<<grid_for(rows * (kpad / 8)), kThreads, 0, s>>>(img, hi, lo, n, h, w, k, stride, pad, ho, wo, kpad, nm);
  else image_im2col_kernel<false><<<grid_for(rows * (kpad / 8)), kThreads, 0, s>>>(img, hi, lo, n, h, w, k, stride, pad, ho, wo, kpad, nm);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}
int b200r_image_im2col_u8(const uint8_t* img, uint16_t* planes, int n, int h, int w, int k, int stride, int pad, int kpad,
                          const float* mean_host, const float* std_host, b200r_stream_t stream) {
  return im2col_common(img, planes, n, h, w, k, stride, pad, kpad, mean_host, std_host, true, as_stream(stream));
}
int b200r_image_im2col_f32(const float* img, uint16_t* planes, int n, int h, int w, int k, int stride, int pad, int kpad,
                           const float* mean_host, const float* std_host, b200r_stream_t stream) {
  return im2col_common(img, planes, n, h, w, k, stride, pad, kpad, mean_host, std_host, false, as_stream(stream));
}

static int stem3_common(const void* img, const float* wgt, const float* scale, const float* bias, uint16_t* y, int n, int h, int w, int cout,
                        int act, const float* mean, const float* stdv, bool u8, cudaStream_t s) {
  B200R_CHECK_ARG(img && wgt && scale && bias && y && mean && stdv, "null pointer");
  B200R_CHECK_ARG(n > 0 && h > 0 && w > 0 && cout > 0 && cout % 8 == 0 && cout <= 64, "image stem: cout must be a multiple of 8, <= 64");
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const size_t cnt = (size_t)n * ho * wo * cout;
  Norm3 nm;
  for (int i = 0; i < 3; ++i) { nm.mean[i] = mean[i]; nm.std[i] = stdv[i]; }
  uint4 *yh = reinterpret_cast<uint4*>(y), *yl = reinterpret_cast<uint4*>(y + cnt);
  if (u8) image_stem3x3s2_kernel<true><<<grid_for(cnt / cout), kThreads, 0, s>>>(img, wgt, scale, bias, yh, yl, n, h, w, ho, wo, cout, act, nm);
  else image_stem3x3s2_kernel<false><<<grid_for(cnt / cout), kThreads, 0, s>>>(img, wgt, scale, bias, yh, yl, n, h, w, ho, wo, cout, act, nm);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}
int b200r_image_stem3x3s2_u8(const uint8_t* img, const float* wgt, const float* scale, const float* bias, uint16_t* y, int n, int h, int w,
                             int cout, int act, const float* mean_host, const float* std_host, b200r_stream_t stream) {
  return stem3_common(img, wgt, scale, bias, y, n, h, w, cout, act, mean_host, std_host, true, as_stream(stream));
}
int b200r_image_stem3x3s2_f32(const float* img, const float* wgt, const float* scale, const float* bias, uint16_t* y, int n, int h, int w,
                              int cout, int act, const float* mean_host, const float* std_host, b200r_stream_t stream) {
  return stem3_common(img, wgt, scale, bias, y, n, h, w, cout, act, mean_host, std_host, false, as_stream(stream));
}


// ---- input-gradient pieces of the mobile families (round 2) ---------------------------------------------------------------
// autograd of mobilenet_v2.py:31-77 / efficientnet.py:312-360 as the attack loops see it (autopgd_base.py:371-376): the depthwise
// convolution's input gradient is the SAME kernel on flipped taps (stride 2: after b200r_dilate2_nhwc), the 1x1 convolutions are
// b200r_conv2d_dgrad_nhwc, the activations b200r_act_bwd_planes; what is new here is the squeeze-excite reduction, a plain add, and
// the transposed 3x3/s2 image stem.
}  // extern "C"  (kernels below live in the anonymous namespace again)
namespace {

// ds[n, c] = sum_p a[n, p, c] * b[n, p, c]  (squeeze-excite: gradient w.r.t. the per-channel scale).  One 256-thread block per
// (image, 8-channel chunk): threads stride over the pixels, fixed-order shared-memory tree.
__global__ void __launch_bounds__(256) channel_dot_kernel(const uint4* __restrict__ ah, const uint4* __restrict__ al, const uint4* __restrict__ bh,
                                                          const uint4* __restrict__ bl, uint16_t* __restrict__ oh, uint16_t* __restrict__ ol, int hw,
                                                          int c8, int s_stride) {
  __shared__ float red[256][8];
  const int im = blockIdx.y, cc = blockIdx.x;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int p = threadIdx.x; p < hw; p += 256) {
    const size_t i = ((size_t)im * hw + p) * c8 + cc;
    float a[8], b[8];
    unpack8(__ldg(ah + i), __ldg(al + i), a);
    unpack8(__ldg(bh + i), __ldg(bl + i), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(a[j], b[j], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = acc[j];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s)
#pragma unroll
      for (int j = 0; j < 8; ++j) red[threadIdx.x][j] += red[threadIdx.x + s][j];
    __syncthreads();
  }
  if (threadIdx.x < 8) {
    uint16_t h, l;
    split_pair(red[0][threadIdx.x], h, l);
    oh[(size_t)im * s_stride + cc * 8 + threadIdx.x] = h;
    ol[(size_t)im * s_stride + cc * 8 + threadIdx.x] = l;
  }
}

__global__ void __launch_bounds__(kThreads) planes_add_kernel(const uint4* __restrict__ ah, const uint4* __restrict__ al, const uint4* __restrict__ bh,
                                                               const uint4* __restrict__ bl, uint4* __restrict__ oh, uint4* __restrict__ ol, size_t count8) {
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < count8; t += (size_t)gridDim.x * kThreads) {
    float a[8], b[8];
    unpack8(__ldg(ah + t), __ldg(al + t), a);
    unpack8(__ldg(bh + t), __ldg(bl + t), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    uint4 h, l;
    pack8(a, h, l);
    oh[t] = h; ol[t] = l;
  }
}

// transposed 3x3/s2/p1 image stem: dx[n, c, iy, ix] = unscale / std[c] * sum_{oc, ky, kx : 2 oy + ky - 1 = iy, 2 ox + kx - 1 = ix} w[oc][ky][kx][c] dy[n, oy, ox, oc]
// thread = input pixel (three channels); weights [27][cout] in shared memory (BN scale folded in by the caller)
__global__ void __launch_bounds__(kThreads) image_stem3x3s2_bwd_kernel(const uint4* __restrict__ dyh, const uint4* __restrict__ dyl,
                                                                        const float* __restrict__ wgt, float* __restrict__ dx, int n, int h, int w,
                                                                        int ho, int wo, int cout, float k0, float k1, float k2) {
  __shared__ __align__(16) float sw[27 * 64];
  for (int i = threadIdx.x; i < 27 * cout; i += kThreads) {          // wgt is [cout][27] (ky, kx, c) -> sw[tap][cout]
    const int t = i / cout, oc = i - t * cout;
    sw[i] = wgt[oc * 27 + t];
  }
  __syncthreads();
  const int c8 = cout / 8;
  const size_t total = (size_t)n * h * w;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int ix = (int)(t % w), iy = (int)((t / w) % h), im = (int)(t / ((size_t)w * h));
    float acc[3] = {0.f, 0.f, 0.f};
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = iy + 1 - ky;
      if (ty < 0 || (ty & 1) || (ty >> 1) >= ho) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = ix + 1 - kx;
        if (tx < 0 || (tx & 1) || (tx >> 1) >= wo) continue;
        const size_t base = (((size_t)im * ho + (ty >> 1)) * wo + (tx >> 1)) * c8;
        const float* wk = sw + (ky * 3 + kx) * 3 * cout;
        for (int cc = 0; cc < c8; ++cc) {
          float g[8];
          unpack8(__ldg(dyh + base + cc), __ldg(dyl + base + cc), g);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc[0] = fmaf(g[j], wk[cc * 8 + j], acc[0]);
            acc[1] = fmaf(g[j], wk[cout + cc * 8 + j], acc[1]);
            acc[2] = fmaf(g[j], wk[2 * cout + cc * 8 + j], acc[2]);
          }
        }
      }
    }
    const size_t o = ((size_t)im * 3 * h + iy) * w + ix;
    dx[o] = acc[0] * k0; dx[o + (size_t)h * w] = acc[1] * k1; dx[o + 2 * (size_t)h * w] = acc[2] * k2;
  }
}

}  // namespace
extern "C" {

/* ds planes [n, s_stride] (first c entries) = sum over pixels of a * b, a / b planes [n, hw, c] */
int b200r_channel_dot(const uint16_t* a, const uint16_t* b, uint16_t* ds, int n, int hw, int c, int s_stride, b200r_stream_t stream) {
  B200R_CHECK_ARG(a && b && ds, "null pointer");
  B200R_CHECK_ARG(c % 8 == 0 && s_stride >= c && n > 0 && n < 65536 && hw > 0, "bad shape (c %% 8 == 0, s_stride >= c)");
  const size_t cnt = (size_t)n * hw * c, scnt = (size_t)n * s_stride;
  B200R_CUDA(cudaMemsetAsync(ds, 0, scnt * 2 * sizeof(uint16_t), as_stream(stream)));       // padding columns of the scale vector stay zero
  channel_dot_kernel<<<dim3(c / 8, n), 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(a + cnt),
                                                                    reinterpret_cast<const uint4*>(b), reinterpret_cast<const uint4*>(b + cnt), ds,
                                                                    ds + scnt, hw, c / 8, s_stride);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

/* out = a + b on split planes; count = elements per plane, multiple of 8 */
int b200r_planes_add(const uint16_t* a, const uint16_t* b, uint16_t* out, size_t count, b200r_stream_t stream) {
  B200R_CHECK_ARG(a && b && out, "null pointer");
  B200R_CHECK_ARG(count % 8 == 0, "count must be a multiple of 8");
  if (!count) return B200R_OK;
  planes_add_kernel<<<grid_for(count / 8), kThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(a + count), reinterpret_cast<const uint4*>(b),
      reinterpret_cast<const uint4*>(b + count), reinterpret_cast<uint4*>(out), reinterpret_cast<uint4*>(out + count), count / 8);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

/* transpose of b200r_image_stem3x3s2_f32 composed with Normalize: dy planes [n, ho, wo, cout] -> float32 NCHW gradient w.r.t. the
 * [0,1] image, times `unscale`; wgt float32 [cout][27] with the BN scale already folded in */
int b200r_image_stem3x3s2_bwd(const uint16_t* dy, const float* wgt, float* dx, int n, int h, int w, int cout, const float* std_host,
                              float unscale, b200r_stream_t stream) {
  B200R_CHECK_ARG(dy && wgt && dx && std_host, "null pointer");
  B200R_CHECK_ARG(cout % 8 == 0 && cout <= 64 && n > 0 && h > 0 && w > 0, "cout must be a multiple of 8, at most 64");
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const size_t cnt = (size_t)n * ho * wo * cout;
  image_stem3x3s2_bwd_kernel<<<grid_for((size_t)n * h * w), kThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(dy + cnt), wgt, dx, n, h, w, ho, wo, cout, unscale / std_host[0],
      unscale / std_host[1], unscale / std_host[2]);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

}  // extern "C"
