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
  for (int j = 0; j < 4; ++j) {
    uint16_t h0, l0, h1, l1;
    split_pair(v[2 * j], h0, l0);
    split_pair(v[2 * j + 1], h1, l1);
    hw[j] = h0 | ((uint32_t)h1 << 16);
    lw[j] = l0 | ((uint32_t)l1 << 16);
  }
  h = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  l = make_uint4(lw[0], lw[1], lw[2], lw[3]);
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
// thread = (output pixel, 8 output channels); weights [27][cout] in shared memory, ToTensor + Normalize on the fly.
template <bool U8>
__global__ void __launch_bounds__(kThreads) image_stem3x3s2_kernel(const void* __restrict__ img, const float* __restrict__ wgt,
                                                                    const float* __restrict__ scale, const float* __restrict__ bias,
                                                                    uint4* __restrict__ yh, uint4* __restrict__ yl, int n, int h, int w,
                                                                    int ho, int wo, int cout, int act, Norm3 nm) {
  __shared__ __align__(16) float sw[27 * 64];
  __shared__ float lut[U8 ? 768 : 1];                                // ToTensor + Normalize of every byte value, per channel
  for (int i = threadIdx.x; i < 27 * cout; i += kThreads) {          // wgt is [cout][27] (ky, kx, c) -> sw[tap][cout]
    const int t = i / cout, oc = i - t * cout;
    sw[i] = wgt[oc * 27 + t];
  }
  if (U8)
    for (int i = threadIdx.x; i < 768; i += kThreads) lut[i] = (__fdiv_rn((float)(i & 255), 255.0f) - nm.mean[i >> 8]) / nm.std[i >> 8];
  __syncthreads();
  const int c8 = cout / 8;
  const size_t total = (size_t)n * ho * wo * c8;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int cc = (int)(t % c8);
    const size_t pix = t / c8;
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
      const float y = fmaf(acc[j], scale[cc * 8 + j], bias[cc * 8 + j]);
      o[j] = act == B200R_ACT_RELU6 ? fminf(fmaxf(y, 0.f), 6.f) : act == B200R_ACT_SWISH ? __fdividef(y, 1.f + __expf(-y)) : act_apply(y, act);
    }
    uint4 hh, ll;
    pack8(o, hh, ll);
    yh[t] = hh; yl[t] = ll;
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
  if (u8) image_stem3x3s2_kernel<true><<<grid_for(cnt / 8), kThreads, 0, s>>>(img, wgt, scale, bias, yh, yl, n, h, w, ho, wo, cout, act, nm);
  else image_stem3x3s2_kernel<false><<<grid_for(cnt / 8), kThreads, 0, s>>>(img, wgt, scale, bias, yh, yl, n, h, w, ho, wo, cout, act, nm);
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
