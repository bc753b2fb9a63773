// Memory-bound pieces of the INPUT-gradient pass (attack loops need d loss / d image only, never weight
// gradients: autopgd_base.py:371-376, foolbox gradient_descent_base.py value_and_grad).  All tensors are
// split-bf16 planes, NHWC, like the forward layers (layers.cu).  The contractions of the backward pass are the
// forward GEMM kernel run on transposed / flipped weights (gemm_sm100.cu); what is left is:
//   relu_bwd      out = (act > 0 ? dy : 0) [+ add]           ReLU backward, optionally joining a second branch
//   dilate2       y[n,2i,2j,:] = x[n,i,j,:], zeros elsewhere  turns a stride-2 conv's dgrad into a stride-1 conv
//   maxpool_bwd   MaxPool2d(3,2,1) backward with PyTorch's first-maximum tie rule
//   avgpool_bwd   AdaptiveAvgPool2d(1) backward
//   stem_col2im   im2col^T of the 7x7/s2 stem + the 1/std of Normalize -> float32 NCHW image gradient
// Every kernel is a template on F16: false = split-bf16 planes (hi, lo), true = ONE fp16 plane (the `*l` pointers are
// unused), the activation format of passes = B200R_PASSES_F16.
#include "common.cuh"
#include <cuda_fp16.h>

namespace {
constexpr int kThreads = 256;

// 8 consecutive channels of a tensor as floats, from one 16-byte piece per plane
template <bool F16>
__device__ __forceinline__ void load8(const uint4* __restrict__ ph, const uint4* __restrict__ pl, size_t i, float (&v)[8]) {
  const uint4 a = __ldg(ph + i);
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
  if (F16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&aw[j]));
      v[2 * j] = f.x; v[2 * j + 1] = f.y;
    }
  } else {
    const uint4 b = __ldg(pl + i);
    const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j)
      v[j] = plane_bits_to_f32((uint16_t)(aw[j >> 1] >> (16 * (j & 1)))) + plane_bits_to_f32((uint16_t)(bw[j >> 1] >> (16 * (j & 1))));
  }
}
template <bool F16>
__device__ __forceinline__ void unpack8r(const uint4& a, const uint4& b, float (&v)[8]) {
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
  if (F16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&aw[j]));
      v[2 * j] = f.x; v[2 * j + 1] = f.y;
    }
  } else {
    const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[2 * j] = plane_lo16_f32(aw[j]) + plane_lo16_f32(bw[j]);
      v[2 * j + 1] = plane_hi16_f32(aw[j]) + plane_hi16_f32(bw[j]);
    }
  }
}
template <bool F16>
__device__ __forceinline__ void store8(uint4* __restrict__ ph, uint4* __restrict__ pl, size_t i, const float (&v)[8]) {
  uint32_t rh[4], rl[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (F16) {
      const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
      rh[j] = *reinterpret_cast<const uint32_t*>(&h);
    } else {
      split_pair2(v[2 * j], v[2 * j + 1], rh[j], rl[j]);
    }
  }
  ph[i] = make_uint4(rh[0], rh[1], rh[2], rh[3]);
  if (!F16) pl[i] = make_uint4(rl[0], rl[1], rl[2], rl[3]);
}

__device__ __forceinline__ uint32_t pack2(uint16_t a, uint16_t b) { return (uint32_t)a | ((uint32_t)b << 16); }
__device__ __forceinline__ float plane_val(uint32_t hw, uint32_t lw, int odd) {
  return plane_bits_to_f32((uint16_t)(hw >> (16 * odd))) + plane_bits_to_f32((uint16_t)(lw >> (16 * odd)));
}

inline unsigned grid_for(size_t items) {
  size_t b = (items + kThreads - 1) / kThreads;
  size_t cap = (size_t)b200r_num_sms() * 16;
  return (unsigned)(b < cap ? (b ? b : 1) : cap);
}

// ---- ReLU backward, 8 elements per thread ---------------------------------------------------------
template <bool ADD, bool F16>
__global__ void __launch_bounds__(kThreads) relu_bwd_kernel(const uint4* __restrict__ dyh, const uint4* __restrict__ dyl,
                                                             const uint4* __restrict__ acth, const uint4* __restrict__ addh,
                                                             const uint4* __restrict__ addl, uint4* __restrict__ oh,
                                                             uint4* __restrict__ ol, size_t count8) {
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < count8; i += (size_t)gridDim.x * kThreads) {
    const uint4 a = __ldg(acth + i), gh = __ldg(dyh + i), gl = F16 ? make_uint4(0, 0, 0, 0) : __ldg(dyl + i);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, hw[4] = {gh.x, gh.y, gh.z, gh.w}, lw[4] = {gl.x, gl.y, gl.z, gl.w};
    uint32_t rh[4], rl[4];
    if (ADD && F16) {
      float g[8], b[8];
      load8<true>(dyh, nullptr, i, g);
      load8<true>(addh, nullptr, i, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint16_t ab = (uint16_t)(aw[j >> 1] >> (16 * (j & 1)));
        g[j] = ((ab != 0 && !(ab & 0x8000u)) ? g[j] : 0.f) + b[j];
      }
      store8<true>(oh, nullptr, i, g);
      continue;
    }
    if (!ADD) {
      // pure masking keeps the planes bit-exact: a positive activation has a positive hi plane (bf16 rounding
      // keeps the sign; 0x0000 / 0x8000 / negative patterns are "not > 0")
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t m0 = ((aw[j] & 0xFFFFu) != 0 && !(aw[j] & 0x8000u)) ? 0x0000FFFFu : 0u;
        const uint32_t m1 = ((aw[j] >> 16) != 0 && !(aw[j] & 0x80000000u)) ? 0xFFFF0000u : 0u;
        rh[j] = hw[j] & (m0 | m1);
        rl[j] = lw[j] & (m0 | m1);
      }
    } else {
      const uint4 bh = __ldg(addh + i), bl = __ldg(addl + i);
      const uint32_t bhw[4] = {bh.x, bh.y, bh.z, bh.w}, blw[4] = {bl.x, bl.y, bl.z, bl.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint16_t h2[2], l2[2];
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const uint16_t ab = (uint16_t)(aw[j] >> (16 * o));
          const bool pos = ab != 0 && !(ab & 0x8000u);
          const float v = (pos ? plane_val(hw[j], lw[j], o) : 0.f) + plane_val(bhw[j], blw[j], o);
          split_pair(v, h2[o], l2[o]);
        }
        rh[j] = pack2(h2[0], h2[1]);
        rl[j] = pack2(l2[0], l2[1]);
      }
    }
    oh[i] = make_uint4(rh[0], rh[1], rh[2], rh[3]);
    if (!F16) ol[i] = make_uint4(rl[0], rl[1], rl[2], rl[3]);
  }
}

// ---- zero insertion: y[n, 2i, 2j, :] = x[n, i, j, :] ----------------------------------------------
template <bool F16>
__global__ void __launch_bounds__(kThreads) dilate2_kernel(const uint4* __restrict__ xh, const uint4* __restrict__ xl,
                                                            uint4* __restrict__ yh, uint4* __restrict__ yl, int n, int h, int w,
                                                            int c8) {
  const size_t total = (size_t)n * (2 * h) * (2 * w) * c8;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int cc = (int)(t % c8);
    const size_t pix = t / c8;
    const int q = (int)(pix % (2 * w)), p = (int)((pix / (2 * w)) % (2 * h)), im = (int)(pix / ((size_t)4 * w * h));
    uint4 vh = make_uint4(0, 0, 0, 0), vl = vh;
    if (!((p | q) & 1)) {
      const size_t src = (((size_t)im * h + (p >> 1)) * w + (q >> 1)) * c8 + cc;
      vh = __ldg(xh + src);
      if (!F16) vl = __ldg(xl + src);
    }
    yh[t] = vh;
    if (!F16) yl[t] = vl;
  }
}

// ---- MaxPool2d(3, 2, 1) backward --------------------------------------------------------------------
// pass 1: per output window, the position (ky*3+kx) of the first maximum, as the forward kernel picks it
// RELU: x is a post-ReLU activation and the ReLU's backward is fused in -- a window whose maximum is not positive routes nothing
template <bool F16, bool RELU = false>
__global__ void __launch_bounds__(kThreads) maxpool_argmax_kernel(const uint4* __restrict__ xh, const uint4* __restrict__ xl,
                                                                   uint2* __restrict__ idx, int n, int h, int w, int c8, int ho,
                                                                   int wo) {
  const size_t total = (size_t)n * ho * wo * c8;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int cc = (int)(t % c8);
    const size_t pix = t / c8;
    const int ox = (int)(pix % wo), oy = (int)((pix / wo) % ho), im = (int)(pix / ((size_t)wo * ho));
    float best[8];
    uint32_t bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0xFu; }
    // the nine window loads first (clamped address, validity kept aside): independent 16-byte loads in flight
    uint4 ra[9], rb[9];
    bool ok[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int iy = oy * 2 - 1 + k / 3, ix = ox * 2 - 1 + k % 3;
      ok[k] = iy >= 0 && iy < h && ix >= 0 && ix < w;
      const size_t s = (((size_t)im * h + (ok[k] ? iy : 0)) * w + (ok[k] ? ix : 0)) * c8 + cc;
      ra[k] = __ldg(xh + s);
      rb[k] = F16 ? make_uint4(0, 0, 0, 0) : __ldg(xl + s);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      if (!ok[k]) continue;
      float v[8];
      unpack8r<F16>(ra[k], rb[k], v);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (v[j] > best[j]) { best[j] = v[j]; bi[j] = (uint32_t)k; }
    }
    if (RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (!(best[j] > 0.f)) bi[j] = 0xFu;
    }
    idx[t] = make_uint2(bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24), bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24));
  }
}
// pass 2: per input position, the sum of dy over the (at most four) windows whose maximum it is
// HI_ONLY: dx is ONE fp16 plane whatever the precision of dy (the stem's gradient GEMM reads a single plane)
template <bool F16, bool HI_ONLY = false>
__global__ void __launch_bounds__(kThreads) maxpool_bwd_kernel(const uint2* __restrict__ idx, const uint4* __restrict__ dyh,
                                                                const uint4* __restrict__ dyl, uint4* __restrict__ dxh,
                                                                uint4* __restrict__ dxl, int n, int h, int w, int c8, int ho, int wo) {
  const size_t total = (size_t)n * h * w * c8;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int cc = (int)(t % c8);
    const size_t pix = t / c8;
    const int q = (int)(pix % w), p = (int)((pix / w) % h), im = (int)(pix / ((size_t)w * h));
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool any = false;
    // windows with 2*oy - 1 <= p <= 2*oy + 1 (one or two per axis): their argmax codes and gradients are loaded together
    uint2 id[4];
    uint4 ga[4], gb[4];
    uint32_t code[4];
    bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int oy = (p >> 1) + (k >> 1), ox = (q >> 1) + (k & 1);
      ok[k] = oy <= ((p + 1) >> 1) && ox <= ((q + 1) >> 1) && oy < ho && ox < wo;
      code[k] = (uint32_t)((p - (2 * oy - 1)) * 3 + (q - (2 * ox - 1)));
      const size_t o = (((size_t)im * ho + (ok[k] ? oy : 0)) * wo + (ok[k] ? ox : 0)) * c8 + cc;
      id[k] = __ldg(idx + o);
      ga[k] = __ldg(dyh + o);
      gb[k] = F16 ? make_uint4(0, 0, 0, 0) : __ldg(dyl + o);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!ok[k]) continue;
      const uint32_t m0 = id[k].x ^ (code[k] * 0x01010101u), m1 = id[k].y ^ (code[k] * 0x01010101u);
      // any byte of m0 / m1 equal to zero = this position is the argmax of that channel
      if (!(((m0 - 0x01010101u) & ~m0 & 0x80808080u) | ((m1 - 0x01010101u) & ~m1 & 0x80808080u))) continue;
      // convert only the channels that match (on average one per window): the hi/lo unpack of all eight was most of this kernel's
      // instructions (ncu: issue 78 %, DRAM 15 %)
      const uint32_t gaw[4] = {ga[k].x, ga[k].y, ga[k].z, ga[k].w}, gbw[4] = {gb[k].x, gb[k].y, gb[k].z, gb[k].w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t b = ((j < 4 ? id[k].x : id[k].y) >> (8 * (j & 3))) & 0xFFu;
        if (b == code[k]) {
          const uint32_t wh = gaw[j >> 1], wl = gbw[j >> 1];
          float gj = (j & 1) ? plane_hi16_f32(wh) : plane_lo16_f32(wh);
          if (!F16) gj += (j & 1) ? plane_hi16_f32(wl) : plane_lo16_f32(wl);
          acc[j] += gj;
          any = true;
        }
      }
    }
    if (any) {
      store8<F16 || HI_ONLY>(dxh, dxl, t, acc);
    } else {
      dxh[t] = make_uint4(0, 0, 0, 0);
      if (!F16 && !HI_ONLY) dxl[t] = make_uint4(0, 0, 0, 0);
    }
  }
}

// ---- global average pool backward: dx[n, p, c] = dy[n, c] / hw ----------------------------------------
template <bool F16>
__global__ void __launch_bounds__(kThreads) avgpool_bwd_kernel(const uint4* __restrict__ dyh, const uint4* __restrict__ dyl,
                                                                uint4* __restrict__ dxh, uint4* __restrict__ dxl, int n, int hw,
                                                                int c8) {
  const size_t total = (size_t)n * hw * c8;
  const float inv = 1.0f / (float)hw;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int cc = (int)(t % c8);
    const int im = (int)(t / ((size_t)hw * c8));
    float g[8];
    load8<F16>(dyh, dyl, (size_t)im * c8 + cc, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= inv;
    store8<F16>(dxh, dxl, t, g);
  }
}

// ---- stem col2im: dcols [n*ho*wo, 192] (column = ky*24 + kx*3 + c) -> float32 NCHW image gradient -----------
// One CTA per (image, input row iy).  Row iy receives from the <= 4 filter rows ky with the parity of iy + 3, i.e. from
// <= 4 output rows oy; their ky-runs (24 values per output pixel, 48 contiguous bytes per plane) are staged once in
// shared memory as floats (pitch 25 words: conflict-free for the strided gather), so every dcols element is read
// from HBM exactly once, in 16-byte pieces.  Each thread then sums the <= 16 taps of one (channel, ix).
struct Inv3 { float v[3]; };
constexpr int kStemK = 192;
constexpr int kRunPitch = 25;
template <bool F16>
__global__ void __launch_bounds__(kThreads) stem_col2im_kernel(const uint16_t* __restrict__ ch, const uint16_t* __restrict__ cl,
                                                                float* __restrict__ dx, int h, int w, int ho, int wo, Inv3 inv_std) {
  // runs[4][wo + 4][kRunPitch]: two zero columns on either side, so the gather below needs no bounds checks and every tap is a
  // compile-time offset from the thread's base (the first version computed kx ranges, bounds and three divisions per element at run
  // time: ncu showed it issue-bound at 70 %, 3.6x its HBM time)
  extern __shared__ float runs[];
  const int iy = blockIdx.x % h, im = blockIdx.x / h;
  const int ky0 = (iy + 1) & 1;                                  // iy = 2*oy - 3 + ky
  const int rowp = (wo + 4) * kRunPitch;
  for (int i = threadIdx.x; i < 4 * 4 * kRunPitch; i += kThreads) {           // the margins
    const int kyi = i / (4 * kRunPitch), r = i - kyi * 4 * kRunPitch, col = r / kRunPitch, e = r - col * kRunPitch;
    runs[kyi * rowp + (col < 2 ? col : wo + col) * kRunPitch + e] = 0.f;
  }
#pragma unroll
  for (int kyi = 0; kyi < 4; ++kyi) {
    const int ky = ky0 + 2 * kyi, oy = (iy + 3 - ky) >> 1;
    const bool valid = ky <= 6 && oy >= 0 && oy < ho;
    const size_t rbase = (((size_t)im * ho + (valid ? oy : 0)) * wo) * kStemK + ky * 24;
    for (int idx = threadIdx.x; idx < wo * 3; idx += kThreads) {
      const int ox = idx / 3, part = idx - ox * 3;
      float vals[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (valid) {
        const size_t base = rbase + (size_t)ox * kStemK + part * 8;
        load8<F16>(reinterpret_cast<const uint4*>(ch + base), reinterpret_cast<const uint4*>(cl + base), 0, vals);
      }
      float* dst = runs + kyi * rowp + (ox + 2) * kRunPitch + part * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = vals[j];
    }
  }
  __syncthreads();
  // thread = (channel, output column pair 2j, 2j + 1): even columns meet kx = 1, 3, 5 at ox = j + 1, j, j - 1; odd columns kx = 0, 2, 4, 6
  // at ox = j + 2, j + 1, j, j - 1
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    for (int j = threadIdx.x; j < wo; j += kThreads) {
      const float* b = runs + (j + 2) * kRunPitch + c;
      float even = 0.f, odd = 0.f;
#pragma unroll
      for (int kyi = 0; kyi < 4; ++kyi) {
        const float* r = b + kyi * rowp;
        even += r[1 * kRunPitch + 1 * 3] + r[0 * kRunPitch + 3 * 3] + r[-1 * kRunPitch + 5 * 3];
        odd += r[2 * kRunPitch + 0 * 3] + r[1 * kRunPitch + 2 * 3] + r[0 * kRunPitch + 4 * 3] + r[-1 * kRunPitch + 6 * 3];
      }
      *reinterpret_cast<float2*>(dx + (((size_t)im * 3 + c) * h + iy) * w + 2 * j) = make_float2(even * inv_std.v[c], odd * inv_std.v[c]);
    }
  }
}
}  // namespace

// ---- host side: one implementation per layer, f16 selects the single-plane instantiation -----------------------
static int relu_bwd_impl(const uint16_t* dy, const uint16_t* act, const uint16_t* add, uint16_t* out, size_t count, bool f16,
                         cudaStream_t s) {
  B200R_CHECK_ARG(dy && act && out, "null pointer");
  B200R_CHECK_ARG(count % 8 == 0, "count must be a multiple of 8");
  if (!count) return B200R_OK;
  const size_t c8 = count / 8;
  const uint4 *dh = reinterpret_cast<const uint4*>(dy), *dl = reinterpret_cast<const uint4*>(dy + count);
  const uint4* ah = reinterpret_cast<const uint4*>(act);
  const uint4 *bh = reinterpret_cast<const uint4*>(add), *bl = add ? reinterpret_cast<const uint4*>(add + count) : nullptr;
  uint4 *oh = reinterpret_cast<uint4*>(out), *ol = reinterpret_cast<uint4*>(out + count);
  if (f16) {
    if (add) relu_bwd_kernel<true, true><<<grid_for(c8), kThreads, 0, s>>>(dh, nullptr, ah, bh, nullptr, oh, nullptr, c8);
    else relu_bwd_kernel<false, true><<<grid_for(c8), kThreads, 0, s>>>(dh, nullptr, ah, nullptr, nullptr, oh, nullptr, c8);
  } else {
    if (add) relu_bwd_kernel<true, false><<<grid_for(c8), kThreads, 0, s>>>(dh, dl, ah, bh, bl, oh, ol, c8);
    else relu_bwd_kernel<false, false><<<grid_for(c8), kThreads, 0, s>>>(dh, dl, ah, nullptr, nullptr, oh, ol, c8);
  }
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

static int dilate2_impl(const uint16_t* x, uint16_t* y, int n, int h, int w, int c, bool f16, cudaStream_t s) {
  B200R_CHECK_ARG(x && y, "null pointer");
  B200R_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "c must be a multiple of 8");
  const size_t xin = (size_t)n * h * w * c, yout = xin * 4;
  if (f16)
    dilate2_kernel<true><<<grid_for(yout / 8), kThreads, 0, s>>>(reinterpret_cast<const uint4*>(x), nullptr, reinterpret_cast<uint4*>(y),
                                                                nullptr, n, h, w, c / 8);
  else
    dilate2_kernel<false><<<grid_for(yout / 8), kThreads, 0, s>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(x + xin),
                                                                 reinterpret_cast<uint4*>(y), reinterpret_cast<uint4*>(y + yout), n, h, w, c / 8);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

static int maxpool_bwd_impl(const uint16_t* x, const uint16_t* dy, uint16_t* dx, void* workspace, size_t ws_bytes, int n, int h, int w,
                            int c, bool f16, cudaStream_t s, bool relu_hi = false) {
  B200R_CHECK_ARG(x && dy && dx && workspace, "null pointer");
  B200R_CHECK_ARG(n > 0 && h > 0 && w > 0 && c % 8 == 0, "c must be a multiple of 8");
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  const size_t xin = (size_t)n * h * w * c, yout = (size_t)n * ho * wo * c;
  B200R_CHECK_ARG(ws_bytes >= yout, "workspace too small: need %zu bytes (one per pooled element)", yout);
  B200R_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "workspace must be 8-byte aligned");
  uint2* idx = reinterpret_cast<uint2*>(workspace);
  const uint4 *xh = reinterpret_cast<const uint4*>(x), *gh = reinterpret_cast<const uint4*>(dy);
  uint4* dh = reinterpret_cast<uint4*>(dx);
  if (relu_hi) {          // ReLU backward fused into the routing, one fp16 plane out
    if (f16) {
      maxpool_argmax_kernel<true, true><<<grid_for(yout / 8), kThreads, 0, s>>>(xh, nullptr, idx, n, h, w, c / 8, ho, wo);
      B200R_LAUNCH_CHECK();
      maxpool_bwd_kernel<true, true><<<grid_for(xin / 8), kThreads, 0, s>>>(idx, gh, nullptr, dh, nullptr, n, h, w, c / 8, ho, wo);
    } else {
      maxpool_argmax_kernel<false, true><<<grid_for(yout / 8), kThreads, 0, s>>>(xh, reinterpret_cast<const uint4*>(x + xin), idx, n, h, w, c / 8, ho, wo);
      B200R_LAUNCH_CHECK();
      maxpool_bwd_kernel<false, true><<<grid_for(xin / 8), kThreads, 0, s>>>(idx, gh, reinterpret_cast<const uint4*>(dy + yout), dh, nullptr, n, h, w, c / 8, ho, wo);
    }
    B200R_LAUNCH_CHECK();
    return B200R_OK;
  }
  if (f16) {
    maxpool_argmax_kernel<true><<<grid_for(yout / 8), kThreads, 0, s>>>(xh, nullptr, idx, n, h, w, c / 8, ho, wo);
    B200R_LAUNCH_CHECK();
    maxpool_bwd_kernel<true><<<grid_for(xin / 8), kThreads, 0, s>>>(idx, gh, nullptr, dh, nullptr, n, h, w, c / 8, ho, wo);
  } else {
    maxpool_argmax_kernel<false><<<grid_for(yout / 8), kThreads, 0, s>>>(xh, reinterpret_cast<const uint4*>(x + xin), idx, n, h, w, c / 8, ho, wo);
    B200R_LAUNCH_CHECK();
    maxpool_bwd_kernel<false><<<grid_for(xin / 8), kThreads, 0, s>>>(idx, gh, reinterpret_cast<const uint4*>(dy + yout), dh,
                                                                    reinterpret_cast<uint4*>(dx + xin), n, h, w, c / 8, ho, wo);
  }
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

static int avgpool_bwd_impl(const uint16_t* dy, uint16_t* dx, int n, int hw, int c, bool f16, cudaStream_t s) {
  B200R_CHECK_ARG(dy && dx, "null pointer");
  B200R_CHECK_ARG(n > 0 && hw > 0 && c % 8 == 0, "c must be a multiple of 8");
  const size_t yin = (size_t)n * c, xout = yin * hw;
  if (f16)
    avgpool_bwd_kernel<true><<<grid_for(xout / 8), kThreads, 0, s>>>(reinterpret_cast<const uint4*>(dy), nullptr, reinterpret_cast<uint4*>(dx),
                                                                    nullptr, n, hw, c / 8);
  else
    avgpool_bwd_kernel<false><<<grid_for(xout / 8), kThreads, 0, s>>>(reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(dy + yin),
                                                                     reinterpret_cast<uint4*>(dx), reinterpret_cast<uint4*>(dx + xout), n, hw, c / 8);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

static int col2im_impl(const uint16_t* dcols, float* dx, int n, int h, int w, const float* std_host, float unscale, bool f16,
                       cudaStream_t s) {
  B200R_CHECK_ARG(dcols && dx && std_host, "null pointer");
  B200R_CHECK_ARG(n > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, "bad shape");
  const int ho = h / 2, wo = w / 2;
  const size_t rows = (size_t)n * ho * wo;
  Inv3 inv;
  for (int i = 0; i < 3; ++i) inv.v[i] = unscale / std_host[i];
  const int smem = 4 * (wo + 4) * kRunPitch * (int)sizeof(float);
  B200R_CHECK_ARG(smem <= 200 * 1024, "image too wide for the staged col2im (w = %d)", w);
  static int configured[2] = {0, 0};
  if (configured[f16] < smem) {
    if (f16) B200R_CUDA(cudaFuncSetAttribute(stem_col2im_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    else B200R_CUDA(cudaFuncSetAttribute(stem_col2im_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[f16] = smem;
  }
  if (f16) stem_col2im_kernel<true><<<(unsigned)((size_t)n * h), kThreads, smem, s>>>(dcols, nullptr, dx, h, w, ho, wo, inv);
  else stem_col2im_kernel<false><<<(unsigned)((size_t)n * h), kThreads, smem, s>>>(dcols, dcols + rows * kStemK, dx, h, w, ho, wo, inv);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

extern "C" {

int b200r_relu_bwd(const uint16_t* dy, const uint16_t* act, const uint16_t* add, uint16_t* out, size_t count, b200r_stream_t stream) {
  return relu_bwd_impl(dy, act, add, out, count, false, as_stream(stream));
}
int b200r_relu_bwd_f16(const uint16_t* dy, const uint16_t* act, const uint16_t* add, uint16_t* out, size_t count, b200r_stream_t stream) {
  return relu_bwd_impl(dy, act, add, out, count, true, as_stream(stream));
}
int b200r_dilate2_nhwc(const uint16_t* x, uint16_t* y, int n, int h, int w, int c, b200r_stream_t stream) {
  return dilate2_impl(x, y, n, h, w, c, false, as_stream(stream));
}
int b200r_dilate2_nhwc_f16(const uint16_t* x, uint16_t* y, int n, int h, int w, int c, b200r_stream_t stream) {
  return dilate2_impl(x, y, n, h, w, c, true, as_stream(stream));
}
int b200r_maxpool3x3s2_bwd_codes_hi(const void* codes, const uint16_t* dy, uint16_t* dx_hi, int n, int h, int w, int c, int planes,
                                    b200r_stream_t stream) {
  B200R_CHECK_ARG(codes && dy && dx_hi, "null pointer");
  B200R_CHECK_ARG(n > 0 && h > 0 && w > 0 && c % 8 == 0, "c must be a multiple of 8");
  B200R_CHECK_ARG(planes == 1 || planes == 2, "planes must be 1 (fp16) or 2 (split)");
  B200R_CHECK_ARG((reinterpret_cast<uintptr_t>(codes) & 7) == 0, "codes must be 8-byte aligned");
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  const size_t xin = (size_t)n * h * w * c, yout = (size_t)n * ho * wo * c;
  const uint2* idx = reinterpret_cast<const uint2*>(codes);
  const uint4* gh = reinterpret_cast<const uint4*>(dy);
  uint4* dh = reinterpret_cast<uint4*>(dx_hi);
  if (planes == 1) maxpool_bwd_kernel<true, true><<<grid_for(xin / 8), kThreads, 0, as_stream(stream)>>>(idx, gh, nullptr, dh, nullptr, n, h, w, c / 8, ho, wo);
  else maxpool_bwd_kernel<false, true><<<grid_for(xin / 8), kThreads, 0, as_stream(stream)>>>(idx, gh, reinterpret_cast<const uint4*>(dy + yout), dh, nullptr, n, h, w, c / 8, ho, wo);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}
int b200r_maxpool3x3s2_relu_bwd_hi(const uint16_t* x, const uint16_t* dy, uint16_t* dx_hi, void* workspace, size_t ws_bytes, int n, int h,
                                   int w, int c, int planes, b200r_stream_t stream) {
  B200R_CHECK_ARG(planes == 1 || planes == 2, "planes must be 1 (fp16) or 2 (split)");
  return maxpool_bwd_impl(x, dy, dx_hi, workspace, ws_bytes, n, h, w, c, planes == 1, as_stream(stream), true);
}
int b200r_maxpool3x3s2_bwd_nhwc(const uint16_t* x, const uint16_t* dy, uint16_t* dx, void* workspace, size_t ws_bytes, int n, int h, int w,
                                int c, b200r_stream_t stream) {
  return maxpool_bwd_impl(x, dy, dx, workspace, ws_bytes, n, h, w, c, false, as_stream(stream));
}
int b200r_maxpool3x3s2_bwd_nhwc_f16(const uint16_t* x, const uint16_t* dy, uint16_t* dx, void* workspace, size_t ws_bytes, int n, int h,
                                    int w, int c, b200r_stream_t stream) {
  return maxpool_bwd_impl(x, dy, dx, workspace, ws_bytes, n, h, w, c, true, as_stream(stream));
}
int b200r_global_avgpool_bwd_nhwc(const uint16_t* dy, uint16_t* dx, int n, int hw, int c, b200r_stream_t stream) {
  return avgpool_bwd_impl(dy, dx, n, hw, c, false, as_stream(stream));
}
int b200r_global_avgpool_bwd_nhwc_f16(const uint16_t* dy, uint16_t* dx, int n, int hw, int c, b200r_stream_t stream) {
  return avgpool_bwd_impl(dy, dx, n, hw, c, true, as_stream(stream));
}
int b200r_stem_col2im_f32(const uint16_t* dcols, float* dx, int n, int h, int w, const float* std_host, b200r_stream_t stream) {
  return col2im_impl(dcols, dx, n, h, w, std_host, 1.0f, false, as_stream(stream));
}
int b200r_stem_col2im_f32_f16(const uint16_t* dcols, float* dx, int n, int h, int w, const float* std_host, float unscale,
                              b200r_stream_t stream) {
  return col2im_impl(dcols, dx, n, h, w, std_host, unscale, true, as_stream(stream));
}

}  // extern "C"
