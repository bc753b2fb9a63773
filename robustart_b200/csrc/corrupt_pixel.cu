// HBM-bound ImageNet-C corruptions: one pass over uint8 NHWC, 16 B / thread / access, no reuse.
//   gaussian_noise (corruptions.py:122-126)   speckle_noise (:143-147)   impulse_noise (:136-140)
//   shot_noise (:129-133)    brightness (:353-361)   saturate (:364-372)   contrast (:345-350)
//   frost (:244-262)         fog (:235-241, plasma_fractal :55-101)
// Algorithmic traffic: 150 528 B read + 150 528 B written per 224x224 image.
#include "corrupt.cuh"
#include <mutex>
#include <vector>
#include <string.h>

namespace {

constexpr int kThreads = 256;

// ---- small byte helpers ---------------------------------------------------------------------
// byte k of `w` as exact float without the XU-pipe I2F: 0x4B0000bb is 8388608 + bb.
__device__ __forceinline__ float byte_f(uint32_t w, int k) {
  uint32_t m = __byte_perm(w, 0x4B000000u, 0x7650u | (uint32_t)k);
  return __uint_as_float(m) - 8388608.0f;
}
// pack the low bytes of four words
__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}
__device__ __forceinline__ uint32_t f01_to_u8bits(float v01) {  // low byte = trunc(v*255)
  return __float_as_uint(__fmaf_rz(v01, 255.0f, 8388608.0f));
}

constexpr float kInv255 = 1.0f / 255.0f;

// =============================================================================================
// gaussian / speckle noise
// =============================================================================================
template <bool SPECKLE, bool EXT>
__global__ void __launch_bounds__(kThreads) normal_noise_kernel(const uint4* __restrict__ in,
                                                                 uint4* __restrict__ out,
                                                                 const float* __restrict__ ext,
                                                                 uint32_t groups_per_image, float c,
                                                                 uint32_t k0, uint32_t k1,
                                                                 uint64_t image_offset) {
  const uint32_t img = blockIdx.y;
  const uint32_t gi = blockIdx.x * kThreads + threadIdx.x;
  if (gi >= groups_per_image) return;
  const size_t g = (size_t)img * groups_per_image + gi;
  uint4 v = ld_stream_u4(in + g);
  float z[16];
  if (EXT) {
    const float4* e = reinterpret_cast<const float4*>(ext) + g * 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 t = ld_stream_f4(e + q);
      z[4 * q] = t.x * c; z[4 * q + 1] = t.y * c; z[4 * q + 2] = t.z * c; z[4 * q + 3] = t.w * c;
    }
  } else {
    const uint64_t gimg = image_offset + img;
#pragma unroll
    for (int call = 0; call < 2; ++call) {
      uint4 r = philox4x32_10(rng_counter(gi, SPECKLE ? RNG_SPECKLE : RNG_GAUSS, call, gimg), k0, k1);
      box_muller16(r.x, z[8 * call + 0], z[8 * call + 1], c);   // normals pre-scaled by sigma
      box_muller16(r.y, z[8 * call + 2], z[8 * call + 3], c);
      box_muller16(r.z, z[8 * call + 4], z[8 * call + 5], c);
      box_muller16(r.w, z[8 * call + 6], z[8 * call + 7], c);
    }
  }
  uint32_t wi[4] = {v.x, v.y, v.z, v.w}, wo[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float bf = byte_f(wi[q], k);
      const float zz = z[4 * q + k];
      // gaussian: x + c*z ; speckle: x + x*c*z  (x = b/255)
      const float r = SPECKLE ? fmaf(bf * kInv255, zz, bf * kInv255) : fmaf(bf, kInv255, zz);
      o[k] = f01_to_u8bits(__saturatef(r));
    }
    wo[q] = pack4(o[0], o[1], o[2], o[3]);
  }
  st_stream_u4(out + g, make_uint4(wo[0], wo[1], wo[2], wo[3]));
}

// =============================================================================================
// impulse noise (skimage random_noise 's&p': flipped = u1 < amount, salted = u2 < 0.5)
// ext layout: [n][2][P] uniforms
// =============================================================================================
template <bool EXT>
__global__ void __launch_bounds__(kThreads) impulse_kernel(const uint4* __restrict__ in,
                                                            uint4* __restrict__ out,
                                                            const float* __restrict__ ext,
                                                            uint32_t groups_per_image, float amount,
                                                            uint32_t k0, uint32_t k1,
                                                            uint64_t image_offset) {
  const uint32_t img = blockIdx.y;
  const uint32_t gi = blockIdx.x * kThreads + threadIdx.x;
  if (gi >= groups_per_image) return;
  const size_t g = (size_t)img * groups_per_image + gi;
  uint4 v = ld_stream_u4(in + g);
  uint32_t wi[4] = {v.x, v.y, v.z, v.w}, wo[4];
  const uint32_t thr = (uint32_t)(amount * 16777216.0f);
  const size_t P = (size_t)groups_per_image * 16;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    bool flip[4], salt[4];
    if (EXT) {
      const float* e = ext + (size_t)img * 2 * P + (size_t)gi * 16 + 4 * q;
      float4 uf = ld_stream_f4(e), us = ld_stream_f4(e + P);
      flip[0] = uf.x < amount; flip[1] = uf.y < amount; flip[2] = uf.z < amount; flip[3] = uf.w < amount;
      salt[0] = us.x < 0.5f; salt[1] = us.y < 0.5f; salt[2] = us.z < 0.5f; salt[3] = us.w < 0.5f;
    } else {
      uint4 r = philox4x32_10(rng_counter(gi, RNG_IMPULSE, q, image_offset + img), k0, k1);
      uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { flip[k] = (rr[k] >> 8) < thr; salt[k] = rr[k] & 1u; }
    }
    uint32_t w = wi[q];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (flip[k]) {
        uint32_t mask = 0xFFu << (8 * k);
        w = salt[k] ? (w | mask) : (w & ~mask);
      }
    }
    wo[q] = w;
  }
  st_stream_u4(out + g, make_uint4(wo[0], wo[1], wo[2], wo[3]));
}

// =============================================================================================
// shot noise: k ~ Poisson(b/255*c); out = trunc(min(k/c,1)*255).
// Exact Poisson by Walker/Vose alias tables: one table of K=128 entries for each of the 256
// possible pixel values, resident in shared memory (128 KB); one 32-bit random word and ONE
// shared-memory lookup per output byte.  entry = (threshold24 << 8) | alias.
// ext layout: [n][P] caller-drawn Poisson counts.
// =============================================================================================
constexpr int kShotK = 128;
constexpr int kShotThreads = 512;

struct ShotTables {
  uint32_t* d_alias = nullptr;  // [256][128]
  uint8_t* d_lut = nullptr;     // [128] count -> output byte
};

template <bool EXT>
__global__ void __launch_bounds__(kShotThreads, 1) shot_kernel(const uint4* __restrict__ in,
                                                                uint4* __restrict__ out,
                                                                const float* __restrict__ ext,
                                                                const uint32_t* __restrict__ alias_g,
                                                                const uint8_t* __restrict__ lut_g,
                                                                uint32_t groups_per_image,
                                                                uint32_t n_images, uint32_t k0,
                                                                uint32_t k1, uint64_t image_offset) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  uint32_t* s_alias = smem_u32;                                    // 256*128
  uint8_t* s_lut = reinterpret_cast<uint8_t*>(smem_u32 + 256 * kShotK);  // 128
  if (!EXT) {
    const uint4* src = reinterpret_cast<const uint4*>(alias_g);
    uint4* dst = reinterpret_cast<uint4*>(s_alias);
    for (int i = threadIdx.x; i < 256 * kShotK / 4; i += kShotThreads) dst[i] = src[i];
  }
  if (threadIdx.x < kShotK) s_lut[threadIdx.x] = lut_g[threadIdx.x];
  __syncthreads();
  const size_t total = (size_t)groups_per_image * n_images;
  for (size_t g = (size_t)blockIdx.x * kShotThreads + threadIdx.x; g < total;
       g += (size_t)gridDim.x * kShotThreads) {
    uint4 v = ld_stream_u4(in + g);
    uint32_t wi[4] = {v.x, v.y, v.z, v.w}, wo[4];
    const uint32_t img = (uint32_t)(g / groups_per_image);
    const uint32_t gi = (uint32_t)(g - (size_t)img * groups_per_image);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t kk[4];
      if (EXT) {
        float4 e = ld_stream_f4(ext + g * 16 + 4 * q);
        kk[0] = (uint32_t)e.x; kk[1] = (uint32_t)e.y; kk[2] = (uint32_t)e.z; kk[3] = (uint32_t)e.w;
      } else {
        uint4 r = philox4x32_10(rng_counter(gi, RNG_SHOT, q, image_offset + img), k0, k1);
        uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t b = (wi[q] >> (8 * k)) & 0xFFu;
          uint32_t idx = rr[k] & (kShotK - 1);
          uint32_t e = s_alias[b * kShotK + idx];
          kk[k] = ((rr[k] >> 8) < (e >> 8)) ? idx : (e & 0xFFu);
        }
      }
      uint32_t o = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) o |= (uint32_t)s_lut[min(kk[k], (uint32_t)(kShotK - 1))] << (8 * k);
      wo[q] = o;
    }
    st_stream_u4(out + g, make_uint4(wo[0], wo[1], wo[2], wo[3]));
  }
}

std::mutex g_shot_mu;
ShotTables g_shot[8][5];  // [device][severity-1]

// Vose alias construction in double precision for one lambda.
void build_alias(double lam, uint32_t* entry /*[K]*/) {
  const int K = kShotK;
  double p[K];
  // pmf by recurrence in log space for stability
  double sum = 0;
  for (int k = 0; k < K; ++k) {
    p[k] = (lam == 0.0) ? (k == 0 ? 1.0 : 0.0) : exp(k * log(lam) - lam - lgamma(k + 1.0));
    sum += p[k];
  }
  double scaled[K];
  int small[K], large[K], ns = 0, nl = 0;
  double prob[K];
  int alias[K];
  for (int k = 0; k < K; ++k) {
    scaled[k] = p[k] / sum * K;
    prob[k] = 1.0;
    alias[k] = k;
  }
  for (int k = 0; k < K; ++k) (scaled[k] < 1.0 ? small[ns++] : large[nl++]) = k;
  while (ns > 0 && nl > 0) {
    int s = small[--ns], l = large[--nl];
    prob[s] = scaled[s];
    alias[s] = l;
    scaled[l] = (scaled[l] + scaled[s]) - 1.0;
    (scaled[l] < 1.0 ? small[ns++] : large[nl++]) = l;
  }
  for (int k = 0; k < K; ++k) {
    double t = prob[k] * 16777216.0;
    uint32_t thr = t >= 16777215.0 ? 0xFFFFFFu : (uint32_t)t;
    if (prob[k] >= 1.0) thr = 0xFFFFFFu, alias[k] = k;  // never take the alias (t < thr fails only at 2^24-1: alias==k anyway)
    entry[k] = (thr << 8) | (uint32_t)alias[k];
  }
}

int get_shot_tables(int severity, ShotTables* out) {
  static const int cs[5] = {60, 25, 12, 5, 3};
  int dev = 0;
  B200R_CUDA(cudaGetDevice(&dev));
  B200R_CHECK_ARG(dev >= 0 && dev < 8, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_shot_mu);
  ShotTables& t = g_shot[dev][severity - 1];
  if (!t.d_alias) {
    const double c = cs[severity - 1];
    std::vector<uint32_t> h_alias(256 * kShotK);
    for (int b = 0; b < 256; ++b) build_alias((b / 255.0) * c, &h_alias[b * kShotK]);
    uint8_t h_lut[kShotK];
    for (int k = 0; k < kShotK; ++k) {
      double v = k / c;                       // np.clip(k / float(c), 0, 1) * 255 -> uint8 truncation
      v = v < 0 ? 0 : (v > 1 ? 1 : v);
      h_lut[k] = (uint8_t)(v * 255);
    }
    B200R_CUDA(cudaMalloc(&t.d_alias, h_alias.size() * 4));
    B200R_CUDA(cudaMalloc(&t.d_lut, kShotK));
    B200R_CUDA(cudaMemcpy(t.d_alias, h_alias.data(), h_alias.size() * 4, cudaMemcpyHostToDevice));
    B200R_CUDA(cudaMemcpy(t.d_lut, h_lut, kShotK, cudaMemcpyHostToDevice));
  }
  *out = t;
  return B200R_OK;
}

// =============================================================================================
// brightness / saturate: HSV round trip per pixel, 16 pixels (48 B) per thread.
// skimage 0.17 rgb2hsv/hsv2rgb operation order (ties: blue > green > red).
// =============================================================================================
__device__ __forceinline__ void hsv_adjust(float r, float g, float b, int mode, float p0, float p1,
                                           float& ro, float& go, float& bo) {
  float v = fmaxf(r, fmaxf(g, b));
  float mn = fminf(r, fminf(g, b));
  float delta = v - mn;
  float s = 0.f, h = 0.f;
  if (delta != 0.f) {
    s = delta / v;
    float h6;
    if (b == v) h6 = 4.f + (r - g) / delta;
    else if (g == v) h6 = 2.f + (b - r) / delta;
    else h6 = (g - b) / delta;
    h = h6 / 6.f;
    h = h - floorf(h);  // % 1.0
  }
  if (mode == 0) v = __saturatef(v + p0);            // brightness: V = clip(V + c, 0, 1)
  else s = __saturatef(fmaf(s, p0, p1));             // saturate:   S = clip(S*c0 + c1, 0, 1)
  float h6 = h * 6.f;
  float hi = floorf(h6);
  float f = h6 - hi;
  float p = v * (1.f - s);
  float q = v * (1.f - f * s);
  float t = v * (1.f - (1.f - f) * s);
  int sel = ((int)hi) % 6;
  switch (sel) {
    case 0: ro = v; go = t; bo = p; break;
    case 1: ro = q; go = v; bo = p; break;
    case 2: ro = p; go = v; bo = t; break;
    case 3: ro = p; go = q; bo = v; break;
    case 4: ro = t; go = p; bo = v; break;
    default: ro = v; go = p; bo = q; break;
  }
}

__global__ void __launch_bounds__(kThreads) hsv_kernel(const uint4* __restrict__ in,
                                                        uint4* __restrict__ out, size_t groups48,
                                                        int mode, float p0, float p1) {
  const size_t g = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (g >= groups48) return;
  uint4 a = ld_stream_u4(in + 3 * g), b = ld_stream_u4(in + 3 * g + 1), c = ld_stream_u4(in + 3 * g + 2);
  uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
  uint32_t ob[48];
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    float ch[3], o[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int byte = 3 * p + k;
      ch[k] = byte_f(w[byte >> 2], byte & 3) * kInv255;
    }
    hsv_adjust(ch[0], ch[1], ch[2], mode, p0, p1, o[0], o[1], o[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) ob[3 * p + k] = f01_to_u8bits(__saturatef(o[k]));
  }
  uint32_t wo[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) wo[q] = pack4(ob[4 * q], ob[4 * q + 1], ob[4 * q + 2], ob[4 * q + 3]);
  st_stream_u4(out + 3 * g, make_uint4(wo[0], wo[1], wo[2], wo[3]));
  st_stream_u4(out + 3 * g + 1, make_uint4(wo[4], wo[5], wo[6], wo[7]));
  st_stream_u4(out + 3 * g + 2, make_uint4(wo[8], wo[9], wo[10], wo[11]));
}

// =============================================================================================
// contrast: per-image per-channel mean (exact integer sums), then affine + clip.
// ws: uint32 [n][4] channel sums (index 3 unused)
// =============================================================================================
__global__ void __launch_bounds__(kThreads) channel_sum_kernel(const uint4* __restrict__ in,
                                                                uint32_t* __restrict__ sums,
                                                                uint32_t groups48_per_image) {
  const uint32_t img = blockIdx.y;
  uint32_t s[3] = {0, 0, 0};
  for (uint32_t gi = blockIdx.x * kThreads + threadIdx.x; gi < groups48_per_image;
       gi += gridDim.x * kThreads) {
    const uint4* p = in + ((size_t)img * groups48_per_image + gi) * 3;
    uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);   // keep in L2 for the apply pass
    uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
    for (int byte = 0; byte < 48; ++byte) s[byte % 3] += (w[byte >> 2] >> (8 * (byte & 3))) & 0xFFu;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) atomicAdd(&sums[img * 4 + k], s[k]);
  }
}

__global__ void __launch_bounds__(kThreads) contrast_apply_kernel(const uint4* __restrict__ in,
                                                                   uint4* __restrict__ out,
                                                                   const uint32_t* __restrict__ sums,
                                                                   uint32_t groups48_per_image,
                                                                   float c, float inv_count255) {
  const uint32_t img = blockIdx.y;
  const uint32_t gi = blockIdx.x * kThreads + threadIdx.x;
  if (gi >= groups48_per_image) return;
  float m[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) m[k] = (float)((double)sums[img * 4 + k] * (double)inv_count255);
  const size_t g = (size_t)img * groups48_per_image + gi;
  uint4 a = ld_stream_u4(in + 3 * g), b = ld_stream_u4(in + 3 * g + 1), cc = ld_stream_u4(in + 3 * g + 2);
  uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, cc.x, cc.y, cc.z, cc.w};
  uint32_t ob[48];
#pragma unroll
  for (int byte = 0; byte < 48; ++byte) {
    float x = byte_f(w[byte >> 2], byte & 3) * kInv255;
    float mm = m[byte % 3];
    ob[byte] = f01_to_u8bits(__saturatef(fmaf(x - mm, c, mm)));
  }
  uint32_t wo[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) wo[q] = pack4(ob[4 * q], ob[4 * q + 1], ob[4 * q + 2], ob[4 * q + 3]);
  st_stream_u4(out + 3 * g, make_uint4(wo[0], wo[1], wo[2], wo[3]));
  st_stream_u4(out + 3 * g + 1, make_uint4(wo[4], wo[5], wo[6], wo[7]));
  st_stream_u4(out + 3 * g + 2, make_uint4(wo[8], wo[9], wo[10], wo[11]));
}

// =============================================================================================
// frost: out = trunc(clip(c0*x + c1*tex[xs+y][ys+x], 0, 255))
// ext layout: [n][3] = texture index, x_start (row), y_start (col)
// =============================================================================================
struct FrostTex { const uint8_t* p; int th, tw; };
struct FrostTexSet { FrostTex t[6]; };
FrostTexSet g_frost[8];
std::mutex g_frost_mu;

__device__ __forceinline__ uint32_t bounded_u32(uint32_t r, uint32_t range) {  // [0, range)
  return (uint32_t)(((uint64_t)r * range) >> 32);
}

__global__ void __launch_bounds__(kThreads) frost_kernel(const uint4* __restrict__ in,
                                                          uint4* __restrict__ out,
                                                          const float* __restrict__ ext, FrostTexSet ts,
                                                          int h, int w, float c0, float c1, uint32_t k0,
                                                          uint32_t k1, uint64_t image_offset) {
  const uint32_t img = blockIdx.y;
  const uint32_t groups48_per_image = (uint32_t)(h * w) / 16;
  const uint32_t gi = blockIdx.x * kThreads + threadIdx.x;
  if (gi >= groups48_per_image) return;
  int idx, xs, ys;
  if (ext) {
    idx = (int)ext[img * 3]; xs = (int)ext[img * 3 + 1]; ys = (int)ext[img * 3 + 2];
  } else {
    uint4 r = philox4x32_10(rng_counter(0, RNG_FROST, 0, image_offset + img), k0, k1);
    idx = (int)bounded_u32(r.x, 5);                      // np.random.randint(5): files 1..5 only
    xs = (int)bounded_u32(r.y, (uint32_t)(ts.t[idx].th - h));
    ys = (int)bounded_u32(r.z, (uint32_t)(ts.t[idx].tw - w));
  }
  const FrostTex tx = ts.t[idx];
  const size_t g = (size_t)img * groups48_per_image + gi;
  uint4 a = ld_stream_u4(in + 3 * g), b = ld_stream_u4(in + 3 * g + 1), cc = ld_stream_u4(in + 3 * g + 2);
  uint32_t wv[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, cc.x, cc.y, cc.z, cc.w};
  const int pix0 = gi * 16;
  const int y = pix0 / w, x0 = pix0 - y * w;  // w % 16 == 0 so the 16 pixels share a row
  const uint8_t* trow = tx.p + ((size_t)(xs + y) * tx.tw + (ys + x0)) * 3;
  uint32_t ob[48];
#pragma unroll
  for (int byte = 0; byte < 48; ++byte) {
    float x = byte_f(wv[byte >> 2], byte & 3);
    float t = (float)__ldg(trow + byte);
    float v = fminf(fmaxf(fmaf(c0, x, c1 * t), 0.f), 255.f);
    ob[byte] = __float_as_uint(__fadd_rz(v, 8388608.0f));
  }
  uint32_t wo[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) wo[q] = pack4(ob[4 * q], ob[4 * q + 1], ob[4 * q + 2], ob[4 * q + 3]);
  st_stream_u4(out + 3 * g, make_uint4(wo[0], wo[1], wo[2], wo[3]));
  st_stream_u4(out + 3 * g + 1, make_uint4(wo[4], wo[5], wo[6], wo[7]));
  st_stream_u4(out + 3 * g + 2, make_uint4(wo[8], wo[9], wo[10], wo[11]));
}

// =============================================================================================
// fog: diamond-square plasma (256x256, 8 levels, 65 535 uniforms per image), per-image max,
// then blend.  ws per image: float map[65536] + float stats[4] = {map_min, map_range, x_max, -}
// ext layout: [n][65535] uniforms in [0,1) in the reference's draw order.
// =============================================================================================
constexpr int kMap = 256;
constexpr int kFogThreads = 1024;

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sbuf) {
  v = is_max ? warp_max(v) : -warp_max(-v);
  if ((threadIdx.x & 31) == 0) sbuf[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sbuf[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = is_max ? fmaxf(r, sbuf[i]) : fminf(r, sbuf[i]);
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kFogThreads, 1) fog_plasma_kernel(const uint8_t* __restrict__ in,
                                                                     float* __restrict__ ws,
                                                                     const float* __restrict__ ext,
                                                                     int image_bytes, float decay,
                                                                     uint32_t k0, uint32_t k1,
                                                                     uint64_t image_offset) {
  __shared__ float sbuf[32];
  const uint32_t img = blockIdx.x;
  float* map = ws + (size_t)img * (kMap * kMap + 4);
  float* stats = map + kMap * kMap;
  const float* e = ext ? ext + (size_t)img * 65535 : nullptr;
  const uint64_t gimg = image_offset + img;
  auto draw = [&](uint32_t i) -> float {
    if (e) return e[i];
    uint4 r = philox4x32_10(rng_counter(i >> 2, RNG_FOG, 0, gimg), k0, k1);
    uint32_t x = (i & 3) == 0 ? r.x : (i & 3) == 1 ? r.y : (i & 3) == 2 ? r.z : r.w;
    return u32_to_unit(x);
  };
  if (threadIdx.x == 0) map[0] = 0.f;
  __syncthreads();
  float wibble = 100.f;
  uint32_t base = 0;
  for (int s = kMap; s >= 2; s >>= 1) {
    const int m = kMap / s, half = s / 2;
    const uint32_t mm = (uint32_t)m * m;
    // squares
    for (uint32_t t = threadIdx.x; t < mm; t += kFogThreads) {
      int i = t / m, j = t - i * m;
      int i1 = (i + 1) % m, j1 = (j + 1) % m;
      float acc = (map[(i * s) * kMap + j * s] + map[(i1 * s) * kMap + j * s]) +
                  (map[(i * s) * kMap + j1 * s] + map[(i1 * s) * kMap + j1 * s]);
      float u = draw(base + t);
      map[(i * s + half) * kMap + j * s + half] = acc * 0.25f + wibble * (-wibble + 2.f * wibble * u);
    }
    __syncthreads();
    // diamonds
    for (uint32_t t = threadIdx.x; t < 2 * mm; t += kFogThreads) {
      const bool second = t >= mm;
      const uint32_t tt = second ? t - mm : t;
      int i = tt / m, j = tt - i * m;
      auto dr = [&](int a, int b) { return map[(((a + m) % m) * s + half) * kMap + ((b + m) % m) * s + half]; };
      auto ul = [&](int a, int b) { return map[(((a + m) % m) * s) * kMap + ((b + m) % m) * s]; };
      float u = draw(base + mm + t);
      float noise = wibble * (-wibble + 2.f * wibble * u);
      if (!second) {
        float acc = (dr(i, j) + dr(i - 1, j)) + (ul(i, j) + ul(i, j + 1));
        map[(i * s) * kMap + j * s + half] = acc * 0.25f + noise;
      } else {
        float acc = (dr(i, j) + dr(i, j - 1)) + (ul(i, j) + ul(i + 1, j));
        map[(i * s + half) * kMap + j * s] = acc * 0.25f + noise;
      }
    }
    __syncthreads();
    base += 3 * mm;
    wibble /= decay;
  }
  float lo = 3.4e38f, hi = -3.4e38f;
  for (int t = threadIdx.x; t < kMap * kMap; t += kFogThreads) {
    float v = map[t];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  lo = block_reduce(lo, false, sbuf);
  hi = block_reduce(hi, true, sbuf);
  // image max (uint8, exact)
  uint32_t mx = 0;
  const uint4* ip = reinterpret_cast<const uint4*>(in + (size_t)img * image_bytes);
  for (int t = threadIdx.x; t < image_bytes / 16; t += kFogThreads) {
    uint4 v = __ldg(ip + t);
    uint32_t m4 = __vmaxu4(__vmaxu4(v.x, v.y), __vmaxu4(v.z, v.w));
    mx = max(mx, max(max(m4 & 0xFF, (m4 >> 8) & 0xFF), max((m4 >> 16) & 0xFF, m4 >> 24)));
  }
  float fm = block_reduce((float)mx, true, sbuf);
  if (threadIdx.x == 0) { stats[0] = lo; stats[1] = hi - lo; stats[2] = fm * kInv255; }
}

__global__ void __launch_bounds__(kThreads) fog_blend_kernel(const uint4* __restrict__ in,
                                                              uint4* __restrict__ out,
                                                              const float* __restrict__ ws, int h, int w,
                                                              float c0) {
  const uint32_t img = blockIdx.y;
  const uint32_t groups48_per_image = (uint32_t)(h * w) / 16;
  const uint32_t gi = blockIdx.x * kThreads + threadIdx.x;
  if (gi >= groups48_per_image) return;
  const float* map = ws + (size_t)img * (kMap * kMap + 4);
  const float lo = map[kMap * kMap], range = map[kMap * kMap + 1], maxv = map[kMap * kMap + 2];
  const float gain = maxv / (maxv + c0);
  const float inv_range = 1.0f / range;
  const size_t g = (size_t)img * groups48_per_image + gi;
  uint4 a = ld_stream_u4(in + 3 * g), b = ld_stream_u4(in + 3 * g + 1), cc = ld_stream_u4(in + 3 * g + 2);
  uint32_t wv[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, cc.x, cc.y, cc.z, cc.w};
  const int pix0 = gi * 16;
  const int y = pix0 / w, x0 = pix0 - y * w;
  const float* mrow = map + y * kMap + x0;
  uint32_t ob[48];
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    float pl = c0 * ((mrow[p] - lo) * inv_range);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int byte = 3 * p + k;
      float x = byte_f(wv[byte >> 2], byte & 3) * kInv255;
      ob[byte] = f01_to_u8bits(__saturatef((x + pl) * gain));
    }
  }
  uint32_t wo[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) wo[q] = pack4(ob[4 * q], ob[4 * q + 1], ob[4 * q + 2], ob[4 * q + 3]);
  st_stream_u4(out + 3 * g, make_uint4(wo[0], wo[1], wo[2], wo[3]));
  st_stream_u4(out + 3 * g + 1, make_uint4(wo[4], wo[5], wo[6], wo[7]));
  st_stream_u4(out + 3 * g + 2, make_uint4(wo[8], wo[9], wo[10], wo[11]));
}

}  // namespace

// ---------------------------------------------------------------------------------------------
extern "C" int b200r_set_frost_texture(int slot, const uint8_t* rgb, int th, int tw) {
  B200R_CHECK_ARG(slot >= 0 && slot < 6, "frost slot %d not in [0,6)", slot);
  B200R_CHECK_ARG(rgb && th > 0 && tw > 0, "bad frost texture");
  int dev = 0;
  B200R_CUDA(cudaGetDevice(&dev));
  B200R_CHECK_ARG(dev >= 0 && dev < 8, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_frost_mu);
  g_frost[dev].t[slot] = FrostTex{rgb, th, tw};
  return B200R_OK;
}

size_t corrupt_pixel_ws(int id, int sev, int n, int h, int w) {
  (void)sev; (void)h; (void)w;
  switch (id) {
    case B200R_CONTRAST: return (size_t)n * 16;
    case B200R_FOG: return (size_t)n * (kMap * kMap + 4) * sizeof(float);
    default: return 0;
  }
}

int corrupt_pixel_family(const CorruptArgs& a) {
  const size_t P = (size_t)a.h * a.w * 3;
  B200R_CHECK_ARG((a.h * a.w) % 16 == 0 && a.w % 16 == 0,
                  "h*w and w must be multiples of 16 (got %dx%d)", a.h, a.w);
  const uint32_t gpi = (uint32_t)(P / 16);   // 16-byte groups per image
  const uint32_t g48 = (uint32_t)(P / 48);   // 16-pixel groups per image
  const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
  const uint4* in = reinterpret_cast<const uint4*>(a.in);
  uint4* out = reinterpret_cast<uint4*>(a.out);
  const int s = a.severity - 1;
  dim3 grid16((gpi + kThreads - 1) / kThreads, a.n), grid48((g48 + kThreads - 1) / kThreads, a.n);
  switch (a.id) {
    case B200R_GAUSSIAN_NOISE: {
      static const float c[5] = {.08f, .12f, 0.18f, 0.26f, 0.38f};
      if (a.ext) normal_noise_kernel<false, true><<<grid16, kThreads, 0, a.stream>>>(in, out, a.ext, gpi, c[s], k0, k1, a.image_offset);
      else normal_noise_kernel<false, false><<<grid16, kThreads, 0, a.stream>>>(in, out, nullptr, gpi, c[s], k0, k1, a.image_offset);
      break;
    }
    case B200R_SPECKLE_NOISE: {
      static const float c[5] = {.15f, .2f, 0.35f, 0.45f, 0.6f};
      if (a.ext) normal_noise_kernel<true, true><<<grid16, kThreads, 0, a.stream>>>(in, out, a.ext, gpi, c[s], k0, k1, a.image_offset);
      else normal_noise_kernel<true, false><<<grid16, kThreads, 0, a.stream>>>(in, out, nullptr, gpi, c[s], k0, k1, a.image_offset);
      break;
    }
    case B200R_IMPULSE_NOISE: {
      static const float c[5] = {.03f, .06f, .09f, 0.17f, 0.27f};
      if (a.ext) impulse_kernel<true><<<grid16, kThreads, 0, a.stream>>>(in, out, a.ext, gpi, c[s], k0, k1, a.image_offset);
      else impulse_kernel<false><<<grid16, kThreads, 0, a.stream>>>(in, out, nullptr, gpi, c[s], k0, k1, a.image_offset);
      break;
    }
    case B200R_SHOT_NOISE: {
      ShotTables t;
      int rc = get_shot_tables(a.severity, &t);
      if (rc) return rc;
      const size_t smem = 256 * kShotK * 4 + kShotK;
      const int grid = b200r_num_sms();
      if (a.ext) {
        B200R_CUDA(cudaFuncSetAttribute(shot_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        shot_kernel<true><<<grid, kShotThreads, smem, a.stream>>>(in, out, a.ext, t.d_alias, t.d_lut, gpi, a.n, k0, k1, a.image_offset);
      } else {
        B200R_CUDA(cudaFuncSetAttribute(shot_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        shot_kernel<false><<<grid, kShotThreads, smem, a.stream>>>(in, out, nullptr, t.d_alias, t.d_lut, gpi, a.n, k0, k1, a.image_offset);
      }
      break;
    }
    case B200R_BRIGHTNESS: {
      static const float c[5] = {.1f, .2f, .3f, .4f, .5f};
      const size_t groups = (size_t)g48 * a.n;
      hsv_kernel<<<(unsigned)((groups + kThreads - 1) / kThreads), kThreads, 0, a.stream>>>(in, out, groups, 0, c[s], 0.f);
      break;
    }
    case B200R_SATURATE: {
      static const float c[5][2] = {{0.3f, 0}, {0.1f, 0}, {2, 0}, {5, 0.1f}, {20, 0.2f}};
      const size_t groups = (size_t)g48 * a.n;
      hsv_kernel<<<(unsigned)((groups + kThreads - 1) / kThreads), kThreads, 0, a.stream>>>(in, out, groups, 1, c[s][0], c[s][1]);
      break;
    }
    case B200R_CONTRAST: {
      static const float c[5] = {0.4f, .3f, .2f, .1f, .05f};
      B200R_CHECK_ARG(a.ws && a.ws_bytes >= (size_t)a.n * 16, "contrast needs %zu workspace bytes", (size_t)a.n * 16);
      uint32_t* sums = static_cast<uint32_t*>(a.ws);
      B200R_CUDA(cudaMemsetAsync(sums, 0, (size_t)a.n * 16, a.stream));
      dim3 gs(4, a.n);
      channel_sum_kernel<<<gs, kThreads, 0, a.stream>>>(in, sums, g48);
      contrast_apply_kernel<<<grid48, kThreads, 0, a.stream>>>(in, out, sums, g48, c[s], (float)(1.0 / (255.0 * a.h * a.w)));
      break;
    }
    case B200R_FROST: {
      static const float c[5][2] = {{1, 0.4f}, {0.8f, 0.6f}, {0.7f, 0.7f}, {0.65f, 0.7f}, {0.6f, 0.75f}};
      int dev = 0;
      B200R_CUDA(cudaGetDevice(&dev));
      FrostTexSet ts;
      {
        std::lock_guard<std::mutex> lk(g_frost_mu);
        ts = g_frost[dev];
      }
      for (int i = 0; i < 5; ++i)
        B200R_CHECK_ARG(ts.t[i].p && ts.t[i].th > a.h && ts.t[i].tw > a.w,
                        "frost texture %d not set (b200r_set_frost_texture) or not larger than the image", i);
      frost_kernel<<<grid48, kThreads, 0, a.stream>>>(in, out, a.ext, ts, a.h, a.w, c[s][0], c[s][1], k0, k1, a.image_offset);
      break;
    }
    case B200R_FOG: {
      static const float c[5][2] = {{1.5f, 2}, {2.f, 2}, {2.5f, 1.7f}, {2.5f, 1.5f}, {3.f, 1.4f}};
      B200R_CHECK_ARG(a.h <= kMap && a.w <= kMap, "fog supports images up to 256x256");
      const size_t need = corrupt_pixel_ws(a.id, a.severity, a.n, a.h, a.w);
      B200R_CHECK_ARG(a.ws && a.ws_bytes >= need, "fog needs %zu workspace bytes", need);
      float* ws = static_cast<float*>(a.ws);
      fog_plasma_kernel<<<a.n, kFogThreads, 0, a.stream>>>(a.in, ws, a.ext, (int)P, c[s][1], k0, k1, a.image_offset);
      fog_blend_kernel<<<grid48, kThreads, 0, a.stream>>>(in, out, ws, a.h, a.w, c[s][0]);
      break;
    }
    default:
      b200r_set_error("corruption id %d is not in the pixel family", a.id);
      return B200R_EINVAL;
  }
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}
