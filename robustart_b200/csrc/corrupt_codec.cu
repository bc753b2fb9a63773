// Integer "codec" corruptions: bit-exact restatements of the library code the reference calls.
//   pixelate (corruptions.py:385-391): PIL Image.resize(BOX) down then up  == Pillow libImaging/Resample.c
//            (precompute_coeffs + normalize_coeffs_8bpc + ImagingResampleHorizontal/Vertical_8bpc):
//            coefficients in 22-bit fixed point, horizontal pass then vertical pass, uint8 in between.
//   jpeg_compression (:375-382): not implemented yet (B200R_ENOTSUP).
#include "corrupt.cuh"
#include <vector>
#include <mutex>
#include <cmath>

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Resample.c PRECISION_BITS
constexpr int kMaxK = 8;

struct ResampleTable {  // one 1-D pass: out[xx] = clip8((2^21 + sum_k in[xmin+k]*coef[k]) >> 22)
  int out_size, ksize;
  std::vector<int> xmin, xcnt, coef;  // coef: [out_size][ksize]
};

// Resample.c precompute_coeffs() with the BOX filter (support 0.5) + normalize_coeffs_8bpc()
ResampleTable make_box_table(int in_size, int out_size) {
  ResampleTable t;
  t.out_size = out_size;
  double scale = (double)in_size / out_size, filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 0.5 * filterscale;
  t.ksize = (int)ceil(support) * 2 + 1;
  t.xmin.resize(out_size); t.xcnt.resize(out_size); t.coef.assign((size_t)out_size * t.ksize, 0);
  std::vector<double> k(t.ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale, ss = 1.0 / filterscale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double a = (x + xmin - center + 0.5) * ss;
      const double w = (a > -0.5 && a <= 0.5) ? 1.0 : 0.0;
      k[x] = w; ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      const double v = k[x] * (1 << kPrecisionBits);
      t.coef[(size_t)xx * t.ksize + x] = (int)(v < 0 ? -0.5 + v : 0.5 + v);
    }
    t.xmin[xx] = xmin; t.xcnt[xx] = xmax;
  }
  return t;
}

struct DevTable { int* xmin; int* xcnt; int* coef; int out_size, ksize; };
struct PixelateTables { DevTable down, up; bool ready = false; int small; };
PixelateTables g_pix[8][5];
std::mutex g_pix_mu;

int upload(const ResampleTable& t, DevTable& d) {
  d.out_size = t.out_size; d.ksize = t.ksize;
  B200R_CUDA(cudaMalloc(&d.xmin, t.xmin.size() * 4));
  B200R_CUDA(cudaMalloc(&d.xcnt, t.xcnt.size() * 4));
  B200R_CUDA(cudaMalloc(&d.coef, t.coef.size() * 4));
  B200R_CUDA(cudaMemcpy(d.xmin, t.xmin.data(), t.xmin.size() * 4, cudaMemcpyHostToDevice));
  B200R_CUDA(cudaMemcpy(d.xcnt, t.xcnt.data(), t.xcnt.size() * 4, cudaMemcpyHostToDevice));
  B200R_CUDA(cudaMemcpy(d.coef, t.coef.data(), t.coef.size() * 4, cudaMemcpyHostToDevice));
  return B200R_OK;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return (uint8_t)min(max(v, 0), 255);
}

constexpr int kPixThreads = 1024;

// one CTA per image.  src (global) [h][w][3] --H--> A [h][s][3] --V--> B [s][s][3] --H--> A [s][w][3] --V--> dst
__global__ void __launch_bounds__(kPixThreads, 1) pixelate_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                                   int h, int w, int s, DevTable down, DevTable up) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* A = smem;                                  // max(h*s, s*w)*3
  uint8_t* B = smem + (size_t)h * s * 3;              // s*s*3
  const int img = blockIdx.x;
  const uint8_t* src = in + (size_t)img * h * w * 3;
  uint8_t* dst = out + (size_t)img * h * w * 3;
  // pass 1: horizontal down, every input row
  for (int i = threadIdx.x; i < h * s; i += kPixThreads) {
    const int y = i / s, xx = i - y * s;
    const int x0 = down.xmin[xx], cnt = down.xcnt[xx];
    const int* k = down.coef + xx * down.ksize;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int j = 0; j < cnt; ++j) {
      const uint8_t* p = src + ((size_t)y * w + x0 + j) * 3;
      const int c = k[j];
      a0 += p[0] * c; a1 += p[1] * c; a2 += p[2] * c;
    }
    uint8_t* o = A + (size_t)i * 3;
    o[0] = clip8(a0); o[1] = clip8(a1); o[2] = clip8(a2);
  }
  __syncthreads();
  // pass 2: vertical down
  for (int i = threadIdx.x; i < s * s; i += kPixThreads) {
    const int yy = i / s, x = i - yy * s;
    const int y0 = down.xmin[yy], cnt = down.xcnt[yy];
    const int* k = down.coef + yy * down.ksize;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int j = 0; j < cnt; ++j) {
      const uint8_t* p = A + ((size_t)(y0 + j) * s + x) * 3;
      const int c = k[j];
      a0 += p[0] * c; a1 += p[1] * c; a2 += p[2] * c;
    }
    uint8_t* o = B + (size_t)i * 3;
    o[0] = clip8(a0); o[1] = clip8(a1); o[2] = clip8(a2);
  }
  __syncthreads();
  // pass 3: horizontal up: B [s][s] -> A [s][w]
  for (int i = threadIdx.x; i < s * w; i += kPixThreads) {
    const int y = i / w, xx = i - y * w;
    const int x0 = up.xmin[xx], cnt = up.xcnt[xx];
    const int* k = up.coef + xx * up.ksize;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int j = 0; j < cnt; ++j) {
      const uint8_t* p = B + ((size_t)y * s + x0 + j) * 3;
      const int c = k[j];
      a0 += p[0] * c; a1 += p[1] * c; a2 += p[2] * c;
    }
    uint8_t* o = A + (size_t)i * 3;
    o[0] = clip8(a0); o[1] = clip8(a1); o[2] = clip8(a2);
  }
  __syncthreads();
  // pass 4: vertical up: A [s][w] -> dst [h][w]
  for (int i = threadIdx.x; i < h * w; i += kPixThreads) {
    const int yy = i / w, x = i - yy * w;
    const int y0 = up.xmin[yy], cnt = up.xcnt[yy];
    const int* k = up.coef + yy * up.ksize;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int j = 0; j < cnt; ++j) {
      const uint8_t* p = A + ((size_t)(y0 + j) * w + x) * 3;
      const int c = k[j];
      a0 += p[0] * c; a1 += p[1] * c; a2 += p[2] * c;
    }
    uint8_t* o = dst + (size_t)i * 3;
    o[0] = clip8(a0); o[1] = clip8(a1); o[2] = clip8(a2);
  }
}

// =============================================================================================
// jpeg_compression (corruptions.py:375-382): PIL save(JPEG, quality) + open == libjpeg(-turbo) baseline
// round trip.  Entropy coding is lossless, so the image only goes through
//   jccolor.c rgb_ycc_convert -> jcsample.c h2v2_downsample -> jfdctint.c jpeg_fdct_islow (x-128)
//   -> jcdctmgr.c quantize (divisor 8q, round half away) -> jidctint.c jpeg_idct_islow (x q)
//   -> jdsample.c h2v2_fancy_upsample -> jdcolor.c ycc_rgb_convert,
// all in integer arithmetic: bit-exact.  oracle/jpeg_restatement.py is the same algorithm in numpy and is
// checked against PIL itself.  One CTA per image; Y / Cb / Cr planes live in shared memory; one thread
// per 8x8 block does FDCT + quant + dequant + IDCT in registers.
// =============================================================================================
struct JpegQ { uint16_t luma[64], chroma[64]; };
constexpr int kJpegThreads = 256;
constexpr int CB = 13, P1 = 2;  // CONST_BITS, PASS1_BITS
#define JFIX(x) ((int)((x) * (1 << CB) + 0.5))

JpegQ make_jpeg_tables(int quality) {
  static const int std_l[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
                                14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                                49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
  static const int std_c[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
                                47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
  quality = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
  const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;   // jpeg_quality_scaling
  JpegQ q;
  for (int i = 0; i < 64; ++i) {
    int l = (std_l[i] * scale + 50) / 100, c = (std_c[i] * scale + 50) / 100;   // force_baseline: clamp to 1..255
    q.luma[i] = (uint16_t)(l < 1 ? 1 : (l > 255 ? 255 : l));
    q.chroma[i] = (uint16_t)(c < 1 ? 1 : (c > 255 ? 255 : c));
  }
  return q;
}

__device__ __forceinline__ int dsc(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// one 1-D pass of jpeg_fdct_islow on 8 values at stride S
template <int S, bool FIRST>
__device__ __forceinline__ void fdct8(int* d) {
  int t0 = d[0] + d[7 * S], t7 = d[0] - d[7 * S], t1 = d[S] + d[6 * S], t6 = d[S] - d[6 * S];
  int t2 = d[2 * S] + d[5 * S], t5 = d[2 * S] - d[5 * S], t3 = d[3 * S] + d[4 * S], t4 = d[3 * S] - d[4 * S];
  const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
  constexpr int n = FIRST ? CB - P1 : CB + P1;
  d[0] = FIRST ? ((t10 + t11) << P1) : dsc(t10 + t11, P1);
  d[4 * S] = FIRST ? ((t10 - t11) << P1) : dsc(t10 - t11, P1);
  int z1 = (t12 + t13) * JFIX(0.541196100);
  d[2 * S] = dsc(z1 + t13 * JFIX(0.765366865), n);
  d[6 * S] = dsc(z1 - t12 * JFIX(1.847759065), n);
  z1 = t4 + t7;
  int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
  const int z5 = (z3 + z4) * JFIX(1.175875602);
  t4 *= JFIX(0.298631336); t5 *= JFIX(2.053119869); t6 *= JFIX(3.072711026); t7 *= JFIX(1.501321110);
  z1 *= -JFIX(0.899976223); z2 *= -JFIX(2.562915447);
  z3 = z3 * -JFIX(1.961570560) + z5; z4 = z4 * -JFIX(0.390180644) + z5;
  d[7 * S] = dsc(t4 + z1 + z3, n); d[5 * S] = dsc(t5 + z2 + z4, n);
  d[3 * S] = dsc(t6 + z2 + z3, n); d[S] = dsc(t7 + z1 + z4, n);
}

// one 1-D pass of jpeg_idct_islow on 8 values at stride S
template <int S, bool FIRST>
__device__ __forceinline__ void idct8(int* w) {
  int z2 = w[2 * S], z3 = w[6 * S];
  int z1 = (z2 + z3) * JFIX(0.541196100);
  int t2 = z1 - z3 * JFIX(1.847759065), t3 = z1 + z2 * JFIX(0.765366865);
  z2 = w[0]; z3 = w[4 * S];
  int t0 = (z2 + z3) << CB, t1 = (z2 - z3) << CB;
  const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
  t0 = w[7 * S]; t1 = w[5 * S]; t2 = w[3 * S]; t3 = w[S];
  z1 = t0 + t3; z2 = t1 + t2; z3 = t0 + t2;
  int z4 = t1 + t3;
  const int z5 = (z3 + z4) * JFIX(1.175875602);
  t0 *= JFIX(0.298631336); t1 *= JFIX(2.053119869); t2 *= JFIX(3.072711026); t3 *= JFIX(1.501321110);
  z1 *= -JFIX(0.899976223); z2 *= -JFIX(2.562915447);
  z3 = z3 * -JFIX(1.961570560) + z5; z4 = z4 * -JFIX(0.390180644) + z5;
  t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
  constexpr int n = FIRST ? CB - P1 : CB + P1 + 3;
  w[0] = dsc(t10 + t3, n); w[7 * S] = dsc(t10 - t3, n);
  w[S] = dsc(t11 + t2, n); w[6 * S] = dsc(t11 - t2, n);
  w[2 * S] = dsc(t12 + t1, n); w[5 * S] = dsc(t12 - t1, n);
  w[3 * S] = dsc(t13 + t0, n); w[4 * S] = dsc(t13 - t0, n);
}

__device__ __forceinline__ void jpeg_block_roundtrip(uint8_t* plane, int pitch, const uint16_t* q) {
  int b[64];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const uint2 v = *reinterpret_cast<const uint2*>(plane + r * pitch);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      b[8 * r + k] = (int)((v.x >> (8 * k)) & 0xFF) - 128;
      b[8 * r + 4 + k] = (int)((v.y >> (8 * k)) & 0xFF) - 128;
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) fdct8<1, true>(b + 8 * r);      // rows
#pragma unroll
  for (int c = 0; c < 8; ++c) fdct8<8, false>(b + c);         // columns
#pragma unroll
  for (int i = 0; i < 64; ++i) {                              // quantise (divisor 8q, half away) + dequantise
    const int qq = q[i], q8 = qq << 3;
    const int a = abs(b[i]);
    const int t = (a + (q8 >> 1)) / q8;
    b[i] = (b[i] < 0 ? -t : t) * qq;
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) idct8<8, true>(b + c);          // pass 1: columns
#pragma unroll
  for (int r = 0; r < 8; ++r) idct8<1, false>(b + 8 * r);     // pass 2: rows
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lo |= (uint32_t)min(max(b[8 * r + k] + 128, 0), 255) << (8 * k);
      hi |= (uint32_t)min(max(b[8 * r + 4 + k] + 128, 0), 255) << (8 * k);
    }
    *reinterpret_cast<uint2*>(plane + r * pitch) = make_uint2(lo, hi);
  }
}

__device__ __forceinline__ int fancy_up(const uint8_t* c, int cw, int chh, int y, int x) {
  // h2v2_fancy_upsample: output (y, x) of the 2x plane from the chroma plane c [chh][cw]
  const int i = y >> 1, j = x >> 1;
  const int io = (y & 1) ? min(i + 1, chh - 1) : max(i - 1, 0);
  auto colsum = [&](int jj) { return (int)c[i * cw + jj] * 3 + (int)c[io * cw + jj]; };
  const int cs = colsum(j);
  if ((x & 1) == 0) return (j == 0) ? (cs * 4 + 8) >> 4 : (cs * 3 + colsum(j - 1) + 8) >> 4;
  return (j == cw - 1) ? (cs * 4 + 7) >> 4 : (cs * 3 + colsum(j + 1) + 7) >> 4;
}

__global__ void __launch_bounds__(kJpegThreads) jpeg_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int h, int w,
                                                             JpegQ q) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int cw = w / 2, chh = h / 2;
  uint8_t* sY = smem;
  uint8_t* sCb = sY + (size_t)h * w;
  uint8_t* sCr = sCb + (size_t)chh * cw;
  __shared__ uint16_t s_q[128];
  if (threadIdx.x < 64) { s_q[threadIdx.x] = q.luma[threadIdx.x]; s_q[64 + threadIdx.x] = q.chroma[threadIdx.x]; }
  const int img = blockIdx.x;
  const uint8_t* src = in + (size_t)img * h * w * 3;
  // 1) colour conversion (SCALEBITS 16) + 2x2 chroma downsample with the alternating 1,2 bias
  constexpr int FR = 19595, FG = 38470, FB = 7471;            // FIX(0.299), FIX(0.587), FIX(0.114)
  constexpr int CBR = 11059, CBG = 21709, CBB = 32768;        // FIX(0.16874), FIX(0.33126), FIX(0.5)
  constexpr int CRG = 27439, CRB = 5329;                      // FIX(0.41869), FIX(0.08131)
  for (int qd = threadIdx.x; qd < chh * cw; qd += kJpegThreads) {
    const int cy = qd / cw, cx = qd - cy * cw;
    int sb = 0, sr = 0;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int y = 2 * cy + dy, x = 2 * cx + dx;
        const uint8_t* p = src + ((size_t)y * w + x) * 3;
        const int r = p[0], g = p[1], b = p[2];
        sY[y * w + x] = (uint8_t)((FR * r + FG * g + FB * b + 32768) >> 16);
        sb += (-CBR * r - CBG * g + CBB * b + (128 << 16) + 32767) >> 16;
        sr += (CBB * r - CRG * g - CRB * b + (128 << 16) + 32767) >> 16;
      }
    const int bias = 1 + (cx & 1);
    sCb[qd] = (uint8_t)((sb + bias) >> 2);
    sCr[qd] = (uint8_t)((sr + bias) >> 2);
  }
  __syncthreads();
  // 2) per 8x8 block: FDCT -> quantise -> dequantise -> IDCT, in place
  const int yb = (h / 8) * (w / 8), cb = (chh / 8) * (cw / 8);
  for (int blk = threadIdx.x; blk < yb + 2 * cb; blk += kJpegThreads) {
    if (blk < yb) {
      const int by = blk / (w / 8), bx = blk - by * (w / 8);
      jpeg_block_roundtrip(sY + (by * 8) * w + bx * 8, w, s_q);
    } else {
      const int k = blk - yb, pl = k / cb, kk = k - pl * cb;
      const int by = kk / (cw / 8), bx = kk - by * (cw / 8);
      jpeg_block_roundtrip((pl ? sCr : sCb) + (by * 8) * cw + bx * 8, cw, s_q + 64);
    }
  }
  __syncthreads();
  // 3) fancy upsampling + YCbCr -> RGB, 4 pixels (12 bytes) per thread
  uint32_t* dst = reinterpret_cast<uint32_t*>(out + (size_t)img * h * w * 3);
  for (int g4 = threadIdx.x; g4 < h * w / 4; g4 += kJpegThreads) {
    const int pix0 = g4 * 4, y = pix0 / w, x0 = pix0 - y * w;
    uint8_t o[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = x0 + k;
      const int yy = sY[y * w + x];
      const int cbv = fancy_up(sCb, cw, chh, y, x) - 128, crv = fancy_up(sCr, cw, chh, y, x) - 128;
      const int r = yy + ((91881 * crv + 32768) >> 16);                       // FIX(1.40200)
      const int g = yy + ((-22554 * cbv + 32768 - 46802 * crv) >> 16);        // FIX(0.34414), FIX(0.71414)
      const int b = yy + ((116130 * cbv + 32768) >> 16);                      // FIX(1.77200)
      o[3 * k] = (uint8_t)min(max(r, 0), 255); o[3 * k + 1] = (uint8_t)min(max(g, 0), 255); o[3 * k + 2] = (uint8_t)min(max(b, 0), 255);
    }
    dst[g4 * 3] = o[0] | (o[1] << 8) | (o[2] << 16) | ((uint32_t)o[3] << 24);
    dst[g4 * 3 + 1] = o[4] | (o[5] << 8) | (o[6] << 16) | ((uint32_t)o[7] << 24);
    dst[g4 * 3 + 2] = o[8] | (o[9] << 8) | (o[10] << 16) | ((uint32_t)o[11] << 24);
  }
}

}  // namespace

size_t corrupt_codec_ws(int, int, int, int, int) { return 0; }

int corrupt_codec_family(const CorruptArgs& a) {
  if (a.id == B200R_PIXELATE) {
    static const double c[5] = {0.6, 0.5, 0.4, 0.3, 0.25};
    B200R_CHECK_ARG(a.h == a.w, "pixelate expects square images (the reference hard-codes 224)");
    const int s = (int)(a.h * c[a.severity - 1]);   // int(224 * c)
    int dev = 0;
    B200R_CUDA(cudaGetDevice(&dev));
    B200R_CHECK_ARG(dev >= 0 && dev < 8, "device index out of range");
    PixelateTables t;
    {
      std::lock_guard<std::mutex> lk(g_pix_mu);
      PixelateTables& g = g_pix[dev][a.severity - 1];
      if (!g.ready || g.small != s) {  // first use: host tables + blocking copies (not capturable)
        int rc = upload(make_box_table(a.h, s), g.down);
        if (rc) return rc;
        rc = upload(make_box_table(s, a.h), g.up);
        if (rc) return rc;
        g.ready = true; g.small = s;
      }
      t = g;
    }
    B200R_CHECK_ARG(t.down.ksize <= kMaxK && t.up.ksize <= kMaxK, "unexpected filter size");
    const size_t smem = (size_t)a.h * s * 3 + (size_t)s * s * 3;
    B200R_CHECK_ARG(smem <= 220 * 1024, "image too large for shared memory");
    B200R_CUDA(cudaFuncSetAttribute(pixelate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pixelate_kernel<<<a.n, kPixThreads, smem, a.stream>>>(a.in, a.out, a.h, a.w, s, t.down, t.up);
    B200R_LAUNCH_CHECK();
    return B200R_OK;
  }
  // ---- jpeg_compression ----
  static const int quality[5] = {25, 18, 15, 10, 7};
  B200R_CHECK_ARG(a.h % 16 == 0 && a.w % 16 == 0, "jpeg_compression needs h, w multiples of 16 (4:2:0 MCUs)");
  const size_t smem = (size_t)a.h * a.w + 2 * (size_t)(a.h / 2) * (a.w / 2);
  B200R_CHECK_ARG(smem <= 220 * 1024, "image too large for shared memory");
  JpegQ q = make_jpeg_tables(quality[a.severity - 1]);
  B200R_CUDA(cudaFuncSetAttribute(jpeg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  jpeg_kernel<<<a.n, kJpegThreads, smem, a.stream>>>(a.in, a.out, a.h, a.w, q);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}
