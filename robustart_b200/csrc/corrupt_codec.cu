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
  b200r_set_error("jpeg_compression is not implemented on the GPU yet");
  return B200R_ENOTSUP;
}
