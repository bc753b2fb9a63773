// cv2.resize for uint8 NHWC batches, INTER_NEAREST, INTER_LINEAR, INTER_AREA, INTER_LANCZOS4 (bit for bit) and INTER_CUBIC (OpenCV imgproc/resize.cpp restated in
// oracle/cv_resize.py and pinned there against cv2 4.13): the `opencv-nearest` / `opencv-bilinear` / `opencv-area` types of the reference's
// ImageNet-S generator (RobustART/noise/utils/imagenet_s_gen.py:28-34,120-148), with the centre crop of the 'val' transform fused
// (only the cropped window is produced).
//   linear: coordinates f = float((d + 0.5) * scale - 0.5), scale = 1 / (dst / src) in double; weights cvRound(w * 2048);
//           horizontal pass  S = p[sx] * a0 + p[sx + 1] * a1   (weight zeroed where the pair leaves the row)
//           vertical pass    (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2   (row indices clipped, weights kept)
//   nearest: src = min(floor(d * scale), size - 1)
//   area:    integer factors on both axes -> block sums, (s + 2) >> 2 for 2x2, else cvRound(float(s) * float(1 / (kx ky)));
//            both axes shrinking -> computeResizeAreaTab weights (float32), rows accumulated in OpenCV's order with one rounding per
//            multiply and per add; an axis growing -> the linear kernel on the "area mode" coefficients
//   lanczos4: OpenCV's 8-tap fixed-point path: weights from interpolateLanczos4 (double sin/cos on the HOST, as OpenCV does; the
//            table is cached per geometry), horizontal sums in int, (sum + 2^21) >> 22
//   cubic:   the opencv-python wheels send INTER_CUBIC to IPP = a float32 Keys cubic (A = -0.75): 4 x 4 float taps, horizontal pass
//            first, one rounding per multiply and add, round-half-even (parity bar: 1 LSB on <= 1e-4 of the pixels, see the oracle)
// One thread per output pixel (3 channels), weights recomputed per thread (a dozen flops against four 3-byte gathers): the kernel
// is bound by the gathers, which hit L1/L2 for every down-scaling ratio the eval transform sees.
// STATUS: checked from source on the host emulator against cv2.resize (tests/test_kernel_emulation_cpu.py); not yet run on a GPU.
#include "common.cuh"
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace {
constexpr int kThreads = 256;

// ---- table-driven K-tap modes (lanczos4: K = 8 fixed point, cubic: K = 4 float) ------------------------------------------------
struct TapTable { int* idx; void* w; };                      // device: anchor per output coordinate, K weights per coordinate
struct AxisTables { TapTable x, y; };

template <int K, bool FIXED>
__global__ void __launch_bounds__(kThreads) resize_taps_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n, int hin, int win,
                                                                const int* __restrict__ xi, const void* __restrict__ xw_,
                                                                const int* __restrict__ yi, const void* __restrict__ yw_, int oy0, int ox0,
                                                                int ch, int cw) {
  const size_t total = (size_t)n * ch * cw;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int x = (int)(t % cw) + ox0, y = (int)((t / cw) % ch) + oy0, im = (int)(t / ((size_t)cw * ch));
    const uint8_t* src = in + (size_t)im * hin * win * 3;
    uint8_t* dst = out + t * 3;
    const int sx = xi[x] - (K / 2 - 1), sy = yi[y] - (K / 2 - 1);
    int cols[K];
#pragma unroll
    for (int k = 0; k < K; ++k) cols[k] = min(max(sx + k, 0), win - 1) * 3;
    if (FIXED) {
      const int* xw = static_cast<const int*>(xw_) + (size_t)x * K;
      const int* yw = static_cast<const int*>(yw_) + (size_t)y * K;
      unsigned acc[3] = {0u, 0u, 0u};                         // OpenCV accumulates in int: wrap-around arithmetic, made explicit
      for (int j = 0; j < K; ++j) {
        const uint8_t* row = src + (size_t)min(max(sy + j, 0), hin - 1) * win * 3;
        int r[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < K; ++k) {
          r[0] += row[cols[k]] * xw[k]; r[1] += row[cols[k] + 1] * xw[k]; r[2] += row[cols[k] + 2] * xw[k];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] += (unsigned)r[c] * (unsigned)yw[j];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int v = (int)(acc[c] + (1u << 21)) >> 22;
        dst[c] = (uint8_t)min(max(v, 0), 255);
      }
    } else {
      const float* xw = static_cast<const float*>(xw_) + (size_t)x * K;
      const float* yw = static_cast<const float*>(yw_) + (size_t)y * K;
      float acc[3] = {0.f, 0.f, 0.f};
      for (int j = 0; j < K; ++j) {
        const uint8_t* row = src + (size_t)min(max(sy + j, 0), hin - 1) * win * 3;
        float r[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
          for (int c = 0; c < 3; ++c) r[c] = __fadd_rn(r[c], __fmul_rn((float)row[cols[k] + c], xw[k]));
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(r[c], yw[j]));
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) dst[c] = (uint8_t)min(max((int)rintf(acc[c]), 0), 255);
    }
  }
}

// host side: the weight tables, computed as OpenCV computes them (libm double sin/cos for Lanczos), cached per geometry
void lanczos4_weights(float x, int* q) {
  static const double s45 = 0.70710678118654752440084436210485;
  static const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
  const double pi = 3.1415926535897932384626433832795;
  float co[8], sum = 0.f;
  const double y0 = -((double)x + 3) * pi * 0.25, s0 = sin(y0), c0 = cos(y0);
  for (int i = 0; i < 8; ++i) {
    const float d = (x + 3.f) - (float)i;
    if (fabsf(d) >= 1e-6f) {
      const double y = -(double)d * pi * 0.25;
      co[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
    } else co[i] = 1e30f;
    sum += co[i];
  }
  sum = 1.f / sum;
  for (int i = 0; i < 8; ++i) q[i] = (int)lrintf((co[i] * sum) * 2048.f);
}
void build_axis(int nin, int nout, bool lanczos, std::vector<int>& idx, std::vector<int>& wi, std::vector<float>& wf) {
  const double scale = 1.0 / ((double)nout / (double)nin);
  idx.resize(nout);
  if (lanczos) wi.resize((size_t)nout * 8); else wf.resize((size_t)nout * 4);
  for (int d = 0; d < nout; ++d) {
    if (lanczos) {
      float f = (float)(((double)d + 0.5) * scale - 0.5);
      const int s = (int)floorf(f);
      f -= (float)s;
      idx[d] = s;
      lanczos4_weights(f, &wi[(size_t)d * 8]);
    } else {
      const double f = ((double)d + 0.5) * scale - 0.5, fl = floor(f), x = f - fl, A = -0.75;
      const double c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
      const double c1 = ((A + 2) * x - (A + 3)) * x * x + 1;
      const double c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
      idx[d] = (int)fl;
      float* w = &wf[(size_t)d * 4];
      w[0] = (float)c0; w[1] = (float)c1; w[2] = (float)c2; w[3] = (float)(1 - c0 - c1 - c2);
    }
  }
}
std::mutex g_tab_mu;
std::map<std::tuple<int, int, int, int, int, int>, AxisTables> g_tabs;
int get_tables(int hin, int win, int hout, int wout, bool lanczos, AxisTables* out) {
  int dev = 0;
  B200R_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_tab_mu);
  const auto key = std::make_tuple(dev, hin, win, hout, wout, lanczos ? 1 : 0);
  auto it = g_tabs.find(key);
  if (it == g_tabs.end()) {
    AxisTables t;
    TapTable* tt[2] = {&t.x, &t.y};
    const int nin[2] = {win, hin}, nout[2] = {wout, hout};
    for (int a = 0; a < 2; ++a) {
      std::vector<int> idx, wi;
      std::vector<float> wf;
      build_axis(nin[a], nout[a], lanczos, idx, wi, wf);
      const size_t wbytes = lanczos ? wi.size() * sizeof(int) : wf.size() * sizeof(float);
      B200R_CUDA(cudaMalloc(reinterpret_cast<void**>(&tt[a]->idx), idx.size() * sizeof(int)));
      B200R_CUDA(cudaMalloc(&tt[a]->w, wbytes));
      B200R_CUDA(cudaMemcpy(tt[a]->idx, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice));
      B200R_CUDA(cudaMemcpy(tt[a]->w, lanczos ? static_cast<const void*>(wi.data()) : static_cast<const void*>(wf.data()), wbytes,
                            cudaMemcpyHostToDevice));
    }
    it = g_tabs.emplace(key, t).first;
  }
  *out = it->second;
  return B200R_OK;
}

struct Tap { int s; int w0, w1; };

__device__ __forceinline__ Tap linear_tap(int d, double scale, int nin, bool clamp) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (clamp) {
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= nin - 1) { s = nin - 1; f = 0.f; }
  }
  Tap t;
  t.s = s;
  t.w0 = (int)rintf(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  t.w1 = (int)rintf(__fmul_rn(f, 2048.f));
  return t;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// INTER_AREA with a growing axis: resize.cpp's area_mode coefficients
__device__ __forceinline__ Tap area_linear_tap(int d, double scale, double inv, int nin, bool clamp) {
  int s = (int)floor((double)d * scale);
  float f = (float)((double)(d + 1) - (double)(s + 1) * inv);
  f = f <= 0.f ? 0.f : __fsub_rn(f, floorf(f));
  if (clamp) {
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= nin - 1) { s = nin - 1; f = 0.f; }
  }
  Tap t;
  t.s = s;
  t.w0 = (int)rintf(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  t.w1 = (int)rintf(__fmul_rn(f, 2048.f));
  return t;
}

// computeResizeAreaTab for one destination index: an optional partial first cell, whole cells [s1, s2), an optional partial last cell
struct AreaTaps { int s1, s2; float wfirst, wmid, wlast; bool first, last; };
__device__ __forceinline__ AreaTaps area_taps(int d, double scale, int nin) {
  const double f1 = (double)d * scale, f2 = f1 + scale;
  const double cell = fmin(scale, (double)nin - f1);
  int s1 = (int)ceil(f1), s2 = (int)floor(f2);
  s2 = min(s2, nin - 1);
  s1 = min(s1, s2);
  AreaTaps t;
  t.s1 = s1; t.s2 = s2;
  t.first = (double)s1 - f1 > 1e-3;
  t.wfirst = (float)(((double)s1 - f1) / cell);
  t.wmid = (float)(1.0 / cell);
  t.last = f2 - (double)s2 > 1e-3;
  t.wlast = (float)(fmin(fmin(f2 - (double)s2, 1.0), cell) / cell);
  return t;
}

// both axes shrinking (MODE_TAB) or integer factors (MODE_FAST)
template <bool FAST>
__global__ void __launch_bounds__(kThreads) resize_area_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n, int hin, int win,
                                                                double sy, double sx, int oy0, int ox0, int ch, int cw) {
  const size_t total = (size_t)n * ch * cw;
  const int kx = (int)sx, ky = (int)sy;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int x = (int)(t % cw) + ox0, y = (int)((t / cw) % ch) + oy0, im = (int)(t / ((size_t)cw * ch));
    const uint8_t* src = in + (size_t)im * hin * win * 3;
    uint8_t* dst = out + t * 3;
    if (FAST) {
      int s[3] = {0, 0, 0};
      for (int j = 0; j < ky; ++j) {
        const uint8_t* p = src + ((size_t)(y * ky + j) * win + (size_t)x * kx) * 3;
        for (int i = 0; i < kx; ++i) { s[0] += p[3 * i]; s[1] += p[3 * i + 1]; s[2] += p[3 * i + 2]; }
      }
      if (kx == 2 && ky == 2) {
#pragma unroll
        for (int c = 0; c < 3; ++c) dst[c] = (uint8_t)((s[c] + 2) >> 2);
      } else {
        const float scale = (float)(1.0 / (double)(kx * ky));
#pragma unroll
        for (int c = 0; c < 3; ++c) dst[c] = (uint8_t)clampi((int)rintf(__fmul_rn((float)s[c], scale)), 0, 255);
      }
      continue;
    }
    const AreaTaps tx = area_taps(x, sx, win), ty = area_taps(y, sy, hin);
    float sum[3] = {0.f, 0.f, 0.f};
    bool started = false;
    const int ny = (ty.first ? 1 : 0) + (ty.s2 - ty.s1) + (ty.last ? 1 : 0);
    for (int k = 0; k < ny; ++k) {
      int yy; float beta;
      if (ty.first && k == 0) { yy = ty.s1 - 1; beta = ty.wfirst; }
      else {
        const int kk = k - (ty.first ? 1 : 0);
        if (kk < ty.s2 - ty.s1) { yy = ty.s1 + kk; beta = ty.wmid; }
        else { yy = ty.s2; beta = ty.wlast; }
      }
      const uint8_t* row = src + (size_t)yy * win * 3;
      float b[3] = {0.f, 0.f, 0.f};
      if (tx.first) {
        const uint8_t* p = row + (size_t)(tx.s1 - 1) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) b[c] = __fadd_rn(b[c], __fmul_rn((float)p[c], tx.wfirst));
      }
      for (int xx = tx.s1; xx < tx.s2; ++xx) {
        const uint8_t* p = row + (size_t)xx * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) b[c] = __fadd_rn(b[c], __fmul_rn((float)p[c], tx.wmid));
      }
      if (tx.last) {
        const uint8_t* p = row + (size_t)tx.s2 * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) b[c] = __fadd_rn(b[c], __fmul_rn((float)p[c], tx.wlast));
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) sum[c] = started ? __fadd_rn(sum[c], __fmul_rn(beta, b[c])) : __fmul_rn(beta, b[c]);
      started = true;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[c] = (uint8_t)clampi((int)rintf(sum[c]), 0, 255);
  }
}

// MODE 0 nearest, 1 linear, 2 linear on the area-mode coefficients
template <int MODE>
__global__ void __launch_bounds__(kThreads) resize_cv_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n, int hin, int win,
                                                              double sy, double sx, double iy, double ix, int oy0, int ox0, int ch, int cw) {
  const size_t total = (size_t)n * ch * cw;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int x = (int)(t % cw), y = (int)((t / cw) % ch), im = (int)(t / ((size_t)cw * ch));
    const uint8_t* src = in + (size_t)im * hin * win * 3;
    uint8_t* dst = out + t * 3;
    if (MODE == 0) {
      const int xs = min((int)floor((double)(x + ox0) * sx), win - 1), ys = min((int)floor((double)(y + oy0) * sy), hin - 1);
      const uint8_t* p = src + ((size_t)ys * win + xs) * 3;
      dst[0] = p[0]; dst[1] = p[1]; dst[2] = p[2];
      continue;
    }
    const Tap tx = MODE == 2 ? area_linear_tap(x + ox0, sx, ix, win, true) : linear_tap(x + ox0, sx, win, true);
    const Tap ty = MODE == 2 ? area_linear_tap(y + oy0, sy, iy, hin, false) : linear_tap(y + oy0, sy, hin, false);
    const int x1 = min(tx.s + 1, win - 1);
    const uint8_t* r0 = src + (size_t)clampi(ty.s, 0, hin - 1) * win * 3;
    const uint8_t* r1 = src + (size_t)clampi(ty.s + 1, 0, hin - 1) * win * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int s0 = r0[tx.s * 3 + c] * tx.w0 + r0[x1 * 3 + c] * tx.w1;
      const int s1 = r1[tx.s * 3 + c] * tx.w0 + r1[x1 * 3 + c] * tx.w1;
      const int v = (((ty.w0 * (s0 >> 4)) >> 16) + ((ty.w1 * (s1 >> 4)) >> 16) + 2) >> 2;
      dst[c] = (uint8_t)clampi(v, 0, 255);
    }
  }
}
}  // namespace

extern "C" int b200r_resize_cv_u8(const uint8_t* in, uint8_t* out, int n, int hin, int win, int hout, int wout, int interpolation, int oy0,
                                  int ox0, int ch, int cw, b200r_stream_t stream) {
  B200R_CHECK_ARG(in && out, "null pointer");
  B200R_CHECK_ARG(n > 0 && hin > 0 && win > 0 && hout > 0 && wout > 0, "bad shape");
  B200R_CHECK_ARG(interpolation == B200R_CV_INTER_NEAREST || interpolation == B200R_CV_INTER_LINEAR || interpolation == B200R_CV_INTER_AREA ||
                      interpolation == B200R_CV_INTER_CUBIC || interpolation == B200R_CV_INTER_LANCZOS4,
                  "interpolation %d not supported (cv2.INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3, INTER_LANCZOS4 = 4)",
                  interpolation);
  B200R_CHECK_ARG(oy0 >= 0 && ox0 >= 0 && ch > 0 && cw > 0 && oy0 + ch <= hout && ox0 + cw <= wout, "crop window outside the resized image");
  const double iy = (double)hout / (double)hin, ix = (double)wout / (double)win;      // resize.cpp: inv_scale, then scale = 1 / inv_scale
  const double sy = 1.0 / iy, sx = 1.0 / ix;
  const size_t total = (size_t)n * ch * cw;
  size_t blocks = (total + kThreads - 1) / kThreads;
  const size_t cap = (size_t)b200r_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  const unsigned g = (unsigned)blocks;
  cudaStream_t st = as_stream(stream);
  if (interpolation == B200R_CV_INTER_CUBIC || interpolation == B200R_CV_INTER_LANCZOS4) {
    const bool lz = interpolation == B200R_CV_INTER_LANCZOS4;
    AxisTables tb;
    const int rc = get_tables(hin, win, hout, wout, lz, &tb);       // first call per geometry: blocking upload of the weight tables
    if (rc != B200R_OK) return rc;
    if (lz) resize_taps_kernel<8, true><<<g, kThreads, 0, st>>>(in, out, n, hin, win, tb.x.idx, tb.x.w, tb.y.idx, tb.y.w, oy0, ox0, ch, cw);
    else resize_taps_kernel<4, false><<<g, kThreads, 0, st>>>(in, out, n, hin, win, tb.x.idx, tb.x.w, tb.y.idx, tb.y.w, oy0, ox0, ch, cw);
  } else if (interpolation == B200R_CV_INTER_LINEAR) {
    resize_cv_kernel<1><<<g, kThreads, 0, st>>>(in, out, n, hin, win, sy, sx, iy, ix, oy0, ox0, ch, cw);
  } else if (interpolation == B200R_CV_INTER_NEAREST) {
    resize_cv_kernel<0><<<g, kThreads, 0, st>>>(in, out, n, hin, win, sy, sx, iy, ix, oy0, ox0, ch, cw);
  } else if (!(sx >= 1.0 && sy >= 1.0)) {
    resize_cv_kernel<2><<<g, kThreads, 0, st>>>(in, out, n, hin, win, sy, sx, iy, ix, oy0, ox0, ch, cw);
  } else {
    const int kx = (int)sx, ky = (int)sy;
    const bool fast = fabs(sx - kx) < 2.220446049250313e-16 && fabs(sy - ky) < 2.220446049250313e-16;     // DBL_EPSILON, as resize.cpp
    if (fast) resize_area_kernel<true><<<g, kThreads, 0, st>>>(in, out, n, hin, win, sy, sx, oy0, ox0, ch, cw);
    else resize_area_kernel<false><<<g, kThreads, 0, st>>>(in, out, n, hin, win, sy, sx, oy0, ox0, ch, cw);
  }
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}
