// spatter, "water" branch (severity 1-3; RobustART/noise/utils/imagenet_c/corruptions.py:305-328), everything after the
// thresholded liquid layer, restated from the OpenCV calls the reference makes (byte-exact against cv2 4.13 on the CPU:
// oracle/spatter_water.py, tests/test_oracle_cpu.py):
//   l8   = uint8(liquid * 255)
//   edge = cv2.Canny(l8, 50, 150)          3x3 Sobel (replicate border), L1 magnitude, fixed-point tan(22.5) sectors, hysteresis
//   dist = min(cv2.distanceTransform(255 - edge, DIST_L2, 5), 20)
//                                          5x5 chamfer metric (1, 1.4, 2.1969) in float32: closed form over the edge pixels of the
//                                          41x41 window (a distance <= 20 cannot come from further away), summed axial moves first
//   u    = uint8(cv2.blur(dist, (3,3)))    double sums * (1/9), reflect-101 border
//   u    = cv2.equalizeHist(u)
//   u    = cv2.filter2D(u, CV_8U, [[-2,-1,0],[-1,1,1],[0,1,2]])
//   b    = float32(cv2.blur(u, (3,3)))     rounded integer mean
//   m    = l8 * b;  m = m / max(m) * c4;   out = uint8(clip(x/255 + m * (175,238,238)/255, 0, 1) * 255)
// One CTA per image walks the stages with __syncthreads between them; the planes live in the caller's workspace (L2-resident:
// 16 bytes per pixel).  The edge map is packed into a shared-memory bit mask so that the distance stage touches only the edge
// pixels of its window.
// STATUS: written after this round's GPU budget was spent -- checked on the host emulator (tests/test_kernel_emulation_cpu.py)
// only; corrupt_stencil.cu calls it when B200R_SPATTER_WATER=1 and keeps returning B200R_ENOTSUP otherwise.
#include "common.cuh"

namespace {
constexpr int kWaterThreads = 1024;
constexpr int kR = 20;                 // truncation radius of the distance transform
constexpr int kMaxBitWords = 8192;     // 32 KB of shared memory for the edge bit mask

__device__ __forceinline__ int reflect101(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void __launch_bounds__(kWaterThreads) spatter_water_kernel(const float* __restrict__ liquid, float* __restrict__ dist,
                                                                       int* __restrict__ grad, uint8_t* __restrict__ planes,
                                                                       const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int h, int w,
                                                                       float c4) {
  __shared__ uint32_t bits[kMaxBitWords];
  __shared__ float cham[(kR + 1) * (kR + 1)];
  __shared__ int hist[256];
  __shared__ uint8_t lut[256];
  __shared__ float red[kWaterThreads / 32];
  __shared__ int flag;
  const int img = blockIdx.x, hw = h * w, tid = threadIdx.x;
  liquid += (size_t)img * hw;
  dist += (size_t)img * hw;
  grad += (size_t)img * hw;
  uint8_t* l8 = planes + (size_t)img * hw * 4;
  uint8_t* st = l8 + hw;
  uint8_t* pa = st + hw;
  uint8_t* pb = pa + hw;
  in += (size_t)img * hw * 3;
  out += (size_t)img * hw * 3;

  // ---- stage 0: uint8(liquid * 255); chamfer table (moves summed in float32: axial, then diagonal, then knight) ----
  for (int i = tid; i < hw; i += kWaterThreads) l8[i] = (uint8_t)(int)(liquid[i] * 255.f);
  for (int i = tid; i < (kR + 1) * (kR + 1); i += kWaterThreads) {
    const int dy = i / (kR + 1), dx = i - dy * (kR + 1);
    const int mx = dx > dy ? dx : dy, mn = dx > dy ? dy : dx;
    int knight, axial, diag;
    if (mx >= 2 * mn) { knight = mn; axial = mx - 2 * mn; diag = 0; }
    else { knight = mx - mn; axial = 0; diag = 2 * mn - mx; }
    float s = 0.f;
    for (int k = 0; k < axial; ++k) s = __fadd_rn(s, 1.0f);
    for (int k = 0; k < diag; ++k) s = __fadd_rn(s, 1.4f);
    for (int k = 0; k < knight; ++k) s = __fadd_rn(s, 2.1969f);
    cham[i] = s;
  }
  __syncthreads();

  // ---- stage 1: Sobel (replicate border), packed (gx, gy) ----
  for (int i = tid; i < hw; i += kWaterThreads) {
    const int y = i / w, x = i - y * w;
    const int y0 = clampi(y - 1, 0, h - 1), y2 = clampi(y + 1, 0, h - 1), x0 = clampi(x - 1, 0, w - 1), x2 = clampi(x + 1, 0, w - 1);
    const int a = l8[y0 * w + x0], b = l8[y0 * w + x], c = l8[y0 * w + x2];
    const int d = l8[y * w + x0], f = l8[y * w + x2];
    const int g = l8[y2 * w + x0], hh = l8[y2 * w + x], k = l8[y2 * w + x2];
    const int gx = (c + 2 * f + k) - (a + 2 * d + g), gy = (g + 2 * hh + k) - (a + 2 * b + c);
    grad[i] = (int)(((uint32_t)gx & 0xFFFFu) | ((uint32_t)gy << 16));
  }
  __syncthreads();

  // ---- stage 2: non-maximum suppression -> 0 none, 1 weak (> 50), 2 strong (> 150) ----
  auto mag_at = [&](int y, int x) -> int {
    if (y < 0 || y >= h || x < 0 || x >= w) return 0;
    const int p = grad[y * w + x];
    const int gx = (int)(short)(p & 0xFFFF), gy = p >> 16;
    return (gx < 0 ? -gx : gx) + (gy < 0 ? -gy : gy);
  };
  for (int i = tid; i < hw; i += kWaterThreads) {
    const int y = i / w, x = i - y * w;
    const int p = grad[i];
    const int gx = (int)(short)(p & 0xFFFF), gy = p >> 16;
    const int ax = gx < 0 ? -gx : gx, ayy = gy < 0 ? -gy : gy, m = ax + ayy;
    uint8_t s = 0;
    if (m > 50) {
      const int ay = ayy << 15, tg22x = ax * 13573, tg67x = tg22x + (ax << 16);
      bool keep;
      if (ay < tg22x) keep = m > mag_at(y, x - 1) && m >= mag_at(y, x + 1);
      else if (ay > tg67x) keep = m > mag_at(y - 1, x) && m >= mag_at(y + 1, x);
      else {
        const int sgn = ((gx ^ gy) < 0) ? -1 : 1;
        keep = m > mag_at(y - 1, x - sgn) && m > mag_at(y + 1, x + sgn);
      }
      if (keep) s = m > 150 ? 2 : 1;
    }
    st[i] = s;
  }
  __syncthreads();

  // ---- stage 3: hysteresis (weak pixels 8-connected to a strong one become strong), to the fixed point ----
  for (;;) {
    if (tid == 0) flag = 0;
    __syncthreads();
    for (int i = tid; i < hw; i += kWaterThreads) {
      if (st[i] != 1) continue;
      const int y = i / w, x = i - y * w;
      bool hit = false;
      for (int dy = -1; dy <= 1 && !hit; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int yy = y + dy, xx = x + dx;
          if (yy >= 0 && yy < h && xx >= 0 && xx < w && st[yy * w + xx] == 2) { hit = true; break; }
        }
      if (hit) { st[i] = 2; flag = 1; }
    }
    __syncthreads();
    const int again = flag;
    __syncthreads();
    if (!again) break;
  }

  // ---- stage 4: edge bit mask, then the truncated chamfer distance ----
  const int wq = (w + 31) >> 5;
  for (int wd = tid; wd < h * wq; wd += kWaterThreads) {
    const int y = wd / wq, xw = wd - y * wq;
    uint32_t v = 0;
    for (int b = 0; b < 32; ++b) {
      const int x = xw * 32 + b;
      if (x < w && st[y * w + x] == 2) v |= 1u << b;
    }
    bits[wd] = v;
  }
  __syncthreads();
  for (int i = tid; i < hw; i += kWaterThreads) {
    const int y = i / w, x = i - y * w;
    float best = (float)kR;
    int xs = x - kR, sh = 0;
    if (xs < 0) { sh = -xs; xs = 0; }
    const int wi = xs >> 5, bo = xs & 31;
    for (int dy = -kR; dy <= kR; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= h) continue;
      const uint32_t* row = bits + yy * wq;
      const uint64_t lo = row[wi], mid = (wi + 1 < wq) ? row[wi + 1] : 0u, hi = (wi + 2 < wq) ? row[wi + 2] : 0u;
      uint64_t v = (lo | (mid << 32)) >> bo;
      if (bo) v |= hi << (64 - bo);
      v = (v << sh) & ((1ull << (2 * kR + 1)) - 1);                 // bit k <-> dx = k - kR
      const float* trow = cham + (dy < 0 ? -dy : dy) * (kR + 1);
      while (v) {
        const int k = __ffsll((long long)v) - 1;
        v &= v - 1;
        const int dx = k - kR;
        best = fminf(best, trow[dx < 0 ? -dx : dx]);
      }
    }
    dist[i] = best;
  }
  __syncthreads();

  // ---- stage 5: uint8(blur3(dist)) -> pa; histogram ----
  for (int i = tid; i < 256; i += kWaterThreads) hist[i] = 0;
  __syncthreads();
  for (int i = tid; i < hw; i += kWaterThreads) {
    const int y = i / w, x = i - y * w;
    double s = 0.0;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) s += (double)dist[reflect101(y + dy, h) * w + reflect101(x + dx, w)];
    const uint8_t u = (uint8_t)(int)(float)(s * (1.0 / 9.0));
    pa[i] = u;
    atomicAdd(&hist[u], 1);
  }
  __syncthreads();

  // ---- stage 6: equalizeHist -> pb ----
  if (tid == 0) {
    int i0 = 0;
    while (hist[i0] == 0) ++i0;
    if (hist[i0] == hw) {
      for (int k = 0; k < 256; ++k) lut[k] = (uint8_t)i0;
    } else {
      const float scale = __fdiv_rn(255.f, (float)(hw - hist[i0]));
      int sum = 0;
      for (int k = 0; k <= i0; ++k) lut[k] = 0;
      for (int k = i0 + 1; k < 256; ++k) {
        sum += hist[k];
        const int r = (int)rintf(__fmul_rn((float)sum, scale));
        lut[k] = (uint8_t)(r > 255 ? 255 : r);
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < hw; i += kWaterThreads) pb[i] = lut[pa[i]];
  __syncthreads();

  // ---- stage 7: filter2D with [[-2,-1,0],[-1,1,1],[0,1,2]] (correlation, reflect-101, saturate) -> pa ----
  for (int i = tid; i < hw; i += kWaterThreads) {
    const int y = i / w, x = i - y * w;
    const int ym = reflect101(y - 1, h) * w, yc = y * w, yp = reflect101(y + 1, h) * w;
    const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
    const int acc = -2 * pb[ym + xm] - pb[ym + x] - pb[yc + xm] + pb[yc + x] + pb[yc + xp] + pb[yp + x] + 2 * pb[yp + xp];
    pa[i] = (uint8_t)clampi(acc, 0, 255);
  }
  __syncthreads();

  // ---- stage 8: blur3 on uint8 (rounded mean) -> pb; m = l8 * pb and its maximum ----
  float mx = 0.f;
  for (int i = tid; i < hw; i += kWaterThreads) {
    const int y = i / w, x = i - y * w;
    int s = 0;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) s += pa[reflect101(y + dy, h) * w + reflect101(x + dx, w)];
    const int b = (2 * s + 9) / 18;
    pb[i] = (uint8_t)b;
    mx = fmaxf(mx, (float)((int)l8[i] * b));
  }
  mx = warp_max(mx);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int k = 1; k < kWaterThreads / 32; ++k) mx = fmaxf(mx, red[k]);

  // ---- stage 9: blend (float32, one rounding per numpy operation: no fused multiply-add) ----
  const float col[3] = {175.f / 255.f, 238.f / 255.f, 238.f / 255.f};
  for (int i = tid; i < hw; i += kWaterThreads) {
    float m = (float)((int)l8[i] * (int)pb[i]);
    m = __fmul_rn(__fdiv_rn(m, mx), c4);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint8_t o = 0;                                                  // max(m) == 0: the reference divides 0/0 and casts NaN -> 0
      if (mx > 0.f) {
        float v = __fadd_rn(__fdiv_rn((float)in[(size_t)i * 3 + c], 255.f), __fmul_rn(m, col[c]));
        v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
        o = (uint8_t)(int)__fmul_rn(v, 255.f);
      }
      out[(size_t)i * 3 + c] = o;
    }
  }
}
}  // namespace

// Not part of the C-ABI (include/b200r.h): called by corrupt_stencil.cu's spatter dispatch, exported for the host emulator test.
// liquid: float [n, h*w] thresholded liquid layer; dist: float scratch [n, h*w]; extra: 8 bytes per pixel (int plane + 4 byte planes).
extern "C" int b200r_spatter_water_planes(const float* liquid, float* dist, void* extra, const uint8_t* in, uint8_t* out, int n, int h,
                                          int w, float c4, b200r_stream_t stream) {
  B200R_CHECK_ARG(liquid && dist && extra && in && out, "null pointer");
  B200R_CHECK_ARG(n > 0 && h >= 2 && w >= 2, "bad shape");
  B200R_CHECK_ARG((size_t)h * ((w + 31) / 32) <= (size_t)kMaxBitWords, "image too large for the spatter water kernel (%d x %d)", h, w);
  const size_t hw = (size_t)h * w;
  int* grad = static_cast<int*>(extra);
  uint8_t* planes = reinterpret_cast<uint8_t*>(grad + (size_t)n * hw);
  spatter_water_kernel<<<n, kWaterThreads, 0, as_stream(stream)>>>(liquid, dist, grad, planes, in, out, h, w, c4);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}
