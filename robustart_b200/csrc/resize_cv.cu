// cv2.resize for uint8 NHWC batches, INTER_NEAREST and INTER_LINEAR, bit for bit (OpenCV imgproc/resize.cpp restated in
// oracle/cv_resize.py and pinned there against cv2 4.13): the `opencv-nearest` / `opencv-bilinear` resize types of the reference's
// ImageNet-S generator (RobustART/noise/utils/imagenet_s_gen.py:28-34,120-148), with the centre crop of the 'val' transform fused
// (only the cropped window is produced).
//   linear: coordinates f = float((d + 0.5) * scale - 0.5), scale = 1 / (dst / src) in double; weights cvRound(w * 2048);
//           horizontal pass  S = p[sx] * a0 + p[sx + 1] * a1   (weight zeroed where the pair leaves the row)
//           vertical pass    (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2   (row indices clipped, weights kept)
//   nearest: src = min(floor(d * scale), size - 1)
// One thread per output pixel (3 channels), weights recomputed per thread (a dozen flops against four 3-byte gathers): the kernel
// is bound by the gathers, which hit L1/L2 for every down-scaling ratio the eval transform sees.
// STATUS: checked from source on the host emulator against cv2.resize (tests/test_kernel_emulation_cpu.py); not yet run on a GPU.
#include "common.cuh"

namespace {
constexpr int kThreads = 256;

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

template <bool LINEAR>
__global__ void __launch_bounds__(kThreads) resize_cv_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n, int hin, int win,
                                                              double sy, double sx, int oy0, int ox0, int ch, int cw) {
  const size_t total = (size_t)n * ch * cw;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int x = (int)(t % cw), y = (int)((t / cw) % ch), im = (int)(t / ((size_t)cw * ch));
    const uint8_t* src = in + (size_t)im * hin * win * 3;
    uint8_t* dst = out + t * 3;
    if (!LINEAR) {
      const int xs = min((int)floor((double)(x + ox0) * sx), win - 1), ys = min((int)floor((double)(y + oy0) * sy), hin - 1);
      const uint8_t* p = src + ((size_t)ys * win + xs) * 3;
      dst[0] = p[0]; dst[1] = p[1]; dst[2] = p[2];
      continue;
    }
    const Tap tx = linear_tap(x + ox0, sx, win, true), ty = linear_tap(y + oy0, sy, hin, false);
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
  B200R_CHECK_ARG(interpolation == B200R_CV_INTER_NEAREST || interpolation == B200R_CV_INTER_LINEAR,
                  "interpolation %d not supported (cv2.INTER_NEAREST = 0, cv2.INTER_LINEAR = 1)", interpolation);
  B200R_CHECK_ARG(oy0 >= 0 && ox0 >= 0 && ch > 0 && cw > 0 && oy0 + ch <= hout && ox0 + cw <= wout, "crop window outside the resized image");
  const double sy = 1.0 / ((double)hout / (double)hin), sx = 1.0 / ((double)wout / (double)win);
  const size_t total = (size_t)n * ch * cw;
  size_t blocks = (total + kThreads - 1) / kThreads;
  const size_t cap = (size_t)b200r_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (interpolation == B200R_CV_INTER_LINEAR)
    resize_cv_kernel<true><<<(unsigned)blocks, kThreads, 0, as_stream(stream)>>>(in, out, n, hin, win, sy, sx, oy0, ox0, ch, cw);
  else
    resize_cv_kernel<false><<<(unsigned)blocks, kThreads, 0, as_stream(stream)>>>(in, out, n, hin, win, sy, sx, oy0, ox0, ch, cw);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}
