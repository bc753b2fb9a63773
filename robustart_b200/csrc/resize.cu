// Image resizing exactly as Pillow does it (libImaging/Resample.c for BOX / BILINEAR / HAMMING / BICUBIC / LANCZOS,
// Geometry.c ImagingScaleAffine for NEAREST), uint8 NHWC batches: the ImageNet-S `pil-*` resize types
// (RobustART/noise/utils/imagenet_s_gen.py:19-26,120-141) and the eval transform's Resize + CenterCrop
// (imagenet_dataloader.py:74-80) -- SURVEY 8f N3 / N4.
//
// Bit-exact by construction: the coefficient tables are built on the host with Resample.c's own double arithmetic
// (precompute_coeffs + normalize_coeffs_8bpc: support scaled by the shrink factor, 22-bit fixed point), the device does the
// integer passes -- horizontal, then vertical over a uint8 intermediate holding only the rows the vertical pass reads --
// with Pillow's rounding (2^21 bias, >> 22, clamp).  A crop of the resized image is folded in: only the requested output
// window (and the intermediate rows / columns it needs) is computed.
// HBM-bound integer work: thread = one output pixel (3 channels); taps come through L1 (ksize <= 2*ceil(3*scale)+1).
#include "common.cuh"
#include <map>
#include <mutex>
#include <tuple>
#include <vector>
#include <cmath>

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Resample.c PRECISION_BITS
constexpr int kThreads = 256;

double f_box(double x) { return (x > -0.5 && x <= 0.5) ? 1.0 : 0.0; }
double f_bilinear(double x) { if (x < 0.0) x = -x; return x < 1.0 ? 1.0 - x : 0.0; }
double f_hamming(double x) {
  if (x < 0.0) x = -x;
  if (x == 0.0) return 1.0;
  if (x >= 1.0) return 0.0;
  x = x * M_PI;
  return sin(x) / x * (0.54f + 0.46f * cos(x));
}
double f_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}
double f_sinc(double x) { if (x == 0.0) return 1.0; x = x * M_PI; return sin(x) / x; }
double f_lanczos(double x) { return (-3.0 <= x && x < 3.0) ? f_sinc(x) * f_sinc(x / 3) : 0.0; }

struct HostTable { int ksize = 0; std::vector<int> xmin, xcnt, coef; };

// filter: B200R_RESIZE_* (1 = box .. 5 = lanczos)
HostTable make_table(int in_size, int out_size, int filter) {
  double (*f)(double) = filter == 1 ? f_box : filter == 2 ? f_bilinear : filter == 3 ? f_hamming : filter == 4 ? f_bicubic : f_lanczos;
  const double support0 = filter == 1 ? 0.5 : filter == 2 ? 1.0 : filter == 3 ? 1.0 : filter == 4 ? 2.0 : 3.0;
  HostTable t;
  double scale = (double)in_size / out_size, filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = support0 * filterscale;
  t.ksize = (int)ceil(support) * 2 + 1;
  t.xmin.resize(out_size); t.xcnt.resize(out_size); t.coef.assign((size_t)out_size * t.ksize, 0);
  std::vector<double> k(t.ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) { const double w = f((x + xmin - center + 0.5) * ss); k[x] = w; ww += w; }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      const double v = k[x] * (1 << kPrecisionBits);
      t.coef[(size_t)xx * t.ksize + x] = (int)(v < 0 ? -0.5 + v : 0.5 + v);
    }
    t.xmin[xx] = xmin; t.xcnt[xx] = xmax;
  }
  return t;
}

// Geometry.c ImagingScaleAffine: xo = a*0.5, xin = (int)xo, xo += a (running double sum)
HostTable make_nearest(int in_size, int out_size) {
  HostTable t;
  t.ksize = 1;
  t.xmin.resize(out_size); t.xcnt.assign(out_size, 1); t.coef.assign(out_size, 1 << kPrecisionBits);
  const double a = (double)in_size / out_size;
  double xo = a * 0.5;
  for (int x = 0; x < out_size; ++x) {
    int xin = xo < 0.0 ? -1 : (int)xo;
    if (xin < 0) xin = 0;
    if (xin > in_size - 1) xin = in_size - 1;
    t.xmin[x] = xin;
    xo += a;
  }
  return t;
}

struct DevTable { const int* xmin; const int* xcnt; const int* coef; int ksize; };
struct Entry { DevTable d; std::vector<int> xmin, xcnt; uint64_t last_use = 0; };   // host copy of the bounds: source range of an output window
std::map<std::tuple<int, int, int, int>, Entry> g_tables;            // (device, in, out, filter)
std::mutex g_mu;
uint64_t g_tick = 0;
// A file-backed run resizes every image at its native size: thousands of distinct (in, out) pairs.  The cache is bounded: beyond
// kMaxTables entries the least recently used one is freed (after a device synchronise -- a kernel may still read it), never one of the
// last 16 looked up (a resize call holds two tables at a time).
constexpr size_t kMaxTables = 512;

void evict_lru_locked() {
  auto victim = g_tables.end();
  for (auto it = g_tables.begin(); it != g_tables.end(); ++it)
    if (it->second.last_use + 16 < g_tick && (victim == g_tables.end() || it->second.last_use < victim->second.last_use)) victim = it;
  if (victim == g_tables.end()) return;
  cudaDeviceSynchronize();
  cudaFree(const_cast<int*>(victim->second.d.xmin));
  cudaFree(const_cast<int*>(victim->second.d.xcnt));
  cudaFree(const_cast<int*>(victim->second.d.coef));
  g_tables.erase(victim);
}

int get_table(int in_size, int out_size, int filter, const Entry** out) {
  int dev = 0;
  B200R_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_mu);
  const auto key = std::make_tuple(dev, in_size, out_size, filter);
  auto it = g_tables.find(key);
  if (it == g_tables.end()) {      // first use: blocking upload (not capturable)
    if (g_tables.size() >= kMaxTables) evict_lru_locked();
    const HostTable t = filter == 0 ? make_nearest(in_size, out_size) : make_table(in_size, out_size, filter);
    int *dx, *dc, *dk;
    B200R_CUDA(cudaMalloc(&dx, t.xmin.size() * 4));
    B200R_CUDA(cudaMalloc(&dc, t.xcnt.size() * 4));
    B200R_CUDA(cudaMalloc(&dk, t.coef.size() * 4));
    B200R_CUDA(cudaMemcpy(dx, t.xmin.data(), t.xmin.size() * 4, cudaMemcpyHostToDevice));
    B200R_CUDA(cudaMemcpy(dc, t.xcnt.data(), t.xcnt.size() * 4, cudaMemcpyHostToDevice));
    B200R_CUDA(cudaMemcpy(dk, t.coef.data(), t.coef.size() * 4, cudaMemcpyHostToDevice));
    Entry e{DevTable{dx, dc, dk, t.ksize}, t.xmin, t.xcnt};
    it = g_tables.emplace(key, std::move(e)).first;
  }
  it->second.last_use = ++g_tick;
  *out = &it->second;
  return B200R_OK;
}

// One pass along x (HORIZ) or y of [n][rows][cols][3] uint8 images with explicit pitches.
//   HORIZ: dst[img][r][xx] = clip8(sum_j src[img][r0 + r][xmin[x0 + xx] + j] * coef[x0 + xx][j])          r < n_r, xx < n_x
//   VERT : dst[img][yy][c] = clip8(sum_j src[img][ymin[y0 + yy] - src_row0 + j][c0 + c] * coef[y0 + yy][j])  yy < n_r, c < n_x
template <bool HORIZ>
__global__ void __launch_bounds__(kThreads) resample_kernel(const uint8_t* __restrict__ src, size_t src_img_pitch, int src_row_pitch,
                                                             uint8_t* __restrict__ dst, size_t dst_img_pitch, int dst_row_pitch,
                                                             int n, int n_r, int n_x, int r0, int x0, int src_row0, DevTable t) {
  const size_t total = (size_t)n * n_r * n_x;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
    const int xx = (int)(i % n_x);
    const int r = (int)((i / n_x) % n_r), img = (int)(i / ((size_t)n_x * n_r));
    const int o = HORIZ ? x0 + xx : r0 + r;                    // index into the table
    const int lo = __ldg(t.xmin + o), cnt = __ldg(t.xcnt + o);
    const int* k = t.coef + (size_t)o * t.ksize;
    const uint8_t* p = HORIZ ? src + img * src_img_pitch + (size_t)(r0 + r) * src_row_pitch + (size_t)lo * 3
                             : src + img * src_img_pitch + (size_t)(lo - src_row0) * src_row_pitch + (size_t)(x0 + xx) * 3;
    const int step = HORIZ ? 3 : src_row_pitch;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int j = 0; j < cnt; ++j) {
      const int c = __ldg(k + j);
      a0 += p[0] * c; a1 += p[1] * c; a2 += p[2] * c;
      p += step;
    }
    uint8_t* q = dst + img * dst_img_pitch + (size_t)r * dst_row_pitch + (size_t)xx * 3;
    q[0] = (uint8_t)min(max(a0 >> kPrecisionBits, 0), 255);
    q[1] = (uint8_t)min(max(a1 >> kPrecisionBits, 0), 255);
    q[2] = (uint8_t)min(max(a2 >> kPrecisionBits, 0), 255);
  }
}

unsigned grid_for(size_t items) {
  size_t b = (items + kThreads - 1) / kThreads;
  const size_t cap = (size_t)b200r_num_sms() * 16;
  return (unsigned)(b < cap ? (b ? b : 1) : cap);
}

struct Plan { bool need_h, need_v; int first, last; };   // first/last: source rows the vertical pass of the window reads

int make_plan(int hin, int win, int hout, int wout, int filter, int oy0, int ch, const Entry** th, const Entry** tv, Plan* p) {
  p->need_h = wout != win;
  p->need_v = hout != hin;
  *th = *tv = nullptr;
  if (p->need_h) { int rc = get_table(win, wout, filter, th); if (rc) return rc; }
  p->first = oy0; p->last = oy0 + ch;
  if (p->need_v) {
    int rc = get_table(hin, hout, filter, tv);
    if (rc) return rc;
    p->first = (*tv)->xmin[oy0];
    p->last = (*tv)->xmin[oy0 + ch - 1] + (*tv)->xcnt[oy0 + ch - 1];
  }
  return B200R_OK;
}

}  // namespace

extern "C" {

int b200r_resize_workspace_bytes(int n, int hin, int win, int hout, int wout, int filter, int oy0, int ox0, int ch, int cw, size_t* bytes) {
  B200R_CHECK_ARG(bytes, "null pointer");
  B200R_CHECK_ARG(n > 0 && hin > 0 && win > 0 && hout > 0 && wout > 0 && filter >= 0 && filter <= 5, "bad shape / filter");
  B200R_CHECK_ARG(oy0 >= 0 && ox0 >= 0 && ch > 0 && cw > 0 && oy0 + ch <= hout && ox0 + cw <= wout, "crop window outside the resized image");
  const Entry *th, *tv;
  Plan p;
  int rc = make_plan(hin, win, hout, wout, filter, oy0, ch, &th, &tv, &p);
  if (rc) return rc;
  *bytes = (p.need_h && p.need_v) ? (size_t)n * (p.last - p.first) * cw * 3 : 0;
  return B200R_OK;
}

int b200r_resize_u8(const uint8_t* in, uint8_t* out, int n, int hin, int win, int hout, int wout, int filter, int oy0, int ox0,
                    int ch, int cw, void* workspace, size_t ws_bytes, b200r_stream_t stream) {
  B200R_CHECK_ARG(in && out, "null pointer");
  size_t need = 0;
  int rc = b200r_resize_workspace_bytes(n, hin, win, hout, wout, filter, oy0, ox0, ch, cw, &need);
  if (rc) return rc;
  if (need > ws_bytes || (need && !workspace)) { b200r_set_error("resize workspace too small: need %zu bytes, got %zu", need, ws_bytes); return B200R_ENOSPC; }
  const Entry *th, *tv;
  Plan p;
  rc = make_plan(hin, win, hout, wout, filter, oy0, ch, &th, &tv, &p);
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  const size_t in_img = (size_t)hin * win * 3, out_img = (size_t)ch * cw * 3;
  if (!p.need_h && !p.need_v) {          // Image.resize returns a copy: the crop window, row by row
    B200R_CUDA(cudaMemcpy2DAsync(out, (size_t)cw * 3, in + ((size_t)oy0 * win + ox0) * 3, (size_t)win * 3, (size_t)cw * 3, (size_t)ch, cudaMemcpyDeviceToDevice, s));
    for (int i = 1; i < n; ++i)
      B200R_CUDA(cudaMemcpy2DAsync(out + i * out_img, (size_t)cw * 3, in + i * in_img + ((size_t)oy0 * win + ox0) * 3, (size_t)win * 3, (size_t)cw * 3,
                                   (size_t)ch, cudaMemcpyDeviceToDevice, s));
    return B200R_OK;
  }
  if (p.need_h && !p.need_v) {
    resample_kernel<true><<<grid_for((size_t)n * ch * cw), kThreads, 0, s>>>(in, in_img, win * 3, out, out_img, cw * 3, n, ch, cw, oy0, ox0, 0, th->d);
  } else if (!p.need_h) {
    resample_kernel<false><<<grid_for((size_t)n * ch * cw), kThreads, 0, s>>>(in, in_img, win * 3, out, out_img, cw * 3, n, ch, cw, oy0, ox0, 0, tv->d);
  } else {
    uint8_t* tmp = static_cast<uint8_t*>(workspace);
    const int rows = p.last - p.first;
    const size_t tmp_img = (size_t)rows * cw * 3;
    resample_kernel<true><<<grid_for((size_t)n * rows * cw), kThreads, 0, s>>>(in, in_img, win * 3, tmp, tmp_img, cw * 3, n, rows, cw, p.first, ox0, 0, th->d);
    B200R_LAUNCH_CHECK();
    resample_kernel<false><<<grid_for((size_t)n * ch * cw), kThreads, 0, s>>>(tmp, tmp_img, cw * 3, out, out_img, cw * 3, n, ch, cw, oy0, 0, p.first, tv->d);
  }
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

}  // extern "C"
