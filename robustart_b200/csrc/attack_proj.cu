// Per-sample constrained projections of the attacks, sort-free: one CTA per sample row, streaming reductions.
//   K-FABproj  FAB's projection_linf  (Attacks/autoattack/fab_projections.py:7-59, called fab_base.py:194-198)
//   FAB's convex-combination update    (fab_base.py:200-232)
//   K-L1proj   APGD-L1's L1_projection (Attacks/autoattack/autopgd_base.py:19-83)
//   PGD-L1 step of ART's ProjectedGradientDescentPyTorch(norm=1) (called from adv/attack.py:44-49; ART is not vendored)
//
// The reference solves each projection by sorting D (or 2 D) values per row, cumulative sums and a log2(D)-step index
// bisection.  Both problems are "find the threshold lambda with  sum_i phi_i(lambda) = target"  for a monotone piecewise-
// linear sum, so a row needs no sort: a few passes of (evaluate the sum and its slope at lambda) over the row -- Newton
// from the left on the concave FAB sum (finite, never overshoots), safeguarded Newton / bisection on the L1 sum -- then
// one pass that writes the result.  Sums are carried in double, so lambda is exact to 1e-15 relative; the reference's
// own float32 cumulative sums are the larger error (parity bar 1e-6, tests/test_autoattack_gpu.py).
// Traffic per row and pass: 2 x D x 4 B (w and t re-read from L2 / HBM); D = 150 528 -> 1.2 MB, typically 4-8 passes.
#include "common.cuh"

namespace {
constexpr int kPT = 1024;          // threads per row

struct D3 { double a, b, c; };

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sums of up to N doubles; every thread gets the totals.  sbuf: [32][N] doubles
template <int N>
__device__ __forceinline__ void block_sum_d(double (&v)[N], double* sbuf) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = warp_sum_d(v[k]);
  __syncthreads();                    // sbuf may still be read from the previous call
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < N; ++k) sbuf[wid * N + k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double t = (lane < nw) ? sbuf[lane * N + k] : 0.0;
    v[k] = warp_sum_d(t);
  }
}
__device__ __forceinline__ float block_min_f(float v, float* sbuf) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_min_f(v);
  __syncthreads();
  if (lane == 0) sbuf[wid] = v;
  __syncthreads();
  return warp_min_f(lane < nw ? sbuf[lane] : 3.4e38f);
}
__device__ __forceinline__ float block_max_f(float v, float* sbuf) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max_f(v);
  __syncthreads();
  if (lane == 0) sbuf[wid] = v;
  __syncthreads();
  return warp_max_f(lane < nw ? sbuf[lane] : -3.4e38f);
}

// =============================================================================================
// FAB projection_linf.  Row r: point t[r, :], hyperplane <w[r, :], x> = b[r].  With s = sign(<w,t> - b) (>= 0 -> +1),
// w' = s w, B = s (<w,t> - b) >= 0: coordinate i can move by at most p_i = |a_i - t_i| towards its bound a_i = [w'_i < 0],
// and an Linf budget lambda buys  g(lambda) = sum_i |w_i| min(lambda, p_i).  The projection is lambda* with g = B
// (d_i = (2 a_i - 1) min(lambda*, p_i)), every coordinate at its bound when g(max p) <= B.  fab_projections.py:7-59 finds
// lambda* through argsort / cumsum / bisection; here: Newton from lambda = 0 on the concave g.
// =============================================================================================
__global__ void __launch_bounds__(kPT) fab_projection_linf_kernel(const float* __restrict__ t, const float* __restrict__ w,
                                                                   const float* __restrict__ b, float* __restrict__ d,
                                                                   float* __restrict__ dmax, int D, int* __restrict__ passes_out) {
  __shared__ double sbuf_d[32 * 4];
  __shared__ float sbuf_f[32];
  const int row = blockIdx.x;
  const float* tr = t + (size_t)row * D;
  const float* wr = w + (size_t)row * D;
  float* dr = d + (size_t)row * D;
  const int D4 = (D % 4 == 0 && ((reinterpret_cast<uintptr_t>(tr) | reinterpret_cast<uintptr_t>(wr) | reinterpret_cast<uintptr_t>(dr)) & 15) == 0) ? D / 4 : 0;

  // ---- pass 0: <w,t>, sum |w|, and for both possible signs sum |w| p and min p ----------------------------------
  double acc[4] = {0, 0, 0, 0};          // <w,t>, sum|w|, sum|w| p (sign +), sum|w| p (sign -)
  float minp = 3.4e38f, minm = 3.4e38f;
  auto visit0 = [&](float wi, float ti) {
    const float aw = fabsf(wi);
    const float pp = (wi < 0.f) ? 1.f - ti : ti;       // sign +: a = [w < 0]
    const float pm = (wi > 0.f) ? 1.f - ti : ti;       // sign -: a = [-w < 0]
    acc[0] += (double)wi * ti; acc[1] += aw; acc[2] += (double)aw * pp; acc[3] += (double)aw * pm;
    minp = fminf(minp, pp); minm = fminf(minm, pm);
  };
  for (int i = threadIdx.x; i < D4; i += kPT) {
    const float4 w4 = reinterpret_cast<const float4*>(wr)[i], t4 = reinterpret_cast<const float4*>(tr)[i];
    visit0(w4.x, t4.x); visit0(w4.y, t4.y); visit0(w4.z, t4.z); visit0(w4.w, t4.w);
  }
  for (int i = 4 * D4 + threadIdx.x; i < D; i += kPT) visit0(wr[i], tr[i]);
  block_sum_d<4>(acc, sbuf_d);
  minp = block_min_f(minp, sbuf_f);
  minm = block_min_f(minm, sbuf_f);
  const double wt = acc[0], sw = acc[1];
  const bool pos = (wt - (double)b[row]) >= 0.0;
  const double B = pos ? (wt - (double)b[row]) : ((double)b[row] - wt);
  const double smax = pos ? acc[2] : acc[3];             // g at lambda = max p: everything at its bound
  const float pmin = pos ? minp : minm;
  // the reference's three cases: c_l (no coordinate saturates), c2 (threshold search), else all coordinates at their bounds
  const bool c_l = (sw * (double)pmin - B) > 0.0;
  const bool c2 = (smax - B > 0.0) && !c_l;
  double lam = 0.0;
  int passes = 1;
  if (c_l) {
    lam = sw > 0.0 ? fmax(B / sw, 0.0) : 0.0;
  } else if (c2) {
    // Newton from the left: g is concave, so the tangent at lambda_k meets B at lambda_{k+1} <= lambda*; when no breakpoint lies
    // in between the step is exact.  Terminates when lambda stops moving (a fixed point of the reference's closed form).
    for (int it = 0; it < 64; ++it) {
      double s[2] = {0, 0};            // sum_{p <= lam} |w| p ,  sum_{p > lam} |w|
      const float lf = (float)lam;
      auto visit = [&](float wi, float ti) {
        const float wsgn = pos ? wi : -wi;
        const float p = (wsgn < 0.f) ? 1.f - ti : ti;
        const float aw = fabsf(wi);
        if ((double)p <= lam) s[0] += (double)aw * p; else s[1] += aw;
      };
      (void)lf;
      for (int i = threadIdx.x; i < D4; i += kPT) {
        const float4 w4 = reinterpret_cast<const float4*>(wr)[i], t4 = reinterpret_cast<const float4*>(tr)[i];
        visit(w4.x, t4.x); visit(w4.y, t4.y); visit(w4.z, t4.z); visit(w4.w, t4.w);
      }
      for (int i = 4 * D4 + threadIdx.x; i < D; i += kPT) visit(wr[i], tr[i]);
      block_sum_d<2>(s, sbuf_d);
      ++passes;
      if (!(s[1] > 0.0)) break;                         // every coordinate saturated (cannot happen while g(max p) > B)
      const double nl = fmax((B - s[0]) / s[1], 0.0);
      if (!(nl > lam * (1.0 + 1e-15))) { lam = fmax(nl, lam); break; }
      lam = nl;
    }
  }
  // ---- final pass: write d (and the row's max |d| for FAB's step-size rule) ---------------------------------------
  float mx = 0.f;
  const float lamf = (float)lam;
  auto outv = [&](float wi, float ti) -> float {
    if (wi == 0.f) return 0.f;
    const float wsgn = pos ? wi : -wi;
    const bool a = wsgn < 0.f;
    const float dm = (a ? 1.f : 0.f) - ti;               // step to the bound: (a - t)
    float v;
    if (c_l) v = a ? lamf : -lamf;
    else if (c2) v = a ? fminf(lamf, dm) : fmaxf(-lamf, dm);
    else v = dm;
    mx = fmaxf(mx, fabsf(v));
    return v;
  };
  for (int i = threadIdx.x; i < D4; i += kPT) {
    const float4 w4 = reinterpret_cast<const float4*>(wr)[i], t4 = reinterpret_cast<const float4*>(tr)[i];
    reinterpret_cast<float4*>(dr)[i] = make_float4(outv(w4.x, t4.x), outv(w4.y, t4.y), outv(w4.z, t4.z), outv(w4.w, t4.w));
  }
  for (int i = 4 * D4 + threadIdx.x; i < D; i += kPT) dr[i] = outv(wr[i], tr[i]);
  if (dmax) {
    mx = block_max_f(mx, sbuf_f);
    if (threadIdx.x == 0) dmax[row] = mx;
  }
  if (passes_out && threadIdx.x == 0) passes_out[row] = passes + 1;
}

// FAB's update (fab_base.py:200-232): a1, a2 = max(|d1|), max(|d2|) clamped at 1e-8, alpha = min(max(a1 / (a1 + a2), 0), alpha_max),
// x1 <- clamp((x1 + eta d1) (1 - alpha) + (x0 + eta d2) alpha, 0, 1).  One block-strided pass; grid.y = sample.
__global__ void __launch_bounds__(256) fab_combine_kernel(float* __restrict__ x1, const float* __restrict__ d1, const float* __restrict__ x0,
                                                          const float* __restrict__ d2, const float* __restrict__ dmax1,
                                                          const float* __restrict__ dmax2, int D, float eta, float alpha_max) {
  const int row = blockIdx.y;
  const float a1 = fmaxf(dmax1[row], 1e-8f), a2 = fmaxf(dmax2[row], 1e-8f);
  const float alpha = fminf(fmaxf(a1 / (a1 + a2), 0.f), alpha_max), om = 1.f - alpha;
  const size_t base = (size_t)row * D;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < D; i += gridDim.x * 256) {
    // torch evaluates (x1 + eta*d1) * (1 - alpha) + (x0 + d2*eta) * alpha with one rounding per operation
    const float u = __fmul_rn(__fadd_rn(x1[base + i], __fmul_rn(eta, d1[base + i])), om);
    const float v = __fmul_rn(__fadd_rn(x0[base + i], __fmul_rn(d2[base + i], eta)), alpha);
    x1[base + i] = fminf(fmaxf(__fadd_rn(u, v), 0.f), 1.f);
  }
}

// =============================================================================================
// L1_projection (autopgd_base.py:19-83): x = centre of the L1 ball, y = current perturbation; returns delta with
// ||y + delta||_1 <= eps and 0 <= x + y + delta <= 1.  Per coordinate: lo_i = max(0, -min(1 - x - y, x + y)) is the shrink the box
// demands, hi_i = |y_i| the shrink that zeroes the coordinate; delta_i = -sign(y_i) clamp(alpha, lo_i, hi_i) with alpha = 0 when
// sum |y| - sum lo <= eps, else the root of  h(alpha) = sum clamp(alpha, lo_i, hi_i) = sum |y| - eps  (h is monotone, piecewise
// linear).  The reference sorts the 2 D breakpoints; here: safeguarded Newton on the bracket [0, max hi].
// =============================================================================================
__global__ void __launch_bounds__(kPT) l1_projection_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ delta,
                                                            int D, float eps) {
  __shared__ double sbuf_d[32 * 3];
  __shared__ float sbuf_f[32];
  const int row = blockIdx.x;
  const float* xr = x + (size_t)row * D;
  const float* yr = y + (size_t)row * D;
  float* dr = delta + (size_t)row * D;
  auto lo_of = [](float xi, float yi) { return fmaxf(0.f, -fminf(1.f - xi - yi, xi + yi)); };   // -u, u = min(0, min(1-x-y, x+y))
  double s0[3] = {0, 0, 0};     // sum |y|, sum lo
  float hmax = 0.f;
  for (int i = threadIdx.x; i < D; i += kPT) {
    const float xi = xr[i], yi = yr[i];
    s0[0] += fabsf(yi); s0[1] += lo_of(xi, yi);
    hmax = fmaxf(hmax, fabsf(yi));
  }
  block_sum_d<3>(s0, sbuf_d);
  hmax = block_max_f(hmax, sbuf_f);
  const double c = (double)eps - s0[0];               // eps - ||y||_1
  const bool need = (s0[1] + c) < 0.0;                 // c5: after the box shrink the point is still outside the ball
  double alpha = 0.0;
  if (need) {
    const double T = -c;                               // h(alpha) must reach sum|y| - eps
    double L = 0.0, U = hmax;
    alpha = 0.0;
    for (int it = 0; it < 100; ++it) {
      double s[3] = {0, 0, 0};                         // h(alpha), number of coordinates with lo < alpha < hi (slope)
      for (int i = threadIdx.x; i < D; i += kPT) {
        const float xi = xr[i], yi = yr[i];
        const double lo = lo_of(xi, yi), hi = fabsf(yi);
        // the reference clamps with max(-u, alpha) then min(., -l): lo wins below, hi wins above (also when lo > hi)
        s[0] += fmin(fmax(lo, alpha), hi);
        if (alpha > lo && alpha < hi) s[1] += 1.0;
      }
      block_sum_d<3>(s, sbuf_d);
      const double r = s[0] - T;
      if (r < 0.0) L = alpha; else U = alpha;
      if (fabs(r) <= 1e-13 * fmax(1.0, fabs(T)) || !(U - L > 1e-14 * fmax(U, 1e-30))) break;
      double nxt = (s[1] > 0.0) ? alpha - r / s[1] : 0.5 * (L + U);
      if (!(nxt > L && nxt < U)) nxt = 0.5 * (L + U);
      alpha = nxt;
    }
  }
  const float af = (float)alpha;
  for (int i = threadIdx.x; i < D; i += kPT) {
    const float xi = xr[i], yi = yr[i];
    const float lo = lo_of(xi, yi), hi = fabsf(yi);
    const float dd = need ? -fminf(fmaxf(lo, af), hi) : -lo;            // d = u, or -min(max(-u, alpha), -l)
    const float sg = (yi > 0.f) ? 1.f : ((yi < 0.f) ? -1.f : 0.f);
    dr[i] = sg * dd;
  }
}

// =============================================================================================
// PGD-L1 step of ART's ProjectedGradientDescentPyTorch (norm = 1; art/attacks/evasion/projected_gradient_descent/
// projected_gradient_descent_pytorch.py, _compute_perturbation / _apply_perturbation / _projection as of ART 1.7-1.16):
//   g <- g / (||g||_1 + 1e-7);  x <- clip(x + eps_step g, 0, 1);  delta = x - x0;  delta <- delta min(1, eps / (||delta||_1 + 1e-7));
//   x <- x0 + delta.      Two reductions per sample -> one CTA per sample, three passes over the row.
// =============================================================================================
__global__ void __launch_bounds__(kPT) pgd_step_l1_kernel(float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ x0,
                                                          int D, float eps_step, float eps) {
  __shared__ double sbuf_d[32];
  const int row = blockIdx.x;
  float* xr = x + (size_t)row * D;
  const float* gr = g + (size_t)row * D;
  const float* x0r = x0 + (size_t)row * D;
  double s[1] = {0};
  for (int i = threadIdx.x; i < D; i += kPT) s[0] += fabsf(gr[i]);
  block_sum_d<1>(s, sbuf_d);
  const float step = eps_step / ((float)s[0] + 1e-7f);
  double n1[1] = {0};
  for (int i = threadIdx.x; i < D; i += kPT) {
    const float v = fminf(fmaxf(__fadd_rn(xr[i], __fmul_rn(step, gr[i])), 0.f), 1.f);
    xr[i] = v;
    n1[0] += fabsf(v - x0r[i]);
  }
  block_sum_d<1>(n1, sbuf_d);
  const float scale = fminf(1.f, eps / ((float)n1[0] + 1e-7f));
  for (int i = threadIdx.x; i < D; i += kPT) xr[i] = __fadd_rn(x0r[i], __fmul_rn(xr[i] - x0r[i], scale));
}

}  // namespace

extern "C" {

int b200r_fab_projection_linf(const float* t, const float* w, const float* b, float* d, float* dmax, int rows, int dim, int* passes,
                              b200r_stream_t stream) {
  B200R_CHECK_ARG(t && w && b && d, "null pointer");
  B200R_CHECK_ARG(rows >= 0 && dim > 0, "bad shape %d x %d", rows, dim);
  if (rows == 0) return B200R_OK;
  fab_projection_linf_kernel<<<rows, kPT, 0, as_stream(stream)>>>(t, w, b, d, dmax, dim, passes);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_fab_combine_linf(float* x1, const float* d1, const float* x0, const float* d2, const float* dmax1, const float* dmax2, int rows,
                           int dim, float eta, float alpha_max, b200r_stream_t stream) {
  B200R_CHECK_ARG(x1 && d1 && x0 && d2 && dmax1 && dmax2, "null pointer");
  B200R_CHECK_ARG(rows >= 0 && rows < 65536 && dim > 0, "bad shape %d x %d", rows, dim);
  if (rows == 0) return B200R_OK;
  int bx = (dim + 256 * 8 - 1) / (256 * 8);
  if (bx < 1) bx = 1;
  fab_combine_kernel<<<dim3(bx, rows), 256, 0, as_stream(stream)>>>(x1, d1, x0, d2, dmax1, dmax2, dim, eta, alpha_max);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_l1_projection(const float* x, const float* y, float* delta, int rows, int dim, float eps, b200r_stream_t stream) {
  B200R_CHECK_ARG(x && y && delta, "null pointer");
  B200R_CHECK_ARG(rows >= 0 && dim > 0, "bad shape %d x %d", rows, dim);
  if (rows == 0) return B200R_OK;
  l1_projection_kernel<<<rows, kPT, 0, as_stream(stream)>>>(x, y, delta, dim, eps);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_pgd_step_l1(float* x, const float* g, const float* x0, int rows, int dim, float eps_step, float eps, b200r_stream_t stream) {
  B200R_CHECK_ARG(x && g && x0, "null pointer");
  B200R_CHECK_ARG(rows >= 0 && dim > 0, "bad shape %d x %d", rows, dim);
  if (rows == 0) return B200R_OK;
  pgd_step_l1_kernel<<<rows, kPT, 0, as_stream(stream)>>>(x, g, x0, dim, eps_step, eps);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

}  // extern "C"
