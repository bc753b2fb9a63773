// Fused elementwise kernels of the gradient attacks + layout/normalisation (all HBM-bound).
//   foolbox 3.3.1 gradient_descent_base.py run(): x += a*normalize(g); x = project(x,x0,eps);
//   x = clip(x, 0, 1)   (called from RobustART/noise/utils/adv/attack.py:20-33)
//   MI-FGSM: RobustART/noise/utils/adv/Attacks/imfgsm_attack.py:85-90
//   normalize(): prototype/prototype/solver/benchmark_eval_adv.py:33-46
// Algorithmic traffic per image per step (fp32, 3x224x224): Linf read x,g,x0 + write x =
// 4 * 602 112 B; L2 / MIM add one extra read of g for the per-sample norm.
#include "common.cuh"
#include "corrupt.cuh"
#ifdef __CUDACC__                       // (the host emulator of tests/emu compiles this file without the cluster kernels)
#include <cooperative_groups.h>
#endif

namespace {
constexpr int kThreads = 256;

__device__ __forceinline__ float sgn(float g) { return (g > 0.f) ? 1.f : ((g < 0.f) ? -1.f : 0.f); }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// ---- u8 NHWC -> float NCHW, (x/255 - mean)/std -----------------------------------------------
struct Norm3 { float mean[3], inv_std[3], std[3]; };

// A thread converts 16 pixels (48 bytes in, three 64-byte runs out) by table look-up: the 96 IEEE divisions per thread of the direct
// form made the kernel ALU-bound; the 3 x 256 table is built with the same two divisions, so results stay bit-identical.  Written straight from the registers, every store instruction
// would put 16 bytes per lane at a 64-byte stride -- half-filled 32-byte sectors, the pattern that halves a kernel's store rate --
// so the CTA's 3 x 2048 floats go through shared memory (16-byte chunks, XOR-swizzled: conflict-free both ways) and leave as
// 512 contiguous bytes per warp and instruction.
constexpr int kCvtThreads = 128;
__global__ void __launch_bounds__(kCvtThreads) u8nhwc_to_f32nchw_kernel(const uint4* __restrict__ in,
                                                                         float* __restrict__ out,
                                                                         uint32_t groups_per_image,
                                                                         uint32_t hw, Norm3 nm) {
  __shared__ float4 stg[3][kCvtThreads * 4];
  __shared__ float lut[3][256];                             // every value a byte can become, with the reference's two true divisions
  for (int e = threadIdx.x; e < 768; e += kCvtThreads) {
    const int ch = e >> 8;
    // torchvision ToTensor: x/255 (true division), Normalize: (t - mean)/std
    lut[ch][e & 255] = (__fdiv_rn((float)(e & 255), 255.0f) - nm.mean[ch]) / nm.std[ch];
  }
  __syncthreads();
  const uint32_t img = blockIdx.y;
  const uint32_t g0 = blockIdx.x * kCvtThreads, gi = g0 + threadIdx.x;
  if (gi < groups_per_image) {
    const uint4* p = in + ((size_t)img * groups_per_image + gi) * 3;
    uint4 a = ld_stream_u4(p), b = ld_stream_u4(p + 1), c = ld_stream_u4(p + 2);
    uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    float v[3][16];
#pragma unroll
    for (int byte = 0; byte < 48; ++byte) {
      const int ch = byte % 3, px = byte / 3;
      v[ch][px] = lut[ch][(w[byte >> 2] >> (8 * (byte & 3))) & 0xFFu];
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t ck = 4 * threadIdx.x + q;
        stg[ch][ck ^ ((ck >> 3) & 7)] = make_float4(v[ch][4 * q], v[ch][4 * q + 1], v[ch][4 * q + 2], v[ch][4 * q + 3]);
      }
  }
  __syncthreads();
  const uint32_t left = groups_per_image - g0;
  const uint32_t chunks = (left < (uint32_t)kCvtThreads ? left : (uint32_t)kCvtThreads) * 4;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float* o = out + ((size_t)img * 3 + ch) * hw + (size_t)g0 * 16;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t ck = threadIdx.x + kCvtThreads * k;
      if (ck < chunks) st_stream_f4(o + 4 * ck, stg[ch][ck ^ ((ck >> 3) & 7)]);
    }
  }
}

__global__ void __launch_bounds__(kThreads) normalize_kernel(const float4* __restrict__ in,
                                                              float4* __restrict__ out, uint32_t hw4,
                                                              Norm3 nm, int inverse) {
  const uint32_t plane = blockIdx.y;  // n*3 + c
  const int ch = plane % 3;
  const float m = nm.mean[ch], s = nm.std[ch];
  for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < hw4; i += gridDim.x * kThreads) {
    float4 v = ld_stream_f4(in + (size_t)plane * hw4 + i);
    if (inverse == 2) {
      v.x = v.x / s; v.y = v.y / s; v.z = v.z / s; v.w = v.w / s;
    } else if (inverse) {  // x * std + mean  (no fma contraction: match torch's two roundings)
      v.x = __fadd_rn(__fmul_rn(v.x, s), m); v.y = __fadd_rn(__fmul_rn(v.y, s), m);
      v.z = __fadd_rn(__fmul_rn(v.z, s), m); v.w = __fadd_rn(__fmul_rn(v.w, s), m);
    } else {
      v.x = (v.x - m) / s; v.y = (v.y - m) / s; v.z = (v.z - m) / s; v.w = (v.w - m) / s;
    }
    st_stream_f4(out + (size_t)plane * hw4 + i, v);
  }
}

// ---- Linf ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) random_start_linf_kernel(const float4* __restrict__ x0,
                                                                      float4* __restrict__ x,
                                                                      const float4* __restrict__ u,
                                                                      size_t chw4, float eps, uint32_t k0,
                                                                      uint32_t k1, uint64_t image_offset,
                                                                      int clip01) {
  const size_t img = blockIdx.y;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < chw4; i += (size_t)gridDim.x * kThreads) {
    float4 a = ld_stream_f4(x0 + img * chw4 + i), r;
    if (u) {
      r = ld_stream_f4(u + img * chw4 + i);
    } else {
      uint4 p = philox4x32_10(rng_counter((uint32_t)i, RNG_PGD, 0, image_offset + img), k0, k1);
      r = make_float4(u32_to_unit(p.x), u32_to_unit(p.y), u32_to_unit(p.z), u32_to_unit(p.w));
    }
    const float two_eps = 2.f * eps;
    // uniform(-eps, eps) = (hi - lo)*u + lo ; then x0 + that ; then clip to [0,1]
    a.x = __fadd_rn(a.x, __fadd_rn(__fmul_rn(two_eps, r.x), -eps));
    a.y = __fadd_rn(a.y, __fadd_rn(__fmul_rn(two_eps, r.y), -eps));
    a.z = __fadd_rn(a.z, __fadd_rn(__fmul_rn(two_eps, r.z), -eps));
    a.w = __fadd_rn(a.w, __fadd_rn(__fmul_rn(two_eps, r.w), -eps));
    if (clip01) {
      a.x = clampf(a.x, 0.f, 1.f); a.y = clampf(a.y, 0.f, 1.f);
      a.z = clampf(a.z, 0.f, 1.f); a.w = clampf(a.w, 0.f, 1.f);
    }
    st_stream_f4(x + img * chw4 + i, a);
  }
}

__device__ __forceinline__ float linf_update(float x, float s_alpha, float x0, float eps) {
  float t = __fadd_rn(x, s_alpha);                   // x + alpha*sign(g)
  float d = clampf(__fsub_rn(t, x0), -eps, eps);     // clip(x - x0, -eps, eps)
  return clampf(__fadd_rn(x0, d), 0.f, 1.f);         // clip(x0 + d, 0, 1)
}

__global__ void __launch_bounds__(kThreads) pgd_step_linf_kernel(float4* __restrict__ x,
                                                                  const float4* __restrict__ g,
                                                                  const float4* __restrict__ x0,
                                                                  size_t total4, float alpha, float eps) {
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total4; i += (size_t)gridDim.x * kThreads) {
    float4 xv = x[i];
    float4 gv = ld_stream_f4(g + i), ov = ld_stream_f4(x0 + i);
    xv.x = linf_update(xv.x, alpha * sgn(gv.x), ov.x, eps);
    xv.y = linf_update(xv.y, alpha * sgn(gv.y), ov.y, eps);
    xv.z = linf_update(xv.z, alpha * sgn(gv.z), ov.z, eps);
    xv.w = linf_update(xv.w, alpha * sgn(gv.w), ov.w, eps);
    x[i] = xv;
  }
}

// ---- per-sample reductions -------------------------------------------------------------------
// mode 0: sum g^2 ; mode 1: sum |g|.  grid (slices, n); atomicAdd into acc[n] (pre-zeroed)
template <int MODE>
__global__ void __launch_bounds__(kThreads) sample_reduce_kernel(const float4* __restrict__ g,
                                                                  float* __restrict__ acc, size_t chw4) {
  const size_t img = blockIdx.y;
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < chw4; i += (size_t)gridDim.x * kThreads) {
    float4 v = __ldg(g + img * chw4 + i);
    if (MODE == 0) s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    else s += fabsf(v.x) + fabsf(v.y) + fabsf(v.z) + fabsf(v.w);
  }
  s = warp_sum(s);
  __shared__ float sb[kThreads / 32];
  if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) t += sb[i];
    atomicAdd(&acc[img], t);
  }
}

// L2 phase 2: x = x + alpha * g / max(||g||,1e-12); accumulate ||x - x0||^2
__global__ void __launch_bounds__(kThreads) l2_ascent_kernel(float4* __restrict__ x,
                                                              const float4* __restrict__ g,
                                                              const float4* __restrict__ x0,
                                                              const float* __restrict__ gnorm2,
                                                              float* __restrict__ dnorm2, size_t chw4,
                                                              float alpha) {
  const size_t img = blockIdx.y;
  const float f = alpha * (1.0f / fmaxf(sqrtf(gnorm2[img]), 1e-12f));
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < chw4; i += (size_t)gridDim.x * kThreads) {
    const size_t k = img * chw4 + i;
    float4 xv = x[k], gv = ld_stream_f4(g + k), ov = __ldg(x0 + k);
    xv.x = __fadd_rn(xv.x, __fmul_rn(f, gv.x)); xv.y = __fadd_rn(xv.y, __fmul_rn(f, gv.y));
    xv.z = __fadd_rn(xv.z, __fmul_rn(f, gv.z)); xv.w = __fadd_rn(xv.w, __fmul_rn(f, gv.w));
    float dx = xv.x - ov.x, dy = xv.y - ov.y, dz = xv.z - ov.z, dw = xv.w - ov.w;
    s += dx * dx + dy * dy + dz * dz + dw * dw;
    x[k] = xv;
  }
  s = warp_sum(s);
  __shared__ float sb[kThreads / 32];
  if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) t += sb[i];
    atomicAdd(&dnorm2[img], t);
  }
}

// L2 phase 3: x = clip01(x0 + (x - x0) * min(1, eps / max(||d||,1e-12)))
__global__ void __launch_bounds__(kThreads) l2_project_kernel(float4* __restrict__ x,
                                                               const float4* __restrict__ x0,
                                                               const float* __restrict__ dnorm2,
                                                               size_t chw4, float eps) {
  const size_t img = blockIdx.y;
  const float f = fminf(1.0f, eps / fmaxf(sqrtf(dnorm2[img]), 1e-12f));
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < chw4; i += (size_t)gridDim.x * kThreads) {
    const size_t k = img * chw4 + i;
    float4 xv = x[k], ov = ld_stream_f4(x0 + k);
    xv.x = clampf(__fadd_rn(ov.x, __fmul_rn(__fsub_rn(xv.x, ov.x), f)), 0.f, 1.f);
    xv.y = clampf(__fadd_rn(ov.y, __fmul_rn(__fsub_rn(xv.y, ov.y), f)), 0.f, 1.f);
    xv.z = clampf(__fadd_rn(ov.z, __fmul_rn(__fsub_rn(xv.z, ov.z), f)), 0.f, 1.f);
    xv.w = clampf(__fadd_rn(ov.w, __fmul_rn(__fsub_rn(xv.w, ov.w), f)), 0.f, 1.f);
    x[k] = xv;
  }
}

#ifdef __CUDACC__
// ---- L2 phases 2 + 3 in one kernel: a cluster of 8 CTAs owns one image -------------------------------------------------------
// The ascent needs ||x' - x0|| of the WHOLE image before the projection can scale any element, which is why the three-kernel form
// above touches x twice more (8 passes over an image per step against the algorithmic 5).  Here the image's 150 528 elements live
// in the registers of 8 x 512 threads (10 float4 of d = x' - x0 each, two CTAs per SM; x0 is read again from L2 at the end), the squared norm is reduced through distributed
// shared memory in fixed order (block tree, then the 8 CTA partials in rank order: deterministic, no atomics), and x is written once.
// The norm of g is the same reduction one phase earlier; g's second read and x0's second read are L2 hits, so HBM sees 4 passes
// (x, g, x0 in, x out) in ONE launch.
namespace cg = cooperative_groups;
constexpr int kClusterCtas = 8, kClusterThreads = 512, kClusterVec = 10, kClusterCtasPerSm = 2;
// sh: one 33-float buffer per call of a kernel (a CTA may write its next partial while a slow peer still reads the previous one);
// the caller ends with cluster_exit_sync() so that no CTA leaves while a peer can still read its shared memory
__device__ __forceinline__ float cluster_sum_ordered(float v, float* sh /* [33] */) {
  cg::cluster_group cluster = cg::this_cluster();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const float t = warp_sum(lane < kClusterThreads / 32 ? sh[lane] : 0.f);
    if (lane == 0) sh[32] = t;
  }
  cluster.sync();                                            // every CTA's partial is in its sh[32]
  float tot = 0.f;
#pragma unroll
  for (int r = 0; r < kClusterCtas; ++r) tot += *cluster.map_shared_rank(sh + 32, r);
  return tot;
}
__device__ __forceinline__ void cluster_exit_sync() { cg::this_cluster().sync(); }
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClusterThreads, kClusterCtasPerSm)
    l2_step_cluster_kernel(float4* __restrict__ x, const float4* __restrict__ g, const float4* __restrict__ x0, size_t chw4, float alpha,
                           float eps) {
  __shared__ float sh[2][33];
  const size_t img = blockIdx.y;
  // phase 1: ||g||^2 (g is read again below: an L2 hit, the cluster's image is 602 KB)
  float s0 = 0.f;
#pragma unroll
  for (int k = 0; k < kClusterVec; ++k) {
    const size_t i = ((size_t)k * kClusterCtas + blockIdx.x) * kClusterThreads + threadIdx.x;
    if (i < chw4) {
      const float4 v = __ldg(g + img * chw4 + i);
      s0 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  }
  const float f = alpha * (1.0f / fmaxf(sqrtf(cluster_sum_ordered(s0, sh[0])), 1e-12f));
  float4 d[kClusterVec];                                    // x0 is read again in the last phase: an L2 hit (a cluster's image is 602 KB)
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kClusterVec; ++k) {
    const size_t i = ((size_t)k * kClusterCtas + blockIdx.x) * kClusterThreads + threadIdx.x;
    if (i < chw4) {
      const size_t at = img * chw4 + i;
      float4 xv = x[at];
      const float4 gv = ld_stream_f4(g + at);
      const float4 ov = __ldg(x0 + at);
      xv.x = __fadd_rn(xv.x, __fmul_rn(f, gv.x)); xv.y = __fadd_rn(xv.y, __fmul_rn(f, gv.y));
      xv.z = __fadd_rn(xv.z, __fmul_rn(f, gv.z)); xv.w = __fadd_rn(xv.w, __fmul_rn(f, gv.w));
      d[k] = make_float4(__fsub_rn(xv.x, ov.x), __fsub_rn(xv.y, ov.y), __fsub_rn(xv.z, ov.z), __fsub_rn(xv.w, ov.w));
      s += d[k].x * d[k].x + d[k].y * d[k].y + d[k].z * d[k].z + d[k].w * d[k].w;
    }
  }
  const float tot = cluster_sum_ordered(s, sh[1]);
  const float f2 = fminf(1.0f, eps / fmaxf(sqrtf(tot), 1e-12f));
#pragma unroll
  for (int k = 0; k < kClusterVec; ++k) {
    const size_t i = ((size_t)k * kClusterCtas + blockIdx.x) * kClusterThreads + threadIdx.x;
    if (i < chw4) {
      const float4 ov = ld_stream_f4(x0 + img * chw4 + i);
      float4 xv;
      xv.x = clampf(__fadd_rn(ov.x, __fmul_rn(d[k].x, f2)), 0.f, 1.f);
      xv.y = clampf(__fadd_rn(ov.y, __fmul_rn(d[k].y, f2)), 0.f, 1.f);
      xv.z = clampf(__fadd_rn(ov.z, __fmul_rn(d[k].z, f2)), 0.f, 1.f);
      xv.w = clampf(__fadd_rn(ov.w, __fmul_rn(d[k].w, f2)), 0.f, 1.f);
      x[img * chw4 + i] = xv;
    }
  }
  cluster_exit_sync();
}
// MI-FGSM the same way: g stays in registers across the mean|g| reduction, so the step is 6 passes (g, x, m, x0 in; x, m out)
// instead of 7, in one launch, with a deterministic reduction.
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClusterThreads, kClusterCtasPerSm)
    mim_step_cluster_kernel(float4* __restrict__ x, float4* __restrict__ mom, const float4* __restrict__ g, const float4* __restrict__ x0,
                            size_t chw4, float step, float eps, float decay) {
  __shared__ float sh[33];
  const size_t img = blockIdx.y;
  float4 gv[kClusterVec];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kClusterVec; ++k) {
    const size_t i = ((size_t)k * kClusterCtas + blockIdx.x) * kClusterThreads + threadIdx.x;
    if (i < chw4) {
      gv[k] = ld_stream_f4(g + img * chw4 + i);
      s += fabsf(gv[k].x) + fabsf(gv[k].y) + fabsf(gv[k].z) + fabsf(gv[k].w);
    }
  }
  const float mean_abs = cluster_sum_ordered(s, sh) / (float)(chw4 * 4);
#pragma unroll
  for (int k = 0; k < kClusterVec; ++k) {
    const size_t i = ((size_t)k * kClusterCtas + blockIdx.x) * kClusterThreads + threadIdx.x;
    if (i < chw4) {
      const size_t at = img * chw4 + i;
      float4 xv = x[at], mv = mom[at];
      const float4 ov = ld_stream_f4(x0 + at);
      mv.x = __fadd_rn(__fmul_rn(decay, mv.x), gv[k].x / mean_abs);
      mv.y = __fadd_rn(__fmul_rn(decay, mv.y), gv[k].y / mean_abs);
      mv.z = __fadd_rn(__fmul_rn(decay, mv.z), gv[k].z / mean_abs);
      mv.w = __fadd_rn(__fmul_rn(decay, mv.w), gv[k].w / mean_abs);
      xv.x = linf_update(xv.x, step * sgn(mv.x), ov.x, eps);
      xv.y = linf_update(xv.y, step * sgn(mv.y), ov.y, eps);
      xv.z = linf_update(xv.z, step * sgn(mv.z), ov.z, eps);
      xv.w = linf_update(xv.w, step * sgn(mv.w), ov.w, eps);
      mom[at] = mv;
      x[at] = xv;
    }
  }
  cluster_exit_sync();
}
#endif  // __CUDACC__

// MIM phase 2: m = decay*m + g/mean|g| ; x = linf_update(x, step*sign(m), x0, eps)
__global__ void __launch_bounds__(kThreads) mim_apply_kernel(float4* __restrict__ x,
                                                              float4* __restrict__ mom,
                                                              const float4* __restrict__ g,
                                                              const float4* __restrict__ x0,
                                                              const float* __restrict__ abs_sum,
                                                              size_t chw4, float step, float eps,
                                                              float decay) {
  const size_t img = blockIdx.y;
  const float mean_abs = abs_sum[img] / (float)(chw4 * 4);
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < chw4; i += (size_t)gridDim.x * kThreads) {
    const size_t k = img * chw4 + i;
    float4 xv = x[k], mv = mom[k], gv = ld_stream_f4(g + k), ov = ld_stream_f4(x0 + k);
    mv.x = __fadd_rn(__fmul_rn(decay, mv.x), gv.x / mean_abs);
    mv.y = __fadd_rn(__fmul_rn(decay, mv.y), gv.y / mean_abs);
    mv.z = __fadd_rn(__fmul_rn(decay, mv.z), gv.z / mean_abs);
    mv.w = __fadd_rn(__fmul_rn(decay, mv.w), gv.w / mean_abs);
    xv.x = linf_update(xv.x, step * sgn(mv.x), ov.x, eps);
    xv.y = linf_update(xv.y, step * sgn(mv.y), ov.y, eps);
    xv.z = linf_update(xv.z, step * sgn(mv.z), ov.z, eps);
    xv.w = linf_update(xv.w, step * sgn(mv.w), ov.w, eps);
    mom[k] = mv;
    x[k] = xv;
  }
}

Norm3 make_norm(const float* mean, const float* std) {
  Norm3 n;
  for (int i = 0; i < 3; ++i) { n.mean[i] = mean[i]; n.std[i] = std[i]; n.inv_std[i] = 1.f / std[i]; }
  return n;
}

inline unsigned slices_for(size_t chw4, size_t n) {
  // enough CTAs to fill 148 SMs x 8 resident CTAs even for small n, capped by the work
  size_t per = (chw4 + kThreads - 1) / kThreads;
  size_t want = (size_t)b200r_num_sms() * 8 / (n ? n : 1) + 1;
  return (unsigned)(per < want ? per : want);
}
}  // namespace

extern "C" {

int b200r_u8nhwc_to_f32nchw(const uint8_t* in, float* out, int n, int h, int w, const float* mean_host,
                            const float* std_host, b200r_stream_t stream) {
  B200R_CHECK_ARG(in && out && mean_host && std_host, "null pointer");
  B200R_CHECK_ARG(n >= 0 && h > 0 && w > 0 && (h * w) % 16 == 0, "h*w must be a multiple of 16");
  if (n == 0) return B200R_OK;
  const uint32_t g48 = (uint32_t)(h * w) / 16;
  dim3 grid((g48 + kCvtThreads - 1) / kCvtThreads, n);
  u8nhwc_to_f32nchw_kernel<<<grid, kCvtThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(in), out, g48, (uint32_t)(h * w), make_norm(mean_host, std_host));
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_normalize_f32nchw(const float* in, float* out, int n, int h, int w, const float* mean_host,
                            const float* std_host, int inverse, b200r_stream_t stream) {
  B200R_CHECK_ARG(in && out && mean_host && std_host, "null pointer");
  B200R_CHECK_ARG(n >= 0 && h > 0 && w > 0 && (h * w) % 4 == 0, "h*w must be a multiple of 4");
  if (n == 0) return B200R_OK;
  const uint32_t hw4 = (uint32_t)(h * w) / 4;
  dim3 grid(min((hw4 + kThreads - 1) / kThreads, 64u), n * 3);
  normalize_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(in),
                                                             reinterpret_cast<float4*>(out), hw4,
                                                             make_norm(mean_host, std_host), inverse);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_random_start_linf(const float* x0, float* x, size_t n, size_t chw, float eps, uint64_t seed,
                            uint64_t image_offset, const float* u, int clip01, b200r_stream_t stream) {
  B200R_CHECK_ARG(x0 && x, "null pointer");
  B200R_CHECK_ARG(chw % 4 == 0 && n < 65536, "chw must be a multiple of 4 and n < 65536");
  if (n == 0) return B200R_OK;
  dim3 grid(slices_for(chw / 4, n), (unsigned)n);
  random_start_linf_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(x0), reinterpret_cast<float4*>(x), reinterpret_cast<const float4*>(u),
      chw / 4, eps, (uint32_t)seed, (uint32_t)(seed >> 32), image_offset, clip01);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_pgd_step_linf(float* x, const float* g, const float* x0, size_t n, size_t chw, float alpha,
                        float eps, b200r_stream_t stream) {
  B200R_CHECK_ARG(x && g && x0, "null pointer");
  B200R_CHECK_ARG((n * chw) % 4 == 0, "n*chw must be a multiple of 4");
  if (n == 0) return B200R_OK;
  const size_t total4 = n * chw / 4;
  size_t blocks = (total4 + kThreads - 1) / kThreads;
  size_t cap = (size_t)b200r_num_sms() * 16;
  pgd_step_linf_kernel<<<(unsigned)(blocks < cap ? blocks : cap), kThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<float4*>(x), reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(x0),
      total4, alpha, eps);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_pgd_step_l2(float* x, const float* g, const float* x0, size_t n, size_t chw, float alpha,
                      float eps, float* workspace, b200r_stream_t stream) {
  B200R_CHECK_ARG(x && g && x0 && workspace, "null pointer");
  B200R_CHECK_ARG(chw % 4 == 0 && n < 65536, "chw must be a multiple of 4 and n < 65536");
  if (n == 0) return B200R_OK;
  cudaStream_t s = as_stream(stream);
#ifdef __CUDACC__
  const bool three_kernels = getenv("B200R_L2_STEP") && !strcmp(getenv("B200R_L2_STEP"), "3k");
  if (chw / 4 <= (size_t)kClusterCtas * kClusterThreads * kClusterVec && !three_kernels) {   // the image fits one cluster's registers
    l2_step_cluster_kernel<<<dim3(kClusterCtas, (unsigned)n), kClusterThreads, 0, s>>>(
        reinterpret_cast<float4*>(x), reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(x0), chw / 4, alpha, eps);
    B200R_LAUNCH_CHECK();
    return B200R_OK;
  }
#endif
  B200R_CUDA(cudaMemsetAsync(workspace, 0, 2 * n * sizeof(float), s));
  dim3 grid(slices_for(chw / 4, n), (unsigned)n);
  sample_reduce_kernel<0><<<grid, kThreads, 0, s>>>(reinterpret_cast<const float4*>(g), workspace, chw / 4);
  l2_ascent_kernel<<<grid, kThreads, 0, s>>>(reinterpret_cast<float4*>(x), reinterpret_cast<const float4*>(g),
                                             reinterpret_cast<const float4*>(x0), workspace, workspace + n,
                                             chw / 4, alpha);
  l2_project_kernel<<<grid, kThreads, 0, s>>>(reinterpret_cast<float4*>(x), reinterpret_cast<const float4*>(x0),
                                              workspace + n, chw / 4, eps);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_mim_step_linf(float* x, float* momentum, const float* g, const float* x0, size_t n, size_t chw,
                        float step, float eps, float decay, float* workspace, b200r_stream_t stream) {
  B200R_CHECK_ARG(x && momentum && g && x0 && workspace, "null pointer");
  B200R_CHECK_ARG(chw % 4 == 0 && n < 65536, "chw must be a multiple of 4 and n < 65536");
  if (n == 0) return B200R_OK;
  cudaStream_t s = as_stream(stream);
#ifdef __CUDACC__
  const bool two_kernels = getenv("B200R_MIM_STEP") && !strcmp(getenv("B200R_MIM_STEP"), "2k");
  if (chw / 4 <= (size_t)kClusterCtas * kClusterThreads * kClusterVec && !two_kernels) {
    mim_step_cluster_kernel<<<dim3(kClusterCtas, (unsigned)n), kClusterThreads, 0, s>>>(
        reinterpret_cast<float4*>(x), reinterpret_cast<float4*>(momentum), reinterpret_cast<const float4*>(g),
        reinterpret_cast<const float4*>(x0), chw / 4, step, eps, decay);
    B200R_LAUNCH_CHECK();
    return B200R_OK;
  }
#endif
  B200R_CUDA(cudaMemsetAsync(workspace, 0, n * sizeof(float), s));
  dim3 grid(slices_for(chw / 4, n), (unsigned)n);
  sample_reduce_kernel<1><<<grid, kThreads, 0, s>>>(reinterpret_cast<const float4*>(g), workspace, chw / 4);
  mim_apply_kernel<<<grid, kThreads, 0, s>>>(reinterpret_cast<float4*>(x), reinterpret_cast<float4*>(momentum),
                                             reinterpret_cast<const float4*>(g),
                                             reinterpret_cast<const float4*>(x0), workspace, chw / 4, step, eps,
                                             decay);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

}  // extern "C"

// =============================================================================================
// AutoAttack device pieces (vendored fra31/auto-attack in the reference:
// RobustART/noise/utils/adv/Attacks/autoattack/autopgd_base.py:332-338, square.py:246-262)
// =============================================================================================
namespace {
__device__ __forceinline__ float apgd_one(float xa, float xo, float g, float x0, float step, float eps, float a) {
  // x_adv_1 = clamp(min(max(x_adv + step*sign(g), x-eps), x+eps), 0, 1)
  float x1 = __fadd_rn(xa, __fmul_rn(step, sgn(g)));
  x1 = clampf(fminf(fmaxf(x1, __fsub_rn(x0, eps)), __fadd_rn(x0, eps)), 0.f, 1.f);
  // x_adv_1 = clamp(min(max(x_adv + (x_adv_1 - x_adv)*a + grad2*(1-a), x-eps), x+eps), 0, 1), grad2 = x_adv - x_adv_old
  float m = __fadd_rn(__fadd_rn(xa, __fmul_rn(__fsub_rn(x1, xa), a)), __fmul_rn(__fsub_rn(xa, xo), __fsub_rn(1.f, a)));
  return clampf(fminf(fmaxf(m, __fsub_rn(x0, eps)), __fadd_rn(x0, eps)), 0.f, 1.f);
}

// x_adv <- APGD update ; x_old <- previous x_adv.  step: per-sample step size [n]
__global__ void __launch_bounds__(kThreads) apgd_step_linf_kernel(float4* __restrict__ xadv, float4* __restrict__ xold,
                                                                   const float4* __restrict__ g, const float4* __restrict__ x0,
                                                                   const float* __restrict__ step, size_t chw4, float eps, float a) {
  const size_t img = blockIdx.y;
  const float st = step[img];
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < chw4; i += (size_t)gridDim.x * kThreads) {
    const size_t k = img * chw4 + i;
    float4 xa = xadv[k], xo = xold[k], gv = ld_stream_f4(g + k), ov = ld_stream_f4(x0 + k), r;
    r.x = apgd_one(xa.x, xo.x, gv.x, ov.x, st, eps, a); r.y = apgd_one(xa.y, xo.y, gv.y, ov.y, st, eps, a);
    r.z = apgd_one(xa.z, xo.z, gv.z, ov.z, st, eps, a); r.w = apgd_one(xa.w, xo.w, gv.w, ov.w, st, eps, a);
    xold[k] = xa;
    xadv[k] = r;
  }
}

// Square attack proposal: out = clamp(min(max(x_best + 2*eps*sign_c on the window, x0-eps), x0+eps), 0, 1)
__global__ void __launch_bounds__(kThreads) square_propose_kernel(const float* __restrict__ xbest, const float* __restrict__ x0,
                                                                   float* __restrict__ out, int c, int h, int w, int vh, int vw, int s,
                                                                   float s0, float s1, float s2, float eps) {
  const size_t img = blockIdx.y;
  const int chw = c * h * w;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < chw; i += gridDim.x * kThreads) {
    const int ch = i / (h * w), rem = i - ch * h * w, y = rem / w, x = rem - y * w;
    const size_t k = img * chw + i;
    float v = xbest[k];
    if (y >= vh && y < vh + s && x >= vw && x < vw + s) v = __fadd_rn(v, __fmul_rn(2.f * eps, ch == 0 ? s0 : (ch == 1 ? s1 : s2)));
    const float o = x0[k];
    out[k] = clampf(fminf(fmaxf(v, __fsub_rn(o, eps)), __fadd_rn(o, eps)), 0.f, 1.f);
  }
}

// dst[i] = mask[i] ? src[i] : dst[i]  (row = sample)
__global__ void __launch_bounds__(kThreads) masked_rows_kernel(float4* __restrict__ dst, const float4* __restrict__ src,
                                                                const uint8_t* __restrict__ mask, size_t chw4) {
  const size_t img = blockIdx.y;
  if (!mask[img]) return;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < chw4; i += (size_t)gridDim.x * kThreads) dst[img * chw4 + i] = src[img * chw4 + i];
}
}  // namespace

extern "C" {

int b200r_apgd_step_linf(float* x_adv, float* x_adv_old, const float* g, const float* x0, const float* step, size_t n, size_t chw,
                         float eps, float a, b200r_stream_t stream) {
  B200R_CHECK_ARG(x_adv && x_adv_old && g && x0 && step, "null pointer");
  B200R_CHECK_ARG(chw % 4 == 0 && n < 65536, "chw must be a multiple of 4 and n < 65536");
  if (n == 0) return B200R_OK;
  dim3 grid(slices_for(chw / 4, n), (unsigned)n);
  apgd_step_linf_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(x_adv), reinterpret_cast<float4*>(x_adv_old),
                                                                  reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(x0),
                                                                  step, chw / 4, eps, a);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_square_propose_linf(const float* x_best, const float* x0, float* out, int n, int c, int h, int w, int vh, int vw, int s,
                              const float* signs_host, float eps, b200r_stream_t stream) {
  B200R_CHECK_ARG(x_best && x0 && out && signs_host, "null pointer");
  B200R_CHECK_ARG(c == 3 && n >= 0 && n < 65536 && s >= 1 && vh >= 0 && vw >= 0 && vh + s <= h && vw + s <= w, "bad square window");
  if (n == 0) return B200R_OK;
  dim3 grid((c * h * w + kThreads * 4 - 1) / (kThreads * 4), n);
  square_propose_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(x_best, x0, out, c, h, w, vh, vw, s, signs_host[0], signs_host[1],
                                                                  signs_host[2], eps);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_masked_rows_copy(float* dst, const float* src, const uint8_t* mask, size_t n, size_t chw, b200r_stream_t stream) {
  B200R_CHECK_ARG(dst && src && mask, "null pointer");
  B200R_CHECK_ARG(chw % 4 == 0 && n < 65536, "chw must be a multiple of 4 and n < 65536");
  if (n == 0) return B200R_OK;
  dim3 grid(slices_for(chw / 4, n), (unsigned)n);
  masked_rows_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(dst), reinterpret_cast<const float4*>(src), mask, chw / 4);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

}  // extern "C"

// ---- L2 random start (foolbox L2ProjectedGradientDescentAttack.get_random_start -> uniform_l2_n_balls) -------------------------
// A uniform point of the unit n-ball = the first n coordinates of a uniform point on the (n+1)-sphere = n + 1 normals over their
// norm.  One CTA per sample; the normals are a pure function of (seed, sample, position), so pass 2 regenerates what pass 1 summed.
namespace {

__device__ __forceinline__ void normals4(uint64_t seed, uint64_t sample, uint32_t group, float z[4]) {
  const uint4 r = philox4x32_10(make_uint4(group, 0x4C32u, (uint32_t)sample, (uint32_t)(sample >> 32)), (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float u1 = ((float)(w[2 * h] >> 8) + 1.0f) * (1.0f / 16777216.0f);     // (0, 1]
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincosf((float)w[2 * h + 1] * (6.283185307179586f / 4294967296.0f), &sn, &cs);
    z[2 * h] = rad * cs;
    z[2 * h + 1] = rad * sn;
  }
}

__global__ void __launch_bounds__(1024) random_start_l2_kernel(const float4* __restrict__ x0, float4* __restrict__ x, uint32_t chw4, float eps,
                                                               uint64_t seed, uint64_t sample0) {
  __shared__ double red[1024];
  __shared__ float s_scale;
  const uint64_t sample = sample0 + blockIdx.x;
  double acc = 0.0;
  for (uint32_t g = threadIdx.x; g <= chw4; g += 1024) {      // group chw4 holds the (n+1)-th normal
    float z[4];
    normals4(seed, sample, g, z);
    if (g < chw4) acc += (double)z[0] * z[0] + (double)z[1] * z[1] + (double)z[2] * z[2] + (double)z[3] * z[3];
    else acc += (double)z[0] * z[0];
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {            // fixed-order tree: the start is reproducible bit for bit
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) s_scale = (float)((double)eps / sqrt(red[0]));
  __syncthreads();
  const float sc = s_scale;
  const float4* x0r = x0 + (size_t)blockIdx.x * chw4;
  float4* xr = x + (size_t)blockIdx.x * chw4;
  for (uint32_t g = threadIdx.x; g < chw4; g += 1024) {
    float z[4];
    normals4(seed, sample, g, z);
    const float4 a = x0r[g];
    float4 o;
    o.x = fminf(fmaxf(fmaf(sc, z[0], a.x), 0.f), 1.f);
    o.y = fminf(fmaxf(fmaf(sc, z[1], a.y), 0.f), 1.f);
    o.z = fminf(fmaxf(fmaf(sc, z[2], a.z), 0.f), 1.f);
    o.w = fminf(fmaxf(fmaf(sc, z[3], a.w), 0.f), 1.f);
    xr[g] = o;
  }
}

}  // namespace

extern "C" int b200r_random_start_l2(const float* x0, float* x, size_t n, size_t chw, float eps, uint64_t seed, uint64_t image_offset,
                                     b200r_stream_t stream) {
  B200R_CHECK_ARG(x0 && x, "null pointer");
  B200R_CHECK_ARG(chw % 4 == 0 && chw / 4 < 0xFFFFFFFFull && n < (1u << 30), "chw must be a multiple of 4");
  B200R_CHECK_ARG(eps >= 0.f, "eps must be non-negative");
  if (n == 0) return B200R_OK;
  random_start_l2_kernel<<<(unsigned)n, 1024, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x0), reinterpret_cast<float4*>(x),
                                                                      (uint32_t)(chw / 4), eps, seed, image_offset);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}
