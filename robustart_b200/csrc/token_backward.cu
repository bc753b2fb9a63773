// Input-gradient pass of the token models (ViT-B/16, MLP-Mixer-B/16) around the tensor-core dgrad GEMMs, all on
// split-bf16 planes [rows, C] like token_layers.cu.  Only INPUT gradients are formed (what autograd.grad(loss, x) returns to
// the attacks: autopgd_base.py:371-376, imfgsm_attack.py:77-80, foolbox value_and_grad) -- no weight gradients.
//   layernorm_bwd      nn.LayerNorm backward w.r.t. its input, residual gradient added in the same pass
//   act / act_bwd      GELU (tanh form, vision_transformer.py:19-37; erf form, nn.GELU) and tanh on a saved pre-activation
//   attention_bwd      d(softmax(q k^T scale) v) w.r.t. packed qkv, probabilities recomputed (nothing saved but qkv)
//   patch_scatter      transpose of the patch gather + Normalize: dcols [n*np, 3*p*p] -> float32 NCHW image gradient
// STATUS: first version, CUDA cores only; the attention backward is the piece to move onto tcgen05.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

size_t b200r_attention_bwd_tc_ws(int n, int tokens, int heads);                                                               // attention_bwd_sm100.cu
int b200r_attention_bwd_tc(const uint16_t* qkv, const uint16_t* dout, uint16_t* dqkv, float* stats, int n, int tokens, int heads, float scale,
                           cudaStream_t stream);

namespace {
constexpr int kThreads = 256;

__device__ __forceinline__ float pl_get(const uint16_t* hi, const uint16_t* lo, size_t i) {
  return plane_bits_to_f32(hi[i]) + plane_bits_to_f32(lo[i]);
}
__device__ __forceinline__ void pl_put(uint16_t* hi, uint16_t* lo, size_t i, float v) {
  uint16_t h, l;
  split_pair(v, h, l);
  hi[i] = h; lo[i] = l;
}
__device__ __forceinline__ void unpack8(uint4 h, uint4 l, float* v) {
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = plane_lo16_f32(hw[j]) + plane_lo16_f32(lw[j]);
    v[2 * j + 1] = plane_hi16_f32(hw[j]) + plane_hi16_f32(lw[j]);
  }
}
__device__ __forceinline__ void pack8(const float* v, uint4& h, uint4& l) {
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_pair2(v[2 * j], v[2 * j + 1], hw[j], lw[j]);
  h = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  l = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// ---- LayerNorm backward: one warp per row, C <= 1024 ------------------------------------------------
//   xhat = (x - mean) rstd,  gy = dy gamma,  dx = rstd (gy - mean(gy) - xhat mean(gy xhat))  [+ add]
// NV = 16-byte vectors per lane (ceil(C / 256)): ViT / Mixer rows (C = 768) keep 2 x 24 values in registers instead of 2 x 32, and the
// residual-gradient operand is requested together with the other two -- three CTAs per SM and no second exposed DRAM latency
// (ncu before: 84 registers -> 2 CTAs per SM, 22 % of the warp slots, top stall long_scoreboard, 44 % of the DRAM peak).
constexpr int kLnMaxVec = 4;
template <int NV>
__global__ void __launch_bounds__(kThreads, NV <= 3 ? 3 : 2) layernorm_bwd_kernel(const uint4* __restrict__ dyh, const uint4* __restrict__ dyl,
                                                                  const uint4* __restrict__ xh, const uint4* __restrict__ xl,
                                                                  const uint4* __restrict__ ah, const uint4* __restrict__ al,
                                                                  uint4* __restrict__ oh, uint4* __restrict__ ol,
                                                                  const float* __restrict__ gamma, int rows, int c8, float eps) {
  const int row = (blockIdx.x * kThreads + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[NV][8], g[NV][8];
  uint4 rah[NV], ral[NV];                                   // residual gradient, raw planes (in flight during the reductions)
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    if (ah && i < c8) { rah[k] = __ldg(ah + (size_t)row * c8 + i); ral[k] = __ldg(al + (size_t)row * c8 + i); }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    if (i < c8) {
      unpack8(__ldg(xh + (size_t)row * c8 + i), __ldg(xl + (size_t)row * c8 + i), v[k]);
      unpack8(__ldg(dyh + (size_t)row * c8 + i), __ldg(dyl + (size_t)row * c8 + i), g[k]);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * i), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * i + 1);
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) g[k][j] *= gm[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[k][j] = 0.f; g[k][j] = 0.f; }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[k][j];
  }
  const float inv_c = 1.f / (float)(c8 * 8);
  const float mean = warp_sum(s) * inv_c;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k)
    if (lane + 32 * k < c8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[k][j] - mean; q += d * d; }
    }
  const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k)
    if (lane + 32 * k < c8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[k][j] = (v[k][j] - mean) * rstd;      // xhat
        s1 += g[k][j];
        s2 += g[k][j] * v[k][j];
      }
    }
  s1 = warp_sum(s1) * inv_c;
  s2 = warp_sum(s2) * inv_c;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    if (i >= c8) continue;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = rstd * (g[k][j] - s1 - v[k][j] * s2);
    if (ah) {
      float a[8];
      unpack8(rah[k], ral[k], a);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += a[j];
    }
    uint4 h, l;
    pack8(o, h, l);
    oh[(size_t)row * c8 + i] = h;
    ol[(size_t)row * c8 + i] = l;
  }
}

// ---- activations on a saved pre-activation ----------------------------------------------------------
__device__ __forceinline__ float act_fwd(float v, int act) {
  switch (act) {
    case B200R_ACT_GELU_TANH: {
      const float u = 0.7978845608028654f * (v + 0.044715f * v * v * v);
      return 0.5f * v * (1.f + tanhf(u));
    }
    case B200R_ACT_GELU_ERF: return 0.5f * v * (1.f + erff(v * 0.7071067811865476f));
    case B200R_ACT_TANH: return tanhf(v);
    case B200R_ACT_RELU: return fmaxf(v, 0.f);
    case B200R_ACT_RELU6: return fminf(fmaxf(v, 0.f), 6.f);
    case B200R_ACT_SWISH: return v / (1.f + expf(-v));
    case B200R_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}
__device__ __forceinline__ float act_deriv(float v, int act) {
  switch (act) {
    case B200R_ACT_GELU_TANH: {
      const float u = 0.7978845608028654f * (v + 0.044715f * v * v * v);
      const float t = tanhf(u);
      return 0.5f * (1.f + t) + 0.5f * v * (1.f - t * t) * 0.7978845608028654f * (1.f + 3.f * 0.044715f * v * v);
    }
    case B200R_ACT_GELU_ERF:
      return 0.5f * (1.f + erff(v * 0.7071067811865476f)) + v * 0.3989422804014327f * expf(-0.5f * v * v);
    case B200R_ACT_TANH: {
      const float t = tanhf(v);
      return 1.f - t * t;
    }
    // ReLU / ReLU6: the derivative is the same function of the pre-activation and of the OUTPUT (0 < v [< 6]), so the saved output
    // of a fused conv + BN + ReLU6 layer serves as `pre` (mobilenet_v2.py:31-47: ReLU6 = hardtanh(0, 6))
    case B200R_ACT_RELU: return v > 0.f ? 1.f : 0.f;
    case B200R_ACT_RELU6: return (v > 0.f && v < 6.f) ? 1.f : 0.f;
    case B200R_ACT_SWISH: {                                  // x sigmoid(x) (efficientnet.py:271-277)
      const float sg = 1.f / (1.f + expf(-v));
      return sg * (1.f + v * (1.f - sg));
    }
    case B200R_ACT_SIGMOID: {
      const float sg = 1.f / (1.f + expf(-v));
      return sg * (1.f - sg);
    }
    default: return 1.f;
  }
}
// The gradient pass's form of the derivatives: one ex2 and one rcp on the MUFU per element (the formulas of the GEMM epilogue's
// activations, gemm_sm100.cu), absolute error ~2e-7 -- the libm forms above cost 40-60 instructions per element and made the
// pass issue-bound (ncu: 75 % issue, 43 % DRAM).  ACT is a template parameter: no switch inside the element loop.
template <int ACT>
__device__ __forceinline__ float act_deriv_fast(float v) {
  if constexpr (ACT == B200R_ACT_GELU_TANH) {
    // gelu = v s, s = 0.5 (1 + tanh u) = 1 / (1 + exp(-2u)), u = k (v + c v^3):  d = s + 2 v s (1 - s) k (1 + 3 c v^2)
    const float v2 = v * v;
    const float u = 0.7978845608028654f * fmaf(0.044715f * v2, v, v);
    const float sg = __fdividef(1.f, 1.f + __expf(-2.f * u));
    return fmaf(2.f * v * sg * (1.f - sg), 0.7978845608028654f * fmaf(3.f * 0.044715f, v2, 1.f), sg);
  } else if constexpr (ACT == B200R_ACT_GELU_ERF) {
    // d = Phi(v) + v phi(v); erf(v / sqrt 2) by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7), sharing exp(-v^2 / 2) with phi
    const float ax = fabsf(v) * 0.7071067811865476f, t = __fdividef(1.f, fmaf(0.3275911f, ax, 1.f));
    const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
    const float e = __expf(-ax * ax);
    const float erfv = copysignf(1.f - poly * e, v);
    return fmaf(v * 0.3989422804014327f, e, 0.5f * (1.f + erfv));
  } else if constexpr (ACT == B200R_ACT_TANH) {
    const float t = 1.f - __fdividef(2.f, 1.f + __expf(2.f * v));
    return 1.f - t * t;
  } else if constexpr (ACT == B200R_ACT_SWISH) {
    const float sg = __fdividef(1.f, 1.f + __expf(-v));
    return sg * fmaf(v, 1.f - sg, 1.f);
  } else if constexpr (ACT == B200R_ACT_SIGMOID) {
    const float sg = __fdividef(1.f, 1.f + __expf(-v));
    return sg * (1.f - sg);
  } else {
    return act_deriv(v, ACT);                                // ReLU / ReLU6: comparisons only
  }
}
template <bool BWD, int ACT = 0>
__global__ void __launch_bounds__(kThreads) act_planes_kernel(const uint4* __restrict__ ph, const uint4* __restrict__ pl,
                                                               const uint4* __restrict__ dyh, const uint4* __restrict__ dyl,
                                                               uint4* __restrict__ oh, uint4* __restrict__ ol, size_t count8, int act) {
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < count8; t += (size_t)gridDim.x * kThreads) {
    float p[8], o[8];
    unpack8(__ldg(ph + t), __ldg(pl + t), p);
    if (BWD) {
      float d[8];
      unpack8(__ldg(dyh + t), __ldg(dyl + t), d);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = d[j] * act_deriv_fast<ACT>(p[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = act_fwd(p[j], act);
    }
    uint4 h, l;
    pack8(o, h, l);
    oh[t] = h; ol[t] = l;
  }
}

// ---- patch scatter: transpose of patch_gather_kernel<false> --------------------------------------------
struct Std3 { float inv[3]; };
__global__ void __launch_bounds__(kThreads) patch_scatter_kernel(const uint4* __restrict__ ch, const uint4* __restrict__ cl,
                                                                  float* __restrict__ dx, int n, int h, int w, int ps, Std3 sd) {
  const int gw = w / ps, gh = h / ps, K = 3 * ps * ps, k8 = K / 8;
  const size_t total = (size_t)n * gh * gw * k8;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int chunk = (int)(t % k8);
    const size_t patch = t / k8;
    const int px = (int)(patch % gw), py = (int)((patch / gw) % gh), im = (int)(patch / ((size_t)gw * gh));
    const int col = chunk * 8;                 // ps % 8 == 0: the 8 columns share (c, ky) and are 8 consecutive pixels
    const int c = col / (ps * ps), rem = col - c * ps * ps, ky = rem / ps, kx = rem - ky * ps;
    float v[8];
    unpack8(__ldg(ch + t), __ldg(cl + t), v);
    float* dst = dx + (((size_t)im * 3 + c) * h + (py * ps + ky)) * w + px * ps + kx;
    const float s = sd.inv[c];
    *reinterpret_cast<float4*>(dst) = make_float4(v[0] * s, v[1] * s, v[2] * s, v[3] * s);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4] * s, v[5] * s, v[6] * s, v[7] * s);
  }
}

// ---- attention backward: one CTA per (image, head); q, k, v, d(out) resident in smem as fp32 (rows padded to D+1) --------
//   P = softmax(scale q k^T), out = P v
//   dV = P^T dO,  dP = dO v^T,  dS = P o (dP - rowsum(P o dP)),  dQ = scale dS k,  dK = scale dS^T q
// phase A (one warp per query row): row max, 1/sum and delta = rowsum(P o dP) kept in smem, dQ written;
// phase B (one warp per key row): probabilities rebuilt from those statistics, dK and dV written.  No atomics: the result
// does not depend on the schedule.
constexpr int kAbThreads = 256;
template <int D>
__global__ void __launch_bounds__(kAbThreads) attention_bwd_kernel(const uint16_t* __restrict__ qh, const uint16_t* __restrict__ ql,
                                                                    const uint16_t* __restrict__ gh, const uint16_t* __restrict__ gl,
                                                                    uint16_t* __restrict__ dh, uint16_t* __restrict__ dl, int T, int H,
                                                                    float scale) {
  extern __shared__ float smf[];
  constexpr int LD = D + 1;
  const int Tp = (T + 31) & ~31;
  float* sQ = smf;
  float* sK = sQ + (size_t)T * LD;
  float* sV = sK + (size_t)T * LD;
  float* sG = sV + (size_t)T * LD;
  float* sM = sG + (size_t)T * LD;          // row max of the scaled scores
  float* sL = sM + T;                        // 1 / row sum
  float* sDl = sL + T;                       // delta
  float* sA = sDl + T;                       // [warps][Tp]
  float* sB = sA + (kAbThreads / 32) * Tp;   // [warps][Tp]
  const int b = blockIdx.x / H, hd = blockIdx.x % H;
  const int C3 = 3 * H * D, C1 = H * D;
  const size_t row0 = (size_t)b * T;
  for (int i = threadIdx.x; i < T * D; i += kAbThreads) {
    const int t = i / D, d = i - t * D;
    const size_t base = (row0 + t) * C3 + hd * D + d;
    sQ[t * LD + d] = pl_get(qh, ql, base);
    sK[t * LD + d] = pl_get(qh, ql, base + (size_t)C1);
    sV[t * LD + d] = pl_get(qh, ql, base + 2 * (size_t)C1);
    sG[t * LD + d] = pl_get(gh, gl, (row0 + t) * C1 + hd * D + d);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* myA = sA + warp * Tp;
  float* myB = sB + warp * Tp;
  // ---- phase A: query rows ----
  for (int i = warp; i < T; i += kAbThreads / 32) {
    const float* qi = sQ + i * LD;
    const float* gi = sG + i * LD;
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) {
      const float* kr = sK + j * LD;
      const float* vr = sV + j * LD;
      float s = 0.f, dp = 0.f;
#pragma unroll 16
      for (int d = 0; d < D; ++d) { s = fmaf(qi[d], kr[d], s); dp = fmaf(gi[d], vr[d], dp); }
      s *= scale;
      myA[j] = s;
      myB[j] = dp;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) { const float e = expf(myA[j] - mx); myA[j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    float delta = 0.f;
    for (int j = lane; j < T; j += 32) delta = fmaf(myA[j], myB[j], delta);
    delta = warp_sum(delta) * inv;
    for (int j = lane; j < T; j += 32) myB[j] = myA[j] * inv * (myB[j] - delta);       // dS
    if (lane == 0) { sM[i] = mx; sL[i] = inv; sDl[i] = delta; }
    __syncwarp();
    float acc[D / 32];
#pragma unroll
    for (int k = 0; k < D / 32; ++k) acc[k] = 0.f;
    for (int j = 0; j < T; ++j) {
      const float ds = myB[j];
#pragma unroll
      for (int k = 0; k < D / 32; ++k) acc[k] = fmaf(ds, sK[j * LD + lane + 32 * k], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < D / 32; ++k) pl_put(dh, dl, (row0 + i) * (size_t)C3 + hd * D + lane + 32 * k, acc[k] * scale);
    __syncwarp();
  }
  __syncthreads();
  // ---- phase B: key rows ----
  for (int j = warp; j < T; j += kAbThreads / 32) {
    const float* kj = sK + j * LD;
    const float* vj = sV + j * LD;
    for (int i = lane; i < T; i += 32) {
      const float* qr = sQ + i * LD;
      const float* gr = sG + i * LD;
      float s = 0.f, dp = 0.f;
#pragma unroll 16
      for (int d = 0; d < D; ++d) { s = fmaf(qr[d], kj[d], s); dp = fmaf(gr[d], vj[d], dp); }
      const float p = expf(s * scale - sM[i]) * sL[i];
      myA[i] = p;
      myB[i] = p * (dp - sDl[i]);
    }
    __syncwarp();
    float accv[D / 32], acck[D / 32];
#pragma unroll
    for (int k = 0; k < D / 32; ++k) { accv[k] = 0.f; acck[k] = 0.f; }
    for (int i = 0; i < T; ++i) {
      const float p = myA[i], ds = myB[i];
#pragma unroll
      for (int k = 0; k < D / 32; ++k) {
        accv[k] = fmaf(p, sG[i * LD + lane + 32 * k], accv[k]);
        acck[k] = fmaf(ds, sQ[i * LD + lane + 32 * k], acck[k]);
      }
    }
    const size_t base = (row0 + j) * (size_t)C3 + hd * D;
#pragma unroll
    for (int k = 0; k < D / 32; ++k) {
      pl_put(dh, dl, base + C1 + lane + 32 * k, acck[k] * scale);
      pl_put(dh, dl, base + 2 * (size_t)C1 + lane + 32 * k, accv[k]);
    }
    __syncwarp();
  }
}

inline unsigned grid_for(size_t items) {
  size_t b = (items + kThreads - 1) / kThreads;
  size_t cap = (size_t)b200r_num_sms() * 16;
  return (unsigned)(b < cap ? (b ? b : 1) : cap);
}
inline bool act_ok(int act) { return act >= B200R_ACT_RELU && act <= B200R_ACT_SIGMOID; }
}  // namespace

extern "C" {

int b200r_layernorm_bwd(const uint16_t* dy, const uint16_t* x, const float* gamma, const uint16_t* add, uint16_t* dx, int rows, int c,
                        float eps, b200r_stream_t stream) {
  B200R_CHECK_ARG(dy && x && gamma && dx, "null pointer");
  B200R_CHECK_ARG(rows > 0 && c % 8 == 0 && c <= 32 * 8 * kLnMaxVec, "c must be a multiple of 8 and <= %d", 32 * 8 * kLnMaxVec);
  const size_t cnt = (size_t)rows * c;
#define B200R_LN_BWD(NV)                                                                                                          \
  layernorm_bwd_kernel<NV><<<(unsigned)(((size_t)rows * 32 + kThreads - 1) / kThreads), kThreads, 0, as_stream(stream)>>>(          \
      reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(dy + cnt), reinterpret_cast<const uint4*>(x),              \
      reinterpret_cast<const uint4*>(x + cnt), reinterpret_cast<const uint4*>(add), reinterpret_cast<const uint4*>(add ? add + cnt : nullptr), \
      reinterpret_cast<uint4*>(dx), reinterpret_cast<uint4*>(dx + cnt), gamma, rows, c / 8, eps)
  switch ((c / 8 + 31) / 32) {
    case 1: B200R_LN_BWD(1); break;
    case 2: B200R_LN_BWD(2); break;
    case 3: B200R_LN_BWD(3); break;
    default: B200R_LN_BWD(4); break;
  }
#undef B200R_LN_BWD
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_act_planes(const uint16_t* pre, uint16_t* out, size_t count, int act, b200r_stream_t stream) {
  B200R_CHECK_ARG(pre && out, "null pointer");
  B200R_CHECK_ARG(count > 0 && count % 8 == 0, "count must be a positive multiple of 8");
  B200R_CHECK_ARG(act_ok(act), "activation %d not supported", act);
  act_planes_kernel<false><<<grid_for(count / 8), kThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(pre), reinterpret_cast<const uint4*>(pre + count), nullptr, nullptr, reinterpret_cast<uint4*>(out),
      reinterpret_cast<uint4*>(out + count), count / 8, act);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_act_bwd_planes(const uint16_t* dy, const uint16_t* pre, uint16_t* dx, size_t count, int act, b200r_stream_t stream) {
  B200R_CHECK_ARG(dy && pre && dx, "null pointer");
  B200R_CHECK_ARG(count > 0 && count % 8 == 0, "count must be a positive multiple of 8");
  B200R_CHECK_ARG(act_ok(act), "activation %d not supported", act);
#define B200R_ACT_BWD(A)                                                                                                           \
  case A:                                                                                                                          \
    act_planes_kernel<true, A><<<grid_for(count / 8), kThreads, 0, as_stream(stream)>>>(                                           \
        reinterpret_cast<const uint4*>(pre), reinterpret_cast<const uint4*>(pre + count), reinterpret_cast<const uint4*>(dy),      \
        reinterpret_cast<const uint4*>(dy + count), reinterpret_cast<uint4*>(dx), reinterpret_cast<uint4*>(dx + count), count / 8, act); \
    break
  switch (act) {
    B200R_ACT_BWD(B200R_ACT_RELU);
    B200R_ACT_BWD(B200R_ACT_RELU6);
    B200R_ACT_BWD(B200R_ACT_GELU_TANH);
    B200R_ACT_BWD(B200R_ACT_GELU_ERF);
    B200R_ACT_BWD(B200R_ACT_SWISH);
    B200R_ACT_BWD(B200R_ACT_TANH);
    default: B200R_ACT_BWD(B200R_ACT_SIGMOID);
  }
#undef B200R_ACT_BWD
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_patch_scatter_f32(const uint16_t* dcols, float* dx, int n, int h, int w, int patch, const float* std_host, b200r_stream_t stream) {
  B200R_CHECK_ARG(dcols && dx && std_host, "null pointer");
  B200R_CHECK_ARG(n > 0 && patch > 0 && patch % 8 == 0 && h % patch == 0 && w % patch == 0, "bad patch geometry (patch must be a multiple of 8)");
  Std3 sd;
  for (int i = 0; i < 3; ++i) sd.inv[i] = 1.0f / std_host[i];
  const size_t cnt = (size_t)n * (h / patch) * (w / patch) * 3 * patch * patch;
  patch_scatter_kernel<<<grid_for(cnt / 8), kThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(dcols), reinterpret_cast<const uint4*>(dcols + cnt), dx, n, h, w, patch, sd);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_attention_bwd(const uint16_t* qkv, const uint16_t* dout, uint16_t* dqkv, int n, int tokens, int heads, int head_dim, float scale,
                        b200r_stream_t stream) {
  B200R_CHECK_ARG(qkv && dout && dqkv, "null pointer");
  B200R_CHECK_ARG(n > 0 && tokens > 0 && heads > 0, "bad shape");
  B200R_CHECK_ARG(head_dim == 64, "head_dim %d not supported (64 only)", head_dim);
  const size_t cin = (size_t)n * tokens * 3 * heads * head_dim, cout = (size_t)n * tokens * heads * head_dim;
  const int Tp = (tokens + 31) & ~31;
  const size_t smem = ((size_t)4 * tokens * (head_dim + 1) + 3 * (size_t)tokens + 2 * (kAbThreads / 32) * (size_t)Tp) * sizeof(float);
  B200R_CHECK_ARG(smem <= 227 * 1024, "sequence too long for the shared-memory attention backward kernel (%d tokens)", tokens);
  B200R_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_bwd_kernel<64><<<n * heads, kAbThreads, smem, as_stream(stream)>>>(qkv, qkv + cin, dout, dout + cout, dqkv, dqkv + cin, tokens,
                                                                              heads, scale);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

// tensor-core path (attention_bwd_sm100.cu) when the geometry allows it and a workspace for the per-query softmax statistics is given
int b200r_attention_bwd_workspace_bytes(int n, int tokens, int heads, size_t* bytes) {
  B200R_CHECK_ARG(bytes && n > 0 && tokens > 0 && heads > 0, "bad argument");
  *bytes = b200r_attention_bwd_tc_ws(n, tokens, heads);
  return B200R_OK;
}

int b200r_attention_bwd_ws(const uint16_t* qkv, const uint16_t* dout, uint16_t* dqkv, void* workspace, size_t ws_bytes, int n, int tokens,
                           int heads, int head_dim, float scale, b200r_stream_t stream) {
  B200R_CHECK_ARG(qkv && dout && dqkv, "null pointer");
  B200R_CHECK_ARG(n > 0 && tokens > 0 && heads > 0, "bad shape");
  B200R_CHECK_ARG(head_dim == 64, "head_dim %d not supported (64 only)", head_dim);
  static int force_cc = -1;
  if (force_cc < 0) { const char* e = getenv("B200R_ATTENTION_BWD"); force_cc = (e && !strcmp(e, "cuda-core")) ? 1 : 0; }
  if (!force_cc && workspace && ws_bytes >= b200r_attention_bwd_tc_ws(n, tokens, heads)) {
    const int rc = b200r_attention_bwd_tc(qkv, dout, dqkv, static_cast<float*>(workspace), n, tokens, heads, scale, as_stream(stream));
    if (rc != B200R_ENOTSUP) return rc;
  }
  return b200r_attention_bwd(qkv, dout, dqkv, n, tokens, heads, head_dim, scale, stream);
}

}  // extern "C"
