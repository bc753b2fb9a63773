// Token-model layers around the tensor-core GEMMs (ViT-B/16, MLP-Mixer-B/16), all on split-bf16 planes
// [rows, C] (planes[0:count] = hi, planes[count:2count] = lo).
//   layernorm          nn.LayerNorm (vision_transformer.py:133,137,190 eps 1e-5; mlp_mixer: vit_base.py:159 eps 1e-6)
//   patch gather       Conv2d(3, D, k=16, s=16) as a GEMM operand: [n*196, 768], column = c*256 + ky*16 + kx
//                      (vision_transformer.py:250-252,320-324), ToTensor+Normalize fused for uint8 input
//   token assemble     cat(cls, x) + pos_embedding (vision_transformer.py:326-330)
//   attention          softmax(q k^T * scale) v per (image, head), fp32 on CUDA cores (vision_transformer.py:80-92)
//   token transpose    [b, t, c] <-> [b, c, t_pad] for Mixer token mixing (mlp_mixer.py:30-40), residual fused
#include "common.cuh"
#include <stdlib.h>

int b200r_attention_tc(const uint16_t* qkv, uint16_t* out, int n, int tokens, int heads, float scale, cudaStream_t stream);   // attention_sm100.cu

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

// ---- LayerNorm: one warp per row, C <= 32*8*MAXV ------------------------------------------------
constexpr int kLnMaxVec = 4;  // up to 1024 channels
__global__ void __launch_bounds__(kThreads) layernorm_kernel(const uint4* __restrict__ xh, const uint4* __restrict__ xl,
                                                              uint4* __restrict__ yh, uint4* __restrict__ yl,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              int rows, int c8, float eps) {
  const int row = (blockIdx.x * kThreads + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  // compile-time slots (lane + 32 k): the row stays in registers (a run-time slot index put it in local memory), and the
  // loads of all slots are issued before the first is consumed
  uint4 rh[kLnMaxVec], rl[kLnMaxVec];
#pragma unroll
  for (int k = 0; k < kLnMaxVec; ++k) {
    const int i = lane + 32 * k;
    if (i < c8) { rh[k] = __ldg(xh + (size_t)row * c8 + i); rl[k] = __ldg(xl + (size_t)row * c8 + i); }
    else { rh[k] = make_uint4(0, 0, 0, 0); rl[k] = make_uint4(0, 0, 0, 0); }
  }
  float v[kLnMaxVec][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kLnMaxVec; ++k) {
    unpack8(rh[k], rl[k], v[k]);                       // absent slots unpack to zeros
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[k][j];
  }
  const float mean = warp_sum(s) / (float)(c8 * 8);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < kLnMaxVec; ++k)
    if (lane + 32 * k < c8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[k][j] - mean; q += d * d; }
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)(c8 * 8) + eps);
#pragma unroll
  for (int k = 0; k < kLnMaxVec; ++k) {
    const int i = lane + 32 * k;
    if (i >= c8) continue;
    float o[8];
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * i), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * i + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * i), b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * i + 1);
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (v[k][j] - mean) * rstd * g[j] + b[j];
    uint4 h, l;
    pack8(o, h, l);
    yh[(size_t)row * c8 + i] = h;
    yl[(size_t)row * c8 + i] = l;
  }
}

// ---- patch gather --------------------------------------------------------------------------------
struct Norm3 { float mean[3], std[3]; };
template <bool U8>
__global__ void __launch_bounds__(kThreads) patch_gather_kernel(const void* __restrict__ img, uint4* __restrict__ yh,
                                                                 uint4* __restrict__ yl, int n, int h, int w, int ps, Norm3 nm) {
  const int gw = w / ps, gh = h / ps, K = 3 * ps * ps, k8 = K / 8;
  const size_t total = (size_t)n * gh * gw * k8;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int chunk = (int)(t % k8);
    const size_t patch = t / k8;
    const int px = (int)(patch % gw), py = (int)((patch / gw) % gh), im = (int)(patch / ((size_t)gw * gh));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = chunk * 8 + j;
      const int c = col / (ps * ps), rem = col - c * ps * ps, ky = rem / ps, kx = rem - ky * ps;
      const int y = py * ps + ky, x = px * ps + kx;
      float xv;
      if (U8) xv = __fdiv_rn((float)static_cast<const uint8_t*>(img)[(((size_t)im * h + y) * w + x) * 3 + c], 255.0f);
      else xv = static_cast<const float*>(img)[(((size_t)im * 3 + c) * h + y) * w + x];
      v[j] = (xv - nm.mean[c]) / nm.std[c];
    }
    uint4 hh, ll;
    pack8(v, hh, ll);
    yh[t] = hh; yl[t] = ll;
  }
}

// ---- cat(cls, x) + pos ----------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) assemble_tokens_kernel(const uint4* __restrict__ xh, const uint4* __restrict__ xl,
                                                                    const float* __restrict__ cls, const float* __restrict__ pos,
                                                                    uint4* __restrict__ yh, uint4* __restrict__ yl, int n, int np, int c8) {
  const size_t total = (size_t)n * (np + 1) * c8;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kThreads) {
    const int cc = (int)(t % c8);
    const size_t tok = t / c8;
    const int ti = (int)(tok % (np + 1)), im = (int)(tok / (np + 1));
    float v[8];
    if (ti == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = cls[cc * 8 + j];
    } else {
      const size_t src = ((size_t)im * np + (ti - 1)) * c8 + cc;
      unpack8(__ldg(xh + src), __ldg(xl + src), v);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += pos[(size_t)ti * c8 * 8 + cc * 8 + j];
    uint4 hh, ll;
    pack8(v, hh, ll);
    yh[t] = hh; yl[t] = ll;
  }
}

// ---- attention: one CTA per (image, head); K and V resident in smem (fp32, rows padded to D+1) ------
constexpr int kAttnThreads = 256;
template <int D>
__global__ void __launch_bounds__(kAttnThreads) attention_kernel(const uint16_t* __restrict__ qh, const uint16_t* __restrict__ ql,
                                                                  uint16_t* __restrict__ oh, uint16_t* __restrict__ ol, int T, int H,
                                                                  float scale) {
  extern __shared__ float smf[];
  float* sK = smf;                       // [T][D+1]
  float* sV = sK + (size_t)T * (D + 1);  // [T][D+1]
  float* sQ = sV + (size_t)T * (D + 1);  // [warps][D]
  float* sP = sQ + (kAttnThreads / 32) * D;  // [warps][T_pad]
  const int Tp = (T + 31) & ~31;
  const int b = blockIdx.x / H, hd = blockIdx.x % H;
  const int C3 = 3 * H * D;
  const size_t row0 = (size_t)b * T;
  for (int i = threadIdx.x; i < T * D; i += kAttnThreads) {
    const int t = i / D, d = i - t * D;
    const size_t base = (row0 + t) * C3 + hd * D + d;
    sK[t * (D + 1) + d] = pl_get(qh, ql, base + (size_t)H * D);
    sV[t * (D + 1) + d] = pl_get(qh, ql, base + 2 * (size_t)H * D);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* myQ = sQ + warp * D;
  float* myP = sP + warp * Tp;
  for (int qi = warp; qi < T; qi += kAttnThreads / 32) {
    for (int d = lane; d < D; d += 32) myQ[d] = pl_get(qh, ql, (row0 + qi) * C3 + hd * D + d);
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) {
      float s = 0.f;
      const float* kr = sK + j * (D + 1);
#pragma unroll 16
      for (int d = 0; d < D; ++d) s = fmaf(myQ[d], kr[d], s);
      s *= scale;
      myP[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) { const float e = expf(myP[j] - mx); myP[j] = e; sum += e; }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.f / sum;
    // out[d] = sum_j p_j V[j][d]; lane owns d = lane and lane + 32
    float acc[D / 32];
#pragma unroll
    for (int k = 0; k < D / 32; ++k) acc[k] = 0.f;
    for (int j = 0; j < T; ++j) {
      const float p = myP[j];
#pragma unroll
      for (int k = 0; k < D / 32; ++k) acc[k] = fmaf(p, sV[j * (D + 1) + lane + 32 * k], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < D / 32; ++k) pl_put(oh, ol, (row0 + qi) * (size_t)(H * D) + hd * D + lane + 32 * k, acc[k] * inv);
    __syncwarp();
  }
}

// ---- Mixer token transposes -----------------------------------------------------------------------
// forward: y[b, c, t] (t < Tp, zero for t >= T) = x[b, t, c]
// 64 x 64 tiles, two elements (one 32-bit word) per thread and plane on both sides: 128-byte warp accesses instead of the
// 64-byte ones of a 16-bit-per-thread transpose; an element travels through shared memory as (hi | lo << 16).
// T, C, Tp even (checked by the callers).
__global__ void __launch_bounds__(kThreads) tokens_to_channels_kernel(const uint16_t* __restrict__ xh, const uint16_t* __restrict__ xl,
                                                                       uint16_t* __restrict__ yh, uint16_t* __restrict__ yl, int B, int T,
                                                                       int C, int Tp) {
  __shared__ uint32_t tile[64][65];
  const int b = blockIdx.z, t0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows per pass
  for (int r = ty; r < 64; r += 8) {
    const int t = t0 + r, c = c0 + 2 * tx;
    uint32_t h = 0, l = 0;
    if (t < T && c < C) {
      const size_t i = ((size_t)b * T + t) * C + c;
      h = *reinterpret_cast<const uint32_t*>(xh + i);
      l = *reinterpret_cast<const uint32_t*>(xl + i);
    }
    tile[r][2 * tx] = (h & 0xFFFFu) | (l << 16);
    tile[r][2 * tx + 1] = (h >> 16) | (l & 0xFFFF0000u);
  }
  __syncthreads();
  for (int r = ty; r < 64; r += 8) {
    const int c = c0 + r, t = t0 + 2 * tx;
    if (c < C && t < Tp) {
      const uint32_t a = tile[2 * tx][r], d = tile[2 * tx + 1][r];
      const size_t o = ((size_t)b * C + c) * Tp + t;
      *reinterpret_cast<uint32_t*>(yh + o) = (a & 0xFFFFu) | (d << 16);
      *reinterpret_cast<uint32_t*>(yl + o) = (a >> 16) | (d & 0xFFFF0000u);
    }
  }
}
// backward with residual: out[b, t, c] = res[b, t, c] + y[b, c, t]
__global__ void __launch_bounds__(kThreads) channels_to_tokens_add_kernel(const uint16_t* __restrict__ yh, const uint16_t* __restrict__ yl,
                                                                           const uint16_t* __restrict__ rh, const uint16_t* __restrict__ rl,
                                                                           uint16_t* __restrict__ oh, uint16_t* __restrict__ ol, int B,
                                                                           int T, int C, int Tp) {
  __shared__ float tile[64][65];
  const int b = blockIdx.z, t0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 64; r += 8) {
    const int c = c0 + r, t = t0 + 2 * tx;
    float v0 = 0.f, v1 = 0.f;
    if (c < C && t < T) {
      const size_t i = ((size_t)b * C + c) * Tp + t;
      const uint32_t h = *reinterpret_cast<const uint32_t*>(yh + i), l = *reinterpret_cast<const uint32_t*>(yl + i);
      v0 = plane_lo16_f32(h) + plane_lo16_f32(l);
      v1 = plane_hi16_f32(h) + plane_hi16_f32(l);
    }
    tile[r][2 * tx] = v0;
    tile[r][2 * tx + 1] = v1;
  }
  __syncthreads();
  for (int r = ty; r < 64; r += 8) {
    const int t = t0 + r, c = c0 + 2 * tx;
    if (t < T && c < C) {
      const size_t i = ((size_t)b * T + t) * C + c;
      const uint32_t h = *reinterpret_cast<const uint32_t*>(rh + i), l = *reinterpret_cast<const uint32_t*>(rl + i);
      const float o0 = plane_lo16_f32(h) + plane_lo16_f32(l) + tile[2 * tx][r];
      const float o1 = plane_hi16_f32(h) + plane_hi16_f32(l) + tile[2 * tx + 1][r];
      uint16_t h0, l0, h1, l1;
      split_pair(o0, h0, l0);
      split_pair(o1, h1, l1);
      *reinterpret_cast<uint32_t*>(oh + i) = h0 | ((uint32_t)h1 << 16);
      *reinterpret_cast<uint32_t*>(ol + i) = l0 | ((uint32_t)l1 << 16);
    }
  }
}

// ---- the same two transposes, four elements per access (t_pad % 4 == 0, c % 4 == 0) --------------------------------------------
// One thread owns a 4 x 4 block: 8-byte plane accesses on both sides (16 lanes = one 128-byte line), the transpose of the block in
// registers, and the exchange between the "lanes along c" and the "lanes along t" mappings through 16-byte shared-memory units with an
// XOR swizzle (unit column ^= (row >> 2) & 7): conflict-free on both sides.  The two-byte version above spent 40-48 instructions per
// element on addressing (ncu: issue 60 %, DRAM 23-33 %).
__device__ __forceinline__ uint32_t pack_lo(uint32_t a, uint32_t b) { return (a & 0xFFFFu) | (b << 16); }
__device__ __forceinline__ uint32_t pack_hi(uint32_t a, uint32_t b) { return (a >> 16) | (b & 0xFFFF0000u); }
__global__ void __launch_bounds__(kThreads) tokens_to_channels_v4_kernel(const uint16_t* __restrict__ xh, const uint16_t* __restrict__ xl,
                                                                          uint16_t* __restrict__ yh, uint16_t* __restrict__ yl, int T, int C,
                                                                          int Tp) {
  __shared__ uint4 tile[64][16];                            // [channel][token quad ^ swizzle]: 4 tokens x (hi | lo << 16)
  const int b = blockIdx.z, t0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  {
    const int cq = threadIdx.x & 15, tq = threadIdx.x >> 4;
    const int c = c0 + 4 * cq;
    uint32_t e[4][4];                                       // [token][channel]
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + 4 * tq + j;
      uint2 h = make_uint2(0u, 0u), l = make_uint2(0u, 0u);
      if (t < T && c < C) {
        const size_t i = ((size_t)b * T + t) * C + c;
        h = __ldg(reinterpret_cast<const uint2*>(xh + i));
        l = __ldg(reinterpret_cast<const uint2*>(xl + i));
      }
      e[j][0] = pack_lo(h.x, l.x); e[j][1] = pack_hi(h.x, l.x);
      e[j][2] = pack_lo(h.y, l.y); e[j][3] = pack_hi(h.y, l.y);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) tile[4 * cq + i][tq ^ (cq & 7)] = make_uint4(e[0][i], e[1][i], e[2][i], e[3][i]);
  }
  __syncthreads();
  {
    const int tq = threadIdx.x & 15, cg = threadIdx.x >> 4;
    const int t = t0 + 4 * tq;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = cg + 16 * i, c = c0 + r;
      if (c < C && t < Tp) {
        const uint4 u = tile[r][tq ^ ((r >> 2) & 7)];
        const size_t o = ((size_t)b * C + c) * Tp + t;
        *reinterpret_cast<uint2*>(yh + o) = make_uint2(pack_lo(u.x, u.y), pack_lo(u.z, u.w));
        *reinterpret_cast<uint2*>(yl + o) = make_uint2(pack_hi(u.x, u.y), pack_hi(u.z, u.w));
      }
    }
  }
}
__global__ void __launch_bounds__(kThreads) channels_to_tokens_add_v4_kernel(const uint16_t* __restrict__ yh, const uint16_t* __restrict__ yl,
                                                                              const uint16_t* __restrict__ rh, const uint16_t* __restrict__ rl,
                                                                              uint16_t* __restrict__ oh, uint16_t* __restrict__ ol, int T, int C,
                                                                              int Tp) {
  __shared__ float4 tile[64][16];                           // [token][channel quad ^ swizzle]: 4 channels, merged to fp32
  const int b = blockIdx.z, t0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  {
    const int tq = threadIdx.x & 15, cq = threadIdx.x >> 4;
    const int t = t0 + 4 * tq;
    float v[4][4];                                          // [channel][token]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + 4 * cq + i;
      uint2 h = make_uint2(0u, 0u), l = make_uint2(0u, 0u);
      if (c < C && t < Tp) {
        const size_t k = ((size_t)b * C + c) * Tp + t;
        h = __ldg(reinterpret_cast<const uint2*>(yh + k));
        l = __ldg(reinterpret_cast<const uint2*>(yl + k));
      }
      v[i][0] = plane_lo16_f32(h.x) + plane_lo16_f32(l.x); v[i][1] = plane_hi16_f32(h.x) + plane_hi16_f32(l.x);
      v[i][2] = plane_lo16_f32(h.y) + plane_lo16_f32(l.y); v[i][3] = plane_hi16_f32(h.y) + plane_hi16_f32(l.y);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) tile[4 * tq + j][cq ^ (tq & 7)] = make_float4(v[0][j], v[1][j], v[2][j], v[3][j]);
  }
  __syncthreads();
  {
    const int cq = threadIdx.x & 15, tg = threadIdx.x >> 4;
    const int c = c0 + 4 * cq;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = tg + 16 * j, t = t0 + r;
      if (t < T && c < C) {
        const float4 u = tile[r][cq ^ ((r >> 2) & 7)];
        const size_t k = ((size_t)b * T + t) * C + c;
        const uint2 h = __ldg(reinterpret_cast<const uint2*>(rh + k)), l = __ldg(reinterpret_cast<const uint2*>(rl + k));
        uint2 qh, ql;
        split_pair2(plane_lo16_f32(h.x) + plane_lo16_f32(l.x) + u.x, plane_hi16_f32(h.x) + plane_hi16_f32(l.x) + u.y, qh.x, ql.x);
        split_pair2(plane_lo16_f32(h.y) + plane_lo16_f32(l.y) + u.z, plane_hi16_f32(h.y) + plane_hi16_f32(l.y) + u.w, qh.y, ql.y);
        *reinterpret_cast<uint2*>(oh + k) = qh;
        *reinterpret_cast<uint2*>(ol + k) = ql;
      }
    }
  }
}

inline unsigned grid_for(size_t items) {
  size_t b = (items + kThreads - 1) / kThreads;
  size_t cap = (size_t)b200r_num_sms() * 16;
  return (unsigned)(b < cap ? (b ? b : 1) : cap);
}
}  // namespace

extern "C" {

int b200r_layernorm(const uint16_t* x, uint16_t* y, const float* gamma, const float* beta, int rows, int c, float eps,
                    b200r_stream_t stream) {
  B200R_CHECK_ARG(x && y && gamma && beta, "null pointer");
  B200R_CHECK_ARG(rows > 0 && c % 8 == 0 && c <= 32 * 8 * kLnMaxVec, "c must be a multiple of 8 and <= %d", 32 * 8 * kLnMaxVec);
  const size_t cnt = (size_t)rows * c;
  layernorm_kernel<<<(unsigned)(((size_t)rows * 32 + kThreads - 1) / kThreads), kThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(x + cnt), reinterpret_cast<uint4*>(y),
      reinterpret_cast<uint4*>(y + cnt), gamma, beta, rows, c / 8, eps);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

static int patch_common(const void* img, uint16_t* y, int n, int h, int w, int ps, const float* mean, const float* stdv, bool u8,
                        cudaStream_t s) {
  B200R_CHECK_ARG(img && y && mean && stdv, "null pointer");
  B200R_CHECK_ARG(n > 0 && ps > 0 && h % ps == 0 && w % ps == 0 && (3 * ps * ps) % 8 == 0, "bad patch geometry");
  Norm3 nm;
  for (int i = 0; i < 3; ++i) { nm.mean[i] = mean[i]; nm.std[i] = stdv[i]; }
  const size_t cnt = (size_t)n * (h / ps) * (w / ps) * 3 * ps * ps;
  if (u8) patch_gather_kernel<true><<<grid_for(cnt / 8), kThreads, 0, s>>>(img, reinterpret_cast<uint4*>(y), reinterpret_cast<uint4*>(y + cnt), n, h, w, ps, nm);
  else patch_gather_kernel<false><<<grid_for(cnt / 8), kThreads, 0, s>>>(img, reinterpret_cast<uint4*>(y), reinterpret_cast<uint4*>(y + cnt), n, h, w, ps, nm);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}
int b200r_patch_gather_u8(const uint8_t* img, uint16_t* y, int n, int h, int w, int patch, const float* mean_host, const float* std_host,
                          b200r_stream_t stream) {
  return patch_common(img, y, n, h, w, patch, mean_host, std_host, true, as_stream(stream));
}
int b200r_patch_gather_f32(const float* img, uint16_t* y, int n, int h, int w, int patch, const float* mean_host, const float* std_host,
                           b200r_stream_t stream) {
  return patch_common(img, y, n, h, w, patch, mean_host, std_host, false, as_stream(stream));
}

int b200r_assemble_tokens(const uint16_t* x, const float* cls, const float* pos, uint16_t* y, int n, int num_patches, int c,
                          b200r_stream_t stream) {
  B200R_CHECK_ARG(x && cls && pos && y, "null pointer");
  B200R_CHECK_ARG(n > 0 && num_patches > 0 && c % 8 == 0, "bad shape");
  const size_t cin = (size_t)n * num_patches * c, cout = (size_t)n * (num_patches + 1) * c;
  assemble_tokens_kernel<<<grid_for(cout / 8), kThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(x + cin), cls, pos, reinterpret_cast<uint4*>(y),
      reinterpret_cast<uint4*>(y + cout), n, num_patches, c / 8);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_attention(const uint16_t* qkv, uint16_t* out, int n, int tokens, int heads, int head_dim, float scale, b200r_stream_t stream) {
  B200R_CHECK_ARG(qkv && out, "null pointer");
  B200R_CHECK_ARG(n > 0 && tokens > 0 && heads > 0, "bad shape");
  B200R_CHECK_ARG(head_dim == 64, "head_dim %d not supported (64 only)", head_dim);
  {
    // tensor-core kernel (attention_sm100.cu) for sequences up to 256 tokens; B200R_ATTN_TC=0 keeps the CUDA-core kernel
    static int use_tc = -1;
    if (use_tc < 0) { const char* e = getenv("B200R_ATTN_TC"); use_tc = (e && e[0] == '0') ? 0 : 1; }
    if (use_tc) {
      const int rc = b200r_attention_tc(qkv, out, n, tokens, heads, scale, as_stream(stream));
      if (rc != B200R_ENOTSUP) return rc;
    }
  }
  const size_t cin = (size_t)n * tokens * 3 * heads * head_dim, cout = (size_t)n * tokens * heads * head_dim;
  const int Tp = (tokens + 31) & ~31;
  const size_t smem = ((size_t)2 * tokens * (head_dim + 1) + (kAttnThreads / 32) * (head_dim + Tp)) * sizeof(float);
  B200R_CHECK_ARG(smem <= 220 * 1024, "sequence too long for the shared-memory attention kernel");
  B200R_CUDA(cudaFuncSetAttribute(attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_kernel<64><<<n * heads, kAttnThreads, smem, as_stream(stream)>>>(qkv, qkv + cin, out, out + cout, tokens, heads, scale);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_tokens_to_channels(const uint16_t* x, uint16_t* y, int b, int t, int c, int t_pad, b200r_stream_t stream) {
  B200R_CHECK_ARG(x && y && b > 0 && t > 0 && c > 0 && t_pad >= t, "bad arguments");
  const bool v4 = t_pad % 4 == 0 && c % 4 == 0;            // any t
  B200R_CHECK_ARG(v4 || (t % 2 == 0 && c % 2 == 0 && t_pad % 2 == 0), "tokens_to_channels needs (t_pad, c) multiples of 4, or even t, c, t_pad");
  const size_t cin = (size_t)b * t * c, cout = (size_t)b * c * t_pad;
  dim3 grid((t_pad + 63) / 64, (c + 63) / 64, b);
  if (v4)
    tokens_to_channels_v4_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(x, x + cin, y, y + cout, t, c, t_pad);
  else
    tokens_to_channels_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(x, x + cin, y, y + cout, b, t, c, t_pad);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_channels_to_tokens_add(const uint16_t* y, const uint16_t* res, uint16_t* out, int b, int t, int c, int t_pad,
                                 b200r_stream_t stream) {
  B200R_CHECK_ARG(y && res && out && b > 0 && t > 0 && c > 0 && t_pad >= t, "bad arguments");
  const bool v4 = t_pad % 4 == 0 && c % 4 == 0;            // any t
  B200R_CHECK_ARG(v4 || (t % 2 == 0 && c % 2 == 0 && t_pad % 2 == 0), "channels_to_tokens_add needs (t_pad, c) multiples of 4, or even t, c, t_pad");
  const size_t cy = (size_t)b * c * t_pad, cx = (size_t)b * t * c;
  dim3 grid((t + 63) / 64, (c + 63) / 64, b);
  if (v4)
    channels_to_tokens_add_v4_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(y, y + cy, res, res + cx, out, out + cx, t, c, t_pad);
  else
    channels_to_tokens_add_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(y, y + cy, res, res + cx, out, out + cx, b, t, c, t_pad);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

}  // extern "C"
