// ResNet stem in one kernel (fp16 precision mode): uint8 NHWC image -> ToTensor + Normalize -> conv1 7x7/s2/p3 (3 -> 64)
// -> folded BN -> ReLU -> MaxPool2d(3, 2, 1) -> fp16 NHWC [n, h/4, w/4, 64]
// (resnet_official.py:221-227,330-334 after imagenet_dataloader.py:78-79).
//
// Why a dedicated kernel.  As an im2col GEMM the stem has K = 147 (padded to 192) and N = 64: tiny per-pixel work, but the
// patch gather reads every input byte 12 times and its 411 MB (batch 256) fp16 output is written and read back by the
// pooling kernel.  The generic GEMM with a gathering producer ran at 0.60 ms + 0.13 ms for the pool -- 9x its roofline,
// bounded by the shared-memory traffic of the gather.  Here nothing is gathered at all:
//
//   * Implicit im2col through OVERLAPPING UMMA descriptors.  One normalised image row is staged in shared memory as
//     fp16 RGB0 pixels (8 bytes each, 4 zero pixels of padding on both sides).  The 7 taps of a kernel row for output
//     pixel ox are the 8 consecutive pixels starting at padded column 2*ox (the first one meets a zero weight), i.e. the
//     64 bytes at byte offset 16*ox.  In the no-swizzle K-major canonical layout an operand row is addressed as
//     start + (m % 8) * 16 + (m / 8) * SBO + j * LBO  (j = 16-byte K chunk); with SBO = 128 and LBO = 16 that is
//     start + 16 * (m + j): row m of the A operand IS the staged row at byte 16*m.  The windows overlap, the tensor core
//     does not care.  One kernel row of one output row = 2 MMAs (128 x 64 x 16) straight from the 1.8 KB row buffer.
//   * Each staged input row is used for the 3-4 output rows it contributes to (ky = p - 2*oy), accumulating into a ring
//     of 8 TMEM accumulators (8 x 64 columns = all 512): an output row is complete 7 input rows after it starts and is
//     drained by the epilogue warps while the MMAs of the next rows proceed.
//   * The epilogue keeps the vertical 3-max of the pooling window in registers (thread = output column, rows arrive in
//     order), exchanges columns through a swizzled 16 KB staging row for the horizontal 3-max and writes the pooled row
//     with coalesced 16-byte stores.  The 112x112 activation never exists in HBM.
//
// Traffic per image: 150 528 B read + 401 408 B written (56*56*64 fp16); tensor work 2*112*112*64*(7*32) = 360 MFLOP
// issued (147/224 useful).  Warp roles (416 threads, one CTA per SM, persistent over quarter-image units):
//   warps 0-3   epilogue (TMEM lane quadrant = warp index)
//   warp  4     TMEM allocation + MMA issue
//   warps 5-12  producers: load a raw uint8 row (24 B per lane), LUT-normalise to fp16, write the row buffer
#include "sm100_ptx.cuh"

namespace {

constexpr int SP_EPI_WARPS = 4;
constexpr int SP_MMA_WARP = 4;
constexpr int SP_PROD_WARP0 = 5;
constexpr int SP_PROD_WARPS = 8;
constexpr int SP_THREADS = (SP_PROD_WARP0 + SP_PROD_WARPS) * 32;   // 416
constexpr int SP_SLOTS = 16;                 // staged input rows in flight
constexpr int SP_SLOT_BYTES = 2176;          // >= 16 * (127 + 4) = 2096 (the 128-row operand reads past the 112 real columns)
constexpr int SP_ACC = 8;                    // TMEM accumulators (64 columns each)
constexpr int SP_UNIT_POOLED = 14;           // pooled rows per work unit
constexpr int SP_OFF_B = 0;                                  // 7 ky x 2 halves x [2 chunks][64 cout][16 B]
constexpr int SP_OFF_ROWS = SP_OFF_B + 14 * 2048;
constexpr int SP_OFF_STAGE = SP_OFF_ROWS + SP_SLOTS * SP_SLOT_BYTES;   // 2 x [128 columns][128 B]
constexpr int SP_OFF_LUT = SP_OFF_STAGE + 2 * 16384;         // [3][256] fp16
constexpr int SP_OFF_SB = SP_OFF_LUT + 3 * 256 * 2;          // scale[64], bias[64] fp32
constexpr int SP_OFF_BARS = SP_OFF_SB + 512;
constexpr int SP_SMEM = SP_OFF_BARS + (2 * SP_SLOTS + 2 * SP_ACC) * 8 + 16 + 128;

struct StemPoolParams {
  const uint8_t* img;      // [n, h, w, 3]
  const __half* wgt;       // [64, 192], column = ky*24 + kx*3 + c
  const float* scale;      // [64] (nullable)
  const float* bias;       // [64] (nullable)
  __half* y;               // [n, h/4, w/4, 64]
  int n, h, w;
  int units_per_img, n_units;
  float mean[3], std[3];
};

// K-major operand without swizzle: 8-row x 16-byte core matrices, `lbo` between the two K chunks of one MMA, `sbo`
// between 8-row groups
__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;      // descriptor version (sm_100); layout type 0 = no swizzle
  return d;
}

__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
  const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

struct Unit {
  int n, py0, py1, r_first, r_last, p_lo, p_hi;
};
__device__ __forceinline__ Unit make_unit(const StemPoolParams& p, int u) {
  Unit t;
  const int ph = p.h >> 2;
  t.n = u / p.units_per_img;
  const int q = u - t.n * p.units_per_img;
  t.py0 = q * SP_UNIT_POOLED;
  t.py1 = min(ph, t.py0 + SP_UNIT_POOLED) - 1;
  t.r_first = max(0, 2 * t.py0 - 1);          // conv rows the unit's pooled rows read
  t.r_last = 2 * t.py1 + 1;
  t.p_lo = max(3, 2 * t.r_first);             // padded input rows (p = iy + 3) with a non-zero contribution
  t.p_hi = min(p.h + 2, 2 * t.r_last + 6);
  return t;
}

__global__ void __launch_bounds__(SP_THREADS, 1) stem_pool_kernel(const StemPoolParams p) {
  extern __shared__ __align__(128) uint8_t sp_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sp_smem_raw) + 127) & ~(uintptr_t)127);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_base = sbase + SP_OFF_BARS;
  auto full_in = [&](int s) { return bar_base + 8u * s; };
  auto empty_in = [&](int s) { return bar_base + 8u * (SP_SLOTS + s); };
  auto acc_full = [&](int a) { return bar_base + 8u * (2 * SP_SLOTS + a); };
  auto acc_empty = [&](int a) { return bar_base + 8u * (2 * SP_SLOTS + SP_ACC + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SP_OFF_BARS + (2 * SP_SLOTS + 2 * SP_ACC) * 8);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int H = p.h, W = p.w, WO = W >> 1, PW = W >> 2, PH = H >> 2;

  // ---- one-time setup: barriers, TMEM, weights in operand layout, LUT, zeroed row ring -----------------------------
  if (threadIdx.x == 0) {
    for (int s = 0; s < SP_SLOTS; ++s) { mbar_init(full_in(s), 1); mbar_init(empty_in(s), 1); }
    for (int a = 0; a < SP_ACC; ++a) { mbar_init(acc_full(a), 1); mbar_init(acc_empty(a), SP_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == SP_MMA_WARP) {
    tmem_alloc(smem_u32(tmem_slot), SP_ACC * 64);
    tmem_relinquish();
  }
  {
    __half* sB = reinterpret_cast<__half*>(smem + SP_OFF_B);
    for (int idx = threadIdx.x; idx < 14 * 1024; idx += SP_THREADS) {
      const int kyh = idx >> 10, j = (idx >> 9) & 1, co = (idx >> 3) & 63, e = idx & 7;
      const int ky = kyh >> 1, t = 4 * (kyh & 1) + 2 * j + (e >> 2), c4 = e & 3;     // t = window position = kx + 1
      sB[idx] = (t >= 1 && c4 < 3) ? p.wgt[co * 192 + ky * 24 + (t - 1) * 3 + c4] : __float2half_rn(0.f);
    }
    __half* lut = reinterpret_cast<__half*>(smem + SP_OFF_LUT);
    for (int idx = threadIdx.x; idx < 768; idx += SP_THREADS) {
      const int c = idx >> 8, b = idx & 255;
      lut[idx] = __float2half_rn(((float)b / 255.0f - p.mean[c]) / p.std[c]);
    }
    float* sb = reinterpret_cast<float*>(smem + SP_OFF_SB);
    for (int idx = threadIdx.x; idx < 128; idx += SP_THREADS)
      sb[idx] = idx < 64 ? (p.scale ? p.scale[idx] : 1.f) : (p.bias ? p.bias[idx - 64] : 0.f);
    uint4* rows = reinterpret_cast<uint4*>(smem + SP_OFF_ROWS);
    for (int idx = threadIdx.x; idx < SP_SLOTS * SP_SLOT_BYTES / 16; idx += SP_THREADS) rows[idx] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= SP_PROD_WARP0) {
    // ================================ producers ================================
    const int pw = warp - SP_PROD_WARP0;
    const unsigned short* lut = reinterpret_cast<const unsigned short*>(smem + SP_OFF_LUT);
    const bool active = lane * 8 < W;
    uint2 cur[3] = {make_uint2(0, 0), make_uint2(0, 0), make_uint2(0, 0)};
    int cur_cnt = -1;
    auto commit_row = [&](const uint2 (&v)[3], int cnt) {
      const int slot = cnt % SP_SLOTS;
      mbar_wait(empty_in(slot), ((cnt / SP_SLOTS) & 1) ^ 1);
      if (active) {
        const uint32_t w6[6] = {v[0].x, v[0].y, v[1].x, v[1].y, v[2].x, v[2].y};
        const uint32_t dst = sbase + SP_OFF_ROWS + slot * SP_SLOT_BYTES + 32 + lane * 64;
#pragma unroll
        for (int pr = 0; pr < 4; ++pr) {
          uint32_t o[4];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int px = 2 * pr + k;
            const uint32_t b0 = (w6[(3 * px) >> 2] >> (8 * ((3 * px) & 3))) & 255u;
            const uint32_t b1 = (w6[(3 * px + 1) >> 2] >> (8 * ((3 * px + 1) & 3))) & 255u;
            const uint32_t b2 = (w6[(3 * px + 2) >> 2] >> (8 * ((3 * px + 2) & 3))) & 255u;
            o[2 * k] = (uint32_t)lut[b0] | ((uint32_t)lut[256 + b1] << 16);
            o[2 * k + 1] = (uint32_t)lut[512 + b2];
          }
          sts_v4(dst + pr * 16, make_uint4(o[0], o[1], o[2], o[3]));
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_in(slot));
    };
    int cnt = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const Unit t = make_unit(p, u);
      for (int pp = t.p_lo; pp <= t.p_hi; ++pp, ++cnt) {
        if ((cnt & (SP_PROD_WARPS - 1)) != pw) continue;
        uint2 nxt[3] = {make_uint2(0, 0), make_uint2(0, 0), make_uint2(0, 0)};
        if (active) {
          const uint2* src = reinterpret_cast<const uint2*>(p.img + ((size_t)t.n * H + (pp - 3)) * (size_t)W * 3) + lane * 3;
          nxt[0] = __ldg(src); nxt[1] = __ldg(src + 1); nxt[2] = __ldg(src + 2);
        }
        if (cur_cnt >= 0) commit_row(cur, cur_cnt);
        cur[0] = nxt[0]; cur[1] = nxt[1]; cur[2] = nxt[2];
        cur_cnt = cnt;
      }
    }
    if (cur_cnt >= 0) commit_row(cur, cur_cnt);
  } else if (warp == SP_MMA_WARP) {
    // ================================ MMA issue ================================
    // fp16 A/B (format bits 0), fp32 accumulator (bit 4), K-major both, N = 64, M = 128
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int cnt = 0, acc_base = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const Unit t = make_unit(p, u);
      for (int pp = t.p_lo; pp <= t.p_hi; ++pp, ++cnt) {
        const int slot = cnt % SP_SLOTS;
        mbar_wait(full_in(slot), (cnt / SP_SLOTS) & 1);
        tc_fence_after();
        const uint32_t a_addr = sbase + SP_OFF_ROWS + slot * SP_SLOT_BYTES;
        const uint64_t a0 = make_nosw_desc(a_addr, 16, 128), a1 = make_nosw_desc(a_addr + 32, 16, 128);
        const int oy_lo = max(t.r_first, (pp - 5) >> 1), oy_hi = min(t.r_last, pp >> 1);   // ky = pp - 2*oy in [0, 6]
        for (int oy = oy_lo; oy <= oy_hi; ++oy) {
          const int ky = pp - 2 * oy;
          const int c = acc_base + (oy - t.r_first), aslot = c % SP_ACC;
          const bool first = (pp == max(2 * oy, 3));
          if (first) {
            mbar_wait(acc_empty(aslot), ((c / SP_ACC) & 1) ^ 1);
            tc_fence_after();
          }
          const uint32_t b_addr = sbase + SP_OFF_B + ky * 4096;
          const uint64_t b0 = make_nosw_desc(b_addr, 1024, 128), b1 = make_nosw_desc(b_addr + 2048, 1024, 128);
          const uint32_t d = tmem_base + aslot * 64;
          if (elect_one()) {
            umma_bf16(d, a0, b0, IDESC, first ? 0u : 1u);
            umma_bf16(d, a1, b1, IDESC, 1u);
          }
          __syncwarp();
        }
        if (elect_one()) {
          umma_commit(empty_in(slot));
          // output rows whose last contributing input row this is
          if (!(pp & 1) && pp >= 6) {
            const int oy = (pp - 6) >> 1;
            if (oy >= t.r_first && oy <= t.r_last) umma_commit(acc_full((acc_base + oy - t.r_first) % SP_ACC));
          }
          if (pp == H + 2) {          // the bottom row loses its last two taps to the padding
            const int oy = (H >> 1) - 1;
            if (oy >= t.r_first && oy <= t.r_last) umma_commit(acc_full((acc_base + oy - t.r_first) % SP_ACC));
          }
        }
        __syncwarp();
      }
      acc_base += t.r_last - t.r_first + 1;
    }
  } else {
    // ================================ epilogue ================================
    const int m = threadIdx.x;                   // output column (TMEM lane)
    const float4* sb4 = reinterpret_cast<const float4*>(smem + SP_OFF_SB);
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    int acc_base = 0, emit = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const Unit t = make_unit(p, u);
      uint32_t prev_odd[32], v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) { prev_odd[i] = 0u; v[i] = 0u; }
      for (int r = t.r_first; r <= t.r_last; ++r) {
        const int c = acc_base + (r - t.r_first), aslot = c % SP_ACC;
        mbar_wait(acc_full(aslot), (c / SP_ACC) & 1);
        tc_fence_after();
        const bool odd = r & 1;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t acc[32];
          tmem_ld32(lane_base + aslot * 64 + half * 32, acc);
          tmem_ld_wait();
          if (half == 1) {               // accumulator drained: hand it back before the arithmetic
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(aslot));
          }
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {           // 4 channels per step: scale / bias as broadcast 128-bit reads
            const float4 s4 = sb4[half * 8 + j4], b4 = sb4[16 + half * 8 + j4];
            const float y0 = fmaxf(fmaf(__uint_as_float(acc[4 * j4]), s4.x, b4.x), 0.f);
            const float y1 = fmaxf(fmaf(__uint_as_float(acc[4 * j4 + 1]), s4.y, b4.y), 0.f);
            const float y2 = fmaxf(fmaf(__uint_as_float(acc[4 * j4 + 2]), s4.z, b4.z), 0.f);
            const float y3 = fmaxf(fmaf(__uint_as_float(acc[4 * j4 + 3]), s4.w, b4.w), 0.f);
            const uint32_t cur0 = cvt_f16x2(y1, y0), cur1 = cvt_f16x2(y3, y2);
            const int i = half * 16 + 2 * j4;
            if (odd) { v[i] = hmax2_u32(v[i], cur0); prev_odd[i] = cur0; v[i + 1] = hmax2_u32(v[i + 1], cur1); prev_odd[i + 1] = cur1; }
            else { v[i] = hmax2_u32(prev_odd[i], cur0); v[i + 1] = hmax2_u32(prev_odd[i + 1], cur1); }
          }
        }
        if (odd && (r >> 1) >= t.py0) {
          // vertical 3-max of pooled row r/2 is in v: exchange columns through the staging row, then the horizontal 3-max
          const uint32_t st = sbase + SP_OFF_STAGE + (emit & 1) * 16384;
          if (m < WO) {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
              sts_v4(st + m * 128 + ((ch ^ (m & 7)) << 4), make_uint4(v[4 * ch], v[4 * ch + 1], v[4 * ch + 2], v[4 * ch + 3]));
          }
          named_bar_sync(1, SP_EPI_WARPS * 32);
          uint4* yrow = reinterpret_cast<uint4*>(p.y + ((size_t)(t.n * PH + (r >> 1)) * PW) * 64);
          for (int i = m; i < PW * 8; i += SP_EPI_WARPS * 32) {
            const int px = i >> 3, ch = i & 7, c1 = 2 * px;
            uint4 a = lds_v4(st + c1 * 128 + ((ch ^ (c1 & 7)) << 4));
            const uint4 b = lds_v4(st + (c1 + 1) * 128 + ((ch ^ ((c1 + 1) & 7)) << 4));
            a.x = hmax2_u32(a.x, b.x); a.y = hmax2_u32(a.y, b.y); a.z = hmax2_u32(a.z, b.z); a.w = hmax2_u32(a.w, b.w);
            if (px > 0) {
              const uint4 l = lds_v4(st + (c1 - 1) * 128 + ((ch ^ ((c1 - 1) & 7)) << 4));
              a.x = hmax2_u32(a.x, l.x); a.y = hmax2_u32(a.y, l.y); a.z = hmax2_u32(a.z, l.z); a.w = hmax2_u32(a.w, l.w);
            }
            yrow[i] = a;
          }
          ++emit;
        }
      }
      acc_base += t.r_last - t.r_first + 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == SP_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, SP_ACC * 64);
  }
}

}  // namespace

extern "C" {

int b200r_stem_pool_u8_f16(const uint8_t* img, const uint16_t* wgt, const float* scale, const float* bias, uint16_t* y,
                           int n, int h, int w, const float* mean_host, const float* std_host, b200r_stream_t stream) {
  B200R_CHECK_ARG(img && wgt && y && mean_host && std_host, "null pointer");
  B200R_CHECK_ARG(n > 0 && h >= 8 && w >= 8, "bad shape %d x %d x %d", n, h, w);
  B200R_CHECK_ARG(h % 4 == 0 && w % 8 == 0 && w <= 248, "stem_pool needs h %% 4 == 0, w %% 8 == 0, w <= 248 (got %d x %d)", h, w);
  B200R_CHECK_ARG((reinterpret_cast<uintptr_t>(img) & 7) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, "img must be 8-byte, y 16-byte aligned");
  StemPoolParams p;
  p.img = img; p.wgt = reinterpret_cast<const __half*>(wgt); p.scale = scale; p.bias = bias; p.y = reinterpret_cast<__half*>(y);
  p.n = n; p.h = h; p.w = w;
  p.units_per_img = ((h >> 2) + SP_UNIT_POOLED - 1) / SP_UNIT_POOLED;
  p.n_units = n * p.units_per_img;
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean_host[c]; p.std[c] = std_host[c]; }
  static bool attr_set[16] = {};
  int dev = 0;
  B200R_CUDA(cudaGetDevice(&dev));
  if (dev < 16 && !attr_set[dev]) {
    B200R_CUDA(cudaFuncSetAttribute(stem_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM));
    attr_set[dev] = true;
  }
  const int grid = p.n_units < b200r_num_sms() ? p.n_units : b200r_num_sms();
  stem_pool_kernel<<<grid, SP_THREADS, SP_SMEM, as_stream(stream)>>>(p);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

}  // extern "C"
