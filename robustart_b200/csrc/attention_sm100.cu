// Multi-head self-attention on the 5th-generation tensor cores (ViT-B/16: 197 tokens, 12 heads of 64;
// vision_transformer.py:80-92): out = softmax(q k^T * scale) v per (image, head), split-bf16 planes in and out.
//
// One CTA per (image, head, 128-query tile); tokens <= 256 so one pass over the keys is the whole softmax -- no online
// rescaling.  544 threads:
//   warp 16     TMA loads (Q tile, all K, all V; hi and lo planes), then the MMA issue:
//                 S[128 x Tk]  = Q K^T          3 x 4 MMAs (lo*hi, hi*lo, hi*hi), A = Q and B = K both K-major SW128 tiles
//                 O[128 x 64]  = P V            3 x Tk/16 MMAs (hi*hi, lo*hi, hi*lo), A = P K-major, B = V **MN-major** (V is
//                                               [key][d] in memory = exactly what TMA delivers; no transpose pass)
//   warps 0-15  softmax: thread = (query row = TMEM lane, quarter of the key chunks); four warps per TMEM lane quadrant
//               (warp % 4) keep each scheduler's issue slots busy -- with one warp per quadrant the 4.6 K dependent
//               instructions of a row were the kernel (18 K of a tile's 26 K cycles).  Row maxima and sums are combined
//               through shared memory.  Pass 1 reads S for the row maximum, pass 2 computes
//               p = 2^((s - max) * scale * log2 e) (masked beyond the last token), accumulates the row sum and writes p as
//               bf16 hi / lo into the A-operand layout (over the dead Q / K tiles); the 1 / sum goes onto O in the epilogue,
//               which re-splits to bf16 planes and stores 128-byte rows.
// fp32-faithful like the GEMMs: every product is hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM.
// Cost per (image, head): 4 * T^2 * 64 FLOP algorithmic; bytes: 3 * T * 64 * 4 read (K, V once per query tile), T * 64 * 4 written.
#include "sm100_ptx.cuh"
#include <mutex>

namespace {

constexpr int AT_D = 64;
constexpr int AT_SM_WARPS = 16;                       // softmax warps: 4 column groups x 4 TMEM lane quadrants
constexpr int AT_THREADS = (AT_SM_WARPS + 1) * 32;
constexpr int AT_MMA_WARP = AT_SM_WARPS;
constexpr int AT_MAXTK = 256;
constexpr int AT_P_PLANE = 4 * 16384;                 // P: up to 4 k-blocks of [128 x 64] per plane

struct AttnParams {
  uint16_t* out_hi; uint16_t* out_lo;
  int T, Tk, H, MT;
  float scale_log2e;
};

// B operand, MN-major (rows = K index, 64 contiguous N elements = one 128-byte swizzled row): 8-row atoms 1024 B apart
__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024 >> 4) << 16;                   // leading byte offset: next 64-element N block (unused, N = 64)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset: next group of 8 K rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int Tk = p.Tk;
  const uint32_t kv_bytes = (uint32_t)Tk * 128;
  // [V hi | V lo | region R], R = Q hi | Q lo | K hi | K lo, later overlaid by P hi | P lo
  const uint32_t off_v = 0, off_r = 2 * kv_bytes;
  const uint32_t off_q = off_r, off_k = off_r + 2 * 16384;
  const uint32_t off_bar = off_r + 2 * AT_P_PLANE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + off_bar);      // ld_full, s_full, p_full, o_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* red = reinterpret_cast<float*>(smem + off_bar + 64);         // [2][4 groups][128 rows]: row maxima, row sums
  const uint32_t bar0 = sbase + off_bar;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;

  const int mt = blockIdx.x % p.MT, hd = (blockIdx.x / p.MT) % p.H, b = blockIdx.x / (p.MT * p.H);
  const int row0 = b * p.T;                       // first token row of this image in the [n*T, .] matrices

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_q);
    prefetch_tmap(&map_kv);
    mbar_init(bar0 + 0, 1);
    mbar_init(bar0 + 8, 1);
    mbar_init(bar0 + 16, AT_SM_WARPS * 32);
    mbar_init(bar0 + 24, 1);
    fence_barrier_init();
  }
  if (warp == AT_MMA_WARP) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_s = tmem_base, tm_o = tmem_base + 256;

  if (warp == AT_MMA_WARP) {
    // fp16 A/B (format bits 7, 10 = 0), fp32 accumulator (bit 4), M = 128
    const uint32_t idesc_s = (1u << 4) | ((uint32_t)(Tk >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_o = (1u << 4) | (1u << 16) /* B is MN-major */ | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (elect_one()) {
      mbar_expect_tx(bar0, 2u * 16384u + 4u * kv_bytes);
      for (int pl = 0; pl < 2; ++pl) {
        tma_load_3d(sbase + off_q + pl * 16384, &map_q, bar0, hd * AT_D, row0 + mt * 128, pl);
        tma_load_3d(sbase + off_k + pl * kv_bytes, &map_kv, bar0, p.H * AT_D + hd * AT_D, row0, pl);
        tma_load_3d(sbase + off_v + pl * kv_bytes, &map_kv, bar0, 2 * p.H * AT_D + hd * AT_D, row0, pl);
      }
    }
    __syncwarp();
    mbar_wait(bar0, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint64_t q_hi = make_sw128_desc(sbase + off_q), q_lo = make_sw128_desc(sbase + off_q + 16384);
      const uint64_t k_hi = make_sw128_desc(sbase + off_k), k_lo = make_sw128_desc(sbase + off_k + kv_bytes);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adv = (uint64_t)((k * 32) >> 4);          // +32 B per K step inside the swizzle row
        umma_bf16(tm_s, q_lo + adv, k_hi + adv, idesc_s, k != 0);
        umma_bf16(tm_s, q_hi + adv, k_lo + adv, idesc_s, 1);
        umma_bf16(tm_s, q_hi + adv, k_hi + adv, idesc_s, 1);
      }
      umma_commit(bar0 + 8);
    }
    __syncwarp();
    mbar_wait(bar0 + 16, 0);                                     // P is in shared memory
    tc_fence_after();
    if (elect_one()) {
      const int ksteps = Tk >> 4;
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint32_t pa = sbase + off_r + (ks >> 2) * 16384 + (ks & 3) * 32;
        const uint32_t vb = sbase + off_v + ks * 2048;
        const uint64_t p_hi = make_sw128_desc(pa), p_lo = make_sw128_desc(pa + AT_P_PLANE);
        const uint64_t v_hi = make_sw128_mn_desc(vb), v_lo = make_sw128_mn_desc(vb + kv_bytes);
        umma_bf16(tm_o, p_lo, v_hi, idesc_o, ks != 0);
        umma_bf16(tm_o, p_hi, v_lo, idesc_o, 1);
        umma_bf16(tm_o, p_hi, v_hi, idesc_o, 1);
      }
      umma_commit(bar0 + 24);
    }
    __syncwarp();
  } else {
    // ================================ softmax + epilogue ================================
    const int quad = warp & 3, grp = warp >> 2;                   // TMEM lane quadrant, column group
    const int row = quad * 32 + lane;                             // query row of the tile = TMEM lane
    const int q = mt * 128 + row;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int chunks = Tk >> 5;                                   // 32-key chunks; this group takes chunks grp, grp + 4, ...
    mbar_wait(bar0 + 8, 0);
    tc_fence_after();
    float mx = -INFINITY;
    for (int c = grp; c < chunks; c += 4) {
      uint32_t v[32];
      tmem_ld32(tm_s + lane_addr + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c * 32 + j < p.T) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    red[grp * 128 + row] = mx;
    named_bar_sync(1, AT_SM_WARPS * 32);
    mx = fmaxf(fmaxf(red[row], red[128 + row]), fmaxf(red[256 + row], red[384 + row]));
    float sum = 0.f;
    const float k2 = p.scale_log2e, mk = -mx * k2;
    for (int c = grp; c < chunks; c += 4) {
      uint32_t v[32];
      tmem_ld32(tm_s + lane_addr + c * 32, v);
      tmem_ld_wait();
      uint8_t* tile = smem + off_r + (c >> 1) * 16384 + row * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g) {                               // 8 keys = one 16-byte chunk per plane
        uint32_t ph[4], pl[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = c * 32 + g * 8 + 2 * j;
          const float e0 = col < p.T ? ex2_approx(fmaf(__uint_as_float(v[g * 8 + 2 * j]), k2, mk)) : 0.f;
          const float e1 = col + 1 < p.T ? ex2_approx(fmaf(__uint_as_float(v[g * 8 + 2 * j + 1]), k2, mk)) : 0.f;
          sum += e0 + e1;
          split_f16x2(e1, e0, ph[j], pl[j]);
        }
        const int chunk = ((c & 1) * 4 + g) ^ (row & 7);          // SWIZZLE_128B position of this 16-byte chunk
        *reinterpret_cast<uint4*>(tile + (chunk << 4)) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        *reinterpret_cast<uint4*>(tile + AT_P_PLANE + (chunk << 4)) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
      }
    }
    red[512 + grp * 128 + row] = sum;
    tc_fence_before();
    fence_proxy_async();
    mbar_arrive(bar0 + 16);
    named_bar_sync(1, AT_SM_WARPS * 32);                          // every group's partial sum is in shared memory
    const float inv = 1.f / ((red[512 + row] + red[640 + row]) + (red[768 + row] + red[896 + row]));
    // ---- O / sum -> bf16 planes: group g takes the 16 output columns [16g, 16g + 16) ----
    mbar_wait(bar0 + 24, 0);
    tc_fence_after();
    const size_t orow = ((size_t)(row0 + q) * p.H + hd) * AT_D + grp * 16;
    {
      uint32_t v[16];
      tmem_ld16(tm_o + lane_addr + grp * 16, v);
      tmem_ld_wait();
      if (q < p.T) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t ph[4], pl[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x0 = __uint_as_float(v[g * 8 + 2 * j]) * inv, x1 = __uint_as_float(v[g * 8 + 2 * j + 1]) * inv;
            split_f16x2(x1, x0, ph[j], pl[j]);
          }
          *reinterpret_cast<uint4*>(p.out_hi + orow + g * 8) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
          *reinterpret_cast<uint4*>(p.out_lo + orow + g * 8) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AT_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn at_get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

}  // namespace

// returns B200R_ENOTSUP when the geometry is outside the tensor-core kernel (the caller then takes the CUDA-core kernel)
int b200r_attention_tc(const uint16_t* qkv, uint16_t* out, int n, int tokens, int heads, float scale, cudaStream_t stream) {
  if (tokens > AT_MAXTK || tokens < 1) return B200R_ENOTSUP;
  EncodeTiledFn enc = at_get_encode();
  if (!enc) { b200r_set_error("cuTensorMapEncodeTiled entry point not found"); return B200R_ECUDA; }
  const int Tk = (tokens + 31) & ~31, MT = (tokens + 127) / 128;
  const cuuint64_t cols = (cuuint64_t)3 * heads * AT_D, rows = (cuuint64_t)n * tokens;
  CUtensorMap mq, mkv;
  cuuint64_t dims[3] = {cols, rows, 2};
  cuuint64_t strides[2] = {cols * 2, rows * cols * 2};
  cuuint32_t estr[3] = {1, 1, 1};
  cuuint32_t box_q[3] = {AT_D, 128, 1}, box_kv[3] = {AT_D, (cuuint32_t)Tk, 1};
  CUresult r = enc(&mq, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<uint16_t*>(qkv), dims, strides, box_q, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS)
    r = enc(&mkv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<uint16_t*>(qkv), dims, strides, box_kv, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { b200r_set_error("cuTensorMapEncodeTiled(attention) failed: %d", (int)r); return B200R_ECUDA; }
  AttnParams p;
  const size_t cout = (size_t)n * tokens * heads * AT_D;
  p.out_hi = out; p.out_lo = out + cout;
  p.T = tokens; p.Tk = Tk; p.H = heads; p.MT = MT;
  p.scale_log2e = scale * 1.4426950408889634f;
  const int smem = 2 * Tk * 128 + 2 * AT_P_PLANE + 64 + 2 * 4 * 128 * 4 + 1024;
  static int configured = 0;
  if (configured < smem) {
    B200R_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  attention_tc_kernel<<<n * heads * MT, AT_THREADS, smem, stream>>>(mq, mkv, p);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}
