// Backward of multi-head self-attention on the 5th-generation tensor cores (ViT-B/16: 197 tokens, 12 heads of 64; autograd of
// vision_transformer.py:80-92, which is what autopgd_base.py:371-376 / foolbox value_and_grad pull through the model every attack
// step): from packed qkv and dO to d qkv, probabilities recomputed (the forward saves nothing but qkv).
//     S = scale Q K^T,  P = softmax(S),  O = P V
//     dV = P^T dO,  dP = dO V^T,  delta_i = sum_j P_ij dP_ij,  dS = P o (dP - delta),  dQ = scale dS K,  dK = scale dS^T Q
// The CUDA-core first version (token_backward.cu) was 79 % of a ViT-B/16 gradient step (99.6 of 126 ms at batch 128,
// profiles/r2_op_times_vit_grad.txt).  Here every contraction is a tcgen05 MMA with operands in their natural layouts; no operand is
// ever transposed in shared memory, because the work is split into two launches of one kernel template:
//   phase A  one CTA per (image, head, 128-QUERY tile):  S = Q_t K^T and dP = dO_t V^T (3-pass split fp16, fp32 in TMEM), row softmax
//            statistics and delta in registers (thread = query row = TMEM lane), dS written as the A operand, dQ_t = dS K with K
//            consumed MN-major as TMA delivered it.  The row statistics (max, 1/sum, delta) go to a small workspace.
//   phase B  one CTA per (image, head, 128-KEY tile):   S^T = K_t Q^T and dP^T = V_t dO^T (the same two products with the roles of the
//            operands swapped), P^T and dS^T rebuilt per element from the per-QUERY statistics of phase A (thread = key row), written as
//            A operands, dV_t = P^T dO and dK_t = dS^T Q with dO / Q consumed MN-major.  No atomics, no cross-CTA accumulation.
// Precision: the score products and phase A's dQ are 3-pass (hi*hi + hi*lo + lo*hi); phase B's two output products take the hi planes
// of P^T / dS^T and of dO / Q (one MMA per product: the P^T / dS^T operand pair only fits shared memory as single fp16 planes).  This is
// the gradient pass -- the attacks consume sign(g) or g / ||g|| -- and holds 2e-4 of max |d qkv| against fp64 autograd
// (tests/test_token_grad_gpu.py::test_attention_bwd).
// Cost per (image, head): 14 T^2 64 FLOP algorithmic (7 T x T x 64 products incl. the recomputed scores, twice for the two phases'
// shared S / dP); bytes: 4 T 64 4 read per phase, 3 T 64 4 written.
#include "sm100_ptx.cuh"
#include <mutex>

namespace {

constexpr int AB_D = 64;
constexpr int AB_SM_WARPS = 16;                       // elementwise warps: 4 column groups x 4 TMEM lane quadrants
constexpr int AB_THREADS = (AB_SM_WARPS + 1) * 32;
constexpr int AB_MMA_WARP = AB_SM_WARPS;
constexpr int AB_MAXTK = 256;
constexpr int AB_OP_PLANE = 4 * 16384;                // an A operand [128 x Tk <= 256]: up to 4 k-blocks of [128 x 64] per plane

struct AttnBwdParams {
  uint16_t* d_hi; uint16_t* d_lo;                     // dqkv planes [n*T, 3*H*64]
  float* stats;                                       // [n*H][3][T]: -max*k2, 1/sum, delta per query row
  int T, Tk, H, MT;
  float scale, scale_log2e;
};

__device__ __forceinline__ uint64_t ab_mn_desc(uint32_t smem_addr) {      // B operand, MN-major, SWIZZLE_128B (see attention_sm100.cu)
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ float ab_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// PHASE 0 = A (rows are queries), 1 = B (rows are keys).  map_xt / map_gt: 128-row boxes of the row-side matrices (A: Q, dO; B: K, V);
// map_ya / map_za: Tk-row boxes of the column-side matrices (A: K, V; B: Q, dO).
template <int PHASE>
__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv_t, const __grid_constant__ CUtensorMap map_qkv_a,
                        const __grid_constant__ CUtensorMap map_do_t, const __grid_constant__ CUtensorMap map_do_a, const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t ab_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ab_smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int Tk = p.Tk;
  const uint32_t kv = (uint32_t)Tk * 128;             // bytes of one plane of a [Tk x 64] matrix
  // R0 (2 kv, stays for the output products) | R1 (128 KB, loaded operands first, the A operands of the output products later)
  //   phase A: R0 = K hi | K lo;   R1 = Q_t hi | Q_t lo | dO_t hi | dO_t lo | V hi | V lo          -> dS hi | dS lo
  //   phase B: R0 = Q hi | dO hi;  R1 = Q lo | dO lo | K_t hi | K_t lo | V_t hi | V_t lo          -> P^T hi | dS^T hi
  const uint32_t off_r1 = 2 * kv;
  uint32_t off_x, off_g, off_y_hi, off_y_lo, off_z_hi, off_z_lo;
  if (PHASE == 0) {
    off_y_hi = 0; off_y_lo = kv;
    off_x = off_r1; off_g = off_r1 + 32768; off_z_hi = off_r1 + 65536; off_z_lo = off_z_hi + kv;
  } else {
    off_y_hi = 0; off_z_hi = kv;
    off_y_lo = off_r1; off_z_lo = off_r1 + kv; off_x = off_r1 + 2 * kv; off_g = off_x + 32768;
  }
  const uint32_t off_bar = off_r1 + 2 * AB_OP_PLANE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + off_bar);      // loads, scores, operands, outputs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* red = reinterpret_cast<float*>(smem + off_bar + 64);         // phase A: [3][4 groups][128 rows]; phase B: [3][256] column statistics
  const uint32_t bar0 = sbase + off_bar;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;

  const int mt = blockIdx.x % p.MT, hd = (blockIdx.x / p.MT) % p.H, b = blockIdx.x / (p.MT * p.H);
  const int row0 = b * p.T;
  const int C1 = p.H * AB_D;
  float* st = p.stats + (size_t)(b * p.H + hd) * 3 * p.T;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_qkv_t); prefetch_tmap(&map_qkv_a); prefetch_tmap(&map_do_t); prefetch_tmap(&map_do_a);
    mbar_init(bar0 + 0, 1);
    mbar_init(bar0 + 8, 1);
    mbar_init(bar0 + 16, AB_SM_WARPS * 32);
    mbar_init(bar0 + 24, 1);
    fence_barrier_init();
  }
  if (warp == AB_MMA_WARP) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  if (PHASE == 1)                                       // per-query statistics of phase A -> shared memory (columns of this phase)
    for (int i = threadIdx.x; i < 3 * AB_MAXTK; i += AB_THREADS) {
      const int s = i / AB_MAXTK, c = i % AB_MAXTK;
      red[i] = c < p.T ? st[s * p.T + c] : 0.f;
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_s = tmem_base, tm_e = tmem_base + 256;       // scores S (or S^T), dP (or dP^T); outputs reuse columns 0.. and 256..

  if (warp == AB_MMA_WARP) {
    const uint32_t idesc_s = (1u << 4) | ((uint32_t)(Tk >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_o = (1u << 4) | (1u << 16) /* B is MN-major */ | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // column offsets inside qkv of the row-side / column-side matrices
    const int col_x = (PHASE == 0 ? 0 : C1) + hd * AB_D, col_y = (PHASE == 0 ? C1 : 0) + hd * AB_D;
    if (elect_one()) {
      mbar_expect_tx(bar0, 4u * 16384u + 4u * kv);
      for (int pl = 0; pl < 2; ++pl) {
        tma_load_3d(sbase + off_x + pl * 16384, &map_qkv_t, bar0, col_x, row0 + mt * 128, pl);                       // Q_t | K_t
        tma_load_3d(sbase + (pl ? off_y_lo : off_y_hi), &map_qkv_a, bar0, col_y, row0, pl);                          // K   | Q
        if (PHASE == 0) {
          tma_load_3d(sbase + off_g + pl * 16384, &map_do_t, bar0, hd * AB_D, row0 + mt * 128, pl);                  // dO_t
          tma_load_3d(sbase + (pl ? off_z_lo : off_z_hi), &map_qkv_a, bar0, 2 * C1 + hd * AB_D, row0, pl);           // V
        } else {
          tma_load_3d(sbase + off_g + pl * 16384, &map_qkv_t, bar0, 2 * C1 + hd * AB_D, row0 + mt * 128, pl);        // V_t
          tma_load_3d(sbase + (pl ? off_z_lo : off_z_hi), &map_do_a, bar0, hd * AB_D, row0, pl);                     // dO
        }
      }
    }
    __syncwarp();
    mbar_wait(bar0, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint64_t x_hi = make_sw128_desc(sbase + off_x), x_lo = make_sw128_desc(sbase + off_x + 16384);
      const uint64_t g_hi = make_sw128_desc(sbase + off_g), g_lo = make_sw128_desc(sbase + off_g + 16384);
      const uint64_t y_hi = make_sw128_desc(sbase + off_y_hi), y_lo = make_sw128_desc(sbase + off_y_lo);
      const uint64_t z_hi = make_sw128_desc(sbase + off_z_hi), z_lo = make_sw128_desc(sbase + off_z_lo);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adv = (uint64_t)((k * 32) >> 4);
        umma_bf16(tm_s, x_lo + adv, y_hi + adv, idesc_s, k != 0);
        umma_bf16(tm_s, x_hi + adv, y_lo + adv, idesc_s, 1);
        umma_bf16(tm_s, x_hi + adv, y_hi + adv, idesc_s, 1);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adv = (uint64_t)((k * 32) >> 4);
        umma_bf16(tm_e, g_lo + adv, z_hi + adv, idesc_s, k != 0);
        umma_bf16(tm_e, g_hi + adv, z_lo + adv, idesc_s, 1);
        umma_bf16(tm_e, g_hi + adv, z_hi + adv, idesc_s, 1);
      }
      umma_commit(bar0 + 8);
    }
    __syncwarp();
    mbar_wait(bar0 + 16, 0);                                     // the A operands are in shared memory, S / dP have been consumed
    tc_fence_after();
    if (elect_one()) {
      const int ksteps = Tk >> 4;
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint32_t oa = sbase + off_r1 + (ks >> 2) * 16384 + (ks & 3) * 32;
        if (PHASE == 0) {                                        // dQ_t = dS K, 3-pass
          const uint64_t a_hi = make_sw128_desc(oa), a_lo = make_sw128_desc(oa + AB_OP_PLANE);
          const uint64_t k_hi = ab_mn_desc(sbase + off_y_hi + ks * 2048), k_lo = ab_mn_desc(sbase + off_y_lo + ks * 2048);
          umma_bf16(tm_s, a_lo, k_hi, idesc_o, ks != 0);
          umma_bf16(tm_s, a_hi, k_lo, idesc_o, 1);
          umma_bf16(tm_s, a_hi, k_hi, idesc_o, 1);
        } else {                                                 // dV_t = P^T dO, dK_t = dS^T Q: hi planes
          umma_bf16(tm_s, make_sw128_desc(oa), ab_mn_desc(sbase + off_z_hi + ks * 2048), idesc_o, ks != 0);
          umma_bf16(tm_e, make_sw128_desc(oa + AB_OP_PLANE), ab_mn_desc(sbase + off_y_hi + ks * 2048), idesc_o, ks != 0);
        }
      }
      umma_commit(bar0 + 24);
    }
    __syncwarp();
  } else {
    // ================================ elementwise stage + epilogue ================================
    const int quad = warp & 3, grp = warp >> 2;                   // TMEM lane quadrant, column group
    const int row = quad * 32 + lane;                             // row of the tile = TMEM lane
    const int r = mt * 128 + row;                                 // query (A) / key (B) index inside the image
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int chunks = Tk >> 5;                                   // 32-column chunks; this group takes chunks grp, grp + 4, ...
    const float k2 = p.scale_log2e;
    mbar_wait(bar0 + 8, 0);
    tc_fence_after();
    if (PHASE == 0) {
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
      named_bar_sync(1, AB_SM_WARPS * 32);
      mx = fmaxf(fmaxf(red[row], red[128 + row]), fmaxf(red[256 + row], red[384 + row]));
      const float mk = -mx * k2;
      float sum = 0.f, dot = 0.f;                                 // sum_j p_ij, sum_j p_ij dP_ij (unnormalised)
      for (int c = grp; c < chunks; c += 4) {
        uint32_t v[32], e[32];
        tmem_ld32(tm_s + lane_addr + c * 32, v);
        tmem_ld32(tm_e + lane_addr + c * 32, e);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c * 32 + j < p.T) {
            const float pe = ab_ex2(fmaf(__uint_as_float(v[j]), k2, mk));
            sum += pe;
            dot = fmaf(pe, __uint_as_float(e[j]), dot);
          }
      }
      red[512 + grp * 128 + row] = sum;
      red[1024 + grp * 128 + row] = dot;
      named_bar_sync(1, AB_SM_WARPS * 32);
      const float inv = 1.f / ((red[512 + row] + red[640 + row]) + (red[768 + row] + red[896 + row]));
      const float delta = ((red[1024 + row] + red[1152 + row]) + (red[1280 + row] + red[1408 + row])) * inv;
      if (grp == 0 && r < p.T) { st[r] = mk; st[p.T + r] = inv; st[2 * p.T + r] = delta; }
      const float ps = inv * p.scale;
      for (int c = grp; c < chunks; c += 4) {
        uint32_t v[32], e[32];
        tmem_ld32(tm_s + lane_addr + c * 32, v);
        tmem_ld32(tm_e + lane_addr + c * 32, e);
        tmem_ld_wait();
        uint8_t* tile = smem + off_r1 + (c >> 1) * 16384 + row * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {                             // 8 columns = one 16-byte chunk per plane
          uint32_t ph[4], pl[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = c * 32 + g * 8 + 2 * j;
            float d0 = 0.f, d1 = 0.f;
            if (col < p.T) d0 = ab_ex2(fmaf(__uint_as_float(v[g * 8 + 2 * j]), k2, mk)) * ps * (__uint_as_float(e[g * 8 + 2 * j]) - delta);
            if (col + 1 < p.T) d1 = ab_ex2(fmaf(__uint_as_float(v[g * 8 + 2 * j + 1]), k2, mk)) * ps * (__uint_as_float(e[g * 8 + 2 * j + 1]) - delta);
            split_f16x2(d1, d0, ph[j], pl[j]);
          }
          const int chunk = ((c & 1) * 4 + g) ^ (row & 7);        // SWIZZLE_128B position of this 16-byte chunk
          *reinterpret_cast<uint4*>(tile + (chunk << 4)) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
          *reinterpret_cast<uint4*>(tile + AB_OP_PLANE + (chunk << 4)) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        }
      }
    } else {
      const float* cmk = red;                                     // per query column: -max * k2, 1 / sum, delta
      const float* cinv = red + AB_MAXTK;
      const float* cdl = red + 2 * AB_MAXTK;
      for (int c = grp; c < chunks; c += 4) {
        uint32_t v[32], e[32];
        tmem_ld32(tm_s + lane_addr + c * 32, v);
        tmem_ld32(tm_e + lane_addr + c * 32, e);
        tmem_ld_wait();
        uint8_t* tile = smem + off_r1 + (c >> 1) * 16384 + row * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t pp[4], dd[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = c * 32 + g * 8 + 2 * j;               // query index
            float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
            if (col < p.T) {
              p0 = ab_ex2(fmaf(__uint_as_float(v[g * 8 + 2 * j]), k2, cmk[col])) * cinv[col];
              d0 = p0 * p.scale * (__uint_as_float(e[g * 8 + 2 * j]) - cdl[col]);
            }
            if (col + 1 < p.T) {
              p1 = ab_ex2(fmaf(__uint_as_float(v[g * 8 + 2 * j + 1]), k2, cmk[col + 1])) * cinv[col + 1];
              d1 = p1 * p.scale * (__uint_as_float(e[g * 8 + 2 * j + 1]) - cdl[col + 1]);
            }
            pp[j] = cvt_f16x2(p1, p0);
            dd[j] = cvt_f16x2(d1, d0);
          }
          const int chunk = ((c & 1) * 4 + g) ^ (row & 7);
          *reinterpret_cast<uint4*>(tile + (chunk << 4)) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
          *reinterpret_cast<uint4*>(tile + AB_OP_PLANE + (chunk << 4)) = make_uint4(dd[0], dd[1], dd[2], dd[3]);
        }
      }
    }
    tc_fence_before();
    fence_proxy_async();
    mbar_arrive(bar0 + 16);
    // ---- outputs -> split planes: group g takes the 16 output columns [16g, 16g + 16) ----
    mbar_wait(bar0 + 24, 0);
    tc_fence_after();
    const int nout = PHASE == 0 ? 1 : 2;
    for (int o = 0; o < nout; ++o) {
      // phase A: dQ (qkv column block 0) from TMEM columns 0..63; phase B: dV (block 2) from columns 0..63, dK (block 1) from 256..319
      const int blockcol = PHASE == 0 ? 0 : (o == 0 ? 2 * C1 : C1);
      const uint32_t tm = o == 0 ? tm_s : tm_e;
      uint32_t v[16];
      tmem_ld16(tm + lane_addr + grp * 16, v);
      tmem_ld_wait();
      if (r < p.T) {
        const size_t orow = (size_t)(row0 + r) * (3 * C1) + blockcol + hd * AB_D + grp * 16;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t ph[4], pl[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) split_f16x2(__uint_as_float(v[g * 8 + 2 * j + 1]), __uint_as_float(v[g * 8 + 2 * j]), ph[j], pl[j]);
          *reinterpret_cast<uint4*>(p.d_hi + orow + g * 8) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
          *reinterpret_cast<uint4*>(p.d_lo + orow + g * 8) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AB_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn ab_get_encode() {
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

int make_map(EncodeTiledFn enc, CUtensorMap* m, const uint16_t* base, cuuint64_t cols, cuuint64_t rows, cuuint32_t box_rows) {
  cuuint64_t dims[3] = {cols, rows, 2};
  cuuint64_t strides[2] = {cols * 2, rows * cols * 2};
  cuuint32_t estr[3] = {1, 1, 1};
  cuuint32_t box[3] = {AB_D, box_rows, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<uint16_t*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { b200r_set_error("cuTensorMapEncodeTiled(attention backward) failed: %d", (int)r); return B200R_ECUDA; }
  return B200R_OK;
}

}  // namespace

size_t b200r_attention_bwd_tc_ws(int n, int tokens, int heads) { return (size_t)n * heads * 3 * tokens * sizeof(float); }

// returns B200R_ENOTSUP when the geometry is outside the tensor-core kernel (the caller then takes the CUDA-core kernel)
int b200r_attention_bwd_tc(const uint16_t* qkv, const uint16_t* dout, uint16_t* dqkv, float* stats, int n, int tokens, int heads, float scale,
                           cudaStream_t stream) {
  if (tokens > AB_MAXTK || tokens < 1) return B200R_ENOTSUP;
  EncodeTiledFn enc = ab_get_encode();
  if (!enc) { b200r_set_error("cuTensorMapEncodeTiled entry point not found"); return B200R_ECUDA; }
  const int Tk = (tokens + 31) & ~31, MT = (tokens + 127) / 128;
  const cuuint64_t rows = (cuuint64_t)n * tokens;
  CUtensorMap m_qkv_t, m_qkv_a, m_do_t, m_do_a;
  int rc;
  if ((rc = make_map(enc, &m_qkv_t, qkv, (cuuint64_t)3 * heads * AB_D, rows, 128)) || (rc = make_map(enc, &m_qkv_a, qkv, (cuuint64_t)3 * heads * AB_D, rows, Tk)) ||
      (rc = make_map(enc, &m_do_t, dout, (cuuint64_t)heads * AB_D, rows, 128)) || (rc = make_map(enc, &m_do_a, dout, (cuuint64_t)heads * AB_D, rows, Tk)))
    return rc;
  AttnBwdParams p;
  const size_t cin = (size_t)n * tokens * 3 * heads * AB_D;
  p.d_hi = dqkv; p.d_lo = dqkv + cin;
  p.stats = stats;
  p.T = tokens; p.Tk = Tk; p.H = heads; p.MT = MT;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  const int smem = 2 * Tk * 128 + 2 * AB_OP_PLANE + 64 + 3 * 4 * 128 * 4 + 1024;
  static int configured = 0;
  if (configured < smem) {
    B200R_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    B200R_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  attention_bwd_tc_kernel<0><<<n * heads * MT, AB_THREADS, smem, stream>>>(m_qkv_t, m_qkv_a, m_do_t, m_do_a, p);
  attention_bwd_tc_kernel<1><<<n * heads * MT, AB_THREADS, smem, stream>>>(m_qkv_t, m_qkv_a, m_do_t, m_do_a, p);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}
