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
//     fp16 (R, G, B, 1) pixels (8 bytes each, 4 pixels of padding on both sides).  The 7 taps of a kernel row for output
//     pixel ox are the 8 consecutive pixels starting at padded column 2*ox (the first one meets a zero weight), i.e. the
//     64 bytes at byte offset 16*ox.  In the no-swizzle K-major canonical layout an operand row is addressed as
//     start + (m % 8) * 16 + (m / 8) * SBO + j * LBO  (j = 16-byte K chunk); with SBO = 128 and LBO = 16 that is
//     start + 16 * (m + j): row m of the A operand IS the staged row at byte 16*m.  The windows overlap, the tensor core
//     does not care.  A kernel row is two K = 16 steps straight from the 1.8 KB row buffer.
//   * One staged input row p feeds the 3-4 output rows oy with ky = p - 2*oy in [0, 6].  Their accumulators sit in
//     consecutive 64-column slots of a ring of 8 in TMEM (all 512 columns), and the weights of ky = 6, 4, 2, 0 (even p)
//     / 5, 3, 1 (odd p) are stacked in that order in shared memory: ONE MMA with N = 256 / 192 updates all of them, the
//     A operand is read once.  Only the row that starts (ky = 0, must overwrite) gets its own N = 64 MMA for the first
//     K step, and a ring wrap splits a group in two.  5 MMAs per two input rows instead of 14.
//   * BN scale is folded into the fp16 weights by the caller; the BN bias rides on the tensor core too: the fourth
//     channel of every staged pixel is 1.0 and meets bias_hi / bias_lo (an fp16 pair, ~2^-22 relative) in the centre
//     kernel row.  The epilogue is conversion and max only.
//   * The epilogue keeps the vertical 3-max of the pooling window in registers (thread = output column, rows arrive in
//     order; ReLU commutes with max and is applied once per pooled value), exchanges columns through a staging row for
//     the horizontal 3-max and writes the pooled row with coalesced 16-byte stores.  The 112x112 activation never
//     exists in HBM.
//   * ToTensor + Normalize in the producers: fp16(fma(byte, 1/(255 std), -mean/std)) in fp32 arithmetic, no tables;
//     lanes own interleaved pixel pairs so that the 16-byte shared-memory stores of a warp are contiguous.
//
// Traffic per image: 150 528 B read + 401 408 B written (56*56*64 fp16); tensor work 2*112*112*64*(7*32) = 360 MFLOP
// issued (147/224 useful).  Warp roles (544 threads, one CTA per SM, persistent over quarter-image units):
//   warps 0-7   epilogue: two groups of four (TMEM lane quadrant = warp % 4), group g takes channels [32g, 32g+32)
//   warp  8     TMEM allocation + MMA issue
//   warps 9-16  producers: raw uint8 row -> normalised fp16 row buffer
//
// Split-precision twin (SPLIT = true, the fp32-faithful mode the evaluation defaults to): the same schedule with
//   * an EXACT single-plane A operand: the staged pixel is (R, G, B) / 256 as fp16 (8 significant bits, no rounding) and a
//     fourth channel that is 1 on real pixels and 0 in the padding; x/255, 1/std, the BN scale and a power of two 2^k that
//     centres the fp16 range are folded into the weights on the host in double (b200r_stem_pool_split_prepare), -mean/std and
//     the BN bias ride on the fourth channel of every tap (padding taps then contribute exactly nothing, as zero padding
//     of the NORMALISED image demands);
//   * the weights as an fp16 hi/lo pair (22 bits): TWO MMAs per product instead of the three a two-plane A operand needs;
//   * fp32 pooling: four epilogue groups of 16 channels (25 warps), fp32 vertical / horizontal 3-max, ReLU, 2^-k, and only
//     then the hi/lo split of the pooled value into the two output planes [2][n, h/4, w/4, 64].
#include <cmath>
#include <vector>

#include "sm100_ptx.cuh"

namespace {

constexpr int SP_PROD_WARPS = 8;
constexpr int SP_SLOTS = 16;                 // staged input rows in flight
constexpr int SP_SLOT_BYTES = 2176;          // >= 16 * (127 + 4) = 2096 (the 128-row operand reads past the 112 real columns)
constexpr int SP_ACC = 8;                    // TMEM accumulators (64 columns each)
constexpr int SP_UNIT_POOLED = 14;           // pooled rows per work unit
constexpr int SP_STAGE_PITCH = 96;           // bytes per column in a group's staging row (64 used): columns c and c+2 land
                                             // in different bank halves, so the pooling reads are conflict free
constexpr int SP_STAGE_BYTES = 128 * SP_STAGE_PITCH;
constexpr int SP_OFF_B = 0;                                  // even tile [2 halves][2 chunks][256 rows][16 B] = 16 KB,
constexpr int SP_B_ODD = 16384;                              // then the odd tile [2][2][192][16 B] = 12 KB
constexpr int SP_B_PLANE = 14 * 2048;                        // SPLIT: the lo-plane tiles follow the hi-plane tiles
constexpr int SP_WROW = 7 * 8 * 4;                           // SPLIT: prepared weights per output channel, [ky][t = kx + 1][R, G, B, 1]

// warp roles and shared-memory map of the two precisions
// MODE: 0 = fp16 planes from uint8 pixels, 1 = split planes from uint8 pixels, 2 = split planes + arg-max codes from float32 NCHW
// pixels in [0, 1] (the attack path's saved forward: A operand as an fp16 hi/lo pair, three MMAs per product)
template <int MODE>
struct SpCfg {
  static constexpr bool SPLIT = MODE >= 1, F32IN = MODE == 2;
  static constexpr int EPI_WARPS = SPLIT ? 16 : 8;           // groups of four (TMEM lane quadrant = warp % 4); group g owns
  static constexpr int GROUPS = EPI_WARPS / 4;               // channels [CG g, CG g + CG) of every output row
  static constexpr int CG = 64 / GROUPS;
  static constexpr int MMA_WARP = EPI_WARPS;
  static constexpr int PROD_WARP0 = EPI_WARPS + 1;
  static constexpr int THREADS = (PROD_WARP0 + SP_PROD_WARPS) * 32;   // 544 / 800
  static constexpr int OFF_ROWS = SP_OFF_B + (SPLIT ? 2 : 1) * SP_B_PLANE;
  static constexpr int OFF_ROWS_LO = OFF_ROWS + SP_SLOTS * SP_SLOT_BYTES;                 // F32IN: the lo plane of the staged rows
  static constexpr int OFF_STAGE = OFF_ROWS + (F32IN ? 2 : 1) * SP_SLOTS * SP_SLOT_BYTES;   // [groups][2 buffers][128 columns][96 B]
  static constexpr int OFF_BARS = OFF_STAGE + GROUPS * 2 * SP_STAGE_BYTES;
  static constexpr int SMEM = OFF_BARS + (2 * SP_SLOTS + 2 * SP_ACC) * 8 + 16 + 128;
};

struct StemPoolParams {
  const uint8_t* img;      // [n, h, w, 3]
  const float* img_f32;    // MODE 2: float32 NCHW [n, 3, h, w] in [0, 1]
  uint8_t* codes;          // MODE 2: [n, h/4, w/4, 64] window position ky*3+kx of the pooled maximum, 0xF where it is not positive
  const __half* wgt;       // [64, 192], column = ky*24 + kx*3 + c, BN scale folded in; SPLIT: prepared planes [2][64][224]
  const float* bias;       // [64] (nullable; SPLIT: unused, the bias is in the prepared weights)
  __half* y;               // [n, h/4, w/4, 64]; SPLIT: hi plane, the lo plane follows at y + n*(h/4)*(w/4)*64
  float out_scale;         // SPLIT: 2^-k
  int n, h, w;
  int units_per_img, n_units;
  float k1[3], k0[3];      // normalised = fma(byte, k1, k0)
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

// ---- MMA issue ------------------------------------------------------------------------------------------------------
// Rings are indexed by ABSOLUTE coordinates: input row p sits in row slot p % 16, output row oy accumulates in TMEM slot
// oy % 8.  Everything an interior input row needs is then a compile-time function of I = p % 16, and the issue loop --
// which bounded the kernel when slots, blocks, wraps and descriptors were computed at run time on the uniform datapath
// (650 cycles per input row) -- is a jump into one of 16 straight-line sequences of 4-5 UTCHMMA.
// descriptors as (hi, lo) words: hi = SBO 128 B | version, identical for A and B; lo = start address | LBO; adding to lo
// moves the start address (everything stays below 256 KB, no carry into the LBO bits)
constexpr uint32_t SP_DESC_HI = (128u >> 4) | (1u << 14);
constexpr uint32_t SP_IDESC0 = (1u << 4) | ((uint32_t)(128 >> 4) << 24);   // fp16 A/B, fp32 accumulator, K-major, M = 128
__device__ __forceinline__ uint64_t sp_desc(uint32_t lo) { return ((uint64_t)SP_DESC_HI << 32) | lo; }

struct MmaBases { uint32_t a_lo0, be_lo0, bo_lo0, tmem; };

// All MMAs of input row p with p % 16 == I.  Output rows are counted from the newest possible one: k = 0 is oy_top = p / 2
// (kernel row ky = p % 2), k is oy_top - k (ky = p % 2 + 2k); its accumulator is TMEM slot (oy_top - k) % 8 and its weights
// block BMAX - k of the parity's stacked tile.  Rows [kmin, kmax] are present (interior rows: all of [0, BMAX]; the edges
// of a unit or of the image clip the range).  Called with literal bounds the whole body folds to 4-5 UTCHMMA with
// immediate descriptor offsets.
template <int I, int MODE>
__device__ __forceinline__ void sp_issue(const MmaBases& mb, int kmin, int kmax, bool all_fresh) {
  constexpr bool SPLIT = MODE >= 1, F32IN = MODE == 2;
  constexpr int PAR = I & 1, S0 = (I >> 1) % SP_ACC, BMAX = 3 - PAR;
  constexpr int SEG0_HI = S0 < BMAX ? S0 : BMAX;       // slots fall with k and wrap below 0: [0, SEG0_HI] and [S0 + 1, BMAX] are contiguous
  auto one = [&](int h, int ka, int kb, uint32_t accumulate) {     // one MMA over the rows [ka, kb], K step h
    const uint32_t d = mb.tmem + (uint32_t)((S0 - kb) & (SP_ACC - 1)) * 64u;
    const uint32_t b = (PAR ? mb.bo_lo0 : mb.be_lo0) + (uint32_t)h * (PAR ? (6144u >> 4) : (8192u >> 4)) + (uint32_t)(BMAX - kb) * 64u;
    const uint64_t a = sp_desc(mb.a_lo0 + I * (SP_SLOT_BYTES >> 4) + 2 * h);
    const uint32_t idesc = SP_IDESC0 | (((uint32_t)(kb - ka + 1) * 8u) << 17);
    umma_bf16(d, a, sp_desc(b), idesc, accumulate);
    if constexpr (SPLIT) umma_bf16(d, a, sp_desc(b + (SP_B_PLANE >> 4)), idesc, 1u);      // the weights' lo plane
    if constexpr (F32IN) umma_bf16(d, a + ((SP_SLOTS * SP_SLOT_BYTES) >> 4), sp_desc(b), idesc, 1u);   // the pixels' lo plane x the weights' hi plane
  };
  auto run = [&](int h, int lo, int hi) {
    const int b0 = hi < SEG0_HI ? hi : SEG0_HI;
    if (lo <= b0) one(h, lo, b0, 1u);
    if constexpr (S0 < BMAX) {
      const int a1 = lo > S0 + 1 ? lo : S0 + 1;
      if (a1 <= hi) one(h, a1, hi, 1u);
    }
  };
  if (all_fresh) {                                     // p == 3, top of the image: every row present starts here
    for (int k = kmin; k <= kmax; ++k) { one(0, k, k, 0u); one(1, k, k, 1u); }
    return;
  }
  int lo = kmin;
  if (PAR == 0 && kmin == 0) { one(0, 0, 0, 0u); lo = 1; }   // ky = 0: the row starts here and overwrites its accumulator
  if (lo <= kmax) run(0, lo, kmax);
  run(1, kmin, kmax);
}

template <int MODE>
__global__ void __launch_bounds__(SpCfg<MODE>::THREADS, 1) stem_pool_kernel(const StemPoolParams p) {
  using C = SpCfg<MODE>;
  constexpr bool SPLIT = C::SPLIT, F32IN = C::F32IN;
  constexpr int SP_EPI_WARPS = C::EPI_WARPS, SP_MMA_WARP = C::MMA_WARP, SP_PROD_WARP0 = C::PROD_WARP0, SP_THREADS = C::THREADS;
  constexpr int SP_OFF_ROWS = C::OFF_ROWS, SP_OFF_STAGE = C::OFF_STAGE, SP_OFF_BARS = C::OFF_BARS;
  extern __shared__ __align__(128) uint8_t sp_smem_raw[];    // nothing here needs more than 16-byte alignment (no swizzle)
  uint8_t* smem = sp_smem_raw;
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
    // weights in operand layout.  Tile `par` stacks the kernel rows ky = (6 - par) - 2*jb as 64-row blocks jb; K step h covers
    // window positions t = 4h .. 4h+3 (t = kx + 1), 4 channels each; chunk j = positions 4h+2j, 4h+2j+1.
    __half* sB = reinterpret_cast<__half*>(smem + SP_OFF_B);
    for (int idx0 = threadIdx.x; idx0 < (SPLIT ? 2 : 1) * 14 * 1024; idx0 += SP_THREADS) {
      const int plane = idx0 >= 14 * 1024, idx = idx0 - plane * 14 * 1024;
      const int par = idx >= 8192, li = idx - par * 8192, rows = par ? 192 : 256;
      const int e = li & 7, r = (li >> 3) % rows, jh = (li >> 3) / rows;       // jh = h * 2 + j
      const int j = jh & 1, h = jh >> 1, jb = r >> 6, co = r & 63;
      const int ky = (6 - par) - 2 * jb, t = 4 * h + 2 * j + (e >> 2), c4 = e & 3;
      __half val = __float2half_rn(0.f);
      if constexpr (SPLIT) {
        val = p.wgt[(plane * 64 + co) * SP_WROW + ky * 32 + t * 4 + c4];
      } else if (c4 < 3) {
        if (t >= 1) val = p.wgt[co * 192 + ky * 24 + (t - 1) * 3 + c4];
      } else if (ky == 3 && p.bias) {
        const float bf = p.bias[co];
        const __half bh = __float2half_rn(bf);
        if (t == 4) val = bh;
        else if (t == 3) val = __float2half_rn(bf - __half2float(bh));
      }
      sB[idx0] = val;
    }
    // row ring: zero, with the constant-one fourth channel on every pixel column (padding included; SPLIT: the padding
    // stays all-zero, the producers write the one with every real pixel)
    uint32_t* rows = reinterpret_cast<uint32_t*>(smem + SP_OFF_ROWS);
    for (int idx = threadIdx.x; idx < (F32IN ? 2 : 1) * SP_SLOTS * SP_SLOT_BYTES / 4; idx += SP_THREADS) {
      const int w4 = idx % (SP_SLOT_BYTES / 4);
      rows[idx] = (!SPLIT && (w4 & 1) && w4 < 2 * (W + 8)) ? 0x3C000000u : 0u;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (F32IN && warp >= SP_PROD_WARP0) {
    // ================================ producers, float32 NCHW pixels in [0, 1] ================================
    // lane owns the pixel pairs q = lane + 32 j: three coalesced 8-byte loads per pair (one per channel plane); the pixel becomes an
    // fp16 hi/lo pair -- (R, G, B, 1) into the hi ring, (R_lo, G_lo, B_lo, 0) into the lo ring, 16-byte stores in 512-byte runs
    const int pw = warp - SP_PROD_WARP0;
    float2 cur[12];
    int cur_cnt = -1;
    uint32_t ph_empty = 0;
    auto commit_row = [&](const float2 (&v)[12], uint32_t slot) {
      mbar_wait(empty_in(slot), ((ph_empty >> slot) & 1) ^ 1);
      ph_empty ^= 1u << slot;
      const uint32_t dst = sbase + SP_OFF_ROWS + slot * SP_SLOT_BYTES + 32;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = lane + 32 * j;
        if (q < WO) {
          const float2 r = v[3 * j], g = v[3 * j + 1], b = v[3 * j + 2];
          uint4 oh, ol;
          split_f16x2(g.x, r.x, oh.x, ol.x);
          split_f16x2(1.0f, b.x, oh.y, ol.y);             // fourth channel: 1 in the hi plane, 0 in the lo plane
          split_f16x2(g.y, r.y, oh.z, ol.z);
          split_f16x2(1.0f, b.y, oh.w, ol.w);
          sts_v4(dst + q * 16, oh);
          sts_v4(dst + SP_SLOTS * SP_SLOT_BYTES + q * 16, ol);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_in(slot));
    };
    const size_t plane = (size_t)H * W;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const Unit t = make_unit(p, u);
      for (int pp = t.p_lo + ((pw - t.p_lo) & (SP_PROD_WARPS - 1)); pp <= t.p_hi; pp += SP_PROD_WARPS) {   // rows with pp % 8 == pw
        float2 nxt[12];
        const float* rowp = p.img_f32 + (size_t)t.n * 3 * plane + (size_t)(pp - 3) * W;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int q = lane + 32 * j;
#pragma unroll
          for (int c = 0; c < 3; ++c) nxt[3 * j + c] = q < WO ? __ldg(reinterpret_cast<const float2*>(rowp + c * plane) + q) : make_float2(0.f, 0.f);
        }
        if (cur_cnt >= 0) commit_row(cur, (uint32_t)cur_cnt);
#pragma unroll
        for (int j = 0; j < 12; ++j) cur[j] = nxt[j];
        cur_cnt = pp % SP_SLOTS;
      }
    }
    if (cur_cnt >= 0) commit_row(cur, (uint32_t)cur_cnt);
  } else if (warp >= SP_PROD_WARP0) {
    // ================================ producers ================================
    const int pw = warp - SP_PROD_WARP0;
    // lane owns the pixel pairs q = lane + 32 j: the warp's four 16-byte stores per row are contiguous 512-byte runs.
    // A pair is 6 bytes at offset 6q: two aligned words from 6q & ~3, shifted by 16 bits when q is odd.
    uint32_t cur[8];
    int cur_cnt = -1;                              // row slot of the staged-in-registers row
    const float k1r = p.k1[0], k1g = p.k1[1], k1b = p.k1[2], k0r = p.k0[0], k0g = p.k0[1], k0b = p.k0[2];
    auto nrm = [](uint32_t w, uint32_t sel, float k1, float k0) {
      return fmaf(__uint_as_float(__byte_perm(w, 0x4B000000u, sel)) - 8388608.0f, k1, k0);
    };
    uint32_t ph_empty = 0;                         // per-slot phase bits (a slot is always filled by the same warp)
    auto commit_row = [&](const uint32_t (&v)[8], uint32_t slot) {
      mbar_wait(empty_in(slot), ((ph_empty >> slot) & 1) ^ 1);
      ph_empty ^= 1u << slot;
      const uint32_t dst = sbase + SP_OFF_ROWS + slot * SP_SLOT_BYTES + 32;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = lane + 32 * j;
        if (q < WO) {
          const uint32_t sh = (lane & 1) * 16;
          const uint32_t lo = __funnelshift_r(v[2 * j], v[2 * j + 1], sh), hi = v[2 * j + 1] >> sh;   // r0 g0 b0 r1 | g1 b1
          uint4 o;
          if constexpr (SPLIT) {
            // byte b under the exponent byte 0x64 is the fp16 number 1024 + b; (1024 + b) / 256 - 4 = b / 256 exactly.
            // The fourth channel: 0x6500 = 1280 -> 1280 / 256 - 4 = 1.
            const uint32_t p1 = __byte_perm(lo, hi, 0x0543);                         // r1 g1 b1 x
            const uint32_t kmul = 0x1C001C00u /* 2^-8 */, kadd = 0xC400C400u /* -4 */, kexp = 0x65000064u;
            auto cv = [&](uint32_t w) {
              const __half2 r = __hfma2(*reinterpret_cast<const __half2*>(&w), *reinterpret_cast<const __half2*>(&kmul), *reinterpret_cast<const __half2*>(&kadd));
              return *reinterpret_cast<const uint32_t*>(&r);
            };
            o.x = cv(__byte_perm(lo, kexp, 0x4140));
            o.y = cv(__byte_perm(lo, kexp, 0x7542));
            o.z = cv(__byte_perm(p1, kexp, 0x4140));
            o.w = cv(__byte_perm(p1, kexp, 0x7542));
          } else {
            o.x = cvt_f16x2(nrm(lo, 0x7651, k1g, k0g), nrm(lo, 0x7650, k1r, k0r));
            o.y = cvt_f16x2(1.0f, nrm(lo, 0x7652, k1b, k0b));
            o.z = cvt_f16x2(nrm(hi, 0x7650, k1g, k0g), nrm(lo, 0x7653, k1r, k0r));
            o.w = cvt_f16x2(1.0f, nrm(hi, 0x7651, k1b, k0b));
          }
          sts_v4(dst + q * 16, o);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_in(slot));
    };
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const Unit t = make_unit(p, u);
      for (int pp = t.p_lo + ((pw - t.p_lo) & (SP_PROD_WARPS - 1)); pp <= t.p_hi; pp += SP_PROD_WARPS) {   // rows with pp % 8 == pw
        uint32_t nxt[8];
        const uint8_t* rowp = p.img + ((size_t)t.n * H + (pp - 3)) * (size_t)W * 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int q = lane + 32 * j;
          nxt[2 * j] = 0u; nxt[2 * j + 1] = 0u;
          if (q < WO) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(rowp + ((6 * q) & ~3));
            nxt[2 * j] = __ldg(src); nxt[2 * j + 1] = __ldg(src + 1);
          }
        }
        if (cur_cnt >= 0) commit_row(cur, (uint32_t)cur_cnt);
#pragma unroll
        for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
        cur_cnt = pp % SP_SLOTS;
      }
    }
    if (cur_cnt >= 0) commit_row(cur, (uint32_t)cur_cnt);
  } else if (warp == SP_MMA_WARP) {
    // ================================ MMA issue ================================
    MmaBases mb;
    mb.a_lo0 = ((sbase + SP_OFF_ROWS) >> 4) | ((16u >> 4) << 16);
    mb.be_lo0 = ((sbase + SP_OFF_B) >> 4) | ((4096u >> 4) << 16);              // even tile: LBO = 256 rows * 16 B
    mb.bo_lo0 = ((sbase + SP_OFF_B + SP_B_ODD) >> 4) | ((3072u >> 4) << 16);   // odd tile: 192 rows
    mb.tmem = tmem_base;
    uint32_t ph_full = 0, ph_aempty = 0;           // per-slot phase bits
    auto take_acc = [&](uint32_t as) { mbar_wait(acc_empty(as), ((ph_aempty >> as) & 1) ^ 1); ph_aempty ^= 1u << as; };
    auto take_row = [&](uint32_t slot) { mbar_wait(full_in(slot), (ph_full >> slot) & 1); ph_full ^= 1u << slot; };
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const Unit t = make_unit(p, u);
      // edge rows: the same schedule with run-time row ranges
      auto edge = [&](int pp) {
        const uint32_t slot = (uint32_t)pp % SP_SLOTS;
        const int oy_top = pp >> 1, par = pp & 1;
        const int kmin = max(0, oy_top - t.r_last), kmax = min(3 - par, oy_top - t.r_first);
        const bool all_fresh = pp == 3;
        if (all_fresh) { for (int k = kmin; k <= kmax; ++k) take_acc((uint32_t)(oy_top - k) % SP_ACC); }
        else if (!par && kmin == 0) take_acc((uint32_t)oy_top % SP_ACC);
        take_row(slot);
        tc_fence_after();
        if (elect_one()) {
          switch (slot) {
            case 0: sp_issue<0, MODE>(mb, kmin, kmax, all_fresh); break;   case 1: sp_issue<1, MODE>(mb, kmin, kmax, all_fresh); break;
            case 2: sp_issue<2, MODE>(mb, kmin, kmax, all_fresh); break;   case 3: sp_issue<3, MODE>(mb, kmin, kmax, all_fresh); break;
            case 4: sp_issue<4, MODE>(mb, kmin, kmax, all_fresh); break;   case 5: sp_issue<5, MODE>(mb, kmin, kmax, all_fresh); break;
            case 6: sp_issue<6, MODE>(mb, kmin, kmax, all_fresh); break;   case 7: sp_issue<7, MODE>(mb, kmin, kmax, all_fresh); break;
            case 8: sp_issue<8, MODE>(mb, kmin, kmax, all_fresh); break;   case 9: sp_issue<9, MODE>(mb, kmin, kmax, all_fresh); break;
            case 10: sp_issue<10, MODE>(mb, kmin, kmax, all_fresh); break; case 11: sp_issue<11, MODE>(mb, kmin, kmax, all_fresh); break;
            case 12: sp_issue<12, MODE>(mb, kmin, kmax, all_fresh); break; case 13: sp_issue<13, MODE>(mb, kmin, kmax, all_fresh); break;
            case 14: sp_issue<14, MODE>(mb, kmin, kmax, all_fresh); break; default: sp_issue<15, MODE>(mb, kmin, kmax, all_fresh); break;
          }
          umma_commit(empty_in(slot));
          // output rows whose last contributing input row this is: ky = 6, or the bottom image row (two taps in the padding)
          if (!par && kmin <= 3 && 3 <= kmax) umma_commit(acc_full((uint32_t)(oy_top - 3) % SP_ACC));
          if (pp == H + 2 && kmin <= 2 && 2 <= kmax) umma_commit(acc_full((uint32_t)(oy_top - 2) % SP_ACC));
        }
        __syncwarp();
      };
      // interior rows [pp_a, pp_b]: all 4 / 3 output rows present
      const int pp_a = max(t.p_lo, 2 * (t.r_first + 3)), pp_b = min(t.p_hi, 2 * t.r_last + 1);
      int pp = t.p_lo;
      for (; pp <= t.p_hi && pp < pp_a; ++pp) edge(pp);
      if (pp <= pp_b) {
        // straight-line sequence of the 16 ring positions, entered at pp % 16 (Duff's device): slots, phases' bit positions,
        // descriptors and commit targets are immediates
#define SP_STEP(I)                                                                           \
        case I: {                                                                            \
          if (pp > pp_b) break;                                                              \
          if (((I) & 1) == 0) take_acc(((I) >> 1) % SP_ACC);                                 \
          take_row(I);                                                                       \
          tc_fence_after();                                                                  \
          if (elect_one()) {                                                                 \
            sp_issue<I, MODE>(mb, 0, 3 - ((I) & 1), false);                                        \
            umma_commit(empty_in(I));                                                        \
            if (((I) & 1) == 0) umma_commit(acc_full((((I) >> 1) + SP_ACC - 3) % SP_ACC));   \
          }                                                                                  \
          __syncwarp();                                                                      \
          ++pp;                                                                              \
        }
        bool more = true;
        switch (pp & 15) {
          do {
            SP_STEP(0) SP_STEP(1) SP_STEP(2) SP_STEP(3) SP_STEP(4) SP_STEP(5) SP_STEP(6) SP_STEP(7)
            SP_STEP(8) SP_STEP(9) SP_STEP(10) SP_STEP(11) SP_STEP(12) SP_STEP(13) SP_STEP(14)
            case 15: {
              if (pp > pp_b) { more = false; break; }
              take_row(15);
              tc_fence_after();
              if (elect_one()) { sp_issue<15, MODE>(mb, 0, 2, false); umma_commit(empty_in(15)); }
              __syncwarp();
              ++pp;
            }
          } while (more && pp <= pp_b);
        }
#undef SP_STEP
      }
      for (; pp <= t.p_hi; ++pp) edge(pp);
    }
  } else if constexpr (F32IN) {
    // ================================ epilogue, fp32 pooling + arg-max codes ================================
    // as the SPLIT epilogue, and every maximum carries the kernel row / column it came from: the vertical pass keeps ky (two bits per
    // channel in one word), the horizontal pass picks, among the columns that tie for the maximum, the smallest ky*3+kx -- the first
    // maximum in the window's scan order, as max_pool2d's backward routes it.  A maximum that is not positive gets code 0xF: the ReLU
    // that follows has derivative 0 there (the backward of pool and ReLU in one code).
    const int grp = warp >> 2, m = threadIdx.x & 127;
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + grp * 16;
    const uint32_t stage0 = sbase + SP_OFF_STAGE + grp * 2 * SP_STAGE_BYTES;
    const size_t plane_stride = (size_t)p.n * PH * PW * 64;
    const float osc = p.out_scale;
    uint32_t ph_afull = 0, emit = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const Unit t = make_unit(p, u);
      float prev_odd[16], v[16];
      uint32_t kyw = 0;                                    // ky of v[i] in bits 2i, 2i+1
#pragma unroll
      for (int i = 0; i < 16; ++i) { prev_odd[i] = 0.f; v[i] = 0.f; }
      for (int r = t.r_first; r <= t.r_last; ++r) {
        const uint32_t aslot = (uint32_t)r % SP_ACC;
        mbar_wait(acc_full(aslot), (ph_afull >> aslot) & 1);
        ph_afull ^= 1u << aslot;
        tc_fence_after();
        uint32_t acc[16];
        tmem_ld16(lane_base + aslot * 64, acc);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(aslot));
        const bool odd = r & 1;
        if (odd) {                                         // conv row 2 py + 1: kernel row 2 of pooled row py, row 0 of pooled row py + 1
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float cur = __uint_as_float(acc[i]);
            if (cur > v[i]) { v[i] = cur; kyw = (kyw & ~(3u << (2 * i))) | (2u << (2 * i)); }
            prev_odd[i] = cur;
          }
        } else {                                           // conv row 2 py: kernel row 1; the previous odd row is kernel row 0
          kyw = 0;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float cur = __uint_as_float(acc[i]);
            v[i] = prev_odd[i];
            if (cur > v[i]) { v[i] = cur; kyw |= 1u << (2 * i); }
          }
        }
        if (odd && (r >> 1) >= t.py0) {
          const uint32_t st = stage0 + (emit & 1) * SP_STAGE_BYTES;
          if (m < WO) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
              sts_v4(st + m * SP_STAGE_PITCH + ch * 16, make_uint4(__float_as_uint(v[4 * ch]), __float_as_uint(v[4 * ch + 1]),
                                                                 __float_as_uint(v[4 * ch + 2]), __float_as_uint(v[4 * ch + 3])));
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(st + m * SP_STAGE_PITCH + 64), "r"(kyw) : "memory");
          }
          named_bar_sync(1 + grp, 128);
          const size_t prow = ((size_t)(t.n * PH + (r >> 1)) * PW) * 64 + grp * 16;
          __half* yrow = p.y + prow;
          for (int i = m; i < PW * 4; i += 128) {
            const int px = i >> 2, ch = i & 3, c1 = 2 * px;
            const uint32_t at = st + c1 * SP_STAGE_PITCH + ch * 16;
            const uint4 a = lds_v4(at), b = lds_v4(at + SP_STAGE_PITCH);
            uint32_t ka, kb, kl = 0;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ka) : "r"(st + c1 * SP_STAGE_PITCH + 64));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(kb) : "r"(st + (c1 + 1) * SP_STAGE_PITCH + 64));
            uint4 l = make_uint4(0, 0, 0, 0);
            if (px > 0) {
              l = lds_v4(at - SP_STAGE_PITCH);
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(kl) : "r"(st + (c1 - 1) * SP_STAGE_PITCH + 64));
            }
            const float fa[4] = {__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w)};
            const float fb[4] = {__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(b.w)};
            const float fl[4] = {__uint_as_float(l.x), __uint_as_float(l.y), __uint_as_float(l.z), __uint_as_float(l.w)};
            float f[4];
            uint32_t cw = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int cidx = 4 * ch + e;
              // scan order: (ky, kx) ascending; columns kx = 0 (left), 1 (centre), 2 (right)
              float best = fa[e];
              uint32_t code = ((ka >> (2 * cidx)) & 3u) * 3u + 1u;
              if (px > 0) {
                const uint32_t cl = ((kl >> (2 * cidx)) & 3u) * 3u;
                if (fl[e] > best || (fl[e] == best && cl < code)) { best = fl[e]; code = cl; }
              }
              const uint32_t cr = ((kb >> (2 * cidx)) & 3u) * 3u + 2u;
              if (fb[e] > best || (fb[e] == best && cr < code)) { best = fb[e]; code = cr; }
              if (!(best > 0.f)) code = 0xFu;
              cw |= code << (8 * e);
              f[e] = fmaxf(best, 0.f) * osc;
            }
            uint2 oh, ol;
            split_f16x2(f[1], f[0], oh.x, ol.x);
            split_f16x2(f[3], f[2], oh.y, ol.y);
            __half* dst = yrow + px * 64 + ch * 4;
            *reinterpret_cast<uint2*>(dst) = oh;
            *reinterpret_cast<uint2*>(dst + plane_stride) = ol;
            *reinterpret_cast<uint32_t*>(p.codes + prow + px * 64 + ch * 4) = cw;
          }
          ++emit;
        }
      }
    }
  } else if constexpr (SPLIT) {
    // ================================ epilogue, fp32 pooling, hi/lo planes out ================================
    // group g = warp / 4 owns channels [16g, 16g + 16); thread = output column (TMEM lane = 32 * (warp % 4) + lane)
    const int grp = warp >> 2, m = threadIdx.x & 127;
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + grp * 16;
    const uint32_t stage0 = sbase + SP_OFF_STAGE + grp * 2 * SP_STAGE_BYTES;
    const size_t plane_stride = (size_t)p.n * PH * PW * 64;
    const float osc = p.out_scale;
    uint32_t ph_afull = 0, emit = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const Unit t = make_unit(p, u);
      float prev_odd[16], v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { prev_odd[i] = 0.f; v[i] = 0.f; }
      for (int r = t.r_first; r <= t.r_last; ++r) {
        const uint32_t aslot = (uint32_t)r % SP_ACC;
        mbar_wait(acc_full(aslot), (ph_afull >> aslot) & 1);
        ph_afull ^= 1u << aslot;
        tc_fence_after();
        uint32_t acc[16];
        tmem_ld16(lane_base + aslot * 64, acc);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(aslot));
        const bool odd = r & 1;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float cur = __uint_as_float(acc[i]);
          if (odd) { v[i] = fmaxf(v[i], cur); prev_odd[i] = cur; }
          else v[i] = fmaxf(prev_odd[i], cur);
        }
        if (odd && (r >> 1) >= t.py0) {
          // vertical 3-max of pooled row r/2 is in v (0 stands in for the padding row: the ReLU follows)
          const uint32_t st = stage0 + (emit & 1) * SP_STAGE_BYTES;
          if (m < WO) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
              sts_v4(st + m * SP_STAGE_PITCH + ch * 16, make_uint4(__float_as_uint(v[4 * ch]), __float_as_uint(v[4 * ch + 1]),
                                                                 __float_as_uint(v[4 * ch + 2]), __float_as_uint(v[4 * ch + 3])));
          }
          named_bar_sync(1 + grp, 128);
          __half* yrow = p.y + ((size_t)(t.n * PH + (r >> 1)) * PW) * 64 + grp * 16;
          for (int i = m; i < PW * 4; i += 128) {
            const int px = i >> 2, ch = i & 3, c1 = 2 * px;
            const uint32_t at = st + c1 * SP_STAGE_PITCH + ch * 16;
            const uint4 a = lds_v4(at), b = lds_v4(at + SP_STAGE_PITCH);
            float f0 = fmaxf(__uint_as_float(a.x), __uint_as_float(b.x)), f1 = fmaxf(__uint_as_float(a.y), __uint_as_float(b.y));
            float f2 = fmaxf(__uint_as_float(a.z), __uint_as_float(b.z)), f3 = fmaxf(__uint_as_float(a.w), __uint_as_float(b.w));
            if (px > 0) {
              const uint4 l = lds_v4(at - SP_STAGE_PITCH);
              f0 = fmaxf(f0, __uint_as_float(l.x)); f1 = fmaxf(f1, __uint_as_float(l.y));
              f2 = fmaxf(f2, __uint_as_float(l.z)); f3 = fmaxf(f3, __uint_as_float(l.w));
            }
            f0 = fmaxf(f0, 0.f) * osc; f1 = fmaxf(f1, 0.f) * osc; f2 = fmaxf(f2, 0.f) * osc; f3 = fmaxf(f3, 0.f) * osc;
            uint2 oh, ol;
            split_f16x2(f1, f0, oh.x, ol.x);
            split_f16x2(f3, f2, oh.y, ol.y);
            __half* dst = yrow + px * 64 + ch * 4;
            *reinterpret_cast<uint2*>(dst) = oh;
            *reinterpret_cast<uint2*>(dst + plane_stride) = ol;
          }
          ++emit;
        }
      }
    }
  } else {
    // ================================ epilogue ================================
    // group g = warp / 4 owns channels [32g, 32g + 32); thread = output column (TMEM lane = 32 * (warp % 4) + lane)
    const int grp = warp >> 2, m = threadIdx.x & 127;
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + grp * 32;
    const uint32_t stage0 = sbase + SP_OFF_STAGE + grp * 2 * SP_STAGE_BYTES;
    uint32_t ph_afull = 0, emit = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const Unit t = make_unit(p, u);
      uint32_t prev_odd[16], v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { prev_odd[i] = 0u; v[i] = 0u; }
      for (int r = t.r_first; r <= t.r_last; ++r) {
        const uint32_t aslot = (uint32_t)r % SP_ACC;
        mbar_wait(acc_full(aslot), (ph_afull >> aslot) & 1);
        ph_afull ^= 1u << aslot;
        tc_fence_after();
        uint32_t acc[32];
        tmem_ld32(lane_base + aslot * 64, acc);
        tmem_ld_wait();
        tc_fence_before();               // accumulator drained: hand it back before the arithmetic
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(aslot));
        const bool odd = r & 1;
        // scale and bias came in through the MMA; the ReLU commutes with the max and is applied once per pooled value
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t cur = cvt_f16x2(__uint_as_float(acc[2 * i + 1]), __uint_as_float(acc[2 * i]));
          if (odd) { v[i] = hmax2_u32(v[i], cur); prev_odd[i] = cur; }
          else v[i] = hmax2_u32(prev_odd[i], cur);
        }
        if (odd && (r >> 1) >= t.py0) {
          // vertical 3-max of pooled row r/2 is in v (0 stands in for the padding row: the ReLU follows): exchange columns
          // through the group's staging row, then the horizontal 3-max, ReLU, store
          const uint32_t st = stage0 + (emit & 1) * SP_STAGE_BYTES;
          if (m < WO) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
              sts_v4(st + m * SP_STAGE_PITCH + ch * 16, make_uint4(v[4 * ch], v[4 * ch + 1], v[4 * ch + 2], v[4 * ch + 3]));
          }
          named_bar_sync(1 + grp, 128);
          uint4* yrow = reinterpret_cast<uint4*>(p.y + ((size_t)(t.n * PH + (r >> 1)) * PW) * 64 + grp * 32);
          for (int i = m; i < PW * 4; i += 128) {
            const int px = i >> 2, ch = i & 3, c1 = 2 * px;
            const uint32_t at = st + c1 * SP_STAGE_PITCH + ch * 16;
            uint4 a = lds_v4(at);
            const uint4 b = lds_v4(at + SP_STAGE_PITCH);
            a.x = hmax2_u32(a.x, b.x); a.y = hmax2_u32(a.y, b.y); a.z = hmax2_u32(a.z, b.z); a.w = hmax2_u32(a.w, b.w);
            if (px > 0) {
              const uint4 l = lds_v4(at - SP_STAGE_PITCH);
              a.x = hmax2_u32(a.x, l.x); a.y = hmax2_u32(a.y, l.y); a.z = hmax2_u32(a.z, l.z); a.w = hmax2_u32(a.w, l.w);
            }
            a.x = hmax2_u32(a.x, 0u); a.y = hmax2_u32(a.y, 0u); a.z = hmax2_u32(a.z, 0u); a.w = hmax2_u32(a.w, 0u);
            yrow[px * 8 + ch] = a;
          }
          ++emit;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == SP_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, SP_ACC * 64);
  }
}

template <int MODE>
int launch_stem_pool(StemPoolParams& p, const float* mean_host, const float* std_host, b200r_stream_t stream) {
  p.units_per_img = ((p.h >> 2) + SP_UNIT_POOLED - 1) / SP_UNIT_POOLED;
  p.n_units = p.n * p.units_per_img;
  for (int c = 0; c < 3; ++c) {
    p.k1[c] = (float)(1.0 / (255.0 * (double)std_host[c]));
    p.k0[c] = (float)(-(double)mean_host[c] / (double)std_host[c]);
  }
  static bool attr_set[16] = {};
  int dev = 0;
  B200R_CUDA(cudaGetDevice(&dev));
  if (dev < 16 && !attr_set[dev]) {
    B200R_CUDA(cudaFuncSetAttribute(stem_pool_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SpCfg<MODE>::SMEM));
    attr_set[dev] = true;
  }
  const int grid = p.n_units < b200r_num_sms() ? p.n_units : b200r_num_sms();
  stem_pool_kernel<MODE><<<grid, SpCfg<MODE>::THREADS, SpCfg<MODE>::SMEM, as_stream(stream)>>>(p);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

}  // namespace

extern "C" {

int b200r_stem_pool_split_prepare(const float* conv1_w, const float* bn_scale, const float* bn_bias, const float* mean_host,
                                  const float* std_host, int f32_input, uint16_t* planes_host, float* out_scale) {
  B200R_CHECK_ARG(conv1_w && mean_host && std_host && planes_host && out_scale, "null pointer");
  // value of tap (co, ky, t = kx + 1, c): colour channels meet pixel / 256, the fourth channel meets 1 on real pixels
  std::vector<double> v((size_t)64 * SP_WROW, 0.0);
  double amax = 0.0;
  for (int co = 0; co < 64; ++co) {
    const double s = bn_scale ? (double)bn_scale[co] : 1.0;
    for (int ky = 0; ky < 7; ++ky)
      for (int kx = 0; kx < 7; ++kx) {
        double* q = &v[(size_t)co * SP_WROW + ky * 32 + (kx + 1) * 4];
        double m = 0.0;
        for (int c = 0; c < 3; ++c) {
          const double w = (double)conv1_w[((co * 3 + c) * 7 + ky) * 7 + kx] * s / (double)std_host[c];
          q[c] = f32_input ? w : w * (256.0 / 255.0);        // the colour channels meet x in [0, 1] (float input) or byte / 256
          m -= w * (double)mean_host[c];
        }
        if (ky == 3 && kx == 3 && bn_bias) m += (double)bn_bias[co];
        q[3] = m;
        for (int c = 0; c < 4; ++c) amax = fmax(amax, fabs(q[c]));
      }
  }
  B200R_CHECK_ARG(std::isfinite(amax), "non-finite stem weights");
  // 2^k puts the largest entry in [2^13, 2^14): the lo plane of every entry down to 2^-13 of it stays a normal fp16 number
  int k = 0;
  if (amax > 0.0) {
    int e = 0;
    frexp(amax, &e);                  // amax = f * 2^e, f in [0.5, 1)
    k = 14 - e;
    if (k > 60) k = 60;
    if (k < -60) k = -60;
  }
  const double up = ldexp(1.0, k);
  for (size_t i = 0; i < v.size(); ++i) {
    const double x = v[i] * up;
    const __half hi = __float2half_rn((float)x);
    const __half lo = __float2half_rn((float)(x - (double)__half2float(hi)));
    planes_host[i] = *reinterpret_cast<const uint16_t*>(&hi);
    planes_host[v.size() + i] = *reinterpret_cast<const uint16_t*>(&lo);
  }
  *out_scale = (float)ldexp(1.0, -k);
  return B200R_OK;
}

int b200r_stem_pool_u8_split(const uint8_t* img, const uint16_t* wplanes, float out_scale, uint16_t* y, int n, int h, int w,
                             b200r_stream_t stream) {
  B200R_CHECK_ARG(img && wplanes && y, "null pointer");
  B200R_CHECK_ARG(n > 0 && h >= 8 && w >= 8, "bad shape %d x %d x %d", n, h, w);
  B200R_CHECK_ARG(h % 4 == 0 && w % 8 == 0 && w <= 248, "stem_pool needs h %% 4 == 0, w %% 8 == 0, w <= 248 (got %d x %d)", h, w);
  B200R_CHECK_ARG((reinterpret_cast<uintptr_t>(img) & 7) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, "img must be 8-byte, y 16-byte aligned");
  B200R_CHECK_ARG(out_scale > 0.f, "out_scale comes from b200r_stem_pool_split_prepare");
  StemPoolParams p;
  p.img = img; p.img_f32 = nullptr; p.codes = nullptr; p.wgt = reinterpret_cast<const __half*>(wplanes); p.bias = nullptr; p.y = reinterpret_cast<__half*>(y);
  p.n = n; p.h = h; p.w = w; p.out_scale = out_scale;
  const float unit[3] = {1.f, 1.f, 1.f}, zero[3] = {0.f, 0.f, 0.f};
  return launch_stem_pool<1>(p, zero, unit, stream);
}

int b200r_stem_pool_f32_split(const float* x01, const uint16_t* wplanes, float out_scale, uint16_t* y, uint8_t* codes, int n, int h,
                              int w, b200r_stream_t stream) {
  B200R_CHECK_ARG(x01 && wplanes && y && codes, "null pointer");
  B200R_CHECK_ARG(n > 0 && h >= 8 && w >= 8, "bad shape %d x %d x %d", n, h, w);
  B200R_CHECK_ARG(h % 4 == 0 && w % 8 == 0 && w <= 248, "stem_pool needs h %% 4 == 0, w %% 8 == 0, w <= 248 (got %d x %d)", h, w);
  B200R_CHECK_ARG((reinterpret_cast<uintptr_t>(x01) & 7) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(codes) & 3) == 0,
                  "x01 must be 8-byte, y 16-byte, codes 4-byte aligned");
  B200R_CHECK_ARG(out_scale > 0.f, "out_scale comes from b200r_stem_pool_split_prepare");
  StemPoolParams p;
  p.img = nullptr; p.img_f32 = x01; p.codes = codes; p.wgt = reinterpret_cast<const __half*>(wplanes); p.bias = nullptr;
  p.y = reinterpret_cast<__half*>(y);
  p.n = n; p.h = h; p.w = w; p.out_scale = out_scale;
  const float unit[3] = {1.f, 1.f, 1.f}, zero[3] = {0.f, 0.f, 0.f};
  return launch_stem_pool<2>(p, zero, unit, stream);
}

int b200r_stem_pool_u8_f16(const uint8_t* img, const uint16_t* wgt, const float* bias, uint16_t* y,
                           int n, int h, int w, const float* mean_host, const float* std_host, b200r_stream_t stream) {
  B200R_CHECK_ARG(img && wgt && y && mean_host && std_host, "null pointer");
  B200R_CHECK_ARG(n > 0 && h >= 8 && w >= 8, "bad shape %d x %d x %d", n, h, w);
  B200R_CHECK_ARG(h % 4 == 0 && w % 8 == 0 && w <= 248, "stem_pool needs h %% 4 == 0, w %% 8 == 0, w <= 248 (got %d x %d)", h, w);
  B200R_CHECK_ARG((reinterpret_cast<uintptr_t>(img) & 7) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, "img must be 8-byte, y 16-byte aligned");
  StemPoolParams p;
  p.img = img; p.img_f32 = nullptr; p.codes = nullptr; p.wgt = reinterpret_cast<const __half*>(wgt); p.bias = bias; p.y = reinterpret_cast<__half*>(y);
  p.n = n; p.h = h; p.w = w; p.out_scale = 1.f;
  return launch_stem_pool<0>(p, mean_host, std_host, stream);
}

}  // extern "C"
