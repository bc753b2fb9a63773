// Dense contractions on the 5th-generation tensor cores (sm_100a): one persistent, warp-specialised
// kernel serves nn.Linear and NHWC convolutions (implicit GEMM, no im2col in HBM).
//
//   D[m, n] = act( (sum_taps sum_c A[pixel(m) shifted by tap, c] * W[n, tap, c]) * scale[n] + bias[n] + res[m, n] )
//
// Precision: "split-bf16".  Every fp32 tensor lives in HBM as two bf16 planes (hi, lo); a product
// is hi*hi + hi*lo + lo*hi (passes = 3) accumulated in fp32 in TMEM -> ~2^-16 relative error per
// product, which is what holds the 1e-3 logit tolerance against the fp32 reference.  passes = 1 is
// plain bf16.  passes = B200R_PASSES_F16 (16): ONE fp16 plane per tensor (kernel template F16) -- operands and stored
// activations are fp16 (10 explicit mantissa bits, like TF32), one MMA per product, fp32 accumulation; half the
// activation bytes and a third of the tensor-pipe work of the split scheme.  Holds the 1e-3 logit tolerance on the
// ResNet golden models with a 5x margin (tests/test_resnet_gpu.py); the smem ring is twice as deep (no lo tiles).
//
// Structure (one CTA per SM, 320 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor (5-D map over [plane, N, H, W, C] for activations --
//               the tap shift is a coordinate offset, padding is TMA out-of-bounds zero fill, stride is
//               the map's elementStrides -- and a 3-D map over [plane, Cout, K] for weights) into a
//               3-stage ring of 128B-swizzled K-major tiles.
//   warp 1      TMEM allocation + single-thread tcgen05.mma issue (UMMA 128 x BN x 16, kind::f16,
//               bf16 inputs, fp32 accumulate), tcgen05.commit onto the ring's "empty" barriers and the
//               accumulator's "full" barrier.
//   warps 2-9   epilogue (STEM variant: warps 2-5 epilogue, warps 6-13 operand producers): tcgen05.ld the 128 x BN fp32 accumulator (double-buffered in TMEM so the next
//               tile's MMAs overlap), folded-BN scale/bias, residual, activation, re-split to bf16
//               planes (and/or fp32), masked stores.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include <mutex>
#include <string.h>
#include <stdlib.h>

namespace {

constexpr int BM = 128;       // UMMA_M (cta_group::1)
constexpr int BK = 64;        // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kStagesDefault = 3;   // BN <= 128: 3 x 64 KB; BN = 256: 2 x 96 KB
constexpr int kThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KB

// ---------------------------------------------------------------------------------------------
struct GemmParams {
  int rows_box;                       // rows TMA delivers per A tile (bn*bh*bw <= 128)
  int bn, bh, bw;                     // A box in output pixels
  int tiles_w, tiles_h, tiles_img;    // M tiling
  int tiles_n;                        // N tiling
  int N, Ho, Wo, Cout;
  int KH, KW, stride, pad, cin_blocks, cin;  // cin_blocks = ceil(Cin / 64)
  int passes, act;
  const float* scale;
  const float* bias;
  const uint16_t* res_hi;
  const uint16_t* res_lo;
  const uint16_t* mask_hi;            // dgrad: output zeroed where this activation plane (hi) is not > 0 (fused ReLU backward)
  uint16_t* y_hi;
  uint16_t* y_lo;
  float* y_f32;
  int res_mma;    // residual added on the tensor core: extra k-blocks R[128 x 64] * I[BN x 64]^T (needs scale == NULL)
  int tma_store;  // planes output leaves through shared memory + cp.async.bulk.tensor stores
  int two_chains; // split planes, BN <= 128: even / odd k-blocks accumulate in two TMEM accumulators, summed in the epilogue
  // where output pixel (n, ho, wo) lives, in pixels of the y / residual / mask tensors: n * o_img + ho * o_h + wo * o_w + o_off
  // (dense: Ho*Wo, Wo, 1, 0; the stride-2 dgrad writes one parity class of a tensor twice as large in h and w)
  long long o_img, o_h, o_w, o_off;
  // STEM variant only: the A operand is gathered from the raw uint8 NHWC image by producer warps
  const uint8_t* img;       // [N, H_in, W_in, 3] uint8 (STEM_MODE 1)  or  float32 [N, 3, H_in, W_in] (STEM_MODE 2)
  float nmean[3], nstd[3];  // STEM_MODE 2: Normalize constants (mode 1 has them baked into the LUT)
  const uint32_t* lut;      // [3][256]: normalised value of byte b in channel c as (hi | lo << 16) fp16 pair
  int H_in, W_in, stem_Ho, stem_Wo;
  long long M_total;        // N * Ho * Wo
};

constexpr int kStemProducerWarps = 8;   // two threads per A-tile row
constexpr int kStemEpiWarps = 4;        // the stem's 64-column epilogue needs only one warp per TMEM lane quarter

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case B200R_ACT_RELU: return fmaxf(v, 0.f);
    case B200R_ACT_RELU6: return fminf(fmaxf(v, 0.f), 6.f);
    case B200R_ACT_GELU_TANH: {  // vision_transformer.py:19-37
      const float k = 0.7978845608028654f;
      return 0.5f * v * (1.f + tanhf(k * (v + 0.044715f * v * v * v)));
    }
    case B200R_ACT_GELU_ERF: return 0.5f * v * (1.f + erff(v * 0.7071067811865476f));
    case B200R_ACT_SWISH: return v / (1.f + expf(-v));
    case B200R_ACT_TANH: return tanhf(v);
    case B200R_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// erf by Abramowitz & Stegun 7.1.26 (|error| < 1.5e-7) on ex2 / rcp: 12 instructions instead of erff's 30
__device__ __forceinline__ float fast_erf(float x) {
  const float ax = fabsf(x), t = __fdividef(1.f, fmaf(0.3275911f, ax, 1.f));
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  return copysignf(1.f - poly * __expf(-ax * ax), x);
}

// STEM_MODE: 0 = TMA-fed convolution / linear, 1 = fused stem from the uint8 NHWC image, 2 = fused stem from a float32
// NCHW image in [0,1] (the attack path)
// LEAN (not the stem): the epilogue instantiation for the common layer -- planes out through TMA stores (Cout % 8 == 0),
// bias / scale / activation / dgrad mask, residual on the tensor core or none, no fp32 output.  The general epilogue carries every variant (ragged columns, LSU
// residual, fp32 outputs, five activations, split planes) behind run-time branches: 26 000 SASS instructions of which a
// layer executes 900, spread over 400 KB of code -- ncu showed 22 % of the epilogue warps' samples stalled on instruction
// fetch (stall_no_inst).  LEAN also double-buffers the staging tile so a store unit costs one named barrier and never waits
// for its own TMA store (ring one stage shorter to pay for the second 32 KB).
// ACTS (LEAN only): 0 = the instantiation knows none / ReLU only (the convolution stacks), 1 = all activations.
// PAIR (LEAN, split planes, BN = 128): launched as clusters of two CTAs on the two SMs of a TPC; the pair runs ONE M = 256
// MMA (tcgen05.mma.cta_group::2) over two vertically adjacent M tiles of the same N tile.  Each CTA loads its own A tile and
// only HALF of the weight tile (the tensor core reads the other half from the peer's shared memory), so the operand bytes a
// CTA pulls from L2 per k-block drop from 64 KB to 48 KB -- tools/gemm_probe.py shows these GEMMs bound by exactly that
// (12.5 TB/s of operand delivery = the L2 throughput cap, not the tensor pipe).  The leader (cluster rank 0) issues the MMAs;
// both CTAs' TMA loads count on the leader's `full` barrier; MMA completion is multicast to both CTAs' `empty` / `tmem_full`
// barriers; both epilogues arrive on the leader's `tmem_empty`.
// DUAL (LEAN, split planes, ACTS = 1): the pre-activation planes leave through a second tensor map (map_y2) before the activation
// is applied -- the forward a gradient pass keeps (ViT / Mixer MLP blocks), one launch instead of GEMM + activation pass.
template <int BN, int STEM_MODE, bool F16, bool LEAN = false, int ACTS = 1, bool PAIR = false, bool DUAL = false>
__global__ void __launch_bounds__(STEM_MODE ? (2 + kStemEpiWarps + kStemProducerWarps) * 32 : kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_i,
            const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_y2, const GemmParams p) {
  constexpr bool STEM = STEM_MODE != 0;
  static_assert(!DUAL || (LEAN && !F16 && ACTS == 1), "DUAL: LEAN split-plane kernel with the full activation set");
  constexpr bool STEM_F32 = STEM_MODE == 2;
  static_assert(!LEAN || !STEM, "LEAN is a non-stem epilogue");
  static_assert(!PAIR || (LEAN && !F16 && BN <= 128), "PAIR: LEAN split-plane kernel with BN = 64 / 128");
  // LEAN + fp16 double-buffers the 32 KB staging area and gives up ring stages for it
  constexpr int kStages = PAIR ? 4 : (LEAN && F16) ? (BN > 128 ? 3 : (BN == 128 ? 5 : 6)) : ((BN > 128) ? 2 : (BN == 64 && !STEM ? 4 : kStagesDefault)) * (F16 ? 2 : 1);
  constexpr int BN_LOAD = PAIR ? BN / 2 : BN;           // weight rows this CTA holds
  constexpr int B_TILE_BYTES = BN_LOAD * BK * 2;
  constexpr int B_OFF = F16 ? A_TILE_BYTES : 2 * A_TILE_BYTES;           // F16: [A | B]; split: [A_hi | A_lo | B_hi | B_lo]
  constexpr int STAGE_BYTES = B_OFF + (F16 ? 1 : 2) * B_TILE_BYTES;
  // tcgen05's fp32 accumulation truncates (profiles/r2_accumulation_bias.txt: -0.77 * 2^-23 relative per 16-wide MMA step, a pure bias
  // that shrinks every activation of a deep network).  With split planes and BN <= 128 the even and the odd k-blocks accumulate in
  // two separate TMEM accumulators (two chains of half the length at about half the magnitude: half the bias), summed with one
  // round-to-nearest add in the epilogue.
  constexpr bool TWO = !F16 && !STEM && BN <= 128;
  constexpr uint32_t ACC_COLS = TWO ? 2 * BN : BN;
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;  // two accumulator buffers (power of two: 128 .. 512)
  // instruction descriptor: fp32 accumulator (bit 4), A/B format (bits 7, 10: 1 = bf16, 0 = fp16), N >> 3, M >> 4
  // (both precisions now feed fp16 operands: split planes are fp16 pairs)
  constexpr uint32_t UMMA_M = PAIR ? 2 * BM : BM;
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(UMMA_M >> 4) << 24);
  constexpr uint32_t IDESC_RES = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(UMMA_M >> 4) << 24);   // residual k-blocks: N = 64

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int BAR_OFF = kStages * STAGE_BYTES;        // 1 KB: mbarriers + TMEM slot
  constexpr int STG_OFF = BAR_OFF + 1024;               // 2 x 16 KB epilogue staging ([2 planes][128 rows][64 B] per warp group)
  constexpr int STEM_OFF = STG_OFF + 32768;             // STEM: staged input rows + LUT (LEAN: the staging area is 2 x 32 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  // bars: [0,kStages) full, [kStages,2kStages) empty, [2k,2k+2) tmem_full, [2k+2,2k+4) tmem_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp index as a uniform value
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    if (p.tma_store) prefetch_tmap(&map_y);
    if constexpr (DUAL) prefetch_tmap(&map_y2);
    if (p.res_mma) { prefetch_tmap(&map_r); prefetch_tmap(&map_i); }
    // STEM: the A tile is written by 128 producer threads (one arrival each) next to the TMA thread's
    // arrive.expect_tx for the weight tile
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), STEM ? 1 + kStemProducerWarps * 32 : 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), STEM ? kStemEpiWarps : (PAIR ? 16 : 8)); }
    fence_barrier_init();
  }
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;  // 0 = the pair's leader
  if (PAIR) cluster_sync_all();                          // the peer's barriers exist before anything signals them
  if (warp == 1) {
    if (PAIR) { tmem_alloc_pair(smem_u32(tmem_slot), TMEM_COLS); tmem_relinquish_pair(); }
    else { tmem_alloc(smem_u32(tmem_slot), TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_img;
  // PAIR: the schedule runs over pairs of M tiles (2 mp, 2 mp + 1); an odd tail tile's partner is out of range (zero-filled
  // loads, clipped stores)
  const int total_tiles = (PAIR ? (m_tiles + 1) / 2 : m_tiles) * p.tiles_n;
  // tile schedule: round-robin over CTAs (PAIR: over clusters); the STEM variant takes a contiguous range instead, so that consecutive
  // tiles of a CTA are consecutive output rows of one image and the staged input rows slide by two
  const int t_first = STEM ? (int)((long long)blockIdx.x * total_tiles / gridDim.x) : (PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x);
  const int t_end = STEM ? (int)((long long)(blockIdx.x + 1) * total_tiles / gridDim.x) : total_tiles;
  const int t_step = STEM ? 1 : (PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x);
  auto m_tile_of = [&](int t) { return PAIR ? 2 * (t / p.tiles_n) + (int)rank : t / p.tiles_n; };
  const int kb_conv = p.KH * p.KW * p.cin_blocks;
  const int kb_res = p.res_mma ? BN / 64 : 0;           // residual k-blocks: acc[:, 64j:64j+64] += R[:, 64j:64j+64] * I64^T
  const int kblocks = kb_conv + kb_res;
  const bool two = TWO && p.two_chains && kb_conv >= 2;
  const uint32_t a_bytes = (uint32_t)p.rows_box * BK * 2;
  const uint32_t tx_bytes = (p.passes == 3 ? 2u : 1u) * ((STEM ? 0u : a_bytes) + (uint32_t)B_TILE_BYTES);

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp runs the (warp-uniform) loop and waits on the barriers; one elected lane issues.  Keeping the
    // control flow uniform lets the compiler hold addresses / coordinates in uniform registers -- a divergent
    // `if (lane == 0)` body costs an ELECT + branch waterfall around every UTMALDG / UTCHMMA.
    int stage = 0;
    uint32_t phase = 0;
    for (int t = t_first; t < t_end; t += t_step) {
      const int nt = t % p.tiles_n, mt = m_tile_of(t);
      const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, ti = mt / (p.tiles_w * p.tiles_h);
      const int w_in0 = tw * p.bw * p.stride - p.pad, h_in0 = th * p.bh * p.stride - p.pad, n0 = ti * p.bn;
      for (int kh = 0; kh < p.KH; ++kh)
        for (int kw = 0; kw < p.KW; ++kw)
          for (int cb = 0; cb < p.cin_blocks; ++cb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            if (elect_one()) {
              const uint32_t sa = smem_base + stage * STAGE_BYTES;
              const int kcol = (kh * p.KW + kw) * p.cin + cb * BK;
              if constexpr (PAIR) {
                const uint32_t lbar = mapa_u32(full_bar(stage), 0);      // the leader's barrier counts both CTAs' bytes
                if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * tx_bytes);
                tma_load_5d_pair(sa, &map_a, lbar, cb * BK, w_in0 + kw, h_in0 + kh, n0, 0);
                tma_load_3d_pair(sa + B_OFF, &map_b, lbar, kcol, nt * BN + (int)rank * BN_LOAD, 0);
                if (p.passes == 3) {
                  tma_load_5d_pair(sa + A_TILE_BYTES, &map_a, lbar, cb * BK, w_in0 + kw, h_in0 + kh, n0, 1);
                  tma_load_3d_pair(sa + B_OFF + B_TILE_BYTES, &map_b, lbar, kcol, nt * BN + (int)rank * BN_LOAD, 1);
                }
              } else {
              mbar_expect_tx(full_bar(stage), tx_bytes);
              if (!STEM) tma_load_5d(sa, &map_a, full_bar(stage), cb * BK, w_in0 + kw, h_in0 + kh, n0, 0);
              tma_load_3d(sa + B_OFF, &map_b, full_bar(stage), kcol, nt * BN, 0);
              if (!F16 && p.passes == 3) {
                if (!STEM) tma_load_5d(sa + A_TILE_BYTES, &map_a, full_bar(stage), cb * BK, w_in0 + kw, h_in0 + kh, n0, 1);
                tma_load_3d(sa + B_OFF + B_TILE_BYTES, &map_b, full_bar(stage), kcol, nt * BN, 1);
              }
              }
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
      for (int j = 0; j < kb_res; ++j) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        if (elect_one()) {
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          // residual channels [64j, 64j+64) only touch accumulator columns [64j, 64j+64): the B operand is the 64 x 64
          // identity (8 KB) and the MMA has N = 64, whatever BN is
          if constexpr (PAIR) {
            // each CTA holds 32 of the identity's 64 rows (map_i's box is 32 rows for a pair launch)
            const uint32_t lbar = mapa_u32(full_bar(stage), 0);
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * (2u * a_bytes + (uint32_t)(32 * BK * 2)));
            tma_load_5d_pair(sa, &map_r, lbar, nt * BN + j * BK, tw * p.bw, th * p.bh, n0, 0);
            tma_load_5d_pair(sa + A_TILE_BYTES, &map_r, lbar, nt * BN + j * BK, tw * p.bw, th * p.bh, n0, 1);
            tma_load_2d_pair(sa + B_OFF, &map_i, lbar, 0, (int)rank * 32);
          } else {
          mbar_expect_tx(full_bar(stage), (F16 ? 1u : 2u) * a_bytes + (uint32_t)(64 * BK * 2));
          tma_load_5d(sa, &map_r, full_bar(stage), nt * BN + j * BK, tw * p.bw, th * p.bh, n0, 0);
          if (!F16) tma_load_5d(sa + A_TILE_BYTES, &map_r, full_bar(stage), nt * BN + j * BK, tw * p.bw, th * p.bh, n0, 1);
          tma_load_2d(sa + B_OFF, &map_i, full_bar(stage), 0, 0);
          }
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues; PAIR: the leader CTA only) =====================
    auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
      if constexpr (PAIR) umma_f16_pair(d, a, b, idesc, accumulate); else umma_bf16(d, a, b, idesc, accumulate);
    };
    auto commit = [&](uint32_t bar) { if constexpr (PAIR) umma_commit_pair(bar); else umma_commit(bar); };
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = (PAIR && rank != 0) ? t_end : t_first; t < t_end; t += t_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_base = tmem_base + (uint32_t)(acc * ACC_COLS);
      for (int kb = 0; kb < kb_conv; ++kb) {
        const uint32_t d_tmem = d_base + ((two && (kb & 1)) ? (uint32_t)BN : 0u);
        const int first_kb = (two && (kb & 1)) ? 1 : 0;          // the k-block that initialises this chain
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint64_t a_hi = make_sw128_desc(sa), a_lo = make_sw128_desc(sa + A_TILE_BYTES);
          const uint64_t b_hi = make_sw128_desc(sa + B_OFF);
          const uint64_t b_lo = make_sw128_desc(sa + B_OFF + B_TILE_BYTES);
          if (!F16 && p.passes == 3) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);  // +32 B per K step inside the swizzle row
              mma(d_tmem, a_lo + adv, b_hi + adv, IDESC, ((kb - first_kb) | k) != 0);
              mma(d_tmem, a_hi + adv, b_lo + adv, IDESC, 1);
              mma(d_tmem, a_hi + adv, b_hi + adv, IDESC, 1);
            }
          } else {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
              mma(d_tmem, a_hi + adv, b_hi + adv, IDESC, ((kb - first_kb) | k) != 0);
            }
          }
          commit(empty_bar(stage));                 // smem slot reusable once these MMAs retire
          if (kb == kblocks - 1) commit(tfull_bar(acc));  // accumulator complete
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      for (int j = 0; j < kb_res; ++j) {                   // residual: (R_hi + R_lo) * identity, exact in the fp32 accumulator
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint64_t a_hi = make_sw128_desc(sa), a_lo = make_sw128_desc(sa + A_TILE_BYTES);
          const uint64_t b_hi = make_sw128_desc(sa + B_OFF);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
            mma(d_base + (uint32_t)(j * 64), a_hi + adv, b_hi + adv, IDESC_RES, 1);
            if (!F16) mma(d_base + (uint32_t)(j * 64), a_lo + adv, b_hi + adv, IDESC_RES, 1);
          }
          commit(empty_bar(stage));
          if (j == kb_res - 1) commit(tfull_bar(acc));
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < 2 + (STEM ? kStemEpiWarps : 8)) {
    // ===================== epilogue (8 warps; 4 in the STEM variant) =====================
    // TMEM lane quarter = warp % 4 (hardware restriction); the two warp groups split the columns.  Planes output
    // leaves through a 16 KB staging buffer per group ([plane][128 rows][32 channels]) and one TMA store per plane
    // and 32-column chunk: no LSU global traffic, ragged rows / columns clipped by the tensor map.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const int box_hw = p.bh * p.bw;
    const int nl = r / box_hw, rem = r - nl * box_hw, hl = rem / p.bw, wl = rem - hl * p.bw;
    if constexpr (LEAN) {
      // fp16: store unit = 64 columns = full 128-byte rows of a SWIZZLE_128B staging tile (16 KB, one TMA store) -- the two
      // 32-column chunks of one warp group (BN >= 128) or one chunk of each group (BN = 64); units alternate between two
      // staging buffers.  split-bf16: store unit = one 32-column chunk, hi and lo planes as two SWIZZLE_64B tiles (8 KB
      // each) in the group's single 16 KB buffer.
      constexpr int kColsPerWarp = BN / 2, kChunks = kColsPerWarp / 32, kPer = kChunks >= 2 ? 2 : 1;
      constexpr bool kSharedUnit = F16 && BN == 64;
      const bool issuer = (quarter == 0) && (lane == 0);
      const bool uissuer = issuer && (!kSharedUnit || half == 0);
      uint32_t sb = 0;                                       // fp16: staging buffer of the next unit
      int it = 0;
      for (int t = t_first; t < t_end; t += t_step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const int nt = t % p.tiles_n, mt = m_tile_of(t);
        const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, ti = mt / (p.tiles_w * p.tiles_h);
        const int n_img = ti * p.bn + nl, ho = th * p.bh + hl, wo = tw * p.bw + wl;
        const bool row_ok = (r < p.rows_box) && (n_img < p.N) && (ho < p.Ho) && (wo < p.Wo);
        const size_t out_row = (size_t)((long long)n_img * p.o_img + (long long)ho * p.o_h + (long long)wo * p.o_w + p.o_off);
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS + half * kColsPerWarp);
#pragma unroll 1
        for (int rd = 0; rd < kChunks / kPer; ++rd) {
          uint32_t vv[kPer][32];
#pragma unroll
          for (int i = 0; i < kPer; ++i) tmem_ld32(t_addr + (uint32_t)((rd * kPer + i) * 32), vv[i]);
          tmem_ld_wait();
          if (two) {                                          // second accumulator chain (odd k-blocks)
            uint32_t ww[kPer][32];
#pragma unroll
            for (int i = 0; i < kPer; ++i) tmem_ld32(t_addr + (uint32_t)(BN + (rd * kPer + i) * 32), ww[i]);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < kPer; ++i)
#pragma unroll
              for (int j = 0; j < 32; ++j) vv[i][j] = __float_as_uint(__uint_as_float(vv[i][j]) + __uint_as_float(ww[i][j]));
          }
          if (rd == kChunks / kPer - 1) {                     // accumulator is in registers: hand the TMEM buffer back early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(tempty_bar(acc), 0)); else mbar_arrive(tempty_bar(acc)); }
          }
#pragma unroll
          for (int ci = 0; ci < kPer; ++ci) {
            const uint32_t (&v)[32] = vv[ci];
            const int col0 = nt * BN + half * kColsPerWarp + (rd * kPer + ci) * 32;
            const bool live = col0 < p.Cout;                  // uniform across the warp group (ragged Cout: whole chunks drop out)
            float f[32];
            // split planes: one 32-column chunk, hi and lo, through the group's staging buffer and two TMA stores
            auto emit_planes = [&](const CUtensorMap* map) {
              // the group's buffer must have been read by the previous chunk's stores before it is rewritten
              uint8_t* stg = smem + STG_OFF + half * 16384;
              if (issuer) bulk_wait_read0();
              named_bar_sync(2 + half, 128);
              if (live) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  uint32_t ph[4], pl[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float x0 = f[8 * q + 2 * j], x1 = f[8 * q + 2 * j + 1];
                    split_f16x2(x1, x0, ph[j], pl[j]);
                  }
                  // 64-byte rows in the TMA SWIZZLE_64B pattern (16-byte chunk index ^= (row >> 1) & 3): conflict-free stores
                  const int pos = (q ^ ((r >> 1) & 3)) * 16;
                  *reinterpret_cast<uint4*>(stg + r * 64 + pos) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                  *reinterpret_cast<uint4*>(stg + 8192 + r * 64 + pos) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                }
              }
              fence_proxy_async();
              named_bar_sync(2 + half, 128);
              if (issuer && live) {
                const uint32_t stg_u32 = smem_base + STG_OFF + half * 16384;
                tma_store_5d(map, stg_u32, col0, tw * p.bw, th * p.bh, ti * p.bn, 0);
                tma_store_5d(map, stg_u32 + 8192, col0, tw * p.bw, th * p.bh, ti * p.bn, 1);
                bulk_commit();
              }
            };
            if (live) {
              if (col0 + 32 <= p.Cout) {
                if (p.scale) {
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.scale + col0) + q);
                    const float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                    f[4 * q + 0] = fmaf(__uint_as_float(v[4 * q + 0]), s4.x, b4.x);
                    f[4 * q + 1] = fmaf(__uint_as_float(v[4 * q + 1]), s4.y, b4.y);
                    f[4 * q + 2] = fmaf(__uint_as_float(v[4 * q + 2]), s4.z, b4.z);
                    f[4 * q + 3] = fmaf(__uint_as_float(v[4 * q + 3]), s4.w, b4.w);
                  }
                } else if (p.bias) {
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + q);
                    f[4 * q + 0] = __uint_as_float(v[4 * q + 0]) + b4.x;
                    f[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + b4.y;
                    f[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + b4.z;
                    f[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + b4.w;
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                }
              } else {                                        // last, partial chunk of a ragged Cout (the TMA store clips it)
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const int col = min(col0 + j, p.Cout - 1);
                  f[j] = fmaf(__uint_as_float(v[j]), p.scale ? __ldg(p.scale + col) : 1.f, p.bias ? __ldg(p.bias + col) : 0.f);
                }
              }
              if (p.mask_hi && row_ok) {                      // ReLU backward fused into the dgrad GEMM: one 16-bit plane read
                const size_t off = out_row * p.Cout + col0;
                if (col0 + 32 <= p.Cout && (p.Cout % 16 == 0)) {
                  uint32_t mw[16];
                  ld_global_v8(p.mask_hi + off, mw);
                  ld_global_v8(p.mask_hi + off + 16, mw + 8);
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const uint32_t a0 = mw[j] & 0xFFFFu, a1 = mw[j] >> 16;
                    if (a0 == 0 || (a0 & 0x8000u)) f[2 * j] = 0.f;
                    if (a1 == 0 || (a1 & 0x8000u)) f[2 * j + 1] = 0.f;
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (col0 + j < p.Cout) {
                      const uint32_t a = p.mask_hi[off + j];
                      if (a == 0 || (a & 0x8000u)) f[j] = 0.f;
                    }
                }
              }
            }
            if constexpr (DUAL) emit_planes(&map_y2);         // the pre-activation, as the plain GEMM would have stored it
            if (live) {
              if constexpr (ACTS == 0) {
                if (p.act == B200R_ACT_RELU) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                }
              } else {
                // activations on the MUFU (ex2 / rcp): each case is one straight run of code, absolute error ~1e-7
                switch (p.act) {
                  case B200R_ACT_NONE: break;
                  case B200R_ACT_RELU:
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                    break;
                  case B200R_ACT_RELU6:
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = fminf(fmaxf(f[j], 0.f), 6.f);
                    break;
                  case B200R_ACT_GELU_TANH:                      // 0.5 v (1 + tanh(u)) = v / (1 + exp(-2u)), u = k (v + 0.044715 v^3)
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                      const float x = f[j], u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
                      f[j] = __fdividef(x, 1.f + __expf(-2.f * u));
                    }
                    break;
                  case B200R_ACT_GELU_ERF:
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = 0.5f * f[j] * (1.f + fast_erf(f[j] * 0.7071067811865476f));
                    break;
                  case B200R_ACT_SWISH:
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __fdividef(f[j], 1.f + __expf(-f[j]));
                    break;
                  case B200R_ACT_TANH:
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = 1.f - __fdividef(2.f, 1.f + __expf(2.f * f[j]));
                    break;
                  default:                                        // B200R_ACT_SIGMOID
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __fdividef(1.f, 1.f + __expf(-f[j]));
                    break;
                }
              }
            }
            if constexpr (F16) {
              if (live) {
                uint8_t* ustg = smem + STG_OFF + sb * 32768 + (kSharedUnit ? 0 : half * 16384);
                const int sub = kSharedUnit ? half : ci;         // which 64-byte half of the 128-byte rows
#pragma unroll
                for (int q = 0; q < 4; ++q)                       // SWIZZLE_128B: 16-byte chunk j of row r sits at chunk j ^ (r & 7)
                  *reinterpret_cast<uint4*>(ustg + r * 128 + (((sub * 4 + q) ^ (r & 7)) << 4)) =
                      make_uint4(cvt_f16x2(f[8 * q + 1], f[8 * q]), cvt_f16x2(f[8 * q + 3], f[8 * q + 2]),
                                 cvt_f16x2(f[8 * q + 5], f[8 * q + 4]), cvt_f16x2(f[8 * q + 7], f[8 * q + 6]));
              }
            } else {
              emit_planes(&map_y);
            }
          }
          if constexpr (F16) {
            fence_proxy_async();
            // the previous unit's store (other buffer) was issued a whole unit ago: once it has read its tile, everybody
            // past the barrier may overwrite that buffer while this unit's store is in flight
            if (uissuer) bulk_wait_read0();
            if (kSharedUnit) named_bar_sync(4, 256); else named_bar_sync(2 + half, 128);
            const int ucol0 = nt * BN + (kSharedUnit ? 0 : half * kColsPerWarp + rd * 64);
            if (uissuer && ucol0 < p.Cout) {
              tma_store_5d(&map_y, smem_base + STG_OFF + sb * 32768 + (kSharedUnit ? 0 : half * 16384), ucol0, tw * p.bw, th * p.bh, ti * p.bn, 0);
              bulk_commit();
            }
            sb ^= 1;
          }
        }
      }
      if (issuer) bulk_wait0();                                 // all stores complete before the CTA exits
    } else {
    uint8_t* stg = smem + STG_OFF + half * 16384;
    const uint32_t stg_u32 = smem_base + STG_OFF + half * 16384;
    const bool issuer = (quarter == 0) && (lane == 0);
    int it = 0;
    for (int t = t_first; t < t_end; t += t_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int nt = t % p.tiles_n, mt = m_tile_of(t);
      const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, ti = mt / (p.tiles_w * p.tiles_h);
      const int n_img = ti * p.bn + nl, ho = th * p.bh + hl, wo = tw * p.bw + wl;
      const bool row_ok = (r < p.rows_box) && (n_img < p.N) && (ho < p.Ho) && (wo < p.Wo);
      const size_t out_row = (size_t)((long long)n_img * p.o_img + (long long)ho * p.o_h + (long long)wo * p.o_w + p.o_off);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      constexpr int kColsPerWarp = STEM ? BN : BN / 2;
      constexpr int kChunks = kColsPerWarp / 32;
      constexpr int kPer = kChunks >= 2 ? 2 : 1;            // accumulator chunks fetched per tcgen05.wait::ld
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS + half * kColsPerWarp);
#pragma unroll
      for (int rd = 0; rd < kChunks / kPer; ++rd) {
      uint32_t vv[kPer][32];
#pragma unroll
      for (int i = 0; i < kPer; ++i) tmem_ld32(t_addr + (uint32_t)((rd * kPer + i) * 32), vv[i]);
      tmem_ld_wait();
      if (two) {                                            // second accumulator chain (odd k-blocks)
        uint32_t ww[kPer][32];
#pragma unroll
        for (int i = 0; i < kPer; ++i) tmem_ld32(t_addr + (uint32_t)(BN + (rd * kPer + i) * 32), ww[i]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < kPer; ++i)
#pragma unroll
          for (int j = 0; j < 32; ++j) vv[i][j] = __float_as_uint(__uint_as_float(vv[i][j]) + __uint_as_float(ww[i][j]));
      }
      if (rd == kChunks / kPer - 1) {                       // accumulator is in registers: hand the TMEM buffer back early
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }
      // F16 + TMA store: a store unit is 64 columns = full 128-byte rows (SWIZZLE_128B staging, one 16 KB store); the
      // 32-column chunks of one warp group (BN >= 128, STEM) or the two groups (BN = 64) fill its halves
      constexpr bool kSharedUnit = F16 && !STEM && BN == 64;
      const bool unit_store = F16 && p.tma_store && p.y_hi;
      const int ucol0 = nt * BN + (kSharedUnit ? 0 : half * kColsPerWarp + rd * 64);
      uint8_t* ustg = kSharedUnit ? smem + STG_OFF : stg;
      const bool uissuer = issuer && (!kSharedUnit || half == 0);
      if (unit_store) {
        if (uissuer) bulk_wait_read0();                    // the previous store has finished reading the staging buffer
        if (kSharedUnit) named_bar_sync(4, 256); else named_bar_sync(2 + half, 128);
      }
#pragma unroll
      for (int ci = 0; ci < kPer; ++ci) {
        const uint32_t (&v)[32] = vv[ci];
        const int c0 = half * kColsPerWarp + (rd * kPer + ci) * 32;
        const int col0 = nt * BN + c0;
        if (col0 >= p.Cout) continue;                     // uniform across the warp group
        const size_t off = out_row * p.Cout + col0;
        const bool full = (col0 + 32 <= p.Cout);
        float f[32];
        if (full) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 s4 = p.scale ? __ldg(reinterpret_cast<const float4*>(p.scale + col0) + q) : make_float4(1.f, 1.f, 1.f, 1.f);
            const float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            f[4 * q + 0] = fmaf(__uint_as_float(v[4 * q + 0]), s4.x, b4.x);
            f[4 * q + 1] = fmaf(__uint_as_float(v[4 * q + 1]), s4.y, b4.y);
            f[4 * q + 2] = fmaf(__uint_as_float(v[4 * q + 2]), s4.z, b4.z);
            f[4 * q + 3] = fmaf(__uint_as_float(v[4 * q + 3]), s4.w, b4.w);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = min(col0 + j, p.Cout - 1);
            f[j] = fmaf(__uint_as_float(v[j]), p.scale ? __ldg(p.scale + col) : 1.f, p.bias ? __ldg(p.bias + col) : 0.f);
          }
        }
        if (p.res_hi && !p.res_mma && row_ok) {           // LSU fallback (scaled convolutions with a residual)
          if (full && (p.Cout % 16 == 0)) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              uint32_t hw[8], lw[8];
              ld_global_v8(p.res_hi + off + 16 * q, hw);
              if (F16) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float2 r2 = f16x2_to_f32(hw[j]);
                  f[16 * q + 2 * j] += r2.x;
                  f[16 * q + 2 * j + 1] += r2.y;
                }
              } else {
                ld_global_v8(p.res_lo + off + 16 * q, lw);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float2 rh = f16x2_to_f32(hw[j]), rl = f16x2_to_f32(lw[j]);
                  f[16 * q + 2 * j] += rh.x + rl.x;
                  f[16 * q + 2 * j + 1] += rh.y + rl.y;
                }
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)   // compile-time j: the arrays must stay in registers
              if (col0 + j < p.Cout)
                f[j] += F16 ? __half2float(__ushort_as_half(p.res_hi[off + j]))
                            : plane_bits_to_f32(p.res_hi[off + j]) + plane_bits_to_f32(p.res_lo[off + j]);
          }
        }
        if (p.mask_hi && row_ok) {                        // ReLU backward fused into the dgrad GEMM: one bf16 plane read
          if (full && (p.Cout % 16 == 0)) {
            uint32_t mw[16];
            ld_global_v8(p.mask_hi + off, mw);
            ld_global_v8(p.mask_hi + off + 16, mw + 8);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const uint32_t a0 = mw[j] & 0xFFFFu, a1 = mw[j] >> 16;
              if (a0 == 0 || (a0 & 0x8000u)) f[2 * j] = 0.f;
              if (a1 == 0 || (a1 & 0x8000u)) f[2 * j + 1] = 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.Cout) {
                const uint32_t a = p.mask_hi[off + j];
                if (a == 0 || (a & 0x8000u)) f[j] = 0.f;
              }
          }
        }
        if (p.act == B200R_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        } else if (p.act != B200R_ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], p.act);
        }
        if (p.y_f32 && row_ok) {
          if (full && (p.Cout % 4 == 0)) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              reinterpret_cast<float4*>(p.y_f32 + off)[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.Cout) p.y_f32[off + j] = f[j];
          }
        }
        if (p.y_hi) {
          uint32_t ph[16], pl[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (F16) {
              ph[j] = cvt_f16x2(f[2 * j + 1], f[2 * j]);
              pl[j] = 0u;
            } else {
              split_f16x2(f[2 * j + 1], f[2 * j], ph[j], pl[j]);
            }
          }
          if (F16 && p.tma_store) {
            const int sub = kSharedUnit ? half : ci;         // which 64-byte half of the 128-byte rows
#pragma unroll
            for (int q = 0; q < 4; ++q)                       // SWIZZLE_128B: 16-byte chunk j of row r sits at chunk j ^ (r & 7)
              *reinterpret_cast<uint4*>(ustg + r * 128 + (((sub * 4 + q) ^ (r & 7)) << 4)) =
                  make_uint4(ph[4 * q], ph[4 * q + 1], ph[4 * q + 2], ph[4 * q + 3]);
          } else if (p.tma_store) {
            if (issuer) bulk_wait_read0();                 // the previous store has finished reading the staging buffer
            named_bar_sync(2 + half, 128);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              // 64-byte rows in the TMA SWIZZLE_64B pattern (16-byte chunk index ^= address bits [7:8] = (row >> 1) & 3):
              // 8 consecutive rows cover all 32 banks, so the 16-byte stores are conflict free
              const int pos = (q ^ ((r >> 1) & 3)) * 16;
              *reinterpret_cast<uint4*>(stg + r * 64 + pos) = make_uint4(ph[4 * q], ph[4 * q + 1], ph[4 * q + 2], ph[4 * q + 3]);
              if (!F16) *reinterpret_cast<uint4*>(stg + 8192 + r * 64 + pos) = make_uint4(pl[4 * q], pl[4 * q + 1], pl[4 * q + 2], pl[4 * q + 3]);
            }
            fence_proxy_async();
            named_bar_sync(2 + half, 128);
            if (issuer) {
              tma_store_5d(&map_y, stg_u32, col0, tw * p.bw, th * p.bh, ti * p.bn, 0);
              if (!F16) tma_store_5d(&map_y, stg_u32 + 8192, col0, tw * p.bw, th * p.bh, ti * p.bn, 1);
              bulk_commit();
            }
          } else if (row_ok && full && (p.Cout % 16 == 0)) {   // direct 256-bit stores
            st_global_v8(p.y_hi + off, ph);
            st_global_v8(p.y_hi + off + 16, ph + 8);
            if (!F16) {
              st_global_v8(p.y_lo + off, pl);
              st_global_v8(p.y_lo + off + 16, pl + 8);
            }
          } else if (row_ok) {                             // ragged direct stores
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.Cout) {
                p.y_hi[off + j] = (uint16_t)(ph[j >> 1] >> (16 * (j & 1)));
                if (!F16) p.y_lo[off + j] = (uint16_t)(pl[j >> 1] >> (16 * (j & 1)));
              }
          }
        }
      }
      if (unit_store) {
        fence_proxy_async();
        if (kSharedUnit) named_bar_sync(4, 256); else named_bar_sync(2 + half, 128);
        if (uissuer && ucol0 < p.Cout) {
          tma_store_5d(&map_y, smem_base + STG_OFF + (kSharedUnit ? 0 : half * 16384), ucol0, tw * p.bw, th * p.bh, ti * p.bn, 0);
          bulk_commit();
        }
      }
      }
    }
    if (issuer) bulk_wait0();                              // all stores complete before the CTA exits
    }
  } else if (STEM) {
    // ===================== stem A producer (8 warps) =====================
    // 7x7 / stride 2 / pad 3 patches of the uint8 NHWC image.  A tile = ONE output row of one image
    // (Wo <= 128 pixels).  The 7 input rows a tile touches live in a 7-slot ring in shared memory (row iy in slot
    // (iy + 3) % 7), already mapped through the 3x256 LUT (ToTensor + Normalize + bf16 split -> hi | lo<<16) and
    // padded with zero columns/rows, so the per-pixel gather needs no bounds checks: K is ordered (ky, kx8, c)
    // with EIGHT kx slots per ky (the 8th carries a zero weight), i.e. 7 runs of 24 consecutive staged words
    // starting at word 6*ox = three aligned 8-word chunks each, read with 64-bit loads.  Consecutive tiles of a CTA are
    // consecutive output rows, so only two new input rows are staged per tile; their global loads are issued
    // before the gather of the current tile and land in the ring after it (the two slots they replace are dead
    // by then).  The 64-element k-block row goes out as 8 16-byte chunks at the 128B-swizzled position
    // (chunk ^ (row & 7)).
    constexpr int kProd = kStemProducerWarps * 32;               // 256 producer threads
    const int ptid = threadIdx.x - (2 + kStemEpiWarps) * 32;
    const int pw = warp - (2 + kStemEpiWarps);                   // producer warp 0..7 (warp-uniform)
    const int r = ((pw & 3) << 5) | lane;                        // tile row = output column ox
    const int sub = pw >> 2;                                     // this warp builds chunks [4*sub, 4*sub + 4) of every k-block
    const int row_bytes = p.W_in * 3;
    const int pitch = 9 + row_bytes + 15;                       // staged words per input row (zero padded)
    const int q_per_row = row_bytes / 16;                       // 16-byte groups per input row (42 for W = 224)
    uint32_t* s_conv = reinterpret_cast<uint32_t*>(smem + STEM_OFF);
    uint32_t* s_lut = s_conv + 7 * pitch;
    if (!STEM_F32)
      for (int i = ptid; i < 768; i += kProd) s_lut[i] = __ldg(p.lut + i);
    for (int i = ptid; i < 7 * pitch; i += kProd) s_conv[i] = 0u;   // pads stay zero for the whole kernel
    asm volatile("bar.sync 1, 256;" ::: "memory");

    // 16-byte load groups per input row: 16 interleaved bytes (uint8 NHWC) or 4 floats of one channel plane (float32 NCHW)
    const int w4 = p.W_in / 4;
    const int g_per_row = STEM_F32 ? 3 * w4 : q_per_row;
    constexpr int kMaxIt = 2;                                    // 2 rows (F32) / 7 rows (uint8) fit in 2 groups per thread
    uint4 v[kMaxIt];
    int dstw[kMaxIt];                                            // staged word offset; -1 = nothing; <= -2 = zero row
    // issue the loads of input rows [iy0, iy0 + nrows) of image n_img
    auto fetch_rows = [&](int n_img, int iy0, int nrows) {
#pragma unroll
      for (int it = 0; it < kMaxIt; ++it) {
        const int idx = ptid + it * kProd;
        dstw[it] = -1;
        v[it] = make_uint4(0, 0, 0, 0);
        if (idx < nrows * g_per_row) {
          const int k = idx / g_per_row, g = idx - k * g_per_row, iy = iy0 + k;
          const int slot = ((iy + 3 + 7) % 7) * pitch + 9;
          int d;
          const uint4* src;
          if (STEM_F32) {
            const int c = g / w4, q = g - c * w4;                 // 4 consecutive pixels of channel c -> words 12q + c (+3 each)
            d = slot + 12 * q + c;
            src = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.img) +
                                                 (((long long)n_img * 3 + c) * p.H_in + iy) * p.W_in) + q;
          } else {
            d = slot + g * 16;
            src = reinterpret_cast<const uint4*>(p.img + ((long long)n_img * p.H_in + iy) * row_bytes) + g;
          }
          if (iy >= 0 && iy < p.H_in) { v[it] = __ldg(src); dstw[it] = d; }
          else dstw[it] = -d - 2;                                 // out-of-image row: store zeros
        }
      }
    };
    // loaded values -> normalised (hi | lo << 16) words in the ring
    auto stage_rows = [&]() {
#pragma unroll
      for (int it = 0; it < kMaxIt; ++it) {
        if (dstw[it] == -1) continue;
        const bool zero = dstw[it] < 0;
        const int d0 = zero ? -(dstw[it] + 2) : dstw[it];
        const uint32_t w4v[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
        const int c0 = (d0 - 9 - (d0 / pitch) * pitch) % 3;     // channel of the first element of this group
        if (STEM_F32) {
          const float mu = p.nmean[c0], sd = p.nstd[c0];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint16_t hh, ll;
            if (F16) { hh = __half_as_ushort(__float2half_rn((__uint_as_float(w4v[j]) - mu) / sd)); ll = 0; }
            else split_pair((__uint_as_float(w4v[j]) - mu) / sd, hh, ll);   // same arithmetic as b200r_stem_im2col_f32
            s_conv[d0 + 3 * j] = zero ? 0u : ((uint32_t)hh | ((uint32_t)ll << 16));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t b = (w4v[j >> 2] >> (8 * (j & 3))) & 0xFFu;
            const int c = (c0 + j) % 3;
            s_conv[d0 + j] = zero ? 0u : s_lut[c * 256 + b];
          }
        }
      }
    };
    // all 7 rows of a tile (first tile of the CTA; F32: also the first tile of every image), two rows at a time
    auto cold_start = [&](int n_img, int oy) {
      for (int k = 0; k < 7; k += 2) {
        fetch_rows(n_img, 2 * oy - 3 + k, (7 - k) < 2 ? (7 - k) : 2);
        stage_rows();
      }
    };

    int stage = 0;
    uint32_t phase = 0;
    if (t_first < t_end) cold_start(t_first / p.tiles_h, t_first % p.tiles_h);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    for (int t = t_first; t < t_end; ++t) {
      const int oy = t % p.tiles_h, n_img = t / p.tiles_h;
      // prefetch what the next tile adds
      const bool has_next = (t + 1 < t_end);
      const bool next_same_image = (oy + 1 < p.tiles_h);
      if (has_next) {
        if (next_same_image) fetch_rows(n_img, 2 * oy + 4, 2);
        else if (!STEM_F32) fetch_rows(n_img + 1, -3, 7);        // 7 uint8 rows still fit the two load slots
      }
      int rowoff[7];
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) rowoff[ky] = ((2 * oy + ky) % 7) * pitch + 6 * r;   // (2*ox - 3)*3 + 9 = 6*ox
#pragma unroll
      for (int kb = 0; kb < 3; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        uint8_t* a_hi = smem + stage * STAGE_BYTES + r * 128;
        uint8_t* a_lo = a_hi + A_TILE_BYTES;
        if (r < p.rows_box) {
#pragma unroll
          for (int q8 = 0; q8 < 8; ++q8) {
            if ((q8 >> 2) != sub) continue;                     // warp-uniform; q stays a compile-time constant
            const int q = q8;
            const int chunk = kb * 8 + q;                       // K chunk 0..23 = (ky, third of the 24-word run)
            uint32_t e[8];
            if (chunk < 21) {
              const uint2* src = reinterpret_cast<const uint2*>(s_conv + rowoff[chunk / 3] + 8 * (chunk % 3));
#pragma unroll
              for (int j = 0; j < 4; ++j) { const uint2 w2 = src[j]; e[2 * j] = w2.x; e[2 * j + 1] = w2.y; }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) e[j] = 0u;
            }
            uint4 h, l;
            h.x = __byte_perm(e[0], e[1], 0x5410); h.y = __byte_perm(e[2], e[3], 0x5410);
            h.z = __byte_perm(e[4], e[5], 0x5410); h.w = __byte_perm(e[6], e[7], 0x5410);
            l.x = __byte_perm(e[0], e[1], 0x7632); l.y = __byte_perm(e[2], e[3], 0x7632);
            l.z = __byte_perm(e[4], e[5], 0x7632); l.w = __byte_perm(e[6], e[7], 0x7632);
            const int pos = (q ^ (r & 7)) << 4;
            *reinterpret_cast<uint4*>(a_hi + pos) = h;
            if (!F16 && p.passes == 3) *reinterpret_cast<uint4*>(a_lo + pos) = l;
          }
        }
        fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
        mbar_arrive(full_bar(stage));
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");            // everyone done reading the ring rows this tile retires
      if (has_next) {
        if (STEM_F32 && !next_same_image) cold_start(n_img + 1, 0);   // image boundary: once per ~112 tiles
        else stage_rows();
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();       // the peer may still be reading this CTA's operands / signalling its barriers
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct GemmMaps { CUtensorMap a, b, r, i, y, y2; };

template <int BN, int STEM, bool F16, bool LEAN = false, int ACTS = 1, bool PAIR = false, bool DUAL = false>
int launch(const GemmMaps& m, const GemmParams& p, cudaStream_t s) {
  constexpr int kStages = PAIR ? 4 : (LEAN && F16) ? (BN > 128 ? 3 : (BN == 128 ? 5 : 6)) : ((BN > 128) ? 2 : (BN == 64 && !STEM ? 4 : kStagesDefault)) * (F16 ? 2 : 1);
  constexpr int STAGE_BYTES = (2 * A_TILE_BYTES + 2 * (PAIR ? BN / 2 : BN) * BK * 2) / (F16 ? 2 : 1);
  // STEM adds the staged input rows (7 x (W_in*3 + 24) words) and the 3 KB LUT behind the barriers
  const int smem = kStages * STAGE_BYTES + 1024 /*align slack*/ + 1024 /*barriers*/ +
                   ((LEAN && F16) ? 65536 : ((STEM || p.tma_store) ? 32768 : 0)) /*epilogue staging*/ + (STEM ? 7 * (p.W_in * 3 + 24) * 4 + 768 * 4 : 0);
  static int configured = 0;
  if (configured < smem) {
    B200R_CUDA((cudaFuncSetAttribute(gemm_kernel<BN, STEM, F16, LEAN, ACTS, PAIR, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)));
    configured = smem;
  }
  if constexpr (PAIR) {
    // one cluster of two CTAs per pair of M tiles; grid = an even number of CTAs, at most one per SM
    const int pairs = ((p.tiles_w * p.tiles_h * p.tiles_img + 1) / 2) * p.tiles_n;
    const int max_clusters = b200r_num_sms() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (pairs < max_clusters ? pairs : max_clusters));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200R_CUDA((cudaLaunchKernelEx(&cfg, gemm_kernel<BN, STEM, F16, LEAN, ACTS, PAIR, DUAL>, m.a, m.b, m.r, m.i, m.y, m.y2, p)));
    return B200R_OK;
  }
  const int total = p.tiles_w * p.tiles_h * p.tiles_img * p.tiles_n;
  const int grid = total < b200r_num_sms() ? total : b200r_num_sms();
  gemm_kernel<BN, STEM, F16, LEAN, ACTS, PAIR, DUAL><<<grid, STEM ? (2 + kStemEpiWarps + kStemProducerWarps) * 32 : kThreads, smem, s>>>(m.a, m.b, m.r, m.i, m.y, m.y2, p);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

// 5-D map over split planes [plane][N][H][W][C] (dims listed innermost first)
// (pw, ph, pn: pixel strides of w, h and n when the tensor is a strided view; 0 = dense)
int make_map5(EncodeTiledFn enc, CUtensorMap* m, const uint16_t* base, int C, int W, int H, int N, size_t plane_elems, int box_c,
              int box_w, int box_h, int box_n, int estride, CUtensorMapSwizzle swz, const char* what, bool f16, long long pw = 0,
              long long ph = 0, long long pn = 0) {
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, f16 ? 1u : 2u};
  cuuint64_t strides[4] = {(cuuint64_t)(pw ? pw : 1) * C * 2, (cuuint64_t)(ph ? ph : W) * C * 2, (cuuint64_t)(pn ? pn : (long long)H * W) * C * 2,
                           (cuuint64_t)plane_elems * 2};
  cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)(box_w * estride), (cuuint32_t)(box_h * estride), (cuuint32_t)box_n, 1};
  cuuint32_t estr[5] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<uint16_t*>(base), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b200r_set_error("cuTensorMapEncodeTiled(%s) failed: %d (N=%d H=%d W=%d C=%d box %d,%d,%d,%d s=%d)", what, (int)r, N, H, W, C, box_n,
                    box_h, box_w, box_c, estride);
    return B200R_ECUDA;
  }
  return B200R_OK;
}

// 256 x 256 identity in bf16 / fp16 (the B operand of the residual k-blocks), one per device and format
uint16_t* g_ident[8][2] = {};
std::mutex g_ident_mu;
int get_identity(const uint16_t** out, bool f16) {
  int dev = 0;
  B200R_CUDA(cudaGetDevice(&dev));
  B200R_CHECK_ARG(dev >= 0 && dev < 8, "device index out of range");
  std::lock_guard<std::mutex> lk(g_ident_mu);
  if (!g_ident[dev][f16]) {  // first use per device: blocking copy (not capturable)
    static uint16_t h[256 * 256];
    memset(h, 0, sizeof(h));
    for (int i = 0; i < 256; ++i) h[i * 256 + i] = 0x3C00;   // 1.0 (fp16 in both precisions)
    B200R_CUDA(cudaMalloc(&g_ident[dev][f16], sizeof(h)));
    B200R_CUDA(cudaMemcpy(g_ident[dev][f16], h, sizeof(h), cudaMemcpyHostToDevice));
  }
  *out = g_ident[dev][f16];
  return B200R_OK;
}

// output / residual / identity maps + flags shared by conv_impl and the fused stem
// tuning switch for experiments: B200R_GEMM_OPTS bit 0 = epilogue stores through the LSU instead of TMA,
// bit 1 = residual through the LSU instead of identity k-blocks, bit 2 = general epilogue instead of the LEAN one,
// bit 3 = one accumulator chain, bit 4 = no CTA pairs, bit 5 = CTA pairs wherever the kernel supports them
int gemm_opts() {        // read at every launch (an A/B test flips it inside one process); launches are graph-captured where it matters
  const char* e = getenv("B200R_GEMM_OPTS");
  return e ? atoi(e) : 0;
}

int finish_maps(EncodeTiledFn enc, GemmMaps* m, GemmParams* p, const uint16_t* res, uint16_t* y, size_t ycount, int BN, bool f16, bool pair = false) {
  // strided output view (GemmParams::o_*): the maps start at the view's first pixel and carry its strides
  const bool dense = p->o_off == 0 && p->o_w == 1 && p->o_h == p->Wo && p->o_img == (long long)p->Ho * p->Wo;
  const long long vw = dense ? 0 : p->o_w, vh = dense ? 0 : p->o_h, vn = dense ? 0 : p->o_img;
  const size_t voff = (size_t)p->o_off * p->Cout;
  p->tma_store = 0;
  p->res_mma = 0;
  m->r = m->b; m->i = m->b; m->y = m->b; m->y2 = m->b;   // placeholders (never dereferenced unless the flag is set)
  if (y && p->Cout % 8 == 0 && !(gemm_opts() & 1)) {
    // split-bf16: 32-channel (64-byte) rows per plane; fp16: 64-channel (128-byte) rows, half as many and twice as wide stores
    int rc = make_map5(enc, &m->y, y + voff, p->Cout, p->Wo, p->Ho, p->N, ycount, f16 ? 64 : 32, p->bw, p->bh, p->bn, 1,
                       f16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, "Y", f16, vw, vh, vn);
    if (rc) return rc;
    p->tma_store = 1;
  }
  if (res && !p->scale && p->Cout % 8 == 0 && !(gemm_opts() & 2)) {
    int rc = make_map5(enc, &m->r, res + voff, p->Cout, p->Wo, p->Ho, p->N, ycount, BK, p->bw, p->bh, p->bn, 1, CU_TENSOR_MAP_SWIZZLE_128B, "R", f16, vw, vh, vn);
    if (rc) return rc;
    const uint16_t* ident = nullptr;
    rc = get_identity(&ident, f16);
    if (rc) return rc;
    cuuint64_t dims[2] = {256, 256};
    cuuint64_t strides[1] = {512};
    cuuint32_t box[2] = {(cuuint32_t)BK, pair ? 32u : 64u};     // the 64 x 64 identity block (the kernel slides the accumulator columns instead); a CTA pair holds 32 rows each
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m->i, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(ident), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b200r_set_error("cuTensorMapEncodeTiled(identity) failed: %d", (int)r); return B200R_ECUDA; }
    p->res_mma = 1;
  }
  return B200R_OK;
}

// pick the A box (bn, bh, bw) in output pixels with bn*bh*bw <= 128 that wastes the fewest MMA rows
void choose_box(int N, int Ho, int Wo, int stride, int& bn, int& bh, int& bw) {
  double best = -1;
  bn = bh = 1; bw = 1;
  for (int w = 1; w <= Wo && w <= 128; ++w) {
    if (w * stride > 256) break;
    for (int h = 1; h <= Ho && h * w <= 128; ++h) {
      if (h * stride > 256) break;
      int n = 1;
      if (h == Ho && w == Wo) n = 128 / (h * w);
      if (n > N) n = N;
      if (n < 1) n = 1;
      const long tiles = (long)((Wo + w - 1) / w) * ((Ho + h - 1) / h) * ((N + n - 1) / n);
      const double util = (double)N * Ho * Wo / (tiles * 128.0);
      // prefer wide boxes on ties (longer contiguous runs per TMA row)
      const double score = util + 1e-6 * w;
      if (score > best) { best = score; bn = n; bh = h; bw = w; }
    }
  }
}

// Output as a strided view of a larger tensor: logical size Ho x Wo per image (overrides the convolution formula: taps beyond the
// input's edge read TMA zero fill), pixel (n, ho, wo) at n * img + ho * h + wo * w + off pixels, planes plane_elems apart.
struct OutView { int Ho, Wo; long long img, h, w, off; size_t plane_elems; };

int conv_impl(const uint16_t* x, const uint16_t* wgt, const float* scale, const float* bias, const uint16_t* res,
              uint16_t* y, float* y_f32, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
              int act, int passes, bool flat2d, cudaStream_t s, const uint16_t* mask = nullptr, const OutView* ov = nullptr,
              uint16_t* pre = nullptr) {
  B200R_CHECK_ARG(x && wgt && (y || y_f32), "null pointer");
  // K tails (Cin % 64 != 0) ride on TMA out-of-bounds zero fill of the activation's channel dimension
  B200R_CHECK_ARG(Cin % 8 == 0, "cin (%d) must be a multiple of 8 (16-byte TMA strides)", Cin);
  B200R_CHECK_ARG(passes == 1 || passes == 3 || passes == B200R_PASSES_F16, "passes must be 1, 3 or B200R_PASSES_F16");
  const bool f16 = (passes == B200R_PASSES_F16);
  if (f16) passes = 1;
  B200R_CHECK_ARG(stride >= 1 && stride <= 8 && KH >= 1 && KW >= 1 && pad >= 0, "bad conv geometry");
  EncodeTiledFn enc = get_encode();
  if (!enc) { b200r_set_error("cuTensorMapEncodeTiled not available from the driver"); return B200R_ECUDA; }
  const int Ho = ov ? ov->Ho : (H + 2 * pad - KH) / stride + 1, Wo = ov ? ov->Wo : (W + 2 * pad - KW) / stride + 1;
  B200R_CHECK_ARG(Ho > 0 && Wo > 0, "empty output");
  B200R_CHECK_ARG(!ov || (!y_f32 && !flat2d), "strided output views are for plane outputs of spatial convolutions");
  const size_t xcount = (size_t)N * H * W * Cin, wcount = (size_t)Cout * KH * KW * Cin, ycount = ov ? ov->plane_elems : (size_t)N * Ho * Wo * Cout;

  GemmParams p{};
  if (flat2d) { p.bn = 1; p.bh = 1; p.bw = 128; }
  else choose_box(N, Ho, Wo, stride, p.bn, p.bh, p.bw);
  p.rows_box = p.bn * p.bh * p.bw;
  p.tiles_w = (Wo + p.bw - 1) / p.bw; p.tiles_h = (Ho + p.bh - 1) / p.bh; p.tiles_img = (N + p.bn - 1) / p.bn;
  // BN = 256 halves the A-operand smem traffic per MMA (96 instead of 128 B/clk), worth it when the layer still
  // yields enough tiles to fill the SMs
  int BN = (Cout <= 64) ? 64 : 128;
  // split planes: two accumulator chains per tile (half the truncation bias of the tensor core's fp32 accumulation); they need
  // 4 x BN TMEM columns, so BN = 256 is left to the fp16 single-plane mode.  B200R_GEMM_OPTS bit 3 switches back.
  p.two_chains = (!f16 && !(gemm_opts() & 8)) ? 1 : 0;
  if (!p.two_chains) {
    const long m_tiles = (long)p.tiles_w * p.tiles_h * p.tiles_img;
    if (Cout % 256 == 0 && m_tiles * (Cout / 256) >= 2L * b200r_num_sms()) BN = 256;
  }
  // CTA pairs (cta_group::2) for the LEAN split-plane BN = 128 kernel: half the weight bytes per CTA (B200R_GEMM_OPTS bit 4 = off)
  const bool will_tma_store = y && Cout % 8 == 0 && !(gemm_opts() & 1);
  const bool will_res_mma = res && !scale && Cout % 8 == 0 && !(gemm_opts() & 2);
  const bool will_lean = will_tma_store && !y_f32 && (!res || will_res_mma) && !(gemm_opts() & 4);
  // ... where operand delivery, not HBM, is the bound: time at the HBM roofline over time at the tensor roofline, per output pixel,
  // = [4 B (Cin + Cout (1 + res)) / 6.5 TB/s] / [3 x 2 K Cout / 1353 TF/s]; measured layer by layer (ResNet-50, batch 256): pairs win
  // below ~1.3 and lose 5-10 % on the HBM-bound 56 x 56 1x1 layers above it (lock-step of the two CTAs, no operand to save)
  const double hbm_over_tensor = 138.8 * ((double)Cin + (double)Cout * (res ? 2.0 : 1.0)) / ((double)KH * KW * Cin * (double)Cout);
  const bool pair = will_lean && !f16 && BN <= 128 && (long)p.tiles_w * p.tiles_h * p.tiles_img >= 2 && b200r_num_sms() >= 2 &&
                    !(gemm_opts() & 16) && (hbm_over_tensor < 1.3 || (gemm_opts() & 32));
  p.tiles_n = (Cout + BN - 1) / BN;
  p.N = N; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout;
  p.o_img = ov ? ov->img : (long long)Ho * Wo; p.o_h = ov ? ov->h : Wo; p.o_w = ov ? ov->w : 1; p.o_off = ov ? ov->off : 0;
  p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad; p.cin_blocks = (Cin + 63) / 64; p.cin = Cin;
  p.passes = passes; p.act = act; p.scale = scale; p.bias = bias;
  p.res_hi = res; p.res_lo = (res && !f16) ? res + ycount : nullptr;
  p.mask_hi = mask;
  p.y_hi = y; p.y_lo = (y && !f16) ? y + ycount : nullptr; p.y_f32 = y_f32;

  GemmMaps m;
  {
    int rc = make_map5(enc, &m.a, x, Cin, W, H, N, xcount, BK, p.bw, p.bh, p.bn, stride, CU_TENSOR_MAP_SWIZZLE_128B, "A", f16);
    if (rc) return rc;
  }
  {
    const cuuint64_t K = (cuuint64_t)KH * KW * Cin;
    cuuint64_t dims[3] = {K, (cuuint64_t)Cout, f16 ? 1u : 2u};
    cuuint64_t strides[2] = {K * 2, (cuuint64_t)wcount * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)(pair ? BN / 2 : BN), 1};     // a CTA pair loads half the weight tile each
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&m.b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<uint16_t*>(wgt), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b200r_set_error("cuTensorMapEncodeTiled(B) failed: %d (Cout=%d K=%llu)", (int)r, Cout, (unsigned long long)K); return B200R_ECUDA; }
  }
  {
    int rc = finish_maps(enc, &m, &p, res, y, ycount, BN, f16, pair);
    if (rc) return rc;
  }
  // the compact epilogue: planes out through TMA stores (Cout % 8 == 0), residual on the tensor core or none, no fp32 output
  const bool lean = p.tma_store && y && !y_f32 && (!res || p.res_mma) && !(gemm_opts() & 4);
  const bool simple_act = (act == B200R_ACT_NONE || act == B200R_ACT_RELU);
#define B200R_LAUNCH_LEAN(F)                                                                                        \
  do {                                                                                                              \
    if (simple_act) {                                                                                               \
      if (BN == 256) return launch<256, 0, F, true, 0>(m, p, s);                                                    \
      return BN == 64 ? launch<64, 0, F, true, 0>(m, p, s) : launch<128, 0, F, true, 0>(m, p, s);                   \
    }                                                                                                               \
    if (BN == 256) return launch<256, 0, F, true, 1>(m, p, s);                                                      \
    return BN == 64 ? launch<64, 0, F, true, 1>(m, p, s) : launch<128, 0, F, true, 1>(m, p, s);                     \
  } while (0)
  if (f16) {
    if (lean) B200R_LAUNCH_LEAN(true);
    if (BN == 256) return launch<256, 0, true>(m, p, s);
    return BN == 64 ? launch<64, 0, true>(m, p, s) : launch<128, 0, true>(m, p, s);
  }
  if (pre) {
    // second output (the pre-activation planes): the DUAL instantiations of the LEAN split-plane kernel
    B200R_CHECK_ARG(lean && !f16 && !simple_act && !ov && BN <= 128, "pre-activation output: split precision, a non-trivial activation, Cout %% 8 == 0, dense planes");
    int rc = make_map5(enc, &m.y2, pre, Cout, Wo, Ho, N, ycount, 32, p.bw, p.bh, p.bn, 1, CU_TENSOR_MAP_SWIZZLE_64B, "Y2", false);
    if (rc) return rc;
    if (pair) return BN == 64 ? launch<64, 0, false, true, 1, true, true>(m, p, s) : launch<128, 0, false, true, 1, true, true>(m, p, s);
    return BN == 64 ? launch<64, 0, false, true, 1, false, true>(m, p, s) : launch<128, 0, false, true, 1, false, true>(m, p, s);
  }
  if (lean && pair) {
    if (BN == 64) return simple_act ? launch<64, 0, false, true, 0, true>(m, p, s) : launch<64, 0, false, true, 1, true>(m, p, s);
    return simple_act ? launch<128, 0, false, true, 0, true>(m, p, s) : launch<128, 0, false, true, 1, true>(m, p, s);
  }
  B200R_CHECK_ARG(!pair, "internal: pair launch planned for a non-LEAN configuration");
  if (lean) B200R_LAUNCH_LEAN(false);
#undef B200R_LAUNCH_LEAN
  if (BN == 256) return launch<256, 0, false>(m, p, s);
  return BN == 64 ? launch<64, 0, false>(m, p, s) : launch<128, 0, false>(m, p, s);
}

// ---- fused stem ---------------------------------------------------------------------------------
struct StemLut { uint32_t* d = nullptr; float key[6] = {0, 0, 0, 0, 0, 0}; };
StemLut g_stem_lut[8][2];
std::mutex g_stem_mu;

uint16_t host_f16(float v) { return __half_as_ushort(__float2half_rn(v)); }

int get_stem_lut(const float* mean, const float* stdv, const uint32_t** out, bool f16) {
  int dev = 0;
  B200R_CUDA(cudaGetDevice(&dev));
  B200R_CHECK_ARG(dev >= 0 && dev < 8, "device index out of range");
  std::lock_guard<std::mutex> lk(g_stem_mu);
  StemLut& L = g_stem_lut[dev][f16];
  bool same = L.d != nullptr;
  for (int i = 0; i < 3 && same; ++i) same = (L.key[i] == mean[i]) && (L.key[3 + i] == stdv[i]);
  if (!same) {  // first use per device (or new constants): host table + blocking copy (not capturable)
    uint32_t h[768];
    for (int c = 0; c < 3; ++c)
      for (int b = 0; b < 256; ++b) {
        const float v = ((float)b / 255.0f - mean[c]) / stdv[c];   // ToTensor + Normalize in fp32
        const uint16_t hi = host_f16(v), lo = f16 ? (uint16_t)0 : host_f16(v - __half2float(__ushort_as_half(hi)));
        h[c * 256 + b] = (uint32_t)hi | ((uint32_t)lo << 16);
      }
    if (!L.d) B200R_CUDA(cudaMalloc(&L.d, sizeof(h)));
    B200R_CUDA(cudaMemcpy(L.d, h, sizeof(h), cudaMemcpyHostToDevice));
    for (int i = 0; i < 3; ++i) { L.key[i] = mean[i]; L.key[3 + i] = stdv[i]; }
  }
  *out = L.d;
  return B200R_OK;
}

}  // namespace

extern "C" {

static int stem_impl(const void* img, bool f32, const uint16_t* wgt, const float* scale, const float* bias, uint16_t* y,
                     int n, int h, int w, const float* mean_host, const float* std_host, int act, int passes,
                     b200r_stream_t stream) {
  B200R_CHECK_ARG(img && wgt && y && mean_host && std_host, "null pointer");
  B200R_CHECK_ARG(n > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, "bad shape");
  B200R_CHECK_ARG(passes == 1 || passes == 3 || passes == B200R_PASSES_F16, "passes must be 1, 3 or B200R_PASSES_F16");
  const bool f16 = (passes == B200R_PASSES_F16);
  if (f16) passes = 1;
  EncodeTiledFn enc = get_encode();
  if (!enc) { b200r_set_error("cuTensorMapEncodeTiled not available from the driver"); return B200R_ECUDA; }
  const uint32_t* lut = nullptr;
  int rc = f32 ? B200R_OK : get_stem_lut(mean_host, std_host, &lut, f16);   // uint8: ToTensor + Normalize + split as a 3x256 table
  if (rc) return rc;
  const int Ho = h / 2, Wo = w / 2, Cout = 64, K = 192;
  GemmParams p{};
  B200R_CHECK_ARG(Wo <= 128 && w % 16 == 0, "fused stem supports input widths up to 256, multiples of 16");
  p.bn = 1; p.bh = 1; p.bw = Wo; p.rows_box = Wo;                    // one tile = one output row of one image
  p.M_total = (long long)n * Ho * Wo;
  p.tiles_w = 1; p.tiles_h = Ho; p.tiles_img = n; p.tiles_n = 1;
  p.N = n; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout;
  p.o_img = (long long)Ho * Wo; p.o_h = Wo; p.o_w = 1; p.o_off = 0;
  p.KH = 1; p.KW = 1; p.stride = 1; p.pad = 0; p.cin_blocks = K / 64; p.cin = K;
  p.passes = passes; p.act = act; p.scale = scale; p.bias = bias;
  p.y_hi = y; p.y_lo = f16 ? nullptr : y + (size_t)p.M_total * Cout;
  p.img = static_cast<const uint8_t*>(img); p.lut = lut; p.H_in = h; p.W_in = w;
  for (int i = 0; i < 3; ++i) { p.nmean[i] = mean_host[i]; p.nstd[i] = std_host[i]; }
  GemmMaps m;
  {
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)Cout, f16 ? 1u : 2u};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)Cout * K * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, 64, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&m.b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<uint16_t*>(wgt), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b200r_set_error("cuTensorMapEncodeTiled(stem B) failed: %d", (int)r); return B200R_ECUDA; }
  }
  m.a = m.b;
  p.stem_Ho = Ho; p.stem_Wo = Wo;   // the producer decodes (n, oy, ox) from the real output geometry
  rc = finish_maps(enc, &m, &p, nullptr, y, (size_t)p.M_total * Cout, 64, f16);
  if (rc) return rc;
  if (f16) return f32 ? launch<64, 2, true>(m, p, as_stream(stream)) : launch<64, 1, true>(m, p, as_stream(stream));
  return f32 ? launch<64, 2, false>(m, p, as_stream(stream)) : launch<64, 1, false>(m, p, as_stream(stream));
}

int b200r_stem_conv7x7_u8(const uint8_t* img, const uint16_t* wgt, const float* scale, const float* bias, uint16_t* y,
                          int n, int h, int w, const float* mean_host, const float* std_host, int act, int passes,
                          b200r_stream_t stream) {
  return stem_impl(img, false, wgt, scale, bias, y, n, h, w, mean_host, std_host, act, passes, stream);
}

int b200r_stem_conv7x7_f32(const float* img, const uint16_t* wgt, const float* scale, const float* bias, uint16_t* y,
                           int n, int h, int w, const float* mean_host, const float* std_host, int act, int passes,
                           b200r_stream_t stream) {
  return stem_impl(img, true, wgt, scale, bias, y, n, h, w, mean_host, std_host, act, passes, stream);
}

int b200r_conv2d_nhwc(const uint16_t* x, const uint16_t* wgt, const float* scale, const float* bias, const uint16_t* res,
                      uint16_t* y, float* y_f32, int n, int h, int w, int cin, int cout, int kh, int kw, int stride,
                      int pad, int act, int passes, b200r_stream_t stream) {
  B200R_CHECK_ARG(n > 0 && h > 0 && w > 0 && cin > 0 && cout > 0, "bad shape");
  // 1x1/s1 convolutions are plain GEMMs over flattened pixels: exact 128-row tiles
  const bool flat = (kh == 1 && kw == 1 && stride == 1 && pad == 0);
  if (flat) return conv_impl(x, wgt, scale, bias, res, y, y_f32, 1, 1, n * h * w, cin, cout, 1, 1, 1, 0, act, passes, true, as_stream(stream));
  return conv_impl(x, wgt, scale, bias, res, y, y_f32, n, h, w, cin, cout, kh, kw, stride, pad, act, passes, false, as_stream(stream));
}

int b200r_conv2d_dgrad_nhwc(const uint16_t* dy, const uint16_t* wgt_t, const uint16_t* res, const uint16_t* mask, uint16_t* dx,
                            int n, int h, int w, int cdy, int cdx, int kh, int kw, int pad, int passes, b200r_stream_t stream) {
  B200R_CHECK_ARG(n > 0 && h > 0 && w > 0 && cdy > 0 && cdx > 0 && dx, "bad shape");
  B200R_CHECK_ARG(h + 2 * pad - kh + 1 == h && w + 2 * pad - kw + 1 == w, "dgrad keeps the spatial size: pad must be (k-1)/2");
  const bool flat = (kh == 1 && kw == 1 && pad == 0);
  if (flat) return conv_impl(dy, wgt_t, nullptr, nullptr, res, dx, nullptr, 1, 1, n * h * w, cdy, cdx, 1, 1, 1, 0, B200R_ACT_NONE, passes, true, as_stream(stream), mask);
  return conv_impl(dy, wgt_t, nullptr, nullptr, res, dx, nullptr, n, h, w, cdy, cdx, kh, kw, 1, pad, B200R_ACT_NONE, passes, false, as_stream(stream), mask);
}

int b200r_conv2d_dgrad3x3s2_nhwc(const uint16_t* dy, const uint16_t* w00, const uint16_t* w01, const uint16_t* w10, const uint16_t* w11,
                                 const uint16_t* res, const uint16_t* mask, uint16_t* dx, int n, int ho, int wo, int cdy, int cdx,
                                 int passes, b200r_stream_t stream) {
  B200R_CHECK_ARG(dy && w00 && w01 && w10 && w11 && dx, "null pointer");
  B200R_CHECK_ARG(n > 0 && ho > 0 && wo > 0 && cdy > 0 && cdx > 0, "bad shape");
  // dx[2i + a, 2j + b] only meets the taps ky' = a + 1 (mod 2), kx' = b + 1 (mod 2) of the flipped kernel: parity class (a, b) is a
  // (1 + a) x (1 + b)-tap stride-1 convolution of dy itself -- 9 tap-GEMMs on the small map instead of 9 on the zero-dilated one
  const uint16_t* wsub[4] = {w00, w01, w10, w11};
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      OutView ov;
      ov.Ho = ho; ov.Wo = wo;
      ov.img = 4LL * ho * wo; ov.h = 4LL * wo; ov.w = 2; ov.off = (long long)a * 2 * wo + b;
      ov.plane_elems = (size_t)n * 4 * ho * wo * cdx;
      int rc = conv_impl(dy, wsub[a * 2 + b], nullptr, nullptr, res, dx, nullptr, n, ho, wo, cdy, cdx, 1 + a, 1 + b, 1, 0, B200R_ACT_NONE, passes,
                         false, as_stream(stream), mask, &ov);
      if (rc) return rc;
    }
  return B200R_OK;
}

int b200r_conv2d_dgrad1x1s2_acc_nhwc(const uint16_t* dy, const uint16_t* wgt_t, const uint16_t* mask, uint16_t* dx, int n, int h, int w,
                                     int cdy, int cdx, int passes, b200r_stream_t stream) {
  B200R_CHECK_ARG(dy && wgt_t && dx, "null pointer");
  B200R_CHECK_ARG(n > 0 && h > 0 && w > 0 && cdy > 0 && cdx > 0, "bad shape");
  B200R_CHECK_ARG(cdx % 8 == 0, "cdx must be a multiple of 8 (the accumulation rides on the tensor core's residual path)");
  // dx[:, 2i, 2j] += W^T dy[:, i, j] in place: output AND residual are the same strided view of dx (every tile reads its own rows
  // before it stores them; tiles are disjoint), so neither the zero-inserted gradient nor a separate add pass exists
  OutView ov;
  ov.Ho = (h - 1) / 2 + 1; ov.Wo = (w - 1) / 2 + 1;
  ov.img = (long long)h * w; ov.h = 2LL * w; ov.w = 2; ov.off = 0;
  ov.plane_elems = (size_t)n * h * w * cdx;
  return conv_impl(dy, wgt_t, nullptr, nullptr, dx, dx, nullptr, n, ov.Ho, ov.Wo, cdy, cdx, 1, 1, 1, 0, B200R_ACT_NONE, passes, false,
                   as_stream(stream), mask, &ov);
}

int b200r_linear(const uint16_t* x, const uint16_t* wgt, const float* scale, const float* bias, const uint16_t* res,
                 uint16_t* y, float* y_f32, int m, int k, int nout, int act, int passes, b200r_stream_t stream) {
  B200R_CHECK_ARG(m > 0 && k > 0 && nout > 0, "bad shape");
  return conv_impl(x, wgt, scale, bias, res, y, y_f32, 1, 1, m, k, nout, 1, 1, 1, 0, act, passes, true, as_stream(stream));
}

int b200r_linear_keep_pre(const uint16_t* x, const uint16_t* wgt, const float* bias, uint16_t* y, uint16_t* pre, int m, int k, int nout,
                          int act, void* stream) {
  B200R_CHECK_ARG(m > 0 && k > 0 && nout > 0, "bad shape");
  B200R_CHECK_ARG(y && pre && y != pre, "two distinct outputs");
  return conv_impl(x, wgt, nullptr, bias, nullptr, y, nullptr, 1, 1, m, k, nout, 1, 1, 1, 0, act, 3, true, as_stream(stream), nullptr, nullptr, pre);
}

}  // extern "C"
