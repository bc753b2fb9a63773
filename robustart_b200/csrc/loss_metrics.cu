// Logit-side kernels: softmax cross-entropy (+ gradient), softmax scores, top-1/top-5 counters.
// One warp per sample; [n, classes] float32 row-major.  Latency-bound (n x 4 KB).
//   F.cross_entropy (foolbox loss; imfgsm_attack.py:81), F.softmax (cls_solver.py:420),
//   accuracy() (prototype/prototype/utils/misc.py:441-455), ImageNetEvaluator.eval
//   (prototype/prototype/data/metrics/imagenet_evaluator.py:49-67).
#include "common.cuh"

namespace {
constexpr int kWarpsPerBlock = 8;

__global__ void __launch_bounds__(kWarpsPerBlock * 32) ce_kernel(const float* __restrict__ logits,
                                                                  const int64_t* __restrict__ labels,
                                                                  float* __restrict__ loss,
                                                                  float* __restrict__ dlogits, int n,
                                                                  int classes, float grad_scale,
                                                                  float* __restrict__ scores) {
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* z = logits + (size_t)row * classes;
  float mx = -INFINITY;
  for (int j = lane; j < classes; j += 32) mx = fmaxf(mx, z[j]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int j = lane; j < classes; j += 32) s += expf(z[j] - mx);
  s = warp_sum(s);
  const float inv = 1.0f / s;
  if (scores) {
    float* o = scores + (size_t)row * classes;
    for (int j = lane; j < classes; j += 32) o[j] = expf(z[j] - mx) * inv;
  }
  if (labels) {
    const int y = (int)labels[row];
    if (loss && lane == 0) loss[row] = (y >= 0 && y < classes) ? (logf(s) + mx - z[y]) : 0.f;
    if (dlogits) {
      float* d = dlogits + (size_t)row * classes;
      for (int j = lane; j < classes; j += 32) d[j] = (expf(z[j] - mx) * inv - (j == y ? 1.f : 0.f)) * grad_scale;
    }
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) topk_kernel(const float* __restrict__ logits,
                                                                    const int64_t* __restrict__ labels,
                                                                    int n, int classes,
                                                                    unsigned long long* __restrict__ counters,
                                                                    int64_t* __restrict__ pred) {
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  __shared__ unsigned int s_cnt[3];
  if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  if (row < n) {
    const float* z = logits + (size_t)row * classes;
    const int y = (int)labels[row];
    const bool valid = (y >= 0 && y < classes);
    const float zy = valid ? z[y] : INFINITY;
    int rank = 0;
    float best = -INFINITY;
    int besti = 0x7fffffff;
    for (int j = lane; j < classes; j += 32) {
      float v = z[j];
      rank += (v > zy) || (v == zy && j < y);   // stable order: lower index wins ties
      if (v > best) { best = v; besti = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      rank += __shfl_xor_sync(0xffffffffu, rank, o);
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if (lane == 0) {
      if (pred) pred[row] = besti;
      if (valid && rank < 1) atomicAdd(&s_cnt[0], 1u);
      if (valid && rank < 5) atomicAdd(&s_cnt[1], 1u);
      atomicAdd(&s_cnt[2], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < 3 && s_cnt[threadIdx.x]) atomicAdd(&counters[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}
}  // namespace

extern "C" {

int b200r_ce_loss_grad(const float* logits, const int64_t* labels, float* loss, float* dlogits, int n,
                       int classes, float grad_scale, b200r_stream_t stream) {
  B200R_CHECK_ARG(logits && labels, "null pointer");
  B200R_CHECK_ARG(n >= 0 && classes > 0, "bad shape");
  if (n == 0) return B200R_OK;
  ce_kernel<<<(n + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, as_stream(stream)>>>(
      logits, labels, loss, dlogits, n, classes, grad_scale, nullptr);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_softmax(const float* logits, float* scores, int n, int classes, b200r_stream_t stream) {
  B200R_CHECK_ARG(logits && scores, "null pointer");
  B200R_CHECK_ARG(n >= 0 && classes > 0, "bad shape");
  if (n == 0) return B200R_OK;
  ce_kernel<<<(n + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, as_stream(stream)>>>(
      logits, nullptr, nullptr, nullptr, n, classes, 1.f, scores);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

int b200r_topk_count(const float* logits, const int64_t* labels, int n, int classes, int64_t* counters,
                     int64_t* pred, b200r_stream_t stream) {
  B200R_CHECK_ARG(logits && labels && counters, "null pointer");
  B200R_CHECK_ARG(n >= 0 && classes > 0, "bad shape");
  if (n == 0) return B200R_OK;
  topk_kernel<<<(n + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, as_stream(stream)>>>(
      logits, labels, n, classes, reinterpret_cast<unsigned long long*>(counters), pred);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}

}  // extern "C"

// =============================================================================================
// DLR loss and its gradient w.r.t. the logits (autopgd_base.py:198-204 untargeted, :599-604 targeted).
//   untargeted: -(z_y - z_other) / (z_(1) - z_(3) + 1e-12), z_other = z_(2) if y is the top class else z_(1)
//   targeted:   -(z_y - z_t)     / (z_(1) - (z_(3) + z_(4))/2 + 1e-12)
// One warp per row finds the four largest logits (value, index), ties broken by index like a stable sort.
// =============================================================================================
namespace {
struct Top4 { float v[4]; int i[4]; };
__device__ __forceinline__ void top4_insert(Top4& t, float v, int i) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (v > t.v[k] || (v == t.v[k] && i > t.i[k])) {   // torch.sort ascending is stable: among equals the LAST index ranks highest
      float tv = t.v[k]; int ti = t.i[k];
      t.v[k] = v; t.i[k] = i; v = tv; i = ti;
    }
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) dlr_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                                                   const int64_t* __restrict__ targets, float* __restrict__ loss,
                                                                   float* __restrict__ dlogits, int n, int classes) {
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* z = logits + (size_t)row * classes;
  Top4 t;
#pragma unroll
  for (int k = 0; k < 4; ++k) { t.v[k] = -INFINITY; t.i[k] = -1; }
  for (int j = lane; j < classes; j += 32) top4_insert(t, z[j], j);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Top4 u;
#pragma unroll
    for (int k = 0; k < 4; ++k) { u.v[k] = __shfl_xor_sync(0xffffffffu, t.v[k], o); u.i[k] = __shfl_xor_sync(0xffffffffu, t.i[k], o); }
#pragma unroll
    for (int k = 0; k < 4; ++k) top4_insert(t, u.v[k], u.i[k]);
  }
  const int y = (int)labels[row];
  const float zy = z[y];
  float num, den;
  int a_idx, b_idx;   // numerator = -(z_a - z_b) with a = y
  int d1 = t.i[0], d2, d3 = -1;
  if (targets) {
    b_idx = (int)targets[row];
    num = -(zy - z[b_idx]);
    den = t.v[0] - 0.5f * (t.v[2] + t.v[3]) + 1e-12f;
    d2 = t.i[2]; d3 = t.i[3];
  } else {
    const bool top_is_y = (t.i[0] == y);
    b_idx = top_is_y ? t.i[1] : t.i[0];
    num = -(zy - (top_is_y ? t.v[1] : t.v[0]));
    den = t.v[0] - t.v[2] + 1e-12f;
    d2 = t.i[2];
  }
  a_idx = y;
  const float L = num / den;
  if (loss && lane == 0) loss[row] = L;
  if (dlogits) {
    float* d = dlogits + (size_t)row * classes;
    for (int j = lane; j < classes; j += 32) d[j] = 0.f;
    __syncwarp();
    if (lane == 0) {
      // L = num/den: dL = dnum/den - L*dden/den
      d[a_idx] += -1.f / den;
      d[b_idx] += 1.f / den;
      d[d1] += -L / den;
      if (targets) { d[d2] += 0.5f * L / den; d[d3] += 0.5f * L / den; }
      else d[d2] += L / den;
    }
  }
}
}  // namespace

extern "C" int b200r_dlr_loss_grad(const float* logits, const int64_t* labels, const int64_t* targets, float* loss, float* dlogits, int n,
                                   int classes, b200r_stream_t stream) {
  B200R_CHECK_ARG(logits && labels, "null pointer");
  B200R_CHECK_ARG(n >= 0 && classes >= 4, "DLR needs at least 4 classes");
  if (n == 0) return B200R_OK;
  dlr_kernel<<<(n + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, as_stream(stream)>>>(logits, labels, targets, loss, dlogits,
                                                                                                         n, classes);
  B200R_LAUNCH_CHECK();
  return B200R_OK;
}
