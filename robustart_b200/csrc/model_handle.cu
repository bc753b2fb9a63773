// Model handles behind the C-ABI: a C caller builds a classifier from the reference's state_dict
// tensors and runs forward / input-gradient passes without any Python.  The handle owns the device weights (BatchNorm
// folded, re-laid out, split into the precision's planes) and an activation arena; every layer is one launch of the
// library's own entry points (b200r_conv2d_nhwc, b200r_stem_*, b200r_linear, ...), in the order of
//   prototype/prototype/model/resnet_official.py:40-140 (BasicBlock / Bottleneck), :221-239 (stem, layers, avgpool, fc),
//   :330-346 (_forward_impl)
// and the input gradient is the reverse chain (what autograd.grad(loss, x) returns to the attacks:
// Attacks/autoattack/autopgd_base.py:371-376; foolbox value_and_grad behind adv/attack.py:20-33).
// robustart_b200/nets.ResNet issues the same launches from Python; tests/test_model_handle_gpu.py checks that both give
// bit-identical logits, tests/c/test_model_handle.c drives this file from plain C.
// Also here: b200r_allreduce_counts, the evaluation's one collective (int64 counter sum over an ncclComm_t).
#include "common.cuh"
#include <cuda_fp16.h>
#include <dlfcn.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

namespace {

constexpr double kBnEps = 1e-5;          // nn.BatchNorm2d default (misc.py:115-143 get_bn)
constexpr float kGradScale = 4096.0f;    // loss scale of the gradient pass (planes are fp16-coded)
const float kMean[3] = {0.485f, 0.456f, 0.406f}, kStd[3] = {0.229f, 0.224f, 0.225f};

struct Planes {              // a tensor in the precision's plane format on the device
  uint16_t* p = nullptr;
  size_t count = 0;          // elements per plane
};

struct ConvBN {
  Planes w, wt;              // forward weight [cout][kh][kw][cin]; dgrad weight [cin][kh][kw][cout] (transposed + flipped)
  Planes wt_s2[4];           // 3x3 / stride 2: the dgrad weight's parity-class sub-kernels [cin][1+a][1+b][cout] (b200r_conv2d_dgrad3x3s2_nhwc)
  float* bias = nullptr;
  int cin = 0, cout = 0, k = 1, stride = 1, pad = 0;
};
struct Block {
  bool bottleneck = false, has_down = false;
  ConvBN c1, c2, c3, down;
};

// Bump allocator over one device buffer.  A forward that must keep its activations (forward_f32, for the gradient pass) bumps
// through the whole buffer; an inference forward alternates between the two halves block by block (a block reads its input from
// one half and puts its output and temporaries into the other), so its footprint is two blocks, not the network.
struct Arena {
  uint8_t* base = nullptr;
  size_t cap = 0, off = 0, limit = 0;
  void reset() { off = 0; limit = cap; }
  void region(int r) { off = r ? cap / 2 : 0; limit = r ? cap : cap / 2; }
  void* take(size_t bytes) {
    off = (off + 255) & ~(size_t)255;
    if (off + bytes > limit) return nullptr;
    void* p = base + off;
    off += bytes;
    return p;
  }
};

}  // namespace

struct VitNet;                 // ViT-B/16 (vision_transformer.py): defined below, owned by the handle
struct MixerNet;               // MLP-Mixer-B/16 (vit/mlp_mixer.py, vit/vit_base.py)
struct MobileNet;              // MobileNetV2 / EfficientNet-B0 (mobilenet_v2.py, efficientnet.py): inference forward

struct b200r_model {
  VitNet* vit = nullptr;
  MixerNet* mixer = nullptr;
  MobileNet* mobile = nullptr;
  int arch = 0, passes = 3, planes = 2, classes = 1000, feat = 512;
  bool f16 = false;
  std::vector<Block> blocks;
  Planes stem_w, stem_w_folded, stem_wt, fc_w, fc_wt;
  uint16_t* stem_wp = nullptr;   // split mode: prepared operand of the one-launch stem [2][64][224]
  float stem_osc = 1.f;
  uint16_t* stem_wpf = nullptr;  // ... and of its float-input twin (b200r_stem_pool_f32_split)
  float stem_oscf = 1.f;
  float *stem_scale = nullptr, *stem_bias = nullptr, *fc_b = nullptr;
  std::vector<void*> owned;          // every cudaMalloc of the weights
  Arena arena;
  // state of the last forward_f32 (for input_grad)
  struct Saved { uint16_t* stem = nullptr; void* pool_codes = nullptr; std::vector<std::vector<uint16_t*>> blocks; int n = 0, h = 0, w = 0; bool valid = false; } saved;
};

namespace {

using Weights = std::map<std::string, std::pair<const float*, int64_t>>;

int dev_alloc(b200r_model* m, void** p, size_t bytes) {
  B200R_CUDA(cudaMalloc(p, bytes));
  m->owned.push_back(*p);
  return B200R_OK;
}

// host float32 -> device planes of the handle's precision (split: hi = rn16(v), lo = rn16(v - hi); f16: hi only)
int upload_planes(b200r_model* m, const std::vector<float>& v, Planes* out) {
  const size_t n = v.size();
  std::vector<uint16_t> h((size_t)m->planes * n);
  for (size_t i = 0; i < n; ++i) {
    const __half hi = __float2half_rn(v[i]);
    h[i] = __half_as_ushort(hi);
    if (m->planes == 2) h[n + i] = __half_as_ushort(__float2half_rn(v[i] - __half2float(hi)));
  }
  void* d = nullptr;
  int rc = dev_alloc(m, &d, h.size() * 2);
  if (rc) return rc;
  B200R_CUDA(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  out->p = static_cast<uint16_t*>(d);
  out->count = n;
  return B200R_OK;
}
// the value the planes hold (hi + lo in float), as the Python mirror's from_planes() sees it
float planes_value(float v, int planes) {
  const __half hi = __float2half_rn(v);
  if (planes == 1) return __half2float(hi);
  return __half2float(hi) + __half2float(__float2half_rn(v - __half2float(hi)));
}
int upload_f32(b200r_model* m, const std::vector<float>& v, float** out) {
  void* d = nullptr;
  int rc = dev_alloc(m, &d, v.size() * 4);
  if (rc) return rc;
  B200R_CUDA(cudaMemcpy(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  *out = static_cast<float*>(d);
  return B200R_OK;
}

int get(const Weights& W, const std::string& key, int64_t numel, const float** out) {
  auto it = W.find(key);
  B200R_CHECK_ARG(it != W.end(), "state_dict tensor '%s' is missing", key.c_str());
  B200R_CHECK_ARG(it->second.second == numel, "state_dict tensor '%s' has %lld elements, expected %lld", key.c_str(),
                  (long long)it->second.second, (long long)numel);
  *out = it->second.first;
  return B200R_OK;
}

// folded-BN scale (double) and bias (float) of a BatchNorm2d in eval mode
int fold_bn(const Weights& W, const std::string& bn, int c, std::vector<double>* scale, std::vector<float>* bias) {
  const float *g, *b, *mu, *var;
  int rc;
  if ((rc = get(W, bn + ".weight", c, &g)) || (rc = get(W, bn + ".bias", c, &b)) || (rc = get(W, bn + ".running_mean", c, &mu)) ||
      (rc = get(W, bn + ".running_var", c, &var)))
    return rc;
  scale->resize(c);
  bias->resize(c);
  for (int i = 0; i < c; ++i) {
    (*scale)[i] = (double)g[i] / sqrt((double)var[i] + kBnEps);
    (*bias)[i] = (float)((double)b[i] - (double)mu[i] * (*scale)[i]);
  }
  return B200R_OK;
}

// conv (bias-free) + BN: weights [cout][cin][k][k] (PyTorch) -> folded, [cout][k][k][cin] planes (+ the dgrad layout)
int make_conv(b200r_model* m, const Weights& W, const std::string& conv, const std::string& bn, int cin, int cout, int k, int stride, int pad,
              ConvBN* out) {
  const float* w;
  int rc = get(W, conv + ".weight", (int64_t)cout * cin * k * k, &w);
  if (rc) return rc;
  std::vector<double> scale;
  std::vector<float> bias;
  if ((rc = fold_bn(W, bn, cout, &scale, &bias))) return rc;
  std::vector<float> f((size_t)cout * k * k * cin), t((size_t)cin * k * k * cout);
  for (int o = 0; o < cout; ++o)
    for (int i = 0; i < cin; ++i)
      for (int y = 0; y < k; ++y)
        for (int x = 0; x < k; ++x) {
          const float v = (float)((double)w[(((size_t)o * cin + i) * k + y) * k + x] * scale[o]);   // BN scale folded (fp64 product)
          f[(((size_t)o * k + y) * k + x) * cin + i] = v;
          t[(((size_t)i * k + (k - 1 - y)) * k + (k - 1 - x)) * cout + o] = v;                       // transposed in (cout, cin), flipped in (ky, kx)
        }
  out->cin = cin; out->cout = cout; out->k = k; out->stride = stride; out->pad = pad;
  if ((rc = upload_planes(m, f, &out->w)) || (rc = upload_planes(m, t, &out->wt)) || (rc = upload_f32(m, bias, &out->bias))) return rc;
  if (k == 3 && stride == 2 && pad == 1) {
    const int taps[2][2] = {{1, 1}, {0, 2}};      // S_0 = {1}, S_1 = {0, 2}
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        const int kh = 1 + a, kw = 1 + b;
        std::vector<float> sub((size_t)cin * kh * kw * cout);
        for (int i = 0; i < cin; ++i)
          for (int y = 0; y < kh; ++y)
            for (int x = 0; x < kw; ++x)
              for (int o = 0; o < cout; ++o)
                sub[(((size_t)i * kh + y) * kw + x) * cout + o] = t[(((size_t)i * 3 + taps[a][y]) * 3 + taps[b][x]) * cout + o];
        if ((rc = upload_planes(m, sub, &out->wt_s2[a * 2 + b]))) return rc;
      }
  }
  return B200R_OK;
}

// conv1 weight [64][3][7][7] -> [64][192]: column = ky*24 + kx*3 + c, every ky run padded from 7 to 8 taps with zeros
std::vector<float> pack_stem(const std::vector<float>& w) {
  std::vector<float> out((size_t)64 * 192, 0.f);
  for (int o = 0; o < 64; ++o)
    for (int c = 0; c < 3; ++c)
      for (int y = 0; y < 7; ++y)
        for (int x = 0; x < 7; ++x) out[(size_t)o * 192 + y * 24 + x * 3 + c] = w[(((size_t)o * 3 + c) * 7 + y) * 7 + x];
  return out;
}

int conv_fwd(b200r_model* m, const ConvBN& c, const uint16_t* x, const uint16_t* res, uint16_t* y, int n, int h, int w, int act, cudaStream_t s) {
  return b200r_conv2d_nhwc(x, c.w.p, nullptr, c.bias, res, y, nullptr, n, h, w, c.cin, c.cout, c.k, c.k, c.stride, c.pad, act, m->passes,
                           reinterpret_cast<b200r_stream_t>(s));
}

uint16_t* take_planes(b200r_model* m, size_t count) { return static_cast<uint16_t*>(m->arena.take(count * 2 * m->planes)); }

#define TAKE(ptr, count)                                                                                                            \
  uint16_t* ptr = take_planes(m, (count));                                                                                          \
  if (!ptr) { b200r_set_error("activation arena too small: call b200r_model_reserve with this batch / image size first"); return B200R_EINVAL; }
#define RC(call) { int rc__ = (call); if (rc__) return rc__; }

size_t vit_arena_need(const b200r_model* m, int n, bool save);
size_t mixer_arena_need(const b200r_model* m, int n, bool save);
size_t mobile_arena_need(const b200r_model* m, int n, int h, int w);
size_t arena_need(const b200r_model* m, int n, int h, int w, bool save) {
  if (m->mobile) return mobile_arena_need(m, n, h, w);
  if (m->vit) return vit_arena_need(m, n, save);
  if (m->mixer) return mixer_arena_need(m, n, save);
  // elements per input pixel, bounded from above by a closed form.  Inference: two halves, each the largest block footprint
  // (ResNet-50 layer1.0: 56^2 x (256 + 256 + 64 + 64) = 40 per pixel; the stem: 112^2 x 64 + 56^2 x 64 = 20).  Saved forward +
  // gradient pass: every activation (ResNet-50: ~220 per pixel) + the gradient pass's own tensors (~280, of which 48 for the stem
  // GEMM's 192-column output): 560 (ResNet-101: 1000); basic-block nets: 300.
  double per_px = 2 * 48.0;
  if (save) per_px = m->feat == 2048 ? (m->arch == B200R_ARCH_RESNET101 ? 1000.0 : 560.0) : 300.0;
  return (size_t)(per_px * n * h * w * 2.0 * m->planes) + (64u << 20);
}

int ensure_arena(b200r_model* m, int n, int h, int w, bool save) {
  const size_t need = arena_need(m, n, h, w, save);
  if (m->arena.cap >= need) return B200R_OK;
  if (m->arena.base) B200R_CUDA(cudaFree(m->arena.base));
  m->arena.base = nullptr; m->arena.cap = 0;
  void* p = nullptr;
  B200R_CUDA(cudaMalloc(&p, need));            // not capturable: reserve before graph capture
  m->arena.base = static_cast<uint8_t*>(p);
  m->arena.cap = need;
  return B200R_OK;
}

// forward from the stem output (post maxpool) to the logits; keeps the activations when `save`
int run_body(b200r_model* m, uint16_t* x, int n, int h, int w, float* logits, bool save, cudaStream_t s) {
  const b200r_stream_t st = reinterpret_cast<b200r_stream_t>(s);
  int c = 64, bi = 0;
  if (save) m->saved.blocks.clear();
  for (const Block& b : m->blocks) {
    if (!save) m->arena.region(bi++ & 1);         // the stem wrote into half 1 (see the callers): block 0 -> half 0, block 1 -> half 1, ...
    const int stride = b.bottleneck ? b.c2.stride : b.c1.stride;
    const int ho = h / stride, wo = w / stride;
    const int cout = b.bottleneck ? b.c3.cout : b.c2.cout;
    const uint16_t* idn = x;
    if (b.has_down) {
      TAKE(d, (size_t)n * ho * wo * cout);
      RC(conv_fwd(m, b.down, x, nullptr, d, n, h, w, B200R_ACT_NONE, s));
      idn = d;
    }
    uint16_t* y;
    if (b.bottleneck) {
      TAKE(a1, (size_t)n * h * w * b.c1.cout);
      RC(conv_fwd(m, b.c1, x, nullptr, a1, n, h, w, B200R_ACT_RELU, s));
      TAKE(a2, (size_t)n * ho * wo * b.c2.cout);
      RC(conv_fwd(m, b.c2, a1, nullptr, a2, n, h, w, B200R_ACT_RELU, s));
      TAKE(o, (size_t)n * ho * wo * cout);
      RC(conv_fwd(m, b.c3, a2, idn, o, n, ho, wo, B200R_ACT_RELU, s));
      y = o;
      if (save) m->saved.blocks.push_back({a1, a2, y});
    } else {
      TAKE(a1, (size_t)n * ho * wo * b.c1.cout);
      RC(conv_fwd(m, b.c1, x, nullptr, a1, n, h, w, B200R_ACT_RELU, s));
      TAKE(o, (size_t)n * ho * wo * cout);
      RC(conv_fwd(m, b.c2, a1, idn, o, n, ho, wo, B200R_ACT_RELU, s));
      y = o;
      if (save) m->saved.blocks.push_back({a1, y});
    }
    x = y; h = ho; w = wo; c = cout;
  }
  if (!save) m->arena.region(bi & 1);
  TAKE(pooled, (size_t)n * c);
  RC((m->f16 ? b200r_global_avgpool_nhwc_f16 : b200r_global_avgpool_nhwc)(x, pooled, n, h * w, c, st));
  RC(b200r_linear(pooled, m->fc_w.p, nullptr, m->fc_b, nullptr, nullptr, logits, n, c, m->classes, B200R_ACT_NONE, m->passes, st));
  return B200R_OK;
}

// ---- ViT-B/16 (prototype/prototype/model/vision_transformer.py:44-349): the launch sequence of nets.ViT --------------------------
struct Lin { Planes w, wt; float* b = nullptr; int k = 0, nout = 0; };      // nn.Linear: w [nout][k], wt [k][nout] for the input gradient
struct VitBlock { float *n1w, *n1b, *n2w, *n2b; Lin qkv, out, m1, m2; };
struct VitSaved { uint16_t *x_in, *qkv, *x_mid, *pre; };
}  // namespace
struct VitNet {
  int depth = 12, dim = 768, heads = 12, patch = 16, hidden = 3072, tokens = 197, rep = 768;
  Lin embed, head, pre;
  bool has_pre = false;
  float *cls = nullptr, *pos = nullptr, *normw = nullptr, *normb = nullptr;
  std::vector<VitBlock> blocks;
  std::vector<VitSaved> saved;
  uint16_t *final_x = nullptr, *pre_logits = nullptr;
  int n = 0, h = 0, w = 0;
  bool valid = false;
};
namespace {

// k_pad / n_pad: the weight is embedded in a zero [n_pad][k_pad] matrix (Mixer's token dimension padded to 256: 16-byte rows, K % 64)
int make_lin(b200r_model* m, const Weights& W, const std::string& name, int nout0, int k0, Lin* out, int n_pad = 0, int k_pad = 0) {
  const float *w, *b;
  int rc;
  if ((rc = get(W, name + ".weight", (int64_t)nout0 * k0, &w)) || (rc = get(W, name + ".bias", nout0, &b))) return rc;
  const int nout = n_pad ? n_pad : nout0, k = k_pad ? k_pad : k0;
  std::vector<float> wv((size_t)nout * k, 0.f), wt(wv.size()), bv(nout, 0.f);
  for (int o = 0; o < nout0; ++o) {
    bv[o] = b[o];
    for (int j = 0; j < k0; ++j) wv[(size_t)o * k + j] = w[(size_t)o * k0 + j];
  }
  for (int o = 0; o < nout; ++o)
    for (int j = 0; j < k; ++j) wt[(size_t)j * nout + o] = planes_value(wv[(size_t)o * k + j], 2);   // transpose of the value the planes hold
  out->k = k; out->nout = nout;
  if ((rc = upload_planes(m, wv, &out->w)) || (rc = upload_planes(m, wt, &out->wt)) || (rc = upload_f32(m, bv, &out->b))) return rc;
  return B200R_OK;
}
int upload_vec(b200r_model* m, const Weights& W, const std::string& key, int64_t numel, float** out) {
  const float* v;
  int rc = get(W, key, numel, &v);
  if (rc) return rc;
  return upload_f32(m, std::vector<float>(v, v + numel), out);
}

int vit_create(b200r_model* m, const Weights& W) {
  VitNet* v = new VitNet();
  m->vit = v;
  const int D = v->dim, P = v->patch;
  int rc;
  auto pw = W.find("pos_embedding");
  B200R_CHECK_ARG(pw != W.end() && pw->second.second % D == 0, "state_dict tensor 'pos_embedding' is missing or not a multiple of %d", D);
  v->tokens = (int)(pw->second.second / D);
  if ((rc = make_lin(m, W, "embedding", D, 3 * P * P, &v->embed)) ||                 // conv weight [D, 3, P, P] == Linear over (c, ky, kx)
      (rc = upload_vec(m, W, "cls_token", D, &v->cls)) || (rc = upload_vec(m, W, "pos_embedding", (int64_t)v->tokens * D, &v->pos)))
    return rc;
  v->blocks.resize(v->depth);
  for (int d = 0; d < v->depth; ++d) {
    const std::string p = "transformer.encoders.encoder_" + std::to_string(d) + ".";
    VitBlock& b = v->blocks[d];
    if ((rc = upload_vec(m, W, p + "norm1.weight", D, &b.n1w)) || (rc = upload_vec(m, W, p + "norm1.bias", D, &b.n1b)) ||
        (rc = upload_vec(m, W, p + "norm2.weight", D, &b.n2w)) || (rc = upload_vec(m, W, p + "norm2.bias", D, &b.n2b)) ||
        (rc = make_lin(m, W, p + "attention.to_qkv", 3 * D, D, &b.qkv)) || (rc = make_lin(m, W, p + "attention.to_out", D, D, &b.out)) ||
        (rc = make_lin(m, W, p + "feedforward.mlp1", v->hidden, D, &b.m1)) || (rc = make_lin(m, W, p + "feedforward.mlp2", D, v->hidden, &b.m2)))
      return rc;
  }
  if ((rc = upload_vec(m, W, "transformer.encoder_norm.weight", D, &v->normw)) || (rc = upload_vec(m, W, "transformer.encoder_norm.bias", D, &v->normb)))
    return rc;
  v->has_pre = W.count("pre_logits.weight") != 0;
  if (v->has_pre) {
    v->rep = (int)(W.at("pre_logits.weight").second / D);
    if ((rc = make_lin(m, W, "pre_logits", v->rep, D, &v->pre))) return rc;
  }
  auto hw = W.find("head.weight");
  B200R_CHECK_ARG(hw != W.end(), "state_dict tensor 'head.weight' is missing");
  m->classes = (int)(hw->second.second / v->rep);
  return make_lin(m, W, "head", m->classes, v->rep, &v->head);
}

size_t vit_arena_need(const b200r_model* m, int n, bool save) {
  const VitNet* v = m->vit;
  const double tok = (double)n * v->tokens;
  // planes elements per token: saved forward keeps (x_in, qkv, x_mid, pre) = 7 D + hidden per block plus ln / attention / act outputs;
  // the gradient pass adds ~3 (D + hidden) of temporaries per block.  Inference: one block's worth, twice.
  const double per_tok = save ? v->depth * (8.0 * v->dim + 2.0 * v->hidden) + 4.0 * v->hidden + 16.0 * v->dim : 2.0 * (10.0 * v->dim + 2.0 * v->hidden);
  return (size_t)(per_tok * tok * 4.0) + (size_t)n * 3 * 224 * 224 * 8 + (64u << 20);
}

int lin_fwd(const Lin& l, const uint16_t* x, const uint16_t* res, uint16_t* y, float* y_f32, int rows, int act, int passes, b200r_stream_t st) {
  return b200r_linear(x, l.w.p, nullptr, l.b, res, y, y_f32, rows, l.k, l.nout, act, passes, st);
}
// pre = Linear(x), y = act(pre), both kept: one launch in split precision, Linear + activation pass in the fp16 mode
int lin_keep_pre(const Lin& l, const uint16_t* x, uint16_t* y, uint16_t* pre, int rows, int act, int passes, b200r_stream_t st) {
  if (passes == 3) return b200r_linear_keep_pre(x, l.w.p, l.b, y, pre, rows, l.k, l.nout, act, st);
  int rc = b200r_linear(x, l.w.p, nullptr, l.b, nullptr, pre, nullptr, rows, l.k, l.nout, B200R_ACT_NONE, passes, st);
  if (rc) return rc;
  return b200r_act_planes(pre, y, (size_t)rows * l.nout, act, st);
}
int lin_dgrad(const Lin& l, const uint16_t* g, uint16_t* dx, int rows, int passes, b200r_stream_t st) {
  return b200r_linear(g, l.wt.p, nullptr, nullptr, nullptr, dx, nullptr, rows, l.nout, l.k, B200R_ACT_NONE, passes, st);
}
// rows [first, first + count) of every image's token block: planes [2][n][T][D] -> planes [2][n][count][D] (class token / patch tokens)
int token_rows(const uint16_t* x, uint16_t* y, int n, int T, int D, int first, int count, bool scatter, cudaStream_t s) {
  for (int p = 0; p < 2; ++p) {
    const uint16_t* src = x + (size_t)p * n * (scatter ? count : T) * D + (scatter ? 0 : (size_t)first * D);
    uint16_t* dst = y + (size_t)p * n * (scatter ? T : count) * D + (scatter ? (size_t)first * D : 0);
    B200R_CUDA(cudaMemcpy2DAsync(dst, (size_t)(scatter ? T : count) * D * 2, src, (size_t)(scatter ? count : T) * D * 2, (size_t)count * D * 2, n,
                                 cudaMemcpyDeviceToDevice, s));
  }
  return B200R_OK;
}

// save == false: the inference sequence of nets.ViT.forward (fused GELU); save == true: nets.ViT.forward_saved
int vit_forward(b200r_model* m, const void* images, bool u8, float* logits, int n, int h, int w, bool save, cudaStream_t s) {
  VitNet* v = m->vit;
  const b200r_stream_t st = reinterpret_cast<b200r_stream_t>(s);
  const int D = v->dim, P = m->passes, pp = v->patch, np = (h / pp) * (w / pp), T = np + 1, hd = D / v->heads;
  B200R_CHECK_ARG(T == v->tokens, "ViT: %d x %d images give %d tokens, the position embedding has %d", h, w, T, v->tokens);
  B200R_CHECK_ARG(m->planes == 2, "ViT handles run in split precision (passes = 3)");
  const float scale = 1.0f / sqrtf((float)hd);
  m->arena.reset();
  const int rows = n * T;
  TAKE(cols, (size_t)n * np * 3 * pp * pp);
  RC(u8 ? b200r_patch_gather_u8(static_cast<const uint8_t*>(images), cols, n, h, w, pp, kMean, kStd, st)
        : b200r_patch_gather_f32(static_cast<const float*>(images), cols, n, h, w, pp, kMean, kStd, st));
  TAKE(emb, (size_t)n * np * D);
  RC(lin_fwd(v->embed, cols, nullptr, emb, nullptr, n * np, B200R_ACT_NONE, P, st));
  TAKE(x0, (size_t)rows * D);
  RC(b200r_assemble_tokens(emb, v->cls, v->pos, x0, n, np, D, st));
  uint16_t* x = x0;
  if (save) { v->saved.clear(); v->valid = false; }
  // inference: two x buffers alternate, every block reuses the same scratch behind them
  uint16_t* xalt = nullptr;
  size_t scratch_mark = 0;
  if (!save) {
    xalt = take_planes(m, (size_t)rows * D);
    B200R_CHECK_ARG(xalt, "activation arena too small");
    scratch_mark = m->arena.off;
  }
  for (const VitBlock& b : v->blocks) {
    if (!save) m->arena.off = scratch_mark;
    TAKE(y1, (size_t)rows * D);
    RC(b200r_layernorm(x, y1, b.n1w, b.n1b, rows, D, 1e-5f, st));
    TAKE(qkv, (size_t)rows * 3 * D);
    RC(lin_fwd(b.qkv, y1, nullptr, qkv, nullptr, rows, B200R_ACT_NONE, P, st));
    TAKE(att, (size_t)rows * D);
    RC(b200r_attention(qkv, att, n, T, v->heads, hd, scale, st));
    TAKE(x_mid, (size_t)rows * D);
    RC(lin_fwd(b.out, att, x, x_mid, nullptr, rows, B200R_ACT_NONE, P, st));            // x_mid = attn(norm1(x)) + x
    TAKE(y2, (size_t)rows * D);
    RC(b200r_layernorm(x_mid, y2, b.n2w, b.n2b, rows, D, 1e-5f, st));
    TAKE(hbuf, (size_t)rows * v->hidden);
    uint16_t* xn;
    if (save) {
      TAKE(act, (size_t)rows * v->hidden);
      RC(lin_keep_pre(b.m1, y2, act, hbuf, rows, B200R_ACT_GELU_TANH, P, st));          // the pre-activation is kept
      TAKE(xo, (size_t)rows * D);
      RC(lin_fwd(b.m2, act, x_mid, xo, nullptr, rows, B200R_ACT_NONE, P, st));
      v->saved.push_back({x, qkv, x_mid, hbuf});
      xn = xo;
    } else {
      RC(lin_fwd(b.m1, y2, nullptr, hbuf, nullptr, rows, B200R_ACT_GELU_TANH, P, st));  // tanh-approximation GELU (:19-37) in the epilogue
      xn = (x == x0) ? xalt : x0;
      RC(lin_fwd(b.m2, hbuf, x_mid, xn, nullptr, rows, B200R_ACT_NONE, P, st));
    }
    x = xn;
  }
  if (!save) m->arena.off = scratch_mark;
  TAKE(xn, (size_t)rows * D);
  RC(b200r_layernorm(x, xn, v->normw, v->normb, rows, D, 1e-5f, st));
  TAKE(cls, (size_t)n * D);
  RC(token_rows(xn, cls, n, T, D, 0, 1, false, s));                                      // x[:, 0]
  const uint16_t* feat = cls;
  v->pre_logits = nullptr;
  if (v->has_pre) {
    TAKE(pl, (size_t)n * v->rep);
    if (save) {
      RC(lin_fwd(v->pre, cls, nullptr, pl, nullptr, n, B200R_ACT_NONE, P, st));
      TAKE(pa, (size_t)n * v->rep);
      RC(b200r_act_planes(pl, pa, (size_t)n * v->rep, B200R_ACT_TANH, st));
      v->pre_logits = pl;
      feat = pa;
    } else {
      RC(lin_fwd(v->pre, cls, nullptr, pl, nullptr, n, B200R_ACT_TANH, P, st));
      feat = pl;
    }
  }
  RC(lin_fwd(v->head, feat, nullptr, nullptr, logits, n, B200R_ACT_NONE, P, st));
  if (save) { v->final_x = x; v->n = n; v->h = h; v->w = w; v->valid = true; }
  return B200R_OK;
}

int vit_input_grad(b200r_model* m, const float* dlogits, float* dx, cudaStream_t s) {
  VitNet* v = m->vit;
  B200R_CHECK_ARG(v->valid, "input_grad needs the activations of a preceding b200r_model_forward_f32 on this handle");
  const b200r_stream_t st = reinterpret_cast<b200r_stream_t>(s);
  const int D = v->dim, P = m->passes, n = v->n, T = v->tokens, rows = n * T, hd = D / v->heads;
  const float scale = 1.0f / sqrtf((float)hd);
  TAKE(g0, (size_t)n * m->classes);
  RC(b200r_split_f32_scaled(dlogits, g0, (size_t)n * m->classes, kGradScale, st));
  TAKE(g1, (size_t)n * v->rep);
  RC(lin_dgrad(v->head, g0, g1, n, P, st));
  uint16_t* gc = g1;
  if (v->has_pre) {
    TAKE(g2, (size_t)n * v->rep);
    RC(b200r_act_bwd_planes(g1, v->pre_logits, g2, (size_t)n * v->rep, B200R_ACT_TANH, st));
    TAKE(g3, (size_t)n * D);
    RC(lin_dgrad(v->pre, g2, g3, n, P, st));
    gc = g3;
  }
  // encoder_norm only feeds x[:, 0]: LayerNorm is row-wise, so its backward runs on the class rows alone
  TAKE(xcls, (size_t)n * D);
  RC(token_rows(v->final_x, xcls, n, T, D, 0, 1, false, s));
  TAKE(gcls, (size_t)n * D);
  RC(b200r_layernorm_bwd(gc, xcls, v->normw, nullptr, gcls, n, D, 1e-5f, st));
  TAKE(g, (size_t)rows * D);
  B200R_CUDA(cudaMemsetAsync(g, 0, (size_t)rows * D * 4, s));
  RC(token_rows(gcls, g, n, T, D, 0, 1, true, s));
  size_t ws_bytes = 0;
  RC(b200r_attention_bwd_workspace_bytes(n, T, v->heads, &ws_bytes));
  void* ws = m->arena.take(ws_bytes + 256);
  B200R_CHECK_ARG(ws, "activation arena too small");
  const size_t mark = m->arena.off;
  uint16_t* gbuf[2] = {g, nullptr};
  gbuf[1] = take_planes(m, (size_t)rows * D);
  B200R_CHECK_ARG(gbuf[1], "activation arena too small");
  const size_t mark2 = m->arena.off;
  (void)mark;
  int cur = 0;
  for (int d = v->depth - 1; d >= 0; --d) {
    const VitBlock& b = v->blocks[d];
    const VitSaved& sv = v->saved[d];
    m->arena.off = mark2;
    TAKE(t1, (size_t)rows * v->hidden);
    RC(lin_dgrad(b.m2, gbuf[cur], t1, rows, P, st));
    TAKE(t2, (size_t)rows * v->hidden);
    RC(b200r_act_bwd_planes(t1, sv.pre, t2, (size_t)rows * v->hidden, B200R_ACT_GELU_TANH, st));
    TAKE(t3, (size_t)rows * D);
    RC(lin_dgrad(b.m1, t2, t3, rows, P, st));
    TAKE(gm, (size_t)rows * D);
    RC(b200r_layernorm_bwd(t3, sv.x_mid, b.n2w, gbuf[cur], gm, rows, D, 1e-5f, st));       // x = x_mid + mlp(norm2(x_mid))
    TAKE(t4, (size_t)rows * D);
    RC(lin_dgrad(b.out, gm, t4, rows, P, st));
    TAKE(t5, (size_t)rows * 3 * D);
    RC(b200r_attention_bwd_ws(sv.qkv, t4, t5, ws, ws_bytes, n, T, v->heads, hd, scale, st));
    TAKE(t6, (size_t)rows * D);
    RC(lin_dgrad(b.qkv, t5, t6, rows, P, st));
    RC(b200r_layernorm_bwd(t6, sv.x_in, b.n1w, gm, gbuf[cur ^ 1], rows, D, 1e-5f, st));    // x_mid = x_in + attn(norm1(x_in))
    cur ^= 1;
  }
  m->arena.off = mark2;
  const int np = T - 1;
  TAKE(gp, (size_t)n * np * D);
  RC(token_rows(gbuf[cur], gp, n, T, D, 1, np, false, s));                                 // drop the class token
  TAKE(dcols, (size_t)n * np * v->embed.k);
  RC(lin_dgrad(v->embed, gp, dcols, n * np, P, st));
  float stdu[3];
  for (int i = 0; i < 3; ++i) stdu[i] = kStd[i] * kGradScale;                              // 1/S rides on the 1/std of the scatter
  return b200r_patch_scatter_f32(dcols, dx, n, v->h, v->w, v->patch, stdu, st);
}


// ---- MLP-Mixer-B/16 (vit/mlp_mixer.py:7-159 on vit/vit_base.py): the launch sequence of nets.Mixer ------------------------------
struct MixBlock { float *n1w, *n1b, *n2w, *n2b; Lin t1, t2, c1, c2; };
struct MixSaved { uint16_t *x_in, *p1, *x_mid, *p2; };
}  // namespace
struct MixerNet {
  int depth = 12, dim = 768, patch = 16, tokens = 196, tok_hidden = 384, ch_hidden = 3072, t_pad = 256;
  Lin embed, head;
  float *normw = nullptr, *normb = nullptr;
  std::vector<MixBlock> blocks;
  std::vector<MixSaved> saved;
  uint16_t* final_x = nullptr;
  int n = 0, h = 0, w = 0;
  bool valid = false;
};
namespace {

int mixer_create(b200r_model* m, const Weights& W) {
  MixerNet* v = new MixerNet();
  m->mixer = v;
  const int D = v->dim, P = v->patch;
  int rc;
  auto f1 = W.find("blocks.0.token_mix.fc1.weight");
  auto c1 = W.find("blocks.0.channel_mix.fc1.weight");
  B200R_CHECK_ARG(f1 != W.end() && c1 != W.end(), "state_dict tensors 'blocks.0.token_mix.fc1.weight' / 'blocks.0.channel_mix.fc1.weight' are missing");
  v->ch_hidden = (int)(c1->second.second / D);
  auto f1b = W.find("blocks.0.token_mix.fc1.bias");
  B200R_CHECK_ARG(f1b != W.end(), "state_dict tensor 'blocks.0.token_mix.fc1.bias' is missing");
  v->tok_hidden = (int)f1b->second.second;
  v->tokens = (int)(f1->second.second / v->tok_hidden);
  B200R_CHECK_ARG(v->tokens <= v->t_pad, "Mixer: %d tokens exceed the padded token dimension %d", v->tokens, v->t_pad);
  if ((rc = make_lin(m, W, "patch_embed.proj", D, 3 * P * P, &v->embed))) return rc;
  v->blocks.resize(v->depth);
  for (int d = 0; d < v->depth; ++d) {
    const std::string p = "blocks." + std::to_string(d) + ".";
    MixBlock& b = v->blocks[d];
    if ((rc = upload_vec(m, W, p + "norm1.weight", D, &b.n1w)) || (rc = upload_vec(m, W, p + "norm1.bias", D, &b.n1b)) ||
        (rc = upload_vec(m, W, p + "norm2.weight", D, &b.n2w)) || (rc = upload_vec(m, W, p + "norm2.bias", D, &b.n2b)) ||
        (rc = make_lin(m, W, p + "token_mix.fc1", v->tok_hidden, v->tokens, &b.t1, 0, v->t_pad)) ||
        (rc = make_lin(m, W, p + "token_mix.fc2", v->tokens, v->tok_hidden, &b.t2, v->t_pad, 0)) ||
        (rc = make_lin(m, W, p + "channel_mix.fc1", v->ch_hidden, D, &b.c1)) || (rc = make_lin(m, W, p + "channel_mix.fc2", D, v->ch_hidden, &b.c2)))
      return rc;
  }
  if ((rc = upload_vec(m, W, "norm.weight", D, &v->normw)) || (rc = upload_vec(m, W, "norm.bias", D, &v->normb))) return rc;
  auto hw = W.find("head.weight");
  B200R_CHECK_ARG(hw != W.end(), "state_dict tensor 'head.weight' is missing");
  m->classes = (int)(hw->second.second / D);
  return make_lin(m, W, "head", m->classes, D, &v->head);
}

size_t mixer_arena_need(const b200r_model* m, int n, bool save) {
  const MixerNet* v = m->mixer;
  // planes elements per image and block: token side D (t_pad + tok_hidden) x 2-3, channel side T (ch_hidden x 2 + D x 4)
  const double tokside = (double)v->dim * (3.0 * v->t_pad + 3.0 * v->tok_hidden), chside = (double)v->tokens * (3.0 * v->ch_hidden + 8.0 * v->dim);
  const double per_img = save ? v->depth * (tokside + chside) + 2.0 * (tokside + chside) : 2.0 * (tokside + chside);
  return (size_t)(per_img * n * 4.0) + (size_t)n * 3 * 224 * 224 * 8 + (64u << 20);
}

int mixer_forward(b200r_model* m, const void* images, bool u8, float* logits, int n, int h, int w, bool save, cudaStream_t s) {
  MixerNet* v = m->mixer;
  const b200r_stream_t st = reinterpret_cast<b200r_stream_t>(s);
  const int D = v->dim, P = m->passes, pp = v->patch, T = (h / pp) * (w / pp), TP = v->t_pad;
  B200R_CHECK_ARG(T == v->tokens, "Mixer: %d x %d images give %d tokens, the token-mixing MLP has %d", h, w, T, v->tokens);
  B200R_CHECK_ARG(m->planes == 2, "Mixer handles run in split precision (passes = 3)");
  m->arena.reset();
  const int rows = n * T, crow = n * D;
  TAKE(cols, (size_t)rows * 3 * pp * pp);
  RC(u8 ? b200r_patch_gather_u8(static_cast<const uint8_t*>(images), cols, n, h, w, pp, kMean, kStd, st)
        : b200r_patch_gather_f32(static_cast<const float*>(images), cols, n, h, w, pp, kMean, kStd, st));
  TAKE(x0, (size_t)rows * D);
  RC(lin_fwd(v->embed, cols, nullptr, x0, nullptr, rows, B200R_ACT_NONE, P, st));
  uint16_t* x = x0;
  if (save) { v->saved.clear(); v->valid = false; }
  uint16_t* xalt = nullptr;
  size_t mark = 0;
  if (!save) {
    xalt = take_planes(m, (size_t)rows * D);
    B200R_CHECK_ARG(xalt, "activation arena too small");
    mark = m->arena.off;
  }
  for (const MixBlock& b : v->blocks) {
    if (!save) m->arena.off = mark;
    TAKE(y1, (size_t)rows * D);
    RC(b200r_layernorm(x, y1, b.n1w, b.n1b, rows, D, 1e-6f, st));                        // vit_base.py:159
    TAKE(yt, (size_t)crow * TP);
    RC(b200r_tokens_to_channels(y1, yt, n, T, D, TP, st));                                // [n*D, 256] (zero padded)
    TAKE(p1, (size_t)crow * v->tok_hidden);
    TAKE(y2, (size_t)crow * TP);
    if (save) {
      TAKE(a1, (size_t)crow * v->tok_hidden);
      RC(lin_keep_pre(b.t1, yt, a1, p1, crow, B200R_ACT_GELU_ERF, P, st));
      RC(lin_fwd(b.t2, a1, nullptr, y2, nullptr, crow, B200R_ACT_NONE, P, st));
    } else {
      RC(lin_fwd(b.t1, yt, nullptr, p1, nullptr, crow, B200R_ACT_GELU_ERF, P, st));       // nn.GELU (erf)
      RC(lin_fwd(b.t2, p1, nullptr, y2, nullptr, crow, B200R_ACT_NONE, P, st));
    }
    TAKE(x_mid, (size_t)rows * D);
    RC(b200r_channels_to_tokens_add(y2, x, x_mid, n, T, D, TP, st));                      // x + token_mix(...)^T
    TAKE(y3, (size_t)rows * D);
    RC(b200r_layernorm(x_mid, y3, b.n2w, b.n2b, rows, D, 1e-6f, st));
    TAKE(p2, (size_t)rows * v->ch_hidden);
    uint16_t* xn;
    if (save) {
      TAKE(a2, (size_t)rows * v->ch_hidden);
      RC(lin_keep_pre(b.c1, y3, a2, p2, rows, B200R_ACT_GELU_ERF, P, st));
      TAKE(xo, (size_t)rows * D);
      RC(lin_fwd(b.c2, a2, x_mid, xo, nullptr, rows, B200R_ACT_NONE, P, st));
      v->saved.push_back({x, p1, x_mid, p2});
      xn = xo;
    } else {
      RC(lin_fwd(b.c1, y3, nullptr, p2, nullptr, rows, B200R_ACT_GELU_ERF, P, st));
      xn = (x == x0) ? xalt : x0;
      RC(lin_fwd(b.c2, p2, x_mid, xn, nullptr, rows, B200R_ACT_NONE, P, st));
    }
    x = xn;
  }
  if (!save) m->arena.off = mark;
  TAKE(xnrm, (size_t)rows * D);
  RC(b200r_layernorm(x, xnrm, v->normw, v->normb, rows, D, 1e-6f, st));
  TAKE(pooled, (size_t)n * D);
  RC(b200r_global_avgpool_nhwc(xnrm, pooled, n, T, D, st));                               // x.mean(dim=1)
  RC(lin_fwd(v->head, pooled, nullptr, nullptr, logits, n, B200R_ACT_NONE, P, st));
  if (save) { v->final_x = x; v->n = n; v->h = h; v->w = w; v->valid = true; }
  return B200R_OK;
}

int mixer_input_grad(b200r_model* m, const float* dlogits, float* dx, cudaStream_t s) {
  MixerNet* v = m->mixer;
  B200R_CHECK_ARG(v->valid, "input_grad needs the activations of a preceding b200r_model_forward_f32 on this handle");
  const b200r_stream_t st = reinterpret_cast<b200r_stream_t>(s);
  const int D = v->dim, P = m->passes, n = v->n, T = v->tokens, rows = n * T, crow = n * D, TP = v->t_pad;
  TAKE(g0, (size_t)n * m->classes);
  RC(b200r_split_f32_scaled(dlogits, g0, (size_t)n * m->classes, kGradScale, st));
  TAKE(g1, (size_t)n * D);
  RC(lin_dgrad(v->head, g0, g1, n, P, st));
  TAKE(g2, (size_t)rows * D);
  RC(b200r_global_avgpool_bwd_nhwc(g1, g2, n, T, D, st));                                 // x.mean(dim=1) backward
  uint16_t* gbuf[2];
  gbuf[0] = take_planes(m, (size_t)rows * D);
  gbuf[1] = take_planes(m, (size_t)rows * D);
  uint16_t* zeros = take_planes(m, (size_t)rows * D);                                     // residual operand of the plain transpose
  B200R_CHECK_ARG(gbuf[0] && gbuf[1] && zeros, "activation arena too small");
  B200R_CUDA(cudaMemsetAsync(zeros, 0, (size_t)rows * D * 4, s));
  RC(b200r_layernorm_bwd(g2, v->final_x, v->normw, nullptr, gbuf[0], rows, D, 1e-6f, st));
  const size_t mark = m->arena.off;
  int cur = 0;
  for (int d = v->depth - 1; d >= 0; --d) {
    const MixBlock& b = v->blocks[d];
    const MixSaved& sv = v->saved[d];
    m->arena.off = mark;
    TAKE(t1, (size_t)rows * v->ch_hidden);
    RC(lin_dgrad(b.c2, gbuf[cur], t1, rows, P, st));
    TAKE(t2, (size_t)rows * v->ch_hidden);
    RC(b200r_act_bwd_planes(t1, sv.p2, t2, (size_t)rows * v->ch_hidden, B200R_ACT_GELU_ERF, st));
    TAKE(t3, (size_t)rows * D);
    RC(lin_dgrad(b.c1, t2, t3, rows, P, st));
    TAKE(gm, (size_t)rows * D);
    RC(b200r_layernorm_bwd(t3, sv.x_mid, b.n2w, gbuf[cur], gm, rows, D, 1e-6f, st));       // x = x_mid + channel_mix(norm2(x_mid))
    TAKE(t4, (size_t)crow * TP);
    RC(b200r_tokens_to_channels(gm, t4, n, T, D, TP, st));
    TAKE(t5, (size_t)crow * v->tok_hidden);
    RC(lin_dgrad(b.t2, t4, t5, crow, P, st));
    TAKE(t6, (size_t)crow * v->tok_hidden);
    RC(b200r_act_bwd_planes(t5, sv.p1, t6, (size_t)crow * v->tok_hidden, B200R_ACT_GELU_ERF, st));
    TAKE(t7, (size_t)crow * TP);
    RC(lin_dgrad(b.t1, t6, t7, crow, P, st));
    TAKE(t8, (size_t)rows * D);
    RC(b200r_channels_to_tokens_add(t7, zeros, t8, n, T, D, TP, st));
    RC(b200r_layernorm_bwd(t8, sv.x_in, b.n1w, gm, gbuf[cur ^ 1], rows, D, 1e-6f, st));    // x_mid = x_in + token_mix(norm1(x_in))
    cur ^= 1;
  }
  m->arena.off = mark;
  TAKE(dcols, (size_t)rows * v->embed.k);
  RC(lin_dgrad(v->embed, gbuf[cur], dcols, rows, P, st));
  float stdu[3];
  for (int i = 0; i < 3; ++i) stdu[i] = kStd[i] * kGradScale;
  return b200r_patch_scatter_f32(dcols, dx, n, v->h, v->w, v->patch, stdu, st);
}

// ---- MobileNetV2 x1.0 (mobilenet_v2.py:80-202) and EfficientNet-B0 (efficientnet.py:91-125,289-495): the inference launch
// sequence of nets.MobileNetV2 / nets.EfficientNetB0 (the ImageNet-C sweep of BASELINE configs[3] is evaluation only) --------------
struct PwConv { Planes w; float *w_f32 = nullptr, *bias = nullptr; int cin = 0, cout = 0; };    // 1x1 conv + folded BN
struct DwConv { float *w = nullptr, *scale = nullptr, *bias = nullptr; int c = 0, k = 3, stride = 1; };
struct MbBlock { bool has_pw = false, has_se = false, res = false; PwConv pw, pl; DwConv dw; Lin se1, se2; };
}  // namespace
struct MobileNet {
  int act = B200R_ACT_RELU6;          // ReLU6 (MobileNetV2) / swish (EfficientNet)
  float *stem_w = nullptr, *stem_scale = nullptr, *stem_bias = nullptr;
  int stem_cout = 32;
  std::vector<MbBlock> blocks;
  PwConv last;
  Lin fc;
};
namespace {

// folded BN as nets._fold_bn returns it: float32 scale and bias (the scale is rounded to float BEFORE it is folded into a weight)
int fold_bn_f32(const Weights& W, const std::string& bn, int c, std::vector<float>* scale, std::vector<float>* bias) {
  std::vector<double> sd;
  int rc = fold_bn(W, bn, c, &sd, bias);
  if (rc) return rc;
  scale->resize(c);
  for (int i = 0; i < c; ++i) (*scale)[i] = (float)sd[i];
  return B200R_OK;
}
int make_pw(b200r_model* m, const Weights& W, const std::string& conv, const std::string& bn, int cin, int cout, PwConv* out) {
  const float* w;
  int rc = get(W, conv + ".weight", (int64_t)cout * cin, &w);
  if (rc) return rc;
  std::vector<float> sc, bi;
  if ((rc = fold_bn_f32(W, bn, cout, &sc, &bi))) return rc;
  std::vector<float> wf((size_t)cout * cin);
  for (int o = 0; o < cout; ++o)
    for (int i = 0; i < cin; ++i) wf[(size_t)o * cin + i] = (float)((double)w[(size_t)o * cin + i] * (double)sc[o]);
  out->cin = cin; out->cout = cout;
  if ((rc = upload_planes(m, wf, &out->w)) || (rc = upload_f32(m, bi, &out->bias))) return rc;
  if (cin % 8 == 0 && cin <= 32) return upload_f32(m, wf, &out->w_f32);       // narrow inputs: the CUDA-core kernel's float32 weights
  return B200R_OK;
}
int make_dw(b200r_model* m, const Weights& W, const std::string& conv, const std::string& bn, int c, int k, int stride, DwConv* out) {
  const float* w;
  int rc = get(W, conv + ".weight", (int64_t)c * k * k, &w);
  if (rc) return rc;
  std::vector<float> wt((size_t)k * k * c), sc, bi;
  for (int ch = 0; ch < c; ++ch)
    for (int t = 0; t < k * k; ++t) wt[(size_t)t * c + ch] = w[(size_t)ch * k * k + t];          // [c, 1, k, k] -> [k*k, c]
  if ((rc = fold_bn_f32(W, bn, c, &sc, &bi))) return rc;
  out->c = c; out->k = k; out->stride = stride;
  if ((rc = upload_f32(m, wt, &out->w)) || (rc = upload_f32(m, sc, &out->scale)) || (rc = upload_f32(m, bi, &out->bias))) return rc;
  return B200R_OK;
}
int make_image_stem(b200r_model* m, const Weights& W, const std::string& conv, const std::string& bn, int cout, MobileNet* v) {
  const float* w;
  int rc = get(W, conv + ".weight", (int64_t)cout * 27, &w);
  if (rc) return rc;
  std::vector<float> w27((size_t)cout * 27), sc, bi;
  for (int o = 0; o < cout; ++o)
    for (int c = 0; c < 3; ++c)
      for (int t = 0; t < 9; ++t) w27[(size_t)o * 27 + t * 3 + c] = w[((size_t)o * 3 + c) * 9 + t];  // [o, c, ky, kx] -> [o, (ky, kx, c)]
  if ((rc = fold_bn_f32(W, bn, cout, &sc, &bi))) return rc;
  v->stem_cout = cout;
  if ((rc = upload_f32(m, w27, &v->stem_w)) || (rc = upload_f32(m, sc, &v->stem_scale)) || (rc = upload_f32(m, bi, &v->stem_bias))) return rc;
  return B200R_OK;
}

int mobilenet_v2_create(b200r_model* m, const Weights& W) {
  MobileNet* v = new MobileNet();
  m->mobile = v;
  v->act = B200R_ACT_RELU6;
  int rc;
  if ((rc = make_image_stem(m, W, "features.0.0", "features.0.1", 32, v))) return rc;
  static const int kSetting[7][4] = {{1, 16, 1, 1}, {6, 24, 2, 2}, {6, 32, 3, 2}, {6, 64, 4, 2}, {6, 96, 3, 1}, {6, 160, 3, 2}, {6, 320, 1, 1}};
  int cin = 32, idx = 1;
  for (const auto& st : kSetting)
    for (int i = 0; i < st[2]; ++i) {
      const int t = st[0], c = st[1], stride = i == 0 ? st[3] : 1, hid = cin * t;
      const std::string p = "features." + std::to_string(idx) + ".conv.";
      MbBlock b;
      b.res = stride == 1 && cin == c;
      int j = 0;
      if (t != 1) {
        b.has_pw = true;
        if ((rc = make_pw(m, W, p + "0.0", p + "0.1", cin, hid, &b.pw))) return rc;
        j = 1;
      }
      if ((rc = make_dw(m, W, p + std::to_string(j) + ".0", p + std::to_string(j) + ".1", hid, 3, stride, &b.dw)) ||
          (rc = make_pw(m, W, p + std::to_string(j + 1), p + std::to_string(j + 2), hid, c, &b.pl)))
        return rc;
      v->blocks.push_back(b);
      cin = c;
      ++idx;
    }
  if ((rc = make_pw(m, W, "features." + std::to_string(idx) + ".0", "features." + std::to_string(idx) + ".1", cin, 1280, &v->last))) return rc;
  auto fw = W.find("classifier.1.weight");
  B200R_CHECK_ARG(fw != W.end(), "state_dict tensor 'classifier.1.weight' is missing");
  m->classes = (int)(fw->second.second / 1280);
  return make_lin(m, W, "classifier.1", m->classes, 1280, &v->fc);
}

int efficientnet_b0_create(b200r_model* m, const Weights& W) {
  MobileNet* v = new MobileNet();
  m->mobile = v;
  v->act = B200R_ACT_SWISH;
  int rc;
  if ((rc = make_image_stem(m, W, "stem.0", "stem.1", 32, v))) return rc;
  static const int kBlocks[7][6] = {{1, 3, 1, 1, 32, 16}, {2, 3, 2, 6, 16, 24}, {2, 5, 2, 6, 24, 40}, {3, 3, 2, 6, 40, 80}, {3, 5, 1, 6, 80, 112},
                                    {4, 5, 2, 6, 112, 192}, {1, 3, 1, 6, 192, 320}};     // (repeat, kernel, stride, expand, in, out), efficientnet.py:101-109
  int bi = 0;
  for (const auto& st : kBlocks)
    for (int r = 0; r < st[0]; ++r) {
      const int k = st[1], e = st[3], ci = r == 0 ? st[4] : st[5], cout = st[5], stride = r == 0 ? st[2] : 1, hid = ci * e;
      const std::string p = "blocks." + std::to_string(bi) + ".";
      MbBlock b;
      b.res = stride == 1 && ci == cout;
      b.has_se = true;
      int j = 0;
      if (e != 1) {
        b.has_pw = true;
        if ((rc = make_pw(m, W, p + "in_conv.0", p + "in_conv.1", ci, hid, &b.pw))) return rc;
        j = 3;
      }
      auto sw = W.find(p + "se_block.conv1.bias");
      B200R_CHECK_ARG(sw != W.end(), "state_dict tensor '%sse_block.conv1.bias' is missing", p.c_str());
      const int se = (int)sw->second.second, sep = (se + 7) / 8 * 8;                       // squeeze width padded to a 16-byte row
      if ((rc = make_dw(m, W, p + "in_conv." + std::to_string(j), p + "in_conv." + std::to_string(j + 1), hid, k, stride, &b.dw)) ||
          (rc = make_lin(m, W, p + "se_block.conv1", se, hid, &b.se1, sep, 0)) || (rc = make_lin(m, W, p + "se_block.conv2", hid, se, &b.se2, 0, sep)) ||
          (rc = make_pw(m, W, p + "out_conv.0", p + "out_conv.1", hid, cout, &b.pl)))
        return rc;
      v->blocks.push_back(b);
      ++bi;
    }
  if ((rc = make_pw(m, W, "head.0", "head.1", 320, 1280, &v->last))) return rc;
  auto fw = W.find("fc.weight");
  B200R_CHECK_ARG(fw != W.end(), "state_dict tensor 'fc.weight' is missing");
  m->classes = (int)(fw->second.second / 1280);
  return make_lin(m, W, "fc", m->classes, 1280, &v->fc);
}

size_t mobile_arena_need(const b200r_model*, int n, int h, int w) {
  // largest block (EfficientNet blocks.1: 96 ch @112^2 + 2 x 96 ch @56^2 + 24 ch @56^2): 37.5 plane elements per input pixel per half
  return (size_t)(2 * 48.0 * n * h * w * 4.0) + (64u << 20);
}

int pw_fwd(b200r_model* m, const PwConv& c, const uint16_t* x, const uint16_t* res, uint16_t* y, int n, int h, int w, int act, b200r_stream_t st) {
  if (c.w_f32) return b200r_pointwise_smallk_nhwc(x, c.w_f32, c.bias, res, y, (size_t)n * h * w, c.cin, c.cout, act, st);
  return b200r_conv2d_nhwc(x, c.w.p, nullptr, c.bias, res, y, nullptr, n, h, w, c.cin, c.cout, 1, 1, 1, 0, act, m->passes, st);
}

int mobile_forward(b200r_model* m, const uint8_t* images, float* logits, int n, int h, int w, cudaStream_t s) {
  MobileNet* v = m->mobile;
  const b200r_stream_t st = reinterpret_cast<b200r_stream_t>(s);
  B200R_CHECK_ARG(m->planes == 2, "the mobile-family handles run in split precision (passes = 3)");
  const int P = m->passes;
  m->arena.region(1);
  int hh = (h - 1) / 2 + 1, ww = (w - 1) / 2 + 1, c = v->stem_cout, bi = 0;
  TAKE(x0, (size_t)n * hh * ww * c);
  RC(b200r_image_stem3x3s2_u8(images, v->stem_w, v->stem_scale, v->stem_bias, x0, n, h, w, c, v->act, kMean, kStd, st));
  uint16_t* x = x0;
  for (const MbBlock& b : v->blocks) {
    m->arena.region(bi++ & 1);                      // the input lives in the other half
    const uint16_t* y = x;
    int hid = c;
    if (b.has_pw) {
      TAKE(a1, (size_t)n * hh * ww * b.pw.cout);
      RC(pw_fwd(m, b.pw, x, nullptr, a1, n, hh, ww, v->act, st));
      y = a1;
      hid = b.pw.cout;
    }
    const int k = b.dw.k, ho = (hh + 2 * (k / 2) - k) / b.dw.stride + 1, wo = (ww + 2 * (k / 2) - k) / b.dw.stride + 1;
    TAKE(a2, (size_t)n * ho * wo * hid);
    RC(b200r_dwconv_nhwc(y, b.dw.w, b.dw.scale, b.dw.bias, a2, n, hh, ww, hid, k, b.dw.stride, k / 2, v->act, st));
    const uint16_t* z = a2;
    if (b.has_se) {
      TAKE(sq, (size_t)n * hid);
      RC(b200r_global_avgpool_nhwc(a2, sq, n, ho * wo, hid, st));                          // squeeze
      TAKE(s1, (size_t)n * b.se1.nout);
      RC(lin_fwd(b.se1, sq, nullptr, s1, nullptr, n, B200R_ACT_SWISH, P, st));
      TAKE(s2, (size_t)n * b.se2.nout);
      RC(lin_fwd(b.se2, s1, nullptr, s2, nullptr, n, B200R_ACT_SIGMOID, P, st));
      TAKE(a3, (size_t)n * ho * wo * hid);
      RC(b200r_channel_scale(a2, s2, a3, n, ho * wo, hid, b.se2.nout, st));                // excite
      z = a3;
    }
    TAKE(o, (size_t)n * ho * wo * b.pl.cout);
    RC(pw_fwd(m, b.pl, z, b.res ? x : nullptr, o, n, ho, wo, B200R_ACT_NONE, st));         // linear bottleneck (+ skip)
    x = o; hh = ho; ww = wo; c = b.pl.cout;
  }
  m->arena.region(bi & 1);
  TAKE(last, (size_t)n * hh * ww * v->last.cout);
  RC(pw_fwd(m, v->last, x, nullptr, last, n, hh, ww, v->act, st));
  TAKE(pooled, (size_t)n * v->last.cout);
  RC(b200r_global_avgpool_nhwc(last, pooled, n, hh * ww, v->last.cout, st));
  return lin_fwd(v->fc, pooled, nullptr, nullptr, logits, n, B200R_ACT_NONE, P, st);
}
}  // namespace

extern "C" {

int b200r_model_create(int arch, const b200r_weight* weights, int n_weights, int passes, b200r_model** out) {
  B200R_CHECK_ARG(out && weights && n_weights > 0, "null argument");
  B200R_CHECK_ARG(arch >= B200R_ARCH_RESNET18 && arch <= B200R_ARCH_EFFICIENTNET_B0, "arch %d: the handle API covers the ResNet family (0..3), ViT-B/16 (4), MLP-Mixer-B/16 (5), MobileNetV2 (6) and EfficientNet-B0 (7)", arch);
  B200R_CHECK_ARG(passes == 3 || (passes == B200R_PASSES_F16 && arch < B200R_ARCH_VIT_B16),
                  "passes must be 3 (split planes, fp32-faithful) or, for the ResNets, B200R_PASSES_F16");
  static const int kLayers[4][4] = {{2, 2, 2, 2}, {3, 4, 6, 3}, {3, 4, 6, 3}, {3, 4, 23, 3}};
  const bool bott = arch >= B200R_ARCH_RESNET50;
  Weights W;
  for (int i = 0; i < n_weights; ++i) {
    B200R_CHECK_ARG(weights[i].name && weights[i].data, "weight %d has a null name / data pointer", i);
    std::string k = weights[i].name;
    for (const char* pre : {"module.", "base_model."})                 // benchmark_eval_adv.py:162-168
      if (k.rfind(pre, 0) == 0) k = k.substr(strlen(pre));
    W[k] = {weights[i].data, weights[i].numel};
  }
  b200r_model* m = new b200r_model();
  m->arch = arch; m->passes = passes; m->f16 = passes == B200R_PASSES_F16; m->planes = m->f16 ? 1 : 2;
  m->feat = bott ? 2048 : 512;
  int rc = B200R_OK;
  auto fail = [&](int code) { b200r_model_destroy(m); return code; };
  if (arch >= B200R_ARCH_VIT_B16) {
    if ((rc = arch == B200R_ARCH_VIT_B16 ? vit_create(m, W) : arch == B200R_ARCH_MIXER_B16 ? mixer_create(m, W)
              : arch == B200R_ARCH_MOBILENET_V2 ? mobilenet_v2_create(m, W) : efficientnet_b0_create(m, W)))
      return fail(rc);
    *out = m;
    return B200R_OK;
  }
  // ---- stem ----
  const float* w1;
  if ((rc = get(W, "conv1.weight", 64 * 3 * 7 * 7, &w1))) return fail(rc);
  std::vector<double> s1;
  std::vector<float> b1;
  if ((rc = fold_bn(W, "bn1", 64, &s1, &b1))) return fail(rc);
  std::vector<float> raw(w1, w1 + 64 * 147), folded(64 * 147), sc(64);
  for (int o = 0; o < 64; ++o) {
    sc[o] = (float)s1[o];
    for (int j = 0; j < 147; ++j) folded[o * 147 + j] = (float)((double)w1[o * 147 + j] * s1[o]);
  }
  const std::vector<float> packed = pack_stem(raw);
  if ((rc = upload_planes(m, packed, &m->stem_w)) || (rc = upload_f32(m, sc, &m->stem_scale)) || (rc = upload_f32(m, b1, &m->stem_bias)))
    return fail(rc);
  if (m->f16 && (rc = upload_planes(m, pack_stem(folded), &m->stem_w_folded))) return fail(rc);
  if (!m->f16) {
    std::vector<uint16_t> wp((size_t)2 * 64 * 224);
    void* d = nullptr;
    if ((rc = b200r_stem_pool_split_prepare(w1, sc.data(), b1.data(), kMean, kStd, 0, wp.data(), &m->stem_osc)) || (rc = dev_alloc(m, &d, wp.size() * 2)))
      return fail(rc);
    if (cudaMemcpy(d, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) { b200r_set_error("cudaMemcpy of the stem operand failed"); return fail(B200R_ECUDA); }
    m->stem_wp = static_cast<uint16_t*>(d);
    void* df = nullptr;
    if ((rc = b200r_stem_pool_split_prepare(w1, sc.data(), b1.data(), kMean, kStd, 1, wp.data(), &m->stem_oscf)) || (rc = dev_alloc(m, &df, wp.size() * 2)))
      return fail(rc);
    if (cudaMemcpy(df, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) { b200r_set_error("cudaMemcpy of the stem operand failed"); return fail(B200R_ECUDA); }
    m->stem_wpf = static_cast<uint16_t*>(df);
  }
  {  // gradient of the stem GEMM: [192][64] = (planes' value of the packed weight * bn scale)^T
    std::vector<float> t((size_t)192 * 64);
    for (int o = 0; o < 64; ++o)
      for (int j = 0; j < 192; ++j) t[(size_t)j * 64 + o] = planes_value(packed[(size_t)o * 192 + j], m->planes) * sc[o];
    // single fp16 plane in both precisions (see nets.ResNet.input_grad): the [n*112*112, 192] output only feeds col2im
    const int planes_saved = m->planes;
    m->planes = 1;
    rc = upload_planes(m, t, &m->stem_wt);
    m->planes = planes_saved;
    if (rc) return fail(rc);
  }
  // ---- residual stages ----
  int inplanes = 64;
  const int widths[4] = {64, 128, 256, 512};
  for (int li = 0; li < 4; ++li)
    for (int bi = 0; bi < kLayers[arch][li]; ++bi) {
      const int planes = widths[li], stride = (bi == 0 && li > 0) ? 2 : 1, exp = bott ? 4 : 1;
      const std::string p = "layer" + std::to_string(li + 1) + "." + std::to_string(bi);
      Block b;
      b.bottleneck = bott;
      if (bott) {
        if ((rc = make_conv(m, W, p + ".conv1", p + ".bn1", inplanes, planes, 1, 1, 0, &b.c1)) ||
            (rc = make_conv(m, W, p + ".conv2", p + ".bn2", planes, planes, 3, stride, 1, &b.c2)) ||       // stride on the 3x3 (:112)
            (rc = make_conv(m, W, p + ".conv3", p + ".bn3", planes, planes * 4, 1, 1, 0, &b.c3)))
          return fail(rc);
      } else {
        if ((rc = make_conv(m, W, p + ".conv1", p + ".bn1", inplanes, planes, 3, stride, 1, &b.c1)) ||
            (rc = make_conv(m, W, p + ".conv2", p + ".bn2", planes, planes, 3, 1, 1, &b.c2)))
          return fail(rc);
      }
      if (stride != 1 || inplanes != planes * exp) {
        b.has_down = true;
        if ((rc = make_conv(m, W, p + ".downsample.0", p + ".downsample.1", inplanes, planes * exp, 1, stride, 0, &b.down))) return fail(rc);
      }
      inplanes = planes * exp;
      m->blocks.push_back(b);
    }
  // ---- head ----
  auto fcw = W.find("fc.weight");
  if (fcw == W.end()) { b200r_set_error("state_dict tensor 'fc.weight' is missing"); return fail(B200R_EINVAL); }
  m->classes = (int)(fcw->second.second / m->feat);
  const float *fw, *fb;
  if ((rc = get(W, "fc.weight", (int64_t)m->classes * m->feat, &fw)) || (rc = get(W, "fc.bias", m->classes, &fb))) return fail(rc);
  std::vector<float> fcv(fw, fw + (size_t)m->classes * m->feat), fct(fcv.size()), fbv(fb, fb + m->classes);
  for (int o = 0; o < m->classes; ++o)
    for (int j = 0; j < m->feat; ++j) fct[(size_t)j * m->classes + o] = planes_value(fcv[(size_t)o * m->feat + j], m->planes);
  if ((rc = upload_planes(m, fcv, &m->fc_w)) || (rc = upload_planes(m, fct, &m->fc_wt)) || (rc = upload_f32(m, fbv, &m->fc_b))) return fail(rc);
  *out = m;
  return B200R_OK;
}

int b200r_model_destroy(b200r_model* m) {
  if (!m) return B200R_OK;
  for (void* p : m->owned) cudaFree(p);
  if (m->arena.base) cudaFree(m->arena.base);
  delete m->vit;
  delete m->mixer;
  delete m->mobile;
  delete m;
  return B200R_OK;
}

int b200r_model_num_classes(const b200r_model* m) { return m ? m->classes : 0; }

int b200r_model_reserve(b200r_model* m, int n, int h, int w, int for_input_grad) {
  B200R_CHECK_ARG(m && n > 0 && h > 0 && w > 0, "bad argument");
  return ensure_arena(m, n, h, w, for_input_grad != 0);
}

int b200r_model_forward_u8(b200r_model* m, const uint8_t* images, float* logits, int n, int h, int w, b200r_stream_t stream) {
  B200R_CHECK_ARG(m && images && logits && n > 0, "bad argument");
  B200R_CHECK_ARG(h % 32 == 0 && w % 32 == 0, "image size must be a multiple of 32 (got %dx%d)", h, w);
  RC(ensure_arena(m, n, h, w, false));
  if (m->vit) return vit_forward(m, images, true, logits, n, h, w, false, as_stream(stream));
  if (m->mixer) return mixer_forward(m, images, true, logits, n, h, w, false, as_stream(stream));
  if (m->mobile) return mobile_forward(m, images, logits, n, h, w, as_stream(stream));
  m->arena.region(1);
  m->saved.valid = false;
  cudaStream_t s = as_stream(stream);
  uint16_t* x;
  if (h % 4 == 0 && w % 8 == 0 && w >= 8 && w <= 248 && h >= 8) {
    // raw pixels -> conv1 + bn1 + relu + maxpool in one launch
    TAKE(p, (size_t)n * (h / 4) * (w / 4) * 64);
    RC(m->f16 ? b200r_stem_pool_u8_f16(images, m->stem_w_folded.p, m->stem_bias, p, n, h, w, kMean, kStd, stream)
              : b200r_stem_pool_u8_split(images, m->stem_wp, m->stem_osc, p, n, h, w, stream));
    x = p;
  } else {
    TAKE(s0, (size_t)n * (h / 2) * (w / 2) * 64);
    RC(b200r_stem_conv7x7_u8(images, m->stem_w.p, m->stem_scale, m->stem_bias, s0, n, h, w, kMean, kStd, B200R_ACT_RELU, m->passes, stream));
    TAKE(p, (size_t)n * (h / 4) * (w / 4) * 64);
    RC((m->f16 ? b200r_maxpool3x3s2_nhwc_f16 : b200r_maxpool3x3s2_nhwc)(s0, p, n, h / 2, w / 2, 64, stream));
    x = p;
  }
  return run_body(m, x, n, h / 4, w / 4, logits, false, s);
}

int b200r_model_forward_f32(b200r_model* m, const float* x01, float* logits, int n, int h, int w, b200r_stream_t stream) {
  B200R_CHECK_ARG(m && x01 && logits && n > 0, "bad argument");
  B200R_CHECK_ARG(h % 32 == 0 && w % 32 == 0 && w <= 256, "float input: image size must be a multiple of 32 and at most 256 wide (got %dx%d)", h, w);
  RC(ensure_arena(m, n, h, w, true));
  if (m->vit) return vit_forward(m, x01, false, logits, n, h, w, true, as_stream(stream));
  if (m->mixer) return mixer_forward(m, x01, false, logits, n, h, w, true, as_stream(stream));
  if (m->mobile) { b200r_set_error("the MobileNetV2 / EfficientNet-B0 handles are inference handles (b200r_model_forward_u8): their input gradient is sequenced by robustart_b200.nets"); return B200R_ENOTSUP; }
  m->arena.reset();
  cudaStream_t s = as_stream(stream);
  if (!m->f16 && h % 4 == 0 && w % 8 == 0 && w >= 8 && w <= 248 && h >= 8) {
    // conv1 + bn1 + relu + maxpool + arg-max codes in one launch: the 112 x 112 activation is never written
    TAKE(pq, (size_t)n * (h / 4) * (w / 4) * 64);
    m->saved.pool_codes = m->arena.take((size_t)n * (h / 4) * (w / 4) * 64 + 8);
    B200R_CHECK_ARG(m->saved.pool_codes, "activation arena too small");
    RC(b200r_stem_pool_f32_split(x01, m->stem_wpf, m->stem_oscf, pq, static_cast<uint8_t*>(m->saved.pool_codes), n, h, w, stream));
    m->saved.stem = nullptr; m->saved.n = n; m->saved.h = h; m->saved.w = w;
    const int rc = run_body(m, pq, n, h / 4, w / 4, logits, true, s);
    m->saved.valid = rc == B200R_OK;
    return rc;
  }
  TAKE(s0, (size_t)n * (h / 2) * (w / 2) * 64);
  RC(b200r_stem_conv7x7_f32(x01, m->stem_w.p, m->stem_scale, m->stem_bias, s0, n, h, w, kMean, kStd, B200R_ACT_RELU, m->passes, stream));
  TAKE(p, (size_t)n * (h / 4) * (w / 4) * 64);
  m->saved.pool_codes = nullptr;
  if (m->f16) {
    RC(b200r_maxpool3x3s2_nhwc_f16(s0, p, n, h / 2, w / 2, 64, stream));
  } else {          // the pool leaves its arg-max codes (stem ReLU's backward folded in): the gradient pass never reads s0 again
    m->saved.pool_codes = m->arena.take((size_t)n * (h / 4) * (w / 4) * 64 + 8);
    B200R_CHECK_ARG(m->saved.pool_codes, "activation arena too small");
    RC(b200r_maxpool3x3s2_nhwc_codes(s0, p, m->saved.pool_codes, n, h / 2, w / 2, 64, stream));
  }
  m->saved.stem = s0; m->saved.n = n; m->saved.h = h; m->saved.w = w;
  int rc = run_body(m, p, n, h / 4, w / 4, logits, true, s);
  m->saved.valid = rc == B200R_OK;
  return rc;
}

// d loss / d x01 from d loss / d logits and the activations of the last b200r_model_forward_f32
int b200r_model_input_grad(b200r_model* m, const float* dlogits, float* dx, b200r_stream_t stream) {
  B200R_CHECK_ARG(m && dlogits && dx, "bad argument");
  if (m->vit) return vit_input_grad(m, dlogits, dx, as_stream(stream));
  if (m->mixer) return mixer_input_grad(m, dlogits, dx, as_stream(stream));
  if (m->mobile) { b200r_set_error("the MobileNetV2 / EfficientNet-B0 handles are inference handles"); return B200R_ENOTSUP; }
  B200R_CHECK_ARG(m->saved.valid, "b200r_model_input_grad needs a preceding b200r_model_forward_f32 on this handle");
  const int n = m->saved.n, H = m->saved.h, Wd = m->saved.w, P = m->passes;
  const bool f16 = m->f16;
  auto relu_bwd = f16 ? b200r_relu_bwd_f16 : b200r_relu_bwd;
  auto dilate2 = f16 ? b200r_dilate2_nhwc_f16 : b200r_dilate2_nhwc;
  // S * dlogits in the precision's planes
  TAKE(g0, (size_t)n * m->classes);
  if (f16) {
    RC(b200r_f32_to_f16(dlogits, g0, (size_t)n * m->classes, kGradScale, stream));
  } else {
    RC(b200r_split_f32_scaled(dlogits, g0, (size_t)n * m->classes, kGradScale, stream));
  }
  int h = H / 32, w = Wd / 32, c = m->feat;
  TAKE(g1, (size_t)n * c);
  RC(b200r_linear(g0, m->fc_wt.p, nullptr, nullptr, nullptr, g1, nullptr, n, m->classes, c, B200R_ACT_NONE, P, stream));
  TAKE(g2, (size_t)n * h * w * c);
  RC((f16 ? b200r_global_avgpool_bwd_nhwc_f16 : b200r_global_avgpool_bwd_nhwc)(g1, g2, n, h * w, c, stream));
  uint16_t* last = m->saved.blocks.back().back();
  TAKE(g, (size_t)n * h * w * c);
  RC(relu_bwd(g2, last, nullptr, g, (size_t)n * h * w * c, stream));        // the last block's output ReLU; all others are fused
  for (int i = (int)m->blocks.size() - 1; i >= 0; --i) {
    const Block& b = m->blocks[i];
    const std::vector<uint16_t*>& sv = m->saved.blocks[i];
    const uint16_t* in_mask = i > 0 ? m->saved.blocks[i - 1].back() : nullptr;     // block input = previous block's output
    const int stride = b.bottleneck ? b.c2.stride : b.c1.stride;
    const int cin = b.c1.cin, hi = h * stride, wi = w * stride;
    // identity path: the 1x1 downsample's dgrad (stride 2: contract on the small map, then zero-insert) or g itself
    // (a stride-2 downsample is accumulated in place into the main branch's result afterwards: no zero-inserted tensor, no residual)
    static const bool ds_acc = !(getenv("B200R_DS_ACC") && atoi(getenv("B200R_DS_ACC")) == 0);
    const bool acc = b.has_down && stride == 2 && ds_acc;
    const uint16_t* r = acc ? nullptr : g;
    if (b.has_down && !acc) {
      TAKE(d0, (size_t)n * h * w * cin);
      RC(b200r_conv2d_dgrad_nhwc(g, b.down.wt.p, nullptr, nullptr, d0, n, h, w, b.down.cout, cin, 1, 1, 0, P, stream));
      if (stride == 2) {
        TAKE(d1, (size_t)n * hi * wi * cin);
        RC(dilate2(d0, d1, n, h, w, cin, stream));
        r = d1;
      } else {
        r = d0;
      }
    }
    uint16_t* t;
    if (b.bottleneck) {
      TAKE(t3, (size_t)n * h * w * b.c3.cin);
      RC(b200r_conv2d_dgrad_nhwc(g, b.c3.wt.p, nullptr, sv[1], t3, n, h, w, b.c3.cout, b.c3.cin, 1, 1, 0, P, stream));
      TAKE(t2, (size_t)n * hi * wi * b.c2.cin);
      if (stride == 2) {        // parity classes: four small convolutions of t3 itself, no zero-dilated tensor
        RC(b200r_conv2d_dgrad3x3s2_nhwc(t3, b.c2.wt_s2[0].p, b.c2.wt_s2[1].p, b.c2.wt_s2[2].p, b.c2.wt_s2[3].p, nullptr, sv[0], t2, n, h, w,
                                        b.c2.cout, b.c2.cin, P, stream));
      } else {
        RC(b200r_conv2d_dgrad_nhwc(t3, b.c2.wt.p, nullptr, sv[0], t2, n, hi, wi, b.c2.cout, b.c2.cin, 3, 3, 1, P, stream));
      }
      TAKE(t1, (size_t)n * hi * wi * cin);
      RC(b200r_conv2d_dgrad_nhwc(t2, b.c1.wt.p, r, in_mask, t1, n, hi, wi, b.c1.cout, cin, 1, 1, 0, P, stream));
      t = t1;
    } else {
      TAKE(t2, (size_t)n * h * w * b.c2.cin);
      RC(b200r_conv2d_dgrad_nhwc(g, b.c2.wt.p, nullptr, sv[0], t2, n, h, w, b.c2.cout, b.c2.cin, 3, 3, 1, P, stream));
      TAKE(t1, (size_t)n * hi * wi * cin);
      if (stride == 2) {
        RC(b200r_conv2d_dgrad3x3s2_nhwc(t2, b.c1.wt_s2[0].p, b.c1.wt_s2[1].p, b.c1.wt_s2[2].p, b.c1.wt_s2[3].p, r, in_mask, t1, n, h, w,
                                        b.c1.cout, cin, P, stream));
      } else {
        RC(b200r_conv2d_dgrad_nhwc(t2, b.c1.wt.p, r, in_mask, t1, n, hi, wi, b.c1.cout, cin, 3, 3, 1, P, stream));
      }
      t = t1;
    }
    if (acc) RC(b200r_conv2d_dgrad1x1s2_acc_nhwc(g, b.down.wt.p, in_mask, t, n, hi, wi, b.down.cout, cin, P, stream));
    g = t; h = hi; w = wi; c = cin;
  }
  // maxpool backward into the 112^2 stem activation, its ReLU, the stem GEMM's gradient, col2im (+ 1/std, 1/S)
  const int h2 = H / 2, w2 = Wd / 2;
  uint16_t* gr = static_cast<uint16_t*>(m->arena.take((size_t)n * h2 * w2 * 64 * 2));               // ONE fp16 plane
  B200R_CHECK_ARG(gr, "activation arena too small");
  if (m->saved.pool_codes) {
    RC(b200r_maxpool3x3s2_bwd_codes_hi(m->saved.pool_codes, g, gr, n, h2, w2, 64, 2, stream));
  } else {
    void* ws = m->arena.take((size_t)n * h * w * 64 + 8);
    B200R_CHECK_ARG(ws, "activation arena too small");
    RC(b200r_maxpool3x3s2_relu_bwd_hi(m->saved.stem, g, gr, ws, (size_t)n * h * w * 64, n, h2, w2, 64, f16 ? 1 : 2, stream));
  }
  uint16_t* dcols = static_cast<uint16_t*>(m->arena.take((size_t)n * h2 * w2 * 192 * 2));            // ONE fp16 plane (hi plane of gr in)
  B200R_CHECK_ARG(dcols, "activation arena too small");
  RC(b200r_linear(gr, m->stem_wt.p, nullptr, nullptr, nullptr, dcols, nullptr, n * h2 * w2, 64, 192, B200R_ACT_NONE, B200R_PASSES_F16, stream));
  return b200r_stem_col2im_f32_f16(dcols, dx, n, H, Wd, kStd, 1.0f / kGradScale, stream);
}

// The evaluation's ONE collective: sum the int64 hit counters over the ranks of an NCCL communicator, in place, on
// `stream`.  NCCL is resolved at run time (the symbol already loaded into the process, else libnccl.so.2), so the library itself
// does not link it.  Replaces the reference's per-rank result files + merge (base_dataset.py:116-133) and its barrier all-reduces
// (linklink/__init__.py:37-41).
int b200r_allreduce_counts(void* nccl_comm, int64_t* dev_counts, int count, b200r_stream_t stream) {
  B200R_CHECK_ARG(nccl_comm && dev_counts && count > 0, "bad argument");
  typedef int (*allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  static allreduce_fn fn = nullptr;
  if (!fn) {
    fn = reinterpret_cast<allreduce_fn>(dlsym(RTLD_DEFAULT, "ncclAllReduce"));
    if (!fn) {
      void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (h) fn = reinterpret_cast<allreduce_fn>(dlsym(h, "ncclAllReduce"));
    }
    if (!fn) { b200r_set_error("ncclAllReduce not found: load NCCL (libnccl.so.2) into the process first"); return B200R_ENOTSUP; }
  }
  const int rc = fn(dev_counts, dev_counts, (size_t)count, /*ncclInt64*/ 4, /*ncclSum*/ 0, nccl_comm, as_stream(stream));
  if (rc != 0) { b200r_set_error("ncclAllReduce failed with ncclResult_t %d", rc); return B200R_ECUDA; }
  return B200R_OK;
}

}  // extern "C"
