/* Plain-C caller of the model-handle API (include/b200r.h): builds a ResNet from a state_dict dumped by the Python test,
 * runs a forward from uint8 pixels, a forward from float pixels + an input-gradient pass, and compares with the values the
 * Python layer sequencing (robustart_b200/nets.py) produced with the same kernels.  No Python, no torch: libb200robust + cudart.
 *
 *   gcc tests/c/test_model_handle.c -I include -L robustart_b200/lib -lb200robust -L/usr/local/cuda/lib64 -lcudart -lm -o t
 *   ./t dump.bin
 * dump.bin: int32 arch, passes, n, h, w, classes, n_weights; per weight: int32 name_len, name, int64 numel, float32 data;
 *           uint8 images[n*h*w*3]; float32 x01[n*3*h*w] (the same pixels / 255, NCHW); float32 logits_u8[n*classes];
 *           float32 dlogits[n*classes]; float32 logits_f32[n*classes]; float32 dx[n*3*h*w]  */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "b200r.h"

#define CHECK(call) do { int rc_ = (call); if (rc_) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, b200r_last_error()); return 1; } } while (0)
#define CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s -> %s\n", #call, cudaGetErrorString(e_)); return 1; } } while (0)

static void* rd(FILE* f, size_t bytes) {
  void* p = malloc(bytes ? bytes : 1);
  if (fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read\n"); exit(2); }
  return p;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s dump.bin\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  int32_t* hd = (int32_t*)rd(f, 7 * 4);
  const int arch = hd[0], passes = hd[1], n = hd[2], h = hd[3], w = hd[4], classes = hd[5], nw = hd[6];
  b200r_weight* W = (b200r_weight*)calloc(nw, sizeof(b200r_weight));
  for (int i = 0; i < nw; ++i) {
    int32_t* len = (int32_t*)rd(f, 4);
    char* name = (char*)rd(f, *len);
    char* z = (char*)malloc(*len + 1);
    memcpy(z, name, *len); z[*len] = 0;
    int64_t* numel = (int64_t*)rd(f, 8);
    W[i].name = z; W[i].numel = *numel; W[i].data = (const float*)rd(f, (size_t)*numel * 4);
  }
  const size_t npx = (size_t)n * h * w * 3, nl = (size_t)n * classes;
  uint8_t* images = (uint8_t*)rd(f, npx);
  float* x = (float*)rd(f, npx * 4);
  float* want_u8 = (float*)rd(f, nl * 4);
  float* dlogits = (float*)rd(f, nl * 4);
  float* want_f32 = (float*)rd(f, nl * 4);
  float* want_dx = (float*)rd(f, npx * 4);
  fclose(f);

  b200r_model* model = NULL;
  CHECK(b200r_model_create(arch, W, nw, passes, &model));
  if (b200r_model_num_classes(model) != classes) { fprintf(stderr, "classes %d != %d\n", b200r_model_num_classes(model), classes); return 1; }
  uint8_t* d_img; float *d_logits, *d_x, *d_dl, *d_dx;
  CUDA(cudaMalloc((void**)&d_img, npx)); CUDA(cudaMalloc((void**)&d_logits, nl * 4)); CUDA(cudaMalloc((void**)&d_x, npx * 4));
  CUDA(cudaMalloc((void**)&d_dl, nl * 4)); CUDA(cudaMalloc((void**)&d_dx, npx * 4));
  CUDA(cudaMemcpy(d_img, images, npx, cudaMemcpyHostToDevice));
  CHECK(b200r_model_reserve(model, n, h, w, 1));
  /* 1. inference from raw pixels */
  CHECK(b200r_model_forward_u8(model, d_img, d_logits, n, h, w, NULL));
  float* got = (float*)malloc(nl * 4);
  CUDA(cudaMemcpy(got, d_logits, nl * 4, cudaMemcpyDeviceToHost));
  double e1 = 0;
  for (size_t i = 0; i < nl; ++i) e1 = fmax(e1, fabs((double)got[i] - want_u8[i]));
  /* 2. float pixels in [0,1] (NCHW) + input gradient */
  CUDA(cudaMemcpy(d_x, x, npx * 4, cudaMemcpyHostToDevice));
  CUDA(cudaMemcpy(d_dl, dlogits, nl * 4, cudaMemcpyHostToDevice));
  CHECK(b200r_model_forward_f32(model, d_x, d_logits, n, h, w, NULL));
  CHECK(b200r_model_input_grad(model, d_dl, d_dx, NULL));
  CUDA(cudaDeviceSynchronize());
  CUDA(cudaMemcpy(got, d_logits, nl * 4, cudaMemcpyDeviceToHost));
  double e2 = 0, e3 = 0, gmax = 0;
  for (size_t i = 0; i < nl; ++i) e2 = fmax(e2, fabs((double)got[i] - want_f32[i]));
  float* gdx = (float*)malloc(npx * 4);
  CUDA(cudaMemcpy(gdx, d_dx, npx * 4, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < npx; ++i) { e3 = fmax(e3, fabs((double)gdx[i] - want_dx[i])); gmax = fmax(gmax, fabs((double)want_dx[i])); }
  /* error path: a missing tensor must be refused, with a message */
  b200r_model* bad = NULL;
  const int rc = b200r_model_create(arch, W, nw - 1, passes, &bad);
  CHECK(b200r_model_destroy(model));
  printf("max|dlogit| u8 %.3g  f32 %.3g  max|d dx| %.3g (|dx| max %.3g)  missing-tensor rc %d (%s)\n", e1, e2, e3, gmax, rc, b200r_last_error());
  if (e1 != 0 || e2 != 0 || e3 != 0 || rc == 0) { printf("FAIL\n"); return 1; }
  printf("PASS\n");
  return 0;
}
