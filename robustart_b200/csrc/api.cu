// C-ABI front door of libb200robust.so: argument validation, dispatch, error string.
#include "corrupt.cuh"
#include <mutex>

static thread_local char g_err[1024] = "";

void b200r_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int b200r_num_sms() {
  static int cached[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 8) return 148;
  if (!cached[dev]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

extern "C" {

const char* b200r_last_error(void) { return g_err; }
int b200r_version(void) { return 100; }

int b200r_sm_count(int* sms) {
  B200R_CHECK_ARG(sms, "null output");
  int dev = 0, v = 0;
  B200R_CUDA(cudaGetDevice(&dev));
  B200R_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
  *sms = v;
  return B200R_OK;
}

static int family_of(int id) {
  switch (id) {
    case B200R_GAUSSIAN_NOISE: case B200R_SHOT_NOISE: case B200R_IMPULSE_NOISE: case B200R_FROST:
    case B200R_FOG: case B200R_BRIGHTNESS: case B200R_CONTRAST: case B200R_SPECKLE_NOISE:
    case B200R_SATURATE:
      return 0;
    case B200R_DEFOCUS_BLUR: case B200R_GLASS_BLUR: case B200R_MOTION_BLUR: case B200R_ZOOM_BLUR:
    case B200R_SNOW: case B200R_ELASTIC_TRANSFORM: case B200R_GAUSSIAN_BLUR: case B200R_SPATTER:
      return 1;
    case B200R_PIXELATE: case B200R_JPEG_COMPRESSION:
      return 2;
    default:
      return -1;
  }
}

int b200r_corrupt_workspace_bytes(int id, int severity, int n, int h, int w, size_t* bytes) {
  B200R_CHECK_ARG(bytes, "null output");
  B200R_CHECK_ARG(family_of(id) >= 0, "unknown corruption id %d", id);
  B200R_CHECK_ARG(severity >= 1 && severity <= 5, "severity %d not in 1..5", severity);
  B200R_CHECK_ARG(n >= 0 && h > 0 && w > 0, "bad shape n=%d h=%d w=%d", n, h, w);
  switch (family_of(id)) {
    case 0: *bytes = corrupt_pixel_ws(id, severity, n, h, w); break;
    case 1: *bytes = corrupt_stencil_ws(id, severity, n, h, w); break;
    default: *bytes = corrupt_codec_ws(id, severity, n, h, w); break;
  }
  return B200R_OK;
}

int b200r_corrupt_ext_noise_count(int id, int severity, int n, int h, int w, size_t* count) {
  B200R_CHECK_ARG(count, "null output");
  B200R_CHECK_ARG(family_of(id) >= 0, "unknown corruption id %d", id);
  B200R_CHECK_ARG(severity >= 1 && severity <= 5, "severity %d not in 1..5", severity);
  *count = corrupt_ext_count(id, severity, n, h, w);
  return B200R_OK;
}

int b200r_corrupt_u8(int id, int severity, const uint8_t* in, uint8_t* out, int n, int h, int w,
                     uint64_t seed, uint64_t image_offset, const float* ext_noise, void* workspace,
                     size_t workspace_bytes, b200r_stream_t stream) {
  B200R_CHECK_ARG(family_of(id) >= 0, "unknown corruption id %d", id);
  // the reference indexes c[severity - 1], so 0 silently means 5 (imagenet_c docs say [0,5]);
  // we reject it instead of mirroring the quirk.
  B200R_CHECK_ARG(severity >= 1 && severity <= 5, "severity %d not in 1..5", severity);
  B200R_CHECK_ARG(n >= 0 && h > 0 && w > 0, "bad shape n=%d h=%d w=%d", n, h, w);
  if (n == 0) return B200R_OK;
  B200R_CHECK_ARG(in && out, "null image pointer");
  B200R_CHECK_ARG((((uintptr_t)in | (uintptr_t)out) & 15) == 0, "image pointers must be 16-byte aligned");
  CorruptArgs a{id, severity, in, out, n, h, w, seed, image_offset, ext_noise, workspace, workspace_bytes,
                as_stream(stream)};
  switch (family_of(id)) {
    case 0: return corrupt_pixel_family(a);
    case 1: return corrupt_stencil_family(a);
    default: return corrupt_codec_family(a);
  }
}

}  // extern "C"
