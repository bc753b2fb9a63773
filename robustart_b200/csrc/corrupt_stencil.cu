// placeholder, replaced below
#include "corrupt.cuh"
size_t corrupt_stencil_ws(int, int, int, int, int) { return 0; }
int corrupt_stencil_family(const CorruptArgs& a) { b200r_set_error("corruption %d not implemented yet", a.id); return B200R_ENOTSUP; }
size_t corrupt_ext_count(int id, int sev, int n, int h, int w) {
  const size_t P = (size_t)h * w * 3;
  switch (id) {
    case B200R_GAUSSIAN_NOISE: case B200R_SPECKLE_NOISE: case B200R_SHOT_NOISE: return n * P;
    case B200R_IMPULSE_NOISE: return 2 * n * P;
    case B200R_FROST: return 3 * (size_t)n;
    case B200R_FOG: return 65535 * (size_t)n;
    default: return 0;
  }
}
