#include "corrupt.cuh"
size_t corrupt_codec_ws(int, int, int, int, int) { return 0; }
int corrupt_codec_family(const CorruptArgs& a) { b200r_set_error("corruption %d not implemented yet", a.id); return B200R_ENOTSUP; }
