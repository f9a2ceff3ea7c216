// Test-only shim: exposes the product's host-side libswscale filter construction (csrc/swscale_filter.h) to Python, so that the
// tables the GPU kernels consume can be checked on a machine without a GPU (tests/test_sws_filter.py).
#include "../libnativecpurenderer_b200/csrc/swscale_filter.h"

extern "C" int ncr_probe_sws_filter(int src, int dst, int align, long one, int* pos_out, int* coef_out, int coef_capacity) {
    const NcrSwsFilter f = ncr_sws_make_filter(src, dst, align, one);
    if ((int)f.coef.size() > coef_capacity) return -1;
    for (int i = 0; i < dst; ++i) pos_out[i] = f.pos[i];
    for (size_t k = 0; k < f.coef.size(); ++k) coef_out[k] = f.coef[k];
    return f.size;
}
