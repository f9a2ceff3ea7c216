// Launch interface between the host runtime (api.cu) and the kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "ncr_cmd.h"

// Alpha read from a 3-channel texture.  The reference leaves it uninitialised (cpp:746-748 with
// cpp:571-573: `a` is only written when the texture has alpha), so its output there is garbage that
// changes from call to call; the product defines it as opaque.  See DESIGN.md "Undefined in the reference".
#define NCR_RGB_TEXTURE_ALPHA 1.0

// Filter tables + geometry of one w x h -> dw x dh libswscale-exact conversion (host builds them: swscale_filter.h).
struct NcrSwsPlan {
    int w, h, dw, dh, cw, cdw, cdh, half;
    const int32_t *hl_pos, *hl_coef, *vl_pos, *vl_coef, *hc_pos, *hc_coef, *vc_pos, *vc_coef;
    int hl_size, vl_size, hc_size, vc_size;
};

extern "C" {
// memset cursors, bin (coarse, fine), composite — all on stream s.  ev: 4 events recorded around the
// three kernels (timing mode) or nullptr.
void ncr_launch_flush(const NcrFlushArgs* A, cudaStream_t s, cudaEvent_t* ev);
void ncr_launch_composite(const NcrFlushArgs* A, cudaStream_t s);
void ncr_launch_convert_u8(const double* fb, unsigned char* out, size_t n, cudaStream_t s);
void ncr_launch_yuv420p(const unsigned char* img, unsigned char* out, int w, int h, int ipp, cudaStream_t s);
// Scaling branch of the present path: `plan` is an NcrSwsPlan (kernels.cu / api.cu) whose table pointers are device memory.
void ncr_launch_sws_scaled(const unsigned char* img, int ipp, const void* plan, short* mid_y, short* mid_u, short* mid_v,
                           unsigned char* out, cudaStream_t s);
void ncr_launch_resample(const NcrCmd* src, void* out, int ow, int oh, cudaStream_t s);
double ncr_measure_f64_rate(cudaStream_t s);   // non-fused DMUL+DADD instructions per second
}
