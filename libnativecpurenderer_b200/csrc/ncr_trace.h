// Binary command-stream ("trace") format.
//
// A trace is a flat sequence of records { uint32 op; uint32 n; double args[n]; } — one record per call
// of the reference C ABI (include/ncr_b200.h §1).  Texture arguments are slot numbers (stored as a
// double) resolved against the texture table handed to the replayer.  Writers: trace.py.  Readers:
// NcrSubmitTrace (api.cu, product) and libnativecpurenderer_b200/csrc/ncr_replay.cpp (drives any library through dlsym).
#pragma once
#include <stdint.h>

enum NcrTraceOp : uint32_t {
    NCR_T_SAVE = 1,
    NCR_T_RESTORE = 2,
    NCR_T_SET_TRANSFORM = 3,     // a b c d e f
    NCR_T_APPLY_TRANSFORM = 4,   // a b c d e f
    NCR_T_SCALE = 5,             // sx sy
    NCR_T_TRANSLATE = 6,         // tx ty
    NCR_T_ROTATE = 7,            // angle
    NCR_T_SET_CT = 8,            // r g b a
    NCR_T_APPLY_CT = 9,          // r g b a
    NCR_T_SET_COLOR = 10,        // r g b a
    NCR_T_FILL_COLOR = 11,       // r g b a
    NCR_T_DRAW_TEXTURE = 12,     // slot x y w h
    NCR_T_DRAW_SPLIT = 13,       // slot x y w h uS uE vS vE
    NCR_T_DRAW_RECT = 14,        // x y w h r g b a
    NCR_T_DRAW_LINE = 15,        // x1 y1 x2 y2 width r g b a
    NCR_T_DRAW_CIRCLE = 16,      // x y radius r g b a
    NCR_T_DRAW_GRD = 17,         // x y w h top(rgba) bottom(rgba)
    NCR_T_SET_PIXEL = 18,        // x y r g b a
    NCR_T_APPLY_PIXEL = 19,      // x y r g b a
    NCR_T_PRESENT = 20,          // end of frame: GetBufferAsUInt8 into the replayer's frame buffer
    // extensions (product only)
    NCR_T_CLIP_SET = 32,         // x y w h
    NCR_T_CLIP_CLEAR = 33,
    NCR_T_SAMPLING = 34,         // mode
    NCR_T_FILL_POLY = 35,        // r g b a x0 y0 x1 y1 ...
    NCR_T_DRAW_PERSP = 36,       // slot h0..h8 x y w h
};

struct NcrTraceRec {
    uint32_t op;
    uint32_t n;
};
