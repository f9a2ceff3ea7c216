// Draw-command encoding shared by the host recorder and the sm_100a kernels.
//
// One NcrCmd is the snapshot the reference takes implicitly when a draw call runs
// (reference src/libNativeCPURenderer.cpp:720-948, 1285-1316): the inverse matrix
// (cpp:472-492), the GetBoarder pixel box (cpp:693-718), the colour transform
// (cpp:525-528) and the call's own parameters.  Everything that is a pure function
// of the call arguments is evaluated ONCE on the host with the reference's expression
// trees; the kernels evaluate only the per-pixel part.
#pragma once
#include <stdint.h>

enum NcrOp : uint32_t {
    NCR_OP_NOP = 0,
    NCR_OP_SET_COLOR = 1,    // cpp:643-657  store (p[0..3]) to every pixel
    NCR_OP_FILL_COLOR = 2,   // cpp:682-691  APPLY (p[0..3]) on every pixel
    NCR_OP_TEX_IDENT = 3,    // cpp:731-751  DrawTexture, "no transform" path
    NCR_OP_TEX = 4,          // cpp:753-778  DrawTexture, inverse-mapped path
    NCR_OP_TEX_SPLIT = 5,    // cpp:781-820  DrawSplittedTexture
    NCR_OP_RECT = 6,         // cpp:847-874
    NCR_OP_GRAD = 7,         // cpp:1285-1316 DrawVerticalGrd
    NCR_OP_CIRCLE = 8,       // cpp:920-948
    NCR_OP_POLY = 9,         // cpp:876-918 DrawLine's 4-gon / extension: N-gon fill
    NCR_OP_SET_PIXEL = 10,   // cpp:494-513
    NCR_OP_APPLY_PIXEL = 11, // cpp:515-549
    NCR_OP_TEX_PERSP = 12,   // extension: projective inverse map (this repo's spec), then DrawTexture's mapped loop
    NCR_OP_COUNT
};

// NcrCmd.flags
enum : uint32_t {
    NCR_F_TEX_ALPHA = 1u << 0,    // texture has 4 channels (else 3; alpha then reads as 1.0)
    NCR_F_TEX_F64 = 1u << 1,      // texels are f64 (else u8 decoded through the k/255.0 table)
    NCR_F_BILINEAR = 1u << 2,     // extension: the cpp:575-620 four-tap formula (pinned to the reference's commented-out code)
    NCR_F_CLIP = 1u << 3,         // extension: box already intersected with the clip rect on the host
    NCR_F_RGB_SPILL = 1u << 4,    // SetColor on a 3-channel canvas, non-uniform colour: alpha lands in the next element (cpp:510)
    NCR_F_ONLY_RED = 1u << 5,     // SET_PIXEL that writes only the red element (the cpp:510 spill of the previous pixel)
    NCR_F_CT_RGB_ONE = 1u << 6,   // ct[0..2] are all exactly 1.0: v * 1.0 == v, the multiplies are skipped
    NCR_F_SPLIT_POW2 = 1u << 7,   // TEX_SPLIT: texture width and height are powers of two; p[6], p[7] = 1/w, 1/h (exact)
    NCR_F_TEX_FAST = 1u << 8,     // RGBA8 texels, nearest sampling, fewer than 2^31 texels: inlined sampler
    NCR_F_FAST_AFFINE = 1u << 9,  // NCR_OP_TEX / NCR_OP_TEX_SPLIT with NCR_F_TEX_FAST: the composite's first-tested path
    NCR_F_ALPHA_LT1 = 1u << 10,   // RGBA8 texture and 0 <= ct[3] < 1: a = fl(texel_a * ct3) <= ct3 < 1 for every texel (texel_a = k/255 <= 1,
                                  // rounding is monotone), so `a != 1` (cpp:533) is true on every pixel: no test, no plain-store case
};

struct alignas(16) NcrCmd {
    uint32_t op;
    uint32_t flags;
    int32_t l, r, t, b;       // pixel box [l,r) x [t,b): loop bounds of the reference (bbox ops) or a conservative cover
    int32_t tex_w, tex_h;
    const void* tex;          // device pointer to texels, row-major [y][x][ipp]
    uint32_t aux_off;         // NCR_OP_POLY: first point in the aux f64 array (x0,y0,x1,y1,...)
    uint32_t aux_n;           // NCR_OP_POLY: number of points
    double inv[6];            // inverse transform (cpp:483-491)
    double x, y, xw, yh;      // x, y, x+width, y+height (the four inclusive bounds, cpp:765-768)
    double sx, sy;            // scaleX, scaleY (cpp:728-729); GRAD: sy = height; CIRCLE: sx = radius
    double ct[4];             // colour transform snapshot
    double p[8];              // op-specific, see the recorder
};
static_assert(sizeof(NcrCmd) == 240, "NcrCmd layout");
#define NCR_CMD_WORDS16 (sizeof(NcrCmd) / 16)

// Compact per-command box, read by the binning kernels (16 B, coalesced).
struct alignas(16) NcrBox {
    int32_t l, r, t, b;
};

#define NCR_TILE 16            // binning tile edge in pixels
#define NCR_COARSE 8           // coarse bin edge in tiles (128 px)
// The composite's work unit is a REGION: the top or bottom 16x8-px half of a tile (region = 2 * tile + half).  ncr_bin_fine
// writes one ordered command list per region.  A list entry is a command index plus tag bits; bit 31 marks the command as INTERIOR to
// the region: every pixel of the region provably passes the command's box and coverage tests (see region_code()).
#define NCR_REGION_W 16
#define NCR_REGION_H 8
#define NCR_REGIONS_PER_TILE 2
#define NCR_ENTRY_INTERIOR 0x80000000u
#define NCR_ENTRY_COVERS 0x40000000u   // the command's pixel box contains the whole region (no box-membership test per pixel); set with INTERIOR too
// Dispatch hints, copied by ncr_bin_fine from the command's op / flags (it reads them anyway): the composite picks the code path
// of a hot command from the entry it already holds in a register, instead of waiting for the staged command's flag word.
#define NCR_ENTRY_FAST_AFFINE 0x20000000u   // NCR_F_FAST_AFFINE
#define NCR_ENTRY_SPLIT 0x10000000u         // op == NCR_OP_TEX_SPLIT
#define NCR_ENTRY_RGB_ONE 0x08000000u       // NCR_F_CT_RGB_ONE
#define NCR_ENTRY_ALPHA_LT1 0x04000000u     // NCR_F_ALPHA_LT1
#define NCR_ENTRY_HINTS 0x3c000000u
#define NCR_ENTRY_INDEX 0x03ffffffu         // 2^26 commands per batch (the recorder submits at 2^20)
static_assert((NCR_ENTRY_INDEX & (NCR_ENTRY_HINTS | NCR_ENTRY_INTERIOR | NCR_ENTRY_COVERS)) == 0 &&
                  (NCR_ENTRY_INDEX | NCR_ENTRY_HINTS | NCR_ENTRY_INTERIOR | NCR_ENTRY_COVERS) == 0xffffffffu &&
                  (NCR_ENTRY_FAST_AFFINE | NCR_ENTRY_SPLIT | NCR_ENTRY_RGB_ONE | NCR_ENTRY_ALPHA_LT1) == NCR_ENTRY_HINTS,
              "list-entry bit fields partition the word");

struct NcrFrameDims {
    int32_t w, h, ipp;
    int32_t tiles_x, tiles_y;
    int32_t bins_x, bins_y;
};

// Launch-side description of one flush (one canvas, one ordered command batch).
struct NcrFlushArgs {
    NcrFrameDims d;
    double* fb;                 // canonical f64 canvas [h][w][ipp]
    unsigned char* u8_out;      // fused (iu8)(v*255) image, or nullptr
    unsigned char* yuv_out;     // YUV 4:2:0 planes of that image (present path; ncr_yuv420p runs right after the composite), or nullptr
    const NcrCmd* cmds;
    const NcrBox* boxes;
    const uint32_t* binboxes;   // per command: first/last 128-px bin touched per axis, 4 x u8 (x0 x1 y0 y1), or nullptr
    const double* aux;
    uint32_t n_cmds;
    uint32_t load_fb;           // 0: every tile's list starts with SET_COLOR, do not read fb
    uint32_t* coarse_list;      // capacity coarse_cap; nullptr: small batch — ncr_bin_coarse is skipped and ncr_bin_fine scans the commands directly
    uint32_t* coarse_off;       // [bins] offset, [bins] count
    uint32_t* fine_list;        // capacity fine_cap; entries: command index | NCR_ENTRY_INTERIOR
    uint32_t* fine_off;         // per region: {offset, count} (uint2[regions])
    uint32_t* cursors;          // [0] coarse cursor, [1] fine cursor, [2..3] blended-pixel counter (u64), [4] overflow flag,
                                // [5] composite work counter (regions handed out), [6] list entries tagged interior (statistics)
    uint32_t coarse_cap, fine_cap;
    uint32_t count_pixels;      // stats mode: count APPLY executions
    uint32_t write_fb;          // 0: present-only flush — the f64 canvas is NOT written back (the host marks it stale and
                                // re-runs this batch with write_fb = 1 if anyone ever reads it; see api.cu materialize())
    uint32_t prefetch;          // composite variant: 1 = next region's header / list / first command fetched during the current one
    // Experiment X5 (profiles/README.md; only read by kernels built with NCR_TMA_IDENT, only filled when NCR_TMA=1): the batch's
    // first full-resolution identity-path DrawTexture of an RGBA8 texture at integer (x, y) — milrenderer's background — whose
    // 16x8 texel box per region can be staged by the TMA unit one region ahead.
    const void* tma_map;        // CUtensorMap of the texture (device memory), or nullptr
    int32_t tma_cmd;            // index of that command in cmds
    int32_t tma_x, tma_y;       // its integer destination origin
    int32_t tma_w, tma_h;       // texture size in texels
};
