// Exact coverage classification of one command against one 16x8-px region (used by ncr_bin_fine).
//
//   0  no pixel of the region can be touched by the command: the entry is dropped from the region's list;
//   1  some pixels may be touched: the composite runs the command's general (per-pixel tested) path;
//   2  EVERY pixel of the region is touched (box covers the region and every pixel passes the command's coverage test):
//      the composite may run the straight-line interior path, which evaluates no test at all.
//
// For the ops whose coverage is the four inclusive bounds of reference cpp:765-768 on the inverse-mapped position
//     X(i,j) = fl(fl(fl(inv0*i) + fl(inv2*j)) + inv4)          (TransformPointFromMatrix, cpp:451-452)
// X is monotone in i and in j separately (IEEE rounding is monotone), so its extrema over a pixel block are attained at
// block corners chosen by the signs of inv0 and inv2.  Evaluating the SAME expression there bounds what every pixel of the
// block computes: max X < x  =>  every pixel fails `invX < x -> continue`; min X >= x and max X <= x+w  =>  every pixel passes
// both x tests (a pixel whose X is NaN also passes them in the reference: NaN compares false).  No tolerance is involved; a
// NaN bound compares false, so it neither rejects nor promotes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ncr_cmd.h"
#include "pixel_math.cuh"

struct NcrQuadRange {
    double x_min, x_max, y_min, y_max;
};

__device__ __forceinline__ NcrQuadRange ncr_quad_range(const double2 m01, const double2 m23, const double2 m45, int i0, int i1,
                                                       int j0, int j1) {
    const double fi0 = (double)i0, fi1 = (double)i1, fj0 = (double)j0, fj1 = (double)j1;
    NcrQuadRange q;
    // X = (inv0*i + inv2*j) + inv4
    const double xa_hi = MUL(m01.x, m01.x >= 0.0 ? fi1 : fi0), xa_lo = MUL(m01.x, m01.x >= 0.0 ? fi0 : fi1);
    const double xb_hi = MUL(m23.x, m23.x >= 0.0 ? fj1 : fj0), xb_lo = MUL(m23.x, m23.x >= 0.0 ? fj0 : fj1);
    q.x_max = ADD(ADD(xa_hi, xb_hi), m45.x);
    q.x_min = ADD(ADD(xa_lo, xb_lo), m45.x);
    // Y = (inv1*i + inv3*j) + inv5
    const double ya_hi = MUL(m01.y, m01.y >= 0.0 ? fi1 : fi0), ya_lo = MUL(m01.y, m01.y >= 0.0 ? fi0 : fi1);
    const double yb_hi = MUL(m23.y, m23.y >= 0.0 ? fj1 : fj0), yb_lo = MUL(m23.y, m23.y >= 0.0 ? fj0 : fj1);
    q.y_max = ADD(ADD(ya_hi, yb_hi), m45.y);
    q.y_min = ADD(ADD(ya_lo, yb_lo), m45.y);
    return q;
}

// Classifies command `c` (pixel box `box` = l, r, t, b) against BOTH regions of the tile whose top-left pixel is (x0, y0):
// returns, per region (top in bits 0..2, bottom in bits 3..5), code | covers << 2 where `covers` = the box contains the whole region,
// and in the NCR_ENTRY_HINTS bits the command's dispatch hints (the same for both regions).
// The caller has already established that the box intersects the tile.
__device__ __forceinline__ uint32_t ncr_region_codes(const NcrCmd* __restrict__ c, const int4 box, int x0, int y0) {
    uint32_t hit[2], covers[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int ry0 = y0 + h * NCR_REGION_H;
        hit[h] = box.z < ry0 + NCR_REGION_H && box.w > ry0 && box.x < x0 + NCR_REGION_W && box.y > x0;
        covers[h] = box.x <= x0 && box.y >= x0 + NCR_REGION_W && box.z <= ry0 && box.w >= ry0 + NCR_REGION_H;
    }
    if (!(hit[0] | hit[1])) return 0u;
    const uint2 of = __ldg((const uint2*)&c->op);   // op, flags
    const uint32_t op = of.x;
    const uint32_t hints = ((of.y & NCR_F_FAST_AFFINE) ? NCR_ENTRY_FAST_AFFINE : 0u) | (op == NCR_OP_TEX_SPLIT ? NCR_ENTRY_SPLIT : 0u) |
                           ((of.y & NCR_F_CT_RGB_ONE) ? NCR_ENTRY_RGB_ONE : 0u) | ((of.y & NCR_F_ALPHA_LT1) ? NCR_ENTRY_ALPHA_LT1 : 0u);
    uint32_t code[2] = {hit[0], hit[1]};
    if (op == NCR_OP_FILL_COLOR || op == NCR_OP_SET_COLOR) {
        // coverage is the box itself (cpp:643-657, 682-691)
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (covers[h]) code[h] = 2u;
    } else if (op == NCR_OP_TEX_IDENT) {
        // cpp:741-745: pixel (i, j) is drawn iff i >= (i64)x, (f64)i < x + width, likewise j (p[0], p[1] = (f64)(i64)x, y)
        const double2 lo = __ldg((const double2*)&c->p[0]);
        const double2 hi = __ldg((const double2*)&c->xw);
        const bool xin = (double)x0 >= lo.x && (double)(x0 + NCR_REGION_W - 1) < hi.x;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ry0 = y0 + h * NCR_REGION_H;
            if (covers[h] && xin && (double)ry0 >= lo.y && (double)(ry0 + NCR_REGION_H - 1) < hi.y) code[h] = 2u;
        }
    } else if (op == NCR_OP_TEX || op == NCR_OP_TEX_SPLIT || op == NCR_OP_RECT || op == NCR_OP_GRAD) {
        const double2 m01 = __ldg((const double2*)&c->inv[0]), m23 = __ldg((const double2*)&c->inv[2]);
        const double2 m45 = __ldg((const double2*)&c->inv[4]);
        const double2 lo = __ldg((const double2*)&c->x), hi = __ldg((const double2*)&c->xw);
        const int i0 = max(x0, box.x), i1 = min(x0 + NCR_REGION_W, box.y) - 1;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (!hit[h]) continue;
            const int ry0 = y0 + h * NCR_REGION_H;
            const NcrQuadRange q = ncr_quad_range(m01, m23, m45, i0, i1, max(ry0, box.z), min(ry0 + NCR_REGION_H, box.w) - 1);
            if (q.x_max < lo.x || q.x_min > hi.x || q.y_max < lo.y || q.y_min > hi.y) code[h] = 0u;
            else if (covers[h] && q.x_min >= lo.x && q.x_max <= hi.x && q.y_min >= lo.y && q.y_max <= hi.y) code[h] = 2u;
        }
    }
    return (code[0] | (code[0] ? covers[0] << 2 : 0u)) | ((code[1] | (code[1] ? covers[1] << 2 : 0u)) << 3) | hints;
}
