// Host-side transform / colour state machine and the per-draw-call scalar math.
//
// These are O(1) per call and feed every pixel, so they are evaluated on the host in f64 with the
// reference's expression trees (same operand order, no FMA contraction: the host side is compiled
// with -ffp-contract=off and without -march).  sin/cos come from the host libm, exactly as in the
// reference (cpp:440-441).
#pragma once
#include <math.h>
#include <stdint.h>
#include <algorithm>
#include <vector>

typedef long i64;   // reference h:2 — LP64 `long`, what ctypes c_long maps to
typedef double f64;
typedef unsigned char iu8;

struct NcrState {
    f64 m[6];    // Canvas2D-style [a b c d e f]: x' = a*x + c*y + e ; y' = b*x + d*y + f   (cpp:451-452)
    f64 ct[4];   // multiplicative RGBA colour transform (cpp:525-528)
};

static inline void ncr_state_reset(NcrState& s) {   // cpp:17-26
    s.m[0] = 1; s.m[1] = 0; s.m[2] = 0; s.m[3] = 1; s.m[4] = 0; s.m[5] = 0;
    s.ct[0] = s.ct[1] = s.ct[2] = s.ct[3] = 1;
}

// M <- M * T, cpp:405-410.
static inline void ncr_apply_transform(f64* m, f64 a, f64 b, f64 c, f64 d, f64 e, f64 f) {
    const f64 o0 = m[0], o1 = m[1], o2 = m[2], o3 = m[3], o4 = m[4], o5 = m[5];
    m[0] = o0 * a + o2 * b;
    m[1] = o1 * a + o3 * b;
    m[2] = o0 * c + o2 * d;
    m[3] = o1 * c + o3 * d;
    m[4] = o0 * e + o2 * f + o4;
    m[5] = o1 * e + o3 * f + o5;
}

// cpp:451-452
static inline void ncr_xform_point(const f64* m, f64 x, f64 y, f64* ox, f64* oy) {
    *ox = m[0] * x + m[2] * y + m[4];
    *oy = m[1] * x + m[3] * y + m[5];
}

// cpp:472-492
static inline void ncr_inverse(const f64* m, f64* inv) {
    const f64 a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5];
    const f64 det = a * d - b * c;
    const f64 inv_det = det != 0 ? 1 / det : 1e9;
    inv[0] = d * inv_det;
    inv[1] = -b * inv_det;
    inv[2] = -c * inv_det;
    inv[3] = a * inv_det;
    inv[4] = (c * f - d * e) * inv_det;
    inv[5] = (b * e - a * f) * inv_det;
}

// cpp:551-553 — a signed sum, not a norm.
static inline bool ncr_is_no_transform(const f64* m) {
    return m[0] - 1 + m[1] + m[2] + m[3] - 1 + m[4] + m[5] < 1e-5;
}

// (i64)v as x86-64 compiles it (cvttsd2si): truncation toward zero; NaN and out-of-range values give
// the "integer indefinite" INT64_MIN.  Doing the range test explicitly keeps this defined C++.
static inline i64 ncr_trunc_i64(f64 v) {
    if (!(v >= -9223372036854775808.0 && v < 9223372036854775808.0)) return INT64_MIN;
    return (i64)v;
}

// GetBoarder, cpp:693-718: forward-transform the four corners, truncate min/max, clamp to the canvas.
static inline void ncr_border(const f64* m, f64 x, f64 y, f64 w, f64 h, i64 cw, i64 ch, i64* l, i64* r, i64* t, i64* b) {
    f64 ltx, lty, rtx, rty, lbx, lby, rbx, rby;
    ncr_xform_point(m, x, y, &ltx, &lty);
    ncr_xform_point(m, x + w, y, &rtx, &rty);
    ncr_xform_point(m, x, y + h, &lbx, &lby);
    ncr_xform_point(m, x + w, y + h, &rbx, &rby);
    const i64 L = ncr_trunc_i64(std::min(std::min(ltx, rtx), std::min(lbx, rbx)));
    const i64 R = ncr_trunc_i64(std::max(std::max(ltx, rtx), std::max(lbx, rbx)));
    const i64 T = ncr_trunc_i64(std::min(std::min(lty, rty), std::min(lby, rby)));
    const i64 B = ncr_trunc_i64(std::max(std::max(lty, rty), std::max(lby, rby)));
    *l = std::max((i64)0, std::min(cw, L));
    *r = std::max((i64)0, std::min(cw, R));
    *t = std::max((i64)0, std::min(ch, T));
    *b = std::max((i64)0, std::min(ch, B));
}
