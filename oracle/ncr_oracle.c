/* TEST INFRASTRUCTURE ONLY — never linked into, loaded by or shipped with the product.
 *
 * CPU restatement ("port") of the reference's 2D draw/composite algorithm, written from the per-pixel
 * specification in SURVEY.md §8a / Appendix A and exporting the render-path subset of the reference
 * C ABI (include/ncr_b200.h §1) so the same host code can drive it.  Every function cites the reference
 * lines it restates (cpp = reference src/libNativeCPURenderer.cpp).
 *
 * PARITY PINNED: tests/test_oracle.py checks this file bit-for-bit against the UNMODIFIED reference
 * build (oracle/_ref, built by oracle/Makefile from /root/reference/src) on the known-answer streams
 * K1-K6 of SURVEY.md §8c and on randomised streams, and against the committed hashes in tests/golden/.
 *
 * Deliberate differences from the reference, all outside its defined behaviour:
 *   - one generic row-major rasteriser with a per-op pixel shader instead of one loop nest per primitive
 *     (within a draw every pixel is touched once, so the visiting order cannot change the result);
 *   - new canvases are zero-filled (reference: uninitialised, cpp:15);
 *   - 3-channel textures read alpha = 1.0 (reference: uninitialised variable, cpp:746 + cpp:571);
 *   - out-of-bounds stores/loads of the reference (cpp:510 on the last pixel, 1-px textures) are dropped.
 *
 * Build: gcc -O2 -ffp-contract=off, no -march (oracle/Makefile) — plain IEEE-754 f64, no FMA.
 */
#include <math.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef long i64;
typedef double f64;
typedef unsigned char u8;

typedef struct { f64 m[6]; f64 ct[4]; } State;

typedef struct Canvas {
    i64 w, h;
    int ipp;
    f64* px;
    State st;
    State* stack;
    i64 depth, cap;
    /* extensions (pinned to reference code, see tests/cases.py): clip rect and sampling mode */
    int clip_on;
    i64 cl, cr, ct, cb;
    int bilinear;
} Canvas;

typedef struct Image {
    i64 w, h;
    int ipp;
    f64* px;
    int borrowed;   /* shares a canvas buffer (cpp:382) */
    Canvas* owner;
} Image;

/* ---------------------------------------------------------------- scalar helpers */

/* (i64)v the way x86-64 does it (cvttsd2si): truncation; NaN / out of range -> INT64_MIN. */
static i64 trunc64(f64 v) {
    if (!(v >= -9223372036854775808.0 && v < 9223372036854775808.0)) return INT64_MIN;
    return (i64)v;
}
static i64 clampi(i64 v, i64 lo, i64 hi) { return v < lo ? lo : (v > hi ? hi : v); }
static f64 fmin2(f64 a, f64 b) { return b < a ? b : a; }   /* std::min */
static f64 fmax2(f64 a, f64 b) { return a < b ? b : a; }   /* std::max */

/* cpp:451-452 */
static void map_point(const f64* m, f64 x, f64 y, f64* ox, f64* oy) {
    *ox = m[0] * x + m[2] * y + m[4];
    *oy = m[1] * x + m[3] * y + m[5];
}

/* cpp:405-410: M <- M * T */
static void concat(f64* m, f64 a, f64 b, f64 c, f64 d, f64 e, f64 f) {
    f64 o[6];
    memcpy(o, m, sizeof o);
    m[0] = o[0] * a + o[2] * b;
    m[1] = o[1] * a + o[3] * b;
    m[2] = o[0] * c + o[2] * d;
    m[3] = o[1] * c + o[3] * d;
    m[4] = o[0] * e + o[2] * f + o[4];
    m[5] = o[1] * e + o[3] * f + o[5];
}

/* cpp:472-492 */
static void invert(const f64* m, f64* inv) {
    f64 det = m[0] * m[3] - m[1] * m[2];
    f64 id = det != 0 ? 1 / det : 1e9;
    inv[0] = m[3] * id;
    inv[1] = -m[1] * id;
    inv[2] = -m[2] * id;
    inv[3] = m[0] * id;
    inv[4] = (m[2] * m[5] - m[3] * m[4]) * id;
    inv[5] = (m[1] * m[4] - m[0] * m[5]) * id;
}

/* cpp:693-718 */
typedef struct { i64 l, r, t, b; } Box;
static Box border(const Canvas* c, f64 x, f64 y, f64 w, f64 h) {
    f64 ax, ay, bx, by, cx, cy, dx, dy;
    map_point(c->st.m, x, y, &ax, &ay);
    map_point(c->st.m, x + w, y, &bx, &by);
    map_point(c->st.m, x, y + h, &cx, &cy);
    map_point(c->st.m, x + w, y + h, &dx, &dy);
    Box o;
    o.l = clampi(trunc64(fmin2(fmin2(ax, bx), fmin2(cx, dx))), 0, c->w);
    o.r = clampi(trunc64(fmax2(fmax2(ax, bx), fmax2(cx, dx))), 0, c->w);
    o.t = clampi(trunc64(fmin2(fmin2(ay, by), fmin2(cy, dy))), 0, c->h);
    o.b = clampi(trunc64(fmax2(fmax2(ay, by), fmax2(cy, dy))), 0, c->h);
    return o;
}

/* extension: every draw's pixel box is intersected with the clip rect */
static Box clipped(const Canvas* c, Box b) {
    if (c->clip_on) {
        if (b.l < c->cl) b.l = c->cl;
        if (b.r > c->cr) b.r = c->cr;
        if (b.t < c->ct) b.t = c->ct;
        if (b.b > c->cb) b.b = c->cb;
    }
    return b;
}

/* cpp:515-549 */
static void blend(Canvas* c, i64 i, i64 j, f64 r, f64 g, f64 b, f64 a) {
    if (i < 0 || i >= c->w || j < 0 || j >= c->h) return;
    r *= c->st.ct[0]; g *= c->st.ct[1]; b *= c->st.ct[2]; a *= c->st.ct[3];
    f64* d = c->px + (j * c->w + i) * c->ipp;
    if (a != 1) {
        r = d[0] * (1 - a) + r * a;
        g = d[1] * (1 - a) + g * a;
        b = d[2] * (1 - a) + b * a;
    }
    d[0] = r; d[1] = g; d[2] = b;
    if (c->ipp == 4) d[3] = a;
}

/* cpp:555-573: nearest, clamp to [0, w-2] x [0, h-2], truncate */
static void texel(const Image* t, f64 u, f64 v, f64* out) {
    if (u < 0) u = 0;
    if (u >= t->w - 1) u = t->w - 2;
    if (v < 0) v = 0;
    if (v >= t->h - 1) v = t->h - 2;
    i64 xi = clampi(trunc64(u), 0, t->w - 1), yi = clampi(trunc64(v), 0, t->h - 1);
    const f64* s = t->px + (yi * t->w + xi) * t->ipp;
    out[0] = s[0]; out[1] = s[1]; out[2] = s[2];
    out[3] = t->ipp == 4 ? s[3] : 1.0;
}

/* extension: the four-tap formula the reference keeps commented out at cpp:575-620 (same clamp as cpp:560-563).  PINNED: bit-identical
 * to the reference translation unit compiled with those lines switched on (Makefile: _ref/libNativeCPURenderer_bilinear.so,
 * tests/golden/golden_bilinear.json). */
static void texel_bilinear(const Image* t, f64 u, f64 v, f64* out) {
    if (u < 0) u = 0;
    if (u >= t->w - 1) u = t->w - 2;
    if (v < 0) v = 0;
    if (v >= t->h - 1) v = t->h - 2;
    i64 xi = clampi(trunc64(u), 0, t->w - 1), yi = clampi(trunc64(v), 0, t->h - 1);
    i64 dx = xi + 1 < t->w ? 1 : 0, dy = yi + 1 < t->h ? 1 : 0;
    f64 fu = u - (f64)xi, fv = v - (f64)yi, mu = 1.0 - fu, mv = 1.0 - fv;
    const f64* s0 = t->px + (yi * t->w + xi) * t->ipp;
    const f64* s1 = t->px + (yi * t->w + xi + dx) * t->ipp;
    const f64* s2 = t->px + ((yi + dy) * t->w + xi) * t->ipp;
    const f64* s3 = t->px + ((yi + dy) * t->w + xi + dx) * t->ipp;
    for (int k = 0; k < 4; ++k) {
        f64 c0 = (k < t->ipp) ? s0[k] : 1.0, c1 = (k < t->ipp) ? s1[k] : 1.0;
        f64 c2 = (k < t->ipp) ? s2[k] : 1.0, c3 = (k < t->ipp) ? s3[k] : 1.0;
        out[k] = c0 * mu * mv + c1 * fu * mv + c2 * mu * fv + c3 * fu * fv;
    }
}

/* cpp:822-845 */
static bool inside_poly(const f64* p, int n, f64 x, f64 y) {
    bool in = false;
    for (int i = 0, j = n - 1; i < n; j = i++) {
        f64 xi = p[2 * i], yi = p[2 * i + 1], xj = p[2 * j], yj = p[2 * j + 1];
        if ((yi > y) != (yj > y) && x < (xj - xi) * (y - yi) / (yj - yi) + xi) in = !in;
    }
    return in;
}

/* ---------------------------------------------------------------- generic rasteriser */
enum { K_RECT, K_TEX, K_SPLIT, K_GRAD, K_CIRCLE, K_POLY, K_PERSP };

typedef struct {
    int kind;
    f64 inv[6];
    f64 x, y, xw, yh;
    f64 sx, sy;
    const Image* tex;
    f64 col[4], dcol[4];
    f64 height, radius;
    f64 us, du, vs, dv;
    const f64* pts;
    int npts;
    f64 hom[3];   /* K_PERSP: third row of the inverse homography (inv[] holds the first two rows, row-major) */
} Shader;

static void raster(Canvas* c, Box bx, const Shader* s) {
    bx = clipped(c, bx);
    for (i64 j = bx.t; j < bx.b; ++j) {
        for (i64 i = bx.l; i < bx.r; ++i) {
            f64 X, Y, rgba[4];
            if (s->kind == K_PERSP) {   /* extension (the map is this repo's own spec; the rest is pinned to DrawTexture's loop): rw = 1 / (h6*i + h7*j + h8);
                                         * X = (h0*i + h1*j + h2) * rw, Y = (h3*i + h4*j + h5) * rw — one division per pixel */
                f64 fi = (f64)i, fj = (f64)j;
                f64 hw = s->hom[0] * fi + s->hom[1] * fj + s->hom[2];
                f64 rw = 1.0 / hw;
                X = (s->inv[0] * fi + s->inv[1] * fj + s->inv[2]) * rw;
                Y = (s->inv[3] * fi + s->inv[4] * fj + s->inv[5]) * rw;
                if (!(hw > 0.0)) continue;
            } else {
                map_point(s->inv, (f64)i, (f64)j, &X, &Y);
            }
            if (s->kind == K_CIRCLE) {   /* cpp:939-943 */
                f64 dx = X - s->x, dy = Y - s->y;
                if (sqrt(dx * dx + dy * dy) > s->radius) continue;
                memcpy(rgba, s->col, sizeof rgba);
            } else if (s->kind == K_POLY) {   /* cpp:913 */
                if (!inside_poly(s->pts, s->npts, X, Y)) continue;
                memcpy(rgba, s->col, sizeof rgba);
            } else {
                if (X < s->x || X > s->xw || Y < s->y || Y > s->yh) continue;   /* cpp:765-768 */
                if (s->kind == K_RECT) {
                    memcpy(rgba, s->col, sizeof rgba);
                } else if (s->kind == K_GRAD) {   /* cpp:1308-1312 */
                    f64 p = (Y - s->y) / s->height;
                    for (int k = 0; k < 4; ++k) rgba[k] = s->col[k] + s->dcol[k] * p;
                } else {
                    f64 u = (X - s->x) * s->sx, v = (Y - s->y) * s->sy;   /* cpp:770-771 */
                    if (s->kind == K_SPLIT) {   /* cpp:812-813 */
                        u = (s->us + s->du * u / s->tex->w) * s->tex->w;
                        v = (s->vs + s->dv * v / s->tex->h) * s->tex->h;
                    }
                    if (c->bilinear) texel_bilinear(s->tex, u, v, rgba);
                    else texel(s->tex, u, v, rgba);
                }
            }
            blend(c, i, j, rgba[0], rgba[1], rgba[2], rgba[3]);
        }
    }
}

/* ---------------------------------------------------------------- C ABI: contexts */
long GetBufferSize(Canvas* c) { return c->w * c->h * c->ipp; }   /* cpp:3-5 */

Canvas* CreateRenderContext(long w, long h, bool alpha) {   /* cpp:7-31 */
    Canvas* c = (Canvas*)calloc(1, sizeof *c);
    c->w = w; c->h = h; c->ipp = alpha ? 4 : 3;
    c->px = (f64*)calloc((size_t)(w * h * c->ipp) + 1, sizeof(f64));
    c->st.m[0] = c->st.m[3] = 1;
    c->st.ct[0] = c->st.ct[1] = c->st.ct[2] = c->st.ct[3] = 1;
    return c;
}
void DestroyRenderContext(Canvas* c) { (void)c; }   /* cpp:33-37: no-op */

void ResizeRenderContext(Canvas* c, long w, long h) {   /* cpp:39-45 */
    free(c->px);
    c->w = w; c->h = h;
    c->px = (f64*)calloc((size_t)(w * h * c->ipp) + 1, sizeof(f64));
}

void SaveContextState(Canvas* c) {   /* cpp:277-290 */
    if (c->depth == c->cap) {
        c->cap = c->cap ? c->cap * 2 : 16;
        c->stack = (State*)realloc(c->stack, (size_t)c->cap * sizeof(State));
    }
    c->stack[c->depth++] = c->st;
}
bool RestoreContextState(Canvas* c) {   /* cpp:292-309 */
    if (!c->depth) return false;
    c->st = c->stack[--c->depth];
    return true;
}

void GetBuffer(Canvas* c, f64* out) { memcpy(out, c->px, (size_t)GetBufferSize(c) * sizeof(f64)); }   /* cpp:311-316 */

/* cpp:52-57 as x86-64 gcc compiles it: mulsd, cvttsd2si (32-bit), low byte. */
static u8 quant(f64 v) {
    f64 s = v * 255;
    int t = (fabs(s) < 2147483648.0) ? (int)s : INT32_MIN;
    return (u8)(t & 0xff);
}
void GetBufferAsUInt8(Canvas* c, u8* out) {
    i64 n = GetBufferSize(c);
    for (i64 k = 0; k < n; ++k) out[k] = quant(c->px[k]);
}

/* ---------------------------------------------------------------- C ABI: state */
void SetTransform(Canvas* c, f64 a, f64 b, f64 cc, f64 d, f64 e, f64 f) {
    c->st.m[0] = a; c->st.m[1] = b; c->st.m[2] = cc; c->st.m[3] = d; c->st.m[4] = e; c->st.m[5] = f;
}
void ApplyTransform(Canvas* c, f64 a, f64 b, f64 cc, f64 d, f64 e, f64 f) { concat(c->st.m, a, b, cc, d, e, f); }
void Scale(Canvas* c, f64 sx, f64 sy) { concat(c->st.m, sx, 0, 0, sy, 0, 0); }       /* cpp:420-426 */
void Translate(Canvas* c, f64 tx, f64 ty) { concat(c->st.m, 1, 0, 0, 1, tx, ty); }   /* cpp:428-434 */
void Rotate(Canvas* c, f64 ang) {                                                     /* cpp:436-444 */
    f64 s = sin(ang), co = cos(ang);
    concat(c->st.m, co, s, -s, co, 0, 0);
}
void TransformPoint(Canvas* c, f64 x, f64 y, f64* ox, f64* oy) { map_point(c->st.m, x, y, ox, oy); }
void GetTransform(Canvas* c, f64* out) { memcpy(out, c->st.m, sizeof c->st.m); }
void GetInverseTransform(Canvas* c, f64* out) { invert(c->st.m, out); }
void SetColorTransform(Canvas* c, f64 r, f64 g, f64 b, f64 a) {
    c->st.ct[0] = r; c->st.ct[1] = g; c->st.ct[2] = b; c->st.ct[3] = a;
}
void ApplyColorTransform(Canvas* c, f64 r, f64 g, f64 b, f64 a) {
    c->st.ct[0] *= r; c->st.ct[1] *= g; c->st.ct[2] *= b; c->st.ct[3] *= a;
}

/* ---------------------------------------------------------------- C ABI: pixel writes */
bool SetPixel(Canvas* c, long x, long y, f64 r, f64 g, f64 b, f64 a) {   /* cpp:494-513 */
    if (x < 0 || x >= c->w || y < 0 || y >= c->h) return false;
    i64 at = (y * c->w + x) * c->ipp;
    c->px[at] = r; c->px[at + 1] = g; c->px[at + 2] = b;
    if (at + 3 < GetBufferSize(c)) c->px[at + 3] = a;   /* on 3-channel canvases this is the next pixel's red */
    return true;
}
bool ApplyPixel(Canvas* c, long x, long y, f64 r, f64 g, f64 b, f64 a) {   /* cpp:515-549 */
    if (x < 0 || x >= c->w || y < 0 || y >= c->h) return false;
    blend(c, x, y, r, g, b, a);
    return true;
}
void SetColor(Canvas* c, f64 r, f64 g, f64 b, f64 a) {   /* cpp:643-657 */
    if (r == g && g == b && b == a) {
        i64 n = GetBufferSize(c);
        for (i64 k = 0; k < n; ++k) c->px[k] = r;
        return;
    }
    for (i64 i = 0; i < c->w; ++i)   /* column-major on purpose: on 3-channel canvases the order is observable */
        for (i64 j = 0; j < c->h; ++j) SetPixel(c, i, j, r, g, b, a);
}
void GetColor(Canvas* c, f64 x, f64 y, f64* r, f64* g, f64* b, f64* a) {   /* cpp:659-680 */
    if (x < 0) x = 0;
    if (x >= c->w) x = c->w - 1;
    if (y < 0) y = 0;
    if (y >= c->h) y = c->h - 1;
    const f64* s = c->px + (trunc64(y) * c->w + trunc64(x)) * c->ipp;
    *r = s[0]; *g = s[1]; *b = s[2];
    if (c->ipp == 4) *a = s[3];
}
void FillColor(Canvas* c, f64 r, f64 g, f64 b, f64 a) {   /* cpp:682-691 */
    Box all = {0, c->w, 0, c->h};
    all = clipped(c, all);
    for (i64 j = all.t; j < all.b; ++j)
        for (i64 i = all.l; i < all.r; ++i) blend(c, i, j, r, g, b, a);
}

/* ---------------------------------------------------------------- C ABI: primitives */
/* The texture operand may alias this canvas (cpp:382); the reference then reads pixels it has already
 * modified in x-outer order.  The port snapshots instead, like the product. */
static const Image* operand(Canvas* c, const Image* t, Image* tmp) {
    if (!t->borrowed) return t;
    *tmp = *t;
    tmp->w = t->owner->w; tmp->h = t->owner->h;
    size_t n = (size_t)(tmp->w * tmp->h * tmp->ipp);
    tmp->px = (f64*)malloc((n + 1) * sizeof(f64));
    memcpy(tmp->px, t->owner->px, n * sizeof(f64));
    tmp->borrowed = 2;
    (void)c;
    return tmp;
}
static void release_operand(const Image* t) { if (t->borrowed == 2) free(t->px); }

void DrawTexture(Canvas* c, Image* tex, f64 x, f64 y, f64 w, f64 h) {   /* cpp:720-779 */
    if (w == 0 || h == 0) return;
    Image tmp;
    const Image* t = operand(c, tex, &tmp);
    f64 sx = t->w / w, sy = t->h / h;
    const f64* m = c->st.m;
    if (m[0] - 1 + m[1] + m[2] + m[3] - 1 + m[4] + m[5] < 1e-5) {   /* cpp:551-553, quirk 1 */
        /* cpp:741-750: i from (i64)x while i < x + w; the matrix is ignored; blend() clips */
        i64 i0 = trunc64(x), j0 = trunc64(y);
        f64 xw = x + w, yh = y + h;
        Box all = {0, c->w, 0, c->h};
        all = clipped(c, all);
        for (i64 j = j0 < all.t ? all.t : j0; j < all.b && (f64)j < yh; ++j)
            for (i64 i = i0 < all.l ? all.l : i0; i < all.r && (f64)i < xw; ++i) {
                f64 rgba[4];
                if (c->bilinear) texel_bilinear(t, ((f64)i - x) * sx, ((f64)j - y) * sy, rgba);
                else texel(t, ((f64)i - x) * sx, ((f64)j - y) * sy, rgba);
                blend(c, i, j, rgba[0], rgba[1], rgba[2], rgba[3]);
            }
        release_operand(t);
        return;
    }
    Shader s;
    memset(&s, 0, sizeof s);
    s.kind = K_TEX; s.tex = t;
    invert(m, s.inv);
    s.x = x; s.y = y; s.xw = x + w; s.yh = y + h; s.sx = sx; s.sy = sy;
    raster(c, border(c, x, y, w, h), &s);
    release_operand(t);
}

void DrawSplittedTexture(Canvas* c, Image* tex, f64 x, f64 y, f64 w, f64 h, f64 us, f64 ue, f64 vs, f64 ve) {   /* cpp:781-820 */
    if (w == 0 || h == 0) return;
    Image tmp;
    const Image* t = operand(c, tex, &tmp);
    Shader s;
    memset(&s, 0, sizeof s);
    s.kind = K_SPLIT; s.tex = t;
    invert(c->st.m, s.inv);
    s.x = x; s.y = y; s.xw = x + w; s.yh = y + h; s.sx = t->w / w; s.sy = t->h / h;
    s.us = us; s.du = ue - us; s.vs = vs; s.dv = ve - vs;
    raster(c, border(c, x, y, w, h), &s);
    release_operand(t);
}

void DrawRect(Canvas* c, f64 x, f64 y, f64 w, f64 h, f64 r, f64 g, f64 b, f64 a) {   /* cpp:847-874 */
    if (w <= 0 || h <= 0) return;
    Shader s;
    memset(&s, 0, sizeof s);
    s.kind = K_RECT;
    invert(c->st.m, s.inv);
    s.x = x; s.y = y; s.xw = x + w; s.yh = y + h;
    s.col[0] = r; s.col[1] = g; s.col[2] = b; s.col[3] = a;
    raster(c, border(c, x, y, w, h), &s);
}

void DrawVerticalGrd(Canvas* c, f64 x, f64 y, f64 w, f64 h, f64 tr, f64 tg, f64 tb, f64 ta, f64 br, f64 bg, f64 bb, f64 ba) {   /* cpp:1285-1316 */
    if (w <= 0 || h <= 0) return;
    Shader s;
    memset(&s, 0, sizeof s);
    s.kind = K_GRAD;
    invert(c->st.m, s.inv);
    s.x = x; s.y = y; s.xw = x + w; s.yh = y + h; s.height = h;
    s.col[0] = tr; s.col[1] = tg; s.col[2] = tb; s.col[3] = ta;
    s.dcol[0] = br - tr; s.dcol[1] = bg - tg; s.dcol[2] = bb - tb; s.dcol[3] = ba - ta;
    raster(c, border(c, x, y, w, h), &s);
}

void DrawCircle(Canvas* c, f64 x, f64 y, f64 radius, f64 r, f64 g, f64 b, f64 a) {   /* cpp:920-948 */
    if (radius <= 0) return;
    Shader s;
    memset(&s, 0, sizeof s);
    s.kind = K_CIRCLE;
    invert(c->st.m, s.inv);
    s.x = x; s.y = y; s.radius = radius;
    s.col[0] = r; s.col[1] = g; s.col[2] = b; s.col[3] = a;
    raster(c, border(c, x - radius, y - radius, 2 * radius, 2 * radius), &s);
}

void DrawLine(Canvas* c, f64 x1, f64 y1, f64 x2, f64 y2, f64 width, f64 r, f64 g, f64 b, f64 a) {   /* cpp:876-918 */
    if (width <= 0) return;
    f64 dx = x2 - x1, dy = y2 - y1;
    f64 len = sqrt(dx * dx + dy * dy);
    if (len == 0) return;
    f64 ux = dx / len, uy = dy / len, nx = -uy, ny = ux, hw = width / 2;
    f64 pts[8] = {x1 - nx * hw, y1 - ny * hw, x1 + nx * hw, y1 + ny * hw,
                  x2 + nx * hw, y2 + ny * hw, x2 - nx * hw, y2 - ny * hw};
    Shader s;
    memset(&s, 0, sizeof s);
    s.kind = K_POLY; s.pts = pts; s.npts = 4;
    invert(c->st.m, s.inv);
    s.col[0] = r; s.col[1] = g; s.col[2] = b; s.col[3] = a;
    Box all = {0, c->w, 0, c->h};   /* cpp:908-909: the whole canvas is scanned */
    raster(c, all, &s);
}

/* ---------------------------------------------------------------- C ABI: textures */
static Image* new_image(i64 w, i64 h, int ipp) {
    Image* t = (Image*)calloc(1, sizeof *t);
    t->w = w; t->h = h; t->ipp = ipp;
    t->px = (f64*)calloc((size_t)(w * h * ipp) + 1, sizeof(f64));
    return t;
}
Image* CreateTexture(long w, long h, bool alpha, f64* data) {   /* cpp:318-335 */
    Image* t = new_image(w, h, alpha ? 4 : 3);
    memcpy(t->px, data, (size_t)(w * h * t->ipp) * sizeof(f64));
    return t;
}
Image* CreateTextureUInt8(long w, long h, bool alpha, u8* data) {   /* cpp:337-354 */
    Image* t = new_image(w, h, alpha ? 4 : 3);
    for (i64 k = 0; k < w * h * t->ipp; ++k) t->px[k] = data[k] / 255.0;
    return t;
}
void DestroyTexture(Image* t) { (void)t; }   /* cpp:356-360: no-op */
Image* CreateTextureFromRenderContext(Canvas* c) {   /* cpp:362-375 */
    Image* t = new_image(c->w, c->h, c->ipp);
    memcpy(t->px, c->px, (size_t)GetBufferSize(c) * sizeof(f64));
    return t;
}
Image* CreateTextureFromRenderContextShared(Canvas* c) {   /* cpp:377-384 */
    Image* t = (Image*)calloc(1, sizeof *t);
    t->w = c->w; t->h = c->h; t->ipp = c->ipp; t->borrowed = 1; t->owner = c;
    return t;
}
long GetTextureWidth(Image* t) { return t->borrowed ? t->owner->w : t->w; }
long GetTextureHeight(Image* t) { return t->borrowed ? t->owner->h : t->h; }
bool GetTextureEnableAlpha(Image* t) { return t->ipp == 4; }

Image* ResampleTexture(Image* tex, long w, long h) {   /* cpp:950-976 */
    Image tmp;
    const Image* t = operand(NULL, tex, &tmp);
    Image* o = new_image(w, h, t->ipp);
    for (i64 j = 0; j < h; ++j)
        for (i64 i = 0; i < w; ++i) {
            f64 rgba[4];
            texel(t, (f64)i / w * t->w, (f64)j / h * t->h, rgba);
            memcpy(o->px + (j * w + i) * o->ipp, rgba, (size_t)o->ipp * sizeof(f64));
        }
    release_operand(t);
    return o;
}

/* ---------------------------------------------------------------- C ABI: hit-effect texture (cpp:1318-1440) */
static f64 frac(f64 v) { return v - floor(v); }
static f64 lattice(f64 x, f64 y) { return frac(sin(x * 12.9898 + y * 78.233) * 43758.5453); }   /* cpp:1339-1341 */
static f64 lerp(f64 a, f64 b, f64 t) { return a + (b - a) * t; }
static f64 vnoise(f64 x, f64 y) {   /* cpp:1372-1383 */
    f64 ix = floor(x), iy = floor(y), ux = frac(x), uy = frac(y);
    f64 a = lattice(ix, iy), b = lattice(ix + 1.0, iy + 0.0), c = lattice(ix + 0.0, iy + 1.0), d = lattice(ix + 1.0, iy + 1.0);
    f64 sx = ux * ux * (3.0 - 2.0 * ux), sy = uy * uy * (3.0 - 2.0 * uy);
    return lerp(lerp(a, b, sx), lerp(c, d, sx), sy);
}
void GetMilthmHitEffectPixel(f64 seed, f64 t, f64 x, f64 y, f64* a) {   /* cpp:1385-1411 */
    f64 cx = x - 0.5, cy = y - 0.5;
    f64 rad = sqrt(cx * cx + cy * cy) * 50.0;
    f64 ang = fabs(atan2(cy, cx));
    if (y > 0.5) ang += sin(ang) * 2.0;
    f64 px = rad + seed * 100.0, py = ang + seed * 100.0;
    f64 n = 0.0;
    n += vnoise(px, py) * 0.7;
    n += vnoise(px * 2.0, py * 2.0) * 0.3;
    n += vnoise(px * 4.0, py * 4.0) * 0.1;
    *a = n < t ? 0.0 : 1.0;
}
Image* CreateMilthmHitEffectTexture(Image* mask, f64 seed, f64 t, f64 r, f64 g, f64 b) {   /* cpp:1417-1440 */
    if (mask->ipp != 4 || mask->borrowed) return NULL;
    Image* o = new_image(mask->w, mask->h, 4);
    for (i64 i = 0; i < mask->w; ++i)
        for (i64 j = 0; j < mask->h; ++j) {
            f64 a;
            GetMilthmHitEffectPixel(seed, t, (f64)i / mask->w, (f64)j / mask->h, &a);
            i64 k = (i * mask->h + j) * 4;   /* transposed indexing on both sides, quirk 9 */
            o->px[k] = r; o->px[k + 1] = g; o->px[k + 2] = b;
            o->px[k + 3] = a * mask->px[k + 3];
        }
    return o;
}

long GetVersion(void) { return 1; }

/* ---------------------------------------------------------------- extensions (include/ncr_b200.h section 2).  The reference has none of
 * these as entry points, but each is PINNED to reference code through the builds of oracle/Makefile (tests/cases.py, DESIGN.md
 * section 5): clip rect = the unmodified reference + outside pixels put back; bilinear = its own commented-out sampler; polygon fill
 * = DrawLine's loop; perspective = DrawTexture's mapped loop behind this repo's projective map. */
void NcrSetClipRect(Canvas* c, long x, long y, long w, long h) {
    c->clip_on = 1;
    c->cl = x < 0 ? 0 : x;
    c->ct = y < 0 ? 0 : y;
    c->cr = x + w > c->w ? c->w : x + w;
    c->cb = y + h > c->h ? c->h : y + h;
}
void NcrClearClipRect(Canvas* c) { c->clip_on = 0; }
void NcrSetSampling(Canvas* c, int mode) { c->bilinear = mode == 1; }

void NcrFillPolygon(Canvas* c, const f64* xy, long n, f64 r, f64 g, f64 b, f64 a) {   /* cpp:822-845 rule for N points */
    if (n <= 0) return;
    Shader s;
    memset(&s, 0, sizeof s);
    s.kind = K_POLY; s.pts = xy; s.npts = (int)n;
    invert(c->st.m, s.inv);
    s.col[0] = r; s.col[1] = g; s.col[2] = b; s.col[3] = a;
    Box all = {0, c->w, 0, c->h};
    raster(c, all, &s);
}

void NcrDrawTexturePerspective(Canvas* c, Image* tex, const f64* hinv, f64 x, f64 y, f64 w, f64 h) {
    if (w == 0 || h == 0) return;
    Image tmp;
    const Image* t = operand(c, tex, &tmp);
    Shader s;
    memset(&s, 0, sizeof s);
    s.kind = K_PERSP; s.tex = t;
    for (int k = 0; k < 6; ++k) s.inv[k] = hinv[k];
    s.hom[0] = hinv[6]; s.hom[1] = hinv[7]; s.hom[2] = hinv[8];
    s.x = x; s.y = y; s.xw = x + w; s.yh = y + h; s.sx = t->w / w; s.sy = t->h / h;
    Box all = {0, c->w, 0, c->h};
    raster(c, all, &s);
    release_operand(t);
}

/* ---------------------------------------------------------------- present path
 * PutRendererContextFrame (cpp:232-256) truncates the canvas to u8 and hands it to libswscale:
 *     sws_getContext(w, h, RGBA | RGB24, w, h, YUV420P, SWS_BILINEAR, 0, 0, 0);  sws_scale(...)          (cap size == canvas size)
 * libswscale is an un-vendored third-party dependency of the reference (FFmpeg; the reference pins no version).  This is a
 * restatement of what libswscale computes on x86-64 for exactly that call — its published algorithm (libswscale/input.c
 * rgb24ToY_c / rgb24ToUV_half_c and their rgb32 twins, utils.c initFilter, x86/yuv2yuvX.asm, swscale.c) with the BT.601
 * limited-range table — PINNED against a real build: libswscale 9.1.100 (FFmpeg 8), the copy bundled with this image's
 * opencv-python-headless wheel, driven through ctypes by tests/golden/make_swscale_fixtures.py; bit-exact Y, U and V for
 * RGBA and RGB24 input on every even size >= 8x8 tried (tests/test_oracle.py).  Odd sizes and heights below 8 follow the same
 * formulas with clamped neighbours — this repo's own definition, not pinned.
 *
 *   luma    Y14 = (8414 R + 16519 G + 3208 B + (32 << 14) + (1 << 8)) >> 9;  Y15 = Y14 << 1;  Y = clip8((Y15 + 64) >> 7)
 *   chroma  horizontal: the SUM of each pixel pair (R2 = R[2i] + R[2i+1], ...):
 *               U14 = (-4865 R2 - 9528 G2 + 14392 B2 + (0x4001 << 9)) >> 10,  V14 = (14392 R2 - 12061 G2 - 2332 B2 + (0x4001 << 9)) >> 10
 *               U15 = min(U14 << 1, 32767)
 *           vertical: bilinear 2:1 = taps {512, 1536, 1536, 512} / 4096 on rows 2c-1 .. 2c+2, taps that fall outside the
 *               image folded onto the edge row (initFilter).  Without SWS_ACCURATE_RND — the reference sets no flag — x86
 *               runs the 16-bit vertical scaler: acc = ((64 + 8 * 3) >> 4) + sum_k ((U15[k] * coeff[k]) >> 16), U = clip8(acc >> 3)
 *               (pmulhw per tap; the bias term compensates its truncation); the LAST chroma row is produced by the C
 *               scaler (swscale.c switches for dstY >= dstH - 2): U = clip8(((64 << 12) + sum_k U15[k] * coeff[k]) >> 19). */
static u8 clip8(i64 v) { return (u8)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

/* The scaling branch (cap size != canvas size) goes through the same machinery with non-trivial filters, so the conversion is
 * written once, for w x h -> dw x dh:
 *   filters    libswscale/utils.c initFilter for SWS_BILINEAR: xInc = ((src << 16) + (dst >> 1)) / dst; |xInc - 65536| < 10 -> one unit
 *              tap per output; else filterSize = 1 + 2 (upscale) or 1 + (2 src + dst - 1) / dst, taps = max(0, 2^30 - |distance|)
 *              scaled by dst/src when shrinking, near-zero ends trimmed (cutoff 0.002), size rounded up to the SIMD alignment
 *              (4 horizontal, 2 vertical on x86), out-of-image taps folded onto the edge, normalised to 2^14 (horizontal) or 2^12
 *              (vertical) with error diffusion.  Chroma source is the pixel-PAIR sum when (dw >> 1) <= (w >> 1), else full width.
 *   horizontal out15 = min((sum_j src14[pos + j] * f[j]) >> 13, 32767)                                     (x86 hscale14to15)
 *   vertical   one tap: (v15 + 64) >> 7; else the 16-bit SIMD scaler acc = ((64 + 8 (n - 1)) >> 4) + sum_j ((v15[j] * f[j]) >> 16),
 *              out = acc >> 3 — except the last two luma rows and the last chroma row, which the C scaler produces:
 *              ((64 << 12) + sum_j v15[j] * f[j]) >> 19.
 * Pinned against libswscale 9.1.100 for shrinking and enlarging, odd and even sizes (tests/test_oracle.py); with an odd SOURCE width
 * on the pair-sum path libswscale reads one pixel past the row end — here the edge pixel is repeated (not pinned). */
typedef struct { int size; int n; int* pos; int* coef; } SwsFilter;   /* coef[i * size + j] */

static i64 sws_cdiv(i64 a, i64 b) { return a / b; }   /* C division truncates toward zero, as in libswscale */
static i64 sws_rounded_div(i64 a, i64 b) { return (a >= 0 ? a + (b >> 1) : a - (b >> 1)) / b; }
static int sws_log2(i64 v) { int n = 0; while (v > 1) { v >>= 1; ++n; } return n; }

static int sws_init_filter(SwsFilter* F, i64 xInc, int srcW, int dstW, int align, i64 one) {
    const i64 fone = (i64)1 << (54 - (sws_log2(srcW / dstW) < 8 ? sws_log2(srcW / dstW) : 8));
    int fs, i, j;
    i64* f;
    int* pos = (int*)malloc((size_t)dstW * sizeof(int));
    if (!pos) return -1;
    if (llabs(xInc - 0x10000) < 10) {   /* unscaled (source and destination sample positions coincide for every plane here) */
        fs = 1;
        f = (i64*)malloc((size_t)dstW * sizeof(i64));
        if (!f) { free(pos); return -1; }
        for (i = 0; i < dstW; ++i) { f[i] = fone; pos[i] = i; }
    } else {
        i64 xDstInSrc = ((128 * xInc) >> 7) - ((128 * (i64)0x10000) >> 7);   /* srcPos = dstPos = 128 */
        fs = (xInc <= (1 << 16)) ? 3 : (int)(1 + (2 * (i64)srcW + dstW - 1) / dstW);
        if (fs > srcW - 2) fs = srcW - 2;
        if (fs < 1) fs = 1;
        f = (i64*)malloc((size_t)dstW * fs * sizeof(i64));
        if (!f) { free(pos); return -1; }
        for (i = 0; i < dstW; ++i) {
            i64 xx = sws_cdiv(xDstInSrc - (fs - 2) * ((i64)1 << 16), (i64)1 << 17);
            pos[i] = (int)xx;
            for (j = 0; j < fs; ++j) {
                i64 d = llabs(xx * ((i64)1 << 17) - xDstInSrc) << 13, coeff;
                if (xInc > (1 << 16)) d = sws_cdiv(d * dstW, srcW);
                coeff = ((i64)1 << 30) - d;
                if (coeff < 0) coeff = 0;
                coeff *= fone >> 30;
                f[i * fs + j] = coeff;
                ++xx;
            }
            xDstInSrc += 2 * xInc;
        }
    }
    /* trim near-zero ends (SWS_MAX_REDUCE_CUTOFF = 0.002) and find the common size */
    int minSize = 0;
    for (i = dstW - 1; i >= 0; --i) {
        int mn = fs;
        double cut = 0;
        for (j = 0; j < fs; ++j) {
            cut += (double)llabs(f[i * fs]);
            if (cut > 0.002 * (double)fone) break;
            if (i < dstW - 1 && pos[i] >= pos[i + 1]) break;
            memmove(&f[i * fs], &f[i * fs + 1], (size_t)(fs - 1) * sizeof(i64));
            f[i * fs + fs - 1] = 0;
            pos[i]++;
        }
        cut = 0;
        for (j = fs - 1; j > 0; --j) {
            cut += (double)llabs(f[i * fs + j]);
            if (cut > 0.002 * (double)fone) break;
            --mn;
        }
        if (mn > minSize) minSize = mn;
    }
    if (minSize == 1 && align == 2) align = 1;
    const int size = (minSize + (align - 1)) & ~(align - 1);
    i64* g = (i64*)calloc((size_t)dstW * size, sizeof(i64));
    int* coef = (int*)malloc((size_t)dstW * size * sizeof(int));
    if (!g || !coef) { free(f); free(pos); free(g); free(coef); return -1; }
    for (i = 0; i < dstW; ++i)
        for (j = 0; j < size; ++j) g[i * size + j] = j < fs ? f[i * fs + j] : 0;
    free(f);
    for (i = 0; i < dstW; ++i) {   /* fold taps that fall outside the image onto the edge */
        i64* r = &g[i * size];
        if (pos[i] < 0) {
            for (j = 1; j < size; ++j) {
                int left = j + pos[i] > 0 ? j + pos[i] : 0;
                r[left] += r[j];
                r[j] = 0;
            }
            pos[i] = 0;
        }
        if (pos[i] + size > srcW) {
            int shift = pos[i] + (size - srcW < 0 ? size - srcW : 0);
            i64 acc = 0;
            for (j = size - 1; j >= 0; --j)
                if (pos[i] + j >= srcW) { acc += r[j]; r[j] = 0; }
            for (j = size - 1; j >= 0; --j) r[j] = j < shift ? 0 : r[j - shift];
            pos[i] -= shift;
            r[srcW - 1 - pos[i]] += acc;
        }
    }
    for (i = 0; i < dstW; ++i) {   /* normalise to `one` with error diffusion */
        i64 err = 0, sum = 0;
        for (j = 0; j < size; ++j) sum += g[i * size + j];
        sum = (sum + one / 2) / one;
        if (!sum) sum = 1;
        for (j = 0; j < size; ++j) {
            i64 v = g[i * size + j] + err;
            i64 q = sws_rounded_div(v, sum);
            coef[i * size + j] = (int)q;
            err = v - q * sum;
        }
    }
    free(g);
    F->size = size; F->n = dstW; F->pos = pos; F->coef = coef;
    return 0;
}
static void sws_free_filter(SwsFilter* F) { free(F->pos); free(F->coef); F->pos = F->coef = NULL; }

/* one plane: src14 [srcH][srcW] -> horizontal (H) -> vertical (V) -> dst [dstH][dstW]; c_rows = rows at the bottom done by the C scaler */
static int sws_plane(const int* src14, int srcW, int srcH, const SwsFilter* H, const SwsFilter* V, int c_rows, u8* dst) {
    const int dstW = H->n, dstH = V->n;
    int* mid = (int*)malloc((size_t)srcH * dstW * sizeof(int));
    if (!mid) return -1;
    for (int y = 0; y < srcH; ++y)
        for (int x = 0; x < dstW; ++x) {
            i64 acc = 0;
            for (int j = 0; j < H->size; ++j) {
                int sx = H->pos[x] + j;
                if (sx > srcW - 1) sx = srcW - 1;   /* only ever multiplies a zero coefficient */
                acc += (i64)src14[y * srcW + sx] * H->coef[x * H->size + j];
            }
            acc >>= 13;
            mid[y * dstW + x] = (int)(acc > 32767 ? 32767 : acc);
        }
    for (int y = 0; y < dstH; ++y)
        for (int x = 0; x < dstW; ++x) {
            i64 acc;
            if (V->size == 1) {
                acc = ((i64)mid[V->pos[y] * dstW + x] + 64) >> 7;
            } else if (y >= dstH - c_rows) {
                acc = (i64)64 << 12;
                for (int j = 0; j < V->size; ++j) {
                    int sy = V->pos[y] + j;
                    if (sy > srcH - 1) sy = srcH - 1;
                    acc += (i64)mid[sy * dstW + x] * V->coef[y * V->size + j];
                }
                acc >>= 19;
            } else {
                acc = (64 + 8 * (V->size - 1)) >> 4;
                for (int j = 0; j < V->size; ++j) {
                    int sy = V->pos[y] + j;
                    if (sy > srcH - 1) sy = srcH - 1;
                    acc += ((i64)mid[sy * dstW + x] * V->coef[y * V->size + j]) >> 16;
                }
                acc >>= 3;
            }
            dst[y * dstW + x] = clip8(acc);
        }
    free(mid);
    return 0;
}

static long yuv420p_of(Canvas* c, i64 dw, i64 dh, u8* out) {
    const i64 w = c->w, h = c->h, cdw = (dw + 1) / 2, cdh = (dh + 1) / 2;
    const int ipp = c->ipp;
    if (w <= 0 || h <= 0 || dw <= 0 || dh <= 0) return 0;
    const int half = (dw >> 1) <= (w >> 1);          /* chroma input: pixel-pair sums, else full width */
    const i64 cw = half ? (w + 1) / 2 : w;
    u8* img = (u8*)malloc((size_t)(w * h * ipp));
    int* y14 = (int*)malloc((size_t)(w * h) * sizeof(int));
    int* u14 = (int*)malloc((size_t)(cw * h) * sizeof(int));
    int* v14 = (int*)malloc((size_t)(cw * h) * sizeof(int));
    SwsFilter hl = {0}, vl = {0}, hc = {0}, vc = {0};
    long rc = -1;
    if (!img || !y14 || !u14 || !v14) goto done;
    GetBufferAsUInt8(c, img);
    for (i64 j = 0; j < h; ++j) {
        for (i64 i = 0; i < w; ++i) {
            const u8* q = img + (j * w + i) * ipp;
            y14[j * w + i] = (int)((8414 * (i64)q[0] + 16519 * (i64)q[1] + 3208 * (i64)q[2] + (32 << 14) + (1 << 8)) >> 9);
        }
        for (i64 i = 0; i < cw; ++i) {
            if (half) {
                const i64 x0 = 2 * i, x1 = (2 * i + 1 < w) ? 2 * i + 1 : w - 1;
                const u8 *p0 = img + (j * w + x0) * ipp, *p1 = img + (j * w + x1) * ipp;
                const i64 r2 = p0[0] + p1[0], g2 = p0[1] + p1[1], b2 = p0[2] + p1[2];
                u14[j * cw + i] = (int)((-4865 * r2 - 9528 * g2 + 14392 * b2 + ((i64)0x4001 << 9)) >> 10);
                v14[j * cw + i] = (int)((14392 * r2 - 12061 * g2 - 2332 * b2 + ((i64)0x4001 << 9)) >> 10);
            } else {
                const u8* q = img + (j * w + i) * ipp;
                u14[j * cw + i] = (int)((-4865 * (i64)q[0] - 9528 * (i64)q[1] + 14392 * (i64)q[2] + (256 << 14) + (1 << 8)) >> 9);
                v14[j * cw + i] = (int)((14392 * (i64)q[0] - 12061 * (i64)q[1] - 2332 * (i64)q[2] + (256 << 14) + (1 << 8)) >> 9);
            }
        }
    }
    if (sws_init_filter(&hl, (((i64)w << 16) + (dw >> 1)) / dw, (int)w, (int)dw, 4, 1 << 14) ||
        sws_init_filter(&vl, (((i64)h << 16) + (dh >> 1)) / dh, (int)h, (int)dh, 2, 1 << 12) ||
        sws_init_filter(&hc, (((i64)cw << 16) + (cdw >> 1)) / cdw, (int)cw, (int)cdw, 4, 1 << 14) ||
        sws_init_filter(&vc, (((i64)h << 16) + (cdh >> 1)) / cdh, (int)h, (int)cdh, 2, 1 << 12))
        goto done;
    {
        u8 *Y = out, *U = out + dw * dh, *V = U + cdw * cdh;
        if (sws_plane(y14, (int)w, (int)h, &hl, &vl, 2, Y) || sws_plane(u14, (int)cw, (int)h, &hc, &vc, 1, U) ||
            sws_plane(v14, (int)cw, (int)h, &hc, &vc, 1, V))
            goto done;
    }
    rc = (long)(dw * dh + 2 * cdw * cdh);
done:
    sws_free_filter(&hl); sws_free_filter(&vl); sws_free_filter(&hc); sws_free_filter(&vc);
    free(img); free(y14); free(u14); free(v14);
    return rc;
}

long NcrYUV420PSize(Canvas* c) {
    if (c->w <= 0 || c->h <= 0) return 0;
    return (long)(c->w * c->h + 2 * ((c->w + 1) / 2) * ((c->h + 1) / 2));
}
long NcrGetBufferAsYUV420P(Canvas* c, u8* out) { return yuv420p_of(c, c->w, c->h, out); }
/* cap size != canvas size: the sws_scale resize of cpp:241-256 */
long NcrGetBufferAsYUV420PScaled(Canvas* c, long dst_w, long dst_h, u8* out) { return yuv420p_of(c, dst_w, dst_h, out); }
