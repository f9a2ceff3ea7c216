// TEST INFRASTRUCTURE ONLY (see Makefile).  Pins the N-gon fill extension (NcrFillPolygon) and the exported ApplyPixel to the
// reference's own functions.
//
// The reference has no polygon entry point, but DrawLine (reference src/libNativeCPURenderer.cpp:876-918) IS a polygon fill: it
// builds the stroke's four corners and then, for every canvas pixel, inverse-maps it (GetInverseTransform +
// TransformPointFromMatrix), tests it with pointInPolygon (cpp:822-845, even-odd rule) and calls ApplyPixel.  This translation
// unit compiles the reference source FROM WHERE IT LIES (-I$(REF_SRC); nothing of it is copied) and adds one entry point that
// runs that same loop, through those same reference functions, on the caller's points instead of the stroke's corners.
//
// It also exports the reference's ApplyPixel (cpp:515-549), which the header declares (h:109) but the source defines `inline`, so
// the plain reference build has no such symbol: the forwarding function below is bound to the symbol name "ApplyPixel".
#include <vector>

#include "libNativeCPURenderer.cpp"

extern "C" bool ncr_shim_apply_pixel(RenderContext* ctx, long x, long y, double r, double g, double b, double a) __asm__("ApplyPixel");
extern "C" bool ncr_shim_apply_pixel(RenderContext* ctx, long x, long y, double r, double g, double b, double a) {
    return ApplyPixel(ctx, (i64)x, (i64)y, r, g, b, a);
}

extern "C" void NcrFillPolygon(RenderContext* ctx, const double* xy, long n_points, double r, double g, double b, double a) {
    if (!ctx || !xy || n_points <= 0) return;
    std::vector<f64> flat(xy, xy + 2 * n_points);
    f64(*points)[2] = reinterpret_cast<f64(*)[2]>(flat.data());
    f64 inv[6];
    GetInverseTransform(ctx, inv);
    for (i64 i = 0; i < ctx->width; ++i) {   // DrawLine's loop, cpp:906-916
        for (i64 j = 0; j < ctx->height; ++j) {
            f64 invX, invY;
            TransformPointFromMatrix(inv, i, j, &invX, &invY);
            if (!pointInPolygon(invX, invY, points, n_points)) continue;
            ApplyPixel(ctx, i, j, r, g, b, a);
        }
    }
}

// Perspective quads (extension X4).  The reference has nothing projective, so the inverse homography is this repo's own spec
// (include/ncr_b200.h: rw = 1 / (h6*i + h7*j + h8), X = (h0*i + h1*j + h2) * rw, Y likewise, pixels with hw <= 0 skipped); everything
// AFTER the map is DrawTexture's own mapped loop (cpp:761-777): the four inclusive bounds, the scale to texels and the reference's
// InterpolateColorFromBuffer and ApplyPixel.
extern "C" void NcrDrawTexturePerspective(RenderContext* ctx, Texture* tex, const double* h, double x, double y, double width,
                                          double height) {
    if (!ctx || !tex || !h) return;
    if (width == 0 || height == 0) return;   // cpp:726
    f64 scaleX = tex->width / width;         // cpp:728-729
    f64 scaleY = tex->height / height;
    for (i64 i = 0; i < ctx->width; ++i) {
        for (i64 j = 0; j < ctx->height; ++j) {
            f64 fi = (f64)i, fj = (f64)j;
            f64 hw = h[6] * fi + h[7] * fj + h[8];
            f64 rw = 1.0 / hw;
            f64 invX = (h[0] * fi + h[1] * fj + h[2]) * rw;
            f64 invY = (h[3] * fi + h[4] * fj + h[5]) * rw;
            if (!(hw > 0.0)) continue;

            if (invX < x) continue;   // cpp:765-768
            if (invX > x + width) continue;
            if (invY < y) continue;
            if (invY > y + height) continue;

            f64 u = (invX - x) * scaleX;   // cpp:770-771
            f64 v = (invY - y) * scaleY;

            f64 r, g, b, a;
            InterpolateColorFromBuffer(tex->buffer, tex->width, tex->height, tex->enableAlpha, u, v, &r, &g, &b, &a);
            ApplyPixel(ctx, i, j, r, g, b, a);
        }
    }
}
