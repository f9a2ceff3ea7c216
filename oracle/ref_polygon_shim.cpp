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
