// Milthm hit-effect alpha (reference cpp:1318-1411): three octaves of value noise in polar coordinates,
// thresholded by t.  Host code on purpose: it depends on libm sin/atan2/floor/sqrt and a threshold, so only the
// host libm reproduces the reference bit for bit.
#pragma once
#include <math.h>

static inline double ncr_he_frac(double v) { return v - floor(v); }

// hash of a lattice point: frac(sin(dot(n, (12.9898, 78.233))) * 43758.5453), cpp:1339-1341
static inline double ncr_he_hash(double nx, double ny) {
    return ncr_he_frac(sin(nx * 12.9898 + ny * 78.233) * 43758.5453);
}

static inline double ncr_he_lerp(double a, double b, double t) { return a + (b - a) * t; }

// bilinear value noise with smoothstep weights, cpp:1372-1383
static inline double ncr_he_noise(double px, double py) {
    const double ix = floor(px), iy = floor(py);
    const double ux = ncr_he_frac(px), uy = ncr_he_frac(py);
    const double h00 = ncr_he_hash(ix, iy);
    const double h10 = ncr_he_hash(ix + 1.0, iy + 0.0);
    const double h01 = ncr_he_hash(ix + 0.0, iy + 1.0);
    const double h11 = ncr_he_hash(ix + 1.0, iy + 1.0);
    const double wx = ux * ux * (3.0 - 2.0 * ux);
    const double wy = uy * uy * (3.0 - 2.0 * uy);
    return ncr_he_lerp(ncr_he_lerp(h00, h10, wx), ncr_he_lerp(h01, h11, wx), wy);
}

// cpp:1385-1411 with density 50
static inline double ncr_hit_effect_alpha(double seed, double t, double x, double y) {
    const double cx = x - 0.5, cy = y - 0.5;
    const double radius = sqrt(cx * cx + cy * cy) * 50.0;
    double angle = fabs(atan2(cy, cx));
    if (y > 0.5) angle += sin(angle) * 2.0;
    const double off = seed * 100.0;
    const double qx = radius + off, qy = angle + off;
    double n = 0.0;
    n += ncr_he_noise(qx, qy) * 0.7;
    n += ncr_he_noise(qx * 2.0, qy * 2.0) * 0.3;
    n += ncr_he_noise(qx * 4.0, qy * 4.0) * 0.1;
    return (n < t) ? 0.0 : 1.0;
}
