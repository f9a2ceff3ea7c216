// ncr_composite — the per-tile raster/composite kernel (sm_100a).
//
// Work unit: one warp owns a 16x8 half of a 16x16 tile; each lane owns four pixels of it (an 8x4 block
// layout: lane = (lx 0..7, ly 0..3); pixel p = block (p&1, p>>1)), held in registers as f64 RGBA while the
// tile's command list is walked in submission order.  Nothing is shared between warps except a replicated
// u8 -> k/255.0 decode table, so there are no block barriers in the command loop; warps pull half-tiles from
// a global counter (persistent CTAs, one launch per flush).
//
// Per command the warp reads the parameters once through the read-only path at a warp-uniform address (one
// L1 wavefront each) and amortises them over its 128 pixels; 8x4 blocks whose pixels all fall outside the
// command's pixel box are skipped with one vote.  Texel fetches of the four pixels are issued back to back
// before any is consumed.
//
// Arithmetic: the reference's f64 expression trees (reference src/libNativeCPURenderer.cpp, cited inline),
// round-to-nearest intrinsics only, no FMA contraction.  Sub-expressions that do not depend on the pixel
// (inv0*x for a pixel column, the colour-transformed constant colour, 1-a, ...) are hoisted: they are the same
// IEEE operations on the same operands, evaluated once.
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.h"
#include "ncr_cmd.h"
#include "pixel_math.cuh"

#define FULL 0xffffffffu
#define NCR_LUT_COPIES 16
#define NCR_COMPOSITE_THREADS 128

namespace {

// InterpolateColorFromBuffer's clamp, reference cpp:560-563, then truncation (cpp:566).
__device__ __forceinline__ void clamp_uv(double& u, double& v, int w, int h) {
    if (u < 0.0) u = 0.0;
    if (u >= (double)(w - 1)) u = (double)(w - 2);
    if (v < 0.0) v = 0.0;
    if (v >= (double)(h - 1)) v = (double)(h - 2);
}

// General texel fetch (any format) — the uncommon formats and the bilinear extension go through here.
__device__ __forceinline__ void fetch_any(const void* tex, uint32_t flags, const double* lut, int l16, long long idx, double& r,
                                          double& g, double& b, double& a) {
    if (!(flags & NCR_F_TEX_F64)) {
        if (flags & NCR_F_TEX_ALPHA) {
            const uint32_t t = __ldg((const uint32_t*)tex + idx);
            r = lut[((t & 255u) << 4) | l16];
            g = lut[(((t >> 8) & 255u) << 4) | l16];
            b = lut[(((t >> 16) & 255u) << 4) | l16];
            a = lut[((t >> 24) << 4) | l16];
        } else {
            const unsigned char* q = (const unsigned char*)tex + idx * 3;
            r = lut[((uint32_t)__ldg(q) << 4) | l16];
            g = lut[((uint32_t)__ldg(q + 1) << 4) | l16];
            b = lut[((uint32_t)__ldg(q + 2) << 4) | l16];
            a = NCR_RGB_TEXTURE_ALPHA;
        }
    } else {
        if (flags & NCR_F_TEX_ALPHA) {
            const double2* q = (const double2*)tex + idx * 2;
            const double2 lo = __ldg(q), hi = __ldg(q + 1);
            r = lo.x; g = lo.y; b = hi.x; a = hi.y;
        } else {
            const double* q = (const double*)tex + idx * 3;
            r = __ldg(q); g = __ldg(q + 1); b = __ldg(q + 2);
            a = NCR_RGB_TEXTURE_ALPHA;
        }
    }
}

// Slow-path sampler: nearest for non-RGBA8 textures, or NCR_F_BILINEAR (extension, parity unpinned): the four-tap
// formula the reference keeps commented out at cpp:575-620 — same clamp, weights (1-u)(1-v), u(1-v), (1-u)v, uv
// applied left to right.  Not inlined: it is off the common path and would otherwise be replicated per pixel slot.
__device__ __noinline__ void sample_slow(const void* tex, uint32_t flags, int w, int h, const double* lut, int l16, double u,
                                         double v, double* out) {
    clamp_uv(u, v, w, h);
    long long xi = (long long)u, yi = (long long)v;
    xi = xi < 0 ? 0 : (xi > w - 1 ? w - 1 : xi);   // memory safety only; no-op for defined inputs
    yi = yi < 0 ? 0 : (yi > h - 1 ? h - 1 : yi);
    const long long idx = yi * w + xi;
    if (!(flags & NCR_F_BILINEAR)) {
        fetch_any(tex, flags, lut, l16, idx, out[0], out[1], out[2], out[3]);
        return;
    }
    const long long dx = xi + 1 < w ? 1 : 0, dy = yi + 1 < h ? w : 0;
    double c0[4], c1[4], c2[4], c3[4];
    fetch_any(tex, flags, lut, l16, idx, c0[0], c0[1], c0[2], c0[3]);
    fetch_any(tex, flags, lut, l16, idx + dx, c1[0], c1[1], c1[2], c1[3]);
    fetch_any(tex, flags, lut, l16, idx + dy, c2[0], c2[1], c2[2], c2[3]);
    fetch_any(tex, flags, lut, l16, idx + dy + dx, c3[0], c3[1], c3[2], c3[3]);
    const double fu = SUB(u, (double)xi), fv = SUB(v, (double)yi);
    const double mu = SUB(1.0, fu), mv = SUB(1.0, fv);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        out[k] = ADD(ADD(ADD(MUL(MUL(c0[k], mu), mv), MUL(MUL(c1[k], fu), mv)), MUL(MUL(c2[k], mu), fv)), MUL(MUL(c3[k], fu), fv));
}

// pointInPolygon, reference cpp:822-845 (even-odd rule; the divide is only evaluated on crossing edges).
__device__ __forceinline__ bool point_in_poly(const double* __restrict__ pts, uint32_t n, double x, double y) {
    bool res = false;
    double xj = __ldg(pts + 2 * (n - 1)), yj = __ldg(pts + 2 * (n - 1) + 1);
    for (uint32_t i = 0; i < n; ++i) {
        const double xi = __ldg(pts + 2 * i), yi = __ldg(pts + 2 * i + 1);
        if ((yi > y) != (yj > y)) {
            const double xc = ADD(DIV(MUL(SUB(xj, xi), SUB(y, yi)), SUB(yj, yi)), xi);
            if (x < xc) res = !res;
        }
        xj = xi;
        yj = yi;
    }
    return res;
}

// ApplyPixel's blend for an already colour-transformed source (reference cpp:533-546).
template <bool ALPHA>
__device__ __forceinline__ void blend(double& dr, double& dg, double& db, double& da, double r, double g, double b, double a) {
    if (a != 1.0) {
        const double om = SUB(1.0, a);
        r = ADD(MUL(dr, om), MUL(r, a));
        g = ADD(MUL(dg, om), MUL(g, a));
        b = ADD(MUL(db, om), MUL(b, a));
    }
    dr = r; dg = g; db = b;
    if (ALPHA) da = a;   // source alpha replaces destination alpha (cpp:544)
}

#define FOR4 _Pragma("unroll") for (int p = 0; p < 4; ++p)

template <bool ALPHA, bool COUNT>
__global__ void __launch_bounds__(NCR_COMPOSITE_THREADS, 4) ncr_composite(NcrFlushArgs A) {
    // u8 / 255.0 (CreateTextureUInt8, cpp:350), IEEE division on both sides.  16 copies, copy c of entry k at
    // [k*16 + c]: a lane only ever reads copy (lane & 15), so the 64-bit lookups of a half-warp never collide on a bank.
    __shared__ double s_lut[256 * NCR_LUT_COPIES];
    for (int e = threadIdx.x; e < 256 * NCR_LUT_COPIES; e += NCR_COMPOSITE_THREADS)
        s_lut[e] = DIV((double)(e >> 4), 255.0);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int l16 = lane & 15;
    const int lx = lane & 7, ly = lane >> 3;
    const int n_tiles = A.d.tiles_x * A.d.tiles_y;
    const int n_tasks = n_tiles * 2;
    constexpr int IPP = ALPHA ? 4 : 3;
    const int W = A.d.w, H = A.d.h;
    unsigned long long n_applied = 0;

    for (;;) {
        int task = 0;
        if (lane == 0) task = (int)atomicAdd(&A.cursors[5], 1u);
        task = __shfl_sync(FULL, task, 0);
        if (task >= n_tasks) break;
        const int tile = task >> 1;
        const uint32_t loff = __ldg(&A.fine_off[tile]);
        const uint32_t lcount = __ldg(&A.fine_off[n_tiles + tile]);
        if (lcount == 0 && A.u8_out == nullptr) continue;

        const int x0 = (tile % A.d.tiles_x) * NCR_TILE;
        const int y0 = (tile / A.d.tiles_x) * NCR_TILE + (task & 1) * 8;
        if (y0 >= H) continue;
        const int xs[2] = {x0 + lx, x0 + 8 + lx};
        const int ys[2] = {y0 + ly, y0 + 4 + ly};
        const double fx[2] = {(double)xs[0], (double)xs[1]};
        const double fy[2] = {(double)ys[0], (double)ys[1]};
        bool valid[4];
        FOR4 valid[p] = xs[p & 1] < W && ys[p >> 1] < H;

        double dr[4], dg[4], db[4], da[4];
        FOR4 { dr[p] = 0.0; dg[p] = 0.0; db[p] = 0.0; da[p] = 0.0; }
        // The canvas is read unless the list starts with a SetColor (then every pixel is overwritten first).
        if (A.load_fb != 0 || lcount == 0) {
            FOR4 if (valid[p]) {
                const double* q = A.fb + ((size_t)ys[p >> 1] * W + xs[p & 1]) * IPP;
                if (ALPHA) {
                    const double2 lo = ((const double2*)q)[0], hi = ((const double2*)q)[1];
                    dr[p] = lo.x; dg[p] = lo.y; db[p] = hi.x; da[p] = hi.y;
                } else {
                    dr[p] = q[0]; dg[p] = q[1]; db[p] = q[2];
                }
            }
        }

        for (uint32_t k0 = 0; k0 < lcount; k0 += 32) {
            const uint32_t mine = (k0 + lane < lcount) ? __ldg(&A.fine_list[loff + k0 + lane]) : 0u;
            const uint32_t nk = min(32u, lcount - k0);
            for (uint32_t kk = 0; kk < nk; ++kk) {
                const uint32_t ci = __shfl_sync(FULL, mine, kk);
                const int4 box = __ldg((const int4*)&A.boxes[ci]);   // l, r, t, b
                if (box.w <= y0 || box.z >= y0 + 8) continue;         // misses this half of the tile
                const NcrCmd* __restrict__ c = A.cmds + ci;
                const uint2 head = __ldg((const uint2*)c);            // op, flags
                const uint32_t op = head.x, flags = head.y;

                // pixel-box membership (the reference's loop bounds) and 8x4-block culling
                bool in[4];
                FOR4 {
                    const int px = xs[p & 1], py = ys[p >> 1];
                    in[p] = valid[p] && px >= box.x && px < box.y && py >= box.z && py < box.w;
                }

                if (op == NCR_OP_SET_COLOR) {   // cpp:643-657
                    const double c0 = __ldg(&c->p[0]), c1 = __ldg(&c->p[1]), c2 = __ldg(&c->p[2]), c3 = __ldg(&c->p[3]);
                    FOR4 if (in[p]) {
                        dr[p] = c0; dg[p] = c1; db[p] = c2;
                        if (ALPHA) da[p] = c3;
                        // 3-channel canvas, non-uniform colour: SetPixel's index+3 store (cpp:510) leaves `a` in the red
                        // of pixel (0, j>=1), because column 0 is written first and the last column spills into it.
                        else if ((flags & NCR_F_RGB_SPILL) && xs[p & 1] == 0 && ys[p >> 1] >= 1 && W > 1) dr[p] = c3;
                    }
                    continue;
                }
                if (op == NCR_OP_SET_PIXEL) {   // cpp:494-513
                    const double c0 = __ldg(&c->p[0]), c1 = __ldg(&c->p[1]), c2 = __ldg(&c->p[2]), c3 = __ldg(&c->p[3]);
                    FOR4 if (in[p]) {
                        dr[p] = c0;
                        if (!(flags & NCR_F_ONLY_RED)) {
                            dg[p] = c1; db[p] = c2;
                            if (ALPHA) da[p] = c3;
                        }
                    }
                    continue;
                }

                if (op == NCR_OP_GRAD) {   // cpp:1299-1314
                    const double i0 = __ldg(&c->inv[0]), i1 = __ldg(&c->inv[1]), i2 = __ldg(&c->inv[2]), i3 = __ldg(&c->inv[3]);
                    const double i4 = __ldg(&c->inv[4]), i5 = __ldg(&c->inv[5]);
                    const double cx = __ldg(&c->x), cy = __ldg(&c->y), cxw = __ldg(&c->xw), cyh = __ldg(&c->yh);
                    const double hgt = __ldg(&c->sy);
                    const double ax[2] = {MUL(i0, fx[0]), MUL(i0, fx[1])}, bx[2] = {MUL(i1, fx[0]), MUL(i1, fx[1])};
                    const double ay[2] = {MUL(i2, fy[0]), MUL(i2, fy[1])}, by[2] = {MUL(i3, fy[0]), MUL(i3, fy[1])};
                    FOR4 {
                        if (!__any_sync(FULL, in[p])) continue;
                        const double X = ADD(ADD(ax[p & 1], ay[p >> 1]), i4);   // cpp:451-452
                        const double Y = ADD(ADD(bx[p & 1], by[p >> 1]), i5);
                        if (in[p] && !(X < cx) && !(X > cxw) && !(Y < cy) && !(Y > cyh)) {
                            const double t = DIV(SUB(Y, cy), hgt);   // cpp:1308; p[4..7] = bottom - top
                            const double r = MUL(ADD(__ldg(&c->p[0]), MUL(__ldg(&c->p[4]), t)), __ldg(&c->ct[0]));
                            const double g = MUL(ADD(__ldg(&c->p[1]), MUL(__ldg(&c->p[5]), t)), __ldg(&c->ct[1]));
                            const double b = MUL(ADD(__ldg(&c->p[2]), MUL(__ldg(&c->p[6]), t)), __ldg(&c->ct[2]));
                            const double a = MUL(ADD(__ldg(&c->p[3]), MUL(__ldg(&c->p[7]), t)), __ldg(&c->ct[3]));
                            blend<ALPHA>(dr[p], dg[p], db[p], da[p], r, g, b, a);
                            if (COUNT) ++n_applied;
                        }
                    }
                    continue;
                }

                if (op == NCR_OP_RECT || op == NCR_OP_CIRCLE || op == NCR_OP_POLY || op == NCR_OP_FILL_COLOR ||
                    op == NCR_OP_APPLY_PIXEL) {
                    // constant colour: p[0..3] = colour * ct, p[4..6] = rgb*a, p[7] = 1-a, all formed on the host with the
                    // same IEEE operations ApplyPixel performs per pixel (cpp:525-536).
                    if (op != NCR_OP_FILL_COLOR && op != NCR_OP_APPLY_PIXEL) {
                        const double i0 = __ldg(&c->inv[0]), i1 = __ldg(&c->inv[1]), i2 = __ldg(&c->inv[2]), i3 = __ldg(&c->inv[3]);
                        const double i4 = __ldg(&c->inv[4]), i5 = __ldg(&c->inv[5]);
                        const double cx = __ldg(&c->x), cy = __ldg(&c->y);
                        const double ax[2] = {MUL(i0, fx[0]), MUL(i0, fx[1])}, bx[2] = {MUL(i1, fx[0]), MUL(i1, fx[1])};
                        const double ay[2] = {MUL(i2, fy[0]), MUL(i2, fy[1])}, by[2] = {MUL(i3, fy[0]), MUL(i3, fy[1])};
                        if (op == NCR_OP_RECT) {   // cpp:866-869
                            const double cxw = __ldg(&c->xw), cyh = __ldg(&c->yh);
                            FOR4 {
                                if (!__any_sync(FULL, in[p])) continue;
                                const double X = ADD(ADD(ax[p & 1], ay[p >> 1]), i4);
                                const double Y = ADD(ADD(bx[p & 1], by[p >> 1]), i5);
                                in[p] = in[p] && !(X < cx) && !(X > cxw) && !(Y < cy) && !(Y > cyh);
                            }
                        } else if (op == NCR_OP_CIRCLE) {   // cpp:939-943
                            const double rad = __ldg(&c->sx);
                            FOR4 {
                                if (!__any_sync(FULL, in[p])) continue;
                                const double X = ADD(ADD(ax[p & 1], ay[p >> 1]), i4);
                                const double Y = ADD(ADD(bx[p & 1], by[p >> 1]), i5);
                                const double ddx = SUB(X, cx), ddy = SUB(Y, cy);
                                const double dist = __dsqrt_rn(ADD(MUL(ddx, ddx), MUL(ddy, ddy)));
                                in[p] = in[p] && !(dist > rad);
                            }
                        } else {   // NCR_OP_POLY, cpp:913
                            const double* pts = A.aux + __ldg(&c->aux_off);
                            const uint32_t npts = __ldg(&c->aux_n);
                            FOR4 {
                                if (!__any_sync(FULL, in[p])) continue;
                                const double X = ADD(ADD(ax[p & 1], ay[p >> 1]), i4);
                                const double Y = ADD(ADD(bx[p & 1], by[p >> 1]), i5);
                                in[p] = in[p] && point_in_poly(pts, npts, X, Y);
                            }
                        }
                    }
                    const double sa = __ldg(&c->p[3]);
                    if (sa != 1.0) {
                        const double q0 = __ldg(&c->p[4]), q1 = __ldg(&c->p[5]), q2 = __ldg(&c->p[6]), om = __ldg(&c->p[7]);
                        FOR4 if (in[p]) {
                            dr[p] = ADD(MUL(dr[p], om), q0);
                            dg[p] = ADD(MUL(dg[p], om), q1);
                            db[p] = ADD(MUL(db[p], om), q2);
                            if (ALPHA) da[p] = sa;
                            if (COUNT) ++n_applied;
                        }
                    } else {
                        const double c0 = __ldg(&c->p[0]), c1 = __ldg(&c->p[1]), c2 = __ldg(&c->p[2]);
                        FOR4 if (in[p]) {
                            dr[p] = c0; dg[p] = c1; db[p] = c2;
                            if (ALPHA) da[p] = sa;
                            if (COUNT) ++n_applied;
                        }
                    }
                    continue;
                }

                // ---- textured ops: DrawTexture (both paths), DrawSplittedTexture, perspective extension ----
                const int2 twh = __ldg((const int2*)&c->tex_w);
                const int tw = twh.x, th = twh.y;
                const void* tex = (const void*)__ldg((const unsigned long long*)&c->tex);
                const double cx = __ldg(&c->x), cy = __ldg(&c->y), cxw = __ldg(&c->xw), cyh = __ldg(&c->yh);
                const double sx = __ldg(&c->sx), sy = __ldg(&c->sy);
                double u[4], v[4];
                if (op == NCR_OP_TEX_IDENT) {   // cpp:741-745: pixels i >= (i64)x with (f64)i < x + width; u = (i - x) * scaleX
                    const double i0 = __ldg(&c->p[0]), j0 = __ldg(&c->p[1]);
                    FOR4 {
                        const double fi = fx[p & 1], fj = fy[p >> 1];
                        in[p] = in[p] && fi >= i0 && fi < cxw && fj >= j0 && fj < cyh;
                        u[p] = MUL(SUB(fi, cx), sx);
                        v[p] = MUL(SUB(fj, cy), sy);
                    }
                } else if (op == NCR_OP_TEX_PERSP) {   // extension: row-major 3x3 inverse homography, then cpp:765-771
                    const double h0 = __ldg(&c->inv[0]), h1 = __ldg(&c->inv[1]), h2 = __ldg(&c->inv[2]);
                    const double h3 = __ldg(&c->inv[3]), h4 = __ldg(&c->inv[4]), h5 = __ldg(&c->inv[5]);
                    const double h6 = __ldg(&c->p[0]), h7 = __ldg(&c->p[1]), h8 = __ldg(&c->p[2]);
                    FOR4 {
                        const double fi = fx[p & 1], fj = fy[p >> 1];
                        const double hw = ADD(ADD(MUL(h6, fi), MUL(h7, fj)), h8);
                        const double X = DIV(ADD(ADD(MUL(h0, fi), MUL(h1, fj)), h2), hw);
                        const double Y = DIV(ADD(ADD(MUL(h3, fi), MUL(h4, fj)), h5), hw);
                        in[p] = in[p] && hw > 0.0 && !(X < cx) && !(X > cxw) && !(Y < cy) && !(Y > cyh);
                        u[p] = MUL(SUB(X, cx), sx);
                        v[p] = MUL(SUB(Y, cy), sy);
                    }
                } else {   // NCR_OP_TEX / NCR_OP_TEX_SPLIT, cpp:763-771
                    const double i0 = __ldg(&c->inv[0]), i1 = __ldg(&c->inv[1]), i2 = __ldg(&c->inv[2]), i3 = __ldg(&c->inv[3]);
                    const double i4 = __ldg(&c->inv[4]), i5 = __ldg(&c->inv[5]);
                    const double ax[2] = {MUL(i0, fx[0]), MUL(i0, fx[1])}, bx[2] = {MUL(i1, fx[0]), MUL(i1, fx[1])};
                    const double ay[2] = {MUL(i2, fy[0]), MUL(i2, fy[1])}, by[2] = {MUL(i3, fy[0]), MUL(i3, fy[1])};
                    FOR4 {
                        u[p] = 0.0; v[p] = 0.0;
                        if (!__any_sync(FULL, in[p])) continue;
                        const double X = ADD(ADD(ax[p & 1], ay[p >> 1]), i4);
                        const double Y = ADD(ADD(bx[p & 1], by[p >> 1]), i5);
                        in[p] = in[p] && !(X < cx) && !(X > cxw) && !(Y < cy) && !(Y > cyh);
                        u[p] = MUL(SUB(X, cx), sx);
                        v[p] = MUL(SUB(Y, cy), sy);
                    }
                    if (op == NCR_OP_TEX_SPLIT) {
                        // cpp:812-813: u = (uStart + (uEnd - uStart) * u / tex->width) * tex->width
                        const double uS = __ldg(&c->p[0]), dU = __ldg(&c->p[1]), vS = __ldg(&c->p[2]), dV = __ldg(&c->p[3]);
                        const double fw = __ldg(&c->p[4]), fh = __ldg(&c->p[5]);
                        if (flags & NCR_F_SPLIT_POW2) {
                            // width and height are powers of two: x / 2^k and x * 2^-k are the same correctly rounded value
                            const double rw = __ldg(&c->p[6]), rh = __ldg(&c->p[7]);
                            FOR4 {
                                u[p] = MUL(ADD(uS, MUL(MUL(dU, u[p]), rw)), fw);
                                v[p] = MUL(ADD(vS, MUL(MUL(dV, v[p]), rh)), fh);
                            }
                        } else {
                            FOR4 if (__any_sync(FULL, in[p])) {
                                u[p] = MUL(ADD(uS, DIV(MUL(dU, u[p]), fw)), fw);
                                v[p] = MUL(ADD(vS, DIV(MUL(dV, v[p]), fh)), fh);
                            }
                        }
                    }
                }

                const double ct3 = __ldg(&c->ct[3]);
                const bool rgb_one = (flags & NCR_F_CT_RGB_ONE) != 0;   // r * 1.0 == r exactly: the multiply is skipped
                double ct0 = 1.0, ct1 = 1.0, ct2 = 1.0;
                if (!rgb_one) { ct0 = __ldg(&c->ct[0]); ct1 = __ldg(&c->ct[1]); ct2 = __ldg(&c->ct[2]); }

                if (flags & NCR_F_TEX_FAST) {   // RGBA8 texels, nearest, < 2^31 texels
                    const uint32_t* t32 = (const uint32_t*)tex;
                    uint32_t tx[4];
                    FOR4 {
                        tx[p] = 0u;
                        if (!__any_sync(FULL, in[p])) continue;
                        clamp_uv(u[p], v[p], tw, th);                       // cpp:560-563
                        int xi = __double2int_rz(u[p]), yi = __double2int_rz(v[p]);   // cpp:566 (i64) truncation
                        xi = max(0, min(tw - 1, xi));                       // memory safety only
                        yi = max(0, min(th - 1, yi));
                        tx[p] = in[p] ? __ldg(t32 + (yi * tw + xi)) : 0u;
                    }
                    FOR4 if (in[p]) {
                        double r = s_lut[((tx[p] & 255u) << 4) | l16];
                        double g = s_lut[(((tx[p] >> 8) & 255u) << 4) | l16];
                        double b = s_lut[(((tx[p] >> 16) & 255u) << 4) | l16];
                        double a = s_lut[((tx[p] >> 24) << 4) | l16];
                        if (!rgb_one) { r = MUL(r, ct0); g = MUL(g, ct1); b = MUL(b, ct2); }   // cpp:525-527
                        a = MUL(a, ct3);                                                       // cpp:528
                        blend<ALPHA>(dr[p], dg[p], db[p], da[p], r, g, b, a);
                        if (COUNT) ++n_applied;
                    }
                } else {
                    FOR4 if (in[p]) {
                        double s[4];
                        sample_slow(tex, flags, tw, th, s_lut, l16, u[p], v[p], s);
                        blend<ALPHA>(dr[p], dg[p], db[p], da[p], MUL(s[0], ct0), MUL(s[1], ct1), MUL(s[2], ct2), MUL(s[3], ct3));
                        if (COUNT) ++n_applied;
                    }
                }
            }
        }

        // tile write-back: canonical f64 canvas (only if something was drawn) and the fused (iu8)(v*255) image
        FOR4 if (valid[p]) {
            const size_t pix = ((size_t)ys[p >> 1] * W + xs[p & 1]) * IPP;
            if (lcount != 0) {
                double* q = A.fb + pix;
                if (ALPHA) {
                    ((double2*)q)[0] = make_double2(dr[p], dg[p]);
                    ((double2*)q)[1] = make_double2(db[p], da[p]);
                } else {
                    q[0] = dr[p]; q[1] = dg[p]; q[2] = db[p];
                }
            }
            if (A.u8_out) {
                if (ALPHA) {
                    const uint32_t o = (uint32_t)ncr_to_u8(dr[p]) | ((uint32_t)ncr_to_u8(dg[p]) << 8) |
                                       ((uint32_t)ncr_to_u8(db[p]) << 16) | ((uint32_t)ncr_to_u8(da[p]) << 24);
                    ((uint32_t*)A.u8_out)[(size_t)ys[p >> 1] * W + xs[p & 1]] = o;
                } else {
                    unsigned char* o = A.u8_out + pix;
                    o[0] = ncr_to_u8(dr[p]); o[1] = ncr_to_u8(dg[p]); o[2] = ncr_to_u8(db[p]);
                }
            }
        }
    }

    if (COUNT) {
        for (int s = 16; s > 0; s >>= 1) n_applied += __shfl_down_sync(FULL, n_applied, s);
        if (lane == 0 && n_applied) atomicAdd((unsigned long long*)(A.cursors + 2), n_applied);
    }
}

int g_grid[4] = {0, 0, 0, 0};

template <bool ALPHA, bool COUNT>
void launch(const NcrFlushArgs& A, cudaStream_t s, int slot) {
    if (g_grid[slot] == 0) {
        int dev = 0, sms = 148, per_sm = 4;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ncr_composite<ALPHA, COUNT>, NCR_COMPOSITE_THREADS, 0);
        g_grid[slot] = sms * (per_sm > 0 ? per_sm : 1);   // persistent: every resident CTA slot, one wave
    }
    const int n_tasks = A.d.tiles_x * A.d.tiles_y * 2;
    const int warps = NCR_COMPOSITE_THREADS / 32;
    int grid = g_grid[slot];
    if (grid * warps > n_tasks) grid = (n_tasks + warps - 1) / warps;
    if (grid < 1) grid = 1;
    ncr_composite<ALPHA, COUNT><<<grid, NCR_COMPOSITE_THREADS, 0, s>>>(A);
}

}   // namespace

extern "C" void ncr_launch_composite(const NcrFlushArgs* A, cudaStream_t s) {
    const bool alpha = A->d.ipp == 4, count = A->count_pixels != 0;
    if (alpha) {
        if (count) launch<true, true>(*A, s, 0);
        else launch<true, false>(*A, s, 1);
    } else {
        if (count) launch<false, true>(*A, s, 2);
        else launch<false, false>(*A, s, 3);
    }
}
