// sm_100a kernels of the draw/composite path.
//
//   ncr_bin_coarse   commands -> ordered per-128x128-px bin lists      (CTA per bin)
//   ncr_bin_fine     bin lists -> ordered per-16x16-px tile lists      (warp per tile, ballot + popc compaction)
//   ncr_composite    CTA per tile, one thread per pixel; the pixel lives in registers while the tile's
//                    commands are applied in submission order; the tile is read from / written to HBM once.
//
// Arithmetic contract (DESIGN.md "Exactness"): every per-pixel expression is the reference's f64
// expression tree (reference src/libNativeCPURenderer.cpp, cited per function) evaluated with
// round-to-nearest mul/add/sub/div/sqrt intrinsics, which the compiler never contracts into FMAs.
// The translation unit is additionally compiled with -fmad=false.
#include <cuda_runtime.h>
#include <stdint.h>
#include "ncr_cmd.h"
#include "kernels.h"

#define MUL(a, b) __dmul_rn((a), (b))
#define ADD(a, b) __dadd_rn((a), (b))
#define SUB(a, b) __dsub_rn((a), (b))
#define DIV(a, b) __ddiv_rn((a), (b))

// ------------------------------------------------------------------------------------------------
// binning
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool box_hits(const NcrBox& bx, int x0, int y0, int x1, int y1) {
    return bx.l < bx.r && bx.t < bx.b && bx.l < x1 && bx.r > x0 && bx.t < y1 && bx.b > y0;
}

__global__ void __launch_bounds__(256) ncr_bin_coarse(NcrFlushArgs A) {
    const int bin = blockIdx.x;
    const int bx = bin % A.d.bins_x, by = bin / A.d.bins_x;
    const int edge = NCR_TILE * NCR_COARSE;
    const int x0 = bx * edge, y0 = by * edge, x1 = x0 + edge, y1 = y0 + edge;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base;

    uint32_t count = 0;
    for (uint32_t base = 0; base < A.n_cmds; base += 256) {
        uint32_t idx = base + tid;
        bool hit = idx < A.n_cmds && box_hits(A.boxes[idx], x0, y0, x1, y1);
        count += __syncthreads_count(hit);
    }
    if (tid == 0) {
        uint32_t off = atomicAdd(&A.cursors[0], count);
        if (off + count > A.coarse_cap) { count = 0; atomicExch(&A.cursors[4], 1u); }
        s_base = off;
        A.coarse_off[bin] = off;
        A.coarse_off[A.d.bins_x * A.d.bins_y + bin] = count;
    }
    __syncthreads();
    if (A.coarse_off[A.d.bins_x * A.d.bins_y + bin] == 0) return;
    uint32_t running = s_base;
    for (uint32_t base = 0; base < A.n_cmds; base += 256) {
        uint32_t idx = base + tid;
        bool hit = idx < A.n_cmds && box_hits(A.boxes[idx], x0, y0, x1, y1);
        uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            uint32_t c = s_warp[w];
            before += (w < warp) ? c : 0;
            total += c;
        }
        if (hit) A.coarse_list[running + before + __popc(m & ((1u << lane) - 1))] = idx;
        running += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) ncr_bin_fine(NcrFlushArgs A) {
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int n_tiles = A.d.tiles_x * A.d.tiles_y;
    if (tile >= n_tiles) return;
    const int tx = tile % A.d.tiles_x, ty = tile / A.d.tiles_x;
    const int x0 = tx * NCR_TILE, y0 = ty * NCR_TILE, x1 = x0 + NCR_TILE, y1 = y0 + NCR_TILE;
    const int bin = (ty / NCR_COARSE) * A.d.bins_x + tx / NCR_COARSE;
    const uint32_t cbase = A.coarse_off[bin];
    const uint32_t ccount = A.coarse_off[A.d.bins_x * A.d.bins_y + bin];

    uint32_t count = 0;
    for (uint32_t k = 0; k < ccount; k += 32) {
        bool hit = false;
        if (k + lane < ccount) hit = box_hits(A.boxes[A.coarse_list[cbase + k + lane]], x0, y0, x1, y1);
        count += __popc(__ballot_sync(0xffffffffu, hit));
    }
    uint32_t off = 0;
    if (lane == 0) {
        off = atomicAdd(&A.cursors[1], count);
        if (off + count > A.fine_cap) { count = 0; atomicExch(&A.cursors[4], 2u); }
        A.fine_off[tile] = off;
        A.fine_off[n_tiles + tile] = count;
    }
    off = __shfl_sync(0xffffffffu, off, 0);
    count = __shfl_sync(0xffffffffu, count, 0);
    if (count == 0) return;
    for (uint32_t k = 0; k < ccount; k += 32) {
        bool hit = false;
        uint32_t idx = 0;
        if (k + lane < ccount) {
            idx = A.coarse_list[cbase + k + lane];
            hit = box_hits(A.boxes[idx], x0, y0, x1, y1);
        }
        uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (hit) A.fine_list[off + __popc(m & ((1u << lane) - 1))] = idx;
        off += __popc(m);
    }
}

// ------------------------------------------------------------------------------------------------
// per-pixel pieces
// ------------------------------------------------------------------------------------------------
struct Px {
    double r, g, b, a;
};

// ApplyPixel, reference cpp:515-549 (bounds are guaranteed by the caller: the thread owns an in-canvas pixel).
template <bool ALPHA>
__device__ __forceinline__ void apply_px(Px& P, double r, double g, double b, double a, const double* ct) {
    r = MUL(r, ct[0]);
    g = MUL(g, ct[1]);
    b = MUL(b, ct[2]);
    a = MUL(a, ct[3]);
    if (a != 1.0) {
        const double om = SUB(1.0, a);
        r = ADD(MUL(P.r, om), MUL(r, a));
        g = ADD(MUL(P.g, om), MUL(g, a));
        b = ADD(MUL(P.b, om), MUL(b, a));
    }
    P.r = r;
    P.g = g;
    P.b = b;
    if (ALPHA) P.a = a;   // source alpha replaces destination alpha (cpp:544; cpp:545 is a dead store)
}

// One texel -> four f64 channels.  u8 texels decode through the k/255.0 table (cpp:350).
__device__ __forceinline__ void fetch_texel(const NcrCmd& c, const double* lut, long long idx, double& r, double& g,
                                            double& b, double& a) {
    if (!(c.flags & NCR_F_TEX_F64)) {
        if (c.flags & NCR_F_TEX_ALPHA) {
            const uint32_t t = __ldg((const uint32_t*)c.tex + idx);
            r = lut[t & 255u];
            g = lut[(t >> 8) & 255u];
            b = lut[(t >> 16) & 255u];
            a = lut[t >> 24];
        } else {
            const unsigned char* q = (const unsigned char*)c.tex + idx * 3;
            r = lut[__ldg(q)];
            g = lut[__ldg(q + 1)];
            b = lut[__ldg(q + 2)];
            a = NCR_RGB_TEXTURE_ALPHA;
        }
    } else {
        if (c.flags & NCR_F_TEX_ALPHA) {
            const double2* q = (const double2*)c.tex + idx * 2;
            const double2 lo = __ldg(q), hi = __ldg(q + 1);
            r = lo.x; g = lo.y; b = hi.x; a = hi.y;
        } else {
            const double* q = (const double*)c.tex + idx * 3;
            r = __ldg(q); g = __ldg(q + 1); b = __ldg(q + 2);
            a = NCR_RGB_TEXTURE_ALPHA;
        }
    }
}

// InterpolateColorFromBuffer, reference cpp:555-573: nearest texel, clamp to [0, w-2] / [0, h-2], truncate.
// NCR_F_BILINEAR (extension, parity unpinned): the four-tap formula the reference keeps commented out at
// cpp:575-620, same clamp, weights (1-u)(1-v), u(1-v), (1-u)v, uv applied left to right.
__device__ __forceinline__ void sample_texture(const NcrCmd& c, const double* lut, double u, double v, double& r,
                                               double& g, double& b, double& a) {
    const int w = c.tex_w, h = c.tex_h;
    if (u < 0.0) u = 0.0;
    if (u >= (double)(w - 1)) u = (double)(w - 2);
    if (v < 0.0) v = 0.0;
    if (v >= (double)(h - 1)) v = (double)(h - 2);
    long long xi = (long long)u, yi = (long long)v;   // cvt.rzi.s64.f64 == C truncation
    // Memory-safety clamp; a no-op for every input the reference defines (w,h >= 2, finite u,v).
    xi = xi < 0 ? 0 : (xi > w - 1 ? w - 1 : xi);
    yi = yi < 0 ? 0 : (yi > h - 1 ? h - 1 : yi);
    const long long idx = yi * w + xi;
    if (!(c.flags & NCR_F_BILINEAR)) {
        fetch_texel(c, lut, idx, r, g, b, a);
        return;
    }
    const long long dx = xi + 1 < w ? 1 : 0, dy = yi + 1 < h ? w : 0;
    double r0, g0, b0, a0, r1, g1, b1, a1, r2, g2, b2, a2, r3, g3, b3, a3;
    fetch_texel(c, lut, idx, r0, g0, b0, a0);
    fetch_texel(c, lut, idx + dx, r1, g1, b1, a1);
    fetch_texel(c, lut, idx + dy, r2, g2, b2, a2);
    fetch_texel(c, lut, idx + dy + dx, r3, g3, b3, a3);
    const double fu = SUB(u, (double)xi), fv = SUB(v, (double)yi);
    const double mu = SUB(1.0, fu), mv = SUB(1.0, fv);
#define NCR_BILERP(c0, c1, c2, c3) \
    ADD(ADD(ADD(MUL(MUL(c0, mu), mv), MUL(MUL(c1, fu), mv)), MUL(MUL(c2, mu), fv)), MUL(MUL(c3, fu), fv))
    r = NCR_BILERP(r0, r1, r2, r3);
    g = NCR_BILERP(g0, g1, g2, g3);
    b = NCR_BILERP(b0, b1, b2, b3);
    a = NCR_BILERP(a0, a1, a2, a3);
#undef NCR_BILERP
}

// pointInPolygon, reference cpp:822-845 (even-odd rule, divide only on crossing edges).
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

// GetBufferAsUInt8, reference cpp:52-57: (iu8)(v * 255) as x86-64 gcc compiles it — cvttsd2si to a 32-bit
// integer (truncate toward zero; NaN / out-of-range give the "integer indefinite" 0x80000000), low byte kept.
__device__ __forceinline__ unsigned char to_u8(double v) {
    const double s = MUL(v, 255.0);
    int t;
    if (!(fabs(s) < 2147483648.0)) t = (int)0x80000000;
    else t = __double2int_rz(s);
    return (unsigned char)(t & 0xff);
}

// ------------------------------------------------------------------------------------------------
// composite
// ------------------------------------------------------------------------------------------------
#define NCR_CHUNK 8

template <bool ALPHA, bool COUNT>
__global__ void __launch_bounds__(256) ncr_composite(NcrFlushArgs A) {
    __shared__ double s_lut[256];
    __shared__ NcrCmd s_cmd[NCR_CHUNK];

    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int n_tiles = A.d.tiles_x * A.d.tiles_y;
    const uint32_t loff = A.fine_off[tile];
    const uint32_t lcount = A.fine_off[n_tiles + tile];
    if (lcount == 0 && A.u8_out == nullptr) return;

    const int tx = tile % A.d.tiles_x, ty = tile / A.d.tiles_x;
    const int i = tx * NCR_TILE + (tid & (NCR_TILE - 1));
    const int j = ty * NCR_TILE + (tid / NCR_TILE);
    const bool valid = i < A.d.w && j < A.d.h;
    const double fi = (double)i, fj = (double)j;
    constexpr int IPP = ALPHA ? 4 : 3;
    const size_t pix = ((size_t)j * A.d.w + i) * IPP;

    s_lut[tid] = DIV((double)tid, 255.0);   // CreateTextureUInt8's u8 / 255.0 (cpp:350), IEEE-exact on both sides

    Px P = {0.0, 0.0, 0.0, 0.0};
    // The canvas is read unless the tile's first command overwrites every pixel (SetColor).
    const bool need_fb = A.load_fb != 0 || lcount == 0;
    if (need_fb && valid) {
        if (ALPHA) {
            const double2* q = (const double2*)(A.fb + pix);
            const double2 lo = q[0], hi = q[1];
            P.r = lo.x; P.g = lo.y; P.b = hi.x; P.a = hi.y;
        } else {
            P.r = A.fb[pix]; P.g = A.fb[pix + 1]; P.b = A.fb[pix + 2];
        }
    }
    unsigned long long n_applied = 0;

    for (uint32_t k0 = 0; k0 < lcount; k0 += NCR_CHUNK) {
        const uint32_t nc = min((uint32_t)NCR_CHUNK, lcount - k0);
        __syncthreads();
        if (tid < (int)(nc * NCR_CMD_WORDS16)) {
            const uint32_t which = tid / NCR_CMD_WORDS16, word = tid % NCR_CMD_WORDS16;
            const uint32_t ci = A.fine_list[loff + k0 + which];
            ((uint4*)&s_cmd[which])[word] = __ldg((const uint4*)&A.cmds[ci] + word);
        }
        __syncthreads();
        if (!valid) continue;
        for (uint32_t k = 0; k < nc; ++k) {
            const NcrCmd& c = s_cmd[k];
            if (i < c.l || i >= c.r || j < c.t || j >= c.b) continue;
            double r, g, b, a;
            switch (c.op) {
                case NCR_OP_SET_COLOR: {   // cpp:643-657
                    P.r = c.p[0]; P.g = c.p[1]; P.b = c.p[2];
                    if (ALPHA) P.a = c.p[3];
                    else if ((c.flags & NCR_F_RGB_SPILL) && i == 0 && j >= 1 && A.d.w > 1) P.r = c.p[3];
                    continue;
                }
                case NCR_OP_SET_PIXEL: {   // cpp:494-513; p[4] != 0: only the red element (3-channel spill of cpp:510)
                    if (c.p[4] != 0.0) { P.r = c.p[0]; continue; }
                    P.r = c.p[0]; P.g = c.p[1]; P.b = c.p[2];
                    if (ALPHA) P.a = c.p[3];
                    continue;
                }
                case NCR_OP_FILL_COLOR:    // cpp:682-691
                case NCR_OP_APPLY_PIXEL: { // cpp:515-549
                    r = c.p[0]; g = c.p[1]; b = c.p[2]; a = c.p[3];
                    break;
                }
                case NCR_OP_TEX_IDENT: {   // cpp:741-750
                    if (!(fi >= c.p[0] && fi < c.xw && fj >= c.p[1] && fj < c.yh)) continue;
                    const double u = MUL(SUB(fi, c.x), c.sx);
                    const double v = MUL(SUB(fj, c.y), c.sy);
                    sample_texture(c, s_lut, u, v, r, g, b, a);
                    break;
                }
                case NCR_OP_TEX_PERSP: {   // extension: row-major 3x3 inverse homography, then cpp:765-775
                    const double hw = ADD(ADD(MUL(c.p[0], fi), MUL(c.p[1], fj)), c.p[2]);
                    const double X = DIV(ADD(ADD(MUL(c.inv[0], fi), MUL(c.inv[1], fj)), c.inv[2]), hw);
                    const double Y = DIV(ADD(ADD(MUL(c.inv[3], fi), MUL(c.inv[4], fj)), c.inv[5]), hw);
                    if (!(hw > 0.0)) continue;
                    if (X < c.x) continue;
                    if (X > c.xw) continue;
                    if (Y < c.y) continue;
                    if (Y > c.yh) continue;
                    sample_texture(c, s_lut, MUL(SUB(X, c.x), c.sx), MUL(SUB(Y, c.y), c.sy), r, g, b, a);
                    break;
                }
                default: {
                    // TransformPointFromMatrix(inv, i, j), cpp:451-452
                    const double X = ADD(ADD(MUL(c.inv[0], fi), MUL(c.inv[2], fj)), c.inv[4]);
                    const double Y = ADD(ADD(MUL(c.inv[1], fi), MUL(c.inv[3], fj)), c.inv[5]);
                    if (c.op == NCR_OP_CIRCLE) {   // cpp:939-943
                        const double dx = SUB(X, c.x), dy = SUB(Y, c.y);
                        const double dist = __dsqrt_rn(ADD(MUL(dx, dx), MUL(dy, dy)));
                        if (dist > c.sx) continue;
                        r = c.p[0]; g = c.p[1]; b = c.p[2]; a = c.p[3];
                    } else if (c.op == NCR_OP_POLY) {   // cpp:913
                        if (!point_in_poly(A.aux + c.aux_off, c.aux_n, X, Y)) continue;
                        r = c.p[0]; g = c.p[1]; b = c.p[2]; a = c.p[3];
                    } else {
                        // the four inclusive bounds, cpp:765-768 (NaN compares false on both sides, as in C)
                        if (X < c.x) continue;
                        if (X > c.xw) continue;
                        if (Y < c.y) continue;
                        if (Y > c.yh) continue;
                        if (c.op == NCR_OP_RECT) {
                            r = c.p[0]; g = c.p[1]; b = c.p[2]; a = c.p[3];
                        } else if (c.op == NCR_OP_GRAD) {   // cpp:1308-1312; p[4..7] = bottom - top
                            const double t = DIV(SUB(Y, c.y), c.sy);
                            r = ADD(c.p[0], MUL(c.p[4], t));
                            g = ADD(c.p[1], MUL(c.p[5], t));
                            b = ADD(c.p[2], MUL(c.p[6], t));
                            a = ADD(c.p[3], MUL(c.p[7], t));
                        } else {
                            double u = MUL(SUB(X, c.x), c.sx);   // cpp:770-771
                            double v = MUL(SUB(Y, c.y), c.sy);
                            if (c.op == NCR_OP_TEX_SPLIT) {   // cpp:812-813; p = {uS, uE-uS, vS, vE-vS, (f64)w, (f64)h}
                                u = MUL(ADD(c.p[0], DIV(MUL(c.p[1], u), c.p[4])), c.p[4]);
                                v = MUL(ADD(c.p[2], DIV(MUL(c.p[3], v), c.p[5])), c.p[5]);
                            }
                            sample_texture(c, s_lut, u, v, r, g, b, a);
                        }
                    }
                    break;
                }
            }
            apply_px<ALPHA>(P, r, g, b, a, c.ct);
            if (COUNT) ++n_applied;
        }
    }

    if (valid) {
        if (lcount != 0) {
            if (ALPHA) {
                double2* q = (double2*)(A.fb + pix);
                q[0] = make_double2(P.r, P.g);
                q[1] = make_double2(P.b, P.a);
            } else {
                A.fb[pix] = P.r; A.fb[pix + 1] = P.g; A.fb[pix + 2] = P.b;
            }
        }
        if (A.u8_out) {
            if (ALPHA) {
                const uint32_t o = (uint32_t)to_u8(P.r) | ((uint32_t)to_u8(P.g) << 8) | ((uint32_t)to_u8(P.b) << 16) |
                                   ((uint32_t)to_u8(P.a) << 24);
                ((uint32_t*)A.u8_out)[(size_t)j * A.d.w + i] = o;
            } else {
                unsigned char* o = A.u8_out + pix;
                o[0] = to_u8(P.r); o[1] = to_u8(P.g); o[2] = to_u8(P.b);
            }
        }
    }
    if (COUNT) {
        for (int s = 16; s > 0; s >>= 1) n_applied += __shfl_down_sync(0xffffffffu, n_applied, s);
        if ((tid & 31) == 0 && n_applied) atomicAdd((unsigned long long*)(A.cursors + 2), n_applied);
    }
}

// f64 canvas -> (iu8)(v*255) image without drawing (readback of an already-flushed canvas).
__global__ void __launch_bounds__(256) ncr_convert_u8(const double* __restrict__ fb, unsigned char* __restrict__ out,
                                                      size_t n) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) out[k] = to_u8(fb[k]);
}

// ResampleTexture, reference cpp:950-976: out(i,j) = nearest(in, (f64)i / width * in.w, (f64)j / height * in.h).
__global__ void __launch_bounds__(256) ncr_resample(NcrCmd src, void* out, int ow, int oh) {
    const int i = blockIdx.x * 16 + (threadIdx.x & 15);
    const int j = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (i >= ow || j >= oh) return;
    double u = MUL(DIV((double)i, (double)ow), (double)src.tex_w);
    double v = MUL(DIV((double)j, (double)oh), (double)src.tex_h);
    const int w = src.tex_w, h = src.tex_h;
    if (u < 0.0) u = 0.0;
    if (u >= (double)(w - 1)) u = (double)(w - 2);
    if (v < 0.0) v = 0.0;
    if (v >= (double)(h - 1)) v = (double)(h - 2);
    long long xi = (long long)u, yi = (long long)v;
    xi = xi < 0 ? 0 : (xi > w - 1 ? w - 1 : xi);
    yi = yi < 0 ? 0 : (yi > h - 1 ? h - 1 : yi);
    const int ipp = (src.flags & NCR_F_TEX_ALPHA) ? 4 : 3;
    const size_t si = ((size_t)yi * w + xi) * ipp, di = ((size_t)j * ow + i) * ipp;
    if (src.flags & NCR_F_TEX_F64) {
        for (int ch = 0; ch < ipp; ++ch) ((double*)out)[di + ch] = ((const double*)src.tex)[si + ch];
    } else {
        for (int ch = 0; ch < ipp; ++ch) ((unsigned char*)out)[di + ch] = ((const unsigned char*)src.tex)[si + ch];
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
extern "C" void ncr_launch_flush(const NcrFlushArgs* A, cudaStream_t s, cudaEvent_t* ev /* 4 or null */) {
    const int n_tiles = A->d.tiles_x * A->d.tiles_y;
    const int n_bins = A->d.bins_x * A->d.bins_y;
    cudaMemsetAsync(A->cursors, 0, 8 * sizeof(uint32_t), s);
    if (ev) cudaEventRecord(ev[0], s);
    if (A->n_cmds) {
        ncr_bin_coarse<<<n_bins, 256, 0, s>>>(*A);
    } else {
        cudaMemsetAsync(A->coarse_off, 0, 2 * n_bins * sizeof(uint32_t), s);
    }
    if (ev) cudaEventRecord(ev[1], s);
    ncr_bin_fine<<<(n_tiles + 7) / 8, 256, 0, s>>>(*A);
    if (ev) cudaEventRecord(ev[2], s);
    const bool alpha = A->d.ipp == 4;
    if (alpha) {
        if (A->count_pixels) ncr_composite<true, true><<<n_tiles, 256, 0, s>>>(*A);
        else ncr_composite<true, false><<<n_tiles, 256, 0, s>>>(*A);
    } else {
        if (A->count_pixels) ncr_composite<false, true><<<n_tiles, 256, 0, s>>>(*A);
        else ncr_composite<false, false><<<n_tiles, 256, 0, s>>>(*A);
    }
    if (ev) cudaEventRecord(ev[3], s);
}

extern "C" void ncr_launch_convert_u8(const double* fb, unsigned char* out, size_t n, cudaStream_t s) {
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    ncr_convert_u8<<<blocks, 256, 0, s>>>(fb, out, n);
}

extern "C" void ncr_launch_resample(const NcrCmd* src, void* out, int ow, int oh, cudaStream_t s) {
    dim3 grid((ow + 15) / 16, (oh + 15) / 16);
    ncr_resample<<<grid, 256, 0, s>>>(*src, out, ow, oh);
}
