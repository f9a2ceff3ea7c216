// ncr_composite — the per-tile raster/composite kernel (sm_100a).
//
// Work unit: one warp owns a 16x8 half of a 16x16 tile; each lane owns four pixels of it (an 8x4 block
// layout: lane = (lx 0..7, ly 0..3); pixel p = block (p&1, p>>1)), held in registers as f64 RGBA while the
// tile's command list is walked in submission order.  Nothing is shared between warps except a replicated
// u8 -> k/255.0 decode table, so there are no block barriers in the command loop; warps pull half-tiles from
// a global counter (persistent CTAs, one launch per flush).
//
// Per command the warp stages the 240-byte command in a per-warp shared slot (the next one is fetched while the
// current one is applied), reads its parameters as warp-uniform LDS broadcasts and amortises them over its 128
// pixels.  The region's list (written by ncr_bin_fine) holds only commands that can touch the region, each tagged
// INTERIOR when every pixel of the region provably passes its box and coverage tests: those run straight-line code
// with no per-pixel test or select on coverage.  The entry also carries the command's dispatch hints (fast textured
// path, split, ct.rgb == 1, a != 1 proven), so the code path of a hot command is chosen from a register before its staged
// copy is touched.  Texel fetches of the four pixels are issued back to back before any is consumed.  The write-back
// produces, in the same pass, the f64 canvas (unless the flush is present-only) and — when asked — the (iu8)(v*255) image;
// its YUV 4:2:0 planes come from ncr_yuv420p right after (present path).
//
// Arithmetic: the reference's f64 expression trees (reference src/libNativeCPURenderer.cpp, cited inline),
// round-to-nearest intrinsics only, no FMA contraction.  Sub-expressions that do not depend on the pixel
// (inv0*x for a pixel column, the colour-transformed constant colour, 1-a, ...) are hoisted: they are the same
// IEEE operations on the same operands, evaluated once.
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "kernels.h"
#include "ncr_cmd.h"
#include "pixel_math.cuh"

// Instruction-cache footprint matters more than call overhead here (ncu: a kernel that outgrows the I-cache stalls on
// `no_inst`): f64 divisions and square roots only occur in the rarer ops, so they are calls, not ~40 inlined instructions each.
__device__ __noinline__ double ncr_div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __noinline__ double ncr_sqrt(double a) { return __dsqrt_rn(a); }
#undef DIV
#define DIV(a, b) ncr_div((a), (b))

#define FULL 0xffffffffu
#ifndef NCR_EARLY_SHFL
#define NCR_EARLY_SHFL 1   // command walk: the shuffle that selects the entry after next is issued before the staging store (see run_region)
#endif
#ifndef NCR_PRELOAD_INV
#define NCR_PRELOAD_INV 1
#endif
#ifndef NCR_USE_COVERS
#define NCR_USE_COVERS 1   // 0: ignore the entries' "box contains the region" bit (A/B builds)
#endif
// u8 -> k/255.0 table: one private copy per lane (32 copies, entries 256 B apart), so a lookup is conflict-free and its
// shared address is a single byte permute of (texel, lane*8): byte 1 <- texel byte, byte 0 <- lane*8.
#define NCR_LUT_COPIES 32
#define NCR_LUT_BYTES (256 * NCR_LUT_COPIES * 8)
// NCR_TMA_IDENT (experiment X5): per warp two 16x8-texel boxes (512 B each, 128-byte aligned) + two mbarriers
#ifdef NCR_TMA_IDENT
#define NCR_TMA_BYTES_PER_WARP (2 * 512 + 128)
#else
#define NCR_TMA_BYTES_PER_WARP 0
#endif
#define NCR_CMD_SLOT_BYTES ((NCR_COMPOSITE_THREADS / 32) * 2 * (int)sizeof(NcrCmd))
#define NCR_TMA_OFFSET ((NCR_LUT_BYTES + NCR_CMD_SLOT_BYTES + 127) & ~127)
#define NCR_SMEM_BYTES (NCR_TMA_OFFSET + (NCR_COMPOSITE_THREADS / 32) * NCR_TMA_BYTES_PER_WARP)
#ifndef NCR_COMPOSITE_THREADS
#define NCR_COMPOSITE_THREADS 384
#endif
// Pixel slots per lane: NCR_NX columns x 2 rows of 8x4 blocks.  NX=2: a warp owns a 16x8 half-tile (4 px per lane);
// NX=1: an 8x8 quarter-tile (2 px per lane: half the register state, twice the warps per tile).
// NCR_ROW4 (experiment X5, profiles/README.md; not the shipped layout): a lane owns FOUR HORIZONTALLY CONTIGUOUS pixels of one
// row (lane = (lx 0..3, ly 0..7)), so that 1:1 identity-path texels are one 128-bit load and the RGBA8 frame one 128-bit store
// per lane.  The region stays 16x8, so ncr_bin_fine's lists are unchanged.
#ifdef NCR_ROW4
#define NCR_NX 4
#define NCR_NY 1
#define NCR_LANE_X(lane) ((lane) & 3)
#define NCR_LANE_Y(lane) ((lane) >> 2)
#define NCR_SLOT_X(lx, k) (4 * (lx) + (k))
#define NCR_SLOT_Y(ly, k) (ly)
#else
#define NCR_NX 2
#define NCR_NY 2
#define NCR_LANE_X(lane) ((lane) & 7)
#define NCR_LANE_Y(lane) ((lane) >> 3)
#define NCR_SLOT_X(lx, k) (8 * (k) + (lx))
#define NCR_SLOT_Y(ly, k) (4 * (k) + (ly))
#endif
#define NCR_RH NCR_REGION_H                      // region height in pixels (ncr_bin_fine writes lists for 16x8 regions)
#define NCR_P (NCR_NX * NCR_NY)                  // 4 pixel slots per lane
#define NCR_RW NCR_REGION_W                      // region width in pixels
static_assert(NCR_P == 4, "four pixel slots per lane");
#define NCR_TASKS_PER_TILE NCR_REGIONS_PER_TILE  // regions per 16x16 tile
#define SX(p) ((p) % NCR_NX)
#define SY(p) ((p) / NCR_NX)
#ifndef NCR_COMPOSITE_MIN_CTAS
#define NCR_COMPOSITE_MIN_CTAS 1   // one persistent 12-warp CTA per SM: one decode table per SM, the rest of the 256 KB stays L1
#endif

namespace {

// Dynamic shared memory: [0, 64 KB) the decode table, then the per-warp command slots.
extern __shared__ __align__(256) unsigned char ncr_smem[];

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
            r = lut[((t & 255u) * NCR_LUT_COPIES) | l16];
            g = lut[(((t >> 8) & 255u) * NCR_LUT_COPIES) | l16];
            b = lut[(((t >> 16) & 255u) * NCR_LUT_COPIES) | l16];
            a = lut[((t >> 24) * NCR_LUT_COPIES) | l16];
        } else {
            const unsigned char* q = (const unsigned char*)tex + idx * 3;
            r = lut[((uint32_t)__ldg(q) * NCR_LUT_COPIES) | l16];
            g = lut[((uint32_t)__ldg(q + 1) * NCR_LUT_COPIES) | l16];
            b = lut[((uint32_t)__ldg(q + 2) * NCR_LUT_COPIES) | l16];
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

// Slow-path sampler: nearest for non-RGBA8 textures, or NCR_F_BILINEAR (extension, pinned to the reference's commented-out code): the four-tap
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
__device__ __noinline__ bool point_in_poly(const double* __restrict__ pts, uint32_t n, double x, double y) {
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

// ApplyPixel's blend for an already colour-transformed source (reference cpp:533-546), applied under the per-lane
// predicate `in` as selects (no branch, so the four pixel slots of a lane interleave):
//     if (a != 1) c = dst*(1 - a) + c*a;   dst.rgb = c;   dst.a = a (canvas with alpha)
// Lanes with a == 1 (plain store, cpp:533 skips the blend) are NOT handled here: the caller collects them in `opaque`
// and applies store_opaque() under one warp-uniform branch, because they are rare.
__device__ __forceinline__ void padd(double& d, double x, double y, bool p) {
    const double n = ADD(x, y);
    d = p ? n : d;
}

template <bool ALPHA>
__device__ __forceinline__ bool blend(double& dr, double& dg, double& db, double& da, double r, double g, double b, double a,
                                      bool in) {
    const bool ne = (a != 1.0);   // true for NaN, as in C
    const bool blended = in && ne;
    const double om = SUB(1.0, a);
    padd(dr, MUL(dr, om), MUL(r, a), blended);
    padd(dg, MUL(dg, om), MUL(g, a), blended);
    padd(db, MUL(db, om), MUL(b, a), blended);
    if (ALPHA) da = in ? a : da;   // source alpha replaces destination alpha (cpp:544)
    return in && !ne;
}

__device__ __forceinline__ void store_opaque(double& dr, double& dg, double& db, double r, double g, double b, bool opaque) {
    dr = opaque ? r : dr;
    dg = opaque ? g : dg;
    db = opaque ? b : db;
}

// Constant-colour blend (cpp:533-546 with the colour-only subexpressions formed on the host): dst = dst*om + q.
__device__ __forceinline__ void blend_const(double& dr, double& dg, double& db, double om, double q0, double q1, double q2,
                                            bool in) {
    padd(dr, MUL(dr, om), q0, in);
    padd(dg, MUL(dg, om), q1, in);
    padd(db, MUL(db, om), q2, in);
}

__device__ __forceinline__ void store_pred(double& d, double v, bool in) { d = in ? v : d; }

// u8 -> k/255.0 through the replicated table: byte K of the packed texel, this lane's copy.  `lane8` = lane * 8 (< 256);
// the byte offset of entry k, copy lane, is (k << 8) | lane8 — one PRMT — and the table base folds into the LDS immediate.
template <int K>
__device__ __forceinline__ double lut_byte(uint32_t lane8, uint32_t texel) {
    const uint32_t off = __byte_perm(texel, lane8, 0x6504 + (K << 4));
    return *(const double*)((const char*)ncr_smem + off);
}

#define FOR4 _Pragma("unroll") for (int p = 0; p < NCR_P; ++p)

__device__ __forceinline__ bool any_slot(const bool (&in)[NCR_P]) {
    bool r = false;
    FOR4 r = r || in[p];
    return r;
}

// The first parameters every inverse-mapped op needs (inv[0..3]), read from the staged command at the top of the command loop —
// before the fetch of the next command and the dispatch — so that their shared-memory latency is covered by those instead of
// stalling the first multiply of the map (NCR_PRELOAD_INV=0: read where used, A/B builds).
struct InvHead {
    double i0, i1, i2, i3;
};

// Per-lane pixel slots of the current half-tile.
struct Slots {
    int xs[NCR_NX], ys[NCR_NY];        // pixel columns / rows owned by this lane
    double fx[NCR_NX], fy[NCR_NY];     // the same as f64 (int -> f64 is exact)
};

// Shared tail of the textured ops: decode four RGBA8 texels and blend (straight-line, predicated).  RGB_ONE: the colour
// transform's rgb are exactly 1.0, so r * 1.0 (cpp:525-527) is the identity and is not issued.
// NE_TRUE: the recorder proved a != 1 on every pixel (NCR_F_ALPHA_LT1): the blend is unconditional under `in`, and the plain-store
// pass does not exist.
template <bool ALPHA, bool COUNT, bool RGB_ONE, bool NE_TRUE = false>
__device__ __forceinline__ void shade_rgba8(const NcrCmd& c, uint32_t lut_base, const uint32_t (&tx)[NCR_P], const bool (&in)[NCR_P],
                                            double (&dr)[NCR_P], double (&dg)[NCR_P], double (&db)[NCR_P], double (&da)[NCR_P],
                                            unsigned long long& n_applied) {
    const double ct3 = c.ct[3];
    double ct0 = 1.0, ct1 = 1.0, ct2 = 1.0;
    if (!RGB_ONE) { ct0 = c.ct[0]; ct1 = c.ct[1]; ct2 = c.ct[2]; }
    bool opaque[NCR_P];
    FOR4 {
        opaque[p] = false;
        const uint32_t t = tx[p];
        double r = lut_byte<0>(lut_base, t), g = lut_byte<1>(lut_base, t), b = lut_byte<2>(lut_base, t);
        double a = lut_byte<3>(lut_base, t);
        if (!RGB_ONE) { r = MUL(r, ct0); g = MUL(g, ct1); b = MUL(b, ct2); }   // cpp:525-527
        a = MUL(a, ct3);                                                       // cpp:528
        if (NE_TRUE) {
            const double om = SUB(1.0, a);
            padd(dr[p], MUL(dr[p], om), MUL(r, a), in[p]);
            padd(dg[p], MUL(dg[p], om), MUL(g, a), in[p]);
            padd(db[p], MUL(db[p], om), MUL(b, a), in[p]);
            if (ALPHA) da[p] = in[p] ? a : da[p];
        } else {
            opaque[p] = blend<ALPHA>(dr[p], dg[p], db[p], da[p], r, g, b, a, in[p]);
        }
        if (COUNT) n_applied += in[p] ? 1 : 0;
    }
    if (!NE_TRUE && __any_sync(FULL, any_slot(opaque))) {   // a == 1: the source is stored as is
        FOR4 {
            const uint32_t t = tx[p];
            double r = lut_byte<0>(lut_base, t), g = lut_byte<1>(lut_base, t), b = lut_byte<2>(lut_base, t);
            if (!RGB_ONE) { r = MUL(r, ct0); g = MUL(g, ct1); b = MUL(b, ct2); }
            store_opaque(dr[p], dg[p], db[p], r, g, b, opaque[p]);
        }
    }
}

// Bilinear extension (the cpp:575-620 formula the reference keeps commented out; pinned bit-exactly to it) on RGBA8 textures, written like the nearest path: straight-line
// over the four pixel slots, the sixteen texel loads issued back to back, decode through the lane-private table.  Performs
// exactly the operations of sample_slow() + the caller's colour transform and blend, in the same order.
template <bool ALPHA, bool COUNT>
__device__ __forceinline__ void shade_bilinear_rgba8(const NcrCmd& c, uint32_t lut_base, int tw, int th, const double (&u)[NCR_P],
                                                     const double (&v)[NCR_P], const bool (&in)[NCR_P], double (&dr)[NCR_P],
                                                     double (&dg)[NCR_P], double (&db)[NCR_P], double (&da)[NCR_P],
                                                     unsigned long long& n_applied) {
    const uint32_t* t32 = (const uint32_t*)c.tex;
    uint32_t t[NCR_P][4];
    double fu[NCR_P], fv[NCR_P];
    FOR4 {
        double uu = u[p], vv = v[p];
        clamp_uv(uu, vv, tw, th);
        int xi = __double2int_rz(uu), yi = __double2int_rz(vv);
        xi = min(max(xi, 0), tw - 1);   // memory safety only; no-op for defined inputs
        yi = min(max(yi, 0), th - 1);
        const int dx = xi + 1 < tw ? 1 : 0, dy = yi + 1 < th ? tw : 0;
        const int idx = yi * tw + xi;
        t[p][0] = in[p] ? __ldg(t32 + idx) : 0u;
        t[p][1] = in[p] ? __ldg(t32 + idx + dx) : 0u;
        t[p][2] = in[p] ? __ldg(t32 + idx + dy) : 0u;
        t[p][3] = in[p] ? __ldg(t32 + idx + dy + dx) : 0u;
        fu[p] = SUB(uu, (double)xi);
        fv[p] = SUB(vv, (double)yi);
    }
    const double ct0 = c.ct[0], ct1 = c.ct[1], ct2 = c.ct[2], ct3 = c.ct[3];
    FOR4 {
        const double mu = SUB(1.0, fu[p]), mv = SUB(1.0, fv[p]);
#define NCR_TAP(K) ADD(ADD(ADD(MUL(MUL(lut_byte<K>(lut_base, t[p][0]), mu), mv), MUL(MUL(lut_byte<K>(lut_base, t[p][1]), fu[p]), mv)), \
                           MUL(MUL(lut_byte<K>(lut_base, t[p][2]), mu), fv[p])), MUL(MUL(lut_byte<K>(lut_base, t[p][3]), fu[p]), fv[p]))
        const double r = MUL(NCR_TAP(0), ct0), g = MUL(NCR_TAP(1), ct1), b = MUL(NCR_TAP(2), ct2), a = MUL(NCR_TAP(3), ct3);
#undef NCR_TAP
        const bool opq = blend<ALPHA>(dr[p], dg[p], db[p], da[p], r, g, b, a, in[p]);
        store_opaque(dr[p], dg[p], db[p], r, g, b, opq);
        if (COUNT) n_applied += in[p] ? 1 : 0;
    }
}

// Fast path of the two hot ops — DrawTexture (inverse-mapped, cpp:753-778) and DrawSplittedTexture (cpp:781-820) on
// RGBA8 textures with nearest sampling.  Written as straight-line code over the four pixel slots (no branch between
// slots), so the f64 dependency chains of the slots interleave.
// INTERIOR: ncr_bin_fine proved that every pixel of the region passes the box and the four bounds: `in` is constant true, the
// bounds are not evaluated and every select on coverage folds away.
template <bool ALPHA, bool COUNT, bool INTERIOR>
__device__ __forceinline__ void tex_fast(const NcrCmd& c, const InvHead& ih, const uint32_t hint /* the list entry's NCR_ENTRY_* bits */, const Slots& S,
                                         const double* lut, uint32_t lut_base, bool (&in)[NCR_P], double (&dr)[NCR_P], double (&dg)[NCR_P],
                                         double (&db)[NCR_P], double (&da)[NCR_P], unsigned long long& n_applied) {
    // TransformPointFromMatrix(inv, i, j), cpp:451-452: (inv0*i + inv2*j) + inv4.  inv0*i depends only on the pixel
    // column and inv2*j only on the row: each product is formed once per lane.
    const double i0 = ih.i0, i1 = ih.i1, i2 = ih.i2, i3 = ih.i3, i4 = c.inv[4], i5 = c.inv[5];
    double ax[NCR_NX], bx[NCR_NX], ay[NCR_NY], by[NCR_NY];
#pragma unroll
    for (int k = 0; k < NCR_NX; ++k) { ax[k] = MUL(i0, S.fx[k]); bx[k] = MUL(i1, S.fx[k]); }
#pragma unroll
    for (int k = 0; k < NCR_NY; ++k) { ay[k] = MUL(i2, S.fy[k]); by[k] = MUL(i3, S.fy[k]); }
    const double cx = c.x, cy = c.y, cxw = c.xw, cyh = c.yh, sx = c.sx, sy = c.sy;
    const int tw = c.tex_w, tw2 = tw - 2, th2 = c.tex_h - 2;
    const uint32_t* t32 = (const uint32_t*)c.tex;
    // cpp:812-813: u = (uStart + (uEnd - uStart) * u / tex->width) * tex->width.  The hot path only takes power-of-two textures
    // here (the recorder routes the others to the general path): x / 2^k and x * 2^-k are the same correctly rounded value.
    const double uS = c.p[0], dU = c.p[1], vS = c.p[2], dV = c.p[3], fw = c.p[4], fh = c.p[5], rw = c.p[6], rh = c.p[7];
    uint32_t tx[NCR_P];
    // Each slot is mapped, clamped and fetched in one go (no u[] / v[] arrays live across the slots).  The split remap is a
    // compile-time variant of the slot loop, chosen by one warp-uniform branch: with the test inside the loop ptxas predicates the
    // eight remap instructions of every slot, which non-split draws then issue for nothing (profiles/README.md, r2 session 3).
    // InterpolateColorFromBuffer, cpp:560-566: clamp u<0 -> 0, u >= w-1 -> w-2, then (i64) truncation.  Done after the
    // truncation here, which is the same function: trunc(u) <= 0 iff u < 1, and because w-1 is an integer,
    // u >= w-1 iff trunc(u) >= w-1 (cvt.rzi saturates, so huge u stays >= w-1).  x >= w-1 ? w-2 : x is min(x, w-2) on
    // integers; max(.., 0) also keeps 1-texel-wide textures in bounds (the reference reads out of bounds there).
    auto map_and_fetch = [&](auto split_c) {
        constexpr bool SPLIT = decltype(split_c)::value;
        FOR4 {
            const double X = ADD(ADD(ax[SX(p)], ay[SY(p)]), i4);
            const double Y = ADD(ADD(bx[SX(p)], by[SY(p)]), i5);
            // the four inclusive bounds, cpp:765-768 (NaN compares false on both sides, as in C)
            if (!INTERIOR) in[p] = in[p] && !(X < cx) && !(X > cxw) && !(Y < cy) && !(Y > cyh);
            double u = MUL(SUB(X, cx), sx);   // cpp:770-771
            double v = MUL(SUB(Y, cy), sy);
            if (SPLIT) {
                u = MUL(ADD(uS, MUL(MUL(dU, u), rw)), fw);
                v = MUL(ADD(vS, MUL(MUL(dV, v), rh)), fh);
            }
            const int xi = __vimin_s32_relu(__double2int_rz(u), tw2);
            const int yi = __vimin_s32_relu(__double2int_rz(v), th2);
            tx[p] = (INTERIOR || in[p]) ? __ldg(t32 + (yi * tw + xi)) : 0u;
        }
    };
    if (hint & NCR_ENTRY_SPLIT) map_and_fetch(std::true_type{});
    else map_and_fetch(std::false_type{});
    if (!INTERIOR && !__any_sync(FULL, any_slot(in))) return;
    if ((hint & (NCR_ENTRY_RGB_ONE | NCR_ENTRY_ALPHA_LT1)) == (NCR_ENTRY_RGB_ONE | NCR_ENTRY_ALPHA_LT1))
        shade_rgba8<ALPHA, COUNT, true, true>(c, lut_base, tx, in, dr, dg, db, da, n_applied);
    else if (hint & NCR_ENTRY_RGB_ONE) shade_rgba8<ALPHA, COUNT, true>(c, lut_base, tx, in, dr, dg, db, da, n_applied);
    else shade_rgba8<ALPHA, COUNT, false>(c, lut_base, tx, in, dr, dg, db, da, n_applied);
}

// A command tagged INTERIOR by ncr_bin_fine: every pixel of the region is inside the command's box and passes its coverage
// test, so the common ops run without any per-pixel test and without selects on coverage (the arithmetic per pixel is the
// same as in apply_cmd — only `in` is the constant true).  Returns false (warp-uniform) for ops that have no interior
// variant; the caller then runs apply_cmd, which is always correct.
template <bool ALPHA, bool COUNT>
__device__ __forceinline__ bool apply_interior(const NcrCmd& c, const InvHead& ih, const uint32_t hint, const Slots& S, const double* lut, uint32_t lut_base,
                                               double (&dr)[NCR_P], double (&dg)[NCR_P], double (&db)[NCR_P], double (&da)[NCR_P],
                                               unsigned long long& n_applied, const uint32_t* tbox = nullptr, int tbox_at = 0) {
    bool in[NCR_P];
    FOR4 in[p] = true;
    // The hot case is chosen from the list entry (a register), not from the staged command: its parameter loads do not wait for
    // the flag word's shared-memory round trip.
    if (hint & NCR_ENTRY_FAST_AFFINE) {
        tex_fast<ALPHA, COUNT, true>(c, ih, hint, S, lut, lut_base, in, dr, dg, db, da, n_applied);
        return true;
    }
    const uint32_t op = c.op, flags = c.flags;
    if (op == NCR_OP_FILL_COLOR || op == NCR_OP_RECT) {   // constant colour, see apply_cmd
        const double sa = c.p[3];
        if (sa != 1.0) {
            const double q0 = c.p[4], q1 = c.p[5], q2 = c.p[6], om = c.p[7];
            FOR4 {
                dr[p] = ADD(MUL(dr[p], om), q0); dg[p] = ADD(MUL(dg[p], om), q1); db[p] = ADD(MUL(db[p], om), q2);
                if (ALPHA) da[p] = sa;
            }
        } else {
            FOR4 {
                dr[p] = c.p[0]; dg[p] = c.p[1]; db[p] = c.p[2];
                if (ALPHA) da[p] = sa;
            }
        }
        if (COUNT) n_applied += NCR_P;
        return true;
    }
    if (op == NCR_OP_SET_COLOR && !(flags & NCR_F_RGB_SPILL)) {   // cpp:643-657
        FOR4 {
            dr[p] = c.p[0]; dg[p] = c.p[1]; db[p] = c.p[2];
            if (ALPHA) da[p] = c.p[3];
        }
        return true;
    }
    if (op == NCR_OP_TEX_IDENT && (flags & NCR_F_TEX_FAST)) {   // cpp:741-751 on RGBA8 texels
        const double cx = c.x, cy = c.y, sx = c.sx, sy = c.sy;
        const int tw = c.tex_w, tw2 = tw - 2, th2 = c.tex_h - 2;
        const uint32_t* t32 = (const uint32_t*)c.tex;
        int xi[NCR_NX], yi[NCR_NY];
#pragma unroll
        for (int k = 0; k < NCR_NX; ++k) xi[k] = __vimin_s32_relu(__double2int_rz(MUL(SUB(S.fx[k], cx), sx)), tw2);
#pragma unroll
        for (int k = 0; k < NCR_NY; ++k) yi[k] = __vimin_s32_relu(__double2int_rz(MUL(SUB(S.fy[k], cy), sy)), th2) * tw;
        uint32_t tx[NCR_P];
#ifdef NCR_TMA_IDENT
        if (tbox) {   // experiment X5: the region's 16x8 texel box was staged in shared memory by the TMA unit
            FOR4 tx[p] = tbox[tbox_at + NCR_SLOT_Y(0, SY(p)) * NCR_RW + NCR_SLOT_X(0, SX(p))];
            if (flags & NCR_F_CT_RGB_ONE) shade_rgba8<ALPHA, COUNT, true>(c, lut_base, tx, in, dr, dg, db, da, n_applied);
            else shade_rgba8<ALPHA, COUNT, false>(c, lut_base, tx, in, dr, dg, db, da, n_applied);
            return true;
        }
#endif
#ifdef NCR_ROW4
        // 1:1 sampling without clamping: the lane's four texels are contiguous; one 128-bit load when they are 16-byte aligned
        const int t0 = yi[0] + xi[0];
        if (xi[3] == xi[0] + 3 && (t0 & 3) == 0) {
            const uint4 q = __ldg((const uint4*)(t32 + t0));
            tx[0] = q.x; tx[1] = q.y; tx[2] = q.z; tx[3] = q.w;
        } else
#endif
        {
            FOR4 tx[p] = __ldg(t32 + (yi[SY(p)] + xi[SX(p)]));
        }
        if (flags & NCR_F_CT_RGB_ONE) shade_rgba8<ALPHA, COUNT, true>(c, lut_base, tx, in, dr, dg, db, da, n_applied);
        else shade_rgba8<ALPHA, COUNT, false>(c, lut_base, tx, in, dr, dg, db, da, n_applied);
        return true;
    }
    return false;
}

// One command applied to the warp's 128 pixels.  `c` lives in shared memory (warp-uniform reads: one wavefront each).
template <bool ALPHA, bool COUNT>
__device__ __forceinline__ void apply_cmd(const NcrCmd& c, const InvHead& ih, const NcrFlushArgs& A, const Slots& S,
                                          const double* lut, uint32_t lut_base /* lane * 8 */,
                                          double (&dr)[NCR_P], double (&dg)[NCR_P], double (&db)[NCR_P], double (&da)[NCR_P],
                                          unsigned long long& n_applied, const bool covers /* warp-uniform: box contains the region */,
                                          const uint32_t hint /* the list entry's NCR_ENTRY_* bits */) {
    // pixel-box membership: the reference's loop bounds (boxes are clamped to the canvas on the host, so a pixel slot
    // outside the canvas is never inside a box).  (unsigned)(v - lo) < (unsigned)(hi - lo)  <=>  lo <= v < hi.
    // ncr_bin_fine tags the entry when the box contains the whole region (the usual case at the slanted edge of a rotated quad,
    // whose bounding box is larger than the quad): every slot is inside, the tests are skipped.
    bool in[NCR_P];
    FOR4 in[p] = true;
    if (!covers) {
        const unsigned wx = (unsigned)(c.r - c.l), wy = (unsigned)(c.b - c.t);
        bool inx[NCR_NX], iny[NCR_NY];
#pragma unroll
        for (int k = 0; k < NCR_NX; ++k) inx[k] = (unsigned)(S.xs[k] - c.l) < wx;
#pragma unroll
        for (int k = 0; k < NCR_NY; ++k) iny[k] = (unsigned)(S.ys[k] - c.t) < wy;
        FOR4 in[p] = inx[SX(p)] && iny[SY(p)];
    }

    if (hint & NCR_ENTRY_FAST_AFFINE) {   // DrawTexture / DrawSplittedTexture on RGBA8, nearest: the hot case, tested first
        tex_fast<ALPHA, COUNT, false>(c, ih, hint, S, lut, lut_base, in, dr, dg, db, da, n_applied);
        return;
    }
    const uint32_t op = c.op, flags = c.flags;

    if (op == NCR_OP_SET_COLOR) {   // cpp:643-657
        FOR4 if (in[p]) {
            dr[p] = c.p[0]; dg[p] = c.p[1]; db[p] = c.p[2];
            if (ALPHA) da[p] = c.p[3];
            // 3-channel canvas, non-uniform colour: SetPixel's index+3 store (cpp:510) leaves `a` in the red of pixel
            // (0, j>=1), because column 0 is written first and the last column of the previous row spills into it.
            else if ((flags & NCR_F_RGB_SPILL) && S.xs[SX(p)] == 0 && S.ys[SY(p)] >= 1 && A.d.w > 1) dr[p] = c.p[3];
        }
        return;
    }
    if (op == NCR_OP_SET_PIXEL) {   // cpp:494-513
        FOR4 if (in[p]) {
            dr[p] = c.p[0];
            if (!(flags & NCR_F_ONLY_RED)) {
                dg[p] = c.p[1]; db[p] = c.p[2];
                if (ALPHA) da[p] = c.p[3];
            }
        }
        return;
    }

    // inverse-mapped source position, TransformPointFromMatrix(inv, i, j), cpp:451-452: (inv0*i + inv2*j) + inv4.
    double X[NCR_P], Y[NCR_P];
    if (op != NCR_OP_FILL_COLOR && op != NCR_OP_APPLY_PIXEL && op != NCR_OP_TEX_IDENT && op != NCR_OP_TEX_PERSP) {
        double ax[NCR_NX], bx[NCR_NX], ay[NCR_NY], by[NCR_NY];
#pragma unroll
        for (int k = 0; k < NCR_NX; ++k) { ax[k] = MUL(ih.i0, S.fx[k]); bx[k] = MUL(ih.i1, S.fx[k]); }
#pragma unroll
        for (int k = 0; k < NCR_NY; ++k) { ay[k] = MUL(ih.i2, S.fy[k]); by[k] = MUL(ih.i3, S.fy[k]); }
        const double i4 = c.inv[4], i5 = c.inv[5];
        FOR4 {
            X[p] = ADD(ADD(ax[SX(p)], ay[SY(p)]), i4);
            Y[p] = ADD(ADD(bx[SX(p)], by[SY(p)]), i5);
        }
        if (op != NCR_OP_CIRCLE && op != NCR_OP_POLY) {
            // the four inclusive bounds, cpp:765-768 (NaN compares false on both sides, as in C)
            const double cx = c.x, cy = c.y, cxw = c.xw, cyh = c.yh;
            FOR4 in[p] = in[p] && !(X[p] < cx) && !(X[p] > cxw) && !(Y[p] < cy) && !(Y[p] > cyh);
        }
    }

    if (op == NCR_OP_GRAD) {   // cpp:1308-1313; p[4..7] = bottom - top
        const double hgt = c.sy, cy = c.y;
        if (!__any_sync(FULL, any_slot(in))) return;
        FOR4 {
            const double t = DIV(SUB(Y[p], cy), hgt);
            const double r = MUL(ADD(c.p[0], MUL(c.p[4], t)), c.ct[0]);
            const double g = MUL(ADD(c.p[1], MUL(c.p[5], t)), c.ct[1]);
            const double b = MUL(ADD(c.p[2], MUL(c.p[6], t)), c.ct[2]);
            const double a = MUL(ADD(c.p[3], MUL(c.p[7], t)), c.ct[3]);
            const bool opq = blend<ALPHA>(dr[p], dg[p], db[p], da[p], r, g, b, a, in[p]);
            store_opaque(dr[p], dg[p], db[p], r, g, b, opq);
            if (COUNT) n_applied += in[p] ? 1 : 0;
        }
        return;
    }

    if (op == NCR_OP_RECT || op == NCR_OP_CIRCLE || op == NCR_OP_POLY || op == NCR_OP_FILL_COLOR || op == NCR_OP_APPLY_PIXEL) {
        if (op == NCR_OP_CIRCLE) {   // cpp:939-943
            const double cx = c.x, cy = c.y, rad = c.sx;
            FOR4 if (__any_sync(FULL, in[p])) {
                const double ddx = SUB(X[p], cx), ddy = SUB(Y[p], cy);
                const double dist = ncr_sqrt(ADD(MUL(ddx, ddx), MUL(ddy, ddy)));
                in[p] = in[p] && !(dist > rad);
            }
        } else if (op == NCR_OP_POLY) {   // cpp:913
            const double* pts = A.aux + c.aux_off;
            const uint32_t npts = c.aux_n;
            FOR4 if (__any_sync(FULL, in[p])) in[p] = in[p] && point_in_poly(pts, npts, X[p], Y[p]);
        }
        // constant colour: p[0..3] = colour * ct, p[4..6] = rgb*a, p[7] = 1-a, formed on the host with the same IEEE
        // operations ApplyPixel performs per pixel (cpp:525-536).
        const double sa = c.p[3];
        if (sa != 1.0) {
            const double q0 = c.p[4], q1 = c.p[5], q2 = c.p[6], om = c.p[7];
            FOR4 {
                blend_const(dr[p], dg[p], db[p], om, q0, q1, q2, in[p]);
                if (ALPHA) store_pred(da[p], sa, in[p]);
                if (COUNT) n_applied += in[p] ? 1 : 0;
            }
        } else {
            const double c0 = c.p[0], c1 = c.p[1], c2 = c.p[2];
            FOR4 {
                store_pred(dr[p], c0, in[p]); store_pred(dg[p], c1, in[p]); store_pred(db[p], c2, in[p]);
                if (ALPHA) store_pred(da[p], sa, in[p]);
                if (COUNT) n_applied += in[p] ? 1 : 0;
            }
        }
        return;
    }

    // ---- textured ops: DrawTexture (both paths), DrawSplittedTexture, perspective extension ----
    const int tw = c.tex_w, th = c.tex_h;
    double u[NCR_P], v[NCR_P];
    if (op == NCR_OP_TEX_IDENT) {   // cpp:741-745: pixels i >= (i64)x with (f64)i < x + width; u = (i - x) * scaleX
        FOR4 {
            const double fi = S.fx[SX(p)], fj = S.fy[SY(p)];
            in[p] = in[p] && fi >= c.p[0] && fi < c.xw && fj >= c.p[1] && fj < c.yh;
            u[p] = MUL(SUB(fi, c.x), c.sx);
            v[p] = MUL(SUB(fj, c.y), c.sy);
        }
    } else if (op == NCR_OP_TEX_PERSP) {   // extension (this repo's own spec): row-major 3x3 inverse homography, one reciprocal
        // of the homogeneous w per pixel and two multiplies, then cpp:765-771.  The per-column / per-row products are hoisted.
        double hx[NCR_NX], ax[NCR_NX], bx[NCR_NX], hy[NCR_NY], ay[NCR_NY], by[NCR_NY];
#pragma unroll
        for (int k = 0; k < NCR_NX; ++k) { hx[k] = MUL(c.p[0], S.fx[k]); ax[k] = MUL(c.inv[0], S.fx[k]); bx[k] = MUL(c.inv[3], S.fx[k]); }
#pragma unroll
        for (int k = 0; k < NCR_NY; ++k) { hy[k] = MUL(c.p[1], S.fy[k]); ay[k] = MUL(c.inv[1], S.fy[k]); by[k] = MUL(c.inv[4], S.fy[k]); }
        FOR4 {
            const double hw = ADD(ADD(hx[SX(p)], hy[SY(p)]), c.p[2]);
            const double rw = DIV(1.0, hw);
            const double Xp = MUL(ADD(ADD(ax[SX(p)], ay[SY(p)]), c.inv[2]), rw);
            const double Yp = MUL(ADD(ADD(bx[SX(p)], by[SY(p)]), c.inv[5]), rw);
            in[p] = in[p] && hw > 0.0 && !(Xp < c.x) && !(Xp > c.xw) && !(Yp < c.y) && !(Yp > c.yh);
            u[p] = MUL(SUB(Xp, c.x), c.sx);
            v[p] = MUL(SUB(Yp, c.y), c.sy);
        }
    } else {   // NCR_OP_TEX / NCR_OP_TEX_SPLIT, cpp:770-771
        const double cx = c.x, cy = c.y, sx = c.sx, sy = c.sy;
        FOR4 {
            u[p] = MUL(SUB(X[p], cx), sx);
            v[p] = MUL(SUB(Y[p], cy), sy);
        }
        if (op == NCR_OP_TEX_SPLIT) {
            // cpp:812-813: u = (uStart + (uEnd - uStart) * u / tex->width) * tex->width
            const double uS = c.p[0], dU = c.p[1], vS = c.p[2], dV = c.p[3], fw = c.p[4], fh = c.p[5];
            if (flags & NCR_F_SPLIT_POW2) {
                // width and height are powers of two: x / 2^k and x * 2^-k are the same correctly rounded value
                const double rw = c.p[6], rh = c.p[7];
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

    if (flags & NCR_F_TEX_FAST) {   // RGBA8 texels, nearest, < 2^31 texels (clamp-after-truncate: see tex_fast)
        const uint32_t* t32 = (const uint32_t*)c.tex;
        uint32_t tx[NCR_P];
        FOR4 {
            const int xi = __vimin_s32_relu(__double2int_rz(u[p]), tw - 2);
            const int yi = __vimin_s32_relu(__double2int_rz(v[p]), th - 2);
            tx[p] = in[p] ? __ldg(t32 + (yi * tw + xi)) : 0u;
        }
        shade_rgba8<ALPHA, COUNT, false>(c, lut_base, tx, in, dr, dg, db, da, n_applied);   // r * 1.0 is exact: one code copy
    } else if ((flags & (NCR_F_BILINEAR | NCR_F_TEX_ALPHA | NCR_F_TEX_F64)) == (NCR_F_BILINEAR | NCR_F_TEX_ALPHA) &&
               (unsigned long long)tw * (unsigned long long)th < (1ull << 31)) {
        if (__any_sync(FULL, any_slot(in))) shade_bilinear_rgba8<ALPHA, COUNT>(c, lut_base, tw, th, u, v, in, dr, dg, db, da, n_applied);
    } else {
        FOR4 if (in[p]) {
            double s[4];
            sample_slow(c.tex, flags, tw, th, lut, threadIdx.x & (NCR_LUT_COPIES - 1), u[p], v[p], s);
            const double r = MUL(s[0], c.ct[0]), g = MUL(s[1], c.ct[1]), b = MUL(s[2], c.ct[2]), a = MUL(s[3], c.ct[3]);
            const bool opq = blend<ALPHA>(dr[p], dg[p], db[p], da[p], r, g, b, a, true);
            store_opaque(dr[p], dg[p], db[p], r, g, b, opq);
            if (COUNT) ++n_applied;
        }
    }
}

#ifdef NCR_TMA_IDENT
// ---- experiment X5: TMA staging of a region's texel box (cp.async.bulk.tensor.2d + mbarrier), see profiles/README.md ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_mbar_init(uint32_t mbar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
}
// one elected lane: expect 512 bytes, then start the bulk tensor copy of the 16x8 box whose first texel is (u0, v0)
__device__ __forceinline__ void tma_load_box(uint32_t dst, const void* map, int u0, int v0, uint32_t mbar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 512;" ::"r"(mbar) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(u0), "r"(v0), "r"(mbar) : "memory");
}
__device__ __forceinline__ void tma_wait(uint32_t mbar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}"
                 ::"r"(mbar), "r"(parity) : "memory");
}
// Can region `task` take its background texels from a TMA box?  (no clamping anywhere in the region: cpp:560-563 are no-ops)
__device__ __forceinline__ bool tma_region_ok(const NcrFlushArgs& A, int task, int& u0, int& v0) {
    const int tile = task / NCR_TASKS_PER_TILE, sub = task % NCR_TASKS_PER_TILE;
    const int x0 = (tile % A.d.tiles_x) * NCR_TILE, y0 = (tile / A.d.tiles_x) * NCR_TILE + sub * NCR_RH;
    u0 = x0 - A.tma_x;
    v0 = y0 - A.tma_y;
    return A.tma_map != nullptr && u0 >= 0 && v0 >= 0 && u0 + NCR_RW - 1 <= A.tma_w - 2 && v0 + NCR_RH - 1 <= A.tma_h - 2 &&
           x0 + NCR_RW <= A.d.w && y0 + NCR_RH <= A.d.h;
}
#endif

// Copies command `e` (list entry) from HBM into a per-warp shared slot: 15 lanes x 16 bytes.
__device__ __forceinline__ void stage_cmd(const NcrFlushArgs& A, NcrCmd* dst, uint32_t e, int lane) {
    constexpr uint32_t WORDS = NCR_CMD_WORDS16;   // 15 x 16 B
    if (lane < (int)WORDS) ((uint4*)dst)[lane] = __ldg((const uint4*)(A.cmds + (e & NCR_ENTRY_INDEX)) + lane);
    __syncwarp();
}

// One region (16x8 px) composited by one warp: pixel set-up, optional canvas read, the region's command list in submission
// order, write-back (f64 canvas unless write_fb == 0, fused u8 image).
//   ents       the first 32 entries of the region's list, one per lane (already loaded);
//   slot       in/out: which of the warp's two shared command slots holds the command to run next;
//   have_cmd0  the region's first command is already staged in s_cmd[slot] (cross-region prefetch);
//   next_valid / next_ents   cross-region prefetch: the first list entries (one per lane) of the region this warp composites
//              next; its first command is fetched while this region's last command is applied (the entries were loaded a whole
//              region earlier and are only touched then).  Returns true when that command is staged in s_cmd[slot].
template <bool ALPHA, bool COUNT>
__device__ __forceinline__ bool run_region(const NcrFlushArgs& A, NcrCmd (*s_cmd)[2], const int task, const uint32_t loff,
                                           const uint32_t lcount, uint32_t ents, int& slot, const bool have_cmd0,
                                           const bool next_valid, const uint32_t next_ents, const double* lut,
                                           unsigned long long& n_applied, const uint32_t* tbox = nullptr, uint32_t tbox_mbar = 0,
                                           uint32_t tbox_parity = 0, bool* tbox_used = nullptr) {
    const int lane = threadIdx.x & 31;
    const uint32_t lut_base = (uint32_t)lane * 8u;
    const int lx = NCR_LANE_X(lane), ly = NCR_LANE_Y(lane);
    constexpr int IPP = ALPHA ? 4 : 3;
    constexpr uint32_t WORDS = NCR_CMD_WORDS16;
    const int W = A.d.w, H = A.d.h;
    const uint32_t* __restrict__ list = A.fine_list;
    const int tile = task / NCR_TASKS_PER_TILE, sub = task % NCR_TASKS_PER_TILE;
    const int x0 = (tile % A.d.tiles_x) * NCR_TILE + (sub % (16 / NCR_RW)) * NCR_RW;
    const int y0 = (tile / A.d.tiles_x) * NCR_TILE + (sub / (16 / NCR_RW)) * NCR_RH;
    if ((lcount == 0 && A.u8_out == nullptr) || y0 >= H || x0 >= W) {   // nothing to do here
        if (next_valid) stage_cmd(A, &s_cmd[0][slot], __shfl_sync(FULL, next_ents, 0), lane);
        return next_valid;
    }
    Slots S;
#pragma unroll
    for (int k = 0; k < NCR_NX; ++k) { S.xs[k] = x0 + NCR_SLOT_X(lx, k); S.fx[k] = (double)S.xs[k]; }
#pragma unroll
    for (int k = 0; k < NCR_NY; ++k) { S.ys[k] = y0 + NCR_SLOT_Y(ly, k); S.fy[k] = (double)S.ys[k]; }
    bool valid[NCR_P];
    FOR4 valid[p] = S.xs[SX(p)] < W && S.ys[SY(p)] < H;

    double dr[NCR_P], dg[NCR_P], db[NCR_P], da[NCR_P];
    FOR4 { dr[p] = 0.0; dg[p] = 0.0; db[p] = 0.0; da[p] = 0.0; }
    // The canvas is read unless the list starts with a SetColor (then every pixel is overwritten first).
    if (A.load_fb != 0 || lcount == 0) {
        FOR4 if (valid[p]) {
            const double* q = A.fb + ((size_t)S.ys[SY(p)] * W + S.xs[SX(p)]) * IPP;
            if (ALPHA) {
                const double2 lo = ((const double2*)q)[0], hi = ((const double2*)q)[1];
                dr[p] = lo.x; dg[p] = lo.y; db[p] = hi.x; da[p] = hi.y;
            } else {
                dr[p] = q[0]; dg[p] = q[1]; db[p] = q[2];
            }
        }
    }

    // Walk the region's list in submission order.  Every entry is a command to run (ncr_bin_fine did the culling); the
    // next command — at the end of the list, the first command of the warp's next region — is fetched into the other
    // shared slot while the current one is applied.
    uint32_t cur = __shfl_sync(FULL, ents, 0);
    if (lcount != 0 && !have_cmd0) stage_cmd(A, &s_cmd[0][slot], cur, lane);
    bool staged_next = false;
    if (lcount == 0 && next_valid) { stage_cmd(A, &s_cmd[0][slot], __shfl_sync(FULL, next_ents, 0), lane); staged_next = true; }
    // Every lane takes part in staging (lanes 15..31 repeat the last 16-byte word: same value to the same address), so the fetch and
    // the store carry no predicate.  The list is walked 32 entries at a time: the next chunk's entries (at the end of the list, the
    // next region's first entries in the prefetch variant) are loaded at the top of a chunk and first touched at its last command, so
    // the inner loop holds no list address and no refill test.  Past the end of everything the "next command" is entry 0 of the
    // current chunk — a valid command index whose copy is staged and never run.
    const uint32_t cmd_word = min(lane, (int)WORDS - 1);
    const uint4* const cmd_words = (const uint4*)A.cmds + cmd_word;    // + 15 * index: this lane's word of command `index`
    for (uint32_t base = 0; base < lcount; base += 32) {
        const uint32_t n_here = min(32u, lcount - base);
        uint32_t ents_n = ents;
        if (base + 32 < lcount) ents_n = (base + 32 + lane < lcount) ? __ldg(list + loff + base + 32 + lane) : 0u;
        else if (next_valid) ents_n = next_ents;
#if NCR_EARLY_SHFL
        // entry of the command after command 1 of this chunk (source lane is taken modulo 32); after the chunk's last command: the
        // next chunk's / region's first.  Inside the chunk the shuffle for command i + 2 is issued right after command i has been
        // applied — before the staging store and the warp sync — so its latency is covered by them instead of stalling the fetch
        // of the next command at the top of the loop.
        uint32_t nxt = __shfl_sync(FULL, n_here == 1 ? ents_n : ents, n_here == 1 ? 0u : 1u);
#endif
        for (uint32_t i = 1; i <= n_here; ++i) {
#if !NCR_EARLY_SHFL
            // command i of this chunk (source lane is taken modulo 32); after the chunk's last one: the next chunk's / region's first
            const bool last = i == n_here;
            const uint32_t nxt = __shfl_sync(FULL, last ? ents_n : ents, last ? 0u : i);
#endif
            const uint4 pre = __ldg(cmd_words + (size_t)(nxt & NCR_ENTRY_INDEX) * WORDS);   // in flight during the apply
            const NcrCmd& c = s_cmd[0][slot];
#if NCR_PRELOAD_INV
            const InvHead ih = {c.inv[0], c.inv[1], c.inv[2], c.inv[3]};
#else
            const InvHead& ih = *(const InvHead*)&c.inv[0];
#endif
#ifdef NCR_TMA_IDENT
            if (tbox && (cur & ~NCR_ENTRY_HINTS) == ((uint32_t)A.tma_cmd | NCR_ENTRY_INTERIOR | NCR_ENTRY_COVERS)) {   // the staged box belongs to this command
                tma_wait(tbox_mbar, tbox_parity);
                *tbox_used = true;
                apply_interior<ALPHA, COUNT>(c, ih, cur, S, lut, lut_base, dr, dg, db, da, n_applied, tbox, ly * NCR_RW + lx);
            } else
#endif
            if (!(cur & NCR_ENTRY_INTERIOR) || !apply_interior<ALPHA, COUNT>(c, ih, cur, S, lut, lut_base, dr, dg, db, da, n_applied))
                apply_cmd<ALPHA, COUNT>(c, ih, A, S, lut, lut_base, dr, dg, db, da, n_applied, NCR_USE_COVERS && (cur & NCR_ENTRY_COVERS) != 0, cur);
#if NCR_EARLY_SHFL
            const bool last2 = i + 1 == n_here;
            const uint32_t nn = __shfl_sync(FULL, last2 ? ents_n : ents, last2 ? 0u : i + 1);   // i == n_here: unused (recomputed above)
#endif
            slot ^= 1;
            ((uint4*)&s_cmd[0][slot])[cmd_word] = pre;
            __syncwarp();
            cur = nxt;
#if NCR_EARLY_SHFL
            nxt = nn;
#endif
        }
        ents = ents_n;
    }
    if (lcount != 0) staged_next = next_valid;

    // region write-back: canonical f64 canvas (only if something was drawn, and not for a present-only flush) and the fused (iu8)(v*255) image
#ifdef NCR_ROW4
    bool u8_done = false;
    if (ALPHA && A.u8_out && valid[3] && (W & 3) == 0) {   // the lane's four RGBA8 pixels: one 128-bit store
        uint32_t o[4];
        FOR4 o[p] = (uint32_t)ncr_to_u8(dr[p]) | ((uint32_t)ncr_to_u8(dg[p]) << 8) | ((uint32_t)ncr_to_u8(db[p]) << 16) |
                    ((uint32_t)ncr_to_u8(da[p]) << 24);
        *(uint4*)((uint32_t*)A.u8_out + (size_t)S.ys[0] * W + S.xs[0]) = make_uint4(o[0], o[1], o[2], o[3]);
        u8_done = true;
    }
#else
    const bool u8_done = false;
#endif
    FOR4 if (valid[p]) {
        const size_t pix = ((size_t)S.ys[SY(p)] * W + S.xs[SX(p)]) * IPP;
        if (lcount != 0 && A.write_fb) {
            double* q = A.fb + pix;
            if (ALPHA) {
                ((double2*)q)[0] = make_double2(dr[p], dg[p]);
                ((double2*)q)[1] = make_double2(db[p], da[p]);
            } else {
                q[0] = dr[p]; q[1] = dg[p]; q[2] = db[p];
            }
        }
        if (A.u8_out && !u8_done) {
            if (ALPHA) {
                const uint32_t o = (uint32_t)ncr_to_u8(dr[p]) | ((uint32_t)ncr_to_u8(dg[p]) << 8) |
                                   ((uint32_t)ncr_to_u8(db[p]) << 16) | ((uint32_t)ncr_to_u8(da[p]) << 24);
                ((uint32_t*)A.u8_out)[(size_t)S.ys[SY(p)] * W + S.xs[SX(p)]] = o;
            } else {
                unsigned char* o = A.u8_out + pix;
                o[0] = ncr_to_u8(dr[p]); o[1] = ncr_to_u8(dg[p]); o[2] = ncr_to_u8(db[p]);
            }
        }
    }
    return staged_next;
}

// PREFETCH = false: claim a region, read its header and list, composite it (the in-region command prefetch is the only
// look-ahead) — the variant for long lists, where the per-region latency chain is amortised over many commands.
// PREFETCH = true: the variant for very short lists (a few full-screen commands per region), where that chain
// (claim -> header -> list -> first command, four dependent round trips to L2) is most of a region's time: regions are
// assigned statically and the rest of the chain is software-pipelined over regions, each link issued one whole region before
// its result is needed.  (A dynamic claim carried across a region was measured first: its result register is spilled right
// after the atomic — the warp waits there — and pinning it any other way costs what it saves; profiles/README.md.)
template <bool ALPHA, bool COUNT, bool PREFETCH>
__global__ void __launch_bounds__(NCR_COMPOSITE_THREADS, NCR_COMPOSITE_MIN_CTAS) ncr_composite(NcrFlushArgs A) {
    // u8 / 255.0 (CreateTextureUInt8, cpp:350), IEEE division on both sides.  32 copies, copy c of entry k at [k*32 + c]:
    // a lane only ever reads its own copy, so lookups never collide on a bank.
    double* s_lut = (double*)ncr_smem;
    // Per-warp double-buffered command slot: the next command of the list is fetched while the current one is applied.
    NcrCmd (*s_cmd_all)[2] = (NcrCmd (*)[2])(ncr_smem + NCR_LUT_BYTES);
    // One division per entry (256 per CTA, not 256 x copies: the init is ~10 % of a low-overdraw launch otherwise), then
    // each value is replicated by its thread.
    for (int k = threadIdx.x; k < 256; k += NCR_COMPOSITE_THREADS) {
        const double v = DIV((double)k, 255.0);
#pragma unroll 8
        for (int c = 0; c < NCR_LUT_COPIES; ++c) s_lut[k * NCR_LUT_COPIES + ((c + k) & (NCR_LUT_COPIES - 1))] = v;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    NcrCmd (*s_cmd)[2] = s_cmd_all + warp;
    const int n_tasks = A.d.tiles_x * A.d.tiles_y * NCR_TASKS_PER_TILE;
    const uint2* __restrict__ hdr = (const uint2*)A.fine_off;   // per region: {offset, count}
    unsigned long long n_applied = 0;
    int slot = 0;

    const uint32_t n = (uint32_t)n_tasks;
    const uint32_t n_warps = gridDim.x * (NCR_COMPOSITE_THREADS / 32);
    const uint32_t gwarp = blockIdx.x * (NCR_COMPOSITE_THREADS / 32) + warp;
    if (!PREFETCH) {
        // Dynamic scheduling: a warp's first region is its own index (no claim: 1,776 simultaneous atomics on one address cost
        // microseconds at kernel start); every later one is claimed from the global counter, which counts from n_warps.
        uint32_t task = gwarp;
        while (task < n) {
            const uint2 h = __ldg(hdr + task);
            uint32_t ents = 0;
            if ((uint32_t)lane < h.y) ents = __ldg(A.fine_list + h.x + lane);   // first 32 entries, one coalesced load
            run_region<ALPHA, COUNT>(A, s_cmd, (int)task, h.x, h.y, ents, slot, false, false, 0u, s_lut, n_applied);
            if (lane == 0) task = atomicAdd(&A.cursors[5], 1u) + n_warps;
            task = __shfl_sync(FULL, task, 0);
        }
    } else {
        // Static round-robin (region r -> warp r mod n_warps; neighbouring regions go to different SMs, so a hot spot of the
        // frame is spread over the machine) and a software pipeline over regions, one stage per region's worth of work between
        // a load's issue and its first use:
        //   iteration i (compositing region R[i]) issues  header(R[i+2]),  entries(R[i+1])
        //   and uses, in its last command, entries(R[i+1]) to fetch that region's first command.
        // No claim at all: this variant only runs frames whose lists are short (host: A.prefetch), where regions cost about the same.
        uint32_t t0 = gwarp, t1 = gwarp + n_warps, t2 = gwarp + 2 * n_warps;
        uint2 h0 = make_uint2(0, 0), h1 = make_uint2(0, 0);
        if (t0 < n) h0 = __ldg(hdr + t0);
        if (t1 < n) h1 = __ldg(hdr + t1);
        uint32_t e0 = 0;
        if ((uint32_t)lane < h0.y) e0 = __ldg(A.fine_list + h0.x + lane);
        bool have0 = false;
#ifdef NCR_TMA_IDENT
        // experiment X5: the background texel box of region R[i+1] is staged by the TMA unit while region R[i] is composited
        unsigned char* tma_base = ncr_smem + NCR_TMA_OFFSET + warp * NCR_TMA_BYTES_PER_WARP;
        const uint32_t mbar0 = smem_u32(tma_base + 1024);
        if (lane == 0) { tma_mbar_init(mbar0); tma_mbar_init(mbar0 + 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
        uint32_t parity = 0, inflight = 0, it = 0;   // bit b: state of buffer b
        {
            int u0, v0;
            if (t0 < n && tma_region_ok(A, (int)t0, u0, v0)) {
                if (lane == 0) tma_load_box(smem_u32(tma_base), A.tma_map, u0, v0, mbar0);
                inflight |= 1u;
            }
        }
#endif
        while (t0 < n) {
            uint2 h2 = make_uint2(0, 0);
            if (t2 < n) h2 = __ldg(hdr + t2);                              // R[i+2]: first used at the top of the next iteration
            uint32_t e1 = 0;
            if ((uint32_t)lane < h1.y) e1 = __ldg(A.fine_list + h1.x + lane);   // R[i+1]: first used in this region's last command
#ifdef NCR_TMA_IDENT
            const uint32_t b = it & 1u, nb = b ^ 1u;
            {   // stage R[i+1]'s box into the other buffer (first drain a box nobody consumed, to keep the barrier's phase)
                int u0, v0;
                if (inflight >> nb & 1u) { tma_wait(mbar0 + 8 * nb, parity >> nb & 1u); parity ^= 1u << nb; inflight &= ~(1u << nb); }
                if (t1 < n && tma_region_ok(A, (int)t1, u0, v0)) {
                    if (lane == 0) tma_load_box(smem_u32(tma_base + 512 * nb), A.tma_map, u0, v0, mbar0 + 8 * nb);
                    inflight |= 1u << nb;
                }
            }
            bool used = false;
            const bool have_box = (inflight >> b & 1u) != 0;
            have0 = run_region<ALPHA, COUNT>(A, s_cmd, (int)t0, h0.x, h0.y, e0, slot, have0, h1.y != 0, e1, s_lut, n_applied,
                                             have_box ? (const uint32_t*)(tma_base + 512 * b) : nullptr, mbar0 + 8 * b, parity >> b & 1u, &used);
            if (used) { parity ^= 1u << b; inflight &= ~(1u << b); }
            __syncwarp();   // every lane has read its texels before the buffer is reused two iterations later
            ++it;
#else
            have0 = run_region<ALPHA, COUNT>(A, s_cmd, (int)t0, h0.x, h0.y, e0, slot, have0, h1.y != 0, e1, s_lut, n_applied);
#endif
            t0 = t1; t1 = t2; t2 += n_warps;
            h0 = h1; h1 = h2;
            e0 = e1;
        }
    }

    if (COUNT) {
        for (int s = 16; s > 0; s >>= 1) n_applied += __shfl_down_sync(FULL, n_applied, s);
        if (lane == 0 && n_applied) atomicAdd((unsigned long long*)(A.cursors + 2), n_applied);
    }
}

// One grid size per (device, kernel variant): every resident CTA slot, one wave.
int g_grid[64][8];

template <bool ALPHA, bool COUNT, bool PREFETCH>
void launch(const NcrFlushArgs& A, cudaStream_t s, int variant) {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (g_grid[dev][variant] == 0) {
        int sms = 148, per_sm = 1;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(ncr_composite<ALPHA, COUNT, PREFETCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, NCR_SMEM_BYTES);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ncr_composite<ALPHA, COUNT, PREFETCH>, NCR_COMPOSITE_THREADS, NCR_SMEM_BYTES);
        g_grid[dev][variant] = sms * (per_sm > 0 ? per_sm : 1);   // persistent: every resident CTA slot, one wave
    }
    const int n_tasks = A.d.tiles_x * A.d.tiles_y * NCR_TASKS_PER_TILE;
    const int warps = NCR_COMPOSITE_THREADS / 32;
    int grid = g_grid[dev][variant];
    if (grid * warps > n_tasks) grid = (n_tasks + warps - 1) / warps;
    if (grid < 1) grid = 1;
    ncr_composite<ALPHA, COUNT, PREFETCH><<<grid, NCR_COMPOSITE_THREADS, NCR_SMEM_BYTES, s>>>(A);
}

}   // namespace

extern "C" void ncr_launch_composite(const NcrFlushArgs* A, cudaStream_t s) {
    const bool alpha = A->d.ipp == 4, count = A->count_pixels != 0, pre = A->prefetch != 0;
    if (alpha) {
        if (count) { if (pre) launch<true, true, true>(*A, s, 0); else launch<true, true, false>(*A, s, 1); }
        else       { if (pre) launch<true, false, true>(*A, s, 2); else launch<true, false, false>(*A, s, 3); }
    } else {
        if (count) { if (pre) launch<false, true, true>(*A, s, 4); else launch<false, true, false>(*A, s, 5); }
        else       { if (pre) launch<false, false, true>(*A, s, 6); else launch<false, false, false>(*A, s, 7); }
    }
}
