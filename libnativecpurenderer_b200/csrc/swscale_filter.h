// Host-side construction of the scaling filters libswscale would use for PutRendererContextFrame's call
//     sws_getContext(w, h, RGBA | RGB24, dw, dh, YUV420P, SWS_BILINEAR, 0, 0, 0)            (reference cpp:241-256)
// when the VideoCap's size differs from the canvas's.  libswscale is an un-vendored third-party dependency of the reference; this
// follows its published algorithm (libswscale/utils.c, initFilter, bilinear branch) and is pinned — through the kernels that
// consume these tables — bit-exactly against libswscale 9.1.100 (tests/test_parity_gpu.py, tests/golden/make_swscale_fixtures.py).
//
//   xInc = ((src << 16) + (dst >> 1)) / dst;  |xInc - 65536| < 10: one unit tap per output sample ("unscaled").
//   otherwise: size = 3 when enlarging, 1 + (2 src + dst - 1) / dst when shrinking (at most src - 2); tap j of output i sits at
//   source sample pos[i] + j and weighs max(0, 2^30 - |distance|), the distance shrunk by dst / src when shrinking; near-zero
//   ends are trimmed (cutoff 0.002 of the unit), the common size is rounded up to the SIMD alignment of the x86 scalers
//   (4 horizontal, 2 vertical), taps outside the image are folded onto the edge sample, and each row is normalised to `one`
//   (2^14 horizontal, 2^12 vertical) with error diffusion.  Source and destination sample positions coincide for every plane
//   of this conversion (both 128/256 in libswscale's units), so no phase term appears.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <vector>

struct NcrSwsFilter {
    int size = 0;                 // taps per output sample
    std::vector<int32_t> pos;     // [n] first source sample
    std::vector<int32_t> coef;    // [n * size]
};

inline int ncr_sws_log2(int64_t v) {
    int n = 0;
    while (v > 1) { v >>= 1; ++n; }
    return n;
}

inline NcrSwsFilter ncr_sws_make_filter(int src, int dst, int align, int64_t one) {
    NcrSwsFilter F;
    const int64_t x_inc = (((int64_t)src << 16) + (dst >> 1)) / dst;
    const int64_t unit = (int64_t)1 << (54 - (ncr_sws_log2(src / dst) < 8 ? ncr_sws_log2(src / dst) : 8));
    int taps;
    std::vector<int64_t> w;
    F.pos.resize((size_t)dst);
    if (llabs(x_inc - 0x10000) < 10) {
        taps = 1;
        w.assign((size_t)dst, unit);
        for (int i = 0; i < dst; ++i) F.pos[i] = i;
    } else {
        taps = x_inc <= (1 << 16) ? 3 : (int)(1 + (2 * (int64_t)src + dst - 1) / dst);
        if (taps > src - 2) taps = src - 2;
        if (taps < 1) taps = 1;
        w.assign((size_t)dst * taps, 0);
        int64_t centre = ((128 * x_inc) >> 7) - ((128 * (int64_t)0x10000) >> 7);
        for (int i = 0; i < dst; ++i) {
            int64_t at = (centre - (taps - 2) * ((int64_t)1 << 16)) / ((int64_t)1 << 17);   // C division: toward zero
            F.pos[i] = (int32_t)at;
            for (int j = 0; j < taps; ++j, ++at) {
                int64_t dist = llabs(at * ((int64_t)1 << 17) - centre) << 13;
                if (x_inc > (1 << 16)) dist = dist * dst / src;
                int64_t c = ((int64_t)1 << 30) - dist;
                if (c < 0) c = 0;
                w[(size_t)i * taps + j] = c * (unit >> 30);
            }
            centre += 2 * x_inc;
        }
    }
    // trim near-zero taps at both ends; the longest remaining row decides the size
    int longest = 0;
    for (int i = dst - 1; i >= 0; --i) {
        int64_t* r = &w[(size_t)i * taps];
        double seen = 0;
        for (int j = 0; j < taps; ++j) {
            seen += (double)llabs(r[0]);
            if (seen > 0.002 * (double)unit) break;
            if (i < dst - 1 && F.pos[i] >= F.pos[i + 1]) break;   // keep the positions monotone
            for (int k = 1; k < taps; ++k) r[k - 1] = r[k];
            r[taps - 1] = 0;
            F.pos[i]++;
        }
        int len = taps;
        seen = 0;
        for (int j = taps - 1; j > 0; --j) {
            seen += (double)llabs(r[j]);
            if (seen > 0.002 * (double)unit) break;
            --len;
        }
        if (len > longest) longest = len;
    }
    if (longest == 1 && align == 2) align = 1;
    F.size = (longest + (align - 1)) & ~(align - 1);
    F.coef.assign((size_t)dst * F.size, 0);
    std::vector<int64_t> row((size_t)F.size);
    for (int i = 0; i < dst; ++i) {
        for (int j = 0; j < F.size; ++j) row[j] = j < taps ? w[(size_t)i * taps + j] : 0;
        if (F.pos[i] < 0) {   // taps left of the image fold onto sample 0
            for (int j = 1; j < F.size; ++j) {
                const int left = j + F.pos[i] > 0 ? j + F.pos[i] : 0;
                row[left] += row[j];
                row[j] = 0;
            }
            F.pos[i] = 0;
        }
        if (F.pos[i] + F.size > src) {   // taps right of the image fold onto the last sample
            const int shift = F.pos[i] + (F.size - src < 0 ? F.size - src : 0);
            int64_t spill = 0;
            for (int j = F.size - 1; j >= 0; --j)
                if (F.pos[i] + j >= src) { spill += row[j]; row[j] = 0; }
            for (int j = F.size - 1; j >= 0; --j) row[j] = j < shift ? 0 : row[j - shift];
            F.pos[i] -= shift;
            row[src - 1 - F.pos[i]] += spill;
        }
        int64_t sum = 0, carry = 0;
        for (int j = 0; j < F.size; ++j) sum += row[j];
        sum = (sum + one / 2) / one;
        if (!sum) sum = 1;
        for (int j = 0; j < F.size; ++j) {
            const int64_t v = row[j] + carry;
            const int64_t q = (v >= 0 ? v + (sum >> 1) : v - (sum >> 1)) / sum;
            F.coef[(size_t)i * F.size + j] = (int32_t)q;
            carry = v - q * sum;
        }
    }
    return F;
}
