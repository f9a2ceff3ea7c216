// Host-side trace replayer (built as lib/libncr_replay.so, separate from the product library).
//
// Trace replayer: dlopen()s any shared library exporting the reference C ABI (the unmodified reference
// build in oracle/_ref, the C restatement libncr_oracle.so, or the product) and feeds it a recorded
// command stream (format: libnativecpurenderer_b200/csrc/ncr_trace.h) through plain function pointers,
// so that a stream of 20,000 draws costs 20,000 C calls instead of 140,000 ctypes crossings
// (1.9 us per ctypes call).  Used by tests (parity on long streams) and by bench.py's
// cpu_baseline / --impl reference legs (single- and multi-threaded: one context per thread, which is
// how a single-threaded CPU renderer uses all host cores on independent frames).
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <chrono>
#include <thread>
#include <vector>

#include "ncr_trace.h"

namespace {

typedef void* H;   // opaque context / texture handle of the loaded library

struct Api {
    void* dl = nullptr;
    H (*CreateRenderContext)(long, long, bool) = nullptr;
    void (*DestroyRenderContext)(H) = nullptr;
    void (*SaveContextState)(H) = nullptr;
    bool (*RestoreContextState)(H) = nullptr;
    void (*SetTransform)(H, double, double, double, double, double, double) = nullptr;
    void (*ApplyTransform)(H, double, double, double, double, double, double) = nullptr;
    void (*Scale)(H, double, double) = nullptr;
    void (*Translate)(H, double, double) = nullptr;
    void (*Rotate)(H, double) = nullptr;
    void (*SetColorTransform)(H, double, double, double, double) = nullptr;
    void (*ApplyColorTransform)(H, double, double, double, double) = nullptr;
    void (*SetColor)(H, double, double, double, double) = nullptr;
    void (*FillColor)(H, double, double, double, double) = nullptr;
    void (*DrawTexture)(H, H, double, double, double, double) = nullptr;
    void (*DrawSplittedTexture)(H, H, double, double, double, double, double, double, double, double) = nullptr;
    void (*DrawRect)(H, double, double, double, double, double, double, double, double) = nullptr;
    void (*DrawLine)(H, double, double, double, double, double, double, double, double, double) = nullptr;
    void (*DrawCircle)(H, double, double, double, double, double, double, double) = nullptr;
    void (*DrawVerticalGrd)(H, double, double, double, double, double, double, double, double, double, double, double, double) = nullptr;
    bool (*SetPixel)(H, long, long, double, double, double, double) = nullptr;
    bool (*ApplyPixel)(H, long, long, double, double, double, double) = nullptr;   // optional
    void (*GetBufferAsUInt8)(H, unsigned char*) = nullptr;
    long (*GetBufferSize)(H) = nullptr;
    void* (*NcrAllocHost)(unsigned long long) = nullptr;   // optional (product): pinned readback destination
    void (*NcrFreeHost)(void*) = nullptr;
    // optional: the additive extension entry points (include/ncr_b200.h section 2)
    void (*NcrSetClipRect)(H, long, long, long, long) = nullptr;
    void (*NcrClearClipRect)(H) = nullptr;
    void (*NcrSetSampling)(H, int) = nullptr;
    void (*NcrFillPolygon)(H, const double*, long, double, double, double, double) = nullptr;
    void (*NcrDrawTexturePerspective)(H, H, const double*, double, double, double, double) = nullptr;
    long (*NcrGetBufferAsYUV420P)(H, unsigned char*) = nullptr;
    int present_mode = 0;   // 0: PRESENT reads back the u8 image (GetBufferAsUInt8); 1: the YUV 4:2:0 planes (video present path)
};

template <class F>
bool bind(void* dl, const char* name, F& fn, bool required = true) {
    fn = (F)dlsym(dl, name);
    if (!fn && required) fprintf(stderr, "ncr_replay: symbol %s missing\n", name);
    return fn != nullptr || !required;
}

// Executes the records of one trace on ctx.  Returns records executed, -1 on a malformed or unsupported stream.
long run_trace(const Api& a, H ctx, const unsigned char* p, long bytes, H const* tex, long ntex, unsigned char* frame,
               long* presents) {
    const unsigned char* end = p + bytes;
    long n = 0;
    while (p + sizeof(NcrTraceRec) <= end) {
        NcrTraceRec rec;
        memcpy(&rec, p, sizeof rec);
        p += sizeof rec;
        if ((size_t)(end - p) < (size_t)rec.n * sizeof(double)) return -1;
        const double* v = (const double*)p;
        p += (size_t)rec.n * sizeof(double);
        H t = nullptr;
        if (rec.op == NCR_T_DRAW_TEXTURE || rec.op == NCR_T_DRAW_SPLIT || rec.op == NCR_T_DRAW_PERSP) {
            long slot = (long)v[0];
            if (slot < 0 || slot >= ntex) return -1;
            t = tex[slot];
        }
        switch (rec.op) {
            case NCR_T_SAVE: a.SaveContextState(ctx); break;
            case NCR_T_RESTORE: a.RestoreContextState(ctx); break;
            case NCR_T_SET_TRANSFORM: a.SetTransform(ctx, v[0], v[1], v[2], v[3], v[4], v[5]); break;
            case NCR_T_APPLY_TRANSFORM: a.ApplyTransform(ctx, v[0], v[1], v[2], v[3], v[4], v[5]); break;
            case NCR_T_SCALE: a.Scale(ctx, v[0], v[1]); break;
            case NCR_T_TRANSLATE: a.Translate(ctx, v[0], v[1]); break;
            case NCR_T_ROTATE: a.Rotate(ctx, v[0]); break;
            case NCR_T_SET_CT: a.SetColorTransform(ctx, v[0], v[1], v[2], v[3]); break;
            case NCR_T_APPLY_CT: a.ApplyColorTransform(ctx, v[0], v[1], v[2], v[3]); break;
            case NCR_T_SET_COLOR: a.SetColor(ctx, v[0], v[1], v[2], v[3]); break;
            case NCR_T_FILL_COLOR: a.FillColor(ctx, v[0], v[1], v[2], v[3]); break;
            case NCR_T_DRAW_TEXTURE: a.DrawTexture(ctx, t, v[1], v[2], v[3], v[4]); break;
            case NCR_T_DRAW_SPLIT: a.DrawSplittedTexture(ctx, t, v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]); break;
            case NCR_T_DRAW_RECT: a.DrawRect(ctx, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]); break;
            case NCR_T_DRAW_LINE: a.DrawLine(ctx, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]); break;
            case NCR_T_DRAW_CIRCLE: a.DrawCircle(ctx, v[0], v[1], v[2], v[3], v[4], v[5], v[6]); break;
            case NCR_T_DRAW_GRD: a.DrawVerticalGrd(ctx, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11]); break;
            case NCR_T_SET_PIXEL: a.SetPixel(ctx, (long)v[0], (long)v[1], v[2], v[3], v[4], v[5]); break;
            case NCR_T_APPLY_PIXEL:
                if (!a.ApplyPixel) return -1;
                a.ApplyPixel(ctx, (long)v[0], (long)v[1], v[2], v[3], v[4], v[5]);
                break;
            case NCR_T_PRESENT:
                if (frame) {
                    if (a.present_mode == 1 && a.NcrGetBufferAsYUV420P) a.NcrGetBufferAsYUV420P(ctx, frame);
                    else a.GetBufferAsUInt8(ctx, frame);
                }
                if (presents) ++*presents;
                break;
            case NCR_T_CLIP_SET:
                if (!a.NcrSetClipRect) return -1;
                a.NcrSetClipRect(ctx, (long)v[0], (long)v[1], (long)v[2], (long)v[3]);
                break;
            case NCR_T_CLIP_CLEAR:
                if (!a.NcrClearClipRect) return -1;
                a.NcrClearClipRect(ctx);
                break;
            case NCR_T_SAMPLING:
                if (!a.NcrSetSampling) return -1;
                a.NcrSetSampling(ctx, (int)v[0]);
                break;
            case NCR_T_FILL_POLY:
                if (!a.NcrFillPolygon || rec.n < 6 || (rec.n & 1)) return -1;
                a.NcrFillPolygon(ctx, v + 4, (rec.n - 4) / 2, v[0], v[1], v[2], v[3]);
                break;
            case NCR_T_DRAW_PERSP:
                if (!a.NcrDrawTexturePerspective || rec.n != 14) return -1;
                a.NcrDrawTexturePerspective(ctx, t, v + 1, v[10], v[11], v[12], v[13]);
                break;
            default: return -1;   // unknown record, or an extension the target library does not export
        }
        ++n;
    }
    if (p != end) return -1;   // trailing partial record
    return n;
}

}   // namespace

extern "C" {

void* ncr_replay_open(const char* path) {
    void* dl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!dl) {
        fprintf(stderr, "ncr_replay: %s\n", dlerror());
        return nullptr;
    }
    Api* a = new Api();
    a->dl = dl;
    bool ok = bind(dl, "CreateRenderContext", a->CreateRenderContext) & bind(dl, "DestroyRenderContext", a->DestroyRenderContext) &
              bind(dl, "SaveContextState", a->SaveContextState) & bind(dl, "RestoreContextState", a->RestoreContextState) &
              bind(dl, "SetTransform", a->SetTransform) & bind(dl, "ApplyTransform", a->ApplyTransform) &
              bind(dl, "Scale", a->Scale) & bind(dl, "Translate", a->Translate) & bind(dl, "Rotate", a->Rotate) &
              bind(dl, "SetColorTransform", a->SetColorTransform) & bind(dl, "ApplyColorTransform", a->ApplyColorTransform) &
              bind(dl, "SetColor", a->SetColor) & bind(dl, "FillColor", a->FillColor) & bind(dl, "DrawTexture", a->DrawTexture) &
              bind(dl, "DrawSplittedTexture", a->DrawSplittedTexture) & bind(dl, "DrawRect", a->DrawRect) &
              bind(dl, "DrawLine", a->DrawLine) & bind(dl, "DrawCircle", a->DrawCircle) &
              bind(dl, "DrawVerticalGrd", a->DrawVerticalGrd) & bind(dl, "SetPixel", a->SetPixel) &
              bind(dl, "GetBufferAsUInt8", a->GetBufferAsUInt8) & bind(dl, "GetBufferSize", a->GetBufferSize);
    bind(dl, "ApplyPixel", a->ApplyPixel, false);
    bind(dl, "NcrAllocHost", a->NcrAllocHost, false);
    bind(dl, "NcrFreeHost", a->NcrFreeHost, false);
    bind(dl, "NcrSetClipRect", a->NcrSetClipRect, false);
    bind(dl, "NcrClearClipRect", a->NcrClearClipRect, false);
    bind(dl, "NcrSetSampling", a->NcrSetSampling, false);
    bind(dl, "NcrFillPolygon", a->NcrFillPolygon, false);
    bind(dl, "NcrDrawTexturePerspective", a->NcrDrawTexturePerspective, false);
    bind(dl, "NcrGetBufferAsYUV420P", a->NcrGetBufferAsYUV420P, false);
    if (!ok) {
        delete a;
        return nullptr;
    }
    return a;
}

// PRESENT records read back the u8 image (mode 0, the reference ABI) or the YUV 4:2:0 planes (mode 1, product only).
// Returns 0, or -1 when the library has no YUV entry point.
int ncr_replay_set_present(void* api, int mode) {
    Api* a = (Api*)api;
    if (mode == 1 && !a->NcrGetBufferAsYUV420P) return -1;
    a->present_mode = mode;
    return 0;
}

// Replays `trace` `repeats` times on an existing context.  frame (may be null) receives every PRESENT readback.
// Returns wall seconds, or a negative value on a malformed stream.
double ncr_replay_run(void* api, void* ctx, const void* trace, long bytes, void* const* textures, long n_textures,
                      unsigned char* frame, int repeats, long* presents_out) {
    const Api& a = *(const Api*)api;
    long presents = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < repeats; ++r)
        if (run_trace(a, ctx, (const unsigned char*)trace, bytes, textures, n_textures, frame, &presents) < 0) return -1.0;
    const auto t1 = std::chrono::steady_clock::now();
    if (presents_out) *presents_out = presents;
    return std::chrono::duration<double>(t1 - t0).count();
}

// n_threads workers, each with its own context of the given shape, each replaying the trace `repeats` times.
// warm_repeats untimed passes come first.  Returns wall seconds of the timed passes until the last worker finishes
// (contexts are created before the clock starts).
// frames_out (may be null): n_threads consecutive buffers of frame_stride bytes; receives each worker's LAST frame after the
// clock has stopped, so that the caller can check what the timed passes rendered.
double ncr_replay_run_threads_ex(void* api, int n_threads, long width, long height, int alpha, const void* trace, long bytes,
                                 void* const* textures, long n_textures, int repeats, int warm_repeats, unsigned char* frames_out,
                                 long frame_stride) {
    const Api& a = *(const Api*)api;
    std::vector<H> ctxs(n_threads);
    std::vector<std::vector<unsigned char>> pageable(n_threads);
    std::vector<unsigned char*> frames(n_threads, nullptr);
    const bool pinned = a.NcrAllocHost && a.NcrFreeHost;
    for (int k = 0; k < n_threads; ++k) {
        ctxs[k] = a.CreateRenderContext(width, height, alpha != 0);
        if (!ctxs[k]) return -1.0;
        const size_t bytes = (size_t)a.GetBufferSize(ctxs[k]);
        if (pinned) frames[k] = (unsigned char*)a.NcrAllocHost(bytes);
        if (!frames[k]) {
            pageable[k].resize(bytes);
            frames[k] = pageable[k].data();
        }
    }
    std::vector<long> rc(n_threads, 0);
    auto pass = [&](int reps) {
        std::vector<std::thread> pool;
        for (int k = 0; k < n_threads; ++k)
            pool.emplace_back([&, k]() {
                for (int r = 0; r < reps && rc[k] >= 0; ++r)
                    rc[k] = run_trace(a, ctxs[k], (const unsigned char*)trace, bytes, textures, n_textures, frames[k], nullptr);
            });
        for (auto& th : pool) th.join();
    };
    if (warm_repeats > 0) pass(warm_repeats);   // untimed: first-use allocations of each context
    const auto t0 = std::chrono::steady_clock::now();
    pass(repeats);
    const auto t1 = std::chrono::steady_clock::now();
    bool bad = false;
    for (int k = 0; k < n_threads; ++k) {
        if (frames_out) {
            long n = a.present_mode == 1 ? frame_stride : (long)a.GetBufferSize(ctxs[k]);
            if (n > frame_stride) n = frame_stride;
            memcpy(frames_out + (size_t)k * frame_stride, frames[k], (size_t)n);
        }
        a.DestroyRenderContext(ctxs[k]);
        if (pinned && pageable[k].empty()) a.NcrFreeHost(frames[k]);
        if (rc[k] < 0) bad = true;
    }
    if (bad) return -1.0;
    return std::chrono::duration<double>(t1 - t0).count();
}

double ncr_replay_run_threads(void* api, int n_threads, long width, long height, int alpha, const void* trace, long bytes,
                              void* const* textures, long n_textures, int repeats, int warm_repeats) {
    return ncr_replay_run_threads_ex(api, n_threads, width, height, alpha, trace, bytes, textures, n_textures, repeats,
                                     warm_repeats, nullptr, 0);
}

}   // extern "C"
