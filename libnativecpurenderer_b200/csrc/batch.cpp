// Frame-parallel batch render: the host-side piece the reference left unfinished
// (MultiThreadedVideoRenderContextPreparer, reference src/libNativeCPURendererPybind.py:302-367: "record N frames,
// render them on a block of contexts, deliver in order").  N recorded frames (trace format: ncr_trace.h) are rendered
// by a pool of worker threads, one RenderContext — one CUDA stream — each, so the recording/state machine of one frame,
// the H2D copy of another and the kernels / readback of a third overlap; frames are handed to the sink strictly in
// frame order.  Built on the public C ABI of this library only.
//
// Frames must be independent: the first drawing call of every frame has to be SetColor, which overwrites the
// whole canvas (reference src/milrenderer.py:866 does exactly that); a frame that is not is refused, because a worker's
// canvas holds frame f - n_workers, not frame f - 1.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/ncr_b200.h"
#include "ncr_trace.h"

namespace {

// True when the first record that touches pixels is a full-canvas SetColor.
bool frame_is_independent(const unsigned char* p, long bytes) {
    const unsigned char* end = p + bytes;
    while (p + sizeof(NcrTraceRec) <= end) {
        NcrTraceRec r;
        memcpy(&r, p, sizeof r);
        const unsigned char* next = p + sizeof r + (size_t)r.n * sizeof(double);
        if (next > end) return false;
        switch (r.op) {
            case NCR_T_SAVE: case NCR_T_RESTORE: case NCR_T_SET_TRANSFORM: case NCR_T_APPLY_TRANSFORM: case NCR_T_SCALE:
            case NCR_T_TRANSLATE: case NCR_T_ROTATE: case NCR_T_SET_CT: case NCR_T_APPLY_CT: case NCR_T_CLIP_SET:
            case NCR_T_CLIP_CLEAR: case NCR_T_SAMPLING:
                break;   // state only
            case NCR_T_SET_COLOR:
                return true;
            default:
                return false;
        }
        p = next;
    }
    return false;
}

void reset_state(RenderContext* ctx) {
    while (RestoreContextState(ctx)) {}          // empty the save stack (cpp:291-309 returns false when empty)
    SetTransform(ctx, 1, 0, 0, 1, 0, 0);
    SetColorTransform(ctx, 1, 1, 1, 1);
    NcrClearClipRect(ctx);
    NcrSetSampling(ctx, 0);
}

}   // namespace

extern "C" long NcrRenderFrames(long width, long height, int alpha, const void* const* traces, const long* trace_bytes,
                                long n_frames, Texture* const* textures, long n_textures, int n_workers, int present,
                                NcrFrameSink sink, void* user) {
    if (n_frames <= 0) return 0;
    if (!traces || !trace_bytes || width <= 0 || height <= 0 || (present != 0 && present != 1)) return -1;
    for (long f = 0; f < n_frames; ++f)
        if (!traces[f] || trace_bytes[f] <= 0 || !frame_is_independent((const unsigned char*)traces[f], trace_bytes[f])) return -2;
    n_workers = (int)std::max<long>(1, std::min<long>(std::min<long>(n_workers, 64), n_frames));

    std::mutex m;
    std::condition_variable cv;
    long next = 0;
    bool failed = false;
    auto fail = [&]() {
        std::lock_guard<std::mutex> g(m);
        failed = true;
        cv.notify_all();
    };

    auto worker = [&](int k) {
        RenderContext* ctx = CreateRenderContext(width, height, alpha != 0);
        if (!ctx) { fail(); return; }
        const long bytes = present == 1 ? NcrYUV420PSize(ctx) : GetBufferSize(ctx);
        unsigned char* buf = (unsigned char*)NcrAllocHost((unsigned long long)bytes);   // pinned: the readback is a direct DMA
        if (!buf) { DestroyRenderContext(ctx); fail(); return; }
        for (long f = k; f < n_frames; f += n_workers) {
            {
                std::lock_guard<std::mutex> g(m);
                if (failed) break;
            }
            reset_state(ctx);
            bool ok = NcrSubmitTrace(ctx, traces[f], trace_bytes[f], textures, n_textures) >= 0;
            if (ok) {
                if (present == 1) ok = NcrGetBufferAsYUV420P(ctx, buf) == bytes;
                else GetBufferAsUInt8(ctx, buf);
            }
            if (ok && NcrFlush(ctx) != 0) ok = false;   // nothing pending: reports the context's sticky device-error state
            if (!ok) { fail(); break; }
            std::unique_lock<std::mutex> g(m);
            cv.wait(g, [&] { return next == f || failed; });
            if (failed) break;
            if (sink) sink(user, f, buf, bytes);   // in frame order, one at a time
            next = f + 1;
            cv.notify_all();
        }
        NcrFreeHost(buf);
        DestroyRenderContext(ctx);
    };

    std::vector<std::thread> pool;
    for (int k = 0; k < n_workers; ++k) pool.emplace_back(worker, k);
    for (auto& t : pool) t.join();
    return failed ? -1 : n_frames;
}
