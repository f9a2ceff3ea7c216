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

struct NcrFramePool {
    long width = 0, height = 0;
    int alpha = 0;
    std::vector<RenderContext*> ctx;
    std::vector<unsigned char*> buf;   // pinned, two per worker (render into one while the other waits for its turn at the sink)
    long u8_bytes = 0, yuv_bytes = 0;
};

extern "C" {

// Contexts, their device buffers and the pinned frame buffers are created once and reused by every render call
// (cudaMalloc / cudaFree synchronise the device: they must not sit inside a render).
NcrFramePool* NcrCreateFramePoolOnDevices(long width, long height, int alpha, int n_workers, const int* devices, int n_devices) {
    if (width <= 0 || height <= 0) return nullptr;
    n_workers = std::max(1, std::min(n_workers, 256));
    NcrFramePool* p = new NcrFramePool();
    p->width = width; p->height = height; p->alpha = alpha;
    for (int k = 0; k < n_workers; ++k) {
        // worker k -> device devices[k % n_devices]: consecutive frames go to different GPUs, so the in-order sink drains them evenly
        RenderContext* c = (devices && n_devices > 0) ? NcrCreateRenderContextOnDevice(width, height, alpha != 0, devices[k % n_devices])
                                                      : CreateRenderContext(width, height, alpha != 0);
        if (!c) break;
        p->u8_bytes = GetBufferSize(c);
        p->yuv_bytes = NcrYUV420PSize(c);
        const unsigned long long fb = (unsigned long long)std::max(p->u8_bytes, p->yuv_bytes);
        unsigned char* b0 = (unsigned char*)NcrAllocHost(fb);
        unsigned char* b1 = b0 ? (unsigned char*)NcrAllocHost(fb) : nullptr;
        if (!b1) { if (b0) NcrFreeHost(b0); DestroyRenderContext(c); break; }
        p->ctx.push_back(c);
        p->buf.push_back(b0);
        p->buf.push_back(b1);
    }
    if (p->ctx.empty()) { delete p; return nullptr; }
    return p;
}

NcrFramePool* NcrCreateFramePool(long width, long height, int alpha, int n_workers) {
    return NcrCreateFramePoolOnDevices(width, height, alpha, std::min(n_workers, 64), nullptr, 0);
}

void NcrDestroyFramePool(NcrFramePool* p) {
    if (!p) return;
    for (unsigned char* b : p->buf) NcrFreeHost(b);
    for (RenderContext* c : p->ctx) DestroyRenderContext(c);
    delete p;
}

int NcrFramePoolWorkers(NcrFramePool* p) { return p ? (int)p->ctx.size() : 0; }

long NcrFramePoolRender(NcrFramePool* p, const void* const* traces, const long* trace_bytes, long n_frames,
                        Texture* const* textures, long n_textures, int present, NcrFrameSink sink, void* user) {
    if (n_frames <= 0) return 0;
    if (!p || !traces || !trace_bytes || (present != 0 && present != 1)) return -1;
    for (long f = 0; f < n_frames; ++f)
        if (!traces[f] || trace_bytes[f] <= 0 || !frame_is_independent((const unsigned char*)traces[f], trace_bytes[f])) return -2;
    const int n_workers = (int)std::min<long>((long)p->ctx.size(), n_frames);
    const long bytes = present == 1 ? p->yuv_bytes : p->u8_bytes;

    // Frame f is rendered by worker f % n_workers into that worker's buffer (f / n_workers) % 2 and handed to a delivery
    // thread, which calls the sink strictly in frame order; the worker goes on with its next frame in its other buffer and
    // only waits when that buffer's previous frame (f - 2 n_workers) has not been delivered yet.  So one slow frame does
    // not stall the pool, and the sink never runs concurrently with itself.
    std::mutex m;
    std::condition_variable cv;
    long delivered = 0;                       // frames [0, delivered) have been through the sink
    std::vector<char> ready((size_t)n_frames, 0);
    bool failed = false;
    auto fail = [&]() {
        std::lock_guard<std::mutex> g(m);
        failed = true;
        cv.notify_all();
    };

    auto worker = [&](int k) {
        RenderContext* ctx = p->ctx[k];
        for (long f = k, j = 0; f < n_frames; f += n_workers, ++j) {
            unsigned char* buf = p->buf[2 * k + (j & 1)];
            {
                std::unique_lock<std::mutex> g(m);
                cv.wait(g, [&] { return failed || delivered > f - 2L * n_workers; });   // this buffer's previous frame is out
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
            std::lock_guard<std::mutex> g(m);
            ready[(size_t)f] = 1;
            cv.notify_all();
        }
    };
    auto deliver = [&]() {
        for (long f = 0; f < n_frames; ++f) {
            {
                std::unique_lock<std::mutex> g(m);
                cv.wait(g, [&] { return failed || ready[(size_t)f]; });
                if (failed) return;
            }
            if (sink) sink(user, f, p->buf[2 * (f % n_workers) + ((f / n_workers) & 1)], bytes);
            std::lock_guard<std::mutex> g(m);
            delivered = f + 1;
            cv.notify_all();
        }
    };

    std::vector<std::thread> pool;
    for (int k = 0; k < n_workers; ++k) pool.emplace_back(worker, k);
    std::thread out(deliver);
    for (auto& t : pool) t.join();
    out.join();
    return failed ? -1 : n_frames;
}

// One-shot convenience: pool for this call only.
long NcrRenderFrames(long width, long height, int alpha, const void* const* traces, const long* trace_bytes, long n_frames,
                     Texture* const* textures, long n_textures, int n_workers, int present, NcrFrameSink sink, void* user) {
    if (n_frames <= 0) return 0;
    if (!traces || !trace_bytes || width <= 0 || height <= 0 || (present != 0 && present != 1)) return -1;
    for (long f = 0; f < n_frames; ++f)
        if (!traces[f] || trace_bytes[f] <= 0 || !frame_is_independent((const unsigned char*)traces[f], trace_bytes[f])) return -2;
    NcrFramePool* p = NcrCreateFramePool(width, height, alpha, (int)std::min<long>(n_workers, n_frames));
    if (!p) return -1;
    const long rc = NcrFramePoolRender(p, traces, trace_bytes, n_frames, textures, n_textures, present, sink, user);
    NcrDestroyFramePool(p);
    return rc;
}

}   // extern "C"
