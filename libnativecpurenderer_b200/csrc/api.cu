// Host runtime + C ABI of the B200-native libNativeCPURenderer.so.
//
//   * contexts own an HBM-resident f64 canvas [h][w][ipp], a CUDA stream and a command recorder;
//   * textures are uploaded once and stay in HBM (RGBA8/RGB8 when born from bytes, f64 otherwise);
//   * state calls (transform / colour stacks) run on the host with the reference's expression trees;
//   * draw calls append one NcrCmd to pinned staging; any canvas read flushes: one H2D copy of the batch,
//     ncr_bin_coarse -> ncr_bin_fine -> ncr_composite on the context's stream, optional fused u8 image.
//
// There is no CPU rendering path in this library: without a usable CUDA device CreateRenderContext and the
// texture constructors print the reason, record it for NcrLastError() and return NULL.
#include <cuda.h>            // CUtensorMap types only; the driver entry point is resolved at run time (experiment X5)
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sched.h>
#include <time.h>
#include <chrono>
#include <atomic>
#include <memory>
#include <mutex>
#include <thread>
#include <algorithm>
#include <vector>

#include "../../include/ncr_b200.h"
#include "hiteffect.h"
#include "kernels.h"
#include "ncr_cmd.h"
#include "ncr_trace.h"
#include "state.h"
#include "swscale_filter.h"

// ------------------------------------------------------------------------------------------------
// device bookkeeping
// ------------------------------------------------------------------------------------------------
namespace {

bool env_flag_early(const char* name) {
    const char* v = getenv(name);
    return v && *v && atoi(v) != 0;
}

std::mutex g_mu;
int g_dev = -1;   // default device of this process: NCR_DEVICE, else LOCAL_RANK, else 0
int g_init = 0;   // 0 untried, 1 ok, -1 failed
char g_err[512] = "";
const int kMaxDevices = 64;
int g_n_devices = 0;
char g_devname[kMaxDevices][256];
std::atomic<int> g_dev_ok[kMaxDevices];   // 0 unchecked, 1 usable (sm_100), -1 not
std::atomic<unsigned long long> g_launches{0};
void* g_l2_scrub[kMaxDevices];
const size_t kL2ScrubBytes = 256u << 20;

void set_error(const char* what, const char* detail) {
    std::lock_guard<std::mutex> lk(g_mu);
    snprintf(g_err, sizeof(g_err), "%s: %s", what, detail);
    fprintf(stderr, "[libNativeCPURenderer/b200] %s\n", g_err);
}

bool ck(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    set_error(what, cudaGetErrorString(e));
    return false;
}
#define CK(call) ck((call), #call)

// Device bookkeeping.  A context (canvas, stream, recorder, device buffers) lives on ONE device, chosen at creation: the
// process default (NCR_DEVICE, else LOCAL_RANK, else 0 — what one-process-per-GPU launchers such as torchrun want) or an
// explicit one (NcrCreateRenderContextOnDevice / NcrCreateFramePoolOnDevices — what a single host process such as
// milrenderer.py needs to use a multi-GPU box).  Every entry point that touches CUDA makes its context's device current on
// the calling thread first; textures are replicated to a device the first time a context on it draws them.
std::mutex g_init_mu;
std::atomic<int> g_init_done{0};   // published copy of g_init for the lock-free fast path

bool init_runtime() {
    if (g_init_done.load(std::memory_order_acquire) != 0) return g_init_done.load(std::memory_order_acquire) == 1;
    std::lock_guard<std::mutex> init_lock(g_init_mu);   // first use from several threads at once (frame pool, tests)
    struct Publish { ~Publish() { g_init_done.store(g_init, std::memory_order_release); } } publish;
    if (g_init != 0) return g_init == 1;
    int want = 0;
    const char* env = getenv("NCR_DEVICE");
    if (!env || !*env) env = getenv("LOCAL_RANK");
    if (env && *env) want = atoi(env);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no usable CUDA device (this library has no CPU rendering path)",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        g_init = -1;
        return false;
    }
    g_n_devices = std::min(count, kMaxDevices);
    if (want < 0 || want >= g_n_devices) want = ((want % g_n_devices) + g_n_devices) % g_n_devices;
    g_dev = want;
    g_init = 1;
    return true;
}

// Makes `dev` current on the calling thread (checking once that it is an sm_100 part).
bool set_device(int dev) {
    if (!init_runtime()) return false;
    if (dev < 0 || dev >= g_n_devices) { set_error("device index", "out of range"); return false; }
    int ok = g_dev_ok[dev].load(std::memory_order_acquire);
    if (ok == 0) {
        std::lock_guard<std::mutex> init_lock(g_init_mu);
        ok = g_dev_ok[dev].load(std::memory_order_acquire);
        if (ok == 0) {
            cudaDeviceProp prop;
            ok = -1;
            if (CK(cudaSetDevice(dev)) && CK(cudaGetDeviceProperties(&prop, dev))) {
                if (prop.major < 10) {
                    char msg[400];
                    snprintf(msg, sizeof(msg), "%.200s is sm_%d%d; the kernels are built for sm_100a only", prop.name, prop.major, prop.minor);
                    set_error("unsupported device", msg);
                } else {
                    std::lock_guard<std::mutex> lk(g_mu);
                    snprintf(g_devname[dev], sizeof(g_devname[dev]), "%s", prop.name);
                    ok = 1;
                }
            }
            g_dev_ok[dev].store(ok, std::memory_order_release);
        }
    }
    if (ok != 1) return false;
    return cudaSetDevice(dev) == cudaSuccess;
}

bool use_device() { return init_runtime() && set_device(g_dev); }

struct DevMem {
    void* p = nullptr;
    size_t bytes = 0;
    ~DevMem() {
        if (p) cudaFree(p);
    }
};
typedef std::shared_ptr<DevMem> DevRef;

DevRef dev_alloc(size_t bytes) {
    DevRef m = std::make_shared<DevMem>();
    if (!CK(cudaMalloc(&m->p, bytes ? bytes : 16))) return nullptr;
    m->bytes = bytes;
    return m;
}

template <class T>
struct DevVec {
    T* p = nullptr;
    size_t cap = 0;
    bool reserve(size_t n) {
        if (n <= cap) return true;
        size_t want = cap ? cap : 1024;
        while (want < n) want *= 2;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        if (!CK(cudaMalloc((void**)&p, want * sizeof(T)))) return false;
        cap = want;
        return true;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

template <class T>
struct PinnedVec {
    T* p = nullptr;
    size_t cap = 0;
    bool reserve(size_t n, size_t keep) {
        if (n <= cap) return true;
        size_t want = cap ? cap : 4096;
        while (want < n) want *= 2;
        T* q = nullptr;
        if (!CK(cudaMallocHost((void**)&q, want * sizeof(T)))) return false;
        if (p) {
            memcpy(q, p, keep * sizeof(T));
            cudaFreeHost(p);
        }
        p = q;
        cap = want;
        return true;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

const uint32_t kMagicCtx = 0x4e435243u;   // "NCRC"
const uint32_t kMagicTex = 0x4e435254u;   // "NCRT"

}   // namespace

struct NcrTexture {
    uint32_t magic = kMagicTex;
    bool dead = false;
    i64 w = 0, h = 0;
    bool alpha = false;
    bool is_f64 = false;
    int dev = 0;                    // device that holds `buf`
    DevRef buf;                     // texels; null for an alias
    NcrContext* alias = nullptr;    // CreateTextureFromRenderContextShared: the canvas itself (cpp:382)
    std::mutex mu;                  // guards replicas, shadow and the alias snapshot (textures are shared between contexts / threads)
    std::vector<DevRef> replicas;   // [device]: copies of `buf` on other devices, made on first use there (slots never move)
    std::atomic<unsigned long long> replica_mask{0};   // bit d: replicas[d] is published
    std::vector<unsigned char> shadow;   // host copy of u8 texels, fetched on demand (hit-effect masks)
    DevRef tmap;                    // experiment X5: device copy of the texture's CUtensorMap (home device, RGBA8, w % 4 == 0)
    bool tmap_tried = false;
    // alias only: snapshot of the source canvas, reused while the source records nothing new (NcrContext::gen)
    DevRef snap;
    unsigned long long snap_gen = 0;
    int snap_dev = -1;
    i64 snap_w = 0, snap_h = 0;
};

struct NcrStaging {
    PinnedVec<NcrCmd> cmds;
    PinnedVec<NcrBox> boxes;
    PinnedVec<uint32_t> binbox;   // the box in 128-px bin coordinates, 4 x u8 (what ncr_bin_coarse reads)
    PinnedVec<double> aux;
    cudaEvent_t done = nullptr;
    bool inflight = false;
};

struct NcrContext {
    uint32_t magic = kMagicCtx;
    bool dead = false;
    int dev = 0;              // the device this context lives on
    unsigned long long gen = 1;   // bumped by everything that changes what the canvas will hold (recorded command, resize)
    i64 w = 0, h = 0;
    bool alpha = false;
    DevRef fb;
    NcrState st;
    std::vector<NcrState> stack;

    // recorder
    NcrStaging stg[2];
    int cur = 0;
    size_t n = 0, n_aux = 0;
    unsigned long long coarse_need = 0, fine_need = 0;
    bool load_fb = true;
    std::vector<DevRef> refs, last_refs;

    // device side
    DevVec<NcrCmd> d_cmds;
    DevVec<NcrBox> d_boxes;
    DevVec<uint32_t> d_binbox;
    DevVec<double> d_aux;
    DevVec<uint32_t> d_coarse, d_coarse_off, d_fine, d_fine_off, d_cursors;
    DevVec<unsigned char> d_u8;
    DevVec<unsigned char> d_yuv;
    bool u8_valid = false;
    bool yuv_valid = false;   // d_yuv holds the planes of the current canvas
    // Present-only flushes (GetBufferAsUInt8 / YUV) do not write the f64 canvas back: a video frame's canvas is overwritten
    // by the next frame's SetColor without ever being read (mil:866).  The canvas is then STALE: its true contents are
    // "`last` applied to what fb holds"; anything that needs them re-runs `last` with write_fb = 1 first (materialize()).
    bool fb_stale = false;
    int elide_block = 0;      // > 0: this many coming present flushes write the canvas (the caller reads canvases between presents)
    uint32_t* h_cursors = nullptr;   // pinned, 8 words
    bool cursors_pending = false;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_sync = nullptr;   // blocking-sync event (NCR_SYNC=block)
    bool ev_pending = false;
    NcrFlushArgs last;
    bool has_last = false;
    int stats_mode = 0;
    NcrStats stats;
    bool failed = false;   // sticky device error

    // present path, scaling branch: filter tables of the last (canvas size -> dw x dh) conversion, kept on the device
    NcrSwsPlan sws;
    bool sws_valid = false;
    DevVec<int32_t> d_sws_tables;
    DevVec<short> d_sws_mid;
    DevVec<unsigned char> d_sws_out;

    // experiment X5: the pending batch's TMA-stageable background draw (see NcrFlushArgs::tma_map)
    const void* tma_map = nullptr;
    int32_t tma_cmd = -1, tma_x = 0, tma_y = 0, tma_w = 0, tma_h = 0;

    // extensions
    bool clip_on = false;
    i64 clip_l = 0, clip_r = 0, clip_t = 0, clip_b = 0;
    int sampling = 0;
};

namespace {

inline NcrContext* live(RenderContext* c) {
    if (!c || c->magic != kMagicCtx || c->dead) return nullptr;
    return c;
}
inline NcrTexture* live(Texture* t) {
    if (!t || t->magic != kMagicTex || t->dead) return nullptr;
    return t;
}
inline int ipp_of(const NcrContext* c) { return c->alpha ? 4 : 3; }
inline bool use_ctx(const NcrContext* c) { return set_device(c->dev); }

void frame_dims(const NcrContext* c, NcrFrameDims* d) {
    d->w = (int32_t)c->w;
    d->h = (int32_t)c->h;
    d->ipp = ipp_of(c);
    d->tiles_x = (int32_t)((c->w + NCR_TILE - 1) / NCR_TILE);
    d->tiles_y = (int32_t)((c->h + NCR_TILE - 1) / NCR_TILE);
    d->bins_x = (d->tiles_x + NCR_COARSE - 1) / NCR_COARSE;
    d->bins_y = (d->tiles_y + NCR_COARSE - 1) / NCR_COARSE;
}

bool alloc_canvas(NcrContext* c, i64 w, i64 h) {
    if (w < 0 || h < 0 || w > 0x3fffffff || h > 0x3fffffff) {
        set_error("canvas size", "out of range");
        return false;
    }
    const size_t bytes = (size_t)w * (size_t)h * (c->alpha ? 4 : 3) * sizeof(double);
    DevRef fb = dev_alloc(bytes);
    if (!fb) return false;
    // The reference leaves new canvases uninitialised (cpp:15); the product defines them as zero.
    if (!CK(cudaMemsetAsync(fb->p, 0, bytes ? bytes : 16, c->stream))) return false;
    c->fb = fb;
    c->w = w;
    c->h = h;
    c->u8_valid = false;
    c->yuv_valid = false;
    c->has_last = false;
    c->fb_stale = false;
    return true;
}

// Collects the results of the previous flush (list sizes, pixel counter, overflow flag, event timings).
void harvest(NcrContext* c) {
    if (c->cursors_pending) {
        c->cursors_pending = false;
        c->stats.coarse_entries = c->h_cursors[0];
        c->stats.fine_entries = c->h_cursors[1];
        c->stats.blended_pixels = (unsigned long long)c->h_cursors[2] | ((unsigned long long)c->h_cursors[3] << 32);
        c->stats.interior_entries = c->h_cursors[6];
        if (c->h_cursors[4]) {
            set_error("tile-list overflow", c->h_cursors[4] == 1 ? "coarse list" : "fine list");
            c->failed = true;
        }
    }
    if (c->ev_pending) {
        c->ev_pending = false;
        cudaEventElapsedTime(&c->stats.ms_bin_coarse, c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&c->stats.ms_bin_fine, c->ev[1], c->ev[2]);
        cudaEventElapsedTime(&c->stats.ms_composite, c->ev[2], c->ev[3]);
        cudaEventElapsedTime(&c->stats.ms_total, c->ev[0], c->ev[3]);
    }
}

// How a host thread waits for its context's stream (NCR_SYNC=spin|yield|block|auto, default auto).
//   spin   cudaStreamSynchronize: the driver spins — lowest latency, one core burnt per waiting context;
//   yield  poll a stream event: sched_yield() between polls for ~150 us, then 80 us naps — the core goes to a sibling context
//          that is still recording its frame; lets a frame-parallel render run more contexts than it has cores (measured:
//          8 contexts on 4 cores reach the GPU-bound rate, 4 spinning ones do not);
//   block  event with cudaEventBlockingSync: the thread sleeps; costs ~40 % on sub-millisecond frames (wake-up latency);
//   auto   yield while this process has more live contexts than its share of the cores (affinity mask / LOCAL_WORLD_SIZE),
//          else spin.
std::atomic<int> g_live_contexts{0};

int cores_for_this_process() {
    static int n = 0;
    if (n == 0) {
        cpu_set_t set;
        int cores = 1;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
        int ranks = 1;
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
        n = std::max(1, cores / ranks);
    }
    return n;
}

int sync_mode() {
    static int mode = -1;   // 0 spin, 1 block, 2 yield, 3 auto
    if (mode < 0) {
        const char* e = getenv("NCR_SYNC");
        if (e && !strcmp(e, "spin")) mode = 0;
        else if (e && !strcmp(e, "block")) mode = 1;
        else if (e && !strcmp(e, "yield")) mode = 2;
        else mode = 3;
    }
    if (mode == 3) return g_live_contexts.load(std::memory_order_relaxed) > cores_for_this_process() ? 2 : 0;
    return mode;
}

bool sync_ctx(NcrContext* c) {
    bool ok;
    if (sync_mode() == 1 && c->ev_sync) {
        ok = CK(cudaEventRecord(c->ev_sync, c->stream)) && CK(cudaEventSynchronize(c->ev_sync));
    } else if (sync_mode() == 2 && c->ev_sync) {
        ok = CK(cudaEventRecord(c->ev_sync, c->stream));
        cudaError_t q = cudaSuccess;
        // Poll: yield the core for the first ~150 us (sub-millisecond frames finish inside this window with no wake-up
        // latency), then sleep between polls so that long waits leave the core to threads that are recording frames —
        // under CFS a yielding poller still takes its fair share of the core from them.
        const auto t0 = std::chrono::steady_clock::now();
        bool nap = false;
        while (ok && (q = cudaEventQuery(c->ev_sync)) == cudaErrorNotReady) {
            if (!nap) {
                sched_yield();
                nap = std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(150);
            } else {
                struct timespec ts = {0, 80 * 1000};
                nanosleep(&ts, nullptr);
            }
        }
        if (ok && q != cudaSuccess) ok = CK(q);
    } else {
        ok = CK(cudaStreamSynchronize(c->stream));
    }
    if (!ok) {
        c->failed = true;
        return false;
    }
    for (int k = 0; k < 2; ++k) c->stg[k].inflight = false;
    harvest(c);
    return !c->failed;
}

// Brings a stale canvas up to date: re-executes the resident last batch, this time writing the f64 canvas (no u8 / YUV output).
// The batch's inputs are intact by construction — fb was not written, the command / list buffers are only replaced by the
// next flush, and last_refs keeps its textures alive.
bool env_flag(const char* name, bool dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    return atoi(v) != 0;
}
const bool kElidePresentWriteback = env_flag("NCR_ELIDE", true);
int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}
const int kPrefetchMode = env_int("NCR_PREFETCH", -1);
const int kPrefetchMaxMean = env_int("NCR_PREFETCH_MAX_MEAN", 6);
const size_t kDirectBinMaxCmds = (size_t)env_int("NCR_DIRECT_BIN_MAX_CMDS", 96);   // up to this many commands per flush: no coarse pass

bool materialize(NcrContext* c) {
    if (!c->fb_stale) return true;
    c->fb_stale = false;
    if (!c->has_last) return true;
    NcrFlushArgs M = c->last;
    M.u8_out = nullptr;
    M.yuv_out = nullptr;
    M.write_fb = 1;
    M.count_pixels = 0;
    ncr_launch_flush(&M, c->stream, nullptr);
    g_launches += 2u + (M.coarse_list ? 1u : 0u);
    c->stats.kernel_launches += 2u + (M.coarse_list ? 1u : 0u);
    c->stats.materialized += 1;
    c->last.write_fb = 1;
    c->elide_block = 8;
    if (!CK(cudaGetLastError())) { c->failed = true; return false; }
    return true;
}

size_t yuv_bytes_of(const NcrContext* c) {
    return (size_t)c->w * c->h + 2 * (size_t)((c->w + 1) / 2) * (size_t)((c->h + 1) / 2);
}

// Submits the recorded batch.  want_u8: also produce the (iu8)(v*255) image in d_u8; want_yuv: the YUV 4:2:0 planes of that
// image in d_yuv (both fused into the composite's tile write-back).  With nothing pending, the outputs that are not
// already current are produced from the canvas by the standalone kernels.
bool flush(NcrContext* c, bool want_u8, bool want_yuv = false) {
    if (!use_ctx(c)) return false;
    const size_t n_elems = (size_t)c->w * c->h * ipp_of(c);
    if (c->n == 0) {
        const bool need_yuv = want_yuv && !c->yuv_valid && n_elems;
        const bool need_u8 = (want_u8 || need_yuv) && !c->u8_valid && n_elems;
        // the canvas itself is wanted (no output requested) or an output has to be converted from it
        if (c->fb_stale && ((!want_u8 && !want_yuv) || need_u8) && !materialize(c)) return false;
        if (need_u8) {
            if (!c->d_u8.reserve(n_elems)) return false;
            ncr_launch_convert_u8((const double*)c->fb->p, c->d_u8.p, n_elems, c->stream);
            g_launches += 1;
            c->stats.kernel_launches += 1;
            c->u8_valid = true;
        }
        if (need_yuv) {
            if (!c->d_yuv.reserve(yuv_bytes_of(c))) return false;
            ncr_launch_yuv420p(c->d_u8.p, c->d_yuv.p, (int)c->w, (int)c->h, ipp_of(c), c->stream);
            g_launches += 1;
            c->stats.kernel_launches += 1;
            c->yuv_valid = true;
        }
        return true;
    }
    if (c->cursors_pending || c->ev_pending) {
        // results of the previous flush live in pinned words that the next one overwrites
        if (!sync_ctx(c)) return false;
    }
    if (c->fb_stale) {
        if (c->load_fb) { if (!materialize(c)) return false; }   // this batch reads the canvas
        else c->fb_stale = false;                                // it starts with SetColor: the old contents are dead
    }
    NcrFlushArgs A;
    memset(&A, 0, sizeof(A));
    frame_dims(c, &A.d);
    const size_t n_tiles = (size_t)A.d.tiles_x * A.d.tiles_y, n_bins = (size_t)A.d.bins_x * A.d.bins_y;
    const size_t n_regions = n_tiles * NCR_REGIONS_PER_TILE;
    NcrStaging& S = c->stg[c->cur];
    bool ok = c->d_cmds.reserve(c->n) && c->d_boxes.reserve(c->n) && c->d_binbox.reserve(c->n) && c->d_aux.reserve(c->n_aux + 2) &&
              c->d_coarse.reserve(c->coarse_need + 1) && c->d_coarse_off.reserve(2 * n_bins) &&
              c->d_fine.reserve(c->fine_need + 1) && c->d_fine_off.reserve(2 * n_regions) && c->d_cursors.reserve(8);
    if (ok && (want_u8 || want_yuv)) ok = c->d_u8.reserve(n_elems);   // the YUV planes are converted from the u8 image
    if (ok && want_yuv) ok = c->d_yuv.reserve(yuv_bytes_of(c));
    if (!ok) { c->failed = true; return false; }
    ok = CK(cudaMemcpyAsync(c->d_cmds.p, S.cmds.p, c->n * sizeof(NcrCmd), cudaMemcpyHostToDevice, c->stream)) &&
         CK(cudaMemcpyAsync(c->d_boxes.p, S.boxes.p, c->n * sizeof(NcrBox), cudaMemcpyHostToDevice, c->stream)) &&
         CK(cudaMemcpyAsync(c->d_binbox.p, S.binbox.p, c->n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    if (ok && c->n_aux)
        ok = CK(cudaMemcpyAsync(c->d_aux.p, S.aux.p, c->n_aux * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (!ok) { c->failed = true; return false; }
    cudaEventRecord(S.done, c->stream);
    S.inflight = true;
    c->stats.h2d_bytes += c->n * (sizeof(NcrCmd) + sizeof(NcrBox) + sizeof(uint32_t)) + c->n_aux * sizeof(double);

    A.fb = (double*)c->fb->p;
    A.u8_out = (want_u8 || want_yuv) ? c->d_u8.p : nullptr;
    A.yuv_out = want_yuv ? c->d_yuv.p : nullptr;
    A.cmds = c->d_cmds.p;
    A.boxes = c->d_boxes.p;
    // bin coordinates fit in a byte up to 255 bins (32,640 px) per axis; larger canvases bin from the full boxes
    A.binboxes = (c->w <= 255 * NCR_TILE * NCR_COARSE && c->h <= 255 * NCR_TILE * NCR_COARSE) ? c->d_binbox.p : nullptr;
    A.aux = c->d_aux.p;
    A.n_cmds = (uint32_t)c->n;
    A.load_fb = c->load_fb ? 1u : 0u;
    A.coarse_list = c->n <= kDirectBinMaxCmds ? nullptr : c->d_coarse.p;
    A.coarse_off = c->d_coarse_off.p;
    A.fine_list = c->d_fine.p;
    A.fine_off = c->d_fine_off.p;
    A.cursors = c->d_cursors.p;
    A.coarse_cap = (uint32_t)c->d_coarse.cap;
    A.fine_cap = (uint32_t)c->d_fine.cap;
    A.tma_map = c->tma_map; A.tma_cmd = c->tma_cmd; A.tma_x = c->tma_x; A.tma_y = c->tma_y; A.tma_w = c->tma_w; A.tma_h = c->tma_h;
    c->tma_map = nullptr; c->tma_cmd = -1;
    A.count_pixels = (c->stats_mode & 1) ? 1u : 0u;
    // present-only flush: skip the canvas write-back unless this caller has been reading canvases between presents
    bool elide = (want_u8 || want_yuv) && kElidePresentWriteback;
    if (elide && c->elide_block > 0) { c->elide_block -= 1; elide = false; }
    A.write_fb = elide ? 0u : 1u;
    // Composite variant: with a handful of commands per region (chart / video frames) the per-region latency chain dominates
    // and the cross-region prefetch variant pays; with long lists it only costs registers.  Decided from the host-side bound of
    // the mean list length (NCR_PREFETCH=0/1 forces a variant; NCR_PREFETCH_MAX_MEAN moves the threshold).
    A.prefetch = kPrefetchMode >= 0 ? (uint32_t)kPrefetchMode
                                    : (c->fine_need <= (unsigned long long)kPrefetchMaxMean * n_regions ? 1u : 0u);
    const bool timed = (c->stats_mode & 2) != 0;
    ncr_launch_flush(&A, c->stream, timed ? c->ev : nullptr);
    const unsigned n_kernels = 2u + (A.coarse_list ? 1u : 0u) + (A.yuv_out ? 1u : 0u);
    g_launches += n_kernels;
    c->stats.kernel_launches += n_kernels;
    c->ev_pending = timed;
    cudaMemcpyAsync(c->h_cursors, c->d_cursors.p, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
    c->cursors_pending = true;
    if (!CK(cudaGetLastError())) { c->failed = true; return false; }

    c->stats.n_cmds = c->n;
    c->stats.flushes += 1;
    c->last = A;
    c->has_last = true;
    c->last_refs.swap(c->refs);
    c->refs.clear();
    c->n = 0;
    c->n_aux = 0;
    c->coarse_need = c->fine_need = 0;
    c->load_fb = true;
    c->u8_valid = want_u8 || want_yuv;
    c->yuv_valid = want_yuv;
    c->fb_stale = elide;
    c->cur ^= 1;
    return true;
}

// Room for one more command (+ extra aux doubles) in the current staging buffers.
bool reserve_cmd(NcrContext* c, size_t extra_aux) {
    NcrStaging& S = c->stg[c->cur];
    // Recording normally touches no CUDA API.  The two cases that do (waiting for a staging buffer still being uploaded,
    // growing pinned staging) first make the context's device current: a fresh worker thread's current device is 0, and
    // an allocation made there would create a primary context on GPU 0 for every rank of a multi-GPU run.
    if (S.inflight) {
        if (!use_ctx(c)) return false;
        cudaEventSynchronize(S.done);
        S.inflight = false;
    }
    if (c->n + 1 > S.cmds.cap) {
        if (!use_ctx(c)) return false;
        if (!S.cmds.reserve(c->n + 1, c->n) || !S.boxes.reserve(S.cmds.cap, c->n) || !S.binbox.reserve(S.cmds.cap, c->n)) return false;
    }
    if (c->n_aux + extra_aux > S.aux.cap) {
        if (!use_ctx(c)) return false;
        if (!S.aux.reserve(c->n_aux + extra_aux, c->n_aux)) return false;
    }
    return true;
}

// A batch is submitted early (recording continues in the other staging buffer while the GPU runs it; the NEXT submit first waits
// for this one, because the list-size words it reports live in one pinned block per context) once it holds this many
// commands or tile-list entries.  NCR_MAX_PENDING_CMDS / NCR_MAX_PENDING_ENTRIES override the defaults (tests use tiny
// values to exercise mid-stream submits).
size_t env_limit(const char* name, size_t dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    const long long n = atoll(v);
    return n > 0 ? (size_t)n : dflt;
}
const unsigned long long kMaxPendingEntries = env_limit("NCR_MAX_PENDING_ENTRIES", 48ull << 20);   // 4 B each
const size_t kMaxPendingCmds = std::min<size_t>(env_limit("NCR_MAX_PENDING_CMDS", 1u << 20), NCR_ENTRY_INDEX);   // a list entry holds the command index in its low bits

// Starts a command covering the pixel box [l,r) x [t,b) (already clamped to the canvas).  Returns nullptr when
// the box is empty — no pixel can be touched, so nothing is recorded.
NcrCmd* begin_cmd(NcrContext* c, uint32_t op, i64 l, i64 r, i64 t, i64 b, bool clip = true, size_t extra_aux = 0) {
    if (c->failed) return nullptr;
    uint32_t flags = 0;
    if (clip && c->clip_on) {
        l = std::max(l, c->clip_l); r = std::min(r, c->clip_r);
        t = std::max(t, c->clip_t); b = std::min(b, c->clip_b);
        flags |= NCR_F_CLIP;
    }
    if (l >= r || t >= b) return nullptr;
    if (c->n >= kMaxPendingCmds || c->fine_need > kMaxPendingEntries) {
        if (!flush(c, false)) return nullptr;
    }
    if (!reserve_cmd(c, extra_aux)) { c->failed = true; return nullptr; }
    NcrStaging& S = c->stg[c->cur];
    NcrCmd* cmd = &S.cmds.p[c->n];
    memset(cmd, 0, sizeof(NcrCmd));
    cmd->op = op;
    cmd->flags = flags;
    cmd->l = (int32_t)l; cmd->r = (int32_t)r; cmd->t = (int32_t)t; cmd->b = (int32_t)b;
    for (int k = 0; k < 4; ++k) cmd->ct[k] = c->st.ct[k];
    if (c->st.ct[0] == 1.0 && c->st.ct[1] == 1.0 && c->st.ct[2] == 1.0) cmd->flags |= NCR_F_CT_RGB_ONE;
    NcrBox& bx = S.boxes.p[c->n];
    bx.l = cmd->l; bx.r = cmd->r; bx.t = cmd->t; bx.b = cmd->b;
    // upper bound of the list entries this command can produce: one per 16x8 region its box touches
    const unsigned long long tx = ((r - 1) / NCR_REGION_W) - (l / NCR_REGION_W) + 1, ty = ((b - 1) / NCR_REGION_H) - (t / NCR_REGION_H) + 1;
    const i64 edge = NCR_TILE * NCR_COARSE;
    if (l < r && t < b) {   // first / last bin touched, per axis (l < x1 <=> l/edge <= bin; r > x0 <=> (r-1)/edge >= bin)
        S.binbox.p[c->n] = (uint32_t)std::min<i64>(l / edge, 255) | ((uint32_t)std::min<i64>((r - 1) / edge, 255) << 8) |
                           ((uint32_t)std::min<i64>(t / edge, 255) << 16) | ((uint32_t)std::min<i64>((b - 1) / edge, 255) << 24);
    } else {
        S.binbox.p[c->n] = 0x00010001u;   // first > last on both axes: touches no bin
    }
    const unsigned long long bxn = ((r - 1) / edge) - (l / edge) + 1, byn = ((b - 1) / edge) - (t / edge) + 1;
    c->fine_need += tx * ty;
    c->coarse_need += bxn * byn;
    c->n += 1;
    c->gen += 1;
    return cmd;
}

void put_inverse(NcrContext* c, NcrCmd* cmd) { ncr_inverse(c->st.m, cmd->inv); }

// Constant-colour primitives: everything ApplyPixel derives from the colour alone (cpp:525-536) is formed here, once,
// with the same IEEE operations: p[0..3] = colour * colorTransform, p[4..6] = rgb * a, p[7] = 1 - a.
void put_const_colour(NcrContext* c, NcrCmd* cmd, double r, double g, double b, double a) {
    r *= c->st.ct[0];
    g *= c->st.ct[1];
    b *= c->st.ct[2];
    a *= c->st.ct[3];
    cmd->p[0] = r; cmd->p[1] = g; cmd->p[2] = b; cmd->p[3] = a;
    cmd->p[4] = r * a; cmd->p[5] = g * a; cmd->p[6] = b * a;
    cmd->p[7] = 1 - a;
}

bool is_pow2(i64 v) { return v > 0 && (v & (v - 1)) == 0; }

// The texels of `tex` on device `dev`: the texture's own buffer, or a replica made (once) with a peer copy.  Returns a
// pointer to a DevRef that stays where it is for the texture's lifetime (no shared_ptr copy — hence no contended atomic —
// on the draw path).
const DevRef* texels_on_device(NcrTexture* tex, int dev) {
    if (dev == tex->dev) return &tex->buf;
    if (dev < 0 || dev >= (int)tex->replicas.size()) return nullptr;
    if (tex->replica_mask.load(std::memory_order_acquire) >> dev & 1ull) return &tex->replicas[dev];
    std::lock_guard<std::mutex> lk(tex->mu);
    if (tex->replica_mask.load(std::memory_order_acquire) >> dev & 1ull) return &tex->replicas[dev];
    if (!set_device(dev)) return nullptr;
    DevRef copy = dev_alloc(tex->buf->bytes);
    if (!copy) return nullptr;
    if (tex->buf->bytes && !CK(cudaMemcpyPeer(copy->p, dev, tex->buf->p, tex->dev, tex->buf->bytes))) return nullptr;
    tex->replicas[dev] = copy;
    tex->replica_mask.fetch_or(1ull << dev, std::memory_order_release);
    return &tex->replicas[dev];
}

// Experiment X5: a CUtensorMap (2-D, 4-byte elements, 16x8 box) over an RGBA8 texture, in device memory, so that a kernel built
// with NCR_TMA_IDENT can stage a region's texel box with cp.async.bulk.tensor.  Only when NCR_TMA=1.
const bool kTmaExperiment = env_flag_early("NCR_TMA");

const void* texture_tmap(NcrTexture* tex, int dev) {
    if (!kTmaExperiment || tex->alias || tex->is_f64 || !tex->alpha || dev != tex->dev || (tex->w & 3) || tex->w < 16 || tex->h < 8) return nullptr;
    std::lock_guard<std::mutex> lk(tex->mu);
    if (tex->tmap_tried) return tex->tmap ? tex->tmap->p : nullptr;
    tex->tmap_tried = true;
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) return nullptr;
    alignas(64) CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)tex->w, (cuuint64_t)tex->h};
    const cuuint64_t strides[1] = {(cuuint64_t)tex->w * 4};
    const cuuint32_t box[2] = {NCR_REGION_W, NCR_REGION_H}, estr[2] = {1, 1};
    if (((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, tex->buf->p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return nullptr;
    if (!set_device(dev)) return nullptr;
    DevRef m = dev_alloc(sizeof(map));
    if (!m || !CK(cudaMemcpy(m->p, &map, sizeof(map), cudaMemcpyHostToDevice))) return nullptr;
    tex->tmap = m;
    return m->p;
}

// Texture operand of a draw on device `dev`.  Aliases of a canvas (cpp:377-384) are resolved to a snapshot of that canvas
// as of this call, which is what an immediate-mode read of the shared buffer would have seen; the snapshot is kept and
// reused until the source context records something new (NcrContext::gen), so n draws of an unchanged shared canvas cost one
// copy, not n.  An alias makes the drawing thread flush and read the SOURCE context: like every other use of one context
// from two threads, that is only legal if the caller serialises them (threading).
// `keep` receives the reference that must outlive the recorded command.
bool bind_texture(int dev, NcrTexture* tex, const void** ptr, uint32_t* flags, const DevRef** keep) {
    if (tex->alias) {
        NcrContext* src = live(tex->alias);
        if (!src) return false;
        std::lock_guard<std::mutex> lk(tex->mu);
        if (!tex->snap || tex->snap_gen != src->gen || tex->snap_dev != dev || tex->snap_w != src->w || tex->snap_h != src->h) {
            if (!flush(src, false) || !sync_ctx(src)) return false;
            if (!set_device(dev)) return false;
            const size_t bytes = (size_t)src->w * src->h * ipp_of(src) * sizeof(double);
            DevRef snap = dev_alloc(bytes);
            if (!snap) return false;
            if (bytes && !CK(cudaMemcpy(snap->p, src->fb->p, bytes, cudaMemcpyDefault))) return false;
            tex->snap = snap;
            tex->snap_gen = src->gen;
            tex->snap_dev = dev;
            tex->snap_w = src->w;
            tex->snap_h = src->h;
        }
        *ptr = tex->snap->p;
        *keep = &tex->snap;   // the caller copies it into the batch's references right away (keep_ref)
        *flags = NCR_F_TEX_F64 | (src->alpha ? NCR_F_TEX_ALPHA : 0);
        return true;
    }
    *keep = texels_on_device(tex, dev);
    if (!*keep || !**keep) return false;
    *ptr = (**keep)->p;
    *flags = (tex->is_f64 ? NCR_F_TEX_F64 : 0) | (tex->alpha ? NCR_F_TEX_ALPHA : 0);
    return true;
}

void tex_dims(NcrTexture* tex, i64* w, i64* h) {
    if (tex->alias && live(tex->alias)) {
        *w = tex->alias->w;
        *h = tex->alias->h;
    } else {
        *w = tex->w;
        *h = tex->h;
    }
}

// Keeps the texels alive until the batch retires.  A frame uses few distinct textures: the recent ones are found by
// pointer comparison, so the shared_ptr is copied (one atomic increment) once per texture per batch, not per draw.
void keep_ref(NcrContext* c, const DevRef& r) {
    const size_t n = c->refs.size();
    for (size_t k = n > 16 ? n - 16 : 0; k < n; ++k)
        if (c->refs[k].get() == r.get()) return;
    c->refs.push_back(r);
}

NcrTexture* new_texture(i64 w, i64 h, bool alpha, bool is_f64, const void* host_data, int dev = -1) {
    if (!init_runtime()) return nullptr;
    if (dev < 0) dev = g_dev;
    if (!set_device(dev)) return nullptr;
    if (w < 0 || h < 0) { set_error("texture size", "negative"); return nullptr; }
    const size_t bytes = (size_t)w * h * (alpha ? 4 : 3) * (is_f64 ? 8 : 1);
    DevRef buf = dev_alloc(bytes);
    if (!buf) return nullptr;
    if (host_data && bytes && !CK(cudaMemcpy(buf->p, host_data, bytes, cudaMemcpyHostToDevice))) return nullptr;
    NcrTexture* t = new NcrTexture();
    t->w = w; t->h = h; t->alpha = alpha; t->is_f64 = is_f64;
    t->dev = dev;
    t->buf = buf;
    t->replicas.resize((size_t)g_n_devices);
    return t;
}

// Materialises a texture operand for setup-time consumers (resample, hit-effect mask) on the device that holds it
// (alias -> snapshot on the source canvas's device).  *dev_out receives that device.
bool texture_view(NcrTexture* tex, NcrCmd* view, DevRef* keep, int* dev_out) {
    memset(view, 0, sizeof(*view));
    const void* p = nullptr;
    uint32_t flags = 0;
    int dev = tex->dev;
    if (tex->alias) {
        NcrContext* src = live(tex->alias);
        if (!src) return false;
        dev = src->dev;
    }
    const DevRef* which = nullptr;
    if (!bind_texture(dev, tex, &p, &flags, &which)) return false;
    *keep = *which;
    *dev_out = dev;
    i64 w, h;
    tex_dims(tex, &w, &h);
    view->tex = p;
    view->flags = flags;
    view->tex_w = (int32_t)w;
    view->tex_h = (int32_t)h;
    return true;
}

}   // namespace

// ------------------------------------------------------------------------------------------------
// context lifecycle & readback
// ------------------------------------------------------------------------------------------------
extern "C" {

long GetBufferSize(RenderContext* ctx) {
    NcrContext* c = live(ctx);
    return c ? c->w * c->h * ipp_of(c) : 0;
}

static RenderContext* create_context(long width, long height, bool enableAlpha, int dev) {
    if (!init_runtime()) return nullptr;
    if (dev < 0) dev = g_dev;
    if (!set_device(dev)) return nullptr;
    NcrContext* c = new NcrContext();
    memset(&c->stats, 0, sizeof(c->stats));
    memset(&c->last, 0, sizeof(c->last));
    c->dev = dev;
    c->alpha = enableAlpha;
    ncr_state_reset(c->st);
    bool ok = CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int k = 0; ok && k < 4; ++k) ok = CK(cudaEventCreate(&c->ev[k]));
    ok = ok && CK(cudaEventCreate(&c->ev_t0)) && CK(cudaEventCreate(&c->ev_t1));
    ok = ok && CK(cudaEventCreateWithFlags(&c->ev_sync, cudaEventBlockingSync | cudaEventDisableTiming));
    for (int k = 0; ok && k < 2; ++k) ok = CK(cudaEventCreateWithFlags(&c->stg[k].done, cudaEventDisableTiming));
    ok = ok && CK(cudaMallocHost((void**)&c->h_cursors, 8 * sizeof(uint32_t)));
    ok = ok && alloc_canvas(c, width, height);
    if (!ok) {
        c->dead = true;
        return nullptr;
    }
    g_live_contexts.fetch_add(1, std::memory_order_relaxed);
    return c;
}

RenderContext* CreateRenderContext(long width, long height, bool enableAlpha) { return create_context(width, height, enableAlpha, -1); }

RenderContext* NcrCreateRenderContextOnDevice(long width, long height, bool enableAlpha, int device) {
    if (device < 0) { set_error("device index", "negative"); return nullptr; }
    return create_context(width, height, enableAlpha, device);
}

int NcrDeviceCount(void) { return init_runtime() ? g_n_devices : 0; }

int NcrContextDevice(RenderContext* ctx) {
    NcrContext* c = live(ctx);
    return c ? c->dev : -1;
}

void DestroyRenderContext(RenderContext* ctx) {
    NcrContext* c = live(ctx);
    if (!c || !use_ctx(c)) return;
    cudaStreamSynchronize(c->stream);
    c->dead = true;
    g_live_contexts.fetch_sub(1, std::memory_order_relaxed);
    c->fb.reset();
    c->refs.clear();
    c->last_refs.clear();
    c->d_cmds.release(); c->d_boxes.release(); c->d_binbox.release(); c->d_aux.release();
    c->d_coarse.release(); c->d_coarse_off.release(); c->d_fine.release(); c->d_fine_off.release();
    c->d_cursors.release(); c->d_u8.release(); c->d_yuv.release();
    c->d_sws_tables.release(); c->d_sws_mid.release(); c->d_sws_out.release();
    for (int k = 0; k < 2; ++k) {
        c->stg[k].cmds.release(); c->stg[k].boxes.release(); c->stg[k].binbox.release(); c->stg[k].aux.release();
        if (c->stg[k].done) cudaEventDestroy(c->stg[k].done);
    }
    for (int k = 0; k < 4; ++k) cudaEventDestroy(c->ev[k]);
    cudaEventDestroy(c->ev_t0);
    cudaEventDestroy(c->ev_t1);
    if (c->ev_sync) cudaEventDestroy(c->ev_sync);
    cudaFreeHost(c->h_cursors);
    cudaStreamDestroy(c->stream);
    // The small host object itself stays allocated (and marked dead) so that a stale handle is detected
    // instead of dereferencing freed memory; the reference never frees anything at all (cpp:33-37).
}

void ResizeRenderContext(RenderContext* ctx, long width, long height) {
    NcrContext* c = live(ctx);
    if (!c || !use_ctx(c)) return;
    // cpp:39-45: pixels are discarded, state is kept — pending draws can no longer be observed.
    cudaStreamSynchronize(c->stream);
    c->n = 0; c->n_aux = 0; c->coarse_need = c->fine_need = 0; c->load_fb = true;
    c->tma_map = nullptr; c->tma_cmd = -1;
    c->gen += 1;
    c->refs.clear();
    alloc_canvas(c, width, height);
}

void GetBuffer(RenderContext* ctx, double* buffer) {
    NcrContext* c = live(ctx);
    if (!c || !buffer) return;
    if (!flush(c, false)) return;
    const size_t bytes = (size_t)c->w * c->h * ipp_of(c) * sizeof(double);
    if (bytes && !CK(cudaMemcpyAsync(buffer, c->fb->p, bytes, cudaMemcpyDeviceToHost, c->stream))) c->failed = true;
    c->stats.d2h_bytes += bytes;
    sync_ctx(c);
}

void GetBufferAsUInt8(RenderContext* ctx, unsigned char* buffer) {
    NcrContext* c = live(ctx);
    if (!c || !buffer) return;
    if (!flush(c, true)) return;
    const size_t bytes = (size_t)c->w * c->h * ipp_of(c);
    if (bytes && !CK(cudaMemcpyAsync(buffer, c->d_u8.p, bytes, cudaMemcpyDeviceToHost, c->stream))) c->failed = true;
    c->stats.d2h_bytes += bytes;
    sync_ctx(c);
}

long NcrYUV420PSize(RenderContext* ctx) {
    NcrContext* c = live(ctx);
    if (!c || c->w <= 0 || c->h <= 0) return 0;
    return (long)(c->w * c->h + 2 * ((c->w + 1) / 2) * ((c->h + 1) / 2));
}

// Present path: the composite writes the YUV 4:2:0 planes of the (iu8)(v*255) image with its tiles (a standalone kernel does
// it when nothing is pending); only the planes, 1.5 B/px, are read back.
// libswscale-exact conversion of the current u8 image to dw x dh planes through the general (filter-table) path: the scaling
// branch of PutRendererContextFrame (cap size != canvas size, cpp:241-256), and same-size conversions of odd / tiny canvases,
// which the fused same-size kernel does not cover.  Leaves the planes in c->d_sws_out.
static bool sws_general(NcrContext* c, long dw, long dh) {
    if (!flush(c, true)) return false;   // composite (+ fused u8 image); the canvas write-back is skipped as for any present
    const int w = (int)c->w, h = (int)c->h;
    if (!c->sws_valid || c->sws.w != w || c->sws.h != h || c->sws.dw != dw || c->sws.dh != dh) {
        NcrSwsPlan P;
        memset(&P, 0, sizeof(P));
        P.w = w; P.h = h; P.dw = (int)dw; P.dh = (int)dh;
        P.cdw = (int)((dw + 1) / 2); P.cdh = (int)((dh + 1) / 2);
        P.half = ((dw >> 1) <= (w >> 1)) ? 1 : 0;              // libswscale: chroma input at half width unless the output needs more
        P.cw = P.half ? (w + 1) / 2 : w;
        const NcrSwsFilter hl = ncr_sws_make_filter(w, P.dw, 4, 1 << 14), vl = ncr_sws_make_filter(h, P.dh, 2, 1 << 12);
        const NcrSwsFilter hc = ncr_sws_make_filter(P.cw, P.cdw, 4, 1 << 14), vc = ncr_sws_make_filter(h, P.cdh, 2, 1 << 12);
        std::vector<int32_t> all;
        size_t off[8];
        const std::vector<int32_t>* parts[8] = {&hl.pos, &hl.coef, &vl.pos, &vl.coef, &hc.pos, &hc.coef, &vc.pos, &vc.coef};
        for (int k = 0; k < 8; ++k) { off[k] = all.size(); all.insert(all.end(), parts[k]->begin(), parts[k]->end()); }
        if (!sync_ctx(c) || !c->d_sws_tables.reserve(all.size())) return false;   // the previous tables may still be in use
        if (!CK(cudaMemcpyAsync(c->d_sws_tables.p, all.data(), all.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream)) ||
            !CK(cudaStreamSynchronize(c->stream)))
            return false;
        const int32_t* base = c->d_sws_tables.p;
        P.hl_pos = base + off[0]; P.hl_coef = base + off[1]; P.vl_pos = base + off[2]; P.vl_coef = base + off[3];
        P.hc_pos = base + off[4]; P.hc_coef = base + off[5]; P.vc_pos = base + off[6]; P.vc_coef = base + off[7];
        P.hl_size = hl.size; P.vl_size = vl.size; P.hc_size = hc.size; P.vc_size = vc.size;
        c->sws = P;
        c->sws_valid = true;
    }
    const NcrSwsPlan& P = c->sws;
    const size_t mid_y = (size_t)P.h * P.dw, mid_c = (size_t)P.h * P.cdw;
    if (!c->d_sws_mid.reserve(mid_y + 2 * mid_c) || !c->d_sws_out.reserve((size_t)P.dw * P.dh + 2 * (size_t)P.cdw * P.cdh)) return false;
    ncr_launch_sws_scaled(c->d_u8.p, ipp_of(c), &P, c->d_sws_mid.p, c->d_sws_mid.p + mid_y, c->d_sws_mid.p + mid_y + mid_c,
                          c->d_sws_out.p, c->stream);
    g_launches += 2;
    c->stats.kernel_launches += 2;
    return CK(cudaGetLastError());
}

long NcrGetBufferAsYUV420PScaled(RenderContext* ctx, long dst_w, long dst_h, unsigned char* out) {
    NcrContext* c = live(ctx);
    if (!c || !out || dst_w <= 0 || dst_h <= 0 || dst_w > 0x3fffffff || dst_h > 0x3fffffff) return -1;
    if (c->w <= 0 || c->h <= 0) return 0;
    if (dst_w == c->w && dst_h == c->h && !(c->w & 1) && !(c->h & 1) && c->w >= 8 && c->h >= 8) return NcrGetBufferAsYUV420P(ctx, out);
    if (!sws_general(c, dst_w, dst_h)) { c->failed = true; return -1; }
    const long bytes = dst_w * dst_h + 2 * ((dst_w + 1) / 2) * ((dst_h + 1) / 2);
    if (!CK(cudaMemcpyAsync(out, c->d_sws_out.p, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream))) { c->failed = true; return -1; }
    c->stats.d2h_bytes += (size_t)bytes;
    return sync_ctx(c) ? bytes : -1;
}

// Present path: the composite writes the (iu8)(v*255) image with its regions and ncr_yuv420p converts it (a standalone pass when
// nothing is pending); only the planes, 1.5 B/px, are read back.
long NcrGetBufferAsYUV420P(RenderContext* ctx, unsigned char* out) {
    NcrContext* c = live(ctx);
    if (!c || !out) return -1;
    const long bytes = NcrYUV420PSize(ctx);
    if (bytes <= 0) return 0;
    if ((c->w & 1) || (c->h & 1) || c->w < 8 || c->h < 8) {   // odd / tiny canvases: the general path (same tables as libswscale builds)
        if (!sws_general(c, c->w, c->h)) { c->failed = true; return -1; }
        if (!CK(cudaMemcpyAsync(out, c->d_sws_out.p, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream))) { c->failed = true; return -1; }
        c->stats.d2h_bytes += (size_t)bytes;
        return sync_ctx(c) ? bytes : -1;
    }
    if (!flush(c, false, true)) return -1;
    if (!CK(cudaGetLastError()) ||
        !CK(cudaMemcpyAsync(out, c->d_yuv.p, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream))) { c->failed = true; return -1; }
    c->stats.d2h_bytes += (size_t)bytes;
    return sync_ctx(c) ? bytes : -1;
}

void GetColor(RenderContext* ctx, double x, double y, double* out_r, double* out_g, double* out_b, double* out_a) {
    NcrContext* c = live(ctx);
    if (!c || c->w <= 0 || c->h <= 0) return;
    // cpp:664-670: clamp then truncate
    if (x < 0) x = 0;
    if (x >= c->w) x = c->w - 1;
    if (y < 0) y = 0;
    if (y >= c->h) y = c->h - 1;
    i64 ix = ncr_trunc_i64(x), iy = ncr_trunc_i64(y);
    ix = std::max((i64)0, std::min(c->w - 1, ix));
    iy = std::max((i64)0, std::min(c->h - 1, iy));
    if (!flush(c, false)) return;
    double px[4] = {0, 0, 0, 0};
    const int ipp = ipp_of(c);
    if (!CK(cudaMemcpyAsync(px, (double*)c->fb->p + (iy * c->w + ix) * ipp, ipp * sizeof(double), cudaMemcpyDeviceToHost,
                            c->stream)))
        return;
    if (!sync_ctx(c)) return;
    if (out_r) *out_r = px[0];
    if (out_g) *out_g = px[1];
    if (out_b) *out_b = px[2];
    if (c->alpha && out_a) *out_a = px[3];
}

// ------------------------------------------------------------------------------------------------
// state machine (host only)
// ------------------------------------------------------------------------------------------------
void SaveContextState(RenderContext* ctx) {
    NcrContext* c = live(ctx);
    if (c) c->stack.push_back(c->st);
}

bool RestoreContextState(RenderContext* ctx) {
    NcrContext* c = live(ctx);
    if (!c || c->stack.empty()) return false;
    c->st = c->stack.back();
    c->stack.pop_back();
    return true;
}

void SetTransform(RenderContext* ctx, double a, double b, double c_, double d, double e, double f) {
    NcrContext* c = live(ctx);
    if (!c) return;
    c->st.m[0] = a; c->st.m[1] = b; c->st.m[2] = c_; c->st.m[3] = d; c->st.m[4] = e; c->st.m[5] = f;
}

void ApplyTransform(RenderContext* ctx, double a, double b, double c_, double d, double e, double f) {
    NcrContext* c = live(ctx);
    if (c) ncr_apply_transform(c->st.m, a, b, c_, d, e, f);
}

void Scale(RenderContext* ctx, double sx, double sy) {
    NcrContext* c = live(ctx);
    if (c) ncr_apply_transform(c->st.m, sx, 0, 0, sy, 0, 0);
}

void Translate(RenderContext* ctx, double tx, double ty) {
    NcrContext* c = live(ctx);
    if (c) ncr_apply_transform(c->st.m, 1, 0, 0, 1, tx, ty);
}

void Rotate(RenderContext* ctx, double angle) {
    NcrContext* c = live(ctx);
    if (!c) return;
    const double s = sin(angle), co = cos(angle);   // host libm, as the reference (cpp:440-441)
    ncr_apply_transform(c->st.m, co, s, -s, co, 0, 0);
}

void TransformPoint(RenderContext* ctx, double x, double y, double* out_x, double* out_y) {
    NcrContext* c = live(ctx);
    if (c && out_x && out_y) ncr_xform_point(c->st.m, x, y, out_x, out_y);
}

void GetTransform(RenderContext* ctx, double out_matrix[6]) {
    NcrContext* c = live(ctx);
    if (c) memcpy(out_matrix, c->st.m, sizeof(c->st.m));
}

void GetInverseTransform(RenderContext* ctx, double out_matrix[6]) {
    NcrContext* c = live(ctx);
    if (c) ncr_inverse(c->st.m, out_matrix);
}

void SetColorTransform(RenderContext* ctx, double r, double g, double b, double a) {
    NcrContext* c = live(ctx);
    if (!c) return;
    c->st.ct[0] = r; c->st.ct[1] = g; c->st.ct[2] = b; c->st.ct[3] = a;
}

void ApplyColorTransform(RenderContext* ctx, double r, double g, double b, double a) {
    NcrContext* c = live(ctx);
    if (!c) return;
    c->st.ct[0] *= r; c->st.ct[1] *= g; c->st.ct[2] *= b; c->st.ct[3] *= a;
}

// ------------------------------------------------------------------------------------------------
// recorded pixel writes
// ------------------------------------------------------------------------------------------------
bool SetPixel(RenderContext* ctx, long x, long y, double r, double g, double b, double a) {
    NcrContext* c = live(ctx);
    if (!c) return false;
    if (x < 0 || x >= c->w || y < 0 || y >= c->h) return false;   // cpp:499-502
    NcrCmd* cmd = begin_cmd(c, NCR_OP_SET_PIXEL, x, x + 1, y, y + 1, false);
    if (!cmd) return true;
    cmd->p[0] = r; cmd->p[1] = g; cmd->p[2] = b; cmd->p[3] = a;
    if (!c->alpha) {
        // cpp:510 stores `a` at index+3 even on a 3-channel canvas: that element is the red of the next pixel
        // in row-major order (past-the-end for the last pixel: out of bounds in the reference, dropped here).
        i64 nx = x + 1, ny = y;
        if (nx >= c->w) { nx = 0; ny = y + 1; }
        if (ny < c->h) {
            NcrCmd* spill = begin_cmd(c, NCR_OP_SET_PIXEL, nx, nx + 1, ny, ny + 1, false);
            if (spill) { spill->p[0] = a; spill->flags |= NCR_F_ONLY_RED; }
        }
    }
    return true;
}

bool ApplyPixel(RenderContext* ctx, long x, long y, double r, double g, double b, double a) {
    NcrContext* c = live(ctx);
    if (!c) return false;
    if (x < 0 || x >= c->w || y < 0 || y >= c->h) return false;   // cpp:520-523
    NcrCmd* cmd = begin_cmd(c, NCR_OP_APPLY_PIXEL, x, x + 1, y, y + 1, false);
    if (cmd) put_const_colour(c, cmd, r, g, b, a);
    return true;
}

void SetColor(RenderContext* ctx, double r, double g, double b, double a) {
    NcrContext* c = live(ctx);
    if (!c || c->w <= 0 || c->h <= 0) return;
    // Every pixel is overwritten (cpp:643-657), so nothing recorded before this call can be observed any more
    // and the composite need not read the canvas back.
    c->n = 0; c->n_aux = 0; c->coarse_need = c->fine_need = 0;
    c->refs.clear();
    c->tma_map = nullptr; c->tma_cmd = -1;
    NcrCmd* cmd = begin_cmd(c, NCR_OP_SET_COLOR, 0, c->w, 0, c->h, false);
    if (!cmd) return;
    c->load_fb = false;
    const bool uniform = (r == g && g == b && b == a);   // cpp:647: std::fill of every element with r
    cmd->p[0] = r; cmd->p[1] = g; cmd->p[2] = b; cmd->p[3] = a;
    if (!c->alpha && !uniform) cmd->flags |= NCR_F_RGB_SPILL;   // SetPixel's index+3 store, see the kernel
}

void FillColor(RenderContext* ctx, double r, double g, double b, double a) {
    NcrContext* c = live(ctx);
    if (!c) return;
    NcrCmd* cmd = begin_cmd(c, NCR_OP_FILL_COLOR, 0, c->w, 0, c->h);
    if (cmd) put_const_colour(c, cmd, r, g, b, a);
}

// ------------------------------------------------------------------------------------------------
// recorded primitives
// ------------------------------------------------------------------------------------------------
static void fill_texture_fields(NcrContext* c, NcrCmd* cmd, const void* ptr, uint32_t tflags, const DevRef& keep, i64 tw, i64 th) {
    cmd->tex = ptr;
    cmd->flags |= tflags;
    if (c->sampling == 1) cmd->flags |= NCR_F_BILINEAR;
    else if ((tflags & NCR_F_TEX_ALPHA) && !(tflags & NCR_F_TEX_F64) && tw * th < 0x7fffffffL) {
        cmd->flags |= NCR_F_TEX_FAST;
        if (c->st.ct[3] >= 0.0 && c->st.ct[3] < 1.0) cmd->flags |= NCR_F_ALPHA_LT1;
        // the composite's hot path: DrawTexture, and DrawSplittedTexture on power-of-two textures (no division)
        if (cmd->op == NCR_OP_TEX || (cmd->op == NCR_OP_TEX_SPLIT && is_pow2(tw) && is_pow2(th))) cmd->flags |= NCR_F_FAST_AFFINE;
    }
    cmd->tex_w = (int32_t)tw;
    cmd->tex_h = (int32_t)th;
    keep_ref(c, keep);
}

void DrawTexture(RenderContext* ctx, Texture* tex_, double x, double y, double width, double height) {
    NcrContext* c = live(ctx);
    NcrTexture* tex = live(tex_);
    if (!c || !tex) return;
    if (width == 0 || height == 0) return;   // cpp:726
    i64 tw, th;
    tex_dims(tex, &tw, &th);
    const double scaleX = tw / width, scaleY = th / height;   // cpp:728-729
    const void* ptr; uint32_t tflags; const DevRef* keep = nullptr;
    if (ncr_is_no_transform(c->st.m)) {
        // cpp:741-742: for (i64 i = x; i < x + width; ++i) — the matrix is ignored, ApplyPixel clips.
        const i64 i0 = ncr_trunc_i64(x), j0 = ncr_trunc_i64(y);
        const double xw = x + width, yh = y + height;
        const double cr = ceil(xw), cb = ceil(yh);
        const i64 l = std::max((i64)0, std::min(c->w, i0)), t = std::max((i64)0, std::min(c->h, j0));
        const i64 r = !(cr > 0) ? 0 : (cr >= (double)c->w ? c->w : (i64)cr);
        const i64 b = !(cb > 0) ? 0 : (cb >= (double)c->h ? c->h : (i64)cb);
        if (l >= r || t >= b) return;
        if (!bind_texture(c->dev, tex, &ptr, &tflags, &keep)) return;
        NcrCmd* cmd = begin_cmd(c, NCR_OP_TEX_IDENT, l, r, t, b);
        if (!cmd) return;
        fill_texture_fields(c, cmd, ptr, tflags, *keep, tw, th);
        cmd->x = x; cmd->y = y; cmd->xw = xw; cmd->yh = yh;
        cmd->sx = scaleX; cmd->sy = scaleY;
        cmd->p[0] = (double)i0; cmd->p[1] = (double)j0;
        if (kTmaExperiment && c->tma_cmd < 0 && scaleX == 1.0 && scaleY == 1.0 && x == (double)i0 && y == (double)j0 &&
            (cmd->flags & NCR_F_TEX_FAST) && !(cmd->flags & NCR_F_CLIP) && fabs(x) < 1e9 && fabs(y) < 1e9) {
            if (const void* map = texture_tmap(tex, c->dev)) {
                c->tma_map = map; c->tma_cmd = (int32_t)(c->n - 1);
                c->tma_x = (int32_t)i0; c->tma_y = (int32_t)j0; c->tma_w = (int32_t)tw; c->tma_h = (int32_t)th;
                keep_ref(c, tex->tmap);
            }
        }
        return;
    }
    i64 l, r, t, b;
    ncr_border(c->st.m, x, y, width, height, c->w, c->h, &l, &r, &t, &b);
    if (l >= r || t >= b) return;
    if (!bind_texture(c->dev, tex, &ptr, &tflags, &keep)) return;
    NcrCmd* cmd = begin_cmd(c, NCR_OP_TEX, l, r, t, b);
    if (!cmd) return;
    fill_texture_fields(c, cmd, ptr, tflags, *keep, tw, th);
    put_inverse(c, cmd);
    cmd->x = x; cmd->y = y; cmd->xw = x + width; cmd->yh = y + height;
    cmd->sx = scaleX; cmd->sy = scaleY;
}

void DrawSplittedTexture(RenderContext* ctx, Texture* tex_, double x, double y, double width, double height, double uStart,
                         double uEnd, double vStart, double vEnd) {
    NcrContext* c = live(ctx);
    NcrTexture* tex = live(tex_);
    if (!c || !tex) return;
    if (width == 0 || height == 0) return;   // cpp:789
    i64 tw, th;
    tex_dims(tex, &tw, &th);
    i64 l, r, t, b;
    ncr_border(c->st.m, x, y, width, height, c->w, c->h, &l, &r, &t, &b);
    if (l >= r || t >= b) return;
    const void* ptr; uint32_t tflags; const DevRef* keep = nullptr;
    if (!bind_texture(c->dev, tex, &ptr, &tflags, &keep)) return;
    NcrCmd* cmd = begin_cmd(c, NCR_OP_TEX_SPLIT, l, r, t, b);
    if (!cmd) return;
    fill_texture_fields(c, cmd, ptr, tflags, *keep, tw, th);
    put_inverse(c, cmd);
    cmd->x = x; cmd->y = y; cmd->xw = x + width; cmd->yh = y + height;
    cmd->sx = tw / width; cmd->sy = th / height;   // cpp:793-794
    // cpp:812-813: u = (uStart + (uEnd - uStart) * u / tex->width) * tex->width
    cmd->p[0] = uStart; cmd->p[1] = uEnd - uStart;
    cmd->p[2] = vStart; cmd->p[3] = vEnd - vStart;
    cmd->p[4] = (double)tw; cmd->p[5] = (double)th;
    if (is_pow2(tw) && is_pow2(th)) {   // x / 2^k == x * 2^-k bit for bit
        cmd->flags |= NCR_F_SPLIT_POW2;
        cmd->p[6] = 1.0 / (double)tw; cmd->p[7] = 1.0 / (double)th;
    }
}

void DrawRect(RenderContext* ctx, double x, double y, double width, double height, double r_, double g, double b_, double a) {
    NcrContext* c = live(ctx);
    if (!c) return;
    if (width <= 0 || height <= 0) return;   // cpp:853
    i64 l, r, t, b;
    ncr_border(c->st.m, x, y, width, height, c->w, c->h, &l, &r, &t, &b);
    NcrCmd* cmd = begin_cmd(c, NCR_OP_RECT, l, r, t, b);
    if (!cmd) return;
    put_inverse(c, cmd);
    cmd->x = x; cmd->y = y; cmd->xw = x + width; cmd->yh = y + height;
    put_const_colour(c, cmd, r_, g, b_, a);
}

void DrawVerticalGrd(RenderContext* ctx, double x, double y, double width, double height, double top_r, double top_g,
                     double top_b, double top_a, double bottom_r, double bottom_g, double bottom_b, double bottom_a) {
    NcrContext* c = live(ctx);
    if (!c) return;
    if (width <= 0 || height <= 0) return;   // cpp:1291
    i64 l, r, t, b;
    ncr_border(c->st.m, x, y, width, height, c->w, c->h, &l, &r, &t, &b);
    NcrCmd* cmd = begin_cmd(c, NCR_OP_GRAD, l, r, t, b);
    if (!cmd) return;
    put_inverse(c, cmd);
    cmd->x = x; cmd->y = y; cmd->xw = x + width; cmd->yh = y + height;
    cmd->sy = height;   // cpp:1308: p = (invY - y) / height
    cmd->p[0] = top_r; cmd->p[1] = top_g; cmd->p[2] = top_b; cmd->p[3] = top_a;
    cmd->p[4] = bottom_r - top_r; cmd->p[5] = bottom_g - top_g;   // cpp:1309-1312
    cmd->p[6] = bottom_b - top_b; cmd->p[7] = bottom_a - top_a;
}

void DrawCircle(RenderContext* ctx, double x, double y, double radius, double r_, double g, double b_, double a) {
    NcrContext* c = live(ctx);
    if (!c) return;
    if (radius <= 0) return;   // cpp:926
    i64 l, r, t, b;
    ncr_border(c->st.m, x - radius, y - radius, 2 * radius, 2 * radius, c->w, c->h, &l, &r, &t, &b);   // cpp:932
    NcrCmd* cmd = begin_cmd(c, NCR_OP_CIRCLE, l, r, t, b);
    if (!cmd) return;
    put_inverse(c, cmd);
    cmd->x = x; cmd->y = y; cmd->sx = radius;
    put_const_colour(c, cmd, r_, g, b_, a);
}

// Polygon fill in inverse-mapped space (cpp:908-916).  The reference scans the whole canvas; the recorded box is
// a conservative cover of the forward-mapped polygon (full canvas whenever the cover cannot be trusted).
static void record_polygon(NcrContext* c, const double* pts, size_t n, double r_, double g, double b_, double a) {
    if (n == 0) return;
    const double* m = c->st.m;
    i64 l = 0, r = c->w, t = 0, b = c->h;
    const double det = m[0] * m[3] - m[1] * m[2];
    const double norm2 = m[0] * m[0] + m[1] * m[1] + m[2] * m[2] + m[3] * m[3];
    bool tight = det != 0 && isfinite(det) && isfinite(norm2) && isfinite(m[4]) && isfinite(m[5]);
    if (tight) {
        double minx = INFINITY, maxx = -INFINITY, miny = INFINITY, maxy = -INFINITY, mag = fabs(m[4]) + fabs(m[5]);
        for (size_t k = 0; k < n; ++k) {
            double fx, fy;
            ncr_xform_point(m, pts[2 * k], pts[2 * k + 1], &fx, &fy);
            if (!isfinite(fx) || !isfinite(fy)) { tight = false; break; }
            minx = std::min(minx, fx); maxx = std::max(maxx, fx);
            miny = std::min(miny, fy); maxy = std::max(maxy, fy);
            mag = std::max(mag, fabs(pts[2 * k]) * sqrt(norm2) + fabs(pts[2 * k + 1]) * sqrt(norm2));
        }
        // Rounding in inv / the per-pixel inverse map moves the boundary by about cond(M) * |coords| * 2^-52 px.
        const double cond = norm2 / fabs(det);
        const double slack = cond * (mag + (double)c->w + (double)c->h) * 1e-14;
        if (tight && slack < 0.25) {
            const double pad = 2.0;
            const double L = floor(minx - pad), R = ceil(maxx + pad) + 1, T = floor(miny - pad), B = ceil(maxy + pad) + 1;
            l = L <= 0 ? 0 : (L >= (double)c->w ? c->w : (i64)L);
            r = R <= 0 ? 0 : (R >= (double)c->w ? c->w : (i64)R);
            t = T <= 0 ? 0 : (T >= (double)c->h ? c->h : (i64)T);
            b = B <= 0 ? 0 : (B >= (double)c->h ? c->h : (i64)B);
        }
    }
    NcrCmd* cmd = begin_cmd(c, NCR_OP_POLY, l, r, t, b, true, 2 * n);
    if (!cmd) return;
    put_inverse(c, cmd);
    NcrStaging& S = c->stg[c->cur];
    cmd->aux_off = (uint32_t)c->n_aux;
    cmd->aux_n = (uint32_t)n;
    memcpy(S.aux.p + c->n_aux, pts, 2 * n * sizeof(double));
    c->n_aux += 2 * n;
    put_const_colour(c, cmd, r_, g, b_, a);
}

void DrawLine(RenderContext* ctx, double x1, double y1, double x2, double y2, double width, double r, double g, double b,
              double a) {
    NcrContext* c = live(ctx);
    if (!c) return;
    if (width <= 0) return;   // cpp:883
    // cpp:888-906: the stroke is the 4-gon endpoints -/+ unit normal * width/2
    const double dx = x2 - x1, dy = y2 - y1;
    const double len = sqrt(dx * dx + dy * dy);
    if (len == 0) return;
    const double ux = dx / len, uy = dy / len;
    const double vx = -uy, vy = ux;
    const double hw = width / 2;
    const double pts[8] = {x1 - vx * hw, y1 - vy * hw, x1 + vx * hw, y1 + vy * hw,
                           x2 + vx * hw, y2 + vy * hw, x2 - vx * hw, y2 - vy * hw};
    record_polygon(c, pts, 4, r, g, b, a);
}

// ------------------------------------------------------------------------------------------------
// textures
// ------------------------------------------------------------------------------------------------
Texture* CreateTexture(long width, long height, bool enableAlpha, double* buffer) {
    return new_texture(width, height, enableAlpha, true, buffer);
}

Texture* CreateTextureUInt8(long width, long height, bool enableAlpha, unsigned char* buffer) {
    // cpp:350 stores u8 / 255.0; the bytes stay bytes in HBM and are decoded with the same division in-kernel.
    return new_texture(width, height, enableAlpha, false, buffer);
}

void DestroyTexture(Texture* tex_) {
    NcrTexture* tex = live(tex_);
    if (!tex) return;
    // Recorded commands hold their own reference to the texels, so dropping ours is safe at any time.
    tex->dead = true;
    tex->buf.reset();
    tex->shadow.clear();
    tex->shadow.shrink_to_fit();
}

Texture* CreateTextureFromRenderContext(RenderContext* ctx) {
    NcrContext* c = live(ctx);
    if (!c) return nullptr;
    if (!flush(c, false)) return nullptr;
    NcrTexture* t = new_texture(c->w, c->h, c->alpha, true, nullptr, c->dev);
    if (!t) return nullptr;
    const size_t bytes = (size_t)c->w * c->h * ipp_of(c) * sizeof(double);
    if (bytes && !CK(cudaMemcpyAsync(t->buf->p, c->fb->p, bytes, cudaMemcpyDeviceToDevice, c->stream))) return nullptr;
    sync_ctx(c);
    return t;
}

Texture* CreateTextureFromRenderContextShared(RenderContext* ctx) {
    NcrContext* c = live(ctx);
    if (!c) return nullptr;
    NcrTexture* t = new NcrTexture();
    t->w = c->w; t->h = c->h; t->alpha = c->alpha; t->is_f64 = true;
    t->dev = c->dev;
    t->alias = c;
    return t;
}

long GetTextureWidth(Texture* tex_) {
    NcrTexture* tex = live(tex_);
    if (!tex) return 0;
    i64 w, h;
    tex_dims(tex, &w, &h);
    return w;
}

long GetTextureHeight(Texture* tex_) {
    NcrTexture* tex = live(tex_);
    if (!tex) return 0;
    i64 w, h;
    tex_dims(tex, &w, &h);
    return h;
}

bool GetTextureEnableAlpha(Texture* tex_) {
    NcrTexture* tex = live(tex_);
    return tex ? tex->alpha : false;
}

Texture* ResampleTexture(Texture* tex_, long width, long height) {
    NcrTexture* tex = live(tex_);
    if (!tex || !init_runtime()) return nullptr;
    NcrCmd view;
    DevRef keep;
    int dev = 0;
    if (!texture_view(tex, &view, &keep, &dev)) return nullptr;
    const bool f64 = (view.flags & NCR_F_TEX_F64) != 0, alpha = (view.flags & NCR_F_TEX_ALPHA) != 0;
    NcrTexture* out = new_texture(width, height, alpha, f64, nullptr, dev);   // leaves `dev` current
    if (!out) return nullptr;
    if (width > 0 && height > 0) {
        // The calling thread's own stream: it does not serialise with the legacy default stream or with any context's
        // (non-blocking) stream; the source texels were written by synchronous copies, so they are visible.
        ncr_launch_resample(&view, out->buf->p, (int)width, (int)height, cudaStreamPerThread);
        g_launches += 1;
        if (!CK(cudaGetLastError()) || !CK(cudaStreamSynchronize(cudaStreamPerThread))) return nullptr;
    }
    return out;
}

void GetMilthmHitEffectPixel(double seed, double t, double x, double y, double* a) {
    if (a) *a = ncr_hit_effect_alpha(seed, t, x, y);
}

Texture* CreateMilthmHitEffectTexture(Texture* mask_, double seed, double t, double r, double g, double b) {
    NcrTexture* mask = live(mask_);
    if (!mask || !init_runtime()) return nullptr;
    if (!mask->alpha) return nullptr;   // cpp:1418
    i64 w, h;
    tex_dims(mask, &w, &h);   // an alias follows its canvas through ResizeRenderContext
    const size_t n = (size_t)w * h;
    // The noise uses libm sin/atan2 (cpp:1339-1389) and is thresholded, so it is evaluated on the host to stay
    // bit-identical; only the mask's alpha plane is needed from the device (fetched once per mask).
    std::vector<double> mask_a(n);
    if (mask->alias || mask->is_f64) {
        NcrCmd view; DevRef keep;
        int vdev = 0;
        if (!texture_view(mask, &view, &keep, &vdev) || !set_device(vdev)) return nullptr;
        std::vector<double> all(n * 4);
        if (n && !CK(cudaMemcpy(all.data(), view.tex, n * 4 * sizeof(double), cudaMemcpyDeviceToHost))) return nullptr;
        for (size_t k = 0; k < n; ++k) mask_a[k] = all[4 * k + 3];
    } else {
        std::lock_guard<std::mutex> lk(mask->mu);   // one mask is shared by every effect texture built from it, possibly from several threads
        if (mask->shadow.size() != n * 4) {
            if (!set_device(mask->dev)) return nullptr;
            std::vector<unsigned char> fetched(n * 4);
            if (n && !CK(cudaMemcpy(fetched.data(), mask->buf->p, n * 4, cudaMemcpyDeviceToHost))) return nullptr;
            mask->shadow.swap(fetched);
        }
        for (size_t k = 0; k < n; ++k) mask_a[k] = mask->shadow[4 * k + 3] / 255.0;
    }
    // cpp:1426-1436 indexes mask and output as [i*height*4 + j*4] (i over width): linear element i*h + j.
    std::vector<double> out(n * 4);
    bool bytes_exact = true;
    // Texels are independent, so the rows are spread over the host cores (milrenderer builds 480 of these 512^2 textures at
    // start-up, mil:856-860: ~58 s on one core); every texel runs the same scalar libm code, so the result stays bit-identical.
    auto rows = [&](i64 i0, i64 i1) {
        for (i64 i = i0; i < i1; ++i) {
            for (i64 j = 0; j < h; ++j) {
                const double av = ncr_hit_effect_alpha(seed, t, (double)i / w, (double)j / h);
                const size_t k = (size_t)i * h + j;
                out[4 * k + 0] = r; out[4 * k + 1] = g; out[4 * k + 2] = b;
                out[4 * k + 3] = av * mask_a[k];
            }
        }
    };
    unsigned n_thr = std::thread::hardware_concurrency();
    if (const char* e = getenv("NCR_HOST_THREADS")) n_thr = (unsigned)atoi(e);
    n_thr = std::max(1u, std::min({n_thr, 32u, (unsigned)std::max<i64>(1, w / 8)}));
    if (n < (size_t)1 << 14) n_thr = 1;
    if (n_thr == 1) {
        rows(0, w);
    } else {
        std::vector<std::thread> pool;
        for (unsigned k = 0; k < n_thr; ++k) pool.emplace_back(rows, w * k / n_thr, w * (k + 1) / n_thr);
        for (auto& th : pool) th.join();
    }
    // Store as RGBA8 when every element is exactly some k/255.0 (true for milrenderer's call, pyb:45-47).
    unsigned char rgb8[3];
    const double rgbv[3] = {r, g, b};
    for (int ch = 0; ch < 3 && bytes_exact; ++ch) {
        const double s = rgbv[ch] * 255.0;
        const long k = lrint(s);
        if (k < 0 || k > 255 || (double)k / 255.0 != rgbv[ch]) bytes_exact = false;
        else rgb8[ch] = (unsigned char)k;
    }
    std::vector<unsigned char> out8;
    if (bytes_exact) {
        out8.resize(n * 4);
        for (size_t k = 0; k < n && bytes_exact; ++k) {
            const double av = out[4 * k + 3];
            const long q = lrint(av * 255.0);
            if (q < 0 || q > 255 || (double)q / 255.0 != av) { bytes_exact = false; break; }
            out8[4 * k + 0] = rgb8[0]; out8[4 * k + 1] = rgb8[1]; out8[4 * k + 2] = rgb8[2];
            out8[4 * k + 3] = (unsigned char)q;
        }
    }
    if (bytes_exact) return new_texture(w, h, true, false, out8.data());
    return new_texture(w, h, true, true, out.data());
}

long GetVersion(void) { return 1; }   // h:9

// ------------------------------------------------------------------------------------------------
// additive entry points
// ------------------------------------------------------------------------------------------------
int NcrFlush(RenderContext* ctx) {
    NcrContext* c = live(ctx);
    if (!c) return -1;
    if (!flush(c, false)) return -1;
    return sync_ctx(c) ? 0 : -1;
}

// The text is copied under the lock that writers hold into a buffer owned by the calling thread, so a reader never sees a
// half-written message and the returned pointer stays valid until this thread's next call.
const char* NcrLastError(void) {
    static thread_local char copy[sizeof(g_err)];
    std::lock_guard<std::mutex> lk(g_mu);
    memcpy(copy, g_err, sizeof(copy));
    return copy;
}

const char* NcrDeviceName(void) {
    static thread_local char copy[256];
    std::lock_guard<std::mutex> lk(g_mu);
    snprintf(copy, sizeof(copy), "%s", g_dev >= 0 ? g_devname[g_dev] : "");
    return copy;
}

void* NcrAllocHost(unsigned long long bytes) {
    if (!use_device()) return nullptr;
    void* p = nullptr;
    if (!CK(cudaMallocHost(&p, bytes ? bytes : 16))) return nullptr;
    return p;
}

void NcrFreeHost(void* p) {
    if (p && use_device()) cudaFreeHost(p);
}

void NcrGetStats(RenderContext* ctx, NcrStats* out) {
    NcrContext* c = live(ctx);
    if (!c || !out) return;
    if (use_ctx(c)) sync_ctx(c);
    *out = c->stats;
}

void NcrSetStatsMode(RenderContext* ctx, int mode) {
    NcrContext* c = live(ctx);
    if (c) c->stats_mode = mode;
}

unsigned long long NcrKernelLaunchCount(void) { return g_launches.load(); }

double NcrMeasureF64Rate(void) {
    if (!use_device()) return 0.0;
    g_launches += 2;
    return ncr_measure_f64_rate(0);
}

// Measurement aid: the rate at which this process can bring frames back, i.e. `streams` concurrent device -> pinned-host
// copies of `bytes_per_copy` each, repeated `iters` times, on the default device.  Returns bytes per second (0 on failure).
// Run by every rank at the same time, the sum is the box's aggregate readback ceiling that frame-sharded renders meet.
double NcrMeasureD2HRate(unsigned long long bytes_per_copy, int streams, int iters) {
    if (!use_device() || bytes_per_copy == 0 || streams < 1 || iters < 1) return 0.0;
    streams = std::min(streams, 16);
    std::vector<cudaStream_t> st((size_t)streams, nullptr);
    std::vector<void*> dev((size_t)streams, nullptr), host((size_t)streams, nullptr);
    bool ok = true;
    for (int k = 0; ok && k < streams; ++k)
        ok = CK(cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking)) && CK(cudaMalloc(&dev[k], bytes_per_copy)) &&
             CK(cudaMallocHost(&host[k], bytes_per_copy)) && CK(cudaMemsetAsync(dev[k], k + 1, bytes_per_copy, st[k]));
    double rate = 0.0;
    if (ok) {
        for (int k = 0; k < streams; ++k) cudaMemcpyAsync(host[k], dev[k], bytes_per_copy, cudaMemcpyDeviceToHost, st[k]);   // warm-up
        for (int k = 0; k < streams; ++k) cudaStreamSynchronize(st[k]);
        const auto t0 = std::chrono::steady_clock::now();
        for (int it = 0; it < iters; ++it)
            for (int k = 0; k < streams; ++k) cudaMemcpyAsync(host[k], dev[k], bytes_per_copy, cudaMemcpyDeviceToHost, st[k]);
        for (int k = 0; k < streams; ++k) ok = CK(cudaStreamSynchronize(st[k])) && ok;
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (ok && s > 0) rate = (double)bytes_per_copy * streams * iters / s;
    }
    for (int k = 0; k < streams; ++k) {
        if (host[k]) cudaFreeHost(host[k]);
        if (dev[k]) cudaFree(dev[k]);
        if (st[k]) cudaStreamDestroy(st[k]);
    }
    return rate;
}

int NcrRerunLastFlushEx(RenderContext* ctx, int iters, int flush_l2, float* ms_out, int write_fb, int prefetch) {
    NcrContext* c = live(ctx);
    if (!c || !c->has_last || !use_ctx(c)) return -1;
    if (!sync_ctx(c)) return -1;
    if (flush_l2 && !g_l2_scrub[c->dev]) {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!g_l2_scrub[c->dev] && cudaMalloc(&g_l2_scrub[c->dev], kL2ScrubBytes) != cudaSuccess) return -1;
    }
    NcrFlushArgs R = c->last;
    if (write_fb >= 0) R.write_fb = write_fb ? 1u : 0u;
    if (prefetch >= 0) R.prefetch = prefetch ? 1u : 0u;
    for (int it = 0; it < iters; ++it) {
        if (flush_l2) cudaMemsetAsync(g_l2_scrub[c->dev], it & 0xff, kL2ScrubBytes, c->stream);
        cudaEventRecord(c->ev_t0, c->stream);
        ncr_launch_flush(&R, c->stream, c->ev);
        cudaEventRecord(c->ev_t1, c->stream);
        g_launches += 2u + (R.coarse_list ? 1u : 0u) + (R.yuv_out ? 1u : 0u);
        c->stats.kernel_launches += 2u + (R.coarse_list ? 1u : 0u) + (R.yuv_out ? 1u : 0u);
        if (!CK(cudaStreamSynchronize(c->stream))) { c->failed = true; return -1; }
        if (ms_out) {
            float* o = ms_out + 4 * it;
            cudaEventElapsedTime(&o[0], c->ev_t0, c->ev_t1);   // whole step: cursor reset + 3 kernels
            cudaEventElapsedTime(&o[1], c->ev[0], c->ev[1]);   // ncr_bin_coarse
            cudaEventElapsedTime(&o[2], c->ev[1], c->ev[2]);   // ncr_bin_fine
            cudaEventElapsedTime(&o[3], c->ev[2], c->ev[3]);   // ncr_composite
        }
    }
    if (iters > 0 && R.write_fb) { c->fb_stale = false; c->last.write_fb = 1; }   // the canvas now holds the batch's result
    c->u8_valid = c->last.u8_out != nullptr;
    c->yuv_valid = c->last.yuv_out != nullptr;
    return 0;
}

int NcrRerunLastFlush(RenderContext* ctx, int iters, int flush_l2, float* ms_out) {
    return NcrRerunLastFlushEx(ctx, iters, flush_l2, ms_out, -1, -1);
}

void NcrSetClipRect(RenderContext* ctx, long x, long y, long width, long height) {
    NcrContext* c = live(ctx);
    if (!c) return;
    c->clip_on = true;
    c->clip_l = std::max((i64)0, (i64)x);
    c->clip_t = std::max((i64)0, (i64)y);
    c->clip_r = std::min(c->w, (i64)(x + width));
    c->clip_b = std::min(c->h, (i64)(y + height));
}

void NcrClearClipRect(RenderContext* ctx) {
    NcrContext* c = live(ctx);
    if (c) c->clip_on = false;
}

void NcrSetSampling(RenderContext* ctx, int mode) {
    NcrContext* c = live(ctx);
    if (c) c->sampling = mode == 1 ? 1 : 0;
}

void NcrFillPolygon(RenderContext* ctx, const double* xy, long n_points, double r, double g, double b, double a) {
    NcrContext* c = live(ctx);
    if (!c || !xy || n_points <= 0) return;
    record_polygon(c, xy, (size_t)n_points, r, g, b, a);
}

void NcrDrawTexturePerspective(RenderContext* ctx, Texture* tex_, const double inv_h[9], double x, double y, double width,
                               double height) {
    NcrContext* c = live(ctx);
    NcrTexture* tex = live(tex_);
    if (!c || !tex || !inv_h) return;
    if (width == 0 || height == 0) return;
    i64 tw, th;
    tex_dims(tex, &tw, &th);
    const void* ptr; uint32_t tflags; const DevRef* keep = nullptr;
    if (!bind_texture(c->dev, tex, &ptr, &tflags, &keep)) return;
    // Pixel box: forward-map the source rectangle through the inverse of inv_h (a projective map keeps the rectangle's
    // image inside the hull of its corners as long as the homogeneous w stays positive on it), pad, and fall back to the
    // whole canvas whenever that cannot be trusted (singular / ill-conditioned matrix, w <= 0 at a corner, non-finite).
    i64 bl = 0, br = c->w, bt = 0, bb = c->h;
    {
        const double* h = inv_h;
        const double A0 = h[4] * h[8] - h[5] * h[7], A1 = h[2] * h[7] - h[1] * h[8], A2 = h[1] * h[5] - h[2] * h[4];
        const double B0 = h[5] * h[6] - h[3] * h[8], B1 = h[0] * h[8] - h[2] * h[6], B2 = h[2] * h[3] - h[0] * h[5];
        const double C0 = h[3] * h[7] - h[4] * h[6], C1 = h[1] * h[6] - h[0] * h[7], C2 = h[0] * h[4] - h[1] * h[3];
        const double det = h[0] * A0 + h[1] * B0 + h[2] * C0;
        bool ok = det != 0 && isfinite(det);
        double minx = INFINITY, maxx = -INFINITY, miny = INFINITY, maxy = -INFINITY;
        const double cxs[4] = {x, x + width, x, x + width}, cys[4] = {y, y, y + height, y + height};
        const double tol = 1e-6 * (fabs(width) + fabs(height));
        for (int k = 0; ok && k < 4; ++k) {
            const double X = A0 * cxs[k] + A1 * cys[k] + A2, Y = B0 * cxs[k] + B1 * cys[k] + B2;
            const double Wd = (C0 * cxs[k] + C1 * cys[k] + C2) / det;
            const double px = X / det / Wd, py = Y / det / Wd;
            if (!(Wd > 0) || !isfinite(px) || !isfinite(py) || fabs(px) > 1e7 || fabs(py) > 1e7) { ok = false; break; }
            // trust the forward-mapped corner only if mapping it back through inv_h reproduces the source corner
            const double bw = h[6] * px + h[7] * py + h[8];
            const double bxs = (h[0] * px + h[1] * py + h[2]) / bw, bys = (h[3] * px + h[4] * py + h[5]) / bw;
            if (!(bw > 0) || !(fabs(bxs - cxs[k]) <= tol) || !(fabs(bys - cys[k]) <= tol)) { ok = false; break; }
            minx = std::min(minx, px); maxx = std::max(maxx, px);
            miny = std::min(miny, py); maxy = std::max(maxy, py);
        }
        if (ok) {
            const double pad = 2.0;
            const double L = floor(minx - pad), R = ceil(maxx + pad) + 1, T = floor(miny - pad), B = ceil(maxy + pad) + 1;
            bl = L <= 0 ? 0 : (L >= (double)c->w ? c->w : (i64)L);
            br = R <= 0 ? 0 : (R >= (double)c->w ? c->w : (i64)R);
            bt = T <= 0 ? 0 : (T >= (double)c->h ? c->h : (i64)T);
            bb = B <= 0 ? 0 : (B >= (double)c->h ? c->h : (i64)B);
        }
    }
    NcrCmd* cmd = begin_cmd(c, NCR_OP_TEX_PERSP, bl, br, bt, bb);
    if (!cmd) return;
    fill_texture_fields(c, cmd, ptr, tflags, *keep, tw, th);
    for (int k = 0; k < 6; ++k) cmd->inv[k] = inv_h[k];
    cmd->p[0] = inv_h[6]; cmd->p[1] = inv_h[7]; cmd->p[2] = inv_h[8];
    cmd->x = x; cmd->y = y; cmd->xw = x + width; cmd->yh = y + height;
    cmd->sx = tw / width; cmd->sy = th / height;
}

// Array form of the sprite idiom every reference host loops over (mil:977-1007, pyb:675-716):
//     save_state; apply_transform(m[k]); apply_color_transform(ct[k]); draw_[splitted_]texture(tex, xywh[k][, uv[k]]); restore_state
// for k = 0..n-1, in order, with ONE FFI crossing (a Python call costs ~2 us; a chart frame has ~16,000 of them).  It calls the
// entry points above, so the result is the one the loop gives, bit for bit.
long NcrDrawTextureBatch(RenderContext* ctx, Texture* tex, long n, const double* m6, const double* ct4, const double* xywh,
                         const double* uv4) {
    if (!live(ctx) || !live(tex) || n < 0 || (n > 0 && !xywh)) return -1;
    for (long k = 0; k < n; ++k) {
        SaveContextState(ctx);
        if (m6) ApplyTransform(ctx, m6[6 * k], m6[6 * k + 1], m6[6 * k + 2], m6[6 * k + 3], m6[6 * k + 4], m6[6 * k + 5]);
        if (ct4) ApplyColorTransform(ctx, ct4[4 * k], ct4[4 * k + 1], ct4[4 * k + 2], ct4[4 * k + 3]);
        const double* r = xywh + 4 * k;
        if (uv4) DrawSplittedTexture(ctx, tex, r[0], r[1], r[2], r[3], uv4[4 * k], uv4[4 * k + 1], uv4[4 * k + 2], uv4[4 * k + 3]);
        else DrawTexture(ctx, tex, r[0], r[1], r[2], r[3]);
        RestoreContextState(ctx);
    }
    return n;
}

long NcrSubmitTrace(RenderContext* ctx, const void* trace, long bytes, Texture* const* textures, long n_textures) {
    NcrContext* c = live(ctx);
    if (!c || !trace || bytes < 0) return -1;
    const unsigned char* p = (const unsigned char*)trace;
    const unsigned char* end = p + bytes;
    long executed = 0;
    while (p + sizeof(NcrTraceRec) <= end) {
        NcrTraceRec rec;
        memcpy(&rec, p, sizeof(rec));
        p += sizeof(rec);
        if ((size_t)(end - p) < (size_t)rec.n * sizeof(double)) return -1;
        const double* a = (const double*)p;   // records are 8-byte aligned by construction
        p += (size_t)rec.n * sizeof(double);
        Texture* tex = nullptr;
        if (rec.op == NCR_T_DRAW_TEXTURE || rec.op == NCR_T_DRAW_SPLIT || rec.op == NCR_T_DRAW_PERSP) {
            if (rec.n < 1) return -1;
            const long slot = (long)a[0];
            if (slot < 0 || slot >= n_textures) return -1;
            tex = textures[slot];
        }
#define NEED(k) if (rec.n != (k)) return -1
        switch (rec.op) {
            case NCR_T_SAVE: SaveContextState(ctx); break;
            case NCR_T_RESTORE: RestoreContextState(ctx); break;
            case NCR_T_SET_TRANSFORM: NEED(6); SetTransform(ctx, a[0], a[1], a[2], a[3], a[4], a[5]); break;
            case NCR_T_APPLY_TRANSFORM: NEED(6); ApplyTransform(ctx, a[0], a[1], a[2], a[3], a[4], a[5]); break;
            case NCR_T_SCALE: NEED(2); Scale(ctx, a[0], a[1]); break;
            case NCR_T_TRANSLATE: NEED(2); Translate(ctx, a[0], a[1]); break;
            case NCR_T_ROTATE: NEED(1); Rotate(ctx, a[0]); break;
            case NCR_T_SET_CT: NEED(4); SetColorTransform(ctx, a[0], a[1], a[2], a[3]); break;
            case NCR_T_APPLY_CT: NEED(4); ApplyColorTransform(ctx, a[0], a[1], a[2], a[3]); break;
            case NCR_T_SET_COLOR: NEED(4); SetColor(ctx, a[0], a[1], a[2], a[3]); break;
            case NCR_T_FILL_COLOR: NEED(4); FillColor(ctx, a[0], a[1], a[2], a[3]); break;
            case NCR_T_DRAW_TEXTURE: NEED(5); DrawTexture(ctx, tex, a[1], a[2], a[3], a[4]); break;
            case NCR_T_DRAW_SPLIT: NEED(9); DrawSplittedTexture(ctx, tex, a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]); break;
            case NCR_T_DRAW_RECT: NEED(8); DrawRect(ctx, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
            case NCR_T_DRAW_LINE: NEED(9); DrawLine(ctx, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]); break;
            case NCR_T_DRAW_CIRCLE: NEED(7); DrawCircle(ctx, a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
            case NCR_T_DRAW_GRD: NEED(12); DrawVerticalGrd(ctx, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11]); break;
            case NCR_T_SET_PIXEL: NEED(6); SetPixel(ctx, (long)a[0], (long)a[1], a[2], a[3], a[4], a[5]); break;
            case NCR_T_APPLY_PIXEL: NEED(6); ApplyPixel(ctx, (long)a[0], (long)a[1], a[2], a[3], a[4], a[5]); break;
            case NCR_T_PRESENT: break;   // frame boundary: the caller reads back
            case NCR_T_CLIP_SET: NEED(4); NcrSetClipRect(ctx, (long)a[0], (long)a[1], (long)a[2], (long)a[3]); break;
            case NCR_T_CLIP_CLEAR: NcrClearClipRect(ctx); break;
            case NCR_T_SAMPLING: NEED(1); NcrSetSampling(ctx, (int)a[0]); break;
            case NCR_T_FILL_POLY:
                if (rec.n < 6 || (rec.n & 1)) return -1;
                NcrFillPolygon(ctx, a + 4, (rec.n - 4) / 2, a[0], a[1], a[2], a[3]);
                break;
            case NCR_T_DRAW_PERSP: NEED(14); NcrDrawTexturePerspective(ctx, tex, a + 1, a[10], a[11], a[12], a[13]); break;
            default: return -1;
        }
#undef NEED
        ++executed;
    }
    if (p != end) return -1;   // trailing partial record
    return executed;
}

}   // extern "C"
