/* C ABI of the B200-native libNativeCPURenderer.so (drop-in boundary).
 *
 * Section 1 is, symbol for symbol, the interface the reference's ctypes binding
 * (reference src/libNativeCPURendererPybind.py:9, `ctypes.CDLL("./libNativeCPURenderer.so")`)
 * and src/milrenderer.py call; each declaration cites the reference declaration
 * (h = src/libNativeCPURenderer.h) and definition (cpp = src/libNativeCPURenderer.cpp) it
 * replaces.  Handles are opaque; `long` is the reference's i64 (h:2), `double` its f64.
 * No function throws across this boundary; device errors are reported on stderr and
 * through NcrLastError().
 *
 * Section 2 is additive (prefix Ncr): batching, pinned host memory, measurement hooks and
 * the extensions BASELINE.json's configs name that the reference does not implement
 * (clip rect, bilinear sampling, polygon fill — each pinned bit-exactly to reference code, DESIGN.md
 * section 5: the unmodified reference with outside pixels put back / the four-tap code it keeps commented
 * out at cpp:575-620 / DrawLine's own pointInPolygon + ApplyPixel loop — and perspective quads, whose
 * projective map is this repo's spec and whose bounds / sampling / blend are DrawTexture's own).
 */
#ifndef NCR_B200_H
#define NCR_B200_H
#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct NcrContext RenderContext;   /* h:32-42 */
typedef struct NcrTexture Texture;         /* h:44-49 */
typedef struct NcrVideoCap VideoCap;       /* h:51-68 */
typedef struct NcrAudioClip AudioClip;     /* h:70-76 */
typedef struct NcrBytes WapperedBytes;     /* h:78-81 (sic) */

/* ---------------------------------------------------------------------------------------------
 * 1. Reference ABI — render path (hot path: HBM-resident canvas, recorded draws, CUDA composite)
 * ------------------------------------------------------------------------------------------- */
long GetBufferSize(RenderContext* ctx);                                   /* h:84  cpp:3-5    W*H*ipp elements */
RenderContext* CreateRenderContext(long width, long height, bool enableAlpha); /* h:85 cpp:7-31; NULL if no usable GPU */
void DestroyRenderContext(RenderContext* ctx);                            /* h:86  cpp:33-37  (reference: no-op) */
void ResizeRenderContext(RenderContext* ctx, long width, long height);    /* h:149 cpp:39-45  discards pixels, keeps state */
void SaveContextState(RenderContext* ctx);                                /* h:92  cpp:277-290 */
bool RestoreContextState(RenderContext* ctx);                             /* h:93  cpp:292-309 false on empty stack */
void GetBuffer(RenderContext* ctx, double* buffer);                       /* h:94  cpp:311-316 flushes; f64 canvas out */
void GetBufferAsUInt8(RenderContext* ctx, unsigned char* buffer);         /* h:95  cpp:52-57  flushes; (iu8)(v*255) */
Texture* CreateTexture(long width, long height, bool enableAlpha, double* buffer);           /* h:96 cpp:318-335 */
Texture* CreateTextureUInt8(long width, long height, bool enableAlpha, unsigned char* buffer); /* h:97 cpp:337-354 */
void DestroyTexture(Texture* tex);                                        /* h:98  cpp:356-360 (reference: no-op) */
Texture* CreateTextureFromRenderContext(RenderContext* ctx);              /* h:99  cpp:362-375 deep copy */
Texture* CreateTextureFromRenderContextShared(RenderContext* ctx);        /* h:148 cpp:377-384 alias of the canvas */
void SetTransform(RenderContext* ctx, double a, double b, double c, double d, double e, double f);   /* h:100 cpp:386-396 */
void ApplyTransform(RenderContext* ctx, double a, double b, double c, double d, double e, double f); /* h:101 cpp:398-411 */
void Scale(RenderContext* ctx, double sx, double sy);                     /* h:102 cpp:420-426 */
void Translate(RenderContext* ctx, double tx, double ty);                 /* h:103 cpp:428-434 */
void Rotate(RenderContext* ctx, double angle);                            /* h:104 cpp:436-444 */
void TransformPoint(RenderContext* ctx, double x, double y, double* out_x, double* out_y); /* h:105 cpp:455-461 (inline in the reference) */
void GetTransform(RenderContext* ctx, double out_matrix[6]);              /* h:106 cpp:463-470 */
void GetInverseTransform(RenderContext* ctx, double out_matrix[6]);       /* h:107 cpp:472-492 */
bool SetPixel(RenderContext* ctx, long x, long y, double r, double g, double b, double a);   /* h:108 cpp:494-513 */
bool ApplyPixel(RenderContext* ctx, long x, long y, double r, double g, double b, double a); /* h:109 cpp:515-549 (inline in the reference) */
void SetColorTransform(RenderContext* ctx, double r, double g, double b, double a);   /* h:110 cpp:623-631 */
void ApplyColorTransform(RenderContext* ctx, double r, double g, double b, double a); /* h:111 cpp:633-641 */
void SetColor(RenderContext* ctx, double r, double g, double b, double a);            /* h:112 cpp:643-657 */
void GetColor(RenderContext* ctx, double x, double y, double* out_r, double* out_g, double* out_b, double* out_a); /* h:113 cpp:659-680 */
void FillColor(RenderContext* ctx, double r, double g, double b, double a);           /* h:114 cpp:682-691 */
void DrawTexture(RenderContext* ctx, Texture* tex, double x, double y, double width, double height); /* h:115 cpp:720-779 */
void DrawRect(RenderContext* ctx, double x, double y, double width, double height, double r, double g, double b, double a); /* h:116 cpp:847-874 */
void DrawLine(RenderContext* ctx, double x1, double y1, double x2, double y2, double width, double r, double g, double b, double a); /* h:117 cpp:876-918 */
void DrawCircle(RenderContext* ctx, double x, double y, double radius, double r, double g, double b, double a); /* h:118 cpp:920-948 */
Texture* ResampleTexture(Texture* tex, long width, long height);          /* h:119 cpp:950-976 */
long GetTextureWidth(Texture* tex);                                       /* h:120 cpp:978-980 */
long GetTextureHeight(Texture* tex);                                      /* h:121 cpp:982-984 */
bool GetTextureEnableAlpha(Texture* tex);                                 /* h:122 cpp:986-988 */
long GetVersion(void);                                                    /* h:143 cpp:1261-1263 */
void DrawVerticalGrd(RenderContext* ctx, double x, double y, double width, double height,
                     double top_r, double top_g, double top_b, double top_a,
                     double bottom_r, double bottom_g, double bottom_b, double bottom_a); /* h:146 cpp:1285-1316 */
void DrawSplittedTexture(RenderContext* ctx, Texture* tex, double x, double y, double width, double height,
                         double uStart, double uEnd, double vStart, double vEnd);          /* h:147 cpp:781-820 */
void GetMilthmHitEffectPixel(double seed, double t, double x, double y, double* a);        /* h:150 cpp:1406-1411 */
Texture* CreateMilthmHitEffectTexture(Texture* mask, double seed, double t, double r, double g, double b); /* h:151 cpp:1417-1440 */

/* Reference ABI — off the hot path; plain host code so the reference's applications keep working. */
VideoCap* CreateVideoCap(long width, long height, double frameRate);      /* h:87  cpp:65-77 */
bool InitializeVideoCap(VideoCap* cap, const char* path, bool hasAudio, AudioClip* aClip, long aBitRate); /* h:88 cpp:79-196; false: built without FFmpeg */
void DestroyVideoCap(VideoCap* cap);                                      /* h:89  cpp:47-50 */
void PutRendererContextFrame(VideoCap* cap, RenderContext* ctx);          /* h:90  cpp:232-275; flush + u8 + YUV420P planes on the device, no encoder */
void ReleaseVideoCap(VideoCap* cap);                                      /* h:91  cpp:198-230 */
bool PutAudioIntoVideoCap(VideoCap* vCap, AudioClip* aClip, long bitRate); /* h:142 declared, never defined in the reference */
long GetAudioClipBufferSizeFromData(long numFrames, long channels);       /* h:123 cpp:990-992 */
long GetAudioClipBufferSize(AudioClip* clip);                             /* h:124 cpp:994-996 */
AudioClip* CreateAudioClipFromBuffer(long sampleRate, long channels, long numFrames, double* buffer);     /* h:125 cpp:998-1014 */
AudioClip* CreateAudioClipFromInt16Buffer(long sampleRate, long channels, long numFrames, short* buffer); /* h:126 cpp:1016-1034 */
AudioClip* CreateSilentAudioClip(long sampleRate, long channels, long numFrames); /* h:127 cpp:1036-1046 */
void DestroyAudioClip(AudioClip* clip);                                   /* h:128 cpp:1048-1052 */
AudioClip* CloneAudioClip(AudioClip* clip);                               /* h:129 cpp:1054-1061 */
void ApplyResampleAudioClip(AudioClip* clip, long sampleRate, long channels); /* h:130 cpp:1063-1120 */
void ResampleAudioClipLike(AudioClip* clip, AudioClip* like);             /* h:131 cpp:1122-1127 */
long OverlayAudioClip(AudioClip* target, AudioClip* source, long startFrame, bool autoResample);        /* h:132 cpp:1129-1154 */
long OverlayAudioClipSecond(AudioClip* target, AudioClip* source, double startSecond, bool autoResample); /* h:133 cpp:1156-1163 */
WapperedBytes* SaveAudioClipAsWav(AudioClip* clip);                       /* h:134 cpp:1165-1228 */
long GetAudioClipSampleRate(AudioClip* clip);                             /* h:135 cpp:1230-1232 */
long GetAudioClipChannels(AudioClip* clip);                               /* h:136 cpp:1234-1236 */
long GetAudioClipNumFrames(AudioClip* clip);                              /* h:137 cpp:1238-1240 */
double GetAudioClipDuration(AudioClip* clip);                             /* h:138 cpp:1242-1244 */
unsigned char* GetWapperedBytesDataPtr(WapperedBytes* bytes);             /* h:139 cpp:1246-1248 */
long GetWapperedBytesDataSize(WapperedBytes* bytes);                      /* h:140 cpp:1250-1252 */
void ApplyVolumeGain(AudioClip* clip, double gain);                       /* h:141 cpp:1254-1259 */
void ApplyCutAudioClip(AudioClip* clip, long startFrame, long endFrame);  /* h:144 cpp:1265-1279 */
void ApplySpeedAudioClip(AudioClip* clip, double speed);                  /* h:145 cpp:1281-1283 */

/* ---------------------------------------------------------------------------------------------
 * 2. Additive entry points
 * ------------------------------------------------------------------------------------------- */
typedef struct NcrStats {
    unsigned long long n_cmds;          /* commands in the last flush */
    unsigned long long coarse_entries;  /* bin-list entries written by ncr_bin_coarse */
    unsigned long long fine_entries;    /* region-list entries written by ncr_bin_fine: (command, 16x8 region) pairs that survive the exact test */
    unsigned long long blended_pixels;  /* ApplyPixel executions in the last flush (stats mode & 1) */
    unsigned long long h2d_bytes;       /* cumulative, this context */
    unsigned long long d2h_bytes;       /* cumulative, this context */
    unsigned long long flushes;         /* cumulative */
    unsigned long long kernel_launches; /* cumulative, this context */
    float ms_bin_coarse, ms_bin_fine, ms_composite, ms_total; /* last flush, CUDA events (stats mode & 2) */
    unsigned long long materialized;    /* cumulative: batches re-run to bring a stale f64 canvas up to date (see NcrRerunLastFlushEx) */
    unsigned long long interior_entries; /* of fine_entries: (command, region) pairs proven to cover their whole region (straight-line path) */
} NcrStats;

/* Devices.  CreateRenderContext / CreateTexture* (section 1) use the process default device: NCR_DEVICE, else LOCAL_RANK,
 * else 0 — what a one-process-per-GPU launcher wants.  A single host process (reference src/milrenderer.py is one,
 * mil:865-1038; its own sketch allocates a block of contexts in-process, pyb:362-364) places contexts explicitly; a texture
 * is copied to a device the first time a context on that device draws it, so textures are created once, as in the reference. */
int NcrDeviceCount(void);                          /* usable CUDA devices, 0 without one */
RenderContext* NcrCreateRenderContextOnDevice(long width, long height, bool enableAlpha, int device);
int NcrContextDevice(RenderContext* ctx);          /* device a context lives on, -1 for a bad handle */
int NcrFlush(RenderContext* ctx);                  /* submit pending draws and wait; 0 on success */
const char* NcrLastError(void);                    /* last device/runtime error text ("" if none) */
const char* NcrDeviceName(void);                   /* name of the CUDA device in use, "" before first use */
void* NcrAllocHost(unsigned long long bytes);      /* pinned host memory: readbacks into it are direct DMA */
void NcrFreeHost(void* p);
/* Replay a recorded command stream (format: libnativecpurenderer_b200/trace.py) through the entry points of
 * section 1 without one FFI crossing per call.  textures[k] resolves texture slot k.  Returns the number of
 * records executed, or -1 on a malformed stream. */
long NcrSubmitTrace(RenderContext* ctx, const void* trace, long bytes, Texture* const* textures, long n_textures);
/* Re-execute the last flushed batch from its HBM-resident command buffers `iters` times (measurement: the
 * "inputs already resident" timing).  ms_out[4*iters] receives, per iteration, CUDA-event times on the context's
 * stream: {whole step, ncr_bin_coarse, ncr_bin_fine, ncr_composite}; flush_l2 != 0 overwrites a buffer larger than
 * L2 before each iteration (outside the timed span). */
int NcrRerunLastFlush(RenderContext* ctx, int iters, int flush_l2, float* ms_out);
/* The same with the two launch modes of a flush made explicit (-1 = as the batch was flushed):
 *   write_fb  1: the composite writes the f64 canvas back (what GetBuffer / a following draw needs);
 *             0: present-only — GetBufferAsUInt8 / NcrGetBufferAsYUV420P flushes skip the canvas write-back (the canvas is
 *                marked stale and the resident batch is re-run with write_fb = 1 if anything ever reads it; a video frame's
 *                canvas never is, its successor starts with SetColor);
 *   prefetch  1/0: composite variant with / without cross-region prefetch of the next region's list and first command. */
int NcrRerunLastFlushEx(RenderContext* ctx, int iters, int flush_l2, float* ms_out, int write_fb, int prefetch);
void NcrGetStats(RenderContext* ctx, NcrStats* out);
void NcrSetStatsMode(RenderContext* ctx, int mode); /* bit 0: count blended pixels, bit 1: per-kernel events */
unsigned long long NcrKernelLaunchCount(void);      /* kernels launched by this library since load (all contexts) */
double NcrMeasureD2HRate(unsigned long long bytes_per_copy, int streams, int iters); /* bytes/s of concurrent device -> pinned-host copies (readback ceiling aid) */
double NcrMeasureD2HRate(unsigned long long bytes_per_copy, int streams, int iters); /* bytes/s of concurrent device -> pinned-host copies on the default device (readback ceiling aid) */
double NcrMeasureF64Rate(void);                     /* measured rate of non-fused f64 mul/add instructions per second (roofline aid) */

/* Extensions the reference does not have as entry points; each is pinned to reference code, the perspective
 * map itself being this repo's spec. */
void NcrSetClipRect(RenderContext* ctx, long x, long y, long width, long height); /* intersects every draw's pixel box; = the reference drawing unclipped + outside pixels put back */
void NcrClearClipRect(RenderContext* ctx);
void NcrSetSampling(RenderContext* ctx, int mode); /* 0 nearest (reference), 1 bilinear (the four-tap code commented out at cpp:575-620; bit-identical to it) */
void NcrFillPolygon(RenderContext* ctx, const double* xy, long n_points, double r, double g, double b, double a); /* DrawLine's loop (cpp:906-916) on N caller points: cpp:822-845 even-odd rule + ApplyPixel */
void NcrDrawTexturePerspective(RenderContext* ctx, Texture* tex, const double inv_h[9], double x, double y, double width, double height); /* rw = 1/(h6 i + h7 j + h8), X = (h0 i + h1 j + h2) rw, Y likewise, hw <= 0 skipped; then DrawTexture's mapped loop, cpp:765-777 */
/* Present path (replaces the f64->u8 loop + sws_scale of PutRendererContextFrame, h:91 cpp:232-256, for
 * cap size == canvas size): flush, convert the canvas to the (iu8)(v*255) image and to planar YUV 4:2:0 on the device exactly as
 * libswscale's sws_scale(RGBA|RGB24 -> YUV420P, SWS_BILINEAR) does (BT.601 limited range, pair-sum chroma, {1,3,3,1}/8 vertical
 * taps), and read back only the planes: Y[h][w], U[ch][cw], V[ch][cw] with cw = (w+1)/2, ch = (h+1)/2,
 * contiguous in `out`.  Returns the bytes written (NcrYUV420PSize), -1 on failure.  Pinned bit-exactly to libswscale 9.1.100
 * for even sizes >= 8x8 (tests/golden/make_swscale_fixtures.py). */
/* n sprites in one call: for k in 0..n-1 { SaveContextState; ApplyTransform(m6[k]) if m6; ApplyColorTransform(ct4[k]) if ct4;
 * DrawSplittedTexture(tex, xywh[k], uv4[k]) if uv4 else DrawTexture(tex, xywh[k]); RestoreContextState } — exactly that call
 * sequence (h:92-93,101,111,115,147), so bit-identical to the loop; m6 is [n][6] (a b c d e f), ct4 [n][4], xywh [n][4],
 * uv4 [n][4] = uStart uEnd vStart vEnd.  Returns n, -1 on bad handles. */
long NcrDrawTextureBatch(RenderContext* ctx, Texture* tex, long n, const double* m6, const double* ct4, const double* xywh,
                         const double* uv4);
/* Frame-parallel batch render (finishes reference pyb:302-367 MultiThreadedVideoRenderContextPreparer):
 * n_frames recorded frames (trace format as NcrSubmitTrace) are rendered by n_workers threads with one context / CUDA
 * stream each and delivered to `sink` strictly in frame order (pixels: the RGB(A)8 image for present 0, the YUV 4:2:0
 * planes for present 1; valid only during the call).  Every frame must start by overwriting the canvas with SetColor.
 * Returns n_frames, -1 on a device/trace error, -2 when a frame is not independent. */
typedef void (*NcrFrameSink)(void* user, long frame_index, const unsigned char* pixels, long bytes);
long NcrRenderFrames(long width, long height, int alpha, const void* const* traces, const long* trace_bytes, long n_frames,
                     Texture* const* textures, long n_textures, int n_workers, int present, NcrFrameSink sink, void* user);
/* The same with the worker contexts, their device buffers and pinned frame buffers kept between calls. */
typedef struct NcrFramePool NcrFramePool;
NcrFramePool* NcrCreateFramePool(long width, long height, int alpha, int n_workers);   /* NULL without a usable device */
/* The same pool spread over several devices of the box: worker k renders on devices[k % n_devices]; frames are still
 * delivered in order to one sink.  One host process drives all the GPUs (no torchrun, no collective). */
NcrFramePool* NcrCreateFramePoolOnDevices(long width, long height, int alpha, int n_workers, const int* devices, int n_devices);
void NcrDestroyFramePool(NcrFramePool* pool);
int NcrFramePoolWorkers(NcrFramePool* pool);
long NcrFramePoolRender(NcrFramePool* pool, const void* const* traces, const long* trace_bytes, long n_frames,
                        Texture* const* textures, long n_textures, int present, NcrFrameSink sink, void* user);
long NcrYUV420PSize(RenderContext* ctx);
long NcrGetBufferAsYUV420P(RenderContext* ctx, unsigned char* out);
/* The scaling branch of PutRendererContextFrame (h:91 cpp:241-256: cap size != canvas size, sws_scale resizes with SWS_BILINEAR):
 * the planes at dst_w x dst_h, Y[dst_h][dst_w], U, V[(dst_h+1)/2][(dst_w+1)/2]; returns the bytes written or -1.  Bit-identical to
 * libswscale (pinned against 9.1.100) for source and destination sizes >= 16 with an even source width. */
long NcrGetBufferAsYUV420PScaled(RenderContext* ctx, long dst_w, long dst_h, unsigned char* out);

#ifdef __cplusplus
}
#endif
#endif
