// Off-path parts of the reference C ABI, kept as plain host C++ so that the reference's applications
// (src/milrenderer.py, src/hjm_mixer.py) still find every symbol they bind:
//   * the audio clip engine (reference cpp:990-1283) — 1-D f64 PCM adds and a linear resampler; nothing here
//     is worth a GPU;
//   * the MP4 writer entry points (reference cpp:59-275).  FFmpeg is not part of this build, so the encoder is
//     absent: InitializeVideoCap reports failure, and PutRendererContextFrame performs the present-side
//     work that IS on the path (flush + fused f64->u8 image + YUV 4:2:0 planes on the device, 1.5 B/px readback) and drops the frame.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>

#include "../../include/ncr_b200.h"

typedef long i64;
typedef double f64;

struct NcrAudioClip {
    i64 sampleRate = 0, channels = 0, numFrames = 0;
    std::vector<f64> pcm;   // interleaved, numFrames * channels
};

struct NcrBytes {
    std::vector<unsigned char> data;
};

struct NcrVideoCap {
    i64 width = 0, height = 0;
    f64 frameRate = 0;
    i64 frames = 0;
    std::vector<unsigned char> last;
};

static NcrAudioClip* make_clip(i64 rate, i64 ch, i64 frames) {
    NcrAudioClip* c = new NcrAudioClip();
    c->sampleRate = rate;
    c->channels = ch;
    c->numFrames = frames;
    const i64 n = frames * ch;
    c->pcm.assign(n > 0 ? (size_t)n : 0, 0.0);
    return c;
}

extern "C" {

long GetAudioClipBufferSizeFromData(long numFrames, long channels) { return numFrames * channels; }
long GetAudioClipBufferSize(AudioClip* clip) { return clip ? clip->numFrames * clip->channels : 0; }

AudioClip* CreateAudioClipFromBuffer(long sampleRate, long channels, long numFrames, double* buffer) {
    NcrAudioClip* c = make_clip(sampleRate, channels, numFrames);
    if (buffer && !c->pcm.empty()) memcpy(c->pcm.data(), buffer, c->pcm.size() * sizeof(f64));
    return c;
}

AudioClip* CreateAudioClipFromInt16Buffer(long sampleRate, long channels, long numFrames, short* buffer) {
    NcrAudioClip* c = make_clip(sampleRate, channels, numFrames);
    if (buffer)
        for (size_t k = 0; k < c->pcm.size(); ++k) c->pcm[k] = (f64)buffer[k] / 32768.0;   // cpp:1029
    return c;
}

AudioClip* CreateSilentAudioClip(long sampleRate, long channels, long numFrames) {
    return make_clip(sampleRate, channels, numFrames);
}

void DestroyAudioClip(AudioClip* clip) { (void)clip; /* reference: no-op (cpp:1048-1052); Python may reuse the pointer */ }

AudioClip* CloneAudioClip(AudioClip* clip) {
    if (!clip) return nullptr;
    return CreateAudioClipFromBuffer(clip->sampleRate, clip->channels, clip->numFrames, clip->pcm.data());
}

double GetAudioClipDuration(AudioClip* clip) { return clip ? (f64)clip->numFrames / (f64)clip->sampleRate : 0.0; }

// Linear-interpolating resampler with the reference's index clamp (cpp:1075-1111), including its use of
// `numFrames - channels` as the upper bound and the channel-averaging branch when the channel count changes.
void ApplyResampleAudioClip(AudioClip* clip, long sampleRate, long channels) {
    if (!clip) return;
    if (clip->sampleRate == sampleRate && clip->channels == channels) return;
    const f64 dur = GetAudioClipDuration(clip);
    const i64 outFrames = (i64)(dur * sampleRate);
    const i64 inCh = clip->channels;
    const i64 hiClamp = clip->numFrames - inCh;
    std::vector<f64> out((size_t)(outFrames > 0 ? outFrames * channels : 0));
    const f64* in = clip->pcm.data();
    for (i64 i = 0; i < outFrames; ++i) {
        const f64 pos = ((f64)i / sampleRate) * clip->sampleRate;
        i64 lo = (i64)floor(pos), hi = (i64)ceil(pos);
        if (lo < 0) lo = 0;
        if (lo >= hiClamp) lo = hiClamp - 1;
        if (hi < 0) hi = 0;
        if (hi >= hiClamp) hi = hiClamp - 1;
        const f64 frac = pos - lo;
        if (inCh == channels) {
            for (i64 ch = 0; ch < channels; ++ch) {
                const f64 a = in[lo * inCh + ch], b = in[hi * inCh + ch];
                out[i * channels + ch] = a + (b - a) * frac;
            }
        } else {
            f64 sumLo = 0, sumHi = 0;
            for (i64 ch = 0; ch < inCh; ++ch) {
                sumLo += in[lo * inCh + ch];
                sumHi += in[hi * inCh + ch];
            }
            const f64 v = sumLo / inCh + (sumHi / inCh - sumLo / inCh) * frac;
            for (i64 ch = 0; ch < channels; ++ch) out[i * channels + ch] = v;
        }
    }
    clip->pcm.swap(out);
    clip->sampleRate = sampleRate;
    clip->channels = channels;
    clip->numFrames = outFrames;
}

void ResampleAudioClipLike(AudioClip* clip, AudioClip* like) {
    if (clip && like) ApplyResampleAudioClip(clip, like->sampleRate, like->channels);
}

long OverlayAudioClip(AudioClip* target, AudioClip* source, long startFrame, bool autoResample) {
    if (!target || !source) return -3;
    NcrAudioClip* tmp = nullptr;
    if (autoResample && (target->sampleRate != source->sampleRate || target->channels != source->channels)) {
        tmp = CloneAudioClip(source);
        ResampleAudioClipLike(tmp, target);
        source = tmp;
    }
    long rc = 0;
    if (target->sampleRate != source->sampleRate) rc = -1;         // cpp:1142
    else if (target->channels != source->channels) rc = -2;        // cpp:1143
    else {
        const i64 ch = source->channels;
        for (i64 i = 0; i < source->numFrames; ++i) {
            const i64 at = startFrame + i;
            if (at >= target->numFrames) break;
            if (at < 0) continue;   // the reference writes before the buffer here; dropped
            for (i64 k = 0; k < ch; ++k) target->pcm[at * ch + k] += source->pcm[i * ch + k];
        }
    }
    delete tmp;
    return rc;
}

long OverlayAudioClipSecond(AudioClip* target, AudioClip* source, double startSecond, bool autoResample) {
    if (!target) return -3;
    return OverlayAudioClip(target, source, (i64)(startSecond * target->sampleRate), autoResample);
}

static void put_u32(unsigned char* p, uint32_t v) { memcpy(p, &v, 4); }
static void put_u16(unsigned char* p, uint16_t v) { memcpy(p, &v, 2); }

// 16-bit PCM RIFF/WAVE image of the clip (cpp:1165-1228): 44-byte header, samples clamped to [-1,1] * 32767.
WapperedBytes* SaveAudioClipAsWav(AudioClip* clip) {
    if (!clip) return nullptr;
    const i64 n = clip->numFrames * clip->channels;
    NcrBytes* out = new NcrBytes();
    out->data.resize(44 + (size_t)n * 2);
    unsigned char* d = out->data.data();
    memcpy(d, "RIFF", 4);
    put_u32(d + 4, (uint32_t)(out->data.size() - 8));
    memcpy(d + 8, "WAVEfmt ", 8);
    put_u32(d + 16, 16);
    put_u16(d + 20, 1);
    put_u16(d + 22, (uint16_t)clip->channels);
    put_u32(d + 24, (uint32_t)clip->sampleRate);
    put_u32(d + 28, (uint32_t)(clip->sampleRate * clip->channels * 2));
    put_u16(d + 32, (uint16_t)(clip->channels * 2));
    put_u16(d + 34, 16);
    memcpy(d + 36, "data", 4);
    put_u32(d + 40, (uint32_t)(n * 2));
    for (i64 k = 0; k < n; ++k) {
        f64 v = clip->pcm[k];
        v = v > 1.0 ? 1.0 : (v < -1.0 ? -1.0 : v);
        put_u16(d + 44 + 2 * k, (uint16_t)(int16_t)(v * 32767.0));
    }
    return out;
}

long GetAudioClipSampleRate(AudioClip* clip) { return clip ? clip->sampleRate : 0; }
long GetAudioClipChannels(AudioClip* clip) { return clip ? clip->channels : 0; }
long GetAudioClipNumFrames(AudioClip* clip) { return clip ? clip->numFrames : 0; }
unsigned char* GetWapperedBytesDataPtr(WapperedBytes* bytes) { return bytes ? bytes->data.data() : nullptr; }
long GetWapperedBytesDataSize(WapperedBytes* bytes) { return bytes ? (long)bytes->data.size() : 0; }

void ApplyVolumeGain(AudioClip* clip, double gain) {
    if (clip)
        for (f64& v : clip->pcm) v *= gain;
}

void ApplyCutAudioClip(AudioClip* clip, long startFrame, long endFrame) {
    if (!clip) return;
    const i64 frames = endFrame - startFrame, ch = clip->channels;
    std::vector<f64> out((size_t)(frames > 0 ? frames * ch : 0), 0.0);   // reference leaves the tail uninitialised
    for (i64 i = 0; i < frames; ++i) {
        const i64 src = startFrame + i;
        if (src >= clip->numFrames) break;
        if (src < 0) continue;
        for (i64 k = 0; k < ch; ++k) out[i * ch + k] = clip->pcm[src * ch + k];
    }
    clip->pcm.swap(out);
    clip->numFrames = frames;
}

void ApplySpeedAudioClip(AudioClip* clip, double speed) {
    if (clip) clip->sampleRate = (i64)(clip->sampleRate * speed);   // cpp:1282: i64 *= f64
}

// ---- MP4 writer entry points (encoder not built: FFmpeg headers are not available) -------------------------
VideoCap* CreateVideoCap(long width, long height, double frameRate) {
    NcrVideoCap* cap = new NcrVideoCap();
    cap->width = width;
    cap->height = height;
    cap->frameRate = frameRate;
    return cap;
}

bool InitializeVideoCap(VideoCap* cap, const char* path, bool hasAudio, AudioClip* aClip, long aBitRate) {
    (void)cap; (void)hasAudio; (void)aClip; (void)aBitRate;
    fprintf(stderr, "[libNativeCPURenderer/b200] InitializeVideoCap(%s): built without FFmpeg, no encoder available\n",
            path ? path : "");
    return false;
}

void DestroyVideoCap(VideoCap* cap) { (void)cap; /* reference: no-op (cpp:47-50) */ }

void PutRendererContextFrame(VideoCap* cap, RenderContext* ctx) {
    if (!cap || !ctx) return;
    // cpp:232-256: f64 -> u8 truncation, then RGB(A) -> YUV420P for the encoder.  Both steps run on the device and only
    // the planes (1.5 B/px) come back; with no encoder linked (FFmpeg absent) the frame is kept as `last` and counted.
    // cap size != canvas size: libswscale resizes while converting (sws_getContext(ctx w, h -> cap w, h, SWS_BILINEAR), cpp:241-246)
    if (cap->width <= 0 || cap->height <= 0) return;
    const long n = cap->width * cap->height + 2 * ((cap->width + 1) / 2) * ((cap->height + 1) / 2);
    cap->last.resize((size_t)n);
    if (NcrGetBufferAsYUV420PScaled(ctx, cap->width, cap->height, cap->last.data()) != n) return;
    cap->frames += 1;
}

void ReleaseVideoCap(VideoCap* cap) {
    if (cap) cap->last.clear();
}

bool PutAudioIntoVideoCap(VideoCap* vCap, AudioClip* aClip, long bitRate) {
    (void)vCap; (void)aClip; (void)bitRate;
    return false;   // declared at h:142 but never defined in the reference
}

}   // extern "C"
