/* TEST INFRASTRUCTURE ONLY.
 * Minimal stand-in declarations for the FFmpeg API surface that the reference's
 * MP4 writer (reference src/libNativeCPURenderer.cpp:59-275) touches, so that the
 * UNMODIFIED reference translation unit compiles in an image without FFmpeg dev
 * headers.  Only names, field names and call shapes matter: every function is
 * defined in ffstub.c as abort(), because the draw/composite path never calls
 * FFmpeg.  The libc includes below are required: the reference relies on FFmpeg's
 * headers to pull in <math.h>/<stdlib.h> (abs() on a double).
 */
#ifndef NCR_FFSTUB_COMMON_H
#define NCR_FFSTUB_COMMON_H
#include <errno.h>
#include <inttypes.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define AV_ERROR_MAX_STRING_SIZE 64
#define AVERROR(e) (-(e))
#define AVERROR_EOF (-541478725)
#define AVFMT_NOFILE 0x0001
#define AVIO_FLAG_WRITE 2
#define SWS_BILINEAR 2

enum AVCodecID { AV_CODEC_ID_NONE = 0, AV_CODEC_ID_H264 = 27, AV_CODEC_ID_AAC = 86018 };
enum AVPixelFormat { AV_PIX_FMT_NONE = -1, AV_PIX_FMT_YUV420P = 0, AV_PIX_FMT_RGB24 = 2, AV_PIX_FMT_RGBA = 26 };
enum AVSampleFormat { AV_SAMPLE_FMT_NONE = -1, AV_SAMPLE_FMT_FLTP = 8 };

typedef struct AVRational { int num; int den; } AVRational;
typedef struct AVCodec { int id; } AVCodec;
typedef struct AVCodecParameters { int codec_id; } AVCodecParameters;
typedef struct AVIOContext { int unused; } AVIOContext;
typedef struct AVOutputFormat { int flags; } AVOutputFormat;
typedef struct SwsContext SwsContext;
typedef struct SwsFilter SwsFilter;
typedef struct AVDictionary AVDictionary;

typedef struct AVCodecContext {
    int width, height;
    AVRational time_base, framerate;
    enum AVPixelFormat pix_fmt;
    int gop_size, max_b_frames;
    enum AVSampleFormat sample_fmt;
    int64_t bit_rate;
    int sample_rate, channels;
    uint64_t channel_layout;
    int frame_size;
} AVCodecContext;

typedef struct AVStream {
    int index;
    AVCodecParameters* codecpar;
    AVRational time_base;
} AVStream;

typedef struct AVFormatContext {
    const AVOutputFormat* oformat;
    AVIOContext* pb;
} AVFormatContext;

typedef struct AVFrame {
    uint8_t* data[8];
    int linesize[8];
    int width, height, nb_samples, format;
    int64_t pts;
    int sample_rate, channels;
    uint64_t channel_layout;
} AVFrame;

typedef struct AVPacket { int stream_index; int64_t pts, dts; } AVPacket;

int av_strerror(int errnum, char* errbuf, size_t errbuf_size);
AVFormatContext* avformat_alloc_context(void);
const AVOutputFormat* av_guess_format(const char* short_name, const char* filename, const char* mime_type);
const AVCodec* avcodec_find_encoder(enum AVCodecID id);
AVStream* avformat_new_stream(AVFormatContext* s, const AVCodec* c);
AVCodecContext* avcodec_alloc_context3(const AVCodec* codec);
int avcodec_parameters_from_context(AVCodecParameters* par, const AVCodecContext* codec);
int avcodec_open2(AVCodecContext* avctx, const AVCodec* codec, AVDictionary** options);
AVFrame* av_frame_alloc(void);
int av_frame_get_buffer(AVFrame* frame, int align);
AVPacket* av_packet_alloc(void);
int64_t av_get_default_channel_layout(int nb_channels);
int avio_open(AVIOContext** s, const char* url, int flags);
int avformat_write_header(AVFormatContext* s, AVDictionary** options);
void av_frame_free(AVFrame** frame);
int avcodec_send_frame(AVCodecContext* avctx, const AVFrame* frame);
int avcodec_receive_packet(AVCodecContext* avctx, AVPacket* avpkt);
void av_packet_rescale_ts(AVPacket* pkt, AVRational tb_src, AVRational tb_dst);
int av_interleaved_write_frame(AVFormatContext* s, AVPacket* pkt);
void av_packet_unref(AVPacket* pkt);
int av_write_trailer(AVFormatContext* s);
int avio_closep(AVIOContext** s);
void avcodec_free_context(AVCodecContext** avctx);
void av_packet_free(AVPacket** pkt);
void sws_freeContext(SwsContext* swsContext);
void avformat_free_context(AVFormatContext* s);
SwsContext* sws_getContext(int srcW, int srcH, enum AVPixelFormat srcFormat, int dstW, int dstH,
                           enum AVPixelFormat dstFormat, int flags, SwsFilter* srcFilter,
                           SwsFilter* dstFilter, const double* param);
int av_image_alloc(uint8_t* pointers[4], int linesizes[4], int w, int h, enum AVPixelFormat pix_fmt, int align);
int sws_scale(SwsContext* c, const uint8_t* const srcSlice[], const int srcStride[], int srcSliceY,
              int srcSliceH, uint8_t* const dst[], const int dstStride[]);
void av_freep(void* ptr);
#endif
