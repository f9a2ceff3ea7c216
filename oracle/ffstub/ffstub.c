/* TEST INFRASTRUCTURE ONLY: abort()-stubs for the FFmpeg symbols the unmodified
 * reference links against (ctypes loads with RTLD_NOW, so they must resolve).
 * The draw/composite path never reaches any of them. */
#include "ffstub_common.h"
#define DIE() do { fprintf(stderr, "ffstub: %s called (FFmpeg is not available here)\n", __func__); abort(); } while (0)
int av_strerror(int a, char* b, size_t c) { (void)a; (void)b; (void)c; DIE(); }
AVFormatContext* avformat_alloc_context(void) { DIE(); }
const AVOutputFormat* av_guess_format(const char* a, const char* b, const char* c) { (void)a; (void)b; (void)c; DIE(); }
const AVCodec* avcodec_find_encoder(enum AVCodecID id) { (void)id; DIE(); }
AVStream* avformat_new_stream(AVFormatContext* s, const AVCodec* c) { (void)s; (void)c; DIE(); }
AVCodecContext* avcodec_alloc_context3(const AVCodec* c) { (void)c; DIE(); }
int avcodec_parameters_from_context(AVCodecParameters* p, const AVCodecContext* c) { (void)p; (void)c; DIE(); }
int avcodec_open2(AVCodecContext* a, const AVCodec* c, AVDictionary** o) { (void)a; (void)c; (void)o; DIE(); }
AVFrame* av_frame_alloc(void) { DIE(); }
int av_frame_get_buffer(AVFrame* f, int a) { (void)f; (void)a; DIE(); }
AVPacket* av_packet_alloc(void) { DIE(); }
int64_t av_get_default_channel_layout(int n) { (void)n; DIE(); }
int avio_open(AVIOContext** s, const char* u, int f) { (void)s; (void)u; (void)f; DIE(); }
int avformat_write_header(AVFormatContext* s, AVDictionary** o) { (void)s; (void)o; DIE(); }
void av_frame_free(AVFrame** f) { (void)f; DIE(); }
int avcodec_send_frame(AVCodecContext* a, const AVFrame* f) { (void)a; (void)f; DIE(); }
int avcodec_receive_packet(AVCodecContext* a, AVPacket* p) { (void)a; (void)p; DIE(); }
void av_packet_rescale_ts(AVPacket* p, AVRational a, AVRational b) { (void)p; (void)a; (void)b; DIE(); }
int av_interleaved_write_frame(AVFormatContext* s, AVPacket* p) { (void)s; (void)p; DIE(); }
void av_packet_unref(AVPacket* p) { (void)p; DIE(); }
int av_write_trailer(AVFormatContext* s) { (void)s; DIE(); }
int avio_closep(AVIOContext** s) { (void)s; DIE(); }
void avcodec_free_context(AVCodecContext** a) { (void)a; DIE(); }
void av_packet_free(AVPacket** p) { (void)p; DIE(); }
void sws_freeContext(SwsContext* c) { (void)c; DIE(); }
void avformat_free_context(AVFormatContext* s) { (void)s; DIE(); }
SwsContext* sws_getContext(int a, int b, enum AVPixelFormat c, int d, int e, enum AVPixelFormat f, int g,
                           SwsFilter* h, SwsFilter* i, const double* j) {
    (void)a; (void)b; (void)c; (void)d; (void)e; (void)f; (void)g; (void)h; (void)i; (void)j; DIE();
}
int av_image_alloc(uint8_t* p[4], int l[4], int w, int h, enum AVPixelFormat f, int a) {
    (void)p; (void)l; (void)w; (void)h; (void)f; (void)a; DIE();
}
int sws_scale(SwsContext* c, const uint8_t* const s[], const int ss[], int y, int h, uint8_t* const d[], const int ds[]) {
    (void)c; (void)s; (void)ss; (void)y; (void)h; (void)d; (void)ds; DIE();
}
void av_freep(void* p) { (void)p; DIE(); }
