// cvs_ffmpeg_ntsc -- the reference program with the B200 engine in the middle: FFmpeg demux / decode -> GPU -> FFmpeg
// encode / mux.  Same command line as ffmpeg_ntsc (ffmpeg_ntsc.cpp main(), :1923-2331; parse_argv(), :972-1282).
//
//   reference (per output field)                                 here
//   ------------------------------------------------------------  -------------------------------------------------------
//   InputFile::next_packet() / decode           :229-543          Input::read(): av_read_frame + avcodec_send_packet /
//                                                                 avcodec_receive_frame (the decode API of FFmpeg >= 4)
//   frame_copy_scale(): sws_scale -> BGRA        :544-613          +
//   composite_layer()                            :2229             |  ONE call per batch of fields: cvs_field_loop_host()
//   line doubling                                :2232-2257        |  (decoder pictures up, encoder YUV down; the
//   sws_scale BGRA -> YUV420P / YUV422P          :2266-2274        +  conversions give libswscale's bytes, include/cvs_ntsc.h)
//   output_frame(): H.264 encode + mux           :1342-1373        Output::video(): avcodec_send_frame / receive_packet
//   process_audio() + write_out_audio()          :1284-1340        cvs_audio_process() on S16 / 44.1 kHz (swr_convert as
//                                                                 InputFile does, :640-714), PCM_S16LE packets
//
// Scheduling is the reference's (:2146-2283): audio packets are processed in demux order; a decoded picture is shown from
// the field its pts names until the next picture's pts; `current` counts output fields at output_field_rate.  Audio
// hiss and video noise draw from ONE rand() stream (:952, :1640), so the fields gathered so far are flushed to the GPU
// before an audio packet is filtered, and the stream position travels between the two engines (cvs_rng_tell /
// cvs_audio_process / cvs_rng_seek).  One input file (the reference composites several with keying, :2213-2230; the C
// ABI takes one layer per call).
//
// This file needs FFmpeg's development headers (libavformat, libavcodec, libavutil, libswresample, libswscale for
// decoder formats other than YUV420P / YUV422P / NV12 / BGRA).  Where they are absent -- as in the image this
// repository is developed in -- it compiles to a stub that says so; `make ffmpeg-host` in csrc/ builds the real thing
// (pkg-config), and tests/test_ffmpeg_host_syntax.py compiles this source against tests/ffmpeg_decl/ (declarations of
// the API subset used here, written for that check; not FFmpeg's headers); the field schedule lives in tools/field_schedule.h
// and is checked on the CPU (tests/test_field_schedule.py).  The raw-frame host tools/cvs_ntsc_raw.cpp is the one exercised
// on the GPU box.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/cvs_ntsc.h"
#include "field_schedule.h"

#if defined(CVS_WITH_FFMPEG) || __has_include(<libavformat/avformat.h>)
extern "C" {
#include <libavcodec/avcodec.h>
#include <libavformat/avformat.h>
#include <libavutil/channel_layout.h>
#include <libavutil/opt.h>
#include <libswresample/swresample.h>
#include <libswscale/swscale.h>
}

namespace {

[[noreturn]] void fail(const char *what, int rc) {
    fprintf(stderr, "%s: %s\n", what, cvs_strerror(rc));
    exit(1);
}
[[noreturn]] void fail_av(const char *what) {
    fprintf(stderr, "%s\n", what);
    exit(1);
}

int pix_of(int avfmt) {
    switch (avfmt) {
    case AV_PIX_FMT_YUV420P: case AV_PIX_FMT_YUVJ420P: return CVS_PIX_YUV420P;
    case AV_PIX_FMT_YUV422P: case AV_PIX_FMT_YUVJ422P: return CVS_PIX_YUV422P;
    case AV_PIX_FMT_NV12: return CVS_PIX_NV12;
    case AV_PIX_FMT_BGRA: return CVS_PIX_BGRA;
    default: return -1;
    }
}

// Decoded pictures of one batch, tightly packed in page-locked memory: plane p of picture k at base[p] + k * stride[p].
struct PictureStore {
    cvs_ctx *ctx = nullptr;
    int fmt = -1, w = 0, h = 0, cap = 0, n = 0;
    uint8_t *base[3] = {nullptr, nullptr, nullptr};
    int linesize[3] = {0, 0, 0};
    long long stride[3] = {0, 0, 0};
    int rows[3] = {0, 0, 0};

    void shape(cvs_ctx *c, int f, int w_, int h_, int cap_) {
        release();
        ctx = c; fmt = f; w = w_; h = h_; cap = cap_; n = 0;
        const int cw = (w + 1) / 2, ch = f == CVS_PIX_YUV422P ? h : (h + 1) / 2;
        const int planes = f == CVS_PIX_BGRA ? 1 : (f == CVS_PIX_NV12 ? 2 : 3);
        for (int p = 0; p < planes; p++) {
            linesize[p] = f == CVS_PIX_BGRA ? 4 * w : (p == 0 ? w : (f == CVS_PIX_NV12 ? 2 * cw : cw));
            rows[p] = p == 0 ? h : ch;
            stride[p] = (long long)linesize[p] * rows[p];
            void *m = nullptr;
            const int rc = cvs_alloc_host(ctx, &m, (size_t)stride[p] * (size_t)cap);
            if (rc != CVS_OK) fail("cvs_alloc_host", rc);
            base[p] = (uint8_t *)m;
        }
    }
    void release() {
        for (uint8_t *&b : base) { if (b) cvs_free_host(ctx, b); b = nullptr; }
    }
    int add(uint8_t *const data[], const int ls[]) {          // copies one decoded picture in, returns its index
        for (int p = 0; p < 3 && base[p]; p++)
            for (int r = 0; r < rows[p]; r++)
                memcpy(base[p] + (size_t)n * (size_t)stride[p] + (size_t)r * (size_t)linesize[p], data[p] + (size_t)r * (size_t)ls[p], (size_t)linesize[p]);
        return n++;
    }
};

struct Input {
    AVFormatContext *fmt = nullptr;
    AVCodecContext *vdec = nullptr, *adec = nullptr;
    int vidx = -1, aidx = -1;
    AVPacket *pkt = nullptr;
    AVFrame *frame = nullptr;
    SwrContext *swr = nullptr;
    SwsContext *to_bgra = nullptr;        // decoder formats the engine does not take: the reference's own conversion
    AVFrame *bgra = nullptr;

    void open(const char *path) {
        if (avformat_open_input(&fmt, path, nullptr, nullptr) < 0) fail_av("Failed to open input");          // :262
        if (avformat_find_stream_info(fmt, nullptr) < 0) fail_av("Failed to probe input");                  // :267
        const AVCodec *vc = nullptr, *ac = nullptr;
        vidx = av_find_best_stream(fmt, AVMEDIA_TYPE_VIDEO, -1, -1, &vc, 0);
        aidx = av_find_best_stream(fmt, AVMEDIA_TYPE_AUDIO, -1, vidx, &ac, 0);
        if (vidx < 0) fail_av("No video stream");
        vdec = avcodec_alloc_context3(vc);
        avcodec_parameters_to_context(vdec, fmt->streams[vidx]->codecpar);
        if (avcodec_open2(vdec, vc, nullptr) < 0) fail_av("Cannot open the video decoder");
        if (aidx >= 0) {
            adec = avcodec_alloc_context3(ac);
            avcodec_parameters_to_context(adec, fmt->streams[aidx]->codecpar);
            if (avcodec_open2(adec, ac, nullptr) < 0) { avcodec_free_context(&adec); aidx = -1; }
        }
        pkt = av_packet_alloc();
        frame = av_frame_alloc();
    }
    void close() {
        if (swr) swr_free(&swr);
        if (to_bgra) sws_freeContext(to_bgra);
        if (bgra) av_frame_free(&bgra);
        av_frame_free(&frame);
        av_packet_free(&pkt);
        if (adec) avcodec_free_context(&adec);
        avcodec_free_context(&vdec);
        avformat_close_input(&fmt);
    }
};

struct Output {
    AVFormatContext *fmt = nullptr;
    AVStream *vst = nullptr, *ast = nullptr;
    AVCodecContext *venc = nullptr, *aenc = nullptr;
    AVFrame *vframe = nullptr;
    AVPacket *pkt = nullptr;
    long long audio_pts = 0;

    void open(const char *path, const cvs_params &p, int w, int h, bool yuv422, int channels) {
        if (avformat_alloc_output_context2(&fmt, nullptr, nullptr, path) < 0) fail_av("Failed to open output file");   // :1942
        const AVCodec *ac = avcodec_find_encoder(AV_CODEC_ID_PCM_S16LE);                                      // :1973
        ast = avformat_new_stream(fmt, nullptr);
        aenc = avcodec_alloc_context3(ac);
        av_channel_layout_default(&aenc->ch_layout, channels);                                                 // :1960-1963
        aenc->sample_rate = 44100;                                                                             // output_audio_rate, :212
        aenc->sample_fmt = AV_SAMPLE_FMT_S16;
        aenc->time_base = AVRational{1, 44100};
        if (fmt->oformat->flags & AVFMT_GLOBALHEADER) aenc->flags |= AV_CODEC_FLAG_GLOBAL_HEADER;
        if (avcodec_open2(aenc, ac, nullptr) < 0) fail_av("Output stream cannot open codec");
        avcodec_parameters_from_context(ast->codecpar, aenc);
        ast->time_base = aenc->time_base;

        const AVCodec *vc = avcodec_find_encoder(AV_CODEC_ID_H264);                                            // :1996
        if (!vc) fail_av("No H.264 encoder in this FFmpeg build");
        vst = avformat_new_stream(fmt, nullptr);
        venc = avcodec_alloc_context3(vc);
        venc->width = w;
        venc->height = h;
        venc->sample_aspect_ratio = AVRational{h * 4, w * 3};                                                  // :1999
        venc->pix_fmt = yuv422 ? AV_PIX_FMT_YUV422P : AV_PIX_FMT_YUV420P;                                      // :2000
        venc->gop_size = 15;
        venc->max_b_frames = 0;
        venc->time_base = p.output_ntsc ? AVRational{1001, 60000} : AVRational{1, 50};                         // field rate, :2003
        venc->colorspace = AVCOL_SPC_SMPTE170M;                                                                // :2100
        venc->color_range = AVCOL_RANGE_MPEG;                                                                  // :2101
        if (fmt->oformat->flags & AVFMT_GLOBALHEADER) venc->flags |= AV_CODEC_FLAG_GLOBAL_HEADER;
        if (avcodec_open2(venc, vc, nullptr) < 0) fail_av("Output stream cannot open codec");
        avcodec_parameters_from_context(vst->codecpar, venc);
        vst->time_base = venc->time_base;
        if (!(fmt->oformat->flags & AVFMT_NOFILE) && avio_open(&fmt->pb, path, AVIO_FLAG_WRITE) < 0) fail_av("Output file cannot open file");
        if (avformat_write_header(fmt, nullptr) < 0) fail_av("Failed to write header");
        vframe = av_frame_alloc();
        vframe->format = venc->pix_fmt;
        vframe->width = w;
        vframe->height = h;
        pkt = av_packet_alloc();
    }
    void drain(AVCodecContext *enc, AVStream *st) {
        while (avcodec_receive_packet(enc, pkt) == 0) {
            pkt->stream_index = st->index;
            av_packet_rescale_ts(pkt, enc->time_base, st->time_base);
            if (av_interleaved_write_frame(fmt, pkt) < 0) fprintf(stderr, "AV write frame failed\n");
            av_packet_unref(pkt);
        }
    }
    // one finished picture: planes point into the batch's page-locked output (no copy)
    void video(uint8_t *y, uint8_t *u, uint8_t *v, int ly, int lc, unsigned long long field) {
        vframe->data[0] = y; vframe->data[1] = u; vframe->data[2] = v;
        vframe->linesize[0] = ly; vframe->linesize[1] = lc; vframe->linesize[2] = lc;
        vframe->pts = (long long)field;                                                                        // :1346
        if (avcodec_send_frame(venc, vframe) < 0) fprintf(stderr, "encode failed\n");
        drain(venc, vst);
        fprintf(stderr, "\rOutput field %llu ", field);                                                        // :1361
    }
    void audio(const int16_t *pcm, int frames, int channels) {                                                 // write_out_audio(), :1292-1340
        AVFrame *f = av_frame_alloc();
        f->format = AV_SAMPLE_FMT_S16;
        f->nb_samples = frames;
        f->sample_rate = 44100;
        av_channel_layout_default(&f->ch_layout, channels);
        if (av_frame_get_buffer(f, 0) < 0) fail_av("audio frame");
        memcpy(f->data[0], pcm, (size_t)frames * (size_t)channels * 2);
        f->pts = audio_pts;
        audio_pts += frames;
        if (avcodec_send_frame(aenc, f) < 0) fprintf(stderr, "audio encode failed\n");
        av_frame_free(&f);
        drain(aenc, ast);
    }
    void close() {
        avcodec_send_frame(venc, nullptr);                                                                     // flush encoder delay, :2286-2306
        drain(venc, vst);
        av_write_trailer(fmt);
        if (!(fmt->oformat->flags & AVFMT_NOFILE)) avio_closep(&fmt->pb);
        av_frame_free(&vframe);
        av_packet_free(&pkt);
        avcodec_free_context(&venc);
        avcodec_free_context(&aenc);
        avformat_free_context(fmt);
    }
};

}  // namespace

int main(int argc, char **argv) {
    std::string in_path, out_path;
    int batch = 32, use_double = 0, fast_noise = 0;
    std::vector<const char *> ref_argv;
    ref_argv.push_back(argv[0]);
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i], *n = a;
        while (*n == '-') n++;
        auto need = [&](const char *what) -> const char * {
            if (i + 1 >= argc) { fprintf(stderr, "%s needs a value\n", what); exit(1); }
            return argv[++i];
        };
        if (a[0] == '-' && !strcmp(n, "i")) { in_path = need("-i"); ref_argv.push_back("-i"); ref_argv.push_back(in_path.c_str()); }
        else if (a[0] == '-' && !strcmp(n, "o")) { out_path = need("-o"); ref_argv.push_back("-o"); ref_argv.push_back(out_path.c_str()); }
        else if (a[0] == '-' && !strcmp(n, "batch")) batch = atoi(need("-batch"));
        else if (a[0] == '-' && !strcmp(n, "double")) use_double = 1;
        else if (a[0] == '-' && !strcmp(n, "fast-noise")) fast_noise = 1;
        else ref_argv.push_back(a);
    }
    cvs_params p;
    cvs_params_default_ntsc(&p);
    int rc = cvs_params_apply_argv(&p, (int)ref_argv.size(), ref_argv.data());
    if (rc != CVS_OK) { fprintf(stderr, "%s: %s\n", argv[0], cvs_strerror(rc)); return 1; }                    // :1925
    if (in_path.empty()) { fprintf(stderr, "No input files specified\n"); return 1; }                          // :1276
    if (out_path.empty()) { fprintf(stderr, "No output file specified\n"); return 1; }                         // :1272
    if (p.output_frame_delay > 1) { fprintf(stderr, "-d > 1 is only available in cvs_ntsc_raw\n"); return 1; } // cvs_field_loop_host keeps a ring of one
    const int w = p.output_width, h = p.output_height;
    const bool yuv422 = p.use_422_colorspace != 0;
    if (batch < 1 || (w & 1)) { fprintf(stderr, "bad geometry (the encoder needs an even width)\n"); return 1; }

    Input in;
    in.open(in_path.c_str());
    cvs_ctx *ctx = nullptr;
    if ((rc = cvs_create(&ctx, &p, 0, w, h, batch)) != CVS_OK) fail("cvs_create", rc);
    cvs_set_precision(ctx, use_double);
    cvs_set_noise_mode(ctx, fast_noise ? CVS_NOISE_FAST : CVS_NOISE_EXACT);
    const int channels = cvs_audio_channels(&p);
    cvs_audio *fx = nullptr;
    if (in.aidx >= 0 && (rc = cvs_audio_create(&fx, &p)) != CVS_OK) fail("cvs_audio_create", rc);
    Output out;
    out.open(out_path.c_str(), p, w, h, yuv422, channels);

    // output pictures of a batch, page-locked: Y then U then V per picture
    const int cw = w / 2, ch = yuv422 ? h : (h + 1) / 2;
    const size_t ysz = (size_t)w * h, csz = (size_t)cw * ch, opic = ysz + 2 * csz;
    void *obuf = nullptr;
    if ((rc = cvs_alloc_host(ctx, &obuf, opic * (size_t)batch)) != CVS_OK) fail("cvs_alloc_host", rc);
    uint8_t *const o = (uint8_t *)obuf;

    PictureStore store;
    FieldSchedule sched(batch);                   // which stored picture every output field shows (tools/field_schedule.h)
    const AVRational field_tb = p.output_ntsc ? AVRational{1001, 60000} : AVRational{1, 50};

    auto flush = [&]() {                          // the gathered fields through the GPU, then to the encoder
        const int n = (int)sched.src_of_field.size();
        if (n > 0) {
            cvs_field_loop d;
            memset(&d, 0, sizeof d);
            d.struct_size = (int32_t)sizeof d;
            d.src_format = store.fmt; d.src_w = store.w; d.src_h = store.h;
            for (int k = 0; k < 3; k++) { d.src[k] = store.base[k]; d.src_linesize[k] = store.linesize[k]; d.src_pic_stride[k] = store.stride[k]; }
            d.nsrc = sched.stored;
            d.src_of_field = sched.src_of_field.data();
            d.w = w; d.h = h;
            d.out_format = yuv422 ? CVS_YUV422P : CVS_YUV420P;
            d.y = o; d.u = o + ysz; d.v = o + ysz + csz;
            d.ly = w; d.lu = cw; d.lv = cw;
            d.y_pic_stride = d.u_pic_stride = d.v_pic_stride = (long long)opic;
            const int rc2 = cvs_field_loop_host(ctx, &d, n, (unsigned long long)sched.current);
            if (rc2 != CVS_OK) fail("cvs_field_loop_host", rc2);
            for (int k = 0; k < n; k++)
                out.video(o + (size_t)k * opic, o + (size_t)k * opic + ysz, o + (size_t)k * opic + ysz + csz, w, cw,
                          (unsigned long long)sched.current + (unsigned long long)k);
        }
        // the picture still on show stays in the store as picture 0
        if (sched.stored > 1)
            for (int pl = 0; pl < 3 && store.base[pl]; pl++)
                memmove(store.base[pl], store.base[pl] + (size_t)(sched.stored - 1) * (size_t)store.stride[pl], (size_t)store.stride[pl]);
        sched.flushed();
        store.n = sched.stored;
    };
    // the zeroed frame the reference starts with (:556-561), in the decoder's format
    auto black = [&](int slot) {
        for (int pl = 0; pl < 3 && store.base[pl]; pl++)
            memset(store.base[pl] + (size_t)slot * (size_t)store.stride[pl], (store.fmt == CVS_PIX_BGRA) ? 0 : (pl == 0 ? 16 : 128), (size_t)store.stride[pl]);
    };

    std::vector<int16_t> pcm;
    bool eof = false;
    while (!eof) {
        const int r = av_read_frame(in.fmt, in.pkt);
        AVCodecContext *dec = nullptr;
        if (r < 0) { eof = true; avcodec_send_packet(in.vdec, nullptr); dec = in.vdec; }
        else if (in.pkt->stream_index == in.vidx) { avcodec_send_packet(in.vdec, in.pkt); dec = in.vdec; }
        else if (in.pkt->stream_index == in.aidx && fx) { avcodec_send_packet(in.adec, in.pkt); dec = in.adec; }
        if (r >= 0) av_packet_unref(in.pkt);
        while (dec && avcodec_receive_frame(dec, in.frame) == 0) {
            AVFrame *f = in.frame;
            if (dec == in.adec) {
                // audio in demux order: whatever was gathered so far draws first (one rand() stream, :952 / :1640)
                if (!sched.src_of_field.empty()) flush();
                if (!in.swr) {
                    AVChannelLayout lay;
                    av_channel_layout_default(&lay, channels);
                    swr_alloc_set_opts2(&in.swr, &lay, AV_SAMPLE_FMT_S16, 44100, &f->ch_layout, (AVSampleFormat)f->format, f->sample_rate, 0, nullptr);
                    if (!in.swr || swr_init(in.swr) < 0) fail_av("audio resampler");
                }
                const int cap = (int)av_rescale_rnd(swr_get_delay(in.swr, f->sample_rate) + f->nb_samples, 44100, f->sample_rate, AV_ROUND_UP);
                pcm.resize((size_t)cap * (size_t)channels);
                uint8_t *dstp[1] = {(uint8_t *)pcm.data()};
                const int got = swr_convert(in.swr, dstp, cap, (const uint8_t **)f->extended_data, f->nb_samples);
                if (got > 0) {
                    unsigned long long pos = 0;
                    cvs_rng_tell(ctx, &pos);
                    if ((rc = cvs_audio_process(fx, pcm.data(), (unsigned)got, &pos)) != CVS_OK) fail("cvs_audio_process", rc);   // :1284-1290
                    cvs_rng_seek(ctx, pos);
                    out.audio(pcm.data(), got, channels);
                }
                continue;
            }
            // a decoded picture: its pts in output fields (:2170-2176)
            long long at = -1;
            const long long pts = f->best_effort_timestamp;
            if (pts != AV_NOPTS_VALUE) at = av_rescale_q(pts, in.fmt->streams[in.vidx]->time_base, field_tb);
            int fmt = pix_of(f->format);
            uint8_t *const *data = f->data;
            const int *ls = f->linesize;
            int sw = f->width, sh = f->height;
            if (fmt < 0) {                        // e.g. 10-bit or 4:4:4 sources: the reference's conversion on the host
                if (!in.to_bgra) {
                    in.to_bgra = sws_getContext(sw, sh, (AVPixelFormat)f->format, w, h, AV_PIX_FMT_BGRA, SWS_BILINEAR, nullptr, nullptr, nullptr);   // :574-585
                    in.bgra = av_frame_alloc();
                    in.bgra->format = AV_PIX_FMT_BGRA; in.bgra->width = w; in.bgra->height = h;
                    if (!in.to_bgra || av_frame_get_buffer(in.bgra, 64) < 0) fail_av("sws_getContext fail");
                }
                sws_scale(in.to_bgra, f->data, f->linesize, 0, sh, in.bgra->data, in.bgra->linesize);          // :603-610
                fmt = CVS_PIX_BGRA; data = in.bgra->data; ls = in.bgra->linesize; sw = w; sh = h;
            }
            if (store.fmt != fmt || store.w != sw || store.h != sh) {       // first picture, or the stream changed shape (:565-572)
                // what is on show belongs to the old shape: it ends here
                if (sched.stored > 0) { sched.show_until(at < 0 ? sched.next_field() : at, flush); flush(); sched.stored = 0; }
                store.shape(ctx, fmt, sw, sh, batch + 1);
            }
            const int slot = sched.picture(at, flush, black);
            store.n = slot;
            store.add(data, ls);
        }
    }
    sched.finish(flush);                          // the last picture: one frame = two fields
    fprintf(stderr, "\n");
    out.close();
    in.close();
    store.release();
    if (fx) cvs_audio_destroy(fx);
    cvs_free_host(ctx, obuf);
    cvs_destroy(ctx);
    return 0;
}

#else   // no FFmpeg development headers on this machine

int main(int, char **argv) {
    fprintf(stderr,
            "%s was built without FFmpeg's development headers (libavformat/avformat.h not found).\n"
            "Use cvs_ntsc_raw with ffmpeg on either side (see its header comment), or rebuild with `make ffmpeg-host`\n"
            "on a machine that has libavformat / libavcodec / libswresample / libswscale.\n", argv[0]);
    return 2;
}

#endif
