// cvs_audio.cpp -- composite_audio_process() of ffmpeg_ntsc (ffmpeg_ntsc.cpp:901-970), host side.
//
// the audio path stays on the CPU (a few one-pole filters per sample at 44.1 kHz) but shares
// the libc rand() stream with the video path: the tape hiss draws one rand() per sample and channel (:951-952), so a
// host that interleaves audio packets and fields like the reference's main loop must hand the stream position back
// and forth (cvs_rng_tell / cvs_rng_seek on the video context, the in/out rng_pos argument here).
// No CUDA in this file; the filters are the reference's LowpassFilter (:74-106) evaluated in the same operation order.
#include <cmath>
#include <cstdint>
#include <new>
#include <vector>

#include "../../include/cvs_ntsc.h"
#include "glibc_rand.h"

namespace {

struct OnePole {                       // LowpassFilter, ffmpeg_ntsc.cpp:74-106
    double alpha = 0, prev = 0;
    void set(double rate, double hz) {                 // :78-86
        const double timeInterval = 1.0 / rate;
        const double tau = 1 / (hz * 2 * M_PI);
        alpha = timeInterval / (tau + timeInterval);
        prev = 0;
    }
    double lowpass(double s) {                         // :90-94
        const double stage1 = s * alpha;
        const double stage2 = prev - (prev * alpha);
        return (prev = (stage1 + stage2));
    }
    double highpass(double s) {                        // :95-99
        const double stage1 = s * alpha;
        const double stage2 = prev - (prev * alpha);
        return s - (prev = (stage1 + stage2));
    }
};

double dBFS(double dB) { return std::pow(10.0, dB / 20.0); }               // :51-58

int clips16(int x) { return x < -32768 ? -32768 : (x > 32767 ? 32767 : x); }   // :892-899

}  // namespace

struct cvs_audio {
    cvs_params p;
    int channels = 2;                  // output_audio_channels after parse_argv (:1227-1262)
    int rate = 44100;                  // output_audio_rate (:212)
    double highpass = 20, lowpass = 20000;
    int hiss_level = 0;                // :1267
    std::vector<std::vector<OnePole>> lo, hi;   // audio_hilopass: [channel][pass], 6 passes (:2032-2036)
    OnePole pre[2], post[2], boost[2]; // :753-754, :890
    unsigned long long proc_count = 0; // audio_proc_count (:889)
    cvs::RandCursor cur;
};

extern "C" {

int cvs_audio_channels(const cvs_params *p) {
    if (!p) return CVS_ERR_INVALID_ARG;
    // :1227-1262 (output_vhs_linear_stereo has no switch and stays false)
    if (p->emulating_vhs && !p->output_vhs_hifi && p->output_vhs_linear_audio) return 1;
    return 2;
}

int cvs_audio_create(cvs_audio **out, const cvs_params *p) {
    if (!out || !p || p->struct_size != (int32_t)sizeof(cvs_params)) return CVS_ERR_INVALID_ARG;
    cvs_audio *a = new (std::nothrow) cvs_audio();
    if (!a) return CVS_ERR_NOMEM;
    a->p = *p;
    // end of parse_argv(), :1227-1262
    a->highpass = 20; a->lowpass = 20000; a->channels = 2;
    if (p->emulating_vhs && !p->output_vhs_hifi && p->output_vhs_linear_audio) {
        a->highpass = 100;
        a->lowpass = p->output_vhs_tape_speed == CVS_VHS_SP ? 10000 : (p->output_vhs_tape_speed == CVS_VHS_LP ? 7000 : 4000);
        a->channels = 1;
    }
    a->hiss_level = (int)(dBFS(p->output_audio_hiss_db) * 5000);           // :1267 (int = double, truncating)
    // "prepare audio filtering", :2031-2066: setCutoff(output_audio_lowpass, output_audio_highpass) (:2034) ->
    // HiLoPair::setFilter(rate, low_hz, high_hz) (:112-115): lo runs at output_audio_lowpass, hi at output_audio_highpass
    a->lo.assign((size_t)a->channels, std::vector<OnePole>(6));
    a->hi.assign((size_t)a->channels, std::vector<OnePole>(6));
    for (int c = 0; c < a->channels; c++)
        for (int i = 0; i < 6; i++) {
            a->lo[(size_t)c][(size_t)i].set(a->rate, a->lowpass);
            a->hi[(size_t)c][(size_t)i].set(a->rate, a->highpass);
        }
    for (int i = 0; i < 2; i++) a->boost[i].set(a->rate, 10000);           // :2039-2040
    const double emph = p->output_vhs_hifi ? 16000 : 8000;                 // :2044-2066
    if (p->emulating_preemphasis) for (int i = 0; i < a->channels; i++) a->pre[i].set(a->rate, emph);
    if (p->emulating_deemphasis) for (int i = 0; i < a->channels; i++) a->post[i].set(a->rate, emph);
    *out = a;
    return CVS_OK;
}

void cvs_audio_destroy(cvs_audio *a) { delete a; }

int cvs_audio_process(cvs_audio *a, int16_t *audio, unsigned samples, unsigned long long *rng_pos) {
    if (!a || (!audio && samples) || !rng_pos) return CVS_ERR_INVALID_ARG;
    if (!a->p.enable_audio_emulation) return CVS_OK;                       // process_audio(), :1288-1289
    if (a->cur.pos() != *rng_pos) a->cur.seek(*rng_pos);
    const cvs_params &p = a->p;
    const int ch = a->channels;
    const double linear_buzz = dBFS(p.output_audio_linear_buzz);           // :903
    const double hsync_hz = p.output_ntsc ? 15734 : 15625;                 // :904
    const int vsync_lines = p.output_ntsc ? 525 : 625;
    const int vpulse_end = p.output_ntsc ? 10 : 12;
    const double hpulse_end = p.output_ntsc ? (hsync_hz * (4.7 / 1000000)) : (hsync_hz * (4.0 / 1000000));
    for (unsigned n = 0; n < samples; n++, audio += ch) {
        for (int c = 0; c < ch; c++) {
            double s = (double)audio[c] / 32768;                           // :913
            // HiLoPass::filter: all lowpasses, then all highpasses (:126-130)
            for (int i = 0; i < 6; i++) s = a->lo[(size_t)c][(size_t)i].lowpass(s);
            for (int i = 0; i < 6; i++) s = a->hi[(size_t)c][(size_t)i].highpass(s);
            if (p.emulating_preemphasis)                                   // :919-923: every channel's filter, on each channel
                for (int i = 0; i < ch; i++) s = s + a->pre[i].highpass(s);
            if (!p.output_vhs_hifi && linear_buzz > 0.000000001) {         // :926-945
                const unsigned oversample = 16;
                for (unsigned oi = 0; oi < oversample; oi++) {
                    const double t = ((((double)a->proc_count * oversample) + oi) * hsync_hz) / a->rate / oversample;
                    const double hpos = std::fmod(t, 1.0);
                    const int vline = (int)std::fmod(std::floor(t + 0.0001 - hpos), (double)vsync_lines / 2);
                    bool pulse = false;
                    if (hpos < hpulse_end) pulse = true;
                    if (vline < vpulse_end) pulse = true;
                    if (pulse) s -= linear_buzz / oversample / 2;
                }
            }
            if (s > 1.0) s = 1.0;                                          // :948-951
            else if (s < -1.0) s = -1.0;
            if (a->hiss_level != 0)                                        // :952-953
                s += ((double)(((int)(a->cur.next() % (unsigned)((a->hiss_level * 2) + 1))) - a->hiss_level)) / 20000;
            if (!p.output_vhs_hifi && p.vhs_linear_high_boost > 0)         // :955-957
                s += a->boost[c].highpass(s) * p.vhs_linear_high_boost;
            if (p.emulating_deemphasis)                                    // :959-963
                for (int i = 0; i < ch; i++) s = a->post[i].lowpass(s);
            audio[c] = (int16_t)clips16((int)(s * 32768));                 // :965
        }
        a->proc_count++;
    }
    *rng_pos = a->cur.pos();
    return CVS_OK;
}

}  // extern "C"
