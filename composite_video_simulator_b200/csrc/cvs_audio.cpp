// cvs_audio.cpp -- the audio step of ffmpeg_ntsc's field loop (composite_audio_process(), ffmpeg_ntsc.cpp:901-970),
// host side, as a packet pipeline.
//
// audio stays on the CPU (a dozen one-pole filters per sample at 44.1 kHz) but shares the libc
// rand() stream with the video path: the tape hiss draws one rand() per sample and channel (:951-952), so a host
// that interleaves audio packets and fields like the reference's main loop hands the stream position back and forth
// (cvs_rng_tell / cvs_rng_seek on the video context, the in/out rng_pos argument here).
//
// Structure (not the reference's): a packet is converted once into a work buffer of doubles and then swept stage
// by stage -- band limiting, pre-emphasis, sync buzz, limiter, hiss, high boost, de-emphasis, quantisation -- each
// stage an object that owns its filter bank.  Every stage's state depends only on its own input sequence, so the
// sweep order gives the reference's per-sample results bit for bit (tests/test_audio.py pins that against the
// reference's own code) while the sync-pulse pattern, which depends on the sample index only, is computed once
// per sample frame into a small table instead of once per channel.  No CUDA in this file.
#include <cmath>
#include <cstdint>
#include <new>
#include <vector>

#include "../../include/cvs_ntsc.h"
#include "glibc_rand.h"

namespace {

constexpr int kSampleRate = 44100;            // output_audio_rate (:212)
constexpr int kBandPasses = 6;                // audio_hilopass.init(6) (:2032)
constexpr unsigned kBuzzOversample = 16;      // (:928)

// One-pole section.  Bit-exactness fixes the three roundings of the update, y = x a + (y - y a)  (:90-94).
class Pole {
public:
    void tune(double cutoff_hz) {                      // setFilter (:78-86)
        const double dt = 1.0 / kSampleRate;
        const double tau = 1 / (cutoff_hz * 2 * M_PI);
        a_ = dt / (tau + dt);
        y_ = 0;
    }
    double low(double x) {
        const double in = x * a_;
        const double keep = y_ - (y_ * a_);
        y_ = in + keep;
        return y_;
    }
    double high(double x) { return x - low(x); }       // (:95-99)

private:
    double a_ = 0, y_ = 0;
};

double db_to_gain(double db) { return std::pow(10.0, db / 20.0); }          // dBFS (:51-58)

// A packet in flight: interleaved doubles, `ch` channels.
struct Packet {
    std::vector<double> v;
    unsigned frames = 0;
    int ch = 1;
};

// Band limiting: per channel, kBandPasses low sections then kBandPasses high sections (HiLoPass::filter, :126-130).
class BandLimiter {
public:
    void configure(int ch, double low_hz, double high_hz) {
        bank_.assign((size_t)ch * 2 * kBandPasses, Pole());
        for (int c = 0; c < ch; c++)
            for (int i = 0; i < kBandPasses; i++) {
                section(c, i).tune(low_hz);                 // the lowpass runs at output_audio_lowpass (:112-115)
                section(c, kBandPasses + i).tune(high_hz);
            }
    }
    void run(Packet &pk) {
        for (int c = 0; c < pk.ch; c++) {
            Pole *s = &section(c, 0);
            double *x = pk.v.data() + c;
            for (unsigned n = 0; n < pk.frames; n++, x += pk.ch) {
                double t = *x;
                for (int i = 0; i < kBandPasses; i++) t = s[i].low(t);
                for (int i = kBandPasses; i < 2 * kBandPasses; i++) t = s[i].high(t);
                *x = t;
            }
        }
    }

private:
    Pole &section(int c, int i) { return bank_[(size_t)c * 2 * kBandPasses + (size_t)i]; }
    std::vector<Pole> bank_;
};

// Pre- and de-emphasis.  The reference runs EVERY channel's section over each channel's sample (:919-923,
// :959-963), so the sections see the interleaved sequence: this stage sweeps the packet as one stream.
class Emphasis {
public:
    void configure(int ch, double hz, bool boost_highs) {
        on_ = true;
        boost_ = boost_highs;
        bank_.assign((size_t)ch, Pole());
        for (Pole &p : bank_) p.tune(hz);
    }
    bool enabled() const { return on_; }
    void run(Packet &pk) {
        for (double &x : pk.v)
            for (Pole &p : bank_) x = boost_ ? x + p.high(x) : p.low(x);
    }

private:
    bool on_ = false, boost_ = false;
    std::vector<Pole> bank_;
};

// Audio/video crosstalk on linear tracks (:926-945): within each sample, 16 sub-instants; each one that falls into a
// horizontal or vertical sync pulse lowers the sample by the same small step.  The pattern depends on the running
// sample index only: one table entry (pulse count 0..16) per sample frame.
class SyncBuzz {
public:
    void configure(bool ntsc, double level_db) {
        step_ = db_to_gain(level_db);
        on_ = step_ > 0.000000001;
        line_hz_ = ntsc ? 15734 : 15625;
        half_lines_ = (ntsc ? 525 : 625) / 2.0;
        vpulse_lines_ = ntsc ? 10 : 12;
        hpulse_ = ntsc ? (line_hz_ * (4.7 / 1000000)) : (line_hz_ * (4.0 / 1000000));
    }
    bool enabled() const { return on_; }
    void run(Packet &pk, unsigned long long first_frame) {
        counts_.resize(pk.frames);
        for (unsigned n = 0; n < pk.frames; n++) {
            unsigned hits = 0;
            for (unsigned k = 0; k < kBuzzOversample; k++) {
                const double t = ((((double)(first_frame + n) * kBuzzOversample) + k) * line_hz_) / kSampleRate / kBuzzOversample;
                const double hpos = std::fmod(t, 1.0);
                const int vline = (int)std::fmod(std::floor(t + 0.0001 - hpos), half_lines_);
                hits += (hpos < hpulse_ || vline < vpulse_lines_) ? 1u : 0u;
            }
            counts_[n] = (uint8_t)hits;
        }
        const double dip = step_ / kBuzzOversample / 2;
        double *x = pk.v.data();
        for (unsigned n = 0; n < pk.frames; n++)
            for (int c = 0; c < pk.ch; c++, x++)
                for (unsigned h = counts_[n]; h != 0; h--) *x -= dip;     // repeated, not multiplied: same roundings
    }

private:
    bool on_ = false;
    double step_ = 0, line_hz_ = 15734, half_lines_ = 262.5, hpulse_ = 0;
    int vpulse_lines_ = 10;
    std::vector<uint8_t> counts_;
};

// Tape hiss (:951-953): one draw of the shared rand() stream per sample and channel, in packet order.
class Hiss {
public:
    void configure(double hiss_db) { level_ = (int)(db_to_gain(hiss_db) * 5000); }   // (:1267, int = double)
    bool enabled() const { return level_ != 0; }
    void run(Packet &pk, cvs::RandCursor &rng) {
        const unsigned span = (unsigned)((level_ * 2) + 1);
        for (double &x : pk.v) x += ((double)(((int)(rng.next() % span)) - level_)) / 20000;
    }

private:
    int level_ = 0;
};

// "some VCRs boost higher frequencies when playing linear tracks" (:955-957): per channel.
class HighBoost {
public:
    void configure(double amount) {
        amount_ = amount;
        for (Pole &p : bank_) p.tune(10000);                                  // (:2039-2040)
    }
    bool enabled() const { return amount_ > 0; }
    void run(Packet &pk) {
        for (int c = 0; c < pk.ch; c++) {
            double *x = pk.v.data() + c;
            for (unsigned n = 0; n < pk.frames; n++, x += pk.ch) *x += bank_[c].high(*x) * amount_;
        }
    }

private:
    double amount_ = 0;
    Pole bank_[2];
};

}  // namespace

struct cvs_audio {
    cvs_params p;
    int channels = 2;                  // output_audio_channels after parse_argv (:1227-1262)
    bool linear_track = false;         // !output_vhs_hifi
    BandLimiter band;
    Emphasis pre, post;
    SyncBuzz buzz;
    Hiss hiss;
    HighBoost boost;
    Packet pk;
    unsigned long long frames_done = 0;   // audio_proc_count (:889)
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
    a->channels = cvs_audio_channels(p);
    a->linear_track = !p->output_vhs_hifi;
    // pass band: the end of parse_argv(), :1227-1262
    double high_hz = 20, low_hz = 20000;
    if (a->channels == 1) {
        high_hz = 100;
        low_hz = p->output_vhs_tape_speed == CVS_VHS_SP ? 10000 : (p->output_vhs_tape_speed == CVS_VHS_LP ? 7000 : 4000);
    }
    a->band.configure(a->channels, low_hz, high_hz);                          // :2031-2036
    const double emph_hz = p->output_vhs_hifi ? 16000 : 8000;                  // :2044-2066
    if (p->emulating_preemphasis) a->pre.configure(a->channels, emph_hz, true);
    if (p->emulating_deemphasis) a->post.configure(a->channels, emph_hz, false);
    a->buzz.configure(p->output_ntsc != 0, p->output_audio_linear_buzz);
    a->hiss.configure(p->output_audio_hiss_db);
    a->boost.configure(p->vhs_linear_high_boost);
    *out = a;
    return CVS_OK;
}

void cvs_audio_destroy(cvs_audio *a) { delete a; }

int cvs_audio_process(cvs_audio *a, int16_t *audio, unsigned samples, unsigned long long *rng_pos) {
    if (!a || (!audio && samples) || !rng_pos) return CVS_ERR_INVALID_ARG;
    if (!a->p.enable_audio_emulation) return CVS_OK;                       // process_audio(), :1288-1289
    if (a->cur.pos() != *rng_pos) a->cur.seek(*rng_pos);
    Packet &pk = a->pk;
    pk.ch = a->channels;
    pk.frames = samples;
    const size_t total = (size_t)samples * (size_t)a->channels;
    pk.v.resize(total);
    for (size_t i = 0; i < total; i++) pk.v[i] = (double)audio[i] / 32768;  // :913
    a->band.run(pk);
    if (a->pre.enabled()) a->pre.run(pk);
    if (a->linear_track && a->buzz.enabled()) a->buzz.run(pk, a->frames_done);
    for (double &x : pk.v) x = x > 1.0 ? 1.0 : (x < -1.0 ? -1.0 : x);       // analog limiting (:948-951)
    if (a->hiss.enabled()) a->hiss.run(pk, a->cur);
    if (a->linear_track && a->boost.enabled()) a->boost.run(pk);
    if (a->post.enabled()) a->post.run(pk);
    for (size_t i = 0; i < total; i++) {                                    // :965, clips16 (:892-899)
        const int q = (int)(pk.v[i] * 32768);
        audio[i] = (int16_t)(q < -32768 ? -32768 : (q > 32767 ? 32767 : q));
    }
    a->frames_done += samples;
    *rng_pos = a->cur.pos();
    return CVS_OK;
}

}  // extern "C"
