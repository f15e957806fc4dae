// refaudio_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin C wrapper around the reference's own composite_audio_process() (ffmpeg_ntsc.cpp:901-970), whose source is
// NOT in this repository: oracle/Makefile extracts it at build time by line range (dBFS :49-58, the filter classes
// :72-204, the globals :205-214 and :750-809, the function and its statics :889-970) into the git-ignored
// oracle/_ref/audio_ref.inc.  Written here: the AVRational shim the globals need, a C entry point that copies a
// cvs_params block onto the reference's globals and then performs the end of parse_argv() (:1227-1267) and main()'s
// "prepare audio filtering" block (:2031-2066) -- both restated, since they sit inside functions that need FFmpeg.
#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
using namespace std;

struct AVRational { int num, den; };

#include "_ref/audio_ref.inc"

#include "../include/cvs_ntsc.h"

extern "C" {

void refaudio_setup(const cvs_params *p) {
    output_ntsc = p->output_ntsc != 0;
    output_pal = !output_ntsc;
    emulating_vhs = p->emulating_vhs != 0;
    output_vhs_tape_speed = p->output_vhs_tape_speed;
    output_vhs_hifi = p->output_vhs_hifi != 0;
    output_vhs_linear_audio = p->output_vhs_linear_audio != 0;
    output_vhs_linear_stereo = false;
    emulating_preemphasis = p->emulating_preemphasis != 0;
    emulating_deemphasis = p->emulating_deemphasis != 0;
    enable_audio_emulation = p->enable_audio_emulation != 0;
    output_audio_hiss_db = p->output_audio_hiss_db;
    output_audio_linear_buzz = p->output_audio_linear_buzz;
    vhs_linear_high_boost = p->vhs_linear_high_boost;
    output_audio_rate = 44100;
    // end of parse_argv(), :1227-1267
    output_audio_highpass = 20; output_audio_lowpass = 20000; output_audio_channels = 2;
    if (emulating_vhs) {
        if (output_vhs_hifi) {
            output_audio_highpass = 20; output_audio_lowpass = 20000; output_audio_channels = 2;
        } else if (output_vhs_linear_audio) {
            switch (output_vhs_tape_speed) {
                case VHS_SP: output_audio_highpass = 100; output_audio_lowpass = 10000; break;
                case VHS_LP: output_audio_highpass = 100; output_audio_lowpass = 7000; break;
                case VHS_EP: output_audio_highpass = 100; output_audio_lowpass = 4000; break;
            }
            output_audio_channels = output_vhs_linear_stereo ? 2 : 1;
        }
    }
    output_audio_hiss_level = dBFS(output_audio_hiss_db) * 5000;
    // main(), :2031-2066 (fresh filter objects: the reference runs this once per process)
    audio_proc_count = 0;
    for (int i = 0; i < 2; i++) {
        audio_linear_preemphasis_pre[i] = LowpassFilter();
        audio_linear_preemphasis_post[i] = LowpassFilter();
        audio_post_vhs_boost[i] = LowpassFilter();
    }
    audio_hilopass = HiLoComboPass();
    audio_hilopass.setChannels(output_audio_channels);
    audio_hilopass.setRate(output_audio_rate);
    audio_hilopass.setCutoff(output_audio_lowpass, output_audio_highpass);
    audio_hilopass.setPasses(6);
    audio_hilopass.init();
    for (unsigned int i = 0; i < 2; i++) audio_post_vhs_boost[i].setFilter(output_audio_rate, 10000);
    if (emulating_preemphasis)
        for (int i = 0; i < output_audio_channels; i++) audio_linear_preemphasis_pre[i].setFilter(output_audio_rate, output_vhs_hifi ? 16000 : 8000);
    if (emulating_deemphasis)
        for (int i = 0; i < output_audio_channels; i++) audio_linear_preemphasis_post[i].setFilter(output_audio_rate, output_vhs_hifi ? 16000 : 8000);
}

int refaudio_channels(void) { return output_audio_channels; }
void refaudio_srand(unsigned seed) { srand(seed); }
int refaudio_rand(void) { return rand(); }

// process_audio(), :1284-1290
void refaudio_process(int16_t *audio, unsigned samples) {
    if (enable_audio_emulation) composite_audio_process(audio, samples);
}

}  // extern "C"
