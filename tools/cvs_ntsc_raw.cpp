// cvs_ntsc_raw -- raw-frame host for the B200 scanline engine.
//
// The reference program (ffmpeg_ntsc.cpp main(), :1923-2331) is FFmpeg demux/decode -> swscale to BGRA
// at the output size -> [composite_layer() per output field -> line doubling] -> swscale to YUV -> H.264,
// with the audio packets of the first input filtered (process_audio(), :1284-1290) in demux order.
// FFmpeg is outside this repository's scope (and not installed here), so this host replaces both ends
// with raw BGRA8 / S16LE files or pipes and keeps the middle exactly as the reference's loop (:2140-2284):
//
//   audio packets that start before the next field: composite_audio_process()               :2157-2163
//   for every output field `current`:
//       composite_layer(ring[idx], source frame, field = (current & 1) ^ 1, fieldno = current)   :2229
//       line-double the field inside ring[idx]                                                     :2232-2257
//       emit ring[idx]; idx = (idx + 1) % delay                                                    :2276-2280
//
// Audio and video draw from ONE rand() stream (tape hiss :952, video noise :1640...), so their interleaving is
// part of the result.  A demuxer delivers an audio packet before the pictures it plays with; this host uses that
// order with fixed-size packets: before field `current` (shown at current * 1001 / 60000 s, or current / 50 s
// for PAL) it runs every audio packet that starts at or before that instant, and the rest after the last field.
// The stream position travels between the two engines (cvs_rng_tell / cvs_audio_process / cvs_rng_seek).
//
// Host side of a batch: pictures live in page-locked memory (cvs_alloc_host), two buffer sets; batch k+1 is read
// while batch k is on the GPU (cvs_composite_fields_host_async) and batch k is written while k+1 computes.
//
// Switches: every switch of the reference's parse_argv() (ffmpeg_ntsc.cpp:972-1282; -i/-o name raw
// files, "-" = stdin/stdout), plus  -height <n> (the reference derives it from -tvstd),
// -fields-per-frame <n> (how many output fields each input frame is shown for; default 2 = 29.97p
// material at the 59.94 field rate), -batch <n> (fields per GPU launch), -double (fp64 validation mode),
// -fast-noise (per-pixel noise from counter generators instead of the exact rand() replay: within +-1 LSB,
// per-line effects unchanged; cvs_set_noise_mode), -audio-in <file> -audio-out <file> (raw interleaved S16LE at
// 44.1 kHz with cvs_audio_channels() channels: 2, or 1 for linear-track VHS audio), -audio-packet <frames>.
//
// Example (ffmpeg on either side does the decode/encode the reference does in-process):
//   ffmpeg -i in.mp4 -vf scale=720:480 -pix_fmt bgra -f rawvideo - |
//   cvs_ntsc_raw -i - -o - -vhs -vhs-speed ep |
//   ffmpeg -f rawvideo -pix_fmt bgra -s 720x480 -r 60000/1001 -i - -c:v libx264 out.mp4
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/cvs_ntsc.h"

namespace {

size_t read_full(FILE *f, void *buf, size_t n) {
    size_t got = 0;
    while (got < n) {
        const size_t r = fread((uint8_t *)buf + got, 1, n - got, f);
        if (r == 0) break;
        got += r;
    }
    return got;
}

[[noreturn]] void die(const char *what, int rc) {
    fprintf(stderr, "%s: %s\n", what, cvs_strerror(rc));
    exit(1);
}

// One batch in flight: n consecutive output fields starting at `first`, their source pictures and results.
struct Batch {
    uint8_t *src = nullptr, *dst = nullptr;      // page-locked, `cap` pictures each
    int n = 0;
    unsigned long long first = 0;
};

// The audio side stream: fixed-size packets, filtered in place, copied through.
struct AudioTrack {
    FILE *in = nullptr, *out = nullptr;
    cvs_audio *fx = nullptr;
    int channels = 2, packet = 1024;
    unsigned long long frames_done = 0;
    bool eof = false;
    std::vector<int16_t> buf;

    // one packet; returns false at the end of the track
    bool step(unsigned long long *rng_pos) {
        if (!in || eof) return false;
        buf.resize((size_t)packet * (size_t)channels);
        const size_t got = read_full(in, buf.data(), buf.size() * 2) / (2 * (size_t)channels);
        if (got == 0) { eof = true; return false; }
        const int rc = cvs_audio_process(fx, buf.data(), (unsigned)got, rng_pos);
        if (rc != CVS_OK) die("cvs_audio_process", rc);
        if (fwrite(buf.data(), 2 * (size_t)channels, got, out) != got) { perror("write audio"); exit(1); }
        frames_done += got;
        if ((int)got < packet) eof = true;
        return true;
    }
};

}  // namespace

int main(int argc, char **argv) {
    std::string in_path, out_path, ain_path, aout_path;
    int height = 0, fields_per_frame = 2, batch = 32, use_double = 0, fast_noise = 0, apacket = 1024;
    std::vector<const char *> ref_argv;
    ref_argv.push_back(argv[0]);
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        const char *n = a;
        while (*n == '-') n++;
        auto need = [&](const char *what) -> const char * {
            if (i + 1 >= argc) { fprintf(stderr, "%s needs a value\n", what); exit(1); }
            return argv[++i];
        };
        if (a[0] == '-' && !strcmp(n, "i")) { in_path = need("-i"); ref_argv.push_back("-i"); ref_argv.push_back(in_path.c_str()); }
        else if (a[0] == '-' && !strcmp(n, "o")) { out_path = need("-o"); ref_argv.push_back("-o"); ref_argv.push_back(out_path.c_str()); }
        else if (a[0] == '-' && !strcmp(n, "height")) height = atoi(need("-height"));
        else if (a[0] == '-' && !strcmp(n, "fields-per-frame")) fields_per_frame = atoi(need("-fields-per-frame"));
        else if (a[0] == '-' && !strcmp(n, "batch")) batch = atoi(need("-batch"));
        else if (a[0] == '-' && !strcmp(n, "audio-in")) ain_path = need("-audio-in");
        else if (a[0] == '-' && !strcmp(n, "audio-out")) aout_path = need("-audio-out");
        else if (a[0] == '-' && !strcmp(n, "audio-packet")) apacket = atoi(need("-audio-packet"));
        else if (a[0] == '-' && !strcmp(n, "double")) use_double = 1;
        else if (a[0] == '-' && !strcmp(n, "fast-noise")) fast_noise = 1;
        else ref_argv.push_back(a);
    }
    cvs_params p;
    cvs_params_default_ntsc(&p);
    int rc = cvs_params_apply_argv(&p, (int)ref_argv.size(), ref_argv.data());
    if (rc != CVS_OK) {
        fprintf(stderr, "%s: %s\n", argv[0], cvs_strerror(rc));
        return 1;                                                         // parse_argv() failure => exit 1 (:1925)
    }
    if (in_path.empty()) { fprintf(stderr, "No input files specified\n"); return 1; }     // :1276-1279
    if (out_path.empty()) { fprintf(stderr, "No output file specified\n"); return 1; }    // :1272-1275
    if (ain_path.empty() != aout_path.empty()) { fprintf(stderr, "-audio-in and -audio-out go together\n"); return 1; }
    const int w = p.output_width, h = height > 0 ? height : p.output_height;
    const int delay = p.output_frame_delay > 0 ? p.output_frame_delay : 1;
    if (fields_per_frame < 1 || batch < 1 || w < 1 || h < 2 || apacket < 1) { fprintf(stderr, "bad geometry\n"); return 1; }

    FILE *fin = in_path == "-" ? stdin : fopen(in_path.c_str(), "rb");
    FILE *fout = out_path == "-" ? stdout : fopen(out_path.c_str(), "wb");
    if (!fin || !fout) { perror("open"); return 1; }

    cvs_ctx *ctx = nullptr;
    rc = cvs_create(&ctx, &p, 0, w, h, batch);
    if (rc != CVS_OK) die("cvs_create", rc);
    cvs_set_bob(ctx, 1);
    cvs_set_precision(ctx, use_double);
    cvs_set_noise_mode(ctx, fast_noise ? CVS_NOISE_FAST : CVS_NOISE_EXACT);

    AudioTrack audio;
    if (!ain_path.empty()) {
        audio.in = fopen(ain_path.c_str(), "rb");
        audio.out = fopen(aout_path.c_str(), "wb");
        if (!audio.in || !audio.out) { perror("open audio"); return 1; }
        audio.channels = cvs_audio_channels(&p);
        audio.packet = apacket;
        rc = cvs_audio_create(&audio.fx, &p);
        if (rc != CVS_OK) die("cvs_audio_create", rc);
    }
    // field period in units of 1 / (44100 * den) s: field `c` is shown at c * num / den seconds
    const unsigned long long f_num = p.output_ntsc ? 1001ull : 1ull, f_den = p.output_ntsc ? 60000ull : 50ull;

    const size_t pic = (size_t)w * h * 4, row = (size_t)w * 4;
    Batch sets[2];
    for (Batch &b : sets) {
        void *a = nullptr, *d = nullptr;
        if ((rc = cvs_alloc_host(ctx, &a, (size_t)batch * pic)) != CVS_OK) die("cvs_alloc_host", rc);
        if ((rc = cvs_alloc_host(ctx, &d, (size_t)batch * pic)) != CVS_OK) die("cvs_alloc_host", rc);
        b.src = (uint8_t *)a;
        b.dst = (uint8_t *)d;
    }
    // frame ring of the reference: `delay` zero-initialised pictures (:2069-2092, index wraps at delay :2277).
    // Only one row of it ever matters: composite_layer() and the line doubling rewrite every row of a picture
    // except row h-1 of field-0 pictures when h is even (:2247), which keeps what the slot held `delay` pictures
    // earlier.  So the ring is `delay` last rows.
    std::vector<std::vector<uint8_t>> ring_last((size_t)delay, std::vector<uint8_t>(row, 0));
    std::vector<uint8_t> frame(pic);
    unsigned long long current = 0;               // next field to be scheduled
    int ring_idx = 0, have_frame = 0, shown = 0;
    bool eof = false;

    // gather the source pictures of the next batch (I/O only: safe while another batch is in flight); pictures that
    // an earlier batch gathered but could not launch (an audio packet was due first) come first
    std::vector<uint8_t> carry;
    auto gather = [&](Batch &b) {
        b.n = (int)(carry.size() / pic);
        memcpy(b.src, carry.data(), carry.size());
        carry.clear();
        while (!eof && b.n < batch) {
            if (!have_frame || shown == fields_per_frame) {
                if (read_full(fin, frame.data(), pic) != pic) { eof = true; break; }
                have_frame = 1;
                shown = 0;
            }
            memcpy(b.src + (size_t)b.n * pic, frame.data(), pic);
            shown++;
            b.n++;
        }
    };
    // the stale row of every picture of the batch, the audio packets that precede its fields, then the launch
    auto launch = [&](Batch &b) {
        b.first = current;
        for (int k = 0; k < b.n; k++)
            memcpy(b.dst + (size_t)k * pic + (size_t)(h - 1) * row, ring_last[(size_t)((ring_idx + k) % delay)].data(), row);
        // The fields of a batch draw consecutively, so audio packets due inside the batch would have to split it:
        // cut the batch at the first field that an audio packet precedes (other than its first field).
        int n = b.n;
        if (audio.in && !audio.eof) {
            unsigned long long pos = 0;
            cvs_rng_tell(ctx, &pos);
            while (!audio.eof && audio.frames_done * f_den <= current * f_num * 44100ull)
                if (!audio.step(&pos)) break;
            cvs_rng_seek(ctx, pos);
            for (int k = 1; k < n; k++)
                if (!audio.eof && audio.frames_done * f_den <= (current + (unsigned long long)k) * f_num * 44100ull) { n = k; break; }
        }
        rc = cvs_composite_fields_host_async(ctx, b.dst, pic, (int)row, b.src, pic, (int)row, w, h, 0, 0, n, current);
        if (rc != CVS_OK) die("cvs_composite_fields_host_async", rc);
        const int rest = b.n - n;
        b.n = n;
        current += (unsigned long long)n;
        return rest;                               // pictures gathered but not launched (an audio packet is due first)
    };
    // after the synchronize: rows that later pictures inherit, output
    auto finish = [&](Batch &b) {
        for (int k = 0; k < b.n; k++) {
            uint8_t *pk = b.dst + (size_t)k * pic;
            // a picture later in the same batch that reuses this ring slot inherits the row the engine does not
            // write for it: the last row, when its parity is not the picture's field (:2247)
            if (k + delay < b.n) {
                const unsigned f2 = (unsigned)(((b.first + (unsigned long long)(k + delay)) & 1) ^ 1);
                if ((unsigned)((h - 1) & 1) != f2)
                    memcpy(b.dst + (size_t)(k + delay) * pic + (size_t)(h - 1) * row, pk + (size_t)(h - 1) * row, row);
            }
        }
        for (int k = (b.n > delay ? b.n - delay : 0); k < b.n; k++)
            memcpy(ring_last[(size_t)((ring_idx + k) % delay)].data(), b.dst + (size_t)k * pic + (size_t)(h - 1) * row, row);
        ring_idx = (ring_idx + b.n) % delay;
    };
    auto emit = [&](Batch &b) {
        for (int k = 0; k < b.n; k++) {
            if (fwrite(b.dst + (size_t)k * pic, 1, pic, fout) != pic) { perror("write"); exit(1); }
            fprintf(stderr, "\rOutput field %llu ", b.first + (unsigned long long)k);          // :1361
        }
    };

    // Pipeline: [gather k+1 | GPU k] -> sync k -> seed + launch k+1 -> [emit k | GPU k+1] -> ...
    int cur_set = 0;
    gather(sets[cur_set]);
    bool in_flight = false;
    Batch *prev = nullptr;
    while (sets[cur_set].n > 0) {
        Batch &b = sets[cur_set];
        if (in_flight) {                          // the previous batch must be complete before its rows seed this one
            if ((rc = cvs_synchronize(ctx)) != CVS_OK) die("cvs_synchronize", rc);
            finish(*prev);
        }
        const int total = b.n;
        const int rest = launch(b);
        if (rest > 0) carry.assign(b.src + (size_t)(total - rest) * pic, b.src + (size_t)total * pic);
        if (in_flight) emit(*prev);               // written while this batch computes
        in_flight = true;
        prev = &b;
        cur_set ^= 1;
        gather(sets[cur_set]);         // read while this batch computes
    }
    if (in_flight) {
        if ((rc = cvs_synchronize(ctx)) != CVS_OK) die("cvs_synchronize", rc);
        finish(*prev);
        emit(*prev);
    }
    if (audio.in) {                               // the rest of the audio track follows the last field
        unsigned long long pos = 0;
        cvs_rng_tell(ctx, &pos);
        while (audio.step(&pos)) {}
        cvs_rng_seek(ctx, pos);
        fclose(audio.in);
        fclose(audio.out);
        cvs_audio_destroy(audio.fx);
    }
    fprintf(stderr, "\n");
    for (Batch &b : sets) { cvs_free_host(ctx, b.src); cvs_free_host(ctx, b.dst); }
    cvs_destroy(ctx);
    if (fout != stdout) fclose(fout);
    if (fin != stdin) fclose(fin);
    return 0;
}
