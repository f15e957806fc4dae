// cvs_ntsc_raw -- raw-frame host for the B200 scanline engine.
//
// The reference program (ffmpeg_ntsc.cpp main(), :1923-2331) is FFmpeg demux/decode -> swscale to BGRA
// at the output size -> [composite_layer() per output field -> line doubling] -> swscale to YUV -> H.264.
// FFmpeg is outside this repository's scope (and not installed here), so this host replaces both ends
// with raw BGRA8 files/pipes and keeps the middle exactly as the reference's field loop (:2202-2282):
//
//   for every output field `current`:
//       composite_layer(ring[idx], source frame, field = (current & 1) ^ 1, fieldno = current)   :2229
//       line-double the field inside ring[idx]                                                     :2232-2257
//       emit ring[idx]; idx = (idx + 1) % delay                                                    :2276-2280
//
// Switches: every switch of the reference's parse_argv() (ffmpeg_ntsc.cpp:972-1282; -i/-o name raw
// files, "-" = stdin/stdout), plus  -height <n> (the reference derives it from -tvstd),
// -fields-per-frame <n> (how many output fields each input frame is shown for; default 2 = 29.97p
// material at the 59.94 field rate), -batch <n> (fields per GPU launch), -double (fp64 validation mode), -fast-noise (per-pixel noise from counter
// generators instead of the exact rand() replay: within +-1 LSB, per-line effects unchanged; cvs_set_noise_mode).
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

static size_t read_full(FILE *f, uint8_t *buf, size_t n) {
    size_t got = 0;
    while (got < n) {
        size_t r = fread(buf + got, 1, n - got, f);
        if (r == 0) break;
        got += r;
    }
    return got;
}

int main(int argc, char **argv) {
    std::string in_path, out_path;
    int height = 0, fields_per_frame = 2, batch = 32, use_double = 0, fast_noise = 0;
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
    const int w = p.output_width, h = height > 0 ? height : p.output_height;
    const int delay = p.output_frame_delay > 0 ? p.output_frame_delay : 1;
    if (fields_per_frame < 1 || batch < 1 || w < 1 || h < 2) { fprintf(stderr, "bad geometry\n"); return 1; }

    FILE *fin = in_path == "-" ? stdin : fopen(in_path.c_str(), "rb");
    FILE *fout = out_path == "-" ? stdout : fopen(out_path.c_str(), "wb");
    if (!fin || !fout) { perror("open"); return 1; }

    cvs_ctx *ctx = nullptr;
    rc = cvs_create(&ctx, &p, 0, w, h, batch);
    if (rc != CVS_OK) { fprintf(stderr, "cvs_create: %s\n", cvs_strerror(rc)); return 1; }
    cvs_set_bob(ctx, 1);
    cvs_set_precision(ctx, use_double);
    cvs_set_noise_mode(ctx, fast_noise ? CVS_NOISE_FAST : CVS_NOISE_EXACT);

    const size_t pic = (size_t)w * h * 4, row = (size_t)w * 4;
    std::vector<uint8_t> frame(pic), src((size_t)batch * pic), dst((size_t)batch * pic);
    // frame ring of the reference: `delay` zero-initialised pictures (:2069-2092, index wraps at delay :2277)
    std::vector<std::vector<uint8_t>> ring((size_t)delay, std::vector<uint8_t>(pic, 0));
    unsigned long long current = 0;
    int ring_idx = 0, have_frame = 0, shown = 0;
    bool eof = false;
    while (!eof) {
        int n = 0;
        while (n < batch) {                                               // gather the next fields' source pictures
            if (!have_frame || shown == fields_per_frame) {
                if (read_full(fin, frame.data(), pic) != pic) { eof = true; break; }
                have_frame = 1;
                shown = 0;
            }
            memcpy(&src[(size_t)n * pic], frame.data(), pic);
            shown++;
            n++;
        }
        if (n == 0) break;
        // every output picture starts from what its ring slot held `delay` pictures earlier: only the row
        // that neither composite_layer() nor the line doubling rewrites matters (row h-1 of field-0
        // pictures when h is even), the rest is overwritten.  Seed the batch with the ring, row by row.
        for (int k = 0; k < n; k++) memcpy(&dst[(size_t)k * pic], ring[(size_t)((ring_idx + k) % delay)].data(), pic);
        rc = cvs_composite_fields_host(ctx, dst.data(), pic, (int)row, src.data(), pic, (int)row, w, h, 0, 0, n, current);
        if (rc != CVS_OK) { fprintf(stderr, "cvs_composite_fields_host: %s\n", cvs_strerror(rc)); return 1; }
        for (int k = 0; k < n; k++) {
            uint8_t *pk = &dst[(size_t)k * pic];
            // a picture later in the same batch that reuses this ring slot inherits the row the engine
            // does not write for it: the last row, when its parity is not the picture's field (:2247)
            if (k + delay < n) {
                const unsigned f2 = (unsigned)(((current + (unsigned long long)(k + delay)) & 1) ^ 1);
                if ((unsigned)((h - 1) & 1) != f2)
                    memcpy(&dst[(size_t)(k + delay) * pic + (size_t)(h - 1) * row], pk + (size_t)(h - 1) * row, row);
            }
            if (fwrite(pk, 1, pic, fout) != pic) { perror("write"); return 1; }
            fprintf(stderr, "\rOutput field %llu ", current + (unsigned long long)k);          // :1361
        }
        for (int k = (n > delay ? n - delay : 0); k < n; k++)
            memcpy(ring[(size_t)((ring_idx + k) % delay)].data(), &dst[(size_t)k * pic], pic);
        ring_idx = (ring_idx + n) % delay;
        current += (unsigned long long)n;
    }
    fprintf(stderr, "\n");
    cvs_destroy(ctx);
    if (fout != stdout) fclose(fout);
    if (fin != stdin) fclose(fin);
    return 0;
}
