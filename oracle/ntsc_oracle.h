/*
 * ntsc_oracle.h -- CPU restatement of the reference's composite_layer() path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker.  The product (libcvs_ntsc.so) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py compares this restatement
 * byte-for-byte with the reference's own code (oracle/_ref/libref.so, extracted at
 * build time from /root/reference/ffmpeg_ntsc.cpp) and with the known-answer hashes of
 * SURVEY.md App. D; tests/golden/ holds committed fixtures generated from that run.
 */
#ifndef NTSC_ORACLE_H
#define NTSC_ORACLE_H

#include <stdint.h>
#include "../include/cvs_ntsc.h"

#ifdef __cplusplus
extern "C" {
#endif

/* glibc rand()/random() TYPE_3 additive-feedback generator. */
typedef struct oracle_rng {
    uint32_t r[31];
    int f, b;                      /* front / back indices (glibc fptr / rptr) */
    unsigned long long pos;        /* draws consumed since seeding             */
} oracle_rng;

void     oracle_rng_seed(oracle_rng *g, unsigned seed);   /* == srand(seed); seed 1 == never seeded */
uint32_t oracle_rng_next(oracle_rng *g);                   /* == (unsigned)rand()                    */

unsigned long long oracle_fnv1a64(const void *buf, unsigned long long n);

/* Number of rand() draws one composite_layer() call makes. */
unsigned long long oracle_draws_per_field(const cvs_params *p, int w, int h, unsigned field);

/*
 * Restatement of composite_layer() (ffmpeg_ntsc.cpp:1570-1921).  Same contract as the
 * reference call: BGRA8 host buffers, writes only dst rows y == field (mod 2), alpha = 0,
 * consumes draws from *g in the reference's order.  Returns 0, or -1 when the reference
 * would return early (ffmpeg_ntsc.cpp:1578-1583).
 */
int oracle_composite_layer(const cvs_params *p, oracle_rng *g,
                           uint8_t *dst, int dst_stride,
                           const uint8_t *src, int src_stride,
                           int w, int h, int src_interlaced, int src_tff,
                           unsigned field, unsigned long long fieldno);

/* Restatement of the line doubling that follows composite_layer() (ffmpeg_ntsc.cpp:2232-2257). */
void oracle_bob(uint8_t *pic, int stride, int w, int h, unsigned field);

/* Optional stage taps for debugging: if non-NULL, receives the int planes (nl*w each, field
   rows packed) after the named stage.  stage ids: 1 = composite signal after luma noise
   (pre head-switch), 2 = after first demod+chroma noise+phase noise, 3 = after the VHS block,
   4 = after output chroma lowpass. */
typedef void (*oracle_tap_fn)(void *user, int stage, int row, int w,
                              const int *Y, const int *I, const int *Q);
void oracle_set_tap(oracle_tap_fn fn, void *user);

#ifdef __cplusplus
}
#endif
#endif
