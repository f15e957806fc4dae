// glibc_rand.h -- exact model of glibc rand() (random_r TYPE_3) with O(log n) jump-ahead.
//
// The reference's only hidden state across composite_layer() calls is the libc rand()
// stream position (the generator is never seeded: SURVEY.md App. C; draws at
// ffmpeg_ntsc.cpp:1640,1655,1729,1731,1744,1896).  To shard fields across lanes, CTAs and
// GPUs and still reproduce the single-threaded reference, every consumer must be able to
// start anywhere in that stream.  The raw sequence obeys
//        q[n] = q[n-31] + q[n-3]   (mod 2^32),      rand() = q[n] >> 1,
// i.e. multiplication by x in Z/2^32[x] / (x^31 - x^28 - 1).  A jump by J is the
// polynomial x^J mod (x^31 - x^28 - 1): q[n+J] = sum_i c_i q[n+i].
#ifndef CVS_GLIBC_RAND_H
#define CVS_GLIBC_RAND_H

#include <cstdint>

namespace cvs {

constexpr int kRandLag = 31;      // long lag (degree of the recurrence)
constexpr int kRandShortLag = 3;  // short lag
constexpr int kRandWindow = 2 * kRandLag - 1;   // 61 words: history + 30 ahead, enough to re-base

struct RandPoly {
    uint32_t c[kRandLag];
};

RandPoly rand_poly_one();                                   // x^0
RandPoly rand_poly_mul(const RandPoly &a, const RandPoly &b);
RandPoly rand_poly_xpow(uint64_t j);                        // x^j mod (x^31 - x^28 - 1)

// A position in the stream: the 31 raw words preceding draw `pos`.
class RandCursor {
public:
    RandCursor() { seed(1); }
    void seed(unsigned s);                 // == srand(s); the reference runs with the default seed 1
    uint32_t next_raw();                   // q[pos++]
    uint32_t next() { return next_raw() >> 1; }   // == (unsigned)rand()
    uint64_t pos() const { return pos_; }
    // out[j] = q[pos - 31 + j], j = 0..60 (does not advance)
    void window(uint32_t out[kRandWindow]) const;
    // advance by j draws using a precomputed polynomial for x^j
    void jump(const RandPoly &xj, uint64_t j);
    void advance(uint64_t j) { jump(rand_poly_xpow(j), j); }
    // absolute seek (draws since seeding); reseeds when going backwards
    void seek(uint64_t abs_pos);

private:
    uint32_t h_[kRandLag];   // h_[(head_ + i) % 31] = q[pos - 31 + i]
    int head_;
    uint64_t pos_;
    unsigned seed_;
};

// hist_out[k] = q[base + j - 31 + k] given window w[i] = q[base - 31 + i] and xj = x^j.
void rand_rebase(const uint32_t w[kRandWindow], const RandPoly &xj, uint32_t hist_out[kRandLag]);

}  // namespace cvs
#endif
