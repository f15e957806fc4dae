// glibc_rand.cpp -- see glibc_rand.h.  Seeding follows glibc 2.39 stdlib/random_r.c
// (__srandom_r, TYPE_3: 31 words by the Lehmer step 16807*x mod 2^31-1 in Schrage form,
// front pointer 3 ahead of the rear, 310 outputs discarded); verified against this
// container's libc in tests/test_rng.py.
#include "glibc_rand.h"

#include <cstring>

namespace cvs {

RandPoly rand_poly_one() {
    RandPoly p;
    std::memset(p.c, 0, sizeof(p.c));
    p.c[0] = 1;
    return p;
}

RandPoly rand_poly_mul(const RandPoly &a, const RandPoly &b) {
    uint32_t t[2 * kRandLag - 1];
    std::memset(t, 0, sizeof(t));
    for (int i = 0; i < kRandLag; i++) {
        const uint32_t ai = a.c[i];
        if (!ai) continue;
        for (int j = 0; j < kRandLag; j++) t[i + j] += ai * b.c[j];
    }
    // reduce: x^d = x^(d-3) + x^(d-31) for d >= 31
    for (int d = 2 * kRandLag - 2; d >= kRandLag; d--) {
        const uint32_t v = t[d];
        if (!v) continue;
        t[d - kRandShortLag] += v;
        t[d - kRandLag] += v;
    }
    RandPoly r;
    std::memcpy(r.c, t, sizeof(r.c));
    return r;
}

RandPoly rand_poly_xpow(uint64_t j) {
    RandPoly result = rand_poly_one();
    RandPoly base;
    std::memset(base.c, 0, sizeof(base.c));
    base.c[1] = 1;   // x
    while (j) {
        if (j & 1) result = rand_poly_mul(result, base);
        j >>= 1;
        if (j) base = rand_poly_mul(base, base);
    }
    return result;
}

void RandCursor::seed(unsigned s) {
    if (s == 0) s = 1;
    seed_ = s;
    uint32_t r[kRandLag];
    int32_t word = (int32_t)s;
    r[0] = s;
    for (int i = 1; i < kRandLag; i++) {
        const long hi = word / 127773, lo = word % 127773;
        word = (int32_t)(16807 * lo - 2836 * hi);
        if (word < 0) word += 2147483647;
        r[i] = (uint32_t)word;
    }
    // glibc: fptr = &state[3], rptr = &state[0]; *fptr += *rptr.  In sequence form the 31
    // seed words are s[0..30] and s[n] = s[n-31] + s[n-3] from n = 34 on, with s[31..33] =
    // s[0..2] (the first three additions land on the wrapped front pointer).  Rotate so that
    // h_ holds the 31 words preceding the next produced word.
    // After k productions the ring holds words s[k+3 .. k+33]; start: s[3..33].
    for (int i = 0; i < kRandLag; i++) h_[i] = r[(i + kRandShortLag) % kRandLag];
    head_ = 0;
    pos_ = 0;
    for (int i = 0; i < 310; i++) (void)next_raw();
    pos_ = 0;
}

uint32_t RandCursor::next_raw() {
    // q[pos] = q[pos-31] + q[pos-3]
    int i3 = head_ + (kRandLag - kRandShortLag);
    if (i3 >= kRandLag) i3 -= kRandLag;
    const uint32_t v = h_[head_] + h_[i3];
    h_[head_] = v;
    if (++head_ == kRandLag) head_ = 0;
    pos_++;
    return v;
}

void RandCursor::window(uint32_t out[kRandWindow]) const {
    for (int i = 0; i < kRandLag; i++) {
        int k = head_ + i;
        if (k >= kRandLag) k -= kRandLag;
        out[i] = h_[k];
    }
    for (int i = kRandLag; i < kRandWindow; i++) out[i] = out[i - kRandLag] + out[i - kRandShortLag];
}

void rand_rebase(const uint32_t w[kRandWindow], const RandPoly &xj, uint32_t hist_out[kRandLag]) {
    for (int k = 0; k < kRandLag; k++) {
        uint32_t acc = 0;
        for (int i = 0; i < kRandLag; i++) acc += xj.c[i] * w[k + i];
        hist_out[k] = acc;
    }
}

void RandCursor::jump(const RandPoly &xj, uint64_t j) {
    uint32_t w[kRandWindow];
    window(w);
    rand_rebase(w, xj, h_);
    head_ = 0;
    pos_ += j;
}

void RandCursor::seek(uint64_t abs_pos) {
    if (abs_pos < pos_) seed(seed_);
    const uint64_t d = abs_pos - pos_;
    if (d == 0) return;
    if (d < 4096) {
        for (uint64_t i = 0; i < d; i++) (void)next_raw();
    } else {
        advance(d);
    }
}

}  // namespace cvs
