// sws_filter.cpp -- see sws_filter.h.
#include "sws_filter.h"

#include <algorithm>
#include <cstdlib>

namespace cvs {
namespace {

typedef long long i64;

struct Row {
    int first;                  // source index of w[0]
    std::vector<i64> w;         // raw weights in 2^54-ish units
};

i64 div_nearest(i64 a, i64 b) { return a >= 0 ? (a + b / 2) / b : -((-a + b / 2) / b); }

// Raw triangle weights.  The library measures positions in 1/2^17 of a source sample relative to the sample centres:
// destination sample i sits at x_i = x_0 + 2 i inc (inc = the 16.16 step), a tap at source sample s at s 2^17, and the
// weight falls linearly from 2^30 to 0 over one destination sample pitch (one source pitch when enlarging).
std::vector<Row> raw_rows(int srcn, int dstn, i64 inc, int width, i64 unit) {
    std::vector<Row> rows((size_t)dstn);
    // centred on both sides: x_0 = (128 inc >> 7) - (128 * 65536 >> 7) = inc - 65536
    i64 x = inc - 65536;
    const bool shrinking = inc > 65536;
    for (int i = 0; i < dstn; i++, x += 2 * inc) {
        Row &r = rows[(size_t)i];
        r.first = (int)((x - (i64)(width - 2) * 65536) / 131072);           // integer division rounds towards zero
        r.w.resize((size_t)width);
        for (int j = 0; j < width; j++) {
            i64 dist = std::llabs((i64)(r.first + j) * 131072 - x) << 13;
            if (shrinking) dist = dist * dstn / srcn;
            const i64 t = ((i64)1 << 30) - dist;
            r.w[(size_t)j] = t > 0 ? t * unit : 0;
        }
    }
    return rows;
}

}  // namespace

FilterBank bilinear_bank(int srcn, int dstn, int one) {
    FilterBank fb;
    fb.pos.resize((size_t)dstn);
    const i64 inc = (((i64)srcn << 16) + (dstn >> 1)) / dstn;
    if (std::llabs(inc - 65536) < 10) {                                      // same sampling grid: the unit filter
        fb.taps = 1;
        fb.coef.assign((size_t)dstn, one);
        for (int i = 0; i < dstn; i++) fb.pos[(size_t)i] = i;
        return fb;
    }
    int ratio_log2 = 0;
    for (unsigned q = (unsigned)std::max(srcn / dstn, 1); q > 1; q >>= 1) ratio_log2++;
    const i64 full = (i64)1 << (54 - std::min(ratio_log2, 8));               // the weight of a whole sample
    int width = inc <= 65536 ? 3 : 1 + (int)(((i64)2 * srcn + dstn - 1) / dstn);
    width = std::max(1, std::min(width, srcn - 2));
    std::vector<Row> rows = raw_rows(srcn, dstn, inc, width, full >> 30);

    // Taps that together weigh less than 0.2 % of a sample are dropped: leading ones by moving the row's window to the
    // right (as long as the windows stay ordered), trailing ones by narrowing the bank to the widest row's need.
    const double negligible = 0.002 * (double)full;
    int taps = 0;
    for (int i = dstn - 1; i >= 0; i--) {
        Row &r = rows[(size_t)i];
        i64 seen = 0;
        for (int moved = 0; moved < width; moved++) {
            seen += std::llabs(r.w[0]);
            if ((double)seen > negligible) break;
            if (i + 1 < dstn && r.first >= rows[(size_t)i + 1].first) break;
            std::rotate(r.w.begin(), r.w.begin() + 1, r.w.end());
            r.w.back() = 0;
            r.first++;
        }
        int need = width;
        seen = 0;
        for (int j = width - 1; j > 0; j--) {
            seen += std::llabs(r.w[(size_t)j]);
            if ((double)seen > negligible) break;
            need--;
        }
        taps = std::max(taps, need);
    }
    fb.taps = taps;
    fb.coef.assign((size_t)dstn * (size_t)taps, 0);
    for (int i = 0; i < dstn; i++) {
        Row &r = rows[(size_t)i];
        r.w.resize((size_t)taps);
        if (r.first < 0) {                                                   // weights left of the picture go to sample 0 ...
            for (int j = 1; j < taps; j++) {
                const int to = std::max(j + r.first, 0);
                r.w[(size_t)to] += r.w[(size_t)j];
                r.w[(size_t)j] = 0;
            }
            r.first = 0;
        }
        if (r.first + taps > srcn) {                                         // ... and those right of it to the last sample
            const int shift = r.first + std::min(taps - srcn, 0);
            i64 outside = 0;
            for (int j = taps - 1; j >= 0; j--)
                if (r.first + j >= srcn) { outside += r.w[(size_t)j]; r.w[(size_t)j] = 0; }
            for (int j = taps - 1; j >= 0; j--) r.w[(size_t)j] = j < shift ? 0 : r.w[(size_t)(j - shift)];
            r.first -= shift;
            r.w[(size_t)(srcn - 1 - r.first)] += outside;
        }
        // scale to integers that sum to `one`: each coefficient is rounded and the remainder travels to the next tap
        i64 total = 0;
        for (i64 v : r.w) total += v;
        i64 per_one = (total + one / 2) / one;
        if (per_one == 0) per_one = 1;
        i64 carry = 0;
        for (int j = 0; j < taps; j++) {
            const i64 v = r.w[(size_t)j] + carry;
            const i64 q = div_nearest(v, per_one);
            fb.coef[(size_t)i * (size_t)taps + (size_t)j] = (int32_t)q;
            carry = v - q * per_one;
        }
        fb.pos[(size_t)i] = r.first;
    }
    return fb;
}

}  // namespace cvs
