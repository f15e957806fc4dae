// sws_filter.h -- the bilinear filter banks of libswscale, rebuilt on the host for the device conversions.
//
// The encoder-side conversion of ffmpeg_ntsc (ffmpeg_ntsc.cpp:2118-2131, 2266-2274) is sws_scale() with SWS_BILINEAR.
// Where that call resamples an axis (chroma rows of 4:2:0, chroma columns of odd-width pictures) libswscale applies a
// table of integer weights it builds once per context.  This builder produces the same table (same taps, same
// coefficients: tests/test_abi.py compares it with the oracle's restatement over a sweep of sizes, and the oracle is
// pinned against the library, tests/test_swscale_pin.py).
#ifndef CVS_SWS_FILTER_H
#define CVS_SWS_FILTER_H
#include <cstdint>
#include <vector>

namespace cvs {

struct FilterBank {
    int taps = 0;                       // coefficients per destination sample
    std::vector<int32_t> pos;           // first source sample of destination sample i
    std::vector<int32_t> coef;          // coef[i * taps + j] weighs source sample pos[i] + j; every row sums to `one`
};

// Destination sample i of dstn, centred sampling (libswscale's default chroma position on both sides), triangle
// kernel widened by the reduction ratio; `one` = 1 << 12 for a vertical bank, 1 << 14 for a horizontal one.
FilterBank bilinear_bank(int srcn, int dstn, int one);

}  // namespace cvs
#endif
