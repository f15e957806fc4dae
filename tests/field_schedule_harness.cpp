// CPU check of tools/field_schedule.h (TEST INFRASTRUCTURE): the schedule driven like tools/cvs_ffmpeg_ntsc.cpp drives it,
// with picture NUMBERS in place of pictures, against a direct restatement of the reference's rule (ffmpeg_ntsc.cpp:2146-2283):
// field f shows the last picture whose start field is <= f, fields before the first picture show the zeroed frame (-1), the
// last picture lasts two fields.  Prints nothing and returns 0 when every scenario agrees.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../tools/field_schedule.h"

static int run(const std::vector<long long> &starts, int batch, bool verbose) {
    FieldSchedule s(batch);
    std::vector<int> store((size_t)batch + 1, -99), shown;      // slot -> picture number; the emitted sequence
    long long first_field_of_next_flush = 0;
    int bad = 0, flushes = 0;
    auto flush = [&]() {
        if (s.current != first_field_of_next_flush) bad++;
        if ((int)s.src_of_field.size() > batch) bad++;
        for (int slot : s.src_of_field) {
            if (slot < 0 || slot >= s.stored) { bad++; continue; }
            shown.push_back(store[(size_t)slot]);
        }
        first_field_of_next_flush += (long long)s.src_of_field.size();
        if (s.stored > 1) store[0] = store[(size_t)s.stored - 1];
        s.flushed();
        flushes++;
    };
    auto black = [&](int slot) { store[(size_t)slot] = -1; };
    for (size_t k = 0; k < starts.size(); k++) {
        const int slot = s.picture(starts[k], flush, black);
        if (slot < 0 || slot > batch) { bad++; break; }
        store[(size_t)slot] = (int)k;
    }
    s.finish(flush);
    // the rule, directly
    std::vector<int> want;
    std::vector<long long> at(starts.size());
    long long cur = 0;
    for (size_t k = 0; k < starts.size(); k++) {               // a picture without pts starts where the previous one was cut
        // (cur = number of fields decided so far when picture k arrives)
        at[k] = starts[k] < 0 ? cur : starts[k];
        if (at[k] > cur) {
            const int prev = k == 0 ? -1 : (int)k - 1;
            for (; cur < at[k]; cur++) want.push_back(prev);
        }
    }
    if (!starts.empty()) { want.push_back((int)starts.size() - 1); want.push_back((int)starts.size() - 1); }
    if (want != shown) bad++;
    if (verbose || bad) {
        printf("batch %d: %zu pictures, %zu fields, %d flushes, %s\n", batch, starts.size(), shown.size(), flushes, bad ? "MISMATCH" : "ok");
        if (bad) {
            for (size_t i = 0; i < want.size() || i < shown.size(); i++)
                if (i >= want.size() || i >= shown.size() || want[i] != shown[i]) { printf("  first difference at field %zu\n", i); break; }
        }
    }
    return bad;
}

int main(int argc, char **) {
    const bool verbose = argc > 1;
    int bad = 0;
    std::vector<std::vector<long long>> scenarios;
    std::vector<long long> v;
    for (int k = 0; k < 100; k++) v.push_back(2 * k);                      // 29.97p material: two fields per picture
    scenarios.push_back(v);
    v.clear(); for (int k = 0; k < 100; k++) v.push_back(k);              // 59.94p: one field per picture
    scenarios.push_back(v);
    v.clear(); for (int k = 0; k < 60; k++) v.push_back((k * 5) / 2);     // 23.976p: 3:2 pulldown (2, 3, 2, 3 ... fields)
    scenarios.push_back(v);
    v.clear(); for (int k = 0; k < 40; k++) v.push_back(7 + 2 * k);       // the stream starts late: black first
    scenarios.push_back(v);
    v.clear(); for (int k = 0; k < 40; k++) v.push_back(-1);              // no pts at all: every picture replaces the last at once
    scenarios.push_back(v);
    v.clear(); for (int k = 0; k < 30; k++) v.push_back(k < 10 ? 2 * k : (k < 12 ? 18 : 2 * k + 50));   // repeats and a long gap
    scenarios.push_back(v);
    v.clear(); v.push_back(0); scenarios.push_back(v);                    // one picture
    v.clear(); v.push_back(400); scenarios.push_back(v);                  // one late picture: 400 black fields
    scenarios.push_back(std::vector<long long>());                        // nothing
    unsigned seed = 12345;
    for (int t = 0; t < 200; t++) {                                        // random monotone schedules with missing pts
        v.clear();
        long long a = (seed = seed * 1664525u + 1013904223u) % 5;
        const int n = 1 + (int)((seed = seed * 1664525u + 1013904223u) % 70);
        for (int k = 0; k < n; k++) {
            seed = seed * 1664525u + 1013904223u;
            v.push_back((seed >> 28) == 0 ? -1 : a);
            a += (seed >> 8) % 7;
        }
        scenarios.push_back(v);
    }
    for (const auto &sc : scenarios)
        for (int batch : {1, 2, 3, 7, 32, 64}) bad += run(sc, batch, verbose);
    return bad ? 1 : 0;
}
