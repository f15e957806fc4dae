// field_schedule.h -- which decoded picture every output field shows: the bookkeeping of the reference's main loop
// (ffmpeg_ntsc.cpp:2146-2283) for hosts that hand fields to the engine in batches.
//
// The reference advances `current` (the output field counter, two per frame at output_field_rate) while it is smaller
// than the pts of the next decoded picture, compositing the picture that is on show (:2203-2230); a picture therefore
// lasts from the field its pts names until the next picture's pts, and fields before the first picture show the zeroed
// frame the program starts with (:556-561).  This class keeps exactly that state for a host that gathers up to `batch`
// fields, sends them through cvs_field_loop_host() in one call and keeps the picture on show for the next batch.
// It owns no pictures: the host stores them (slot numbers are handed out here) and does the GPU call in `flush`.
// tests/test_field_schedule.py drives it on the CPU against a direct restatement of the reference's rule.
#ifndef CVS_FIELD_SCHEDULE_H
#define CVS_FIELD_SCHEDULE_H
#include <cstdint>
#include <vector>

struct FieldSchedule {
    int batch;                          // fields per GPU call at most; the host's store holds batch + 1 pictures
    long long current = 0;              // first field of the batch being gathered (the reference's `current` at its start)
    int stored = 0;                     // pictures in the host's store; the newest (stored - 1) is the one on show
    std::vector<int32_t> src_of_field;  // per gathered field: the slot of the picture it shows

    explicit FieldSchedule(int batch_) : batch(batch_) {}
    long long next_field() const { return current + (long long)src_of_field.size(); }

    // The host calls this from its flush: n fields starting at `current` were sent (slots as in src_of_field) and the
    // picture on show was moved to slot 0.
    void flushed() {
        current += (long long)src_of_field.size();
        src_of_field.clear();
        if (stored > 0) stored = 1;
    }
    // Fields [next_field(), upto) show the newest picture.  `flush()` must send what is gathered and call flushed().
    template <class Flush>
    void show_until(long long upto, Flush &&flush) {
        while (stored > 0 && next_field() < upto) {
            src_of_field.push_back(stored - 1);
            if ((int)src_of_field.size() == batch) flush();
        }
    }
    // A decoded picture that starts at field `at` (a negative value: no pts, it starts now).  Returns the slot the host
    // must store it in.  `black()` is asked for once, when the stream's first picture starts after field 0: the host
    // stores a black picture in the slot it is given, which the fields before `at` then show (:556-561).
    template <class Flush, class Black>
    int picture(long long at, Flush &&flush, Black &&black) {
        if (at < 0) at = next_field();
        if (stored == 0 && at > next_field()) {
            black(0);
            stored = 1;
        }
        show_until(at, flush);              // the previous picture lasts until this one starts
        if (stored == batch + 1) flush();   // the store is full of pictures that the gathered fields still need
        return stored++;
    }
    // End of the stream: the last picture lasts one frame (two fields), then everything goes out.
    template <class Flush>
    void finish(Flush &&flush) {
        show_until(next_field() + 2, flush);
        if (!src_of_field.empty()) flush();
    }
};
#endif
