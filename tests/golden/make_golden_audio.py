"""Generate tests/golden/audio_ntsc.npz from the REFERENCE's own composite_audio_process()
(oracle/_ref/librefaudio.so, extracted from /root/reference/ffmpeg_ntsc.cpp at build time):

    python tests/golden/make_golden_audio.py

For every case of tests/test_audio.py: the processed PCM of the test signal and the number of rand() draws."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402
import test_audio  # noqa: E402


def main():
    ref = helpers.load_refaudio()
    assert ref is not None, "needs /root/reference"
    out = {}
    for name, argv in test_audio.CASES.items():
        if not test_audio.known_switch(argv):
            continue
        p = helpers.params(*argv)
        ref.refaudio_setup(helpers.C.byref(p))
        ch = ref.refaudio_channels()
        want, _ = helpers.run_refaudio(ref, p, test_audio.packets(ch))
        out[name] = np.concatenate(want)
        # draws = samples x channels when the hiss level is non-zero (:952)
        lvl = int(10.0 ** (p.output_audio_hiss_db / 20.0) * 5000)
        out[name + "_draws"] = np.int64(out[name].size if (lvl != 0 and p.enable_audio_emulation) else 0)
    np.savez_compressed(os.path.join(HERE, "audio_ntsc.npz"), **out)
    print("wrote", sorted(out))


if __name__ == "__main__":
    main()
