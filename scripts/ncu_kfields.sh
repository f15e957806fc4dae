#!/bin/bash
# One `ncu --set full` capture of the headline kernel of a build: scripts/ncu_kfields.sh <tag> [variant] [preset w h batch]
# -> gpurun_out/prof_<tag>.ncu-rep (+ raw csv).  Numbers under the profiler are cold-cache and serialised.
tag=$1; v=${2:-default}; preset=${3:-sp}; w=${4:-1920}; h=${5:-1080}; batch=${6:-320}
if [ "$v" != default ]; then export CVS_NTSC_LIB=$PWD/variants/libcvs_$v.so; fi
ncu --set full --clock-control none --import-source on -k regex:k_fields -s 3 -c 1 -f -o gpurun_out/prof_$tag \
    python bench.py --quick --no-noise-side --preset $preset --width $w --height $h --batch $batch --steps 1 --warmup 3 > gpurun_out/prof_$tag.bench.json 2> gpurun_out/prof_$tag.err
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_$tag.raw.csv 2>/dev/null
unset CVS_NTSC_LIB
