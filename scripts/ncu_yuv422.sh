#!/bin/bash
# ncu --set full capture of one k_yuv422 launch (1080p -vhs -vhs-speed sp, 339 fields): scripts/ncu_yuv422.sh <tag> [lib]
tag=$1; lib=$2
[ -n "$lib" ] && export CVS_NTSC_LIB=$PWD/variants/libcvs_$lib.so
ncu --set full --clock-control none --import-source on -k regex:k_yuv422 -s 3 -c 1 -f -o gpurun_out/prof_${tag}_yuv422 \
    python -c "
import sys; sys.argv=['x','--steps','1','--warmup','3','--cpu-fields','0']
sys.path.insert(0,'scripts'); import runpy; runpy.run_path('scripts/bench_yuv422.py', run_name='__main__')" > /dev/null 2> gpurun_out/prof_${tag}_yuv422.err
