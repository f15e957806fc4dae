#!/bin/bash
# A/B builds of the 4:2:2 kernel: recompile yuv422_kernels.cu with extra -D flags and link it with the default
# objects into variants/libcvs_<name>.so (git-ignored; select with CVS_NTSC_LIB).
#   scripts/build_variant422.sh <name> "<-D flags>"
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../composite_video_simulator_b200/csrc"
mkdir -p build_var/$name ../../variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -Xcompiler -ffp-contract=off \
  $flags -Xptxas -v -c yuv422_kernels.cu -o build_var/$name/yuv422_kernels.o 2>&1 | grep -A2 "k_yuv422E" | grep -E "spill|Used" | tr '\n' ' '
echo
objs=""
for o in build/*.o; do b=$(basename $o); if [ -f build_var/$name/$b ]; then objs="$objs build_var/$name/$b"; else objs="$objs $o"; fi; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/libcvs_$name.so $objs
echo built variants/libcvs_$name.so
