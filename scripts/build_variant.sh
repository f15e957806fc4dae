#!/bin/bash
# A/B builds: recompile only the headline kernels (VHS-SP exact + fast noise, composite-only) with extra -D flags
# and link them with the default objects into variants/libcvs_<name>.so (git-ignored; select with CVS_NTSC_LIB).
#   scripts/build_variant.sh <name> "<-D flags>"
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../composite_video_simulator_b200/csrc"
make -j16 >/dev/null
mkdir -p build_var/$name ../../variants
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -Xcompiler -ffp-contract=off"
for k in kern_f_sp_tv kern_n_sp_tv kern_f_comp_tv kern_f_ep_tv; do
  $NV $flags -c $k.cu -o build_var/$name/$k.o &
done
wait
objs=""
for o in build/*.o; do b=$(basename $o); if [ -f build_var/$name/$b ]; then objs="$objs build_var/$name/$b"; else objs="$objs $o"; fi; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/libcvs_$name.so $objs
echo built variants/libcvs_$name.so
