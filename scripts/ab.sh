#!/bin/bash
# A/B of build variants on the GPU box: scripts/ab.sh <preset> <w> <h> <batch> name1 name2 ...   ("default" = the in-tree library)
preset=$1; w=$2; h=$3; batch=$4; shift 4
for v in "$@"; do
  if [ "$v" = default ]; then unset CVS_NTSC_LIB; else export CVS_NTSC_LIB=$PWD/variants/libcvs_$v.so; fi
  for i in 1 2; do
    python bench.py --quick --preset $preset --width $w --height $h --batch $batch --steps 20 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); o=d['config']['other_noise_mode']
print('%-10s %s %dx%d value %.0f kernel_ms %.4f frac %.4f | fast-noise %.0f frac %.4f'%('$v','$preset',$w,$h,d['value'],d['roofline']['kernel_ms_per_launch'],d['roofline']['frac'],o['value'],o['roofline_frac']))"
  done
done
unset CVS_NTSC_LIB
