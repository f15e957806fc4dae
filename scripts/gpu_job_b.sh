#!/bin/bash
# A/B of build variants (variants/libcvs_*.so, see the EXTRA flags in the commit message / DESIGN.md): for each,
# the fp32 == CPU-emulation bit-exactness test and a short bench of the headline workload.
mkdir -p gpurun_out
for v in default t1 t2 t2c2 t2u2; do
  if [ $v = default ]; then unset CVS_NTSC_LIB; else export CVS_NTSC_LIB=$PWD/variants/libcvs_$v.so; fi
  python -m pytest tests/test_gpu_parity.py -q -x -k "bit_identical or 1080p_vhs_sp" 2>&1 | tail -1 | sed "s/^/$v tests: /"
  python bench.py --steps 20 --warmup 3 --cpu-fields 0 --e2e-batch 16 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python bench.py --steps 20 --warmup 3 --cpu-fields 0 --e2e-batch 16 > gpurun_out/ab2_$v.json 2> /dev/null
done 2>&1 | tee gpurun_out/ab_tests.txt
python - <<'PY'
import json
for v in ("default", "t1", "t2", "t2c2", "t2u2"):
    for pre in ("ab", "ab2"):
        try:
            d = json.load(open("gpurun_out/%s_%s.json" % (pre, v)))
            print(v, pre, "value %.0f frac %.4f kernel_ms %.3f sm %s" % (d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["clocks"]["sm_mhz"]))
        except Exception as e:
            print(v, pre, "failed", e)
PY
