#!/bin/bash
# full GPU parity suite on the current build (packed rows, XU conversions), bench, and two more build variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/gpu_tests.txt
python bench.py --steps 20 --warmup 3 --cpu-fields 0 --e2e-batch 16 > gpurun_out/c_default.json 2> gpurun_out/c_default.err; tail -3 gpurun_out/c_default.err
CVS_PACKED_ROWS=0 python bench.py --steps 20 --warmup 3 --cpu-fields 0 --e2e-batch 16 > gpurun_out/c_unpacked.json 2>/dev/null
for v in u0t1 u0t2; do
  CVS_NTSC_LIB=$PWD/variants/libcvs_$v.so python bench.py --steps 20 --warmup 3 --cpu-fields 0 --e2e-batch 16 > gpurun_out/c_$v.json 2>/dev/null
done
python - <<'PY'
import json
for v in ("default", "unpacked", "u0t1", "u0t2"):
    try:
        d = json.load(open("gpurun_out/c_%s.json" % v))
        print(v, "value %.0f batch %d frac %.4f kernel_ms %.3f sm %s" % (d["value"], d["config"]["fields_per_step_per_gpu"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["clocks"]["sm_mhz"]))
    except Exception as e:
        print(v, "failed", e)
PY
