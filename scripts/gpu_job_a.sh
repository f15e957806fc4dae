#!/bin/bash
# GPU job of this session (1 GPU): parity tests, smoke, bench of the default build and of the one-step-per-iteration
# variant (variants/libcvs_u1.so, built with EXTRA=-DCVS_FAST_UNROLL=1), launch list, one full ncu capture.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err
if [ -f variants/libcvs_u1.so ]; then
  CVS_NTSC_LIB=$PWD/variants/libcvs_u1.so python bench.py --steps 20 --warmup 3 --cpu-fields 0 --e2e-batch 16 > gpurun_out/bench_u1.json 2> gpurun_out/bench_u1.err
fi
python bench.py --steps 20 --warmup 3 --cpu-fields 0 --e2e-batch 16 > gpurun_out/bench_n1_b.json 2>/dev/null
for cfg in "ep 1920 1080 320" "comp 3840 2160 80" "comp 720 480 1024" "sp 720 480 1024"; do set -- $cfg
python bench.py --preset $1 --width $2 --height $3 --steps 20 --warmup 3 --cpu-fields 0 --e2e-batch 16 --batch $4 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 $2x$3 B=%d value %.0f kernel_ms %.3f achieved %.1f GB/s frac %.4f'%(d['config']['fields_per_step_per_gpu'],d['value'],d['roofline']['kernel_ms_per_launch'],d['roofline']['achieved'],d['roofline']['frac']))
"
done | tee gpurun_out/bench_presets.txt
python - <<'PY'
import json
for n in ("bench_n1", "bench_u1", "bench_n1_b"):
    try:
        d = json.load(open("gpurun_out/%s.json" % n))
        print(n, "value %.0f e2e %.0f frac %.4f kernel_ms %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["clocks"]))
    except Exception as e:
        print(n, "failed", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_fields|k_headswitch" -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --e2e-batch 32 --cpu-fields 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fields -s 3 -c 1 -o gpurun_out/prof_kfields python bench.py --steps 1 --warmup 3 --e2e-batch 16 --cpu-fields 0 > /dev/null 2>&1
ls -la gpurun_out | tail -12
