#!/bin/bash
# What the round-end evidence in profiles/ was produced with (run under gpurun, 1 GPU):
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash scripts/round_end_gpu.sh'
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; tail -2 gpurun_out/bench_r1_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err
for cfg in "ep 1920 1080 320" "comp 3840 2160 80" "comp 720 480 1024" "sp 720 480 1024"; do set -- $cfg
python bench.py --preset $1 --width $2 --height $3 --steps 20 --warmup 3 --cpu-fields 0 --e2e-batch 16 --batch $4 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 $2x$3 B=%d value %.0f kernel_ms %.3f achieved %.1f GB/s frac %.4f'%(d['config']['fields_per_step_per_gpu'],d['value'],d['roofline']['kernel_ms_per_launch'],d['roofline']['achieved'],d['roofline']['frac']))
"
done | tee gpurun_out/bench_r1_presets.txt
python scripts/bench_yuv422.py --steps 5 --warmup 3 > gpurun_out/bench_r1_yuv422.json 2> gpurun_out/bench_r1_yuv422.err
# the dominant kernel, once (cold-cache, serialised: compare shares, not absolute times)
ncu --set full --clock-control none --import-source on -k regex:k_fields -s 3 -c 1 -o gpurun_out/prof_r1i_kfields python bench.py --steps 1 --warmup 3 --e2e-batch 16 --cpu-fields 0 > /dev/null 2>&1
# every launch of the step with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_fields|k_headswitch" -c 40 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 4 --warmup 3 --e2e-batch 32 --cpu-fields 0 > /dev/null 2>&1
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_n1.json')); print('N1 value %.0f e2e %.0f frac %.4f cpu %.2f clocks %s'%(d['value'],d['e2e']['value'],d['roofline']['frac'],d['cpu_baseline']['value'],d['clocks']))
d=json.load(open('gpurun_out/bench_r1_ref.json')); print('ref arm %.1f fields/s cores %d'%(d['value'],d['cpu_baseline']['cores']))"
