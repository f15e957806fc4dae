#!/bin/bash
# What the round-end evidence in profiles/ was produced with (run under gpurun, 1 GPU):
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash scripts/round_end_gpu.sh'
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err
# every launch of the step with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_fields|k_headswitch" -c 40 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --e2e-batch 32 --cpu-fields 0 > /dev/null 2>&1
# the dominant kernel, once
ncu --set full --clock-control none --import-source on -k regex:k_fields -s 3 -c 1 -o gpurun_out/prof_kfields \
    python bench.py --steps 1 --warmup 3 --e2e-batch 16 --cpu-fields 0 > /dev/null 2>&1
