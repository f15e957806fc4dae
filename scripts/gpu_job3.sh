set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -3 gpurun_out/bench3.err; cat gpurun_out/bench3.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_fields|k_headswitch" -c 40 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 4 --warmup 3 --batch 256 --e2e-batch 32 --cpu-fields 0 > gpurun_out/bench_ncu.json 2>&1
tail -5 gpurun_out/launches_r1.csv | cut -c1-400
