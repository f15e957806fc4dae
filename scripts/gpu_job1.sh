set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; tail -3 gpurun_out/bench1.err; cat gpurun_out/bench1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -3 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
nproc; lscpu | grep "Model name"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --batch 64 --e2e-batch 8 --cpu-fields 0 > gpurun_out/bench_ncu.json 2>&1
tail -12 gpurun_out/launches_r1.csv
ncu --set full --clock-control none --import-source on -k regex:k_fields -s 3 -c 1 -o gpurun_out/prof_r1_kfields python bench.py --steps 1 --warmup 3 --batch 64 --e2e-batch 8 --cpu-fields 0 > gpurun_out/bench_ncu2.json 2>&1
ls -la gpurun_out
