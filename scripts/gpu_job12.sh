mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err
for cfg in "ep 1920 1080" "comp 3840 2160" "comp 720 480" "sp 720 480"; do set -- $cfg
python bench.py --preset $1 --width $2 --height $3 --steps 20 --warmup 3 --cpu-fields 0 --e2e-batch 16 --batch $(( 256 * 1920 * 1080 / ($2 * $3) > 1024 ? 1024 : 256 * 1920 * 1080 / ($2 * $3) )) 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 $2x$3 value %.0f kernel_ms %.3f achieved %.1f GB/s frac %.4f'%(d['value'],d['roofline']['kernel_ms_per_launch'],d['roofline']['achieved'],d['roofline']['frac']))
"
done | tee gpurun_out/bench_r1_presets.txt
ncu --set full --clock-control none --import-source on -k regex:k_fields -s 3 -c 1 -o gpurun_out/prof_r1g_kfields python bench.py --steps 1 --warmup 3 --batch 64 --e2e-batch 16 --cpu-fields 0 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_fields|k_headswitch" -c 40 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 4 --warmup 3 --batch 256 --e2e-batch 32 --cpu-fields 0 > /dev/null 2>&1
cat gpurun_out/bench_r1_n1.json | cut -c1-400
