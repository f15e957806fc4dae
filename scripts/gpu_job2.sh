set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for lib in libcvs_ntsc.so libcvs_ntsc_b3.so libcvs_ntsc_b4.so; do
  CVS_NTSC_LIB=$PWD/composite_video_simulator_b200/$lib python bench.py --steps 10 --warmup 3 --cpu-fields 0 --e2e-batch 16 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib', 'value %.0f'%d['value'], 'kernel_ms %.3f'%d['roofline']['kernel_ms_per_launch'], 'frac %.4f'%d['roofline']['frac'], 'ms/step %.3f'%d['ms_per_step'], 'e2e %.0f'%d['e2e']['value'])
    else: print(l.rstrip())
"
done
ncu --set full --clock-control none --import-source on -k regex:k_fields -s 3 -c 1 -o gpurun_out/prof_r1b_kfields python bench.py --steps 1 --warmup 3 --batch 64 --e2e-batch 8 --cpu-fields 0 > gpurun_out/bench_ncu3.json 2>&1
