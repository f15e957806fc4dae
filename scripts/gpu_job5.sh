set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for ch in 16 8 32; do
CVS_HOST_CHUNK=$ch python bench.py --steps 20 --warmup 3 --cpu-fields 0 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('chunk $ch value %.0f e2e %.0f kernel_ms %.3f frac %.4f ms/step %.3f'%(d['value'],d['e2e']['value'],d['roofline']['kernel_ms_per_launch'],d['roofline']['frac'],d['ms_per_step']))
    else: print(l.rstrip())
"
done
ncu --set full --clock-control none --import-source on -k regex:k_fields -s 3 -c 1 -o gpurun_out/prof_r1d_kfields python bench.py --steps 1 --warmup 3 --batch 64 --e2e-batch 16 --cpu-fields 0 > /dev/null 2>&1
