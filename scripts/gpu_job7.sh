set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --cpu-fields 0 --e2e-batch 64 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f e2e %.0f kernel_ms %.3f frac %.4f ms/step %.3f'%(d['value'],d['e2e']['value'],d['roofline']['kernel_ms_per_launch'],d['roofline']['frac'],d['ms_per_step']))
    else: print(l.rstrip())
"
