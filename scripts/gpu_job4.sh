set -x
mkdir -p gpurun_out
python scripts/pcie_probe.py 2>&1 | tail -8
python bench.py --steps 20 --warmup 3 --cpu-fields 0 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -3 gpurun_out/bench4.err; python -c "
import json; d=json.load(open('gpurun_out/bench4.json')); print('value %.0f e2e %.0f kernel_ms %.3f frac %.4f ms/step %.3f'%(d['value'],d['e2e']['value'],d['roofline']['kernel_ms_per_launch'],d['roofline']['frac'],d['ms_per_step']))"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_fields -s 3 -c 2 --csv --log-file gpurun_out/traffic_r1c.csv python bench.py --steps 2 --warmup 3 --batch 64 --e2e-batch 16 --cpu-fields 0 > /dev/null 2>&1
tail -6 gpurun_out/traffic_r1c.csv | cut -d, -f5,13-
