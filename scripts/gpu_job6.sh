set -x
mkdir -p gpurun_out
nvidia-smi -L
for lib in libcvs_ntsc.so libcvs_ntsc_b3.so; do
CVS_NTSC_LIB=$PWD/composite_video_simulator_b200/$lib python bench.py --steps 20 --warmup 3 --cpu-fields 0 --e2e-batch 64 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib value %.0f e2e %.0f kernel_ms %.3f frac %.4f ms/step %.3f'%(d['value'],d['e2e']['value'],d['roofline']['kernel_ms_per_launch'],d['roofline']['frac'],d['ms_per_step']))
    else: print(l.rstrip())
"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -5 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; tail -3 gpurun_out/bench_ref_n2.err; cat gpurun_out/bench_ref_n2.json | cut -c1-300
