mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_fields -s 3 -c 1 -o gpurun_out/prof_r1e_kfields python bench.py --steps 1 --warmup 3 --batch 64 --e2e-batch 16 --cpu-fields 0 > /dev/null 2>&1
ls -la gpurun_out/prof_r1e_kfields.ncu-rep
