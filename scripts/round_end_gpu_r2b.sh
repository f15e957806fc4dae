#!/bin/bash
# Final GPU job of round 2 (one gpurun call, 1 GPU; the budget left allowed ~7 minutes of box time):
#   /usr/local/graft/bin/gpurun --timeout 430 -- 'bash scripts/round_end_gpu_r2b.sh'
# Order = importance; every step has its own time limit so that a slow one cannot take the later ones with it.
mkdir -p gpurun_out
O=gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee $O/r2f_smoke.txt
# A. what changed since the last full GPU run: the two picture conversions (libswscale pin), the field loop / CLI around them
timeout 120 python -m pytest tests/test_gpu_scale.py tests/test_gpu_yuv_convert.py tests/test_gpu_cli.py -q -n 4 2>&1 | tail -6 | tee $O/r2f_tests_conversions.txt
# B. the bench line
timeout 240 python bench.py --steps 20 --warmup 3 > $O/bench_r2f_n1.json 2> $O/bench_r2f_n1.err; tail -2 $O/bench_r2f_n1.err
# C. the conversions on their own, and the scaler with the one-row kernel for comparison
timeout 60 python scripts/bench_yuv_convert.py > $O/bench_r2f_convert.json 2> $O/bench_r2f_convert.err
CVS_SWS_ROWS=0 timeout 60 python scripts/bench_yuv_convert.py > $O/bench_r2f_convert_onerow.json 2>> $O/bench_r2f_convert.err
cat $O/bench_r2f_convert.json; grep sws $O/bench_r2f_convert_onerow.json
# D. the rest of the GPU suite
timeout 170 python -m pytest tests/test_gpu_yuv422.py tests/test_gpu_parity.py tests/test_gpu_multi.py -q -n 8 2>&1 | tail -6 | tee $O/r2f_tests_rest.txt
# E. ncu of the two new streaming conversion kernels (cold-cache, serialised: shares and pipe figures, not absolutes)
timeout 90 ncu --set full --clock-control none --import-source on -k "regex:k_bgra_to_yuv420_tiled|k_sws_yuv_to_bgra_rows" -s 2 -c 2 -f -o $O/prof_r2f_convert \
    python scripts/bench_yuv_convert.py > /dev/null 2> $O/prof_r2f_convert.err
timeout 30 ncu -i $O/prof_r2f_convert.ncu-rep --page raw --csv > $O/prof_r2f_convert.raw.csv 2>/dev/null
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_r2f_n1.json'))
    print('N1 value %.0f e2e %.0f frac %.4f parity %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity']))
    o = d['config']['other_workloads']
    print({k: (v.get('value'), v.get('error')) if 'value' in v or 'error' in v else {kk: vv.get('roofline', {}).get('frac') for kk, vv in v.items()} for k, v in o.items()})
except Exception as e:
    print('bench line unreadable:', e)
PY
