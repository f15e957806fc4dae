#!/bin/bash
# What the round-2 evidence in profiles/ was produced with (run under gpurun, 1 GPU):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/round_end_gpu.sh'
# Everything lands in gpurun_out/ (scratch); the summaries that are judged are copied to profiles/ by hand.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -2 gpurun_out/bench_r2_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err
python scripts/bench_yuv_convert.py > gpurun_out/bench_r2_convert.json 2> gpurun_out/bench_r2_convert.err
# the dominant kernel of the three BASELINE presets, once each (cold-cache, serialised: compare shares, not absolutes)
bash scripts/ncu_kfields.sh r2_sp default sp 1920 1080 320
bash scripts/ncu_kfields.sh r2_ep default ep 1920 1080 320
bash scripts/ncu_kfields.sh r2_comp default comp 3840 2160 80
# the 4:2:2 kernel
ncu --set full --clock-control none --import-source on -k regex:k_yuv422 -s 3 -c 1 -f -o gpurun_out/prof_r2_yuv422 \
    python -c "
import sys; sys.argv=['x','--steps','1','--warmup','3','--cpu-fields','0']
sys.path.insert(0,'scripts'); import runpy; runpy.run_path('scripts/bench_yuv422.py', run_name='__main__')" > /dev/null 2> gpurun_out/prof_r2_yuv422.err
ncu -i gpurun_out/prof_r2_yuv422.ncu-rep --page raw --csv > gpurun_out/prof_r2_yuv422.raw.csv 2>/dev/null
# every launch of a bench step with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_fields|k_headswitch" -c 40 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --quick --steps 4 --warmup 3 > /dev/null 2>&1
python - <<'PY'
import csv, json, hashlib, os
def load(fn):
    rows = list(csv.reader(open(fn))); return {h: v for h, v in zip(rows[0], rows[2])}
def f(x): return float(x.replace(',', ''))
caps = {}
for tag, key, fields_key in (("r2_sp", "sp_1920x1080", None), ("r2_ep", "ep_1920x1080", None), ("r2_comp", "comp_3840x2160", None)):
    try:
        m = load("gpurun_out/prof_%s.raw.csv" % tag)
        rd, wr = f(m["dram__bytes_read.sum"]), f(m["dram__bytes_write.sum"])
        # ncu prints Gbyte / Mbyte units in row 2: read the unit row
        rows = list(csv.reader(open("gpurun_out/prof_%s.raw.csv" % tag)))
        unit = dict(zip(rows[0], rows[1]))
        scale = lambda u: {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
        total = rd * scale(unit["dram__bytes_read.sum"]) + wr * scale(unit["dram__bytes_write.sum"])
        w, h = (3840, 2160) if "comp" in tag else (1920, 1080)
        line = [l for l in open("gpurun_out/prof_%s.bench.json" % tag) if l.startswith("{")][-1]     # (ncu prints its banner there too)
        fields = json.loads(line)["config"]["fields_per_step_per_gpu"]
        caps[key] = {"dram_bytes": total, "fields": fields, "grid": f(m["launch__grid_size"]),
                     "algorithmic_bytes": 8.0 * w * ((h + 1) // 2) * fields}
    except Exception as e:
        print("no capture for", tag, e)
hsh = hashlib.sha256()
for fn in ("lane_pipeline.cuh", "scanline_kernels.cuh"):
    hsh.update(open(os.path.join("composite_video_simulator_b200", "csrc", fn), "rb").read())
json.dump({"kernel_source_sha256": hsh.hexdigest()[:16], "captures": caps,
           "source": "one ncu --set full capture per preset (scripts/round_end_gpu.sh): dram__bytes_read.sum + dram__bytes_write.sum"},
          open("gpurun_out/ncu_traffic.json", "w"), indent=1)
d = json.load(open('gpurun_out/bench_r2_n1.json'))
print('N1 value %.0f e2e %.0f frac %.4f cpu %.2f parity %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['parity']))
PY
