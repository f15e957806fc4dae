import json,subprocess,sys,os
for v in sys.argv[1:]:
    env=dict(os.environ)
    if v!="default": env["CVS_NTSC_LIB"]=os.path.abspath("variants/libcvs_%s.so"%v)
    for i in range(2):
        out=subprocess.run([sys.executable,"scripts/bench_yuv422.py","--steps","5","--warmup","3","--cpu-fields","0"],env=env,capture_output=True,text=True).stdout
        d=json.loads(out.strip().splitlines()[-1])
        print("%-8s yuv422 value %.0f kernel_ms %.3f GB/s %.1f"%(v,d["value"],d["kernel_ms_per_step"],d["algorithmic_GBps"]))
