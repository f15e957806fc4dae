import sys, os, subprocess, ctypes as C, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import helpers, composite_video_simulator_b200 as cvs
from test_gpu_cli import reference_loop, TOOL
oracle = helpers.load_oracle()
for (w,h,delay,batch,argv) in [(160,120,1,3,["-vhs"]),(164,121,3,4,[])]:
    frames=[helpers.stream_frame(w,h,k) for k in range(4)]
    np.stack(frames).tofile('/tmp/in.bgra')
    subprocess.run([TOOL,"-i","/tmp/in.bgra","-o","/tmp/out.bgra","-width",str(w),"-height",str(h),"-d",str(delay),"-batch",str(batch),"-fields-per-frame","2","-double"]+argv,check=True,stderr=subprocess.DEVNULL)
    got=np.fromfile('/tmp/out.bgra',dtype=np.uint32).reshape(-1,h,w)
    p=helpers.params(*(["-width",str(w)]+argv))
    want=reference_loop(oracle,p,frames,w,h,2,delay)
    for k in range(len(want)):
        bad=np.nonzero((got[k]!=want[k]).any(axis=1))[0]
        print(w,h,'delay',delay,'pic',k,'field',(k&1)^1,'bad rows',bad.tolist()[:12], len(bad))
    # engine directly: batch host call with bob
    src=np.stack([frames[k//2] for k in range(batch)])
    dst=np.zeros_like(src)
    with cvs.Engine(params=p,max_w=w,max_h=h,max_batch=batch) as eng:
        eng.set_precision(True); eng.set_bob(True)
        eng.composite_fields_host(dst,src,0)
    # sequential
    seq=np.zeros_like(src)
    with cvs.Engine(params=p,max_w=w,max_h=h,max_batch=batch) as eng:
        eng.set_precision(True); eng.set_bob(True)
        for k in range(batch): eng.composite_layer(seq[k],src[k],(k&1)^1,k)
    for k in range(batch):
        bad=np.nonzero((dst[k]!=seq[k]).any(axis=1))[0]
        print('   engine batch vs sequential pic',k,'bad rows',bad.tolist()[:12])
