for lib in libcvs_ntsc.so libcvs_ntsc_k4b2.so libcvs_ntsc_k4b3.so libcvs_ntsc_k4b4.so; do for preset in sp comp; do
CVS_NTSC_LIB=$PWD/composite_video_simulator_b200/$lib python bench.py --preset $preset --steps 20 --warmup 3 --cpu-fields 0 --e2e-batch 16 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib $preset value %.0f kernel_ms %.3f frac %.4f'%(d['value'],d['roofline']['kernel_ms_per_launch'],d['roofline']['frac']))
"
done; done
CVS_NTSC_LIB=$PWD/composite_video_simulator_b200/libcvs_ntsc_k4b3.so python -m pytest tests -m gpu -x -q 2>&1 | tail -3
