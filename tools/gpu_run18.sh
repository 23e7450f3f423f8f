set -x
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests18.log 2>&1; tail -4 gpurun_out/r2_tests18.log
for k in pald ntsc3d secam niir proto pald1080 secam1080; do f=256; case $k in *1080) f=64;; esac; timeout 120 python tools/kt.py $k $f; done 2>&1 | tee gpurun_out/r2_kt18.log
timeout 200 python tools/latency.py --reps 100 > gpurun_out/r2_latency18.log 2>&1; cut -c1-260 gpurun_out/r2_latency18.log | head -3
