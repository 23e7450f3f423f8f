timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests23.log 2>&1; tail -3 gpurun_out/r2_tests23.log
for k in pald ntsc3d ntsc secam niir proto pald1080 secam1080; do f=256; case $k in *1080) f=64;; esac; timeout 120 python tools/kt.py $k $f; done 2>&1 | tee gpurun_out/r2_kt23.log
