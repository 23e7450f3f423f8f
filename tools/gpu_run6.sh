set -x
python -m pytest tests -m gpu -q -x -k "pal or ntsc" > gpurun_out/r2_tests6.log 2>&1; tail -4 gpurun_out/r2_tests6.log
for k in pald ntsc3d pal3d pald1080 ntsc3d1080; do f=256; case $k in *1080) f=64;; esac; python tools/kt.py $k $f; done 2>&1 | tee gpurun_out/r2_kt6.log
