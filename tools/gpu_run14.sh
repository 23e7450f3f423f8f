set -x
python -m pytest tests -m gpu -q -x -k "niir or secam or pal or ntsc" > gpurun_out/r2_tests14.log 2>&1; tail -5 gpurun_out/r2_tests14.log
for k in niir secam pald ntsc3d secam1080 pald1080; do f=256; case $k in *1080) f=64;; esac; python tools/kt.py $k $f; done 2>&1 | tee gpurun_out/r2_kt14.log
