set -x
python -m pytest tests -m gpu -q -x -k "mac or pal or ntsc" > gpurun_out/r2_tests8.log 2>&1; tail -4 gpurun_out/r2_tests8.log
for k in mac mac7 mac1080 pald ntsc pald1080 ntsc3d1080; do f=256; case $k in *1080) f=64;; esac; python tools/kt.py $k $f; CM_ROWS_V1=1 python tools/kt.py $k $f; done 2>&1 | tee gpurun_out/r2_kt8.log
