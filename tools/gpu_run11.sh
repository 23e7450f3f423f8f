set -x
python -m pytest tests -m gpu -q -x -k "mac or secam" > gpurun_out/r2_tests11.log 2>&1; tail -6 gpurun_out/r2_tests11.log
for k in mac mac1080 secam secam1080; do f=256; case $k in *1080) f=64;; esac; python tools/kt.py $k $f; done 2>&1 | tee gpurun_out/r2_kt11.log
