set -x
python -m pytest tests -m gpu -q -x -k "mac or secam" > gpurun_out/r2_tests10.log 2>&1; tail -6 gpurun_out/r2_tests10.log
for k in mac mac7 mac1080 secam secam1080; do f=256; case $k in *1080) f=64;; esac; python tools/kt.py $k $f; CM_ROWS_V1=1 python tools/kt.py $k $f; done 2>&1 | tee gpurun_out/r2_kt10.log
