set -x
python -m pytest tests -m gpu -q -x -k "proto" > gpurun_out/r2_tests15.log 2>&1; tail -5 gpurun_out/r2_tests15.log
for k in proto proto1080; do f=256; case $k in *1080) f=64;; esac; python tools/kt.py $k $f; CM_ROWS_V1=1 python tools/kt.py $k $f; done 2>&1 | tee gpurun_out/r2_kt15.log
