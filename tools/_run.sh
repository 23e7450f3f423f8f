timeout 400 python -m pytest tests -m gpu -q -x -k "mac" > gpurun_out/r2_tests41.log 2>&1; tail -3 gpurun_out/r2_tests41.log
for r in 4 2 1; do for k in mac mac7 mac1080; do f=256; case $k in *1080) f=64;; esac; echo "rows_max=$r"; CM_ROWS_MAX=$r timeout 120 python tools/kt.py $k $f; done; done 2>&1 | cut -c1-170
