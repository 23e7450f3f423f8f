timeout 400 python -m pytest tests -m gpu -q -x -k "mac" > gpurun_out/r2_tests43.log 2>&1; tail -2 gpurun_out/r2_tests43.log
for k in mac mac7 mac1080; do f=256; case $k in *1080) f=64;; esac; timeout 120 python tools/kt.py $k $f; done 2>&1 | cut -c1-190
