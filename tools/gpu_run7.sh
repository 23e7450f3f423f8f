set -x
python -m pytest tests -m gpu -q -x -k "mac" > gpurun_out/r2_tests7.log 2>&1; tail -4 gpurun_out/r2_tests7.log
for k in mac mac7 mac1080; do f=256; case $k in *1080) f=64;; esac; python tools/kt.py $k $f; done 2>&1 | tee gpurun_out/r2_kt7.log
