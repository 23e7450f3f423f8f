set -x
python -m pytest tests -m gpu -q -x -k "mac" > gpurun_out/r2_tests9.log 2>&1; tail -4 gpurun_out/r2_tests9.log
for k in mac mac7 mac1080; do f=256; case $k in *1080) f=64;; esac; python tools/kt.py $k $f; done 2>&1 | tee gpurun_out/r2_kt9.log
ncu --set full --clock-control none --import-source on -k regex:k_mac_encode -s 2 -c 1 -o gpurun_out/r2_prof_mac1080_v2 python tools/kt.py mac1080 16 > gpurun_out/r2_ncu9.log 2>&1
