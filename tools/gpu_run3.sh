python -m pytest tests -m gpu -q -x -k "not fullsize" > gpurun_out/r2_tests3.log 2>&1; tail -25 gpurun_out/r2_tests3.log
for k in pald ntsc3d pal3d; do CM_ROWS_V1=1 python tools/kt.py $k 256; python tools/kt.py $k 256; done 2>&1 | tee gpurun_out/r2_kt3.log
for k in pald1080 ntsc3d1080; do CM_ROWS_V1=1 python tools/kt.py $k 64; python tools/kt.py $k 64; done 2>&1 | tee -a gpurun_out/r2_kt3.log
