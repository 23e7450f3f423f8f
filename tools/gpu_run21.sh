timeout 600 python -m pytest tests -m gpu -q -x -k "ntsc or pal_s or image or kit or perline" > gpurun_out/r2_tests21.log 2>&1; tail -4 gpurun_out/r2_tests21.log
for k in ntsc pals; do timeout 120 python tools/kt.py $k 256; CM_ROWS_V1=1 timeout 120 python tools/kt.py $k 256; done 2>&1 | tee gpurun_out/r2_kt21.log
