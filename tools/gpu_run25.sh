timeout 300 python -m pytest tests -m gpu -q -x -k "niir" > gpurun_out/r2_tests25.log 2>&1; tail -3 gpurun_out/r2_tests25.log
timeout 120 python tools/kt.py niir 256 2>&1 | tee gpurun_out/r2_kt25.log
