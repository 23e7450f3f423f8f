set -x
python -m pytest tests -m gpu -q -k "niir or mac" > gpurun_out/r2_tests12.log 2>&1; tail -12 gpurun_out/r2_tests12.log
for k in niir mac; do python tools/kt.py $k 256; CM_ROWS_V1=1 python tools/kt.py $k 256; done 2>&1 | tee gpurun_out/r2_kt12.log
python tests/diag_gpu.py niir 2>&1 | tee gpurun_out/r2_diag_niir.log
python tools/latency.py --json gpurun_out/r2_latency.json 2>&1 | tee gpurun_out/r2_latency.log
