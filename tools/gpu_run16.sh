set -x
python -m pytest tests -m gpu -q > gpurun_out/r2_tests16.log 2>&1; tail -6 gpurun_out/r2_tests16.log
python bench.py > gpurun_out/r2_bench16.json 2> gpurun_out/r2_bench16.err; echo rc=$?; tail -3 gpurun_out/r2_bench16.err; cut -c1-1500 gpurun_out/r2_bench16.json
python bench.py --impl reference > gpurun_out/r2_bench16_ref.json 2>> gpurun_out/r2_bench16.err; cut -c1-300 gpurun_out/r2_bench16_ref.json
