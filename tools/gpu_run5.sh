set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests5.log 2>&1; tail -8 gpurun_out/r2_tests5.log
for k in pald ntsc3d pal3d; do for f in 256 1024; do CM_OVERLAP=0 python tools/kt.py $k $f; python tools/kt.py $k $f; CM_CHUNK=64 python tools/kt.py $k $f; done; done 2>&1 | tee gpurun_out/r2_kt5.log
for k in pald1080 ntsc3d1080; do CM_OVERLAP=0 python tools/kt.py $k 64; python tools/kt.py $k 64;  done 2>&1 | tee -a gpurun_out/r2_kt5.log
python bench.py > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo rc=$?; tail -3 gpurun_out/r2_bench5.err; cat gpurun_out/r2_bench5.json
python bench.py --workload ntsc3d600 > gpurun_out/r2_bench5_ntsc.json 2>> gpurun_out/r2_bench5.err; echo rc=$?; cat gpurun_out/r2_bench5_ntsc.json
