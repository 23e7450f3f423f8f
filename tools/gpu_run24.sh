for k in pald ntsc3d pal3d ntsc secam1080; do f=256; case $k in *1080) f=64;; esac; timeout 120 python tools/kt.py $k $f; done 2>&1 | tee gpurun_out/r2_kt24.log
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests24.log 2>&1; tail -3 gpurun_out/r2_tests24.log
timeout 400 python bench.py > gpurun_out/r2_bench24.json 2> gpurun_out/r2_bench24.err; echo rc=$?; tail -2 gpurun_out/r2_bench24.err; cut -c1-260 gpurun_out/r2_bench24.json
timeout 400 python bench.py --impl reference > gpurun_out/r2_bench24_ref.json 2>> gpurun_out/r2_bench24.err; cut -c1-200 gpurun_out/r2_bench24_ref.json
timeout 300 python tools/sweep.py --frames 256 --json gpurun_out/r2_sweep24.json > gpurun_out/r2_sweep24.log 2>&1; tail -30 gpurun_out/r2_sweep24.log
