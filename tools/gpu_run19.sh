set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench19_n2.json 2> gpurun_out/r2_bench19_n2.err; echo rc=$?; tail -3 gpurun_out/r2_bench19_n2.err; cut -c1-300 gpurun_out/r2_bench19_n2.json
timeout 300 $TR bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench19_n2_ref.json 2>> gpurun_out/r2_bench19_n2.err; cut -c1-200 gpurun_out/r2_bench19_n2_ref.json
timeout 300 $TR bench.py --gpus 2 --workload ntsc3d600 > gpurun_out/r2_bench19_ntsc_n2.json 2>> gpurun_out/r2_bench19_n2.err; cut -c1-300 gpurun_out/r2_bench19_ntsc_n2.json
timeout 400 $TR bench.py --gpus 2 --workload sweep1080 --steps 5 --warmup 3 > gpurun_out/r2_bench19_sweep_n2.json 2>> gpurun_out/r2_bench19_n2.err; cut -c1-300 gpurun_out/r2_bench19_sweep_n2.json
timeout 300 python bench.py --workload sweep1080 --steps 5 --warmup 3 > gpurun_out/r2_bench19_sweep_n1.json 2>> gpurun_out/r2_bench19_n2.err; cut -c1-300 gpurun_out/r2_bench19_sweep_n1.json
timeout 200 python bench.py --workload ntsc3d600 > gpurun_out/r2_bench19_ntsc_n1.json 2>> gpurun_out/r2_bench19_n2.err; cut -c1-200 gpurun_out/r2_bench19_ntsc_n1.json
tail -5 gpurun_out/r2_bench19_n2.err
