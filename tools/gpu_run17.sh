set -x
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_tests17.log 2>&1; tail -4 gpurun_out/r2_tests17.log
timeout 400 python bench.py > gpurun_out/r2_bench17.json 2> gpurun_out/r2_bench17.err; echo rc=$?; tail -3 gpurun_out/r2_bench17.err; cut -c1-400 gpurun_out/r2_bench17.json
timeout 400 python bench.py --impl reference > gpurun_out/r2_bench17_ref.json 2>> gpurun_out/r2_bench17.err; cut -c1-300 gpurun_out/r2_bench17_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 60 --csv --log-file gpurun_out/r2_launches17.csv python bench.py --steps 2 --warmup 1 --frames 512 --no-extras --no-cpu --e2e-seconds 0.05 > gpurun_out/r2_launches17.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_qam_rows2|k_qam_combine|k_qam_encode_row2" -s 6 -c 3 -o gpurun_out/r2_prof_pald_v17 python tools/kt.py pald 64 > gpurun_out/r2_ncu17.log 2>&1
timeout 200 python tools/latency.py --json gpurun_out/r2_latency17.json > gpurun_out/r2_latency17.log 2>&1; tail -3 gpurun_out/r2_latency17.log | cut -c1-300
