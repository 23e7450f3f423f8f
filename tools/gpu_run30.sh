timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_tests30.log 2>&1; tail -3 gpurun_out/r2_tests30.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench30.json 2> gpurun_out/r2_bench30.err; echo rc=$?; cut -c1-200 gpurun_out/r2_bench30.json
timeout 200 python tools/latency.py --reps 100 --json gpurun_out/r2_latency30.json > gpurun_out/r2_latency30.log 2>&1; python - <<'PY'
import json
for r in json.load(open('gpurun_out/r2_latency30.json')):
    print(r['modem'][:40], 'rpc', r['rows_per_cta'], 'dev %.1f graph %.1f host %.1f' % (r['device_us']['median'], r['graph_us'].get('median', -1), r['host_us']['median']))
PY
