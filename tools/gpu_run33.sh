TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench33_n8.json 2> gpurun_out/r2_bench33.err; echo rc=$?
timeout 200 $TR bench.py --gpus 8 --workload ntsc3d600 > gpurun_out/r2_bench33_ntsc_n8.json 2>> gpurun_out/r2_bench33.err; echo rc=$?
timeout 300 $TR bench.py --gpus 8 --workload sweep1080 --steps 5 --warmup 3 > gpurun_out/r2_bench33_sweep_n8.json 2>> gpurun_out/r2_bench33.err; echo rc=$?
python - <<'PY'
import json
def load(p):
    for ln in open(p):
        if ln.startswith('{'): return json.loads(ln)
d=load('gpurun_out/r2_bench33_n8.json'); e=d['e2e']
print('N8 value %.0f e2e %.0f seq %.0f trans %.0f frac %.3f peak %.1f'%(d['value'],e['value'],e['sequential'],e['transcode']['value'],e['frac_of_copy_peak'],e['copy_peak']['min_over_ranks_both_each_way_gbs']))
print([(o['workload'][:28], round(o['frames_per_s'])) for o in d['other_workloads']])
print('ntsc3d600 N8', load('gpurun_out/r2_bench33_ntsc_n8.json')['value'])
print('sweep1080 N8', load('gpurun_out/r2_bench33_sweep_n8.json')['value'])
PY
tail -3 gpurun_out/r2_bench33.err
