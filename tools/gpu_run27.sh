timeout 300 python bench.py --steps 5 --no-extras --no-cpu > gpurun_out/r2_bench27_n1.json 2> gpurun_out/r2_bench27.err; python - <<'PY'
import json
for ln in open('gpurun_out/r2_bench27_n1.json'):
    if ln.startswith('{'):
        d=json.loads(ln); e=d['e2e']; print('N1 value %.0f e2e %.0f seq %.0f trans %.0f frac %.3f node %s peak %s'%(d['value'],e['value'],e['sequential'],e['transcode']['value'],e['frac_of_copy_peak'],e['numa_node_of_rank0'],e['copy_peak']['both_each_way_gbs']))
PY
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --no-extras --no-cpu > gpurun_out/r2_bench27_n2.json 2>> gpurun_out/r2_bench27.err; python - <<'PY'
import json
for ln in open('gpurun_out/r2_bench27_n2.json'):
    if ln.startswith('{'):
        d=json.loads(ln); e=d['e2e']; print('N2 value %.0f e2e %.0f seq %.0f trans %.0f frac %.3f node %s peak %s'%(d['value'],e['value'],e['sequential'],e['transcode']['value'],e['frac_of_copy_peak'],e['numa_node_of_rank0'],e['copy_peak']['both_each_way_gbs']))
PY
tail -3 gpurun_out/r2_bench27.err
lscpu | grep -i numa; nvidia-smi topo -m 2>/dev/null | head -8
