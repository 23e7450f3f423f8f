python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rpc in 1 2 4; do echo -n "RPC=$rpc "; CM_RPC=$rpc python tools/kt.py pald | cut -c1-120; done
python tools/kt.py ntsc3d | cut -c1-120
