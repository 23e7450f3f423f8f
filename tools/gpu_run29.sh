for r in 1 2 3 4 6; do for k in pald ntsc3d; do CM_RPC=$r timeout 120 python tools/kt.py $k 256 | sed "s/^/rpc=$r /"; done; done 2>&1 | tee gpurun_out/r2_kt29.log
