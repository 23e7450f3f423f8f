set -x
python -m pytest tests -m gpu -q > gpurun_out/r2_tests4.log 2>&1; tail -25 gpurun_out/r2_tests4.log
for k in pald1080 ntsc3d1080; do CM_ROWS_V1=1 python tools/kt.py $k 64; python tools/kt.py $k 64; done 2>&1 | tee gpurun_out/r2_kt4.log
for k in pald ntsc3d; do python tools/kt.py $k 256; done 2>&1 | tee -a gpurun_out/r2_kt4.log
ncu --set full --clock-control none --import-source on -k regex:k_qam_rows2 -s 2 -c 1 -o gpurun_out/r2_prof_rows2_pald python tools/kt.py pald 64 > gpurun_out/r2_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_qam_rows2 -s 2 -c 1 -o gpurun_out/r2_prof_rows2_pald1080 python tools/kt.py pald1080 16 >> gpurun_out/r2_ncu4.log 2>&1
tail -3 gpurun_out/r2_ncu4.log
