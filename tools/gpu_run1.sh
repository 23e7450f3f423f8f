set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests1.log 2>&1; tail -15 gpurun_out/r2_tests1.log
for k in pald ntsc3d ntsc secam niir proto mac mac7; do python tools/kt.py $k 256; done > gpurun_out/r2_kt0.log 2>&1
for k in pald1080 ntsc3d1080 secam1080 proto1080 mac1080; do python tools/kt.py $k 64; done >> gpurun_out/r2_kt0.log 2>&1
cat gpurun_out/r2_kt0.log
ncu --set full --clock-control none --import-source on -k regex:k_qam_rows -s 2 -c 1 -o gpurun_out/r2_prof_pald1080 python tools/kt.py pald1080 16 > gpurun_out/r2_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mac -s 2 -c 2 -o gpurun_out/r2_prof_mac1080 python tools/kt.py mac1080 16 >> gpurun_out/r2_ncu1.log 2>&1
tail -3 gpurun_out/r2_ncu1.log
