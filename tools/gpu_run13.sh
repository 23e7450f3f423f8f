set -x
python -m pytest tests -m gpu -q > gpurun_out/r2_tests13.log 2>&1; tail -5 gpurun_out/r2_tests13.log
python tools/sweep.py --frames 256 --json gpurun_out/r2_sweep13.json 2>&1 | tee gpurun_out/r2_sweep13.log
ncu --set full --clock-control none --import-source on -k regex:k_niir_decode2 -s 2 -c 1 -o gpurun_out/r2_prof_niir2 python tools/kt.py niir 64 > gpurun_out/r2_ncu13.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_secam_decode2 -s 2 -c 1 -o gpurun_out/r2_prof_secam2 python tools/kt.py secam 64 >> gpurun_out/r2_ncu13.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_secam_encode_row2 -s 2 -c 1 -o gpurun_out/r2_prof_secam_enc2 python tools/kt.py secam 64 >> gpurun_out/r2_ncu13.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_qam_encode_row2 -s 2 -c 1 -o gpurun_out/r2_prof_qam_enc2 python tools/kt.py pald 64 >> gpurun_out/r2_ncu13.log 2>&1
tail -2 gpurun_out/r2_ncu13.log
