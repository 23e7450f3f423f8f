timeout 300 python -m pytest tests -m gpu -q -x -k "niir" > gpurun_out/r2_tests22.log 2>&1; tail -3 gpurun_out/r2_tests22.log
timeout 120 python tools/kt.py niir 256 2>&1 | tee gpurun_out/r2_kt22.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_secam_decode2|k_secam_encode_row2" -s 4 -c 2 -o gpurun_out/r2_prof_secam1080 python tools/kt.py secam1080 16 > gpurun_out/r2_ncu22.log 2>&1; tail -2 gpurun_out/r2_ncu22.log
