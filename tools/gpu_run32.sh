timeout 1500 compute-sanitizer --tool racecheck --print-limit 200 python tools/sanitize.py > gpurun_out/r2_racecheck2.log 2>&1; grep -E "^========= (Error|Warning)" gpurun_out/r2_racecheck2.log | sed 's/+0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -12; grep "RACECHECK SUMMARY" gpurun_out/r2_racecheck2.log
for k in pald ntsc3d secam; do timeout 120 python tools/kt.py $k 256; done 2>&1 | cut -c1-200
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests32.log 2>&1; tail -2 gpurun_out/r2_tests32.log
