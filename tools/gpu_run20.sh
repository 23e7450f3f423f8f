for v in "" tools/variants/cmb8.so tools/variants/rows10.so; do
  for k in pald ntsc3d; do CM_B200_LIB=$v timeout 120 python tools/kt.py $k 256; done
done 2>&1 | tee gpurun_out/r2_kt20.log
