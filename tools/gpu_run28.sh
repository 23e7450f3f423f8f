for v in "" tools/variants/firfast.so; do
  for k in pald ntsc3d ntsc secam; do CM_B200_LIB=$v timeout 120 python tools/kt.py $k 256; done
done 2>&1 | tee gpurun_out/r2_kt28.log
CM_B200_LIB=tools/variants/firfast.so timeout 120 python tools/ab.py pald 64; timeout 120 python tools/ab.py pald 64
