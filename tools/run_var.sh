for lib in "" build/variants/nopf.so; do
  for w in pald ntsc3d ntsc secam niir; do CM_B200_LIB=$lib python tools/kt.py $w 2>&1 | sed 's/ | encode.*launches.step) | / | /'; done
done
