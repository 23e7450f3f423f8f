#!/bin/bash
# Tuning aid: build variants of the library with extra -D flags into tools/variants/<name>.so (objects under build/variants/<name>/)
#   tools/variants.sh name1 "-DFOO=1 -DBAR=2" name2 "..."      (then: CM_B200_LIB=tools/variants/name1.so python tools/kt.py)
set -e
cd "$(dirname "$0")/.."
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-extended-lambda -Xcompiler -fPIC"
mkdir -p build/variants
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  d=build/variants/$name; mkdir -p $d
  (
    nvcc $FLAGS $defs -c -o $d/api.o color_modem_b200/csrc/cm_api.cu &
    nvcc $FLAGS $defs -c -o $d/util.o color_modem_b200/csrc/cm_util.cu &
    for t in F32 F64; do
      for f in mac niir secam; do nvcc $FLAGS $defs -DCM_INST_$t -c -o $d/${f}_$t.o color_modem_b200/csrc/cm_$f.cu & done
      for part in 0 1 2; do nvcc $FLAGS $defs -DCM_INST_$t -DCM_QAM_PART=$part -c -o $d/qam_${part}_$t.o color_modem_b200/csrc/cm_qam.cu & done
    done
    nvcc $FLAGS $defs -DCM_QAM_PART=3 -c -o $d/qam_3.o color_modem_b200/csrc/cm_qam.cu &
    wait
  ) 2>&1 | grep -E "error" || true
  mkdir -p tools/variants
  nvcc -shared -o tools/variants/$name.so $d/*.o
  echo built tools/variants/$name.so
done
