python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/kt.py secam | cut -c1-200
CM_B200_LIB=tools/variants/secam_p2.so python tools/kt.py secam | cut -c1-200
python tools/kt.py niir | cut -c1-200
CM_B200_LIB=tools/variants/secam_p2.so python -m pytest tests -m gpu -x -q -k secam 2>&1 | tail -2
