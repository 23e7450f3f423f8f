python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in niir proto secam; do python tools/kt.py $w | cut -c1-230; done
for r in 1 2; do echo -n "R=$r "; CM_ROWS_MAX=$r python tools/kt.py proto | cut -c1-200; done
for r in 1 2 3; do echo -n "R=$r "; CM_ROWS_MAX=$r python tools/kt.py niir | cut -c1-200; done
