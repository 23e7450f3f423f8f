python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in niir proto; do python tools/kt.py $w | cut -c1-200; done
