for w in secam niir proto mac pald; do
  for r in 1 2 3 4 8; do for mw in 2 4; do
  echo -n "R=$r minw=$mw "; CM_ROWS_MAX=$r CM_MIN_WARPS=$mw python tools/kt.py $w 2>&1 | sed 's/default //; s/(. launches.step)//g' | cut -c1-150
  done; done
done
