for d in 4 8; do
  echo "== pipeline $d"
  LINNE_B200_PIPELINE=$d python bench.py --c5-files 0 --no-streaming --no-refine --no-inlib --c4-seconds 0 --steps 3 --warmup 1 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c3=d['c3_decode']; print({k:c3[k] for k in c3 if k in ('ms','e2e','e2e_packed')})"
done
