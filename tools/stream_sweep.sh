for cfg in "4 50" "6 50" "8 50" "4 25" "8 25" "4 100" "2 100"; do
  set -- $cfg
  python bench.py --steps 6 --warmup 3 --streams $1 --batch $2 --no-extra --no-cpu-baseline --no-determinism 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('streams $1 batch $2 value %.0f e2e %.0f ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
