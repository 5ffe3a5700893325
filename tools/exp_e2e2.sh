run() { timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 30 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(sys.argv[1:], 'e2e %.2f ms' % d['e2e']['ms_per_step'], d['e2e']['ms_each_step'], d['e2e']['ms_inside_library'])" "$@"; }
for a in "$@"; do run $a; done
