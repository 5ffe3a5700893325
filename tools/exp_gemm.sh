run() { timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(sys.argv[1:], 'step %.2f ms contract %.2f e2e %.2f ms stages %s' % (d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['ms_per_step'], d['stages_ms']))" "$@"; }
for v in "$@"; do run --pair 0 --debug-no-tma $v; done
