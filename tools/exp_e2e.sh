run() { timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(sys.argv[1:], 'step %.2f ms contract %.2f e2e %.2f ms each %s dev %s' % (d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['ms_per_step'], d['e2e']['ms_each_step'], d['e2e']['device_ms']))" "$@"; }
run --pair 1
run --pair 0
run --pair 1
run --pair 0
