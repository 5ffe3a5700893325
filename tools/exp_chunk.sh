run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(sys.argv[1:], 'step %.2f ms contract %.2f (launches %d) e2e %.2f ms (contract %.2f) clocks %s' % (d['ms_per_step'], d['stages_ms']['contract'], d['roofline']['launches_per_step'], d['e2e']['ms_per_step'], d['e2e']['device_ms']['ms_contract'], d['clocks']))" "$@"; }
for a in "$@"; do run $a; done
