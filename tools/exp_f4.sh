run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(sys.argv[1:], 'step %.2f ms contract %.2f (launches %d, frac %.2f) layout %.2f e2e %.2f ms (contract %.2f) clocks %s bits %s ov %s' % (d['ms_per_step'], d['stages_ms']['contract'], d['roofline']['launches_per_step'], d['roofline']['frac'], d['stages_ms']['counts_layout'], d['e2e']['ms_per_step'], d['e2e']['device_ms']['ms_contract'], d['clocks'], d['roofline'].get('operand_bits'), d['roofline'].get('overflow_columns_per_step')))" "$@"; }
for a in "$@"; do run $a; done
