run() { timeout 600 python bench.py --workload cfg5_atlas_shard --cells 10240 --steps 3 --warmup 2 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(sys.argv[1:], 'step %.2f ms contract %.2f (%.0f TOP/s) e2e %.2f ms stages %s clocks %s' % (d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['achieved'], d['e2e']['ms_per_step'], d['stages_ms'], d['clocks']))" "$@"; }
run --pair 0
run --pair 0 --debug-no-tma 1
run --pair 1
