run() { python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity 8 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); r=d['roofline']
print(sys.argv[1:], round(d['value']), 'q/s', round(d['ms_per_step'],1), 'ms  e2e', round(d['e2e']['value']), d['e2e']['host_ms'], 'hot', r['hot_terms'], round(r['hot_decode_ms'],1), {k:round(v,1) for k,v in r['class_ms'].items() if v}, d['parity_sample']['mismatches'])" "$@"; }
run --opt eager_hot=1
run --opt hot_div=100
run --opt hot_div=50
run --opt hot_div=400
run --opt hot_min_uses=3
run --opt hot_min_uses=5
