python -m pytest tests -m gpu -x -q -k "negated or cfg4 or dnf or cfg2 or fuzz or boolean" 2>&1 | tail -3
run() { python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity 48 "$@" 2>gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); r=d['roofline']
print(sys.argv[1:], round(d['value']), 'q/s', round(d['ms_per_step'],1), 'ms  e2e', round(d['e2e']['value']), 'K0', round(r['hot_decode_ms'],1), r['hot_terms'], {k:round(v,1) for k,v in r['class_ms'].items() if v}, d['parity_sample']['mismatches'], d['parity_sample_e2e']['mismatches'])" "$@"; }
run --opt eager_hot=1
