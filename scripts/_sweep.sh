python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity 48 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); r=d['roofline']
print('cfg2', round(d['value']), 'q/s', round(d['ms_per_step'],1), 'ms  e2e', round(d['e2e']['value']), {k:round(v,1) for k,v in r['class_ms'].items() if v}, d['parity_sample']['mismatches'], d['parity_sample_e2e']['mismatches'])"
