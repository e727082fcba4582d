O=gpurun_out
for c in 2 3 4; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['baseline_config'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['roofline']['frac'])"
done
MAU_NO_COL3=1 timeout 600 python bench.py --config 4 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nocol3', d['config']['baseline_config'], round(d['value']), d['ms_per_step'])"
