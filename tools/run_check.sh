O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu17.log 2>&1; echo pytest rc=$?; tail -3 $O/pytest_gpu17.log
for c in 3 2 4; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --profile-layers 2> $O/layers_c${c}_v20.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['baseline_config'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['roofline']['frac'])"
done
grep "dgrad" $O/layers_c3_v20.txt | head -20
