O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu18.log 2>&1; echo pytest rc=$?; tail -3 $O/pytest_gpu18.log
timeout 300 python tools/bw_bench.py > $O/bw_bench4.txt 2>&1; cat $O/bw_bench4.txt
for c in 3 4; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['baseline_config'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['roofline']['frac'])"
done
