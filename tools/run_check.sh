# quick single-GPU check of a change (gpurun -- bash tools/run_check.sh): parity suite, smoke, the two headline configs
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu_check.log 2>&1; echo pytest rc=$?; tail -2 $O/pytest_gpu_check.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for c in 2 3; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --profile-layers 2> $O/layers_c${c}_check.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['baseline_config'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['roofline']['frac'])"
done
