O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu10.log 2>&1; echo pytest rc=$?; tail -6 $O/pytest_gpu10.log
for c in 3 2; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --profile-layers > $O/bench_c${c}_v12.json 2> $O/layers_c${c}_v12.txt; echo config $c rc=$?; cut -c1-330 $O/bench_c${c}_v12.json
done
