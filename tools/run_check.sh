O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu14.log 2>&1; echo pytest rc=$?; tail -3 $O/pytest_gpu14.log
for c in 4 5 3; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c${c}_v17.json 2> /dev/null; echo config $c rc=$?; cut -c1-230 $O/bench_c${c}_v17.json
done
