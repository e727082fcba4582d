O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu9.log 2>&1; echo pytest rc=$?; tail -6 $O/pytest_gpu9.log
timeout 300 python tools/bw_bench.py --json $O/bw_bench2.json > $O/bw_bench2.txt 2>&1; cat $O/bw_bench2.txt
for c in 2 3; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --profile-layers > $O/bench_c${c}_v11.json 2> $O/layers_c${c}_v11.txt; echo config $c rc=$?; cut -c1-420 $O/bench_c${c}_v11.json
done
