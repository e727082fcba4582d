O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu13.log 2>&1; echo pytest rc=$?; tail -3 $O/pytest_gpu13.log
for c in 2 3; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --profile-layers > $O/bench_c${c}_v15.json 2> $O/layers_c${c}_v15.txt; echo config $c rc=$?; cut -c1-330 $O/bench_c${c}_v15.json
done
grep "k:" $O/layers_c2_v15.txt | awk '{a+=$2; print} END {print a}'
