O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu15.log 2>&1; echo pytest rc=$?; tail -3 $O/pytest_gpu15.log
for c in 2 3; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --profile-layers > $O/bench_c${c}_v18.json 2> $O/layers_c${c}_v18.txt; echo config $c rc=$?; cut -c1-230 $O/bench_c${c}_v18.json
done
grep "k:conv0_" $O/layers_c2_v18.txt; grep "k:conv0_" $O/layers_c3_v18.txt
MAU_NO_BRES=1 timeout 600 python bench.py --config 2 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200
