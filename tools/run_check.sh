O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu22.log 2>&1; echo pytest rc=$?; tail -2 $O/pytest_gpu22.log
BW_ONLY=bn_bwd_apply,bn_apply_relu timeout 300 python tools/bw_bench.py 2>&1 | grep bn_
timeout 600 python bench.py --config 3 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
