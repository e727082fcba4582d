# final single-GPU artefacts of the round (gpurun -- bash tools/run_final.sh)
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu_final.log 2>&1; echo pytest rc=$?; tail -2 $O/pytest_gpu_final.log
timeout 300 python tools/bw_bench.py --json $O/bw_bench_final.json > $O/bw_bench_final.txt 2>&1; echo bw rc=$?
timeout 300 python tools/conv_bench.py wgrad > $O/wgrad_bench_final.txt 2>&1; echo wgrad rc=$?
timeout 600 python bench.py --steps 20 --warmup 5 --profile-layers > $O/bench_c2_final.json 2> $O/layers_c2_final.txt; echo c2 rc=$?
timeout 600 python bench.py --config 3 --steps 20 --warmup 5 --profile-layers > $O/bench_c3_final.json 2> $O/layers_c3_final.txt; echo c3 rc=$?
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c4_final.json 2>/dev/null; echo c4 rc=$?
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c5_final.json 2>/dev/null; echo c5 rc=$?
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_final.json 2>/dev/null; echo ref rc=$?
for f in c2 c3 c4 c5 ref; do python - <<PY
import json
d=json.loads(open("$O/bench_${f}_final.json").read().strip().splitlines()[-1])
print("$f", round(d["value"],1), d.get("ms_per_step"), "e2e", round(d["e2e"]["value"],1), "roof", d.get("roofline",{}).get("frac"), d.get("cpu_baseline"))
PY
done
