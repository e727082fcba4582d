O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu_final.log 2>&1; echo pytest rc=$?; tail -2 $O/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > $O/bench_default_final.json 2>/dev/null; echo default rc=$?
for c in 4 5; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c${c}_final.json 2>/dev/null; done
timeout 600 python bench.py --config 3 --steps 20 --warmup 5 > $O/bench_c3_final.json 2>/dev/null
timeout 600 python bench.py --config 2 --steps 20 --warmup 5 > $O/bench_c2_final.json 2>/dev/null
python - <<'PY'
import json
for f in ["default","c2","c3","c4","c5"]:
    d=json.loads(open(f"gpurun_out/bench_{f}_final.json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["frac"],3), (d.get("training") or {}).get("value"))
PY
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/plain_col3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"col3" -s 1 -c 2 -o $O/prof_col3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_col3.log 2>&1
echo ncu rc=$?
ncu -i $O/prof_col3.ncu-rep --page raw --csv > $O/prof_col3_raw.csv 2>/dev/null; rm -f $O/prof_col3.ncu-rep
