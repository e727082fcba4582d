O=gpurun_out
timeout 900 python bench.py > $O/bench_default_final.json 2> $O/bench_default_final.err; echo rc=$?; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_default_final.json").read().strip().splitlines()[-1])
print(d["metric"], round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["frac"],3), d["clocks"], d["cpu_baseline"])
t=d["training"]; print(t["metric"], round(t["value"]), t["ms_per_step"], "e2e", round(t["e2e"]["value"]), "roof", round(t["roofline"]["frac"],3), t["clocks"])
PY
wc -l $O/bench_default_final.json
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-160
