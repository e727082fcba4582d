# multi-GPU evidence run: gpurun --gpus N -- bash tools/run_dp8.sh N
O=gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_default_n$N.json 2> $O/bench_default_n$N.err; echo default rc=$?
timeout 400 $TR bench.py --gpus $N --config 4 --steps 10 --warmup 3 > $O/bench_c4_n$N.json 2> $O/bench_c4_n$N.err; echo c4 rc=$?
timeout 300 $TR tools/dp_parity.py > $O/dp_parity_n$N.log 2>&1; echo parity rc=$?; grep dp_parity $O/dp_parity_n$N.log | cut -c1-330
python - <<PY
import json
d=json.loads(open("$O/bench_default_n$N.json").read().strip().splitlines()[-1])
print("infer", d["n_gpus"], round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]))
t=d["training"]; print("train", round(t["value"]), t["ms_per_step"], "e2e", round(t["e2e"]["value"]), t["config"]["parallelism"])
d=json.loads(open("$O/bench_c4_n$N.json").read().strip().splitlines()[-1])
print("unet++ train", d["n_gpus"], round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]))
PY
