O=gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --config 3 --steps 10 --warmup 3 > $O/bench_train_n$N.json 2> $O/bench_train_n$N.err; echo train rc=$?; cut -c1-400 $O/bench_train_n$N.json; tail -3 $O/bench_train_n$N.err | cut -c1-300
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_infer_n$N.json 2> $O/bench_infer_n$N.err; echo infer rc=$?; cut -c1-400 $O/bench_infer_n$N.json
timeout 300 $TR bench.py --gpus $N --config 3 --steps 10 --warmup 3 --comm-ctas 8 > $O/bench_train_n${N}_c8.json 2>/dev/null; echo train8 rc=$?; cut -c1-230 $O/bench_train_n${N}_c8.json
timeout 300 $TR tools/dp_parity.py > $O/dp_parity_n$N.log 2>&1; echo parity rc=$?; grep dp_parity $O/dp_parity_n$N.log
