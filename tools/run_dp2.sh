O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/dp_parity.py > $O/dp_parity_n2.log 2>&1; echo parity rc=$?; grep dp_parity $O/dp_parity_n2.log; tail -3 $O/dp_parity_n2.log | cut -c1-300
for ctas in 0 4 8 16; do
  if [ $ctas = 0 ]; then EXTRA="--comm-ctas 32"; export MAU_SM_RESERVE=0; else EXTRA="--comm-ctas $ctas"; unset MAU_SM_RESERVE; fi
  timeout 300 $TR bench.py --gpus 2 --config 3 --steps 10 --warmup 3 $EXTRA > $O/bench_train_n2_c$ctas.json 2> $O/bench_train_n2_c$ctas.err; echo ctas=$ctas rc=$?
  python -c "
import json,sys
d=json.loads(open('$O/bench_train_n2_c$ctas.json').read().strip().splitlines()[-1]); print('ctas $ctas', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
timeout 300 $TR bench.py --gpus 2 --config 3 --steps 10 --warmup 3 --sync-bn > $O/bench_train_n2_syncbn.json 2> $O/bench_train_n2_syncbn.err; echo syncbn rc=$?; cut -c1-200 $O/bench_train_n2_syncbn.json
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_infer_n2_v11.json 2>/dev/null; cut -c1-200 $O/bench_infer_n2_v11.json
