O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 tools/dp_parity.py > $O/r02n8_dp_parity.txt 2>&1; grep dp_parity $O/r02n8_dp_parity.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29626 bench.py --gpus 8 --config 2 --no-cpu-baseline --sustain-s 1 > $O/r02n8_bench_c2.json 2> /dev/null; echo "bench n8 c2 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29627 bench.py --gpus 4 --config 2 --no-cpu-baseline --sustain-s 0 > $O/r02n4_bench_c2.json 2> /dev/null; echo "bench n4 c2 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29628 bench.py --gpus 8 --config 3 --no-cpu-baseline --sustain-s 1 > $O/r02n8_bench_c3.json 2> /dev/null; echo "bench n8 c3 rc=$?"
python -c "
import json
for f in ('r02n8_bench_c2','r02n4_bench_c2','r02n8_bench_c3'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', d['e2e'])
    except Exception as e: print(f, 'ERR', e)
"
