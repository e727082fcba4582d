# 8-GPU: NCCL protocol / algorithm / bucket-size variants of the training step (12 CTAs)
O=gpurun_out
run() { name=$1; shift; timeout 300 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --config 3 --no-cpu-baseline --sustain-s 1 --no-e2e $EXTRA > $O/r02n8_bench_c3_$name.json 2> $O/r02n8_$name.err; echo "$name rc=$?"; }
EXTRA="" run base X=1
EXTRA="" run simple NCCL_PROTO=Simple
EXTRA="" run nvls NCCL_ALGO=NVLS
EXTRA="--bucket-mb 32" run bucket32 X=1
EXTRA="--comm-ctas 4" run nvls_ctas4 NCCL_ALGO=NVLS
python -c "
import json
for f in ('base','simple','nvls','bucket32','nvls_ctas4'):
    try:
        d=json.load(open('gpurun_out/r02n8_bench_c3_'+f+'.json')); print(f, round(d['value'],1), round(d['ms_per_step'],3), 'sust', d['sustained'] and round(d['sustained']['ms_per_step'],3))
    except Exception as e: print(f, 'ERR', e)
"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --config 2 --no-cpu-baseline --sustain-s 0 > $O/r02n8_bench_c2.json 2> /dev/null; echo "bench n8 c2 rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02n8_bench_c2.json')); print('c2 n8', round(d['value'],1), d['e2e'])
"
