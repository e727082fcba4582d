# 8-GPU: in-switch (NVLS) all-reduce for the gradient buckets, fewer CTAs
O=gpurun_out
run() { name=$1; shift; timeout 300 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29660 bench.py --gpus 8 --config 3 --no-cpu-baseline --sustain-s 1 --no-e2e $EXTRA > $O/r02n8_bench_c3_$name.json 2> $O/r02n8_$name.err; echo "$name rc=$?"; }
EXTRA="--comm-ctas 12" run base2 X=1
EXTRA="--comm-ctas 12" run nvls12 NCCL_ALGO=allreduce:nvls
EXTRA="--comm-ctas 8" run nvls8 NCCL_ALGO=allreduce:nvls
EXTRA="--comm-ctas 4" run nvls4 NCCL_ALGO=allreduce:nvls
python -c "
import json
for f in ('base2','nvls12','nvls8','nvls4'):
    try:
        d=json.load(open('gpurun_out/r02n8_bench_c3_'+f+'.json')); print(f, round(d['value'],1), round(d['ms_per_step'],3), 'sust', d['sustained'] and round(d['sustained']['ms_per_step'],3))
    except Exception as e: print(f, 'ERR', e)
"
grep -h -i "error" $O/r02n8_nvls12.err | head -3
