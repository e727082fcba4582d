# 8-GPU: NCCL CTA sweep for the training step, data-parallel parity
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 tools/dp_parity.py > $O/r02n8_dp_parity.txt 2>&1; grep dp_parity $O/r02n8_dp_parity.txt
for c in 12 16 24 32; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2964${c:0:1} bench.py --gpus 8 --config 3 --no-cpu-baseline --sustain-s 1 --comm-ctas $c --no-e2e > $O/r02n8_bench_c3_ctas$c.json 2> /dev/null; echo "bench n8 ctas$c rc=$?"
done
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r02n8_bench_c3_ctas*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value'],1), round(d['ms_per_step'],3), 'sust', d['sustained'] and round(d['sustained']['ms_per_step'],3))
    except Exception as e: print(f, 'ERR', e)
"
